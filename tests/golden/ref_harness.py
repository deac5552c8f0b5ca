"""Import the (read-only) Python reference: from /root/reference in the build container, else from the copy
``baseline/install_ref.py`` put under ``baseline/_ref/`` (git-ignored; it travels to the GPU box).

Used by ``make_golden.py`` (CPU goldens), by ``tests/test_gpu_reference.py`` (the reference on CUDA against those
goldens) and by ``bench.py``'s ``gpu_reference`` figure.  Never by the product path.

The reference cannot run on CPU unmodified: two lines build a
``cuda:-1`` device string.  They are patched in memory at import; the files under
/root/reference are never touched:
  * src/conv_onet/models/decoder.py:316  ``f'cuda:{p.get_device()}'`` -> ``p.device``
  * src/common.py:202                    ``.to(quad.get_device())``  -> ``.to(quad.device)``
"""
from __future__ import annotations

import importlib
import importlib.util
import os
import sys
import types

import numpy as np
import torch

_LOCAL = os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "baseline", "_ref")
REF_ROOT = "/root/reference" if os.path.isdir("/root/reference/src") else _LOCAL


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "src", "utils", "Renderer.py"))


def _load_patched(modname: str, relpath: str, replacements):
    path = os.path.join(REF_ROOT, relpath)
    with open(path, "r") as f:
        src = f.read()
    for old, new in replacements:
        assert old in src, f"shim anchor not found in {relpath}: {old!r}"
        src = src.replace(old, new)
    spec = importlib.util.spec_from_loader(modname, loader=None, origin=path)
    mod = importlib.util.module_from_spec(spec)
    mod.__file__ = path
    sys.modules[modname] = mod
    exec(compile(src, path, "exec"), mod.__dict__)
    return mod


_loaded = {}


def load():
    """Returns a namespace with the reference's hot-path modules (patched for CPU)."""
    if _loaded:
        return types.SimpleNamespace(**_loaded)
    if not available():
        raise RuntimeError("the reference is not present (neither /root/reference nor baseline/_ref)")
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import src  # noqa: F401  (the reference's package root)
    common = _load_patched("src.common", "src/common.py",
                           [(".to(quad.get_device())", ".to(quad.device)")])
    sys.modules["src"].common = common
    import src.conv_onet.models  # noqa: F401  -- triggers decoder import via package __init__
    # re-load decoder with the device shim (package import above used the unpatched one)
    decoder = _load_patched("src.conv_onet.models.decoder", "src/conv_onet/models/decoder.py",
                            [("device = f'cuda:{p.get_device()}'", "device = p.device")])
    import src.conv_onet.models as models_pkg
    models_pkg.decoder = decoder
    models_pkg.decoder_dict = {"nice": decoder.NICE, "imap": decoder.MLP}
    import src.conv_onet.config as onet_config
    import src.config as config
    renderer = importlib.import_module("src.utils.Renderer")
    _loaded.update(common=common, decoder=decoder, onet_config=onet_config, config=config,
                   Renderer=renderer.Renderer)
    return types.SimpleNamespace(**_loaded)


def load_cfg(rel="configs/Replica/room0.yaml"):
    ref = load()
    cwd = os.getcwd()
    os.chdir(REF_ROOT)
    try:
        cfg = ref.config.load_config(rel, "configs/nice_slam.yaml")
    finally:
        os.chdir(cwd)
    return cfg


def build_reference(scene, cfg=None, device="cpu"):
    """Reference NICE decoders + grids + Renderer for a synthetic ``Scene`` (tensors on ``device``)."""
    ref = load()
    cfg = cfg or load_cfg()
    model = ref.onet_config.get_model(cfg, nice=True)
    bound = torch.from_numpy(scene.bound.copy()).to(device)
    model.bound = bound
    model.middle_decoder.bound = bound
    model.fine_decoder.bound = bound
    model.color_decoder.bound = bound
    model.coarse_decoder.bound = bound * cfg["model"]["coarse_bound_enlarge"]
    for lv in ("coarse", "middle", "fine", "color"):
        dec = getattr(model, lv + "_decoder")
        sd = {k: torch.from_numpy(v.copy()) for k, v in scene.decoders[lv].items()}
        dec.load_state_dict(sd)
    model = model.to(device)
    c = {k: torch.from_numpy(v.copy()).to(device) for k, v in scene.grids.items()}
    cam = scene.cam
    slam = types.SimpleNamespace(nice=True, bound=bound, H=cam.H, W=cam.W, fx=cam.fx, fy=cam.fy,
                                 cx=cam.cx, cy=cam.cy)
    renderer = ref.Renderer(cfg, None, slam)
    return model, c, renderer, cfg
