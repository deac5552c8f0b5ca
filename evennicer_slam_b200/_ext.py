"""The thin torch C++ extension over the C ABI (``csrc/ens_torch.cpp``): autograd plumbing of the hot calls in C++.

Built in-tree by ``__graft_entry__.build()`` (``build()`` below, ``torch.utils.cpp_extension``) into
``evennicer_slam_b200/_ext/ens_torch_ext.so`` and linked against ``libens_render.so``.  ``module()`` imports the built file;
it never compiles at import time.  ``ENS_TORCH_EXT=0`` selects the pure-Python plumbing of ``functional.py`` instead (the
same kernels through ctypes; bench.py uses it for per-call CUDA-event timing)."""
from __future__ import annotations

import importlib.machinery
import importlib.util
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_DIR = os.path.join(_HERE, "_ext")
_NAME = "ens_torch_ext"
_mod = None
_tried = False
ENABLED = os.environ.get("ENS_TORCH_EXT", "1") != "0"


def build(verbose: bool = False):
    """Compile csrc/ens_torch.cpp (host C++ only; the CUDA kernels live in libens_render.so)."""
    import torch  # noqa: F401
    from torch.utils import cpp_extension
    os.makedirs(_DIR, exist_ok=True)
    return cpp_extension.load(
        name=_NAME, sources=[os.path.join(_HERE, "csrc", "ens_torch.cpp")], build_directory=_DIR, with_cuda=True,
        extra_cflags=["-O2", "-std=c++17"], extra_ldflags=[f"-L{_HERE}", "-lens_render", "-Wl,-rpath,'$$ORIGIN/..'"],
        verbose=verbose)


def module():
    """The built extension module, or None if it has not been built (the Python plumbing is used then)."""
    global _mod, _tried
    if _mod is not None or _tried:
        return _mod
    _tried = True
    path = os.path.join(_DIR, _NAME + ".so")
    if not os.path.exists(path):
        return None
    import torch  # noqa: F401  (libtorch must be loaded first)
    from . import _lib
    _lib.lib()                       # libens_render.so (also found through the extension's rpath)
    loader = importlib.machinery.ExtensionFileLoader(_NAME, path)
    spec = importlib.util.spec_from_loader(_NAME, loader)
    mod = importlib.util.module_from_spec(spec)
    loader.exec_module(mod)
    _mod = mod
    return _mod
