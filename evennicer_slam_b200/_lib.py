"""ctypes binding of ``libens_render.so`` (the C ABI declared in ``include/ens_render.h``).

There is no fallback: if the shared library is missing or fails to load, importing a
compute entry point raises.  The library is built in-tree by ``__graft_entry__.build()``
(``make -C evennicer_slam_b200/csrc``).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libens_render.so")

STAGES = {"coarse": 0, "middle": 1, "fine": 2, "color": 3}
LEVELS = ("coarse", "middle", "fine", "color")
# levels (grids + decoders) a stage touches -- decoder.py:312-342
STAGE_LEVELS = {"coarse": ("coarse",), "middle": ("middle",), "fine": ("middle", "fine"),
                "color": ("middle", "fine", "color")}

ENS_OK = 0
ABI_VERSION = 6            # include/ens_render.h: ENS_ABI_VERSION


class EnsScene(C.Structure):
    _fields_ = [("grid", C.c_void_p * 4),
                ("dims", (C.c_int32 * 3) * 4),
                ("bound", (C.c_double * 2) * 3),
                ("coarse_bound", (C.c_double * 2) * 3),
                ("weights", C.c_void_p * 4)]


class EnsRenderCfg(C.Structure):
    _fields_ = [("n_samples", C.c_int32), ("n_surface", C.c_int32), ("n_importance", C.c_int32),
                ("lindisp", C.c_int32), ("perturb", C.c_float), ("occupancy", C.c_int32),
                ("t_vals", C.c_void_p), ("t_vals_surface", C.c_void_p)]


class EnsAdamLevel(C.Structure):
    _fields_ = [("grid", C.c_void_p), ("grad", C.c_void_p), ("exp_avg", C.c_void_p), ("exp_avg_sq", C.c_void_p),
                ("voxel_index", C.c_void_p), ("n_selected", C.c_int64), ("n_voxels", C.c_int64), ("lr", C.c_double)]


class EnsGrads(C.Structure):
    _fields_ = [("grid", C.c_void_p * 4), ("decoder", C.c_void_p * 4),
                ("rays_o", C.c_void_p), ("rays_d", C.c_void_p)]


_SIGNATURES = {
    "ens_version": (C.c_int, []),
    "ens_strerror": (C.c_char_p, [C.c_int]),
    "ens_frustum_workspace_bytes": (C.c_int64, [C.c_int, C.c_int, C.c_int]),
    "ens_frustum_mask": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                   C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "ens_keyframe_overlap": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_float, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "ens_unet_input": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                 C.c_void_p, C.c_void_p]),
    "ens_unet_input_bwd": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "ens_grid_touched": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "ens_grid_compact": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int, C.c_void_p,
                                   C.c_void_p]),
    "ens_last_error": (C.c_char_p, []),
    "ens_packed_decoder_floats": (C.c_int64, [C.c_int]),
    "ens_decoder_grad_floats": (C.c_int64, [C.c_int]),
    "ens_decoder_num_tensors": (C.c_int, [C.c_int]),
    "ens_bwd_workspace_bytes": (C.c_int64, [C.c_int64, C.c_int, C.c_int]),
    "ens_fwd_saved_bytes": (C.c_int64, [C.c_int64, C.c_int, C.c_int, C.c_int]),
    "ens_fwd_saved_bytes_kind": (C.c_int64, [C.c_int64, C.c_int, C.c_int, C.c_int]),
    "ens_fwd_scratch_bytes": (C.c_int64, [C.c_int64, C.c_int, C.c_int]),
    "ens_grid_to_native": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "ens_grid_from_native": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p]),
    "ens_pack_decoder": (C.c_int, [C.c_int, C.POINTER(C.c_void_p), C.c_int, C.c_void_p, C.c_void_p]),
    "ens_depth_max": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "ens_sample_rays": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                  C.c_float, C.c_float, C.c_float, C.c_float, C.c_void_p, C.c_int,
                                  C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ens_lattice_rays": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_float,
                                   C.c_float, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ens_rays_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_float, C.c_float, C.c_float,
                               C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ens_pose_fwd": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "ens_pose_bwd": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ens_eval_points": (C.c_int, [C.POINTER(EnsScene), C.c_int, C.c_void_p, C.c_int, C.c_int64, C.c_int,
                                  C.c_void_p, C.c_void_p]),
    "ens_render_fwd": (C.c_int, [C.POINTER(EnsScene), C.POINTER(EnsRenderCfg), C.c_int, C.c_void_p, C.c_void_p,
                                 C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                 C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p,
                                 C.c_int64, C.c_void_p]),
    "ens_render_bwd": (C.c_int, [C.POINTER(EnsScene), C.POINTER(EnsRenderCfg), C.c_int, C.c_void_p, C.c_void_p,
                                 C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                 C.c_void_p, C.POINTER(EnsGrads), C.c_void_p, C.c_int64, C.c_void_p, C.c_int64,
                                 C.c_int, C.c_void_p]),
    "ens_grid_adam_step": (C.c_int, [C.POINTER(EnsAdamLevel), C.c_int, C.c_double, C.c_double, C.c_double, C.c_int64,
                                     C.c_void_p, C.c_int, C.c_void_p]),
    "ens_tensors_adam_step": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_int64),
                                        C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_double), C.c_int, C.c_void_p,
                                        C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_int64, C.c_void_p,
                                        C.c_void_p]),
    "ens_event_loss": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int),
                                 C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_int, C.c_float, C.c_void_p,
                                 C.c_void_p, C.c_void_p]),
    "ens_rgbd_loss": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int64,
                                C.c_int, C.c_double, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                C.c_void_p]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib = None


def lib():
    """The loaded library (loads on first use).  Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(make -C evennicer_slam_b200/csrc). There is no CPU or PyTorch fallback for the render path.")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(handle, name)      # AttributeError here = header/library mismatch
            fn.restype = res
            fn.argtypes = args
        if handle.ens_version() != ABI_VERSION:
            raise RuntimeError("libens_render.so ABI version mismatch")
        _lib = handle
    return _lib


def check(rc: int, what: str = ""):
    if rc != ENS_OK:
        msg = lib().ens_strerror(rc).decode()
        if rc == -3:
            msg += ": " + lib().ens_last_error().decode()
        raise RuntimeError(f"libens_render: {what} failed: {msg} (code {rc})")


def ptr(t):
    """Device pointer of a tensor (or None)."""
    return None if t is None else C.c_void_p(t.data_ptr())


def cur_stream(device):
    """Current stream of ``device`` for an ABI call.  The library launches on the calling thread's CURRENT device (one
    process per GPU, ``torch.cuda.set_device(LOCAL_RANK)``): tensors on another device would fail inside the launch with an
    opaque 'invalid resource handle', so say it here."""
    import torch
    device = torch.device(device)
    cur = torch.cuda.current_device()
    if device.index is not None and device.index != cur:
        raise RuntimeError(f"libens_render: tensors live on cuda:{device.index} but the current device is cuda:{cur}; "
                           f"call torch.cuda.set_device({device.index}) or wrap the call in torch.cuda.device({device.index})")
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)
