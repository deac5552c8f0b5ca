"""CPU checks of the boundary: the library builds/loads and exports every symbol the header declares;
host-side helpers behave.  No compute calls (no GPU here)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as ge
    ge.build()
    from evennicer_slam_b200 import _lib
    return _lib


def test_header_symbols_exported(built):
    hdr = open(os.path.join(ROOT, "include", "ens_render.h")).read()
    declared = set(re.findall(r"\b(ens_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 15
    handle = ctypes.CDLL(built.LIB_PATH)
    for name in declared:
        assert hasattr(handle, name), f"{name} declared in ens_render.h but not exported"
    assert declared == set(built.EXPORTED_SYMBOLS), declared ^ set(built.EXPORTED_SYMBOLS)


def test_sizes_match_reference_parameter_counts(built):
    L = built.lib()
    # middle 15 800, fine 20 920, colour 15 899, coarse 6 337 parameters
    assert [L.ens_decoder_grad_floats(i) for i in range(4)] == [6337, 15800, 20920, 15899]
    assert [L.ens_decoder_num_tensors(i) for i in range(4)] == [12, 23, 23, 23]
    assert L.ens_packed_decoder_floats(1) % 4 == 0 and L.ens_packed_decoder_floats(2) % 4 == 0
    # pose-only backward: optional scratch of the tcgen05 backward = per point g_raw/g_c rows (24 + 8 + 16 + 72 B) + four
    # raw weight-gradient accumulator sets of 17 152 floats + 256 B of counters
    tc = 48000 * (24 + 8 + 16 + 72) + 4 * 17152 * 4 + 512
    assert L.ens_bwd_workspace_bytes(1000, 48, 0) == tc
    # 1000 rays x 48 samples = 1500 tiles of 32 points: split-backward scratch = g_h tiles (3 decoders x 5 x 4 KB per
    # tile) + mask words (3 x 5 x 128 B) + points (1 KB); it exceeds the recompute / FFMA activation scratch
    # (ENS_BWD_TC_SPLIT=1 adds the g_u rows of the split tcgen05 backward: [3 decoders][375 tiles][5 blocks][128][32] f32)
    assert L.ens_bwd_workspace_bytes(1000, 48, 1) == max(tc, 1500 * (3 * 5 * 4096 + 3 * 5 * 128 + 1024))
    assert L.ens_bwd_workspace_bytes(1000, 48, 1) >= (48000 + 1) * 160 * 4
    # saved-for-backward buffer of the forward: masks only / masks + five activation tiles per decoder
    assert L.ens_fwd_saved_bytes(1000, 48, 3, 0) == 3 * 1500 * 5 * 128
    assert L.ens_fwd_saved_bytes(1000, 48, 3, 1) == 3 * 1500 * 5 * 128 + 3 * 1500 * 5 * 4096
    assert L.ens_fwd_saved_bytes(1000, 48, 0, 1) == 0            # coarse stage: FFMA kernels, nothing saved
    assert L.ens_fwd_saved_bytes(1000, 40, 3, 1) == 0            # 40 samples per ray do not tile the CTAs
    # tcgen05-path scratch: points f64[3] + z f64 + raw f32[4] + one f32[4] output plane per decoder, per sample point
    assert L.ens_fwd_scratch_bytes(1000, 48, 3) == 48000 * 96 and L.ens_fwd_scratch_bytes(1000, 48, 0) == 0
    import evennicer_slam_b200.synthetic as syn
    for li, lv in enumerate(syn.LEVELS):
        n = sum(int(np.prod(s)) for _, s in syn.decoder_param_shapes(lv))
        assert n == L.ens_decoder_grad_floats(li)


def test_error_codes_without_gpu(built):
    L = built.lib()
    assert L.ens_strerror(0) == b"ok"
    assert L.ens_grid_to_native(None, None, 1, None) == -1          # ENS_EINVAL before any CUDA call
    assert L.ens_eval_points(None, 0, None, 1, 0, 1, None, None) == -1
    assert L.ens_pack_decoder(7, None, 0, None, None) == -1


def test_dropin_state_dict_keys_match_reference_layout():
    import torch
    from evennicer_slam_b200.decoder import NICE
    import evennicer_slam_b200.synthetic as syn
    m = NICE(coarse=True)
    for lv in syn.LEVELS:
        dec = getattr(m, lv + "_decoder")
        got = [(k, tuple(v.shape)) for k, v in dec.state_dict().items()]
        assert got == syn.decoder_param_shapes(lv), lv
    import copy
    m2 = copy.deepcopy(m)
    assert m2._ens_cache is None and len(list(m2.parameters())) == len(list(m.parameters()))


def test_render_path_fails_loudly_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    import cases
    from evennicer_slam_b200 import harness
    scene = cases.tiny_scene()
    decoders, c, renderer, cfg = harness.build(scene, "cpu")
    with pytest.raises((RuntimeError, ValueError)):
        renderer.render_batch_ray(c, decoders, torch.zeros(4, 3), torch.zeros(4, 3), "cpu", "color")


def test_host_wrappers_refuse_cpu_tensors():
    """No CPU fallback anywhere on the product path: the optimiser / loss wrappers raise on CPU tensors instead of
    computing something in torch."""
    import pytest
    import torch
    from evennicer_slam_b200 import losses, optim
    g = torch.zeros(1, 32, 2, 3, 4)
    with pytest.raises(ValueError):
        optim.FrustumGridAdam({"grid_fine": g}, None)
    with pytest.raises(ValueError):
        optim.FusedAdam([torch.zeros(3, requires_grad=True)])
    with pytest.raises(RuntimeError):
        losses.event_loss(torch.zeros(20, 30, 2), torch.zeros(20, 30, 2))
    with pytest.raises(RuntimeError):
        losses.mapper_loss(torch.zeros(4), torch.zeros(4, 3), torch.zeros(4, dtype=torch.float64), torch.zeros(4, 3))
    with pytest.raises(RuntimeError):
        losses.tracker_loss(torch.zeros(4), torch.zeros(4, 3), torch.zeros(4, dtype=torch.float64),
                            torch.zeros(4, dtype=torch.float64), torch.zeros(4, 3))


def test_gaussian_kernel_matches_torchvision_default_sigma():
    import numpy as np
    import torch
    from torchvision.transforms import _functional_tensor as FT
    from evennicer_slam_b200.losses import gaussian_kernel1d
    for ks in (3, 9, 15):
        want = FT._get_gaussian_kernel1d(ks, ks * 0.15 + 0.35, torch.float32, torch.device("cpu")).numpy()
        assert np.array_equal(gaussian_kernel1d(ks), want)


def test_decoder_tensors_cache_follows_the_live_parameters():
    """scene.decoder_tensors walks the module tree once per decoder object and then reads the live _parameters dicts:
    state_dict order, and re-assigned / loaded / copied parameters are still the ones returned."""
    import copy
    import torch
    from evennicer_slam_b200.decoder import NICE
    from evennicer_slam_b200.scene import decoder_tensors, _bound_floats
    m = NICE(coarse=True)
    for lv in ("coarse", "middle", "fine", "color"):
        dec = getattr(m, lv + "_decoder")
        want = [p for _, p in dec.named_parameters()]
        for _ in range(2):                                   # second call: cached slots
            got = decoder_tensors(m, lv)
            assert len(got) == len(want) and all(a is b for a, b in zip(got, want))
    m.load_state_dict(copy.deepcopy(m.state_dict()))         # in-place copy: same objects
    assert all(a is b for a, b in zip(decoder_tensors(m, "fine"), [p for _, p in m.fine_decoder.named_parameters()]))
    new = torch.nn.Parameter(torch.zeros_like(m.fine_decoder.output_linear.bias))
    m.fine_decoder.output_linear.bias = new                  # a re-assigned Parameter is picked up
    assert decoder_tensors(m, "fine")[-1] is new
    m2 = copy.deepcopy(m)                                    # a copy has its own parameters
    assert all(a is b for a, b in zip(decoder_tensors(m2, "color"), [p for _, p in m2.color_decoder.named_parameters()]))
    assert not any(a is b for a, b in zip(decoder_tensors(m2, "color"), decoder_tensors(m, "color")))
    b = torch.tensor([[-1.0, 2.0], [-3.0, 4.0], [-5.0, 6.5]], dtype=torch.float64)
    assert _bound_floats(b) == (-1.0, 2.0, -3.0, 4.0, -5.0, 6.5)
    b[2, 1] = 7.0                                            # in-place edit bumps the version
    assert _bound_floats(b)[-1] == 7.0


def test_round2_dropins_have_no_cpu_fallback():
    """mapper_ops / event_net must refuse CPU tensors loudly instead of computing something else."""
    import torch
    from evennicer_slam_b200 import event_net, mapper_ops
    with pytest.raises(RuntimeError, match="CUDA"):
        mapper_ops.FrustumSelector(8, 8, 5.0, 5.0, 3.5, 3.5, [[-1, 1], [-1, 1], [-1, 1]], "cpu")
    with pytest.raises(RuntimeError, match="CUDA"):
        event_net.assemble_input(torch.zeros(4, 4, 3), torch.zeros(4, 4, 3))

