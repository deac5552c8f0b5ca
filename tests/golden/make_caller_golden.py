"""Golden outputs of the reference's CALLER loop bodies (tests/golden/caller_loops.py) run on the reference itself (CPU).

    python tests/golden/make_caller_golden.py        # build container: /root/reference must exist

Stores, per iteration: the pixel draws (torch.randint on the CPU generator; the CUDA test replays them, since the CUDA
generator draws a different stream), the loss, the surviving ray counts and every gradient the loop produces.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, HERE]

import ref_harness as rh                    # noqa: E402
import caller_loops as cl                   # noqa: E402
from cases import tiny_scene, SEED          # noqa: E402


def main():
    ref = rh.load()
    scene = tiny_scene()
    model, c, renderer, cfg = rh.build_reference(scene)
    cam = scene.cam
    ns = types.SimpleNamespace(get_samples=ref.common.get_samples, get_camera_from_tensor=ref.common.get_camera_from_tensor,
                               renderer=renderer, H=cam.H, W=cam.W, fx=cam.fx, fy=cam.fy, cx=cam.cx, cy=cam.cy)
    bound = torch.from_numpy(scene.bound.copy())
    frames = cl.caller_inputs()
    draws = []
    orig_randint = torch.randint

    def spy_randint(*a, **k):
        r = orig_randint(*a, **k)
        draws.append(r.clone().numpy())
        return r
    out = {}
    # ---- mapper: 3 keyframes + current frame, 12 px each, BA on, every stage ----
    for stage in ("middle", "fine", "color"):
        torch.manual_seed(SEED)
        draws.clear()
        for p in model.parameters():
            p.grad = None
        cg = {k: v.clone().requires_grad_(True) for k, v in c.items()}
        cams = [torch.from_numpy(f[0].copy()).requires_grad_(True) for f in frames[1:]]
        kfs = [{"depth": torch.from_numpy(f[1]), "color": torch.from_numpy(f[2]),
                "est_c2w": ref.common.get_camera_from_tensor(torch.from_numpy(f[0].copy())).detach()} for f in frames[:3]]
        cur = {"depth": torch.from_numpy(frames[3][1]), "color": torch.from_numpy(frames[3][2])}
        torch.randint = spy_randint
        try:
            loss, n_in, depth, color = cl.mapper_iteration(ns, cg, model, kfs, cams, cur, bound, stage, 12, "cpu")
        finally:
            torch.randint = orig_randint
        tag = "map." + stage
        out[tag + ".loss"] = np.float64(loss); out[tag + ".n_inside"] = np.int64(n_in)
        out[tag + ".depth"] = depth.numpy(); out[tag + ".color"] = color.numpy()
        out[tag + ".draws"] = np.stack(draws)
        for i, ct in enumerate(cams):
            out[f"{tag}.g_cam{i}"] = ct.grad.numpy().copy()
        for k, v in cg.items():
            if v.grad is not None:
                out[f"{tag}.ggrid.{k}"] = v.grad.numpy().copy()
        for lv in ("middle", "fine", "color"):
            for k, p in getattr(model, lv + "_decoder").named_parameters():
                if p.grad is not None and k in ("pts_linears.0.weight", "fc_c.4.weight", "output_linear.weight", "embedder._B"):
                    out[f"{tag}.gdec.{lv}.{k}"] = p.grad.numpy().copy()
        print(tag, "loss", loss, "inside", n_in)
    # ---- tracker: 40 px from the cropped region, pose gradient only ----
    torch.manual_seed(SEED + 1)
    draws.clear()
    ct = torch.from_numpy(frames[3][0].copy()).requires_grad_(True)
    cg = {k: v.clone() for k, v in c.items()}
    torch.randint = spy_randint
    try:
        loss, n_in, n_mask = cl.tracker_iteration(ns, cg, model, ct, torch.from_numpy(frames[3][1]), torch.from_numpy(frames[3][2]),
                                                 bound, 40, 2, 2, "cpu")
    finally:
        torch.randint = orig_randint
    out["track.loss"] = np.float64(loss); out["track.n_inside"] = np.int64(n_in); out["track.n_mask"] = np.int64(n_mask)
    out["track.draws"] = np.stack(draws)
    out["track.g_cam"] = ct.grad.numpy().copy()
    print("track loss", loss, "inside", n_in, "mask", n_mask)
    np.savez_compressed(os.path.join(HERE, "caller_loops.npz"), **out)
    print("caller_loops.npz", os.path.getsize(os.path.join(HERE, "caller_loops.npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()
