"""Golden vectors for the frustum feature selection and the keyframe overlap test, from the reference ITSELF.

    python tests/golden/make_frustum_golden.py          # in the build container (needs /root/reference and cv2)

``Mapper.get_mask_from_c2w`` (src/Mapper.py:115-186) and the per-keyframe projection of ``keyframe_selection_overlap``
(:188-250) are taken from the reference's source file by name (``ast``), compiled unmodified and run on seeded synthetic
inputs; src/Mapper.py as a whole cannot be imported here (colorama, the dataset modules).  The overlap function is run
with ``np.random.permutation`` and ``sorted`` observed: the golden holds each keyframe's ``percent_inside``.
Writes tests/golden/frustum.npz (masks bit-packed).
"""
import ast
import os
import sys
import types

import cv2
import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, HERE)

import frustum_cases as fc  # noqa: E402


def reference_functions():
    src = open("/root/reference/src/Mapper.py").read()
    tree = ast.parse(src)
    cls = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "Mapper"][0]
    want = {"get_mask_from_c2w", "keyframe_selection_overlap"}
    fns = [n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name in want]
    mod = ast.Module(body=fns, type_ignores=[])
    import ref_harness
    ref = ref_harness.load()
    ns = {"np": np, "cv2": cv2, "torch": torch, "get_samples": ref.common.get_samples}
    exec(compile(mod, "/root/reference/src/Mapper.py", "exec"), ns)
    return ns["get_mask_from_c2w"], ns["keyframe_selection_overlap"]


def main():
    get_mask, kf_overlap = reference_functions()
    out = {}
    for name, case in fc.mask_cases().items():
        slf = types.SimpleNamespace(H=case["cam"][0], W=case["cam"][1], fx=case["cam"][2], fy=case["cam"][3],
                                    cx=case["cam"][4], cy=case["cam"][5], bound=torch.from_numpy(case["bound"]))
        for key, shape in case["shapes"].items():
            m = get_mask(slf, torch.from_numpy(case["c2w"]), key, list(shape), case["depth"])
            assert m.shape == (shape[2], shape[1], shape[0])
            out[f"{name}.{key}.mask"] = np.packbits(m.reshape(-1))
            out[f"{name}.{key}.count"] = np.int64(m.sum())
            print(name, key, shape, int(m.sum()), m.size)
    for name, case in fc.overlap_cases().items():
        cam = case["cam"]
        slf = types.SimpleNamespace(H=cam[0], W=cam[1], fx=cam[2], fy=cam[3], cx=cam[4], cy=cam[5], device="cpu")
        seen = []
        real_sorted = sorted

        def spy_sorted(lst, **kw):
            seen.append([(d["id"], float(d["percent_inside"])) for d in lst])
            return real_sorted(lst, **kw)
        kf_overlap.__globals__["sorted"] = spy_sorted
        torch.manual_seed(case["seed"]); np.random.seed(case["seed"])
        kfd = [{"est_c2w": torch.from_numpy(c)} for c in case["kf_c2w"]]
        sel = kf_overlap(slf, torch.from_numpy(case["color"]), torch.from_numpy(case["depth"]),
                         torch.from_numpy(case["c2w"]), kfd, case["k"], 16, case["pixels"])
        out[f"{name}.percent_inside"] = np.array([p for _, p in seen[0]], np.float64)
        out[f"{name}.selected"] = np.array(sel, np.int64)
        print(name, out[f"{name}.percent_inside"], sel)
    np.savez_compressed(os.path.join(HERE, "frustum.npz"), **out)


if __name__ == "__main__":
    main()
