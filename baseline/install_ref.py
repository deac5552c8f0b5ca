"""Install the UNMODIFIED reference modules of the hot path under ``baseline/_ref/`` (git-ignored, travels to the GPU box).

    python baseline/install_ref.py            # in the build container, where /root/reference exists

The reference (cs-vision/EvenNICER-SLAM) is a plain Python tree without ``setup.py`` / ``pyproject.toml``, so ``pip install``
cannot be used; this recipe copies exactly the files the ray-rendering path needs -- nothing is edited:

    src/__init__.py, src/common.py, src/config.py, src/conv_onet/**, src/utils/Renderer.py,
    configs/nice_slam.yaml, configs/Replica/{replica,room0}.yaml

They are used (a) by ``bench.py`` for the ``gpu_reference`` figure (the reference's own eager-PyTorch renderer on the same
B200, same inputs) and (b) by ``tests/test_gpu_reference.py`` (reference-on-CUDA against the reference-on-CPU goldens: the
measured relu-kink sensitivity of the decoder gradients).  ``__graft_entry__.build()`` runs it when /root/reference exists.
"""
import os
import shutil
import sys

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
FILES = ["src/__init__.py", "src/common.py", "src/config.py", "src/conv_onet/__init__.py", "src/conv_onet/config.py",
         "src/conv_onet/models/__init__.py", "src/conv_onet/models/decoder.py", "src/utils/Renderer.py",
         "configs/nice_slam.yaml", "configs/Replica/replica.yaml", "configs/Replica/room0.yaml"]


def install() -> bool:
    if not os.path.isdir(os.path.join(REF, "src")):
        return False
    for rel in FILES:
        src, dst = os.path.join(REF, rel), os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
    return True


if __name__ == "__main__":
    ok = install()
    print("installed %d reference files under %s" % (len(FILES), DST) if ok else "/root/reference not present: nothing done")
    sys.exit(0 if ok else 1)
