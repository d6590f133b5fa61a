"""Copy the reference's Python sources to baseline/_ref (git-ignored, travels to the GPU box with gpurun) so the
full-model end-to-end parity test (tests/test_gpu_model_e2e.py) can run the UNMODIFIED reference model next to the
patched one on a B200.  TEST INFRASTRUCTURE; run in the build container:  python -m oracle.install_ref
(The reference has no packaging metadata, so `pip install --target baseline/_ref /root/reference` is not possible.)"""
import os
import shutil
import sys

SRC = "/root/reference"
DST = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref")


def main():
    if not os.path.isdir(SRC):
        sys.exit("no reference tree at /root/reference")
    os.makedirs(DST, exist_ok=True)
    n = 0
    for root, dirs, files in os.walk(SRC):
        dirs[:] = [d for d in dirs if d not in ("__pycache__", ".git", "imgs")]
        for f in files:
            if f.endswith(".py"):
                rel = os.path.relpath(os.path.join(root, f), SRC)
                os.makedirs(os.path.dirname(os.path.join(DST, rel)) or DST, exist_ok=True)
                shutil.copyfile(os.path.join(root, f), os.path.join(DST, rel))
                n += 1
    print(f"copied {n} files to {DST}")


if __name__ == "__main__":
    main()
