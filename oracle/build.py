"""TEST INFRASTRUCTURE ONLY -- builds the oracle's compiled column loops.

    python -m oracle.build

gcc -O2 (no -ffast-math: float64 semantics must match numba's default).
Output: oracle/_build/liboracle_kernels.so (git-ignored, travels with gpurun).
The reference itself is pure Python (no C sources), so there is no
``oracle/_ref`` binary to compile; see DESIGN.md.
"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))


def build(force=False):
    src = os.path.join(HERE, "oracle_kernels.c")
    out_dir = os.path.join(HERE, "_build")
    out = os.path.join(out_dir, "liboracle_kernels.so")
    os.makedirs(out_dir, exist_ok=True)
    if (not force and os.path.exists(out)
            and os.path.getmtime(out) >= os.path.getmtime(src)):
        return out
    subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-o", out, src, "-lm"])
    return out


if __name__ == "__main__":
    print(build(force=True))
