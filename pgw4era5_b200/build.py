"""Compile libpgw_b200.so in-tree with nvcc for sm_100a:  python -m pgw4era5_b200.build"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def build(force=False, verbose=False):
    csrc = os.path.join(HERE, "csrc")
    cmd = ["make", "-C", csrc, "-j", str(min(8, os.cpu_count() or 1))]
    if force:
        subprocess.check_call(["make", "-C", csrc, "clean"], stdout=subprocess.DEVNULL)
    out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or out.returncode != 0:
        sys.stderr.write(out.stdout)
    if out.returncode != 0:
        raise RuntimeError("building libpgw_b200.so failed")
    return os.path.join(HERE, "libpgw_b200.so")


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
