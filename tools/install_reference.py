"""Install the UNMODIFIED reference sources of the hot path under baseline/_ref/ (git-ignored; travels to the GPU
box with gpurun) so that `bench.py --impl reference` and `cpu_baseline` time the real reference:

    models/__init__.py, models/common.py, models/actnorm.py,
    models/networks_iid_hlgvit_crs_gd4_cfs_v3.py     (define_G :93-100, dec_ipt.forward :392-1020)
    models/networks_iid_hlgvit_crs_gd4_cfs.py        (sibling generator)

The reference is a plain Python source tree without setup.py / pyproject.toml, so "install" is a byte-for-byte copy
of the files the path imports (verified by sha256 below).  Nothing is copied into tracked paths.
"""
import hashlib
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FILES = ("models/__init__.py", "models/common.py", "models/actnorm.py",
         "models/networks_iid_hlgvit_crs_gd4_cfs_v3.py", "models/networks_iid_hlgvit_crs_gd4_cfs.py")


def _sha(path):
    return hashlib.sha256(open(path, "rb").read()).hexdigest()


def install(src="/root/reference", dst=None, verbose=False):
    """Returns the list of installed files, [] when the reference checkout is absent (GPU box)."""
    dst = dst or os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(src):
        return []
    done = []
    for rel in FILES:
        a, b = os.path.join(src, rel), os.path.join(dst, rel)
        if not os.path.exists(a):
            continue
        os.makedirs(os.path.dirname(b), exist_ok=True)
        if not os.path.exists(b) or _sha(a) != _sha(b):
            shutil.copyfile(a, b)
        assert _sha(a) == _sha(b)
        done.append(rel)
        if verbose:
            print("baseline/_ref/%s  sha256 %s" % (rel, _sha(b)[:16]))
    return done


if __name__ == "__main__":
    print("\n".join(install(verbose=True)) or "reference checkout not found", file=sys.stderr)
