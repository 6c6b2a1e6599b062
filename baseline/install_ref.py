"""Copy the UNMODIFIED reference packages the GPU baseline needs into the git-ignored ``baseline/_ref/`` so that they
travel to the GPU box with the repo snapshot (``/root/reference`` does not exist there).  Nothing under ``baseline/_ref``
is committed, edited or imported by the product path (``editor_b200/``): it is the measured denominator only
(``bench.py``'s ``torch_eager_gpu`` / ``do_train_dropin`` keys, ``tests/test_dropin_do_train_gpu.py``).

The reference has no ``setup.py`` / ``pyproject.toml``, so ``pip install --target baseline/_ref
/root/reference`` has nothing to install; a verbatim directory copy of its eight packages is the equivalent.
"""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DST = os.path.join(ROOT, "baseline", "_ref")
PACKAGES = ("config", "configs", "engine", "layers", "modeling", "solver", "utils", "pytorch_wavelets")


def install(src="/root/reference", dst=DST, quiet=False):
    if not os.path.isdir(os.path.join(src, "modeling")):
        if not quiet:
            print("reference tree not present at %s: baseline/_ref left as it is" % src)
        return os.path.isdir(os.path.join(dst, "modeling"))
    os.makedirs(dst, exist_ok=True)
    for name in PACKAGES:
        d = os.path.join(dst, name)
        if os.path.isdir(d):
            shutil.rmtree(d)
        shutil.copytree(os.path.join(src, name), d,
                        ignore=shutil.ignore_patterns("__pycache__", "*.pyc", "basic_cnn_params", "*.pth"))
    if not quiet:
        print("reference packages copied to", dst)
    return True


if __name__ == "__main__":
    sys.exit(0 if install(*(sys.argv[1:2])) else 1)
