"""Build recipe for ``oracle/_ref/`` — the REAL reference's only native code, compiled where it lies.

TEST INFRASTRUCTURE ONLY.  Nothing under ``tuotuo_b200/`` may import this.

What it does
------------
Compiles ``/root/reference/tuotuo/cutils.pyx`` (the reference's Cython Dirichlet-expectation / AS-103
digamma, cutils.pyx:11-93) into ``oracle/_ref/cutils<EXT_SUFFIX>`` with the image's Cython + gcc:

    cython -3 -o oracle/_ref/cutils.c /root/reference/tuotuo/cutils.pyx
    gcc -O2 -fPIC -shared -I<python> -I<numpy> oracle/_ref/cutils.c -o oracle/_ref/cutils.<abi>.so

Only *outputs* are written (``oracle/_ref/`` is git-ignored, travels to the GPU box via gpurun);
no reference source is copied into the repository.  The reference's own setup.py is NOT run.

When ``/root/reference`` is absent (the GPU box) this is a no-op: the prebuilt ``.so`` is used as is.
"""
from __future__ import annotations

import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
REF_PYX = "/root/reference/tuotuo/cutils.pyx"


def ref_so_path() -> str:
    return os.path.join(REF_DIR, "cutils" + sysconfig.get_config_var("EXT_SUFFIX"))


def build_ref(force: bool = False, verbose: bool = False) -> str | None:
    """Compile the reference cutils into oracle/_ref/.  Returns the .so path or None if unavailable."""
    so = ref_so_path()
    if not os.path.exists(REF_PYX):
        return so if os.path.exists(so) else None
    if os.path.exists(so) and not force and os.path.getmtime(so) >= os.path.getmtime(REF_PYX):
        return so
    import numpy

    os.makedirs(REF_DIR, exist_ok=True)
    c_file = os.path.join(REF_DIR, "cutils.c")
    out = None if verbose else subprocess.DEVNULL
    subprocess.check_call(
        [sys.executable, "-m", "cython", "-3", "-o", c_file, REF_PYX], stdout=out, stderr=out
    )
    subprocess.check_call(
        [
            "gcc", "-O2", "-fPIC", "-shared", "-fno-fast-math", "-ffp-contract=off",
            "-DNPY_NO_DEPRECATED_API=NPY_1_7_API_VERSION",
            "-I" + sysconfig.get_paths()["include"],
            "-I" + numpy.get_include(),
            c_file, "-o", so, "-lm",
        ],
        stdout=out, stderr=out,
    )
    return so


def load_ref_cutils():
    """Import the compiled reference module (as top-level ``cutils``) or return None."""
    so = ref_so_path()
    if not os.path.exists(so):
        return None
    import importlib.util

    spec = importlib.util.spec_from_file_location("cutils", so)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    p = build_ref(force="--force" in sys.argv, verbose=True)
    print("oracle/_ref:", p)
