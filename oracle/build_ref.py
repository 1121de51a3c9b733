"""Build recipe for ``oracle/_ref/`` — the REAL reference's only native code, compiled where it lies.

TEST INFRASTRUCTURE ONLY.  Nothing under ``tuotuo_b200/`` may import this.

What it does
------------
Compiles ``/root/reference/tuotuo/cutils.pyx`` (the reference's Cython Dirichlet-expectation / AS-103
digamma, cutils.pyx:11-93) into ``oracle/_ref/cutils<EXT_SUFFIX>`` with the image's Cython + gcc:

    cython -3 -o oracle/_ref/cutils.c /root/reference/tuotuo/cutils.pyx
    gcc -O2 -fPIC -shared -I<python> -I<numpy> oracle/_ref/cutils.c -o oracle/_ref/cutils.<abi>.so

and STAGES the reference's pure-Python modules (``tuotuo/{__init__,lda_model,document,utils,text_pre_processor,
generator}.py``, byte for byte, unmodified) next to it as ``oracle/_ref/tuotuo/`` — the equivalent of the
``pip install --target baseline/_ref`` of a Python reference — so that ``bench.py --impl reference`` can time the
REAL reference's ``fit`` path on the GPU box's host cores (``/root/reference`` does not exist there).

``oracle/_ref/`` is git-ignored build output (it travels to the GPU box via gpurun, like the built ``.so`` files);
nothing from the reference enters the repository's history.  The reference's own setup.py is NOT run.

When ``/root/reference`` is absent (the GPU box) this is a no-op: the prebuilt / staged files are used as they are.
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


REF_PY = ("__init__.py", "lda_model.py", "document.py", "utils.py", "text_pre_processor.py", "generator.py")


def stage_ref_python() -> str | None:
    """Copy the reference's .py modules (unmodified) into oracle/_ref/tuotuo/.  Returns that directory, or None."""
    src = os.path.dirname(REF_PYX)
    dst = os.path.join(REF_DIR, "tuotuo")
    if not os.path.isdir(src):
        return dst if os.path.exists(os.path.join(dst, "lda_model.py")) else None
    import shutil

    os.makedirs(dst, exist_ok=True)
    for f in REF_PY:
        s, d = os.path.join(src, f), os.path.join(dst, f)
        if os.path.exists(s) and (not os.path.exists(d) or os.path.getmtime(d) < os.path.getmtime(s)):
            shutil.copyfile(s, d)
    return dst


def build_ref(force: bool = False, verbose: bool = False) -> str | None:
    """Compile the reference cutils into oracle/_ref/ and stage its .py modules.  Returns the .so path or None."""
    so = ref_so_path()
    stage_ref_python()
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
