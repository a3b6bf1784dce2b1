"""ONE C program written against the reference headers (tests/c/dropin_demo.c): its output
when linked with the unmodified reference (tests/golden/dropin_demo.out, produced by
oracle/Makefile in the build container) must equal its output when recompiled against
include/compat and linked with liboblas_b200.so on the B200."""
import os
import subprocess

import pytest

from helpers import ROOT

pytestmark = pytest.mark.gpu


def test_same_c_program_same_output(tmp_path):
    exe = str(tmp_path / "dropin_demo_b200")
    libdir = os.path.join(ROOT, "oblas_b200")
    r = subprocess.run(["gcc", "-std=c99", "-O2", "-I", os.path.join(ROOT, "include", "compat"),
                        os.path.join(ROOT, "tests", "c", "dropin_demo.c"), "-L", libdir, "-loblas_b200",
                        "-Wl,-rpath," + libdir, "-o", exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    want = open(os.path.join(ROOT, "tests", "golden", "dropin_demo.out")).read()
    assert out.stdout == want
    ref_exe = os.path.join(ROOT, "oracle", "_ref", "dropin_demo_ref")
    if os.path.exists(ref_exe):  # the reference-linked binary travels with the snapshot
        ref = subprocess.run([ref_exe], capture_output=True, text=True, timeout=300)
        assert ref.stdout == out.stdout
