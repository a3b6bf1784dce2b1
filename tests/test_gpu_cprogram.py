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


def test_elimination_loop_through_the_op_queue(tmp_path):
    """tests/c/elim_loop.c = the caller-side loop of SURVEY 3.7 (one oaxpy per row and column) built
    against include/compat + liboblas_b200.so: the same digests with one launch + synchronise per call
    (mode 0) and with the op queue (oblas_b200_set_async(2): one launch per batch of independent row
    operations), equal to the reference-linked binary's (tests/golden/elim_loop_256_320.out)."""
    exe = str(tmp_path / "elim_loop_b200")
    libdir = os.path.join(ROOT, "oblas_b200")
    r = subprocess.run(["gcc", "-std=c99", "-O2", "-D_DEFAULT_SOURCE", "-DOBLAS_B200_QUEUE", "-I", os.path.join(ROOT, "include", "compat"),
                        os.path.join(ROOT, "tests", "c", "elim_loop.c"), "-L", libdir, "-loblas_b200",
                        "-Wl,-rpath," + libdir, "-o", exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    want = open(os.path.join(ROOT, "tests", "golden", "elim_loop_256_320.out")).read()
    outs = {}
    for mode in ("0", "2"):
        out = subprocess.run([exe, "256", "320", mode], capture_output=True, text=True, timeout=300)
        assert out.returncode == 0, out.stderr
        assert out.stdout == want, (mode, out.stdout, want)
        outs[mode] = out.stderr
    stats = [ln for ln in outs["2"].splitlines() if ln.startswith("queue launches")][0].split()
    launches, ops = int(stats[2]), int(stats[4])
    assert ops > 100000 and launches < ops // 50, stats      # ~4 launches per column instead of ~512
    ref_exe = os.path.join(ROOT, "oracle", "_ref", "elim_loop_ref")
    if os.path.exists(ref_exe):
        ref = subprocess.run([ref_exe, "256", "320"], capture_output=True, text=True, timeout=300)
        assert ref.stdout == want
