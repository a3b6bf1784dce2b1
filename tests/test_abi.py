"""The C-ABI library loads on a GPU-less host and exports every symbol the headers
declare; the compat headers compile as C99 against a caller written for the
reference.  No compute calls (they need a GPU and must FAIL here, not fall back)."""
import ctypes as C
import hashlib
import os
import re
import subprocess
import tempfile

import pytest

from helpers import ROOT
from oblas_b200 import capi


def test_library_exports_reference_api():
    L = capi.lib()
    assert len(capi.REFERENCE_SYMBOLS) == 39  # `nm -g liboblas.a`, SURVEY.md section 2
    for name in capi.REFERENCE_SYMBOLS + capi.EXTENSION_SYMBOLS:
        assert hasattr(L, name), name


def test_headers_and_library_agree():
    """every function prototype in include/*.h is an exported symbol"""
    L = capi.lib()
    names = set()
    for h in ("oblas_b200.h", "oblas_compat.h"):
        src = open(os.path.join(ROOT, "include", h)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        src = "\n".join(l for l in src.splitlines() if not l.lstrip().startswith("#"))
        for m in re.finditer(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\(", src):
            n = m.group(1)
            if n.startswith(("oblas_", "octmat_", "gfmat_", "binmat_", "bf", "oaxpy", "oadd", "oscal", "oswap")):
                names.add(n)
    assert len(names) > 70
    for n in sorted(names):
        assert hasattr(L, n), n


def test_exported_tables_match_golden(golden):
    for name, n in capi.TABLE_SYMBOLS.items():
        b = capi.table(name)
        g = golden["tables"][name]
        assert len(b) == g["len"] == n
        assert hashlib.sha256(b).hexdigest() == g["sha256"], name


def test_make_tables_reproduces_the_reference_tables_and_rejects_bad_polynomials():
    """oblas_b200_make_tables (row f-4, host-only) with the reference's polynomials (tablegen.c:7-12)
    gives the shipped GF(256)/GF(16) tables byte for byte and, for GF(4), the LO table plus the
    corrected HI (= LO << 4; the shipped HI carries the defect of SURVEY 9.2)."""
    L = capi.lib()
    for bits, poly, tag in ((8, 285, "GF2_8"), (4, 19, "GF2_4"), (2, 7, "GF2_2")):
        n = 1 << bits
        lo, hi, inv = (C.c_uint8 * (16 * n))(), (C.c_uint8 * (16 * n))(), (C.c_uint8 * n)()
        assert L.oblas_b200_make_tables(bits, poly, lo, hi, inv) == 0
        assert bytes(lo) == capi.table(tag + "_SHUF_LO")
        assert bytes(inv) == capi.table(tag + "_INV")
        if bits == 2:
            assert bytes(hi) == bytes((b << 4) & 0xFF for b in bytes(lo))
        else:
            assert bytes(hi) == capi.table(tag + "_SHUF_HI")
    lo, hi, inv = (C.c_uint8 * 4096)(), (C.c_uint8 * 4096)(), (C.c_uint8 * 256)()
    assert L.oblas_b200_make_tables(8, 0x11B, lo, hi, inv) == 0          # AES polynomial
    assert inv[0x53] == 0xCA                                             # FIPS-197 example: {53}^-1 = {CA}
    assert L.oblas_b200_make_tables(8, 0x100, lo, hi, inv) != 0          # reducible
    assert L.oblas_b200_make_tables(8, 0x1D, lo, hi, inv) != 0           # wrong degree
    assert L.oblas_b200_make_tables(3, 0xB, lo, hi, inv) != 0            # unsupported width


def test_struct_layouts(golden):
    assert [C.sizeof(capi.Octmat), C.sizeof(capi.Gfmat), C.sizeof(capi.Binmat)] == golden["layout"]["sizeof"]
    assert [capi.Octmat.bits.offset, capi.Gfmat.bits.offset, capi.Binmat.bits.offset] == golden["layout"]["bits_offset"]


def test_host_helpers_without_gpu():
    L = capi.lib()
    assert L.bfx_32(0xABCD1234, 8, 4) == 2
    assert L.bfd_32(0, 4, 4, 0xF) == 0xF0


def test_no_cpu_fallback():
    """without a CUDA device compute entry points must fail loudly"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    L = capi.lib()
    assert L.oblas_b200_init(-1) != 0
    assert b"cuda" in L.oblas_b200_last_error().lower()
    assert L.oblas_b200_rowops(3, 0, 1, -1, None, 16, None, 16, 16, None, None, None, 1) != 0


CALLER = r"""
#include "octmat.h"
#include "gfmat.h"
#include "binmat.h"
#include "oblas.h"
int main(void) {
  octmat *a = octmat_new(4, 100), *b = octmat_new(4, 100);
  om_A(a, 1, 2) = OCT_INV[7];
  oaxpy(a, b, 0, 1, GF_MUL(GF2_8, 3, 5));
  oscal(a, 1, OCT_EXP[OCT_LOG[9]]); oaddrow(a, b, 2, 3); oswaprow(a, 0, 1);
  gfmat *g = gfmat_new(GF2_4, 3, 40); gfmat_set(g, 0, 1, 5); gfmat_axpy(g, g, 1, 0, 2);
  binmat *m = binmat_new(3, 70); binmat_set(m, 0, 3, 1); binmat_add(m, m, 1, 0);
  unsigned r = octmat_solve(a, b, NULL) + gfmat_nnz(g, 0, 0, 40) + binmat_nnz(m, 1, 0, 70);
  uint8_t *p = oblas_alloc(4, 100, 16); oblas_free(p)
  octmat_free(a); octmat_free(b); gfmat_free(g); binmat_free(m);
  return (int)r + (int)sizeof(GF2_8_SHUF_LO) - 4096;
}
"""


def test_reference_style_caller_compiles_and_links():
    """a C99 program written against the reference headers builds against include/compat
    and links with liboblas_b200.so (not run: it needs a GPU)"""
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "caller.c")
        open(src, "w").write(CALLER)
        exe = os.path.join(d, "caller")
        cmd = ["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include", "compat"), src,
               "-L", os.path.join(ROOT, "oblas_b200"), "-loblas_b200", "-Wl,-rpath," + os.path.join(ROOT, "oblas_b200"),
               "-o", exe]
        r = subprocess.run(cmd, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
