"""Generate tests/golden/golden.json by RUNNING THE UNMODIFIED REFERENCE
(oracle/_ref/liboblas_{ref,avx2,avx512}.so, built from /root/reference by
oracle/Makefile).  The reference ships no tests or vectors of its own (SURVEY.md
section 4), so these outputs are the fixtures that pin the oracle and the CUDA
path.  Run in the build container:  python tests/golden/make_golden.py
TEST INFRASTRUCTURE."""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

import cases  # noqa: E402
from helpers import GF4, GF16, GF256, ref_table, reference  # noqa: E402


def main():
    ref = reference("ref")
    assert ref is not None, "oracle/_ref/liboblas_ref.so missing: run make -C oracle ref"
    g = {"generator": "tests/golden/make_golden.py on unmodified /root/reference (oblas_ref.c build)"}

    # tables (gf2_8_tables.h etc. as compiled into the reference)
    names = {0: "LOG", 1: "EXP", 2: "INV", 3: "SHUF_LO", 4: "SHUF_HI"}
    g["tables"] = {}
    for bits in (2, 4, 8):
        for which, nm in names.items():
            t = ref_table(ref, bits, which)
            g["tables"]["GF2_%d_%s" % (bits, nm)] = {"len": int(len(t)), "sha256": hashlib.sha256(t.tobytes()).hexdigest()}
    # the corrected GF(4) table as ref_driver.c feeds it to oblas_axpy
    import numpy as np
    fix_hi = np.ctypeslib.as_array(ref.refdrv_gf4_fixed_table(1), shape=(64,)).copy()
    g["tables"]["GF2_2_SHUF_HI_fixed"] = {"len": 64, "sha256": hashlib.sha256(fix_hi.tobytes()).hexdigest()}
    # table self-check against GF_MUL (oblas.h:11-12): SURVEY.md section 10
    bad = {}
    for bits, field in ((2, GF4), (4, GF16), (8, GF256)):
        lo, hi = ref_table(ref, bits, 3), ref_table(ref, bits, 4)
        n = 1 << bits
        cnt = 0
        for u in range(n):
            for x in range(256):
                got = hi[u * 16 + (x >> 4)] ^ lo[u * 16 + (x & 15)]
                want = 0
                for k in range(8 // bits):
                    e = (x >> (k * bits)) & (n - 1)
                    want |= ref.refdrv_gf_mul(bits, u, e) << (k * bits)
                cnt += int(got != want)
        bad[str(bits)] = cnt
    g["table_selfcheck_bad"] = bad

    # layout
    lay = {"octmat_align": {}, "sizeof": [ref.refdrv_sizeof(i) for i in range(3)],
           "bits_offset": [ref.refdrv_bits_offset(i) for i in range(3)]}
    for v in ("ref", "avx2", "avx512"):
        lib = reference(v)
        if lib is None:
            continue
        al = lib.refdrv_octmat_align()
        ent = {"octmat": [], "gfmat": [], "binmat": []}
        for cols in (1, 15, 16, 17, 100, 1024, 1280, 1281, 10000):
            m = lib.octmat_new(2, cols)
            ent["octmat"].append([cols, m.contents.stride])
            lib.octmat_free(m)
        for field in (0, 1, 2, 3):
            for cols in (1, 33, 100, 2048, 8192):
                m = lib.gfmat_new(field, 2, cols)
                ent["gfmat"].append([field, cols, m.contents.stride, m.contents.align])
                lib.gfmat_free(m)
        for cols in (1, 100, 256, 257, 8192):
            m = lib.binmat_new(2, cols)
            ent["binmat"].append([cols, m.contents.stride])
            lib.binmat_free(m)
        lay["octmat_align"][str(al)] = ent
    g["layout"] = lay

    # row ops: every build must agree
    per_variant = {}
    for v in ("ref", "avx2", "avx512"):
        lib = reference(v)
        if lib is None:
            continue
        per_variant[v] = {
            "octmat": cases.octmat_rowop_digest(lib),
            "gfmat": {str(f): cases.gfmat_rowop_digest(lib, f) for f in (GF4, GF16, GF256)},
            "b32": cases.b32_digest(lib),
        }
    r = per_variant["ref"]
    for v, d in per_variant.items():
        assert d["octmat"]["digest"] == r["octmat"]["digest"], v
        for f in d["gfmat"]:
            assert d["gfmat"][f]["digest"] == r["gfmat"][f]["digest"], (v, f)
    if "avx2" in per_variant:
        assert per_variant["avx2"]["b32"] == r["b32"]
    g["rowops"] = {"octmat": r["octmat"], "gfmat": r["gfmat"], "b32": r["b32"],
                   "octmat_odd_stride": cases.octmat_odd_stride_digest(ref),
                   "b32_avx512_divergent": per_variant.get("avx512", {}).get("b32")}
    g["binmat"] = cases.binmat_ops(ref)
    g["gfmat_aux"] = {str(f): cases.gfmat_aux(ref, f) for f in (0, GF4, GF16, GF256)}
    g["l1"] = cases.l1_ops(ref, lambda bits, which: ref_table(ref, bits, which))

    # caller-side loops, composed of reference calls only (oracle/ref_driver.c)
    g["solve"] = [dict(case=list(c), **cases.solve_case(ref, c, cases.ref_solve_fns(ref))) for c in cases.SOLVE_CASES]
    # the AVX-512 build must leave the same bytes (GF256 cases; strides differ -> compare rank/pivots only)
    a512 = reference("avx512")
    if a512 is not None:
        for c, want in zip(cases.SOLVE_CASES, g["solve"]):
            got = cases.solve_case(a512, c, cases.ref_solve_fns(a512))
            assert got["rank"] == want["rank"] and got["pivcols"] == want["pivcols"], c
    g["apply"] = [dict(case=list(c), **cases.apply_case(ref, c, ref.refdrv_octmat_apply)) for c in cases.APPLY_CASES]

    out = os.path.join(HERE, "golden.json")
    with open(out, "w") as f:
        json.dump(g, f, indent=1, sort_keys=True)
    print("wrote", out, os.path.getsize(out), "bytes")
    print("table self-check bad counts:", bad)
    print("b32 ref:", r["b32"], "avx512:", g["rowops"]["b32_avx512_divergent"])


if __name__ == "__main__":
    main()
