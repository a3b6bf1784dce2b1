"""Shared test plumbing: synthetic data, loaders for the checkers (oracle, compiled
reference, kernel emulation).  TEST INFRASTRUCTURE -- nothing here is imported by
the product package."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

GF2, GF4, GF16, GF256 = 0, 1, 2, 3
FIELD_BITS = {GF2: 1, GF4: 2, GF16: 4, GF256: 8}
OP_AXPY, OP_ADD, OP_SCAL, OP_SWAP, OP_ZERO, OP_AXPY_B32 = range(6)
MUL_LIN, MUL_LUT = 0, 1
TS_GF4_SHIPPED, TS_GF4_FIXED, TS_GF16, TS_GF256 = range(4)

u8p = C.POINTER(C.c_uint8)
u32p = C.POINTER(C.c_uint32)
i32p = C.POINTER(C.c_int32)


def ptr(a, t=u8p):
    return a.ctypes.data_as(t) if a is not None else None


# --------------------------------------------------------------------------- data
def splitmix_bytes(seed, nbytes, byte_offset=0):
    """byte n of the splitmix64 counter stream `seed` (== orc_fill_random and the
    library's fill kernel).  Never a GF(2)-linear generator."""
    first = byte_offset >> 3
    last = (byte_offset + nbytes + 7) >> 3
    n = np.arange(first, last, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = np.uint64(seed) + (n + np.uint64(1)) * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    b = z.view(np.uint8)
    lo = byte_offset - (first << 3)
    return b[lo:lo + nbytes].copy()


def fnv1a(buf):
    h = 0xCBF29CE484222325
    for x in np.asarray(buf, dtype=np.uint8).ravel().tolist():
        h = ((h ^ x) * 0x100000001B3) & 0xFFFFFFFFFFFFFFFF
    return h


def cpu_flags():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    return set(line.split(":", 1)[1].split())
    except OSError:
        pass
    return set()


# ------------------------------------------------------------------------ loaders
def _make(target_dir, *args):
    subprocess.run(["make", "-s", "-C", target_dir, *args], check=True,
                   stdout=subprocess.DEVNULL, stderr=subprocess.PIPE)


_oracle = None


def oracle():
    """ctypes handle of oracle/liboblas_oracle.so (our C restatement)."""
    global _oracle
    if _oracle is None:
        d = os.path.join(ROOT, "oracle")
        so = os.path.join(d, "liboblas_oracle.so")
        if not os.path.exists(so):
            _make(d, so)
        lib = C.CDLL(so)
        lib.orc_table.restype = u8p
        lib.orc_table.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_uint)]
        lib.orc_gf_mul.restype = C.c_uint8
        lib.orc_gf_mul.argtypes = [C.c_int, C.c_uint8, C.c_uint8]
        lib.orc_gf_inv.restype = C.c_uint8
        lib.orc_gf_inv.argtypes = [C.c_int, C.c_uint8]
        lib.orc_xor.argtypes = [u8p, u8p, C.c_size_t]
        lib.orc_swap.argtypes = [u8p, u8p, C.c_size_t]
        lib.orc_axpy.argtypes = [u8p, u8p, C.c_size_t, u8p, u8p]
        lib.orc_scal.argtypes = [u8p, C.c_size_t, u8p, u8p]
        lib.orc_axpy_b32.argtypes = [u8p, u32p, C.c_size_t, C.c_uint8]
        for f in ("orc_octmat_stride",):
            getattr(lib, f).restype = C.c_uint
        lib.orc_octmat_stride.argtypes = [C.c_uint, C.c_uint]
        lib.orc_gfmat_stride.restype = C.c_uint
        lib.orc_gfmat_stride.argtypes = [C.c_int, C.c_uint, C.c_uint]
        lib.orc_binmat_stride.restype = C.c_uint
        lib.orc_binmat_stride.argtypes = [C.c_uint]
        lib.orc_row_axpy.argtypes = [C.c_int, C.c_int, u8p, C.c_size_t, u8p, C.c_size_t,
                                     C.c_uint, C.c_uint, C.c_uint8]
        lib.orc_row_scal.argtypes = [C.c_int, C.c_int, u8p, C.c_size_t, C.c_uint, C.c_uint8]
        lib.orc_row_add.argtypes = [u8p, C.c_size_t, u8p, C.c_size_t, C.c_uint, C.c_uint]
        lib.orc_row_swap.argtypes = [u8p, C.c_size_t, C.c_uint, C.c_uint]
        lib.orc_row_zero.argtypes = [u8p, C.c_size_t, C.c_uint]
        lib.orc_row_axpy_b32.argtypes = [u8p, C.c_size_t, u32p, C.c_uint, C.c_uint8]
        lib.orc_get.restype = C.c_uint8
        lib.orc_get.argtypes = [C.c_int, u32p, C.c_uint, C.c_uint, C.c_uint, C.c_uint, C.c_uint]
        lib.orc_set.argtypes = [C.c_int, u32p, C.c_uint, C.c_uint, C.c_uint, C.c_uint, C.c_uint, C.c_uint8]
        lib.orc_binmat_set.argtypes = [u32p, C.c_uint, C.c_uint, C.c_uint, C.c_uint, C.c_uint, C.c_uint8]
        lib.orc_nnz.restype = C.c_uint
        lib.orc_nnz.argtypes = [C.c_int, u32p, C.c_uint, C.c_uint, C.c_uint, C.c_uint, C.c_uint, C.c_uint]
        lib.orc_fill.argtypes = [C.c_int, u32p, C.c_uint, C.c_uint, u8p]
        lib.orc_swapcol.argtypes = [C.c_int, u32p, C.c_uint, C.c_uint, C.c_uint, C.c_uint]
        lib.orc_solve.restype = C.c_uint
        lib.orc_solve.argtypes = [C.c_int, C.c_int, u8p, C.c_size_t, C.c_uint, C.c_uint, u8p,
                                  C.c_size_t, u32p]
        lib.orc_gf256_apply.argtypes = [u8p, C.c_size_t, C.c_uint, C.c_uint, u8p, C.c_size_t,
                                        u8p, C.c_size_t, C.c_size_t]
        lib.orc_fill_random.argtypes = [u8p, C.c_size_t, C.c_uint64, C.c_uint64]
        lib.orc_fnv1a.restype = C.c_uint64
        lib.orc_fnv1a.argtypes = [u8p, C.c_size_t]
        _oracle = lib
    return _oracle


def oracle_table(field, which, variant=0):
    n = C.c_uint()
    p = oracle().orc_table(field, which, variant, C.byref(n))
    return np.ctypeslib.as_array(p, shape=(n.value,)).copy()


# ---- the compiled reference (oracle/_ref), struct API ---------------------------
# struct layouts are the interface (octmat.h:19-24, gfmat.h:18-25, binmat.h:6-11): one definition
from oblas_b200.capi import Binmat, Gfmat, Octmat  # noqa: E402


def bind_struct_api(lib):
    """argtypes of the reference's 39-function API -- used both for oracle/_ref and
    for the drop-in symbols of liboblas_b200.so (same ABI: that is the point)."""
    PO, PG, PB = C.POINTER(Octmat), C.POINTER(Gfmat), C.POINTER(Binmat)
    lib.octmat_new.restype = PO
    lib.octmat_new.argtypes = [C.c_uint, C.c_uint]
    lib.octmat_free.argtypes = [PO]
    lib.oswaprow.argtypes = [PO, C.c_uint, C.c_uint]
    lib.oaxpy.argtypes = [PO, PO, C.c_uint, C.c_uint, C.c_uint8]
    lib.oaddrow.argtypes = [PO, PO, C.c_uint, C.c_uint]
    lib.oscal.argtypes = [PO, C.c_uint, C.c_uint8]
    lib.oaxpy_b32.argtypes = [PO, u32p, C.c_uint, C.c_uint8]
    lib.gfmat_new.restype = PG
    lib.gfmat_new.argtypes = [C.c_int, C.c_uint, C.c_uint]
    lib.gfmat_copy.restype = PG
    lib.gfmat_copy.argtypes = [PG]
    lib.gfmat_free.argtypes = [PG]
    lib.gfmat_get.restype = C.c_uint8
    lib.gfmat_get.argtypes = [PG, C.c_uint, C.c_uint]
    lib.gfmat_set.argtypes = [PG, C.c_uint, C.c_uint, C.c_uint8]
    lib.gfmat_add.argtypes = [PG, PG, C.c_uint, C.c_uint]
    lib.gfmat_swaprow.argtypes = [PG, C.c_uint, C.c_uint]
    lib.gfmat_axpy.argtypes = [PG, PG, C.c_uint, C.c_uint, C.c_uint8]
    lib.gfmat_scal.argtypes = [PG, C.c_uint, C.c_uint8]
    lib.gfmat_zero.argtypes = [PG, C.c_uint]
    lib.gfmat_fill.argtypes = [PG, C.c_uint, u8p]
    lib.gfmat_nnz.restype = C.c_uint
    lib.gfmat_nnz.argtypes = [PG, C.c_uint, C.c_uint, C.c_uint]
    lib.gfmat_swapcol.argtypes = [PG, C.c_uint, C.c_uint]
    lib.binmat_new.restype = PB
    lib.binmat_new.argtypes = [C.c_uint, C.c_uint]
    lib.binmat_free.argtypes = [PB]
    lib.binmat_get.restype = C.c_uint8
    lib.binmat_get.argtypes = [PB, C.c_uint, C.c_uint]
    lib.binmat_set.argtypes = [PB, C.c_uint, C.c_uint, C.c_uint8]
    lib.binmat_add.argtypes = [PB, PB, C.c_uint, C.c_uint]
    lib.binmat_swaprow.argtypes = [PB, C.c_uint, C.c_uint]
    lib.binmat_zero.argtypes = [PB, C.c_uint]
    lib.binmat_fill.argtypes = [PB, C.c_uint, u8p]
    lib.binmat_expand.argtypes = [PB, u32p, C.c_uint, C.c_uint8]
    lib.binmat_nnz.restype = C.c_uint
    lib.binmat_nnz.argtypes = [PB, C.c_uint, C.c_uint, C.c_uint]
    lib.bfd_32.restype = C.c_uint32
    lib.bfd_32.argtypes = [C.c_uint32, C.c_uint8, C.c_uint8, C.c_uint8]
    lib.bfx_32.restype = C.c_uint32
    lib.bfx_32.argtypes = [C.c_uint32, C.c_uint8, C.c_uint8]
    lib.oblas_alloc.restype = C.c_void_p
    lib.oblas_alloc.argtypes = [C.c_size_t, C.c_size_t, C.c_size_t]
    lib.oblas_xor.argtypes = [u8p, u8p, C.c_size_t]
    lib.oblas_axpy.argtypes = [u8p, u8p, C.c_size_t, u8p, u8p]
    lib.oblas_scal.argtypes = [u8p, C.c_size_t, u8p, u8p]
    lib.oblas_swap.argtypes = [u8p, u8p, C.c_size_t]
    lib.oblas_axpy_gf2_gf256_32.argtypes = [u8p, u32p, C.c_size_t, C.c_uint8]
    return lib


REF_VARIANTS = {"ref": None, "avx2": "avx2", "avx512": "avx512bw"}
_refs = {}


def reference(variant="ref"):
    """ctypes handle of oracle/_ref/liboblas_<variant>.so = the unmodified reference
    sources + oracle/ref_driver.c, or None when it is not available / the host CPU
    lacks the ISA."""
    if variant in _refs:
        return _refs[variant]
    need = REF_VARIANTS[variant]
    so = os.path.join(ROOT, "oracle", "_ref", "liboblas_%s.so" % variant)
    lib = None
    if not os.path.exists(so) and os.path.exists("/root/reference/oblas.c"):
        _make(os.path.join(ROOT, "oracle"), "ref")
    if os.path.exists(so) and (need is None or need in cpu_flags()):
        lib = bind_struct_api(C.CDLL(so))
        PO, PG, PB = C.POINTER(Octmat), C.POINTER(Gfmat), C.POINTER(Binmat)
        lib.refdrv_octmat_align.restype = C.c_uint
        lib.refdrv_sizeof.restype = C.c_uint
        lib.refdrv_sizeof.argtypes = [C.c_int]
        lib.refdrv_bits_offset.restype = C.c_uint
        lib.refdrv_bits_offset.argtypes = [C.c_int]
        lib.refdrv_table.restype = u8p
        lib.refdrv_table.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_uint)]
        lib.refdrv_gf_mul.restype = C.c_uint8
        lib.refdrv_gf_mul.argtypes = [C.c_int, C.c_uint8, C.c_uint8]
        lib.refdrv_gf4_fixed_table.restype = u8p
        lib.refdrv_gf4_fixed_table.argtypes = [C.c_int]
        lib.refdrv_octmat_solve.restype = C.c_uint
        lib.refdrv_octmat_solve.argtypes = [PO, PO, u32p]
        lib.refdrv_gfmat_solve.restype = C.c_uint
        lib.refdrv_gfmat_solve.argtypes = [PG, PG, u32p, C.c_int]
        lib.refdrv_binmat_rref.restype = C.c_uint
        lib.refdrv_binmat_rref.argtypes = [PB, PB, u32p]
        lib.refdrv_octmat_apply.argtypes = [PO, PO, PO]
        lib.refdrv_oaxpy_many.argtypes = [PO, PO, u32p, u32p, u8p, C.c_size_t]
        lib.refdrv_oaddrow_many.argtypes = [PO, PO, u32p, u32p, C.c_size_t]
        lib.refdrv_oscal_many.argtypes = [PO, u32p, u8p, C.c_size_t]
        lib.refdrv_oswaprow_many.argtypes = [PO, u32p, u32p, C.c_size_t]
        lib.refdrv_gfmat_axpy_many.argtypes = [PG, PG, u32p, u32p, u8p, C.c_size_t]
        lib.refdrv_gfmat_scal_many.argtypes = [PG, u32p, u8p, C.c_size_t]
        lib.refdrv_binmat_add_many.argtypes = [PB, PB, u32p, u32p, C.c_size_t]
    _refs[variant] = lib
    return lib


def ref_table(lib, field_bits, which):
    n = C.c_uint()
    p = lib.refdrv_table(field_bits, which, C.byref(n))
    return np.ctypeslib.as_array(p, shape=(n.value,)).copy()


def mat_view(m, dtype=np.uint8):
    """numpy view [rows, stride_bytes] over a struct-API matrix's storage."""
    mm = m.contents
    if isinstance(mm, Octmat):
        sb = mm.stride
    else:
        sb = mm.stride * 4
    buf = np.ctypeslib.as_array(C.cast(mm.bits, u8p), shape=(mm.rows, sb))
    return buf if dtype == np.uint8 else buf.view(dtype)


# ---- kernel emulation -------------------------------------------------------------
_emu = None


def emu():
    global _emu
    if _emu is None:
        d = os.path.join(ROOT, "tests", "emu")
        _make(d)
        lib = C.CDLL(os.path.join(d, "libob2_emu.so"))
        lib.emu_table.restype = u8p
        lib.emu_table.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_uint)]
        lib.emu_table_linear.argtypes = [C.c_int]
        lib.emu_rowops.argtypes = [C.c_int, C.c_int, C.c_int, u8p, C.c_size_t, u8p, C.c_size_t,
                                   C.c_uint32, u32p, u32p, u8p, C.c_uint32, C.c_uint32]
        lib.emu_rowop_explicit.argtypes = [C.c_int, C.c_int, u8p, u8p, C.c_uint32, u8p, u8p]
        lib.emu_rowop_bytes.argtypes = [C.c_int, u8p, u8p, C.c_size_t, u8p, u8p, C.c_uint32]
        lib.emu_fastdiv_check.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32]
        lib.emu_solve.argtypes = [C.c_int, u8p, C.c_size_t, C.c_size_t, C.c_uint32, C.c_uint32,
                                  C.c_uint32, u8p, C.c_size_t, C.c_size_t, C.c_uint32, i32p,
                                  u32p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_size_t, C.c_int, C.c_int, C.c_int]
        lib.emu_apply.argtypes = [u8p, C.c_size_t, C.c_uint32, C.c_uint32, u8p, C.c_size_t,
                                  u8p, C.c_size_t, C.c_uint32, C.c_int, C.c_uint32, C.c_int]
        lib.emu_solve_outer.argtypes = [u8p, C.c_size_t, C.c_size_t, C.c_uint32, C.c_uint32, C.c_uint32,
                                        u8p, C.c_size_t, C.c_size_t, C.c_uint32, i32p, u32p, C.c_uint32,
                                        C.c_uint32, C.c_uint32, C.c_size_t, C.c_int]
        lib.emu_solve_outer_lazy.argtypes = lib.emu_solve_outer.argtypes + [C.c_uint32, u32p]
        _emu = lib
    return _emu


# ---- oracle-side convenience -------------------------------------------------------
def field_row_bytes(field, cols, align=16):
    """padded row length in bytes of the reference containers"""
    o = oracle()
    if field == GF256:
        return o.orc_octmat_stride(cols, align)
    return 4 * o.orc_gfmat_stride(field, cols, align)


def oracle_solve(field, A, B, cols, variant=1):
    """A: [rows, a_bytes] uint8 (modified in place), B likewise or None.
    Returns (rank, pivcols)."""
    rows = A.shape[0]
    piv = np.full(rows, 0xFFFFFFFF, dtype=np.uint32)
    rank = oracle().orc_solve(field, variant, ptr(A), A.strides[0], rows, cols,
                              ptr(B) if B is not None else None,
                              B.strides[0] if B is not None else 0, ptr(piv, u32p))
    return rank, piv


def random_matrix(field, rows, cols, seed, align=16, rank_deficient=0):
    """[rows, padded bytes] uint8 with `cols` uniform field elements per row and zero
    padding; optionally rank-deficient (duplicate rows, zero column)."""
    nb = field_row_bytes(field, cols, align)
    bits = FIELD_BITS[field]
    used = (cols * bits + 7) // 8
    M = np.zeros((rows, nb), dtype=np.uint8)
    M[:, :used] = splitmix_bytes(seed, rows * used).reshape(rows, used)
    tail = cols * bits - (used - 1) * 8  # valid bits in the last used byte
    if tail < 8:
        M[:, used - 1] &= (1 << tail) - 1
    if rank_deficient >= 1 and rows > 3:
        M[rows // 2] = M[1]
    if rank_deficient >= 2 and cols > 5:
        # zero one whole column (element 3)
        e = 3
        byte, sh = (e * bits) // 8, (e * bits) % 8
        M[:, byte] &= np.uint8(~(((1 << bits) - 1) << sh) & 0xFF)
    return M
