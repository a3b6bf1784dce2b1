"""ctypes binding of liboblas_b200.so (include/oblas_b200.h, include/oblas_compat.h).

The shared library IS the product; this module only loads it and declares the
argument types.  There is no Python or CPU implementation behind these calls:
if the library has not been built (`python -c "import __graft_entry__ as g; g.build()"`
or `make -C oblas_b200/csrc`) importing fails, and every entry point fails on a
machine without a CUDA device.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboblas_b200.so")

GF2, GF4, GF16, GF256 = 0, 1, 2, 3
FIELD_BITS = {GF2: 1, GF4: 2, GF16: 4, GF256: 8}
TABLES_SHIPPED, TABLES_FIXED = 0, 1
OP_AXPY, OP_ADD, OP_SCAL, OP_SWAP, OP_ZERO, OP_AXPY_B32 = range(6)
MUL_AUTO, MUL_LIN, MUL_LUT = -1, 0, 1

u8p = C.POINTER(C.c_uint8)
u32p = C.POINTER(C.c_uint32)
i32p = C.POINTER(C.c_int32)


class Octmat(C.Structure):  # octmat.h:19-24
    _fields_ = [("rows", C.c_uint), ("cols", C.c_uint), ("stride", C.c_uint), ("bits", u8p)]


class Gfmat(C.Structure):  # gfmat.h:18-25
    _fields_ = [("field", C.c_int), ("rows", C.c_uint), ("cols", C.c_uint),
                ("stride", C.c_uint), ("align", C.c_uint), ("bits", u32p)]


class Binmat(C.Structure):  # binmat.h:6-11
    _fields_ = [("rows", C.c_uint), ("cols", C.c_uint), ("stride", C.c_uint), ("bits", u32p)]


class OblasB200Error(RuntimeError):
    pass


_lib = None


def lib():
    """The loaded library.  Raises if it was not built -- there is no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise OblasB200Error(
            "%s is missing: build it with `make -C oblas_b200/csrc` (nvcc, sm_100a). "
            "oblas_b200 has no CPU or PyTorch fallback." % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, sz, u32, u64, ci = C.c_void_p, C.c_size_t, C.c_uint32, C.c_uint64, C.c_int
    L.oblas_b200_init.argtypes = [ci]
    L.oblas_b200_set_device.argtypes = [ci]
    L.oblas_b200_set_wait_limit_ms.argtypes = [ci]
    L.oblas_b200_set_fault_injection.argtypes = [ci]
    L.oblas_b200_set_workspace_budget.argtypes = [sz]
    L.oblas_b200_set_workspace_budget.restype = sz
    L.oblas_b200_device.restype = ci
    L.oblas_b200_sm_count.restype = ci
    L.oblas_b200_last_error.restype = C.c_char_p
    L.oblas_b200_set_stream.argtypes = [vp]
    L.oblas_b200_get_stream.restype = vp
    L.oblas_b200_set_async.argtypes = [ci]
    L.oblas_b200_queue_stats.argtypes = [C.POINTER(u64), C.POINTER(u64)]
    L.oblas_b200_set_tuning.argtypes = [ci, ci]
    L.oblas_b200_make_tables.argtypes = [C.c_uint, C.c_uint, vp, vp, vp]
    L.oblas_b200_register_tables.argtypes = [C.c_uint, vp, vp, vp]
    L.oblas_b200_set_field_tables.argtypes = [ci, ci]
    L.oblas_b200_swapcols_batch.argtypes = [ci, vp, sz, sz, C.c_uint32, vp, vp, sz]
    L.oblas_b200_get_elements.argtypes = [ci, vp, sz, vp, vp, sz, vp]
    L.oblas_b200_set_elements.argtypes = [ci, vp, sz, vp, vp, vp, sz]
    L.oblas_b200_solve_tc_times.argtypes = [ci, C.POINTER(C.c_double), C.POINTER(C.c_uint64)]
    L.oblas_b200_solve_tc_times.restype = ci
    L.oblas_b200_set_panel_path.argtypes = [ci]
    L.oblas_b200_solve_tc_times_ex.argtypes = [ci, C.POINTER(C.c_double), C.POINTER(C.c_uint64), ci]
    L.oblas_b200_solve_tc_times_ex.restype = ci
    L.oblas_b200_set_lazy_panels.argtypes = [ci]
    L.oblas_b200_set_tall_mode.argtypes = [ci]
    L.oblas_b200_set_tall_mode.restype = None
    L.oblas_b200_lazy_stats.argtypes = [C.POINTER(u64), ci]
    L.oblas_b200_lazy_stats.restype = ci
    L.oblas_b200_solve_phases.argtypes = [ci, vp]
    L.oblas_b200_update_phases.argtypes = [ci, vp]
    L.oblas_b200_update_phases.restype = ci
    L.oblas_b200_launch_count.restype = u64
    L.oblas_b200_dev_alloc.restype = vp
    L.oblas_b200_dev_alloc.argtypes = [sz]
    L.oblas_b200_dev_free.argtypes = [vp]
    L.oblas_b200_host_alloc.restype = vp
    L.oblas_b200_host_alloc.argtypes = [sz]
    L.oblas_b200_host_free.argtypes = [vp]
    L.oblas_b200_managed_alloc.restype = vp
    L.oblas_b200_managed_alloc.argtypes = [sz, sz]
    L.oblas_b200_free.argtypes = [vp]
    L.oblas_b200_copy.argtypes = [vp, vp, sz]
    L.oblas_b200_memset.argtypes = [vp, ci, sz]
    L.oblas_b200_fill_random.argtypes = [vp, sz, u64, u64]
    L.oblas_b200_rowops.argtypes = [ci, ci, ci, ci, vp, sz, vp, sz, sz, vp, vp, vp, sz]
    L.oblas_b200_solve_batch.argtypes = [ci, vp, sz, sz, u32, u32, sz, vp, sz, sz, sz, vp, vp, sz]
    L.oblas_b200_solve_plan.argtypes = [ci, u32, u32, sz, sz, sz, C.POINTER(ci), C.POINTER(sz), C.POINTER(ci)]
    L.oblas_b200_solve_batch_host.argtypes = [ci, vp, sz, sz, u32, u32, sz, vp, sz, sz, sz, vp, vp, sz, sz]
    L.oblas_b200_solve_batch_host_ex.argtypes = [ci, vp, sz, sz, u32, u32, sz, vp, sz, sz, sz, vp, vp, sz, sz, C.c_uint]
    L.oblas_b200_device_count.restype = ci
    L.oblas_b200_solve_batch_host_multi.argtypes = [ci, vp, sz, sz, u32, u32, sz, vp, sz, sz, sz, vp, vp, sz, vp, ci, C.c_uint]
    L.oblas_b200_gf256_apply_sharded.argtypes = [vp, sz, u32, u32, vp, sz, vp, sz, sz, vp, ci, C.POINTER(C.c_double)]
    L.oblas_b200_gf256_apply_b32.argtypes = [vp, sz, u32, u32, vp, sz, vp, sz, sz, ci]
    L.oblas_b200_gf256_apply.argtypes = [vp, sz, u32, u32, vp, sz, vp, sz, sz, ci]
    L.oblas_b200_nnz_batch.argtypes = [ci, vp, sz, u32, vp, sz, u32, u32, vp]
    _bind_compat(L)
    _lib = L
    return L


def _bind_compat(L):
    """argtypes of the reference API (39 symbols) + batched / solve additions."""
    PO, PG, PB = C.POINTER(Octmat), C.POINTER(Gfmat), C.POINTER(Binmat)
    cu, u8, sz = C.c_uint, C.c_uint8, C.c_size_t
    L.octmat_new.restype = PO
    L.octmat_new.argtypes = [cu, cu]
    L.octmat_new_aligned.restype = PO
    L.octmat_new_aligned.argtypes = [cu, cu, cu]
    L.octmat_free.argtypes = [PO]
    L.oswaprow.argtypes = [PO, cu, cu]
    L.oaxpy.argtypes = [PO, PO, cu, cu, u8]
    L.oaddrow.argtypes = [PO, PO, cu, cu]
    L.oscal.argtypes = [PO, cu, u8]
    L.oaxpy_b32.argtypes = [PO, u32p, cu, u8]
    L.gfmat_new.restype = PG
    L.gfmat_new.argtypes = [C.c_int, cu, cu]
    L.gfmat_new_aligned.restype = PG
    L.gfmat_new_aligned.argtypes = [C.c_int, cu, cu, cu]
    L.gfmat_copy.restype = PG
    L.gfmat_copy.argtypes = [PG]
    L.gfmat_free.argtypes = [PG]
    L.gfmat_get.restype = u8
    L.gfmat_get.argtypes = [PG, cu, cu]
    L.gfmat_set.argtypes = [PG, cu, cu, u8]
    L.gfmat_add.argtypes = [PG, PG, cu, cu]
    L.gfmat_swaprow.argtypes = [PG, cu, cu]
    L.gfmat_axpy.argtypes = [PG, PG, cu, cu, u8]
    L.gfmat_scal.argtypes = [PG, cu, u8]
    L.gfmat_zero.argtypes = [PG, cu]
    L.gfmat_fill.argtypes = [PG, cu, u8p]
    L.gfmat_nnz.restype = cu
    L.gfmat_nnz.argtypes = [PG, cu, cu, cu]
    L.gfmat_swapcol.argtypes = [PG, cu, cu]
    L.binmat_new.restype = PB
    L.binmat_new.argtypes = [cu, cu]
    L.binmat_free.argtypes = [PB]
    L.binmat_get.restype = u8
    L.binmat_get.argtypes = [PB, cu, cu]
    L.binmat_set.argtypes = [PB, cu, cu, u8]
    L.binmat_add.argtypes = [PB, PB, cu, cu]
    L.binmat_swaprow.argtypes = [PB, cu, cu]
    L.binmat_zero.argtypes = [PB, cu]
    L.binmat_fill.argtypes = [PB, cu, u8p]
    L.binmat_expand.argtypes = [PB, u32p, cu, u8]
    L.binmat_nnz.restype = cu
    L.binmat_nnz.argtypes = [PB, cu, cu, cu]
    L.bfd_32.restype = C.c_uint32
    L.bfd_32.argtypes = [C.c_uint32, u8, u8, u8]
    L.bfx_32.restype = C.c_uint32
    L.bfx_32.argtypes = [C.c_uint32, u8, u8]
    L.oblas_alloc.restype = C.c_void_p
    L.oblas_alloc.argtypes = [sz, sz, sz]
    L.oblas_xor.argtypes = [u8p, u8p, sz]
    L.oblas_axpy.argtypes = [u8p, u8p, sz, u8p, u8p]
    L.oblas_scal.argtypes = [u8p, sz, u8p, u8p]
    L.oblas_swap.argtypes = [u8p, u8p, sz]
    L.oblas_axpy_gf2_gf256_32.argtypes = [u8p, u32p, sz, u8]
    # additions
    L.oaxpy_batch.argtypes = [PO, PO, u32p, u32p, u8p, sz]
    L.oaddrow_batch.argtypes = [PO, PO, u32p, u32p, sz]
    L.oscal_batch.argtypes = [PO, u32p, u8p, sz]
    L.oswaprow_batch.argtypes = [PO, u32p, u32p, sz]
    L.gfmat_axpy_batch.argtypes = [PG, PG, u32p, u32p, u8p, sz]
    L.gfmat_add_batch.argtypes = [PG, PG, u32p, u32p, sz]
    L.gfmat_scal_batch.argtypes = [PG, u32p, u8p, sz]
    L.gfmat_swaprow_batch.argtypes = [PG, u32p, u32p, sz]
    L.binmat_add_batch.argtypes = [PB, PB, u32p, u32p, sz]
    L.binmat_swaprow_batch.argtypes = [PB, u32p, u32p, sz]
    L.octmat_solve.restype = cu
    L.octmat_solve.argtypes = [PO, PO, u32p]
    L.gfmat_solve.restype = cu
    L.gfmat_solve.argtypes = [PG, PG, u32p]
    L.binmat_rref.restype = cu
    L.binmat_rref.argtypes = [PB, PB, u32p]
    L.octmat_solve_batch.argtypes = [C.POINTER(PO), C.POINTER(PO), u32p, sz]
    L.gfmat_solve_batch.argtypes = [C.POINTER(PG), C.POINTER(PG), u32p, sz]
    L.binmat_rref_batch.argtypes = [C.POINTER(PB), C.POINTER(PB), u32p, sz]
    L.octmat_apply.argtypes = [PO, PO, PO]


# symbols every build must export: the reference's 39 + tables + our additions
REFERENCE_SYMBOLS = [
    "bfd_32", "bfx_32", "oblas_alloc", "oblas_axpy", "oblas_axpy_gf2_gf256_32", "oblas_scal",
    "oblas_swap", "oblas_xor",
    "gfmat_new", "gfmat_copy", "gfmat_free", "gfmat_get", "gfmat_set", "gfmat_print", "gfmat_add",
    "gfmat_swaprow", "gfmat_axpy", "gfmat_scal", "gfmat_zero", "gfmat_fill", "gfmat_nnz",
    "gfmat_swapcol",
    "binmat_new", "binmat_free", "binmat_get", "binmat_set", "binmat_add", "binmat_swaprow",
    "binmat_zero", "binmat_fill", "binmat_expand", "binmat_nnz",
    "octmat_new", "octmat_free", "oswaprow", "oaxpy", "oaddrow", "oscal", "oaxpy_b32",
]
TABLE_SYMBOLS = {
    "GF2_8_LOG": 256, "GF2_8_EXP": 510, "GF2_8_INV": 256, "GF2_8_SHUF_LO": 4096, "GF2_8_SHUF_HI": 4096,
    "GF2_4_LOG": 16, "GF2_4_EXP": 30, "GF2_4_INV": 16, "GF2_4_SHUF_LO": 256, "GF2_4_SHUF_HI": 256,
    "GF2_2_LOG": 4, "GF2_2_EXP": 6, "GF2_2_INV": 4, "GF2_2_SHUF_LO": 64, "GF2_2_SHUF_HI": 64,
}
EXTENSION_SYMBOLS = [
    "oblas_b200_init", "oblas_b200_set_device", "oblas_b200_set_workspace_budget", "oblas_b200_set_wait_limit_ms", "oblas_b200_set_fault_injection", "oblas_b200_shutdown", "oblas_b200_device", "oblas_b200_sm_count",
    "oblas_b200_last_error", "oblas_b200_set_stream", "oblas_b200_use_own_stream", "oblas_b200_get_stream", "oblas_b200_sync",
    "oblas_b200_set_async", "oblas_b200_queue_stats", "oblas_b200_set_tuning", "oblas_b200_solve_tc_times", "oblas_b200_solve_tc_times_ex", "oblas_b200_set_lazy_panels", "oblas_b200_set_tall_mode", "oblas_b200_lazy_stats", "oblas_b200_set_panel_path", "oblas_b200_solve_phases", "oblas_b200_update_phases", "oblas_b200_launch_count", "oblas_b200_dev_alloc", "oblas_b200_dev_free",
    "oblas_b200_host_alloc", "oblas_b200_host_free", "oblas_b200_managed_alloc", "oblas_b200_free",
    "oblas_b200_copy", "oblas_b200_memset", "oblas_b200_fill_random", "oblas_b200_rowops",
    "oblas_b200_solve_batch", "oblas_b200_solve_plan", "oblas_b200_solve_batch_host", "oblas_b200_solve_batch_host_ex", "oblas_b200_solve_batch_host_multi", "oblas_b200_gf256_apply_sharded", "oblas_b200_device_count", "oblas_b200_gf256_apply", "oblas_b200_gf256_apply_b32",
    "oblas_b200_nnz_batch", "oblas_b200_swapcols_batch", "oblas_b200_make_tables", "oblas_b200_register_tables", "oblas_b200_set_field_tables", "oblas_b200_get_elements", "oblas_b200_set_elements",
    "octmat_new_aligned", "gfmat_new_aligned",
    "oaxpy_batch", "oaddrow_batch", "oscal_batch", "oswaprow_batch", "gfmat_axpy_batch",
    "gfmat_add_batch", "gfmat_scal_batch", "gfmat_swaprow_batch", "binmat_add_batch",
    "binmat_swaprow_batch", "octmat_solve", "gfmat_solve", "binmat_rref", "octmat_solve_batch",
    "gfmat_solve_batch", "binmat_rref_batch", "octmat_apply",
]


def table(name):
    """bytes of an exported table symbol"""
    n = TABLE_SYMBOLS[name]
    arr = (C.c_uint8 * n).in_dll(lib(), name)
    return bytes(arr)


def check(rc, what="oblas_b200"):
    if rc != 0:
        raise OblasB200Error("%s failed: %s" % (what, lib().oblas_b200_last_error().decode()))


def init(device=-1):
    check(lib().oblas_b200_init(device), "oblas_b200_init")
