/*
 * oblas_b200.h -- C ABI of liboblas_b200.so, the B200 (sm_100a) implementation of
 * the oblas hot path.  Plain pointers and sizes only; no CUDA or torch types.
 *
 * Two layers are exported by the same shared library:
 *
 *  (1) the reference's own API, unchanged (include/oblas_compat.h and the shim
 *      headers include/compat/{oblas,octmat,gfmat,binmat}.h): the 39 functions of
 *      liboblas.a with identical signatures and struct layouts, plus *_batch forms.
 *      A C program written against /root/reference/{oblas,octmat,gfmat,binmat}.h
 *      recompiles against include/compat and links -loblas_b200 without edits.
 *
 *  (2) this file: the device-resident / batched entry points the throughput
 *      numbers are measured on.  Every pointer named "device-visible" may be a
 *      cudaMalloc pointer, a managed pointer (what octmat_new & co. return) or
 *      host memory registered with CUDA; plain host memory is staged through the
 *      device where the function says so.
 *
 * All functions return 0 on success and a negative code on error unless stated;
 * oblas_b200_last_error() gives the text.  There is NO CPU fallback: without a
 * usable CUDA device every entry point fails (the drop-in void functions print to
 * stderr and exit(EIO), mirroring the reference's exit(ENOMEM) convention,
 * oblas.c:16-17).
 */
#ifndef OBLAS_B200_H
#define OBLAS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* field ids = gfmat_type of the reference (gfmat.h:6) */
#define OBLAS_B200_GF2 0
#define OBLAS_B200_GF4 1
#define OBLAS_B200_GF16 2
#define OBLAS_B200_GF256 3

/* table variant: 0 = split-nibble tables exactly as the reference ships them
 * (gf2_2_tables.h:27-33 carries a defect in GF2_2_SHUF_HI), 1 = corrected GF(4)
 * table (HI = LO << 4).  Identical for GF(16)/GF(256). */
#define OBLAS_B200_TABLES_SHIPPED 0
#define OBLAS_B200_TABLES_FIXED 1

/* row-operation kinds of oblas_b200_rowops() */
#define OBLAS_B200_OP_AXPY 0     /* oaxpy      octmat.c:32-46 / gfmat_axpy gfmat.c:133-145 */
#define OBLAS_B200_OP_ADD 1      /* oaddrow    octmat.c:48-52 / gfmat_add / binmat_add      */
#define OBLAS_B200_OP_SCAL 2     /* oscal      octmat.c:54-63 / gfmat_scal gfmat.c:147-156  */
#define OBLAS_B200_OP_SWAP 3     /* oswaprow   octmat.c:24-30 / gfmat_swaprow / binmat_swaprow */
#define OBLAS_B200_OP_ZERO 4     /* gfmat_zero gfmat.c:158-161 / binmat_zero binmat.c:59-62 */
#define OBLAS_B200_OP_AXPY_B32 5 /* oaxpy_b32  octmat.c:65-68 -> oblas_ref.c:19-28          */

/* multiply back-end of AXPY/SCAL: -1 = pick (linear bit-plane path when the table
 * set is GF(2)-linear, PRMT nibble look-up otherwise), 0 = force bit-plane,
 * 1 = force PRMT look-up (always valid). */
#define OBLAS_B200_MUL_AUTO (-1)
#define OBLAS_B200_MUL_LIN 0
#define OBLAS_B200_MUL_LUT 1

/* ---- context ------------------------------------------------------------ */
/* One context per CUDA device.  oblas_b200_init(device) creates the context of `device` if needed
 * and makes it the one the CALLING HOST THREAD works on (thread-local); the first device ever
 * initialised is the process default for threads that never chose.  device < 0: the default (or,
 * before any init, CUDA's current device).  Idempotent.  oblas_b200_set_device is the same call
 * under the name the multi-GPU callers use: one host thread per GPU, each thread calls
 * oblas_b200_set_device(d) once and then any other entry point.  Stream, tuning and workspace
 * settings belong to the device context. */
int oblas_b200_init(int device);
int oblas_b200_set_device(int device);
void oblas_b200_shutdown(void);           /* releases the contexts of all devices */
int oblas_b200_device(void);
int oblas_b200_sm_count(void);
const char *oblas_b200_last_error(void);
/* all launches go to this stream (a cudaStream_t passed as void*; NULL is CUDA's
 * legacy default stream).  After init the library uses its own non-blocking
 * stream; oblas_b200_use_own_stream() goes back to it. */
void oblas_b200_set_stream(void *cuda_stream);
void oblas_b200_use_own_stream(void);
void *oblas_b200_get_stream(void);
/* wait for the current stream; also collects the status words of the tensor-core kernels: returns
 * an error if a barrier wait of a launch since the last collection ran into its time limit (the
 * result of that launch is invalid).  oblas_b200_solve_batch / oblas_b200_gf256_apply only enqueue:
 * callers must sync (here, or on the stream they passed) before using the results, and should do it
 * through this call.  The host-buffer entry points and the struct API collect it themselves. */
int oblas_b200_sync(void);
/* drop-in (layer 1) calls synchronise before returning unless async != 0, in
 * which case they only enqueue and the caller must oblas_b200_sync() before
 * touching `bits` from the host */
void oblas_b200_set_async(int async);
/* async == 2: OP QUEUE.  The drop-in row operations on matrices made by *_new (oaxpy, oaddrow, oscal,
 * oswaprow, gfmat_add/axpy/scal/swaprow/zero, binmat_add/swaprow/zero) are collected instead of launched:
 * consecutive calls of one kind on one matrix pair become ONE kernel launch.  A batch ends -- and is
 * launched, in order, on the stream -- when the next call is of another kind or on another matrix, when it
 * touches a row a queued operation writes (or writes a row a queued operation reads: the operations of a
 * launch are unordered), at 65536 operations, and before every other entry point of this library does its
 * work.  oblas_b200_sync, the accessors (gfmat_get/set/print/fill/nnz/swapcol, binmat_*) and *_free launch
 * the queue AND wait.  As with async == 1 the caller must oblas_b200_sync() before dereferencing `bits`
 * directly (om_A, om_P): a macro cannot be intercepted.  An elimination loop written against the
 * reference (one oaxpy per row and column) then costs one or two launches per column instead of one per
 * row; INTEGRATION.md has the measured difference. */
/* launches made for / operations that went through the op queue so far (either may be NULL) */
void oblas_b200_queue_stats(uint64_t *batches, uint64_t *ops);
/* Four-Russians table geometry of the elimination / apply kernels (tuning knob for
 * measurements): 0 = 4-bit groups, 256-byte tiles; 1 = 6-bit groups, 128-byte tiles;
 * 2 = 7-bit groups, 64-byte tiles; 4 = 5-bit groups, 128-byte tiles; -1 = library default (picked by
 * what fits shared memory, oblas_b200_solve_plan reports it).  Results are identical.
 * Value 3 forces the tensor-core path (tcgen05 kind::mxf4 bit-sliced GF(2) products,
 * apply_tc.cuh): for apply_geo the dense apply, for solve_geo the GF(256) elimination in outer
 * panels of 128 columns with tcgen05 rank-128 updates.  -1 picks it by size (apply: M >= 256,
 * Kc >= 64, nbytes >= 480; solve: GF(256), rows >= 512, cols >= 256, strided batches) and the
 * shared-memory table kernels otherwise. */
void oblas_b200_set_tuning(int solve_geo, int apply_geo);
/* Bytes of device workspace (fp4 operand blocks, coefficient matrices E, row maps) the tensor-core
 * paths may hold per stream; default 3 GiB, 0 restores it.  A batch / column range that needs more
 * runs as several chunks of systems (elimination) or several column passes (apply) -- identical
 * results.  Returns the previous value. */
size_t oblas_b200_set_workspace_budget(size_t bytes);
/* Every pipeline barrier wait inside the tensor-core kernels is bounded by WALL time (%globaltimer):
 * default 10 s, ms <= 0 restores it.  A wait that gives up sets the launch's status word (see
 * oblas_b200_sync); the kernel then drains, releases its tensor memory and exits. */
void oblas_b200_set_wait_limit_ms(int ms);
/* TESTS ONLY: bits = 8 makes the operand producer of the tensor-core kernels stop delivering, which
 * forces the give-up path above; 0 = normal operation. */
void oblas_b200_set_fault_injection(int bits);
/* diagnostics/tests (identical results, slower): bit 0 disables the register-resident fast panel
 * factorisation, bit 1 the look-ahead of the outer-panel kernel (candidate elimination of the next
 * 16-column step overlapped with the current step's tile pass) */
void oblas_b200_set_panel_path(int general_only);
/* diagnostics: enable != 0 starts per-phase cycle counters of the elimination kernels (SM cycles
 * of thread 0 summed over CTAs); out (8 entries, may be NULL) receives the counters collected so far
 * (entries 5..7, outer-panel kernel only: inside the factorisation -- candidate elimination, tables
 * of the panel's bit matrix, coefficient vectors of all rows).
 * Table kernel: panel load, panel factorisation, trailing update, permutation, 0.  Outer-panel
 * kernel of the tensor-core path: panel load + clear, factorisation, unit injection / direct E
 * columns, tile passes, per-system prologue and epilogue. */
int oblas_b200_solve_phases(int enable, uint64_t *out);
/* diagnostics: enable != 0 starts the role counters of the tcgen05 rank-128 update kernel (SM
 * cycles of one thread per role, summed over CTAs); out (16 entries, may be NULL) receives what was
 * collected so far.  MMA thread: 0 wait for the A operand, 1 wait for drained accumulators, 2 wait
 * for a B stage, 3 whole loop; epilogue warp 0: 4 wait for complete accumulators, 5 read-out and
 * stores, 6 whole loop; producer: 7 wait for a free stage, 8 whole loop; A builder warp 0: 9 wait
 * for a free buffer, 10 build, 11 whole loop; 12 = column tiles processed. */
int oblas_b200_update_phases(int enable, uint64_t *out);
/* diagnostics: CUDA-event device times of the kernel classes of the tensor-core elimination,
 * recorded on the launching stream.  enable != 0 starts recording; ms_out / launches_out (4
 * entries each, may be NULL) receive what was recorded so far (0 outer-panel factorisation,
 * 1 pivot-row expansion, 2 tcgen05 rank-128 update, 3 row permutation) and reset it. */
int oblas_b200_solve_tc_times(int enable, double *ms_out, uint64_t *launches_out);
/* the same with nclasses (1..5) entries: class 4 = lazy outer panels, the full kernel's pass over the systems
 * the compact pass gave up (the 4-entry form counts it under class 0) */
int oblas_b200_solve_tc_times_ex(int enable, double *ms_out, uint64_t *launches_out, int nclasses);
/* Lazy outer panels of the tensor-core elimination.  Per outer panel of 128 columns only the windows of the
 * first `compact_rows` not yet pivoted rows (in position order) are factorised in shared memory; those rows are
 * updated against the OLD pivot rows (their own coefficient rows), every other row against the NEW pivot rows with
 * its own 128-byte window as the coefficient row -- the rows x 128 x 128 products of the coefficient matrix
 * disappear.  A system whose panel has a column without a pivot among the compact rows (rank deficiency, pivots
 * far down) is redone by the kernel that looks at all rows; results are identical either way.
 * compact_rows: rounded up to a multiple of 16, at least 128; 0 = off; < 0 = default (144, and only for batches of at
 * least 2 x (SM count) systems: smaller batches cannot put two compact CTAs on every SM and take the full kernel).
 * The compact pass is bound by the chain of dependent column eliminations of its look-ahead, not by table look-ups,
 * so it runs with the small 4-bit tables as two CTAs of 256 threads per SM.  Measured on B200 (K = 1024, 1280-byte
 * symbols, 592 systems): outer-panel kernel 7.11 -> 3.68 ms, second expansion of the pivot rows +1.40 ms, update
 * launches +0.4 ms: 21 520 -> 22 780 solves/s with compact sets of 144 rows (192: 22 390; 128: 21 670, the compact
 * pass then gives up 48 of 4736 panels); bit-exact either way (pytest -m gpu runs both). */
void oblas_b200_set_lazy_panels(int compact_rows);
/* out[0] = (system, outer panel) pairs factorised by the compact pass so far, out[1] = pairs it gave up;
 * reset != 0 clears the counters */
int oblas_b200_lazy_stats(uint64_t *out, int reset);
/* Shared-memory table kernel, systems too tall for 128-bit panels in shared memory (GF(2) K = 8192 ...): mode 1
 * (default) keeps the 128-bit panels and the big tables and holds the per-row state (35 bytes per row) in a per-CTA
 * block of global memory; mode 0 falls back to 64-bit panels in shared memory.  Results are identical. */
void oblas_b200_set_tall_mode(int mode);
/* number of kernels this library has launched so far */
uint64_t oblas_b200_launch_count(void);

/* ---- memory ------------------------------------------------------------- */
void *oblas_b200_dev_alloc(size_t bytes);
void oblas_b200_dev_free(void *p);
void *oblas_b200_host_alloc(size_t bytes); /* pinned */
void oblas_b200_host_free(void *p);
void *oblas_b200_managed_alloc(size_t bytes, size_t align);
void oblas_b200_free(void *p);             /* any of the above, or malloc memory */
int oblas_b200_copy(void *dst, const void *src, size_t bytes); /* async on the stream */
int oblas_b200_memset(void *dst, int value, size_t bytes);
/* synthetic data: byte n of the splitmix64 counter stream `seed` (same generator
 * as oracle/oblas_oracle.c:orc_fill_random), written by a kernel */
int oblas_b200_fill_random(void *dst_dev, size_t bytes, uint64_t seed, uint64_t byte_offset);

/* ---- batched row operations (rows a-2 .. a-8) ----------------------------- */
/* nops independent operations of kind `op` on rows of `row_bytes` bytes
 * (row_bytes % 16 == 0, base pointers 16-byte aligned, strides % 16 == 0):
 *     a[dst_rows[k]]  op=  u[k] (x) b[src_rows[k]]
 * a, b: device-visible matrix bases; strides in bytes; b may equal a.
 * dst_rows/src_rows/u: nops entries each, device-visible OR plain host memory
 * (host arrays are copied to the device first).  src_rows may be NULL for
 * SCAL/ZERO, u may be NULL for ADD/SWAP/ZERO.  For OP_AXPY_B32, b is a matrix of
 * uint32 mask words (row_bytes/32 words per row, oblas_ref.c:19-28).
 * Destination rows must be pairwise distinct and disjoint from source rows of
 * other operations in the same call (the operations are unordered). */
int oblas_b200_rowops(int field, int variant, int op, int mul_mode,
                      void *a, size_t a_stride, const void *b, size_t b_stride,
                      size_t row_bytes, const uint32_t *dst_rows,
                      const uint32_t *src_rows, const uint8_t *u, size_t nops);

/* ---- batched elimination + RHS update (rows a-9 .. a-11) ------------------- */
/* For every system s < nsys, in place:
 *   A_s (rows x cols field elements, row stride a_row_stride, a_row_bytes bytes of
 *   each row are processed = the padded stride of the reference container) and
 *   optionally B_s (rows x b_row_bytes bytes) go through the column-skipping
 *   Gauss-Jordan loop of DESIGN.md / oracle/ref_driver.c: rank_out[s] = rank,
 *   pivcols_out[s*rows + t] = column of pivot t (t < rank), A_s = RREF(A_s) with
 *   pivot rows on top, B_s transformed by the same row operations.  For a full
 *   rank square system A_s = I and B_s = A_s^-1 B_s.
 * A, B, rank_out, pivcols_out are device-visible.  B and pivcols_out may be NULL.
 * GF(4) always uses true field arithmetic (the corrected table). */
int oblas_b200_solve_batch(int field, void *A, size_t a_row_stride, size_t a_sys_stride,
                           uint32_t rows, uint32_t cols, size_t a_row_bytes,
                           void *B, size_t b_row_stride, size_t b_sys_stride,
                           size_t b_row_bytes, int32_t *rank_out,
                           uint32_t *pivcols_out, size_t nsys);

/* How oblas_b200_solve_batch would run a batch of this shape with the current settings:
 * *path = 1 tensor-core elimination (GF(256), outer panels of 128 columns + tcgen05 rank-128
 * updates), 0 shared-memory table kernel; *chunk_systems = systems per chunk of the tensor-core path
 * (workspace budget / per-system workspace), nsys on the table kernel; geometry[3] = {panel width in
 * 32-bit words (NBW), coefficient bits per look-up table (KB), table entry / tile width in bytes (WT)}
 * of the Four-Russians kernel that runs.  Any output may be NULL.  Fails if the rows do not fit. */
int oblas_b200_solve_plan(int field, uint32_t rows, uint32_t cols, size_t a_row_bytes, size_t b_row_bytes,
                          size_t nsys, int *path, size_t *chunk_systems, int *geometry);

/* Same job with HOST buffers (pinned for full speed, pageable works): systems are
 * streamed host -> device -> host in chunks, copies overlapped with the
 * elimination of the neighbouring chunks.  rank_out/pivcols_out are host arrays.
 * chunk_systems = 0 picks a default. */
int oblas_b200_solve_batch_host(int field, void *A_host, size_t a_row_stride, size_t a_sys_stride,
                                uint32_t rows, uint32_t cols, size_t a_row_bytes,
                                void *B_host, size_t b_row_stride, size_t b_sys_stride,
                                size_t b_row_bytes, int32_t *rank_out,
                                uint32_t *pivcols_out, size_t nsys, size_t chunk_systems);

/* The same with flags.  OBLAS_B200_SOLVE_SKIP_IDENTITY_A: for square systems the reduced A of a
 * full-rank system is the identity matrix -- known in advance -- so it is NOT copied back: A_host
 * keeps the input for systems with rank_out[s] == rows and receives the reduced form only for the
 * rank-deficient ones.  Saves rows*cols of the (rows*cols + rows*b_row_bytes) bytes returned per
 * system (44 % at K=1024 with 1280-byte symbols).  With A_host in pinned (cudaHostAlloc / cudaHostRegister)
 * memory a kernel writes the rank-deficient systems' A straight into the host buffer, so the host never waits
 * for the ranks of a chunk: 19.8 k against 17.7 k solves/s for the full copy-back on the same B200 box
 * (K=1024, 1280-byte symbols); pageable memory takes a deferred cudaMemcpyAsync per deficient system. */
#define OBLAS_B200_SOLVE_SKIP_IDENTITY_A 1u
int oblas_b200_solve_batch_host_ex(int field, void *A_host, size_t a_row_stride, size_t a_sys_stride,
                                   uint32_t rows, uint32_t cols, size_t a_row_bytes,
                                   void *B_host, size_t b_row_stride, size_t b_sys_stride,
                                   size_t b_row_bytes, int32_t *rank_out,
                                   uint32_t *pivcols_out, size_t nsys, size_t chunk_systems,
                                   unsigned flags);

/* ---- dense apply (row a-12) ---------------------------------------------- */
/* Y (M x nbytes) = C (M x Kc GF(256) coefficients) * D (Kc x nbytes), i.e. the
 * loop  for i,j: oaxpy(Y, D, i, j, C[i][j])  on a zeroed Y; accumulate != 0 keeps
 * Y's previous contents (Y ^= C*D).  nbytes % 16 == 0; device-visible pointers. */
int oblas_b200_gf256_apply(const void *C, size_t c_stride, uint32_t M, uint32_t Kc,
                           const void *D, size_t d_stride, void *Y, size_t y_stride,
                           size_t nbytes, int accumulate);

/* ---- mixed-field apply ("next" row f-2) --------------------------------------------------- */
/* Y (M x ncols bytes) = C (M x Kc GF(256) coefficients) * expand(Dbits): the Kc rows of D are GF(2) rows in
 * binmat layout (one bit per column, 32 columns per uint32 word, row stride d_stride_words) and every bit
 * counts as the field element 0 / 1 -- exactly what the loop
 *     for i, j:  oaxpy_b32(Y, Dbits + j * d_stride_words, i, C[i][j])        (octmat.c:65-68, oblas_ref.c:19-28)
 * leaves in a zeroed Y (accumulate != 0: Y ^= ...).  The bit -> element expansion happens inside the
 * operand expansion of the tcgen05 apply: M * Kc row operations become one tensor-core product.
 * ncols % 16 == 0; device-visible pointers. */
int oblas_b200_gf256_apply_b32(const void *C, size_t c_stride, uint32_t M, uint32_t Kc,
                               const uint32_t *Dbits, size_t d_stride_words, void *Y, size_t y_stride,
                               size_t ncols, int accumulate);

/* ---- several GPUs from ONE process ------------------------------------ */
/* One context per device (oblas_b200_set_device); the two calls below run one host thread per
 * device.  devices: ndev device ordinals, NULL = 0 .. ndev-1; ndev <= 0 = all visible devices. */
int oblas_b200_device_count(void);
/* oblas_b200_solve_batch_host_ex with the batch cut into one contiguous block of systems per
 * device (independent systems: no collective). */
int oblas_b200_solve_batch_host_multi(int field, void *A_host, size_t a_row_stride, size_t a_sys_stride,
                                      uint32_t rows, uint32_t cols, size_t a_row_bytes,
                                      void *B_host, size_t b_row_stride, size_t b_sys_stride,
                                      size_t b_row_bytes, int32_t *rank_out, uint32_t *pivcols_out,
                                      size_t nsys, const int *devices, int ndev, unsigned flags);
/* Y = C * D over GF(256) with HOST matrices (octmat.c:32-46 semantics, Y overwritten), the symbol-byte
 * columns of D and Y sharded over the devices in 256-byte aligned slices; C goes to the first device
 * once and from there to the others over NVLink (peer copies along a binary tree).  seconds_out (3
 * entries, may be NULL): coefficient upload + broadcast, slowest device's copy-in + apply + copy-out,
 * whole call. */
int oblas_b200_gf256_apply_sharded(const void *C_host, size_t c_stride, uint32_t M, uint32_t Kc,
                                   const void *D_host, size_t d_stride, void *Y_host, size_t y_stride,
                                   size_t nbytes, const int *devices, int ndev, double *seconds_out);

/* ---- row-degree kernels ("next" row f-1) ---------------------------------- */
/* out[k] = number of non-zero elements of row rows_idx[k] in columns [s, e) with
 * CORRECT counting (the reference's gfmat_nnz/binmat_nnz are only right for s==0,
 * gfmat.c:193-195; the drop-in gfmat_nnz/binmat_nnz symbols mirror the reference
 * bug-for-bug on the host).  bits: device-visible packed matrix. */
int oblas_b200_nnz_batch(int field, const void *bits, size_t row_stride, uint32_t cols,
                         const uint32_t *rows_idx, size_t nrows, uint32_t s, uint32_t e,
                         uint32_t *out);

/* ---- column operations and accessors on device-resident matrices ("next" row f-3) ---------- */
/* gfmat_swapcol (gfmat.c:215-231) for a batch: in system s (base + s * sys_stride, `rows` rows of
 * row_stride bytes) columns col_i[s] and col_j[s] are exchanged.  Elements are packed as in gfmat
 * (gfmat.c:88,97); an octmat is field GF256.  col_i/col_j: nsys entries, device-visible or host. */
int oblas_b200_swapcols_batch(int field, void *bits, size_t row_stride, size_t sys_stride, uint32_t rows,
                              const uint32_t *col_i, const uint32_t *col_j, size_t nsys);
/* gfmat_get / gfmat_set (gfmat.c:83-98) for n (row, column) pairs of one matrix.  Index arrays and
 * vals: device-visible or host; out_dev: device-visible.  Pairs of one set call must be distinct. */
int oblas_b200_get_elements(int field, const void *bits, size_t row_stride, const uint32_t *rows_idx,
                            const uint32_t *cols_idx, size_t n, uint8_t *out_dev);
int oblas_b200_set_elements(int field, void *bits, size_t row_stride, const uint32_t *rows_idx,
                            const uint32_t *cols_idx, const uint8_t *vals, size_t n);

/* ---- tables for other field polynomials ("next" row f-4) ---------------------------------- */
/* Host helper: split-nibble tables in the reference generator's layout (tablegen.c:66-106) for an
 * irreducible polynomial `poly` (leading term included, e.g. 0x11B) of degree exp_bits = 2, 4 or 8.
 * lo, hi: 16 << exp_bits bytes, inv: 1 << exp_bits bytes.  The rows lo + 16*u / hi + 16*u are what
 * oblas_axpy(a, b, k, lo_row, hi_row) / oblas_scal take (oblas.h:24-27). */
int oblas_b200_make_tables(unsigned exp_bits, unsigned poly, uint8_t *lo, uint8_t *hi, uint8_t *inv);
/* Registers a table set with the device; returns a handle (>= 16) to pass as `variant` to
 * oblas_b200_rowops together with the field of the same element width, and to
 * oblas_b200_set_field_tables for the elimination / apply kernels. */
int oblas_b200_register_tables(unsigned exp_bits, const uint8_t *lo, const uint8_t *hi, const uint8_t *inv);
/* Which table set the ELIMINATION (oblas_b200_solve_batch*, octmat_solve / gfmat_solve) and the GF(256) APPLY
 * use for `field` (GF4 / GF16 / GF256) from now on, in the calling thread's device context: variant 0 = the
 * reference's polynomial (7 / 19 / 285, tablegen.c:7-12; GF(4) with true field arithmetic), a handle of
 * oblas_b200_register_tables = the field polynomial those tables belong to.  Registration regenerates the
 * tables from the polynomial they imply; a handle whose tables are not those of a field (e.g. the shipped
 * GF(4) pair) is refused here.  All kernels -- the Four-Russians table kernels, the outer-panel
 * factorisation and the tcgen05 bit-matrix products -- take the polynomial as a run-time parameter. */
int oblas_b200_set_field_tables(int field, int variant);

#ifdef __cplusplus
}
#endif
#endif
