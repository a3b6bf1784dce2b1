// oblas_b200.cu -- liboblas_b200.so: context, memory model and every exported entry
// point (include/oblas_b200.h, include/oblas_compat.h).  Kernels live in
// rowops.cuh / solve.cuh / apply.cuh.  sm_100a only; no CPU path: every compute
// entry point ends in a kernel launch or fails.
#define OBLAS_B200_BUILDING 1
#include <cuda_runtime.h>
#include <errno.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <string>
#include <mutex>
#include <thread>
#include <unordered_set>
#include <vector>

#include "apply.cuh"
#include "apply_tc.cuh"
#include "rowops.cuh"
#include "solve.cuh"
#include "tables_host.h"

#include "../../include/oblas_b200.h"

using namespace ob2;

// ============================================================================
// context
// ============================================================================
namespace {

struct Ctx {
  bool ready = false;
  int device = -1;
  int sm_count = 0;
  size_t smem_optin = 0;
  cudaStream_t own_stream = nullptr;
  cudaStream_t stream = nullptr;
  int async = 0;   // 0: drop-in calls synchronise; 1: enqueue only; 2: op queue (see OpQueue)
  // Op queue behind the drop-in row operations (oblas_b200_set_async(2)): consecutive calls of one kind
  // on one matrix pair are collected and become ONE rowop_kernel launch when the batch has to end -- a
  // different kind / matrix, a row that a queued operation writes or reads (the operations of a launch
  // are unordered), the capacity, or anything that lets the host observe results (oblas_b200_sync, the
  // accessors, *_free, every other entry point).
  struct OpQueue {
    std::mutex mu;
    bool flushing = false;
    int field = 0, op = 0;
    uint8_t *a = nullptr;
    const uint8_t *b = nullptr;
    size_t a_stride = 0, b_stride = 0, row_bytes = 0;
    std::vector<uint32_t> dst, src;
    std::vector<uint8_t> u;
    // rows of matrix `a` written / rows of matrix `b` read by the queued operations: stamp == batch id
    std::vector<uint32_t> written, read;
    uint32_t stamp = 1;
    const void *ok_a = nullptr, *ok_b = nullptr;   // matrices already checked to be device-visible
    uint64_t batches = 0, ops = 0;
  } q;
  int solve_geo = -1, apply_geo = -1;  // table geometry overrides (-1 = default)
  // tensor-core apply (apply_tc.cuh): expanded-D workspace and the device-written status word
  // one workspace per stream that has used the tensor-core paths (the host-streaming lanes run
  // concurrently on their own streams and must not share operand blocks / row maps)
  struct TcWs {
    cudaStream_t stream = nullptr;
    void *p = nullptr;
    size_t bytes = 0;
    bool used = false;
  } tc_slots[8];
  void *tc_ws = nullptr;      // workspace of the current stream (set by tc_ws_reserve)
  size_t tc_ws_bytes = 0;
  // pinned + mapped, one word per workspace slot: a kernel sets its slot's word when a barrier wait
  // gave up (time limit); a failure on one stream does not stop the kernels of the other lanes
  int *tc_status = nullptr;   // [8 slots][2]: {flag, wait limit in ms}
  int tc_wait_limit_ms = ob2::tc::kWaitLimitMs;
  int tc_inject = 0;          // tests: fault-injection bits OR-ed into ApplyTcArgs::dbg
  int tc_cur_slot = 0;
  size_t tc_ws_budget = (size_t)3 << 30;   // bytes of operand workspace per stream (oblas_b200_set_workspace_budget)
  bool tc_attr_set = false, perm_attr_set = false;   // cudaFuncSetAttribute is per device
  int max_clusters = -1;
  uint32_t *d_zero = nullptr;  // a device word holding row index 0 (single-row launches)
  // diagnostics: CUDA-event timing of the kernel classes of the tensor-core elimination
  bool tc_timing = false;
  std::vector<std::pair<int, std::pair<cudaEvent_t, cudaEvent_t>>> tc_events;  // (class, (start, stop))
  int no_fast_panel = 0;               // diagnostics: general panel path only
  int solve_tall = 1;                  // table kernel: per-row state in a global workspace when 128-bit panels do not fit shared memory (oblas_b200_set_tall_mode)
  unsigned long long *d_lazy_stats = nullptr;   // {compact, given up} (system, outer panel) pairs
  int tc_lazy_rows = 144;              // tensor-core elimination: rows of the compact set of the lazy outer panels (0 = off); 144 measured best
  bool tc_lazy_auto = true;            // default setting: lazy only for chunks of at least 2 x (SM count) systems (two compact CTAs per SM)
  unsigned long long *d_phase = nullptr;  // solve-kernel phase counters (diagnostics)
  unsigned long long *d_uprof = nullptr;  // update_tc_kernel role counters (diagnostics)
  DevTableSet *d_sets = nullptr;  // [TS_COUNT] in device memory
  DevTableSet h_sets[TS_COUNT];
  // table sets registered at run time for other field polynomials (row f-4): handle = 16 + index
  std::vector<DevTableSet *> d_custom;
  std::vector<int> custom_linear, custom_bits;
  std::vector<unsigned> custom_poly;   // field polynomial if the registered tables are those of a field, else 0
  int field_variant[4] = {0, 0, 0, 0};  // table set of the elimination / apply kernels per field (oblas_b200_set_field_tables)
  // device scratch for index arrays handed over from plain host memory
  void *scratch = nullptr;
  size_t scratch_bytes = 0;
  std::mutex mu;
  std::atomic<uint64_t> launches{0};
  // lanes of the host-streaming pipeline (oblas_b200_solve_batch_host): chunk c runs on lane c % kLanes
  // (own stream, own device buffers); kept between calls (cudaMalloc/cudaFree of GB-sized blocks
  // costs more than the elimination of a chunk)
  static constexpr int kLanes = 3;
  struct HostLane {
    cudaStream_t s = nullptr;
    uint8_t *A = nullptr, *B = nullptr;
    int32_t *rank = nullptr;
    uint32_t *piv = nullptr;
    size_t a_cap = 0, b_cap = 0, r_cap = 0, p_cap = 0;
  } lane[kLanes];
  void *pin = nullptr;  // pinned staging for rank / pivcols of a whole host-streaming call
  size_t pin_cap = 0;
};
// One context per CUDA device.  A host thread works on the device it selected with
// oblas_b200_set_device (thread-local), else on the process default = the device of the first
// oblas_b200_init.  Settings made through oblas_b200_set_* apply to the calling thread's device.
constexpr int kMaxDev = 16;
Ctx g_ctx[kMaxDev];
std::atomic<int> g_default_dev{-1};
thread_local int tl_dev = -1;
inline Ctx &cx() {
  int d = tl_dev >= 0 ? tl_dev : g_default_dev.load(std::memory_order_relaxed);
  return g_ctx[d < 0 ? 0 : d];
}
thread_local char tl_err[512] = "";
void lanes_release(Ctx &c);  // host-streaming lanes, defined with oblas_b200_solve_batch_host

int fail(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(tl_err, sizeof tl_err, fmt, ap);
  va_end(ap);
  return -1;
}
#define CK(call)                                                                   \
  do {                                                                             \
    cudaError_t e_ = (call);                                                       \
    if (e_ != cudaSuccess)                                                         \
      return fail("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

int queue_flush();
int ensure_init(bool flush_queue = true) {
  Ctx &c = cx();
  if (!c.ready) return oblas_b200_init(tl_dev);
  // entry points may be called from any host thread: make the context's device current
  int cur = -1;
  if (cudaGetDevice(&cur) != cudaSuccess || cur != c.device) CK(cudaSetDevice(c.device));
  // queue mode: whatever the drop-in calls have collected goes first (stream order does the rest)
  if (flush_queue && c.async == 2 && !c.q.flushing && !c.q.dst.empty()) return queue_flush();
  return 0;
}

// The void functions of the reference API cannot report errors; like the
// reference's exit(ENOMEM) (oblas.c:16-17) they terminate loudly.
[[noreturn]] void die(const char *what) {
  fprintf(stderr, "liboblas_b200: %s: %s\n", what, tl_err);
  exit(EIO);
}
#define MUST(expr, what)                                                           \
  do {                                                                             \
    if ((expr) != 0) die(what);                                                    \
  } while (0)

enum PtrKind { PK_HOST = 0, PK_DEVVIS = 1 };
PtrKind classify(const void *p) {
  if (!p) return PK_HOST;
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
    cudaGetLastError();
    return PK_HOST;
  }
  if (at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged ||
      at.type == cudaMemoryTypeHost)
    return PK_DEVVIS;
  return PK_HOST;
}

int scratch_reserve(size_t bytes) {
  if (bytes <= cx().scratch_bytes) return 0;
  if (cx().scratch) {
    CK(cudaStreamSynchronize(cx().stream));
    CK(cudaFree(cx().scratch));
    cx().scratch = nullptr;
    cx().scratch_bytes = 0;
  }
  size_t want = bytes < (1u << 20) ? (1u << 20) : bytes * 2;
  CK(cudaMalloc(&cx().scratch, want));
  cx().scratch_bytes = want;
  return 0;
}

int after_launch(const char *what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail("%s launch failed: %s", what, cudaGetErrorString(e));
  cx().launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

int finish_dropin() {
  if (cx().async) return 0;
  CK(cudaStreamSynchronize(cx().stream));
  return 0;
}

constexpr size_t kQueueCapacity = 1u << 16;
// launches what is queued (asynchronously, on the context's stream)
int queue_flush() {
  Ctx &c = cx();
  Ctx::OpQueue &q = c.q;
  std::lock_guard<std::mutex> lk(q.mu);
  if (q.flushing || q.dst.empty()) return 0;
  q.flushing = true;
  const bool needs_src = q.op == OP_AXPY || q.op == OP_ADD || q.op == OP_SWAP;
  const int rc = oblas_b200_rowops(q.field, 0, q.op, OBLAS_B200_MUL_AUTO, q.a, q.a_stride, needs_src ? q.b : q.a,
                                   needs_src ? q.b_stride : q.a_stride, q.row_bytes, q.dst.data(),
                                   q.src.data(), q.u.data(), q.dst.size());
  q.batches++;
  q.ops += q.dst.size();
  q.dst.clear(); q.src.clear(); q.u.clear();
  if (++q.stamp == 0) {   // stamp wrapped: forget everything
    std::fill(q.written.begin(), q.written.end(), 0u);
    std::fill(q.read.begin(), q.read.end(), 0u);
    q.stamp = 1;
  }
  q.flushing = false;
  return rc;
}
// the host is about to look at results (or hands memory back): launch the queue and wait
int queue_observe() {
  Ctx &c = cx();
  if (!c.ready || c.async != 2) return 0;
  if (queue_flush()) return -1;
  CK(cudaStreamSynchronize(c.stream));
  return 0;
}
int queue_push(int field, int op, uint8_t *a, size_t a_stride, const uint8_t *b, size_t b_stride, size_t row_bytes,
               uint32_t i, uint32_t j, uint8_t u) {
  Ctx::OpQueue &q = cx().q;
  const bool has_src = op == OP_AXPY || op == OP_ADD || op == OP_SWAP;
  if (!has_src) { b = a; b_stride = a_stride; j = i; }
  const bool same = (a == b);   // one matrix: a row may be written by one operation and read by another
  bool flush;
  {
    std::lock_guard<std::mutex> lk(q.mu);
    auto hit = [&](const std::vector<uint32_t> &v, uint32_t r) { return r < v.size() && v[r] == q.stamp; };
    flush = !q.dst.empty() &&
            (q.field != field || q.op != op || q.a != a || q.b != b || q.a_stride != a_stride || q.b_stride != b_stride ||
             q.row_bytes != row_bytes || q.dst.size() >= kQueueCapacity ||
             hit(q.written, i) ||                                   // written twice
             (same && (hit(q.read, i) ||                            // a queued operation still reads row i
                       (j != i && hit(q.written, j)) ||             // row j is the result of a queued operation
                       (op == OP_SWAP && hit(q.read, j)))));
  }
  if (flush && queue_flush()) return -1;
  std::lock_guard<std::mutex> lk(q.mu);
  if (q.dst.empty()) {
    q.field = field; q.op = op; q.a = a; q.b = b; q.a_stride = a_stride; q.b_stride = b_stride; q.row_bytes = row_bytes;
  }
  auto mark = [&](std::vector<uint32_t> &v, uint32_t r) {
    if (r >= v.size()) v.resize((size_t)r + 1024, 0u);
    v[r] = q.stamp;
  };
  q.dst.push_back(i);
  q.src.push_back(j);
  q.u.push_back(u);
  mark(q.written, i);
  if (op == OP_SWAP) mark(q.written, j);
  else if (same && j != i) mark(q.read, j);
  return 0;
}

// host streaming with OBLAS_B200_SOLVE_SKIP_IDENTITY_A and a pinned (device-visible) A_host: the A of the systems
// that are NOT of full rank goes straight from the lane buffer into host memory -- a handful of systems per chunk,
// decided on the device, so the host never has to wait for the ranks before it can reuse the lane
__global__ void copy_deficient_kernel(const uint8_t *src, uint8_t *dst, const int32_t *rank, uint32_t rows,
                                      uint32_t nsys, size_t sys_bytes) {
  for (uint32_t s = blockIdx.x; s < nsys; s += gridDim.x) {
    if ((uint32_t)rank[s] == rows) continue;
    const uint4 *a = reinterpret_cast<const uint4 *>(src + (size_t)s * sys_bytes);
    uint4 *d = reinterpret_cast<uint4 *>(dst + (size_t)s * sys_bytes);
    for (size_t i = threadIdx.x; i < sys_bytes / 16; i += blockDim.x) d[i] = a[i];
  }
}

__global__ void fill_random_kernel(uint64_t *dst, size_t nwords, uint64_t seed, uint64_t word0) {
  for (size_t n = (size_t)blockIdx.x * blockDim.x + threadIdx.x; n < nwords;
       n += (size_t)gridDim.x * blockDim.x) {
    uint64_t z = seed + (word0 + n + 1) * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    dst[n] = z ^ (z >> 31);
  }
}

// one warp per requested row: count non-zero elements in columns [s, e)
template <int EXP>
__global__ void nnz_kernel(const uint8_t *bits, size_t row_stride, const uint32_t *rows_idx,
                           uint32_t nrows, uint32_t s, uint32_t e, uint32_t *out) {
  constexpr uint32_t VPW = 32 / EXP;
  uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= nrows) return;
  const uint32_t *row = reinterpret_cast<const uint32_t *>(bits + (size_t)rows_idx[warp] * row_stride);
  uint32_t w0 = s / VPW, w1 = (e + VPW - 1) / VPW, cnt = 0;
  for (uint32_t w = w0 + lane; w < w1; w += 32) {
    uint32_t x = row[w], z = x;
    // fold every element onto its lowest bit
    if (EXP >= 2) z |= z >> 1;
    if (EXP >= 4) z |= z >> 2;
    if (EXP >= 8) z |= z >> 4;
    z &= (EXP == 1) ? 0xffffffffu : (EXP == 2) ? 0x55555555u : (EXP == 4) ? 0x11111111u : 0x01010101u;
    // keep elements with index in [s, e)
    uint32_t first = w * VPW;
    for (uint32_t k = 0; k < VPW; k++) {
      uint32_t col = first + k;
      if (col < s || col >= e) z &= ~(1u << (k * EXP));
    }
    cnt += __popc(z);
  }
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if (lane == 0) out[warp] = cnt;
}

// ---- column operations / accessors on packed matrices -----------
// element j of a row = bits [(j % vpw) * EXP, +EXP) of 32-bit word j / vpw (gfmat.c:88,97;
// for octmat bytes the same with EXP = 8 on a little-endian machine)
template <int EXP>
__device__ __forceinline__ uint32_t elem_get(uint32_t w, uint32_t at) { return (w >> at) & ((1u << EXP) - 1u); }
template <int EXP>
__device__ __forceinline__ uint32_t elem_put(uint32_t w, uint32_t at, uint32_t v) {
  const uint32_t m = ((1u << EXP) - 1u) << at;
  return (w & ~m) | ((v << at) & m);
}
// gfmat_swapcol (gfmat.c:215-231) for a batch: system s swaps columns ci[s] and cj[s] in all rows
template <int EXP>
__global__ void swapcols_kernel(uint8_t *bits, size_t row_stride, size_t sys_stride, uint32_t rows,
                                const uint32_t *ci, const uint32_t *cj, uint32_t nsys) {
  constexpr uint32_t VPW = 32 / EXP;
  const size_t total = (size_t)nsys * rows;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    const uint32_t sys = (uint32_t)(idx / rows), r = (uint32_t)(idx - (size_t)sys * rows);
    const uint32_t i = ci[sys], j = cj[sys];
    if (i == j) continue;
    uint32_t *row = reinterpret_cast<uint32_t *>(bits + (size_t)sys * sys_stride + (size_t)r * row_stride);
    const uint32_t p = i / VPW, q = j / VPW, pa = (i % VPW) * EXP, qa = (j % VPW) * EXP;
    if (p == q) {
      uint32_t w = row[p];
      const uint32_t vi = elem_get<EXP>(w, pa), vj = elem_get<EXP>(w, qa);
      row[p] = elem_put<EXP>(elem_put<EXP>(w, pa, vj), qa, vi);
    } else {
      const uint32_t wp = row[p], wq = row[q];
      row[p] = elem_put<EXP>(wp, pa, elem_get<EXP>(wq, qa));
      row[q] = elem_put<EXP>(wq, qa, elem_get<EXP>(wp, pa));
    }
  }
}
// gfmat_get / gfmat_set (gfmat.c:83-98) for n (row, column) pairs; set = clear + or with atomics so
// that different elements of one word may be written by different threads of the same launch
template <int EXP>
__global__ void get_elements_kernel(const uint8_t *bits, size_t row_stride, const uint32_t *ri, const uint32_t *ci,
                                    uint32_t n, uint8_t *out) {
  constexpr uint32_t VPW = 32 / EXP;
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const uint32_t *row = reinterpret_cast<const uint32_t *>(bits + (size_t)ri[k] * row_stride);
  out[k] = (uint8_t)elem_get<EXP>(row[ci[k] / VPW], (ci[k] % VPW) * EXP);
}
template <int EXP>
__global__ void set_elements_kernel(uint8_t *bits, size_t row_stride, const uint32_t *ri, const uint32_t *ci,
                                    const uint8_t *vals, uint32_t n) {
  constexpr uint32_t VPW = 32 / EXP;
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  uint32_t *w = reinterpret_cast<uint32_t *>(bits + (size_t)ri[k] * row_stride) + ci[k] / VPW;
  const uint32_t at = (ci[k] % VPW) * EXP, m = ((1u << EXP) - 1u) << at;
  atomicAnd(w, ~m);
  atomicOr(w, ((uint32_t)vals[k] << at) & m);
}

}  // namespace

// ============================================================================
// layer 2: oblas_b200_* C ABI
// ============================================================================
extern "C" {

int oblas_b200_init(int device) {
  int n = 0;
  CK(cudaGetDeviceCount(&n));
  if (n <= 0) return fail("no CUDA device");
  if (device < 0) device = g_default_dev.load();
  if (device < 0) CK(cudaGetDevice(&device));
  if (device >= n || device >= kMaxDev) return fail("no CUDA device %d (%d visible, at most %d supported)", device, n, kMaxDev);
  Ctx &c = g_ctx[device];
  std::lock_guard<std::mutex> lk(c.mu);
  CK(cudaSetDevice(device));
  if (!c.ready) {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
      return fail("device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
    c.device = device;
    c.sm_count = prop.multiProcessorCount;
    c.smem_optin = prop.sharedMemPerBlockOptin;
    CK(cudaStreamCreateWithFlags(&c.own_stream, cudaStreamNonBlocking));
    c.stream = c.own_stream;
    fill_all_table_sets(c.h_sets);
    CK(cudaMalloc(&c.d_sets, sizeof(DevTableSet) * TS_COUNT));
    CK(cudaMemcpy(c.d_sets, c.h_sets, sizeof(DevTableSet) * TS_COUNT, cudaMemcpyHostToDevice));
    c.ready = true;
  }
  int none = -1;
  g_default_dev.compare_exchange_strong(none, device);
  tl_dev = device;   // the calling thread works on this device from now on
  return 0;
}

// Selects the device the calling host thread works on (its context is created on first use).
int oblas_b200_set_device(int device) { return oblas_b200_init(device); }

static void shutdown_ctx(Ctx &c) {
  std::lock_guard<std::mutex> lk(c.mu);
  if (!c.ready) return;
  cudaSetDevice(c.device);
  cudaStreamSynchronize(c.stream);
  if (c.scratch) cudaFree(c.scratch);
  c.scratch = nullptr;
  c.scratch_bytes = 0;
  for (auto &w : c.tc_slots) {
    if (w.p) cudaFree(w.p);
    w = Ctx::TcWs();
  }
  c.tc_ws = nullptr;
  c.tc_ws_bytes = 0;
  if (c.tc_status) cudaFreeHost(c.tc_status);
  c.tc_status = nullptr;
  c.tc_attr_set = c.perm_attr_set = false;
  c.max_clusters = -1;
  if (c.d_zero) cudaFree(c.d_zero);
  c.d_zero = nullptr;
  if (c.d_phase) cudaFree(c.d_phase);
  if (c.d_uprof) cudaFree(c.d_uprof);
  if (c.d_lazy_stats) cudaFree(c.d_lazy_stats);
  c.d_phase = c.d_uprof = c.d_lazy_stats = nullptr;
  lanes_release(c);
  for (auto *d : c.d_custom) cudaFree(d);
  c.d_custom.clear();
  c.custom_linear.clear();
  c.custom_bits.clear();
  c.custom_poly.clear();
  for (int &v : c.field_variant) v = 0;
  cudaFree(c.d_sets);
  c.d_sets = nullptr;
  cudaStreamDestroy(c.own_stream);
  c.own_stream = c.stream = nullptr;
  c.ready = false;
}
// releases the contexts of ALL devices (the thread-local device selection of other threads
// becomes stale: their next call re-initialises that device)
void oblas_b200_shutdown(void) {
  for (auto &c : g_ctx) shutdown_ctx(c);
  g_default_dev.store(-1);
  tl_dev = -1;
}

int oblas_b200_device(void) { return cx().ready ? cx().device : -1; }
int oblas_b200_sm_count(void) { return ensure_init() ? -1 : cx().sm_count; }
const char *oblas_b200_last_error(void) { return tl_err; }
void oblas_b200_set_stream(void *s) {
  if (ensure_init()) return;
  cx().stream = (cudaStream_t)s;  // NULL is the legacy default stream
}
void oblas_b200_use_own_stream(void) {
  if (ensure_init()) return;
  cx().stream = cx().own_stream;
}
void *oblas_b200_get_stream(void) { return ensure_init() ? nullptr : (void *)cx().stream; }
int tc_check_status();
int oblas_b200_sync(void) {
  if (ensure_init()) return -1;
  CK(cudaStreamSynchronize(cx().stream));
  return tc_check_status();
}
void oblas_b200_set_async(int async) {
  if (ensure_init()) return;   // (flushes a queue that is being switched off)
  cx().async = async < 0 ? 0 : (async > 2 ? 2 : async);
}
// diagnostics: launches made for / operations that went through the op queue (set_async(2)) so far
void oblas_b200_queue_stats(uint64_t *batches, uint64_t *ops) {
  if (batches) *batches = cx().q.batches;
  if (ops) *ops = cx().q.ops;
}
void oblas_b200_set_panel_path(int general_only) { cx().no_fast_panel = general_only & 3; }
// Lazy outer panels of the tensor-core elimination (solve_tc_run): compact_rows = rows of the compact set
// (rounded up to a multiple of 16, at least 128), 0 = off (every outer panel factorised over all rows), < 0 = default (144)
void oblas_b200_set_lazy_panels(int compact_rows) {
  cx().tc_lazy_auto = compact_rows < 0;
  if (compact_rows < 0) compact_rows = 144;
  if (compact_rows > 0) {
    if (compact_rows < 128) compact_rows = 128;
    if (compact_rows > 1024) compact_rows = 1024;
    compact_rows = (compact_rows + 15) / 16 * 16;
  }
  cx().tc_lazy_rows = compact_rows;
}
// Table kernel, systems whose per-row state does not fit shared memory beside the tables of the 128-bit panels:
// 1 (default) = keep the 128-bit panels and hold the per-row state in a global workspace, 0 = 64-bit panels in
// shared memory (the behaviour before; rows <= ~9800)
void oblas_b200_set_tall_mode(int mode) { cx().solve_tall = mode < 0 ? 1 : (mode > 2 ? 2 : mode); }
void oblas_b200_set_tuning(int solve_geo, int apply_geo) {
  cx().solve_geo = solve_geo;
  cx().apply_geo = apply_geo;
}
uint64_t oblas_b200_launch_count(void) { return cx().launches.load(); }
// Bytes of operand / coefficient workspace the tensor-core paths may hold per stream (default 3 GiB).
// Batches and column ranges that need more run as several chunks / passes; results are identical.
// 0 restores the default.  Returns the previous value.
// Time limit of a pipeline barrier wait inside the tensor-core kernels (default 10 s; <= 0 restores it).
void oblas_b200_set_wait_limit_ms(int ms) {
  if (ensure_init()) return;
  std::lock_guard<std::mutex> lk(cx().mu);
  cx().tc_wait_limit_ms = ms > 0 ? ms : ob2::tc::kWaitLimitMs;
}
// TESTS ONLY.  bits = 8: the B-operand producer of the tensor-core kernels never delivers, so every
// consumer wait runs into the time limit -- the give-up path (status word, TMEM release) under test.
void oblas_b200_set_fault_injection(int bits) {
  if (ensure_init()) return;
  cx().tc_inject = bits & 8;
}
size_t oblas_b200_set_workspace_budget(size_t bytes) {
  if (ensure_init()) return 0;
  std::lock_guard<std::mutex> lk(cx().mu);
  const size_t old = cx().tc_ws_budget;
  cx().tc_ws_budget = bytes ? bytes : (size_t)3 << 30;
  return old;
}
// diagnostics: enable != 0 starts (and zeroes) per-phase cycle counters of the elimination
// kernel; out (may be NULL) receives 4 counters: panel load, panel factorisation,
// trailing update, permutation -- SM cycles of thread 0 summed over all CTAs
int oblas_b200_solve_phases(int enable, uint64_t *out) {
  if (ensure_init()) return -1;
  if (out && cx().d_phase) {
    CK(cudaStreamSynchronize(cx().stream));
    CK(cudaMemcpy(out, cx().d_phase, 8 * sizeof(uint64_t), cudaMemcpyDeviceToHost));
  }
  if (enable) {
    if (!cx().d_phase) CK(cudaMalloc(&cx().d_phase, 8 * sizeof(uint64_t)));
    CK(cudaMemset(cx().d_phase, 0, 8 * sizeof(uint64_t)));
  } else if (cx().d_phase) {
    CK(cudaFree(cx().d_phase));
    cx().d_phase = nullptr;
  }
  return 0;
}

// ---- memory -----------------------------------------------------------------
void *oblas_b200_dev_alloc(size_t bytes) {
  if (ensure_init()) return nullptr;
  void *p = nullptr;
  if (cudaMalloc(&p, bytes ? bytes : 1) != cudaSuccess) {
    fail("cudaMalloc(%zu) failed", bytes);
    cudaGetLastError();
    return nullptr;
  }
  return p;
}
void oblas_b200_dev_free(void *p) {
  if (p) cudaFree(p);
}
void *oblas_b200_host_alloc(size_t bytes) {
  if (ensure_init()) return nullptr;
  void *p = nullptr;
  if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) {
    fail("cudaHostAlloc(%zu) failed", bytes);
    cudaGetLastError();
    return nullptr;
  }
  return p;
}
void oblas_b200_host_free(void *p) {
  if (p) cudaFreeHost(p);
}
void *oblas_b200_managed_alloc(size_t bytes, size_t align) {
  if (ensure_init()) return nullptr;
  (void)align;  // cudaMallocManaged returns >= 256-byte aligned blocks
  void *p = nullptr;
  if (cudaMallocManaged(&p, bytes ? bytes : 1, cudaMemAttachGlobal) != cudaSuccess) {
    fail("cudaMallocManaged(%zu) failed", bytes);
    cudaGetLastError();
    return nullptr;
  }
  return p;
}
void oblas_b200_free(void *p) {
  if (!p) return;
  cx().q.ok_a = cx().q.ok_b = nullptr;   // the op queue's "known device-visible" pointers may go stale
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
    cudaGetLastError();
    free(p);
    return;
  }
  switch (at.type) {
  case cudaMemoryTypeDevice:
  case cudaMemoryTypeManaged:
    if (cx().ready) cudaStreamSynchronize(cx().stream);
    cudaFree(p);
    break;
  case cudaMemoryTypeHost:
    cudaFreeHost(p);
    break;
  default:
    free(p);
  }
}
int oblas_b200_copy(void *dst, const void *src, size_t bytes) {
  if (ensure_init()) return -1;
  CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, cx().stream));
  return 0;
}
int oblas_b200_memset(void *dst, int value, size_t bytes) {
  if (ensure_init()) return -1;
  CK(cudaMemsetAsync(dst, value, bytes, cx().stream));
  return 0;
}
int oblas_b200_fill_random(void *dst, size_t bytes, uint64_t seed, uint64_t byte_offset) {
  if (ensure_init()) return -1;
  if ((bytes & 7) || (byte_offset & 7) || (reinterpret_cast<uintptr_t>(dst) & 7))
    return fail("fill_random: dst, bytes and byte_offset must be multiples of 8");
  size_t nwords = bytes / 8;
  if (!nwords) return 0;
  size_t blocks = (nwords + 255) / 256;
  if (blocks > (size_t)cx().sm_count * 16) blocks = (size_t)cx().sm_count * 16;
  fill_random_kernel<<<(unsigned)blocks, 256, 0, cx().stream>>>((uint64_t *)dst, nwords, seed, byte_offset / 8);
  return after_launch("fill_random");
}

// ---- row operations ------------------------------------------------------------
// Bring an index/scalar array to the device if it lives in plain host memory.
// `slot` is a byte offset inside the scratch block.
static int stage_array(const void *src, size_t bytes, size_t slot, const void **out) {
  if (!src) {
    *out = nullptr;
    return 0;
  }
  if (classify(src) == PK_DEVVIS) {
    *out = src;
    return 0;
  }
  void *dst = (char *)cx().scratch + slot;
  CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, cx().stream));
  *out = dst;
  return 0;
}

int oblas_b200_rowops(int field, int variant, int op, int mul_mode, void *a, size_t a_stride,
                      const void *b, size_t b_stride, size_t row_bytes,
                      const uint32_t *dst_rows, const uint32_t *src_rows, const uint8_t *u,
                      size_t nops) {
  if (ensure_init()) return -1;
  if (nops == 0 || row_bytes == 0) return 0;
  if (op < OP_AXPY || op > OP_AXPY_B32) return fail("rowops: bad op %d", op);
  if ((row_bytes & 15) || (a_stride & 15) || (reinterpret_cast<uintptr_t>(a) & 15))
    return fail("rowops: a, a_stride and row_bytes must be multiples of 16");
  const bool needs_b = (op == OP_AXPY || op == OP_ADD || op == OP_SWAP || op == OP_AXPY_B32);
  if (needs_b) {
    if (!b || !src_rows) return fail("rowops: op %d needs b and src_rows", op);
    size_t balign = (op == OP_AXPY_B32) ? 3 : 15;
    if ((b_stride & balign) || (reinterpret_cast<uintptr_t>(b) & balign))
      return fail("rowops: b / b_stride misaligned");
  }
  const bool needs_u = (op == OP_AXPY || op == OP_SCAL || op == OP_AXPY_B32);
  if (needs_u && !u) return fail("rowops: op %d needs u", op);
  if (classify(a) != PK_DEVVIS || (needs_b && classify(b) != PK_DEVVIS))
    return fail("rowops: matrix pointers must be device-visible (device, managed or pinned)");
  const bool needs_tab = (op == OP_AXPY || op == OP_SCAL);
  int ts = -1, mode = MUL_LIN;
  std::lock_guard<std::mutex> lk(cx().mu);   // before the table-handle lookup: register_tables appends concurrently
  if (needs_tab && field != OB2_GF2) {
    bool linear;
    if (variant >= 16) {   // table set registered with oblas_b200_register_tables
      const size_t ci = (size_t)variant - 16;
      if (ci >= cx().d_custom.size()) return fail("rowops: unknown table handle %d", variant);
      if (cx().custom_bits[ci] != (int)kBits[field]) return fail("rowops: table handle %d is for %d-bit elements", variant, cx().custom_bits[ci]);
      ts = 1000 + (int)ci;
      linear = cx().custom_linear[ci] != 0;
    } else {
      ts = table_set_id(field, variant);
      if (ts < 0) return fail("rowops: bad field %d", field);
      linear = cx().h_sets[ts].linear != 0;
    }
    if (mul_mode == OBLAS_B200_MUL_AUTO) mode = linear ? MUL_LIN : MUL_LUT;
    else if (mul_mode == OBLAS_B200_MUL_LIN) {
      if (!linear) return fail("rowops: table set %d is not GF(2)-linear, MUL_LIN is invalid", ts);
      mode = MUL_LIN;
    } else mode = MUL_LUT;
  }
  // GF(2) has no tables (gfmat.c:14-19): the only scalars are 0 and 1 (u == 1 is XOR; the reference
  // dereferences NULL for u > 1).  u > 1 is a no-op here, in single calls and in batches alike
  // (RowOpArgs::u_max); the kernel still gets a valid table set.
  if (needs_tab && field == OB2_GF2) ts = TS_GF4_FIXED;
  const size_t idx_bytes = nops * sizeof(uint32_t);
  const size_t slot1 = (idx_bytes + 255) & ~(size_t)255, slot2 = 2 * slot1;
  if (scratch_reserve(slot2 + nops + 256)) return -1;
  const void *d_dst, *d_src, *d_u;
  if (stage_array(dst_rows, idx_bytes, 0, &d_dst)) return -1;
  if (stage_array(src_rows, idx_bytes, slot1, &d_src)) return -1;
  if (stage_array(needs_u ? u : nullptr, nops, slot2, &d_u)) return -1;

  const uint32_t cpr = (uint32_t)(row_bytes / 16);
  const size_t max_ops = (size_t)0x7fffffffu / cpr;
  for (size_t first = 0; first < nops; first += max_ops) {
    size_t cnt = nops - first < max_ops ? nops - first : max_ops;
    RowOpArgs p;
    memset(&p, 0, sizeof p);
    p.a = (uint8_t *)a;
    p.b = (const uint8_t *)b;
    p.a_stride = a_stride;
    p.b_stride = b_stride;
    p.chunks_per_row = cpr;
    FastDiv fd = make_fastdiv(cpr);
    p.div_mul = fd.mul;
    p.div_shift = fd.shift;
    p.total_chunks = (uint32_t)(cnt * cpr);
    p.dst_rows = (const uint32_t *)d_dst + first;
    p.src_rows = d_src ? (const uint32_t *)d_src + first : nullptr;
    p.u = d_u ? (const uint8_t *)d_u + first : nullptr;
    p.ts = ts >= 1000 ? cx().d_custom[ts - 1000] : (ts >= 0 ? cx().d_sets + ts : nullptr);
    p.u_max = field == OB2_GF2 ? 1u : 255u;
    launch_rowops(op, mode, p, cx().stream);
    if (after_launch("rowops")) return -1;
  }
  return 0;
}

// ---- elimination -----------------------------------------------------------------
int solve_tc_run(SolveArgs &p, size_t nsys);   // tensor-core path (GF(256)), defined with apply_tc_run
int tc_ws_reserve(size_t need);                // per-stream workspace slot, defined with the tensor-core host code
static bool solve_tc_eligible(int field, const SolveArgs &p, size_t nsys) {
  if (field != OB2_GF256 || p.a_ptrs || p.b_ptrs) return false;
  // rows: u16 row map; the final row permutation stages at least 16 bytes of every row in shared memory.
  // Systems too tall for the shared-memory row state of the outer-panel kernel (> ~2700 rows) keep it in a
  // global workspace (panel_kernel<.., GROWS>): that is what makes e.g. a K = 10000 GF(256) solve possible.
  if (p.rows > 65535u || 2 * (((size_t)p.rows * 2 + 15) / 16 * 16) + 80 + (size_t)p.rows * 16 > cx().smem_optin) return false;
  if (cx().solve_geo == 3) return true;
  if (cx().solve_geo >= 0) return false;
  // auto: the rank-128 tensor-core update pays once the systems are big enough to fill its tiles
  return p.rows >= 512 && p.cols >= 256 && nsys >= 1;
}
// inverse table (device) and polynomial the elimination / apply kernels use for `field`: the reference's
// (GF(4): true field arithmetic) unless oblas_b200_set_field_tables selected a registered set
static int field_arith(int field, const uint8_t **inv, unsigned *poly) {
  static const unsigned kRefPoly[4] = {3, 7, 19, 285};   // tablegen.c:7-12 (GF(2): x + 1, unused)
  Ctx &c = cx();
  const int v = c.field_variant[field];
  if (v >= 16) {
    const size_t ci = (size_t)v - 16;
    if (ci >= c.d_custom.size() || c.custom_bits[ci] != (int)kBits[field] || !c.custom_poly[ci])
      return fail("field tables: handle %d cannot be used for eliminations / applies over this field", v);
    *inv = c.d_custom[ci]->inv;
    *poly = c.custom_poly[ci];
    return 0;
  }
  const int ts = table_set_id(field, 1);
  *inv = ts >= 0 ? c.d_sets[ts].inv : nullptr;
  *poly = kRefPoly[field];
  return 0;
}
static int solve_common(int field, SolveArgs &p, size_t nsys) {
  if (field < OB2_GF2 || field > OB2_GF256) return fail("solve: bad field %d", field);
  if (p.rows == 0 || p.cols == 0 || nsys == 0) return 0;
  if ((p.a_bytes & 15) || (p.a_rs & 15)) return fail("solve: A row bytes / stride must be multiples of 16");
  if (p.B && ((p.b_bytes & 15) || (p.b_rs & 15))) return fail("solve: B row bytes / stride must be multiples of 16");
  uint32_t ebits = kBits[field];
  if ((size_t)p.cols * ebits > (size_t)p.a_bytes * 8) return fail("solve: cols exceed a_row_bytes");
  {
    const uint8_t *inv = nullptr;
    unsigned poly = 0;
    if (field_arith(field, &inv, &poly)) return -1;
    p.inv = inv;
    p.red = field_red(ebits, poly);
  }
  p.nsys = (uint32_t)nsys;
  p.phase = cx().d_phase;
  p.no_fast_panel = cx().no_fast_panel & 1;
  p.no_look_ahead = (cx().no_fast_panel >> 1) & 1;
  if (solve_tc_eligible(field, p, nsys)) return solve_tc_run(p, nsys);
  uint32_t grid = nsys < (size_t)cx().sm_count ? (uint32_t)nsys : (uint32_t)cx().sm_count;
  const int geo = cx().solve_geo == 3 ? -1 : cx().solve_geo;
  // Tall systems (rows x 35 bytes of per-row state do not fit beside the tables of the 128-bit panels): rather than
  // 64-bit panels in shared memory, 128-bit panels with the per-row state in a per-CTA block of global memory
  if (cx().solve_tall && (geo < 0 || geo == 1 || geo == 4) && !p.a_ptrs && !p.b_ptrs) {
    int plan[3] = {0, 0, 0};
    const bool fits = plan_solve(field, geo, p.rows, cx().smem_optin, plan);
    const int ggeo = geo == 4 ? 4 : 1;
    if ((!fits || plan[0] != 4 || (cx().solve_tall == 2 && geo < 0 && plan[1] != 6)) && solve_grows_fits(p.rows, ggeo)) {
      const size_t per_cta = solve_row_ws_bytes<4>(p.rows);
      const int wrc = tc_ws_reserve(per_cta * grid + 256);
      if (wrc < 0) return -1;
      if (wrc == 0) {
        p.row_ws = (uint8_t *)(((uintptr_t)cx().tc_ws + 255) & ~(uintptr_t)255);
        p.row_ws_cta = per_cta;
        int rc = launch_solve_grows(field, ggeo, p, grid, 512, cx().smem_optin, cx().stream);
        if (rc == 0) return after_launch("solve (tall)");
        if (rc != -1) return fail("solve: cudaFuncSetAttribute failed");
        p.row_ws = nullptr;
        p.row_ws_cta = 0;
      }
    }
  }
  int rc = launch_solve(field, geo, p, grid, 512, cx().smem_optin, cx().stream);
  if (rc == -1)
    return fail("solve: %u rows do not fit the shared-memory panel (limit %zu bytes); no fallback", p.rows, cx().smem_optin);
  if (rc) return fail("solve: cudaFuncSetAttribute failed");
  return after_launch("solve");
}

int oblas_b200_solve_batch(int field, void *A, size_t a_row_stride, size_t a_sys_stride,
                           uint32_t rows, uint32_t cols, size_t a_row_bytes, void *B,
                           size_t b_row_stride, size_t b_sys_stride, size_t b_row_bytes,
                           int32_t *rank_out, uint32_t *pivcols_out, size_t nsys) {
  if (ensure_init()) return -1;
  if (!A || !rank_out) return fail("solve: A and rank_out are required");
  if (classify(A) != PK_DEVVIS || (B && classify(B) != PK_DEVVIS) || classify(rank_out) != PK_DEVVIS ||
      (pivcols_out && classify(pivcols_out) != PK_DEVVIS))
    return fail("solve_batch: pointers must be device-visible; use oblas_b200_solve_batch_host for host buffers");
  if ((reinterpret_cast<uintptr_t>(A) & 15) || (a_sys_stride & 15) ||
      (B && ((reinterpret_cast<uintptr_t>(B) & 15) || (b_sys_stride & 15))))
    return fail("solve: base pointers / system strides must be multiples of 16");
  SolveArgs p;
  memset(&p, 0, sizeof p);
  p.A = (uint8_t *)A; p.a_rs = a_row_stride; p.a_ss = a_sys_stride;
  p.rows = rows; p.cols = cols; p.a_bytes = (uint32_t)a_row_bytes;
  p.B = (uint8_t *)B; p.b_rs = b_row_stride; p.b_ss = b_sys_stride;
  p.b_bytes = B ? (uint32_t)b_row_bytes : 0;
  p.rank = rank_out; p.pivcols = pivcols_out;
  std::lock_guard<std::mutex> lk(cx().mu);
  return solve_common(field, p, nsys);
}

}  // extern "C"

namespace {
using HostLane = Ctx::HostLane;
constexpr int kLanes = Ctx::kLanes;

int lane_reserve(HostLane &L, size_t a, size_t b, size_t r, size_t pv) {
  if (!L.s) CK(cudaStreamCreateWithFlags(&L.s, cudaStreamNonBlocking));
  auto grow = [&](void **p, size_t &cap, size_t want) -> int {
    if (want <= cap) return 0;
    if (*p) CK(cudaFree(*p));
    *p = nullptr;
    cap = 0;
    CK(cudaMalloc(p, want));
    cap = want;
    return 0;
  };
  if (grow((void **)&L.A, L.a_cap, a)) return -1;
  if (grow((void **)&L.B, L.b_cap, b)) return -1;
  if (grow((void **)&L.rank, L.r_cap, r)) return -1;
  if (grow((void **)&L.piv, L.p_cap, pv)) return -1;
  return 0;
}
void lanes_release(Ctx &c) {
  for (auto &L : c.lane) {
    if (L.s) cudaStreamSynchronize(L.s);
    cudaFree(L.A); cudaFree(L.B); cudaFree(L.rank); cudaFree(L.piv);
    if (L.s) cudaStreamDestroy(L.s);
    L = HostLane();
  }
  if (c.pin) cudaFreeHost(c.pin);
  c.pin = nullptr;
  c.pin_cap = 0;
}
}  // namespace

extern "C" {

int oblas_b200_solve_batch_host_ex(int field, void *A_host, size_t a_row_stride, size_t a_sys_stride,
                                   uint32_t rows, uint32_t cols, size_t a_row_bytes, void *B_host,
                                   size_t b_row_stride, size_t b_sys_stride, size_t b_row_bytes,
                                   int32_t *rank_out, uint32_t *pivcols_out, size_t nsys,
                                   size_t chunk_systems, unsigned flags) {
  if (ensure_init()) return -1;
  // OBLAS_B200_SOLVE_SKIP_IDENTITY_A: the reduced A of a full-rank square system is the identity --
  // known in advance, so it is not copied back (A_host keeps the input for those systems)
  const bool skip_ident = (flags & OBLAS_B200_SOLVE_SKIP_IDENTITY_A) && rows == cols;
  if (!A_host || !rank_out) return fail("solve_host: A and rank_out are required");
  if (nsys == 0) return 0;
  if ((a_sys_stride & 15) || (B_host && (b_sys_stride & 15)))
    return fail("solve_host: system strides must be multiples of 16");
  std::lock_guard<std::mutex> lk(cx().mu);
  const size_t per_sys = a_sys_stride + (B_host ? b_sys_stride : 0) + 4 + (pivcols_out ? (size_t)rows * 4 : 0);
  if (chunk_systems == 0) {
    // one wave of the persistent grid per chunk: no tail inside a chunk, fine-grained
    // overlap of the copies with the neighbouring chunks' elimination (measured best)
    chunk_systems = (size_t)cx().sm_count;
    size_t free_b = 0, total_b = 0;
    CK(cudaMemGetInfo(&free_b, &total_b));
    size_t have = 0;
    for (auto &L : cx().lane) have += L.a_cap + L.b_cap;
    while (chunk_systems > 1 && chunk_systems * per_sys * kLanes > (free_b + have) / 2) chunk_systems /= 2;
  }
  if (chunk_systems > nsys) chunk_systems = nsys;
  for (auto &L : cx().lane)
    if (lane_reserve(L, chunk_systems * a_sys_stride, B_host ? chunk_systems * b_sys_stride : 0,
                     chunk_systems * 4, pivcols_out ? chunk_systems * (size_t)rows * 4 : 0))
      return -1;
  // results are staged in pinned memory: a device-to-host copy into the caller's
  // (usually pageable) rank array would block the host and serialise the lanes
  const size_t pin_rank = (nsys * 4 + 255) & ~(size_t)255;
  const size_t pin_need = pin_rank + (pivcols_out ? nsys * (size_t)rows * 4 : 0);
  if (pin_need > cx().pin_cap) {
    if (cx().pin) CK(cudaFreeHost(cx().pin));
    cx().pin = nullptr;
    cx().pin_cap = 0;
    CK(cudaHostAlloc(&cx().pin, pin_need, cudaHostAllocDefault));
    cx().pin_cap = pin_need;
  }
  int32_t *pin_r = (int32_t *)cx().pin;
  uint32_t *pin_p = (uint32_t *)((uint8_t *)cx().pin + pin_rank);

  // skip_ident with pinned host memory: a kernel copies the A of the rank-deficient systems back (no host wait);
  // pageable host memory takes the deferred cudaMemcpyAsync path below
  uint8_t *A_host_dev = nullptr;
  if (skip_ident && (reinterpret_cast<uintptr_t>(A_host) & 15) == 0) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, A_host) == cudaSuccess && at.type == cudaMemoryTypeHost && at.devicePointer)
      A_host_dev = (uint8_t *)at.devicePointer;
    else
      cudaGetLastError();
  }
  const bool trace = getenv("OBLAS_B200_TRACE") != nullptr;
  std::vector<cudaEvent_t> ev;
  cudaEvent_t ev_start = nullptr;
  if (trace) {
    cudaEventCreate(&ev_start);
    cudaEventRecord(ev_start, cx().lane[0].s);
  }
  auto mark = [&](cudaStream_t st) {
    if (!trace) return;
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e, st);
    ev.push_back(e);
  };
  cudaStream_t saved = cx().stream;
  int rc = 0;
  size_t c = 0;
  // skip_ident: the copy-back of A waits until the chunk's ranks are on the host; it is issued when
  // the lane is about to be reused (3 chunks later) or at the end, for the deficient systems only
  struct Pending { size_t first = 0, cnt = 0; cudaEvent_t done = nullptr; } pend[kLanes];
  auto flush_lane = [&](int li) -> int {
    Pending &pd = pend[li];
    if (!pd.cnt) return 0;
    HostLane &L = cx().lane[li];
    CK(cudaEventSynchronize(pd.done));
    for (size_t k = 0; k < pd.cnt;) {
      if ((uint32_t)pin_r[pd.first + k] == rows) { k++; continue; }
      size_t k1 = k + 1;
      while (k1 < pd.cnt && (uint32_t)pin_r[pd.first + k1] != rows) k1++;
      CK(cudaMemcpyAsync((uint8_t *)A_host + (pd.first + k) * a_sys_stride, L.A + k * a_sys_stride,
                         (k1 - k) * a_sys_stride, cudaMemcpyDeviceToHost, L.s));
      k = k1;
    }
    pd.cnt = 0;
    return 0;
  };
  for (size_t first = 0; first < nsys && rc == 0; first += chunk_systems, c++) {
    HostLane &L = cx().lane[c % kLanes];
    size_t cnt = nsys - first < chunk_systems ? nsys - first : chunk_systems;
    if (skip_ident && !A_host_dev && flush_lane((int)(c % kLanes))) { rc = -1; break; }
    cudaError_t e = cudaMemcpyAsync(L.A, (uint8_t *)A_host + first * a_sys_stride, cnt * a_sys_stride,
                                    cudaMemcpyHostToDevice, L.s);
    if (e == cudaSuccess && B_host)
      e = cudaMemcpyAsync(L.B, (uint8_t *)B_host + first * b_sys_stride, cnt * b_sys_stride,
                          cudaMemcpyHostToDevice, L.s);
    if (e == cudaSuccess && pivcols_out) e = cudaMemsetAsync(L.piv, 0xff, cnt * (size_t)rows * 4, L.s);
    if (e != cudaSuccess) { rc = fail("solve_host: copy to device failed: %s", cudaGetErrorString(e)); break; }
    mark(L.s);
    SolveArgs p;
    memset(&p, 0, sizeof p);
    p.A = L.A; p.a_rs = a_row_stride; p.a_ss = a_sys_stride;
    p.rows = rows; p.cols = cols; p.a_bytes = (uint32_t)a_row_bytes;
    p.B = B_host ? L.B : nullptr; p.b_rs = b_row_stride; p.b_ss = b_sys_stride;
    p.b_bytes = B_host ? (uint32_t)b_row_bytes : 0;
    p.rank = L.rank; p.pivcols = pivcols_out ? L.piv : nullptr;
    cx().stream = L.s;
    rc = solve_common(field, p, cnt);
    cx().stream = saved;
    if (rc) break;
    mark(L.s);
    e = cudaSuccess;
    if (!skip_ident)
      e = cudaMemcpyAsync((uint8_t *)A_host + first * a_sys_stride, L.A, cnt * a_sys_stride,
                          cudaMemcpyDeviceToHost, L.s);
    if (e == cudaSuccess && B_host)
      e = cudaMemcpyAsync((uint8_t *)B_host + first * b_sys_stride, L.B, cnt * b_sys_stride,
                          cudaMemcpyDeviceToHost, L.s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(pin_r + first, L.rank, cnt * 4, cudaMemcpyDeviceToHost, L.s);
    if (e == cudaSuccess && pivcols_out)
      e = cudaMemcpyAsync(pin_p + first * rows, L.piv, cnt * (size_t)rows * 4, cudaMemcpyDeviceToHost, L.s);
    if (e != cudaSuccess) { rc = fail("solve_host: copy to host failed: %s", cudaGetErrorString(e)); break; }
    if (skip_ident && A_host_dev) {
      const uint32_t g = cnt < 64 ? (uint32_t)cnt : 64u;
      copy_deficient_kernel<<<g, 256, 0, L.s>>>(L.A, A_host_dev + first * a_sys_stride, L.rank, rows, (uint32_t)cnt, a_sys_stride);
      if (after_launch("solve_host (copy-back of the rank-deficient systems)")) { rc = -1; break; }
    } else if (skip_ident) {
      Pending &pd = pend[c % kLanes];
      if (!pd.done && cudaEventCreateWithFlags(&pd.done, cudaEventDisableTiming) != cudaSuccess) { rc = fail("solve_host: event"); break; }
      cudaEventRecord(pd.done, L.s);
      pd.first = first;
      pd.cnt = cnt;
    }
    mark(L.s);
  }
  for (int li = 0; li < kLanes && skip_ident; li++) {
    if (rc == 0 && flush_lane(li)) rc = -1;
    if (pend[li].done) cudaEventDestroy(pend[li].done);
  }
  for (auto &L : cx().lane) {
    cudaError_t e = cudaStreamSynchronize(L.s);
    if (e != cudaSuccess && rc == 0) rc = fail("solve_host: %s", cudaGetErrorString(e));
  }
  // a tensor-core kernel of one of the lanes may have given up on a barrier wait: its chunk is invalid
  if (rc == 0) rc = tc_check_status();
  if (rc == 0) {
    memcpy(rank_out, pin_r, nsys * 4);
    if (pivcols_out) memcpy(pivcols_out, pin_p, nsys * (size_t)rows * 4);
  }
  if (trace) {
    for (size_t k = 0; k + 2 < ev.size() + 1; k += 3) {
      float a = 0, b = 0, d = 0;
      cudaEventElapsedTime(&a, ev_start, ev[k]);
      cudaEventElapsedTime(&b, ev_start, ev[k + 1]);
      cudaEventElapsedTime(&d, ev_start, ev[k + 2]);
      fprintf(stderr, "chunk %zu: h2d done %.1f ms, kernel done %.1f ms, d2h done %.1f ms\n", k / 3, a, b, d);
    }
    for (auto e : ev) cudaEventDestroy(e);
    cudaEventDestroy(ev_start);
  }
  return rc;
}

int oblas_b200_solve_batch_host(int field, void *A_host, size_t a_row_stride, size_t a_sys_stride,
                                uint32_t rows, uint32_t cols, size_t a_row_bytes, void *B_host,
                                size_t b_row_stride, size_t b_sys_stride, size_t b_row_bytes,
                                int32_t *rank_out, uint32_t *pivcols_out, size_t nsys,
                                size_t chunk_systems) {
  return oblas_b200_solve_batch_host_ex(field, A_host, a_row_stride, a_sys_stride, rows, cols, a_row_bytes, B_host,
                                        b_row_stride, b_sys_stride, b_row_bytes, rank_out, pivcols_out, nsys,
                                        chunk_systems, 0);
}

// ---- tensor-core apply: host side ---------------------------------------------------
// Column window by column window: expand D into the fp4 operand blocks (workspace), then the
// persistent tcgen05 kernel.  The workspace is bounded (tc_ws_budget) and cached.
// Status words of the tensor-core kernels: one per workspace slot (= per stream).  A kernel whose
// barrier wait exceeded the time limit sets its slot's word; the result of that launch is invalid.
// Collected by oblas_b200_sync, by the host-streaming solve before it returns, and by tc_prepare
// (an uncollected failure of an earlier asynchronous call on the same stream fails the next call
// once instead of being lost).
int tc_check_status() {
  Ctx &c = cx();
  if (!c.tc_status) return 0;
  int bad = -1;
  for (int k = 0; k < 8; k++)
    if (reinterpret_cast<volatile int *>(c.tc_status)[2 * k] != 0) {
      c.tc_status[2 * k] = 0;
      bad = k;
    }
  if (bad >= 0)
    return fail("tensor-core path: a pipeline barrier wait exceeded its time limit (workspace slot %d); result invalid", bad);
  return 0;
}
int tc_ws_slot() {   // slot of the current stream (claims a free one)
  Ctx &c = cx();
  for (int k = 0; k < 8; k++)
    if (c.tc_slots[k].used && c.tc_slots[k].stream == c.stream) return k;
  for (int k = 0; k < 8; k++)
    if (!c.tc_slots[k].used) {
      c.tc_slots[k].used = true;
      c.tc_slots[k].stream = c.stream;
      return k;
    }
  return -1;
}
int tc_prepare(int **dstatus) {
  using namespace ob2::tc;
  Ctx &c = cx();
  if (!c.tc_status) {
    CK(cudaHostAlloc((void **)&c.tc_status, 16 * sizeof(int), cudaHostAllocMapped));
    memset(c.tc_status, 0, 16 * sizeof(int));
  }
  for (int k = 0; k < 8; k++) c.tc_status[2 * k + 1] = c.tc_wait_limit_ms;
  if (!c.tc_attr_set) {
    CK(cudaFuncSetAttribute(apply_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg<1>::kSmemBytes));
    CK(cudaFuncSetAttribute(apply_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg<2>::kSmemBytes));
    CK(cudaFuncSetAttribute(update_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kUSmemBytes));
    CK(cudaFuncSetAttribute(update_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kUSmemBytes));
    CK(cudaFuncSetAttribute(update_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kU2SmemBytes));
    c.tc_attr_set = true;
  }
  const int slot = tc_ws_slot();
  if (slot < 0) return fail("tensor-core workspace: more than 8 streams in use");
  c.tc_cur_slot = slot;
  if (reinterpret_cast<volatile int *>(c.tc_status)[2 * slot] != 0) {
    // a kernel of an earlier call on this stream gave up and nobody collected it: the kernels of this
    // call would bail out of their waits at once -- report it now, with a clean slate afterwards
    CK(cudaStreamSynchronize(c.stream));
    c.tc_status[2 * slot] = 0;
    return fail("tensor-core path: an earlier launch on this stream exceeded its barrier time limit and was never "
                "collected with oblas_b200_sync; its result was invalid (status cleared, call again)");
  }
  int *d = nullptr;
  CK(cudaHostGetDevicePointer((void **)&d, c.tc_status, 0));
  *dstatus = d + 2 * slot;
  return 0;
}
int tc_ws_reserve(size_t need) {
  Ctx &c = cx();
  const int k = tc_ws_slot();
  if (k < 0) return fail("tensor-core workspace: more than 8 streams in use");
  Ctx::TcWs *slot = &c.tc_slots[k];
  if (need > slot->bytes) {
    if (slot->p) {
      CK(cudaStreamSynchronize(c.stream));
      CK(cudaFree(slot->p));
      slot->p = nullptr;
      slot->bytes = 0;
    }
    if (cudaMalloc(&slot->p, need) != cudaSuccess) {
      cudaGetLastError();
      slot->p = nullptr;
      c.tc_ws = nullptr;
      c.tc_ws_bytes = 0;
      return 1;   // caller shrinks its request
    }
    slot->bytes = need;
  }
  c.tc_ws = slot->p;
  c.tc_ws_bytes = slot->bytes;
  return 0;
}
// OBLAS_B200_TC_DEBUG switches parts of the tensor-core kernels off (timing experiments: results are
// WRONG).  Only honoured by builds made for that purpose: make EXTRA=-DOB2_TC_TIMING_EXPERIMENTS
static int tc_debug_switches() {
#ifdef OB2_TC_TIMING_EXPERIMENTS
  static const int v = getenv("OBLAS_B200_TC_DEBUG") ? atoi(getenv("OBLAS_B200_TC_DEBUG")) : 0;
  return v | cx().tc_inject;
#else
  return cx().tc_inject;
#endif
}
// OBLAS_B200_TC_PAIR=1: the CTA-pair variant of the update kernel (fixed tile geometry only)
static bool tc_pair_enabled() {
  static const bool pair = getenv("OBLAS_B200_TC_PAIR") && atoi(getenv("OBLAS_B200_TC_PAIR")) != 0;
  return pair;
}
// Column tiles of the update launches in the EVEN geometry (apply_tc.cuh: TileGeo) unless OBLAS_B200_TC_EVEN=0
// (same-box A/B runs) or the CTA-pair kernel is selected
static int tc_even_tiles() {
  static const bool off = getenv("OBLAS_B200_TC_EVEN") && atoi(getenv("OBLAS_B200_TC_EVEN")) == 0;
  return (off || tc_pair_enabled()) ? 0 : 1;
}
// persistent launch of the tcgen05 GEMM kernel: one CTA per SM.  nh = half tiles (240 byte
// columns) per tile: 2 = one accumulator set, long inner dimension; 1 = double-buffered accumulators
int tc_launch(const ob2::tc::ApplyTcArgs &a, int nh) {
  using namespace ob2::tc;
  const uint64_t units64 = (uint64_t)a.nsys * a.ntiles * a.itiles;
  if (units64 >= (1ull << 32)) return fail("apply (tcgen05): too many tiles in one launch");
  uint32_t grid = (uint32_t)cx().sm_count;
  if (units64 < grid) grid = (uint32_t)units64;
  if (grid == 0) return 0;
  if (nh == 2) {
    apply_tc_kernel<2><<<grid, apply_threads<2>(), Cfg<2>::kSmemBytes, cx().stream>>>(a);
  } else if (nh == 1) {
    apply_tc_kernel<1><<<grid, apply_threads<1>(), Cfg<1>::kSmemBytes, cx().stream>>>(a);
  } else {   // rank-128 update with the A operand resident per (system, row tile)
    const uint64_t nunits = (uint64_t)a.nsys * a.itiles;
    grid = (uint32_t)cx().sm_count;
    if (nunits < grid) grid = (uint32_t)nunits;
    // OBLAS_B200_TC_PAIR=1: the CTA-pair variant (tcgen05 cta_group::2).  Bit-exact (the whole GPU suite
    // passes on it) but as built 1.75x slower than the single-CTA kernel -- see DESIGN.md 4.5 -- so off by default.
    const bool pair = tc_pair_enabled();
    if (pair && !a.prof && !a.units) {
      // CTA pairs (cta_group::2): a cluster of 2 per pair of row tiles
      const uint64_t npairs = (uint64_t)a.nsys * ((a.itiles + 1) / 2);
      grid = (uint32_t)cx().sm_count & ~1u;
      // a persistent kernel must not launch more clusters than can be resident at once
      int &max_clusters = cx().max_clusters;
      if (max_clusters < 0) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = kU2SmemBytes;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        if (cudaOccupancyMaxActiveClusters(&max_clusters, update_tc2_kernel, &cfg) != cudaSuccess) {
          cudaGetLastError();
          max_clusters = (int)grid / 2;
        }
        if (getenv("OBLAS_B200_TRACE")) fprintf(stderr, "oblas_b200: update_tc2_kernel: %d clusters of 2 resident\n", max_clusters);
      }
      if (max_clusters > 0 && (uint32_t)(2 * max_clusters) < grid) grid = (uint32_t)(2 * max_clusters);
      if (2 * npairs < grid) grid = (uint32_t)(2 * npairs);
      update_tc2_kernel<<<grid, kThreads, kU2SmemBytes, cx().stream>>>(a);
    } else if (a.prof) update_tc_kernel<true><<<grid, kUThreads, kUSmemBytes, cx().stream>>>(a);
    else update_tc_kernel<false><<<grid, kUThreads, kUSmemBytes, cx().stream>>>(a);
  }
  return after_launch("apply (tcgen05)");
}
int apply_tc_run(const uint8_t *C, size_t c_stride, uint32_t M, uint32_t Kc, const uint8_t *D, size_t d_stride,
                 uint8_t *Y, size_t y_stride, size_t nbytes, int accumulate, unsigned poly, int d_bit_rows = 0) {
  using namespace ob2::tc;
  int *dstatus = nullptr;
  if (tc_prepare(&dstatus)) return -1;
  const uint32_t ksteps = ksteps_for(Kc, Cfg<2>::kSPS);
  constexpr int kTileN = Cfg<2>::kTileN;
  const size_t tile_bytes = (size_t)ksteps * 2 * kHBlock;
  const uint32_t ntiles_all = (uint32_t)((nbytes + kTileN - 1) / kTileN);
  uint32_t tiles_per_pass = (uint32_t)(cx().tc_ws_budget / tile_bytes);
  if (tiles_per_pass == 0) tiles_per_pass = 1;
  if (tiles_per_pass > ntiles_all) tiles_per_pass = ntiles_all;
  for (;;) {   // shrink the column window until the workspace fits
    int rc = tc_ws_reserve((size_t)tiles_per_pass * tile_bytes);
    if (rc < 0) return -1;
    if (rc == 0) break;
    if (tiles_per_pass == 1) return fail("apply: cannot allocate %zu bytes of operand workspace", tile_bytes);
    tiles_per_pass = (tiles_per_pass + 1) / 2;
  }
  const uint32_t itiles = (M + kRowsI - 1) / kRowsI;
  for (uint32_t t0 = 0; t0 < ntiles_all; t0 += tiles_per_pass) {
    const uint32_t nt = (ntiles_all - t0 < tiles_per_pass) ? (ntiles_all - t0) : tiles_per_pass;
    const uint32_t n0 = t0 * kTileN;
    const uint32_t ncols = (uint32_t)((nbytes - n0 < (size_t)nt * kTileN) ? (nbytes - n0) : (size_t)nt * kTileN);
    ExpandArgs e;
    memset(&e, 0, sizeof e);
    e.reg[0].base = const_cast<uint8_t *>(D); e.reg[0].rs = d_stride; e.reg[0].ss = 0;
    e.reg[0].col0 = n0; e.reg[0].ncols = ncols; e.reg[0].v0 = 0;
    e.nreg = 1; e.Kc = Kc; e.rowmap = nullptr; e.rowmap_stride = 0; e.nrows = nullptr; e.bit_rows = d_bit_rows;
    e.Dx = (uint8_t *)cx().tc_ws; e.ksteps = ksteps; e.nhalf = 2 * nt; e.nsys = 1;
    const size_t work = (size_t)nt * ksteps * (kTileN / 4);
    uint32_t eg = (uint32_t)((work + 255) / 256);
    if (eg > (uint32_t)cx().sm_count * 16) eg = (uint32_t)cx().sm_count * 16;
    expand_d_kernel<<<eg, 256, 0, cx().stream>>>(e);
    if (after_launch("apply (operand expansion)")) return -1;
    ApplyTcArgs a;
    memset(&a, 0, sizeof a);
    a.C = C; a.c_stride = c_stride; a.c_ss = 0; a.M = M; a.Kc = Kc; a.Dx = (const uint8_t *)cx().tc_ws;
    a.ksteps = ksteps; a.ntiles = nt; a.itiles = itiles; a.nsys = 1;
    a.reg[0].base = Y; a.reg[0].rs = y_stride; a.reg[0].ss = 0; a.reg[0].col0 = n0; a.reg[0].ncols = ncols;
    a.reg[0].v0 = 0; a.nreg = 1;
    a.accumulate = accumulate; a.status = dstatus; a.poly = poly;
    a.dbg = tc_debug_switches();
    if (tc_launch(a, 2)) return -1;
  }
  return 0;
}


// ---- diagnostics: where the roles of update_tc_kernel spend their cycles ------------------
extern "C" int oblas_b200_update_phases(int enable, uint64_t *out) {
  if (ensure_init()) return -1;
  std::lock_guard<std::mutex> lk(cx().mu);
  if (out && cx().d_uprof) {
    CK(cudaStreamSynchronize(cx().stream));
    CK(cudaMemcpy(out, cx().d_uprof, 16 * sizeof(uint64_t), cudaMemcpyDeviceToHost));
  }
  if (enable) {
    if (!cx().d_uprof) CK(cudaMalloc(&cx().d_uprof, 16 * sizeof(uint64_t)));
    CK(cudaMemset(cx().d_uprof, 0, 16 * sizeof(uint64_t)));
  } else if (cx().d_uprof) {
    CK(cudaFree(cx().d_uprof));
    cx().d_uprof = nullptr;
  }
  return 0;
}

// ---- diagnostics: per-kernel-class device times of the tensor-core elimination ----------
// classes: 0 outer-panel factorisation, 1 pivot-row expansion, 2 tcgen05 update, 3 permutation
struct TcTimer {
  int cls;
  cudaEvent_t a = nullptr, b = nullptr;
  explicit TcTimer(int c) : cls(c) {
    if (!cx().tc_timing) return;
    if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess) { a = b = nullptr; return; }
    cudaEventRecord(a, cx().stream);
  }
  ~TcTimer() {
    if (!a) return;
    cudaEventRecord(b, cx().stream);
    cx().tc_events.push_back({cls, {a, b}});
  }
};
// classes of the _ex form: 0 outer-panel factorisation (lazy: the compact pass), 1 pivot-row expansion, 2 tcgen05
// update, 3 permutation, 4 lazy: the full kernel's pass over the systems the compact pass gave up.  The 4-entry
// form adds class 4 to class 0.
static int solve_tc_times_impl(int enable, double *ms_out, uint64_t *launches_out, int nclasses);
extern "C" int oblas_b200_solve_tc_times(int enable, double *ms_out, uint64_t *launches_out) {
  return solve_tc_times_impl(enable, ms_out, launches_out, 4);
}
extern "C" int oblas_b200_solve_tc_times_ex(int enable, double *ms_out, uint64_t *launches_out, int nclasses) {
  if (nclasses < 1 || nclasses > 5) return fail("solve_tc_times_ex: 1..5 classes");
  return solve_tc_times_impl(enable, ms_out, launches_out, nclasses);
}
// lazy outer panels so far on this context: out[0] = (system, outer panel) pairs factorised by the compact pass,
// out[1] = pairs it gave up (redone over all rows); reset != 0 clears the counters afterwards
extern "C" int oblas_b200_lazy_stats(uint64_t *out, int reset) {
  if (ensure_init()) return -1;
  std::lock_guard<std::mutex> lk(cx().mu);
  uint64_t v[2] = {0, 0};
  if (cx().d_lazy_stats) {
    CK(cudaStreamSynchronize(cx().stream));
    CK(cudaMemcpy(v, cx().d_lazy_stats, sizeof v, cudaMemcpyDeviceToHost));
    if (reset) CK(cudaMemset(cx().d_lazy_stats, 0, sizeof v));
  }
  if (out) { out[0] = v[0]; out[1] = v[1]; }
  return 0;
}
static int solve_tc_times_impl(int enable, double *ms_out, uint64_t *launches_out, int nclasses) {
  if (ensure_init()) return -1;
  std::lock_guard<std::mutex> lk(cx().mu);
  double ms[5] = {0, 0, 0, 0, 0};
  uint64_t n[5] = {0, 0, 0, 0, 0};
  if (!cx().tc_events.empty()) CK(cudaDeviceSynchronize());
  for (auto &e : cx().tc_events) {
    float t = 0;
    if (cudaEventElapsedTime(&t, e.second.first, e.second.second) == cudaSuccess) {
      ms[e.first] += t;
      n[e.first]++;
    }
    cudaEventDestroy(e.second.first);
    cudaEventDestroy(e.second.second);
  }
  cx().tc_events.clear();
  if (nclasses < 5) {
    ms[0] += ms[4];
    n[0] += n[4];
  }
  for (int k = 0; k < nclasses; k++) {
    if (ms_out) ms_out[k] = ms[k];
    if (launches_out) launches_out[k] = n[k];
  }
  cx().tc_timing = enable != 0;
  return 0;
}

// ---- tensor-core elimination: host side ------------------------------------------------
// Outer panels of 128 columns.  Per outer panel and chunk of systems:
//   panel_kernel    (solve.cuh)    factorises the 128-column window in shared memory, leaves the
//                                  coefficient matrix E (rows x 128) and the pivot rows
//   expand_d_kernel (apply_tc.cuh) old contents of the pivot rows, A's remaining columns and B,
//                                  -> fp4 operand blocks
//   apply_tc_kernel                [A_rest | B] ^= E * P_old  on tcgen05
// then permute_kernel puts the rows into position order and writes the ranks.
// workspace bytes of one system on the tensor-core elimination path
// ... of which the lazy outer panels: gathered windows + coefficient rows of the compact set, its row list, the flag,
// the two unit lists
static size_t solve_tc_lazy_ws_per_system(uint32_t rows) {
  const size_t cmax = (size_t)cx().tc_lazy_rows;
  if (cmax == 0) return 0;
  return 2 * cmax * ob2::kOuterBytes + cmax * 2 + 16 + (cmax / 16 + (rows + 15) / 16) * 4;
}
static size_t solve_tc_ws_per_system(uint32_t rows, uint32_t a_bytes, uint32_t b_bytes) {
  using namespace ob2::tc;
  const uint32_t a_rest0 = a_bytes > (uint32_t)kOuterBytes ? a_bytes - kOuterBytes : 0;
  const uint32_t ntiles_max = (a_rest0 + b_bytes + kHalfN - 1) / kHalfN;
  return (size_t)ntiles_max * kUSteps * kHBlock + (size_t)rows * kOuterBytes + ((size_t)rows * 2 + 15) / 16 * 16 +
         2 * kOuterBytes + 16 + solve_tc_lazy_ws_per_system(rows);
}
// How a batch would run: *path = 1 tensor-core elimination (outer panels + tcgen05 updates), 0 = shared-memory
// table kernel; *chunk_systems = systems per chunk of the tensor-core path (the whole batch on the table kernel).
extern "C" int oblas_b200_solve_plan(int field, uint32_t rows, uint32_t cols, size_t a_row_bytes, size_t b_row_bytes,
                                     size_t nsys, int *path, size_t *chunk_systems, int *geometry) {
  if (ensure_init()) return -1;
  if (field < OB2_GF2 || field > OB2_GF256) return fail("solve_plan: bad field %d", field);
  SolveArgs p;
  memset(&p, 0, sizeof p);
  p.rows = rows; p.cols = cols; p.a_bytes = (uint32_t)a_row_bytes; p.b_bytes = (uint32_t)b_row_bytes;
  std::lock_guard<std::mutex> lk(cx().mu);
  const bool tc = solve_tc_eligible(field, p, nsys);
  if (path) *path = tc ? 1 : 0;
  size_t chunk = nsys;
  if (tc) {
    chunk = cx().tc_ws_budget / solve_tc_ws_per_system(rows, p.a_bytes, p.b_bytes);
    if (chunk < 1) chunk = 1;
    if (chunk > nsys) chunk = nsys;
  }
  if (chunk_systems) *chunk_systems = chunk;
  if (geometry) {
    geometry[0] = geometry[1] = geometry[2] = 0;
    if (tc) {   // the outer-panel kernel: 128-bit inner panels, 6-bit groups, one 128-byte tile
      geometry[0] = 4; geometry[1] = 6; geometry[2] = 128;
    } else {
      const int geo = cx().solve_geo == 3 ? -1 : cx().solve_geo;
      const bool fits = plan_solve(field, geo, rows, cx().smem_optin, geometry);
      const int ggeo = geo == 4 ? 4 : 1;
      if (cx().solve_tall && (geo < 0 || geo == 1 || geo == 4) &&
          (!fits || geometry[0] != 4 || (cx().solve_tall == 2 && geo < 0 && geometry[1] != 6)) && solve_grows_fits(rows, ggeo)) {
        geometry[0] = 4; geometry[1] = ggeo == 4 ? 5 : 6; geometry[2] = 128;   // per-row state in global memory
      } else if (!fits) {
        return fail("solve_plan: %u rows do not fit the shared-memory panel (limit %zu bytes); no fallback", rows, cx().smem_optin);
      }
    }
  }
  return 0;
}
int solve_tc_run(SolveArgs &p, size_t nsys) {
  using namespace ob2::tc;
  int *dstatus = nullptr;
  if (tc_prepare(&dstatus)) return -1;
  const uint32_t rows = p.rows;
  const uint32_t nouter = (p.cols + kOuterBytes - 1) / kOuterBytes;
  const uint32_t ksteps = ksteps_for(kOuterBytes, Cfg<1>::kSPS);   // 16 = kUSteps
  const uint32_t itiles = (rows + kRowsI - 1) / kRowsI;
  // column tiles of the widest trailing update (outer panel 0)
  constexpr int kTileN = Cfg<1>::kTileN;   // 240: double-buffered accumulators
  auto tiles_of = [](uint32_t bytes) { return (bytes + kTileN - 1) / kTileN; };
  const uint32_t a_rest0 = p.a_bytes > (uint32_t)kOuterBytes ? p.a_bytes - kOuterBytes : 0;
  const uint32_t ntiles_max = tiles_of(a_rest0 + (p.B ? p.b_bytes : 0));
  const size_t dx_per_sys = (size_t)ntiles_max * ksteps * kHBlock;
  const size_t e_per_sys = (size_t)rows * kOuterBytes;
  const size_t map_per_sys = ((size_t)rows * 2 + 15) / 16 * 16;
  const size_t small_per_sys = 2 * kOuterBytes + 16;   // piv_all (u16 x 128) + tcount + npiv + pad
  // lazy outer panels (header of panel_kernel, solve.cuh): compact set of cmax rows per system
  // (default setting: only for batches of at least 2 x (SM count) systems -- the compact pass needs two CTAs per SM
  // to overlap their look-ahead chains; the 148-system chunks of the host-streaming call take the full kernel)
  const uint32_t cmax = (rows >= (uint32_t)kOuterBytes && (!cx().tc_lazy_auto || nsys >= 2 * (size_t)cx().sm_count))
                            ? (uint32_t)cx().tc_lazy_rows : 0u;
  const bool lazy = cmax != 0;
  const size_t wc_per_sys = (size_t)cmax * kOuterBytes;
  const size_t lazy_per_sys = solve_tc_lazy_ws_per_system(rows);
  const size_t per_sys = dx_per_sys + e_per_sys + map_per_sys + small_per_sys + lazy_per_sys;
  // tall systems: the per-row state of the outer-panel kernel lives in global memory, one block per CTA
  const bool grows = panel_smem_bytes(rows) > cx().smem_optin;
  const size_t row_ws_cta = grows ? panel_row_ws_bytes(rows) : 0;
  const size_t row_ws_all = row_ws_cta * (size_t)cx().sm_count;
  size_t chunk = cx().tc_ws_budget / per_sys;
  if (chunk < 1) chunk = 1;
  if (chunk > nsys) chunk = nsys;
  if (lazy && chunk > 65535) chunk = 65535;   // unit lists: (system << 16 | row tile)
  for (;;) {
    int rc = tc_ws_reserve(chunk * per_sys + 256 + row_ws_all + 256 + 256);
    if (rc < 0) return -1;
    if (rc == 0) break;
    if (chunk == 1) return fail("solve: cannot allocate %zu bytes of workspace", per_sys);
    chunk = (chunk + 1) / 2;
  }
  uint8_t *ws = (uint8_t *)cx().tc_ws;
  uint8_t *Dx = ws;                                        ws += chunk * dx_per_sys;
  uint8_t *E = ws;                                         ws += chunk * e_per_sys;
  uint16_t *maps = (uint16_t *)ws;                         ws += chunk * map_per_sys;
  uint16_t *piv_all = (uint16_t *)ws;                      ws += chunk * 2 * kOuterBytes;
  uint32_t *tcount = (uint32_t *)ws;                       ws += chunk * 4;
  uint32_t *npiv = (uint32_t *)ws;                         ws += chunk * 4;
  uint8_t *Wc = nullptr, *Ec = nullptr, *flags = nullptr;
  uint16_t *crow = nullptr;
  uint32_t *unitsA = nullptr, *unitsB = nullptr, *ncount = nullptr;
  if (lazy) {
    ws = (uint8_t *)(((uintptr_t)ws + 15) & ~(uintptr_t)15);
    Wc = ws;                                               ws += chunk * wc_per_sys;
    Ec = ws;                                               ws += chunk * wc_per_sys;
    unitsA = (uint32_t *)ws;                               ws += chunk * (cmax / 16) * 4;
    unitsB = (uint32_t *)ws;                               ws += chunk * (size_t)itiles * 4;
    ncount = (uint32_t *)ws;                               ws += 16;
    crow = (uint16_t *)ws;                                 ws += chunk * (size_t)cmax * 2;
    flags = ws;                                            ws += (chunk + 15) / 16 * 16;
  }
  uint8_t *row_ws = grows ? (uint8_t *)(((uintptr_t)ws + 255) & ~(uintptr_t)255) : nullptr;
  // permutation kernel: row map + staging area in shared memory
  const size_t pmap = ((size_t)rows * 2 + 15) / 16 * 16;
  // stage = count + list of moved rows (u16 x rows) + at least 16 bytes per row
  if (pmap + 64 + 16 + pmap + (size_t)rows * 16 > cx().smem_optin) return fail("solve: %u rows do not fit the permutation stage", rows);
  const uint32_t stage_bytes = (uint32_t)((cx().smem_optin - pmap - 64) & ~(size_t)15);
  if (!cx().perm_attr_set) {
    CK(cudaFuncSetAttribute(permute_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cx().smem_optin));
    cx().perm_attr_set = true;
  }
  for (size_t first = 0; first < nsys; first += chunk) {
    const uint32_t cnt = (uint32_t)((nsys - first < chunk) ? (nsys - first) : chunk);
    uint8_t *A = p.A + first * p.a_ss;
    uint8_t *B = p.B ? p.B + first * p.b_ss : nullptr;
    const uint32_t grid = cnt < (uint32_t)cx().sm_count ? cnt : (uint32_t)cx().sm_count;
    for (uint32_t outer = 0; outer < nouter; outer++) {
      PanelArgs pa;
      memset(&pa, 0, sizeof pa);
      pa.A = A; pa.a_rs = p.a_rs; pa.a_ss = p.a_ss; pa.rows = rows; pa.cols = p.cols; pa.a_bytes = p.a_bytes;
      pa.E = E; pa.pos2phys = maps; pa.tcount = tcount; pa.piv_all = piv_all; pa.npiv = npiv;
      pa.pivcols = p.pivcols ? p.pivcols + first * rows : nullptr;
      pa.nsys = cnt; pa.outer = outer; pa.inv = p.inv; pa.red = p.red; pa.no_fast_panel = p.no_fast_panel;
      pa.no_look_ahead = p.no_look_ahead;
      pa.phase = cx().d_phase;
      pa.row_ws = row_ws; pa.row_ws_cta = row_ws_cta;
      // Lazy outer panel (a whole 128-column window): the compact kernel factorises the windows of the first cmax
      // unpivoted rows only and leaves two unit lists; the systems it gives up (a column without a pivot among
      // those rows: rank deficiency, or pivots far down) are redone by the full kernel, which puts all their row
      // tiles on list B.  Otherwise (narrow last panel, lazy off): the full kernel for every system.
      const uint32_t win_end = (outer + 1) * kOuterBytes;
      const bool lazy_panel = lazy && win_end <= p.cols && win_end <= p.a_bytes;
      if (lazy_panel) {
        CK(cudaMemsetAsync(ncount, 0, 16, cx().stream));
        PanelArgs pc = pa;
        pc.cmax = cmax; pc.Wc = Wc; pc.Ec = Ec; pc.crow = crow; pc.flags = flags;
        pc.unitsA = unitsA; pc.nunitsA = ncount; pc.unitsB = unitsB; pc.nunitsB = ncount + 1;
        pc.row_ws = nullptr; pc.row_ws_cta = 0;
        if (!cx().d_lazy_stats) {
          CK(cudaMalloc(&cx().d_lazy_stats, 16));
          CK(cudaMemset(cx().d_lazy_stats, 0, 16));
        }
        pc.stats = cx().d_lazy_stats;
        {
          TcTimer tt(0);
          // The compact pass is bound by the look-ahead chain (16 dependent column eliminations per step on two
          // warps: ~25 k clk per step, against ~6 k clk of tile pass over 192 rows), not by table look-ups: with
          // the 4-bit tables (64 KB) and 256 threads two CTAs share an SM and their chains overlap.
          // OBLAS_B200_LAZY_GEO=6: one CTA of 512 threads per SM with the 6-bit tables (A/B runs).
          static const bool geo6 = getenv("OBLAS_B200_LAZY_GEO") && atoi(getenv("OBLAS_B200_LAZY_GEO")) == 6;
          // (measured: 256 threads x 2 CTAs per SM 3.73 ms per 592 systems; 192 x 2: 3.78; 160 x 3: 4.03; 128 x 4: 4.14)
          const uint32_t grid2 = cnt < 2u * (uint32_t)cx().sm_count ? cnt : 2u * (uint32_t)cx().sm_count;
          int rc = geo6 ? launch_panel_t<6, 128>(pc, grid, cx().smem_optin, cx().stream)
                        : launch_panel_t<4, 128>(pc, grid2, cx().smem_optin, cx().stream, 256);
          if (rc) return fail("solve (tensor-core path): compact panel kernel launch configuration failed (%d)", rc);
        }
        if (after_launch("solve (compact outer panel)")) return -1;
        pa.flags = flags; pa.only_flagged = 1; pa.unitsB = unitsB; pa.nunitsB = ncount + 1;
      }
      // the row map lives in global memory as [sys][rows] u16 with a 16-byte aligned pitch
      {
        TcTimer tt(lazy_panel ? 4 : 0);
        int rc = launch_panel_t<6, 128>(pa, grid, cx().smem_optin, cx().stream);
        if (rc) return fail("solve (tensor-core path): panel kernel launch configuration failed (%d)", rc);
      }
      if (after_launch("solve (outer panel)")) return -1;
      const uint32_t a_col0 = (outer + 1) * kOuterBytes;
      const uint32_t a_rest = p.a_bytes > a_col0 ? p.a_bytes - a_col0 : 0;
      const uint32_t b_cols = B ? p.b_bytes : 0;
      if (a_rest == 0 && b_cols == 0) continue;
      TcRegion reg[2];
      memset(reg, 0, sizeof reg);
      // A's remaining columns and B form one range of virtual columns; tiles are cut from the range
      uint32_t nreg = 0, vcols = 0;
      if (a_rest) {
        reg[nreg].base = A; reg[nreg].rs = p.a_rs; reg[nreg].ss = p.a_ss; reg[nreg].col0 = a_col0;
        reg[nreg].ncols = a_rest; reg[nreg].v0 = vcols;
        vcols += a_rest;
        nreg++;
      }
      if (b_cols) {
        reg[nreg].base = B; reg[nreg].rs = p.b_rs; reg[nreg].ss = p.b_ss; reg[nreg].col0 = 0;
        reg[nreg].ncols = b_cols; reg[nreg].v0 = vcols;
        vcols += b_cols;
        nreg++;
      }
      const uint32_t ntiles = tiles_of(vcols);
      ExpandArgs e;
      memset(&e, 0, sizeof e);
      e.reg[0] = reg[0]; e.reg[1] = reg[1]; e.nreg = nreg; e.Kc = kOuterBytes;
      e.rowmap = piv_all; e.rowmap_stride = kOuterBytes; e.nrows = npiv;
      e.Dx = Dx; e.ksteps = ksteps; e.nhalf = ntiles; e.nsys = cnt;
      e.even = tc_even_tiles();
      const size_t work = (size_t)cnt * ntiles * ksteps * (kTileN / 4);
      uint32_t eg = (uint32_t)((work + 255) / 256);
      if (eg > (uint32_t)cx().sm_count * 16) eg = (uint32_t)cx().sm_count * 16;
      ApplyTcArgs a;
      memset(&a, 0, sizeof a);
      a.C = E; a.c_stride = kOuterBytes; a.c_ss = e_per_sys; a.M = rows; a.Kc = kOuterBytes; a.Dx = Dx;
      a.ksteps = ksteps; a.ntiles = ntiles; a.itiles = itiles; a.nsys = cnt;
      a.reg[0] = reg[0]; a.reg[1] = reg[1]; a.nreg = nreg;
      a.accumulate = 1; a.status = dstatus; a.prof = cx().d_uprof;
      a.even_tiles = e.even;
      a.poly = 0x100u | (p.red & 0xffu);
      a.dbg = tc_debug_switches();
      // lazy: launch A -- the compact rows (gathered through crow) ^= their coefficient rows * OLD pivot rows --
      // then the pivot rows are expanded again and launch B -- all other rows ^= their old windows * NEW pivot rows
      for (int pass = lazy_panel ? 0 : 1; pass < 2; pass++) {
        {
          TcTimer tt(1);
          expand_d_kernel<<<eg, 256, 0, cx().stream>>>(e);
        }
        if (after_launch("solve (pivot-row expansion)")) return -1;
        ApplyTcArgs u = a;
        if (lazy_panel && pass == 0) {
          u.C = Ec; u.c_ss = wc_per_sys; u.M = cmax; u.itiles = cmax / kRowsI;
          u.units = unitsA; u.nunits_dev = ncount;
          u.y_rowmap = crow; u.y_rowmap_stride = cmax;
        } else if (lazy_panel) {
          u.units = unitsB; u.nunits_dev = ncount + 1;
        }
        {
          TcTimer tt(2);
          if (tc_launch(u, (!lazy_panel && getenv("OBLAS_B200_TC_STREAM_A")) ? 1 : 0)) return -1;
        }
      }
    }
    PermuteArgs pm;
    memset(&pm, 0, sizeof pm);
    pm.A = A; pm.a_rs = p.a_rs; pm.a_ss = p.a_ss; pm.a_bytes = p.a_bytes;
    pm.B = B; pm.b_rs = p.b_rs; pm.b_ss = p.b_ss; pm.b_bytes = B ? p.b_bytes : 0;
    pm.rows = rows; pm.nsys = cnt; pm.pos2phys = maps; pm.tcount = tcount; pm.rank = p.rank + first;
    pm.stage_bytes = stage_bytes;
    {
      TcTimer tt(3);
      permute_kernel<<<grid, 512, pmap + stage_bytes + 16, cx().stream>>>(pm);
    }
    if (after_launch("solve (row permutation)")) return -1;
  }
  return 0;
}

// ---- apply ---------------------------------------------------------------------
int oblas_b200_gf256_apply(const void *C, size_t c_stride, uint32_t M, uint32_t Kc, const void *D,
                           size_t d_stride, void *Y, size_t y_stride, size_t nbytes, int accumulate) {
  if (ensure_init()) return -1;
  if (M == 0 || nbytes == 0) return 0;
  if (!C || !D || !Y) return fail("apply: null pointer");
  if ((nbytes & 15) || (d_stride & 15) || (y_stride & 15) || (reinterpret_cast<uintptr_t>(D) & 15) ||
      (reinterpret_cast<uintptr_t>(Y) & 15))
    return fail("apply: D, Y, their strides and nbytes must be multiples of 16");
  if (classify(C) != PK_DEVVIS || classify(D) != PK_DEVVIS || classify(Y) != PK_DEVVIS)
    return fail("apply: pointers must be device-visible");
  std::lock_guard<std::mutex> lk(cx().mu);
  // path: apply_geo 0..2 force the shared-memory Four-Russians kernel with that geometry, 3 forces
  // the tensor-core kernel, -1 picks by size (the fp4 expansion of D is amortised over M/16 row tiles)
  const bool tc_ok = Kc > 0 && Kc < (1u << 20);
  bool use_tc = cx().apply_geo == 3 || (cx().apply_geo < 0 && M >= 256 && Kc >= 64 && nbytes >= 480);
  if (use_tc && !tc_ok) {
    if (cx().apply_geo == 3) return fail("apply: tensor-core path needs 0 < Kc < 2^20");
    use_tc = false;
  }
  const uint8_t *inv_unused = nullptr;
  unsigned poly = 0;
  if (field_arith(OB2_GF256, &inv_unused, &poly)) return -1;
  if (use_tc)
    return apply_tc_run((const uint8_t *)C, c_stride, M, Kc, (const uint8_t *)D, d_stride, (uint8_t *)Y,
                        y_stride, nbytes, accumulate, poly);
  ApplyArgs p;
  memset(&p, 0, sizeof p);
  p.C = (const uint8_t *)C; p.c_stride = c_stride; p.M = M; p.Kc = Kc;
  p.D = (const uint8_t *)D; p.d_stride = d_stride;
  p.Y = (uint8_t *)Y; p.y_stride = y_stride; p.nbytes = (uint32_t)nbytes; p.accumulate = accumulate;
  p.red = field_red(8, poly);
  int rc = launch_apply(p, cx().apply_geo == 3 ? -1 : cx().apply_geo, 512, cx().stream);
  if (rc) return fail("apply: launch configuration failed (%d)", rc);
  return after_launch("apply");
}

// ---- mixed-field apply ("next" row f-2): the rows of D are GF(2) rows in binmat layout -------
// Y (M x ncols bytes) = C (M x Kc over GF(256)) * expand(Dbits), expand = every bit as the field element 0 / 1:
// what  for i, j: oaxpy_b32(Y, Dbits row j, i, C[i][j])  leaves (octmat.c:65-68 -> oblas_ref.c:19-28), the
// expansion fused into the operand expansion of the tensor-core apply.
int oblas_b200_gf256_apply_b32(const void *C, size_t c_stride, uint32_t M, uint32_t Kc, const uint32_t *Dbits,
                               size_t d_stride_words, void *Y, size_t y_stride, size_t ncols, int accumulate) {
  if (ensure_init()) return -1;
  if (M == 0 || ncols == 0) return 0;
  if (!C || !Dbits || !Y) return fail("apply_b32: null pointer");
  if ((ncols & 15) || (y_stride & 15) || (reinterpret_cast<uintptr_t>(Y) & 15))
    return fail("apply_b32: Y, its stride and ncols must be multiples of 16");
  if (d_stride_words * 32 < ncols) return fail("apply_b32: d_stride_words * 32 < ncols");
  if (Kc == 0 || Kc >= (1u << 20)) return fail("apply_b32: needs 0 < Kc < 2^20");
  if (classify(C) != PK_DEVVIS || classify(Dbits) != PK_DEVVIS || classify(Y) != PK_DEVVIS)
    return fail("apply_b32: pointers must be device-visible");
  std::lock_guard<std::mutex> lk(cx().mu);
  const uint8_t *inv_unused = nullptr;
  unsigned poly = 0;
  if (field_arith(OB2_GF256, &inv_unused, &poly)) return -1;
  return apply_tc_run((const uint8_t *)C, c_stride, M, Kc, (const uint8_t *)Dbits, d_stride_words * 4, (uint8_t *)Y, y_stride,
                      ncols, accumulate, poly, 1);
}

// ---- several GPUs from one process ---------------------------------
// One host thread and one context per device; no data-path collective for batches of independent
// systems, one coefficient broadcast over NVLink (peer copies, a binary tree) for the column-sharded
// apply.  torch.distributed / NCCL drive the one-process-per-GPU form of the same split (bench.py).
int oblas_b200_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n < kMaxDev ? n : kMaxDev;
}

}  // extern "C"
namespace {
// runs fn(k) for k < ndev on its own host thread bound to device devs[k]; first error wins
template <typename F>
int on_devices(const int *devs, int ndev, F fn) {
  std::vector<std::thread> th;
  std::vector<int> rc((size_t)ndev, 0);
  std::vector<std::string> err((size_t)ndev);
  for (int k = 0; k < ndev; k++)
    th.emplace_back([&, k]() {
      rc[(size_t)k] = oblas_b200_set_device(devs[k]);
      if (rc[(size_t)k] == 0) rc[(size_t)k] = fn(k);
      if (rc[(size_t)k]) err[(size_t)k] = tl_err;
    });
  for (auto &t : th) t.join();
  for (int k = 0; k < ndev; k++)
    if (rc[(size_t)k]) return fail("device %d: %s", devs[k], err[(size_t)k].c_str());
  return 0;
}
int pick_devices(const int *devices, int ndev, std::vector<int> &out) {
  const int have = oblas_b200_device_count();
  if (have <= 0) return fail("no CUDA device");
  if (ndev <= 0) ndev = have;
  out.clear();
  for (int k = 0; k < ndev; k++) {
    const int d = devices ? devices[k] : k;
    if (d < 0 || d >= have) return fail("no CUDA device %d (%d visible)", d, have);
    for (int q : out)
      if (q == d) return fail("device %d listed twice", d);
    out.push_back(d);
  }
  return 0;
}
}  // namespace
extern "C" {

// oblas_b200_solve_batch_host over several GPUs: device k gets the contiguous block of systems
// [k * nsys / ndev, (k + 1) * nsys / ndev) -- independent systems, nothing is exchanged.
int oblas_b200_solve_batch_host_multi(int field, void *A_host, size_t a_row_stride, size_t a_sys_stride,
                                      uint32_t rows, uint32_t cols, size_t a_row_bytes, void *B_host,
                                      size_t b_row_stride, size_t b_sys_stride, size_t b_row_bytes,
                                      int32_t *rank_out, uint32_t *pivcols_out, size_t nsys,
                                      const int *devices, int ndev, unsigned flags) {
  std::vector<int> devs;
  if (pick_devices(devices, ndev, devs)) return -1;
  if (!A_host || !rank_out) return fail("solve_host_multi: A and rank_out are required");
  const int nd = (int)devs.size();
  return on_devices(devs.data(), nd, [&](int k) -> int {
    const size_t lo = nsys * (size_t)k / (size_t)nd, hi = nsys * (size_t)(k + 1) / (size_t)nd;
    if (hi == lo) return 0;
    return oblas_b200_solve_batch_host_ex(field, (uint8_t *)A_host + lo * a_sys_stride, a_row_stride, a_sys_stride, rows, cols,
                                          a_row_bytes, B_host ? (uint8_t *)B_host + lo * b_sys_stride : nullptr, b_row_stride,
                                          b_sys_stride, b_row_bytes, rank_out + lo, pivcols_out ? pivcols_out + lo * rows : nullptr,
                                          hi - lo, 0, flags);
  });
}

// Y = C * D with HOST matrices over several GPUs (BASELINE configs[4]): the symbol-byte columns of D and Y
// are cut into one slice per device (boundaries multiples of 256 bytes); the coefficient matrix goes to the
// first device over PCIe once and from there to the others over NVLink (peer copies along a binary tree:
// ceil(log2 ndev) rounds); every device then runs the tcgen05 apply on its slice.  seconds_out (may be
// NULL, 3 entries): coefficient upload + broadcast, slowest device's H2D + apply + D2H, whole call.
int oblas_b200_gf256_apply_sharded(const void *C_host, size_t c_stride, uint32_t M, uint32_t Kc, const void *D_host,
                                   size_t d_stride, void *Y_host, size_t y_stride, size_t nbytes, const int *devices,
                                   int ndev, double *seconds_out) {
  std::vector<int> devs;
  if (pick_devices(devices, ndev, devs)) return -1;
  if (!C_host || !D_host || !Y_host) return fail("apply_sharded: null pointer");
  if (M == 0 || Kc == 0 || nbytes == 0) return 0;
  if ((nbytes & 15) || c_stride < Kc) return fail("apply_sharded: nbytes must be a multiple of 16 and c_stride >= Kc");
  const int nd = (int)devs.size();
  const size_t cpitch = ((size_t)Kc + 15) & ~(size_t)15;
  std::vector<void *> dC((size_t)nd, nullptr), dD((size_t)nd, nullptr), dY((size_t)nd, nullptr);
  std::vector<double> t_dev((size_t)nd, 0.0);
  auto now = []() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  const double t0 = now();
  // slices: 256-byte aligned boundaries
  const size_t tiles = (nbytes + 255) / 256;
  auto slice = [&](int k, size_t &lo, size_t &hi) {
    lo = std::min(nbytes, (tiles * (size_t)k / (size_t)nd) * 256);
    hi = std::min(nbytes, (tiles * (size_t)(k + 1) / (size_t)nd) * 256);
  };
  int rc = on_devices(devs.data(), nd, [&](int k) -> int {
    size_t lo, hi;
    slice(k, lo, hi);
    CK(cudaMalloc(&dC[(size_t)k], (size_t)M * cpitch));
    if (hi > lo) {
      CK(cudaMalloc(&dD[(size_t)k], (size_t)Kc * (hi - lo)));
      CK(cudaMalloc(&dY[(size_t)k], (size_t)M * (hi - lo)));
    }
    return 0;
  });
  double t_bcast = 0;
  if (rc == 0) {
    // coefficients: host -> first device, then a binary tree of peer copies
    rc = on_devices(devs.data(), 1, [&](int) -> int {
      CK(cudaMemcpy2DAsync(dC[0], cpitch, C_host, c_stride, Kc, M, cudaMemcpyHostToDevice, cx().stream));
      if (cpitch > Kc) CK(cudaMemset2DAsync((uint8_t *)dC[0] + Kc, cpitch, 0, cpitch - Kc, M, cx().stream));
      CK(cudaStreamSynchronize(cx().stream));
      return 0;
    });
    for (int have = 1; have < nd && rc == 0; have *= 2) {
      std::vector<int> src_dev;
      const int cnt = std::min(have, nd - have);
      for (int q = 0; q < cnt; q++) src_dev.push_back(devs[(size_t)q]);
      rc = on_devices(src_dev.data(), cnt, [&](int q) -> int {
        const int dst = have + q;
        int can = 0;
        CK(cudaDeviceCanAccessPeer(&can, devs[(size_t)q], devs[(size_t)dst]));
        if (can) {
          cudaError_t e = cudaDeviceEnablePeerAccess(devs[(size_t)dst], 0);
          if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return fail("cudaDeviceEnablePeerAccess: %s", cudaGetErrorString(e));
          cudaGetLastError();
        }
        CK(cudaMemcpyPeerAsync(dC[(size_t)dst], devs[(size_t)dst], dC[(size_t)q], devs[(size_t)q], (size_t)M * cpitch, cx().stream));
        CK(cudaStreamSynchronize(cx().stream));
        return 0;
      });
    }
    t_bcast = now() - t0;
  }
  if (rc == 0)
    rc = on_devices(devs.data(), nd, [&](int k) -> int {
      size_t lo, hi;
      slice(k, lo, hi);
      if (hi == lo) return 0;
      const double s0 = now();
      const size_t w = hi - lo;
      CK(cudaMemcpy2DAsync(dD[(size_t)k], w, (const uint8_t *)D_host + lo, d_stride, w, Kc, cudaMemcpyHostToDevice, cx().stream));
      if (oblas_b200_gf256_apply(dC[(size_t)k], cpitch, M, Kc, dD[(size_t)k], w, dY[(size_t)k], w, w, 0)) return -1;
      CK(cudaMemcpy2DAsync((uint8_t *)Y_host + lo, y_stride, dY[(size_t)k], w, w, M, cudaMemcpyDeviceToHost, cx().stream));
      if (oblas_b200_sync()) return -1;
      t_dev[(size_t)k] = now() - s0;
      return 0;
    });
  on_devices(devs.data(), nd, [&](int k) -> int {
    cudaFree(dC[(size_t)k]);
    cudaFree(dD[(size_t)k]);
    cudaFree(dY[(size_t)k]);
    return 0;
  });
  if (seconds_out) {
    seconds_out[0] = t_bcast;
    seconds_out[1] = 0;
    for (double t : t_dev) seconds_out[1] = std::max(seconds_out[1], t);
    seconds_out[2] = now() - t0;
  }
  return rc;
}

// ---- row degree -------------------------------------------------------------------
int oblas_b200_nnz_batch(int field, const void *bits, size_t row_stride, uint32_t cols,
                         const uint32_t *rows_idx, size_t nrows, uint32_t s, uint32_t e, uint32_t *out) {
  if (ensure_init()) return -1;
  if (nrows == 0) return 0;
  if (e > cols) e = cols;
  if (s > e) s = e;
  if (classify(bits) != PK_DEVVIS || classify(rows_idx) != PK_DEVVIS || classify(out) != PK_DEVVIS)
    return fail("nnz_batch: pointers must be device-visible");
  std::lock_guard<std::mutex> lk(cx().mu);
  unsigned blocks = (unsigned)((nrows * 32 + 255) / 256);
  const uint8_t *p = (const uint8_t *)bits;
  switch (field) {
  case OB2_GF2: nnz_kernel<1><<<blocks, 256, 0, cx().stream>>>(p, row_stride, rows_idx, (uint32_t)nrows, s, e, out); break;
  case OB2_GF4: nnz_kernel<2><<<blocks, 256, 0, cx().stream>>>(p, row_stride, rows_idx, (uint32_t)nrows, s, e, out); break;
  case OB2_GF16: nnz_kernel<4><<<blocks, 256, 0, cx().stream>>>(p, row_stride, rows_idx, (uint32_t)nrows, s, e, out); break;
  case OB2_GF256: nnz_kernel<8><<<blocks, 256, 0, cx().stream>>>(p, row_stride, rows_idx, (uint32_t)nrows, s, e, out); break;
  default: return fail("nnz_batch: bad field %d", field);
  }
  return after_launch("nnz");
}

// ---- tables for other field polynomials ("next" row f-4) -------------------------------------
// Split-nibble tables in the layout of the reference's generator (tablegen.c:66-106):
//   8-bit elements: LO[u][j] = u*j, HI[u][j] = u*(j<<4);  4-bit: LO[u][j] = u*j, HI = LO<<4;
//   2-bit: LO[u][j] = (u*(j>>2))<<2 | u*(j&3), HI = LO<<4 (the corrected form).
// poly includes the leading term (285 = x^8+x^4+x^3+x^2+1 is the reference's GF(256) polynomial).
int oblas_b200_make_tables(unsigned exp_bits, unsigned poly, uint8_t *lo, uint8_t *hi, uint8_t *inv) {
  if (exp_bits != 2 && exp_bits != 4 && exp_bits != 8) return fail("make_tables: element width must be 2, 4 or 8 bits");
  if (!lo || !hi || !inv) return fail("make_tables: null pointer");
  if ((poly >> exp_bits) != 1u) return fail("make_tables: polynomial 0x%x is not of degree %u", poly, exp_bits);
  const unsigned len = 1u << exp_bits;
  auto mul = [&](unsigned a, unsigned b) { return (unsigned)gf_mul_slow(exp_bits, poly, a, b); };
  for (unsigned u = 0; u < len; u++) {
    inv[u] = 0;
    for (unsigned v = 1; v < len && u; v++)
      if (mul(u, v) == 1) inv[u] = (uint8_t)v;
    if (u && !inv[u]) return fail("make_tables: 0x%x is not irreducible (element %u has no inverse)", poly, u);
    for (unsigned j = 0; j < 16; j++) {
      uint8_t l, h;
      if (exp_bits == 8) {
        l = (uint8_t)mul(u, j);
        h = (uint8_t)mul(u, j << 4);
      } else if (exp_bits == 4) {
        l = (uint8_t)mul(u, j);
        h = (uint8_t)(l << 4);
      } else {
        l = (uint8_t)((mul(u, j >> 2) << 2) | mul(u, j & 3));
        h = (uint8_t)(l << 4);
      }
      lo[u * 16 + j] = l;
      hi[u * 16 + j] = h;
    }
  }
  return 0;
}
// Which table set the elimination and apply kernels use for `field` from now on (this context): 0 / 1 = the
// reference's polynomial (GF(4): always the corrected, true-field table), a handle of
// oblas_b200_register_tables = that field polynomial.  The handle must be for the field's element width and
// its tables must be those of a field (checked at registration: regenerated from the polynomial they imply).
int oblas_b200_set_field_tables(int field, int variant) {
  if (ensure_init()) return -1;
  if (field < OB2_GF4 || field > OB2_GF256) return fail("set_field_tables: field must be GF(4), GF(16) or GF(256)");
  std::lock_guard<std::mutex> lk(cx().mu);
  if (variant >= 16) {
    const size_t ci = (size_t)variant - 16;
    if (ci >= cx().d_custom.size()) return fail("set_field_tables: unknown table handle %d", variant);
    if (cx().custom_bits[ci] != (int)kBits[field]) return fail("set_field_tables: handle %d is for %d-bit elements", variant, cx().custom_bits[ci]);
    if (!cx().custom_poly[ci]) return fail("set_field_tables: the tables of handle %d are not those of a field polynomial", variant);
  } else if (variant < 0) {
    return fail("set_field_tables: bad variant %d", variant);
  } else {
    variant = 0;
  }
  cx().field_variant[field] = variant;
  return 0;
}
// Makes a table set usable by oblas_b200_rowops: returns a handle to pass as `variant` (>= 16) with
// the field of the same element width, or a negative code.  lo/hi: 16 * 2^exp_bits bytes, inv: 2^exp_bits.
int oblas_b200_register_tables(unsigned exp_bits, const uint8_t *lo, const uint8_t *hi, const uint8_t *inv) {
  if (ensure_init()) return -1;
  if (exp_bits != 2 && exp_bits != 4 && exp_bits != 8) return fail("register_tables: element width must be 2, 4 or 8 bits");
  if (!lo || !hi || !inv) return fail("register_tables: null pointer");
  std::lock_guard<std::mutex> lk(cx().mu);
  DevTableSet h;
  fill_table_set(h, lo, hi, inv, 1u << exp_bits);
  DevTableSet *d = nullptr;
  CK(cudaMalloc(&d, sizeof(DevTableSet)));
  CK(cudaMemcpy(d, &h, sizeof h, cudaMemcpyHostToDevice));
  cx().d_custom.push_back(d);
  cx().custom_linear.push_back(h.linear);
  cx().custom_bits.push_back((int)exp_bits);
  {
    // are these the tables of a field?  x * x^(e-1) gives the polynomial; the generator must reproduce all of them
    const unsigned e = exp_bits, n = 1u << e;
    unsigned low = (e == 8) ? hi[2 * 16 + 8] : (e == 4) ? (unsigned)lo[2 * 16 + 8] : ((unsigned)lo[2 * 16 + 2] & 3u);
    const unsigned poly = n | low;
    std::vector<uint8_t> l2(16 * n), h2(16 * n), i2(n);
    unsigned ok_poly = 0;
    if (oblas_b200_make_tables(e, poly, l2.data(), h2.data(), i2.data()) == 0 && !memcmp(l2.data(), lo, 16 * n) &&
        !memcmp(h2.data(), hi, 16 * n) && !memcmp(i2.data(), inv, n))
      ok_poly = poly;
    cx().custom_poly.push_back(ok_poly);
  }
  return 16 + (int)cx().d_custom.size() - 1;
}

// ---- column operations / accessors on device-resident matrices (row f-3) --------------------
#define OB2_FIELD_SWITCH(KERNEL, GRID, BLOCK, ...)                                              \
  switch (field) {                                                                             \
  case OB2_GF2: KERNEL<1><<<GRID, BLOCK, 0, cx().stream>>>(__VA_ARGS__); break;                    \
  case OB2_GF4: KERNEL<2><<<GRID, BLOCK, 0, cx().stream>>>(__VA_ARGS__); break;                    \
  case OB2_GF16: KERNEL<4><<<GRID, BLOCK, 0, cx().stream>>>(__VA_ARGS__); break;                   \
  case OB2_GF256: KERNEL<8><<<GRID, BLOCK, 0, cx().stream>>>(__VA_ARGS__); break;                  \
  default: return fail("bad field %d", field);                                                 \
  }
int oblas_b200_swapcols_batch(int field, void *bits, size_t row_stride, size_t sys_stride, uint32_t rows,
                              const uint32_t *col_i, const uint32_t *col_j, size_t nsys) {
  if (ensure_init()) return -1;
  if (nsys == 0 || rows == 0) return 0;
  if (!bits || !col_i || !col_j) return fail("swapcols: null pointer");
  if (classify(bits) != PK_DEVVIS) return fail("swapcols: matrix must be device-visible");
  if ((row_stride & 3) || (sys_stride & 3) || (reinterpret_cast<uintptr_t>(bits) & 3))
    return fail("swapcols: base / strides must be multiples of 4");
  std::lock_guard<std::mutex> lk(cx().mu);
  const size_t ib = nsys * sizeof(uint32_t), slot = (ib + 255) & ~(size_t)255;
  if (scratch_reserve(2 * slot + 256)) return -1;
  const void *di, *dj;
  if (stage_array(col_i, ib, 0, &di) || stage_array(col_j, ib, slot, &dj)) return -1;
  const size_t total = nsys * (size_t)rows;
  unsigned blocks = (unsigned)((total + 255) / 256);
  if (blocks > (unsigned)cx().sm_count * 32) blocks = (unsigned)cx().sm_count * 32;
  OB2_FIELD_SWITCH(swapcols_kernel, blocks, 256, (uint8_t *)bits, row_stride, sys_stride, rows,
                   (const uint32_t *)di, (const uint32_t *)dj, (uint32_t)nsys)
  return after_launch("swapcols");
}
int oblas_b200_get_elements(int field, const void *bits, size_t row_stride, const uint32_t *rows_idx,
                            const uint32_t *cols_idx, size_t n, uint8_t *out_dev) {
  if (ensure_init()) return -1;
  if (n == 0) return 0;
  if (!bits || !rows_idx || !cols_idx || !out_dev) return fail("get_elements: null pointer");
  if (classify(bits) != PK_DEVVIS || classify(out_dev) != PK_DEVVIS)
    return fail("get_elements: matrix and output must be device-visible");
  if (n > 0x7fffffffu) return fail("get_elements: too many elements");
  std::lock_guard<std::mutex> lk(cx().mu);
  const size_t ib = n * sizeof(uint32_t), slot = (ib + 255) & ~(size_t)255;
  if (scratch_reserve(2 * slot + 256)) return -1;
  const void *dr, *dc;
  if (stage_array(rows_idx, ib, 0, &dr) || stage_array(cols_idx, ib, slot, &dc)) return -1;
  const unsigned blocks = (unsigned)((n + 255) / 256);
  OB2_FIELD_SWITCH(get_elements_kernel, blocks, 256, (const uint8_t *)bits, row_stride, (const uint32_t *)dr,
                   (const uint32_t *)dc, (uint32_t)n, out_dev)
  return after_launch("get_elements");
}
int oblas_b200_set_elements(int field, void *bits, size_t row_stride, const uint32_t *rows_idx,
                            const uint32_t *cols_idx, const uint8_t *vals, size_t n) {
  if (ensure_init()) return -1;
  if (n == 0) return 0;
  if (!bits || !rows_idx || !cols_idx || !vals) return fail("set_elements: null pointer");
  if (classify(bits) != PK_DEVVIS) return fail("set_elements: matrix must be device-visible");
  if (n > 0x7fffffffu) return fail("set_elements: too many elements");
  std::lock_guard<std::mutex> lk(cx().mu);
  const size_t ib = n * sizeof(uint32_t), slot = (ib + 255) & ~(size_t)255;
  if (scratch_reserve(2 * slot + n + 512)) return -1;
  const void *dr, *dc, *dv;
  if (stage_array(rows_idx, ib, 0, &dr) || stage_array(cols_idx, ib, slot, &dc) || stage_array(vals, n, 2 * slot, &dv))
    return -1;
  const unsigned blocks = (unsigned)((n + 255) / 256);
  OB2_FIELD_SWITCH(set_elements_kernel, blocks, 256, (uint8_t *)bits, row_stride, (const uint32_t *)dr,
                   (const uint32_t *)dc, (const uint8_t *)dv, (uint32_t)n)
  return after_launch("set_elements");
}
#undef OB2_FIELD_SWITCH

}  // extern "C"

// ============================================================================
// layer 1: the reference's API (include/oblas_compat.h)
// ============================================================================
#include "../../include/oblas_compat.h"

namespace {

unsigned round_up_u(unsigned k, unsigned a) { return (k + a - 1) / a * a; }

// One-row operation through the batched kernel, explicit tables, operands that
// are already device-visible and 16-byte friendly.
int single_vec(int op, int mode, uint8_t *a, const uint8_t *b, size_t k, const uint8_t *lo,
               const uint8_t *hi, uint32_t u) {
  uint32_t *&d_zero = cx().d_zero;  // a device word holding row index 0
  if (!d_zero) {
    CK(cudaMalloc(&d_zero, 4));
    CK(cudaMemset(d_zero, 0, 4));
  }
  RowOpArgs p;
  memset(&p, 0, sizeof p);
  p.a = a;
  p.b = b;
  p.chunks_per_row = (uint32_t)(k / 16);
  FastDiv fd = make_fastdiv(p.chunks_per_row);
  p.div_mul = fd.mul;
  p.div_shift = fd.shift;
  p.total_chunks = p.chunks_per_row;
  p.dst_rows = d_zero;
  p.src_rows = d_zero;
  p.u0 = u;
  p.u_max = 255u;
  if (lo) {
    p.use_explicit = 1;
    make_lut(p.lut, lo, hi);
    make_basis(p.basis, lo, hi);
  }
  launch_rowops(op, mode, p, cx().stream);
  return after_launch("row op");
}

// L1 entry: any pointers, any k.  u = 2 makes the kernels take the table path.
int l1_op(int op, uint8_t *a, const uint8_t *b, size_t k, const uint8_t *lo, const uint8_t *hi, uint32_t u) {
  if (ensure_init()) return -1;
  if (k == 0) return 0;
  std::lock_guard<std::mutex> lk(cx().mu);
  const bool has_b = (op == OP_AXPY || op == OP_ADD || op == OP_SWAP || op == OP_AXPY_B32);
  const size_t b_bytes = (op == OP_AXPY_B32) ? ((k + 31) / 32) * 4 : k;
  int mode = MUL_LUT;
  if (lo && nibble_table_linear(lo) && nibble_table_linear(hi)) mode = MUL_LIN;
  const bool a_dev = classify(a) == PK_DEVVIS, b_dev = !has_b || classify(b) == PK_DEVVIS;
  uint8_t *da = a;
  const uint8_t *db = b;
  size_t ka = (k + 15) & ~(size_t)15, kb = (b_bytes + 15) & ~(size_t)15;
  if (!a_dev || !b_dev) {  // stage plain host operands through device scratch
    if (scratch_reserve(ka + kb + 64)) return -1;
    if (!a_dev) {
      da = (uint8_t *)cx().scratch;
      CK(cudaMemsetAsync(da + (ka - 16), 0, 16, cx().stream));
      CK(cudaMemcpyAsync(da, a, k, cudaMemcpyHostToDevice, cx().stream));
    }
    if (has_b && !b_dev) {
      uint8_t *t = (uint8_t *)cx().scratch + ka;
      CK(cudaMemsetAsync(t + (kb - 16), 0, 16, cx().stream));
      CK(cudaMemcpyAsync(t, b, b_bytes, cudaMemcpyHostToDevice, cx().stream));
      db = t;
    }
  }
  const bool a_staged = !a_dev, b_staged = has_b && !b_dev;
  const bool vec_ok = ((reinterpret_cast<uintptr_t>(da) & 15) == 0) && (a_staged || (k & 15) == 0) &&
                      (!has_b || (((reinterpret_cast<uintptr_t>(db) & (op == OP_AXPY_B32 ? 3 : 15)) == 0) &&
                                  (b_staged || op == OP_AXPY_B32 || (k & 15) == 0)));
  if (vec_ok) {
    // staged buffers are padded to 16 bytes, so the vector kernel may run over ka
    size_t kk = a_staged ? ka : k;
    if (op == OP_AXPY_B32 && !b_staged && (k & 31)) {
      // mask words beyond ceil(k/32) must not be read: fall through to bytes
      Lut16x2 lt;
      memset(&lt, 0, sizeof lt);
      launch_rowop_bytes(op, da, db, k, lt, u, cx().stream);
      if (after_launch("row op (bytes)")) return -1;
    } else if (single_vec(op, mode, da, db, kk, lo, hi, u)) {
      return -1;
    }
  } else {
    Lut16x2 lt;
    memset(&lt, 0, sizeof lt);
    if (lo) make_lut(lt, lo, hi);
    launch_rowop_bytes(op, da, db, k, lt, u, cx().stream);
    if (after_launch("row op (bytes)")) return -1;
  }
  if (a_staged) CK(cudaMemcpyAsync(a, da, k, cudaMemcpyDeviceToHost, cx().stream));
  if (op == OP_SWAP && b_staged) CK(cudaMemcpyAsync(const_cast<uint8_t *>(b), db, k, cudaMemcpyDeviceToHost, cx().stream));
  if (a_staged || b_staged) CK(cudaStreamSynchronize(cx().stream));
  return 0;
}

// matrix-level single row op on managed/device matrices: one launch, one row
int mat_rowop(int field, int variant, int op, uint8_t *abits, size_t astride, const uint8_t *bbits,
              size_t bstride, size_t row_bytes, unsigned i, unsigned j, uint8_t u) {
  uint32_t di = i, dj = j;
  if (cx().async == 2 && variant == 0 && op != OP_AXPY_B32 && !(row_bytes & 15) && !(astride & 15) &&
      !(reinterpret_cast<uintptr_t>(abits) & 15)) {
    Ctx::OpQueue &q = cx().q;
    // cudaPointerGetAttributes costs more than everything else here: only for a matrix not seen last time
    const bool vis = (abits == q.ok_a || classify(abits) == PK_DEVVIS) && (!bbits || bbits == q.ok_b || bbits == q.ok_a || classify(bbits) == PK_DEVVIS);
    if (vis) {
      q.ok_a = abits;
      if (bbits) q.ok_b = bbits;
      if (ensure_init(false)) return -1;
      return queue_push(field, op, abits, astride, bbits, bstride, row_bytes, di, dj, u);
    }
  }
  return oblas_b200_rowops(field, variant, op, OBLAS_B200_MUL_AUTO, abits, astride, bbits, bstride,
                           row_bytes, &di, &dj, &u, 1);
}

int batch_rowop(int field, int variant, int op, uint8_t *abits, size_t astride, const uint8_t *bbits,
                size_t bstride, size_t row_bytes, const unsigned *i, const unsigned *j,
                const uint8_t *u, size_t n) {
  return oblas_b200_rowops(field, variant, op, OBLAS_B200_MUL_AUTO, abits, astride, bbits, bstride,
                           row_bytes, (const uint32_t *)i, (const uint32_t *)j, u, n);
}

const unsigned kFieldBits[4] = {1, 2, 4, 8};

// solve n systems given as separate allocations (struct API)
int solve_ptrs(int field, uint8_t **abits, uint8_t **bbits, size_t a_rs, unsigned rows, unsigned cols,
               size_t a_bytes, size_t b_rs, size_t b_bytes, unsigned *rank, unsigned *pivcols, size_t n) {
  if (ensure_init()) return -1;
  if (n == 0) return 0;
  std::lock_guard<std::mutex> lk(cx().mu);
  const size_t ptr_bytes = ((n * 8) + 255) & ~(size_t)255;
  const size_t rank_bytes = ((n * 4) + 255) & ~(size_t)255;
  const size_t piv_bytes = pivcols ? n * (size_t)rows * 4 : 0;
  if (scratch_reserve(2 * ptr_bytes + rank_bytes + piv_bytes + 256)) return -1;
  uint8_t *s = (uint8_t *)cx().scratch;
  unsigned long long *d_a = (unsigned long long *)s, *d_b = (unsigned long long *)(s + ptr_bytes);
  int32_t *d_rank = (int32_t *)(s + 2 * ptr_bytes);
  uint32_t *d_piv = pivcols ? (uint32_t *)(s + 2 * ptr_bytes + rank_bytes) : nullptr;
  CK(cudaMemcpyAsync(d_a, abits, n * 8, cudaMemcpyHostToDevice, cx().stream));
  if (bbits) CK(cudaMemcpyAsync(d_b, bbits, n * 8, cudaMemcpyHostToDevice, cx().stream));
  if (d_piv) CK(cudaMemsetAsync(d_piv, 0xff, piv_bytes, cx().stream));
  SolveArgs p;
  memset(&p, 0, sizeof p);
  p.A = abits[0]; p.a_rs = a_rs; p.rows = rows; p.cols = cols; p.a_bytes = (uint32_t)a_bytes;
  p.B = bbits ? bbits[0] : nullptr; p.b_rs = b_rs; p.b_bytes = bbits ? (uint32_t)b_bytes : 0;
  p.rank = d_rank; p.pivcols = d_piv;
  p.a_ptrs = d_a;
  p.b_ptrs = bbits ? d_b : nullptr;
  if (solve_common(field, p, n)) return -1;
  CK(cudaMemcpyAsync(rank, d_rank, n * 4, cudaMemcpyDeviceToHost, cx().stream));
  CK(cudaStreamSynchronize(cx().stream));
  if (tc_check_status()) return -1;
  if (pivcols) {
    // like the caller-side loop (oracle/ref_driver.c) only the first rank[k] entries of a system's
    // pivcols are written: a caller may size the array by min(rows, cols)
    for (size_t k = 0; k < n; k++)
      if (rank[k]) CK(cudaMemcpyAsync(pivcols + k * rows, d_piv + k * rows, (size_t)rank[k] * 4, cudaMemcpyDeviceToHost, cx().stream));
    CK(cudaStreamSynchronize(cx().stream));
  }
  return 0;
}

// the *_solve_batch wrappers take the geometry from system 0: every system must have the same
template <typename M>
int same_shape(M **A, M **B, size_t n, const char *what) {
  for (size_t k = 0; k < n; k++) {
    if (!A[k] || !A[k]->bits) return fail("%s: A[%zu] is null", what, k);
    if (A[k]->rows != A[0]->rows || A[k]->cols != A[0]->cols || A[k]->stride != A[0]->stride)
      return fail("%s: A[%zu] is %ux%u (stride %u), A[0] is %ux%u (stride %u): one batch = one shape", what, k,
                  A[k]->rows, A[k]->cols, A[k]->stride, A[0]->rows, A[0]->cols, A[0]->stride);
    if (!B) continue;
    if (!B[k] || !B[k]->bits) return fail("%s: B[%zu] is null", what, k);
    if (B[k]->stride != B[0]->stride || B[k]->rows < A[k]->rows)
      return fail("%s: B[%zu] has %u rows, stride %u; needs >= %u rows and stride %u", what, k, B[k]->rows,
                  B[k]->stride, A[k]->rows, B[0]->stride);
  }
  return 0;
}
int same_field(gfmat **A, gfmat **B, size_t n, const char *what) {
  for (size_t k = 0; k < n; k++)
    if (A[k]->field != A[0]->field || (B && B[k]->field != A[0]->field))
      return fail("%s: system %zu is over another field than system 0", what, k);
  return 0;
}

}  // namespace

extern "C" {

// ---- oblas.h ------------------------------------------------------------------
// oblas.c:4-11
uint32_t bfd_32(uint32_t word, uint8_t at, uint8_t len, uint8_t val) {
  uint32_t mask = (uint32_t)((1 << len) - 1) << at;
  return (word & ~mask) | ((uint32_t)val << at);
}
uint32_t bfx_32(uint32_t word, uint8_t at, uint8_t len) { return (word >> at) & (uint32_t)((1 << len) - 1); }

// oblas.c:13-19: nmemb blocks of `size` rounded up to `align`; here CUDA managed
void *oblas_alloc(size_t nmemb, size_t size, size_t align) {
  size_t each = (size + align - 1) / align * align;
  void *p = oblas_b200_managed_alloc(nmemb * each, align);
  if (!p) {
    fprintf(stderr, "liboblas_b200: oblas_alloc: %s\n", tl_err);
    exit(ENOMEM);
  }
  return p;
}

void oblas_xor(uint8_t *a, uint8_t *b, size_t k) {
  MUST(l1_op(OP_ADD, a, b, k, nullptr, nullptr, 1), "oblas_xor");
  MUST(finish_dropin(), "oblas_xor");
}
void oblas_axpy(uint8_t *a, const uint8_t *b, size_t k, const uint8_t *u_lo, const uint8_t *u_hi) {
  MUST(l1_op(OP_AXPY, a, b, k, u_lo, u_hi, 2), "oblas_axpy");
  MUST(finish_dropin(), "oblas_axpy");
}
void oblas_scal(uint8_t *a, size_t k, const uint8_t *u_lo, const uint8_t *u_hi) {
  MUST(l1_op(OP_SCAL, a, nullptr, k, u_lo, u_hi, 2), "oblas_scal");
  MUST(finish_dropin(), "oblas_scal");
}
void oblas_swap(uint8_t *a, uint8_t *b, size_t k) {
  MUST(l1_op(OP_SWAP, a, b, k, nullptr, nullptr, 1), "oblas_swap");
  MUST(finish_dropin(), "oblas_swap");
}
void oblas_axpy_gf2_gf256_32(uint8_t *a, uint32_t *b, size_t k, uint8_t u) {
  MUST(l1_op(OP_AXPY_B32, a, (const uint8_t *)b, k, nullptr, nullptr, u), "oblas_axpy_gf2_gf256_32");
  MUST(finish_dropin(), "oblas_axpy_gf2_gf256_32");
}

// ---- octmat.h -------------------------------------------------------------------
// octmat.c:7-15
octmat *octmat_new_aligned(unsigned rows, unsigned cols, unsigned align) {
  octmat *m = (octmat *)calloc(1, sizeof(octmat));
  if (!m) exit(ENOMEM);
  if (align < 16) align = 16;
  m->rows = rows;
  m->cols = cols;
  m->stride = round_up_u(cols, align);
  m->bits = (uint8_t *)oblas_alloc(rows, m->stride, align);
  memset(m->bits, 0, (size_t)rows * m->stride);
  return m;
}
octmat *octmat_new(unsigned rows, unsigned cols) { return octmat_new_aligned(rows, cols, 16); }
// octmat.c:17-22
void octmat_free(octmat *m) {
  MUST(queue_observe(), "octmat_free");
  if (m && m->bits) oblas_b200_free(m->bits);
  if (m) free(m);
}
void oswaprow(octmat *a, unsigned i, unsigned j) {
  if (i == j) return;  // octmat.c:25-26
  MUST(mat_rowop(OB2_GF256, 0, OP_SWAP, a->bits, a->stride, a->bits, a->stride, a->stride, i, j, 1), "oswaprow");
  MUST(finish_dropin(), "oswaprow");
}
void oaxpy(octmat *a, octmat *b, unsigned i, unsigned j, uint8_t u) {
  if (u == 0) return;  // octmat.c:36-37
  MUST(mat_rowop(OB2_GF256, 0, OP_AXPY, a->bits, a->stride, b->bits, b->stride, a->stride, i, j, u), "oaxpy");
  MUST(finish_dropin(), "oaxpy");
}
void oaddrow(octmat *a, octmat *b, unsigned i, unsigned j) {
  MUST(mat_rowop(OB2_GF256, 0, OP_ADD, a->bits, a->stride, b->bits, b->stride, a->stride, i, j, 1), "oaddrow");
  MUST(finish_dropin(), "oaddrow");
}
void oscal(octmat *a, unsigned i, uint8_t u) {
  if (u < 2) return;  // octmat.c:57-58
  MUST(mat_rowop(OB2_GF256, 0, OP_SCAL, a->bits, a->stride, nullptr, 0, a->stride, i, i, u), "oscal");
  MUST(finish_dropin(), "oscal");
}
void oaxpy_b32(octmat *a, uint32_t *b, unsigned i, uint8_t u) {
  // octmat.c:65-68: b is a caller-owned row of stride/32 mask words
  MUST(l1_op(OP_AXPY_B32, a->bits + (size_t)i * a->stride, (const uint8_t *)b, a->stride, nullptr, nullptr, u), "oaxpy_b32");
  MUST(finish_dropin(), "oaxpy_b32");
}

void oaxpy_batch(octmat *a, octmat *b, const unsigned *i, const unsigned *j, const uint8_t *u, size_t n) {
  MUST(batch_rowop(OB2_GF256, 0, OP_AXPY, a->bits, a->stride, b->bits, b->stride, a->stride, i, j, u, n), "oaxpy_batch");
  MUST(finish_dropin(), "oaxpy_batch");
}
void oaddrow_batch(octmat *a, octmat *b, const unsigned *i, const unsigned *j, size_t n) {
  MUST(batch_rowop(OB2_GF256, 0, OP_ADD, a->bits, a->stride, b->bits, b->stride, a->stride, i, j, nullptr, n), "oaddrow_batch");
  MUST(finish_dropin(), "oaddrow_batch");
}
void oscal_batch(octmat *a, const unsigned *i, const uint8_t *u, size_t n) {
  MUST(batch_rowop(OB2_GF256, 0, OP_SCAL, a->bits, a->stride, nullptr, 0, a->stride, i, nullptr, u, n), "oscal_batch");
  MUST(finish_dropin(), "oscal_batch");
}
void oswaprow_batch(octmat *a, const unsigned *i, const unsigned *j, size_t n) {
  MUST(batch_rowop(OB2_GF256, 0, OP_SWAP, a->bits, a->stride, a->bits, a->stride, a->stride, i, j, nullptr, n), "oswaprow_batch");
  MUST(finish_dropin(), "oswaprow_batch");
}

// ---- gfmat.h ---------------------------------------------------------------------
// gfmat.c:48-60: stride in WORDS, ceil(cols/vpw) rounded up to `align` words,
// align = sizeof(void*) for GF2_1 else OCTMAT_ALIGN
gfmat *gfmat_new_aligned(gfmat_type field, unsigned rows, unsigned cols, unsigned align) {
  gfmat *m = (gfmat *)calloc(1, sizeof(gfmat));
  if (!m) exit(ENOMEM);
  unsigned vpw = 32 / kFieldBits[field];
  m->field = field;
  m->rows = rows;
  m->cols = cols;
  m->align = (field == GF2_1) ? (unsigned)sizeof(void *) : (align < 16 ? 16 : align);
  m->stride = round_up_u((cols + vpw - 1) / vpw, m->align);
  m->bits = (oblas_word *)oblas_alloc(rows, (size_t)m->stride * sizeof(oblas_word), m->align);
  memset(m->bits, 0, (size_t)rows * m->stride * sizeof(oblas_word));
  return m;
}
gfmat *gfmat_new(gfmat_type field, unsigned rows, unsigned cols) { return gfmat_new_aligned(field, rows, cols, 16); }
// gfmat.c:62-74.  The reference copies stride*cols words (line 72), which over- or
// under-runs for non-square matrices; we copy the whole matrix (stride*rows).
gfmat *gfmat_copy(gfmat *src) {
  if (!src || !src->bits) return NULL;
  MUST(oblas_b200_sync(), "gfmat_copy");
  gfmat *m = (gfmat *)calloc(1, sizeof(gfmat));
  if (!m) exit(ENOMEM);
  *m = *src;
  m->bits = (oblas_word *)oblas_alloc(m->rows, (size_t)m->stride * sizeof(oblas_word), m->align);
  memcpy(m->bits, src->bits, sizeof(oblas_word) * (size_t)m->stride * m->rows);
  return m;
}
void gfmat_free(gfmat *m) {
  MUST(queue_observe(), "gfmat_free");
  if (m && m->bits) oblas_b200_free(m->bits);
  if (m) free(m);
}
// host-side element access on managed memory (gfmat.c:83-98)
uint8_t gfmat_get(gfmat *m, unsigned i, unsigned j) {
  MUST(queue_observe(), "gfmat_get");
  if (i >= m->rows || j >= m->cols) return 0;
  unsigned e = kFieldBits[m->field], vpw = 32 / e;
  oblas_word *a = m->bits + (size_t)i * m->stride;
  return (uint8_t)bfx_32(a[j / vpw], (uint8_t)((j % vpw) * e), (uint8_t)e);
}
void gfmat_set(gfmat *m, unsigned i, unsigned j, uint8_t b) {
  MUST(queue_observe(), "gfmat_set");
  if (i >= m->rows || j >= m->cols) return;
  unsigned e = kFieldBits[m->field], vpw = 32 / e;
  oblas_word *a = m->bits + (size_t)i * m->stride;
  a[j / vpw] = bfd_32(a[j / vpw], (uint8_t)((j % vpw) * e), (uint8_t)e, (uint8_t)(b % (1u << e)));
}
// gfmat.c:100-116
void gfmat_print(gfmat *m, FILE *stream) {
  MUST(queue_observe(), "gfmat_print");
  static const char *fn[] = {"gf2", "gf4", "gf16", "gf256"};
  fprintf(stream, "%s [%ux%u]\n", fn[m->field], m->rows, m->cols);
  fprintf(stream, "|     ");
  for (unsigned j = 0; j < m->cols; j++) fprintf(stream, "| %03d ", j);
  fprintf(stream, "|\n");
  for (unsigned i = 0; i < m->rows; i++) {
    fprintf(stream, "| %03d | %3d ", i, gfmat_get(m, i, 0));
    for (unsigned j = 1; j < m->cols; j++) fprintf(stream, "| %3d ", gfmat_get(m, i, j));
    fprintf(stream, "|\n");
  }
}
void gfmat_add(gfmat *a, gfmat *b, unsigned i, unsigned j) {
  MUST(mat_rowop(a->field, 0, OP_ADD, (uint8_t *)a->bits, (size_t)a->stride * 4, (uint8_t *)b->bits, (size_t)b->stride * 4, (size_t)a->stride * 4, i, j, 1), "gfmat_add");
  MUST(finish_dropin(), "gfmat_add");
}
void gfmat_swaprow(gfmat *m, unsigned i, unsigned j) {
  if (i == j) return;
  MUST(mat_rowop(m->field, 0, OP_SWAP, (uint8_t *)m->bits, (size_t)m->stride * 4, (uint8_t *)m->bits, (size_t)m->stride * 4, (size_t)m->stride * 4, i, j, 1), "gfmat_swaprow");
  MUST(finish_dropin(), "gfmat_swaprow");
}
void gfmat_axpy(gfmat *a, gfmat *b, unsigned i, unsigned j, uint8_t u) {
  if (u == 0) return;
  // GF2_1 has no tables (gfmat.c:14-19): only u==1 is defined; we treat u>1 as XOR-free no-op
  if (a->field == GF2_1 && u > 1) return;
  MUST(mat_rowop(a->field, 0, OP_AXPY, (uint8_t *)a->bits, (size_t)a->stride * 4, (uint8_t *)b->bits, (size_t)b->stride * 4, (size_t)a->stride * 4, i, j, u), "gfmat_axpy");
  MUST(finish_dropin(), "gfmat_axpy");
}
void gfmat_scal(gfmat *a, unsigned i, uint8_t u) {
  if (u < 2 || a->field == GF2_1) return;
  MUST(mat_rowop(a->field, 0, OP_SCAL, (uint8_t *)a->bits, (size_t)a->stride * 4, nullptr, 0, (size_t)a->stride * 4, i, i, u), "gfmat_scal");
  MUST(finish_dropin(), "gfmat_scal");
}
void gfmat_zero(gfmat *a, unsigned i) {
  MUST(mat_rowop(a->field, 0, OP_ZERO, (uint8_t *)a->bits, (size_t)a->stride * 4, nullptr, 0, (size_t)a->stride * 4, i, i, 0), "gfmat_zero");
  MUST(finish_dropin(), "gfmat_zero");
}
// gfmat.c:163-175, truncation into uint8_t kept
void gfmat_fill(gfmat *m, unsigned i, uint8_t *dst) {
  MUST(queue_observe(), "gfmat_fill");
  unsigned e = kFieldBits[m->field], vpw = 32 / e;
  oblas_word *a = m->bits + (size_t)i * m->stride;
  for (unsigned idx = 0; idx < m->stride; idx++)
    for (unsigned tz = 0; tz < 32; tz++)
      if ((a[idx] >> tz) & 1u) dst[tz / e + idx * vpw] |= (uint8_t)(1u << (tz % vpw));
}
static unsigned elem_flags_host(uint32_t w, unsigned e) {
  unsigned z = 0;
  for (unsigned bit = 0; bit < 32; bit++)
    if ((w >> bit) & 1u) z |= 1u << (bit / e);
  return z;
}
static unsigned nnz_compat(const oblas_word *a, unsigned e, unsigned s, unsigned en) {
  // gfmat.c:183-212 / binmat.c:87-116 bug-for-bug: after a partial first word the
  // element index s is bumped and reused as the word index of the middle loop
  unsigned vpw = 32 / e, n = 0;
  unsigned sq = s / vpw, eq = en / vpw, sr = s % vpw, er = en % vpw;
  uint32_t m_first = ~((1u << sr) - 1u), m_last = (1u << er) - 1u;
  if (sr) {
    n += (unsigned)__builtin_popcount(elem_flags_host(a[sq], e) & m_first);
    s++;
  }
  for (unsigned w = s; w < eq; w++) n += (unsigned)__builtin_popcount(elem_flags_host(a[w], e));
  if (en > eq && m_last) n += (unsigned)__builtin_popcount(elem_flags_host(a[eq], e) & m_last);
  return n;
}
unsigned gfmat_nnz(gfmat *m, unsigned i, unsigned s, unsigned e) {
  MUST(queue_observe(), "gfmat_nnz");
  if (i >= m->rows || s > e || e > (m->cols + 1)) return 0;
  return nnz_compat(m->bits + (size_t)i * m->stride, kFieldBits[m->field], s, e);
}
// gfmat.c:215-231
void gfmat_swapcol(gfmat *m, unsigned i, unsigned j) {
  MUST(queue_observe(), "gfmat_swapcol");
  if (i == j) return;
  unsigned e = kFieldBits[m->field], vpw = 32 / e;
  for (unsigned r = 0; r < m->rows; r++) {
    oblas_word *row = m->bits + (size_t)r * m->stride;
    uint8_t vi = (uint8_t)bfx_32(row[i / vpw], (uint8_t)((i % vpw) * e), (uint8_t)e);
    uint8_t vj = (uint8_t)bfx_32(row[j / vpw], (uint8_t)((j % vpw) * e), (uint8_t)e);
    row[i / vpw] = bfd_32(row[i / vpw], (uint8_t)((i % vpw) * e), (uint8_t)e, vj);
    row[j / vpw] = bfd_32(row[j / vpw], (uint8_t)((j % vpw) * e), (uint8_t)e, vi);
  }
}
void gfmat_axpy_batch(gfmat *a, gfmat *b, const unsigned *i, const unsigned *j, const uint8_t *u, size_t n) {
  MUST(batch_rowop(a->field, 0, OP_AXPY, (uint8_t *)a->bits, (size_t)a->stride * 4, (uint8_t *)b->bits, (size_t)b->stride * 4, (size_t)a->stride * 4, i, j, u, n), "gfmat_axpy_batch");
  MUST(finish_dropin(), "gfmat_axpy_batch");
}
void gfmat_add_batch(gfmat *a, gfmat *b, const unsigned *i, const unsigned *j, size_t n) {
  MUST(batch_rowop(a->field, 0, OP_ADD, (uint8_t *)a->bits, (size_t)a->stride * 4, (uint8_t *)b->bits, (size_t)b->stride * 4, (size_t)a->stride * 4, i, j, nullptr, n), "gfmat_add_batch");
  MUST(finish_dropin(), "gfmat_add_batch");
}
void gfmat_scal_batch(gfmat *a, const unsigned *i, const uint8_t *u, size_t n) {
  if (a->field == GF2_1) return;
  MUST(batch_rowop(a->field, 0, OP_SCAL, (uint8_t *)a->bits, (size_t)a->stride * 4, nullptr, 0, (size_t)a->stride * 4, i, nullptr, u, n), "gfmat_scal_batch");
  MUST(finish_dropin(), "gfmat_scal_batch");
}
void gfmat_swaprow_batch(gfmat *m, const unsigned *i, const unsigned *j, size_t n) {
  MUST(batch_rowop(m->field, 0, OP_SWAP, (uint8_t *)m->bits, (size_t)m->stride * 4, (uint8_t *)m->bits, (size_t)m->stride * 4, (size_t)m->stride * 4, i, j, nullptr, n), "gfmat_swaprow_batch");
  MUST(finish_dropin(), "gfmat_swaprow_batch");
}

// ---- binmat.h ---------------------------------------------------------------------
// binmat.c:8-17
binmat *binmat_new(unsigned rows, unsigned cols) {
  binmat *m = (binmat *)calloc(1, sizeof(binmat));
  if (!m) exit(ENOMEM);
  m->rows = rows;
  m->cols = cols;
  m->stride = round_up_u((cols + 31) / 32, (unsigned)sizeof(void *));
  m->bits = (oblas_word *)oblas_alloc(rows, (size_t)m->stride * sizeof(oblas_word), sizeof(void *));
  memset(m->bits, 0, (size_t)rows * m->stride * sizeof(oblas_word));
  return m;
}
void binmat_free(binmat *m) {
  MUST(queue_observe(), "binmat_free");
  if (m && m->bits) oblas_b200_free(m->bits);
  if (m) free(m);
}
uint8_t binmat_get(binmat *m, unsigned i, unsigned j) {
  MUST(queue_observe(), "binmat_get");
  if (i >= m->rows || j >= m->cols) return 0;
  return (uint8_t)((m->bits[(size_t)i * m->stride + j / 32] >> (j % 32)) & 1u);
}
// binmat.c:35-42: the bit is always set, `b` is ignored (line 41) -- kept
void binmat_set(binmat *m, unsigned i, unsigned j, uint8_t b) {
  MUST(queue_observe(), "binmat_set");
  (void)b;
  if (i >= m->rows || j >= m->cols) return;
  m->bits[(size_t)i * m->stride + j / 32] |= 1u << (j % 32);
}
void binmat_add(binmat *a, binmat *b, unsigned i, unsigned j) {
  MUST(mat_rowop(OB2_GF2, 0, OP_ADD, (uint8_t *)a->bits, (size_t)a->stride * 4, (uint8_t *)b->bits, (size_t)b->stride * 4, (size_t)a->stride * 4, i, j, 1), "binmat_add");
  MUST(finish_dropin(), "binmat_add");
}
void binmat_swaprow(binmat *m, unsigned i, unsigned j) {
  if (i == j) return;
  MUST(mat_rowop(OB2_GF2, 0, OP_SWAP, (uint8_t *)m->bits, (size_t)m->stride * 4, (uint8_t *)m->bits, (size_t)m->stride * 4, (size_t)m->stride * 4, i, j, 1), "binmat_swaprow");
  MUST(finish_dropin(), "binmat_swaprow");
}
void binmat_zero(binmat *a, unsigned i) {
  MUST(mat_rowop(OB2_GF2, 0, OP_ZERO, (uint8_t *)a->bits, (size_t)a->stride * 4, nullptr, 0, (size_t)a->stride * 4, i, i, 0), "binmat_zero");
  MUST(finish_dropin(), "binmat_zero");
}
// binmat.c:64-75
void binmat_fill(binmat *m, unsigned i, uint8_t *dst) {
  MUST(queue_observe(), "binmat_fill");
  oblas_word *a = m->bits + (size_t)i * m->stride;
  for (unsigned idx = 0; idx < m->stride; idx++)
    for (unsigned tz = 0; tz < 32; tz++)
      if ((a[idx] >> tz) & 1u) dst[tz + idx * 32] |= (uint8_t)(1u << tz);
}
// binmat.c:77-80
void binmat_expand(binmat *m, uint32_t *src, unsigned i, uint8_t u) {
  MUST(l1_op(OP_AXPY_B32, (uint8_t *)(m->bits + (size_t)i * m->stride), (const uint8_t *)src, (size_t)m->stride * 4, nullptr, nullptr, u), "binmat_expand");
  MUST(finish_dropin(), "binmat_expand");
}
unsigned binmat_nnz(binmat *m, unsigned i, unsigned s, unsigned e) {
  MUST(queue_observe(), "binmat_nnz");
  if (i >= m->rows || s > e || e > (m->cols + 1)) return 0;
  return nnz_compat(m->bits + (size_t)i * m->stride, 1, s, e);
}
void binmat_add_batch(binmat *a, binmat *b, const unsigned *i, const unsigned *j, size_t n) {
  MUST(batch_rowop(OB2_GF2, 0, OP_ADD, (uint8_t *)a->bits, (size_t)a->stride * 4, (uint8_t *)b->bits, (size_t)b->stride * 4, (size_t)a->stride * 4, i, j, nullptr, n), "binmat_add_batch");
  MUST(finish_dropin(), "binmat_add_batch");
}
void binmat_swaprow_batch(binmat *m, const unsigned *i, const unsigned *j, size_t n) {
  MUST(batch_rowop(OB2_GF2, 0, OP_SWAP, (uint8_t *)m->bits, (size_t)m->stride * 4, (uint8_t *)m->bits, (size_t)m->stride * 4, (size_t)m->stride * 4, i, j, nullptr, n), "binmat_swaprow_batch");
  MUST(finish_dropin(), "binmat_swaprow_batch");
}

// ---- caller-side loops as single calls ------------------------------------------------
void octmat_solve_batch(octmat **A, octmat **B, unsigned *rank, size_t n) {
  if (n == 0) return;
  MUST(same_shape(A, B, n, "octmat_solve_batch"), "octmat_solve_batch");
  std::vector<uint8_t *> ab(n), bb(n);
  for (size_t k = 0; k < n; k++) {
    ab[k] = A[k]->bits;
    if (B) bb[k] = B[k]->bits;
  }
  MUST(solve_ptrs(OB2_GF256, ab.data(), B ? bb.data() : nullptr, A[0]->stride, A[0]->rows, A[0]->cols,
                  A[0]->stride, B ? B[0]->stride : 0, B ? B[0]->stride : 0, rank, nullptr, n), "octmat_solve_batch");
}
unsigned octmat_solve(octmat *A, octmat *B, unsigned *pivcols) {
  unsigned rank = 0;
  uint8_t *ab = A->bits, *bb = B ? B->bits : nullptr;
  MUST(solve_ptrs(OB2_GF256, &ab, B ? &bb : nullptr, A->stride, A->rows, A->cols, A->stride,
                  B ? B->stride : 0, B ? B->stride : 0, &rank, pivcols, 1), "octmat_solve");
  return rank;
}
void gfmat_solve_batch(gfmat **A, gfmat **B, unsigned *rank, size_t n) {
  if (n == 0) return;
  MUST(same_shape(A, B, n, "gfmat_solve_batch"), "gfmat_solve_batch");
  MUST(same_field(A, B, n, "gfmat_solve_batch"), "gfmat_solve_batch");
  std::vector<uint8_t *> ab(n), bb(n);
  for (size_t k = 0; k < n; k++) {
    ab[k] = (uint8_t *)A[k]->bits;
    if (B) bb[k] = (uint8_t *)B[k]->bits;
  }
  MUST(solve_ptrs(A[0]->field, ab.data(), B ? bb.data() : nullptr, (size_t)A[0]->stride * 4, A[0]->rows, A[0]->cols,
                  (size_t)A[0]->stride * 4, B ? (size_t)B[0]->stride * 4 : 0, B ? (size_t)B[0]->stride * 4 : 0, rank, nullptr, n), "gfmat_solve_batch");
}
unsigned gfmat_solve(gfmat *A, gfmat *B, unsigned *pivcols) {
  unsigned rank = 0;
  uint8_t *ab = (uint8_t *)A->bits, *bb = B ? (uint8_t *)B->bits : nullptr;
  MUST(solve_ptrs(A->field, &ab, B ? &bb : nullptr, (size_t)A->stride * 4, A->rows, A->cols, (size_t)A->stride * 4,
                  B ? (size_t)B->stride * 4 : 0, B ? (size_t)B->stride * 4 : 0, &rank, pivcols, 1), "gfmat_solve");
  return rank;
}
void binmat_rref_batch(binmat **A, binmat **B, unsigned *rank, size_t n) {
  if (n == 0) return;
  MUST(same_shape(A, B, n, "binmat_rref_batch"), "binmat_rref_batch");
  std::vector<uint8_t *> ab(n), bb(n);
  for (size_t k = 0; k < n; k++) {
    ab[k] = (uint8_t *)A[k]->bits;
    if (B) bb[k] = (uint8_t *)B[k]->bits;
  }
  MUST(solve_ptrs(OB2_GF2, ab.data(), B ? bb.data() : nullptr, (size_t)A[0]->stride * 4, A[0]->rows, A[0]->cols,
                  (size_t)A[0]->stride * 4, B ? (size_t)B[0]->stride * 4 : 0, B ? (size_t)B[0]->stride * 4 : 0, rank, nullptr, n), "binmat_rref_batch");
}
unsigned binmat_rref(binmat *A, binmat *B, unsigned *pivcols) {
  unsigned rank = 0;
  uint8_t *ab = (uint8_t *)A->bits, *bb = B ? (uint8_t *)B->bits : nullptr;
  MUST(solve_ptrs(OB2_GF2, &ab, B ? &bb : nullptr, (size_t)A->stride * 4, A->rows, A->cols, (size_t)A->stride * 4,
                  B ? (size_t)B->stride * 4 : 0, B ? (size_t)B->stride * 4 : 0, &rank, pivcols, 1), "binmat_rref");
  return rank;
}
void octmat_apply(octmat *C, octmat *D, octmat *Y) {
  unsigned nbytes = D->stride < Y->stride ? D->stride : Y->stride;
  MUST(oblas_b200_gf256_apply(C->bits, C->stride, C->rows, C->cols, D->bits, D->stride, Y->bits, Y->stride, nbytes, 0), "octmat_apply");
  MUST(oblas_b200_sync(), "octmat_apply");
}

}  // extern "C"
