// Host runtime behind include/opty_b200.h: owns the device-resident trajectory
// matrix, the residual / Jacobian buffers, pinned host buffers, the streams and
// the TMA descriptors, loads the generated sm_100a modules and launches them.
//
// It replaces the NumPy / Cython glue of the reference's callback path
// (opty/utils.py:277-326 parse_free, opty/direct_collocation.py:2891-2926
// _merge_fixed_free, :2382-2446 constraints, :2816-2887 constraints_jacobian),
// the Python index loop (opty/direct_collocation.py:2628-2684) and the NumPy
// quadrature of create_objective_function (opty/utils.py:329-470).
//
// Driver-API entry points (module loading, tensor-map encoding, launches) are
// resolved through cudaGetDriverEntryPoint so that this library has no
// load-time dependency on libcuda.so.1: it can be dlopen'ed on a machine
// without a GPU (symbol checks), and fails loudly at opty_colloc_create there.
//
// NVTX ranges (opty_b200 domain) mark upload / pre-pass + eval launches /
// device->host copies / quadrature for timeline tools; they cost nothing when
// no tool is attached.

#include <cuda.h>
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/opty_b200.h"
#include "colloc_params.h"

namespace {

thread_local std::string g_err;

int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}

#define RT_CHECK(expr)                                                                     \
  do {                                                                                     \
    cudaError_t e__ = (expr);                                                              \
    if (e__ != cudaSuccess) {                                                              \
      return fail(OPTY_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorName(e__) + ": " + \
                                     cudaGetErrorString(e__));                             \
    }                                                                                      \
  } while (0)

struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
};

// ---- driver API, resolved lazily ------------------------------------------
struct DriverApi {
  bool ready = false;
  CUresult (*ModuleLoadData)(CUmodule*, const void*) = nullptr;
  CUresult (*ModuleUnload)(CUmodule) = nullptr;
  CUresult (*ModuleGetFunction)(CUfunction*, CUmodule, const char*) = nullptr;
  CUresult (*ModuleGetGlobal)(CUdeviceptr*, size_t*, CUmodule, const char*) = nullptr;
  CUresult (*FuncSetAttribute)(CUfunction, CUfunction_attribute, int) = nullptr;
  CUresult (*LaunchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned,
                           unsigned, CUstream, void**, void**) = nullptr;
  CUresult (*LaunchKernelEx)(const CUlaunchConfig*, CUfunction, void**, void**) = nullptr;
  CUresult (*GetErrorString)(CUresult, const char**) = nullptr;
  CUresult (*OccupancyMaxActiveBlocksPerMultiprocessor)(int*, CUfunction, int, size_t) = nullptr;
  CUresult (*TensorMapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                   const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                   const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill) = nullptr;
};

DriverApi g_drv;

template <typename F>
int load_entry(const char* name, F* fn) {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaError_t e = cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q);
  if (e != cudaSuccess || p == nullptr || q != cudaDriverEntryPointSuccess) {
    return fail(OPTY_ERR_CUDA, std::string("cannot resolve CUDA driver entry point ") + name +
                                   (e != cudaSuccess ? std::string(": ") + cudaGetErrorString(e) : ""));
  }
  *fn = reinterpret_cast<F>(p);
  return OPTY_OK;
}

int init_driver() {
  if (g_drv.ready) return OPTY_OK;
  int rc;
  if ((rc = load_entry("cuModuleLoadData", &g_drv.ModuleLoadData))) return rc;
  if ((rc = load_entry("cuModuleUnload", &g_drv.ModuleUnload))) return rc;
  if ((rc = load_entry("cuModuleGetFunction", &g_drv.ModuleGetFunction))) return rc;
  if ((rc = load_entry("cuModuleGetGlobal", &g_drv.ModuleGetGlobal))) return rc;
  if ((rc = load_entry("cuFuncSetAttribute", &g_drv.FuncSetAttribute))) return rc;
  if ((rc = load_entry("cuLaunchKernel", &g_drv.LaunchKernel))) return rc;
  if ((rc = load_entry("cuLaunchKernelEx", &g_drv.LaunchKernelEx))) return rc;
  if ((rc = load_entry("cuGetErrorString", &g_drv.GetErrorString))) return rc;
  if ((rc = load_entry("cuOccupancyMaxActiveBlocksPerMultiprocessor", &g_drv.OccupancyMaxActiveBlocksPerMultiprocessor)))
    return rc;
  if ((rc = load_entry("cuTensorMapEncodeTiled", &g_drv.TensorMapEncodeTiled))) return rc;
  g_drv.ready = true;
  return OPTY_OK;
}

std::string drv_err(CUresult r) {
  const char* s = nullptr;
  if (g_drv.GetErrorString) g_drv.GetErrorString(r, &s);
  return s ? std::string(s) : std::string("CUresult ") + std::to_string((int)r);
}

#define DRV_CHECK(expr)                                                     \
  do {                                                                      \
    CUresult r__ = (expr);                                                  \
    if (r__ != CUDA_SUCCESS) {                                              \
      return fail(OPTY_ERR_CUDA, std::string(#expr) + ": " + drv_err(r__)); \
    }                                                                       \
  } while (0)

inline int64_t round_up(int64_t v, int64_t m) { return (v + m - 1) / m * m; }

// ---- Jacobian structure kernel ---------------------------------------------
// One thread per COO entry; entry e of node i, equation j, partial c.  Column
// formulas are those of opty/direct_collocation.py:2655-2675, row formula of
// :2644, entry order of :2677-2684.
__global__ void opty_jac_indices_kernel(long long first, long long count, long long N, long long n,
                                        long long q, long long M, long long P, int method,
                                        long long* __restrict__ rows, long long* __restrict__ cols) {
  const long long MP = M * P;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < count;
       t += (long long)gridDim.x * blockDim.x) {
    const long long e = first + t;
    const long long i = e / MP;
    const long long rem = e - i * MP;
    const long long j = rem / P;
    const long long c = rem - j * P;
    long long col;
    if (method == OPTY_BACKWARD_EULER) {
      if (c < n) col = c * N + i + 1;
      else if (c < 2 * n) col = (c - n) * N + i;
      else if (c < 2 * n + q) col = n * N + (c - 2 * n) * N + i + 1;
      else col = (n + q) * N + (c - 2 * n - q);
    } else {
      if (c < n) col = c * N + i;
      else if (c < 2 * n) col = (c - n) * N + i + 1;
      else if (c < 2 * n + q) col = n * N + (c - 2 * n) * N + i;
      else if (c < 2 * n + 2 * q) col = n * N + (c - 2 * n - q) * N + i + 1;
      else col = (n + q) * N + (c - 2 * n - 2 * q);
    }
    rows[t] = j * (N - 1) + i;
    cols[t] = col;
  }
}

// ---- quadrature kernels (opty/utils.py:418-434) -----------------------------
// `vals` is the node-major output of an integrand module: per point
// [f, df/darg_0 .. df/darg_{na-1}, df/dconst_0 .. df/dconst_{nc-1}].
// Stage 1: grid-stride over the points; every thread writes the weighted
// partials with respect to the array arguments straight to grad[a*N + i]
// (coalesced over i) and accumulates f and the partials with respect to the
// scalar arguments; fixed-shape block reduction, one partial sum per block.
// Stage 2: one block adds the per-block partial sums in a fixed order, so the
// result does not depend on scheduling (no floating-point atomics).
#define OPTY_QUAD_THREADS 256
#define OPTY_QUAD_MAX_SCALARS 32

__device__ __forceinline__ double quad_weight_sum(int rule, int i, int N) {
  // weight of the summed quantities (objective value, parameter partials)
  if (rule == OPTY_MIDPOINT) return i < N - 1 ? 1.0 : 0.0;
  return i > 0 ? 1.0 : 0.0;
}

__device__ __forceinline__ double quad_weight_time(int rule, int i, int N) {
  // weight of the partials with respect to trajectory values
  if (rule == OPTY_MIDPOINT) return (i == 0 || i == N - 1) ? 0.5 : 1.0;
  return i > 0 ? 1.0 : 0.0;
}

__global__ void __launch_bounds__(OPTY_QUAD_THREADS)
opty_quadrature_stage1(const double* __restrict__ vals, int N, int P, int na, int nc, int rule, double scale,
                       double* __restrict__ grad, double* __restrict__ partial) {
  __shared__ double red[OPTY_QUAD_THREADS / 32][1 + OPTY_QUAD_MAX_SCALARS];
  double acc[1 + OPTY_QUAD_MAX_SCALARS];
  for (int s = 0; s <= nc; ++s) acc[s] = 0.0;
  for (int i = blockIdx.x * OPTY_QUAD_THREADS + threadIdx.x; i < N; i += gridDim.x * OPTY_QUAD_THREADS) {
    const double* row = vals + (long long)i * P;
    const double wt = quad_weight_time(rule, i, N);
    for (int a = 0; a < na; ++a) grad[(long long)a * N + i] = wt != 0.0 ? scale * wt * row[1 + a] : 0.0;
    const double ws = quad_weight_sum(rule, i, N);
    if (ws != 0.0) {  // points of weight zero are skipped, not multiplied: they may hold Inf / NaN
      acc[0] += ws * row[0];
      for (int s = 0; s < nc; ++s) acc[1 + s] += ws * row[1 + na + s];
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int s = 0; s <= nc; ++s) {
    double v = acc[s];
    for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
    if (lane == 0) red[warp][s] = v;
  }
  __syncthreads();
  if (threadIdx.x <= nc) {
    double v = 0.0;
    for (int w = 0; w < OPTY_QUAD_THREADS / 32; ++w) v += red[w][threadIdx.x];
    partial[(long long)blockIdx.x * (1 + OPTY_QUAD_MAX_SCALARS) + threadIdx.x] = v;
  }
}

__global__ void opty_quadrature_stage2(const double* __restrict__ partial, int blocks, int nc, double scale,
                                       double* __restrict__ out) {
  // out[0] = value, out[1 + s] = partial with respect to scalar argument s
  const int s = threadIdx.x;
  if (s > nc) return;
  double v = 0.0;
  for (int b = 0; b < blocks; ++b) v += partial[(long long)b * (1 + OPTY_QUAD_MAX_SCALARS) + s];
  out[s] = scale * v;
}

// layout of the generated modules' `opty_module_info` table (codegen.py)
enum {
  INFO_MAGIC = 0,
  INFO_VERSION = 1,
  INFO_WARPS = 2,
  INFO_GROUPS = 3,
  INFO_DERIVED = 4,
  INFO_PRE_GROUPS = 5,
  INFO_TMA_LOAD = 6,
  INFO_TMA_STORE = 7,
  INFO_TILE_BUFS = 8,
  INFO_TILE_DOUBLES = 9,
  INFO_NMAPS = 10,
  INFO_MAP_WIDTH0 = 11,  // .. 18
  INFO_NUM_INV = 19,
  INFO_R = 20,
  INFO_M = 21,
  INFO_P = 22,
  INFO_HAS_AUX = 23,
  INFO_PERSISTENT = 24,  // 0 grid kernel, 1 code-stationary work queue, 2 row-stationary static schedule
  INFO_MIN_BLOCKS = 25,
  INFO_SMEM_BYTES = 26,  // row-stationary kernel: dynamic shared memory per block
  INFO_SLOTS = 27,       //   blocks to launch (one per slot of the emitted schedule)
  INFO_TILES = 28,       //   node tiles the schedule was laid out for
  INFO_CVAL0 = 29,       //   first entry of the invariants table that belongs to the constant column runs
  INFO_FUSED_PRE = 30,   //   1: the main kernel does the pre-pass's work itself (no pre-pass launch)
  INFO_XROWS = 31,       //   rows of the input window a block fetches per item (box height of the input map)
  INFO_NCONST = 32,      // doubles in the constant column runs
  INFO_CONST_KERNEL = 33,  // 1: the module has opty_colloc_const, which writes them (else the main kernels do)
  INFO_WORDS = 40
};
const int kInfoMagic = 0x4f505459;
const int kEmitterVersion = 8;
const int kMaxMaps = 8;

}  // namespace

struct opty_colloc {
  opty_colloc_cfg cfg;
  int nn = 0;          // evaluation nodes of this handle
  int ncols = 0;       // trajectory columns held (nn + 1; elementwise: nn)
  int R = 0;           // trajectory rows n + q + k
  int RD = 0;          // ... plus derived rows
  int K = 0;           // M * P
  int64_t ldt = 0;
  size_t free_len = 0;
  bool elementwise = false;

  // kernel geometry, read from the primary module
  int warps = 0, num_derived = 0, pre_groups = 0, tma_load = 0, tma_store = 0, tile_bufs = 0,
      tile_doubles = 0, num_inv = 0, persistent = 0, stat_smem = 0, stat_slots = 0, stat_tiles = 0, stat_cval0 = 0, fused_pre = 0, stat_xrows = 0, num_const = 0, const_kernel = 0;

  struct Module {
    CUmodule mod = nullptr;
    CUfunction f_eval = nullptr;
    CUdeviceptr ci_sym = 0;
    int num_groups = 0;
    int nmaps = 0;
    int widths[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    std::vector<std::vector<unsigned char>> tmaps;  // per ring slot: OptyTmaps blob (in + out[nmaps])
    int* d_work = nullptr;      // persistent kernel: tile counter per group + departure counter
    unsigned long long* d_ready = nullptr;  // row-stationary kernel: progress counters of the fused pre-pass
    unsigned persist_grid = 0;  // resident blocks on the whole device
  };
  std::vector<Module> modules;  // [0] = primary (carries opty_colloc_inv / opty_colloc_pre)
  CUfunction f_inv = nullptr;
  CUfunction f_pre = nullptr;
  CUfunction f_const = nullptr;  // grid kernel: writes the constant column runs (opty_colloc_const)
  int num_sms = 0;

  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr;   // speculative Jacobian D2H
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  cudaEvent_t ev_copy = nullptr, ev_con = nullptr;
  bool jac_inflight = false;
  std::vector<char> slot_copying;       // ring slots a speculative copy may still be reading
  uint64_t eval_seq = 0;                // evaluations launched so far
  uint64_t copy_seq = 0;                // evaluation the most recent speculative copy belongs to

  double* d_traj = nullptr;
  double* d_tiled = nullptr;  // direct-input modules: trajectory + derived rows tile by tile (colloc_params.h)
  double* d_uni = nullptr;
  double* d_inv = nullptr;
  std::vector<double*> d_con, d_jac;
  int ring = -1;

  double* h_free = nullptr;    // pinned staging copy of the free vector
  double* h_con = nullptr;
  double* h_jacs[2] = {nullptr, nullptr};  // [1] only with prefetch_jac: speculative copies never touch the
                                            // buffer the caller may still be reading
  double* h_jac = nullptr;                 // buffer holding the most recently fetched Jacobian
  int jac_cur = 0;                         // its index
  int jac_target = 0;                      // destination of the copy in flight
  bool full_fetch[2] = {true, true};       // constant Jacobian columns not yet in that host buffer
  double* ext_con = nullptr;               // full-problem host vectors of opty_colloc_set_host_outputs
  double* ext_jac = nullptr;
  bool ext_full_fetch = true;

  bool known_set = false;
  bool free_valid = false;
  bool inv_dirty = true;
  bool evaluated = false;
  bool con_fetched = false, jac_fetched = false;
  bool pending_con = false, pending_jac = false;  // opty_colloc_begin without opty_colloc_finish

  // quadrature
  double* d_quad_partial = nullptr;
  double* d_quad_out = nullptr;
  double* d_quad_grad = nullptr;
  double* h_quad = nullptr;
  int quad_blocks = 0;

  std::vector<int32_t> d2h_begin, d2h_end;
  unsigned smem_bytes = 0;
  unsigned grid_x = 0;
  int64_t launches = 0;
  bool have_ms = false;
};

namespace {

int encode_2d(CUtensorMap* map, void* base, uint64_t dim0, uint64_t dim1, uint64_t stride1_bytes,
              uint32_t box0, uint32_t box1) {
  cuuint64_t gdim[2] = {dim0, dim1};
  cuuint64_t gstr[1] = {stride1_bytes};
  cuuint32_t box[2] = {box0, box1};
  cuuint32_t estr[2] = {1, 1};
  DRV_CHECK(g_drv.TensorMapEncodeTiled(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, base, gdim, gstr, box, estr,
                                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                       CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE));
  return OPTY_OK;
}

int build_tmaps(opty_colloc* h, opty_colloc::Module& m, int slot) {
  std::vector<unsigned char>& blob = m.tmaps[slot];
  blob.assign(sizeof(CUtensorMap) * (1 + (m.nmaps > 0 ? m.nmaps : 1)), 0);
  CUtensorMap* maps = reinterpret_cast<CUtensorMap*>(blob.data());
  int rc;
  if (h->tma_load == 1) {
    const uint32_t threads = 32u * h->warps;
    const uint32_t xbox = (threads <= 128u ? threads : 128u) + 2u;
    // (row-stationary kernel: a window of the rows, the height of the largest one any group reads)
    const uint32_t xrows = h->persistent == 2 ? (uint32_t)h->stat_xrows : (uint32_t)h->RD;
    if ((rc = encode_2d(&maps[0], h->d_traj, (uint64_t)h->ncols, (uint64_t)h->RD, (uint64_t)h->ldt * 8, xbox, xrows)))
      return rc;
  }
  if (h->tma_store) {
    // one map per sub-tile width over the whole node-major block: {K, nodes}, box {w, 32}; node rows
    // beyond the shard are clipped by the TMA unit
    for (int i = 0; i < m.nmaps; ++i) {
      if ((rc = encode_2d(&maps[1 + i], h->d_jac[slot], (uint64_t)h->K, (uint64_t)h->nn, (uint64_t)h->K * 8,
                          (uint32_t)m.widths[i], 32u)))
        return rc;
    }
  }
  return OPTY_OK;
}

int read_module_info(CUmodule mod, int* info) {
  CUdeviceptr sym = 0;
  size_t bytes = 0;
  DRV_CHECK(g_drv.ModuleGetGlobal(&sym, &bytes, mod, "opty_module_info"));
  if (bytes < sizeof(int) * INFO_WORDS) return fail(OPTY_ERR_ARG, "module info table too small");
  RT_CHECK(cudaMemcpy(info, reinterpret_cast<void*>(sym), sizeof(int) * INFO_WORDS, cudaMemcpyDeviceToHost));
  if (info[INFO_MAGIC] != kInfoMagic || info[INFO_VERSION] != kEmitterVersion)
    return fail(OPTY_ERR_ARG, "module was not generated by this version of the opty_b200 emitter");
  return OPTY_OK;
}

int load_module(opty_colloc* h, const void* cubin, bool primary) {
  opty_colloc::Module m;
  DRV_CHECK(g_drv.ModuleLoadData(&m.mod, cubin));
  int info[INFO_WORDS];
  int rc = read_module_info(m.mod, info);
  auto bail = [&](int code) {
    g_drv.ModuleUnload(m.mod);
    return code;
  };
  if (rc) return bail(rc);
  if (info[INFO_M] != h->cfg.M || info[INFO_P] != h->cfg.P || info[INFO_R] != h->R)
    return bail(fail(OPTY_ERR_ARG, "module was generated for a different problem (M, P or trajectory rows)"));
  if (primary) {
    if (!info[INFO_HAS_AUX]) return bail(fail(OPTY_ERR_ARG, "the first module must carry the invariants and pre-pass kernels"));
    h->warps = info[INFO_WARPS];
    h->num_derived = info[INFO_DERIVED];
    h->pre_groups = info[INFO_PRE_GROUPS];
    h->tma_load = info[INFO_TMA_LOAD];
    h->tma_store = info[INFO_TMA_STORE];
    h->tile_bufs = info[INFO_TILE_BUFS];
    h->tile_doubles = info[INFO_TILE_DOUBLES];
    h->num_inv = info[INFO_NUM_INV];
    h->persistent = info[INFO_PERSISTENT];
    h->stat_smem = info[INFO_SMEM_BYTES];
    h->stat_slots = info[INFO_SLOTS];
    h->stat_tiles = info[INFO_TILES];
    h->stat_cval0 = info[INFO_CVAL0];
    h->fused_pre = info[INFO_FUSED_PRE];
    h->stat_xrows = info[INFO_XROWS];
    h->num_const = info[INFO_NCONST];
    if (h->num_const > 0) h->stat_cval0 = info[INFO_CVAL0];
    h->const_kernel = info[INFO_CONST_KERNEL];
    if (h->persistent == 2 && (h->stat_smem < 1 || h->stat_slots < 1 || h->stat_tiles < 1 || h->stat_xrows < 1 ||
                               h->stat_xrows > 256))
      return bail(fail(OPTY_ERR_ARG, "invalid row-stationary geometry in the module info table"));
    if (h->warps < 1 || h->warps > 32 || h->tile_bufs < 1 || h->tile_bufs > 2 || h->tile_doubles < 64 ||
        h->num_derived < 0 || (h->num_derived > 0 && h->pre_groups < 1))
      return bail(fail(OPTY_ERR_ARG, "invalid kernel geometry in the module info table"));
  } else if (info[INFO_WARPS] != h->warps || info[INFO_DERIVED] != h->num_derived ||
             info[INFO_TMA_LOAD] != h->tma_load || info[INFO_TMA_STORE] != h->tma_store ||
             info[INFO_TILE_BUFS] != h->tile_bufs || info[INFO_TILE_DOUBLES] != h->tile_doubles ||
             info[INFO_NUM_INV] != h->num_inv || info[INFO_PERSISTENT] != h->persistent) {
    return bail(fail(OPTY_ERR_ARG, "additional module does not match the geometry of the first one"));
  }
  m.num_groups = info[INFO_GROUPS];
  m.nmaps = info[INFO_NMAPS];

  if (m.num_groups < 1 || m.num_groups > OPTY_MAX_GROUPS || m.nmaps < 1 || m.nmaps > kMaxMaps)
    return bail(fail(OPTY_ERR_ARG, "invalid group / tensor-map count in the module info table"));
  for (int i = 0; i < m.nmaps; ++i) {
    m.widths[i] = info[INFO_MAP_WIDTH0 + i];
    if (m.widths[i] < 1 || m.widths[i] > 256 || (h->tma_store && (m.widths[i] & 1)))
      return bail(fail(OPTY_ERR_ARG, "invalid staging tile width in the module info table"));
  }
  CUresult r1 = g_drv.ModuleGetFunction(&m.f_eval, m.mod, "opty_colloc_eval");
  size_t ci_bytes = 0;
  CUresult r2 = r1 == CUDA_SUCCESS ? g_drv.ModuleGetGlobal(&m.ci_sym, &ci_bytes, m.mod, "opty_ci") : r1;
  if (r2 != CUDA_SUCCESS) return bail(fail(OPTY_ERR_CUDA, "module lacks opty_colloc_eval / opty_ci: " + drv_err(r2)));
  // (the values of the constant column runs sit behind the invariants in d_inv and are read from there)
  const size_t ci_count = h->num_const > 0 ? (size_t)h->stat_cval0 : (size_t)h->num_inv;
  if (ci_bytes < ci_count * 8) return bail(fail(OPTY_ERR_ARG, "module's invariant table is too small"));
  if (primary) {
    CUresult r3 = g_drv.ModuleGetFunction(&h->f_inv, m.mod, "opty_colloc_inv");
    if (r3 == CUDA_SUCCESS) r3 = g_drv.ModuleGetFunction(&h->f_pre, m.mod, "opty_colloc_pre");
    if (r3 == CUDA_SUCCESS && h->num_const > 0 && h->const_kernel) {
      r3 = g_drv.ModuleGetFunction(&h->f_const, m.mod, "opty_colloc_const");
      if (r3 == CUDA_SUCCESS)
        r3 = g_drv.FuncSetAttribute(h->f_const, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, h->num_const * 8);
    }
    if (r3 != CUDA_SUCCESS) return bail(fail(OPTY_ERR_CUDA, "module lacks the invariants / pre-pass kernels: " + drv_err(r3)));
  }
  h->modules.push_back(std::move(m));
  return OPTY_OK;
}

int finish_module(opty_colloc* h, opty_colloc::Module& m) {
  DRV_CHECK(g_drv.FuncSetAttribute(m.f_eval, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)h->smem_bytes));
  m.tmaps.resize(h->cfg.out_ring);
  for (int s = 0; s < h->cfg.out_ring; ++s) {
    int rc = build_tmaps(h, m, s);
    if (rc) return rc;
  }
  if (h->persistent == 2) {
    // one block per slot of the schedule the emitter laid out; `d_work` holds the launch number that last
    // claimed each slot
    m.persist_grid = (unsigned)h->stat_slots;
    RT_CHECK(cudaMalloc(&m.d_work, (size_t)h->stat_slots * sizeof(int)));
    RT_CHECK(cudaMemset(m.d_work, 0, (size_t)h->stat_slots * sizeof(int)));
    RT_CHECK(cudaMalloc(&m.d_ready, (size_t)(h->stat_tiles + 1) * sizeof(unsigned long long)));
    RT_CHECK(cudaMemset(m.d_ready, 0, (size_t)(h->stat_tiles + 1) * sizeof(unsigned long long)));
  } else if (h->persistent) {
    // as many blocks as fit on the device at once: every block loops over node tiles
    int per_sm = 0;
    DRV_CHECK(g_drv.OccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, m.f_eval, 32 * h->warps, h->smem_bytes));
    if (per_sm < 1) return fail(OPTY_ERR_ARG, "the persistent kernel does not fit on an SM");
    m.persist_grid = (unsigned)per_sm * (unsigned)h->num_sms;
    RT_CHECK(cudaMalloc(&m.d_work, (size_t)(m.num_groups + 1) * sizeof(int)));
    RT_CHECK(cudaMemset(m.d_work, 0, (size_t)(m.num_groups + 1) * sizeof(int)));
  }
  return OPTY_OK;
}

int launch_eval(opty_colloc* h, bool record_events = false) {
  const opty_colloc_cfg& c = h->cfg;
  if (!h->known_set) return fail(OPTY_ERR_STATE, "opty_colloc_set_known must be called before evaluating");
  if (!h->free_valid) return fail(OPTY_ERR_STATE, "no free vector resident on the device");
  NvtxRange range("opty_b200:eval");
  {
    // Speculative Jacobian copies of earlier evaluations may still be reading their ring slot (IPOPT's
    // line search asks for g at trial points without ever asking for jac_g there).  The new kernels only
    // have to wait when they are about to overwrite a slot such a copy reads -- with out_ring >= 2 a
    // rejected trial point does not stall the next evaluation behind 40-80 MB of PCIe traffic.
    const int next_slot = (h->ring + 1) % c.out_ring;
    if (h->slot_copying.size() != (size_t)c.out_ring) h->slot_copying.assign(c.out_ring, 0);
    if (h->slot_copying[next_slot]) {
      RT_CHECK(cudaStreamWaitEvent(h->stream, h->ev_copy, 0));  // ev_copy follows every copy queued so far
      h->slot_copying.assign(c.out_ring, 0);
      h->jac_inflight = false;
    }
  }
  // per-evaluation timing events only on request: every event is one more operation in the stream
  if (record_events) RT_CHECK(cudaEventRecord(h->ev0, h->stream));
  if (h->inv_dirty && h->num_inv > 0) {
    void* args[2] = {&h->d_uni, &h->d_inv};
    DRV_CHECK(g_drv.LaunchKernel(h->f_inv, 1, 1, 1, 32, 1, 1, 0, (CUstream)h->stream, args, nullptr));
    h->launches++;
    for (auto& m : h->modules)
      if (h->num_const == 0 || h->stat_cval0 > 0)
        RT_CHECK(cudaMemcpyAsync(reinterpret_cast<void*>(m.ci_sym), h->d_inv,
                                 (size_t)(h->num_const > 0 ? h->stat_cval0 : h->num_inv) * 8,
                                 cudaMemcpyDeviceToDevice, h->stream));
  }
  h->inv_dirty = false;
  h->ring = (h->ring + 1) % c.out_ring;
  h->eval_seq++;
  OptyParams p;
  p.traj = h->d_traj;
  p.tiled = h->d_tiled;
  p.con = h->d_con[h->ring];
  p.jac = h->d_jac[h->ring];
  p.ldt = h->ldt;
  p.ldc = h->nn;
  p.n_nodes = h->nn;
  p.n_cols = h->ncols;
  p.n_tiles = (int)h->grid_x;
  p.work = nullptr;
  p.ready = nullptr;
  p.epoch = (int)(h->eval_seq & 0x7fffffff);  // launches of this handle so far, this one included
  p.cvals = h->d_inv + h->stat_cval0;
  bool pre_launched = false;
  {
    // pre-pass: derived rows, and for direct-input modules the tile-by-tile copy of the trajectory rows
    // (chunks of 16 rows in grid.y behind the groups of derived rows)
    const unsigned copy_groups = h->tma_load == 2 ? (unsigned)((h->R + 15) / 16) : 0u;
    const unsigned gy = (unsigned)h->pre_groups + copy_groups;
    if (gy > 0 && !h->fused_pre) {
      pre_launched = getenv("OPTY_B200_NO_PDL") == nullptr;
      void* pargs[1] = {&p};
      if (pre_launched && h->persistent == 2) {
        // Programmatic serialisation is a permission: behind a copy or an event the launch is ordered as
        // usual; directly behind the main kernel of the previous evaluation (back-to-back evaluations of a
        // resident point) the launch overhead overlaps that kernel's tail.  The pre-pass reads nothing that
        // kernel writes and starts only after its last block has exited (the main kernel never triggers
        // its dependents early), so it cannot overwrite derived rows that are still being read.
        CUlaunchConfig lc;
        memset(&lc, 0, sizeof(lc));
        lc.gridDimX = (unsigned)((h->nn + 1 + 127) / 128);
        lc.gridDimY = gy;
        lc.gridDimZ = 1;
        lc.blockDimX = 128;
        lc.blockDimY = lc.blockDimZ = 1;
        lc.sharedMemBytes = 0;
        lc.hStream = (CUstream)h->stream;
        CUlaunchAttribute attr;
        memset(&attr, 0, sizeof(attr));
        attr.id = CU_LAUNCH_ATTRIBUTE_PROGRAMMATIC_STREAM_SERIALIZATION;
        attr.value.programmaticStreamSerializationAllowed = 1;
        lc.attrs = &attr;
        lc.numAttrs = 1;
        DRV_CHECK(g_drv.LaunchKernelEx(&lc, h->f_pre, pargs, nullptr));
      } else {
        DRV_CHECK(g_drv.LaunchKernel(h->f_pre, (unsigned)((h->nn + 1 + 127) / 128), gy, 1, 128, 1, 1, 0,
                                     (CUstream)h->stream, pargs, nullptr));
      }
      h->launches++;
    }
  }
  if (h->f_const) {
    // constant column runs (grid kernel): one persistent wave of 256-thread blocks, bulk copies from a copy of
    // the runs in shared memory
    void* cargs[1] = {&p};
    DRV_CHECK(g_drv.LaunchKernel(h->f_const, (unsigned)h->num_sms, 1, 1, 256, 1, 1, (unsigned)h->num_const * 8u,
                                 (CUstream)h->stream, cargs, nullptr));
    h->launches++;
  }
  for (auto& m : h->modules) {
    p.work = m.d_work;
    p.ready = m.d_ready;
    void* args[2] = {m.tmaps[h->ring].data(), &p};
    if (h->persistent == 2 && pre_launched) {
      // row-stationary kernel behind the pre-pass: programmatic dependent launch -- its blocks set themselves
      // up while the pre-pass drains and wait (griddepcontrol.wait) before they fetch their first input rows
      CUlaunchConfig lc;
      memset(&lc, 0, sizeof(lc));
      lc.gridDimX = m.persist_grid;
      lc.gridDimY = lc.gridDimZ = 1;
      lc.blockDimX = 32u * h->warps;
      lc.blockDimY = lc.blockDimZ = 1;
      lc.sharedMemBytes = h->smem_bytes;
      lc.hStream = (CUstream)h->stream;
      CUlaunchAttribute attr;
      memset(&attr, 0, sizeof(attr));
      attr.id = CU_LAUNCH_ATTRIBUTE_PROGRAMMATIC_STREAM_SERIALIZATION;
      attr.value.programmaticStreamSerializationAllowed = 1;
      lc.attrs = &attr;
      lc.numAttrs = 1;
      DRV_CHECK(g_drv.LaunchKernelEx(&lc, m.f_eval, args, nullptr));
    } else if (h->persistent) {
      // code-stationary persistent kernel: resident blocks pull (group, tile) work items
      DRV_CHECK(g_drv.LaunchKernel(m.f_eval, m.persist_grid, 1, 1, 32u * h->warps, 1, 1, h->smem_bytes,
                                   (CUstream)h->stream, args, nullptr));
    } else {
      DRV_CHECK(g_drv.LaunchKernel(m.f_eval, h->grid_x, (unsigned)m.num_groups, 1, 32u * h->warps, 1, 1,
                                   h->smem_bytes, (CUstream)h->stream, args, nullptr));
    }
    h->launches++;
  }
  if (record_events) {
    RT_CHECK(cudaEventRecord(h->ev1, h->stream));
    h->have_ms = true;
  }
  h->evaluated = true;
  h->con_fetched = h->jac_fetched = false;
  return OPTY_OK;
}

int ensure_host_jac(opty_colloc* h) {
  const size_t bytes = ((size_t)h->nn * h->K + h->cfg.jac_tail) * 8;
  if (!h->h_jacs[0]) {
    RT_CHECK(cudaHostAlloc(&h->h_jacs[0], bytes, cudaHostAllocDefault));
    h->h_jac = h->h_jacs[0];
    h->full_fetch[0] = true;
  }
  if (h->cfg.prefetch_jac && !h->h_jacs[1]) {
    RT_CHECK(cudaHostAlloc(&h->h_jacs[1], bytes, cudaHostAllocDefault));
    h->full_fetch[1] = true;
  }
  return OPTY_OK;
}

// copies the Jacobian block of the current ring slot to `dst` (node-major, row pitch K), all columns or
// only the registered ranges of call-dependent columns
int copy_jac_block(opty_colloc* h, double* dst, bool full, cudaStream_t st) {
  NvtxRange range("opty_b200:jac_d2h");
  if (h->d2h_begin.empty() || full) {
    RT_CHECK(cudaMemcpyAsync(dst, h->d_jac[h->ring], (size_t)h->nn * h->K * 8, cudaMemcpyDeviceToHost, st));
  } else {
    for (size_t i = 0; i < h->d2h_begin.size(); ++i) {
      const int b = h->d2h_begin[i], e = h->d2h_end[i];
      RT_CHECK(cudaMemcpy2DAsync(dst + b, (size_t)h->K * 8, h->d_jac[h->ring] + b, (size_t)h->K * 8,
                                 (size_t)(e - b) * 8, h->nn, cudaMemcpyDeviceToHost, st));
    }
  }
  return OPTY_OK;
}

int enqueue_jac_copy(opty_colloc* h, cudaStream_t st) {
  int rc0 = ensure_host_jac(h);
  if (rc0) return rc0;
  const int which = h->h_jacs[1] ? (h->jac_cur ^ 1) : 0;
  h->jac_target = which;
  // the first fetch into a buffer brings the constant columns to the host once
  int rc = copy_jac_block(h, h->h_jacs[which], h->full_fetch[which], st);
  if (rc) return rc;
  h->full_fetch[which] = false;
  return OPTY_OK;
}

int upload(opty_colloc* h, const double* free_host, bool* changed_out) {
  const opty_colloc_cfg& c = h->cfg;
  const int nrows = c.n + c.q;
  // IPOPT evaluates g and jac_g at the same point back to back: compare with
  // the staged copy (a vector handed over in the pinned buffer itself cannot
  // be compared and always counts as new).  Only the part of the free vector
  // this handle evaluates is compared and staged: its column window of every
  // trajectory row plus the parameter / time-interval tail, so a shard's host
  // work does not grow with the size of the whole problem.
  const size_t win_bytes = (size_t)h->ncols * 8;
  const size_t tail_off = (size_t)nrows * c.N;
  const size_t tail_bytes = (size_t)(c.r + c.s) * 8;
  bool changed = !h->free_valid || free_host == h->h_free;
  if (!changed) {
    for (int r = 0; r < nrows && !changed; ++r) {
      const size_t off = (size_t)r * c.N + c.node_lo;
      changed = memcmp(free_host + off, h->h_free + off, win_bytes) != 0;
    }
    if (!changed && tail_bytes) changed = memcmp(free_host + tail_off, h->h_free + tail_off, tail_bytes) != 0;
  }
  if (changed_out) *changed_out = changed;
  if (!changed) return OPTY_OK;
  NvtxRange range("opty_b200:upload");
  if (free_host != h->h_free) {
    for (int r = 0; r < nrows; ++r) {
      const size_t off = (size_t)r * c.N + c.node_lo;
      memcpy(h->h_free + off, free_host + off, win_bytes);
    }
    if (tail_bytes) memcpy(h->h_free + tail_off, free_host + tail_off, tail_bytes);
  }
  // rows of the free vector are [row][N]; this handle keeps columns node_lo..node_hi
  if (nrows > 0)
    RT_CHECK(cudaMemcpy2DAsync(h->d_traj, (size_t)h->ldt * 8, h->h_free + c.node_lo, (size_t)c.N * 8,
                               (size_t)h->ncols * 8, nrows, cudaMemcpyHostToDevice, h->stream));
  if (c.r + c.s > 0) {
    RT_CHECK(cudaMemcpyAsync(h->d_uni + c.pk, h->h_free + (size_t)nrows * c.N, (size_t)(c.r + c.s) * 8,
                             cudaMemcpyHostToDevice, h->stream));
    h->inv_dirty = true;
  }
  h->free_valid = true;
  h->evaluated = false;
  return OPTY_OK;
}

// residual D2H into the handle's own buffer or into its M strided segments of the full-problem vector
int enqueue_con_copy(opty_colloc* h) {
  NvtxRange range("opty_b200:con_d2h");
  const opty_colloc_cfg& c = h->cfg;
  if (h->ext_con) {
    RT_CHECK(cudaMemcpy2DAsync(h->ext_con + c.node_lo, (size_t)(c.N - 1) * 8, h->d_con[h->ring], (size_t)h->nn * 8,
                               (size_t)h->nn * 8, c.M, cudaMemcpyDeviceToHost, h->stream));
  } else {
    RT_CHECK(cudaMemcpyAsync(h->h_con, h->d_con[h->ring], (size_t)c.M * h->nn * 8, cudaMemcpyDeviceToHost,
                             h->stream));
  }
  return OPTY_OK;
}

}  // namespace

extern "C" {

const char* opty_colloc_last_error(void) { return g_err.c_str(); }

int opty_b200_abi_version(void) { return OPTY_B200_ABI_VERSION; }

int opty_host_alloc(size_t bytes, void** ptr) {
  if (!ptr || bytes == 0) return fail(OPTY_ERR_ARG, "invalid argument");
  RT_CHECK(cudaHostAlloc(ptr, bytes, cudaHostAllocPortable));
  return OPTY_OK;
}

int opty_host_free(void* ptr) {
  if (ptr) RT_CHECK(cudaFreeHost(ptr));
  return OPTY_OK;
}

int opty_colloc_create(const opty_colloc_cfg* cfg, const void* cubin, size_t cubin_bytes, opty_colloc_t** out) {
  if (!cfg || !cubin || !out || cubin_bytes == 0) return fail(OPTY_ERR_ARG, "null argument");
  *out = nullptr;
  if (cfg->abi_version != OPTY_B200_ABI_VERSION) return fail(OPTY_ERR_ARG, "ABI version mismatch");
  const bool elementwise = cfg->method == OPTY_ELEMENTWISE;
  if (cfg->method != OPTY_MIDPOINT && cfg->method != OPTY_BACKWARD_EULER && !elementwise)
    return fail(OPTY_ERR_ARG, "invalid method");
  if (cfg->N < (elementwise ? 1 : 2) || cfg->n < (elementwise ? 0 : 1) || cfg->M < 1 || cfg->P < 1 || cfg->q < 0 ||
      cfg->k < 0 || cfg->r < 0 || cfg->pk < 0 || (cfg->s != 0 && cfg->s != 1) || cfg->con_tail < 0 ||
      cfg->jac_tail < 0)
    return fail(OPTY_ERR_ARG, "invalid problem dimensions");
  if (cfg->node_lo < 0 || cfg->node_hi > cfg->N - (elementwise ? 0 : 1) || cfg->node_lo >= cfg->node_hi)
    return fail(OPTY_ERR_ARG, "invalid node range");
  if (cfg->out_ring < 1 || cfg->out_ring > 64) return fail(OPTY_ERR_ARG, "invalid out_ring");
  const int expectP = (cfg->method == OPTY_MIDPOINT ? 2 * cfg->n + 2 * cfg->q : 2 * cfg->n + cfg->q) + cfg->r + cfg->s;
  if (!elementwise && cfg->P != expectP)
    return fail(OPTY_ERR_ARG, "P does not match n, q, r, s and the integration method");

  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(OPTY_ERR_CUDA, std::string("no CUDA device available: ") + cudaGetErrorString(e));
  if (cfg->device < 0 || cfg->device >= ndev) return fail(OPTY_ERR_ARG, "invalid device ordinal");
  RT_CHECK(cudaSetDevice(cfg->device));
  RT_CHECK(cudaFree(0));
  int rc = init_driver();
  if (rc) return rc;

  opty_colloc* h = new opty_colloc();
  h->cfg = *cfg;
  h->elementwise = elementwise;
  h->nn = cfg->node_hi - cfg->node_lo;
  // one column more than nodes: the neighbour of the last node (elementwise handles keep it as a zero
  // column so that midpoint-rule integrands may read it at the last point, whose weight is zero)
  h->ncols = h->nn + 1;
  h->R = cfg->n + cfg->q + cfg->k;
  h->K = cfg->M * cfg->P;
  h->free_len = (size_t)(cfg->n + cfg->q) * cfg->N + cfg->r + cfg->s;

#define CREATE_CHECK(stmt)        \
  do {                            \
    int rc__ = (stmt);            \
    if (rc__) {                   \
      opty_colloc_destroy(h);     \
      return rc__;                \
    }                             \
  } while (0)
#define CREATE_RT(expr) CREATE_CHECK([&]() -> int { RT_CHECK(expr); return OPTY_OK; }())
#define CREATE_DRV(expr) CREATE_CHECK([&]() -> int { DRV_CHECK(expr); return OPTY_OK; }())

  CREATE_CHECK(load_module(h, cubin, /*primary=*/true));
  h->RD = h->R + h->num_derived;
  h->ldt = round_up(h->ncols, 16);
  if (elementwise) h->ncols = h->nn;  // valid columns; column nn exists and stays zero
  if (h->tma_load == 1 && h->RD > 256) {
    opty_colloc_destroy(h);
    return fail(OPTY_ERR_ARG, "TMA input staging supports at most 256 trajectory rows");
  }
  if (h->tma_store && ((h->K & 1) != 0)) {
    opty_colloc_destroy(h);
    return fail(OPTY_ERR_ARG, "TMA Jacobian stores need an even M*P");
  }
  CREATE_RT(cudaDeviceGetAttribute(&h->num_sms, cudaDevAttrMultiProcessorCount, cfg->device));

  CREATE_RT(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  CREATE_RT(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
  CREATE_RT(cudaEventCreate(&h->ev0));
  CREATE_RT(cudaEventCreate(&h->ev1));
  CREATE_RT(cudaEventCreateWithFlags(&h->ev_copy, cudaEventDisableTiming));
  CREATE_RT(cudaEventCreateWithFlags(&h->ev_con, cudaEventDisableTiming));

  // (+ one input row of the widest block: the row-stationary kernel's bulk copies fetch whole rows of
  // 32*W + 2 columns, the last tile's reach past the valid columns)
  const size_t traj_bytes = (size_t)h->RD * h->ldt * 8 + (32 * 32 + 2) * 8;
  CREATE_RT(cudaMalloc(&h->d_traj, traj_bytes));
  CREATE_RT(cudaMemsetAsync(h->d_traj, 0, traj_bytes, h->stream));
  const int nuni = cfg->pk + cfg->r + 1;
  CREATE_RT(cudaMalloc(&h->d_uni, (size_t)nuni * 8));
  CREATE_RT(cudaMemsetAsync(h->d_uni, 0, (size_t)nuni * 8, h->stream));
  if (cfg->s == 0) {
    CREATE_RT(cudaMemcpyAsync(h->d_uni + cfg->pk + cfg->r, &cfg->h, 8, cudaMemcpyHostToDevice, h->stream));
    CREATE_RT(cudaStreamSynchronize(h->stream));
  }
  CREATE_RT(cudaMalloc(&h->d_inv, (size_t)(h->num_inv > 0 ? h->num_inv : 1) * 8));
  h->d_con.assign(cfg->out_ring, nullptr);
  h->d_jac.assign(cfg->out_ring, nullptr);
  for (int s = 0; s < cfg->out_ring; ++s) {
    CREATE_RT(cudaMalloc(&h->d_con[s], (size_t)cfg->M * h->nn * 8));
    CREATE_RT(cudaMalloc(&h->d_jac[s], (size_t)h->nn * h->K * 8));
  }

  CREATE_RT(cudaHostAlloc(&h->h_free, h->free_len * 8, cudaHostAllocDefault));
  CREATE_RT(cudaHostAlloc(&h->h_con, ((size_t)cfg->M * h->nn + cfg->con_tail) * 8, cudaHostAllocDefault));
  // the pinned Jacobian buffers (8.4 GB each at BASELINE config 5) are allocated on first use:
  // device-resident consumers never need them

  const unsigned tiles_bytes = (unsigned)h->warps * (unsigned)h->tile_bufs * (unsigned)h->tile_doubles * 8u;
  const unsigned threads = 32u * h->warps;
  const unsigned xseg = threads <= 128u ? threads : 128u;
  const unsigned nseg = threads / xseg;
  const unsigned xin_bytes =
      h->tma_load == 2 ? 0u : nseg * (unsigned)round_up((int64_t)h->RD * (xseg + 2u) * 8, 128);
  h->smem_bytes = tiles_bytes + xin_bytes + 128u;
  if (h->persistent == 2) h->smem_bytes = (unsigned)h->stat_smem;
  if (const char* pad = getenv("OPTY_B200_DEBUG_SMEM_FLOOR")) {
    // measurement aid: a larger dynamic shared-memory request caps the resident blocks per SM
    const unsigned floor_bytes = (unsigned)atoi(pad);
    if (floor_bytes > h->smem_bytes) h->smem_bytes = floor_bytes;
  }
  if (h->smem_bytes > 227u * 1024u) {
    opty_colloc_destroy(h);
    return fail(OPTY_ERR_ARG, "kernel needs more than 227 KB of shared memory per block");
  }
  CREATE_CHECK(finish_module(h, h->modules[0]));

  h->grid_x = (unsigned)((h->nn + 32 * h->warps - 1) / (32 * h->warps));
  if (h->persistent == 2 && (int)h->grid_x != h->stat_tiles) {
    opty_colloc_destroy(h);
    return fail(OPTY_ERR_ARG, "the module's static schedule was laid out for a different number of nodes");
  }
  if (h->tma_load == 2) {
    const size_t tiled_bytes = (size_t)h->grid_x * h->RD * (32 * h->warps + 2) * 8;
    CREATE_RT(cudaMalloc(&h->d_tiled, tiled_bytes));
    CREATE_RT(cudaMemsetAsync(h->d_tiled, 0, tiled_bytes, h->stream));
  }
  CREATE_RT(cudaStreamSynchronize(h->stream));
  *out = h;
  return OPTY_OK;
}

int opty_colloc_destroy(opty_colloc_t* h) {
  if (!h) return OPTY_OK;
  cudaSetDevice(h->cfg.device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  if (h->copy_stream) cudaStreamSynchronize(h->copy_stream);
  for (double* p : h->d_con) cudaFree(p);
  for (double* p : h->d_jac) cudaFree(p);
  cudaFree(h->d_traj);
  cudaFree(h->d_tiled);
  cudaFree(h->d_uni);
  cudaFree(h->d_inv);
  cudaFree(h->d_quad_partial);
  cudaFree(h->d_quad_out);
  cudaFree(h->d_quad_grad);
  if (h->h_quad) cudaFreeHost(h->h_quad);
  if (h->h_free) cudaFreeHost(h->h_free);
  if (h->h_con) cudaFreeHost(h->h_con);
  if (h->h_jacs[0]) cudaFreeHost(h->h_jacs[0]);
  if (h->h_jacs[1]) cudaFreeHost(h->h_jacs[1]);
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  if (h->ev_copy) cudaEventDestroy(h->ev_copy);
  if (h->ev_con) cudaEventDestroy(h->ev_con);
  if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
  if (h->stream) cudaStreamDestroy(h->stream);
  for (auto& m : h->modules) {
    cudaFree(m.d_work);
    cudaFree(m.d_ready);
    if (m.mod && g_drv.ModuleUnload) g_drv.ModuleUnload(m.mod);
  }
  delete h;
  return OPTY_OK;
}

int opty_colloc_add_module(opty_colloc_t* h, const void* cubin, size_t cubin_bytes) {
  if (!h || !cubin || cubin_bytes == 0) return fail(OPTY_ERR_ARG, "null argument");
  RT_CHECK(cudaSetDevice(h->cfg.device));
  int rc = load_module(h, cubin, /*primary=*/false);
  if (rc) return rc;
  rc = finish_module(h, h->modules.back());
  if (rc) {
    g_drv.ModuleUnload(h->modules.back().mod);
    h->modules.pop_back();
    return rc;
  }
  h->inv_dirty = true;
  h->evaluated = false;
  return OPTY_OK;
}

int opty_colloc_set_known(opty_colloc_t* h, const double* traj, const double* params) {
  if (!h) return fail(OPTY_ERR_ARG, "null handle");
  const opty_colloc_cfg& c = h->cfg;
  if ((c.k > 0 && !traj) || (c.pk > 0 && !params)) return fail(OPTY_ERR_ARG, "known values missing");
  RT_CHECK(cudaSetDevice(c.device));
  if (c.k > 0) {
    RT_CHECK(cudaMemcpy2DAsync(h->d_traj + (size_t)(c.n + c.q) * h->ldt, (size_t)h->ldt * 8, traj + c.node_lo,
                               (size_t)c.N * 8, (size_t)h->ncols * 8, c.k, cudaMemcpyHostToDevice, h->stream));
  }
  if (c.pk > 0) {
    RT_CHECK(cudaMemcpyAsync(h->d_uni, params, (size_t)c.pk * 8, cudaMemcpyHostToDevice, h->stream));
  }
  RT_CHECK(cudaStreamSynchronize(h->stream));
  h->known_set = true;
  h->inv_dirty = true;
  h->evaluated = false;
  h->full_fetch[0] = h->full_fetch[1] = true;
  h->ext_full_fetch = true;
  return OPTY_OK;
}

int opty_colloc_upload_free(opty_colloc_t* h, const double* free_host) {
  if (!h || !free_host) return fail(OPTY_ERR_ARG, "null argument");
  RT_CHECK(cudaSetDevice(h->cfg.device));
  int rc = upload(h, free_host, nullptr);
  if (rc) return rc;
  RT_CHECK(cudaStreamSynchronize(h->stream));
  return OPTY_OK;
}

int opty_colloc_eval_device(opty_colloc_t* h, int sync) {
  if (!h) return fail(OPTY_ERR_ARG, "null handle");
  RT_CHECK(cudaSetDevice(h->cfg.device));
  int rc = launch_eval(h, /*record_events=*/true);
  if (rc) return rc;
  if (sync) RT_CHECK(cudaStreamSynchronize(h->stream));
  return OPTY_OK;
}

int opty_colloc_constraints(opty_colloc_t* h, const double* free_host, double* con_host) {
  if (!h || !free_host) return fail(OPTY_ERR_ARG, "null argument");
  if (h->ext_con) return fail(OPTY_ERR_STATE, "host outputs are redirected: use opty_colloc_begin / opty_colloc_finish");
  RT_CHECK(cudaSetDevice(h->cfg.device));
  int rc = upload(h, free_host, nullptr);
  if (rc) return rc;
  bool launched = false;
  if (!h->evaluated) {
    if ((rc = launch_eval(h))) return rc;
    launched = true;
  }
  const size_t bytes = (size_t)h->cfg.M * h->nn * 8;
  if (!h->con_fetched && (rc = enqueue_con_copy(h))) return rc;
  if (launched && h->cfg.prefetch_jac) {
    // IPOPT asks for the Jacobian at the point it just evaluated g at: start
    // moving it now, on a second stream, behind the kernels and the residuals,
    // into the host buffer the caller is not looking at
    // (ordered behind the small residual copy: both go through the same
    // device-to-host copy engine, which would otherwise serve the big one first)
    RT_CHECK(cudaEventRecord(h->ev_con, h->stream));
    RT_CHECK(cudaStreamWaitEvent(h->copy_stream, h->ev_con, 0));
    if ((rc = enqueue_jac_copy(h, h->copy_stream))) return rc;
    RT_CHECK(cudaEventRecord(h->ev_copy, h->copy_stream));
    h->jac_inflight = true;
    h->copy_seq = h->eval_seq;
    if (h->slot_copying.size() != (size_t)h->cfg.out_ring) h->slot_copying.assign(h->cfg.out_ring, 0);
    h->slot_copying[h->ring] = 1;
  }
  if (!h->con_fetched) {
    RT_CHECK(cudaStreamSynchronize(h->stream));
    h->con_fetched = true;
  }
  if (con_host && con_host != h->h_con) memcpy(con_host, h->h_con, bytes);
  return OPTY_OK;
}

int opty_colloc_jacobian(opty_colloc_t* h, const double* free_host, double* jac_host) {
  if (!h || !free_host) return fail(OPTY_ERR_ARG, "null argument");
  if (h->ext_jac) return fail(OPTY_ERR_STATE, "host outputs are redirected: use opty_colloc_begin / opty_colloc_finish");
  RT_CHECK(cudaSetDevice(h->cfg.device));
  int rc = upload(h, free_host, nullptr);
  if (rc) return rc;
  if (!h->evaluated && (rc = launch_eval(h))) return rc;
  const size_t bytes = (size_t)h->nn * h->K * 8;
  if (!h->jac_fetched) {
    if (h->jac_inflight && h->copy_seq == h->eval_seq) {
      // the speculative copy started by constraints() at this very point
      RT_CHECK(cudaEventSynchronize(h->ev_copy));
      h->jac_inflight = false;
      h->slot_copying.assign(h->slot_copying.size(), 0);
    } else {
      if (h->jac_inflight) {
        // copies of other points are still in flight (rejected trial points): they target the same
        // pinned buffer, so this one is ordered behind them
        RT_CHECK(cudaStreamWaitEvent(h->stream, h->ev_copy, 0));
        h->jac_inflight = false;
        h->slot_copying.assign(h->slot_copying.size(), 0);
      }
      if ((rc = enqueue_jac_copy(h, h->stream))) return rc;
      RT_CHECK(cudaStreamSynchronize(h->stream));
    }
    h->jac_cur = h->jac_target;
    h->h_jac = h->h_jacs[h->jac_cur];
    h->jac_fetched = true;
  }
  if (jac_host && jac_host != h->h_jac) memcpy(jac_host, h->h_jac, bytes);
  return OPTY_OK;
}

int opty_colloc_set_host_outputs(opty_colloc_t* h, double* con_full, double* jac_full) {
  if (!h) return fail(OPTY_ERR_ARG, "null handle");
  if ((con_full == nullptr) != (jac_full == nullptr)) return fail(OPTY_ERR_ARG, "give both host vectors or neither");
  if (h->elementwise) return fail(OPTY_ERR_ARG, "elementwise handles have no full-problem layout");
  RT_CHECK(cudaSetDevice(h->cfg.device));
  RT_CHECK(cudaStreamSynchronize(h->stream));
  RT_CHECK(cudaStreamSynchronize(h->copy_stream));
  h->ext_con = con_full;
  h->ext_jac = jac_full;
  h->ext_full_fetch = true;
  h->con_fetched = h->jac_fetched = false;
  h->jac_inflight = false;
  h->slot_copying.assign(h->slot_copying.size(), 0);
  return OPTY_OK;
}

int opty_colloc_begin(opty_colloc_t* h, const double* free_host, int want_con, int want_jac) {
  if (!h || !free_host) return fail(OPTY_ERR_ARG, "null argument");
  if (!h->ext_con || !h->ext_jac) return fail(OPTY_ERR_STATE, "opty_colloc_set_host_outputs must be called first");
  RT_CHECK(cudaSetDevice(h->cfg.device));
  int rc = upload(h, free_host, nullptr);
  if (rc) return rc;
  if (!h->evaluated && (rc = launch_eval(h))) return rc;
  if (want_con && !h->con_fetched) {
    if ((rc = enqueue_con_copy(h))) return rc;
    h->pending_con = true;
  }
  if (want_jac && !h->jac_fetched) {
    double* dst = h->ext_jac + (size_t)h->cfg.node_lo * h->K;
    if ((rc = copy_jac_block(h, dst, h->ext_full_fetch, h->stream))) return rc;
    h->ext_full_fetch = false;
    h->pending_jac = true;
  }
  return OPTY_OK;
}

int opty_colloc_finish(opty_colloc_t* h) {
  if (!h) return fail(OPTY_ERR_ARG, "null handle");
  RT_CHECK(cudaSetDevice(h->cfg.device));
  RT_CHECK(cudaStreamSynchronize(h->stream));
  if (h->pending_con) h->con_fetched = true;
  if (h->pending_jac) h->jac_fetched = true;
  h->pending_con = h->pending_jac = false;
  return OPTY_OK;
}

int opty_colloc_host_buffers(opty_colloc_t* h, double** free_pinned, double** con_pinned, double** jac_pinned) {
  if (!h) return fail(OPTY_ERR_ARG, "null handle");
  if (jac_pinned) {
    RT_CHECK(cudaSetDevice(h->cfg.device));
    int rc = ensure_host_jac(h);
    if (rc) return rc;
  }
  if (free_pinned) *free_pinned = h->h_free;
  if (con_pinned) *con_pinned = h->h_con;
  if (jac_pinned) *jac_pinned = h->h_jac;
  return OPTY_OK;
}

int opty_colloc_device_buffers(opty_colloc_t* h, void** traj, int64_t* ldt, void** con, void** jac, void** uni) {
  if (!h) return fail(OPTY_ERR_ARG, "null handle");
  const int slot = h->ring < 0 ? 0 : h->ring;
  if (traj) *traj = h->d_traj;
  if (ldt) *ldt = h->ldt;
  if (con) *con = h->d_con[slot];
  if (jac) *jac = h->d_jac[slot];
  if (uni) *uni = h->d_uni;
  return OPTY_OK;
}

int opty_colloc_set_d2h_columns(opty_colloc_t* h, int num_ranges, const int32_t* col_begin, const int32_t* col_end) {
  if (!h || num_ranges < 0) return fail(OPTY_ERR_ARG, "invalid argument");
  std::vector<int32_t> b, e;
  int prev = 0;
  for (int i = 0; i < num_ranges; ++i) {
    if (!col_begin || !col_end || col_begin[i] < prev || col_end[i] <= col_begin[i] || col_end[i] > h->K)
      return fail(OPTY_ERR_ARG, "column ranges must be sorted, non-empty and inside [0, M*P)");
    b.push_back(col_begin[i]);
    e.push_back(col_end[i]);
    prev = col_end[i];
  }
  h->d2h_begin.swap(b);
  h->d2h_end.swap(e);
  return opty_colloc_invalidate_host_jacobian(h);
}

int opty_colloc_invalidate_host_jacobian(opty_colloc_t* h) {
  if (!h) return fail(OPTY_ERR_ARG, "null handle");
  h->jac_fetched = false;
  h->full_fetch[0] = h->full_fetch[1] = true;
  h->ext_full_fetch = true;
  if (h->jac_inflight) {
    // a speculative copy that skipped the constant columns must not be taken for the full one
    RT_CHECK(cudaSetDevice(h->cfg.device));
    RT_CHECK(cudaEventSynchronize(h->ev_copy));
    h->jac_inflight = false;
    h->slot_copying.assign(h->slot_copying.size(), 0);
  }
  return OPTY_OK;
}

int opty_colloc_quadrature(opty_colloc_t* h, const double* free_host, double scale, int rule, double* value,
                           double* grad) {
  if (!h || !free_host || !value || !grad) return fail(OPTY_ERR_ARG, "null argument");
  const opty_colloc_cfg& c = h->cfg;
  if (!h->elementwise || c.M != 1 || c.k != 0 || c.q != 0 || c.s != 0 || c.P != 1 + c.n + c.r)
    return fail(OPTY_ERR_ARG,
                "quadrature needs an elementwise handle whose module returns [f, df/darg.., df/dconst..] per point");
  if (rule != OPTY_BACKWARD_EULER && rule != OPTY_MIDPOINT) return fail(OPTY_ERR_ARG, "invalid quadrature rule");
  if (c.r > OPTY_QUAD_MAX_SCALARS) return fail(OPTY_ERR_ARG, "too many scalar arguments for the quadrature kernel");
  if (c.node_lo != 0 || c.node_hi != c.N) return fail(OPTY_ERR_ARG, "quadrature handles cover all nodes");
  RT_CHECK(cudaSetDevice(c.device));
  const int N = c.N;
  if (!h->d_quad_out) {
    h->quad_blocks = (N + OPTY_QUAD_THREADS - 1) / OPTY_QUAD_THREADS;
    if (h->quad_blocks > 4 * h->num_sms) h->quad_blocks = 4 * h->num_sms;
    RT_CHECK(cudaMalloc(&h->d_quad_partial, (size_t)h->quad_blocks * (1 + OPTY_QUAD_MAX_SCALARS) * 8));
    RT_CHECK(cudaMalloc(&h->d_quad_out, (size_t)(1 + OPTY_QUAD_MAX_SCALARS) * 8));
    RT_CHECK(cudaMalloc(&h->d_quad_grad, ((size_t)c.n * N + 1) * 8));
    RT_CHECK(cudaHostAlloc(&h->h_quad, ((size_t)c.n * N + 1 + OPTY_QUAD_MAX_SCALARS) * 8, cudaHostAllocDefault));
  }
  int rc = upload(h, free_host, nullptr);
  if (rc) return rc;
  if (!h->evaluated && (rc = launch_eval(h))) return rc;
  {
    NvtxRange range("opty_b200:quadrature");
    opty_quadrature_stage1<<<h->quad_blocks, OPTY_QUAD_THREADS, 0, h->stream>>>(
        h->d_jac[h->ring], N, c.P, c.n, c.r, rule, scale, h->d_quad_grad, h->d_quad_partial);
    opty_quadrature_stage2<<<1, 64, 0, h->stream>>>(h->d_quad_partial, h->quad_blocks, c.r, scale, h->d_quad_out);
    RT_CHECK(cudaGetLastError());
    h->launches += 2;
    RT_CHECK(cudaMemcpyAsync(h->h_quad, h->d_quad_out, (size_t)(1 + c.r) * 8, cudaMemcpyDeviceToHost, h->stream));
    if (c.n > 0)
      RT_CHECK(cudaMemcpyAsync(h->h_quad + 1 + OPTY_QUAD_MAX_SCALARS, h->d_quad_grad, (size_t)c.n * N * 8,
                               cudaMemcpyDeviceToHost, h->stream));
    RT_CHECK(cudaStreamSynchronize(h->stream));
  }
  *value = h->h_quad[0];
  memcpy(grad, h->h_quad + 1 + OPTY_QUAD_MAX_SCALARS, (size_t)c.n * N * 8);
  for (int s = 0; s < c.r; ++s) grad[(size_t)c.n * N + s] = h->h_quad[1 + s];
  return OPTY_OK;
}

int opty_colloc_last_kernel_ms(opty_colloc_t* h, float* ms) {
  if (!h || !ms) return fail(OPTY_ERR_ARG, "null argument");
  if (!h->have_ms) return fail(OPTY_ERR_STATE, "no evaluation has been launched yet");
  RT_CHECK(cudaEventSynchronize(h->ev1));
  RT_CHECK(cudaEventElapsedTime(ms, h->ev0, h->ev1));
  return OPTY_OK;
}

int opty_colloc_time_device_evals(opty_colloc_t* h, int steps, float* total_ms) {
  if (!h || !total_ms || steps < 1) return fail(OPTY_ERR_ARG, "invalid argument");
  RT_CHECK(cudaSetDevice(h->cfg.device));
  cudaEvent_t a, b;
  RT_CHECK(cudaEventCreate(&a));
  RT_CHECK(cudaEventCreate(&b));
  RT_CHECK(cudaStreamSynchronize(h->stream));
  RT_CHECK(cudaEventRecord(a, h->stream));
  int rc = OPTY_OK;
  for (int i = 0; i < steps && rc == OPTY_OK; ++i) rc = launch_eval(h);
  cudaEventRecord(b, h->stream);
  cudaError_t e = cudaEventSynchronize(b);
  if (rc == OPTY_OK && e == cudaSuccess) e = cudaEventElapsedTime(total_ms, a, b);
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  if (rc) return rc;
  if (e != cudaSuccess) return fail(OPTY_ERR_CUDA, std::string("timing failed: ") + cudaGetErrorString(e));
  return OPTY_OK;
}

int opty_colloc_launch_count(opty_colloc_t* h, int64_t* count) {
  if (!h || !count) return fail(OPTY_ERR_ARG, "null argument");
  *count = h->launches;
  return OPTY_OK;
}

int opty_colloc_jacobian_indices(int device, int N, int node_lo, int node_hi, int n, int q, int r, int s, int M,
                                 int method, int64_t* rows, int64_t* cols) {
  if (!rows || !cols) return fail(OPTY_ERR_ARG, "null argument");
  if (N < 2 || node_lo < 0 || node_hi > N - 1 || node_lo >= node_hi || n < 1 || q < 0 || r < 0 || s < 0 || M < 1)
    return fail(OPTY_ERR_ARG, "invalid dimensions");
  if (method != OPTY_MIDPOINT && method != OPTY_BACKWARD_EULER) return fail(OPTY_ERR_ARG, "invalid method");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(OPTY_ERR_CUDA, std::string("no CUDA device available: ") + cudaGetErrorString(e));
  RT_CHECK(cudaSetDevice(device));
  NvtxRange range("opty_b200:jacobian_indices");
  const long long P = (method == OPTY_MIDPOINT ? 2LL * n + 2LL * q : 2LL * n + q) + r + s;
  const long long MP = (long long)M * P;
  const long long first = (long long)node_lo * MP;
  const long long total = (long long)(node_hi - node_lo) * MP;
  const long long chunk = 1LL << 25;  // 32 Mi entries = 2 x 256 MiB of int64 per pass
  long long *d_rows = nullptr, *d_cols = nullptr;
  const long long cap = total < chunk ? total : chunk;
  RT_CHECK(cudaMalloc(&d_rows, (size_t)cap * 8));
  cudaError_t e2 = cudaMalloc(&d_cols, (size_t)cap * 8);
  if (e2 != cudaSuccess) {
    cudaFree(d_rows);
    return fail(OPTY_ERR_CUDA, std::string("cudaMalloc: ") + cudaGetErrorString(e2));
  }
  int rc = OPTY_OK;
  for (long long done = 0; done < total && rc == OPTY_OK; done += cap) {
    const long long cnt = (total - done) < cap ? (total - done) : cap;
    const int threads = 256;
    long long blocks = (cnt + threads - 1) / threads;
    if (blocks > 148LL * 16) blocks = 148LL * 16;
    opty_jac_indices_kernel<<<(unsigned)blocks, threads>>>(first + done, cnt, N, n, q, M, P, method, d_rows, d_cols);
    cudaError_t ek = cudaGetLastError();
    if (ek == cudaSuccess) ek = cudaMemcpy(rows + done, d_rows, (size_t)cnt * 8, cudaMemcpyDeviceToHost);
    if (ek == cudaSuccess) ek = cudaMemcpy(cols + done, d_cols, (size_t)cnt * 8, cudaMemcpyDeviceToHost);
    if (ek != cudaSuccess) rc = fail(OPTY_ERR_CUDA, std::string("jacobian_indices: ") + cudaGetErrorString(ek));
  }
  cudaFree(d_rows);
  cudaFree(d_cols);
  return rc;
}

}  // extern "C"
