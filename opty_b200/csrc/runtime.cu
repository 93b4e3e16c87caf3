// Host runtime behind include/opty_b200.h: owns the device-resident trajectory
// matrix, the residual / Jacobian buffers, pinned host buffers, the stream and
// the TMA descriptors, loads the generated sm_100a module and launches it.
//
// It replaces the NumPy / Cython glue of the reference's callback path
// (opty/utils.py:277-326 parse_free, opty/direct_collocation.py:2891-2926
// _merge_fixed_free, :2382-2446 constraints, :2816-2887 constraints_jacobian)
// and the Python index loop (opty/direct_collocation.py:2628-2684).
//
// Driver-API entry points (module loading, tensor-map encoding, launches) are
// resolved through cudaGetDriverEntryPoint so that this library has no
// load-time dependency on libcuda.so.1: it can be dlopen'ed on a machine
// without a GPU (symbol checks), and fails loudly at opty_colloc_create there.

#include <cuda.h>
#include <cuda_runtime.h>
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
  CUresult (*LaunchCooperativeKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned,
                                      unsigned, CUstream, void**) = nullptr;
  CUresult (*GetErrorString)(CUresult, const char**) = nullptr;
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
  if ((rc = load_entry("cuLaunchCooperativeKernel", &g_drv.LaunchCooperativeKernel))) return rc;
  if ((rc = load_entry("cuGetErrorString", &g_drv.GetErrorString))) return rc;
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


// ---- constant-run replicator ------------------------------------------------
// Column runs of the node block whose entries are literals or node-invariant
// are the same for every node (at the 10-link pendulum: the 506 partials of
// the 11 kinematic equations, half of all columns).  The generated group bodies
// skip them; this kernel builds one shared-memory image [rows][w] of a chunk
// (<= 254 columns of one run) and replicates it down the node rows with 2-D
// TMA tile stores (32 KB per store): no per-node arithmetic, no per-node
// staging, long contiguous row segments.  grid = (node-tile batches, chunks).
#define OPTY_REPL_THREADS 128
#define OPTY_REPL_MAX_CHUNKS 96
struct OptyReplMaps {
  CUtensorMap m[OPTY_REPL_MAX_CHUNKS];  // chunk c of jac as {w_c, nodes}, box {w_c, rows}
};

__global__ void __launch_bounds__(OPTY_REPL_THREADS)
opty_replicate_kernel(const __grid_constant__ OptyReplMaps maps, const double* __restrict__ lit,
                      const int* __restrict__ inv_idx, const double* __restrict__ inv,
                      const int* __restrict__ ch_off, const int* __restrict__ ch_w, int n_nodes, int rows,
                      int tiles_per_block) {
  extern __shared__ __align__(128) unsigned char repl_smem[];
  double* img = reinterpret_cast<double*>(repl_smem);
  const int c = blockIdx.y;
  const int w = ch_w[c];
  const int off = ch_off[c];
  for (int j = threadIdx.x; j < w; j += OPTY_REPL_THREADS) {
    const int k = inv_idx[off + j];
    const double v = k >= 0 ? inv[k] : lit[off + j];
    for (int r = 0; r < rows; ++r) img[r * w + j] = v;
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t src = (uint32_t)__cvta_generic_to_shared(img);
    for (int t = 0; t < tiles_per_block; ++t) {
      const int node0 = (blockIdx.x * tiles_per_block + t) * rows;
      if (node0 >= n_nodes) break;
      asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(
                       reinterpret_cast<uint64_t>(&maps.m[c])),
                   "r"(0), "r"(node0), "r"(src)
                   : "memory");
    }
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  }
}

// Same job with plain coalesced 16-byte stores: a warp writes a node row's run
// as consecutive 512-byte pieces straight from the shared-memory image (the
// fill pattern that reaches the measured HBM write ceiling).  grid.x = batches
// of `nodes_per_block` nodes.
__global__ void __launch_bounds__(OPTY_REPL_THREADS)
opty_replicate_st_kernel(double* __restrict__ jac, long long K, const double* __restrict__ lit,
                         const int* __restrict__ inv_idx, const double* __restrict__ inv,
                         const int* __restrict__ run_col0, const int* __restrict__ run_len,
                         const int* __restrict__ run_off, int num_runs, int ncc, int n_nodes,
                         int nodes_per_block) {
  extern __shared__ __align__(128) unsigned char repl_smem[];
  double* img = reinterpret_cast<double*>(repl_smem);
  int* tab = reinterpret_cast<int*>(img + ncc);  // [3][num_runs]: col0, len, off
  for (int j = threadIdx.x; j < ncc; j += OPTY_REPL_THREADS) {
    const int k = inv_idx[j];
    img[j] = k >= 0 ? inv[k] : lit[j];
  }
  for (int r = threadIdx.x; r < num_runs; r += OPTY_REPL_THREADS) {
    tab[r] = run_col0[r];
    tab[num_runs + r] = run_len[r];
    tab[2 * num_runs + r] = run_off[r];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int node_end = min(n_nodes, (int)(blockIdx.x + 1) * nodes_per_block);
  for (int r = 0; r < num_runs; ++r) {
    const int col0 = tab[r];
    const int n2 = tab[num_runs + r] >> 1;
    const double2* src = reinterpret_cast<const double2*>(img + tab[2 * num_runs + r]);
    // a lane keeps its 16-byte pieces of the run in registers across the nodes it writes
    for (int j0 = 0; j0 < n2; j0 += 32 * 8) {
      double2 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int j = j0 + u * 32 + lane;
        v[u] = j < n2 ? src[j] : make_double2(0.0, 0.0);
      }
      for (int node = blockIdx.x * nodes_per_block + warp; node < node_end; node += OPTY_REPL_THREADS / 32) {
        double2* dst = reinterpret_cast<double2*>(jac + (long long)node * K + col0);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int j = j0 + u * 32 + lane;
          if (j < n2) dst[j] = v[u];
        }
      }
    }
  }
}

}  // namespace

// mirrors OptyPersist of csrc/colloc_persistent.cuh
struct OptyPersistArgs {
  const void* sched;
  unsigned int* barrier;
  unsigned int barrier_target;
  long long* block_clocks;
  int pre_units;
};

struct opty_colloc {
  opty_colloc_cfg cfg;
  // persistent main kernel
  int sched_blocks = 0;
  void* d_sched = nullptr;          // int4 per block
  unsigned int* d_barrier = nullptr;
  long long* d_block_clocks = nullptr;
  unsigned int barrier_epoch = 0;
  int nn = 0;          // constraint nodes of this handle
  int ncols = 0;       // trajectory columns held (nn + 1)
  int R = 0;           // trajectory rows n + q + k
  int K = 0;           // M * P
  int64_t ldt = 0;
  size_t free_len = 0;

  CUmodule mod = nullptr;
  CUfunction f_eval = nullptr;
  CUfunction f_inv = nullptr;
  CUfunction f_pre = nullptr;
  int num_sms = 0;
  int RD = 0;          // trajectory rows incl. derived rows
  int n_tiles = 0;
  CUdeviceptr ci_sym = 0;
  size_t ci_bytes = 0;

  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr;   // speculative Jacobian D2H
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  cudaEvent_t ev_copy = nullptr, ev_con = nullptr;
  bool jac_inflight = false;
  std::vector<char> slot_copying;       // ring slots a speculative copy may still be reading
  uint64_t eval_seq = 0;                // evaluations launched so far
  uint64_t copy_seq = 0;                // evaluation the most recent speculative copy belongs to

  double* d_traj = nullptr;
  double* d_uni = nullptr;
  double* d_inv = nullptr;
  std::vector<double*> d_con, d_jac;
  std::vector<std::vector<unsigned char>> tmaps;  // per ring slot: OptyTmaps blob
  int ring = -1;

  double* h_free = nullptr;    // pinned staging copy of the free vector
  double* h_con = nullptr;
  double* h_jacs[2] = {nullptr, nullptr};  // [1] only with prefetch_jac: speculative copies never touch the
                                            // buffer the caller may still be reading
  double* h_jac = nullptr;                 // buffer holding the most recently fetched Jacobian
  int jac_cur = 0;                         // its index
  int jac_target = 0;                      // destination of the copy in flight
  bool full_fetch[2] = {true, true};       // constant Jacobian columns not yet in that host buffer

  bool known_set = false;
  bool free_valid = false;
  bool inv_dirty = true;
  bool evaluated = false;
  bool con_fetched = false, jac_fetched = false;

  // additional modules of a problem compiled in several pieces (groups beyond the primary module's)
  struct ExtraModule {
    CUmodule mod = nullptr;
    CUfunction f_eval = nullptr;
    CUdeviceptr ci_sym = 0;
    int seg_first = 0, seg_count = 0, num_groups = 0;
    std::vector<std::vector<unsigned char>> tmaps;  // per ring slot
  };
  std::vector<ExtraModule> extra;

  // constant-run replicator
  int repl_chunks = 0;                 // 0: no constant runs registered
  int repl_mode = 0;                   // 0: coalesced stores, 1: TMA tile stores
  int repl_runs = 0;
  int repl_nodes_per_block = 32;
  int* d_run_col0 = nullptr;
  int* d_run_len = nullptr;
  int* d_run_off = nullptr;
  int repl_rows = 16;                  // node rows per TMA store
  int repl_tiles_per_block = 4;
  size_t repl_smem = 0;
  std::vector<int32_t> repl_col0, repl_w, repl_off;   // per chunk
  double* d_repl_lit = nullptr;
  int* d_repl_inv = nullptr;
  int* d_repl_off = nullptr;
  int* d_repl_w = nullptr;
  std::vector<OptyReplMaps> repl_maps;  // per ring slot
  cudaStream_t repl_stream = nullptr;
  cudaEvent_t ev_repl_go = nullptr, ev_repl_done = nullptr;

  std::vector<int32_t> d2h_begin, d2h_end;
  unsigned smem_bytes = 0;
  unsigned grid_x = 0;
  int64_t launches = 0;
  float last_ms = 0.f;
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

int build_tmaps_into(opty_colloc* h, int slot, int seg_first, int seg_count, std::vector<unsigned char>& blob);

int build_tmaps(opty_colloc* h, int slot) {
  return build_tmaps_into(h, slot, 0, h->cfg.primary_segments, h->tmaps[slot]);
}

int build_tmaps_into(opty_colloc* h, int slot, int seg_first, int seg_count, std::vector<unsigned char>& blob) {
  const opty_colloc_cfg& c = h->cfg;
  blob.assign(sizeof(CUtensorMap) * (1 + (seg_count > 0 ? seg_count : 1)), 0);
  CUtensorMap* maps = reinterpret_cast<CUtensorMap*>(blob.data());
  int rc;
  if (c.tma_load == 1) {
    const uint32_t threads = c.persistent ? 32u : 32u * c.warps_per_block;  // persistent: per-warp slices
    const uint32_t xbox = (threads <= 128u ? threads : 128u) + 2u;
    if ((rc = encode_2d(&maps[0], h->d_traj, (uint64_t)h->ncols, (uint64_t)h->RD, (uint64_t)h->ldt * 8, xbox,
                        (uint32_t)h->RD)))
      return rc;
  }
  if (c.tma_store) {
    for (int g = 0; g < seg_count; ++g) {
      if ((rc = encode_2d(&maps[1 + g], h->d_jac[slot] + c.seg_col0[seg_first + g],
                          (uint64_t)c.seg_ncols[seg_first + g], (uint64_t)h->nn, (uint64_t)h->K * 8,
                          (uint32_t)c.tile_cols, c.persistent == 1 ? 32u * c.warps_per_block : 32u)))
        return rc;
    }
  }
  return OPTY_OK;
}

int launch_eval(opty_colloc* h, bool record_events = false) {
  const opty_colloc_cfg& c = h->cfg;
  if (!h->known_set) return fail(OPTY_ERR_STATE, "opty_colloc_set_known must be called before evaluating");
  if (!h->free_valid) return fail(OPTY_ERR_STATE, "no free vector resident on the device");
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
  if (h->inv_dirty && c.num_inv > 0) {
    void* args[2] = {&h->d_uni, &h->d_inv};
    DRV_CHECK(g_drv.LaunchKernel(h->f_inv, 1, 1, 1, 32, 1, 1, 0, (CUstream)h->stream, args, nullptr));
    h->launches++;
    RT_CHECK(cudaMemcpyAsync(reinterpret_cast<void*>(h->ci_sym), h->d_inv, (size_t)c.num_inv * 8,
                             cudaMemcpyDeviceToDevice, h->stream));
    for (auto& em : h->extra)
      RT_CHECK(cudaMemcpyAsync(reinterpret_cast<void*>(em.ci_sym), h->d_inv, (size_t)c.num_inv * 8,
                               cudaMemcpyDeviceToDevice, h->stream));
  }
  {
    int covered = c.primary_segments;
    for (auto& em : h->extra) covered += em.seg_count;
    if (covered != c.num_segments)
      return fail(OPTY_ERR_STATE, "not all modules of this problem have been added (opty_colloc_add_module)");
  }
  h->inv_dirty = false;
  h->ring = (h->ring + 1) % c.out_ring;
  h->eval_seq++;
  if (c.const_image_doubles > 0) {
    if (h->repl_chunks == 0)
      return fail(OPTY_ERR_STATE, "opty_colloc_set_const_runs must be called before evaluating");
    // the replicator only depends on the invariants table: it runs on its own stream, next to the
    // pre-pass and the main kernel, and is joined below
    RT_CHECK(cudaEventRecord(h->ev_repl_go, h->stream));
    RT_CHECK(cudaStreamWaitEvent(h->repl_stream, h->ev_repl_go, 0));
    if (h->repl_mode == 1) {
      const int n_tiles = (h->nn + h->repl_rows - 1) / h->repl_rows;
      dim3 grid((unsigned)((n_tiles + h->repl_tiles_per_block - 1) / h->repl_tiles_per_block),
                (unsigned)h->repl_chunks);
      opty_replicate_kernel<<<grid, OPTY_REPL_THREADS, h->repl_smem, h->repl_stream>>>(
          h->repl_maps[h->ring], h->d_repl_lit, h->d_repl_inv, h->d_inv, h->d_repl_off, h->d_repl_w, h->nn,
          h->repl_rows, h->repl_tiles_per_block);
    } else {
      const unsigned blocks = (unsigned)((h->nn + h->repl_nodes_per_block - 1) / h->repl_nodes_per_block);
      opty_replicate_st_kernel<<<blocks, OPTY_REPL_THREADS, (size_t)c.const_image_doubles * 8 + (size_t)h->repl_runs * 12, h->repl_stream>>>(
          h->d_jac[h->ring], (long long)h->K, h->d_repl_lit, h->d_repl_inv, h->d_inv, h->d_run_col0, h->d_run_len,
          h->d_run_off, h->repl_runs, c.const_image_doubles, h->nn, h->repl_nodes_per_block);
    }
    RT_CHECK(cudaGetLastError());
    RT_CHECK(cudaEventRecord(h->ev_repl_done, h->repl_stream));
    h->launches++;
  }
  OptyParams p;
  p.traj = h->d_traj;
  p.con = h->d_con[h->ring];
  p.jac = h->d_jac[h->ring];
  p.ldt = h->ldt;
  p.ldc = h->nn;
  p.n_nodes = h->nn;
  p.n_cols = h->ncols;
  if (c.persistent) {
    // one cooperative launch: phase 0 = pre-pass, grid barrier, then every block walks its schedule entry
    if (h->sched_blocks < 1) return fail(OPTY_ERR_STATE, "opty_colloc_set_schedule must be called before evaluating");
    OptyPersistArgs ps;
    ps.sched = h->d_sched;
    ps.barrier = h->d_barrier;
    h->barrier_epoch += (unsigned)h->sched_blocks;
    ps.barrier_target = h->barrier_epoch;
    ps.block_clocks = h->d_block_clocks;
    ps.pre_units = c.num_derived > 0 ? c.pre_groups : 0;
    void* pargs[3] = {h->tmaps[h->ring].data(), &p, &ps};
    DRV_CHECK(g_drv.LaunchCooperativeKernel(h->f_eval, (unsigned)h->sched_blocks, 1, 1, 32u * c.warps_per_block, 1, 1,
                                            h->smem_bytes, (CUstream)h->stream, pargs));
    h->launches++;
  } else {
  if (c.num_derived > 0) {
    void* pargs[1] = {&p};
    DRV_CHECK(g_drv.LaunchKernel(h->f_pre, (unsigned)((h->nn + 127) / 128), (unsigned)c.pre_groups, 1, 128, 1, 1, 0,
                                 (CUstream)h->stream, pargs, nullptr));
    h->launches++;
  }
  void* args[2] = {h->tmaps[h->ring].data(), &p};
  DRV_CHECK(g_drv.LaunchKernel(h->f_eval, h->grid_x, (unsigned)c.num_groups, 1, 32u * c.warps_per_block, 1, 1, h->smem_bytes,
                               (CUstream)h->stream, args, nullptr));
  h->launches++;
  }
  for (auto& em : h->extra) {
    void* eargs[2] = {em.tmaps[h->ring].data(), &p};
    DRV_CHECK(g_drv.LaunchKernel(em.f_eval, h->grid_x, (unsigned)em.num_groups, 1, 32u * c.warps_per_block, 1, 1,
                                 h->smem_bytes, (CUstream)h->stream, eargs, nullptr));
    h->launches++;
  }
  if (c.const_image_doubles > 0) RT_CHECK(cudaStreamWaitEvent(h->stream, h->ev_repl_done, 0));
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

int enqueue_jac_copy(opty_colloc* h, cudaStream_t st) {
  int rc0 = ensure_host_jac(h);
  if (rc0) return rc0;
  const size_t bytes = (size_t)h->nn * h->K * 8;
  const int which = h->h_jacs[1] ? (h->jac_cur ^ 1) : 0;
  double* dst = h->h_jacs[which];
  h->jac_target = which;
  if (h->d2h_begin.empty() || h->full_fetch[which]) {
    // first fetch into this buffer (and every fetch without column ranges):
    // the whole block, which also brings the constant columns to the host once
    RT_CHECK(cudaMemcpyAsync(dst, h->d_jac[h->ring], bytes, cudaMemcpyDeviceToHost, st));
    h->full_fetch[which] = false;
  } else {
    for (size_t i = 0; i < h->d2h_begin.size(); ++i) {
      const int b = h->d2h_begin[i], e = h->d2h_end[i];
      RT_CHECK(cudaMemcpy2DAsync(dst + b, (size_t)h->K * 8, h->d_jac[h->ring] + b, (size_t)h->K * 8,
                                 (size_t)(e - b) * 8, h->nn, cudaMemcpyDeviceToHost, st));
    }
  }
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
  if (free_host != h->h_free) {
    for (int r = 0; r < nrows; ++r) {
      const size_t off = (size_t)r * c.N + c.node_lo;
      memcpy(h->h_free + off, free_host + off, win_bytes);
    }
    if (tail_bytes) memcpy(h->h_free + tail_off, free_host + tail_off, tail_bytes);
  }
  const int rows = c.n + c.q;
  // rows of the free vector are [row][N]; this handle keeps columns node_lo..node_hi
  RT_CHECK(cudaMemcpy2DAsync(h->d_traj, (size_t)h->ldt * 8, h->h_free + c.node_lo, (size_t)c.N * 8,
                             (size_t)h->ncols * 8, rows, cudaMemcpyHostToDevice, h->stream));
  if (c.r + c.s > 0) {
    RT_CHECK(cudaMemcpyAsync(h->d_uni + c.pk, h->h_free + (size_t)rows * c.N, (size_t)(c.r + c.s) * 8,
                             cudaMemcpyHostToDevice, h->stream));
    h->inv_dirty = true;
  }
  h->free_valid = true;
  h->evaluated = false;
  return OPTY_OK;
}

}  // namespace

extern "C" {

const char* opty_colloc_last_error(void) { return g_err.c_str(); }

int opty_b200_abi_version(void) { return OPTY_B200_ABI_VERSION; }

int opty_colloc_create(const opty_colloc_cfg* cfg, const void* cubin, size_t cubin_bytes, opty_colloc_t** out) {
  if (!cfg || !cubin || !out || cubin_bytes == 0) return fail(OPTY_ERR_ARG, "null argument");
  *out = nullptr;
  if (cfg->abi_version != OPTY_B200_ABI_VERSION) return fail(OPTY_ERR_ARG, "ABI version mismatch");
  if (cfg->N < (cfg->method == OPTY_ELEMENTWISE ? 1 : 2) || cfg->n < (cfg->method == OPTY_ELEMENTWISE ? 0 : 1) || cfg->M < 1 || cfg->P < 1 || cfg->q < 0 || cfg->k < 0 || cfg->r < 0 ||
      cfg->pk < 0 || (cfg->s != 0 && cfg->s != 1))
    return fail(OPTY_ERR_ARG, "invalid problem dimensions");
  const bool elementwise = cfg->method == OPTY_ELEMENTWISE;
  if (cfg->node_lo < 0 || cfg->node_hi > cfg->N - (elementwise ? 0 : 1) || cfg->node_lo >= cfg->node_hi)
    return fail(OPTY_ERR_ARG, "invalid node range");
  if (cfg->num_groups < 1 || cfg->num_groups > OPTY_MAX_GROUPS) return fail(OPTY_ERR_ARG, "invalid group count");
  if (cfg->primary_segments > 240)
    return fail(OPTY_ERR_ARG, "a module's TMA descriptors are one kernel parameter: at most 240 store segments per module");
  if (cfg->warps_per_block < 1 || cfg->warps_per_block > 32 ||
      (cfg->warps_per_block > 4 && cfg->warps_per_block % 4 != 0 && !cfg->persistent) || cfg->tile_cols < 2 || (cfg->tile_cols & 1) ||
      cfg->tile_cols > 256)
    return fail(OPTY_ERR_ARG, "invalid kernel geometry");
  if (cfg->out_ring < 1 || cfg->out_ring > 64) return fail(OPTY_ERR_ARG, "invalid out_ring");
  if (cfg->num_derived < 0 || cfg->pre_groups < 0 || (cfg->num_derived > 0 && cfg->pre_groups < 1) ||
      cfg->tile_bufs < 1 || cfg->tile_bufs > 4)
    return fail(OPTY_ERR_ARG, "invalid num_derived / pre_groups / tile_bufs");
  const int expectP = (cfg->method == OPTY_MIDPOINT ? 2 * cfg->n + 2 * cfg->q : 2 * cfg->n + cfg->q) + cfg->r + cfg->s;
  if (cfg->method != OPTY_MIDPOINT && cfg->method != OPTY_BACKWARD_EULER && !elementwise)
    return fail(OPTY_ERR_ARG, "invalid method");
  if (!elementwise && cfg->P != expectP)
    return fail(OPTY_ERR_ARG, "P does not match n, q, r, s and the integration method");
  {
    if (cfg->num_segments < 0 || cfg->num_segments > OPTY_MAX_SEGMENTS || cfg->const_image_doubles < 0 ||
        cfg->primary_segments < 0 || cfg->primary_segments > cfg->num_segments)
      return fail(OPTY_ERR_ARG, "invalid segment count");
    long long covered = cfg->const_image_doubles, prev_end = 0;
    for (int g = 0; g < cfg->num_segments; ++g) {
      if (cfg->seg_col0[g] < prev_end || cfg->seg_ncols[g] < 1)
        return fail(OPTY_ERR_ARG, "store segments must be sorted, non-empty and disjoint");
      prev_end = (long long)cfg->seg_col0[g] + cfg->seg_ncols[g];
      covered += cfg->seg_ncols[g];
    }
    if (prev_end > (long long)cfg->M * cfg->P || covered != (long long)cfg->M * cfg->P)
      return fail(OPTY_ERR_ARG, "store segments and constant runs must tile the M*P columns");
    if (cfg->const_image_doubles > 0 && !cfg->tma_store)
      return fail(OPTY_ERR_ARG, "constant runs need TMA stores (even M*P)");
  }

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
  h->nn = cfg->node_hi - cfg->node_lo;
  h->ncols = h->nn + (elementwise ? 0 : 1);
  h->R = cfg->n + cfg->q + cfg->k;
  h->RD = h->R + cfg->num_derived;
  h->K = cfg->M * cfg->P;
  h->ldt = round_up(h->ncols, 16);
  h->free_len = (size_t)(cfg->n + cfg->q) * cfg->N + cfg->r + cfg->s;
  if (cfg->tma_load == 1 && h->RD > 256) {
    delete h;
    return fail(OPTY_ERR_ARG, "TMA input staging supports at most 256 trajectory rows");
  }
  if (cfg->tma_store && ((h->K & 1) != 0)) {
    delete h;
    return fail(OPTY_ERR_ARG, "TMA Jacobian stores need an even M*P");
  }

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

  CREATE_DRV(g_drv.ModuleLoadData(&h->mod, cubin));
  CREATE_DRV(g_drv.ModuleGetFunction(&h->f_eval, h->mod, "opty_colloc_eval"));
  CREATE_DRV(g_drv.ModuleGetFunction(&h->f_inv, h->mod, "opty_colloc_inv"));
  CREATE_DRV(g_drv.ModuleGetFunction(&h->f_pre, h->mod, "opty_colloc_pre"));
  CREATE_RT(cudaDeviceGetAttribute(&h->num_sms, cudaDevAttrMultiProcessorCount, cfg->device));
  CREATE_DRV(g_drv.ModuleGetGlobal(&h->ci_sym, &h->ci_bytes, h->mod, "opty_ci"));
  if (h->ci_bytes < (size_t)cfg->num_inv * 8) {
    opty_colloc_destroy(h);
    return fail(OPTY_ERR_ARG, "module's invariant table is smaller than cfg.num_inv");
  }

  CREATE_RT(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  CREATE_RT(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
  CREATE_RT(cudaEventCreate(&h->ev0));
  CREATE_RT(cudaEventCreate(&h->ev1));
  CREATE_RT(cudaEventCreateWithFlags(&h->ev_copy, cudaEventDisableTiming));
  CREATE_RT(cudaEventCreateWithFlags(&h->ev_con, cudaEventDisableTiming));
  CREATE_RT(cudaStreamCreateWithFlags(&h->repl_stream, cudaStreamNonBlocking));
  CREATE_RT(cudaEventCreateWithFlags(&h->ev_repl_go, cudaEventDisableTiming));
  CREATE_RT(cudaEventCreateWithFlags(&h->ev_repl_done, cudaEventDisableTiming));

  CREATE_RT(cudaMalloc(&h->d_traj, (size_t)h->RD * h->ldt * 8));
  CREATE_RT(cudaMemsetAsync(h->d_traj, 0, (size_t)h->RD * h->ldt * 8, h->stream));
  const int nuni = cfg->pk + cfg->r + 1;
  CREATE_RT(cudaMalloc(&h->d_uni, (size_t)nuni * 8));
  CREATE_RT(cudaMemsetAsync(h->d_uni, 0, (size_t)nuni * 8, h->stream));
  if (cfg->s == 0) {
    CREATE_RT(cudaMemcpyAsync(h->d_uni + cfg->pk + cfg->r, &cfg->h, 8, cudaMemcpyHostToDevice, h->stream));
    CREATE_RT(cudaStreamSynchronize(h->stream));
  }
  CREATE_RT(cudaMalloc(&h->d_inv, (size_t)(cfg->num_inv > 0 ? cfg->num_inv : 1) * 8));
  h->d_con.assign(cfg->out_ring, nullptr);
  h->d_jac.assign(cfg->out_ring, nullptr);
  h->tmaps.resize(cfg->out_ring);
  for (int s = 0; s < cfg->out_ring; ++s) {
    CREATE_RT(cudaMalloc(&h->d_con[s], (size_t)cfg->M * h->nn * 8));
    CREATE_RT(cudaMalloc(&h->d_jac[s], (size_t)h->nn * h->K * 8));
    CREATE_CHECK(build_tmaps(h, s));
  }

  CREATE_RT(cudaHostAlloc(&h->h_free, h->free_len * 8, cudaHostAllocDefault));
  CREATE_RT(cudaHostAlloc(&h->h_con, ((size_t)cfg->M * h->nn + cfg->con_tail) * 8, cudaHostAllocDefault));
  // the pinned Jacobian buffers (8.4 GB each at BASELINE config 5) are allocated on first use:
  // device-resident consumers never need them

  const unsigned tiles_bytes = (unsigned)cfg->warps_per_block * (unsigned)cfg->tile_bufs * 32u * cfg->tile_cols * 8u;
  const unsigned threads = 32u * cfg->warps_per_block;
  const unsigned xseg = threads <= 128u ? threads : 128u;
  const unsigned nseg = threads / xseg;
  const unsigned xin_bytes =
      cfg->tma_load == 2 ? 0u : nseg * (unsigned)round_up((int64_t)h->RD * (xseg + 2u) * 8, 128);
  h->smem_bytes = tiles_bytes + xin_bytes + 128u;
  if (cfg->persistent) {
    if (cfg->warps_per_block > 8)
      return (opty_colloc_destroy(h), fail(OPTY_ERR_ARG, "the persistent kernel supports at most 8 warps per block"));
    if (cfg->tma_load != 1 || !cfg->tma_store || elementwise)
      return (opty_colloc_destroy(h), fail(OPTY_ERR_ARG, "the persistent kernel needs TMA input staging and TMA stores"));
    // per warp: its staging tiles and its own [R+D][34] input slice, one mbarrier each
    const unsigned slice = (unsigned)round_up((int64_t)h->RD * 34 * 8, 128);
    h->smem_bytes = tiles_bytes + (unsigned)cfg->warps_per_block * slice + 8u * cfg->warps_per_block + 128u;
    CREATE_RT(cudaMalloc(&h->d_barrier, sizeof(unsigned int)));
    CREATE_RT(cudaMemset(h->d_barrier, 0, sizeof(unsigned int)));
  }
  if (const char* pad = getenv("OPTY_B200_DEBUG_SMEM_FLOOR")) {
    // measurement aid: a larger dynamic shared-memory request caps the resident blocks per SM
    const unsigned floor_bytes = (unsigned)atoi(pad);
    if (floor_bytes > h->smem_bytes) h->smem_bytes = floor_bytes;
  }
  if (h->smem_bytes > 227u * 1024u) {
    opty_colloc_destroy(h);
    return fail(OPTY_ERR_ARG, "kernel needs more than 227 KB of shared memory per block");
  }
  CREATE_DRV(g_drv.FuncSetAttribute(h->f_eval, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)h->smem_bytes));

  h->n_tiles = (h->nn + 32 * cfg->warps_per_block - 1) / (32 * cfg->warps_per_block);
  h->grid_x = (unsigned)h->n_tiles;
  CREATE_RT(cudaStreamSynchronize(h->stream));
  *out = h;
  return OPTY_OK;
}

int opty_colloc_destroy(opty_colloc_t* h) {
  if (!h) return OPTY_OK;
  cudaSetDevice(h->cfg.device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  if (h->copy_stream) cudaStreamSynchronize(h->copy_stream);
  if (h->repl_stream) cudaStreamSynchronize(h->repl_stream);
  cudaFree(h->d_sched);
  cudaFree(h->d_barrier);
  cudaFree(h->d_block_clocks);
  cudaFree(h->d_repl_lit);
  cudaFree(h->d_repl_inv);
  cudaFree(h->d_repl_off);
  cudaFree(h->d_repl_w);
  cudaFree(h->d_run_col0);
  cudaFree(h->d_run_len);
  cudaFree(h->d_run_off);
  if (h->ev_repl_go) cudaEventDestroy(h->ev_repl_go);
  if (h->ev_repl_done) cudaEventDestroy(h->ev_repl_done);
  if (h->repl_stream) cudaStreamDestroy(h->repl_stream);
  for (double* p : h->d_con) cudaFree(p);
  for (double* p : h->d_jac) cudaFree(p);
  cudaFree(h->d_traj);
  cudaFree(h->d_uni);
  cudaFree(h->d_inv);
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
  if (h->mod && g_drv.ModuleUnload) g_drv.ModuleUnload(h->mod);
  for (auto& em : h->extra)
    if (em.mod && g_drv.ModuleUnload) g_drv.ModuleUnload(em.mod);
  delete h;
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
  RT_CHECK(cudaSetDevice(h->cfg.device));
  int rc = upload(h, free_host, nullptr);
  if (rc) return rc;
  bool launched = false;
  if (!h->evaluated) {
    if ((rc = launch_eval(h))) return rc;
    launched = true;
  }
  const size_t bytes = (size_t)h->cfg.M * h->nn * 8;
  if (!h->con_fetched) {
    RT_CHECK(cudaMemcpyAsync(h->h_con, h->d_con[h->ring], bytes, cudaMemcpyDeviceToHost, h->stream));
  }
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

int opty_colloc_set_d2h_columns(opty_colloc_t* h, int num_ranges, const int32_t* col_begin, const int32_t* col_end,
                                const double* fill) {
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
  if (fill) {
    int rcf = ensure_host_jac(h);
    if (rcf) return rcf;
    // pre-write the per-node constant pattern once
    for (int w = 0; w < 2; ++w)
      if (h->h_jacs[w])
        for (int64_t i = 0; i < h->nn; ++i) memcpy(h->h_jacs[w] + i * h->K, fill, (size_t)h->K * 8);
  }
  h->d2h_begin.swap(b);
  h->d2h_end.swap(e);
  h->jac_fetched = false;
  if (!fill) h->full_fetch[0] = h->full_fetch[1] = true;
  return OPTY_OK;
}

int opty_colloc_add_module(opty_colloc_t* h, const void* cubin, size_t cubin_bytes, int seg_first, int seg_count,
                           int num_groups) {
  if (!h || !cubin || cubin_bytes == 0) return fail(OPTY_ERR_ARG, "null argument");
  const opty_colloc_cfg& c = h->cfg;
  int expect_first = c.primary_segments;
  for (auto& em : h->extra) expect_first += em.seg_count;
  if (seg_first != expect_first || seg_count < 0 || seg_first + seg_count > c.num_segments || num_groups < 1 ||
      num_groups > OPTY_MAX_GROUPS || seg_count > 240)
    return fail(OPTY_ERR_ARG, "modules must be added in segment order, stay inside cfg.num_segments and hold at most 240 segments");
  RT_CHECK(cudaSetDevice(c.device));
  opty_colloc::ExtraModule em;
  em.seg_first = seg_first;
  em.seg_count = seg_count;
  em.num_groups = num_groups;
  DRV_CHECK(g_drv.ModuleLoadData(&em.mod, cubin));
  CUresult r1 = g_drv.ModuleGetFunction(&em.f_eval, em.mod, "opty_colloc_eval");
  size_t ci_bytes = 0;
  CUresult r2 = r1 == CUDA_SUCCESS ? g_drv.ModuleGetGlobal(&em.ci_sym, &ci_bytes, em.mod, "opty_ci") : r1;
  CUresult r3 = r2 == CUDA_SUCCESS ? g_drv.FuncSetAttribute(em.f_eval, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES,
                                                            (int)h->smem_bytes)
                                   : r2;
  if (r3 != CUDA_SUCCESS || ci_bytes < (size_t)c.num_inv * 8) {
    g_drv.ModuleUnload(em.mod);
    return fail(r3 != CUDA_SUCCESS ? OPTY_ERR_CUDA : OPTY_ERR_ARG,
                r3 != CUDA_SUCCESS ? "opty_colloc_add_module: " + drv_err(r3)
                                   : std::string("module's invariant table is smaller than cfg.num_inv"));
  }
  em.tmaps.resize(c.out_ring);
  for (int s = 0; s < c.out_ring; ++s) {
    int rc = build_tmaps_into(h, s, seg_first, seg_count, em.tmaps[s]);
    if (rc) {
      g_drv.ModuleUnload(em.mod);
      return rc;
    }
  }
  h->extra.push_back(std::move(em));
  h->inv_dirty = true;
  h->evaluated = false;
  return OPTY_OK;
}

int opty_colloc_set_schedule(opty_colloc_t* h, int num_blocks, const int32_t* triples) {
  if (!h || !triples || num_blocks < 1) return fail(OPTY_ERR_ARG, "invalid argument");
  const opty_colloc_cfg& c = h->cfg;
  if (!c.persistent) return fail(OPTY_ERR_ARG, "the module was not emitted with the persistent kernel");
  if (num_blocks > h->num_sms) return fail(OPTY_ERR_ARG, "at most one block per SM");
  const int n_tiles = (h->nn + 31) / 32;
  // every (group, tile) pair exactly once
  std::vector<int> covered((size_t)c.num_groups * n_tiles, 0);
  std::vector<int32_t> table((size_t)num_blocks * 4, 0);
  for (int b = 0; b < num_blocks; ++b) {
    const int g = triples[3 * b], t0 = triples[3 * b + 1], t1 = triples[3 * b + 2];
    if (g < 0 || g >= c.num_groups || t0 < 0 || t1 < t0 || t1 > n_tiles)
      return fail(OPTY_ERR_ARG, "schedule entry out of range");
    for (int t = t0; t < t1; ++t) covered[(size_t)g * n_tiles + t]++;
    table[4 * b] = g;
    table[4 * b + 1] = t0;
    table[4 * b + 2] = t1;
  }
  for (int v : covered)
    if (v != 1) return fail(OPTY_ERR_ARG, "the schedule must cover every (group, tile) pair exactly once");
  RT_CHECK(cudaSetDevice(c.device));
  RT_CHECK(cudaStreamSynchronize(h->stream));
  cudaFree(h->d_sched);
  cudaFree(h->d_block_clocks);
  h->d_sched = nullptr;
  h->d_block_clocks = nullptr;
  RT_CHECK(cudaMalloc(&h->d_sched, (size_t)num_blocks * 16));
  RT_CHECK(cudaMalloc(&h->d_block_clocks, (size_t)num_blocks * 8));
  RT_CHECK(cudaMemcpy(h->d_sched, table.data(), (size_t)num_blocks * 16, cudaMemcpyHostToDevice));
  RT_CHECK(cudaMemset(h->d_block_clocks, 0, (size_t)num_blocks * 8));
  // the barrier counter restarts with the new grid size
  RT_CHECK(cudaMemset(h->d_barrier, 0, sizeof(unsigned int)));
  h->barrier_epoch = 0;
  h->sched_blocks = num_blocks;
  h->evaluated = false;
  return OPTY_OK;
}

int opty_colloc_block_clocks(opty_colloc_t* h, int num_blocks, int64_t* clocks) {
  if (!h || !clocks || num_blocks != h->sched_blocks || !h->d_block_clocks)
    return fail(OPTY_ERR_ARG, "invalid argument");
  RT_CHECK(cudaSetDevice(h->cfg.device));
  RT_CHECK(cudaStreamSynchronize(h->stream));
  RT_CHECK(cudaMemcpy(clocks, h->d_block_clocks, (size_t)num_blocks * 8, cudaMemcpyDeviceToHost));
  return OPTY_OK;
}

int opty_colloc_set_const_runs(opty_colloc_t* h, int num_runs, const int32_t* col0, const int32_t* len,
                               const double* lit, const int32_t* inv_idx) {
  if (!h || num_runs < 1 || !col0 || !len || !lit || !inv_idx) return fail(OPTY_ERR_ARG, "invalid argument");
  const opty_colloc_cfg& c = h->cfg;
  if (!c.tma_store) return fail(OPTY_ERR_ARG, "constant runs need TMA stores (even M*P)");
  RT_CHECK(cudaSetDevice(c.device));
  long long total = 0;
  int prev_end = 0;
  std::vector<int32_t> ch_col0, ch_w, ch_off;
  for (int r = 0; r < num_runs; ++r) {
    if (col0[r] < prev_end || len[r] < 2 || (col0[r] & 1) || (len[r] & 1) || col0[r] + len[r] > h->K)
      return fail(OPTY_ERR_ARG, "constant runs must be sorted, disjoint, inside [0, M*P), with even start and length");
    // chunks of at most 254 columns (TMA boxes are at most 256 elements wide, rows 16-byte multiples)
    for (int done = 0; done < len[r];) {
      const int w = (len[r] - done) < 254 ? (len[r] - done) : 254;
      ch_col0.push_back(col0[r] + done);
      ch_w.push_back(w);
      ch_off.push_back((int32_t)(total + done));
      done += w;
    }
    total += len[r];
    prev_end = col0[r] + len[r];
  }
  if (total != c.const_image_doubles) return fail(OPTY_ERR_ARG, "constant runs do not match cfg.const_image_doubles");
  if ((int)ch_w.size() > OPTY_REPL_MAX_CHUNKS) return fail(OPTY_ERR_ARG, "too many constant-run chunks");
  for (long long i = 0; i < total; ++i)
    if (inv_idx[i] >= c.num_inv) return fail(OPTY_ERR_ARG, "invariant index out of range");
  if (const char* e = getenv("OPTY_B200_REPL_MODE")) h->repl_mode = atoi(e);
  if (const char* e = getenv("OPTY_B200_REPL_NODES")) h->repl_nodes_per_block = atoi(e);
  if (h->repl_nodes_per_block < 1) return fail(OPTY_ERR_ARG, "invalid replicator geometry");
  if ((size_t)c.const_image_doubles * 8 > 200u * 1024u)
    return fail(OPTY_ERR_ARG, "constant-run image exceeds 200 KB of shared memory");
  RT_CHECK(cudaFuncSetAttribute(opty_replicate_st_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  {
    std::vector<int32_t> offs(num_runs);
    int32_t acc = 0;
    for (int r = 0; r < num_runs; ++r) {
      offs[r] = acc;
      acc += len[r];
    }
    cudaFree(h->d_run_col0);
    cudaFree(h->d_run_len);
    cudaFree(h->d_run_off);
    h->d_run_col0 = h->d_run_len = h->d_run_off = nullptr;
    RT_CHECK(cudaMalloc(&h->d_run_col0, (size_t)num_runs * 4));
    RT_CHECK(cudaMalloc(&h->d_run_len, (size_t)num_runs * 4));
    RT_CHECK(cudaMalloc(&h->d_run_off, (size_t)num_runs * 4));
    RT_CHECK(cudaMemcpy(h->d_run_col0, col0, (size_t)num_runs * 4, cudaMemcpyHostToDevice));
    RT_CHECK(cudaMemcpy(h->d_run_len, len, (size_t)num_runs * 4, cudaMemcpyHostToDevice));
    RT_CHECK(cudaMemcpy(h->d_run_off, offs.data(), (size_t)num_runs * 4, cudaMemcpyHostToDevice));
    h->repl_runs = num_runs;
  }
  if (const char* e = getenv("OPTY_B200_REPL_ROWS")) h->repl_rows = atoi(e);
  if (const char* e = getenv("OPTY_B200_REPL_TILES")) h->repl_tiles_per_block = atoi(e);
  if (h->repl_rows < 1 || h->repl_rows > 256 || h->repl_tiles_per_block < 1)
    return fail(OPTY_ERR_ARG, "invalid replicator geometry");
  int wmax = 0;
  for (int w : ch_w) wmax = w > wmax ? w : wmax;
  h->repl_smem = (size_t)h->repl_rows * wmax * 8;
  if (h->repl_smem > 200u * 1024u) return fail(OPTY_ERR_ARG, "replicator image exceeds 200 KB of shared memory");
  RT_CHECK(cudaFuncSetAttribute(opty_replicate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  cudaFree(h->d_repl_lit);
  cudaFree(h->d_repl_inv);
  cudaFree(h->d_repl_off);
  cudaFree(h->d_repl_w);
  h->d_repl_lit = nullptr;
  h->d_repl_inv = h->d_repl_off = h->d_repl_w = nullptr;
  const int nch = (int)ch_w.size();
  RT_CHECK(cudaMalloc(&h->d_repl_lit, (size_t)total * 8));
  RT_CHECK(cudaMalloc(&h->d_repl_inv, (size_t)total * 4));
  RT_CHECK(cudaMalloc(&h->d_repl_off, (size_t)nch * 4));
  RT_CHECK(cudaMalloc(&h->d_repl_w, (size_t)nch * 4));
  RT_CHECK(cudaMemcpy(h->d_repl_lit, lit, (size_t)total * 8, cudaMemcpyHostToDevice));
  RT_CHECK(cudaMemcpy(h->d_repl_inv, inv_idx, (size_t)total * 4, cudaMemcpyHostToDevice));
  RT_CHECK(cudaMemcpy(h->d_repl_off, ch_off.data(), (size_t)nch * 4, cudaMemcpyHostToDevice));
  RT_CHECK(cudaMemcpy(h->d_repl_w, ch_w.data(), (size_t)nch * 4, cudaMemcpyHostToDevice));
  h->repl_maps.assign(c.out_ring, OptyReplMaps());
  for (int s = 0; s < c.out_ring; ++s) {
    memset(&h->repl_maps[s], 0, sizeof(OptyReplMaps));
    for (int k = 0; k < nch; ++k) {
      int rc = encode_2d(&h->repl_maps[s].m[k], h->d_jac[s] + ch_col0[k], (uint64_t)ch_w[k], (uint64_t)h->nn,
                         (uint64_t)h->K * 8, (uint32_t)ch_w[k], (uint32_t)h->repl_rows);
      if (rc) return rc;
    }
  }
  h->repl_col0 = ch_col0;
  h->repl_w = ch_w;
  h->repl_off = ch_off;
  h->repl_chunks = nch;
  h->evaluated = false;
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
