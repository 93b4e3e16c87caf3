// Hand-written sm_100a skeleton of the collocation constraint + Jacobian kernel.
//
// The generated module (opty_b200/codegen.py) defines the problem sizes
// (OPTY_M, OPTY_P, OPTY_K, OPTY_R, OPTY_C, ...) and one straight-line
// `opty_group_<g>` device function per output group, then includes this file,
// which supplies everything around the arithmetic:
//
//   * staging of the block's slice of the trajectory matrix into shared memory
//     with ONE 2-D TMA tile load (cp.async.bulk.tensor, mbarrier completion),
//   * the per-warp, double-buffered, bank-conflict-free staging tile for the
//     node-major Jacobian block and its drain by 2-D TMA tile stores,
//   * a coalesced warp-per-node fallback for shapes TMA cannot describe.
//
// Work mapping: lane = collocation node (so that all lanes of a warp execute
// the same generated instruction stream), warp = 32 consecutive nodes x one
// output group, grid = (node tiles, groups).  It replaces the node loop of the
// reference's generated Cython (`for i in prange(n)`, opty/utils.py:524-526)
// and the per-node `eval_matrix` C function (opty/utils.py:483-494).
//
// Data layout (all float64):
//   traj : [R][ldt]   rows = states, unknown inputs, known inputs; col = node
//   con  : [M][ldc]   eom-major residuals, node i at column i
//                     (layout of opty/direct_collocation.py:2446 without the
//                     transpose copy)
//   jac  : [nodes][K] node-major partials, K = M*P, incl. structural zeros
//                     (layout of opty/direct_collocation.py:2814, 2887)
#pragma once

#include <cuda.h>
#include <stdint.h>

#include "colloc_params.h"

#define OPTY_THREADS (OPTY_WARPS * 32)
// The block's slice of the trajectory matrix is staged in segments of
// OPTY_XSEG nodes; a segment holds one column per node plus the right
// neighbour, rounded up to an even count (TMA rows are 16-byte multiples, TMA
// boxes at most 256 elements wide).  Blocks wider than 128 threads use several
// overlapping segments.
#if OPTY_THREADS <= 128
#define OPTY_XSEG OPTY_THREADS
#else
#define OPTY_XSEG 128
#endif
#define OPTY_XBOX (OPTY_XSEG + 2)
#define OPTY_NSEG (OPTY_THREADS / OPTY_XSEG)
#define OPTY_XSEG_BYTES (((OPTY_R * OPTY_XBOX * 8) + 127) / 128 * 128)
#define OPTY_TILE_DOUBLES (32 * OPTY_C)
#ifndef OPTY_BLOCK_SYNC
#define OPTY_BLOCK_SYNC 0
#endif
#ifndef OPTY_DEBUG_NOSTORE
#define OPTY_DEBUG_NOSTORE 0   // measurement aid: skip the Jacobian tile stores
#endif

struct OptyTmaps {
  CUtensorMap in;                  // traj as {cols, R}
  CUtensorMap out[OPTY_NGROUPS];   // group g's columns of jac as {ncols_g, nodes}
};

// node-invariant sub-expressions, filled by the host from opty_colloc_inv
__constant__ double opty_ci[OPTY_NINV];
#define CI(k) opty_ci[k]

struct OptyCtx {
  const double* xs;     // this lane's column in its staged segment; row pitch OPTY_XBOX
  double* con;          // &con[node]
  double* trow0;        // this lane's row in tile buffer 0
  double* trow1;        // this lane's row in tile buffer 1
  double* tile0;        // warp's tile buffer 0
  double* jac;          // p.jac
  const OptyTmaps* tm;
  long long ldc;
  int node;             // first node of the warp
  int lane;
  int n_nodes;
  bool active;          // node + lane < n_nodes
};

static __device__ __forceinline__ double opty_sign(double x) {
  return (double)((x > 0.0) - (x < 0.0));
}

static __device__ __forceinline__ uint32_t opty_smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}

static __device__ __forceinline__ void opty_mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(opty_smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

static __device__ __forceinline__ void opty_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(opty_smem_u32(bar)), "r"(bytes)
               : "memory");
}

static __device__ __forceinline__ void opty_mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "OPTY_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra OPTY_DONE;\n"
      "bra OPTY_WAIT;\n"
      "OPTY_DONE:\n"
      "}" ::"r"(opty_smem_u32(bar)),
      "r"(parity)
      : "memory");
}

static __device__ __forceinline__ void opty_tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1,
                                                        uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
          opty_smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(opty_smem_u32(bar))
      : "memory");
}

static __device__ __forceinline__ void opty_tma_store_2d(const CUtensorMap* map, const void* src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(c0), "r"(c1), "r"(opty_smem_u32(src))
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}

// trajectory value of row r at this lane's node (A) and at the next node (B)
#define XA(r) ctx.xs[(r) * OPTY_XBOX]
#define XB(r) ctx.xs[(r) * OPTY_XBOX + 1]

#define OPTY_CON(j, val)                                   \
  do {                                                     \
    if (ctx.active) ctx.con[(long long)(j) * ctx.ldc] = (val); \
  } while (0)

#define OPTY_TROW(buf) ((buf) ? ctx.trow1 : ctx.trow0)
#define OPTY_JS2(buf, tc, v0, v1) *reinterpret_cast<double2*>(OPTY_TROW(buf) + (tc)) = make_double2((v0), (v1))
#define OPTY_JS1(buf, tc, v0) OPTY_TROW(buf)[(tc)] = (v0)

// Hands the warp's finished tile (chunk `q` of group `g`: node rows
// ctx.node..+31, Jacobian columns col0 + q*C .. + ncols) to the TMA unit, or
// copies it out with coalesced warp-per-node stores.
template <int G, int Q, int COL0, int NCOLS>
static __device__ __forceinline__ void opty_flush(const OptyCtx& ctx) {
  double* tile = ctx.tile0 + (Q & 1) * OPTY_TILE_DOUBLES;
#if OPTY_TMA_STORE
  // make the generic-proxy st.shared visible to the async proxy, then one
  // lane issues the tile store; at most one older store may still be reading
  // its buffer (the other one) when the warp continues
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
#if OPTY_BLOCK_SYNC
  // all warps of the block run the same group: meeting here keeps them on the
  // same instruction-cache lines (the kernel is instruction-fetch bound
  // otherwise)
  __syncthreads();
#else
  __syncwarp();
#endif
  if (ctx.lane == 0 && ctx.node < ctx.n_nodes && !OPTY_DEBUG_NOSTORE) {
    opty_tma_store_2d(&ctx.tm->out[G], tile, Q * OPTY_C, ctx.node);
    asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
  }
  __syncwarp();
#else
#if OPTY_BLOCK_SYNC
  __syncthreads();
#else
  __syncwarp();
#endif
  const int rows = min(32, ctx.n_nodes - ctx.node);
  for (int r = 0; r < rows; ++r) {
    double* dst = ctx.jac + (long long)(ctx.node + r) * OPTY_K + COL0 + Q * OPTY_C;
    const double* src = tile + r * OPTY_C;
#pragma unroll
    for (int c = 0; c < NCOLS; c += 32)
      if (c + ctx.lane < NCOLS) dst[c + ctx.lane] = src[c + ctx.lane];
  }
  __syncwarp();
#endif
}
#define OPTY_FLUSH(g, q, col0, ncols) opty_flush<g, q, col0, ncols>(ctx)

#if OPTY_TMA_STORE
#define OPTY_DRAIN()                                                              \
  do {                                                                            \
    if (ctx.lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); \
    __syncwarp();                                                                 \
  } while (0)
#else
#define OPTY_DRAIN() \
  do {               \
  } while (0)
#endif

// dynamic shared memory: [WARPS][2][32][C] Jacobian tiles | [NSEG][R][XBOX]
// trajectory segments | mbarrier
#define OPTY_SMEM_TILES_BYTES (OPTY_WARPS * 2 * OPTY_TILE_DOUBLES * 8)
#define OPTY_SMEM_XIN_BYTES (OPTY_NSEG * OPTY_XSEG_BYTES)
#define OPTY_SMEM_BYTES (OPTY_SMEM_TILES_BYTES + OPTY_SMEM_XIN_BYTES + 128)

#if OPTY_TMA_LOAD
#define OPTY_STAGE_INPUT()                                                              \
  if (threadIdx.x == 0) opty_mbar_init(bar, 1);                                         \
  __syncthreads();                                                                      \
  if (threadIdx.x == 0) {                                                               \
    opty_mbar_expect_tx(bar, OPTY_NSEG * OPTY_R * OPTY_XBOX * 8);                       \
    for (int sgm = 0; sgm < OPTY_NSEG; ++sgm)                                           \
      opty_tma_load_2d(xin_bytes + sgm * OPTY_XSEG_BYTES, &tm.in, block_node0 + sgm * OPTY_XSEG, 0, bar); \
  }                                                                                     \
  opty_mbar_wait(bar, 0);
#else
#define OPTY_STAGE_INPUT()                                                              \
  for (int sgm = 0; sgm < OPTY_NSEG; ++sgm) {                                           \
    double* dstseg = reinterpret_cast<double*>(xin_bytes + sgm * OPTY_XSEG_BYTES);      \
    for (int r = 0; r < OPTY_R; ++r)                                                    \
      for (int c = threadIdx.x; c < OPTY_XBOX; c += OPTY_THREADS) {                     \
        const int col = block_node0 + sgm * OPTY_XSEG + c;                              \
        dstseg[r * OPTY_XBOX + c] = (col < p.n_cols) ? __ldg(p.traj + (long long)r * p.ldt + col) : 0.0; \
      }                                                                                 \
  }                                                                                     \
  __syncthreads();
#endif

#if OPTY_BLOCK_SYNC
#define OPTY_SKIP_IDLE_WARP()
#else
#define OPTY_SKIP_IDLE_WARP() if (ctx.node >= p.n_nodes) return;
#endif

#define OPTY_PROLOGUE()                                                                 \
  extern __shared__ __align__(128) unsigned char opty_smem[];                           \
  double* tiles = reinterpret_cast<double*>(opty_smem);                                 \
  unsigned char* xin_bytes = opty_smem + OPTY_SMEM_TILES_BYTES;                         \
  uint64_t* bar = reinterpret_cast<uint64_t*>(opty_smem + OPTY_SMEM_TILES_BYTES + OPTY_SMEM_XIN_BYTES); \
  const int block_node0 = blockIdx.x * OPTY_THREADS;                                    \
  (void)bar;                                                                            \
  OPTY_STAGE_INPUT()                                                                    \
  OptyCtx ctx;                                                                          \
  ctx.lane = threadIdx.x & 31;                                                          \
  ctx.node = block_node0 + (threadIdx.x & ~31);                                         \
  ctx.n_nodes = p.n_nodes;                                                              \
  ctx.active = (block_node0 + (int)threadIdx.x) < p.n_nodes;                            \
  ctx.xs = reinterpret_cast<const double*>(xin_bytes + (threadIdx.x / OPTY_XSEG) * OPTY_XSEG_BYTES) + \
           (threadIdx.x % OPTY_XSEG);                                                   \
  ctx.con = p.con + block_node0 + threadIdx.x;                                          \
  ctx.ldc = p.ldc;                                                                      \
  ctx.tile0 = tiles + (threadIdx.x >> 5) * 2 * OPTY_TILE_DOUBLES;                       \
  ctx.trow0 = ctx.tile0 + ctx.lane * OPTY_C;                                            \
  ctx.trow1 = ctx.trow0 + OPTY_TILE_DOUBLES;                                            \
  ctx.jac = p.jac;                                                                      \
  ctx.tm = &tm;                                                                         \
  OPTY_SKIP_IDLE_WARP()
