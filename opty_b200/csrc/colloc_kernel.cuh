// Hand-written sm_100a skeleton of the collocation constraint + Jacobian kernels.
//
// The generated module (opty_b200/codegen.py) defines the problem sizes
// (OPTY_M, OPTY_P, OPTY_K, OPTY_R, OPTY_D, OPTY_C, ...), the straight-line body
// of the pre-pass kernel and one straight-line `opty_group_<g>` device function
// per output group, then includes this file, which supplies everything around
// the arithmetic:
//
//   * the block -> (node tile, output group) mapping,
//   * staging of a tile's slice of the trajectory matrix (plus the derived rows
//     written by the pre-pass) into shared memory with 2-D TMA tile loads
//     (cp.async.bulk.tensor, mbarrier completion),
//   * the per-warp, double-buffered, bank-conflict-free staging tile for the
//     node-major Jacobian block and its drain by 2-D TMA tile stores,
//   * a coalesced warp-per-node fallback for shapes TMA cannot describe.
//
// Work mapping: lane = collocation node (all lanes of a warp execute the same
// generated instruction stream), warp = 32 consecutive nodes x one output
// group.  It replaces the node loop of the reference's generated Cython
// (`for i in prange(n)`, opty/utils.py:524-526) and the per-node `eval_matrix`
// C function (opty/utils.py:483-494).
//
// Data layout (all float64):
//   traj : [R + D][ldt]  rows = states, unknown inputs, known inputs, then the D
//                        derived rows (shared transcendental sub-expressions,
//                        one value per constraint node); col = node
//   con  : [M][ldc]      eom-major residuals, node i at column i (layout of
//                        opty/direct_collocation.py:2446 without the transpose)
//   jac  : [nodes][K]    node-major partials, K = M*P, incl. structural zeros
//                        (layout of opty/direct_collocation.py:2814, 2887)
#pragma once

#include <cuda.h>
#include <stdint.h>

#include "colloc_params.h"

#define OPTY_THREADS (OPTY_WARPS * 32)
#define OPTY_RD (OPTY_R + OPTY_D)
// A tile's slice of the trajectory matrix is staged in segments of OPTY_XSEG
// nodes; a segment holds one column per node plus the right neighbour, rounded
// up to an even count (TMA rows are 16-byte multiples, TMA boxes at most 256
// elements wide).  Blocks wider than 128 threads use several overlapping
// segments.
#if OPTY_THREADS <= 128
#define OPTY_XSEG OPTY_THREADS
#else
#define OPTY_XSEG 128
#endif
#define OPTY_XBOX (OPTY_XSEG + 2)
#define OPTY_NSEG (OPTY_THREADS / OPTY_XSEG)
#define OPTY_XSEG_BYTES (((OPTY_RD * OPTY_XBOX * 8) + 127) / 128 * 128)
#define OPTY_TILE_DOUBLES (32 * OPTY_C)
#ifndef OPTY_NBUF
#define OPTY_NBUF 2  // staging tiles per warp (TMA stores in flight + 1)
#endif
#ifndef OPTY_DEBUG_NOSTORE
#define OPTY_DEBUG_NOSTORE 0  // measurement aid: skip the Jacobian tile stores
#endif

struct OptyTmaps {
  CUtensorMap in;                 // traj as {cols, R + D}
  CUtensorMap out[OPTY_NSEGS];    // store segment s (a run of columns written by one group) of jac as
                                  // {ncols_s, nodes}
};

// node-invariant sub-expressions, filled by the host from opty_colloc_inv
__constant__ double opty_ci[OPTY_NINV];
#define CI(k) opty_ci[k]

struct OptyCtx {
  const double* xs;  // this lane's column in its staged segment (row pitch OPTY_XBOX) or, with direct
                     // input loads, in the trajectory matrix itself (row pitch ldt)
  long long ldt;
  double* con;       // &con[node of this lane]
  double* trow0;     // this lane's row in tile buffer 0 (buffer b: + b*OPTY_TILE_DOUBLES)
  double* tile0;     // warp's tile buffer 0
  double* jac;       // p.jac
  const OptyTmaps* tm;
  long long ldc;
  int node;          // first node of the warp
  int lane;
  int n_nodes;
  bool active;       // node + lane < n_nodes
};

static __device__ __forceinline__ double opty_sign(double x) { return (double)((x > 0.0) - (x < 0.0)); }

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

// ---------------------------------------------------------------------------
// main kernel: operand and output macros used by the generated group bodies
// ---------------------------------------------------------------------------
// trajectory value of row r at this lane's node (A) and at the next node (B);
// derived row d (pre-pass output) at this lane's node
#if OPTY_TMA_LOAD == 2
// direct mode: no staging; lanes read consecutive columns of a row (coalesced,
// read-only path), the neighbour column comes from the same cache lines
#define XA(r) __ldg(ctx.xs + (long long)(r) * ctx.ldt)
#define XB(r) __ldg(ctx.xs + (long long)(r) * ctx.ldt + 1)
#define XD(d) __ldg(ctx.xs + (long long)(OPTY_R + (d)) * ctx.ldt)
#else
#define XA(r) ctx.xs[(r) * OPTY_XBOX]
#define XB(r) ctx.xs[(r) * OPTY_XBOX + 1]
#define XD(d) ctx.xs[(OPTY_R + (d)) * OPTY_XBOX]
#endif

#define OPTY_CON(j, val)                                       \
  do {                                                         \
    if (ctx.active) ctx.con[(long long)(j) * ctx.ldc] = (val); \
  } while (0)

#define OPTY_TROW(buf) (ctx.trow0 + (buf) * OPTY_TILE_DOUBLES)
#define OPTY_JS2(buf, tc, v0, v1) *reinterpret_cast<double2*>(OPTY_TROW(buf) + (tc)) = make_double2((v0), (v1))
#define OPTY_JS1(buf, tc, v0) OPTY_TROW(buf)[(tc)] = (v0)

// Hands the warp's finished tile (chunk `Q` of store segment `SEG`: node rows
// ctx.node..+31, Jacobian columns SEGCOL0 + Q*C .. + NCOLS, staged in tile
// buffer `BUF`) to the TMA unit, or copies it out with coalesced
// warp-per-node stores.
template <int SEG, int Q, int BUF, int SEGCOL0, int NCOLS>
static __device__ __forceinline__ void opty_flush(const OptyCtx& ctx) {
  double* tile = ctx.tile0 + BUF * OPTY_TILE_DOUBLES;
#if OPTY_TMA_STORE
  // make the generic-proxy st.shared visible to the async proxy, then one
  // lane issues the tile store; at most OPTY_NBUF-1 older stores may still be
  // reading their buffers when the warp continues with the next buffer
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  // OPTY_DEBUG_NOSTORE (measurement aid, wrong results): 1 = no tile stores,
  // 2 = all tile stores go to node rows 0..31 (no HBM write stream),
  // 3 = stores without waiting for the staging buffer to be free again
  if (ctx.lane == 0 && ctx.node < ctx.n_nodes && OPTY_DEBUG_NOSTORE != 1) {
    opty_tma_store_2d(&ctx.tm->out[SEG], tile, Q * OPTY_C, OPTY_DEBUG_NOSTORE == 2 ? 0 : ctx.node);
    if (OPTY_DEBUG_NOSTORE != 3) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(OPTY_NBUF - 1) : "memory");
  }
  __syncwarp();
#else
  __syncwarp();
  const int rows = min(32, ctx.n_nodes - ctx.node);
  for (int r = 0; r < rows; ++r) {
    double* dst = ctx.jac + (long long)(ctx.node + r) * OPTY_K + SEGCOL0 + Q * OPTY_C;
    const double* src = tile + r * OPTY_C;
#pragma unroll
    for (int c = 0; c < NCOLS; c += 32)
      if (c + ctx.lane < NCOLS) dst[c + ctx.lane] = src[c + ctx.lane];
  }
  __syncwarp();
#endif
}
#define OPTY_FLUSH(seg, q, buf, segcol0, ncols) opty_flush<seg, q, buf, segcol0, ncols>(ctx)

// end of a group body: the warp's tile buffers are reused by its next tile
#if OPTY_TMA_STORE
#define OPTY_DRAIN()                                                                  \
  do {                                                                                \
    if (ctx.lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); \
    __syncwarp();                                                                     \
  } while (0)
#else
#define OPTY_DRAIN() \
  do {               \
  } while (0)
#endif

// dynamic shared memory: [WARPS][NBUF][32][C] Jacobian tiles | [NSEG][R+D][XBOX]
// trajectory segments | mbarrier, tile slot
#define OPTY_SMEM_TILES_BYTES (OPTY_WARPS * OPTY_NBUF * OPTY_TILE_DOUBLES * 8)
#if OPTY_TMA_LOAD == 2
#define OPTY_SMEM_XIN_BYTES 0
#else
#define OPTY_SMEM_XIN_BYTES (OPTY_NSEG * OPTY_XSEG_BYTES)
#endif
#define OPTY_SMEM_BYTES (OPTY_SMEM_TILES_BYTES + OPTY_SMEM_XIN_BYTES + 128)

#if OPTY_TMA_LOAD == 2
#define OPTY_STAGE_INPUT()
#elif OPTY_TMA_LOAD == 1
#define OPTY_STAGE_INPUT()                                                                               \
  if (threadIdx.x == 0) {                                                                                \
    opty_mbar_expect_tx(bar, OPTY_NSEG * OPTY_RD * OPTY_XBOX * 8);                                       \
    for (int sgm = 0; sgm < OPTY_NSEG; ++sgm)                                                            \
      opty_tma_load_2d(xin_bytes + sgm * OPTY_XSEG_BYTES, &tm.in, tile_node0 + sgm * OPTY_XSEG, 0, bar); \
  }                                                                                                      \
  opty_mbar_wait(bar, phase);                                                                            \
  phase ^= 1u;
#else
#define OPTY_STAGE_INPUT()                                                                               \
  for (int sgm = 0; sgm < OPTY_NSEG; ++sgm) {                                                            \
    double* dstseg = reinterpret_cast<double*>(xin_bytes + sgm * OPTY_XSEG_BYTES);                       \
    for (int r = 0; r < OPTY_RD; ++r)                                                                    \
      for (int c = threadIdx.x; c < OPTY_XBOX; c += OPTY_THREADS) {                                      \
        const int col = tile_node0 + sgm * OPTY_XSEG + c;                                                \
        dstseg[r * OPTY_XBOX + c] = (col < p.n_cols) ? __ldcg(p.traj + (long long)r * p.ldt + col) : 0.0; \
      }                                                                                                  \
  }                                                                                                      \
  __syncthreads();
#endif

// One block = one tile of 32*W nodes x one output group; grid = (tiles, groups).
// blockIdx.y walks the groups in the order the emitter chose (most expensive
// first, so that the cheap groups fill the tail of the launch); the hardware
// block scheduler balances the SMs (measured: a persistent variant with global
// tile counters or a static schedule loses more to scheduling round trips and
// imbalance than it gains in instruction-cache locality).
//
// The generated kernel body sits between OPTY_KERNEL_BEGIN and OPTY_KERNEL_END
// and dispatches on `opty_g`.
#ifndef OPTY_TILE_MAJOR
#define OPTY_TILE_MAJOR 0
#endif
#if OPTY_TILE_MAJOR
// tile-major dispatch: consecutive blocks are the groups of ONE node tile, so a
// node row's 8 KB are written within a short time window (DRAM page locality of
// the write-back stream) instead of in one phase per group
#define OPTY_BLOCK_TO_WORK()                                                        \
  const unsigned opty_lin = blockIdx.y * gridDim.x + blockIdx.x;                    \
  const int opty_g = opty_group_order[opty_lin % OPTY_NGROUPS];                     \
  const int tile_node0 = (int)(opty_lin / OPTY_NGROUPS) * OPTY_THREADS;
#else
#define OPTY_BLOCK_TO_WORK()                              \
  const int opty_g = opty_group_order[blockIdx.y];        \
  const int tile_node0 = blockIdx.x * OPTY_THREADS;
#endif
#define OPTY_KERNEL_BEGIN()                                                                              \
  extern __shared__ __align__(128) unsigned char opty_smem[];                                            \
  double* tiles = reinterpret_cast<double*>(opty_smem);                                                  \
  unsigned char* xin_bytes = opty_smem + OPTY_SMEM_TILES_BYTES;                                          \
  uint64_t* bar = reinterpret_cast<uint64_t*>(opty_smem + OPTY_SMEM_TILES_BYTES + OPTY_SMEM_XIN_BYTES);  \
  uint32_t phase = 0;                                                                                    \
  (void)phase;                                                                                           \
  if (OPTY_TMA_LOAD == 1) {                                                                              \
    if (threadIdx.x == 0) opty_mbar_init(bar, 1);                                                        \
    __syncthreads();                                                                                     \
  }                                                                                                      \
  OPTY_BLOCK_TO_WORK()                                                                                   \
  OPTY_STAGE_INPUT()                                                                                     \
  OptyCtx ctx;                                                                                           \
  ctx.lane = threadIdx.x & 31;                                                                           \
  ctx.n_nodes = p.n_nodes;                                                                               \
  ctx.ldt = p.ldt;                                                                                       \
  if (OPTY_TMA_LOAD == 2)                                                                                \
    ctx.xs = p.traj + min(tile_node0 + (int)threadIdx.x, p.n_nodes - 1);                                 \
  else                                                                                                   \
    ctx.xs = reinterpret_cast<const double*>(xin_bytes + (threadIdx.x / OPTY_XSEG) * OPTY_XSEG_BYTES) +  \
             (threadIdx.x % OPTY_XSEG);                                                                  \
  ctx.ldc = p.ldc;                                                                                       \
  ctx.tile0 = tiles + (threadIdx.x >> 5) * OPTY_NBUF * OPTY_TILE_DOUBLES;                                \
  ctx.trow0 = ctx.tile0 + ctx.lane * OPTY_C;                                                             \
  ctx.jac = p.jac;                                                                                       \
  ctx.tm = &tm;                                                                                          \
  ctx.node = tile_node0 + (threadIdx.x & ~31);                                                           \
  ctx.active = (tile_node0 + (int)threadIdx.x) < p.n_nodes;                                              \
  ctx.con = p.con + tile_node0 + threadIdx.x;                                                            \
  if (ctx.node >= p.n_nodes) return;

#define OPTY_KERNEL_END()

// ---------------------------------------------------------------------------
// pre-pass kernel: one thread per node, derived rows written coalesced
// ---------------------------------------------------------------------------
#define OPTY_PRE_THREADS 128
#define GA(r) __ldcg(xg + (long long)(r) * p.ldt)
#define GB(r) __ldcg(xg + (long long)(r) * p.ldt + 1)
#define OPTY_DRV(d, val) drv[(long long)(d) * p.ldt] = (val)
#define OPTY_PRE_BEGIN()                                        \
  const int node = blockIdx.x * OPTY_PRE_THREADS + threadIdx.x; \
  if (node >= p.n_nodes) return;                                \
  const double* xg = p.traj + node;                             \
  double* drv = p.traj + (long long)OPTY_R * p.ldt + node;      \
  const int opty_pg = blockIdx.y;

