// Hand-written sm_100a skeleton of the collocation constraint + Jacobian kernels.
//
// The generated module (opty_b200/codegen.py) defines the problem sizes
// (OPTY_M, OPTY_P, OPTY_K, OPTY_R, OPTY_D, ...), the straight-line body of the
// pre-pass kernel and one straight-line `opty_group_<g>` device function per
// output group, then includes this file, which supplies everything around the
// arithmetic:
//
//   * the block -> (node tile, output group) mapping,
//   * staging of a tile's slice of the trajectory matrix (plus the derived rows
//     written by the pre-pass) into shared memory with 2-D TMA tile loads
//     (cp.async.bulk.tensor, mbarrier completion),
//   * the per-warp shared-memory staging buffers for the node-major Jacobian
//     block and their drain by 2-D TMA tile stores,
//   * a coalesced warp-per-node fallback for shapes TMA cannot describe.
//
// Work mapping: lane = collocation node (all lanes of a warp execute the same
// generated instruction stream), warp = 32 consecutive nodes x one output
// group.  It replaces the node loop of the reference's generated Cython
// (`for i in prange(n)`, opty/utils.py:524-526) and the per-node `eval_matrix`
// C function (opty/utils.py:483-494).
//
// A group body is a sequence of PHASES.  A phase owns a few column runs of the
// node block ("sub-tiles": [32 nodes][w columns], dense, one after the other in
// one of the warp's OPTY_NBUF staging buffers); the body writes the phase's
// entries in whatever order the register-pressure scheduler (schedule.py)
// produced them, then hands every sub-tile to the TMA unit with one tile store.
// Sub-tiles of the same width share a tensor map over the whole Jacobian
// ({K, nodes}, box {w, 32}); rows beyond the shard are clipped by the TMA unit.
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
#include <stdio.h>

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
#define OPTY_TW (OPTY_THREADS + 2)  // row pitch of the tiled trajectory layout (direct input loads)
#define OPTY_NSEG (OPTY_THREADS / OPTY_XSEG)
#define OPTY_XSEG_BYTES (((OPTY_RD * OPTY_XBOX * 8) + 127) / 128 * 128)
#ifndef OPTY_PERSISTENT
#define OPTY_PERSISTENT 0
#endif
#ifndef OPTY_NBUF
#define OPTY_NBUF 2  // staging buffers per warp (phases whose TMA stores may be in flight + 1)
#endif
#ifndef OPTY_DEBUG_NOSTORE
#define OPTY_DEBUG_NOSTORE 0  // measurement aid: 1 = skip the Jacobian tile stores (wrong results)
#endif

struct OptyTmaps {
  CUtensorMap in;               // traj as {cols, R + D}, box {OPTY_XBOX, R + D}
  CUtensorMap out[OPTY_NMAPS];  // jac as {K, nodes}, box {OPTY_MAP_WIDTH[i], 32}
};

// node-invariant sub-expressions, filled by the host from opty_colloc_inv
__constant__ double opty_ci[OPTY_NINV];
#define CI(k) opty_ci[k]

struct OptyCtx {
#if OPTY_TMA_LOAD == 2
  const double* xg;  // this lane's column of its tile in the tiled trajectory layout (row pitch OPTY_TW)
#else
  uint32_t xs;       // shared-memory address of this lane's column in its staged segment (row pitch OPTY_XBOX)
  const double* xp;  // the same as a pointer (plain loads)
#endif
  long long ldt;
  double* con;       // &con[node of this lane]
  double* tile0;     // the warp's staging buffer 0 (buffer b: + b * OPTY_TILE_DOUBLES)
#if OPTY_PERSISTENT == 2
  double* tile1;     // staging buffer 1 (the two swap after an item with an odd number of phases)
  int* cn;           // constant runs: next node of this warp's early share, end of the share, nodes per
  int ce, ctick;     // OPTY_TICK(), the block's copy of the runs in shared memory
  const double* cbuf;
#endif
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

// (input tiles are read by every group: evict_last keeps the few MB of the trajectory matrix in L2 beside
// the output stream -- without the hint the row-stationary kernel re-read 20 MB of it from DRAM per launch)
static __device__ __forceinline__ void opty_tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1,
                                                        uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], "
      "[%4], %5;" ::"r"(opty_smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(opty_smem_u32(bar)), "l"(0x14F0000000000000ull)
      : "memory");
}

// L2 policy of the Jacobian stores.  The output (81 MB per evaluation at BASELINE config 2, 126 MB of L2)
// is written once and read by nobody on the device; with the default policy every launch pushes the
// module's code and the trajectory matrix out of L2, and the next launch re-reads both through a cache
// full of dirty lines.  evict_first makes the freshly written lines the first candidates for replacement.
#ifndef OPTY_STORE_HINT
#define OPTY_STORE_HINT 0  // 0 none, 1 evict_first, 2 evict_last (measurement), 3 no_allocate-like evict_unchanged
#endif
#if OPTY_STORE_HINT == 1
#define OPTY_STORE_POLICY 0x12F0000000000000ull
#elif OPTY_STORE_HINT == 2
#define OPTY_STORE_POLICY 0x14F0000000000000ull
#else
#define OPTY_STORE_POLICY 0x1000000000000000ull
#endif
static __device__ __forceinline__ void opty_tma_store_2d(const CUtensorMap* map, const void* src, int c0, int c1) {
#if OPTY_STORE_HINT
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%1, %2}], [%3], %4;" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(c0), "r"(c1), "r"(opty_smem_u32(src)), "l"(OPTY_STORE_POLICY)
               : "memory");
#else
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(c0), "r"(c1), "r"(opty_smem_u32(src))
               : "memory");
#endif
}

// 1-D bulk copy shared -> global (constant column runs)
static __device__ __forceinline__ void opty_bulk_store_1d(void* dst, const void* src, uint32_t bytes) {
#if OPTY_STORE_HINT
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(dst),
               "r"(opty_smem_u32(src)), "r"(bytes), "l"(OPTY_STORE_POLICY)
               : "memory");
#else
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(opty_smem_u32(src)),
               "r"(bytes)
               : "memory");
#endif
}

// ---------------------------------------------------------------------------
// constant rows with the grid kernel: `opty_colloc_const`
// ---------------------------------------------------------------------------
// Equations whose partials are all literals or node-invariant values (x' = v and the like: half of the
// Jacobian of a mechanical system in first-order form) are not output groups.  Their part of every node's
// Jacobian row is the same run of numbers (values p.cvals, filled by the invariants kernel; runs opty_crun):
// a persistent grid of 256-thread blocks keeps one copy in shared memory and every thread sends the runs of
// its nodes with bulk copies of at most 16 KB.  (The row-stationary kernel sends them itself, in slices.)
#define OPTY_CONST_KERNEL_BODY()                                                                          \
  extern __shared__ __align__(128) unsigned char opty_smem[];                                             \
  double* cbuf_ = reinterpret_cast<double*>(opty_smem);                                                   \
  for (int i_ = threadIdx.x; i_ < OPTY_NCONST; i_ += blockDim.x) cbuf_[i_] = p.cvals[i_];                 \
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");                                            \
  __syncthreads();                                                                                        \
  for (int n_ = blockIdx.x * blockDim.x + threadIdx.x; n_ < p.n_nodes; n_ += gridDim.x * blockDim.x)      \
    for (int r_ = 0; r_ < OPTY_NCRUNS; ++r_) {                                                            \
      const int len_ = opty_crun[r_][1] * 2; /* doubles */                                                \
      for (int o_ = 0; o_ < len_; o_ += 2048)                                                             \
        opty_bulk_store_1d(p.jac + (long long)n_ * OPTY_K + opty_crun[r_][0] + o_, cbuf_ + opty_crun[r_][2] + o_, \
                           (uint32_t)min(2048, len_ - o_) * 8u);                                          \
    }                                                                                                     \
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");                                               \
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");

// ---------------------------------------------------------------------------
// main kernel: operand and output macros used by the generated group bodies
// ---------------------------------------------------------------------------
// Input loads.  With OPTY_VOLATILE_LOADS they are `asm volatile`: every textual
// load is one load instruction.  The scheduler decides where an input is held
// in a register and where it is read again (schedule.py); a compiler that
// merged the loads would pin the register for the whole distance between
// them -- what large bodies cannot afford.  Small bodies are better off with
// plain loads that nvcc may merge (fewer instructions to fetch).
#ifndef OPTY_VOLATILE_LOADS
#define OPTY_VOLATILE_LOADS 1
#endif
#if OPTY_PERSISTENT == 2
// row-stationary kernel: the block's input buffer holds the rows
// OPTY_XROW0 .. of the item's group (the emitter defines OPTY_XROW0 in front of
// every body) in segments of OPTY_XSEG nodes, row pitch OPTY_XBOX doubles
#define OPTY_PXBOX OPTY_XBOX
#if OPTY_VOLATILE_LOADS
static __device__ __forceinline__ double opty_ldin(uint32_t a) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
  return v;
}
#define XA(r) opty_ldin(ctx.xs + ((r) - OPTY_XROW0) * (OPTY_PXBOX * 8))
#define XB(r) opty_ldin(ctx.xs + ((r) - OPTY_XROW0) * (OPTY_PXBOX * 8) + 8)
#define XD(d) opty_ldin(ctx.xs + ((OPTY_R + (d)) - OPTY_XROW0) * (OPTY_PXBOX * 8))
#else
#define XA(r) ctx.xp[((r) - OPTY_XROW0) * OPTY_PXBOX]
#define XB(r) ctx.xp[((r) - OPTY_XROW0) * OPTY_PXBOX + 1]
#define XD(d) ctx.xp[((OPTY_R + (d)) - OPTY_XROW0) * OPTY_PXBOX]
#endif
#elif OPTY_TMA_LOAD == 2
// direct mode: no shared-memory staging; the pre-pass kernel lays the
// trajectory matrix and the derived rows out tile by tile
// ([tiles][R + D][OPTY_TW], OPTY_TW = nodes of a block + 2), so that every row
// of a block's slice sits at a compile-time offset from one base pointer: a
// load is `ld.global.nc [base + immediate]`, lanes read consecutive columns
// (coalesced), the neighbour column comes from the same cache lines.  (With a
// run-time row pitch every row needs its own 64-bit address register; ptxas
// keeps hundreds of them alive in large bodies and spills.)
static __device__ __forceinline__ double opty_ldin(const double* p) {
#if OPTY_VOLATILE_LOADS
  double v;
  asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
#else
  return __ldg(p);
#endif
}
#define XA(r) opty_ldin(ctx.xg + (r) * OPTY_TW)
#define XB(r) opty_ldin(ctx.xg + (r) * OPTY_TW + 1)
#define XD(d) opty_ldin(ctx.xg + (OPTY_R + (d)) * OPTY_TW)
#elif OPTY_VOLATILE_LOADS
static __device__ __forceinline__ double opty_ldin(uint32_t a) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
  return v;
}
#define XA(r) opty_ldin(ctx.xs + (r) * (OPTY_XBOX * 8))
#define XB(r) opty_ldin(ctx.xs + (r) * (OPTY_XBOX * 8) + 8)
#define XD(d) opty_ldin(ctx.xs + (OPTY_R + (d)) * (OPTY_XBOX * 8))
#else
#define XA(r) ctx.xp[(r) * OPTY_XBOX]
#define XB(r) ctx.xp[(r) * OPTY_XBOX + 1]
#define XD(d) ctx.xp[(OPTY_R + (d)) * OPTY_XBOX]
#endif

// scheduling fence: no memory operation moves across it (see codegen.py, `fence_every`)
#define OPTY_FENCE() __syncwarp()

#define OPTY_CON(j, val)                                       \
  do {                                                         \
    if (ctx.active) ctx.con[(long long)(j) * ctx.ldc] = (val); \
  } while (0)

// this lane's row of the sub-tile that starts `off` doubles into staging buffer `buf` and is `w` columns wide
#if OPTY_PERSISTENT == 2
#define OPTY_TBUF(buf) ((buf) ? ctx.tile1 : ctx.tile0)
#else
#define OPTY_TBUF(buf) (ctx.tile0 + (buf) * OPTY_TILE_DOUBLES)
#endif
#define OPTY_TROW(buf, off, w) (OPTY_TBUF(buf) + (off) + ctx.lane * (w))
#define OPTY_JS2(buf, off, w, c, v0, v1) \
  *reinterpret_cast<double2*>(OPTY_TROW(buf, off, w) + (c)) = make_double2((v0), (v1))
#define OPTY_JS1(buf, off, w, c, v0) OPTY_TROW(buf, off, w)[(c)] = (v0)

// an entry stored straight to global memory (outputs the scheduler moved to the
// end of the body, schedule.py `deferrable`); OPTY_DRAIN_WRITES comes first:
// the tile store of the phase that owns the column has then completed
#define OPTY_JG(col, val)                                                              \
  do {                                                                                 \
    if (ctx.active) ctx.jac[(long long)(ctx.node + ctx.lane) * OPTY_K + (col)] = (val); \
  } while (0)

// first statement of phase `t` (static index inside the body): its staging
// buffer is free again once at most OPTY_NBUF-1 younger store groups may still
// be reading theirs
#if OPTY_TMA_STORE
// (row-stationary kernel: a warp's store groups run on from item to item, so
// every phase waits)
#define OPTY_PHASE_BEGIN(t)                                                                                   \
  do {                                                                                                        \
    if ((t) >= OPTY_NBUF || OPTY_PERSISTENT == 2) {                                                           \
      if (ctx.lane == 0) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(OPTY_NBUF - 1) : "memory"); \
      __syncwarp();                                                                                           \
    }                                                                                                         \
  } while (0)
// the phase's entries are in shared memory: make the generic-proxy stores
// visible to the async proxy, then one lane issues the tile stores
#define OPTY_FLUSH_BEGIN()                                      \
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); \
  __syncwarp();
#define OPTY_TSTORE(map, buf, off, w, col0)                                                                   \
  if (ctx.lane == 0 && (OPTY_DEBUG_NOSTORE & 1) == 0)                                                               \
    opty_tma_store_2d(&ctx.tm->out[map], OPTY_TBUF(buf) + (off), (col0), ctx.node);
#define OPTY_FLUSH_END() \
  if (ctx.lane == 0) asm volatile("cp.async.bulk.commit_group;" ::: "memory");
#if OPTY_PERSISTENT == 2
#define OPTY_DRAIN() \
  do {               \
  } while (0)
#else
#define OPTY_DRAIN()                                                                  \
  do {                                                                                \
    if (ctx.lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); \
    __syncwarp();                                                                     \
  } while (0)
#endif
#define OPTY_DRAIN_WRITES()                                                      \
  do {                                                                           \
    if (ctx.lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); \
    __syncwarp();                                                                \
  } while (0)
#else
// shapes TMA cannot describe (odd M*P): the warp copies a sub-tile out itself,
// one node row per iteration, coalesced
#define OPTY_PHASE_BEGIN(t) __syncwarp()
#define OPTY_FLUSH_BEGIN() __syncwarp();
#define OPTY_TSTORE(map, buf, off, w, col0)                                            \
  {                                                                                    \
    const double* src_ = ctx.tile0 + (buf) * OPTY_TILE_DOUBLES + (off);                \
    const int rows_ = min(32, ctx.n_nodes - ctx.node);                                 \
    for (int r_ = 0; r_ < rows_; ++r_) {                                               \
      double* dst_ = ctx.jac + (long long)(ctx.node + r_) * OPTY_K + (col0);           \
      for (int c_ = ctx.lane; c_ < (w); c_ += 32) dst_[c_] = src_[r_ * (w) + c_];      \
    }                                                                                  \
  }
#define OPTY_FLUSH_END()
#define OPTY_DRAIN() \
  do {               \
  } while (0)
#define OPTY_DRAIN_WRITES() __syncwarp()
#endif

// dynamic shared memory: [WARPS][NBUF][OPTY_TILE_DOUBLES] staging buffers |
// [NSEG][R+D][XBOX] trajectory segments | mbarrier
#define OPTY_SMEM_TILES_BYTES (OPTY_WARPS * OPTY_NBUF * OPTY_TILE_DOUBLES * 8)
#if OPTY_TMA_LOAD == 2
#define OPTY_SMEM_XIN_BYTES 0
#else
#define OPTY_SMEM_XIN_BYTES (OPTY_NSEG * OPTY_XSEG_BYTES)
#endif
#define OPTY_SMEM_BYTES (OPTY_SMEM_TILES_BYTES + OPTY_SMEM_XIN_BYTES + 128)

// stages the slice of node tile `tile_node0` (mbarrier parity `xphase`)
#if OPTY_TMA_LOAD == 2
#define OPTY_STAGE_INIT()
#define OPTY_STAGE_INPUT()
#elif OPTY_TMA_LOAD == 1
#define OPTY_STAGE_INIT()                     \
  if (threadIdx.x == 0) opty_mbar_init(bar, 1); \
  __syncthreads();
#define OPTY_STAGE_INPUT()                                                                               \
  if (threadIdx.x == 0) {                                                                                \
    opty_mbar_expect_tx(bar, OPTY_NSEG * OPTY_RD * OPTY_XBOX * 8);                                       \
    for (int sgm = 0; sgm < OPTY_NSEG; ++sgm)                                                            \
      opty_tma_load_2d(xin_bytes + sgm * OPTY_XSEG_BYTES, &tm.in, tile_node0 + sgm * OPTY_XSEG, 0, bar); \
  }                                                                                                      \
  opty_mbar_wait(bar, xphase);                                                                           \
  xphase ^= 1u;
#else
#define OPTY_STAGE_INIT()
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

#if OPTY_TMA_LOAD == 2
#define OPTY_CTX_INPUT() \
  ctx.xg = p.tiled + (long long)(tile_node0 / OPTY_THREADS) * (OPTY_RD * OPTY_TW) + threadIdx.x;
#else
#define OPTY_CTX_INPUT()                                                                          \
  ctx.xp = reinterpret_cast<const double*>(xin_bytes + (threadIdx.x / OPTY_XSEG) * OPTY_XSEG_BYTES) + \
           (threadIdx.x % OPTY_XSEG);                                                             \
  ctx.xs = opty_smem_u32(ctx.xp);
#endif

#define OPTY_SMEM_SETUP()                                                                                \
  extern __shared__ __align__(128) unsigned char opty_smem[];                                            \
  double* tiles = reinterpret_cast<double*>(opty_smem);                                                  \
  unsigned char* xin_bytes = opty_smem + OPTY_SMEM_TILES_BYTES;                                          \
  uint64_t* bar = reinterpret_cast<uint64_t*>(opty_smem + OPTY_SMEM_TILES_BYTES + OPTY_SMEM_XIN_BYTES);  \
  uint32_t xphase = 0;                                                                                   \
  (void)bar;                                                                                             \
  (void)xin_bytes;                                                                                       \
  (void)xphase;

#define OPTY_CTX_SETUP()                                                  \
  OptyCtx ctx;                                                            \
  ctx.lane = threadIdx.x & 31;                                            \
  ctx.n_nodes = p.n_nodes;                                                \
  ctx.ldt = p.ldt;                                                        \
  OPTY_CTX_INPUT()                                                        \
  ctx.ldc = p.ldc;                                                        \
  ctx.tile0 = tiles + (threadIdx.x >> 5) * OPTY_NBUF * OPTY_TILE_DOUBLES; \
  ctx.jac = p.jac;                                                        \
  ctx.tm = &tm;                                                           \
  ctx.node = tile_node0 + (threadIdx.x & ~31);                            \
  ctx.active = (tile_node0 + (int)threadIdx.x) < p.n_nodes;               \
  ctx.con = p.con + tile_node0 + threadIdx.x;

#if !OPTY_PERSISTENT
// Grid kernel: one block = one tile of 32*W nodes x one output group; grid =
// (tiles, groups).  blockIdx.y walks the groups in the order the emitter chose
// (most expensive first, so that the cheap groups fill the tail of the
// launch); the hardware block scheduler balances the SMs.
//
// The generated kernel body sits between OPTY_KERNEL_BEGIN and OPTY_KERNEL_END
// and dispatches on `opty_g`.
// OPTY_TILE_MAJOR: consecutive blocks (the hardware deals them out in the order x fastest, then y) are the
// groups of ONE node tile instead of the node tiles of one group, so that the pieces of a node's Jacobian
// row are written at about the same time (a DRAM page written in several visits costs several activations,
// tools/write_path_bench.cu variants I/J)
#ifndef OPTY_TILE_MAJOR
#define OPTY_TILE_MAJOR 0
#endif
#if OPTY_TILE_MAJOR
#define OPTY_BLOCK_TO_WORK()                                                   \
  const unsigned opty_lin = blockIdx.y * gridDim.x + blockIdx.x;               \
  const int opty_g = opty_group_order[opty_lin % OPTY_NGROUPS];                \
  const int tile_node0 = (int)(opty_lin / OPTY_NGROUPS) * OPTY_THREADS;
#else
#define OPTY_BLOCK_TO_WORK()                         \
  const int opty_g = opty_group_order[blockIdx.y];   \
  const int tile_node0 = blockIdx.x * OPTY_THREADS;
#endif
// Constant rows with the grid kernel (OPTY_GRID_CONST): the block of (node tile, group g) first sends the
// constant column runs of every OPTY_NGROUPS_ALL-th node of its tile -- plain streaming stores, threads =
// consecutive 16-byte pieces of a run, values from the tail of the invariants table in L2 -- so that they
// drain while the body runs and reach memory together with the tile's other pieces.
#ifndef OPTY_GRID_CONST
#define OPTY_GRID_CONST 0
#endif
#if OPTY_GRID_CONST && OPTY_NCRUNS > 0
#define OPTY_GRID_CONST_SHARE()                                                                        \
  if ((OPTY_DEBUG_NOSTORE & 1) == 0) {                                                                 \
    const int last_ = min(tile_node0 + OPTY_THREADS, p.n_nodes);                                       \
    int n_ = tile_node0 + ((OPTY_G0 + opty_g) - tile_node0 % OPTY_NGROUPS_ALL + OPTY_NGROUPS_ALL) % OPTY_NGROUPS_ALL; \
    for (; n_ < last_; n_ += OPTY_NGROUPS_ALL)                                                         \
      for (int r_ = 0; r_ < OPTY_NCRUNS; ++r_) {                                                       \
        const double2* v_ = reinterpret_cast<const double2*>(p.cvals + opty_crun[r_][2]);              \
        double2* d_ = reinterpret_cast<double2*>(p.jac + (long long)n_ * OPTY_K + opty_crun[r_][0]);   \
        for (int k_ = threadIdx.x; k_ < opty_crun[r_][1]; k_ += OPTY_THREADS) __stcs(d_ + k_, __ldg(v_ + k_)); \
      }                                                                                                \
  }
#else
#define OPTY_GRID_CONST_SHARE()
#endif
#define OPTY_KERNEL_BEGIN()                          \
  OPTY_SMEM_SETUP()                                  \
  OPTY_BLOCK_TO_WORK()                               \
  OPTY_GRID_CONST_SHARE()                            \
  OPTY_STAGE_INIT()                                  \
  OPTY_STAGE_INPUT()                                 \
  OPTY_CTX_SETUP()                                   \
  if (ctx.node >= p.n_nodes) return;

#define OPTY_KERNEL_END()

#elif OPTY_PERSISTENT == 2
// Row-stationary persistent kernel: grid = OPTY_NSLOTS resident blocks (one per
// SM).  The emitter lays the work out statically (codegen.stationary_schedule):
// a slot is a list of segments (group, first node tile, tiles); heavy groups
// -- one equation row each -- keep their slots for the whole launch, so an SM
// executes ONE body, tile after tile and launch after launch, out of its
// instruction cache (a block takes the slot of its %smid; the grid kernel
// above streams 25-55 KB of once-used code per block through the GPC-level
// instruction caches and spends > 60 % of its stall samples waiting for
// instructions, profiles/).  Everything else a body needs from L2 is taken off
// its critical path as well:
//   * the input rows of the NEXT item (the contiguous row window the body reads,
//     opty_group_xrow0/xrows) are fetched by bulk copies
//     (cp.async.bulk.shared.global, one row of OPTY_THREADS + 2 columns per
//     lane of warp 0, mbarrier completion) into the second input buffer while
//     the block works on the current item;
//   * a warp's Jacobian tile stores run on across items: staging buffers
//     alternate (an item with an odd number of phases swaps them), a phase only
//     waits until the store before the previous one has read its buffer.
// All warps of a block run the same body on adjacent node tiles; one
// __syncthreads per item hands the consumed input buffer back to the prefetch.
#define OPTY_PXSEG_BYTES (((OPTY_XROWS_MAX * OPTY_XBOX * 8) + 127) / 128 * 128)
#define OPTY_XBUF_BYTES (OPTY_NSEG * OPTY_PXSEG_BYTES)
#undef OPTY_SMEM_XIN_BYTES
#define OPTY_SMEM_XIN_BYTES (2 * OPTY_XBUF_BYTES)

// one lane: fetch the input window (OPTY_XROWS_MAX rows from row0) of node tile t into `dst` -- one 2-D TMA
// tile load per segment of OPTY_XSEG nodes (tm.in: box {OPTY_XBOX, OPTY_XROWS_MAX}; rows and columns beyond
// the matrix arrive as zeros).  (One bulk copy per row and lane was tried first: the 32 copies of a warp
// are issued one after the other and kept warp 0 a microsecond behind the others in every item.)
static __device__ __forceinline__ void opty_issue_input_rows(const CUtensorMap* map, int row0, int t,
                                                             unsigned char* dst, uint64_t* bar) {
  opty_mbar_expect_tx(bar, OPTY_NSEG * OPTY_XROWS_MAX * OPTY_XBOX * 8);
  for (int sgm = 0; sgm < OPTY_NSEG; ++sgm)
    opty_tma_load_2d(dst + sgm * OPTY_PXSEG_BYTES, map, t * OPTY_THREADS + sgm * OPTY_XSEG, row0, bar);
}
// (the window of a segment's group is packed into the segment entry: .w = first row | rows << 16)
#define opty_issue_input(p, sg, t, dst, bar) \
  if ((threadIdx.x & 31) == 0) opty_issue_input_rows(&tm.in, (sg).w & 0xffff, t, dst, bar)

// Phase 0 (OPTY_FUSED_PRE): the derived rows and the residuals of the constant rows -- the work of the
// pre-pass kernel -- are computed by the blocks of this launch: block = one case of the pre-pass (its code
// stays small) x one chunk of the nodes.  `p.ready[0]` counts the (node, case) pairs done over all launches
// of the handle (64-bit, never reset); a block fetches its first input window once the count has reached
// this launch's target.  All blocks are resident (one per SM) and produce before they consume.
#ifndef OPTY_FUSED_PRE
#define OPTY_FUSED_PRE 0
#endif
// the pre-pass kernel writes the constant runs of the first OPTY_CONST_PRE_PCT per cent of the nodes: the
// memory system has nothing else to do while it runs
#ifndef OPTY_CONST_PRE_PCT
#define OPTY_CONST_PRE_PCT 0
#endif
#define OPTY_CONST_PRE_NODES(n) ((int)(((long long)(n) * OPTY_CONST_PRE_PCT) / 100))
#if OPTY_FUSED_PRE
static __device__ __forceinline__ unsigned long long opty_ld_acquire(const unsigned long long* a) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(a) : "memory");
  return v;
}
// warp 0, before it fetches its first input window: every (node, case) pair of this launch is done
static __device__ __forceinline__ void opty_wait_ready(const OptyParams& p, int& all_ready) {
  if (!all_ready) {
    if ((threadIdx.x & 31) == 0) {
      const unsigned long long need_ = (unsigned long long)p.epoch * OPTY_PRE_GROUPS * (unsigned long long)p.n_nodes;
      while (opty_ld_acquire(p.ready) < need_) {
      }
    }
    __syncwarp();
    all_ready = 1;
    // the rows were written through the generic proxy (by other SMs), the bulk copies read through the async proxy
    asm volatile("fence.proxy.async.global;" ::: "memory");
  }
}
// a thread takes up to OPTY_PRE_ILP nodes at a time: their (long, serial) sine / cosine chains interleave
#define OPTY_PRE_ILP 3
#define OPTY_PRE_PHASE()                                                                                   \
  {                                                                                                        \
    const int pg_ = opty_slot % OPTY_PRE_GROUPS, ci_ = opty_slot / OPTY_PRE_GROUPS;                        \
    const int nc_ = (OPTY_NSLOTS - pg_ + OPTY_PRE_GROUPS - 1) / OPTY_PRE_GROUPS;                           \
    const int per_ = ((p.n_nodes + nc_ - 1) / nc_ + 31) & ~31;                                             \
    const int b0_ = ci_ * per_, b1_ = min(p.n_nodes, b0_ + per_);                                          \
    for (int base_ = b0_ + (int)threadIdx.x; base_ < b1_; base_ += OPTY_PRE_ILP * OPTY_THREADS) {          \
      int nodes_[OPTY_PRE_ILP]; /* (a repeated node writes the same values again) */                        \
      _Pragma("unroll") for (int u_ = 0; u_ < OPTY_PRE_ILP; ++u_)                                          \
        nodes_[u_] = min(base_ + u_ * OPTY_THREADS, b1_ - 1);                                              \
      opty_pre_case(p, nodes_, pg_);                                                                       \
    }                                                                                                      \
    __threadfence();                                                                                       \
    __syncthreads();                                                                                       \
    if (threadIdx.x == 0 && b1_ > b0_) atomicAdd(p.ready, (unsigned long long)(b1_ - b0_));                \
  }
#define OPTY_WAIT_READY(t) opty_wait_ready(p, opty_all_ready);
#else
#define OPTY_PRE_PHASE()
// launched behind the pre-pass with programmatic stream serialisation: its results are visible from here
// (no effect after an ordinary launch)
#define OPTY_WAIT_READY(t)                                     \
  if (!opty_all_ready) {                                       \
    asm volatile("griddepcontrol.wait;" ::: "memory");         \
    opty_all_ready = 1;                                        \
  }
#endif

// measurement aid (debug_nostore bit 1): thread 0 prints when its block started, how long every item took
#if OPTY_DEBUG_NOSTORE & 2
static __device__ __forceinline__ unsigned long long opty_gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define OPTY_TIMING_BEGIN()                     \
  unsigned long long opty_t0 = opty_gtime();    \
  unsigned long long opty_tl = opty_t0;         \
  unsigned opty_dt[12];                         \
  int opty_gi[12];                              \
  for (int i_ = 0; i_ < 12; ++i_) { opty_dt[i_] = 0; opty_gi[i_] = -1; } \
  unsigned opty_st[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#define OPTY_TIMING_STAMP(k) \
  if ((k) < 8) opty_st[(k)] = (unsigned)(opty_gtime() - opty_t0);
#define OPTY_TIMING_ITEM()                                           \
  if (threadIdx.x == 0 && opty_it < 12) {                            \
    const unsigned long long n_ = opty_gtime();                      \
    opty_dt[opty_it] = (unsigned)(n_ - opty_tl);                     \
    opty_gi[opty_it] = opty_g;                                       \
    opty_tl = n_;                                                    \
  }
#define OPTY_TIMING_END()                                                                                         \
  if (threadIdx.x == 0)                                                                                           \
    printf("T slot %d t0 %llu total %llu items %u : %d %u | %d %u | %d %u | %d %u | %d %u | %d %u | %d %u | %d %u | %d %u | %d %u\n", \
           opty_slot, opty_t0, opty_gtime() - opty_t0, opty_it, opty_gi[0], opty_dt[0], opty_gi[1], opty_dt[1],      \
           opty_gi[2], opty_dt[2], opty_gi[3], opty_dt[3], opty_gi[4], opty_dt[4], opty_gi[5], opty_dt[5], opty_gi[6],    \
           opty_dt[6], opty_gi[7], opty_dt[7], opty_gi[8], opty_dt[8], opty_gi[9], opty_dt[9]);                          \
  if ((threadIdx.x & 31) == 0)                                                                                    \
    printf("W slot %d warp %d stamps %u %u %u %u %u %u %u %u\n", opty_slot, threadIdx.x >> 5, opty_st[0], opty_st[1], \
           opty_st[2], opty_st[3], opty_st[4], opty_st[5], opty_st[6], opty_st[7]);
#else
#define OPTY_TIMING_STAMP(k)
#define OPTY_TIMING_BEGIN()
#define OPTY_TIMING_ITEM()
#define OPTY_TIMING_END()
#endif

// constant rows: the node-invariant column runs (values p.cvals, filled by the invariants kernel) are the
// same bytes for every node.  The block keeps one copy in shared memory; every lane of the launch owns a few
// nodes and sends each run to its node rows with one bulk copy (cp.async.bulk.global.shared::cta) -- the
// copies drain in the background while the warps work on their items.  (Plain 16-byte stores were tried
// first: 41 MB of them at the start of the launch stall every warp for 7-13 us on the full store queues,
// profiles/r02s_*.)
#if OPTY_NCRUNS > 0
#define OPTY_SMEM_CONST_BYTES ((OPTY_NCONST * 8 + 127) / 128 * 128)
#define OPTY_CONST_FILL()                                                                                      \
  double* opty_cbuf = reinterpret_cast<double*>(opty_smem + OPTY_SMEM_TILES_BYTES + OPTY_SMEM_XIN_BYTES + 128); \
  for (int i_ = threadIdx.x; i_ < OPTY_NCONST; i_ += OPTY_THREADS) opty_cbuf[i_] = p.cvals[i_];                \
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
// The tile stores of the items and these copies share the SM's one in-order TMA queue, and a warp has ONE
// staging buffer: a tile store that queues behind a backlog of copies stalls the next item until it has been
// read.  The memory system on the other hand is the bottleneck of the whole launch (81 MB at ~4.5 TB/s) and
// must not idle while the blocks wait for the pre-pass and work on their first items.  So a warp sends its
// nodes in slices: OPTY_CONST_HEAD_PCT per cent when the block starts, OPTY_CONST_FIRST_PCT when its first
// input window has arrived, OPTY_CONST_ITEM_PCT at the start of every later item (behind the tile store of
// the item before), the rest after its last item.  Lanes 1..31 only while items follow: lane 0 waits on its
// own store groups at every phase.
#ifndef OPTY_CONST_HEAD_PCT
#define OPTY_CONST_HEAD_PCT 25
#endif
#ifndef OPTY_CONST_FIRST_PCT
#define OPTY_CONST_FIRST_PCT 35
#endif
#ifndef OPTY_CONST_ITEM_PCT
#define OPTY_CONST_ITEM_PCT 15
#endif
// Nodes from OPTY_CONST_ALIGN_PCT per cent of the way on are not sent in slices but together with the tiles: the
// block that has just stored its row's tile of a node tile sends the constant runs of every OPTY_NGROUPS-th
// node of that tile, so that the two halves of these node rows reach memory at the same time (a DRAM page
// written in two visits costs two activations; the early slices trade that for the otherwise idle start of
// the launch).
#ifndef OPTY_CONST_ALIGN_PCT
#define OPTY_CONST_ALIGN_PCT 100
#endif
#define OPTY_CTX_CONST()                                                \
  ctx.cn = &opty_cn;                                                    \
  ctx.ce = opty_ce;                                                     \
  ctx.ctick = (opty_cnpw * OPTY_CONST_TICK_PCT + 99) / 100;             \
  ctx.cbuf = opty_cbuf;
#define OPTY_CONST_INIT()                                                                              \
  const int opty_cpre = OPTY_CONST_PRE_NODES(p.n_nodes); /* written by the pre-pass kernel */          \
  const int opty_csplit = OPTY_CONST_ALIGN_PCT >= 100                                                  \
                              ? p.n_nodes                                                              \
                              : max(opty_cpre, (int)((long long)p.n_nodes * OPTY_CONST_ALIGN_PCT / 100) / OPTY_THREADS * OPTY_THREADS); \
  const int opty_cnpw = (opty_csplit - opty_cpre + OPTY_NSLOTS * OPTY_WARPS - 1) / (OPTY_NSLOTS * OPTY_WARPS); \
  int opty_cn = opty_cpre + (opty_slot * OPTY_WARPS + (int)(threadIdx.x >> 5)) * opty_cnpw;            \
  const int opty_ce = min(opty_csplit, opty_cn + opty_cnpw);
// inside a body (the emitter places a few of these per body): the next ctx.ctick nodes of the early share --
// a trickle that never piles up in front of a tile store
#ifndef OPTY_CONST_TICK_PCT
#define OPTY_CONST_TICK_PCT 0
#endif
#define OPTY_TICK()                                                                                    \
  do {                                                                                                 \
    if ((OPTY_DEBUG_NOSTORE & 1) == 0 && OPTY_CONST_TICK_PCT > 0 && *ctx.cn < ctx.ce) {                \
      const int n1_ = min(ctx.ce, *ctx.cn + ctx.ctick);                                                \
      if (ctx.lane >= 1)                                                                               \
        for (int n_ = *ctx.cn + ctx.lane - 1; n_ < n1_; n_ += 31)                                      \
          for (int r_ = 0; r_ < OPTY_NCRUNS; ++r_)                                                     \
            opty_bulk_store_1d(ctx.jac + (long long)n_ * OPTY_K + opty_crun[r_][0], ctx.cbuf + opty_crun[r_][2], \
                               (uint32_t)opty_crun[r_][1] * 16u);                                      \
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");                                        \
      *ctx.cn = n1_;                                                                                   \
    }                                                                                                  \
  } while (0)
// after the tile store of item (opty_g, tile): lane l sends node ctx.node + l - 1 .. (lane 0 sends nothing: it
// waits on its own store groups), lane 1 also the warp's last node
#define OPTY_CONST_ALIGNED()                                                                           \
  if ((OPTY_DEBUG_NOSTORE & 1) == 0 && OPTY_CONST_ALIGN_PCT < 100 && tile_node0 >= opty_csplit) {      \
    const int l_ = threadIdx.x & 31;                                                                   \
    for (int q_ = 0; q_ < 2; ++q_) {                                                                   \
      const int n_ = ctx.node + (q_ == 0 ? l_ - 1 : 31);                                               \
      if (l_ >= 1 && (q_ == 0 || l_ == 1) && n_ < p.n_nodes && (n_ % OPTY_NGROUPS) == opty_g)          \
        for (int r_ = 0; r_ < OPTY_NCRUNS; ++r_)                                                       \
          opty_bulk_store_1d(p.jac + (long long)n_ * OPTY_K + opty_crun[r_][0], opty_cbuf + opty_crun[r_][2], \
                             (uint32_t)opty_crun[r_][1] * 16u);                                        \
    }                                                                                                  \
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");                                          \
  }
// the next `pct` per cent of the warp's nodes (everything that is left if `all`)
#define OPTY_CONST_SLICE(pct, all)                                                                     \
  if ((OPTY_DEBUG_NOSTORE & 1) == 0 && opty_cn < opty_ce) {                                            \
    const int n1_ = (all) ? opty_ce : min(opty_ce, opty_cn + (opty_cnpw * (pct) + 99) / 100);          \
    const int l_ = (threadIdx.x & 31) - ((all) ? 0 : 1), nl_ = (all) ? 32 : 31;                        \
    if (l_ >= 0)                                                                                       \
      for (int n_ = opty_cn + l_; n_ < n1_; n_ += nl_)                                                 \
        for (int r_ = 0; r_ < OPTY_NCRUNS; ++r_)                                                       \
          opty_bulk_store_1d(p.jac + (long long)n_ * OPTY_K + opty_crun[r_][0], opty_cbuf + opty_crun[r_][2], \
                             (uint32_t)opty_crun[r_][1] * 16u);                                        \
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");                                          \
    opty_cn = n1_;                                                                                     \
  }
#else
#define OPTY_SMEM_CONST_BYTES 0
#define OPTY_CONST_FILL()
#define OPTY_CONST_INIT()
#define OPTY_CONST_SLICE(pct, all)
#define OPTY_CONST_ALIGNED()
#define OPTY_TICK()
#define OPTY_CTX_CONST()
#endif

#define OPTY_KERNEL_BEGIN()                                                                              \
  OPTY_SMEM_SETUP()                                                                                      \
  /* slot = block: with one block per SM and an otherwise idle device the hardware deals the blocks of   \
     every launch out to the SMs in the same order, which is what keeps a body in the same instruction   \
     caches from launch to launch; nothing depends on it but speed */                                     \
  const int opty_slot = blockIdx.x;                                                                      \
  if (threadIdx.x == 0) {                                                                                \
    opty_mbar_init(&bar[0], 1);                                                                          \
    opty_mbar_init(&bar[1], 1);                                                                          \
  }                                                                                                      \
  OPTY_CONST_FILL()                                                                                      \
  __syncthreads();                                                                                       \
  const int opty_seg_end = opty_sched_slot[opty_slot + 1];                                               \
  int opty_seg = opty_sched_slot[opty_slot];                                                             \
  int4 opty_sg = opty_sched_seg[opty_seg < opty_seg_end ? opty_seg : 0];                                 \
  int opty_k = 0;                                                                                        \
  unsigned opty_it = 0;                                                                                  \
  int opty_flip = 0;                                                                                     \
  int opty_all_ready = 0;                                                                                \
  (void)opty_all_ready;                                                                                  \
  OPTY_TIMING_BEGIN()                                                                                    \
  OPTY_PRE_PHASE()                                                                                       \
  /* the first part of the constant runs leaves while the block waits for the pre-pass */                \
  OPTY_CONST_INIT()                                                                                      \
  OPTY_CONST_SLICE(OPTY_CONST_HEAD_PCT, false)                                                           \
  if (opty_seg < opty_seg_end && threadIdx.x < 32) {                                                     \
    OPTY_WAIT_READY(opty_sg.y)                                                                           \
    opty_issue_input(p, opty_sg, opty_sg.y, xin_bytes, &bar[0]);                                         \
  }                                                                                                      \
  OPTY_TIMING_STAMP(0)                                                                                   \
  while (opty_seg < opty_seg_end) {                                                                      \
    const int opty_g = opty_sg.x;                                                                        \
    const int tile_node0 = (opty_sg.y + opty_k) * OPTY_THREADS;                                          \
    int opty_nseg = opty_seg, opty_nk = opty_k + 1;                                                      \
    int4 opty_nsg = opty_sg;                                                                             \
    if (opty_nk >= opty_sg.z) {                                                                          \
      ++opty_nseg;                                                                                       \
      opty_nk = 0;                                                                                       \
      if (opty_nseg < opty_seg_end) opty_nsg = opty_sched_seg[opty_nseg];                                \
    }                                                                                                    \
    /* the other buffer was handed back by the __syncthreads that ended the previous item */             \
    if (opty_nseg < opty_seg_end && threadIdx.x < 32) {                                                  \
      OPTY_WAIT_READY(opty_nsg.y + opty_nk)                                                              \
      opty_issue_input(p, opty_nsg, opty_nsg.y + opty_nk, xin_bytes + ((opty_it + 1) & 1) * OPTY_XBUF_BYTES, \
                       &bar[(opty_it + 1) & 1]);                                                         \
    }                                                                                                    \
    opty_mbar_wait(&bar[opty_it & 1], (opty_it >> 1) & 1);                                               \
    if (opty_it == 0) {                                                                                  \
      OPTY_CONST_SLICE(OPTY_CONST_FIRST_PCT, false)                                                      \
    } else {                                                                                             \
      OPTY_CONST_SLICE(OPTY_CONST_ITEM_PCT, false)                                                       \
    }                                                                                                    \
    OPTY_TIMING_STAMP(1 + 2 * opty_it)                                                                   \
    OptyCtx ctx;                                                                                         \
    ctx.lane = threadIdx.x & 31;                                                                         \
    ctx.n_nodes = p.n_nodes;                                                                             \
    ctx.ldt = p.ldt;                                                                                     \
    ctx.xp = reinterpret_cast<const double*>(xin_bytes + (opty_it & 1) * OPTY_XBUF_BYTES +               \
                                             (threadIdx.x / OPTY_XSEG) * OPTY_PXSEG_BYTES) +             \
             (threadIdx.x % OPTY_XSEG);                                                                  \
    ctx.xs = opty_smem_u32(ctx.xp);                                                                      \
    ctx.ldc = p.ldc;                                                                                     \
    ctx.tile0 = tiles + ((threadIdx.x >> 5) * OPTY_NBUF + (OPTY_NBUF == 2 ? opty_flip : 0)) * OPTY_TILE_DOUBLES;     \
    ctx.tile1 = tiles + ((threadIdx.x >> 5) * OPTY_NBUF + (OPTY_NBUF == 2 ? 1 - opty_flip : 0)) * OPTY_TILE_DOUBLES; \
    ctx.jac = p.jac;                                                                                     \
    ctx.tm = &tm;                                                                                        \
    ctx.node = tile_node0 + (threadIdx.x & ~31);                                                         \
    ctx.active = (tile_node0 + (int)threadIdx.x) < p.n_nodes;                                            \
    ctx.con = p.con + tile_node0 + threadIdx.x;                                                          \
    OPTY_CTX_CONST()                                                                                     \
    if (ctx.node < p.n_nodes) {

#define OPTY_KERNEL_END()                                                                              \
      OPTY_CONST_ALIGNED()                                                                             \
    }                                                                                                  \
    OPTY_TIMING_STAMP(2 + 2 * opty_it)                                                                 \
    opty_flip ^= opty_group_odd[opty_g];                                                               \
    __syncthreads();                                                                                   \
    OPTY_TIMING_ITEM()                                                                                 \
    opty_seg = opty_nseg;                                                                              \
    opty_k = opty_nk;                                                                                  \
    opty_sg = opty_nsg;                                                                                \
    ++opty_it;                                                                                         \
  }                                                                                                    \
  OPTY_CONST_SLICE(100, true)                                                                                \
  /* shared memory must outlive the tile stores that still read it */                                  \
  if (OPTY_TMA_STORE) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");                   \
  OPTY_TIMING_END()

#else
// Persistent, code-stationary kernel: grid = (resident blocks per SM) x SMs.
// Every warp of the generated code walks its straight-line body exactly once
// per node tile, so a block that changes its group with every tile streams
// its instructions through the GPC-level cache and stalls on instruction
// fetch for more than half of its cycles (profiles/).  Here a block keeps ONE
// group for as long as that group has tiles left -- the body stays in the SM's
// instruction cache -- and all blocks of an SM start on the same group
// (`opty_sm_group`, indexed by %smid: SMs are dealt out to the groups in
// proportion to their cost).  Tiles are handed out by one atomic counter per
// group (`p.work`), asked for one tile ahead so that the round trip hides
// behind the arithmetic; a block whose group has run dry moves to the group
// with the most work left, which balances the load without a cost model that
// has to be right.  The last block to leave resets the counters for the next
// launch.
// next (slot, tile) for a block whose current slot has run dry: the slot with
// the most work left (one look at all counters, then one atomic)
static __device__ __noinline__ int2 opty_steal(const OptyParams& p, const int* opty_slot_cost) {
  for (int tries = 0; tries < OPTY_NGROUPS; ++tries) {
    int best = -1;
    long long best_left = 0;
    for (int s_ = 0; s_ < OPTY_NGROUPS; ++s_) {
      const int done = *reinterpret_cast<volatile int*>(p.work + s_);
      const long long left = (long long)(p.n_tiles - done) * opty_slot_cost[s_];
      if (left > best_left) {
        best_left = left;
        best = s_;
      }
    }
    if (best < 0) break;
    const int t_ = atomicAdd(p.work + best, 1);
    if (t_ < p.n_tiles) return make_int2(best, t_);
  }
  return make_int2(-1, 0);
}

#define OPTY_KERNEL_BEGIN()                                                                            \
  OPTY_SMEM_SETUP()                                                                                    \
  __shared__ int opty_next[2];                                                                         \
  OPTY_STAGE_INIT()                                                                                    \
  unsigned opty_smid;                                                                                  \
  asm("mov.u32 %0, %%smid;" : "=r"(opty_smid));                                                        \
  if (threadIdx.x == 0) {                                                                              \
    const int s0_ = opty_sm_group[opty_smid % OPTY_SM_TABLE];                                          \
    int t0_ = atomicAdd(p.work + s0_, 1);                                                              \
    int2 a_ = t0_ < p.n_tiles ? make_int2(s0_, t0_) : opty_steal(p, opty_slot_cost);                                   \
    opty_next[0] = a_.x;                                                                               \
    opty_next[1] = a_.y;                                                                               \
  }                                                                                                    \
  __syncthreads();                                                                                     \
  int opty_slot = opty_next[0], opty_t = opty_next[1];                                                 \
  while (opty_slot >= 0) {                                                                             \
    __syncthreads(); /* everybody has read opty_next and is done with the previous input slice */      \
    /* the next tile of this slot is asked for now and looked at after the body */                     \
    int opty_raw = 0;                                                                                  \
    if (threadIdx.x == 0) opty_raw = atomicAdd(p.work + opty_slot, 1);                                 \
    const int opty_g = opty_group_order[opty_slot];                                                    \
    const int tile_node0 = opty_t * OPTY_THREADS;                                                      \
    OPTY_STAGE_INPUT()                                                                                 \
    OPTY_CTX_SETUP()                                                                                   \
    if (ctx.node < p.n_nodes) {

#define OPTY_KERNEL_END()                                                                              \
    }                                                                                                  \
    if (threadIdx.x == 0) {                                                                            \
      int2 a_ = opty_raw < p.n_tiles ? make_int2(opty_slot, opty_raw) : opty_steal(p, opty_slot_cost);                 \
      opty_next[0] = a_.x;                                                                             \
      opty_next[1] = a_.y;                                                                             \
    }                                                                                                  \
    __syncthreads();                                                                                   \
    opty_slot = opty_next[0];                                                                          \
    opty_t = opty_next[1];                                                                             \
  }                                                                                                    \
  if (threadIdx.x == 0) {                                                                              \
    __threadfence();                                                                                   \
    if (atomicInc(reinterpret_cast<unsigned*>(p.work + OPTY_NGROUPS), gridDim.x - 1) == gridDim.x - 1) \
      for (int g_ = 0; g_ < OPTY_NGROUPS; ++g_) p.work[g_] = 0;                                        \
  }
#endif

// ---------------------------------------------------------------------------
// pre-pass kernel: one thread per node, derived rows written coalesced
// ---------------------------------------------------------------------------
// With direct input loads (OPTY_TMA_LOAD == 2) the derived rows go to the
// tiled layout, and blockIdx.y >= OPTY_PRE_GROUPS copies chunks of
// OPTY_COPY_ROWS trajectory rows there: thread = column, its value goes to its
// own tile and, for the first two columns of a tile, also to the halo slots of
// the tile before.
#define OPTY_PRE_THREADS 128
// last statement of the pre-pass kernel.  The kernel may have been launched with programmatic stream
// serialisation directly behind the main kernel of the previous evaluation and have started before that
// kernel's stores were flushed: it reads and writes nothing that kernel touches, but it must not COMPLETE
// before it -- the next main kernel waits for this grid only and writes output sets whose previous contents
// have to be final by then.  (No effect after an ordinary launch.)
#define OPTY_PRE_END() asm volatile("griddepcontrol.wait;" ::: "memory");
#define OPTY_COPY_ROWS 16
#define GA(r) __ldcg(xg + (long long)(r) * p.ldt)
#define GB(r) __ldcg(xg + (long long)(r) * p.ldt + 1)
#if OPTY_TMA_LOAD == 2
#define OPTY_DRV(d, val) drv[(d) * OPTY_TW] = (val)
#define OPTY_PCON(j, val) p.con[(long long)(j) * p.ldc + node] = (val)
#define OPTY_PRE_BEGIN()                                                                              \
  const int node = blockIdx.x * OPTY_PRE_THREADS + threadIdx.x;                                       \
  const int opty_pg = blockIdx.y;                                                                     \
  if (opty_pg >= OPTY_PRE_GROUPS) {                                                                   \
    if (node < p.n_cols) {                                                                            \
      const int t_ = node / OPTY_THREADS, j_ = node % OPTY_THREADS;                                   \
      const int r0_ = (opty_pg - OPTY_PRE_GROUPS) * OPTY_COPY_ROWS;                                   \
      double* own_ = p.tiled + (long long)t_ * (OPTY_RD * OPTY_TW) + j_;                              \
      for (int r_ = r0_; r_ < min(r0_ + OPTY_COPY_ROWS, OPTY_R); ++r_) {                              \
        const double v_ = __ldcg(p.traj + (long long)r_ * p.ldt + node);                              \
        own_[r_ * OPTY_TW] = v_;                                                                      \
        if (j_ < 2 && t_ > 0) own_[r_ * OPTY_TW - (OPTY_RD * OPTY_TW) + OPTY_THREADS] = v_;           \
      }                                                                                               \
    }                                                                                                 \
    return;                                                                                           \
  }                                                                                                   \
  if (node >= p.n_nodes) return;                                                                      \
  const double* xg = p.traj + node;                                                                   \
  double* drv = p.tiled + (long long)(node / OPTY_THREADS) * (OPTY_RD * OPTY_TW) +                    \
                (long long)OPTY_R * OPTY_TW + node % OPTY_THREADS;
#else
#define OPTY_DRV(d, val) drv[(long long)(d) * p.ldt] = (val)
#define OPTY_PCON(j, val) p.con[(long long)(j) * p.ldc + node] = (val)
#if OPTY_PERSISTENT == 2 && OPTY_CONST_PRE_PCT > 0 && OPTY_NCRUNS > 0
// last OPTY_PRE_CONST_SLICES slices of the grid: lanes = consecutive 16-byte pieces of a run, a warp = 4 nodes
// (many warps with few stores each: a warp's stores leave one after the other)
#define OPTY_PRE_CONST_SLICES 8
#define OPTY_PRE_CONST_SLICE()                                                                          \
  if (blockIdx.y >= OPTY_PRE_GROUPS - OPTY_PRE_CONST_SLICES) {                                          \
    const int n0_ = blockIdx.x * OPTY_PRE_THREADS + (int)(threadIdx.x & ~31u) +                         \
                    (int)(blockIdx.y - (OPTY_PRE_GROUPS - OPTY_PRE_CONST_SLICES)) * (32 / OPTY_PRE_CONST_SLICES); \
    const int n1_ = min(OPTY_CONST_PRE_NODES(p.n_nodes), n0_ + 32 / OPTY_PRE_CONST_SLICES);             \
    for (int r_ = 0; r_ < OPTY_NCRUNS; ++r_) {                                                          \
      const double2* v_ = reinterpret_cast<const double2*>(p.cvals + opty_crun[r_][2]);                 \
      for (int k_ = threadIdx.x & 31; k_ < opty_crun[r_][1]; k_ += 32) {                                \
        const double2 v2_ = v_[k_];                                                                     \
        double2* d_ = reinterpret_cast<double2*>(p.jac + (long long)n0_ * OPTY_K + opty_crun[r_][0]) + k_; \
        for (int n_ = n0_; n_ < n1_; ++n_, d_ += OPTY_K / 2) __stcs(d_, v2_);                           \
      }                                                                                                 \
    }                                                                                                   \
    return;                                                                                             \
  }
#else
#define OPTY_PRE_CONST_SLICE()
#endif
#define OPTY_PRE_BEGIN()                                        \
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); \
  OPTY_PRE_CONST_SLICE()                                        \
  const int node = blockIdx.x * OPTY_PRE_THREADS + threadIdx.x; \
  if (node >= p.n_nodes) return;                                \
  const double* xg = p.traj + node;                             \
  double* drv = p.traj + (long long)OPTY_R * p.ldt + node;      \
  const int opty_pg = blockIdx.y;
#endif
