// Persistent variant of the collocation kernel skeleton (generated modules
// include it after colloc_kernel.cuh when the `persistent` option is set).
//
// Why: the straight-line group bodies are bound by instruction delivery -- every
// (block, body) pair of the grid kernel streams its body once from the GPC-level
// instruction cache.  Here one block per SM is bound to ONE
// output group for the whole launch and its warps loop over node tiles, so an
// SM fetches a single body (<= ~30 KB) once and then runs it out of its own
// instruction cache.  The block -> (group, tile range) schedule is a table in
// device memory, computed on the host from measured per-group tile times
// (`opty_colloc_set_schedule`, `opty_colloc_block_clocks`): no atomics, no work
// stealing.  The pre-pass (derived rows) runs as phase 0 of the same launch,
// separated from the group bodies by a grid-wide barrier (all blocks are
// co-resident: cooperative launch, one block per SM).
//
// Per warp: its own [R+D][34] input slice staged by a TMA tile load on its own
// mbarrier, its own Jacobian staging tiles.  The base skeleton is included with
// OPTY_WARPS = 1 (per-warp geometry); OPTY_PWARPS is the real block width.
#pragma once

struct OptyPersist {
  const int4* sched;        // per block: {group (local index), first tile, end tile, unused}; tiles of 32 nodes
  unsigned int* barrier;    // grid barrier counter (monotonic)
  unsigned int barrier_target;  // value the counter reaches when every block of THIS launch has arrived
  long long* block_clocks;  // per block: clock64 ticks spent in the group phase (schedule tuning)
  int pre_units;            // derived-row groups of the pre-pass (0: no phase 0)
};

#define OPTY_PTHREADS (OPTY_PWARPS * 32)

// Block-wide Jacobian tile stores.  The warps of a block work on adjacent
// node tiles of the same group in lock step, so their staging tiles form one
// [32*W nodes][C columns] block: ONE TMA store per chunk instead of W (the
// per-SM TMA unit spends ~250 cycles per operation besides ~21 B/clk, and with
// [32 x 30] tiles it is the busiest unit of the SM).  Buffer b of warp w sits
// at (b * W + w) * 32 * C doubles.  Warps whose tile index lies beyond the
// block's range compute the neighbouring block's tile again (same group, same
// values), so every row of the box is valid; rows beyond the last node are
// clipped by the tensor map.
#ifndef OPTY_PERSIST_BLOCK_STORES
#define OPTY_PERSIST_BLOCK_STORES 1
#endif
#if OPTY_PERSIST_BLOCK_STORES
#undef OPTY_TROW
#define OPTY_TROW(buf) (ctx.trow0 + (buf) * (OPTY_PWARPS * OPTY_TILE_DOUBLES))

template <int SEG, int Q, int BUF, int SEGCOL0, int NCOLS>
static __device__ __forceinline__ void opty_flush_block(const OptyCtx& ctx) {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0 && OPTY_DEBUG_NOSTORE != 1) {
    // ctx.tile0 of warp 0 is the block's buffer 0; ctx.node of warp 0 the block's first node
    opty_tma_store_2d(&ctx.tm->out[SEG], ctx.tile0 + BUF * (OPTY_PWARPS * OPTY_TILE_DOUBLES), Q * OPTY_C, ctx.node);
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(OPTY_NBUF - 1) : "memory");
  }
  __syncthreads();
}
#undef OPTY_FLUSH
#define OPTY_FLUSH(seg, q, buf, segcol0, ncols) opty_flush_block<seg, q, buf, segcol0, ncols>(ctx)
#undef OPTY_DRAIN
#define OPTY_DRAIN()                                                                          \
  do {                                                                                        \
    if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");      \
    __syncthreads();                                                                          \
  } while (0)
#define OPTY_PTILE0(warp) (tiles + (warp) * OPTY_TILE_DOUBLES)
#define OPTY_PROUND_SYNC() __syncthreads()
#else
// per-warp [32 x C] tile stores of the base skeleton (warps run independently)
#define OPTY_PTILE0(warp) (tiles + (warp) * OPTY_NBUF * OPTY_TILE_DOUBLES)
#define OPTY_PROUND_SYNC() __syncwarp()
#endif
#define OPTY_PSLICE_BYTES OPTY_XSEG_BYTES  // [R+D][34] doubles, 128-byte multiple
#define OPTY_PSMEM_TILES_BYTES (OPTY_PWARPS * OPTY_NBUF * OPTY_TILE_DOUBLES * 8)
#define OPTY_PSMEM_BYTES (OPTY_PSMEM_TILES_BYTES + OPTY_PWARPS * OPTY_PSLICE_BYTES + 8 * OPTY_PWARPS + 128)

static __device__ __forceinline__ void opty_grid_barrier(unsigned int* counter, unsigned int target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(counter, 1u);
    unsigned int seen;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory");
    } while ((int)(seen - target) < 0);
  }
  __syncthreads();
  // derived rows written by other SMs with ordinary stores are read below by
  // TMA (async proxy)
  asm volatile("fence.proxy.async;" ::: "memory");
}

// Phase 0: the blocks share the pre-pass work (node x derived-row group units).
#define OPTY_PERSIST_PRE()                                                                        \
  if (ps.pre_units > 0) {                                                                         \
    const long long units = (long long)p.n_nodes * ps.pre_units;                                  \
    for (long long u = (long long)blockIdx.x * OPTY_PTHREADS + threadIdx.x; u < units;            \
         u += (long long)gridDim.x * OPTY_PTHREADS) {                                             \
      const int pg = (int)(u / p.n_nodes);                                                        \
      const int node = (int)(u - (long long)pg * p.n_nodes);                                      \
      opty_pre_unit(p, node, pg);                                                                 \
    }                                                                                             \
    opty_grid_barrier(ps.barrier, ps.barrier_target);                                             \
  }

// Sets up the warp's context and loops over its tiles; the generated dispatch
// `switch (opty_g)` sits between OPTY_PERSIST_LOOP_BEGIN and _END.
#define OPTY_PERSIST_BEGIN()                                                                      \
  extern __shared__ __align__(128) unsigned char opty_smem[];                                     \
  const int opty_warp = threadIdx.x >> 5;                                                         \
  double* tiles = reinterpret_cast<double*>(opty_smem);                                           \
  unsigned char* slice = opty_smem + OPTY_PSMEM_TILES_BYTES + opty_warp * OPTY_PSLICE_BYTES;      \
  uint64_t* bar = reinterpret_cast<uint64_t*>(opty_smem + OPTY_PSMEM_TILES_BYTES +                \
                                              OPTY_PWARPS * OPTY_PSLICE_BYTES) + opty_warp;       \
  if ((threadIdx.x & 31) == 0) opty_mbar_init(bar, 1);                                            \
  __syncthreads();                                                                                \
  OPTY_PERSIST_PRE()                                                                              \
  const int4 opty_sch = ps.sched[blockIdx.x];                                                     \
  const int opty_g = opty_sch.x;                                                                  \
  uint32_t phase = 0;                                                                             \
  OptyCtx ctx;                                                                                    \
  ctx.lane = threadIdx.x & 31;                                                                    \
  ctx.n_nodes = p.n_nodes;                                                                        \
  ctx.ldt = p.ldt;                                                                                \
  ctx.xs = reinterpret_cast<const double*>(slice) + ctx.lane;                                     \
  ctx.ldc = p.ldc;                                                                                \
  ctx.tile0 = OPTY_PTILE0(opty_warp);                                                             \
  ctx.trow0 = ctx.tile0 + ctx.lane * OPTY_C;                                                      \
  ctx.jac = p.jac;                                                                                \
  ctx.tm = &tm;                                                                                   \
  const long long opty_t0 = clock64();

#define OPTY_PERSIST_LOOP_BEGIN()                                                                 \
  for (int opty_r0 = opty_sch.y; opty_r0 < opty_sch.z; opty_r0 += OPTY_PWARPS) {                  \
    /* all warps run every round (block-wide stores need the whole block) */                      \
    const int tile_node0 = (opty_r0 + opty_warp) * 32;                                            \
    if (ctx.lane == 0) {                                                                          \
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");                                \
      opty_mbar_expect_tx(bar, OPTY_RD * OPTY_XBOX * 8);                                          \
      opty_tma_load_2d(slice, &tm.in, tile_node0, 0, bar);                                        \
    }                                                                                             \
    opty_mbar_wait(bar, phase);                                                                   \
    phase ^= 1u;                                                                                  \
    ctx.node = tile_node0;                                                                        \
    ctx.active = (tile_node0 + ctx.lane) < p.n_nodes;                                             \
    ctx.con = p.con + tile_node0 + ctx.lane;

#define OPTY_PERSIST_LOOP_END()                                                                   \
    OPTY_PROUND_SYNC();                                                                           \
  }

#define OPTY_PERSIST_END()                                                                        \
  __syncthreads();                                                                                \
  if (threadIdx.x == 0 && ps.block_clocks) ps.block_clocks[blockIdx.x] = clock64() - opty_t0;
