// TEST INFRASTRUCTURE ONLY -- never used by the product path.
//
// Host stand-in for opty_b200/csrc/colloc_kernel.cuh: lets the CPU test-suite
// compile an emitted module with g++ and run its pre-pass and group bodies
// node by node, so that the SymPy -> tape -> CUDA-C emitter (lowering,
// differentiation, derived rows, chunk / column bookkeeping) can be checked
// against the oracle without a GPU.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#define __global__
#define __device__
#define __forceinline__ inline
#define __restrict__
#define __grid_constant__
#define __launch_bounds__(...)

struct OptyTmaps { int unused; };
struct OptyParams {
  double* traj; double* tiled; double* con; double* jac;
  long long ldt; long long ldc; int n_nodes; int n_cols; int n_tiles; int* work;
};
struct Dim3 { unsigned x, y, z; };
static Dim3 blockIdx, threadIdx;

static double opty_ci[OPTY_NINV];
#define CI(k) opty_ci[k]

struct OptyCtx {
  const double* xs; long long ldt; double* con; long long ldc; double* jac;
  int node; double tile[OPTY_NBUF][OPTY_TILE_DOUBLES / 32];
};

static inline double opty_sign(double x) { return (double)((x > 0.0) - (x < 0.0)); }

#define XA(r) ctx.xs[(long long)(r) * ctx.ldt]
#define XB(r) ctx.xs[(long long)(r) * ctx.ldt + 1]
#define XD(d) ctx.xs[(long long)(OPTY_R + (d)) * ctx.ldt]
#define OPTY_CON(j, val) ctx.con[(long long)(j) * ctx.ldc] = (val)
// one node per "warp": a sub-tile that starts `off` doubles into the 32-lane buffer starts off/32 here
#define OPTY_JS2(buf, off, w, c, v0, v1) do { double* t_ = const_cast<OptyCtx&>(ctx).tile[buf] + (off) / 32 + (c); t_[0] = (v0); t_[1] = (v1); } while (0)
#define OPTY_JS1(buf, off, w, c, v0) const_cast<OptyCtx&>(ctx).tile[buf][(off) / 32 + (c)] = (v0)
#define OPTY_PHASE_BEGIN(t) do { } while (0)
#define OPTY_FENCE() do { } while (0)
#define OPTY_FLUSH_BEGIN()
#define OPTY_TSTORE(map, buf, off, w, col0) \
  memcpy(ctx.jac + (long long)ctx.node * OPTY_K + (col0), ctx.tile[buf] + (off) / 32, (w) * sizeof(double));
#define OPTY_FLUSH_END()
#define OPTY_DRAIN() do { } while (0)
#define OPTY_DRAIN_WRITES() do { } while (0)
#define OPTY_JG(col, val) ctx.jac[(long long)ctx.node * OPTY_K + (col)] = (val)
#define OPTY_THREADS 1
#define OPTY_PRE_THREADS 1
#define OPTY_KERNEL_BEGIN() \
  OptyCtx ctx; ctx.node = (int)blockIdx.x; ctx.xs = p.traj + ctx.node; ctx.ldt = p.ldt; \
  ctx.con = p.con + ctx.node; ctx.ldc = p.ldc; ctx.jac = p.jac; (void)tm; \
  const int opty_g = opty_group_order[blockIdx.y];
#define OPTY_KERNEL_END()

#define GA(r) xg[(long long)(r) * p.ldt]
#define GB(r) xg[(long long)(r) * p.ldt + 1]
#define OPTY_PRE_END()
#define OPTY_DRV(d, val) drv[(long long)(d) * p.ldt] = (val)
#define OPTY_PRE_BEGIN() \
  const int node = (int)blockIdx.x; \
  const double* xg = p.traj + node; \
  double* drv = p.traj + (long long)OPTY_R * p.ldt + node; \
  const int opty_pg = (int)blockIdx.y;

extern "C" void opty_colloc_inv(const double* uni, double* inv);
extern "C" void host_get_invariants(double* out) { memcpy(out, opty_ci, sizeof(opty_ci)); }
extern "C" void opty_colloc_pre(const OptyParams p);
extern "C" void opty_colloc_eval(const OptyTmaps tm, const OptyParams p);

// traj must have OPTY_R + OPTY_D rows
extern "C" void host_eval(const double* uni, double* traj, long long ldt, int n_nodes,
                          double* con, double* jac) {
  threadIdx.x = threadIdx.y = 0; blockIdx.x = blockIdx.y = 0;
  opty_colloc_inv(uni, opty_ci);
  OptyTmaps tm; OptyParams p;
  p.traj = traj; p.tiled = nullptr; p.con = con; p.jac = jac; p.ldt = ldt; p.ldc = n_nodes; p.n_nodes = n_nodes; p.n_cols = n_nodes + 1;
  for (int pg = 0; pg < 300; ++pg)
    for (int i = 0; i < n_nodes; ++i) { blockIdx.x = i; blockIdx.y = pg; opty_colloc_pre(p); }
  for (int g = 0; g < OPTY_NGROUPS; ++g)
    for (int i = 0; i < n_nodes; ++i) { blockIdx.x = i; blockIdx.y = g; opty_colloc_eval(tm, p); }
}
