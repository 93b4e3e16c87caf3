/*
 * opty_b200 C-ABI: the drop-in boundary of the B200 collocation-constraint
 * engine.
 *
 * Every entry point replaces one piece of the reference's generated-code
 * boundary (csu-hmc/opty @ 911d150c, paths relative to /root/reference):
 *
 *   - the per-node C function
 *         void eval_matrix(double matrix[K], double a0_, ..., double aL_)
 *     (template opty/utils.py:483-494) and the Cython node loop
 *         eval_matrix_loop(matrix, *args)
 *     (template opty/utils.py:500-529) that `ufuncify_matrix`
 *     (opty/utils.py:639-928) compiles and imports, and
 *   - the NumPy glue around it: `parse_free` (opty/utils.py:277-326),
 *     `_merge_fixed_free` (opty/direct_collocation.py:2891-2926), the slicing
 *     and result allocation in `constraints` (opty/direct_collocation.py:
 *     2382-2446) and `constraints_jacobian` (opty/direct_collocation.py:
 *     2816-2887), and the index loop of `jacobian_indices`
 *     (opty/direct_collocation.py:2628-2684).
 *
 * Plain pointers and sizes only.  All values are IEEE float64, all indices
 * int64.  Every function returns 0 on success and a negative code on failure;
 * `opty_colloc_last_error()` then describes the failure (thread local).
 *
 * A handle is NOT reentrant (the reference's persistent Jacobian buffer,
 * opty/direct_collocation.py:2814, makes its `jacobian()` non-reentrant too).
 */
#ifndef OPTY_B200_H
#define OPTY_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OPTY_B200_ABI_VERSION 6
#define OPTY_MAX_GROUPS 1024
#define OPTY_MAX_SEGMENTS 1024

#define OPTY_OK 0
#define OPTY_ERR_ARG -1     /* invalid argument / configuration */
#define OPTY_ERR_CUDA -2    /* CUDA runtime or driver error */
#define OPTY_ERR_STATE -3   /* call sequence error (e.g. known values not set) */

/* OPTY_ELEMENTWISE: the generic `ufuncify_matrix` operator (opty/utils.py:
 * 639-670): N evaluation points, n array arguments of length N, r scalar
 * ("const") arguments after them in `free`, an M x P output matrix per point
 * (returned through opty_colloc_jacobian), no residuals, no neighbour column. */
enum { OPTY_BACKWARD_EULER = 0, OPTY_MIDPOINT = 1, OPTY_ELEMENTWISE = 2 };

/* Problem + kernel geometry.  Symbols follow the reference's notation
 * (opty/direct_collocation.py:101-114). */
typedef struct opty_colloc_cfg {
  int32_t abi_version;      /* OPTY_B200_ABI_VERSION */
  int32_t device;           /* CUDA device ordinal */
  int32_t N;                /* collocation nodes of the whole problem */
  int32_t node_lo;          /* this handle evaluates constraint nodes */
  int32_t node_hi;          /*   [node_lo, node_hi) of the N-1 (a shard) */
  int32_t n;                /* states */
  int32_t q;                /* unknown input trajectories */
  int32_t k;                /* known input trajectories */
  int32_t r;                /* unknown parameters */
  int32_t s;                /* 1 if the node time interval is free, else 0 */
  int32_t pk;               /* known parameters */
  int32_t M;                /* equations of motion */
  int32_t P;                /* partials per equation (2n+q+r+s or 2n+2q+r+s) */
  int32_t method;           /* OPTY_BACKWARD_EULER / OPTY_MIDPOINT */
  int32_t num_inv;          /* entries of the node-invariant table */
  int32_t num_groups;       /* output groups */
  int32_t num_derived;      /* D: derived rows written by the pre-pass kernel */
  int32_t tile_cols;        /* C: columns of the Jacobian staging tile */
  int32_t warps_per_block;
  int32_t pre_groups;       /* grid.y of the pre-pass kernel (groups of derived rows) */
  int32_t tile_bufs;        /* staging tiles per warp (1..4) */
  int32_t tma_load;         /* input staging of the module: 1 TMA tile loads, 0 plain loads into shared
                               memory, 2 none (lanes read the trajectory matrix directly) */
  int32_t tma_store;        /* module was emitted with TMA Jacobian stores */
  int32_t out_ring;         /* number of device output sets to rotate (>=1) */
  int32_t con_tail;         /* extra host slots after the M*(N-1) residuals */
  int32_t jac_tail;         /* extra host slots after the (N-1)*M*P partials */
  int32_t prefetch_jac;     /* opty_colloc_constraints starts the Jacobian D2H speculatively */
  int32_t persistent;       /* != 0: module holds the persistent main kernel (1: block-wide [32*W x C] TMA
                               tile stores, 2: per-warp [32 x C] stores): one block per SM bound to one
                               group, launched along the schedule of opty_colloc_set_schedule, the
                               pre-pass as phase 0 of the same (cooperative) launch */
  int32_t num_segments;     /* store segments: column runs of the node block written by the group bodies */
  int32_t primary_segments; /* segments [0, primary_segments) belong to the module given to
                               opty_colloc_create; the rest to modules added with opty_colloc_add_module */
  int32_t const_image_doubles; /* total length of the constant column runs that the pre-pass kernel
                                  replicates into every node row (0: none); segments and constant
                                  runs together tile the M*P columns */
  int32_t seg_col0[OPTY_MAX_SEGMENTS];   /* first Jacobian column of segment s */
  int32_t seg_ncols[OPTY_MAX_SEGMENTS];  /* number of columns of segment s */
  double h;                 /* fixed node time interval (ignored when s=1) */
} opty_colloc_cfg;

typedef struct opty_colloc opty_colloc_t;

/* Loads the generated sm_100a module (`cubin`, produced by nvcc from the
 * emitter's CUDA-C) on `cfg->device`, allocates the device-resident
 * trajectory matrix, residual / Jacobian buffers and pinned host buffers.
 * Replaces importing the compiled Cython module (opty/utils.py:909-916). */
int opty_colloc_create(const opty_colloc_cfg* cfg, const void* cubin, size_t cubin_bytes,
                       opty_colloc_t** out);

int opty_colloc_destroy(opty_colloc_t* h);

/* Known input trajectories `traj` as [k][N] (row-major, full problem length N)
 * and known parameter values `params` [pk], in the collocator's symbol order.
 * Replaces the known-value half of `_merge_fixed_free`
 * (opty/direct_collocation.py:2911-2926).  Must be called once before the
 * first evaluation (also when k = pk = 0) and again whenever a known
 * trajectory changes. */
int opty_colloc_set_known(opty_colloc_t* h, const double* traj, const double* params);

/* Copies the free vector (length n*N + q*N + r + s, layout of
 * opty/direct_collocation.py:116-125) to the device.  `free_host` may be the
 * handle's own pinned buffer (see opty_colloc_host_buffers) to skip staging. */
int opty_colloc_upload_free(opty_colloc_t* h, const double* free_host);

/* Evaluates residuals and Jacobian partials of the handle's node range from
 * the device-resident free vector into the next device output set.  No host
 * traffic.  `sync` != 0 waits for completion. */
int opty_colloc_eval_device(opty_colloc_t* h, int sync);

/* `Problem.constraints(free)` (opty/direct_collocation.py:498-525): uploads
 * `free_host` (skipped if it is bit-identical to the vector already resident),
 * evaluates, and returns the M*(node_hi-node_lo) residuals eom-major in
 * `con_host` (NULL: leave them in the pinned residual buffer). */
int opty_colloc_constraints(opty_colloc_t* h, const double* free_host, double* con_host);

/* `Problem.jacobian(free)` (opty/direct_collocation.py:552-562): as above for
 * the (node_hi-node_lo)*M*P partials, node-major. */
int opty_colloc_jacobian(opty_colloc_t* h, const double* free_host, double* jac_host);

/* Pinned host buffers owned by the handle: the free-vector staging buffer,
 * the residual buffer (M*nodes + con_tail) and the Jacobian buffer
 * (nodes*M*P + jac_tail).  Valid until opty_colloc_destroy. */
int opty_colloc_host_buffers(opty_colloc_t* h, double** free_pinned, double** con_pinned,
                             double** jac_pinned);

/* Device pointers of the most recently written output set and of the
 * trajectory matrix, for on-device consumers (e.g. torch views, NCCL). */
int opty_colloc_device_buffers(opty_colloc_t* h, void** traj, int64_t* ldt, void** con, void** jac,
                               void** uni);

/* Restricts Jacobian device->host copies to the column ranges
 * [col_begin[i], col_end[i]) of every node block.  Columns outside the ranges
 * must hold values that do not change between calls (literals, or functions
 * of known parameters only); they reach the pinned Jacobian buffer through
 * one full copy after this call / after opty_colloc_set_known (or from
 * `fill`, a K-entry per-node pattern, if given).  `num_ranges` = 0 restores
 * full copies. */
int opty_colloc_set_d2h_columns(opty_colloc_t* h, int num_ranges, const int32_t* col_begin,
                                const int32_t* col_end, const double* fill);

/* Problems too large for one nvcc run are compiled as several modules (contiguous
 * ranges of output groups) in parallel -- the reference compiles its single
 * generated C function serially (opty/utils.py:866-907).  Adds the module that
 * writes store segments [seg_first, seg_first + seg_count) with `num_groups`
 * output groups; modules are added in segment order and all of them before the
 * first evaluation.  The module given to opty_colloc_create carries the
 * invariants and pre-pass kernels. */
int opty_colloc_add_module(opty_colloc_t* h, const void* cubin, size_t cubin_bytes, int seg_first,
                           int seg_count, int num_groups);

/* Persistent main kernel only: the block -> work table, one (group, first tile,
 * end tile) triple per block (group = index inside the module, tiles of 32
 * nodes); every (group, tile) pair must be covered exactly once and there may
 * be at most one block per SM.  Replaces the `prange` node loop's static
 * OpenMP schedule (opty/utils.py:716-741) with a measured one. */
int opty_colloc_set_schedule(opty_colloc_t* h, int num_blocks, const int32_t* triples);

/* clock64 ticks every block of the last persistent launch spent in its group
 * phase (input for re-balancing the schedule). */
int opty_colloc_block_clocks(opty_colloc_t* h, int num_blocks, int64_t* clocks);

/* Registers the constant column runs of the node block (cfg.const_image_doubles
 * columns in total): run i covers columns [col0[i], col0[i]+len[i]) (even start
 * and length); `lit` / `inv_idx` give, in run order, each column's literal value
 * or (inv_idx >= 0) its index in the node-invariant table.  These are the
 * entries the reference recomputes for every node although they do not depend
 * on it (opty/utils.py:483-494 evaluates the full matrix per node); here one
 * image is replicated into all node rows by TMA tile stores.  Must be called
 * once after opty_colloc_create when cfg.const_image_doubles > 0. */
int opty_colloc_set_const_runs(opty_colloc_t* h, int num_runs, const int32_t* col0, const int32_t* len,
                               const double* lit, const int32_t* inv_idx);

/* CUDA-event duration (ms) of the kernels of the last evaluation. */
int opty_colloc_last_kernel_ms(opty_colloc_t* h, float* ms);

/* Measurement aid: launches `steps` device-resident evaluations back to back
 * on the handle's stream, bracketed by two CUDA events recorded on that same
 * stream, waits, and returns the elapsed device time of the batch in ms. */
int opty_colloc_time_device_evals(opty_colloc_t* h, int steps, float* total_ms);

/* Number of kernel launches issued by this handle so far. */
int opty_colloc_launch_count(opty_colloc_t* h, int64_t* count);

/* COO structure of the constraint Jacobian, bit-equal to the Python loop of
 * `jacobian_indices` (opty/direct_collocation.py:2628-2684) for the node range
 * [node_lo, node_hi): writes (node_hi-node_lo)*M*P row and column indices
 * (int64) to host memory.  Generated on `device`. */
int opty_colloc_jacobian_indices(int device, int N, int node_lo, int node_hi, int n, int q, int r,
                                 int s, int M, int method, int64_t* rows, int64_t* cols);

const char* opty_colloc_last_error(void);

int opty_b200_abi_version(void);

#ifdef __cplusplus
}
#endif

#endif /* OPTY_B200_H */
