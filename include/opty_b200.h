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
 *     2816-2887), the index loop of `jacobian_indices`
 *     (opty/direct_collocation.py:2628-2684), and the quadrature of
 *     `create_objective_function` (opty/utils.py:329-470).
 *
 * Plain pointers and sizes only.  All values are IEEE float64, all indices
 * int64.  Every function returns 0 on success and a negative code on failure;
 * `opty_colloc_last_error()` then describes the failure (thread local).
 *
 * The configuration holds the PROBLEM in the reference's notation only.  The
 * kernel geometry (block size, staging tiles, tensor-map shapes ...) is a
 * property of the generated module and is read from the module itself (its
 * `opty_module_info` table); it does not cross this boundary.
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

#define OPTY_B200_ABI_VERSION 7
#define OPTY_MAX_GROUPS 1024

#define OPTY_OK 0
#define OPTY_ERR_ARG -1     /* invalid argument / configuration */
#define OPTY_ERR_CUDA -2    /* CUDA runtime or driver error */
#define OPTY_ERR_STATE -3   /* call sequence error (e.g. known values not set) */

/* OPTY_ELEMENTWISE: the generic `ufuncify_matrix` operator (opty/utils.py:
 * 639-670): N evaluation points, n array arguments of length N, r scalar
 * ("const") arguments after them in `free`, an M x P output matrix per point
 * (returned through opty_colloc_jacobian), no residuals, no neighbour column. */
enum { OPTY_BACKWARD_EULER = 0, OPTY_MIDPOINT = 1, OPTY_ELEMENTWISE = 2 };

/* The problem, in the reference's notation (opty/direct_collocation.py:101-114). */
typedef struct opty_colloc_cfg {
  int32_t abi_version;   /* OPTY_B200_ABI_VERSION */
  int32_t device;        /* CUDA device ordinal */
  int32_t N;             /* collocation nodes of the whole problem */
  int32_t node_lo;       /* this handle evaluates constraint nodes */
  int32_t node_hi;       /*   [node_lo, node_hi) of the N-1 (a shard) */
  int32_t n, q, k;       /* states, unknown / known input trajectories */
  int32_t r, s, pk;      /* unknown parameters, 1 if h is free, known parameters */
  int32_t M, P;          /* equations of motion, partials per equation */
  int32_t method;        /* OPTY_BACKWARD_EULER / OPTY_MIDPOINT / OPTY_ELEMENTWISE */
  int32_t out_ring;      /* device output sets to rotate through (>= 1) */
  int32_t prefetch_jac;  /* opty_colloc_constraints starts the Jacobian D2H speculatively */
  int32_t con_tail;      /* extra host slots after the M*(N-1) residuals (instance constraints) */
  int32_t jac_tail;      /* extra host slots after the (N-1)*M*P partials */
  double h;              /* fixed node time interval (ignored when s = 1) */
} opty_colloc_cfg;

typedef struct opty_colloc opty_colloc_t;

/* Loads the generated sm_100a module (`cubin`, produced by nvcc from the
 * emitter's CUDA-C) on `cfg->device`, allocates the device-resident
 * trajectory matrix, residual / Jacobian buffers and pinned host buffers.
 * Replaces importing the compiled Cython module (opty/utils.py:909-916). */
int opty_colloc_create(const opty_colloc_cfg* cfg, const void* cubin, size_t cubin_bytes,
                       opty_colloc_t** out);

int opty_colloc_destroy(opty_colloc_t* h);

/* Problems too large for one nvcc run are compiled as several modules
 * (contiguous ranges of output groups) in parallel -- the reference compiles
 * its single generated C function serially (opty/utils.py:866-907).  Adds one
 * of them; all must be added before the first evaluation.  The module given
 * to opty_colloc_create carries the invariants and pre-pass kernels. */
int opty_colloc_add_module(opty_colloc_t* h, const void* cubin, size_t cubin_bytes);

/* Known input trajectories `traj` as [k][N] (row-major, full problem length N)
 * and known parameter values `params` [pk], in the collocator's symbol order.
 * Replaces the known-value half of `_merge_fixed_free`
 * (opty/direct_collocation.py:2911-2926).  Must be called once before the
 * first evaluation (also when k = pk = 0) and again whenever a known value
 * changes. */
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

/* Several handles (one per GPU, each a shard of the constraint nodes) driven
 * by ONE host thread: `begin` uploads, launches and queues the device->host
 * copies of what is asked for without waiting, `finish` waits.  With
 * opty_colloc_set_host_outputs every handle writes straight into its slice of
 * one full-problem host vector: the node-major Jacobian block of a shard is
 * contiguous (opty/direct_collocation.py:2681-2684), its residuals are M
 * strided segments of the eom-major vector (opty/direct_collocation.py:2446). */
int opty_colloc_begin(opty_colloc_t* h, const double* free_host, int want_con, int want_jac);
int opty_colloc_finish(opty_colloc_t* h);

/* Full-problem host vectors (from opty_host_alloc) that this handle's
 * device->host copies target instead of its own pinned buffers: residuals
 * M*(N-1) (+ tail), Jacobian (N-1)*M*P (+ tail).  NULL, NULL restores the
 * handle's own buffers.  Speculative Jacobian copies are off in this mode. */
int opty_colloc_set_host_outputs(opty_colloc_t* h, double* con_full, double* jac_full);

/* Page-locked host memory usable from every device (cudaHostAllocPortable). */
int opty_host_alloc(size_t bytes, void** ptr);
int opty_host_free(void* ptr);

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
 * of known parameters only); they reach the host Jacobian buffer through one
 * full copy after this call / after opty_colloc_set_known /
 * opty_colloc_invalidate_host_jacobian.  `num_ranges` = 0 restores full
 * copies. */
int opty_colloc_set_d2h_columns(opty_colloc_t* h, int num_ranges, const int32_t* col_begin,
                                const int32_t* col_end);

/* The next Jacobian fetch copies every column again (for consumers that
 * modified the returned buffer in place). */
int opty_colloc_invalidate_host_jacobian(opty_colloc_t* h);

/* Quadrature of a running cost over the node grid, the numeric half of
 * `create_objective_function` (opty/utils.py:329-470): the handle (created
 * with method = OPTY_ELEMENTWISE from a module that evaluates the integrand
 * in column 0 and its partials with respect to the n array arguments in
 * columns 1..n at every node) returns
 *     value = scale * sum_i w_i * integrand(node i)
 *     grad[a*N + i] = scale * w_i * d integrand / d arg_a (node i)
 * with backward-Euler weights w_0 = 0, w_i = 1 (opty/utils.py:418-423) or
 * midpoint weights (method = OPTY_MIDPOINT: the integrand module is then
 * evaluated at the N-1 midpoints, opty/utils.py:424-434).  `grad` has n*N
 * entries. */
int opty_colloc_quadrature(opty_colloc_t* h, const double* free_host, double scale, int rule,
                           double* value, double* grad);

/* CUDA-event duration (ms) of the kernels of the last evaluation started with
 * opty_colloc_eval_device. */
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
