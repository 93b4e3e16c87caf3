// Development aid: which store path fills the node-major Jacobian ([nodes][K]
// float64, K = 1012, ring of 4 x 81 MB > L2) fastest from a persistent grid of
// one 256-thread block per SM, with no arithmetic.
//
//   A  linear fill, 16-byte stores
//   B  first half of every node row (4048 B) with coalesced 16-byte stores          (41 MB)
//   C  first half of every node row with one 1-D bulk copy from shared memory       (41 MB)
//   D  whole node rows (8096 B) with one 1-D bulk copy each                         (81 MB)
//   E  2-D TMA tile stores [32 nodes x 46 columns], all 22 equation rows            (81 MB)
//   F  C + tile stores for the second half (the row-stationary kernel's pattern)    (81 MB)
//   G  B + tile stores for the second half                                          (81 MB)
//   H  like F with tiles of [32 x 92]                                               (81 MB)
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/bin/write_path_bench tools/write_path_bench.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x)                                                               \
  do {                                                                      \
    cudaError_t e = (x);                                                    \
    if (e != cudaSuccess) {                                                 \
      printf("%s failed: %s\n", #x, cudaGetErrorString(e));                 \
      exit(1);                                                              \
    }                                                                       \
  } while (0)

constexpr int K = 1012, P = 46, M = 22, NODES = 9999, HALF = 506;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(c0), "r"(c1), "r"(smem_u32(src))
               : "memory");
}
__device__ __forceinline__ void bulk_store_1d(void* dst, const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

struct Maps {
  CUtensorMap w46, w92;
};

// tiles of the second half: (tile of 32 nodes, row j in 11..21), dealt out to warps round robin
template <int W>
__device__ void second_half_tiles(const CUtensorMap* map, double* tile, int gw, int nw, int lane, int rows0) {
  const int ntiles = (NODES + 31) / 32;
  const int per = (M - rows0) * P / W;  // column chunks per node tile
  for (int it = gw; it < ntiles * per; it += nw) {
    const int t = it / per, c = rows0 * P + (it % per) * W;
    if (lane == 0) wait_read0();
    __syncwarp();
    for (int k = 0; k < W; k += 2) *reinterpret_cast<double2*>(tile + lane * W + k) = make_double2(it, k);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0) {
      tma_store_2d(map, tile, c, t * 32);
      commit();
    }
  }
}

// all 22 column chunks as tiles; a warp keeps ONE chunk (like an SM that keeps one equation) and walks the
// node tiles.  shift = 0: the 22 pieces of a node tile are written in the same step by 22 warps;
// shift > 0: at different steps (chunk c is `shift * c` tiles ahead)
__device__ void chunk_stationary_tiles(const CUtensorMap* map, double* tile, int gw, int lane, int shift) {
  const int ntiles = (NODES + 31) / 32, per = M, lanes_per = 53, span = 318;
  const int c = gw % per, i = gw / per;
  if (i >= lanes_per) return;
  for (int k = 0; k * lanes_per < span; ++k) {
    const int t = (i + k * lanes_per + shift * c) % span;
    if (t >= ntiles) continue;
    if (lane == 0) wait_read0();
    __syncwarp();
    for (int q = 0; q < P; q += 2) *reinterpret_cast<double2*>(tile + lane * P + q) = make_double2(t, q);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0) {
      tma_store_2d(map, tile, c * P, t * 32);
      commit();
    }
  }
}

template <int MODE>
__global__ void __launch_bounds__(256, 1) wk(const __grid_constant__ Maps tm, double* out) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int gw = blockIdx.x * 8 + warp, nw = gridDim.x * 8;
  double* rowbuf = reinterpret_cast<double*>(smem);                       // 8096 B
  double* tile = reinterpret_cast<double*>(smem + 8192 + warp * 32 * 92 * 8);  // per-warp staging
  if (MODE == 0) {
    double2* o = reinterpret_cast<double2*>(out);
    const size_t n = (size_t)NODES * K / 2;
    for (size_t i = blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) o[i] = make_double2(1.0, 2.0);
    return;
  }
  for (int i = threadIdx.x; i < K; i += 256) rowbuf[i] = i;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  const int npw = (NODES + nw - 1) / nw;
  const int n0 = gw * npw, n1 = min(NODES, n0 + npw);
  if (MODE == 1 || MODE == 6) {  // B / G: coalesced 16-byte stores of the first half
    for (int k = lane; k < HALF / 2; k += 32) {
      const double2 v = reinterpret_cast<const double2*>(rowbuf)[k];
      double2* d = reinterpret_cast<double2*>(out + (long long)n0 * K) + k;
      for (int n = n0; n < n1; ++n, d += K / 2) *d = v;
    }
  }
  if (MODE == 2 || MODE == 5 || MODE == 7) {  // C / F / H: one bulk copy per node, lanes 1..31
    if (lane > 0)
      for (int n = n0 + lane - 1; n < n1; n += 31) bulk_store_1d(out + (long long)n * K, rowbuf, HALF * 8);
    commit();
  }
  if (MODE == 3) {  // D: whole rows
    for (int n = n0 + lane; n < n1; n += 32) bulk_store_1d(out + (long long)n * K, rowbuf, K * 8);
    commit();
  }
  if (MODE == 4) second_half_tiles<46>(&tm.w46, tile, gw, nw, lane, 0);
  if (MODE == 5 || MODE == 6) second_half_tiles<46>(&tm.w46, tile, gw, nw, lane, 11);
  if (MODE == 7) second_half_tiles<92>(&tm.w92, tile, gw, nw, lane, 11);
  if (MODE == 8) chunk_stationary_tiles(&tm.w46, tile, gw, lane, 0);
  if (MODE == 9) chunk_stationary_tiles(&tm.w46, tile, gw, lane, 14);
  if (MODE == 10) chunk_stationary_tiles(&tm.w46, tile, gw, lane, 1);
  wait_read0();
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static CUtensorMap make_map(EncodeFn enc, double* base, int w) {
  CUtensorMap m;
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)NODES};
  cuuint64_t strides[1] = {(cuuint64_t)K * 8};
  cuuint32_t box[2] = {(cuuint32_t)w, 32};
  cuuint32_t es[2] = {1, 1};
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    printf("encode failed %d\n", (int)r);
    exit(1);
  }
  return m;
}

template <int MODE>
static void run(const char* name, EncodeFn enc, double** ring, int nring, double bytes) {
  const int smem = 8192 + 8 * 32 * 92 * 8;
  CK(cudaFuncSetAttribute(wk<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  Maps maps[8];
  for (int i = 0; i < nring; ++i) {
    maps[i].w46 = make_map(enc, ring[i], 46);
    maps[i].w92 = make_map(enc, ring[i], 92);
  }
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int i = 0; i < 8; ++i) wk<MODE><<<148, 256, smem>>>(maps[i % nring], ring[i % nring]);
  CK(cudaDeviceSynchronize());
  cudaEventRecord(e0);
  const int reps = 100;
  for (int i = 0; i < reps; ++i) wk<MODE><<<148, 256, smem>>>(maps[i % nring], ring[i % nring]);
  cudaEventRecord(e1);
  CK(cudaEventSynchronize(e1));
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  printf("%-62s %7.2f us  %7.1f GB/s\n", name, ms * 1e3 / reps, bytes / (ms * 1e-3 / reps) / 1e9);
}

int main() {
  EncodeFn enc = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &q));
  const int nring = 4;
  double* ring[nring];
  const size_t bytes = (size_t)NODES * K * 8;
  for (auto& p : ring) CK(cudaMalloc(&p, bytes + 4096));
  const double full = (double)bytes, half = full / 2;
  run<0>("A linear fill, 16-byte stores", enc, ring, nring, full);
  run<1>("B first half of each node row, coalesced 16-byte stores", enc, ring, nring, half);
  run<2>("C first half of each node row, one 4048-byte bulk copy", enc, ring, nring, half);
  run<3>("D whole node rows, one 8096-byte bulk copy each", enc, ring, nring, full);
  run<4>("E tile stores [32 x 46], all rows", enc, ring, nring, full);
  run<5>("F bulk copies (first half) + tile stores [32 x 46]", enc, ring, nring, full);
  run<6>("G 16-byte stores (first half) + tile stores [32 x 46]", enc, ring, nring, full);
  run<7>("H bulk copies (first half) + tile stores [32 x 92]", enc, ring, nring, full);
  run<8>("I chunk-stationary warps, pieces of a node tile in the same step", enc, ring, nring, full);
  run<9>("J chunk-stationary warps, pieces 14 tiles apart", enc, ring, nring, full);
  run<10>("K chunk-stationary warps, pieces 1 tile apart", enc, ring, nring, full);
  return 0;
}
