// Development aid: store-path microbenchmark for the node-major Jacobian
// buffer.  Writes an [nodes][K] float64 matrix (row pitch Kp) the way the
// collocation kernel does -- lane = node, per-warp shared-memory tile of 32
// rows x C columns drained by 2-D TMA tile stores -- but with no arithmetic,
// to find the store pattern that reaches HBM speed.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o store_bench tools/store_bench.cu -lcuda
//   ./store_bench
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e = (x);                                                           \
    if (e != cudaSuccess) {                                                        \
      printf("%s failed: %s\n", #x, cudaGetErrorString(e));                        \
      exit(1);                                                                     \
    }                                                                              \
  } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(c0), "r"(c1), "r"(smem_u32(src))
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}

// mode 0: per-warp tile [32][C] (dense), TMA store per chunk
// mode 1: same with 128B swizzle (C must be 16)
// mode 2: direct per-lane st.global.v2.f64 (lane = node, strided)
// mode 3: warp-per-row coalesced 16-byte stores from the smem tile
template <int MODE>
__global__ void __launch_bounds__(256) store_kernel(const __grid_constant__ CUtensorMap tm, double* out, int nodes,
                                                   int K, long long Kp, int C, int nbuf, int work, int G) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int warps = blockDim.x >> 5;
  const int node0 = (blockIdx.x * warps + warp) * 32;
  if (node0 >= nodes) return;
  const int tile_bytes = 32 * C * 8;
  unsigned char* mytiles = smem + (size_t)warp * nbuf * tile_bytes;
  const int nchunks_all = (K + C - 1) / C;
  const int per = (nchunks_all + G - 1) / G;
  const int q_begin = blockIdx.y * per;
  const int q_end = min(nchunks_all, q_begin + per);
  double v = (double)(node0 + lane);
  for (int q = q_begin; q < q_end; ++q) {
    const int ncols = min(C, K - q * C);
    // a little dependent arithmetic per entry so that the loop is not empty
    for (int w = 0; w < work; ++w) v = v * 1.0000001 + 0.5;
    if (MODE == 2) {
      if (node0 + lane < nodes) {
        double* dst = out + (long long)(node0 + lane) * Kp + q * C;
        for (int c = 0; c + 1 < ncols; c += 2) *reinterpret_cast<double2*>(dst + c) = make_double2(v, v + c);
      }
      continue;
    }
    if (MODE == 7) {
      // tile rows padded by one 16-byte vector: pitch (C + 2) doubles
      const int pitch = C + 2;
      double* t7 = reinterpret_cast<double*>(smem + (size_t)warp * 32 * pitch * 8);
      for (int c = 0; c + 1 < ncols; c += 2)
        *reinterpret_cast<double2*>(t7 + lane * pitch + c) = make_double2(v, v + c);
      __syncwarp();
      const int V = ncols >> 1;          // 16-byte vectors per row
      const int rows = min(32, nodes - node0);
      for (int idx = lane; idx < rows * V; idx += 32) {
        const int r = idx / V, vv = idx - r * V;
        *reinterpret_cast<double2*>(out + (long long)(node0 + r) * Kp + q * C + 2 * vv) =
            *reinterpret_cast<const double2*>(t7 + r * pitch + 2 * vv);
      }
      __syncwarp();
      continue;
    }
    unsigned char* tile = mytiles + (q % nbuf) * tile_bytes;
    for (int c = 0; c + 1 < ncols; c += 2) {
      int chunk16 = c >> 1;
      if (MODE == 1) chunk16 ^= (lane & 7);
      *reinterpret_cast<double2*>(tile + lane * C * 8 + chunk16 * 16) = make_double2(v, v + c);
    }
    if (MODE == 3) {
      __syncwarp();
      const int rows = min(32, nodes - node0);
      for (int r = 0; r < rows; ++r) {
        double* dst = out + (long long)(node0 + r) * Kp + q * C;
        const double* src = reinterpret_cast<const double*>(tile + r * C * 8);
        for (int c = lane * 2; c + 1 < ncols; c += 64) *reinterpret_cast<double2*>(dst + c) = *reinterpret_cast<const double2*>(src + c);
      }
      __syncwarp();
      continue;
    }
    if (MODE != 5) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (MODE == 6) continue;
    if (lane == 0) {
      tma_store_2d(&tm, tile, q * C, node0);
      if (nbuf == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
      else if (nbuf == 3) asm volatile("cp.async.bulk.wait_group.read 2;" ::: "memory");
      else if (nbuf == 4) asm volatile("cp.async.bulk.wait_group.read 3;" ::: "memory");
      else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
    __syncwarp();
  }
  if ((MODE < 2 || MODE == 5) && lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

struct Variant {
  const char* name;
  int mode, K, Kp, C, warps, nbuf, work, G;
  int smem_floor = 0;  // > 0: request at least this much dynamic shared memory (caps the resident blocks per SM)
};

int main() {
  CK(cudaSetDevice(0));
  CK(cudaFree(0));
  const int nodes = 9999;
  const int nring = 4;
  const size_t maxbytes = (size_t)nodes * 1024 * 8;
  std::vector<double*> bufs(nring);
  for (auto& b : bufs) CK(cudaMalloc(&b, maxbytes));
  cudaStream_t st;
  CK(cudaStreamCreate(&st));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));

  // session 2: does the TMA store stream overlap with arithmetic at the real kernel's occupancy
  // (4 blocks x 2 warps per SM = 56 KB of shared memory per block)?
  const int F = 56 * 1024;
  std::vector<Variant> vs = {
      // does a 128-byte aligned row pitch (1024 columns instead of 1012) raise the store floor?
      {"tma store only  pitch 1012  C30 occ8", 0, 1012, 1012, 30, 2, 2, 0, 8, F},
      {"tma store only  pitch 1024  C30 occ8", 0, 1012, 1024, 30, 2, 2, 0, 8, F},
      {"tma store only  pitch 1012  C62 occ8", 0, 1012, 1012, 62, 2, 2, 0, 8, F},
      {"tma store only  pitch 1024  C62 occ8", 0, 1012, 1024, 62, 2, 2, 0, 8, F},
      {"tma store only  pitch 1024  C16 swizzled occ8", 1, 1012, 1024, 16, 2, 2, 0, 8, F},
      {"tma store only  pitch 1024  C46 occ8", 0, 1012, 1024, 46, 2, 2, 0, 8, F},
      {"tma store only  pitch 1012  C46 occ8", 0, 1012, 1012, 46, 2, 2, 0, 8, F},
      {"tma store+compute w250 pitch 1024 C30 occ8", 0, 1012, 1024, 30, 2, 2, 250, 8, F},
      {"tma store+compute w250 pitch 1012 C30 occ8", 0, 1012, 1012, 30, 2, 2, 250, 8, F},
  };

  for (const Variant& v : vs) {
    std::vector<CUtensorMap> maps(nring);
    bool ok = true;
    for (int r = 0; r < nring; ++r) {
      cuuint64_t gdim[2] = {(cuuint64_t)v.K, (cuuint64_t)nodes};
      cuuint64_t gstr[1] = {(cuuint64_t)v.Kp * 8};
      cuuint32_t box[2] = {(cuuint32_t)v.C, 32};
      cuuint32_t estr[2] = {1, 1};
      CUresult res = cuTensorMapEncodeTiled(&maps[r], CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, bufs[r], gdim, gstr, box,
                                            estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                            v.mode == 1 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                                            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (res != CUDA_SUCCESS) {
        printf("%-44s tensor map encode failed (%d)\n", v.name, (int)res);
        ok = false;
        break;
      }
    }
    if (!ok) continue;
    const int threads = v.warps * 32;
    const int blocks = (nodes + threads - 1) / threads;
    size_t smem = (size_t)v.warps * v.nbuf * 32 * (v.C + 2) * 8 + 1024;
    if ((size_t)v.smem_floor > smem) smem = (size_t)v.smem_floor;
    auto launch = [&](int r) {
      switch (v.mode) {
        case 0:
          cudaFuncSetAttribute(store_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
          store_kernel<0><<<dim3(blocks, v.G), threads, smem, st>>>(maps[r], bufs[r], nodes, v.K, v.Kp, v.C, v.nbuf, v.work, v.G);
          break;
        case 1:
          cudaFuncSetAttribute(store_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
          store_kernel<1><<<dim3(blocks, v.G), threads, smem, st>>>(maps[r], bufs[r], nodes, v.K, v.Kp, v.C, v.nbuf, v.work, v.G);
          break;
        case 2:
          store_kernel<2><<<dim3(blocks, v.G), threads, smem, st>>>(maps[r], bufs[r], nodes, v.K, v.Kp, v.C, v.nbuf, v.work, v.G);
          break;
        case 5:
          cudaFuncSetAttribute(store_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
          store_kernel<5><<<dim3(blocks, v.G), threads, smem, st>>>(maps[r], bufs[r], nodes, v.K, v.Kp, v.C, v.nbuf, v.work, v.G);
          break;
        case 7:
          cudaFuncSetAttribute(store_kernel<7>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
          store_kernel<7><<<dim3(blocks, v.G), threads, smem, st>>>(maps[r], bufs[r], nodes, v.K, v.Kp, v.C, v.nbuf, v.work, v.G);
          break;
        case 6:
          cudaFuncSetAttribute(store_kernel<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
          store_kernel<6><<<dim3(blocks, v.G), threads, smem, st>>>(maps[r], bufs[r], nodes, v.K, v.Kp, v.C, v.nbuf, v.work, v.G);
          break;
        default:
          cudaFuncSetAttribute(store_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
          store_kernel<3><<<dim3(blocks, v.G), threads, smem, st>>>(maps[r], bufs[r], nodes, v.K, v.Kp, v.C, v.nbuf, v.work, v.G);
      }
    };
    for (int i = 0; i < 8; ++i) launch(i % nring);
    CK(cudaStreamSynchronize(st));
    const int reps = 100;
    CK(cudaEventRecord(e0, st));
    for (int i = 0; i < reps; ++i) launch(i % nring);
    CK(cudaEventRecord(e1, st));
    CK(cudaEventSynchronize(e1));
    CK(cudaGetLastError());
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    const double us = 1e3 * ms / reps;
    const double bytes = (double)nodes * v.K * 8;
    printf("%-44s %8.2f us  %8.1f GB/s\n", v.name, us, bytes / us / 1e3);
  }
  return 0;
}
