// Development aid: FP64 dependent-issue latency and per-SM throughput on the
// device (cycles per DFMA for ILP independent chains and W warps per block,
// one block per SM).
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void chain(double* out, long long* cyc, int iters, double a, double b) {
  double x[ILP];
  for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x * 1e-3 + i;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u)
#pragma unroll
      for (int i = 0; i < ILP; ++i) x[i] = fma(x[i], a, b);
  }
  long long t1 = clock64();
  double s = 0;
  for (int i = 0; i < ILP; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int ILP>
void run(int warps, int blocks) {
  double* out;
  long long* cyc;
  cudaMalloc(&out, sizeof(double) * blocks * warps * 32);
  cudaMalloc(&cyc, sizeof(long long) * blocks);
  const int iters = 2000;
  chain<ILP><<<blocks, warps * 32>>>(out, cyc, iters, 0.999, 1e-3);
  chain<ILP><<<blocks, warps * 32>>>(out, cyc, iters, 0.999, 1e-3);
  long long h[1];
  cudaMemcpy(h, cyc, sizeof(long long), cudaMemcpyDeviceToHost);
  double per = (double)h[0] / (iters * 8.0 * ILP);
  printf("ILP %d warps/SM %2d: %.2f cycles per DFMA per warp, %.2f warp-DFMA/cycle/SM\n", ILP, warps, per,
         warps / per);
  cudaFree(out);
  cudaFree(cyc);
}

int main() {
  for (int w : {1, 4, 8, 16, 32}) {
    run<1>(w, 148);
    run<2>(w, 148);
    run<4>(w, 148);
    run<8>(w, 148);
  }
  return 0;
}
