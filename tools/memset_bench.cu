// Development aid: write-only HBM bandwidth (coalesced 16-byte stores and
// cudaMemsetAsync) over a ring of buffers larger than L2, as the ceiling for
// the Jacobian store path.
#include <cuda_runtime.h>
#include <stdio.h>
#include <vector>
__global__ void fill(double2* out, size_t n, double v) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    out[i] = make_double2(v, v);
}
__global__ void copyk(const double2* in, double2* out, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    out[i] = in[i];
}
int main() {
  const size_t bytes = (size_t)9999 * 1012 * 8;
  const int nring = 6;
  std::vector<double2*> b(nring);
  for (auto& p : b) cudaMalloc(&p, bytes);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms;
  for (int blocks : {148 * 4, 148 * 8, 148 * 16, 148 * 32}) {
    for (int i = 0; i < 10; ++i) fill<<<blocks, 256>>>(b[i % nring], bytes / 16, 1.0);
    cudaEventRecord(e0);
    for (int i = 0; i < 100; ++i) fill<<<blocks, 256>>>(b[i % nring], bytes / 16, 1.0);
    cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
    printf("fill kernel   blocks %5d: %7.2f us  %7.1f GB/s (write only)\n", blocks, ms * 10, bytes / (ms * 10) / 1e3);
  }
  cudaEventRecord(e0);
  for (int i = 0; i < 100; ++i) cudaMemsetAsync(b[i % nring], 0, bytes);
  cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
  printf("cudaMemsetAsync           : %7.2f us  %7.1f GB/s (write only)\n", ms * 10, bytes / (ms * 10) / 1e3);
  for (int i = 0; i < 10; ++i) copyk<<<148 * 16, 256>>>(b[i % nring], b[(i + 3) % nring], bytes / 16);
  cudaEventRecord(e0);
  for (int i = 0; i < 100; ++i) copyk<<<148 * 16, 256>>>(b[i % nring], b[(i + 3) % nring], bytes / 16);
  cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
  printf("copy kernel               : %7.2f us  %7.1f GB/s (read+write)\n", ms * 10, 2 * bytes / (ms * 10) / 1e3);
  // large buffer write (1 GiB) for the asymptote
  double2* big; size_t bigb = (size_t)1 << 30; cudaMalloc(&big, bigb);
  fill<<<148 * 16, 256>>>(big, bigb / 16, 1.0);
  cudaEventRecord(e0);
  for (int i = 0; i < 10; ++i) fill<<<148 * 16, 256>>>(big, bigb / 16, 1.0);
  cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
  printf("fill 1 GiB                : %7.2f us  %7.1f GB/s (write only)\n", ms * 100, bigb / (ms * 100) / 1e3);
  return 0;
}
