// Microbenchmark: FP64 vs FP32 FMA throughput and double log/exp cost on this GPU.
#include <cstdio>
#include <cuda_runtime.h>
template <typename T> __global__ void fma_loop(T* out, int iters) {
  T a = (T)threadIdx.x * (T)1e-3, b = (T)1.000001, c = (T)1e-7;
  T x0 = a, x1 = a + 1, x2 = a + 2, x3 = a + 3;
  for (int i = 0; i < iters; ++i) { x0 = x0 * b + c; x1 = x1 * b + c; x2 = x2 * b + c; x3 = x3 * b + c; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3;
}
__global__ void dlog_loop(float* out, int iters) {
  float v = 1.5f + threadIdx.x * 1e-3f; float acc = 0;
  for (int i = 0; i < iters; ++i) { acc += (float)log2((double)(v + acc * 1e-9f)); }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
__global__ void flog_loop(float* out, int iters) {
  float v = 1.5f + threadIdx.x * 1e-3f; float acc = 0;
  for (int i = 0; i < iters; ++i) { acc += log2f(v + acc * 1e-9f); }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
int main() {
  void* buf; cudaMalloc(&buf, 148 * 8 * 256 * 8);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms; int iters = 20000;
  for (int rep = 0; rep < 2; ++rep) {
    cudaEventRecord(e0); fma_loop<float><<<148 * 8, 256>>>((float*)buf, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1); printf("fp32 fma: %.2f TFLOP/s\n", 2.0 * 4 * iters * 148 * 8 * 256 / ms / 1e9);
    cudaEventRecord(e0); fma_loop<double><<<148 * 8, 256>>>((double*)buf, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1); printf("fp64 fma: %.2f TFLOP/s\n", 2.0 * 4 * iters * 148 * 8 * 256 / ms / 1e9);
    cudaEventRecord(e0); dlog_loop<<<148, 32>>>((float*)buf, 1000); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1); printf("double log2, 1 warp/SM, 1000 dependent: %.3f us each\n", ms * 1e3 / 1000);
    cudaEventRecord(e0); flog_loop<<<148, 32>>>((float*)buf, 1000); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1); printf("float log2f, 1 warp/SM, 1000 dependent: %.3f us each\n", ms * 1e3 / 1000);
  }
  return 0;
}
