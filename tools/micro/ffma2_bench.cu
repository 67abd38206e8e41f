// Micro-benchmark: issue rate and dependent-chain latency of packed fp32 FFMA2 vs scalar FFMA on sm_100a.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ffma2_bench tools/micro/ffma2_bench.cu && /tmp/ffma2_bench
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  float2 d;
  asm volatile("{.reg .b64 ra, rb, rc, rd;\n mov.b64 ra, {%2,%3};\n mov.b64 rb, {%4,%5};\n mov.b64 rc, {%6,%7};\n"
      " fma.rn.f32x2 rd, ra, rb, rc;\n mov.b64 {%0,%1}, rd;}"
      : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return d;
}

template <int CHAINS, bool PACKED, bool BCAST>
__global__ void k(float* out, float w, int iters, long long* cyc) {
  float2 acc[CHAINS];
  for (int j = 0; j < CHAINS; ++j) acc[j] = make_float2(threadIdx.x * 1e-3f + j, j * 0.5f);
  float2 a = make_float2(1.0f + threadIdx.x * 1e-6f, 1.0f - threadIdx.x * 1e-6f);
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < CHAINS; ++j) {
      if (PACKED) {
        acc[j] = ffma2(BCAST ? make_float2(a.x, a.x) : a, make_float2(w, w), acc[j]);
      } else {
        asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(acc[j].x) : "f"(a.x), "f"(w));
        asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(acc[j].y) : "f"(a.y), "f"(w));
      }
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
  for (int j = 0; j < CHAINS; ++j) s += acc[j].x + acc[j].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int CHAINS, bool PACKED, bool BCAST>
void run(const char* name, int warps) {
  float* out; long long* cyc; long long h;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
  const int iters = 4096;
  k<CHAINS, PACKED, BCAST><<<148, warps * 32>>>(out, 0.999f, iters, cyc);
  k<CHAINS, PACKED, BCAST><<<148, warps * 32>>>(out, 0.999f, iters, cyc);
  cudaDeviceSynchronize();
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  const double fma_per_thread = (double)iters * CHAINS * 2;
  printf("%-28s chains=%d warps/SM=%2d: %7.2f cycles per loop iteration, %6.2f fp32 FMA/clk/SM\n", name, CHAINS, warps,
         (double)h / iters, fma_per_thread * warps * 32 / (double)h);
  cudaFree(out); cudaFree(cyc);
}

int main() {
  run<1, false, false>("scalar FFMA x2", 1);
  run<1, true, false>("FFMA2", 1);
  run<1, true, true>("FFMA2 bcast", 1);
  run<4, false, false>("scalar FFMA x2", 1);
  run<4, true, false>("FFMA2", 1);
  run<4, true, true>("FFMA2 bcast", 1);
  run<8, false, false>("scalar FFMA x2", 4);
  run<8, true, false>("FFMA2", 4);
  run<8, false, false>("scalar FFMA x2", 16);
  run<8, true, false>("FFMA2", 16);
  run<8, true, true>("FFMA2 bcast", 16);
  run<8, false, false>("scalar FFMA x2", 32);
  run<8, true, false>("FFMA2", 32);
  return 0;
}
