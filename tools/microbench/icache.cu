// Instruction-cache microbenchmark for sm_100a: straight-line FFMA body of B instructions inside a loop, W warps per SM,
// warps either in lock step or skewed by a per-warp delay.  Prints cycles per instruction per warp.
#include <cstdio>
#include <cuda_runtime.h>

template <int B>
__global__ void __launch_bounds__(1024) body(float* out, int reps, int skew) {
  float a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const float m = 1.0001f, c = 0.5f;
  if (skew) {  // desynchronise the warps: warp w waits w * skew clocks
    long long t0 = clock64();
    long long wait = (long long)(threadIdx.x / 32) * skew;
    while (clock64() - t0 < wait) {}
  }
  long long t0 = clock64();
  for (int r = 0; r < reps; ++r) {
#pragma unroll
    for (int i = 0; i < B / 8; ++i) {
      asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a0) : "f"(m), "f"(c));
      asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a1) : "f"(m), "f"(c));
      asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a2) : "f"(m), "f"(c));
      asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a3) : "f"(m), "f"(c));
      asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a4) : "f"(m), "f"(c));
      asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a5) : "f"(m), "f"(c));
      asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a6) : "f"(m), "f"(c));
      asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a7) : "f"(m), "f"(c));
    }
  }
  long long t1 = clock64();
  if (threadIdx.x % 32 == 0) out[blockIdx.x * 32 + threadIdx.x / 32] = (float)(t1 - t0) / ((float)reps * B);
  if (a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 == 12345.f) out[0] = 0;
}

template <int B>
void run(int warps, int skew) {
  float* d;
  cudaMalloc(&d, 148 * 32 * sizeof(float));
  int reps = (1 << 22) / B;
  body<B><<<148, warps * 32>>>(d, 2, skew);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  body<B><<<148, warps * 32>>>(d, reps, skew);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  float h[32]; cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
  float mx = 0; for (int i = 0; i < warps; ++i) mx = h[i] > mx ? h[i] : mx;
  // chip-wide warp-instructions per SM-cycle
  double ipc_sm = (double)reps * B * warps / (ms * 1e-3 * 1.965e9);
  printf("body %6d instrs (%4d KB)  warps/SM %2d  skew %6d : %.3f cycles/instr/warp  (SM IPC %.2f)\n", B, B * 16 / 1024, warps, skew, mx, ipc_sm);
  cudaFree(d);
}

int main() {
  for (int skew = 0; skew <= 1; ++skew)
    for (int w : {4, 8, 16}) {
      int sk = skew ? 3001 : 0;
      run<256>(w, sk); run<1024>(w, sk); run<2048>(w, sk); run<3072>(w, sk); run<4096>(w, sk); run<8192>(w, sk); run<16384>(w, sk);
    }
  return 0;
}
