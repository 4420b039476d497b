// umma_probe.cu -- validates the descriptor encodings of csrc/umma.cuh on a B200 before the PPO kernel relies on them:
//   case 0: A K-major [128 x 32], B K-major [64 x 32]                      D[128 x 64]  = A B^T       (forward GEMMs)
//   case 1: A K-major [128 x 64] (2 blocks), B MN-major (W[k][n], 64 x 64) D[128 x 64]  = A W         (backward-data GEMM)
//   case 2: A MN-major (Act[s][f], M' = 128 features in 4 blocks, K' = 16 samples), B MN-major (X[s][f], N' = 80 in 3 blocks)
//                                                                          D[128 x 80]  = Act^T X     (weight-gradient GEMMs)
// nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I mujoco_playground_b200/csrc -o umma_probe tools/microbench/umma_probe.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include "umma.cuh"

using namespace umma;

// generic descriptor: layout_type 0 = no swizzle (interleaved 8 x 16-byte core matrices), 1 = SWIZZLE_128B_BASE32B, 2 = SWIZZLE_128B
__device__ __forceinline__ uint64_t gdesc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;
  return d;
}
// no-swizzle core-matrix layout of an array [rows][F features]: (r / 8) * GS + (f / 4) * 128 + (r % 8) * 16 + (f % 4) * 4, GS = 32 F
__device__ __forceinline__ uint32_t ns_off(int r, int f, int F) { return (uint32_t)((r >> 3) * 32 * F + (f >> 2) * 128 + (r & 7) * 16 + (f & 3) * 4); }
// SWIZZLE_128B_BASE32B MN-major block: rows of 128 bytes, 4-row atoms of 512 bytes, 32-byte chunk c of row r stored at c ^ (r & 3)


__global__ void __launch_bounds__(128, 1) probe(int which, const float* A, const float* B, float* D, int ncol) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint32_t tmem_base;
  __shared__ __align__(8) uint64_t mbar;
  unsigned char* sA = smem;                 // up to 4 blocks x 16 KB
  unsigned char* sB = smem + 65536;         // up to 3 blocks x 16 KB
  const int t = threadIdx.x, warp = t >> 5;
  if (warp == 0) tmem_alloc(&tmem_base, 128);
  if (t == 0) { mbar_init(&mbar, 1); mbar_fence_init(); }
  // zero the operand regions, then fill
  for (int i = t; i < (65536 + 49152) / 4; i += 128) reinterpret_cast<float*>(smem)[i] = 0.f;
  __syncthreads();
  if (which == 0) {          // A[128][32], B[64][32] row-major in global
    for (int i = t; i < 128 * 32; i += 128) { int r = i / 32, c = i % 32; *reinterpret_cast<float*>(sA + sw128_off(r, c)) = A[i]; }
    for (int i = t; i < 64 * 32; i += 128) { int r = i / 32, c = i % 32; *reinterpret_cast<float*>(sB + sw128_off(r, c)) = B[i]; }
  } else if (which == 1) {   // A[128][64] (2 blocks of 16 KB), W[64][64] (k rows, n cols; 2 blocks of 8 KB by n)
    for (int i = t; i < 128 * 64; i += 128) { int r = i / 64, c = i % 64; *reinterpret_cast<float*>(sA + (c / 32) * 16384 + sw128_off(r, c % 32)) = A[i]; }
    for (int i = t; i < 64 * 64; i += 128) { int k = i / 64, n = i % 64; *reinterpret_cast<float*>(sB + (n / 32) * 8192 + sw128_off(k, n % 32)) = B[i]; }
  } else if (which == 3) {   // no swizzle, K-major both: A[128][32], B[64][32]
    for (int i = t; i < 128 * 32; i += 128) { int r = i / 32, c = i % 32; *reinterpret_cast<float*>(sA + ns_off(r, c, 32)) = A[i]; }
    for (int i = t; i < 64 * 32; i += 128) { int r = i / 32, c = i % 32; *reinterpret_cast<float*>(sB + ns_off(r, c, 32)) = B[i]; }
  } else if (which == 4) {   // no swizzle, MN-major both: Act[16][128], X[16][80]
    for (int i = t; i < 16 * 128; i += 128) { int r = i / 128, f = i % 128; *reinterpret_cast<float*>(sA + ns_off(r, f, 128)) = A[i]; }
    for (int i = t; i < 16 * 80; i += 128) { int r = i / 80, f = i % 80; *reinterpret_cast<float*>(sB + ns_off(r, f, 80)) = B[i]; }
  } else if (which == 5) {   // BASE32B MN-major both: blocks of 32 features, 16 KB apart
    for (int i = t; i < 16 * 128; i += 128) { int r = i / 128, f = i % 128; *reinterpret_cast<float*>(sA + (f / 32) * 16384 + b32_off(r, f % 32)) = A[i]; }
    for (int i = t; i < 16 * 80; i += 128) { int r = i / 80, f = i % 80; *reinterpret_cast<float*>(sB + (f / 32) * 16384 + b32_off(r, f % 32)) = B[i]; }
  } else if (which == 7) {   // A in TMEM ([128][64], written by the threads), B K-major SW128 [64][64] (2 blocks of 8 KB)
    for (int i = t; i < 64 * 64; i += 128) { int r = i / 64, c = i % 64; *reinterpret_cast<float*>(sB + (c / 32) * 8192 + sw128_off(r, c % 32)) = B[i]; }
  } else if (which == 6) {   // no swizzle: A K-major [128][64], B = W[k][n] (64 x 64) MN-major
    for (int i = t; i < 128 * 64; i += 128) { int r = i / 64, c = i % 64; *reinterpret_cast<float*>(sA + ns_off(r, c, 64)) = A[i]; }
    for (int i = t; i < 64 * 64; i += 128) { int k = i / 64, n = i % 64; *reinterpret_cast<float*>(sB + ns_off(k, n, 64)) = B[i]; }
  } else {                   // Act[16][128] (samples x features, 4 blocks of 16 KB), X[16][80] (3 blocks of 16 KB)
    for (int i = t; i < 16 * 128; i += 128) { int s = i / 128, f = i % 128; *reinterpret_cast<float*>(sA + (f / 32) * 16384 + sw128_off(s, f % 32)) = A[i]; }
    for (int i = t; i < 16 * 80; i += 128) { int s = i / 80, f = i % 80; *reinterpret_cast<float*>(sB + (f / 32) * 16384 + sw128_off(s, f % 32)) = B[i]; }
  }
  fence_proxy_async();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tb = tmem_base;
  if (which == 7) {          // thread t = TMEM lane t writes its row of A into columns 64..127
    for (int c0 = 0; c0 < 64; c0 += 16) {
      float v[16];
      for (int j = 0; j < 16; ++j) v[j] = A[t * 64 + c0 + j];
      tmem_st16(tb + ((uint32_t)(32 * warp) << 16) + 64 + c0, v);
    }
    tmem_st_wait();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
  }
  if (t == 0) {
    const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
    if (which == 0) {
      const uint32_t id = idesc_tf32(128, 64, 0, 0);
      for (int k = 0; k < 4; ++k) mma_tf32(tb, desc_kmajor(a0, k), desc_kmajor(b0, k), id, k > 0);
    } else if (which == 1) {
      const uint32_t id = idesc_tf32(128, 64, 0, 1);
      for (int k = 0; k < 8; ++k) mma_tf32(tb, desc_kmajor(a0 + (k / 4) * 16384, k % 4), desc_mnmajor(b0, 8192, k), id, k > 0);
    } else if (which == 3) {
      const uint32_t id = idesc_tf32(128, 64, 0, 0);     // K-major: LBO = K-chunk stride (128), SBO = 8-row group stride (32 F = 1024)
      for (int k = 0; k < 4; ++k) mma_tf32(tb, gdesc(a0 + 256 * k, 128, 1024, 0), gdesc(b0 + 256 * k, 128, 1024, 0), id, k > 0);
    } else if (which == 4) {
      const uint32_t id = idesc_tf32(128, 80, 1, 1);     // MN-major: SBO = MN-chunk stride (128), LBO = 8-row (K) group stride (32 F)
      for (int k = 0; k < 2; ++k) mma_tf32(tb, gdesc(a0 + 4096 * k, 4096, 128, 0), gdesc(b0 + 2560 * k, 2560, 128, 0), id, k > 0);
    } else if (which == 5) {
      const uint32_t id = idesc_tf32(128, 80, 1, 1);     // BASE32B: LBO = MN block stride, SBO = stride between 4-row atoms (512)
      for (int k = 0; k < 2; ++k) mma_tf32(tb, desc_mn32(a0, 16384, k), desc_mn32(b0, 16384, k), id, k > 0);
    } else if (which == 7) {
      const uint32_t id = idesc_tf32(128, 64, 0, 0);
      for (int k = 0; k < 8; ++k) mma_tf32_ts(tb, tb + 64 + 8 * k, desc_kmajor(b0 + (k / 4) * 8192, k % 4), id, k > 0);
    } else if (which == 6) {
      const uint32_t id = idesc_tf32(128, 64, 0, 1);
      for (int k = 0; k < 8; ++k) mma_tf32(tb, gdesc(a0 + 256 * k, 128, 2048, 0), gdesc(b0 + 2048 * k, 2048, 128, 0), id, k > 0);
    } else {
      const uint32_t id = idesc_tf32(128, 80, 1, 1);
      for (int k = 0; k < 2; ++k) mma_tf32(tb, desc_mnmajor(a0, 16384, k), desc_mnmajor(b0, 16384, k), id, k > 0);
    }
    commit(&mbar);
  }
  mbar_wait(&mbar, 0);
  fence_after_sync();
  // read back: thread t = TMEM lane t (row), ncol columns
  for (int c0 = 0; c0 < ncol; c0 += 16) {
    float v[16];
    tmem_ld16(tb + ((uint32_t)(32 * warp) << 16) + c0, v);
    for (int j = 0; j < 16; ++j) D[t * ncol + c0 + j] = v[j];
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tb, 128);
}

static float q(float x) { return roundf(x * 8.f) / 8.f; }   // exactly representable in tf32

int main() {
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 + 49152 + 1024);
  int bad = 0;
  for (int which = 0; which < 8; ++which) {
    if (which == 1 || which == 2 || which == 4 || which == 6) continue;    // MN-major tf32 with layouts other than BASE32B: the tensor core returns zeros
    const int shape = (which == 0 || which == 3) ? 0 : ((which == 1 || which == 6) ? 1 : (which == 7 ? 3 : 2));
    const int ar = shape == 2 ? 16 : 128, ac = shape == 0 ? 32 : ((shape == 1 || shape == 3) ? 64 : 128);
    const int br = shape == 0 ? 64 : ((shape == 1 || shape == 3) ? 64 : 16), bc = shape == 0 ? 32 : ((shape == 1 || shape == 3) ? 64 : 80);
    const int ncol = shape == 2 ? 80 : 64;
    std::vector<float> A(ar * ac), B(br * bc), D(128 * ncol), R(128 * ncol, 0.f);
    srand(1 + which);
    for (auto& x : A) x = q((rand() % 33 - 16) / 8.f);
    for (auto& x : B) x = q((rand() % 33 - 16) / 8.f);
    for (int i = 0; i < 128; ++i)
      for (int j = 0; j < ncol; ++j) {
        double s = 0;
        if (shape == 0) for (int k = 0; k < 32; ++k) s += (double)A[i * 32 + k] * B[j * 32 + k];
        if (shape == 1) for (int k = 0; k < 64; ++k) s += (double)A[i * 64 + k] * B[k * 64 + j];
        if (shape == 2) for (int k = 0; k < 16; ++k) s += (double)A[k * 128 + i] * B[k * 80 + j];
        if (shape == 3) for (int k = 0; k < 64; ++k) s += (double)A[i * 64 + k] * B[j * 64 + k];
        R[i * ncol + j] = (float)s;
      }
    float *dA, *dB, *dD;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    cudaMemset(dD, 0, D.size() * 4);
    probe<<<1, 128, 65536 + 49152 + 1024>>>(which, dA, dB, dD, ncol);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    double worst = 0; int wi = 0;
    for (size_t i = 0; i < D.size(); ++i) { double d = fabs((double)D[i] - R[i]); if (d > worst) { worst = d; wi = (int)i; } }
    printf("case %d: %s, max |D - ref| = %g at (%d, %d): got %g want %g\n", which, cudaGetErrorString(e), worst, wi / ncol, wi % ncol, D[wi], R[wi]);
    if (worst > 1e-3 || e != cudaSuccess) { bad = 1; for (int j = 0; j < 4; ++j) printf("   row0 col%d got %g want %g | row1 got %g want %g\n", j, D[j], R[j], D[ncol + j], R[ncol + j]); }
    cudaFree(dA); cudaFree(dB); cudaFree(dD);
  }
  printf(bad ? "PROBE FAILED\n" : "PROBE OK\n");
  return bad;
}
