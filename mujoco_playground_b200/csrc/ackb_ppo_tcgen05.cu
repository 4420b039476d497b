// ackb_ppo_tcgen05.cu -- PPO minibatch gradient on the Blackwell tensor-core path (tcgen05.mma kind::tf32, accumulators in TMEM).
//
// Same contract as ppo_grad_kernel_tc (ackb_ppo.cu; include/ackb_ppo.h: ackb_ppo_minibatch_grad): for a minibatch of the rollout,
// forward of the two 79 -> 64 -> 64 tanh MLPs of SB3's MlpPolicy, PPO loss derivatives per sample, backward, and the weight
// gradients summed over the minibatch (stable_baselines3 PPO.train(), called through model.learn, src/rl/train.py:175-179).
//
// Mapping: one persistent CTA per SM walks its tiles of 128 samples (= UMMA M) once with the policy net and once with the value net
// (the two nets share nothing but the observation rows, whose second read comes from L2).  Thread = (sample row, half of the 64
// columns); a thread's sample row is also its TMEM lane, so everything a thread computes stays in its lane:
//   * operands whose reduction index is the FEATURE (forward and backward-data GEMMs) are read by the tensor core straight from
//     TMEM: the epilogue writes H1 / dZ2 back IN PLACE over the accumulator it just consumed (tcgen05.st), X is copied once from
//     shared memory into TMEM.  No K-major activation copies exist in shared memory.
//   * operands whose reduction index is the SAMPLE (weight-gradient GEMMs) are the same arrays stored [sample][feature] in shared
//     memory in SWIZZLE_128B_BASE32B, the MN-major layout the tensor core accepts for 32-bit operands (umma.cuh; probed in
//     tools/microbench/umma_probe.cu: MN-major tf32 with any other layout type returns zeros).
//   * five GEMM groups per tile, issued by ONE thread (58 tcgen05.mma):
//       (1) Z1  [128 x 64] = X  W1^T        A = X   (TMEM),  B = W1   K-major SW128     -> TMEM ZA
//       (2) Z2  [128 x 64] = H1 W2^T        A = H1  (TMEM, in place of Z1), B = W2      -> TMEM ZB (over X)
//       (3) dH1 [128 x 64] = dZ2 W2         A = dZ2 (TMEM, in place of Z2), B = W2^T    -> TMEM ZA
//       (4) dW2 [ 64 x 65] += dZ2^T [H1|1]  A = dZ2 MN-major, B = [H1 | ones] MN-major  -> TMEM GW2 (kept across tiles)
//       (5) dW1 [ 64 x 80] += dZ1^T X       A = dZ1 MN-major, B = X MN-major            -> TMEM GW1 (kept across tiles)
//     Weight-gradient accumulators leave TMEM once per net; bias gradients come out of the same GEMMs (column 79 of X and the ones
//     block are 1).  (4) and (5) run with M = 128: rows 64..127 of the A operand address the buffer that follows dZ in shared memory
//     and produce accumulator rows that are never read.
//   * epilogues (bias + tanh.approx, tanh', Gaussian / value heads, clipped-surrogate derivatives) on all 256 threads;
//     head-weight gradients (3 x 64) are reduced over the samples of a warp with a transpose-reduction of shuffles.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "ackb.h"
#include "ackb_ppo_common.cuh"
#include "umma.cuh"

namespace ackb_ppo {
namespace {

using namespace umma;

constexpr int TM = 128;          // samples per tile
constexpr int NT5 = 256;
constexpr uint32_t BLK = 16384;  // one SW128 block of 128 rows

// shared-memory map (bytes from a 1024-aligned base)
constexpr uint32_t O_DZ = 0;                 // [2 blocks] BASE32B  dZ2, then dZ1 [sample][64]
constexpr uint32_t O_H1 = O_DZ + 2 * BLK;    // [3 blocks] BASE32B  H1 [sample][64] | block 2: column 0 = 1   (blocks 0, 1 double as rows 64..127 of the padded A operand)
constexpr uint32_t O_X = O_H1 + 3 * BLK;     // [3 blocks] BASE32B  observation tile [sample][80], column 79 = 1 for real samples
constexpr uint32_t O_W1 = O_X + 3 * BLK;     // [3 blocks of 64 rows] K-major SW128  W1 of the current net
constexpr uint32_t O_W2 = O_W1 + 3 * 8192;   // [2 blocks of 64 rows] K-major SW128  W2   [out][in]
constexpr uint32_t O_W2T = O_W2 + 2 * 8192;  // [2 blocks of 64 rows] K-major SW128  W2^T [in][out]
constexpr uint32_t O_F = O_W2T + 2 * 8192;   // floats from here
constexpr int F_B1 = 0, F_B2 = 128, F_W3 = 256, F_B3 = 448, F_LS = 452, F_SC = 456, F_DO = F_SC + TM * 8, F_PART = F_DO + TM * 4,
              F_END = F_PART + TM * 8;
constexpr uint32_t O_BAR = O_F + F_END * 4;      // mbarrier (8 bytes) + TMEM base (4 bytes)
constexpr uint32_t SMEM_BYTES = O_BAR + 16 + 1024;   // + alignment slack
static_assert(SMEM_BYTES <= 232448, "shared memory budget of one SM (227 KB)");

// TMEM columns: ZA = Z1 -> H1 (in place) -> dH1;  ZB = X (80 columns) -> Z2 -> dZ2 (in place);  weight-gradient accumulators of the net
constexpr uint32_t C_ZA = 0, C_ZB = 64, C_GW1 = 144, C_GW2 = 224;

__device__ __forceinline__ float tf32r(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}
__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// read-only global load that the compiler may NOT sink towards its use (volatile asm keeps its place among the other volatile
// statements): the prefetch of the next tile must be issued a whole tile ahead of its consumer
__device__ __forceinline__ float ldg_early(const float* p) {
  float v;
  asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ float4 ldg_early4(const float* p) {
  float4 v;
  asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}

// sum over the 32 lanes of a warp of 16 per-lane values; afterwards lane l holds the total of column
// 8 * bit4(l) + 4 * bit3(l) + 2 * bit2(l) + bit1(l)  (both lanes of a pair hold the same column)
__device__ __forceinline__ float transpose_reduce16(const float (&v)[16], int lane) {
  float a[8], b[4], c[2];
  {
    const bool hi = lane & 16;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float send = hi ? v[i] : v[i + 8], keep = hi ? v[i + 8] : v[i];
      a[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
  }
  {
    const bool hi = lane & 8;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float send = hi ? a[i] : a[i + 4], keep = hi ? a[i + 4] : a[i];
      b[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
  }
  {
    const bool hi = lane & 4;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const float send = hi ? b[i] : b[i + 2], keep = hi ? b[i + 2] : b[i];
      c[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
  }
  float d;
  {
    const bool hi = lane & 2;
    const float send = hi ? c[0] : c[1], keep = hi ? c[1] : c[0];
    d = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  }
  return d + __shfl_xor_sync(0xffffffffu, d, 1);
}
__device__ __forceinline__ int transpose_reduce16_col(int lane) { return ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1); }

#ifdef ACKB_T5_PROFILE
__device__ long long g_t5_prof[16];
#define T5_MARK(i) do { if (blockIdx.x == 0 && threadIdx.x == 0) { const long long c_ = clock64(); g_t5_prof[i] += c_ - t5_last; t5_last = c_; } } while (0)
#else
#define T5_MARK(i) do { } while (0)
#endif

__global__ void __launch_bounds__(NT5, 1) ppo_grad_kernel_tcgen05(PpoArgs a) {
#ifdef ACKB_T5_PROFILE
  long long t5_last = clock64();
#endif
  extern __shared__ unsigned char smem_raw[];
  unsigned char* const sm = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  float* const F = reinterpret_cast<float*>(sm + O_F);
  uint64_t* const mbar = reinterpret_cast<uint64_t*>(sm + O_BAR);
  uint32_t* const tmem_slot = reinterpret_cast<uint32_t*>(sm + O_BAR + 8);
  const uint32_t sb = smem_u32(sm);

  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int q4 = warp & 3, hf = warp >> 2;          // TMEM lane quarter of this warp, column half of this thread
  const int row = 32 * q4 + lane;                   // sample row of the tile = TMEM lane
  const uint32_t rowoff = (uint32_t)((row >> 2) * 512 + (row & 3) * 128);   // BASE32B: 4-row atoms, 32-byte chunks XOR (row & 3)
  const int rx = row & 3;
  const int D = a.D;
  const Offsets o = offsets(D);
  const float* P = a.params;

  // ---- one-time setup: TMEM, mbarrier, ones block, small vectors of both nets
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  if (t == 32) { mbar_init(mbar, 1); mbar_fence_init(); }
  for (int i = t; i < TM * 32; i += NT5) *reinterpret_cast<float*>(sm + O_H1 + 2 * BLK + b32_off(i >> 5, i & 31)) = ((i & 31) == 0) ? 1.0f : 0.0f;
  for (int i = t; i < 3 * H; i += NT5) F[F_W3 + i] = (i < 2 * H) ? P[o.Wa + i] : P[o.Wv + (i - 2 * H)];
  if (t < 128) {
    const int net = t >> 6, r = t & 63;
    F[F_B1 + t] = P[(net ? o.b1v : o.b1p) + r];
    F[F_B2 + t] = P[(net ? o.b2v : o.b2p) + r];
  }
  if (t < 2) { F[F_B3 + t] = P[o.ba + t]; F[F_LS + t] = P[o.ls + t]; }
  if (t == 2) F[F_B3 + 2] = P[o.bv];
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tb = *tmem_slot;
  const uint32_t tlane = tb + ((uint32_t)(32 * q4) << 16);     // this warp's TMEM lane quarter

  const float adv_mean = a.adv_stats[0], adv_istd = 1.0f / (a.adv_stats[1] + 1e-8f);
  const float inv_mb = 1.0f / (float)a.mb;
  const float ls0 = F[F_LS], ls1 = F[F_LS + 1];
  const float iv0 = expf(-2.0f * ls0), iv1 = expf(-2.0f * ls1);

  // per-thread sums over all tiles of this CTA
  float g3[3][2] = {{0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}};   // head weights: [mean0, mean1, value][16-column chunk], column = 32 hf + 16 chunk + col(lane)
  float gb3[3] = {0.f, 0.f, 0.f}, gls0 = 0.f, gls1 = 0.f, d_pg = 0.f, d_vl = 0.f, d_kl = 0.f, d_cf = 0.f;   // threads with hf == 0
  uint32_t phase = 0;
  float* G = a.grads;

  // issue a group of MMAs (thread 0) once every thread's shared-memory / TMEM accesses of the previous phase are done, wait for it
  auto run_mma = [&](auto&& issue) {
    tmem_st_wait();
    fence_proxy_async();
    fence_before_sync();
    __syncthreads();
    if (t == 0) {
      fence_after_sync();
      issue();
      commit(mbar);
    }
    mbar_wait(mbar, phase);
    phase ^= 1u;
    fence_after_sync();
  };

  const int ntiles = (a.mb + TM - 1) / TM;
  // Gather of a tile's observation rows into registers, one tile ahead of its use.  Two mappings (CTA-uniform choice):
  //   vec   (row pitch a multiple of 4 floats and a 16-byte aligned base, e.g. pitch 80): 4 threads per row, 16-byte loads -- thread
  //         (r4 = t >> 2, c = t & 3) holds the float4s c, c + 4, .. c + 16 of rows r4 and r4 + 64;
  //   scalar (dense 79-float rows): 8 threads per row, 4-byte loads -- thread (rq = t >> 3, kk = t & 7) holds columns kk, kk + 8, ..
  //         of rows rq, rq + 32, rq + 64, rq + 96.
  // Threads t < 128 also hold the five scalars of row t.  The loads are issued in four parts spread over the tile body.
  const bool vec = (a.pitch % 4 == 0) && ((reinterpret_cast<uintptr_t>(a.obs) & 15) == 0);
  const int PT = a.pitch;
  float px[40], psc[5];
  int64_t prow[4], prow_sc = 0;          // global row numbers of the tile fetched NEXT (loaded one fetch earlier: no dependent-load stall)
  const int rq = t >> 3, kk = t & 7, r4 = t >> 2, c4i = t & 3;
  auto fetch_rows = [&](int tile) {
    const int sbase = tile * TM, ns = tile < ntiles ? min(TM, a.mb - sbase) : 0;
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      const int r = vec ? (r4 + 64 * (p & 1)) : (rq + 32 * p);
      prow[p] = (r < ns) ? (a.idx ? a.idx[sbase + r] : (int64_t)(sbase + r)) : -1;
    }
    prow_sc = (t < ns) ? (a.idx ? a.idx[sbase + t] : (int64_t)(sbase + t)) : -1;
  };
  auto fetch_part = [&](int p) {
    if (vec) {        // part p: row pass p >> 1, float4s j = 0..2 (even p) or 3..4 (odd p)
      const int pass = p >> 1;
      const bool live = prow[pass] >= 0;
      const float* src = a.obs + (live ? prow[pass] : 0) * (int64_t)PT;
#pragma unroll
      for (int j = 0; j < 5; ++j) {
        if ((j < 3) == ((p & 1) == 0)) {
          const int f = 4 * (c4i + 4 * j);
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (live && f < PT) v = ldg_early4(src + f);
          float* d = &px[20 * pass + 4 * j];
          d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
        }
      }
    } else {
      const bool live = prow[p] >= 0;
      const float* src = a.obs + (live ? prow[p] : 0) * (int64_t)PT;
#pragma unroll
      for (int j = 0; j < KP / 8; ++j) {
        const int k = kk + 8 * j;
        px[10 * p + j] = (live && k < D) ? ldg_early(src + k) : 0.0f;
      }
    }
  };
  auto fetch_scalars = [&]() {
    if (prow_sc >= 0) {
      const int64_t grow = prow_sc;
      psc[0] = ldg_early(&a.act[grow * 2]); psc[1] = ldg_early(&a.act[grow * 2 + 1]); psc[2] = ldg_early(&a.old_logp[grow]);
      psc[3] = ldg_early(&a.adv[grow]); psc[4] = ldg_early(&a.ret[grow]);
    }
  };
  auto fetch_tile = [&](int tile) {      // everything at once (first tile of a pass)
#pragma unroll
    for (int p = 0; p < 4; ++p) fetch_part(p);
    fetch_scalars();
    fetch_rows(tile + gridDim.x);        // row numbers of the tile after this one
  };
  // registers -> BASE32B operand layout, rounded to TF32; columns >= D are zero except column 79 = 1 on real samples (bias gradient)
  auto store_tile = [&](int ns) {
    if (vec) {
#pragma unroll
      for (int pass = 0; pass < 2; ++pass) {
        const int r = r4 + 64 * pass;
        const float one = (r < ns) ? 1.0f : 0.0f;
        const uint32_t xrow = O_X + (uint32_t)((r >> 2) * 512 + (r & 3) * 128);
#pragma unroll
        for (int j = 0; j < 5; ++j) {
          const int f = 4 * (c4i + 4 * j);
          float v[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) v[i] = (f + i < D) ? tf32r(px[20 * pass + 4 * j + i]) : ((f + i == KP - 1) ? one : 0.0f);
          *reinterpret_cast<float4*>(sm + xrow + (uint32_t)(f >> 5) * BLK + (uint32_t)((((((f & 31) >> 3) ^ (r & 3)) & 3) << 5) + (f & 7) * 4)) =
              make_float4(v[0], v[1], v[2], v[3]);
        }
      }
    } else {
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        const int r = rq + 32 * p;
        const float one = (r < ns) ? 1.0f : 0.0f;
        const uint32_t xrow = O_X + (uint32_t)((r >> 2) * 512 + (r & 3) * 128);
#pragma unroll
        for (int j = 0; j < KP / 8; ++j) {
          const int k = kk + 8 * j;     // feature k: block k / 32, 32-byte chunk (k % 32) / 8 = j % 4, position kk inside the chunk
          const float v = (k < D) ? tf32r(px[10 * p + j]) : ((k == KP - 1) ? one : 0.0f);
          *reinterpret_cast<float*>(sm + xrow + (uint32_t)(k >> 5) * BLK + (uint32_t)(((((j & 3) ^ (r & 3)) & 3) << 5) + kk * 4)) = v;
        }
      }
    }
    if (t < TM && t < ns) {
#pragma unroll
      for (int i = 0; i < 5; ++i) F[F_SC + t * 8 + i] = psc[i];
    }
  };
#pragma unroll 1
  for (int net = 0; net < 2; ++net) {
    // ---- first tile of this pass on its way, then the weights of this net (TF32-rounded) in the K-major operand layout:
    // W1 [64][80], W2 [out][in], W2^T [in][out]
    __syncthreads();
    bool first = true;
    fetch_rows(blockIdx.x);
    if ((int)blockIdx.x < ntiles) fetch_tile(blockIdx.x);
    {
      const float* W1g = P + (net ? o.W1v : o.W1p);
      const float* W2g = P + (net ? o.W2v : o.W2p);
      // W1: thread t owns row r = t / 4 and columns k = (t % 4) + 4 j: consecutive threads read consecutive words
      {
        const int r = t >> 2, k0 = t & 3;
#pragma unroll 5
        for (int j = 0; j < KP / 4; ++j) {
          const int k = k0 + 4 * j;
          const float w = (k < D) ? tf32r(__ldg(W1g + r * D + k)) : 0.0f;
          *reinterpret_cast<float*>(sm + O_W1 + (k >> 5) * 8192 + sw128_off(r, k & 31)) = w;
        }
      }
#pragma unroll 4
      for (int i = t; i < H * H; i += NT5) {
        const int r = i >> 6, k = i & 63;
        const float w = tf32r(__ldg(W2g + i));
        *reinterpret_cast<float*>(sm + O_W2 + (k >> 5) * 8192 + sw128_off(r, k & 31)) = w;
        *reinterpret_cast<float*>(sm + O_W2T + (r >> 5) * 8192 + sw128_off(k, r & 31)) = w;
      }
    }
    T5_MARK(10);
#pragma unroll 1
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int sbase = tile * TM;
      const int ns = min(TM, a.mb - sbase);
      // ---- this tile's observation rows and scalars were fetched into registers one tile ahead (global latency hidden behind the
      // previous tile's compute): round to TF32 and store into the BASE32B operand layout, then fetch the next tile
      store_tile(ns);
      const bool more = tile + (int)gridDim.x < ntiles;      // CTA-uniform
      if (more) { fetch_part(0); fetch_scalars(); }
      __syncthreads();
      T5_MARK(0);
      // ---- X -> TMEM (A operand of (1)): this thread's row, features 40 hf .. 40 hf + 39, into columns C_ZB + 40 hf ..
      {
        float v[16], w8[8];
#pragma unroll
        for (int part = 0; part < 3; ++part) {
          const int nv = part < 2 ? 4 : 2;      // 16 + 16 + 8 features
#pragma unroll
          for (int c4 = 0; c4 < 4; ++c4) {
            if (c4 < nv) {
              const int f = 40 * hf + 16 * part + 4 * c4;
              const float4 x = *reinterpret_cast<const float4*>(sm + O_X + (uint32_t)(f >> 5) * BLK + rowoff + (uint32_t)((((((f & 31) >> 3) ^ rx) & 3) << 5) + (f & 7) * 4));
              if (part < 2) { v[4 * c4] = x.x; v[4 * c4 + 1] = x.y; v[4 * c4 + 2] = x.z; v[4 * c4 + 3] = x.w; }
              else { w8[4 * c4] = x.x; w8[4 * c4 + 1] = x.y; w8[4 * c4 + 2] = x.z; w8[4 * c4 + 3] = x.w; }
            }
          }
          if (part < 2) tmem_st16(tlane + C_ZB + 40 * hf + 16 * part, v);
          else tmem_st8(tlane + C_ZB + 40 * hf + 32, w8);
        }
      }
      // ---- (1) Z1 = X W1^T
      run_mma([&] {
        const uint32_t id = idesc_tf32(128, 64, 0, 0);
#pragma unroll
        for (int k = 0; k < KP / 8; ++k) mma_tf32_ts(tb + C_ZA, tb + C_ZB + 8 * k, desc_kmajor(sb + O_W1 + (k >> 2) * 8192, k & 3), id, k > 0);
      });
      T5_MARK(1);
      // ---- H1 = tanh(Z1 + b1): back into TMEM in place (A operand of (2)) and into shared memory (B operand of (4), tanh' later)
#pragma unroll
      for (int ch = 0; ch < 2; ++ch) {
        float v[16];
        tmem_ld16(tlane + C_ZA + 32 * hf + 16 * ch, v);
        const float* b1 = &F[F_B1 + 64 * net + 32 * hf + 16 * ch];
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = tf32r(tanh_fast(v[j] + b1[j]));
        tmem_st16(tlane + C_ZA + 32 * hf + 16 * ch, v);
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) {
          const int cc = 4 * ch + c4;           // 16-byte vector index inside the 128-byte row of block hf
          *reinterpret_cast<float4*>(sm + O_H1 + hf * BLK + rowoff + ((((cc >> 1) ^ rx) & 3) << 5) + (cc & 1) * 16) = make_float4(v[4 * c4], v[4 * c4 + 1], v[4 * c4 + 2], v[4 * c4 + 3]);
        }
      }
      if (more) fetch_part(1);
      T5_MARK(2);
      // ---- (2) Z2 = H1 W2^T
      run_mma([&] {
        const uint32_t id = idesc_tf32(128, 64, 0, 0);
#pragma unroll
        for (int k = 0; k < 8; ++k) mma_tf32_ts(tb + C_ZB, tb + C_ZA + 8 * k, desc_kmajor(sb + O_W2 + (k >> 2) * 8192, k & 3), id, k > 0);
      });
      T5_MARK(3);
      // ---- heads: H2 = tanh(Z2 + b2) of this thread's 32 columns (kept in registers), partial dot products with the head weights
      float h2v[32];
      {
        float p0 = 0.f, p1 = 0.f;
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
          float v[16];
          tmem_ld16(tlane + C_ZB + 32 * hf + 16 * ch, v);
          const int c0 = 32 * hf + 16 * ch;
          const float* b2 = &F[F_B2 + 64 * net + c0];
          const float* wa = &F[F_W3 + (net ? 2 * H : 0) + c0];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float h2 = tanh_fast(v[j] + b2[j]);
            h2v[16 * ch + j] = h2;
            p0 = fmaf(h2, wa[j], p0);
            if (net == 0) p1 = fmaf(h2, wa[H + j], p1);
          }
        }
        F[F_PART + row * 8 + hf * 4] = p0; F[F_PART + row * 8 + hf * 4 + 1] = p1;
      }
      if (more) fetch_part(2);
      __syncthreads();
      T5_MARK(4);
      // ---- PPO loss derivatives, one thread per sample (hf == 0)
      if (hf == 0) {
        float d0 = 0.f, d1 = 0.f;
        if (row < ns) {
          const float* SC = &F[F_SC + row * 8];
          const float s0 = F[F_PART + row * 8] + F[F_PART + row * 8 + 4], s1 = F[F_PART + row * 8 + 1] + F[F_PART + row * 8 + 5];
          if (net == 0) {
            const float m0 = s0 + F[F_B3], m1 = s1 + F[F_B3 + 1];
            const float e0 = SC[0] - m0, e1 = SC[1] - m1;
            const float q0 = e0 * e0 * iv0, q1 = e1 * e1 * iv1;
            const float logp = -0.5f * q0 - ls0 - 0.9189385332046727f - 0.5f * q1 - ls1 - 0.9189385332046727f;
            const float A_ = (SC[3] - adv_mean) * adv_istd;
            const float lr = logp - SC[2];
            const float r = expf(lr);
            const float rc = fminf(fmaxf(r, 1.0f - a.clip), 1.0f + a.clip);
            const float u1 = A_ * r, u2 = A_ * rc;
            const float dlogp = (u1 <= u2) ? -A_ * r : 0.0f;      // d(-min(u1, u2)) / d logp (the clipped branch has zero slope)
            d0 = dlogp * e0 * iv0 * inv_mb; d1 = dlogp * e1 * iv1 * inv_mb;
            gls0 += dlogp * (q0 - 1.0f) * inv_mb; gls1 += dlogp * (q1 - 1.0f) * inv_mb;
            d_pg += -fminf(u1, u2); d_kl += (r - 1.0f) - lr; d_cf += (fabsf(r - 1.0f) > a.clip) ? 1.0f : 0.0f;
            gb3[0] += d0; gb3[1] += d1;
          } else {
            const float v = s0 + F[F_B3 + 2], R = SC[4];
            d0 = a.vf_coef * 2.0f * (v - R) * inv_mb;
            d_vl += (v - R) * (v - R);
            gb3[2] += d0;
          }
        }
        F[F_DO + row * 4] = d0; F[F_DO + row * 4 + 1] = d1;
      }
      __syncthreads();
      T5_MARK(5);
      // ---- dZ2 = (dOut W3) (1 - H2^2): into TMEM in place (A operand of (3)) and shared memory (A operand of (4));
      //      head-weight gradients dOut^T H2 by transpose-reduction over the warp's rows
      {
        const float d0 = F[F_DO + row * 4], d1 = F[F_DO + row * 4 + 1];
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
          float v[16], ga[16], gb[16];
          const int c0 = 32 * hf + 16 * ch;
          const float* wa = &F[F_W3 + (net ? 2 * H : 0) + c0];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float h2 = h2v[16 * ch + j];
            ga[j] = d0 * h2;
            gb[j] = d1 * h2;
            const float g = (net == 0) ? fmaf(d1, wa[H + j], d0 * wa[j]) : d0 * wa[j];
            v[j] = tf32r(g * (1.0f - h2 * h2));
          }
          tmem_st16(tlane + C_ZB + 32 * hf + 16 * ch, v);
#pragma unroll
          for (int c4 = 0; c4 < 4; ++c4) {
            const int cc = 4 * ch + c4;
            *reinterpret_cast<float4*>(sm + O_DZ + hf * BLK + rowoff + ((((cc >> 1) ^ rx) & 3) << 5) + (cc & 1) * 16) = make_float4(v[4 * c4], v[4 * c4 + 1], v[4 * c4 + 2], v[4 * c4 + 3]);
          }
          const float ra = transpose_reduce16(ga, lane);
          if (net == 0) {
            g3[0][ch] += ra;
            g3[1][ch] += transpose_reduce16(gb, lane);
          } else g3[2][ch] += ra;
        }
      }
      if (more) { fetch_part(3); fetch_rows(tile + 2 * gridDim.x); }
      T5_MARK(6);
      // ---- (3) dH1 = dZ2 W2   and   (4) dW2 += dZ2^T [H1 | 1]
      run_mma([&] {
        const uint32_t id3 = idesc_tf32(128, 64, 0, 0);
#pragma unroll
        for (int k = 0; k < 8; ++k) mma_tf32_ts(tb + C_ZA, tb + C_ZB + 8 * k, desc_kmajor(sb + O_W2T + (k >> 2) * 8192, k & 3), id3, k > 0);
        const uint32_t id4 = idesc_tf32(128, 80, 1, 1);
#pragma unroll
        for (int k = 0; k < TM / 8; ++k) mma_tf32(tb + C_GW2, desc_mn32(sb + O_DZ, BLK, k), desc_mn32(sb + O_H1, BLK, k), id4, !first || k > 0);
      });
      T5_MARK(7);
      // ---- dZ1 = dH1 (1 - H1^2) -> shared memory (overwrites dZ2: (3) and (4) have completed)
#pragma unroll
      for (int ch = 0; ch < 2; ++ch) {
        float v[16];
        tmem_ld16(tlane + C_ZA + 32 * hf + 16 * ch, v);
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) {
          const int cc = 4 * ch + c4;
          const uint32_t off = hf * BLK + rowoff + ((((cc >> 1) ^ rx) & 3) << 5) + (cc & 1) * 16;
          const float4 h = *reinterpret_cast<const float4*>(sm + O_H1 + off);
          *reinterpret_cast<float4*>(sm + O_DZ + off) = make_float4(tf32r(v[4 * c4] * (1.0f - h.x * h.x)), tf32r(v[4 * c4 + 1] * (1.0f - h.y * h.y)),
                                                                    tf32r(v[4 * c4 + 2] * (1.0f - h.z * h.z)), tf32r(v[4 * c4 + 3] * (1.0f - h.w * h.w)));
        }
      }
      T5_MARK(8);
      // ---- (5) dW1 += dZ1^T X
      run_mma([&] {
        const uint32_t id5 = idesc_tf32(128, 80, 1, 1);
#pragma unroll
        for (int k = 0; k < TM / 8; ++k) mma_tf32(tb + C_GW1, desc_mn32(sb + O_DZ, BLK, k), desc_mn32(sb + O_X, BLK, k), id5, !first || k > 0);
      });
      T5_MARK(9);
      first = false;
    }

    // ---- weight-gradient accumulators of this net: TMEM -> global (rows 0..63 of each accumulator are real)
    if (!first && q4 < 2) {
      const int i = row;      // output row of the layer (0..63)
#pragma unroll 1
      for (int ch = hf; ch < 5; ch += 2) {
        float v[16];
        tmem_ld16(tlane + C_GW1 + 16 * ch, v);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int k = 16 * ch + j;
          if (k < D) atomicAdd(&G[(net ? o.W1v : o.W1p) + i * D + k], v[j]);
          else if (k == KP - 1) atomicAdd(&G[(net ? o.b1v : o.b1p) + i], v[j]);
        }
        tmem_ld16(tlane + C_GW2 + 16 * ch, v);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int k = 16 * ch + j;
          if (k < H) atomicAdd(&G[(net ? o.W2v : o.W2p) + i * H + k], v[j]);
          else if (k == H) atomicAdd(&G[(net ? o.b2v : o.b2p) + i], v[j]);
        }
      }
    }
    T5_MARK(11);
    fence_before_sync();     // the accumulator reads above are ordered before the next net's MMAs by the barrier at the loop top
  }

  if ((lane & 1) == 0) {   // head weights: even lanes hold the column totals of their warp's 32 rows
    const int col = transpose_reduce16_col(lane);
#pragma unroll
    for (int ch = 0; ch < 2; ++ch) {
      const int n = 32 * hf + 16 * ch + col;
      atomicAdd(&G[o.Wa + n], g3[0][ch]);
      atomicAdd(&G[o.Wa + H + n], g3[1][ch]);
      atomicAdd(&G[o.Wv + n], g3[2][ch]);
    }
  }
  if (hf == 0) {           // per-sample sums: one value per thread of warps 0..3
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      gls0 += __shfl_xor_sync(0xffffffffu, gls0, off); gls1 += __shfl_xor_sync(0xffffffffu, gls1, off);
      d_pg += __shfl_xor_sync(0xffffffffu, d_pg, off); d_vl += __shfl_xor_sync(0xffffffffu, d_vl, off);
      d_kl += __shfl_xor_sync(0xffffffffu, d_kl, off); d_cf += __shfl_xor_sync(0xffffffffu, d_cf, off);
      gb3[0] += __shfl_xor_sync(0xffffffffu, gb3[0], off); gb3[1] += __shfl_xor_sync(0xffffffffu, gb3[1], off);
      gb3[2] += __shfl_xor_sync(0xffffffffu, gb3[2], off);
    }
    if (lane == 0) {
      atomicAdd(&G[o.ls], gls0); atomicAdd(&G[o.ls + 1], gls1);
      atomicAdd(&G[o.ba], gb3[0]); atomicAdd(&G[o.ba + 1], gb3[1]); atomicAdd(&G[o.bv], gb3[2]);
      atomicAdd(&a.diag[0], d_pg * inv_mb); atomicAdd(&a.diag[1], d_vl * inv_mb);
      atomicAdd(&a.diag[3], d_kl * inv_mb); atomicAdd(&a.diag[4], d_cf * inv_mb);
    }
  }
  if (blockIdx.x == 0 && t == 0) {
    // entropy of the state-independent Gaussian: sum_j (0.5 + 0.5 log 2 pi + log_std_j); -ent_coef * H enters the loss
    a.diag[2] = 2.0f * 1.4189385332046727f + ls0 + ls1;
    atomicAdd(&G[o.ls], -a.ent_coef); atomicAdd(&G[o.ls + 1], -a.ent_coef);
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tb, 512);
}

}  // namespace

#ifdef ACKB_T5_PROFILE
extern "C" int ackb_ppo_t5_profile(long long* out16) {
  cudaDeviceSynchronize();
  if (cudaMemcpyFromSymbol(out16, g_t5_prof, sizeof(long long) * 16) != cudaSuccess) return -1;
  long long z[16] = {0};
  cudaMemcpyToSymbol(g_t5_prof, z, sizeof z);
  return 0;
}
#endif

int launch_grad_tcgen05(const PpoArgs& a, cudaStream_t stream) {
  if (a.D >= KP) return ACKB_ERR_ARG;     // the bias gradient of layer 1 rides in column KP - 1 of the observation tile
  static bool attr_done[64] = {false};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return ACKB_ERR_NO_DEVICE;
  if (dev < 64 && !attr_done[dev]) {
    if (cudaFuncSetAttribute(ppo_grad_kernel_tcgen05, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES) != cudaSuccess) return ACKB_ERR_CUDA;
    attr_done[dev] = true;
  }
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int ntiles = (a.mb + TM - 1) / TM;
  ppo_grad_kernel_tcgen05<<<ntiles < sms ? ntiles : sms, NT5, SMEM_BYTES, stream>>>(a);
  return cudaGetLastError() == cudaSuccess ? ACKB_OK : ACKB_ERR_CUDA;
}

}  // namespace ackb_ppo
