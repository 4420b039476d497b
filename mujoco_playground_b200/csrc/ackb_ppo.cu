// ackb_ppo.cu -- fused PPO minibatch gradient for the reference's MlpPolicy (include/ackb_ppo.h).
//
// One persistent CTA per SM walks tiles of 64 samples.  Weights (75 KB) and the tile's activations stay in shared memory:
// forward (two 79->64->64 tanh MLPs, Gaussian head, value head), PPO loss derivatives per sample, backward through both MLPs
// and the weight-gradient accumulation (in registers, one fixed slice of every weight matrix per thread) happen without the
// activations ever touching HBM; the only per-sample HBM traffic is the gathered observation row (316 B) and 20 B of scalars.
// The eager torch path moves ~10 KB per sample through HBM for the same arithmetic.  CUDA-core fp32 FMAs (the matrices are
// 64 wide), gradients agree with torch autograd to fp32 rounding (tests/test_ppo.py).
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

#include "ackb.h"
#include "ackb_ppo.h"
#include "ackb_ppo_common.cuh"

using namespace ackb_ppo;

namespace {

constexpr int TS = 64;       // samples per tile (weights 105 KB + activations 117 KB fill the 227 KB of an SM)
constexpr int NT = 256;      // threads per CTA

// shared-memory layout (floats)
constexpr int S_W1T = 0;                       // [KP][128]   k-major, n: 0..63 policy, 64..127 value
constexpr int S_W2T = S_W1T + KP * 128;        // [2][64][64] [net][k][n]
constexpr int S_W2 = S_W2T + 2 * H * H;        // [2][64][64] [net][n][k]
constexpr int S_W3 = S_W2 + 2 * H * H;         // [3][64]     mean0, mean1, value
constexpr int S_B1 = S_W3 + 3 * H;             // [128]
constexpr int S_B2 = S_B1 + 128;               // [128]
constexpr int S_B3 = S_B2 + 128;               // [4]  (ba0, ba1, bv, -)
constexpr int S_LS = S_B3 + 4;                 // [4]  (log_std0, log_std1, -, -)
constexpr int S_X = S_LS + 4;                  // [TS][KP]
constexpr int S_H1 = S_X + TS * KP;            // [TS][128]
constexpr int S_H2 = S_H1 + TS * 128;          // [TS][128]
constexpr int S_DH = S_H2 + TS * 128;          // [TS][128]   dH2, then dH1
constexpr int S_DO = S_DH + TS * 128;          // [TS][4]     d loss / d (mean0, mean1, value)
constexpr int S_TOTAL = S_DO + TS * 4;

__global__ void __launch_bounds__(NT, 1) ppo_grad_kernel(PpoArgs a) {
  extern __shared__ __align__(16) float sm[];
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int D = a.D;
  const Offsets o = offsets(D);
  const float* P = a.params;

  // ---- weights into shared memory (transposed copies where the contraction index must come first)
  for (int i = t; i < KP * 128; i += NT) {
    const int k = i >> 7, n = i & 127, net = n >> 6, r = n & 63;
    sm[S_W1T + i] = (k < D) ? P[(net ? o.W1v : o.W1p) + r * D + k] : 0.0f;
  }
  for (int i = t; i < 2 * H * H; i += NT) {
    const int net = i >> 12, n = (i >> 6) & 63, k = i & 63;
    const float w = P[(net ? o.W2v : o.W2p) + n * H + k];
    sm[S_W2 + i] = w;
    sm[S_W2T + (net << 12) + k * H + n] = w;
  }
  for (int i = t; i < 3 * H; i += NT) sm[S_W3 + i] = (i < 2 * H) ? P[o.Wa + i] : P[o.Wv + (i - 2 * H)];
  if (t < 128) {
    const int net = t >> 6, r = t & 63;
    sm[S_B1 + t] = P[(net ? o.b1v : o.b1p) + r];
    sm[S_B2 + t] = P[(net ? o.b2v : o.b2p) + r];
  }
  if (t < 2) { sm[S_B3 + t] = P[o.ba + t]; sm[S_LS + t] = P[o.ls + t]; }
  if (t == 2) sm[S_B3 + 2] = P[o.bv];
  __syncthreads();

  const float adv_mean = a.adv_stats[0], adv_istd = 1.0f / (a.adv_stats[1] + 1e-8f);
  const float inv_mb = 1.0f / (float)a.mb;
  const float ls0 = sm[S_LS], ls1 = sm[S_LS + 1];
  const float iv0 = expf(-2.0f * ls0), iv1 = expf(-2.0f * ls1);

  // ---- per-thread slices of the weight gradients, accumulated over all tiles of this CTA
  float g1[8][5];      // dW1[n][k]: n = (t >> 4) * 8 + i (both nets, 0..127), k = (t & 15) * 5 + j
  float g2[4][8];      // dW2[net][n][k]: net = t >> 7, n = ((t & 127) >> 3) * 4 + i, k = (t & 7) * 8 + j
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 5; ++j) g1[i][j] = 0.0f;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) g2[i][j] = 0.0f;
  float g3 = 0.0f, gb1 = 0.0f, gb2 = 0.0f, gb3 = 0.0f;   // dW3 (t < 192), db1 / db2 (t < 128), db3 (t < 3)
  float gls0 = 0.0f, gls1 = 0.0f;                         // d log_std (sample-leader lanes)
  float d_pg = 0.0f, d_vl = 0.0f, d_kl = 0.0f, d_cf = 0.0f;

  // GEMM thread mapping: 4 samples x 8 outputs per thread
  constexpr int MS = 4;
  const int ng = t & 15, n0 = ng * 8, s0 = (t >> 4) * MS;
  const int gnet = n0 >> 6, nn0 = n0 & 63;
  // dW2 mapping
  const int w2net = t >> 7, w2n0 = ((t & 127) >> 3) * 4, w2k0 = (t & 7) * 8;
  // dW1 mapping
  const int w1n0 = (t >> 4) * 8, w1k0 = (t & 15) * 5;

  const int ntiles = (a.mb + TS - 1) / TS;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int sbase = tile * TS;
    const int ns = min(TS, a.mb - sbase);
    // ---- 1. gather the observation rows of the tile
    for (int i = t; i < TS * KP; i += NT) {
      const int s = i / KP, k = i - s * KP;
      float v = 0.0f;
      if (s < ns && k < D) {
        const int64_t row = a.idx ? a.idx[sbase + s] : (int64_t)(sbase + s);
        v = a.obs[row * a.pitch + k];
      }
      sm[S_X + i] = v;
    }
    __syncthreads();
    // ---- 2. layer 1 of both nets: H1 = tanh(X W1^T + b1)
    {
      float acc[MS][8];
#pragma unroll
      for (int i = 0; i < MS; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;
#pragma unroll 2
      for (int k = 0; k < D; ++k) {
        float av[MS];
#pragma unroll
        for (int i = 0; i < MS; ++i) av[i] = sm[S_X + (s0 + i) * KP + k];
        const float4 wa = *reinterpret_cast<const float4*>(&sm[S_W1T + k * 128 + n0]);
        const float4 wb = *reinterpret_cast<const float4*>(&sm[S_W1T + k * 128 + n0 + 4]);
        const float w[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
        for (int i = 0; i < MS; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], w[j], acc[i][j]);
      }
#pragma unroll
      for (int i = 0; i < MS; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) sm[S_H1 + (s0 + i) * 128 + n0 + j] = tanhf(acc[i][j] + sm[S_B1 + n0 + j]);
    }
    __syncthreads();
    // ---- 3. layer 2 (block diagonal: each net reads its own 64 columns of H1)
    {
      float acc[MS][8];
#pragma unroll
      for (int i = 0; i < MS; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;
#pragma unroll 4
      for (int k = 0; k < H; ++k) {
        float av[MS];
#pragma unroll
        for (int i = 0; i < MS; ++i) av[i] = sm[S_H1 + (s0 + i) * 128 + gnet * H + k];
        const float4 wa = *reinterpret_cast<const float4*>(&sm[S_W2T + (gnet << 12) + k * H + nn0]);
        const float4 wb = *reinterpret_cast<const float4*>(&sm[S_W2T + (gnet << 12) + k * H + nn0 + 4]);
        const float w[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
        for (int i = 0; i < MS; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], w[j], acc[i][j]);
      }
#pragma unroll
      for (int i = 0; i < MS; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) sm[S_H2 + (s0 + i) * 128 + n0 + j] = tanhf(acc[i][j] + sm[S_B2 + n0 + j]);
    }
    __syncthreads();
    // ---- 4. heads + PPO loss derivatives: warp w owns samples (TS/8) w .. (TS/8) w + TS/8 - 1
#pragma unroll 1
    for (int si = 0; si < TS / 8; ++si) {
      const int s = warp * (TS / 8) + si;
      const float* h2 = &sm[S_H2 + s * 128];
      float p0 = h2[lane] * sm[S_W3 + lane] + h2[lane + 32] * sm[S_W3 + lane + 32];
      float p1 = h2[lane] * sm[S_W3 + H + lane] + h2[lane + 32] * sm[S_W3 + H + lane + 32];
      float pv = h2[H + lane] * sm[S_W3 + 2 * H + lane] + h2[H + lane + 32] * sm[S_W3 + 2 * H + lane + 32];
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        p0 += __shfl_xor_sync(0xffffffffu, p0, off); p1 += __shfl_xor_sync(0xffffffffu, p1, off); pv += __shfl_xor_sync(0xffffffffu, pv, off);
      }
      if (lane == 0) {
        float dm0 = 0.0f, dm1 = 0.0f, dv = 0.0f;
        if (s < ns) {
          const int64_t row = a.idx ? a.idx[sbase + s] : (int64_t)(sbase + s);
          const float m0 = p0 + sm[S_B3], m1 = p1 + sm[S_B3 + 1], v = pv + sm[S_B3 + 2];
          const float e0 = a.act[row * 2] - m0, e1 = a.act[row * 2 + 1] - m1;
          const float q0 = e0 * e0 * iv0, q1 = e1 * e1 * iv1;
          const float logp = -0.5f * q0 - ls0 - 0.9189385332046727f - 0.5f * q1 - ls1 - 0.9189385332046727f;
          const float A = (a.adv[row] - adv_mean) * adv_istd;
          const float lr = logp - a.old_logp[row];
          const float r = expf(lr);
          const float rc = fminf(fmaxf(r, 1.0f - a.clip), 1.0f + a.clip);
          const float s1 = A * r, s2 = A * rc;
          const float dlogp = (s1 <= s2) ? -A * r : 0.0f;      // d(-min(s1, s2)) / d logp (clipped branch has zero slope)
          dm0 = dlogp * e0 * iv0 * inv_mb; dm1 = dlogp * e1 * iv1 * inv_mb;
          gls0 += dlogp * (q0 - 1.0f) * inv_mb; gls1 += dlogp * (q1 - 1.0f) * inv_mb;
          const float R = a.ret[row];
          dv = a.vf_coef * 2.0f * (v - R) * inv_mb;
          d_pg += -fminf(s1, s2); d_vl += (v - R) * (v - R); d_kl += (r - 1.0f) - lr; d_cf += (fabsf(r - 1.0f) > a.clip) ? 1.0f : 0.0f;
        }
        sm[S_DO + s * 4] = dm0; sm[S_DO + s * 4 + 1] = dm1; sm[S_DO + s * 4 + 2] = dv; sm[S_DO + s * 4 + 3] = 0.0f;
      }
    }
    __syncthreads();
    // ---- 5. dH2 = (dOut W3) * (1 - H2^2)
    for (int i = t; i < TS * 128; i += NT) {
      const int s = i >> 7, n = i & 127;
      const float h = sm[S_H2 + i];
      const float g = (n < H) ? sm[S_DO + s * 4] * sm[S_W3 + n] + sm[S_DO + s * 4 + 1] * sm[S_W3 + H + n]
                              : sm[S_DO + s * 4 + 2] * sm[S_W3 + 2 * H + (n - H)];
      sm[S_DH + i] = g * (1.0f - h * h);
    }
    __syncthreads();
    // ---- 6. weight gradients of the heads and of layer 2, bias gradients
    if (t < 3 * H) {
      const int r = t >> 6, c = t & 63, hoff = (r < 2 ? 0 : H) + c;
#pragma unroll 4
      for (int s = 0; s < TS; ++s) g3 = fmaf(sm[S_DO + s * 4 + r], sm[S_H2 + s * 128 + hoff], g3);
    }
    if (t < 3)
      for (int s = 0; s < TS; ++s) gb3 += sm[S_DO + s * 4 + t];
    if (t < 128)
      for (int s = 0; s < TS; ++s) gb2 += sm[S_DH + s * 128 + t];
#pragma unroll 4
    for (int s = 0; s < TS; ++s) {
      const float4 dn = *reinterpret_cast<const float4*>(&sm[S_DH + s * 128 + w2net * H + w2n0]);
      const float4 ha = *reinterpret_cast<const float4*>(&sm[S_H1 + s * 128 + w2net * H + w2k0]);
      const float4 hb = *reinterpret_cast<const float4*>(&sm[S_H1 + s * 128 + w2net * H + w2k0 + 4]);
      const float d[4] = {dn.x, dn.y, dn.z, dn.w};
      const float hk[8] = {ha.x, ha.y, ha.z, ha.w, hb.x, hb.y, hb.z, hb.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) g2[i][j] = fmaf(d[i], hk[j], g2[i][j]);
    }
    // ---- 7. dH1 = (dH2 W2) * (1 - H1^2)
    {
      float acc[MS][8];
#pragma unroll
      for (int i = 0; i < MS; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;
#pragma unroll 4
      for (int n = 0; n < H; ++n) {
        float av[MS];
#pragma unroll
        for (int i = 0; i < MS; ++i) av[i] = sm[S_DH + (s0 + i) * 128 + gnet * H + n];
        const float4 wa = *reinterpret_cast<const float4*>(&sm[S_W2 + (gnet << 12) + n * H + nn0]);
        const float4 wb = *reinterpret_cast<const float4*>(&sm[S_W2 + (gnet << 12) + n * H + nn0 + 4]);
        const float w[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
        for (int i = 0; i < MS; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], w[j], acc[i][j]);
      }
      __syncthreads();   // every read of dH2 (steps 6 and 7) is done: the buffer now takes dH1
#pragma unroll
      for (int i = 0; i < MS; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float h = sm[S_H1 + (s0 + i) * 128 + n0 + j];
          sm[S_DH + (s0 + i) * 128 + n0 + j] = acc[i][j] * (1.0f - h * h);
        }
    }
    __syncthreads();
    // ---- 8. weight gradient of layer 1 and its bias
    if (t < 128)
      for (int s = 0; s < TS; ++s) gb1 += sm[S_DH + s * 128 + t];
#pragma unroll 4
    for (int s = 0; s < TS; ++s) {
      const float4 da = *reinterpret_cast<const float4*>(&sm[S_DH + s * 128 + w1n0]);
      const float4 db = *reinterpret_cast<const float4*>(&sm[S_DH + s * 128 + w1n0 + 4]);
      const float d[8] = {da.x, da.y, da.z, da.w, db.x, db.y, db.z, db.w};
      float xk[5];
#pragma unroll
      for (int j = 0; j < 5; ++j) xk[j] = sm[S_X + s * KP + w1k0 + j];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 5; ++j) g1[i][j] = fmaf(d[i], xk[j], g1[i][j]);
    }
    __syncthreads();
  }

  // ---- flush this CTA's gradient slices
  float* G = a.grads;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int n = w1n0 + i, net = n >> 6, r = n & 63;
#pragma unroll
    for (int j = 0; j < 5; ++j) {
      const int k = w1k0 + j;
      if (k < D) atomicAdd(&G[(net ? o.W1v : o.W1p) + r * D + k], g1[i][j]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) atomicAdd(&G[(w2net ? o.W2v : o.W2p) + (w2n0 + i) * H + w2k0 + j], g2[i][j]);
  if (t < 3 * H) {
    const int r = t >> 6, c = t & 63;
    atomicAdd(&G[r < 2 ? o.Wa + r * H + c : o.Wv + c], g3);
  }
  if (t < 128) {
    const int net = t >> 6, r = t & 63;
    atomicAdd(&G[(net ? o.b1v : o.b1p) + r], gb1);
    atomicAdd(&G[(net ? o.b2v : o.b2p) + r], gb2);
  }
  if (t < 2) atomicAdd(&G[o.ba + t], gb3);
  if (t == 2) atomicAdd(&G[o.bv], gb3);
  // per-warp reduction of the sample-leader sums (only lane 0 of each warp holds values)
  if (lane == 0) {
    atomicAdd(&G[o.ls], gls0); atomicAdd(&G[o.ls + 1], gls1);
    atomicAdd(&a.diag[0], d_pg * inv_mb); atomicAdd(&a.diag[1], d_vl * inv_mb);
    atomicAdd(&a.diag[3], d_kl * inv_mb); atomicAdd(&a.diag[4], d_cf * inv_mb);
  }
  if (blockIdx.x == 0 && t == 0) {
    // entropy of the state-independent Gaussian: sum_j (0.5 + 0.5 log 2 pi + log_std_j); -ent_coef * H enters the loss
    if (a.diag_keep) atomicAdd(&a.diag[2], 2.0f * 1.4189385332046727f + ls0 + ls1);
    else a.diag[2] = 2.0f * 1.4189385332046727f + ls0 + ls1;
    atomicAdd(&G[o.ls], -a.ent_coef); atomicAdd(&G[o.ls + 1], -a.ent_coef);
  }
}


// =================================================================================================================
// Tensor-core variant: the five GEMMs of a tile (layer 1, layer 2, dH1, dW2, dW1) run on mma.sync m16n8k8 TF32 with fp32
// accumulation; operands are rounded to TF32 once, when they are written to shared memory.  Tiles of 32 samples; row strides are
// padded (84 / 132 / 136 / 72 floats) so that fragment loads are (nearly) bank-conflict free.  The weight-gradient accumulators
// are mma C fragments that live in registers across all tiles of the CTA.  Heads, loss derivatives and the bias / head
// gradients reuse the CUDA-core code paths.
// =================================================================================================================
constexpr int TT = 32;                 // samples per tile
constexpr int XS = 84, AS = 132, W1S = 136, W2S = 72;
// Gradient kernel: X, H1, H2 and dH are read both as row-major A operands (bank = g ld + t) and, with the sample index as the
// reduction dimension of the weight-gradient GEMMs, as transposed-A / B operands (bank = t ld + g).  No padding serves both, an
// XOR swizzle does: column c of row r lives at c ^ swz(r) with row strides that are multiples of 32 floats.
constexpr int XSW = 96, ASW = 128;
__device__ __forceinline__ int swz(int r) { return ((r & 3) << 3) | (r & 4); }
constexpr int T_W1T = 0;                          // [KP][W1S]
constexpr int T_W2T = T_W1T + KP * W1S;           // [2][64][W2S]   [net][k][n]
constexpr int T_W2 = T_W2T + 2 * H * W2S;         // [2][64][W2S]   [net][n][k]
constexpr int T_W3 = T_W2 + 2 * H * W2S;          // [3][64]
constexpr int T_B1 = T_W3 + 3 * H;
constexpr int T_B2 = T_B1 + 128;
constexpr int T_B3 = T_B2 + 128;
constexpr int T_LS = T_B3 + 4;
constexpr int T_X = T_LS + 4;                     // [2][TT][XS]  double buffered: the next tile is fetched with cp.async while this one computes
constexpr int T_H1 = T_X + 2 * TT * XSW;          // [TT][AS] (rollout kernel) / [TT][ASW] swizzled (gradient kernel)
constexpr int T_H2 = T_H1 + TT * AS;
constexpr int T_DH = T_H2 + TT * AS;
constexpr int T_DO = T_DH + TT * AS;              // [TT][4]
constexpr int T_W3B = T_DO + TT * 4;              // [128][8]  heads as one mma B operand: col 0/1 = Wa rows (k < 64), col 2 = Wv (k >= 64)
constexpr int T_OUT = T_W3B + 128 * 8;            // [TT][4]   head outputs (mean0, mean1, value) before the bias
constexpr int T_SC = T_OUT + TT * 4;              // [2][TT][8] per-sample scalars (act0, act1, old_logp, adv, ret), double buffered
constexpr int T_IDX = T_SC + 2 * TT * 8;         // [2][TT] int64 row numbers of a tile (two floats each), double buffered
constexpr int T_TOTAL = T_IDX + 2 * TT * 2;
static_assert(T_IDX % 2 == 0, "row numbers are 8-byte aligned");

__device__ __forceinline__ float tf32r(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}
// MUFU.TANH: one instruction, relative error about 2^-11 -- the precision the TF32 operands of the next GEMM keep anyway (the
// branchy libm tanhf was a quarter of the instructions of the tensor-core kernels: both of its paths run in every warp)
__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
// A fragment (16 x 8, row major) of a matrix stored [row][col] with row stride `ld`: rows row0 .. row0+15, cols col0 .. col0+7
__device__ __forceinline__ void lda_rowmajor(uint32_t (&a)[4], const float* m, int ld, int row0, int col0, int g, int t) {
  a[0] = __float_as_uint(m[(row0 + g) * ld + col0 + t]);
  a[1] = __float_as_uint(m[(row0 + g + 8) * ld + col0 + t]);
  a[2] = __float_as_uint(m[(row0 + g) * ld + col0 + t + 4]);
  a[3] = __float_as_uint(m[(row0 + g + 8) * ld + col0 + t + 4]);
}
// A fragment of the TRANSPOSE of a matrix stored [k][row] (row stride ld): element (row, k) = m[k * ld + row]
__device__ __forceinline__ void lda_transposed(uint32_t (&a)[4], const float* m, int ld, int row0, int k0, int g, int t) {
  a[0] = __float_as_uint(m[(k0 + t) * ld + row0 + g]);
  a[1] = __float_as_uint(m[(k0 + t) * ld + row0 + g + 8]);
  a[2] = __float_as_uint(m[(k0 + t + 4) * ld + row0 + g]);
  a[3] = __float_as_uint(m[(k0 + t + 4) * ld + row0 + g + 8]);
}
// B fragment (8 x 8, k x n) of a matrix stored [k][n] with row stride ld
__device__ __forceinline__ void ldb(uint32_t (&b)[2], const float* m, int ld, int k0, int n0, int g, int t) {
  b[0] = __float_as_uint(m[(k0 + t) * ld + n0 + g]);
  b[1] = __float_as_uint(m[(k0 + t + 4) * ld + n0 + g]);
}

// the same three fragment loaders for swizzled arrays (row0, col0, k0, n0 multiples of 8)
__device__ __forceinline__ void lda_rowmajor_sw(uint32_t (&a)[4], const float* m, int ld, int row0, int col0, int g, int t) {
  const int f = swz(g);
  a[0] = __float_as_uint(m[(row0 + g) * ld + ((col0 + t) ^ f)]);
  a[1] = __float_as_uint(m[(row0 + g + 8) * ld + ((col0 + t) ^ f)]);
  a[2] = __float_as_uint(m[(row0 + g) * ld + ((col0 + t + 4) ^ f)]);
  a[3] = __float_as_uint(m[(row0 + g + 8) * ld + ((col0 + t + 4) ^ f)]);
}
__device__ __forceinline__ void lda_transposed_sw(uint32_t (&a)[4], const float* m, int ld, int row0, int k0, int g, int t) {
  const int f0 = swz(t), f1 = swz(t + 4);
  a[0] = __float_as_uint(m[(k0 + t) * ld + ((row0 + g) ^ f0)]);
  a[1] = __float_as_uint(m[(k0 + t) * ld + ((row0 + g + 8) ^ f0)]);
  a[2] = __float_as_uint(m[(k0 + t + 4) * ld + ((row0 + g) ^ f1)]);
  a[3] = __float_as_uint(m[(k0 + t + 4) * ld + ((row0 + g + 8) ^ f1)]);
}
__device__ __forceinline__ void ldb_sw(uint32_t (&b)[2], const float* m, int ld, int k0, int n0, int g, int t) {
  b[0] = __float_as_uint(m[(k0 + t) * ld + ((n0 + g) ^ swz(t))]);
  b[1] = __float_as_uint(m[(k0 + t + 4) * ld + ((n0 + g) ^ swz(t + 4))]);
}

__device__ __forceinline__ void cp_async4(float* smem_dst, const float* gsrc) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sa), "l"(gsrc));
}
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(gsrc));
}
// row numbers of a tile -> shared memory, asynchronously (joins the cp.async group that is committed next): the gather of a
// tile then takes its rows from shared memory instead of stalling every thread on a dependent global load
__device__ __forceinline__ void stage_rows_async(const PpoArgs& a, int tile, long long* rows, int t) {
  if (t < TT && a.idx && tile * TT + t < a.mb) cp_async8(&rows[t], &a.idx[tile * TT + t]);
}
// asynchronous gather of the observation rows of one tile into xbuf (padding written directly); rows = the tile's row numbers
// in shared memory (index mode) or null (contiguous rows)
__device__ __forceinline__ void gather_tile_async(const PpoArgs& a, int tile, float* xbuf, float* scbuf, const long long* rows, int t) {
  const int sbase = tile * TT, ns = min(TT, a.mb - sbase), D = a.D;
  if (t < TT * 5) {   // per-sample scalars ride along with the observation rows
    const int s = t / 5, j = t - s * 5;
    if (s < ns) {
      const int64_t row = rows ? (int64_t)rows[s] : (int64_t)(sbase + s);
      const float* src = j < 2 ? &a.act[row * 2 + j] : (j == 2 ? &a.old_logp[row] : (j == 3 ? &a.adv[row] : &a.ret[row]));
      cp_async4(&scbuf[s * 8 + j], src);
    }
  }
  {   // 8 threads per row, thread (s, kk) owns columns kk, kk + 8, ... of row s: one row pointer per thread
    static_assert(NT == TT * 8 && KP % 8 == 0, "gather mapping");
    const int s = t >> 3, kk = t & 7, f = swz(s);
    float* const dst = xbuf + s * XSW;
    const int64_t row = s < ns ? (rows ? (int64_t)rows[s] : (int64_t)(sbase + s)) : 0;
    const float* const src = a.obs + row * a.pitch;
#pragma unroll
    for (int j = 0; j < KP / 8; ++j) {
      const int k = kk + 8 * j;
      if (s < ns && k < D) cp_async4(&dst[k ^ f], &src[k]);
      // padding; with D < KP the last column is 1 for real samples, so that dW1 = dH1^T X also yields db1 (column KP - 1)
      else dst[k ^ f] = (k == KP - 1 && D < KP && s < ns) ? 1.0f : 0.0f;
    }
  }
  asm volatile("cp.async.commit_group;");
}

__global__ void __launch_bounds__(NT, 1) ppo_grad_kernel_tc(PpoArgs a) {
  extern __shared__ __align__(16) float sm[];
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int g = lane >> 2, q = lane & 3;      // mma fragment coordinates
  const int D = a.D;
  const Offsets o = offsets(D);
  const float* P = a.params;

  for (int i = t; i < KP * 128; i += NT) {
    const int k = i >> 7, n = i & 127, net = n >> 6, r = n & 63;
    sm[T_W1T + k * W1S + n] = (k < D) ? tf32r(P[(net ? o.W1v : o.W1p) + r * D + k]) : 0.0f;
  }
  for (int i = t; i < 2 * H * H; i += NT) {
    const int net = i >> 12, n = (i >> 6) & 63, k = i & 63;
    const float w = tf32r(P[(net ? o.W2v : o.W2p) + n * H + k]);
    sm[T_W2 + net * H * W2S + n * W2S + k] = w;
    sm[T_W2T + net * H * W2S + k * W2S + n] = w;
  }
  for (int i = t; i < 3 * H; i += NT) sm[T_W3 + i] = (i < 2 * H) ? P[o.Wa + i] : P[o.Wv + (i - 2 * H)];
  for (int i = t; i < 128 * 8; i += NT) {
    const int k = i >> 3, n = i & 7;
    float w = 0.0f;
    if (n < 2 && k < H) w = P[o.Wa + n * H + k];
    if (n == 2 && k >= H) w = P[o.Wv + (k - H)];
    sm[T_W3B + i] = tf32r(w);
  }
  if (t < 128) {
    const int net = t >> 6, r = t & 63;
    sm[T_B1 + t] = P[(net ? o.b1v : o.b1p) + r];
    sm[T_B2 + t] = P[(net ? o.b2v : o.b2p) + r];
  }
  if (t < 2) { sm[T_B3 + t] = P[o.ba + t]; sm[T_LS + t] = P[o.ls + t]; }
  if (t == 2) sm[T_B3 + 2] = P[o.bv];
  __syncthreads();

  const float adv_mean = a.adv_stats[0], adv_istd = 1.0f / (a.adv_stats[1] + 1e-8f);
  const float inv_mb = 1.0f / (float)a.mb;
  const float ls0 = sm[T_LS], ls1 = sm[T_LS + 1];
  const float iv0 = expf(-2.0f * ls0), iv1 = expf(-2.0f * ls1);

  // weight-gradient accumulators (mma C fragments):
  //   dW1[n][k]: warp w owns rows n = 16 w .. 16 w + 15 (both nets, 0..127), 10 column tiles of 8 over k = 0..79
  //   dW2[net][n][k]: warp w owns net = w >> 2, rows n = 16 (w & 3) .. + 15, 8 column tiles over k = 0..63
  float acc1[10][4], acc2[8][4];
#pragma unroll
  for (int i = 0; i < 10; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc1[i][j] = 0.0f;
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc2[i][j] = 0.0f;
  //   db2: one more column tile of the dW2 product, against a ones vector (column 0 of accb2)
  //   db1: column KP - 1 of dW1 (the gather writes a ones column into X when D < KP)
  //   head weights dW3[r][n] = sum_s dOut[s][r] H2[s][n]: rows r < 3 of a 16-row tile, warp w owns column tiles 2 w, 2 w + 1
  float accb2[4] = {0.0f, 0.0f, 0.0f, 0.0f}, acc3[2][4];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc3[i][j] = 0.0f;
  float gb1 = 0.0f, gb3 = 0.0f, gls0 = 0.0f, gls1 = 0.0f;
  float d_pg = 0.0f, d_vl = 0.0f, d_kl = 0.0f, d_cf = 0.0f;
  const bool b1_by_mma = D < KP;               // the ones column of X exists
  uint32_t ones_b[2];                          // B fragment of an [8 x 8] matrix whose column 0 is all ones
  ones_b[0] = ones_b[1] = (g == 0) ? __float_as_uint(1.0f) : 0u;

  const int nb0 = 16 * warp;                 // output columns of this warp in the activation GEMMs
  const int gnet = warp >> 2;                // net of those columns
  const int w2net = warp >> 2, w2row0 = 16 * (warp & 3);

  const int ntiles = (a.mb + TT - 1) / TT;
  int buf = 0;
  long long* const rows_sm = a.idx ? reinterpret_cast<long long*>(&sm[T_IDX]) : nullptr;   // [2][TT]
  if (rows_sm) {   // row numbers of this CTA's first two tiles (later ones are staged one tile ahead of their gather)
    if (t < TT && blockIdx.x * TT + t < a.mb) rows_sm[t] = a.idx[blockIdx.x * TT + t];
    if (t < TT && (blockIdx.x + gridDim.x) * TT + t < a.mb) rows_sm[TT + t] = a.idx[(blockIdx.x + gridDim.x) * TT + t];
    __syncthreads();
  }
  if ((int)blockIdx.x < ntiles) gather_tile_async(a, blockIdx.x, &sm[T_X], &sm[T_SC], rows_sm, t);
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, buf ^= 1) {
    const int sbase = tile * TT;
    const int ns = min(TT, a.mb - sbase);
    float* const X = &sm[T_X + buf * TT * XSW];
    // ---- 1. this tile's observation rows have been fetched asynchronously: wait, round to TF32 (each thread its own elements),
    // then start fetching the next tile into the other buffer
    asm volatile("cp.async.wait_all;");
    {   // same (row, column) ownership as the gather
      const int s = t >> 3, kk = t & 7, f = swz(s);
#pragma unroll
      for (int j = 0; j < KP / 8; ++j) {
        const int xi = s * XSW + ((kk + 8 * j) ^ f);
        X[xi] = tf32r(X[xi]);
      }
    }
    __syncthreads();
    if (tile + (int)gridDim.x < ntiles) {
      // this tile's row numbers are no longer needed (its gather has completed): their slot takes those of the tile after next
      if (rows_sm && tile + 2 * (int)gridDim.x < ntiles) stage_rows_async(a, tile + 2 * gridDim.x, rows_sm + buf * TT, t);
      gather_tile_async(a, tile + gridDim.x, &sm[T_X + (buf ^ 1) * TT * XSW], &sm[T_SC + (buf ^ 1) * TT * 8],
                        rows_sm ? rows_sm + (buf ^ 1) * TT : nullptr, t);
    }
    const float* const SC = &sm[T_SC + buf * TT * 8];
    // ---- 2. layer 1: H1[s][n] = tanh(sum_k X[s][k] W1t[k][n] + b1[n])
    {
      float c[2][2][4];
#pragma unroll
      for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
          for (int e = 0; e < 4; ++e) c[m][j][e] = 0.0f;
#pragma unroll 2
      for (int k0 = 0; k0 < KP; k0 += 8) {
        uint32_t af[2][4], bf[2][2];
        lda_rowmajor_sw(af[0], X, XSW, 0, k0, g, q);
        lda_rowmajor_sw(af[1], X, XSW, 16, k0, g, q);
        ldb(bf[0], &sm[T_W1T], W1S, k0, nb0, g, q);
        ldb(bf[1], &sm[T_W1T], W1S, k0, nb0 + 8, g, q);
#pragma unroll
        for (int m = 0; m < 2; ++m)
#pragma unroll
          for (int j = 0; j < 2; ++j) mma_tf32(c[m][j], af[m], bf[j]);
      }
#pragma unroll
      for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int col = nb0 + 8 * j + 2 * q, row = 16 * m + g, pc = col ^ swz(g);
          sm[T_H1 + row * ASW + pc] = tf32r(tanh_fast(c[m][j][0] + sm[T_B1 + col]));
          sm[T_H1 + row * ASW + pc + 1] = tf32r(tanh_fast(c[m][j][1] + sm[T_B1 + col + 1]));
          sm[T_H1 + (row + 8) * ASW + pc] = tf32r(tanh_fast(c[m][j][2] + sm[T_B1 + col]));
          sm[T_H1 + (row + 8) * ASW + pc + 1] = tf32r(tanh_fast(c[m][j][3] + sm[T_B1 + col + 1]));
        }
    }
    __syncthreads();
    // ---- 3. layer 2 (block diagonal)
    {
      float c[2][2][4];
#pragma unroll
      for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
          for (int e = 0; e < 4; ++e) c[m][j][e] = 0.0f;
      const float* A = &sm[T_H1 + gnet * H];
      const float* B = &sm[T_W2T + gnet * H * W2S];
      const int nn0 = nb0 - gnet * H;
#pragma unroll 2
      for (int k0 = 0; k0 < H; k0 += 8) {
        uint32_t af[2][4], bf[2][2];
        lda_rowmajor_sw(af[0], A, ASW, 0, k0, g, q);
        lda_rowmajor_sw(af[1], A, ASW, 16, k0, g, q);
        ldb(bf[0], B, W2S, k0, nn0, g, q);
        ldb(bf[1], B, W2S, k0, nn0 + 8, g, q);
#pragma unroll
        for (int m = 0; m < 2; ++m)
#pragma unroll
          for (int j = 0; j < 2; ++j) mma_tf32(c[m][j], af[m], bf[j]);
      }
#pragma unroll
      for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int col = nb0 + 8 * j + 2 * q, row = 16 * m + g, pc = col ^ swz(g);
          sm[T_H2 + row * ASW + pc] = tf32r(tanh_fast(c[m][j][0] + sm[T_B2 + col]));
          sm[T_H2 + row * ASW + pc + 1] = tf32r(tanh_fast(c[m][j][1] + sm[T_B2 + col + 1]));
          sm[T_H2 + (row + 8) * ASW + pc] = tf32r(tanh_fast(c[m][j][2] + sm[T_B2 + col]));
          sm[T_H2 + (row + 8) * ASW + pc + 1] = tf32r(tanh_fast(c[m][j][3] + sm[T_B2 + col + 1]));
        }
    }
    __syncthreads();
    // ---- 4. heads as one small GEMM [32 x 128] x [128 x 8] (warps 0 and 1, one 16-row tile each), then the PPO loss
    // derivatives with one thread per sample
    if (warp < 2) {
      float c[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll 4
      for (int k0 = 0; k0 < 128; k0 += 8) {
        uint32_t af[4], bf[2];
        lda_rowmajor_sw(af, &sm[T_H2], ASW, 16 * warp, k0, g, q);
        ldb(bf, &sm[T_W3B], 8, k0, 0, g, q);
        mma_tf32(c, af, bf);
      }
      if (q < 2) {   // columns 2q, 2q+1 of rows g and g+8
        const int row = 16 * warp + g;
        sm[T_OUT + row * 4 + 2 * q] = c[0]; sm[T_OUT + row * 4 + 2 * q + 1] = c[1];
        sm[T_OUT + (row + 8) * 4 + 2 * q] = c[2]; sm[T_OUT + (row + 8) * 4 + 2 * q + 1] = c[3];
      }
    }
    __syncthreads();
    if (t < TT) {
      const int s = t;
      float dm0 = 0.0f, dm1 = 0.0f, dv = 0.0f;
      if (s < ns) {
        const float m0 = sm[T_OUT + s * 4] + sm[T_B3], m1 = sm[T_OUT + s * 4 + 1] + sm[T_B3 + 1], v = sm[T_OUT + s * 4 + 2] + sm[T_B3 + 2];
        const float e0 = SC[s * 8] - m0, e1 = SC[s * 8 + 1] - m1;
        const float q0 = e0 * e0 * iv0, q1 = e1 * e1 * iv1;
        const float logp = -0.5f * q0 - ls0 - 0.9189385332046727f - 0.5f * q1 - ls1 - 0.9189385332046727f;
        const float A_ = (SC[s * 8 + 3] - adv_mean) * adv_istd;
        const float lr = logp - SC[s * 8 + 2];
        const float r = expf(lr);
        const float rc = fminf(fmaxf(r, 1.0f - a.clip), 1.0f + a.clip);
        const float s1 = A_ * r, s2 = A_ * rc;
        const float dlogp = (s1 <= s2) ? -A_ * r : 0.0f;      // d(-min(s1, s2)) / d logp (clipped branch has zero slope)
        dm0 = dlogp * e0 * iv0 * inv_mb; dm1 = dlogp * e1 * iv1 * inv_mb;
        gls0 += dlogp * (q0 - 1.0f) * inv_mb; gls1 += dlogp * (q1 - 1.0f) * inv_mb;
        const float R = SC[s * 8 + 4];
        dv = a.vf_coef * 2.0f * (v - R) * inv_mb;
        d_pg += -fminf(s1, s2); d_vl += (v - R) * (v - R); d_kl += (r - 1.0f) - lr; d_cf += (fabsf(r - 1.0f) > a.clip) ? 1.0f : 0.0f;
      }
      sm[T_DO + s * 4] = dm0; sm[T_DO + s * 4 + 1] = dm1; sm[T_DO + s * 4 + 2] = dv; sm[T_DO + s * 4 + 3] = 0.0f;
    }
    __syncthreads();
    // ---- 5. dH2 = (dOut W3) * (1 - H2^2)   (TF32-rounded: it feeds two GEMMs)
    {   // thread = one hidden column n, every second sample: the head weights of the column stay in registers
      static_assert(NT == 256 && T_DO % 4 == 0, "dH2 mapping");
      const int n = t & 127, s_off = t >> 7;
      const bool pol = n < H;                       // warp-uniform
      const float w_a = pol ? sm[T_W3 + n] : sm[T_W3 + 2 * H + (n - H)];
      const float w_b = pol ? sm[T_W3 + H + n] : 0.0f;
#pragma unroll 4
      for (int j = 0; j < TT / 2; ++j) {
        const int s = s_off + 2 * j;
        const float4 d = *reinterpret_cast<const float4*>(&sm[T_DO + s * 4]);
        const int pn = s * ASW + (n ^ swz(s));
        const float h = sm[T_H2 + pn];
        const float gg = pol ? d.x * w_a + d.y * w_b : d.z * w_a;
        sm[T_DH + pn] = tf32r(gg * (1.0f - h * h));
      }
    }
    __syncthreads();
    // ---- 6. head weights dW3 += dOut^T H2, dW2 += dH2^T H1 and db2 += dH2^T 1 (tensor cores)
    if (t < 3)
      for (int s = 0; s < TT; ++s) gb3 += sm[T_DO + s * 4 + t];
    {
      const float* B3 = &sm[T_H2];                  // H2[s][n] -> B[k = s][n]
#pragma unroll
      for (int s0 = 0; s0 < TT; s0 += 8) {
        uint32_t af[4];                             // A = dOut^T: element (r, s) = dOut[s][r]; rows r >= 4 are zero
        af[0] = (g < 4) ? __float_as_uint(tf32r(sm[T_DO + (s0 + q) * 4 + g])) : 0u;
        af[2] = (g < 4) ? __float_as_uint(tf32r(sm[T_DO + (s0 + q + 4) * 4 + g])) : 0u;
        af[1] = af[3] = 0u;
#pragma unroll
        for (int jj = 0; jj < 2; ++jj) {
          uint32_t bf[2];
          ldb_sw(bf, B3, ASW, s0, 8 * (2 * warp + jj), g, q);
          mma_tf32(acc3[jj], af, bf);
        }
      }
      const float* At = &sm[T_DH + w2net * H];      // dH2[s][net*64 + n]  -> A = transpose, element (n, s)
      const float* B = &sm[T_H1 + w2net * H];       // H1[s][net*64 + k]   -> B[k = s][n = k]
#pragma unroll
      for (int s0 = 0; s0 < TT; s0 += 8) {
        uint32_t af[4];
        lda_transposed_sw(af, At, ASW, w2row0, s0, g, q);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          uint32_t bf[2];
          ldb_sw(bf, B, ASW, s0, 8 * j, g, q);
          mma_tf32(acc2[j], af, bf);
        }
        mma_tf32(accb2, af, ones_b);
      }
    }
    // ---- 7. dH1 = (dH2 W2) * (1 - H1^2)
    {
      float c[2][2][4];
#pragma unroll
      for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
          for (int e = 0; e < 4; ++e) c[m][j][e] = 0.0f;
      const float* A = &sm[T_DH + gnet * H];
      const float* B = &sm[T_W2 + gnet * H * W2S];
      const int nn0 = nb0 - gnet * H;
#pragma unroll 2
      for (int k0 = 0; k0 < H; k0 += 8) {
        uint32_t af[2][4], bf[2][2];
        lda_rowmajor_sw(af[0], A, ASW, 0, k0, g, q);
        lda_rowmajor_sw(af[1], A, ASW, 16, k0, g, q);
        ldb(bf[0], B, W2S, k0, nn0, g, q);
        ldb(bf[1], B, W2S, k0, nn0 + 8, g, q);
#pragma unroll
        for (int m = 0; m < 2; ++m)
#pragma unroll
          for (int j = 0; j < 2; ++j) mma_tf32(c[m][j], af[m], bf[j]);
      }
      __syncthreads();   // all reads of dH2 (steps 6 and 7) are done: the buffer now takes dH1
#pragma unroll
      for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int col = nb0 + 8 * j + 2 * q, row = 16 * m + g, pc = col ^ swz(g);
          const float h00 = sm[T_H1 + row * ASW + pc], h01 = sm[T_H1 + row * ASW + pc + 1];
          const float h10 = sm[T_H1 + (row + 8) * ASW + pc], h11 = sm[T_H1 + (row + 8) * ASW + pc + 1];
          sm[T_DH + row * ASW + pc] = tf32r(c[m][j][0] * (1.0f - h00 * h00));
          sm[T_DH + row * ASW + pc + 1] = tf32r(c[m][j][1] * (1.0f - h01 * h01));
          sm[T_DH + (row + 8) * ASW + pc] = tf32r(c[m][j][2] * (1.0f - h10 * h10));
          sm[T_DH + (row + 8) * ASW + pc + 1] = tf32r(c[m][j][3] * (1.0f - h11 * h11));
        }
    }
    __syncthreads();
    // ---- 8. db1, dW1 += dH1^T X
    if (!b1_by_mma && t < 128)
      for (int s = 0; s < TT; ++s) gb1 += sm[T_DH + s * ASW + (t ^ swz(s))];
    {
      const float* At = &sm[T_DH];                  // dH1[s][n] -> element (n, s)
      const float* B = X;                           // X[s][k]   -> B[k = s][n = k]
#pragma unroll
      for (int s0 = 0; s0 < TT; s0 += 8) {
        uint32_t af[4];
        lda_transposed_sw(af, At, ASW, 16 * warp, s0, g, q);
#pragma unroll
        for (int j = 0; j < 10; ++j) {
          uint32_t bf[2];
          ldb_sw(bf, B, XSW, s0, 8 * j, g, q);
          mma_tf32(acc1[j], af, bf);
        }
      }
    }
    __syncthreads();
  }

  // ---- flush: C fragment element e of tile j: row = r0 + g (+8 for e >= 2), col = 8 j + 2 q + (e & 1)
  float* G = a.grads;
#pragma unroll
  for (int j = 0; j < 10; ++j)
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int n = 16 * warp + g + ((e & 2) ? 8 : 0), k = 8 * j + 2 * q + (e & 1);
      const int net = n >> 6, r = n & 63;
      if (k < D) atomicAdd(&G[(net ? o.W1v : o.W1p) + r * D + k], acc1[j][e]);
      else if (b1_by_mma && k == KP - 1) atomicAdd(&G[(net ? o.b1v : o.b1p) + r], acc1[j][e]);
    }
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int n = w2row0 + g + ((e & 2) ? 8 : 0), k = 8 * j + 2 * q + (e & 1);
      atomicAdd(&G[(w2net ? o.W2v : o.W2p) + n * H + k], acc2[j][e]);
    }
  if (q == 0) {   // db2: column 0 of the ones-vector tile, rows g and g + 8
    atomicAdd(&G[(w2net ? o.b2v : o.b2p) + w2row0 + g], accb2[0]);
    atomicAdd(&G[(w2net ? o.b2v : o.b2p) + w2row0 + g + 8], accb2[2]);
  }
  if (g < 3) {    // head weights: row r = g of the two column tiles of this warp (policy heads read H2[:, :64], the value head H2[:, 64:])
#pragma unroll
    for (int jj = 0; jj < 2; ++jj)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int n = 8 * (2 * warp + jj) + 2 * q + e;
        if (g < 2 && n < H) atomicAdd(&G[o.Wa + g * H + n], acc3[jj][e]);
        if (g == 2 && n >= H) atomicAdd(&G[o.Wv + (n - H)], acc3[jj][e]);
      }
  }
  if (!b1_by_mma && t < 128) {
    const int net = t >> 6, r = t & 63;
    atomicAdd(&G[(net ? o.b1v : o.b1p) + r], gb1);
  }
  if (t < 2) atomicAdd(&G[o.ba + t], gb3);
  if (t == 2) atomicAdd(&G[o.bv], gb3);
  if (warp == 0) {   // the per-sample sums live in the 32 lanes of warp 0
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      gls0 += __shfl_xor_sync(0xffffffffu, gls0, off); gls1 += __shfl_xor_sync(0xffffffffu, gls1, off);
      d_pg += __shfl_xor_sync(0xffffffffu, d_pg, off); d_vl += __shfl_xor_sync(0xffffffffu, d_vl, off);
      d_kl += __shfl_xor_sync(0xffffffffu, d_kl, off); d_cf += __shfl_xor_sync(0xffffffffu, d_cf, off);
    }
    if (lane == 0) {
      atomicAdd(&G[o.ls], gls0); atomicAdd(&G[o.ls + 1], gls1);
      atomicAdd(&a.diag[0], d_pg * inv_mb); atomicAdd(&a.diag[1], d_vl * inv_mb);
      atomicAdd(&a.diag[3], d_kl * inv_mb); atomicAdd(&a.diag[4], d_cf * inv_mb);
    }
  }
  if (blockIdx.x == 0 && t == 0) {
    if (a.diag_keep) atomicAdd(&a.diag[2], 2.0f * 1.4189385332046727f + ls0 + ls1);
    else a.diag[2] = 2.0f * 1.4189385332046727f + ls0 + ls1;
    atomicAdd(&G[o.ls], -a.ent_coef); atomicAdd(&G[o.ls + 1], -a.ent_coef);
  }
}


// =================================================================================================================
// Rollout side: policy / value forward for a whole batch of observations in one launch (what SB3's policy.forward() does per
// step inside collect_rollouts): mean = pi(obs), value = V(obs), and, if requested, a sampled action a = mean + exp(log_std) * eps
// with its log-probability.  eps ~ N(0, 1) from Philox4x32-10 keyed by (seed, step, row) through Box-Muller.  Same TF32 tiles as
// the gradient kernel (forward half), one CTA per SM walking 32-row tiles.
// =================================================================================================================
struct ActArgs {
  const float* obs;
  int n, D;
  const float* params;
  float *mean, *value, *action, *logp;     // mean [n][2] / action [n][2] / logp [n] may be null
  unsigned long long seed;
  unsigned step;
  int value_only;                           // skip the policy net (bootstrap values of terminal observations)
  // time-limit bootstrap (value_only): reward_out = reward + gamma V(obs) where the episode was truncated but not terminated,
  // reward elsewhere; done_out = terminated | truncated as float.  Tiles without such a row skip the network.
  const uint8_t *term, *trunc;
  const float* reward;
  float gamma;
  float *reward_out, *done_out;
  int pitch;                                // floats between consecutive observation rows (>= D)
};

__device__ __forceinline__ void philox_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t* out) {
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// asynchronous fetch of the observation rows of one tile (contiguous rows, stride XS, zero padding written directly)
__device__ __forceinline__ void fetch_obs_tile_async(const ActArgs& a, int tile, float* xbuf, int t) {
  const int sbase = tile * TT, ns = min(TT, a.n - sbase), D = a.D;
  for (int i = t; i < TT * KP; i += NT) {
    const int s = i / KP, k = i - s * KP;
    if (s < ns && k < D) cp_async4(&xbuf[s * XS + k], &a.obs[(size_t)(sbase + s) * a.pitch + k]);
    else xbuf[s * XS + k] = 0.0f;
  }
}
// does this tile need the network?  (bootstrap mode: only tiles with a truncated-but-not-terminated row)
__device__ __forceinline__ bool tile_needs_value(const ActArgs& a, int tile, int t) {
  if (!a.reward_out) return true;
  const int row = tile * TT + t;
  const bool need = t < TT && row < a.n && a.trunc[row] != 0 && a.term[row] == 0;
  return __syncthreads_or(need) != 0;
}

__global__ void __launch_bounds__(NT, 1) ppo_act_kernel(ActArgs a) {
  extern __shared__ __align__(16) float sm[];
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int g = lane >> 2, q = lane & 3;
  const int D = a.D;
  const Offsets o = offsets(D);
  const float* P = a.params;
  if (a.reward_out) {
    // bootstrap mode: most calls have no truncated episode at all -- look at this CTA's rows first and leave before the weights are
    // staged (the staging is 30 us of a call that otherwise only copies rewards and flags)
    // (all 256 threads scan: NT / TT tiles per pass, independent loads -- a per-tile loop of dependent byte loads costs 20 us)
    const int nt = (a.n + TT - 1) / TT;
    bool any = false;
    for (int k = t / TT; blockIdx.x + k * gridDim.x < nt; k += NT / TT) {
      const int row = (blockIdx.x + k * gridDim.x) * TT + (t % TT);
      if (row < a.n) any = any | ((a.trunc[row] != 0) & (a.term[row] == 0));
    }
    if (!__syncthreads_or(any)) {
      for (int k = t / TT; blockIdx.x + k * gridDim.x < nt; k += NT / TT) {
        const int row = (blockIdx.x + k * gridDim.x) * TT + (t % TT);
        if (row < a.n) {
          a.reward_out[row] = a.reward[row];
          a.done_out[row] = (a.term[row] != 0 || a.trunc[row] != 0) ? 1.0f : 0.0f;
        }
      }
      return;
    }
  }
  for (int i = t; i < KP * 128; i += NT) {
    const int k = i >> 7, n = i & 127, net = n >> 6, r = n & 63;
    sm[T_W1T + k * W1S + n] = (k < D) ? tf32r(P[(net ? o.W1v : o.W1p) + r * D + k]) : 0.0f;
  }
  for (int i = t; i < 2 * H * H; i += NT) {
    const int net = i >> 12, n = (i >> 6) & 63, k = i & 63;
    sm[T_W2T + net * H * W2S + k * W2S + n] = tf32r(P[(net ? o.W2v : o.W2p) + n * H + k]);
  }
  for (int i = t; i < 3 * H; i += NT) sm[T_W3 + i] = (i < 2 * H) ? P[o.Wa + i] : P[o.Wv + (i - 2 * H)];
  if (t < 128) {
    const int net = t >> 6, r = t & 63;
    sm[T_B1 + t] = P[(net ? o.b1v : o.b1p) + r];
    sm[T_B2 + t] = P[(net ? o.b2v : o.b2p) + r];
  }
  if (t < 2) { sm[T_B3 + t] = P[o.ba + t]; sm[T_LS + t] = P[o.ls + t]; }
  if (t == 2) sm[T_B3 + 2] = P[o.bv];
  __syncthreads();
  const float ls0 = sm[T_LS], ls1 = sm[T_LS + 1];
  const float sd0 = expf(ls0), sd1 = expf(ls1);
  const int nb0 = 16 * warp, gnet = warp >> 2;
  const bool skip = a.value_only && gnet == 0;     // warps 0..3 own the policy net's columns
  const int ntiles = (a.n + TT - 1) / TT;
  // the observation rows of the next tile are fetched with cp.async into the other buffer while this tile computes
  int buf = 0;
  bool need = (int)blockIdx.x < ntiles && tile_needs_value(a, blockIdx.x, t);
  if (need) fetch_obs_tile_async(a, blockIdx.x, &sm[T_X], t);
  asm volatile("cp.async.commit_group;");
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, buf ^= 1) {
    const int sbase = tile * TT, ns = min(TT, a.n - sbase);
    float* const X = &sm[T_X + buf * TT * XSW];
    const bool need_this = need;
    asm volatile("cp.async.wait_all;");
    if (need_this)
      for (int i = t; i < TT * KP; i += NT) {      // round to TF32 (each thread its own elements)
        const int s = i / KP, k = i - s * KP;
        X[s * XS + k] = tf32r(X[s * XS + k]);
      }
    __syncthreads();
    const int nxt = tile + gridDim.x;
    need = nxt < ntiles && tile_needs_value(a, nxt, t);
    if (need) fetch_obs_tile_async(a, nxt, &sm[T_X + (buf ^ 1) * TT * XSW], t);
    asm volatile("cp.async.commit_group;");
    if (!need_this) {      // nothing to bootstrap in this tile: rewards pass through
      if (t < ns) {
        a.reward_out[sbase + t] = a.reward[sbase + t];
        a.done_out[sbase + t] = (a.term[sbase + t] != 0 || a.trunc[sbase + t] != 0) ? 1.0f : 0.0f;
      }
      continue;
    }
    if (!skip) {
      float c[2][2][4];
#pragma unroll
      for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
          for (int e = 0; e < 4; ++e) c[m][j][e] = 0.0f;
#pragma unroll 2
      for (int k0 = 0; k0 < KP; k0 += 8) {
        uint32_t af[2][4], bf[2][2];
        lda_rowmajor(af[0], X, XS, 0, k0, g, q);
        lda_rowmajor(af[1], X, XS, 16, k0, g, q);
        ldb(bf[0], &sm[T_W1T], W1S, k0, nb0, g, q);
        ldb(bf[1], &sm[T_W1T], W1S, k0, nb0 + 8, g, q);
#pragma unroll
        for (int m = 0; m < 2; ++m)
#pragma unroll
          for (int j = 0; j < 2; ++j) mma_tf32(c[m][j], af[m], bf[j]);
      }
#pragma unroll
      for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int col = nb0 + 8 * j + 2 * q, row = 16 * m + g;
          sm[T_H1 + row * AS + col] = tf32r(tanh_fast(c[m][j][0] + sm[T_B1 + col]));
          sm[T_H1 + row * AS + col + 1] = tf32r(tanh_fast(c[m][j][1] + sm[T_B1 + col + 1]));
          sm[T_H1 + (row + 8) * AS + col] = tf32r(tanh_fast(c[m][j][2] + sm[T_B1 + col]));
          sm[T_H1 + (row + 8) * AS + col + 1] = tf32r(tanh_fast(c[m][j][3] + sm[T_B1 + col + 1]));
        }
    }
    __syncthreads();
    if (!skip) {
      float c[2][2][4];
#pragma unroll
      for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
          for (int e = 0; e < 4; ++e) c[m][j][e] = 0.0f;
      const float* A = &sm[T_H1 + gnet * H];
      const float* B = &sm[T_W2T + gnet * H * W2S];
      const int nn0 = nb0 - gnet * H;
#pragma unroll 2
      for (int k0 = 0; k0 < H; k0 += 8) {
        uint32_t af[2][4], bf[2][2];
        lda_rowmajor(af[0], A, AS, 0, k0, g, q);
        lda_rowmajor(af[1], A, AS, 16, k0, g, q);
        ldb(bf[0], B, W2S, k0, nn0, g, q);
        ldb(bf[1], B, W2S, k0, nn0 + 8, g, q);
#pragma unroll
        for (int m = 0; m < 2; ++m)
#pragma unroll
          for (int j = 0; j < 2; ++j) mma_tf32(c[m][j], af[m], bf[j]);
      }
#pragma unroll
      for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int col = nb0 + 8 * j + 2 * q, row = 16 * m + g;
          sm[T_H2 + row * AS + col] = tanh_fast(c[m][j][0] + sm[T_B2 + col]);
          sm[T_H2 + row * AS + col + 1] = tanh_fast(c[m][j][1] + sm[T_B2 + col + 1]);
          sm[T_H2 + (row + 8) * AS + col] = tanh_fast(c[m][j][2] + sm[T_B2 + col]);
          sm[T_H2 + (row + 8) * AS + col + 1] = tanh_fast(c[m][j][3] + sm[T_B2 + col + 1]);
        }
    }
    __syncthreads();
    // heads: warp w owns rows 4w .. 4w+3 of the tile
#pragma unroll 1
    for (int si = 0; si < TT / 8; ++si) {
      const int s = warp * (TT / 8) + si;
      const float* h2 = &sm[T_H2 + s * AS];
      float p0 = 0.0f, p1 = 0.0f;
      if (!a.value_only) {
        p0 = h2[lane] * sm[T_W3 + lane] + h2[lane + 32] * sm[T_W3 + lane + 32];
        p1 = h2[lane] * sm[T_W3 + H + lane] + h2[lane + 32] * sm[T_W3 + H + lane + 32];
      }
      float pv = h2[H + lane] * sm[T_W3 + 2 * H + lane] + h2[H + lane + 32] * sm[T_W3 + 2 * H + lane + 32];
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        p0 += __shfl_xor_sync(0xffffffffu, p0, off); p1 += __shfl_xor_sync(0xffffffffu, p1, off); pv += __shfl_xor_sync(0xffffffffu, pv, off);
      }
      if (lane == 0 && s < ns) {
        const size_t row = (size_t)(sbase + s);
        const float v = pv + sm[T_B3 + 2];
        if (a.value) a.value[row] = v;
        if (a.reward_out) {
          const bool tm = a.term[row] != 0, tr = a.trunc[row] != 0;
          a.reward_out[row] = a.reward[row] + ((tr && !tm) ? a.gamma * v : 0.0f);
          a.done_out[row] = (tm || tr) ? 1.0f : 0.0f;
        }
        if (!a.value_only) {
          const float m0 = p0 + sm[T_B3], m1 = p1 + sm[T_B3 + 1];
          if (a.mean) { a.mean[row * 2] = m0; a.mean[row * 2 + 1] = m1; }
          if (a.action) {
            uint32_t r[4];
            philox_10(a.step, (uint32_t)row, (uint32_t)(row >> 32), 0x50504F41u, (uint32_t)a.seed, (uint32_t)(a.seed >> 32), r);
            // Box-Muller on two uniforms in (0, 1]
            const float u1 = ((float)(r[0] >> 8) + 1.0f) * (1.0f / 16777216.0f), u2 = (float)(r[1] >> 8) * (1.0f / 16777216.0f);
            const float rad = sqrtf(-2.0f * logf(u1));
            float sn, cs;
            sincosf(6.283185307179586f * u2, &sn, &cs);
            const float e0 = rad * cs, e1 = rad * sn;
            a.action[row * 2] = m0 + sd0 * e0; a.action[row * 2 + 1] = m1 + sd1 * e1;
            if (a.logp) a.logp[row] = -0.5f * e0 * e0 - ls0 - 0.9189385332046727f - 0.5f * e1 * e1 - ls1 - 0.9189385332046727f;
          }
        }
      }
    }
    __syncthreads();
  }
}

// generalised advantage estimate, one thread per environment walking its T rollout steps backwards (arrays are [T][n], so the
// loads of a warp are coalesced): delta_t = r_t + gamma V_{t+1} (1 - done_t) - V_t,  A_t = delta_t + gamma lam (1 - done_t) A_{t+1}
__global__ void gae_kernel(const float* __restrict__ rew, const float* __restrict__ val, const float* __restrict__ done,
                           const float* __restrict__ last_val, int T, int n, float gamma, float lam, float* __restrict__ adv,
                           float* __restrict__ ret) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  float next = last_val[e], last = 0.0f;
  for (int t = T - 1; t >= 0; --t) {
    const size_t i = (size_t)t * n + e;
    const float nt = 1.0f - done[i], v = val[i];
    const float delta = rew[i] + gamma * next * nt - v;
    last = delta + gamma * lam * nt * last;
    adv[i] = last;
    ret[i] = last + v;
    next = v;
  }
}

// Pseudo-random permutation of 0 .. n-1 without a sort: a keyed balanced Feistel network is a bijection of [0, 2^(2h)), and
// cycle walking (re-encrypt until the value is < n) restricts it to [0, n); 2^(2h) < 4 n, so a value walks < 4 times on average.
__device__ __forceinline__ uint32_t feistel_f(uint32_t x, uint32_t key) {
  x = (x ^ key) * 0x9E3779B1u; x ^= x >> 15;
  x *= 0x85EBCA77u; x ^= x >> 13;
  x *= 0xC2B2AE3Du; x ^= x >> 16;
  return x;
}
__global__ void permutation_kernel(int64_t* __restrict__ perm, long long n, int half_bits, uint64_t seed, uint32_t stream_id) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t mask = (1u << half_bits) - 1u;
  uint32_t keys[8];
#pragma unroll
  for (int r = 0; r < 8; ++r) keys[r] = feistel_f((uint32_t)seed + 0x632BE5ABu * (uint32_t)r, (uint32_t)(seed >> 32) ^ (stream_id * 0x9E3779B9u + (uint32_t)r));
  unsigned long long v = (unsigned long long)i;
  do {
    uint32_t L = (uint32_t)(v >> half_bits) & mask, R = (uint32_t)v & mask;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const uint32_t nl = R, nr = L ^ (feistel_f(R, keys[r]) & mask);
      L = nl; R = nr;
    }
    v = ((unsigned long long)L << half_bits) | R;
  } while (v >= (unsigned long long)n);
  perm[i] = (int64_t)v;
}

// block-wide sum of one double per thread (1024 threads); every thread receives the total
__device__ __forceinline__ double block_sum_1024(double v, double* scratch) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  __syncthreads();                     // scratch may still be read from a previous call
  if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
  __syncthreads();
  double tot = 0.0;
#pragma unroll 8
  for (int w = 0; w < 32; ++w) tot += scratch[w];
  return tot;
}

// mean and unbiased standard deviation (torch.Tensor.std) of the advantages of one minibatch: every CTA adds its partial sums
// (double) to two device-scope accumulators, the last CTA to finish writes the result and clears them for the next call
// (calls on one device must not overlap: they share the accumulators)
__device__ double g_adv_ws[3] = {0.0, 0.0, 0.0};   // default workspace of ackb_ppo_adv_stats (one learner per device, one stream)
__global__ void __launch_bounds__(256) adv_stats_kernel(const float* __restrict__ adv, const int64_t* __restrict__ idx, int n,
                                                        float* __restrict__ mean_std, double* __restrict__ ws) {
  adv_stats_block(adv, idx, n, mean_std, ws, blockIdx.x, gridDim.x);      // ackb_ppo_common.cuh
}

// global-norm gradient clipping (torch.nn.utils.clip_grad_norm_) + Adam (torch.optim.Adam, no weight decay / amsgrad) on the flat
// parameter vector; `step` is the optimiser's device-side step counter (float, as torch keeps it when capturable).
// One thread-block cluster of ADAM_CTAS CTAs: every CTA computes the full gradient norm itself (75 KB, L2 resident; four loads in
// flight per thread) and updates its 1/ADAM_CTAS share of the elements; the cluster barrier orders all reads of `step` before its
// update.  (The single-CTA version took 28 us per optimiser step, 18 dependent load round trips twice: 8 % of a PPO iteration.)
constexpr int ADAM_CTAS = 8;
__global__ void __cluster_dims__(ADAM_CTAS, 1, 1) __launch_bounds__(1024, 1)
    clip_adam_kernel(float* __restrict__ p, const float* __restrict__ gr, float* __restrict__ m, float* __restrict__ v, float* __restrict__ step,
                     int n, float max_norm, float lr, float b1, float b2, float eps) {
  __shared__ double scratch[32];
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
  int i = threadIdx.x;
  for (; i + 3 * 1024 < n; i += 4 * 1024) {
    const double g0 = (double)gr[i], g1 = (double)gr[i + 1024], g2 = (double)gr[i + 2048], g3 = (double)gr[i + 3072];
    s0 += g0 * g0; s1 += g1 * g1; s2 += g2 * g2; s3 += g3 * g3;
  }
  for (; i < n; i += 1024) { const double g = (double)gr[i]; s0 += g * g; }
  const float norm = (float)sqrt(block_sum_1024((s0 + s1) + (s2 + s3), scratch));
  const float coef = fminf(max_norm / (norm + 1e-6f), 1.0f);
  const float t = step[0] + 1.0f;
  const float bc1 = 1.0f - powf(b1, t), bc2s = sqrtf(1.0f - powf(b2, t));
  const float step_size = lr / bc1;
  for (int k = blockIdx.x * 1024 + threadIdx.x; k < n; k += ADAM_CTAS * 1024) {
    const float g = gr[k] * coef;
    const float mi = m[k] + (1.0f - b1) * (g - m[k]);          // lerp, as torch does
    const float vi = b2 * v[k] + (1.0f - b2) * g * g;
    m[k] = mi; v[k] = vi;
    p[k] -= step_size * mi / (sqrtf(vi) / bc2s + eps);
  }
  asm volatile("barrier.cluster.arrive.aligned;\n\tbarrier.cluster.wait.aligned;" ::: "memory");     // every CTA has read step[0]
  if (blockIdx.x == 0 && threadIdx.x == 0) step[0] = t;
}

// clip_adam_kernel with the data-parallel gradient all-reduce in front (include/ackb_ppo.h: ackb_ppo_clip_adam_allreduce): flags and
// gradients of the peers are read / written through NVLink peer mappings with system-scope release / acquire.
__device__ __forceinline__ void st_release_sys_u32(uint32_t* p, uint32_t v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t ld_acquire_sys_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float ld_relaxed_sys_f32(const float* p) {
  float v;
  asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
  return v;
}
__global__ void __cluster_dims__(ADAM_CTAS, 1, 1) __launch_bounds__(1024, 1)
    clip_adam_allreduce_kernel(float* __restrict__ p, const uint64_t* __restrict__ peer_g, const uint64_t* __restrict__ peer_f,
                               const int* __restrict__ cur_buf, int buf_stride, int world, int rank, float* __restrict__ gsum,
                               uint32_t* __restrict__ epoch, int* __restrict__ error, float* __restrict__ m, float* __restrict__ v,
                               float* __restrict__ step, int n, float max_norm, float lr, float b1, float b2, float eps) {
  __shared__ double scratch[32];
  const uint32_t e = epoch[0] + 1u;                       // every thread reads it; CTA 0 stores the new value after the last cluster barrier
  const size_t goff = (size_t)(cur_buf[0] & 1) * (size_t)buf_stride;
  // 1. my gradient (written by the kernels before this one on the stream) is complete: tell every rank, myself included
  if (blockIdx.x == 0 && (int)threadIdx.x < world) {
    __threadfence_system();
    st_release_sys_u32(reinterpret_cast<uint32_t*>(peer_f[threadIdx.x]) + rank, e);
  }
  // 2. wait until every rank has announced this step (bounded: a missing peer raises *error instead of hanging the GPU)
  if ((int)threadIdx.x < world) {
    const uint32_t* mine = reinterpret_cast<const uint32_t*>(peer_f[rank]) + threadIdx.x;
    const long long t0 = clock64();
    while ((int32_t)(ld_acquire_sys_u32(mine) - e) < 0) {
      if (clock64() - t0 > 40000000000ll) { atomicExch(error, 1); break; }      // ~20 s: ranks may be seconds apart at start-up
      __nanosleep(200);
    }
  }
  __syncthreads();
  // 3. this CTA's share of the elements: sum over the ranks in rank order (same bits everywhere), mean -> gsum
  const float inv_w = 1.0f / (float)world;
  for (int k = blockIdx.x * 1024 + threadIdx.x; k < n; k += ADAM_CTAS * 1024) {
    float s = 0.0f;
    for (int q = 0; q < world; ++q) s += ld_relaxed_sys_f32(reinterpret_cast<const float*>(peer_g[q]) + goff + k);
    gsum[k] = s * inv_w;
  }
  __threadfence();
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  // 4. clip + Adam on the averaged gradient, as clip_adam_kernel
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
  int i = threadIdx.x;
  for (; i + 3 * 1024 < n; i += 4 * 1024) {
    const double g0 = (double)__ldcg(gsum + i), g1 = (double)__ldcg(gsum + i + 1024), g2 = (double)__ldcg(gsum + i + 2048),
                 g3 = (double)__ldcg(gsum + i + 3072);
    s0 += g0 * g0; s1 += g1 * g1; s2 += g2 * g2; s3 += g3 * g3;
  }
  for (; i < n; i += 1024) { const double g = (double)__ldcg(gsum + i); s0 += g * g; }
  const float norm = (float)sqrt(block_sum_1024((s0 + s1) + (s2 + s3), scratch));
  const float coef = fminf(max_norm / (norm + 1e-6f), 1.0f);
  const float t = step[0] + 1.0f;
  const float bc1 = 1.0f - powf(b1, t), bc2s = sqrtf(1.0f - powf(b2, t));
  const float step_size = lr / bc1;
  for (int k = blockIdx.x * 1024 + threadIdx.x; k < n; k += ADAM_CTAS * 1024) {
    const float g = __ldcg(gsum + k) * coef;
    const float mi = m[k] + (1.0f - b1) * (g - m[k]);
    const float vi = b2 * v[k] + (1.0f - b2) * g * g;
    m[k] = mi; v[k] = vi;
    p[k] -= step_size * mi / (sqrtf(vi) / bc2s + eps);
  }
  asm volatile("barrier.cluster.arrive.aligned;\n\tbarrier.cluster.wait.aligned;" ::: "memory");     // every CTA has read step[0] and epoch[0]
  if (blockIdx.x == 0 && threadIdx.x == 0) { step[0] = t; epoch[0] = e; }
}

}  // namespace

static int g_use_tc = -1;
namespace { struct NvtxRange { explicit NvtxRange(const char* n) { nvtxRangePushA(n); } ~NvtxRange() { nvtxRangePop(); } }; }

extern "C" {

int ackb_ppo_num_params(int obs_dim) { return offsets(obs_dim).total; }

int ackb_ppo_set_mode(int tensor_cores) { g_use_tc = tensor_cores ? 1 : 0; return ACKB_OK; }

int ackb_ppo_minibatch_grad(const float* obs, const float* act, const float* old_logp, const float* adv, const float* ret,
                            const int64_t* idx, int mb, int obs_dim, const float* adv_mean_std, const float* params, float* grads,
                            float* diag, float clip_range, float vf_coef, float ent_coef, void* stream) {
  return ackb_ppo_minibatch_grad_mode(obs, act, old_logp, adv, ret, idx, mb, obs_dim, adv_mean_std, params, grads, diag, clip_range, vf_coef,
                                      ent_coef, ACKB_PPO_MODE_DEFAULT, stream);
}

int ackb_ppo_minibatch_grad_mode(const float* obs, const float* act, const float* old_logp, const float* adv, const float* ret,
                                 const int64_t* idx, int mb, int obs_dim, const float* adv_mean_std, const float* params, float* grads,
                                 float* diag, float clip_range, float vf_coef, float ent_coef, int mode, void* stream) {
  return ackb_ppo_minibatch_grad_pitched(obs, obs_dim, act, old_logp, adv, ret, idx, mb, obs_dim, adv_mean_std, params, grads, diag, clip_range,
                                         vf_coef, ent_coef, mode, stream);
}

int ackb_ppo_minibatch_grad_pitched(const float* obs, int obs_pitch, const float* act, const float* old_logp, const float* adv, const float* ret,
                                    const int64_t* idx, int mb, int obs_dim, const float* adv_mean_std, const float* params, float* grads,
                                    float* diag, float clip_range, float vf_coef, float ent_coef, int mode, void* stream) {
  NvtxRange nvtx("ackb_ppo_minibatch_grad");
  const int diag_keep = (mode >= 0 && (mode & ACKB_PPO_DIAG_ACCUMULATE)) ? 1 : 0;
  if (mode >= 0) mode &= ~ACKB_PPO_DIAG_ACCUMULATE;
  if (obs_pitch < obs_dim) return ACKB_ERR_ARG;
  if (!obs || !act || !old_logp || !adv || !ret || !adv_mean_std || !params || !grads || !diag || mb <= 0) return ACKB_ERR_ARG;
  if (obs_dim <= 0 || obs_dim > KP) return ACKB_ERR_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  static bool attr_done[64] = {false};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return ACKB_ERR_NO_DEVICE;
  // ACKB_PPO_TC=0 (or ackb_ppo_set_mode(0)) selects the fp32 CUDA-core kernel; default: TF32 tensor-core kernel
  if (g_use_tc < 0) { const char* ev = getenv("ACKB_PPO_TC"); g_use_tc = ev ? (atoi(ev) != 0) : 1; }
  if (mode != ACKB_PPO_MODE_DEFAULT && mode != ACKB_PPO_MODE_FP32 && mode != ACKB_PPO_MODE_TF32 && mode != ACKB_PPO_MODE_TCGEN05) return ACKB_ERR_ARG;
  if (mode == ACKB_PPO_MODE_TCGEN05) {     // Blackwell path: tcgen05.mma kind::tf32, accumulators in TMEM (ackb_ppo_tcgen05.cu)
    PpoArgs a5{obs, act, old_logp, adv, ret, idx, mb, obs_dim, adv_mean_std, params, grads, diag, clip_range, vf_coef, ent_coef, obs_pitch, diag_keep};
    return launch_grad_tcgen05(a5, s, nullptr, nullptr);
  }
  const int use_tc = mode == ACKB_PPO_MODE_DEFAULT ? g_use_tc : (mode == ACKB_PPO_MODE_TF32 ? 1 : 0);
  const size_t smem = (size_t)(use_tc ? T_TOTAL : S_TOTAL) * sizeof(float);
  if (dev < 64 && !attr_done[dev]) {
    if (cudaFuncSetAttribute(ppo_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(S_TOTAL * sizeof(float))) != cudaSuccess) return ACKB_ERR_CUDA;
    if (cudaFuncSetAttribute(ppo_grad_kernel_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(T_TOTAL * sizeof(float))) != cudaSuccess) return ACKB_ERR_CUDA;
    attr_done[dev] = true;
  }
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const Offsets o = offsets(obs_dim);
  if (cudaMemsetAsync(grads, 0, sizeof(float) * o.total, s) != cudaSuccess) return ACKB_ERR_CUDA;
  if (!diag_keep && cudaMemsetAsync(diag, 0, sizeof(float) * 5, s) != cudaSuccess) return ACKB_ERR_CUDA;
  PpoArgs a{obs, act, old_logp, adv, ret, idx, mb, obs_dim, adv_mean_std, params, grads, diag, clip_range, vf_coef, ent_coef, obs_pitch, diag_keep};
  const int ntiles = (mb + (use_tc ? TT : TS) - 1) / (use_tc ? TT : TS);
  const int grid = ntiles < sms ? ntiles : sms;
  if (use_tc) ppo_grad_kernel_tc<<<grid, NT, smem, s>>>(a);
  else ppo_grad_kernel<<<grid, NT, smem, s>>>(a);
  return cudaGetLastError() == cudaSuccess ? ACKB_OK : ACKB_ERR_CUDA;
}

int ackb_ppo_permutation(int64_t* perm, long long n, uint64_t seed, uint32_t stream_id, void* stream) {
  if (!perm || n <= 0 || n > (1ll << 40)) return ACKB_ERR_ARG;
  int half_bits = 1;
  while ((1ull << (2 * half_bits)) < (unsigned long long)n) ++half_bits;
  permutation_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(perm, n, half_bits, seed, stream_id);
  return cudaGetLastError() == cudaSuccess ? ACKB_OK : ACKB_ERR_CUDA;
}

int ackb_ppo_minibatch_grad_stats(const float* obs, int obs_pitch, const float* act, const float* old_logp, const float* adv, const float* ret,
                                  const int64_t* idx, int mb, int obs_dim, float* adv_mean_std, double* adv_workspace, const float* params,
                                  float* grads, float* diag, float clip_range, float vf_coef, float ent_coef, int mode, void* stream) {
  if (!adv || !adv_mean_std || !adv_workspace || mb <= 0) return ACKB_ERR_ARG;
  const int diag_keep = (mode >= 0 && (mode & ACKB_PPO_DIAG_ACCUMULATE)) ? 1 : 0;
  if ((mode >= 0 ? (mode & ~ACKB_PPO_DIAG_ACCUMULATE) : mode) == ACKB_PPO_MODE_TCGEN05 && obs_dim < KP) {      // statistics in the prologue launch of the tcgen05 kernel
    NvtxRange nvtx("ackb_ppo_minibatch_grad");
    if (!obs || !act || !old_logp || !ret || !params || !grads || !diag || obs_dim <= 0 || obs_pitch < obs_dim) return ACKB_ERR_ARG;
    PpoArgs a5{obs, act, old_logp, adv, ret, idx, mb, obs_dim, adv_mean_std, params, grads, diag, clip_range, vf_coef, ent_coef, obs_pitch, diag_keep};
    return launch_grad_tcgen05(a5, (cudaStream_t)stream, adv_mean_std, adv_workspace);
  }
  const int rc = ackb_ppo_adv_stats_ws(adv, idx, mb, adv_mean_std, adv_workspace, stream);
  if (rc != ACKB_OK) return rc;
  return ackb_ppo_minibatch_grad_pitched(obs, obs_pitch, act, old_logp, adv, ret, idx, mb, obs_dim, adv_mean_std, params, grads, diag, clip_range,
                                         vf_coef, ent_coef, mode, stream);
}

int ackb_ppo_adv_stats(const float* adv, const int64_t* idx, int n, float* mean_std, void* stream) {
  double* ws = nullptr;
  if (cudaGetSymbolAddress((void**)&ws, g_adv_ws) != cudaSuccess) return ACKB_ERR_CUDA;
  return ackb_ppo_adv_stats_ws(adv, idx, n, mean_std, ws, stream);
}

int ackb_ppo_adv_stats_ws(const float* adv, const int64_t* idx, int n, float* mean_std, double* workspace, void* stream) {
  if (!adv || !mean_std || !workspace || n <= 0) return ACKB_ERR_ARG;
  int blocks = (n + 1023) / 1024;                 // >= 4 values per thread
  if (blocks > 296) blocks = 296;
  if (blocks < 1) blocks = 1;
  adv_stats_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(adv, idx, n, mean_std, workspace);
  return cudaGetLastError() == cudaSuccess ? ACKB_OK : ACKB_ERR_CUDA;
}

int ackb_ppo_clip_adam(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, float* step, int n, float max_grad_norm,
                       float lr, float beta1, float beta2, float eps, void* stream) {
  if (!params || !grads || !exp_avg || !exp_avg_sq || !step || n <= 0) return ACKB_ERR_ARG;
  clip_adam_kernel<<<ADAM_CTAS, 1024, 0, (cudaStream_t)stream>>>(params, grads, exp_avg, exp_avg_sq, step, n, max_grad_norm, lr, beta1, beta2, eps);
  return cudaGetLastError() == cudaSuccess ? ACKB_OK : ACKB_ERR_CUDA;
}

int ackb_ppo_clip_adam_allreduce(float* params, const uint64_t* peer_grad_ptrs, const uint64_t* peer_flag_ptrs, const int* cur_buf,
                                 int buf_stride, int world, int rank, float* gsum, uint32_t* epoch, int* error, float* exp_avg,
                                 float* exp_avg_sq, float* step, int n, float max_grad_norm, float lr, float beta1, float beta2,
                                 float eps, void* stream) {
  if (!params || !peer_grad_ptrs || !peer_flag_ptrs || !cur_buf || !gsum || !epoch || !error || !exp_avg || !exp_avg_sq || !step) return ACKB_ERR_ARG;
  if (n <= 0 || world < 1 || world > 32 || rank < 0 || rank >= world || buf_stride < n) return ACKB_ERR_ARG;
  NvtxRange nvtx("ackb_ppo_clip_adam_allreduce");
  clip_adam_allreduce_kernel<<<ADAM_CTAS, 1024, 0, (cudaStream_t)stream>>>(params, peer_grad_ptrs, peer_flag_ptrs, cur_buf, buf_stride, world, rank,
                                                                            gsum, epoch, error, exp_avg, exp_avg_sq, step, n, max_grad_norm, lr,
                                                                            beta1, beta2, eps);
  return cudaGetLastError() == cudaSuccess ? ACKB_OK : ACKB_ERR_CUDA;
}

int ackb_ppo_gae(const float* rew, const float* val, const float* done, const float* last_val, int n_steps, int n, float gamma,
                 float gae_lambda, float* adv, float* ret, void* stream) {
  if (!rew || !val || !done || !last_val || !adv || !ret || n_steps <= 0 || n <= 0) return ACKB_ERR_ARG;
  gae_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(rew, val, done, last_val, n_steps, n, gamma, gae_lambda, adv, ret);
  return cudaGetLastError() == cudaSuccess ? ACKB_OK : ACKB_ERR_CUDA;
}

static int launch_act(const ActArgs& a, cudaStream_t s) {
  NvtxRange nvtx(a.done_out ? "ackb_ppo_bootstrap" : "ackb_ppo_act");
  static bool attr_done[64] = {false};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return ACKB_ERR_NO_DEVICE;
  const size_t smem = (size_t)T_TOTAL * sizeof(float);
  if (dev < 64 && !attr_done[dev]) {
    if (cudaFuncSetAttribute(ppo_act_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return ACKB_ERR_CUDA;
    attr_done[dev] = true;
  }
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int ntiles = (a.n + TT - 1) / TT;
  ppo_act_kernel<<<ntiles < sms ? ntiles : sms, NT, smem, s>>>(a);
  return cudaGetLastError() == cudaSuccess ? ACKB_OK : ACKB_ERR_CUDA;
}

int ackb_ppo_act(const float* obs, int n, int obs_dim, const float* params, float* mean, float* value, float* action, float* logp,
                 uint64_t seed, uint32_t step, int value_only, void* stream) {
  return ackb_ppo_act_pitched(obs, obs_dim, n, obs_dim, params, mean, value, action, logp, seed, step, value_only, stream);
}

int ackb_ppo_act_pitched(const float* obs, int obs_pitch, int n, int obs_dim, const float* params, float* mean, float* value, float* action,
                         float* logp, uint64_t seed, uint32_t step, int value_only, void* stream) {
  if (!obs || !params || !value || n <= 0 || obs_dim <= 0 || obs_dim > KP || obs_pitch < obs_dim) return ACKB_ERR_ARG;
  // full forward on rows the bulk-copy engine can fetch (pitch a multiple of 4 floats, <= 80, 16-byte aligned base): tcgen05 / TMEM
  // kernel (ackb_ppo_tcgen05.cu); ACKB_PPO_ACT_T5=0 keeps the mma.sync kernel
  static int use_t5 = -1;
  if (use_t5 < 0) { const char* ev = getenv("ACKB_PPO_ACT_T5"); use_t5 = ev ? (atoi(ev) != 0) : 1; }
  // (value-only calls take it too: the policy net rides along, still half the time of the mma.sync kernel's value-only pass)
  if (use_t5 && (obs_pitch & 3) == 0 && obs_pitch <= KP && (reinterpret_cast<uintptr_t>(obs) & 15) == 0) {
    NvtxRange nvtx("ackb_ppo_act");
    ActT5Args a5{obs, n, obs_dim, obs_pitch, params, value_only ? nullptr : mean, value, value_only ? nullptr : action, value_only ? nullptr : logp,
                 (unsigned long long)seed, step};
    return launch_act_tcgen05(a5, (cudaStream_t)stream);
  }
  ActArgs a{obs, n, obs_dim, params, mean, value, action, logp, (unsigned long long)seed, step, value_only,
            nullptr, nullptr, nullptr, 0.0f, nullptr, nullptr, obs_pitch};
  return launch_act(a, (cudaStream_t)stream);
}

int ackb_ppo_bootstrap(const float* terminal_obs, const uint8_t* terminated, const uint8_t* truncated, const float* reward, int n,
                       int obs_dim, const float* params, float gamma, float* reward_out, float* done_out, void* stream) {
  return ackb_ppo_bootstrap_pitched(terminal_obs, obs_dim, terminated, truncated, reward, n, obs_dim, params, gamma, reward_out, done_out, stream);
}

int ackb_ppo_bootstrap_pitched(const float* terminal_obs, int obs_pitch, const uint8_t* terminated, const uint8_t* truncated, const float* reward,
                               int n, int obs_dim, const float* params, float gamma, float* reward_out, float* done_out, void* stream) {
  if (!terminal_obs || !terminated || !truncated || !reward || !params || !reward_out || !done_out || n <= 0 || obs_dim <= 0 ||
      obs_dim > KP || obs_pitch < obs_dim)
    return ACKB_ERR_ARG;
  ActArgs a{terminal_obs, n, obs_dim, params, nullptr, nullptr, nullptr, nullptr, 0ull, 0u, 1,
            terminated, truncated, reward, gamma, reward_out, done_out, obs_pitch};
  return launch_act(a, (cudaStream_t)stream);
}

}  // extern "C"
