// ackb_ppo_common.cuh -- parameter layout and kernel arguments shared by the PPO learner kernels (ackb_ppo.cu, ackb_ppo_tcgen05.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ackb_ppo {

constexpr int H = 64;        // hidden width of both MLPs
constexpr int KP = 80;       // padded observation width in shared memory (obs_dim <= KP)

struct Offsets {
  int W1p, b1p, W2p, b2p, W1v, b1v, W2v, b2v, Wa, ba, Wv, bv, ls, total;
};
__host__ __device__ inline Offsets offsets(int D) {
  Offsets o;
  o.W1p = 0; o.b1p = o.W1p + H * D; o.W2p = o.b1p + H; o.b2p = o.W2p + H * H;
  o.W1v = o.b2p + H; o.b1v = o.W1v + H * D; o.W2v = o.b1v + H; o.b2v = o.W2v + H * H;
  o.Wa = o.b2v + H; o.ba = o.Wa + 2 * H; o.Wv = o.ba + 2; o.bv = o.Wv + H; o.ls = o.bv + 1; o.total = o.ls + 2;
  return o;
}

struct PpoArgs {
  const float *obs, *act, *old_logp, *adv, *ret;
  const int64_t* idx;
  int mb, D;
  const float* adv_stats;
  const float* params;
  float* grads;
  float* diag;
  float clip, vf_coef, ent_coef;
  int pitch;       // floats between consecutive observation rows (>= D)
  int diag_keep;   // ACKB_PPO_DIAG_ACCUMULATE: diag is not zeroed by the call, this minibatch's five values are added to it
};

// tcgen05 / TMEM gradient kernel (ackb_ppo_tcgen05.cu); zeroes grads and diag itself (in its prologue launch, which also writes the
// weight images and, if adv_out / adv_ws are given, the advantage statistics of the minibatch).  Returns an ackb_status.
int launch_grad_tcgen05(const PpoArgs& a, cudaStream_t stream, float* adv_out, double* adv_ws);

// Mean and unbiased standard deviation (torch.Tensor.std) of the advantages of one minibatch, one CTA of 256 threads out of `nblocks`:
// every CTA adds its partial sums (double) to two device-scope accumulators in `ws`, the last CTA to finish writes the result and
// clears them for the next call (ws: double[3], zero-initialised once; calls sharing a workspace must not overlap).
#ifdef __CUDACC__
__device__ __forceinline__ void adv_stats_block(const float* __restrict__ adv, const int64_t* __restrict__ idx, int n, float* __restrict__ mean_std,
                                                double* __restrict__ ws, int bid, int nblocks) {
  double* acc = ws;                                           // [0..1]: sum, sum of squares
  unsigned* done_p = reinterpret_cast<unsigned*>(ws + 2);     // CTAs finished
  __shared__ double sh[2][8];
  double s = 0.0, q = 0.0;
  for (int i = bid * 256 + threadIdx.x; i < n; i += nblocks * 256) {
    const double x = (double)adv[idx ? idx[i] : (int64_t)i];
    s += x; q += x * x;
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, off); q += __shfl_xor_sync(0xffffffffu, q, off); }
  if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = s; sh[1][threadIdx.x >> 5] = q; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double ts = 0.0, tq = 0.0;
    for (int w = 0; w < 8; ++w) { ts += sh[0][w]; tq += sh[1][w]; }
    atomicAdd(&acc[0], ts); atomicAdd(&acc[1], tq);
    __threadfence();
    if (atomicAdd(done_p, 1u) == (unsigned)nblocks - 1u) {
      __threadfence();
      const double S = atomicAdd(&acc[0], 0.0), Q = atomicAdd(&acc[1], 0.0);
      const double mean = S / (double)n;
      const double var = (Q - S * mean) / (double)(n > 1 ? n - 1 : 1);
      mean_std[0] = (float)mean; mean_std[1] = (float)sqrt(var > 0.0 ? var : 0.0);
      acc[0] = 0.0; acc[1] = 0.0; *done_p = 0u;
      __threadfence();
    }
  }
}
#endif

// rollout forward on the same path (ackb_ppo_tcgen05.cu): mean [n][2] (may be null), value [n], and, if `action` is given, the sampled
// action and (if `logp` is given) its log-probability, same random stream as ppo_act_kernel.  Rows `pitch` floats apart, pitch a
// multiple of 4 and obs 16-byte aligned (the tiles are fetched with bulk copies).
struct ActT5Args {
  const float* obs;
  int n, D, pitch;
  const float* params;
  float *mean, *value, *action, *logp;
  unsigned long long seed;
  unsigned step;
};
int launch_act_tcgen05(const ActT5Args& a, cudaStream_t stream);

}  // namespace ackb_ppo
