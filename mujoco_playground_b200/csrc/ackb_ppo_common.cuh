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
};

// tcgen05 / TMEM gradient kernel (ackb_ppo_tcgen05.cu); zeroes grads and diag itself (in its weight-image launch).  Returns an ackb_status.
int launch_grad_tcgen05(const PpoArgs& a, cudaStream_t stream);

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
