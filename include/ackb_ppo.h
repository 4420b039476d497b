/* ackb_ppo.h -- C ABI of the fused PPO minibatch-gradient kernel (learner side of config 5).
 *
 * Replaces, for one minibatch, what the reference obtains from Stable-Baselines3's PPO.train() inner loop
 * (stable_baselines3/ppo/ppo.py: evaluate_actions -> ratio / clipped surrogate / value loss / entropy -> loss.backward();
 * called through model.learn, src/rl/train.py:175-179) for the MlpPolicy the reference trains
 * (separate tanh MLPs obs -> 64 -> 64 for policy and value, state-independent log_std; 18 757 parameters for obs = 79).
 * The optimiser step (gradient clipping + Adam) is a separate entry point, ackb_ppo_clip_adam, on caller-owned state.
 *
 * All pointers are caller-owned DEVICE memory.  Parameter / gradient layout (floats, D = obs_dim):
 *   W1p[64][D] b1p[64] W2p[64][64] b2p[64]  W1v[64][D] b1v[64] W2v[64][64] b2v[64]  Wa[2][64] ba[2]  Wv[64] bv[1]  log_std[2]
 * i.e. the parameters of mujoco_playground_b200.ppo.ActorCritic in state_dict order with log_std moved to the end.
 */
#ifndef ACKB_PPO_H_
#define ACKB_PPO_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* number of floats of the parameter / gradient vector for an observation of obs_dim floats */
int ackb_ppo_num_params(int obs_dim);

/* Arithmetic of the gradient kernel: 1 = TF32 tensor-core mma with fp32 accumulation (default; same operand precision as
 * torch.backends.cuda.matmul.allow_tf32), 0 = fp32 CUDA-core FMAs.  Process-wide; the environment variable ACKB_PPO_TC sets the
 * initial value. */
int ackb_ppo_set_mode(int tensor_cores);

/* Gradient of the PPO loss  mean_i[-min(A_i r_i, A_i clip(r_i, 1-c, 1+c))] + vf_coef * mean_i (V_i - R_i)^2 - ent_coef * H
 * over the minibatch rows idx[0 .. mb-1] of the rollout arrays (A normalised with adv_mean_std = {mean, std} of the minibatch,
 * A = (adv - mean) / (std + 1e-8)).  grads (zeroed by the call) receives d loss / d params; diag receives
 * {policy loss, value loss (mean squared error), entropy, approx_kl, clip fraction}.  Asynchronous on `stream`.
 * Returns 0 or a negative error code (same codes as ackb.h). */
int ackb_ppo_minibatch_grad(const float* obs, const float* act, const float* old_logp, const float* adv, const float* ret,
                            const int64_t* idx, int mb, int obs_dim, const float* adv_mean_std, const float* params, float* grads,
                            float* diag, float clip_range, float vf_coef, float ent_coef, void* stream);

/* Same call with the arithmetic chosen per call (nothing process-wide is read or written): one of ACKB_PPO_MODE_*. */
enum {
  ACKB_PPO_MODE_DEFAULT = -1,
  ACKB_PPO_MODE_FP32 = 0,      /* fp32 CUDA-core FMAs (exact-fp32 comparisons)                                                  */
  ACKB_PPO_MODE_TF32 = 1,      /* TF32 mma.sync m16n8k8 fragments (round-1 kernel)                                              */
  ACKB_PPO_MODE_TCGEN05 = 2,   /* TF32 tcgen05.mma: operands by shared-memory descriptor, accumulators in TMEM (obs_dim < 80)   */
  /* flag, OR-ed onto one of the three modes above: diag is NOT zeroed by the call; the minibatch's five values are added to it
   * (zero it once per update and divide by the number of optimiser steps: no per-step accumulation launch on the caller's side) */
  ACKB_PPO_DIAG_ACCUMULATE = 0x100
};
int ackb_ppo_minibatch_grad_mode(const float* obs, const float* act, const float* old_logp, const float* adv, const float* ret,
                                 const int64_t* idx, int mb, int obs_dim, const float* adv_mean_std, const float* params, float* grads,
                                 float* diag, float clip_range, float vf_coef, float ent_coef, int mode, void* stream);

/* Pitched variants: observation rows obs_pitch floats apart (>= obs_dim).  With obs_pitch = 80 every 79-float row of the rollout
 * buffer starts on a 16-byte boundary (ackb.h: ackb_set_obs_pitch): the tcgen05 gradient kernel then gathers rows with 16-byte
 * vector loads over whole 32-byte sectors (column 79 of the buffer is never read). */
int ackb_ppo_minibatch_grad_pitched(const float* obs, int obs_pitch, const float* act, const float* old_logp, const float* adv, const float* ret,
                                    const int64_t* idx, int mb, int obs_dim, const float* adv_mean_std, const float* params, float* grads,
                                    float* diag, float clip_range, float vf_coef, float ent_coef, int mode, void* stream);
/* ackb_ppo_adv_stats_ws + ackb_ppo_minibatch_grad_pitched in one call: adv_mean_std (OUT, then used by the gradient) = mean / std of
 * adv over the minibatch, adv_workspace as for ackb_ppo_adv_stats_ws.  In ACKB_PPO_MODE_TCGEN05 the statistics are computed by extra
 * CTAs of the kernel's prologue launch (weight images + gradient zeroing), which saves a launch per optimiser step. */
int ackb_ppo_minibatch_grad_stats(const float* obs, int obs_pitch, const float* act, const float* old_logp, const float* adv, const float* ret,
                                  const int64_t* idx, int mb, int obs_dim, float* adv_mean_std, double* adv_workspace, const float* params,
                                  float* grads, float* diag, float clip_range, float vf_coef, float ent_coef, int mode, void* stream);
int ackb_ppo_act_pitched(const float* obs, int obs_pitch, int n, int obs_dim, const float* params, float* mean, float* value, float* action,
                         float* logp, uint64_t seed, uint32_t step, int value_only, void* stream);
int ackb_ppo_bootstrap_pitched(const float* terminal_obs, int obs_pitch, const uint8_t* terminated, const uint8_t* truncated, const float* reward,
                               int n, int obs_dim, const float* params, float gamma, float* reward_out, float* done_out, void* stream);

/* Rollout side (SB3 policy.forward() inside collect_rollouts): value[n] = V(obs) and, unless value_only, mean[n][2] = pi(obs);
 * if `action` is given also a sample action = mean + exp(log_std) * eps (eps ~ N(0,1), Philox keyed by (seed, step, row)) and,
 * if `logp` is given, its log-probability.  mean / action / logp may be NULL.  TF32 tensor-core tiles.  Asynchronous on `stream`. */
int ackb_ppo_act(const float* obs, int n, int obs_dim, const float* params, float* mean, float* value, float* action, float* logp,
                 uint64_t seed, uint32_t step, int value_only, void* stream);

/* Time-limit bootstrap of one rollout step (SB3 collect_rollouts: rewards[idx] += gamma * V(terminal_observation) where the
 * episode was truncated but not terminated): reward_out[n] = reward + gamma * V(terminal_obs) for those rows, reward elsewhere;
 * done_out[n] = (terminated | truncated) as float.  The value network only runs for 32-row tiles that contain such a row. */
int ackb_ppo_bootstrap(const float* terminal_obs, const uint8_t* terminated, const uint8_t* truncated, const float* reward, int n,
                       int obs_dim, const float* params, float gamma, float* reward_out, float* done_out, void* stream);

/* Pseudo-random permutation of 0 .. n-1 into perm[n] (the per-epoch shuffle of SB3's RolloutBuffer.get, np.random.permutation):
 * a keyed Feistel bijection with cycle walking, one thread per element, no sort.  (seed, stream_id) select the permutation. */
int ackb_ppo_permutation(int64_t* perm, long long n, uint64_t seed, uint32_t stream_id, void* stream);

/* {mean, unbiased std} of the advantages of one minibatch (rows idx[0 .. n-1] of adv, or the first n rows if idx is NULL): the
 * adv_mean_std operand of ackb_ppo_minibatch_grad (SB3 normalises advantages per minibatch, ppo.py train()).  Calls on one
 * device share a pair of device-side accumulators: issue them on one stream (or otherwise ordered). */
int ackb_ppo_adv_stats(const float* adv, const int64_t* idx, int n, float* mean_std, void* stream);
/* Same with caller-owned accumulators: `workspace` = 3 doubles of device memory, zeroed once by the caller (the call leaves them
 * zeroed).  Learners with their own workspace may run concurrently on one device. */
int ackb_ppo_adv_stats_ws(const float* adv, const int64_t* idx, int n, float* mean_std, double* workspace, void* stream);

/* Optimiser step on the flat parameter vector: global-norm clipping of grads to max_grad_norm (torch.nn.utils.clip_grad_norm_)
 * followed by Adam (torch.optim.Adam without weight decay / amsgrad: SB3's PPO optimiser).  exp_avg / exp_avg_sq / step are the
 * optimiser state (step: one float on the device, incremented by the call).  grads is not modified. */
int ackb_ppo_clip_adam(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, float* step, int n, float max_grad_norm,
                       float lr, float beta1, float beta2, float eps, void* stream);

/* The same optimiser step for data-parallel learners on the GPUs of one node, with the gradient all-reduce INSIDE the kernel (no
 * NCCL call between the gradient kernel and the optimiser step; SB3 has one learner, the reference's multi-GPU counterpart is the
 * 75 KB all-reduce of SURVEY.md section 8(e)).  Every rank passes
 *   peer_grad_ptrs  device array [world] of addresses: rank q's gradient storage, mapped into this process (CUDA IPC / symmetric
 *                   memory), two buffers of buf_stride floats each; *cur_buf (device int, 0 / 1) selects the buffer that the
 *                   gradient kernels of ALL ranks wrote for this step -- consecutive steps must alternate the buffers;
 *   peer_flag_ptrs  device array [world] of addresses of each rank's flag words (uint32[world], zero-initialised, peer mapped);
 *   epoch           this rank's step counter (device uint32, starts at 0, incremented by the call);
 *   gsum            scratch of n floats (this rank).
 * The kernel announces "my gradient is complete" on every peer's flags, waits for all peers, sums the world gradients in rank
 * order over NVLink peer loads (identical bits on every rank), divides by world, then clips and applies Adam as ackb_ppo_clip_adam.
 * A peer that does not arrive within ~20 s sets *error (device int) instead of hanging.  Asynchronous on `stream`. */
int ackb_ppo_clip_adam_allreduce(float* params, const uint64_t* peer_grad_ptrs, const uint64_t* peer_flag_ptrs, const int* cur_buf,
                                 int buf_stride, int world, int rank, float* gsum, uint32_t* epoch, int* error, float* exp_avg,
                                 float* exp_avg_sq, float* step, int n, float max_grad_norm, float lr, float beta1, float beta2,
                                 float eps, void* stream);

/* Generalised advantage estimation over a rollout (SB3 RolloutBuffer.compute_returns_and_advantage, called from
 * collect_rollouts): rew / val / done / adv / ret are [n_steps][n] device arrays, done[t] = 1 if the episode ended after step t,
 * last_val[n] = V(observation after the last step).  ret = adv + val.  Asynchronous on `stream`. */
int ackb_ppo_gae(const float* rew, const float* val, const float* done, const float* last_val, int n_steps, int n, float gamma,
                 float gae_lambda, float* adv, float* ret, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ACKB_PPO_H_ */
