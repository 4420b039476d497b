/*
 * ackb.h -- C ABI of the B200-native batched Ackermann simulator (libackb.so).
 *
 * Drop-in boundary for the env-step hot path of ulusoyn/mujoco_playground.  Each entry point
 * names the reference interface it replaces (paths relative to the reference repository).
 * Plain pointers and sizes only; no exceptions cross the ABI: every function returns 0 on
 * success and a negative ackb_status otherwise (ackb_last_error gives the text).
 *
 * One handle owns the structure-of-arrays state of `num_envs` independent environments on one
 * GPU.  All `dev_*` pointers are caller-owned DEVICE memory (e.g. torch tensors), all `host_*`
 * pointers are caller-owned HOST memory.  Calls taking a stream are asynchronous on it.  A handle
 * is not thread-safe; different handles are independent.
 */
#ifndef ACKB_H_
#define ACKB_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ackb_handle ackb_handle;

typedef enum {
  ACKB_OK = 0,
  ACKB_ERR_ARG = -1,      /* bad argument (null pointer, size mismatch, unknown dtype ...) */
  ACKB_ERR_CUDA = -2,     /* a CUDA runtime call failed */
  ACKB_ERR_NO_DEVICE = -3 /* no usable CUDA device: there is no CPU fallback */
} ackb_status;

enum { ACKB_F32 = 0, ACKB_F64 = 1 };

/* Device-side episode statistics accumulated by ackb_step since the last ackb_stats_reset
 * (replaces the per-episode r/l bookkeeping of SB3's Monitor wrapper, src/rl/train.py:70). */
typedef struct {
  unsigned long long episodes;      /* episodes finished (terminated or truncated)          */
  unsigned long long successes;     /* ... of which terminated (goal reached)               */
  unsigned long long env_steps;     /* env steps taken                                       */
  unsigned long long collisions;    /* env steps with the reference's lidar "collision" flag */
  unsigned long long unsupported;   /* env steps that left the supported contact set        */
  unsigned long long solver_iters;  /* Newton iterations summed over substeps                */
  double return_sum;                /* sum of finished episodes' returns                     */
  double length_sum;                /* sum of finished episodes' lengths                     */
  unsigned long long obstacle_steps;/* env steps with >= 1 wheel-vs-obstacle contact in their last substep (scene)   */
  unsigned long long contacts_sum;  /* contacts detected in the last substep of every env step, summed (mean ncon)   */
  unsigned long long bad_state;     /* simulation states reset by the bad-state guard (MuJoCo's mj_checkPos / mj_checkVel /
                                       mj_checkAcc -> mj_resetData: NaN or |x| > 1e10 in qpos, qvel or qacc)            */
} ackb_stats_t;

/* Size, in doubles, of the model-constants blob this build expects (layout: ackb_consts.def). */
int ackb_consts_len(void);

/* Replaces model loading in SimpleMapSpawner.load_random_environment
 * (src/rl/envs/simple_map_spawner.py:37-38: MjModel.from_xml_path + MjData) and the env constructor
 * (src/rl/envs/ackermann_env.py:51-124).  `consts` is the host blob produced by the model compiler.
 * lanes_per_env: 0 = automatic (1 lane per environment for flat-floor batches of >= 32768 environments, else 4 = one lane per
 * wheel), or 1, 4, 8 (8: one lane per floor contact, flat-floor model only). */
int ackb_create(const double* consts, size_t consts_len, int num_envs, int device, int dtype, uint64_t seed,
                int lanes_per_env, ackb_handle** out);
int ackb_destroy(ackb_handle* h);
/* Multi-GPU sharding: global id of this handle's environment 0.  Every random stream (goals, spawn jitter, maze
 * cells, synthetic actions) is keyed by (seed, GLOBAL environment id, episode / step), so a batch sharded over R handles with the
 * same seed reproduces the single-handle batch bit for bit.  Default 0.  Call before the first ackb_reset. */
int ackb_set_env_id_base(ackb_handle* h, uint64_t env_id_base);

/* Row pitch, in floats, of the dev_obs / dev_terminal_obs arrays of ackb_reset / ackb_step (default obs_dim = dense rows).  A pitch
 * of 80 makes every 79-float observation row start on a 16-byte boundary: the PPO learner then fetches rollout rows with vector
 * loads and whole 32-byte sectors (include/ackb_ppo.h, *_pitched entry points).  The host arrays of ackb_step_host stay dense. */
int ackb_set_obs_pitch(ackb_handle* h, int pitch_floats);

int ackb_num_envs(const ackb_handle* h);
int ackb_obs_dim(const ackb_handle* h);   /* 79 for ackermann_robot_v2 and the maze models (ackermann_env.py:95-100), 43 for the scene */
int ackb_dtype(const ackb_handle* h);

/* Replaces AckermannRobotEnv.reset (src/rl/envs/ackermann_env.py:143-185) for every environment, or for
 * those with dev_mask[i] != 0.  Writes the fresh observation rows of the reset environments into dev_obs
 * ([num_envs, obs_dim] float32); other rows are left untouched. */
int ackb_reset(ackb_handle* h, const uint8_t* dev_mask_or_null, float* dev_obs, void* stream);

/* Replaces AckermannRobotEnv.step (src/rl/envs/ackermann_env.py:187-229): action clip/scale,
 * BicycleController.apply_cmd_vel (src/core/controller.py:136-140), frame_skip x mujoco.mj_step (:200),
 * _get_observation (:231-265), _calculate_reward (:267-312), step/truncation bookkeeping (:216-220),
 * and -- when auto_reset != 0 -- the auto-reset that SB3's DummyVecEnv performs around it (src/rl/train.py:76):
 * a finished environment is reset in the same call, dev_obs gets the first observation of the new episode and
 * dev_terminal_obs (if given) the last observation of the finished one.
 *   dev_action   [num_envs, 2] float32, or NULL: synthetic U(-1,1) actions from Philox(seed, step, env)
 *   dev_obs      [num_envs, obs_dim] float32
 *   dev_reward   [num_envs] float32; dev_terminated / dev_truncated [num_envs] uint8
 *   dev_terminal_obs  [num_envs, obs_dim] float32 or NULL (rows of finished environments only)
 *   dev_ncon     [num_envs] int32 or NULL: contacts detected in the last substep (MuJoCo's data.ncon) */
int ackb_step(ackb_handle* h, const float* dev_action, int frame_skip, int auto_reset, float* dev_obs, float* dev_reward,
              uint8_t* dev_terminated, uint8_t* dev_truncated, float* dev_terminal_obs, int32_t* dev_ncon, void* stream);

/* Same call with HOST buffers, synchronising before it returns: the end-to-end path a host-side caller such as
 * src/rl/train.py:199-205 sees.  Pinned buffers (cudaHostAlloc / cudaHostRegister) are read and written by the kernel directly
 * (mapped host memory, no staging copies); pageable buffers go through cudaMemcpyAsync staging.
 * Stream order: the call runs on a stream of the handle and returns after it completed; it waits for the ackb_step / ackb_reset
 * work the caller issued last on another stream.  Calls on two different caller streams are the caller's to order. */
int ackb_step_host(ackb_handle* h, const float* host_action, int frame_skip, int auto_reset, float* host_obs, float* host_reward,
                   uint8_t* host_terminated, uint8_t* host_truncated);

/* State access for parity tests (replaces reads/writes of data.qpos / data.qvel / data.qacc_warmstart).
 * Host arrays, row major, MuJoCo ordering: qpos [num_envs,13], qvel [num_envs,12], warm [num_envs,12]. Synchronous. */
int ackb_get_state(ackb_handle* h, double* host_qpos, double* host_qvel, double* host_warm);
int ackb_set_state(ackb_handle* h, const double* host_qpos, const double* host_qvel, const double* host_warm);
/* Episode data: goal [num_envs,2] and odometry reference [num_envs,2] (ackermann_env.py:163-172), step_count [num_envs]. */
int ackb_get_episode(ackb_handle* h, double* host_goal, double* host_ref, int32_t* host_step_count);
int ackb_set_episode(ackb_handle* h, const double* host_goal, const double* host_ref, const int32_t* host_step_count);

int ackb_stats(ackb_handle* h, ackb_stats_t* out); /* synchronous */
int ackb_stats_reset(ackb_handle* h);
/* Kernels launched by this handle so far (for bench.py's gpu_launches). */
unsigned long long ackb_launch_count(const ackb_handle* h);
/* Fill dev_action [num_envs,2] with the same synthetic actions ackb_step(NULL action) would draw at its next call. */
int ackb_random_actions(ackb_handle* h, float* dev_action, void* stream);

const char* ackb_last_error(const ackb_handle* h_or_null);

#ifdef __cplusplus
}
#endif
#endif /* ACKB_H_ */
