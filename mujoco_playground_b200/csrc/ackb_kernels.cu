// ackb_kernels.cu -- sm_100a kernels and the C ABI (include/ackb.h) of the batched Ackermann simulator.
//
// HBM layout (per handle): structure of arrays, env index fastest:
//   qpos[13][N] qvel[12][N] warm[12][N] goal[2][N] ref[2][N]   (T = float | double)
//   step_count[N] i32, episode[N] u32, ep_return[N] f32
// One fused kernel per env.step(): LANES lanes per environment (4 = one per wheel), state loaded once,
// frame_skip substeps in registers, observation staged per warp in shared memory and written with
// coalesced stores, auto-reset fused.  No tensor cores: the largest dense object is 8 x 8.
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>   // header-only NVTX v3: ranges cost one pointer test unless a profiler (nsys / ncu --nvtx) is attached
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <new>
#include <string>

#include "ackb.h"
#include "ackb_env.cuh"

using namespace ackb;

namespace {

// Model constants travel as a __grid_constant__ kernel parameter (2.7 KB in fp32, 5.5 KB in fp64; the parameter space is
// constant bank 0, so reads are the same broadcast LDC / c[0][..] operands a __constant__ symbol gives).  Every handle owns its
// own host copy: handles of different models or precisions on one device never share or re-upload anything, and a launch
// needs no synchronisation (stream-capture safe).

template <typename T>
struct DevState {
  T *qpos, *qvel, *warm, *goal, *ref;
  int32_t* step_count;
  uint32_t* episode;
  float* ep_return;
  int n;
};

struct StepArgs {
  const float* action;
  float* obs;
  float* reward;
  uint8_t* terminated;
  uint8_t* truncated;
  float* terminal_obs;
  int32_t* ncon;
  const uint8_t* mask;
  ackb_stats_t* stats;
  unsigned long long seed;
  uint32_t step_index;
  int frame_skip, auto_reset, obs_dim;
  int obs_pitch;        // floats between consecutive rows of obs / terminal_obs (>= obs_dim; ackb_set_obs_pitch)
  uint8_t* done_mask;   // defer_reset: per-env done flag for the masked reset launch that follows
  int defer_reset;      // models with settle steps: the step kernel only marks finished environments, reset_kernel does the rest
  uint32_t env_base;    // global id of environment 0 of this handle (Philox streams are keyed by the GLOBAL environment id)
  // Regime split of the flat-floor model: its fast kernels carry two floor-contact slots per wheel.  Environments tilted beyond
  // kGeneralTilt at the start of a step (rolled over, on their side, nose down: wheel caps or chassis plates can reach the floor) are
  // left to a second launch of the general kernel (NC = 4: two extra contact slots per wheel).  regime 0: every environment;
  // 1: fast pass (skips the tilted ones and lists them); 2: general pass (only the listed ones).
  // The fast pass appends the tilted environments to gen_list (length in gen_count[0]); the general pass is a small fixed grid
  // that walks the list, leaves at once when it is empty, and zeroes the counter for the next step (the next fast pass starts
  // only after this grid has finished).
  int regime;
  int* gen_list;
  int* gen_count;
  int cta_sync;   // multi-lane kernels: re-converge the CTA once per substep (pays off only when several warps share a scheduler)
};

constexpr float kGeneralTilt = 0.8f;   // cos(36.9 deg): a plate needs > 60 deg of tilt to reach the floor, a wheel cap ~88 deg

template <typename T>
struct SoAAcc {
  const DevState<T>& s;
  int env;
  __device__ __forceinline__ T qpos(int i) const { return s.qpos[(size_t)i * s.n + env]; }
  __device__ __forceinline__ T qvel(int i) const { return s.qvel[(size_t)i * s.n + env]; }
  __device__ __forceinline__ T warm(int i) const { return s.warm[(size_t)i * s.n + env]; }
  __device__ __forceinline__ void set_qpos(int i, T v) { s.qpos[(size_t)i * s.n + env] = v; }
  __device__ __forceinline__ void set_qvel(int i, T v) { s.qvel[(size_t)i * s.n + env] = v; }
  __device__ __forceinline__ void set_warm(int i, T v) { s.warm[(size_t)i * s.n + env] = v; }
};
template <bool ALIAS>
struct RowSink {
  static constexpr bool kAliasesWheels = ALIAS;
  float* row;
  __device__ __forceinline__ void put(int slot, float v) { row[slot] = v; }
};
struct PredRowSink {
  static constexpr bool kAliasesWheels = false;
  float* row;
  bool on;
  __device__ __forceinline__ void put(int slot, float v) { if (on) row[slot] = v; }
};

__device__ __forceinline__ unsigned warp_sum_u32(unsigned v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Per-(T, LANES) launch geometry.  With fewer than 4 lanes per environment the wheel records of a thread live in
// shared memory with an odd per-thread stride (conflict free) and the per-warp observation tile aliases that region
// (the tile is flushed to HBM before the dynamics of the last substep reuse it).
template <typename T, int LANES, int NC>
struct Geo {
  static constexpr int WPL = LANES == 8 ? 1 : 4 / LANES;
  static constexpr bool kSmemWheels = !Sim<T, LANES, NC>::kRegRecords;
  static constexpr int kWheelUnits = sizeof(Wheel<T, NC>) / sizeof(T);          // record size in units of T
  static constexpr int kStride = WPL * kWheelUnits + ((WPL * kWheelUnits) % 2 == 0 ? 1 : 0);   // odd
  // shared-memory layouts: pick the CTA size (128, 64 or 32 threads) that fits the most warps into the 227 KB of an SM
  // (1 KB per CTA is reserved by the system); registers allow 2 CTAs of 128 threads at most (<= 255 registers per thread)
  static constexpr int warps_per_sm(int block) {
    const long per_cta = (long)block * kStride * (long)sizeof(T) + 1024;
    long ctas = (227L * 1024) / per_cta;
    const long reg_ctas = 65536 / (232L * block);      // register allocation of this kernel: 232 per thread
    if (ctas > reg_ctas) ctas = reg_ctas;
    return (int)(ctas * block / 32);
  }
  static constexpr int pick_block() {
#ifdef ACKB_L1_BLOCK
    return sizeof(T) == 4 ? ACKB_L1_BLOCK : ACKB_L1_BLOCK / 2;
#else
    int best = sizeof(T) == 4 ? 128 : 64;
    if (sizeof(T) == 4) {   // fp64 kernels sit at the 255-register cap and measured slower with smaller CTAs
      if (warps_per_sm(64) > warps_per_sm(best)) best = 64;
      if (warps_per_sm(32) > warps_per_sm(best)) best = 32;
    }
    return best;
#endif
  }
#ifndef ACKB_L1_MINB
#define ACKB_L1_MINB (warps_per_sm(pick_block()) * 32 / pick_block())
#endif
#ifndef ACKB_L4_BLOCK
#define ACKB_L4_BLOCK 128
#endif
  static constexpr int kBlock = kSmemWheels ? pick_block() : ACKB_L4_BLOCK;
  static constexpr int kMinBlocks = kSmemWheels ? ACKB_L1_MINB : (256 / ACKB_L4_BLOCK);
  static size_t smem_bytes(int obs_dim) {
    const size_t tile = (size_t)(kBlock / LANES) * obs_dim * sizeof(float);
    const size_t wheels = kSmemWheels ? (size_t)kBlock * kStride * sizeof(T) : 0;
    return tile > wheels ? tile : wheels;
  }
};
static_assert(sizeof(Wheel<float, 2>) % sizeof(float) == 0 && sizeof(Wheel<double, 2>) % sizeof(double) == 0 && sizeof(Wheel<double, 4>) % sizeof(double) == 0, "Wheel must be a whole number of T");

// synthetic action of (step, env): Philox(seed) with counter (step, env, 2, tag) -> U(-1, 1)^2
__device__ __forceinline__ void synth_action(unsigned long long seed, uint32_t step, uint32_t env, float* a0, float* a1) {
  uint32_t r[4];
  philox4x32(step, env, 2u, 0x41434B42u, (uint32_t)seed, (uint32_t)(seed >> 32), r);
  *a0 = 2.0f * u01(r[0]) - 1.0f;
  *a1 = 2.0f * u01(r[1]) - 1.0f;
}

// wheel records of this thread: registers (4 lanes per env) or the thread's slice of shared memory
template <typename T, int LANES, int NC>
struct WheelStore {
  Wheel<T, NC> reg[Geo<T, LANES, NC>::kSmemWheels ? 1 : Geo<T, LANES, NC>::WPL];
  __device__ __forceinline__ Wheel<T, NC>* get(unsigned char* smem) {
    if constexpr (Geo<T, LANES, NC>::kSmemWheels) return reinterpret_cast<Wheel<T, NC>*>(reinterpret_cast<T*>(smem) + (size_t)threadIdx.x * Geo<T, LANES, NC>::kStride);
    else return reg;
  }
};

// LIST = general pass of the regime split: the environments come from the list the fast pass wrote (blk = virtual block index of
// the grid-stride loop in step_kernel_list, count = list length); otherwise thread -> environment is the identity map.
template <typename T, int LANES, int NC, bool LIST>
__device__ __forceinline__ void step_body(DevState<T>& st, const StepArgs& a, const Consts<T>& C, int blk, int count) {
  using E = EnvOps<T, LANES, NC>;
  using G = Geo<T, LANES, NC>;
  constexpr int EPW = 32 / LANES;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = blk * blockDim.x + threadIdx.x;
  const int env_raw = tid / LANES, lane = tid % LANES;
  const bool valid_in = env_raw < (LIST ? count : st.n);
  const int env = LIST ? a.gen_list[valid_in ? env_raw : 0] : (valid_in ? env_raw : st.n - 1);
  const int warp = threadIdx.x >> 5, lid = threadIdx.x & 31;
#if defined(__CUDA_ARCH__)
  // programmatic dependent launch of the general pass: it may be scheduled while this grid drains
  if (!LIST && a.regime == 1) asm volatile("griddepcontrol.launch_dependents;");
#endif
  if (LIST && !__any_sync(0xffffffffu, valid_in)) return;   // no CTA barriers in this pass (cta_sync = 0)
  if (!LIST) stage_tables(C);
  // per-warp observation tile: its own region (4 lanes) or aliasing the warp's wheel records
  float* wtile = G::kSmemWheels ? reinterpret_cast<float*>(reinterpret_cast<T*>(smem_raw) + (size_t)warp * 32 * G::kStride)
                                : reinterpret_cast<float*>(smem_raw) + (size_t)warp * EPW * a.obs_dim;
  RowSink<G::kSmemWheels> sink{wtile + (lid / LANES) * a.obs_dim};
  WheelStore<T, LANES, NC> store;
  Wheel<T, NC>* wh = store.get(smem_raw);

  typename E::State e;
  SoAAcc<T> acc{st, env};
  E::load_state(acc, lane, e, wh);
  bool valid = valid_in;
  if (!LIST && a.regime == 1) {     // CTA-uniform
    // cosine of the tilt angle = R_zz of the chassis = (w^2 - x^2 - y^2 + z^2) / |q|^2
    const T qq = e.q[0] * e.q[0] + e.q[1] * e.q[1] + e.q[2] * e.q[2] + e.q[3] * e.q[3];
    const bool general = !((e.q[0] * e.q[0] - e.q[1] * e.q[1] - e.q[2] * e.q[2] + e.q[3] * e.q[3]) >= T(kGeneralTilt) * qq);   // NaN -> general
    if (general && valid_in && lane == 0) {      // rare: left to the general pass
      const int slot = atomicAdd(a.gen_count, 1);
      if (slot < st.n) a.gen_list[slot] = env;
    }
    valid = valid_in && !general;
  }
  Episode<T> ep;
  ep.goal[0] = st.goal[env]; ep.goal[1] = st.goal[(size_t)st.n + env];
  ep.ref[0] = st.ref[env]; ep.ref[1] = st.ref[(size_t)st.n + env];
  ep.step_count = st.step_count[env];
  ep.episode = st.episode[env];
  float a0, a1;
  if (a.action) { float2 v = reinterpret_cast<const float2*>(a.action)[env]; a0 = v.x; a1 = v.y; }
  else synth_action(a.seed, a.step_index, (uint32_t)env + a.env_base, &a0, &a1);

  const int env0 = (blk * blockDim.x + warp * 32) / LANES;     // LIST: position in the list
  int nrow = (LIST ? count : st.n) - env0;
  nrow = nrow < 0 ? 0 : (nrow > EPW ? EPW : nrow);
  // coalesced store of the warp's observation tile (the warp's environments are consecutive rows), issued as soon as
  // the observation exists so that the tile storage can be reused by the last substep
  // rows of this warp's tile that belong to environments this pass owns (regime split): bit (row * LANES)
  const unsigned rowmask = __ballot_sync(0xffffffffu, valid && lane == 0);
  const unsigned fullmask = __ballot_sync(0xffffffffu, valid_in && lane == 0);
  auto emit = [&]() {
    __syncwarp();
    if constexpr (LIST) {                // rows of listed environments: scattered
      for (int r = 0; r < nrow; ++r) {
        float* drow = a.obs + (size_t)__shfl_sync(0xffffffffu, env, r * LANES) * a.obs_pitch;
        for (int j = lid; j < a.obs_dim; j += 32) drow[j] = wtile[r * a.obs_dim + j];
      }
      __syncwarp();
      return;
    }
    float* dst = a.obs + (size_t)env0 * a.obs_pitch;
    if (rowmask != fullmask) {           // warp-uniform: some rows belong to the other pass
      for (int r = 0; r < nrow; ++r)
        if ((rowmask >> (r * LANES)) & 1u)
          for (int j = lid; j < a.obs_dim; j += 32) dst[(size_t)r * a.obs_pitch + j] = wtile[r * a.obs_dim + j];
    } else if (a.obs_pitch == a.obs_dim) {      // warp-uniform: contiguous rows, one flat coalesced copy
      const int total = nrow * a.obs_dim;
      for (int i = lid; i < total; i += 32) dst[i] = wtile[i];
    } else {                             // pitched rows (e.g. 80 floats: 16-byte aligned rows for the learner's vector loads)
      for (int r = 0; r < nrow; ++r)
        for (int j = lid; j < a.obs_dim; j += 32) dst[(size_t)r * a.obs_pitch + j] = wtile[r * a.obs_dim + j];
    }
    __syncwarp();
  };

  StepOut<T> out;
  StepDiag diag{};
  E::step_env(C, e, wh, ep, a0, a1, a.frame_skip, lane, sink, emit, out, diag, (DebugTap<T>*)nullptr, a.cta_sync != 0, G::kSmemWheels ? G::kStride : 0);
  __syncwarp();

  // contact count of the last substep, summed over the lanes of the environment
  int ncon = diag.ncon, unsup = diag.unsupported;
  ncon = Team<LANES>::sum(ncon);
  unsup = Team<LANES>::sum(unsup);
  const int nbox = Team<LANES>::sum(diag.nbox);

  const bool done = out.terminated || out.truncated;
  float ret = st.ep_return[env] + out.reward;
  const int ep_len = ep.step_count;

  // statistics: one atomic per warp and counter
  if (a.stats) {
    const bool lead = valid && lane == 0;
    unsigned v_done = warp_sum_u32(lead && done ? 1u : 0u), v_succ = warp_sum_u32(lead && out.terminated ? 1u : 0u);
    unsigned v_coll = warp_sum_u32(lead && out.collision ? 1u : 0u), v_uns = warp_sum_u32(lead && unsup ? 1u : 0u);
    unsigned v_it = warp_sum_u32(lead ? (unsigned)diag.niter : 0u);
    unsigned v_box = warp_sum_u32(lead && nbox > 0 ? 1u : 0u), v_con = warp_sum_u32(lead ? (unsigned)ncon : 0u);
    unsigned v_bad = warp_sum_u32(lead ? (unsigned)diag.bad : 0u);
    float r_sum = lead && done ? ret : 0.f, l_sum = lead && done ? (float)ep_len : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { r_sum += __shfl_xor_sync(0xffffffffu, r_sum, o); l_sum += __shfl_xor_sync(0xffffffffu, l_sum, o); }
    if (lid == 0) {
      if (v_done) { atomicAdd(&a.stats->episodes, (unsigned long long)v_done); atomicAdd(&a.stats->return_sum, (double)r_sum); atomicAdd(&a.stats->length_sum, (double)l_sum); }
      if (v_succ) atomicAdd(&a.stats->successes, (unsigned long long)v_succ);
      if (v_coll) atomicAdd(&a.stats->collisions, (unsigned long long)v_coll);
      if (v_uns) atomicAdd(&a.stats->unsupported, (unsigned long long)v_uns);
      atomicAdd(&a.stats->solver_iters, (unsigned long long)v_it);
      if (v_box) atomicAdd(&a.stats->obstacle_steps, (unsigned long long)v_box);
      atomicAdd(&a.stats->contacts_sum, (unsigned long long)v_con);
      if (v_bad) atomicAdd(&a.stats->bad_state, (unsigned long long)v_bad);
    }
  }

  const bool do_reset = valid && done && a.auto_reset && !a.defer_reset;
  if (a.defer_reset) {   // CTA-uniform
    if (valid && done && a.terminal_obs)   // the finished episode's last observation was flushed to dev_obs by emit()
      for (int j = lane; j < a.obs_dim; j += LANES) a.terminal_obs[(size_t)env * a.obs_pitch + j] = a.obs[(size_t)env * a.obs_pitch + j];
    if (valid && lane == 0) a.done_mask[env] = (done && a.auto_reset) ? 1 : 0;
  }
  if (__any_sync(0xffffffffu, do_reset)) {   // warp-uniform: the team collectives inside need the whole warp
    if (do_reset && valid) {
      if (a.terminal_obs)   // the finished episode's last observation was flushed to dev_obs by emit()
        for (int j = lane; j < a.obs_dim; j += LANES) a.terminal_obs[(size_t)env * a.obs_pitch + j] = a.obs[(size_t)env * a.obs_pitch + j];
    }
    __syncwarp();
    if (do_reset) E::reset_env(C, e, wh, ep, lane, a.seed, (uint32_t)env + a.env_base);
    Kin<T> k;
    E::S::kinematics(e, k);
    T dist, minl;
    PredRowSink psink{sink.row, do_reset};
    // the wheel records in shared memory are about to be overwritten by the tile: keep the spin state in registers
    T keep[3 * G::WPL];
    if (G::kSmemWheels) for (int s = 0; s < G::WPL; ++s) { keep[3 * s] = wh[s].sp; keep[3 * s + 1] = wh[s].dsp; keep[3 * s + 2] = wh[s].warm; }
    __syncwarp();
    E::observe(C, e, k, ep, lane, psink, &dist, &minl);
    __syncwarp();
    if (do_reset && valid)
      for (int j = lane; j < a.obs_dim; j += LANES) a.obs[(size_t)env * a.obs_pitch + j] = sink.row[j];
    __syncwarp();
    if (G::kSmemWheels) for (int s = 0; s < G::WPL; ++s) { wh[s].sp = keep[3 * s]; wh[s].dsp = keep[3 * s + 1]; wh[s].warm = keep[3 * s + 2]; }
  }
  __syncwarp();

  if (valid) {
    E::store_state(acc, lane, e, wh);
    if (lane == 0) {
      st.goal[env] = ep.goal[0]; st.goal[(size_t)st.n + env] = ep.goal[1];
      st.ref[env] = ep.ref[0]; st.ref[(size_t)st.n + env] = ep.ref[1];
      st.step_count[env] = ep.step_count;
      st.episode[env] = ep.episode;
      st.ep_return[env] = do_reset ? 0.f : ret;
      a.reward[env] = out.reward;
      a.terminated[env] = out.terminated;
      a.truncated[env] = out.truncated;
      if (a.ncon) a.ncon[env] = ncon;
    }
  }
}

// (A device-side tail launch of the general pass from the last CTA of the fast pass -- CUDA dynamic parallelism, -rdc=true -- was
// built and measured: 0.892 ms per 131072-env step against 0.873 ms for the host-side dependent launch below; the relocatable build
// slows the fast kernel by more than the second launch costs.  Removed.)
template <typename T, int LANES, int NC>
__global__ void __launch_bounds__(Geo<T, LANES, NC>::kBlock, Geo<T, LANES, NC>::kMinBlocks) step_kernel(DevState<T> st, StepArgs a, const __grid_constant__ Consts<T> C) {
  step_body<T, LANES, NC, false>(st, a, C, (int)blockIdx.x, 0);
}

// General pass of the regime split (flat-floor model): a small fixed grid, launched as a programmatic dependent of the fast pass,
// walks the list of tilted environments.  Almost always the list is empty and every CTA leaves after one load.
template <typename T, int LANES, int NC>
__global__ void __launch_bounds__(Geo<T, LANES, NC>::kBlock, Geo<T, LANES, NC>::kMinBlocks) step_kernel_list(DevState<T> st, StepArgs a, const __grid_constant__ Consts<T> C) {
#if defined(__CUDA_ARCH__)
  asm volatile("griddepcontrol.wait;" ::: "memory");     // nothing is read before the fast pass's writes are visible
#endif
  // list length; the last CTA to have read it zeroes it for the next step (ticket in gen_count[1]: no step parity baked into the
  // launch, so a captured step can be replayed)
  __shared__ int s_count;
  if (threadIdx.x == 0) {
    const int c = *reinterpret_cast<volatile int*>(a.gen_count);
    __threadfence();
    if (atomicAdd(a.gen_count + 1, 1) == (int)gridDim.x - 1) { a.gen_count[0] = 0; a.gen_count[1] = 0; }
    s_count = c;
  }
  __syncthreads();
  int count = s_count;
  count = count > st.n ? st.n : count;
  const int per_block = Geo<T, LANES, NC>::kBlock / LANES;
  if ((long long)blockIdx.x * per_block >= count) return;
  stage_tables(C);
  for (int blk = blockIdx.x; (long long)blk * per_block < count; blk += gridDim.x) {
    step_body<T, LANES, NC, true>(st, a, C, blk, count);
    __syncwarp();
  }
}

template <typename T, int LANES, int NC>
__global__ void __launch_bounds__(Geo<T, LANES, 2>::kBlock) reset_kernel(DevState<T> st, StepArgs a, const __grid_constant__ Consts<T> C) {
  using E = EnvOps<T, LANES, NC>;
  using G = Geo<T, LANES, 2>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  stage_tables(C);
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  const int env_raw = tid / LANES, lane = tid % LANES;
  const bool valid = env_raw < st.n;
  const int env = valid ? env_raw : st.n - 1;
  const bool sel = (a.mask ? (a.mask[env] != 0) : true) && valid;   // team-uniform
  // masked launch (deferred auto-reset of the maze models after every step): a warp without a selected environment leaves at
  // once -- the settle steps below are three full physics substeps.  All collectives are intra-warp and no CTA barrier follows.
  if (!__any_sync(0xffffffffu, sel)) return;
  // the reset kernel keeps wheel records in local storage and uses shared memory for the observation rows only
  Wheel<T, NC> wh[G::WPL];
  float* row = reinterpret_cast<float*>(smem_raw) + (size_t)(threadIdx.x / LANES) * a.obs_dim;
  RowSink<false> sink{row};
  typename E::State e;
  Episode<T> ep;
  ep.episode = st.episode[env];
  E::reset_env(C, e, wh, ep, lane, a.seed, (uint32_t)env + a.env_base);
  T dist, minl;
  if (!E::settle_and_observe(C, e, wh, ep, lane, sink, &dist, &minl)) {   // maze scenes settle first; everything else observes the spawn state
    Kin<T> k;
    E::S::kinematics(e, k);
    E::observe(C, e, k, ep, lane, sink, &dist, &minl);
  }
  __syncwarp();
  if (!sel) return;
  for (int j = lane; j < a.obs_dim; j += LANES) a.obs[(size_t)env * a.obs_pitch + j] = sink.row[j];
  SoAAcc<T> acc{st, env};
  E::store_state(acc, lane, e, wh);
  if (lane == 0) {
    st.goal[env] = ep.goal[0]; st.goal[(size_t)st.n + env] = ep.goal[1];
    st.ref[env] = ep.ref[0]; st.ref[(size_t)st.n + env] = ep.ref[1];
    st.step_count[env] = 0;
    st.episode[env] = ep.episode;
    st.ep_return[env] = 0.f;
  }
}

__global__ void random_action_kernel(float* action, int n, unsigned long long seed, uint32_t step, uint32_t env_base) {
  int env = blockIdx.x * blockDim.x + threadIdx.x;
  if (env >= n) return;
  float a0, a1;
  synth_action(seed, step, (uint32_t)env + env_base, &a0, &a1);
  reinterpret_cast<float2*>(action)[env] = make_float2(a0, a1);
}

thread_local std::string g_last_error;

// NVTX range around one C-ABI call (tracing): ackb_step / ackb_reset / ackb_step_host show up as named
// ranges on the calling thread's timeline, with the kernels they launch underneath
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
};

}  // namespace

// ------------------------------------------------------------------------------------------------
// handle
// ------------------------------------------------------------------------------------------------
struct ackb_handle {
  int n = 0, device = 0, dtype = ACKB_F32, lanes = 4, obs_dim = 0;
  int obs_pitch = 0;            // row pitch (floats) of the caller's observation arrays (ackb_set_obs_pitch; default obs_dim)
  int num_sms = 148;
  int cta_sync = -1;            // -1 = auto (by grid size), 0 / 1 forced through ACKB_CTA_SYNC (tuning)
  int zero_copy = 1;            // ackb_step_host: let the kernel access pinned caller buffers directly (ACKB_ZERO_COPY=0 disables)
  unsigned long long seed = 0;
  uint32_t env_base = 0;        // global id of local environment 0 (ackb_set_env_id_base)
  uint32_t step_index = 0;
  unsigned long long stat_steps = 0;
  unsigned long long launches = 0;
  double* consts_host = nullptr;
  Consts<float>* consts_f = nullptr;    // the handle's own copy in the kernels' precision (passed by value at every launch)
  void* state = nullptr;        // one allocation holding every SoA array
  size_t elem = 4;
  DevState<float> sf{};
  DevState<double> sd{};
  ackb_stats_t* stats = nullptr;
  // staging for ackb_step_host
  float *d_action = nullptr, *d_obs = nullptr, *d_reward = nullptr;
  uint8_t *d_term = nullptr, *d_trunc = nullptr;
  uint8_t* d_done = nullptr;    // per-env done mask of the deferred reset (models with settle steps)
  int* d_gen_list = nullptr;    // flat-floor regime split: environments the fast pass left to the general contact pass ...
  int* d_gen_count = nullptr;   // ... [0] their number, [1] the general pass's read ticket (both back to zero after every step)
  int general_pass = 1;         // flat-floor model: run the general-contact pass after the fast kernel (ACKB_GENERAL_PASS=0 disables)
  cudaStream_t own_stream = nullptr;
  // cross-stream ordering: ackb_step_host runs on own_stream, so it has to wait for work the caller issued through
  // ackb_step / ackb_reset on another stream (last_stream) before it touches the state arrays
  cudaStream_t last_stream = nullptr;
  bool last_is_foreign = false;
  cudaEvent_t order_ev = nullptr;
  std::string err;
};

namespace {
int fail(ackb_handle* h, int code, const std::string& msg) {
  g_last_error = msg;
  if (h) h->err = msg;
  return code;
}
#define CK(call)                                                                                          \
  do {                                                                                                    \
    cudaError_t e_ = (call);                                                                              \
    if (e_ != cudaSuccess) return fail(h, ACKB_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
  } while (0)

template <typename T>
void carve(ackb_handle* h, DevState<T>& s) {
  T* p = static_cast<T*>(h->state);
  const size_t n = h->n;
  s.n = h->n;
  s.qpos = p; p += 13 * n;
  s.qvel = p; p += 12 * n;
  s.warm = p; p += 12 * n;
  s.goal = p; p += 2 * n;
  s.ref = p; p += 2 * n;
  s.step_count = reinterpret_cast<int32_t*>(p);
  s.episode = reinterpret_cast<uint32_t*>(s.step_count + n);
  s.ep_return = reinterpret_cast<float*>(s.episode + n);
}

template <typename T> const Consts<T>& handle_consts(const ackb_handle* h);
template <> const Consts<float>& handle_consts<float>(const ackb_handle* h) { return *h->consts_f; }
template <> const Consts<double>& handle_consts<double>(const ackb_handle* h) { return *reinterpret_cast<const Consts<double>*>(h->consts_host); }

// cudaFuncAttributeMaxDynamicSharedMemorySize is per (kernel, device) state: remember the largest value set so far
template <typename K>
int ensure_smem_attr(ackb_handle* h, K kernel, std::atomic<int>* set_per_device, size_t smem) {
  if (smem <= 48 * 1024 || (int)smem <= set_per_device[h->device].load(std::memory_order_acquire)) return ACKB_OK;
  CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  set_per_device[h->device].store((int)smem, std::memory_order_release);
  return ACKB_OK;
}

template <typename T, int LANES, int NC>
int launch_one(ackb_handle* h, DevState<T>& st, const StepArgs& a, cudaStream_t stream, bool is_reset) {
  using G = Geo<T, LANES, NC>;
  const long long threads = (long long)h->n * LANES;
  const int grid = (int)((threads + G::kBlock - 1) / G::kBlock);
  if (is_reset) {
    using GR = Geo<T, LANES, 2>;   // the reset kernel has its own geometry (records in local storage)
    const int rgrid = (int)((threads + GR::kBlock - 1) / GR::kBlock);
    const size_t smem = (size_t)(GR::kBlock / LANES) * a.obs_dim * sizeof(float);
    static std::atomic<int> rattr[64];
    if (int rc = ensure_smem_attr(h, reset_kernel<T, LANES, NC>, rattr, smem)) return rc;
    reset_kernel<T, LANES, NC><<<rgrid, GR::kBlock, smem, stream>>>(st, a, handle_consts<T>(h));
  } else {
    const size_t smem = G::smem_bytes(a.obs_dim);
    StepArgs a2 = a;
    a2.cta_sync = h->cta_sync >= 0 ? h->cta_sync : 1;   // measured on B200: faster at every batch size from 4096 to 131072 envs
    static std::atomic<int> attr[64];
    if constexpr (LANES == 4 && NC == 4) if (a.regime == 2) {
      // general pass: a grid of at most one CTA per SM walks the list of tilted environments; programmatic dependent launch:
      // the grid is scheduled while the fast pass drains (its blocks sit in griddepcontrol.wait)
      a2.cta_sync = 0;
      static std::atomic<int> lattr[64];
      if (int rc = ensure_smem_attr(h, step_kernel_list<T, LANES, NC>, lattr, smem)) return rc;
      const int lgrid = grid < h->num_sms ? grid : h->num_sms;
      cudaLaunchConfig_t cfg{};
      cfg.gridDim = dim3((unsigned)lgrid); cfg.blockDim = dim3((unsigned)G::kBlock); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
      cudaLaunchAttribute at{};
      at.id = cudaLaunchAttributeProgrammaticStreamSerialization;
      at.val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = &at; cfg.numAttrs = 1;
      CK(cudaLaunchKernelEx(&cfg, step_kernel_list<T, LANES, NC>, st, a2, handle_consts<T>(h)));
      h->launches++;
      CK(cudaGetLastError());
      return ACKB_OK;
    }
    if (int rc = ensure_smem_attr(h, step_kernel<T, LANES, NC>, attr, smem)) return rc;
    step_kernel<T, LANES, NC><<<grid, G::kBlock, smem, stream>>>(st, a2, handle_consts<T>(h));
  }
  h->launches++;
  CK(cudaGetLastError());
  return ACKB_OK;
}

template <typename T>
int launch_step(ackb_handle* h, DevState<T>& st, const StepArgs& a, cudaStream_t stream, bool is_reset) {
  h->last_stream = stream; h->last_is_foreign = stream != h->own_stream;
  const bool scene = h->consts_host[0] != 0.0;   // model_kind: the obstacle scene needs the two box-contact slots per wheel
#ifdef ACKB_TUNE_MIN   // tuning builds (tools/gpu/build_variant.sh): fp32 flat-floor kernels only, to keep compile times short
  if constexpr (sizeof(T) == 8) return fail(h, ACKB_ERR_ARG, "tuning build: fp32 only");
  else return h->lanes >= 4 ? launch_one<T, 4, 2>(h, st, a, stream, is_reset) : launch_one<T, 1, 2>(h, st, a, stream, is_reset);
#else
  if (!scene && !is_reset && h->general_pass) {
    // flat-floor model: fast kernel (two floor-contact slots per wheel) for the upright environments, then the general kernel
    // (extra slots: cap-down wheel points, chassis plates) for the environments the fast pass marked as tilted
    StepArgs f = a;
    f.regime = 1; f.gen_list = h->d_gen_list; f.gen_count = h->d_gen_count;
    int rc = h->lanes == 8 ? launch_one<T, 8, 1>(h, st, f, stream, false)
                           : (h->lanes >= 4 ? launch_one<T, 4, 2>(h, st, f, stream, false) : launch_one<T, 1, 2>(h, st, f, stream, false));
    if (rc) return rc;
    StepArgs g = f;
    g.regime = 2;
    return launch_one<T, 4, 4>(h, st, g, stream, false);
  }
  if (h->lanes == 8 && !scene) return launch_one<T, 8, 1>(h, st, a, stream, is_reset);   // one lane per floor contact (flat-floor model)
  if (h->lanes >= 4) return scene ? launch_one<T, 4, 4>(h, st, a, stream, is_reset) : launch_one<T, 4, 2>(h, st, a, stream, is_reset);
  return scene ? launch_one<T, 1, 4>(h, st, a, stream, is_reset) : launch_one<T, 1, 2>(h, st, a, stream, is_reset);
#endif
}
}  // namespace

extern "C" {

int ackb_consts_len(void) { return kNumConsts; }

const char* ackb_last_error(const ackb_handle* h) { return h ? h->err.c_str() : g_last_error.c_str(); }

int ackb_create(const double* consts, size_t consts_len, int num_envs, int device, int dtype, uint64_t seed, int lanes_per_env,
                ackb_handle** out) {
  ackb_handle* h = nullptr;
  if (!consts || !out || num_envs <= 0) return fail(nullptr, ACKB_ERR_ARG, "ackb_create: null pointer or num_envs <= 0");
  if ((int)consts_len != kNumConsts) return fail(nullptr, ACKB_ERR_ARG, "ackb_create: constants blob has the wrong length");
  if (dtype != ACKB_F32 && dtype != ACKB_F64) return fail(nullptr, ACKB_ERR_ARG, "ackb_create: dtype must be ACKB_F32 or ACKB_F64");
  if (lanes_per_env == 0) {   // throughput layout (1 lane) for big flat-floor batches, latency layout (4 lanes) otherwise and for the obstacle scene
    const bool scene_model = consts[0] != 0.0;
    lanes_per_env = (num_envs >= 32768 && !scene_model) ? 1 : 4;
  }
  if (lanes_per_env != 1 && lanes_per_env != 4 && lanes_per_env != 8) return fail(nullptr, ACKB_ERR_ARG, "ackb_create: lanes_per_env must be 1, 4 or 8");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(nullptr, ACKB_ERR_NO_DEVICE, "ackb_create: no CUDA device (there is no CPU fallback)");
  if (device < 0 || device >= ndev || device >= 64) return fail(nullptr, ACKB_ERR_ARG, "ackb_create: bad device index");
  h = new (std::nothrow) ackb_handle();
  if (!h) return fail(nullptr, ACKB_ERR_ARG, "ackb_create: out of host memory");
  h->n = num_envs; h->device = device; h->dtype = dtype; h->seed = seed; h->lanes = lanes_per_env;
  h->elem = dtype == ACKB_F32 ? 4 : 8;
  h->consts_host = new double[kNumConsts];
  memcpy(h->consts_host, consts, sizeof(double) * kNumConsts);
  h->consts_f = new Consts<float>;
  for (int i = 0; i < kNumConsts; ++i) reinterpret_cast<float*>(h->consts_f)[i] = (float)consts[i];
  h->obs_dim = (int)reinterpret_cast<const Consts<double>*>(consts)->nbeam[0] + 7;
  h->obs_pitch = h->obs_dim;
  CK(cudaSetDevice(device));
  CK(cudaDeviceGetAttribute(&h->num_sms, cudaDevAttrMultiProcessorCount, device));
  if (const char* ev = getenv("ACKB_CTA_SYNC")) h->cta_sync = atoi(ev);
  if (const char* ev = getenv("ACKB_ZERO_COPY")) h->zero_copy = atoi(ev);
  const size_t n = num_envs;
  const size_t bytes = (13 + 12 + 12 + 2 + 2) * n * h->elem + 3 * n * 4;
  CK(cudaMalloc(&h->state, bytes));
  CK(cudaMemset(h->state, 0, bytes));
  if (dtype == ACKB_F32) carve(h, h->sf); else carve(h, h->sd);
  CK(cudaMalloc(&h->stats, sizeof(ackb_stats_t)));
  CK(cudaMemset(h->stats, 0, sizeof(ackb_stats_t)));
  CK(cudaMalloc(&h->d_action, n * 2 * sizeof(float)));
  CK(cudaMalloc(&h->d_obs, n * h->obs_dim * sizeof(float)));
  CK(cudaMalloc(&h->d_reward, n * sizeof(float)));
  CK(cudaMalloc(&h->d_term, n));
  CK(cudaMalloc(&h->d_trunc, n));
  CK(cudaMalloc(&h->d_done, n));
  CK(cudaMemset(h->d_done, 0, n));
  CK(cudaMalloc(&h->d_gen_list, n * sizeof(int)));
  CK(cudaMemset(h->d_gen_list, 0, n * sizeof(int)));
  CK(cudaMalloc(&h->d_gen_count, 2 * sizeof(int)));
  CK(cudaMemset(h->d_gen_count, 0, 2 * sizeof(int)));
  if (const char* ev = getenv("ACKB_GENERAL_PASS")) h->general_pass = atoi(ev);

  CK(cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking));
  CK(cudaEventCreateWithFlags(&h->order_ev, cudaEventDisableTiming));
  *out = h;
  return ACKB_OK;
}

int ackb_destroy(ackb_handle* h) {
  if (!h) return ACKB_ERR_ARG;
  cudaSetDevice(h->device);
  cudaFree(h->state); cudaFree(h->stats); cudaFree(h->d_action); cudaFree(h->d_obs); cudaFree(h->d_reward);
  cudaFree(h->d_term); cudaFree(h->d_trunc); cudaFree(h->d_done); cudaFree(h->d_gen_list); cudaFree(h->d_gen_count);
  if (h->order_ev) cudaEventDestroy(h->order_ev);
  if (h->own_stream) cudaStreamDestroy(h->own_stream);
  delete[] h->consts_host;
  delete h->consts_f;
  delete h;
  return ACKB_OK;
}

int ackb_set_env_id_base(ackb_handle* h, uint64_t env_id_base) {
  if (!h) return ACKB_ERR_ARG;
  if (env_id_base + (uint64_t)h->n > 0xffffffffull) return fail(h, ACKB_ERR_ARG, "ackb_set_env_id_base: global environment ids must fit 32 bits");
  h->env_base = (uint32_t)env_id_base;
  return ACKB_OK;
}

int ackb_set_obs_pitch(ackb_handle* h, int pitch_floats) {
  if (!h) return ACKB_ERR_ARG;
  if (pitch_floats < h->obs_dim) return fail(h, ACKB_ERR_ARG, "ackb_set_obs_pitch: pitch must be >= obs_dim");
  h->obs_pitch = pitch_floats;
  return ACKB_OK;
}

int ackb_num_envs(const ackb_handle* h) { return h ? h->n : ACKB_ERR_ARG; }
int ackb_obs_dim(const ackb_handle* h) { return h ? h->obs_dim : ACKB_ERR_ARG; }
int ackb_dtype(const ackb_handle* h) { return h ? h->dtype : ACKB_ERR_ARG; }
unsigned long long ackb_launch_count(const ackb_handle* h) { return h ? h->launches : 0ull; }

int ackb_reset(ackb_handle* h, const uint8_t* dev_mask, float* dev_obs, void* stream) {
  NvtxRange nvtx("ackb_reset");
  if (!h || !dev_obs) return fail(h, ACKB_ERR_ARG, "ackb_reset: null pointer");
  CK(cudaSetDevice(h->device));
  StepArgs a{};
  a.obs = dev_obs; a.mask = dev_mask; a.seed = h->seed; a.obs_dim = h->obs_dim; a.obs_pitch = h->obs_pitch; a.stats = h->stats; a.env_base = h->env_base;
  return h->dtype == ACKB_F32 ? launch_step(h, h->sf, a, (cudaStream_t)stream, true) : launch_step(h, h->sd, a, (cudaStream_t)stream, true);
}

int ackb_step(ackb_handle* h, const float* dev_action, int frame_skip, int auto_reset, float* dev_obs, float* dev_reward,
              uint8_t* dev_terminated, uint8_t* dev_truncated, float* dev_terminal_obs, int32_t* dev_ncon, void* stream) {
  NvtxRange nvtx("ackb_step");
  if (!h || !dev_obs || !dev_reward || !dev_terminated || !dev_truncated) return fail(h, ACKB_ERR_ARG, "ackb_step: null output pointer");
  if (frame_skip < 1) return fail(h, ACKB_ERR_ARG, "ackb_step: frame_skip must be >= 1");
  CK(cudaSetDevice(h->device));
  int rc = ACKB_OK;
  StepArgs a{};
  a.env_base = h->env_base;
  a.action = dev_action; a.obs = dev_obs; a.reward = dev_reward; a.terminated = dev_terminated; a.truncated = dev_truncated;
  a.terminal_obs = dev_terminal_obs; a.ncon = dev_ncon; a.stats = h->stats; a.seed = h->seed; a.step_index = h->step_index++;
  h->stat_steps += (unsigned long long)h->n;
  a.frame_skip = frame_skip; a.auto_reset = auto_reset; a.obs_dim = h->obs_dim; a.obs_pitch = h->obs_pitch;
  // models whose reset includes settle steps (maze scenes): the step kernel only marks the finished environments, a masked
  // reset launch on the same stream then resets, settles and observes them
  const bool defer = auto_reset && reinterpret_cast<const Consts<double>*>(h->consts_host)->settle_steps[0] > 0.0;
  a.defer_reset = defer ? 1 : 0;
  a.done_mask = h->d_done;
  rc = h->dtype == ACKB_F32 ? launch_step(h, h->sf, a, (cudaStream_t)stream, false) : launch_step(h, h->sd, a, (cudaStream_t)stream, false);
  if (rc || !defer) return rc;
  StepArgs r{};
  r.obs = dev_obs; r.mask = h->d_done; r.seed = h->seed; r.obs_dim = h->obs_dim; r.obs_pitch = h->obs_pitch; r.stats = h->stats; r.env_base = h->env_base;
  return h->dtype == ACKB_F32 ? launch_step(h, h->sf, r, (cudaStream_t)stream, true) : launch_step(h, h->sd, r, (cudaStream_t)stream, true);
}

// Device-visible alias of a host pointer: pinned (cudaHostAlloc / cudaHostRegister) memory is mapped into the device address
// space under unified addressing, so the kernel can read the actions from it and write its results into it directly over
// PCIe/NVLink-C2C, overlapped with the computation (no staging copies).  Returns null for pageable memory.
static void* device_alias(const void* host_ptr) {
  cudaPointerAttributes at{};
  if (cudaPointerGetAttributes(&at, host_ptr) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  if (at.type == cudaMemoryTypeHost && at.devicePointer) return at.devicePointer;
  return nullptr;
}

int ackb_step_host(ackb_handle* h, const float* host_action, int frame_skip, int auto_reset, float* host_obs, float* host_reward,
                   uint8_t* host_terminated, uint8_t* host_truncated) {
  NvtxRange nvtx("ackb_step_host");
  if (!h || !host_action || !host_obs || !host_reward || !host_terminated || !host_truncated) return fail(h, ACKB_ERR_ARG, "ackb_step_host: null pointer");
  CK(cudaSetDevice(h->device));
  const size_t n = h->n;
  cudaStream_t s = h->own_stream;
  if (h->last_is_foreign) {   // order after the caller's earlier ackb_step / ackb_reset on its own stream
    if (cudaEventRecord(h->order_ev, h->last_stream) == cudaSuccess) CK(cudaStreamWaitEvent(s, h->order_ev, 0));
    else { cudaGetLastError(); CK(cudaDeviceSynchronize()); }   // that stream no longer exists
  }
  // pinned caller buffers: zero-copy (the kernel's coalesced observation stores go straight to host memory)
  float* z_act = (float*)device_alias(host_action);
  float* z_obs = (float*)device_alias(host_obs);
  float* z_rew = (float*)device_alias(host_reward);
  uint8_t* z_term = (uint8_t*)device_alias(host_terminated);
  uint8_t* z_trunc = (uint8_t*)device_alias(host_truncated);
  const bool zero_copy = h->zero_copy && z_act && z_obs && z_rew && z_term && z_trunc;
  if (zero_copy) {
    const int pitch_keep = h->obs_pitch;
    h->obs_pitch = h->obs_dim;            // host observation rows are dense
    int rc = ackb_step(h, z_act, frame_skip, auto_reset, z_obs, z_rew, z_term, z_trunc, nullptr, nullptr, s);
    h->obs_pitch = pitch_keep;
    if (rc) return rc;
    CK(cudaStreamSynchronize(s));
    return ACKB_OK;
  }
  CK(cudaMemcpyAsync(h->d_action, host_action, n * 2 * sizeof(float), cudaMemcpyHostToDevice, s));
  const int pitch_keep = h->obs_pitch;
  h->obs_pitch = h->obs_dim;              // host observation rows are dense
  int rc = ackb_step(h, h->d_action, frame_skip, auto_reset, h->d_obs, h->d_reward, h->d_term, h->d_trunc, nullptr, nullptr, s);
  h->obs_pitch = pitch_keep;
  if (rc) return rc;
  CK(cudaMemcpyAsync(host_obs, h->d_obs, n * h->obs_dim * sizeof(float), cudaMemcpyDeviceToHost, s));
  CK(cudaMemcpyAsync(host_reward, h->d_reward, n * sizeof(float), cudaMemcpyDeviceToHost, s));
  CK(cudaMemcpyAsync(host_terminated, h->d_term, n, cudaMemcpyDeviceToHost, s));
  CK(cudaMemcpyAsync(host_truncated, h->d_trunc, n, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  return ACKB_OK;
}

int ackb_random_actions(ackb_handle* h, float* dev_action, void* stream) {
  if (!h || !dev_action) return fail(h, ACKB_ERR_ARG, "ackb_random_actions: null pointer");
  CK(cudaSetDevice(h->device));
  random_action_kernel<<<(h->n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(dev_action, h->n, h->seed, h->step_index, h->env_base);
  h->launches++;
  CK(cudaGetLastError());
  return ACKB_OK;
}

}  // extern "C"

// ---- state access (tests): row-major host doubles <-> SoA device arrays -----------------------------
namespace {
template <typename T>
int xfer(ackb_handle* h, T* dev, double* host, int rows, bool to_host) {
  const size_t n = h->n;
  CK(cudaDeviceSynchronize());   // the caller's streams may be non-blocking ones that a plain cudaMemcpy does not wait for
  T* tmp = new T[rows * n];
  if (to_host) {
    CK(cudaMemcpy(tmp, dev, rows * n * sizeof(T), cudaMemcpyDeviceToHost));
    for (size_t e = 0; e < n; ++e) for (int r = 0; r < rows; ++r) host[e * rows + r] = (double)tmp[r * n + e];
  } else {
    for (size_t e = 0; e < n; ++e) for (int r = 0; r < rows; ++r) tmp[r * n + e] = (T)host[e * rows + r];
    CK(cudaMemcpy(dev, tmp, rows * n * sizeof(T), cudaMemcpyHostToDevice));
  }
  delete[] tmp;
  return ACKB_OK;
}
template <typename T>
int state_io(ackb_handle* h, DevState<T>& s, double* qpos, double* qvel, double* warm, bool to_host) {
  int rc = 0;
  if (qpos && (rc = xfer(h, s.qpos, qpos, 13, to_host))) return rc;
  if (qvel && (rc = xfer(h, s.qvel, qvel, 12, to_host))) return rc;
  if (warm && (rc = xfer(h, s.warm, warm, 12, to_host))) return rc;
  return ACKB_OK;
}
template <typename T>
int episode_io(ackb_handle* h, DevState<T>& s, double* goal, double* ref, int32_t* sc, bool to_host) {
  int rc = 0;
  if (goal && (rc = xfer(h, s.goal, goal, 2, to_host))) return rc;
  if (ref && (rc = xfer(h, s.ref, ref, 2, to_host))) return rc;
  if (sc) {
    CK(cudaDeviceSynchronize());
    if (to_host) CK(cudaMemcpy(sc, s.step_count, h->n * sizeof(int32_t), cudaMemcpyDeviceToHost));
    else CK(cudaMemcpy(s.step_count, sc, h->n * sizeof(int32_t), cudaMemcpyHostToDevice));
  }
  return ACKB_OK;
}
}  // namespace

extern "C" {

int ackb_get_state(ackb_handle* h, double* q, double* v, double* w) {
  if (!h) return ACKB_ERR_ARG;
  CK(cudaSetDevice(h->device));
  CK(cudaDeviceSynchronize());
  return h->dtype == ACKB_F32 ? state_io(h, h->sf, q, v, w, true) : state_io(h, h->sd, q, v, w, true);
}
int ackb_set_state(ackb_handle* h, const double* q, const double* v, const double* w) {
  if (!h) return ACKB_ERR_ARG;
  CK(cudaSetDevice(h->device));
  CK(cudaDeviceSynchronize());
  return h->dtype == ACKB_F32 ? state_io(h, h->sf, (double*)q, (double*)v, (double*)w, false)
                              : state_io(h, h->sd, (double*)q, (double*)v, (double*)w, false);
}
int ackb_get_episode(ackb_handle* h, double* goal, double* ref, int32_t* sc) {
  if (!h) return ACKB_ERR_ARG;
  CK(cudaSetDevice(h->device));
  CK(cudaDeviceSynchronize());
  return h->dtype == ACKB_F32 ? episode_io(h, h->sf, goal, ref, sc, true) : episode_io(h, h->sd, goal, ref, sc, true);
}
int ackb_set_episode(ackb_handle* h, const double* goal, const double* ref, const int32_t* sc) {
  if (!h) return ACKB_ERR_ARG;
  CK(cudaSetDevice(h->device));
  CK(cudaDeviceSynchronize());
  return h->dtype == ACKB_F32 ? episode_io(h, h->sf, (double*)goal, (double*)ref, (int32_t*)sc, false)
                              : episode_io(h, h->sd, (double*)goal, (double*)ref, (int32_t*)sc, false);
}

int ackb_stats(ackb_handle* h, ackb_stats_t* out) {
  if (!h || !out) return ACKB_ERR_ARG;
  CK(cudaSetDevice(h->device));
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(out, h->stats, sizeof(ackb_stats_t), cudaMemcpyDeviceToHost));
  out->env_steps = h->stat_steps;
  return ACKB_OK;
}
int ackb_stats_reset(ackb_handle* h) {
  if (!h) return ACKB_ERR_ARG;
  CK(cudaSetDevice(h->device));
  CK(cudaDeviceSynchronize());
  CK(cudaMemset(h->stats, 0, sizeof(ackb_stats_t)));
  h->stat_steps = 0;
  return ACKB_OK;
}

}  // extern "C"
