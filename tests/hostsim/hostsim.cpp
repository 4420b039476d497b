// hostsim.cpp -- TEST-ONLY host build of the per-environment device code (LANES = 1).
// Lets the arithmetic of mujoco_playground_b200/csrc/ackb_core.cuh / ackb_env.cuh be checked against
// the oracle on a machine without a GPU.  It is not part of the product and is never loaded by it.
#include <cstring>
#include <thread>
#include <vector>
#include "../../mujoco_playground_b200/csrc/ackb_env.cuh"

using namespace ackb;

static int g_general = 0;
namespace {
template <typename T>
struct ArrAcc {
  double *qp, *qv, *wm;
  T qpos(int i) const { return (T)qp[i]; }
  T qvel(int i) const { return (T)qv[i]; }
  T warm(int i) const { return (T)wm[i]; }
  void set_qpos(int i, T v) { qp[i] = (double)v; }
  void set_qvel(int i, T v) { qv[i] = (double)v; }
  void set_warm(int i, T v) { wm[i] = (double)v; }
};
struct ObsSink {
  static constexpr bool kAliasesWheels = false;
  float* o;
  void put(int slot, float v) { o[slot] = v; }
};
template <typename T>
void to_consts(const double* blob, Consts<T>& C) {
  T* dst = reinterpret_cast<T*>(&C);
  for (int i = 0; i < kNumConsts; ++i) dst[i] = (T)blob[i];
}

template <typename T, int NC>
void substep(const double* blob, double* qpos, double* qvel, double* warm, const double* ctrl_in, int nsteps, double* tap_out, int* diag_out) {
  using E = EnvOps<T, 1, NC>;
  Consts<T> C;
  to_consts(blob, C);
  typename E::State e;
  Wheel<T, NC> wh[4];
  ArrAcc<T> acc{qpos, qvel, warm};
  E::load_state(acc, 0, e, wh);
  T ctrl[4];
  for (int i = 0; i < 4; ++i) ctrl[i] = (T)ctrl_in[i];
  StepDiag diag{};
  DebugTap<T> tap;
  for (int s = 0; s < nsteps; ++s) {
    Kin<T> k;
    E::S::kinematics(e, k);
    diag.ncon = 0;
    E::S::dynamics(C, e, k, ctrl, 0, wh, diag, &tap);
  }
  E::store_state(acc, 0, e, wh);
  if (tap_out) {
    for (int i = 0; i < 12; ++i) { tap_out[i] = tap.tau[i]; tap_out[12 + i] = tap.a_smooth[i]; tap_out[24 + i] = tap.a[i]; tap_out[36 + i] = tap.fc[i]; }
    tap_out[48] = tap.niter; tap_out[49] = tap.nls;
  }
  if (diag_out) { diag_out[0] = diag.ncon; diag_out[1] = diag.unsupported; diag_out[2] = diag.niter; diag_out[3] = diag.bad; }
}

template <typename T, int NC>
void env_step(const double* blob, double* qpos, double* qvel, double* warm, double* epd /*goal2 ref2*/, int* epi /*step_count episode*/,
              const float* action, int frame_skip, float* obs, float* out, int* diag_out) {
  using E = EnvOps<T, 1, NC>;
  Consts<T> C;
  to_consts(blob, C);
  typename E::State e;
  Wheel<T, NC> wh[4];
  ArrAcc<T> acc{qpos, qvel, warm};
  E::load_state(acc, 0, e, wh);
  Episode<T> ep;
  ep.goal[0] = (T)epd[0]; ep.goal[1] = (T)epd[1]; ep.ref[0] = (T)epd[2]; ep.ref[1] = (T)epd[3];
  ep.step_count = epi[0]; ep.episode = (uint32_t)epi[1];
  ObsSink sink{obs};
  StepOut<T> so;
  StepDiag diag{};
  E::step_env(C, e, wh, ep, action[0], action[1], frame_skip, 0, sink, [] {}, so, diag, (DebugTap<T>*)nullptr);
  E::store_state(acc, 0, e, wh);
  epi[0] = ep.step_count;
  out[0] = so.reward; out[1] = so.terminated; out[2] = so.truncated; out[3] = so.collision; out[4] = so.goal_distance; out[5] = so.min_lidar;
  if (diag_out) { diag_out[0] = diag.ncon; diag_out[1] = diag.unsupported; diag_out[2] = diag.niter; diag_out[3] = diag.bad; }
}

template <typename T, int NC>
void env_reset(const double* blob, double* qpos, double* qvel, double* warm, double* epd, int* epi, unsigned long long seed, unsigned env_id, float* obs) {
  using E = EnvOps<T, 1, NC>;
  Consts<T> C;
  to_consts(blob, C);
  typename E::State e;
  Wheel<T, NC> wh[4];
  Episode<T> ep;
  ep.episode = (uint32_t)epi[1];
  E::reset_env(C, e, wh, ep, 0, seed, env_id);
  ObsSink sink{obs};
  T dist, minl;
  if (!E::settle_and_observe(C, e, wh, ep, 0, sink, &dist, &minl)) {
    Kin<T> k;
    E::S::kinematics(e, k);
    E::observe(C, e, k, ep, 0, sink, &dist, &minl);
  }
  ArrAcc<T> acc{qpos, qvel, warm};
  E::store_state(acc, 0, e, wh);
  epd[0] = (double)ep.goal[0]; epd[1] = (double)ep.goal[1]; epd[2] = (double)ep.ref[0]; epd[3] = (double)ep.ref[1];
  epi[0] = ep.step_count; epi[1] = (int)ep.episode;
}
// LANES host threads emulate the lanes of one environment (collectives meet at a std::barrier): exercises the
// multi-lane code paths (4 lanes = one per wheel, 8 lanes = one per floor contact) without a GPU.
template <typename T, int LANES, int NC>
void env_step_team(const double* blob, double* qpos, double* qvel, double* warm, double* epd, int* epi, const float* action, int frame_skip,
                   float* obs, float* out, int* diag_out) {
  using E = EnvOps<T, LANES, NC>;
  Consts<T> C;
  to_consts(blob, C);
  std::barrier<> bar(LANES);
  HostTeamCtx ctx{LANES, &bar, {0}};
  std::vector<std::thread> th;
  int ncon_lane[LANES], unsup_lane[LANES], niter_lane[LANES];
  for (int lane = 0; lane < LANES; ++lane)
    th.emplace_back([&, lane] {
      g_host_team = &ctx; g_host_lane = lane;
      typename E::State e;
      Wheel<T, NC> wh[E::WPL];
      ArrAcc<T> acc{qpos, qvel, warm};
      E::load_state(acc, lane, e, wh);
      Episode<T> ep;
      ep.goal[0] = (T)epd[0]; ep.goal[1] = (T)epd[1]; ep.ref[0] = (T)epd[2]; ep.ref[1] = (T)epd[3];
      ep.step_count = epi[0]; ep.episode = (uint32_t)epi[1];
      ObsSink sink{obs};
      StepOut<T> so;
      StepDiag diag{};
      E::step_env(C, e, wh, ep, action[0], action[1], frame_skip, lane, sink, [] {}, so, diag, (DebugTap<T>*)nullptr);
      bar.arrive_and_wait();           // every lane finished reading the shared state arrays
      E::store_state(acc, lane, e, wh);
      ncon_lane[lane] = diag.ncon; unsup_lane[lane] = diag.unsupported; niter_lane[lane] = diag.niter;
      if (lane == 0) {
        epi[0] = ep.step_count;
        out[0] = so.reward; out[1] = so.terminated; out[2] = so.truncated; out[3] = so.collision; out[4] = so.goal_distance; out[5] = so.min_lidar;
      }
    });
  for (auto& t : th) t.join();
  if (diag_out) {
    diag_out[0] = diag_out[1] = 0; diag_out[2] = niter_lane[0];
    for (int l = 0; l < LANES; ++l) { diag_out[0] += ncon_lane[l]; diag_out[1] |= unsup_lane[l]; }
  }
}
}  // namespace

extern "C" {
// multi-lane emulation (v2 model only): lanes = 4 or 8
void hs_env_step_lanes(int lanes, int f32, const double* blob, double* qpos, double* qvel, double* warm, double* epd, int* epi, const float* action,
                       int frame_skip, float* obs, float* out, int* diag) {
  if (lanes == 4 && (blob[0] != 0.0 || g_general)) { if (f32) env_step_team<float, 4, 4>(blob, qpos, qvel, warm, epd, epi, action, frame_skip, obs, out, diag); else env_step_team<double, 4, 4>(blob, qpos, qvel, warm, epd, epi, action, frame_skip, obs, out, diag); }
  else if (lanes == 4) { if (f32) env_step_team<float, 4, 2>(blob, qpos, qvel, warm, epd, epi, action, frame_skip, obs, out, diag); else env_step_team<double, 4, 2>(blob, qpos, qvel, warm, epd, epi, action, frame_skip, obs, out, diag); }
  else { if (f32) env_step_team<float, 8, 1>(blob, qpos, qvel, warm, epd, epi, action, frame_skip, obs, out, diag); else env_step_team<double, 8, 1>(blob, qpos, qvel, warm, epd, epi, action, frame_skip, obs, out, diag); }
}
int hs_nconsts() { return kNumConsts; }
// general = 1: use the NC = 4 variants (extra contact slots: cap-down wheel points, chassis plates) for the flat-floor model too,
// as the launcher does for tilted environments (ackb_kernels.cu, regime split)
void hs_set_general(int v) { g_general = v; }
void hs_substep(int f32, const double* blob, double* qpos, double* qvel, double* warm, const double* ctrl, int nsteps, double* tap, int* diag) {
  const bool scene = blob[0] != 0.0 || g_general;   // model_kind: the obstacle scene has two extra box-contact slots per wheel
  if (f32) { if (scene) substep<float, 4>(blob, qpos, qvel, warm, ctrl, nsteps, tap, diag); else substep<float, 2>(blob, qpos, qvel, warm, ctrl, nsteps, tap, diag); }
  else { if (scene) substep<double, 4>(blob, qpos, qvel, warm, ctrl, nsteps, tap, diag); else substep<double, 2>(blob, qpos, qvel, warm, ctrl, nsteps, tap, diag); }
}
void hs_env_step(int f32, const double* blob, double* qpos, double* qvel, double* warm, double* epd, int* epi, const float* action,
                 int frame_skip, float* obs, float* out, int* diag) {
  const bool scene = blob[0] != 0.0 || g_general;
  if (f32) { if (scene) env_step<float, 4>(blob, qpos, qvel, warm, epd, epi, action, frame_skip, obs, out, diag); else env_step<float, 2>(blob, qpos, qvel, warm, epd, epi, action, frame_skip, obs, out, diag); }
  else { if (scene) env_step<double, 4>(blob, qpos, qvel, warm, epd, epi, action, frame_skip, obs, out, diag); else env_step<double, 2>(blob, qpos, qvel, warm, epd, epi, action, frame_skip, obs, out, diag); }
}
void hs_env_reset(int f32, const double* blob, double* qpos, double* qvel, double* warm, double* epd, int* epi, unsigned long long seed,
                  unsigned env_id, float* obs) {
  const bool scene = blob[0] != 0.0 || g_general;
  if (f32) { if (scene) env_reset<float, 4>(blob, qpos, qvel, warm, epd, epi, seed, env_id, obs); else env_reset<float, 2>(blob, qpos, qvel, warm, epd, epi, seed, env_id, obs); }
  else { if (scene) env_reset<double, 4>(blob, qpos, qvel, warm, epd, epi, seed, env_id, obs); else env_reset<double, 2>(blob, qpos, qvel, warm, epd, epi, seed, env_id, obs); }
}
// observation of a given state (qpos) with given episode data; out = goal distance, min lidar
void hs_observe(int f32, const double* blob, double* qpos, double* qvel, double* warm, const double* epd, float* obs, double* out) {
  if (f32) {
    using E = EnvOps<float, 1, 4>; Consts<float> C; to_consts(blob, C);
    E::State e; Wheel<float, 4> wh[4]; ArrAcc<float> acc{qpos, qvel, warm}; E::load_state(acc, 0, e, wh);
    Episode<float> ep; ep.goal[0] = epd[0]; ep.goal[1] = epd[1]; ep.ref[0] = epd[2]; ep.ref[1] = epd[3]; ep.step_count = 0; ep.episode = 0;
    Kin<float> k; E::S::kinematics(e, k); ObsSink sink{obs}; float d, m; E::observe(C, e, k, ep, 0, sink, &d, &m); out[0] = d; out[1] = m;
  } else {
    using E = EnvOps<double, 1, 4>; Consts<double> C; to_consts(blob, C);
    E::State e; Wheel<double, 4> wh[4]; ArrAcc<double> acc{qpos, qvel, warm}; E::load_state(acc, 0, e, wh);
    Episode<double> ep; ep.goal[0] = epd[0]; ep.goal[1] = epd[1]; ep.ref[0] = epd[2]; ep.ref[1] = epd[3]; ep.step_count = 0; ep.episode = 0;
    Kin<double> k; E::S::kinematics(e, k); ObsSink sink{obs}; double d, m; E::observe(C, e, k, ep, 0, sink, &d, &m); out[0] = d; out[1] = m;
  }
}
// reward / flags from (goal distance, min lidar, step counter before the step); out = reward, terminated, truncated, collision, new counter
void hs_reward(int f32, const double* blob, double dist, double min_lidar, int step_count, double* out) {
  if (f32) {
    Consts<float> C; to_consts(blob, C); Episode<float> ep; ep.step_count = step_count; StepOut<float> so;
    EnvOps<float, 1, 2>::reward_done(C, ep, (float)dist, (float)min_lidar, so);
    out[0] = so.reward; out[1] = so.terminated; out[2] = so.truncated; out[3] = so.collision; out[4] = ep.step_count;
  } else {
    Consts<double> C; to_consts(blob, C); Episode<double> ep; ep.step_count = step_count; StepOut<double> so;
    EnvOps<double, 1, 2>::reward_done(C, ep, dist, min_lidar, so);
    out[0] = so.reward; out[1] = so.terminated; out[2] = so.truncated; out[3] = so.collision; out[4] = ep.step_count;
  }
}
void hs_action_to_ctrl(int f32, const double* blob, float a0, float a1, double* ctrl) {
  if (f32) { Consts<float> C; to_consts(blob, C); float c[4]; action_to_ctrl<float>(C, a0, a1, c); for (int i = 0; i < 4; ++i) ctrl[i] = c[i]; }
  else { Consts<double> C; to_consts(blob, C); double c[4]; action_to_ctrl<double>(C, a0, a1, c); for (int i = 0; i < 4; ++i) ctrl[i] = c[i]; }
}
}
