// ackb_env.cuh -- environment semantics around the physics core: observation, reward, termination,
// episode bookkeeping and reset, exactly as src/rl/envs/ackermann_env.py defines them
// (reset :143-185, step :187-229, _get_observation :231-265, _calculate_reward :267-312) together with
// src/core/odometry.py:62-103,154-170 and src/rl/envs/simple_map_spawner.py:37-52.
#pragma once
#include "ackb_core.cuh"

namespace ackb {

template <typename T>
struct Episode {
  T goal[2];      // goal position in the odometry frame (ackermann_env.py:167-172)
  T ref[2];       // odometry reference = chassis xpos at reset (odometry.py:46-60); z is never read
  int step_count;
  uint32_t episode;  // number of resets so far (Philox counter for the goal / spawn stream)
};

template <typename T>
struct StepOut {
  float reward;
  uint8_t terminated, truncated, collision;
  float goal_distance, min_lidar;
};

template <typename T, int LANES, int NC>
struct EnvOps {
  using S = Sim<T, LANES, NC>;
  using WheelT = Wheel<T, NC>;
  using N = Num<T>;
  using Tm = Team<LANES>;
  static constexpr int WPL = S::WPL;
  using State = typename S::State;

  // MuJoCo dof / qpos addresses of the hinges in internal order (steer L, steer R, spin RL, RR, FL, FR)
  ACKB_HD static int hinge_qadr(int h) { const int t[6] = {9, 11, 7, 8, 10, 12}; return t[h]; }
  ACKB_HD static int hinge_dadr(int h) { const int t[6] = {8, 10, 6, 7, 9, 11}; return t[h]; }

  // A is any accessor with T qpos(int i), T qvel(int i), T warm(int i) and the matching setters
  template <class A>
  ACKB_HD static void load_state(const A& a, int lane, State& e, WheelT* wh) {
    for (int i = 0; i < 3; ++i) { e.p[i] = a.qpos(i); e.vw[i] = a.qvel(i); e.om[i] = a.qvel(3 + i); e.warm_l[i] = a.warm(i); e.warm_a[i] = a.warm(3 + i); }
    for (int i = 0; i < 4; ++i) e.q[i] = a.qpos(3 + i);
    for (int i = 0; i < 2; ++i) { e.st[i] = a.qpos(hinge_qadr(i)); e.dst[i] = a.qvel(hinge_dadr(i)); e.warm_st[i] = a.warm(hinge_dadr(i)); }
#pragma unroll 1
    for (int s = 0; s < WPL; ++s) {
      int h = 2 + S::wheel_index(lane, s);
      wh[s].sp = a.qpos(hinge_qadr(h)); wh[s].dsp = a.qvel(hinge_dadr(h)); wh[s].warm = a.warm(hinge_dadr(h));
    }
  }
  template <class A>
  ACKB_HD static void store_state(A& a, int lane, const State& e, const WheelT* wh) {
    if (lane == 0) {
      for (int i = 0; i < 3; ++i) { a.set_qpos(i, e.p[i]); a.set_qvel(i, e.vw[i]); a.set_qvel(3 + i, e.om[i]); a.set_warm(i, e.warm_l[i]); a.set_warm(3 + i, e.warm_a[i]); }
      for (int i = 0; i < 4; ++i) a.set_qpos(3 + i, e.q[i]);
      for (int i = 0; i < 2; ++i) { a.set_qpos(hinge_qadr(i), e.st[i]); a.set_qvel(hinge_dadr(i), e.dst[i]); a.set_warm(hinge_dadr(i), e.warm_st[i]); }
    }
#pragma unroll 1
    for (int s = 0; s < WPL; ++s) {
      int h = 2 + S::wheel_index(lane, s);
      if (S::PAIR && (lane & 1)) continue;
      a.set_qpos(hinge_qadr(h), wh[s].sp); a.set_qvel(hinge_dadr(h), wh[s].dsp); a.set_warm(hinge_dadr(h), wh[s].warm);
    }
  }

  // observation + reward inputs from the kinematics `k` of state `e`.
  // Sink: void put(int slot, float v).  Every lane computes the scalars; lanes share the lidar slots.
  template <class Sink>
  ACKB_HD static void observe(const Consts<T>& C, const State& e, const Kin<T>& k, const Episode<T>& ep, int lane, Sink& sink,
                              T* goal_distance, T* min_lidar) {
    const int nbeam = (int)C.nbeam[0];
    T mn = T(1e30);
    typename S::LidarBase lb;
    S::lidar_base(C, e, k, lb);
    for (int slot = lane; slot < nbeam; slot += LANES) {
      T d = S::lidar_ray(C, k, lb, tbl_lidar_map(C, slot));
      sink.put(slot, (float)d);
      mn = d < mn ? d : mn;
    }
    mn = Tm::min(mn);
    // odometry (odometry.py:79-87): position relative to the reset reference, yaw from the body quaternion
    const T px = e.p[0] - ep.ref[0], py = e.p[1] - ep.ref[1];
    const T w = e.q[0], x = e.q[1], y = e.q[2], z = e.q[3];
    const T yaw = N::atan2_(T(2) * (w * z + x * y), T(1) - T(2) * (y * y + z * z));
    const T gx = ep.goal[0] - px, gy = ep.goal[1] - py;
    const T dist = N::sqrt_(gx * gx + gy * gy);
    T ang = N::atan2_(gy, gx) - yaw;
    ang = N::wrap_pi(ang);
    if (lane == 0) {
      sink.put(nbeam + 0, (float)px); sink.put(nbeam + 1, (float)py); sink.put(nbeam + 2, (float)yaw);
      sink.put(nbeam + 3, (float)gx); sink.put(nbeam + 4, (float)gy); sink.put(nbeam + 5, (float)dist); sink.put(nbeam + 6, (float)ang);
    }
    *goal_distance = dist;
    *min_lidar = mn;
  }

  // reward / termination (ackermann_env.py:267-312, :216-220), same operation order as the reference
  ACKB_HD static void reward_done(const Consts<T>& C, Episode<T>& ep, T dist, T min_lidar, StepOut<T>& out) {
    const bool term = dist < C.goal_threshold[0];
    const bool coll = min_lidar < C.collision_threshold[0];
    T r = T(0);
    r -= dist * T(0.1);
    if (term) r += T(100);
    if (coll) r -= T(50);
    r -= T(0.01);
    ep.step_count += 1;
    out.reward = (float)r;
    out.terminated = term;
    out.truncated = ep.step_count >= (int)C.max_episode_steps[0];
    out.collision = coll;
    out.goal_distance = (float)dist;
    out.min_lidar = (float)min_lidar;
  }

  // reset (ackermann_env.py:143-172 + simple_map_spawner.py:37-52): spawn pose, zero velocity and warm start,
  // odometry reference := chassis position, goal at U(dmin, dmax) metres in a U(0, 2pi) direction.
  ACKB_HD static void reset_env(const Consts<T>& C, State& e, WheelT* wh, Episode<T>& ep, int lane, uint64_t seed, uint32_t env_id) {
    uint32_t r[4];
    philox4x32(ep.episode, env_id, 0u, 0x41434B42u, (uint32_t)seed, (uint32_t)(seed >> 32), r);
    for (int i = 0; i < 3; ++i) { e.p[i] = C.spawn_qpos[i]; e.vw[i] = e.om[i] = e.warm_l[i] = e.warm_a[i] = T(0); }
    for (int i = 0; i < 4; ++i) e.q[i] = C.spawn_qpos[3 + i];
    if (C.spawn_yaw_range[0] > T(0) || C.spawn_xy_jitter[0] > T(0)) {
      T yaw = C.spawn_yaw_range[0] * (T(2) * (T)u01(r[2]) - T(1));
      uint32_t r2[4];
      philox4x32(ep.episode, env_id, 1u, 0x41434B42u, (uint32_t)seed, (uint32_t)(seed >> 32), r2);
      e.p[0] += C.spawn_xy_jitter[0] * (T(2) * (T)u01(r2[0]) - T(1));
      e.p[1] += C.spawn_xy_jitter[0] * (T(2) * (T)u01(r2[1]) - T(1));
      T cz = N::cos_(yaw * T(0.5)), sz = N::sin_(yaw * T(0.5));
      T q0 = e.q[0], q1 = e.q[1], q2 = e.q[2], q3 = e.q[3];  // q <- qz(yaw) * q
      e.q[0] = cz * q0 - sz * q3; e.q[1] = cz * q1 - sz * q2; e.q[2] = cz * q2 + sz * q1; e.q[3] = cz * q3 + sz * q0;
    }
    const bool maze = C.maze_on[0] != T(0);
    T maze_goal[2] = {T(0), T(0)};
    if (maze) {
      // PointMaze reset (gymnasium_robotics maze.generate_target_goal / generate_reset_pos, restated in compiler/maze.py): goal cell
      // and a different start cell drawn uniformly from the free cells, each with uniform xy noise; identity orientation
      uint32_t r3[4];
      philox4x32(ep.episode, env_id, 3u, 0x41434B42u, (uint32_t)seed, (uint32_t)(seed >> 32), r3);
      const int nx = (int)C.grid_nx[0], ny = (int)C.grid_ny[0];
      int nfree = 0;
      for (int iy = 0; iy < ny; ++iy) { unsigned m = (unsigned)C.maze_free_rows[iy]; for (int ix = 0; ix < nx; ++ix) nfree += (int)((m >> ix) & 1u); }
      int gi = (int)(u01(r3[0]) * (float)nfree); gi = gi >= nfree ? nfree - 1 : gi;
      int si = (int)(u01(r3[1]) * (float)(nfree - 1)); si = si >= nfree - 1 ? nfree - 2 : si;
      if (nfree > 1 && si >= gi) si += 1;          // start cell != goal cell
      if (nfree <= 1) si = gi;
      int cnt = 0;
      T sc[2] = {T(0), T(0)};
      for (int iy = 0; iy < ny; ++iy) {
        const unsigned m = (unsigned)C.maze_free_rows[iy];
        for (int ix = 0; ix < nx; ++ix) {
          if (!((m >> ix) & 1u)) continue;
          const T cx = C.grid_x0[0] + (T(ix) + T(0.5)) * C.grid_pitch[0], cy = C.grid_y0[0] + (T(iy) + T(0.5)) * C.grid_pitch[0];
          if (cnt == gi) { maze_goal[0] = cx; maze_goal[1] = cy; }
          if (cnt == si) { sc[0] = cx; sc[1] = cy; }
          ++cnt;
        }
      }
      const T nz = C.maze_xy_noise[0];
      maze_goal[0] += nz * (T(2) * (T)u01(r3[2]) - T(1)); maze_goal[1] += nz * (T(2) * (T)u01(r3[3]) - T(1));
      e.p[0] = sc[0] + nz * (T(2) * (T)u01(r[2]) - T(1)); e.p[1] = sc[1] + nz * (T(2) * (T)u01(r[3]) - T(1));
    }
    for (int i = 0; i < 2; ++i) { e.st[i] = C.spawn_qpos[hinge_qadr(i)]; e.dst[i] = e.warm_st[i] = T(0); }
#pragma unroll 1
    for (int s = 0; s < WPL; ++s) { wh[s].sp = C.spawn_qpos[hinge_qadr(2 + S::wheel_index(lane, s))]; wh[s].dsp = wh[s].warm = T(0); }
    ep.ref[0] = e.p[0]; ep.ref[1] = e.p[1];
    ep.step_count = 0;
    const T d = C.goal_dmin[0] + (C.goal_dmax[0] - C.goal_dmin[0]) * (T)u01(r[0]);
    const T th = T(6.283185307179586) * (T)u01(r[1]);
    ep.goal[0] = d * N::cos_(th);  // robot_start_position is the odometry origin (0, 0)
    ep.goal[1] = d * N::sin_(th);
    if (maze) { ep.goal[0] = maze_goal[0]; ep.goal[1] = maze_goal[1]; }   // maze goals are WORLD coordinates (…maze_env.py:405-412), quirk kept
    ep.episode += 1;
  }

  // Reset observation of models with settle steps (maze scenes: 3 x mj_step with zero controls before the first odometry call,
  // ackermann_gymnasium_maze_env.py:222-235).  As in step(), what the reference then reads (xpos, sensordata) stems from the
  // forward pass of the LAST settle step, i.e. from the state before its integration: the odometry reference and the reset
  // observation are taken there.  Returns false (and does nothing) when the model has no settle steps.  Warp-uniform.
  template <class Sink>
  ACKB_HD static bool settle_and_observe(const Consts<T>& C, State& e, WheelT* wh, Episode<T>& ep, int lane, Sink& sink, T* dist, T* minl) {
    const int n = (int)C.settle_steps[0];
    if (n <= 0) return false;
    const T ctrl[4] = {T(0), T(0), T(0), T(0)};
    StepDiag diag{};
#pragma unroll 1
    for (int i = 0; i < n; ++i) {
      Kin<T> k;
      S::kinematics(e, k);
      if (i == n - 1) {
        ep.ref[0] = e.p[0]; ep.ref[1] = e.p[1];
        observe(C, e, k, ep, lane, sink, dist, minl);
      }
      S::dynamics(C, e, k, ctrl, lane, wh, diag, (DebugTap<T>*)nullptr, 0);
    }
    return true;
  }

  // mj_resetData after a bad state (mj_checkPos / mj_checkVel / mj_checkAcc): qpos = qpos0, velocities and
  // warm start zero.  MuJoCo also zeroes data.ctrl; the reference env rewrites ctrl at its next step(), so the remaining
  // substeps of THIS env step run with zero controls.  The episode (goal, odometry reference, step counter) is the env's, not
  // MuJoCo's, and is left alone -- exactly what happens to the reference, which never notices the warning.
  ACKB_HD static void reset_data(const Consts<T>& C, State& e, WheelT* wh, int lane) {
    for (int i = 0; i < 3; ++i) { e.p[i] = C.qpos0[i]; e.vw[i] = e.om[i] = e.warm_l[i] = e.warm_a[i] = T(0); }
    for (int i = 0; i < 4; ++i) e.q[i] = C.qpos0[3 + i];
    for (int i = 0; i < 2; ++i) { e.st[i] = C.qpos0[hinge_qadr(i)]; e.dst[i] = e.warm_st[i] = T(0); }
#pragma unroll 1
    for (int s = 0; s < WPL; ++s) { wh[s].sp = C.qpos0[hinge_qadr(2 + S::wheel_index(lane, s))]; wh[s].dsp = wh[s].warm = T(0); }
  }
  // mj_checkPos + mj_checkVel of one environment (team-uniform result)
  ACKB_HD static bool state_is_bad(const State& e, const WheelT* wh) {
    bool bad = false;
    for (int i = 0; i < 3; ++i) bad = bad || is_bad(e.p[i]) || is_bad(e.vw[i]) || is_bad(e.om[i]);
    for (int i = 0; i < 4; ++i) bad = bad || is_bad(e.q[i]);
    for (int i = 0; i < 2; ++i) bad = bad || is_bad(e.st[i]) || is_bad(e.dst[i]);
#pragma unroll 1
    for (int s = 0; s < WPL; ++s) bad = bad || is_bad(wh[s].sp) || is_bad(wh[s].dsp);
    return Tm::sum(bad ? 1 : 0) != 0;
  }

  // one env.step(): frame_skip x mj_step, observation from the kinematics of the last substep (quirk Q3).
  // `emit` is called right after the observation has been written into the sink (before the last substep's dynamics),
  // so that the sink's storage may alias the wheel records.
  template <class Sink, class Emit>
  ACKB_HD static void step_env(const Consts<T>& C, State& e, WheelT* wh, Episode<T>& ep, float a0, float a1, int frame_skip, int lane,
                               Sink& sink, Emit&& emit, StepOut<T>& out, StepDiag& diag, DebugTap<T>* tap, bool cta_sync = false, int rec_stride = 0) {
    T ctrl[4];
    action_to_ctrl<T>(C, a0, a1, ctrl);
    T dist = T(0), minl = T(0);
    for (int s = 0; s < frame_skip; ++s) {
      if (cta_sync) Tm::block_sync();   // CTA-uniform flag (see Team::block_sync)
      // mj_step begins with mj_checkPos / mj_checkVel: a NaN or |x| > 1e10 anywhere in qpos / qvel resets the data
      if (state_is_bad(e, wh)) {
        reset_data(C, e, wh, lane);
        for (int i = 0; i < 4; ++i) ctrl[i] = T(0);
        diag.bad += 1;
      }
      Kin<T> k;
      S::kinematics(e, k);
      if (s == frame_skip - 1) {
        // the sink's storage may alias the wheel records (1 lane per environment): park the live spin state in registers
        T keep[Sink::kAliasesWheels ? 3 * WPL : 1];
        if (Sink::kAliasesWheels) {
#pragma unroll
          for (int i = 0; i < WPL; ++i) { keep[3 * i] = wh[i].sp; keep[3 * i + 1] = wh[i].dsp; keep[3 * i + 2] = wh[i].warm; }
          Tm::warp_sync();
        }
        observe(C, e, k, ep, lane, sink, &dist, &minl);
        emit();
        if (Sink::kAliasesWheels) {
#pragma unroll
          for (int i = 0; i < WPL; ++i) { wh[i].sp = keep[3 * i]; wh[i].dsp = keep[3 * i + 1]; wh[i].warm = keep[3 * i + 2]; }
        }
      }
      // the lidar walk is data dependent: bring the warps of the CTA back into step before the long dynamics code (instruction fetch
      // is shared only while they walk the same lines)
      if (cta_sync && s == frame_skip - 1) Tm::block_sync();
      diag.ncon = 0; diag.nbox = 0; diag.bad_acc = 0;
      S::dynamics(C, e, k, ctrl, lane, wh, diag, tap, rec_stride);
      // mj_checkAcc: a bad qacc resets the data as well.  (MuJoCo then re-runs mj_forward on the reset state and integrates
      // that step; here the reset state is kept as it is -- one substep of free motion from rest is skipped, documented.)
      if (Tm::sum(diag.bad_acc) != 0) {
        reset_data(C, e, wh, lane);
        for (int i = 0; i < 4; ++i) ctrl[i] = T(0);
        diag.bad += 1;
      }
    }
    reward_done(C, ep, dist, minl, out);
  }
};

}  // namespace ackb
