// ackb_core.cuh -- per-environment physics of the Ackermann robot, specialised to its topology.
//
// Replaces, for one environment, what the reference obtains from `mujoco.mj_step`
// (src/rl/envs/ackermann_env.py:200) plus the controller / observation / reward glue around it
// (src/core/controller.py:98-140, src/core/odometry.py:62-103,154-170,
//  src/rl/envs/ackermann_env.py:187-312).
//
// Design:
//   * The kinematic tree is fixed (free chassis + 2 rear spin hinges + 2 x (steer hinge -> spin hinge)),
//     every hinge-mounted body has isotropic inertia with its centre of mass on the hinge anchor
//     (checked by the model compiler).  Expressed in the chassis body frame with generalised
//     accelerations a~ = [R^T a_lin ; a_ang ; hinges], the CRB mass matrix collapses to
//     compile-time constants plus the steer-dependent wheel axes, and the RNE bias has the closed
//     form implemented in `smooth_forces`.
//   * LANES lanes cooperate on one environment (LANES = 4: one lane per wheel).  Each lane owns
//     WPL = 4/LANES wheels: their spin dof, contacts and pyramid rows.  The 8 "shared" dofs
//     (3 lin, 3 ang, 2 steer) are replicated in every lane; wheel contributions to them are summed
//     with warp shuffles (`Team::sum`).
//   * Newton solver on MuJoCo's convex primal cost.  The Hessian is arrow shaped: the spin dofs are
//     eliminated per lane (1x1 pivots), the 8x8 Schur complement is LDL^T-factorised in registers.
//   * The same header compiles for the host with LANES = 1 (tests/hostsim) so that the arithmetic
//     can be debugged against the oracle without a GPU.  That build is test-only.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define ACKB_HD __host__ __device__ __forceinline__
#define ACKB_D __device__ __forceinline__
#ifdef ACKB_COLD_INLINE                                 // tuning: everything inline, as before
#define ACKB_COLD __host__ __device__ __forceinline__
#else
#define ACKB_COLD __host__ __device__ __noinline__      // rarely executed blocks kept out of the hot instruction stream
#endif
#else
#define ACKB_HD inline
#define ACKB_D inline
#define ACKB_COLD inline
#endif

namespace ackb {

// ------------------------------------------------------------------------------------------------
// constants block
// ------------------------------------------------------------------------------------------------
template <typename T>
struct Consts {
#define ACKB_FIELD(name, count) T name[count];
#include "ackb_consts.def"
#undef ACKB_FIELD
};
constexpr int kNumConsts = sizeof(Consts<double>) / sizeof(double);

template <typename T>
ACKB_HD T mjmin(T a, T b) { return a < b ? a : b; }
template <typename T>
ACKB_HD T mjmax(T a, T b) { return a > b ? a : b; }
template <typename T>
ACKB_HD T mjclip(T x, T lo, T hi) { return x < lo ? lo : (x > hi ? hi : x); }  // NaN passes through, like mju_clip

template <typename T> struct Num;
template <> struct Num<float> {
  static constexpr float minval = 1e-15f;
  // device: MUFU.RSQ refined by one Newton step (no IEEE slow path; <= 1 ulp for normal inputs, sqrt(0) = 0, NaN propagates)
  ACKB_HD static float sqrt_(float x) {
#if defined(__CUDA_ARCH__)
    const float r = rsqrtf(x);
    float s = x * r;
    s = fmaf(fmaf(-s, s, x), 0.5f * r, s);
    return x == 0.0f ? 0.0f : s;
#else
    return sqrtf(x);
#endif
  }
  ACKB_HD static float abs_(float x) { return fabsf(x); }
  ACKB_HD static float sin_(float x) { return sinf(x); }
  ACKB_HD static float cos_(float x) { return cosf(x); }
  ACKB_HD static float atan_(float x) { return atanf(x); }
  ACKB_HD static float tan_(float x) { return tanf(x); }
  ACKB_HD static float atan2_(float y, float x) { return atan2f(y, x); }
  // sin/cos of a bounded angle (steer angle, |x| < 1): one MUFU each on the device (abs. error 2^-21.4)
  ACKB_HD static void sincos_small(float x, float* s, float* c) {
#if defined(__CUDA_ARCH__)
    *s = __sinf(x); *c = __cosf(x);
#else
    *s = sinf(x); *c = cosf(x);
#endif
  }
  // sin/cos of a tiny angle (half the rotation of one time step): Taylor series, exact to fp32 for |x| < 0.2
  ACKB_HD static void sincos_tiny(float x, float* s, float* c) {
    if (fabsf(x) < 0.2f) {
      const float x2 = x * x;
      *s = x * (1.0f + x2 * (-1.0f / 6.0f + x2 * (1.0f / 120.0f - x2 * (1.0f / 5040.0f))));
      *c = 1.0f + x2 * (-0.5f + x2 * (1.0f / 24.0f - x2 * (1.0f / 720.0f)));
    } else { *s = sinf(x); *c = cosf(x); }
  }
  // wrap an angle difference of two atan2 results (|x| <= 2 pi) into [-pi, pi]
  ACKB_HD static float wrap_pi(float x) {
    const float pi = 3.14159265358979f;
    return x > pi ? x - 2.0f * pi : (x < -pi ? x + 2.0f * pi : x);
  }
  // reciprocal without the IEEE slow path: MUFU.RCP refined by one Newton step (<= 1 ulp for normal inputs)
  ACKB_HD static float rcp_(float x) {
#if defined(__CUDA_ARCH__)
    float r = __fdividef(1.0f, x);
    return fmaf(r, fmaf(-x, r, 1.0f), r);
#else
    return 1.0f / x;
#endif
  }
  // Newton exit thresholds usable at this precision
  static constexpr float tol_floor = 1e-6f;
  static constexpr float ls_rel = 1e-3f;
};
template <> struct Num<double> {
  static constexpr double minval = 1e-15;
  ACKB_HD static double sqrt_(double x) { return sqrt(x); }
  ACKB_HD static double abs_(double x) { return fabs(x); }
  ACKB_HD static double sin_(double x) { return sin(x); }
  ACKB_HD static double cos_(double x) { return cos(x); }
  ACKB_HD static double atan_(double x) { return atan(x); }
  ACKB_HD static double tan_(double x) { return tan(x); }
  ACKB_HD static double atan2_(double y, double x) { return atan2(y, x); }
  ACKB_HD static void sincos_small(double x, double* s, double* c) { *s = sin(x); *c = cos(x); }
  ACKB_HD static void sincos_tiny(double x, double* s, double* c) { *s = sin(x); *c = cos(x); }
  // the reference wraps with arctan2(sin, cos) (ackermann_env.py:252); kept verbatim in fp64 mode
  ACKB_HD static double wrap_pi(double x) { return atan2(sin(x), cos(x)); }
  ACKB_HD static double rcp_(double x) { return 1.0 / x; }
  static constexpr double tol_floor = 0.0;
  static constexpr double ls_rel = 1e-10;
};

// ------------------------------------------------------------------------------------------------
// lookup tables that lanes of one warp index with DIFFERENT indices (occupancy-grid rows, lidar beam directions and the
// slot -> beam map): on the device they are staged in shared memory once per CTA (`stage_tables`), because a divergent index
// into __constant__ memory is serialised by the constant cache; the host build reads the constants block directly.
// ------------------------------------------------------------------------------------------------
#if defined(__CUDA_ARCH__)
__shared__ unsigned ackb_s_grid[16];
__shared__ float ackb_s_lcos[72], ackb_s_lsin[72];
__shared__ unsigned char ackb_s_lmap[72];
template <typename T>
__device__ __forceinline__ void stage_tables(const Consts<T>& C) {   // every thread of the CTA must call it (ends with a barrier)
  for (int i = threadIdx.x; i < 72; i += blockDim.x) {
    ackb_s_lcos[i] = (float)C.lidar_cos[i]; ackb_s_lsin[i] = (float)C.lidar_sin[i]; ackb_s_lmap[i] = (unsigned char)C.lidar_map[i];
    if (i < 16) ackb_s_grid[i] = (unsigned)C.grid_rows[i];
  }
  __syncthreads();
}
template <typename T> __device__ __forceinline__ unsigned tbl_grid_row(const Consts<T>&, int iy) { return ackb_s_grid[iy]; }
template <typename T> __device__ __forceinline__ int tbl_lidar_map(const Consts<T>&, int slot) { return (int)ackb_s_lmap[slot]; }
// beam directions: fp32 tables for the fp32 kernels, the exact constants for fp64 (the index is lane-uniform enough there)
__device__ __forceinline__ void tbl_beam(const Consts<float>&, int beam, float* c, float* s) { *c = ackb_s_lcos[beam]; *s = ackb_s_lsin[beam]; }
__device__ __forceinline__ void tbl_beam(const Consts<double>& C, int beam, double* c, double* s) { *c = C.lidar_cos[beam]; *s = C.lidar_sin[beam]; }
#else
template <typename T> inline unsigned tbl_grid_row(const Consts<T>& C, int iy) { return (unsigned)C.grid_rows[iy]; }
template <typename T> inline int tbl_lidar_map(const Consts<T>& C, int slot) { return (int)C.lidar_map[slot]; }
template <typename T> inline void tbl_beam(const Consts<T>& C, int beam, T* c, T* s) { *c = C.lidar_cos[beam]; *s = C.lidar_sin[beam]; }
#endif

// ------------------------------------------------------------------------------------------------
// team of LANES lanes working on one environment
// ------------------------------------------------------------------------------------------------
#if !defined(__CUDACC__)
}  // namespace ackb
#include <barrier>
namespace ackb {
// Host emulation of a team of lanes (tests/hostsim only): LANES std::threads run the per-lane code in lock step and meet
// at every collective.  The device code keeps all collectives warp-uniform, so the same call sequence is valid here.
struct HostTeamCtx {
  int lanes;
  std::barrier<>* bar;
  double slot[32];
};
inline thread_local HostTeamCtx* g_host_team = nullptr;
inline thread_local int g_host_lane = 0;
template <typename T, typename F>
inline T host_collective(T v, F&& combine) {
  HostTeamCtx* c = g_host_team;
  c->slot[g_host_lane] = (double)v;
  c->bar->arrive_and_wait();
  T r = (T)c->slot[0];
  for (int i = 1; i < c->lanes; ++i) r = combine(r, (T)c->slot[i]);
  c->bar->arrive_and_wait();
  return r;
}
#endif

template <int LANES>
struct Team {
  // All team collectives are executed by the whole (converged) warp with the full mask: control flow around them is
  // kept warp-uniform with `any()`, lanes of finished environments simply discard their results.
  template <typename T>
  ACKB_D static T sum(T v) {
#if defined(__CUDA_ARCH__)
#pragma unroll
    for (int o = LANES / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
#elif !defined(__CUDACC__)
    if (LANES > 1) v = host_collective(v, [](T a, T b) { return a + b; });
#endif
    return v;
  }
  // sum over the two lanes that share a wheel (8 lanes per environment: lane = 2 * wheel + contact)
  template <typename T>
  ACKB_D static T pair_sum(T v) {
#if defined(__CUDA_ARCH__)
    v += __shfl_xor_sync(0xffffffffu, v, 1);
#elif !defined(__CUDACC__)
    if (LANES > 1) {
      HostTeamCtx* c = g_host_team;
      c->slot[g_host_lane] = (double)v;
      c->bar->arrive_and_wait();
      const T lo = (T)c->slot[g_host_lane & ~1], hi = (T)c->slot[g_host_lane | 1];
      c->bar->arrive_and_wait();
      v = lo + hi;
    }
#endif
    return v;
  }
  template <typename T>
  ACKB_D static T min(T v) {
#if defined(__CUDA_ARCH__)
#pragma unroll
    for (int o = LANES / 2; o > 0; o >>= 1) { T other = __shfl_xor_sync(0xffffffffu, v, o); v = other < v ? other : v; }
#elif !defined(__CUDACC__)
    if (LANES > 1) v = host_collective(v, [](T a, T b) { return b < a ? b : a; });
#endif
    return v;
  }
  template <typename T, int N>
  ACKB_D static void sum_n(T (&v)[N]) {
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] = sum(v[i]);
  }
  // Instruction-cache locality: the warps of a CTA re-converge here so that they walk the same code together and share
  // the fetched lines (the kernel is instruction-fetch bound).  Must be reached by every thread of the CTA.
  // Measured on B200 for every layout (1-lane layout with tail mode: 136 M -> 152 M env-steps/s at 131072 envs, frame_skip 4).
  ACKB_D static void block_sync() {
#if defined(__CUDA_ARCH__)
    __syncthreads();
#endif
  }
  ACKB_D static void warp_sync() {
#if defined(__CUDA_ARCH__)
    __syncwarp();
#endif
  }
  // loop predicate of the solver: warp-wide by default, CTA-wide when ACKB_SYNC_PASS is defined
  ACKB_D static bool loop_any(bool p) {
#if defined(__CUDA_ARCH__) && defined(ACKB_SYNC_PASS)
    return __syncthreads_or(p ? 1 : 0) != 0;
#else
    return any(p);
#endif
  }
  // true if the predicate holds for any lane of the warp (host: any lane of the single environment)
  ACKB_D static bool any(bool p) {
#if defined(__CUDA_ARCH__)
    return __any_sync(0xffffffffu, p) != 0;
#elif !defined(__CUDACC__)
    if (LANES > 1) return host_collective(p ? 1 : 0, [](int a, int b) { return a | b; }) != 0;
    return p;
#else
    return p;
#endif
  }
};

// ------------------------------------------------------------------------------------------------
// small vector helpers
// ------------------------------------------------------------------------------------------------
template <typename T>
ACKB_HD T dot3(const T* a, const T* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
template <typename T>
ACKB_HD void cross3(T* r, const T* a, const T* b) {
  T t0 = a[1] * b[2] - a[2] * b[1], t1 = a[2] * b[0] - a[0] * b[2], t2 = a[0] * b[1] - a[1] * b[0];
  r[0] = t0; r[1] = t1; r[2] = t2;
}

// per-environment state held in registers.  Velocities and warm start use MuJoCo's convention
// (linear part in the world frame, angular part in the chassis frame).
// The spin angle / rate / warm start of each wheel live in its Wheel record (below).
template <typename T>
struct EnvState {
  T p[3], q[4], st[2];
  T vw[3], om[3], dst[2];
  T warm_l[3], warm_a[3], warm_st[2];
};

// quantities of the position stage that the observation needs (pre-integration, reference quirk Q3)
template <typename T>
struct Kin {
  T R[9];      // chassis rotation (row major), from the normalised quaternion
  T n[3], t1[3], t2[3];  // floor contact frame expressed in the body frame
};

struct StepDiag {
  int ncon;         // contacts detected (dist <= 0) over all wheels of this lane
  int unsupported;  // geometry outside the supported contact set
  int niter;
  int nbox;         // wheel-vs-obstacle contacts among ncon
  int bad;          // bad-state resets in this env step (mj_checkPos / mj_checkVel / mj_checkAcc)
  int bad_acc;      // this lane saw a bad solver acceleration in the last substep (consumed by step_env)
};

// mju_isBad: NaN or magnitude beyond mjMAXVAL (1e10)
template <typename T>
ACKB_HD bool is_bad(T x) { return !(Num<T>::abs_(x) <= T(1e10)); }

// index of element (i, j), i >= j, in a packed lower triangle
ACKB_HD constexpr int tri(int i, int j) { return i * (i + 1) / 2 + j; }

// ------------------------------------------------------------------------------------------------
// impedance  d(|pos - margin|)  
// ------------------------------------------------------------------------------------------------
template <typename T>
ACKB_HD T impedance(const T* solimp, T pos) {
  const T lo = T(0.0001), hi = T(0.9999);
  T d0 = mjclip(solimp[0], lo, hi), d1 = mjclip(solimp[1], lo, hi), width = mjmax(T(0), solimp[2]);
  T mid = mjclip(solimp[3], lo, hi), power = mjmax(T(1), solimp[4]);
  if (d0 == d1 || width <= Num<T>::minval) return T(0.5) * (d0 + d1);
  T x = Num<T>::abs_(pos) * Num<T>::rcp_(width);
  if (x >= T(1)) return d1;
  if (x <= T(0)) return d0;
  T y;
  // solimp power is 1 or 2 in both models (checked by the model compiler); other powers are not compiled in
  if (power == T(1)) y = x;
  else y = (x <= mid) ? x * x * Num<T>::rcp_(mid) : T(1) - (T(1) - x) * (T(1) - x) * Num<T>::rcp_(T(1) - mid);
  return d0 + y * (d1 - d0);
}

// ------------------------------------------------------------------------------------------------
// controllers (reference: src/core/controller.py)
// ------------------------------------------------------------------------------------------------
// BicycleController.cmd_vel_to_controls + apply_cmd_vel (controller.py:98-140), every branch and
// epsilon kept, including the inf*0 = NaN corner that MuJoCo answers by zeroing all controls.
template <typename T>
ACKB_HD void bicycle_controller(const Consts<T>& C, T v, T omega, T* ctrl) {
  const T eps = T(1e-5);
  const T L = C.wheelbase[0], Tw = C.track_width[0], rw = C.wheel_radius[0];
  T delta;
  if (Num<T>::abs_(omega) < T(1e-6)) delta = T(0);
  else {
    T sgn = omega > T(0) ? T(1) : (omega < T(0) ? T(-1) : T(0));
    T den = (Num<T>::abs_(v) > eps) ? v : sgn * eps;
    delta = Num<T>::atan_((L * omega) / den);
  }
  const T lim = T(0.6108652381980153);  // deg2rad(35)
  delta = mjclip(delta, -lim, lim);
  T vl, vr;
  if (Num<T>::abs_(delta) < T(1e-6)) vl = vr = v;
  else {
    T tn = Num<T>::tan_(delta);
    T Rt = (Num<T>::abs_(tn) > eps) ? L / tn : T(INFINITY);
    T omega_turn = (Num<T>::abs_(Rt) > eps) ? v / Rt : T(0);
    vl = omega_turn * (Rt - Tw / T(2));
    vr = omega_turn * (Rt + Tw / T(2));
  }
  ctrl[0] = mjclip(delta, T(-0.61), T(0.61));
  ctrl[1] = mjclip(vl / rw, T(-50), T(50));
  ctrl[2] = mjclip(vr / rw, T(-50), T(50));
  ctrl[3] = T(0);
}

// AckermannController (controller.py:42-78) for the 4-actuator scene model.
template <typename T>
ACKB_HD void ackermann_controller(const Consts<T>& C, T v, T omega, T* ctrl) {
  const T L = C.wheelbase[0], Tw = C.track_width[0], rw = C.wheel_radius[0];
  T dl, dr, vl, vr;
  if (Num<T>::abs_(omega) < T(1e-4)) { dl = dr = T(0); vl = vr = v; }
  else {
    T Rt = v / omega;
    T Ri = Rt - Tw / T(2), Ro = Rt + Tw / T(2);
    T ai = Num<T>::atan_(L / Ri), ao = Num<T>::atan_(L / Ro);  // Ri == 0 raises in Python; here atan(inf)
    if (omega > T(0)) { dl = ai; dr = ao; } else { dl = ao; dr = ai; }
    vl = omega * Ri; vr = omega * Ro;
  }
  ctrl[0] = mjclip(dl, T(-0.61), T(0.61));
  ctrl[1] = mjclip(dr, T(-0.61), T(0.61));
  ctrl[2] = mjclip(vl / rw, T(-50), T(50));
  ctrl[3] = mjclip(vr / rw, T(-50), T(50));
}

// action (float32, as in the env) -> ctrl; mirrors ackermann_env.py:190-197
template <typename T>
ACKB_HD void action_to_ctrl(const Consts<T>& C, float a0, float a1, T* ctrl) {
  a0 = mjclip(a0, -1.0f, 1.0f);
  a1 = mjclip(a1, -1.0f, 1.0f);
  float lin = a0 * (float)C.max_linear_velocity[0];   // stays float32 in the reference (NumPy scalar rules)
  float ang = a1 * (float)C.max_angular_velocity[0];
  if (C.ctrl_kind[0] == T(0)) bicycle_controller<T>(C, (T)lin, (T)ang, ctrl);
  else ackermann_controller<T>(C, (T)lin, (T)ang, ctrl);
  // MuJoCo: a bad number in ctrl zeroes all controls for this step (mjWARN_BADCTRL)
  bool bad = false;
  for (int i = 0; i < 4; ++i) bad = bad || !(Num<T>::abs_(ctrl[i]) <= T(1e10));
  if (bad) for (int i = 0; i < 4; ++i) ctrl[i] = T(0);
}

// ------------------------------------------------------------------------------------------------
// per-wheel working set.  With 4 lanes per environment each lane keeps ONE record in registers; with 1 lane per
// environment the 4 records of a thread live in shared memory (odd per-thread stride, conflict free) and the wheel /
// contact loops stay rolled, which keeps the instruction footprint of the solver loop inside the instruction cache.
// ------------------------------------------------------------------------------------------------
template <typename T>
struct Contact {
  T x[3];     // contact point relative to the chassis origin, WORLD axes (the solver works in world axes, see Sim::solve_loop)
  T D;        // 1/R of its pyramid rows (0: contact absent or excluded)
  T z[3];     // residual F Jp a - aref along (n, t1, t2) at the current point
  T zv[3];    // image F Jp x of the current step direction
};

// Contact frames.  Slots 0,1 of a wheel are floor contacts (frame of the floor normal +z).  Slots 2,3 (obstacle scene)
// are wheel-vs-box contacts whose frame is one of the world-axis-aligned frames below, selected by a face code; the
// contact normal points from the wheel (geom1) to the box (geom2), so the rows act on the wheel with a minus sign,
// which is folded into the frame (all three rows negated).
//   code 0: floor (+z)   1: normal +x   2: normal -x   3: normal +y   4: normal -y   5: normal -z (box top)
template <typename T>
ACKB_HD void contact_frame(const Kin<T>& k, int code, T* n, T* t1, T* t2) {
  // rows of R are the world axes expressed in the body frame
  const T* Rx = k.R; const T* Ry = k.R + 3; const T* Rz = k.R + 6;
  T sn = T(1), s2 = T(1);
  const T *pn = Rz, *p1 = Ry, *p2 = Rx;
  switch (code) {
    case 0: sn = T(1); pn = Rz; p1 = Ry; p2 = Rx; s2 = T(-1); break;                    // (+z, +y, -x)
    case 1: pn = Rx; p1 = Ry; p2 = Rz; sn = T(-1); s2 = T(-1); break;                    // -( +x, +y, +z)
    case 2: pn = Rx; p1 = Ry; p2 = Rz; sn = T(1); s2 = T(1); break;                      // -( -x, +y, -z)
    case 3: pn = Ry; p1 = Rz; p2 = Rx; sn = T(-1); s2 = T(-1); break;                    // -( +y, +z, +x)
    case 4: pn = Ry; p1 = Rz; p2 = Rx; sn = T(1); s2 = T(1); break;                      // -( -y, +z, -x)
    default: pn = Rz; p1 = Ry; p2 = Rx; sn = T(1); s2 = T(-1); break;                    // -( -z, +y, +x)
  }
  const T s1 = (code == 0) ? T(1) : T(-1);
#pragma unroll
  for (int i = 0; i < 3; ++i) { n[i] = sn * pn[i]; t1[i] = s1 * p1[i]; t2[i] = s2 * p2[i]; }
}

// Every contact frame of the two models is a signed permutation of the world axes (floor: (n, t1, t2) = (+z, +y, -x); box
// faces: see contact_frame).  In world axes a frame is therefore applied with selects instead of 3 x 3 products:
//   pat 0: (n, t1, t2) along (x, y, z)   pat 1: along (y, z, x)   pat 2: along (z, y, x);  sn, s1, s2 = +-1.
// Packed form (8 bits): pat | (sn < 0) << 2 | (s1 < 0) << 3 | (s2 < 0) << 4.
template <typename T>
struct Perm {
  int pat;
  T sn, s1, s2;
};
ACKB_HD constexpr unsigned perm_pack(int code) {
  // code 0 floor (+z,+y,-x); 1: -(+x,+y,+z); 2: -(-x,+y,-z); 3: -(+y,+z,+x); 4: -(-y,+z,-x); 5: -(-z,+y,+x)   (contact_frame)
  return code == 0 ? (2u | 16u) : code == 1 ? (0u | 4u | 8u | 16u) : code == 2 ? (0u | 8u) : code == 3 ? (1u | 4u | 8u | 16u)
       : code == 4 ? (1u | 8u) : (2u | 8u | 16u);
}
template <typename T>
ACKB_HD Perm<T> perm_unpack(unsigned p) {
  Perm<T> q;
  q.pat = (int)(p & 3u);
  q.sn = (p & 4u) ? T(-1) : T(1); q.s1 = (p & 8u) ? T(-1) : T(1); q.s2 = (p & 16u) ? T(-1) : T(1);
  return q;
}
// frame components (n, t1, t2) of a world vector y
template <typename T>
ACKB_HD void perm_apply(const Perm<T>& q, const T* y, T* out) {
  out[0] = q.sn * (q.pat == 0 ? y[0] : (q.pat == 1 ? y[1] : y[2]));
  out[1] = q.s1 * (q.pat == 1 ? y[2] : y[1]);
  out[2] = q.s2 * (q.pat == 0 ? y[2] : y[0]);
}
// world vector of frame components phi
template <typename T>
ACKB_HD void perm_apply_t(const Perm<T>& q, const T* phi, T* out) {
  const T a = q.sn * phi[0], b = q.s1 * phi[1], c = q.s2 * phi[2];
  out[0] = q.pat == 0 ? a : c;
  out[1] = q.pat == 1 ? a : b;
  out[2] = q.pat == 0 ? c : (q.pat == 1 ? b : a);
}
// world 3 x 3 (symmetric) of the frame-space matrix [[A, E, F], [E, B, 0], [F, 0, Cc]] (rows / columns n, t1, t2)
template <typename T>
ACKB_HD void perm_sym(const Perm<T>& q, T A, T B, T Cc, T E, T F, T (*S3)[3]) {
  E *= q.sn * q.s1; F *= q.sn * q.s2;
  S3[0][0] = q.pat == 0 ? A : Cc;
  S3[1][1] = q.pat == 1 ? A : B;
  S3[2][2] = q.pat == 0 ? Cc : (q.pat == 1 ? B : A);
  S3[0][1] = S3[1][0] = q.pat == 0 ? E : (q.pat == 1 ? F : T(0));
  S3[0][2] = S3[2][0] = q.pat == 1 ? T(0) : F;
  S3[1][2] = S3[2][1] = q.pat == 0 ? T(0) : E;
}

template <typename T, int NC>
struct Wheel {
  T sp, dsp, warm;   // state: spin angle, rate, warm-start acceleration
  T ax, ay;          // spin axis (ax, ay, 0) in the body frame
  T axw[3], cww[3];  // spin axis and wheel centre (relative to the chassis origin) in world axes
  T a, x, tau;       // solver: spin acceleration, step direction, smooth force
  T g, cw;           // gradient entry and Schur pivot of the spin dof at the last assembly
  T b[7];            // coupling of the spin dof with (lin3, ang3, own steer)
  unsigned zone0;    // activity pattern of the wheel's rows at the last assembly
  unsigned fcode;    // packed frame permutations of the box contact slots (8 bits each, slots 2 and 3; see Perm)
  Contact<T> con[NC];
};

// wheel constants by wheel index wi (RL, RR, FL, FR)
template <typename T>
struct WheelK {
  T c[3], isL, isR, J, cdiag, mu, flf, flR, flD, flB, damp;
  int hidx;
};
template <typename T>
ACKB_HD WheelK<T> wheel_consts(const Consts<T>& C, int wi) {
  WheelK<T> k;
  k.c[0] = C.w_center[3 * wi]; k.c[1] = C.w_center[3 * wi + 1]; k.c[2] = C.w_center[3 * wi + 2];
  k.isL = (wi == 2) ? T(1) : T(0);
  k.isR = (wi == 3) ? T(1) : T(0);
  k.hidx = 2 + wi;
  k.J = C.h_inertia[2 + wi];
  k.cdiag = k.J + C.h_armature[2 + wi];
  k.mu = C.w_mu[wi];
  k.flf = C.h_floss[2 + wi]; k.flR = C.h_flR[2 + wi]; k.flD = C.h_flD[2 + wi]; k.flB = C.h_flB[2 + wi];
  k.damp = C.h_damping[2 + wi];
  return k;
}

// spin column u = axis x r and steer column w = ez x r of the point Jacobian (r = X - centre)
template <typename T, int NC>
ACKB_HD void contact_cols(const Wheel<T, NC>& w, const WheelK<T>& wk, const T* X, T* u, T* wv) {
  const T r0 = X[0] - wk.c[0], r1 = X[1] - wk.c[1], r2 = X[2] - wk.c[2];
  u[0] = w.ay * r2; u[1] = -w.ax * r2; u[2] = w.ax * r1 - w.ay * r0;   // (ax, ay, 0) x r
  const T st = wk.isL + wk.isR;
  wv[0] = -r1 * st; wv[1] = r0 * st;
}

// world axes: spin column u = axis_w x r and steer column wv = st * (ez_w x r) of the point Jacobian, r = X - centre_w
template <typename T, int NC>
ACKB_HD void contact_cols_w(const Wheel<T, NC>& w, T st, const T* ezw, const T* X, T* u, T* wv) {
  const T r[3] = {X[0] - w.cww[0], X[1] - w.cww[1], X[2] - w.cww[2]};
  cross3(u, w.axw, r);
  cross3(wv, ezw, r);
  wv[0] *= st; wv[1] *= st; wv[2] *= st;
}
// world axes: point "acceleration" y = a_lin + a_ang x X + a_spin u + a_steer wv
template <typename T>
ACKB_HD void point_accel_w(const T* X, const T* u, const T* wv, const T* alin, const T* aang, T aspin, T asteer, T* y) {
  cross3(y, aang, X);
#pragma unroll
  for (int i = 0; i < 3; ++i) y[i] += alin[i] + aspin * u[i] + asteer * wv[i];
}

// point "acceleration" y = a_lin + a_ang x X + a_spin u + a_steer w, projected on the contact frame (n, t1, t2)
template <typename T>
ACKB_HD void project_point(const T* fn, const T* ft1, const T* ft2, const T* X, const T* u, const T* wv, const T* alin, const T* aang, T aspin,
                           T asteer, T* out) {
  T y[3];
  cross3(y, aang, X);
  y[0] += alin[0] + aspin * u[0] + asteer * wv[0];
  y[1] += alin[1] + aspin * u[1] + asteer * wv[1];
  y[2] += alin[2] + aspin * u[2];
  out[0] = dot3(fn, y); out[1] = dot3(ft1, y); out[2] = dot3(ft2, y);
}

// forces of the four pyramid rows of a contact at residual z (3-vector in the contact frame);
// returns cost, fills phi = contact-frame force and the quadratic-zone flags q[4]
template <typename T>
ACKB_HD T pyramid_rows(T D, T mu, const T* z, T* phi, T* q) {
  T x[4] = {z[0] + mu * z[1], z[0] - mu * z[1], z[0] + mu * z[2], z[0] - mu * z[2]};
  T f[4], cost = T(0);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    bool act = x[i] < T(0);
    q[i] = act ? T(1) : T(0);
    f[i] = act ? -D * x[i] : T(0);
    cost += act ? T(0.5) * D * x[i] * x[i] : T(0);
  }
  phi[0] = f[0] + f[1] + f[2] + f[3];
  phi[1] = mu * (f[0] - f[1]);
  phi[2] = mu * (f[2] - f[3]);
  return cost;
}

// friction-loss (Huber) row: returns cost, force and quadratic flag
// (select form: no divergent branches; D = 1 / R is passed in)
template <typename T>
ACKB_HD T floss_row(T x, T f, T R, T D, T* force, T* quad) {
  const T rf = R * f;
  const bool lo = x <= -rf, hi = x >= rf;
  const bool lin = lo || hi;
  *force = lo ? f : (hi ? -f : -D * x);
  *quad = lin ? T(0) : T(1);
  return lin ? f * (T(-0.5) * rf + Num<T>::abs_(x)) : T(0.5) * D * x * x;
}

// rows that involve only the shared dofs: steering equality, steer friction loss, steer limits
template <typename T>
struct SharedRows {
  T eqD, eq_aref;            // D = 0 when the model has no equality
  T flD[2], flR[2], flf[2], fl_aref[2];
  T limD[2], lim_sign[2], lim_aref[2];  // sign = +1 lower limit active, -1 upper; D = 0 inactive
};

// evaluate shared rows at steer accelerations (aL, aR): cost, generalised force on (sL, sR) and
// the Hessian contribution (3 numbers: LL, LR, RR)
template <typename T>
ACKB_HD T shared_rows_eval(const SharedRows<T>& s, T aL, T aR, T* fL, T* fR, T* hLL, T* hLR, T* hRR) {
  T cost = T(0);
  *fL = *fR = T(0); *hLL = *hLR = *hRR = T(0);
  {  // equality  (aL - aR) - aref
    T x = aL - aR - s.eq_aref, f = -s.eqD * x;
    cost += T(0.5) * s.eqD * x * x;
    *fL += f; *fR -= f; *hLL += s.eqD; *hRR += s.eqD; *hLR -= s.eqD;
  }
  const T a[2] = {aL, aR};
  T* fo[2] = {fL, fR};
  T* ho[2] = {hLL, hRR};
  for (int i = 0; i < 2; ++i) {
    T f, q;
    cost += floss_row(a[i] - s.fl_aref[i], s.flf[i], s.flR[i], s.flD[i], &f, &q);
    *fo[i] += f; *ho[i] += q * s.flD[i];
    T x = s.lim_sign[i] * a[i] - s.lim_aref[i];
    bool act = x < T(0);
    T fl = act ? -s.limD[i] * x : T(0);
    cost += act ? T(0.5) * s.limD[i] * x * x : T(0);
    *fo[i] += s.lim_sign[i] * fl;
    *ho[i] += act ? s.limD[i] : T(0);
  }
  return cost;
}


// which zone each shared row is in at steer accelerations (aL, aR): friction loss (-, quadratic, +), limit on/off
template <typename T>
ACKB_HD unsigned shared_rows_zone(const SharedRows<T>& s, T aL, T aR) {
  const T a[2] = {aL, aR};
  unsigned zn = 0u;
  for (int i = 0; i < 2; ++i) {
    const T x = a[i] - s.fl_aref[i], rf = s.flR[i] * s.flf[i];
    zn = (zn << 2) | (x <= -rf ? 0u : (x >= rf ? 2u : 1u));
    zn = (zn << 1) | ((s.limD[i] > T(0) && s.lim_sign[i] * a[i] - s.lim_aref[i] < T(0)) ? 1u : 0u);
  }
  return zn;
}

// ------------------------------------------------------------------------------------------------
// LDL^T of a packed symmetric positive definite 8x8, in registers (fully unrolled).
// On exit the strict lower triangle holds L and the diagonal holds 1/d.
// ------------------------------------------------------------------------------------------------
// All loops have constant trip counts (the triangular bounds are expressed as conditions on the unrolled indices) so
// that every access has a compile-time index and S stays in registers.
template <typename T>
ACKB_HD void ldl8_factor(T* S) {
  T d[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    T v[8];
    T dj = S[tri(j, j)];
#pragma unroll
    for (int k = 0; k < 8; ++k)
      if (k < j) { v[k] = S[tri(j, k)] * d[k]; dj -= S[tri(j, k)] * v[k]; }
    d[j] = dj;
    T dinv = Num<T>::rcp_(dj);
    S[tri(j, j)] = dinv;
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (i > j) {
        T s = S[tri(i, j)];
#pragma unroll
        for (int k = 0; k < 8; ++k)
          if (k < j) s -= S[tri(i, k)] * v[k];
        S[tri(i, j)] = s * dinv;
      }
  }
}
template <typename T>
ACKB_HD void ldl8_solve(const T* S, T* x /* in: rhs, out: solution */) {
#pragma unroll
  for (int i = 1; i < 8; ++i)
#pragma unroll
    for (int k = 0; k < 8; ++k)
      if (k < i) x[i] -= S[tri(i, k)] * x[k];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] *= S[tri(i, i)];
#pragma unroll
  for (int i = 6; i >= 0; --i)
#pragma unroll
    for (int k = 0; k < 8; ++k)
      if (k > i) x[i] -= S[tri(k, i)] * x[k];
}

// Philox4x32-10 counter-based generator (Salmon et al. 2011), used for goals, spawn jitter and the
// synthetic random actions of the benchmark.
ACKB_HD void philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t* out) {
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
ACKB_HD float u01(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }  // [0, 1)

// debugging taps for the host build (tests/hostsim); null on the device
template <typename T>
struct DebugTap {
  T tau[12], a_smooth[12], a[12], fc[12];  // internal order: lin3 ang3 sL sR spinRL spinRR spinFL spinFR
  int niter, nls;
};

// ------------------------------------------------------------------------------------------------
// the simulator
// ------------------------------------------------------------------------------------------------
// LANES lanes per environment.  1, 2, 4: each lane owns WPL = 4/LANES whole wheels (records with NC contact slots).
// 8: each lane owns ONE floor contact of a wheel (lane = 2 * wheel + contact, records with NC = 1 slot); the spin dof of
// the wheel is replicated in the two lanes of the pair, its contact-dependent pivots are pair-summed, and every per-wheel
// term that enters a team sum carries the weight 1/2.
template <typename T, int LANES, int NC>
struct Sim {
  static constexpr bool PAIR = (LANES == 8);
  static constexpr int WPL = PAIR ? 1 : 4 / LANES;
  static constexpr int NCF = PAIR ? 1 : 2;          // floor-contact slots of a record
  using WheelT = Wheel<T, NC>;
  ACKB_HD static int wheel_index(int lane, int s) { return PAIR ? (lane >> 1) : lane * WPL + s; }
  ACKB_HD static T pair_weight() { return PAIR ? T(0.5) : T(1); }
  using Tm = Team<LANES>;
  using N = Num<T>;
  using State = EnvState<T>;
  // contact loops are unrolled when the wheel record lives in registers (WPL == 1) and rolled when it is in shared memory
#ifndef ACKB_CU_SMEM
#define ACKB_CU_SMEM 2   // contacts of a wheel unrolled (ILP), wheels rolled (code size)
#endif
  // records live in registers when a lane owns one wheel.  (Measured alternative for NC = 4, ACKB_REG_RECORDS_NC4=0: records in
  // shared memory with a rolled contact-group loop -- smaller solver loop, but 5 % slower on the scene: 52.0 vs 54.8 M env-steps/s.)
#ifndef ACKB_REG_RECORDS_NC4
#define ACKB_REG_RECORDS_NC4 1
#endif
  static constexpr bool kRegRecords = (WPL == 1) && (NC <= 2 || ACKB_REG_RECORDS_NC4 != 0);
  static constexpr int CU = kRegRecords ? NC : ACKB_CU_SMEM;
  // Contact loops of the solver: slots are visited in groups of CSTEP; groups at or beyond `ncs` (a warp-uniform count: 2 when no
  // environment of the warp has a wheel-box contact in this substep) are skipped.  Register records: fully unrolled (static
  // indices); shared-memory records: the group loop stays rolled.
  static constexpr int CSTEP = (NC >= 2) ? (kRegRecords ? 2 : (ACKB_CU_SMEM >= 2 ? 2 : 1)) : 1;
  static constexpr int CO = kRegRecords ? (NC / CSTEP) : 1;
#define ACKB_CONTACTS_BEGIN(c)                                                     \
  _Pragma("unroll(CO)") for (int c##_g = 0; c##_g < NC; c##_g += CSTEP) {          \
    if (c##_g < ncs) {                                                             \
      _Pragma("unroll") for (int c = c##_g; c < c##_g + CSTEP; ++c) {
#define ACKB_CONTACTS_END }}}
#ifndef ACKB_TAIL_MODE
#define ACKB_TAIL_MODE 1
#endif
#ifndef ACKB_TAIL_ENVS
#define ACKB_TAIL_ENVS 8
#endif
  static constexpr int kTailEnvs = ACKB_TAIL_ENVS;
  // team sum inside the solver loop: the layout's own team, or the 4-lane teams of tail mode
  template <int NN>
  ACKB_D static void team_sum_n(T (&v)[NN], bool TL) {
#if defined(__CUDA_ARCH__)
    if (TL) {
#pragma unroll
      for (int i = 0; i < NN; ++i) { v[i] += __shfl_xor_sync(0xffffffffu, v[i], 1); v[i] += __shfl_xor_sync(0xffffffffu, v[i], 2); }
      return;
    }
#endif
    Tm::sum_n(v);
  }

  // frame / kind code of contact slot c of a wheel record: slots 0, 1 are floor contacts of the wheel; slots >= 2 carry a packed frame
  // permutation (bits 0..4, see Perm) and bit 5 = chassis-plate contact (acts on the chassis only, its own friction coefficient)
  ACKB_HD static unsigned slot_code(const WheelT& w, int c) { return (NC <= 2 || c < 2) ? perm_pack(0) : ((w.fcode >> (8 * (c - 2))) & 255u); }
  ACKB_HD static T slot_coupling(unsigned code) { return (NC > 2 && (code & 32u)) ? T(0) : T(1); }
  ACKB_HD static T slot_mu(const Consts<T>& C, const WheelK<T>& wk, unsigned code) { return (NC > 2 && (code & 32u)) ? C.pl_mu[0] : wk.mu; }

  // Team sum of the 45 assembly partials (36 packed Hessian entries, 8 reduced right-hand sides, 1 scalar).  With one lane per
  // wheel (4-lane layout, tail mode) the steer rows are produced by a single lane each -- row / rhs 6 by the front-left wheel's
  // lane (2), row / rhs 7 by the front-right one (3) -- so those 16 entries are broadcast instead of butterfly-summed
  // (bit-identical: the other lanes hold exact zeros).
  ACKB_D static void team_sum_part(T (&part)[45], bool TL) {
#if defined(__CUDA_ARCH__)
    if (TL || LANES == 4) {
      const int base = (int)(threadIdx.x & 28u);     // first lane of this lane's team of 4
#pragma unroll
      for (int i = 0; i < 45; ++i) {
        const bool row6 = (i >= tri(6, 0) && i <= tri(6, 6)) || i == 36 + 6;
        const bool row7 = (i >= tri(7, 0) && i <= tri(7, 7)) || i == 36 + 7;
        if (row6) part[i] = __shfl_sync(0xffffffffu, part[i], base + 2);
        else if (row7) part[i] = __shfl_sync(0xffffffffu, part[i], base + 3);
        else { part[i] += __shfl_xor_sync(0xffffffffu, part[i], 1); part[i] += __shfl_xor_sync(0xffffffffu, part[i], 2); }
      }
      return;
    }
#endif
    team_sum_n(part, TL);
  }

  // ---- B1 kinematics: normalise the quaternion (written back, like mj_kinematics), rotation, floor frame
  ACKB_HD static void kinematics(State& e, Kin<T>& k) {
    T qn = N::sqrt_(e.q[0] * e.q[0] + e.q[1] * e.q[1] + e.q[2] * e.q[2] + e.q[3] * e.q[3]);
    if (qn < N::minval) { e.q[0] = T(1); e.q[1] = e.q[2] = e.q[3] = T(0); }
    else { T inv = N::rcp_(qn); for (int i = 0; i < 4; ++i) e.q[i] *= inv; }
    const T w = e.q[0], x = e.q[1], y = e.q[2], z = e.q[3];
    k.R[0] = w * w + x * x - y * y - z * z; k.R[1] = T(2) * (x * y - w * z); k.R[2] = T(2) * (x * z + w * y);
    k.R[3] = T(2) * (x * y + w * z); k.R[4] = w * w - x * x + y * y - z * z; k.R[5] = T(2) * (y * z - w * x);
    k.R[6] = T(2) * (x * z - w * y); k.R[7] = T(2) * (y * z + w * x); k.R[8] = w * w - x * x - y * y + z * z;
    // contact frame of a floor contact in world axes is (n, t1, t2) = (+z, +y, -x); body components = R^T d
    for (int i = 0; i < 3; ++i) { k.n[i] = k.R[6 + i]; k.t1[i] = k.R[3 + i]; k.t2[i] = -k.R[i]; }
  }

  // ---- B6/B7 floor contacts of one wheel (plane vs cylinder, in the body frame) and their row parameters
  // `tri` (may be null): receives the two extra "triangle" points mjc_PlaneCylinder adds when the cap faces the floor
  // (tri[0..2], tri[3..5] = body-frame positions, tri[6] = their common distance, tri[7] = 1 if present).
  ACKB_HD static void collide_wheel(const Consts<T>& C, const State& e, const Kin<T>& k, const T* vb, int wi, int cown, const WheelK<T>& wk,
                                    WheelT& w, StepDiag& diag, T* tri = nullptr) {
    const T s = wk.isL * e.st[0] + wk.isR * e.st[1];
    T sn, cs;
    N::sincos_small(s, &sn, &cs);
    w.ax = -sn; w.ay = cs;                       // Rz(s) * (0, 1, 0)
    const T dsteer = wk.isL * e.dst[0] + wk.isR * e.dst[1];
    const T r = C.w_radius[wi], hl = C.w_halflen[wi];
    const T hO = e.p[2] - C.plane_z[0];
    T ax[3] = {w.ax, w.ay, T(0)};
    T prjaxis = dot3(k.n, ax);
    if (prjaxis > T(0)) { ax[0] = -ax[0]; ax[1] = -ax[1]; prjaxis = -prjaxis; }
    const T dist = dot3(k.n, wk.c) + hO;
    T vec[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) vec[i] = ax[i] * prjaxis - k.n[i];
    T len = N::sqrt_(dot3(vec, vec));
    if (len < N::minval) {
      // disk parallel to the floor: mjc_PlaneCylinder takes the cylinder's own x axis, Rz(steer) Ry(spin) e_x in the body frame
      const T sp_s = N::sin_(w.sp), sp_c = N::cos_(w.sp);
      vec[0] = cs * sp_c; vec[1] = sn * sp_c; vec[2] = -sp_s;
      len = T(1);
    }
    const T rlen = r * N::rcp_(len);
#pragma unroll
    for (int i = 0; i < 3; ++i) vec[i] *= rlen;
    const T prjvec = dot3(vec, k.n);
    ax[0] *= hl; ax[1] *= hl;
    prjaxis *= hl;
    const T d0 = dist + prjaxis + prjvec, d1 = dist - prjaxis + prjvec;
    const bool has0 = d0 <= T(0), has1 = has0 && (d1 <= T(0));
    {   // cap faces the floor: two more points at +-(sqrt(3)/2) r along vec x axis, half a radius back from the rim point
      const T dtri = dist + prjaxis - T(0.5) * prjvec;
      const bool has_tri = has0 && (dtri <= T(0));
      if (tri) {
        tri[7] = has_tri ? T(1) : T(0);
        if (has_tri) {
          T v1[3];
          cross3(v1, vec, ax);
          const T l1 = N::sqrt_(dot3(v1, v1));
          const T sc = (l1 < N::minval) ? T(0) : r * T(0.8660254037844386) * N::rcp_(l1);
          // a degenerate cross product (vec parallel to the axis) cannot occur here: vec is the radial direction
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            const T base = wk.c[i] + ax[i] - vec[i] * T(0.5) - k.n[i] * dtri * T(0.5);
            tri[i] = base + sc * v1[i];
            tri[3 + i] = base - sc * v1[i];
          }
          tri[6] = dtri;
          if (l1 < N::minval) diag.unsupported = 1;
        }
      } else if (has_tri) diag.unsupported = 1;
    }
    if (PAIR) diag.ncon += (cown == 0) ? (has0 ? 1 : 0) : (has1 ? 1 : 0);
    else diag.ncon += (has0 ? 1 : 0) + (has1 ? 1 : 0);
    const T mu = wk.mu;
#pragma unroll(CU)
    for (int cc = 0; cc < NCF; ++cc) {
      const int c = PAIR ? cown : cc;
      Contact<T>& con = w.con[cc];
      const T dd = (c == 0) ? d0 : d1, sg = (c == 0) ? T(1) : T(-1);
      const bool has = (c == 0) ? has0 : has1;
      con.x[0] = wk.c[0] + vec[0] + sg * ax[0] - k.n[0] * dd * T(0.5);
      con.x[1] = wk.c[1] + vec[1] + sg * ax[1] - k.n[1] * dd * T(0.5);
      con.x[2] = wk.c[2] + vec[2] - k.n[2] * dd * T(0.5);
      const bool active = has && (dd < T(0));  // dist >= includemargin(0): counted but excluded
      const T imp = impedance(&C.w_solimp[5 * wi], dd);
      const T R0 = mjmax(N::minval, (T(1) - imp) * N::rcp_(imp) * C.w_tran[wi] * (T(1) + mu * mu));
      con.D = active ? N::rcp_(T(2) * C.w_mureg2[wi] * R0) : T(0);
      T u[3], wv[2], vel[3];
      contact_cols(w, wk, con.x, u, wv);
      project_point(k.n, k.t1, k.t2, con.x, u, wv, vb, e.om, w.dsp, dsteer, vel);
      // residual at a = 0 is minus the reference acceleration
      con.z[0] = C.w_B[wi] * vel[0] + C.w_K[wi] * imp * dd;
      con.z[1] = C.w_B[wi] * vel[1];
      con.z[2] = C.w_B[wi] * vel[2];
      // the solver works in world axes: keep the point as R X (relative to the chassis origin)
      const T xb[3] = {con.x[0], con.x[1], con.x[2]};
#pragma unroll
      for (int i = 0; i < 3; ++i) con.x[i] = k.R[3 * i] * xb[0] + k.R[3 * i + 1] * xb[1] + k.R[3 * i + 2] * xb[2];
    }
  }

  // ---- wheel vs static axis-aligned boxes (obstacle scene).  MuJoCo sends cylinder-box pairs to its general convex
  // collider (one contact along the minimum-penetration direction); here that contact is computed in closed form over the
  // box's face normals with the cylinder's support point (see oracle cylinder_box, DESIGN.md for the deviation).
  // Up to two box contacts per wheel are kept (slots 2, 3); more are flagged as unsupported.
  ACKB_HD static void collide_boxes(const Consts<T>& C, const State& e, const Kin<T>& k, const T* vb, int wi, const WheelK<T>& wk,
                                    WheelT& w, StepDiag& diag, int& nfound, unsigned& fcode) {
    const T r = C.w_radius[wi], hl = C.w_halflen[wi];
    // wheel centre and axis in the world frame
    T cw[3], aw[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      cw[i] = e.p[i] + k.R[3 * i] * wk.c[0] + k.R[3 * i + 1] * wk.c[1] + k.R[3 * i + 2] * wk.c[2];
      aw[i] = k.R[3 * i] * w.ax + k.R[3 * i + 1] * w.ay;
    }
    const T dsteer = wk.isL * e.dst[0] + wk.isR * e.dst[1];
    const T reach = N::sqrt_(r * r + hl * hl);
    const int nbox = (int)C.nbox[0];
    // candidates: every box (no grid), or the occupied cells under the wheel's bounding square, visited in box order
    const bool grid = C.grid_on[0] != T(0);
    int gx0 = 0, gy0 = 0, gny = 1, ncand = nbox;
    if (grid) {
      const T ip = N::rcp_(C.grid_pitch[0]);
      gx0 = (int)floor((cw[0] - reach - C.grid_x0[0]) * ip); gy0 = (int)floor((cw[1] - reach - C.grid_y0[0]) * ip);
      const int gx1 = (int)floor((cw[0] + reach - C.grid_x0[0]) * ip), gy1 = (int)floor((cw[1] + reach - C.grid_y0[0]) * ip);
      gny = gy1 - gy0 + 1;
      ncand = (gx1 - gx0 + 1) * gny;
      if (ncand > 16) ncand = 16;   // reach << pitch: 1 to 4 cells
    }
#pragma unroll 1
    for (int bi = 0; bi < ncand; ++bi) {
      T bcx, bcy;
      if (grid) {
        const int ix = gx0 + bi / gny, iy = gy0 + bi % gny;
        if (ix < 0 || iy < 0 || ix >= (int)C.grid_nx[0] || iy >= (int)C.grid_ny[0]) continue;
        if (!((tbl_grid_row(C, iy) >> ix) & 1u)) continue;
        bcx = C.grid_x0[0] + (T(ix) + T(0.5)) * C.grid_pitch[0]; bcy = C.grid_y0[0] + (T(iy) + T(0.5)) * C.grid_pitch[0];
      } else { bcx = C.box_cx[bi]; bcy = C.box_cy[bi]; }
      const T c[3] = {cw[0] - bcx, cw[1] - bcy, cw[2] - C.box_z[0]};
      if (N::abs_(c[0]) > C.box_half[0] + reach || N::abs_(c[1]) > C.box_half[1] + reach || N::abs_(c[2]) > C.box_half[2] + reach) continue;
      T best = T(-1e30), bp[3] = {T(0), T(0), T(0)};
      int bk = 0, bs = 1;
#pragma unroll
      for (int kk = 0; kk < 3; ++kk)
#pragma unroll
        for (int sgn = -1; sgn <= 1; sgn += 2) {
          const T an = aw[kk] * T(sgn);
          T wv3[3] = {an * aw[0], an * aw[1], an * aw[2]};
          wv3[kk] -= T(sgn);
          const T wl = N::sqrt_(dot3(wv3, wv3));
          const T tcl = mjclip(-an * T(50), T(-1), T(1));     // smooth axial support: full rim corner beyond ~1.1 deg of tilt
          const T sc = wl > T(1e-12) ? r / wl : T(0);
          T pp[3];
#pragma unroll
          for (int i = 0; i < 3; ++i) pp[i] = c[i] + tcl * hl * aw[i] + sc * wv3[i];
          const T sep = T(sgn) * pp[kk] - C.box_half[kk];
          if (sep > best) { best = sep; bk = kk; bs = sgn; bp[0] = pp[0]; bp[1] = pp[1]; bp[2] = pp[2]; }
        }
      if (best > T(0)) continue;
      diag.ncon += 1; diag.nbox += 1;
      if (nfound >= 2) { diag.unsupported = 1; continue; }
      // contact point midway between the surfaces, world -> body frame
      T pw[3] = {bp[0] + bcx - e.p[0], bp[1] + bcy - e.p[1], bp[2] + C.box_z[0] - e.p[2]};
      pw[bk] -= T(bs) * best * T(0.5);
      Contact<T> con;   // filled here, then copied to slot NC-2 or NC-1 with static indices (keeps register records in registers)
      con.zv[0] = con.zv[1] = con.zv[2] = T(0);
      T xb[3];          // body-frame point for the velocity of the contact; the record keeps the world-axes point pw
#pragma unroll
      for (int i = 0; i < 3; ++i) { xb[i] = k.R[i] * pw[0] + k.R[3 + i] * pw[1] + k.R[6 + i] * pw[2]; con.x[i] = pw[i]; }
      // normal from the wheel to the box is -s e_k:  +x 1, -x 2, +y 3, -y 4, -z 5 (+z: wheel below a box, never happens)
      const int code = (bk == 0) ? (bs < 0 ? 1 : 2) : ((bk == 1) ? (bs < 0 ? 3 : 4) : 5);
      fcode |= perm_pack(code) << (8 * nfound);
      const T mu = wk.mu;
      const T imp = impedance(&C.w_solimp[5 * wi], best);
      const T R0 = mjmax(N::minval, (T(1) - imp) * N::rcp_(imp) * C.w_tran[wi] * (T(1) + mu * mu));
      con.D = (best < T(0)) ? N::rcp_(T(2) * C.w_mureg2[wi] * R0) : T(0);
      T u[3], wv[2], vel[3], fn[3], ft1[3], ft2[3];
      contact_cols(w, wk, xb, u, wv);
      contact_frame(k, code, fn, ft1, ft2);
      project_point(fn, ft1, ft2, xb, u, wv, vb, e.om, w.dsp, dsteer, vel);
      con.z[0] = C.w_B[wi] * vel[0] + C.w_K[wi] * imp * best;
      con.z[1] = C.w_B[wi] * vel[1];
      con.z[2] = C.w_B[wi] * vel[2];
      if (NC > 2) { if (nfound == 0) w.con[NC > 2 ? NC - 2 : 0] = con; else w.con[NC - 1] = con; }
      ++nfound;
    }
  }


  // ---- chassis plates (convex hulls welded to the chassis): contacts of plate `p` in MuJoCo's generation order -- floor first
  // (mjc_PlaneConvex: support vertex, then its hull-graph neighbours within the margin and at least pl_tol away from the first
  // contact, <= 3), then the obstacle boxes in box order (one contact per box: the box face with the largest separation over the
  // hull's support distances, point = blend of the deepest vertices; see oracle box_convex).  Up to kMaxPlate contacts are
  // returned (body-frame point, distance, packed frame permutation | 32); the true count is the return value.
  static constexpr int kMaxPlate = 4;
  ACKB_HD static int plate_adj(const Consts<T>& C, int p, int v, int ei) {
    const unsigned wv = (unsigned)C.pl_adj[(p * 32 + v) * 6 + (ei >> 2)];
    return (int)((wv >> (6 * (ei & 3))) & 63u);
  }
  // out-of-line entry for the step kernels: state and kinematics travel BY VALUE, so that the caller's copies stay in registers (an
  // address passed to a real call would pin them in local memory for the whole substep); executed only when a plate can touch
  // something (warp-uniform test in dynamics), which is rare
  ACKB_COLD static int plate_contacts_cold(const Consts<T>& C, State e, Kin<T> k, int p, T (*xb)[3], T* dist, unsigned* code) {
    return plate_contacts(C, e, k, p, xb, dist, code);
  }
  ACKB_HD static int plate_contacts(const Consts<T>& C, const State& e, const Kin<T>& k, int p, T (*xb)[3], T* dist, unsigned* code) {
    int cnt = 0;
    const int nv = (int)C.pl_nvert[p];
    const T* V = &C.pl_vert[p * 96];
    if (C.pl_floor[0] != T(0)) {
      const T hO = e.p[2] - C.plane_z[0];
      int best = -1;
      T dbest = T(1e30);
#pragma unroll 1
      for (int v = 0; v < nv; ++v) {
        const T d = dot3(k.n, V + 3 * v) + hO;
        if (d < dbest) { dbest = d; best = v; }
      }
      if (best >= 0 && dbest <= T(0)) {
#pragma unroll
        for (int i = 0; i < 3; ++i) xb[0][i] = V[3 * best + i] - k.n[i] * dbest * T(0.5);
        dist[0] = dbest; code[0] = perm_pack(0) | 32u;
        cnt = 1;
        const T tol2 = C.pl_tol[p] * C.pl_tol[p];
#pragma unroll 1
        for (int ei = 0; ei < 24 && cnt < 3; ++ei) {
          const int v = plate_adj(C, p, best, ei);
          if (v == 63) break;
          const T dx = V[3 * v] - V[3 * best], dy = V[3 * v + 1] - V[3 * best + 1], dz = V[3 * v + 2] - V[3 * best + 2];
          if (dx * dx + dy * dy + dz * dz < tol2) continue;
          const T d = dot3(k.n, V + 3 * v) + hO;
          if (d > T(0)) continue;
#pragma unroll
          for (int i = 0; i < 3; ++i) xb[cnt][i] = V[3 * v + i] - k.n[i] * d * T(0.5);
          dist[cnt] = d; code[cnt] = perm_pack(0) | 32u;
          ++cnt;
        }
      }
    }
    if (C.pl_box[0] != T(0) && C.nbox[0] > T(0)) {
      // plate centre in the world, candidate boxes under its bounding sphere
      T cw[3];
      const T* pc = &C.pl_center[3 * p];
#pragma unroll
      for (int i = 0; i < 3; ++i) cw[i] = e.p[i] + k.R[3 * i] * pc[0] + k.R[3 * i + 1] * pc[1] + k.R[3 * i + 2] * pc[2];
      const T reach = C.pl_radius[p];
      const bool grid = C.grid_on[0] != T(0);
      int gx0 = 0, gy0 = 0, gny = 1, ncand = (int)C.nbox[0];
      if (grid) {
        const T ip = N::rcp_(C.grid_pitch[0]);
        gx0 = (int)floor((cw[0] - reach - C.grid_x0[0]) * ip); gy0 = (int)floor((cw[1] - reach - C.grid_y0[0]) * ip);
        const int gx1 = (int)floor((cw[0] + reach - C.grid_x0[0]) * ip), gy1 = (int)floor((cw[1] + reach - C.grid_y0[0]) * ip);
        gny = gy1 - gy0 + 1;
        ncand = (gx1 - gx0 + 1) * gny;
        if (ncand > 16) ncand = 16;
      }
#pragma unroll 1
      for (int bi = 0; bi < ncand; ++bi) {
        T bcx, bcy;
        if (grid) {
          const int ix = gx0 + bi / gny, iy = gy0 + bi % gny;
          if (ix < 0 || iy < 0 || ix >= (int)C.grid_nx[0] || iy >= (int)C.grid_ny[0]) continue;
          if (!((tbl_grid_row(C, iy) >> ix) & 1u)) continue;
          bcx = C.grid_x0[0] + (T(ix) + T(0.5)) * C.grid_pitch[0]; bcy = C.grid_y0[0] + (T(iy) + T(0.5)) * C.grid_pitch[0];
        } else { bcx = C.box_cx[bi]; bcy = C.box_cy[bi]; }
        const T bc[3] = {bcx, bcy, C.box_z[0]};
        if (N::abs_(cw[0] - bc[0]) > C.box_half[0] + reach || N::abs_(cw[1] - bc[1]) > C.box_half[1] + reach ||
            N::abs_(cw[2] - bc[2]) > C.box_half[2] + reach) continue;
        // support distances of the hull over the six face normals of the box (vertices in the box frame = world axes)
        T dmin[6] = {T(1e30), T(1e30), T(1e30), T(1e30), T(1e30), T(1e30)};
        const T off[3] = {e.p[0] - bc[0], e.p[1] - bc[1], e.p[2] - bc[2]};
#pragma unroll 1
        for (int v = 0; v < nv; ++v) {
#pragma unroll
          for (int kk = 0; kk < 3; ++kk) {
            const T pk = off[kk] + k.R[3 * kk] * V[3 * v] + k.R[3 * kk + 1] * V[3 * v + 1] + k.R[3 * kk + 2] * V[3 * v + 2];
            dmin[2 * kk] = mjmin(dmin[2 * kk], -pk - C.box_half[kk]);
            dmin[2 * kk + 1] = mjmin(dmin[2 * kk + 1], pk - C.box_half[kk]);
          }
        }
        T best = T(-1e30);
        int bk = 0, bs = 1;
#pragma unroll
        for (int kk = 0; kk < 3; ++kk) {      // same visiting order as the oracle: axis, then sign -1, +1
          if (dmin[2 * kk] > best) { best = dmin[2 * kk]; bk = kk; bs = -1; }
          if (dmin[2 * kk + 1] > best) { best = dmin[2 * kk + 1]; bk = kk; bs = 1; }
        }
        if (best > T(0)) continue;
        if (cnt < kMaxPlate) {
          T wsum = T(0), pp[3] = {T(0), T(0), T(0)};
          const T iblend = T(1000);      // 1 / PLATE_BLEND (1 mm)
#pragma unroll 1
          for (int v = 0; v < nv; ++v) {
            T pv[3];
#pragma unroll
            for (int kk = 0; kk < 3; ++kk) pv[kk] = off[kk] + k.R[3 * kk] * V[3 * v] + k.R[3 * kk + 1] * V[3 * v + 1] + k.R[3 * kk + 2] * V[3 * v + 2];
            const T d = T(bs) * pv[bk] - C.box_half[bk];
            const T wgt = T(1) - (d - best) * iblend;
            if (wgt <= T(0)) continue;
            wsum += wgt;
            pp[0] += wgt * pv[0]; pp[1] += wgt * pv[1]; pp[2] += wgt * pv[2];
          }
          const T iw = N::rcp_(wsum);
          pp[0] *= iw; pp[1] *= iw; pp[2] *= iw;
          pp[bk] -= T(bs) * best * T(0.5);
          // relative to the chassis origin (world axes), then into the body frame
          const T pw[3] = {pp[0] - off[0], pp[1] - off[1], pp[2] - off[2]};
#pragma unroll
          for (int i = 0; i < 3; ++i) xb[cnt][i] = k.R[i] * pw[0] + k.R[3 + i] * pw[1] + k.R[6 + i] * pw[2];
          dist[cnt] = best;
          // frame of the normal bs * e_bk pointing from the box to the plate (oracle make_frame), acting on the chassis with a plus sign
          const unsigned packed = (bk == 0) ? (bs > 0 ? 0u : 20u) : ((bk == 1) ? (bs > 0 ? 1u : 21u) : (bs > 0 ? 18u : 6u));
          code[cnt] = packed | 32u;
        }
        ++cnt;
      }
    }
    return cnt;
  }

  // one extra (slot >= 2) contact record from a body-frame point: row parameters with the given friction / solref / solimp class
  ACKB_HD static void make_extra(const Consts<T>& C, const State& e, const Kin<T>& k, const T* vb, const WheelT& w, const WheelK<T>& wk, const T* xb, T dd,
                                 unsigned code, T dsteer, Contact<T>& con) {
    const bool chassis = (code & 32u) != 0u;
    const T mu = chassis ? C.pl_mu[0] : wk.mu;
    const T* solimp = chassis ? C.pl_solimp : &C.w_solimp[5 * (wk.hidx - 2)];
    const T tran = chassis ? C.pl_tran[(wk.hidx - 2) >> 1] : C.w_tran[wk.hidx - 2];
    const T mureg2 = chassis ? C.pl_mureg2[0] : C.w_mureg2[wk.hidx - 2];
    const T Kc = chassis ? C.pl_K[0] : C.w_K[wk.hidx - 2], Bc = chassis ? C.pl_B[0] : C.w_B[wk.hidx - 2];
    const T imp = impedance(solimp, dd);
    const T R0 = mjmax(N::minval, (T(1) - imp) * N::rcp_(imp) * tran * (T(1) + mu * mu));
    con.D = (dd < T(0)) ? N::rcp_(T(2) * mureg2 * R0) : T(0);
    T u[3], wv[2], vel[3], fn[3], ft1[3], ft2[3];
    contact_cols(w, wk, xb, u, wv);
    if (chassis) { u[0] = u[1] = u[2] = T(0); wv[0] = wv[1] = T(0); }
    // frame axes in the body frame from the packed permutation: world axis j in body coordinates is row j of R
    {
      const Perm<T> pf = perm_unpack<T>(code);
      const T* Rx = k.R; const T* Ry = k.R + 3; const T* Rz = k.R + 6;
      const T* an = pf.pat == 0 ? Rx : (pf.pat == 1 ? Ry : Rz);
      const T* a1 = pf.pat == 1 ? Rz : Ry;
      const T* a2 = pf.pat == 0 ? Rz : Rx;
#pragma unroll
      for (int i = 0; i < 3; ++i) { fn[i] = pf.sn * an[i]; ft1[i] = pf.s1 * a1[i]; ft2[i] = pf.s2 * a2[i]; }
    }
    project_point(fn, ft1, ft2, xb, u, wv, vb, e.om, chassis ? T(0) : w.dsp, chassis ? T(0) : dsteer, vel);
    con.z[0] = Bc * vel[0] + Kc * imp * dd;
    con.z[1] = Bc * vel[1];
    con.z[2] = Bc * vel[2];
    con.zv[0] = con.zv[1] = con.zv[2] = T(0);
#pragma unroll
    for (int i = 0; i < 3; ++i) con.x[i] = k.R[3 * i] * xb[0] + k.R[3 * i + 1] * xb[1] + k.R[3 * i + 2] * xb[2];
  }
  // place an extra contact into the next free extra slot of the wheel record (static slot indices keep register records in registers)
  ACKB_HD static void place_extra(WheelT& w, const Contact<T>& con, unsigned code, int& nfound, unsigned& fcode, StepDiag& diag) {
    if (NC > 2) {
      if (nfound == 0) w.con[NC > 2 ? NC - 2 : 0] = con;
      else if (nfound == 1) w.con[NC - 1] = con;
      else { diag.unsupported = 1; return; }
      fcode |= (code & 255u) << (8 * nfound);
      ++nfound;
    } else diag.unsupported = 1;
  }

  ACKB_HD static void make_shared_rows(const Consts<T>& C, const State& e, SharedRows<T>& s) {
    if (C.has_eq[0] != T(0)) {
      T pos = e.st[0] - e.st[1];
      T imp = impedance(C.eq_solimp, pos);
      T R = mjmax(N::minval, (T(1) - imp) * N::rcp_(imp) * C.eq_invweight[0]);
      s.eqD = N::rcp_(R);
      s.eq_aref = -C.eq_B[0] * (e.dst[0] - e.dst[1]) - C.eq_K[0] * imp * pos;
    } else { s.eqD = T(0); s.eq_aref = T(0); }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      s.flf[i] = C.h_floss[i]; s.flR[i] = C.h_flR[i]; s.flD[i] = C.h_flD[i];
      s.fl_aref[i] = -C.h_flB[i] * e.dst[i];
      s.limD[i] = T(0); s.lim_sign[i] = T(1); s.lim_aref[i] = T(0);
      if (C.st_limited[i] != T(0)) {
        T dlo = e.st[i] - C.st_lo[i], dhi = C.st_hi[i] - e.st[i];
        T pos = T(0), sign = T(0);
        if (dlo < T(0)) { pos = dlo; sign = T(1); }
        else if (dhi < T(0)) { pos = dhi; sign = T(-1); }
        if (sign != T(0)) {
          T imp = impedance(&C.lim_solimp[5 * i], pos);
          T R = mjmax(N::minval, (T(1) - imp) * N::rcp_(imp) * C.h_invweight[i]);
          s.limD[i] = N::rcp_(R); s.lim_sign[i] = sign;
          s.lim_aref[i] = -C.lim_B[i] * (sign * e.dst[i]) - C.lim_K[i] * imp * pos;
        }
      }
    }
  }

  // ---- actuator force on hinge h (B12); sums every actuator that targets it
  ACKB_HD static T actuator_force(const Consts<T>& C, const T* ctrl, int h, T len, T vel) {
    T f = T(0);
    const int nact = (int)C.nact[0];
#pragma unroll 1
    for (int u = 0; u < nact; ++u) {
      if ((int)C.act_hinge[u] != h) continue;
      T c = ctrl[u];
      if (C.act_ctrllimited[u] != T(0)) c = mjclip(c, C.act_ctrlrange[2 * u], C.act_ctrlrange[2 * u + 1]);
      T fu = C.act_gain[u] * c + C.act_bias[3 * u] + C.act_bias[3 * u + 1] * len + C.act_bias[3 * u + 2] * vel;
      if (C.act_forcelimited[u] != T(0)) fu = mjclip(fu, C.act_forcerange[2 * u], C.act_forcerange[2 * u + 1]);
      f += fu;
    }
    return f;
  }

  // solver variables that live across the Newton loop (kept in registers; a struct so that the loop can be instantiated twice)
  struct SolverVars {
    T a_sh[8], Ma_sh[8], x_sh[8], tau_sh[8];   // shared dofs in WORLD axes: lin3, ang3, steer L, steer R
    T mcw[3], Iw[6], ezw[3];                    // m c, inertia about the chassis origin (xx yy zz xy xz yz), steer axis: world axes
    SharedRows<T> sr;
    int iter, nls, phase;
    bool first;
    T lam2;
    unsigned szone0;
  };

  // shared block of M~ in world axes (packed lower triangle)
  ACKB_HD static void shared_mass(const Consts<T>& C, const SolverVars& sv, T* S) {
#pragma unroll
    for (int i = 0; i < 36; ++i) S[i] = T(0);
    const T m = C.mass[0], cx = sv.mcw[0], cy = sv.mcw[1], cz = sv.mcw[2];
    S[tri(0, 0)] = S[tri(1, 1)] = S[tri(2, 2)] = m;
    // (ang, lin) block = [m c]x
    S[tri(3, 1)] = -cz; S[tri(3, 2)] = cy;
    S[tri(4, 0)] = cz;  S[tri(4, 2)] = -cx;
    S[tri(5, 0)] = -cy; S[tri(5, 1)] = cx;
    const T* I = sv.Iw;
    S[tri(3, 3)] = I[0]; S[tri(4, 4)] = I[1]; S[tri(5, 5)] = I[2];
    S[tri(4, 3)] = I[3]; S[tri(5, 3)] = I[4]; S[tri(5, 4)] = I[5];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
#pragma unroll
      for (int j = 0; j < 3; ++j) S[tri(6 + i, 3 + j)] = C.h_inertia[i] * sv.ezw[j];
      S[tri(6 + i, 6 + i)] = C.h_inertia[i] + C.h_armature[i];
    }
  }

  // shared part of M~ x (world axes) given the team-summed wheel term acc = sum_w J axis_w x_spin,w
  ACKB_HD static void mul_M_shared(const Consts<T>& C, const SolverVars& sv, const T* x_sh, const T* acc, T* y_sh) {
    T c1[3], c2[3];
    cross3(c1, sv.mcw, x_sh + 3);  // m c x a_ang
    cross3(c2, sv.mcw, x_sh);      // m c x a_lin
    const T* I = sv.Iw;
    const T* a = x_sh + 3;
    const T hs = C.h_inertia[0] * x_sh[6] + C.h_inertia[1] * x_sh[7];
    const T eza = dot3(sv.ezw, a);
#pragma unroll
    for (int i = 0; i < 3; ++i) y_sh[i] = C.mass[0] * x_sh[i] - c1[i];
    y_sh[3] = c2[0] + I[0] * a[0] + I[3] * a[1] + I[4] * a[2] + acc[0] + sv.ezw[0] * hs;
    y_sh[4] = c2[1] + I[3] * a[0] + I[1] * a[1] + I[5] * a[2] + acc[1] + sv.ezw[1] * hs;
    y_sh[5] = c2[2] + I[4] * a[0] + I[5] * a[1] + I[2] * a[2] + acc[2] + sv.ezw[2] * hs;
#pragma unroll
    for (int i = 0; i < 2; ++i) y_sh[6 + i] = C.h_inertia[i] * eza + (C.h_inertia[i] + C.h_armature[i]) * x_sh[6 + i];
  }

  // ---- the Newton loop (B13-B15), see `dynamics`.  TL = false: the layout's own team (LANES lanes, WPL wheels per lane).
  // TL = true (tail mode of the 1-lane layout): 4 lanes per environment, one wheel record each, reached through `whp`.
  // (TL is a warp-uniform run-time flag and the loop is instantiated ONCE: a second copy of the body costs more in instruction
  // fetch than tail mode saves -- measured: 134 M -> 106 M env-steps/s at 131072 envs, frame_skip 4.)
  // The shared dofs are expressed in WORLD axes (lin3, ang3, steer L, steer R): every contact frame is then a signed permutation
  // of the axes (Perm), so rows, forces and the 3 x 3 contact Hessians need no frame products.
  ACKB_HD static void solve_loop(const Consts<T>& C, WheelT* whp, const int wbase, const int ncs, SolverVars& sv,
                                 const bool exit_for_tail, const bool TL) {
    const int wpl = TL ? 1 : WPL;
    T (&a_sh)[8] = sv.a_sh; T (&Ma_sh)[8] = sv.Ma_sh; T (&x_sh)[8] = sv.x_sh; T (&tau_sh)[8] = sv.tau_sh;
    SharedRows<T>& sr = sv.sr;
    int& iter = sv.iter; int& nls = sv.nls; int& phase = sv.phase;
    bool& first = sv.first;
    T& lam2 = sv.lam2;
    unsigned& szone0 = sv.szone0;
    const T tol = mjmax(C.tolerance[0], N::tol_floor);
    const int maxit = (int)C.iterations[0], maxls_exact = (int)C.ls_iterations[0];
    const int ls_fast_cap = (int)C.ls_fast_cap[0], ls_fast_iters = (int)C.ls_fast_iters[0];
    const int ls_mid_cap = (int)C.ls_mid_cap[0], ls_mid_iters = (int)C.ls_mid_iters[0];
    while (Tm::loop_any(phase != 2)) {
#if defined(__CUDA_ARCH__)
      // plain 1-lane pass: leave as soon as few enough environments of the warp are still iterating (the caller re-spreads them)
      if (!TL && exit_for_tail && __popc(__ballot_sync(0xffffffffu, phase != 2)) <= kTailEnvs) return;
#endif
      const bool stepping = (phase == 0);
      // ---- move along x: exact line search (safeguarded Newton on f'(alpha)), then update the point
      {
        // M~ x (shared part), contact images of x, line-search constants
        T acc[5] = {T(0), T(0), T(0), T(0), T(0)};  // sum_w J a_w x_w (3), private parts of x.Mx and x.(Ma - tau)
#pragma unroll 1
        for (int s = 0; s < wpl; ++s) {
          const int wi = wbase + s;
          const WheelK<T> wk = wheel_consts(C, wi);
          WheelT& w = whp[s];
          const T aw_aang_x = dot3(w.axw, x_sh + 3), aw_aang_a = dot3(w.axw, a_sh + 3);
          const T Mv_sp = wk.J * aw_aang_x + wk.cdiag * w.x, Ma_sp = wk.J * aw_aang_a + wk.cdiag * w.a;
          const T pwt = pair_weight();
          const T jx = pwt * wk.J * w.x;
          acc[0] += jx * w.axw[0]; acc[1] += jx * w.axw[1]; acc[2] += jx * w.axw[2];
          acc[3] += pwt * w.x * Mv_sp; acc[4] += pwt * w.x * (Ma_sp - w.tau);
          const T ast = wk.isL * x_sh[6] + wk.isR * x_sh[7];
          const T stw = wk.isL + wk.isR;
          ACKB_CONTACTS_BEGIN(c)
            Contact<T>& con = w.con[c];
            T u[3], wv[3], y[3];
            const unsigned code = slot_code(w, c);
            const T cpl = slot_coupling(code);      // 0 for a chassis-plate contact (no spin / steer column), 1 for a wheel contact
            contact_cols_w(w, stw * cpl, sv.ezw, con.x, u, wv);
            point_accel_w(con.x, u, wv, x_sh, x_sh + 3, w.x * cpl, ast, y);
            perm_apply(perm_unpack<T>(code), y, con.zv);
          ACKB_CONTACTS_END
        }
        team_sum_n(acc, TL);
        T Mv_sh[8];
        mul_M_shared(C, sv, x_sh, acc, Mv_sh);
        T sMs = acc[3], sg = acc[4];
#pragma unroll
        for (int i = 0; i < 8; ++i) { sMs += x_sh[i] * Mv_sh[i]; sg += x_sh[i] * (Ma_sh[i] - tau_sh[i]); }
        T alpha = T(1);
        bool exact = false;
        bool ls_on = stepping && !first;
        T lo = T(0), hi = T(-1);
        // The search is exact (safeguarded Newton to ls_rel) only for environments that are slow to converge; during their
        // first ls_fast_iters iterations at most ls_fast_cap evaluations are spent and the next candidate is taken, which keeps
        // the warp-wide trip count of this loop small (the outer Newton iteration absorbs the inexactness).
        const int my_maxls = (iter < ls_fast_iters) ? ls_fast_cap : ((iter < ls_mid_iters) ? ls_mid_cap : maxls_exact);
        for (int ls = 0; ls < maxls_exact && Tm::any(ls_on); ++ls) {
          T d[3] = {T(0), T(0), T(0)};
#pragma unroll 1
          for (int s = 0; s < wpl; ++s) {
            const int wi = wbase + s;
            const WheelK<T> wk = wheel_consts(C, wi);
            const WheelT& w = whp[s];
            T f, q;
            floss_row(w.a + alpha * w.x + wk.flB * w.dsp, wk.flf, wk.flR, wk.flD, &f, &q);
            d[0] -= pair_weight() * f * w.x;
            d[1] += pair_weight() * q * wk.flD * w.x * w.x;
            unsigned zone = (q != T(0)) ? 1u : (f > T(0) ? 0u : 2u);
            ACKB_CONTACTS_BEGIN(c)
              const Contact<T>& con = w.con[c];
              const T mu = slot_mu(C, wk, slot_code(w, c));
              // rows x_r = z_n +- mu z_t at the trial point and their slopes j_r (select form, no divergent branches)
              const T jn = con.zv[0], j1 = mu * con.zv[1], j2 = mu * con.zv[2];
              const T xn = con.z[0] + alpha * jn, x1 = mu * con.z[1] + alpha * j1, x2 = mu * con.z[2] + alpha * j2;
              const T jr[4] = {jn + j1, jn - j1, jn + j2, jn - j2};
              const T xr[4] = {xn + x1, xn - x1, xn + x2, xn - x2};
              T s0 = T(0), s1 = T(0);
              unsigned zb = 0u;
#pragma unroll
              for (int r = 0; r < 4; ++r) {
                const bool act = xr[r] < T(0);
                const T jm = act ? jr[r] : T(0);      // one select, then two fused multiply-adds
                s0 += jm * xr[r];
                s1 += jm * jr[r];
                zb |= act ? (1u << r) : 0u;
              }
              d[0] += con.D * s0; d[1] += con.D * s1;       // D = 0 for an absent or excluded contact
              zone = (zone << 4) | ((con.D > T(0)) ? zb : 0u);
            ACKB_CONTACTS_END
            d[2] += (zone != w.zone0) ? T(1) : T(0);
          }
          team_sum_n(d, TL);
          T gL, gR, kLL, kLR, kRR;
          shared_rows_eval(sr, a_sh[6] + alpha * x_sh[6], a_sh[7] + alpha * x_sh[7], &gL, &gR, &kLL, &kLR, &kRR);
          const T d1 = alpha * sMs + sg + d[0] - gL * x_sh[6] - gR * x_sh[7];
          const T d2 = sMs + d[1] + kLL * x_sh[6] * x_sh[6] + T(2) * kLR * x_sh[6] * x_sh[7] + kRR * x_sh[7] * x_sh[7];
          if (ls_on) {
            ++nls;
            // a full Newton step that keeps every row in its zone lands on the exact minimiser
            if (ls == 0 && d[2] == T(0) && shared_rows_zone(sr, a_sh[6] + x_sh[6], a_sh[7] + x_sh[7]) == szone0) { exact = true; ls_on = false; }
            else if (N::abs_(d1) <= N::ls_rel * lam2) ls_on = false;
            else {
              if (d1 < T(0)) lo = alpha; else hi = alpha;
              T an = alpha - d1 * N::rcp_(d2);
              if (hi >= T(0) && (an <= lo || an >= hi)) an = T(0.5) * (lo + hi);
              if (an == alpha || ls + 1 >= my_maxls) ls_on = false;
              alpha = an;
            }
          }
        }
        if (stepping) {
          if (!first) ++iter;
#pragma unroll
          for (int i = 0; i < 8; ++i) { a_sh[i] += alpha * x_sh[i]; Ma_sh[i] += alpha * Mv_sh[i]; }
#pragma unroll 1
          for (int s = 0; s < wpl; ++s) {
            WheelT& w = whp[s];
            w.a += alpha * w.x;
            ACKB_CONTACTS_BEGIN(c)
#pragma unroll
              for (int i = 0; i < 3; ++i) w.con[c].z[i] += alpha * w.con[c].zv[i];
            ACKB_CONTACTS_END
          }
          first = false;
          if (exact || iter >= maxit) phase = 2;
        }
      }
#if !defined(ACKB_SYNC_PASS)
      if (!Tm::any(phase != 2)) break;   // every environment of the warp landed on its exact minimiser
#endif

      // ---- assemble the arrow system at the current point; per-lane parts: S(36), reduced rhs(8), sum gsp^2/c (1)
      T part[45];
#pragma unroll
      for (int i = 0; i < 45; ++i) part[i] = T(0);
#pragma unroll 1
      for (int s = 0; s < wpl; ++s) {
        const int wi = wbase + s;
        const WheelK<T> wk = wheel_consts(C, wi);
        WheelT& w = whp[s];
        T f, q;
        floss_row(w.a + wk.flB * w.dsp, wk.flf, wk.flR, wk.flD, &f, &q);
        unsigned zone = (q != T(0)) ? 1u : (f > T(0) ? 0u : 2u);
        T gs_sp = T(0), cs = T(0);   // contact parts first (pair-summed with 8 lanes), wheel parts added below
        T Hll[6] = {T(0), T(0), T(0), T(0), T(0), T(0)};   // 00 10 11 20 21 22
        T Hal[3][3] = {{T(0), T(0), T(0)}, {T(0), T(0), T(0)}, {T(0), T(0), T(0)}};
        T Haa[6] = {T(0), T(0), T(0), T(0), T(0), T(0)};
        T Hsl[3] = {T(0), T(0), T(0)}, Hsa[3] = {T(0), T(0), T(0)}, Hss = T(0);
        T bl[3] = {T(0), T(0), T(0)}, ba[3] = {T(0), T(0), T(0)}, bs = T(0);
        T gl[3] = {T(0), T(0), T(0)}, ga[3] = {T(0), T(0), T(0)}, gst = T(0);
        ACKB_CONTACTS_BEGIN(c)
          const Contact<T>& con = w.con[c];
          T phi[3], qq[4];
          const unsigned code = slot_code(w, c);
          const T mu = slot_mu(C, wk, code);
          const T cpl = slot_coupling(code);
          const Perm<T> pf = perm_unpack<T>(code);
          pyramid_rows(con.D, mu, con.z, phi, qq);
          {
            const unsigned zb = (qq[0] != T(0) ? 1u : 0u) | (qq[1] != T(0) ? 2u : 0u) | (qq[2] != T(0) ? 4u : 0u) | (qq[3] != T(0) ? 8u : 0u);
            zone = (zone << 4) | ((con.D > T(0)) ? zb : 0u);
          }
          // S3 = D F^T W F with W the active-row weights on (n, t1, t2); F is a signed permutation of the world axes
          const T W00 = qq[0] + qq[1] + qq[2] + qq[3], W01 = mu * (qq[0] - qq[1]), W02 = mu * (qq[2] - qq[3]);
          const T W11 = mu * mu * (qq[0] + qq[1]), W22 = mu * mu * (qq[2] + qq[3]);
          T S3[3][3];
          perm_sym(pf, con.D * W00, con.D * W11, con.D * W22, con.D * W01, con.D * W02, S3);
          const T X[3] = {con.x[0], con.x[1], con.x[2]};
          T u[3], wv[3];
          contact_cols_w(w, (wk.isL + wk.isR) * cpl, sv.ezw, X, u, wv);
#pragma unroll
          for (int i = 0; i < 3; ++i) u[i] *= cpl;
          Hll[0] += S3[0][0]; Hll[1] += S3[1][0]; Hll[2] += S3[1][1]; Hll[3] += S3[2][0]; Hll[4] += S3[2][1]; Hll[5] += S3[2][2];
          // (ang, lin) block: column j = X x S3[:, j];  (ang, ang) block: column j = X x (S3 g_j), g_j = e_j x X
          T Cj[3][3];
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            const T v[3] = {S3[0][j], S3[1][j], S3[2][j]};
            cross3(Cj[j], X, v);
#pragma unroll
            for (int i = 0; i < 3; ++i) Hal[i][j] += Cj[j][i];
          }
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            const T Sg[3] = {Cj[0][j], Cj[1][j], Cj[2][j]};  // S3 g_j
            T col[3];
            cross3(col, X, Sg);
            if (j == 0) { Haa[0] += col[0]; Haa[1] += col[1]; Haa[3] += col[2]; }
            if (j == 1) { Haa[2] += col[1]; Haa[4] += col[2]; }
            if (j == 2) { Haa[5] += col[2]; }
          }
          // spin column u and steer column wv
          T Su[3], Sw[3], XSu[3], XSw[3];
#pragma unroll
          for (int i = 0; i < 3; ++i) { Su[i] = S3[i][0] * u[0] + S3[i][1] * u[1] + S3[i][2] * u[2]; Sw[i] = S3[i][0] * wv[0] + S3[i][1] * wv[1] + S3[i][2] * wv[2]; }
          cross3(XSu, X, Su);
          cross3(XSw, X, Sw);
#pragma unroll
          for (int i = 0; i < 3; ++i) { bl[i] += Su[i]; ba[i] += XSu[i]; Hsl[i] += Sw[i]; Hsa[i] += XSw[i]; }
          cs += dot3(u, Su);
          bs += dot3(wv, Su);
          Hss += dot3(wv, Sw);
          // body-frame contact force and its generalised image (enters the gradient with a minus sign)
          T Phi[3], XF[3];
          perm_apply_t(pf, phi, Phi);
          cross3(XF, X, Phi);
#pragma unroll
          for (int i = 0; i < 3; ++i) { gl[i] -= Phi[i]; ga[i] -= XF[i]; }
          gst -= dot3(wv, Phi);
          gs_sp -= dot3(u, Phi);
        ACKB_CONTACTS_END
        if (PAIR) {
          gs_sp = Tm::pair_sum(gs_sp); cs = Tm::pair_sum(cs); bs = Tm::pair_sum(bs);
#pragma unroll
          for (int i = 0; i < 3; ++i) { bl[i] = Tm::pair_sum(bl[i]); ba[i] = Tm::pair_sum(ba[i]); }
        }
        gs_sp += wk.J * dot3(w.axw, a_sh + 3) + wk.cdiag * w.a - w.tau - f;
        cs += wk.cdiag + q * wk.flD;
        ba[0] += wk.J * w.axw[0]; ba[1] += wk.J * w.axw[1]; ba[2] += wk.J * w.axw[2];
        if (phase == 0) w.zone0 = zone;
        w.g = gs_sp; w.cw = cs;
        T b8[8], gsh_w[8];
#pragma unroll
        for (int i = 0; i < 3; ++i) { b8[i] = bl[i]; b8[3 + i] = ba[i]; w.b[i] = bl[i]; w.b[3 + i] = ba[i]; gsh_w[i] = gl[i]; gsh_w[3 + i] = ga[i]; }
        w.b[6] = bs;
        b8[6] = wk.isL * bs; b8[7] = wk.isR * bs; gsh_w[6] = wk.isL * gst; gsh_w[7] = wk.isR * gst;
        part[tri(0, 0)] += Hll[0]; part[tri(1, 0)] += Hll[1]; part[tri(1, 1)] += Hll[2];
        part[tri(2, 0)] += Hll[3]; part[tri(2, 1)] += Hll[4]; part[tri(2, 2)] += Hll[5];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int j = 0; j < 3; ++j) part[tri(3 + i, j)] += Hal[i][j];
        part[tri(3, 3)] += Haa[0]; part[tri(4, 3)] += Haa[1]; part[tri(4, 4)] += Haa[2];
        part[tri(5, 3)] += Haa[3]; part[tri(5, 4)] += Haa[4]; part[tri(5, 5)] += Haa[5];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          part[tri(6, j)] += wk.isL * Hsl[j]; part[tri(6, 3 + j)] += wk.isL * Hsa[j];
          part[tri(7, j)] += wk.isR * Hsl[j]; part[tri(7, 3 + j)] += wk.isR * Hsa[j];
        }
        part[tri(6, 6)] += wk.isL * Hss; part[tri(7, 7)] += wk.isR * Hss;
        const T ci = pair_weight() * N::rcp_(cs);     // the Schur terms of a wheel are counted once per team
#pragma unroll
        for (int a = 0; a < 8; ++a) {
          const T bc = b8[a] * ci;
#pragma unroll
          for (int bb = 0; bb <= a; ++bb) part[tri(a, bb)] -= bc * b8[bb];
          part[36 + a] += gsh_w[a] - bc * gs_sp;
        }
        part[44] += gs_sp * gs_sp * ci;
      }
      team_sum_part(part, TL);
      T S[36];
      shared_mass(C, sv, S);
#pragma unroll
      for (int i = 0; i < 36; ++i) S[i] += part[i];
      T fL, fR, hLL, hLR, hRR;
      shared_rows_eval(sr, a_sh[6], a_sh[7], &fL, &fR, &hLL, &hLR, &hRR);
      S[tri(6, 6)] += hLL; S[tri(7, 6)] += hLR; S[tri(7, 7)] += hRR;
      T rhs[8], y_sh[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) rhs[i] = Ma_sh[i] - tau_sh[i] + part[36 + i];
      rhs[6] -= fL; rhs[7] -= fR;
#pragma unroll
      for (int i = 0; i < 8; ++i) y_sh[i] = rhs[i];
      ldl8_factor(S);
      ldl8_solve(S, y_sh);
      if (phase != 2) {   // converged environments keep their solution
        szone0 = shared_rows_zone(sr, a_sh[6], a_sh[7]);
        lam2 = part[44];  // Newton decrement g^T H^-1 g
#pragma unroll
        for (int i = 0; i < 8; ++i) lam2 += rhs[i] * y_sh[i];
        // x = -H^-1 g
#pragma unroll 1
        for (int s = 0; s < wpl; ++s) {
          const int wi = wbase + s;
          WheelT& w = whp[s];
          T dotb = T(0);
#pragma unroll
          for (int a = 0; a < 6; ++a) dotb += w.b[a] * y_sh[a];
          dotb += w.b[6] * ((wi == 2) ? y_sh[6] : ((wi == 3) ? y_sh[7] : T(0)));
          w.x = -(w.g - dotb) * N::rcp_(w.cw);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) x_sh[i] = -y_sh[i];
        // predicted improvement of a full Newton step is lam2/2: below tolerance the point is converged
        if (!(C.solver_scale[0] * T(0.5) * lam2 >= tol)) phase = 2;
      }
    }
  }

  // ---- one physics substep (mj_step): everything between kinematics and integration.
  // `k` must hold the kinematics of the current state; `wh` are the WPL wheel records of this lane.
  //
  // Solver structure (B13-B16): ONE loop body.  Every trip moves along the current direction x with an exact line
  // search and then assembles the arrow-shaped Newton system  H x = -g  (H = M~ + J^T D J, g = M~ a - tau - J^T f),
  // LDL^T-factorises its 8 x 8 Schur complement and solves it.  The iteration starts at a = 0 with a unit step along
  // the warm start (previous qacc), as MuJoCo does when its cost beats qacc_smooth; the minimiser is unique, so the
  // starting point only affects the iteration count (qacc_smooth is never needed).
  ACKB_HD static void dynamics(const Consts<T>& C, State& e, const Kin<T>& k, const T* ctrl, int lane, WheelT* wh, StepDiag& diag,
                               DebugTap<T>* tap, int rec_stride = 0) {
    const T h = C.timestep[0];
    T vb[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) vb[i] = k.R[i] * e.vw[0] + k.R[3 + i] * e.vw[1] + k.R[6 + i] * e.vw[2];

    // ---- wheels: collision, spin-dof smooth force; wheel-derived parts of the chassis / steer bias (closed-form RNE)
    T bpart[8];   // [hrel(3), gyro(3), steerL, steerR]
#pragma unroll
    for (int i = 0; i < 8; ++i) bpart[i] = T(0);
    bool warm_ok = true;
    // can a chassis plate touch anything in this pose?  floor: a bounding point of the plates at or below the plane; boxes: the
    // chassis within reach of an occupied cell (checked per plate inside plate_contacts).  Warp-uniform (the exact routines diverge badly).
    bool plates_near = false;
    if (NC > 2 && C.pl_count[0] != T(0)) {
      bool near = C.pl_box[0] != T(0);
      const int nh = (int)C.nhull[0];
      const T hO = e.p[2] - C.plane_z[0];
#pragma unroll 1
      for (int i = 0; i < nh; ++i) near = near || (dot3(k.n, &C.hull_pts[3 * i]) + hO <= T(0));
      plates_near = Tm::any(near);
    }
#pragma unroll 1
    for (int s = 0; s < WPL; ++s) {
      const int wi = wheel_index(lane, s);
      const WheelK<T> wk = wheel_consts(C, wi);
      WheelT& w = wh[s];
      T tri[8];
      tri[7] = T(0);
      collide_wheel(C, e, k, vb, wi, lane & 1, wk, w, diag, NC > 2 ? tri : nullptr);
      if (NC > 2) {
        // extra contact slots of this wheel (NC - 2, NC - 1): wheel-vs-box contacts, the cap-down triangle points of the wheel, and this
        // wheel's share of the chassis-plate contacts (plate wi / 2: entries 0, 1 of its list go to wheel 2p, entries 2, 3 to wheel 2p + 1)
        int nfound = 0;
        unsigned fcode = 0u;
        w.con[NC - 2].D = T(0); w.con[NC - 1].D = T(0);
        for (int i = 0; i < 3; ++i) { w.con[NC - 2].x[i] = w.con[NC - 1].x[i] = T(0); w.con[NC - 2].z[i] = w.con[NC - 1].z[i] = T(0); }
        collide_boxes(C, e, k, vb, wi, wk, w, diag, nfound, fcode);
        const T dst_w = wk.isL * e.dst[0] + wk.isR * e.dst[1];
        if (tri[7] != T(0)) {
          diag.ncon += 2;
          Contact<T> con;
          make_extra(C, e, k, vb, w, wk, tri, tri[6], perm_pack(0), dst_w, con);
          place_extra(w, con, perm_pack(0), nfound, fcode, diag);
          make_extra(C, e, k, vb, w, wk, tri + 3, tri[6], perm_pack(0), dst_w, con);
          place_extra(w, con, perm_pack(0), nfound, fcode, diag);
        }
        if (C.pl_count[0] != T(0) && plates_near) {
          T pxb[kMaxPlate][3], pdist[kMaxPlate];
          unsigned pcode[kMaxPlate];
          const int np = plate_contacts_cold(C, e, k, wi >> 1, pxb, pdist, pcode);
          const int part = wi & 1;
          for (int j = 0; j < 2; ++j) {
            const int idx = 2 * part + j;
            if (idx < np) {
              diag.ncon += 1;
              Contact<T> con;
              make_extra(C, e, k, vb, w, wk, pxb[idx], pdist[idx], pcode[idx], dst_w, con);
              place_extra(w, con, pcode[idx], nfound, fcode, diag);
            }
          }
          if (part == 0 && np > kMaxPlate) { diag.ncon += np - kMaxPlate; diag.unsupported = 1; }
        }
        w.fcode = fcode;
      }
#pragma unroll
      for (int i = 0; i < 3; ++i) {   // spin axis and wheel centre in world axes (solver)
        w.axw[i] = k.R[3 * i] * w.ax + k.R[3 * i + 1] * w.ay;
        w.cww[i] = k.R[3 * i] * wk.c[0] + k.R[3 * i + 1] * wk.c[1] + k.R[3 * i + 2] * wk.c[2];
      }
      const T dsteer = wk.isL * e.dst[0] + wk.isR * e.dst[1];
      const T aw[3] = {w.ax, w.ay, T(0)};
      const T ez[3] = {T(0), T(0), T(1)};
      T ezxa[3], omxa[3], omxez[3];
      cross3(ezxa, ez, aw);
      cross3(omxa, e.om, aw);
      cross3(omxez, e.om, ez);
      const T pwt = pair_weight();
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        bpart[i] += pwt * wk.J * w.dsp * aw[i];
        bpart[3 + i] += pwt * wk.J * w.dsp * dsteer * ezxa[i];
      }
      const T bias_steer = pwt * wk.J * w.dsp * omxa[2];  // ez . (om x a)
      bpart[6] += wk.isL * bias_steer;
      bpart[7] += wk.isR * bias_steer;
      const T bias_spin = wk.J * dsteer * dot3(aw, omxez);
      w.tau = -wk.damp * w.dsp - bias_spin + actuator_force(C, ctrl, wk.hidx, w.sp, w.dsp);
      w.a = T(0);
      w.x = w.warm;
      w.zone0 = 0u;
      if (NC <= 2) w.fcode = 0u;
      warm_ok = warm_ok && (N::abs_(w.warm) <= T(1e10));
    }
    Tm::sum_n(bpart);
    // number of contact slots the solver has to look at (warp-uniform): the box slots only if some environment of the warp uses them
    int ncs = NC;
    if (NC > 2) {
      bool has_box = false;
#pragma unroll 1
      for (int s = 0; s < WPL; ++s) has_box = has_box || (wh[s].con[NC - 2].D > T(0)) || (wh[s].con[NC - 1].D > T(0));
      ncs = Tm::any(has_box) ? NC : 2;
    }
    // plate hull vs floor in a kernel without extra contact slots (NC = 2): flagged; the launcher routes tilted environments to the
    // NC = 4 kernel before this can happen (ackb_kernels.cu, regime split), so the flag only fires if that guard was outrun
    if (NC <= 2 && lane == 0) {
      const int nh = (int)C.nhull[0];
      const T hO = e.p[2] - C.plane_z[0];
#pragma unroll 1
      for (int i = 0; i < nh; ++i) if (dot3(k.n, &C.hull_pts[3 * i]) + hO <= T(0)) diag.unsupported = 1;
    }
    SolverVars sv;
    SharedRows<T>& sr = sv.sr;
    make_shared_rows(C, e, sr);

    // ---- B10/B12/B13 shared smooth force  tau = passive - bias + actuation  (closed-form RNE, see header)
    T (&tau_sh)[8] = sv.tau_sh;
    {
      T gb[3];  // gravity in the body frame
#pragma unroll
      for (int i = 0; i < 3; ++i) gb[i] = k.R[i] * C.gravity[0] + k.R[3 + i] * C.gravity[1] + k.R[6 + i] * C.gravity[2];
      T Iw[3], hh[3], t0[3], t1[3], t2[3], t3[3];
      const T* I = C.inertiaO;
      Iw[0] = I[0] * e.om[0] + I[3] * e.om[1] + I[4] * e.om[2];
      Iw[1] = I[3] * e.om[0] + I[1] * e.om[1] + I[5] * e.om[2];
      Iw[2] = I[4] * e.om[0] + I[5] * e.om[1] + I[2] * e.om[2];
#pragma unroll
      for (int i = 0; i < 3; ++i) hh[i] = Iw[i] + bpart[i];
      hh[2] += C.h_inertia[0] * e.dst[0] + C.h_inertia[1] * e.dst[1];
      cross3(t0, e.om, hh);                // om x (I_O om + h_rel)
      cross3(t1, C.mcom, gb);              // m c x g
      cross3(t2, e.om, C.mcom);
      cross3(t3, e.om, t2);                // om x (om x m c)
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        tau_sh[i] = -(t3[i] - C.mass[0] * gb[i]);
        tau_sh[3 + i] = -(t0[i] + bpart[3 + i] - t1[i]);
      }
#pragma unroll
      for (int i = 0; i < 2; ++i)   // unrolled: a rolled loop would index tau_sh / bpart / e.st dynamically and push them to local memory
        tau_sh[6 + i] = -C.h_damping[i] * e.dst[i] - bpart[6 + i] + actuator_force(C, ctrl, i, e.st[i], e.dst[i]);
    }
    if (tap) for (int i = 0; i < 8; ++i) tap->tau[i] = tau_sh[i];   // body-frame smooth force (MuJoCo's qfrc_smooth up to the frame)
    // ---- world-axes quantities of this substep for the solver: tau, m c, I_O, steer axis
    {
      const T tl[3] = {tau_sh[0], tau_sh[1], tau_sh[2]}, ta[3] = {tau_sh[3], tau_sh[4], tau_sh[5]};
      T RI[9];   // R I_O
      const T* I = C.inertiaO;
      const T Im[9] = {I[0], I[3], I[4], I[3], I[1], I[5], I[4], I[5], I[2]};
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        tau_sh[i] = k.R[3 * i] * tl[0] + k.R[3 * i + 1] * tl[1] + k.R[3 * i + 2] * tl[2];
        tau_sh[3 + i] = k.R[3 * i] * ta[0] + k.R[3 * i + 1] * ta[1] + k.R[3 * i + 2] * ta[2];
        sv.mcw[i] = k.R[3 * i] * C.mcom[0] + k.R[3 * i + 1] * C.mcom[1] + k.R[3 * i + 2] * C.mcom[2];
        sv.ezw[i] = k.R[3 * i + 2];
#pragma unroll
        for (int j = 0; j < 3; ++j) RI[3 * i + j] = k.R[3 * i] * Im[j] + k.R[3 * i + 1] * Im[3 + j] + k.R[3 * i + 2] * Im[6 + j];
      }
      // I_w = (R I_O) R^T, symmetric: xx yy zz xy xz yz
      sv.Iw[0] = RI[0] * k.R[0] + RI[1] * k.R[1] + RI[2] * k.R[2];
      sv.Iw[1] = RI[3] * k.R[3] + RI[4] * k.R[4] + RI[5] * k.R[5];
      sv.Iw[2] = RI[6] * k.R[6] + RI[7] * k.R[7] + RI[8] * k.R[8];
      sv.Iw[3] = RI[0] * k.R[3] + RI[1] * k.R[4] + RI[2] * k.R[5];
      sv.Iw[4] = RI[0] * k.R[6] + RI[1] * k.R[7] + RI[2] * k.R[8];
      sv.Iw[5] = RI[3] * k.R[6] + RI[4] * k.R[7] + RI[5] * k.R[8];
    }

    // current point a (starts at 0 and takes a unit step along the warm start), M~ a
    T (&a_sh)[8] = sv.a_sh; T (&Ma_sh)[8] = sv.Ma_sh; T (&x_sh)[8] = sv.x_sh;
#pragma unroll
    for (int i = 0; i < 3; ++i) {   // MuJoCo keeps the linear part in world axes and the angular part in body axes
      x_sh[i] = e.warm_l[i];
      x_sh[3 + i] = k.R[3 * i] * e.warm_a[0] + k.R[3 * i + 1] * e.warm_a[1] + k.R[3 * i + 2] * e.warm_a[2];
    }
    x_sh[6] = e.warm_st[0]; x_sh[7] = e.warm_st[1];
#pragma unroll
    for (int i = 0; i < 8; ++i) { warm_ok = warm_ok && (N::abs_(x_sh[i]) <= T(1e10)); a_sh[i] = T(0); Ma_sh[i] = T(0); }
    if (Tm::sum(warm_ok ? 0 : 1) != 0) {   // unusable warm start: begin at a = 0
#pragma unroll
      for (int i = 0; i < 8; ++i) x_sh[i] = T(0);
#pragma unroll 1
      for (int s = 0; s < WPL; ++s) wh[s].x = T(0);
    }

    sv.iter = 0; sv.nls = 0;
    // per-environment phase: 0 = iterating (step along x, then Newton pass), 2 = converged.
    // The loop itself is warp-uniform: it runs until every environment of the warp has converged.
    sv.phase = 0; sv.first = true; sv.lam2 = T(0); sv.szone0 = 0u;
#if defined(__CUDA_ARCH__)
    constexpr bool TAIL = (LANES == 1) && (sizeof(T) == 4) && (ACKB_TAIL_MODE != 0);   // fp64 kernels are at the register cap already
    const bool tail_ok = TAIL && rec_stride != 0;   // needs the records in shared memory (rec_stride = per-thread stride in units of T)
#else
    const bool tail_ok = false;
#endif
    // pass 0: the layout's own team.  pass 1 (tail mode, 1-lane layout with records in shared memory only): once at most kTailEnvs
    // environments of the warp are still iterating, pass 0 returns and they are re-spread over the warp with 4 lanes each
    // (lane = wheel, as in the 4-lane layout): the solver state of environment number r moves to lanes 4r .. 4r+3 by shuffles,
    // the wheel records are reached in place.
    WheelT* whp = wh;
    int wb = wheel_index(lane, 0);
    bool tail = false;
#if defined(__CUDA_ARCH__)
    T a_keep[8];
    int iter_keep = 0, my_src = 0;
    bool my_active = false;
#endif
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
      solve_loop(C, whp, wb, ncs, sv, tail_ok && pass == 0, tail);
#if defined(__CUDA_ARCH__)
      if (!TAIL) break;
      const unsigned act = __ballot_sync(0xffffffffu, sv.phase != 2);
      if (pass == 1 || !tail_ok || act == 0u) break;   // warp-uniform
      const int cnt = __popc(act);
      const int lid = (int)(threadIdx.x & 31u);
      my_active = (sv.phase != 2);
      my_src = my_active ? 4 * __popc(act & ((1u << lid) - 1u)) : lid;
#pragma unroll
      for (int i = 0; i < 8; ++i) a_keep[i] = sv.a_sh[i];
      iter_keep = sv.iter;
      const int slot = lid >> 2;
      const bool serve = slot < cnt;
      const int owner = serve ? (int)__fns(act, 0u, slot + 1) : (__ffs((int)~act) - 1);   // idle lanes shadow a finished environment
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        sv.a_sh[i] = __shfl_sync(0xffffffffu, sv.a_sh[i], owner); sv.Ma_sh[i] = __shfl_sync(0xffffffffu, sv.Ma_sh[i], owner);
        sv.x_sh[i] = __shfl_sync(0xffffffffu, sv.x_sh[i], owner); sv.tau_sh[i] = __shfl_sync(0xffffffffu, sv.tau_sh[i], owner);
      }
#pragma unroll
      for (int i = 0; i < 3; ++i) { sv.mcw[i] = __shfl_sync(0xffffffffu, sv.mcw[i], owner); sv.ezw[i] = __shfl_sync(0xffffffffu, sv.ezw[i], owner); }
#pragma unroll
      for (int i = 0; i < 6; ++i) sv.Iw[i] = __shfl_sync(0xffffffffu, sv.Iw[i], owner);
      SharedRows<T>& sr2 = sv.sr;
      sr2.eqD = __shfl_sync(0xffffffffu, sr2.eqD, owner); sr2.eq_aref = __shfl_sync(0xffffffffu, sr2.eq_aref, owner);
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        sr2.flD[i] = __shfl_sync(0xffffffffu, sr2.flD[i], owner); sr2.flR[i] = __shfl_sync(0xffffffffu, sr2.flR[i], owner);
        sr2.flf[i] = __shfl_sync(0xffffffffu, sr2.flf[i], owner); sr2.fl_aref[i] = __shfl_sync(0xffffffffu, sr2.fl_aref[i], owner);
        sr2.limD[i] = __shfl_sync(0xffffffffu, sr2.limD[i], owner); sr2.lim_sign[i] = __shfl_sync(0xffffffffu, sr2.lim_sign[i], owner);
        sr2.lim_aref[i] = __shfl_sync(0xffffffffu, sr2.lim_aref[i], owner);
      }
      sv.iter = __shfl_sync(0xffffffffu, sv.iter, owner);
      sv.lam2 = __shfl_sync(0xffffffffu, sv.lam2, owner);
      sv.szone0 = __shfl_sync(0xffffffffu, sv.szone0, owner);
      sv.first = __shfl_sync(0xffffffffu, sv.first ? 1 : 0, owner) != 0;
      sv.phase = serve ? 0 : 2;
      wb = lid & 3;
      whp = reinterpret_cast<WheelT*>(reinterpret_cast<T*>(wh) + (owner - lid) * rec_stride) + wb;
      tail = true;
      __syncwarp();
#else
      break;
#endif
    }
#if defined(__CUDA_ARCH__)
    if (TAIL && tail) {   // warp-uniform: hand the accelerations back to the lanes that own the environments
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 8; ++i) { const T v = __shfl_sync(0xffffffffu, sv.a_sh[i], my_src); sv.a_sh[i] = my_active ? v : a_keep[i]; }
      const int it2 = __shfl_sync(0xffffffffu, sv.iter, my_src);
      sv.iter = my_active ? it2 : iter_keep;
    }
#endif
    const int iter = sv.iter, nls = sv.nls;
    {   // back to the chassis frame for the rest of the step: a~ = [R^T a_lin ; R^T a_ang ; hinges]
      const T al[3] = {a_sh[0], a_sh[1], a_sh[2]}, aa[3] = {a_sh[3], a_sh[4], a_sh[5]};
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        a_sh[i] = k.R[i] * al[0] + k.R[3 + i] * al[1] + k.R[6 + i] * al[2];
        a_sh[3 + i] = k.R[i] * aa[0] + k.R[3 + i] * aa[1] + k.R[6 + i] * aa[2];
      }
    }
    diag.niter = iter;
    {   // mj_checkAcc: a bad solver acceleration resets the simulation state (handled by the caller, see EnvOps::step_env)
      bool bad = false;
#pragma unroll
      for (int i = 0; i < 8; ++i) bad = bad || is_bad(a_sh[i]);
#pragma unroll 1
      for (int s = 0; s < WPL; ++s) bad = bad || is_bad(wh[s].a);
      diag.bad_acc = bad ? 1 : 0;
    }

    // ---- B16 implicit joint damping.  At the minimiser M~ a = tau + J^T f, so MuJoCo's integration acceleration
    // (M~ + hB)^-1 (tau + J^T f) equals a - y with (M~ + hB) y = hB a.  B acts on the hinges only; after eliminating the
    // spin dofs the 8x8 system matrix is a constant (inverse P precomputed by the model compiler) minus a rank-2 term
    // from the two front spin axes, handled with the Woodbury identity.
    T yi_sh[8];
    {
      T part3[3] = {T(0), T(0), T(0)};
#pragma unroll 1
      for (int s = 0; s < WPL; ++s) {
        const int wi = wheel_index(lane, s);
        const WheelK<T> wk = wheel_consts(C, wi);
        const WheelT& w = wh[s];
        const T coef = pair_weight() * wk.J * (h * wk.damp * w.a) * C.w_cEinv[wi];
        part3[0] -= coef * w.ax; part3[1] -= coef * w.ay;
      }
      Tm::sum_n(part3);
      T r[8] = {T(0), T(0), T(0), part3[0], part3[1], part3[2], h * C.h_damping[0] * a_sh[6], h * C.h_damping[1] * a_sh[7]};
      T t[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        T acc = T(0);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc += C.eulerP[i >= j ? tri(i, j) : tri(j, i)] * r[j];
        t[i] = acc;
      }
      // front spin axes (both are needed by every lane): u_w = (-sin s_w, cos s_w, 0) in the angular rows
      T ux[2], uy[2];
#pragma unroll
      for (int f = 0; f < 2; ++f) { T sn, cs; N::sincos_small(e.st[f], &sn, &cs); ux[f] = -sn; uy[f] = cs; }
      const T Pxx = C.eulerP[tri(3, 3)], Pxy = C.eulerP[tri(4, 3)], Pyy = C.eulerP[tri(4, 4)];
      T G[2][2], ut[2];
#pragma unroll
      for (int f = 0; f < 2; ++f) {
        ut[f] = ux[f] * t[3] + uy[f] * t[4];
#pragma unroll
        for (int g = 0; g < 2; ++g) G[f][g] = -(ux[f] * (Pxx * ux[g] + Pxy * uy[g]) + uy[f] * (Pxy * ux[g] + Pyy * uy[g]));
        G[f][f] += T(1) / C.euler_kappa[f];
      }
      const T det = G[0][0] * G[1][1] - G[0][1] * G[1][0];
      const T idet = N::rcp_(det);
      const T m0 = (G[1][1] * ut[0] - G[0][1] * ut[1]) * idet, m1 = (G[0][0] * ut[1] - G[1][0] * ut[0]) * idet;
      const T vx = ux[0] * m0 + ux[1] * m1, vy = uy[0] * m0 + uy[1] * m1;   // U m, angular x / y rows
#pragma unroll
      for (int i = 0; i < 8; ++i)
        yi_sh[i] = t[i] + C.eulerP[i >= 3 ? tri(i, 3) : tri(3, i)] * vx + C.eulerP[i >= 4 ? tri(i, 4) : tri(4, i)] * vy;
    }

    if (tap) {
      for (int i = 0; i < 8; ++i) { tap->a_smooth[i] = T(0); tap->a[i] = a_sh[i]; tap->fc[i] = T(0); }
      for (int s = 0; s < WPL; ++s) {
        int wi = wheel_index(lane, s);
        tap->tau[8 + wi] = wh[s].tau; tap->a_smooth[8 + wi] = T(0); tap->a[8 + wi] = wh[s].a; tap->fc[8 + wi] = T(0);
      }
      tap->niter = iter; tap->nls = nls;
    }

    // warm start for the next step = solver acceleration (MuJoCo coordinates); B16: velocities advance with the
    // implicitly damped acceleration a - y, then positions with the new velocities (semi-implicit Euler)
#pragma unroll 1
    for (int s = 0; s < WPL; ++s) {
      const int wi = wheel_index(lane, s);
      const WheelK<T> wk = wheel_consts(C, wi);
      WheelT& w = wh[s];
      const T yi = (h * wk.damp * w.a - wk.J * (w.ax * yi_sh[3] + w.ay * yi_sh[4])) * C.w_cEinv[wi];
      w.warm = w.a;
      w.dsp += h * (w.a - yi);
      w.sp += h * w.dsp;
    }
    T ai[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) ai[i] = a_sh[i] - yi_sh[i];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      e.warm_l[i] = k.R[3 * i] * a_sh[0] + k.R[3 * i + 1] * a_sh[1] + k.R[3 * i + 2] * a_sh[2];
      e.warm_a[i] = a_sh[3 + i];
      e.vw[i] += h * (k.R[3 * i] * ai[0] + k.R[3 * i + 1] * ai[1] + k.R[3 * i + 2] * ai[2]);
      e.om[i] += h * ai[3 + i];
    }
    e.warm_st[0] = a_sh[6]; e.warm_st[1] = a_sh[7];
#pragma unroll
    for (int i = 0; i < 2; ++i) { e.dst[i] += h * ai[6 + i]; e.st[i] += h * e.dst[i]; }
#pragma unroll
    for (int i = 0; i < 3; ++i) e.p[i] += h * e.vw[i];
    {
      T wn = N::sqrt_(dot3(e.om, e.om));
      T ax[3] = {T(1), T(0), T(0)};
      if (wn >= N::minval) { const T iw = N::rcp_(wn); ax[0] = e.om[0] * iw; ax[1] = e.om[1] * iw; ax[2] = e.om[2] * iw; }
      T half = T(0.5) * h * wn, sn, cs;
      N::sincos_tiny(half, &sn, &cs);
      T r0 = cs, r1 = ax[0] * sn, r2 = ax[1] * sn, r3 = ax[2] * sn;
      T q0 = e.q[0], q1 = e.q[1], q2 = e.q[2], q3 = e.q[3];
      e.q[0] = q0 * r0 - q1 * r1 - q2 * r2 - q3 * r3;
      e.q[1] = q0 * r1 + q1 * r0 + q2 * r3 - q3 * r2;
      e.q[2] = q0 * r2 - q1 * r3 + q2 * r0 + q3 * r1;
      e.q[3] = q0 * r3 + q1 * r2 - q2 * r1 + q3 * r0;
    }
  }

  // nearest intersection (ray parameter >= 0) of the ray lp + t dw with one obstacle box centred at the origin, or -1
  ACKB_HD static T box_ray(const Consts<T>& C, T lx, T ly, T lz, const T* dw) {
    const T lp[3] = {lx, ly, lz};
    T best = T(-1);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      if (N::abs_(dw[i]) <= N::minval) continue;
      const int a = (i + 1) % 3, b = (i + 2) % 3;
      const T id = N::rcp_(dw[i]);
#pragma unroll
      for (int side = -1; side <= 1; side += 2) {
        const T sol = (T(side) * C.box_half[i] - lp[i]) * id;
        const T pa = lp[a] + sol * dw[a], pb = lp[b] + sol * dw[b];
        const bool ok = sol >= T(0) && N::abs_(pa) <= C.box_half[a] && N::abs_(pb) <= C.box_half[b] && (best < T(0) || sol < best);
        best = ok ? sol : best;
      }
    }
    return best;
  }

  // ---- B9 rangefinder of observation slot `slot` (ray against the floor plane; boxes in the scene)
  // Per-environment part, computed once for all beams: B = world position of the lidar centre (beam origins sit on a circle of
  // radius lidar_r around it, along the beam direction: o = B + r dw), h0 = its height above the floor plane.
  struct LidarBase { T B[3]; T h0; };
  ACKB_HD static void lidar_base(const Consts<T>& C, const State& e, const Kin<T>& k, LidarBase& b) {
#pragma unroll
    for (int i = 0; i < 3; ++i) b.B[i] = e.p[i] + k.R[3 * i] * C.lidar_pos[0] + k.R[3 * i + 1] * C.lidar_pos[1] + k.R[3 * i + 2] * C.lidar_pos[2];
    b.h0 = b.B[2] - C.plane_z[0];
  }
  ACKB_HD static T lidar_ray(const Consts<T>& C, const Kin<T>& k, const LidarBase& b, int beam) {
    T cb, sb;
    tbl_beam(C, beam, &cb, &sb);
    T best = T(-1);
    // ray direction in world axes; its z component is the floor normal (body frame) dotted with the beam direction
    const T lvz = k.n[0] * cb + k.n[1] * sb;
    const T dw[3] = {k.R[0] * cb + k.R[1] * sb, k.R[3] * cb + k.R[4] * sb, lvz};
    const T r = C.lidar_r[0];
    // floor: the origin's height is h0 + r lvz, so the ray parameter of the plane is x = -(h0 + r lvz) / lvz = s - r with
    // s = -h0 / lvz, and the hit point is B + s dw (mj_ray's plane test: x >= 0, hit point inside the plane's half sizes)
    if (!(lvz > -N::minval)) {
      const T sp = -b.h0 * N::rcp_(lvz);
      const T x = sp - r;
      if (x >= T(0)) {
        const T px = b.B[0] + sp * dw[0], py = b.B[1] + sp * dw[1];
        const bool in = (C.plane_half[0] <= T(0) || N::abs_(px) <= C.plane_half[0]) && (C.plane_half[1] <= T(0) || N::abs_(py) <= C.plane_half[1]);
        best = in ? x : best;
      }
    }
    if (C.nbox[0] == T(0)) {      // no obstacles (flat-floor model): done
      if (C.lidar_cutoff[0] > T(0) && best > C.lidar_cutoff[0]) best = C.lidar_cutoff[0];
      return best;
    }
    const T ow[3] = {b.B[0] + r * dw[0], b.B[1] + r * dw[1], b.B[2] + r * dw[2]};
    const int nbox = (int)C.nbox[0];
    if (C.grid_on[0] != T(0)) {
      // boxes on a lattice: walk the cells along the ray (2-D DDA) and test only the boxes met; the first box hit is the
      // nearest one because the boxes are disjoint and cells are visited in order of entry distance
      const T pitch = C.grid_pitch[0], ip = N::rcp_(pitch);
      const T gx = (ow[0] - C.grid_x0[0]) * ip, gy = (ow[1] - C.grid_y0[0]) * ip;
      int ix = (int)floor(gx), iy = (int)floor(gy);
      const int nx = (int)C.grid_nx[0], ny = (int)C.grid_ny[0];
      if (ix >= 0 && iy >= 0 && ix < nx && iy < ny) {
        const int sx = dw[0] > T(0) ? 1 : -1, sy = dw[1] > T(0) ? 1 : -1;
        const T big = T(1e30);
        const bool mx = N::abs_(dw[0]) > N::minval, my = N::abs_(dw[1]) > N::minval;
        const T idx = mx ? N::rcp_(dw[0]) : T(0), idy = my ? N::rcp_(dw[1]) : T(0);
        const T tdx = mx ? N::abs_(pitch * idx) : big, tdy = my ? N::abs_(pitch * idy) : big;
        T tmx = mx ? ((T)(ix + (sx > 0 ? 1 : 0)) - gx) * pitch * idx : big;
        T tmy = my ? ((T)(iy + (sy > 0 ? 1 : 0)) - gy) * pitch * idy : big;
        const int budget = nx + ny + 2;
        const bool mz = N::abs_(dw[2]) > N::minval;
        const T idz = mz ? N::rcp_(dw[2]) : T(0);
        int entered = -1;                                   // axis of the cell boundary crossed last (-1: the origin's own cell)
#pragma unroll 1
        for (int step = 0; step < budget; ++step) {
          if ((tbl_grid_row(C, iy) >> ix) & 1u) {
            const T lx = ow[0] - (C.grid_x0[0] + (T(ix) + T(0.5)) * pitch), ly = ow[1] - (C.grid_y0[0] + (T(iy) + T(0.5)) * pitch);
            const T lz = ow[2] - C.box_z[0];
            T h;
            if (entered < 0) h = box_ray(C, lx, ly, lz, dw);      // origin over / inside this box: the general six-face test
            else {
              // the box fills the cell, so from outside it can only be hit on the lateral face just crossed or on the z face
              // turned towards the ray: the same face formulas as box_ray, evaluated for these two faces only
              h = T(-1);
              const bool ex = entered == 0;
              const T side = ex ? T(-sx) : T(-sy);
              const T sol = ex ? (side * C.box_half[0] - lx) * idx : (side * C.box_half[1] - ly) * idy;
              const T pa = ex ? ly + sol * dw[1] : lz + sol * dw[2], pb = ex ? lz + sol * dw[2] : lx + sol * dw[0];
              const T ha = ex ? C.box_half[1] : C.box_half[2], hb = ex ? C.box_half[2] : C.box_half[0];
              if (sol >= T(0) && N::abs_(pa) <= ha && N::abs_(pb) <= hb) h = sol;
              if (mz) {
                const T sz = dw[2] > T(0) ? T(-1) : T(1);
                const T solz = (sz * C.box_half[2] - lz) * idz;
                const T px = lx + solz * dw[0], py = ly + solz * dw[1];
                if (solz >= T(0) && N::abs_(px) <= C.box_half[0] && N::abs_(py) <= C.box_half[1] && (h < T(0) || solz < h)) h = solz;
              }
            }
            if (h >= T(0)) { if (best < T(0) || h < best) best = h; break; }
          }
          const T tn = tmx < tmy ? tmx : tmy;             // ray parameter at which the next cell is entered
          if (best >= T(0) && tn > best) break;            // the floor is hit first
          if (tmx < tmy) { ix += sx; tmx += tdx; entered = 0; } else { iy += sy; tmy += tdy; entered = 1; }
          if (ix < 0 || iy < 0 || ix >= nx || iy >= ny) break;
        }
      } else {
        for (int bi = 0; bi < nbox; ++bi) {
          const T h = box_ray(C, ow[0] - C.box_cx[bi], ow[1] - C.box_cy[bi], ow[2] - C.box_z[0], dw);
          if (h >= T(0) && (best < T(0) || h < best)) best = h;
        }
      }
    } else {
      for (int bi = 0; bi < nbox; ++bi) {
        const T h = box_ray(C, ow[0] - C.box_cx[bi], ow[1] - C.box_cy[bi], ow[2] - C.box_z[0], dw);
        if (h >= T(0) && (best < T(0) || h < best)) best = h;
      }
    }
    if (C.lidar_cutoff[0] > T(0) && best > C.lidar_cutoff[0]) best = C.lidar_cutoff[0];
    return best;
  }
};

}  // namespace ackb
