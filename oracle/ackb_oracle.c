/*
 * ackb_oracle.c -- CPU fp64 ORACLE for the Ackermann env-step hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product package imports, links or
 * executes this file; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs do (as the checker / CPU baseline).
 *
 * PARITY UNPINNED for the physics: the arithmetic of the reference's hot path
 * lives in the third-party `mujoco` wheel (requirements.txt:4, `mujoco>=3.0.0`,
 * unpinned, source not under /root/reference, not installable here).  This file
 * restates MuJoCo's published computation pipeline (mj_step = mj_forward + Euler)
 * for exactly the feature subset that models/ackermann_robot_v2.xml and
 * models/environments/ackermann_maze_flat.xml exercise, anchored on the reference's
 * call sites:
 *     src/rl/envs/ackermann_env.py:200            mujoco.mj_step(model, data)
 *     src/rl/envs/simple_map_spawner.py:37-52     from_xml_path / MjData / mj_forward
 *     src/core/controller.py:138-140              data.ctrl[...] writes
 *     src/rl/envs/ackermann_env.py:234-237        data.sensordata[...] reads
 *     src/core/odometry.py:79-80                  data.xpos / data.xquat reads
 * The reference holds no tests or golden vectors for this path, so
 * the oracle is pinned only by analytic invariants (tests/test_oracle_physics.py).
 *
 * The implementation is table driven (general kinematic tree, dense nv x nv algebra)
 * and deliberately shares no code with the topology-specialised CUDA kernels.
 *
 * Pipeline: kinematics -> comPos -> CRB -> factor M ->
 * collision -> makeConstraint -> sensorPos (rays) -> comVel/passive/RNE -> sensorVel ->
 * actuation -> acceleration -> constraint solve (Newton, exact line search) ->
 * Euler with implicit joint damping.
 */
#include <math.h>
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define MAXB 64   /* bodies   */
#define MAXJ 16   /* joints   */
#define MAXV 16   /* dofs     */
#define MAXQ 20   /* qpos     */
#define MAXG 96   /* geoms    */
#define MAXS 80   /* sites    */
#define MAXSEN 96 /* sensors  */
#define MAXU 8
#define MAXEQ 4
#define MAXCON 64
#define MAXEFC 300
#define MAXHULL 128
#define MAXADJ 24

#define MINVAL 1e-15
#define MINIMP 0.0001
#define MAXIMP 0.9999

enum { JNT_FREE = 0, JNT_BALL = 1, JNT_SLIDE = 2, JNT_HINGE = 3 };
enum { G_PLANE = 0, G_SPHERE = 2, G_CAPSULE = 3, G_CYLINDER = 5, G_BOX = 6, G_MESH = 7 };
enum { C_EQUALITY = 0, C_FRICTION_DOF = 1, C_LIMIT = 3, C_CONTACT_PYR = 6 };

typedef struct {
  /* sizes */
  int nq, nv, nu, nbody, njnt, ngeom, nsite, nsensor, neq, nsensordata, nhullvert;
  /* options */
  double opt_timestep, opt_gravity[3], opt_impratio, opt_tolerance, opt_ls_tolerance;
  int opt_iterations, opt_ls_iterations;
  double stat_meaninertia;
  /* bodies */
  int body_parentid[MAXB], body_rootid[MAXB], body_weldid[MAXB], body_jntnum[MAXB], body_jntadr[MAXB],
      body_dofnum[MAXB], body_dofadr[MAXB];
  double body_pos[MAXB * 3], body_quat[MAXB * 4], body_ipos[MAXB * 3], body_iquat[MAXB * 4], body_mass[MAXB],
      body_inertia[MAXB * 3], body_invweight0[MAXB * 2];
  /* joints */
  int jnt_type[MAXJ], jnt_qposadr[MAXJ], jnt_dofadr[MAXJ], jnt_bodyid[MAXJ], jnt_limited[MAXJ];
  double jnt_pos[MAXJ * 3], jnt_axis[MAXJ * 3], jnt_range[MAXJ * 2], jnt_margin[MAXJ], jnt_solref[MAXJ * 2],
      jnt_solimp[MAXJ * 5];
  double qpos0[MAXQ];
  /* dofs */
  int dof_bodyid[MAXV], dof_jntid[MAXV], dof_parentid[MAXV];
  double dof_armature[MAXV], dof_damping[MAXV], dof_frictionloss[MAXV], dof_invweight0[MAXV], dof_solref[MAXV * 2],
      dof_solimp[MAXV * 5];
  /* geoms */
  int geom_type[MAXG], geom_bodyid[MAXG], geom_contype[MAXG], geom_conaffinity[MAXG], geom_condim[MAXG],
      geom_priority[MAXG], geom_hulladr[MAXG], geom_hullnum[MAXG];
  double geom_size[MAXG * 3], geom_pos[MAXG * 3], geom_quat[MAXG * 4], geom_friction[MAXG * 3], geom_solref[MAXG * 2],
      geom_solimp[MAXG * 5], geom_solmix[MAXG], geom_margin[MAXG], geom_gap[MAXG], geom_alpha[MAXG];
  double hull_vert[MAXHULL * 3];
  int hull_adj[MAXHULL * MAXADJ];   /* hull graph: neighbours of every hull vertex (local indices, -1 padded), MuJoCo's MakeGraph order */
  double geom_rbound[MAXG];
  /* sites / sensors */
  int site_bodyid[MAXS];
  double site_pos[MAXS * 3], site_quat[MAXS * 4];
  int sensor_type[MAXSEN], sensor_objid[MAXSEN], sensor_adr[MAXSEN];
  double sensor_cutoff[MAXSEN];
  /* equality */
  int eq_obj1id[MAXEQ], eq_obj2id[MAXEQ];
  double eq_data[MAXEQ * 5], eq_solref[MAXEQ * 2], eq_solimp[MAXEQ * 5];
  /* actuators */
  int actuator_trnid[MAXU], actuator_ctrllimited[MAXU], actuator_forcelimited[MAXU];
  double actuator_gear[MAXU], actuator_gainprm[MAXU], actuator_biasprm[MAXU * 3], actuator_ctrlrange[MAXU * 2],
      actuator_forcerange[MAXU * 2];
} OModel;

typedef struct {
  int geom1, geom2, dim, exclude;
  double dist, pos[3], frame[9], includemargin, friction[5], solref[2], solimp[5], mu;
} OContact;

typedef struct {
  /* state */
  double time, qpos[MAXQ], qvel[MAXV], ctrl[MAXU], qacc_warmstart[MAXV];
  /* position stage */
  double xpos[MAXB * 3], xquat[MAXB * 4], xmat[MAXB * 9], xipos[MAXB * 3], ximat[MAXB * 9], xanchor[MAXJ * 3],
      xaxis[MAXJ * 3], geom_xpos[MAXG * 3], geom_xmat[MAXG * 9], site_xpos[MAXS * 3], site_xmat[MAXS * 9],
      subtree_com[MAXB * 3], cinert[MAXB * 10], crb[MAXB * 10], cdof[MAXV * 6], cdof_dot[MAXV * 6], cvel[MAXB * 6];
  double qM[MAXV * MAXV], qL[MAXV * MAXV]; /* dense mass matrix and its Cholesky factor */
  /* velocity / force stage */
  double qfrc_bias[MAXV], qfrc_passive[MAXV], qfrc_actuator[MAXV], actuator_force[MAXU], qfrc_smooth[MAXV],
      qacc_smooth[MAXV], qacc[MAXV], qfrc_constraint[MAXV];
  /* contacts / constraints */
  int ncon, nefc, unsupported_contact;
  OContact contact[MAXCON];
  int efc_type[MAXEFC], efc_id[MAXEFC];
  double efc_J[MAXEFC * MAXV], efc_pos[MAXEFC], efc_margin[MAXEFC], efc_frictionloss[MAXEFC], efc_diagApprox[MAXEFC],
      efc_R[MAXEFC], efc_D[MAXEFC], efc_aref[MAXEFC], efc_force[MAXEFC], efc_solref[MAXEFC * 2], efc_solimp[MAXEFC * 5],
      efc_KBIP[MAXEFC * 4];
  double sensordata[MAXSEN];
  int solver_niter, solver_lsiter;
  double solver_cost, solver_gradnorm;
  int nbad;   /* states reset by mj_checkPos / mj_checkVel / mj_checkAcc */
} OData;

/* ------------------------------------------------------------------------- */
/*  small vector / quaternion algebra                                          */
/* ------------------------------------------------------------------------- */
static double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static void cross3(double* r, const double* a, const double* b) {
  double t0 = a[1] * b[2] - a[2] * b[1], t1 = a[2] * b[0] - a[0] * b[2], t2 = a[0] * b[1] - a[1] * b[0];
  r[0] = t0; r[1] = t1; r[2] = t2;
}
static double norm3(const double* a) { return sqrt(dot3(a, a)); }
static double normalize3(double* a) {
  double n = norm3(a);
  if (n < MINVAL) { a[0] = 1; a[1] = 0; a[2] = 0; } else { a[0] /= n; a[1] /= n; a[2] /= n; }
  return n;
}
static void normalize4(double* q) {
  double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  if (n < MINVAL) { q[0] = 1; q[1] = q[2] = q[3] = 0; } else { for (int i = 0; i < 4; i++) q[i] /= n; }
}
static void mulquat(double* r, const double* a, const double* b) {
  double t[4] = {a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3], a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2],
                 a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1], a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0]};
  memcpy(r, t, sizeof t);
}
static void quat2mat(double* m, const double* q) {
  double q00 = q[0] * q[0], q01 = q[0] * q[1], q02 = q[0] * q[2], q03 = q[0] * q[3], q11 = q[1] * q[1], q12 = q[1] * q[2],
         q13 = q[1] * q[3], q22 = q[2] * q[2], q23 = q[2] * q[3], q33 = q[3] * q[3];
  m[0] = q00 + q11 - q22 - q33; m[4] = q00 - q11 + q22 - q33; m[8] = q00 - q11 - q22 + q33;
  m[1] = 2 * (q12 - q03); m[2] = 2 * (q13 + q02); m[3] = 2 * (q12 + q03);
  m[5] = 2 * (q23 - q01); m[6] = 2 * (q13 - q02); m[7] = 2 * (q23 + q01);
}
static void mulmatvec3(double* r, const double* m, const double* v) {
  double t0 = m[0] * v[0] + m[1] * v[1] + m[2] * v[2], t1 = m[3] * v[0] + m[4] * v[1] + m[5] * v[2],
         t2 = m[6] * v[0] + m[7] * v[1] + m[8] * v[2];
  r[0] = t0; r[1] = t1; r[2] = t2;
}
static void mulmatTvec3(double* r, const double* m, const double* v) {
  double t0 = m[0] * v[0] + m[3] * v[1] + m[6] * v[2], t1 = m[1] * v[0] + m[4] * v[1] + m[7] * v[2],
         t2 = m[2] * v[0] + m[5] * v[1] + m[8] * v[2];
  r[0] = t0; r[1] = t1; r[2] = t2;
}
static void rotvecquat(double* r, const double* v, const double* q) {
  double m[9];
  quat2mat(m, q);
  mulmatvec3(r, m, v);
}
static void axisangle2quat(double* q, const double* axis, double ang) {
  double s = sin(ang * 0.5);
  q[0] = cos(ang * 0.5); q[1] = axis[0] * s; q[2] = axis[1] * s; q[3] = axis[2] * s;
}

/* spatial algebra, MuJoCo layout: motion/force vectors are [rot(3); lin(3)], cinert is
   [Ixx Iyy Izz Ixy Ixz Iyz  m*cx m*cy m*cz  m] about the frame origin */
static void mul_inert_vec(double* r, const double* i, const double* v) {
  r[0] = i[0] * v[0] + i[3] * v[1] + i[4] * v[2] - i[8] * v[4] + i[7] * v[5];
  r[1] = i[3] * v[0] + i[1] * v[1] + i[5] * v[2] + i[8] * v[3] - i[6] * v[5];
  r[2] = i[4] * v[0] + i[5] * v[1] + i[2] * v[2] - i[7] * v[3] + i[6] * v[4];
  r[3] = i[8] * v[1] - i[7] * v[2] + i[9] * v[3];
  r[4] = i[6] * v[2] - i[8] * v[0] + i[9] * v[4];
  r[5] = i[7] * v[0] - i[6] * v[1] + i[9] * v[5];
}
static void cross_motion(double* r, const double* vel, const double* v) {
  double a[3], b[3], c[3];
  cross3(a, vel, v);          /* w x v_rot */
  cross3(b, vel, v + 3);      /* w x v_lin */
  cross3(c, vel + 3, v);      /* vlin x v_rot */
  for (int k = 0; k < 3; k++) { r[k] = a[k]; r[3 + k] = b[k] + c[k]; }
}
static void cross_force(double* r, const double* vel, const double* f) {
  double a[3], b[3], c[3];
  cross3(a, vel, f);          /* w x f_rot */
  cross3(b, vel + 3, f + 3);  /* vlin x f_lin */
  cross3(c, vel, f + 3);      /* w x f_lin */
  for (int k = 0; k < 3; k++) { r[k] = a[k] + b[k]; r[3 + k] = c[k]; }
}

/* ------------------------------------------------------------------------- */
/*  B1 kinematics                                                              */
/* ------------------------------------------------------------------------- */
static void o_kinematics(const OModel* m, OData* d) {
  memset(d->xpos, 0, 3 * sizeof(double));
  d->xquat[0] = 1; d->xquat[1] = d->xquat[2] = d->xquat[3] = 0;
  quat2mat(d->xmat, d->xquat);
  memset(d->xipos, 0, 3 * sizeof(double));
  quat2mat(d->ximat, d->xquat);
  /* normalise quaternions stored in qpos */
  for (int j = 0; j < m->njnt; j++)
    if (m->jnt_type[j] == JNT_FREE) normalize4(d->qpos + m->jnt_qposadr[j] + 3);

  for (int i = 1; i < m->nbody; i++) {
    double* xpos = d->xpos + 3 * i; double* xquat = d->xquat + 4 * i;
    int ja = m->body_jntadr[i], jn = m->body_jntnum[i], pid = m->body_parentid[i];
    if (jn == 1 && m->jnt_type[ja] == JNT_FREE) {
      int qa = m->jnt_qposadr[ja];
      memcpy(xpos, d->qpos + qa, 3 * sizeof(double));
      memcpy(xquat, d->qpos + qa + 3, 4 * sizeof(double));
      memcpy(d->xanchor + 3 * ja, xpos, 3 * sizeof(double));
      memcpy(d->xaxis + 3 * ja, m->jnt_axis + 3 * ja, 3 * sizeof(double));
    } else {
      double t[3];
      mulmatvec3(t, d->xmat + 9 * pid, m->body_pos + 3 * i);
      for (int k = 0; k < 3; k++) xpos[k] = d->xpos[3 * pid + k] + t[k];
      mulquat(xquat, d->xquat + 4 * pid, m->body_quat + 4 * i);
      for (int j = ja; j < ja + jn; j++) {
        double* xanchor = d->xanchor + 3 * j; double* xaxis = d->xaxis + 3 * j;
        rotvecquat(xaxis, m->jnt_axis + 3 * j, xquat);
        rotvecquat(xanchor, m->jnt_pos + 3 * j, xquat);
        for (int k = 0; k < 3; k++) xanchor[k] += xpos[k];
        int qa = m->jnt_qposadr[j];
        if (m->jnt_type[j] == JNT_HINGE) {
          double ql[4], vec[3];
          axisangle2quat(ql, m->jnt_axis + 3 * j, d->qpos[qa] - m->qpos0[qa]);
          mulquat(xquat, xquat, ql);
          rotvecquat(vec, m->jnt_pos + 3 * j, xquat);
          for (int k = 0; k < 3; k++) xpos[k] = xanchor[k] - vec[k];
        } else if (m->jnt_type[j] == JNT_SLIDE) {
          for (int k = 0; k < 3; k++) xpos[k] += xaxis[k] * (d->qpos[qa] - m->qpos0[qa]);
        }
      }
    }
    normalize4(xquat);
    quat2mat(d->xmat + 9 * i, xquat);
  }
  for (int i = 1; i < m->nbody; i++) {
    double t[3], q[4];
    mulmatvec3(t, d->xmat + 9 * i, m->body_ipos + 3 * i);
    for (int k = 0; k < 3; k++) d->xipos[3 * i + k] = d->xpos[3 * i + k] + t[k];
    mulquat(q, d->xquat + 4 * i, m->body_iquat + 4 * i);
    quat2mat(d->ximat + 9 * i, q);
  }
  for (int g = 0; g < m->ngeom; g++) {
    int b = m->geom_bodyid[g]; double t[3], q[4];
    mulmatvec3(t, d->xmat + 9 * b, m->geom_pos + 3 * g);
    for (int k = 0; k < 3; k++) d->geom_xpos[3 * g + k] = d->xpos[3 * b + k] + t[k];
    mulquat(q, d->xquat + 4 * b, m->geom_quat + 4 * g);
    quat2mat(d->geom_xmat + 9 * g, q);
  }
  for (int s = 0; s < m->nsite; s++) {
    int b = m->site_bodyid[s]; double t[3], q[4];
    mulmatvec3(t, d->xmat + 9 * b, m->site_pos + 3 * s);
    for (int k = 0; k < 3; k++) d->site_xpos[3 * s + k] = d->xpos[3 * b + k] + t[k];
    mulquat(q, d->xquat + 4 * b, m->site_quat + 4 * s);
    quat2mat(d->site_xmat + 9 * s, q);
  }
}

/* ------------------------------------------------------------------------- */
/*  B2 comPos: subtree_com, cinert, cdof                                       */
/* ------------------------------------------------------------------------- */
static void o_compos(const OModel* m, OData* d) {
  double mass_subtree[MAXB];
  for (int i = 0; i < m->nbody; i++) {
    mass_subtree[i] = m->body_mass[i];
    for (int k = 0; k < 3; k++) d->subtree_com[3 * i + k] = m->body_mass[i] * d->xipos[3 * i + k];
  }
  for (int i = m->nbody - 1; i > 0; i--) {
    int p = m->body_parentid[i];
    mass_subtree[p] += mass_subtree[i];
    for (int k = 0; k < 3; k++) d->subtree_com[3 * p + k] += d->subtree_com[3 * i + k];
  }
  for (int i = 0; i < m->nbody; i++) {
    if (mass_subtree[i] < MINVAL) memcpy(d->subtree_com + 3 * i, d->xipos + 3 * i, 3 * sizeof(double));
    else for (int k = 0; k < 3; k++) d->subtree_com[3 * i + k] /= mass_subtree[i];
  }
  memset(d->cinert, 0, 10 * sizeof(double));
  for (int i = 1; i < m->nbody; i++) {
    const double* R = d->ximat + 9 * i; const double* I = m->body_inertia + 3 * i;
    double mass = m->body_mass[i], dif[3], tmp[9];
    for (int k = 0; k < 3; k++) dif[k] = d->xipos[3 * i + k] - d->subtree_com[3 * m->body_rootid[i] + k];
    for (int r = 0; r < 3; r++)
      for (int c = 0; c < 3; c++) tmp[3 * r + c] = R[3 * r] * I[0] * R[3 * c] + R[3 * r + 1] * I[1] * R[3 * c + 1] + R[3 * r + 2] * I[2] * R[3 * c + 2];
    double* ci = d->cinert + 10 * i;
    ci[0] = tmp[0] + mass * (dif[1] * dif[1] + dif[2] * dif[2]);
    ci[1] = tmp[4] + mass * (dif[0] * dif[0] + dif[2] * dif[2]);
    ci[2] = tmp[8] + mass * (dif[0] * dif[0] + dif[1] * dif[1]);
    ci[3] = tmp[1] - mass * dif[0] * dif[1];
    ci[4] = tmp[2] - mass * dif[0] * dif[2];
    ci[5] = tmp[5] - mass * dif[1] * dif[2];
    ci[6] = mass * dif[0]; ci[7] = mass * dif[1]; ci[8] = mass * dif[2]; ci[9] = mass;
  }
  for (int j = 0; j < m->njnt; j++) {
    int b = m->jnt_bodyid[j], da = m->jnt_dofadr[j];
    double off[3];
    for (int k = 0; k < 3; k++) off[k] = d->subtree_com[3 * m->body_rootid[b] + k] - d->xanchor[3 * j + k];
    if (m->jnt_type[j] == JNT_FREE) {
      memset(d->cdof + 6 * da, 0, 18 * sizeof(double));
      d->cdof[6 * da + 3] = 1; d->cdof[6 * (da + 1) + 4] = 1; d->cdof[6 * (da + 2) + 5] = 1;
      for (int k = 0; k < 3; k++) {
        double ax[3] = {d->xmat[9 * b + k], d->xmat[9 * b + 3 + k], d->xmat[9 * b + 6 + k]};
        double* c = d->cdof + 6 * (da + 3 + k);
        memcpy(c, ax, sizeof ax);
        cross3(c + 3, ax, off);
      }
    } else if (m->jnt_type[j] == JNT_HINGE) {
      double* c = d->cdof + 6 * da;
      memcpy(c, d->xaxis + 3 * j, 3 * sizeof(double));
      cross3(c + 3, d->xaxis + 3 * j, off);
    } else { /* slide */
      double* c = d->cdof + 6 * da;
      c[0] = c[1] = c[2] = 0;
      memcpy(c + 3, d->xaxis + 3 * j, 3 * sizeof(double));
    }
  }
}

/* ------------------------------------------------------------------------- */
/*  B4/B5 composite rigid body + dense Cholesky                                */
/* ------------------------------------------------------------------------- */
static int chol_factor(double* L, const double* A, int n, int ld) {
  /* lower Cholesky A = L L^T (dense, row-major, leading dimension ld) */
  for (int i = 0; i < n; i++)
    for (int j = 0; j <= i; j++) {
      double s = A[i * ld + j];
      for (int k = 0; k < j; k++) s -= L[i * ld + k] * L[j * ld + k];
      if (i == j) { if (s < MINVAL) return -1; L[i * ld + i] = sqrt(s); }
      else L[i * ld + j] = s / L[j * ld + j];
    }
  return 0;
}
static void chol_solve(double* x, const double* L, const double* b, int n, int ld) {
  double y[MAXV];
  for (int i = 0; i < n; i++) { double s = b[i]; for (int k = 0; k < i; k++) s -= L[i * ld + k] * y[k]; y[i] = s / L[i * ld + i]; }
  for (int i = n - 1; i >= 0; i--) { double s = y[i]; for (int k = i + 1; k < n; k++) s -= L[k * ld + i] * x[k]; x[i] = s / L[i * ld + i]; }
}

static void o_crb(const OModel* m, OData* d) {
  int nv = m->nv;
  memcpy(d->crb, d->cinert, sizeof(double) * 10 * m->nbody);
  for (int i = m->nbody - 1; i > 0; i--) {
    int p = m->body_parentid[i];
    if (p > 0) for (int k = 0; k < 10; k++) d->crb[10 * p + k] += d->crb[10 * i + k];
  }
  memset(d->qM, 0, sizeof d->qM);
  for (int i = 0; i < nv; i++) {
    double buf[6];
    mul_inert_vec(buf, d->crb + 10 * m->dof_bodyid[i], d->cdof + 6 * i);
    for (int j = i; j >= 0; j = m->dof_parentid[j]) {
      double s = 0;
      for (int k = 0; k < 6; k++) s += d->cdof[6 * j + k] * buf[k];
      d->qM[i * MAXV + j] = d->qM[j * MAXV + i] = s;
    }
    d->qM[i * MAXV + i] += m->dof_armature[i];
  }
  chol_factor(d->qL, d->qM, nv, MAXV);
}

/* ------------------------------------------------------------------------- */
/*  Jacobian of a world point attached to a body                               */
/* ------------------------------------------------------------------------- */
static void o_jac(const OModel* m, const OData* d, double* jacp, double* jacr, const double* point, int body) {
  int nv = m->nv;
  if (jacp) memset(jacp, 0, sizeof(double) * 3 * nv);
  if (jacr) memset(jacr, 0, sizeof(double) * 3 * nv);
  if (body <= 0) return;
  /* offset of point from the root subtree COM (cdof is expressed there) */
  double off[3];
  for (int k = 0; k < 3; k++) off[k] = point[k] - d->subtree_com[3 * m->body_rootid[body] + k];
  /* last dof of the nearest moving ancestor */
  int b = body;
  while (b > 0 && m->body_dofnum[b] == 0) b = m->body_parentid[b];
  if (b <= 0) return;
  int i = m->body_dofadr[b] + m->body_dofnum[b] - 1;
  for (; i >= 0; i = m->dof_parentid[i]) {
    const double* c = d->cdof + 6 * i;
    if (jacr) for (int k = 0; k < 3; k++) jacr[k * nv + i] = c[k];
    if (jacp) { double t[3]; cross3(t, c, off); for (int k = 0; k < 3; k++) jacp[k * nv + i] = c[3 + k] + t[k]; }
  }
}

/* ------------------------------------------------------------------------- */
/*  B6 collision                                                               */
/* ------------------------------------------------------------------------- */
static void make_frame(double* f) {
  normalize3(f);
  if (norm3(f + 3) < 0.5) {
    f[3] = f[4] = f[5] = 0;
    if (f[1] < 0.5 && f[1] > -0.5) f[4] = 1; else f[5] = 1;
  }
  double dp = dot3(f, f + 3);
  for (int k = 0; k < 3; k++) f[3 + k] -= dp * f[k];
  normalize3(f + 3);
  cross3(f + 6, f, f + 3);
}

static int plane_cylinder(const double* pos1, const double* mat1, const double* pos2, const double* mat2,
                          const double* size2, double margin, double* dist_out, double* pos_out) {
  /* plane normal = plane z axis; cylinder axis = its z axis.  Contacts: nearest rim point,
     opposite rim point, then two "triangle" points when the cap faces the plane. */
  double normal[3] = {mat1[2], mat1[5], mat1[8]}, axis[3] = {mat2[2], mat2[5], mat2[8]};
  double vec[3], dif[3];
  int cnt = 0;
  double prjaxis = dot3(normal, axis);
  if (prjaxis > 0) { for (int k = 0; k < 3; k++) axis[k] = -axis[k]; prjaxis = -prjaxis; }
  for (int k = 0; k < 3; k++) dif[k] = pos2[k] - pos1[k];
  double dist = dot3(normal, dif);
  for (int k = 0; k < 3; k++) vec[k] = axis[k] * prjaxis - normal[k];
  double len = norm3(vec);
  if (len >= MINVAL) for (int k = 0; k < 3; k++) vec[k] *= size2[0] / len;
  else { vec[0] = mat2[0] * size2[0]; vec[1] = mat2[3] * size2[0]; vec[2] = mat2[6] * size2[0]; }
  double prjvec = dot3(vec, normal);
  for (int k = 0; k < 3; k++) axis[k] *= size2[1];
  prjaxis *= size2[1];
  if (dist + prjaxis + prjvec <= margin) {
    double dd = dist + prjaxis + prjvec;
    dist_out[cnt] = dd;
    for (int k = 0; k < 3; k++) pos_out[3 * cnt + k] = pos2[k] + vec[k] + axis[k] - normal[k] * dd * 0.5;
    cnt++;
  } else return 0;
  if (dist - prjaxis + prjvec <= margin) {
    double dd = dist - prjaxis + prjvec;
    dist_out[cnt] = dd;
    for (int k = 0; k < 3; k++) pos_out[3 * cnt + k] = pos2[k] + vec[k] - axis[k] - normal[k] * dd * 0.5;
    cnt++;
  }
  double prjvec1 = -prjvec * 0.5;
  if (dist + prjaxis + prjvec1 <= margin) {
    double vec1[3];
    cross3(vec1, vec, axis);
    normalize3(vec1);
    double sc = size2[0] * sqrt(3.0) * 0.5;
    double dd = dist + prjaxis + prjvec1;
    for (int s = -1; s <= 1; s += 2) {
      dist_out[cnt] = dd;
      for (int k = 0; k < 3; k++) pos_out[3 * cnt + k] = pos2[k] + s * sc * vec1[k] + axis[k] - vec[k] * 0.5 - normal[k] * dd * 0.5;
      cnt++;
    }
  }
  return cnt;
}

/* Cylinder (g1) vs box (g2): analytic single contact.
   MuJoCo sends this pair to its general convex collider (one contact along the minimum-penetration direction).  Here that
   contact is computed in closed form: over the six face normals n of the box, take the cylinder's support point in
   direction -n (axial part blended over ~1 degree of tilt so that a rim parallel to the face contacts at its centre);
   the face with the largest separation is the contact face; contact if that separation <= margin.  Edge/corner axes are
   not tested (documented deviation).  Normal points from the cylinder to the box; pos is midway between the surfaces. */
static int cylinder_box(const double* cpos, const double* cmat, const double* csize, const double* bpos,
                        const double* bmat, const double* bsize, double margin, double* dist_out, double* pos_out,
                        double* normal_out) {
  double rel[3], c[3], a[3];
  for (int k = 0; k < 3; k++) rel[k] = cpos[k] - bpos[k];
  mulmatTvec3(c, bmat, rel);
  double axw[3] = {cmat[2], cmat[5], cmat[8]};
  mulmatTvec3(a, bmat, axw);
  /* the spin axis direction sign is irrelevant for the support function */
  double r = csize[0], h = csize[1];
  double best = -1e300; int bestk = 0, bests = 1; double bestp[3] = {0, 0, 0};
  for (int k = 0; k < 3; k++)
    for (int s = -1; s <= 1; s += 2) {
      double an = a[k] * s;
      double w[3] = {an * a[0], an * a[1], an * a[2]};
      w[k] -= s;
      double wl = norm3(w);
      double tcl = -an * 50.0;
      tcl = tcl < -1 ? -1 : (tcl > 1 ? 1 : tcl);
      double sc = wl > 1e-12 ? r / wl : 0.0;
      double p[3];
      for (int q = 0; q < 3; q++) p[q] = c[q] + tcl * h * a[q] + sc * w[q];
      double sep = s * p[k] - bsize[k];
      if (sep > best) { best = sep; bestk = k; bests = s; memcpy(bestp, p, sizeof p); }
    }
  if (best > margin) return 0;
  double n[3] = {0, 0, 0}, pmid[3];
  n[bestk] = -bests; /* from the cylinder into the box */
  memcpy(pmid, bestp, sizeof pmid);
  pmid[bestk] -= bests * best * 0.5;
  double pw[3], nw[3];
  mulmatvec3(pw, bmat, pmid);
  mulmatvec3(nw, bmat, n);
  for (int k = 0; k < 3; k++) { pos_out[k] = pw[k] + bpos[k]; normal_out[k] = nw[k]; }
  dist_out[0] = best;
  return 1;
}

/* Plane (g1) vs convex mesh (g2), MuJoCo's mjc_PlaneConvex [MJ-knowledge, VERIFY against a golden dump]: the support vertex of the
   hull in the direction -normal gives the first contact (none if its distance exceeds the margin); for meshes the neighbours of
   that vertex in the hull graph are then tried in graph order: a neighbour within the margin and at least tolplanemesh * rbound
   (0.3 x bounding radius) away from the first contact adds a contact, up to 3 contacts in all.  pos is midway between the surfaces. */
static int plane_convex(const OModel* m, int g2, const double* pos1, const double* mat1, const double* pos2, const double* mat2,
                        double margin, double* dist_out, double* pos_out) {
  const double normal[3] = {mat1[2], mat1[5], mat1[8]};
  const int adr = m->geom_hulladr[g2], nv = m->geom_hullnum[g2];
  double pw[MAXHULL * 3], dist[MAXHULL];
  int best = -1;
  for (int v = 0; v < nv; v++) {
    double t[3], dif[3];
    mulmatvec3(t, mat2, m->hull_vert + 3 * (adr + v));
    for (int k = 0; k < 3; k++) { pw[3 * v + k] = t[k] + pos2[k]; dif[k] = pw[3 * v + k] - pos1[k]; }
    dist[v] = dot3(dif, normal);
    if (best < 0 || dist[v] < dist[best]) best = v;
  }
  if (best < 0 || dist[best] > margin) return 0;
  int cnt = 0;
  dist_out[0] = dist[best];
  for (int k = 0; k < 3; k++) pos_out[k] = pw[3 * best + k] - normal[k] * dist[best] * 0.5;
  cnt = 1;
  const double tol = 0.3 * m->geom_rbound[g2];
  for (int e = 0; e < MAXADJ && cnt < 3; e++) {
    const int v = m->hull_adj[(adr + best) * MAXADJ + e];
    if (v < 0) break;
    double d[3] = {pw[3 * v] - pw[3 * best], pw[3 * v + 1] - pw[3 * best + 1], pw[3 * v + 2] - pw[3 * best + 2]};
    if (norm3(d) < tol) continue;
    if (dist[v] > margin) continue;
    dist_out[cnt] = dist[v];
    for (int k = 0; k < 3; k++) pos_out[3 * cnt + k] = pw[3 * v + k] - normal[k] * dist[v] * 0.5;
    cnt++;
  }
  return cnt;
}

/* Box (g1) vs convex mesh (g2): analytic single contact, the counterpart of cylinder_box for the chassis plates against the maze
   blocks.  MuJoCo sends the pair to its general convex collider (one contact along the minimum-penetration direction).  Here: over
   the box's face normals n, the hull's support distance d(n) = min over vertices of (n . v - half); the face with the largest d is
   the contact face, contact if d <= margin.  The contact point is the weighted mean of the hull vertices within `blend` of the
   deepest one (weights fall linearly to zero), so that an edge or a face of the plate resting against the block gives its
   middle and the point moves continuously with the pose.  Normal from the box to the mesh.  Edge / corner axes are not tested. */
#define PLATE_BLEND 1e-3
static int box_convex(const OModel* m, int g2, const double* bpos, const double* bmat, const double* bsize, const double* pos2,
                      const double* mat2, double margin, double* dist_out, double* pos_out, double* normal_out) {
  const int adr = m->geom_hulladr[g2], nv = m->geom_hullnum[g2];
  double pl[MAXHULL * 3];
  for (int v = 0; v < nv; v++) {
    double t[3], rel[3];
    mulmatvec3(t, mat2, m->hull_vert + 3 * (adr + v));
    for (int k = 0; k < 3; k++) rel[k] = t[k] + pos2[k] - bpos[k];
    mulmatTvec3(pl + 3 * v, bmat, rel);       /* vertex in the box frame */
  }
  double best = -1e300; int bk = 0, bs = 1;
  for (int k = 0; k < 3; k++)
    for (int sg = -1; sg <= 1; sg += 2) {
      double dmin = 1e300;
      for (int v = 0; v < nv; v++) { double d = sg * pl[3 * v + k] - bsize[k]; if (d < dmin) dmin = d; }
      if (dmin > best) { best = dmin; bk = k; bs = sg; }
    }
  if (best > margin) return 0;
  double wsum = 0, p[3] = {0, 0, 0};
  for (int v = 0; v < nv; v++) {
    double d = bs * pl[3 * v + bk] - bsize[bk];
    double w = 1.0 - (d - best) / PLATE_BLEND;
    if (w <= 0) continue;
    wsum += w;
    for (int k = 0; k < 3; k++) p[k] += w * pl[3 * v + k];
  }
  for (int k = 0; k < 3; k++) p[k] /= wsum;
  p[bk] -= bs * best * 0.5;                    /* midway between the surfaces */
  double n[3] = {0, 0, 0}, pw[3], nw[3];
  n[bk] = bs;                                  /* from the box towards the mesh */
  mulmatvec3(pw, bmat, p);
  mulmatvec3(nw, bmat, n);
  for (int k = 0; k < 3; k++) { pos_out[k] = pw[k] + bpos[k]; normal_out[k] = nw[k]; }
  dist_out[0] = best;
  return 1;
}

static void mix_params(const OModel* m, OContact* con, int g1, int g2) {
  /* equal priority assumed (all geoms priority 0): condim max, friction max, solref/solimp solmix-weighted */
  con->dim = m->geom_condim[g1] > m->geom_condim[g2] ? m->geom_condim[g1] : m->geom_condim[g2];
  const double* f1 = m->geom_friction + 3 * g1; const double* f2 = m->geom_friction + 3 * g2;
  double fr[3];
  for (int k = 0; k < 3; k++) fr[k] = f1[k] > f2[k] ? f1[k] : f2[k];
  con->friction[0] = fr[0]; con->friction[1] = fr[0]; con->friction[2] = fr[1]; con->friction[3] = fr[2]; con->friction[4] = fr[2];
  double s1 = m->geom_solmix[g1], s2 = m->geom_solmix[g2], mix;
  if (s1 >= MINVAL && s2 >= MINVAL) mix = s1 / (s1 + s2);
  else if (s1 < MINVAL && s2 < MINVAL) mix = 0.5;
  else mix = s1 < MINVAL ? 0.0 : 1.0;
  const double* r1 = m->geom_solref + 2 * g1; const double* r2 = m->geom_solref + 2 * g2;
  if (r1[0] > 0 && r2[0] > 0) for (int k = 0; k < 2; k++) con->solref[k] = mix * r1[k] + (1 - mix) * r2[k];
  else for (int k = 0; k < 2; k++) con->solref[k] = r1[k] < r2[k] ? r1[k] : r2[k];
  for (int k = 0; k < 5; k++) con->solimp[k] = mix * m->geom_solimp[5 * g1 + k] + (1 - mix) * m->geom_solimp[5 * g2 + k];
}

static void add_contact(const OModel* m, OData* d, int g1, int g2, double dist, const double* pos, const double* normal,
                        double margin, double gap) {
  if (d->ncon >= MAXCON) return;
  OContact* c = d->contact + d->ncon++;
  memset(c, 0, sizeof *c);
  c->geom1 = g1; c->geom2 = g2; c->dist = dist;
  memcpy(c->pos, pos, 3 * sizeof(double));
  memcpy(c->frame, normal, 3 * sizeof(double));
  make_frame(c->frame);
  c->includemargin = margin - gap;
  mix_params(m, c, g1, g2);
  c->exclude = (dist >= c->includemargin);
}

static void o_collision(const OModel* m, OData* d) {
  d->ncon = 0;
  d->unsupported_contact = 0;
  for (int ga = 0; ga < m->ngeom; ga++)
    for (int gb = ga + 1; gb < m->ngeom; gb++) {
      int g1 = ga, g2 = gb;
      int b1 = m->geom_bodyid[g1], b2 = m->geom_bodyid[g2];
      int w1 = m->body_weldid[b1], w2 = m->body_weldid[b2];
      if (w1 == w2) continue;                         /* same (welded) body, incl. static-static */
      if (!((m->geom_contype[g1] & m->geom_conaffinity[g2]) || (m->geom_contype[g2] & m->geom_conaffinity[g1]))) continue;
      int p1 = m->body_weldid[m->body_parentid[w1]], p2 = m->body_weldid[m->body_parentid[w2]];
      if (w1 != 0 && w2 != 0 && (w1 == p2 || w2 == p1)) continue; /* parent-child filter */
      if (m->geom_type[g1] > m->geom_type[g2]) { int t = g1; g1 = g2; g2 = t; }
      int t1 = m->geom_type[g1], t2 = m->geom_type[g2];
      double margin = m->geom_margin[g1] > m->geom_margin[g2] ? m->geom_margin[g1] : m->geom_margin[g2];
      double gap = m->geom_gap[g1] > m->geom_gap[g2] ? m->geom_gap[g1] : m->geom_gap[g2];
      const double *pos1 = d->geom_xpos + 3 * g1, *mat1 = d->geom_xmat + 9 * g1, *pos2 = d->geom_xpos + 3 * g2,
                   *mat2 = d->geom_xmat + 9 * g2;
      if (t1 == G_PLANE && t2 == G_CYLINDER) {
        double dist[4], pos[12], normal[3] = {mat1[2], mat1[5], mat1[8]};
        int n = plane_cylinder(pos1, mat1, pos2, mat2, m->geom_size + 3 * g2, margin, dist, pos);
        for (int k = 0; k < n; k++) add_contact(m, d, g1, g2, dist[k], pos + 3 * k, normal, margin, gap);
      } else if (t1 == G_PLANE && t2 == G_MESH) {
        double dist[3], pos[9], normal[3] = {mat1[2], mat1[5], mat1[8]};
        int n = plane_convex(m, g2, pos1, mat1, pos2, mat2, margin, dist, pos);
        for (int k = 0; k < n; k++) add_contact(m, d, g1, g2, dist[k], pos + 3 * k, normal, margin, gap);
      } else if (t1 == G_BOX && t2 == G_MESH) {
        double dist[1], pos[3], normal[3];
        if (box_convex(m, g2, pos1, mat1, m->geom_size + 3 * g1, pos2, mat2, margin, dist, pos, normal))
          add_contact(m, d, g1, g2, dist[0], pos, normal, margin, gap);
      } else if (t1 == G_CYLINDER && t2 == G_BOX) {
        double dist[1], pos[3], normal[3];
        if (cylinder_box(pos1, mat1, m->geom_size + 3 * g1, pos2, mat2, m->geom_size + 3 * g2, margin, dist, pos, normal))
          add_contact(m, d, g1, g2, dist[0], pos, normal, margin, gap);
      } else if (t1 == G_CYLINDER && t2 == G_CYLINDER) {
        double dif[3];
        for (int k = 0; k < 3; k++) dif[k] = pos2[k] - pos1[k];
        double r1 = sqrt(m->geom_size[3 * g1] * m->geom_size[3 * g1] + m->geom_size[3 * g1 + 1] * m->geom_size[3 * g1 + 1]);
        double r2 = sqrt(m->geom_size[3 * g2] * m->geom_size[3 * g2] + m->geom_size[3 * g2 + 1] * m->geom_size[3 * g2 + 1]);
        if (norm3(dif) <= r1 + r2 + margin) d->unsupported_contact = 1; /* wheels never touch each other */
      } else {
        /* other pairs do not occur in the two models */
      }
    }
}

/* ------------------------------------------------------------------------- */
/*  B7 constraint rows                                                         */
/* ------------------------------------------------------------------------- */
static int add_row(const OModel* m, OData* d, const double* J, double pos, double margin, double frictionloss, int type,
                   int id, const double* solref, const double* solimp, double diagApprox) {
  if (d->nefc >= MAXEFC) return -1;
  int r = d->nefc++;
  memcpy(d->efc_J + r * MAXV, J, sizeof(double) * m->nv);
  d->efc_pos[r] = pos; d->efc_margin[r] = margin; d->efc_frictionloss[r] = frictionloss;
  d->efc_type[r] = type; d->efc_id[r] = id; d->efc_diagApprox[r] = diagApprox;
  memcpy(d->efc_solref + 2 * r, solref, 2 * sizeof(double));
  memcpy(d->efc_solimp + 5 * r, solimp, 5 * sizeof(double));
  return r;
}

static void get_impedance(const double* solimp_in, double pos, double margin, double* imp, double* impP) {
  double si[5];
  memcpy(si, solimp_in, sizeof si);
  si[0] = fmin(MAXIMP, fmax(MINIMP, si[0]));
  si[1] = fmin(MAXIMP, fmax(MINIMP, si[1]));
  si[2] = fmax(0, si[2]);
  si[3] = fmin(MAXIMP, fmax(MINIMP, si[3]));
  si[4] = fmax(1, si[4]);
  *impP = 0;
  if (si[0] == si[1] || si[2] <= MINVAL) { *imp = 0.5 * (si[0] + si[1]); return; }
  double x = (pos - margin) / si[2];
  if (x < 0) x = -x;
  if (x >= 1 || x <= 0) { *imp = (x >= 1) ? si[1] : si[0]; return; }
  double y;
  if (si[4] == 1) y = x;
  else if (x <= si[3]) y = pow(x, si[4]) / pow(si[3], si[4] - 1);
  else y = 1 - pow(1 - x, si[4]) / pow(1 - si[3], si[4] - 1);
  *imp = si[0] + y * (si[1] - si[0]);
}

static void o_make_constraint(const OModel* m, OData* d) {
  int nv = m->nv;
  double J[2 * MAXV];
  d->nefc = 0;
  /* equality: joint coupling  q1 - poly(q2) = 0 */
  for (int e = 0; e < m->neq; e++) {
    int j1 = m->eq_obj1id[e], j2 = m->eq_obj2id[e];
    const double* c = m->eq_data + 5 * e;
    memset(J, 0, sizeof(double) * nv);
    double pos1 = d->qpos[m->jnt_qposadr[j1]] - m->qpos0[m->jnt_qposadr[j1]], cpos, diag = m->dof_invweight0[m->jnt_dofadr[j1]];
    J[m->jnt_dofadr[j1]] = 1;
    if (j2 >= 0) {
      double dif = d->qpos[m->jnt_qposadr[j2]] - m->qpos0[m->jnt_qposadr[j2]];
      double ref = c[0] + dif * (c[1] + dif * (c[2] + dif * (c[3] + dif * c[4])));
      double deriv = c[1] + dif * (2 * c[2] + dif * (3 * c[3] + dif * 4 * c[4]));
      cpos = pos1 - ref;
      J[m->jnt_dofadr[j2]] = -deriv;
      diag += m->dof_invweight0[m->jnt_dofadr[j2]];
    } else cpos = pos1 - c[0];
    add_row(m, d, J, cpos, 0, 0, C_EQUALITY, e, m->eq_solref + 2 * e, m->eq_solimp + 5 * e, diag);
  }
  /* dof friction loss */
  for (int i = 0; i < nv; i++)
    if (m->dof_frictionloss[i] > 0) {
      memset(J, 0, sizeof(double) * nv);
      J[i] = 1;
      add_row(m, d, J, 0, 0, m->dof_frictionloss[i], C_FRICTION_DOF, i, m->dof_solref + 2 * i, m->dof_solimp + 5 * i,
              m->dof_invweight0[i]);
    }
  /* joint limits (hinge / slide) */
  for (int j = 0; j < m->njnt; j++)
    if (m->jnt_limited[j] && (m->jnt_type[j] == JNT_HINGE || m->jnt_type[j] == JNT_SLIDE)) {
      double value = d->qpos[m->jnt_qposadr[j]], margin = m->jnt_margin[j];
      for (int side = -1; side <= 1; side += 2) {
        double dist = side * (m->jnt_range[2 * j + (side + 1) / 2] - value);
        if (dist < margin) {
          memset(J, 0, sizeof(double) * nv);
          J[m->jnt_dofadr[j]] = -side;
          add_row(m, d, J, dist, margin, 0, C_LIMIT, j, m->jnt_solref + 2 * j, m->jnt_solimp + 5 * j,
                  m->dof_invweight0[m->jnt_dofadr[j]]);
        }
      }
    }
  /* contacts: pyramidal friction cone, 2*(dim-1) rows per contact */
  for (int ci = 0; ci < d->ncon; ci++) {
    OContact* con = d->contact + ci;
    if (con->exclude) continue;
    int b1 = m->geom_bodyid[con->geom1], b2 = m->geom_bodyid[con->geom2];
    double jp1[3 * MAXV], jp2[3 * MAXV], jac[3 * MAXV];
    o_jac(m, d, jp1, NULL, con->pos, b1);
    o_jac(m, d, jp2, NULL, con->pos, b2);
    for (int r = 0; r < 3; r++)
      for (int c = 0; c < nv; c++) {
        double s = 0;
        for (int k = 0; k < 3; k++) s += con->frame[3 * r + k] * (jp2[k * nv + c] - jp1[k * nv + c]);
        jac[r * nv + c] = s;
      }
    double tran = m->body_invweight0[2 * b1] + m->body_invweight0[2 * b2];
    double rot = m->body_invweight0[2 * b1 + 1] + m->body_invweight0[2 * b2 + 1];
    if (con->dim == 1) {
      add_row(m, d, jac, con->dist, con->includemargin, 0, C_CONTACT_PYR, ci, con->solref, con->solimp, tran);
      continue;
    }
    for (int k = 1; k < con->dim; k++) {
      double fri = con->friction[k - 1];
      double diag = tran + fri * fri * ((k - 1) < 2 ? tran : rot);
      /* rows beyond the 3 translational directions would need the rotational Jacobian (condim 4/6): not used */
      for (int sgn = 1; sgn >= -1; sgn -= 2) {
        for (int c = 0; c < nv; c++) J[c] = jac[c] + sgn * fri * jac[k * nv + c];
        add_row(m, d, J, con->dist, con->includemargin, 0, C_CONTACT_PYR, ci, con->solref, con->solimp, diag);
      }
    }
  }
  /* impedance, regulariser, reference acceleration */
  for (int r = 0; r < d->nefc; r++) {
    double imp, impP;
    const double* sr = d->efc_solref + 2 * r; const double* si = d->efc_solimp + 5 * r;
    get_impedance(si, d->efc_pos[r], d->efc_margin[r], &imp, &impP);
    double K, B;
    if (sr[0] > 0) {
      double tc = fmax(sr[0], 2 * m->opt_timestep); /* refsafe */
      double dmax = fmin(MAXIMP, fmax(MINIMP, si[1]));
      double dr = sr[1];
      K = 1 / fmax(MINVAL, dmax * dmax * tc * tc * dr * dr);
      B = 2 / fmax(MINVAL, dmax * tc);
    } else { /* direct stiffness / damping */
      double dmax = fmin(MAXIMP, fmax(MINIMP, si[1]));
      K = -sr[0] / fmax(MINVAL, dmax * dmax);
      B = -sr[1] / fmax(MINVAL, dmax);
    }
    if (d->efc_type[r] == C_FRICTION_DOF) K = 0;
    d->efc_KBIP[4 * r] = K; d->efc_KBIP[4 * r + 1] = B; d->efc_KBIP[4 * r + 2] = imp; d->efc_KBIP[4 * r + 3] = impP;
    d->efc_R[r] = fmax(MINVAL, (1 - imp) / imp * d->efc_diagApprox[r]);
  }
  /* friction-cone adjustment of R: all rows of a pyramidal contact share Rpy = 2 mu^2 R0 */
  for (int r = 0; r < d->nefc; r++)
    if (d->efc_type[r] == C_CONTACT_PYR) {
      OContact* con = d->contact + d->efc_id[r];
      if (con->dim == 1) continue;
      int nrow = 2 * (con->dim - 1);
      double R1 = d->efc_R[r] / fmax(MINVAL, m->opt_impratio);
      con->mu = con->friction[0] * sqrt(R1 / d->efc_R[r]);
      double Rpy = 2 * con->mu * con->mu * d->efc_R[r];
      for (int k = 0; k < nrow; k++) d->efc_R[r + k] = Rpy;
      r += nrow - 1;
    }
  for (int r = 0; r < d->nefc; r++) {
    d->efc_D[r] = 1 / d->efc_R[r];
    double vel = 0;
    for (int c = 0; c < nv; c++) vel += d->efc_J[r * MAXV + c] * d->qvel[c];
    d->efc_aref[r] = -d->efc_KBIP[4 * r + 1] * vel - d->efc_KBIP[4 * r] * d->efc_KBIP[4 * r + 2] * (d->efc_pos[r] - d->efc_margin[r]);
  }
}

/* ------------------------------------------------------------------------- */
/*  B9 rays / sensors                                                          */
/* ------------------------------------------------------------------------- */
static double ray_geom(const double* gpos, const double* gmat, const double* size, int type, const double* pnt,
                       const double* vec) {
  double dif[3], lp[3], lv[3];
  for (int k = 0; k < 3; k++) dif[k] = pnt[k] - gpos[k];
  mulmatTvec3(lp, gmat, dif);
  mulmatTvec3(lv, gmat, vec);
  if (type == G_PLANE) {
    if (lv[2] > -MINVAL) return -1;
    double x = -lp[2] / lv[2];
    if (x < 0) return -1;
    double p0 = lp[0] + x * lv[0], p1 = lp[1] + x * lv[1];
    if ((size[0] <= 0 || fabs(p0) <= size[0]) && (size[1] <= 0 || fabs(p1) <= size[1])) return x;
    return -1;
  }
  if (type == G_BOX) {
    double best = -1;
    for (int i = 0; i < 3; i++) {
      if (fabs(lv[i]) <= MINVAL) continue;
      for (int side = -1; side <= 1; side += 2) {
        double sol = (side * size[i] - lp[i]) / lv[i];
        if (sol < 0) continue;
        int a = (i + 1) % 3, b = (i + 2) % 3;
        double pa = lp[a] + sol * lv[a], pb = lp[b] + sol * lv[b];
        if (fabs(pa) <= size[a] && fabs(pb) <= size[b] && (best < 0 || sol < best)) best = sol;
      }
    }
    return best;
  }
  if (type == G_CYLINDER) {
    double best = -1;
    /* caps */
    if (fabs(lv[2]) > MINVAL)
      for (int side = -1; side <= 1; side += 2) {
        double sol = (side * size[1] - lp[2]) / lv[2];
        if (sol < 0) continue;
        double px = lp[0] + sol * lv[0], py = lp[1] + sol * lv[1];
        if (px * px + py * py <= size[0] * size[0] && (best < 0 || sol < best)) best = sol;
      }
    /* side */
    double a = lv[0] * lv[0] + lv[1] * lv[1], b = lv[0] * lp[0] + lv[1] * lp[1], c = lp[0] * lp[0] + lp[1] * lp[1] - size[0] * size[0];
    if (a > MINVAL) {
      double det = b * b - a * c;
      if (det >= 0) {
        double sq = sqrt(det);
        double sols[2] = {(-b - sq) / a, (-b + sq) / a};
        for (int k = 0; k < 2; k++)
          if (sols[k] >= 0 && fabs(lp[2] + sols[k] * lv[2]) <= size[1] && (best < 0 || sols[k] < best)) best = sols[k];
      }
    }
    return best;
  }
  return -1; /* mesh: rays lie in the plane z=+0.03 of the chassis frame, above every robot geom */
}

static void o_sensor_pos(const OModel* m, OData* d) {
  for (int s = 0; s < m->nsensor; s++) {
    int adr = m->sensor_adr[s], obj = m->sensor_objid[s];
    if (m->sensor_type[s] == 0) d->sensordata[adr] = d->qpos[m->jnt_qposadr[obj]];
    else if (m->sensor_type[s] == 2) {
      const double* pnt = d->site_xpos + 3 * obj; const double* sm = d->site_xmat + 9 * obj;
      double vec[3] = {sm[2], sm[5], sm[8]};
      int bex = m->site_bodyid[obj];
      double best = -1;
      for (int g = 0; g < m->ngeom; g++) {
        if (m->geom_bodyid[g] == bex) continue;
        if (m->geom_alpha[g] == 0) continue;
        double x = ray_geom(d->geom_xpos + 3 * g, d->geom_xmat + 9 * g, m->geom_size + 3 * g, m->geom_type[g], pnt, vec);
        if (x >= 0 && (best < 0 || x < best)) best = x;
      }
      if (m->sensor_cutoff[s] > 0 && best > m->sensor_cutoff[s]) best = m->sensor_cutoff[s];
      d->sensordata[adr] = best;
    }
  }
}
static void o_sensor_vel(const OModel* m, OData* d) {
  for (int s = 0; s < m->nsensor; s++)
    if (m->sensor_type[s] == 1) {
      double v = d->qvel[m->jnt_dofadr[m->sensor_objid[s]]], c = m->sensor_cutoff[s];
      if (c > 0) v = fmin(c, fmax(-c, v));
      d->sensordata[m->sensor_adr[s]] = v;
    }
}

/* ------------------------------------------------------------------------- */
/*  B10 velocity stage: comVel, passive, RNE bias                              */
/* ------------------------------------------------------------------------- */
static void o_comvel(const OModel* m, OData* d) {
  memset(d->cvel, 0, 6 * sizeof(double));
  for (int i = 1; i < m->nbody; i++) {
    double cvel[6];
    memcpy(cvel, d->cvel + 6 * m->body_parentid[i], sizeof cvel);
    int ja = m->body_jntadr[i], jn = m->body_jntnum[i];
    for (int j = ja; j < ja + jn; j++) {
      int da = m->jnt_dofadr[j];
      if (m->jnt_type[j] == JNT_FREE) {
        memset(d->cdof_dot + 6 * da, 0, 18 * sizeof(double));
        for (int k = 0; k < 3; k++) for (int c = 0; c < 6; c++) cvel[c] += d->cdof[6 * (da + k) + c] * d->qvel[da + k];
        for (int k = 3; k < 6; k++) cross_motion(d->cdof_dot + 6 * (da + k), cvel, d->cdof + 6 * (da + k));
        for (int k = 3; k < 6; k++) for (int c = 0; c < 6; c++) cvel[c] += d->cdof[6 * (da + k) + c] * d->qvel[da + k];
      } else {
        cross_motion(d->cdof_dot + 6 * da, cvel, d->cdof + 6 * da);
        for (int c = 0; c < 6; c++) cvel[c] += d->cdof[6 * da + c] * d->qvel[da];
      }
    }
    memcpy(d->cvel + 6 * i, cvel, sizeof cvel);
  }
}

static void o_rne_bias(const OModel* m, OData* d) {
  double cacc[MAXB * 6], cfrc[MAXB * 6];
  memset(cacc, 0, sizeof cacc);
  for (int k = 0; k < 3; k++) cacc[3 + k] = -m->opt_gravity[k];
  memset(cfrc, 0, 6 * sizeof(double));
  for (int i = 1; i < m->nbody; i++) {
    int da = m->body_dofadr[i], dn = m->body_dofnum[i];
    double* a = cacc + 6 * i;
    memcpy(a, cacc + 6 * m->body_parentid[i], 6 * sizeof(double));
    for (int k = 0; k < dn; k++) for (int c = 0; c < 6; c++) a[c] += d->cdof_dot[6 * (da + k) + c] * d->qvel[da + k];
    double t0[6], t1[6], t2[6];
    mul_inert_vec(t0, d->cinert + 10 * i, a);
    mul_inert_vec(t1, d->cinert + 10 * i, d->cvel + 6 * i);
    cross_force(t2, d->cvel + 6 * i, t1);
    for (int c = 0; c < 6; c++) cfrc[6 * i + c] = t0[c] + t2[c];
  }
  for (int i = m->nbody - 1; i > 0; i--) {
    int p = m->body_parentid[i];
    if (p > 0) for (int c = 0; c < 6; c++) cfrc[6 * p + c] += cfrc[6 * i + c];
  }
  for (int v = 0; v < m->nv; v++) {
    double s = 0;
    for (int c = 0; c < 6; c++) s += d->cdof[6 * v + c] * cfrc[6 * m->dof_bodyid[v] + c];
    d->qfrc_bias[v] = s;
  }
}

static void o_fwd_velocity(const OModel* m, OData* d) {
  o_comvel(m, d);
  for (int v = 0; v < m->nv; v++) d->qfrc_passive[v] = -m->dof_damping[v] * d->qvel[v];
  o_rne_bias(m, d);
}

/* ------------------------------------------------------------------------- */
/*  B12/B13 actuation and smooth acceleration                                  */
/* ------------------------------------------------------------------------- */
static void o_fwd_actuation(const OModel* m, OData* d) {
  memset(d->qfrc_actuator, 0, sizeof d->qfrc_actuator);
  for (int u = 0; u < m->nu; u++) {
    int j = m->actuator_trnid[u], dof = m->jnt_dofadr[j];
    double gear = m->actuator_gear[u];
    double length = gear * d->qpos[m->jnt_qposadr[j]], velocity = gear * d->qvel[dof];
    double ctrl = d->ctrl[u];
    if (m->actuator_ctrllimited[u]) ctrl = fmin(m->actuator_ctrlrange[2 * u + 1], fmax(m->actuator_ctrlrange[2 * u], ctrl));
    double f = m->actuator_gainprm[u] * ctrl + m->actuator_biasprm[3 * u] + m->actuator_biasprm[3 * u + 1] * length +
               m->actuator_biasprm[3 * u + 2] * velocity;
    if (m->actuator_forcelimited[u]) f = fmin(m->actuator_forcerange[2 * u + 1], fmax(m->actuator_forcerange[2 * u], f));
    d->actuator_force[u] = f;
    d->qfrc_actuator[dof] += gear * f;
  }
}
static void o_fwd_acceleration(const OModel* m, OData* d) {
  for (int v = 0; v < m->nv; v++) d->qfrc_smooth[v] = d->qfrc_passive[v] - d->qfrc_bias[v] + d->qfrc_actuator[v];
  chol_solve(d->qacc_smooth, d->qL, d->qfrc_smooth, m->nv, MAXV);
}

/* ------------------------------------------------------------------------- */
/*  B14 constraint solver: Newton on the convex primal cost, exact line search */
/* ------------------------------------------------------------------------- */
typedef struct {
  double Ma[MAXV], jar[MAXEFC], force[MAXEFC];
  int quad[MAXEFC]; /* row currently in its quadratic zone (contributes D to the Hessian) */
  double cost;
} SolState;

/* per-row cost: value, force = -ds/dx, and whether the quadratic zone is active */
static double row_cost(const OData* d, int r, double x, double* force, int* quad) {
  double D = d->efc_D[r], R = d->efc_R[r];
  switch (d->efc_type[r]) {
    case C_EQUALITY:
      *force = -D * x; *quad = 1; return 0.5 * D * x * x;
    case C_FRICTION_DOF: {
      double f = d->efc_frictionloss[r], rf = R * f;
      if (x <= -rf) { *force = f; *quad = 0; return f * (-0.5 * rf - x); }
      if (x >= rf) { *force = -f; *quad = 0; return f * (-0.5 * rf + x); }
      *force = -D * x; *quad = 1; return 0.5 * D * x * x;
    }
    default: /* limit, pyramidal contact row */
      if (x < 0) { *force = -D * x; *quad = 1; return 0.5 * D * x * x; }
      *force = 0; *quad = 0; return 0;
  }
}

static void sol_update(const OModel* m, const OData* d, const double* qacc, SolState* s) {
  int nv = m->nv;
  double c = 0;
  for (int r = 0; r < d->nefc; r++) c += row_cost(d, r, s->jar[r], &s->force[r], &s->quad[r]);
  double g = 0;
  for (int v = 0; v < nv; v++) g += (s->Ma[v] - d->qfrc_smooth[v]) * (qacc[v] - d->qacc_smooth[v]);
  s->cost = c + 0.5 * g;
}

static void sol_init(const OModel* m, const OData* d, const double* qacc, SolState* s) {
  int nv = m->nv;
  for (int i = 0; i < nv; i++) { double t = 0; for (int j = 0; j < nv; j++) t += d->qM[i * MAXV + j] * qacc[j]; s->Ma[i] = t; }
  for (int r = 0; r < d->nefc; r++) {
    double t = -d->efc_aref[r];
    for (int j = 0; j < nv; j++) t += d->efc_J[r * MAXV + j] * qacc[j];
    s->jar[r] = t;
  }
  sol_update(m, d, qacc, s);
}

static void o_fwd_constraint(const OModel* m, OData* d) {
  int nv = m->nv, nefc = d->nefc;
  d->solver_niter = 0; d->solver_lsiter = 0;
  memset(d->qfrc_constraint, 0, sizeof d->qfrc_constraint);
  if (nefc == 0) { memcpy(d->qacc, d->qacc_smooth, sizeof(double) * nv); return; }
  static SolState s, s2; /* single-threaded oracle */
  /* warm start: previous qacc if its cost beats qacc_smooth */
  sol_init(m, d, d->qacc_warmstart, &s);
  sol_init(m, d, d->qacc_smooth, &s2);
  double qacc[MAXV];
  if (s.cost > s2.cost || !(s.cost == s.cost)) { memcpy(qacc, d->qacc_smooth, sizeof(double) * nv); s = s2; }
  else memcpy(qacc, d->qacc_warmstart, sizeof(double) * nv);

  double scale = 1.0 / (m->stat_meaninertia * (nv > 1 ? nv : 1));
  double grad[MAXV], search[MAXV], H[MAXV * MAXV], L[MAXV * MAXV], Mv[MAXV], jv[MAXEFC];
  double improvement = 1e300;
  for (int iter = 0; iter < m->opt_iterations; iter++) {
    /* gradient  g = M a - qfrc_smooth - J^T f */
    for (int v = 0; v < nv; v++) grad[v] = s.Ma[v] - d->qfrc_smooth[v];
    for (int r = 0; r < nefc; r++) if (s.force[r] != 0) for (int v = 0; v < nv; v++) grad[v] -= d->efc_J[r * MAXV + v] * s.force[r];
    double gn = 0;
    for (int v = 0; v < nv; v++) gn += grad[v] * grad[v];
    gn = sqrt(gn);
    d->solver_gradnorm = scale * gn;
    if (iter > 0 && (improvement < m->opt_tolerance || scale * gn < m->opt_tolerance)) break;
    /* Newton direction: (M + J^T D_quad J) search = -g */
    memcpy(H, d->qM, sizeof H);
    for (int r = 0; r < nefc; r++)
      if (s.quad[r]) {
        const double* Jr = d->efc_J + r * MAXV; double D = d->efc_D[r];
        for (int i = 0; i < nv; i++) if (Jr[i] != 0) for (int j = 0; j < nv; j++) H[i * MAXV + j] += D * Jr[i] * Jr[j];
      }
    chol_factor(L, H, nv, MAXV);
    chol_solve(search, L, grad, nv, MAXV);
    for (int v = 0; v < nv; v++) search[v] = -search[v];
    /* exact line search on the piecewise-quadratic, convex 1-D cost (safeguarded Newton on f') */
    for (int i = 0; i < nv; i++) { double t = 0; for (int j = 0; j < nv; j++) t += d->qM[i * MAXV + j] * search[j]; Mv[i] = t; }
    for (int r = 0; r < nefc; r++) { double t = 0; for (int j = 0; j < nv; j++) t += d->efc_J[r * MAXV + j] * search[j]; jv[r] = t; }
    double sMs = 0, sg = 0;
    for (int v = 0; v < nv; v++) { sMs += search[v] * Mv[v]; sg += search[v] * (s.Ma[v] - d->qfrc_smooth[v]); }
    if (sMs < MINVAL) break;
    double alpha = 0, lo = 0, hi = -1, d1_0 = 0; /* hi<0: no upper bracket yet */
    for (int ls = 0; ls < m->opt_ls_iterations; ls++) {
      double d1 = alpha * sMs + sg, d2 = sMs;
      for (int r = 0; r < nefc; r++) {
        double f; int q;
        row_cost(d, r, s.jar[r] + alpha * jv[r], &f, &q);
        d1 -= f * jv[r];
        if (q) d2 += d->efc_D[r] * jv[r] * jv[r];
      }
      d->solver_lsiter++;
      if (ls == 0) d1_0 = fabs(d1);
      if (fabs(d1) <= 1e-13 * d1_0 || d1_0 == 0) break;
      if (d1 < 0) lo = alpha; else hi = alpha;
      double an = alpha - d1 / d2;
      if (hi >= 0 && (an <= lo || an >= hi)) an = 0.5 * (lo + hi);
      if (an == alpha) break;
      alpha = an;
    }
    if (alpha == 0) break;
    double oldcost = s.cost;
    for (int v = 0; v < nv; v++) { qacc[v] += alpha * search[v]; s.Ma[v] += alpha * Mv[v]; }
    for (int r = 0; r < nefc; r++) s.jar[r] += alpha * jv[r];
    sol_update(m, d, qacc, &s);
    d->solver_niter++;
    improvement = scale * (oldcost - s.cost);
  }
  memcpy(d->qacc, qacc, sizeof(double) * nv);
  memcpy(d->efc_force, s.force, sizeof(double) * nefc);
  d->solver_cost = s.cost;
  for (int r = 0; r < nefc; r++) if (s.force[r] != 0) for (int v = 0; v < nv; v++) d->qfrc_constraint[v] += d->efc_J[r * MAXV + v] * s.force[r];
}

/* ------------------------------------------------------------------------- */
/*  B16 Euler with implicit joint damping                                      */
/* ------------------------------------------------------------------------- */
static void o_euler(const OModel* m, OData* d) {
  int nv = m->nv;
  double h = m->opt_timestep, qacc[MAXV];
  int damped = 0;
  for (int v = 0; v < nv; v++) if (m->dof_damping[v] > 0) damped = 1;
  if (!damped) memcpy(qacc, d->qacc, sizeof(double) * nv);
  else {
    double H[MAXV * MAXV], L[MAXV * MAXV], f[MAXV];
    memcpy(H, d->qM, sizeof H);
    for (int v = 0; v < nv; v++) { H[v * MAXV + v] += h * m->dof_damping[v]; f[v] = d->qfrc_smooth[v] + d->qfrc_constraint[v]; }
    chol_factor(L, H, nv, MAXV);
    chol_solve(qacc, L, f, nv, MAXV);
  }
  for (int v = 0; v < nv; v++) d->qvel[v] += h * qacc[v];
  for (int j = 0; j < m->njnt; j++) {
    int qa = m->jnt_qposadr[j], da = m->jnt_dofadr[j];
    if (m->jnt_type[j] == JNT_FREE) {
      for (int k = 0; k < 3; k++) d->qpos[qa + k] += h * d->qvel[da + k];
      double w[3] = {d->qvel[da + 3], d->qvel[da + 4], d->qvel[da + 5]}, qr[4];
      double ang = h * normalize3(w);
      axisangle2quat(qr, w, ang);
      normalize4(d->qpos + qa + 3);
      mulquat(d->qpos + qa + 3, d->qpos + qa + 3, qr);
    } else d->qpos[qa] += h * d->qvel[da];
  }
  d->time += h;
}

/* ------------------------------------------------------------------------- */
/*  public entry points (ctypes)                                               */
/* ------------------------------------------------------------------------- */
void orc_forward(const OModel* m, OData* d) {
  o_kinematics(m, d);
  o_compos(m, d);
  o_crb(m, d);
  o_collision(m, d);
  o_make_constraint(m, d);
  o_sensor_pos(m, d);
  o_fwd_velocity(m, d);
  o_sensor_vel(m, d);
  o_fwd_actuation(m, d);
  o_fwd_acceleration(m, d);
  o_fwd_constraint(m, d);
  memcpy(d->qacc_warmstart, d->qacc, sizeof(double) * m->nv);
}
void orc_reset(const OModel* m, OData* d);
/* mju_isBad: NaN or |x| > mjMAXVAL */
static int is_bad(double x) { return !(fabs(x) <= 1e10); }
/* mj_step: mj_checkPos, mj_checkVel -> mj_resetData on a bad number; mj_forward; mj_checkAcc -> mj_resetData +
   mj_forward; integrate.  `nbad` counts the resets (MuJoCo's warning counters). */
void orc_step(const OModel* m, OData* d) {
  int bad = 0, nbad = d->nbad;
  for (int i = 0; i < m->nq; i++) bad |= is_bad(d->qpos[i]);
  for (int i = 0; i < m->nv; i++) bad |= is_bad(d->qvel[i]);
  if (bad) { orc_reset(m, d); d->nbad = ++nbad; }
  orc_forward(m, d);
  bad = 0;
  for (int i = 0; i < m->nv; i++) bad |= is_bad(d->qacc[i]);
  if (bad) { orc_reset(m, d); d->nbad = ++nbad; orc_forward(m, d); }
  o_euler(m, d);
}
void orc_step_n(const OModel* m, OData* d, int n) { for (int i = 0; i < n; i++) orc_step(m, d); }
void orc_reset(const OModel* m, OData* d) {
  memset(d, 0, sizeof *d);
  memcpy(d->qpos, m->qpos0, sizeof(double) * m->nq);
}
OModel* orc_model_new(void) { return (OModel*)calloc(1, sizeof(OModel)); }
OData* orc_data_new(void) { return (OData*)calloc(1, sizeof(OData)); }
void orc_free(void* p) { free(p); }

/* name -> (offset, count, is_int) tables so the Python side needs no struct mirror */
typedef struct { const char* name; size_t off; int count; int is_int; } Field;
#define MF(f, isint) {#f, offsetof(OModel, f), (int)(sizeof(((OModel*)0)->f) / (isint ? sizeof(int) : sizeof(double))), isint}
static const Field model_fields[] = {
    MF(nq, 1), MF(nv, 1), MF(nu, 1), MF(nbody, 1), MF(njnt, 1), MF(ngeom, 1), MF(nsite, 1), MF(nsensor, 1), MF(neq, 1),
    MF(nsensordata, 1), MF(nhullvert, 1), MF(opt_timestep, 0), MF(opt_gravity, 0), MF(opt_impratio, 0), MF(opt_tolerance, 0),
    MF(opt_ls_tolerance, 0), MF(opt_iterations, 1), MF(opt_ls_iterations, 1), MF(stat_meaninertia, 0),
    MF(body_parentid, 1), MF(body_rootid, 1), MF(body_weldid, 1), MF(body_jntnum, 1), MF(body_jntadr, 1), MF(body_dofnum, 1),
    MF(body_dofadr, 1), MF(body_pos, 0), MF(body_quat, 0), MF(body_ipos, 0), MF(body_iquat, 0), MF(body_mass, 0),
    MF(body_inertia, 0), MF(body_invweight0, 0), MF(jnt_type, 1), MF(jnt_qposadr, 1), MF(jnt_dofadr, 1), MF(jnt_bodyid, 1),
    MF(jnt_limited, 1), MF(jnt_pos, 0), MF(jnt_axis, 0), MF(jnt_range, 0), MF(jnt_margin, 0), MF(jnt_solref, 0),
    MF(jnt_solimp, 0), MF(qpos0, 0), MF(dof_bodyid, 1), MF(dof_jntid, 1), MF(dof_parentid, 1), MF(dof_armature, 0),
    MF(dof_damping, 0), MF(dof_frictionloss, 0), MF(dof_invweight0, 0), MF(dof_solref, 0), MF(dof_solimp, 0),
    MF(geom_type, 1), MF(geom_bodyid, 1), MF(geom_contype, 1), MF(geom_conaffinity, 1), MF(geom_condim, 1),
    MF(geom_priority, 1), MF(geom_hulladr, 1), MF(geom_hullnum, 1), MF(geom_size, 0), MF(geom_pos, 0), MF(geom_quat, 0),
    MF(geom_friction, 0), MF(geom_solref, 0), MF(geom_solimp, 0), MF(geom_solmix, 0), MF(geom_margin, 0), MF(geom_gap, 0),
    MF(geom_alpha, 0), MF(hull_vert, 0), MF(hull_adj, 1), MF(geom_rbound, 0), MF(site_bodyid, 1), MF(site_pos, 0), MF(site_quat, 0), MF(sensor_type, 1),
    MF(sensor_objid, 1), MF(sensor_adr, 1), MF(sensor_cutoff, 0), MF(eq_obj1id, 1), MF(eq_obj2id, 1), MF(eq_data, 0),
    MF(eq_solref, 0), MF(eq_solimp, 0), MF(actuator_trnid, 1), MF(actuator_ctrllimited, 1), MF(actuator_forcelimited, 1),
    MF(actuator_gear, 0), MF(actuator_gainprm, 0), MF(actuator_biasprm, 0), MF(actuator_ctrlrange, 0),
    MF(actuator_forcerange, 0), {NULL, 0, 0, 0}};
#define DF(f, isint) {#f, offsetof(OData, f), (int)(sizeof(((OData*)0)->f) / (isint ? sizeof(int) : sizeof(double))), isint}
static const Field data_fields[] = {
    DF(time, 0), DF(qpos, 0), DF(qvel, 0), DF(ctrl, 0), DF(qacc_warmstart, 0), DF(xpos, 0), DF(xquat, 0), DF(xmat, 0),
    DF(xipos, 0), DF(ximat, 0), DF(xanchor, 0), DF(xaxis, 0), DF(geom_xpos, 0), DF(geom_xmat, 0), DF(site_xpos, 0),
    DF(site_xmat, 0), DF(subtree_com, 0), DF(cinert, 0), DF(crb, 0), DF(cdof, 0), DF(cdof_dot, 0), DF(cvel, 0), DF(qM, 0),
    DF(qfrc_bias, 0), DF(qfrc_passive, 0), DF(qfrc_actuator, 0), DF(actuator_force, 0), DF(qfrc_smooth, 0),
    DF(qacc_smooth, 0), DF(qacc, 0), DF(qfrc_constraint, 0), DF(ncon, 1), DF(nefc, 1), DF(unsupported_contact, 1),
    DF(efc_type, 1), DF(efc_id, 1), DF(efc_J, 0), DF(efc_pos, 0), DF(efc_margin, 0), DF(efc_frictionloss, 0),
    DF(efc_diagApprox, 0), DF(efc_R, 0), DF(efc_D, 0), DF(efc_aref, 0), DF(efc_force, 0), DF(sensordata, 0),
    DF(solver_niter, 1), DF(solver_lsiter, 1), DF(nbad, 1), DF(solver_cost, 0), DF(solver_gradnorm, 0), {NULL, 0, 0, 0}};

static const Field* find_field(const Field* t, const char* name) {
  for (; t->name; t++) if (!strcmp(t->name, name)) return t;
  return NULL;
}
/* returns element capacity, or -1 if unknown; *is_int tells the element type */
int orc_model_field(OModel* m, const char* name, void** ptr, int* is_int) {
  const Field* f = find_field(model_fields, name);
  if (!f) return -1;
  *ptr = (char*)m + f->off; *is_int = f->is_int;
  return f->count;
}
int orc_data_field(OData* d, const char* name, void** ptr, int* is_int) {
  const Field* f = find_field(data_fields, name);
  if (!f) return -1;
  *ptr = (char*)d + f->off; *is_int = f->is_int;
  return f->count;
}
int orc_maxv(void) { return MAXV; }
/* contact accessor: fills out[0..17] = geom1, geom2, dim, exclude, dist, pos[3], frame[9], mu */
int orc_contact(const OData* d, int i, double* out) {
  if (i < 0 || i >= d->ncon) return -1;
  const OContact* c = d->contact + i;
  out[0] = c->geom1; out[1] = c->geom2; out[2] = c->dim; out[3] = c->exclude; out[4] = c->dist;
  memcpy(out + 5, c->pos, 3 * sizeof(double));
  memcpy(out + 8, c->frame, 9 * sizeof(double));
  out[17] = c->mu;
  return 0;
}

/* ------------------------------------------------------------------------- */
/*  CPU baseline loop: the reference's per-step work, restated in C            */
/*  (action -> BicycleController (src/core/controller.py:98-140) -> mj_step),  */
/*  used by bench.py's cpu_baseline / --impl reference legs only.              */
/* ------------------------------------------------------------------------- */
static void bicycle_ctrl_c(double v, double omega, double* ctrl) {
  const double eps = 1e-5, L = 0.20, Tw = 0.174, rw = 0.0325;
  double delta;
  if (fabs(omega) < 1e-6) delta = 0;
  else {
    double sgn = omega > 0 ? 1.0 : (omega < 0 ? -1.0 : 0.0);
    delta = atan((L * omega) / (fabs(v) > eps ? v : sgn * eps));
  }
  const double lim = 0.6108652381980153;
  delta = delta < -lim ? -lim : (delta > lim ? lim : delta);
  double vl, vr;
  if (fabs(delta) < 1e-6) vl = vr = v;
  else {
    double tn = tan(delta), R = fabs(tn) > eps ? L / tn : INFINITY;
    double ot = fabs(R) > eps ? v / R : 0.0;
    vl = ot * (R - Tw / 2); vr = ot * (R + Tw / 2);
  }
  ctrl[0] = delta < -0.61 ? -0.61 : (delta > 0.61 ? 0.61 : delta);
  double wl = vl / rw, wr = vr / rw;
  ctrl[1] = wl < -50 ? -50 : (wl > 50 ? 50 : wl);
  ctrl[2] = wr < -50 ? -50 : (wr > 50 ? 50 : wr);
  if (!(fabs(ctrl[0]) <= 1e10 && fabs(ctrl[1]) <= 1e10 && fabs(ctrl[2]) <= 1e10)) ctrl[0] = ctrl[1] = ctrl[2] = 0;
}
/* n_steps env steps of frame_skip substeps with U(-1,1) actions (xorshift), episodes of max_steps; returns substeps done */
long orc_rollout(const OModel* m, OData* d, long n_steps, int frame_skip, int max_steps, unsigned long long seed, const double* spawn_qpos) {
  unsigned long long s = seed * 2685821657736338717ULL + 1442695040888963407ULL;
  long sub = 0;
  int ep = 0;
  for (long i = 0; i < n_steps; i++) {
    if (ep == 0) { orc_reset(m, d); memcpy(d->qpos, spawn_qpos, sizeof(double) * m->nq); }
    double a[2];
    for (int k = 0; k < 2; k++) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; a[k] = (double)(float)(2.0 * ((s >> 11) * (1.0 / 9007199254740992.0)) - 1.0); }
    bicycle_ctrl_c(a[0], a[1], d->ctrl);
    for (int f = 0; f < frame_skip; f++) { orc_step(m, d); sub++; }
    if (++ep >= max_steps) ep = 0;
  }
  return sub;
}
