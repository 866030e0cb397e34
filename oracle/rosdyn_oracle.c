/*
 * rosdyn_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C CPU restatement of the hot path of rosdyn_core's `rosdyn::Chain`, operation by operation in the
 * order of the reference's *direct* evaluation paths.  It is the parity oracle for the CUDA engine and the
 * "port" CPU baseline of bench.py.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may link or call it; the product (rosdyn_b200/) never does.
 *
 * PINNING.  The reference's own tests hold nothing to pin against (rosdyn_core/test/test.cpp has no EXPECT/ASSERT, ships no golden
 * vectors and its URDF is external), and Eigen3 / urdfdom / roscpp are not installed here.
 * This file is therefore pinned against OUTPUTS OF THE REFERENCE ITSELF: oracle/_ref/librosdyn_ref.so is the reference's own
 * rosdyn::Chain (primitives.h, internal/primitives_impl.h, spacevect_algebra.h, urdf_parser.h compiled where they lie under
 * /root/reference by oracle/Makefile `ref`), with the absent third-party headers replaced by the stand-ins of oracle/shim/
 * (mini_eigen.h: dense products as plain loops; urdf structs; ros logging dropped).  tests/test_reference_build.py compares every
 * output of this file with that build live (<= 1e-12) and with its committed outputs tests/golden/ref_*.npz (generator
 * tests/golden/make_golden_ref.py).  What that build does NOT pin is Eigen's own rounding order inside fixed-size products
 * (Eigen3 is un-vendored and unpinned by the reference, CMakeLists.txt:31): a few ulp, far inside the 1e-10 acceptance bound.
 * Also kept: (1) the independent numpy transcription oracle/numpy_transcription.py -> tests/golden/<chain>.npz, (2) the algebraic
 * invariants of SURVEY.md section 4 and the UR10 zero-pose known answer.
 * Third-party arithmetic restated here: Eigen3 fixed-size products/cross/transposes (version unpinned by the
 * reference, CMakeLists.txt:31) and libm sin/cos.
 *
 * Citations: SA.h = rosdyn_core/include/rosdyn_core/spacevect_algebra.h,
 *            PI.h = rosdyn_core/include/rosdyn_core/internal/primitives_impl.h.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../include/rosdyn_b200.h" /* descriptor structs only (the interface both sides share) */

#define OR_MAXJ RDB_MAX_JOINTS
#define OR_MAXL (RDB_MAX_JOINTS + 1)

typedef struct
{
  int type;
  int input_index;
  double T_pj[16];  /* 4x4 row-major, PI.h:54 */
  double R_pj[9];   /* PI.h:68 */
  double axis_j[3]; /* PI.h:55-59 */
  double axis_p[3]; /* PI.h:69 */
  double K[9];      /* skew(axis_j)    PI.h:62 */
  double K2[9];     /* K*K             PI.h:63 */
  double screw_p[6]; /* PI.h:25-35 */
} or_joint;

typedef struct
{
  double mass;
  double cog[3];
  double I_cc[36];     /* spatial inertia about the link origin, SA.h:232-239 */
  double E[10][36];    /* m_Inertia_cc_single_term, PI.h:342-396 */
  double nominal[10];  /* PI.h:399-417 */
} or_link;

typedef struct oracle_chain
{
  int nJ, nL, n_in;
  double g[3];
  or_joint joint[OR_MAXJ];
  or_link link[OR_MAXL];
} oracle_chain;

/* ------------------------------------------------------------------ small dense helpers (Eigen stand-ins) */
static void mat3_mul(const double* a, const double* b, double* c)
{
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++)
    {
      double s = 0;
      for (int k = 0; k < 3; k++) s += a[3 * i + k] * b[3 * k + j];
      c[3 * i + j] = s;
    }
}
static void mat3_vec(const double* a, const double* x, double* y)
{
  for (int i = 0; i < 3; i++) y[i] = a[3 * i] * x[0] + a[3 * i + 1] * x[1] + a[3 * i + 2] * x[2];
}
static void mat3T_vec(const double* a, const double* x, double* y)
{
  for (int i = 0; i < 3; i++) y[i] = a[i] * x[0] + a[3 + i] * x[1] + a[6 + i] * x[2];
}
static void mat4_mul(const double* a, const double* b, double* c)
{
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++)
    {
      double s = 0;
      for (int k = 0; k < 4; k++) s += a[4 * i + k] * b[4 * k + j];
      c[4 * i + j] = s;
    }
}
static void mat6_vec(const double* a, const double* x, double* y)
{
  for (int i = 0; i < 6; i++)
  {
    double s = 0;
    for (int k = 0; k < 6; k++) s += a[6 * i + k] * x[k];
    y[i] = s;
  }
}
static void cross3(const double* a, const double* b, double* c)
{
  c[0] = a[1] * b[2] - a[2] * b[1];
  c[1] = a[2] * b[0] - a[0] * b[2];
  c[2] = a[0] * b[1] - a[1] * b[0];
}
/* SA.h:69-76 */
static void skew(const double* v, double* m)
{
  m[0] = 0; m[1] = -v[2]; m[2] = v[1];
  m[3] = v[2]; m[4] = 0; m[5] = -v[0];
  m[6] = -v[1]; m[7] = v[0]; m[8] = 0;
}
/* SA.h:88-93  spatialCrossProduct (twist x twist) */
static void spatial_cross(const double* a, const double* b, double* r)
{
  double t1[3], t2[3];
  cross3(a + 3, b + 3, r + 3);
  cross3(a + 3, b, t1);
  cross3(a, b + 3, t2);
  for (int i = 0; i < 3; i++) r[i] = t1[i] + t2[i];
}
/* SA.h:108-113 spatialDualCrossProduct (twist x* wrench) */
static void spatial_dual_cross(const double* a, const double* w, double* r)
{
  double t1[3], t2[3];
  cross3(a + 3, w + 3, t1);
  cross3(a, w, t2);
  for (int i = 0; i < 3; i++) r[3 + i] = t1[i] + t2[i];
  cross3(a + 3, w, r);
}
/* SA.h:129-133 spatialTranslation */
static void spatial_translation(const double* t, const double* d, double* r)
{
  double c[3];
  cross3(t + 3, d, c);
  for (int i = 0; i < 3; i++)
  {
    r[i] = t[i] + c[i];
    r[3 + i] = t[3 + i];
  }
}
/* SA.h:150-154 spatialDualTranslation */
static void spatial_dual_translation(const double* w, const double* d, double* r)
{
  double c[3];
  cross3(w, d, c);
  for (int i = 0; i < 3; i++)
  {
    r[i] = w[i];
    r[3 + i] = w[3 + i] + c[i];
  }
}
/* SA.h:172-175 spatialRotation with R given row-major */
static void spatial_rotation(const double* x, const double* R, double* r)
{
  mat3_vec(R, x, r);
  mat3_vec(R, x + 3, r + 3);
}
static void spatial_rotation_T(const double* x, const double* R, double* r)
{
  mat3T_vec(R, x, r);
  mat3T_vec(R, x + 3, r + 3);
}
/* SA.h:193-197 spatialTranformation (twist form; the reference also applies it to ext. wrenches, PI.h:1255) */
static void spatial_transformation(const double* x, const double* R, const double* p, double* r)
{
  double a[3], b[3], c[3];
  mat3_vec(R, x, a);
  mat3_vec(R, x + 3, b);
  cross3(b, p, c);
  for (int i = 0; i < 3; i++)
  {
    r[i] = a[i] + c[i];
    r[3 + i] = b[i];
  }
}
/* SA.h:232-239 computeSpatialInertiaMatrix */
static void spatial_inertia(const double* inertia, const double* cog, double mass, double* S)
{
  double cs[9], cst[9], cc[9];
  skew(cog, cs);
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) cst[3 * i + j] = cs[3 * j + i];
  mat3_mul(cs, cst, cc);
  memset(S, 0, 36 * sizeof(double));
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++)
    {
      S[6 * i + j] = (i == j) ? mass : 0.0;
      S[6 * i + 3 + j] = mass * cst[3 * i + j];
      S[6 * (3 + i) + j] = mass * cs[3 * i + j];
      S[6 * (3 + i) + 3 + j] = inertia[3 * i + j] + mass * cc[3 * i + j];
    }
}

/* ------------------------------------------------------------------ model build (Joint/Link::fromUrdf) */
const char* oracle_kind(void) { return "port"; }

oracle_chain* oracle_chain_create(const rdb_chain_desc* d)
{
  if (!d || d->n_joints < 0 || d->n_joints > OR_MAXJ) return NULL;
  oracle_chain* c = (oracle_chain*)calloc(1, sizeof(oracle_chain));
  c->nJ = d->n_joints;
  c->nL = d->n_joints + 1;
  c->n_in = d->n_inputs;
  memcpy(c->g, d->gravity, sizeof(c->g));
  for (int j = 0; j < c->nJ; j++)
  {
    const rdb_joint_desc* s = &d->joints[j];
    or_joint* o = &c->joint[j];
    o->type = s->type;
    o->input_index = s->input_index;
    memcpy(o->R_pj, s->rot, sizeof(o->R_pj));
    memset(o->T_pj, 0, sizeof(o->T_pj));
    for (int r = 0; r < 3; r++)
    {
      for (int k = 0; k < 3; k++) o->T_pj[4 * r + k] = s->rot[3 * r + k];
      o->T_pj[4 * r + 3] = s->xyz[r];
    }
    o->T_pj[15] = 1.0;
    double n = sqrt(s->axis[0] * s->axis[0] + s->axis[1] * s->axis[1] + s->axis[2] * s->axis[2]);
    for (int k = 0; k < 3; k++) o->axis_j[k] = (n > 0) ? s->axis[k] / n : s->axis[k]; /* PI.h:58-59 */
    skew(o->axis_j, o->K);
    mat3_mul(o->K, o->K, o->K2);
    mat3_vec(o->R_pj, o->axis_j, o->axis_p);
    memset(o->screw_p, 0, sizeof(o->screw_p));
    if (o->type == RDB_JOINT_REVOLUTE) memcpy(o->screw_p + 3, o->axis_p, 3 * sizeof(double));
    else if (o->type == RDB_JOINT_PRISMATIC) memcpy(o->screw_p, o->axis_p, 3 * sizeof(double));
  }
  for (int l = 0; l < c->nL; l++)
  {
    const rdb_link_desc* s = &d->links[l];
    or_link* o = &c->link[l];
    o->mass = s->mass;
    memcpy(o->cog, s->cog, sizeof(o->cog));
    /* PI.h:295-317: inertia(p) = R_p_cog * inertia_cog * R_p_cog^T */
    double I[9] = {s->inertia[0], s->inertia[1], s->inertia[2], s->inertia[1], s->inertia[3],
                   s->inertia[4], s->inertia[2], s->inertia[4], s->inertia[5]};
    double Rt[9], t[9], Ir[9];
    for (int i = 0; i < 3; i++)
      for (int k = 0; k < 3; k++) Rt[3 * i + k] = s->inertial_rot[3 * k + i];
    mat3_mul(s->inertial_rot, I, t);
    mat3_mul(t, Rt, Ir);
    spatial_inertia(Ir, o->cog, o->mass, o->I_cc);
    /* PI.h:342-396 basis matrices */
    memset(o->E, 0, sizeof(o->E));
    for (int i = 0; i < 3; i++) o->E[0][6 * i + i] = 1.0;
    for (int k = 0; k < 3; k++)
    {
      double e[3] = {0, 0, 0}, sk[9];
      e[k] = 1.0;
      skew(e, sk);
      for (int i = 0; i < 3; i++)
        for (int jj = 0; jj < 3; jj++)
        {
          o->E[1 + k][6 * i + 3 + jj] = sk[3 * jj + i];
          o->E[1 + k][6 * (3 + i) + jj] = sk[3 * i + jj];
        }
    }
    o->E[4][6 * 3 + 3] = 1;
    o->E[5][6 * 3 + 4] = 1; o->E[5][6 * 4 + 3] = 1;
    o->E[6][6 * 3 + 5] = 1; o->E[6][6 * 5 + 3] = 1;
    o->E[7][6 * 4 + 4] = 1;
    o->E[8][6 * 4 + 5] = 1; o->E[8][6 * 5 + 4] = 1;
    o->E[9][6 * 5 + 5] = 1;
    /* PI.h:399-417 */
    o->nominal[0] = o->mass;
    for (int k = 0; k < 3; k++) o->nominal[1 + k] = o->cog[k] * o->mass;
    o->nominal[4] = o->I_cc[6 * 3 + 3];
    o->nominal[5] = o->I_cc[6 * 3 + 4];
    o->nominal[6] = o->I_cc[6 * 3 + 5];
    o->nominal[7] = o->I_cc[6 * 4 + 4];
    o->nominal[8] = o->I_cc[6 * 4 + 5];
    o->nominal[9] = o->I_cc[6 * 5 + 5];
  }
  return c;
}
void oracle_chain_destroy(oracle_chain* c) { free(c); }
int oracle_chain_joints_number(const oracle_chain* c) { return c->nJ; }
int oracle_chain_inputs_number(const oracle_chain* c) { return c->n_in; }

/* PI.h:1382-1391 */
void oracle_nominal_parameters(const oracle_chain* c, double* out)
{
  for (int nl = c->nL - 1; nl > 0; nl--) memcpy(out + 10 * (nl - 1), c->link[nl].nominal, 10 * sizeof(double));
}

/* ------------------------------------------------------------------ per-sample state */
typedef struct
{
  double sq[OR_MAXJ], sdq[OR_MAXJ], sddq[OR_MAXJ], sdddq[OR_MAXJ]; /* sorted = S*x, PI.h:865,984,1086,1188 */
  double T[OR_MAXL][16];  /* m_T_bl, 4x4 row-major */
  double R[OR_MAXL][9];   /* T.linear() */
  double p[OR_MAXL][3];   /* T.translation() */
  double s[OR_MAXL][6];   /* m_screws_of_c_in_b */
  double v[OR_MAXL][6], a[OR_MAXL][6];
} or_state;

static void scatter(const oracle_chain* c, const double* x, double* sorted)
{
  for (int j = 0; j < c->nJ; j++)
  {
    int ii = c->joint[j].input_index;
    sorted[j] = (x && ii >= 0) ? x[ii] : 0.0;
  }
}

/* Joint::computedTpc, PI.h:38-47 (4x4 row-major) */
static void joint_T_pc(const or_joint* jn, double q, double* T)
{
  memcpy(T, jn->T_pj, 16 * sizeof(double));
  if (jn->type == RDB_JOINT_REVOLUTE)
  {
    double sq = sin(q), cq = cos(q), Rjc[9], Rpc[9];
    for (int i = 0; i < 9; i++) Rjc[i] = ((i % 4 == 0) ? 1.0 : 0.0) + sq * jn->K[i] + (1 - cq) * jn->K2[i];
    mat3_mul(jn->R_pj, Rjc, Rpc);
    for (int r = 0; r < 3; r++)
      for (int k = 0; k < 3; k++) T[4 * r + k] = Rpc[3 * r + k];
  }
  else if (jn->type == RDB_JOINT_PRISMATIC)
  {
    for (int r = 0; r < 3; r++) T[4 * r + 3] = jn->T_pj[4 * r + 3] + jn->axis_p[r] * q;
  }
}

/* computeFrames PI.h:863-872 + computeScrews PI.h:874-882 */
static void frames_and_screws(const oracle_chain* c, or_state* st)
{
  memset(st->T[0], 0, sizeof(st->T[0]));
  st->T[0][0] = st->T[0][5] = st->T[0][10] = st->T[0][15] = 1.0;
  for (int nl = 1; nl < c->nL; nl++)
  {
    double Tpc[16];
    joint_T_pc(&c->joint[nl - 1], st->sq[nl - 1], Tpc);
    mat4_mul(st->T[nl - 1], Tpc, st->T[nl]);
  }
  for (int nl = 0; nl < c->nL; nl++)
    for (int r = 0; r < 3; r++)
    {
      for (int k = 0; k < 3; k++) st->R[nl][3 * r + k] = st->T[nl][4 * r + k];
      st->p[nl][r] = st->T[nl][4 * r + 3];
    }
  memset(st->s[0], 0, sizeof(st->s[0]));
  for (int nl = 1; nl < c->nL; nl++) spatial_rotation(c->joint[nl - 1].screw_p, st->R[nl - 1], st->s[nl]);
}

static void dp(const or_state* st, int a, int b, double* d) /* p[a] - p[b] */
{
  for (int i = 0; i < 3; i++) d[i] = st->p[a][i] - st->p[b][i];
}

/* getTwist PI.h:1004-1009 */
static void twists(const oracle_chain* c, or_state* st)
{
  memset(st->v[0], 0, sizeof(st->v[0]));
  for (int nl = 1; nl < c->nL; nl++)
  {
    double d[3], t[6];
    dp(st, nl, nl - 1, d);
    spatial_translation(st->v[nl - 1], d, t);
    for (int i = 0; i < 6; i++) st->v[nl][i] = t[i] + st->s[nl][i] * st->sdq[nl - 1];
  }
}
/* getDTwist direct path PI.h:1113-1118 */
static void dtwists(const oracle_chain* c, or_state* st)
{
  memset(st->a[0], 0, sizeof(st->a[0]));
  for (int nl = 1; nl < c->nL; nl++)
  {
    double d[3], t[6], x[6];
    dp(st, nl, nl - 1, d);
    spatial_translation(st->a[nl - 1], d, t);
    spatial_cross(st->v[nl], st->s[nl], x);
    for (int i = 0; i < 6; i++) st->a[nl][i] = t[i] + x[i] * st->sdq[nl - 1] + st->s[nl][i] * st->sddq[nl - 1];
  }
}

/* ------------------------------------------------------------------ public per-sample evaluation */
typedef struct oracle_out
{
  double* T_links;        /* [nL][12] 3x4 row-major */
  double* jacobian;       /* [n_in][6] column-major 6 x n_in: jac[col*6+row] */
  double* twist;          /* [nL][6] */
  double* dtwist;         /* [nL][6] */
  double* dtwist_lin;     /* [nL][6] */
  double* dtwist_nonlin;  /* [nL][6] */
  double* ddtwist;        /* [nL][6] */
  double* ddtwist_lin;    /* [nL][6] */
  double* ddtwist_nonlin; /* [nL][6] */
  double* wrench;         /* [nL][6] getWrench */
  const double* ext;      /* [nL][6] ext_wrenches_in_link_frame (PI.h:1225), NULL = zero */
  double* jac_link;       /* [n_in][6] getJacobianLink of link `link` (PI.h:951-979) */
  int link;
  double* torque;         /* [n_in] */
  double* regressor;      /* [10 nJ][n_in] column-major n_in x 10nJ: phi[col*n_in+row] */
  double* inertia;        /* [n_in][n_in] column-major */
} oracle_out;

void oracle_eval(const oracle_chain* c, const double* q, const double* dq, const double* ddq, const double* dddq,
                 const oracle_out* o)
{
  or_state st;
  const int nJ = c->nJ, nL = c->nL, n_in = c->n_in;
  scatter(c, q, st.sq);
  scatter(c, dq, st.sdq);
  scatter(c, ddq, st.sddq);
  scatter(c, dddq, st.sdddq);
  frames_and_screws(c, &st);

  if (o->T_links)
    for (int nl = 0; nl < nL; nl++)
      for (int r = 0; r < 3; r++)
        for (int k = 0; k < 4; k++) o->T_links[12 * nl + 4 * r + k] = st.T[nl][4 * r + k];

  /* getJacobian PI.h:939-945 */
  if (o->jacobian)
  {
    memset(o->jacobian, 0, sizeof(double) * 6 * n_in);
    for (int nj = 0; nj < nJ; nj++)
    {
      int idx = c->joint[nj].input_index;
      if (idx < 0 || c->joint[nj].type == RDB_JOINT_FIXED) continue;
      double d[3];
      dp(&st, nL - 1, nj + 1, d);
      spatial_translation(st.s[nj + 1], d, o->jacobian + 6 * idx);
    }
  }

  /* getJacobianLink PI.h:966-976: joints = active joints between base and link; the loop indexes m_active_joints BY THE
   * COUNTER (PI.h:970), i.e. the first joints.size() entries of the input-ordered active list */
  if (o->jac_link)
  {
    int k_before = 0;
    memset(o->jac_link, 0, sizeof(double) * 6 * n_in);
    for (int nj = 0; nj < o->link && nj < nJ; nj++)
      if (c->joint[nj].input_index >= 0) k_before++;
    for (int nj = 0; nj < nJ; nj++)
    {
      int idx = c->joint[nj].input_index;
      if (idx < 0 || idx >= k_before || c->joint[nj].type == RDB_JOINT_FIXED) continue;
      double d[3];
      dp(&st, o->link, nj + 1, d);
      spatial_translation(st.s[nj + 1], d, o->jac_link + 6 * idx);
    }
  }

  twists(c, &st);
  dtwists(c, &st);
  if (o->twist) memcpy(o->twist, st.v, sizeof(double) * 6 * nL);
  if (o->dtwist) memcpy(o->dtwist, st.a, sizeof(double) * 6 * nL);

  /* getDTwistLinearPart PI.h:1052-1057 / NonLinearPart PI.h:1071-1076 */
  if (o->dtwist_lin || o->dtwist_nonlin)
  {
    double al[OR_MAXL][6], an[OR_MAXL][6];
    memset(al[0], 0, sizeof(al[0]));
    memset(an[0], 0, sizeof(an[0]));
    for (int nl = 1; nl < nL; nl++)
    {
      double d[3], t[6], x[6];
      dp(&st, nl, nl - 1, d);
      spatial_translation(al[nl - 1], d, t);
      for (int i = 0; i < 6; i++) al[nl][i] = t[i] + st.s[nl][i] * st.sddq[nl - 1];
      spatial_translation(an[nl - 1], d, t);
      spatial_cross(st.v[nl], st.s[nl], x);
      for (int i = 0; i < 6; i++) an[nl][i] = t[i] + x[i] * st.sdq[nl - 1];
    }
    if (o->dtwist_lin) memcpy(o->dtwist_lin, al, sizeof(double) * 6 * nL);
    if (o->dtwist_nonlin) memcpy(o->dtwist_nonlin, an, sizeof(double) * 6 * nL);
  }

  /* getDDTwist direct PI.h:1210-1219, LinearPart PI.h:1145-1150, NonLinearPart PI.h:1171-1179.
   * The 1x coefficient on (v x s) DDq is the reference's and is mirrored, not fixed. */
  if (o->ddtwist || o->ddtwist_lin || o->ddtwist_nonlin)
  {
    double jf[OR_MAXL][6], jl[OR_MAXL][6], jn[OR_MAXL][6];
    memset(jf[0], 0, sizeof(jf[0]));
    memset(jl[0], 0, sizeof(jl[0]));
    memset(jn[0], 0, sizeof(jn[0]));
    for (int nl = 1; nl < nL; nl++)
    {
      int nj = nl - 1;
      double d[3], t[6], vxs[6], axs[6], vvxs[6];
      dp(&st, nl, nl - 1, d);
      spatial_cross(st.v[nl], st.s[nl], vxs);
      spatial_cross(st.a[nl], st.s[nl], axs);
      spatial_cross(st.v[nl], vxs, vvxs);
      spatial_translation(jf[nl - 1], d, t);
      for (int i = 0; i < 6; i++)
        jf[nl][i] = t[i] + st.s[nl][i] * st.sdddq[nj] + vxs[i] * st.sddq[nj] + (axs[i] + vvxs[i]) * st.sdq[nj];
      spatial_translation(jl[nl - 1], d, t);
      for (int i = 0; i < 6; i++) jl[nl][i] = t[i] + st.s[nl][i] * st.sdddq[nj];
      spatial_translation(jn[nl - 1], d, t);
      for (int i = 0; i < 6; i++) jn[nl][i] = t[i] + vxs[i] * st.sddq[nj] + (axs[i] + vvxs[i]) * st.sdq[nj];
    }
    if (o->ddtwist) memcpy(o->ddtwist, jf, sizeof(double) * 6 * nL);
    if (o->ddtwist_lin) memcpy(o->ddtwist_lin, jl, sizeof(double) * 6 * nL);
    if (o->ddtwist_nonlin) memcpy(o->ddtwist_nonlin, jn, sizeof(double) * 6 * nL);
  }

  /* getWrench PI.h:1231-1258 with zero external wrenches, getJointTorque PI.h:1267-1273 */
  if (o->wrench || o->torque)
  {
    double w[OR_MAXL][6];
    for (int nl = nL - 1; nl >= 0; nl--)
    {
      double inertial[6] = {0, 0, 0, 0, 0, 0}, grav[6] = {0, 0, 0, 0, 0, 0}, ext[6];
      if (nl > 0)
      {
        const or_link* L = &c->link[nl];
        double al[6], vl[6], Ia[6], Iv[6], vl2[6], x[6], sum[6];
        spatial_rotation_T(st.a[nl], st.R[nl], al);
        mat6_vec(L->I_cc, al, Ia);
        spatial_rotation_T(st.v[nl], st.R[nl], vl);
        spatial_rotation_T(st.v[nl], st.R[nl], vl2); /* the reference evaluates R^T v twice, PI.h:1245,1247 */
        mat6_vec(L->I_cc, vl2, Iv);
        spatial_dual_cross(vl, Iv, x);
        for (int i = 0; i < 6; i++) sum[i] = Ia[i] + x[i];
        spatial_rotation(sum, st.R[nl], inertial);
        double Rc[3], mg[3], cr[3];
        for (int i = 0; i < 3; i++) grav[i] = -L->mass * c->g[i];
        mat3_vec(st.R[nl], L->cog, Rc);
        for (int i = 0; i < 3; i++) mg[i] = L->mass * c->g[i];
        cross3(Rc, mg, cr);
        for (int i = 0; i < 3; i++) grav[3 + i] = -cr[i];
      }
      double zero6[6] = {-0.0, -0.0, -0.0, -0.0, -0.0, -0.0}; /* -ext with ext == 0 */
      if (o->ext)
        for (int i = 0; i < 6; i++) zero6[i] = -o->ext[6 * nl + i];
      spatial_transformation(zero6, st.R[nl], st.p[nl], ext); /* PI.h:1255: the twist transform, applied to a wrench */
      if (nl < nL - 1)
      {
        double d[3], tr[6];
        dp(&st, nl, nl + 1, d);
        spatial_dual_translation(w[nl + 1], d, tr);
        for (int i = 0; i < 6; i++) w[nl][i] = ext[i] + inertial[i] + grav[i] + tr[i];
      }
      else
        for (int i = 0; i < 6; i++) w[nl][i] = ext[i] + inertial[i] + grav[i];
    }
    if (o->wrench) memcpy(o->wrench, w, sizeof(double) * 6 * nL);
    if (o->torque)
    {
      for (int i = 0; i < n_in; i++) o->torque[i] = 0.0;
      for (int nj = 0; nj < nJ; nj++)
      {
        double t = 0;
        for (int i = 0; i < 6; i++) t += w[nj + 1][i] * st.s[nj + 1][i];
        if (c->joint[nj].input_index >= 0) o->torque[c->joint[nj].input_index] = t;
      }
    }
  }

  /* getRegressor PI.h:1321-1352 */
  if (o->regressor)
  {
    static const int unit[3] = {0, 1, 2};
    double W[OR_MAXL][60]; /* 6 x 10 row-major per link */
    const int P = 10 * nJ;
    double* Rext = (double*)calloc((size_t)nJ * P, sizeof(double)); /* m_regressor_extended nJ x 10nJ row-major */
    for (int nl = nL - 1; nl > 0; nl--)
    {
      const or_link* L = &c->link[nl];
      for (int ip = 0; ip < 10; ip++)
      {
        double al[6], vl[6], vl2[6], Ea[6], Ev[6], x[6], sum[6], col[6];
        spatial_rotation_T(st.a[nl], st.R[nl], al);
        mat6_vec(L->E[ip], al, Ea);
        spatial_rotation_T(st.v[nl], st.R[nl], vl);
        spatial_rotation_T(st.v[nl], st.R[nl], vl2);
        mat6_vec(L->E[ip], vl2, Ev);
        spatial_dual_cross(vl, Ev, x);
        for (int i = 0; i < 6; i++) sum[i] = Ea[i] + x[i];
        spatial_rotation(sum, st.R[nl], col);
        for (int i = 0; i < 6; i++) W[nl][10 * i + ip] = col[i];
      }
      for (int i = 0; i < 3; i++) W[nl][10 * i + 0] -= c->g[i]; /* PI.h:1336 */
      for (int k = 0; k < 3; k++)                                 /* PI.h:1337-1339 */
      {
        double e[3] = {0, 0, 0}, Re[3], cr[3];
        e[unit[k]] = 1.0;
        mat3_vec(st.R[nl], e, Re);
        cross3(Re, c->g, cr);
        for (int i = 0; i < 3; i++) W[nl][10 * (3 + i) + 1 + k] -= cr[i];
      }
      for (int ip = 0; ip < 10; ip++) /* PI.h:1341 */
      {
        double t = 0;
        for (int i = 0; i < 6; i++) t += st.s[nl][i] * W[nl][10 * i + ip];
        Rext[(nl - 1) * P + (nl - 1) * 10 + ip] = t;
      }
      for (int nlf = nl + 1; nlf < nL; nlf++) /* PI.h:1343-1347 */
      {
        double d[3];
        dp(&st, nl, nlf, d);
        for (int ip = 0; ip < 10; ip++)
        {
          double col[6], tr[6], t = 0;
          for (int i = 0; i < 6; i++) col[i] = W[nlf][10 * i + ip];
          spatial_dual_translation(col, d, tr);
          for (int i = 0; i < 6; i++) t += st.s[nl][i] * tr[i];
          Rext[(nl - 1) * P + (nlf - 1) * 10 + ip] = t;
        }
      }
    }
    /* result = S^T * R_ext, PI.h:1352 */
    memset(o->regressor, 0, sizeof(double) * (size_t)n_in * P);
    for (int nj = 0; nj < nJ; nj++)
    {
      int r = c->joint[nj].input_index;
      if (r < 0) continue;
      for (int col = 0; col < P; col++) o->regressor[(size_t)col * n_in + r] = Rext[nj * P + col];
    }
    free(Rext);
  }

  /* getJointInertia PI.h:1361-1377 */
  if (o->inertia)
  {
    double* Mext = (double*)calloc((size_t)nJ * nJ, sizeof(double));
    double* Jn = (double*)malloc(sizeof(double) * 6 * nJ);  /* 6 x nJ row-major */
    double* IJ = (double*)malloc(sizeof(double) * 6 * nJ);
    for (int nj = 0; nj < nJ; nj++)
    {
      memset(Jn, 0, sizeof(double) * 6 * nJ);
      for (int ij = 0; ij <= nj; ij++)
      {
        int il = ij + 1;
        if (c->joint[ij].type == RDB_JOINT_FIXED) continue;
        double d[3], t[6], r[6];
        dp(&st, nj + 1, il, d);
        spatial_translation(st.s[il], d, t);
        spatial_rotation_T(t, st.R[nj + 1], r);
        for (int i = 0; i < 6; i++) Jn[i * nJ + ij] = r[i];
      }
      const double* I = c->link[nj + 1].I_cc;
      for (int i = 0; i < 6; i++)
        for (int k = 0; k < nJ; k++)
        {
          double s = 0;
          for (int m = 0; m < 6; m++) s += I[6 * i + m] * Jn[m * nJ + k];
          IJ[i * nJ + k] = s;
        }
      for (int a = 0; a < nJ; a++)
        for (int b = 0; b < nJ; b++)
        {
          double s = 0;
          for (int m = 0; m < 6; m++) s += Jn[m * nJ + a] * IJ[m * nJ + b];
          Mext[a * nJ + b] += s;
        }
    }
    memset(o->inertia, 0, sizeof(double) * (size_t)n_in * n_in);
    for (int a = 0; a < nJ; a++)
      for (int b = 0; b < nJ; b++)
      {
        int ia = c->joint[a].input_index, ib = c->joint[b].input_index;
        if (ia >= 0 && ib >= 0) o->inertia[(size_t)ib * n_in + ia] = Mext[a * nJ + b];
      }
    free(Mext);
    free(Jn);
    free(IJ);
  }
}

/* ------------------------------------------------------------------ batched drivers (SoA planes, like the ABI) */
/* ------------------------------------------------------------------ local IK (PI.h:1398-1468, frame_distance.h:44-49) */
/* Eigen::AngleAxisd(Matrix3d) = AngleAxis(Quaternion(m)); returns angle * axis (m row-major).
 * Quaternion from matrix: Eigen/src/Geometry/Quaternion.h (quaternionbase_assign_impl<Other,3,3>); angle-axis from quaternion:
 * Eigen/src/Geometry/AngleAxis.h (operator=(QuaternionBase)).  Eigen is un-vendored by the reference; published algorithm restated. */
static void angle_axis_of_matrix(const double* m, double* out)
{
  double q[4]; /* x y z w */
  double t = m[0] + m[4] + m[8];
  if (t > 0.0)
  {
    t = sqrt(t + 1.0);
    q[3] = 0.5 * t;
    t = 0.5 / t;
    q[0] = (m[7] - m[5]) * t;
    q[1] = (m[2] - m[6]) * t;
    q[2] = (m[3] - m[1]) * t;
  }
  else
  {
    int i = 0;
    if (m[4] > m[0]) i = 1;
    if (m[8] > m[4 * i]) i = 2;
    int j = (i + 1) % 3, k = (j + 1) % 3;
    t = sqrt(m[4 * i] - m[4 * j] - m[4 * k] + 1.0);
    q[i] = 0.5 * t;
    t = 0.5 / t;
    q[3] = (m[3 * k + j] - m[3 * j + k]) * t;
    q[j] = (m[3 * j + i] + m[3 * i + j]) * t;
    q[k] = (m[3 * k + i] + m[3 * i + k]) * t;
  }
  double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2]);
  if (n != 0.0)
  {
    double angle = 2.0 * atan2(n, fabs(q[3]));
    if (q[3] < 0.0) n = -n;
    for (int c = 0; c < 3; c++) out[c] = angle * (q[c] / n);
  }
  else
    out[0] = out[1] = out[2] = 0.0;
}

/* getFrameDistance(T_wa, T_wb) with T = [R | p] 3x4 row-major (frame_distance.h:44-49) */
void oracle_frame_distance(const double* Ta, const double* Tb, double* d)
{
  double Ra[9], Rb[9], Rab[9], aa[3];
  for (int r = 0; r < 3; r++)
    for (int k = 0; k < 3; k++)
    {
      Ra[3 * r + k] = Ta[4 * r + k];
      Rb[3 * r + k] = Tb[4 * r + k];
    }
  for (int r = 0; r < 3; r++) d[r] = Ta[4 * r + 3] - Tb[4 * r + 3];
  for (int a = 0; a < 3; a++)
    for (int b = 0; b < 3; b++) Rab[3 * a + b] = Ra[a] * Rb[b] + Ra[3 + a] * Rb[3 + b] + Ra[6 + a] * Rb[6 + b]; /* R_a^-1 R_b */
  angle_axis_of_matrix(Rab, aa);
  for (int a = 0; a < 3; a++) d[3 + a] = -(Ra[3 * a] * aa[0] + Ra[3 * a + 1] * aa[1] + Ra[3 * a + 2] * aa[2]);
}

#include "box_qp.h"
void oracle_box_qp(int n, const double* H, const double* f, const double* lo, const double* hi, double* x) { oracle_box_qp_impl(n, H, f, lo, hi, x); }

/* computeLocalIk (weight == NULL, PI.h:1398-1432) / computeWeigthedLocalIk (PI.h:1435-1468) for n targets.  The reference's wall-clock
 * budget is an iteration budget here: at most max_iter steps, max_iter + 1 convergence checks.  target[12][ld] (3x4 row-major planes),
 * seed / sol [n_in][ld]; status 1 = converged (the reference's return value). */
void oracle_local_ik_batch(const oracle_chain* c, int64_t n, int64_t ld, const double* target, const double* seed, const double* q_min,
                           const double* q_max, const double* weight, double toll, int max_iter, double* sol_out, int32_t* status,
                           int32_t* iters, double* err_norm)
{
  const int n_in = c->n_in, nL = c->nL;
  if (n_in > OR_IK_MAXN) return;
  for (int64_t i = 0; i < n; i++)
  {
    double Ta[12], sol[OR_IK_MAXN], bT[OR_MAXL * 12], J[OR_IK_MAXN * 6], e[6], en = 0.0;
    for (int k = 0; k < 12; k++) Ta[k] = target[(int64_t)k * ld + i];
    for (int r = 0; r < n_in; r++) sol[r] = seed[(int64_t)r * ld + i];
    int done = 0, it = 0;
    for (;; it++)
    {
      oracle_out o;
      memset(&o, 0, sizeof(o));
      o.T_links = bT;
      o.jacobian = J;
      oracle_eval(c, sol, NULL, NULL, NULL, &o);
      oracle_frame_distance(Ta, bT + 12 * (nL - 1), e);
      en = 0.0;
      for (int k = 0; k < 6; k++)
      {
        double we = weight ? weight[k] * e[k] : e[k];
        en += we * we;
      }
      en = sqrt(en);
      if (en < toll) { done = 1; break; }
      if (it >= max_iter) break;
      double H[OR_IK_MAXN * OR_IK_MAXN], f[OR_IK_MAXN], lo[OR_IK_MAXN], hi[OR_IK_MAXN], dq[OR_IK_MAXN];
      for (int a = 0; a < n_in; a++)
      {
        for (int b = 0; b < n_in; b++)
        {
          double s = 0.0;
          for (int k = 0; k < 6; k++) s += J[6 * a + k] * (weight ? weight[k] : 1.0) * J[6 * b + k];
          H[a * n_in + b] = s;
        }
        double s = 0.0;
        for (int k = 0; k < 6; k++) s += J[6 * a + k] * (weight ? weight[k] : 1.0) * e[k];
        f[a] = -s;
        lo[a] = (q_min ? q_min[a] : -1e10) - sol[a];
        hi[a] = (q_max ? q_max[a] : 1e10) - sol[a];
      }
      oracle_box_qp(n_in, H, f, lo, hi, dq);
      for (int a = 0; a < n_in; a++) sol[a] += dq[a];
    }
    for (int r = 0; r < n_in; r++) sol_out[(int64_t)r * ld + i] = sol[r];
    if (status) status[i] = done;
    if (iters) iters[i] = it;
    if (err_norm) err_norm[i] = en;
  }
}

static int or_threads(int nthreads)
{
#ifdef _OPENMP
  return nthreads > 0 ? nthreads : omp_get_max_threads();
#else
  (void)nthreads;
  return 1;
#endif
}
int oracle_max_threads(void) { return or_threads(0); }

static void gather_in(const double* x, int n_in, int64_t ld, int64_t i, double* v)
{
  for (int j = 0; j < n_in; j++) v[j] = x ? x[(int64_t)j * ld + i] : 0.0;
}

/* Everything the ABI's rdb_kinematics_out holds; pointers optional; planes [..][ld_out]. */
void oracle_kinematics_batch(const oracle_chain* c, int64_t n, int64_t ld, const double* q, const double* dq, const double* ddq,
                             const double* dddq, int64_t ld_out, double* T_tool, double* T_links, double* jacobian, double* twist,
                             double* dtwist, double* dtwist_lin, double* dtwist_nonlin, double* ddtwist, double* ddtwist_lin,
                             double* ddtwist_nonlin, double* torque, int nthreads)
{
  const int nL = c->nL, n_in = c->n_in;
  (void)nthreads;
#pragma omp parallel for num_threads(or_threads(nthreads)) schedule(static)
  for (int64_t i = 0; i < n; i++)
  {
    double vq[OR_MAXJ], vdq[OR_MAXJ], vddq[OR_MAXJ], vdddq[OR_MAXJ];
    double bT[OR_MAXL * 12], bJ[OR_MAXJ * 6], bv[OR_MAXL * 6], ba[OR_MAXL * 6], bal[OR_MAXL * 6], ban[OR_MAXL * 6];
    double bj[OR_MAXL * 6], bjl[OR_MAXL * 6], bjn[OR_MAXL * 6], bt[OR_MAXJ];
    gather_in(q, n_in, ld, i, vq);
    gather_in(dq, n_in, ld, i, vdq);
    gather_in(ddq, n_in, ld, i, vddq);
    gather_in(dddq, n_in, ld, i, vdddq);
    oracle_out o;
    memset(&o, 0, sizeof(o));
    o.T_links = (T_tool || T_links) ? bT : NULL;
    o.jacobian = jacobian ? bJ : NULL;
    o.twist = twist ? bv : NULL;
    o.dtwist = dtwist ? ba : NULL;
    o.dtwist_lin = dtwist_lin ? bal : NULL;
    o.dtwist_nonlin = dtwist_nonlin ? ban : NULL;
    o.ddtwist = ddtwist ? bj : NULL;
    o.ddtwist_lin = ddtwist_lin ? bjl : NULL;
    o.ddtwist_nonlin = ddtwist_nonlin ? bjn : NULL;
    o.torque = torque ? bt : NULL;
    oracle_eval(c, vq, vdq, vddq, vdddq, &o);
    if (T_tool)
      for (int k = 0; k < 12; k++) T_tool[(int64_t)k * ld_out + i] = bT[12 * (nL - 1) + k];
    if (T_links)
      for (int k = 0; k < 12 * nL; k++) T_links[(int64_t)k * ld_out + i] = bT[k];
    if (jacobian)
      for (int k = 0; k < 6 * n_in; k++) jacobian[(int64_t)k * ld_out + i] = bJ[k];
#define OR_PUT6(dst, src)                                                  \
  if (dst)                                                                 \
    for (int k = 0; k < 6 * nL; k++) dst[(int64_t)k * ld_out + i] = src[k];
    OR_PUT6(twist, bv)
    OR_PUT6(dtwist, ba)
    OR_PUT6(dtwist_lin, bal)
    OR_PUT6(dtwist_nonlin, ban)
    OR_PUT6(ddtwist, bj)
    OR_PUT6(ddtwist_lin, bjl)
    OR_PUT6(ddtwist_nonlin, bjn)
#undef OR_PUT6
    if (torque)
      for (int k = 0; k < n_in; k++) torque[(int64_t)k * ld_out + i] = bt[k];
  }
}

/* getJointTorque + getRegressor per sample (config 1 of BASELINE.json); phi planes col*n_in+row. */
void oracle_regressor_torque_batch(const oracle_chain* c, int64_t n, int64_t ld, const double* q, const double* dq, const double* ddq,
                                   int64_t ld_out, double* phi, double* torque, int nthreads)
{
  const int n_in = c->n_in, P = 10 * c->nJ;
  (void)nthreads;
#pragma omp parallel for num_threads(or_threads(nthreads)) schedule(static)
  for (int64_t i = 0; i < n; i++)
  {
    double vq[OR_MAXJ], vdq[OR_MAXJ], vddq[OR_MAXJ], bt[OR_MAXJ];
    double* bphi = (double*)malloc(sizeof(double) * (size_t)n_in * P); /* the reference returns MatrixXd by value */
    gather_in(q, n_in, ld, i, vq);
    gather_in(dq, n_in, ld, i, vdq);
    gather_in(ddq, n_in, ld, i, vddq);
    oracle_out o;
    memset(&o, 0, sizeof(o));
    o.torque = bt;
    oracle_eval(c, vq, vdq, vddq, NULL, &o); /* getJointTorque */
    memset(&o, 0, sizeof(o));
    o.regressor = bphi;
    oracle_eval(c, vq, vdq, vddq, NULL, &o); /* getRegressor */
    if (phi)
      for (int k = 0; k < n_in * P; k++) phi[(int64_t)k * ld_out + i] = bphi[k];
    if (torque)
      for (int k = 0; k < n_in; k++) torque[(int64_t)k * ld_out + i] = bt[k];
    free(bphi);
  }
}

void oracle_inertia_batch(const oracle_chain* c, int64_t n, int64_t ld, const double* q, int64_t ld_out, double* inertia, int nthreads)
{
  const int n_in = c->n_in;
  (void)nthreads;
#pragma omp parallel for num_threads(or_threads(nthreads)) schedule(static)
  for (int64_t i = 0; i < n; i++)
  {
    double vq[OR_MAXJ], bM[OR_MAXJ * OR_MAXJ];
    gather_in(q, n_in, ld, i, vq);
    oracle_out o;
    memset(&o, 0, sizeof(o));
    o.inertia = bM;
    oracle_eval(c, vq, NULL, NULL, NULL, &o);
    for (int k = 0; k < n_in * n_in; k++) inertia[(int64_t)k * ld_out + i] = bM[k];
  }
}

/* getWrench / getJointTorque with external wrenches (PI.h:1225-1274) and getJacobianLink (PI.h:951-979), batched */
void oracle_wrench_batch(const oracle_chain* c, int64_t n, int64_t ld, const double* q, const double* dq, const double* ddq,
                         const double* ext, int64_t ld_ext, int64_t ld_out, double* torque, double* wrenches)
{
  const int nL = c->nL, n_in = c->n_in;
  for (int64_t i = 0; i < n; i++)
  {
    double vq[OR_MAXJ], vdq[OR_MAXJ], vddq[OR_MAXJ], bt[OR_MAXJ], bw[OR_MAXL * 6], be[OR_MAXL * 6];
    gather_in(q, n_in, ld, i, vq);
    gather_in(dq, n_in, ld, i, vdq);
    gather_in(ddq, n_in, ld, i, vddq);
    if (ext)
      for (int k = 0; k < 6 * nL; k++) be[k] = ext[(int64_t)k * ld_ext + i];
    oracle_out o;
    memset(&o, 0, sizeof(o));
    o.torque = bt;
    o.wrench = bw;
    o.ext = ext ? be : NULL;
    oracle_eval(c, vq, vdq, vddq, NULL, &o);
    if (torque)
      for (int k = 0; k < n_in; k++) torque[(int64_t)k * ld_out + i] = bt[k];
    if (wrenches)
      for (int k = 0; k < 6 * nL; k++) wrenches[(int64_t)k * ld_out + i] = bw[k];
  }
}
void oracle_jacobian_link_batch(const oracle_chain* c, int64_t n, int64_t ld, const double* q, int link, int64_t ld_out, double* jac)
{
  const int n_in = c->n_in;
  for (int64_t i = 0; i < n; i++)
  {
    double vq[OR_MAXJ], bJ[OR_MAXJ * 6];
    gather_in(q, n_in, ld, i, vq);
    oracle_out o;
    memset(&o, 0, sizeof(o));
    o.jac_link = bJ;
    o.link = link;
    oracle_eval(c, vq, NULL, NULL, NULL, &o);
    for (int k = 0; k < 6 * n_in; k++) jac[(int64_t)k * ld_out + i] = bJ[k];
  }
}

/* Normal equations with long-double accumulation (reference sums for the fused Gram kernel):
 * gram[P*P] column-major, rhs[P], tau_sq[1].  tau = tau_meas if given else getJointTorque. */
void oracle_regressor_gram(const oracle_chain* c, int64_t n, int64_t ld, const double* q, const double* dq, const double* ddq,
                           const double* tau_meas, double* gram, double* rhs, double* tau_sq)
{
  const int n_in = c->n_in, P = 10 * c->nJ;
  long double* G = (long double*)calloc((size_t)P * P, sizeof(long double));
  long double* b = (long double*)calloc((size_t)P, sizeof(long double));
  long double tt = 0;
  double* bphi = (double*)malloc(sizeof(double) * (size_t)n_in * P);
  for (int64_t i = 0; i < n; i++)
  {
    double vq[OR_MAXJ], vdq[OR_MAXJ], vddq[OR_MAXJ], bt[OR_MAXJ];
    gather_in(q, n_in, ld, i, vq);
    gather_in(dq, n_in, ld, i, vdq);
    gather_in(ddq, n_in, ld, i, vddq);
    oracle_out o;
    memset(&o, 0, sizeof(o));
    o.torque = bt;
    o.regressor = bphi;
    oracle_eval(c, vq, vdq, vddq, NULL, &o);
    if (tau_meas) gather_in(tau_meas, n_in, ld, i, bt);
    for (int r = 0; r < n_in; r++)
    {
      tt += (long double)bt[r] * bt[r];
      for (int a = 0; a < P; a++)
      {
        double pa = bphi[(size_t)a * n_in + r];
        if (pa == 0.0) continue;
        b[a] += (long double)pa * bt[r];
        for (int bb = 0; bb < P; bb++) G[(size_t)bb * P + a] += (long double)pa * bphi[(size_t)bb * n_in + r];
      }
    }
  }
  for (int k = 0; k < P * P; k++) gram[k] = (double)G[k];
  for (int k = 0; k < P; k++) rhs[k] = (double)b[k];
  if (tau_sq) *tau_sq = (double)tt;
  free(G);
  free(b);
  free(bphi);
}

/* same generator as rdb_fill_uniform_host (include/rosdyn_b200.h) */
static uint64_t splitmix64(uint64_t x)
{
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
/* ------------------------------------------------------------------ additive joint components
 * FirstOrderPolynomialFriction::computeRegressor   friction_polynomial1.h:45-52 (+ ctor rules :74-87)
 * SecondOrderPolynomialFriction::computeRegressor  friction_polynomial2.h:42-58 (+ ctor rules :84-96, threshold quirk included)
 * IdealSpring::getRegressor                        ideal_spring.h:64-70
 * phi_c planes col*n_in+row; a component's columns are zero except in the row of its joint (dense m_regressor). */
static int or_comp_cols(int type) { return type == RDB_COMPONENT_FRICTION_POLY2 ? 3 : 2; }
void oracle_components_regressor_batch(int n_comp, const rdb_component_desc* comp, int n_in, int64_t n, int64_t ld, const double* q,
                                       const double* dq, int64_t ld_out, double* phi_c)
{
  int col = 0;
  for (int k = 0; k < n_comp; k++)
  {
    const rdb_component_desc* c = &comp[k];
    double thr = c->min_velocity, vmax = c->max_velocity;
    if (c->type != RDB_COMPONENT_IDEAL_SPRING)
    {
      if (thr < 1e-6) thr = 1.0e-6;
      if (c->type == RDB_COMPONENT_FRICTION_POLY1 && vmax <= 0) vmax = 1.0e6;
      if (c->type == RDB_COMPONENT_FRICTION_POLY2 && vmax < 0) thr = 1.0e6;
    }
    const int nc = or_comp_cols(c->type);
    for (int64_t i = 0; i < n; i++)
    {
      double v[3] = {0, 0, 0};
      if (c->type == RDB_COMPONENT_IDEAL_SPRING)
      {
        v[0] = q ? q[(int64_t)c->input_index * ld + i] : 0.0;
        v[1] = 1.0;
      }
      else
      {
        const double d = dq ? dq[(int64_t)c->input_index * ld + i] : 0.0;
        const double omega = fmin(fmax(d, -vmax), vmax);
        if (c->type == RDB_COMPONENT_FRICTION_POLY1)
        {
          v[0] = fmin(fmax(omega / thr, -1.0), 1.0);
          v[1] = omega;
        }
        else
        {
          double sg;
          if (omega == 0) sg = 0;
          else if (omega > thr) sg = 1.0;
          else if (omega < -thr) sg = -1.0;
          else sg = omega / thr;
          v[0] = sg;
          v[1] = omega;
          v[2] = pow(omega, 2.0) * sg;
        }
      }
      for (int p = 0; p < nc; p++)
        for (int r = 0; r < n_in; r++) phi_c[((int64_t)(col + p) * n_in + r) * ld_out + i] = (r == c->input_index) ? v[p] : 0.0;
    }
    col += nc;
  }
}

void oracle_fill_uniform(double* x, int n_planes, int64_t n, int64_t ld, uint64_t seed, int stream_id)
{
  for (int j = 0; j < n_planes; j++)
    for (int64_t i = 0; i < n; i++)
    {
      uint64_t z = splitmix64(seed + ((uint64_t)i << 8) + ((uint64_t)stream_id << 6) + (uint64_t)j);
      x[(int64_t)j * ld + i] = 2.0 * ((double)(z >> 11) * (1.0 / 9007199254740992.0)) - 1.0;
    }
}
