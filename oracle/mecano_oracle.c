/*
 * mecano_oracle.c -- CPU restatement of Mecano's RNEA / ABA / CRBA (see mecano_oracle.h).
 *
 * TEST INFRASTRUCTURE ONLY; PARITY UNPINNED against the JVM (no JDK here, no golden vectors in
 * the reference).  Every function cites the reference file:line it follows.  "M/" below means
 * /root/reference/src/main/java/us/ihmc/mecano/.
 *
 * The restatement is deliberately reference-shaped rather than fast:
 *   - every frame keeps a transform-to-root, and every changeFrame() goes through the root like
 *     Euclid's ReferenceFrame.getTransformToDesiredFrame does (M/spatial/Twist.java:247-259);
 *   - RNEA works in body CoM frames, ABA and CRBA in frameAfterJoint, exactly like the reference;
 *   - built with -ffp-contract=off so products and sums round like Java doubles (no FMA).
 *
 * Euclid 0.21.0 and EJML 0.39 (the reference's un-vendored dependencies, build.gradle.kts:16-23)
 * supply only textbook 3-D / small dense linear algebra on this path; it is restated inline.
 */
#define _POSIX_C_SOURCE 200809L
#include "mecano_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

#define MO_MAXB 256

typedef struct { double R[9]; double t[3]; } xf_t; /* child coords -> parent coords: p' = R p + t */
typedef struct { double w[3]; double v[3]; } sv_t; /* angular part, linear part */

/* ------------------------------------------------------------------ 3-D helpers (Euclid) */

static void v3_cross(const double *a, const double *b, double *c)
{
   double x = a[1] * b[2] - a[2] * b[1];
   double y = a[2] * b[0] - a[0] * b[2];
   double z = a[0] * b[1] - a[1] * b[0];
   c[0] = x; c[1] = y; c[2] = z;
}

static void v3_add_cross(const double *a, const double *b, double *c) /* c += a x b */
{
   double t[3];
   v3_cross(a, b, t);
   c[0] += t[0]; c[1] += t[1]; c[2] += t[2];
}

static void m3_mulv(const double *M, const double *x, double *y)
{
   double a = M[0] * x[0] + M[1] * x[1] + M[2] * x[2];
   double b = M[3] * x[0] + M[4] * x[1] + M[5] * x[2];
   double c = M[6] * x[0] + M[7] * x[1] + M[8] * x[2];
   y[0] = a; y[1] = b; y[2] = c;
}

static void m3_tmulv(const double *M, const double *x, double *y) /* y = M^T x */
{
   double a = M[0] * x[0] + M[3] * x[1] + M[6] * x[2];
   double b = M[1] * x[0] + M[4] * x[1] + M[7] * x[2];
   double c = M[2] * x[0] + M[5] * x[1] + M[8] * x[2];
   y[0] = a; y[1] = b; y[2] = c;
}

static void m3_mul(const double *A, const double *B, double *C)
{
   double T[9];
   for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++)
         T[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
   memcpy(C, T, sizeof T);
}

static void m3_mul_bt(const double *A, const double *B, double *C) /* C = A B^T */
{
   double T[9];
   for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++)
         T[3 * i + j] = A[3 * i] * B[3 * j] + A[3 * i + 1] * B[3 * j + 1] + A[3 * i + 2] * B[3 * j + 2];
   memcpy(C, T, sizeof T);
}

/* R M R^T (RotationMatrix.transform(Matrix3D); MecanoTools.transformSymmetricMatrix3D, M/tools/MecanoTools.java:1053-1070) */
static void m3_rot_congruence(const double *R, double *M)
{
   double T[9];
   m3_mul(R, M, T);
   m3_mul_bt(T, R, M);
}

static void xf_identity(xf_t *x)
{
   memset(x, 0, sizeof *x);
   x->R[0] = x->R[4] = x->R[8] = 1.0;
}

/* RigidBodyTransform.multiply: (a o b)(p) = a(b(p)) */
static void xf_mul(const xf_t *a, const xf_t *b, xf_t *c)
{
   xf_t r;
   m3_mul(a->R, b->R, r.R);
   m3_mulv(a->R, b->t, r.t);
   r.t[0] += a->t[0]; r.t[1] += a->t[1]; r.t[2] += a->t[2];
   *c = r;
}

static void xf_inv(const xf_t *a, xf_t *c)
{
   xf_t r;
   for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++)
         r.R[3 * i + j] = a->R[3 * j + i];
   m3_mulv(r.R, a->t, r.t);
   r.t[0] = -r.t[0]; r.t[1] = -r.t[1]; r.t[2] = -r.t[2];
   *c = r;
}

/* ReferenceFrame.getTransformToDesiredFrame: inverse(desired.transformToRoot) * this.transformToRoot */
static void xf_rel(const xf_t *cur_to_root, const xf_t *desired_to_root, xf_t *out)
{
   xf_t inv;
   xf_inv(desired_to_root, &inv);
   xf_mul(&inv, cur_to_root, out);
}

/* ------------------------------------------------------------------ spatial vector transforms */

/* M/spatial/interfaces/FixedFrameSpatialMotionBasics.java:311-322 */
static void motion_apply(const xf_t *x, sv_t *m)
{
   m3_mulv(x->R, m->w, m->w);
   m3_mulv(x->R, m->v, m->v);
   v3_add_cross(x->t, m->w, m->v);
}

/* M/spatial/interfaces/FixedFrameSpatialMotionBasics.java:343-353 */
static void motion_apply_inv(const xf_t *x, sv_t *m)
{
   v3_add_cross(m->w, x->t, m->v);
   m3_tmulv(x->R, m->w, m->w);
   m3_tmulv(x->R, m->v, m->v);
}

/* M/spatial/interfaces/FixedFrameSpatialForceBasics.java:249-259 (w = moment, v = force) */
static void force_apply(const xf_t *x, sv_t *f)
{
   m3_mulv(x->R, f->w, f->w);
   m3_mulv(x->R, f->v, f->v);
   v3_add_cross(x->t, f->v, f->w);
}

static void sv_add(sv_t *a, const sv_t *b)
{
   for (int k = 0; k < 3; k++) { a->w[k] += b->w[k]; a->v[k] += b->v[k]; }
}

static void sv_sub(sv_t *a, const sv_t *b)
{
   for (int k = 0; k < 3; k++) { a->w[k] -= b->w[k]; a->v[k] -= b->v[k]; }
}

/* ------------------------------------------------------------------ joint transforms */

/* Euclid RotationMatrixConversion.convertAxisAngleToMatrix (Rodrigues), used by the general-axis
 * RevoluteJointTransformUpdater, M/tools/MecanoFactories.java:231-260.  The roll/pitch/yaw shortcuts
 * taken there for axis ~ X/Y/Z (eps 1e-7) are the same matrix with exact zeros. */
static void rot_axis_angle(const double *u, double q, double *R)
{
   const double eps = 1.0e-7;
   double c = cos(q), s = sin(q);
   if (fabs(u[0] - 1.0) < eps && fabs(u[1]) < eps && fabs(u[2]) < eps)
   {
      double T[9] = {1, 0, 0, 0, c, -s, 0, s, c};
      memcpy(R, T, sizeof T);
      return;
   }
   if (fabs(u[0]) < eps && fabs(u[1] - 1.0) < eps && fabs(u[2]) < eps)
   {
      double T[9] = {c, 0, s, 0, 1, 0, -s, 0, c};
      memcpy(R, T, sizeof T);
      return;
   }
   if (fabs(u[0]) < eps && fabs(u[1]) < eps && fabs(u[2] - 1.0) < eps)
   {
      double T[9] = {c, -s, 0, s, c, 0, 0, 0, 1};
      memcpy(R, T, sizeof T);
      return;
   }
   double n = sqrt(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]);
   double ux = u[0] / n, uy = u[1] / n, uz = u[2] / n;
   double t = 1.0 - c;
   double xy = t * ux * uy, xz = t * ux * uz, yz = t * uy * uz;
   double sx = s * ux, sy = s * uy, sz = s * uz;
   R[0] = t * ux * ux + c; R[1] = xy - sz;         R[2] = xz + sy;
   R[3] = xy + sz;         R[4] = t * uy * uy + c; R[5] = yz - sx;
   R[6] = xz - sy;         R[7] = yz + sx;         R[8] = t * uz * uz + c;
}

/* Euclid RotationMatrixConversion.convertQuaternionToMatrix; the quaternion is normalised on set
 * (M/multiBodySystem/interfaces/SixDoFJointBasics.java:104-109). */
static void rot_quaternion(const double *q4, double *R)
{
   double qx = q4[0], qy = q4[1], qz = q4[2], qs = q4[3];
   double n = sqrt(qx * qx + qy * qy + qz * qz + qs * qs);
   if (n < 1.0e-14)
   {
      double I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
      memcpy(R, I, sizeof I);
      return;
   }
   n = 1.0 / n;
   qx *= n; qy *= n; qz *= n; qs *= n;
   double yy2 = 2.0 * qy * qy, zz2 = 2.0 * qz * qz, xx2 = 2.0 * qx * qx;
   double xy2 = 2.0 * qx * qy, sz2 = 2.0 * qs * qz, xz2 = 2.0 * qx * qz;
   double sy2 = 2.0 * qs * qy, yz2 = 2.0 * qy * qz, sx2 = 2.0 * qs * qx;
   R[0] = 1.0 - yy2 - zz2; R[1] = xy2 - sz2;       R[2] = xz2 + sy2;
   R[3] = xy2 + sz2;       R[4] = 1.0 - xx2 - zz2; R[5] = yz2 - sx2;
   R[6] = xz2 - sy2;       R[7] = yz2 + sx2;       R[8] = 1.0 - xx2 - yy2;
}

/* joint configuration -> transform of frameAfterJoint in frameBeforeJoint
 * revolute M/tools/MecanoFactories.java:231-260, prismatic M/multiBodySystem/interfaces/PrismaticJointReadOnly.java:18-22,
 * floating M/multiBodySystem/interfaces/FloatingJointReadOnly.java:34-37 */
static void joint_transform(const mo_tree *t, int i, const double *q, xf_t *x)
{
   const double *u = t->axis + 3 * i;
   const double *qi = q + t->cfg_off[i];
   xf_identity(x);
   switch (t->jtype[i])
   {
      case MO_REVOLUTE:
         rot_axis_angle(u, qi[0], x->R);
         break;
      case MO_PRISMATIC:
         x->t[0] = qi[0] * u[0]; x->t[1] = qi[0] * u[1]; x->t[2] = qi[0] * u[2];
         break;
      case MO_SPHERICAL: /* SphericalJointReadOnly.java:31-35: setRotationAndZeroTranslation(jointOrientation) */
         rot_quaternion(qi, x->R);
         break;
      case MO_PLANAR:
      {
         /* PlanarJointReadOnly.java:40-48 (pitch, x, z) of a planar pose: rotation about y, translation in the x-z plane
          * (M/tools/MecanoFactories.java newPlanarPose3DBasics) */
         const double c = cos(qi[0]), s = sin(qi[0]);
         double T[9] = {c, 0, s, 0, 1, 0, -s, 0, c};
         memcpy(x->R, T, sizeof T);
         x->t[0] = qi[1]; x->t[1] = 0.0; x->t[2] = qi[2];
         break;
      }
      default:
         rot_quaternion(qi, x->R);
         x->t[0] = qi[4]; x->t[1] = qi[5]; x->t[2] = qi[6];
   }
}

static int joint_ndof(const mo_tree *t, int i)
{
   switch (t->jtype[i])
   {
      case MO_SIXDOF: return 6;
      case MO_SPHERICAL: case MO_PLANAR: return 3;
      default: return 1;
   }
}

/* S * x for this joint, expressed in frameAfterJoint (motion subspace: M/multiBodySystem/interfaces/JointReadOnly.java:201-207,
 * M/tools/MecanoTools.java:964-995: revolute [axis;0], prismatic [0;axis], SixDoF identity) */
static void joint_S_times(const mo_tree *t, int i, const double *x, sv_t *out)
{
   const double *u = t->axis + 3 * i;
   memset(out, 0, sizeof *out);
   switch (t->jtype[i])
   {
      case MO_REVOLUTE:
         out->w[0] = u[0] * x[0]; out->w[1] = u[1] * x[0]; out->w[2] = u[2] * x[0];
         break;
      case MO_PRISMATIC:
         out->v[0] = u[0] * x[0]; out->v[1] = u[1] * x[0]; out->v[2] = u[2] * x[0];
         break;
      case MO_SPHERICAL: /* unit twists = the three angular components (SphericalJoint.java, MecanoTools.computeSphericalJointMotionSubspace) */
         for (int k = 0; k < 3; k++) out->w[k] = x[k];
         break;
      case MO_PLANAR: /* twistComponentIndex = {1, 3, 5}: w_y, v_x, v_z (MecanoTools.java:938) */
         out->w[1] = x[0]; out->v[0] = x[1]; out->v[2] = x[2];
         break;
      default:
         for (int k = 0; k < 3; k++) { out->w[k] = x[k]; out->v[k] = x[3 + k]; }
   }
}

/* column k of S as a 6-vector */
static void joint_S_col(const mo_tree *t, int i, int k, double *col)
{
   double e[6] = {0, 0, 0, 0, 0, 0};
   sv_t s;
   e[k] = 1.0;
   joint_S_times(t, i, e, &s);
   memcpy(col, s.w, 3 * sizeof(double));
   memcpy(col + 3, s.v, 3 * sizeof(double));
}

/* ------------------------------------------------------------------ frame tree
 * M/multiBodySystem/interfaces/RigidBodyBasics.java:104-112 (updateFramesRecursively),
 * M/tools/MecanoFactories.java:61-125 (frame construction), M/frames/MovingReferenceFrame.java:279-311 (twist of frame) */

typedef struct
{
   xf_t after[MO_MAXB]; /* frameAfterJoint  -> root */
   xf_t com[MO_MAXB];   /* bodyFixedFrame   -> root */
   sv_t tw_after[MO_MAXB]; /* twist of frameAfterJoint wrt world, in frameAfterJoint */
   sv_t tw_com[MO_MAXB];   /* twist of bodyFixedFrame wrt world, in bodyFixedFrame */
   xf_t root;              /* root body frame (identity) */
} frames_t;

static void update_frames(const mo_tree *t, const double *q, const double *qd, frames_t *F)
{
   xf_identity(&F->root);
   for (int i = 0; i < t->nb; i++)
   {
      int p = t->parent[i];
      const xf_t *parent_after = p < 0 ? &F->root : &F->after[p];
      xf_t off, before, xj, pose, rel;
      memcpy(off.R, t->off_R + 9 * i, sizeof off.R);
      memcpy(off.t, t->off_p + 3 * i, sizeof off.t);
      xf_mul(parent_after, &off, &before); /* frameBeforeJoint -> root */
      joint_transform(t, i, q, &xj);
      xf_mul(&before, &xj, &F->after[i]);
      memcpy(pose.R, t->com_R + 9 * i, sizeof pose.R);
      memcpy(pose.t, t->com_p + 3 * i, sizeof pose.t);
      xf_mul(&F->after[i], &pose, &F->com[i]);

      if (qd)
      {
         /* frameBeforeJoint is fixed in the parent's frameAfterJoint: its twist is the parent's, re-expressed */
         sv_t tw;
         if (p < 0)
            memset(&tw, 0, sizeof tw);
         else
         {
            tw = F->tw_after[p];
            xf_rel(parent_after, &before, &rel);
            motion_apply(&rel, &tw);
         }
         /* frameAfterJoint: parent's twist changed to this frame + twist relative to parent (the joint twist) */
         xf_rel(&before, &F->after[i], &rel);
         motion_apply(&rel, &tw);
         sv_t jt;
         joint_S_times(t, i, qd + t->dof_off[i], &jt);
         sv_add(&tw, &jt);
         F->tw_after[i] = tw;
         /* bodyFixedFrame is fixed in frameAfterJoint */
         xf_rel(&F->after[i], &F->com[i], &rel);
         motion_apply(&rel, &tw);
         F->tw_com[i] = tw;
      }
      else
      {
         memset(&F->tw_after[i], 0, sizeof(sv_t));
         memset(&F->tw_com[i], 0, sizeof(sv_t));
      }
   }
}

/* ------------------------------------------------------------------ Newton-Euler wrench at the CoM
 * M/spatial/interfaces/SpatialInertiaReadOnly.java:229-277 (CoM offset zero branch) with
 * M/tools/MecanoTools.java:571-598 (computeDynamicMomentFast) and :728-752 (computeDynamicForceFast).
 * acc / tw may be NULL. */
static void dynamic_wrench(const double *J, double m, const sv_t *acc, const sv_t *tw, sv_t *W)
{
   if (tw)
   {
      m3_mulv(J, tw->w, W->w);       /* J w */
      v3_cross(tw->w, W->w, W->w);   /* w x J w */
   }
   else
      W->w[0] = W->w[1] = W->w[2] = 0.0;
   if (acc)
   {
      double mx = W->w[0], my = W->w[1], mz = W->w[2];
      m3_mulv(J, acc->w, W->w); /* J wdot */
      W->w[0] += mx; W->w[1] += my; W->w[2] += mz;
   }
   if (tw)
   {
      v3_cross(tw->w, tw->v, W->v);
      if (acc) { W->v[0] += acc->v[0]; W->v[1] += acc->v[1]; W->v[2] += acc->v[2]; }
      W->v[0] *= m; W->v[1] *= m; W->v[2] *= m;
   }
   else
   {
      W->v[0] = W->v[1] = W->v[2] = 0.0;
      if (acc) { W->v[0] = acc->v[0] * m; W->v[1] = acc->v[1] * m; W->v[2] = acc->v[2] * m; }
   }
}

/* ================================================================== RNEA
 * M/algorithms/InverseDynamicsCalculator.java:496-501 compute(), :873-917 passOne(), :930-966 passTwo() */

static void rnea_impl(const mo_tree *t, const double *g, const double *q, const double *qd, const double *qdd, const double *fext,
                      int flags, double *tau, double *acc_out, double *wr_out)
{
   frames_t *F = (frames_t *)malloc(sizeof(frames_t));
   sv_t *acc = (sv_t *)malloc(sizeof(sv_t) * (size_t)t->nb);
   sv_t *wr = (sv_t *)malloc(sizeof(sv_t) * (size_t)t->nb);
   int coriolis = !(flags & MO_NO_CORIOLIS), accel = !(flags & MO_NO_ACCELERATIONS);
   xf_t rel;

   update_frames(t, q, qd, F);

   /* setGravitationalAcceleration: root acceleration = -gravity, :397-403 */
   sv_t root_acc;
   memset(&root_acc, 0, sizeof root_acc);
   root_acc.v[0] = -g[0]; root_acc.v[1] = -g[1]; root_acc.v[2] = -g[2];

   /* pass one: bodies are listed in DFS pre-order, so a plain loop visits them like passOneRecursive() */
   for (int i = 0; i < t->nb; i++)
   {
      int p = t->parent[i];
      const xf_t *pred = p < 0 ? &F->root : &F->com[p]; /* predecessor bodyFixedFrame */
      sv_t a = p < 0 ? root_acc : acc[p];              /* :880 */

      if (coriolis)
      {
         /* joint.getPredecessorTwist(): -(joint twist) re-expressed in the predecessor frame
          * M/multiBodySystem/interfaces/JointReadOnly.java:270-281, OneDoFJointReadOnly.java:292-296 */
         sv_t d;
         joint_S_times(t, i, qd + t->dof_off[i], &d);
         for (int k = 0; k < 3; k++) { d.w[k] = -d.w[k]; d.v[k] = -d.v[k]; }
         xf_rel(&F->after[i], pred, &rel);
         motion_apply(&rel, &d);
         /* SpatialAccelerationBasics.changeFrame(desired, deltaTwist, bodyTwist), cross products first
          * M/spatial/interfaces/SpatialAccelerationBasics.java:179-202 (flipCrossProducts == false) */
         if (p >= 0)
         {
            const sv_t *b = &F->tw_com[p];
            v3_add_cross(d.v, b->w, a.v); /* v_old x omega_body */
            v3_add_cross(d.w, b->v, a.v); /* omega_old x v_body */
            v3_add_cross(d.w, b->w, a.w); /* omega_old x omega_body */
         }
      }
      xf_rel(pred, &F->com[i], &rel);
      motion_apply(&rel, &a); /* changeFrame(bodyFixedFrame) */

      if (accel)
      {
         /* :894-910 */
         sv_t ja;
         joint_S_times(t, i, qdd + t->dof_off[i], &ja);
         xf_rel(&F->after[i], &F->com[i], &rel);
         motion_apply(&rel, &ja);
         sv_add(&a, &ja);
      }
      acc[i] = a;
   }
   if (acc_out)
      memcpy(acc_out, acc, sizeof(sv_t) * (size_t)t->nb);

   /* pass two: reverse pre-order visits children before parents like passTwoRecursive(); the order in which
    * sibling wrenches are added differs from the Java recursion only in rounding. To keep the child order of
    * :949-950 we accumulate children in increasing index below. */
   if (tau)
   {
      for (int i = t->nb - 1; i >= 0; i--)
      {
         sv_t W;
         dynamic_wrench(t->J + 9 * i, t->mass[i], &acc[i], coriolis ? &F->tw_com[i] : NULL, &W); /* :943 */
         if (fext)
         {
            sv_t e;
            memcpy(e.w, fext + 6 * i, 3 * sizeof(double));
            memcpy(e.v, fext + 6 * i + 3, 3 * sizeof(double));
            sv_sub(&W, &e); /* :946 */
         }
         xf_rel(&F->com[i], &F->after[i], &rel);
         force_apply(&rel, &W); /* :947 */
         for (int c = i + 1; c < t->nb; c++)
            if (t->parent[c] == i)
            {
               sv_t Wc = wr[c]; /* :961-966 */
               xf_rel(&F->after[c], &F->after[i], &rel);
               force_apply(&rel, &Wc);
               sv_add(&W, &Wc);
            }
         wr[i] = W;
         if (wr_out) /* getComputedJointWrench(joint): the joint wrench, expressed in frameAfterJoint (:947, :593-602) */
            memcpy(wr_out + 6 * i, &W, sizeof(sv_t));
         /* tau = S^T W, :952-958 */
         int nd = joint_ndof(t, i);
         for (int k = 0; k < nd; k++)
         {
            double col[6];
            joint_S_col(t, i, k, col);
            tau[t->dof_off[i] + k] = col[0] * W.w[0] + col[1] * W.w[1] + col[2] * W.w[2] + col[3] * W.v[0] + col[4] * W.v[1]
                                     + col[5] * W.v[2];
         }
      }
   }
   free(F); free(acc); free(wr);
}

void mo_rnea(const mo_tree *t, const double *g, const double *q, const double *qd, const double *qdd, const double *fext, int flags,
             double *tau)
{
   rnea_impl(t, g, q, qd, qdd, fext, flags, tau, NULL, NULL);
}

void mo_rnea_full(const mo_tree *t, const double *g, const double *q, const double *qd, const double *qdd, const double *fext, int flags,
                  double *tau, double *body_acc, double *joint_wrench)
{
   rnea_impl(t, g, q, qd, qdd, fext, flags, tau, body_acc, joint_wrench);
}

void mo_rnea_body_accelerations(const mo_tree *t, const double *g, const double *q, const double *qd, const double *qdd, int flags,
                                double *acc)
{
   rnea_impl(t, g, q, qd, qdd, NULL, flags, NULL, acc, NULL);
}

/* ================================================================== articulated-body inertia
 * M/algorithms/ArticulatedBodyInertia.java: three 3x3 blocks, 6x6 = [[A, C],[C^T, L]] (see mult() in
 * M/algorithms/ForwardDynamicsCalculator.java:1448-1505) */

typedef struct { double A[9], L[9], C[9]; } abi_t;

/* SpatialInertia (moment about the frame origin, mass, CoM offset) */
typedef struct { double I[9]; double m; double c[3]; } si_t;

/* M/spatial/interfaces/SpatialInertiaBasics.java:222-239 with M/tools/MecanoTools.java:483-547 */
static void si_apply(const xf_t *x, si_t *s)
{
   m3_rot_congruence(x->R, s->I);
   m3_mulv(x->R, s->c, s->c);
   {
      double xp = x->t[0], yp = x->t[1], zp = x->t[2];
      double xc = s->c[0], yc = s->c[1], zc = s->c[2], m = s->m;
      double xp_xp = xp * xp, yp_yp = yp * yp, zp_zp = zp * zp;
      double two_xc_xp = 2.0 * xc * xp, two_yc_yp = 2.0 * yc * yp, two_zc_zp = 2.0 * zc * zp;
      double txx = m * (two_yc_yp + two_zc_zp + yp_yp + zp_zp);
      double tyy = m * (two_xc_xp + two_zc_zp + xp_xp + zp_zp);
      double tzz = m * (two_xc_xp + two_yc_yp + xp_xp + yp_yp);
      double txy = m * (-xc * yp - yc * xp - xp * yp);
      double txz = m * (-xc * zp - zc * xp - xp * zp);
      double tyz = m * (-yc * zp - zc * yp - yp * zp);
      s->I[0] += txx; s->I[1] += txy; s->I[2] += txz;
      s->I[3] += txy; s->I[4] += tyy; s->I[5] += tyz;
      s->I[6] += txz; s->I[7] += tyz; s->I[8] += tzz;
      s->c[0] += xp; s->c[1] += yp; s->c[2] += zp;
   }
}

/* M/spatial/interfaces/FixedFrameSpatialInertiaBasics.java:167-176 */
static void si_add(si_t *a, const si_t *b)
{
   for (int k = 0; k < 9; k++) a->I[k] += b->I[k];
   for (int k = 0; k < 3; k++) a->c[k] = a->c[k] * a->m;
   for (int k = 0; k < 3; k++) a->c[k] = b->m * b->c[k] + a->c[k];
   a->m = a->m + b->m;
   if (fabs(a->m) >= 1.0e-7)
   {
      double inv = 1.0 / a->m;
      for (int k = 0; k < 3; k++) a->c[k] *= inv;
   }
}

/* body inertia (about CoM, in bodyFixedFrame) re-expressed in frameAfterJoint: spatialInertia.changeFrame(frameAfterJoint) */
static void body_inertia_at_after(const mo_tree *t, int i, const frames_t *F, si_t *s)
{
   xf_t rel;
   memcpy(s->I, t->J + 9 * i, sizeof s->I);
   s->m = t->mass[i];
   s->c[0] = s->c[1] = s->c[2] = 0.0;
   xf_rel(&F->com[i], &F->after[i], &rel);
   si_apply(&rel, s);
}

/* M/algorithms/ArticulatedBodyInertia.java:176-186 */
static void abi_from_si(const si_t *s, abi_t *a)
{
   memcpy(a->A, s->I, sizeof a->A);
   memset(a->L, 0, sizeof a->L);
   a->L[0] = a->L[4] = a->L[8] = s->m;
   /* tilde(c) * m */
   double x = s->c[0], y = s->c[1], z = s->c[2];
   double T[9] = {0, -z, y, z, 0, -x, -y, x, 0};
   for (int k = 0; k < 9; k++) a->C[k] = T[k] * s->m;
}

/* M/algorithms/ArticulatedBodyInertia.java:359-375 with ArticulatedBodyInertiaAlorigthmTools.java:32-163 */
static void abi_apply(const xf_t *xf, abi_t *a)
{
   m3_rot_congruence(xf->R, a->A);
   m3_rot_congruence(xf->R, a->L);
   m3_rot_congruence(xf->R, a->C);
   {
      double x = xf->t[0], y = xf->t[1], z = xf->t[2];
      double xx = x * x, yy = y * y, zz = z * z, xy = x * y, xz = x * z, yz = y * z;
      double mxx = a->L[0], myy = a->L[4], mzz = a->L[8], mxy = a->L[1], mxz = a->L[2], myz = a->L[5];
      double ixx = a->A[0], iyy = a->A[4], izz = a->A[8], ixy = a->A[1], ixz = a->A[2], iyz = a->A[5];
      double c00 = a->C[0], c01 = a->C[1], c02 = a->C[2], c10 = a->C[3], c11 = a->C[4], c12 = a->C[5], c20 = a->C[6],
             c21 = a->C[7], c22 = a->C[8];

      ixx += yy * mzz + zz * myy - 2.0 * yz * myz;
      iyy += xx * mzz + zz * mxx - 2.0 * xz * mxz;
      izz += xx * myy + yy * mxx - 2.0 * xy * mxy;
      ixy += -zz * mxy - xy * mzz + xz * myz + yz * mxz;
      ixz += -yy * mxz + xy * myz - xz * myy + yz * mxy;
      iyz += -xx * myz + xy * mxz + xz * mxy - yz * mxx;

      ixx += 2.0 * (y * c02 - z * c01);
      iyy += 2.0 * (-x * c12 + z * c10);
      izz += 2.0 * (x * c21 - y * c20);
      ixy += -x * c02 + y * c12 + z * (c00 - c11);
      ixz += x * c01 - y * (c00 - c22) - z * c21;
      iyz += x * (c11 - c22) - y * c10 + z * c20;

      a->A[0] = ixx; a->A[1] = ixy; a->A[2] = ixz;
      a->A[3] = ixy; a->A[4] = iyy; a->A[5] = iyz;
      a->A[6] = ixz; a->A[7] = iyz; a->A[8] = izz;

      /* translateCrossInertia */
      c00 += y * mxz - z * mxy;
      c11 += -x * myz + z * mxy;
      c22 += x * myz - y * mxz;
      c01 += y * myz - z * myy;
      c02 += y * mzz - z * myz;
      c12 += -x * mzz + z * mxz;
      c10 += -x * mxz + z * mxx;
      c20 += x * mxy - y * mxx;
      c21 += x * myy - y * mxy;
      a->C[0] = c00; a->C[1] = c01; a->C[2] = c02;
      a->C[3] = c10; a->C[4] = c11; a->C[5] = c12;
      a->C[6] = c20; a->C[7] = c21; a->C[8] = c22;
   }
}

/* y = IA x (6-vectors), M/algorithms/ForwardDynamicsCalculator.java:1448-1505 */
static void abi_mulv(const abi_t *a, const double *x, double *y)
{
   double r[6];
   for (int i = 0; i < 3; i++)
   {
      r[i] = a->A[3 * i] * x[0] + a->A[3 * i + 1] * x[1] + a->A[3 * i + 2] * x[2] + a->C[3 * i] * x[3] + a->C[3 * i + 1] * x[4]
             + a->C[3 * i + 2] * x[5];
      r[3 + i] = a->C[i] * x[0] + a->C[3 + i] * x[1] + a->C[6 + i] * x[2] + a->L[3 * i] * x[3] + a->L[3 * i + 1] * x[4]
                 + a->L[3 * i + 2] * x[5];
   }
   memcpy(y, r, sizeof r);
}

/* inverse of a symmetric positive-definite n x n matrix by Cholesky (EJML LinearSolverFactory_DDRM.symmPosDef(6),
 * M/algorithms/ForwardDynamicsCalculator.java:1040, :1193-1197); n <= 6 */
static void spd_inverse(int n, const double *D, double *Dinv)
{
   double L[36], Y[36];
   memset(L, 0, sizeof L);
   for (int i = 0; i < n; i++)
      for (int j = 0; j <= i; j++)
      {
         double s = D[i * n + j];
         for (int k = 0; k < j; k++) s -= L[i * n + k] * L[j * n + k];
         L[i * n + j] = (i == j) ? sqrt(s) : s / L[j * n + j];
      }
   /* solve L Y = I, then L^T X = Y */
   for (int c = 0; c < n; c++)
   {
      for (int i = 0; i < n; i++)
      {
         double s = (i == c) ? 1.0 : 0.0;
         for (int k = 0; k < i; k++) s -= L[i * n + k] * Y[k * n + c];
         Y[i * n + c] = s / L[i * n + i];
      }
      for (int i = n - 1; i >= 0; i--)
      {
         double s = Y[i * n + c];
         for (int k = i + 1; k < n; k++) s -= L[k * n + i] * Dinv[k * n + c];
         Dinv[i * n + c] = s / L[i * n + i];
      }
   }
}

/* ================================================================== ABA
 * M/algorithms/ForwardDynamicsCalculator.java:508-520 compute(), :1085-1127 passOne(), :1136-1254 passTwo(),
 * :1259-1310 passThree() (all joints EFFORT_SOURCE) */

typedef struct
{
   xf_t X;       /* transformToParentJointFrame */
   sv_t p;       /* biasWrench in frameAfterJoint */
   double c[6];  /* biasAcceleration */
   abi_t Ia;     /* articulatedInertiaForParent */
   double pa[6]; /* articulatedBiasWrenchForParent */
   double S[36]; /* 6 x nd, column k at S[6*k..] */
   double U[36]; /* 6 x nd, column-wise */
   double Dinv[36];
   double u[6];
   int nd;
} aba_step_t;

/* accsrc (nullable, [nb]): non-zero = the joint is an ACCELERATION_SOURCE (ForwardDynamicsCalculator.java:45-57, :400-403):
 * its acceleration is taken from qdd_in and its effort is computed (pass four, :1315-1363) into tau_out (nullable, [nv]; the
 * rows of EFFORT_SOURCE joints repeat the input, getJointTauMatrix() :566-590). */
static void aba_impl(const mo_tree *t, const double *g, const double *q, const double *qd, const double *tau, const double *fext,
                     const int *accsrc, const double *qdd_in, double *qdd, double *tau_out)
{
   frames_t *F = (frames_t *)malloc(sizeof(frames_t));
   aba_step_t *st = (aba_step_t *)malloc(sizeof(aba_step_t) * (size_t)t->nb);
   sv_t *acc = (sv_t *)malloc(sizeof(sv_t) * (size_t)t->nb);
   xf_t rel;

   update_frames(t, q, qd, F);

   /* pass one */
   for (int i = 0; i < t->nb; i++)
   {
      aba_step_t *s = &st[i];
      int p = t->parent[i];
      s->nd = joint_ndof(t, i);
      for (int k = 0; k < s->nd; k++) joint_S_col(t, i, k, s->S + 6 * k);
      xf_rel(&F->after[i], p < 0 ? &F->root : &F->after[p], &s->X); /* :1096-1099 */

      dynamic_wrench(t->J + 9 * i, t->mass[i], NULL, &F->tw_com[i], &s->p); /* :1109 */
      if (fext)
      {
         sv_t e;
         memcpy(e.w, fext + 6 * i, 3 * sizeof(double));
         memcpy(e.v, fext + 6 * i + 3, 3 * sizeof(double));
         sv_sub(&s->p, &e); /* :1111 */
      }
      xf_rel(&F->com[i], &F->after[i], &rel);
      force_apply(&rel, &s->p); /* :1112 */

      /* :1114-1118 biasAcceleration: zero, then changeFrame(after, jointTwist, twistOfFrame(after)); both twists are
       * expressed in the desired frame and deltaTwist's body is the desired frame -> flipCrossProducts == true,
       * M/spatial/interfaces/SpatialAccelerationBasics.java:204-217 */
      {
         sv_t jt, c;
         const sv_t *b = &F->tw_after[i];
         joint_S_times(t, i, qd + t->dof_off[i], &jt);
         memset(&c, 0, sizeof c);
         v3_add_cross(b->w, jt.v, c.v); /* omega_body x v_new */
         v3_add_cross(b->v, jt.w, c.v); /* v_body x omega_new */
         v3_add_cross(b->w, jt.w, c.w); /* omega_body x omega_new */
         memcpy(s->c, c.w, 3 * sizeof(double));
         memcpy(s->c + 3, c.v, 3 * sizeof(double));
      }
   }

   /* pass two, leaves to root */
   for (int i = t->nb - 1; i >= 0; i--)
   {
      aba_step_t *s = &st[i];
      si_t si;
      abi_t IA;
      sv_t pA = s->p; /* :1153-1154 (already in frameAfterJoint) */
      int nd = s->nd;

      body_inertia_at_after(t, i, F, &si); /* :1149-1150 */
      abi_from_si(&si, &IA);               /* :1151 */

      for (int c = i + 1; c < t->nb; c++)
         if (t->parent[c] == i)
         {
            /* :1156-1166 */
            aba_step_t *ch = &st[c];
            abi_t Ic = ch->Ia;
            sv_t pc;
            memcpy(pc.w, ch->pa, 3 * sizeof(double));
            memcpy(pc.v, ch->pa + 3, 3 * sizeof(double));
            abi_apply(&ch->X, &Ic);
            force_apply(&ch->X, &pc);
            for (int k = 0; k < 9; k++) { IA.A[k] += Ic.A[k]; IA.L[k] += Ic.L[k]; IA.C[k] += Ic.C[k]; }
            sv_add(&pA, &pc);
         }

      /* U = IA S, D = S^T U, :1176-1179 */
      double D[36];
      for (int k = 0; k < nd; k++) abi_mulv(&IA, s->S + 6 * k, s->U + 6 * k);
      for (int a = 0; a < nd; a++)
         for (int b = 0; b < nd; b++)
         {
            double d = 0.0;
            for (int r = 0; r < 6; r++) d += s->S[6 * a + r] * s->U[6 * b + r];
            D[a * nd + b] = d;
         }
      if (nd == 1)
         s->Dinv[0] = 1.0 / D[0]; /* :1181-1184 */
      else
         spd_inverse(nd, D, s->Dinv); /* :1193-1197 */

      /* u = tau - S^T pA, :1199-1215 */
      double pAv[6];
      memcpy(pAv, pA.w, 3 * sizeof(double));
      memcpy(pAv + 3, pA.v, 3 * sizeof(double));
      for (int k = 0; k < nd; k++)
      {
         double d = 0.0;
         for (int r = 0; r < 6; r++) d += -1.0 * s->S[6 * k + r] * pAv[r];
         s->u[k] = d + tau[t->dof_off[i] + k];
      }

      if (accsrc && accsrc[i])
      {
         /* :1237-1253: nothing is removed from the articulated inertia, the known joint acceleration enters the bias wrench */
         if (t->parent[i] >= 0)
         {
            double ca[6], Iac[6];
            s->Ia = IA;
            for (int r = 0; r < 6; r++)
            {
               double d = 0.0;
               for (int k = 0; k < nd; k++) d += s->S[6 * k + r] * qdd_in[t->dof_off[i] + k]; /* getJointAcceleration().get(a) */
               ca[r] = d;
            }
            abi_mulv(&s->Ia, s->c, Iac);
            for (int r = 0; r < 6; r++) s->pa[r] = pAv[r] + Iac[r];
            abi_mulv(&s->Ia, ca, Iac);
            for (int r = 0; r < 6; r++) s->pa[r] += Iac[r];
         }
      }
      else if (t->parent[i] >= 0)
      {
         /* :1217-1235 */
         double UD[36]; /* U Dinv, 6 x nd column-wise */
         double UDU[36];
         for (int k = 0; k < nd; k++)
            for (int r = 0; r < 6; r++)
            {
               double d = 0.0;
               for (int m = 0; m < nd; m++) d += s->U[6 * m + r] * s->Dinv[m * nd + k];
               UD[6 * k + r] = d;
            }
         for (int r = 0; r < 6; r++)
            for (int c = 0; c < 6; c++)
            {
               double d = 0.0;
               for (int k = 0; k < nd; k++) d += UD[6 * k + r] * s->U[6 * k + c];
               UDU[6 * r + c] = d;
            }
         s->Ia = IA;
         for (int r = 0; r < 3; r++)
            for (int c = 0; c < 3; c++)
            {
               s->Ia.A[3 * r + c] -= UDU[6 * r + c];
               s->Ia.C[3 * r + c] -= UDU[6 * r + 3 + c];
               s->Ia.L[3 * r + c] -= UDU[6 * (r + 3) + 3 + c];
            }
         double Iac[6];
         abi_mulv(&s->Ia, s->c, Iac);
         for (int r = 0; r < 6; r++)
         {
            double d = pAv[r] + Iac[r];
            for (int k = 0; k < nd; k++) d += UD[6 * k + r] * s->u[k];
            s->pa[r] = d;
         }
      }
   }

   /* pass three, root to leaves */
   sv_t root_acc;
   memset(&root_acc, 0, sizeof root_acc);
   root_acc.v[0] = -g[0]; root_acc.v[1] = -g[1]; root_acc.v[2] = -g[2];
   for (int i = 0; i < t->nb; i++)
   {
      aba_step_t *s = &st[i];
      int p = t->parent[i], nd = s->nd;
      sv_t a = p < 0 ? root_acc : acc[p];
      motion_apply_inv(&s->X, &a); /* :1270-1272 */
      double av[6];
      for (int k = 0; k < 3; k++) { av[k] = a.w[k] + s->c[k]; av[3 + k] = a.v[k] + s->c[3 + k]; } /* :1273 */
      /* qdd = Dinv (u - U^T a'), :1279-1282 */
      double tmp[6], qddi[6];
      for (int k = 0; k < nd; k++)
      {
         double d = 0.0;
         for (int r = 0; r < 6; r++) d += -1.0 * s->U[6 * k + r] * av[r];
         tmp[k] = d + s->u[k];
      }
      for (int k = 0; k < nd; k++)
      {
         double d = 0.0;
         for (int m = 0; m < nd; m++) d += s->Dinv[k * nd + m] * tmp[m];
         if (accsrc && accsrc[i])
            d = qdd_in[t->dof_off[i] + k]; /* :1286-1298 */
         qddi[k] = d;
         qdd[t->dof_off[i] + k] = d;
      }
      /* a = a' + S qdd, :1299-1305 */
      for (int r = 0; r < 6; r++)
      {
         double d = 0.0;
         for (int k = 0; k < nd; k++) d += s->S[6 * k + r] * qddi[k];
         av[r] += d;
      }
      memcpy(acc[i].w, av, 3 * sizeof(double));
      memcpy(acc[i].v, av + 3, 3 * sizeof(double));
   }

   if (tau_out)
   {
      /* pass four (:1315-1363): joint wrenches from the body accelerations, leaves to root.  The body inertia is taken in the
       * body's CoM frame (zero CoM offset), so the velocity terms are the bias wrench of pass one (:1335-1340) */
      sv_t *wr = (sv_t *)malloc(sizeof(sv_t) * (size_t)t->nb);
      for (int i = t->nb - 1; i >= 0; i--)
      {
         aba_step_t *s = &st[i];
         sv_t a = acc[i], W;
         xf_rel(&F->after[i], &F->com[i], &rel);
         motion_apply(&rel, &a); /* rigidBodyAcceleration.changeFrame(getBodyFixedFrame()) */
         dynamic_wrench(t->J + 9 * i, t->mass[i], &a, NULL, &W);
         xf_rel(&F->com[i], &F->after[i], &rel);
         force_apply(&rel, &W);
         sv_add(&W, &s->p);
         for (int c = i + 1; c < t->nb; c++)
            if (t->parent[c] == i)
            {
               sv_t Wc = wr[c]; /* addJointWrenchFromChild :1358-1363 */
               force_apply(&st[c].X, &Wc);
               sv_add(&W, &Wc);
            }
         wr[i] = W;
         for (int k = 0; k < s->nd; k++)
         {
            if (accsrc && accsrc[i])
            {
               const double *col = s->S + 6 * k;
               tau_out[t->dof_off[i] + k] = col[0] * W.w[0] + col[1] * W.w[1] + col[2] * W.w[2] + col[3] * W.v[0] + col[4] * W.v[1] + col[5] * W.v[2];
            }
            else
               tau_out[t->dof_off[i] + k] = tau[t->dof_off[i] + k];
         }
      }
      free(wr);
   }
   free(F); free(st); free(acc);
}

void mo_aba(const mo_tree *t, const double *g, const double *q, const double *qd, const double *tau, const double *fext, double *qdd)
{
   aba_impl(t, g, q, qd, tau, fext, NULL, NULL, qdd, NULL);
}

void mo_aba_sources(const mo_tree *t, const double *g, const double *q, const double *qd, const double *tau, const double *qdd_in,
                    const double *fext, const int *accel_source, double *qdd, double *tau_out)
{
   aba_impl(t, g, q, qd, tau, fext, accel_source, qdd_in, qdd, tau_out);
}

/* ================================================================== CRBA
 * M/algorithms/CompositeRigidBodyMassMatrixCalculator.java:286-303 (reset/update), :588-667, :700-707, :772-797 */

/* Momentum.compute(inertia, twist) = inertia.transform(twist), M/spatial/interfaces/SpatialInertiaReadOnly.java:334-357 */
static void si_momentum(const si_t *s, const double *tw6, double *h6)
{
   const double *w = tw6, *v = tw6 + 3;
   if (s->c[0] == 0.0 && s->c[1] == 0.0 && s->c[2] == 0.0)
   {
      m3_mulv(s->I, w, h6);
      h6[3] = s->m * v[0]; h6[4] = s->m * v[1]; h6[5] = s->m * v[2];
      return;
   }
   double ang[3], lin[3], Iw[3];
   v3_cross(s->c, v, ang);
   ang[0] *= s->m; ang[1] *= s->m; ang[2] *= s->m;
   m3_mulv(s->I, w, Iw);
   ang[0] += Iw[0]; ang[1] += Iw[1]; ang[2] += Iw[2];
   v3_cross(w, s->c, lin);
   lin[0] += v[0]; lin[1] += v[1]; lin[2] += v[2];
   lin[0] *= s->m; lin[1] *= s->m; lin[2] *= s->m;
   memcpy(h6, ang, sizeof ang);
   memcpy(h6 + 3, lin, sizeof lin);
}

/* frame: MO_FRAME_WORLD = the inertial (root) frame; MO_FRAME_COM = axes of the inertial frame, origin at the centre of mass of the
 * whole system (what a CenterOfMassReferenceFrame handed to setCentroidalMomentumFrame() is, :380-387).
 * cmm (nullable) [6][nv] row-major: getCentroidalMomentumMatrix() :801-809; com4 (nullable): CoM in the root frame, total mass. */
static void crba_impl(const mo_tree *t, const double *q, double *M, int frame, double *cmm, double *com4)
{
   frames_t *F = (frames_t *)malloc(sizeof(frames_t));
   si_t *comp = (si_t *)malloc(sizeof(si_t) * (size_t)t->nb);
   xf_t *X = (xf_t *)malloc(sizeof(xf_t) * (size_t)t->nb);
   int nv = t->nv;

   update_frames(t, q, NULL, F);
   memset(M, 0, sizeof(double) * (size_t)nv * (size_t)nv); /* :296 */

   for (int i = 0; i < t->nb; i++)
      if (t->parent[i] >= 0)
         xf_rel(&F->after[i], &F->after[t->parent[i]], &X[i]); /* :592-593 */
      else
         xf_identity(&X[i]);

   for (int i = t->nb - 1; i >= 0; i--)
   {
      body_inertia_at_after(t, i, F, &comp[i]); /* :648-651 */
      for (int c = i + 1; c < t->nb; c++)
         if (t->parent[c] == i)
         {
            si_t ch = comp[c]; /* :653-661 */
            si_apply(&X[c], &ch);
            si_add(&comp[i], &ch);
         }

      int nd = joint_ndof(t, i);
      double S[36], F2[36];
      for (int k = 0; k < nd; k++)
      {
         joint_S_col(t, i, k, S + 6 * k);
         si_momentum(&comp[i], S + 6 * k, F2 + 6 * k); /* :663-667 */
         if (cmm) /* computeCentroidalMomentumMatrix() :801-809: F2.changeFrame(centroidalMomentumFrame), here to the root frame first */
         {
            sv_t f;
            memcpy(f.w, F2 + 6 * k, 3 * sizeof(double));
            memcpy(f.v, F2 + 6 * k + 3, 3 * sizeof(double));
            force_apply(&F->after[i], &f);
            for (int r = 0; r < 3; r++)
            {
               cmm[r * nv + t->dof_off[i] + k] = f.w[r];
               cmm[(3 + r) * nv + t->dof_off[i] + k] = f.v[r];
            }
         }
      }
      for (int a = 0; a < nd; a++) /* :700-707 */
         for (int b = 0; b < nd; b++)
         {
            double d = 0.0;
            for (int r = 0; r < 6; r++) d += S[6 * a + r] * F2[6 * b + r];
            M[(t->dof_off[i] + a) * nv + t->dof_off[i] + b] = d;
            M[(t->dof_off[i] + b) * nv + t->dof_off[i] + a] = d;
         }
      /* :772-797 walk the ancestors */
      int prev = i, anc = t->parent[i];
      while (anc >= 0)
      {
         int nda = joint_ndof(t, anc);
         for (int j = 0; j < nd; j++)
         {
            sv_t f;
            memcpy(f.w, F2 + 6 * j, 3 * sizeof(double));
            memcpy(f.v, F2 + 6 * j + 3, 3 * sizeof(double));
            force_apply(&X[prev], &f);
            memcpy(F2 + 6 * j, f.w, 3 * sizeof(double));
            memcpy(F2 + 6 * j + 3, f.v, 3 * sizeof(double));
            for (int a = 0; a < nda; a++)
            {
               double col[6], d = 0.0;
               joint_S_col(t, anc, a, col);
               for (int r = 0; r < 6; r++) d += col[r] * F2[6 * j + r];
               M[(t->dof_off[anc] + a) * nv + t->dof_off[i] + j] = d;
               M[(t->dof_off[i] + j) * nv + t->dof_off[anc] + a] = d;
            }
         }
         prev = anc;
         anc = t->parent[anc];
      }
   }
   if (cmm || com4)
   {
      /* centre of mass of the system = CoM of the composite inertias of the root's children, in the root frame */
      double c[3] = {0, 0, 0}, m = 0.0;
      for (int i = 0; i < t->nb; i++)
         if (t->parent[i] < 0)
         {
            si_t w = comp[i];
            si_apply(&F->after[i], &w);
            for (int k = 0; k < 3; k++) c[k] += w.m * w.c[k];
            m += w.m;
         }
      for (int k = 0; k < 3; k++) c[k] /= m;
      if (com4) { memcpy(com4, c, sizeof c); com4[3] = m; }
      if (cmm && frame == MO_FRAME_COM)
         for (int j = 0; j < nv; j++)
         {
            /* same axes, origin moved to c: n' = n - c x f */
            double f[3] = {cmm[3 * nv + j], cmm[4 * nv + j], cmm[5 * nv + j]}, cxf[3];
            v3_cross(c, f, cxf);
            for (int r = 0; r < 3; r++) cmm[r * nv + j] -= cxf[r];
         }
   }
   free(F); free(comp); free(X);
}

void mo_crba(const mo_tree *t, const double *q, double *M)
{
   crba_impl(t, q, M, MO_FRAME_WORLD, NULL, NULL);
}

void mo_crba_centroidal(const mo_tree *t, const double *q, int frame, double *M, double *cmm, double *com4)
{
   crba_impl(t, q, M, frame, cmm, com4);
}

/* getCentroidalConvectiveTerm() (CompositeRigidBodyMassMatrixCalculator.java:811-839): the accelerations of pass one of inverse
 * dynamics with zero joint accelerations and no gravity (coriolisBodyAcceleration, :826-829 = InverseDynamicsCalculator.java:880-892),
 * each body's net wrench (:831) re-expressed in the centroidal frame and summed (:832-833). */
void mo_centroidal_convective_term(const mo_tree *t, const double *q, const double *qd, int frame, double *out6)
{
   const double g0[3] = {0.0, 0.0, 0.0};
   sv_t *acc = (sv_t *)malloc(sizeof(sv_t) * (size_t)t->nb);
   frames_t *F = (frames_t *)malloc(sizeof(frames_t));
   double *zero = (double *)calloc((size_t)t->nv, sizeof(double));
   sv_t sum;
   memset(&sum, 0, sizeof sum);
   rnea_impl(t, g0, q, qd, zero, NULL, MO_NO_ACCELERATIONS, NULL, (double *)acc, NULL);
   update_frames(t, q, qd, F);
   for (int i = 0; i < t->nb; i++)
   {
      sv_t W;
      dynamic_wrench(t->J + 9 * i, t->mass[i], &acc[i], &F->tw_com[i], &W);
      force_apply(&F->com[i], &W);
      sv_add(&sum, &W);
   }
   if (frame == MO_FRAME_COM)
   {
      double com4[4], cxf[3];
      double *M = (double *)malloc(sizeof(double) * (size_t)t->nv * (size_t)t->nv);
      crba_impl(t, q, M, MO_FRAME_WORLD, NULL, com4);
      v3_cross(com4, sum.v, cxf);
      for (int r = 0; r < 3; r++) sum.w[r] -= cxf[r];
      free(M);
   }
   memcpy(out6, sum.w, 3 * sizeof(double));
   memcpy(out6 + 3, sum.v, 3 * sizeof(double));
   free(acc); free(F); free(zero);
}

/* ================================================================== Coriolis and centrifugal matrix
 * CompositeRigidBodyMassMatrixCalculator with setEnableCoriolisMatrixCalculation(true) (:278-281, getCoriolisMatrix :358-366):
 * computeMassMatrix() :588-799 with the factorized body inertia of M/algorithms/FactorizedBodyInertia.java (:149-175 construction
 * from a spatial inertia and the body twist, :295-312 applyTransform, :201-247 transform / addTransform / transposeTransform). */

typedef struct { double A[9], L[9], TR[9], BL[9]; } fbi_t; /* angular, linear, top-right, bottom-left 3x3 blocks */

static void m3_tilde(const double *p, double *T)
{
   T[0] = 0; T[1] = -p[2]; T[2] = p[1];
   T[3] = p[2]; T[4] = 0; T[5] = -p[0];
   T[6] = -p[1]; T[7] = p[0]; T[8] = 0;
}

/* FactorizedBodyInertia.tildeTimesTilde :379-391 */
static void m3_tilde_times_tilde(const double *p1, const double *p2, double *C)
{
   C[0] = -p1[2] * p2[2] - p1[1] * p2[1]; C[1] = p1[1] * p2[0]; C[2] = p1[2] * p2[0];
   C[3] = p1[0] * p2[1]; C[4] = -p1[2] * p2[2] - p1[0] * p2[0]; C[5] = p1[2] * p2[1];
   C[6] = p1[0] * p2[2]; C[7] = p1[1] * p2[2]; C[8] = -p1[1] * p2[1] - p1[0] * p2[0];
}

/* setIncludingFrame(spatialInertia, bodyTwist) :149-175 */
static void fbi_from_si(const si_t *s, const sv_t *tw, fbi_t *b)
{
   double T[9], WJ[9];
   m3_tilde_times_tilde(tw->v, s->c, b->A); /* w x J - m v x c x */
   for (int k = 0; k < 9; k++) b->A[k] *= -s->m;
   m3_tilde(tw->w, T);
   m3_mul(T, s->I, WJ);
   for (int k = 0; k < 9; k++) b->A[k] += WJ[k];
   m3_tilde_times_tilde(tw->w, s->c, b->BL); /* -m w x c x */
   for (int k = 0; k < 9; k++) b->BL[k] *= -s->m;
   m3_tilde(tw->v, b->TR); /* m v x + m w x c x */
   for (int k = 0; k < 9; k++) b->TR[k] = b->TR[k] * s->m - b->BL[k];
   m3_tilde(tw->w, b->L); /* m w x */
   for (int k = 0; k < 9; k++) b->L[k] *= s->m;
}

/* applyTransform(RigidBodyTransform) :295-312 */
static void fbi_apply(const xf_t *x, fbi_t *b)
{
   double T[9], P[9];
   m3_rot_congruence(x->R, b->A);
   m3_rot_congruence(x->R, b->L);
   m3_rot_congruence(x->R, b->TR);
   m3_rot_congruence(x->R, b->BL);
   m3_tilde(x->t, T);
   m3_mul(T, b->L, P);  for (int k = 0; k < 9; k++) b->TR[k] += P[k]; /* addTildeTimesMatrix(t, linear, topRight) */
   m3_mul(T, b->BL, P); for (int k = 0; k < 9; k++) b->A[k] += P[k];  /* addTildeTimesMatrix(t, bottomLeft, angular) */
   m3_mul(b->TR, T, P); for (int k = 0; k < 9; k++) b->A[k] -= P[k];  /* subMatrixTimesTilde(topRight, t, angular) */
   m3_mul(b->L, T, P);  for (int k = 0; k < 9; k++) b->BL[k] -= P[k]; /* subMatrixTimesTilde(linear, t, bottomLeft) */
}

static void fbi_mulv(const fbi_t *b, const double *x6, double *y6, int add) /* transform / addTransform :201-228 */
{
   double a[3], c[3];
   m3_mulv(b->A, x6, a);      m3_mulv(b->TR, x6 + 3, c);
   for (int k = 0; k < 3; k++) y6[k] = (add ? y6[k] : 0.0) + a[k] + c[k];
   m3_mulv(b->BL, x6, a);     m3_mulv(b->L, x6 + 3, c);
   for (int k = 0; k < 3; k++) y6[3 + k] = (add ? y6[3 + k] : 0.0) + a[k] + c[k];
}

static void fbi_tmulv(const fbi_t *b, const double *x6, double *y6) /* transposeTransform :230-247 */
{
   double a[3], c[3];
   m3_tmulv(b->A, x6, a);     m3_tmulv(b->BL, x6 + 3, c);
   for (int k = 0; k < 3; k++) y6[k] = a[k] + c[k];
   m3_tmulv(b->TR, x6, a);    m3_tmulv(b->L, x6 + 3, c);
   for (int k = 0; k < 3; k++) y6[3 + k] = a[k] + c[k];
}

static double dot6(const double *a, const double *b)
{
   return a[0] * b[0] + a[1] * b[1] + a[2] * b[2] + a[3] * b[3] + a[4] * b[4] + a[5] * b[5];
}

static void force6_apply(const xf_t *x, double *f6)
{
   sv_t f;
   memcpy(f.w, f6, 3 * sizeof(double));
   memcpy(f.v, f6 + 3, 3 * sizeof(double));
   force_apply(x, &f);
   memcpy(f6, f.w, 3 * sizeof(double));
   memcpy(f6 + 3, f.v, 3 * sizeof(double));
}

void mo_coriolis(const mo_tree *t, const double *q, const double *qd, double *M, double *C)
{
   frames_t *F = (frames_t *)malloc(sizeof(frames_t));
   si_t *comp = (si_t *)malloc(sizeof(si_t) * (size_t)t->nb);
   fbi_t *fcomp = (fbi_t *)malloc(sizeof(fbi_t) * (size_t)t->nb);
   xf_t *X = (xf_t *)malloc(sizeof(xf_t) * (size_t)t->nb);
   double (*Sd)[36] = (double (*)[36])malloc(sizeof(double[36]) * (size_t)t->nb); /* unitTwistDots, per body */
   int nv = t->nv;

   update_frames(t, q, qd, F);
   memset(C, 0, sizeof(double) * (size_t)nv * (size_t)nv); /* :298-299 */
   if (M) memset(M, 0, sizeof(double) * (size_t)nv * (size_t)nv);

   for (int i = 0; i < t->nb; i++)
   {
      if (t->parent[i] >= 0)
         xf_rel(&F->after[i], &F->after[t->parent[i]], &X[i]);
      else
         xf_identity(&X[i]);
      /* unitTwistDots :604-630 (constant motion subspace): [w x s_w ; v x s_w + w x s_v] with the twist of frameAfterJoint */
      const sv_t *tw = &F->tw_after[i];
      for (int k = 0; k < joint_ndof(t, i); k++)
      {
         double s[6], *d = Sd[i] + 6 * k;
         joint_S_col(t, i, k, s);
         v3_cross(tw->w, s, d);
         v3_cross(tw->v, s, d + 3);
         v3_add_cross(tw->w, s + 3, d + 3);
      }
   }

   for (int i = t->nb - 1; i >= 0; i--)
   {
      si_t body;
      body_inertia_at_after(t, i, F, &body); /* :648-651 */
      comp[i] = body;
      fbi_from_si(&body, &F->tw_after[i], &fcomp[i]); /* :671-673 */
      for (int c = i + 1; c < t->nb; c++)
         if (t->parent[c] == i)
         {
            si_t ch = comp[c]; /* :653-661 */
            fbi_t fch = fcomp[c]; /* :675-683 */
            si_apply(&X[c], &ch);
            si_add(&comp[i], &ch);
            fbi_apply(&X[c], &fch);
            for (int k = 0; k < 9; k++)
            {
               fcomp[i].A[k] += fch.A[k]; fcomp[i].L[k] += fch.L[k]; fcomp[i].TR[k] += fch.TR[k]; fcomp[i].BL[k] += fch.BL[k];
            }
         }

      int nd = joint_ndof(t, i), di = t->dof_off[i];
      double S[36], F1[36], F2[36], F3[36];
      for (int k = 0; k < nd; k++)
      {
         joint_S_col(t, i, k, S + 6 * k);
         si_momentum(&comp[i], S + 6 * k, F2 + 6 * k);      /* :663-667 */
         si_momentum(&comp[i], Sd[i] + 6 * k, F1 + 6 * k);  /* :687-688 */
         fbi_mulv(&fcomp[i], S + 6 * k, F1 + 6 * k, 1);     /* :689 */
         fbi_tmulv(&fcomp[i], S + 6 * k, F3 + 6 * k);       /* :692 */
      }
      for (int a = 0; a < nd; a++) /* :700-725, in the order of the Java loops (later writes win) */
         for (int b = 0; b < nd; b++)
         {
            if (M)
            {
               double m = dot6(S + 6 * a, F2 + 6 * b);
               M[(di + a) * nv + di + b] = m;
               M[(di + b) * nv + di + a] = m;
            }
            C[(di + a) * nv + di + b] = dot6(S + 6 * a, F1 + 6 * b);
            if (a != b)
               C[(di + b) * nv + di + a] = dot6(Sd[i] + 6 * a, F2 + 6 * b) + dot6(S + 6 * a, F3 + 6 * b);
         }
      /* :730-766 */
      int prev = i, anc = t->parent[i];
      while (anc >= 0)
      {
         int nda = joint_ndof(t, anc), da = t->dof_off[anc];
         for (int j = 0; j < nd; j++)
         {
            force6_apply(&X[prev], F1 + 6 * j);
            force6_apply(&X[prev], F2 + 6 * j);
            force6_apply(&X[prev], F3 + 6 * j);
            for (int a = 0; a < nda; a++)
            {
               double col[6];
               joint_S_col(t, anc, a, col);
               if (M)
               {
                  double m = dot6(col, F2 + 6 * j);
                  M[(da + a) * nv + di + j] = m;
                  M[(di + j) * nv + da + a] = m;
               }
               C[(da + a) * nv + di + j] = dot6(col, F1 + 6 * j);
               C[(di + j) * nv + da + a] = dot6(Sd[anc] + 6 * a, F2 + 6 * j) + dot6(col, F3 + 6 * j);
            }
         }
         prev = anc;
         anc = t->parent[anc];
      }
   }
   free(F); free(comp); free(fcomp); free(X); free(Sd);
}

/* ================================================================== batched drivers (CPU baseline: one
 * calculator instance per thread, like one cloned MultiBodySystem + calculator per thread in Java,
 * M/tools/MultiBodySystemFactories.java:310).  Plain pthreads, static contiguous slices. */

/* ------------------------------------------------------------------------------------------------
 * State integrator: MultiBodySystemStateIntegrator.doubleIntegrateFromAcceleration
 * (tools/MultiBodySystemStateIntegrator.java:365-470 dispatch, :503-560 floating joints, :710-733 one-DoF joints).
 * q, qd and -- for SixDoF joints -- qdd are updated in place, exactly as the Java code updates the joint objects.
 * Third-party arithmetic restated from its published definition (Euclid 0.21.0, not vendored):
 *   Quaternion.setRotationVector (RotationVectorConversion -> QuaternionConversion): q = [r/|r| sin(|r|/2), cos(|r|/2)],
 *   identity below |r| = 1e-12;  Quaternion.append = Hamilton product q0 * qi;  Quaternion.transform /
 *   inverseTransform = rotation by the (normalised) quaternion and by its conjugate.
 * ------------------------------------------------------------------------------------------------ */
static void quat_mul(const double *a, const double *b, double *c) /* (x y z s) Hamilton product a * b */
{
   const double ax = a[0], ay = a[1], az = a[2], as = a[3], bx = b[0], by = b[1], bz = b[2], bs = b[3];
   c[0] = as * bx + ax * bs + ay * bz - az * by;
   c[1] = as * by - ax * bz + ay * bs + az * bx;
   c[2] = as * bz + ax * by - ay * bx + az * bs;
   c[3] = as * bs - ax * bx - ay * by - az * bz;
}

static void quat_from_rotation_vector(const double *r, double *q4)
{
   const double n = sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
   if (n < 1.0e-12)
   {
      q4[0] = q4[1] = q4[2] = 0.0;
      q4[3] = 1.0;
      return;
   }
   const double sh = sin(0.5 * n) / n;
   q4[0] = r[0] * sh;
   q4[1] = r[1] * sh;
   q4[2] = r[2] * sh;
   q4[3] = cos(0.5 * n);
}

void mo_integrate(const mo_tree *t, double dt, double *q, double *qd, double *qdd)
{
   const double half_dt_dt = 0.5 * dt * dt;
   for (int i = 0; i < t->nb; i++)
   {
      double *qi = q + t->cfg_off[i], *vi = qd + t->dof_off[i], *ai = qdd + t->dof_off[i];
      if (t->jtype[i] == MO_SPHERICAL)
      {
         /* :449-452, :575-594 doubleIntegrate(angularAcceleration, angularVelocity, orientation):
          * orientation = orientation * exp(dt w + 0.5 dt^2 wd) ; w += dt wd */
         double rv[3], qint[4], qfin[4];
         for (int k = 0; k < 3; k++)
            rv[k] = dt * vi[k] + half_dt_dt * ai[k];
         quat_from_rotation_vector(rv, qint);
         quat_mul(qi, qint, qfin);
         for (int k = 0; k < 4; k++)
            qi[k] = qfin[k];
         for (int k = 0; k < 3; k++)
            vi[k] = dt * ai[k] + vi[k];
         continue;
      }
      if (t->jtype[i] == MO_PLANAR)
      {
         /* a PlanarJoint is a FloatingJointBasics (PlanarJointBasics.java:17): :421-424 runs the floating-joint update :503-560 on its
          * planar pose / twist / acceleration.  With everything in the x-z plane the rotation vector is along y, so the pitch-only
          * orientation composes additively; the remaining steps are those of the SixDoF branch below restricted to the plane. */
         const double th0 = qi[0], wy = vi[0], vx = vi[1], vz = vi[2], wdy = ai[0];
         /* origin acceleration a + w x v with w = (0, wy, 0), v = (vx, 0, vz): (wy vz, 0, -wy vx) */
         const double lx = ai[1] + wy * vz, lz = ai[2] - wy * vx;
         const double dth = dt * wy + half_dt_dt * wdy;
         const double c0 = cos(th0), s0 = sin(th0), ci = cos(dth), si = sin(dth);
         const double tx = dt * vx + half_dt_dt * lx, tz = dt * vz + half_dt_dt * lz;
         /* R_y(th) (x, 0, z) = (c x + s z, 0, -s x + c z) */
         qi[1] += c0 * tx + s0 * tz;
         qi[2] += -s0 * tx + c0 * tz;
         qi[0] = th0 + dth;
         const double ux = dt * lx + vx, uz = dt * lz + vz;
         /* R_y(dth)^T (x, 0, z) = (c x - s z, 0, s x + c z) */
         const double vfx = ci * ux - si * uz, vfz = si * ux + ci * uz;
         const double wf = dt * wdy + wy;
         const double l2x = ci * lx - si * lz, l2z = si * lx + ci * lz;
         vi[0] = wf; vi[1] = vfx; vi[2] = vfz;
         /* linear = a_origin' + v' x w' with v' = (vfx, 0, vfz), w' = (0, wf, 0): (-vfz wf, 0, vfx wf) */
         ai[1] = l2x - vfz * wf;
         ai[2] = l2z + vfx * wf;
         continue;
      }
      if (t->jtype[i] != MO_SIXDOF)
      {
         /* :710-733  q += 0.5 dt^2 qdd + dt qd ; qd += dt qdd */
         const double q0 = qi[0], v0 = vi[0], a0 = ai[0];
         qi[0] = half_dt_dt * a0 + dt * v0 + q0;
         vi[0] = dt * a0 + v0;
         continue;
      }
      /* :503-560, doubleIntegrate(spatialAcceleration, initialTwist, initialPose, finalTwist, finalPose) */
      const double *w0 = vi, *v0 = vi + 3, *wd = ai;
      double lin[3], R0[9], rv[3], qint[4], Ri[9], dp[3], tmp[3], vfin[3], wfin[3], qfin[4];
      /* linear acceleration of the body origin (SpatialAccelerationReadOnly.java:197-204): a + w x v */
      lin[0] = ai[3]; lin[1] = ai[4]; lin[2] = ai[5];
      v3_add_cross(w0, v0, lin);
      for (int k = 0; k < 3; k++)
         rv[k] = dt * w0[k] + half_dt_dt * wd[k];
      quat_from_rotation_vector(rv, qint);
      for (int k = 0; k < 3; k++)
         wfin[k] = dt * wd[k] + w0[k];
      /* position: p += R(q0) (dt v + 0.5 dt^2 a_origin) */
      rot_quaternion(qi, R0);
      for (int k = 0; k < 3; k++)
         tmp[k] = dt * v0[k] + half_dt_dt * lin[k];
      m3_mulv(R0, tmp, dp);
      /* linear velocity: (v + dt a_origin) re-expressed in the new body frame */
      rot_quaternion(qint, Ri);
      for (int k = 0; k < 3; k++)
         tmp[k] = dt * lin[k] + v0[k];
      m3_tmulv(Ri, tmp, vfin);
      /* orientation: q0 * q_integrated */
      quat_mul(qi, qint, qfin);
      /* acceleration: origin acceleration in the new frame, back to spatial form with the final twist
       * (FixedFrameSpatialAccelerationBasics.java:81-90): linear = a_origin' + v' x w' */
      double lin2[3];
      m3_tmulv(Ri, lin, lin2);
      v3_add_cross(vfin, wfin, lin2);
      for (int k = 0; k < 4; k++)
         qi[k] = qfin[k];
      for (int k = 0; k < 3; k++)
      {
         qi[4 + k] += dp[k];
         vi[k] = wfin[k];
         vi[3 + k] = vfin[k];
         ai[3 + k] = lin2[k];
      }
   }
}

void mo_integrate_batch(const mo_tree *t, double dt, long n, long ld, double *q, double *qd, double *qdd)
{
   double qs[512], vs[512], as[512];
   for (long s = 0; s < n; s++)
   {
      for (int k = 0; k < t->nq; k++) qs[k] = q[k * ld + s];
      for (int k = 0; k < t->nv; k++) { vs[k] = qd[k * ld + s]; as[k] = qdd[k * ld + s]; }
      mo_integrate(t, dt, qs, vs, as);
      for (int k = 0; k < t->nq; k++) q[k * ld + s] = qs[k];
      for (int k = 0; k < t->nv; k++) { qd[k * ld + s] = vs[k]; qdd[k * ld + s] = as[k]; }
   }
}

int mo_max_threads(void)
{
   long n = sysconf(_SC_NPROCESSORS_ONLN);
   return n < 1 ? 1 : (int)n;
}

typedef struct
{
   int kind; /* 0 rnea, 1 aba, 2 crba */
   const mo_tree *t;
   const double *g;
   long s0, s1, ld;
   const double *q, *qd, *in3, *fext;
   double *out;
   int flags;
} job_t;

static void *job_run(void *arg)
{
   job_t *j = (job_t *)arg;
   const mo_tree *t = j->t;
   long ld = j->ld;
   double *b = (double *)malloc(sizeof(double) * (size_t)(t->nq + 3 * t->nv + 6 * t->nb + t->nv * t->nv));
   double *qs = b, *qds = qs + t->nq, *ins = qds + t->nv, *outs = ins + t->nv, *fs = outs + t->nv, *Ms = fs + 6 * t->nb;
   for (long s = j->s0; s < j->s1; s++)
   {
      for (int k = 0; k < t->nq; k++) qs[k] = j->q[k * ld + s];
      if (j->kind == 2)
      {
         mo_crba(t, qs, Ms);
         for (long k = 0; k < (long)t->nv * t->nv; k++) j->out[k * ld + s] = Ms[k];
         continue;
      }
      for (int k = 0; k < t->nv; k++) { qds[k] = j->qd[k * ld + s]; ins[k] = j->in3[k * ld + s]; }
      if (j->fext)
         for (int k = 0; k < 6 * t->nb; k++) fs[k] = j->fext[k * ld + s];
      if (j->kind == 0)
         mo_rnea(t, j->g, qs, qds, ins, j->fext ? fs : NULL, j->flags, outs);
      else
         mo_aba(t, j->g, qs, qds, ins, j->fext ? fs : NULL, outs);
      for (int k = 0; k < t->nv; k++) j->out[k * ld + s] = outs[k];
   }
   free(b);
   return NULL;
}

static void run_jobs(job_t proto, long n, int nthreads)
{
   int mx = mo_max_threads();
   int nt = (nthreads <= 0 || nthreads > mx) ? mx : nthreads;
   if ((long)nt > n) nt = n > 0 ? (int)n : 1;
   pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)nt);
   job_t *jobs = (job_t *)malloc(sizeof(job_t) * (size_t)nt);
   for (int i = 0; i < nt; i++)
   {
      jobs[i] = proto;
      jobs[i].s0 = n * i / nt;
      jobs[i].s1 = n * (i + 1) / nt;
      if (i > 0) pthread_create(&th[i], NULL, job_run, &jobs[i]);
   }
   job_run(&jobs[0]);
   for (int i = 1; i < nt; i++) pthread_join(th[i], NULL);
   free(th); free(jobs);
}

void mo_rnea_batch(const mo_tree *t, const double *g, long n, long ld, const double *q, const double *qd, const double *qdd,
                   const double *fext, int flags, double *tau, int nthreads)
{
   job_t j = {0, t, g, 0, 0, ld, q, qd, qdd, fext, tau, flags};
   run_jobs(j, n, nthreads);
}

void mo_aba_batch(const mo_tree *t, const double *g, long n, long ld, const double *q, const double *qd, const double *tau,
                  const double *fext, double *qdd, int nthreads)
{
   job_t j = {1, t, g, 0, 0, ld, q, qd, tau, fext, qdd, 0};
   run_jobs(j, n, nthreads);
}

void mo_crba_batch(const mo_tree *t, long n, long ld, const double *q, double *M, int nthreads)
{
   job_t j = {2, t, NULL, 0, 0, ld, q, NULL, NULL, NULL, M, 0};
   run_jobs(j, n, nthreads);
}
