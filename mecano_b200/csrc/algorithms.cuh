// algorithms.cuh -- one state of RNEA / ABA / CRBA, written once against a small "context" policy so
// that the same source is (a) inlined into the sm_100a kernels (kernels.cu) and (b) compiled for the
// host by the kernel-source emulation harness that the no-GPU tests use (tests/emu).
//
// What each routine restates (M/ = /root/reference/src/main/java/us/ihmc/mecano/):
//   rnea_state : InverseDynamicsCalculator.compute()            M/algorithms/InverseDynamicsCalculator.java:496-501, 873-966
//   aba_state  : ForwardDynamicsCalculator.compute()            M/algorithms/ForwardDynamicsCalculator.java:508-520, 1085-1310
//   crba_state : CompositeRigidBodyMassMatrixCalculator         M/algorithms/CompositeRigidBodyMassMatrixCalculator.java:588-667, 700-707, 772-797
// including the "set state -> updateFramesRecursively()" prologue (RigidBodyBasics.java:104-112,
// MovingReferenceFrame.java:279-311) which Mecano keeps in the frame tree and which is fused here.
//
// Unlike the reference (CoM frames for RNEA, through-the-root frame changes) everything is expressed
// in the canonical joint frames of program.h with local parent<->child transforms; results are
// frame-independent and agree with the oracle to round-off.
//
// Context policy (all methods inline):
//   T   ld_q(row) ld_qd(row) ld_x(row)      inputs in Mecano row order (x = qdd for RNEA, tau for ABA)
//   T   ld_fext(ext_body, comp)             external wrench rows
//   void st_out(row, T)                     tau (RNEA) / qdd (ABA)
//   void st_M(row, col, T)                  mass-matrix entry (CRBA)
//   T   stk_ld(i) / void stk_st(i, T)       per-state stack (shared memory on the GPU)
//   T   aux_ld(i) / aux_st, rec_ld / rec_st per-state branch-save and record areas (local memory)
//   const T* cst(body)                      constant record of a body (shared memory on the GPU)
#pragma once
#include "program.h"
#include "spatial.cuh"
#include "jointmath.cuh"
#include "rnea.cuh"
#include "crba.cuh"

namespace mb
{
// ---- stack helpers
template <class T, class Ctx> MB_HD void stk1_st_sv(Ctx &c, int i, const SvT<T> &v)
{
   c.stk_st(i + 0, v.a.x); c.stk_st(i + 1, v.a.y); c.stk_st(i + 2, v.a.z);
   c.stk_st(i + 3, v.l.x); c.stk_st(i + 4, v.l.y); c.stk_st(i + 5, v.l.z);
}
template <class T, class Ctx> MB_HD SvT<T> stk1_ld_sv(Ctx &c, int i)
{
   SvT<T> v;
   v.a = v3<T>(c.stk_ld(i + 0), c.stk_ld(i + 1), c.stk_ld(i + 2));
   v.l = v3<T>(c.stk_ld(i + 3), c.stk_ld(i + 4), c.stk_ld(i + 5));
   return v;
}
// Joint parameters: what must be kept to rebuild the joint transform on the way back up.
template <class T> struct JpT
{
   T s, c;      // revolute: sin/cos; prismatic: s = q
   XfT<T> X;    // SixDoF: the whole transform
};

// (a1) joint transform X_J(q) composed with the fixed offset (MecanoFactories.java:231-260,
// PrismaticJointReadOnly.java:18-22, FloatingJointReadOnly.java:34-37), in canonical frames
template <class T, class Ctx> MB_HD XfT<T> joint_transform(Ctx &c, const MbBody &B, const T *C, JpT<T> &jp, T q1)
{
   XfT<T> X;
   const M3T<T> R0 = ld_m3(C + MB_C_R);
   const V3T<T> p0 = ld_v3(C + MB_C_P);
   if (B.jtype == MB_REVOLUTE)
   {
      mb_sincos(q1, &jp.s, &jp.c);
      X.R = mul_rz(R0, jp.s, jp.c);
      X.p = p0;
   }
   else if (B.jtype == MB_PRISMATIC)
   {
      jp.s = q1;
      X.R = R0;
      X.p = p0 + jp.s * v3<T>(R0.xz, R0.yz, R0.zz);
   }
   else
   {
      const int r = B.cfg_off;
      const M3T<T> Rq = quat_to_rot(c.ld_q(r), c.ld_q(r + 1), c.ld_q(r + 2), c.ld_q(r + 3));
      X.R = mul(R0, Rq);
      X.p = p0 + mul(R0, v3<T>(c.ld_q(r + 4), c.ld_q(r + 5), c.ld_q(r + 6)));
      jp.X = X;
   }
   return X;
}

template <class T> MB_HD XfT<T> rebuild_transform(int jtype, const T *C, const JpT<T> &jp)
{
   if (jtype == MB_SIXDOF)
      return jp.X;
   XfT<T> X;
   const M3T<T> R0 = ld_m3(C + MB_C_R);
   const V3T<T> p0 = ld_v3(C + MB_C_P);
   if (jtype == MB_REVOLUTE)
   {
      X.R = mul_rz(R0, jp.s, jp.c);
      X.p = p0;
   }
   else
   {
      X.R = R0;
      X.p = p0 + jp.s * v3<T>(R0.xz, R0.yz, R0.zz);
   }
   return X;
}

template <class T, class Ctx> MB_HD void stk_st_jp(Ctx &c, int i, int jtype, const JpT<T> &jp)
{
   if (jtype == MB_REVOLUTE)
   {
      c.stk_st(i, jp.s);
      c.stk_st(i + 1, jp.c);
   }
   else if (jtype == MB_PRISMATIC)
      c.stk_st(i, jp.s);
   else
   {
      const M3T<T> &R = jp.X.R;
      c.stk_st(i + 0, R.xx); c.stk_st(i + 1, R.xy); c.stk_st(i + 2, R.xz);
      c.stk_st(i + 3, R.yx); c.stk_st(i + 4, R.yy); c.stk_st(i + 5, R.yz);
      c.stk_st(i + 6, R.zx); c.stk_st(i + 7, R.zy); c.stk_st(i + 8, R.zz);
      c.stk_st(i + 9, jp.X.p.x); c.stk_st(i + 10, jp.X.p.y); c.stk_st(i + 11, jp.X.p.z);
   }
}
template <class T, class Ctx> MB_HD void stk_ld_jp(Ctx &c, int i, int jtype, JpT<T> &jp)
{
   if (jtype == MB_REVOLUTE)
   {
      jp.s = c.stk_ld(i);
      jp.c = c.stk_ld(i + 1);
   }
   else if (jtype == MB_PRISMATIC)
      jp.s = c.stk_ld(i);
   else
   {
      M3T<T> &R = jp.X.R;
      R.xx = c.stk_ld(i + 0); R.xy = c.stk_ld(i + 1); R.xz = c.stk_ld(i + 2);
      R.yx = c.stk_ld(i + 3); R.yy = c.stk_ld(i + 4); R.yz = c.stk_ld(i + 5);
      R.zx = c.stk_ld(i + 6); R.zy = c.stk_ld(i + 7); R.zz = c.stk_ld(i + 8);
      jp.X.p = v3<T>(c.stk_ld(i + 9), c.stk_ld(i + 10), c.stk_ld(i + 11));
   }
}

// S * x for the joint (motion subspace in canonical frames: revolute [e_z;0], prismatic [0;e_z], SixDoF 1_6;
// JointReadOnly.java:201-207, MecanoTools.java:964-995)
template <class T, class F> MB_HD SvT<T> joint_motion(int jtype, int row, F ld)
{
   SvT<T> r = sv_zero<T>();
   if (jtype == MB_REVOLUTE)
      r.a.z = ld(row);
   else if (jtype == MB_PRISMATIC)
      r.l.z = ld(row);
   else
   {
      r.a = v3<T>(ld(row), ld(row + 1), ld(row + 2));
      r.l = v3<T>(ld(row + 3), ld(row + 4), ld(row + 5));
   }
   return r;
}

// S * x with the scalar of a 1-DoF joint already in a register (SixDoF joints load their six rows)
template <class T, class F> MB_HD SvT<T> joint_motion_pf(int jtype, int row, T x1, F ld)
{
   SvT<T> r = sv_zero<T>();
   if (jtype == MB_REVOLUTE)
      r.a.z = x1;
   else if (jtype == MB_PRISMATIC)
      r.l.z = x1;
   else
   {
      r.a = v3<T>(ld(row), ld(row + 1), ld(row + 2));
      r.l = v3<T>(ld(row + 3), ld(row + 4), ld(row + 5));
   }
   return r;
}

// Software prefetch: while op k runs, the global loads of op k+1 (the scalars of a 1-DoF joint) are already in
// flight, so their HBM latency overlaps one whole body of arithmetic instead of stalling the warp.
template <class T> struct PfT
{
   T q, qd, x;
};

// external wrench on a body, given in its CoM frame, re-expressed in the canonical joint frame
template <class T, class Ctx> MB_HD SvT<T> external_wrench_b(Ctx &c, const MbBody &B, const T *C)
{
   SvT<T> w, r;
   const int e = B.ext_index;
   w.a = v3<T>(c.ld_fext(e, 0), c.ld_fext(e, 1), c.ld_fext(e, 2));
   w.l = v3<T>(c.ld_fext(e, 3), c.ld_fext(e, 4), c.ld_fext(e, 5));
   const M3T<T> E = ld_m3(C + MB_C_E);
   r.l = mul(E, w.l);
   r.a = mul(E, w.a) + cross(ld_v3(C + MB_C_C), r.l);
   return r;
}

// ======================================================================================== ABA
template <class T, class Ctx> MB_HD void aux_st_abi(Ctx &c, int i, const AbiT<T> &I, const SvT<T> &p)
{
   c.aux_st(i + 0, I.A.xx); c.aux_st(i + 1, I.A.xy); c.aux_st(i + 2, I.A.xz); c.aux_st(i + 3, I.A.yy); c.aux_st(i + 4, I.A.yz); c.aux_st(i + 5, I.A.zz);
   c.aux_st(i + 6, I.C.xx); c.aux_st(i + 7, I.C.xy); c.aux_st(i + 8, I.C.xz); c.aux_st(i + 9, I.C.yx); c.aux_st(i + 10, I.C.yy); c.aux_st(i + 11, I.C.yz);
   c.aux_st(i + 12, I.C.zx); c.aux_st(i + 13, I.C.zy); c.aux_st(i + 14, I.C.zz);
   c.aux_st(i + 15, I.L.xx); c.aux_st(i + 16, I.L.xy); c.aux_st(i + 17, I.L.xz); c.aux_st(i + 18, I.L.yy); c.aux_st(i + 19, I.L.yz); c.aux_st(i + 20, I.L.zz);
   aux_st_sv<T>(c, i + 21, p);
}
template <class T, class Ctx> MB_HD void aux_ld_abi(Ctx &c, int i, AbiT<T> &I, SvT<T> &p)
{
   I.A.xx = c.aux_ld(i + 0); I.A.xy = c.aux_ld(i + 1); I.A.xz = c.aux_ld(i + 2); I.A.yy = c.aux_ld(i + 3); I.A.yz = c.aux_ld(i + 4); I.A.zz = c.aux_ld(i + 5);
   I.C.xx = c.aux_ld(i + 6); I.C.xy = c.aux_ld(i + 7); I.C.xz = c.aux_ld(i + 8); I.C.yx = c.aux_ld(i + 9); I.C.yy = c.aux_ld(i + 10); I.C.yz = c.aux_ld(i + 11);
   I.C.zx = c.aux_ld(i + 12); I.C.zy = c.aux_ld(i + 13); I.C.zz = c.aux_ld(i + 14);
   I.L.xx = c.aux_ld(i + 15); I.L.xy = c.aux_ld(i + 16); I.L.xz = c.aux_ld(i + 17); I.L.yy = c.aux_ld(i + 18); I.L.yz = c.aux_ld(i + 19); I.L.zz = c.aux_ld(i + 20);
   p = aux_ld_sv<T>(c, i + 21);
}

template <class T, class Ctx, bool FEXT> MB_HD void aba_state(const MbProgram &P, Ctx &c, const T *grav)
{
   SvT<T> v = sv_zero<T>(), vj = sv_zero<T>(), pacc = sv_zero<T>();
   AbiT<T> acc = AbiT<T>();
   XfT<T> X;
   JpT<T> jp;
   X.R = M3T<T>();
   X.p = v3<T>(0, 0, 0);
   const int nops = P.nops;
   PfT<T> pf;
   pf.q = pf.qd = pf.x = (T)0;
   auto prefetch = [&](uint32_t wn) {
      const MbBody &Bn = P.body[MB_OP_BODY(wn)];
      if (Bn.jtype == MB_SIXDOF)
         return;
      if (wn & MB_OP_ASCEND)
         pf.x = c.ld_x(Bn.dof_off); // tau of the joint whose subtree is about to be folded
      else
      {
         pf.q = c.ld_q(Bn.cfg_off);
         pf.qd = c.ld_qd(Bn.dof_off);
      }
   };
   prefetch(P.op[0]);
   // ---- passes one and two interleaved along the depth-first traversal
   for (int k = 0; k < nops; k++)
   {
      const uint32_t w = P.op[k];
      const int i = MB_OP_BODY(w);
      const MbBody &B = P.body[i];
      const T *C = c.cst(i);
      const PfT<T> cur = pf;
      if (k + 1 < nops)
         prefetch(P.op[k + 1]);
      if (!(w & MB_OP_ASCEND))
      {
         // twist of the body (the frame tree's lazy twist-of-frame, MovingReferenceFrame.java:279-311)
         SvT<T> vp;
         if (w & MB_F_ROOT_PARENT)
            vp = sv_zero<T>();
         else if (w & MB_F_LOAD_PARENT)
            vp = stk1_ld_sv<T>(c, P.body[B.parent].slot);
         else
            vp = v;
         X = joint_transform<T>(c, B, C, jp, cur.q);
         vj = joint_motion_pf<T>(B.jtype, B.dof_off, cur.qd, [&](int r) { return c.ld_qd(r); });
         v = motion_to_child(X, vp) + vj;
         if (!(w & MB_F_LEAF))
         {
            stk1_st_sv<T>(c, B.slot, v);
            const int njp = mb_jp_size(B.jtype);
            stk_st_jp<T>(c, B.slot + 6, B.jtype, jp);
            if (B.jtype == MB_REVOLUTE)
               c.stk_st(B.slot + 6 + njp, vj.a.z);
            else if (B.jtype == MB_PRISMATIC)
               c.stk_st(B.slot + 6 + njp, vj.l.z);
            else
               stk1_st_sv<T>(c, B.slot + 6 + njp, vj);
         }
      }
      else
      {
         if (!(w & MB_F_LEAF))
         {
            v = stk1_ld_sv<T>(c, B.slot);
            const int njp = mb_jp_size(B.jtype);
            stk_ld_jp<T>(c, B.slot + 6, B.jtype, jp);
            X = rebuild_transform<T>(B.jtype, C, jp);
            vj = sv_zero<T>();
            if (B.jtype == MB_REVOLUTE)
               vj.a.z = c.stk_ld(B.slot + 6 + njp);
            else if (B.jtype == MB_PRISMATIC)
               vj.l.z = c.stk_ld(B.slot + 6 + njp);
            else
               vj = stk1_ld_sv<T>(c, B.slot + 6 + njp);
         }
         // pass one quantities (ForwardDynamicsCalculator.java:1109-1118): bias wrench and bias acceleration
         const RbiT<T> I = ld_rbi(C);
         SvT<T> pA = cross_force(v, mul(I, v));
         if (FEXT)
            pA = pA - external_wrench_b<T>(c, B, C);
         const SvT<T> cb = cross_motion(v, vj);
         // pass two (:1136-1254)
         AbiT<T> IA = abi_from_rbi(I);
         if (!(w & MB_F_LEAF))
         {
            IA = IA + acc;
            pA = pA + pacc;
         }
         AbiT<T> Ia;
         SvT<T> pa;
         const bool to_parent = !(w & MB_F_ROOT_PARENT);
         if (B.jtype != MB_SIXDOF)
         {
            SvT<T> U;
            T D, u;
            const T tau = cur.x;
            if (B.jtype == MB_REVOLUTE)
            {
               U.a = v3<T>(IA.A.xz, IA.A.yz, IA.A.zz);
               U.l = v3<T>(IA.C.zx, IA.C.zy, IA.C.zz);
               D = IA.A.zz;
               u = tau - pA.a.z;
            }
            else
            {
               U.a = v3<T>(IA.C.xz, IA.C.yz, IA.C.zz);
               U.l = v3<T>(IA.L.xz, IA.L.yz, IA.L.zz);
               D = IA.L.zz;
               u = tau - pA.l.z;
            }
            const T Dinv = (T)1 / D;
            SvT<T> g;
            g.a = Dinv * U.a;
            g.l = Dinv * U.l;
            const T k0 = Dinv * u;
            // record for pass three: qdd = k0 - g . a'
            c.rec_st(B.rec + 0, g.a.x); c.rec_st(B.rec + 1, g.a.y); c.rec_st(B.rec + 2, g.a.z);
            c.rec_st(B.rec + 3, g.l.x); c.rec_st(B.rec + 4, g.l.y); c.rec_st(B.rec + 5, g.l.z);
            c.rec_st(B.rec + 6, k0);
            c.rec_st(B.rec + 7, jp.s);
            if (B.jtype == MB_REVOLUTE)
               c.rec_st(B.rec + 8, jp.c);
            if (to_parent)
            {
               Ia = abi_downdate(IA, U, g);       // I^a = I^A - U D^-1 U^T
               pa = pA + mul(Ia, cb);             // p^a = p^A + I^a c + U D^-1 u
               pa.a = pa.a + k0 * U.a;
               pa.l = pa.l + k0 * U.l;
            }
         }
         else
         {
            SvT<T> tau6 = joint_motion<T>(MB_SIXDOF, B.dof_off, [&](int r) { return c.ld_x(r); });
            // D = I^A, U = I^A: a_i = D^-1 u, and the joint transmits nothing but tau to its parent
            const SvT<T> x = abi_solve(IA, tau6 - pA);
            c.rec_st(B.rec + 0, x.a.x); c.rec_st(B.rec + 1, x.a.y); c.rec_st(B.rec + 2, x.a.z);
            c.rec_st(B.rec + 3, x.l.x); c.rec_st(B.rec + 4, x.l.y); c.rec_st(B.rec + 5, x.l.z);
            c.rec_st(B.rec + 6, X.R.xx); c.rec_st(B.rec + 7, X.R.xy); c.rec_st(B.rec + 8, X.R.xz);
            c.rec_st(B.rec + 9, X.R.yx); c.rec_st(B.rec + 10, X.R.yy); c.rec_st(B.rec + 11, X.R.yz);
            c.rec_st(B.rec + 12, X.R.zx); c.rec_st(B.rec + 13, X.R.zy); c.rec_st(B.rec + 14, X.R.zz);
            c.rec_st(B.rec + 15, X.p.x); c.rec_st(B.rec + 16, X.p.y); c.rec_st(B.rec + 17, X.p.z);
            if (to_parent)
            {
               Ia = AbiT<T>();
               pa = tau6;
            }
         }
         if (to_parent)
         {
            const AbiT<T> K = abi_to_parent(X, Ia); // :1159-1165
            const SvT<T> Pp = force_to_parent(X, pa);
            const int pa_off = P.body[B.parent].aux;
            if (w & MB_F_FIRST_CHILD)
            {
               acc = K;
               pacc = Pp;
            }
            else
            {
               aux_ld_abi<T>(c, pa_off, acc, pacc);
               acc = acc + K;
               pacc = pacc + Pp;
            }
            if (w & MB_F_STORE_ACC)
               aux_st_abi<T>(c, pa_off, acc, pacc);
         }
      }
   }
   // ---- pass three (:1259-1310), root to leaves, in the same depth-first order
   SvT<T> a = sv_zero<T>();
   v = sv_zero<T>();
   auto prefetch3 = [&](int from) {
      for (int kk = from; kk < nops; kk++)
      {
         const uint32_t wn = P.op[kk];
         if (wn & MB_OP_ASCEND)
            continue;
         const MbBody &Bn = P.body[MB_OP_BODY(wn)];
         if (Bn.jtype != MB_SIXDOF)
            pf.qd = c.ld_qd(Bn.dof_off);
         return;
      }
   };
   prefetch3(0);
   for (int k = 0; k < nops; k++)
   {
      const uint32_t w = P.op[k];
      if (w & MB_OP_ASCEND)
         continue;
      const T qd1 = pf.qd;
      prefetch3(k + 1);
      const int i = MB_OP_BODY(w);
      const MbBody &B = P.body[i];
      const T *C = c.cst(i);
      SvT<T> vp, ap;
      if (w & MB_F_ROOT_PARENT)
      {
         vp = sv_zero<T>();
         ap = sv_zero<T>();
         ap.l = v3<T>(-grav[0], -grav[1], -grav[2]);
      }
      else if (w & MB_F_LOAD_PARENT)
      {
         const int pa_off = P.body[B.parent].aux;
         vp = aux_ld_sv<T>(c, pa_off);
         ap = aux_ld_sv<T>(c, pa_off + 6);
      }
      else
      {
         vp = v;
         ap = a;
      }
      vj = joint_motion_pf<T>(B.jtype, B.dof_off, qd1, [&](int r) { return c.ld_qd(r); });
      if (B.jtype != MB_SIXDOF)
      {
         SvT<T> g;
         g.a = v3<T>(c.rec_ld(B.rec + 0), c.rec_ld(B.rec + 1), c.rec_ld(B.rec + 2));
         g.l = v3<T>(c.rec_ld(B.rec + 3), c.rec_ld(B.rec + 4), c.rec_ld(B.rec + 5));
         const T k0 = c.rec_ld(B.rec + 6);
         jp.s = c.rec_ld(B.rec + 7);
         if (B.jtype == MB_REVOLUTE)
            jp.c = c.rec_ld(B.rec + 8);
         X = rebuild_transform<T>(B.jtype, C, jp);
         v = motion_to_child(X, vp) + vj;
         const SvT<T> a1 = motion_to_child(X, ap) + cross_motion(v, vj); // a' = X^-1 a_parent + c
         const T qdd = k0 - (dot(g.a, a1.a) + dot(g.l, a1.l));            // D^-1 (u - U^T a')
         c.st_out(B.dof_off, qdd);
         a = a1;
         if (B.jtype == MB_REVOLUTE)
            a.a.z += qdd;
         else
            a.l.z += qdd;
      }
      else
      {
         SvT<T> x;
         x.a = v3<T>(c.rec_ld(B.rec + 0), c.rec_ld(B.rec + 1), c.rec_ld(B.rec + 2));
         x.l = v3<T>(c.rec_ld(B.rec + 3), c.rec_ld(B.rec + 4), c.rec_ld(B.rec + 5));
         X.R.xx = c.rec_ld(B.rec + 6); X.R.xy = c.rec_ld(B.rec + 7); X.R.xz = c.rec_ld(B.rec + 8);
         X.R.yx = c.rec_ld(B.rec + 9); X.R.yy = c.rec_ld(B.rec + 10); X.R.yz = c.rec_ld(B.rec + 11);
         X.R.zx = c.rec_ld(B.rec + 12); X.R.zy = c.rec_ld(B.rec + 13); X.R.zz = c.rec_ld(B.rec + 14);
         X.p = v3<T>(c.rec_ld(B.rec + 15), c.rec_ld(B.rec + 16), c.rec_ld(B.rec + 17));
         v = motion_to_child(X, vp) + vj;
         const SvT<T> a1 = motion_to_child(X, ap) + cross_motion(v, vj);
         const SvT<T> qdd = x - a1;
         c.st_out(B.dof_off + 0, qdd.a.x); c.st_out(B.dof_off + 1, qdd.a.y); c.st_out(B.dof_off + 2, qdd.a.z);
         c.st_out(B.dof_off + 3, qdd.l.x); c.st_out(B.dof_off + 4, qdd.l.y); c.st_out(B.dof_off + 5, qdd.l.z);
         a = x;
      }
      if (w & MB_F_SAVE_STATE)
      {
         aux_st_sv<T>(c, B.aux, v);
         aux_st_sv<T>(c, B.aux + 6, a);
      }
   }
}

} // namespace mb
