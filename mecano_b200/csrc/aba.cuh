// aba.cuh -- one state of the articulated-body algorithm (ForwardDynamicsCalculator.compute(),
// M/algorithms/ForwardDynamicsCalculator.java:508-520; passOne :1085-1127, passTwo :1136-1254, passThree :1259-1310;
// ArticulatedBodyInertia.java:176-186, :359-375), with the frame update fused in like rnea.cuh.
//
// Passes one and two are interleaved along the depth-first traversal (DESCEND = twist of the body, ASCEND =
// articulated inertia / bias force folded into the parent), so only the current root-to-leaf path is live on the
// stack (wide part in tensor memory, narrow part in shared memory); pass three is a second, DESCEND-only sweep.  What
// pass three needs from pass two (g = U / D, k0 = u / D and the sin/cos of the joint: eight doubles per body, program.h:
// MB_ABA_REC) goes through a global workspace with one column per resident thread of the persistent grid, streamed back by
// cp.async through a ring; the twists are recomputed instead of stored.
#pragma once
#include "jointmath.cuh"

namespace mb
{
// pass-three record: MB_ABA_REC doubles per body (program.h)

// record of a one-DoF joint: g without its unit component along the joint axis, k0, sin/cos of the joint angle (prismatic: q, 1)
template <class T, class Ctx, bool REV> MB_HD void aba_rec_st_1dof(Ctx &c, int r, const SvT<T> &g, T k0, T s, T cs)
{
   if (REV)
   {
      c.rec_st2(r + 0, g.a.x, g.a.y);
      c.rec_st2(r + 1, g.l.x, g.l.y);
      c.rec_st2(r + 2, g.l.z, k0);
   }
   else
   {
      c.rec_st2(r + 0, g.a.x, g.a.y);
      c.rec_st2(r + 1, g.a.z, g.l.x);
      c.rec_st2(r + 2, g.l.y, k0);
   }
   c.rec_st2(r + 3, s, cs);
}

template <class T> struct AbaPipe
{
   T s, c;        // current op: sin/cos (prismatic: s = q)
   T qd, x;       // current op: joint velocity, joint effort
   T mq;          // next op: raw configuration
   T ls, lc;      // sin/cos of the last DESCEND (leaf)
};

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

// ---- pass one: twist of the body (the frame tree's lazy twist-of-frame, MovingReferenceFrame.java:279-311)
template <class T, class Ctx, bool REV, bool SC>
MB_HD void aba_descend_1dof(Ctx &c, const MbOp2 o, SvT<T> &v, AbaPipe<T> &pp, T &ns, T &nc)
{
   if (SC)
      mb_sincos(pp.mq, &ns, &nc);
   const XfT<T> X = joint_xf_1dof<T, REV>(c.cst(o.body), pp.s, pp.c);
   v = motion_to_child(X, v);
   if (REV) v.a.z += pp.qd;
   else v.l.z += pp.qd;
   pp.ls = pp.s;
   pp.lc = pp.c;
   if (!(o.flags & MB2_LEAF))
   {
      c.acc_st(o.slot, o.wslot, v.a.x, v.a.y, v.a.z, v.l.x, v.l.y, v.l.z);
      c.jp_st2(o.slot, o.nslot, 0, pp.s, pp.c);
   }
}

// ---- fold a child's articulated inertia / bias wrench (already in the parent frame) into the parent's accumulator
template <class T, class Ctx> MB_HD void aba_fold(Ctx &c, const MbOp2 o, const AbiT<T> &K, const SvT<T> &Pp, AbiT<T> &acc, SvT<T> &pacc)
{
   if (o.flags & MB2_FIRST_CHILD)
   {
      acc = K;
      pacc = Pp;
   }
   else
   {
      aux_ld_abi<T>(c, o.paux, acc, pacc);
      acc = acc + K;
      pacc = pacc + Pp;
   }
   if (o.flags & MB2_STORE_ACC)
      aux_st_abi<T>(c, o.paux, acc, pacc);
}

} // namespace mb
#include "multidof_aba.cuh"
namespace mb
{
// ---- pass two of a joint in JointSourceMode.ACCELERATION_SOURCE (ForwardDynamicsCalculator.java:45-57, :1237-1253): its
// acceleration is an input, so nothing is removed from the articulated inertia and the known S qdd enters the bias wrench:
// I^a = I^A, p^a = p^A + I^A (c + S qdd).  Pass three reads qdd back through the same record (g = 0, k0 = qdd).
template <class T, class Ctx, bool REV, class CP>
MB_HD void aba_ascend_1dof_locked(Ctx &c, const MbOp2 o, const CP C, const SvT<T> &vb, T qd, T s, T cs, const AbiT<T> &IA, const SvT<T> &pA,
                                  AbiT<T> &acc, SvT<T> &pacc)
{
   const T qdd = c.ld_x2(o.dof);
   const int r = o.body * (MB_ABA_REC / 2);
   c.rec_st2(r + 0, (T)0, (T)0);
   c.rec_st2(r + 1, (T)0, (T)0);
   c.rec_st2(r + 2, (T)0, qdd); // pass three takes k0 as the acceleration (MB2_ACCSRC on its record)
   c.rec_st2(r + 3, s, cs);
   if (!(o.flags & MB2_ROOT_PARENT))
   {
      SvT<T> cc;
      if (REV)
      {
         cc.a = v3<T>(vb.a.y * qd, -(vb.a.x * qd), qdd);
         cc.l = v3<T>(vb.l.y * qd, -(vb.l.x * qd), (T)0);
      }
      else
      {
         cc.a = v3<T>((T)0, (T)0, (T)0);
         cc.l = v3<T>(vb.a.y * qd, -(vb.a.x * qd), qdd);
      }
      const SvT<T> pa = pA + mul(IA, cc);
      const XfT<T> X = joint_xf_1dof<T, REV>(C, s, cs);
      aba_fold<T>(c, o, abi_to_parent<T, 0>(X, IA), force_to_parent(X, pa), acc, pacc);
   }
}

// ---- pass one quantities (bias wrench / bias acceleration, :1109-1118) and pass two (:1136-1254) for body i
template <class T, class Ctx, bool FEXT, bool REV, bool SC>
MB_HD void aba_ascend_1dof(Ctx &c, const MbOp2 o, int ext, const SvT<T> &v, AbiT<T> &acc, SvT<T> &pacc, AbaPipe<T> &pp, T &ns, T &nc)
{
   if (SC)
      mb_sincos(pp.mq, &ns, &nc);
   const auto C = c.cst(o.body);
   // twist and sin/cos of the body: still in registers for a leaf (its DESCEND was the op before), otherwise on the stack
   T s, cs;
   SvT<T> vb;
   if (o.flags & MB2_LEAF)
   {
      vb = v;
      s = pp.ls;
      cs = pp.lc;
   }
   else
   {
      c.acc_ld(o.slot, o.wslot, vb.a.x, vb.a.y, vb.a.z, vb.l.x, vb.l.y, vb.l.z);
      c.jp_ld2(o.slot, o.nslot, 0, s, cs);
   }
   const T qd = pp.qd, tau = pp.x;
   const RbiT<T> I = ld_rbi<T>(C);
   // p^A = v x* (I v) [- f_ext] + the children's, I^A = I + the children's (the accumulated terms ride on the multiply-add chains)
   SvT<T> pA;
   AbiT<T> IA;
   {
      const SvT<T> Iv = mul(I, vb);
      if (o.flags & MB2_LEAF)
      {
         pA = cross_force(vb, Iv);
         IA = abi_from_rbi(I);
      }
      else
      {
         pA = cross_force_add(vb, Iv, pacc);
         IA = abi_add_rbi(acc, I);
      }
   }
   if (FEXT && c.has_fext())
      pA = pA - external_wrench<T>(c, ext, C);
   if (FEXT && (o.flags & MB2_ACCSRC))
   {
      aba_ascend_1dof_locked<T, Ctx, REV>(c, o, C, vb, qd, s, cs, IA, pA, acc, pacc);
      return;
   }
   SvT<T> U;
   T D, u;
   if (REV)
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
   const T Dinv = mb_rcp(D);
   SvT<T> g;
   g.a = Dinv * U.a;
   g.l = Dinv * U.l;
   const T k0 = Dinv * u;
   // record for pass three: qdd = k0 - g . a'  (the component of g along the joint axis is D / D = 1 and is not stored)
   aba_rec_st_1dof<T, Ctx, REV>(c, o.body * (MB_ABA_REC / 2), g, k0, s, cs);
   if (!(o.flags & MB2_ROOT_PARENT))
   {
      // bias acceleration c = v x (S qd): only x / y components
      T cax, cay, clx, cly;
      if (REV)
      {
         cax = vb.a.y * qd; cay = -(vb.a.x * qd);
         clx = vb.l.y * qd; cly = -(vb.l.x * qd);
      }
      else
      {
         cax = (T)0; cay = (T)0;
         clx = vb.a.y * qd; cly = -(vb.a.x * qd);
      }
      constexpr int Z = REV ? 1 : 2;
      const AbiT<T> Ia = abi_downdate<T, Z>(IA, U, g); // I^a = I^A - U D^-1 U^T: zero along the joint's own direction
      SvT<T> pa = pA;                               // p^a = p^A + I^a c + U D^-1 u
      if (REV)
      {
         pa.a.x = fmad(Ia.A.xx, cax, fmad(Ia.A.xy, cay, fmad(Ia.C.xx, clx, fmad(Ia.C.xy, cly, fmad(k0, U.a.x, pa.a.x)))));
         pa.a.y = fmad(Ia.A.xy, cax, fmad(Ia.A.yy, cay, fmad(Ia.C.yx, clx, fmad(Ia.C.yy, cly, fmad(k0, U.a.y, pa.a.y)))));
         pa.a.z = fmad(k0, U.a.z, pa.a.z);
         pa.l.x = fmad(Ia.C.xx, cax, fmad(Ia.C.yx, cay, fmad(Ia.L.xx, clx, fmad(Ia.L.xy, cly, fmad(k0, U.l.x, pa.l.x)))));
         pa.l.y = fmad(Ia.C.xy, cax, fmad(Ia.C.yy, cay, fmad(Ia.L.xy, clx, fmad(Ia.L.yy, cly, fmad(k0, U.l.y, pa.l.y)))));
         pa.l.z = fmad(Ia.C.xz, cax, fmad(Ia.C.yz, cay, fmad(Ia.L.xz, clx, fmad(Ia.L.yz, cly, fmad(k0, U.l.z, pa.l.z)))));
      }
      else
      {
         // c = [0; w x e_z qd]: only the linear x / y components
         pa.a.x = fmad(Ia.C.xx, clx, fmad(Ia.C.xy, cly, fmad(k0, U.a.x, pa.a.x)));
         pa.a.y = fmad(Ia.C.yx, clx, fmad(Ia.C.yy, cly, fmad(k0, U.a.y, pa.a.y)));
         pa.a.z = fmad(Ia.C.zx, clx, fmad(Ia.C.zy, cly, fmad(k0, U.a.z, pa.a.z)));
         pa.l.x = fmad(Ia.L.xx, clx, fmad(Ia.L.xy, cly, fmad(k0, U.l.x, pa.l.x)));
         pa.l.y = fmad(Ia.L.xy, clx, fmad(Ia.L.yy, cly, fmad(k0, U.l.y, pa.l.y)));
         pa.l.z = fmad(k0, U.l.z, pa.l.z);
      }
      const XfT<T> X = joint_xf_1dof<T, REV>(C, s, cs);
      aba_fold<T>(c, o, abi_to_parent<T, Z>(X, Ia), force_to_parent(X, pa), acc, pacc); // :1159-1165
   }
}

template <class T, class Ctx> MB_HD void aba_descend_6dof(Ctx &c, const MbOp2 o, SvT<T> &v)
{
   const int sub = mb_sub_of<Ctx>(o);
   const SvT<T> vj = ld_svj<T>(o.dof, sub, [&](int r) { return c.ld_qd(r); });
   if (o.flags & MB2_ROOT_PARENT)
      v = vj; // a floating base: the root body is at rest, and nothing is folded into it, so passes one and two need no transform
   else
   {
      const XfT<T> X = joint_xf_multi<T>(c, c.cst(o.body), o.cfg, sub);
      v = motion_to_child(X, v) + vj;
      jp_st_xf<T>(c, o.slot, o.nslot, X);
   }
   c.acc_st(o.slot, o.wslot, v.a.x, v.a.y, v.a.z, v.l.x, v.l.y, v.l.z);
}

template <class T, class Ctx, bool FEXT>
MB_HD void aba_ascend_6dof(Ctx &c, const MbOp2 o, int ext, int rec_hi, AbiT<T> &acc, SvT<T> &pacc)
{
   if (mb_sub_of<Ctx>(o) == MB_SUB_SPHERICAL)
      return aba_ascend_3dof<T, Ctx, FEXT, false>(c, o, ext, rec_hi, acc, pacc);
   if (mb_sub_of<Ctx>(o) == MB_SUB_PLANAR)
      return aba_ascend_3dof<T, Ctx, FEXT, true>(c, o, ext, rec_hi, acc, pacc);
   const auto C = c.cst(o.body);
   SvT<T> vb;
   c.acc_ld(o.slot, o.wslot, vb.a.x, vb.a.y, vb.a.z, vb.l.x, vb.l.y, vb.l.z);
   const RbiT<T> I = ld_rbi<T>(C);
   // p^A = v x* (I v) [- f_ext] + the children's, I^A = I + the children's (the accumulated terms ride on the multiply-add chains)
   SvT<T> pA;
   AbiT<T> IA;
   {
      const SvT<T> Iv = mul(I, vb);
      if (o.flags & MB2_LEAF)
      {
         pA = cross_force(vb, Iv);
         IA = abi_from_rbi(I);
      }
      else
      {
         pA = cross_force_add(vb, Iv, pacc);
         IA = abi_add_rbi(acc, I);
      }
   }
   if (FEXT && c.has_fext())
      pA = pA - external_wrench<T>(c, ext, C);
   const int r = o.body * (MB_ABA_REC / 2);
   if (FEXT && (o.flags & MB2_ACCSRC))
   {
      // ACCELERATION_SOURCE (:1237-1253): the record carries qdd itself (pass three, MB2_ACCSRC on its record: a = a' + qdd)
      const SvT<T> qdd6 = ld_sv6<T>(o.dof, [&](int rr) { return c.ld_x2(rr); });
      c.rec_st2(r + 0, qdd6.a.x, qdd6.a.y);
      c.rec_st2(r + 1, qdd6.a.z, qdd6.l.x);
      c.rec_st2(r + 2, qdd6.l.y, qdd6.l.z);
      c.rec_st2(r + 3, (T)0, (T)0);
      if (!(o.flags & MB2_ROOT_PARENT))
      {
         const XfT<T> X = jp_ld_xf<T>(c, o.slot, o.nslot);
         const SvT<T> vj = ld_sv6<T>(o.dof, [&](int rr) { return c.ld_qd(rr); });
         const SvT<T> pa = pA + mul(IA, cross_motion(vb, vj) + qdd6);
         aba_fold<T>(c, o, abi_to_parent<T, 0>(X, IA), force_to_parent(X, pa), acc, pacc);
      }
      return;
   }
   const SvT<T> tau6 = ld_sv6<T>(o.dof, [&](int r) { return c.ld_x(r); });
   if (o.flags & MB2_ROOT_PARENT)
   {
      // a floating base ascends last: pass three starts with it and reads its quaternion and velocity rows directly -- ask L2 for
      // them now (the wait for them was the largest single stall of the kernel, 2.4 % of its samples)
      c.warm_q(o.cfg); c.warm_q(o.cfg + 1); c.warm_q(o.cfg + 2); c.warm_q(o.cfg + 3);
      c.warm_qd(o.dof); c.warm_qd(o.dof + 1); c.warm_qd(o.dof + 2); c.warm_qd(o.dof + 3); c.warm_qd(o.dof + 4); c.warm_qd(o.dof + 5);
   }
   // D = I^A, U = I^A: a_i = D^-1 u, and the joint transmits nothing but tau to its parent
   const SvT<T> x = abi_solve(IA, tau6 - pA);
   c.rec_st2(r + 0, x.a.x, x.a.y);
   c.rec_st2(r + 1, x.a.z, x.l.x);
   c.rec_st2(r + 2, x.l.y, x.l.z);
   c.rec_st2(r + 3, (T)0, (T)0); // (the whole record travels through the pass-three ring)
   if (!(o.flags & MB2_ROOT_PARENT))
   {
      const XfT<T> X = jp_ld_xf<T>(c, o.slot, o.nslot);
      const SvT<T> Pp = force_to_parent(X, tau6);
      if (o.flags & MB2_FIRST_CHILD)
      {
         acc = AbiT<T>();
         pacc = Pp;
      }
      else
      {
         aux_ld_abi<T>(c, o.paux, acc, pacc);
         pacc = pacc + Pp;
      }
      if (o.flags & MB2_STORE_ACC)
         aux_st_abi<T>(c, o.paux, acc, pacc);
   }
}

// ---- pass three (:1259-1310): accelerations, root to leaves
// LOCKS: the instantiation that serves joints in ACCELERATION_SOURCE mode (their records hold the given acceleration)
template <class T, class Ctx, bool REV, bool LOCKS>
MB_HD void aba_pass3_1dof(Ctx &c, const MbOp2 o, int st, SvT<T> &v, SvT<T> &a, T qd)
{
   T g0, g1, g2, g3, g4, k0, s, cs;
   c.pf3_ld2(st, 1, g0, g1);
   c.pf3_ld2(st, 2, g2, g3);
   c.pf3_ld2(st, 3, g4, k0);
   c.pf3_ld2(st, 4, s, cs);
   c.rec_discard(o.body * (MB_ABA_REC / 2));
   const XfT<T> X = joint_xf_1dof<T, REV>(c.cst(o.body), s, cs);
   v = motion_to_child(X, v);
   a = motion_to_child(X, a); // a' = X^-1 a_parent + c
   if (REV)
   {
      a.a.x += v.a.y * qd; a.a.y -= v.a.x * qd;
      a.l.x += v.l.y * qd; a.l.y -= v.l.x * qd;
      v.a.z += qd;
   }
   else
   {
      a.l.x += v.a.y * qd; a.l.y -= v.a.x * qd;
      v.l.z += qd;
   }
   // D^-1 (u - U^T a'): the stored five components of g, plus a' along the joint axis (g = 1 there)
   T qdd;
   if (REV)
      qdd = k0 - (fmad(g0, a.a.x, fmad(g1, a.a.y, fmad(g2, a.l.x, fmad(g3, a.l.y, g4 * a.l.z)))) + a.a.z);
   else
      qdd = k0 - (fmad(g0, a.a.x, fmad(g1, a.a.y, fmad(g2, a.a.z, fmad(g3, a.l.x, g4 * a.l.y)))) + a.l.z);
   if (LOCKS && (o.flags & MB2_ACCSRC))
      qdd = k0; // the joint's given acceleration
   c.st_out(o.dof, qdd);
   if (REV) a.a.z += qdd;
   else a.l.z += qdd;
   if (o.flags & MB2_SAVE_STATE)
   {
      aux_st_sv<T>(c, o.aux, v);
      aux_st_sv<T>(c, o.aux + 6, a);
   }
}

template <class T, class Ctx, bool LOCKS> MB_HD void aba_pass3_6dof(Ctx &c, const MbOp2 o, int st, int rec_hi, SvT<T> &v, SvT<T> &a)
{
   if (mb_sub_of<Ctx>(o) == MB_SUB_SPHERICAL)
      return aba_pass3_3dof<T, Ctx, LOCKS, false>(c, o, st, rec_hi, v, a);
   if (mb_sub_of<Ctx>(o) == MB_SUB_PLANAR)
      return aba_pass3_3dof<T, Ctx, LOCKS, true>(c, o, st, rec_hi, v, a);
   SvT<T> x;
   c.pf3_ld2(st, 1, x.a.x, x.a.y);
   c.pf3_ld2(st, 2, x.a.z, x.l.x);
   c.pf3_ld2(st, 3, x.l.y, x.l.z);
   c.rec_discard(o.body * (MB_ABA_REC / 2));
   const SvT<T> vj = ld_sv6<T>(o.dof, [&](int rr) { return c.ld_qd(rr); });
   SvT<T> a1;
   if (o.flags & MB2_ROOT_PARENT)
   {
      // a floating base: the root body is at rest and its acceleration is -gravity, purely linear: v = vj, v x vj = 0, and only
      // the rotation of the joint is needed
      const auto C = c.cst(o.body);
      M3T<T> R0;
      V3T<T> p0;
      ld_xf0<T>(C, R0, p0);
      const M3T<T> R = mul(R0, quat_to_rot<T, Ctx::kFastQuat>(c.ld_q(o.cfg), c.ld_q(o.cfg + 1), c.ld_q(o.cfg + 2), c.ld_q(o.cfg + 3)));
      v = vj;
      a1.a = v3<T>((T)0, (T)0, (T)0);
      a1.l = mulT(R, a.l);
   }
   else
   {
      const XfT<T> X = joint_xf_6dof<T>(c, c.cst(o.body), o.cfg);
      v = motion_to_child(X, v) + vj;
      a1 = motion_to_child(X, a) + cross_motion(v, vj);
   }
   // effort source: a = x, qdd = x - a'; acceleration source (the record holds the given acceleration): qdd = x, a = a' + x
   SvT<T> qdd;
   if (LOCKS && (o.flags & MB2_ACCSRC))
   {
      qdd = x;
      x = a1 + x;
   }
   else
   {
      qdd.a = x.a - a1.a;
      qdd.l = x.l - a1.l;
   }
   c.st_out(o.dof + 0, qdd.a.x); c.st_out(o.dof + 1, qdd.a.y); c.st_out(o.dof + 2, qdd.a.z);
   c.st_out(o.dof + 3, qdd.l.x); c.st_out(o.dof + 4, qdd.l.y); c.st_out(o.dof + 5, qdd.l.z);
   a = x;
   if (o.flags & MB2_SAVE_STATE)
   {
      aux_st_sv<T>(c, o.aux, v);
      aux_st_sv<T>(c, o.aux + 6, a);
   }
}

// which scalars an op needs from the prefetch ring: DESCEND q + qd, ASCEND qd + tau
MB_HD int aba_pf_mask(const MbOp2 &o) { return MB2_JT(o.code) == MB_SIXDOF ? 0 : ((o.code & MB2_ASCEND) ? 6 : 3); }

// ---- the part of an op of passes one + two that does not depend on its kind (see rnea_pre)
// rev1: op k is a revolute DESCEND (its sin/cos were evaluated during the previous op from the raw angle still in pp.mq: rnea_pre)
template <class T, class Ctx>
MB_HD void aba_pre(Ctx &c, const int k, const MbOp2 &o, const bool onedof, const bool asc, const bool rev1, SvT<T> &v, AbaPipe<T> &pp)
{
   if (rev1 && mb_angle_large(pp.mq))
      mb_sincos_redo(pp.mq, pp.s, pp.c);
   if (asc)
      c.stk_fence(); // an ASCEND reads its twist back from the wide stack (a DESCEND only when it reloads its parent's, below)
   if (o.pf & (MB2_PF_D1 | MB2_PF_A1))
      c.pf_issue((k + MB_PF_DIST) & (MB_PF_STAGES - 1), o.pfcfg, o.pfdof, (o.pf & MB2_PF_A1) ? 6 : 3);
   c.pf_commit();
   c.template pf_wait<MB_PF_DIST - 1>();
   pp.qd = pp.x = pp.mq = (T)0;
   if (onedof)
   {
      pp.qd = c.pf_ld(k & (MB_PF_STAGES - 1), 1);
      if (asc)
         pp.x = c.pf_ld(k & (MB_PF_STAGES - 1), 2);
   }
   if (o.pf & MB2_PF_NEXT1)
      pp.mq = c.pf_ld((k + 1) & (MB_PF_STAGES - 1), 0); // raw: checked when op k + 1 starts
   // twist of the parent: carried along a chain, zero for the root body, otherwise on the parent's stack slot
   if (!asc && (o.flags & (MB2_ROOT_PARENT | MB2_LOAD_PARENT)))
   {
      if (o.flags & MB2_ROOT_PARENT)
         v = sv_zero<T>();
      else
      {
         c.stk_fence();
         c.acc_ld(o.pslot, o.pwslot, v.a.x, v.a.y, v.a.z, v.l.x, v.l.y, v.l.z);
      }
   }
}

// ---- one op of passes one + two (the unit the tree-specialised kernels are generated from; see rnea_op)
template <class T, class Ctx, bool FEXT>
MB_HD void aba_op(Ctx &c, const int k, const MbOp2 o, const int ext, SvT<T> &v, AbiT<T> &acc, SvT<T> &pacc, AbaPipe<T> &pp)
{
   c.op_sync(k);
   aba_pre<T, Ctx>(c, k, o, MB2_JT(o.code) != MB_SIXDOF, (o.code & MB2_ASCEND) != 0, !(o.code & MB2_ASCEND) && MB2_JT(o.code) == MB_REVOLUTE, v, pp);
   T ns = pp.mq, nc = (T)1;
   switch (o.code & 0xfu)
   {
      case 0 | (MB_REVOLUTE << 1): aba_descend_1dof<T, Ctx, true, false>(c, o, v, pp, ns, nc); break;
      case 0 | (MB_REVOLUTE << 1) | MB2_SC: aba_descend_1dof<T, Ctx, true, true>(c, o, v, pp, ns, nc); break;
      case 1 | (MB_REVOLUTE << 1): aba_ascend_1dof<T, Ctx, FEXT, true, false>(c, o, ext, v, acc, pacc, pp, ns, nc); break;
      case 1 | (MB_REVOLUTE << 1) | MB2_SC: aba_ascend_1dof<T, Ctx, FEXT, true, true>(c, o, ext, v, acc, pacc, pp, ns, nc); break;
      case 0 | (MB_PRISMATIC << 1): aba_descend_1dof<T, Ctx, false, false>(c, o, v, pp, ns, nc); break;
      case 0 | (MB_PRISMATIC << 1) | MB2_SC: aba_descend_1dof<T, Ctx, false, true>(c, o, v, pp, ns, nc); break;
      case 1 | (MB_PRISMATIC << 1): aba_ascend_1dof<T, Ctx, FEXT, false, false>(c, o, ext, v, acc, pacc, pp, ns, nc); break;
      case 1 | (MB_PRISMATIC << 1) | MB2_SC: aba_ascend_1dof<T, Ctx, FEXT, false, true>(c, o, ext, v, acc, pacc, pp, ns, nc); break;
      default:
         if (o.code & MB2_SC)
            mb_sincos(pp.mq, &ns, &nc);
         if (o.code & MB2_ASCEND)
            aba_ascend_6dof<T, Ctx, FEXT>(c, o, ext, -1, acc, pacc); // (tree-specialised kernels: no three-DoF joints, api.cu)
         else
            aba_descend_6dof<T, Ctx>(c, o, v);
         break;
   }
   pp.s = ns;
   pp.c = nc;
}

template <class T, class Ctx>
MB_HD void aba_begin(Ctx &c, const MbOp2 o0, const MbOp2 o1, const MbOp2 o2, SvT<T> &v, AbiT<T> &acc, SvT<T> &pacc, AbaPipe<T> &pp)
{
   static_assert(MB_PF_DIST == 3, "prologue written for a prefetch distance of 3");
   v = sv_zero<T>(); pacc = sv_zero<T>();
   acc = AbiT<T>();
   pp.s = pp.qd = pp.x = pp.mq = pp.ls = (T)0;
   pp.c = pp.lc = (T)1;
   if (aba_pf_mask(o0)) c.pf_issue(0, o0.cfg, o0.dof, aba_pf_mask(o0));
   c.pf_commit();
   if (aba_pf_mask(o1)) c.pf_issue(1, o1.cfg, o1.dof, aba_pf_mask(o1));
   c.pf_commit();
   if (aba_pf_mask(o2)) c.pf_issue(2, o2.cfg, o2.dof, aba_pf_mask(o2));
   c.pf_commit();
   if (mb2_is_1dof_descend(o0))
   {
      // (only a one-DoF first op needs its angle now; a floating base reads its rows directly, and waiting here would put the latency
      // of the ring in front of the latency of those loads instead of next to it)
      c.template pf_wait<0>();
      const T q0 = c.pf_ld(0, 0);
      if (MB2_JT(o0.code) == MB_REVOLUTE) mb_sincos(mb_reduce_angle(q0), &pp.s, &pp.c);
      else pp.s = q0; // a prismatic displacement is not an angle
   }
}

// ---- pass three: DESCEND records only.  Same software pipeline, but the ring (Ctx::pf3_*, overlaid on the now idle
// stack area) carries the pass-two record of each body -- which includes the sin/cos of its joint -- and its joint velocity
template <class T, class Ctx>
MB_HD void aba_pass3_begin(Ctx &c, const MbOp2 o0, const MbOp2 o1, const MbOp2 o2, SvT<T> &v, SvT<T> &a, AbaPipe<T> &pp)
{
   c.template pf_wait<0>();
   c.pass_fence(); // the records written in pass two are read back below (same thread)
   a = sv_zero<T>();
   v = sv_zero<T>();
   c.pf3_issue(0, o0.cfg, o0.dof, o0.body * (MB_ABA_REC / 2), mb2_is_1dof_descend(o0) ? 2 : 0);
   c.pf_commit();
   c.pf3_issue(1, o1.cfg, o1.dof, o1.body * (MB_ABA_REC / 2), mb2_is_1dof_descend(o1) ? 2 : 0);
   c.pf_commit();
   c.pf3_issue(2, o2.cfg, o2.dof, o2.body * (MB_ABA_REC / 2), mb2_is_1dof_descend(o2) ? 2 : 0);
   c.pf_commit();
   (void)pp;
}

// ---- the kind-independent part of a pass-three op; returns the joint velocity of a one-DoF joint
template <class T, class Ctx>
MB_HD T aba_pass3_pre(Ctx &c, const int k, const MbOp2 &o, const bool onedof, const T *grav, SvT<T> &v, SvT<T> &a)
{
   c.pf3_issue((k + MB_PF_DIST) & (MB_PF_STAGES - 1), o.pfcfg, o.pfdof, o.pfbody * (MB_ABA_REC / 2), (o.pf & MB2_PF_D1) ? 2 : 0);
   c.pf_commit();
   c.template pf_wait<MB_PF_DIST - 1>();
   T qd = (T)0;
   if (onedof)
   {
      T unused;
      c.pf3_ld2(k & (MB_PF_STAGES - 1), 0, unused, qd);
   }
   if (o.flags & (MB2_ROOT_PARENT | MB2_LOAD_PARENT))
   {
      if (o.flags & MB2_ROOT_PARENT)
      {
         v = sv_zero<T>();
         a = sv_zero<T>();
         a.l = v3<T>(-grav[0], -grav[1], -grav[2]); // root acceleration = -gravity (ForwardDynamicsCalculator.java:313-319)
      }
      else
      {
         v = aux_ld_sv<T>(c, o.paux);
         a = aux_ld_sv<T>(c, o.paux + 6);
      }
   }
   return qd;
}

template <class T, class Ctx, bool LOCKS = false>
MB_HD void aba_pass3_op(Ctx &c, const int k, const MbOp2 o, const T *grav, SvT<T> &v, SvT<T> &a, AbaPipe<T> &pp)
{
   (void)pp;
   c.op_sync(k);
   const T qd = aba_pass3_pre<T, Ctx>(c, k, o, MB2_JT(o.code) != MB_SIXDOF, grav, v, a);
   const int st = k & (MB_PF_STAGES - 1);
   switch ((o.code >> 1) & 3u)
   {
      case MB_REVOLUTE: aba_pass3_1dof<T, Ctx, true, LOCKS>(c, o, st, v, a, qd); break;
      case MB_PRISMATIC: aba_pass3_1dof<T, Ctx, false, LOCKS>(c, o, st, v, a, qd); break;
      default: aba_pass3_6dof<T, Ctx, LOCKS>(c, o, st, -1, v, a); break;
   }
}

// ---- run steps: as aba_op / aba_pass3_op with the kind (ASCEND / joint type) fixed at compile time (see rnea.cuh).  Whether
// the op also evaluates the sin/cos of the next joint (SC) is part of the kind for the small ops -- DESCEND and pass three, where
// the sin/cos chain then shares a basic block with the op's own arithmetic instead of running serially in front of it -- and
// tested at run time for the large ASCEND ops (one loop body per joint type: their code has to stay in the instruction cache)
template <class T, class Ctx, bool FEXT, int KIND>
MB_HD void aba_run_step(const MbProgram &P, Ctx &c, const int k, SvT<T> &v, AbiT<T> &acc, SvT<T> &pacc, AbaPipe<T> &pp)
{
   constexpr bool ASC = (KIND & MB2_ASCEND) != 0, SCK = (KIND & MB2_SC) != 0, PLAIN = (KIND & MB_RUN_PLAIN) != 0;
   constexpr int JT = (KIND >> 1) & 3;
   static_assert(!(ASC && SCK), "ASCEND kinds do not carry the SC bit");
   constexpr bool DYN_SC = ASC || !MB_ABA_D_SC_SPLIT; // SC tested at run time (not part of the run kind)
   MbOp2 o = P.op2[k];
   // plain runs (rnea.cuh: rnea_run_step): the flags of the op are those of the common case of its kind, i.e. constants
   if (PLAIN)
   {
      o.flags = (uint8_t)mb_run_plain_flags(MB_ABA, KIND & 0xf, false);
      o.pf = (!ASC && SCK) ? (uint8_t)(o.pf | MB2_PF_NEXT1) : (ASC ? o.pf : (uint8_t)(o.pf & ~MB2_PF_NEXT1));
   }
   else if (!DYN_SC && SCK)
      o.pf |= MB2_PF_NEXT1; // the SC bit implies that a one-DoF DESCEND follows
   aba_pre<T, Ctx>(c, k, o, JT != MB_SIXDOF, ASC, !ASC && JT == MB_REVOLUTE, v, pp);
   const int ext = FEXT ? P.body[o.body].ext_index : 0;
   T ns = pp.mq, nc = (T)1;
   if (DYN_SC || JT == MB_SIXDOF)
   {
      if (DYN_SC ? (o.code & MB2_SC) != 0 : SCK)
         mb_sincos(pp.mq, &ns, &nc);
   }
   if (JT == MB_SIXDOF)
   {
      if (ASC) aba_ascend_6dof<T, Ctx, FEXT>(c, o, ext, P.body[o.body].rec, acc, pacc);
      else aba_descend_6dof<T, Ctx>(c, o, v);
   }
   else if (ASC)
      aba_ascend_1dof<T, Ctx, FEXT, JT == MB_REVOLUTE, false>(c, o, ext, v, acc, pacc, pp, ns, nc);
   else
      aba_descend_1dof<T, Ctx, JT == MB_REVOLUTE, SCK>(c, o, v, pp, ns, nc);
   pp.s = ns;
   pp.c = nc;
}

template <class T, class Ctx, bool LOCKS, int KIND>
MB_HD void aba_pass3_run_step(const MbProgram &P, Ctx &c, const int k, const T *grav, SvT<T> &v, SvT<T> &a)
{
   constexpr int JT = (KIND >> 1) & 3;
   constexpr bool PLAIN = (KIND & MB_RUN_PLAIN) != 0;
   MbOp2 o = P.op3[k];
   if (PLAIN)
      o.flags = (uint8_t)mb_run_plain_flags(MB_ABA, KIND & 0xf, true);
   const T qd = aba_pass3_pre<T, Ctx>(c, k, o, JT != MB_SIXDOF, grav, v, a);
   const int st = k & (MB_PF_STAGES - 1);
   if (JT == MB_SIXDOF)
      aba_pass3_6dof<T, Ctx, LOCKS>(c, o, st, P.body[o.body].rec, v, a);
   else
      aba_pass3_1dof<T, Ctx, JT == MB_REVOLUTE, LOCKS>(c, o, st, v, a, qd);
}

template <class T, class Ctx, bool FEXT> MB_HD void aba_state(const MbProgram &P, Ctx &c, const T *grav)
{
   SvT<T> v, pacc, a;
   AbiT<T> acc;
   AbaPipe<T> pp;
   // =========================== passes one and two, interleaved; one tight loop per run of same-kind ops
   aba_begin<T, Ctx>(c, P.op2[0], P.op2[1], P.op2[2], v, acc, pacc, pp);
   const int nruns = P.nruns;
#pragma unroll 1
   for (int r = 0; r < nruns; r++)
   {
      const MbRun R = P.run[r];
      int k = R.k0;
      const int k1 = k + R.n;
#define MB_RUN_CASE(KIND)                                                                          \
   case KIND:                                                                                       \
      _Pragma("unroll 1") do { aba_run_step<T, Ctx, FEXT, KIND>(P, c, k, v, acc, pacc, pp); } while (++k < k1); \
      break;
#define MB_RUN_CASE_IF(COND, KIND)                                                                 \
   case KIND:                                                                                       \
      if constexpr ((COND) != 0)                                                                    \
         _Pragma("unroll 1") do { aba_run_step<T, Ctx, FEXT, KIND>(P, c, k, v, acc, pacc, pp); } while (++k < k1); \
      break;
      switch (R.kind)
      {
         MB_RUN_CASE(0) MB_RUN_CASE(1) MB_RUN_CASE(2) MB_RUN_CASE(3) MB_RUN_CASE(4) MB_RUN_CASE(5)
#if MB_ABA_D_SC_SPLIT
         MB_RUN_CASE(8) MB_RUN_CASE(10) MB_RUN_CASE(12)
#endif
#if MB_PLAIN_ABA & 7
         MB_RUN_CASE_IF(MB_PLAIN_ABA & 1, MB_RUN_PLAIN | 0) MB_RUN_CASE_IF(MB_PLAIN_ABA & 2, MB_RUN_PLAIN | 1) MB_RUN_CASE_IF(MB_PLAIN_ABA & 4, MB_RUN_PLAIN | 8)
#endif
         default: break;
      }
#undef MB_RUN_CASE
#undef MB_RUN_CASE_IF
   }
   // =========================== pass three
   aba_pass3_begin<T, Ctx>(c, P.op3[0], P.op3[1], P.op3[2], v, a, pp);
   const int nruns3 = P.nruns3;
#pragma unroll 1
   for (int r = 0; r < nruns3; r++)
   {
      const MbRun R = P.run3[r];
      int k = R.k0;
      const int k1 = k + R.n;
#define MB_RUN_CASE(KIND)                                                                       \
   case KIND:                                                                                    \
      _Pragma("unroll 1") do { aba_pass3_run_step<T, Ctx, FEXT, KIND>(P, c, k, grav, v, a); } while (++k < k1); \
      break;
#define MB_RUN_CASE_IF(COND, KIND)                                                              \
   case KIND:                                                                                    \
      if constexpr ((COND) != 0)                                                                 \
         _Pragma("unroll 1") do { aba_pass3_run_step<T, Ctx, FEXT, KIND>(P, c, k, grav, v, a); } while (++k < k1); \
      break;
      switch (R.kind)
      {
         MB_RUN_CASE(0) MB_RUN_CASE(2) MB_RUN_CASE(4)
#if MB_PLAIN_ABA & 8
         MB_RUN_CASE_IF(MB_PLAIN_ABA & 8, MB_RUN_PLAIN | 0)
#endif
         default: break;
      }
#undef MB_RUN_CASE
#undef MB_RUN_CASE_IF
   }
   c.template pf_wait<0>();
}
} // namespace mb
