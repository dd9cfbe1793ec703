// rnea.cuh -- one state of the recursive Newton-Euler algorithm (InverseDynamicsCalculator.compute(),
// M/algorithms/InverseDynamicsCalculator.java:496-501, passOne :873-917, passTwo :930-966), including the
// "set state -> updateFramesRecursively()" prologue that Mecano keeps in the frame tree
// (RigidBodyBasics.java:104-112, MovingReferenceFrame.java:279-311).
//
// Shape of the code (what the sm_100a kernel needs; the host emulation harness compiles the same source):
//   * the traversal program is pre-decoded into 28-byte records (MbOp2) and cut into runs of same-kind ops (MbRun); a run
//     is one tight loop over a straight-line routine specialised on <ASCEND, joint type, SC[, plain flags]>, so the FP64 work
//     of one op sits in one basic block;
//   * scalars are software-pipelined: while op k runs, the q/qd/qdd of op k+3 are in flight from HBM (cp.async ring) and the
//     sin/cos of op k+1 (the only long serial FP64 chain) is evaluated next to op k's spatial algebra (the SC variants), which
//     gives the in-order SM independent work to issue;
//   * per-level data lives on a per-state stack of double2, state-minor: the accumulated wrench in tensor memory (or shared
//     memory), sin/cos in shared memory.
#pragma once
#include "program.h"
#include "spatial.cuh"
#include "jointmath.cuh"

namespace mb
{
// scalars travelling through the software pipeline
template <class T> struct RneaPipe
{
   T s, c;   // current op: sin/cos (prismatic: s = q)
   T qd, x;  // current op: joint velocity, joint acceleration (read from the prefetch ring at the top of the op)
   T mq;     // next op: raw configuration (read from the prefetch ring)
   T ls, lc; // sin/cos of the last DESCEND (a leaf's ASCEND follows immediately and reuses them)
};

template <class T, class Ctx, bool FEXT, bool REV, bool SC>
MB_HD void rnea_descend_1dof(Ctx &c, const MbOp2 o, int ext, SvT<T> &v, SvT<T> &a, SvT<T> &f, RneaPipe<T> &pp, T &ns, T &nc)
{
   if (SC)
      mb_sincos(pp.mq, &ns, &nc);
   const auto C = c.cst(o.body);
   const XfT<T> X = joint_xf_1dof<T, REV>(C, pp.s, pp.c);
   // pass one (:873-917): twist and acceleration of the body, in its joint frame
   v = motion_to_child(X, v);
   a = motion_to_child(X, a);
   if (REV)
   {
      // a += v x (S qd) + S qdd with S = [e_z; 0]
      a.a.x += v.a.y * pp.qd; a.a.y -= v.a.x * pp.qd;
      a.l.x += v.l.y * pp.qd; a.l.y -= v.l.x * pp.qd;
      v.a.z += pp.qd;
      a.a.z += pp.x;
   }
   else
   {
      // S = [0; e_z]: v x (S qd) = [0; w x e_z qd]
      a.l.x += v.a.y * pp.qd; a.l.y -= v.a.x * pp.qd;
      v.l.z += pp.qd;
      a.l.z += pp.x;
   }
   // Newton-Euler (SpatialInertiaReadOnly.java:229-296), about the joint-frame origin
   S3T<T> J;
   V3T<T> cp;
   T m;
   ld_com_inertia<T>(C, J, cp, m);
   f = newton_euler(J, cp, m, v, a);
   if (FEXT)
   {
      // the FEXT instantiation serves every call with optional buffers: external wrenches in, by-products out
      if (c.has_fext())
         f = f - external_wrench<T>(c, ext, C); // :946
      if (c.has_acc())
         rnea_store_body_acc<T>(c, ext, C, a);
   }
   pp.ls = pp.s;
   pp.lc = pp.c;
   if (!(o.flags & MB2_LEAF))
   {
      c.acc_st(o.slot, o.wslot, f.a.x, f.a.y, f.a.z, f.l.x, f.l.y, f.l.z);
      c.jp_st2(o.slot, o.nslot, 0, pp.s, pp.c);
   }
   if (o.flags & MB2_SAVE_STATE)
   {
      aux_st_sv<T>(c, o.aux, v);
      aux_st_sv<T>(c, o.aux + 6, a);
   }
}

template <class T, class Ctx, bool FEXT, bool REV, bool SC>
MB_HD void rnea_ascend_1dof(Ctx &c, const MbOp2 o, int ext, SvT<T> &f, RneaPipe<T> &pp, T &ns, T &nc)
{
   if (SC)
      mb_sincos(pp.mq, &ns, &nc);
   // pass two (:930-966); f holds the wrench of the whole subtree
   c.st_out(o.dof, REV ? f.a.z : f.l.z); // tau = S^T W (:952-958)
   if (FEXT && c.has_wr())
      rnea_store_joint_wrench<T>(c, ext, c.cst(o.body), f);
   if (FEXT && c.has_rootw() && (o.flags & MB2_ROOT_PARENT))
   {
      // wrench the root body exerts on this subtree, re-expressed in the root frame and summed over the root's children
      T s = pp.ls, cs = pp.lc;
      if (!(o.flags & MB2_LEAF))
         c.jp_ld2(o.slot, o.nslot, 0, s, cs);
      rnea_add_root_wrench<T>(c, force_up_1dof<T, REV>(c.cst(o.body), s, cs, f));
   }
   if (!(o.flags & MB2_ROOT_PARENT))
   {
      T s = pp.ls, cs = pp.lc;
      if (!(o.flags & MB2_LEAF))
         c.jp_ld2(o.slot, o.nslot, 0, s, cs);
      SvT<T> acc;
      c.acc_ld(o.pslot, o.pwslot, acc.a.x, acc.a.y, acc.a.z, acc.l.x, acc.l.y, acc.l.z);
      f = force_up_1dof_add<T, REV>(c.cst(o.body), s, cs, f, acc); // addJointWrenchFromChild (:961-966)
      if (o.flags & MB2_STORE_ACC)
         c.acc_st(o.pslot, o.pwslot, f.a.x, f.a.y, f.a.z, f.l.x, f.l.y, f.l.z);
   }
}

// multi-DoF joints (SixDoF, Spherical, Planar; sub-type in the op code, multidof.cuh): S is a selection of components
template <class T, class Ctx, bool FEXT>
MB_HD void rnea_descend_6dof(Ctx &c, const MbOp2 o, int ext, SvT<T> &v, SvT<T> &a, SvT<T> &f)
{
   const auto C = c.cst(o.body);
   const int sub = mb_sub_of<Ctx>(o);
   const XfT<T> X = joint_xf_multi<T>(c, C, o.cfg, sub);
   const SvT<T> vj = ld_svj<T>(o.dof, sub, [&](int r) { return c.ld_qd(r); });
   const SvT<T> aj = ld_svj<T>(o.dof, sub, [&](int r) { return c.ld_x(r); });
   v = motion_to_child(X, v) + vj;
   a = motion_to_child(X, a) + cross_motion(v, vj) + aj;
   S3T<T> J;
   V3T<T> cp;
   T m;
   ld_com_inertia<T>(C, J, cp, m);
   f = newton_euler(J, cp, m, v, a);
   if (FEXT)
   {
      if (c.has_fext())
         f = f - external_wrench<T>(c, ext, C);
      if (c.has_acc())
         rnea_store_body_acc<T>(c, ext, C, a);
   }
   c.acc_st(o.slot, o.wslot, f.a.x, f.a.y, f.a.z, f.l.x, f.l.y, f.l.z);
   if (!(o.flags & MB2_ROOT_PARENT))
      jp_st_xf<T>(c, o.slot, o.nslot, X);
   if (o.flags & MB2_SAVE_STATE)
   {
      aux_st_sv<T>(c, o.aux, v);
      aux_st_sv<T>(c, o.aux + 6, a);
   }
}

template <class T, class Ctx, bool FEXT> MB_HD void rnea_ascend_6dof(Ctx &c, const MbOp2 o, int ext, SvT<T> &f)
{
   const int sub = mb_sub_of<Ctx>(o);
   if (FEXT && c.has_wr())
      rnea_store_joint_wrench<T>(c, ext, c.cst(o.body), f);
   if (FEXT && c.has_rootw() && (o.flags & MB2_ROOT_PARENT))
      rnea_add_root_wrench<T>(c, force_to_parent(joint_xf_multi<T>(c, c.cst(o.body), o.cfg, sub), f));
   st_svj<T>(o.dof, sub, f, [&](int r, T x) { c.st_out(r, x); }); // tau = S^T W (:952-958)
   if (!(o.flags & MB2_ROOT_PARENT))
   {
      const XfT<T> X = jp_ld_xf<T>(c, o.slot, o.nslot);
      SvT<T> acc;
      c.acc_ld(o.pslot, o.pwslot, acc.a.x, acc.a.y, acc.a.z, acc.l.x, acc.l.y, acc.l.z);
      f = acc + force_to_parent(X, f);
      if (o.flags & MB2_STORE_ACC)
         c.acc_st(o.pslot, o.pwslot, f.a.x, f.a.y, f.a.z, f.l.x, f.l.y, f.l.z);
   }
}

// Prefetch ring (Ctx::pf_*): MB_PF_STAGES slots of (q, qd, x) per state.  Op k issues the asynchronous copies of
// op k + MB_PF_DIST (cp.async on the GPU: no register and no scoreboard is held while the data is in flight),
// evaluates the sin/cos of op k + 1 and consumes the velocity / acceleration of op k.  What op k has to request for
// op k + MB_PF_DIST is pre-decoded into its own record (MbOp2::pf bits, pfcfg, pfdof).
// setConsiderCoriolisAndCentrifugalForces(false) / setConsiderJointAccelerations(false) (:291-306) never reach this
// code: the launcher substitutes a zero row with stride 0 for qd / qdd, which is the same computation.
#define MB_PF_STAGES 4
#define MB_PF_DIST 3

// The part of an op that does not depend on its kind: prefetch of op k + MB_PF_DIST, scalars of op k, raw angle of op
// k + 1, kinematic state of the parent.  DESC1: op k is a 1-DoF DESCEND;  DESC: op k is a DESCEND.
// REV1: op k is a revolute DESCEND (its sin/cos were evaluated during the previous op from the raw angle still in pp.mq)
template <class T, class Ctx>
MB_HD void rnea_pre(Ctx &c, const int k, const MbOp2 &o, const bool desc1, const bool desc, const bool rev1, const T *grav, SvT<T> &v, SvT<T> &a,
                    RneaPipe<T> &pp)
{
   if (rev1 && mb_angle_large(pp.mq))
      mb_sincos_redo(pp.mq, pp.s, pp.c);
   if (!desc)
      c.stk_fence(); // only an ASCEND reads the wide stack back
   if (o.pf & MB2_PF_D1)
      c.pf_issue((k + MB_PF_DIST) & (MB_PF_STAGES - 1), o.pfcfg, o.pfdof, 7);
   c.pf_commit();
   c.template pf_wait<MB_PF_DIST - 1>(); // everything up to the group of op k + 1 has landed
   pp.qd = pp.x = pp.mq = (T)0;
   if (desc1)
   {
      pp.qd = c.pf_ld(k & (MB_PF_STAGES - 1), 1);
      pp.x = c.pf_ld(k & (MB_PF_STAGES - 1), 2);
   }
   if (o.pf & MB2_PF_NEXT1) // op k + 1 is a 1-DoF DESCEND
      pp.mq = c.pf_ld((k + 1) & (MB_PF_STAGES - 1), 0); // raw: checked when op k + 1 starts (mb_angle_large)
   // kinematic state of the parent: carried in registers along a chain, otherwise the root acceleration
   // (= -gravity, InverseDynamicsCalculator.java:397-403) or the state saved by the branching ancestor
   if (desc && (o.flags & (MB2_ROOT_PARENT | MB2_LOAD_PARENT)))
   {
      if (o.flags & MB2_ROOT_PARENT)
      {
         v = sv_zero<T>();
         a = sv_zero<T>();
         a.l = v3<T>(-grav[0], -grav[1], -grav[2]);
      }
      else
      {
         v = aux_ld_sv<T>(c, o.paux);
         a = aux_ld_sv<T>(c, o.paux + 6);
      }
   }
}

// One op of the traversal program: the unit the tree-specialised kernels (specialize.cpp) are generated from -- there
// every argument except the context and the carried state is a literal, so the flag tests, the dispatch switch and all
// table lookups fold away at compile time.   k: index of the op (selects the prefetch-ring stage);  o: this op
template <class T, class Ctx, bool FEXT>
MB_HD void rnea_op(Ctx &c, const int k, const MbOp2 o, const int ext, const T *grav, SvT<T> &v, SvT<T> &a, SvT<T> &f, RneaPipe<T> &pp)
{
   c.op_sync(k);
   rnea_pre<T, Ctx>(c, k, o, mb2_is_1dof_descend(o), !(o.code & MB2_ASCEND), !(o.code & MB2_ASCEND) && MB2_JT(o.code) == MB_REVOLUTE, grav, v, a, pp);
   T ns = pp.mq, nc = (T)1; // prismatic next op: "s" carries q
   switch (o.code & 0xfu)
   {
      case 0 | (MB_REVOLUTE << 1): rnea_descend_1dof<T, Ctx, FEXT, true, false>(c, o, ext, v, a, f, pp, ns, nc); break;
      case 0 | (MB_REVOLUTE << 1) | MB2_SC: rnea_descend_1dof<T, Ctx, FEXT, true, true>(c, o, ext, v, a, f, pp, ns, nc); break;
      case 1 | (MB_REVOLUTE << 1): rnea_ascend_1dof<T, Ctx, FEXT, true, false>(c, o, ext, f, pp, ns, nc); break;
      case 1 | (MB_REVOLUTE << 1) | MB2_SC: rnea_ascend_1dof<T, Ctx, FEXT, true, true>(c, o, ext, f, pp, ns, nc); break;
      case 0 | (MB_PRISMATIC << 1): rnea_descend_1dof<T, Ctx, FEXT, false, false>(c, o, ext, v, a, f, pp, ns, nc); break;
      case 0 | (MB_PRISMATIC << 1) | MB2_SC: rnea_descend_1dof<T, Ctx, FEXT, false, true>(c, o, ext, v, a, f, pp, ns, nc); break;
      case 1 | (MB_PRISMATIC << 1): rnea_ascend_1dof<T, Ctx, FEXT, false, false>(c, o, ext, f, pp, ns, nc); break;
      case 1 | (MB_PRISMATIC << 1) | MB2_SC: rnea_ascend_1dof<T, Ctx, FEXT, false, true>(c, o, ext, f, pp, ns, nc); break;
      default:
         if (o.code & MB2_SC)
            mb_sincos(pp.mq, &ns, &nc);
         if (o.code & MB2_ASCEND)
            rnea_ascend_6dof<T, Ctx, FEXT>(c, o, ext, f);
         else
            rnea_descend_6dof<T, Ctx, FEXT>(c, o, ext, v, a, f);
         break;
   }
   pp.s = ns;
   pp.c = nc;
}

// prologue: the scalars of ops 0 .. MB_PF_DIST-1 are requested, the sin/cos of op 0 evaluated
template <class T, class Ctx>
MB_HD void rnea_begin(Ctx &c, const MbOp2 o0, const MbOp2 o1, const MbOp2 o2, SvT<T> &v, SvT<T> &a, SvT<T> &f, RneaPipe<T> &pp)
{
   static_assert(MB_PF_DIST == 3, "prologue written for a prefetch distance of 3");
   v = sv_zero<T>(); a = sv_zero<T>(); f = sv_zero<T>();
   pp.s = pp.qd = pp.x = pp.mq = pp.ls = (T)0;
   pp.c = pp.lc = (T)1;
   if (mb2_is_1dof_descend(o0)) c.pf_issue(0, o0.cfg, o0.dof, 7);
   c.pf_commit();
   if (mb2_is_1dof_descend(o1)) c.pf_issue(1, o1.cfg, o1.dof, 7);
   c.pf_commit();
   if (mb2_is_1dof_descend(o2)) c.pf_issue(2, o2.cfg, o2.dof, 7);
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

// One op inside a run: the same steps as rnea_op with the kind (ASCEND / joint type / SC) fixed at compile time, so a run
// is a tight loop over one straight-line routine.
// KIND & MB_RUN_PLAIN: every op of the run has the flags of the common case of its kind (mb_run_plain_flags: the interior body of
// a chain, or the leaf that ends it), so they are compile-time constants here and their tests -- a uniform-datapath test and a
// branch each, some 15 cycles of latency apiece at the head of the op -- fold away.  The SC bit implies that the next op is a
// one-DoF DESCEND (MB2_PF_NEXT1) in every kind.
template <class T, class Ctx, bool FEXT, int KIND>
MB_HD void rnea_run_step(const MbProgram &P, Ctx &c, const int k, const T *grav, SvT<T> &v, SvT<T> &a, SvT<T> &f, RneaPipe<T> &pp)
{
   constexpr bool ASC = (KIND & MB2_ASCEND) != 0, SC = (KIND & MB2_SC) != 0, PLAIN = (KIND & MB_RUN_PLAIN) != 0;
   constexpr bool DYN_SC = !MB_RNEA_SC_SPLIT; // SC tested at run time instead of being part of the run kind
   constexpr int JT = (KIND >> 1) & 3;
   MbOp2 o = P.op2[k];
   if (!DYN_SC && SC)
      o.pf |= MB2_PF_NEXT1;
   if (PLAIN)
   {
      o.flags = (uint8_t)mb_run_plain_flags(MB_RNEA, KIND & 0xf, false);
      if (!DYN_SC && !SC)
         o.pf &= (uint8_t)~MB2_PF_NEXT1;
   }
   rnea_pre<T, Ctx>(c, k, o, !ASC && JT != MB_SIXDOF, !ASC, !ASC && JT == MB_REVOLUTE, grav, v, a, pp);
   const int ext = FEXT ? P.body[o.body].ext_index : 0;
   T ns = pp.mq, nc = (T)1;
   if (DYN_SC)
   {
      if (o.code & MB2_SC)
         mb_sincos(pp.mq, &ns, &nc);
   }
   if (JT == MB_SIXDOF)
   {
      if (SC) mb_sincos(pp.mq, &ns, &nc);
      if (ASC) rnea_ascend_6dof<T, Ctx, FEXT>(c, o, ext, f);
      else rnea_descend_6dof<T, Ctx, FEXT>(c, o, ext, v, a, f);
   }
   else if (ASC)
      rnea_ascend_1dof<T, Ctx, FEXT, JT == MB_REVOLUTE, SC>(c, o, ext, f, pp, ns, nc);
   else
      rnea_descend_1dof<T, Ctx, FEXT, JT == MB_REVOLUTE, SC>(c, o, ext, v, a, f, pp, ns, nc);
   pp.s = ns;
   pp.c = nc;
}

template <class T, class Ctx, bool FEXT> MB_HD void rnea_state(const MbProgram &P, Ctx &c, const T *grav)
{
   SvT<T> v, a, f;
   RneaPipe<T> pp;
   rnea_begin<T, Ctx>(c, P.op2[0], P.op2[1], P.op2[2], v, a, f, pp);
   const int nruns = P.nruns;
#pragma unroll 1
   for (int r = 0; r < nruns; r++)
   {
      const MbRun R = P.run[r];
      int k = R.k0;
      const int k1 = k + R.n;
#define MB_RUN_CASE(KIND)                                                                                  \
   case KIND:                                                                                               \
      _Pragma("unroll 1") do { rnea_run_step<T, Ctx, FEXT, KIND>(P, c, k, grav, v, a, f, pp); } while (++k < k1); \
      break;
#define MB_RUN_CASE_IF(COND, KIND)                                                                          \
   case KIND:                                                                                               \
      if constexpr ((COND) != 0)                                                                            \
         _Pragma("unroll 1") do { rnea_run_step<T, Ctx, FEXT, KIND>(P, c, k, grav, v, a, f, pp); } while (++k < k1); \
      break;
      switch (R.kind)
      {
         MB_RUN_CASE(0) MB_RUN_CASE(1) MB_RUN_CASE(2) MB_RUN_CASE(3) MB_RUN_CASE(4) MB_RUN_CASE(5)
         MB_RUN_CASE(8) MB_RUN_CASE(9) MB_RUN_CASE(10) MB_RUN_CASE(11) MB_RUN_CASE(12) MB_RUN_CASE(13)
         MB_RUN_CASE_IF(MB_PLAIN_RNEA & 1, MB_RUN_PLAIN | 0) MB_RUN_CASE_IF(MB_PLAIN_RNEA & 2, MB_RUN_PLAIN | 1) MB_RUN_CASE_IF(MB_PLAIN_RNEA & 4, MB_RUN_PLAIN | 8)
         default: break;
      }
#undef MB_RUN_CASE
#undef MB_RUN_CASE_IF
   }
   c.template pf_wait<0>();
}
} // namespace mb
