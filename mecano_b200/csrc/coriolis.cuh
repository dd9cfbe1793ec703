// coriolis.cuh -- one state of the Coriolis and centrifugal matrix C(q, qd) together with the mass matrix
// (CompositeRigidBodyMassMatrixCalculator with setEnableCoriolisMatrixCalculation(true),
//  M/algorithms/CompositeRigidBodyMassMatrixCalculator.java:278-281, getCoriolisMatrix :358-366, computeMassMatrix :588-799;
//  the factorized body inertia of M/algorithms/FactorizedBodyInertia.java:149-175, :201-247, :295-312).
//
// Same depth-first interleaving as crba.cuh: DESCEND = joint transform and twist of the body, ASCEND = composite inertia
// and composite factorized inertia of the subtree, the three force columns per DoF
//     F1 = Ic Sdot + Bc S,   F2 = Ic S,   F3 = Bc^T S        (:685-693)
// and the walk over the ancestors (:730-766).  All frames are the kernels' canonical joint frames (axis = +z), in which the
// projections onto S and Sdot = v x S pick single components.  Per level of the tree the stack holds the twist and sin/cos
// (SixDoF: the joint transform); the branch save area holds the two accumulators (10 + 36 doubles).
#pragma once
#include "crba.cuh"

namespace mb
{
// four general 3x3 blocks of a 6x6 matrix [[A, TR], [BL, L]] acting on (angular, linear)
template <class T> struct FbiT { M3T<T> A, TR, BL, L; };

template <class T> MB_HD M3T<T> m3_zero()
{
   M3T<T> r;
   r.xx = r.xy = r.xz = r.yx = r.yy = r.yz = r.zx = r.zy = r.zz = (T)0;
   return r;
}
template <class T> MB_HD M3T<T> m3_add(const M3T<T> &a, const M3T<T> &b)
{
   M3T<T> r;
   r.xx = a.xx + b.xx; r.xy = a.xy + b.xy; r.xz = a.xz + b.xz;
   r.yx = a.yx + b.yx; r.yy = a.yy + b.yy; r.yz = a.yz + b.yz;
   r.zx = a.zx + b.zx; r.zy = a.zy + b.zy; r.zz = a.zz + b.zz;
   return r;
}
template <class T> MB_HD M3T<T> m3_sub(const M3T<T> &a, const M3T<T> &b)
{
   M3T<T> r;
   r.xx = a.xx - b.xx; r.xy = a.xy - b.xy; r.xz = a.xz - b.xz;
   r.yx = a.yx - b.yx; r.yy = a.yy - b.yy; r.yz = a.yz - b.yz;
   r.zx = a.zx - b.zx; r.zy = a.zy - b.zy; r.zz = a.zz - b.zz;
   return r;
}
// p~ M: row i of the result = (p x column j of M)_i, i.e. every column crossed from the left
template <class T> MB_HD M3T<T> tilde_mul(const V3T<T> &p, const M3T<T> &m)
{
   M3T<T> r;
   r.xx = p.y * m.zx - p.z * m.yx; r.xy = p.y * m.zy - p.z * m.yy; r.xz = p.y * m.zz - p.z * m.yz;
   r.yx = p.z * m.xx - p.x * m.zx; r.yy = p.z * m.xy - p.x * m.zy; r.yz = p.z * m.xz - p.x * m.zz;
   r.zx = p.x * m.yx - p.y * m.xx; r.zy = p.x * m.yy - p.y * m.xy; r.zz = p.x * m.yz - p.y * m.xz;
   return r;
}
// M p~
template <class T> MB_HD M3T<T> mul_tilde(const M3T<T> &m, const V3T<T> &p)
{
   M3T<T> r;
   r.xx = m.xy * p.z - m.xz * p.y; r.xy = m.xz * p.x - m.xx * p.z; r.xz = m.xx * p.y - m.xy * p.x;
   r.yx = m.yy * p.z - m.yz * p.y; r.yy = m.yz * p.x - m.yx * p.z; r.yz = m.yx * p.y - m.yy * p.x;
   r.zx = m.zy * p.z - m.zz * p.y; r.zy = m.zz * p.x - m.zx * p.z; r.zz = m.zx * p.y - m.zy * p.x;
   return r;
}
// a~ b~ = b a^T - (a . b) 1
template <class T> MB_HD M3T<T> tilde_tilde(const V3T<T> &a, const V3T<T> &b)
{
   const T d = dot(a, b);
   M3T<T> r;
   r.xx = b.x * a.x - d; r.xy = b.x * a.y; r.xz = b.x * a.z;
   r.yx = b.y * a.x; r.yy = b.y * a.y - d; r.yz = b.y * a.z;
   r.zx = b.z * a.x; r.zy = b.z * a.y; r.zz = b.z * a.z - d;
   return r;
}
template <class T> MB_HD M3T<T> m3_from_sym(const S3T<T> &s)
{
   M3T<T> r;
   r.xx = s.xx; r.xy = s.xy; r.xz = s.xz; r.yx = s.xy; r.yy = s.yy; r.yz = s.yz; r.zx = s.xz; r.zy = s.yz; r.zz = s.zz;
   return r;
}

// FactorizedBodyInertia.setIncludingFrame(spatialInertia, bodyTwist) (:149-175) for a body given by its inertia about the
// frame origin, first moment h = m c and mass:  A = w~ J - v~ h~,  BL = -w~ h~,  TR = m v~ - BL,  L = m w~
template <class T> MB_HD FbiT<T> fbi_from_rbi(const RbiT<T> &I, const SvT<T> &v)
{
   FbiT<T> b;
   const M3T<T> wh = tilde_tilde(v.a, I.h);
   b.A = m3_sub(tilde_mul(v.a, m3_from_sym(I.I)), tilde_tilde(v.l, I.h));
   b.BL = m3_sub(m3_zero<T>(), wh);
   M3T<T> mv = m3_zero<T>();
   mv.xy = -I.m * v.l.z; mv.xz = I.m * v.l.y; mv.yx = I.m * v.l.z; mv.yz = -I.m * v.l.x; mv.zx = -I.m * v.l.y; mv.zy = I.m * v.l.x;
   b.TR = m3_add(mv, wh);
   b.L = m3_zero<T>();
   b.L.xy = -I.m * v.a.z; b.L.xz = I.m * v.a.y; b.L.yx = I.m * v.a.z; b.L.yz = -I.m * v.a.x; b.L.zx = -I.m * v.a.y; b.L.zy = I.m * v.a.x;
   return b;
}
template <class T> MB_HD FbiT<T> operator+(const FbiT<T> &a, const FbiT<T> &b)
{
   FbiT<T> r;
   r.A = m3_add(a.A, b.A); r.TR = m3_add(a.TR, b.TR); r.BL = m3_add(a.BL, b.BL); r.L = m3_add(a.L, b.L);
   return r;
}
// FactorizedBodyInertia.applyTransform (:295-312): rotate the four blocks, then with t = X.p, in this order:
//   TR += t~ L,  A += t~ BL,  A -= TR t~,  BL -= L t~
template <class T> MB_HD FbiT<T> fbi_to_parent(const XfT<T> &X, const FbiT<T> &b)
{
   FbiT<T> r;
   r.A = rot_gen(X.R, b.A); r.L = rot_gen(X.R, b.L); r.TR = rot_gen(X.R, b.TR); r.BL = rot_gen(X.R, b.BL);
   r.TR = m3_add(r.TR, tilde_mul(X.p, r.L));
   r.A = m3_add(r.A, tilde_mul(X.p, r.BL));
   r.A = m3_sub(r.A, mul_tilde(r.TR, X.p));
   r.BL = m3_sub(r.BL, mul_tilde(r.L, X.p));
   return r;
}
template <class T> MB_HD SvT<T> mul(const FbiT<T> &b, const SvT<T> &x) // transform (:201-213)
{
   SvT<T> r;
   r.a = mul(b.A, x.a) + mul(b.TR, x.l);
   r.l = mul(b.BL, x.a) + mul(b.L, x.l);
   return r;
}
template <class T> MB_HD SvT<T> mulT(const FbiT<T> &b, const SvT<T> &x) // transposeTransform (:230-247)
{
   SvT<T> r;
   r.a = mulT(b.A, x.a) + mulT(b.BL, x.l);
   r.l = mulT(b.TR, x.a) + mulT(b.L, x.l);
   return r;
}

template <class T, class Ctx> MB_HD void aux_st_m3(Ctx &c, int i, const M3T<T> &m)
{
   c.aux_st(i + 0, m.xx); c.aux_st(i + 1, m.xy); c.aux_st(i + 2, m.xz); c.aux_st(i + 3, m.yx); c.aux_st(i + 4, m.yy); c.aux_st(i + 5, m.yz);
   c.aux_st(i + 6, m.zx); c.aux_st(i + 7, m.zy); c.aux_st(i + 8, m.zz);
}
template <class T, class Ctx> MB_HD M3T<T> aux_ld_m3(Ctx &c, int i)
{
   M3T<T> m;
   m.xx = c.aux_ld(i + 0); m.xy = c.aux_ld(i + 1); m.xz = c.aux_ld(i + 2); m.yx = c.aux_ld(i + 3); m.yy = c.aux_ld(i + 4); m.yz = c.aux_ld(i + 5);
   m.zx = c.aux_ld(i + 6); m.zy = c.aux_ld(i + 7); m.zz = c.aux_ld(i + 8);
   return m;
}
template <class T, class Ctx> MB_HD void aux_st_fbi(Ctx &c, int i, const FbiT<T> &b)
{
   aux_st_m3<T>(c, i, b.A); aux_st_m3<T>(c, i + 9, b.TR); aux_st_m3<T>(c, i + 18, b.BL); aux_st_m3<T>(c, i + 27, b.L);
}
template <class T, class Ctx> MB_HD FbiT<T> aux_ld_fbi(Ctx &c, int i)
{
   FbiT<T> b;
   b.A = aux_ld_m3<T>(c, i); b.TR = aux_ld_m3<T>(c, i + 9); b.BL = aux_ld_m3<T>(c, i + 18); b.L = aux_ld_m3<T>(c, i + 27);
   return b;
}

// stack slot of a non-leaf body (double2 units): [twist: 3 | sin/cos: 1] or, SixDoF, [twist: 3 | joint transform: 6]
#define MB_COR_TWIST 0
#define MB_COR_JP 3
template <class T, class Ctx> MB_HD void cor_st_twist(Ctx &c, int slot2, const SvT<T> &v)
{
   c.stk_st2(slot2, 0, v.a.x, v.a.y); c.stk_st2(slot2, 1, v.a.z, v.l.x); c.stk_st2(slot2, 2, v.l.y, v.l.z);
}
template <class T, class Ctx> MB_HD SvT<T> cor_ld_twist(Ctx &c, int slot2)
{
   SvT<T> v;
   c.stk_ld2(slot2, 0, v.a.x, v.a.y); c.stk_ld2(slot2, 1, v.a.z, v.l.x); c.stk_ld2(slot2, 2, v.l.y, v.l.z);
   return v;
}

// entries of one column `col` (a DoF of a descendant, force columns F1, F2, F3 expressed in the frame of body `jt`, twist v)
// against the DoFs of that body: C[i, col] = S_i . F1, C[col, i] = Sdot_i . F2 + S_i . F3, M[i, col] = M[col, i] = S_i . F2
// (:745-760), with Sdot_i = v x S_i (:604-630) so that Sdot_i . F2 = -(v x* F2)_i
// element i of a six-array without indexing it at run time (the arrays stay in registers)
template <class T> MB_HD T pick6(const T f[6], int i)
{
   return i == 0 ? f[0] : (i == 1 ? f[1] : (i == 2 ? f[2] : (i == 3 ? f[3] : (i == 4 ? f[4] : f[5]))));
}

// sub: MB_SUB_* of a multi-DoF joint (DoF r = component mb_sub_component(sub, r) of the spatial vectors, multidof.cuh)
template <class T, class Ctx>
MB_HD void cor_project(Ctx &c, int jt, int sub, int di, int col, const SvT<T> &v, const SvT<T> &F1, const SvT<T> &F2, const SvT<T> &F3)
{
   const int nv = c.n_dofs();
   if (jt == MB_REVOLUTE)
   {
      c.st_C(di * nv + col, F1.a.z);
      c.st_C(col * nv + di, v.a.y * F2.a.x - v.a.x * F2.a.y + v.l.y * F2.l.x - v.l.x * F2.l.y + F3.a.z);
      c.st_M(di * nv + col, F2.a.z);
      c.st_M(col * nv + di, F2.a.z);
   }
   else if (jt == MB_PRISMATIC)
   {
      c.st_C(di * nv + col, F1.l.z);
      c.st_C(col * nv + di, v.a.y * F2.l.x - v.a.x * F2.l.y + F3.l.z);
      c.st_M(di * nv + col, F2.l.z);
      c.st_M(col * nv + di, F2.l.z);
   }
   else
   {
      const SvT<T> d = cross_force(v, F2);
      const T f1[6] = {F1.a.x, F1.a.y, F1.a.z, F1.l.x, F1.l.y, F1.l.z};
      const T f2[6] = {F2.a.x, F2.a.y, F2.a.z, F2.l.x, F2.l.y, F2.l.z};
      const T g[6] = {F3.a.x - d.a.x, F3.a.y - d.a.y, F3.a.z - d.a.z, F3.l.x - d.l.x, F3.l.y - d.l.y, F3.l.z - d.l.z};
      if (sub == MB_SUB_SIX)
      {
#pragma unroll
         for (int r = 0; r < 6; r++)
         {
            c.st_C((di + r) * nv + col, f1[r]);
            c.st_C(col * nv + di + r, g[r]);
            c.st_M((di + r) * nv + col, f2[r]);
            c.st_M(col * nv + di + r, f2[r]);
         }
      }
      else
      {
#pragma unroll
         for (int r = 0; r < 3; r++)
         {
            const int cr = mb_sub_component(sub, r);
            c.st_C((di + r) * nv + col, pick6(f1, cr));
            c.st_C(col * nv + di + r, pick6(g, cr));
            c.st_M((di + r) * nv + col, pick6(f2, cr));
            c.st_M(col * nv + di + r, pick6(f2, cr));
         }
      }
   }
}

// :730-766: carry the three force columns of DoF `col` of body b up to the child of the root body
template <class T, class Ctx>
MB_HD void cor_walk(const MbProgram &P, Ctx &c, int b, int col, T s, T cs, SvT<T> F1, SvT<T> F2, SvT<T> F3)
{
   MbWalk w = P.walk[b];
   while (!(w.flags & 1u))
   {
      const auto C = c.cst(b);
      if (w.jtype == MB_REVOLUTE)
      {
         F1 = force_up_1dof<T, true>(C, s, cs, F1); F2 = force_up_1dof<T, true>(C, s, cs, F2); F3 = force_up_1dof<T, true>(C, s, cs, F3);
      }
      else if (w.jtype == MB_PRISMATIC)
      {
         F1 = force_up_1dof<T, false>(C, s, cs, F1); F2 = force_up_1dof<T, false>(C, s, cs, F2); F3 = force_up_1dof<T, false>(C, s, cs, F3);
      }
      else
      {
         const XfT<T> X = stk_ld_xf<T>(c, w.slot + MB_COR_JP);
         F1 = force_to_parent(X, F1); F2 = force_to_parent(X, F2); F3 = force_to_parent(X, F3);
      }
      b = w.parent;
      w = P.walk[b];
      cor_project<T>(c, w.jtype, mb_sub_of<Ctx>(w), w.dof, col, cor_ld_twist<T>(c, w.slot), F1, F2, F3);
      if (w.jtype != MB_SIXDOF)
         c.stk_ld2(w.slot, MB_COR_JP, s, cs);
   }
}

template <class T, class Ctx> MB_HD void coriolis_state(const MbProgram &P, Ctx &c)
{
   // entries coupling joints of unrelated branches (massMatrix.zero(), coriolisMatrix.zero(), :296-299), spread over the ops
   const int zpart = c.zero_parts(P.nops);
   RbiT<T> acc = RbiT<T>();
   FbiT<T> bacc;
   bacc.A = bacc.TR = bacc.BL = bacc.L = m3_zero<T>();
   SvT<T> v = sv_zero<T>();
   T ls = (T)0, lc = (T)1;
   const int nops = P.nops;
   const int nv = c.n_dofs();
#pragma unroll 1
   for (int k = 0; k < nops; k++)
   {
      const MbOp2 o = P.op2[k];
      c.stk_fence();
      c.zero_fill_mc_part(k, zpart);
      const int jt = MB2_JT(o.code);
      const auto C = c.cst(o.body);
      if (!(o.code & MB2_ASCEND))
      {
         // ---- joint transform and twist of frameAfterJoint (MovingReferenceFrame.java:279-311)
         if (o.flags & MB2_ROOT_PARENT)
            v = sv_zero<T>();
         else if (o.flags & MB2_LOAD_PARENT)
            v = cor_ld_twist<T>(c, o.pslot);
         if (jt == MB_SIXDOF)
         {
            const XfT<T> X = joint_xf_multi<T>(c, C, o.cfg, mb_sub_of<Ctx>(o));
            v = motion_to_child(X, v) + ld_svj<T>(o.dof, mb_sub_of<Ctx>(o), [&](int r) { return c.ld_qd(r); });
            stk_st_xf<T>(c, o.slot + MB_COR_JP, X);
            cor_st_twist<T>(c, o.slot, v);
         }
         else
         {
            const T q = c.ld_q(o.cfg), qd = c.ld_qd(o.dof);
            XfT<T> X;
            if (jt == MB_REVOLUTE)
            {
               mb_sincos(mb_reduce_angle(q), &ls, &lc);
               X = joint_xf_1dof<T, true>(C, ls, lc);
               v = motion_to_child(X, v);
               v.a.z += qd;
            }
            else
            {
               ls = q;
               lc = (T)1;
               X = joint_xf_1dof<T, false>(C, ls, lc);
               v = motion_to_child(X, v);
               v.l.z += qd;
            }
            if (!(o.flags & MB2_LEAF))
            {
               cor_st_twist<T>(c, o.slot, v);
               c.stk_st2(o.slot, MB_COR_JP, ls, lc);
            }
         }
      }
      else
      {
         // ---- composite inertia (:648-661) and composite factorized inertia (:671-683) of the subtree, about this joint frame
         SvT<T> vb = v;
         T js = ls, jc = lc;
         if (jt == MB_SIXDOF || !(o.flags & MB2_LEAF))
         {
            vb = cor_ld_twist<T>(c, o.slot);
            if (jt != MB_SIXDOF)
               c.stk_ld2(o.slot, MB_COR_JP, js, jc);
         }
         const RbiT<T> Ib = ld_rbi<T>(C);
         RbiT<T> Ic = Ib;
         FbiT<T> Bc = fbi_from_rbi(Ib, vb);
         if (!(o.flags & MB2_LEAF))
         {
            Ic = Ic + acc;
            Bc = Bc + bacc;
         }
         const int d = o.dof;
         // fold into the parent: childInertia.applyTransform(child.transformToParent) (:658), childFactorizedInertia (:679)
         auto fold = [&]() {
            XfT<T> X;
            if (jt == MB_REVOLUTE) X = joint_xf_1dof<T, true>(C, js, jc);
            else if (jt == MB_PRISMATIC) X = joint_xf_1dof<T, false>(C, js, jc);
            else X = stk_ld_xf<T>(c, o.slot + MB_COR_JP);
            const RbiT<T> K = rbi_to_parent(X, Ic);
            const FbiT<T> KB = fbi_to_parent(X, Bc);
            if (o.flags & MB2_FIRST_CHILD)
            {
               acc = K;
               bacc = KB;
            }
            else
            {
               acc = aux_ld_rbi<T>(c, o.paux) + K;
               bacc = aux_ld_fbi<T>(c, o.paux + 10) + KB;
            }
            if (o.flags & MB2_STORE_ACC)
            {
               aux_st_rbi<T>(c, o.paux, acc);
               aux_st_fbi<T>(c, o.paux + 10, bacc);
            }
         };
         if (jt != MB_SIXDOF)
         {
            // S = e_z (angular: revolute, linear: prismatic), Sdot = v x S (:604-630); F2 = Ic S (:663-667),
            // F1 = Ic Sdot + Bc S (:687-689), F3 = Bc^T S (:692)
            SvT<T> S = sv_zero<T>();
            if (jt == MB_REVOLUTE) S.a.z = (T)1;
            else S.l.z = (T)1;
            const SvT<T> F2 = mul(Ic, S);
            const SvT<T> F1 = mul(Ic, cross_motion(vb, S)) + mul(Bc, S);
            const SvT<T> F3 = mulT(Bc, S);
            c.st_M(d * nv + d, jt == MB_REVOLUTE ? F2.a.z : F2.l.z);
            c.st_C(d * nv + d, jt == MB_REVOLUTE ? F1.a.z : F1.l.z);
            if (!(o.flags & MB2_ROOT_PARENT))
            {
               // the accumulators first, so that the composite inertias are dead while the three columns walk up the ancestors
               fold();
               cor_walk<T>(P, c, o.body, d, js, jc, F1, F2, F3);
            }
         }
         else
         {
            const int sub = mb_sub_of<Ctx>(o), nd = mb_sub_ndof(sub);
#pragma unroll 1
            for (int col = 0; col < nd; col++)
            {
               const int cc = mb_sub_component(sub, col);
               SvT<T> S = sv_zero<T>();
               if (cc == 0) S.a.x = (T)1; else if (cc == 1) S.a.y = (T)1; else if (cc == 2) S.a.z = (T)1;
               else if (cc == 3) S.l.x = (T)1; else if (cc == 4) S.l.y = (T)1; else S.l.z = (T)1;
               const SvT<T> F2 = mul(Ic, S);
               const SvT<T> F1 = mul(Ic, cross_motion(vb, S)) + mul(Bc, S);
               const SvT<T> F3 = mulT(Bc, S);
               const int dc = d + col;
               // the 6 x 6 block of the joint, as the Java loops leave it (:700-725, later writes win): on and below the diagonal
               // C[a, b] = S_a . F1_b, above it C[a, b] = Sdot_b . F2_a + S_b . F3_a; this column b = col holds F*_b, i.e. the
               // entries (a >= col, col) and (col, a > col)
               const SvT<T> dd = cross_force(vb, F2);
               const T f1[6] = {F1.a.x, F1.a.y, F1.a.z, F1.l.x, F1.l.y, F1.l.z};
               const T f2[6] = {F2.a.x, F2.a.y, F2.a.z, F2.l.x, F2.l.y, F2.l.z};
               const T g[6] = {F3.a.x - dd.a.x, F3.a.y - dd.a.y, F3.a.z - dd.a.z, F3.l.x - dd.l.x, F3.l.y - dd.l.y, F3.l.z - dd.l.z};
#pragma unroll
               for (int r = 0; r < 6; r++)
               {
                  if (r >= nd)
                     continue;
                  const int cr = sub == MB_SUB_SIX ? r : mb_sub_component(sub, r);
                  c.st_M((d + r) * nv + dc, pick6(f2, cr));
                  if (r >= col)
                     c.st_C((d + r) * nv + dc, pick6(f1, cr));
                  if (r > col)
                     c.st_C(dc * nv + d + r, pick6(g, cr));
               }
               if (!(o.flags & MB2_ROOT_PARENT))
                  cor_walk<T>(P, c, o.body, dc, js, jc, F1, F2, F3);
            }
            if (!(o.flags & MB2_ROOT_PARENT))
               fold();
         }
      }
   }
}
} // namespace mb
