// crba.cuh -- one state of the composite-rigid-body algorithm: the joint-space mass matrix
// (CompositeRigidBodyMassMatrixCalculator, M/algorithms/CompositeRigidBodyMassMatrixCalculator.java:
//  reset()/getMassMatrix() :286-303, :344-348; computeMassMatrix() :588-667, :700-707, :772-797).
//
// The kernel is bound by the HBM write of the dense nv x nv result, so the code is organised around the stores:
//   * the structural zeros (joints of unrelated branches; massMatrix.zero() :296) come from a list of entry indices
//     built at flatten time and are written first, fire-and-forget;
//   * every non-zero entry costs one integer multiply-add for the entry index, one for the address, one store;
//   * the only per-level state is sin/cos on the shared-memory stack; the ancestor walk re-applies
//     R0 * Rz(q) to the unit momentum without materialising the rotation matrix.
#pragma once
#include "jointmath.cuh"

namespace mb
{
template <class T, class Ctx> MB_HD void aux_st_rbi(Ctx &c, int i, const RbiT<T> &I)
{
   c.aux_st(i + 0, I.I.xx); c.aux_st(i + 1, I.I.xy); c.aux_st(i + 2, I.I.xz); c.aux_st(i + 3, I.I.yy); c.aux_st(i + 4, I.I.yz); c.aux_st(i + 5, I.I.zz);
   c.aux_st(i + 6, I.h.x); c.aux_st(i + 7, I.h.y); c.aux_st(i + 8, I.h.z); c.aux_st(i + 9, I.m);
}
template <class T, class Ctx> MB_HD RbiT<T> aux_ld_rbi(Ctx &c, int i)
{
   RbiT<T> I;
   I.I.xx = c.aux_ld(i + 0); I.I.xy = c.aux_ld(i + 1); I.I.xz = c.aux_ld(i + 2); I.I.yy = c.aux_ld(i + 3); I.I.yz = c.aux_ld(i + 4); I.I.zz = c.aux_ld(i + 5);
   I.h = v3<T>(c.aux_ld(i + 6), c.aux_ld(i + 7), c.aux_ld(i + 8));
   I.m = c.aux_ld(i + 9);
   return I;
}

// M[dof_j .. , col] = S_j^T F and the mirrored entries (setSymmetricEntry, :704-705, :790-791)
// PK (packed layout, MECANO_B200_CRBA_PACKED): one store per unique entry, at packed row pk (+ r for the DoFs of a SixDoF joint):
// the mirrored entry and the structural zeros are not materialised
// sub: MB_SUB_* of a multi-DoF joint j (its DoFs are a selection of the components of F, multidof.cuh)
template <class T, class Ctx, bool PK = false> MB_HD void crba_project(Ctx &c, int jt, int sub, int dj, int col, int pk, const SvT<T> &F)
{
   const int nv = c.n_dofs();
   if (jt == MB_SIXDOF && sub != MB_SUB_SIX)
   {
      // three-DoF joint: S^T F picks three components
      const T e0 = sub == MB_SUB_PLANAR ? F.a.y : F.a.x, e1 = sub == MB_SUB_PLANAR ? F.l.x : F.a.y, e2 = sub == MB_SUB_PLANAR ? F.l.z : F.a.z;
      if (PK)
      {
         c.st_M(pk + 0, e0); c.st_M(pk + 1, e1); c.st_M(pk + 2, e2);
      }
      else
      {
         c.st_M((dj + 0) * nv + col, e0); c.st_M(col * nv + dj + 0, e0);
         c.st_M((dj + 1) * nv + col, e1); c.st_M(col * nv + dj + 1, e1);
         c.st_M((dj + 2) * nv + col, e2); c.st_M(col * nv + dj + 2, e2);
      }
      return;
   }
   if (PK)
   {
      if (jt == MB_REVOLUTE)
         c.st_M(pk, F.a.z);
      else if (jt == MB_PRISMATIC)
         c.st_M(pk, F.l.z);
      else
      {
         c.st_M(pk + 0, F.a.x); c.st_M(pk + 1, F.a.y); c.st_M(pk + 2, F.a.z);
         c.st_M(pk + 3, F.l.x); c.st_M(pk + 4, F.l.y); c.st_M(pk + 5, F.l.z);
      }
      return;
   }
   if (jt == MB_REVOLUTE)
   {
      c.st_M(dj * nv + col, F.a.z);
      c.st_M(col * nv + dj, F.a.z);
   }
   else if (jt == MB_PRISMATIC)
   {
      c.st_M(dj * nv + col, F.l.z);
      c.st_M(col * nv + dj, F.l.z);
   }
   else
   {
      const T e[6] = {F.a.x, F.a.y, F.a.z, F.l.x, F.l.y, F.l.z};
#pragma unroll
      for (int r = 0; r < 6; r++)
      {
         c.st_M((dj + r) * nv + col, e[r]);
         c.st_M(col * nv + dj + r, e[r]);
      }
   }
}

// walk from body b (whose force column F is expressed in its own frame, transform given by (s, cs) / the stack) up to
// the root, filling the off-diagonal blocks of column `col` (:772-797)
// BY (by-products instantiation): the walk also runs for the children of the root body and ends by re-expressing the unit
// momentum in the root frame, which is column `col` of the centroidal momentum matrix (computeCentroidalMomentumMatrix(), :801-809)
// pcol (PK): first packed row of column `col`
template <class T, class Ctx, bool BY = false, bool PK = false> MB_HD void crba_walk(const MbProgram &P, Ctx &c, int b, int col, int pcol, T s, T cs, SvT<T> F)
{
   MbWalk w = P.walk[b];
   while (!(w.flags & 1u)) // until the parent is the root body
   {
      const auto C = c.cst(b);
      if (w.jtype == MB_REVOLUTE)
         F = force_up_1dof<T, true>(C, s, cs, F);
      else if (w.jtype == MB_PRISMATIC)
         F = force_up_1dof<T, false>(C, s, cs, F);
      else
         F = force_to_parent(stk_ld_xf<T>(c, w.slot), F);
      b = w.parent;
      w = P.walk[b];
      crba_project<T, Ctx, PK>(c, w.jtype, mb_sub_of<Ctx>(w), w.dof, col, pcol + w.above, F);
      if (w.jtype != MB_SIXDOF)
         c.stk_ld2(w.slot, 0, s, cs);
   }
   if (BY)
   {
      const auto C = c.cst(b);
      if (w.jtype == MB_REVOLUTE)
         F = force_up_1dof<T, true>(C, s, cs, F);
      else if (w.jtype == MB_PRISMATIC)
         F = force_up_1dof<T, false>(C, s, cs, F);
      else
         F = force_to_parent(stk_ld_xf<T>(c, w.slot), F);
      const int nv = c.n_dofs();
      c.st_cmm(0 * nv + col, F.a.x); c.st_cmm(1 * nv + col, F.a.y); c.st_cmm(2 * nv + col, F.a.z);
      c.st_cmm(3 * nv + col, F.l.x); c.st_cmm(4 * nv + col, F.l.y); c.st_cmm(5 * nv + col, F.l.z);
   }
}

template <class T, class Ctx, bool BY = false, bool PK = false> MB_HD void crba_state(const MbProgram &P, Ctx &c)
{
   static_assert(!(BY && PK), "the by-product instantiation writes the dense entry-major matrix");
   // entries coupling joints of unrelated branches are zero
   // (the zeros are spread over the ops of the traversal instead of being written in one burst up front: a steadier store stream)
   const int zpart = c.zero_parts(P.nops);
   // BY: a launch without a matrix (mecano_b200_center_of_mass) keeps the composite-inertia recursion and the com rows only
   bool full = true;
   if constexpr (BY)
      full = !c.com_only();
   RbiT<T> acc = RbiT<T>();
   T s = (T)0, cs = (T)1, ls = (T)0, lc = (T)1, mq = (T)0;
   const int nops = P.nops;
   const int nv = c.n_dofs();
#pragma unroll
   for (int k = 0; k < MB_PF_DIST; k++)
   {
      const MbOp2 o = P.op2[k];
      if (mb2_is_1dof_descend(o))
         c.pf_issue(k, o.cfg, o.dof, 1);
      c.pf_commit();
   }
   {
      const MbOp2 o0 = P.op2[0];
      if (mb2_is_1dof_descend(o0))
      {
         c.template pf_wait<0>(); // (only a one-DoF first op needs its angle now, see rnea_begin)
         const T q0 = c.pf_ld(0, 0);
         if (MB2_JT(o0.code) == MB_REVOLUTE) mb_sincos(mb_reduce_angle(q0), &s, &cs);
         else s = q0; // a prismatic displacement is not an angle
      }
   }
#pragma unroll 1
   for (int k = 0; k < nops; k++)
   {
      const MbOp2 o = P.op2[k];
      c.stk_fence();
      if (!PK && full)
         c.zero_fill_part(k, zpart);
      if (o.pf & MB2_PF_D1)
         c.pf_issue((k + MB_PF_DIST) & (MB_PF_STAGES - 1), o.pfcfg, o.pfdof, 1);
      c.pf_commit();
      c.template pf_wait<MB_PF_DIST - 1>();
      mq = (T)0;
      if (o.pf & MB2_PF_NEXT1)
         mq = c.pf_ld((k + 1) & (MB_PF_STAGES - 1), 0);
      T ns = mq, nc = (T)1;
      if (o.code & MB2_SC)
         mb_sincos(mb_reduce_angle(mq), &ns, &nc);
      const int jt = MB2_JT(o.code);
      const auto C = c.cst(o.body);
      if (!(o.code & MB2_ASCEND))
      {
         // ---- joint transform of body i (the frame update of updateFramesRecursively())
         if (jt == MB_SIXDOF)
            stk_st_xf<T>(c, o.slot, joint_xf_multi<T>(c, C, o.cfg, mb_sub_of<Ctx>(o)));
         else
         {
            ls = s;
            lc = cs;
            if (!(o.flags & MB2_LEAF))
               c.stk_st2(o.slot, 0, s, cs);
         }
      }
      else
      {
         // ---- composite inertia of the subtree, about this joint frame (:648-661)
         RbiT<T> Ic = ld_rbi<T>(C);
         if (!(o.flags & MB2_LEAF))
            Ic = Ic + acc;
         T js = ls, jc = lc;
         if (jt != MB_SIXDOF && !(o.flags & MB2_LEAF))
            c.stk_ld2(o.slot, 0, js, jc);
         const int d = o.dof;
         // packed layout: first packed row of this joint's first column, position of the joint's own DoFs within a column
         const int pcol = PK ? (int)P.walk[o.body].pcol : 0, above = PK ? (int)P.walk[o.body].above : 0;
         // unit momenta F = Ic S (:663-667), diagonal block (:700-707), ancestors (:772-797)
         // (not for a centre-of-mass-only launch: the composite inertias are all it needs)
         if (full)
         {
            if (jt == MB_REVOLUTE)
            {
               SvT<T> F;
               F.a = v3<T>(Ic.I.xz, Ic.I.yz, Ic.I.zz);
               F.l = v3<T>(-Ic.h.y, Ic.h.x, (T)0);
               c.st_M(PK ? pcol + above : d * nv + d, F.a.z);
               if (BY || !(o.flags & MB2_ROOT_PARENT))
                  crba_walk<T, Ctx, BY, PK>(P, c, o.body, d, pcol, js, jc, F);
            }
            else if (jt == MB_PRISMATIC)
            {
               SvT<T> F;
               F.a = v3<T>(Ic.h.y, -Ic.h.x, (T)0);
               F.l = v3<T>((T)0, (T)0, Ic.m);
               c.st_M(PK ? pcol + above : d * nv + d, F.l.z);
               if (BY || !(o.flags & MB2_ROOT_PARENT))
                  crba_walk<T, Ctx, BY, PK>(P, c, o.body, d, pcol, js, jc, F);
            }
            else
            {
               // multi-DoF joint: one unit momentum per DoF; DoF k is component comp(k) of the spatial vector (all six for a SixDoF
               // joint, three for a spherical / planar one, multidof.cuh)
               const int sub = mb_sub_of<Ctx>(o), nd = mb_sub_ndof(sub);
               int pc = pcol; // packed: column `col` of this joint starts at pcol + col * above + col (col + 1) / 2
   #pragma unroll 1
               for (int col = 0; col < nd; col++)
               {
                  const int cc = mb_sub_component(sub, col);
                  SvT<T> e = sv_zero<T>();
                  if (cc == 0) e.a.x = 1; else if (cc == 1) e.a.y = 1; else if (cc == 2) e.a.z = 1;
                  else if (cc == 3) e.l.x = 1; else if (cc == 4) e.l.y = 1; else e.l.z = 1;
                  const SvT<T> F = mul(Ic, e);
                  const int dc = d + col;
                  const T f6[6] = {F.a.x, F.a.y, F.a.z, F.l.x, F.l.y, F.l.z};
                  if (sub == MB_SUB_SIX)
                  {
                     if (PK)
                     {
                        // upper triangle of the diagonal block: rows 0 .. col of this column
   #pragma unroll
                        for (int r = 0; r < 6; r++)
                           if (r <= col)
                              c.st_M(pc + above + r, f6[r]);
                     }
                     else
                     {
                        c.st_M((d + 0) * nv + dc, F.a.x); c.st_M((d + 1) * nv + dc, F.a.y); c.st_M((d + 2) * nv + dc, F.a.z);
                        c.st_M((d + 3) * nv + dc, F.l.x); c.st_M((d + 4) * nv + dc, F.l.y); c.st_M((d + 5) * nv + dc, F.l.z);
                     }
                  }
                  else
                  {
                     // diagonal block of a three-DoF joint: rows = the joint's own components of F
                     const T g0 = sub == MB_SUB_PLANAR ? f6[1] : f6[0], g1 = sub == MB_SUB_PLANAR ? f6[3] : f6[1], g2 = sub == MB_SUB_PLANAR ? f6[5] : f6[2];
                     if (PK)
                     {
                        c.st_M(pc + above + 0, g0);
                        if (col >= 1) c.st_M(pc + above + 1, g1);
                        if (col >= 2) c.st_M(pc + above + 2, g2);
                     }
                     else
                     {
                        c.st_M((d + 0) * nv + dc, g0); c.st_M((d + 1) * nv + dc, g1); c.st_M((d + 2) * nv + dc, g2);
                     }
                  }
                  if (BY || !(o.flags & MB2_ROOT_PARENT))
                     crba_walk<T, Ctx, BY, PK>(P, c, o.body, dc, pc, js, jc, F);
                  pc += above + col + 1;
               }
            }
         }
         if (BY && (o.flags & MB2_ROOT_PARENT))
         {
            // first moment and mass of this subtree in the root frame, summed over the children of the root body into the
            // centre-of-mass rows (zeroed before the launch): the origin of a centre-of-mass centroidal frame
            XfT<T> X;
            if (jt == MB_REVOLUTE) X = joint_xf_1dof<T, true>(C, js, jc);
            else if (jt == MB_PRISMATIC) X = joint_xf_1dof<T, false>(C, js, jc);
            else X = stk_ld_xf<T>(c, o.slot);
            const V3T<T> hw = mul(X.R, Ic.h) + Ic.m * X.p;
            c.add_com(0, hw.x); c.add_com(1, hw.y); c.add_com(2, hw.z); c.add_com(3, Ic.m);
         }
         if (!(o.flags & MB2_ROOT_PARENT))
         {
            // childInertia.applyTransform(child.transformToParent) (:658)
            XfT<T> X;
            if (jt == MB_REVOLUTE) X = joint_xf_1dof<T, true>(C, js, jc);
            else if (jt == MB_PRISMATIC) X = joint_xf_1dof<T, false>(C, js, jc);
            else X = stk_ld_xf<T>(c, o.slot);
            const RbiT<T> K = rbi_to_parent(X, Ic);
            if (o.flags & MB2_FIRST_CHILD)
               acc = K;
            else
               acc = aux_ld_rbi<T>(c, o.paux) + K;
            if (o.flags & MB2_STORE_ACC)
               aux_st_rbi<T>(c, o.paux, acc);
         }
      }
      s = ns;
      cs = nc;
   }
   c.template pf_wait<0>();
}
} // namespace mb
