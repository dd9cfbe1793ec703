// multidof_aba.cuh -- forward dynamics of the three-DoF joints (SphericalJoint, PlanarJoint); see multidof.cuh.  Included by
// aba.cuh once aba_fold is defined.
#pragma once

namespace mb
{
// ---- ABA pass two of a three-DoF joint (:1136-1254).  PLANAR fixes the component sets at compile time:
//   J = the joint's components (0 1 2 / 1 3 5), O = the others (3 4 5 / 0 2 4)
// Record for pass three (twelve doubles = two record slots): k0 = D^-1 u (3) and the rows O of G = U D^-1 (3 x 3); the rows J of
// G are the identity.  lo half at the body's regular record slot (it travels through the pass-three ring), hi half at rec_hi.
template <class T, class Ctx, bool FEXT, bool PLANAR>
MB_HD void aba_ascend_3dof(Ctx &c, const MbOp2 o, int ext, int rec_hi, AbiT<T> &acc, SvT<T> &pacc)
{
   constexpr int J[3] = {PLANAR ? 1 : 0, PLANAR ? 3 : 1, PLANAR ? 5 : 2};
   constexpr int O[3] = {PLANAR ? 0 : 3, PLANAR ? 2 : 4, PLANAR ? 4 : 5};
   const int sub = PLANAR ? MB_SUB_PLANAR : MB_SUB_SPHERICAL;
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
   const SvT<T> vj = ld_svj<T>(o.dof, sub, [&](int rr) { return c.ld_qd(rr); });
   if (FEXT && (o.flags & MB2_ACCSRC))
   {
      // ACCELERATION_SOURCE (:1237-1253): nothing is removed from the inertia, the known S qdd enters the bias wrench; the record
      // carries the given acceleration (pass three, MB2_ACCSRC on its record)
      const SvT<T> qdd6 = ld_svj<T>(o.dof, sub, [&](int rr) { return c.ld_x2(rr); });
      c.rec_st2(r + 0, sv_get(qdd6, J[0]), sv_get(qdd6, J[1]));
      c.rec_st2(r + 1, sv_get(qdd6, J[2]), (T)0);
      c.rec_st2(r + 2, (T)0, (T)0);
      c.rec_st2(r + 3, (T)0, (T)0);
      c.rec_st2(rec_hi + 0, (T)0, (T)0); // (pass three reads the whole record before it looks at the flag)
      c.rec_st2(rec_hi + 1, (T)0, (T)0);
      c.rec_st2(rec_hi + 2, (T)0, (T)0);
      if (!(o.flags & MB2_ROOT_PARENT))
      {
         const XfT<T> X = jp_ld_xf<T>(c, o.slot, o.nslot);
         const SvT<T> pa = pA + mul(IA, cross_motion(vb, vj) + qdd6);
         aba_fold<T>(c, o, abi_to_parent<T, 0>(X, IA), force_to_parent(X, pa), acc, pacc);
      }
      return;
   }
   T m[6][6];
   abi_to_dense(IA, m);
   const SvT<T> tau6 = ld_svj<T>(o.dof, sub, [&](int rr) { return c.ld_x(rr); });
   T D[3][3], Di[3][3], u[3], k0[3], G[3][3]; // G[i][k]: row O[i] of U D^-1
#pragma unroll
   for (int a = 0; a < 3; a++)
   {
      u[a] = sv_get(tau6, J[a]) - sv_get(pA, J[a]); // u = tau - S^T p^A (:1199-1215)
#pragma unroll
      for (int b = 0; b < 3; b++)
         D[a][b] = m[J[a]][J[b]];
   }
   spd3_inverse(D, Di);
#pragma unroll
   for (int k = 0; k < 3; k++)
   {
      k0[k] = Di[k][0] * u[0] + Di[k][1] * u[1] + Di[k][2] * u[2];
#pragma unroll
      for (int i = 0; i < 3; i++)
         G[i][k] = m[O[i]][J[0]] * Di[0][k] + m[O[i]][J[1]] * Di[1][k] + m[O[i]][J[2]] * Di[2][k];
   }
   c.rec_st2(r + 0, k0[0], k0[1]);
   c.rec_st2(r + 1, k0[2], G[0][0]);
   c.rec_st2(r + 2, G[0][1], G[0][2]);
   c.rec_st2(r + 3, (T)0, (T)0); // (the whole first slot travels through the pass-three ring)
   c.rec_st2(rec_hi + 0, G[1][0], G[1][1]);
   c.rec_st2(rec_hi + 1, G[1][2], G[2][0]);
   c.rec_st2(rec_hi + 2, G[2][1], G[2][2]);
   if (!(o.flags & MB2_ROOT_PARENT))
   {
      // I^a = I^A - U D^-1 U^T (:1217-1226): zero along the joint's own components, G U^T on the others
      T Ia[6][6];
#pragma unroll
      for (int i = 0; i < 6; i++)
#pragma unroll
         for (int j = 0; j < 6; j++)
            Ia[i][j] = (T)0;
#pragma unroll
      for (int i = 0; i < 3; i++)
#pragma unroll
         for (int j = i; j < 3; j++)
         {
            const T e = m[O[i]][O[j]] - (G[i][0] * m[O[j]][J[0]] + G[i][1] * m[O[j]][J[1]] + G[i][2] * m[O[j]][J[2]]);
            Ia[O[i]][O[j]] = e;
            Ia[O[j]][O[i]] = e;
         }
      // p^a = p^A + I^a c + U D^-1 u (:1228-1235), c = v x (S qd); along the joint's own components this is tau
      const SvT<T> cc = cross_motion(vb, vj);
      T pa6[6];
#pragma unroll
      for (int a = 0; a < 3; a++)
         pa6[J[a]] = sv_get(tau6, J[a]);
#pragma unroll
      for (int i = 0; i < 3; i++)
      {
         T e = sv_get(pA, O[i]);
#pragma unroll
         for (int j = 0; j < 3; j++)
            e += Ia[O[i]][O[j]] * sv_get(cc, O[j]) + m[O[i]][J[j]] * k0[j];
         pa6[O[i]] = e;
      }
      SvT<T> pa;
      pa.a = v3<T>(pa6[0], pa6[1], pa6[2]);
      pa.l = v3<T>(pa6[3], pa6[4], pa6[5]);
      const XfT<T> X = jp_ld_xf<T>(c, o.slot, o.nslot);
      aba_fold<T>(c, o, abi_to_parent<T, 0>(X, abi_from_dense(Ia)), force_to_parent(X, pa), acc, pacc);
   }
}

// ---- ABA pass three of a three-DoF joint (:1259-1310): qdd = D^-1 (u - U^T a') = k0 - G^T a', a = a' + S qdd
template <class T, class Ctx, bool LOCKS, bool PLANAR>
MB_HD void aba_pass3_3dof(Ctx &c, const MbOp2 o, int st, int rec_hi, SvT<T> &v, SvT<T> &a)
{
   constexpr int J[3] = {PLANAR ? 1 : 0, PLANAR ? 3 : 1, PLANAR ? 5 : 2};
   constexpr int O[3] = {PLANAR ? 0 : 3, PLANAR ? 2 : 4, PLANAR ? 4 : 5};
   const int sub = PLANAR ? MB_SUB_PLANAR : MB_SUB_SPHERICAL;
   T k0[3], G[3][3];
   c.pf3_ld2(st, 1, k0[0], k0[1]);
   c.pf3_ld2(st, 2, k0[2], G[0][0]);
   c.pf3_ld2(st, 3, G[0][1], G[0][2]);
   c.rec_ld2(rec_hi + 0, G[1][0], G[1][1]);
   c.rec_ld2(rec_hi + 1, G[1][2], G[2][0]);
   c.rec_ld2(rec_hi + 2, G[2][1], G[2][2]);
   c.rec_discard(o.body * (MB_ABA_REC / 2));
   c.rec_discard(rec_hi);
   const XfT<T> X = joint_xf_multi<T>(c, c.cst(o.body), o.cfg, sub);
   const SvT<T> vj = ld_svj<T>(o.dof, sub, [&](int rr) { return c.ld_qd(rr); });
   v = motion_to_child(X, v) + vj;
   const SvT<T> a1 = motion_to_child(X, a) + cross_motion(v, vj);
   T qdd[3];
#pragma unroll
   for (int k = 0; k < 3; k++)
   {
      qdd[k] = k0[k] - (sv_get(a1, J[k]) + G[0][k] * sv_get(a1, O[0]) + G[1][k] * sv_get(a1, O[1]) + G[2][k] * sv_get(a1, O[2]));
      if (LOCKS && (o.flags & MB2_ACCSRC))
         qdd[k] = k0[k]; // the record holds the joint's given acceleration
      c.st_out(o.dof + k, qdd[k]);
   }
   SvT<T> sq = sv_zero<T>();
   if (PLANAR)
   {
      sq.a.y = qdd[0]; sq.l.x = qdd[1]; sq.l.z = qdd[2];
   }
   else
      sq.a = v3<T>(qdd[0], qdd[1], qdd[2]);
   a = a1 + sq;
   if (o.flags & MB2_SAVE_STATE)
   {
      aux_st_sv<T>(c, o.aux, v);
      aux_st_sv<T>(c, o.aux + 6, a);
   }
}
} // namespace mb
