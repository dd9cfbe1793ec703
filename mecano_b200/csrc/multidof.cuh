// multidof.cuh -- the multi-DoF joints beside SixDoFJoint: SphericalJoint and PlanarJoint
// (M/multiBodySystem/SphericalJoint.java, PlanarJoint.java; interfaces/SphericalJointReadOnly.java:31-71,
//  PlanarJointReadOnly.java:20-58; motion subspaces M/tools/MecanoTools.java:880-952).
//
// The kernels keep one class for joints with more than one DoF (MB_SIXDOF: joint transform on the stack, rows read straight from
// global memory) and tell the three apart by sub-type (MB_SUB_*).  In frameAfterJoint the motion subspace of all three is a
// SELECTION of components of the spatial vector [wx wy wz vx vy vz]:
//    SixDoF     all six                      configuration [qx qy qz qs x y z]
//    Spherical  (wx, wy, wz) = 0, 1, 2       configuration [qx qy qz qs]          R = R(quat), p = 0
//    Planar     (wy, vx, vz) = 1, 3, 5       configuration [pitch x z]            R = Ry(pitch), p = (x, 0, z)
// so S q_dot is an embedding, S^T f a pick, U = I^A S three columns of the articulated inertia and D = S^T I^A S a principal
// 3 x 3 block.  Forward dynamics of a three-DoF joint (ForwardDynamicsCalculator.java:1176-1235 with the 2..5-DoF inverse of
// :1189-1192) is written out here once, for both sub-types, over a dense 6 x 6 view of the inertia: these joints are rare next to
// the one-DoF joints the hot loops are built around.
#pragma once
// (included at the end of jointmath.cuh)

namespace mb
{
// sub-type of a multi-DoF op.  Ctx::kM3 says whether this instantiation handles three-DoF joints at all: the kernels for trees
// without them (the common case, humanoids included) are compiled with kM3 = false, every test below folds away and they
// are the kernels they were before these joints existed (with the tests in: H37 RNEA +3 %, ABA +3 %, profiles/r04e_variants.md)
template <class Ctx> MB_HD int mb_sub_of(const MbOp2 &o) { return Ctx::kM3 ? (int)MB2_SUB(o.code) : MB_SUB_SIX; }
template <class Ctx> MB_HD int mb_sub_of(const MbWalk &w) { return Ctx::kM3 ? (int)w.sub : MB_SUB_SIX; }

// ---- joint transform composed with the fixed offset (MecanoFactories.java:111-117 -> joint.getJointConfiguration(transform))
template <class T, class Ctx, class CP> MB_HD XfT<T> joint_xf_multi(Ctx &c, const CP C, int r, int sub)
{
   if (sub == MB_SUB_SIX)
      return joint_xf_6dof<T>(c, C, r);
   XfT<T> X;
   M3T<T> R0;
   V3T<T> p0;
   ld_xf0<T>(C, R0, p0);
   if (sub == MB_SUB_SPHERICAL)
   {
      // SphericalJointReadOnly.java:31-35: setRotationAndZeroTranslation(jointOrientation)
      X.R = mul(R0, quat_to_rot<T, Ctx::kFastQuat>(c.ld_q(r), c.ld_q(r + 1), c.ld_q(r + 2), c.ld_q(r + 3)));
      X.p = p0;
   }
   else
   {
      // planar pose (MecanoFactories.newPlanarPose3DBasics): rotation about y by the pitch, translation in the x-z plane
      T s, cs;
      mb_sincos(mb_reduce_angle(c.ld_q(r)), &s, &cs);
      M3T<T> Ry;
      Ry.xx = cs; Ry.xy = (T)0; Ry.xz = s;
      Ry.yx = (T)0; Ry.yy = (T)1; Ry.yz = (T)0;
      Ry.zx = -s; Ry.zy = (T)0; Ry.zz = cs;
      X.R = mul(R0, Ry);
      X.p = p0 + mul(R0, v3<T>(c.ld_q(r + 1), (T)0, c.ld_q(r + 2)));
   }
   return X;
}

// ---- S x: the joint's velocity-like rows embedded in a spatial vector (ld(row) reads one row)
template <class T, class F> MB_HD SvT<T> ld_svj(int row, int sub, F ld)
{
   if (sub == MB_SUB_SIX)
      return ld_sv6<T>(row, ld);
   SvT<T> r = sv_zero<T>();
   if (sub == MB_SUB_SPHERICAL)
      r.a = v3<T>(ld(row), ld(row + 1), ld(row + 2));
   else
   {
      r.a.y = ld(row);
      r.l.x = ld(row + 1);
      r.l.z = ld(row + 2);
   }
   return r;
}
// ---- S^T f into the joint's rows (st(row, value) writes one row)
template <class T, class F> MB_HD void st_svj(int row, int sub, const SvT<T> &f, F st)
{
   if (sub == MB_SUB_SIX)
   {
      st(row + 0, f.a.x); st(row + 1, f.a.y); st(row + 2, f.a.z);
      st(row + 3, f.l.x); st(row + 4, f.l.y); st(row + 5, f.l.z);
   }
   else if (sub == MB_SUB_SPHERICAL)
   {
      st(row + 0, f.a.x); st(row + 1, f.a.y); st(row + 2, f.a.z);
   }
   else
   {
      st(row + 0, f.a.y); st(row + 1, f.l.x); st(row + 2, f.l.z);
   }
}

// component i of a spatial vector, i in [wx wy wz vx vy vz] order (compile-time i after unrolling)
template <class T> MB_HD T sv_get(const SvT<T> &v, int i)
{
   return i == 0 ? v.a.x : (i == 1 ? v.a.y : (i == 2 ? v.a.z : (i == 3 ? v.l.x : (i == 4 ? v.l.y : v.l.z))));
}
// DoF k of a multi-DoF joint -> component of the spatial vector
MB_HD int mb_sub_component(int sub, int k) { return sub == MB_SUB_PLANAR ? 2 * k + 1 : k; }

// ---- dense 6 x 6 view of an articulated inertia [[A, C], [C^T, L]]
template <class T> MB_HD void abi_to_dense(const AbiT<T> &I, T m[6][6])
{
   m[0][0] = I.A.xx; m[0][1] = I.A.xy; m[0][2] = I.A.xz; m[1][1] = I.A.yy; m[1][2] = I.A.yz; m[2][2] = I.A.zz;
   m[0][3] = I.C.xx; m[0][4] = I.C.xy; m[0][5] = I.C.xz;
   m[1][3] = I.C.yx; m[1][4] = I.C.yy; m[1][5] = I.C.yz;
   m[2][3] = I.C.zx; m[2][4] = I.C.zy; m[2][5] = I.C.zz;
   m[3][3] = I.L.xx; m[3][4] = I.L.xy; m[3][5] = I.L.xz; m[4][4] = I.L.yy; m[4][5] = I.L.yz; m[5][5] = I.L.zz;
#pragma unroll
   for (int i = 1; i < 6; i++)
#pragma unroll
      for (int j = 0; j < i; j++)
         m[i][j] = m[j][i];
}
template <class T> MB_HD AbiT<T> abi_from_dense(const T m[6][6])
{
   AbiT<T> I;
   I.A.xx = m[0][0]; I.A.xy = m[0][1]; I.A.xz = m[0][2]; I.A.yy = m[1][1]; I.A.yz = m[1][2]; I.A.zz = m[2][2];
   I.C.xx = m[0][3]; I.C.xy = m[0][4]; I.C.xz = m[0][5];
   I.C.yx = m[1][3]; I.C.yy = m[1][4]; I.C.yz = m[1][5];
   I.C.zx = m[2][3]; I.C.zy = m[2][4]; I.C.zz = m[2][5];
   I.L.xx = m[3][3]; I.L.xy = m[3][4]; I.L.xz = m[3][5]; I.L.yy = m[4][4]; I.L.yz = m[4][5]; I.L.zz = m[5][5];
   return I;
}

// inverse of a symmetric positive-definite 3 x 3 by cofactors (what UnrolledInverseFromMinor_DDRM.inv computes for the 2..5-DoF
// joints, ForwardDynamicsCalculator.java:1189-1192)
template <class T> MB_HD void spd3_inverse(const T d[3][3], T inv[3][3])
{
   const T c00 = d[1][1] * d[2][2] - d[1][2] * d[1][2];
   const T c01 = d[0][2] * d[1][2] - d[0][1] * d[2][2];
   const T c02 = d[0][1] * d[1][2] - d[0][2] * d[1][1];
   const T c11 = d[0][0] * d[2][2] - d[0][2] * d[0][2];
   const T c12 = d[0][1] * d[0][2] - d[0][0] * d[1][2];
   const T c22 = d[0][0] * d[1][1] - d[0][1] * d[0][1];
   const T det = d[0][0] * c00 + d[0][1] * c01 + d[0][2] * c02;
   const T r = (T)1 / det;
   inv[0][0] = c00 * r; inv[0][1] = c01 * r; inv[0][2] = c02 * r;
   inv[1][0] = inv[0][1]; inv[1][1] = c11 * r; inv[1][2] = c12 * r;
   inv[2][0] = inv[0][2]; inv[2][1] = inv[1][2]; inv[2][2] = c22 * r;
}
} // namespace mb
