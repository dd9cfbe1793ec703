// spatial.cuh -- 6-D spatial algebra held in registers (fp64), shared by all kernels.
// Compiles for the device (nvcc) and, for the CPU emulation harness used by the no-GPU tests, for the
// host (tests/emu).  Everything is angular-first like Mecano (SpatialVectorReadOnly.java:268-272).
//
// A transform Xf = (R, p) locates a child frame in its parent: x_parent = R x_child + p, the same
// convention as Euclid's RigidBodyTransform, so
//   motion_to_child  == FixedFrameSpatialMotionBasics.applyInverseTransform (:343-353)
//   force_to_parent  == FixedFrameSpatialForceBasics.applyTransform        (:249-259)
#pragma once

#if defined(__CUDACC__)
#define MB_HD __host__ __device__ __forceinline__
#else
#define MB_HD inline
#endif

#if !defined(__CUDACC_RTC__)
#include <math.h>
#endif

namespace mb
{
template <class T> struct V3T { T x, y, z; };
template <class T> struct M3T { T xx, xy, xz, yx, yy, yz, zx, zy, zz; }; // row-major
template <class T> struct S3T { T xx, xy, xz, yy, yz, zz; };             // symmetric
template <class T> struct XfT { M3T<T> R; V3T<T> p; };
template <class T> struct SvT { V3T<T> a, l; };                         // angular, linear
// rigid-body / composite inertia about the frame origin: [[I, h~],[h~^T, m 1]]
template <class T> struct RbiT { S3T<T> I; V3T<T> h; T m; };
// articulated-body inertia [[A, C],[C^T, L]] (ArticulatedBodyInertia.java: angular, cross, linear)
template <class T> struct AbiT { S3T<T> A; M3T<T> C; S3T<T> L; };

#define MB_T template <class T> MB_HD

// Reciprocal without the library's special-case branch (which would split the basic block of an ABA op): hardware
// seed (rcp.approx.ftz.f64, ~2^-20) + three Newton steps.  Used for well-scaled positive numbers: the joint-space inertia
// D = S^T I^A S of a 1-DoF joint and the pivots of the 6 x 6 LDL^T of a SixDoF joint (an IEEE division is some twenty instructions
// with a slow-path call, six times per solve: ABA 1.551 -> 1.538 ms, r06q).
MB_HD double mb_rcp(double x)
{
#if defined(__CUDA_ARCH__)
   double r;
   asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
   r = __fma_rn(r, __fma_rn(-x, r, 1.0), r);
   r = __fma_rn(r, __fma_rn(-x, r, 1.0), r);
   r = __fma_rn(r, __fma_rn(-x, r, 1.0), r);
   return r;
#else
   return 1.0 / x;
#endif
}
MB_HD float mb_rcp(float x) { return 1.0f / x; }


MB_T V3T<T> v3(T x, T y, T z) { V3T<T> r; r.x = x; r.y = y; r.z = z; return r; }
MB_T V3T<T> operator+(const V3T<T> &a, const V3T<T> &b) { return v3<T>(a.x + b.x, a.y + b.y, a.z + b.z); }
MB_T V3T<T> operator-(const V3T<T> &a, const V3T<T> &b) { return v3<T>(a.x - b.x, a.y - b.y, a.z - b.z); }
MB_T V3T<T> operator*(T s, const V3T<T> &a) { return v3<T>(s * a.x, s * a.y, s * a.z); }
MB_T T dot(const V3T<T> &a, const V3T<T> &b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
MB_T V3T<T> cross(const V3T<T> &a, const V3T<T> &b)
{
   return v3<T>(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
// ---- accumulate forms: the addend rides on the multiply-add chain (one FMA per product) instead of costing an extra add.
// fmad(a, b, c) = a * b + c; for double it is a real FMA on both sides so that the host emulation follows the same path.
MB_T T fmad(T a, T b, T c) { return a * b + c; }
#if defined(__CUDA_ARCH__)
template <> MB_HD double fmad<double>(double a, double b, double c) { return __fma_rn(a, b, c); }
#else
template <> inline double fmad<double>(double a, double b, double c) { return fma(a, b, c); }
#endif
// c + a x b
MB_T V3T<T> cross_add(const V3T<T> &a, const V3T<T> &b, const V3T<T> &c)
{
   return v3<T>(fmad(a.y, b.z, fmad(-a.z, b.y, c.x)), fmad(a.z, b.x, fmad(-a.x, b.z, c.y)), fmad(a.x, b.y, fmad(-a.y, b.x, c.z)));
}
// c + R v
MB_T V3T<T> mul_add(const M3T<T> &R, const V3T<T> &v, const V3T<T> &c)
{
   return v3<T>(fmad(R.xx, v.x, fmad(R.xy, v.y, fmad(R.xz, v.z, c.x))), fmad(R.yx, v.x, fmad(R.yy, v.y, fmad(R.yz, v.z, c.y))),
                fmad(R.zx, v.x, fmad(R.zy, v.y, fmad(R.zz, v.z, c.z))));
}
// c + S v for symmetric S
MB_T V3T<T> mul_add(const S3T<T> &S, const V3T<T> &v, const V3T<T> &c)
{
   return v3<T>(fmad(S.xx, v.x, fmad(S.xy, v.y, fmad(S.xz, v.z, c.x))), fmad(S.xy, v.x, fmad(S.yy, v.y, fmad(S.yz, v.z, c.y))),
                fmad(S.xz, v.x, fmad(S.yz, v.y, fmad(S.zz, v.z, c.z))));
}

MB_T V3T<T> mul(const M3T<T> &R, const V3T<T> &v)
{
   return v3<T>(R.xx * v.x + R.xy * v.y + R.xz * v.z, R.yx * v.x + R.yy * v.y + R.yz * v.z, R.zx * v.x + R.zy * v.y + R.zz * v.z);
}
MB_T V3T<T> mulT(const M3T<T> &R, const V3T<T> &v)
{
   return v3<T>(R.xx * v.x + R.yx * v.y + R.zx * v.z, R.xy * v.x + R.yy * v.y + R.zy * v.z, R.xz * v.x + R.yz * v.y + R.zz * v.z);
}
MB_T V3T<T> mul(const S3T<T> &S, const V3T<T> &v)
{
   return v3<T>(S.xx * v.x + S.xy * v.y + S.xz * v.z, S.xy * v.x + S.yy * v.y + S.yz * v.z, S.xz * v.x + S.yz * v.y + S.zz * v.z);
}
MB_T SvT<T> operator+(const SvT<T> &a, const SvT<T> &b) { SvT<T> r; r.a = a.a + b.a; r.l = a.l + b.l; return r; }
MB_T SvT<T> operator-(const SvT<T> &a, const SvT<T> &b) { SvT<T> r; r.a = a.a - b.a; r.l = a.l - b.l; return r; }
MB_T SvT<T> sv_zero() { SvT<T> r; r.a = v3<T>(0, 0, 0); r.l = v3<T>(0, 0, 0); return r; }

// M = A * B
MB_T M3T<T> mul(const M3T<T> &A, const M3T<T> &B)
{
   M3T<T> r;
   r.xx = A.xx * B.xx + A.xy * B.yx + A.xz * B.zx; r.xy = A.xx * B.xy + A.xy * B.yy + A.xz * B.zy; r.xz = A.xx * B.xz + A.xy * B.yz + A.xz * B.zz;
   r.yx = A.yx * B.xx + A.yy * B.yx + A.yz * B.zx; r.yy = A.yx * B.xy + A.yy * B.yy + A.yz * B.zy; r.yz = A.yx * B.xz + A.yy * B.yz + A.yz * B.zz;
   r.zx = A.zx * B.xx + A.zy * B.yx + A.zz * B.zx; r.zy = A.zx * B.xy + A.zy * B.yy + A.zz * B.zy; r.zz = A.zx * B.xz + A.zy * B.yz + A.zz * B.zz;
   return r;
}

// R * Rz(angle) given (s, c): only the first two columns change
MB_T M3T<T> mul_rz(const M3T<T> &R, T s, T c)
{
   M3T<T> r;
   r.xx = c * R.xx + s * R.xy; r.xy = c * R.xy - s * R.xx; r.xz = R.xz;
   r.yx = c * R.yx + s * R.yy; r.yy = c * R.yy - s * R.yx; r.yz = R.yz;
   r.zx = c * R.zx + s * R.zy; r.zy = c * R.zy - s * R.zx; r.zz = R.zz;
   return r;
}

// rotation matrix of a (not necessarily unit) quaternion (qx qy qz qs); Mecano normalises on set
// (SixDoFJointBasics.java:104-109)
// FAST: the branch-free reciprocal instead of the IEEE division (chosen per kernel: the ABA kernels gain 1 % from it, the RNEA kernel
// loses 2 % -- code placement, r06q)
template <class T, bool FAST = false> MB_HD M3T<T> quat_to_rot(T qx, T qy, T qz, T qs)
{
   T n2 = qx * qx + qy * qy + qz * qz + qs * qs;
   T k = n2 > (T)1e-28 ? (FAST ? (T)2 * mb_rcp(n2) : (T)2 / n2) : (T)0;
   M3T<T> r;
   T xx = k * qx * qx, yy = k * qy * qy, zz = k * qz * qz;
   T xy = k * qx * qy, xz = k * qx * qz, yz = k * qy * qz;
   T sx = k * qs * qx, sy = k * qs * qy, sz = k * qs * qz;
   r.xx = (T)1 - yy - zz; r.xy = xy - sz;         r.xz = xz + sy;
   r.yx = xy + sz;         r.yy = (T)1 - xx - zz; r.yz = yz - sx;
   r.zx = xz - sy;         r.zy = yz + sx;         r.zz = (T)1 - xx - yy;
   return r;
}

// ---- spatial vector transforms
MB_T SvT<T> motion_to_child(const XfT<T> &X, const SvT<T> &m)
{
   SvT<T> r;
   r.a = mulT(X.R, m.a);
   r.l = mulT(X.R, cross_add(m.a, X.p, m.l));
   return r;
}
MB_T SvT<T> force_to_parent(const XfT<T> &X, const SvT<T> &f) // f.a = moment, f.l = force
{
   SvT<T> r;
   r.l = mul(X.R, f.l);
   r.a = mul_add(X.R, f.a, cross(X.p, r.l));
   return r;
}
// v x m  (motion cross motion), crm(v) m
MB_T SvT<T> cross_motion(const SvT<T> &v, const SvT<T> &m)
{
   SvT<T> r;
   r.a = cross(v.a, m.a);
   r.l = cross_add(v.l, m.a, cross(v.a, m.l));
   return r;
}
// v x* f  (motion cross force), crf(v) f
MB_T SvT<T> cross_force(const SvT<T> &v, const SvT<T> &f)
{
   SvT<T> r;
   r.a = cross_add(v.l, f.l, cross(v.a, f.a));
   r.l = cross(v.a, f.l);
   return r;
}
// p + v x* f: the addend rides on the multiply-add chains
MB_T SvT<T> cross_force_add(const SvT<T> &v, const SvT<T> &f, const SvT<T> &p)
{
   SvT<T> r;
   r.a = cross_add(v.l, f.l, cross_add(v.a, f.a, p.a));
   r.l = cross_add(v.a, f.l, p.l);
   return r;
}
// I * m for a rigid-body inertia about the frame origin
MB_T SvT<T> mul(const RbiT<T> &I, const SvT<T> &m)
{
   SvT<T> r;
   r.a = mul_add(I.I, m.a, cross(I.h, m.l));
   r.l = cross_add(m.a, I.h, I.m * m.l);
   return r;
}
// Newton-Euler wrench of a rigid body about the frame origin, from the quantities at its centre of mass
// (SpatialInertiaReadOnly.computeDynamicWrench :229-296, MecanoTools.computeDynamicMomentFast :571 / ForceFast :728, then
// shifted from the CoM to the origin): with vc = v + w x c, ac = a + wd x c (spatial acceleration a, wd of the origin)
//   f = m (ac + w x vc),   n = J wd + w x (J w) + c x f      == I a + v x* (I v) with I about the origin, in fewer operations
MB_T SvT<T> newton_euler(const S3T<T> &J, const V3T<T> &c, T m, const SvT<T> &v, const SvT<T> &a)
{
   const V3T<T> vc = cross_add(v.a, c, v.l);
   const V3T<T> ac = cross_add(a.a, c, a.l);
   SvT<T> r;
   r.l = m * cross_add(v.a, vc, ac);
   r.a = mul_add(J, a.a, cross_add(v.a, mul(J, v.a), cross(c, r.l)));
   return r;
}
// IA * m for an articulated inertia
MB_T SvT<T> mul(const AbiT<T> &I, const SvT<T> &m)
{
   SvT<T> r;
   r.a = mul_add(I.A, m.a, mul(I.C, m.l));
   r.l = mul_add(I.L, m.l, mulT(I.C, m.a));
   return r;
}

// ---- inertia transforms (child frame -> parent frame)
// R S R^T for symmetric S
MB_T S3T<T> rot_sym(const M3T<T> &R, const S3T<T> &S)
{
   // T = R S
   T t00 = R.xx * S.xx + R.xy * S.xy + R.xz * S.xz, t01 = R.xx * S.xy + R.xy * S.yy + R.xz * S.yz, t02 = R.xx * S.xz + R.xy * S.yz + R.xz * S.zz;
   T t10 = R.yx * S.xx + R.yy * S.xy + R.yz * S.xz, t11 = R.yx * S.xy + R.yy * S.yy + R.yz * S.yz, t12 = R.yx * S.xz + R.yy * S.yz + R.yz * S.zz;
   T t20 = R.zx * S.xx + R.zy * S.xy + R.zz * S.xz, t21 = R.zx * S.xy + R.zy * S.yy + R.zz * S.yz, t22 = R.zx * S.xz + R.zy * S.yz + R.zz * S.zz;
   S3T<T> r;
   r.xx = t00 * R.xx + t01 * R.xy + t02 * R.xz;
   r.xy = t00 * R.yx + t01 * R.yy + t02 * R.yz;
   r.xz = t00 * R.zx + t01 * R.zy + t02 * R.zz;
   r.yy = t10 * R.yx + t11 * R.yy + t12 * R.yz;
   r.yz = t10 * R.zx + t11 * R.zy + t12 * R.zz;
   r.zz = t20 * R.zx + t21 * R.zy + t22 * R.zz;
   return r;
}
// R C R^T for general C
MB_T M3T<T> rot_gen(const M3T<T> &R, const M3T<T> &C)
{
   M3T<T> t = mul(R, C), r;
   r.xx = t.xx * R.xx + t.xy * R.xy + t.xz * R.xz; r.xy = t.xx * R.yx + t.xy * R.yy + t.xz * R.yz; r.xz = t.xx * R.zx + t.xy * R.zy + t.xz * R.zz;
   r.yx = t.yx * R.xx + t.yy * R.xy + t.yz * R.xz; r.yy = t.yx * R.yx + t.yy * R.yy + t.yz * R.yz; r.yz = t.yx * R.zx + t.yy * R.zy + t.yz * R.zz;
   r.zx = t.zx * R.xx + t.zy * R.xy + t.zz * R.xz; r.zy = t.zx * R.yx + t.zy * R.yy + t.zz * R.yz; r.zz = t.zx * R.zx + t.zy * R.zy + t.zz * R.zz;
   return r;
}

// The same congruences for the structured matrices left by the rank-one downdate of a 1-DoF joint whose axis is the local z
// axis (abi_downdate<Z>): the row and column of the joint's own direction are exactly zero.
// R S R^T for symmetric S with S.xz = S.yz = S.zz = 0
MB_T S3T<T> rot_sym_z0(const M3T<T> &R, const S3T<T> &S)
{
   T t00 = R.xx * S.xx + R.xy * S.xy, t01 = R.xx * S.xy + R.xy * S.yy;
   T t10 = R.yx * S.xx + R.yy * S.xy, t11 = R.yx * S.xy + R.yy * S.yy;
   T t20 = R.zx * S.xx + R.zy * S.xy, t21 = R.zx * S.xy + R.zy * S.yy;
   S3T<T> r;
   r.xx = t00 * R.xx + t01 * R.xy;
   r.xy = t00 * R.yx + t01 * R.yy;
   r.xz = t00 * R.zx + t01 * R.zy;
   r.yy = t10 * R.yx + t11 * R.yy;
   r.yz = t10 * R.zx + t11 * R.zy;
   r.zz = t20 * R.zx + t21 * R.zy;
   return r;
}
// R C R^T for C with a zero third row (C.zx = C.zy = C.zz = 0)
MB_T M3T<T> rot_gen_rowz0(const M3T<T> &R, const M3T<T> &C)
{
   M3T<T> t, r;
   t.xx = R.xx * C.xx + R.xy * C.yx; t.xy = R.xx * C.xy + R.xy * C.yy; t.xz = R.xx * C.xz + R.xy * C.yz;
   t.yx = R.yx * C.xx + R.yy * C.yx; t.yy = R.yx * C.xy + R.yy * C.yy; t.yz = R.yx * C.xz + R.yy * C.yz;
   t.zx = R.zx * C.xx + R.zy * C.yx; t.zy = R.zx * C.xy + R.zy * C.yy; t.zz = R.zx * C.xz + R.zy * C.yz;
   r.xx = t.xx * R.xx + t.xy * R.xy + t.xz * R.xz; r.xy = t.xx * R.yx + t.xy * R.yy + t.xz * R.yz; r.xz = t.xx * R.zx + t.xy * R.zy + t.xz * R.zz;
   r.yx = t.yx * R.xx + t.yy * R.xy + t.yz * R.xz; r.yy = t.yx * R.yx + t.yy * R.yy + t.yz * R.yz; r.yz = t.yx * R.zx + t.yy * R.zy + t.yz * R.zz;
   r.zx = t.zx * R.xx + t.zy * R.xy + t.zz * R.xz; r.zy = t.zx * R.yx + t.zy * R.yy + t.zz * R.yz; r.zz = t.zx * R.zx + t.zy * R.zy + t.zz * R.zz;
   return r;
}
// R C R^T for C with a zero third column (C.xz = C.yz = C.zz = 0)
MB_T M3T<T> rot_gen_colz0(const M3T<T> &R, const M3T<T> &C)
{
   M3T<T> r;
   T t00 = R.xx * C.xx + R.xy * C.yx + R.xz * C.zx, t01 = R.xx * C.xy + R.xy * C.yy + R.xz * C.zy;
   T t10 = R.yx * C.xx + R.yy * C.yx + R.yz * C.zx, t11 = R.yx * C.xy + R.yy * C.yy + R.yz * C.zy;
   T t20 = R.zx * C.xx + R.zy * C.yx + R.zz * C.zx, t21 = R.zx * C.xy + R.zy * C.yy + R.zz * C.zy;
   r.xx = t00 * R.xx + t01 * R.xy; r.xy = t00 * R.yx + t01 * R.yy; r.xz = t00 * R.zx + t01 * R.zy;
   r.yx = t10 * R.xx + t11 * R.xy; r.yy = t10 * R.yx + t11 * R.yy; r.yz = t10 * R.zx + t11 * R.zy;
   r.zx = t20 * R.xx + t21 * R.xy; r.zy = t20 * R.yx + t21 * R.yy; r.zz = t20 * R.zx + t21 * R.zy;
   return r;
}

// rigid-body inertia of a child expressed in (and about the origin of) the parent frame.
// With w = R h + (m/2) p:  I' = R I R^T + 2 (w.p) 1 - (w p^T + p w^T),  h' = R h + m p
// (same result as SpatialInertiaBasics.applyTransform :222-239 with MecanoTools.translateMomentOfInertia :483-547)
MB_T RbiT<T> rbi_to_parent(const XfT<T> &X, const RbiT<T> &I)
{
   RbiT<T> r;
   S3T<T> Ir = rot_sym(X.R, I.I);
   V3T<T> hr = mul(X.R, I.h);
   V3T<T> p = X.p;
   V3T<T> w = hr + ((T)0.5 * I.m) * p;
   T wp2 = (T)2 * dot(w, p);
   r.I.xx = Ir.xx + wp2 - (T)2 * w.x * p.x;
   r.I.yy = Ir.yy + wp2 - (T)2 * w.y * p.y;
   r.I.zz = Ir.zz + wp2 - (T)2 * w.z * p.z;
   r.I.xy = Ir.xy - (w.x * p.y + p.x * w.y);
   r.I.xz = Ir.xz - (w.x * p.z + p.x * w.z);
   r.I.yz = Ir.yz - (w.y * p.z + p.y * w.z);
   r.h = hr + I.m * p;
   r.m = I.m;
   return r;
}
MB_T RbiT<T> operator+(const RbiT<T> &a, const RbiT<T> &b)
{
   RbiT<T> r;
   r.I.xx = a.I.xx + b.I.xx; r.I.xy = a.I.xy + b.I.xy; r.I.xz = a.I.xz + b.I.xz;
   r.I.yy = a.I.yy + b.I.yy; r.I.yz = a.I.yz + b.I.yz; r.I.zz = a.I.zz + b.I.zz;
   r.h = a.h + b.h;
   r.m = a.m + b.m;
   return r;
}

MB_T AbiT<T> abi_from_rbi(const RbiT<T> &I)
{
   AbiT<T> r;
   r.A = I.I;
   r.C.xx = 0;      r.C.xy = -I.h.z; r.C.xz = I.h.y;
   r.C.yx = I.h.z;  r.C.yy = 0;      r.C.yz = -I.h.x;
   r.C.zx = -I.h.y; r.C.zy = I.h.x;  r.C.zz = 0;
   r.L.xx = I.m; r.L.yy = I.m; r.L.zz = I.m; r.L.xy = 0; r.L.xz = 0; r.L.yz = 0;
   return r;
}
// acc + abi_from_rbi(I) without the additions of the structural zeros (a compiler may not fold x + 0.0) and with the negated
// entries of h~ as subtractions: 15 additions instead of 27 and three negations
MB_T AbiT<T> abi_add_rbi(const AbiT<T> &a, const RbiT<T> &I)
{
   AbiT<T> r;
   r.A.xx = a.A.xx + I.I.xx; r.A.xy = a.A.xy + I.I.xy; r.A.xz = a.A.xz + I.I.xz; r.A.yy = a.A.yy + I.I.yy; r.A.yz = a.A.yz + I.I.yz; r.A.zz = a.A.zz + I.I.zz;
   r.C.xx = a.C.xx;           r.C.xy = a.C.xy - I.h.z; r.C.xz = a.C.xz + I.h.y;
   r.C.yx = a.C.yx + I.h.z;  r.C.yy = a.C.yy;           r.C.yz = a.C.yz - I.h.x;
   r.C.zx = a.C.zx - I.h.y;  r.C.zy = a.C.zy + I.h.x;  r.C.zz = a.C.zz;
   r.L.xx = a.L.xx + I.m; r.L.yy = a.L.yy + I.m; r.L.zz = a.L.zz + I.m; r.L.xy = a.L.xy; r.L.xz = a.L.xz; r.L.yz = a.L.yz;
   return r;
}
MB_T AbiT<T> operator+(const AbiT<T> &a, const AbiT<T> &b)
{
   AbiT<T> r;
   r.A.xx = a.A.xx + b.A.xx; r.A.xy = a.A.xy + b.A.xy; r.A.xz = a.A.xz + b.A.xz; r.A.yy = a.A.yy + b.A.yy; r.A.yz = a.A.yz + b.A.yz; r.A.zz = a.A.zz + b.A.zz;
   r.L.xx = a.L.xx + b.L.xx; r.L.xy = a.L.xy + b.L.xy; r.L.xz = a.L.xz + b.L.xz; r.L.yy = a.L.yy + b.L.yy; r.L.yz = a.L.yz + b.L.yz; r.L.zz = a.L.zz + b.L.zz;
   r.C.xx = a.C.xx + b.C.xx; r.C.xy = a.C.xy + b.C.xy; r.C.xz = a.C.xz + b.C.xz;
   r.C.yx = a.C.yx + b.C.yx; r.C.yy = a.C.yy + b.C.yy; r.C.yz = a.C.yz + b.C.yz;
   r.C.zx = a.C.zx + b.C.zx; r.C.zy = a.C.zy + b.C.zy; r.C.zz = a.C.zz + b.C.zz;
   return r;
}

// articulated inertia of a child expressed in the parent frame: the congruence X* IA X^-1
// (ArticulatedBodyInertia.applyTransform :359-375).  Rotate the three blocks, then with t = p:
//   C' = C + t~ L,   A' = A + t~ C^T + C' t~^T,   L' = L
// row_i(M t~^T) = t x row_i(M), and t~ C^T = (C t~^T)^T.
// Z = 0: general;  Z = 1 / 2: I comes out of abi_downdate<1 / 2> (revolute / prismatic joint about the local z axis)
template <class T, int Z = 0> MB_HD AbiT<T> abi_to_parent(const XfT<T> &X, const AbiT<T> &I)
{
   AbiT<T> r;
   S3T<T> A = Z == 1 ? rot_sym_z0(X.R, I.A) : rot_sym(X.R, I.A);
   S3T<T> L = Z == 2 ? rot_sym_z0(X.R, I.L) : rot_sym(X.R, I.L);
   M3T<T> C = Z == 1 ? rot_gen_rowz0(X.R, I.C) : (Z == 2 ? rot_gen_colz0(X.R, I.C) : rot_gen(X.R, I.C));
   V3T<T> t = X.p;
   // columns of C' = C + t~ L: C'[:,j] = C[:,j] + t x L[:,j]
   const V3T<T> c0 = cross_add(t, v3<T>(L.xx, L.xy, L.xz), v3<T>(C.xx, C.yx, C.zx));
   const V3T<T> c1 = cross_add(t, v3<T>(L.xy, L.yy, L.yz), v3<T>(C.xy, C.yy, C.zy));
   const V3T<T> c2 = cross_add(t, v3<T>(L.xz, L.yz, L.zz), v3<T>(C.xz, C.yz, C.zz));
   M3T<T> Cn;
   Cn.xx = c0.x; Cn.xy = c1.x; Cn.xz = c2.x;
   Cn.yx = c0.y; Cn.yy = c1.y; Cn.yz = c2.y;
   Cn.zx = c0.z; Cn.zy = c1.z; Cn.zz = c2.z;
   // A'_ij = A_ij + (t x row_i(C'))_j + (t x row_j(C))_i, every product on one multiply-add chain
   r.A.xx = fmad(t.y, Cn.xz, fmad(-t.z, Cn.xy, fmad(t.y, C.xz, fmad(-t.z, C.xy, A.xx))));
   r.A.xy = fmad(t.z, Cn.xx, fmad(-t.x, Cn.xz, fmad(t.y, C.yz, fmad(-t.z, C.yy, A.xy))));
   r.A.xz = fmad(t.x, Cn.xy, fmad(-t.y, Cn.xx, fmad(t.y, C.zz, fmad(-t.z, C.zy, A.xz))));
   r.A.yy = fmad(t.z, Cn.yx, fmad(-t.x, Cn.yz, fmad(t.z, C.yx, fmad(-t.x, C.yz, A.yy))));
   r.A.yz = fmad(t.x, Cn.yy, fmad(-t.y, Cn.yx, fmad(t.z, C.zx, fmad(-t.x, C.zz, A.yz))));
   r.A.zz = fmad(t.x, Cn.zy, fmad(-t.y, Cn.zx, fmad(t.x, C.zy, fmad(-t.y, C.zx, A.zz))));
   r.C = Cn;
   r.L = L;
   return r;
}

// IA - U U^T / D, with g = U / D  (rank-one downdate of the three blocks).  U is the column of IA along the joint's motion
// subspace, so that column (and row) of the result is zero: Z = 1 (revolute about z: U = [A(:,z); C(z,:)^T]) and Z = 2
// (prismatic along z: U = [C(:,z); L(:,z)]) write exact zeros there instead of computing the cancellation.
template <class T, int Z = 0> MB_HD AbiT<T> abi_downdate(const AbiT<T> &I, const SvT<T> &U, const SvT<T> &g)
{
   AbiT<T> r;
   r.A.xx = I.A.xx - U.a.x * g.a.x; r.A.xy = I.A.xy - U.a.x * g.a.y;
   r.A.yy = I.A.yy - U.a.y * g.a.y;
   r.L.xx = I.L.xx - U.l.x * g.l.x; r.L.xy = I.L.xy - U.l.x * g.l.y;
   r.L.yy = I.L.yy - U.l.y * g.l.y;
   r.C.xx = I.C.xx - U.a.x * g.l.x; r.C.xy = I.C.xy - U.a.x * g.l.y;
   r.C.yx = I.C.yx - U.a.y * g.l.x; r.C.yy = I.C.yy - U.a.y * g.l.y;
   if (Z == 1)
   {
      r.A.xz = r.A.yz = r.A.zz = (T)0;
      r.C.zx = r.C.zy = r.C.zz = (T)0;
   }
   else
   {
      r.A.xz = I.A.xz - U.a.x * g.a.z; r.A.yz = I.A.yz - U.a.y * g.a.z; r.A.zz = I.A.zz - U.a.z * g.a.z;
      r.C.zx = I.C.zx - U.a.z * g.l.x; r.C.zy = I.C.zy - U.a.z * g.l.y;
   }
   if (Z == 2)
   {
      r.L.xz = r.L.yz = r.L.zz = (T)0;
      r.C.xz = r.C.yz = r.C.zz = (T)0;
   }
   else
   {
      r.L.xz = I.L.xz - U.l.x * g.l.z; r.L.yz = I.L.yz - U.l.y * g.l.z; r.L.zz = I.L.zz - U.l.z * g.l.z;
      r.C.xz = I.C.xz - U.a.x * g.l.z; r.C.yz = I.C.yz - U.a.y * g.l.z;
   }
   if (Z == 0)
      r.C.zz = I.C.zz - U.a.z * g.l.z;
   return r;
}

// Solve IA x = b for a symmetric positive-definite 6x6 (LDL^T, fully unrolled => registers).
// Replaces EJML's LinearSolverFactory_DDRM.symmPosDef(6) used for SixDoF joints (ForwardDynamicsCalculator.java:1040, :1193-1197).
// (MB_SOLVE_ROLLED: the loops stay loops and the factor lives in local memory -- a SixDoF joint is solved once per state, and the
// unrolled solve is several KB of code that evicts the hot one-DoF loops from the instruction cache)
#if defined(MB_SOLVE_ROLLED) && defined(__CUDA_ARCH__)
#define MB_SOLVE_UNROLL _Pragma("unroll 1")
#else
#define MB_SOLVE_UNROLL _Pragma("unroll")
#endif
MB_T SvT<T> abi_solve(const AbiT<T> &I, const SvT<T> &b)
{
   T a[6][6];
   a[0][0] = I.A.xx; a[1][0] = I.A.xy; a[2][0] = I.A.xz; a[1][1] = I.A.yy; a[2][1] = I.A.yz; a[2][2] = I.A.zz;
   a[3][0] = I.C.xx; a[3][1] = I.C.yx; a[3][2] = I.C.zx; // C^T rows
   a[4][0] = I.C.xy; a[4][1] = I.C.yy; a[4][2] = I.C.zy;
   a[5][0] = I.C.xz; a[5][1] = I.C.yz; a[5][2] = I.C.zz;
   a[3][3] = I.L.xx; a[4][3] = I.L.xy; a[5][3] = I.L.xz; a[4][4] = I.L.yy; a[5][4] = I.L.yz; a[5][5] = I.L.zz;
   T x[6] = {b.a.x, b.a.y, b.a.z, b.l.x, b.l.y, b.l.z};
   T dinv[6], w[6];
MB_SOLVE_UNROLL
   for (int j = 0; j < 6; j++)
   {
      // after column j is done a[i][j] (i > j) holds L[i][j]; w[k] = L[j][k] * D[k]
      T d = a[j][j];
MB_SOLVE_UNROLL
      for (int k = 0; k < j; k++)
      {
         w[k] = a[j][k] * a[k][k];
         d -= a[j][k] * w[k];
      }
      a[j][j] = d; // D[j]
      dinv[j] = mb_rcp(d);
MB_SOLVE_UNROLL
      for (int i = j + 1; i < 6; i++)
      {
         T s = a[i][j];
MB_SOLVE_UNROLL
         for (int k = 0; k < j; k++)
            s -= a[i][k] * w[k];
         a[i][j] = s * dinv[j];
      }
   }
   // forward: L y = b
MB_SOLVE_UNROLL
   for (int i = 0; i < 6; i++)
MB_SOLVE_UNROLL
      for (int k = 0; k < i; k++)
         x[i] -= a[i][k] * x[k];
   // diagonal
MB_SOLVE_UNROLL
   for (int i = 0; i < 6; i++)
      x[i] *= dinv[i];
   // backward: L^T x = y
MB_SOLVE_UNROLL
   for (int i = 5; i >= 0; i--)
MB_SOLVE_UNROLL
      for (int k = i + 1; k < 6; k++)
         x[i] -= a[k][i] * x[k];
   SvT<T> r;
   r.a = v3<T>(x[0], x[1], x[2]);
   r.l = v3<T>(x[3], x[4], x[5]);
   return r;
}

#undef MB_SOLVE_UNROLL
#undef MB_T
} // namespace mb
