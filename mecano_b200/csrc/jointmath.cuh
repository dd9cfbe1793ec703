// jointmath.cuh -- helpers shared by the per-algorithm state routines (rnea.cuh, aba.cuh, crba.cuh): constant-record
// loads, joint transforms in canonical frames, stack / save-area accessors.
//
// Context policy (all methods inline; GPU: kernels.cu, host emulation: tests/emu/emu.cpp):
//   T    ld_q(row) ld_qd(row) ld_x(row)          inputs in Mecano row order (x = qdd for RNEA, tau for ABA)
//   T    ld_fext(ext_body, comp)                 external wrench rows
//   void st_out(row, T)                          tau (RNEA) / qdd (ABA)
//   void st_M(row, col, T)                       mass-matrix entry (CRBA)
//   void stk_ld2(slot2, j, T&, T&) / stk_st2     per-state stack of double2 (shared memory on the GPU)
//   void acc_ld(slot2, wslot, T& x6) / acc_st    RNEA / ABA: the first three double2 of a slot ("wide" area, MbOp2::wslot)
//   void jp_ld2(slot2, nslot, j, T&, T&) / jp_st2  RNEA / ABA: the rest of a slot ("narrow" area, MbOp2::nslot), j = 0, 1, ..
//   T    aux_ld(i) / aux_st, rec_ld / rec_st     per-state branch-save and record areas (local memory)
//   const T* cst(body)                           constant record of a body (shared memory on the GPU)
#pragma once
#include "program.h"
#include "spatial.cuh"
#if !defined(__CUDACC_RTC__)
#include <string.h>
#endif

namespace mb
{
// ---- sin/cos
// Branch-free double-precision sincos for |x| <= MB_SINCOS_FAST_LIMIT, so that it can be scheduled inside the same
// basic block as the spatial algebra of an op (the CUDA library routine carries a Payne-Hanek slow-path call that
// splits the block).  Cody-Waite reduction by pi/2 in three FMA steps, then the fdlibm kernel polynomials
// (|r| <= pi/4, error < 1 ulp).  Arguments beyond the limit are first brought into range by mb_reduce_angle(), off
// the hot path.  Java's Math.sin/cos (what Mecano calls through Euclid, MecanoFactories.java:231-260) are specified
// to 1 ulp; the difference is far below the 1e-9 parity tolerance.
#define MB_SINCOS_FAST_LIMIT 1.0e5
#if defined(__CUDA_ARCH__)
#define MB_CONST_TABLE static __constant__
#else
#define MB_CONST_TABLE static const
#endif
MB_CONST_TABLE double mb_sc_tab[18] = {
   6.36619772367581382433e-01,                                                             // 0: 2/pi
   1.57079632679489655800e+00, 6.12323399573676603587e-17, -1.49738490485916983151e-33,    // 1-3: pi/2 split (hi, mid, lo)
   -1.66666666666666324348e-01, 8.33333333332248946124e-03, -1.98412698298579493134e-04,   // 4-9: S1..S6
   2.75573137070700676789e-06, -2.50507602534068634195e-08, 1.58969099521155010221e-10,
   4.16666666666666019037e-02, -1.38888888888741095749e-03, 2.48015872894767294178e-05,    // 10-15: C1..C6
   -2.75573143513906633035e-07, 2.08757232129817482790e-09, -1.13596475577881948265e-11,
   6755399441055744.0, 0.0};                                                               // 16: 1.5 * 2^52 (round-to-integer magic)

MB_HD double mb_fma(double a, double b, double c)
{
#if defined(__CUDA_ARCH__)
   return __fma_rn(a, b, c);
#else
   return fma(a, b, c);
#endif
}

MB_HD void mb_sincos(double x, double *sn, double *cs)
{
   const double *t = mb_sc_tab;
   // k = round(x * 2/pi): the integer lands in the low mantissa bits of kd + magic
   const double km = mb_fma(x, t[0], t[16]);
   const double kd = km - t[16];
#if defined(__CUDA_ARCH__)
   const int k = __double2loint(km);
#else
   long long kb;
   memcpy(&kb, &km, sizeof kb);
   const int k = (int)(kb & 0xffffffffll);
#endif
   double r = mb_fma(kd, -t[1], x);
   r = mb_fma(kd, -t[2], r);
   r = mb_fma(kd, -t[3], r);
   const double z = r * r;
   double ps = mb_fma(z, t[9], t[8]);
   double pc = mb_fma(z, t[15], t[14]);
   ps = mb_fma(z, ps, t[7]);
   pc = mb_fma(z, pc, t[13]);
   ps = mb_fma(z, ps, t[6]);
   pc = mb_fma(z, pc, t[12]);
   ps = mb_fma(z, ps, t[5]);
   pc = mb_fma(z, pc, t[11]);
   ps = mb_fma(z, ps, t[4]);
   pc = mb_fma(z, pc, t[10]);
   const double sr = mb_fma(z * r, ps, r);                    // sin(r)
   const double cr = mb_fma(z, mb_fma(z, pc, -0.5), 1.0);     // cos(r)
   // quadrant
   const double s0 = (k & 1) ? cr : sr, c0 = (k & 1) ? sr : cr;
   *sn = (k & 2) ? -s0 : s0;
   *cs = ((k + 1) & 2) ? -c0 : c0;
}
MB_HD void mb_sincos(float x, float *s, float *c)
{
#if defined(__CUDA_ARCH__)
   sincosf(x, s, c);
#else
   *s = sinf(x);
   *c = cosf(x);
#endif
}
// Bring an angle of any magnitude into the fast range without losing accuracy.  The rare path is kept out of line so
// that it does not sit in the instruction stream of every op (the unrolled, tree-specialised kernels are bound by
// instruction fetch).
#if defined(__CUDA_ARCH__)
__device__ __noinline__ static double mb_reduce_angle_slow(double x)
{
   double s, c;
   sincos(x, &s, &c);
   return atan2(s, c);
}
#else
inline double mb_reduce_angle_slow(double x) { return atan2(sin(x), cos(x)); }
#endif
MB_HD double mb_reduce_angle(double x)
{
   if (!(fabs(x) > MB_SINCOS_FAST_LIMIT))
      return x;
   return mb_reduce_angle_slow(x);
}
MB_HD float mb_reduce_angle(float x) { return x; }

// The same safeguard off the critical path (RNEA / ABA thread-per-state kernels): the fast sincos is evaluated on the raw
// angle inside the previous op; when the op that uses it starts -- a basic-block boundary anyway -- the angle is tested with
// an integer compare on its high word (|x| >= 1e5, NaN and infinities included) and, if it was out of the fast range, sin/cos
// are redone by the library routine (exact argument reduction, like Java's Math.sin / cos).  A test in front of the sincos costs
// the latency of the shared-memory load and of a DSETP at the top of every op.
MB_HD bool mb_angle_large(double x)
{
#if defined(__CUDA_ARCH__)
   return (unsigned)(__double2hiint(x) & 0x7fffffff) >= 0x40F86A00u;
#else
   return !(fabs(x) < MB_SINCOS_FAST_LIMIT);
#endif
}
template <class T> MB_HD bool mb_angle_large(T) { return false; } // float: sincosf covers every magnitude
#if defined(__CUDA_ARCH__)
__device__ __noinline__ static double2 mb_sincos_slow2(double x) // by value: an address-taken sin/cos would live in local memory
{
   double2 r;
   sincos(x, &r.x, &r.y);
   return r;
}
MB_HD void mb_sincos_redo(double x, double &s, double &c)
{
   const double2 r = mb_sincos_slow2(x);
   s = r.x;
   c = r.y;
}
#else
inline void mb_sincos_redo(double x, double &s, double &c) { s = sin(x); c = cos(x); }
#endif
template <class T> MB_HD void mb_sincos_redo(T, T &, T &) {}

// (mb_rcp: spatial.cuh)

MB_HD bool mb2_is_1dof_descend(const MbOp2 &o) { return !(o.code & MB2_ASCEND) && MB2_JT(o.code) != MB_SIXDOF; }

template <class T> MB_HD M3T<T> ld_m3(const T *p)
{
   M3T<T> r;
   r.xx = p[0]; r.xy = p[1]; r.xz = p[2]; r.yx = p[3]; r.yy = p[4]; r.yz = p[5]; r.zx = p[6]; r.zy = p[7]; r.zz = p[8];
   return r;
}
template <class T> MB_HD V3T<T> ld_v3(const T *p) { return v3<T>(p[0], p[1], p[2]); }

// Constant records are read through a handle `C` (what Ctx::cst(body) returns) and the free function
// cst_ld2(C, i2, a, b): doubles 2 * i2 and 2 * i2 + 1 of the record.  A plain pointer is a handle (warp kernels, host
// emulation, specialised kernels); the thread-per-state GPU context returns a 32-bit shared-memory address and
// reads with ld.shared.v2.f64 (gpu_ctx.cuh).
template <class T> MB_HD void cst_ld2(const T *C, int i2, T &a, T &b)
{
   a = C[2 * i2];
   b = C[2 * i2 + 1];
}
// fixed offset of the joint: R0 (doubles 0..8), p0 (9..11)
template <class T, class CP> MB_HD void ld_xf0(const CP C, M3T<T> &R0, V3T<T> &p0)
{
   cst_ld2(C, 0, R0.xx, R0.xy); cst_ld2(C, 1, R0.xz, R0.yx); cst_ld2(C, 2, R0.yy, R0.yz);
   cst_ld2(C, 3, R0.zx, R0.zy); cst_ld2(C, 4, R0.zz, p0.x); cst_ld2(C, 5, p0.y, p0.z);
}
// inertia about the joint-frame origin: I (12..17), h (18..20), m (21)
template <class T, class CP> MB_HD RbiT<T> ld_rbi(const CP C)
{
   static_assert(MB_C_I == 12 && MB_C_H == 18 && MB_C_M == 21, "record layout");
   RbiT<T> r;
   cst_ld2(C, 6, r.I.xx, r.I.xy); cst_ld2(C, 7, r.I.xz, r.I.yy); cst_ld2(C, 8, r.I.yz, r.I.zz);
   cst_ld2(C, 9, r.h.x, r.h.y); cst_ld2(C, 10, r.h.z, r.m);
   return r;
}
// the body as Newton-Euler sees it: inertia about the CoM J (34..39), CoM position c (31..33), mass (21)
template <class T, class CP> MB_HD void ld_com_inertia(const CP C, S3T<T> &J, V3T<T> &cp, T &m)
{
   static_assert(MB_C_J == 34 && MB_C_C == 31 && MB_C_M == 21, "record layout");
   T u0, u1;
   cst_ld2(C, 17, J.xx, J.xy); cst_ld2(C, 18, J.xz, J.yy); cst_ld2(C, 19, J.yz, J.zz);
   cst_ld2(C, 15, u0, cp.x); cst_ld2(C, 16, cp.y, cp.z);
   cst_ld2(C, 10, u1, m);
}
// CoM pose in the joint frame: E (22..30), c (31..33)
template <class T, class CP> MB_HD void ld_com(const CP C, M3T<T> &E, V3T<T> &cp)
{
   static_assert(MB_C_E == 22 && MB_C_C == 31, "record layout");
   cst_ld2(C, 11, E.xx, E.xy); cst_ld2(C, 12, E.xz, E.yx); cst_ld2(C, 13, E.yy, E.yz);
   cst_ld2(C, 14, E.zx, E.zy); cst_ld2(C, 15, E.zz, cp.x); cst_ld2(C, 16, cp.y, cp.z);
}

// rotation canonical frame -> frameAfterJoint: Q (40..48)
template <class T, class CP> MB_HD M3T<T> ld_q_rot(const CP C)
{
   static_assert(MB_C_Q == 40, "record layout");
   M3T<T> Q;
   T pad;
   cst_ld2(C, 20, Q.xx, Q.xy); cst_ld2(C, 21, Q.xz, Q.yx); cst_ld2(C, 22, Q.yy, Q.yz); cst_ld2(C, 23, Q.zx, Q.zy); cst_ld2(C, 24, Q.zz, pad);
   return Q;
}
// RNEA by-products of one body (InverseDynamicsCalculator.java:578-602): its spatial acceleration re-expressed in the CoM
// frame, the wrench of its joint re-expressed in frameAfterJoint
template <class T, class Ctx, class CP> MB_HD void rnea_store_body_acc(Ctx &c, int ext, const CP C, const SvT<T> &a)
{
   XfT<T> X;
   ld_com<T>(C, X.R, X.p);
   const SvT<T> ac = motion_to_child(X, a);
   c.st_acc(ext, 0, ac.a.x); c.st_acc(ext, 1, ac.a.y); c.st_acc(ext, 2, ac.a.z);
   c.st_acc(ext, 3, ac.l.x); c.st_acc(ext, 4, ac.l.y); c.st_acc(ext, 5, ac.l.z);
}
template <class T, class Ctx, class CP> MB_HD void rnea_store_joint_wrench(Ctx &c, int ext, const CP C, const SvT<T> &f)
{
   const M3T<T> Q = ld_q_rot<T>(C);
   const V3T<T> n = mul(Q, f.a), l = mul(Q, f.l);
   c.st_wr(ext, 0, n.x); c.st_wr(ext, 1, n.y); c.st_wr(ext, 2, n.z);
   c.st_wr(ext, 3, l.x); c.st_wr(ext, 4, l.y); c.st_wr(ext, 5, l.z);
}

template <class T, class Ctx> MB_HD void rnea_add_root_wrench(Ctx &c, const SvT<T> &f)
{
   c.add_rootw(0, f.a.x); c.add_rootw(1, f.a.y); c.add_rootw(2, f.a.z);
   c.add_rootw(3, f.l.x); c.add_rootw(4, f.l.y); c.add_rootw(5, f.l.z);
}

// (a1) joint transform X_J(q) composed with the fixed offset, canonical frames (axis = +z):
// revolute (MecanoFactories.java:231-260): R = R0 Rz(q), p = p0;  prismatic (PrismaticJointReadOnly.java:18-22): R = R0, p = p0 + q R0 e_z
template <class T, bool REV, class CP> MB_HD XfT<T> joint_xf_1dof(const CP C, T s, T c)
{
   XfT<T> X;
   M3T<T> R0;
   V3T<T> p0;
   ld_xf0<T>(C, R0, p0);
   if (REV)
   {
      X.R = mul_rz(R0, s, c);
      X.p = p0;
   }
   else
   {
      X.R = R0;
      X.p = p0 + s * v3<T>(R0.xz, R0.yz, R0.zz);
   }
   return X;
}

// acc + F expressed in the parent of a 1-DoF joint: (R0 Rz(q), p0) applied to a force vector without forming the matrix, the
// parent's accumulator riding on the multiply-add chains
template <class T, bool REV, class CP> MB_HD SvT<T> force_up_1dof_add(const CP C, T s, T cs, const SvT<T> &f, const SvT<T> &acc)
{
   M3T<T> R0;
   V3T<T> p;
   ld_xf0<T>(C, R0, p);
   SvT<T> g = f, r;
   if (REV)
   {
      g.a.x = cs * f.a.x - s * f.a.y; g.a.y = s * f.a.x + cs * f.a.y;
      g.l.x = cs * f.l.x - s * f.l.y; g.l.y = s * f.l.x + cs * f.l.y;
   }
   else
      p = p + s * v3<T>(R0.xz, R0.yz, R0.zz);
   const V3T<T> l0 = mul(R0, g.l);
   // (letting the accumulator -- a TMEM load issued just before -- enter last instead, three more additions with its latency off
   // the chain, measured no faster: 0.608 vs 0.600 ms, r02o)
   r.a = mul_add(R0, g.a, cross_add(p, l0, acc.a));
   r.l = l0 + acc.l;
   return r;
}
template <class T, bool REV, class CP> MB_HD SvT<T> force_up_1dof(const CP C, T s, T cs, const SvT<T> &f)
{
   M3T<T> R0;
   V3T<T> p;
   ld_xf0<T>(C, R0, p);
   SvT<T> g = f, r;
   if (REV)
   {
      g.a.x = cs * f.a.x - s * f.a.y; g.a.y = s * f.a.x + cs * f.a.y;
      g.l.x = cs * f.l.x - s * f.l.y; g.l.y = s * f.l.x + cs * f.l.y;
   }
   else
      p = p + s * v3<T>(R0.xz, R0.yz, R0.zz);
   r.l = mul(R0, g.l);
   r.a = mul_add(R0, g.a, cross(p, r.l));
   return r;
}

// SixDoF (FloatingJointReadOnly.java:34-37): R = R0 R(quat), p = p0 + R0 pos; configuration rows [qx qy qz qs x y z]
template <class T, class Ctx, class CP> MB_HD XfT<T> joint_xf_6dof(Ctx &c, const CP C, int r)
{
   XfT<T> X;
   M3T<T> R0;
   V3T<T> p0;
   ld_xf0<T>(C, R0, p0);
   const M3T<T> Rq = quat_to_rot<T, Ctx::kFastQuat>(c.ld_q(r), c.ld_q(r + 1), c.ld_q(r + 2), c.ld_q(r + 3));
   X.R = mul(R0, Rq);
   X.p = p0 + mul(R0, v3<T>(c.ld_q(r + 4), c.ld_q(r + 5), c.ld_q(r + 6)));
   return X;
}

template <class T, class F> MB_HD SvT<T> ld_sv6(int row, F ld)
{
   SvT<T> r;
   r.a = v3<T>(ld(row), ld(row + 1), ld(row + 2));
   r.l = v3<T>(ld(row + 3), ld(row + 4), ld(row + 5));
   return r;
}

template <class T, class Ctx> MB_HD void aux_st_sv(Ctx &c, int i, const SvT<T> &v)
{
   c.aux_st(i + 0, v.a.x); c.aux_st(i + 1, v.a.y); c.aux_st(i + 2, v.a.z);
   c.aux_st(i + 3, v.l.x); c.aux_st(i + 4, v.l.y); c.aux_st(i + 5, v.l.z);
}
template <class T, class Ctx> MB_HD SvT<T> aux_ld_sv(Ctx &c, int i)
{
   SvT<T> v;
   v.a = v3<T>(c.aux_ld(i + 0), c.aux_ld(i + 1), c.aux_ld(i + 2));
   v.l = v3<T>(c.aux_ld(i + 3), c.aux_ld(i + 4), c.aux_ld(i + 5));
   return v;
}

template <class T, class Ctx> MB_HD void stk_st_sv(Ctx &c, int slot2, const SvT<T> &v)
{
   c.stk_st2(slot2, 0, v.a.x, v.a.y);
   c.stk_st2(slot2, 1, v.a.z, v.l.x);
   c.stk_st2(slot2, 2, v.l.y, v.l.z);
}
template <class T, class Ctx> MB_HD SvT<T> stk_ld_sv(Ctx &c, int slot2)
{
   SvT<T> v;
   c.stk_ld2(slot2, 0, v.a.x, v.a.y);
   c.stk_ld2(slot2, 1, v.a.z, v.l.x);
   c.stk_ld2(slot2, 2, v.l.y, v.l.z);
   return v;
}
// a whole transform on the stack: 6 double2
template <class T, class Ctx> MB_HD void stk_st_xf(Ctx &c, int slot2, const XfT<T> &X)
{
   c.stk_st2(slot2, 0, X.R.xx, X.R.xy); c.stk_st2(slot2, 1, X.R.xz, X.R.yx); c.stk_st2(slot2, 2, X.R.yy, X.R.yz);
   c.stk_st2(slot2, 3, X.R.zx, X.R.zy); c.stk_st2(slot2, 4, X.R.zz, X.p.x); c.stk_st2(slot2, 5, X.p.y, X.p.z);
}
template <class T, class Ctx> MB_HD XfT<T> stk_ld_xf(Ctx &c, int slot2)
{
   XfT<T> X;
   c.stk_ld2(slot2, 0, X.R.xx, X.R.xy); c.stk_ld2(slot2, 1, X.R.xz, X.R.yx); c.stk_ld2(slot2, 2, X.R.yy, X.R.yz);
   c.stk_ld2(slot2, 3, X.R.zx, X.R.zy); c.stk_ld2(slot2, 4, X.R.zz, X.p.x); c.stk_ld2(slot2, 5, X.p.y, X.p.z);
   return X;
}

// the narrow part of a SixDoF slot: the joint transform, 6 double2
template <class T, class Ctx> MB_HD void jp_st_xf(Ctx &c, int slot2, int nslot, const XfT<T> &X)
{
   c.jp_st2(slot2, nslot, 0, X.R.xx, X.R.xy); c.jp_st2(slot2, nslot, 1, X.R.xz, X.R.yx); c.jp_st2(slot2, nslot, 2, X.R.yy, X.R.yz);
   c.jp_st2(slot2, nslot, 3, X.R.zx, X.R.zy); c.jp_st2(slot2, nslot, 4, X.R.zz, X.p.x); c.jp_st2(slot2, nslot, 5, X.p.y, X.p.z);
}
template <class T, class Ctx> MB_HD XfT<T> jp_ld_xf(Ctx &c, int slot2, int nslot)
{
   XfT<T> X;
   c.jp_ld2(slot2, nslot, 0, X.R.xx, X.R.xy); c.jp_ld2(slot2, nslot, 1, X.R.xz, X.R.yx); c.jp_ld2(slot2, nslot, 2, X.R.yy, X.R.yz);
   c.jp_ld2(slot2, nslot, 3, X.R.zx, X.R.zy); c.jp_ld2(slot2, nslot, 4, X.R.zz, X.p.x); c.jp_ld2(slot2, nslot, 5, X.p.y, X.p.z);
   return X;
}

// external wrench on a body, given in its CoM frame (InverseDynamicsCalculator.java:819), re-expressed in the canonical joint frame
template <class T, class Ctx, class CP> MB_HD SvT<T> external_wrench(Ctx &c, int e, const CP C)
{
   SvT<T> w, r;
   w.a = v3<T>(c.ld_fext(e, 0), c.ld_fext(e, 1), c.ld_fext(e, 2));
   w.l = v3<T>(c.ld_fext(e, 3), c.ld_fext(e, 4), c.ld_fext(e, 5));
   M3T<T> E;
   V3T<T> cp;
   ld_com<T>(C, E, cp);
   r.l = mul(E, w.l);
   r.a = mul(E, w.a) + cross(cp, r.l);
   return r;
}
} // namespace mb

#include "multidof.cuh"
