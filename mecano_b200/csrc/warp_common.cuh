// warp_common.cuh -- what the two body-parallel kernel families share: one lane / thread per BODY of one state
// (warp_kernels.cu: a warp per state, trees of up to 32 bodies, exchange by shuffles; team_kernels.cu: a team of two to four warps
// per state, trees of up to 128 bodies, exchange through shared memory).
#pragma once
#include "algorithms.cuh"
#include "kernels.h"

namespace mb
{
namespace
{
__device__ __forceinline__ AbiT<double> abi_zero()
{
   AbiT<double> r;
   r.A.xx = r.A.xy = r.A.xz = r.A.yy = r.A.yz = r.A.zz = 0.0;
   r.L = r.A;
   r.C.xx = r.C.xy = r.C.xz = r.C.yx = r.C.yy = r.C.yz = r.C.zx = r.C.zy = r.C.zz = 0.0;
   return r;
}
__device__ __forceinline__ RbiT<double> rbi_zero()
{
   RbiT<double> r;
   r.I.xx = r.I.xy = r.I.xz = r.I.yy = r.I.yz = r.I.zz = 0.0;
   r.h = v3<double>(0, 0, 0);
   r.m = 0.0;
   return r;
}

// What a lane knows about its body (read once per block from the constant bank / the constant records)
struct Lane
{
   int body;    // internal (depth-first) index = lane; -1 for lanes beyond the tree
   int parent;  // lane of the parent body, -1 = root body
   int jt, dof, cfg, depth, ext;
   unsigned children; // lanes of the child bodies
   XfT<double> X0;    // fixed offset (canonical frames)
   RbiT<double> I;    // inertia about the joint-frame origin
   M3T<double> E;     // CoM frame -> joint frame (external wrenches)
   V3T<double> C;
};

struct Io
{
   const double *q, *qd, *x, *fext;
   double *out;
   long long ld;
   __device__ __forceinline__ double ld_q(int r) const { return __ldg(q + (long long)r * ld); }
   __device__ __forceinline__ double ld_fext(int b, int k) const { return __ldg(fext + (long long)(6 * b + k) * ld); }
   __device__ __forceinline__ void st_out(int r, double v) const { out[(long long)r * ld] = v; }
};

// joint transform of the lane's body (a1; MecanoFactories.java:231-260, PrismaticJointReadOnly.java:18-22, FloatingJointReadOnly.java:34-37)
__device__ __forceinline__ XfT<double> lane_xf(const Lane &L, const Io &io)
{
   XfT<double> X = L.X0;
   if (L.body < 0) return X;
   if (L.jt == MB_SIXDOF)
   {
      const M3T<double> Rq = quat_to_rot(io.ld_q(L.cfg), io.ld_q(L.cfg + 1), io.ld_q(L.cfg + 2), io.ld_q(L.cfg + 3));
      X.R = mul(L.X0.R, Rq);
      X.p = L.X0.p + mul(L.X0.R, v3<double>(io.ld_q(L.cfg + 4), io.ld_q(L.cfg + 5), io.ld_q(L.cfg + 6)));
   }
   else
   {
      const double q = io.ld_q(L.cfg);
      if (L.jt == MB_REVOLUTE)
      {
         double s, c;
         mb_sincos(mb_reduce_angle(q), &s, &c);
         X.R = mul_rz(L.X0.R, s, c);
      }
      else
         X.p = L.X0.p + q * v3<double>(L.X0.R.xz, L.X0.R.yz, L.X0.R.zz);
   }
   return X;
}

// joint velocity-like 6-vector S * x in the joint frame (canonical axis = +z)
__device__ __forceinline__ SvT<double> lane_joint_vec(const Lane &L, const double *base, long long ld, bool use)
{
   SvT<double> r = sv_zero<double>();
   if (L.body < 0 || !use) return r;
   if (L.jt == MB_SIXDOF)
   {
      r.a = v3<double>(__ldg(base + (long long)L.dof * ld), __ldg(base + (long long)(L.dof + 1) * ld), __ldg(base + (long long)(L.dof + 2) * ld));
      r.l = v3<double>(__ldg(base + (long long)(L.dof + 3) * ld), __ldg(base + (long long)(L.dof + 4) * ld), __ldg(base + (long long)(L.dof + 5) * ld));
   }
   else if (L.jt == MB_REVOLUTE)
      r.a.z = __ldg(base + (long long)L.dof * ld);
   else
      r.l.z = __ldg(base + (long long)L.dof * ld);
   return r;
}

template <bool FEXT> __device__ __forceinline__ SvT<double> lane_fext(const Lane &L, const Io &io)
{
   SvT<double> r = sv_zero<double>();
   if (!FEXT || L.body < 0) return r;
   SvT<double> w;
   w.a = v3<double>(io.ld_fext(L.ext, 0), io.ld_fext(L.ext, 1), io.ld_fext(L.ext, 2));
   w.l = v3<double>(io.ld_fext(L.ext, 3), io.ld_fext(L.ext, 4), io.ld_fext(L.ext, 5));
   r.l = mul(L.E, w.l);
   r.a = mul(L.E, w.a) + cross(L.C, r.l);
   return r;
}

} // namespace
} // namespace mb
