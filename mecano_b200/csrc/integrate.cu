// integrate.cu -- batched MultiBodySystemStateIntegrator.doubleIntegrateFromAcceleration
// (M/tools/MultiBodySystemStateIntegrator.java:365-470 dispatch, :503-560 floating joints, :710-733 one-DoF joints):
// the step that follows ForwardDynamicsCalculator.compute() in a simulation loop, with the state resident in HBM.
//
// HBM-bound, no reuse: every (joint, state) pair is touched once.  A one-DoF joint is an element-wise update of three
// rows (read q, qd, qdd; write q, qd: 40 bytes), so the grid is 2-D: blockIdx.y selects the joint, threads run along
// the states of its rows with 128-bit accesses (two states per thread and access) when the rows are 16-byte aligned.
// A SixDoF joint (read 19 rows, write 16: 280 bytes) is one thread per state doing the SE(3) update in registers.
#include "kernels.h"

namespace mb
{
namespace
{
__device__ __forceinline__ void onedof_update(double dt, double hdt2, double &q, double &v, const double a)
{
   q = fma(hdt2, a, fma(dt, v, q)); // :710-733  q += 0.5 dt^2 qdd + dt qd
   v = fma(dt, a, v);               //           qd += dt qdd
}

// quaternion (x y z s) of a rotation vector (Euclid's Quaternion.setRotationVector)
__device__ __forceinline__ void quat_from_rv(double rx, double ry, double rz, double &x, double &y, double &z, double &s)
{
   const double n = sqrt(rx * rx + ry * ry + rz * rz);
   if (n < 1.0e-12)
   {
      x = y = z = 0.0;
      s = 1.0;
      return;
   }
   double sh, ch;
   sincos(0.5 * n, &sh, &ch);
   sh /= n;
   x = rx * sh; y = ry * sh; z = rz * sh; s = ch;
}

// rotation matrix of a quaternion, normalised first like Euclid does
__device__ __forceinline__ void quat_rot(double qx, double qy, double qz, double qs, double *R)
{
   double n = sqrt(qx * qx + qy * qy + qz * qz + qs * qs);
   if (n < 1.0e-14)
   {
      R[0] = R[4] = R[8] = 1.0;
      R[1] = R[2] = R[3] = R[5] = R[6] = R[7] = 0.0;
      return;
   }
   n = 1.0 / n;
   qx *= n; qy *= n; qz *= n; qs *= n;
   const double yy2 = 2.0 * qy * qy, zz2 = 2.0 * qz * qz, xx2 = 2.0 * qx * qx;
   const double xy2 = 2.0 * qx * qy, sz2 = 2.0 * qs * qz, xz2 = 2.0 * qx * qz;
   const double sy2 = 2.0 * qs * qy, yz2 = 2.0 * qy * qz, sx2 = 2.0 * qs * qx;
   R[0] = 1.0 - yy2 - zz2; R[1] = xy2 - sz2;       R[2] = xz2 + sy2;
   R[3] = xy2 + sz2;       R[4] = 1.0 - xx2 - zz2; R[5] = yz2 - sx2;
   R[6] = xz2 - sy2;       R[7] = yz2 + sx2;       R[8] = 1.0 - xx2 - yy2;
}

// doubleIntegrate(spatialAcceleration, initialTwist, initialPose, finalTwist, finalPose)  (:503-560)
__device__ __forceinline__ void sixdof_update(double dt, double hdt2, double *q7, double *v6, double *a6)
{
   const double wx = v6[0], wy = v6[1], wz = v6[2], vx = v6[3], vy = v6[4], vz = v6[5];
   const double ax = a6[0], ay = a6[1], az = a6[2];
   // linear acceleration of the body origin (SpatialAccelerationReadOnly.java:197-204): a + w x v
   const double lx = a6[3] + (wy * vz - wz * vy), ly = a6[4] + (wz * vx - wx * vz), lz = a6[5] + (wx * vy - wy * vx);
   double ix, iy, iz, is;
   quat_from_rv(dt * wx + hdt2 * ax, dt * wy + hdt2 * ay, dt * wz + hdt2 * az, ix, iy, iz, is);
   const double fwx = dt * ax + wx, fwy = dt * ay + wy, fwz = dt * az + wz;
   double R0[9], Ri[9];
   quat_rot(q7[0], q7[1], q7[2], q7[3], R0);
   quat_rot(ix, iy, iz, is, Ri);
   const double dx = dt * vx + hdt2 * lx, dy = dt * vy + hdt2 * ly, dz = dt * vz + hdt2 * lz;
   q7[4] += R0[0] * dx + R0[1] * dy + R0[2] * dz;
   q7[5] += R0[3] * dx + R0[4] * dy + R0[5] * dz;
   q7[6] += R0[6] * dx + R0[7] * dy + R0[8] * dz;
   // linear velocity and origin acceleration re-expressed in the new body frame (integrated.inverseTransform)
   const double ux = dt * lx + vx, uy = dt * ly + vy, uz = dt * lz + vz;
   const double fvx = Ri[0] * ux + Ri[3] * uy + Ri[6] * uz, fvy = Ri[1] * ux + Ri[4] * uy + Ri[7] * uz, fvz = Ri[2] * ux + Ri[5] * uy + Ri[8] * uz;
   const double mx = Ri[0] * lx + Ri[3] * ly + Ri[6] * lz, my = Ri[1] * lx + Ri[4] * ly + Ri[7] * lz, mz = Ri[2] * lx + Ri[5] * ly + Ri[8] * lz;
   // orientation: q0 * q_integrated (Quaternion.append)
   const double ox = q7[0], oy = q7[1], oz = q7[2], os = q7[3];
   q7[0] = os * ix + ox * is + oy * iz - oz * iy;
   q7[1] = os * iy - ox * iz + oy * is + oz * ix;
   q7[2] = os * iz + ox * iy - oy * ix + oz * is;
   q7[3] = os * is - ox * ix - oy * iy - oz * iz;
   v6[0] = fwx; v6[1] = fwy; v6[2] = fwz;
   v6[3] = fvx; v6[4] = fvy; v6[5] = fvz;
   // setBasedOnOriginAcceleration (FixedFrameSpatialAccelerationBasics.java:81-90): linear = a_origin' + v' x w'
   a6[3] = mx + (fvy * fwz - fvz * fwy);
   a6[4] = my + (fvz * fwx - fvx * fwz);
   a6[5] = mz + (fvx * fwy - fvy * fwx);
}

// SphericalJoint (:449-452, :575-594): orientation = orientation * exp(dt w + 0.5 dt^2 wd), w += dt wd
__device__ __forceinline__ void spherical_update(double dt, double hdt2, double *q4, double *w3, const double *a3)
{
   double ix, iy, iz, is;
   quat_from_rv(dt * w3[0] + hdt2 * a3[0], dt * w3[1] + hdt2 * a3[1], dt * w3[2] + hdt2 * a3[2], ix, iy, iz, is);
   const double ox = q4[0], oy = q4[1], oz = q4[2], os = q4[3];
   q4[0] = os * ix + ox * is + oy * iz - oz * iy;
   q4[1] = os * iy - ox * iz + oy * is + oz * ix;
   q4[2] = os * iz + ox * iy - oy * ix + oz * is;
   q4[3] = os * is - ox * ix - oy * iy - oz * iz;
   w3[0] = fma(dt, a3[0], w3[0]); w3[1] = fma(dt, a3[1], w3[1]); w3[2] = fma(dt, a3[2], w3[2]);
}

// PlanarJoint: a FloatingJointBasics (PlanarJointBasics.java:17), so :421-424 runs the floating-joint update :503-560 on its planar
// pose / twist / acceleration.  Everything stays in the x-z plane: the rotation vector is along y and the pitch-only orientation
// composes additively.  q3 = (pitch, x, z), v3 / a3 = (w_y, v_x, v_z) in the body frame; a3's linear rows are updated like SixDoF's.
__device__ __forceinline__ void planar_update(double dt, double hdt2, double *q3, double *v3, double *a3)
{
   const double th0 = q3[0], wy = v3[0], vx = v3[1], vz = v3[2], wdy = a3[0];
   const double lx = a3[1] + wy * vz, lz = a3[2] - wy * vx; // origin acceleration a + w x v
   const double dth = dt * wy + hdt2 * wdy;
   double s0, c0, si, ci;
   sincos(th0, &s0, &c0);
   sincos(dth, &si, &ci);
   const double tx = dt * vx + hdt2 * lx, tz = dt * vz + hdt2 * lz;
   q3[1] += c0 * tx + s0 * tz; // R_y(th) (x, 0, z) = (c x + s z, 0, -s x + c z)
   q3[2] += -s0 * tx + c0 * tz;
   q3[0] = th0 + dth;
   const double ux = dt * lx + vx, uz = dt * lz + vz;
   const double vfx = ci * ux - si * uz, vfz = si * ux + ci * uz; // R_y(dth)^T
   const double wf = dt * wdy + wy;
   const double l2x = ci * lx - si * lz, l2z = si * lx + ci * lz;
   v3[0] = wf; v3[1] = vfx; v3[2] = vfz;
   a3[1] = l2x - vfz * wf; // a_origin' + v' x w'
   a3[2] = l2z + vfx * wf;
}

template <bool VEC2> __global__ void __launch_bounds__(256) integrate_kernel(const __grid_constant__ IntegrateJoints J, const IntegrateArgs a)
{
   const int j = blockIdx.y;
   const double dt = a.dt, hdt2 = 0.5 * a.dt * a.dt;
   const long long stride = (long long)gridDim.x * blockDim.x;
   const int cfg = J.cfg[j], dof = J.dof[j];
   if (J.type[j] != MB_SIXDOF)
   {
      double *q = a.q + (long long)cfg * a.ld, *v = a.qd + (long long)dof * a.ld;
      const double *acc = a.qdd + (long long)dof * a.ld;
      if (VEC2)
      {
         const long long n2 = a.n >> 1;
         for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += stride)
         {
            double2 q2 = reinterpret_cast<double2 *>(q)[i], v2 = reinterpret_cast<double2 *>(v)[i];
            const double2 a2 = __ldcs(reinterpret_cast<const double2 *>(acc) + i);
            onedof_update(dt, hdt2, q2.x, v2.x, a2.x);
            onedof_update(dt, hdt2, q2.y, v2.y, a2.y);
            reinterpret_cast<double2 *>(q)[i] = q2;
            reinterpret_cast<double2 *>(v)[i] = v2;
         }
         if ((a.n & 1) && blockIdx.x == 0 && threadIdx.x == 0)
            onedof_update(dt, hdt2, q[a.n - 1], v[a.n - 1], acc[a.n - 1]);
      }
      else
         for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += stride)
            onedof_update(dt, hdt2, q[i], v[i], acc[i]);
      return;
   }
   if (J.sub[j] != MB_SUB_SIX)
   {
      const bool planar = J.sub[j] == MB_SUB_PLANAR;
      const int nc = planar ? 3 : 4;
      for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += stride)
      {
         double q4[4], v3[3], a3[3];
         for (int k = 0; k < nc; k++) q4[k] = a.q[(long long)(cfg + k) * a.ld + i];
#pragma unroll
         for (int k = 0; k < 3; k++) { v3[k] = a.qd[(long long)(dof + k) * a.ld + i]; a3[k] = a.qdd[(long long)(dof + k) * a.ld + i]; }
         if (planar) planar_update(dt, hdt2, q4, v3, a3);
         else spherical_update(dt, hdt2, q4, v3, a3);
         for (int k = 0; k < nc; k++) a.q[(long long)(cfg + k) * a.ld + i] = q4[k];
#pragma unroll
         for (int k = 0; k < 3; k++) a.qd[(long long)(dof + k) * a.ld + i] = v3[k];
         if (planar)
         {
            a.qdd[(long long)(dof + 1) * a.ld + i] = a3[1];
            a.qdd[(long long)(dof + 2) * a.ld + i] = a3[2];
         }
      }
      return;
   }
   for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += stride)
   {
      double q7[7], v6[6], a6[6];
#pragma unroll
      for (int k = 0; k < 7; k++) q7[k] = a.q[(long long)(cfg + k) * a.ld + i];
#pragma unroll
      for (int k = 0; k < 6; k++) { v6[k] = a.qd[(long long)(dof + k) * a.ld + i]; a6[k] = a.qdd[(long long)(dof + k) * a.ld + i]; }
      sixdof_update(dt, hdt2, q7, v6, a6);
#pragma unroll
      for (int k = 0; k < 7; k++) a.q[(long long)(cfg + k) * a.ld + i] = q7[k];
#pragma unroll
      for (int k = 0; k < 6; k++) a.qd[(long long)(dof + k) * a.ld + i] = v6[k];
#pragma unroll
      for (int k = 3; k < 6; k++) a.qdd[(long long)(dof + k) * a.ld + i] = a6[k];
   }
}
} // namespace

cudaError_t launch_integrate_kernel(const IntegrateJoints &J, const IntegrateArgs &a, int sm_count, cudaStream_t stream)
{
   if (a.n <= 0 || J.nb <= 0)
      return cudaSuccess;
   const bool vec2 = ((a.ld & 1) == 0) && (((uintptr_t)a.q | (uintptr_t)a.qd | (uintptr_t)a.qdd) & 15) == 0;
   const long long work = vec2 ? (a.n + 1) / 2 : a.n;
   // enough blocks per joint row to fill the machine even for a single joint, capped so that the grid stays a few waves
   long long bx = (work + 255) / 256;
   const long long cap = std::max<long long>(1, (long long)sm_count * 8 / std::max(1, J.nb) + 1) * 4;
   bx = std::max<long long>(1, std::min(bx, cap));
   const dim3 grid((unsigned)bx, (unsigned)J.nb);
   if (vec2)
      integrate_kernel<true><<<grid, 256, 0, stream>>>(J, a);
   else
      integrate_kernel<false><<<grid, 256, 0, stream>>>(J, a);
   return cudaGetLastError();
}
} // namespace mb
