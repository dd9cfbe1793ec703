// centroidal.cu -- last step of the centroidal by-products of CompositeRigidBodyMassMatrixCalculator
// (getCentroidalMomentumMatrix() / getCentroidalConvectiveTerm(), M/algorithms/CompositeRigidBodyMassMatrixCalculator.java:380-440,
// :801-839).  The CRBA / RNEA kernels leave the unit momenta (and the convective wrench) expressed in the root frame and the
// first moment of mass of the system in the com rows; this element-wise pass turns (mass * CoM, mass) into (CoM, mass) and, when
// the centroidal momentum frame is the centre-of-mass frame (axes of the root frame, origin at the CoM: what Mecano's
// CenterOfMassReferenceFrame is), moves the moments there: n' = n - c x f for every column.
//
// HBM-bound, no reuse: one thread per state, columns in a loop; every access of a warp is one 256-byte segment.
#include "kernels.h"

namespace mb
{
namespace
{
__global__ void __launch_bounds__(256) centroidal_finish_kernel(const CentroidalArgs a)
{
   const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
   if (s >= a.n)
      return;
   double cx = 0.0, cy = 0.0, cz = 0.0;
   if (a.com)
   {
      double *c = a.com + s;
      cx = c[0]; cy = c[a.ld]; cz = c[2 * a.ld];
      if (a.normalize_com)
      {
         const double inv = 1.0 / c[3 * a.ld];
         cx *= inv; cy *= inv; cz *= inv;
         c[0] = cx; c[a.ld] = cy; c[2 * a.ld] = cz;
      }
   }
   if (!a.shift)
      return;
   double *m = a.cols + s;
   const long long rs = (long long)a.ncols * a.ld; // stride between the six rows of one column
#pragma unroll 1
   for (int j = 0; j < a.ncols; j++, m += a.ld)
   {
      const double fx = m[3 * rs], fy = m[4 * rs], fz = m[5 * rs];
      m[0] -= cy * fz - cz * fy;
      m[rs] -= cz * fx - cx * fz;
      m[2 * rs] -= cx * fy - cy * fx;
   }
}
} // namespace

cudaError_t launch_centroidal_finish(const CentroidalArgs &a, cudaStream_t stream)
{
   if (a.n <= 0)
      return cudaSuccess;
   const unsigned grid = (unsigned)((a.n + 255) / 256);
   centroidal_finish_kernel<<<grid, 256, 0, stream>>>(a);
   return cudaGetLastError();
}
} // namespace mb
