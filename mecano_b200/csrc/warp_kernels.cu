// warp_kernels.cu -- sm_100a kernels, one WARP per state: lane = body (trees of up to 32 bodies).
//
// The thread-per-state kernels (kernels.cu) need >= 148 SMs x 256..512 states to fill the machine and one state costs
// a full serial walk over the tree; for small batches (latency-bound callers: one controller tick over a few thousand
// candidate states) that leaves most of the GPU idle.  Here the 32 lanes of a warp hold the 32 bodies of ONE state:
//   * everything that does not depend on other bodies -- joint transforms (one sincos per lane), Newton-Euler body
//     wrenches, bias terms, unit momenta -- is evaluated by all lanes at once;
//   * the two sweeps of each recursion run level by level (tree depth, not body count, iterations): a lane reads its
//     parent's twist / acceleration with warp shuffles, and a parent sums the wrenches / inertias of its children --
//     the sibling-subtree reduction -- by shuffling them in, child after child in lane order (deterministic);
//   * all spatial quantities of a body live in the registers of its lane: no shared memory, no stack, no workspace
//     (ABA's pass-two results g = U/D, u/D stay in the lane until pass three).
// Same canonical joint frames, constant records and 6-D routines (spatial.cuh, jointmath.cuh) as the thread-per-state
// kernels, so the two variants agree to round-off.  Global accesses are one 8-byte word per lane and row (uncoalesced):
// irrelevant in the latency-bound regime this variant is for, and the reason it loses at large batches.
//
// Reference semantics: InverseDynamicsCalculator.java:873-966, ForwardDynamicsCalculator.java:1085-1310,
// CompositeRigidBodyMassMatrixCalculator.java:588-667, 700-707, 772-797 (as rnea.cuh / aba.cuh / crba.cuh).
#include <algorithm>

#include "warp_common.cuh"

namespace mb
{
namespace
{
constexpr unsigned kFull = 0xffffffffu;
constexpr int kWarpBlock = 128; // 4 warps = 4 states per block

__device__ __forceinline__ double shfl(double x, int src) { return __shfl_sync(kFull, x, src); }
__device__ __forceinline__ V3T<double> shfl(const V3T<double> &v, int src) { return v3<double>(shfl(v.x, src), shfl(v.y, src), shfl(v.z, src)); }
__device__ __forceinline__ SvT<double> shfl(const SvT<double> &v, int src)
{
   SvT<double> r;
   r.a = shfl(v.a, src);
   r.l = shfl(v.l, src);
   return r;
}
__device__ __forceinline__ S3T<double> shfl(const S3T<double> &s, int src)
{
   S3T<double> r;
   r.xx = shfl(s.xx, src); r.xy = shfl(s.xy, src); r.xz = shfl(s.xz, src); r.yy = shfl(s.yy, src); r.yz = shfl(s.yz, src); r.zz = shfl(s.zz, src);
   return r;
}
__device__ __forceinline__ M3T<double> shfl(const M3T<double> &m, int src)
{
   M3T<double> r;
   r.xx = shfl(m.xx, src); r.xy = shfl(m.xy, src); r.xz = shfl(m.xz, src);
   r.yx = shfl(m.yx, src); r.yy = shfl(m.yy, src); r.yz = shfl(m.yz, src);
   r.zx = shfl(m.zx, src); r.zy = shfl(m.zy, src); r.zz = shfl(m.zz, src);
   return r;
}
__device__ __forceinline__ XfT<double> shfl(const XfT<double> &X, int src)
{
   XfT<double> r;
   r.R = shfl(X.R, src);
   r.p = shfl(X.p, src);
   return r;
}
__device__ __forceinline__ RbiT<double> shfl(const RbiT<double> &I, int src)
{
   RbiT<double> r;
   r.I = shfl(I.I, src);
   r.h = shfl(I.h, src);
   r.m = shfl(I.m, src);
   return r;
}
__device__ __forceinline__ AbiT<double> shfl(const AbiT<double> &I, int src)
{
   AbiT<double> r;
   r.A = shfl(I.A, src);
   r.C = shfl(I.C, src);
   r.L = shfl(I.L, src);
   return r;
}
// The traversal program is read from global memory here (a pointer argument): passing the 17 KB MbProgram by value, as the
// thread-per-state kernels do to get it into the constant bank, costs ~20 us of launch time -- more than a whole batch
// in the regime this variant serves -- and the program is only consulted once per block.
template <bool FEXT> __device__ __forceinline__ Lane lane_setup(const MbProgram &P, const double *consts)
{
   Lane L;
   const int lane = threadIdx.x & 31;
   const bool on = lane < P.nb;
   const int b = on ? lane : 0;
   const MbBody B = P.body[b];
   L.body = on ? lane : -1;
   L.parent = on ? B.parent : -1;
   L.jt = B.jtype; L.dof = B.dof_off; L.cfg = B.cfg_off; L.depth = on ? B.depth : -1; L.ext = B.ext_index;
   const double *C = consts + (size_t)b * MB_CONST_STRIDE;
   L.X0.R = ld_m3(C + MB_C_R);
   L.X0.p = ld_v3(C + MB_C_P);
   L.I = ld_rbi<double>(C);
   if (FEXT)
   {
      L.E = ld_m3(C + MB_C_E);
      L.C = ld_v3(C + MB_C_C);
   }
   // children: lanes whose parent is this lane
   L.children = 0;
   for (int p = 0; p < P.nb; p++)
   {
      const unsigned m = __ballot_sync(kFull, on && L.parent == p);
      if (lane == p) L.children = m;
   }
   return L;
}

// sum over the children of each lane of `contrib` (held by the child lanes), child after child in lane order
template <class V, class Add> __device__ __forceinline__ void gather_children(const Lane &L, bool parent_active, int maxc, const V &contrib, V &acc, Add add)
{
   unsigned m = parent_active ? L.children : 0u;
   const int lane = threadIdx.x & 31;
   for (int it = 0; it < maxc; it++)
   {
      const int src = m ? (__ffs(m) - 1) : lane;
      const V val = shfl(contrib, src);
      if (m) add(acc, val);
      m &= m - 1;
   }
}

// ------------------------------------------------------------------------------------------------ RNEA
template <bool FEXT> __global__ void __launch_bounds__(kWarpBlock) warp_rnea_kernel(const MbProgram *__restrict__ Pp, const KernelArgs a, int maxc)
{
   const MbProgram &P = *Pp;
   const Lane L = lane_setup<FEXT>(P, a.consts);
   const bool use_qd = !(a.flags & 1u), use_qdd = !(a.flags & 2u);
   const long long nwarps = (long long)gridDim.x * (kWarpBlock / 32);
   const int nlev = P.max_depth;
   for (long long s = (long long)blockIdx.x * (kWarpBlock / 32) + (threadIdx.x >> 5); s < a.n; s += nwarps)
   {
      Io io{a.q + s, a.qd + s, a.x + s, a.fext + s, a.out + s, a.ld};
      // ---- all bodies at once: joint transforms, joint twists / accelerations
      const XfT<double> X = lane_xf(L, io);
      const SvT<double> vj = lane_joint_vec(L, io.qd, a.ld, use_qd);
      const SvT<double> aj = lane_joint_vec(L, io.x, a.ld, use_qdd);
      // ---- pass one (:873-917), level by level: parent quantities arrive by shuffle
      SvT<double> v = sv_zero<double>(), acc = sv_zero<double>();
      const int psrc = L.parent < 0 ? (threadIdx.x & 31) : L.parent;
      for (int lev = 0; lev < nlev; lev++)
      {
         SvT<double> pv = shfl(v, psrc), pa = shfl(acc, psrc);
         if (L.depth == lev)
         {
            if (L.parent < 0)
            {
               pv = sv_zero<double>();
               pa = sv_zero<double>();
               pa.l = v3<double>(-a.grav[0], -a.grav[1], -a.grav[2]); // root acceleration = -gravity (:397-403)
            }
            v = motion_to_child(X, pv) + vj;
            acc = motion_to_child(X, pa) + cross_motion(v, vj) + aj;
         }
      }
      // ---- Newton-Euler wrench of every body at once (SpatialInertiaReadOnly.java:229-296), about the joint-frame origin
      SvT<double> f = mul(L.I, acc) + cross_force(v, mul(L.I, v));
      if (FEXT) f = f - lane_fext<FEXT>(L, io);
      if (L.body < 0) f = sv_zero<double>();
      // ---- pass two (:930-966), leaves first: tau = S^T W, parents sum their children's wrenches
      for (int lev = nlev - 1; lev >= 0; lev--)
      {
         SvT<double> contrib = sv_zero<double>();
         if (L.depth == lev)
         {
            if (L.jt == MB_SIXDOF)
            {
               io.st_out(L.dof + 0, f.a.x); io.st_out(L.dof + 1, f.a.y); io.st_out(L.dof + 2, f.a.z);
               io.st_out(L.dof + 3, f.l.x); io.st_out(L.dof + 4, f.l.y); io.st_out(L.dof + 5, f.l.z);
            }
            else
               io.st_out(L.dof, L.jt == MB_REVOLUTE ? f.a.z : f.l.z);
            contrib = force_to_parent(X, f);
         }
         if (lev > 0)
            gather_children(L, L.depth == lev - 1, maxc, contrib, f, [](SvT<double> &x, const SvT<double> &y) { x = x + y; });
      }
   }
}

// ------------------------------------------------------------------------------------------------ ABA
struct AbaContrib
{
   AbiT<double> K;
   SvT<double> P;
};
__device__ __forceinline__ AbaContrib shfl(const AbaContrib &c, int src)
{
   AbaContrib r;
   r.K = shfl(c.K, src);
   r.P = shfl(c.P, src);
   return r;
}

template <bool FEXT> __global__ void __launch_bounds__(kWarpBlock) warp_aba_kernel(const MbProgram *__restrict__ Pp, const KernelArgs a, int maxc)
{
   const MbProgram &P = *Pp;
   const Lane L = lane_setup<FEXT>(P, a.consts);
   const long long nwarps = (long long)gridDim.x * (kWarpBlock / 32);
   const int nlev = P.max_depth;
   const int lane = threadIdx.x & 31;
   for (long long s = (long long)blockIdx.x * (kWarpBlock / 32) + (threadIdx.x >> 5); s < a.n; s += nwarps)
   {
      Io io{a.q + s, a.qd + s, a.x + s, a.fext + s, a.out + s, a.ld};
      const XfT<double> X = lane_xf(L, io);
      const SvT<double> vj = lane_joint_vec(L, io.qd, a.ld, true);
      const SvT<double> tauj = lane_joint_vec(L, io.x, a.ld, true); // S * tau (only the joint's own components are non-zero)
      // ---- pass one (:1085-1127): twists level by level, then bias terms for all bodies at once
      SvT<double> v = sv_zero<double>();
      const int psrc = L.parent < 0 ? lane : L.parent;
      for (int lev = 0; lev < nlev; lev++)
      {
         SvT<double> pv = shfl(v, psrc);
         if (L.depth == lev)
         {
            if (L.parent < 0) pv = sv_zero<double>();
            v = motion_to_child(X, pv) + vj;
         }
      }
      const SvT<double> cb = cross_motion(v, vj); // bias acceleration (:1114-1118)
      SvT<double> pA = cross_force(v, mul(L.I, v));
      if (FEXT) pA = pA - lane_fext<FEXT>(L, io);
      AbiT<double> IA = abi_from_rbi(L.I);
      if (L.body < 0)
      {
         pA = sv_zero<double>();
         IA = abi_zero();
      }
      // ---- pass two (:1136-1254), leaves first; g = U / D and k0 = u / D stay in the lane for pass three
      SvT<double> g = sv_zero<double>();
      double k0 = 0.0;
      SvT<double> x6 = sv_zero<double>(); // SixDoF: (I^A)^-1 (tau - p^A)
      for (int lev = nlev - 1; lev >= 0; lev--)
      {
         AbaContrib c;
         c.K = abi_zero();
         c.P = sv_zero<double>();
         if (L.depth == lev)
         {
            if (L.jt == MB_SIXDOF)
            {
               x6 = abi_solve(IA, tauj - pA);
               if (L.parent >= 0) c.P = force_to_parent(X, tauj); // the joint transmits nothing but tau
            }
            else
            {
               const bool rev = L.jt == MB_REVOLUTE;
               SvT<double> U;
               double D, u;
               if (rev)
               {
                  U.a = v3<double>(IA.A.xz, IA.A.yz, IA.A.zz);
                  U.l = v3<double>(IA.C.zx, IA.C.zy, IA.C.zz);
                  D = IA.A.zz;
                  u = tauj.a.z - pA.a.z;
               }
               else
               {
                  U.a = v3<double>(IA.C.xz, IA.C.yz, IA.C.zz);
                  U.l = v3<double>(IA.L.xz, IA.L.yz, IA.L.zz);
                  D = IA.L.zz;
                  u = tauj.l.z - pA.l.z;
               }
               const double Dinv = mb_rcp(D);
               g.a = Dinv * U.a;
               g.l = Dinv * U.l;
               k0 = Dinv * u;
               if (L.parent >= 0)
               {
                  const AbiT<double> Ia = abi_downdate(IA, U, g);
                  SvT<double> pa = pA + mul(Ia, cb); // p^a = p^A + I^a c + U D^-1 u
                  pa.a = pa.a + k0 * U.a;
                  pa.l = pa.l + k0 * U.l;
                  c.K = abi_to_parent(X, Ia);
                  c.P = force_to_parent(X, pa);
               }
            }
         }
         if (lev > 0)
         {
            AbaContrib accum;
            accum.K = IA;
            accum.P = pA;
            gather_children(L, L.depth == lev - 1, maxc, c, accum, [](AbaContrib &x, const AbaContrib &y) {
               x.K = x.K + y.K;
               x.P = x.P + y.P;
            });
            IA = accum.K;
            pA = accum.P;
         }
      }
      // ---- pass three (:1259-1310), root first
      SvT<double> acc = sv_zero<double>();
      for (int lev = 0; lev < nlev; lev++)
      {
         SvT<double> pa = shfl(acc, psrc);
         if (L.depth == lev)
         {
            if (L.parent < 0)
            {
               pa = sv_zero<double>();
               pa.l = v3<double>(-a.grav[0], -a.grav[1], -a.grav[2]);
            }
            const SvT<double> a1 = motion_to_child(X, pa) + cb;
            if (L.jt == MB_SIXDOF)
            {
               const SvT<double> qdd = x6 - a1;
               io.st_out(L.dof + 0, qdd.a.x); io.st_out(L.dof + 1, qdd.a.y); io.st_out(L.dof + 2, qdd.a.z);
               io.st_out(L.dof + 3, qdd.l.x); io.st_out(L.dof + 4, qdd.l.y); io.st_out(L.dof + 5, qdd.l.z);
               acc = x6;
            }
            else
            {
               const double qdd = k0 - (dot(g.a, a1.a) + dot(g.l, a1.l));
               io.st_out(L.dof, qdd);
               acc = a1;
               if (L.jt == MB_REVOLUTE) acc.a.z += qdd;
               else acc.l.z += qdd;
            }
         }
      }
   }
}

// ------------------------------------------------------------------------------------------------ CRBA
template <bool STATE_MAJOR>
__global__ void __launch_bounds__(kWarpBlock) warp_crba_kernel(const MbProgram *__restrict__ Pp, const KernelArgs a, int maxc, int max_ndof)
{
   const MbProgram &P = *Pp;
   const Lane L = lane_setup<false>(P, a.consts);
   const long long nwarps = (long long)gridDim.x * (kWarpBlock / 32);
   const int nlev = P.max_depth, nv = a.nv;
   const int lane = threadIdx.x & 31;
   const long long mstride = STATE_MAJOR ? 1 : a.ld;
   for (long long s = (long long)blockIdx.x * (kWarpBlock / 32) + (threadIdx.x >> 5); s < a.n; s += nwarps)
   {
      Io io{a.q + s, nullptr, nullptr, nullptr, nullptr, a.ld};
      double *M = STATE_MAJOR ? a.out + s * (long long)nv * nv : a.out + s;
      // entries coupling joints of unrelated branches are zero (massMatrix.zero(), :296)
      for (int k = lane; k < a.n_zero; k += 32)
         M[(long long)a.zero_entries[k] * mstride] = 0.0;
      const XfT<double> X = lane_xf(L, io);
      // ---- composite inertias (:648-661), leaves first
      RbiT<double> Ic = L.body < 0 ? rbi_zero() : L.I;
      for (int lev = nlev - 1; lev > 0; lev--)
      {
         RbiT<double> contrib = rbi_zero();
         if (L.depth == lev) contrib = rbi_to_parent(X, Ic);
         gather_children(L, L.depth == lev - 1, maxc, contrib, Ic, [](RbiT<double> &x, const RbiT<double> &y) { x = x + y; });
      }
      // ---- columns: unit momenta F = Ic S (:663-667), diagonal block (:700-707), walk to the root (:772-797)
      for (int col = 0; col < max_ndof; col++)
      {
         const int ndof = L.body < 0 ? 0 : (L.jt == MB_SIXDOF ? 6 : 1);
         const bool has_col = col < ndof;
         SvT<double> e = sv_zero<double>();
         if (L.jt == MB_SIXDOF)
         {
            if (col == 0) e.a.x = 1; else if (col == 1) e.a.y = 1; else if (col == 2) e.a.z = 1;
            else if (col == 3) e.l.x = 1; else if (col == 4) e.l.y = 1; else e.l.z = 1;
         }
         else if (L.jt == MB_REVOLUTE) e.a.z = 1;
         else e.l.z = 1;
         SvT<double> F = mul(Ic, e);
         const int dc = L.dof + col;
         if (has_col)
         {
            if (L.jt == MB_SIXDOF)
            {
               M[(long long)((L.dof + 0) * nv + dc) * mstride] = F.a.x; M[(long long)((L.dof + 1) * nv + dc) * mstride] = F.a.y;
               M[(long long)((L.dof + 2) * nv + dc) * mstride] = F.a.z; M[(long long)((L.dof + 3) * nv + dc) * mstride] = F.l.x;
               M[(long long)((L.dof + 4) * nv + dc) * mstride] = F.l.y; M[(long long)((L.dof + 5) * nv + dc) * mstride] = F.l.z;
            }
            else
               M[(long long)(dc * nv + dc) * mstride] = L.jt == MB_REVOLUTE ? F.a.z : F.l.z;
         }
         // every lane walks up its own ancestor chain; the ancestors' transforms and joint data arrive by shuffle
         int j = L.body < 0 ? lane : L.body; // body whose frame F is expressed in
         XfT<double> Xj = X;
         int pj = L.parent;
         for (int step = 1; step < nlev; step++)
         {
            const bool go = has_col && pj >= 0;
            if (go) F = force_to_parent(Xj, F);
            const int src = go ? pj : lane;
            // joint data of the ancestor
            const XfT<double> Xn = shfl(X, src);
            const int jt_n = __shfl_sync(kFull, L.jt, src), dof_n = __shfl_sync(kFull, L.dof, src), par_n = __shfl_sync(kFull, L.parent, src);
            if (go)
            {
               j = pj;
               if (jt_n == MB_SIXDOF)
               {
                  const double ev[6] = {F.a.x, F.a.y, F.a.z, F.l.x, F.l.y, F.l.z};
#pragma unroll
                  for (int r = 0; r < 6; r++)
                  {
                     M[(long long)((dof_n + r) * nv + dc) * mstride] = ev[r];
                     M[(long long)(dc * nv + dof_n + r) * mstride] = ev[r];
                  }
               }
               else
               {
                  const double val = jt_n == MB_REVOLUTE ? F.a.z : F.l.z;
                  M[(long long)(dof_n * nv + dc) * mstride] = val;
                  M[(long long)(dc * nv + dof_n) * mstride] = val;
               }
               Xj = Xn;
               pj = par_n;
            }
         }
         (void)j;
      }
   }
}
} // namespace

bool warp_variant_supports(const MbProgram &P)
{
   if (P.nb > MB_MAX_BODIES)
      return false; // (up to 32 bodies: a warp per state; 33 to 128: a team of warps, team_kernels.cu)
   for (int i = 0; i < P.nb; i++)
      if (P.body[i].sub != MB_SUB_SIX)
         return false; // spherical / planar joints run on the thread-per-state kernels
   return true;
}

// P: device copy of the traversal program

cudaError_t launch_warp_kernel(int algo, const MbProgram *P, const KernelArgs &a, int max_children, int max_ndof, int sm_count, cudaStream_t stream)
{
   if (a.n <= 0)
      return cudaSuccess;
   const long long blocks_needed = (a.n + (kWarpBlock / 32) - 1) / (kWarpBlock / 32);
   const unsigned grid = (unsigned)std::min<long long>(blocks_needed, (long long)sm_count * 16);
   if (algo == MB_RNEA)
   {
      if (a.fext) warp_rnea_kernel<true><<<grid, kWarpBlock, 0, stream>>>(P, a, max_children);
      else warp_rnea_kernel<false><<<grid, kWarpBlock, 0, stream>>>(P, a, max_children);
   }
   else if (algo == MB_ABA)
   {
      if (a.fext) warp_aba_kernel<true><<<grid, kWarpBlock, 0, stream>>>(P, a, max_children);
      else warp_aba_kernel<false><<<grid, kWarpBlock, 0, stream>>>(P, a, max_children);
   }
   else
   {
      if (a.flags & 1u) warp_crba_kernel<true><<<grid, kWarpBlock, 0, stream>>>(P, a, max_children, max_ndof);
      else warp_crba_kernel<false><<<grid, kWarpBlock, 0, stream>>>(P, a, max_children, max_ndof);
   }
   return cudaGetLastError();
}

cudaError_t warp_kernel_attributes(int algo, bool fext, cudaFuncAttributes *attr)
{
   if (algo == MB_RNEA) return fext ? cudaFuncGetAttributes(attr, warp_rnea_kernel<true>) : cudaFuncGetAttributes(attr, warp_rnea_kernel<false>);
   if (algo == MB_ABA) return fext ? cudaFuncGetAttributes(attr, warp_aba_kernel<true>) : cudaFuncGetAttributes(attr, warp_aba_kernel<false>);
   return cudaFuncGetAttributes(attr, warp_crba_kernel<false>);
}
} // namespace mb
