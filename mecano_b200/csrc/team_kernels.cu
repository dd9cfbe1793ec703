// team_kernels.cu -- sm_100a kernels, one TEAM of two to four warps per state: thread = body, trees of 33 to 128 bodies.
//
// The body-parallel counterpart of warp_kernels.cu for trees that do not fit one warp (BASELINE.json config 5: "tree-size sweep
// 7-100 bodies (thread- vs warp-per-state)").  A block of 32 * ceil(nb / 32) threads holds one state at a time:
//   * as in the warp kernels, everything that does not depend on other bodies (joint transforms, Newton-Euler wrenches, bias
//     terms, unit momenta) is evaluated by all threads at once, and the sweeps run level by level (tree depth iterations);
//   * bodies of different warps cannot shuffle, so the exchange goes through shared memory, laid out component-major /
//     body-minor ([k][body]: a warp's own-slot accesses are conflict-free): a body publishes its twist / acceleration for its
//     children, and children ADD their wrench / inertia contribution into their parent's slot, one sibling rank per round
//     (round r: the r-th child of every parent, in body order) -- the sibling-subtree reduction, deterministic, with one
//     __syncthreads() per round;
//   * the per-body results that pass three needs (ABA: g = U / D, u / D) stay in the registers of the body's thread.
// Same canonical joint frames, constant records and 6-D routines as every other variant; agrees with them to round-off
// (the order of the sibling sums differs).  Latency variant: the crossover against thread-per-state is measured in
// profiles/ (tree sweep).
//
// Reference semantics: InverseDynamicsCalculator.java:873-966, ForwardDynamicsCalculator.java:1085-1310,
// CompositeRigidBodyMassMatrixCalculator.java:588-667, 700-707, 772-797 (as rnea.cuh / aba.cuh / crba.cuh).
#include <algorithm>

#include "warp_common.cuh"

namespace mb
{
namespace
{
constexpr int kTeamMax = 128; // = MB_MAX_BODIES

// one body's column of a [k][T] shared array
struct Col
{
   double *p;
   int T;
   __device__ __forceinline__ double &operator[](int k) const { return p[k * T]; }
};

__device__ __forceinline__ void col_st(const Col c, const SvT<double> &v)
{
   c[0] = v.a.x; c[1] = v.a.y; c[2] = v.a.z; c[3] = v.l.x; c[4] = v.l.y; c[5] = v.l.z;
}
__device__ __forceinline__ SvT<double> col_ld_sv(const Col c)
{
   SvT<double> v;
   v.a = v3<double>(c[0], c[1], c[2]);
   v.l = v3<double>(c[3], c[4], c[5]);
   return v;
}
__device__ __forceinline__ void col_add(const Col c, const SvT<double> &v)
{
   c[0] += v.a.x; c[1] += v.a.y; c[2] += v.a.z; c[3] += v.l.x; c[4] += v.l.y; c[5] += v.l.z;
}
__device__ __forceinline__ void s3_to(double *d, const S3T<double> &s) { d[0] = s.xx; d[1] = s.xy; d[2] = s.xz; d[3] = s.yy; d[4] = s.yz; d[5] = s.zz; }
__device__ __forceinline__ S3T<double> s3_from(const double *d)
{
   S3T<double> s;
   s.xx = d[0]; s.xy = d[1]; s.xz = d[2]; s.yy = d[3]; s.yz = d[4]; s.zz = d[5];
   return s;
}
__device__ __forceinline__ void m3_to(double *d, const M3T<double> &m)
{
   d[0] = m.xx; d[1] = m.xy; d[2] = m.xz; d[3] = m.yx; d[4] = m.yy; d[5] = m.yz; d[6] = m.zx; d[7] = m.zy; d[8] = m.zz;
}
__device__ __forceinline__ M3T<double> m3_from(const double *d)
{
   M3T<double> m;
   m.xx = d[0]; m.xy = d[1]; m.xz = d[2]; m.yx = d[3]; m.yy = d[4]; m.yz = d[5]; m.zx = d[6]; m.zy = d[7]; m.zz = d[8];
   return m;
}
// articulated inertia (21) + bias wrench (6)
constexpr int kAbaCols = 27;
__device__ __forceinline__ void abi_flat(const AbiT<double> &I, const SvT<double> &p, double d[kAbaCols])
{
   s3_to(d, I.A); m3_to(d + 6, I.C); s3_to(d + 15, I.L);
   d[21] = p.a.x; d[22] = p.a.y; d[23] = p.a.z; d[24] = p.l.x; d[25] = p.l.y; d[26] = p.l.z;
}
__device__ __forceinline__ void abi_unflat(const double d[kAbaCols], AbiT<double> &I, SvT<double> &p)
{
   I.A = s3_from(d); I.C = m3_from(d + 6); I.L = s3_from(d + 15);
   p.a = v3<double>(d[21], d[22], d[23]);
   p.l = v3<double>(d[24], d[25], d[26]);
}
// composite rigid-body inertia (10)
constexpr int kRbiCols = 10;
__device__ __forceinline__ void rbi_flat(const RbiT<double> &I, double d[kRbiCols])
{
   s3_to(d, I.I); d[6] = I.h.x; d[7] = I.h.y; d[8] = I.h.z; d[9] = I.m;
}
__device__ __forceinline__ RbiT<double> rbi_unflat(const double d[kRbiCols])
{
   RbiT<double> I;
   I.I = s3_from(d); I.h = v3<double>(d[6], d[7], d[8]); I.m = d[9];
   return I;
}
constexpr int kXfCols = 12;

// thread = body.  `ipar` (shared, T ints) receives the parent of every body; *rank = position of this body among its parent's
// children in body order (the round in which it adds its contribution).
template <bool FEXT> __device__ __forceinline__ Lane team_setup(const MbProgram &P, const double *consts, int *ipar, int *rank)
{
   Lane L;
   const int t = threadIdx.x;
   const bool on = t < P.nb;
   const int b = on ? t : 0;
   const MbBody B = P.body[b];
   L.body = on ? t : -1;
   L.parent = on ? B.parent : -1;
   L.jt = B.jtype; L.dof = B.dof_off; L.cfg = B.cfg_off; L.depth = on ? B.depth : -1; L.ext = B.ext_index;
   L.children = 0;
   const double *C = consts + (size_t)b * MB_CONST_STRIDE;
   L.X0.R = ld_m3(C + MB_C_R);
   L.X0.p = ld_v3(C + MB_C_P);
   L.I = ld_rbi<double>(C);
   if (FEXT)
   {
      L.E = ld_m3(C + MB_C_E);
      L.C = ld_v3(C + MB_C_C);
   }
   ipar[t] = on ? B.parent : -2;
   __syncthreads();
   int r = 0;
   for (int c = 0; c < t; c++)
      r += (on && ipar[c] == L.parent) ? 1 : 0;
   *rank = r;
   return L;
}

// ------------------------------------------------------------------------------------------------ RNEA
// shared: V[6][T] (twists, then the wrench accumulators), A[6][T] (accelerations), ipar[T]
template <bool FEXT> __global__ void __launch_bounds__(kTeamMax) team_rnea_kernel(const MbProgram *__restrict__ Pp, const KernelArgs a, int maxc)
{
   extern __shared__ double tsm[];
   const MbProgram &P = *Pp;
   const int T = blockDim.x, t = threadIdx.x;
   int rank;
   const Lane L = team_setup<FEXT>(P, a.consts, reinterpret_cast<int *>(tsm + 12 * T), &rank);
   const bool use_qd = !(a.flags & 1u), use_qdd = !(a.flags & 2u);
   const int nlev = P.max_depth;
   const Col V{tsm + t, T}, A{tsm + 6 * T + t, T};
   const int pp = L.parent < 0 ? t : L.parent;
   const Col PV{tsm + pp, T}, PA{tsm + 6 * T + pp, T};
   for (long long s = blockIdx.x; s < a.n; s += gridDim.x)
   {
      Io io{a.q + s, a.qd + s, a.x + s, a.fext + s, a.out + s, a.ld};
      const XfT<double> X = lane_xf(L, io);
      const SvT<double> vj = lane_joint_vec(L, io.qd, a.ld, use_qd);
      const SvT<double> aj = lane_joint_vec(L, io.x, a.ld, use_qdd);
      // ---- pass one (:873-917), level by level: a body reads its parent's twist / acceleration from the parent's slot
      SvT<double> v = sv_zero<double>(), acc = sv_zero<double>();
      for (int lev = 0; lev < nlev; lev++)
      {
         if (L.depth == lev)
         {
            SvT<double> pv = sv_zero<double>(), pa = sv_zero<double>();
            if (L.parent < 0)
               pa.l = v3<double>(-a.grav[0], -a.grav[1], -a.grav[2]); // root acceleration = -gravity (:397-403)
            else
            {
               pv = col_ld_sv(PV);
               pa = col_ld_sv(PA);
            }
            v = motion_to_child(X, pv) + vj;
            acc = motion_to_child(X, pa) + cross_motion(v, vj) + aj;
            col_st(V, v);
            col_st(A, acc);
         }
         __syncthreads();
      }
      // ---- Newton-Euler wrench of every body at once, about the joint-frame origin; V becomes the wrench accumulator
      SvT<double> f = mul(L.I, acc) + cross_force(v, mul(L.I, v));
      if (FEXT) f = f - lane_fext<FEXT>(L, io);
      if (L.body < 0) f = sv_zero<double>();
      col_st(V, f);
      __syncthreads();
      // ---- pass two (:930-966), leaves first: tau = S^T W; children add their wrench to the parent's slot, rank by rank
      for (int lev = nlev - 1; lev >= 0; lev--)
      {
         const bool mine = L.depth == lev;
         SvT<double> contrib = sv_zero<double>();
         if (mine)
         {
            f = col_ld_sv(V);
            if (L.jt == MB_SIXDOF)
            {
               io.st_out(L.dof + 0, f.a.x); io.st_out(L.dof + 1, f.a.y); io.st_out(L.dof + 2, f.a.z);
               io.st_out(L.dof + 3, f.l.x); io.st_out(L.dof + 4, f.l.y); io.st_out(L.dof + 5, f.l.z);
            }
            else
               io.st_out(L.dof, L.jt == MB_REVOLUTE ? f.a.z : f.l.z);
            if (L.parent >= 0) contrib = force_to_parent(X, f);
         }
         if (lev > 0)
            for (int r = 0; r < maxc; r++)
            {
               if (mine && rank == r) col_add(PV, contrib);
               __syncthreads();
            }
      }
      __syncthreads(); // the slots are reused by the next state
   }
}

// ------------------------------------------------------------------------------------------------ ABA
// shared: V[6][T] (twists in pass one, accelerations in pass three), K[27][T] (articulated inertia + bias wrench accumulators), ipar[T]
template <bool FEXT> __global__ void __launch_bounds__(kTeamMax) team_aba_kernel(const MbProgram *__restrict__ Pp, const KernelArgs a, int maxc)
{
   extern __shared__ double tsm[];
   const MbProgram &P = *Pp;
   const int T = blockDim.x, t = threadIdx.x;
   int rank;
   const Lane L = team_setup<FEXT>(P, a.consts, reinterpret_cast<int *>(tsm + (6 + kAbaCols) * T), &rank);
   const int nlev = P.max_depth;
   const int pp = L.parent < 0 ? t : L.parent;
   const Col V{tsm + t, T}, PV{tsm + pp, T};
   const Col K{tsm + 6 * T + t, T}, PK{tsm + 6 * T + pp, T};
   for (long long s = blockIdx.x; s < a.n; s += gridDim.x)
   {
      Io io{a.q + s, a.qd + s, a.x + s, a.fext + s, a.out + s, a.ld};
      const XfT<double> X = lane_xf(L, io);
      const SvT<double> vj = lane_joint_vec(L, io.qd, a.ld, true);
      const SvT<double> tauj = lane_joint_vec(L, io.x, a.ld, true); // S * tau
      // ---- pass one (:1085-1127)
      SvT<double> v = sv_zero<double>();
      for (int lev = 0; lev < nlev; lev++)
      {
         if (L.depth == lev)
         {
            const SvT<double> pv = L.parent < 0 ? sv_zero<double>() : col_ld_sv(PV);
            v = motion_to_child(X, pv) + vj;
            col_st(V, v);
         }
         __syncthreads();
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
      {
         double d[kAbaCols];
         abi_flat(IA, pA, d);
#pragma unroll
         for (int k = 0; k < kAbaCols; k++) K[k] = d[k];
      }
      __syncthreads();
      // ---- pass two (:1136-1254), leaves first; g = U / D and k0 = u / D stay in the thread for pass three
      SvT<double> g = sv_zero<double>();
      double k0 = 0.0;
      SvT<double> x6 = sv_zero<double>(); // SixDoF: (I^A)^-1 (tau - p^A)
      for (int lev = nlev - 1; lev >= 0; lev--)
      {
         const bool mine = L.depth == lev;
         double d[kAbaCols];
#pragma unroll
         for (int k = 0; k < kAbaCols; k++) d[k] = 0.0;
         if (mine)
         {
#pragma unroll
            for (int k = 0; k < kAbaCols; k++) d[k] = K[k];
            abi_unflat(d, IA, pA);
            AbiT<double> cK = abi_zero();
            SvT<double> cP = sv_zero<double>();
            if (L.jt == MB_SIXDOF)
            {
               x6 = abi_solve(IA, tauj - pA);
               if (L.parent >= 0) cP = force_to_parent(X, tauj); // the joint transmits nothing but tau
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
                  cK = abi_to_parent(X, Ia);
                  cP = force_to_parent(X, pa);
               }
            }
            abi_flat(cK, cP, d);
         }
         if (lev > 0)
            for (int r = 0; r < maxc; r++)
            {
               if (mine && rank == r && L.parent >= 0)
               {
#pragma unroll
                  for (int k = 0; k < kAbaCols; k++) PK[k] += d[k];
               }
               __syncthreads();
            }
      }
      __syncthreads();
      // ---- pass three (:1259-1310), root first; V carries the accelerations
      for (int lev = 0; lev < nlev; lev++)
      {
         if (L.depth == lev)
         {
            SvT<double> pa = sv_zero<double>();
            if (L.parent < 0)
               pa.l = v3<double>(-a.grav[0], -a.grav[1], -a.grav[2]);
            else
               pa = col_ld_sv(PV);
            const SvT<double> a1 = motion_to_child(X, pa) + cb;
            SvT<double> acc;
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
            col_st(V, acc);
         }
         __syncthreads();
      }
   }
}

// ------------------------------------------------------------------------------------------------ CRBA
// shared: XS[12][T] (joint transforms), IC[10][T] (composite inertia accumulators), then ints jt[T], dof[T], par[T]
template <bool STATE_MAJOR>
__global__ void __launch_bounds__(kTeamMax) team_crba_kernel(const MbProgram *__restrict__ Pp, const KernelArgs a, int maxc, int max_ndof)
{
   extern __shared__ double tsm[];
   const MbProgram &P = *Pp;
   const int T = blockDim.x, t = threadIdx.x;
   int *ijt = reinterpret_cast<int *>(tsm + (kXfCols + kRbiCols) * T), *idof = ijt + T, *ipar = idof + T;
   int rank;
   const Lane L = team_setup<false>(P, a.consts, ipar, &rank);
   ijt[t] = L.jt;
   idof[t] = L.dof;
   const int nlev = P.max_depth, nv = a.nv;
   const long long mstride = STATE_MAJOR ? 1 : a.ld;
   const int pp = L.parent < 0 ? t : L.parent;
   double *XS = tsm;
   const Col IC{tsm + kXfCols * T + t, T}, PIC{tsm + kXfCols * T + pp, T};
   for (long long s = blockIdx.x; s < a.n; s += gridDim.x)
   {
      Io io{a.q + s, nullptr, nullptr, nullptr, nullptr, a.ld};
      double *M = STATE_MAJOR ? a.out + s * (long long)nv * nv : a.out + s;
      // entries coupling joints of unrelated branches are zero (massMatrix.zero(), :296)
      for (int k = t; k < a.n_zero; k += T)
         M[(long long)a.zero_entries[k] * mstride] = 0.0;
      const XfT<double> X = lane_xf(L, io);
      {
         double d[kXfCols];
         m3_to(d, X.R);
         d[9] = X.p.x; d[10] = X.p.y; d[11] = X.p.z;
#pragma unroll
         for (int k = 0; k < kXfCols; k++) XS[k * T + t] = d[k];
         double e[kRbiCols];
         rbi_flat(L.body < 0 ? rbi_zero() : L.I, e);
#pragma unroll
         for (int k = 0; k < kRbiCols; k++) IC[k] = e[k];
      }
      __syncthreads();
      // ---- composite inertias (:648-661), leaves first
      for (int lev = nlev - 1; lev > 0; lev--)
      {
         const bool mine = L.depth == lev;
         double e[kRbiCols];
#pragma unroll
         for (int k = 0; k < kRbiCols; k++) e[k] = 0.0;
         if (mine)
         {
#pragma unroll
            for (int k = 0; k < kRbiCols; k++) e[k] = IC[k];
            rbi_flat(rbi_to_parent(X, rbi_unflat(e)), e);
         }
         for (int r = 0; r < maxc; r++)
         {
            if (mine && rank == r)
            {
#pragma unroll
               for (int k = 0; k < kRbiCols; k++) PIC[k] += e[k];
            }
            __syncthreads();
         }
      }
      RbiT<double> Ic;
      {
         double e[kRbiCols];
#pragma unroll
         for (int k = 0; k < kRbiCols; k++) e[k] = IC[k];
         Ic = rbi_unflat(e);
      }
      // ---- columns: unit momenta F = Ic S (:663-667), diagonal block (:700-707), walk to the root (:772-797); every body walks
      // up its own ancestor chain, the ancestors' transforms and joint data come from shared memory
      const int ndof = L.body < 0 ? 0 : (L.jt == MB_SIXDOF ? 6 : 1);
      for (int col = 0; col < ndof; col++)
      {
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
         if (L.jt == MB_SIXDOF)
         {
            M[(long long)((L.dof + 0) * nv + dc) * mstride] = F.a.x; M[(long long)((L.dof + 1) * nv + dc) * mstride] = F.a.y;
            M[(long long)((L.dof + 2) * nv + dc) * mstride] = F.a.z; M[(long long)((L.dof + 3) * nv + dc) * mstride] = F.l.x;
            M[(long long)((L.dof + 4) * nv + dc) * mstride] = F.l.y; M[(long long)((L.dof + 5) * nv + dc) * mstride] = F.l.z;
         }
         else
            M[(long long)(dc * nv + dc) * mstride] = L.jt == MB_REVOLUTE ? F.a.z : F.l.z;
         int j = t;
         int pj = L.parent;
         while (pj >= 0)
         {
            XfT<double> Xj;
            {
               double d[kXfCols];
#pragma unroll
               for (int k = 0; k < kXfCols; k++) d[k] = XS[k * T + j];
               Xj.R = m3_from(d);
               Xj.p = v3<double>(d[9], d[10], d[11]);
            }
            F = force_to_parent(Xj, F);
            j = pj;
            const int jt_n = ijt[j], dof_n = idof[j];
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
            pj = ipar[j];
         }
      }
      (void)max_ndof;
      __syncthreads(); // the slots are reused by the next state
   }
}

size_t team_smem(int algo, int T)
{
   const int cols = algo == MB_RNEA ? 12 : (algo == MB_ABA ? 6 + kAbaCols : kXfCols + kRbiCols);
   return sizeof(double) * (size_t)cols * T + sizeof(int) * 3 * (size_t)T;
}
} // namespace

int team_threads(const MbProgram &P) { return 32 * ((P.nb + 31) / 32); }

// P: device copy of the traversal program, nb: its body count
cudaError_t launch_team_kernel(int algo, const MbProgram *P, int nb, const KernelArgs &a, int max_children, int max_ndof, int sm_count, cudaStream_t stream)
{
   if (a.n <= 0)
      return cudaSuccess;
   const int T = 32 * ((nb + 31) / 32);
   const size_t sm = team_smem(algo, T);
   // a few resident teams per SM (shared memory: ABA 4 x 128 threads x 33 columns = 135 KB)
   const unsigned grid = (unsigned)std::min<long long>(a.n, (long long)sm_count * (512 / T));
   if (algo == MB_RNEA)
   {
      if (a.fext) team_rnea_kernel<true><<<grid, T, sm, stream>>>(P, a, max_children);
      else team_rnea_kernel<false><<<grid, T, sm, stream>>>(P, a, max_children);
   }
   else if (algo == MB_ABA)
   {
      if (a.fext) team_aba_kernel<true><<<grid, T, sm, stream>>>(P, a, max_children);
      else team_aba_kernel<false><<<grid, T, sm, stream>>>(P, a, max_children);
   }
   else
   {
      if (a.flags & 1u) team_crba_kernel<true><<<grid, T, sm, stream>>>(P, a, max_children, max_ndof);
      else team_crba_kernel<false><<<grid, T, sm, stream>>>(P, a, max_children, max_ndof);
   }
   return cudaGetLastError();
}

cudaError_t team_kernel_attributes(int algo, bool fext, cudaFuncAttributes *attr)
{
   if (algo == MB_RNEA) return fext ? cudaFuncGetAttributes(attr, team_rnea_kernel<true>) : cudaFuncGetAttributes(attr, team_rnea_kernel<false>);
   if (algo == MB_ABA) return fext ? cudaFuncGetAttributes(attr, team_aba_kernel<true>) : cudaFuncGetAttributes(attr, team_aba_kernel<false>);
   return cudaFuncGetAttributes(attr, team_crba_kernel<false>);
}
} // namespace mb
