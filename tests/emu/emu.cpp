// emu.cpp -- kernel-source emulation harness (TEST INFRASTRUCTURE, never shipped or loaded by the product).
// Compiles the exact per-state algorithm templates that the sm_100a kernels inline
// (mecano_b200/csrc/algorithms.cuh) plus the host flattener for the CPU, so that the no-GPU test suite can
// check the kernel mathematics and the traversal programs against the oracle.  The product library
// (libmecano_b200.so) contains no such host path: its entry points launch CUDA kernels or fail.
#include <algorithm>
#include <cstring>
#include <string>
#include <vector>

#include <math.h>

// ---- counting scalar: instantiating the per-state routines with it yields the ALGORITHMIC operation count of the
// local-transform formulation: add / sub / mul / div = 1 flop each (so a fused multiply-add = 2), a
// sin/cos pair counted separately.  Defined before the algorithm headers so that their overload sets see it.
struct Cnt
{
   double v;
   Cnt() : v(0) {}
   Cnt(double x) : v(x) {}
   explicit operator double() const { return v; }
};
struct CntTotals
{
   long add, mul, div, sincos;
};
static thread_local CntTotals g_cnt = {0, 0, 0, 0};
inline Cnt operator+(Cnt a, Cnt b) { g_cnt.add++; return Cnt(a.v + b.v); }
inline Cnt operator-(Cnt a, Cnt b) { g_cnt.add++; return Cnt(a.v - b.v); }
inline Cnt operator*(Cnt a, Cnt b) { g_cnt.mul++; return Cnt(a.v * b.v); }
inline Cnt operator/(Cnt a, Cnt b) { g_cnt.div++; return Cnt(a.v / b.v); }
inline Cnt operator-(Cnt a) { return Cnt(-a.v); }
inline Cnt &operator+=(Cnt &a, Cnt b) { a = a + b; return a; }
inline Cnt &operator-=(Cnt &a, Cnt b) { a = a - b; return a; }
inline Cnt &operator*=(Cnt &a, Cnt b) { a = a * b; return a; }
inline bool operator>(Cnt a, Cnt b) { return a.v > b.v; }
inline bool operator<(Cnt a, Cnt b) { return a.v < b.v; }
namespace mb
{
inline void mb_sincos(Cnt x, Cnt *s, Cnt *c) { g_cnt.sincos++; *s = Cnt(sin(x.v)); *c = Cnt(cos(x.v)); }
inline Cnt mb_rcp(Cnt x) { g_cnt.div++; return Cnt(1.0 / x.v); }
inline Cnt mb_reduce_angle(Cnt x) { return x; }
} // namespace mb

#include "../../mecano_b200/csrc/algorithms.cuh"
#include "../../mecano_b200/csrc/flatten.h"

namespace
{
template <class T> struct CpuCtx
{
   static constexpr bool kM3 = true; // three-DoF joints compiled in (multidof.cuh)
   static constexpr bool kFastQuat = false;
   void warm_q(int) const {}
   void warm_qd(int) const {}
   void warm_x(int) const {}
   const double *q, *qd, *x, *fext;
   double *out, *M;
   long ld, s;
   long ldd, ldx; // row strides of qd / x (0: a single row of zeros, the way the launcher handles the RNEA flags)
   int nv;
   T *stk, *aux, *rec;
   const T *consts;
   T ld_q(int r) const { return (T)q[r * ld + s]; }
   T ld_qd(int r) const { return (T)qd[r * ldd + s]; }
   T ld_x(int r) const { return (T)x[r * ldx + s]; }
   T ld_fext(int b, int k) const { return (T)fext[(6 * b + k) * ld + s]; }
   void st_out(int r, T v) { out[r * ld + s] = (double)v; }
   double *acc_out = nullptr, *wr_out = nullptr;
   const double *x2 = nullptr;
   double *cmm = nullptr, *com = nullptr, *rootw = nullptr;
   double *Cm = nullptr; // Coriolis matrix, entry-major like M
   void st_C(int e, T v) { Cm[(long)e * ld + s] = (double)v; }
   void zero_fill_mc_part(int k, int parts)
   {
      const int n8 = (int)zl->size() / 8, k1 = std::min(n8, (k + 1) * parts);
      for (int i = k * parts; i < k1; i++)
         for (int j = 0; j < 8; j++) { st_M((*zl)[8 * i + j], (T)0); st_C((*zl)[8 * i + j], (T)0); }
   }
   bool has_rootw() const { return rootw != nullptr; }
   bool com_only() const { return M == nullptr; } // gpu_ctx.cuh: a by-product CRBA launch without a matrix
   void st_cmm(int row, T v) { cmm[(long)row * ld + s] = (double)v; }
   void add_com(int r, T v) { com[r * ld + s] += (double)v; }
   void add_rootw(int r, T v) { rootw[r * ld + s] += (double)v; }
   T ld_x2(int r) const { return (T)x2[r * ld + s]; }
   bool has_fext() const { return fext != nullptr; }
   bool has_acc() const { return acc_out != nullptr; }
   bool has_wr() const { return wr_out != nullptr; }
   void st_acc(int b, int k, T v) { acc_out[(6 * b + k) * ld + s] = (double)v; }
   void st_wr(int b, int k, T v) { wr_out[(6 * b + k) * ld + s] = (double)v; }
   void st_M(int e, T v) { M[(long)e * ld + s] = (double)v; }
   int n_dofs() const { return nv; }
   const std::vector<uint16_t> *zl;
   void zero_fill() { for (uint16_t e : *zl) st_M(e, (T)0); }
   int zero_parts(int nops) const { return ((int)zl->size() / 8 + nops - 1) / nops; }
   void zero_fill_part(int k, int parts)
   {
      const int n8 = (int)zl->size() / 8, k1 = std::min(n8, (k + 1) * parts);
      for (int i = k * parts; i < k1; i++)
         for (int j = 0; j < 8; j++) st_M((*zl)[8 * i + j], (T)0);
   }
   void stk_ld2(int slot2, int j, T &a, T &b) const { a = stk[2 * (slot2 + j)]; b = stk[2 * (slot2 + j) + 1]; }
   void stk_st2(int slot2, int j, T a, T b) { stk[2 * (slot2 + j)] = a; stk[2 * (slot2 + j) + 1] = b; }
   // split view of a slot (MbOp2::wslot / nslot): the wide area behind the combined stack so that both index sets are exercised
   T *wide, *narrow;
   void acc_ld(int, int wslot, T &x0, T &x1, T &x2, T &x3, T &x4, T &x5) const
   {
      const T *p = wide + 2 * wslot;
      x0 = p[0]; x1 = p[1]; x2 = p[2]; x3 = p[3]; x4 = p[4]; x5 = p[5];
   }
   void acc_st(int, int wslot, T x0, T x1, T x2, T x3, T x4, T x5)
   {
      T *p = wide + 2 * wslot;
      p[0] = x0; p[1] = x1; p[2] = x2; p[3] = x3; p[4] = x4; p[5] = x5;
   }
   void jp_ld2(int, int nslot, int j, T &a, T &b) const { a = narrow[2 * (nslot + j)]; b = narrow[2 * (nslot + j) + 1]; }
   void jp_st2(int, int nslot, int j, T a, T b) { narrow[2 * (nslot + j)] = a; narrow[2 * (nslot + j) + 1] = b; }
   T ring[4][3];
   void pf_issue(int stage, int cfg, int dof, int mask)
   {
      const T nan = (T)(0.0 / 0.0);
      ring[stage][0] = (mask & 1) ? ld_q(cfg) : nan;
      ring[stage][1] = (mask & 2) ? ld_qd(dof) : nan;
      ring[stage][2] = (mask & 4) ? ld_x(dof) : nan;
   }
   void pf_commit() {}
   // cache warming: nothing to do on the host
   void pf_six(int, int, int) const {}
   void pf_six_next_tile(int, int, int) const {}
   void rec_prefetch_far(int, int, int, int) const {}
   void stk_fence() const {}
   void op_sync(int) const {}
   template <int N> void pf_wait() {}
   T pf_ld(int stage, int j) const { return ring[stage][j]; }
   T stk_ld(int i) const { return stk[i]; }
   void stk_st(int i, T v) { stk[i] = v; }
   T aux_ld(int i) const { return aux[i]; }
   void aux_st(int i, T v) { aux[i] = v; }
   T rec_ld(int i) const { return rec[i]; }
   void rec_st2(int i2, T a, T b) { rec[2 * i2] = a; rec[2 * i2 + 1] = b; }
   void rec_ld2(int i2, T &a, T &b) const { a = rec[2 * i2]; b = rec[2 * i2 + 1]; }
   T ring3[4][2 + MB_ABA_REC];
   void pf3_issue(int stage, int cfg, int dof, int rec2, int mask)
   {
      const T nan = (T)(0.0 / 0.0);
      ring3[stage][0] = (mask & 1) ? ld_q(cfg) : nan;
      ring3[stage][1] = (mask & 2) ? ld_qd(dof) : nan;
      for (int j = 0; j < MB_ABA_REC; j++)
         ring3[stage][2 + j] = rec[2 * rec2 + j];
   }
   // the GPU context tells L2 that the record is dead; here: poison it, so that a second read shows up as NaN
   void rec_discard(int rec2)
   {
      for (int j = 0; j < MB_ABA_REC; j++)
         rec[2 * rec2 + j] = (T)(0.0 / 0.0);
   }
   void pf3_ld2(int stage, int row, T &a, T &b) const { a = ring3[stage][2 * row]; b = ring3[stage][2 * row + 1]; }
   void pass_fence() const {}
   const T *cst(int b) const { return consts + (size_t)b * MB_CONST_STRIDE; }
};

template <class T>
int run(int algo, const mecano_b200_tree_desc *d, const double *g, long n, long ld, const double *q, const double *qd, const double *x,
        const double *fext, double *out, unsigned flags, char *err, int errlen, double *acc_out = nullptr, double *wr_out = nullptr,
        const int32_t *accel_source = nullptr, const double *x2 = nullptr, double *cmm = nullptr, double *com = nullptr, double *rootw = nullptr, double *Cm = nullptr)
{
   mb::FlatTree ft;
   std::string e;
   int rc = mb::flatten_tree(d, ft, e);
   if (rc != 0)
   {
      if (err) { std::strncpy(err, e.c_str(), errlen - 1); err[errlen - 1] = 0; }
      return rc;
   }
   std::vector<std::pair<int, int>> effort_runs;
   const int n_locked = mb::apply_source_modes(ft, accel_source, effort_runs);
   const MbProgram &P = ft.prog[algo];
   std::vector<T> consts(ft.consts.begin(), ft.consts.end());
   // poison the work areas so that a read-before-write shows up as NaN
   const T nan = (T)(0.0 / 0.0);
   std::vector<T> stk(std::max(P.stack_doubles, 2 * P.stack2) + 2, nan), aux(P.aux_doubles + 1, nan), rec(P.rec_doubles + 64, nan);
   std::vector<T> wide(2 * P.wstack2 + 2, nan), narrow(2 * P.nstack2 + 2, nan);
   const T grav[3] = {(T)g[0], (T)g[1], (T)g[2]};
   const std::vector<double> zero_row((size_t)std::max(n, 1l), 0.0);
   for (long s = 0; s < n; s++)
   {
      std::fill(stk.begin(), stk.end(), nan);
      std::fill(aux.begin(), aux.end(), nan);
      std::fill(rec.begin(), rec.end(), nan);
      std::fill(wide.begin(), wide.end(), nan);
      std::fill(narrow.begin(), narrow.end(), nan);
      CpuCtx<T> c{q, qd, x, fext, out, out, ld, s, ld, ld, ft.nv, stk.data(), aux.data(), rec.data(), consts.data()};
      if (algo == MB_RNEA && (flags & 1u)) { c.qd = zero_row.data(); c.ldd = 0; } // api.cu: run()
      if (algo == MB_RNEA && (flags & 2u)) { c.x = zero_row.data(); c.ldx = 0; }
      c.wide = wide.data();
      c.narrow = narrow.data();
      c.acc_out = acc_out;
      c.wr_out = wr_out;
      c.x2 = x2;
      c.cmm = cmm; c.com = com; c.rootw = rootw;
      c.Cm = Cm;
      c.zl = &ft.zero_entries;
      if (algo == MB_RNEA)
      {
         if (fext || acc_out || wr_out || rootw) mb::rnea_state<T, CpuCtx<T>, true>(P, c, grav);
         else mb::rnea_state<T, CpuCtx<T>, false>(P, c, grav);
      }
      else if (algo == MB_ABA)
      {
         if (fext || n_locked > 0) mb::aba_state<T, CpuCtx<T>, true>(P, c, grav);
         else mb::aba_state<T, CpuCtx<T>, false>(P, c, grav);
      }
      else if (algo == MB_CORIOLIS)
         mb::coriolis_state<T, CpuCtx<T>>(P, c);
      else if (cmm || com)
         mb::crba_state<T, CpuCtx<T>, true>(P, c);
      else if (flags & 4u) // MECANO_B200_CRBA_PACKED: out = [packed rows][ld]
         mb::crba_state<T, CpuCtx<T>, false, true>(P, c);
      else
         mb::crba_state<T, CpuCtx<T>>(P, c);
   }
   return 0;
}
} // namespace

extern "C" int emu_run(int algo, int fp32, const mecano_b200_tree_desc *d, const double *g, long n, long ld, const double *q, const double *qd,
                       const double *x, const double *fext, double *out, unsigned flags, char *err, int errlen)
{
   return fp32 ? run<float>(algo, d, g, n, ld, q, qd, x, fext, out, flags, err, errlen)
               : run<double>(algo, d, g, n, ld, q, qd, x, fext, out, flags, err, errlen);
}

// RNEA with its by-products (body accelerations in CoM frames, joint wrenches in frameAfterJoint), rows [6 * w + c]
extern "C" int emu_rnea_full(const mecano_b200_tree_desc *d, const double *g, long n, long ld, const double *q, const double *qd, const double *x,
                             const double *fext, double *tau, double *acc, double *wr, unsigned flags, char *err, int errlen)
{
   return run<double>(MB_RNEA, d, g, n, ld, q, qd, x, fext, tau, flags, err, errlen, acc, wr);
}

// ABA with joints in ACCELERATION_SOURCE mode (passes one to three; accel_source [n_bodies] in the order of the description)
extern "C" int emu_aba_sources(const mecano_b200_tree_desc *d, const double *g, long n, long ld, const double *q, const double *qd, const double *tau,
                               const double *qdd_in, const double *fext, const int32_t *accel_source, double *qdd, char *err, int errlen)
{
   return run<double>(MB_ABA, d, g, n, ld, q, qd, tau, fext, qdd, 0u, err, errlen, nullptr, nullptr, accel_source, qdd_in);
}

// CRBA with by-products: M [nv * nv][ld], centroidal momentum matrix [6 * nv][ld] in the root frame, com [4][ld] = (mass * CoM, mass)
// accumulated over the root's children (zeroed here, as the C ABI does before the launch)
extern "C" int emu_crba_centroidal(const mecano_b200_tree_desc *d, long n, long ld, const double *q, double *M, double *cmm, double *com, char *err,
                                   int errlen)
{
   const double g[3] = {0, 0, 0};
   std::fill(com, com + 4 * ld, 0.0);
   return run<double>(MB_CRBA, d, g, n, ld, q, nullptr, nullptr, nullptr, M, 0u, err, errlen, nullptr, nullptr, nullptr, nullptr, cmm, com, nullptr);
}

// the centre-of-mass-only launch of the by-product CRBA routine (mecano_b200_center_of_mass): no matrix, no momentum matrix (a
// write through either would fault here), com [4][ld] = (mass * CoM, mass)
extern "C" int emu_center_of_mass(const mecano_b200_tree_desc *d, long n, long ld, const double *q, double *com, char *err, int errlen)
{
   const double g[3] = {0, 0, 0};
   std::fill(com, com + 4 * ld, 0.0);
   return run<double>(MB_CRBA, d, g, n, ld, q, nullptr, nullptr, nullptr, nullptr, 0u, err, errlen, nullptr, nullptr, nullptr, nullptr, nullptr, com, nullptr);
}

// mass matrix M and Coriolis matrix C, both [nv * nv][ld] entry-major
extern "C" int emu_coriolis(const mecano_b200_tree_desc *d, long n, long ld, const double *q, const double *qd, double *M, double *C, char *err,
                            int errlen)
{
   const double g[3] = {0, 0, 0};
   return run<double>(MB_CORIOLIS, d, g, n, ld, q, qd, nullptr, nullptr, M, 0u, err, errlen, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, C);
}

// RNEA with zero joint accelerations and no gravity, its wrench at the root summed into rootw [6][ld] (root frame): the centroidal
// convective term about the origin of the root frame
extern "C" int emu_rnea_root_wrench(const mecano_b200_tree_desc *d, long n, long ld, const double *q, const double *qd, double *tau, double *rootw,
                                    char *err, int errlen)
{
   const double g[3] = {0, 0, 0};
   std::fill(rootw, rootw + 6 * ld, 0.0);
   return run<double>(MB_RNEA, d, g, n, ld, q, qd, qd, nullptr, tau, 2u, err, errlen, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, rootw);
}

// Algorithmic operation counts of one state: out5 = {add, mul, div, sincos, flops = add + mul + div}
extern "C" int emu_count_flops(int algo, const mecano_b200_tree_desc *d, const double *q, const double *qd, const double *x, long *out5)
{
   const double g[3] = {0, 0, -9.81};
   mb::FlatTree ft;
   std::string e;
   int rc = mb::flatten_tree(d, ft, e);
   if (rc != 0) return rc;
   std::vector<double> out((size_t)ft.nv * ft.nv + ft.nv, 0.0);
   g_cnt = {0, 0, 0, 0};
   rc = run<Cnt>(algo, d, g, 1, 1, q, qd, x, nullptr, out.data(), 0, nullptr, 0);
   out5[0] = g_cnt.add; out5[1] = g_cnt.mul; out5[2] = g_cnt.div; out5[3] = g_cnt.sincos;
   out5[4] = g_cnt.add + g_cnt.mul + g_cnt.div;
   return rc;
}

// packed mass-matrix layout: number of packed rows and the dense (row, col) of each (row / col may be NULL)
extern "C" int emu_packed_index(const mecano_b200_tree_desc *d, int32_t *row, int32_t *col)
{
   mb::FlatTree ft;
   std::string e;
   int rc = mb::flatten_tree(d, ft, e);
   if (rc != 0) return rc;
   if (row) std::copy(ft.packed_row.begin(), ft.packed_row.end(), row);
   if (col) std::copy(ft.packed_col.begin(), ft.packed_col.end(), col);
   return (int)ft.packed_row.size();
}

// the run table of a traversal program (program.h: MbRun): kinds and lengths; pass3 != 0: the ABA pass-three list
extern "C" int emu_program_runs(const mecano_b200_tree_desc *d, int algo, int pass3, int *kinds, int *lens, int cap)
{
   mb::FlatTree ft;
   std::string e;
   int rc = mb::flatten_tree(d, ft, e);
   if (rc != 0) return -1;
   const MbProgram &P = ft.prog[algo];
   const int n = pass3 ? P.nruns3 : P.nruns;
   for (int r = 0; r < n && r < cap; r++)
   {
      const MbRun &R = pass3 ? P.run3[r] : P.run[r];
      kinds[r] = R.kind;
      lens[r] = R.n;
   }
   return n;
}

extern "C" int emu_program_info(const mecano_b200_tree_desc *d, int algo, int *out8)
{
   mb::FlatTree ft;
   std::string e;
   int rc = mb::flatten_tree(d, ft, e);
   if (rc != 0) return rc;
   const MbProgram &P = ft.prog[algo];
   out8[0] = P.nb; out8[1] = P.nops; out8[2] = 2 * P.stack2; out8[3] = P.aux_doubles; out8[4] = P.rec_doubles; out8[5] = P.max_depth;
   out8[6] = P.nv; out8[7] = P.nq;
   return 0;
}

// ---- the spatial-algebra primitives of the kernels (spatial.cuh), one call each, for unit tests against dense 6 x 6 numpy
// (the reference tests its own unrolled helpers the same way: ArticulatedBodyInertiaTest.java:21-64, SpatialInertiaBasicsTest,
// MecanoToolsTest.testComputeDynamicWrench).  Flat layouts: X = R row-major (9), p (3); spatial vector = angular (3), linear (3);
// rigid-body inertia = I (xx xy xz yy yz zz), h (3), m; articulated inertia = A (6, symmetric), C (9, row-major), L (6).
namespace
{
using D = double;
mb::XfT<D> rd_xf(const D *p)
{
   mb::XfT<D> X;
   X.R.xx = p[0]; X.R.xy = p[1]; X.R.xz = p[2]; X.R.yx = p[3]; X.R.yy = p[4]; X.R.yz = p[5]; X.R.zx = p[6]; X.R.zy = p[7]; X.R.zz = p[8];
   X.p = mb::v3<D>(p[9], p[10], p[11]);
   return X;
}
mb::SvT<D> rd_sv(const D *p)
{
   mb::SvT<D> v;
   v.a = mb::v3<D>(p[0], p[1], p[2]);
   v.l = mb::v3<D>(p[3], p[4], p[5]);
   return v;
}
mb::S3T<D> rd_s3(const D *p)
{
   mb::S3T<D> s;
   s.xx = p[0]; s.xy = p[1]; s.xz = p[2]; s.yy = p[3]; s.yz = p[4]; s.zz = p[5];
   return s;
}
mb::RbiT<D> rd_rbi(const D *p)
{
   mb::RbiT<D> I;
   I.I = rd_s3(p);
   I.h = mb::v3<D>(p[6], p[7], p[8]);
   I.m = p[9];
   return I;
}
mb::AbiT<D> rd_abi(const D *p)
{
   mb::AbiT<D> I;
   I.A = rd_s3(p);
   I.C.xx = p[6]; I.C.xy = p[7]; I.C.xz = p[8]; I.C.yx = p[9]; I.C.yy = p[10]; I.C.yz = p[11]; I.C.zx = p[12]; I.C.zy = p[13]; I.C.zz = p[14];
   I.L = rd_s3(p + 15);
   return I;
}
void wr_sv(D *o, const mb::SvT<D> &v) { o[0] = v.a.x; o[1] = v.a.y; o[2] = v.a.z; o[3] = v.l.x; o[4] = v.l.y; o[5] = v.l.z; }
void wr_s3(D *o, const mb::S3T<D> &s) { o[0] = s.xx; o[1] = s.xy; o[2] = s.xz; o[3] = s.yy; o[4] = s.yz; o[5] = s.zz; }
void wr_abi(D *o, const mb::AbiT<D> &I)
{
   wr_s3(o, I.A);
   o[6] = I.C.xx; o[7] = I.C.xy; o[8] = I.C.xz; o[9] = I.C.yx; o[10] = I.C.yy; o[11] = I.C.yz; o[12] = I.C.zx; o[13] = I.C.zy; o[14] = I.C.zz;
   wr_s3(o + 15, I.L);
}
} // namespace

extern "C" int emu_spatial(int op, const double *in, double *out)
{
   using namespace mb;
   switch (op)
   {
      case 0: wr_sv(out, motion_to_child(rd_xf(in), rd_sv(in + 12))); return 6;
      case 1: wr_sv(out, force_to_parent(rd_xf(in), rd_sv(in + 12))); return 6;
      case 2:
      {
         const RbiT<D> r = rbi_to_parent(rd_xf(in), rd_rbi(in + 12));
         wr_s3(out, r.I);
         out[6] = r.h.x; out[7] = r.h.y; out[8] = r.h.z; out[9] = r.m;
         return 10;
      }
      case 3: wr_abi(out, abi_to_parent<D, 0>(rd_xf(in), rd_abi(in + 12))); return 21;
      case 4: wr_sv(out, newton_euler(rd_s3(in), v3<D>(in[6], in[7], in[8]), in[9], rd_sv(in + 10), rd_sv(in + 16))); return 6;
      case 5: wr_sv(out, abi_solve(rd_abi(in), rd_sv(in + 21))); return 6;
      case 6: wr_sv(out, mul(rd_abi(in), rd_sv(in + 21))); return 6;
      case 7: wr_sv(out, mul(rd_rbi(in), rd_sv(in + 10))); return 6;
      case 8: wr_sv(out, cross_motion(rd_sv(in), rd_sv(in + 6))); return 6;
      case 9: wr_sv(out, cross_force(rd_sv(in), rd_sv(in + 6))); return 6;
      case 10: wr_abi(out, abi_downdate<D, 0>(rd_abi(in), rd_sv(in + 21), rd_sv(in + 27))); return 21;
      // the zero-structured pair the one-DoF joints use: downdate along z (1: revolute, 2: prismatic), then the congruence that
      // skips the zero row / column; in = X (12), IA (21), U (6), g (6); out = downdated (21), transformed (21)
      case 11:
      {
         const AbiT<D> d = abi_downdate<D, 1>(rd_abi(in + 12), rd_sv(in + 33), rd_sv(in + 39));
         wr_abi(out, d);
         wr_abi(out + 21, abi_to_parent<D, 1>(rd_xf(in), d));
         return 42;
      }
      case 12:
      {
         const AbiT<D> d = abi_downdate<D, 2>(rd_abi(in + 12), rd_sv(in + 33), rd_sv(in + 39));
         wr_abi(out, d);
         wr_abi(out + 21, abi_to_parent<D, 2>(rd_xf(in), d));
         return 42;
      }
      case 13: wr_sv(out, cross_force_add(rd_sv(in), rd_sv(in + 6), rd_sv(in + 12))); return 6;
      case 14: wr_abi(out, abi_add_rbi(rd_abi(in), rd_rbi(in + 21))); return 21;
      default: return -1;
   }
}

// the kernels' branch-free sin/cos (jointmath.cuh: mb_sincos, with the large-angle handling of the ops around it) on n angles
extern "C" void emu_sincos(long n, const double *x, double *s, double *c)
{
   for (long i = 0; i < n; i++)
   {
      mb::mb_sincos(mb::mb_reduce_angle(x[i]), s + i, c + i);
      // the off-critical-path form of the same safeguard (RNEA / ABA thread-per-state kernels): fast path on the raw angle, redone if large
      double s2, c2;
      mb::mb_sincos(x[i], &s2, &c2);
      if (mb::mb_angle_large(x[i]))
         mb::mb_sincos_redo(x[i], s2, c2);
      if (!(fabs(s2 - s[i]) <= 4.0e-16 && fabs(c2 - c[i]) <= 4.0e-16))
         s[i] = c[i] = 0.0 / 0.0; // the two forms disagree: poison, the test fails
   }
}
