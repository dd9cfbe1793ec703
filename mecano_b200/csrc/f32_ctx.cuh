// f32_ctx.cuh -- the optional fp32 variant (north star: "an optional fp32 variant is reported separately with its stated
// tolerance"): the same per-state routines instantiated with T = float over a float view of the fp64 context.
//
// What changes: the arithmetic (FP32 pipe: twice the lanes of the FP64 pipe, half the registers per value), the constant
// records (staged in shared memory as floats) and the per-state stack, which holds floats in the cells of the fp64 layout
// (shared memory: the low half of each double2; tensor memory: six of the twelve columns of a wide slot), so the stack
// traffic needs no conversion.  What does not: the C ABI and the buffers in HBM stay fp64 (q, qd, tau, M ... are converted at
// the load / store), and so do the branch save area and the ABA records, whose accessors convert on the way.
// Accuracy (tests/test_gpu_parity.py::test_fp32_variant): RNEA / CRBA ~1e-5 relative, ABA ~1e-3 on the test trees (the
// articulated-inertia recursion amplifies rounding with depth); tolerances 2e-4 / 2e-4 / 5e-2 as in the emulation test.
#pragma once
#include "gpu_ctx.cuh"

// 1: the per-state stack holds floats (no conversion at the access); 0: it holds doubles, converted at the access
#ifndef MB_F32_NATIVE_STACK
#define MB_F32_NATIVE_STACK 1
#endif

namespace mb
{
struct SmemCstF
{
   unsigned a;
};
__device__ __forceinline__ void cst_ld2(const SmemCstF C, int i2, float &a, float &b)
{
   asm("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(a), "=f"(b) : "r"(C.a + 8u * (unsigned)i2));
}

template <class Ctx> struct F32Ctx
{
   static constexpr bool kM3 = false; // (api.cu refuses the fp32 variant for trees with three-DoF joints)
   Ctx &c;
   __device__ __forceinline__ explicit F32Ctx(Ctx &ctx) : c(ctx) {}
   // ---- inputs / outputs (fp64 in HBM)
   __device__ __forceinline__ float ld_q(int r) const { return (float)c.ld_q(r); }
   __device__ __forceinline__ float ld_qd(int r) const { return (float)c.ld_qd(r); }
   __device__ __forceinline__ float ld_x(int r) const { return (float)c.ld_x(r); }
   __device__ __forceinline__ void st_out(int r, float v) { c.st_out(r, (double)v); }
   __device__ __forceinline__ void st_M(int e, float v) const { c.st_M(e, (double)v); }
   __device__ __forceinline__ int n_dofs() const { return c.n_dofs(); }
   __device__ __forceinline__ void zero_fill() const { c.zero_fill(); }
   __device__ __forceinline__ int zero_parts(int nops) const { return c.zero_parts(nops); }
   __device__ __forceinline__ void zero_fill_part(int k, int parts) const { c.zero_fill_part(k, parts); }
   // ---- the optional buffers belong to the fp64 "general" instantiation only
   __device__ __forceinline__ bool has_fext() const { return false; }
   __device__ __forceinline__ bool has_acc() const { return false; }
   __device__ __forceinline__ bool has_wr() const { return false; }
   __device__ __forceinline__ bool has_rootw() const { return false; }
   __device__ __forceinline__ float ld_fext(int, int) const { return 0.0f; }
   __device__ __forceinline__ float ld_x2(int) const { return 0.0f; }
   __device__ __forceinline__ void st_acc(int, int, float) {}
   __device__ __forceinline__ void st_wr(int, int, float) {}
   __device__ __forceinline__ void st_cmm(int, float) {}
   __device__ __forceinline__ void add_com(int, float) {}
   __device__ __forceinline__ void add_rootw(int, float) {}
#if MB_F32_NATIVE_STACK
   // ---- per-state stack: floats in the cells of the fp64 layout (same addresses, half of each cell used)
   static constexpr int BLOCK = Ctx::kBlock, TM = Ctx::kTM;
   static_assert(!Ctx::kPartial, "the fp32 variant is not compiled for the partial-TMEM block size");
   static __device__ __forceinline__ void lds2f(unsigned addr, float &a, float &b) { asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(a), "=f"(b) : "r"(addr)); }
   static __device__ __forceinline__ void sts2f(unsigned addr, float a, float b) { asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(a), "f"(b)); }
   __device__ __forceinline__ void stk_ld2(int slot2, int j, float &a, float &b) const { lds2f(c.sb + (unsigned)((slot2 + j) * (BLOCK * 16)), a, b); }
   __device__ __forceinline__ void stk_st2(int slot2, int j, float a, float b) { sts2f(c.sb + (unsigned)((slot2 + j) * (BLOCK * 16)), a, b); }
   __device__ __forceinline__ void acc_ld(int slot2, int wslot, float &x0, float &x1, float &x2, float &x3, float &x4, float &x5) const
   {
      if (TM > 0)
      {
         const unsigned t = c.tm0 + 4u * (unsigned)wslot;
         unsigned r[6];
#pragma unroll
         for (int i = 0; i < 3; i++)
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(r[2 * i]), "=r"(r[2 * i + 1]) : "r"(t + 2u * i));
         asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]));
         x0 = __uint_as_float(r[0]); x1 = __uint_as_float(r[1]); x2 = __uint_as_float(r[2]);
         x3 = __uint_as_float(r[3]); x4 = __uint_as_float(r[4]); x5 = __uint_as_float(r[5]);
      }
      else
      {
         const unsigned t = c.sb + (unsigned)(slot2 * (BLOCK * 16));
         lds2f(t, x0, x1);
         lds2f(t + BLOCK * 16, x2, x3);
         lds2f(t + 2 * BLOCK * 16, x4, x5);
      }
   }
   __device__ __forceinline__ void acc_st(int slot2, int wslot, float x0, float x1, float x2, float x3, float x4, float x5)
   {
      if (TM > 0)
      {
         const unsigned t = c.tm0 + 4u * (unsigned)wslot;
         const float x[6] = {x0, x1, x2, x3, x4, x5};
#pragma unroll
         for (int i = 0; i < 3; i++)
            asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(t + 2u * i), "r"(__float_as_uint(x[2 * i])), "r"(__float_as_uint(x[2 * i + 1])));
      }
      else
      {
         const unsigned t = c.sb + (unsigned)(slot2 * (BLOCK * 16));
         sts2f(t, x0, x1);
         sts2f(t + BLOCK * 16, x2, x3);
         sts2f(t + 2 * BLOCK * 16, x4, x5);
      }
   }
   __device__ __forceinline__ void jp_ld2(int slot2, int nslot, int j, float &a, float &b) const
   {
      lds2f(c.sb + (unsigned)((TM > 0 ? nslot + j : slot2 + 3 + j) * (BLOCK * 16)), a, b);
   }
   __device__ __forceinline__ void jp_st2(int slot2, int nslot, int j, float a, float b)
   {
      sts2f(c.sb + (unsigned)((TM > 0 ? nslot + j : slot2 + 3 + j) * (BLOCK * 16)), a, b);
   }
#else
   // ---- per-state stack: fp64 storage, converted at the access
   __device__ __forceinline__ void stk_ld2(int slot2, int j, float &a, float &b) const
   {
      double x, y;
      c.stk_ld2(slot2, j, x, y);
      a = (float)x; b = (float)y;
   }
   __device__ __forceinline__ void stk_st2(int slot2, int j, float a, float b) { c.stk_st2(slot2, j, (double)a, (double)b); }
   __device__ __forceinline__ void acc_ld(int slot2, int wslot, float &x0, float &x1, float &x2, float &x3, float &x4, float &x5) const
   {
      double d0, d1, d2, d3, d4, d5;
      c.acc_ld(slot2, wslot, d0, d1, d2, d3, d4, d5);
      x0 = (float)d0; x1 = (float)d1; x2 = (float)d2; x3 = (float)d3; x4 = (float)d4; x5 = (float)d5;
   }
   __device__ __forceinline__ void acc_st(int slot2, int wslot, float x0, float x1, float x2, float x3, float x4, float x5)
   {
      c.acc_st(slot2, wslot, (double)x0, (double)x1, (double)x2, (double)x3, (double)x4, (double)x5);
   }
   __device__ __forceinline__ void jp_ld2(int slot2, int nslot, int j, float &a, float &b) const
   {
      double x, y;
      c.jp_ld2(slot2, nslot, j, x, y);
      a = (float)x; b = (float)y;
   }
   __device__ __forceinline__ void jp_st2(int slot2, int nslot, int j, float a, float b) { c.jp_st2(slot2, nslot, j, (double)a, (double)b); }
#endif
   // ---- save area, records: fp64 storage, converted at the access
   __device__ __forceinline__ float aux_ld(int i) const { return (float)c.aux_ld(i); }
   __device__ __forceinline__ void aux_st(int i, float v) { c.aux_st(i, (double)v); }
   __device__ __forceinline__ void rec_st2(int i2, float a, float b) { c.rec_st2(i2, (double)a, (double)b); }
   // ---- prefetch rings
   __device__ __forceinline__ void pf_issue(int stage, int cfg, int dof, int mask) const { c.pf_issue(stage, cfg, dof, mask); }
   __device__ __forceinline__ void pf_commit() const { c.pf_commit(); }
   template <int N> __device__ __forceinline__ void pf_wait() const { c.template pf_wait<N>(); }
   __device__ __forceinline__ float pf_ld(int stage, int j) const { return (float)c.pf_ld(stage, j); }
   static constexpr bool kFastQuat = false;
   __device__ __forceinline__ void warm_q(int r) const { c.warm_q(r); }
   __device__ __forceinline__ void warm_qd(int r) const { c.warm_qd(r); }
   __device__ __forceinline__ void warm_x(int r) const { c.warm_x(r); }
   __device__ __forceinline__ void pf3_issue(int stage, int cfg, int dof, int rec2, int mask) const { c.pf3_issue(stage, cfg, dof, rec2, mask); }
   __device__ __forceinline__ void pf3_ld2(int stage, int row, float &a, float &b) const
   {
      double x, y;
      c.pf3_ld2(stage, row, x, y);
      a = (float)x; b = (float)y;
   }
   __device__ __forceinline__ void rec_discard(int rec2) const { c.rec_discard(rec2); }
   __device__ __forceinline__ void rec_ld2(int i2, float &a, float &b) const
   {
      double x, y;
      c.rec_ld2(i2, x, y);
      a = (float)x; b = (float)y;
   }
   __device__ __forceinline__ void pf_six(int cfg, int dof, int mask) const { c.pf_six(cfg, dof, mask); }
   __device__ __forceinline__ void pf_six_next_tile(int cfg, int dof, int mask) const { c.pf_six_next_tile(cfg, dof, mask); }
   __device__ __forceinline__ void rec_prefetch_far(int rec2, int cfg, int dof, int mask) const { c.rec_prefetch_far(rec2, cfg, dof, mask); }
   __device__ __forceinline__ void pass_fence() const { c.pass_fence(); }
   __device__ __forceinline__ void stk_fence() const { c.stk_fence(); }
   __device__ __forceinline__ void op_sync(int k) const { c.op_sync(k); }
   // ---- constant records: floats, staged by thread_kernel_f32 at the front of shared memory
   __device__ __forceinline__ SmemCstF cst(int b) const { return SmemCstF{c.cb + (unsigned)(b * (MB_CONST_STRIDE * 4))}; }
};
} // namespace mb
