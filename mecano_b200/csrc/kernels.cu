// kernels.cu -- sm_100a kernels: one thread per state (thread-per-state variant).
//
// Layout of one block:
//   * the traversal program and topology tables arrive as a __grid_constant__ kernel parameter, i.e.
//     they live in the constant bank and are read with uniform (warp-wide) LDC;
//   * the per-body constant records (fixed transforms, inertias) are staged once into shared memory
//     and read as broadcasts;
//   * each thread owns one state: its spatial quantities stay in fp64 registers, and the data that
//     must survive from the downward to the upward sweep lives in a per-thread stack in shared memory,
//     laid out state-minor (stack[slot * blockDim + tid]) so that a warp touches 32 consecutive
//     doubles: conflict-free;
//   * q / qd / qdd / tau / wrench buffers are DoF-major, state-minor in HBM, so every global access of
//     a warp is one fully used 256-byte segment.
// No tensor cores: the recursion is not a dense contraction.
#include <algorithm>
#include <cstdio>
#include <type_traits>

#include "algorithms.cuh"
#include "kernels.h"

// the dynamic shared memory of a block: [constant records | per-thread stack, state-minor]
extern __shared__ double mb_smem[];

namespace mb
{
namespace
{
// BLOCK (threads per block = stack stride) is a compile-time constant so that stack addresses are base + immediate;
// shared memory is addressed through the mb_smem symbol so that the compiler emits LDS/STS (a pointer kept in a
// struct degrades to generic LD/ST).
// Per-thread context: per-thread base pointers (one IMAD.WIDE per global access), stack of double2 (LDS.128 / STS.128)
template <int BLOCK> struct GpuCtx2
{
   const char *qb, *qdb, *xb, *fb;
   char *ob;
   unsigned ld8; // bytes between consecutive rows (the launcher keeps ld * 8 < 2^32)
   int stk0;     // index (double2 units) of stack slot 0 of this thread
   double *aux; // local memory

   __device__ __forceinline__ double ld_q(int r) const { return __ldg((const double *)(qb + (unsigned long long)(unsigned)r * ld8)); }
   __device__ __forceinline__ double ld_qd(int r) const { return __ldg((const double *)(qdb + (unsigned long long)(unsigned)r * ld8)); }
   __device__ __forceinline__ double ld_x(int r) const { return __ldg((const double *)(xb + (unsigned long long)(unsigned)r * ld8)); }
   __device__ __forceinline__ double ld_fext(int b, int k) const { return __ldg((const double *)(fb + (unsigned long long)(unsigned)(6 * b + k) * ld8)); }
   __device__ __forceinline__ void st_out(int r, double v) { *(double *)(ob + (unsigned long long)(unsigned)r * ld8) = v; }
   __device__ __forceinline__ void stk_ld2(int slot2, int j, double &a, double &b) const
   {
      const double2 t = reinterpret_cast<const double2 *>(mb_smem)[stk0 + (slot2 + j) * BLOCK];
      a = t.x;
      b = t.y;
   }
   __device__ __forceinline__ void stk_st2(int slot2, int j, double a, double b) { reinterpret_cast<double2 *>(mb_smem)[stk0 + (slot2 + j) * BLOCK] = make_double2(a, b); }
   __device__ __forceinline__ double aux_ld(int i) const { return aux[i]; }
   __device__ __forceinline__ void aux_st(int i, double v) { aux[i] = v; }
   // ABA pass-two records: global workspace of double2, one column per resident thread (coalesced 16-byte accesses),
   // read back in pass three through the ring below with cp.async.cg (L2, the coherence point of the earlier stores)
   double2 *wsb;       // workspace + column of this thread
   long long ws_ld;
   int ring3_0;        // index (double2) of this thread's element of stage 0, row 0 of the pass-three ring
   __device__ __forceinline__ void rec_st2(int i2, double a, double b) { wsb[i2 * ws_ld] = make_double2(a, b); }
   // pass-three ring: [stage][(q, qd) | rec0 .. rec3][BLOCK] double2, overlaid on the (then idle) stack area
   __device__ __forceinline__ void pf3_issue(int stage, int cfg, int dof, int rec2, int mask) const
   {
      const unsigned dst = (unsigned)__cvta_generic_to_shared(reinterpret_cast<double2 *>(mb_smem) + ring3_0 + stage * 5 * BLOCK);
      if (mask & 1)
         asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(qb + (unsigned long long)(unsigned)cfg * ld8) : "memory");
      if (mask & 2)
         asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst + 8), "l"(qdb + (unsigned long long)(unsigned)dof * ld8) : "memory");
      const double2 *src = wsb + rec2 * ws_ld;
#pragma unroll
      for (int j = 0; j < 4; j++)
         asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + (1 + j) * BLOCK * 16), "l"(src + j * ws_ld) : "memory");
   }
   __device__ __forceinline__ void pf3_ld2(int stage, int row, double &a, double &b) const
   {
      const double2 t = reinterpret_cast<const double2 *>(mb_smem)[ring3_0 + (stage * 5 + row) * BLOCK];
      a = t.x;
      b = t.y;
   }
   __device__ __forceinline__ void pass_fence() const { __threadfence(); }
   __device__ __forceinline__ const double *cst(int b) const { return mb_smem + b * MB_CONST_STRIDE; }
   // mass matrix: entry e = row * nv + col lives at mbase + e * mstride (entry-major: mstride = ld8; state-major: 8)
   char *mbase;
   unsigned mstride;
   int nv;
   const uint4 *zlist;
   int nz8;
   __device__ __forceinline__ int n_dofs() const { return nv; }
   __device__ __forceinline__ void st_M(int e, double v) const { __stcs((double *)(mbase + (unsigned long long)(unsigned)e * mstride), v); }
   __device__ __forceinline__ void zero_fill() const
   {
#pragma unroll 1
      for (int k = 0; k < nz8; k++)
      {
         const uint4 u = __ldg(zlist + k);
         st_M(u.x & 0xffffu, 0.0); st_M(u.x >> 16, 0.0); st_M(u.y & 0xffffu, 0.0); st_M(u.y >> 16, 0.0);
         st_M(u.z & 0xffffu, 0.0); st_M(u.z >> 16, 0.0); st_M(u.w & 0xffffu, 0.0); st_M(u.w >> 16, 0.0);
      }
   }
   // prefetch ring: [stage][q | qd | x][BLOCK] doubles in shared memory, filled by cp.async (LDGSTS)
   int ring0; // index (doubles) of this thread's element of stage 0, row 0
   // mask: 1 = q[cfg], 2 = qd[dof], 4 = x[dof]
   __device__ __forceinline__ void pf_issue(int stage, int cfg, int dof, int mask) const
   {
      const unsigned dst = (unsigned)__cvta_generic_to_shared(mb_smem + ring0 + stage * 3 * BLOCK);
      if (mask & 1)
         asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(qb + (unsigned long long)(unsigned)cfg * ld8) : "memory");
      if (mask & 2)
         asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst + BLOCK * 8), "l"(qdb + (unsigned long long)(unsigned)dof * ld8) : "memory");
      if (mask & 4)
         asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst + 2 * BLOCK * 8), "l"(xb + (unsigned long long)(unsigned)dof * ld8) : "memory");
   }
   __device__ __forceinline__ void pf_commit() const { asm volatile("cp.async.commit_group;" ::: "memory"); }
   template <int N> __device__ __forceinline__ void pf_wait() const { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
   __device__ __forceinline__ double pf_ld(int stage, int j) const { return mb_smem[ring0 + (stage * 3 + j) * BLOCK]; }
};

template <int ALGO, bool FEXT, bool STATE_MAJOR, int BLOCK, int AUXN, int RECN>
__global__ void __launch_bounds__(BLOCK) thread_kernel(const __grid_constant__ MbProgram P, const KernelArgs a)
{
   const int ncst = P.nb * MB_CONST_STRIDE;
   for (int i = threadIdx.x; i < ncst; i += BLOCK)
      mb_smem[i] = a.consts[i];
   __syncthreads();
   double aux[AUXN > 0 ? AUXN : 1];
   GpuCtx2<BLOCK> c2;
   c2.ld8 = (unsigned)(a.ld * 8);
   c2.stk0 = (((ncst + 1) & ~1) >> 1) + threadIdx.x;
   c2.ring3_0 = c2.stk0;
   c2.ring0 = ((ncst + 1) & ~1) + 2 * P.stack2 * BLOCK + threadIdx.x;
   c2.aux = aux;
   c2.nv = a.nv;
   c2.mstride = STATE_MAJOR ? 8u : c2.ld8;
   c2.zlist = (const uint4 *)a.zero_entries;
   c2.nz8 = a.n_zero >> 3;
   c2.wsb = reinterpret_cast<double2 *>(a.ws) + ((long long)blockIdx.x * BLOCK + threadIdx.x);
   c2.ws_ld = a.ws_ld;
   // persistent grid: each block walks over tiles of BLOCK states.  The per-state areas (stack, rings) are private to a
   // thread and the constant records are read-only, so the threads of a block never synchronise again.
   const long long ntiles = (a.n + BLOCK - 1) / BLOCK;
   for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x)
   {
      const long long s = tile * BLOCK + threadIdx.x;
      if (s >= a.n)
         break;
      c2.qb = (const char *)(a.q + s); c2.qdb = (const char *)(a.qd + s); c2.xb = (const char *)(a.x + s);
      c2.fb = (const char *)(a.fext + s); c2.ob = (char *)(a.out + s);
      c2.mbase = STATE_MAJOR ? (char *)(a.out + s * (long long)a.nv * a.nv) : (char *)(a.out + s);
      if constexpr (ALGO == MB_RNEA)
         rnea_state<double, GpuCtx2<BLOCK>, FEXT>(P, c2, a.grav, !(a.flags & 1u), !(a.flags & 2u));
      else if constexpr (ALGO == MB_ABA)
         aba_state<double, GpuCtx2<BLOCK>, FEXT>(P, c2, a.grav);
      else
         crba_state<double, GpuCtx2<BLOCK>>(P, c2);
   }
}

// compiled work-area classes (local memory per thread): {aux, rec}
//   class 0: up to 4 nested branching bodies, 32 one-DoF-equivalent records (humanoids); blocks of 256 / 128
//   class 1: up to 16 nested branching bodies, 128 bodies; blocks of 128 / 64 / 32 (deeper stacks)
constexpr int kRnaAux0 = 12 * 4, kRnaAux1 = 12 * 16;
constexpr int kAbaAux0 = 27 * 4, kAbaAux1 = 27 * 16;
constexpr int kCrbAux0 = 10 * 4, kCrbAux1 = 10 * 16;
constexpr int kAbaRec0 = MB_ABA_REC * 33, kAbaRec1 = MB_ABA_REC * 128;
constexpr int kNumCfg = 8;
constexpr int kCfgClass[kNumCfg] = {0, 0, 0, 0, 0, 1, 1, 1};
constexpr int kCfgBlock[kNumCfg] = {384, 320, 256, 192, 128, 128, 64, 32};

typedef void (*KernelFn)(const MbProgram, const KernelArgs);

template <int ALGO, bool FEXT, bool SM> KernelFn pick_cfg(int cfg)
{
   constexpr int a0 = ALGO == MB_RNEA ? kRnaAux0 : (ALGO == MB_ABA ? kAbaAux0 : kCrbAux0);
   constexpr int a1 = ALGO == MB_RNEA ? kRnaAux1 : (ALGO == MB_ABA ? kAbaAux1 : kCrbAux1);
   constexpr int r0 = ALGO == MB_ABA ? kAbaRec0 : 0, r1 = ALGO == MB_ABA ? kAbaRec1 : 0;
   switch (cfg)
   {
      case 0: return thread_kernel<ALGO, FEXT, SM, 384, a0, r0>;
      case 1: return thread_kernel<ALGO, FEXT, SM, 320, a0, r0>;
      case 2: return thread_kernel<ALGO, FEXT, SM, 256, a0, r0>;
      case 3: return thread_kernel<ALGO, FEXT, SM, 192, a0, r0>;
      case 4: return thread_kernel<ALGO, FEXT, SM, 128, a0, r0>;
      case 5: return thread_kernel<ALGO, FEXT, SM, 128, a1, r1>;
      case 6: return thread_kernel<ALGO, FEXT, SM, 64, a1, r1>;
      default: return thread_kernel<ALGO, FEXT, SM, 32, a1, r1>;
   }
}

KernelFn pick(int algo, bool fext, bool state_major, int cfg)
{
   if (algo == MB_RNEA) return fext ? pick_cfg<MB_RNEA, true, false>(cfg) : pick_cfg<MB_RNEA, false, false>(cfg);
   if (algo == MB_ABA) return fext ? pick_cfg<MB_ABA, true, false>(cfg) : pick_cfg<MB_ABA, false, false>(cfg);
   return state_major ? pick_cfg<MB_CRBA, false, true>(cfg) : pick_cfg<MB_CRBA, false, false>(cfg);
}

int class_of(int algo, const MbProgram &P)
{
   const int aux0 = algo == MB_RNEA ? kRnaAux0 : (algo == MB_ABA ? kAbaAux0 : kCrbAux0);
   const int aux1 = algo == MB_RNEA ? kRnaAux1 : (algo == MB_ABA ? kAbaAux1 : kCrbAux1);
   const int rec0 = algo == MB_ABA ? kAbaRec0 : 0, rec1 = algo == MB_ABA ? kAbaRec1 : 0;
   if (P.aux_doubles <= aux0 && P.rec_doubles <= rec0) return 0;
   if (P.aux_doubles <= aux1 && P.rec_doubles <= rec1) return 1;
   return -1;
}

size_t smem_bytes(int algo, const MbProgram &P, int block)
{
   const int ncst = (P.nb * MB_CONST_STRIDE + 1) & ~1;
   (void)algo;
   return sizeof(double) * ((size_t)ncst + (2 * (size_t)P.stack2 + 3 * MB_PF_STAGES) * block);
}
} // namespace

cudaError_t plan_thread_kernel(int algo, const MbProgram &P, bool fext, LaunchPlan &plan, bool *fits)
{
   *fits = false;
   const int cls = class_of(algo, P);
   if (cls < 0)
      return cudaSuccess;
   int dev = 0, max_optin = 0;
   cudaError_t e = cudaGetDevice(&dev);
   if (e != cudaSuccess) return e;
   e = cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
   if (e != cudaSuccess) return e;
   int best_threads = 0;
   for (int cfg = 0; cfg < kNumCfg; cfg++)
   {
      if (kCfgClass[cfg] < cls)
         continue; // work areas too small
      const int b = kCfgBlock[cfg];
      const size_t sm = smem_bytes(algo, P, b);
      if (sm > (size_t)max_optin)
         continue;
      // every variant of this configuration gets the opt-in so that later launches cannot fail on it
      for (int f = 0; f < 2; f++)
         for (int st = 0; st < 2; st++)
         {
            e = cudaFuncSetAttribute((const void *)pick(algo, f != 0, st != 0, cfg), cudaFuncAttributeMaxDynamicSharedMemorySize, max_optin);
            if (e != cudaSuccess) return e;
         }
      KernelFn fn = pick(algo, fext, false, cfg);
      int nblk = 0;
      e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nblk, (const void *)fn, b, sm);
      if (e != cudaSuccess) return e;
      // prefer more resident states; on ties the smaller work-area class (first in the table) wins
      if (nblk * b > best_threads)
      {
         best_threads = nblk * b;
         plan.block = b;
         plan.smem = sm;
         plan.blocks_per_sm = nblk;
         plan.size_class = cfg;
      }
   }
   if (best_threads == 0)
      return cudaSuccess; // the stack of even a 32-state block does not fit in shared memory
   cudaFuncAttributes attr;
   e = cudaFuncGetAttributes(&attr, (const void *)pick(algo, fext, false, plan.size_class));
   if (e != cudaSuccess) return e;
   int sms = 0;
   cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
   plan.grid = sms * plan.blocks_per_sm;
   plan.ws_doubles = algo == MB_ABA ? (size_t)plan.grid * plan.block * (size_t)std::max(P.rec_doubles, 1) : 0;
   plan.regs = attr.numRegs;
   plan.local_bytes = (int)attr.localSizeBytes;
   plan.static_smem = (int)attr.sharedSizeBytes;
   *fits = true;
   return cudaSuccess;
}

cudaError_t launch_thread_kernel(int algo, const MbProgram &P, const KernelArgs &a, const LaunchPlan &plan, cudaStream_t stream)
{
   if (a.n <= 0)
      return cudaSuccess;
   const bool state_major = algo == MB_CRBA && (a.flags & 1u);
   KernelFn fn = pick(algo, a.fext != nullptr, state_major, plan.size_class);
   const long long ntiles = (a.n + plan.block - 1) / plan.block;
   // ABA runs as a persistent grid (its pass-two records live in a workspace with one column per resident thread);
   // RNEA and CRBA measured faster with one block per tile (hardware block scheduling keeps the SMs evenly loaded)
   const unsigned grid = algo == MB_ABA ? (unsigned)std::min<long long>(ntiles, plan.grid) : (unsigned)ntiles;
   KernelArgs b = a;
   b.ws_ld = (long long)plan.grid * plan.block;
   fn<<<grid, plan.block, plan.smem, stream>>>(P, b);
   return cudaGetLastError();
}

// ---------------------------------------------------------------------------- roofline denominators
namespace
{
__global__ void __launch_bounds__(256) dfma_chain_kernel(double *out, int iters, double seed)
{
   // 8 independent FMA chains per thread: enough ILP to saturate the FP64 pipe at modest occupancy
   double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
   const double m = 1.0000001, b = 1e-9;
   for (int i = 0; i < iters; i++)
   {
      a0 = fma(a0, m, b); a1 = fma(a1, m, b); a2 = fma(a2, m, b); a3 = fma(a3, m, b);
      a4 = fma(a4, m, b); a5 = fma(a5, m, b); a6 = fma(a6, m, b); a7 = fma(a7, m, b);
   }
   out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

__global__ void __launch_bounds__(256) copy_kernel(const double2 *__restrict__ src, double2 *__restrict__ dst, long long n2)
{
   const long long stride = (long long)gridDim.x * blockDim.x;
   for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += stride)
      dst[i] = src[i];
}
} // namespace

cudaError_t measure_fp64_peak(double *tflops)
{
   int dev = 0, sms = 0;
   cudaError_t e = cudaGetDevice(&dev);
   if (e != cudaSuccess) return e;
   cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
   const int blocks = sms * 8, threads = 256, iters = 1 << 15;
   double *out = nullptr;
   e = cudaMalloc(&out, sizeof(double) * blocks * threads);
   if (e != cudaSuccess) return e;
   cudaEvent_t t0, t1;
   cudaEventCreate(&t0);
   cudaEventCreate(&t1);
   double best = 0;
   for (int rep = 0; rep < 6; rep++)
   {
      cudaEventRecord(t0);
      dfma_chain_kernel<<<blocks, threads>>>(out, iters, 1.0 + rep);
      cudaEventRecord(t1);
      e = cudaEventSynchronize(t1);
      if (e != cudaSuccess) break;
      float ms = 0;
      cudaEventElapsedTime(&ms, t0, t1);
      const double flops = 2.0 * 8.0 * (double)iters * blocks * threads;
      if (rep > 0) best = std::max(best, flops / (ms * 1e-3) / 1e12);
   }
   cudaEventDestroy(t0);
   cudaEventDestroy(t1);
   cudaFree(out);
   *tflops = best;
   return e;
}

cudaError_t measure_hbm_peak(double *gbs)
{
   const long long bytes = 1ll << 31; // 2 GiB each way, far beyond L2
   double2 *src = nullptr, *dst = nullptr;
   cudaError_t e = cudaMalloc(&src, bytes);
   if (e != cudaSuccess) return e;
   e = cudaMalloc(&dst, bytes);
   if (e != cudaSuccess) { cudaFree(src); return e; }
   cudaMemset(src, 1, bytes);
   int dev = 0, sms = 0;
   cudaGetDevice(&dev);
   cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
   cudaEvent_t t0, t1;
   cudaEventCreate(&t0);
   cudaEventCreate(&t1);
   double best = 0;
   for (int rep = 0; rep < 6; rep++)
   {
      cudaEventRecord(t0);
      copy_kernel<<<sms * 16, 256>>>(src, dst, bytes / (long long)sizeof(double2));
      cudaEventRecord(t1);
      e = cudaEventSynchronize(t1);
      if (e != cudaSuccess) break;
      float ms = 0;
      cudaEventElapsedTime(&ms, t0, t1);
      if (rep > 0) best = std::max(best, 2.0 * bytes / (ms * 1e-3) / 1e9);
   }
   cudaEventDestroy(t0);
   cudaEventDestroy(t1);
   cudaFree(src);
   cudaFree(dst);
   *gbs = best;
   return e;
}
} // namespace mb
