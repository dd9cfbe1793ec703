// tmem_stack.cu -- is tensor memory usable as a per-thread fp64 stack?  Measures, on one B200:
//   (1) dependent-chain latency of a double2 round trip  st -> ld -> fma   through TMEM (STTM/LDTM) and shared memory (STS/LDS.128)
//   (2) throughput of independent double2 loads with 8 / 16 warps per SM
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_stack tmem_stack.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

extern __shared__ double2 sm2[];

__device__ __forceinline__ void tm_st2(uint32_t addr, double a, double b)
{
   asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(__double2loint(a)), "r"(__double2hiint(a)),
                "r"(__double2loint(b)), "r"(__double2hiint(b))
                : "memory");
}
__device__ __forceinline__ void tm_ld2_nowait(uint32_t addr, int &r0, int &r1, int &r2, int &r3)
{
   asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr) : "memory");
}
__device__ __forceinline__ void tm_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tm_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

template <bool TMEM, int BLOCK> __global__ void __launch_bounds__(BLOCK) bench(double *out, long long *cyc, int iters, int mode)
{
   __shared__ uint32_t base_s;
   uint32_t my = 0;
   if (TMEM)
   {
      if (threadIdx.x < 32)
      {
         asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t)__cvta_generic_to_shared(&base_s)) : "memory");
         asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncthreads();
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t warp = threadIdx.x >> 5;
      my = base_s + (((warp & 3u) * 32u) << 16) + (warp >> 2) * (512u / (BLOCK / 128));
   }
   constexpr int SLOTS = 16; // double2 slots per thread exercised
   double x = 1.0 + threadIdx.x * 1e-3, y = 2.0;
   for (int j = 0; j < SLOTS; j++)
   {
      if (TMEM) tm_st2(my + 4 * j, x + j, y);
      else sm2[j * BLOCK + threadIdx.x] = make_double2(x + j, y);
   }
   if (TMEM) tm_wait_st();
   __syncthreads();
   const long long t0 = clock64();
   if (mode == 0)
   {
      // dependent chain: load, one fma, store back, load again
      for (int i = 0; i < iters; i++)
      {
         const int j = i & (SLOTS - 1);
         double a, b;
         if (TMEM)
         {
            int r0, r1, r2, r3;
            tm_ld2_nowait(my + 4 * j, r0, r1, r2, r3);
            tm_wait_ld();
            a = __hiloint2double(r1, r0); b = __hiloint2double(r3, r2);
         }
         else
         {
            const double2 t = sm2[j * BLOCK + threadIdx.x];
            a = t.x; b = t.y;
         }
         x = fma(a, 1.0000001, x);
         y = b + x;
         const int jn = (i + 1) & (SLOTS - 1);
         if (TMEM) { tm_st2(my + 4 * jn, x, y); tm_wait_st(); }
         else sm2[jn * BLOCK + threadIdx.x] = make_double2(x, y);
      }
   }
   else
   {
      // throughput: 8 independent loads per iteration, consumed by 8 independent fma chains
      double c0 = 0, c1 = 0, c2 = 0, c3 = 0, c4 = 0, c5 = 0, c6 = 0, c7 = 0;
      for (int i = 0; i < iters; i++)
      {
         double a[8], b[8];
         if (TMEM)
         {
            int r[8][4];
#pragma unroll
            for (int j = 0; j < 8; j++) tm_ld2_nowait(my + 4 * (j + (i & 8)), r[j][0], r[j][1], r[j][2], r[j][3]);
            tm_wait_ld();
#pragma unroll
            for (int j = 0; j < 8; j++) { a[j] = __hiloint2double(r[j][1], r[j][0]); b[j] = __hiloint2double(r[j][3], r[j][2]); }
         }
         else
         {
#pragma unroll
            for (int j = 0; j < 8; j++) { const double2 t = sm2[(j + (i & 8)) * BLOCK + threadIdx.x]; a[j] = t.x; b[j] = t.y; }
         }
         c0 = fma(a[0], b[0], c0); c1 = fma(a[1], b[1], c1); c2 = fma(a[2], b[2], c2); c3 = fma(a[3], b[3], c3);
         c4 = fma(a[4], b[4], c4); c5 = fma(a[5], b[5], c5); c6 = fma(a[6], b[6], c6); c7 = fma(a[7], b[7], c7);
      }
      x = c0 + c1 + c2 + c3 + c4 + c5 + c6 + c7;
   }
   const long long t1 = clock64();
   out[blockIdx.x * BLOCK + threadIdx.x] = x + y;
   if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
   __syncthreads();
   if (TMEM && threadIdx.x < 32)
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(base_s) : "memory");
}

template <bool TMEM, int BLOCK> void run(const char *name, int blocks, int mode, int iters)
{
   double *out; long long *cyc;
   cudaMalloc(&out, sizeof(double) * blocks * BLOCK);
   cudaMalloc(&cyc, sizeof(long long) * blocks);
   const size_t smem = TMEM ? 0 : sizeof(double2) * 16 * BLOCK;
   cudaFuncSetAttribute(bench<TMEM, BLOCK>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
   cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
   float best = 1e30f;
   for (int rep = 0; rep < 3; rep++)
   {
      cudaEventRecord(e0);
      bench<TMEM, BLOCK><<<blocks, BLOCK, smem>>>(out, cyc, iters, mode);
      cudaEventRecord(e1);
      cudaError_t e = cudaEventSynchronize(e1);
      if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (ms < best) best = ms;
   }
   long long c0; cudaMemcpy(&c0, cyc, sizeof c0, cudaMemcpyDeviceToHost);
   const double per_iter = (double)c0 / iters;
   const double accesses = (mode == 0 ? 2.0 : 8.0);
   printf("%-34s block %3d blocks %4d  %8.1f cyc/iter (block 0)  %6.2f cyc per double2 access per warp   %.3f ms\n", name, BLOCK, blocks, per_iter,
          per_iter / accesses, best);
   if (mode == 1)
   {
      const double bytes = (double)blocks * BLOCK * iters * 8.0 * 16.0;
      printf("%-34s    aggregate %.1f TB/s = %.1f B/clk/SM at 1.9 GHz over %d SMs\n", "", bytes / (best * 1e-3) / 1e12, bytes / (best * 1e-3) / 1.9e9 / 148, 148);
   }
   cudaFree(out); cudaFree(cyc);
}

int main()
{
   const int it = 1 << 14;
   run<false, 256>("smem latency chain, 1 block", 1, 0, it);
   run<true, 256>("tmem latency chain, 1 block", 1, 0, it);
   run<false, 256>("smem throughput 8 warps/SM", 148, 1, it);
   run<true, 256>("tmem throughput 8 warps/SM", 148, 1, it);
   run<false, 512>("smem throughput 16 warps/SM", 148, 1, it);
   run<true, 512>("tmem throughput 16 warps/SM", 148, 1, it);
   return 0;
}
