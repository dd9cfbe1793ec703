// entry_major_store.cu -- the store pattern of the CRBA kernel without its arithmetic: thread = state, every thread writes E
// rows of an entry-major [E][n] matrix (a warp store = one 256-byte segment, consecutive stores of a warp land n * 8 bytes
// apart).  Compares the streaming ceiling of that pattern with a linear fill.  nvcc -arch=sm_100a -O3 -o ems entry_major_store.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE> __device__ __forceinline__ void st(double *p, double v)
{
   if (MODE == 0) asm volatile("st.global.cs.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
   else if (MODE == 1) asm volatile("st.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
   else if (MODE == 2) asm volatile("st.global.cg.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
   else asm volatile("st.global.wt.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}

// the same scattered pattern with each cache operator, and with two states per thread (16-byte stores, 512-byte warp segments)
template <int MODE> __global__ void __launch_bounds__(256) entry_major_op(double *out, long long n, int E)
{
   const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
   if (s >= n) return;
   double v = (double)s;
   for (int k = 0; k < E; k++)
   {
      const int e = (int)(((long long)k * 997) % E);
      st<MODE>(out + (long long)e * n + s, v);
      v += 1.0;
   }
}
__global__ void __launch_bounds__(256) entry_major_x2(double *out, long long n, int E)
{
   const long long s = 2 * ((long long)blockIdx.x * blockDim.x + threadIdx.x);
   if (s >= n) return;
   double v = (double)s;
   for (int k = 0; k < E; k++)
   {
      const int e = (int)(((long long)k * 997) % E);
      asm volatile("st.global.cs.v2.f64 [%0], {%1, %2};" ::"l"(out + (long long)e * n + s), "d"(v), "d"(v + 0.5) : "memory");
      v += 1.0;
   }
}

__global__ void __launch_bounds__(256) entry_major(double *out, long long n, int E, int order)
{
   const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
   if (s >= n) return;
   double v = (double)s;
   for (int k = 0; k < E; k++)
   {
      // order 0: rows in sequence; order 1: a scattered sequence like the depth-first entry order of the mass matrix
      const int e = order ? (int)(((long long)k * 997) % E) : k;
      asm volatile("st.global.cs.f64 [%0], %1;" ::"l"(out + (long long)e * n + s), "d"(v) : "memory");
      v += 1.0;
   }
}

__global__ void __launch_bounds__(256) linear_fill(double2 *out, long long n2)
{
   const long long stride = (long long)gridDim.x * blockDim.x;
   for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += stride)
      out[i] = make_double2(1.0, 2.0);
}

int main()
{
   const long long n = 1 << 20;
   const int E = 1369;
   double *out;
   cudaMalloc(&out, sizeof(double) * n * E);
   cudaEvent_t t0, t1;
   cudaEventCreate(&t0);
   cudaEventCreate(&t1);
   for (int mode = 10; mode < 15; mode++)
   {
      float best = 1e30f;
      for (int rep = 0; rep < 5; rep++)
      {
         cudaEventRecord(t0);
         if (mode == 10) entry_major_op<0><<<(unsigned)(n / 256), 256>>>(out, n, E);
         else if (mode == 11) entry_major_op<1><<<(unsigned)(n / 256), 256>>>(out, n, E);
         else if (mode == 12) entry_major_op<2><<<(unsigned)(n / 256), 256>>>(out, n, E);
         else if (mode == 13) entry_major_op<3><<<(unsigned)(n / 256), 256>>>(out, n, E);
         else entry_major_x2<<<(unsigned)(n / 512), 256>>>(out, n, E);
         cudaEventRecord(t1);
         cudaEventSynchronize(t1);
         float ms;
         cudaEventElapsedTime(&ms, t0, t1);
         if (rep > 0 && ms < best) best = ms;
      }
      const char *names[5] = {"scattered rows, st.cs", "scattered rows, st (write-back)", "scattered rows, st.cg", "scattered rows, st.wt", "scattered rows, two states per thread (st.cs.v2)"};
      printf("{\"pattern\": \"%s\", \"ms\": %.4f, \"gbs\": %.1f}\n", names[mode - 10], best, 8.0 * n * E / (best * 1e-3) / 1e9);
   }
   for (int mode = 0; mode < 3; mode++)
   {
      float best = 1e30f;
      for (int rep = 0; rep < 5; rep++)
      {
         cudaEventRecord(t0);
         if (mode < 2)
            entry_major<<<(unsigned)(n / 256), 256>>>(out, n, E, mode);
         else
            linear_fill<<<148 * 16, 256>>>((double2 *)out, n * E / 2);
         cudaEventRecord(t1);
         cudaEventSynchronize(t1);
         float ms;
         cudaEventElapsedTime(&ms, t0, t1);
         if (rep > 0 && ms < best) best = ms;
      }
      printf("{\"pattern\": \"%s\", \"ms\": %.4f, \"gbs\": %.1f}\n", mode == 0 ? "entry-major, rows in sequence" : (mode == 1 ? "entry-major, rows scattered" : "linear fill"),
             best, 8.0 * n * E / (best * 1e-3) / 1e9);
   }
   return 0;
}
