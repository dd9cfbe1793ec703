// Does an FP64 instruction (16 lanes per sub-partition: two passes per warp) hold the issue port for both passes, or can
// the scheduler issue to other pipes in between?  Kernel A: 8 independent DFMA chains.  Kernel B/C/D: the same plus
// 1 / 2 / 3 independent integer (IMAD) instructions per DFMA.  If the port is free during the second pass, B costs the
// same as A (one integer instruction fits in the shadow of every DFMA); if not, time grows by one issue slot per instruction.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_coissue fp64_coissue.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int NI> __global__ void __launch_bounds__(256) mix_kernel(double *out, int *iout, int iters, double seed, int iseed)
{
   double a[8];
   int b[8];
#pragma unroll
   for (int j = 0; j < 8; j++) { a[j] = seed + threadIdx.x + j; b[j] = iseed + threadIdx.x * 3 + j; }
   const double m = 1.0000001, c = 1e-9;
   const int im = iseed | 1;
   for (int i = 0; i < iters; i++)
   {
#pragma unroll
      for (int j = 0; j < 8; j++)
      {
         a[j] = fma(a[j], m, c);
#pragma unroll
         for (int t = 0; t < NI; t++)
            b[(j + t) & 7] = b[(j + t) & 7] * im + 12345; // IMAD, independent of the DFMA chains
      }
   }
   double s = 0;
   int si = 0;
#pragma unroll
   for (int j = 0; j < 8; j++) { s += a[j]; si += b[j]; }
   out[blockIdx.x * blockDim.x + threadIdx.x] = s;
   iout[blockIdx.x * blockDim.x + threadIdx.x] = si;
}

template <int NI> float run(double *out, int *iout, int blocks, int iters)
{
   cudaEvent_t t0, t1;
   cudaEventCreate(&t0);
   cudaEventCreate(&t1);
   float best = 1e30f;
   for (int rep = 0; rep < 4; rep++)
   {
      cudaEventRecord(t0);
      mix_kernel<NI><<<blocks, 256>>>(out, iout, iters, 1.0 + rep, 7 + rep);
      cudaEventRecord(t1);
      cudaEventSynchronize(t1);
      float ms;
      cudaEventElapsedTime(&ms, t0, t1);
      if (rep > 0 && ms < best) best = ms;
   }
   return best;
}

int main()
{
   int sms = 0;
   cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
   for (int wps = 1; wps <= 4; wps *= 2) // resident warps per sub-partition: 2, 4, 8 (blocks of 8 warps)
   {
      const int blocks = sms * wps, iters = 1 << 14;
      double *out;
      int *iout;
      cudaMalloc(&out, sizeof(double) * blocks * 256);
      cudaMalloc(&iout, sizeof(int) * blocks * 256);
      const float t0 = run<0>(out, iout, blocks, iters), t1 = run<1>(out, iout, blocks, iters), t2 = run<2>(out, iout, blocks, iters),
                  t3 = run<3>(out, iout, blocks, iters);
      const double dfma = 8.0 * iters * blocks * 8; // warp-level DFMA instructions
      const double cyc = 1.965e9 * 1e-3 * sms * 4; // sub-partition cycles per ms
      printf("warps/SMSP %d: cycles per DFMA per sub-partition with 0/1/2/3 IMAD per DFMA: %.2f %.2f %.2f %.2f   (ms %.3f %.3f %.3f %.3f)\n", 2 * wps,
             t0 * cyc / dfma, t1 * cyc / dfma, t2 * cyc / dfma, t3 * cyc / dfma, t0, t1, t2, t3);
      cudaFree(out);
      cudaFree(iout);
   }
   return 0;
}
