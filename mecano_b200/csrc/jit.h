// jit.h -- compile and load tree-specialised kernels (specialize.h) at run time: NVRTC (libnvrtc, loaded lazily with
// dlopen) produces an sm_100a cubin from the generated source plus the embedded algorithm headers; the CUDA driver
// entry points needed to load and launch it are obtained through cudaGetDriverEntryPoint, so the library links against
// neither libcuda nor libnvrtc and still loads on a machine without a GPU.  Cubins are cached on disk.
#pragma once
#include <cuda_runtime.h>

#include <string>

#include "flatten.h"
#include "kernel_args.h"
#include "specialize.h"

namespace mb
{
struct SpecKernel
{
   void *module = nullptr;   // CUmodule
   void *function = nullptr; // CUfunction
   SpecOptions opt;
   size_t smem = 0;          // dynamic shared memory per block
   int regs = 0, local_bytes = 0, static_smem = 0;
   int blocks_per_sm = 0, grid = 0; // resident blocks per SM / on the device
   double compile_seconds = 0.0;
   bool from_cache = false;
   bool ready() const { return function != nullptr; }
};

// Generates, compiles (or fetches from the cache), loads and plans one specialised kernel on the current device.
// Returns 0 or a MECANO_B200_* / CUDA error code with the text in `err`.
int spec_build(int algo, const FlatTree &tree, const SpecOptions &opt, SpecKernel &out, std::string &err);
cudaError_t spec_launch(const SpecKernel &k, const KernelArgs &a, unsigned grid, cudaStream_t stream);
void spec_unload(SpecKernel &k);
// NVRTC only (no device needed): compiles a generated source to an sm_100a cubin and reports its size
int spec_compile_only(const std::string &src, size_t *cubin_bytes, std::string &err);
// dynamic shared memory a specialised kernel needs
size_t spec_smem_bytes(int algo, const MbProgram &P, int block, int tm);
} // namespace mb
