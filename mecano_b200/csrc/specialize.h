// specialize.h -- source generator for tree-specialised kernels.
//
// The generic kernels (kernels.cu) interpret the traversal program of a tree from the constant bank: one compiled
// kernel serves every tree, at the price of decoding, flag tests, dispatch and register shuffling on every op (more
// than half of the issued instructions for a 32-body humanoid).  A calculator, however, is constructed once per
// MultiBodySystem and then evaluated millions of times -- Mecano builds its recursion-step objects in the constructor
// for the same reason (InverseDynamicsCalculator.java:253-282).  generate_source() therefore unrolls the traversal
// program of one tree into straight-line CUDA C++: a sequence of calls of the per-op routines of rnea.cuh / aba.cuh /
// crba.cuh whose op records, stack slots, save-area offsets and constant records are literals.  jit.cpp compiles that
// text with NVRTC for sm_100a when the handle is created.
#pragma once
#include <string>

#include "flatten.h"

namespace mb
{
struct SpecOptions
{
   int block = 256;       // threads per block
   int tm = 0;            // stack slots (double2) held in tensor memory
   bool fext = false;     // external wrenches (RNEA / ABA)
   bool state_major = false; // CRBA output layout
   int sync_every = 1;    // block barrier every this many ops (0 = never): keeps the warps on one instruction stream
};

// Returns the translation unit for one algorithm of one flattened tree.  The kernel is `extern "C" mb_spec_kernel`.
std::string generate_source(int algo, const FlatTree &tree, const SpecOptions &opt);
} // namespace mb
