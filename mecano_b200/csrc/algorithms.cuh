// algorithms.cuh -- the per-state routines of the three calculators, written once against a small "context" policy so
// that the same source is (a) inlined into the sm_100a kernels (kernels.cu) and (b) compiled for the host by the
// kernel-source emulation harness that the no-GPU tests use (tests/emu).
//   rnea.cuh : InverseDynamicsCalculator.compute()            M/algorithms/InverseDynamicsCalculator.java:496-501, 873-966
//   aba.cuh  : ForwardDynamicsCalculator.compute()            M/algorithms/ForwardDynamicsCalculator.java:508-520, 1085-1310
//   crba.cuh : CompositeRigidBodyMassMatrixCalculator         M/algorithms/CompositeRigidBodyMassMatrixCalculator.java:588-667, 700-707, 772-797
// Unlike the reference (CoM frames for RNEA, through-the-root frame changes) everything is expressed in the canonical
// joint frames of program.h with local parent<->child transforms; results are frame-independent and agree with the
// oracle to round-off.
#pragma once
#include "program.h"
#include "spatial.cuh"
#include "jointmath.cuh"
#include "rnea.cuh"
#include "aba.cuh"
#include "crba.cuh"
#include "coriolis.cuh"
