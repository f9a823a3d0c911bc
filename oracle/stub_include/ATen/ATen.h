// Minimal stand-in for <ATen/ATen.h>, used ONLY when compiling the reference's
// chamfer3D.cu into oracle/_ref (test infrastructure).  The reference kernel
// (NmDistanceKernel) is plain CUDA; only its host wrapper mentions at::Tensor.
// The wrapper is compiled against this stub and never called: ref_shims.cu
// launches the reference kernel through raw pointers with the reference's own
// launch configuration.
#pragma once
namespace at {
struct Tensor {
  long size(int) const { return 0; }
  template <typename T> T* data() const { return nullptr; }
};
}  // namespace at
