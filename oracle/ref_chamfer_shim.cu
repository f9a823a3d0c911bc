// TEST INFRASTRUCTURE (oracle/_ref): compiles the reference's own chamfer kernel
// (utils/ChamferDistancePytorch/chamfer3D/chamfer3D.cu:12-134) from where it lies
// and launches it with the reference's launch shape (chamfer3D.cu:142-143).
// Nothing in the product path links or loads this.
#include "/root/reference/utils/ChamferDistancePytorch/chamfer3D/chamfer3D.cu"

extern "C" int ref_chamfer_launch(int b, int n, const float* xyz1, int m, const float* xyz2,
                                  float* dist1, int* idx1, float* dist2, int* idx2) {
  NmDistanceKernel<<<dim3(32, 16, 1), 512>>>(b, n, xyz1, m, xyz2, dist1, idx1);
  NmDistanceKernel<<<dim3(32, 16, 1), 512>>>(b, m, xyz2, n, xyz1, dist2, idx2);
  return (int)cudaDeviceSynchronize();
}
