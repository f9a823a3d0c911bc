// TEST INFRASTRUCTURE (oracle/_ref): compiles the reference's own FPS kernel
// from where it lies under /root/reference and exposes a raw-pointer launcher.
// Nothing in the product path links or loads this.
//
// The reference's sampling_gpu.h pulls in torch headers only to declare the
// at::Tensor wrapper (ops/fps/src/sampling_gpu.h:4-11); the kernel and its
// launcher (ops/fps/src/sampling_gpu.cu:24-184) are plain CUDA, so the header is
// skipped by pre-defining its include guard.
#define _SAMPLING_GPU_H
void furthest_point_sampling_kernel_launcher(int b, int n, int m, const float* dataset, float* temp, int* idxs);
#include "/root/reference/ops/fps/src/sampling_gpu.cu"

// points: (b,n,3) device f32; temp: (b,n) device f32 pre-filled with 1e10 (ops/fps/fps_utils.py:26);
// idx: (b,m) device i32.  Same launch as ops/fps/src/sampling.cpp:35.
extern "C" int ref_fps_launch(int b, int n, int m, const float* points, float* temp, int* idx) {
  furthest_point_sampling_kernel_launcher(b, n, m, points, temp, idx);
  return (int)cudaDeviceSynchronize();
}
