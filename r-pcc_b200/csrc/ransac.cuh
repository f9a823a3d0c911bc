// ransac.cuh -- pieces shared by the ground fit (ground.cu) and the per-cluster plane models (plane.cu):
// a counter-based generator and open3d's least-squares plane.
#pragma once
#include "common.cuh"

namespace rpcc {

__device__ __forceinline__ unsigned long long splitmix64(unsigned long long x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

// open3d GetPlaneFromPoints: centroid + second moments, normal along the axis with the largest
// 2x2 determinant.  sums = {n, sx, sy, sz, sxx, sxy, sxz, syy, syz, szz} (f64).  Returns false if degenerate.
__device__ __forceinline__ bool plane_from_sums(const double* s, double* plane) {
  const double n = s[0];
  if (n < 3.0) return false;
  const double cx = s[1] / n, cy = s[2] / n, cz = s[3] / n;
  const double xx = s[4] - n * cx * cx, xy = s[5] - n * cx * cy, xz = s[6] - n * cx * cz;
  const double yy = s[7] - n * cy * cy, yz = s[8] - n * cy * cz, zz = s[9] - n * cz * cz;
  const double dx = yy * zz - yz * yz, dy = xx * zz - xz * xz, dz = xx * yy - xy * xy;
  const double dmax = fmax(dx, fmax(dy, dz));
  if (!(dmax > 0.0)) return false;
  double a, b, c;
  if (dmax == dx) { a = dx; b = xz * yz - xy * zz; c = xy * yz - xz * yy; }
  else if (dmax == dy) { a = xz * yz - xy * zz; b = dy; c = xy * xz - yz * xx; }
  else { a = xy * yz - xz * yy; b = xy * xz - yz * xx; c = dz; }
  const double nn = sqrt(a * a + b * b + c * c);
  if (!(nn > 0.0)) return false;
  a /= nn; b /= nn; c /= nn;
  plane[0] = a; plane[1] = b; plane[2] = c; plane[3] = -(a * cx + b * cy + c * cz);
  return true;
}

}  // namespace rpcc
