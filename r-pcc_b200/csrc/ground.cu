// ground.cu -- stage 2 prologue: deterministic ground-plane RANSAC on the device.
//
// Stands in for open3d's PointCloud.segment_plane(distance_threshold=0.1, ransac_n=10,
// num_iterations=100) as PointCloudSegment.segment calls it (utils/segment_utils.py:74-82,101-108).
// open3d is a third-party dependency that is neither vendored nor pinned by the reference, and the
// reference feeds it an UNSEEDED random subsample (segment_utils.py:103), so two reference runs do
// not produce the same bytes: there is nothing to be bit-exact against ("parity unpinned").
// What is kept: the candidate rule (z < -1.5; more than 5000 -> subsample to 5000; fewer than 800 ->
// every pixel), the hypothesis shape (10-point least-squares planes, 100 of them), the score
// (inlier count at 0.1 m, ties by lower rmse) and the final least-squares refit on the inliers of
// the best hypothesis.  What changes: sampling is counter-based and keyed by
// (seed, frame), the subsample is an even stride over the candidates in raster order, so the same
// frame always yields the same plane, on any GPU count.
//
// One CTA per frame; the <= 5000 candidates live in shared memory, one warp scores one hypothesis.
#include "ransac.cuh"

namespace rpcc {

#ifndef RPCC_GF_THREADS
#define RPCC_GF_THREADS 1024
#endif
#ifndef RPCC_GF_OCC
#define RPCC_GF_OCC 1
#endif
constexpr int kGfThreads = RPCC_GF_THREADS;
constexpr int kGfMaxPts = 5000;
constexpr int kGfIters = 100;
constexpr int kGfSample = 10;

__global__ void __launch_bounds__(kGfThreads, RPCC_GF_OCC)
ground_fit_kernel(const float* __restrict__ range, const float* __restrict__ lut, int HW, unsigned long long seed,
                  float z_below, float inlier_thr, float* __restrict__ ground) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* px = reinterpret_cast<float*>(smem_raw);
  float* py = px + kGfMaxPts;
  float* pz = py + kGfMaxPts;
  unsigned* s_mask = reinterpret_cast<unsigned*>(pz + kGfMaxPts);   // [NW][ceil(seg_len / 32)] candidate bits
  __shared__ int s_warp[kGfThreads / 32];
  __shared__ double s_plane[kGfIters][4];
  __shared__ unsigned long long s_score[kGfIters];  // (inliers << 32) | ~quantised rmse  (max wins)
  __shared__ double s_sums[10];
  __shared__ double s_part[kGfThreads / 32][10];
  __shared__ double s_best[4];

  const int f = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* rg = range + (size_t)f * HW;

  // Each warp owns a contiguous segment of the image, so the candidate order is the raster order
  // and no block-wide synchronisation is needed inside the two passes.
  constexpr int NW = kGfThreads / 32;
  const int seg_len = (HW + NW - 1) / NW;
  const int p_begin = warp * seg_len;
  const int p_end = min(HW, p_begin + seg_len);
  // pass 1: count candidates (z < z_below; empty pixels have z = 0) and remember them as one bit per pixel,
  // in words private to the warp's segment
  unsigned* wmask = s_mask + warp * ((seg_len + 31) / 32);
  int cnt = 0;
#pragma unroll 4
  for (int p0 = p_begin; p0 < p_end; p0 += 32) {
    const int p = p0 + lane;
    const bool c = p < p_end && (__ldg(rg + p) * __ldg(lut + 3 * p + 2) < z_below);
    const unsigned b = __ballot_sync(0xffffffffu, c);
    if (lane == 0) wmask[(p0 - p_begin) >> 5] = b;
    cnt += __popc(b);
  }
  if (lane == 0) s_warp[warp] = cnt;
  __syncthreads();
  int nc = 0, base = 0;
  for (int q = 0; q < NW; ++q) { const int c = s_warp[q]; if (q < warp) base += c; nc += c; }
  const bool use_all = nc < 800;  // segment_utils.py:105-106
  if (use_all) { nc = HW; base = p_begin; }
  const int ns = nc < kGfMaxPts ? nc : kGfMaxPts;
  const bool narrow = (unsigned long long)HW * (unsigned long long)kGfMaxPts < 0xFFFFFFFFull;

  // pass 2: keep an even stride of the candidates, in raster order; only the kept ones are loaded
  for (int p0 = p_begin; p0 < p_end; p0 += 32) {
    const int p = p0 + lane;
    const unsigned b = use_all ? __ballot_sync(0xffffffffu, p < p_end) : wmask[(p0 - p_begin) >> 5];
    const bool cand = (b >> lane) & 1u;
    if (cand) {
      // candidate number r goes to slot floor(r * ns / nc) and is kept if it is the first one there
      const int r = base + __popc(b & lanemask_lt());
      int slot = r;
      bool keep = true;
      if (ns != nc) {
        if (narrow) {   // r * ns < 2^32: one 32-bit division each instead of the 64-bit routine
          const unsigned a = (unsigned)r * (unsigned)ns;
          slot = (int)(a / (unsigned)nc);
          keep = r == 0 || slot != (int)((a - (unsigned)ns) / (unsigned)nc);
        } else {
          slot = (int)((long long)r * ns / nc);
          keep = r == 0 || slot != (int)((long long)(r - 1) * ns / nc);
        }
      }
      if (keep) {
        const float rr = __ldg(rg + p);
        px[slot] = rr * __ldg(lut + 3 * p); py[slot] = rr * __ldg(lut + 3 * p + 1); pz[slot] = rr * __ldg(lut + 3 * p + 2);
      }
    }
    base += __popc(b);
  }
  __syncthreads();

  // hypotheses: one warp each
  for (int it = warp; it < kGfIters; it += kGfThreads / 32) {
    double s[10];
    // sample j is the (j+1)-th link of one splitmix64 chain: lane j walks the chain to its own link and
    // reduces it modulo ns (the 64-bit remainder is the expensive part), lane 0 then adds the ten points in order
    int myk = 0;
    if (lane < kGfSample) {
      unsigned long long st = splitmix64(splitmix64(seed + (unsigned long long)f) ^ ((unsigned long long)it << 40));
      for (int j = 0; j <= lane; ++j) st = splitmix64(st);
      myk = (int)(st % (unsigned long long)ns);
    }
    int ks[kGfSample];
#pragma unroll
    for (int j = 0; j < kGfSample; ++j) ks[j] = __shfl_sync(0xffffffffu, myk, j);
    if (lane == 0) {
      for (int q = 0; q < 10; ++q) s[q] = 0.0;
#pragma unroll
      for (int j = 0; j < kGfSample; ++j) {
        const int k = ks[j];
        const double x = px[k], y = py[k], z = pz[k];
        s[0] += 1.0; s[1] += x; s[2] += y; s[3] += z;
        s[4] += x * x; s[5] += x * y; s[6] += x * z; s[7] += y * y; s[8] += y * z; s[9] += z * z;
      }
      double pl[4] = {0, 0, 0, 0};
      const bool ok = plane_from_sums(s, pl);
      s_plane[it][0] = ok ? pl[0] : 0.0; s_plane[it][1] = ok ? pl[1] : 0.0;
      s_plane[it][2] = ok ? pl[2] : 0.0; s_plane[it][3] = ok ? pl[3] : 0.0;
    }
    __syncwarp();
    const float a = (float)s_plane[it][0], b = (float)s_plane[it][1], c = (float)s_plane[it][2], d = (float)s_plane[it][3];
    int inl = 0;
    float err = 0.f;
    if (a != 0.f || b != 0.f || c != 0.f) {
      for (int k = lane; k < ns; k += 32) {
        const float dist = fabsf(a * px[k] + b * py[k] + c * pz[k] + d);
        if (dist < inlier_thr) { ++inl; err += dist * dist; }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      inl += __shfl_xor_sync(0xffffffffu, inl, o);
      err += __shfl_xor_sync(0xffffffffu, err, o);
    }
    if (lane == 0) {
      const float rmse = inl > 0 ? sqrtf(err / (float)inl) : 1e9f;
      // larger is better: inliers first, then smaller rmse, then smaller iteration index
      s_score[it] = ((unsigned long long)inl << 40) | ((unsigned long long)(0xFFFFFFu - min(0xFFFFFFu, (unsigned)(rmse * 1e6f))) << 8) |
                    (unsigned long long)(0xFFu - it);
    }
  }
  __syncthreads();
  if (tid == 0) {
    int bi = 0;
    for (int it = 1; it < kGfIters; ++it) if (s_score[it] > s_score[bi]) bi = it;
    for (int q = 0; q < 4; ++q) s_best[q] = s_plane[bi][q];
    for (int q = 0; q < 10; ++q) s_sums[q] = 0.0;
  }
  __syncthreads();
  // refit on the inliers of the best hypothesis (f64 sums)
  {
    const float a = (float)s_best[0], b = (float)s_best[1], c = (float)s_best[2], d = (float)s_best[3];
    double s[10];
    for (int q = 0; q < 10; ++q) s[q] = 0.0;
    for (int k = tid; k < ns; k += kGfThreads) {
      const double x = px[k], y = py[k], z = pz[k];
      if (fabsf(a * px[k] + b * py[k] + c * pz[k] + d) < inlier_thr) {
        s[0] += 1.0; s[1] += x; s[2] += y; s[3] += z;
        s[4] += x * x; s[5] += x * y; s[6] += x * z; s[7] += y * y; s[8] += y * z; s[9] += z * z;
      }
    }
    // fixed reduction order (xor tree inside the warp, warps in index order): the plane must be
    // bit-reproducible from run to run
    for (int q = 0; q < 10; ++q) {
      double v = s[q];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) s_part[warp][q] = v;
    }
  }
  __syncthreads();
  if (tid == 0) {
    for (int q = 0; q < 10; ++q) {
      double v = 0.0;
      for (int w = 0; w < kGfThreads / 32; ++w) v += s_part[w][q];
      s_sums[q] = v;
    }
    double pl[4];
    if (!plane_from_sums(s_sums, pl)) {
      if (s_best[0] != 0.0 || s_best[1] != 0.0 || s_best[2] != 0.0) { for (int q = 0; q < 4; ++q) pl[q] = s_best[q]; }
      else { pl[0] = 0.0; pl[1] = 0.0; pl[2] = 1.0; pl[3] = 1.73; }  // no usable plane: a level ground at sensor height
    }
    for (int q = 0; q < 4; ++q) ground[f * 4 + q] = (float)pl[q];
  }
}

}  // namespace rpcc

using namespace rpcc;

extern "C" int rpcc_ground_fit_batch(const float* range, const float* lut, int B, int H, int W, uint64_t seed,
                                     float* ground, void* stream) {
  RPCC_REQUIRE(range && lut && ground, "null pointer");
  if (B == 0) return RPCC_OK;
  const size_t smem = sizeof(float) * 3 * kGfMaxPts + sizeof(unsigned) * (size_t)(kGfThreads / 32) * (((H * W + kGfThreads / 32 - 1) / (kGfThreads / 32) + 31) / 32);
  RPCC_CUDA(cudaFuncSetAttribute(ground_fit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  ground_fit_kernel<<<B, kGfThreads, smem, as_stream(stream)>>>(range, lut, H * W, (unsigned long long)seed, -1.5f, 0.1f, ground);
  RPCC_LAUNCH_CHECK("ground_fit_kernel");
  return RPCC_OK;
}
