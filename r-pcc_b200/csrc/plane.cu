// plane.cu -- stage 3, model_method = 'plane': per-cluster RANSAC plane models.
//
// Replaces the Python loop of PointCloudSegment.cluster_modeling
// (reference utils/segment_utils.py:188-216): for every label >= 2
//   < 30 pixels                                   -> point model [0,0,0,mean range]
//   else open3d segment_plane(0.1, ransac_n=4, num_iterations=10) on the cluster's points and
//        plane_angle_validation (:84-93)          -> [a,b,c,d]   or, if rejected, the point model
// which costs the reference ~100 np.where scans of the image plus 100 open3d calls per frame.
//
// open3d is third-party, absent and randomised: there is nothing to be bit-exact against ("parity
// unpinned").  What is kept: hypotheses are least-squares planes of
// ransac_n distinct cluster points, scored by inlier count at the distance threshold with ties broken
// by the lower rmse; the winner is refitted on its inliers; the angle test reproduces the reference's
// expression including its precedence quirk (|n.s| / |n| * |s|) and numpy's NaN
// semantics (arccos of a value above 1 is NaN, a NaN maximum compares false => the plane is kept).
// What changes: sampling is counter-based and keyed by (seed, frame key, label) -- the frame key is 0
// unless the caller supplies one per frame, so a frame always gets the same planes wherever it sits in a
// batch or a datalist; the fallback mean is the exactly rounded one of point_model_kernel (the
// reference's plane branch uses numpy's float32 pairwise mean here, <= 1 ulp away).
//
// oracle/rpcc_oracle.c:orc_plane_models restates plane_model_kernel (samples, summation orders, 128 threads per CTA) and
// tests/test_gpu_plane.py compares the model rows byte for byte: a change of the arithmetic here must be mirrored there.
//
// Two kernels: label_order_kernel lists every non-empty pixel in the label-major stable order of the
// symbol stream (so a cluster's pixels are one contiguous slice, found through the same tile offsets
// quantize.cu uses); plane_model_kernel runs one CTA per (frame, cluster).
#include "book.cuh"
#include "ransac.cuh"

namespace rpcc {

constexpr int kOrdWarps = 8;

// order[f][pos] = flat pixel index, pos = position of the pixel's symbol in the frame's stream.
__global__ void __launch_bounds__(kOrdWarps * 32)
label_order_kernel(const uint8_t* __restrict__ labels, Book bk, int HW, int K, int T, unsigned* __restrict__ order,
                   size_t order_stride) {
  extern __shared__ unsigned s_cnt[];   // [kOrdWarps][K]
  const int f = blockIdx.y;
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int tile = blockIdx.x * kOrdWarps + (int)warp;
  if (tile >= T) return;
  unsigned* cnt = s_cnt + warp * K;
  for (int l = lane; l < K; l += 32) cnt[l] = bk.tile_off[((size_t)f * T + tile) * K + l];
  __syncwarp();
  const uint8_t* lb = labels + (size_t)f * HW;
  unsigned* out = order + (size_t)f * order_stride;
  for (int s = 0; s < RPCC_TILE / 32; ++s) {
    const int p = tile * RPCC_TILE + s * 32 + (int)lane;
    int l = p < HW ? (int)lb[p] : 1;
    if (l >= K) l = 1;
    const unsigned grp = __match_any_sync(0xffffffffu, l);
    const int leader = __ffs(grp) - 1;
    unsigned base = 0;
    if ((int)lane == leader) { base = cnt[l]; cnt[l] = base + __popc(grp); }
    base = __shfl_sync(0xffffffffu, base, leader);
    __syncwarp();
    if (l != 1) out[base + __popc(grp & lanemask_lt())] = (unsigned)p;
  }
}

constexpr int kPlThreads = 128;
constexpr int kPlWarps = kPlThreads / 32;
constexpr int kPlMaxIter = 16;
constexpr int kPlMaxSample = 16;
constexpr int kPlCap = 2048;     // cluster pixels staged in shared memory (ray + range, 16 bytes each); larger clusters re-read them

// Sums over the CTA in a fixed order (xor tree inside a warp, warps in index order), so that a frame always gets the
// same planes: the warp partials go to shared memory, the caller synchronises once and adds them up.
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// One CTA per (frame, cluster).  The cluster's pixels (a contiguous slice of `order`) are staged in shared memory once
// -- unit ray and range, so that x = f32(range * ray) is formed exactly as PCTransformer.range_image_to_point_cloud
// does -- and every later pass (hypothesis sampling, one scoring pass per hypothesis, the refit, the angle test) reads
// them from there.  Registers stay low (one hypothesis at a time), so 7 CTAs share an SM.
__global__ void __launch_bounds__(kPlThreads, 7)
plane_model_kernel(const float* __restrict__ range, const float* __restrict__ lut, const unsigned* __restrict__ order,
                   size_t order_stride, Book bk, int HW, int K, int T, int min_pixels, float dist_thr, int ransac_n, int iters,
                   double cos_thr, unsigned long long seed, const unsigned long long* __restrict__ frame_keys,
                   float* __restrict__ model) {
  extern __shared__ __align__(16) float4 s_pt[];     // [kPlCap] ray.x, ray.y, ray.z, range
  __shared__ double s_plane[kPlMaxIter][4];
  __shared__ double s_err[kPlMaxIter][kPlWarps];
  __shared__ int s_inl[kPlMaxIter][kPlWarps];
  __shared__ double s_part[10][kPlWarps];
  __shared__ double s_best[4];
  __shared__ int s_flag;
  const int f = blockIdx.y, l = blockIdx.x + 2, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (l >= K) return;
  const unsigned n = bk.label_cnt[(size_t)f * K + l];
  if ((int)n < min_pixels) return;              // utils/segment_utils.py:203-204: the point model stays
  // first symbol of the label in the frame's label-major stream = its offset in tile 0 (model.cu)
  const unsigned base = bk.tile_off[(size_t)f * T * K + l];
  const unsigned* pix = order + (size_t)f * order_stride + base;
  const float* rg = range + (size_t)f * HW;
  const bool staged = n <= (unsigned)kPlCap;
  auto fetch = [&](unsigned k) {
    const unsigned p = pix[k];
    return make_float4(__ldg(lut + (size_t)p * 3), __ldg(lut + (size_t)p * 3 + 1), __ldg(lut + (size_t)p * 3 + 2), rg[p]);
  };
  if (staged)
    for (unsigned k = tid; k < n; k += kPlThreads) s_pt[k] = fetch(k);
  if (tid == 0) s_flag = 0;
  __syncthreads();
  auto ray = [&](unsigned k) { return staged ? s_pt[k] : fetch(k); };
  auto point = [&](unsigned k, double& x, double& y, double& z) {
    // PCTransformer.range_image_to_point_cloud (dataset/transformer.py:94-98): f32 products, then f64 (open3d)
    const float4 v = ray(k);
    x = (double)(v.w * v.x); y = (double)(v.w * v.y); z = (double)(v.w * v.z);
  };

  // ---- hypotheses: thread `it` draws ransac_n distinct points and fits their least-squares plane
  if (tid < iters) {
    unsigned long long st = splitmix64(splitmix64(seed + (frame_keys ? frame_keys[f] : 0ull)) ^
                                       ((unsigned long long)l << 48) ^ ((unsigned long long)tid << 32));
    unsigned pick[kPlMaxSample];
    double s[10];
    for (int q = 0; q < 10; ++q) s[q] = 0.0;
    for (int j = 0; j < ransac_n; ++j) {
      unsigned k;
      bool dup;
      do {                                       // n >= min_pixels > ransac_n: terminates quickly
        st = splitmix64(st);
        k = (unsigned)(st % (unsigned long long)n);
        dup = false;
        for (int e = 0; e < j; ++e) dup = dup || (pick[e] == k);
      } while (dup);
      pick[j] = k;
      double x, y, z;
      point(k, x, y, z);
      s[0] += 1.0; s[1] += x; s[2] += y; s[3] += z;
      s[4] += x * x; s[5] += x * y; s[6] += x * z; s[7] += y * y; s[8] += y * z; s[9] += z * z;
    }
    double pl[4] = {0, 0, 0, 0};
    const bool ok = plane_from_sums(s, pl);
    for (int q = 0; q < 4; ++q) s_plane[tid][q] = ok ? pl[q] : 0.0;
  }
  __syncthreads();

  // ---- score the hypotheses, one pass over the staged cluster each
  for (int h = 0; h < iters; ++h) {
    const double p0 = s_plane[h][0], p1 = s_plane[h][1], p2 = s_plane[h][2], p3 = s_plane[h][3];
    int inl = 0;
    double err = 0.0;
    for (unsigned k = tid; k < n; k += kPlThreads) {
      double x, y, z;
      point(k, x, y, z);
      const double d = fabs(p0 * x + p1 * y + p2 * z + p3);
      if (d < (double)dist_thr) { ++inl; err += d * d; }
    }
    inl = __reduce_add_sync(0xffffffffu, inl);
    err = warp_sum(err);
    if (lane == 0) { s_inl[h][warp] = inl; s_err[h][warp] = err; }
  }
  __syncthreads();
  int best = -1;
  double best_cnt = 0.0, best_rmse = 0.0;
  for (int h = 0; h < iters; ++h) {              // every thread takes the same decision from the same numbers
    int ci = 0;
    double e = 0.0;
    for (int w = 0; w < kPlWarps; ++w) { ci += s_inl[h][w]; e += s_err[h][w]; }
    const double c = (double)ci;
    const bool valid = s_plane[h][0] != 0.0 || s_plane[h][1] != 0.0 || s_plane[h][2] != 0.0;
    if (!valid || c <= 0.0) continue;
    const double rmse = sqrt(e / c);
    // open3d: higher fitness, then lower inlier rmse; earlier iteration wins a full tie
    if (best < 0 || c > best_cnt || (c == best_cnt && rmse < best_rmse)) { best = h; best_cnt = c; best_rmse = rmse; }
  }
  if (best < 0) return;                          // no usable hypothesis: the point model stays (uniform decision)

  // ---- refit on the inliers of the best hypothesis
  const double b0 = s_plane[best][0], b1 = s_plane[best][1], b2 = s_plane[best][2], b3 = s_plane[best][3];
  {
    double s[10];
    for (int q = 0; q < 10; ++q) s[q] = 0.0;
    for (unsigned k = tid; k < n; k += kPlThreads) {
      double x, y, z;
      point(k, x, y, z);
      if (fabs(b0 * x + b1 * y + b2 * z + b3) < (double)dist_thr) {
        s[0] += 1.0; s[1] += x; s[2] += y; s[3] += z;
        s[4] += x * x; s[5] += x * y; s[6] += x * z; s[7] += y * y; s[8] += y * z; s[9] += z * z;
      }
    }
#pragma unroll
    for (int q = 0; q < 10; ++q) {
      const double v = warp_sum(s[q]);
      if (lane == 0) s_part[q][warp] = v;
    }
  }
  __syncthreads();
  if (tid == 0) {
    double s[10];
    for (int q = 0; q < 10; ++q) {
      double t = 0.0;
      for (int w = 0; w < kPlWarps; ++w) t += s_part[q][w];
      s[q] = t;
    }
    double pl[4];
    if (!plane_from_sums(s, pl)) { pl[0] = b0; pl[1] = b1; pl[2] = b2; pl[3] = b3; }
    for (int q = 0; q < 4; ++q) s_best[q] = pl[q];
  }
  __syncthreads();

  // ---- plane_angle_validation (utils/segment_utils.py:84-93), f64 like numpy:
  //      alpha = arccos(|n.s| / |n| * |s|); reject if max(alpha) > threshold; arccos(>1) = NaN poisons the
  //      maximum, and NaN > threshold is false, so such a plane is kept.
  const double a = s_best[0], b = s_best[1], c = s_best[2], d = s_best[3];
  const double nn = sqrt(a * a + b * b + c * c);
  int flag = 0;   // bit 0: some alpha above the threshold, bit 1: some alpha is NaN
  for (unsigned k = tid; k < n; k += kPlThreads) {
    const float4 rv = ray(k);
    const double sx = rv.x, sy = rv.y, sz = rv.z;
    const double v = fabs(a * sx + b * sy + c * sz) / nn * sqrt(sx * sx + sy * sy + sz * sz);
    if (!(v <= 1.0)) flag |= 2;               // arccos -> NaN (also v itself NaN)
    else if (v < cos_thr) flag |= 1;          // arccos(v) > threshold
  }
  flag = __reduce_or_sync(0xffffffffu, flag);
  if (lane == 0 && flag) atomicOr(&s_flag, flag);
  __syncthreads();
  const bool keep = (s_flag & 2) || !(s_flag & 1);
  if (tid == 0 && keep) {
    // the model rows are narrowed to f32 when they are packed (utils/compress_utils.py:161)
    float4 row = make_float4((float)a, (float)b, (float)c, (float)d);
    reinterpret_cast<float4*>(model)[(size_t)f * K + l] = row;
  }
}

}  // namespace rpcc

using namespace rpcc;

extern "C" int rpcc_label_order_batch(const uint8_t* labels, void* book, int B, int H, int W, int K, uint32_t* order,
                                      size_t order_stride, void* stream) {
  RPCC_REQUIRE(labels && book && order, "null pointer");
  RPCC_REQUIRE(K >= 2 && K <= 254, "K must be in [2, 254]");
  RPCC_REQUIRE(B <= 65535, "at most 65535 frames per launch");
  if (B == 0) return RPCC_OK;
  const int HW = H * W, T = (HW + RPCC_TILE - 1) / RPCC_TILE;
  const Book bk = make_book(book, B, T, K);
  label_order_kernel<<<dim3((T + kOrdWarps - 1) / kOrdWarps, B), kOrdWarps * 32, sizeof(unsigned) * kOrdWarps * K,
                       as_stream(stream)>>>(labels, bk, HW, K, T, order, order_stride);
  RPCC_LAUNCH_CHECK("label_order_kernel");
  return RPCC_OK;
}

extern "C" int rpcc_plane_model_batch(const float* range, const float* lut, const uint32_t* order, size_t order_stride,
                                      void* book, int B, int H, int W, int K, int min_pixels, float dist_thr, int ransac_n,
                                      int iterations, float angle_threshold_deg, uint64_t seed, const uint64_t* frame_keys,
                                      float* model, void* stream) {
  RPCC_REQUIRE(range && lut && order && book && model, "null pointer");
  RPCC_REQUIRE(K >= 2 && K <= 254, "K must be in [2, 254]");
  RPCC_REQUIRE(ransac_n >= 3 && ransac_n <= kPlMaxSample, "ransac_n must be in [3, 16]");
  RPCC_REQUIRE(iterations >= 1 && iterations <= kPlMaxIter, "iterations must be in [1, 16]");
  RPCC_REQUIRE(min_pixels > ransac_n, "min_pixels must exceed ransac_n");
  RPCC_REQUIRE(B <= 65535, "at most 65535 frames per launch");
  if (B == 0 || K <= 2) return RPCC_OK;
  const int HW = H * W, T = (HW + RPCC_TILE - 1) / RPCC_TILE;
  const Book bk = make_book(book, B, T, K);
  const double cos_thr = cos(3.14159265358979323846 * ((double)angle_threshold_deg / 180.0));
  plane_model_kernel<<<dim3(K - 2, B), kPlThreads, sizeof(float4) * kPlCap, as_stream(stream)>>>(
      range, lut, order, order_stride, bk, HW, K, T, min_pixels, dist_thr, ransac_n, iterations, cos_thr,
      (unsigned long long)seed, reinterpret_cast<const unsigned long long*>(frame_keys), model);
  RPCC_LAUNCH_CHECK("plane_model_kernel");
  return RPCC_OK;
}
