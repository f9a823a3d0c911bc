// assign.cu -- stage 2b + stage 3 accumulation.
//
// Replaces the torch eager block of PointCloudSegment.segment
// (utils/segment_utils.py:143-148,168-169: ground residual along the ray, 100 centre distances,
// argmax of -|.|, relabel) -- which materialises a (H,W,100,3) temporary of 153.6 MB per frame --
// by one fused per-pixel kernel, and accumulates what point_modeling
// (ops/cpp_modules/src/cpp_modules.cpp:471-518) and the stable label-major symbol order need.
//
// Arithmetic: every torch op is its own kernel, so each elementwise result is
// rounded to f32 before the next; size-3 reductions associate as torch_sum3 (common.cuh).
//   channel 0   : | range - ( -g3 / sum3(g*lut) ) |
//   channel c>=1: sqrt( sum3( (pc - centre_c)^2 ) )          pc = range * lut
//   label = first argmin; label>0 -> +1; range==0 -> 1
// The sqrt is only taken when a squared distance undercuts every earlier one (sqrt is monotone,
// so a candidate that is not a running minimum of the squares cannot be a strict minimum of the
// roots); this keeps the first-index tie rule of torch.max bit-exact.
//
// Per-label statistics are exact: range * 2^28 is an integer for range in [2^-5, 256), summed in
// u64 -- identical to the reference's double accumulation in raster order, whose partial sums are
// then all exactly representable.  Ranges outside that interval raise flags bit 0 and
// the frame's means are recomputed sequentially by model.cu.
#include "book.cuh"

namespace rpcc {

constexpr int kTile = RPCC_TILE;  // 1024 threads, one pixel each

// Centres are kept sorted by their distance to the sensor (sort_centers_kernel, once per frame); a
// pixel at range r only has to look at the centres whose norm lies within its current best distance
// of r (| |p| - |c| | <= |p - c|), walking outwards from r and stopping as soon as the nearer side of
// the window is out of reach.  The skip test carries a slack that covers every float rounding involved
// (|p| vs r, the computed norms, the reference's own distance arithmetic), so a skipped centre can
// neither beat nor tie the current best; the evaluated ones use the reference arithmetic verbatim,
// with torch.max's first-index rule.
__global__ void __launch_bounds__(128)
sort_centers_kernel(const float* __restrict__ centers, int m, float4* __restrict__ sorted_xyzi, float* __restrict__ sorted_norm) {
  extern __shared__ float s_n[];
  const int f = blockIdx.x;
  for (int c = threadIdx.x; c < m; c += blockDim.x) {
    const float* cp = centers + ((size_t)f * m + c) * 3;
    s_n[c] = sqrtf(cp[0] * cp[0] + cp[1] * cp[1] + cp[2] * cp[2]);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < m; c += blockDim.x) {
    const float cn = s_n[c];
    int rank = 0;  // position in (norm, index) order
    for (int q = 0; q < m; ++q) {
      const float o = s_n[q];
      rank += (o < cn || (o == cn && q < c)) ? 1 : 0;
    }
    const float* cp = centers + ((size_t)f * m + c) * 3;
    sorted_xyzi[(size_t)f * m + rank] = make_float4(cp[0], cp[1], cp[2], __int_as_float(c + 1));
    sorted_norm[(size_t)f * m + rank] = cn;
  }
}

__global__ void __launch_bounds__(kTile, 2)
assign_labels_kernel(const float* __restrict__ range, const float* __restrict__ lut, const float* __restrict__ ground,
                     const float4* __restrict__ sorted_xyzi, const float* __restrict__ sorted_norm, int HW, int W, int m,
                     int T, uint8_t* __restrict__ labels, Book bk) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int K = m + 2;
  float4* s_c = reinterpret_cast<float4*>(smem_raw);                               // [m] x, y, z, bits(centre index + 1)
  unsigned long long* s_sum = reinterpret_cast<unsigned long long*>(s_c + m);     // [K]
  unsigned* s_cnt = reinterpret_cast<unsigned*>(s_sum + K);                       // [K]
  unsigned* s_flag = s_cnt + K;
  unsigned* s_ccnt = s_flag + 1;
  unsigned* s_last = s_ccnt + 1;                                                  // [32]
  float* s_norm = reinterpret_cast<float*>(s_last + 32);                          // [m + 2]: -inf, norms ascending, +inf

  const int f = blockIdx.y, tile = blockIdx.x, tid = threadIdx.x;
  if (tid < m) {
    s_c[tid] = sorted_xyzi[(size_t)f * m + tid];
    s_norm[tid + 1] = sorted_norm[(size_t)f * m + tid];
  }
  if (tid == 0) { s_norm[0] = __int_as_float(0xff800000); s_norm[m + 1] = __int_as_float(0x7f800000); }
  for (int l = tid; l < K; l += kTile) { s_cnt[l] = 0; s_sum[l] = 0; }
  if (tid == 0) { *s_flag = 0; *s_ccnt = 0; }
  __syncthreads();

  const int p = tile * kTile + tid;
  const bool inb = p < HW;
  int label = 1;
  float r = 0.f;
  if (inb) {
    r = range[(size_t)f * HW + p];
    if (r != 0.0f) {
      const float t0 = lut[(size_t)p * 3], t1 = lut[(size_t)p * 3 + 1], t2 = lut[(size_t)p * 3 + 2];
      const float g0 = ground[f * 4], g1 = ground[f * 4 + 1], g2 = ground[f * 4 + 2], g3 = ground[f * 4 + 3];
      const float x = r * t0, y = r * t1, z = r * t2;
      const float rplane = (-g3) / torch_sum3(g0 * t0, g1 * t1, g2 * t2);
      float best = fabsf(r - rplane);
      int bi = 0;
      // lower bound in the padded norm array: first slot (1-based) whose norm is >= r
      int lo = 1, n = m;
      while (n > 0) {
        const int half = n >> 1;
        const bool right = s_norm[lo + half] < r;
        lo = right ? lo + half + 1 : lo;
        n = right ? n - half - 1 : half;
      }
      int hi = lo;      // next slot above r (m + 1 = exhausted, norm +inf)
      lo = lo - 1;      // next slot below r (0 = exhausted, norm -inf)
      float dlo = r - s_norm[lo], dhi = s_norm[hi] - r;
      const float slack = 2e-6f * (r + r), kscale = 0.99999f;
      while (true) {
        const bool take_lo = dlo <= dhi;
        const float dn = take_lo ? dlo : dhi;
        // | |p| - |c| | with every rounding on the safe side (|c| <= r + dn); +inf ends the walk; a NaN best
        // (degenerate ground plane) never skips
        if (!(dn < 3.0e38f) || (dn - slack - 2e-6f * dn) * kscale > best) break;
        const int slot = take_lo ? lo : hi;
        const float4 cc = s_c[slot - 1];
        if (take_lo) { --lo; dlo = r - s_norm[lo]; } else { ++hi; dhi = s_norm[hi] - r; }
        const float dx = x - cc.x, dy = y - cc.y, dz = z - cc.z;
        const float v = sqrtf(torch_sum3(dx * dx, dy * dy, dz * dz));
        const int ci = __float_as_int(cc.w);
        if (v < best || (v == best && ci < bi)) { best = v; bi = ci; }
      }
      label = bi > 0 ? bi + 1 : 0;
    }
    labels[(size_t)f * HW + p] = (uint8_t)label;
  }
  warp_label_stats(label, inb, r, s_cnt, s_sum, s_flag);
  tile_contour_count(label, inb, p, W, s_last, s_ccnt);
  __syncthreads();
  flush_tile_stats(K, f, tile, T, s_cnt, s_sum, s_flag, s_ccnt, bk);
}

// statistics only (caller-supplied labels)
__global__ void __launch_bounds__(kTile, 2)
label_stats_kernel(const float* __restrict__ range, const uint8_t* __restrict__ labels, int HW, int W, int K, int T,
                   Book bk) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  unsigned long long* s_sum = reinterpret_cast<unsigned long long*>(smem_raw);
  unsigned* s_cnt = reinterpret_cast<unsigned*>(s_sum + K);
  unsigned* s_flag = s_cnt + K;
  unsigned* s_ccnt = s_flag + 1;
  unsigned* s_last = s_ccnt + 1;
  const int f = blockIdx.y, tile = blockIdx.x, tid = threadIdx.x;
  for (int l = tid; l < K; l += kTile) { s_cnt[l] = 0; s_sum[l] = 0; }
  if (tid == 0) { *s_flag = 0; *s_ccnt = 0; }
  __syncthreads();
  const int p = tile * kTile + tid;
  const bool inb = p < HW;
  int label = 1;
  float r = 0.f;
  if (inb) {
    label = labels[(size_t)f * HW + p];
    r = range[(size_t)f * HW + p];
    if (label >= K) { label = K - 1; atomicOr(s_flag, 2u); }
  }
  warp_label_stats(label, inb, r, s_cnt, s_sum, s_flag);
  tile_contour_count(label, inb, p, W, s_last, s_ccnt);
  __syncthreads();
  flush_tile_stats(K, f, tile, T, s_cnt, s_sum, s_flag, s_ccnt, bk);
}

}  // namespace rpcc

using namespace rpcc;

static int zero_book(const Book& bk, int B, int K, cudaStream_t st) {
  RPCC_CUDA(cudaMemsetAsync(bk.label_sum, 0, book_zero_bytes(B, K), st));
  RPCC_CUDA(cudaMemsetAsync(bk.flags, 0, sizeof(unsigned) * (size_t)B, st));
  return RPCC_OK;
}

extern "C" size_t rpcc_book_bytes(int B, int H, int W, int K) {
  const int T = (H * W + kTile - 1) / kTile;
  return book_bytes(B, T, K);
}

extern "C" size_t rpcc_assign_workspace_bytes(int B, int m) { return (size_t)B * m * 20 + 64; }

extern "C" int rpcc_assign_labels_batch(const float* range, const float* lut, const float* ground, const float* centers,
                                        int B, int H, int W, int m, uint8_t* labels, void* book, void* workspace,
                                        void* stream) {
  RPCC_REQUIRE(range && lut && ground && centers && labels && book && workspace, "null pointer");
  RPCC_REQUIRE(m >= 1 && m + 2 <= RPCC_MAX_LABELS, "cluster_num must be in [1, 254]");
  RPCC_REQUIRE(B <= 65535, "at most 65535 frames per launch");
  if (B == 0) return RPCC_OK;
  const int HW = H * W, K = m + 2, T = (HW + kTile - 1) / kTile;
  cudaStream_t st = as_stream(stream);
  const Book bk = make_book(book, B, T, K);
  int rc = zero_book(bk, B, K, st);
  if (rc != RPCC_OK) return rc;
  float4* sorted_xyzi = reinterpret_cast<float4*>((reinterpret_cast<uintptr_t>(workspace) + 15) & ~(uintptr_t)15);
  float* sorted_norm = reinterpret_cast<float*>(sorted_xyzi + (size_t)B * m);
  sort_centers_kernel<<<B, 128, sizeof(float) * m, st>>>(centers, m, sorted_xyzi, sorted_norm);
  RPCC_LAUNCH_CHECK("sort_centers_kernel");
  const size_t smem = sizeof(float4) * m + (sizeof(unsigned long long) + sizeof(unsigned)) * K + sizeof(unsigned) * 34 +
                      sizeof(float) * (m + 2) + 16;
  assign_labels_kernel<<<dim3(T, B), kTile, smem, st>>>(range, lut, ground, sorted_xyzi, sorted_norm, HW, W, m, T, labels, bk);
  RPCC_LAUNCH_CHECK("assign_labels_kernel");
  return RPCC_OK;
}

extern "C" int rpcc_label_stats_batch(const float* range, const uint8_t* labels, int B, int H, int W, int K,
                                      void* book, void* stream) {
  RPCC_REQUIRE(range && labels && book, "null pointer");
  RPCC_REQUIRE(K >= 2 && K <= RPCC_MAX_LABELS, "K must be in [2, 256]");
  RPCC_REQUIRE(B <= 65535, "at most 65535 frames per launch");
  if (B == 0) return RPCC_OK;
  const int HW = H * W, T = (HW + kTile - 1) / kTile;
  cudaStream_t st = as_stream(stream);
  const Book bk = make_book(book, B, T, K);
  int rc = zero_book(bk, B, K, st);
  if (rc != RPCC_OK) return rc;
  const size_t smem = (sizeof(unsigned long long) + sizeof(unsigned)) * K + sizeof(unsigned) * 34;
  label_stats_kernel<<<dim3(T, B), kTile, smem, st>>>(range, labels, HW, W, K, T, bk);
  RPCC_LAUNCH_CHECK("label_stats_kernel");
  return RPCC_OK;
}
