// features.cu -- non-uniform framework: LOAM-style key points per range-image row and the
// per-cluster salience level / quantisation step.
//
// Replaces feature_extractor_cpp.extract_features_with_segment + mark_as_picked
// (ops/cpp_modules/src/cpp_modules.cpp:28-121, 10-25) and the salience rule inside
// quantization_utils_cpp.nonuniform_quantize (:376-405).  SURVEY A.6b is the contract.
//
// Per row (one CTA of 8 warps, one warp per 1/8-row segment):
//   1. compact the pixels whose label is not ground/empty (ballot + scan), keep range and column;
//   2. curvature c[s] = ((sum_{k=-R..R} (r[s+k]-r[s]))^2 / (2R)) / r[s], f32, adds in k order;
//   3. each warp bitonic-sorts its segment's (c, s) pairs in shared memory (ascending, as
//      std::sort on pair<float,int>), then walks it from the top ("sharp": labels 3 then 2) and
//      from the bottom ("flat": label 1) with the reference's counters and break rules.  The walk
//      evaluates the gap test of mark_as_picked for 32 entries at a time and resolves the
//      sequential counters with ballot/popc.
// The reference's `cloud_neighbors_picked` never blocks a candidate (mark_as_picked only ever marks
// the candidate itself and every entry is examined at most once), so it is not materialised.
#include "book.cuh"

namespace rpcc {

struct Levels { int kp[8]; float acc[8]; int n; int ground_level; };

constexpr int kFeatThreads = 256;  // 8 warps = the reference's default `segments`
constexpr float kGap = 0.3f;       // mark_as_picked gap_threshold (cpp_modules.cpp:11)

__device__ __forceinline__ bool gap_accept(const float* __restrict__ rrow, int w, int region) {
  // cpp_modules.cpp:15-23 on the RAW row (empty pixels are 0 and therefore always reject)
  const float r = rrow[w];
  bool ok = true;
  for (int i = -region; i <= region; ++i) ok = ok && !((r - rrow[w + i]) > kGap);
  return ok;
}

// Bitonic sort of 32 * E keys by one warp, in registers: element i = lane * E + r lives in register r of `lane`, so the
// strides below E are compare-exchanges inside a lane and the others one shuffle per key; no shared-memory round trips
// and no warp barriers between the stages.  Ascending, like std::sort on pair<float,int> (the keys are distinct).
template <int E>
__device__ __forceinline__ void warp_sort_regs(unsigned long long* __restrict__ key, int lane) {
  unsigned long long v[E];
#pragma unroll
  for (int r = 0; r < E; ++r) v[r] = key[r * 32 + lane];        // any input order will do: conflict-free reads
  __syncwarp();
#pragma unroll
  for (int k = 2; k <= 32 * E; k <<= 1) {
#pragma unroll
    for (int st = k >> 1; st > 0; st >>= 1) {
      if (st < E) {
#pragma unroll
        for (int r = 0; r < E; ++r) {
          if ((r & st) == 0) {
            const bool up = ((lane * E + r) & k) == 0;
            const unsigned long long a = v[r], b = v[r | st];
            const bool sw = (a > b) == up;
            v[r] = sw ? b : a;
            v[r | st] = sw ? a : b;
          }
        }
      } else {
        const int ls = st / E;
        const bool lower = (lane & ls) == 0;
#pragma unroll
        for (int r = 0; r < E; ++r) {
          const bool up = ((lane * E + r) & k) == 0;
          const unsigned long long other = __shfl_xor_sync(0xffffffffu, v[r], ls);
          const bool keep_min = lower == up;
          const bool other_smaller = other < v[r];
          v[r] = (keep_min == other_smaller) ? other : v[r];
        }
      }
    }
  }
#pragma unroll
  for (int r = 0; r < E; ++r) key[lane * E + r] = v[r];
  __syncwarp();
}

__global__ void __launch_bounds__(kFeatThreads)
keypoints_kernel(const float* __restrict__ range, const uint8_t* __restrict__ labels, int H, int W, int K, int region,
                 int segments, int sharp_num, int less_sharp_num, int flat_num, int N,
                 uint8_t* __restrict__ key_points, float* __restrict__ feat, unsigned* __restrict__ kp_cnt) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  unsigned long long* s_key = reinterpret_cast<unsigned long long*>(smem_raw);     // [segments][N]
  float* s_row = reinterpret_cast<float*>(s_key + (size_t)segments * N);           // [W] raw range row
  float* s_vr = s_row + W;                                                         // [W] compacted range
  unsigned short* s_col = reinterpret_cast<unsigned short*>(s_vr + W);             // [W] compacted column
  __shared__ int s_warp_cnt[kFeatThreads / 32];
  __shared__ int s_total;

  const int f = blockIdx.y, h = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const size_t rowoff = ((size_t)f * H + h) * W;
  const float* rg = range + rowoff;
  const uint8_t* lb = labels + rowoff;

  // 1. load the raw row, compact non-ground / non-empty pixels in column order
  int base = 0;
  for (int w0 = 0; w0 < W; w0 += kFeatThreads) {
    const int w = w0 + tid;
    float r = 0.f;
    bool keep = false;
    if (w < W) {
      r = rg[w];
      s_row[w] = r;
      const int l = lb[w];
      keep = l != 0 && l != 1;
    }
    const unsigned b = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) s_warp_cnt[warp] = __popc(b);
    __syncthreads();
    int before = 0, total = 0;
    for (int q = 0; q < kFeatThreads / 32; ++q) {
      const int c = s_warp_cnt[q];
      if (q < warp) before += c;
      total += c;
    }
    if (keep) {
      const int pos = base + before + __popc(b & lanemask_lt());
      s_vr[pos] = r;
      s_col[pos] = (unsigned short)w;
    }
    base += total;
    __syncthreads();
  }
  const int L0 = base;
  if (L0 < segments + region * 2 + 1) return;  // cpp_modules.cpp:59

  // 2. curvature and sort keys
  const int nf = L0 - 2 * region;
  const int per = nf / segments;  // the tail nf % segments entries are never examined (:76-77)
  for (int i = tid; i < segments * N; i += kFeatThreads) s_key[i] = ~0ull;
  __syncthreads();
  for (int e = tid; e < nf; e += kFeatThreads) {
    const int s = e + region;
    const float c0 = s_vr[s];
    float a = 0.0f;
    for (int k = -region; k <= region; ++k) a += s_vr[s + k] - c0;
    a = a * a;
    a /= (float)(2 * region);
    a /= c0;
    if (feat) feat[rowoff + s_col[s]] = a;
    const int j = e / per;
    if (j < segments) s_key[(size_t)j * N + (e - j * per)] = ((unsigned long long)__float_as_uint(a) << 32) | (unsigned)s;
  }
  __syncthreads();

  // 3. one warp per segment
  for (int j = warp; j < segments; j += kFeatThreads / 32) {
    unsigned long long* key = s_key + (size_t)j * N;
    // bitonic sort, ascending; padding keys (~0) end up past `per`
    bool sort_small = false;   // (also: rows wider than 16384 pixels) the generic shared-memory network
    // only the first `per` keys of the segment are real (the rest is ~0 padding up to the allocation N): sort the
    // smallest power of two that holds them -- a row with few non-ground pixels sorts a fraction of N
    if (N < 32) sort_small = true;
    else if (per <= 32) warp_sort_regs<1>(key, lane);
    else if (per <= 64) warp_sort_regs<2>(key, lane);
    else if (per <= 128) warp_sort_regs<4>(key, lane);
    else if (per <= 256) warp_sort_regs<8>(key, lane);
    else if (per <= 512) warp_sort_regs<16>(key, lane);
    else sort_small = true;
    if (sort_small) {
      for (int k = 2; k <= N; k <<= 1) {
        for (int st = k >> 1; st > 0; st >>= 1) {
          for (int i = lane; i < N / 2; i += 32) {
            const int lo = ((i & ~(st - 1)) << 1) | (i & (st - 1));
            const int hi = lo | st;
            const bool up = (lo & k) == 0;
            const unsigned long long a = key[lo], b = key[hi];
            if ((a > b) == up) { key[lo] = b; key[hi] = a; }
          }
          __syncwarp();
        }
      }
    }
    // sharp walk: largest curvature first (cpp_modules.cpp:79-95)
    int picked = 0;
    int e1_low = 0;       // lowest sorted index examined (all of those are consumed for the flat walk)
    bool done = false;
    for (int top = per - 1; top >= 0 && !done; top -= 32) {
      const int i = top - lane;
      bool acc = false;
      int w = 0;
      if (i >= 0) {
        w = s_col[(unsigned)(key[i] & 0xFFFFFFFFull)];
        acc = gap_accept(s_row, w, region);
      }
      const unsigned am = __ballot_sync(0xffffffffu, acc);
      const int n = picked + __popc(am & (lanemask_lt() | (1u << lane)));  // running count including me
      if (acc && n < less_sharp_num) {
        const uint8_t v = n < sharp_num ? 3 : 2;
        key_points[rowoff + w] = v;
        atomicAdd(&kp_cnt[(size_t)f * K + lb[w]], 1u);
      }
      // the entry whose acceptance brings the count to less_sharp_num ends the walk (consumed, unlabelled)
      const unsigned stop = __ballot_sync(0xffffffffu, acc && n == less_sharp_num);
      if (stop) {
        done = true;
        e1_low = top - (__ffs(stop) - 1);
      }
      picked += __popc(am);
    }
    if (!done) e1_low = 0;
    // flat walk: smallest non-zero curvature first, consumed entries skipped (:97-112)
    picked = 0;
    done = false;
    for (int bot = 0; bot < e1_low && !done; bot += 32) {
      const int i = bot + lane;
      bool acc = false;
      int w = 0;
      if (i < e1_low) {
        const unsigned long long kk = key[i];
        if ((unsigned)(kk >> 32) != 0u) {  // exactly-zero curvature counts as already consumed (:100)
          w = s_col[(unsigned)(kk & 0xFFFFFFFFull)];
          acc = gap_accept(s_row, w, region);
        }
      }
      const unsigned am = __ballot_sync(0xffffffffu, acc);
      const int n = picked + __popc(am & (lanemask_lt() | (1u << lane)));
      if (acc && n < flat_num) {
        key_points[rowoff + w] = 1;
        atomicAdd(&kp_cnt[(size_t)f * K + lb[w]], 1u);
      }
      if (__ballot_sync(0xffffffffu, acc && n == flat_num)) done = true;
      picked += __popc(am);
    }
  }
}

// cpp_modules.cpp:388-405
__global__ void salience_kernel(const unsigned* __restrict__ label_cnt, const unsigned* __restrict__ kp_cnt, int B, int K,
                                Levels lv, uint8_t* __restrict__ salience, float* __restrict__ step_per_label) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * K) return;
  const int l = i % K;
  int lev = 0;
  if (l == 0) lev = lv.ground_level;
  else if (l == 1) lev = lv.n - 1;
  else if (label_cnt[i] < 30u) lev = lv.n - 1;
  else {
    for (int q = 0; q < lv.n; ++q) if ((int)kp_cnt[i] >= lv.kp[q]) { lev = q; break; }
  }
  salience[i] = (uint8_t)lev;
  step_per_label[i] = lv.acc[lev];
}

}  // namespace rpcc

using namespace rpcc;

extern "C" int rpcc_keypoints_salience_batch(const float* range, const uint8_t* labels, const void* book, int B, int H, int W,
                                             int K, int region, int segments, int sharp_num, int less_sharp_num, int flat_num,
                                             const int32_t* level_kp_num, const float* level_acc, int level_num,
                                             int ground_level, uint8_t* key_points, float* feat, uint8_t* salience,
                                             float* step_per_label, uint32_t* kp_cnt, void* stream) {
  RPCC_REQUIRE(range && labels && key_points && kp_cnt, "null pointer");
  RPCC_REQUIRE(K >= 2 && K <= 254, "K must be in [2, 254]");
  RPCC_REQUIRE(region >= 1 && region <= 16 && segments >= 1 && segments <= 64, "bad feature_region / segments");
  RPCC_REQUIRE(W <= 65535 && H <= 65535 && B <= 65535, "shape too large");
  RPCC_REQUIRE(level_num <= 8, "at most 8 salience levels");
  if (B == 0) return RPCC_OK;
  cudaStream_t st = as_stream(stream);
  const int HW = H * W;
  RPCC_CUDA(cudaMemsetAsync(key_points, 0, (size_t)B * HW, st));
  RPCC_CUDA(cudaMemsetAsync(kp_cnt, 0, sizeof(uint32_t) * (size_t)B * K, st));
  if (feat) RPCC_CUDA(cudaMemsetAsync(feat, 0, sizeof(float) * (size_t)B * HW, st));
  int per_max = (W - 2 * region) / segments;
  if (per_max < 1) per_max = 1;
  int N = 2;
  while (N < per_max) N <<= 1;
  const size_t smem = sizeof(unsigned long long) * (size_t)segments * N + sizeof(float) * 2 * (size_t)W + sizeof(unsigned short) * (size_t)W;
  RPCC_REQUIRE(smem <= 200 * 1024, "row too wide for the key-point kernel");
  RPCC_CUDA(cudaFuncSetAttribute(keypoints_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  keypoints_kernel<<<dim3(H, B), kFeatThreads, smem, st>>>(range, labels, H, W, K, region, segments, sharp_num, less_sharp_num,
                                                           flat_num, N, key_points, feat, kp_cnt);
  RPCC_LAUNCH_CHECK("keypoints_kernel");
  if (salience && step_per_label) {
    RPCC_REQUIRE(book && level_kp_num && level_acc && level_num >= 1, "salience needs the book and the level tables");
    const int T = (HW + RPCC_TILE - 1) / RPCC_TILE;
    const Book bk = make_book(const_cast<void*>(book), B, T, K);
    Levels lv;
    for (int q = 0; q < 8; ++q) { lv.kp[q] = q < level_num ? level_kp_num[q] : 0; lv.acc[q] = q < level_num ? level_acc[q] : 0.f; }
    lv.n = level_num; lv.ground_level = ground_level;
    salience_kernel<<<(B * K + 255) / 256, 256, 0, st>>>(bk.label_cnt, kp_cnt, B, K, lv, salience, step_per_label);
    RPCC_LAUNCH_CHECK("salience_kernel");
  }
  return RPCC_OK;
}
