// model.cu -- stage 3: per-cluster point models and the bookkeeping of the label-major order.
//
// Replaces segment_utils_cpp.point_modeling (ops/cpp_modules/src/cpp_modules.cpp:471-518) and the
// numpy reshaping around it (utils/segment_utils.py:182-185, tools/compress.py:101-102):
//   model[f][0] = f32(ground), model[f][1] = 0, model[f][l] = [0,0,0, f32(sum_l / n_l)]  (l >= 2)
// with sum_l the exact sum of the ranges (u64 at scale 2^28 from assign.cu); an empty cluster
// yields the reference's 0.0/0 quiet NaN (x86 default NaN, 0xFFC00000 after narrowing).
// model_rows = highest label present + 1, the row count the reference writes (cpp_modules.cpp:481-486).
// Also turns the per-tile label histograms into tile_off[f][tile][l]: the position, inside the
// frame's symbol stream, of the first symbol of label l in that tile (stream order is label
// 0, 2, 3, ...; label 1 = empty pixels is skipped; raster order inside a label,
// cpp_modules.cpp:311-331), and the per-tile contour counts into tile_coff (idx_sequence positions).
#include "book.cuh"

namespace rpcc {

__global__ void __launch_bounds__(256)
point_model_kernel(const float* __restrict__ range, const uint8_t* __restrict__ labels, const float* __restrict__ ground,
                   Book bk, int HW, int W, int K, int T, float* __restrict__ model, rpcc_frame_result* __restrict__ results) {
  __shared__ unsigned s_cnt[RPCC_MAX_LABELS];
  __shared__ unsigned s_full[RPCC_MAX_LABELS];
  const int f = blockIdx.x, l = threadIdx.x;
  if (l < RPCC_MAX_LABELS) {
    const unsigned c = l < K ? bk.label_cnt[(size_t)f * K + l] : 0u;
    s_full[l] = c;
    s_cnt[l] = l == 1 ? 0u : c;
  }
  __syncthreads();

  // contour prefix over tiles: one thread, T <= a few hundred
  if (l == 255) {
    const uint8_t* lb = labels + (size_t)f * HW;
    unsigned run = 0;
    for (int t = 0; t < T; ++t) {
      bk.tile_coff[(size_t)f * T + t] = run;
      const int p = t * RPCC_TILE;
      const unsigned first = (p % W == 0 || lb[p] != lb[p - 1]) ? 1u : 0u;
      run += bk.tile_ccnt[(size_t)f * T + t] + first;
    }
    results[f].seq_count = run;
  }
  if (l == 254) {
    unsigned total = 0;
    int rows = 1;
    for (int q = 0; q < K; ++q) { total += s_cnt[q]; if (s_full[q]) rows = q + 1; }
    results[f].sym_count = total;
    results[f].model_rows = (unsigned)rows;
    results[f].flags = bk.flags[f];
  }
  if (l >= K) return;
  unsigned base = 0;
  for (int q = 0; q < l; ++q) base += s_cnt[q];

  float4 row = make_float4(0.f, 0.f, 0.f, 0.f);
  if (model == nullptr) {
    // offsets only (decode path)
  } else if (l == 0) {
    row = make_float4(ground[f * 4], ground[f * 4 + 1], ground[f * 4 + 2], ground[f * 4 + 3]);
  } else if (l >= 2) {
    const unsigned n = s_cnt[l];
    if (n == 0) {
      row.w = __int_as_float(0xFFC00000);
    } else if (!(bk.flags[f] & 1u)) {
      const double S = (double)bk.label_sum[(size_t)f * K + l] * (1.0 / 268435456.0);
      row.w = (float)(S / (double)n);
    } else {
      // exactness guard tripped (a range outside [2^-5, 256)): redo the reference's own loop,
      // double accumulation in raster order (cpp_modules.cpp:494-514).  Rare and slow by design.
      double S = 0.0;
      const float* rg = range + (size_t)f * HW;
      const uint8_t* lb = labels + (size_t)f * HW;
      for (int p = 0; p < HW; ++p) if (lb[p] == l) S += (double)rg[p];
      row.w = (float)(S / (double)n);
    }
  }
  if (model != nullptr) reinterpret_cast<float4*>(model)[(size_t)f * K + l] = row;

  unsigned run = base;
  for (int t = 0; t < T; ++t) {
    const size_t o = ((size_t)f * T + t) * K + l;
    bk.tile_off[o] = run;
    run += (l == 1) ? 0u : bk.tile_hist[o];
  }
}

}  // namespace rpcc

using namespace rpcc;

extern "C" int rpcc_point_model_batch(const float* range, const uint8_t* labels, const float* ground, void* book,
                                      int B, int H, int W, int K, float* model, rpcc_frame_result* results, void* stream) {
  RPCC_REQUIRE(labels && book && results, "null pointer");
  RPCC_REQUIRE(model == nullptr || (range && ground), "range and ground are needed to build models");
  RPCC_REQUIRE(K >= 2 && K <= 254, "K must be in [2, 254]");
  if (B == 0) return RPCC_OK;
  const int HW = H * W, T = (HW + RPCC_TILE - 1) / RPCC_TILE;
  const Book bk = make_book(book, B, T, K);
  point_model_kernel<<<B, 256, 0, as_stream(stream)>>>(range, labels, ground, bk, HW, W, K, T, model, results);
  RPCC_LAUNCH_CHECK("point_model_kernel");
  return RPCC_OK;
}
