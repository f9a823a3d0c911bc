// decode.cu -- stage 5: .rpcc sections -> labels -> dequantised residual -> range -> xyz.
//
// Replaces, for a batch of frames:
//   contour_utils_cpp.recover_map          (ops/cpp_modules/src/cpp_modules.cpp:561-593)
//   QuantizationModule.dequantize_residual (utils/compress_utils.py:114-132, a Python loop of ~100
//                                           np.where scans, 21 ms/frame in the reference)
//   segment_utils_cpp.intra_predict        (cpp_modules.cpp:248-285)
//   range_rec = pred + residual, xyz = range_rec * LUT (tools/decompress.py:108-110,
//                                           dataset/transformer.py:94-101)
// Arithmetic: residual = f32( (double)q * step_f64 ), range_rec = pred + residual (f32),
// xyz = range_rec * lut (f32).
//   label(p)  = seq[ #contour bits in pixels 1..p ]                (pixel 0 always takes seq[0])
//   symbol(p) = symbols[ tile_off[tile][label] + rank of p inside (tile,label) ]   label != 1
// so decoding is the encoder's bookkeeping run backwards: labels -> histograms -> offsets -> gather.
#include "book.cuh"

namespace rpcc {

constexpr int kDTile = RPCC_TILE;

// number of set contour bits in the 32 pixels of word `w` (MSB-first bytes)
__device__ __forceinline__ unsigned load_word_be(const uint8_t* __restrict__ bits, int cbytes, int w) {
  const int b0 = w * 4;
  unsigned v = 0;
  if ((cbytes & 3) == 0 && b0 + 4 <= cbytes) {
    v = __byte_perm(*reinterpret_cast<const unsigned*>(bits + b0), 0, 0x0123);
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j) if (b0 + j < cbytes) v |= (unsigned)bits[b0 + j] << (24 - 8 * j);
  }
  return v;  // pixel i of the word is bit 31-i
}

// tile_coff[f][t] = contour bits in pixels [1, t*1024)   (pixel 0 excluded, see recover_map)
__global__ void __launch_bounds__(256)
contour_prefix_kernel(const uint8_t* __restrict__ contour_bits, int cbytes, int HW, int T, Book bk) {
  __shared__ unsigned s_cnt[1024];
  const int f = blockIdx.x;
  const uint8_t* bits = contour_bits + (size_t)f * cbytes;
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    unsigned c = 0;
    for (int j = 0; j < kDTile / 32; ++j) {
      const int w = t * (kDTile / 32) + j;
      if (w * 32 < HW) {
        unsigned v = load_word_be(bits, cbytes, w);
        if (w == 0) v &= 0x7FFFFFFFu;
        c += __popc(v);
      }
    }
    s_cnt[t] = c;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned run = 0;
    for (int t = 0; t < T; ++t) { bk.tile_coff[(size_t)f * T + t] = run; run += s_cnt[t]; }
  }
}

__global__ void __launch_bounds__(kDTile, 2)
decode_labels_kernel(const uint8_t* __restrict__ contour_bits, int cbytes, const uint16_t* __restrict__ seq,
                     size_t seq_stride, const uint32_t* __restrict__ seq_count, int HW, int W, int K, int T,
                     uint8_t* __restrict__ labels, Book bk) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  unsigned long long* s_sum = reinterpret_cast<unsigned long long*>(smem_raw);
  unsigned* s_cnt = reinterpret_cast<unsigned*>(s_sum + K);
  unsigned* s_flag = s_cnt + K;
  unsigned* s_ccnt = s_flag + 1;
  unsigned* s_wc = s_ccnt + 1;  // [32]
  unsigned* s_last = s_wc + 32; // [32]
  const int f = blockIdx.y, tile = blockIdx.x, tid = threadIdx.x;
  const unsigned lane = tid & 31, warp = tid >> 5;
  for (int l = tid; l < K; l += kDTile) { s_cnt[l] = 0; s_sum[l] = 0; }
  if (tid == 0) { *s_flag = 0; *s_ccnt = 0; }
  const int p = tile * kDTile + tid;
  const bool inb = p < HW;
  const int word = p >> 5;
  unsigned v = 0;
  if (word * 32 < HW) {
    v = load_word_be(contour_bits + (size_t)f * cbytes, cbytes, word);
    if (word == 0) v &= 0x7FFFFFFFu;
  }
  if (lane == 0) s_wc[warp] = __popc(v);
  __syncthreads();
  unsigned before = 0;
  {
    const unsigned wc = s_wc[lane];
    unsigned incl = wc;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned u = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= (unsigned)o) incl += u;
    }
    before = __shfl_sync(0xffffffffu, incl - wc, warp);
  }
  int label = 1;
  if (inb) {
    const unsigned idx = bk.tile_coff[(size_t)f * T + tile] + before + __popc(v >> (31 - lane));
    const unsigned L = seq_count ? seq_count[f] : 0xFFFFFFFFu;
    int l = idx < L ? (int)seq[(size_t)f * seq_stride + idx] : 1;
    if (idx >= L) atomicOr(s_flag, 4u);            // sequence shorter than the contour map asks for
    if (l >= K) { l = 1; atomicOr(s_flag, 2u); }   // label without a model row
    label = l;
    labels[(size_t)f * HW + p] = (uint8_t)label;
  }
  warp_label_stats(label, inb, 1.0f, s_cnt, s_sum, s_flag);
  tile_contour_count(label, inb, p, W, s_last, s_ccnt);  // run count of the decoded map (validation)
  __syncthreads();
  flush_tile_stats(K, f, tile, T, s_cnt, s_sum, s_flag, s_ccnt, bk);
}

__global__ void __launch_bounds__(kDTile, 1)
dequant_reconstruct_kernel(const uint8_t* __restrict__ labels, const int16_t* __restrict__ symbols, size_t sym_stride,
                           const uint32_t* __restrict__ sym_count, const float* __restrict__ model,
                           const double* __restrict__ steps, const float* __restrict__ lut, Book bk, int HW, int K, int T,
                           float* __restrict__ range_rec, float* __restrict__ xyz) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float4* s_model = reinterpret_cast<float4*>(smem_raw);        // [K]
  double* s_step = reinterpret_cast<double*>(s_model + K);      // [K]
  unsigned* s_tb = reinterpret_cast<unsigned*>(s_step + K);     // [K]
  uint16_t* s_wcnt = reinterpret_cast<uint16_t*>(s_tb + K);     // [32][K]
  const int f = blockIdx.y, tile = blockIdx.x, tid = threadIdx.x;
  for (int l = tid; l < K; l += kDTile) {
    s_model[l] = reinterpret_cast<const float4*>(model)[(size_t)f * K + l];
    s_step[l] = steps[(size_t)f * K + l];
    s_tb[l] = bk.tile_off[((size_t)f * T + tile) * K + l];
  }
  const int p = tile * kDTile + tid;
  const bool inb = p < HW;
  const int label = inb ? labels[(size_t)f * HW + p] : 1;
  const unsigned rank = tile_label_rank(label, K, s_wcnt);  // syncs cover the smem fills above
  if (!inb) return;
  float res = 0.0f;
  if (label != 1) {
    const unsigned pos = s_tb[label] + rank;
    const unsigned n = sym_count ? sym_count[f] : 0xFFFFFFFFu;
    const int q = pos < n ? (int)symbols[(size_t)f * sym_stride + pos] : 0;
    res = (float)((double)q * s_step[label]);
  }
  const float4 m = s_model[label];
  const float* t = lut + (size_t)p * 3;
  float pred;
  if (m.x + m.y + m.z == 0) pred = m.w;
  else pred = -m.w / (m.x * t[0] + m.y * t[1] + m.z * t[2]);
  const float rec = pred + res;
  range_rec[(size_t)f * HW + p] = rec;
  if (xyz) {
    float* o = xyz + ((size_t)f * HW + p) * 3;
    o[0] = rec * t[0]; o[1] = rec * t[1]; o[2] = rec * t[2];
  }
}

}  // namespace rpcc

using namespace rpcc;


// labels known -> residual / range / xyz.  The second half of decoding, also the batched form of
// QuantizationModule.dequantize_residual (utils/compress_utils.py:114-132).
extern "C" int rpcc_dequantize_batch(const uint8_t* labels, const int16_t* symbols, size_t sym_stride,
                                     const uint32_t* sym_count, const float* model, const double* steps, const float* lut,
                                     int B, int H, int W, int K, float* range_rec, float* xyz, void* book,
                                     rpcc_frame_result* results, int stats_ready, void* stream) {
  RPCC_REQUIRE(labels && symbols && model && steps && lut && range_rec && book && results, "null pointer");
  RPCC_REQUIRE(K >= 2 && K <= 254, "K must be in [2, 254]");
  RPCC_REQUIRE(B <= 65535, "at most 65535 frames per launch");
  if (B == 0) return RPCC_OK;
  const int HW = H * W, T = (HW + kDTile - 1) / kDTile;
  cudaStream_t st = as_stream(stream);
  const Book bk = make_book(book, B, T, K);
  int rc;
  if (!stats_ready) {
    // label histograms per tile (the range argument only feeds the mean accumulators, unused here)
    rc = rpcc_label_stats_batch(range_rec, labels, B, H, W, K, book, stream);
    if (rc != RPCC_OK) return rc;
  }
  // offsets (tile_off) + the counts a well-formed stream must have; no models are built on this path
  rc = rpcc_point_model_batch(nullptr, labels, nullptr, book, B, H, W, K, nullptr, results, stream);
  if (rc != RPCC_OK) return rc;
  const size_t smem2 = (sizeof(float4) + sizeof(double) + sizeof(unsigned)) * K + sizeof(uint16_t) * 32 * (size_t)K + 16;
  dequant_reconstruct_kernel<<<dim3(T, B), kDTile, smem2, st>>>(labels, symbols, sym_stride, sym_count, model, steps, lut,
                                                                bk, HW, K, T, range_rec, xyz);
  RPCC_LAUNCH_CHECK("dequant_reconstruct_kernel");
  return RPCC_OK;
}

extern "C" int rpcc_decode_batch(const uint8_t* contour_bits, const uint16_t* seq, size_t seq_stride,
                                 const uint32_t* seq_count, const int16_t* symbols, size_t sym_stride,
                                 const uint32_t* sym_count, const float* model, const double* steps, const float* lut,
                                 int B, int H, int W, int K, uint8_t* labels, float* range_rec, float* xyz, void* book,
                                 rpcc_frame_result* results, void* stream) {
  RPCC_REQUIRE(contour_bits && seq && symbols && model && steps && lut && labels && range_rec && book && results, "null pointer");
  RPCC_REQUIRE(K >= 2 && K <= 254, "K must be in [2, 254]");
  RPCC_REQUIRE(B <= 65535, "at most 65535 frames per launch");
  if (B == 0) return RPCC_OK;
  const int HW = H * W, T = (HW + kDTile - 1) / kDTile, cbytes = (HW + 7) / 8;
  RPCC_REQUIRE(T <= 1024, "range image too large");
  cudaStream_t st = as_stream(stream);
  const Book bk = make_book(book, B, T, K);
  RPCC_CUDA(cudaMemsetAsync(bk.label_sum, 0, book_zero_bytes(B, K), st));
  RPCC_CUDA(cudaMemsetAsync(bk.flags, 0, sizeof(unsigned) * (size_t)B, st));
  contour_prefix_kernel<<<B, 256, 0, st>>>(contour_bits, cbytes, HW, T, bk);
  RPCC_LAUNCH_CHECK("contour_prefix_kernel");
  const size_t smem1 = (sizeof(unsigned long long) + sizeof(unsigned)) * K + sizeof(unsigned) * 66;
  decode_labels_kernel<<<dim3(T, B), kDTile, smem1, st>>>(contour_bits, cbytes, seq, seq_stride, seq_count, HW, W, K, T,
                                                          labels, bk);
  RPCC_LAUNCH_CHECK("decode_labels_kernel");
  return rpcc_dequantize_batch(labels, symbols, sym_stride, sym_count, model, steps, lut, B, H, W, K, range_rec, xyz, book,
                               results, 1, stream);
}
