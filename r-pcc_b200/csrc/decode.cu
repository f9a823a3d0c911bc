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

constexpr int kDWarps = 8;   // one 1024-pixel tile per warp, 8 tiles of one frame per CTA (as assign.cu / quantize.cu)

// One warp walks one tile, 32 pixels (one contour word) at a time; nothing is shared between the warps of a CTA, so
// there is no block-level synchronisation.  label(p) = seq[contour bits in pixels 1..p]; the per-(tile,label) pixel counts
// go to bins private to the warp (one leader lane per label and slice, found with match_any), the run count of the
// decoded map is kept for validation (point_model_kernel compares it with the sequence length).
__global__ void __launch_bounds__(kDWarps * 32, 6)
decode_labels_kernel(const uint8_t* __restrict__ contour_bits, int cbytes, const uint16_t* __restrict__ seq,
                     size_t seq_stride, const uint32_t* __restrict__ seq_count, const unsigned long long* __restrict__ seq_base,
                     int HW, int W, int K, int T, uint8_t* __restrict__ labels, Book bk) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  unsigned* s_cnt = reinterpret_cast<unsigned*>(smem_raw);      // [kDWarps][K]
  const int f = blockIdx.y, tid = threadIdx.x;
  const unsigned lane = tid & 31, warp = tid >> 5;
  const int tile = blockIdx.x * kDWarps + (int)warp;
  if (tile >= T) return;
  unsigned* cnt = s_cnt + warp * K;
  for (int l = lane; l < K; l += 32) cnt[l] = 0;
  __syncwarp();

  const int p_tile = tile * kDTile;
  const uint8_t* bits = contour_bits + (size_t)f * cbytes;
  // lane j holds the contour word of slice j (pixel i of a word is bit 31 - i; pixel 0 of the image never counts)
  unsigned v_mine = 0;
  {
    const int w = (p_tile >> 5) + (int)lane;
    if (w * 32 < HW) {
      v_mine = load_word_be(bits, cbytes, w);
      if (w == 0) v_mine &= 0x7FFFFFFFu;
    }
  }
  unsigned running = bk.tile_coff[(size_t)f * T + tile];
  // strided streams with optional counts, or streams packed back to back (seq_base [B+1])
  const unsigned L = seq_base ? (unsigned)(seq_base[f + 1] - seq_base[f]) : (seq_count ? seq_count[f] : 0xFFFFFFFFu);
  const uint16_t* sq = seq_base ? seq + seq_base[f] : seq + (size_t)f * seq_stride;
  uint8_t* lb = labels + (size_t)f * HW;
  const unsigned lt = lanemask_lt();
  int next_row = ((p_tile + W - 1) / W) * W;       // next pixel that starts an image row
  int carry = -1;
  unsigned ccnt = 0, flag = 0;
#pragma unroll 2
  for (int s = 0; s < kDTile / 32; ++s) {
    const int p0 = p_tile + s * 32, p = p0 + (int)lane;
    if (p0 >= HW) break;
    const bool inb = p < HW;
    const unsigned v = __shfl_sync(0xffffffffu, v_mine, s);
    const unsigned idx = running + __popc(v >> (31 - lane));
    running += __popc(v);
    int label = 1;
    if (inb) {
      int l = 1;
      if (idx < L) l = (int)__ldg(sq + idx); else flag |= 4u;     // sequence shorter than the contour map asks for
      if (l >= K) { l = 1; flag |= 2u; }                          // label without a model row
      label = l;
      lb[p] = (uint8_t)label;
    }
    // per-label pixel counts of the tile (out-of-image lanes form a group of their own that is not counted)
    {
      const unsigned grp = __match_any_sync(0xffffffffu, inb ? label : 0x7fffffff);
      if (inb && (grp & lt) == 0u) cnt[label] += (unsigned)__popc(grp);
      __syncwarp();
    }
    // run count of the decoded map inside the tile, its first pixel excluded (extract_contour, cpp_modules.cpp:534-545)
    {
      int left = __shfl_up_sync(0xffffffffu, label, 1);
      if (lane == 0) left = carry;
      carry = __shfl_sync(0xffffffffu, label, 31);
      bool rowstart = false;
      if (W >= 32) {
        if (next_row < p0 + 32) { rowstart = (p == next_row); next_row += W; }
      } else {
        rowstart = (p % W) == 0;
      }
      const bool c = inb && p != p_tile && (rowstart || label != left);
      ccnt += __popc(__ballot_sync(0xffffffffu, c));
    }
  }
  flag = __reduce_or_sync(0xffffffffu, flag);
  for (int l = lane; l < K; l += 32) {
    const unsigned c = cnt[l];
    bk.tile_hist[((size_t)f * T + tile) * K + l] = (uint16_t)c;
    if (c) atomicAdd(&bk.label_cnt[(size_t)f * K + l], c);
  }
  if (lane == 0) {
    bk.tile_ccnt[(size_t)f * T + tile] = (uint16_t)ccnt;
    if (flag) atomicOr(&bk.flags[f], flag);
  }
}

// The encoder's stable scatter run backwards: the warp's private counters start at tile_off (the position, in the frame's
// label-major symbol stream, of the tile's first symbol of each label) and advance slice by slice; a pixel's symbol sits
// at counter + (same-label lanes below it).  Then residual = f32((double)q * step), range = pred + residual, xyz = range * LUT.
__global__ void __launch_bounds__(kDWarps * 32, 6)
dequant_reconstruct_kernel(const uint8_t* __restrict__ labels, const int16_t* __restrict__ symbols, size_t sym_stride,
                           const uint32_t* __restrict__ sym_count, const unsigned long long* __restrict__ sym_base,
                           const float* __restrict__ model, const double* __restrict__ steps, const float* __restrict__ lut,
                           Book bk, int HW, int K, int T, float* __restrict__ range_rec, float* __restrict__ xyz) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float4* s_model = reinterpret_cast<float4*>(smem_raw);        // [K]
  double* s_step = reinterpret_cast<double*>(s_model + K);      // [K]
  unsigned* s_cnt = reinterpret_cast<unsigned*>(s_step + K);    // [kDWarps][K]
  const int f = blockIdx.y, tid = threadIdx.x;
  const unsigned lane = tid & 31, warp = tid >> 5;
  const int tile = blockIdx.x * kDWarps + (int)warp;
  for (int l = tid; l < K; l += kDWarps * 32) {
    s_model[l] = reinterpret_cast<const float4*>(model)[(size_t)f * K + l];
    s_step[l] = steps[(size_t)f * K + l];
  }
  unsigned* cnt = s_cnt + warp * K;
  if (tile < T)
    for (int l = lane; l < K; l += 32) cnt[l] = bk.tile_off[((size_t)f * T + tile) * K + l];
  __syncthreads();
  if (tile >= T) return;

  const int p_tile = tile * kDTile;
  const uint8_t* lb = labels + (size_t)f * HW;
  const int16_t* sym = sym_base ? symbols + sym_base[f] : symbols + (size_t)f * sym_stride;
  const unsigned n = sym_base ? (unsigned)(sym_base[f + 1] - sym_base[f]) : (sym_count ? sym_count[f] : 0xFFFFFFFFu);
  float* rr = range_rec + (size_t)f * HW;
  const unsigned lt = lanemask_lt();
#pragma unroll 2
  for (int s = 0; s < kDTile / 32; ++s) {
    const int p0 = p_tile + s * 32, p = p0 + (int)lane;
    if (p0 >= HW) break;
    const bool inb = p < HW;
    int label = inb ? (int)__ldg(lb + p) : 1;
    if (label >= K) label = 1;                       // caller-supplied label maps: flagged by label_stats, decoded as empty
    float t0 = 0.f, t1 = 0.f, t2 = 0.f;
    if (inb) { const float* t = lut + (size_t)p * 3; t0 = __ldg(t); t1 = __ldg(t + 1); t2 = __ldg(t + 2); }
    const unsigned grp = __match_any_sync(0xffffffffu, label);
    const unsigned base = cnt[label];                // same address for the whole group: one broadcast read
    __syncwarp();
    if ((grp & lt) == 0u) cnt[label] = base + (unsigned)__popc(grp);
    __syncwarp();
    float res = 0.0f;
    if (label != 1) {
      const unsigned pos = base + (unsigned)__popc(grp & lt);
      const int q = pos < n ? (int)__ldg(sym + pos) : 0;
      res = (float)((double)q * s_step[label]);      // utils/compress_utils.py:128-131
    }
    const float4 m = s_model[label];
    float pred;
    if (m.x + m.y + m.z == 0) pred = m.w;            // cpp_modules.cpp:271-279
    else pred = -m.w / (m.x * t0 + m.y * t1 + m.z * t2);
    const float rec = pred + res;
    if (inb) {
      rr[p] = rec;
      if (xyz) {
        float* o = xyz + ((size_t)f * HW + p) * 3;
        o[0] = rec * t0; o[1] = rec * t1; o[2] = rec * t2;
      }
    }
  }
}

}  // namespace rpcc

using namespace rpcc;


// labels known -> residual / range / xyz.  The second half of decoding, also the batched form of
// QuantizationModule.dequantize_residual (utils/compress_utils.py:114-132).
static int dequantize_impl(const uint8_t* labels, const int16_t* symbols, size_t sym_stride, const uint32_t* sym_count,
                           const uint64_t* sym_base, const float* model, const double* steps, const float* lut,
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
  const size_t smem2 = (sizeof(float4) + sizeof(double)) * K + sizeof(unsigned) * (size_t)kDWarps * K;
  dequant_reconstruct_kernel<<<dim3((T + kDWarps - 1) / kDWarps, B), kDWarps * 32, smem2, st>>>(
      labels, symbols, sym_stride, sym_count, reinterpret_cast<const unsigned long long*>(sym_base), model, steps, lut, bk, HW, K,
      T, range_rec, xyz);
  RPCC_LAUNCH_CHECK("dequant_reconstruct_kernel");
  return RPCC_OK;
}

extern "C" int rpcc_dequantize_batch(const uint8_t* labels, const int16_t* symbols, size_t sym_stride,
                                     const uint32_t* sym_count, const float* model, const double* steps, const float* lut,
                                     int B, int H, int W, int K, float* range_rec, float* xyz, void* book,
                                     rpcc_frame_result* results, int stats_ready, void* stream) {
  return dequantize_impl(labels, symbols, sym_stride, sym_count, nullptr, model, steps, lut, B, H, W, K, range_rec, xyz, book,
                         results, stats_ready, stream);
}

static int decode_impl(const uint8_t* contour_bits, const uint16_t* seq, size_t seq_stride, const uint32_t* seq_count,
                       const uint64_t* seq_base, const int16_t* symbols, size_t sym_stride, const uint32_t* sym_count,
                       const uint64_t* sym_base, const float* model, const double* steps, const float* lut,
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
  const size_t smem1 = sizeof(unsigned) * (size_t)kDWarps * K;
  decode_labels_kernel<<<dim3((T + kDWarps - 1) / kDWarps, B), kDWarps * 32, smem1, st>>>(
      contour_bits, cbytes, seq, seq_stride, seq_count, reinterpret_cast<const unsigned long long*>(seq_base), HW, W, K, T, labels, bk);
  RPCC_LAUNCH_CHECK("decode_labels_kernel");
  return dequantize_impl(labels, symbols, sym_stride, sym_count, sym_base, model, steps, lut, B, H, W, K, range_rec, xyz, book,
                         results, 1, stream);
}

extern "C" int rpcc_decode_batch(const uint8_t* contour_bits, const uint16_t* seq, size_t seq_stride,
                                 const uint32_t* seq_count, const int16_t* symbols, size_t sym_stride,
                                 const uint32_t* sym_count, const float* model, const double* steps, const float* lut,
                                 int B, int H, int W, int K, uint8_t* labels, float* range_rec, float* xyz, void* book,
                                 rpcc_frame_result* results, void* stream) {
  return decode_impl(contour_bits, seq, seq_stride, seq_count, nullptr, symbols, sym_stride, sym_count, nullptr, model, steps,
                     lut, B, H, W, K, labels, range_rec, xyz, book, results, stream);
}

extern "C" int rpcc_decode_packed_batch(const uint8_t* contour_bits, const uint16_t* seq, const uint64_t* seq_base,
                                        const int16_t* symbols, const uint64_t* sym_base, const float* model,
                                        const double* steps, const float* lut, int B, int H, int W, int K, uint8_t* labels,
                                        float* range_rec, float* xyz, void* book, rpcc_frame_result* results, void* stream) {
  RPCC_REQUIRE(seq_base && sym_base, "null pointer");
  return decode_impl(contour_bits, seq, 0, nullptr, seq_base, symbols, 0, nullptr, sym_base, model, steps, lut, B, H, W, K,
                     labels, range_rec, xyz, book, results, stream);
}
