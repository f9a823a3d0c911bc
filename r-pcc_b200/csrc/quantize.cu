// quantize.cu -- stage 4: intra-prediction + residual quantisation + stable label-major scatter
// + contour bitmap / index sequence, fused in one pass over the range image.
//
// Replaces, per pixel and bit for bit:
//   segment_utils_cpp.intra_predict        (ops/cpp_modules/src/cpp_modules.cpp:248-285)
//   residual = range - pred                (tools/compress.py:106, numpy f32)
//   quantization_utils_cpp.uniform_quantize / nonuniform_quantize   (:288-334 / :337-424):
//       q = (int)roundf(residual / step_l), emitted label-major (0,2,3,..), raster inside a label
//   .astype(np.int16)                      (utils/compress_utils.py:142, wraps)
//   contour_utils_cpp.extract_contour + np.packbits + astype(uint16)   (:521-558, compress_utils.py:155-160)
//
// HBM-bound: reads 4 B (range) + 1 B (label) per pixel, writes 2 B per valid pixel + 1 bit per pixel
// + 2 B per run.  One CTA = one 1024-pixel tile (one pixel per thread).  The stable rank of a pixel
// inside (tile, label) comes from match_any inside its warp plus a per-label scan over the 32 warps
// in shared memory; the tile's base comes from tile_off (model.cu).
#include "book.cuh"

namespace rpcc {

constexpr int kQWarps = 8;                  // one 1024-pixel tile per warp, 8 tiles per CTA
constexpr int kQThreads = kQWarps * 32;
#ifndef RPCC_QSLICES
#define RPCC_QSLICES 4
#endif
#ifndef RPCC_QOCC
#define RPCC_QOCC 6
#endif
constexpr int kQSlices = RPCC_QSLICES;      // 32-pixel slices per step: that many independent loads in flight per lane
constexpr int kQSteps = RPCC_TILE / (32 * kQSlices);

// Per-label table staged in shared memory: .x = the constant prediction m.w, .y = step, .z = 1 / step,
// .w = 1 if the row is a plane (m.x + m.y + m.z != 0, cpp_modules.cpp:271), else 0.
__device__ __forceinline__ float predict_plane(const float4 m, const float* __restrict__ lut3) {
  // cpp_modules.cpp:273-279
  return -m.w / (m.x * __ldg(lut3) + m.y * __ldg(lut3 + 1) + m.z * __ldg(lut3 + 2));
}

// One warp walks one tile in raster order, 32 consecutive pixels (a slice) at a time, and needs no
// block-level synchronisation: its private counters s_cnt[warp][label] start at tile_off (the position,
// in the frame's label-major symbol stream, of the tile's first symbol of that label) and advance as the
// slices go by, so that the stable rank of a pixel is counter + (same-label lanes below it) -- one
// match_any per slice.  Contour bits come from one ballot per slice, MSB-first bytes from brev.
// FULL: the whole tile lies inside the image (no bounds predicates).
template <typename SymT, bool FULL>
__device__ __forceinline__ void quantize_tile(const float* __restrict__ rg, const uint8_t* __restrict__ lb,
                                              const float* __restrict__ lut, const float4* __restrict__ s_model,
                                              const float4* __restrict__ s_tab, unsigned* __restrict__ cnt,
                                              SymT* __restrict__ sym, uint16_t* __restrict__ sq,
                                              uint8_t* __restrict__ cbits, int cbytes, int p_tile, int HW, int W, int K,
                                              unsigned lane) {
  // label left of the tile's first pixel (-1: none, the pixel starts a run anyway)
  int carry = (p_tile > 0 && p_tile < HW) ? (int)lb[p_tile - 1] : -1;
  int next_row = ((p_tile + W - 1) / W) * W;     // next pixel that starts an image row (cpp_modules.cpp:537)
  unsigned nseq = 0;                             // idx_sequence entries emitted by this tile so far
  const unsigned lt = lanemask_lt();

#pragma unroll 1
  for (int s = 0; s < kQSteps; ++s) {
    const int p_step = p_tile + s * (32 * kQSlices);
    if (!FULL && p_step >= HW) break;
    float r[kQSlices];
    int lab[kQSlices];
#pragma unroll
    for (int j = 0; j < kQSlices; ++j) {
      const int p = p_step + j * 32 + (int)lane;
      const bool inb = FULL || p < HW;
      r[j] = inb ? ld_stream_f(rg + p) : 0.f;
      lab[j] = inb ? (int)__ldg(lb + p) : 1;
    }
    unsigned words[kQSlices];
#pragma unroll
    for (int j = 0; j < kQSlices; ++j) {
      const int p0 = p_step + j * 32, p = p0 + (int)lane;
      const bool inb = FULL || p < HW;
      int l = lab[j];
      if (l >= K) l = 1;  // flagged by label_stats; never emitted
      // ---- symbol (cpp_modules.cpp:264-281, tools/compress.py:106, cpp_modules.cpp:311-331)
      const float4 tb = s_tab[l];
      float pred = tb.x;
      if (tb.w != 0.0f) pred = predict_plane(s_model[l], lut + (size_t)p * 3);
      const float res = r[j] - pred;
      // q = (int)roundf(res / step) (cpp_modules.cpp:318).  The reciprocal product t is within 1.8e-7 |t| of the
      // IEEE quotient, so unless t lies within 1e-6 |t| of a rounding boundary (k + 0.5) both round to the same
      // integer -- and away from a boundary round-to-nearest-even IS round-half-away.  The few that are near
      // (and NaN / huge values, whose conversion does not round-trip) take the division.
      const float t = res * tb.z;
      int q = __float2int_rn(t);
      if (!(__fmaf_rn(fabsf(t), 1e-6f, fabsf(t - (float)q)) < 0.5f)) q = (int)roundf(res / tb.y);
      // ---- stable position: private counter + rank among the slice's lanes with the same label
      const unsigned grp = __match_any_sync(0xffffffffu, l);
      const unsigned base = cnt[l];                  // same address for the whole group: one broadcast read
      __syncwarp();
      if ((grp & lt) == 0u) cnt[l] = base + __popc(grp);   // the group's lowest lane
      __syncwarp();
      if (l != 1) sym[base + __popc(grp & lt)] = (SymT)q;  // int16: wraps like astype(np.int16)
      // ---- contour bit (cpp_modules.cpp:534-545) and idx_sequence
      int left = __shfl_up_sync(0xffffffffu, l, 1);
      if (lane == 0) left = carry;
      carry = __shfl_sync(0xffffffffu, l, 31);
      bool rowstart = false;
      if (W >= 32) {
        if (next_row < p0 + 32) { rowstart = (p == next_row); next_row += W; }
      } else {
        rowstart = (p % W) == 0;
      }
      const bool c = inb && (rowstart || l != left);
      const unsigned cb = __ballot_sync(0xffffffffu, c);
      if (c) sq[nseq + __popc(cb & lt)] = (uint16_t)l;
      nseq += __popc(cb);
      words[j] = __byte_perm(__brev(cb), 0, 0x0123);   // np.packbits: pixel p0+i -> byte i/8, bit 7-(i%8)
    }
    // ---- contour bytes of the step: 4 bytes per slice
    const int byte0 = p_step >> 3;
    if ((cbytes & 15) == 0 && (FULL || byte0 + 4 * kQSlices <= cbytes)) {        // every frame's bitmap is 16-byte aligned
      if (lane == 0) {
        uint4* dst = reinterpret_cast<uint4*>(cbits + byte0);
#pragma unroll
        for (int j = 0; j < kQSlices; j += 4) dst[j >> 2] = make_uint4(words[j], words[j + 1], words[j + 2], words[j + 3]);
      }
    } else if ((cbytes & 3) == 0 && byte0 + 4 * kQSlices <= cbytes) {
      if (lane == 0) {
        unsigned* dst = reinterpret_cast<unsigned*>(cbits + byte0);
#pragma unroll
        for (int j = 0; j < kQSlices; ++j) dst[j] = words[j];
      }
    } else {
#pragma unroll
      for (int j = 0; j < kQSlices; ++j) {
        const int b = byte0 + 4 * j + (int)lane;
        if (lane < 4 && b < cbytes) cbits[b] = (uint8_t)(words[j] >> (8 * lane));
      }
    }
  }
}

template <typename SymT>
__global__ void __launch_bounds__(kQThreads, RPCC_QOCC)
quantize_pack_kernel(const float* __restrict__ range, const uint8_t* __restrict__ labels, const float* __restrict__ model,
                     const float* __restrict__ lut, Book bk, const float* __restrict__ step_per_label, float step,
                     int HW, int W, int K, int T, SymT* __restrict__ symbols, size_t sym_stride,
                     uint8_t* __restrict__ contour_bits, int cbytes, uint16_t* __restrict__ seq, size_t seq_stride,
                     const unsigned long long* __restrict__ sym_base, const unsigned long long* __restrict__ seq_base) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float4* s_model = reinterpret_cast<float4*>(smem_raw);            // [K]
  float4* s_tab = s_model + K;                                      // [K] constant prediction, step, 1 / step, plane flag
  unsigned* s_cnt = reinterpret_cast<unsigned*>(s_tab + K);         // [kQWarps][K] next symbol position per label

  const int f = blockIdx.y, tid = threadIdx.x;
  const unsigned lane = tid & 31, warp = tid >> 5;
  const int tile = blockIdx.x * kQWarps + (int)warp;
  for (int l = tid; l < K; l += kQThreads) {
    const float4 m = reinterpret_cast<const float4*>(model)[(size_t)f * K + l];
    s_model[l] = m;
    const float st = step_per_label ? step_per_label[(size_t)f * K + l] : step;
    s_tab[l] = make_float4(m.w, st, 1.0f / st, (m.x + m.y + m.z == 0) ? 0.0f : 1.0f);
  }
  unsigned* cnt = s_cnt + warp * K;
  if (tile < T)
    for (int l = lane; l < K; l += 32) cnt[l] = bk.tile_off[((size_t)f * T + tile) * K + l];
  __syncthreads();
  if (tile >= T) return;

  const size_t fbase = (size_t)f * HW;
  SymT* sym = symbols + (sym_base ? (size_t)sym_base[f] : (size_t)f * sym_stride);
  uint16_t* sq = seq + (seq_base ? (size_t)seq_base[f] : (size_t)f * seq_stride) + bk.tile_coff[(size_t)f * T + tile];
  const int p_tile = tile * RPCC_TILE;
  if (p_tile + RPCC_TILE <= HW)
    quantize_tile<SymT, true>(range + fbase, labels + fbase, lut, s_model, s_tab, cnt, sym, sq,
                              contour_bits + (size_t)f * cbytes, cbytes, p_tile, HW, W, K, lane);
  else
    quantize_tile<SymT, false>(range + fbase, labels + fbase, lut, s_model, s_tab, cnt, sym, sq,
                               contour_bits + (size_t)f * cbytes, cbytes, p_tile, HW, W, K, lane);
}

// exclusive scan of the per-frame symbol / sequence counts (one CTA; B <= 65535)
__global__ void __launch_bounds__(1024)
frame_offsets_kernel(const rpcc_frame_result* __restrict__ results, int B, unsigned long long* __restrict__ sym_base,
                     unsigned long long* __restrict__ seq_base) {
  __shared__ unsigned long long s_a[32], s_b[32];
  __shared__ unsigned long long s_run[2];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) { s_run[0] = 0; s_run[1] = 0; }
  __syncthreads();
  for (int b0 = 0; b0 < B; b0 += 1024) {
    const int b = b0 + tid;
    const unsigned long long a = b < B ? results[b].sym_count : 0ull;
    const unsigned long long c = b < B ? results[b].seq_count : 0ull;
    unsigned long long ia = a, ic = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned long long va = __shfl_up_sync(0xffffffffu, ia, o), vc = __shfl_up_sync(0xffffffffu, ic, o);
      if (lane >= o) { ia += va; ic += vc; }
    }
    if (lane == 31) { s_a[warp] = ia; s_b[warp] = ic; }
    __syncthreads();
    unsigned long long wa = 0, wc = 0;
    for (int q = 0; q < warp; ++q) { wa += s_a[q]; wc += s_b[q]; }
    if (b < B) { sym_base[b] = s_run[0] + wa + ia - a; seq_base[b] = s_run[1] + wc + ic - c; }
    __syncthreads();
    if (tid == 1023) { s_run[0] += wa + ia; s_run[1] += wc + ic; }
    __syncthreads();
  }
  if (tid == 0) { sym_base[B] = s_run[0]; seq_base[B] = s_run[1]; }
}

}  // namespace rpcc

using namespace rpcc;

extern "C" int rpcc_frame_offsets_batch(const rpcc_frame_result* results, int B, uint64_t* sym_base, uint64_t* seq_base,
                                        void* stream) {
  RPCC_REQUIRE(results && sym_base && seq_base, "null pointer");
  frame_offsets_kernel<<<1, 1024, 0, as_stream(stream)>>>(results, B, reinterpret_cast<unsigned long long*>(sym_base),
                                                          reinterpret_cast<unsigned long long*>(seq_base));
  RPCC_LAUNCH_CHECK("frame_offsets_kernel");
  return RPCC_OK;
}

namespace rpcc {
// SymT = int16_t: the bitstream type (utils/compress_utils.py:142); int32_t: what
// uniform_quantize / nonuniform_quantize themselves return (cpp_modules.cpp:322,408).
template <typename SymT>
int quantize_pack_launch(const float* range, const uint8_t* labels, const float* model, const float* lut, void* book,
                         const float* step_per_label, float step, int B, int H, int W, int K, SymT* symbols,
                         size_t sym_stride, uint8_t* contour_bits, uint16_t* seq, size_t seq_stride,
                         const uint64_t* sym_base, const uint64_t* seq_base, void* stream) {
  RPCC_REQUIRE(range && labels && model && lut && book && symbols && contour_bits && seq, "null pointer");
  RPCC_REQUIRE(K >= 2 && K <= 254, "K must be in [2, 254]");
  RPCC_REQUIRE(B <= 65535, "at most 65535 frames per launch");
  if (B == 0) return RPCC_OK;
  const int HW = H * W, T = (HW + RPCC_TILE - 1) / RPCC_TILE;
  const Book bk = make_book(book, B, T, K);
  const size_t smem = 2 * sizeof(float4) * K + sizeof(unsigned) * (size_t)kQWarps * K;
  quantize_pack_kernel<SymT><<<dim3((T + kQWarps - 1) / kQWarps, B), kQThreads, smem, as_stream(stream)>>>(
      range, labels, model, lut, bk, step_per_label, step, HW, W, K, T, symbols, sym_stride, contour_bits,
      (HW + 7) / 8, seq, seq_stride, reinterpret_cast<const unsigned long long*>(sym_base),
      reinterpret_cast<const unsigned long long*>(seq_base));
  RPCC_LAUNCH_CHECK("quantize_pack_kernel");
  return RPCC_OK;
}
template int quantize_pack_launch<int32_t>(const float*, const uint8_t*, const float*, const float*, void*, const float*, float,
                                           int, int, int, int, int32_t*, size_t, uint8_t*, uint16_t*, size_t, const uint64_t*,
                                           const uint64_t*, void*);
}  // namespace rpcc

extern "C" int rpcc_quantize_pack_batch(const float* range, const uint8_t* labels, const float* model, const float* lut,
                                        void* book, const float* step_per_label, float step, int B, int H, int W, int K,
                                        int16_t* symbols, size_t sym_stride, uint8_t* contour_bits, uint16_t* seq,
                                        size_t seq_stride, const uint64_t* sym_base, const uint64_t* seq_base,
                                        void* stream) {
  return quantize_pack_launch<int16_t>(range, labels, model, lut, book, step_per_label, step, B, H, W, K, symbols, sym_stride,
                                       contour_bits, seq, seq_stride, sym_base, seq_base, stream);
}
