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

constexpr int kQThreads = 256;              // 8 warps per 1024-pixel tile
constexpr int kQPer = RPCC_TILE / kQThreads; // 4 pixels per thread, strided by 256 (coalesced)
constexpr int kQChunks = RPCC_TILE / 32;     // 32 warp-sized chunks per tile, chunk = j*8 + warp

__device__ __forceinline__ float predict_range(const float4 m, const float* __restrict__ lut3) {
  // cpp_modules.cpp:271-279
  if (m.x + m.y + m.z == 0) return m.w;
  return -m.w / (m.x * lut3[0] + m.y * lut3[1] + m.z * lut3[2]);
}

template <typename SymT>
__global__ void __launch_bounds__(kQThreads, 4)
quantize_pack_kernel(const float* __restrict__ range, const uint8_t* __restrict__ labels, const float* __restrict__ model,
                     const float* __restrict__ lut, Book bk, const float* __restrict__ step_per_label, float step,
                     int HW, int W, int K, int T, SymT* __restrict__ symbols, size_t sym_stride,
                     uint8_t* __restrict__ contour_bits, int cbytes, uint16_t* __restrict__ seq, size_t seq_stride,
                     const unsigned long long* __restrict__ sym_base, const unsigned long long* __restrict__ seq_base) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float4* s_model = reinterpret_cast<float4*>(smem_raw);            // [K]
  float* s_step = reinterpret_cast<float*>(s_model + K);            // [K]
  unsigned* s_tb = reinterpret_cast<unsigned*>(s_step + K);         // [K] first symbol of (tile,label) in the frame stream
  unsigned* s_last = s_tb + K;                                      // [32] last label of each chunk
  unsigned* s_wc = s_last + kQChunks;                               // [32] contour bits per chunk
  unsigned* s_ccnt = s_wc + kQChunks;                               // [32*Kp/2] per-chunk label counts -> offsets (u16 pairs)
  const int Kp = (K + 1) & ~1;                                      // row pitch in u16, even so rows are u32-aligned
  uint16_t* s_cnt = reinterpret_cast<uint16_t*>(s_ccnt);
  uint16_t* s_tcnt = s_cnt + kQChunks * Kp;                         // [K] pixels of each label in this tile

  const int f = blockIdx.y, tile = blockIdx.x, tid = threadIdx.x;
  const unsigned lane = tid & 31, warp = tid >> 5;
  for (int l = tid; l < K; l += kQThreads) {
    s_model[l] = reinterpret_cast<const float4*>(model)[(size_t)f * K + l];
    s_step[l] = step_per_label ? step_per_label[(size_t)f * K + l] : step;
    s_tb[l] = bk.tile_off[((size_t)f * T + tile) * K + l];
    s_tcnt[l] = bk.tile_hist[((size_t)f * T + tile) * K + l];
  }
  for (int i = tid; i < kQChunks * Kp / 2; i += kQThreads) s_ccnt[i] = 0;
  __syncthreads();

  int label[kQPer], q[kQPer];
  unsigned rank[kQPer];
  const size_t fbase = (size_t)f * HW;
#pragma unroll
  for (int j = 0; j < kQPer; ++j) {
    const int p = tile * RPCC_TILE + j * kQThreads + tid;
    label[j] = 1;
    q[j] = 0;
    if (p < HW) {
      int l = labels[fbase + p];
      if (l >= K) l = 1;  // flagged by label_stats; never emitted
      const float r = range[fbase + p];
      const float pred = predict_range(s_model[l], lut + (size_t)p * 3);
      const float res = r - pred;
      q[j] = (int)roundf(res / s_step[l]);
      label[j] = l;
    }
    const unsigned grp = __match_any_sync(0xffffffffu, label[j]);
    rank[j] = __popc(grp & lanemask_lt());
    const int c = j * (kQThreads / 32) + warp;
    if (rank[j] == 0) s_cnt[c * Kp + label[j]] = (uint16_t)__popc(grp);
    if (lane == 31) s_last[c] = (unsigned)label[j];
  }
  __syncthreads();

  // contour bits (cpp_modules.cpp:534-545), MSB-first packing (np.packbits): pixel p0+i -> byte i/8, bit 7-(i%8)
  unsigned cb[kQPer];
#pragma unroll
  for (int j = 0; j < kQPer; ++j) {
    const int c = j * (kQThreads / 32) + warp;
    const int p = tile * RPCC_TILE + j * kQThreads + tid;
    const bool inb = p < HW;
    int left = __shfl_up_sync(0xffffffffu, label[j], 1);
    if (lane == 0) left = c > 0 ? (int)s_last[c - 1] : (p > 0 && inb ? (int)labels[fbase + p - 1] : -1);
    const bool cbit = inb && ((p % W) == 0 || label[j] != left);
    cb[j] = __ballot_sync(0xffffffffu, cbit);
    if (lane == 0) s_wc[c] = __popc(cb[j]);
    const unsigned word = __byte_perm(__brev(cb[j]), 0, 0x0123);
    const int byte0 = (tile * RPCC_TILE + c * 32) >> 3;
    uint8_t* dst = contour_bits + (size_t)f * cbytes + byte0;
    if ((cbytes & 3) == 0 && byte0 + 4 <= cbytes) {
      if (lane == 0) *reinterpret_cast<unsigned*>(dst) = word;
    } else if (lane < 4 && byte0 + (int)lane < cbytes) {
      dst[lane] = (uint8_t)(word >> (8 * lane));
    }
  }
  // exclusive scan over the 32 chunks, only for the labels present in this tile; warp w owns l = w, w+8, ...
  for (int l = warp; l < K; l += kQThreads / 32) {
    if (s_tcnt[l] == 0) continue;
    const unsigned cnt = s_cnt[lane * Kp + l];
    unsigned incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= (unsigned)o) incl += v;
    }
    s_cnt[lane * Kp + l] = (uint16_t)(incl - cnt);
  }
  __syncthreads();

  const size_t sbase = sym_base ? (size_t)sym_base[f] : (size_t)f * sym_stride;
  const size_t qbase = (seq_base ? (size_t)seq_base[f] : (size_t)f * seq_stride) + bk.tile_coff[(size_t)f * T + tile];
  // contour bits before each chunk
  unsigned cex;
  {
    const unsigned wc = s_wc[lane];
    unsigned incl = wc;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= (unsigned)o) incl += v;
    }
    cex = incl - wc;
  }
#pragma unroll
  for (int j = 0; j < kQPer; ++j) {
    const int c = j * (kQThreads / 32) + warp;
    const int p = tile * RPCC_TILE + j * kQThreads + tid;
    if (p < HW && label[j] != 1) {
      const unsigned pos = s_tb[label[j]] + s_cnt[c * Kp + label[j]] + rank[j];
      symbols[sbase + pos] = (SymT)q[j];  // int16: wraps like astype(np.int16)
    }
    const unsigned before_chunk = __shfl_sync(0xffffffffu, cex, c);
    if ((cb[j] >> lane) & 1u) seq[qbase + before_chunk + __popc(cb[j] & lanemask_lt())] = (uint16_t)label[j];
  }
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
  const int Kp = (K + 1) & ~1;
  const size_t smem = (sizeof(float4) + sizeof(float) + sizeof(unsigned)) * K + sizeof(unsigned) * 2 * kQChunks +
                      sizeof(uint16_t) * ((size_t)kQChunks * Kp + K) + 16;
  quantize_pack_kernel<SymT><<<dim3(T, B), kQThreads, smem, as_stream(stream)>>>(
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
