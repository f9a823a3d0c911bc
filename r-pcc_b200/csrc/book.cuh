// book.cuh -- layout of the per-batch bookkeeping workspace shared by assign / model / quantize.
#pragma once
#include "common.cuh"

namespace rpcc {

struct Book {
  unsigned long long* label_sum;  // [B][K]  sum of range * 2^28 per label (exact)
  unsigned* label_cnt;            // [B][K]  pixels per label
  unsigned* tile_off;             // [B][T][K] first symbol position of (tile,label) in the frame stream
  unsigned* tile_coff;            // [B][T]  contour bits before the tile
  unsigned* flags;                // [B]     bit0: exact-mean guard tripped, bit1: label >= K seen
  uint16_t* tile_hist;            // [B][T][K] pixels per (tile,label)
  uint16_t* tile_ccnt;            // [B][T]  contour bits inside the tile, its first pixel excluded
};

__host__ __device__ inline size_t align16(size_t x) { return (x + 15) & ~(size_t)15; }

inline size_t book_bytes(int B, int T, int K) {
  size_t n = 0;
  n += align16(sizeof(unsigned long long) * (size_t)B * K);
  n += align16(sizeof(unsigned) * (size_t)B * K);
  n += align16(sizeof(unsigned) * (size_t)B * T * K);
  n += align16(sizeof(unsigned) * (size_t)B * T);
  n += align16(sizeof(unsigned) * (size_t)B);
  n += align16(sizeof(uint16_t) * (size_t)B * T * K);
  n += align16(sizeof(uint16_t) * (size_t)B * T);
  return n;
}

inline Book make_book(void* ws, int B, int T, int K) {
  unsigned char* p = static_cast<unsigned char*>(ws);
  Book b;
  b.label_sum = reinterpret_cast<unsigned long long*>(p); p += align16(sizeof(unsigned long long) * (size_t)B * K);
  b.label_cnt = reinterpret_cast<unsigned*>(p);           p += align16(sizeof(unsigned) * (size_t)B * K);
  b.tile_off = reinterpret_cast<unsigned*>(p);            p += align16(sizeof(unsigned) * (size_t)B * T * K);
  b.tile_coff = reinterpret_cast<unsigned*>(p);           p += align16(sizeof(unsigned) * (size_t)B * T);
  b.flags = reinterpret_cast<unsigned*>(p);               p += align16(sizeof(unsigned) * (size_t)B);
  b.tile_hist = reinterpret_cast<uint16_t*>(p);           p += align16(sizeof(uint16_t) * (size_t)B * T * K);
  b.tile_ccnt = reinterpret_cast<uint16_t*>(p);
  return b;
}

// bytes from label_sum up to (not including) tile_off: the part that must be zero before assign
inline size_t book_zero_bytes(int B, int K) {
  return align16(sizeof(unsigned long long) * (size_t)B * K) + align16(sizeof(unsigned) * (size_t)B * K);
}

#ifdef __CUDACC__
// One warp-level pass per distinct label in the warp: count + exact range sum into shared bins.
__device__ __forceinline__ void warp_label_stats(int label, bool active, float r, unsigned* s_cnt,
                                                 unsigned long long* s_sum, unsigned* s_flag) {
  const unsigned lane = threadIdx.x & 31;
  unsigned todo = __ballot_sync(0xffffffffu, active);
  const bool exact = !(active && label >= 2) || (r >= 0.03125f && r < 256.0f);
  if (__any_sync(0xffffffffu, !exact) && lane == 0) atomicOr(s_flag, 1u);
  const unsigned long long v = active ? (unsigned long long)((double)r * 268435456.0) : 0ull;
  while (todo) {
    const int leader = __ffs(todo) - 1;
    const int l = __shfl_sync(0xffffffffu, label, leader);
    const bool mine = active && label == l;
    const unsigned grp = __ballot_sync(0xffffffffu, mine);
    // v < 2^36: split so that 32 addends cannot overflow 32 bits
    const unsigned lo = mine ? (unsigned)(v & 0xFFFFFu) : 0u;
    const unsigned hi = mine ? (unsigned)(v >> 20) : 0u;
    const unsigned slo = __reduce_add_sync(0xffffffffu, lo);
    const unsigned shi = __reduce_add_sync(0xffffffffu, hi);
    if (lane == (unsigned)leader) {
      atomicAdd(&s_cnt[l], (unsigned)__popc(grp));
      if (l >= 2) atomicAdd(&s_sum[l], ((unsigned long long)shi << 20) + slo);
    }
    todo &= ~grp;
  }
}

// Contour bits inside the tile (extract_contour, cpp_modules.cpp:534-545): a pixel starts a run when
// it is in column 0 or its label differs from its left neighbour.  The tile's first pixel needs the
// previous tile's last label, so it is left to model.cu; everything else is counted here.
__device__ __forceinline__ void tile_contour_count(int label, bool inb, int p, int W, unsigned* s_last, unsigned* s_ccnt) {
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 31) s_last[warp] = (unsigned)label;
  __syncthreads();
  int left = __shfl_up_sync(0xffffffffu, label, 1);
  if (lane == 0 && warp > 0) left = (int)s_last[warp - 1];
  const bool c = inb && threadIdx.x > 0 && ((p % W) == 0 || label != left);
  const unsigned b = __ballot_sync(0xffffffffu, c);
  if (lane == 0 && b) atomicAdd(s_ccnt, (unsigned)__popc(b));
}

__device__ __forceinline__ void flush_tile_stats(int K, int f, int tile, int T, const unsigned* s_cnt,
                                                 const unsigned long long* s_sum, const unsigned* s_flag,
                                                 const unsigned* s_ccnt, const Book& bk) {
  for (int l = threadIdx.x; l < K; l += blockDim.x) {
    const unsigned c = s_cnt[l];
    bk.tile_hist[((size_t)f * T + tile) * K + l] = (uint16_t)c;
    if (c) {
      atomicAdd(&bk.label_cnt[(size_t)f * K + l], c);
      if (l >= 2) atomicAdd(&bk.label_sum[(size_t)f * K + l], s_sum[l]);
    }
  }
  if (threadIdx.x == 0) {
    bk.tile_ccnt[(size_t)f * T + tile] = (uint16_t)*s_ccnt;
    if (*s_flag) atomicOr(&bk.flags[f], *s_flag);
  }
}


// Stable rank of each thread's pixel inside (tile, label): match_any inside the warp plus a
// per-label exclusive scan over the 32 warps.  s_wcnt is [32][K] u16 scratch (zeroed here).
// Contains three __syncthreads; every thread of a 1024-thread CTA must call it.
__device__ __forceinline__ unsigned tile_label_rank(int label, int K, uint16_t* s_wcnt) {
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 32 * K; i += blockDim.x) s_wcnt[i] = 0;
  __syncthreads();
  const unsigned grp = __match_any_sync(0xffffffffu, label);
  const unsigned rank_in_warp = __popc(grp & lanemask_lt());
  if (rank_in_warp == 0) s_wcnt[warp * K + label] = (uint16_t)__popc(grp);
  __syncthreads();
  for (int l = warp; l < K; l += 32) {
    const unsigned c = s_wcnt[lane * K + l];
    unsigned incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= (unsigned)o) incl += v;
    }
    s_wcnt[lane * K + l] = (uint16_t)(incl - c);
  }
  __syncthreads();
  return s_wcnt[warp * K + label] + rank_in_warp;
}
#endif  // __CUDACC__

}  // namespace rpcc
