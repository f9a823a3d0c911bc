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

}  // namespace rpcc
