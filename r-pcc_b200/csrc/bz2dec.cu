// bz2dec.cu -- a bzip2 decoder for the sections of a `.rpcc` (the inverse of bz2enc.cu / libbz2's
// BZ2_bzBuffToBuffDecompress, which is what the reference's bz2.decompress runs: utils/compress_utils.py:300-302).
// Host code only.
//
// Why: the decode tools are bound by the entropy decoder on the host threads (5-6 ms per 64E frame
// and core in libbz2, the device stages take 1.2 us).  libbz2 decodes one Huffman symbol per loop trip with a
// length-by-length search and refills its bit buffer a byte at a time; here the codes are looked up in a 2^11-entry table
// (symbol and length in one load, longer codes fall back to the canonical limit search), the bit buffer is 64 bits wide,
// and the move-to-front list is a plain byte array with the small-position case inlined.  The inverse transform is the
// textbook one (bzip2's "fast" mode: one 32-bit word per byte, pointer chase).
//
// What is NOT handled: "randomised" blocks (bzip2 < 0.9.5) and anything malformed -- rpcc_bz2_decompress returns
// RPCC_BZ2_DECLINED and the caller gives the stream to libbz2, whose verdict (and error code) is then the one reported.
// Both CRCs are checked.  tests/test_host_logic.py decodes thousands of bz2.compress outputs and compares.
#include <stdlib.h>
#include <string.h>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif

#include <vector>

#include "common.cuh"

namespace {

constexpr int kMaxAlpha = 258, kGroups = 6, kGroupSize = 50, kMaxSelectors = 18002, kFastBits = 11;

// bzip2's CRC (MSB-first, polynomial 0x04c11db7), four bytes per step: t[k][b] = the CRC of byte b followed by k zero bytes
struct CrcTable {
  unsigned t[4][256];
  CrcTable() {
    for (unsigned i = 0; i < 256; ++i) {
      unsigned c = i << 24;
      for (int k = 0; k < 8; ++k) c = (c & 0x80000000u) ? (c << 1) ^ 0x04c11db7u : (c << 1);
      t[0][i] = c;
    }
    for (int k = 1; k < 4; ++k)
      for (unsigned i = 0; i < 256; ++i) t[k][i] = (t[k - 1][i] << 8) ^ t[0][t[k - 1][i] >> 24];
  }
};
const CrcTable g_crc;
inline unsigned crc_update(unsigned crc, const unsigned char* p, size_t n) {
  while (n >= 4) {
    crc ^= ((unsigned)p[0] << 24) | ((unsigned)p[1] << 16) | ((unsigned)p[2] << 8) | (unsigned)p[3];
    crc = g_crc.t[3][crc >> 24] ^ g_crc.t[2][(crc >> 16) & 255u] ^ g_crc.t[1][(crc >> 8) & 255u] ^ g_crc.t[0][crc & 255u];
    p += 4; n -= 4;
  }
  while (n--) crc = (crc << 8) ^ g_crc.t[0][(crc >> 24) ^ *p++];
  return crc;
}

// MSB-first bit reader over a byte buffer; past the end it reads zeros, and over() tells whether any of them were consumed
struct BitReader {
  const unsigned char* src;
  size_t n, pos = 0;                           // bytes fetched so far (may run past n: those are the zeros)
  unsigned long long buf = 0;
  int bits = 0;
  BitReader(const unsigned char* s, size_t len) : src(s), n(len) {}
  inline void refill() {
    if (pos + 8 <= n) {                        // eight bytes at once, keep the whole ones that fit
      unsigned long long w;
      memcpy(&w, src + pos, 8);
      w = __builtin_bswap64(w);
      buf |= bits ? (w >> bits) : w;
      const int take = (64 - bits) >> 3;
      pos += (size_t)take;
      bits += take << 3;
      return;
    }
    while (bits <= 56) {
      const unsigned long long b = pos < n ? src[pos] : 0;
      ++pos;
      buf |= b << (56 - bits);
      bits += 8;
    }
  }
  inline unsigned peek(int k) { return (unsigned)(buf >> (64 - k)); }          // k in [1, 32], bits >= k
  inline void drop(int k) { buf <<= k; bits -= k; }
  inline unsigned get(int k) { if (bits < k) refill(); const unsigned v = peek(k); drop(k); return v; }
  inline bool over() const { return pos * 8 - (size_t)bits > n * 8; }
};

struct Table {
  unsigned short fast[1 << kFastBits];      // (symbol << 5) | length, 0 = longer than kFastBits
  int limit[22], base[22], perm[kMaxAlpha];
  int minLen, maxLen;
};

// canonical code of bzip2 (huffman.c BZ2_hbCreateDecodeTables / BZ2_hbAssignCodes): by length, then by symbol
bool build_table(Table& t, const unsigned char* len, int alphaSize) {
  int minLen = 32, maxLen = 0;
  for (int v = 0; v < alphaSize; ++v) { if (len[v] > maxLen) maxLen = len[v]; if (len[v] < minLen) minLen = len[v]; }
  if (minLen < 1 || maxLen > 20) return false;
  t.minLen = minLen; t.maxLen = maxLen;
  int pp = 0;
  for (int l = minLen; l <= maxLen; ++l) for (int v = 0; v < alphaSize; ++v) if (len[v] == l) t.perm[pp++] = v;
  int count[22];
  for (int l = 0; l < 22; ++l) count[l] = 0;
  for (int v = 0; v < alphaSize; ++v) ++count[len[v]];
  // limit[l] = largest code of length l (left-aligned comparisons are done on l-bit values); base[l] = first index - first code
  int code = 0, idx = 0;
  for (int l = 0; l < 22; ++l) { t.limit[l] = -1; t.base[l] = 0; }
  memset(t.fast, 0, sizeof(t.fast));
  for (int l = minLen; l <= maxLen; ++l) {
    t.base[l] = idx - code;
    for (int k = 0; k < count[l]; ++k) {
      if (l <= kFastBits) {
        const unsigned first = (unsigned)code << (kFastBits - l), n = 1u << (kFastBits - l);
        if (first + n > (1u << kFastBits)) return false;                   // over-subscribed code
        const unsigned short e = (unsigned short)((t.perm[idx] << 5) | l);
        for (unsigned q = 0; q < n; ++q) t.fast[first + q] = e;
      }
      ++code; ++idx;
    }
    t.limit[l] = code - 1;
    if (code > (1 << l)) return false;
    code <<= 1;
  }
  return true;
}

// g_mtf_mask + 15 - pos: sixteen bytes, 0xFF at index <= pos
alignas(16) const unsigned char g_mtf_mask[32] = {255, 255, 255, 255, 255, 255, 255, 255, 255, 255, 255, 255, 255, 255, 255, 255,
                                                  0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};

struct Scratch {
  std::vector<unsigned> tt;
  std::vector<Table> tab;
};

}  // namespace

// dst receives the decoded bytes of the .bz2 stream src[0..n).  Returns RPCC_OK, RPCC_BZ2_DECLINED (hand the stream to
// libbz2: a feature this decoder leaves out, or a stream it finds malformed) or RPCC_ERR_CAPACITY (dst too small).
extern "C" int rpcc_bz2_decompress(const uint8_t* src, size_t n, uint8_t* dst, size_t cap, size_t* out_len) {
  if (!src || !dst || !out_len) return RPCC_ERR_ARG;
  static thread_local Scratch S;
  if (n < 14 || src[0] != 'B' || src[1] != 'Z' || src[2] != 'h' || src[3] < '1' || src[3] > '9') return RPCC_BZ2_DECLINED;
  const int blockMax = 100000 * (src[3] - '0');
  BitReader br(src + 4, n - 4);
  size_t out = 0;
  unsigned combined = 0;
  if (S.tab.size() < (size_t)kGroups) S.tab.resize(kGroups);
  for (;;) {
    br.refill();
    const unsigned m1 = br.get(24), m2 = br.get(24);
    if (m1 == 0x177245u && m2 == 0x385090u) {                    // end of stream
      const unsigned want = br.get(32);
      if (br.over() || want != combined) return RPCC_BZ2_DECLINED;
      *out_len = out;
      return RPCC_OK;                                             // (bytes after the stream are ignored, as BuffToBuff does)
    }
    if (m1 != 0x314159u || m2 != 0x265359u) return RPCC_BZ2_DECLINED;
    const unsigned blockCrc = br.get(32);
    if (br.get(1)) return RPCC_BZ2_DECLINED;                      // randomised block
    const unsigned origPtr = br.get(24);
    // ---- symbol map
    unsigned char seqToUnseq[256];
    int nInUse = 0;
    {
      const unsigned used16 = br.get(16);
      for (int a = 0; a < 16; ++a) {
        if (used16 & (0x8000u >> a)) {
          const unsigned bitsA = br.get(16);
          for (int b = 0; b < 16; ++b) if (bitsA & (0x8000u >> b)) seqToUnseq[nInUse++] = (unsigned char)(a * 16 + b);
        }
      }
    }
    if (nInUse == 0) return RPCC_BZ2_DECLINED;
    const int alphaSize = nInUse + 2, EOB = nInUse + 1;
    const int nGroups = (int)br.get(3);
    const int nSelectors = (int)br.get(15);
    if (nGroups < 2 || nGroups > kGroups || nSelectors < 1 || nSelectors > kMaxSelectors) return RPCC_BZ2_DECLINED;
    static thread_local unsigned char selector[kMaxSelectors];
    {
      unsigned char pos[kGroups];
      for (int g = 0; g < nGroups; ++g) pos[g] = (unsigned char)g;
      for (int q = 0; q < nSelectors; ++q) {
        int j = 0;
        while (br.get(1)) { if (++j >= nGroups) return RPCC_BZ2_DECLINED; }
        const unsigned char tmp = pos[j];
        for (; j > 0; --j) pos[j] = pos[j - 1];
        pos[0] = tmp;
        selector[q] = tmp;
      }
    }
    for (int g = 0; g < nGroups; ++g) {
      unsigned char len[kMaxAlpha];
      int cur = (int)br.get(5);
      for (int v = 0; v < alphaSize; ++v) {
        for (;;) {
          if (cur < 1 || cur > 20) return RPCC_BZ2_DECLINED;
          if (!br.get(1)) break;
          cur += br.get(1) ? -1 : 1;
        }
        len[v] = (unsigned char)cur;
      }
      if (!build_table(S.tab[g], len, alphaSize)) return RPCC_BZ2_DECLINED;
    }
    if (br.over()) return RPCC_BZ2_DECLINED;
    // ---- Huffman symbols -> move-to-front -> the last column, one word per byte, with the byte counts
    if (S.tt.size() < (size_t)blockMax) S.tt.resize(blockMax);
    unsigned* tt = S.tt.data();
    int unzftab[256];
    for (int q = 0; q < 256; ++q) unzftab[q] = 0;
    alignas(16) unsigned char yy[256 + 16];
    for (int q = 0; q < 256 + 16; ++q) yy[q] = (unsigned char)q;
    int nblock = 0, groupNo = -1, groupPos = 0;
    const Table* T = nullptr;
    int runLen = 0, runBit = 1;                                   // pending zero run: RUNA adds runBit, RUNB 2 * runBit
    for (;;) {
      if (groupPos == 0) {
        if (++groupNo >= nSelectors) return RPCC_BZ2_DECLINED;
        groupPos = kGroupSize;
        T = &S.tab[selector[groupNo]];
      }
      --groupPos;
      if (br.bits < 32) br.refill();
      int sym;
      {
        const unsigned short e = T->fast[br.peek(kFastBits)];
        if (e) {
          sym = e >> 5;
          br.drop(e & 31);
        } else {
          int l = kFastBits + 1;
          if (l < T->minLen) l = T->minLen;
          for (;; ++l) {
            if (l > T->maxLen) return RPCC_BZ2_DECLINED;
            const int code = (int)br.peek(l);
            if (code <= T->limit[l]) { const int ix = code + T->base[l]; if (ix < 0 || ix >= alphaSize) return RPCC_BZ2_DECLINED; sym = T->perm[ix]; break; }
          }
          br.drop(l);
        }
      }
      if (sym <= 1) {                                             // RUNA / RUNB
        if (runBit > (1 << 21)) return RPCC_BZ2_DECLINED;
        runLen += runBit << sym;
        runBit <<= 1;
        continue;
      }
      if (runLen) {
        if (nblock + runLen > blockMax) return RPCC_BZ2_DECLINED;
        const unsigned char uc = seqToUnseq[yy[0]];
        unzftab[uc] += runLen;
        for (int q = 0; q < runLen; ++q) tt[nblock + q] = uc;
        nblock += runLen;
        runLen = 0; runBit = 1;
      }
      if (sym == EOB) break;
      if (nblock >= blockMax) return RPCC_BZ2_DECLINED;
      {
        const int pos = sym - 1;                                  // 1 .. nInUse - 1
        const unsigned char v = yy[pos];
#if defined(__SSE2__)
        if (pos < 16) {
          // the first 16 entries in one register: shifted up by one where the index is <= pos, the moved value in front
          const __m128i cur = _mm_load_si128(reinterpret_cast<const __m128i*>(yy));
          const __m128i sh = _mm_or_si128(_mm_slli_si128(cur, 1), _mm_cvtsi32_si128((int)v));
          const __m128i m = _mm_loadu_si128(reinterpret_cast<const __m128i*>(g_mtf_mask + 15 - pos));   // 0xFF for index <= pos
          _mm_store_si128(reinterpret_cast<__m128i*>(yy), _mm_or_si128(_mm_and_si128(sh, m), _mm_andnot_si128(m, cur)));
        } else {
          memmove(yy + 1, yy, (size_t)pos);
          yy[0] = v;
        }
#else
        if (pos < 8) { for (int q = pos; q > 0; --q) yy[q] = yy[q - 1]; }
        else memmove(yy + 1, yy, (size_t)pos);
        yy[0] = v;
#endif
        const unsigned char uc = seqToUnseq[v];
        ++unzftab[uc];
        tt[nblock++] = uc;
      }
    }
    if (br.over() || origPtr >= (unsigned)nblock) return RPCC_BZ2_DECLINED;
    // ---- inverse transform (decompress.c, the fast path) + the initial run-length layer + CRC
    {
      int cftab[257];
      cftab[0] = 0;
      for (int q = 1; q <= 256; ++q) cftab[q] = cftab[q - 1] + unzftab[q - 1];
      for (int i = 0; i < nblock; ++i) {
        const unsigned uc = tt[i] & 0xffu;
        tt[cftab[uc]++] |= (unsigned)i << 8;
      }
    }
    const size_t out0 = out;
    unsigned tPos = tt[origPtr] >> 8;
    int run = 0, prev = -1;
    for (int i = 0; i < nblock; ++i) {
      tPos = tt[tPos];
      const unsigned char ch = (unsigned char)(tPos & 0xffu);
      tPos >>= 8;
      if (run == 4) {                                             // a count byte: ch more copies of prev
        if (out + ch > cap) return RPCC_ERR_CAPACITY;
        memset(dst + out, prev, ch);
        out += ch;
        run = 0; prev = -1;
        continue;
      }
      if (out >= cap) return RPCC_ERR_CAPACITY;
      dst[out++] = ch;
      if (ch == prev) ++run; else { run = 1; prev = ch; }
    }
    const unsigned crc = ~crc_update(0xffffffffu, dst + out0, out - out0);   // (off the pointer chase's dependency chain)
    if (crc != blockCrc) return RPCC_BZ2_DECLINED;
    combined = ((combined << 1) | (combined >> 31)) ^ crc;
  }
}
