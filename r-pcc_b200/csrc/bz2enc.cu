// bz2enc.cu -- a bzip2 (level 9) encoder for the sections of a `.rpcc`, byte-identical to libbz2's
// BZ2_bzBuffToBuffCompress(..., 9, 0, workFactor) and therefore to CPython's bz2.compress, which is what the reference
// writes (utils/compress_utils.py:296-298).  Host code only.
//
// Why: the entropy coder is the whole cost of the batch tool once the GPU stages are batched (about 20 ms per 64E frame
// and core), and three quarters of libbz2's time on these sections goes into its block sort, which
// degrades on exactly this kind of input -- int16 residuals (every other byte 0x00 / 0xFF) and label sequences are
// long chains of repeated contexts.  The bitstream does not depend on HOW the rotations are sorted, only on the sorted
// order, so the sort is replaced: the block is rotated to its least rotation (then it is a Lyndon word, for which the
// order of the cyclic rotations equals the order of the suffixes) and the suffixes are sorted in linear time by induced
// sorting (SA-IS, Nong / Zhang / Chan 2009).  Everything else -- the initial run-length pass, move-to-front with
// RUNA / RUNB, the six-table Huffman optimisation with libbz2's own code-length routine, selectors, CRCs, bit layout --
// follows bzip2 1.0.x step for step (compress.c, huffman.c, bzlib.c), because every choice there shows in the bytes.
//
// What is NOT reproduced: the order libbz2 gives to IDENTICAL rotations (a block that is a whole number of repetitions
// of a shorter string; it changes origPtr) and inputs beyond one 900 kB block.  rpcc_bz2_compress returns
// RPCC_BZ2_DECLINED for those and the caller uses libbz2 (hostio.cu).  tests/test_host_logic.py compares the bytes
// with bz2.compress on thousands of inputs.
#include <stdlib.h>
#include <string.h>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif

#include <vector>
#ifdef RPCC_BZ2_PROFILE
#include <chrono>
#include <stdio.h>
static double g_t[8];
static inline double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
#define TICK(i) do { double _n = now(); g_t[i] += _n - _t0; _t0 = _n; } while (0)
extern "C" void rpcc_bz2_profile_dump() { fprintf(stderr, "rle %.3f periodic %.3f rot %.3f sais %.3f mtf %.3f huff %.3f write %.3f ms\n", g_t[0]*1e3, g_t[1]*1e3, g_t[2]*1e3, g_t[3]*1e3, g_t[4]*1e3, g_t[5]*1e3, g_t[6]*1e3); for (int i=0;i<8;++i) g_t[i]=0; }
#else
#define TICK(i)
#endif

#include "common.cuh"

namespace {

constexpr int kRunA = 0, kRunB = 1;
constexpr int kMaxAlpha = 258, kGroups = 6, kGroupSize = 50, kIters = 4, kMaxSelectors = 18002;
constexpr int kBlockMax = 900000 - 19;
// kMtfMask + 15 - pos: sixteen bytes, 0xFF at index <= pos
alignas(16) const unsigned char kMtfMask[32] = {255, 255, 255, 255, 255, 255, 255, 255, 255, 255, 255, 255, 255, 255, 255, 255,
                                                0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};

// ------------------------------------------------------------------------------------------------ CRC (bzlib crctable.c)
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

// ------------------------------------------------------------------------------------------------ suffix sorting (SA-IS)
// s[0..n): symbols in [0, K), s[n-1] = 0 is the unique smallest one.  SA receives the suffix array.  The string is
// packed in place -- symbol << 1 | type (1 = S-type, 0 = L-type) -- so that the induction passes, whose accesses into
// the string are random, touch one array instead of two (they are what the sort spends its time on).
inline void get_buckets(const int* cnt, int* bkt, int K, bool end) {
  int sum = 0;
  for (int i = 0; i < K; ++i) { sum += cnt[i]; bkt[i] = end ? sum : sum - cnt[i]; }
}
template <typename Ch>
void induce_l(int* SA, const Ch* s, const int* cnt, int* bkt, int n, int K) {
  get_buckets(cnt, bkt, K, false);
  for (int i = 0; i < n; ++i) {
    const int j = SA[i] - 1;
    if (j >= 0) {
      const unsigned v = (unsigned)s[j];
      if (!(v & 1u)) SA[bkt[v >> 1]++] = j;
    }
  }
}
template <typename Ch>
void induce_s(int* SA, const Ch* s, const int* cnt, int* bkt, int n, int K) {
  get_buckets(cnt, bkt, K, true);
  for (int i = n - 1; i >= 0; --i) {
    const int j = SA[i] - 1;
    if (j >= 0) {
      const unsigned v = (unsigned)s[j];
      if (v & 1u) SA[--bkt[v >> 1]] = j;
    }
  }
}
template <typename Ch>
void sais(Ch* s, int* SA, int n, int K) {
  std::vector<int> bucket((size_t)K * 2);
  int* bkt = bucket.data();
  int* cnt = bkt + K;
  for (int i = 0; i < K; ++i) cnt[i] = 0;
  for (int i = 0; i < n; ++i) ++cnt[s[i]];
  {                                                        // types, packed into the string (back to front)
    bool tn = true;
    unsigned nxt = (unsigned)s[n - 1];
    s[n - 1] = (Ch)((nxt << 1) | 1u);
    for (int i = n - 2; i >= 0; --i) {
      const unsigned cur = (unsigned)s[i];
      const bool ti = cur < nxt || (cur == nxt && tn);
      s[i] = (Ch)((cur << 1) | (ti ? 1u : 0u));
      tn = ti;
      nxt = cur;
    }
  }
  auto is_lms = [&](int i) { return i > 0 && (s[i] & 1) && !(s[i - 1] & 1); };
  // stage 1: sort the LMS substrings
  get_buckets(cnt, bkt, K, true);
  for (int i = 0; i < n; ++i) SA[i] = -1;
  for (int i = 1; i < n; ++i) if ((s[i] & 1) && !(s[i - 1] & 1)) SA[--bkt[s[i] >> 1]] = i;
  induce_l(SA, s, cnt, bkt, n, K);
  induce_s(SA, s, cnt, bkt, n, K);
  int n1 = 0;
  for (int i = 0; i < n; ++i) if (is_lms(SA[i])) SA[n1++] = SA[i];
  for (int i = n1; i < n; ++i) SA[i] = -1;
  int name = 0, prev = -1;
  for (int i = 0; i < n1; ++i) {
    int pos = SA[i];
    bool diff = false;
    for (int d = 0; d < n; ++d) {
      if (prev == -1 || s[pos + d] != s[prev + d]) { diff = true; break; }       // symbol and type at once
      if (d > 0 && (is_lms(pos + d) || is_lms(prev + d))) break;
    }
    if (diff) { ++name; prev = pos; }
    pos >>= 1;
    SA[n1 + pos] = name - 1;
  }
  for (int i = n - 1, j = n - 1; i >= n1; --i) if (SA[i] >= 0) SA[j--] = SA[i];
  // stage 2: the reduced problem
  int* SA1 = SA;
  int* s1 = SA + n - n1;
  if (name < n1) {
    sais<int>(s1, SA1, n1, name);
  } else {
    for (int i = 0; i < n1; ++i) SA1[s1[i]] = i;
  }
  // stage 3: induce the result
  get_buckets(cnt, bkt, K, true);
  for (int i = 1, j = 0; i < n; ++i) if ((s[i] & 1) && !(s[i - 1] & 1)) s1[j++] = i;
  for (int i = 0; i < n1; ++i) SA1[i] = s1[SA1[i]];
  for (int i = n1; i < n; ++i) SA[i] = -1;
  for (int i = n1 - 1; i >= 0; --i) {
    const int j = SA[i];
    SA[i] = -1;
    SA[--bkt[s[j] >> 1]] = j;
  }
  induce_l(SA, s, cnt, bkt, n, K);
  induce_s(SA, s, cnt, bkt, n, K);
}

// start of the lexicographically least rotation
// d = the block written twice (2n bytes + 8 of padding): no index ever wraps, and the common prefix of two rotations is
// measured eight bytes at a time
int least_rotation(const unsigned char* d, int n) {
  int i = 0, j = 1, k = 0;
  while (i < n && j < n && k < n) {
    unsigned long long a, b;
    memcpy(&a, d + i + k, 8);
    memcpy(&b, d + j + k, 8);
    if (a == b) { k += 8; continue; }
    k += (int)(__builtin_ctzll(a ^ b) >> 3);                  // little-endian: the lowest differing byte comes first
    if (k >= n) break;
    if (d[i + k] > d[j + k]) i += k + 1; else j += k + 1;
    if (i == j) ++j;
    k = 0;
  }
  return i < j ? i : j;
}

// true if the block is a whole number (>= 2) of repetitions of a shorter string: some proper divisor p of n is a period.
// (A prefix-function pass over the block answered the same question in 0.6 ms per frame; a mismatch within the first few
// bytes settles nearly every divisor.)
bool is_periodic(const unsigned char* s, int n) {
  for (int d = 1; (long long)d * d <= n; ++d) {
    if (n % d) continue;
    const int e = n / d;
    if (d < n && memcmp(s, s + d, (size_t)(n - d)) == 0) return true;
    if (e < n && e != d && memcmp(s, s + e, (size_t)(n - e)) == 0) return true;
  }
  return false;
}

// ------------------------------------------------------------------------------------------------ bit writer (bzlib bsW)
struct BitWriter {
  unsigned char* out;
  size_t cap, pos = 0;
  unsigned buff = 0;
  int live = 0;
  bool overflow = false;
  inline void put(int n, unsigned v) {
    while (live >= 8) {
      if (pos < cap) out[pos++] = (unsigned char)(buff >> 24); else overflow = true;
      buff <<= 8;
      live -= 8;
    }
    buff |= v << (32 - live - n);
    live += n;
  }
  inline void put_u8(unsigned c) { put(8, c); }
  inline void put_u32(unsigned u) { put(8, (u >> 24) & 0xff); put(8, (u >> 16) & 0xff); put(8, (u >> 8) & 0xff); put(8, u & 0xff); }
  void finish() {
    while (live > 0) {
      if (pos < cap) out[pos++] = (unsigned char)(buff >> 24); else overflow = true;
      buff <<= 8;
      live -= 8;
    }
  }
};

// ------------------------------------------------------------------------------------------------ huffman.c
void make_code_lengths(unsigned char* len, const int* freq, int alphaSize, int maxLen) {
  int heap[kMaxAlpha + 2], weight[kMaxAlpha * 2], parent[kMaxAlpha * 2];
  for (int i = 0; i < alphaSize; ++i) weight[i + 1] = (freq[i] == 0 ? 1 : freq[i]) << 8;
  for (;;) {
    int nNodes = alphaSize, nHeap = 0;
    heap[0] = 0; weight[0] = 0; parent[0] = -2;
    auto upheap = [&](int z) {
      int zz = z;
      const int tmp = heap[zz];
      while (weight[tmp] < weight[heap[zz >> 1]]) { heap[zz] = heap[zz >> 1]; zz >>= 1; }
      heap[zz] = tmp;
    };
    auto downheap = [&](int z) {
      int zz = z;
      const int tmp = heap[zz];
      for (;;) {
        int yy = zz << 1;
        if (yy > nHeap) break;
        if (yy < nHeap && weight[heap[yy + 1]] < weight[heap[yy]]) ++yy;
        if (weight[tmp] < weight[heap[yy]]) break;
        heap[zz] = heap[yy];
        zz = yy;
      }
      heap[zz] = tmp;
    };
    for (int i = 1; i <= alphaSize; ++i) { parent[i] = -1; ++nHeap; heap[nHeap] = i; upheap(nHeap); }
    while (nHeap > 1) {
      const int n1 = heap[1]; heap[1] = heap[nHeap]; --nHeap; downheap(1);
      const int n2 = heap[1]; heap[1] = heap[nHeap]; --nHeap; downheap(1);
      ++nNodes;
      parent[n1] = parent[n2] = nNodes;
      const int d1 = weight[n1] & 0xff, d2 = weight[n2] & 0xff;
      weight[nNodes] = (int)(((unsigned)weight[n1] & 0xffffff00u) + ((unsigned)weight[n2] & 0xffffff00u)) | (1 + (d1 > d2 ? d1 : d2));
      parent[nNodes] = -1;
      ++nHeap; heap[nHeap] = nNodes; upheap(nHeap);
    }
    bool tooLong = false;
    for (int i = 1; i <= alphaSize; ++i) {
      int j = 0, k = i;
      while (parent[k] >= 0) { k = parent[k]; ++j; }
      len[i - 1] = (unsigned char)j;
      if (j > maxLen) tooLong = true;
    }
    if (!tooLong) break;
    for (int i = 1; i <= alphaSize; ++i) {
      int j = weight[i] >> 8;
      j = 1 + (j / 2);
      weight[i] = j << 8;
    }
  }
}

struct Scratch {
  std::vector<unsigned char> block, rot;
  std::vector<int> sa, pi;
  std::vector<unsigned short> mtfv;
};

}  // namespace

// dst receives the complete .bz2 stream of src[0..n).  Returns RPCC_OK, RPCC_BZ2_DECLINED (the caller should use
// libbz2: periodic block or more than one block) or RPCC_ERR_CAPACITY.
extern "C" int rpcc_bz2_compress(const uint8_t* src, size_t n, uint8_t* dst, size_t cap, size_t* out_len) {
  RPCC_REQUIRE((src || n == 0) && dst && out_len, "null pointer");
  static thread_local Scratch S;
  BitWriter bw{dst, cap};
  bw.put_u8('B'); bw.put_u8('Z'); bw.put_u8('h'); bw.put_u8('9');
  unsigned combined = 0;
#ifdef RPCC_BZ2_PROFILE
  double _t0 = now();
#endif
  if (n > 0) {
    if (n > (size_t)kBlockMax - 1024) return RPCC_BZ2_DECLINED;   // more than one block (or close to it): libbz2's business
    // ---- initial run-length pass (bzlib.c add_pair_to_block) + block CRC over the original bytes
    std::vector<unsigned char>& blk = S.block;
    if (blk.size() < n + n / 4 + 16) blk.resize(n + n / 4 + 16);   // scratch buffers only ever grow: a resize back up would zero-fill
    unsigned char* w = blk.data();
    bool inUse[256];
    memset(inUse, 0, sizeof(inUse));
    unsigned crc = crc_update(0xffffffffu, src, n);
    size_t i = 0;
    while (i < n) {
      const unsigned char ch = src[i];
      size_t run = 1;
      while (i + run < n && src[i + run] == ch && run < 255) ++run;
      inUse[ch] = true;
      if (run < 4) {
        for (size_t q = 0; q < run; ++q) *w++ = ch;
      } else {
        inUse[run - 4] = true;
        w[0] = ch; w[1] = ch; w[2] = ch; w[3] = ch; w[4] = (unsigned char)(run - 4);
        w += 5;
      }
      i += run;
    }
    const size_t blk_len = (size_t)(w - blk.data());
    crc = ~crc;
    combined = ((combined << 1) | (combined >> 31)) ^ crc;
    const int nb = (int)blk_len;
    if (nb > kBlockMax - 1024) return RPCC_BZ2_DECLINED;
    const unsigned char* block = blk.data();
    TICK(0);
    if (is_periodic(block, nb)) return RPCC_BZ2_DECLINED;
    TICK(1);
    // ---- sorted rotations: suffix array of the least rotation (a Lyndon word) with a sentinel
    if (S.rot.size() < (size_t)2 * nb + 8) S.rot.resize((size_t)2 * nb + 8);
    memcpy(S.rot.data(), block, (size_t)nb);
    memcpy(S.rot.data() + nb, block, (size_t)nb);
    memset(S.rot.data() + 2 * (size_t)nb, 0, 8);
    const int r = least_rotation(S.rot.data(), nb);
    const unsigned char* rot = S.rot.data() + r;             // the least rotation, nb bytes
    if (S.sa.size() < (size_t)nb + 1) S.sa.resize((size_t)nb + 1);
    int* SA = S.sa.data();
    TICK(2);
    if (nb == 1) {
      SA[1] = 0;
    } else {
      // symbols shifted by one so that the sentinel 0 is unique and smallest: done on a 16-bit copy
      static thread_local std::vector<unsigned short> sym;
      if (sym.size() < (size_t)nb + 1) sym.resize((size_t)nb + 1);
      for (int q = 0; q < nb; ++q) sym[q] = (unsigned short)(rot[q] + 1);
      sym[nb] = 0;
      sais<unsigned short>(sym.data(), SA, nb + 1, 257);
    }
    TICK(3);
    // SA[0] is the sentinel suffix; ptr[i] = start of the i-th smallest rotation in the original block
    int origPtr = -1;
    for (int q = 1; q <= nb; ++q) {
      int p = SA[q] + r;
      if (p >= nb) p -= nb;
      SA[q] = p;
      if (p == 0) origPtr = q - 1;
    }
    const int* ptr = SA + 1;
    // ---- block header
    bw.put_u8(0x31); bw.put_u8(0x41); bw.put_u8(0x59); bw.put_u8(0x26); bw.put_u8(0x53); bw.put_u8(0x59);
    bw.put_u32(crc);
    bw.put(1, 0);
    bw.put(24, (unsigned)origPtr);
    // ---- move-to-front + zero-run coding (compress.c generateMTFValues)
    unsigned char unseqToSeq[256];
    int nInUse = 0;
    for (int q = 0; q < 256; ++q) if (inUse[q]) unseqToSeq[q] = (unsigned char)nInUse++;
    const int EOB = nInUse + 1;
    int mtfFreq[kMaxAlpha];
    for (int q = 0; q <= EOB; ++q) mtfFreq[q] = 0;
    if (S.mtfv.size() < (size_t)nb + 2) S.mtfv.resize((size_t)nb + 2);
    unsigned short* mtfv = S.mtfv.data();
    alignas(16) unsigned char yy[256 + 16];                   // (padded: the search below reads 16 entries at a time)
    for (int q = 0; q < 256 + 16; ++q) yy[q] = (unsigned char)q;
    int wr = 0, zPend = 0;
    auto flush_zeros = [&]() {
      if (zPend > 0) {
        --zPend;
        for (;;) {
          if (zPend & 1) { mtfv[wr++] = kRunB; ++mtfFreq[kRunB]; } else { mtfv[wr++] = kRunA; ++mtfFreq[kRunA]; }
          if (zPend < 2) break;
          zPend = (zPend - 2) / 2;
        }
        zPend = 0;
      }
    };
    for (int q = 0; q < nb; ++q) {
      int j = ptr[q] - 1;
      if (j < 0) j += nb;
      const unsigned char ll = unseqToSeq[block[j]];
      if (yy[0] == ll) { ++zPend; continue; }
      flush_zeros();
#if defined(__SSE2__)
      // position of ll in the list, 16 entries per compare; the list is a permutation of the symbols in use, so ll is
      // found before the padding.  Then the entries in front of it move up by one: inside one register when pos < 16.
      int pos;
      {
        const __m128i key = _mm_set1_epi8((char)ll);
        int base = 0;
        for (;;) {
          const int m = _mm_movemask_epi8(_mm_cmpeq_epi8(_mm_load_si128(reinterpret_cast<const __m128i*>(yy + base)), key));
          if (m) { pos = base + __builtin_ctz((unsigned)m); break; }
          base += 16;
        }
      }
      if (pos < 16) {
        const __m128i cur = _mm_load_si128(reinterpret_cast<const __m128i*>(yy));
        const __m128i sh = _mm_or_si128(_mm_slli_si128(cur, 1), _mm_cvtsi32_si128((int)ll));
        const __m128i m = _mm_loadu_si128(reinterpret_cast<const __m128i*>(kMtfMask + 15 - pos));   // 0xFF for index <= pos
        _mm_store_si128(reinterpret_cast<__m128i*>(yy), _mm_or_si128(_mm_and_si128(sh, m), _mm_andnot_si128(m, cur)));
      } else {
        memmove(yy + 1, yy, (size_t)pos);
        yy[0] = ll;
      }
#else
      int pos = 1;
      unsigned char carry = yy[0];
      while (yy[pos] != ll) { const unsigned char t2 = yy[pos]; yy[pos] = carry; carry = t2; ++pos; }
      yy[pos] = carry;
      yy[0] = ll;
#endif
      mtfv[wr++] = (unsigned short)(pos + 1);
      ++mtfFreq[pos + 1];
    }
    flush_zeros();
    mtfv[wr++] = (unsigned short)EOB;
    ++mtfFreq[EOB];
    const int nMTF = wr;
    TICK(4);
    // ---- Huffman tables (compress.c sendMTFValues)
    const int alphaSize = nInUse + 2;
    static thread_local unsigned char len[kGroups][kMaxAlpha];
    static thread_local int code[kGroups][kMaxAlpha];
    static thread_local int rfreq[kGroups][kMaxAlpha];
    static thread_local int rfreq2[kGroups][kMaxAlpha];      // (all zero between uses)
    static thread_local unsigned char selector[kMaxSelectors], selectorMtf[kMaxSelectors];
    for (int t = 0; t < kGroups; ++t) for (int v = 0; v < alphaSize; ++v) len[t][v] = 15;
    const int nGroups = nMTF < 200 ? 2 : nMTF < 600 ? 3 : nMTF < 1200 ? 4 : nMTF < 2400 ? 5 : 6;
    {
      int nPart = nGroups, remF = nMTF, gs = 0;
      while (nPart > 0) {
        const int tFreq = remF / nPart;
        int ge = gs - 1, aFreq = 0;
        while (aFreq < tFreq && ge < alphaSize - 1) { ++ge; aFreq += mtfFreq[ge]; }
        if (ge > gs && nPart != nGroups && nPart != 1 && ((nGroups - nPart) % 2 == 1)) { aFreq -= mtfFreq[ge]; --ge; }
        for (int v = 0; v < alphaSize; ++v) len[nPart - 1][v] = (v >= gs && v <= ge) ? 0 : 15;
        --nPart;
        gs = ge + 1;
        remF -= aFreq;
      }
    }
    int nSelectors = 0;
    for (int iter = 0; iter < kIters; ++iter) {
      for (int t = 0; t < nGroups; ++t) for (int v = 0; v < alphaSize; ++v) rfreq[t][v] = 0;
      // the costs of the (up to six) tables for one symbol, ten bits each, in one 64-bit word: a group of 50 symbols
      // costs at most 50 * 17 < 1024 per table, so the six sums never touch each other
      static thread_local unsigned long long pack[kMaxAlpha];
      for (int v = 0; v < alphaSize; ++v) {
        unsigned long long w = 0;
        for (int t = 0; t < nGroups; ++t) w |= (unsigned long long)len[t][v] << (10 * t);
        pack[v] = w;
      }
      nSelectors = 0;
      int gs = 0;
      while (gs < nMTF) {
        int ge = gs + kGroupSize - 1;
        if (ge >= nMTF) ge = nMTF - 1;
        unsigned long long acc = 0;
        for (int q = gs; q <= ge; ++q) acc += pack[mtfv[q]];
        int bt = -1;
        unsigned bc = 999999999u;
        for (int t = 0; t < nGroups; ++t) {
          const unsigned c = (unsigned)((acc >> (10 * t)) & 1023u);
          if (c < bc) { bc = c; bt = t; }
        }
        selector[nSelectors++] = (unsigned char)bt;
        // two sets of counters, alternate symbols: consecutive symbols are often the same one (RUNA, 1, 2), and a single
        // counter then serialises on its own store-to-load latency
        int* rf = rfreq[bt];
        int* rg2 = rfreq2[bt];
        int q = gs;
        for (; q + 1 <= ge; q += 2) { ++rf[mtfv[q]]; ++rg2[mtfv[q + 1]]; }
        if (q <= ge) ++rf[mtfv[q]];
        gs = ge + 1;
      }
      for (int t = 0; t < nGroups; ++t) {
        for (int v = 0; v < alphaSize; ++v) { rfreq[t][v] += rfreq2[t][v]; rfreq2[t][v] = 0; }
        make_code_lengths(len[t], rfreq[t], alphaSize, 17);
      }
    }
    {
      unsigned char pos[kGroups];
      for (int q = 0; q < nGroups; ++q) pos[q] = (unsigned char)q;
      for (int q = 0; q < nSelectors; ++q) {
        const unsigned char ll = selector[q];
        int j = 0;
        unsigned char tmp = pos[j];
        while (ll != tmp) { ++j; const unsigned char tmp2 = tmp; tmp = pos[j]; pos[j] = tmp2; }
        pos[0] = tmp;
        selectorMtf[q] = (unsigned char)j;
      }
    }
    for (int t = 0; t < nGroups; ++t) {
      int minLen = 32, maxLen = 0;
      for (int v = 0; v < alphaSize; ++v) { if (len[t][v] > maxLen) maxLen = len[t][v]; if (len[t][v] < minLen) minLen = len[t][v]; }
      int vec = 0;
      for (int l = minLen; l <= maxLen; ++l) {
        for (int v = 0; v < alphaSize; ++v) if (len[t][v] == l) { code[t][v] = vec; ++vec; }
        vec <<= 1;
      }
    }
    TICK(5);
    // ---- mapping table, selectors, coding tables, data
    {
      bool inUse16[16];
      for (int a = 0; a < 16; ++a) { inUse16[a] = false; for (int b = 0; b < 16; ++b) if (inUse[a * 16 + b]) inUse16[a] = true; }
      for (int a = 0; a < 16; ++a) bw.put(1, inUse16[a] ? 1 : 0);
      for (int a = 0; a < 16; ++a) if (inUse16[a]) for (int b = 0; b < 16; ++b) bw.put(1, inUse[a * 16 + b] ? 1 : 0);
    }
    bw.put(3, (unsigned)nGroups);
    bw.put(15, (unsigned)nSelectors);
    for (int q = 0; q < nSelectors; ++q) { for (int j = 0; j < selectorMtf[q]; ++j) bw.put(1, 1); bw.put(1, 0); }
    for (int t = 0; t < nGroups; ++t) {
      int curr = len[t][0];
      bw.put(5, (unsigned)curr);
      for (int v = 0; v < alphaSize; ++v) {
        while (curr < len[t][v]) { bw.put(2, 2); ++curr; }
        while (curr > len[t][v]) { bw.put(2, 3); --curr; }
        bw.put(1, 0);
      }
    }
    {
      int selCtr = 0, gs = 0;
      while (gs < nMTF) {
        int ge = gs + kGroupSize - 1;
        if (ge >= nMTF) ge = nMTF - 1;
        const unsigned char* L = len[selector[selCtr]];
        const int* Cd = code[selector[selCtr]];
        for (int q = gs; q <= ge; ++q) bw.put(L[mtfv[q]], (unsigned)Cd[mtfv[q]]);
        gs = ge + 1;
        ++selCtr;
      }
    }
  }
  TICK(6);
  bw.put_u8(0x17); bw.put_u8(0x72); bw.put_u8(0x45); bw.put_u8(0x38); bw.put_u8(0x50); bw.put_u8(0x90);
  bw.put_u32(combined);
  bw.finish();
  if (bw.overflow) { rpcc::set_error("rpcc_bz2_compress: output buffer too small"); return RPCC_ERR_CAPACITY; }
  *out_len = bw.pos;
  return RPCC_OK;
}
