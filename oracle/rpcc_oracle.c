/*
 * rpcc_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain-C, single-threaded CPU restatement of the arithmetic of R-PCC's
 * per-frame compression hot path.  It exists only so that tests/, the smoke
 * check and bench.py's cpu_baseline leg can check (and time) the CUDA path
 * against an independent implementation.  Nothing under r-pcc_b200/ may
 * import, link or execute it.
 *
 * Every function cites the reference file:line (relative to the reference
 * repo root) whose behaviour it restates.  The restatement is pinned against
 * the reference's own compiled code (oracle/_ref, built by oracle/Makefile
 * from the sources where they lie) by tests/test_oracle_pins.py and against
 * the committed fixtures in tests/golden/.
 *
 * PARITY UNPINNED for two functions at the end of this file, orc_ground_fit and
 * orc_plane_models: the reference delegates both fits to open3d's randomised
 * segment_plane (third-party, absent, fed an unseeded subsample), so there is
 * nothing of the reference to restate bit for bit.  They restate the PRODUCT's
 * own deterministic RANSACs (keeping what the reference fixes: candidate rules,
 * sample sizes, iteration counts, thresholds, the angle validation) and serve
 * as a regression pin for those two kernels only.
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math -shared -fPIC (see Makefile).
 * -ffp-contract=off matters: the reference's C++ is built for baseline x86-64
 * (no FMA), so every float product and sum is rounded separately; the places
 * where the reference's CUDA kernels DO fuse (FPS, chamfer) call fmaf()
 * explicitly below.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------- */
/* glibc 2.39 atan2f / atanf restated (fdlibm float sequence).                */
/* The reference calls libm's atan2f (cpp_modules.cpp:447,450); the CUDA      */
/* projection kernel carries a device port of exactly this sequence, and      */
/* tests compare this restatement with the box's own libm bit for bit.        */
/* ------------------------------------------------------------------------- */
static inline float orc_bits2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline int32_t orc_f2bits(float f) { int32_t i; memcpy(&i, &f, 4); return i; }

static const uint32_t kAtanHi[4] = {0x3eed6338u, 0x3f490fdau, 0x3f7b985eu, 0x3fc90fdau};
static const uint32_t kAtanLo[4] = {0x31ac3769u, 0x33222168u, 0x33140fb4u, 0x33a22168u};
static const uint32_t kAT[11] = {0x3eaaaaabu, 0xbe4ccccdu, 0x3e124925u, 0xbde38e38u,
                                 0x3dba2e6eu, 0xbd9d8795u, 0x3d886b35u, 0xbd6ef16bu,
                                 0x3d4bda59u, 0xbd15a221u, 0x3c8569d7u};

ORC_API float orc_atanf(float x) {
  int32_t hx = orc_f2bits(x);
  int32_t ix = hx & 0x7fffffff;
  int id;
  if (ix >= 0x50800000) { /* |x| >= 2^34 */
    if (ix > 0x7f800000) return x + x;
    float r = orc_bits2f(kAtanHi[3]) + orc_bits2f(kAtanLo[3]);
    return hx > 0 ? r : -r;
  }
  if (ix < 0x3ee00000) {     /* |x| < 0.4375 */
    if (ix < 0x31000000) return x; /* |x| < 2^-29 */
    id = -1;
  } else {
    x = fabsf(x);
    if (ix < 0x3f980000) {   /* |x| < 1.1875 */
      if (ix < 0x3f300000) { id = 0; x = (2.0f * x - 1.0f) / (2.0f + x); }
      else                 { id = 1; x = (x - 1.0f) / (x + 1.0f); }
    } else {
      if (ix < 0x401c0000) { id = 2; x = (x - 1.5f) / (1.0f + 1.5f * x); }
      else                 { id = 3; x = -1.0f / x; }
    }
  }
  float z = x * x;
  float w = z * z;
  const float a0 = orc_bits2f(kAT[0]), a1 = orc_bits2f(kAT[1]), a2 = orc_bits2f(kAT[2]),
              a3 = orc_bits2f(kAT[3]), a4 = orc_bits2f(kAT[4]), a5 = orc_bits2f(kAT[5]),
              a6 = orc_bits2f(kAT[6]), a7 = orc_bits2f(kAT[7]), a8 = orc_bits2f(kAT[8]),
              a9 = orc_bits2f(kAT[9]), a10 = orc_bits2f(kAT[10]);
  float s1 = z * (a0 + w * (a2 + w * (a4 + w * (a6 + w * (a8 + w * a10)))));
  float s2 = w * (a1 + w * (a3 + w * (a5 + w * (a7 + w * a9))));
  if (id < 0) return x - x * (s1 + s2);
  z = orc_bits2f(kAtanHi[id]) - ((x * (s1 + s2) - orc_bits2f(kAtanLo[id])) - x);
  return hx < 0 ? -z : z;
}

ORC_API float orc_atan2f(float y, float x) {
  const float tiny = 1.0e-30f;
  const float pi_o_2 = orc_bits2f(0x3fc90fdbu), pi = orc_bits2f(0x40490fdbu);
  const float pi_lo = orc_bits2f(0xb3bbbd2eu);
  int32_t hx = orc_f2bits(x), hy = orc_f2bits(y);
  int32_t ix = hx & 0x7fffffff, iy = hy & 0x7fffffff;
  if (ix > 0x7f800000 || iy > 0x7f800000) return x + y;
  if (hx == 0x3f800000) return orc_atanf(y);
  int m = ((hy >> 31) & 1) | ((hx >> 30) & 2);
  if (iy == 0) {
    switch (m) {
      case 0: case 1: return y;
      case 2: return pi + tiny;
      default: return -pi - tiny;
    }
  }
  if (ix == 0) return (hy < 0) ? -pi_o_2 - tiny : pi_o_2 + tiny;
  if (ix == 0x7f800000) {
    const float pi_o_4 = orc_bits2f(0x3f490fdbu);
    if (iy == 0x7f800000) {
      switch (m) {
        case 0: return pi_o_4 + tiny;
        case 1: return -pi_o_4 - tiny;
        case 2: return 3.0f * pi_o_4 + tiny;
        default: return -3.0f * pi_o_4 - tiny;
      }
    } else {
      switch (m) {
        case 0: return 0.0f;
        case 1: return -0.0f;
        case 2: return pi + tiny;
        default: return -pi - tiny;
      }
    }
  }
  if (iy == 0x7f800000) return (hy < 0) ? -pi_o_2 - tiny : pi_o_2 + tiny;
  int k = (iy - ix) >> 23;
  float z;
  if (k > 24) z = pi_o_2 + 0.5f * pi_lo;
  else if (hx < 0 && k < -26) z = 0.0f;
  else z = orc_atanf(fabsf(y / x));
  switch (m) {
    case 0: return z;
    case 1: return -z;
    case 2: return pi - (z - pi_lo);
    default: return (z - pi_lo) - pi;
  }
}

/* libm's own atan2f, exported so tests can compare on the box they run on. */
ORC_API float orc_libm_atan2f(float y, float x) { return atan2f(y, x); }

ORC_API void orc_atan2f_array(const float* y, const float* x, int64_t n, float* restated, float* libm) {
  for (int64_t i = 0; i < n; ++i) {
    restated[i] = orc_atan2f(y[i], x[i]);
    libm[i] = atan2f(y[i], x[i]);
  }
}

/* ------------------------------------------------------------------------- */
/* a1: transform map (dataset/transformer.py:41-54).  f64 trig, cast to f32.  */
/* ------------------------------------------------------------------------- */
ORC_API void orc_transform_map(int H, int W, double hfov, double vmax, double vmin, float* lut) {
  double vfov = vmax - vmin;
  for (int h = 0; h < H; ++h) {
    double alt = vfov * ((double)h / (double)(H - 1)) + vmin;
    for (int w = 0; w < W; ++w) {
      double az = hfov * ((double)w / (double)W);
      float* o = lut + ((size_t)h * W + w) * 3;
      o[0] = (float)(cos(alt) * cos(az));
      o[1] = (float)(cos(alt) * sin(az));
      o[2] = (float)sin(alt);
    }
  }
}

/* ------------------------------------------------------------------------- */
/* a2: spherical projection (cpp_modules.cpp:427-467).                        */
/* pts: n points, `stride` floats apart (3 for xyz, 4 for KITTI .bin rows).   */
/* use_restated_atan2 != 0 swaps libm's atan2f for orc_atan2f (to show the    */
/* device port's arithmetic yields the same image).                           */
/* ------------------------------------------------------------------------- */
ORC_API void orc_project(const float* pts, int stride, int64_t n, int H, int W, float hfov,
                         float vmax, float vmin, int use_restated_atan2, float* ri) {
  memset(ri, 0, sizeof(float) * (size_t)H * W);
  for (int64_t i = 0; i < n; ++i) {
    float x = pts[i * stride], y = pts[i * stride + 1], z = pts[i * stride + 2];
    float depth = sqrtf(x * x + y * y + z * z);
    float ha = use_restated_atan2 ? orc_atan2f(y, x) : atan2f(y, x);
    if (ha < 0) ha = (float)((double)ha + 2 * 3.14159265);
    float planar = sqrtf(x * x + y * y);
    float va = use_restated_atan2 ? orc_atan2f(z, planar) : atan2f(z, planar);
    int col = (int)roundf(ha / hfov * (float)W);
    col = col % W;
    float vres = (vmax - vmin) / (float)(H - 1);
    int row = (int)roundf((va - vmin) / vres);
    if (row >= H) row = H - 1;
    if (row < 0) row = 0;
    float* px = ri + (size_t)row * W + col;
    if (*px == 0 || depth < *px) *px = depth;
  }
}

/* a3: range image -> xyz (dataset/transformer.py:94-101): one f32 multiply. */
ORC_API void orc_range_to_xyz(const float* ri, const float* lut, int64_t hw, float* xyz) {
  for (int64_t i = 0; i < hw; ++i) {
    xyz[i * 3 + 0] = ri[i] * lut[i * 3 + 0];
    xyz[i * 3 + 1] = ri[i] * lut[i * 3 + 1];
    xyz[i * 3 + 2] = ri[i] * lut[i * 3 + 2];
  }
}

/* ------------------------------------------------------------------------- */
/* a5: furthest point sampling, restating the CUDA kernel                     */
/* (ops/fps/src/sampling_gpu.cu:24-140) thread by thread: per-thread strided  */
/* scan with strict '>', then the shared-memory tree whose merge keeps the    */
/* lower slot on ties (sampling_gpu.cu:16-21).  The distance uses the FMA     */
/* contraction nvcc emits for sampling_gpu.cu:64 (fma_mode 0:                 */
/* fma(dz,dz,fma(dx,dx,dy*dy)); 1: fma(dz,dz,fma(dy,dy,dx*dx)); 2: no FMA).   */
/* temp is (n) scratch; launcher picks block = min(1024, 2^floor(log2 n))     */
/* (sampling_gpu.cu:9-13).                                                    */
/* ------------------------------------------------------------------------- */
static inline float orc_fps_d(float dx, float dy, float dz, int fma_mode) {
  if (fma_mode == 0) return fmaf(dz, dz, fmaf(dx, dx, dy * dy));
  if (fma_mode == 1) return fmaf(dz, dz, fmaf(dy, dy, dx * dx));
  return dx * dx + dy * dy + dz * dz;
}

ORC_API void orc_fps(const float* pts, int n, int m, int fma_mode, float* temp, int32_t* idx) {
  if (m <= 0) return;
  int bs = 1;
  while (bs * 2 <= n && bs < 1024) bs *= 2;
  float* best = (float*)malloc(sizeof(float) * bs);
  int32_t* besti = (int32_t*)malloc(sizeof(int32_t) * bs);
  for (int k = 0; k < n; ++k) temp[k] = 1e10f;
  int old = 0;
  idx[0] = 0;
  for (int j = 1; j < m; ++j) {
    float x1 = pts[old * 3], y1 = pts[old * 3 + 1], z1 = pts[old * 3 + 2];
    for (int t = 0; t < bs; ++t) { best[t] = -1.0f; besti[t] = 0; }
    for (int k = 0; k < n; ++k) {
      int t = k & (bs - 1);
      float dx = pts[k * 3] - x1, dy = pts[k * 3 + 1] - y1, dz = pts[k * 3 + 2] - z1;
      float d = orc_fps_d(dx, dy, dz, fma_mode);
      float d2 = fminf(d, temp[k]);
      temp[k] = d2;
      if (d2 > best[t]) { best[t] = d2; besti[t] = k; }
    }
    for (int half = bs / 2; half >= 1; half /= 2) {
      for (int t = 0; t < half; ++t) {
        float v1 = best[t], v2 = best[t + half];
        if (v2 > v1) { best[t] = v2; besti[t] = besti[t + half]; }
      }
    }
    old = besti[0];
    idx[j] = old;
  }
  free(best);
  free(besti);
}

/* ------------------------------------------------------------------------- */
/* a4: segment(), GPU branch (utils/segment_utils.py:133-148,168-169),        */
/* restating the torch float32 eager ops.  assoc selects the association of   */
/* torch's size-3 reductions: 0 = (t0+t2)+t1, 1 = (t0+t1)+t2.                 */
/* xyz is the (H,W,3) point cloud the caller computed as range*LUT.           */
/* ------------------------------------------------------------------------- */
static inline float orc_sum3(float a, float b, float c, int assoc) {
  return assoc == 0 ? (a + c) + b : (a + b) + c;
}

ORC_API void orc_nonground_points(const float* xyz, const float* g, int64_t hw, float thr,
                                  int assoc, float* out) {
  float gn = sqrtf(orc_sum3(g[0] * g[0], g[1] * g[1], g[2] * g[2], assoc));
  for (int64_t i = 0; i < hw; ++i) {
    const float* p = xyz + i * 3;
    float s = orc_sum3(p[0] * g[0], p[1] * g[1], p[2] * g[2], assoc);
    float dif = fabsf(s + g[3]) / gn;
    float mk = dif > thr ? 1.0f : 0.0f;
    out[i * 3] = p[0] * mk; out[i * 3 + 1] = p[1] * mk; out[i * 3 + 2] = p[2] * mk;
  }
}

ORC_API void orc_assign_labels(const float* ri, const float* xyz, const float* lut, const float* g,
                               const float* centers, int ncenter, int64_t hw, int assoc,
                               int32_t* seg) {
  for (int64_t i = 0; i < hw; ++i) {
    const float* p = xyz + i * 3;
    const float* t = lut + i * 3;
    float den = orc_sum3(g[0] * t[0], g[1] * t[1], g[2] * t[2], assoc);
    float rplane = (-g[3]) / den;
    float bestv = -fabsf(ri[i] - rplane);
    int besti = 0;
    for (int c = 0; c < ncenter; ++c) {
      float dx = p[0] - centers[c * 3], dy = p[1] - centers[c * 3 + 1], dz = p[2] - centers[c * 3 + 2];
      float v = -fabsf(sqrtf(orc_sum3(dx * dx, dy * dy, dz * dz, assoc)));
      if (v > bestv) { bestv = v; besti = c + 1; }
    }
    int lab = besti > 0 ? besti + 1 : 0;
    if (ri[i] == 0) lab = 1;
    seg[i] = lab;
  }
}

/* ------------------------------------------------------------------------- */
/* a6: point_modeling (cpp_modules.cpp:471-518).  out has max_label+1 floats; */
/* returns that count.  Empty cluster -> 0.0/0 = NaN, as the reference.       */
/* ------------------------------------------------------------------------- */
ORC_API int orc_point_modeling(const float* ri, const int32_t* seg, int64_t hw, float* out, int cap) {
  int kmax = 0;
  for (int64_t i = 0; i < hw; ++i) if (seg[i] > kmax) kmax = seg[i];
  int K = kmax + 1;
  if (K > cap) return -K;
  double* sum = (double*)calloc(K, sizeof(double));
  int64_t* cnt = (int64_t*)calloc(K, sizeof(int64_t));
  for (int64_t i = 0; i < hw; ++i) {
    int l = seg[i];
    if (l != 0 && l != 1) { sum[l] += (double)ri[i]; cnt[l]++; }
  }
  for (int l = 0; l < K; ++l) {
    if (l < 2) out[l] = 0.0f;
    else out[l] = (float)(sum[l] / (double)(size_t)cnt[l]);
  }
  free(sum); free(cnt);
  return K;
}

/* a7: intra_predict (cpp_modules.cpp:248-285). */
ORC_API void orc_intra_predict(const int32_t* seg, const float* model, const float* lut, int64_t hw,
                               float* pred) {
  for (int64_t i = 0; i < hw; ++i) {
    const float* m = model + (size_t)seg[i] * 4;
    if (m[0] + m[1] + m[2] == 0) pred[i] = m[3];
    else pred[i] = -m[3] / (m[0] * lut[i * 3] + m[1] * lut[i * 3 + 1] + m[2] * lut[i * 3 + 2]);
  }
}

/* a8: uniform_quantize (cpp_modules.cpp:288-334).  Returns symbol count. */
ORC_API int64_t orc_uniform_quantize(const int32_t* seg, const float* res, int64_t hw, float step,
                                     int32_t* out) {
  int kmax = 0;
  for (int64_t i = 0; i < hw; ++i) if (seg[i] > kmax) kmax = seg[i];
  int K = kmax + 1;
  int64_t* off = (int64_t*)calloc(K + 1, sizeof(int64_t));
  for (int64_t i = 0; i < hw; ++i) if (seg[i] != 1) off[seg[i] + 1]++;
  for (int l = 0; l < K; ++l) off[l + 1] += off[l];
  int64_t total = off[K];
  for (int64_t i = 0; i < hw; ++i) {
    int l = seg[i];
    if (l == 1) continue;
    out[off[l]++] = (int32_t)roundf(res[i] / step);
  }
  free(off);
  return total;
}

/* ------------------------------------------------------------------------- */
/* a9: extract_features_with_segment + mark_as_picked                         */
/* (cpp_modules.cpp:28-121, 10-25).  kp (H*W int32) and feat (H*W f32) are    */
/* zero-filled here (the reference relies on fresh pages).         */
/* ------------------------------------------------------------------------- */
typedef struct { float c; int s; } orc_feat;

static int orc_feat_cmp(const void* a, const void* b) {
  const orc_feat* p = (const orc_feat*)a; const orc_feat* q = (const orc_feat*)b;
  if (p->c < q->c) return -1;
  if (q->c < p->c) return 1;
  return (p->s > q->s) - (p->s < q->s);
}

static int orc_accept(const float* ri, uint8_t* visited, int W, int h, int w, int region) {
  float r = ri[(size_t)h * W + w];
  int ok = 1;
  for (int i = -region; i <= region; ++i) {
    float dif = r - ri[(size_t)h * W + w + i];
    if (fabsf(dif) < 0.2f) visited[(size_t)h * W + w] = 1;
    if (dif > 0.3f) ok = 0;
  }
  return ok;
}

ORC_API void orc_extract_features(const float* ri, const int32_t* seg, int H, int W, int region,
                                  int segments, int sharp_num, int less_sharp_num, int flat_num,
                                  float* feat, int32_t* kp) {
  memset(feat, 0, sizeof(float) * (size_t)H * W);
  memset(kp, 0, sizeof(int32_t) * (size_t)H * W);
  uint8_t* visited = (uint8_t*)calloc((size_t)H * W, 1);
  float* vr = (float*)malloc(sizeof(float) * W);
  int* cols = (int*)malloc(sizeof(int) * W);
  orc_feat* F = (orc_feat*)malloc(sizeof(orc_feat) * W);
  for (int h = 0; h < H; ++h) {
    int L = 0;
    for (int w = 0; w < W; ++w) {
      int l = seg[(size_t)h * W + w];
      if (l != 0 && l != 1) { vr[L] = ri[(size_t)h * W + w]; cols[L] = w; ++L; }
    }
    if (L < segments + region * 2 + 1) continue;
    int nf = 0;
    for (int s = region; s < L - region; ++s) {
      float a = 0.0f;
      for (int k = -region; k <= region; ++k) a += vr[s + k] - vr[s];
      a = a * a;
      a /= (float)(2 * region);
      a /= vr[s];
      feat[(size_t)h * W + cols[s]] = a;
      F[nf].c = a; F[nf].s = s; ++nf;
    }
    int per = nf / segments;
    for (int j = 0; j < segments; ++j) {
      orc_feat* S = F + per * j;
      qsort(S, per, sizeof(orc_feat), orc_feat_cmp);
      int picked = 0;
      for (int i = per - 1; i >= 0; --i) {
        int w = cols[S[i].s];
        S[i].c = 0;
        if (!visited[(size_t)h * W + w] && orc_accept(ri, visited, W, h, w, region)) {
          ++picked;
          if (picked < sharp_num) kp[(size_t)h * W + w] = 3;
          else if (picked < less_sharp_num) kp[(size_t)h * W + w] = 2;
          else break;
        }
      }
      qsort(S, per, sizeof(orc_feat), orc_feat_cmp);
      picked = 0;
      for (int i = 0; i < per; ++i) {
        if (S[i].c == 0) continue;
        int w = cols[S[i].s];
        S[i].c = 0;
        if (!visited[(size_t)h * W + w] && orc_accept(ri, visited, W, h, w, region)) {
          ++picked;
          if (picked < flat_num) kp[(size_t)h * W + w] = 1;
          else break;
        }
      }
    }
  }
  free(visited); free(vr); free(cols); free(F);
}

/* a9: nonuniform_quantize (cpp_modules.cpp:337-424).  salience has K entries
 * (returned through *K_out); returns symbol count. */
ORC_API int64_t orc_nonuniform_quantize(const int32_t* seg, const float* res, const int32_t* kp,
                                        int64_t hw, const int32_t* level_kp_num,
                                        const float* level_acc, int level_num, int ground_level,
                                        int32_t* out, int32_t* salience, int* K_out) {
  int kmax = 0;
  for (int64_t i = 0; i < hw; ++i) if (seg[i] > kmax) kmax = seg[i];
  int K = kmax + 1;
  *K_out = K;
  int64_t* off = (int64_t*)calloc(K + 1, sizeof(int64_t));
  int* kpn = (int*)calloc(K, sizeof(int));
  int* pn = (int*)calloc(K, sizeof(int));
  float* acc = (float*)calloc(K, sizeof(float));
  for (int64_t i = 0; i < hw; ++i) {
    int l = seg[i];
    if (l == 1) continue;
    if (kp[i] > 0) kpn[l]++;
    pn[l]++;
    off[l + 1]++;
  }
  for (int l = 0; l < K; ++l) {
    int lev = 0;
    if (l == 0) lev = ground_level;
    else if (l == 1) lev = level_num - 1;
    else if (pn[l] < 30) lev = level_num - 1;
    else {
      for (int q = 0; q < level_num; ++q) if (kpn[l] >= level_kp_num[q]) { lev = q; break; }
    }
    salience[l] = lev;
    acc[l] = level_acc[lev];
  }
  for (int l = 0; l < K; ++l) off[l + 1] += off[l];
  int64_t total = off[K];
  for (int64_t i = 0; i < hw; ++i) {
    int l = seg[i];
    if (l == 1) continue;
    out[off[l]++] = (int32_t)roundf(res[i] / acc[l]);
  }
  free(off); free(kpn); free(pn); free(acc);
  return total;
}

/* a10: extract_contour (cpp_modules.cpp:521-558).  Returns sequence length. */
ORC_API int64_t orc_extract_contour(const int32_t* seg, int H, int W, int32_t* contour, int32_t* seq) {
  int64_t L = 0;
  for (int h = 0; h < H; ++h) {
    const int32_t* row = seg + (size_t)h * W;
    int32_t* c = contour + (size_t)h * W;
    c[0] = 1; seq[L++] = row[0];
    for (int w = 1; w < W; ++w) {
      if (row[w] != row[w - 1]) { c[w] = 1; seq[L++] = row[w]; }
      else c[w] = 0;
    }
  }
  return L;
}

/* a11: recover_map (cpp_modules.cpp:561-593). */
ORC_API void orc_recover_map(const int32_t* contour, const int32_t* seq, int64_t L, int64_t hw,
                             int32_t* seg) {
  int64_t p = 0;
  for (int64_t i = 0; i < L && p < hw; ++i) {
    int32_t v = seq[i];
    seg[p++] = v;
    while (p < hw && contour[p] == 0) seg[p++] = v;
  }
}

/* a11: dequantize_residual (utils/compress_utils.py:114-132) under numpy 2.x
 * promotion: f32( (double)q * step_f64 ), scattered back label-major.
 * steps: per-label f64 step (uniform callers pass the same value K times).
 * Returns symbols consumed. */
ORC_API int64_t orc_dequantize(const int16_t* q, const int32_t* seg, int64_t hw, const double* steps,
                               int K, float* res) {
  int64_t* off = (int64_t*)calloc(K + 1, sizeof(int64_t));
  for (int64_t i = 0; i < hw; ++i) { res[i] = 0.0f; if (seg[i] != 1 && seg[i] < K) off[seg[i] + 1]++; }
  for (int l = 0; l < K; ++l) off[l + 1] += off[l];
  int64_t total = off[K];
  for (int64_t i = 0; i < hw; ++i) {
    int l = seg[i];
    if (l == 1 || l >= K) continue;
    res[i] = (float)((double)q[off[l]++] * steps[l]);
  }
  free(off);
  return total;
}

/* ------------------------------------------------------------------------- */
/* a12: chamfer NN (chamfer3D.cu:12-134): for each point of A the min squared */
/* distance to B and its first argmin.  fma_mode as compiled by nvcc for      */
/* chamfer3D.cu:32-35 (0: fma(z,z,fma(y,y,x*x)), 1: fma(z,z,fma(x,x,y*y)),    */
/* 2: none).  Scan order over B is ascending with strict '<' inside a 512     */
/* batch and strict '>' across batches (chamfer3D.cu:126), i.e. first min.    */
/* ------------------------------------------------------------------------- */
ORC_API void orc_chamfer_nn(const float* a, int64_t n, const float* b, int64_t m, int fma_mode,
                            float* dist, int32_t* idx) {
  for (int64_t i = 0; i < n; ++i) {
    float x1 = a[i * 3], y1 = a[i * 3 + 1], z1 = a[i * 3 + 2];
    float best = 0; int32_t bi = 0;
    for (int64_t k = 0; k < m; ++k) {
      float x2 = b[k * 3] - x1, y2 = b[k * 3 + 1] - y1, z2 = b[k * 3 + 2] - z1;
      float d;
      if (fma_mode == 0) d = fmaf(z2, z2, fmaf(y2, y2, x2 * x2));
      else if (fma_mode == 1) d = fmaf(z2, z2, fmaf(x2, x2, y2 * y2));
      else d = x2 * x2 + y2 * y2 + z2 * z2;
      if (k == 0 || d < best) { best = d; bi = (int32_t)k; }
    }
    dist[i] = best; idx[i] = bi;
  }
}

/* ------------------------------------------------------------------------------------------------
 * Ground plane: restatement of the PRODUCT's deterministic RANSAC (r-pcc_b200/csrc/ground.cu), not of
 * the reference -- utils/segment_utils.py:74-82,101-108 hands an unseeded random subsample to open3d's
 * segment_plane, which is third-party, absent and randomised ("parity unpinned").  What the reference
 * fixes is kept (candidates z < -1.5, at most 5000 of them, fewer than 800 -> every pixel, 10-point
 * least-squares hypotheses, 100 of them, inliers at 0.1 m, refit on the winner's inliers); this function
 * repeats the device's counter-based samples and its summation orders so that the fitted plane can be
 * compared bit for bit (a regression pin for the kernel; tests/test_gpu_stages.py).
 * `threads` = threads per CTA of ground_fit_kernel (512): it fixes the order of the refit's partial sums. */
static uint64_t orc_splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

/* open3d GetPlaneFromPoints from the ten moments (ransac.cuh: plane_from_sums) */
static int orc_plane_from_sums(const double* s, double* plane) {
  const double n = s[0];
  if (n < 3.0) return 0;
  const double cx = s[1] / n, cy = s[2] / n, cz = s[3] / n;
  const double xx = s[4] - n * cx * cx, xy = s[5] - n * cx * cy, xz = s[6] - n * cx * cz;
  const double yy = s[7] - n * cy * cy, yz = s[8] - n * cy * cz, zz = s[9] - n * cz * cz;
  const double dx = yy * zz - yz * yz, dy = xx * zz - xz * xz, dz = xx * yy - xy * xy;
  const double dmax = fmax(dx, fmax(dy, dz));
  if (!(dmax > 0.0)) return 0;
  double a, b, c;
  if (dmax == dx) { a = dx; b = xz * yz - xy * zz; c = xy * yz - xz * yy; }
  else if (dmax == dy) { a = xz * yz - xy * zz; b = dy; c = xy * xz - yz * xx; }
  else { a = xy * yz - xz * yy; b = xy * xz - yz * xx; c = dz; }
  const double nn = sqrt(a * a + b * b + c * c);
  if (!(nn > 0.0)) return 0;
  a /= nn; b /= nn; c /= nn;
  plane[0] = a; plane[1] = b; plane[2] = c; plane[3] = -(a * cx + b * cy + c * cz);
  return 1;
}

static inline float orc_plane_dist(float a, float b, float c, float d, float x, float y, float z) {
  return fabsf(fmaf(a, x, fmaf(b, y, fmaf(c, z, d))));
}

ORC_API void orc_ground_fit(const float* range, const float* lut, int64_t hw, uint64_t seed, uint64_t frame,
                            int threads, float* ground) {
  enum { kMax = 5000, kIters = 100, kSample = 10 };
  const float z_below = -1.5f, thr = 0.1f;
  /* candidates in raster order */
  int64_t nc = 0;
  for (int64_t p = 0; p < hw; ++p) nc += (range[p] * lut[3 * p + 2] < z_below);
  const int use_all = nc < 800;
  if (use_all) nc = hw;
  const int ns = (int)(nc < kMax ? nc : kMax);
  float* px = (float*)malloc(sizeof(float) * 3 * (size_t)ns);
  float *py = px + ns, *pz = py + ns;
  {
    int64_t r = -1, p = -1;                 /* candidate number of pixel p */
    for (int s = 0; s < ns; ++s) {
      const int64_t want = (ns != nc) ? (int64_t)(((uint32_t)s * (uint32_t)nc + (uint32_t)ns - 1u) / (uint32_t)ns) : s;
      while (r < want) { ++p; if (use_all || range[p] * lut[3 * p + 2] < z_below) ++r; }
      const float rr = range[p];
      px[s] = rr * lut[3 * p]; py[s] = rr * lut[3 * p + 1]; pz[s] = rr * lut[3 * p + 2];
    }
  }
  const int ns4 = (ns + 3) >> 2;
  double planes[kIters][4];
  uint64_t best = 0;
  for (int it = 0; it < kIters; ++it) {
    uint64_t st = orc_splitmix64(orc_splitmix64(seed + frame) ^ ((uint64_t)it << 40));
    double s[10] = {0};
    for (int j = 0; j < kSample; ++j) {
      st = orc_splitmix64(st);              /* lane j walks j + 1 links of the one chain */
      const int k = (int)(st % (uint64_t)ns);
      const double x = px[k], y = py[k], z = pz[k];
      s[0] += 1.0; s[1] += x; s[2] += y; s[3] += z;
      s[4] += x * x; s[5] += x * y; s[6] += x * z; s[7] += y * y; s[8] += y * z; s[9] += z * z;
    }
    double pl[4] = {0, 0, 0, 0};
    if (!orc_plane_from_sums(s, pl)) { pl[0] = pl[1] = pl[2] = pl[3] = 0.0; }
    for (int q = 0; q < 4; ++q) planes[it][q] = pl[q];
    const float a = (float)pl[0], b = (float)pl[1], c = (float)pl[2], d = (float)pl[3];
    int inl_l[32];
    float err_l[32];
    for (int L = 0; L < 32; ++L) { inl_l[L] = 0; err_l[L] = 0.f; }
    if (a != 0.f || b != 0.f || c != 0.f) {
      for (int L = 0; L < 32; ++L)
        for (int k4 = L; k4 < ns4; k4 += 32)
          for (int e = 0; e < 4; ++e) {
            const int k = 4 * k4 + e;
            if (k >= ns) continue;          /* the device pads with NaN: never an inlier */
            const float dd = orc_plane_dist(a, b, c, d, px[k], py[k], pz[k]);
            if (dd < thr) { ++inl_l[L]; err_l[L] += dd * dd; }
          }
    }
    for (int o = 16; o > 0; o >>= 1) {      /* the warp's xor tree */
      int ni[32];
      float ne[32];
      for (int L = 0; L < 32; ++L) { ni[L] = inl_l[L] + inl_l[L ^ o]; ne[L] = err_l[L] + err_l[L ^ o]; }
      memcpy(inl_l, ni, sizeof ni);
      memcpy(err_l, ne, sizeof ne);
    }
    const int inl = inl_l[0];
    const float rmse = inl > 0 ? sqrtf(err_l[0] / (float)inl) : 1e9f;
    const float scaled = rmse * 1e6f;
    const uint32_t qr = scaled >= 4294967296.0f ? 0xFFFFFFFFu : (uint32_t)scaled;    /* cvt.rzi.u32.f32 saturates */
    const uint32_t qc = qr < 0xFFFFFFu ? qr : 0xFFFFFFu;
    const uint64_t score = ((uint64_t)inl << 40) | ((uint64_t)(0xFFFFFFu - qc) << 8) | (uint64_t)(0xFFu - it);
    if (score > best) best = score;
  }
  const int bi = 0xFF - (int)(best & 0xFFull);
  const double* sb = planes[bi];
  /* refit: thread t adds its points k = t, t + threads, ... ; xor tree inside each warp; warps in index order */
  double sums[10];
  {
    const float a = (float)sb[0], b = (float)sb[1], c = (float)sb[2], d = (float)sb[3];
    const int nw = threads / 32;
    double* part = (double*)calloc((size_t)threads * 10, sizeof(double));
    for (int t = 0; t < threads; ++t)
      for (int k = t; k < ns; k += threads) {
        const double x = px[k], y = py[k], z = pz[k];
        if (orc_plane_dist(a, b, c, d, px[k], py[k], pz[k]) < thr) {
          double* s = part + (size_t)t * 10;
          s[0] += 1.0; s[1] += x; s[2] += y; s[3] += z;
          s[4] += x * x; s[5] += x * y; s[6] += x * z; s[7] += y * y; s[8] += y * z; s[9] += z * z;
        }
      }
    for (int q = 0; q < 10; ++q) {
      double total = 0.0;
      for (int w = 0; w < nw; ++w) {
        double v[32], nv[32];
        for (int L = 0; L < 32; ++L) v[L] = part[(size_t)(w * 32 + L) * 10 + q];
        for (int o = 16; o > 0; o >>= 1) {
          for (int L = 0; L < 32; ++L) nv[L] = v[L] + v[L ^ o];
          memcpy(v, nv, sizeof v);
        }
        total += v[0];
      }
      sums[q] = total;
    }
    free(part);
  }
  double pl[4];
  if (!orc_plane_from_sums(sums, pl)) {
    if (sb[0] != 0.0 || sb[1] != 0.0 || sb[2] != 0.0) { for (int q = 0; q < 4; ++q) pl[q] = sb[q]; }
    else { pl[0] = 0.0; pl[1] = 0.0; pl[2] = 1.0; pl[3] = 1.73; }
  }
  for (int q = 0; q < 4; ++q) ground[q] = (float)pl[q];
  free(px);
}

/* ------------------------------------------------------------------------------------------------
 * Per-cluster plane models: restatement of the PRODUCT's deterministic RANSAC
 * (r-pcc_b200/csrc/plane.cu: plane_model_kernel), for the same reason as orc_ground_fit -- the
 * reference's cluster_modeling(model_method='plane') (utils/segment_utils.py:188-216) calls open3d's
 * randomised segment_plane per cluster.  Kept from the reference: clusters under 30 pixels stay point
 * models, 4-point least-squares hypotheses, 10 of them, inliers at 0.1 m by count then rmse, refit on
 * the winner's inliers, plane_angle_validation (:84-93) with its precedence quirk and numpy's NaN
 * semantics.  `model` holds the point-model rows on entry ([K][4] f32); rows of accepted planes are
 * overwritten.  `threads` = threads per CTA of plane_model_kernel (128): it fixes the summation order. */
ORC_API void orc_plane_models(const float* range, const float* lut, const int32_t* seg, int64_t hw, int K,
                              uint64_t seed, uint64_t frame, int min_pixels, float dist_thr, int ransac_n,
                              int iters, float angle_threshold_deg, int threads, float* model) {
  const double thr = (double)dist_thr;
  const double cos_thr = cos(3.14159265358979323846 * ((double)angle_threshold_deg / 180.0));
  const int nw = threads / 32;
  int64_t* pix = (int64_t*)malloc(sizeof(int64_t) * (size_t)hw);
  double* part = (double*)malloc(sizeof(double) * (size_t)threads * 10);
  for (int l = 2; l < K; ++l) {
    int64_t n = 0;
    for (int64_t p = 0; p < hw; ++p) if (seg[p] == l) pix[n++] = p;      /* raster order inside the label */
    if (n < min_pixels) continue;
#define ORC_PT(k, x, y, z) do { const int64_t p_ = pix[k]; const float r_ = range[p_];                      \
      (x) = (double)(r_ * lut[3 * p_]); (y) = (double)(r_ * lut[3 * p_ + 1]); (z) = (double)(r_ * lut[3 * p_ + 2]); } while (0)
    double planes[16][4];
    for (int it = 0; it < iters; ++it) {
      uint64_t st = orc_splitmix64(orc_splitmix64(seed + frame) ^ ((uint64_t)l << 48) ^ ((uint64_t)it << 32));
      uint32_t pick[16];
      double s[10] = {0};
      for (int j = 0; j < ransac_n; ++j) {
        uint32_t k;
        int dup;
        do {
          st = orc_splitmix64(st);
          k = (uint32_t)(st % (uint64_t)n);
          dup = 0;
          for (int e = 0; e < j; ++e) dup = dup || (pick[e] == k);
        } while (dup);
        pick[j] = k;
        double x, y, z;
        ORC_PT(k, x, y, z);
        s[0] += 1.0; s[1] += x; s[2] += y; s[3] += z;
        s[4] += x * x; s[5] += x * y; s[6] += x * z; s[7] += y * y; s[8] += y * z; s[9] += z * z;
      }
      double pl[4] = {0, 0, 0, 0};
      if (!orc_plane_from_sums(s, pl)) { pl[0] = pl[1] = pl[2] = pl[3] = 0.0; }
      for (int q = 0; q < 4; ++q) planes[it][q] = pl[q];
    }
    /* scores: thread t adds k = t, t + threads, ...; xor tree inside a warp; warps in index order */
    int best = -1;
    double best_cnt = 0.0, best_rmse = 0.0;
    for (int h = 0; h < iters; ++h) {
      const double p0 = planes[h][0], p1 = planes[h][1], p2 = planes[h][2], p3 = planes[h][3];
      int ci = 0;
      double e = 0.0;
      for (int w = 0; w < nw; ++w) {
        double v[32], nv[32];
        for (int L = 0; L < 32; ++L) {
          double err = 0.0;
          for (int64_t k = w * 32 + L; k < n; k += threads) {
            double x, y, z;
            ORC_PT(k, x, y, z);
            const double d = fabs(p0 * x + p1 * y + p2 * z + p3);
            if (d < thr) { ++ci; err += d * d; }
          }
          v[L] = err;
        }
        for (int o = 16; o > 0; o >>= 1) {
          for (int L = 0; L < 32; ++L) nv[L] = v[L] + v[L ^ o];
          memcpy(v, nv, sizeof v);
        }
        e += v[0];
      }
      const double c = (double)ci;
      const int valid = p0 != 0.0 || p1 != 0.0 || p2 != 0.0;
      if (!valid || c <= 0.0) continue;
      const double rmse = sqrt(e / c);
      if (best < 0 || c > best_cnt || (c == best_cnt && rmse < best_rmse)) { best = h; best_cnt = c; best_rmse = rmse; }
    }
    if (best < 0) continue;
    const double b0 = planes[best][0], b1 = planes[best][1], b2 = planes[best][2], b3 = planes[best][3];
    memset(part, 0, sizeof(double) * (size_t)threads * 10);
    for (int t = 0; t < threads; ++t)
      for (int64_t k = t; k < n; k += threads) {
        double x, y, z;
        ORC_PT(k, x, y, z);
        if (fabs(b0 * x + b1 * y + b2 * z + b3) < thr) {
          double* s = part + (size_t)t * 10;
          s[0] += 1.0; s[1] += x; s[2] += y; s[3] += z;
          s[4] += x * x; s[5] += x * y; s[6] += x * z; s[7] += y * y; s[8] += y * z; s[9] += z * z;
        }
      }
    double sums[10];
    for (int q = 0; q < 10; ++q) {
      double total = 0.0;
      for (int w = 0; w < nw; ++w) {
        double v[32], nv[32];
        for (int L = 0; L < 32; ++L) v[L] = part[(size_t)(w * 32 + L) * 10 + q];
        for (int o = 16; o > 0; o >>= 1) {
          for (int L = 0; L < 32; ++L) nv[L] = v[L] + v[L ^ o];
          memcpy(v, nv, sizeof v);
        }
        total += v[0];
      }
      sums[q] = total;
    }
    double pl[4];
    if (!orc_plane_from_sums(sums, pl)) { pl[0] = b0; pl[1] = b1; pl[2] = b2; pl[3] = b3; }
    /* plane_angle_validation, f64 like numpy: arccos(|n.s| / |n| * |s|) */
    const double a = pl[0], b = pl[1], c = pl[2], d = pl[3];
    const double nn = sqrt(a * a + b * b + c * c);
    int flag = 0;
    for (int64_t k = 0; k < n; ++k) {
      const int64_t p = pix[k];
      const double sx = lut[3 * p], sy = lut[3 * p + 1], sz = lut[3 * p + 2];
      const double v = fabs(a * sx + b * sy + c * sz) / nn * sqrt(sx * sx + sy * sy + sz * sz);
      if (!(v <= 1.0)) flag |= 2;
      else if (v < cos_thr) flag |= 1;
    }
    if ((flag & 2) || !(flag & 1)) {
      model[4 * l] = (float)a; model[4 * l + 1] = (float)b; model[4 * l + 2] = (float)c; model[4 * l + 3] = (float)d;
    }
#undef ORC_PT
  }
  free(part);
  free(pix);
}
