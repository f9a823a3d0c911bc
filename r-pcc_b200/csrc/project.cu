// project.cu -- stage 1: spherical range-image projection, B frames per launch.
//
// Replaces dataset_utils_cpp.point_cloud_to_range_image_even
// (reference ops/cpp_modules/src/cpp_modules.cpp:427-467) and the numpy multiply of
// dataset/transformer.py:94-101.
//
// Design (HBM-bound: 16 B/point in, 4 B/pixel out):
//   one persistent CTA of 1024 threads per SM; a CTA owns a whole frame at a time, so the
//   frame's range image lives in L2 from its sentinel fill, through the z-buffer reductions
//   (RED.MIN.U32 on the order-preserving bit pattern of the non-negative depth), to the
//   in-place finalisation -- DRAM sees the points once and the image once.
//   Points are read as coalesced float4 rows (x,y,z,intensity) with L1 bypass.
// Exactness: pixel indices use a device port of glibc's atan2f (the reference calls libm;
// CUDA's own atan2f differs in the last ulp and flips ~10 pixels per frame).  All float ops
// are unfused (-fmad=false), division and sqrt are IEEE.
#include "common.cuh"

namespace rpcc {

// ---- glibc 2.39 atanf/atan2f float sequence (fdlibm), see oracle/rpcc_oracle.c and SURVEY App. F
__device__ __forceinline__ float dev_atanf(float x) {
  const int hx = __float_as_int(x);
  const int ix = hx & 0x7fffffff;
  int id;
  if (ix >= 0x50800000) {
    if (ix > 0x7f800000) return x + x;
    const float r = __int_as_float(0x3fc90fda) + __int_as_float(0x33a22168);
    return hx > 0 ? r : -r;
  }
  float hi = 0.f, lo = 0.f;
  if (ix < 0x3ee00000) {
    if (ix < 0x31000000) return x;
    id = -1;
  } else {
    x = fabsf(x);
    if (ix < 0x3f980000) {
      if (ix < 0x3f300000) { id = 0; x = (2.0f * x - 1.0f) / (2.0f + x); hi = __int_as_float(0x3eed6338); lo = __int_as_float(0x31ac3769); }
      else                 { id = 1; x = (x - 1.0f) / (x + 1.0f);        hi = __int_as_float(0x3f490fda); lo = __int_as_float(0x33222168); }
    } else {
      if (ix < 0x401c0000) { id = 2; x = (x - 1.5f) / (1.0f + 1.5f * x); hi = __int_as_float(0x3f7b985e); lo = __int_as_float(0x33140fb4); }
      else                 { id = 3; x = -1.0f / x;                       hi = __int_as_float(0x3fc90fda); lo = __int_as_float(0x33a22168); }
    }
  }
  const float z = x * x;
  const float w = z * z;
  const float s1 = z * (__int_as_float(0x3eaaaaab) + w * (__int_as_float(0x3e124925) + w * (__int_as_float(0x3dba2e6e) +
                   w * (__int_as_float(0x3d886b35) + w * (__int_as_float(0x3d4bda59) + w * __int_as_float(0x3c8569d7))))));
  const float s2 = w * (__int_as_float(0xbe4ccccd) + w * (__int_as_float(0xbde38e38) + w * (__int_as_float(0xbd9d8795) +
                   w * (__int_as_float(0xbd6ef16b) + w * __int_as_float(0xbd15a221)))));
  if (id < 0) return x - x * (s1 + s2);
  const float r = hi - ((x * (s1 + s2) - lo) - x);
  return hx < 0 ? -r : r;
}

__device__ __forceinline__ float dev_atan2f(float y, float x) {
  const float tiny = 1.0e-30f;
  const float pi_o_2 = __int_as_float(0x3fc90fdb), pi = __int_as_float(0x40490fdb);
  const float pi_lo = __int_as_float(0xb3bbbd2e);
  const int hx = __float_as_int(x), hy = __float_as_int(y);
  const int ix = hx & 0x7fffffff, iy = hy & 0x7fffffff;
  if (ix > 0x7f800000 || iy > 0x7f800000) return x + y;
  if (hx == 0x3f800000) return dev_atanf(y);
  const int m = ((hy >> 31) & 1) | ((hx >> 30) & 2);
  if (iy == 0) return m < 2 ? y : (m == 2 ? pi + tiny : -pi - tiny);
  if (ix == 0) return hy < 0 ? -pi_o_2 - tiny : pi_o_2 + tiny;
  if (ix == 0x7f800000) {
    const float pi_o_4 = __int_as_float(0x3f490fdb);
    if (iy == 0x7f800000) {
      return m == 0 ? pi_o_4 + tiny : m == 1 ? -pi_o_4 - tiny : m == 2 ? 3.0f * pi_o_4 + tiny : -3.0f * pi_o_4 - tiny;
    }
    return m == 0 ? 0.0f : m == 1 ? -0.0f : m == 2 ? pi + tiny : -pi - tiny;
  }
  if (iy == 0x7f800000) return hy < 0 ? -pi_o_2 - tiny : pi_o_2 + tiny;
  const int k = (iy - ix) >> 23;
  float z;
  if (k > 24) z = pi_o_2 + 0.5f * pi_lo;
  else if (hx < 0 && k < -26) z = 0.0f;
  else z = dev_atanf(fabsf(y / x));
  switch (m) {
    case 0: return z;
    case 1: return -z;
    case 2: return pi - (z - pi_lo);
    default: return (z - pi_lo) - pi;
  }
}

struct Pixel { int pix; float depth; };

// cpp_modules.cpp:443-458, expression for expression.
__device__ __forceinline__ Pixel point_to_pixel(float x, float y, float z, int H, int W, float hfov, float vmin, float vres) {
  Pixel p;
  p.depth = sqrtf(x * x + y * y + z * z);
  float ha = dev_atan2f(y, x);
  if (ha < 0) ha = (float)((double)ha + 2 * 3.14159265);
  const float va = dev_atan2f(z, sqrtf(x * x + y * y));
  int col = (int)roundf(ha / hfov * (float)W);
  col = col % W;
  int row = (int)roundf((va - vmin) / vres);
  row = row >= H ? H - 1 : row;
  row = row < 0 ? 0 : row;
  p.pix = row * W + col;
  return p;
}

constexpr unsigned kEmpty = 0xFFFFFFFFu;
constexpr int kProjThreads = 1024;
constexpr int kDeferCap = 2560;   // points per frame whose pixel is re-derived with the exact libm sequence

// Fast pixel derivation with a proof obligation instead of exactness.  A division-free atan2
// (degree-17 odd minimax polynomial on [0,1], |error| <= 1.0e-7 rad in f32, plus the quadrant
// fix-ups) and reciprocal multiplies give continuous column / row coordinates that differ from the
// reference's (glibc atan2f <= 1 ulp, then IEEE divisions) by at most
//   col: ~6e-4 px at W = 2000 (7e-7 rad of angle error scaled by W / hfov, a few ulp(W) of rounding)
//   row: ~6e-5 rows at vres = 0.00745 rad
// `mcol` / `mrow` (host, from the lidar table) are >= 4x those bounds.  If the fast coordinate is
// farther than the margin from every rounding boundary (k + 0.5), rounding it gives the reference's
// integer; otherwise -- under 1 % of the points -- `ok` is false and the caller re-derives the pixel with
// the exact sequence.  Zero / tiny / huge coordinates always take the exact path.
__device__ __forceinline__ float fast_atan2(float y, float x) {
  const float ax = fabsf(x), ay = fabsf(y);
  const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
  const float t = __fdividef(mn, mx);
  const float u = __fmul_rn(t, t);
  float p = 0.002456609858199954f;
  p = __fmaf_rn(p, u, -0.01440086867660284f);
  p = __fmaf_rn(p, u, 0.03978036344051361f);
  p = __fmaf_rn(p, u, -0.07234777510166168f);
  p = __fmaf_rn(p, u, 0.10498903691768646f);
  p = __fmaf_rn(p, u, -0.14161217212677002f);
  p = __fmaf_rn(p, u, 0.19985905289649963f);
  p = __fmaf_rn(p, u, -0.33332598209381104f);
  p = __fmaf_rn(p, u, 0.9999998807907104f);
  float a = __fmul_rn(p, t);
  a = ay > ax ? 1.5707963705062866f - a : a;
  a = x < 0.0f ? 3.1415927410125732f - a : a;
  return copysignf(a, y);
}

__device__ __forceinline__ int fast_pixel(float x, float y, float z, int H, int W, float col_scale, float vmin,
                                          float row_scale, float mcol, float mrow, bool& ok) {
  const float p2 = x * x + y * y;
  const float planar = __fsqrt_rn(p2);
  float ha = fast_atan2(y, x);
  ha = ha < 0.0f ? ha + 6.2831853071795862f : ha;
  const float va = fast_atan2(z, planar);
  const float ct = __fmaf_rn(ha, col_scale, 0.5f);          // column coordinate + 0.5
  const float rt = __fmaf_rn(va - vmin, row_scale, 0.5f);   // row coordinate + 0.5
  const float cfl = floorf(ct), rfl = floorf(rt);
  const float cfr = ct - cfl, rfr = rt - rfl;               // distance above the lower rounding boundary
  const bool row_far = rt < -1.0f || rt > (float)H + 1.0f;  // clamped to the same edge either way
  ok = (cfr > mcol) && (cfr < 1.0f - mcol) && (row_far || ((rfr > mrow) && (rfr < 1.0f - mrow))) &&
       (p2 > 1e-24f) && (p2 < 1e24f) && (fabsf(z) < 1e12f);
  int col = (int)cfl;
  col = (col >= W || col < 0) ? col % W : col;
  int row = (int)fminf(fmaxf(rfl, 0.0f), (float)(H - 1));
  return row * W + col;
}

template <int STRIDE>
__device__ __forceinline__ void load_point(const float* __restrict__ pts, int64_t i, float& x, float& y, float& z) {
  if (STRIDE == 4) {
    const float4 v = ld_stream_f4(reinterpret_cast<const float4*>(pts) + i);
    x = v.x; y = v.y; z = v.z;
  } else {
    x = ld_stream_f(pts + i * 3); y = ld_stream_f(pts + i * 3 + 1); z = ld_stream_f(pts + i * 3 + 2);
  }
}

// scratch: int32[B][4] = {lastZero+1 (slot col==0), pix, lastZero+1 (slot col!=0), pix}, zeroed by the caller.
template <int STRIDE>
__global__ void __launch_bounds__(kProjThreads, 1)
project_kernel(const float* __restrict__ points, const int64_t* __restrict__ offsets, int B, int H, int W,
               float hfov, float vmax, float vmin, float mcol, float mrow, unsigned* __restrict__ range,
               int* __restrict__ scratch) {
  const int HW = H * W;
  const float vres = (vmax - vmin) / (float)(H - 1);
  const float col_scale = (float)W / hfov, row_scale = 1.0f / vres;
  const int tid = threadIdx.x;
  __shared__ int s_zero[4];
  __shared__ unsigned s_fix[2];
  __shared__ int s_ndefer;
  __shared__ float4 s_defer[kDeferCap];   // x, y, z, bits(index in frame)

  for (int f = blockIdx.x; f < B; f += gridDim.x) {
    unsigned* img = range + (size_t)f * HW;
    // phase 1: sentinel fill (stays in L2)
    {
      uint4* img4 = reinterpret_cast<uint4*>(img);
      const int n4 = (HW & 3) ? 0 : (HW >> 2);  // vector path needs every frame 16-byte aligned
      const uint4 e = make_uint4(kEmpty, kEmpty, kEmpty, kEmpty);
      for (int i = tid; i < n4; i += kProjThreads) img4[i] = e;
      for (int i = (n4 << 2) + tid; i < HW; i += kProjThreads) img[i] = kEmpty;
    }
    __syncthreads();
    // phase 2: z-buffer
    const int64_t p0 = offsets[f], p1 = offsets[f + 1];
    int* zs = scratch + (size_t)f * 4;
    if (tid == 0) s_ndefer = 0;
    __syncthreads();
    // exact derivation + z-buffer update of one point (also the zero-depth bookkeeping)
    auto exact_update = [&](float x, float y, float z, int rel) {
      const Pixel p = point_to_pixel(x, y, z, H, W, hfov, vmin, vres);
      if (p.depth == 0.0f) {
        // zero depth re-opens the pixel in the reference's sequential loop (cpp_modules.cpp:459)
        const int slot = (p.pix % W == 0) ? 0 : 2;
        atomicMax(&zs[slot], rel + 1);
        zs[slot + 1] = p.pix;
      } else {
        atomicMin(&img[p.pix], __float_as_uint(p.depth));
      }
    };
    const int64_t npts = p1 - p0;
    for (int64_t i0 = 0; i0 < npts; i0 += kProjThreads) {   // whole warps stay in the loop (ballot below)
      const int64_t i = i0 + tid;
      const bool have = i < npts;
      float x = 1.f, y = 0.f, z = 0.f;
      if (have) load_point<STRIDE>(points, p0 + i, x, y, z);
      bool ok;
      const int pix = fast_pixel(x, y, z, H, W, col_scale, vmin, row_scale, mcol, mrow, ok);
      const float depth = sqrtf(x * x + y * y + z * z);
      ok = ok && depth > 0.0f;
      if (have && ok) atomicMin(&img[pix], __float_as_uint(depth));
      // the rest is parked in shared memory and redone densely below
      const unsigned need = __ballot_sync(0xffffffffu, have && !ok);
      if (need) {
        int base = 0;
        if ((tid & 31) == 0) base = atomicAdd(&s_ndefer, __popc(need));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (have && !ok) {
          const int slot = base + __popc(need & lanemask_lt());
          if (slot < kDeferCap) s_defer[slot] = make_float4(x, y, z, __int_as_float((int)i));
          else exact_update(x, y, z, (int)i);              // list full: do it in place
        }
      }
    }
    __syncthreads();
    {
      const int nd = min(s_ndefer, kDeferCap);
      for (int e = tid; e < nd; e += kProjThreads) {
        const float4 v = s_defer[e];
        exact_update(v.x, v.y, v.z, __float_as_int(v.w));
      }
    }
    __threadfence();
    __syncthreads();
    if (tid < 4) s_zero[tid] = ((volatile int*)zs)[tid];
    if (tid < 2) s_fix[tid] = kEmpty;
    __syncthreads();
    if (s_zero[0] > 0 || s_zero[2] > 0) {
      // rare: only points after the last zero-depth hit of a pixel count for that pixel
      for (int64_t i = p0 + tid; i < p1; i += kProjThreads) {
        float x, y, z;
        load_point<STRIDE>(points, i, x, y, z);
        const Pixel p = point_to_pixel(x, y, z, H, W, hfov, vmin, vres);
        if (p.depth == 0.0f) continue;
        const int rel = (int)(i - p0);
        if (s_zero[0] > 0 && p.pix == s_zero[1] && rel >= s_zero[0]) atomicMin(&s_fix[0], __float_as_uint(p.depth));
        if (s_zero[2] > 0 && p.pix == s_zero[3] && rel >= s_zero[2]) atomicMin(&s_fix[1], __float_as_uint(p.depth));
      }
      __syncthreads();
      if (tid == 0) {
        if (s_zero[0] > 0) img[s_zero[1]] = s_fix[0];
        if (s_zero[2] > 0) img[s_zero[3]] = s_fix[1];
        __threadfence();
      }
      __syncthreads();
    }
    // phase 3: sentinel -> 0.0f in place (read through L2)
    {
      uint4* img4 = reinterpret_cast<uint4*>(img);
      const int n4 = (HW & 3) ? 0 : (HW >> 2);
      for (int i = tid; i < n4; i += kProjThreads) {
        uint4 v = __ldcg(img4 + i);
        v.x = v.x == kEmpty ? 0u : v.x; v.y = v.y == kEmpty ? 0u : v.y;
        v.z = v.z == kEmpty ? 0u : v.z; v.w = v.w == kEmpty ? 0u : v.w;
        img4[i] = v;
      }
      for (int i = (n4 << 2) + tid; i < HW; i += kProjThreads) {
        unsigned v = __ldcg(img + i);
        img[i] = v == kEmpty ? 0u : v;
      }
    }
    __syncthreads();
  }
}

__global__ void range_to_xyz_kernel(const float* __restrict__ range, const float* __restrict__ lut, int64_t total,
                                    int HW, float* __restrict__ xyz) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const float r = range[i];
  const int p = (int)(i % HW);
  xyz[i * 3 + 0] = r * lut[p * 3 + 0];
  xyz[i * 3 + 1] = r * lut[p * 3 + 1];
  xyz[i * 3 + 2] = r * lut[p * 3 + 2];
}

}  // namespace rpcc

using namespace rpcc;

extern "C" int rpcc_project_batch(const float* points, int stride, const int64_t* offsets, int B, int H, int W,
                                  float hfov, float vmax, float vmin, float* range, int32_t* scratch, void* stream) {
  RPCC_REQUIRE(points && offsets && range && scratch, "null pointer");
  RPCC_REQUIRE(stride == 3 || stride == 4, "stride must be 3 or 4");
  RPCC_REQUIRE(B >= 0 && H >= 2 && W >= 1, "bad shape");
  RPCC_REQUIRE(stride != 4 || (reinterpret_cast<uintptr_t>(points) & 15) == 0, "stride-4 points must be 16-byte aligned");
  if (B == 0) return RPCC_OK;
  cudaStream_t st = as_stream(stream);
  RPCC_CUDA(cudaMemsetAsync(scratch, 0, sizeof(int32_t) * 4 * (size_t)B, st));
  const int grid = B < sm_count() ? B : sm_count();
  // margins of fast_pixel(): 4x the worst-case distance between the fast and the reference
  // coordinate.  col: 3 ulp of an angle in [4,8) (4.77e-7 each) scaled to columns + 4 ulp of W;
  // row: 3 ulp of an angle in [1,2) (1.19e-7 each) scaled to rows + 4 ulp of H.
  const float vres = (vmax - vmin) / (float)(H - 1);
  const float ulp_w = ldexpf(1.0f, ilogbf((float)(W > 1 ? W : 2)) - 23), ulp_h = ldexpf(1.0f, ilogbf((float)H) - 23);
  float mcol = 4.0f * (3.0f * 4.77e-7f * (float)W / fabsf(hfov) + 4.0f * ulp_w);
  float mrow = 4.0f * (3.0f * 1.2e-7f / fabsf(vres) + 4.0f * ulp_h);
  if (!(mcol < 0.25f)) mcol = 1.0f;   // margins this large disable the fast path (everything goes exact)
  if (!(mrow < 0.25f)) mrow = 1.0f;
  if (stride == 4)
    project_kernel<4><<<grid, kProjThreads, 0, st>>>(points, offsets, B, H, W, hfov, vmax, vmin, mcol, mrow,
                                                     reinterpret_cast<unsigned*>(range), scratch);
  else
    project_kernel<3><<<grid, kProjThreads, 0, st>>>(points, offsets, B, H, W, hfov, vmax, vmin, mcol, mrow,
                                                     reinterpret_cast<unsigned*>(range), scratch);
  RPCC_LAUNCH_CHECK("project_kernel");
  return RPCC_OK;
}

extern "C" int rpcc_range_to_xyz_batch(const float* range, const float* lut, int B, int HW, float* xyz, void* stream) {
  RPCC_REQUIRE(range && lut && xyz, "null pointer");
  const int64_t total = (int64_t)B * HW;
  if (total == 0) return RPCC_OK;
  range_to_xyz_kernel<<<(unsigned)((total + 255) / 256), 256, 0, as_stream(stream)>>>(range, lut, total, HW, xyz);
  RPCC_LAUNCH_CHECK("range_to_xyz_kernel");
  return RPCC_OK;
}
