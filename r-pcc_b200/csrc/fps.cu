// fps.cu -- stage 2a: furthest point sampling.
//
// Replaces ops/fps/src/sampling_gpu.cu:24-184 (furthest_point_sampling_kernel + launcher) and,
// in the fused form, the mask + FPS lines of PointCloudSegment.segment
// (utils/segment_utils.py:137-141).
//
// Exactness contract: the seed sequence is identical to the reference kernel's:
//   d  = fma(dz,dz, fma(dx,dx, dy*dy))            (the contraction nvcc emits for sampling_gpu.cu:64)
//   t[k] = min(d, t[k]), t starts at 1e10
//   winner = max t, ties -> min bitrev(k mod bs) over log2(bs) bits, then min k
// (the reference's per-thread strided scan with strict '>' followed by a shared-memory tree that
// keeps the lower slot on ties, sampling_gpu.cu:16-21,54-134; bs = min(1024, 2^floor(log2 n))).
// Both kernels below carry the tie rule inside a 64-bit key: (bits(t) << 32) | ~tiekey with
// tiekey = (bitrev(k mod bs) << (32-log2 bs)) | (k / bs), so one unsigned max does it.
//
// Two kernels:
//  * fps_generic_kernel: any (B,n,3) input, one CTA per batch item, state in global memory --
//    the drop-in for furthest_point_sampling_wrapper.
//  * segment_fps_kernel: the pipeline path.  One thread-block CLUSTER of 8 CTAs per frame keeps
//    every candidate point on chip for all m-1 rounds (4 slots per thread in registers, the rest
//    in shared memory), builds the masked point cloud straight from range x LUT, keeps a single
//    representative of the identical origin points (ground / empty pixels) per residue class,
//    and exchanges the per-CTA winners through distributed shared memory once per round.
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace rpcc {

__device__ __forceinline__ float fps_dist(float x, float y, float z, float x1, float y1, float z1) {
  const float dx = x - x1, dy = y - y1, dz = z - z1;
  return __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
}

__device__ __forceinline__ unsigned long long warp_key_max(unsigned long long k) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long other = __shfl_xor_sync(0xffffffffu, k, o);
    k = other > k ? other : k;
  }
  return k;
}

__device__ __forceinline__ unsigned long long make_key(float d2, unsigned tiekey) {
  // d2 >= 0 (NaN coordinates are not supported): its bit pattern orders like the value
  return ((unsigned long long)__float_as_uint(d2) << 32) | (unsigned long long)(0xFFFFFFFFu - tiekey);
}

// ------------------------------------------------------------------------------------------------
// generic kernel: one CTA per batch item, temp in global memory (the reference's buffer)
// ------------------------------------------------------------------------------------------------
constexpr int kGenThreads = 1024;

__global__ void __launch_bounds__(kGenThreads, 1)
fps_generic_kernel(const float* __restrict__ points, int n, int m, int log2bs, float* __restrict__ temp,
                   int* __restrict__ idx) {
  if (m <= 0) return;
  const int b = blockIdx.x;
  points += (size_t)b * n * 3;
  temp += (size_t)b * n;
  idx += (size_t)b * m;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const unsigned bsmask = (1u << log2bs) - 1u;
  __shared__ unsigned long long s_part[2][32];

  for (int k = tid; k < n; k += kGenThreads) temp[k] = 1e10f;
  if (tid == 0) idx[0] = 0;
  int old = 0;
  for (int j = 1; j < m; ++j) {
    const float x1 = points[(size_t)old * 3], y1 = points[(size_t)old * 3 + 1], z1 = points[(size_t)old * 3 + 2];
    unsigned long long best = 0;
    for (int k = tid; k < n; k += kGenThreads) {
      const float d = fps_dist(points[(size_t)k * 3], points[(size_t)k * 3 + 1], points[(size_t)k * 3 + 2], x1, y1, z1);
      const float d2 = fminf(d, temp[k]);
      temp[k] = d2;
      const unsigned t = (unsigned)k & bsmask;
      const unsigned tiekey = log2bs ? (__brev(t) & ~(0xFFFFFFFFu >> log2bs)) | ((unsigned)k >> log2bs) : (unsigned)k;
      const unsigned long long key = make_key(d2, tiekey);
      best = key > best ? key : best;
    }
    best = warp_key_max(best);
    if (lane == 0) s_part[j & 1][warp] = best;
    __syncthreads();
    best = warp_key_max(s_part[j & 1][lane]);
    const unsigned tiekey = 0xFFFFFFFFu - (unsigned)(best & 0xFFFFFFFFull);
    if (log2bs) {
      const unsigned t = __brev(tiekey & ~(0xFFFFFFFFu >> log2bs));
      old = (int)(((tiekey & (0xFFFFFFFFu >> log2bs)) << log2bs) | t);
    } else {
      old = (int)tiekey;
    }
    if (tid == 0) idx[j] = old;
  }
}

// ------------------------------------------------------------------------------------------------
// fused mask + FPS, one 8-CTA cluster per frame
// ------------------------------------------------------------------------------------------------
constexpr int kCl = 8;            // CTAs per cluster
constexpr int kFpsThreads = 1024; // = the reference's block size, so thread id == residue class
constexpr int kRegSlots = 4;

struct __align__(16) FpsRecord { unsigned long long key; float x, y, z; int pad; };

// torch float32 semantics of utils/segment_utils.py:137-139 for one pixel:
// |sum(pc*g)+g3| / norm(g) > thr ? pc : 0   (size-3 reductions associate as (t0+t2)+t1, see assign.cu)
__device__ __forceinline__ void masked_point(float r, const float* __restrict__ lut3, float g0, float g1, float g2,
                                             float g3, float gnorm, float thr, float& x, float& y, float& z) {
  x = r * lut3[0]; y = r * lut3[1]; z = r * lut3[2];
  const float s = torch_sum3(x * g0, y * g1, z * g2);
  const float dif = fabsf(s + g3) / gnorm;
  if (!(dif > thr)) { x = 0.f; y = 0.f; z = 0.f; }
}

template <int SLOTS>
__global__ void __launch_bounds__(kFpsThreads, 1)
segment_fps_kernel(const float* __restrict__ range, const float* __restrict__ lut, const float* __restrict__ ground,
                   int B, int HW, int m, float thr, int* __restrict__ center_idx, float* __restrict__ centers) {
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int ncluster = gridDim.x / kCl;
  const int cid = blockIdx.x / kCl;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int SM = SLOTS - kRegSlots;  // shared-memory slots

  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* sx = reinterpret_cast<float*>(smem_raw);
  float* sy = sx + SM * kFpsThreads;
  float* sz = sy + SM * kFpsThreads;
  unsigned char* sj = reinterpret_cast<unsigned char*>(sz + SM * kFpsThreads);   // [SLOTS][1024] row of each slot
  unsigned long long* s_part = reinterpret_cast<unsigned long long*>(sj + SLOTS * kFpsThreads);  // [2][32]
  FpsRecord* s_rec = reinterpret_cast<FpsRecord*>(s_part + 64);                   // [2][kCl]

  const int J = (HW + kFpsThreads - 1) / kFpsThreads;
  const unsigned tie_hi = __brev((unsigned)tid) & 0xFFC00000u;  // bitrev10(tid) in the top 10 bits
  unsigned par = 0;

  for (int f = cid; f < B; f += ncluster) {
    const float* rg = range + (size_t)f * HW;
    const float g0 = ground[f * 4], g1 = ground[f * 4 + 1], g2 = ground[f * 4 + 2], g3 = ground[f * 4 + 3];
    const float gnorm = sqrtf(torch_sum3(g0 * g0, g1 * g1, g2 * g2));

    // ---- gather this thread's share of residue class `tid`: items are the non-origin points plus
    //      the first origin pixel, dealt round-robin (by rank inside the class) to the 8 CTAs.
    float rx[kRegSlots], ry[kRegSlots], rz[kRegSlots];
#pragma unroll
    for (int s = 0; s < kRegSlots; ++s) { rx[s] = 0.f; ry[s] = 0.f; rz[s] = 0.f; }
    int mine = 0, seen = 0;
    bool origin_seen = false;
#pragma unroll 5
    for (int j = 0; j < J; ++j) {
      const int k = j * kFpsThreads + tid;
      if (k >= HW) break;
      float x, y, z;
      masked_point(rg[k], lut + (size_t)k * 3, g0, g1, g2, g3, gnorm, thr, x, y, z);
      const bool origin = (x == 0.f) && (y == 0.f) && (z == 0.f);
      if (origin && origin_seen) continue;
      origin_seen = origin_seen || origin;
      const int r = seen++;
      if ((r & (kCl - 1)) != rank) continue;
      if (mine < kRegSlots) {
#pragma unroll
        for (int s = 0; s < kRegSlots; ++s) if (mine == s) { rx[s] = x; ry[s] = y; rz[s] = z; }
      } else if (mine < SLOTS) {
        const int o = (mine - kRegSlots) * kFpsThreads + tid;
        sx[o] = x; sy[o] = y; sz[o] = z;
      }
      if (mine < SLOTS) sj[mine * kFpsThreads + tid] = (unsigned char)j;
      ++mine;
    }
    mine = mine < SLOTS ? mine : SLOTS;
    int wmax = mine;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) wmax = max(wmax, __shfl_xor_sync(0xffffffffu, wmax, o));

    float temp[SLOTS];
#pragma unroll
    for (int s = 0; s < SLOTS; ++s) temp[s] = 1e10f;

    // seed 0 is flat index 0 (sampling_gpu.cu:44-46)
    float x1, y1, z1;
    masked_point(rg[0], lut, g0, g1, g2, g3, gnorm, thr, x1, y1, z1);
    if (rank == 0 && tid == 0) {
      center_idx[(size_t)f * m] = 0;
      centers[(size_t)f * m * 3 + 0] = x1; centers[(size_t)f * m * 3 + 1] = y1; centers[(size_t)f * m * 3 + 2] = z1;
    }

    for (int j = 1; j < m; ++j) {
      par ^= 1u;
      float best = -1.f;
      int bs = 0;
#pragma unroll
      for (int s = 0; s < SLOTS; ++s) {
        if (s >= wmax) break;  // warp-uniform
        if (s < mine) {
          float px, py, pz;
          if (s < kRegSlots) { px = rx[s < kRegSlots ? s : 0]; py = ry[s < kRegSlots ? s : 0]; pz = rz[s < kRegSlots ? s : 0]; }
          else { const int o = (s - kRegSlots) * kFpsThreads + tid; px = sx[o]; py = sy[o]; pz = sz[o]; }
          const float d2 = fminf(fps_dist(px, py, pz, x1, y1, z1), temp[s]);
          temp[s] = d2;
          if (d2 > best) { best = d2; bs = s; }
        }
      }
      unsigned long long mykey = 0;
      if (mine > 0) mykey = make_key(best, tie_hi | (unsigned)sj[bs * kFpsThreads + tid]);
      unsigned long long key = warp_key_max(mykey);
      if (lane == 0) s_part[par * 32 + warp] = key;
      __syncthreads();
      key = warp_key_max(s_part[par * 32 + lane]);
      // the owner of the CTA winner publishes it to every CTA of the cluster
      if ((key != 0 && mykey == key) || (key == 0 && tid == 0)) {
        FpsRecord rec;
        rec.key = key; rec.pad = 0;
        if (key != 0) {
          if (bs < kRegSlots) {
            rec.x = rx[0]; rec.y = ry[0]; rec.z = rz[0];
#pragma unroll
            for (int s = 1; s < kRegSlots; ++s) if (bs == s) { rec.x = rx[s]; rec.y = ry[s]; rec.z = rz[s]; }
          } else {
            const int o = (bs - kRegSlots) * kFpsThreads + tid;
            rec.x = sx[o]; rec.y = sy[o]; rec.z = sz[o];
          }
        } else {
          rec.x = 0.f; rec.y = 0.f; rec.z = 0.f;
        }
#pragma unroll
        for (int r = 0; r < kCl; ++r) {
          FpsRecord* dst = cluster.map_shared_rank(s_rec, r) + par * kCl + rank;
          *dst = rec;
        }
      }
      cluster.sync();
      FpsRecord w = s_rec[par * kCl];
#pragma unroll
      for (int r = 1; r < kCl; ++r) {
        const FpsRecord o = s_rec[par * kCl + r];
        if (o.key > w.key) w = o;
      }
      x1 = w.x; y1 = w.y; z1 = w.z;
      if (rank == 0 && tid == 0) {
        const unsigned tiekey = 0xFFFFFFFFu - (unsigned)(w.key & 0xFFFFFFFFull);
        const int k = (int)(((tiekey & 0x3FFFFFu) << 10) | (__brev(tiekey & 0xFFC00000u)));
        center_idx[(size_t)f * m + j] = k;
        float* c = centers + ((size_t)f * m + j) * 3;
        c[0] = x1; c[1] = y1; c[2] = z1;
      }
    }
    // slots are rewritten by the next frame's gather only after every thread of this CTA left the
    // round loop; remote records use the other parity buffer first, so no extra cluster barrier.
    __syncthreads();
  }
}

template <int SLOTS>
static size_t fps_smem_bytes() {
  return (size_t)(SLOTS - kRegSlots) * kFpsThreads * 12 + (size_t)SLOTS * kFpsThreads + 64 * 8 + 2 * kCl * sizeof(FpsRecord);
}

template <int SLOTS>
static int launch_segment_fps(const float* range, const float* lut, const float* ground, int B, int HW, int m,
                              float thr, int* center_idx, float* centers, cudaStream_t st) {
  auto kern = segment_fps_kernel<SLOTS>;
  const size_t smem = fps_smem_bytes<SLOTS>();
  RPCC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kCl; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.blockDim = dim3(kFpsThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cfg.gridDim = dim3(kCl);
  int nclusters = 0;
  if (cudaOccupancyMaxActiveClusters(&nclusters, kern, &cfg) != cudaSuccess || nclusters <= 0) {
    cudaGetLastError();
    nclusters = sm_count() / kCl;
  }
  if (nclusters > B) nclusters = B;
  cfg.gridDim = dim3(kCl * nclusters);
  RPCC_CUDA(cudaLaunchKernelEx(&cfg, kern, range, lut, ground, B, HW, m, thr, center_idx, centers));
  count_launch();
  return RPCC_OK;
}

}  // namespace rpcc

using namespace rpcc;

extern "C" int rpcc_fps_batch(const float* points, int B, int n, int m, float* temp, int32_t* idx, void* stream) {
  RPCC_REQUIRE(points && temp && idx, "null pointer");
  RPCC_REQUIRE(B >= 0 && n >= 1 && m >= 0, "bad shape");
  if (B == 0 || m == 0) return RPCC_OK;
  int log2bs = 0;
  while ((2 << log2bs) <= n && log2bs < 10) ++log2bs;  // bs = min(1024, 2^floor(log2 n)), sampling_gpu.cu:9-13
  fps_generic_kernel<<<B, kGenThreads, 0, as_stream(stream)>>>(points, n, m, log2bs, temp, idx);
  RPCC_LAUNCH_CHECK("fps_generic_kernel");
  return RPCC_OK;
}

extern "C" int rpcc_segment_fps_batch(const float* range, const float* lut, const float* ground, int B, int H, int W,
                                      int m, float ground_thr, int32_t* center_idx, float* centers, void* stream) {
  RPCC_REQUIRE(range && lut && ground && center_idx && centers, "null pointer");
  const int HW = H * W;
  RPCC_REQUIRE(HW >= 1024, "range image must have at least 1024 pixels");
  RPCC_REQUIRE(m >= 1, "need at least one seed");
  if (B == 0) return RPCC_OK;
  const int J = (HW + kFpsThreads - 1) / kFpsThreads;
  const int need = (J + kCl - 1) / kCl;  // slots per thread
  cudaStream_t st = as_stream(stream);
  if (need <= 4) return launch_segment_fps<4>(range, lut, ground, B, HW, m, ground_thr, center_idx, centers, st);
  if (need <= 10) return launch_segment_fps<10>(range, lut, ground, B, HW, m, ground_thr, center_idx, centers, st);
  if (need <= 16) return launch_segment_fps<16>(range, lut, ground, B, HW, m, ground_thr, center_idx, centers, st);
  if (need <= 20) return launch_segment_fps<20>(range, lut, ground, B, HW, m, ground_thr, center_idx, centers, st);
  set_error("rpcc_segment_fps_batch: H*W = %d exceeds the on-chip capacity (163840 pixels)", HW);
  return RPCC_ERR_CAPACITY;
}
