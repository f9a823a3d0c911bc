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
#include <stdlib.h>
#include <string.h>

#include <type_traits>

#include "common.cuh"

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
constexpr int kFpsThreads = 1024; // = the reference's block size, so thread id == residue class k mod 1024
constexpr int kRegSlots = 6;      // candidate points per thread also kept in registers (all are mirrored in smem)
constexpr unsigned kRecBytes = 32;

// ---- raw shared-memory / cluster / mbarrier primitives (sm_90+ PTX)
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float lds_f32(unsigned a) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a)); return v; }
__device__ __forceinline__ unsigned lds_u8(unsigned a) { unsigned v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint4 lds_v4(unsigned a) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts_f32(unsigned a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }
__device__ __forceinline__ void sts_u8(unsigned a, unsigned v) { asm volatile("st.shared.u8 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts_v4(unsigned a, uint4 v) {
  asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ unsigned mapa_u32(unsigned a, unsigned rank) {
  unsigned r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_init(unsigned a, unsigned count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned a, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned a, unsigned parity) {
  unsigned ok;
  do {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(a), "r"(parity) : "memory");
  } while (!ok);
}
// 16 bytes into a peer CTA's shared memory; the peer's mbarrier is credited with the bytes when they land
__device__ __forceinline__ void st_async_v4(unsigned raddr, uint4 v, unsigned rmbar) {
  asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1,%2,%3,%4}, [%5];"
               ::"r"(raddr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"(rmbar) : "memory");
}
__device__ __forceinline__ void cluster_barrier() {
  asm volatile("barrier.cluster.arrive.aligned;\n barrier.cluster.wait.aligned;" ::: "memory");
}
__device__ __forceinline__ unsigned cluster_ctarank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }

// torch float32 semantics of utils/segment_utils.py:137-139 for one pixel:
// |sum(pc*g)+g3| / norm(g) > thr ? pc : 0   (size-3 reductions associate as torch_sum3)
__device__ __forceinline__ void masked_point(float r, const float* __restrict__ lut3, float g0, float g1, float g2,
                                             float g3, float gnorm, float thr, float& x, float& y, float& z) {
  x = r * lut3[0]; y = r * lut3[1]; z = r * lut3[2];
  const float s = torch_sum3(x * g0, y * g1, z * g2);
  const float dif = fabsf(s + g3) / gnorm;
  if (!(dif > thr)) { x = 0.f; y = 0.f; z = 0.f; }
}

// winner among the lanes of a warp: max d (bit pattern of a non-negative float), then min tie key.
// Returns the owning lane (same answer in every lane); lanes without a candidate pass d = 0, tk = ~0.
__device__ __forceinline__ int warp_winner(unsigned d, unsigned tk, unsigned& dmax, unsigned& tkmin) {
  dmax = __reduce_max_sync(0xffffffffu, d);
  const unsigned t = d == dmax ? tk : 0xFFFFFFFFu;
  tkmin = __reduce_min_sync(0xffffffffu, t);
  return __ffs(__ballot_sync(0xffffffffu, t == tkmin)) - 1;
}

// Shared-memory map (bytes):  [0,1024) per-warp winners 2 x 32 x 16
//                             [1024,1536) per-CTA winners 2 x 8 x 32 (written by the peers, st.async)
//                             [1536,1552) two mbarriers
//                             [2048, ...) point slots: x[SLOTS][1024], y, z (f32) then row[SLOTS][1024] (u8)
template <int SLOTS>
__global__ void __launch_bounds__(kFpsThreads, 1)
segment_fps_kernel(const float* __restrict__ range, const float* __restrict__ lut, const float* __restrict__ ground,
                   int B, int HW, int m, float thr, int* __restrict__ center_idx, float* __restrict__ centers) {
  const unsigned rank = cluster_ctarank();
  const int ncluster = gridDim.x / kCl;
  const int cid = blockIdx.x / kCl;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int RS = SLOTS < kRegSlots ? SLOTS : kRegSlots;

  extern __shared__ __align__(128) unsigned char smem_raw[];
  const unsigned sbase = smem_u32(smem_raw);
  const unsigned a_part = sbase, a_rec = sbase + 1024, a_bar = sbase + 1536;
  constexpr unsigned kPlane = SLOTS * kFpsThreads * 4;
  const unsigned a_x = sbase + 2048 + tid * 4, a_y = a_x + kPlane, a_z = a_y + kPlane;   // + slot * 4096
  const unsigned a_j = sbase + 2048 + 3 * kPlane + tid;                                  // + slot * 1024

  if (tid == 0) {
    mbar_init(a_bar, 1);
    mbar_init(a_bar + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  cluster_barrier();   // peers may now signal our mbarriers

  const unsigned tie_hi = __brev((unsigned)tid) & 0xFFC00000u;  // bitrev10(tid) in the top 10 bits
  unsigned round = 0;                                            // rounds since kernel start (buffer / phase bookkeeping)

  for (int f = cid; f < B; f += ncluster) {
    const float* rg = range + (size_t)f * HW;
    const float g0 = ground[f * 4], g1 = ground[f * 4 + 1], g2 = ground[f * 4 + 2], g3 = ground[f * 4 + 3];
    const float gnorm = sqrtf(torch_sum3(g0 * g0, g1 * g1, g2 * g2));

    // ---- gather: CTA `rank` owns the rows j = rank, rank+8, ... of the reference's (row j, thread tid)
    //      layout k = j*1024 + tid.  Kept per thread, in increasing k: every non-origin point and the
    //      first origin point (ground / empty pixels are all the same point; the first one carries the
    //      best tie key of its residue class inside this CTA).  Unused slots get temp = 0: they can never
    //      exceed a running maximum, so the round loop needs no per-lane guard.
    float rx[RS], ry[RS], rz[RS];
#pragma unroll
    for (int s = 0; s < RS; ++s) { rx[s] = 0.f; ry[s] = 0.f; rz[s] = 0.f; }
    int mine = 0;
    bool origin_seen = false;
#pragma unroll 4
    for (int jj = 0; jj < SLOTS; ++jj) {
      const int j = jj * kCl + (int)rank;
      const int k = j * kFpsThreads + tid;
      if (k < HW) {
        float x, y, z;
        masked_point(rg[k], lut + (size_t)k * 3, g0, g1, g2, g3, gnorm, thr, x, y, z);
        const bool origin = (x == 0.f) && (y == 0.f) && (z == 0.f);
        if (!(origin && origin_seen)) {
          origin_seen = origin_seen || origin;
#pragma unroll
          for (int s = 0; s < RS; ++s) if (mine == s) { rx[s] = x; ry[s] = y; rz[s] = z; }
          sts_f32(a_x + mine * 4096, x); sts_f32(a_y + mine * 4096, y); sts_f32(a_z + mine * 4096, z);
          sts_u8(a_j + mine * 1024, (unsigned)j);
          ++mine;
        }
      }
    }
    int wmax = mine;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) wmax = max(wmax, __shfl_xor_sync(0xffffffffu, wmax, o));

    float temp[SLOTS];
#pragma unroll
    for (int s = 0; s < SLOTS; ++s) temp[s] = s < mine ? 1e10f : 0.0f;

    // seed 0 is flat index 0 (sampling_gpu.cu:44-46)
    float x1, y1, z1;
    masked_point(rg[0], lut, g0, g1, g2, g3, gnorm, thr, x1, y1, z1);
    if (rank == 0 && tid == 0) {
      center_idx[(size_t)f * m] = 0;
      centers[(size_t)f * m * 3 + 0] = x1; centers[(size_t)f * m * 3 + 1] = y1; centers[(size_t)f * m * 3 + 2] = z1;
    }

    for (int j = 1; j < m; ++j) {
      const unsigned par = round & 1u, phase = (round >> 1) & 1u;
      ++round;
      if (tid == 0) mbar_expect_tx(a_bar + par * 8, kCl * kRecBytes);
      float best = -1.f;
      int bs = 0;
#pragma unroll
      for (int s = 0; s < SLOTS; ++s) {
        if (s >= RS && s >= wmax) break;  // warp-uniform
        float px, py, pz;
        if (s < RS) { px = rx[s < RS ? s : 0]; py = ry[s < RS ? s : 0]; pz = rz[s < RS ? s : 0]; }
        else { px = lds_f32(a_x + s * 4096); py = lds_f32(a_y + s * 4096); pz = lds_f32(a_z + s * 4096); }
        const float d2 = fminf(fps_dist(px, py, pz, x1, y1, z1), temp[s]);
        temp[s] = d2;
        if (d2 > best) { best = d2; bs = s; }
      }
      // ---- warp winner -> shared (16 bytes: d, tie key, slot)
      {
        const unsigned d = mine > 0 ? __float_as_uint(best) : 0u;
        const unsigned dmax = __reduce_max_sync(0xffffffffu, d);
        // the tie key (one shared-memory read) is only needed by lanes that hold the warp maximum
        const unsigned tk = (mine > 0 && d == dmax) ? (tie_hi | lds_u8(a_j + bs * 1024)) : 0xFFFFFFFFu;
        const unsigned tkmin = __reduce_min_sync(0xffffffffu, tk);
        if (tk == tkmin && (tk != 0xFFFFFFFFu || lane == 0))
          sts_v4(a_part + (par * 32 + warp) * 16, make_uint4(dmax, tkmin, (unsigned)bs, 0u));
      }
      __syncthreads();
      // ---- CTA winner: warp 0 only; its coordinates come from the shared mirror; published to every CTA
      //      of the cluster with st.async, which credits the receiver's mbarrier
      if (warp == 0) {
        const uint4 pr = lds_v4(a_part + (par * 32 + lane) * 16);
        unsigned dmax, tkmin;
        const int own = warp_winner(pr.x, pr.y, dmax, tkmin);
        if (lane == own) {
          float wx = 0.f, wy = 0.f, wz = 0.f;
          if (tkmin != 0xFFFFFFFFu) {
            const unsigned t = __brev(tkmin & 0xFFC00000u);           // owning thread
            const unsigned o = pr.z * 4096 + t * 4 - tid * 4;         // a_x is relative to this thread
            wx = lds_f32(a_x + o); wy = lds_f32(a_y + o); wz = lds_f32(a_z + o);
          }
          const uint4 r0 = make_uint4(dmax, tkmin, __float_as_uint(wx), __float_as_uint(wy));
          const uint4 r1 = make_uint4(__float_as_uint(wz), 0u, 0u, 0u);
          const unsigned dst = a_rec + (par * kCl + rank) * kRecBytes;
#pragma unroll
          for (unsigned r = 0; r < kCl; ++r) {
            const unsigned rd = mapa_u32(dst, r), rb = mapa_u32(a_bar + par * 8, r);
            st_async_v4(rd, r0, rb);
            st_async_v4(rd + 16, r1, rb);
          }
        }
      }
      mbar_wait(a_bar + par * 8, phase);
      // ---- cluster winner: every warp reduces the 8 records on its own
      {
        uint4 w = make_uint4(0u, 0xFFFFFFFFu, 0u, 0u);
        float wz = 0.f;
        if (lane < kCl) {
          w = lds_v4(a_rec + (par * kCl + lane) * kRecBytes);
          wz = lds_f32(a_rec + (par * kCl + lane) * kRecBytes + 16);
        }
        unsigned dmax, tkmin;
        const int own = warp_winner(w.x, w.y, dmax, tkmin);
        x1 = __uint_as_float(__shfl_sync(0xffffffffu, w.z, own));
        y1 = __uint_as_float(__shfl_sync(0xffffffffu, w.w, own));
        z1 = __shfl_sync(0xffffffffu, wz, own);
        if (rank == 0 && tid == 0) {
          const int k = (int)(((tkmin & 0x3FFFFFu) << 10) | (__brev(tkmin & 0xFFC00000u)));
          center_idx[(size_t)f * m + j] = k;
          float* c = centers + ((size_t)f * m + j) * 3;
          c[0] = x1; c[1] = y1; c[2] = z1;
        }
      }
    }
    // the next frame's gather rewrites the shared slots: every thread must have left the round loop
    __syncthreads();
  }
  cluster_barrier();   // no peer may still be writing into this CTA's shared memory when it exits
}

template <int SLOTS>
static size_t fps_smem_bytes() {
  return 2048 + (size_t)SLOTS * kFpsThreads * 13;
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


// ------------------------------------------------------------------------------------------------
// fused mask + FPS with exact bucket pruning: one CTA per frame, state in L2
// ------------------------------------------------------------------------------------------------
// Round j of FPS only lowers the running distance t[k] of points that are closer to the new centre
// than to every earlier one -- a few per cent of the image.  Pixels are grouped into buckets of 32
// consecutive pixels (neighbours on one beam); a bucket keeps its bounding box, its maximum t
// and the tie key of the point holding it.  A bucket whose box is farther from the new centre than
// sqrt(max t) cannot change (the test carries a 1e-5 relative slack, two orders of magnitude above
// the rounding of either side), so a round touches only the buckets around the new centre and the
// argmax comes from the per-bucket maxima.  Touched points are updated with the reference's own
// arithmetic, and maxima / tie keys are exact, so the seed sequence is the reference's, bit for bit.
// Per frame this replaces 99 passes over 128 000 points by one pass plus ~100 buckets per round:
// a single CTA per frame suffices, no cross-SM exchange, 148+ frames in flight.
//   state in global memory (L2-resident): t[HW] per CTA, bit 31 = "masked to the origin"
//   state on chip: boxes in shared memory (6 floats per bucket), max t / tie key in registers.
constexpr unsigned kNoTie = 0xFFFFFFFFu;
#ifndef RPCC_FPS_ONEWINNER
#define RPCC_FPS_ONEWINNER 1
#endif

__device__ __forceinline__ int ford(float f) { const int b = __float_as_int(f); return b ^ ((b >> 31) & 0x7fffffff); }
__device__ __forceinline__ float ordf(int i) { return __int_as_float(i ^ ((i >> 31) & 0x7fffffff)); }

// ---- first pass as its own kernel (the default): round 1 over every bucket is plain streaming work -- it wants registers
//      and many independent warps, the round loop below wants two resident CTAs of 1024 threads at 32 registers.  One warp
//      takes a run of consecutive buckets of one frame and leaves, per bucket, a 32-byte record {box, max t, tie key} and,
//      per pixel, t (bit 31 = "masked to the origin") for the round kernel.
struct __align__(16) FpsBucket { float x0, y0, z0, x1, y1, z1; unsigned bmax, btk; };
constexpr int kFirstThreads = 256;

// round 1 for the 32 pixels of bucket b (one per lane); every lane returns the same record
__device__ __forceinline__ FpsBucket fps_first_bucket(const float* __restrict__ rg, const float* __restrict__ lut, int b, int lane,
                                                      int HW, float g0, float g1, float g2, float g3, float gnorm, float thr,
                                                      float x1, float y1, float z1, unsigned* __restrict__ temp) {
  const float INF = __int_as_float(0x7f800000);
  FpsBucket o = {INF, INF, INF, -INF, -INF, -INF, 0u, kNoTie};
  const int p = (b << 5) + lane;
  const bool inb = p < HW;
  float x = 0.f, y = 0.f, z = 0.f;
  if (inb) masked_point(ld_stream_f(rg + p), lut + (size_t)p * 3, g0, g1, g2, g3, gnorm, thr, x, y, z);
  const bool origin = (x == 0.f) && (y == 0.f) && (z == 0.f);
  const float tv = fminf(fps_dist(x, y, z, x1, y1, z1), 1e10f);
  if (inb) temp[p] = __float_as_uint(tv) | (origin ? 0x80000000u : 0u);
  const unsigned nb = inb ? __float_as_uint(tv) : 0u;
  const unsigned newmax = __reduce_max_sync(0xffffffffu, nb);
  // A bucket whose running distances are all zero -- ground and empty pixels when seed 0 is one of them, about half
  // of a frame's buckets -- can never change, and can only be chosen when every point of the frame sits at zero, in
  // which case the reference's tie rule picks pixel 0 (tie key 0): bucket 0 is always kept, the others keep the
  // defaults (maximum 0, no tie key, empty box).
  if (newmax == 0u && b != 0) return o;                     // warp-uniform
  const unsigned tk = (inb && nb == newmax) ? ((__brev((unsigned)p) & 0xFFC00000u) | ((unsigned)p >> 10)) : kNoTie;
  o.bmax = newmax;
  o.btk = __reduce_min_sync(0xffffffffu, tk);
  // box over the points that can still change (t > 0).  Coordinates inside +-250 m are shifted into (0, 512), where
  // the bit patterns order like the values (one REDUX each, no order-preserving transform); the shift rounds to
  // 3e-5 m and the box is widened by 1e-4 m on every side, which keeps the pruning conservative.
  const bool live = inb && tv > 0.f;
  if (!__any_sync(0xffffffffu, live && !(fmaxf(fmaxf(fabsf(x), fabsf(y)), fabsf(z)) < 250.f))) {
    const unsigned ux = __float_as_uint(x + 256.f), uy = __float_as_uint(y + 256.f), uz = __float_as_uint(z + 256.f);
    o.x0 = __uint_as_float(__reduce_min_sync(0xffffffffu, live ? ux : 0x7f800000u)) - 256.0001f;
    o.y0 = __uint_as_float(__reduce_min_sync(0xffffffffu, live ? uy : 0x7f800000u)) - 256.0001f;
    o.z0 = __uint_as_float(__reduce_min_sync(0xffffffffu, live ? uz : 0x7f800000u)) - 256.0001f;
    o.x1 = __uint_as_float(__reduce_max_sync(0xffffffffu, live ? ux : 0u)) - 255.9999f;
    o.y1 = __uint_as_float(__reduce_max_sync(0xffffffffu, live ? uy : 0u)) - 255.9999f;
    o.z1 = __uint_as_float(__reduce_max_sync(0xffffffffu, live ? uz : 0u)) - 255.9999f;
  } else {                                                  // coordinates beyond any lidar's range (or NaN): exact order
    const int big = 0x7fffffff;
    o.x0 = ordf(__reduce_min_sync(0xffffffffu, live ? ford(x) : big));
    o.y0 = ordf(__reduce_min_sync(0xffffffffu, live ? ford(y) : big));
    o.z0 = ordf(__reduce_min_sync(0xffffffffu, live ? ford(z) : big));
    o.x1 = ordf(__reduce_max_sync(0xffffffffu, live ? ford(x) : -big));
    o.y1 = ordf(__reduce_max_sync(0xffffffffu, live ? ford(y) : -big));
    o.z1 = ordf(__reduce_max_sync(0xffffffffu, live ? ford(z) : -big));
  }
  return o;
}

__global__ void __launch_bounds__(kFirstThreads)
fps_first_pass_kernel(const float* __restrict__ range, const float* __restrict__ lut, const float* __restrict__ ground,
                      int B, int HW, int NB, int runs, float thr, unsigned* __restrict__ temp_ws, FpsBucket* __restrict__ rec,
                      int* __restrict__ work) {
  const int lane = threadIdx.x & 31;
  const int w = blockIdx.x * (kFirstThreads / 32) + (threadIdx.x >> 5);
  const int f = w / runs, c = w - f * runs;
  if (f >= B) return;                                        // warp-uniform
  int nlive = 0;                                             // buckets of this warp's run that can still change
  const int per = (NB + runs - 1) / runs;
  const int b0 = c * per, b1 = min(NB, b0 + per);
  const float* rg = range + (size_t)f * HW;
  unsigned* temp = temp_ws + (size_t)f * HW;
  const float g0 = ground[f * 4], g1 = ground[f * 4 + 1], g2 = ground[f * 4 + 2], g3 = ground[f * 4 + 3];
  const float gnorm = sqrtf(torch_sum3(g0 * g0, g1 * g1, g2 * g2));
  float x1, y1, z1;                                          // seed 0 is flat index 0 (sampling_gpu.cu:44-46)
  masked_point(rg[0], lut, g0, g1, g2, g3, gnorm, thr, x1, y1, z1);
#pragma unroll 2
  for (int b = b0; b < b1; ++b) {
    const FpsBucket o = fps_first_bucket(rg, lut, b, lane, HW, g0, g1, g2, g3, gnorm, thr, x1, y1, z1, temp);
    nlive += o.bmax != 0u ? 1 : 0;
    if (lane == 0) {
      float4* dst = reinterpret_cast<float4*>(rec + (size_t)f * NB + b);
      dst[0] = make_float4(o.x0, o.y0, o.z0, o.x1);
      dst[1] = make_float4(o.y1, o.z1, __uint_as_float(o.bmax), __uint_as_float(o.btk));
    }
  }
  if (work && lane == 0 && nlive) atomicAdd(work + f, nlive);
}

// The same with four consecutive pixels per lane (H*W a multiple of 4): 128-bit loads and stores, a bucket is 8 lanes, a warp
// does 4 buckets at a time and the per-bucket reductions are 3-step shuffles.  The ground mask is decided by a * (1/|g|)
// wherever that is more than 1e-6 (relative) away from the threshold -- the correctly rounded quotient the reference compares
// is then on the same side -- and by the division itself for the lanes that are not.
__device__ __forceinline__ float seg8_min(float v) {
  v = fminf(v, __shfl_xor_sync(0xffffffffu, v, 1)); v = fminf(v, __shfl_xor_sync(0xffffffffu, v, 2));
  return fminf(v, __shfl_xor_sync(0xffffffffu, v, 4));
}
__device__ __forceinline__ float seg8_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1)); v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
  return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 4));
}
__device__ __forceinline__ unsigned seg8_umin(unsigned v) {
  v = min(v, __shfl_xor_sync(0xffffffffu, v, 1)); v = min(v, __shfl_xor_sync(0xffffffffu, v, 2));
  return min(v, __shfl_xor_sync(0xffffffffu, v, 4));
}
__device__ __forceinline__ unsigned seg8_umax(unsigned v) {
  v = max(v, __shfl_xor_sync(0xffffffffu, v, 1)); v = max(v, __shfl_xor_sync(0xffffffffu, v, 2));
  return max(v, __shfl_xor_sync(0xffffffffu, v, 4));
}

__global__ void __launch_bounds__(kFirstThreads)
fps_first_pass4_kernel(const float* __restrict__ range, const float* __restrict__ lut, const float* __restrict__ ground,
                       int B, int HW, int NB, int runs, float thr, unsigned* __restrict__ temp_ws, FpsBucket* __restrict__ rec,
                       int* __restrict__ work) {
  const int lane = threadIdx.x & 31;
  const int w = blockIdx.x * (kFirstThreads / 32) + (threadIdx.x >> 5);
  const int f = w / runs, c = w - f * runs;
  if (f >= B) return;                                        // warp-uniform
  int nlive = 0;                                             // buckets of this warp's run that can still change
  const int NG = (HW + 127) >> 7;                            // groups of 4 buckets
  const int per = (NG + runs - 1) / runs;
  const int ga = c * per, gb = min(NG, ga + per);
  const float* rg = range + (size_t)f * HW;
  uint4* temp4 = reinterpret_cast<uint4*>(temp_ws + (size_t)f * HW);
  const float4* lut4 = reinterpret_cast<const float4*>(lut);
  const float g0 = ground[f * 4], g1 = ground[f * 4 + 1], g2 = ground[f * 4 + 2], g3 = ground[f * 4 + 3];
  const float gnorm = sqrtf(torch_sum3(g0 * g0, g1 * g1, g2 * g2));
  const float rgn = 1.0f / gnorm, thi = thr + fabsf(thr) * 1e-6f, tlo = thr - fabsf(thr) * 1e-6f;
  float sx, sy, sz;                                          // seed 0 is flat index 0 (sampling_gpu.cu:44-46)
  masked_point(rg[0], lut, g0, g1, g2, g3, gnorm, thr, sx, sy, sz);
  const float INF = __int_as_float(0x7f800000), QNAN = __int_as_float(0x7fc00000);
  for (int g = ga; g < gb; ++g) {
    const int p0 = (g << 7) + 4 * lane;
    const bool inb = p0 < HW;                                // H*W % 4 == 0: all four pixels or none
    float4 r = make_float4(0.f, 0.f, 0.f, 0.f), l0 = r, l1 = r, l2 = r;
    if (inb) {
      r = ld_stream_f4(reinterpret_cast<const float4*>(rg + p0));
      const float4* lp = lut4 + (size_t)(p0 >> 2) * 3;
      l0 = __ldg(lp); l1 = __ldg(lp + 1); l2 = __ldg(lp + 2);
    }
    float x[4] = {r.x * l0.x, r.y * l0.w, r.z * l1.z, r.w * l2.y};
    float y[4] = {r.x * l0.y, r.y * l1.x, r.z * l1.w, r.w * l2.z};
    float z[4] = {r.x * l0.z, r.y * l1.y, r.z * l2.x, r.w * l2.w};
    // ---- ground mask (utils/segment_utils.py:137-139)
    float a[4];
    bool keep[4], unsure = false;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      a[i] = fabsf(torch_sum3(x[i] * g0, y[i] * g1, z[i] * g2) + g3);
      const float qa = a[i] * rgn;
      keep[i] = qa > thi;
      unsure = unsure || !(keep[i] || qa < tlo);
    }
    if (__any_sync(0xffffffffu, unsure)) {
#pragma unroll
      for (int i = 0; i < 4; ++i) keep[i] = a[i] / gnorm > thr;
    }
    // ---- round 1: t = min(d(point, seed 0), 1e10)
    float t[4];
    unsigned word[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (!keep[i]) { x[i] = 0.f; y[i] = 0.f; z[i] = 0.f; }
      const bool origin = (x[i] == 0.f) && (y[i] == 0.f) && (z[i] == 0.f);
      t[i] = fminf(fps_dist(x[i], y[i], z[i], sx, sy, sz), 1e10f);
      word[i] = __float_as_uint(t[i]) | (origin ? 0x80000000u : 0u);
    }
    if (inb) temp4[p0 >> 2] = make_uint4(word[0], word[1], word[2], word[3]);
    const unsigned t0 = inb ? __float_as_uint(t[0]) : 0u, t1 = inb ? __float_as_uint(t[1]) : 0u;
    const unsigned t2 = inb ? __float_as_uint(t[2]) : 0u, t3 = inb ? __float_as_uint(t[3]) : 0u;
    const unsigned bmax = seg8_umax(max(max(t0, t1), max(t2, t3)));
    const int b = (g << 2) + (lane >> 3);                    // this lane's bucket
    nlive += __popc(__ballot_sync(0xffffffffu, (lane & 7) == 0 && b < NB && bmax != 0u));
    FpsBucket o = {INF, INF, INF, -INF, -INF, -INF, 0u, kNoTie};
    // buckets whose running distances are all zero stay at the defaults (see fps_first_bucket); bucket 0 is always kept
    if (__any_sync(0xffffffffu, bmax != 0u) || g == 0) {
      // tie key of pixel p0 + i = key(p0) | bitrev(i): the smallest comes first in the order i = 0, 2, 1, 3
      const unsigned key0 = (__brev((unsigned)p0) & 0xFFC00000u) | ((unsigned)p0 >> 10);
      unsigned tk = kNoTie;
      tk = (inb && t3 == bmax) ? (key0 | 0xC0000000u) : tk;
      tk = (inb && t1 == bmax) ? (key0 | 0x80000000u) : tk;
      tk = (inb && t2 == bmax) ? (key0 | 0x40000000u) : tk;
      tk = (inb && t0 == bmax) ? key0 : tk;
      o.bmax = bmax;
      o.btk = seg8_umin(tk);
      // box over the points that can still change (t > 0): the others are NaN, which min / max skip
      float mn[3] = {QNAN, QNAN, QNAN}, mx[3] = {QNAN, QNAN, QNAN};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const bool live = inb && t[i] > 0.f;
        const float lx = live ? x[i] : QNAN, ly = live ? y[i] : QNAN, lz = live ? z[i] : QNAN;
        mn[0] = fminf(mn[0], lx); mn[1] = fminf(mn[1], ly); mn[2] = fminf(mn[2], lz);
        mx[0] = fmaxf(mx[0], lx); mx[1] = fmaxf(mx[1], ly); mx[2] = fmaxf(mx[2], lz);
      }
      o.x0 = seg8_min(mn[0]); o.y0 = seg8_min(mn[1]); o.z0 = seg8_min(mn[2]);
      o.x1 = seg8_max(mx[0]); o.y1 = seg8_max(mx[1]); o.z1 = seg8_max(mx[2]);
      if (!(o.x0 == o.x0)) { o.x0 = INF; o.y0 = INF; o.z0 = INF; o.x1 = -INF; o.y1 = -INF; o.z1 = -INF; }   // no live point
      if (bmax == 0u && b != 0) o.btk = kNoTie;
    }
    if ((lane & 7) == 0 && b < NB) {
      float4* dst = reinterpret_cast<float4*>(rec + (size_t)f * NB + b);
      dst[0] = make_float4(o.x0, o.y0, o.z0, o.x1);
      dst[1] = make_float4(o.y1, o.z1, __uint_as_float(o.bmax), __uint_as_float(o.btk));
    }
  }
  if (work && lane == 0 && nlive) atomicAdd(work + f, nlive);
}

// The round kernels give one frame at a time to a CTA, and a frame's rounds cost in proportion to its buckets that can
// change (30 ... 60 % of them).  Handing the frames out longest first keeps the CTAs' finishing times together: queue[i] =
// the frame with the i-th largest count (ties by index; any order gives the same seeds), followed by `tail` entries B that
// send the CTAs home.  work == nullptr: index order.
constexpr int kOrderMaxFrames = 8192;
__global__ void __launch_bounds__(256)
fps_order_kernel(const int* __restrict__ work, int B, int tail, int* __restrict__ queue) {
  __shared__ int s_w[256];
  const int f = blockIdx.x * 256 + threadIdx.x;
  if (work == nullptr) {                                     // index order
    if (f < B + tail) queue[f] = f < B ? f : B;
    return;
  }
  const int w = f < B ? work[f] : 0;
  int rank = 0;
  for (int g0 = 0; g0 < B; g0 += 256) {                      // every thread of the block walks the same counts: a tile at a time
    __syncthreads();
    s_w[threadIdx.x] = g0 + (int)threadIdx.x < B ? work[g0 + threadIdx.x] : -1;
    __syncthreads();
    const int n = min(256, B - g0);
#pragma unroll 8
    for (int i = 0; i < n; ++i) {
      const int v = s_w[i];
      rank += (v > w || (v == w && g0 + i < f)) ? 1 : 0;
    }
  }
  if (f < B) queue[rank] = f;
  else if (f < B + tail) queue[f] = B;
}

template <int THREADS, int Q, int MINB, bool SPLIT>   // Q buckets per lane: the warp owns buckets warp + NW * (q * 32 + lane)
__global__ void __launch_bounds__(THREADS, MINB)
segment_fps_pruned_kernel(const float* __restrict__ range, const float* __restrict__ lut, const float* __restrict__ ground,
                          int B, int HW, int m, float thr, unsigned* __restrict__ temp_ws, const FpsBucket* __restrict__ rec,
                          int* __restrict__ next_frame, int* __restrict__ center_idx, float* __restrict__ centers) {
  constexpr int NW = THREADS / 32;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int NB = (HW + 31) >> 5;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float4* s_boxa = reinterpret_cast<float4*>(smem_raw);          // [Q][THREADS]: x0, y0, z0, x1 of bucket (q, tid)
  float2* s_boxb = reinterpret_cast<float2*>(s_boxa + Q * THREADS);   // [Q][THREADS]: y1, z1
  __shared__ uint2 s_part[2][32];
  __shared__ float4 s_win[2];
  __shared__ int s_frame;
  unsigned* temp = temp_ws + (size_t)blockIdx.x * HW;      // (SPLIT: per frame, set below)
  const float INF = __int_as_float(0x7f800000);

  // Frames are handed out by a counter: the number of bucket updates (and so the time) differs by tens of per cent
  // from frame to frame, and a static round-robin leaves the SMs with the short frames idle at the end of the launch.
  for (;;) {
    if (tid == 0) s_frame = atomicAdd(next_frame, 1);
    __syncthreads();
    const int f = s_frame;
    if (f >= B) break;
    const float* rg = range + (size_t)f * HW;
    const float g0 = ground[f * 4], g1 = ground[f * 4 + 1], g2 = ground[f * 4 + 2], g3 = ground[f * 4 + 3];
    const float gnorm = sqrtf(torch_sum3(g0 * g0, g1 * g1, g2 * g2));
    // seed 0 is flat index 0 (sampling_gpu.cu:44-46)
    float x1, y1, z1;
    masked_point(rg[0], lut, g0, g1, g2, g3, gnorm, thr, x1, y1, z1);
    const bool seed0_origin = (x1 == 0.f) && (y1 == 0.f) && (z1 == 0.f);
    if (tid == 0) {
      center_idx[(size_t)f * m] = 0;
      float* c = centers + (size_t)f * m * 3;
      c[0] = x1; c[1] = y1; c[2] = z1;
    }
    // ---- round 1 over every bucket: t = min(d, 1e10), boxes, maxima, tie keys
    unsigned bmax[Q], btk[Q];
    if (SPLIT) {                                              // done by fps_first_pass_kernel: fetch this thread's records
      temp = temp_ws + (size_t)f * HW;
#pragma unroll
      for (int q = 0; q < Q; ++q) {
        const int b = warp + NW * (q * 32 + lane);
        float4 ra = make_float4(INF, INF, INF, -INF), rb = make_float4(-INF, -INF, 0.f, __uint_as_float(kNoTie));
        if (b < NB) {
          const float4* src = reinterpret_cast<const float4*>(rec + (size_t)f * NB + b);
          ra = ld_stream_f4(src); rb = ld_stream_f4(src + 1);
        }
        s_boxa[q * THREADS + tid] = ra;
        s_boxb[q * THREADS + tid] = make_float2(rb.x, rb.y);
        bmax[q] = __float_as_uint(rb.z); btk[q] = __float_as_uint(rb.w);
      }
    } else {
#pragma unroll
    for (int q = 0; q < Q; ++q) {
      bmax[q] = 0u; btk[q] = kNoTie;
      float mx0 = INF, my0 = INF, mz0 = INF, mx1 = -INF, my1 = -INF, mz1 = -INF;
#pragma unroll 2
      for (int t = 0; t < 32; ++t) {
        const int b = warp + NW * (q * 32 + t);
        if (b >= NB) break;                                   // warp-uniform
        const int p = (b << 5) + lane;
        const bool inb = p < HW;
        float x = 0.f, y = 0.f, z = 0.f;
        if (inb) masked_point(rg[p], lut + (size_t)p * 3, g0, g1, g2, g3, gnorm, thr, x, y, z);
        const bool origin = (x == 0.f) && (y == 0.f) && (z == 0.f);
        const float tv = fminf(fps_dist(x, y, z, x1, y1, z1), 1e10f);
        if (inb) temp[p] = __float_as_uint(tv) | (origin ? 0x80000000u : 0u);
        const unsigned nb = inb ? __float_as_uint(tv) : 0u;
        const unsigned newmax = __reduce_max_sync(0xffffffffu, nb);
        // A bucket whose running distances are all zero -- ground and empty pixels when seed 0 is one of them, about half
        // of a frame's buckets -- can never change, and can only be chosen when every point of the frame sits at zero, in
        // which case the reference's tie rule picks pixel 0 (tie key 0): bucket 0 is always kept, the others keep the
        // defaults (maximum 0, no tie key, empty box).
        if (newmax == 0u && b != 0) continue;                   // warp-uniform
        const unsigned tk = (inb && nb == newmax) ? ((__brev((unsigned)p) & 0xFFC00000u) | ((unsigned)p >> 10)) : kNoTie;
        const unsigned ntk = __reduce_min_sync(0xffffffffu, tk);
        // box over the points that can still change (t > 0).  Coordinates inside +-250 m are shifted into (0, 512), where
        // the bit patterns order like the values (one REDUX each, no order-preserving transform); the shift rounds to
        // 3e-5 m and the box is widened by 1e-4 m on every side, which keeps the pruning conservative.
        const bool live = inb && tv > 0.f;
        float a0, a1, a2, a3, a4, a5;
        if (!__any_sync(0xffffffffu, live && !(fmaxf(fmaxf(fabsf(x), fabsf(y)), fabsf(z)) < 250.f))) {
          const unsigned ux = __float_as_uint(x + 256.f), uy = __float_as_uint(y + 256.f), uz = __float_as_uint(z + 256.f);
          a0 = __uint_as_float(__reduce_min_sync(0xffffffffu, live ? ux : 0x7f800000u)) - 256.0001f;
          a1 = __uint_as_float(__reduce_min_sync(0xffffffffu, live ? uy : 0x7f800000u)) - 256.0001f;
          a2 = __uint_as_float(__reduce_min_sync(0xffffffffu, live ? uz : 0x7f800000u)) - 256.0001f;
          a3 = __uint_as_float(__reduce_max_sync(0xffffffffu, live ? ux : 0u)) - 255.9999f;
          a4 = __uint_as_float(__reduce_max_sync(0xffffffffu, live ? uy : 0u)) - 255.9999f;
          a5 = __uint_as_float(__reduce_max_sync(0xffffffffu, live ? uz : 0u)) - 255.9999f;
        } else {                                               // coordinates beyond any lidar's range (or NaN): exact order
          const int big = 0x7fffffff;
          a0 = ordf(__reduce_min_sync(0xffffffffu, live ? ford(x) : big));
          a1 = ordf(__reduce_min_sync(0xffffffffu, live ? ford(y) : big));
          a2 = ordf(__reduce_min_sync(0xffffffffu, live ? ford(z) : big));
          a3 = ordf(__reduce_max_sync(0xffffffffu, live ? ford(x) : -big));
          a4 = ordf(__reduce_max_sync(0xffffffffu, live ? ford(y) : -big));
          a5 = ordf(__reduce_max_sync(0xffffffffu, live ? ford(z) : -big));
        }
        if (lane == t) { bmax[q] = newmax; btk[q] = ntk; mx0 = a0; my0 = a1; mz0 = a2; mx1 = a3; my1 = a4; mz1 = a5; }
      }
      // an all-dead / missing bucket keeps (+big, -big): its lower bound is huge and its maximum 0
      s_boxa[q * THREADS + tid] = make_float4(mx0, my0, mz0, mx1);
      s_boxb[q * THREADS + tid] = make_float2(my1, mz1);
    }
    }
    // (boxes are private to their owner thread: no synchronisation needed before they are read back)

    for (int j = 1; j < m; ++j) {
      // ---- winner of the previous round (in round 1: of the first pass)
      {
        unsigned d = 0u, tk = kNoTie;
#pragma unroll
        for (int q = 0; q < Q; ++q) {
          const bool better = bmax[q] > d || (bmax[q] == d && btk[q] < tk);
          d = better ? bmax[q] : d; tk = better ? btk[q] : tk;
        }
        const unsigned dw = __reduce_max_sync(0xffffffffu, d);
        const unsigned tw = __reduce_min_sync(0xffffffffu, d == dw ? tk : kNoTie);
        if (lane == 0) s_part[j & 1][warp] = make_uint2(dw, tw);
        __syncthreads();
#if RPCC_FPS_ONEWINNER
        // warp 0 alone reduces the warps' winners, fetches the centre and publishes it: a second block sync, but 31 warps
        // skip ~30 instructions a round
        if (warp == 0) {
          const uint2 v = lane < NW ? s_part[j & 1][lane] : make_uint2(0u, kNoTie);
          const unsigned dmax = __reduce_max_sync(0xffffffffu, v.x);
          const unsigned tkmin = __reduce_min_sync(0xffffffffu, v.x == dmax ? v.y : kNoTie);
          const int k = (int)(((tkmin & 0x3FFFFFu) << 10) | __brev(tkmin & 0xFFC00000u));
          if (lane == 0) {
            const float r = ld_stream_f(rg + k);
            const float wx = ld_stream_f(lut + (size_t)k * 3), wy = ld_stream_f(lut + (size_t)k * 3 + 1), wz = ld_stream_f(lut + (size_t)k * 3 + 2);
            const bool org = (temp[k] >> 31) != 0u;
            const float cx = org ? 0.f : r * wx, cy = org ? 0.f : r * wy, cz = org ? 0.f : r * wz;
            s_win[j & 1] = make_float4(cx, cy, cz, 0.f);
            center_idx[(size_t)f * m + j] = k;
            float* c = centers + ((size_t)f * m + j) * 3;
            c[0] = cx; c[1] = cy; c[2] = cz;
          }
        }
        __syncthreads();
        { const float4 w = s_win[j & 1]; x1 = w.x; y1 = w.y; z1 = w.z; }
#else
        const uint2 v = lane < NW ? s_part[j & 1][lane] : make_uint2(0u, kNoTie);
        const unsigned dmax = __reduce_max_sync(0xffffffffu, v.x);
        const unsigned tkmin = __reduce_min_sync(0xffffffffu, v.x == dmax ? v.y : kNoTie);
        const int k = (int)(((tkmin & 0x3FFFFFu) << 10) | __brev(tkmin & 0xFFC00000u));
        // all five loads leave together (volatile asm: the compiler would otherwise predicate the last four on the first)
        const float r = ld_stream_f(rg + k);
        const float wx = ld_stream_f(lut + (size_t)k * 3), wy = ld_stream_f(lut + (size_t)k * 3 + 1), wz = ld_stream_f(lut + (size_t)k * 3 + 2);
        const bool org = (temp[k] >> 31) != 0u;
        x1 = org ? 0.f : r * wx; y1 = org ? 0.f : r * wy; z1 = org ? 0.f : r * wz;
        if (tid == 0) {
          center_idx[(size_t)f * m + j] = k;
          float* c = centers + ((size_t)f * m + j) * 3;
          c[0] = x1; c[1] = y1; c[2] = z1;
        }
#endif
      }
      if (j == m - 1) break;                                  // the last centre needs no update pass
      // ---- which of my buckets can change?
      unsigned act[Q];
#pragma unroll
      for (int q = 0; q < Q; ++q) {
        const float4 ba = s_boxa[q * THREADS + tid];            // two vector loads per box (six scalar ones before)
        const float2 bb = s_boxb[q * THREADS + tid];
        const float ox = fmaxf(fmaxf(ba.x - x1, x1 - ba.w), 0.f);
        const float oy = fmaxf(fmaxf(ba.y - y1, y1 - bb.x), 0.f);
        const float oz = fmaxf(fmaxf(ba.z - z1, z1 - bb.y), 0.f);
        const float lb = __fmaf_rn(oz, oz, __fmaf_rn(oy, oy, ox * ox));
        act[q] = __ballot_sync(0xffffffffu, lb * 0.99999f < __uint_as_float(bmax[q]));
      }
      // ---- update them with the reference arithmetic.  When seed 0 is an origin point (a ground or empty pixel 0: the
      //      usual case) every origin point sits at t = 0 for good -- min(d, 0) = 0 whatever coordinates go in -- so
      //      the masked points need not be re-zeroed (SIMPLE); otherwise their origin bit decides.
      auto update = [&](auto simple) {
#pragma unroll
        for (int q = 0; q < Q; ++q) {
          unsigned a = act[q];
          while (a) {
            const int t = __ffs(a) - 1;
            a &= a - 1;
            const int b = warp + NW * (q * 32 + t);
            const int p = (b << 5) + lane;
            const bool inb = p < HW;
            unsigned nb = 0u;
            if (inb) {
              const unsigned tb = temp[p];
              const float r = __ldg(rg + p);
              const float lx = __ldg(lut + (size_t)p * 3), ly = __ldg(lut + (size_t)p * 3 + 1), lz = __ldg(lut + (size_t)p * 3 + 2);
              float x = r * lx, y = r * ly, z = r * lz;
              if (!decltype(simple)::value) {
                const bool org = (tb >> 31) != 0u;
                x = org ? 0.f : x; y = org ? 0.f : y; z = org ? 0.f : z;
              }
              const float told = __uint_as_float(tb & 0x7fffffffu);
              const float tn = fminf(fps_dist(x, y, z, x1, y1, z1), told);          // sampling_gpu.cu:64-66
              nb = __float_as_uint(tn);
              if (tn != told) temp[p] = nb | (tb & 0x80000000u);
            }
            const unsigned newmax = __reduce_max_sync(0xffffffffu, nb);
            const unsigned tk = (inb && nb == newmax) ? ((__brev((unsigned)p) & 0xFFC00000u) | ((unsigned)p >> 10)) : kNoTie;
            const unsigned ntk = __reduce_min_sync(0xffffffffu, tk);
            if (lane == t) { bmax[q] = newmax; btk[q] = ntk; }
          }
        }
      };
      if (seed0_origin) update(std::true_type()); else update(std::false_type());
    }
    __syncthreads();   // the next frame's first pass rewrites temp[], s_part and s_frame
  }
}

static cudaError_t launch_first_pass(const float* range, const float* lut, const float* ground, int B, int HW, int NB, float thr,
                                     unsigned* temp_ws, FpsBucket* rec, cudaStream_t st, int* work = nullptr) {
  const int runs = 32;                                       // warps per frame: ~125 buckets each at 64 x 2000
  const long long warps = (long long)B * runs;
  const int wpb = kFirstThreads / 32;
  const unsigned blocks = (unsigned)((warps + wpb - 1) / wpb);
  const bool vec4 = HW % 4 == 0 && (reinterpret_cast<uintptr_t>(range) | reinterpret_cast<uintptr_t>(lut)) % 16 == 0;
  if (vec4) fps_first_pass4_kernel<<<blocks, kFirstThreads, 0, st>>>(range, lut, ground, B, HW, NB, runs, thr, temp_ws, rec, work);
  else fps_first_pass_kernel<<<blocks, kFirstThreads, 0, st>>>(range, lut, ground, B, HW, NB, runs, thr, temp_ws, rec, work);
  count_launch();
  return cudaGetLastError();
}

__device__ __forceinline__ unsigned fps_tie_key(unsigned p) { return (__brev(p) & 0xFFC00000u) | (p >> 10); }

// ---- the rounds with 64-pixel buckets (the default).  More than half of the instructions of the kernel above, and the
//      longest dependent chain of a round, are the bucket updates: a warp walks its touched buckets one at a time, five
//      32-bit loads and one trip to L2 each.  Here a bucket is one warp wide at two consecutive pixels per lane (64-bit
//      loads): half as many boxes to keep (48 KB per frame at 64 x 2000) and to test, half as many update steps.  The
//      box of a bucket is the union of the two 32-pixel boxes the first pass leaves, its maximum the larger of the two.
//      The tie key of a bucket is no longer kept: a round's winner is found by scanning the one bucket that holds the
//      frame's maximum (the load warp 0 made anyway to fetch the centre's coordinates, 32 lanes wide now); only when
//      several buckets hold the very same maximum -- the origin points when seed 0 is not one of them, or equal
//      distances -- do all warps scan their candidates (TIE below).  Needs H*W even and 8-byte aligned images (every
//      lidar of the reference); the kernel above remains for the rest.
//      Measured on 1184 frames of 64 x 2000 (profiles/r02h_fps_pair_ab.txt, r02j_fps_wide_ab.txt): one 32-pixel bucket
//      per step 2.70 ms, two per step (16 lanes x 2 pixels each) 2.43-2.47 ms, 64-pixel buckets 2.26 ms at two CTAs of
//      1024 threads per SM -- and 2.51 ms at FOUR CTAs of 512 threads (a third fewer instructions, 46 % issue
//      utilisation): what shortens a round is warps working on the same frame, not more frames per SM.
//      Also measured (profiles/r02j_fps_wide_smemcoords_ab.txt): the point holding every bucket's maximum kept in shared
//      memory (first pass and update steps leave it), so that all warps reduce the candidates themselves -- one barrier
//      per round, no global load before the box tests.  Seeds identical, but a third more instructions (the capture in
//      every update step, the redundant reduction) at the same 63 % issue utilisation: 2.57-2.66 ms.  Not kept.
template <int THREADS, int Q, int MINB>
__global__ void __launch_bounds__(THREADS, MINB)
segment_fps_wide_kernel(const float* __restrict__ range, const float* __restrict__ lut, const float* __restrict__ ground,
                        int B, int HW, int m, float thr, unsigned* __restrict__ temp_ws, const FpsBucket* __restrict__ rec,
                        int* __restrict__ next_frame, int* __restrict__ center_idx, float* __restrict__ centers) {
  constexpr int NW = THREADS / 32;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int NB = (HW + 31) >> 5;                             // the first pass's 32-pixel buckets
  const int NB2 = (HW + 63) >> 6;                            // this kernel's
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float4* s_boxa = reinterpret_cast<float4*>(smem_raw);          // [Q][THREADS]: x0, y0, z0, x1 of bucket (q, tid)
  float2* s_boxb = reinterpret_cast<float2*>(s_boxa + Q * THREADS);   // [Q][THREADS]: y1, z1
  __shared__ uint2 s_part[2][32];
  __shared__ float4 s_win[2];
  __shared__ int s_frame;
  const float INF = __int_as_float(0x7f800000);

  for (;;) {
    // queue: next_frame[0] = head, next_frame[64 ...] = the frames, longest first, then one B per CTA (fps_order_kernel)
    if (tid == 0) s_frame = next_frame[64 + atomicAdd(next_frame, 1)];
    __syncthreads();
    const int f = s_frame;
    if (f >= B) break;
    const float* rg = range + (size_t)f * HW;
    unsigned* temp = temp_ws + (size_t)f * HW;
    float sx, sy, sz;                                          // seed 0 is flat index 0 (sampling_gpu.cu:44-46)
    {
      const float g0 = ground[f * 4], g1 = ground[f * 4 + 1], g2 = ground[f * 4 + 2], g3 = ground[f * 4 + 3];
      masked_point(rg[0], lut, g0, g1, g2, g3, sqrtf(torch_sum3(g0 * g0, g1 * g1, g2 * g2)), thr, sx, sy, sz);
    }
    const bool seed0_origin = (sx == 0.f) && (sy == 0.f) && (sz == 0.f);
    if (tid == 0) {
      center_idx[(size_t)f * m] = 0;
      float* c = centers + (size_t)f * m * 3;
      c[0] = sx; c[1] = sy; c[2] = sz;
    }
    // ---- round 1 was done by the first-pass kernel: two of its records make a bucket
    unsigned bmax[Q];
#pragma unroll
    for (int q = 0; q < Q; ++q) {
      const int b = warp + NW * (q * 32 + lane);
      float4 ra = make_float4(INF, INF, INF, -INF), rb = make_float4(-INF, -INF, 0.f, 0.f);
      if (b < NB2) {
        const float4* src = reinterpret_cast<const float4*>(rec + (size_t)f * NB + 2 * b);
        ra = ld_stream_f4(src); rb = ld_stream_f4(src + 1);
        if (2 * b + 1 < NB) {
          const float4 sa = ld_stream_f4(src + 2), sb = ld_stream_f4(src + 3);
          ra = make_float4(fminf(ra.x, sa.x), fminf(ra.y, sa.y), fminf(ra.z, sa.z), fmaxf(ra.w, sa.w));
          rb = make_float4(fmaxf(rb.x, sb.x), fmaxf(rb.y, sb.y), __uint_as_float(max(__float_as_uint(rb.z), __float_as_uint(sb.z))), 0.f);
        }
      }
      s_boxa[q * THREADS + tid] = ra;
      s_boxb[q * THREADS + tid] = make_float2(rb.x, rb.y);
      bmax[q] = __float_as_uint(rb.z);
    }
    float x1 = sx, y1 = sy, z1 = sz;

    for (int j = 1; j < m; ++j) {
      // ---- winner of the previous round: the largest bucket maximum, then the point inside that bucket
      {
        unsigned d = bmax[0];
        int bq = 0, same = 1;                                  // how many of this lane's buckets hold d
#pragma unroll
        for (int q = 1; q < Q; ++q) {
          same = bmax[q] > d ? 1 : same + (bmax[q] == d ? 1 : 0);
          bq = bmax[q] > d ? q : bq;
          d = max(d, bmax[q]);
        }
        const unsigned dw = __reduce_max_sync(0xffffffffu, d);
        const unsigned holders = __reduce_add_sync(0xffffffffu, d == dw ? (unsigned)same : 0u);
        const unsigned code = __reduce_min_sync(0xffffffffu, d == dw ? (unsigned)(bq * 32 + lane) : 0xFFFFu);
        if (lane == 0) s_part[j & 1][warp] = make_uint2(dw, (unsigned)(warp + NW * (int)code) | (holders > 1u ? 0x80000000u : 0u));
        __syncthreads();
        if (warp == 0) {
          const uint2 v = lane < NW ? s_part[j & 1][lane] : make_uint2(0u, 0u);
          const unsigned dmax = __reduce_max_sync(0xffffffffu, v.x);
          const unsigned cand = __ballot_sync(0xffffffffu, lane < NW && v.x == dmax);
          const bool several = (cand & (cand - 1u)) != 0u || __any_sync(0xffffffffu, v.x == dmax && (v.y >> 31) != 0u);
          if (dmax == 0u) {
            // every running distance is zero: the tie rule picks pixel 0, which is seed 0 again
            if (lane == 0) {
              s_win[j & 1] = make_float4(sx, sy, sz, 0.f);
              center_idx[(size_t)f * m + j] = 0;
              float* c = centers + ((size_t)f * m + j) * 3;
              c[0] = sx; c[1] = sy; c[2] = sz;
            }
          } else if (!several) {
            const int bstar = (int)(__shfl_sync(0xffffffffu, v.y, __ffs(cand) - 1) & 0x7fffffffu);
            const int p = (bstar << 6) + (lane << 1);
            uint2 tb = make_uint2(0u, 0u);
            float2 r = make_float2(0.f, 0.f), la = r, lb2 = r, lc = r;
            if (p < HW) {                                      // H*W is even: both pixels or none
              tb = *reinterpret_cast<const uint2*>(temp + p);
              r = __ldg(reinterpret_cast<const float2*>(rg + p));
              const float2* l2 = reinterpret_cast<const float2*>(lut + (size_t)p * 3);
              la = __ldg(l2); lb2 = __ldg(l2 + 1); lc = __ldg(l2 + 2);
            }
            const unsigned ka = (p < HW && (tb.x & 0x7fffffffu) == dmax) ? fps_tie_key((unsigned)p) : kNoTie;
            const unsigned kb = (p < HW && (tb.y & 0x7fffffffu) == dmax) ? fps_tie_key((unsigned)p + 1u) : kNoTie;
            const unsigned tk = min(ka, kb);
            const unsigned tkmin = __reduce_min_sync(0xffffffffu, tk);
            if (tk == tkmin) {                                 // tie keys are unique: one lane (the bucket holds dmax)
              const bool second = kb < ka;
              const bool org = ((second ? tb.y : tb.x) >> 31) != 0u;
              const float rr = second ? r.y : r.x;
              const float wx = second ? lb2.y : la.x, wy = second ? lc.x : la.y, wz = second ? lc.y : lb2.x;
              const float cx = org ? 0.f : rr * wx, cy = org ? 0.f : rr * wy, cz = org ? 0.f : rr * wz;
              s_win[j & 1] = make_float4(cx, cy, cz, 0.f);
              center_idx[(size_t)f * m + j] = p + (second ? 1 : 0);
              float* c = centers + ((size_t)f * m + j) * 3;
              c[0] = cx; c[1] = cy; c[2] = cz;
            }
          } else if (lane == 0) {
            s_win[j & 1] = make_float4(__uint_as_float(dmax), 0.f, 0.f, 1.f);    // TIE: resolved by all warps below
          }
        }
        __syncthreads();
        float4 w = s_win[j & 1];
        if (w.w != 0.f) {                                      // block-uniform
          // TIE: every warp scans its buckets that hold the maximum for the smallest tie key
          const unsigned dmax = __float_as_uint(w.x);
          unsigned tkw = kNoTie;
#pragma unroll
          for (int q = 0; q < Q; ++q) {
            unsigned cm = __ballot_sync(0xffffffffu, bmax[q] == dmax);
            while (cm) {
              const int t = __ffs(cm) - 1;
              cm &= cm - 1;
              const int p = ((warp + NW * (q * 32 + t)) << 6) + (lane << 1);
              unsigned tk = kNoTie;
              if (p < HW) {
                const uint2 tb = *reinterpret_cast<const uint2*>(temp + p);
                if ((tb.x & 0x7fffffffu) == dmax) tk = fps_tie_key((unsigned)p);
                if ((tb.y & 0x7fffffffu) == dmax) tk = min(tk, fps_tie_key((unsigned)p + 1u));
              }
              tkw = min(tkw, __reduce_min_sync(0xffffffffu, tk));
            }
          }
          if (lane == 0) s_part[j & 1][warp] = make_uint2(tkw, 0u);      // (warp 0 read the maxima before the last barrier)
          __syncthreads();
          if (warp == 0) {
            const unsigned tkmin = __reduce_min_sync(0xffffffffu, lane < NW ? s_part[j & 1][lane].x : kNoTie);
            if (lane == 0) {
              const int k = (int)(((tkmin & 0x3FFFFFu) << 10) | __brev(tkmin & 0xFFC00000u));
              const float r = ld_stream_f(rg + k);
              const float wx = ld_stream_f(lut + (size_t)k * 3), wy = ld_stream_f(lut + (size_t)k * 3 + 1), wz = ld_stream_f(lut + (size_t)k * 3 + 2);
              const bool org = (temp[k] >> 31) != 0u;
              const float cx = org ? 0.f : r * wx, cy = org ? 0.f : r * wy, cz = org ? 0.f : r * wz;
              s_win[j & 1] = make_float4(cx, cy, cz, 0.f);
              center_idx[(size_t)f * m + j] = k;
              float* c = centers + ((size_t)f * m + j) * 3;
              c[0] = cx; c[1] = cy; c[2] = cz;
            }
          }
          __syncthreads();
          w = s_win[j & 1];
        }
        x1 = w.x; y1 = w.y; z1 = w.z;
      }
      if (j == m - 1) break;                                  // the last centre needs no update pass
      // ---- per q: which of my buckets can change, then update them with the reference arithmetic.  When seed 0 is an
      //      origin point (a ground or empty pixel 0: the usual case) every origin point sits at t = 0 for good --
      //      min(d, 0) = 0 whatever coordinates go in -- so the masked points need not be re-zeroed.
      auto update = [&](auto simple) {
#pragma unroll
        for (int q = 0; q < Q; ++q) {
          unsigned act;
          {
            const float4 ba = s_boxa[q * THREADS + tid];
            const float2 bb = s_boxb[q * THREADS + tid];
            const float ox = fmaxf(fmaxf(ba.x - x1, x1 - ba.w), 0.f);
            const float oy = fmaxf(fmaxf(ba.y - y1, y1 - bb.x), 0.f);
            const float oz = fmaxf(fmaxf(ba.z - z1, z1 - bb.y), 0.f);
            const float lb = __fmaf_rn(oz, oz, __fmaf_rn(oy, oy, ox * ox));
            act = __ballot_sync(0xffffffffu, lb * 0.99999f < __uint_as_float(bmax[q]));
          }
          while (act) {                                        // warp-uniform
            const int t = __ffs(act) - 1;
            act &= act - 1;
            const int p = ((warp + NW * (q * 32 + t)) << 6) + (lane << 1);
            unsigned nb = 0u;
            if (p < HW) {                                      // H*W is even: both pixels or none
              const uint2 told = *reinterpret_cast<const uint2*>(temp + p);
              const float2 r = __ldg(reinterpret_cast<const float2*>(rg + p));
              const float2* l2 = reinterpret_cast<const float2*>(lut + (size_t)p * 3);
              const float2 la = __ldg(l2), lb2 = __ldg(l2 + 1), lc = __ldg(l2 + 2);
              float xa = r.x * la.x, ya = r.x * la.y, za = r.x * lb2.x;
              float xb = r.y * lb2.y, yb = r.y * lc.x, zb = r.y * lc.y;
              if (!decltype(simple)::value) {
                const bool oa = (told.x >> 31) != 0u, ob = (told.y >> 31) != 0u;
                xa = oa ? 0.f : xa; ya = oa ? 0.f : ya; za = oa ? 0.f : za;
                xb = ob ? 0.f : xb; yb = ob ? 0.f : yb; zb = ob ? 0.f : zb;
              }
              const float oa_ = __uint_as_float(told.x & 0x7fffffffu), ob_ = __uint_as_float(told.y & 0x7fffffffu);
              const float na = fminf(fps_dist(xa, ya, za, x1, y1, z1), oa_);          // sampling_gpu.cu:64-66
              const float nb_ = fminf(fps_dist(xb, yb, zb, x1, y1, z1), ob_);
              if (na != oa_ || nb_ != ob_)
                *reinterpret_cast<uint2*>(temp + p) = make_uint2(__float_as_uint(na) | (told.x & 0x80000000u),
                                                                 __float_as_uint(nb_) | (told.y & 0x80000000u));
              nb = max(__float_as_uint(na), __float_as_uint(nb_));
            }
            const unsigned newmax = __reduce_max_sync(0xffffffffu, nb);
            bmax[q] = lane == t ? newmax : bmax[q];
          }
        }
      };
      if (seed0_origin) update(std::true_type()); else update(std::false_type());
    }
    __syncthreads();   // the next frame rewrites the boxes, s_part and s_frame
  }
}

template <int THREADS, int Q, int MINB>
static int launch_fps_wide(const float* range, const float* lut, const float* ground, int B, int HW, int m, float thr,
                           int* center_idx, float* centers, cudaStream_t st) {
  auto kern = segment_fps_wide_kernel<THREADS, Q, MINB>;
  const size_t smem = sizeof(float) * 6 * Q * THREADS;
  RPCC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  RPCC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, THREADS, smem));
  if (per_sm < 1) per_sm = 1;
  int grid = per_sm * sm_count();
  if (grid > B) grid = B;
  const int NB = (HW + 31) >> 5;
  void* ws = nullptr;                                        // running distances + bucket records + frame queue head
  cudaMemPool_t pool = nullptr;
  { const int rc = scratch_pool(&pool); if (rc != RPCC_OK) return rc; }
  const size_t temp_bytes = sizeof(unsigned) * (size_t)B * HW;
  const size_t rec_bytes = sizeof(FpsBucket) * (size_t)B * NB;
  const size_t rec_off = (temp_bytes + 255) & ~(size_t)255;
  const size_t ctr_off = rec_off + ((rec_bytes + 255) & ~(size_t)255);
  static const bool order_ok = !(getenv("RPCC_FPS_ORDER") && atoi(getenv("RPCC_FPS_ORDER")) == 0);   // 0: frames in index order (A/B)
  const bool ordered = order_ok && B > grid && B <= kOrderMaxFrames;      // more frames than CTAs: longest first
  RPCC_CUDA(cudaMallocFromPoolAsync(&ws, ctr_off + 256 + sizeof(int) * ((size_t)2 * B + grid), pool, st));
  unsigned char* base = static_cast<unsigned char*>(ws);
  FpsBucket* rec = reinterpret_cast<FpsBucket*>(base + rec_off);
  int* counter = reinterpret_cast<int*>(base + ctr_off);     // queue head (256 bytes), the queue (B + grid entries), live buckets per frame (B)
  int* queue = counter + 64;
  int* work = queue + B + grid;
  cudaError_t le = cudaMemsetAsync(counter, 0, 256, st);
  if (le == cudaSuccess && ordered) le = cudaMemsetAsync(work, 0, sizeof(int) * (size_t)B, st);
  if (le == cudaSuccess) le = launch_first_pass(range, lut, ground, B, HW, NB, thr, static_cast<unsigned*>(ws), rec, st, ordered ? work : nullptr);
  if (le == cudaSuccess) {
    fps_order_kernel<<<(B + grid + 255) / 256, 256, 0, st>>>(ordered ? work : nullptr, B, grid, queue);
    le = cudaGetLastError();
    count_launch();
  }
  if (le == cudaSuccess) {
    kern<<<grid, THREADS, smem, st>>>(range, lut, ground, B, HW, m, thr, static_cast<unsigned*>(ws), rec, counter, center_idx, centers);
    le = cudaGetLastError();
  }
  RPCC_CUDA(cudaFreeAsync(ws, st));
  RPCC_CUDA(le);
  count_launch();
  return RPCC_OK;
}

template <int THREADS, int Q, int MINB, bool SPLIT>
static int launch_fps_pruned(const float* range, const float* lut, const float* ground, int B, int HW, int m, float thr,
                             int* center_idx, float* centers, cudaStream_t st) {
  auto kern = segment_fps_pruned_kernel<THREADS, Q, MINB, SPLIT>;
  const size_t smem = sizeof(float) * 6 * Q * THREADS;
  RPCC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  RPCC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, THREADS, smem));
  if (per_sm < 1) per_sm = 1;
  int grid = per_sm * sm_count();
  if (grid > B) grid = B;
  const int NB = (HW + 31) >> 5;
  // running distances (SPLIT: of every frame, else of the frames in flight) and the buckets' records: stream-ordered
  // scratch, no synchronisation
  void* ws = nullptr;
  cudaMemPool_t pool = nullptr;
  { const int rc = scratch_pool(&pool); if (rc != RPCC_OK) return rc; }
  const size_t temp_bytes = sizeof(unsigned) * (size_t)(SPLIT ? B : grid) * HW;
  const size_t rec_bytes = SPLIT ? sizeof(FpsBucket) * (size_t)B * NB : 0;
  const size_t rec_off = (temp_bytes + 255) & ~(size_t)255;
  const size_t ctr_off = rec_off + ((rec_bytes + 255) & ~(size_t)255);
  RPCC_CUDA(cudaMallocFromPoolAsync(&ws, ctr_off + 256, pool, st));
  unsigned char* base = static_cast<unsigned char*>(ws);
  FpsBucket* rec = reinterpret_cast<FpsBucket*>(base + rec_off);
  int* counter = reinterpret_cast<int*>(base + ctr_off);     // frame queue head
  cudaError_t le = cudaMemsetAsync(counter, 0, sizeof(int), st);
  if (le == cudaSuccess && SPLIT) le = launch_first_pass(range, lut, ground, B, HW, NB, thr, static_cast<unsigned*>(ws), rec, st);
  if (le == cudaSuccess) {
    kern<<<grid, THREADS, smem, st>>>(range, lut, ground, B, HW, m, thr, static_cast<unsigned*>(ws), rec, counter, center_idx, centers);
    le = cudaGetLastError();
  }
  RPCC_CUDA(cudaFreeAsync(ws, st));
  RPCC_CUDA(le);
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
  static const bool use_cluster = getenv("RPCC_FPS_IMPL") && !strcmp(getenv("RPCC_FPS_IMPL"), "cluster");
  static const int fps_threads = getenv("RPCC_FPS_THREADS") ? atoi(getenv("RPCC_FPS_THREADS")) : 1024;
  if (!use_cluster && m >= 2) {
    cudaStream_t st = as_stream(stream);
    const int NB = (HW + 31) / 32;
    static const int minb = getenv("RPCC_FPS_MINB") ? atoi(getenv("RPCC_FPS_MINB")) : 2;
    static const bool wide_ok = !(getenv("RPCC_FPS_WIDE") && atoi(getenv("RPCC_FPS_WIDE")) == 0);   // 0: the 32-pixel kernel (A/B, tests)
    const int NB2 = (HW + 63) / 64;
    if (wide_ok && NB2 <= 2048 && HW % 2 == 0 && (reinterpret_cast<uintptr_t>(range) | reinterpret_cast<uintptr_t>(lut)) % 8 == 0) {
      if (fps_threads == 512) {                              // A/B: four CTAs of 512 threads per SM
        const int q = (NB2 + 511) / 512;
        if (q <= 1) return launch_fps_wide<512, 1, 4>(range, lut, ground, B, HW, m, ground_thr, center_idx, centers, st);
        if (q <= 2) return launch_fps_wide<512, 2, 4>(range, lut, ground, B, HW, m, ground_thr, center_idx, centers, st);
        return launch_fps_wide<512, 4, 4>(range, lut, ground, B, HW, m, ground_thr, center_idx, centers, st);
      }
      if (NB2 <= 1024) return launch_fps_wide<1024, 1, 2>(range, lut, ground, B, HW, m, ground_thr, center_idx, centers, st);
      return launch_fps_wide<1024, 2, 2>(range, lut, ground, B, HW, m, ground_thr, center_idx, centers, st);
    }
    static const bool split = !(getenv("RPCC_FPS_SPLIT") && atoi(getenv("RPCC_FPS_SPLIT")) == 0);
#define RPCC_FPS_GO(T, Q, M) return split ? launch_fps_pruned<T, Q, M, true>(range, lut, ground, B, HW, m, ground_thr, center_idx, centers, st) \
                                          : launch_fps_pruned<T, Q, M, false>(range, lut, ground, B, HW, m, ground_thr, center_idx, centers, st)
    if (fps_threads == 512) {
      const int q = (NB + 511) / 512;
      if (q <= 2) RPCC_FPS_GO(512, 2, 2);
      if (q <= 4) RPCC_FPS_GO(512, 4, 2);
      if (q <= 8) RPCC_FPS_GO(512, 8, 2);
    } else if (minb == 2) {
      const int q = (NB + 1023) / 1024;
      if (q <= 1) RPCC_FPS_GO(1024, 1, 2);
      if (q <= 2) RPCC_FPS_GO(1024, 2, 2);
      if (q <= 4) RPCC_FPS_GO(1024, 4, 2);
    } else {
      const int q = (NB + 1023) / 1024;
      if (q <= 1) RPCC_FPS_GO(1024, 1, 1);
      if (q <= 2) RPCC_FPS_GO(1024, 2, 1);
      if (q <= 4) RPCC_FPS_GO(1024, 4, 1);
    }
#undef RPCC_FPS_GO
    // larger images fall through to the cluster kernel (which has its own capacity check)
  }
  const int J = (HW + kFpsThreads - 1) / kFpsThreads;
  const int need = (J + kCl - 1) / kCl;  // slots per thread
  cudaStream_t st = as_stream(stream);
  if (need <= 4) return launch_segment_fps<4>(range, lut, ground, B, HW, m, ground_thr, center_idx, centers, st);
  if (need <= 10) return launch_segment_fps<10>(range, lut, ground, B, HW, m, ground_thr, center_idx, centers, st);
  if (need <= 16) return launch_segment_fps<16>(range, lut, ground, B, HW, m, ground_thr, center_idx, centers, st);
  set_error("rpcc_segment_fps_batch: H*W = %d exceeds the on-chip capacity (131072 pixels)", HW);
  return RPCC_ERR_CAPACITY;
}
