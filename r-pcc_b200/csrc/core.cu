// core.cu -- error reporting, launch accounting and the one-time host-side LUT.
#include <math.h>
#include <stdarg.h>
#include <atomic>
#include <mutex>

#include "common.cuh"

namespace rpcc {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return RPCC_OK;
  set_error("CUDA error in %s: %s", what, cudaGetErrorString(e));
  return e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver ? RPCC_ERR_NO_DEVICE : RPCC_ERR_CUDA;
}

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

// Stream-ordered scratch comes from a pool owned by the library, one per device, that keeps its memory
// across synchronisation points (release threshold = max).  The device's default pool hands everything
// back to the driver at every event/stream synchronisation -- which the pipelined host path does once
// per chunk -- and re-mapping ~100 MB per launch stalls the copy queue.
int scratch_pool(cudaMemPool_t* out) {
  static std::mutex mu;
  static cudaMemPool_t pools[64] = {nullptr};
  int dev = 0;
  RPCC_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) { set_error("scratch_pool: device ordinal %d out of range", dev); return RPCC_ERR_ARG; }
  std::lock_guard<std::mutex> lock(mu);
  if (!pools[dev]) {
    cudaMemPoolProps props = {};
    props.allocType = cudaMemAllocationTypePinned;
    props.handleTypes = cudaMemHandleTypeNone;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id = dev;
    cudaMemPool_t pool = nullptr;
    RPCC_CUDA(cudaMemPoolCreate(&pool, &props));
    unsigned long long keep = ~0ull;
    RPCC_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
    // a block freed on one stream is reused on another only once that free has completed: never by making
    // the second stream wait for the first (the encoder's stream slots must stay independent)
    int off = 0;
    RPCC_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolReuseAllowInternalDependencies, &off));
    pools[dev] = pool;
  }
  *out = pools[dev];
  return RPCC_OK;
}

}  // namespace rpcc

extern "C" const char* rpcc_last_error(void) { return rpcc::g_err; }
extern "C" int rpcc_version(void) { return 100; }
extern "C" long long rpcc_launch_count(void) { return rpcc::g_launches.load(); }

extern "C" int rpcc_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

// dataset/transformer.py:41-54: double-precision trig through libm (what Python's math module
// calls), narrowed to f32 at the end.  One-time, host side.
extern "C" int rpcc_transform_map(int H, int W, double hfov, double vmax, double vmin, float* lut) {
  RPCC_REQUIRE(lut != nullptr && H >= 2 && W >= 1, "bad argument");
  const double vfov = vmax - vmin;
  for (int h = 0; h < H; ++h) {
    const double alt = vfov * ((double)h / (double)(H - 1)) + vmin;
    const double ca = cos(alt), sa = sin(alt);
    for (int w = 0; w < W; ++w) {
      const double az = hfov * ((double)w / (double)W);
      float* o = lut + ((size_t)h * W + w) * 3;
      o[0] = (float)(ca * cos(az));
      o[1] = (float)(ca * sin(az));
      o[2] = (float)sa;
    }
  }
  return RPCC_OK;
}
