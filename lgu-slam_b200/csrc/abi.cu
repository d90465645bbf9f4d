// abi.cu -- library identification + error plumbing of the C ABI (include/lgu_corr.h).
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <unordered_map>
#include <vector>
#define LGU_STR2(x) #x
#define LGU_STR(x) LGU_STR2(x)
#include "common.cuh"

namespace lgu {
static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return LGU_ERR_LAUNCH;
  }
  return LGU_OK;
}

// ---- one-time launch plumbing ------------------------------------------------------------------------------------
static std::mutex g_plumb_mutex;

int optin_smem(const void* kernel, int bytes, const char* who) {
  struct Done { int dev; const void* kernel; int bytes; };
  static std::vector<Done> done;                                // a few dozen (device, kernel) pairs at most
  int dev = 0;
  cudaGetDevice(&dev);
  {
    std::lock_guard<std::mutex> lk(g_plumb_mutex);
    for (const auto& d : done)
      if (d.dev == dev && d.kernel == kernel && d.bytes >= bytes) return LGU_OK;
  }
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) {
    set_error("%s: cannot opt in to %d B of shared memory: %s", who, bytes, cudaGetErrorString(e));
    return LGU_ERR_LAUNCH;
  }
  std::lock_guard<std::mutex> lk(g_plumb_mutex);
  for (auto& d : done)
    if (d.dev == dev && d.kernel == kernel) { d.bytes = bytes; return LGU_OK; }
  done.push_back({dev, kernel, bytes});
  return LGU_OK;
}

namespace {
struct KeyHash {
  size_t operator()(const MapKey& k) const {
    uint64_t h = 1469598103934665603ull;
    for (uint64_t x : k.v) h = (h ^ x) * 1099511628211ull;
    return (size_t)h;
  }
};
struct KeyEq {
  bool operator()(const MapKey& a, const MapKey& b) const { return memcmp(a.v, b.v, sizeof(a.v)) == 0; }
};
struct Map128 {
  unsigned char b[128];
};
constexpr size_t kMapCacheMax = 512;
std::unordered_map<MapKey, Map128, KeyHash, KeyEq>& map_cache() {
  static std::unordered_map<MapKey, Map128, KeyHash, KeyEq> c;
  return c;
}
}  // namespace

bool map_cache_get(const MapKey& key, void* map128) {
  std::lock_guard<std::mutex> lk(g_plumb_mutex);
  auto it = map_cache().find(key);
  if (it == map_cache().end()) return false;
  memcpy(map128, it->second.b, 128);
  return true;
}
void map_cache_put(const MapKey& key, const void* map128) {
  std::lock_guard<std::mutex> lk(g_plumb_mutex);
  if (map_cache().size() >= kMapCacheMax) map_cache().clear();   // geometry churn: start over (maps are cheap to redo)
  Map128 m;
  memcpy(m.b, map128, 128);
  map_cache()[key] = m;
}
}  // namespace lgu

extern "C" {
// ---- peer-visible device memory (the gathered output buffer of the sharded backend, lgu-slam_b200/sharded.py) ---------
int lgu_peer_alloc(long long bytes, void** ptr, void* handle64) {
  LGU_REQUIRE(bytes > 0 && ptr != nullptr && handle64 != nullptr, "lgu_peer_alloc: bad arguments");
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, (size_t)bytes);                 // a plain cudaMalloc allocation: exportable by CUDA IPC
  if (e != cudaSuccess) {
    lgu::set_error("lgu_peer_alloc: cudaMalloc(%lld) failed: %s", bytes, cudaGetErrorString(e));
    return LGU_ERR_LAUNCH;
  }
  cudaIpcMemHandle_t h;
  e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    lgu::set_error("lgu_peer_alloc: cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e));
    return LGU_ERR_LAUNCH;
  }
  static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
  memcpy(handle64, &h, 64);
  *ptr = p;
  return LGU_OK;
}
int lgu_peer_open(const void* handle64, void** ptr) {
  LGU_REQUIRE(handle64 != nullptr && ptr != nullptr, "lgu_peer_open: bad arguments");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  // opened with the CALLER's device current: the mapping (and peer access to the owning GPU) belongs to the device whose
  // kernels will store through it
  cudaError_t e = cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) {
    lgu::set_error("lgu_peer_open: cudaIpcOpenMemHandle failed: %s", cudaGetErrorString(e));
    return LGU_ERR_LAUNCH;
  }
  return LGU_OK;
}
int lgu_peer_close(void* ptr) {
  cudaError_t e = cudaIpcCloseMemHandle(ptr);
  if (e != cudaSuccess) {
    lgu::set_error("lgu_peer_close: %s", cudaGetErrorString(e));
    return LGU_ERR_LAUNCH;
  }
  return LGU_OK;
}
int lgu_peer_free(void* ptr) {
  cudaError_t e = cudaFree(ptr);
  if (e != cudaSuccess) {
    lgu::set_error("lgu_peer_free: %s", cudaGetErrorString(e));
    return LGU_ERR_LAUNCH;
  }
  return LGU_OK;
}

int lgu_abi_version(void) { return 1; }
const char* lgu_build_info(void) {
  return "lgu_corr sm_100a, nvcc " LGU_STR(__CUDACC_VER_MAJOR__) "." LGU_STR(__CUDACC_VER_MINOR__) ", built " __DATE__;
}
const char* lgu_last_error_string(void) { return lgu::g_err; }
}
