// Runtime shim: device/stream ownership, buffers, error state, residency classification.
// This is the part of the C ABI Jalla's Haskell side would bind for a GPU-resident
// CMatrix instance (replacing mallocForeignPtrArray, /root/reference/Numeric/Jalla/Matrix.hs:785-787).
#include "jb_common.cuh"
#include <cstdarg>
#include <mutex>
#include <map>
#include <string>

namespace jb {

std::atomic<uint64_t> g_launches{0};
static std::mutex g_mu;
static std::map<std::string, int> g_options;
static std::atomic<int> g_device{-1};
static std::atomic<int> g_sms{0};

ThreadState &tls() {
  static thread_local ThreadState st;
  return st;
}

void set_error(int code, const char *fmt, ...) {
  ThreadState &t = tls();
  t.err = code;
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(t.msg, sizeof(t.msg), fmt, ap);
  va_end(ap);
}

static std::atomic<unsigned> g_opt_version{0};
// Hot paths ask for options on every kernel launch (hundreds per blocked factorisation): the mutex + string-map lookup
// happens once per thread and key until jb_set_option changes something.
int option(const char *key) {
  struct Slot { const char *key; unsigned ver; int val; };
  static thread_local Slot cache[32];
  const unsigned ver = g_opt_version.load(std::memory_order_acquire) + 1;     // 0 = empty slot
  Slot &sl = cache[((uintptr_t)key >> 3) & 31];
  if (sl.key == key && sl.ver == ver) return sl.val;
  int v;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_options.find(key);
    v = it == g_options.end() ? -1 : it->second;
  }
  sl.key = key; sl.ver = ver; sl.val = v;
  return v;
}

int ensure_device() {
  int dev = g_device.load();
  if (dev < 0) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
      set_error(JB_ERR_CUDA, "no usable CUDA device (%s); jalla_b200 has no CPU path",
                e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
      cudaGetLastError();
      return JB_ERR_CUDA;
    }
    int cur = 0;
    JB_CUDA(cudaGetDevice(&cur));
    cudaDeviceProp p;
    JB_CUDA(cudaGetDeviceProperties(&p, cur));
    if (p.major != 10) {
      set_error(JB_ERR_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only", cur, p.major, p.minor);
      return JB_ERR_CUDA;
    }
    g_sms.store(p.multiProcessorCount);
    g_device.store(cur);
    dev = cur;
  }
  // `ccall unsafe` can arrive on any OS thread: bind the device per call (H11).
  if (tls().device >= 0) dev = tls().device;
  int cur = -1;
  if (cudaGetDevice(&cur) != cudaSuccess || cur != dev) JB_CUDA(cudaSetDevice(dev));
  return 0;
}

int num_sms() { return g_sms.load() > 0 ? g_sms.load() : 148; }

Residency classify(const void *p) {
  if (!p) return Residency::Device;
  cudaPointerAttributes at;
  cudaError_t e = cudaPointerGetAttributes(&at, p);
  if (e != cudaSuccess) { cudaGetLastError(); return Residency::Host; }
  if (at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged) return Residency::Device;
  return Residency::Host;  // unregistered (pageable) or pinned host memory: staged
}

int ws_alloc(void **p, size_t bytes, cudaStream_t s) {
  if (bytes == 0) bytes = 16;
  cudaError_t e = cudaMallocAsync(p, bytes, s);
  if (e != cudaSuccess) {
    cudaGetLastError();
    set_error(JB_ERR_ALLOC, "workspace allocation of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
    return JB_ERR_ALLOC;
  }
  return 0;
}
void ws_free(void *p, cudaStream_t s) { if (p) cudaFreeAsync(p, s); }

int Staged::open(const void *p, long r, long c, int ld_, size_t esz_, bool copy_in, cudaStream_t s) {
  rows = r; cols = c; esz = esz_;
  if (r <= 0 || c <= 0 || classify(p) == Residency::Device) {
    dev = const_cast<void *>(p); ld = ld_; staged = false;
    return 0;
  }
  staged = true; host = const_cast<void *>(p); host_ld = ld_; ld = (int)(r > 1 ? r : 1);
  if (r & 1) ld = (int)r + 1;   // even leading dimension keeps 16-byte alignment for the TMA paths
  int rc = ws_alloc(&dev, (size_t)ld * (size_t)c * esz, s);
  if (rc) return rc;
  if (copy_in)
    if (int r2 = stage.h2d(dev, (size_t)ld * esz, host, (size_t)host_ld * esz, (size_t)r * esz, (size_t)c, s)) { ws_free(dev, s); staged = false; return r2; }
  return 0;
}

int Staged::close(bool copy_out, cudaStream_t s) {
  if (!staged) return 0;
  int rc = 0;
  if (copy_out) rc = stage.d2h(host, (size_t)host_ld * esz, dev, (size_t)ld * esz, (size_t)rows * esz, (size_t)cols, s);
  ws_free(dev, s);
  if (int r2 = stage.finish()) rc = rc ? rc : r2;      // pageable destination: returns when the drain thread has delivered it
  staged = false;
  return rc;
}

}  // namespace jb

using namespace jb;

extern "C" {

int jb_init(int device) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    cudaGetLastError();
    set_error(JB_ERR_CUDA, "no usable CUDA device; jalla_b200 has no CPU path");
    return JB_ERR_CUDA;
  }
  if (device >= 0) {
    if (device >= n) { set_error(JB_ERR_ARG, "device %d out of range (%d devices)", device, n); return JB_ERR_ARG; }
    JB_CUDA(cudaSetDevice(device));
    g_device.store(-1);
  }
  int rc = ensure_device();
  if (rc) return rc;
  // keep freed workspace memory cached in the stream-ordered pool
  cudaMemPool_t pool;
  int dev = 0;
  cudaGetDevice(&dev);
  if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
    uint64_t thr = UINT64_MAX;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
  }
  return 0;
}

void jb_shutdown(void) { cudaDeviceSynchronize(); }

int jb_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

int jb_last_error(void) { int e = tls().err; tls().err = 0; return e; }
const char *jb_last_error_string(void) { return tls().msg; }
const char *jb_version(void) { return "jalla_b200 0.1 (sm_100a; DMMA/FFMA hand-written kernels; no cuBLAS/cuSOLVER)"; }

void *jb_malloc(size_t bytes, int kind) {
  if (kind != JB_MEM_PINNED && ensure_device()) return nullptr;
  void *p = nullptr;
  if (bytes == 0) bytes = 16;
  cudaError_t e;
  if (kind == JB_MEM_DEVICE) e = cudaMalloc(&p, bytes);
  else if (kind == JB_MEM_MANAGED) e = cudaMallocManaged(&p, bytes);
  else if (kind == JB_MEM_PINNED) e = cudaHostAlloc(&p, bytes, cudaHostAllocPortable);
  else { set_error(JB_ERR_ARG, "jb_malloc: unknown kind %d", kind); return nullptr; }
  if (e != cudaSuccess) { cudaGetLastError(); set_error(JB_ERR_ALLOC, "jb_malloc(%zu,%d): %s", bytes, kind, cudaGetErrorString(e)); return nullptr; }
  return p;
}

// Context-agnostic and thread-safe: GHC runs ForeignPtr finalizers on arbitrary threads (H11).
void jb_free(void *p) {
  if (!p) return;
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return; }
  if (at.type == cudaMemoryTypeHost) cudaFreeHost(p);
  else if (at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged) cudaFree(p);
}

int jb_memcpy_h2d(void *dst, const void *src, size_t bytes) {
  if (int rc = ensure_device()) return rc;
  JB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, tls().stream));
  JB_CUDA(cudaStreamSynchronize(tls().stream));
  return 0;
}
int jb_memcpy_d2h(void *dst, const void *src, size_t bytes) {
  if (int rc = ensure_device()) return rc;
  JB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, tls().stream));
  JB_CUDA(cudaStreamSynchronize(tls().stream));
  return 0;
}
int jb_memcpy_d2d(void *dst, const void *src, size_t bytes) {
  if (int rc = ensure_device()) return rc;
  JB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, tls().stream));
  return 0;
}
int jb_sync(void) {
  if (int rc = ensure_device()) return rc;
  JB_CUDA(cudaStreamSynchronize(tls().stream));
  return 0;
}
// ---- peer memory between the one-process-per-GPU ranks of a node ------------------------------------
// A rank exports a jb_malloc'd device buffer as a 64-byte CUDA IPC handle, its peers map it and pull panels from it with
// copy-engine transfers over NVLink (jb_memcpy_async on a side stream) — no collective kernel competes with the
// persistent GEMM for SMs.
int jb_ipc_export(const void *dev_ptr, void *handle64) {
  if (int rc = ensure_device()) return rc;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
  JB_CUDA(cudaIpcGetMemHandle((cudaIpcMemHandle_t *)handle64, const_cast<void *>(dev_ptr)));
  return 0;
}
void *jb_ipc_open(const void *handle64) {
  if (ensure_device()) return nullptr;
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  void *p = nullptr;
  cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) { cudaGetLastError(); set_error(JB_ERR_CUDA, "jb_ipc_open: %s", cudaGetErrorString(e)); return nullptr; }
  return p;
}
int jb_ipc_close(void *p) {
  if (!p) return 0;
  JB_CUDA(cudaIpcCloseMemHandle(p));
  return 0;
}
// dst / src: any mix of device, peer-mapped, pinned or pageable memory (UVA); ordered on `cuda_stream`
int jb_memcpy_async(void *dst, const void *src, size_t bytes, void *cuda_stream) {
  if (int rc = ensure_device()) return rc;
  JB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, (cudaStream_t)cuda_stream));
  return 0;
}
// 2-D form: `height` rows of `width_bytes`, pitches in bytes (a K panel of a column-major block is a strided sub-matrix)
int jb_memcpy2d_async(void *dst, size_t dpitch, const void *src, size_t spitch, size_t width_bytes, size_t height, void *cuda_stream) {
  if (int rc = ensure_device()) return rc;
  JB_CUDA(cudaMemcpy2DAsync(dst, dpitch, src, spitch, width_bytes, height, cudaMemcpyDefault, (cudaStream_t)cuda_stream));
  return 0;
}
void jb_set_stream(void *s) { tls().stream = (cudaStream_t)s; }
void *jb_get_stream(void) { return (void *)tls().stream; }
uint64_t jb_kernel_launches(void) { return g_launches.load(); }
int jb_set_option(const char *key, int value) {
  if (!key) return JB_ERR_ARG;
  std::lock_guard<std::mutex> lk(g_mu);
  g_options[key] = value;
  g_opt_version.fetch_add(1, std::memory_order_release);
  return 0;
}

}  // extern "C"
