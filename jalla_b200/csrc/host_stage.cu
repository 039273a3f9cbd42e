// Pageable host operands (a stock Jalla Matrix is a mallocForeignPtrArray buffer,
// /root/reference/Numeric/Jalla/Matrix.hs:785-787 — ordinary pageable memory the CUDA copy engines cannot read directly).
// cudaMemcpy2DAsync from such memory is synchronous and staged by the driver on ONE thread (measured: cblas_dgemm N = 16384
// from numpy arrays 1019 ms against 259 ms from pinned ones).  Here the staging is the library's own: a ring of pinned
// slots, filled (H2D) or emptied (D2H) by a few host threads row by row, the device side of every slot an ordinary
// asynchronous copy.  H2D blocks the caller only for the host-side memcpy of a slot; D2H hands the slot to a drain
// thread and returns.  Pinned or registered host memory bypasses all of this (one cudaMemcpy2DAsync).
#include "jb_common.cuh"

#include <atomic>
#include <condition_variable>
#include <cstring>
#include <deque>
#include <mutex>
#include <thread>
#include <vector>

namespace jb {

bool is_pageable(const void *p) {
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return true; }
  return at.type == cudaMemoryTypeUnregistered;
}

namespace {

constexpr size_t SLOT_BYTES = 32u << 20;
constexpr size_t STAGE_MIN_BYTES = 16u << 20;      // smaller copies stay with the driver (no pool start-up for small calls)
constexpr int H2D_SLOTS = 4, D2H_SLOTS = 6;

// rows of `width` bytes between two pitched host buffers, split over the copy threads (the caller takes a share)
class RowCopier {
 public:
  RowCopier() {
    unsigned hc = std::thread::hardware_concurrency();
    int want = option("host.copy_threads");
    // measured on the 16-core host of a B200 box (tools/pageable_e2e.py, cblas_dgemm N = 16384, ms per call): 4 threads 350,
    // 6: 333, 12: 302, 16: 296 (the driver's own staging: 1019)
    nthreads_ = want > 0 ? want : (hc >= 4 ? (int)(hc * 3 / 4) : 2);
    if (nthreads_ > 16) nthreads_ = 16;
    for (int i = 1; i < nthreads_; i++) workers_.emplace_back([this, i] { loop(i); });
  }
  ~RowCopier() {
    { std::lock_guard<std::mutex> g(m_); stop_ = true; gen_++; }
    cv_.notify_all();
    for (auto &t : workers_) t.join();
  }
  void copy(char *dst, size_t dpitch, const char *src, size_t spitch, size_t width, size_t rows) {
    if (rows * width < (4u << 20) || nthreads_ == 1) { run(dst, dpitch, src, spitch, width, 0, rows); return; }
    {
      std::lock_guard<std::mutex> g(m_);
      dst_ = dst; dpitch_ = dpitch; src_ = src; spitch_ = spitch; width_ = width; rows_ = rows;
      pending_ = nthreads_ - 1; gen_++;
    }
    cv_.notify_all();
    share(0);
    std::unique_lock<std::mutex> g(m_);
    done_.wait(g, [this] { return pending_ == 0; });
  }

 private:
  static void run(char *dst, size_t dpitch, const char *src, size_t spitch, size_t width, size_t r0, size_t r1) {
    if (dpitch == width && spitch == width) { std::memcpy(dst + r0 * width, src + r0 * width, (r1 - r0) * width); return; }
    for (size_t r = r0; r < r1; r++) std::memcpy(dst + r * dpitch, src + r * spitch, width);
  }
  void share(int i) {
    const size_t per = (rows_ + nthreads_ - 1) / nthreads_, r0 = per * i, r1 = r0 + per < rows_ ? r0 + per : rows_;
    if (r0 < r1) run(dst_, dpitch_, src_, spitch_, width_, r0, r1);
  }
  void loop(int i) {
    unsigned long seen = 0;
    for (;;) {
      {
        std::unique_lock<std::mutex> g(m_);
        cv_.wait(g, [&] { return gen_ != seen; });
        seen = gen_;
        if (stop_) return;
      }
      share(i);
      bool last;
      { std::lock_guard<std::mutex> g(m_); last = --pending_ == 0; }
      if (last) done_.notify_all();
    }
  }
  int nthreads_ = 1;
  std::vector<std::thread> workers_;
  std::mutex m_;
  std::condition_variable cv_, done_;
  unsigned long gen_ = 0;
  int pending_ = 0;
  bool stop_ = false;
  char *dst_ = nullptr; const char *src_ = nullptr;
  size_t dpitch_ = 0, spitch_ = 0, width_ = 0, rows_ = 0;
};

struct Slot { char *buf = nullptr; cudaEvent_t ev = nullptr; bool busy = false; };

struct DrainTask { int slot; char *dst; size_t dpitch, width, rows; };

class Stager {
 public:
  // one staging CALL at a time per process: callers on several OS threads serialise here; recursive because one call may
  // hold several HostStage objects (the operands of a LAPACKE call)
  std::recursive_mutex big;

  int init() {
    if (ready_) return 0;
    for (int i = 0; i < H2D_SLOTS + D2H_SLOTS; i++) {
      Slot &s = slots_[i];
      if (cudaHostAlloc((void **)&s.buf, SLOT_BYTES, cudaHostAllocDefault) != cudaSuccess ||
          cudaEventCreateWithFlags(&s.ev, cudaEventDisableTiming) != cudaSuccess) {
        cudaGetLastError();
        set_error(JB_ERR_ALLOC, "pinned staging buffers (%d x %zu bytes) could not be allocated", H2D_SLOTS + D2H_SLOTS, SLOT_BYTES);
        return JB_ERR_ALLOC;
      }
    }
    drain_ = std::thread([this] { drain_loop(); });
    ready_ = true;
    return 0;
  }
  ~Stager() {
    if (!ready_) return;
    { std::lock_guard<std::mutex> g(qm_); stop_ = true; }
    qcv_.notify_all();
    drain_.join();
    // the pinned slots are left to process teardown: the CUDA context may already be gone in a static destructor
  }

  int h2d(char *dst, size_t dpitch, const char *src, size_t spitch, size_t width, size_t rows, cudaStream_t st) {
    const size_t per = SLOT_BYTES / width;
    if (per == 0) {           // a row wider than a slot: let the driver stage it
      return cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, rows, cudaMemcpyHostToDevice, st) == cudaSuccess ? 0 : fail("H2D");
    }
    for (size_t r0 = 0; r0 < rows; r0 += per) {
      const size_t nr = rows - r0 < per ? rows - r0 : per;
      Slot &s = slots_[h2d_next_];
      h2d_next_ = (h2d_next_ + 1) % H2D_SLOTS;
      if (s.busy && cudaEventSynchronize(s.ev) != cudaSuccess) return fail("H2D slot wait");
      copier_.copy(s.buf, width, src + r0 * spitch, spitch, width, nr);
      if (cudaMemcpy2DAsync(dst + r0 * dpitch, dpitch, s.buf, width, width, nr, cudaMemcpyHostToDevice, st) != cudaSuccess) return fail("H2D");
      if (cudaEventRecord(s.ev, st) != cudaSuccess) return fail("H2D event");
      s.busy = true;
    }
    return 0;
  }

  int d2h(char *dst, size_t dpitch, const char *src, size_t spitch, size_t width, size_t rows, cudaStream_t st) {
    const size_t per = SLOT_BYTES / width;
    if (per == 0) {
      return cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, rows, cudaMemcpyDeviceToHost, st) == cudaSuccess ? 0 : fail("D2H");
    }
    for (size_t r0 = 0; r0 < rows; r0 += per) {
      const size_t nr = rows - r0 < per ? rows - r0 : per;
      const int si = H2D_SLOTS + d2h_next_;
      d2h_next_ = (d2h_next_ + 1) % D2H_SLOTS;
      Slot &s = slots_[si];
      {
        std::unique_lock<std::mutex> g(qm_);
        qdone_.wait(g, [&] { return !s.busy; });
        s.busy = true;
      }
      if (cudaMemcpy2DAsync(s.buf, width, src + r0 * spitch, spitch, width, nr, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
          cudaEventRecord(s.ev, st) != cudaSuccess) {
        { std::lock_guard<std::mutex> g(qm_); s.busy = false; }
        return fail("D2H");
      }
      { std::lock_guard<std::mutex> g(qm_); q_.push_back(DrainTask{si, dst + r0 * dpitch, dpitch, width, nr}); inflight_++; }
      qcv_.notify_all();
    }
    return 0;
  }

  // every D2H handed to the drain thread has reached the caller's memory
  int drain() {
    std::unique_lock<std::mutex> g(qm_);
    qdone_.wait(g, [&] { return inflight_ == 0; });
    const int rc = drain_rc_;
    drain_rc_ = 0;
    return rc;
  }

 private:
  int fail(const char *what) {
    set_error(JB_ERR_CUDA, "host staging: %s failed: %s", what, cudaGetErrorString(cudaGetLastError()));
    return JB_ERR_CUDA;
  }
  void drain_loop() {
    for (;;) {
      DrainTask t;
      {
        std::unique_lock<std::mutex> g(qm_);
        qcv_.wait(g, [&] { return stop_ || !q_.empty(); });
        if (q_.empty()) return;
        t = q_.front(); q_.pop_front();
      }
      Slot &s = slots_[t.slot];
      const bool ok = cudaEventSynchronize(s.ev) == cudaSuccess;
      if (ok) drain_copier_.copy(t.dst, t.dpitch, s.buf, t.width, t.width, t.rows);
      {
        std::lock_guard<std::mutex> g(qm_);
        if (!ok) drain_rc_ = JB_ERR_CUDA;
        s.busy = false;
        inflight_--;
      }
      qdone_.notify_all();
    }
  }
  bool ready_ = false;
  Slot slots_[H2D_SLOTS + D2H_SLOTS];
  int h2d_next_ = 0, d2h_next_ = 0;
  RowCopier copier_, drain_copier_;
  std::thread drain_;
  std::mutex qm_;
  std::condition_variable qcv_, qdone_;
  std::deque<DrainTask> q_;
  int inflight_ = 0, drain_rc_ = 0;
  bool stop_ = false;
};

Stager &stager() {
  static Stager *s = new Stager();      // never destroyed: worker threads must not outlive a static destructor's CUDA context
  return *s;
}

}  // namespace

// ---- HostStage: the staging of ONE library call -----------------------------------------------------------------------
HostStage::HostStage() {}
HostStage::~HostStage() { if (locked_) { stager().drain(); stager().big.unlock(); } }

int HostStage::acquire() {
  if (locked_) return 0;
  stager().big.lock();
  locked_ = true;
  if (int rc = stager().init()) { stager().big.unlock(); locked_ = false; return rc; }
  return 0;
}

int HostStage::h2d(void *dst, size_t dpitch, const void *src, size_t spitch, size_t width, size_t rows, cudaStream_t st) {
  if (rows == 0 || width == 0) return 0;
  if (option("host.stage") == 0 || width * rows < STAGE_MIN_BYTES || !is_pageable(src)) {
    if (cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, rows, cudaMemcpyHostToDevice, st) != cudaSuccess) {
      set_error(JB_ERR_CUDA, "H2D copy failed: %s", cudaGetErrorString(cudaGetLastError()));
      return JB_ERR_CUDA;
    }
    return 0;
  }
  if (int rc = acquire()) return rc;
  return stager().h2d((char *)dst, dpitch, (const char *)src, spitch, width, rows, st);
}

int HostStage::d2h(void *dst, size_t dpitch, const void *src, size_t spitch, size_t width, size_t rows, cudaStream_t st) {
  if (rows == 0 || width == 0) return 0;
  if (option("host.stage") == 0 || width * rows < STAGE_MIN_BYTES || !is_pageable(dst)) {
    if (cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, rows, cudaMemcpyDeviceToHost, st) != cudaSuccess) {
      set_error(JB_ERR_CUDA, "D2H copy failed: %s", cudaGetErrorString(cudaGetLastError()));
      return JB_ERR_CUDA;
    }
    return 0;
  }
  if (int rc = acquire()) return rc;
  return stager().d2h((char *)dst, dpitch, (const char *)src, spitch, width, rows, st);
}

int HostStage::finish() {
  if (!locked_) return 0;
  const int rc = stager().drain();
  stager().big.unlock();
  locked_ = false;
  if (rc) set_error(JB_ERR_CUDA, "host staging: a device-to-host slot failed");
  return rc;
}

}  // namespace jb
