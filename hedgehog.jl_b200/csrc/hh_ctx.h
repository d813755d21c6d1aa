// hh_ctx.h — the opaque context behind the C ABI: one CUDA device, one stream, growable device
// scratch, last-error text. Host-side only.
#pragma once
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <string>

#include "../../include/hedgehog_mc.h"

namespace hh {

// HH_DEBUG_GUARDS=1 (read once per process): every scratch buffer is allocated with a 4 KB guard band on each side, the
// whole allocation is filled with 0xFF bytes (NaN as f64, huge as integers: a read of memory the kernels never wrote
// shows up as a non-finite result) and hh_debug_check_guards counts guard bytes that changed (a write out of bounds).
// compute-sanitizer is closed on the GPU pool this library is developed on; this is the stand-in for its memcheck /
// initcheck on the small cases of tools/sanity_small.py (tests/test_gpu_guards.py).
constexpr size_t kGuardBytes = 4096;
inline bool debug_guards() {
  static const bool on = getenv("HH_DEBUG_GUARDS") && atoi(getenv("HH_DEBUG_GUARDS")) != 0;
  return on;
}

struct DeviceBuffer {
  void *ptr = nullptr;
  size_t cap = 0;
  size_t guard = 0;  // bytes of guard band on each side of [ptr, ptr + cap)
  cudaError_t ensure(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    release();
    size_t want = bytes + bytes / 8 + 256;
    want = (want + 255) & ~(size_t)255;
    const size_t g = debug_guards() ? kGuardBytes : 0;
    void *base = nullptr;
    cudaError_t e = cudaMalloc(&base, want + 2 * g);
    if (e != cudaSuccess) return e;
    if (g) {
      e = cudaMemset(base, 0xFF, want + 2 * g);
      if (e != cudaSuccess) {
        cudaFree(base);
        return e;
      }
    }
    ptr = static_cast<char *>(base) + g;
    cap = want;
    guard = g;
    return cudaSuccess;
  }
  void release() {
    if (ptr) cudaFree(static_cast<char *>(ptr) - guard);
    ptr = nullptr;
    cap = 0;
    guard = 0;
  }
  // guard bytes that no longer hold 0xFF (0 without HH_DEBUG_GUARDS); the device must be idle
  long long guard_violations() const {
    if (!ptr || !guard) return 0;
    unsigned char host[kGuardBytes];
    long long bad = 0;
    for (int side = 0; side < 2; ++side) {
      const char *src = side ? static_cast<char *>(ptr) + cap : static_cast<char *>(ptr) - guard;
      if (cudaMemcpy(host, src, guard, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
      for (size_t i = 0; i < guard; ++i) bad += host[i] != 0xFF;
    }
    return bad;
  }
  template <class T>
  T *as() const {
    return static_cast<T *>(ptr);
  }
};

// "Done once PER DEVICE": the opt-in above 48 KB of dynamic shared memory and the constant-table uploads are per device,
// and contexts on different GPUs may reach the same call site from different threads (each context has its own mutex).
struct PerDeviceOnce {
  std::atomic<unsigned long long> mask{0};
  bool done(int dev) const { return dev >= 0 && dev < 64 && ((mask.load(std::memory_order_acquire) >> dev) & 1ull); }
  void set(int dev) {
    if (dev >= 0 && dev < 64) mask.fetch_or(1ull << dev, std::memory_order_release);
  }
};

// cudaFuncAttributeMaxDynamicSharedMemorySize for `kern` on the CURRENT device, once per device.
template <class K>
inline cudaError_t smem_opt_in(PerDeviceOnce &once, K kern, int bytes) {
  int dev = -1;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (once.done(dev)) return cudaSuccess;
  e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess) once.set(dev);
  return e;
}

struct PendingEuropean {
  bool active = false;
  int npay = 0;
  int nblocks = 0;
  int64_t n = 0;
  bool anti = false;
  bool want_terminal = false;
  bool bk = false;  // Broadie-Kaya run: counters live in d_counters
  // the simulation was launched in `nseg` segments of trajectories so that the copy of a segment's terminal values to the
  // host overlaps the kernel of the next one; seg_end[i] = first trajectory after segment i
  int nseg = 0;
  int64_t seg_end[16] = {};
};

}  // namespace hh

struct hh_ctx {
  int device = 0;
  int sm_count = 0;
  int cc_major = 0, cc_minor = 0;
  size_t total_mem = 0;
  size_t l2_bytes = 0, l2_persist_max = 0, l2_window_max = 0;  // L2 size, largest persisting carve-out, largest policy window
  cudaStream_t own_stream = nullptr;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr;
  std::mutex mu;
  std::string err;

  hh::DeviceBuffer d_payoffs, d_partials, d_final, d_terminal, d_seeds, d_normals, d_tangents;
  hh::DeviceBuffer d_grid, d_cash, d_tau, d_lsm_partials, d_lsm_state, d_misc, d_counters, d_bk_slab, d_bk_work, d_lsm_uab;
  // peer mailboxes (hh_peer_*): own buffer + the peers' buffers mapped through CUDA IPC
  void *mailbox = nullptr;
  void *peer_mail[HH_MAX_PEERS] = {};
  int peer_rank = 0, peer_world = 1;
  unsigned long long peer_epoch = 0;  // advances identically on every rank (one per exchanged date)
  double peer_timeout_s = 30.0;       // in-kernel limit on one peer wait (hh_peer_set_timeout)
  double sm_clock_hz = 1.965e9;       // clock64() rate used to turn that limit into cycles
  double bk_stats[5] = {0, 0, 0, 0, 0};
  void *h_pinned = nullptr;  // small pinned staging area for results
  size_t h_pinned_cap = 0;
  void *h_stage[2] = {nullptr, nullptr};  // pinned double buffer for large device -> pageable-host copies
  cudaEvent_t ev_stage[2] = {nullptr, nullptr};
  cudaStream_t copy_stream = nullptr;     // device -> host copies that overlap the kernels on `stream`
  cudaEvent_t ev_seg[16] = {};            // segment i of a segmented launch has finished
  hh::PendingEuropean pend;

  template <class F>
  void for_each_buffer(F f) {
    hh::DeviceBuffer *bufs[] = {&d_payoffs, &d_partials, &d_final, &d_terminal, &d_seeds, &d_normals, &d_tangents, &d_grid, &d_cash,
                                &d_tau, &d_lsm_partials, &d_lsm_state, &d_misc, &d_counters, &d_bk_slab, &d_bk_work, &d_lsm_uab};
    for (auto *b : bufs) f(*b);
  }

  int fail(int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    err = buf;
    return code;
  }
  int fail_cuda(cudaError_t e, const char *what, const char *file, int line) {
    int code = (e == cudaErrorMemoryAllocation) ? HH_ERR_NOMEM : HH_ERR_CUDA;
    return fail(code, "CUDA error %d (%s) in %s at %s:%d", (int)e, cudaGetErrorString(e), what, file, line);
  }
};

// NVTX range around a C-ABI call (visible in nsys / ncu --nvtx; a no-op costing a few ns when no tool is attached).
struct NvtxRange {
  explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
  NvtxRange(const NvtxRange &) = delete;
  NvtxRange &operator=(const NvtxRange &) = delete;
};

#define HH_CUDA(ctx, call)                                                   \
  do {                                                                       \
    cudaError_t _e = (call);                                                 \
    if (_e != cudaSuccess) return (ctx)->fail_cuda(_e, #call, __FILE__, __LINE__); \
  } while (0)

namespace hh {
// implemented in hh_european.cu
int european_launch(hh_ctx *ctx, const hh_model *model, const hh_sim *sim, const hh_payoff *payoffs, int npayoffs,
                    int want_terminal);
int european_collect(hh_ctx *ctx, double discount, hh_result *results, double *terminal, size_t terminal_len);
int validate_model_sim(hh_ctx *ctx, const hh_model *model, const hh_sim *sim);
// implemented in hh_pathdep.cu
int path_dependent(hh_ctx *ctx, const hh_model *model, const hh_sim *sim, int monitor_every, const hh_path_payoff *payoffs,
                   int npayoffs, double discount, hh_result *results, double *path_stats, size_t path_stats_len);
// implemented in hh_bk.cu
int bk_path_stats_launch(hh_ctx *ctx, const hh_model *model, const hh_sim *sim, int monitor_every, double *d_stats);
int bk_read_counters(hh_ctx *ctx, int64_t *n_fallback);
// spots of every trajectory at dates 0..n_steps into the date-major grid of the LSM driver (hh_lsm.cu)
int bk_path_grid_launch(hh_ctx *ctx, const hh_model *model, const hh_sim *sim, double *d_grid, int64_t grid_stride);
// device -> the caller's pageable buffer through a pinned double buffer with parallel host copies (large outputs)
int copy_to_pageable_host(hh_ctx *ctx, void *dst, const void *src_dev, size_t bytes, cudaStream_t st);
}  // namespace hh
