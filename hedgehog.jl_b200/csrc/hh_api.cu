// hh_api.cu — extern "C" surface of libhedgehog_mc.so: context lifetime and the thin wrappers
// that lock the context and forward to the kernels' host drivers. See include/hedgehog_mc.h.
#include <cstdlib>
#include <cstring>
#include <new>

#include "hh_ctx.h"

static thread_local std::string g_create_error;

namespace hh {
int tangent_sums(hh_ctx *ctx, const hh_model *model, const hh_tangent *tangents, int ntangents, const hh_sim *sim,
                 const hh_payoff *payoffs, int npayoffs, double *sums, double spot_bump, double *second_sums, double *kernel_ms);
int lsm_american(hh_ctx *ctx, const hh_model *model, const hh_sim *sim, const hh_payoff *payoff, int degree,
                 double step_discount, const hh_comm *comm, hh_lsm_result *out, int32_t *stop_idx, double *stop_val,
                 double *spot_paths);
int bk_chf(hh_ctx *ctx, const hh_model *model, double tau, const double *V0, const double *VT, int n, const double *a,
           int na, double *out_re, double *out_im);
int bk_log_besseli(hh_ctx *ctx, double nu, const double *z_re, const double *z_im, int n, double *out_re,
                   double *out_im);
int bk_european_launch(hh_ctx *ctx, const hh_model *model, const hh_sim *sim, const hh_payoff *payoffs, int npayoffs,
                       int want_terminal);
int bk_integral(hh_ctx *ctx, const hh_model *m, double tau, const hh_bk_config *cfg, const double *V0, const double *VT,
                const double *U, int n, double *out8);
int bk_variance(hh_ctx *ctx, const hh_model *m, double tau, const double *V0, int n, uint64_t seed, double *VT);
int bk_elementary(hh_ctx *ctx, int kind, const double *x, const double *y, int n, double *out_a, double *out_b);
int fp64_peak(hh_ctx *ctx, double *tflops, double *ms);
int heston_ablation(hh_ctx *ctx, int64_t n_paths, int n_steps, int rng_mode, int part, double *ms);
}  // namespace hh

extern "C" {

int hh_version(void) { return HH_VERSION; }

void hh_default_bk_config(hh_bk_config *out) {
  if (!out) return;
  out->n_std = 5;              // sample_from_cf.jl:27
  out->h_fd = 1e-2;            // :50
  out->cf_tol = 1e-3;          // :75
  out->atol = 1e-4;            // :110
  out->maxiter_newton = 10;    // :111
  out->maxiter_bisection = 100;  // :112
  out->max_terms = 4096;
}

int hh_create(hh_ctx **out, int device) {
  if (!out) return HH_ERR_ARG;
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    g_create_error = std::string("no usable CUDA device: ") + cudaGetErrorString(e) +
                     " (libhedgehog_mc has no CPU fallback)";
    return HH_ERR_CUDA;
  }
  if (device < 0 || device >= ndev) {
    g_create_error = "device index out of range";
    return HH_ERR_ARG;
  }
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) {
    g_create_error = cudaGetErrorString(e);
    return HH_ERR_CUDA;
  }
  if (prop.major != 10) {
    g_create_error = "libhedgehog_mc is built for sm_100a only; device is sm_" + std::to_string(prop.major) +
                     std::to_string(prop.minor);
    return HH_ERR_CUDA;
  }
  hh_ctx *ctx = new (std::nothrow) hh_ctx();
  if (!ctx) return HH_ERR_NOMEM;
  ctx->device = device;
  ctx->sm_count = prop.multiProcessorCount;
  ctx->cc_major = prop.major;
  ctx->cc_minor = prop.minor;
  ctx->total_mem = prop.totalGlobalMem;
  ctx->l2_bytes = (size_t)prop.l2CacheSize;
  ctx->l2_persist_max = (size_t)prop.persistingL2CacheMaxSize;
  ctx->l2_window_max = (size_t)prop.accessPolicyMaxWindowSize;
  int khz = 0;
  if (cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, device) == cudaSuccess && khz > 0) ctx->sm_clock_hz = 1e3 * (double)khz;
  if (const char *ts = getenv("HH_PEER_TIMEOUT_S")) {
    const double t = atof(ts);
    if (t > 0.0) ctx->peer_timeout_s = t;
  }
  bool ok = cudaSetDevice(device) == cudaSuccess &&
            cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) == cudaSuccess &&
            cudaEventCreate(&ctx->ev0) == cudaSuccess && cudaEventCreate(&ctx->ev1) == cudaSuccess &&
            cudaEventCreate(&ctx->ev2) == cudaSuccess;
  if (!ok) {
    g_create_error = std::string("context setup failed: ") + cudaGetErrorString(cudaGetLastError());
    delete ctx;
    return HH_ERR_CUDA;
  }
  ctx->stream = ctx->own_stream;
  *out = ctx;
  return HH_OK;
}

int hh_destroy(hh_ctx *ctx) {
  if (!ctx) return HH_ERR_ARG;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  ctx->for_each_buffer([](hh::DeviceBuffer &b) { b.release(); });
  for (int q = 0; q < ctx->peer_world; ++q)
    if (q != ctx->peer_rank && ctx->peer_mail[q]) cudaIpcCloseMemHandle(ctx->peer_mail[q]);
  if (ctx->mailbox) cudaFree(ctx->mailbox);
  if (ctx->h_pinned) cudaFreeHost(ctx->h_pinned);
  for (int i = 0; i < 16; ++i)
    if (ctx->ev_seg[i]) cudaEventDestroy(ctx->ev_seg[i]);
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  for (int b = 0; b < 2; ++b) {
    if (ctx->h_stage[b]) cudaFreeHost(ctx->h_stage[b]);
    if (ctx->ev_stage[b]) cudaEventDestroy(ctx->ev_stage[b]);
  }
  cudaEventDestroy(ctx->ev0);
  cudaEventDestroy(ctx->ev1);
  cudaEventDestroy(ctx->ev2);
  cudaStreamDestroy(ctx->own_stream);
  delete ctx;
  return HH_OK;
}

// mailbox layout (doubles): [parity 2][rank HH_MAX_PEERS][kMailSlot] payload, then flags [parity 2][rank] as uint64
static constexpr size_t kMailSlot = 32;
// ... and one "abort" word behind the flags (hh_lsm.cu: mail_abort), padded to 16 bytes
static constexpr size_t kMailBytes =
    2 * HH_MAX_PEERS * kMailSlot * sizeof(double) + 2 * HH_MAX_PEERS * sizeof(unsigned long long) + 16;

int hh_peer_export(hh_ctx *ctx, unsigned char handle[HH_IPC_HANDLE_BYTES]) {
  if (!ctx || !handle) return HH_ERR_ARG;
  std::lock_guard<std::mutex> lk(ctx->mu);
  static_assert(sizeof(cudaIpcMemHandle_t) <= HH_IPC_HANDLE_BYTES, "IPC handle size");
  HH_CUDA(ctx, cudaSetDevice(ctx->device));
  if (!ctx->mailbox) {
    HH_CUDA(ctx, cudaMalloc(&ctx->mailbox, kMailBytes));
    HH_CUDA(ctx, cudaMemset(ctx->mailbox, 0, kMailBytes));
    HH_CUDA(ctx, cudaDeviceSynchronize());
  }
  cudaIpcMemHandle_t h;
  HH_CUDA(ctx, cudaIpcGetMemHandle(&h, ctx->mailbox));
  memset(handle, 0, HH_IPC_HANDLE_BYTES);
  memcpy(handle, &h, sizeof h);
  return HH_OK;
}

int hh_peer_connect(hh_ctx *ctx, int rank, int world, const unsigned char *handles) {
  if (!ctx || !handles) return HH_ERR_ARG;
  std::lock_guard<std::mutex> lk(ctx->mu);
  if (world < 1 || world > HH_MAX_PEERS || rank < 0 || rank >= world)
    return ctx->fail(HH_ERR_ARG, "peer connect: rank %d / world %d out of range (max %d)", rank, world, HH_MAX_PEERS);
  if (!ctx->mailbox) return ctx->fail(HH_ERR_ARG, "peer connect: call hh_peer_export first");
  HH_CUDA(ctx, cudaSetDevice(ctx->device));
  // epochs restart at zero with every connection: clear flags left by an earlier one. The caller synchronises the ranks
  // between connect and the first exchange (nobody posts into a mailbox before everybody has connected).
  HH_CUDA(ctx, cudaMemset(ctx->mailbox, 0, kMailBytes));
  HH_CUDA(ctx, cudaDeviceSynchronize());
  for (int q = 0; q < world; ++q) {
    if (q == rank) {
      ctx->peer_mail[q] = ctx->mailbox;
      continue;
    }
    cudaIpcMemHandle_t h;
    memcpy(&h, handles + (size_t)q * HH_IPC_HANDLE_BYTES, sizeof h);
    void *p = nullptr;
    HH_CUDA(ctx, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    ctx->peer_mail[q] = p;
  }
  ctx->peer_rank = rank;
  ctx->peer_world = world;
  ctx->peer_epoch = 0;
  return HH_OK;
}

int hh_peer_set_timeout(hh_ctx *ctx, double seconds) {
  if (!ctx) return HH_ERR_ARG;
  std::lock_guard<std::mutex> lk(ctx->mu);
  if (!(seconds > 0.0) || seconds > 3600.0) return ctx->fail(HH_ERR_ARG, "peer timeout must be in (0, 3600] seconds (got %g)", seconds);
  ctx->peer_timeout_s = seconds;
  return HH_OK;
}

int hh_peer_disconnect(hh_ctx *ctx) {
  if (!ctx) return HH_ERR_ARG;
  std::lock_guard<std::mutex> lk(ctx->mu);
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  for (int q = 0; q < ctx->peer_world; ++q) {
    if (q != ctx->peer_rank && ctx->peer_mail[q]) cudaIpcCloseMemHandle(ctx->peer_mail[q]);
    ctx->peer_mail[q] = nullptr;
  }
  ctx->peer_world = 1;
  ctx->peer_rank = 0;
  (void)cudaGetLastError();
  return HH_OK;
}

const char *hh_last_error(hh_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int hh_set_stream(hh_ctx *ctx, void *cuda_stream) {
  if (!ctx) return HH_ERR_ARG;
  std::lock_guard<std::mutex> lk(ctx->mu);
  ctx->stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : ctx->own_stream;
  return HH_OK;
}

int hh_device_info(hh_ctx *ctx, int32_t *sm_count, int32_t *cc_major, int32_t *cc_minor, size_t *total_mem) {
  if (!ctx) return HH_ERR_ARG;
  if (sm_count) *sm_count = ctx->sm_count;
  if (cc_major) *cc_major = ctx->cc_major;
  if (cc_minor) *cc_minor = ctx->cc_minor;
  if (total_mem) *total_mem = ctx->total_mem;
  return HH_OK;
}

int hh_bench_fp64_peak(hh_ctx *ctx, double *tflops, double *ms) {
  if (!ctx) return HH_ERR_ARG;
  std::lock_guard<std::mutex> lk(ctx->mu);
  NvtxRange nv("hh_bench_fp64_peak");
  return hh::fp64_peak(ctx, tflops, ms);
}

static int european_launch_locked(hh_ctx *ctx, const hh_model *model, const hh_sim *sim, const hh_payoff *payoffs,
                                  int npayoffs, int want_terminal) {
  if (sim && sim->scheme == HH_SCHEME_HESTON_BK) {
    int rc = hh::validate_model_sim(ctx, model, sim);
    if (rc) return rc;
    return hh::bk_european_launch(ctx, model, sim, payoffs, npayoffs, want_terminal);
  }
  return hh::european_launch(ctx, model, sim, payoffs, npayoffs, want_terminal);
}

int hh_bench_heston_ablation(hh_ctx *ctx, int64_t n_paths, int n_steps, int rng_mode, int part, double *ms) {
  if (!ctx) return HH_ERR_ARG;
  std::lock_guard<std::mutex> lk(ctx->mu);
  NvtxRange nv("hh_bench_heston_ablation");
  return hh::heston_ablation(ctx, n_paths, n_steps, rng_mode, part, ms);
}

int hh_mc_european_launch(hh_ctx *ctx, const hh_model *model, const hh_sim *sim, const hh_payoff *payoffs, int npayoffs,
                          int want_terminal) {
  if (!ctx) return HH_ERR_ARG;
  std::lock_guard<std::mutex> lk(ctx->mu);
  NvtxRange nv("hh_mc_european_launch");
  return european_launch_locked(ctx, model, sim, payoffs, npayoffs, want_terminal);
}

int hh_mc_european_collect(hh_ctx *ctx, double discount, hh_result *results, double *terminal, size_t terminal_len) {
  if (!ctx) return HH_ERR_ARG;
  std::lock_guard<std::mutex> lk(ctx->mu);
  NvtxRange nv("hh_mc_european_collect");
  return hh::european_collect(ctx, discount, results, terminal, terminal_len);
}

// The blocking form holds the context mutex across launch AND collect: two threads sharing a context can never collect
// each other's results (the header promises serialised calls).
int hh_mc_european(hh_ctx *ctx, const hh_model *model, const hh_sim *sim, const hh_payoff *payoffs, int npayoffs,
                   double discount, hh_result *results, double *terminal, size_t terminal_len) {
  if (!ctx) return HH_ERR_ARG;
  std::lock_guard<std::mutex> lk(ctx->mu);
  NvtxRange nv("hh_mc_european");
  int rc = european_launch_locked(ctx, model, sim, payoffs, npayoffs, terminal != nullptr);
  if (rc) return rc;
  return hh::european_collect(ctx, discount, results, terminal, terminal_len);
}

int hh_mc_path_dependent(hh_ctx *ctx, const hh_model *model, const hh_sim *sim, int monitor_every,
                         const hh_path_payoff *payoffs, int npayoffs, double discount, hh_result *results,
                         double *path_stats, size_t path_stats_len) {
  if (!ctx) return HH_ERR_ARG;
  std::lock_guard<std::mutex> lk(ctx->mu);
  NvtxRange nv("hh_mc_path_dependent");
  return hh::path_dependent(ctx, model, sim, monitor_every, payoffs, npayoffs, discount, results, path_stats, path_stats_len);
}

int hh_mc_european_tangent_sums(hh_ctx *ctx, const hh_model *model, const hh_tangent *tangents, int ntangents,
                                const hh_sim *sim, const hh_payoff *payoffs, int npayoffs, double *sums,
                                double spot_bump, double *second_sums, double *kernel_ms) {
  if (!ctx) return HH_ERR_ARG;
  std::lock_guard<std::mutex> lk(ctx->mu);
  NvtxRange nv("hh_mc_european_tangent_sums");
  return hh::tangent_sums(ctx, model, tangents, ntangents, sim, payoffs, npayoffs, sums, spot_bump, second_sums, kernel_ms);
}

int hh_mc_european_tangent(hh_ctx *ctx, const hh_model *model, const hh_tangent *tangents, int ntangents,
                           const hh_sim *sim, const hh_payoff *payoffs, int npayoffs, double discount,
                           hh_result *results, double *tangent_results, double *tangent_stderr) {
  if (!ctx) return HH_ERR_ARG;
  if (!results || !tangent_results || ntangents < 1 || npayoffs < 1) {
    std::lock_guard<std::mutex> lk(ctx->mu);
    return ctx->fail(HH_ERR_ARG, "tangent: results/tangent_results NULL or empty request");
  }
  const int stride = 2 + 2 * ntangents;
  double *sums = new (std::nothrow) double[(size_t)npayoffs * stride];
  if (!sums) return HH_ERR_NOMEM;
  double ms = 0.0;
  int rc = hh_mc_european_tangent_sums(ctx, model, tangents, ntangents, sim, payoffs, npayoffs, sums, 0.0, nullptr, &ms);
  if (rc == HH_OK) {
    const double N = (double)sim->n_paths;
    for (int k = 0; k < npayoffs; ++k) {
      const double *s = sums + (size_t)k * stride;
      hh_result *r = &results[k];
      memset(r, 0, sizeof *r);
      r->sum = s[0];
      r->sumsq = s[1];
      r->n = sim->n_paths;
      const double mean = s[0] / N;
      r->price = discount * mean;
      double var = N > 1 ? (s[1] - N * mean * mean) / (N - 1) : 0.0;
      r->std_error = discount * sqrt((var > 0 ? var : 0) / N);
      r->kernel_ms = ms;
      for (int p = 0; p < ntangents; ++p) {
        // price = D * mean(payoff)  =>  d price = dD * mean(payoff) + D * mean(d payoff)   (montecarlo.jl:489-490)
        const double dmean = s[2 + p] / N;
        tangent_results[(size_t)k * ntangents + p] = tangents[p].ddiscount * mean + discount * dmean;
        if (tangent_stderr) {
          double tv = N > 1 ? (s[2 + ntangents + p] - N * dmean * dmean) / (N - 1) : 0.0;
          tangent_stderr[(size_t)k * ntangents + p] = discount * sqrt((tv > 0 ? tv : 0) / N);
        }
      }
    }
  }
  delete[] sums;
  return rc;
}

int hh_lsm_american(hh_ctx *ctx, const hh_model *model, const hh_sim *sim, const hh_payoff *payoff, int degree,
                    double step_discount, const hh_comm *comm, hh_lsm_result *out, int32_t *stop_idx, double *stop_val,
                    double *spot_paths) {
  if (!ctx) return HH_ERR_ARG;
  std::lock_guard<std::mutex> lk(ctx->mu);
  NvtxRange nv("hh_lsm_american");
  return hh::lsm_american(ctx, model, sim, payoff, degree, step_discount, comm, out, stop_idx, stop_val, spot_paths);
}

int hh_bk_chf(hh_ctx *ctx, const hh_model *model, double tau, const double *V0, const double *VT, int n, const double *a,
              int na, double *out_re, double *out_im) {
  if (!ctx) return HH_ERR_ARG;
  std::lock_guard<std::mutex> lk(ctx->mu);
  NvtxRange nv("hh_bk_chf");
  return hh::bk_chf(ctx, model, tau, V0, VT, n, a, na, out_re, out_im);
}

int hh_bk_log_besseli(hh_ctx *ctx, double nu, const double *z_re, const double *z_im, int n, double *out_re,
                      double *out_im) {
  if (!ctx) return HH_ERR_ARG;
  std::lock_guard<std::mutex> lk(ctx->mu);
  NvtxRange nv("hh_bk_log_besseli");
  return hh::bk_log_besseli(ctx, nu, z_re, z_im, n, out_re, out_im);
}

int hh_bk_integral(hh_ctx *ctx, const hh_model *model, double tau, const hh_bk_config *cfg, const double *V0,
                   const double *VT, const double *u, int n, double *out8) {
  if (!ctx) return HH_ERR_ARG;
  std::lock_guard<std::mutex> lk(ctx->mu);
  NvtxRange nv("hh_bk_integral");
  return hh::bk_integral(ctx, model, tau, cfg, V0, VT, u, n, out8);
}

int hh_bk_variance(hh_ctx *ctx, const hh_model *model, double tau, const double *V0, int n, uint64_t seed, double *VT) {
  if (!ctx) return HH_ERR_ARG;
  std::lock_guard<std::mutex> lk(ctx->mu);
  NvtxRange nv("hh_bk_variance");
  return hh::bk_variance(ctx, model, tau, V0, n, seed, VT);
}

int hh_bk_elementary(hh_ctx *ctx, int kind, const double *x, const double *y, int n, double *out_a, double *out_b) {
  if (!ctx) return HH_ERR_ARG;
  std::lock_guard<std::mutex> lk(ctx->mu);
  NvtxRange nv("hh_bk_elementary");
  return hh::bk_elementary(ctx, kind, x, y, n, out_a, out_b);
}

int hh_debug_check_guards(hh_ctx *ctx, int64_t *violations) {
  if (!ctx || !violations) return HH_ERR_ARG;
  std::lock_guard<std::mutex> lk(ctx->mu);
  HH_CUDA(ctx, cudaSetDevice(ctx->device));
  HH_CUDA(ctx, cudaDeviceSynchronize());
  long long bad = 0;
  bool failed = false;
  ctx->for_each_buffer([&](hh::DeviceBuffer &b) {
    const long long v = b.guard_violations();
    if (v < 0) failed = true;
    else bad += v;
  });
  if (failed) return ctx->fail(HH_ERR_CUDA, "hh_debug_check_guards: reading a guard band failed");
  *violations = hh::debug_guards() ? (int64_t)bad : -1;
  return HH_OK;
}

int hh_bk_last_stats(hh_ctx *ctx, double *out5) {
  if (!ctx || !out5) return HH_ERR_ARG;
  std::lock_guard<std::mutex> lk(ctx->mu);
  for (int i = 0; i < 5; ++i) out5[i] = ctx->bk_stats[i];
  return HH_OK;
}

}  // extern "C"
