// temporary: entry points under construction
#include "hh_ctx.h"
namespace hh {
int bk_chf(hh_ctx *ctx, const hh_model *, double, const double *, const double *, int, const double *, int, double *,
           double *) {
  return ctx->fail(HH_ERR_UNSUPPORTED, "BK not built yet");
}
int bk_log_besseli(hh_ctx *ctx, double, const double *, const double *, int, double *, double *) {
  return ctx->fail(HH_ERR_UNSUPPORTED, "BK not built yet");
}
int bk_european_launch(hh_ctx *ctx, const hh_model *, const hh_sim *, const hh_payoff *, int, int) {
  return ctx->fail(HH_ERR_UNSUPPORTED, "BK not built yet");
}
}  // namespace hh
