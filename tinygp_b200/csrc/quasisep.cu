// placeholder until the scan kernels land (next commit): every entry point fails loudly.
#include "common.cuh"
#define QS_NI(ctxexpr)                                              \
    b200gp_ctx* c_ = (ctxexpr);                                     \
    if (c_) c_->err = "quasisep path not implemented yet";          \
    return 4;
struct b200gp_qs { b200gp_ctx* ctx; };
extern "C" {
int b200gp_qs_check_sorted(b200gp_ctx* ctx, const double*, int64_t, int*) { QS_NI(ctx) }
int b200gp_qs_create(b200gp_ctx* ctx, const double*, int, const double*, int64_t, const double*, int, b200gp_qs**, int*, int*) { QS_NI(ctx) }
int b200gp_qs_create_dev(b200gp_ctx* ctx, const double*, int, const double*, int64_t, const double*, int, b200gp_qs**, int*, int*) { QS_NI(ctx) }
int b200gp_qs_free(b200gp_qs* s) { (void)s; return 0; }
int b200gp_qs_state_dim(b200gp_qs* s, int*) { QS_NI(s ? s->ctx : nullptr) }
int b200gp_qs_logdet_half(b200gp_qs* s, double*) { QS_NI(s ? s->ctx : nullptr) }
int b200gp_qs_variance(b200gp_qs* s, double*) { QS_NI(s ? s->ctx : nullptr) }
int b200gp_qs_get_factor(b200gp_qs* s, double*, double*) { QS_NI(s ? s->ctx : nullptr) }
int b200gp_qs_get_generators(b200gp_qs* s, double*, double*, double*, double*) { QS_NI(s ? s->ctx : nullptr) }
int b200gp_qs_solve_triangular(b200gp_qs* s, double*, int64_t, int) { QS_NI(s ? s->ctx : nullptr) }
int b200gp_qs_dot_triangular(b200gp_qs* s, double*, int64_t) { QS_NI(s ? s->ctx : nullptr) }
int b200gp_qs_matmul(b200gp_qs* s, double*, int64_t) { QS_NI(s ? s->ctx : nullptr) }
int b200gp_qs_log_probability(b200gp_ctx* ctx, const double*, int, const double*, int64_t, const double*, const double*, int, int*, double*) { QS_NI(ctx) }
int b200gp_qs_log_probability_dev(b200gp_ctx* ctx, const double*, int, const double*, int64_t, const double*, const double*, int, int*, double*) { QS_NI(ctx) }
int b200gp_searchsorted_right_m1(b200gp_ctx* ctx, const double*, int64_t, const double*, int64_t, int64_t*) { QS_NI(ctx) }
int b200gp_dense_log_probability_batched(b200gp_ctx* ctx, const double*, int, int64_t, const double*, int64_t, int, const double*, const double*, double*) { QS_NI(ctx) }
}
