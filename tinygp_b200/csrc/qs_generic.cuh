// Generic (any block structure) quasiseparable scans for ONE state dimension J: the kernels and the host drivers that launch
// them.  Included by quasisep.cu (J = 1..6) and by quasisep_j7.cu / quasisep_j8.cu, which hold the explicit instantiations of the
// two widest state dimensions (their kernels dominate the compile time: three translation units build in parallel).
#pragma once
#include "common.cuh"
#include <limits.h>
#include "qs_core.cuh"

#include "qs_tree.cuh"

// ---------------------------------------------------------------------------------------------
// scan kernels: thin wrappers over the __host__ __device__ bodies of qs_core.cuh (one thread = one chunk)
// ---------------------------------------------------------------------------------------------
template <int J>
__global__ void __launch_bounds__(QS_THREADS) chol_chunk_kernel(const __grid_constant__ QsModel m, const double* __restrict__ t,
                                                                const double* __restrict__ diag, int64_t n,
                                                                double* comp, int64_t nchunks) {
    const int64_t ch = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (ch < nchunks) chol_chunk_body<J>(m, t, diag, n, comp, nchunks, ch);
}
template <int J>
__global__ void __launch_bounds__(QS_THREADS) chol_replay_kernel(const __grid_constant__ QsModel m, const double* __restrict__ t,
                                                                 const double* __restrict__ diag, int64_t n,
                                                                 const double* fstart, int64_t nchunks, double* c_out,
                                                                 double* w_out, double* logc_part, int* info,
                                                                 const double* __restrict__ x_fuse, double* aff_comp) {
    const int64_t ch = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (ch < nchunks) chol_replay_body<J>(m, t, diag, n, fstart, nchunks, c_out, w_out, logc_part, info, x_fuse, aff_comp, ch);
}
template <int J, int OP>
__global__ void __launch_bounds__(QS_THREADS) affine_chunk_kernel(const __grid_constant__ QsModel m, const double* __restrict__ t,
                                                                  const double* __restrict__ c, const double* __restrict__ w,
                                                                  const double* __restrict__ x, int64_t n, double* comp,
                                                                  int64_t nchunks) {
    const int64_t ch = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (ch < nchunks) affine_chunk_body<J, OP>(m, t, c, w, x, n, comp, nchunks, ch);
}
template <int J, int OP>
__global__ void __launch_bounds__(QS_THREADS) affine_replay_kernel(const __grid_constant__ QsModel m, const double* __restrict__ t,
                                                                   const double* __restrict__ diag, const double* __restrict__ c,
                                                                   const double* __restrict__ w, const double* x, int64_t n,
                                                                   const double* gstart, int64_t nchunks, double* out,
                                                                   double* sq_part) {
    const int64_t ch = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (ch < nchunks) affine_replay_body<J, OP>(m, t, diag, c, w, x, n, gstart, nchunks, out, sq_part, ch);
}

// ---------------------------------------------------------------------------------------------
// diag((L L^T)^-1) by a backward scan: bodies in qs_core.cuh (shared with the host check in tests/csrc)
// ---------------------------------------------------------------------------------------------
template <int J>
__global__ void __launch_bounds__(QS_THREADS) gram_chunk_kernel(const __grid_constant__ QsModel m, const double* __restrict__ t,
                                                                const double* __restrict__ c, const double* __restrict__ w,
                                                                int64_t n, double* comp, int64_t nchunks) {
    const int64_t ch = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (ch < nchunks) gram_chunk_body<J>(m, t, c, w, n, comp, nchunks, ch);
}
template <int J>
__global__ void __launch_bounds__(QS_THREADS) gram_replay_kernel(const __grid_constant__ QsModel m, const double* __restrict__ t,
                                                                 const double* __restrict__ c, const double* __restrict__ w,
                                                                 int64_t n, const double* tstart, int64_t nchunks, double* out) {
    const int64_t ch = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (ch < nchunks) gram_replay_body<J>(m, t, c, w, n, tstart, nchunks, out, ch);
}

// GeneralQSM.matmul epilogue (kernels/quasisep.py:118-145 + general.py:84-104), one thread per test point:
//   idx = searchsorted(t2, x, right) - 1;
//   lower = [h T(t2[idx], x)^T] . f[idx]           if 0 <= idx < n
//   upper = [(h Pinf) T(x, t2[idx+1])] . g[idx+1]  if -1 <= idx < n-1
template <int J>
__global__ void general_gather_kernel(const __grid_constant__ QsModel m, const double* __restrict__ t2, int64_t n,
                                      const double* __restrict__ t1, int64_t mtest, const double* __restrict__ F,
                                      const double* __restrict__ G, double* out, int64_t out_stride) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= mtest) return;
    const double x = t1[i];
    int64_t lo = 0, hi = n;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (t2[mid] <= x) lo = mid + 1; else hi = mid;
    }
    const int64_t idx = lo - 1;
    double acc = 0.0;
    double a[J][J], p[J], st[J];
    if (idx >= 0) {
        qs_gen<J>(m, x - t2[idx], a, p);   // a = T(t2[idx], x)^T, p = h a = pl
        ldrow<J>(F, idx, st);
#pragma unroll
        for (int j = 0; j < J; ++j) acc += p[j] * st[j];
    }
    if (idx < n - 1) {
        qs_gen<J>(m, t2[idx + 1] - x, a, p);   // a = T(x, t2[idx+1])^T ; qu = q T = q a^T
        ldrow<J>(G, idx + 1, st);
        double up = 0.0;
#pragma unroll
        for (int j = 0; j < J; ++j) {
            double qu = 0.0;
#pragma unroll
            for (int k = 0; k < J; ++k) qu += m.q[k] * a[j][k];
            up += qu * st[j];
        }
        acc += up;
    }
    out[i * out_stride] = acc;
}
template <int J>
__global__ void generators_kernel(const __grid_constant__ QsModel m, const double* t, const double* diag, int64_t n,
                                  double* d, double* p, double* q, double* a) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    double aa[J][J], pp[J];
    qs_gen<J>(m, (k == 0) ? 0.0 : (t[k] - t[k - 1]), aa, pp);
    d[k] = m.d0 + (diag ? diag[k] : 0.0);
    for (int i = 0; i < J; ++i) {
        p[k * J + i] = pp[i];
        q[k * J + i] = m.q[i];
        for (int j = 0; j < J; ++j) a[(k * J + i) * J + j] = aa[i][j];
    }
}

template <int J>
void qs_factor_J(b200gp_qs* s, int* info_dev, double* logdet_dev, const double* x_fuse) {
    b200gp_ctx* ctx = s->ctx;
    const int64_t n = s->n, nch = (n + s->model.chunk - 1) / s->model.chunk;
    const size_t cb = (size_t)Riccati<J>::SIZE * nch * 8, sb = (size_t)J * J * nch * 8;
    double* comp = (double*)ctx->alloc(cb);
    double* fstart = (double*)ctx->alloc(sb);
    double* part = (double*)ctx->alloc((size_t)nch * 8);
    chol_chunk_kernel<J><<<nblk(nch, QS_THREADS), QS_THREADS, 0, ctx->stream>>>(s->model, s->t, s->diag, n, comp, nch);
    ctx->launches++;
    run_tree<Riccati<J>>(ctx, comp, nch, fstart);
    if (x_fuse) {
        s->fused_comp_bytes = (size_t)Affine<J>::SIZE * nch * 8;
        s->fused_comp = (double*)ctx->alloc(s->fused_comp_bytes);
    }
    chol_replay_kernel<J><<<nblk(nch, QS_THREADS), QS_THREADS, 0, ctx->stream>>>(s->model, s->t, s->diag, n, fstart, nch,
                                                                                 s->c, s->w, part, info_dev, x_fuse,
                                                                                 s->fused_comp);
    ctx->launches++;
    sum_partials(ctx, part, nch, logdet_dev);
    CUDA_CHECK(cudaGetLastError());
    ctx->release(comp, cb);
    ctx->release(fstart, sb);
    ctx->release(part, (size_t)nch * 8);
}

// one affine scan over a device vector x -> out ; optional sum of squares of the emitted values
template <int J, int OP>
void qs_affine_J(b200gp_qs* s, const double* x, double* out, double* sumsq_dev, double* precomputed_comp = nullptr) {
    b200gp_ctx* ctx = s->ctx;
    const int64_t n = s->n, nch = (n + s->model.chunk - 1) / s->model.chunk;
    const size_t cb = (size_t)Affine<J>::SIZE * nch * 8, sb = (size_t)J * nch * 8;
    double* comp = precomputed_comp ? precomputed_comp : (double*)ctx->alloc(cb);
    double* gstart = (double*)ctx->alloc(sb);
    double* part = sumsq_dev ? (double*)ctx->alloc((size_t)nch * 8) : nullptr;
    if (!precomputed_comp) {
        affine_chunk_kernel<J, OP><<<nblk(nch, QS_THREADS), QS_THREADS, 0, ctx->stream>>>(s->model, s->t, s->c, s->w, x, n, comp, nch);
        ctx->launches++;
    }
    run_tree<Affine<J>>(ctx, comp, nch, gstart);
    affine_replay_kernel<J, OP><<<nblk(nch, QS_THREADS), QS_THREADS, 0, ctx->stream>>>(s->model, s->t, s->diag, s->c, s->w, x, n,
                                                                                       gstart, nch, out, part);
    ctx->launches++;
    if (sumsq_dev) {
        sum_partials(ctx, part, nch, sumsq_dev);
        ctx->release(part, (size_t)nch * 8);
    }
    CUDA_CHECK(cudaGetLastError());
    if (!precomputed_comp) ctx->release(comp, cb);
    ctx->release(gstart, sb);
}

template <int J>
void qs_inv_diag_J(b200gp_qs* s, double* out_dev) {
    b200gp_ctx* ctx = s->ctx;
    const int64_t n = s->n, nch = (n + s->model.chunk - 1) / s->model.chunk;
    const size_t cb = (size_t)GramBack<J>::SIZE * nch * 8, sb = (size_t)J * J * nch * 8;
    Scratch comp(ctx, cb), tstart(ctx, sb);
    gram_chunk_kernel<J><<<nblk(nch, QS_THREADS), QS_THREADS, 0, ctx->stream>>>(s->model, s->t, s->c, s->w, n, comp.f64(),
                                                                              nch);
    ctx->launches++;
    run_tree<GramBack<J>>(ctx, comp.f64(), nch, tstart.f64());
    gram_replay_kernel<J><<<nblk(nch, QS_THREADS), QS_THREADS, 0, ctx->stream>>>(s->model, s->t, s->c, s->w, n, tstart.f64(),
                                                                               nch, out_dev);
    ctx->launches++;
    CUDA_CHECK(cudaGetLastError());
}

// the two point-wise kernels that are launched directly (generators for the QSM algebra, GeneralQSM.matmul's gather)
template <int J>
void qs_generators_J(b200gp_ctx* ctx, const QsModel& model, const double* t, const double* diag, int64_t n, double* d, double* p,
                     double* q, double* a) {
    generators_kernel<J><<<nblk(n, 128), 128, 0, ctx->stream>>>(model, t, diag, n, d, p, q, a);
}
template <int J>
void qs_general_gather_J(b200gp_ctx* ctx, const QsModel& model, const double* t2, int64_t n, const double* t1, int64_t m,
                         const double* F, const double* G, double* out, int64_t out_stride) {
    general_gather_kernel<J><<<nblk(m, 128), 128, 0, ctx->stream>>>(model, t2, n, t1, m, F, G, out, out_stride);
}

// explicit instantiation of every host driver for one J: `QS_FOR_J(template, 7)` defines them (quasisep_j7.cu),
// `QS_FOR_J(extern template, 7)` tells quasisep.cu that another translation unit does
#define QS_FOR_J(KW, JV)                                                                                                     \
    KW void qs_factor_J<JV>(b200gp_qs*, int*, double*, const double*);                                                       \
    KW void qs_affine_J<JV, OP_LOWER_SOLVE>(b200gp_qs*, const double*, double*, double*, double*);                           \
    KW void qs_affine_J<JV, OP_UPPER_SOLVE>(b200gp_qs*, const double*, double*, double*, double*);                           \
    KW void qs_affine_J<JV, OP_LOWER_DOT>(b200gp_qs*, const double*, double*, double*, double*);                             \
    KW void qs_affine_J<JV, OP_SYMM_LOWER>(b200gp_qs*, const double*, double*, double*, double*);                            \
    KW void qs_affine_J<JV, OP_SYMM_UPPER>(b200gp_qs*, const double*, double*, double*, double*);                            \
    KW void qs_affine_J<JV, OP_GEN_LOWER>(b200gp_qs*, const double*, double*, double*, double*);                             \
    KW void qs_affine_J<JV, OP_GEN_UPPER>(b200gp_qs*, const double*, double*, double*, double*);                             \
    KW void qs_inv_diag_J<JV>(b200gp_qs*, double*);                                                                          \
    KW void qs_generators_J<JV>(b200gp_ctx*, const QsModel&, const double*, const double*, int64_t, double*, double*, double*, \
                                double*);                                                                                    \
    KW void qs_general_gather_J<JV>(b200gp_ctx*, const QsModel&, const double*, int64_t, const double*, int64_t, const double*, \
                                    const double*, double*, int64_t);
