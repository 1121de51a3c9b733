// estep.cu — document-major pass: E-step gamma update fused with the ELBO token term (incoming state), the theta
// refresh and the theta-prior ELBO term (outgoing state); and the stand-alone ELBO token kernel.
//
// Reference: LDASmoothed.e_step tuotuo/lda_model.py:197-270; approx_elbo token loop :125-139; theta-prior term
// :25-46,:142-148; Dirichlet expectation tuotuo/cutils.pyx:41-93.
//
// Mapping: ONE WARP PER DOCUMENT, the K topics live in registers: lane l owns the double2 slices
// k = 2l + 64j (+0,+1), j < KP, so that one warp-wide 16-byte load covers 512 contiguous bytes of a word's
// K-vector in the word-major exp(E[log beta]) table [V][ld].  Per nonzero (w, c):
//     p = sum_k eT_k * eB[w,k]            (2*KP DFMA per lane + butterfly)
//     r = c / (p + 1e-100)                (lda_model.py:242)
//     acc_k += r * eB[w,k]                (lda_model.py:248, the (cts/norm) . eB^T product)
// and, after the row, gamma_k = alpha_k + eT_k * acc_k.  The CSR (word, count) pairs are fetched 32 at a time
// (one coalesced load per lane) and broadcast with shuffles; the next word's K-vector is prefetched into
// registers while the current one is consumed.  Bound: HBM/L2 gather of 8K bytes per nonzero.
#include "ttlda_common.cuh"

namespace ttlda {

struct EStepParams {
    const int64_t* rowptr;
    const int32_t* colidx;
    const int32_t* counts;
    int64_t D;
    int K;
    int64_t ld;
    const double* alpha_vec;
    double alpha_sum;
    const double* eT_in;
    const double* eB;
    double* gamma;
    double* eT_out;
    double* sstats_atomic;
    int64_t max_iter;
    double tol;
    unsigned long long* iters;  // [2] device counters (passes, capped docs) or NULL
    double* records;
};

template <int KP>
__device__ __forceinline__ void load_row(double2 (&v)[KP], const double* __restrict__ row, int lane, int ld)
{
#pragma unroll
    for (int j = 0; j < KP; ++j) {
        const int k = 2 * lane + 64 * j;
        v[j] = (k < ld) ? ldg2(row + k) : make_double2(0., 0.);
    }
}

template <int KP>
__device__ __forceinline__ double dot_row(const double2 (&t)[KP], const double2 (&b)[KP])
{
    double p = 0.;
#pragma unroll
    for (int j = 0; j < KP; ++j) {
        p = fma(t[j].x, b[j].x, p);
        p = fma(t[j].y, b[j].y, p);
    }
    return warp_sum(p);
}

template <int KP, bool PREFETCH, bool ESTEP>
__global__ void __launch_bounds__(256) doc_pass_kernel(const EStepParams P)
{
    __shared__ double sm[3 * 32];
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int ld = (int)P.ld, K = P.K;

    double lga = 0., lg_asum = 0.;
    if (ESTEP) {
        for (int k = lane; k < K; k += 32) lga += lgamma(P.alpha_vec[k]);
        lga = warp_sum(lga);
        lg_asum = lgamma(P.alpha_sum);
    }
    double tok_acc = 0., theta_acc = 0., cnt_acc = 0.;
    unsigned long long passes = 0, capped = 0;

    for (int64_t d = warp0; d < P.D; d += nwarps) {
        const int64_t b = P.rowptr[d], e = P.rowptr[d + 1];
        double2 t[KP], acc[KP];
        load_row<KP>(t, P.eT_in + d * P.ld, lane, ld);
#pragma unroll
        for (int j = 0; j < KP; ++j) acc[j] = make_double2(0., 0.);

        for (int64_t n0 = b; n0 < e; n0 += 32) {
            const int cnt = (int)((e - n0 < 32) ? (e - n0) : 32);
            int my_w = 0, my_c = 0;
            if (lane < cnt) {
                my_w = ld_stream_s32(P.colidx + n0 + lane);
                my_c = ld_stream_s32(P.counts + n0 + lane);
            }
            double my_p = 1.;
            double2 cur[KP], nxt[KP];
            {
                const int w = __shfl_sync(0xffffffffu, my_w, 0);
                load_row<KP>(cur, P.eB + (int64_t)w * P.ld, lane, ld);
            }
            for (int i = 0; i < cnt; ++i) {
                if (PREFETCH && i + 1 < cnt) {
                    const int wn = __shfl_sync(0xffffffffu, my_w, i + 1);
                    load_row<KP>(nxt, P.eB + (int64_t)wn * P.ld, lane, ld);
                }
                const int w = __shfl_sync(0xffffffffu, my_w, i);
                const int c = __shfl_sync(0xffffffffu, my_c, i);
                const double p = dot_row<KP>(t, cur);
                if (lane == i) my_p = p;
                if (ESTEP) {
                    const double r = (double)c / (p + 1e-100);
#pragma unroll
                    for (int j = 0; j < KP; ++j) {
                        acc[j].x = fma(r, cur[j].x, acc[j].x);
                        acc[j].y = fma(r, cur[j].y, acc[j].y);
                    }
                    if (P.sstats_atomic) {
                        double* srow = P.sstats_atomic + (int64_t)w * P.ld;
#pragma unroll
                        for (int j = 0; j < KP; ++j) {
                            const int k = 2 * lane + 64 * j;
                            if (k < K) atomicAdd(srow + k, t[j].x * r);
                            if (k + 1 < K) atomicAdd(srow + k + 1, t[j].y * r);
                        }
                    }
                }
                if (PREFETCH) {
#pragma unroll
                    for (int j = 0; j < KP; ++j) cur[j] = nxt[j];
                } else if (i + 1 < cnt) {
                    const int wn = __shfl_sync(0xffffffffu, my_w, i + 1);
                    load_row<KP>(cur, P.eB + (int64_t)wn * P.ld, lane, ld);
                }
            }
            // one log per nonzero, evaluated by the lane that owns it
            if (lane < cnt) {
                tok_acc += (double)my_c * log(my_p);
                cnt_acc += (double)my_c;
            }
        }

        if (ESTEP) {
            // gamma_k = alpha_k + eT_k * acc_k  (lda_model.py:248), then theta refresh + theta-prior term
            double* g = P.gamma + d * P.ld;
            double2 gn[KP];
            double s = 0., l1 = 0.;
#pragma unroll
            for (int j = 0; j < KP; ++j) {
                const int k = 2 * lane + 64 * j;
                gn[j] = make_double2(0., 0.);
                if (k < K) {
                    gn[j].x = P.alpha_vec[k] + t[j].x * acc[j].x;
                    s += gn[j].x;
                    if (P.iters) l1 += fabs(gn[j].x - g[k]);
                }
                if (k + 1 < K) {
                    gn[j].y = P.alpha_vec[k + 1] + t[j].y * acc[j].y;
                    s += gn[j].y;
                    if (P.iters) l1 += fabs(gn[j].y - g[k + 1]);
                }
            }
            s = warp_sum(s);
            if (P.iters) {
                l1 = warp_sum(l1);
                // the reference's while-loop (lda_model.py:236): pass 1 always runs; pass 2 runs iff the first
                // moved gamma by more than tol and the cap allows it, and then reproduces pass 1 (delta = 0).
                const unsigned long long np = (l1 > P.tol && 1 <= P.max_iter) ? 2ull : 1ull;
                passes += np;
                capped += ((int64_t)np > P.max_iter) ? 1ull : 0ull;
            }
            const double pt = psi_as103(s);
            double th = 0.;
#pragma unroll
            for (int j = 0; j < KP; ++j) {
                const int k = 2 * lane + 64 * j;
                double2 et = make_double2(0., 0.);
                if (k < K) {
                    const double el = psi_as103(gn[j].x) - pt;
                    th += (P.alpha_vec[k] - gn[j].x) * el + lgamma(gn[j].x);
                    et.x = exp(el);
                }
                if (k + 1 < K) {
                    const double el = psi_as103(gn[j].y) - pt;
                    th += (P.alpha_vec[k + 1] - gn[j].y) * el + lgamma(gn[j].y);
                    et.y = exp(el);
                }
                if (k < ld) {
                    *reinterpret_cast<double2*>(g + k) = gn[j];
                    *reinterpret_cast<double2*>(P.eT_out + d * P.ld + k) = et;
                }
            }
            th = warp_sum(th);
            theta_acc += th - lga + (lg_asum - lgamma(s));
        }
    }

    double v[3] = {tok_acc, (lane == 0) ? theta_acc : 0., cnt_acc};
    block_sum<3>(v, sm);
    if (threadIdx.x == 0) {
        double* rec = P.records + (size_t)blockIdx.x * NPART;
        rec[0] = v[0];
        rec[1] = v[1];
        rec[2] = v[2];
#pragma unroll
        for (int j = 3; j < NPART; ++j) rec[j] = 0.;
    }
    if (ESTEP && P.iters && lane == 0 && (passes | capped)) {
        atomicAdd(P.iters, passes);
        if (capped) atomicAdd(P.iters + 1, capped);
    }
}

template <bool ESTEP>
static int launch_doc_pass(const EStepParams& P, int grid, cudaStream_t s)
{
    const int need = (int)((P.ld + 63) / 64);
#define TTLDA_CASE(KP, PF)                                            \
    if (need <= KP) {                                                 \
        doc_pass_kernel<KP, PF, ESTEP><<<grid, 256, 0, s>>>(P);        \
        return TTLDA_OK;                                              \
    }
    TTLDA_CASE(1, true)
    TTLDA_CASE(2, true)
    TTLDA_CASE(3, true)
    TTLDA_CASE(4, true)
    TTLDA_CASE(6, false)
    TTLDA_CASE(8, false)
    TTLDA_CASE(12, false)
    TTLDA_CASE(16, false)
#undef TTLDA_CASE
    set_error("K too large (ld=%lld > %d)", (long long)P.ld, TTLDA_MAX_K);
    return TTLDA_EINVAL;
}

}  // namespace ttlda

using namespace ttlda;

extern "C" {

int ttlda_estep_f64(const int64_t* rowptr, const int32_t* colidx, const int32_t* counts, int64_t D, int64_t K,
                    int64_t V, int64_t ld, const double* alpha_vec, double alpha_sum,
                    const double* exp_elogtheta_in, const double* exp_elogbeta, double* gamma,
                    double* exp_elogtheta_out, double* sstats_atomic, int64_t max_iter, double tol,
                    int64_t* iters_out, double* partial_out, void* workspace, int64_t workspace_bytes, void* stream)
{
    TTLDA_REQUIRE(rowptr && alpha_vec && exp_elogtheta_in && exp_elogbeta && gamma && exp_elogtheta_out &&
                      partial_out && workspace,
                  "NULL pointer");
    TTLDA_REQUIRE(D >= 0 && V >= 0 && K >= 1 && K <= TTLDA_MAX_K, "K must be in [1, 1024]");
    TTLDA_REQUIRE(ld >= K && ld % 4 == 0, "ld must be a multiple of 4 and >= K");
    TTLDA_REQUIRE(exp_elogtheta_in != exp_elogtheta_out, "exp_elogtheta_out must not alias the input");
    TTLDA_REQUIRE(max_iter >= 0, "max_iter must be >= 0");
    if (workspace_bytes < reduce_ws_bytes()) {
        set_error("estep: workspace %lld < %lld", (long long)workspace_bytes, (long long)reduce_ws_bytes());
        return TTLDA_EWORKSPACE;
    }
    cudaStream_t s = (cudaStream_t)stream;
    if (iters_out) {
        cudaError_t e = cudaMemsetAsync(iters_out, 0, 2 * sizeof(int64_t), s);
        if (e != cudaSuccess) return cuda_fail(e, "memset iters");
    }
    EStepParams P{rowptr, colidx, counts, D, (int)K, ld, alpha_vec, alpha_sum, exp_elogtheta_in, exp_elogbeta,
                  gamma, exp_elogtheta_out, sstats_atomic, max_iter, tol, (unsigned long long*)iters_out,
                  (double*)workspace};
    const int grid = grid_for_warps(D, 8, 8);
    int rc = launch_doc_pass<true>(P, grid, s);
    if (rc != TTLDA_OK) return rc;
    TTLDA_LAUNCH_CHECK("estep");
    return launch_reduce_records((const double*)workspace, grid, partial_out, s);
}

int ttlda_elbo_tokens_f64(const int64_t* rowptr, const int32_t* colidx, const int32_t* counts, int64_t D, int64_t K,
                          int64_t ld, const double* exp_elogtheta, const double* exp_elogbeta, double* partial_out,
                          void* workspace, int64_t workspace_bytes, void* stream)
{
    TTLDA_REQUIRE(rowptr && exp_elogtheta && exp_elogbeta && partial_out && workspace, "NULL pointer");
    TTLDA_REQUIRE(D >= 0 && K >= 1 && K <= TTLDA_MAX_K, "K must be in [1, 1024]");
    TTLDA_REQUIRE(ld >= K && ld % 4 == 0, "ld must be a multiple of 4 and >= K");
    if (workspace_bytes < reduce_ws_bytes()) {
        set_error("elbo_tokens: workspace too small");
        return TTLDA_EWORKSPACE;
    }
    cudaStream_t s = (cudaStream_t)stream;
    EStepParams P{rowptr, colidx, counts, D, (int)K, ld, nullptr, 0., exp_elogtheta, exp_elogbeta,
                  nullptr, nullptr, nullptr, 0, 0., nullptr, (double*)workspace};
    const int grid = grid_for_warps(D, 8, 8);
    int rc = launch_doc_pass<false>(P, grid, s);
    if (rc != TTLDA_OK) return rc;
    TTLDA_LAUNCH_CHECK("elbo_tokens");
    return launch_reduce_records((const double*)workspace, grid, partial_out, s);
}

}  // extern "C"
