// estep_fp.cu — OPT-IN E-step with a true per-document fixed point.
//
// The reference's inner loop (tuotuo/lda_model.py:236-254) never refreshes exp(E[log theta_d]) and therefore
// degenerates to one update (estep.cu reproduces exactly that).  This kernel runs the loop the way Blei/Hoffman and
// scikit-learn's _update_doc_distribution do: after every gamma_d update, exp(E[log theta_d]) is recomputed (AS-103
// digamma, as everywhere on this path) and the update is repeated until ||gamma_new - gamma_old||_1 <= tol or the
// cap — the reference's own loop condition (`while l1 > threshold and it <= step`).  The statistics pass then reads
// the REFRESHED exp(E[log theta]) table, i.e. the normaliser is re-evaluated with the final theta like scikit-learn
// does.  Each extra inner iteration re-gathers the document's exp(E[log beta]) rows (L1/L2 hits: same rows as the
// pass before) — algorithmic bytes are those of ttlda_estep_f64 times the mean inner iteration count, reported by
// iters_out.  Same warp-per-document / lane-group mapping as estep.cu (traverse.cuh).
#include "traverse.cuh"

namespace ttlda {

struct FixedPointParams {
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
    const double* gamma_in;
    double* gamma_out;
    double* eT_out;
    int64_t max_iter;
    double tol;
    unsigned long long* iters;
    double* records;
};

constexpr int kFpThreads = 128;

template <int G, int E>
__global__ void __launch_bounds__(kFpThreads) doc_fixed_point_kernel(const FixedPointParams P)
{
    constexpr int R = 32 / G;
    constexpr int LDP = 2 * G * E;
    constexpr int MMAX = (LDP + 31) / 32;
    constexpr bool PF = (E <= 6);
    extern __shared__ double smem[];
    double* red = smem;
    double* wbuf = smem + 96 + (threadIdx.x >> 5) * ((R + 2) * LDP);  // R group partials + theta + previous gamma
    double* tbuf = wbuf + R * LDP;
    double* gbuf = tbuf + LDP;

    const int lane = threadIdx.x & 31;
    const int g = lane % G, r = lane / G;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int ld = (int)P.ld, K = P.K;
    const bool last_ok = KSlice<G, E>::last_slice_ok(g, ld);
    const uint64_t pol = l2_evict_first_policy();

    double lga = 0.;
    for (int k = lane; k < K; k += 32) lga += lgamma(P.alpha_vec[k]);
    lga = warp_sum(lga);
    const double lg_asum = lgamma(P.alpha_sum);
    double tok_acc = 0., theta_acc = 0., cnt_acc = 0.;
    unsigned long long passes = 0, capped = 0;

    for (int64_t d = warp0; d < P.D; d += nwarps) {
        const int64_t b = P.rowptr[d], e = P.rowptr[d + 1];
        KSlice<G, E> t;
        t.load(P.eT_in + d * P.ld, g, last_ok);
        // 32-lane-layout copies of exp(E[log theta_d]) and of the previous gamma_d live in shared memory (tbuf, gbuf)
#pragma unroll
        for (int m = 0; m < MMAX; ++m) {
            const int k = lane + 32 * m;
            if (k < LDP) {
                tbuf[k] = (k < K) ? P.eT_in[d * P.ld + k] : 0.;
                gbuf[k] = (k < K) ? P.gamma_in[d * P.ld + k] : 0.;
            }
        }
        __syncwarp();
        int64_t it = 0;
        double l1, s;
        do {
            KSlice<G, E> acc;
            acc.zero();
            for (int64_t n0 = b; n0 < e; n0 += 32) {
                const int cnt = (int)((e - n0 < 32) ? (e - n0) : 32);
                int my_w = 0, my_c = 0;
                if (lane < cnt) {
                    my_w = ld_stream_s32(P.colidx + n0 + lane, pol);
                    my_c = ld_stream_s32(P.counts + n0 + lane, pol);
                }
                double my_p = 1.;
                walk_rows<G, E, PF>(P.eB, (int)P.ld, my_w, cnt, g, r, last_ok, [&](int sb, const KSlice<G, E>& row) {
                    const int c = __shfl_sync(0xffffffffu, my_c, sb * R + r);
                    const double p = group_sum<G>(t.dot(row));
                    acc.axpy(fast_div_pos((double)c, p + 1e-100), row);  // lda_model.py:242,248
                    const double pv = __shfl_sync(0xffffffffu, p, (lane % R) * G);
                    if (lane / R == sb) my_p = pv;
                });
                if (it == 0 && lane < cnt) {  // ELBO token term of the incoming state: first pass only
                    tok_acc += (double)my_c * log(my_p);
                    cnt_acc += (double)my_c;
                }
            }
            acc.to_smem(wbuf + r * LDP, g);
            __syncwarp();
            s = 0.;
            l1 = 0.;
#pragma unroll
            for (int m = 0; m < MMAX; ++m) {
                const int k = lane + 32 * m;
                if (k < K) {
                    double a = 0.;
#pragma unroll
                    for (int q = 0; q < R; ++q) a += wbuf[q * LDP + k];
                    const double gnew = P.alpha_vec[k] + tbuf[k] * a;  // lda_model.py:248
                    s += gnew;
                    l1 += fabs(gnew - gbuf[k]);  // :251-252
                    gbuf[k] = gnew;
                }
            }
            s = warp_sum(s);
            l1 = warp_sum(l1);
            // refresh exp(E[log theta_d]) — the step the reference's loop omits
            const double pt = psi_as103(s);
#pragma unroll
            for (int m = 0; m < MMAX; ++m) {
                const int k = lane + 32 * m;
                if (k < K) tbuf[k] = exp(psi_as103(gbuf[k]) - pt);
            }
            __syncwarp();
#pragma unroll
            for (int j = 0; j < E; ++j) t.v[j] = *reinterpret_cast<const double2*>(tbuf + 2 * (g + G * j));
            ++it;
        } while (l1 > P.tol && it <= P.max_iter);  // lda_model.py:236
        passes += (unsigned long long)it;
        capped += (it > P.max_iter) ? 1ull : 0ull;  // :258

        double pt, lg_s;
        psi_lgamma_as103(s, pt, lg_s);
        double th = 0.;
#pragma unroll
        for (int m = 0; m < MMAX; ++m) {
            const int k = lane + 32 * m;
            if (k < K) {
                const double gk = gbuf[k];
                double ps, lg;
                psi_lgamma_as103(gk, ps, lg);
                th += (P.alpha_vec[k] - gk) * (ps - pt) + lg;
                P.gamma_out[d * P.ld + k] = gk;
                P.eT_out[d * P.ld + k] = tbuf[k];
            }
        }
        if (lane < ld - K) {
            P.gamma_out[d * P.ld + K + lane] = 0.;
            P.eT_out[d * P.ld + K + lane] = 0.;
        }
        __syncwarp();
        th = warp_sum(th);
        theta_acc += th - lga + (lg_asum - lg_s);
    }

    double v[3] = {tok_acc, (lane == 0) ? theta_acc : 0., cnt_acc};
    block_sum<3>(v, red);
    if (threadIdx.x == 0) {
        double* rec = P.records + (size_t)blockIdx.x * NPART;
        rec[0] = v[0];
        rec[1] = v[1];
        rec[2] = v[2];
#pragma unroll
        for (int j = 3; j < NPART; ++j) rec[j] = 0.;
    }
    if (P.iters && lane == 0 && (passes | capped)) {
        atomicAdd(P.iters, passes);
        if (capped) atomicAdd(P.iters + 1, capped);
    }
}

template <int G, int E>
static int launch_fp(const FixedPointParams& P, int grid, cudaStream_t s)
{
    constexpr int R = 32 / G;
    constexpr size_t smem = (96 + (kFpThreads / 32) * (R + 2) * 2 * G * E) * sizeof(double);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(doc_fixed_point_kernel<G, E>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)smem);
        if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(fixed_point)");
    }
    doc_fixed_point_kernel<G, E><<<grid, kFpThreads, smem, s>>>(P);
    return TTLDA_OK;
}

}  // namespace ttlda

using namespace ttlda;

extern "C" int ttlda_estep_fixed_point_f64(const int64_t* rowptr, const int32_t* colidx, const int32_t* counts,
                                           int64_t D, int64_t K, int64_t V, int64_t ld, const double* alpha_vec,
                                           double alpha_sum, const double* exp_elogtheta_in,
                                           const double* exp_elogbeta, const double* gamma_in, double* gamma_out,
                                           double* exp_elogtheta_out, int64_t max_iter, double tol, int64_t* iters_out,
                                           double* partial_out, void* workspace, int64_t workspace_bytes, void* stream)
{
    TTLDA_REQUIRE(rowptr && alpha_vec && exp_elogtheta_in && exp_elogbeta && gamma_in && gamma_out &&
                      exp_elogtheta_out && partial_out && workspace,
                  "NULL pointer");
    TTLDA_REQUIRE(D >= 0 && V >= 0 && K >= 1 && K <= TTLDA_MAX_K, "K must be in [1, 1024]");
    TTLDA_REQUIRE(ld >= K && ld % 4 == 0 && ld < K + 4, "ld must be K rounded up to a multiple of 4");
    TTLDA_REQUIRE(exp_elogtheta_in != exp_elogtheta_out && gamma_in != gamma_out, "outputs must not alias the inputs");
    TTLDA_REQUIRE(max_iter >= 0, "max_iter must be >= 0");
    if (workspace_bytes < reduce_ws_bytes()) {
        set_error("estep_fixed_point: workspace too small");
        return TTLDA_EWORKSPACE;
    }
    cudaStream_t s = (cudaStream_t)stream;
    if (iters_out) {
        cudaError_t e = cudaMemsetAsync(iters_out, 0, 2 * sizeof(int64_t), s);
        if (e != cudaSuccess) return cuda_fail(e, "memset iters");
    }
    FixedPointParams P{rowptr, colidx, counts, D, (int)K, ld, alpha_vec, alpha_sum, exp_elogtheta_in, exp_elogbeta,
                       gamma_in, gamma_out, exp_elogtheta_out, max_iter, tol, (unsigned long long*)iters_out,
                       (double*)workspace};
    const int grid = grid_for_warps(D, kFpThreads / 32, 16);
    const GroupShape gs = group_shape_for(ld);
    int rc = TTLDA_EINVAL;
    bool launched = false;
#define TTLDA_CASE(GG, EE)                        \
    if (!launched && gs.G == GG && gs.E == EE) {  \
        rc = launch_fp<GG, EE>(P, grid, s);       \
        launched = true;                          \
    }
    TTLDA_FOR_EACH_SHAPE(TTLDA_CASE)
#undef TTLDA_CASE
    if (!launched) {
        set_error("no kernel instantiated for ld=%lld", (long long)ld);
        return TTLDA_EINVAL;
    }
    if (rc != TTLDA_OK) return rc;
    TTLDA_LAUNCH_CHECK("estep_fixed_point");
    return launch_reduce_records((const double*)workspace, grid, partial_out, s);
}
