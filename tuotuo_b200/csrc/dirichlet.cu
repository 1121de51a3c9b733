// dirichlet.cu — row-wise Dirichlet expectation (AS-103 digamma), theta refresh + theta-prior ELBO term,
// exact-digamma Newton statistics.
//
// Reference: tuotuo/cutils.pyx:11-93 (psi, _dirichlet_expectation_1d/_2d); call sites tuotuo/lda_model.py:116,
// 120,256,281 ; expec_real_dist_minus_suro_dist tuotuo/lda_model.py:25-46 ; update_alpha/eta :449-462,:522-528.
//
// HBM-bound elementwise/row kernels: rows are read once with coalesced accesses, reductions are warp shuffles
// (rows <= 2048 columns: one warp per row) or a CTA per row (long rows, e.g. a (K,V) matrix).
#include "ttlda_common.cuh"
#include "traverse.cuh"

namespace ttlda {

// ---- generic rows: warp per row ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) dirichlet_rows_warp_kernel(const double* __restrict__ a, int64_t rows,
                                                                  int64_t cols, int64_t ld,
                                                                  double* __restrict__ out_elog,
                                                                  double* __restrict__ out_exp)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp0; r < rows; r += nwarps) {
        const double* row = a + r * ld;
        double s = 0.;
        for (int64_t j = lane; j < cols; j += 32) s += row[j];
        s = warp_sum(s);
        const double pt = psi_as103(s);
        for (int64_t j = lane; j < ld; j += 32) {
            const bool live = j < cols;
            const double e = live ? psi_as103(row[j]) - pt : 0.;
            if (out_elog) out_elog[r * ld + j] = e;
            if (out_exp) out_exp[r * ld + j] = live ? exp(e) : 0.;
        }
    }
}

// ---- long rows: CTA per row ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) dirichlet_rows_block_kernel(const double* __restrict__ a, int64_t rows,
                                                                   int64_t cols, int64_t ld,
                                                                   double* __restrict__ out_elog,
                                                                   double* __restrict__ out_exp)
{
    __shared__ double sm[32];
    __shared__ double s_pt;
    for (int64_t r = blockIdx.x; r < rows; r += gridDim.x) {
        const double* row = a + r * ld;
        double v[1] = {0.};
        for (int64_t j = threadIdx.x; j < cols; j += blockDim.x) v[0] += row[j];
        block_sum<1>(v, sm);
        if (threadIdx.x == 0) s_pt = psi_as103(v[0]);
        __syncthreads();
        const double pt = s_pt;
        for (int64_t j = threadIdx.x; j < ld; j += blockDim.x) {
            const bool live = j < cols;
            const double e = live ? psi_as103(row[j]) - pt : 0.;
            if (out_elog) out_elog[r * ld + j] = e;
            if (out_exp) out_exp[r * ld + j] = live ? exp(e) : 0.;
        }
        __syncthreads();
    }
}

__global__ void dirichlet_1d_kernel(double* __restrict__ doc_topic, double prior, double* __restrict__ out, int64_t n)
{
    __shared__ double sm[32];
    __shared__ double s_pt;
    double v[1] = {0.};
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
        const double dt = doc_topic[i] + prior;  // cutils.pyx:31-33 (in-place prior add)
        doc_topic[i] = dt;
        v[0] += dt;
    }
    block_sum<1>(v, sm);
    if (threadIdx.x == 0) s_pt = psi_as103(v[0]);
    __syncthreads();
    const double pt = s_pt;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) out[i] = exp(psi_as103(doc_topic[i]) - pt);
}

// ---- theta refresh: eT = exp(DE(gamma)), theta-prior term ------------------------------------------------------
// LPD lanes per document (32 / LPD documents per warp); lane l of a group owns topics l, l + LPD, ...  The digamma /
// log-gamma evaluations are warp-wide instructions, so what counts is evaluations per document:
// (ceil(K / LPD) + 1 for the row sum) * LPD / 32 — K = 100: 5 with a full warp, 3.5 with 8 lanes; K = 20: 2 vs 0.75
// with 4 lanes.  Partial record per CTA -> fixed-order final reduction.
template <int LPD>
__global__ void __launch_bounds__(256) theta_refresh_kernel(const double* __restrict__ gamma, int64_t D, int K,
                                                            int64_t ld, const double* __restrict__ alpha_vec,
                                                            double alpha_sum, double* __restrict__ eT,
                                                            double* __restrict__ records)
{
    constexpr int DPW = 32 / LPD;  // documents per warp
    __shared__ double sm[32];
    const int lane = threadIdx.x & 31;
    const int l = lane % LPD, grp = lane / LPD;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    // per-warp constants: sum_k lgamma(alpha_k), lgamma(sum(alpha))
    double lga = 0.;
    for (int k = lane; k < K; k += 32) lga += lgamma(alpha_vec[k]);
    lga = warp_sum(lga);
    const double lg_asum = lgamma(alpha_sum);
    double term = 0.;
    const int64_t ngroups = (D + DPW - 1) / DPW;
    for (int64_t w = warp0; w < ngroups; w += nwarps) {
        const int64_t d = w * DPW + grp;
        const bool live = d < D;  // the last warp's trailing groups idle (they still take part in the shuffles)
        const double* g = gamma + (live ? d : 0) * ld;
        double s = 0.;
        for (int k = l; k < K; k += LPD) s += g[k];
        s = group_sum<LPD>(s);
        double pt, lg_s;
        psi_lgamma_as103(s, pt, lg_s);
        double t = 0., pp = 1.;  // pp: product of this lane's shift products (each in [1e-4, 4e5])
        int nf = 0;
        for (int k = l; k < (int)ld; k += LPD) {
            if (k < K) {
                const double gk = g[k];
                double ps, lg, p;
                psi_lgamma_as103_split(gk, ps, lg, p);
                pp *= p;
                if (++nf == 32) {  // keep the running product inside the double range for any K / LPD
                    t -= log(pp);
                    pp = 1.;
                    nf = 0;
                }
                const double e = ps - pt;
                t += (alpha_vec[k] - gk) * e + lg;
                if (eT && live) eT[d * ld + k] = exp(e);
            } else if (eT && live) {
                eT[d * ld + k] = 0.;
            }
        }
        t -= log(pp);  // sum_k lgamma(gamma_k) = sum_k Stirling_k - log(prod_k p_k): one log per lane
        t = group_sum<LPD>(t);
        if (live && l == 0) term += t - lga + (lg_asum - lg_s);
    }
    double v[1] = {term};
    block_sum<1>(v, sm);
    if (threadIdx.x == 0) {
        double* rec = records + (size_t)blockIdx.x * NPART;
        rec[0] = v[0];
#pragma unroll
        for (int j = 1; j < NPART; ++j) rec[j] = 0.;
    }
}

// ---- per-column exact-digamma statistics: out[k] = sum_r (psi(x_rk) - psi(sum_j x_rj)) --------------------------------
// (the gradient term of the non-exchangeable update_alpha, lda_model.py:452-456).  One warp per row, lane l owns
// columns l, l+32, ...; per-CTA partial column sums go to the workspace, hyper_colstats_finish_kernel adds them in CTA
// order => bit-reproducible.
constexpr int kColMax = TTLDA_MAX_K / 32;  // columns per lane

__global__ void __launch_bounds__(256) hyper_colstats_kernel(const double* __restrict__ x, int64_t rows, int cols,
                                                             int64_t ld, double* __restrict__ part)
{
    extern __shared__ double sm_col[];  // [warps][cols]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    double acc[kColMax];
#pragma unroll
    for (int m = 0; m < kColMax; ++m) acc[m] = 0.;
    for (int64_t r = warp0; r < rows; r += nwarps) {
        const double* row = x + r * ld;
        double s = 0.;
        for (int k = lane; k < cols; k += 32) s += row[k];
        const double ps = psi_exact(warp_sum(s));
#pragma unroll
        for (int m = 0; m < kColMax; ++m) {
            const int k = lane + 32 * m;
            if (k < cols) acc[m] += psi_exact(row[k]) - ps;
        }
    }
#pragma unroll
    for (int m = 0; m < kColMax; ++m) {
        const int k = lane + 32 * m;
        if (k < cols) sm_col[warp * cols + k] = acc[m];
    }
    __syncthreads();
    for (int k = threadIdx.x; k < cols; k += blockDim.x) {
        double t = 0.;
        for (int w = 0; w < nw; ++w) t += sm_col[w * cols + k];
        part[(int64_t)blockIdx.x * cols + k] = t;
    }
}

__global__ void hyper_colstats_finish_kernel(const double* __restrict__ part, int nblocks, int cols,
                                             double* __restrict__ out)
{
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < cols; k += gridDim.x * blockDim.x) {
        double t = 0.;
        for (int b = 0; b < nblocks; ++b) t += part[(int64_t)b * cols + k];
        out[k] = t;
    }
}

// ---- exact-digamma statistics ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) hyper_stats_kernel(const double* __restrict__ x, int64_t rows, int64_t cols,
                                                          int64_t ld, double* __restrict__ records)
{
    __shared__ double sm[2 * 32];
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    double a = 0., b = 0.;
    for (int64_t r = warp0; r < rows; r += nwarps) {
        const double* row = x + r * ld;
        double s = 0., p = 0.;
        for (int64_t j = lane; j < cols; j += 32) {
            const double v = row[j];
            s += v;
            p += psi_exact(v);
        }
        s = warp_sum(s);
        a += p;
        if (lane == 0) b += psi_exact(s);
    }
    double v[2] = {a, b};
    block_sum<2>(v, sm);
    if (threadIdx.x == 0) {
        double* rec = records + (size_t)blockIdx.x * NPART;
        rec[0] = v[0];
        rec[1] = v[1];
#pragma unroll
        for (int j = 2; j < NPART; ++j) rec[j] = 0.;
    }
}

__global__ void psi_exact_kernel(const double* __restrict__ x, int64_t n, double* __restrict__ out)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = psi_exact(x[i]);
}

}  // namespace ttlda

using namespace ttlda;

extern "C" {

int ttlda_dirichlet_expectation_f64(const double* a, int64_t rows, int64_t cols, int64_t ld, double* out_elog,
                                    double* out_exp, void* stream)
{
    TTLDA_REQUIRE(a != nullptr, "a is NULL");
    TTLDA_REQUIRE(out_elog || out_exp, "both outputs NULL");
    TTLDA_REQUIRE(rows >= 0 && cols >= 0 && ld >= cols, "bad shape");
    if (rows == 0 || ld == 0) return TTLDA_OK;
    cudaStream_t s = (cudaStream_t)stream;
    if (cols <= 2048) {
        int grid = grid_for_warps(rows, 8, 8);
        dirichlet_rows_warp_kernel<<<grid, 256, 0, s>>>(a, rows, cols, ld, out_elog, out_exp);
    } else {
        int grid = (int)(rows < (int64_t)kSMs * 8 ? rows : (int64_t)kSMs * 8);
        dirichlet_rows_block_kernel<<<grid, 256, 0, s>>>(a, rows, cols, ld, out_elog, out_exp);
    }
    TTLDA_LAUNCH_CHECK("dirichlet_expectation");
    return TTLDA_OK;
}

int ttlda_dirichlet_expectation_1d_f64(double* doc_topic, double prior, double* out, int64_t size, void* stream)
{
    TTLDA_REQUIRE(doc_topic && out, "NULL pointer");
    TTLDA_REQUIRE(size >= 0, "negative size");
    if (size == 0) return TTLDA_OK;
    dirichlet_1d_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(doc_topic, prior, out, size);
    TTLDA_LAUNCH_CHECK("dirichlet_expectation_1d");
    return TTLDA_OK;
}

int ttlda_theta_refresh_f64(const double* gamma, int64_t D, int64_t K, int64_t ld, const double* alpha_vec,
                            double alpha_sum, double* exp_elogtheta, double* partial_out, void* workspace,
                            int64_t workspace_bytes, void* stream)
{
    TTLDA_REQUIRE(gamma && alpha_vec && partial_out && workspace, "NULL pointer");
    TTLDA_REQUIRE(D >= 0 && K >= 1 && ld >= K && ld % 4 == 0, "bad shape");
    if (workspace_bytes < reduce_ws_bytes()) {
        set_error("theta_refresh: workspace %lld < %lld", (long long)workspace_bytes, (long long)reduce_ws_bytes());
        return TTLDA_EWORKSPACE;
    }
    cudaStream_t s = (cudaStream_t)stream;
    const int lpd = K > 256 ? 32 : (K > 128 ? 16 : (K > 32 ? 8 : 4));
    int grid = grid_for_warps((D + 32 / lpd - 1) / (32 / lpd), 8, 16);
    double* rec = (double*)workspace;
    if (lpd == 32) theta_refresh_kernel<32><<<grid, 256, 0, s>>>(gamma, D, (int)K, ld, alpha_vec, alpha_sum, exp_elogtheta, rec);
    else if (lpd == 16) theta_refresh_kernel<16><<<grid, 256, 0, s>>>(gamma, D, (int)K, ld, alpha_vec, alpha_sum, exp_elogtheta, rec);
    else if (lpd == 8) theta_refresh_kernel<8><<<grid, 256, 0, s>>>(gamma, D, (int)K, ld, alpha_vec, alpha_sum, exp_elogtheta, rec);
    else theta_refresh_kernel<4><<<grid, 256, 0, s>>>(gamma, D, (int)K, ld, alpha_vec, alpha_sum, exp_elogtheta, rec);
    TTLDA_LAUNCH_CHECK("theta_refresh");
    return launch_reduce_records((const double*)workspace, grid, partial_out, s);
}

int ttlda_hyper_stats_f64(const double* x, int64_t rows, int64_t cols, int64_t ld, double* partial_out,
                          void* workspace, int64_t workspace_bytes, void* stream)
{
    TTLDA_REQUIRE(x && partial_out && workspace, "NULL pointer");
    TTLDA_REQUIRE(rows >= 0 && cols >= 0 && ld >= cols, "bad shape");
    if (workspace_bytes < reduce_ws_bytes()) {
        set_error("hyper_stats: workspace too small");
        return TTLDA_EWORKSPACE;
    }
    cudaStream_t s = (cudaStream_t)stream;
    int grid = grid_for_warps(rows, 8, 8);
    hyper_stats_kernel<<<grid, 256, 0, s>>>(x, rows, cols, ld, (double*)workspace);
    TTLDA_LAUNCH_CHECK("hyper_stats");
    return launch_reduce_records((const double*)workspace, grid, partial_out, s);
}

int64_t ttlda_hyper_colstats_workspace_bytes(int64_t rows, int64_t cols)
{
    if (rows < 0 || cols < 1) return TTLDA_EINVAL;
    return (int64_t)grid_for_warps(rows, 8, 8) * cols * (int64_t)sizeof(double) + 256;
}

int ttlda_hyper_colstats_f64(const double* x, int64_t rows, int64_t cols, int64_t ld, double* out, void* workspace,
                             int64_t workspace_bytes, void* stream)
{
    TTLDA_REQUIRE(x && out && workspace, "NULL pointer");
    TTLDA_REQUIRE(rows >= 0 && cols >= 1 && cols <= TTLDA_MAX_K && ld >= cols, "bad shape");
    if (workspace_bytes < ttlda_hyper_colstats_workspace_bytes(rows, cols)) {
        set_error("hyper_colstats: workspace too small");
        return TTLDA_EWORKSPACE;
    }
    cudaStream_t s = (cudaStream_t)stream;
    const int grid = grid_for_warps(rows, 8, 8);
    hyper_colstats_kernel<<<grid, 256, 8 * cols * sizeof(double), s>>>(x, rows, (int)cols, ld, (double*)workspace);
    TTLDA_LAUNCH_CHECK("hyper_colstats");
    hyper_colstats_finish_kernel<<<(int)((cols + 255) / 256), 256, 0, s>>>((const double*)workspace, grid, (int)cols, out);
    TTLDA_LAUNCH_CHECK("hyper_colstats_finish");
    return TTLDA_OK;
}

int ttlda_psi_f64(const double* x, int64_t n, double* out, void* stream)
{
    TTLDA_REQUIRE(x && out, "NULL pointer");
    if (n <= 0) return TTLDA_OK;
    int grid = (int)((n + 255) / 256 < (int64_t)kSMs * 8 ? (n + 255) / 256 : (int64_t)kSMs * 8);
    psi_exact_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, n, out);
    TTLDA_LAUNCH_CHECK("psi");
    return TTLDA_OK;
}

}  // extern "C"
