// mstep.cu — M-step on the word-major tables: lambda = eta + sstats * exp(E[log beta]); Dirichlet expectation of the
// new lambda over the vocabulary; exp of it; beta-prior ELBO term.
//
// Reference: `lambda_update *= np.exp(expec_log_beta)` tuotuo/lda_model.py:264; LDASmoothed.m_step :273-283;
// _dirichlet_expectation_2d tuotuo/cutils.pyx:41-68 (row k of the reference's (K,V) lambda is column k here);
// expec_real_dist_minus_suro_dist(Elogbeta, eta, lambda) :25-46 with call site :157-163.
//
// Three HBM-streaming kernels over the [V][ld] tables (K*V*8 bytes each, coalesced, read/written once):
//   1. mstep_lambda_kernel  : lambda = eta + S*eB, per-CTA column partial sums
//   2. colsum_finish_kernel : column sums in fixed CTA order, psi~(colsum_k), sum_k (lgamma(eta_sum) - lgamma(colsum_k))
//   3. beta_refresh_kernel  : Elogbeta = psi~(lambda) - psi~(colsum), eB = exp(Elogbeta), beta-prior term records
#include "ttlda_common.cuh"

namespace ttlda {

constexpr int kColBlocks = 592;  // 148 SMs x 4: CTAs of the column-sum pass (fixed => reproducible partial order)
constexpr int kMaxJ = 8;         // columns per thread: ld <= TX * kMaxJ with TX >= 128 for ld > 64

// blockDim = (TX, 256/TX).  Thread (tx,ty) owns columns k = tx + TX*j and words w0+ty, w0+ty+TY, ...
// MODE 0: column sums of an existing lambda;  1: lambda = eta + S*eB (batch M-step);  2: online blend
// lambda = (1 - rho) * lambda + rho * (eta + scale * S*eB)   (Hoffman, Blei & Bach 2010, eq. 7-9)
template <int MODE>
__global__ void __launch_bounds__(256) mstep_lambda_kernel(const double* __restrict__ S, const double* __restrict__ eB,
                                                           double* __restrict__ lambda, int64_t V, int K, int64_t ld,
                                                           double eta, double scale, double rho,
                                                           double* __restrict__ colpart)
{
    extern __shared__ double sm_cols[];  // [TY][ld]
    const int TX = blockDim.x, TY = blockDim.y, tx = threadIdx.x, ty = threadIdx.y;
    const int64_t per = (V + gridDim.x - 1) / gridDim.x;
    const int64_t w0 = (int64_t)blockIdx.x * per, w1 = (w0 + per < V) ? w0 + per : V;
    double cs[kMaxJ];
#pragma unroll
    for (int j = 0; j < kMaxJ; ++j) cs[j] = 0.;
    for (int64_t w = w0 + ty; w < w1; w += TY) {
#pragma unroll
        for (int j = 0; j < kMaxJ; ++j) {
            const int k = tx + TX * j;
            if (k < (int)ld) {
                const int64_t idx = w * ld + k;
                double lam;
                if (MODE == 1) {
                    lam = (k < K) ? eta + S[idx] * eB[idx] : 0.;  // lda_model.py:264 then :280
                    lambda[idx] = lam;
                } else if (MODE == 2) {
                    lam = (k < K) ? fma(rho, fma(scale, S[idx] * eB[idx], eta), (1. - rho) * lambda[idx]) : 0.;
                    lambda[idx] = lam;
                } else {
                    lam = (k < K) ? lambda[idx] : 0.;
                }
                cs[j] += lam;
            }
        }
    }
#pragma unroll
    for (int j = 0; j < kMaxJ; ++j) {
        const int k = tx + TX * j;
        if (k < (int)ld) sm_cols[ty * ld + k] = cs[j];
    }
    __syncthreads();
    if (ty == 0) {
#pragma unroll
        for (int j = 0; j < kMaxJ; ++j) {
            const int k = tx + TX * j;
            if (k < (int)ld) {
                double s = 0.;
                for (int y = 0; y < TY; ++y) s += sm_cols[y * ld + k];
                colpart[(int64_t)blockIdx.x * ld + k] = s;
            }
        }
    }
}

// colsum[k], psi~(colsum[k]) and the third ELBO term; single CTA.
__global__ void colsum_finish_kernel(const double* __restrict__ colpart, int nblocks, int K, int64_t ld,
                                     double eta_sum, double* __restrict__ colsum, double* __restrict__ psi_colsum,
                                     double* __restrict__ extra)
{
    __shared__ double sm[32];
    double t[1] = {0.};
    const double lg_eta_sum = lgamma(eta_sum);
    for (int k = threadIdx.x; k < (int)ld; k += blockDim.x) {
        double s = 0.;
        for (int b = 0; b < nblocks; ++b) s += colpart[(int64_t)b * ld + k];
        colsum[k] = s;
        if (k < K) {
            psi_colsum[k] = psi_as103(s);  // cutils.pyx:63
            t[0] += lg_eta_sum - lgamma(s);  // lda_model.py:42-44
        } else {
            psi_colsum[k] = 0.;
        }
    }
    block_sum<1>(t, sm);
    if (threadIdx.x == 0) extra[0] = t[0];
}

__global__ void __launch_bounds__(256) beta_refresh_kernel(const double* __restrict__ lambda, int64_t V, int K,
                                                           int64_t ld, double eta, const double* __restrict__ psi_colsum,
                                                           const double* __restrict__ extra, double* __restrict__ eB,
                                                           double* __restrict__ elogbeta, double* __restrict__ records)
{
    __shared__ double sm[32];
    const int64_t total = V * ld;
    const double lg_eta = lgamma(eta);
    double t[1] = {0.};
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int k = (int)(i % ld);
        double e = 0., x = 0.;
        if (k < K) {
            const double lam = lambda[i];
            double ps, lg;
            psi_lgamma_as103(lam, ps, lg);
            e = ps - psi_colsum[k];                    // cutils.pyx:65-66
            x = exp(e);
            t[0] += (eta - lam) * e + (lg - lg_eta);  // lda_model.py:40-41
        }
        eB[i] = x;
        if (elogbeta) elogbeta[i] = e;
    }
    block_sum<1>(t, sm);
    if (threadIdx.x == 0) {
        double* rec = records + (size_t)blockIdx.x * NPART;
        rec[0] = t[0] + (blockIdx.x == 0 ? extra[0] : 0.);
#pragma unroll
        for (int j = 1; j < NPART; ++j) rec[j] = 0.;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Sharded M-step (documents AND the vocabulary are split over the ranks): rank r owns the rows [w0, w0 + n) of lambda.
// The two kernels below are the per-slice halves of the three above, with the cross-rank exchanges FUSED into them when
// the statistics / exp(E[log beta]) tables live in peer-mapped memory (NVLink): phase A sums the `nsrc` ranks' statistics
// rows with plain loads through the peer pointers, in rank order (the reduce-scatter, bit-identical on every rank);
// phase B stores every new exp(E[log beta]) row into all `ndst` ranks' tables (the all-gather).  With nsrc = ndst = 1
// they run on buffers a library collective has reduced / will gather.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kMaxPeers = 8;  // ranks of one NVLink domain (one node)
struct PeerTables {
    const double* src[kMaxPeers];
    double* dst[kMaxPeers];
    const double* src_mc;  // multicast address of the statistics tables (NVSwitch), or NULL
    double* dst_mc;        // multicast address of the exp(E[log beta]) tables, or NULL
};

// NVLink SHARP through the switch's multicast objects: ONE load returns the sum of the same address in every rank's
// table (the reduction happens inside the NVSwitch: this rank's link carries the slice once, not once per peer), ONE
// store lands in every rank's table.
__device__ __forceinline__ double multimem_sum_f64(const double* mc)
{
    double v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.f64 %0, [%1];" : "=d"(v) : "l"(mc) : "memory");
    return v;
}

__device__ __forceinline__ void multimem_store_f64(double* mc, double v)
{
    asm volatile("multimem.st.relaxed.sys.global.f64 [%0], %1;" ::"l"(mc), "d"(v) : "memory");
}

__global__ void __launch_bounds__(256) mstep_slice_lambda_kernel(const PeerTables T, int nsrc, int64_t src_row0,
                                                                 int64_t w0, int64_t n, int K, int64_t ld, double eta,
                                                                 const double* __restrict__ eB,
                                                                 double* __restrict__ lambda,
                                                                 double* __restrict__ colpart)
{
    extern __shared__ double sm_cols[];  // [TY][ld]
    const int TX = blockDim.x, TY = blockDim.y, tx = threadIdx.x, ty = threadIdx.y;
    const int64_t per = (n + gridDim.x - 1) / gridDim.x;
    const int64_t b0 = (int64_t)blockIdx.x * per, b1 = (b0 + per < n) ? b0 + per : n;
    double cs[kMaxJ];
#pragma unroll
    for (int j = 0; j < kMaxJ; ++j) cs[j] = 0.;
    // kRowIlp rows per trip: the cross-rank loads are NVLink round trips (~2 us), so each thread keeps
    // kRowIlp * (columns per thread) * (1 multicast | nsrc peer) of them in flight before it touches the first result
    constexpr int kRowIlp = 4;
    for (int64_t wl0 = b0 + ty; wl0 < b1; wl0 += (int64_t)TY * kRowIlp) {
        double sum[kRowIlp][kMaxJ];
#pragma unroll
        for (int u = 0; u < kRowIlp; ++u) {
            const int64_t wl = wl0 + (int64_t)u * TY;
#pragma unroll
            for (int j = 0; j < kMaxJ; ++j) {
                const int k = tx + TX * j;
                sum[u][j] = 0.;
                if (wl < b1 && k < K) {
                    const int64_t is = (src_row0 + wl) * ld + k;
                    if (T.src_mc) {
                        sum[u][j] = multimem_sum_f64(T.src_mc + is);  // statistics of all ranks, summed in the switch
                    } else {
                        // ... or by plain loads through the peer pointers, in rank order (lda_model.py:262 over shards)
                        double v[kMaxPeers];
#pragma unroll
                        for (int i = 0; i < kMaxPeers; ++i) v[i] = (i < nsrc) ? T.src[i][is] : 0.;
                        double a = v[0];
#pragma unroll
                        for (int i = 1; i < kMaxPeers; ++i) a += v[i];  // + 0.0 for i >= nsrc: exact
                        sum[u][j] = a;
                    }
                }
            }
        }
#pragma unroll
        for (int u = 0; u < kRowIlp; ++u) {
            const int64_t wl = wl0 + (int64_t)u * TY;
            if (wl >= b1) continue;
#pragma unroll
            for (int j = 0; j < kMaxJ; ++j) {
                const int k = tx + TX * j;
                if (k < (int)ld) {
                    const int64_t idx = (w0 + wl) * ld + k;
                    const double lam = (k < K) ? eta + sum[u][j] * eB[idx] : 0.;  // lda_model.py:264 then :280
                    lambda[idx] = lam;
                    cs[j] += lam;
                }
            }
        }
    }
#pragma unroll
    for (int j = 0; j < kMaxJ; ++j) {
        const int k = tx + TX * j;
        if (k < (int)ld) sm_cols[ty * ld + k] = cs[j];
    }
    __syncthreads();
    if (ty == 0) {
#pragma unroll
        for (int j = 0; j < kMaxJ; ++j) {
            const int k = tx + TX * j;
            if (k < (int)ld) {
                double s = 0.;
                for (int y = 0; y < TY; ++y) s += sm_cols[y * ld + k];
                colpart[(int64_t)blockIdx.x * ld + k] = s;
            }
        }
    }
}

// column sums of the slice in fixed CTA order
__global__ void colsum_slice_kernel(const double* __restrict__ colpart, int nblocks, int64_t ld, double* __restrict__ out)
{
    for (int k = threadIdx.x; k < (int)ld; k += blockDim.x) {
        double s = 0.;
        for (int b = 0; b < nblocks; ++b) s += colpart[(int64_t)b * ld + k];
        out[k] = s;
    }
}

__global__ void __launch_bounds__(256) beta_slice_kernel(const double* __restrict__ lambda, int64_t w0, int64_t n, int K,
                                                         int64_t ld, double eta, const double* __restrict__ psi_colsum,
                                                         const double* __restrict__ extra, int add_extra,
                                                         const PeerTables T, int ndst, double* __restrict__ records)
{
    __shared__ double sm[32];
    const int64_t total = n * ld, base = w0 * ld;
    const double lg_eta = lgamma(eta);
    double t[1] = {0.};
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int k = (int)(i % ld);
        double x = 0.;
        if (k < K) {
            const double lam = lambda[base + i];
            double ps, lg;
            psi_lgamma_as103(lam, ps, lg);
            const double e = ps - psi_colsum[k];      // cutils.pyx:65-66
            x = exp(e);
            t[0] += (eta - lam) * e + (lg - lg_eta);  // lda_model.py:40-41
        }
        if (T.dst_mc) {
            multimem_store_f64(T.dst_mc + base + i, x);  // one store, replicated by the switch into every rank's table
        } else {
            for (int d = 0; d < ndst; ++d) T.dst[d][base + i] = x;  // this rank's table and, over NVLink, every peer's
        }
    }
    block_sum<1>(t, sm);
    if (threadIdx.x == 0) {
        double* rec = records + (size_t)blockIdx.x * NPART;
        rec[0] = t[0] + ((blockIdx.x == 0 && add_extra) ? extra[0] : 0.);
#pragma unroll
        for (int j = 1; j < NPART; ++j) rec[j] = 0.;
    }
}

static size_t align256(size_t x) { return (x + 255) / 256 * 256; }

static int64_t mstep_ws_bytes(int64_t ld)
{
    return (int64_t)(align256((size_t)reduce_ws_bytes()) + align256((size_t)kColBlocks * ld * sizeof(double)) +
                     align256((size_t)ld * sizeof(double)) + 256);
}

template <int MODE>
static int run_mstep(const double* S, int64_t V, int64_t K, int64_t ld, double eta, double scale, double rho, double* eB,
                     double* lambda,
                     double* elogbeta, double* colsum_out, double* partial_out, void* workspace,
                     int64_t workspace_bytes, cudaStream_t s)
{
    if (workspace_bytes < mstep_ws_bytes(ld)) {
        set_error("mstep: workspace %lld < %lld", (long long)workspace_bytes, (long long)mstep_ws_bytes(ld));
        return TTLDA_EWORKSPACE;
    }
    char* ws = (char*)workspace;
    double* records = (double*)ws;
    double* colpart = (double*)(ws + align256((size_t)reduce_ws_bytes()));
    double* psi_colsum = (double*)((char*)colpart + align256((size_t)kColBlocks * ld * sizeof(double)));
    double* extra = (double*)((char*)psi_colsum + align256((size_t)ld * sizeof(double)));

    int TX = 32;
    while (TX < 256 && TX < ld) TX *= 2;
    if (ld > 64 && TX < 128) TX = 128;
    const int TY = 256 / TX;
    int nblocks = (int)(V < kColBlocks ? (V > 0 ? V : 1) : kColBlocks);
    const size_t smem = (size_t)TY * ld * sizeof(double);
    mstep_lambda_kernel<MODE><<<nblocks, dim3(TX, TY), smem, s>>>(S, eB, lambda, V, (int)K, ld, eta, scale, rho,
                                                                   colpart);
    TTLDA_LAUNCH_CHECK("mstep_lambda");
    // eta is a scalar prior: np.sum(eta) is eta itself
    colsum_finish_kernel<<<1, 256, 0, s>>>(colpart, nblocks, (int)K, ld, eta, colsum_out, psi_colsum, extra);
    TTLDA_LAUNCH_CHECK("colsum_finish");
    const int64_t total = V * ld;
    int grid = (int)((total + 255) / 256 < (int64_t)kSMs * 8 ? (total + 255) / 256 : (int64_t)kSMs * 8);
    if (grid < 1) grid = 1;
    beta_refresh_kernel<<<grid, 256, 0, s>>>(lambda, V, (int)K, ld, eta, psi_colsum, extra, eB, elogbeta, records);
    TTLDA_LAUNCH_CHECK("beta_refresh");
    return launch_reduce_records(records, grid, partial_out, s);
}

}  // namespace ttlda

using namespace ttlda;

extern "C" {

int64_t ttlda_mstep_workspace_bytes(int64_t V, int64_t ld)
{
    (void)V;
    return mstep_ws_bytes(ld);
}

int ttlda_mstep_f64(const double* sstats, int64_t V, int64_t K, int64_t ld, double eta, double* exp_elogbeta,
                    double* lambda, double* elogbeta_out, double* colsum_out, double* partial_out, void* workspace,
                    int64_t workspace_bytes, void* stream)
{
    TTLDA_REQUIRE(sstats && exp_elogbeta && lambda && colsum_out && partial_out && workspace, "NULL pointer");
    TTLDA_REQUIRE(V >= 0 && K >= 1 && K <= TTLDA_MAX_K && ld >= K && ld % 4 == 0, "bad shape");
    return run_mstep<1>(sstats, V, K, ld, eta, 1., 1., exp_elogbeta, lambda, elogbeta_out, colsum_out, partial_out,
                        workspace, workspace_bytes, (cudaStream_t)stream);
}

int ttlda_mstep_online_f64(const double* sstats, int64_t V, int64_t K, int64_t ld, double eta, double scale, double rho,
                           double* exp_elogbeta, double* lambda, double* elogbeta_out, double* colsum_out,
                           double* partial_out, void* workspace, int64_t workspace_bytes, void* stream)
{
    TTLDA_REQUIRE(sstats && exp_elogbeta && lambda && colsum_out && partial_out && workspace, "NULL pointer");
    TTLDA_REQUIRE(V >= 0 && K >= 1 && K <= TTLDA_MAX_K && ld >= K && ld % 4 == 0, "bad shape");
    TTLDA_REQUIRE(rho > 0. && rho <= 1. && scale > 0., "need 0 < rho <= 1 and scale > 0");
    return run_mstep<2>(sstats, V, K, ld, eta, scale, rho, exp_elogbeta, lambda, elogbeta_out, colsum_out, partial_out,
                        workspace, workspace_bytes, (cudaStream_t)stream);
}

int ttlda_mstep_slice_lambda_f64(const double* const* src_tables, int64_t nsrc, const double* src_multicast,
                                 int64_t src_row0, int64_t w0, int64_t n,
                                 int64_t K, int64_t ld, double eta, const double* exp_elogbeta, double* lambda,
                                 double* colsum_part_out, void* workspace, int64_t workspace_bytes, void* stream)
{
    TTLDA_REQUIRE(src_tables && exp_elogbeta && lambda && colsum_part_out && workspace, "NULL pointer");
    TTLDA_REQUIRE(nsrc >= 1 && nsrc <= kMaxPeers, "1 <= nsrc <= 8");
    TTLDA_REQUIRE(w0 >= 0 && n >= 0 && src_row0 >= 0 && K >= 1 && K <= TTLDA_MAX_K && ld >= K && ld % 4 == 0, "bad shape");
    if (workspace_bytes < mstep_ws_bytes(ld)) {
        set_error("mstep_slice_lambda: workspace %lld < %lld", (long long)workspace_bytes, (long long)mstep_ws_bytes(ld));
        return TTLDA_EWORKSPACE;
    }
    cudaStream_t s = (cudaStream_t)stream;
    PeerTables T{};
    for (int i = 0; i < nsrc; ++i) {
        TTLDA_REQUIRE(src_tables[i], "NULL source table");
        T.src[i] = src_tables[i];
    }
    T.src_mc = src_multicast;
    double* colpart = (double*)((char*)workspace + align256((size_t)reduce_ws_bytes()));
    int TX = 32;
    while (TX < 256 && TX < ld) TX *= 2;
    if (ld > 64 && TX < 128) TX = 128;
    const int TY = 256 / TX;
    const int nblocks = (int)(n < kColBlocks ? (n > 0 ? n : 1) : kColBlocks);
    mstep_slice_lambda_kernel<<<nblocks, dim3(TX, TY), (size_t)TY * ld * sizeof(double), s>>>(
        T, (int)nsrc, src_row0, w0, n, (int)K, ld, eta, exp_elogbeta, lambda, colpart);
    TTLDA_LAUNCH_CHECK("mstep_slice_lambda");
    colsum_slice_kernel<<<1, 256, 0, s>>>(colpart, nblocks, ld, colsum_part_out);
    TTLDA_LAUNCH_CHECK("colsum_slice");
    return TTLDA_OK;
}

int ttlda_mstep_slice_beta_f64(const double* colsum_parts, int64_t nparts, int64_t w0, int64_t n, int64_t K, int64_t ld,
                               double eta, const double* lambda, double* const* dst_tables, int64_t ndst,
                               double* dst_multicast, int add_colsum_term, double* colsum_out, double* partial_out,
                               void* workspace, int64_t workspace_bytes, void* stream)
{
    TTLDA_REQUIRE(colsum_parts && lambda && dst_tables && colsum_out && partial_out && workspace, "NULL pointer");
    TTLDA_REQUIRE(nparts >= 1 && nparts <= kColBlocks && ndst >= 1 && ndst <= kMaxPeers, "bad rank counts");
    TTLDA_REQUIRE(w0 >= 0 && n >= 0 && K >= 1 && K <= TTLDA_MAX_K && ld >= K && ld % 4 == 0, "bad shape");
    if (workspace_bytes < mstep_ws_bytes(ld)) {
        set_error("mstep_slice_beta: workspace %lld < %lld", (long long)workspace_bytes, (long long)mstep_ws_bytes(ld));
        return TTLDA_EWORKSPACE;
    }
    cudaStream_t s = (cudaStream_t)stream;
    PeerTables T{};
    for (int i = 0; i < ndst; ++i) {
        TTLDA_REQUIRE(dst_tables[i], "NULL destination table");
        T.dst[i] = dst_tables[i];
    }
    T.dst_mc = dst_multicast;
    char* ws = (char*)workspace;
    double* records = (double*)ws;
    double* psi_colsum = (double*)(ws + align256((size_t)reduce_ws_bytes()) + align256((size_t)kColBlocks * ld * sizeof(double)));
    double* extra = (double*)((char*)psi_colsum + align256((size_t)ld * sizeof(double)));
    // global column sums = the ranks' partials in rank order (identical bits everywhere), psi~ of them, colsum-only term
    colsum_finish_kernel<<<1, 256, 0, s>>>(colsum_parts, (int)nparts, (int)K, ld, eta, colsum_out, psi_colsum, extra);
    TTLDA_LAUNCH_CHECK("colsum_finish");
    const int64_t total = n * ld;
    int grid = (int)((total + 255) / 256 < (int64_t)kSMs * 8 ? (total + 255) / 256 : (int64_t)kSMs * 8);
    if (grid < 1) grid = 1;
    beta_slice_kernel<<<grid, 256, 0, s>>>(lambda, w0, n, (int)K, ld, eta, psi_colsum, extra, add_colsum_term, T,
                                           (int)ndst, records);
    TTLDA_LAUNCH_CHECK("beta_slice");
    return launch_reduce_records(records, grid, partial_out, s);
}

int ttlda_beta_refresh_f64(const double* lambda, int64_t V, int64_t K, int64_t ld, double eta, double* exp_elogbeta,
                           double* elogbeta_out, double* colsum_out, double* partial_out, void* workspace,
                           int64_t workspace_bytes, void* stream)
{
    TTLDA_REQUIRE(lambda && exp_elogbeta && colsum_out && partial_out && workspace, "NULL pointer");
    TTLDA_REQUIRE(V >= 0 && K >= 1 && K <= TTLDA_MAX_K && ld >= K && ld % 4 == 0, "bad shape");
    return run_mstep<0>(nullptr, V, K, ld, eta, 1., 1., exp_elogbeta, const_cast<double*>(lambda), elogbeta_out,
                        colsum_out, partial_out, workspace, workspace_bytes, (cudaStream_t)stream);
}

}  // extern "C"
