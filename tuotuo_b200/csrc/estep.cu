// estep.cu — document-major pass: E-step gamma update fused with the ELBO token term (incoming state), the theta
// refresh and the theta-prior ELBO term (outgoing state); and the stand-alone ELBO token kernel.
//
// Reference: LDASmoothed.e_step tuotuo/lda_model.py:197-270; approx_elbo token loop :125-139; theta-prior term
// :25-46,:142-148; Dirichlet expectation tuotuo/cutils.pyx:41-93.
//
// Mapping (traverse.cuh): one WARP per document, the document's exp(E[log theta_d]) K-vector lives in registers,
// replicated over the warp's R = 32/G lane groups; each group gathers the exp(E[log beta_w]) K-vector of a
// different nonzero of the document from the word-major [V][ld] table, so R nonzeros advance per instruction:
//     p = sum_k eT_k * eB[w,k]            (2E DFMA per lane + log2(G) shuffle steps)
//     r = c / (p + 1e-100)                (lda_model.py:242)
//     acc_k += r * eB[w,k]                (lda_model.py:248, the (cts/norm) . eB^T product; per-group partials)
// The (word, count) pairs are fetched 32 at a time (one coalesced load per lane) and routed to the groups by
// shuffle; the next sub-batch's K-vectors are prefetched into registers.  p of every nonzero is handed back to
// the lane that loaded it, so the token term costs ONE log per nonzero per lane.  Per-document epilogue: group
// partials are transposed through shared memory to a 32-lane layout (k = lane + 32m), then
// gamma_k = alpha_k + eT_k * acc_k, AS-103 psi, exp, lgamma => new exp(E[log theta_d]) and the theta-prior term.
// Bound: gather of 8K bytes per nonzero (L2 when the 8*K*V-byte table fits, else HBM).
#include "tma_gather.cuh"

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
    const double* gamma_in;  // previous gamma (only read for the loop statistics), may be NULL
    double* gamma_out;
    double* eT_out;
    const uint32_t* csr2csc;  // mirror position of every CSR position, or NULL
    double* ratio_csc;        // c/N per nonzero in mirror order (for the statistics pass), or NULL
    int64_t max_iter;
    double tol;
    unsigned long long* iters;  // [2] device counters (passes, capped docs) or NULL
    double* records;
};

// (TTLDA_DOC_THREADS / TTLDA_DOC_MINB: same-box A/B builds of the CTA shape, tools/build_variant.sh)
#ifndef TTLDA_DOC_THREADS
#define TTLDA_DOC_THREADS 128
#endif
#ifndef TTLDA_DOC_MINB
#define TTLDA_DOC_MINB 3
#endif
constexpr int kDocThreads = TTLDA_DOC_THREADS;  // 4 warps per CTA

// Per-document epilogue shared by the LDG- and TMA-fed kernels: group partials -> 32-lane layout through shared
// memory, gamma_d = alpha + eT_d * acc (lda_model.py:248), the reference's loop statistics, and (MODE 1) the theta
// refresh + theta-prior ELBO term.
template <int G, int E, int MODE>
__device__ __forceinline__ void doc_epilogue(const EStepParams& P, int64_t d, const KSlice<G, E>& acc, double* wbuf,
                                             int lane, int g, int r, uint64_t pol, double lga, double lg_asum,
                                             double& theta_acc, unsigned long long& passes,
                                             unsigned long long& capped)
{
    constexpr bool REFRESH = MODE == 1;
    constexpr int R = 32 / G;
    constexpr int LDP = 2 * G * E;
    constexpr int MMAX = (LDP + 31) / 32;
    const int ld = (int)P.ld, K = P.K;
    // group partials -> 32-lane layout through shared memory
    acc.to_smem(wbuf + r * LDP, g);
    __syncwarp();
    double gn[MMAX];
    double s = 0., l1 = 0.;
#pragma unroll
    for (int m = 0; m < MMAX; ++m) {
        const int k = lane + 32 * m;
        gn[m] = 0.;
        if (k < K) {
            double a = 0.;
#pragma unroll
            for (int q = 0; q < R; ++q) a += wbuf[q * LDP + k];
            gn[m] = P.alpha_vec[k] + ld_stream_f64(P.eT_in + d * P.ld + k, pol) * a;  // lda_model.py:248
            s += gn[m];
            if (P.iters) l1 += fabs(gn[m] - ld_stream_f64(P.gamma_in + d * P.ld + k, pol));
        }
    }
    __syncwarp();
    s = warp_sum(s);
    if (P.iters) {
        l1 = warp_sum(l1);
        // the reference's while-loop (lda_model.py:236): pass 1 always runs; pass 2 runs iff the first
        // moved gamma by more than tol and the cap allows it, and then reproduces pass 1 (delta = 0).
        const unsigned long long np = (l1 > P.tol && 1 <= P.max_iter) ? 2ull : 1ull;
        passes += np;
        capped += ((int64_t)np > P.max_iter) ? 1ull : 0ull;
    }
    if (REFRESH) {
        double pt, lg_s;
        psi_lgamma_as103(s, pt, lg_s);
        double th = 0.;
#pragma unroll
        for (int m = 0; m < MMAX; ++m) {
            const int k = lane + 32 * m;
            if (k < K) {
                double ps, lg;
                psi_lgamma_as103(gn[m], ps, lg);
                const double el = ps - pt;
                th += (P.alpha_vec[k] - gn[m]) * el + lg;
                st_stream_f64(P.gamma_out + d * P.ld + k, gn[m], pol);
                st_stream_f64(P.eT_out + d * P.ld + k, exp(el), pol);
            }
        }
        if (lane < ld - K) {  // padding columns
            P.gamma_out[d * P.ld + K + lane] = 0.;
            P.eT_out[d * P.ld + K + lane] = 0.;
        }
        th = warp_sum(th);
        theta_acc += th - lga + (lg_asum - lg_s);
    } else {
#pragma unroll
        for (int m = 0; m < MMAX; ++m) {
            const int k = lane + 32 * m;
            if (k < K) st_stream_f64(P.gamma_out + d * P.ld + k, gn[m], pol);
        }
        if (lane < ld - K) P.gamma_out[d * P.ld + K + lane] = 0.;
    }
}

// MODE 0: ELBO token term only;  1: E-step with the theta refresh + theta-prior term fused into the per-document
// epilogue;  2: E-step writing gamma only (the refresh is then run as the separate, fully parallel theta_refresh pass:
// the epilogue's digamma/log-gamma/exp chains otherwise sit on the critical path of a warp that has nothing else to do)
template <int G, int E, bool PREFETCH, int MODE>
#ifdef TTLDA_DOC_MAXNREG
__global__ void __maxnreg__(TTLDA_DOC_MAXNREG) doc_pass_kernel(const EStepParams P)
#else
__global__ void __launch_bounds__(kDocThreads, (E <= 8 ? TTLDA_DOC_MINB : 1)) doc_pass_kernel(const EStepParams P)
#endif
{
    constexpr bool ESTEP = MODE > 0, REFRESH = MODE == 1;
    constexpr int R = 32 / G;
    constexpr int LDP = 2 * G * E;
    constexpr int MMAX = (LDP + 31) / 32;
    extern __shared__ double smem[];
    double* red = smem;  // 3*32 doubles for the CTA reduction
    double* wbuf = smem + 96 + (threadIdx.x >> 5) * (R * LDP);

    const int lane = threadIdx.x & 31;
    const int g = lane % G, r = lane / G;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int ld = (int)P.ld, K = P.K;
    const bool last_ok = KSlice<G, E>::last_slice_ok(g, ld);
    const uint64_t pol = l2_evict_first_policy();

    double lga = 0., lg_asum = 0.;
    if (REFRESH) {
        for (int k = lane; k < K; k += 32) lga += lgamma(P.alpha_vec[k]);
        lga = warp_sum(lga);
        lg_asum = lgamma(P.alpha_sum);
    }
    double tok_acc = 0., theta_acc = 0., cnt_acc = 0.;
    unsigned long long passes = 0, capped = 0;

    for (int64_t d = warp0; d < P.D; d += nwarps) {
        const int64_t b = P.rowptr[d], e = P.rowptr[d + 1];
        KSlice<G, E> t, acc;
        t.load(P.eT_in + d * P.ld, g, last_ok);
        acc.zero();

        // (word, count) pairs: 32 per coalesced load, fetched ONE batch ahead so that their latency is hidden
        int nx_w = 0, nx_c = 0, nx_pos = 0;
        if (b + lane < e) {
            nx_w = ld_stream_s32(P.colidx + b + lane, pol);
            nx_c = ld_stream_s32(P.counts + b + lane, pol);
            if (ESTEP && P.ratio_csc) nx_pos = ld_stream_s32(reinterpret_cast<const int32_t*>(P.csr2csc) + b + lane, pol);
        }
        KSlice<G, E> rowA;  // sub-batch 0 of the coming batch (carried across batches when PREFETCH)
        if (PREFETCH && b < e) rowA.load_at(row_at(P.eB + 2 * g, __shfl_sync(0xffffffffu, nx_w, r), ld), last_ok);
        for (int64_t n0 = b; n0 < e; n0 += 32) {
            const int cnt = (int)((e - n0 < 32) ? (e - n0) : 32);
            const int my_w = nx_w, my_c = nx_c, my_pos = nx_pos;
            nx_w = 0;
            nx_c = 0;
            if (n0 + 32 + lane < e) {
                nx_w = ld_stream_s32(P.colidx + n0 + 32 + lane, pol);
                nx_c = ld_stream_s32(P.counts + n0 + 32 + lane, pol);
                if (ESTEP && P.ratio_csc)
                    nx_pos = ld_stream_s32(reinterpret_cast<const int32_t*>(P.csr2csc) + n0 + 32 + lane, pol);
            }
            double my_p = 1.;
            auto body = [&](int s, const KSlice<G, E>& row) {
                const int i = s * R + r;                          // nonzero handled by this group in this sub-batch
                const int c = __shfl_sync(0xffffffffu, my_c, i);  // 0 for i >= cnt: the group idles harmlessly
                // N = theta.beta + 1e-100 (lda_model.py:242): the 1e-100 enters as the start value of one DFMA chain
                // (1e-100 / G per lane is exact, G a power of two) instead of as one more dependent DADD
                const double p = group_sum<G>(t.dot(row, ESTEP ? 1e-100 / G : 0.));
                if (ESTEP) acc.axpy(fast_div_pos((double)c, p), row);
                // hand p back to the lane that loaded nonzero `lane` (group lane % R of sub-batch lane / R)
                const double pv = __shfl_sync(0xffffffffu, p, (lane % R) * G);
                if (lane / R == s) my_p = pv;
            };
            if (PREFETCH) walk_rows_carried<G, E>(P.eB, ld, my_w, cnt, nx_w, n0 + 32 < e, rowA, g, r, last_ok, body);
            else walk_rows<G, E, false>(P.eB, ld, my_w, cnt, g, r, last_ok, body);
            if (lane < cnt) {  // one log per nonzero (of theta.beta itself: the token term has no 1e-100, :125-139)
                tok_acc += (double)my_c * log(ESTEP ? my_p - 1e-100 : my_p);
                cnt_acc += (double)my_c;
                // leave c/N for the statistics pass, in mirror (word-major) order: one scattered 8-byte store
                if (ESTEP && P.ratio_csc)
                    st_stream_f64(P.ratio_csc + (uint32_t)my_pos, fast_div_pos((double)my_c, my_p), pol);
            }
        }

        if (ESTEP)
            doc_epilogue<G, E, MODE>(P, d, acc, wbuf, lane, g, r, pol, lga, lg_asum, theta_acc, passes, capped);
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
    if (ESTEP && P.iters && lane == 0 && (passes | capped)) {
        atomicAdd(P.iters, passes);
        if (capped) atomicAdd(P.iters + 1, capped);
    }
}

// The same pass fed by the TMA engine (tma_gather.cuh): one gather4 per sub-batch lands the 4 exp(E[log beta]) rows in
// shared memory, kTmaStages sub-batches ahead, continuously across the document's 32-item batches.  G = 8 only.
template <int E, int MODE>
__global__ void __launch_bounds__(kDocThreads, 4) doc_pass_tma_kernel(const __grid_constant__ CUtensorMap map,
                                                                      const EStepParams P)
{
    constexpr bool ESTEP = MODE > 0, REFRESH = MODE == 1;
    constexpr int G = 8, R = 4;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* red = reinterpret_cast<double*>(smem_raw);  // 3*32 doubles for the CTA reduction (768 B)
    unsigned char* wbase = smem_raw + 768 + (size_t)(threadIdx.x >> 5) * Gather4Pipe<E>::bytes_per_warp(P.ld);

    const int lane = threadIdx.x & 31;
    const int g = lane % G, r = lane / G;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int ld = (int)P.ld, K = P.K;
    const bool last_ok = KSlice<G, E>::last_slice_ok(g, ld);
    const uint64_t pol = l2_evict_first_policy();

    Gather4Pipe<E> pipe;
    pipe.init(wbase, ld, lane);
    double* wbuf = reinterpret_cast<double*>(wbase);  // the epilogue's transpose buffer aliases the drained stages

    double lga = 0., lg_asum = 0.;
    if (REFRESH) {
        for (int k = lane; k < K; k += 32) lga += lgamma(P.alpha_vec[k]);
        lga = warp_sum(lga);
        lg_asum = lgamma(P.alpha_sum);
    }
    double tok_acc = 0., theta_acc = 0., cnt_acc = 0.;
    unsigned long long passes = 0, capped = 0;
    const bool handoff = ESTEP && P.ratio_csc != nullptr;

    for (int64_t d = warp0; d < P.D; d += nwarps) {
        const int64_t b = P.rowptr[d], n = P.rowptr[d + 1] - b;
        KSlice<G, E> t, acc;
        t.load(P.eT_in + d * P.ld, g, last_ok);
        acc.zero();
        int c_cur = 0, c_nxt = 0, pos_cur = 0, pos_nxt = 0;  // per-item payload of the current / next 32-item batch
        double my_p = 1.;
        tma_walk<E>(
            pipe, &map, P.colidx + b, n, lane, g, r, last_ok, pol,
            [&](int64_t base) {  // payload of the batch starting at item `base`
                c_nxt = 0;
                pos_nxt = 0;
                if (base + lane < n) {
                    c_nxt = ld_stream_s32(P.counts + b + base + lane, pol);
                    if (handoff) pos_nxt = ld_stream_s32(reinterpret_cast<const int32_t*>(P.csr2csc) + b + base + lane, pol);
                }
            },
            [&]() {
                c_cur = c_nxt;
                pos_cur = pos_nxt;
            },
            [&](int sb, const KSlice<G, E>& row) {
                const int c = __shfl_sync(0xffffffffu, c_cur, sb * R + r);  // 0 past the end: the group idles
                const double p = group_sum<G>(t.dot(row));
                if (ESTEP) acc.axpy(fast_div_pos((double)c, p + 1e-100), row);  // lda_model.py:242,248
                // hand p back to the lane that loaded nonzero `lane` of the batch (group lane % R, sub-batch lane / R)
                const double pv = __shfl_sync(0xffffffffu, p, (lane % R) * G);
                if (lane / R == sb) my_p = pv;
            },
            [&](int cnt) {
                if (lane < cnt) {  // one log per nonzero
                    tok_acc += (double)c_cur * log(my_p);
                    cnt_acc += (double)c_cur;
                    if (handoff) st_stream_f64(P.ratio_csc + (uint32_t)pos_cur, (double)c_cur / (my_p + 1e-100), pol);
                }
            });
        if (ESTEP)
            doc_epilogue<G, E, MODE>(P, d, acc, wbuf, lane, g, r, pol, lga, lg_asum, theta_acc, passes, capped);
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
    if (ESTEP && P.iters && lane == 0 && (passes | capped)) {
        atomicAdd(P.iters, passes);
        if (capped) atomicAdd(P.iters + 1, capped);
    }
}

// PERSISTENT grid: exactly as many CTAs as are resident at once (occupancy x SMs), documents strided over their warps.
// With more CTAs than fit (the usual "16 per SM"), a CTA retires only when its slowest warp is done — the other
// warps' slots idle meanwhile — and every CTA pays its prologue / record epilogue: same-box A/B at cfg3, 16 CTAs/SM
// 15.59 ms, 6/SM 14.78 ms, resident (3/SM) 14.53 ms.
template <class Kernel>
static int resident_grid(Kernel kernel, size_t smem, int64_t D)
{
    static int per_sm = 0;  // one instance per kernel instantiation (Kernel is a distinct function type per shape/mode
                            // only through the caller's template: the query is repeated when the pointer changes)
    static Kernel cached = nullptr;
    if (cached != kernel || per_sm < 1) {
        int n = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, kDocThreads, smem) != cudaSuccess || n < 1) n = 1;
        per_sm = n;
        cached = kernel;
    }
    return grid_for_warps(D, kDocThreads / 32, per_sm);
}

template <int E, int MODE>
static int launch_tma(const EStepParams& P, int* grid_out, int64_t table_rows, cudaStream_t s)
{
    CUtensorMap map;
    int rc = make_row_gather_map(P.eB, table_rows, P.ld, &map);
    if (rc != TTLDA_OK) return rc;
    const size_t smem = 768 + (kDocThreads / 32) * Gather4Pipe<E>::bytes_per_warp(P.ld);
    cudaError_t e = cudaFuncSetAttribute(doc_pass_tma_kernel<E, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(doc_pass_tma)");
    const int grid = resident_grid(doc_pass_tma_kernel<E, MODE>, smem, P.D);
    doc_pass_tma_kernel<E, MODE><<<grid, kDocThreads, smem, s>>>(map, P);
    *grid_out = grid;
    return TTLDA_OK;
}

template <int G, int E, int MODE>
static int launch_shape(const EStepParams& P, int* grid_out, cudaStream_t s)
{
    constexpr int R = 32 / G;
    constexpr size_t smem = (96 + (kDocThreads / 32) * R * 2 * G * E) * sizeof(double);
    constexpr bool PF = (E <= 8);  // register budget: t, acc, cur, nxt = 8E doubles
    const int grid = resident_grid(doc_pass_kernel<G, E, PF, MODE>, smem, P.D);
    doc_pass_kernel<G, E, PF, MODE><<<grid, kDocThreads, smem, s>>>(P);
    *grid_out = grid;
    return TTLDA_OK;
}

template <int MODE>
static int launch_doc_pass(const EStepParams& P, int64_t V, int* grid, cudaStream_t s)
{
    GroupShape gs = group_shape_for(P.ld);
    if (const char* ov = getenv("TTLDA_GROUP")) {  // tuning override: lanes per K-vector
        const int G = atoi(ov);
        if (G == 2 || G == 4 || G == 8 || G == 16 || G == 32) {
            const int E = (int)((P.ld + 2 * G - 1) / (2 * G));
            gs = {G, E};
        }
    }
    if (tma_gather_applicable(gs.G, P.ld) && tma_gather_enabled() && V > 0) {  // rows fed by TMA gather4
        switch (gs.E) {
            case 5: return launch_tma<5, MODE>(P, grid, V, s);
            case 6: return launch_tma<6, MODE>(P, grid, V, s);
            case 7: return launch_tma<7, MODE>(P, grid, V, s);
            case 8: return launch_tma<8, MODE>(P, grid, V, s);
            default: break;
        }
    }
#define TTLDA_CASE(GG, EE) \
    if (gs.G == GG && gs.E == EE) return launch_shape<GG, EE, MODE>(P, grid, s);
    TTLDA_FOR_EACH_SHAPE(TTLDA_CASE)
#undef TTLDA_CASE
    set_error("no kernel instantiated for ld=%lld (G=%d, E=%d)", (long long)P.ld, gs.G, gs.E);
    return TTLDA_EINVAL;
}

}  // namespace ttlda

using namespace ttlda;

extern "C" {

int ttlda_estep_f64(const int64_t* rowptr, const int32_t* colidx, const int32_t* counts, int64_t D, int64_t K,
                    int64_t V, int64_t ld, const double* alpha_vec, double alpha_sum,
                    const double* exp_elogtheta_in, const double* exp_elogbeta, const double* gamma_in,
                    double* gamma_out, double* exp_elogtheta_out, const uint32_t* csr2csc, double* ratio_csc,
                    int64_t max_iter, double tol,
                    int64_t* iters_out, double* partial_out, void* workspace, int64_t workspace_bytes, void* stream)
{
    TTLDA_REQUIRE(rowptr && alpha_vec && exp_elogtheta_in && exp_elogbeta && gamma_out && partial_out && workspace,
                  "NULL pointer");
    TTLDA_REQUIRE(D >= 0 && V >= 0 && K >= 1 && K <= TTLDA_MAX_K, "K must be in [1, 1024]");
    TTLDA_REQUIRE(ld >= K && ld % 4 == 0 && ld < K + 4, "ld must be K rounded up to a multiple of 4");
    TTLDA_REQUIRE(exp_elogtheta_in != exp_elogtheta_out, "exp_elogtheta_out must not alias the input");
    TTLDA_REQUIRE(max_iter >= 0, "max_iter must be >= 0");
    TTLDA_REQUIRE(!iters_out || gamma_in, "iters_out needs gamma_in (the previous gamma)");
    TTLDA_REQUIRE((csr2csc == nullptr) == (ratio_csc == nullptr), "csr2csc and ratio_csc go together");
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
                  gamma_in, gamma_out, exp_elogtheta_out, csr2csc, ratio_csc, max_iter, tol,
                  (unsigned long long*)iters_out, (double*)workspace};
    int grid = 0;
    int rc = exp_elogtheta_out ? launch_doc_pass<1>(P, V, &grid, s) : launch_doc_pass<2>(P, V, &grid, s);
    if (rc != TTLDA_OK) return rc;
    TTLDA_LAUNCH_CHECK("estep");
    return launch_reduce_records((const double*)workspace, grid, partial_out, s);
}

int ttlda_elbo_tokens_f64(const int64_t* rowptr, const int32_t* colidx, const int32_t* counts, int64_t D, int64_t K,
                          int64_t V, int64_t ld, const double* exp_elogtheta, const double* exp_elogbeta, double* partial_out,
                          void* workspace, int64_t workspace_bytes, void* stream)
{
    TTLDA_REQUIRE(rowptr && exp_elogtheta && exp_elogbeta && partial_out && workspace, "NULL pointer");
    TTLDA_REQUIRE(D >= 0 && K >= 1 && K <= TTLDA_MAX_K, "K must be in [1, 1024]");
    TTLDA_REQUIRE(ld >= K && ld % 4 == 0 && ld < K + 4, "ld must be K rounded up to a multiple of 4");
    if (workspace_bytes < reduce_ws_bytes()) {
        set_error("elbo_tokens: workspace too small");
        return TTLDA_EWORKSPACE;
    }
    cudaStream_t s = (cudaStream_t)stream;
    EStepParams P{rowptr, colidx, counts, D, (int)K, ld, nullptr, 0., exp_elogtheta, exp_elogbeta,
                  nullptr, nullptr, nullptr, nullptr, nullptr, 0, 0., nullptr, (double*)workspace};
    int grid = 0;
    int rc = launch_doc_pass<0>(P, V, &grid, s);
    if (rc != TTLDA_OK) return rc;
    TTLDA_LAUNCH_CHECK("elbo_tokens");
    return launch_reduce_records((const double*)workspace, grid, partial_out, s);
}

}  // extern "C"
