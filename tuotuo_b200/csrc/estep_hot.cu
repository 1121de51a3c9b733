// estep_hot.cu — the E-step with the most frequent words' exp(E[log beta_w]) K-vectors held in shared memory.
//
// Reference: LDASmoothed.e_step tuotuo/lda_model.py:197-270 (gamma update :242,:248) and the token loop of approx_elbo
// :125-139 — the same arithmetic as estep.cu's doc_pass_kernel, one warp per document, K topics in registers.
//
// Why: word frequencies are Zipfian, so a few hundred words carry a third of all (document, word) nonzeros (cfg3: the
// 289 most frequent of 50 000 words = 31 % of the nonzeros), yet every nonzero costs one 8K-byte row gather from L2
// (the per-SM L1 catches < 9 % of them: the cold rows stream through it, profiles/r01m).  Here each persistent CTA
// copies the H hottest rows into its shared memory once per launch (H = as many as fit in 227 KB: 289 rows of 800 B);
// the document's nonzeros arrive in an "E-step stream" in which the hot ones come FIRST and carry their cache slot
// instead of their word id (ttlda_build_estream, built once per corpus like the CSC mirror), so a warp runs a
// homogeneous low-latency shared-memory phase (LDS.128) and then the L2 phase (LDG.128) over the remaining rows; which
// of the two a sub-batch takes is warp-uniform.  The traversal is otherwise the one of traverse.cuh.
//
// Also different from doc_pass_kernel: the per-document epilogue stays in registers — the R lane groups' partial
// accumulators are combined by log2(R) butterfly steps (every lane then holds the document's full sums for its own
// topic slices, identical bits), gamma = alpha + eT * acc uses the eT slices the lane already holds, and each lane
// group stores a quarter of the row with 16-byte streaming stores: no shared-memory transpose, no second read of
// exp(E[log theta_d]).
#include "traverse.cuh"

namespace ttlda {

constexpr int kHotThreads = 384;                       // 12 warps: 168 registers per thread, one CTA per SM
constexpr size_t kHotSmemBytes = 227 * 1024;           // opt-in maximum of dynamic shared memory per CTA on sm_100
constexpr int kHotHeaderDoubles = 96;                  // CTA reduction scratch in front of the cache

struct HotParams {
    const int64_t* rowptr;
    const int32_t* e_idx;     // per nonzero, hot ones first within a document: slot | 0x80000000, or the word id
    const int32_t* e_cnt;
    const uint32_t* e_pos;    // mirror position of each stream entry, or NULL
    const int32_t* hot_words; // [H] word id cached in slot s
    int H;
    int64_t D;
    int K;
    int ld;
    const double* alpha_vec;
    const double* eT_in;
    const double* eB;
    const double* gamma_in;   // previous gamma (loop statistics), or NULL
    double* gamma_out;
    double* ratio_csc;        // or NULL
    int64_t max_iter;
    double tol;
    unsigned long long* iters;
    double* records;
};

__device__ __forceinline__ void st_stream_f64x2(double* p, double2 v, uint64_t pol)
{
    asm volatile("st.global.L1::no_allocate.L2::cache_hint.v2.f64 [%0], {%1, %2}, %3;" ::"l"(p), "d"(v.x), "d"(v.y),
                 "l"(pol)
                 : "memory");
}

__device__ __forceinline__ double2 ld_stream_f64x2(const double* p, uint64_t pol)
{
    double2 v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;"
                 : "=d"(v.x), "=d"(v.y)
                 : "l"(p), "l"(pol));
    return v;
}

// E 16-byte loads of a cached row (shared memory; lane offset folded into the pointer)
template <int G, int E>
__device__ __forceinline__ void load_cached(KSlice<G, E>& s, const double* rowg, bool last_ok)
{
#pragma unroll
    for (int j = 0; j < E - 1; ++j) s.v[j] = *reinterpret_cast<const double2*>(rowg + 2 * G * j);
    s.v[E - 1] = last_ok ? *reinterpret_cast<const double2*>(rowg + 2 * G * (E - 1)) : make_double2(0., 0.);
}

template <int G, int E>
__global__ void __launch_bounds__(kHotThreads, 1) doc_pass_hot_kernel(const HotParams P)
{
    constexpr int R = 32 / G;
    extern __shared__ __align__(16) double smem_hot[];
    double* red = smem_hot;
    double* cache = smem_hot + kHotHeaderDoubles;

    const int lane = threadIdx.x & 31;
    const int g = lane % G, r = lane / G;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int ld = P.ld, K = P.K;
    const bool last_ok = KSlice<G, E>::last_slice_ok(g, ld);
    const uint64_t pol = l2_evict_first_policy();

    // the hot rows: global -> shared, once per CTA (148 x 227 KB = 34 MB of L2 reads per launch)
    {
        const int half = ld >> 1;  // double2 per row (ld % 4 == 0)
        const int total = P.H * half;
        for (int i = threadIdx.x; i < total; i += kHotThreads) {
            const int row = i / half, c2 = i - row * half;
            reinterpret_cast<double2*>(cache)[i] = ldg2(P.eB + (size_t)(uint32_t)P.hot_words[row] * (uint32_t)ld + 2 * c2);
        }
    }
    __syncthreads();

    const double* cache_g = cache + 2 * g;  // this lane's first slice of slot 0 (shared memory)
    const double* table_g = P.eB + 2 * g;
    // Sub-batch `sb` of a batch whose first `nh` entries are hot: all R rows cached -> LDS from the cache; otherwise all R
    // rows come from the global table (a hot entry in a mixed sub-batch — at most one sub-batch per document — is read
    // from the table through the word id its lane looked up when it loaded the entry).  Warp-uniform either way.
    auto load_sub = [&](KSlice<G, E>& row, int raw, int wid, int sb, int nh) {
        const int src = sb * R + r;
        const int v_raw = __shfl_sync(0xffffffffu, raw, src), v_wid = __shfl_sync(0xffffffffu, wid, src);
        if ((sb + 1) * R <= nh)
            load_cached<G, E>(row, cache_g + (uint32_t)(v_raw & 0x7fffffff) * (uint32_t)ld, last_ok);
        else
            row.load_at(row_at(table_g, v_wid, ld), last_ok);
    };

    double tok_acc = 0., cnt_acc = 0.;
    unsigned long long passes = 0, capped = 0;

    for (int64_t d = warp0; d < P.D; d += nwarps) {
        const int64_t b = P.rowptr[d], e = P.rowptr[d + 1];
        KSlice<G, E> t, acc;
        t.load(P.eT_in + d * ld, g, last_ok);
        acc.zero();

        // (row, count, mirror position) triples: 32 per coalesced load, fetched ONE batch ahead
        int nx_w = 0, nx_id = 0, nx_c = 0, nx_pos = 0;
        if (b + lane < e) {
            nx_w = ld_stream_s32(P.e_idx + b + lane, pol);
            nx_c = ld_stream_s32(P.e_cnt + b + lane, pol);
            if (P.ratio_csc) nx_pos = ld_stream_s32(reinterpret_cast<const int32_t*>(P.e_pos) + b + lane, pol);
        }
        nx_id = (nx_w < 0) ? P.hot_words[nx_w & 0x7fffffff] : nx_w;
        int nx_nh = __popc(__ballot_sync(0xffffffffu, nx_w < 0));  // hot entries lead the document (and so the batch)
        KSlice<G, E> rowA, rowB;
        if (b < e) load_sub(rowA, nx_w, nx_id, 0, nx_nh);
        for (int64_t n0 = b; n0 < e; n0 += 32) {
            const int cnt = (int)((e - n0 < 32) ? (e - n0) : 32);
            const int my_w = nx_w, my_id = nx_id, my_c = nx_c, my_pos = nx_pos, my_nh = nx_nh;
            nx_w = 0;
            nx_c = 0;
            if (n0 + 32 + lane < e) {
                nx_w = ld_stream_s32(P.e_idx + n0 + 32 + lane, pol);
                nx_c = ld_stream_s32(P.e_cnt + n0 + 32 + lane, pol);
                if (P.ratio_csc) nx_pos = ld_stream_s32(reinterpret_cast<const int32_t*>(P.e_pos) + n0 + 32 + lane, pol);
            }
            nx_id = (nx_w < 0) ? P.hot_words[nx_w & 0x7fffffff] : nx_w;
            nx_nh = __popc(__ballot_sync(0xffffffffu, nx_w < 0));
            const bool has_next = n0 + 32 < e;
            double my_p = 1.;
            auto body = [&](int s, const KSlice<G, E>& row) {
                const int i = s * R + r;                          // nonzero handled by this group in this sub-batch
                const int c = __shfl_sync(0xffffffffu, my_c, i);  // 0 for i >= cnt: the group idles harmlessly
                // N = theta.beta + 1e-100 (lda_model.py:242), the 1e-100 as the start value of one DFMA chain
                const double p = group_sum<G>(t.dot(row, 1e-100 / G));
                acc.axpy(fast_div_pos((double)c, p), row);  // lda_model.py:248
                const double pv = __shfl_sync(0xffffffffu, p, (lane % R) * G);
                if (lane / R == s) my_p = pv;
            };
            // the row pipeline of traverse.cuh's walk_rows_carried, rows from the cache or from the table
            const int nsub = (cnt + R - 1) / R;
            int s = 0;
            while (true) {
                if (s + 1 < nsub) load_sub(rowB, my_w, my_id, s + 1, my_nh);
                body(s, rowA);
                if (++s >= nsub) {
                    if (has_next) load_sub(rowA, nx_w, nx_id, 0, nx_nh);
                    break;
                }
                if (s + 1 < nsub)
                    load_sub(rowA, my_w, my_id, s + 1, my_nh);
                else if (has_next)
                    load_sub(rowA, nx_w, nx_id, 0, nx_nh);
                body(s, rowB);
                if (++s >= nsub) break;
            }
            if (lane < cnt) {  // one log per nonzero (of theta.beta itself: the token term has no 1e-100, :125-139)
                tok_acc += (double)my_c * log(my_p - 1e-100);
                cnt_acc += (double)my_c;
                if (P.ratio_csc) st_stream_f64(P.ratio_csc + (uint32_t)my_pos, fast_div_pos((double)my_c, my_p), pol);
            }
        }

        // ---- per-document epilogue, in registers ---------------------------------------------------------------------
#pragma unroll
        for (int j = 0; j < E; ++j) {
#pragma unroll
            for (int o = G; o < 32; o <<= 1) {
                acc.v[j].x += __shfl_xor_sync(0xffffffffu, acc.v[j].x, o);
                acc.v[j].y += __shfl_xor_sync(0xffffffffu, acc.v[j].y, o);
            }
        }
        double l1 = 0.;
#pragma unroll
        for (int j = 0; j < E; ++j) {
            if ((j % R) != r) continue;  // group r stores the slices j = r, r + R, ...
            if (j == E - 1 && !last_ok) continue;
            const int k = 2 * (g + G * j);
            const double a0 = (k < K) ? P.alpha_vec[k] : 0., a1 = (k + 1 < K) ? P.alpha_vec[k + 1] : 0.;
            double2 gn;  // gamma_d = alpha + eT_d * acc (lda_model.py:248); padding columns come out as 0
            gn.x = fma(t.v[j].x, acc.v[j].x, a0);
            gn.y = fma(t.v[j].y, acc.v[j].y, a1);
            if (P.iters) {
                const double2 go = ld_stream_f64x2(P.gamma_in + d * ld + k, pol);
                l1 += fabs(gn.x - go.x) + fabs(gn.y - go.y);
            }
            st_stream_f64x2(P.gamma_out + d * ld + k, gn, pol);
        }
        if (P.iters) {
            l1 = warp_sum(l1);
            // the reference's while-loop (lda_model.py:236): pass 1 always runs; pass 2 runs iff the first moved gamma
            // by more than tol and the cap allows it, and then reproduces pass 1 (delta = 0)
            const unsigned long long np = (l1 > P.tol && 1 <= P.max_iter) ? 2ull : 1ull;
            passes += np;
            capped += ((int64_t)np > P.max_iter) ? 1ull : 0ull;
        }
    }

    double v[3] = {tok_acc, 0., cnt_acc};
    block_sum<3>(v, red);
    if (threadIdx.x == 0) {
        double* rec = P.records + (size_t)blockIdx.x * NPART;
        rec[0] = v[0];
        rec[1] = 0.;
        rec[2] = v[2];
#pragma unroll
        for (int j = 3; j < NPART; ++j) rec[j] = 0.;
    }
    if (P.iters && lane == 0 && (passes | capped)) {
        atomicAdd(P.iters, passes);
        if (capped) atomicAdd(P.iters + 1, capped);
    }
}

// E-step stream of a CSR corpus: within every document the nonzeros whose word has a cache slot come first (stable),
// carrying `slot | 0x80000000`, the others follow with their word id; counts and mirror positions move with them.
// One warp per document, two passes over its row (count the hot ones, then a ballot-ranked stable partition).
__global__ void __launch_bounds__(256) estream_build_kernel(const int64_t* __restrict__ rowptr,
                                                            const int32_t* __restrict__ colidx,
                                                            const int32_t* __restrict__ counts,
                                                            const uint32_t* __restrict__ csr2csc, int64_t D,
                                                            const int32_t* __restrict__ slot_of_word,
                                                            int32_t* __restrict__ e_idx, int32_t* __restrict__ e_cnt,
                                                            uint32_t* __restrict__ e_pos)
{
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t d = warp0; d < D; d += nwarps) {
        const int64_t b = rowptr[d], e = rowptr[d + 1];
        int nh = 0;
        for (int64_t p0 = b; p0 < e; p0 += 32) {
            const bool valid = p0 + lane < e;
            const bool hot = valid && slot_of_word[colidx[p0 + lane]] >= 0;
            nh += __popc(__ballot_sync(0xffffffffu, hot));
        }
        int64_t hbase = b, cbase = b + nh;
        for (int64_t p0 = b; p0 < e; p0 += 32) {
            const int64_t p = p0 + lane;
            const bool valid = p < e;
            int w = 0, s = -1;
            if (valid) {
                w = colidx[p];
                s = slot_of_word[w];
            }
            const bool hot = valid && s >= 0;
            const unsigned mh = __ballot_sync(0xffffffffu, hot), mc = __ballot_sync(0xffffffffu, valid && !hot);
            if (valid) {
                const int64_t dst = hot ? hbase + __popc(mh & lt) : cbase + __popc(mc & lt);
                e_idx[dst] = hot ? (int32_t)((uint32_t)s | 0x80000000u) : w;
                e_cnt[dst] = counts[p];
                if (csr2csc) e_pos[dst] = csr2csc[p];
            }
            hbase += __popc(mh);
            cbase += __popc(mc);
        }
    }
}

static int hot_rows_that_fit(int64_t ld)
{
    const int64_t n = ((int64_t)kHotSmemBytes - kHotHeaderDoubles * 8) / (ld * 8);
    return (int)(n < 0 ? 0 : n);
}

template <int G, int E>
static int launch_hot(const HotParams& P, int* grid_out, cudaStream_t s)
{
    const size_t smem = (kHotHeaderDoubles + (size_t)P.H * P.ld) * sizeof(double);
    static bool configured = false;  // per instantiation
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(doc_pass_hot_kernel<G, E>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)kHotSmemBytes);
        if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(doc_pass_hot)");
        configured = true;
    }
    const int grid = grid_for_warps(P.D, kHotThreads / 32, 1);  // persistent: at most one CTA per SM
    doc_pass_hot_kernel<G, E><<<grid, kHotThreads, smem, s>>>(P);
    *grid_out = grid;
    return TTLDA_OK;
}

}  // namespace ttlda

using namespace ttlda;

extern "C" {

int64_t ttlda_estep_hot_rows(int64_t ld)
{
    if (ld < 4 || ld % 4) return TTLDA_EINVAL;
    const GroupShape gs = group_shape_for(ld);
    return gs.E <= 8 ? hot_rows_that_fit(ld) : 0;  // wider rows run the plain kernel (few rows would fit anyway)
}

int ttlda_build_estream(const int64_t* rowptr, const int32_t* colidx, const int32_t* counts, const uint32_t* csr2csc,
                        int64_t D, int64_t V, const int32_t* slot_of_word, int32_t* e_idx, int32_t* e_cnt,
                        uint32_t* e_pos, void* stream)
{
    TTLDA_REQUIRE(rowptr && slot_of_word, "NULL pointer");
    TTLDA_REQUIRE(D >= 0 && V >= 0, "negative size");
    TTLDA_REQUIRE((csr2csc == nullptr) == (e_pos == nullptr), "csr2csc and e_pos go together");
    if (D == 0) return TTLDA_OK;
    TTLDA_REQUIRE(colidx && counts && e_idx && e_cnt, "NULL pointer");
    estream_build_kernel<<<grid_for_warps(D, 8, 16), 256, 0, (cudaStream_t)stream>>>(rowptr, colidx, counts, csr2csc, D,
                                                                                   slot_of_word, e_idx, e_cnt, e_pos);
    TTLDA_LAUNCH_CHECK("estream_build");
    return TTLDA_OK;
}

int ttlda_estep_hot_f64(const int64_t* rowptr, const int32_t* e_idx, const int32_t* e_cnt, const uint32_t* e_pos,
                        const int32_t* hot_words, int64_t H, int64_t D, int64_t K, int64_t V, int64_t ld,
                        const double* alpha_vec, const double* exp_elogtheta_in, const double* exp_elogbeta,
                        const double* gamma_in, double* gamma_out, double* ratio_csc, int64_t max_iter, double tol,
                        int64_t* iters_out, double* partial_out, void* workspace, int64_t workspace_bytes, void* stream)
{
    TTLDA_REQUIRE(rowptr && alpha_vec && exp_elogtheta_in && exp_elogbeta && gamma_out && partial_out && workspace,
                  "NULL pointer");
    TTLDA_REQUIRE(D >= 0 && V >= 0 && K >= 1 && K <= TTLDA_MAX_K, "K must be in [1, 1024]");
    TTLDA_REQUIRE(ld >= K && ld % 4 == 0 && ld < K + 4, "ld must be K rounded up to a multiple of 4");
    TTLDA_REQUIRE(max_iter >= 0, "max_iter must be >= 0");
    TTLDA_REQUIRE(!iters_out || gamma_in, "iters_out needs gamma_in (the previous gamma)");
    TTLDA_REQUIRE((e_pos == nullptr) == (ratio_csc == nullptr), "e_pos and ratio_csc go together");
    TTLDA_REQUIRE(H >= 0 && H <= V && (H == 0 || hot_words), "bad hot-word list");
    const GroupShape gs = group_shape_for(ld);
    TTLDA_REQUIRE(gs.E <= 8 && H <= hot_rows_that_fit(ld), "H exceeds ttlda_estep_hot_rows(ld)");
    if (workspace_bytes < reduce_ws_bytes()) {
        set_error("estep_hot: workspace %lld < %lld", (long long)workspace_bytes, (long long)reduce_ws_bytes());
        return TTLDA_EWORKSPACE;
    }
    cudaStream_t s = (cudaStream_t)stream;
    if (iters_out) {
        cudaError_t e = cudaMemsetAsync(iters_out, 0, 2 * sizeof(int64_t), s);
        if (e != cudaSuccess) return cuda_fail(e, "memset iters");
    }
    HotParams P{rowptr, e_idx, e_cnt, e_pos, hot_words, (int)H, D, (int)K, (int)ld, alpha_vec, exp_elogtheta_in,
                exp_elogbeta, gamma_in, gamma_out, ratio_csc, max_iter, tol, (unsigned long long*)iters_out,
                (double*)workspace};
    int grid = 0, rc = TTLDA_EINVAL;
    bool launched = false;
#define TTLDA_CASE(GG, EE)                               \
    if (!launched && gs.G == GG && gs.E == EE) {         \
        rc = launch_hot<GG, EE>(P, &grid, s);            \
        launched = true;                                 \
    }
    TTLDA_FOR_EACH_SMALL_SHAPE(TTLDA_CASE)
#undef TTLDA_CASE
    if (!launched) {
        set_error("no hot-cache kernel instantiated for ld=%lld (G=%d, E=%d)", (long long)ld, gs.G, gs.E);
        return TTLDA_EINVAL;
    }
    if (rc != TTLDA_OK) return rc;
    TTLDA_LAUNCH_CHECK("estep_hot");
    return launch_reduce_records((const double*)workspace, grid, partial_out, s);
}

}  // extern "C"
