// sstats.cu — M-step sufficient statistics, word-major pass over the CSC mirror.
//
// Reference: lambda_update[:, ids] += np.outer(exp_expec_log_theta_d, d_word_cts.T/phi_d_norm)
// (tuotuo/lda_model.py:262).  The reference scatters K-vectors into a (K,V) matrix document by document; here the
// same sum is evaluated from the word's side so that the K accumulators of a word live in registers and nothing
// is scattered:   S[w,k] = sum_{d in postings(w)} eT[d,k] * c_dw / (eT[d,:].eB[w,:] + 1e-100).
// phi (D x N x K) is never materialised; not even the per-nonzero normaliser is stored — it is recomputed from the
// gathered eT row (K FMAs), which is cheaper than 12 more bytes per nonzero of HBM traffic.
//
// Work decomposition: the nnz CSC positions are cut into fixed chunks (nnz-balanced, immune to Zipfian posting
// lists); one warp per chunk walks the words intersecting it.  A word that lies wholly inside a chunk is written
// with a plain store; a word cut by a chunk boundary contributes one partial K-vector per chunk, written to a
// carry buffer and summed in chunk order by a second kernel => bit-reproducible, no atomics.
// Bound: HBM gather of 8K bytes per nonzero (eT rows, D x K table, far larger than L2 at the target sizes).
#include "ttlda_common.cuh"

namespace ttlda {

constexpr int kChunk = 512;  // CSC positions per warp work item

struct SStatsParams {
    const int64_t* colptr;
    const int32_t* docidx;
    const int32_t* ccounts;
    int64_t V;
    int K;
    int64_t ld;
    int64_t nnz;
    int64_t nchunks;
    const double* eT;
    const double* eB;
    double* S;
    double* carry;       // [nchunks][2][ld]  head / tail partials of words cut by a chunk boundary
    int32_t* carry_word; // [nchunks][2]      word id of each carry slot, -1 = unused
};

template <int KP>
__device__ __forceinline__ void load_row2(double2 (&v)[KP], const double* __restrict__ row, int lane, int ld)
{
#pragma unroll
    for (int j = 0; j < KP; ++j) {
        const int k = 2 * lane + 64 * j;
        v[j] = (k < ld) ? ldg2(row + k) : make_double2(0., 0.);
    }
}

template <int KP>
__device__ __forceinline__ void store_row2(const double2 (&v)[KP], double* __restrict__ row, int lane, int ld)
{
#pragma unroll
    for (int j = 0; j < KP; ++j) {
        const int k = 2 * lane + 64 * j;
        if (k < ld) *reinterpret_cast<double2*>(row + k) = v[j];
    }
}

template <int KP, bool PREFETCH>
__global__ void __launch_bounds__(256) sstats_csc_kernel(const SStatsParams P)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int ld = (int)P.ld;

    for (int64_t c = warp0; c < P.nchunks; c += nwarps) {
        const int64_t cb = c * kChunk, ce = (cb + kChunk < P.nnz) ? cb + kChunk : P.nnz;
        // first word with colptr[w+1] > cb  (upper_bound over colptr[1..V])
        int64_t lo = 0, hi = P.V;  // search w in [0, V)
        while (lo < hi) {
            const int64_t mid = (lo + hi) >> 1;
            if (P.colptr[mid + 1] > cb) hi = mid; else lo = mid + 1;
        }
        int64_t w = lo;
        if (lane == 0) {
            P.carry_word[c * 2 + 0] = -1;
            P.carry_word[c * 2 + 1] = -1;
        }
        int64_t pos = cb;
        while (pos < ce) {
            int64_t wb = P.colptr[w], we = P.colptr[w + 1];
            while (we <= pos) {  // skip words with empty posting lists
                ++w;
                wb = we;
                we = P.colptr[w + 1];
            }
            const int64_t se = (we < ce) ? we : ce;  // segment [pos, se) of word w
            double2 b[KP], acc[KP];
            load_row2<KP>(b, P.eB + w * P.ld, lane, ld);
#pragma unroll
            for (int j = 0; j < KP; ++j) acc[j] = make_double2(0., 0.);

            for (int64_t n0 = pos; n0 < se; n0 += 32) {
                const int cnt = (int)((se - n0 < 32) ? (se - n0) : 32);
                int my_d = 0, my_c = 0;
                if (lane < cnt) {
                    my_d = ld_stream_s32(P.docidx + n0 + lane);
                    my_c = ld_stream_s32(P.ccounts + n0 + lane);
                }
                double2 cur[KP], nxt[KP];
                {
                    const int d0 = __shfl_sync(0xffffffffu, my_d, 0);
                    load_row2<KP>(cur, P.eT + (int64_t)d0 * P.ld, lane, ld);
                }
                for (int i = 0; i < cnt; ++i) {
                    if (PREFETCH && i + 1 < cnt) {
                        const int dn = __shfl_sync(0xffffffffu, my_d, i + 1);
                        load_row2<KP>(nxt, P.eT + (int64_t)dn * P.ld, lane, ld);
                    }
                    const int cc = __shfl_sync(0xffffffffu, my_c, i);
                    double p = 0.;
#pragma unroll
                    for (int j = 0; j < KP; ++j) {
                        p = fma(cur[j].x, b[j].x, p);
                        p = fma(cur[j].y, b[j].y, p);
                    }
                    p = warp_sum(p);
                    const double r = (double)cc / (p + 1e-100);  // lda_model.py:242,262
#pragma unroll
                    for (int j = 0; j < KP; ++j) {
                        acc[j].x = fma(r, cur[j].x, acc[j].x);
                        acc[j].y = fma(r, cur[j].y, acc[j].y);
                    }
                    if (PREFETCH) {
#pragma unroll
                        for (int j = 0; j < KP; ++j) cur[j] = nxt[j];
                    } else if (i + 1 < cnt) {
                        const int dn = __shfl_sync(0xffffffffu, my_d, i + 1);
                        load_row2<KP>(cur, P.eT + (int64_t)dn * P.ld, lane, ld);
                    }
                }
            }
            const bool whole = (pos == wb) && (se == we);
            if (whole) {
                store_row2<KP>(acc, P.S + w * P.ld, lane, ld);
            } else {
                // head slot: the word was started by an earlier chunk; tail slot: it continues in a later one.
                // A word spanning the entire chunk (neither starts nor ends here) uses the head slot.
                const int slot = (pos != wb) ? 0 : 1;
                store_row2<KP>(acc, P.carry + (c * 2 + slot) * P.ld, lane, ld);
                if (lane == 0) P.carry_word[c * 2 + slot] = (int32_t)w;
            }
            pos = se;
            if (se == we) ++w;
        }
    }
}

// Sum the carries of each cut word in chunk order.  One warp per chunk whose TAIL slot is in use (= the chunk
// where a cut word starts); it walks the following chunks' HEAD slots while they name the same word.
__global__ void __launch_bounds__(256) sstats_carry_kernel(const SStatsParams P)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t c = warp0; c < P.nchunks; c += nwarps) {
        // a cut word starts in chunk c either in its tail slot, or — if chunk c begins exactly at the word's first
        // posting and the word outlives the chunk — also in the tail slot (pos == wb there), so tail is the start.
        const int32_t w = P.carry_word[c * 2 + 1];
        if (w < 0) continue;
        for (int k = lane; k < (int)P.ld; k += 32) {
            double s = P.carry[(c * 2 + 1) * P.ld + k];
            for (int64_t c2 = c + 1; c2 < P.nchunks && P.carry_word[c2 * 2 + 0] == w; ++c2)
                s += P.carry[(c2 * 2 + 0) * P.ld + k];
            P.S[(int64_t)w * P.ld + k] = s;
        }
    }
}

static int64_t nchunks_for(int64_t nnz) { return (nnz + kChunk - 1) / kChunk; }

static size_t align256(size_t x) { return (x + 255) / 256 * 256; }

}  // namespace ttlda

using namespace ttlda;

extern "C" {

int64_t ttlda_sstats_workspace_bytes(int64_t nnz, int64_t V, int64_t ld)
{
    (void)V;
    const int64_t nc = nchunks_for(nnz);
    return (int64_t)(align256((size_t)nc * 2 * ld * sizeof(double)) + align256((size_t)nc * 2 * sizeof(int32_t)) + 256);
}

int ttlda_sstats_csc_f64(const int64_t* colptr, const int32_t* docidx, const int32_t* ccounts, int64_t V, int64_t K,
                         int64_t ld, int64_t nnz, const double* exp_elogtheta, const double* exp_elogbeta,
                         double* sstats, void* workspace, int64_t workspace_bytes, void* stream)
{
    TTLDA_REQUIRE(colptr && exp_elogtheta && exp_elogbeta && sstats && workspace, "NULL pointer");
    TTLDA_REQUIRE(V >= 0 && nnz >= 0 && K >= 1 && K <= TTLDA_MAX_K, "K must be in [1, 1024]");
    TTLDA_REQUIRE(ld >= K && ld % 4 == 0, "ld must be a multiple of 4 and >= K");
    if (workspace_bytes < ttlda_sstats_workspace_bytes(nnz, V, ld)) {
        set_error("sstats_csc: workspace %lld < %lld", (long long)workspace_bytes,
                  (long long)ttlda_sstats_workspace_bytes(nnz, V, ld));
        return TTLDA_EWORKSPACE;
    }
    cudaStream_t s = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(sstats, 0, (size_t)V * ld * sizeof(double), s);  // words with no postings
    if (e != cudaSuccess) return cuda_fail(e, "memset sstats");
    if (nnz == 0 || V == 0) return TTLDA_OK;
    const int64_t nc = nchunks_for(nnz);
    SStatsParams P{colptr, docidx, ccounts, V, (int)K, ld, nnz, nc, exp_elogtheta, exp_elogbeta, sstats,
                   (double*)workspace,
                   (int32_t*)((char*)workspace + align256((size_t)nc * 2 * ld * sizeof(double)))};
    const int grid = grid_for_warps(nc, 8, 8);
    const int need = (int)((ld + 63) / 64);
    bool launched = false;
#define TTLDA_CASE(KP, PF)                                   \
    if (!launched && need <= KP) {                           \
        sstats_csc_kernel<KP, PF><<<grid, 256, 0, s>>>(P);   \
        launched = true;                                     \
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
    TTLDA_LAUNCH_CHECK("sstats_csc");
    sstats_carry_kernel<<<grid, 256, 0, s>>>(P);
    TTLDA_LAUNCH_CHECK("sstats_carry");
    return TTLDA_OK;
}

}  // extern "C"
