// sstats.cu — M-step sufficient statistics, word-major pass over the CSC mirror.
//
// Reference: lambda_update[:, ids] += np.outer(exp_expec_log_theta_d, d_word_cts.T/phi_d_norm)
// (tuotuo/lda_model.py:262).  The reference scatters K-vectors into a (K,V) matrix document by document; here the
// same sum is evaluated from the word's side so that the K accumulators of a word live in registers and nothing
// is scattered:   S[w,k] = sum_{d in postings(w)} eT[d,k] * c_dw / (eT[d,:].eB[w,:] + 1e-100).
// phi (D x N x K) is never materialised.  Two forms: with the ratio c/N of every nonzero left by the E-step in mirror
// order (USE_RATIO: a pure gather-accumulate, the default), or recomputing the normaliser from the gathered eT row
// and the word's eB K-vector (callers without the csr->csc map: fixed-point mode).
//
// Mapping (traverse.cuh): the word's exp(E[log beta_w]) K-vector is replicated over the warp's R = 32/G lane
// groups; each group gathers the exp(E[log theta_d]) K-vector of a different posting, R postings per instruction;
// group partials are summed through shared memory when the word's segment ends.
//
// Work decomposition: the nnz CSC positions are cut into fixed chunks (nnz-balanced, immune to Zipfian posting
// lists), handed out to warps in mirror order by a ticket counter; a warp walks the words intersecting its chunk.
// A word that lies wholly inside a chunk is written with a plain store; a word cut by a chunk boundary contributes
// one partial K-vector per chunk, written to a carry buffer and summed in chunk order by a second kernel =>
// bit-reproducible, no data atomics (the ticket is the only atomic: it decides WHO computes a chunk, never WHAT).
//
// The CSC mirror may be DOCUMENT-BLOCKED (ttlda_csr_to_csc with nblocks > 1): "virtual word" b*V + w holds the
// postings of word w inside document block b, so that consecutive chunks gather eT rows from one block-sized
// window of the D x K table that stays resident in the 126 MB L2; the per-block partial statistics are summed in
// block order by sstats_fold_kernel.  Bound: gather of 8K bytes per nonzero (HBM unblocked, L2 blocked).
#include "tma_gather.cuh"

namespace ttlda {

constexpr int kChunk = 1024;      // CSC positions per warp work item
constexpr int kWordThreads = 128;  // 4 warps per CTA

struct SStatsParams {
    const int64_t* colptr;  // [VV+1], VV = nblocks * V virtual words
    const int32_t* docidx;
    const int32_t* ccounts;
    int64_t V;
    int64_t VV;
    int K;
    int64_t ld;        // row stride of eT / eB (doubles)
    int kw;            // width of the topic window [k0, k0 + kw) this launch accumulates (eT, S, carry already point at
                       // column k0); kw == ld: the whole K-vector
    int64_t s_ld;      // row stride of S   (ld when S is the final table, kw for per-block partials of a window)
    int64_t c_ld;      // row stride of carry
    int64_t pos_base;  // mirror positions [pos_base, pos_end) are processed, cut into nchunks chunks (pos_base != 0:
    int64_t pos_end;   // one document block of a corpus-wide mirror, ttlda_sstats_csc_block_f64)
    int64_t nchunks;
    const double* eT;
    const double* eB;
    const double* ratio; // c/N per posting in mirror order (left by the E-step), or NULL => recompute
    double* S;           // [VV][ld] per-(block, word) statistics (== the final [V][ld] table when nblocks == 1)
    double* carry;       // [nchunks][2][ld]  head / tail partials of words cut by a chunk boundary
    int32_t* carry_word; // [nchunks][2]      virtual word id of each carry slot, -1 = unused
    unsigned long long* ticket;  // next chunk to hand out (zeroed before the launch)
    int accumulate;              // add to S instead of overwriting it (blocks processed one launch at a time)
};

// USE_RATIO: the E-step left c/N per posting (ttlda_estep_f64 ratio_csc): the pass is then a pure gather-accumulate
// S[w,:] += ratio * eT[d,:] — no beta K-vector in registers, no dot product, shuffle reduction or division.
template <int G, int E, bool PREFETCH, bool USE_RATIO>
__global__ void __launch_bounds__(kWordThreads, (E <= 8 ? (USE_RATIO ? 4 : 3) : 1)) sstats_csc_kernel(const SStatsParams P)
{
    constexpr int R = 32 / G;
    constexpr int LDP = 2 * G * E;
    constexpr int MMAX = (LDP + 31) / 32;
    extern __shared__ double smem[];
    double* wbuf = smem + (threadIdx.x >> 5) * (R * LDP);

    const int lane = threadIdx.x & 31;
    const int g = lane % G, r = lane / G;
    const int ld = (int)P.ld, kw = P.kw;
    const bool last_ok = KSlice<G, E>::last_slice_ok(g, kw);
    const uint64_t pol = l2_evict_first_policy();

    // Chunks are handed out in mirror order through a global ticket counter, so that the warps resident at any moment
    // work on ~one narrow range of positions => on ONE document block's exp(E[log theta]) window, which then stays in
    // L2 until the whole block is done (a static chunk -> warp map lets waves of CTAs and fast warps drift across
    // blocks and re-read every window several times: 73 GB of DRAM traffic per pass at cfg3, profiles/r01h).
    for (;;) {
        unsigned long long ticket = 0;
        if (lane == 0) ticket = atomicAdd(P.ticket, 1ull);
        const int64_t c = (int64_t)__shfl_sync(0xffffffffu, ticket, 0);
        if (c >= P.nchunks) break;
        // chunks sit on the corpus-wide grid of kChunk positions (so a block processed on its own cuts its words exactly
        // where the one-call pass over the whole mirror does => identical bits)
        const int64_t cg = (P.pos_base / kChunk + c) * kChunk;
        const int64_t cb = (cg > P.pos_base) ? cg : P.pos_base, ce = (cg + kChunk < P.pos_end) ? cg + kChunk : P.pos_end;
        // first virtual word with colptr[w+1] > cb: warp-cooperative 32-ary search (4 dependent loads for 10^6 words
        // instead of 20 — the chunk's start-up latency is not hidden by anything else in this warp)
        int64_t lo = 0, hi = P.VV;  // the answer lies in [lo, hi); colptr[VV] = nnz > cb
        while (hi - lo > 1) {
            const int64_t step = (hi - lo + 31) >> 5;
            int64_t wi = lo + (int64_t)(lane + 1) * step - 1;
            if (wi > hi - 1) wi = hi - 1;
            const unsigned hit = __ballot_sync(0xffffffffu, P.colptr[wi + 1] > cb);  // monotone: 0..0 1..1
            const int j = __ffs(hit) - 1;                                            // lane 31 probes hi-1 => hit != 0
            const int64_t wj = lo + (int64_t)(j + 1) * step - 1;
            const int64_t new_hi = ((wj > hi - 1) ? hi - 1 : wj) + 1;
            if (j > 0) lo = lo + (int64_t)j * step;  // = w_{j-1} + 1
            hi = new_hi;
        }
        int64_t w = lo;
        if (lane == 0) {
            P.carry_word[c * 2 + 0] = -1;
            P.carry_word[c * 2 + 1] = -1;
        }
        int64_t pos = cb;
        while (pos < ce) {
            int64_t wb = P.colptr[w], we = P.colptr[w + 1];
            while (we <= pos) {  // skip (virtual) words with empty posting lists
                ++w;
                wb = we;
                we = P.colptr[w + 1];
            }
            const int64_t se = (we < ce) ? we : ce;  // segment [pos, se) of virtual word w
            KSlice<G, E> b, acc;
            if (!USE_RATIO) b.load(P.eB + (w % P.V) * P.ld, g, last_ok);
            acc.zero();

            // (document, count | ratio) pairs: 32 per coalesced load, fetched ONE batch ahead
            int nx_d = 0, nx_c = 0;
            double nx_r = 0.;
            if (pos + lane < se) {
                nx_d = ld_stream_s32(P.docidx + pos + lane, pol);
                if (USE_RATIO) nx_r = ld_stream_f64(P.ratio + pos + lane, pol);
                else nx_c = ld_stream_s32(P.ccounts + pos + lane, pol);
            }
            KSlice<G, E> rowA;  // sub-batch 0 of the coming batch (carried across batches when PREFETCH)
            if (PREFETCH) rowA.load_at(row_at(P.eT + 2 * g, __shfl_sync(0xffffffffu, nx_d, r), ld), last_ok);
            for (int64_t n0 = pos; n0 < se; n0 += 32) {
                const int cnt = (int)((se - n0 < 32) ? (se - n0) : 32);
                const int my_d = nx_d, my_c = nx_c;
                const double my_r = nx_r;
                nx_d = 0;
                nx_c = 0;
                nx_r = 0.;
                if (n0 + 32 + lane < se) {
                    nx_d = ld_stream_s32(P.docidx + n0 + 32 + lane, pol);
                    if (USE_RATIO) nx_r = ld_stream_f64(P.ratio + n0 + 32 + lane, pol);
                    else nx_c = ld_stream_s32(P.ccounts + n0 + 32 + lane, pol);
                }
                auto body = [&](int s, const KSlice<G, E>& row) {
                    if (USE_RATIO) {
                        acc.axpy(__shfl_sync(0xffffffffu, my_r, s * R + r), row);  // 0 past the end of the segment
                    } else {
                        const int cc = __shfl_sync(0xffffffffu, my_c, s * R + r);  // 0 past the end of the segment
                        const double p = group_sum<G>(b.dot(row));
                        acc.axpy(fast_div_pos((double)cc, p + 1e-100), row);  // lda_model.py:242,262
                    }
                };
                if (PREFETCH)
                    walk_rows_carried<G, E>(P.eT, ld, my_d, cnt, nx_d, n0 + 32 < se, rowA, g, r, last_ok, body);
                else
                    walk_rows<G, E, false>(P.eT, ld, my_d, cnt, g, r, last_ok, body);
            }
            // group partials -> one K-vector, written with coalesced stores
            acc.to_smem(wbuf + r * LDP, g);
            __syncwarp();
            const bool whole = (pos == wb) && (se == we);
            // head slot: the word was started by an earlier chunk; tail slot: it continues in a later one.
            // A word spanning the entire chunk (neither starts nor ends here) uses the head slot.
            const int slot = (pos != wb) ? 0 : 1;
            double* dst = whole ? P.S + w * P.s_ld : P.carry + (c * 2 + slot) * P.c_ld;
#pragma unroll
            for (int m = 0; m < MMAX; ++m) {
                const int k = lane + 32 * m;
                if (k < kw) {
                    double a = 0.;
#pragma unroll
                    for (int q = 0; q < R; ++q) a += wbuf[q * LDP + k];
                    if (whole && P.accumulate) a += dst[k];  // serial-block mode: S[w] += this block's partial
                    st_stream_f64(dst + k, a, pol);
                }
            }
            __syncwarp();
            if (!whole && lane == 0) P.carry_word[c * 2 + slot] = (int32_t)w;
            pos = se;
            if (se == we) ++w;
        }
    }
}

// The ratio variant fed by the TMA engine (tma_gather.cuh): one gather4 per sub-batch lands 4 exp(E[log theta]) rows in
// shared memory, kTmaStages sub-batches ahead, continuously across a segment's 32-item batches.  G = 8 only.
template <int E>
__global__ void __launch_bounds__(kWordThreads, 4) sstats_tma_kernel(const __grid_constant__ CUtensorMap map,
                                                                     const SStatsParams P)
{
    constexpr int G = 8, R = 4;
    constexpr int LDP = 2 * G * E;
    constexpr int MMAX = (LDP + 31) / 32;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char* wbase = smem_raw + (size_t)(threadIdx.x >> 5) * Gather4Pipe<E>::bytes_per_warp(P.ld);

    const int lane = threadIdx.x & 31;
    const int g = lane % G, r = lane / G;
    const int ld = (int)P.ld;
    const bool last_ok = KSlice<G, E>::last_slice_ok(g, ld);
    const uint64_t pol = l2_evict_first_policy();

    Gather4Pipe<E> pipe;
    pipe.init(wbase, ld, lane);
    double* wbuf = reinterpret_cast<double*>(wbase);  // the segment epilogue's transpose buffer aliases the drained stages

    // Chunks are handed out in mirror order through a global ticket counter, so that the warps resident at any moment
    // work on ~one narrow range of positions => on ONE document block's exp(E[log theta]) window, which then stays in
    // L2 until the whole block is done (a static chunk -> warp map lets waves of CTAs and fast warps drift across
    // blocks and re-read every window several times: 73 GB of DRAM traffic per pass at cfg3, profiles/r01h).
    for (;;) {
        unsigned long long ticket = 0;
        if (lane == 0) ticket = atomicAdd(P.ticket, 1ull);
        const int64_t c = (int64_t)__shfl_sync(0xffffffffu, ticket, 0);
        if (c >= P.nchunks) break;
        // chunks sit on the corpus-wide grid of kChunk positions (so a block processed on its own cuts its words exactly
        // where the one-call pass over the whole mirror does => identical bits)
        const int64_t cg = (P.pos_base / kChunk + c) * kChunk;
        const int64_t cb = (cg > P.pos_base) ? cg : P.pos_base, ce = (cg + kChunk < P.pos_end) ? cg + kChunk : P.pos_end;
        int64_t lo = 0, hi = P.VV;  // warp-cooperative 32-ary search for the first word with colptr[w+1] > cb
        while (hi - lo > 1) {
            const int64_t step = (hi - lo + 31) >> 5;
            int64_t wi = lo + (int64_t)(lane + 1) * step - 1;
            if (wi > hi - 1) wi = hi - 1;
            const unsigned hit = __ballot_sync(0xffffffffu, P.colptr[wi + 1] > cb);
            const int j = __ffs(hit) - 1;
            const int64_t wj = lo + (int64_t)(j + 1) * step - 1;
            const int64_t new_hi = ((wj > hi - 1) ? hi - 1 : wj) + 1;
            if (j > 0) lo = lo + (int64_t)j * step;
            hi = new_hi;
        }
        int64_t w = lo;
        if (lane == 0) {
            P.carry_word[c * 2 + 0] = -1;
            P.carry_word[c * 2 + 1] = -1;
        }
        int64_t pos = cb;
        while (pos < ce) {
            int64_t wb = P.colptr[w], we = P.colptr[w + 1];
            while (we <= pos) {
                ++w;
                wb = we;
                we = P.colptr[w + 1];
            }
            const int64_t se = (we < ce) ? we : ce;
            KSlice<G, E> acc;
            acc.zero();
            double r_cur = 0., r_nxt = 0.;
            const int64_t n = se - pos;
            tma_walk<E>(
                pipe, &map, P.docidx + pos, n, lane, g, r, last_ok, pol,
                [&](int64_t base) { r_nxt = (base + lane < n) ? ld_stream_f64(P.ratio + pos + base + lane, pol) : 0.; },
                [&]() { r_cur = r_nxt; },
                [&](int sb, const KSlice<G, E>& row) { acc.axpy(__shfl_sync(0xffffffffu, r_cur, sb * R + r), row); },
                [](int) {});
            acc.to_smem(wbuf + r * LDP, g);
            __syncwarp();
            const bool whole = (pos == wb) && (se == we);
            const int slot = (pos != wb) ? 0 : 1;
            double* dst = whole ? P.S + w * P.s_ld : P.carry + (c * 2 + slot) * P.c_ld;
#pragma unroll
            for (int m = 0; m < MMAX; ++m) {
                const int k = lane + 32 * m;
                if (k < ld) {
                    double a = 0.;
#pragma unroll
                    for (int q = 0; q < R; ++q) a += wbuf[q * LDP + k];
                    if (whole && P.accumulate) a += dst[k];  // serial-block mode: S[w] += this block's partial
                    st_stream_f64(dst + k, a, pol);
                }
            }
            __syncwarp();
            if (!whole && lane == 0) P.carry_word[c * 2 + slot] = (int32_t)w;
            pos = se;
            if (se == we) ++w;
        }
    }
}

template <int E>
static int launch_sstats_tma(const SStatsParams& P, int grid, int64_t table_rows, cudaStream_t s)
{
    CUtensorMap map;
    int rc = make_row_gather_map(P.eT, table_rows, P.ld, &map);
    if (rc != TTLDA_OK) return rc;
    const size_t smem = (kWordThreads / 32) * Gather4Pipe<E>::bytes_per_warp(P.ld);
    cudaError_t e = cudaFuncSetAttribute(sstats_tma_kernel<E>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(sstats_tma)");
    sstats_tma_kernel<E><<<grid, kWordThreads, smem, s>>>(map, P);
    return TTLDA_OK;
}

// Sum the carries of each cut word in chunk order.  One warp per chunk whose TAIL slot is in use (= the chunk in
// which a cut word starts); it walks the following chunks' HEAD slots while they name the same word.
__global__ void __launch_bounds__(256) sstats_carry_kernel(const SStatsParams P)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t c = warp0; c < P.nchunks; c += nwarps) {
        const int32_t w = P.carry_word[c * 2 + 1];
        if (w < 0) continue;
        int64_t last = c;
        while (last + 1 < P.nchunks && P.carry_word[(last + 1) * 2 + 0] == w) ++last;
        for (int k = lane; k < P.kw; k += 32) {
            double s = P.carry[(c * 2 + 1) * P.c_ld + k];
            for (int64_t c2 = c + 1; c2 <= last; ++c2) s += P.carry[(c2 * 2 + 0) * P.c_ld + k];
            P.S[(int64_t)w * P.s_ld + k] = P.accumulate ? P.S[(int64_t)w * P.s_ld + k] + s : s;
        }
    }
}

// S[w][k] = sum_b Sb[b*V + w][k] in block order (document-blocked mirror only)
__global__ void __launch_bounds__(256) sstats_fold_kernel(const double* __restrict__ Sb, int64_t V, int64_t ld,
                                                          int nblocks, double* __restrict__ S)
{
    const int64_t total = V * ld;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        double s = 0.;
        for (int b = 0; b < nblocks; ++b) s += Sb[(int64_t)b * total + i];
        S[i] = s;
    }
}

// The same for one topic window: S[w][k] (k < kw; S already points at column k0, row stride ld) = sum_b Sb[b*V + w][k],
// Sb rows kw wide
__global__ void __launch_bounds__(256) sstats_fold_window_kernel(const double* __restrict__ Sb, int64_t V, int kw,
                                                                 int64_t ld, int nblocks, double* __restrict__ S)
{
    const int64_t total = V * kw;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        double s = 0.;
        for (int b = 0; b < nblocks; ++b) s += Sb[(int64_t)b * total + i];
        S[(i / kw) * ld + (i % kw)] = s;
    }
}

static int64_t nchunks_for(int64_t nnz) { return (nnz + kChunk - 1) / kChunk; }
// chunks of the corpus-wide grid that intersect [pos_base, pos_base + nnz)
static int64_t nchunks_for(int64_t pos_base, int64_t nnz) { return nchunks_for(pos_base % kChunk + nnz); }

static size_t align256(size_t x) { return (x + 255) / 256 * 256; }

template <int G, int E>
static void launch_shape(const SStatsParams& P, int grid, cudaStream_t s)
{
    constexpr int R = 32 / G;
    constexpr size_t smem = (size_t)(kWordThreads / 32) * R * 2 * G * E * sizeof(double);
    constexpr bool PF = (E <= 8);
    if (P.ratio) sstats_csc_kernel<G, E, PF, true><<<grid, kWordThreads, smem, s>>>(P);
    else sstats_csc_kernel<G, E, PF, false><<<grid, kWordThreads, smem, s>>>(P);
}

// One-pass alternative (the north-star's literal formulation): document-major scatter with red.global.add.f64.
// One warp per document, lane l owns topics l, l+32, ...; the normaliser is recomputed per nonzero.  Kept for
// comparison: REDG throughput bounds it well below the two-pass form, and the sums are not bit-reproducible.
__global__ void __launch_bounds__(256) sstats_atomic_kernel(const int64_t* __restrict__ rowptr,
                                                            const int32_t* __restrict__ colidx,
                                                            const int32_t* __restrict__ counts, int64_t D, int K,
                                                            int64_t ld, const double* __restrict__ eT,
                                                            const double* __restrict__ eB, double* __restrict__ S)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t d = warp0; d < D; d += nwarps) {
        const double* t = eT + d * ld;
        for (int64_t n = rowptr[d]; n < rowptr[d + 1]; ++n) {
            const int64_t w = colidx[n];
            const double* b = eB + w * ld;
            double p = 0.;
            for (int k = lane; k < K; k += 32) p = fma(t[k], b[k], p);
            p = warp_sum(p);
            const double rr = (double)counts[n] / (p + 1e-100);  // lda_model.py:242,262
            for (int k = lane; k < K; k += 32) atomicAdd(S + w * ld + k, t[k] * rr);
        }
    }
}

}  // namespace ttlda

using namespace ttlda;

extern "C" {

int64_t ttlda_sstats_workspace_bytes(int64_t nnz, int64_t V, int64_t ld, int64_t nblocks)
{
    const int64_t nc = nchunks_for(nnz) + 1;  // + 1: a block that does not start on the chunk grid
    // [256: ticket][carry][carry_word][per-block partials, nblocks > 1 only — always the last bytes]
    size_t b = 256 + align256((size_t)nc * 2 * ld * sizeof(double)) + align256((size_t)nc * 2 * sizeof(int32_t));
    if (nblocks > 1) b += (size_t)nblocks * V * ld * sizeof(double);
    return (int64_t)b;
}

static int sstats_csc_impl(const int64_t* colptr, const int32_t* docidx, const int32_t* ccounts, int64_t D, int64_t V,
                           int64_t K, int64_t ld, int64_t pos_base, int64_t nnz, int64_t nblocks,
                           const double* exp_elogtheta, const double* exp_elogbeta, const double* ratio_csc,
                           double* sstats, int accumulate, void* workspace, int64_t workspace_bytes, cudaStream_t s)
{
    TTLDA_REQUIRE(colptr && exp_elogtheta && (exp_elogbeta || ratio_csc) && sstats && workspace, "NULL pointer");
    TTLDA_REQUIRE(D >= 0 && V >= 0 && nnz >= 0 && pos_base >= 0 && K >= 1 && K <= TTLDA_MAX_K, "K must be in [1, 1024]");
    TTLDA_REQUIRE(ld >= K && ld % 4 == 0 && ld < K + 4, "ld must be K rounded up to a multiple of 4");
    TTLDA_REQUIRE(nblocks >= 1 && nblocks * V < ((int64_t)1 << 31), "bad nblocks");
    if (workspace_bytes < ttlda_sstats_workspace_bytes(nnz, V, ld, nblocks)) {
        set_error("sstats_csc: workspace %lld < %lld", (long long)workspace_bytes,
                  (long long)ttlda_sstats_workspace_bytes(nnz, V, ld, nblocks));
        return TTLDA_EWORKSPACE;
    }
    const int64_t nc = nchunks_for(pos_base, nnz);
    char* ws = (char*)workspace;
    unsigned long long* ticket = (unsigned long long*)ws;
    ws += 256;
    const int64_t nc_cap = nchunks_for(nnz) + 1;  // the layout ttlda_sstats_workspace_bytes sized (nc <= nc_cap)
    double* carry = (double*)ws;
    int32_t* carry_word = (int32_t*)(ws + align256((size_t)nc_cap * 2 * ld * sizeof(double)));
    // the per-block partials are the TAIL of the workspace (callers of the block entry points may use that region as
    // their [nblocks][V][ld] buffer: it is not touched by single-block calls)
    double* Sb = (nblocks > 1)
                     ? (double*)((char*)carry_word + align256((size_t)nc_cap * 2 * sizeof(int32_t)))
                     : sstats;
    const int64_t VV = nblocks * V;
    cudaError_t e = cudaSuccess;
    if (!accumulate) e = cudaMemsetAsync(Sb, 0, (size_t)VV * ld * sizeof(double), s);  // words with no postings
    if (e != cudaSuccess) return cuda_fail(e, "memset sstats");
    if (nnz == 0 || V == 0) {
        if (nblocks > 1) {
            e = cudaMemsetAsync(sstats, 0, (size_t)V * ld * sizeof(double), s);
            if (e != cudaSuccess) return cuda_fail(e, "memset sstats");
        }
        return TTLDA_OK;
    }
    SStatsParams P{colptr, docidx, ccounts, V, VV, (int)K, ld, (int)ld, ld, ld, pos_base, pos_base + nnz, nc,
                   exp_elogtheta, exp_elogbeta, ratio_csc, Sb, carry, carry_word, ticket, accumulate};
    e = cudaMemsetAsync(ticket, 0, sizeof(unsigned long long), s);
    if (e != cudaSuccess) return cuda_fail(e, "memset ticket");
    const int grid = grid_for_warps(nc, kWordThreads / 32, 16);
    GroupShape gs = group_shape_for(ld);
    if (const char* ov = getenv("TTLDA_GROUP")) {
        const int G = atoi(ov);
        if (G == 2 || G == 4 || G == 8 || G == 16 || G == 32) gs = {G, (int)((ld + 2 * G - 1) / (2 * G))};
    }
    bool launched = false;
    if (ratio_csc && tma_gather_applicable(gs.G, ld) && tma_gather_enabled() && D > 0) {  // rows fed by TMA gather4
        int rc = TTLDA_OK;
        switch (gs.E) {
            case 5: rc = launch_sstats_tma<5>(P, grid, D, s); launched = true; break;
            case 6: rc = launch_sstats_tma<6>(P, grid, D, s); launched = true; break;
            case 7: rc = launch_sstats_tma<7>(P, grid, D, s); launched = true; break;
            case 8: rc = launch_sstats_tma<8>(P, grid, D, s); launched = true; break;
            default: break;
        }
        if (rc != TTLDA_OK) return rc;
    }
#define TTLDA_CASE(GG, EE)                        \
    if (!launched && gs.G == GG && gs.E == EE) {  \
        launch_shape<GG, EE>(P, grid, s);         \
        launched = true;                          \
    }
    TTLDA_FOR_EACH_SHAPE(TTLDA_CASE)
#undef TTLDA_CASE
    if (!launched) {
        set_error("no kernel instantiated for ld=%lld (G=%d, E=%d)", (long long)ld, gs.G, gs.E);
        return TTLDA_EINVAL;
    }
    TTLDA_LAUNCH_CHECK("sstats_csc");
    sstats_carry_kernel<<<(unsigned)((nc + 7) / 8), 256, 0, s>>>(P);  // one chunk per warp: the pass is three dependent round trips per chunk
    TTLDA_LAUNCH_CHECK("sstats_carry");
    if (nblocks > 1) {
        const int64_t total = V * ld;
        int fg = (int)((total + 255) / 256 < (int64_t)kSMs * 8 ? (total + 255) / 256 : (int64_t)kSMs * 8);
        sstats_fold_kernel<<<fg, 256, 0, s>>>(Sb, V, ld, (int)nblocks, sstats);
        TTLDA_LAUNCH_CHECK("sstats_fold");
    }
    return TTLDA_OK;
}

int64_t ttlda_sstats_windowed_workspace_bytes(int64_t nnz, int64_t V, int64_t kw, int64_t nblocks)
{
    if (nnz < 0 || V < 0 || kw < 4 || kw % 4 || nblocks < 1) return TTLDA_EINVAL;
    const int64_t nc = nchunks_for(nnz) + 1;
    // [256: ticket][carry, kw wide][carry_word][per-block partials of ONE window, nblocks > 1 only]
    size_t b = 256 + align256((size_t)nc * 2 * kw * sizeof(double)) + align256((size_t)nc * 2 * sizeof(int32_t));
    if (nblocks > 1) b += (size_t)nblocks * V * kw * sizeof(double);
    return (int64_t)b;
}

// Topic-windowed form of the ratio pass: the K columns are cut into windows of kw, and the WHOLE mirror is walked once
// per window, gathering only columns [k0, k0 + kw) of every exp(E[log theta]) row.  What a document block contributes to
// the L2 working set is then block_docs * kw * 8 bytes instead of block_docs * K * 8: at K = 1000 a 50 MB window holds
// 51k documents instead of 6.5k, so (block, word) segments stay long and the per-block partial statistics
// (nblocks * V * kw doubles, re-used by every window) stay small — the two things that made document blocking alone
// unaffordable at cfg5-class shapes.  The price: docidx and ratio are re-read once per window
// (12 bytes against kw * 8 gathered).  Chunks, carries and the block fold are per column, so the result is
// bit-reproducible run to run; against the unwindowed pass it differs only in the order in which a word's postings are
// dealt to the warp's lane groups (R = 32/G depends on the row width): ~1 ulp.
int ttlda_sstats_csc_windowed_f64(const int64_t* colptr, const int32_t* docidx, int64_t D, int64_t V, int64_t K,
                                  int64_t ld, int64_t nnz, int64_t nblocks, int64_t kw, const double* exp_elogtheta,
                                  const double* ratio_csc, double* sstats, void* workspace, int64_t workspace_bytes,
                                  void* stream)
{
    TTLDA_REQUIRE(colptr && exp_elogtheta && ratio_csc && sstats && workspace, "NULL pointer");
    TTLDA_REQUIRE(D >= 0 && V >= 0 && nnz >= 0 && K >= 1 && K <= TTLDA_MAX_K, "K must be in [1, 1024]");
    TTLDA_REQUIRE(ld >= K && ld % 4 == 0 && ld < K + 4, "ld must be K rounded up to a multiple of 4");
    TTLDA_REQUIRE(kw >= 16 && kw % 16 == 0 && kw <= 256, "window width must be a multiple of 16 in [16, 256]");
    TTLDA_REQUIRE(nblocks >= 1 && nblocks * V < ((int64_t)1 << 31), "bad nblocks");
    const int64_t need = ttlda_sstats_windowed_workspace_bytes(nnz, V, kw, nblocks);
    if (workspace_bytes < need) {
        set_error("sstats_csc_windowed: workspace %lld < %lld", (long long)workspace_bytes, (long long)need);
        return TTLDA_EWORKSPACE;
    }
    cudaStream_t s = (cudaStream_t)stream;
    cudaError_t e;
    if (nnz == 0 || V == 0) {
        e = cudaMemsetAsync(sstats, 0, (size_t)V * ld * sizeof(double), s);
        return e == cudaSuccess ? TTLDA_OK : cuda_fail(e, "memset sstats");
    }
    const int64_t nc = nchunks_for(nnz), nc_cap = nc + 1;
    char* ws = (char*)workspace;
    unsigned long long* ticket = (unsigned long long*)ws;
    ws += 256;
    double* carry = (double*)ws;
    int32_t* carry_word = (int32_t*)(ws + align256((size_t)nc_cap * 2 * kw * sizeof(double)));
    double* Sb = (double*)((char*)carry_word + align256((size_t)nc_cap * 2 * sizeof(int32_t)));
    const int64_t VV = nblocks * V;
    const int grid = grid_for_warps(nc, kWordThreads / 32, 16);
    if (nblocks == 1) {  // words without postings: zero the final table once
        e = cudaMemsetAsync(sstats, 0, (size_t)V * ld * sizeof(double), s);
        if (e != cudaSuccess) return cuda_fail(e, "memset sstats");
    }
    for (int64_t k0 = 0; k0 < ld; k0 += kw) {
        const int64_t w = (ld - k0 < kw) ? ld - k0 : kw;  // a multiple of 4
        SStatsParams P{colptr, docidx, nullptr, V, VV, (int)K, ld, (int)w,
                       nblocks > 1 ? w : ld, w, 0, nnz, nc, exp_elogtheta + k0, nullptr, ratio_csc,
                       nblocks > 1 ? Sb : sstats + k0, carry, carry_word, ticket, 0};
        e = cudaMemsetAsync(ticket, 0, sizeof(unsigned long long), s);
        if (e != cudaSuccess) return cuda_fail(e, "memset ticket");
        if (nblocks > 1) {
            e = cudaMemsetAsync(Sb, 0, (size_t)VV * w * sizeof(double), s);
            if (e != cudaSuccess) return cuda_fail(e, "memset window partials");
        }
        const GroupShape gs = group_shape_for(w);
        bool launched = false;
#define TTLDA_CASE(GG, EE)                        \
    if (!launched && gs.G == GG && gs.E == EE) {  \
        launch_shape<GG, EE>(P, grid, s);         \
        launched = true;                          \
    }
        TTLDA_FOR_EACH_SHAPE(TTLDA_CASE)
#undef TTLDA_CASE
        if (!launched) {
            set_error("no kernel instantiated for window width %lld (G=%d, E=%d)", (long long)w, gs.G, gs.E);
            return TTLDA_EINVAL;
        }
        TTLDA_LAUNCH_CHECK("sstats_csc (window)");
        sstats_carry_kernel<<<(unsigned)((nc + 7) / 8), 256, 0, s>>>(P);
        TTLDA_LAUNCH_CHECK("sstats_carry (window)");
        if (nblocks > 1) {
            const int64_t total = V * w;
            int fg = (int)((total + 255) / 256 < (int64_t)kSMs * 8 ? (total + 255) / 256 : (int64_t)kSMs * 8);
            sstats_fold_window_kernel<<<fg, 256, 0, s>>>(Sb, V, (int)w, ld, (int)nblocks, sstats + k0);
            TTLDA_LAUNCH_CHECK("sstats_fold (window)");
        }
    }
    return TTLDA_OK;
}

int ttlda_sstats_csc_f64(const int64_t* colptr, const int32_t* docidx, const int32_t* ccounts, int64_t D, int64_t V,
                         int64_t K, int64_t ld, int64_t nnz, int64_t nblocks, const double* exp_elogtheta,
                         const double* exp_elogbeta, const double* ratio_csc, double* sstats, void* workspace,
                         int64_t workspace_bytes, void* stream)
{
    return sstats_csc_impl(colptr, docidx, ccounts, D, V, K, ld, 0, nnz, nblocks, exp_elogtheta, exp_elogbeta, ratio_csc,
                           sstats, 0, workspace, workspace_bytes, (cudaStream_t)stream);
}

int ttlda_sstats_csc_block_f64(const int64_t* colptr_block, const int32_t* docidx, const int32_t* ccounts, int64_t D,
                               int64_t V, int64_t K, int64_t ld, int64_t pos_base, int64_t nnz_block,
                               const double* exp_elogtheta, const double* exp_elogbeta, const double* ratio_csc,
                               double* sstats_block, int accumulate, void* workspace, int64_t workspace_bytes,
                               void* stream)
{
    return sstats_csc_impl(colptr_block, docidx, ccounts, D, V, K, ld, pos_base, nnz_block, 1, exp_elogtheta,
                           exp_elogbeta, ratio_csc, sstats_block, accumulate != 0, workspace, workspace_bytes,
                           (cudaStream_t)stream);
}

int ttlda_sstats_fold_f64(const double* sstats_blocks, int64_t nblocks, int64_t V, int64_t ld, double* sstats,
                          void* stream)
{
    TTLDA_REQUIRE(sstats_blocks && sstats, "NULL pointer");
    TTLDA_REQUIRE(nblocks >= 1 && nblocks < ((int64_t)1 << 31) && V >= 0 && ld >= 1, "bad shape");
    const int64_t total = V * ld;
    if (total == 0) return TTLDA_OK;
    int fg = (int)((total + 255) / 256 < (int64_t)kSMs * 8 ? (total + 255) / 256 : (int64_t)kSMs * 8);
    sstats_fold_kernel<<<fg, 256, 0, (cudaStream_t)stream>>>(sstats_blocks, V, ld, (int)nblocks, sstats);
    TTLDA_LAUNCH_CHECK("sstats_fold");
    return TTLDA_OK;
}

int ttlda_sstats_atomic_f64(const int64_t* rowptr, const int32_t* colidx, const int32_t* counts, int64_t D, int64_t K,
                            int64_t V, int64_t ld, const double* exp_elogtheta, const double* exp_elogbeta,
                            double* sstats, void* stream)
{
    TTLDA_REQUIRE(rowptr && exp_elogtheta && exp_elogbeta && sstats, "NULL pointer");
    TTLDA_REQUIRE(D >= 0 && V >= 0 && K >= 1 && K <= TTLDA_MAX_K, "K must be in [1, 1024]");
    TTLDA_REQUIRE(ld >= K && ld % 4 == 0 && ld < K + 4, "ld must be K rounded up to a multiple of 4");
    cudaStream_t s = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(sstats, 0, (size_t)V * ld * sizeof(double), s);
    if (e != cudaSuccess) return cuda_fail(e, "memset sstats");
    if (D == 0) return TTLDA_OK;
    sstats_atomic_kernel<<<grid_for_warps(D, 8, 8), 256, 0, s>>>(rowptr, colidx, counts, D, (int)K, ld, exp_elogtheta,
                                                                exp_elogbeta, sstats);
    TTLDA_LAUNCH_CHECK("sstats_atomic");
    return TTLDA_OK;
}

}  // extern "C"
