// corpus.cu — bag-of-words construction on the device: token-id stream -> CSR doc-term matrix, and its CSC mirror.
//
// Reference: get_vocab_from_docs / get_np_wct (tuotuo/utils.py:177-206) build a dense V x M float64 count matrix
// with Python dict/list loops; Document exposes its transpose (tuotuo/document.py:33-36) and the E-step reads rows
// through np.nonzero(X[d,:]) (tuotuo/lda_model.py:226-227), i.e. ascending word ids with their counts.  The CSR
// built here holds exactly those (ids, counts) per document; integer work, bit-exact.
//
// Sorting/scan plumbing uses CUB (ships with the CUDA toolkit); key packing, run splitting and pointer
// construction are hand-written.  One-time cost per corpus, outside the EM loop.
#include <cub/cub.cuh>

#include "ttlda_common.cuh"

namespace ttlda {

static size_t align256(size_t x) { return (x + 255) / 256 * 256; }

static int bits_for(int64_t n)
{
    int b = 1;
    while (((int64_t)1 << b) < n) ++b;
    return b;
}

// key = (doc << vbits) | word for every token; doc found by binary search in doc_offsets.
__global__ void pack_keys_kernel(const int32_t* __restrict__ tok, const int64_t* __restrict__ offs, int64_t T,
                                 int64_t D, int vbits, uint64_t* __restrict__ keys)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < T; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t lo = 0, hi = D;  // last d with offs[d] <= i
        while (hi - lo > 1) {
            const int64_t mid = (lo + hi) >> 1;
            if (offs[mid] <= i) lo = mid; else hi = mid;
        }
        keys[i] = ((uint64_t)lo << vbits) | (uint64_t)(uint32_t)tok[i];
    }
}

// unique (doc,word) keys -> colidx, and rowptr from the doc boundaries (documents may be empty).
__global__ void split_runs_kernel(const uint64_t* __restrict__ ukeys, const int64_t* __restrict__ nruns_p, int64_t D,
                                  int vbits, int32_t* __restrict__ colidx, int64_t* __restrict__ rowptr,
                                  int64_t* __restrict__ nnz_out)
{
    const int64_t n = *nruns_p;
    const uint64_t vmask = ((uint64_t)1 << vbits) - 1;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i <= n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t dcur = (i < n) ? (int64_t)(ukeys[i] >> vbits) : D;
        const int64_t dprev = (i > 0) ? (int64_t)(ukeys[i - 1] >> vbits) : -1;
        if (i < n) colidx[i] = (int32_t)(ukeys[i] & vmask);
        for (int64_t d = dprev + 1; d <= dcur; ++d) rowptr[d] = i;  // rowptr[D] = n when i == n
        if (i == n) *nnz_out = n;
    }
}

// key of CSR position p = (doc(p) / block_docs) * V + word(p).  One warp per document walks its row (coalesced);
// also materialises doc(p) and the identity permutation for the pair sort.
__global__ void __launch_bounds__(256) csc_keys_kernel(const int64_t* __restrict__ rowptr,
                                                       const int32_t* __restrict__ colidx, int64_t D, int64_t V,
                                                       int64_t nnz, int64_t block_docs, int64_t doc_base,
                                                       int64_t pos_base, uint32_t* __restrict__ keys,
                                                       uint32_t* __restrict__ docs, uint32_t* __restrict__ pos)
{
    // doc_base / pos_base != 0: rowptr is a slice of a larger CSR (ttlda_csr_to_csc_block) — document d of the slice is
    // document doc_base + d of the corpus, its entries sit at absolute positions rowptr[d].. = pos_base + local
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t d = warp0; d < D; d += nwarps) {
        const int64_t b = rowptr[d] - pos_base, e = rowptr[d + 1] - pos_base;
        const uint32_t base = (uint32_t)((d / block_docs) * V);
        for (int64_t p = b + lane; p < e; p += 32) {
            keys[p] = base + (uint32_t)colidx[p];
            docs[p] = (uint32_t)(doc_base + d);
            pos[p] = (uint32_t)p;
        }
    }
}

// after the stable sort by (block, word): gather doc id and count of each CSR position; build colptr over the
// VV = nblocks * V virtual words.  (Keeping four independent gathers in flight per thread was tried: 36.2 vs 35.7 ms for
// the streamed step — no gain, the L2-resident random accesses are not what bounds it.)
__global__ void csc_gather_kernel(const uint32_t* __restrict__ skeys, const uint32_t* __restrict__ spos,
                                  const uint32_t* __restrict__ docs, const int32_t* __restrict__ counts, int64_t VV,
                                  int64_t nnz, int64_t pos_base, int64_t* __restrict__ colptr,
                                  int32_t* __restrict__ docidx, int32_t* __restrict__ ccounts,
                                  uint32_t* __restrict__ csr2csc)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i <= nnz; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t wcur = (i < nnz) ? (int64_t)skeys[i] : VV;
        const int64_t wprev = (i > 0) ? (int64_t)skeys[i - 1] : -1;
        for (int64_t w = wprev + 1; w <= wcur; ++w) colptr[w] = pos_base + i;
        if (i < nnz) {
            const int64_t p = spos[i];
            docidx[i] = (int32_t)docs[p];
            if (ccounts) ccounts[i] = counts[p];  // NULL: the caller's statistics pass runs on the E-step's ratios
            if (csr2csc) csr2csc[p] = (uint32_t)(pos_base + i);  // where CSR position p lives in the mirror
        }
    }
}

// Compact host formats (uint16 word ids when V <= 65536, uint8 / uint16 counts) are widened to the int32 arrays the
// kernels read right after they land on the device: 3 bytes per nonzero cross PCIe instead of 8.
// bound > 0: items >= bound are reported in *flags (bit 1) and stored as 0, so that nothing downstream can index out
// of the [V][ld] tables with them (the check rides on a pass that touches every item anyway).
template <typename T>
__global__ void __launch_bounds__(256) widen_kernel(const T* __restrict__ src, int32_t* __restrict__ dst, int64_t n,
                                                    uint32_t bound, int32_t* __restrict__ flags)
{
    bool bad = false;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        uint32_t v = (uint32_t)src[i];
        if (bound && v >= bound) {
            bad = true;
            v = 0;
        }
        dst[i] = (int32_t)v;
    }
    if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicOr(flags, TTLDA_BAD_COLIDX);
}

// One pass over a CSR corpus that is about to drive device addressing (gathers of eB + w*ld, eT + d*ld, scatters
// through rowptr): rowptr must start at pos_base, be monotone and end at pos_base + nnz; word ids must lie in [0, V);
// counts must be >= 0.  Violations are reported in *flags; with `sanitize` they are also made harmless (rowptr clamped
// into [pos_base, pos_base + nnz] and made monotone by the clamp of each entry against the endpoints only — enough to
// keep every [b, e) range inside the arrays —, word ids set to 0), so that a caller that checks the flags AFTER queueing
// its kernels (the streaming step) cannot have corrupted memory in between.
__global__ void __launch_bounds__(256) validate_csr_kernel(int64_t* __restrict__ rowptr, int64_t D, int64_t pos_base,
                                                           int32_t* __restrict__ colidx, int32_t* __restrict__ counts,
                                                           int64_t nnz, int64_t V, int sanitize,
                                                           int32_t* __restrict__ flags)
{
    int bad = 0;
    const int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, step = (int64_t)gridDim.x * blockDim.x;
    if (rowptr) {
        const int64_t lo = pos_base, hi = pos_base + nnz;
        for (int64_t d = i0; d <= D; d += step) {
            const int64_t r = rowptr[d];
            bool b = r < lo || r > hi || (d == 0 && r != lo) || (d == D && r != hi);
            if (!b && d > 0) {
                const int64_t q = rowptr[d - 1];
                b = q > r;
            }
            if (b) {
                bad |= TTLDA_BAD_ROWPTR;
                if (sanitize && (r < lo || r > hi)) rowptr[d] = r < lo ? lo : hi;
            }
        }
    }
    if (colidx) {
        for (int64_t p = i0; p < nnz; p += step) {
            if ((uint64_t)(int64_t)colidx[p] >= (uint64_t)V) {
                bad |= TTLDA_BAD_COLIDX;
                if (sanitize) colidx[p] = 0;
            }
            if (counts && counts[p] < 0) {
                bad |= TTLDA_BAD_COUNT;
                if (sanitize) counts[p] = 0;
            }
        }
    }
    bad = __reduce_or_sync(0xffffffffu, bad);
    if (bad && (threadIdx.x & 31) == 0) atomicOr(flags, bad);
}

struct BowWs {
    uint64_t *keys, *skeys, *ukeys;
    int64_t* nruns;
    void* cub;
    size_t cub_bytes, total;
};

static cudaError_t bow_layout(int64_t T, int64_t D, int64_t V, char* base, BowWs& w)
{
    size_t sort_b = 0, rle_b = 0;
    const int end_bit = bits_for(V) + bits_for(D);
    cudaError_t e = cub::DeviceRadixSort::SortKeys(nullptr, sort_b, (uint64_t*)nullptr, (uint64_t*)nullptr, T, 0, end_bit);
    if (e != cudaSuccess) return e;
    e = cub::DeviceRunLengthEncode::Encode(nullptr, rle_b, (uint64_t*)nullptr, (uint64_t*)nullptr, (int32_t*)nullptr,
                                           (int64_t*)nullptr, (int)T);
    if (e != cudaSuccess) return e;
    size_t off = 0;
    w.keys = (uint64_t*)(base + off);  off += align256((size_t)T * 8);
    w.skeys = (uint64_t*)(base + off); off += align256((size_t)T * 8);
    w.ukeys = (uint64_t*)(base + off); off += align256((size_t)T * 8);
    w.nruns = (int64_t*)(base + off);  off += 256;
    w.cub = base + off;
    w.cub_bytes = align256(sort_b > rle_b ? sort_b : rle_b);
    off += w.cub_bytes;
    w.total = off;
    return cudaSuccess;
}

struct CscWs {
    uint32_t *keys, *docs, *skeys, *pos, *spos;
    void* cub;
    size_t cub_bytes, total;
};

static cudaError_t csc_layout(int64_t nnz, int64_t VV, char* base, CscWs& w)
{
    size_t sort_b = 0;
    cudaError_t e = cub::DeviceRadixSort::SortPairs(nullptr, sort_b, (uint32_t*)nullptr, (uint32_t*)nullptr,
                                                    (uint32_t*)nullptr, (uint32_t*)nullptr, nnz, 0, bits_for(VV));
    if (e != cudaSuccess) return e;
    size_t off = 0;
    w.keys = (uint32_t*)(base + off);  off += align256((size_t)nnz * 4);
    w.docs = (uint32_t*)(base + off);  off += align256((size_t)nnz * 4);
    w.skeys = (uint32_t*)(base + off); off += align256((size_t)nnz * 4);
    w.pos = (uint32_t*)(base + off);   off += align256((size_t)nnz * 4);
    w.spos = (uint32_t*)(base + off);  off += align256((size_t)nnz * 4);
    w.cub = base + off;
    w.cub_bytes = align256(sort_b);
    off += w.cub_bytes;
    w.total = off;
    return cudaSuccess;
}

static int grid_1d(int64_t n)
{
    int64_t g = (n + 255) / 256;
    if (g > (int64_t)kSMs * 16) g = (int64_t)kSMs * 16;
    return (int)(g < 1 ? 1 : g);
}

// ---------------------------------------------------------------------------------------------------------------------
// Mirror of ONE document block by a hand-written, comparison-free counting sort: the "slice walk".
// (Default for blocks whose vocabulary fits the shared-memory counter table, V <= ~110 000; TTLDA_CSC_WALK=0 or a larger
// vocabulary selects the generic CUB pair sort above.)
//
// A CSR block is already sorted by (document, word); its mirror is the same entries sorted by (word, document) — a
// transpose.  The mirror position of entry (d, w) is
//     colptr[w] + #{documents d' < d of the block that contain w},
// so what has to be computed per entry is a RANK among the earlier documents.  The block's documents are cut into S
// contiguous slices (S = 148 or 296: one or two resident CTAs per SM); then
//   1 walk   one CTA per slice keeps a 16-bit counter per word in shared memory (V * 2 bytes: 100 KB at V = 50k, 205 KB at
//            V = 102 660) and visits its documents IN ORDER, one barrier per document: every entry takes its word's counter
//            value as its rank inside the slice (rank[p], uint16) and increments it.  A document holds a word at most once,
//            so inside a document no two threads meet on a counter and the ranks do not depend on thread order:
//            deterministic.  The entries reach the walk through a double-buffered shared-memory stage filled by cp.async a
//            group of documents ahead (a per-thread register queue of loads was tried first: with more loads in flight than
//            the six scoreboards a thread has, waiting for the oldest waits for the youngest, and every step paid a full
//            memory latency — 152 us per block instead of ~25).  At the end the counter table IS the slice's word
//            histogram: hist[slice][w].
//   2 scan   per word, the exclusive prefix of hist over the slices -> off[slice][w], the totals, and (a second small kernel,
//            one scan over the per-CTA sums) colptr.
//   3 place  q = colptr[w] + off[slice(d)][w] + rank[p];  docidx[q] = d, (ccounts[q] = counts[p],) csr2csc[p] = q — one CTA
//            per slice stages the slice's bases colptr + off in shared memory with coalesced loads and then runs flat over
//            the slice's entries (document of an entry by binary search in the staged row offsets); csr2csc is one coalesced
//            stream in CSR order.
// Against the generic path (key packing, histogram, two 8-bit onesweep passes over (key, value) pairs, a dependent random
// gather of doc ids through the sorted permutation) there is no pair sort and no 20-bytes-per-entry scratch (2 bytes per
// entry + 6 bytes per (slice, word)).
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kWalkThreads = 256;      // entries of one document handled per step (longer documents loop)
constexpr int kWalkDocs = 256;         // documents per chunk of staged row offsets
constexpr int kWalkSmemSM = 227 * 1024;
constexpr int kWalkMinStage = 2048;    // entries per stage buffer (2 buffers): at least
constexpr int kWalkMaxStage = 8192;    //                                        at most
constexpr int kPlaceThreads = 1024;

struct WalkPlan {
    bool ok;
    int64_t vp, dps, slices, parts, partw;
    int stage_cap;  // entries per stage buffer
    size_t smem, rank_b, hist_b, off_b, total_b, total;
};

static WalkPlan walk_plan(int64_t D, int64_t V, int64_t nnz)
{
    WalkPlan w{};
    w.vp = (V + 7) / 8 * 8;  // counters are cleared / flushed as 16-byte vectors of eight
    // One walk CTA per slice holds the whole vocabulary's counters when they fit next to the stage buffers (V <= ~105 000);
    // larger vocabularies are cut into `parts` word ranges, one CTA per (slice, range): every CTA of a slice sees all the
    // slice's entries and counts those of its range.  (Ranges were also tried as a way to put four small CTAs on an SM
    // at V = 50k: the walk got 2x SLOWER — it is bound by the instructions of a step, not by their latency, and every
    // range repeats the steps.)
    const size_t budget1 = (size_t)kWalkSmemSM - 4 * 1024;  // one CTA per SM: dynamic shared memory available
    const size_t stage_min = (size_t)kWalkMinStage * 8;
    w.parts = ((size_t)w.vp * 2 + (budget1 - stage_min) - 1) / (budget1 - stage_min);
    if (w.parts < 1) w.parts = 1;
    w.partw = ((w.vp + w.parts - 1) / w.parts + 7) / 8 * 8;
    const size_t table = (size_t)w.partw * 2;
    // two CTAs per SM when two tables + two minimal stage pairs (+ static shared memory and the per-CTA reserve) fit
    const int64_t per = (w.parts == 1 && 2 * (table + stage_min + 4 * 1024) <= (size_t)kWalkSmemSM) ? 2 : 1;
    const size_t budget = (size_t)kWalkSmemSM / per - 4 * 1024;
    size_t stage = (budget - table) / 8;  // entries per buffer (two buffers of 4-byte entries)
    if (stage > (size_t)kWalkMaxStage) stage = kWalkMaxStage;
    w.stage_cap = (int)(stage / 4 * 4);
    w.smem = table + (size_t)w.stage_cap * 8;
    const int64_t target = (int64_t)kSMs * per;  // slices: one wave of walk CTAs (parts waves for a cut vocabulary)
    w.dps = (D + target - 1) / target;
    if (w.dps < 16) w.dps = 16;
    w.slices = D > 0 ? (D + w.dps - 1) / w.dps : 1;
    w.ok = V >= 1 && w.parts <= 64 && w.dps <= 65535 && nnz < ((int64_t)1 << 31) && target * w.vp < ((int64_t)1 << 32);
    // buffer sizes from a bound on the slice count that is monotone in D (the workspace is sized for the largest block, a
    // call may bring a smaller one: slices <= target and slices <= ceil(D / 16) always hold)
    int64_t cap = (D + 15) / 16;
    if (cap > target) cap = target;
    if (cap < 1) cap = 1;
    w.rank_b = align256((size_t)(nnz > 0 ? nnz : 1) * 4);
    w.hist_b = align256((size_t)cap * w.vp * 2);
    w.off_b = align256((size_t)cap * w.vp * 4);
    w.total_b = align256((size_t)w.vp * 4 + 16) + align256((size_t)(w.vp / 512 + 2) * 8);  // totals, redo flag, CTA sums
    w.total = w.rank_b + w.hist_b + w.off_b + w.total_b;
    return w;
}

static bool walk_path_applies(int64_t D, int64_t V, int64_t nnz, int64_t nblocks)
{
    if (nblocks != 1) return false;
    const char* e = getenv("TTLDA_CSC_WALK");
    if (e && e[0] == '0') return false;
    return walk_plan(D, V, nnz).ok;
}

__device__ __forceinline__ void walk_cp_async4(uint32_t dst, const void* src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void walk_cp_async16(uint32_t dst, const void* src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void walk_cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void walk_cp_wait()
{
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// 1: the walk.  rank[p] = documents of the slice before p's that contain p's word;  hist[slice][*] = the slice's word counts.
// Also fills docidx[.] of the slice's entry range with a valid document id, so that whatever a malformed corpus makes of
// the placement (3), no slot of the mirror is left holding garbage that a gather would follow.
//
// ATOMIC = false (the fast form): the counter update is a plain 16-bit LDS + STS.  That is only right when no document
// repeats a word, which holds for every row the reference can produce (np.nonzero) and is what Document.from_csr checks; the
// kernel checks it again on the fly (each entry against its predecessor: strictly ascending) and raises *redo if a row
// fails.  ATOMIC = true is the same walk with atomicAdd on the 32-bit word that holds two counters: launched right behind
// the fast form, it returns at once unless *redo is set, and then recomputes the block's ranks and histograms correctly for
// ANY row order / repetition (distinct, in-range ranks).
template <bool ATOMIC>
__global__ void __launch_bounds__(kWalkThreads) csc_walk_kernel(const int64_t* __restrict__ rowptr,
                                                                const int32_t* __restrict__ colidx, int64_t D, int64_t V,
                                                                int64_t vp, int64_t partw, int64_t dps, int cap,
                                                                int64_t doc_base, int64_t pos_base,
                                                                uint32_t* __restrict__ rank, uint32_t* __restrict__ hist32,
                                                                int32_t* __restrict__ docidx, int32_t* __restrict__ redo)
{
    extern __shared__ __align__(16) uint32_t walk_tab[];  // partw / 2 words (two 16-bit counters each), 2 stages of `cap`
    __shared__ int s_off[kWalkDocs + 1];                  // entry offsets of up to kWalkDocs documents
    if (ATOMIC && *redo == 0) return;
    const int t = threadIdx.x;
    const int wlo = (int)(blockIdx.y * partw);                                   // this CTA's word range [wlo, wlo + wn)
    const int wn = (int)((wlo + partw <= vp) ? partw : (vp > wlo ? vp - wlo : 0));
    const int half = wn >> 1;
    const bool first_part = blockIdx.y == 0;
    int32_t* stage = reinterpret_cast<int32_t*>(walk_tab + (partw >> 1));
    const uint32_t stage_u32 = (uint32_t)__cvta_generic_to_shared(stage);
    const int64_t d_lo = (int64_t)blockIdx.x * dps;
    const int64_t d_hi = (d_lo + dps < D) ? d_lo + dps : D;
    for (int i = t * 4; i < half; i += kWalkThreads * 4) *reinterpret_cast<uint4*>(walk_tab + i) = make_uint4(0u, 0u, 0u, 0u);
    uint16_t* tab16 = reinterpret_cast<uint16_t*>(walk_tab);
    bool unordered = false;

    // jl: the document's index inside the slice; what is left for the placement is (jl << 16) | rank
    auto place = [&](int w, int64_t p, unsigned jl) {
        if (!ATOMIC && first_part) docidx[p] = (int32_t)doc_base;
        const unsigned x = (unsigned)(w - wlo);
        if (x < (unsigned)wn) {  // a word of this CTA's range (ids outside [0, vp) are nobody's: the placement skips them)
            unsigned r;
            if (ATOMIC) {
                const unsigned sh = (x & 1u) << 4;
                r = (atomicAdd(walk_tab + (x >> 1), 1u << sh) >> sh) & 0xffffu;
            } else {
                r = tab16[x];
                tab16[x] = (uint16_t)(r + 1);
            }
            rank[p] = (jl << 16) | r;
        }
    };

    for (int64_t c0 = d_lo; c0 < d_hi; c0 += kWalkDocs) {
        const int n = (int)((d_hi - c0 < kWalkDocs) ? d_hi - c0 : kWalkDocs);
        __syncthreads();  // the table is cleared / the previous chunk's offsets and stages are no longer read
        const int64_t base = rowptr[c0] - pos_base;
        for (int i = t; i <= n; i += kWalkThreads) s_off[i] = (int)(rowptr[c0 + i] - pos_base - base);
        __syncthreads();
        const int32_t* ci = colidx + base;
        // the documents [j0, j1) whose entries fit one stage buffer; j1 == j0: document j0 alone is longer than a buffer
        // (a stage buffer also holds up to 3 entries in front of the group's first one: the copies are 16 bytes wide and
        // start at the 16-byte boundary at or below it — 4-byte cp.async costs the LSU as much per instruction and made
        // the fill, not the walk, the bound)
        auto sub_end = [&](int j0) {
            const int lim = s_off[j0] + cap - 4;
            int lo = j0, hi = n + 1;  // largest j in [j0, n] with s_off[j] <= lim
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (s_off[mid] <= lim) lo = mid; else hi = mid;
            }
            return lo;
        };
        // returns the number of leading pad entries: entry s_off[j0] + i of the corpus sits at stage[buf][lead + i]
        auto fill = [&](int buf, int j0, int j1) {
            const int b = s_off[j0], cnt = s_off[j1] - b;
            const int lead = (int)((reinterpret_cast<uintptr_t>(ci + b) & 15) >> 2);
            const int32_t* src = ci + b - lead;  // 16-byte aligned; the pad entries belong to this block or to the array
                                                 // in front of it (b - lead >= 0 whenever lead > 0: the base is aligned)
            const uint32_t dst = stage_u32 + (uint32_t)buf * (uint32_t)cap * 4u;
            const int nvec = (lead + cnt) >> 2;  // whole 16-byte vectors
            for (int i = t; i < nvec; i += kWalkThreads) walk_cp_async16(dst + 16u * (uint32_t)i, src + 4 * i);
            const int tail = (lead + cnt) & 3;  // the last 1..3 entries: 4-byte copies, nothing is read past the group
            if (t < tail) walk_cp_async4(dst + 4u * (uint32_t)(4 * nvec + t), src + 4 * nvec + t);
            return lead;
        };
        int j0 = 0, buf = 0, lead0 = 0, lead1 = 0;
        int j1 = sub_end(0);
        bool long0 = (j1 == 0);
        if (long0) j1 = 1; else lead0 = fill(0, 0, j1);
        walk_cp_commit();
        while (j0 < n) {
            int j2 = j1;
            bool long1 = false;
            if (j1 < n) {
                j2 = sub_end(j1);
                long1 = (j2 == j1);
                if (long1) j2 = j1 + 1; else lead1 = fill(buf ^ 1, j1, j2);
            }
            walk_cp_commit();
            walk_cp_wait<1>();  // this thread's copies of the current group have landed ...
            __syncthreads();    // ... and so have everybody else's
            if (long0) {  // one document longer than a stage buffer: straight from global memory
                const int b = s_off[j0], e = s_off[j0 + 1];
                const unsigned jl = (unsigned)(c0 - d_lo) + (unsigned)j0;
                for (int i = b + t; i < e; i += kWalkThreads) {
                    const int w = ci[i];
                    if (i > b) unordered |= (w <= ci[i - 1]);
                    place(w, base + i, jl);
                }
                __syncthreads();
            } else {
                // the entry a thread handles in document j + 1 is fetched from the stage while document j is being counted:
                // between two barriers only the counter update (LDS -> STS) is left on the critical path
                const int32_t* st = stage + (size_t)buf * cap + lead0 - s_off[j0];
                int b = s_off[j0], e = s_off[j0 + 1];
                int wn = -1, wpn = -1;
                if (b + t < e) {
                    wn = st[b + t];
                    if (t > 0) wpn = st[b + t - 1];
                }
                for (int j = j0; j < j1; ++j) {
                    const int cb = b, ce = e, w = wn, wp = wpn;
                    if (j + 1 < j1) {
                        b = ce;
                        e = s_off[j + 2];
                        wn = -1;
                        wpn = -1;
                        if (b + t < e) {
                            wn = st[b + t];
                            if (t > 0) wpn = st[b + t - 1];
                        }
                    }
                    const unsigned jl = (unsigned)(c0 - d_lo) + (unsigned)j;
                    if (cb + t < ce) {
                        unordered |= (w <= wp);  // wp = -1 for a document's first entry
                        place(w, base + cb + t, jl);
                    }
                    for (int i = cb + t + kWalkThreads; i < ce; i += kWalkThreads) {  // long documents
                        const int wl = st[i];
                        unordered |= (wl <= st[i - 1]);
                        place(wl, base + i, jl);
                    }
                    __syncthreads();  // document j is counted before document j + 1 reads the counters
                }
            }
            j0 = j1;
            j1 = j2;
            long0 = long1;
            lead0 = lead1;
            buf ^= 1;
        }
        walk_cp_wait<0>();
    }
    if (!ATOMIC && unordered) *redo = 1;
    __syncthreads();
    uint32_t* out = hist32 + (size_t)blockIdx.x * (size_t)(vp >> 1) + (wlo >> 1);
    for (int i = t * 4; i < half; i += kWalkThreads * 4)
        *reinterpret_cast<uint4*>(out + i) = *reinterpret_cast<const uint4*>(walk_tab + i);
}

// 2a: per word, exclusive prefix of the slice counts and the total; one thread per PAIR of words (the packed counters);
// bsum[CTA] = the sum of the CTA's 512 totals (for 2b)
__global__ void __launch_bounds__(256) csc_walk_scan_kernel(const uint32_t* __restrict__ hist32, int64_t slices, int64_t vp,
                                                            uint32_t* __restrict__ off, uint32_t* __restrict__ total,
                                                            int64_t* __restrict__ bsum)
{
    __shared__ int64_t wsum[8];
    const int64_t half = vp >> 1;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t r0 = 0, r1 = 0;
    if (i < half) {
        int64_t s = 0;
        for (; s + 16 <= slices; s += 16) {
            uint32_t h[16];
#pragma unroll
            for (int u = 0; u < 16; ++u) h[u] = hist32[(s + u) * half + i];
#pragma unroll
            for (int u = 0; u < 16; ++u) {
                *reinterpret_cast<uint2*>(off + (s + u) * vp + 2 * i) = make_uint2(r0, r1);
                r0 += h[u] & 0xffffu;
                r1 += h[u] >> 16;
            }
        }
        for (; s < slices; ++s) {
            const uint32_t h = hist32[s * half + i];
            *reinterpret_cast<uint2*>(off + s * vp + 2 * i) = make_uint2(r0, r1);
            r0 += h & 0xffffu;
            r1 += h >> 16;
        }
        *reinterpret_cast<uint2*>(total + 2 * i) = make_uint2(r0, r1);
    }
    int64_t x = (int64_t)r0 + (int64_t)r1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = x;
    __syncthreads();
    if (threadIdx.x == 0) {
        int64_t sum = 0;
        for (int k = 0; k < 8; ++k) sum += wsum[k];
        bsum[blockIdx.x] = sum;
    }
}

// 2b: colptr[w] = pos_base + sum_{w' < w} total[w'] (never beyond the block: pos_base + nnz), colptr[V] = pos_base + nnz.
// Same CTA partition as 2a: a CTA adds up the sums of the CTAs before it, then scans its own 512 totals.
__global__ void __launch_bounds__(256) csc_colptr_kernel(const uint32_t* __restrict__ total,
                                                         const int64_t* __restrict__ bsum, int64_t V, int64_t vp,
                                                         int64_t pos_base, int64_t nnz, int64_t* __restrict__ colptr)
{
    __shared__ int64_t wsum[8];
    __shared__ int64_t before_s;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int64_t before = 0;
    for (int c = tid; c < (int)blockIdx.x; c += 256) before += bsum[c];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) before += __shfl_xor_sync(0xffffffffu, before, o);
    if (lane == 0) wsum[warp] = before;
    __syncthreads();
    if (tid == 0) {
        int64_t sum = 0;
        for (int k = 0; k < 8; ++k) sum += wsum[k];
        before_s = sum;
    }
    __syncthreads();
    before = before_s;
    const int64_t i = (int64_t)blockIdx.x * 256 + tid;  // word pair
    uint2 tt = make_uint2(0u, 0u);
    if (2 * i < vp) tt = *reinterpret_cast<const uint2*>(total + 2 * i);
    const int64_t v = (int64_t)tt.x + (int64_t)tt.y;
    int64_t x = v;  // inclusive scan over the CTA's pairs
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int64_t y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    __syncthreads();
    if (lane == 31) wsum[warp] = x;
    __syncthreads();
    int64_t wbefore = 0;
    for (int k = 0; k < warp; ++k) wbefore += wsum[k];
    const int64_t excl = before + wbefore + (x - v);
    if (2 * i < V) colptr[2 * i] = pos_base + (excl < nnz ? excl : nnz);
    if (2 * i + 1 < V) {
        const int64_t e1 = excl + (int64_t)tt.x;
        colptr[2 * i + 1] = pos_base + (e1 < nnz ? e1 : nnz);
    }
    if (blockIdx.x == gridDim.x - 1 && tid == 0) colptr[V] = pos_base + nnz;
}

// 3 (generic form, vocabularies whose 32-bit bases do not fit shared memory): one warp per document
__global__ void __launch_bounds__(256) csc_walk_place_kernel(const int64_t* __restrict__ rowptr,
                                                             const int32_t* __restrict__ colidx,
                                                             const int32_t* __restrict__ counts, int64_t D, int64_t V,
                                                             int64_t vp, int64_t dps, int64_t doc_base, int64_t pos_base,
                                                             int64_t nnz, const uint32_t* __restrict__ rank,
                                                             const uint32_t* __restrict__ off,
                                                             const int64_t* __restrict__ colptr, int32_t* __restrict__ docidx,
                                                             int32_t* __restrict__ ccounts, uint32_t* __restrict__ csr2csc)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t d = warp0; d < D; d += nwarps) {
        const int64_t b = rowptr[d] - pos_base, e = rowptr[d + 1] - pos_base;
        const uint32_t* offs = off + (d / dps) * vp;
        for (int64_t p = b + lane; p < e; p += 32) {
            const int w = colidx[p];
            int64_t q = p;  // a word id outside [0, V) (malformed, and not sanitised by the caller): stay in bounds
            if ((unsigned)w < (unsigned)V) {
                q = (colptr[w] - pos_base) + (int64_t)offs[w] + (int64_t)(rank[p] & 0xffffu);
                if (q >= nnz) q = nnz - 1;  // only a corpus that overflows a 16-bit counter (> 65 535 repeats) gets here
                docidx[q] = (int32_t)(doc_base + d);
                if (ccounts) ccounts[q] = counts[p];
            }
            if (csr2csc) csr2csc[p] = (uint32_t)(pos_base + q);
        }
    }
}

// 3, when a 32-bit entry per word fits shared memory (V <= ~55 000): one CTA per SLICE first stages the slice's placement
// bases colptr[w] + off[slice][w] with coalesced loads (the generic form above fetches them with two dependent random
// 4/8-byte reads per entry: a 32-byte sector each), then runs FLAT over the slice's entries — the walk left the document's
// index inside the slice next to the rank, so an entry needs nothing but its own (word, rank | doc) pair.  Four entries per
// thread are processed while the next four are in flight.
__global__ void __launch_bounds__(kPlaceThreads) csc_walk_place_slice_kernel(
    const int64_t* __restrict__ rowptr, const int32_t* __restrict__ colidx, const int32_t* __restrict__ counts, int64_t D,
    int64_t V, int64_t vp, int64_t dps, int64_t doc_base, int64_t pos_base, int64_t nnz, const uint32_t* __restrict__ rank,
    const uint32_t* __restrict__ off, const int64_t* __restrict__ colptr, int32_t* __restrict__ docidx,
    int32_t* __restrict__ ccounts, uint32_t* __restrict__ csr2csc)
{
    extern __shared__ __align__(16) uint32_t place_base[];  // vp words
    const int t = threadIdx.x;
    const int64_t d_lo = (int64_t)blockIdx.x * dps;
    const int64_t d_hi = (d_lo + dps < D) ? d_lo + dps : D;
    const int p0 = (int)(rowptr[d_lo] - pos_base), p1 = (int)(rowptr[d_hi] - pos_base);
    constexpr int U = 4, STEP = U * kPlaceThreads;
    int w[U], rk[U], c[U];
    auto fetch = [&](int pb) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int p = pb + u * kPlaceThreads;
            w[u] = -1;
            rk[u] = 0;
            c[u] = 0;
            if (p < p1) {
                w[u] = colidx[p];
                rk[u] = (int)rank[p];
                if (ccounts) c[u] = counts[p];
            }
        }
    };
    fetch(p0 + t);  // in flight while the bases are staged
    const uint32_t* offs = off + (size_t)blockIdx.x * vp;
    for (int64_t w0 = t; w0 < V; w0 += 8 * kPlaceThreads) {  // eight independent (colptr, off) pairs in flight per thread
        int64_t cp[8];
        uint32_t of[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int64_t ww = w0 + u * kPlaceThreads;
            cp[u] = 0;
            of[u] = 0;
            if (ww < V) {
                cp[u] = colptr[ww];
                of[u] = offs[ww];
            }
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int64_t ww = w0 + u * kPlaceThreads;
            if (ww < V) place_base[ww] = (uint32_t)(cp[u] - pos_base) + of[u];
        }
    }
    __syncthreads();
    for (int pb = p0 + t; pb < p1; pb += STEP) {
        int cw[U], crk[U], cc[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            cw[u] = w[u];
            crk[u] = rk[u];
            cc[u] = c[u];
        }
        fetch(pb + STEP);  // the next four entries: in flight while these four are placed
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int p = pb + u * kPlaceThreads;
            if (p >= p1) continue;
            int64_t q = p;  // a word id outside [0, V): stay in bounds (see the generic form)
            if ((unsigned)cw[u] < (unsigned)V) {
                q = (int64_t)place_base[cw[u]] + (int64_t)((unsigned)crk[u] & 0xffffu);
                if (q >= nnz) q = nnz - 1;
                docidx[q] = (int32_t)(doc_base + d_lo + (int64_t)((unsigned)crk[u] >> 16));
                if (ccounts) ccounts[q] = cc[u];
            }
            if (csr2csc) csr2csc[p] = (uint32_t)(pos_base + q);
        }
    }
}

static int csr_to_csc_walk(const int64_t* rowptr, const int32_t* colidx, const int32_t* counts, int64_t D, int64_t V,
                           int64_t nnz, int64_t doc_base, int64_t pos_base, int64_t* colptr, int32_t* docidx,
                           int32_t* ccounts, uint32_t* csr2csc, void* workspace, int64_t workspace_bytes, cudaStream_t s)
{
    const WalkPlan w = walk_plan(D, V, nnz);
    if ((int64_t)w.total > workspace_bytes) {
        set_error("csr_to_csc: workspace %lld < %lld", (long long)workspace_bytes, (long long)w.total);
        return TTLDA_EWORKSPACE;
    }
    char* base = (char*)workspace;
    uint32_t* rank = (uint32_t*)base;  // (document index inside the slice << 16) | rank inside the slice, per entry
    uint32_t* hist32 = (uint32_t*)(base + w.rank_b);
    uint32_t* off = (uint32_t*)(base + w.rank_b + w.hist_b);
    uint32_t* total = (uint32_t*)(base + w.rank_b + w.hist_b + w.off_b);
    int32_t* redo = (int32_t*)(total + w.vp);  // "a row repeats or misorders words: redo the walk with atomics"
    int64_t* bsum = (int64_t*)((char*)total + align256((size_t)w.vp * 4 + 16));
    const unsigned scan_grid = (unsigned)((w.vp / 2 + 255) / 256);
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(csc_walk_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             kWalkSmemSM - 4 * 1024);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(csc_walk_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     kWalkSmemSM - 4 * 1024);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(csc_walk_place_slice_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     kWalkSmemSM - 12 * 1024);
        if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(csc_walk)");
        configured = true;
    }
    if (D > 0 && nnz > 0) {
        cudaError_t e = cudaMemsetAsync(redo, 0, sizeof(int32_t), s);
        if (e != cudaSuccess) return cuda_fail(e, "memset redo flag");
        const dim3 wgrid((unsigned)w.slices, (unsigned)w.parts);
        csc_walk_kernel<false><<<wgrid, kWalkThreads, w.smem, s>>>(rowptr, colidx, D, V, w.vp, w.partw, w.dps, w.stage_cap,
                                                                  doc_base, pos_base, rank, hist32, docidx, redo);
        TTLDA_LAUNCH_CHECK("csc_walk");
        csc_walk_kernel<true><<<wgrid, kWalkThreads, w.smem, s>>>(rowptr, colidx, D, V, w.vp, w.partw, w.dps, w.stage_cap,
                                                                 doc_base, pos_base, rank, hist32, docidx, redo);
        TTLDA_LAUNCH_CHECK("csc_walk (atomic redo)");
        csc_walk_scan_kernel<<<scan_grid, 256, 0, s>>>(hist32, w.slices, w.vp, off, total, bsum);
        TTLDA_LAUNCH_CHECK("csc_walk_scan");
    } else {
        cudaError_t e = cudaMemsetAsync(total, 0, (size_t)((char*)(bsum + scan_grid) - (char*)total), s);
        if (e != cudaSuccess) return cuda_fail(e, "memset totals");
    }
    csc_colptr_kernel<<<scan_grid, 256, 0, s>>>(total, bsum, V, w.vp, pos_base, nnz, colptr);
    TTLDA_LAUNCH_CHECK("csc_colptr");
    if (D > 0 && nnz > 0) {
        const size_t place_smem = (size_t)w.vp * 4;
        // (staging V bases per slice only pays when the slice has entries to place with them)
        if (place_smem + 12 * 1024 <= (size_t)kWalkSmemSM && nnz / w.slices >= w.vp / 4) {
            csc_walk_place_slice_kernel<<<(unsigned)w.slices, kPlaceThreads, place_smem, s>>>(
                rowptr, colidx, counts, D, V, w.vp, w.dps, doc_base, pos_base, nnz, rank, off, colptr, docidx, ccounts,
                csr2csc);
        } else {
            csc_walk_place_kernel<<<grid_for_warps(D, 8, 16), 256, 0, s>>>(rowptr, colidx, counts, D, V, w.vp, w.dps,
                                                                           doc_base, pos_base, nnz, rank, off, colptr,
                                                                           docidx, ccounts, csr2csc);
        }
        TTLDA_LAUNCH_CHECK("csc_walk_place");
    }
    return TTLDA_OK;
}

}  // namespace ttlda

using namespace ttlda;

extern "C" {

int64_t ttlda_bow_to_csr_workspace_bytes(int64_t T, int64_t D, int64_t V)
{
    if (T < 0 || D < 0 || V < 0) return TTLDA_EINVAL;
    BowWs w;
    cudaError_t e = bow_layout(T, D, V, nullptr, w);
    if (e != cudaSuccess) return cuda_fail(e, "bow_to_csr workspace query");
    return (int64_t)w.total;
}

int ttlda_bow_to_csr(const int32_t* token_ids, const int64_t* doc_offsets, int64_t T, int64_t D, int64_t V,
                     int64_t* rowptr, int32_t* colidx, int32_t* counts, int64_t* nnz_out, void* workspace,
                     int64_t workspace_bytes, void* stream)
{
    TTLDA_REQUIRE(doc_offsets && rowptr && nnz_out && workspace, "NULL pointer");
    TTLDA_REQUIRE(T == 0 || (token_ids && colidx && counts), "NULL pointer");
    TTLDA_REQUIRE(T >= 0 && D >= 0 && V >= 0, "negative size");
    TTLDA_REQUIRE(T < ((int64_t)1 << 31), "token stream per call must be < 2^31 (split the corpus)");
    TTLDA_REQUIRE(bits_for(V) + bits_for(D) <= 64, "D*V key does not fit 64 bits");
    cudaStream_t s = (cudaStream_t)stream;
    BowWs w;
    cudaError_t e = bow_layout(T, D, V, (char*)workspace, w);
    if (e != cudaSuccess) return cuda_fail(e, "bow_to_csr layout");
    if ((int64_t)w.total > workspace_bytes) {
        set_error("bow_to_csr: workspace %lld < %lld", (long long)workspace_bytes, (long long)w.total);
        return TTLDA_EWORKSPACE;
    }
    const int vbits = bits_for(V);
    if (T > 0) {
        pack_keys_kernel<<<grid_1d(T), 256, 0, s>>>(token_ids, doc_offsets, T, D, vbits, w.keys);
        TTLDA_LAUNCH_CHECK("pack_keys");
        e = cub::DeviceRadixSort::SortKeys(w.cub, w.cub_bytes, w.keys, w.skeys, T, 0, vbits + bits_for(D), s);
        if (e != cudaSuccess) return cuda_fail(e, "radix sort (bow)");
        e = cub::DeviceRunLengthEncode::Encode(w.cub, w.cub_bytes, w.skeys, w.ukeys, counts, w.nruns, (int)T, s);
        if (e != cudaSuccess) return cuda_fail(e, "run-length encode");
    } else {
        e = cudaMemsetAsync(w.nruns, 0, sizeof(int64_t), s);
        if (e != cudaSuccess) return cuda_fail(e, "memset");
    }
    split_runs_kernel<<<grid_1d(T + 1), 256, 0, s>>>(w.ukeys, w.nruns, D, vbits, colidx, rowptr, nnz_out);
    TTLDA_LAUNCH_CHECK("split_runs");
    return TTLDA_OK;
}

int64_t ttlda_csr_to_csc_workspace_bytes(int64_t nnz, int64_t D, int64_t V, int64_t nblocks)
{
    if (nnz < 0 || V < 0 || D < 0 || nblocks < 1) return TTLDA_EINVAL;
    CscWs w;
    cudaError_t e = csc_layout(nnz, nblocks * V, nullptr, w);
    if (e != cudaSuccess) return cuda_fail(e, "csr_to_csc workspace query");
    size_t need = w.total;
    if (nblocks == 1) {  // whichever path runs (slice walk / pair sort): the larger of the two layouts
        const WalkPlan t = walk_plan(D, V, nnz);
        if (t.ok && t.total > need) need = t.total;
    }
    return (int64_t)need;
}

static int csr_to_csc_impl(const int64_t* rowptr, const int32_t* colidx, const int32_t* counts, int64_t D, int64_t V,
                           int64_t nnz, int64_t nblocks, int64_t block_docs, int64_t doc_base, int64_t pos_base,
                           int64_t* colptr, int32_t* docidx, int32_t* ccounts, uint32_t* csr2csc, void* workspace,
                           int64_t workspace_bytes, cudaStream_t s)
{
    if (walk_path_applies(D, V, nnz, nblocks))
        return csr_to_csc_walk(rowptr, colidx, counts, D, V, nnz, doc_base, pos_base, colptr, docidx, ccounts, csr2csc,
                                workspace, workspace_bytes, s);
    const int64_t VV = nblocks * V;
    CscWs w;
    cudaError_t e = csc_layout(nnz, VV, (char*)workspace, w);
    if (e != cudaSuccess) return cuda_fail(e, "csr_to_csc layout");
    if ((int64_t)w.total > workspace_bytes) {
        set_error("csr_to_csc: workspace %lld < %lld", (long long)workspace_bytes, (long long)w.total);
        return TTLDA_EWORKSPACE;
    }
    if (nnz > 0) {
        csc_keys_kernel<<<grid_for_warps(D, 8, 16), 256, 0, s>>>(rowptr, colidx, D, V, nnz, block_docs, doc_base,
                                                                  pos_base, w.keys, w.docs, w.pos);
        TTLDA_LAUNCH_CHECK("csc_keys");
        // LSD radix sort is stable: within a (block, word), CSR positions (hence documents) stay ascending.
        e = cub::DeviceRadixSort::SortPairs(w.cub, w.cub_bytes, w.keys, w.skeys, w.pos, w.spos, nnz, 0, bits_for(VV), s);
        if (e != cudaSuccess) return cuda_fail(e, "radix sort (csc)");
    }
    csc_gather_kernel<<<grid_1d(nnz + 1), 256, 0, s>>>(w.skeys, w.spos, w.docs, counts, VV, nnz, pos_base, colptr,
                                                       docidx, ccounts, csr2csc);
    TTLDA_LAUNCH_CHECK("csc_gather");
    return TTLDA_OK;
}

int ttlda_csr_to_csc(const int64_t* rowptr, const int32_t* colidx, const int32_t* counts, int64_t D, int64_t V,
                     int64_t nnz, int64_t nblocks, int64_t block_docs, int64_t* colptr, int32_t* docidx,
                     int32_t* ccounts, uint32_t* csr2csc, void* workspace, int64_t workspace_bytes, void* stream)
{
    TTLDA_REQUIRE(rowptr && colptr && workspace, "NULL pointer");
    TTLDA_REQUIRE(nnz == 0 || (colidx && counts && docidx && ccounts), "NULL pointer");
    TTLDA_REQUIRE(nnz >= 0 && D >= 0 && V >= 0, "negative size");
    TTLDA_REQUIRE(nnz < ((int64_t)1 << 32), "nnz per device must be < 2^32");
    TTLDA_REQUIRE(nblocks >= 1 && block_docs >= 1 && nblocks * block_docs >= D, "document blocks must cover D");
    TTLDA_REQUIRE(nblocks * V < ((int64_t)1 << 31), "nblocks * V must be < 2^31");
    return csr_to_csc_impl(rowptr, colidx, counts, D, V, nnz, nblocks, block_docs, 0, 0, colptr, docidx, ccounts, csr2csc,
                           workspace, workspace_bytes, (cudaStream_t)stream);
}

int ttlda_csr_to_csc_block(const int64_t* rowptr_block, const int32_t* colidx, const int32_t* counts, int64_t D_block,
                           int64_t V, int64_t doc_base, int64_t pos_base, int64_t nnz_block, int64_t* colptr_block,
                           int32_t* docidx, int32_t* ccounts, uint32_t* csr2csc, void* workspace,
                           int64_t workspace_bytes, void* stream)
{
    TTLDA_REQUIRE(rowptr_block && colptr_block && workspace, "NULL pointer");
    TTLDA_REQUIRE(nnz_block == 0 || (colidx && docidx && (counts || !ccounts)), "NULL pointer");
    TTLDA_REQUIRE(nnz_block >= 0 && D_block >= 0 && V >= 0 && doc_base >= 0 && pos_base >= 0, "negative size");
    TTLDA_REQUIRE(pos_base + nnz_block < ((int64_t)1 << 32), "nnz per device must be < 2^32");
    TTLDA_REQUIRE(doc_base + D_block < ((int64_t)1 << 31), "documents per device must be < 2^31");
    TTLDA_REQUIRE(V < ((int64_t)1 << 31), "V must be < 2^31");
    // the block's entries occupy [pos_base, pos_base + nnz_block) of the corpus-wide arrays, in CSR and mirror alike
    return csr_to_csc_impl(rowptr_block, colidx ? colidx + pos_base : colidx, counts ? counts + pos_base : counts, D_block,
                           V, nnz_block, 1, D_block > 0 ? D_block : 1, doc_base, pos_base, colptr_block,
                           docidx ? docidx + pos_base : docidx, ccounts ? ccounts + pos_base : ccounts,
                           csr2csc ? csr2csc + pos_base : csr2csc, workspace, workspace_bytes, (cudaStream_t)stream);
}

int ttlda_widen_i32(const void* src, int64_t src_bytes_per_item, int32_t* dst, int64_t n, int64_t bound,
                    int32_t* flags, void* stream)
{
    TTLDA_REQUIRE(n >= 0 && (n == 0 || (src && dst)), "NULL pointer");
    TTLDA_REQUIRE(src_bytes_per_item == 1 || src_bytes_per_item == 2, "source items must be uint8 or uint16");
    TTLDA_REQUIRE(bound >= 0 && bound <= 65536 && (bound == 0 || flags), "bound needs a flags word; items are < 65536");
    if (n == 0) return TTLDA_OK;
    if (src_bytes_per_item == 1)
        widen_kernel<uint8_t><<<grid_1d(n), 256, 0, (cudaStream_t)stream>>>((const uint8_t*)src, dst, n,
                                                                            (uint32_t)bound, flags);
    else
        widen_kernel<uint16_t><<<grid_1d(n), 256, 0, (cudaStream_t)stream>>>((const uint16_t*)src, dst, n,
                                                                             (uint32_t)bound, flags);
    TTLDA_LAUNCH_CHECK("widen");
    return TTLDA_OK;
}

int ttlda_validate_csr(int64_t* rowptr, int64_t D, int64_t pos_base, int32_t* colidx, int32_t* counts, int64_t nnz,
                       int64_t V, int sanitize, int32_t* flags, void* stream)
{
    TTLDA_REQUIRE(flags, "NULL flags pointer");
    TTLDA_REQUIRE(D >= 0 && nnz >= 0 && V >= 0 && pos_base >= 0, "negative size");
    TTLDA_REQUIRE(nnz == 0 || colidx || rowptr, "nothing to validate");
    const int64_t n = (rowptr ? D + 1 : 0) > (colidx ? nnz : 0) ? D + 1 : nnz;
    if (n == 0) return TTLDA_OK;
    validate_csr_kernel<<<grid_1d(n), 256, 0, (cudaStream_t)stream>>>(rowptr, D, pos_base, colidx, counts, nnz, V,
                                                                      sanitize, flags);
    TTLDA_LAUNCH_CHECK("validate_csr");
    return TTLDA_OK;
}

}  // extern "C"
