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
// Mirror of ONE document block by a hand-written, comparison-free, atomic-order-free counting sort: the "slice walk".
// (Default for blocks whose vocabulary fits the shared-memory counter table, V <= ~115 000; TTLDA_CSC_WALK=0 or a larger
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
//            deterministic.  (The increment is a shared-memory atomicAdd on the 32-bit word that holds two counters — so a
//            malformed row with a repeated word still gets distinct, in-range ranks.)  The entries of the next kPF
//            documents are already in registers when their turn comes.  At the end the counter table IS the slice's word
//            histogram: hist[slice][w].
//   2 scan   per word, the exclusive prefix of hist over the slices -> off[slice][w]; the totals -> colptr (one CTA).
//   3 place  one warp per document: q = colptr[w] + off[slice(d)][w] + rank[p];  docidx[q] = d, (ccounts[q] = counts[p],)
//            csr2csc[p] = q — csr2csc is one coalesced stream in CSR order.
// Against the generic path (key packing, histogram, two 8-bit onesweep passes over (key, value) pairs, a dependent random
// gather of doc ids through the sorted permutation: ~0.68 ms per 12.7M-entry block at cfg3) there is no pair sort, no
// 20-bytes-per-entry scratch (2 bytes per entry + 6 bytes per (slice, word)), and the only serial part — the walk — is
// ~200 barrier-separated steps per CTA.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kWalkThreads = 512;  // entries of one document handled per step (longer documents loop)
constexpr int kWalkPF = 8;         // documents whose entries are prefetched into registers ahead of the walk
constexpr int kWalkSmemSM = 227 * 1024;

struct WalkPlan {
    bool ok;
    int64_t vp, dps, slices;
    size_t smem, rank_b, hist_b, off_b, total_b, total;
};

static WalkPlan walk_plan(int64_t D, int64_t V, int64_t nnz)
{
    WalkPlan w{};
    w.vp = (V + 7) / 8 * 8;  // counters are cleared / flushed as 16-byte vectors of eight
    w.smem = (size_t)w.vp * 2;
    // two CTAs per SM when two tables (+ 2 KB of static shared memory and the per-CTA reserve each) fit
    const int per_sm = (2 * (w.smem + 3 * 1024) <= (size_t)kWalkSmemSM) ? 2 : 1;
    const int64_t target = (int64_t)kSMs * per_sm;
    w.dps = (D + target - 1) / target;
    if (w.dps < 16) w.dps = 16;
    w.slices = D > 0 ? (D + w.dps - 1) / w.dps : 1;
    w.ok = V >= 1 && w.smem + 3 * 1024 <= (size_t)kWalkSmemSM && w.dps <= 65535 && nnz < ((int64_t)1 << 31) &&
           target * w.vp < ((int64_t)1 << 32);
    // buffer sizes from a bound on the slice count that is monotone in D (the workspace is sized for the largest block, a
    // call may bring a smaller one: slices <= target and slices <= ceil(D / 16) always hold)
    int64_t cap = (D + 15) / 16;
    if (cap > target) cap = target;
    if (cap < 1) cap = 1;
    w.rank_b = align256((size_t)(nnz > 0 ? nnz : 1) * 2);
    w.hist_b = align256((size_t)cap * w.vp * 2);
    w.off_b = align256((size_t)cap * w.vp * 4);
    w.total_b = align256((size_t)w.vp * 4 + 16);  // + the redo flag
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

// 1: the walk.  rank[p] = documents of the slice before p's that contain p's word;  hist[slice][*] = the slice's word counts.
// Also fills docidx[.] of the slice's entry range with a valid document id, so that whatever a malformed corpus makes of
// the placement (3), no slot of the mirror is left holding garbage that a gather would follow.
//
// ATOMIC = false (the fast form): the counter update is a plain 16-bit LDS + STS — a shared-memory atomic costs ~2 cycles
// per LANE on this part (64 per warp instruction) and made the walk 3x slower than the rest of the sort.  That is only
// right when no document repeats a word, which holds for every row the reference can produce (np.nonzero) and is what
// Document.from_csr checks; the kernel checks it again on the fly (each entry against its predecessor: strictly ascending)
// and raises *redo if a row fails.  ATOMIC = true is the same walk with atomicAdd on the 32-bit word that holds two
// counters: launched right behind the fast form, it returns at once unless *redo is set, and then recomputes the block's
// ranks and histograms correctly for ANY row order / repetition (distinct, in-range ranks).
template <bool ATOMIC>
__global__ void __launch_bounds__(kWalkThreads) csc_walk_kernel(const int64_t* __restrict__ rowptr,
                                                                const int32_t* __restrict__ colidx, int64_t D, int64_t V,
                                                                int64_t vp, int64_t dps, int64_t doc_base, int64_t pos_base,
                                                                uint16_t* __restrict__ rank, uint32_t* __restrict__ hist32,
                                                                int32_t* __restrict__ docidx, int32_t* __restrict__ redo)
{
    extern __shared__ __align__(16) uint32_t walk_tab[];  // vp / 2 words, two 16-bit counters each
    __shared__ int s_off[kWalkThreads + 1];               // entry offsets of up to kWalkThreads documents
    if (ATOMIC && *redo == 0) return;
    const int t = threadIdx.x;
    const int half = (int)(vp >> 1);
    const int64_t d_lo = (int64_t)blockIdx.x * dps;
    const int64_t d_hi = (d_lo + dps < D) ? d_lo + dps : D;
    for (int i = t * 4; i < half; i += kWalkThreads * 4) *reinterpret_cast<uint4*>(walk_tab + i) = make_uint4(0u, 0u, 0u, 0u);
    uint16_t* tab16 = reinterpret_cast<uint16_t*>(walk_tab);
    bool unordered = false;

    auto place = [&](int w, int64_t p) {
        if (!ATOMIC) docidx[p] = (int32_t)doc_base;
        unsigned r = 0;
        if ((unsigned)w < (unsigned)V) {
            if (ATOMIC) {
                const unsigned sh = (unsigned)(w & 1) << 4;
                r = atomicAdd(walk_tab + (w >> 1), 1u << sh) >> sh;
            } else {
                r = tab16[w];
                tab16[w] = (uint16_t)(r + 1);
            }
        }
        rank[p] = (uint16_t)r;
    };

    for (int64_t c0 = d_lo; c0 < d_hi; c0 += kWalkThreads) {
        const int n = (int)((d_hi - c0 < kWalkThreads) ? d_hi - c0 : kWalkThreads);
        __syncthreads();  // the table is cleared / the previous chunk's offsets are no longer read
        const int64_t base = rowptr[c0] - pos_base;
        for (int i = t; i <= n; i += kWalkThreads) s_off[i] = (int)(rowptr[c0 + i] - pos_base - base);
        __syncthreads();
        const int32_t* ci = colidx + base;
        int wq[kWalkPF], pq[kWalkPF];  // this thread's entry of the next kWalkPF documents, and the entry before it
#pragma unroll
        for (int u = 0; u < kWalkPF; ++u) {
            wq[u] = -1;
            pq[u] = -1;
            if (u < n && s_off[u] + t < s_off[u + 1]) {
                wq[u] = ci[s_off[u] + t];
                if (!ATOMIC && t > 0) pq[u] = ci[s_off[u] + t - 1];
            }
        }
        for (int j0 = 0; j0 < n; j0 += kWalkPF) {
#pragma unroll
            for (int u = 0; u < kWalkPF; ++u) {
                const int j = j0 + u;
                if (j < n) {  // CTA-uniform
                    const int b = s_off[j], e = s_off[j + 1];
                    const int w = wq[u], wp = pq[u];
                    const bool mine = b + t < e;
                    wq[u] = -1;
                    pq[u] = -1;
                    if (j + kWalkPF < n && s_off[j + kWalkPF] + t < s_off[j + kWalkPF + 1]) {
                        wq[u] = ci[s_off[j + kWalkPF] + t];
                        if (!ATOMIC && t > 0) pq[u] = ci[s_off[j + kWalkPF] + t - 1];
                    }
                    if (mine) {
                        unordered |= (w <= wp);  // wp = -1 for a document's first entry; ids are >= 0 (else flagged later)
                        place(w, base + b + t);
                    }
                    for (int i = b + t + kWalkThreads; i < e; i += kWalkThreads) {  // long documents
                        const int wl = ci[i];
                        unordered |= (wl <= ci[i - 1]);
                        place(wl, base + i);
                    }
                    __syncthreads();  // document j is counted before document j + 1 reads the counters
                }
            }
        }
    }
    if (!ATOMIC && unordered) *redo = 1;
    __syncthreads();
    uint32_t* out = hist32 + (size_t)blockIdx.x * half;
    for (int i = t * 4; i < half; i += kWalkThreads * 4)
        *reinterpret_cast<uint4*>(out + i) = *reinterpret_cast<const uint4*>(walk_tab + i);
}

// 2: per word, exclusive prefix of the slice counts and the total; one thread per PAIR of words (the packed counters)
__global__ void __launch_bounds__(256) csc_walk_scan_kernel(const uint32_t* __restrict__ hist32, int64_t slices, int64_t vp,
                                                            uint32_t* __restrict__ off, uint32_t* __restrict__ total)
{
    const int64_t half = vp >> 1;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= half) return;
    uint32_t r0 = 0, r1 = 0;
    int64_t s = 0;
    for (; s + 8 <= slices; s += 8) {
        uint32_t h[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) h[u] = hist32[(s + u) * half + i];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
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

// colptr[w] = pos_base + sum_{w' < w} total[w'] (never beyond the block: pos_base + nnz), colptr[V] = pos_base + nnz.
// One CTA scans the totals 1024 at a time (coalesced loads, warp shuffles, a running carry).
__global__ void __launch_bounds__(1024) csc_colptr_kernel(const uint32_t* __restrict__ total, int64_t V, int64_t pos_base,
                                                          int64_t nnz, int64_t* __restrict__ colptr)
{
    __shared__ int64_t wsum[32];
    __shared__ int64_t carry_s;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) carry_s = 0;
    __syncthreads();
    for (int64_t c0 = 0; c0 < V; c0 += 1024) {
        const int64_t w = c0 + tid;
        const int64_t v = (w < V) ? (int64_t)total[w] : 0;
        int64_t x = v;  // inclusive scan inside the warp
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int64_t y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) wsum[warp] = x;
        __syncthreads();
        if (warp == 0) {
            int64_t s = wsum[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int64_t y = __shfl_up_sync(0xffffffffu, s, o);
                if (lane >= o) s += y;
            }
            wsum[lane] = s;  // inclusive over the warps
        }
        __syncthreads();
        const int64_t carry = carry_s;
        const int64_t excl = carry + (warp ? wsum[warp - 1] : 0) + (x - v);
        if (w < V) colptr[w] = pos_base + (excl < nnz ? excl : nnz);
        __syncthreads();
        if (tid == 1023) carry_s = carry + wsum[31];
        __syncthreads();
    }
    if (tid == 0) colptr[V] = pos_base + nnz;
}

// 3: placement, one warp per document
__global__ void __launch_bounds__(256) csc_walk_place_kernel(const int64_t* __restrict__ rowptr,
                                                             const int32_t* __restrict__ colidx,
                                                             const int32_t* __restrict__ counts, int64_t D, int64_t V,
                                                             int64_t vp, int64_t dps, int64_t doc_base, int64_t pos_base,
                                                             int64_t nnz, const uint16_t* __restrict__ rank,
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
                q = (colptr[w] - pos_base) + (int64_t)offs[w] + (int64_t)rank[p];
                if (q >= nnz) q = nnz - 1;  // only a corpus that overflows a 16-bit counter (> 65 535 repeats) gets here
                docidx[q] = (int32_t)(doc_base + d);
                if (ccounts) ccounts[q] = counts[p];
            }
            if (csr2csc) csr2csc[p] = (uint32_t)(pos_base + q);
        }
    }
}

// 3, when a 32-bit entry per word fits shared memory (V <= ~57 000): one CTA per SLICE first stages the slice's placement
// bases colptr[w] + off[slice][w] with coalesced loads (the generic form above fetches them with two dependent random
// 4/8-byte reads per entry: a 32-byte sector each), then places its documents' entries from shared memory.
constexpr int kPlaceThreads = 1024;
__global__ void __launch_bounds__(kPlaceThreads) csc_walk_place_slice_kernel(
    const int64_t* __restrict__ rowptr, const int32_t* __restrict__ colidx, const int32_t* __restrict__ counts, int64_t D,
    int64_t V, int64_t vp, int64_t dps, int64_t doc_base, int64_t pos_base, int64_t nnz, const uint16_t* __restrict__ rank,
    const uint32_t* __restrict__ off, const int64_t* __restrict__ colptr, int32_t* __restrict__ docidx,
    int32_t* __restrict__ ccounts, uint32_t* __restrict__ csr2csc)
{
    extern __shared__ __align__(16) uint32_t place_base[];  // vp words
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const uint32_t* offs = off + (size_t)blockIdx.x * vp;
    for (int64_t w = t; w < V; w += kPlaceThreads) place_base[w] = (uint32_t)(colptr[w] - pos_base) + offs[w];
    __syncthreads();
    const int64_t d_lo = (int64_t)blockIdx.x * dps;
    const int64_t d_hi = (d_lo + dps < D) ? d_lo + dps : D;
    for (int64_t d = d_lo + warp; d < d_hi; d += kPlaceThreads / 32) {
        const int64_t b = rowptr[d] - pos_base, e = rowptr[d + 1] - pos_base;
        for (int64_t p = b + lane; p < e; p += 32) {
            const int w = colidx[p];
            int64_t q = p;  // a word id outside [0, V): stay in bounds (see the generic form)
            if ((unsigned)w < (unsigned)V) {
                q = (int64_t)place_base[w] + (int64_t)rank[p];
                if (q >= nnz) q = nnz - 1;
                docidx[q] = (int32_t)(doc_base + d);
                if (ccounts) ccounts[q] = counts[p];
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
    uint16_t* rank = (uint16_t*)base;
    uint32_t* hist32 = (uint32_t*)(base + w.rank_b);
    uint32_t* off = (uint32_t*)(base + w.rank_b + w.hist_b);
    uint32_t* total = (uint32_t*)(base + w.rank_b + w.hist_b + w.off_b);
    int32_t* redo = (int32_t*)(total + w.vp);  // "a row repeats or misorders words: redo the walk with atomics"
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(csc_walk_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             kWalkSmemSM - 3 * 1024);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(csc_walk_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     kWalkSmemSM - 3 * 1024);
        if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(csc_walk)");
        configured = true;
    }
    if (D > 0 && nnz > 0) {
        cudaError_t e = cudaMemsetAsync(redo, 0, sizeof(int32_t), s);
        if (e != cudaSuccess) return cuda_fail(e, "memset redo flag");
        csc_walk_kernel<false><<<(unsigned)w.slices, kWalkThreads, w.smem, s>>>(rowptr, colidx, D, V, w.vp, w.dps, doc_base,
                                                                                  pos_base, rank, hist32, docidx, redo);
        TTLDA_LAUNCH_CHECK("csc_walk");
        csc_walk_kernel<true><<<(unsigned)w.slices, kWalkThreads, w.smem, s>>>(rowptr, colidx, D, V, w.vp, w.dps, doc_base,
                                                                                 pos_base, rank, hist32, docidx, redo);
        TTLDA_LAUNCH_CHECK("csc_walk (atomic redo)");
        csc_walk_scan_kernel<<<(unsigned)((w.vp / 2 + 255) / 256), 256, 0, s>>>(hist32, w.slices, w.vp, off, total);
        TTLDA_LAUNCH_CHECK("csc_walk_scan");
    } else {
        cudaError_t e = cudaMemsetAsync(total, 0, (size_t)w.vp * 4, s);
        if (e != cudaSuccess) return cuda_fail(e, "memset totals");
    }
    csc_colptr_kernel<<<1, 1024, 0, s>>>(total, V, pos_base, nnz, colptr);
    TTLDA_LAUNCH_CHECK("csc_colptr");
    if (D > 0 && nnz > 0) {
        const size_t place_smem = (size_t)w.vp * 4;
        // (staging V bases per slice only pays when the slice has entries to place with them)
        if (place_smem + 3 * 1024 <= (size_t)kWalkSmemSM && nnz / w.slices >= w.vp / 4) {
            static bool place_configured = false;
            if (!place_configured) {
                cudaError_t e = cudaFuncSetAttribute(csc_walk_place_slice_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                     kWalkSmemSM - 3 * 1024);
                if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(csc_walk_place_slice)");
                place_configured = true;
            }
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
