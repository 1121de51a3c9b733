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
// Mirror of ONE document block by a hand-written counting sort (opt-in, TTLDA_CSC_TILES=1, for blocks of <= 65 535
// documents; the CUB pair sort above is the default — see tile_path_applies for the measurement).
//
// A CSR block is already sorted by (document, word); its mirror is the same entries sorted by (word, document) — a
// transpose.  Keys are bounded by V and distinct within a document, so the sort needs no comparisons; what it needs is,
// for every entry, the number of EARLIER documents of the block that contain the same word (stability = documents
// ascending within a word).  That rank is made order-free, hence fully parallel and deterministic:
//   tiles   a CTA owns a tile of 32 consecutive documents; shared memory holds one 32-bit mask per word, bit i = "document
//           i of the tile contains the word", built with atomicOr (an OR commutes: any arrival order gives the same mask);
//           an entry's rank inside its tile is popc(mask & bits below its document).
//   A hist  hist[tile][w] = popc(mask[w])                                                        (uint8)
//   B scan  off[tile][w]  = sum over earlier tiles (uint16: a block holds < 65 536 documents); total[w]; colptr by an
//           exclusive scan of the totals
//   C place mirror position = colptr[w] + off[tile][w] + rank;  docidx / ccounts are written there, csr2csc[p] — one
//           coalesced stream in CSR order — records it.
// Vocabularies larger than the shared-memory mask array (58 112 words) are covered in word RANGES: the tile re-reads its
// entries once per range.  Against the generic path (key packing, two 8-bit onesweep passes over (key, value) pairs, a
// gather through the sorted permutation) there is no dependent random gather of 4-byte items and no 20-bytes-per-entry
// scratch.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kTileDocs = 32;
constexpr int kTileThreads = 512;                                // 16 warps: two documents of the tile per warp
constexpr int64_t kTileRangeWords = (227 * 1024 / 4) / 16 * 16;  // words per range: one 32-bit mask each

struct TilePlan {
    int64_t tiles, vp, range_words;
    int ranges;
    size_t hist_b, off_b, total_b, total;
};

static TilePlan tile_plan(int64_t D, int64_t V)
{
    TilePlan t;
    t.tiles = (D + kTileDocs - 1) / kTileDocs;
    t.vp = (V + 15) / 16 * 16;
    t.range_words = t.vp < kTileRangeWords ? (t.vp > 0 ? t.vp : 16) : kTileRangeWords;
    t.ranges = (int)((t.vp + t.range_words - 1) / t.range_words);
    if (t.ranges < 1) t.ranges = 1;
    t.hist_b = align256((size_t)(t.tiles > 0 ? t.tiles : 1) * t.vp);
    t.off_b = align256((size_t)(t.tiles > 0 ? t.tiles : 1) * t.vp * 2);
    t.total_b = align256((size_t)(V + 1) * 4);
    t.total = t.hist_b + t.off_b + t.total_b;
    return t;
}

static bool tile_path_applies(int64_t D, int64_t V, int64_t nblocks)
{
    if (nblocks != 1 || D > 65535 || V < 1) return false;
    // OPT-IN (TTLDA_CSC_TILES=1).  Same-box A/B of the streamed step at cfg3: 39.6 ms with this path against 36.3 ms
    // with the CUB pair sort (cfg4: 56.1 vs 54.2): per tile the mask array is cleared and scanned densely — O(V) work
    // for ~7k entries — and that outweighs the sort's two passes.  Kept because it is exact, allocation-light (0.3 GB
    // instead of 20 bytes per entry) and the right starting point for a sparse-tile version.
    const char* e = getenv("TTLDA_CSC_TILES");
    if (!e || e[0] != '1') return false;
    const TilePlan t = tile_plan(D, V);
    return t.total <= ((size_t)3 << 30);  // scratch cap; beyond it the generic sort's 20 bytes per entry are smaller
}

// MODE 0: histogram (kernel A);  MODE 1: placement (kernel C).  One CTA per tile of 32 documents.
template <int MODE>
__global__ void __launch_bounds__(kTileThreads) csc_tile_kernel(
    const int64_t* __restrict__ rowptr, const int32_t* __restrict__ colidx, const int32_t* __restrict__ counts, int64_t D,
    int64_t V, int64_t vp, int64_t range_words, int ranges, int64_t doc_base, int64_t pos_base,
    uint8_t* __restrict__ hist, const uint16_t* __restrict__ off, const int64_t* __restrict__ colptr,
    int32_t* __restrict__ docidx, int32_t* __restrict__ ccounts, uint32_t* __restrict__ csr2csc)
{
    extern __shared__ __align__(16) uint32_t tile_mask[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int kWarps = kTileThreads / 32;
    const int64_t tile = blockIdx.x;
    const int64_t d_lo = tile * kTileDocs;
    const int nd = (int)((d_lo + kTileDocs < D) ? kTileDocs : D - d_lo);
    for (int rg = 0; rg < ranges; ++rg) {
        const int64_t lo = (int64_t)rg * range_words, hi = (lo + range_words < vp) ? lo + range_words : vp;
        const int nw = (int)(hi - lo);
        for (int i = threadIdx.x * 4; i < nw; i += kTileThreads * 4)
            *reinterpret_cast<uint4*>(tile_mask + i) = make_uint4(0u, 0u, 0u, 0u);
        __syncthreads();
        // which documents of the tile contain each word (an OR commutes: no ordering between lanes / warps needed)
        for (int dl = warp; dl < nd; dl += kWarps) {
            const int64_t b = rowptr[d_lo + dl] - pos_base, e = rowptr[d_lo + dl + 1] - pos_base;
            for (int64_t p = b + lane; p < e; p += 32) {
                const int64_t x = (int64_t)colidx[p] - lo;
                if (x >= 0 && x < nw) atomicOr(tile_mask + x, 1u << dl);
            }
        }
        __syncthreads();
        if (MODE == 0) {
            uint8_t* out = hist + tile * vp + lo;
            for (int i = threadIdx.x * 4; i < nw; i += kTileThreads * 4) {
                const uint4 m = *reinterpret_cast<const uint4*>(tile_mask + i);
                *reinterpret_cast<uint32_t*>(out + i) = (uint32_t)__popc(m.x) | ((uint32_t)__popc(m.y) << 8) |
                                                        ((uint32_t)__popc(m.z) << 16) | ((uint32_t)__popc(m.w) << 24);
            }
        } else {
            for (int dl = warp; dl < nd; dl += kWarps) {
                const int64_t b = rowptr[d_lo + dl] - pos_base, e = rowptr[d_lo + dl + 1] - pos_base;
                const uint32_t below = (1u << dl) - 1u;
                for (int64_t p = b + lane; p < e; p += 32) {
                    const int w = colidx[p];
                    const int64_t x = (int64_t)w - lo;
                    if (x < 0 || x >= nw) continue;
                    // documents before this one that contain w: inside the tile (mask), in earlier tiles (off), in the
                    // block's words before w (colptr)
                    const int64_t q = (colptr[w] - pos_base) + (int64_t)off[tile * vp + w] + __popc(tile_mask[x] & below);
                    docidx[q] = (int32_t)(doc_base + d_lo + dl);
                    if (ccounts) ccounts[q] = counts[p];
                    if (csr2csc) csr2csc[p] = (uint32_t)(pos_base + q);
                }
            }
        }
        __syncthreads();
    }
}

// kernel B: per word, exclusive prefix of its tile counts (uint16) and its total
__global__ void __launch_bounds__(256) csc_tile_scan_kernel(const uint8_t* __restrict__ hist, int64_t tiles, int64_t V,
                                                            int64_t vp, uint16_t* __restrict__ off,
                                                            uint32_t* __restrict__ total)
{
    const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= V) return;
    uint32_t run = 0;
    int64_t t = 0;
    for (; t + 8 <= tiles; t += 8) {
        uint8_t c[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) c[u] = hist[(t + u) * vp + w];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            off[(t + u) * vp + w] = (uint16_t)run;
            run += c[u];
        }
    }
    for (; t < tiles; ++t) {
        off[t * vp + w] = (uint16_t)run;
        run += hist[t * vp + w];
    }
    total[w] = run;
}

// colptr[w] = pos_base + sum_{w' < w} total[w'], colptr[V] = pos_base + nnz; one CTA
__global__ void __launch_bounds__(1024) csc_colptr_kernel(const uint32_t* __restrict__ total, int64_t V, int64_t pos_base,
                                                          int64_t* __restrict__ colptr)
{
    __shared__ int64_t part[1024];
    const int tid = threadIdx.x;
    const int64_t per = (V + 1023) / 1024, b = (int64_t)tid * per, e = (b + per < V) ? b + per : V;
    int64_t s = 0;
    for (int64_t w = b; w < e; ++w) s += total[w];
    part[tid] = s;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {  // inclusive Hillis-Steele scan of the 1024 partials
        const int64_t v = (tid >= o) ? part[tid - o] : 0;
        __syncthreads();
        part[tid] += v;
        __syncthreads();
    }
    int64_t run = pos_base + part[tid] - s;
    for (int64_t w = b; w < e; ++w) {
        colptr[w] = run;
        run += total[w];
    }
    if (tid == 1023) colptr[V] = pos_base + part[1023];
}

static int csr_to_csc_tiles(const int64_t* rowptr, const int32_t* colidx, const int32_t* counts, int64_t D, int64_t V,
                            int64_t nnz, int64_t doc_base, int64_t pos_base, int64_t* colptr, int32_t* docidx,
                            int32_t* ccounts, uint32_t* csr2csc, void* workspace, int64_t workspace_bytes, cudaStream_t s)
{
    const TilePlan t = tile_plan(D, V);
    if ((int64_t)t.total > workspace_bytes) {
        set_error("csr_to_csc: workspace %lld < %lld", (long long)workspace_bytes, (long long)t.total);
        return TTLDA_EWORKSPACE;
    }
    uint8_t* hist = (uint8_t*)workspace;
    uint16_t* off = (uint16_t*)((char*)workspace + t.hist_b);
    uint32_t* total = (uint32_t*)((char*)workspace + t.hist_b + t.off_b);
    const size_t smem = (size_t)t.range_words * 4;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(csc_tile_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(csc_tile_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(csc_tile)");
        configured = true;
    }
    if (t.tiles > 0 && nnz > 0) {
        const unsigned grid = (unsigned)t.tiles;
        csc_tile_kernel<0><<<grid, kTileThreads, smem, s>>>(rowptr, colidx, counts, D, V, t.vp, t.range_words, t.ranges,
                                                               doc_base, pos_base, hist, nullptr, nullptr, nullptr, nullptr,
                                                               nullptr);
        TTLDA_LAUNCH_CHECK("csc_tile_hist");
        csc_tile_scan_kernel<<<(unsigned)((V + 255) / 256), 256, 0, s>>>(hist, t.tiles, V, t.vp, off, total);
        TTLDA_LAUNCH_CHECK("csc_tile_scan");
    } else {
        cudaError_t e = cudaMemsetAsync(total, 0, (size_t)V * 4, s);
        if (e != cudaSuccess) return cuda_fail(e, "memset totals");
    }
    csc_colptr_kernel<<<1, 1024, 0, s>>>(total, V, pos_base, colptr);
    TTLDA_LAUNCH_CHECK("csc_colptr");
    if (t.tiles > 0 && nnz > 0) {
        const unsigned grid = (unsigned)t.tiles;
        csc_tile_kernel<1><<<grid, kTileThreads, smem, s>>>(rowptr, colidx, counts, D, V, t.vp, t.range_words, t.ranges,
                                                               doc_base, pos_base, nullptr, off, colptr, docidx, ccounts,
                                                               csr2csc);
        TTLDA_LAUNCH_CHECK("csc_tile_place");
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
    if (tile_path_applies(D, V, nblocks)) {  // whichever path runs: the larger of the two layouts
        const TilePlan t = tile_plan(D, V);
        if (t.total > need) need = t.total;
    }
    return (int64_t)need;
}

static int csr_to_csc_impl(const int64_t* rowptr, const int32_t* colidx, const int32_t* counts, int64_t D, int64_t V,
                           int64_t nnz, int64_t nblocks, int64_t block_docs, int64_t doc_base, int64_t pos_base,
                           int64_t* colptr, int32_t* docidx, int32_t* ccounts, uint32_t* csr2csc, void* workspace,
                           int64_t workspace_bytes, cudaStream_t s)
{
    if (tile_path_applies(D, V, nblocks))
        return csr_to_csc_tiles(rowptr, colidx, counts, D, V, nnz, doc_base, pos_base, colptr, docidx, ccounts, csr2csc,
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
