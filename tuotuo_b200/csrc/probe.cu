// probe.cu — the gather ceiling of THIS box, measured with the product's own traversal.
//
// The E-step (estep.cu) and the statistics pass (sstats.cu) are gathers of 8K-byte rows of an fp64 table, one row per
// nonzero of the doc-term matrix (tuotuo/lda_model.py:242,248,262 — the np.dot / np.outer over the columns
// `ids = np.nonzero(X[d])`).  Whether such a gather is bounded by HBM or by the L2 -> SM path depends on whether the
// table (and the window of it the index stream touches) fits the 126 MB L2, so the roofline `bench.py` reports is not
// a constant: it is measured in the same process, on the same table and the same index stream, by this kernel — the
// product traversal (traverse.cuh: G lanes per row, R = 32/G rows per step, 32 indices per coalesced batch, register
// double-buffering) with every arithmetic instruction removed.  What is left is the row loads and one add per loaded
// value (so that the loads cannot be eliminated).  Test/bench infrastructure of the library, not part of the EM path.
#include "traverse.cuh"

namespace ttlda {

constexpr int kProbeThreads = 128;
constexpr int kProbeChunk = 1024;  // indices per warp visit (as the statistics pass: 1024-position chunks)

template <int G, int E>
__global__ void __launch_bounds__(kProbeThreads, (E <= 8 ? 4 : 1))
    probe_gather_kernel(const double* __restrict__ table, int ld, const int32_t* __restrict__ idx, int64_t n,
                        double* __restrict__ sink)
{
    const int lane = threadIdx.x & 31;
    const int g = lane % G, r = lane / G;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const bool last_ok = KSlice<G, E>::last_slice_ok(g, ld);
    const uint64_t pol = l2_evict_first_policy();
    double s0 = 0., s1 = 0.;
    for (int64_t c0 = warp0 * kProbeChunk; c0 < n; c0 += nwarps * kProbeChunk) {
        const int64_t c1 = (c0 + kProbeChunk < n) ? c0 + kProbeChunk : n;
        int nx = (c0 + lane < c1) ? ld_stream_s32(idx + c0 + lane, pol) : 0;
        for (int64_t n0 = c0; n0 < c1; n0 += 32) {
            const int cnt = (int)((c1 - n0 < 32) ? (c1 - n0) : 32);
            const int my = nx;
            nx = (n0 + 32 + lane < c1) ? ld_stream_s32(idx + n0 + 32 + lane, pol) : 0;
            walk_rows<G, E, (E <= 8)>(table, ld, my, cnt, g, r, last_ok, [&](int, const KSlice<G, E>& row) {
#pragma unroll
                for (int j = 0; j < E; ++j) {
                    s0 += row.v[j].x;
                    s1 += row.v[j].y;
                }
            });
        }
    }
    // one conditional store per thread keeps the loads alive without adding traffic
    const double s = s0 + s1;
    if (s == 1.2345678912345e-300) sink[0] = s;
}

template <int G, int E>
static int launch_probe(const double* table, int64_t ld, const int32_t* idx, int64_t n, double* sink, cudaStream_t s)
{
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, probe_gather_kernel<G, E>, kProbeThreads, 0) !=
            cudaSuccess || per_sm < 1)
        per_sm = 1;
    const int64_t chunks = (n + kProbeChunk - 1) / kProbeChunk;
    const int grid = grid_for_warps(chunks, kProbeThreads / 32, per_sm);
    probe_gather_kernel<G, E><<<grid, kProbeThreads, 0, s>>>(table, (int)ld, idx, n, sink);
    return TTLDA_OK;
}

}  // namespace ttlda

using namespace ttlda;

extern "C" int ttlda_probe_gather_f64(const double* table, int64_t rows, int64_t ld, const int32_t* idx, int64_t n,
                                      double* sink, void* stream)
{
    TTLDA_REQUIRE(table && sink && (n == 0 || idx), "NULL pointer");
    TTLDA_REQUIRE(rows >= 1 && n >= 0, "negative size");
    TTLDA_REQUIRE(ld >= 4 && ld % 4 == 0 && ld <= TTLDA_MAX_K, "ld must be a multiple of 4 in [4, 1024]");
    TTLDA_REQUIRE(rows * ld < ((int64_t)1 << 40), "table too large");
    if (n == 0) return TTLDA_OK;
    const GroupShape gs = group_shape_for(ld);
#define TTLDA_CASE(GG, EE) \
    if (gs.G == GG && gs.E == EE) { \
        int rc = launch_probe<GG, EE>(table, ld, idx, n, sink, (cudaStream_t)stream); \
        if (rc != TTLDA_OK) return rc; \
        TTLDA_LAUNCH_CHECK("probe_gather"); \
        return TTLDA_OK; \
    }
    TTLDA_FOR_EACH_SHAPE(TTLDA_CASE)
#undef TTLDA_CASE
    set_error("no probe instantiated for ld=%lld", (long long)ld);
    return TTLDA_EINVAL;
}
