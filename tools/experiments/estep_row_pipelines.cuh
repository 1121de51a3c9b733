// tools/experiments/estep_row_pipelines.cuh — NOT compiled into libttlda: two round-2 E-step experiments kept for the record.
//
// Both were built into csrc/estep.cu behind environment switches, passed the parity tests (bit-identical to the default
// register walk: same arithmetic in the same order) and were measured same-box at cfg3 (1M documents, K = 100, V = 50k):
//
//   default walk (two register row buffers, 12 warps/SM at 166 registers)                      E-step 14.4 ms
//   TTLDA_RING=1   rows staged through a per-lane cp.async ring in shared memory (3-4 stages,
//                  16 warps/SM at 128 registers)                                                19.4 ms (cfg4 27.1 vs 26.9,
//                                                                                               cfg5 shard 153.5 vs 152.3)
//   TTLDA_SKEW=3   three register buffers, two sub-batches in flight, 10 one-warp CTAs/SM       24.5 ms
//   TTLDA_SKEW=4   four buffers, two in flight + dot/reduce of q+1 issued next to rcp/axpy of q   21.9 ms
//   (and, with the default walk: 13 / 14 / 15 / 16 warps per SM by capping registers at 152 / 144 / 136 / 128 — small
//   spills — 17.4 / 19.9 / 22.7 / 24.6 ms; one-warp CTAs at 12 warps: 14.38 ms, i.e. free)
//   profiles/r02m_bench_*_{base,ring}.json, profiles/r02v_*, profiles/r02w_*, DESIGN.md section 5.
//
// Why they were tried: the source-level profile (profiles/r02l_estep_source_hotloop_cfg3.txt) puts 24 % of the E-step's
// stall samples on the first use of the prefetched row buffer — one sub-batch per warp in flight is not enough to cover the
// ~1200-cycle loaded L2 latency.  Why they lost: per warp and sub-batch every variant needs MORE time than the default walk
// (985 cycles: ring 1770, SKEW=3 1400, SKEW=4 1120) — the generic per-sub-batch index routing of these walks (two shuffles
// + select per issue, 64-bit sub-batch bookkeeping: ~130 instructions per step against ~95) and, for the ring, the
// LDGSTS + LDS round trip cost more than the extra bytes in flight return; with fewer or register-starved warps the
// dependent chain, not the memory latency, sets the pace.  A fair re-test needs the third buffer inside the hand-tuned
// batch-structured walk (walk_rows_carried), with the buffer roles rotating across 32-row batches.
//
// To resurrect: include this file from csrc/estep.cu after tma_gather.cuh and re-add the dispatch at the top of
// launch_doc_pass (see the `dispatch` comment at the end).
#pragma once



#include "traverse.cuh"

namespace ttlda {

inline bool ring_gather_applicable(int G, int E) { return G >= 8 && E >= 5 && E <= 16; }
inline bool ring_gather_enabled()
{
    const char* e = getenv("TTLDA_RING");
    return e && e[0] == '1';
}

// sub-batches per warp in the ring (NSTAGE - 1 in flight while one is consumed), by shape: what 16 / 12 / 8 resident warps
// (128 / 168 / 255 registers) can hold in ~200 KB of shared memory — a stage is 32/G rows of up to 2 G E doubles
// (the TTLDA_RING_* macros exist for same-box A/B builds: tools/build_variant.sh)
#ifndef TTLDA_RING_STAGES_SMALL
#define TTLDA_RING_STAGES_SMALL 4
#endif
#ifndef TTLDA_RING_STAGES_LARGE
#define TTLDA_RING_STAGES_LARGE 3
#endif
#ifndef TTLDA_RING_CTAS_SMALL
#define TTLDA_RING_CTAS_SMALL 4
#endif
__host__ __device__ constexpr int ring_stages(int G, int E)
{
    return (G == 8 && E <= 7) ? TTLDA_RING_STAGES_SMALL : TTLDA_RING_STAGES_LARGE;
}
__host__ __device__ constexpr int ring_min_ctas(int E)  // CTAs of 4 warps each
{
    return E <= 7 ? TTLDA_RING_CTAS_SMALL : (E <= 12 ? 3 : 2);
}

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src)
{
#ifdef TTLDA_RING_CA
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
#else
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
#endif
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait()
{
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

template <int G, int E>
struct RingPipe {
    static constexpr int R = 32 / G, kRingStages = ring_stages(G, E);
    // a stage holds the R rows of a sub-batch as they lie in the table (ld * 8 bytes each); lane (g, r) owns the 16-byte
    // slices 16 (g + G j) of row r — a quarter-warp (8 consecutive lanes) reads 128 contiguous bytes: conflict-free LDS.128
    uint32_t base_u32;          // this warp's ring (shared-space address), the lane's offset folded in
    const unsigned char* base;  // generic address of the same byte
    uint32_t stage_bytes;

    static __host__ __device__ size_t bytes_per_warp(int64_t ld) { return (size_t)kRingStages * R * ld * 8; }

    __device__ __forceinline__ void init(unsigned char* warp_base, int ld, int g, int r)
    {
        stage_bytes = (uint32_t)(R * ld * 8);
        base = warp_base + r * ld * 8 + g * 16;
        base_u32 = (uint32_t)__cvta_generic_to_shared(base);
    }
    // copy this lane's slices of row `rowg` (lane offset already folded in) into stage q % NSTAGE; always commits a group
    // (an empty one when `go` is false) so that the number of groups outstanding is the same in every iteration
    __device__ __forceinline__ void produce(uint32_t q, const double* rowg, bool go, bool last_ok) const
    {
        if (go) {
            const uint32_t dst = base_u32 + (q % kRingStages) * stage_bytes;
#pragma unroll
            for (int j = 0; j < E - 1; ++j) cp_async16(dst + j * (16 * G), rowg + 2 * G * j);
            if (last_ok) cp_async16(dst + (E - 1) * (16 * G), rowg + 2 * G * (E - 1));
        }
        cp_async_commit();
    }
    __device__ __forceinline__ void read(KSlice<G, E>& out, uint32_t q, bool last_ok) const
    {
        const unsigned char* src = base + (q % kRingStages) * stage_bytes;
#pragma unroll
        for (int j = 0; j < E - 1; ++j) out.v[j] = *reinterpret_cast<const double2*>(src + j * (16 * G));
        out.v[E - 1] = last_ok ? *reinterpret_cast<const double2*>(src + (E - 1) * (16 * G)) : make_double2(0., 0.);
    }
};

// ring bytes per warp; never less than the per-document epilogue's transpose buffer (R group partials of 2 G E doubles),
// which aliases the drained ring
template <int G, int E>
__host__ __device__ inline size_t ring_bytes_per_warp(int64_t ld)
{
    const size_t ring = RingPipe<G, E>::bytes_per_warp(ld), wbuf = (size_t)(32 / G) * 2 * G * E * sizeof(double);
    return ((ring > wbuf ? ring : wbuf) + 127) / 128 * 128;
}

// Walk the rows of the n items [0, n) whose row indices are streamed from idx_ptr (32 per coalesced load, one batch
// ahead); sub-batch q gives group r the item R q + r.  The caller keeps its per-item payload in registers:
//   load_next(base)   the walk is fetching the indices of the batch starting at item `base` (one batch ahead)
//   rotate()          the batch fetched last becomes the current one
//   on_sub(sb, row)   sub-batch sb of the current batch: this lane's slices of this group's row
//   on_batch_end(cnt) the current batch (cnt items) has been consumed
template <int G, int E, class LoadNext, class Rotate, class OnSub, class OnBatchEnd>
__device__ __forceinline__ void ring_walk(const RingPipe<G, E>& pipe, const double* __restrict__ table, int ld32,
                                          const int32_t* __restrict__ idx_ptr, int64_t n, int lane, int g, int r,
                                          bool last_ok, uint64_t pol, LoadNext&& load_next, Rotate&& rotate,
                                          OnSub&& on_sub, OnBatchEnd&& on_batch_end)
{
    constexpr int R = 32 / G, SPB = 32 / R, AHEAD = ring_stages(G, E) - 1;
    static_assert(AHEAD < SPB, "the look-ahead must stay inside the next index batch");
    if (n <= 0) return;
    const int64_t Q = (n + R - 1) / R;
    const double* tg = table + 2 * g;

    int w_cur = 0, w_nxt = 0;  // lanes past the end keep row 0: a harmless in-bounds gather
    if (lane < n) w_cur = ld_stream_s32(idx_ptr + lane, pol);
    load_next((int64_t)0);
    rotate();
    if (32 + lane < n) w_nxt = ld_stream_s32(idx_ptr + 32 + lane, pol);
    load_next((int64_t)32);

    auto produce = [&](int64_t qp, int64_t q) {  // qp: sub-batch to fetch, q: sub-batch being consumed (qp - q <= AHEAD)
        const int src = (int)(qp % SPB) * R + r;
        const int w = (qp / SPB != q / SPB) ? __shfl_sync(0xffffffffu, w_nxt, src) : __shfl_sync(0xffffffffu, w_cur, src);
        pipe.produce((uint32_t)qp, row_at(tg, w, ld32), qp < Q, last_ok);
    };
#pragma unroll
    for (int qp = 0; qp < AHEAD; ++qp) produce(qp, 0);
    for (int64_t q = 0; q < Q; ++q) {
        const int sb = (int)(q % SPB);
        produce(q + AHEAD, q);  // into the stage consumed at q - 1 (its LDS results were used before this point)
        cp_async_wait<AHEAD>();  // all but the AHEAD newest groups have landed: sub-batch q is in the ring
        KSlice<G, E> row;
        pipe.read(row, (uint32_t)q, last_ok);
        on_sub(sb, row);
        if (sb == SPB - 1 || q == Q - 1) {
            const int64_t base = (q / SPB) * 32;
            on_batch_end((int)((n - base < 32) ? (n - base) : 32));
            w_cur = w_nxt;
            rotate();
            w_nxt = 0;
            if (base + 64 + lane < n) w_nxt = ld_stream_s32(idx_ptr + base + 64 + lane, pol);
            load_next(base + 64);
        }
    }
    cp_async_wait<0>();
}

// the shapes the cp.async ring feeds (ring_gather.cuh): every ld in (64, 1024]
#define TTLDA_FOR_EACH_RING_SHAPE(X)                                                                 \
    X(8, 5) X(8, 6) X(8, 7) X(8, 8)                                                                   \
    X(16, 5) X(16, 6) X(16, 7) X(16, 8)                                                               \
    X(32, 5) X(32, 6) X(32, 7) X(32, 8) X(32, 9) X(32, 10) X(32, 11) X(32, 12) X(32, 13) X(32, 14) X(32, 15) X(32, 16)


// Deeper register pipelines for the document-major pass (opt-in, TTLDA_SKEW=3|4; doc_pass_skew_kernel in estep.cu).
// The default walk holds two row buffers: while sub-batch s is consumed, s + 1 is in flight — ONE sub-batch per warp, and
// the source-level profile shows the warp waiting for it a quarter of the time (pace (L + C) / 2: L the loaded latency,
// C the dependent dot -> reduce -> reciprocal -> axpy chain).
//   NB = 3  three buffers, TWO sub-batches in flight (q + 1, q + 2 while q is consumed): pace max(C, (L + C) / 3)
//   NB = 4  four buffers, two in flight AND the chain skewed: the dot / reduce half of sub-batch q + 1 and the
//           reciprocal / axpy half of sub-batch q are issued in ONE basic block, so the in-order warp overlaps them:
//           pace max(C', (L + C') / 3) with C' ~ the longer half
// Price: 28 registers per extra buffer at E = 7 => one-warp CTAs, 10 (NB = 3) or 9 (NB = 4) per SM instead of 12 warps.
//   first(row) -> p      the dot / reduce half of a sub-batch
//   second(sb, row, p)   the rest of it
//   load_next(base), rotate(), on_batch_end(cnt): the caller's per-item payload, as in the ring / TMA walks
template <int G, int E, int NB, class LoadNext, class Rotate, class First, class Second, class OnBatchEnd>
__device__ __forceinline__ void deep_walk(const double* __restrict__ table, int ld32, const int32_t* __restrict__ idx_ptr,
                                          int64_t n, int lane, int g, int r, bool last_ok, uint64_t pol,
                                          LoadNext&& load_next, Rotate&& rotate, First&& first, Second&& second,
                                          OnBatchEnd&& on_batch_end)
{
    static_assert(NB == 3 || NB == 4, "");
    constexpr int R = 32 / G, SPB = 32 / R, AHEAD = NB - 1;
    static_assert(SPB > AHEAD, "the look-ahead must stay inside the next index batch");
    if (n <= 0) return;
    const int64_t Q = (n + R - 1) / R;
    const double* tg = table + 2 * g;

    int w_cur = 0, w_nxt = 0;  // lanes past the end keep row 0: a harmless in-bounds gather
    if (lane < n) w_cur = ld_stream_s32(idx_ptr + lane, pol);
    load_next((int64_t)0);
    rotate();
    if (32 + lane < n) w_nxt = ld_stream_s32(idx_ptr + 32 + lane, pol);
    load_next((int64_t)32);

    auto issue = [&](KSlice<G, E>& buf, int64_t qp, int64_t q) {  // rows of sub-batch qp, while q is the current one
        if (qp < Q) {
            const int src = (int)(qp % SPB) * R + r;
            const int w = (qp / SPB != q / SPB) ? __shfl_sync(0xffffffffu, w_nxt, src) : __shfl_sync(0xffffffffu, w_cur, src);
            buf.load_at(row_at(tg, w, ld32), last_ok);
        }
    };
    auto end_of = [&](int64_t q) {
        const int sb = (int)(q % SPB);
        if (sb == SPB - 1 || q == Q - 1) {
            const int64_t base = (q / SPB) * 32;
            on_batch_end((int)((n - base < 32) ? (n - base) : 32));
            w_cur = w_nxt;
            rotate();
            w_nxt = 0;
            if (base + 64 + lane < n) w_nxt = ld_stream_s32(idx_ptr + base + 64 + lane, pol);
            load_next(base + 64);
        }
    };
    int64_t q = 0;
    if constexpr (NB == 3) {
        KSlice<G, E> R0, R1, R2;
        issue(R0, 0, 0);
        issue(R1, 1, 0);
#define TTLDA_STEP(CUR, TGT)                      \
    {                                             \
        issue(TGT, q + 2, q);                     \
        const double p = first(CUR);              \
        second((int)(q % SPB), CUR, p);           \
        end_of(q);                                \
        if (++q >= Q) break;                      \
    }
        while (true) {
            TTLDA_STEP(R0, R2)
            TTLDA_STEP(R1, R0)
            TTLDA_STEP(R2, R1)
        }
#undef TTLDA_STEP
    } else {
        KSlice<G, E> R0, R1, R2, R3;
        R1.zero();  // first() also runs on the buffer AFTER the last sub-batch (result unused): keep it finite
        R2.zero();
        R3.zero();
        issue(R0, 0, 0);
        issue(R1, 1, 0);
        issue(R2, 2, 0);
        double p = first(R0);
#define TTLDA_STEP(CUR, NXT, TGT)                 \
    {                                             \
        issue(TGT, q + 3, q);                     \
        const double pn = first(NXT);             \
        second((int)(q % SPB), CUR, p);           \
        p = pn;                                   \
        end_of(q);                                \
        if (++q >= Q) break;                      \
    }
        while (true) {
            TTLDA_STEP(R0, R1, R3)
            TTLDA_STEP(R1, R2, R0)
            TTLDA_STEP(R2, R3, R1)
            TTLDA_STEP(R3, R0, R2)
        }
#undef TTLDA_STEP
    }
}


// The same pass with the rows staged through a per-lane shared-memory ring by cp.async (ring_gather.cuh): kRingStages - 1
// sub-batches in flight per warp without holding destination registers, one row buffer instead of two => 16 warps/SM.
template <int G, int E, int MODE>
__global__ void __launch_bounds__(kDocThreads, ring_min_ctas(E)) doc_pass_ring_kernel(const EStepParams P)
{
    constexpr bool ESTEP = MODE > 0, REFRESH = MODE == 1;
    constexpr int R = 32 / G;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* red = reinterpret_cast<double*>(smem_raw);  // 3*32 doubles for the CTA reduction (768 B)
    unsigned char* wbase = smem_raw + 768 + (size_t)(threadIdx.x >> 5) * ring_bytes_per_warp<G, E>(P.ld);

    const int lane = threadIdx.x & 31;
    const int g = lane % G, r = lane / G;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int ld = (int)P.ld, K = P.K;
    const bool last_ok = KSlice<G, E>::last_slice_ok(g, ld);
    const uint64_t pol = l2_evict_first_policy();

    RingPipe<G, E> pipe;
    pipe.init(wbase, ld, g, r);
    double* wbuf = reinterpret_cast<double*>(wbase);  // the epilogue's transpose buffer aliases the drained ring

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
        ring_walk<G, E>(
            pipe, P.eB, ld, P.colidx + b, n, lane, g, r, last_ok, pol,
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
                const double p = group_sum<G>(t.dot(row, ESTEP ? 1e-100 / G : 0.));  // lda_model.py:242
                if (ESTEP) acc.axpy(fast_div_pos((double)c, p), row);                // lda_model.py:248
                // hand p back to the lane that loaded nonzero `lane` of the batch (group lane % R, sub-batch lane / R)
                const double pv = __shfl_sync(0xffffffffu, p, (lane % R) * G);
                if (lane / R == sb) my_p = pv;
            },
            [&](int cnt) {
                if (lane < cnt) {  // one log per nonzero (of theta.beta itself: the token term has no 1e-100, :125-139)
                    tok_acc += (double)c_cur * log(ESTEP ? my_p - 1e-100 : my_p);
                    cnt_acc += (double)c_cur;
                    if (handoff) st_stream_f64(P.ratio_csc + (uint32_t)pos_cur, fast_div_pos((double)c_cur, my_p), pol);
                }
            });
        if (ESTEP) {
            __syncwarp();  // every lane is done with the ring before it becomes the transpose buffer
            doc_epilogue<G, E, MODE>(P, d, acc, wbuf, lane, g, r, pol, lga, lg_asum, theta_acc, passes, capped);
            __syncwarp();
        }
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


// The same pass on the deeper register pipelines of traverse.cuh deep_walk (NB = 3 or 4 row buffers): ONE-warp CTAs capped
// at 200 / 224 registers, so that ten / nine of them share an SM.  Same arithmetic, same order as doc_pass_kernel:
// bit-identical results.
template <int G, int E, int MODE, int NB>
__global__ void __maxnreg__(NB == 3 ? 200 : 224) doc_pass_skew_kernel(const EStepParams P)
{
    constexpr bool ESTEP = MODE > 0, REFRESH = MODE == 1;
    constexpr int R = 32 / G;
    constexpr int LDP = 2 * G * E;
    extern __shared__ double smem[];
    double* red = smem;        // 3*32 doubles for the CTA reduction
    double* wbuf = smem + 96;  // R * LDP doubles: the epilogue's transpose buffer (one warp per CTA)

    const int lane = threadIdx.x & 31;
    const int g = lane % G, r = lane / G;
    const int64_t warp0 = blockIdx.x;
    const int64_t nwarps = gridDim.x;
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
    const bool handoff = ESTEP && P.ratio_csc != nullptr;

    for (int64_t d = warp0; d < P.D; d += nwarps) {
        const int64_t b = P.rowptr[d], n = P.rowptr[d + 1] - b;
        KSlice<G, E> t, acc;
        t.load(P.eT_in + d * P.ld, g, last_ok);
        acc.zero();
        int c_cur = 0, c_nxt = 0, pos_cur = 0, pos_nxt = 0;  // per-item payload of the current / next 32-item batch
        double my_p = 1.;
        deep_walk<G, E, NB>(
            P.eB, ld, P.colidx + b, n, lane, g, r, last_ok, pol,
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
            [&](const KSlice<G, E>& row) { return group_sum<G>(t.dot(row, ESTEP ? 1e-100 / G : 0.)); },  // lda_model.py:242
            [&](int sb, const KSlice<G, E>& row, double p) {
                const int c = __shfl_sync(0xffffffffu, c_cur, sb * R + r);  // 0 past the end: the group idles
                if (ESTEP) acc.axpy(fast_div_pos((double)c, p), row);       // lda_model.py:248
                // hand p back to the lane that loaded nonzero `lane` of the batch (group lane % R, sub-batch lane / R)
                const double pv = __shfl_sync(0xffffffffu, p, (lane % R) * G);
                if (lane / R == sb) my_p = pv;
            },
            [&](int cnt) {
                if (lane < cnt) {  // one log per nonzero (of theta.beta itself: the token term has no 1e-100, :125-139)
                    tok_acc += (double)c_cur * log(ESTEP ? my_p - 1e-100 : my_p);
                    cnt_acc += (double)c_cur;
                    if (handoff) st_stream_f64(P.ratio_csc + (uint32_t)pos_cur, fast_div_pos((double)c_cur, my_p), pol);
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


template <int G, int E, int MODE>
static int launch_ring(const EStepParams& P, int* grid_out, cudaStream_t s)
{
    constexpr int R = 32 / G;
    size_t per_warp = ring_bytes_per_warp<G, E>(P.ld);
    const size_t smem = 768 + (kDocThreads / 32) * per_warp;
    cudaError_t e = cudaFuncSetAttribute(doc_pass_ring_kernel<G, E, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(doc_pass_ring)");
    (void)R;
    const int grid = resident_grid(doc_pass_ring_kernel<G, E, MODE>, smem, P.D);
    doc_pass_ring_kernel<G, E, MODE><<<grid, kDocThreads, smem, s>>>(P);
    *grid_out = grid;
    return TTLDA_OK;
}


inline int skew_buffers()  // TTLDA_SKEW=3 | 4: row buffers of the deep register pipeline (0: the default walk)
{
    const char* e = getenv("TTLDA_SKEW");
    return (e && (e[0] == '3' || e[0] == '4')) ? e[0] - '0' : 0;
}

template <int G, int E, int MODE, int NB>
static int launch_skew(const EStepParams& P, int* grid_out, cudaStream_t s)
{
    constexpr int R = 32 / G;
    constexpr size_t smem = (96 + R * 2 * G * E) * sizeof(double);
    static int per_sm = 0;
    if (per_sm < 1) {
        int n = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, doc_pass_skew_kernel<G, E, MODE, NB>, 32, smem) !=
                cudaSuccess ||
            n < 1)
            n = 1;
        per_sm = n;
    }
    const int grid = grid_for_warps(P.D, 1, per_sm);
    doc_pass_skew_kernel<G, E, MODE, NB><<<grid, 32, smem, s>>>(P);
    *grid_out = grid;
    return TTLDA_OK;
}


}  // namespace ttlda

/* dispatch (inside launch_doc_pass, before the TMA branch):
    if (const int nb = skew_buffers(); gs.G == 8 && nb) {  // deep register pipeline, one-warp CTAs
#define TTLDA_CASE(EE)                                                         \
    if (gs.E == EE) return nb == 3 ? launch_skew<8, EE, MODE, 3>(P, grid, s) : launch_skew<8, EE, MODE, 4>(P, grid, s);
        TTLDA_CASE(5) TTLDA_CASE(6) TTLDA_CASE(7) TTLDA_CASE(8)
#undef TTLDA_CASE
    }
    if (ring_gather_applicable(gs.G, gs.E) && ring_gather_enabled()) {  // rows staged through the cp.async ring
#define TTLDA_CASE(GG, EE) \
    if (gs.G == GG && gs.E == EE) return launch_ring<GG, EE, MODE>(P, grid, s);
        TTLDA_FOR_EACH_RING_SHAPE(TTLDA_CASE)
#undef TTLDA_CASE
    }
*/
