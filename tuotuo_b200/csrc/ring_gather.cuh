// ring_gather.cuh — K-vector rows staged through a per-lane shared-memory ring by `cp.async` (SASS LDGSTS.128): the bytes
// in flight live in shared memory instead of in destination registers.
//
// Why: the register-staged walk of traverse.cuh keeps ONE sub-batch (R rows) per warp in flight — its destination
// registers are the second of two row buffers — and at 166 registers only 12 warps fit an SM: ~43 KB in flight per SM.
// The E-step at cfg3 then runs at the pace (L + C) / 2 per sub-batch (L: loaded L2 latency, C: the dependent
// dot -> reduce -> reciprocal -> axpy chain), i.e. it waits for its own loads a third of the time (ncu: long-scoreboard
// stalls).  Here every lane copies ITS OWN 16-byte slices of its group's row into a private slot of a ring of NSTAGE
// sub-batches and reads them back with LDS.128 when their turn comes: no lane ever reads bytes another lane copied, so the
// only synchronisation is the lane's own `cp.async.wait_group`; the register file holds one row buffer instead of two
// (=> 16 warps/SM), and (NSTAGE - 1) x 16 sub-batches are in flight per SM.  Price: every row crosses the shared-memory
// pipe twice (fill + LDS).
//
// Same contract as tma_gather.cuh's walk (the TMA engine's gather4 does the same job with one instruction per sub-batch,
// but its issue rate, not the memory system, bounds it: 13.3 TB/s in tools/membench4.cu).
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

}  // namespace ttlda
