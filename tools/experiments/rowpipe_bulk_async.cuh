// rowpipe.cuh — per-warp asynchronous gather pipeline: K-vector rows travel global -> shared memory as bulk async
// copies (cp.async.bulk, SASS UBLKCP — the TMA engine's non-tensor path) that complete on an mbarrier, NSTAGE
// sub-batches ahead of the arithmetic.
//
// Why: the passes are gathers of 8K-byte rows.  With LDG the rows in flight are bounded by registers (168 regs at
// 12 warps/SM left ncu showing long-scoreboard + wait stalls at 26-33 % issue utilisation, profiles/r01c); with bulk
// copies ONE lane issues ONE instruction per row, the data in flight lives in shared memory (NSTAGE x R rows per
// warp) and the register file only holds the slice being consumed.
//
// Protocol (single warp is producer and consumer, so no "empty" barriers are needed):
//   produce(q):  lanes 0..R-1 each issue the bulk copy of one row of sub-batch q into stage q % NSTAGE; lane 0
//                arms the stage's mbarrier with expect_tx = R * row_bytes.
//   consume(q):  every lane waits on the stage's mbarrier phase ((q / NSTAGE) & 1), then reads its 16-byte slices
//                of its group's row with LDS.128.
//   A stage is refilled only after __syncwarp() following its last read (WAR across the generic/async proxies).
#pragma once

#include "traverse.cuh"

namespace ttlda {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!ok);
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// Per-warp pipeline state.  Shared-memory layout per warp: NSTAGE stages x R row slots x row_doubles, then NSTAGE
// mbarriers (8 bytes each).
template <int G, int E>
struct RowPipe {
    static constexpr int R = 32 / G;
    double* rows;        // this warp's stage buffers
    uint32_t rows_u32;   // same, shared-window address
    uint32_t bars_u32;   // this warp's mbarriers
    int nstage;          // power of two
    int row_doubles;     // slot stride (= ld)
    uint32_t row_bytes;  // bytes copied per row (= ld * 8)
    uint32_t gq;         // sub-batches consumed so far (stage / parity bookkeeping), warp-uniform

    static __host__ __device__ size_t bytes_per_warp(int nstage, int64_t ld)
    {
        return ((size_t)nstage * R * ld * sizeof(double) + (size_t)nstage * 8 + 127) / 128 * 128;
    }

    __device__ __forceinline__ void init(void* base, int nstage_, int ld, int lane)
    {
        nstage = nstage_;
        row_doubles = ld;
        row_bytes = (uint32_t)ld * 8u;
        rows = reinterpret_cast<double*>(base);
        rows_u32 = smem_u32(base);
        bars_u32 = rows_u32 + (uint32_t)(nstage * R * ld * 8);
        gq = 0;
        if (lane == 0)
            for (int s = 0; s < nstage; ++s) mbar_init(bars_u32 + 8 * s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        __syncwarp();
    }

    // Issue the copies of local sub-batch `qp` (global number gq0 + qp).  `idx` = this lane's row index if lane < R.
    __device__ __forceinline__ void produce(const double* __restrict__ table, uint32_t q_global, int idx, int lane)
    {
        const uint32_t st = q_global & (uint32_t)(nstage - 1);
        const uint32_t bar = bars_u32 + 8 * st;
        if (lane < R) {
            if (lane == 0) {
                fence_proxy_async_smem();
                mbar_expect_tx(bar, (uint32_t)R * row_bytes);
            }
            bulk_g2s(rows_u32 + ((st * R + lane) * row_bytes), table + (int64_t)idx * row_doubles, row_bytes, bar);
        }
    }

    __device__ __forceinline__ void wait(uint32_t q_global) const
    {
        const uint32_t st = q_global & (uint32_t)(nstage - 1);
        mbar_wait(bars_u32 + 8 * st, (q_global / (uint32_t)nstage) & 1u);
    }

    // this lane's slices of its group's row in the stage of sub-batch q_global
    __device__ __forceinline__ void read(KSlice<G, E>& out, uint32_t q_global, int g, int r, bool last_ok) const
    {
        const uint32_t st = q_global & (uint32_t)(nstage - 1);
        const double* row = rows + (size_t)(st * R + r) * row_doubles;
#pragma unroll
        for (int j = 0; j < E - 1; ++j) out.v[j] = *reinterpret_cast<const double2*>(row + 2 * (g + G * j));
        out.v[E - 1] = last_ok ? *reinterpret_cast<const double2*>(row + 2 * (g + G * (E - 1))) : make_double2(0., 0.);
    }
};

// Walk the rows of the `n` items [0, n) whose row indices / counts are read 32 at a time from idx_ptr / cnt_ptr.
// Sub-batch q gives group r the item q*R + r.  `body(c, row, p_handback)` style work is supplied by two callbacks:
//   on_sub(sb, c, row)   per sub-batch: sb = sub-batch index inside the current 32-item batch, c = this group's
//                        count (0 past the end), row = this lane's slices of this group's row
//   on_batch(my_c, cnt)  when a 32-item batch has been consumed: my_c = the count lane `lane` loaded, cnt = items
//                        in the batch (lets the E-step take one log per nonzero)
template <int G, int E, class OnSub, class OnBatch>
__device__ __forceinline__ void pipe_walk(RowPipe<G, E>& pipe, const double* __restrict__ table,
                                          const int32_t* __restrict__ idx_ptr, const int32_t* __restrict__ cnt_ptr,
                                          int64_t n, int lane, int g, int r, bool last_ok, OnSub&& on_sub,
                                          OnBatch&& on_batch)
{
    constexpr int R = 32 / G;
    constexpr int SPB = 32 / R;  // sub-batches per 32-item batch (= G)
    if (n <= 0) return;
    const int64_t Q = (n + R - 1) / R;
    const uint32_t gq0 = pipe.gq;
    const int ahead = pipe.nstage - 1;

    int w_cur = 0, c_cur = 0, w_nxt = 0, c_nxt = 0;
    if (lane < n) {
        w_cur = ld_stream_s32(idx_ptr + lane);
        c_cur = ld_stream_s32(cnt_ptr + lane);
    }
    if (32 + lane < n) {
        w_nxt = ld_stream_s32(idx_ptr + 32 + lane);
        c_nxt = ld_stream_s32(cnt_ptr + 32 + lane);
    }
    // prologue: fill the pipeline (sub-batches 0 .. ahead-1 all lie in batch 0 or 1 because ahead <= SPB)
    for (int qp = 0; qp < ahead && qp < Q; ++qp) {
        const int src = (qp % SPB) * R + (lane % R);
        int idx;
        if (qp / SPB) idx = __shfl_sync(0xffffffffu, w_nxt, src);  // warp-uniform branch: do not touch w_nxt (a
        else idx = __shfl_sync(0xffffffffu, w_cur, src);           // load still in flight) unless it is needed
        pipe.produce(table, gq0 + qp, idx, lane);
    }
    for (int64_t q = 0; q < Q; ++q) {
        const int sb = (int)(q % SPB);
        const int64_t qp = q + ahead;
        if (qp < Q) {
            const int src = (int)(qp % SPB) * R + (lane % R);
            int idx;
            if (qp / SPB != q / SPB) idx = __shfl_sync(0xffffffffu, w_nxt, src);
            else idx = __shfl_sync(0xffffffffu, w_cur, src);
            __syncwarp();  // every lane has finished reading the stage being refilled (consumed at q-1)
            pipe.produce(table, gq0 + (uint32_t)qp, idx, lane);
        }
        const int c = __shfl_sync(0xffffffffu, c_cur, sb * R + r);
        pipe.wait(gq0 + (uint32_t)q);
        KSlice<G, E> row;
        pipe.read(row, gq0 + (uint32_t)q, g, r, last_ok);
        on_sub(sb, c, row);
        if (sb == SPB - 1 || q == Q - 1) {
            const int64_t base = (q / SPB) * 32;
            on_batch(c_cur, (int)((n - base < 32) ? (n - base) : 32));
            w_cur = w_nxt;
            c_cur = c_nxt;
            w_nxt = 0;
            c_nxt = 0;
            if (base + 64 + lane < n) {
                w_nxt = ld_stream_s32(idx_ptr + base + 64 + lane);
                c_nxt = ld_stream_s32(cnt_ptr + base + 64 + lane);
            }
        }
    }
    __syncwarp();
    pipe.gq = gq0 + (uint32_t)Q;
}

// pipeline depth for a row of `ld` doubles handled by groups of G lanes: as many stages (power of two, <= 8) as
// fit ~14 KB per warp, never running more than one 32-item batch ahead (nstage - 1 <= sub-batches per batch)
inline int pipe_stages_for(int G, int64_t ld)
{
    const int R = 32 / G;
    const int64_t stage_bytes = (int64_t)R * ld * 8;
    int ns = 8;
    while (ns > 2 && (ns * stage_bytes > 14336 || ns - 1 > G)) ns >>= 1;
    return ns;
}

}  // namespace ttlda
