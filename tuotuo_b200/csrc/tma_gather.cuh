// tma_gather.cuh — K-vector rows gathered by the TMA engine: ONE `cp.async.bulk.tensor.2d ... tile::gather4`
// (SASS UTMALDG.2D.GATHER4, new on sm_100) fetches the 4 rows of a sub-batch from a [rows][ld] fp64 table straight
// into shared memory and completes on an mbarrier, NSTAGE sub-batches ahead of the arithmetic.
//
// Why (measured, tools/membench4.cu vs membench2.cu, 800-byte rows from a 40 MB table, same dot/reduce/divide/axpy
// chain): register-staged LDG 10.3-11.9 TB/s, gather4 13.3 TB/s; the bytes in flight live in shared memory
// (NSTAGE x 4 rows per warp), the register file only holds the slice being consumed (=> 16 warps/SM instead of 12), and
// the LSU issues E LDS.128 per lane instead of E LDG.128.  (One bulk copy PER ROW was tried first and lost: the copy
// engine's per-operation rate became the bound — tools/experiments/rowpipe_bulk_async.cuh.)
//
// Applies to tables whose rows are handled by groups of G = 8 lanes (R = 4 rows per sub-batch = one gather4) and fit
// one TMA box (ld <= 256 doubles); other shapes keep the LDG walk of traverse.cuh.
//
// Protocol: one warp is producer and consumer, so no "empty" barriers: lane 0 arms stage q % NSTAGE with
// expect_tx(4 * ld * 8) and issues the gather4; every lane waits on the stage's phase ((q / NSTAGE) & 1) and reads its
// 16-byte slices of its group's row; a stage is re-armed only after a __syncwarp() that follows its last read.
#pragma once

#include <cuda.h>  // CUtensorMap (types only: the encoder is fetched through cudaGetDriverEntryPoint)

#include "traverse.cuh"

namespace ttlda {

// ---- host: tensor map of a [rows][ld] fp64 table with a {ld, 1} box (the gather4 box has 1 in the gathered dim) ----
int make_row_gather_map(const double* table, int64_t rows, int64_t ld, CUtensorMap* out);  // api.cu

inline bool tma_gather_applicable(int G, int64_t ld) { return G == 8 && ld <= 256; }
// OPT-IN (TTLDA_TMA=1): in the full kernels the TMA-fed variants are correct (tests/test_gpu_parity.py runs them) but
// not faster than the register-staged walk at cfg3 — E-step 20.7 vs 20.5 ms, statistics pass 15.6 vs 14.5 ms: the
// per-segment pipeline restarts and the mbarrier wait on the critical path eat the microbenchmark's advantage.
inline bool tma_gather_enabled()
{
    const char* e = getenv("TTLDA_TMA");
    return e && e[0] == '1';
}

// ---- device -----------------------------------------------------------------------------------------------------------
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
__device__ __forceinline__ void tma_gather4(uint32_t dst, const CUtensorMap* map, int r0, int r1, int r2, int r3,
                                            uint32_t bar)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes "
        "[%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::"r"(dst),
        "l"(map), "r"(0), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(bar)
        : "memory");
}

constexpr int kTmaStages = 4;  // sub-batches in flight per warp (4 x 4 rows x ld x 8 bytes of shared memory)

template <int E>
struct Gather4Pipe {
    static constexpr int G = 8, R = 4;
    const unsigned char* stages;  // this warp's stage buffers (generic address)
    uint32_t stages_u32, bars_u32;
    uint32_t row_bytes, stage_bytes;
    uint32_t gq;  // sub-batches consumed so far (stage / parity bookkeeping), warp-uniform

    static __host__ __device__ size_t bytes_per_warp(int64_t ld) { return (size_t)kTmaStages * R * ld * 8 + 128; }

    __device__ __forceinline__ void init(unsigned char* base, int ld, int lane)
    {
        stages = base;
        stages_u32 = smem_u32(base);
        row_bytes = (uint32_t)ld * 8u;
        stage_bytes = R * row_bytes;
        bars_u32 = stages_u32 + kTmaStages * stage_bytes;
        gq = 0;
        if (lane == 0)
            for (int s = 0; s < kTmaStages; ++s) mbar_init(bars_u32 + 8 * s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        __syncwarp();
    }
    __device__ __forceinline__ void produce(const CUtensorMap* map, uint32_t q, int r0, int r1, int r2, int r3,
                                            int lane) const
    {
        if (lane == 0) {
            const uint32_t st = q % kTmaStages;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // prior generic reads of the stage
            mbar_expect_tx(bars_u32 + 8 * st, stage_bytes);
            tma_gather4(stages_u32 + st * stage_bytes, map, r0, r1, r2, r3, bars_u32 + 8 * st);
        }
    }
    __device__ __forceinline__ void wait(uint32_t q) const
    {
        mbar_wait(bars_u32 + 8 * (q % kTmaStages), (q / kTmaStages) & 1u);
    }
    __device__ __forceinline__ void read(KSlice<G, E>& out, uint32_t q, int g, int r, bool last_ok) const
    {
        const double* row = reinterpret_cast<const double*>(stages + (q % kTmaStages) * stage_bytes + r * row_bytes);
#pragma unroll
        for (int j = 0; j < E - 1; ++j) out.v[j] = *reinterpret_cast<const double2*>(row + 2 * (g + G * j));
        out.v[E - 1] = last_ok ? *reinterpret_cast<const double2*>(row + 2 * (g + G * (E - 1))) : make_double2(0., 0.);
    }
};

// Walk the rows of the n items [0, n) whose row indices are streamed from idx_ptr (32 per coalesced load, one batch
// ahead); sub-batch q gives group r the item 4q + r.  The row pipeline runs continuously across 32-item batches.
// The caller keeps its own per-item payload (counts, ...) in registers and is told when to load / rotate it:
//   load_next(base)   the walk is fetching the indices of the batch starting at item `base` (one batch ahead)
//   rotate()          the batch fetched last becomes the current one
//   on_sub(sb, row)   sub-batch sb (0..7) of the current batch: this lane's slices of this group's row
//   on_batch_end(cnt) the current batch (cnt items) has been consumed
template <int E, class LoadNext, class Rotate, class OnSub, class OnBatchEnd>
__device__ __forceinline__ void tma_walk(Gather4Pipe<E>& pipe, const CUtensorMap* map,
                                         const int32_t* __restrict__ idx_ptr, int64_t n, int lane, int g, int r,
                                         bool last_ok, uint64_t pol, LoadNext&& load_next, Rotate&& rotate,
                                         OnSub&& on_sub, OnBatchEnd&& on_batch_end)
{
    constexpr int R = 4, SPB = 8, AHEAD = kTmaStages - 1;
    if (n <= 0) return;
    const int64_t Q = (n + R - 1) / R;
    const uint32_t gq0 = pipe.gq;

    int w_cur = 0, w_nxt = 0;  // lanes past the end keep row 0: a harmless in-bounds gather
    if (lane < n) w_cur = ld_stream_s32(idx_ptr + lane, pol);
    load_next((int64_t)0);
    rotate();
    if (32 + lane < n) w_nxt = ld_stream_s32(idx_ptr + 32 + lane, pol);
    load_next((int64_t)32);

    auto produce = [&](int64_t qp, int64_t q) {
        const int src = (int)(qp % SPB) * R;
        int r0, r1, r2, r3;
        if (qp / SPB != q / SPB) {  // warp-uniform: touch w_nxt (a load possibly still in flight) only when needed
            r0 = __shfl_sync(0xffffffffu, w_nxt, src);
            r1 = __shfl_sync(0xffffffffu, w_nxt, src + 1);
            r2 = __shfl_sync(0xffffffffu, w_nxt, src + 2);
            r3 = __shfl_sync(0xffffffffu, w_nxt, src + 3);
        } else {
            r0 = __shfl_sync(0xffffffffu, w_cur, src);
            r1 = __shfl_sync(0xffffffffu, w_cur, src + 1);
            r2 = __shfl_sync(0xffffffffu, w_cur, src + 2);
            r3 = __shfl_sync(0xffffffffu, w_cur, src + 3);
        }
        pipe.produce(map, gq0 + (uint32_t)qp, r0, r1, r2, r3, lane);
    };
    for (int64_t qp = 0; qp < AHEAD && qp < Q; ++qp) produce(qp, 0);
    for (int64_t q = 0; q < Q; ++q) {
        const int sb = (int)(q % SPB);
        if (q + AHEAD < Q) {
            __syncwarp();  // every lane has finished reading the stage being re-armed (consumed at q - 1)
            produce(q + AHEAD, q);
        }
        pipe.wait(gq0 + (uint32_t)q);
        KSlice<8, E> row;
        pipe.read(row, gq0 + (uint32_t)q, g, r, last_ok);
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
    __syncwarp();
    pipe.gq = gq0 + (uint32_t)Q;
}

}  // namespace ttlda
