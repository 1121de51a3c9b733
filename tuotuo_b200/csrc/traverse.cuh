// traverse.cuh — the group-mapped K-vector traversal shared by the document-major (E-step / ELBO) and word-major
// (sufficient statistics) passes.
//
// A K-vector (one row of a [.][ld] fp64 table) is owned by a GROUP of G lanes, G in {2,4,8,16,32}; a warp holds
// R = 32/G groups and therefore processes R nonzeros at once.  Lane g of a group owns the E double2 slices
//     k = 2*(g + G*j), j = 0..E-1                      (ld <= 2*G*E)
// so each load instruction reads G*16 contiguous bytes per group (coalesced, 16-byte aligned since ld % 4 == 0).
//
// Why groups instead of one warp per K-vector: per nonzero the fixed costs — the shuffle reduction of the dot
// product (log2 G steps instead of 5), the fp64 division (~25 instructions) and the index broadcasts — are paid
// once per R nonzeros, and each lane has E (not K/64) independent 16-byte loads in flight, E more when the next
// sub-batch is prefetched.  For K = 100: G = 8, E = 7, R = 4 => ~20 warp instructions per nonzero instead of ~130
// (ncu, profiles/r01a), which moves the passes from issue-bound to memory-bound.
#pragma once

#include "ttlda_common.cuh"

namespace ttlda {

template <int G, int E>
struct KSlice {
    double2 v[E];

    // Slices j < E-1 are always inside the row (E is minimal for ld); only the last one needs a predicate, which is
    // lane-constant (`last_ok`, see last_slice_ok) — so a load is E LDG.128 and nothing else.
    __device__ __forceinline__ void load(const double* __restrict__ row, int g, bool last_ok)
    {
#pragma unroll
        for (int j = 0; j < E - 1; ++j) v[j] = ldg2(row + 2 * (g + G * j));
        v[E - 1] = last_ok ? ldg2(row + 2 * (g + G * (E - 1))) : make_double2(0., 0.);
    }
    // the same with the lane's offset (2 g doubles) already folded into the pointer: offsets are compile-time constants
    __device__ __forceinline__ void load_at(const double* __restrict__ rowg, bool last_ok)
    {
#pragma unroll
        for (int j = 0; j < E - 1; ++j) v[j] = ldg2(rowg + 2 * G * j);
        v[E - 1] = last_ok ? ldg2(rowg + 2 * G * (E - 1)) : make_double2(0., 0.);
    }
    static __device__ __forceinline__ bool last_slice_ok(int g, int ld) { return 2 * (g + G * (E - 1)) < ld; }
    __device__ __forceinline__ void zero()
    {
#pragma unroll
        for (int j = 0; j < E; ++j) v[j] = make_double2(0., 0.);
    }
    // partial dot product over this lane's slices (+ init: lets the caller fold a constant into the sum for free)
    __device__ __forceinline__ double dot(const KSlice& o, double init = 0.) const
    {
        double p0 = init, p1 = 0., p2 = 0., p3 = 0.;  // four independent DFMA chains
#pragma unroll
        for (int j = 0; j < E; ++j) {
            if (j & 1) {
                p2 = fma(v[j].x, o.v[j].x, p2);
                p3 = fma(v[j].y, o.v[j].y, p3);
            } else {
                p0 = fma(v[j].x, o.v[j].x, p0);
                p1 = fma(v[j].y, o.v[j].y, p1);
            }
        }
        return (p0 + p1) + (p2 + p3);
    }
    __device__ __forceinline__ void axpy(double a, const KSlice& x)
    {
#pragma unroll
        for (int j = 0; j < E; ++j) {
            v[j].x = fma(a, x.v[j].x, v[j].x);
            v[j].y = fma(a, x.v[j].y, v[j].y);
        }
    }
    // write this lane's slices into a [.. 2*G*E] shared-memory row (group-partial K-vector)
    __device__ __forceinline__ void to_smem(double* row, int g) const
    {
#pragma unroll
        for (int j = 0; j < E; ++j) *reinterpret_cast<double2*>(row + 2 * (g + G * j)) = v[j];
    }
};

// row `idx` of a table whose lane offset is already folded into `tg`: one unsigned 32 x 32 -> 64 multiply-add
__device__ __forceinline__ const double* row_at(const double* tg, int idx, int ld32)
{
    return tg + (size_t)(uint32_t)idx * (uint32_t)ld32;
}

// (ld32: the leading dimension as a 32-bit int, so that row * ld is ONE IMAD.WIDE instead of a 64 x 64 multiply)
// Software-pipelined walk over the `cnt` (<= 32) rows whose indices the lanes hold in `my_idx` (lane i holds row
// i): sub-batch s gives group r the row of lane s*R + r.  Two register buffers alternate (explicit ping-pong: no
// register copies), the loads of sub-batch s+1 are issued before sub-batch s is consumed.
// body(s, slice) is called once per sub-batch, warp-uniformly.
template <int G, int E, bool PREFETCH, class Body>
__device__ __forceinline__ void walk_rows(const double* __restrict__ table, int ld32, int my_idx, int cnt, int g,
                                          int r, bool last_ok, Body&& body)
{
    constexpr int R = 32 / G;
    const int nsub = (cnt + R - 1) / R;
    const double* tg = table + 2 * g;
    KSlice<G, E> A;
    if (PREFETCH) {
        KSlice<G, E> B;
        A.load_at(row_at(tg, __shfl_sync(0xffffffffu, my_idx, r), ld32), last_ok);
        int s = 0;
        while (true) {
            if (s + 1 < nsub) B.load_at(row_at(tg, __shfl_sync(0xffffffffu, my_idx, (s + 1) * R + r), ld32), last_ok);
            body(s, A);
            if (++s >= nsub) break;
            if (s + 1 < nsub) A.load_at(row_at(tg, __shfl_sync(0xffffffffu, my_idx, (s + 1) * R + r), ld32), last_ok);
            body(s, B);
            if (++s >= nsub) break;
        }
    } else {
        for (int s = 0; s < nsub; ++s) {
            A.load_at(row_at(tg, __shfl_sync(0xffffffffu, my_idx, s * R + r), ld32), last_ok);
            body(s, A);
        }
    }
}

// The same walk with the row pipeline carried ACROSS 32-row batches: A arrives holding sub-batch 0 of this batch
// (loaded by the caller or by the previous call) and leaves holding sub-batch 0 of the next batch (rows `next_idx`,
// fetched by the caller one batch ahead), whose loads are issued as soon as A's last use in this batch is over.
// Without this every batch starts with one fully exposed load latency — 27 % of the E-step's stall samples sat on
// that first use (profiles/r01h source view).
template <int G, int E, class Body>
__device__ __forceinline__ void walk_rows_carried(const double* __restrict__ table, int ld32, int my_idx, int cnt,
                                                  int next_idx, bool has_next, KSlice<G, E>& A, int g, int r,
                                                  bool last_ok, Body&& body)
{
    constexpr int R = 32 / G;
    const int nsub = (cnt + R - 1) / R;
    const double* tg = table + 2 * g;
    KSlice<G, E> B;
    int s = 0;
    while (true) {
        if (s + 1 < nsub) B.load_at(row_at(tg, __shfl_sync(0xffffffffu, my_idx, (s + 1) * R + r), ld32), last_ok);
        body(s, A);
        if (++s >= nsub) {  // odd number of sub-batches (a partial batch): A is free only now
            if (has_next) A.load_at(row_at(tg, __shfl_sync(0xffffffffu, next_idx, r), ld32), last_ok);
            break;
        }
        if (s + 1 < nsub)
            A.load_at(row_at(tg, __shfl_sync(0xffffffffu, my_idx, (s + 1) * R + r), ld32), last_ok);
        else if (has_next)
            A.load_at(row_at(tg, __shfl_sync(0xffffffffu, next_idx, r), ld32), last_ok);
        body(s, B);
        if (++s >= nsub) break;
    }
}

// Sum over each aligned quad of lanes on the FP64 tensor core: one DMMA m8n8k4 against a matrix of ones
// (A[lane / 4][lane % 4] = this lane's value, B = 1, C = 0  =>  D[lane / 4][*] = the quad's sum, in every lane of
// the quad).  It replaces two dependent shuffle + add steps — each a 64-bit SHFL pair and a ~30-cycle DADD that an
// in-order warp cannot issue past — on the dot -> reduce -> divide -> axpy chain of every sub-batch
// (tools/membench7.cu: a lone warp's sub-batch drops from 738 to 547 cycles, 12 warps/SM go from 11.8 to 14.2 TB/s).
// Must be executed by all 32 lanes (mma.sync).
__device__ __forceinline__ void dmma_8x8x4(double a, double b, double& d0, double& d1)
{
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%4, %5};"
        : "=d"(d0), "=d"(d1)
        : "d"(a), "d"(b), "d"(0.0), "d"(0.0));
}

__device__ __forceinline__ double quad_sum_mma(double p)
{
    double d0, d1;
    dmma_8x8x4(p, 1.0, d0, d1);
    return d0;
}

// Sum over 16 or 32 lanes with three DMMAs and no shuffle at all (a 64-bit shuffle + add step costs an in-order warp
// ~100 cycles, a DMMA a fraction of that):
//   1. A = the lanes' values, B = ones            -> every lane holds the sum of its quad (row lane/4 of D)
//   2. B[k][n] = quad sum held by lane 4n + k, A = selector of k = 0
//                                                 -> D[i][n] = quad sum n: lane l receives quads 2j, 2j+1 (j = l % 4);
//                                                    their sum is the sum of lanes 8j .. 8j+7
//   3. A[i][k] = that 8-lane sum k (masked to the pairs of the row's own group for G = 16), B = ones
//                                                 -> every lane holds the sum of its group.
// All additions are exact-order FMAs inside the tensor core: deterministic, identical in every lane of the group.
template <int G>
__device__ __forceinline__ double wide_sum_mma(double p, int lane)
{
    static_assert(G == 16 || G == 32, "");
    double d0, d1;
    dmma_8x8x4(p, 1.0, d0, d1);
    dmma_8x8x4((lane & 3) == 0 ? 1.0 : 0.0, d0, d0, d1);
    double s = d0 + d1;  // sum of lanes 8j .. 8j+7, j = lane % 4
    if (G == 16 && ((lane & 3) >> 1) != (lane >> 4)) s = 0.;
    dmma_8x8x4(s, 1.0, d0, d1);
    return d0;
}

// sum over the G lanes of a group (result in every lane of the group; identical bits in all of them).  Warp-uniform.
template <int G>
__device__ __forceinline__ double group_sum(double p)
{
    if constexpr (G >= 16) {
        unsigned lane;
        asm("mov.u32 %0, %%laneid;" : "=r"(lane));
        return wide_sum_mma<G>(p, (int)lane);
    } else if constexpr (G >= 4) {
        p = quad_sum_mma(p);
#pragma unroll
        for (int o = 4; o < G; o <<= 1) p += __shfl_xor_sync(0xffffffffu, p, o);
    } else {
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) p += __shfl_xor_sync(0xffffffffu, p, o);
    }
    return p;
}

// (G, E) for a leading dimension: the smallest group that keeps E <= 8 slices per lane (E <= 16 for full warps)
struct GroupShape {
    int G, E;
};

inline GroupShape group_shape_for(int64_t ld)
{
    for (int G = 2; G <= 32; G *= 2) {
        const int E = (int)((ld + 2 * G - 1) / (2 * G));
        if (E <= 8 || G == 32) return {G, E};
    }
    return {32, 16};
}

// Expands X(G, E) for every instantiated shape.
#define TTLDA_FOR_EACH_SHAPE(X)                                                                      \
    X(2, 1) X(2, 2) X(2, 3) X(2, 4) X(2, 5) X(2, 6) X(2, 7) X(2, 8)                                   \
    X(4, 5) X(4, 6) X(4, 7) X(4, 8)                                                                   \
    X(8, 5) X(8, 6) X(8, 7) X(8, 8)                                                                   \
    X(16, 2) X(16, 3) X(16, 4) X(16, 5) X(16, 6) X(16, 7) X(16, 8)                                    \
    X(32, 2) X(32, 3) X(32, 4) X(32, 5) X(32, 6) X(32, 7) X(32, 8) X(32, 9) X(32, 10) X(32, 11) X(32, 12) X(32, 13) X(32, 14) \
    X(32, 15) X(32, 16)

// the shapes with E <= 8 (register double-buffering of rows): every ld <= 512
#define TTLDA_FOR_EACH_SMALL_SHAPE(X)                                                                \
    X(2, 1) X(2, 2) X(2, 3) X(2, 4) X(2, 5) X(2, 6) X(2, 7) X(2, 8)                                   \
    X(4, 5) X(4, 6) X(4, 7) X(4, 8)                                                                   \
    X(8, 5) X(8, 6) X(8, 7) X(8, 8)                                                                   \
    X(16, 5) X(16, 6) X(16, 7) X(16, 8)                                                               \
    X(32, 5) X(32, 6) X(32, 7) X(32, 8)

}  // namespace ttlda
