// ttlda_common.cuh — shared device helpers for libttlda (sm_100a).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "ttlda.h"

namespace ttlda {

// ---------------------------------------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define TTLDA_REQUIRE(cond, msg)                                   \
    do {                                                           \
        if (!(cond)) {                                             \
            ::ttlda::set_error("%s: %s", __func__, msg);           \
            return TTLDA_EINVAL;                                   \
        }                                                          \
    } while (0)

#define TTLDA_LAUNCH_CHECK(what)                                   \
    do {                                                           \
        cudaError_t e__ = cudaGetLastError();                      \
        if (e__ != cudaSuccess) return ::ttlda::cuda_fail(e__, what); \
    } while (0)

constexpr int kSMs = 148;             // B200: 2 dies x 74 SMs
constexpr int kMaxReduceBlocks = 4096; // upper bound on the grid of any kernel that emits partial records
constexpr int NPART = TTLDA_NPART;

inline int64_t reduce_ws_bytes() { return (int64_t)kMaxReduceBlocks * NPART * sizeof(double); }

// ---------------------------------------------------------------------------------------------------------
// warp / block reductions (fixed order => bit-reproducible)
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Sum `NV` values across the CTA; valid result in thread 0.  smem: double[NV * 32].
template <int NV>
__device__ __forceinline__ void block_sum(double (&v)[NV], double* smem)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = warp_sum(v[i]);
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) smem[i * 32 + warp] = v[i];
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            double x = (lane < nwarps) ? smem[i * 32 + lane] : 0.0;
            v[i] = warp_sum(x);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// AS-103 digamma — op-for-op restatement of tuotuo/cutils.pyx:75-93 (the reference's "psi", optimised for speed
// not accuracy).  Parity on gamma/lambda at 1e-9 requires this approximation, not an exact digamma.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double psi_as103(double x)
{
    if (x <= 1e-6) return -0.577215664901532860606512090082402431 - 1. / x;  // cutils.pyx:76-78
    double result = 0.;
    while (x < 6.) {  // cutils.pyx:83-85
        result -= 1. / x;
        x += 1.;
    }
    double r = 1. / x;  // cutils.pyx:89-92
    result += log(x) - .5 * r;
    r = r * r;
    result -= r * ((1. / 12.) - r * ((1. / 120.) - r * (1. / 252.)));
    return result;
}

// Fused AS-103 digamma + log-gamma for the ELBO prior terms (the reference evaluates psi~(x) in cutils.pyx:75-93 and
// scipy.special.gammaln(x) in lda_model.py:41 separately).  Both need the same upward shift to x >= 6, log(x) and
// 1/x, so they are evaluated together:
//   * the AS-103 recurrence  result -= 1/x  (one IEEE division per step, up to six steps for the many gamma_dk ~ 1)
//     is folded into ONE division: sum_i 1/(x+i) = q/p with p = prod_i (x+i), q accumulated by FMA — same value to
//     ~1e-16 relative (parity needs 1e-9), ~5x fewer instructions;
//   * lgamma(x) = Stirling(x_shifted) - log(p), series through x^-13 (|err| < 6e-14 absolute for x_shifted >= 6).
//   * callers that only need SUMS of lgamma (the ELBO prior terms) take the Stirling part and the shift product p
//     separately (psi_lgamma_as103_split): sum_k log(p_k) = log(prod_k p_k) costs one log per lane instead of one
//     per element; the two remaining quotients use the reciprocal approximation + Newton (fast_rcp_pos, ~2 ulp).
__device__ __forceinline__ double fast_rcp_pos(double x)
{
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    y = y * fma(-x, y, 2.);
    y = y * fma(-x, y, 2.);
    return y;
}

// psi_out = AS-103 digamma(x);  lgamma(x) = lgam_stirling - log(p_out)   (p_out = 1 when no shift was needed)
__device__ __forceinline__ void psi_lgamma_as103_split(double x, double& psi_out, double& lgam_stirling, double& p_out)
{
    if (x <= 1e-6) {  // cutils.pyx:76-78
        psi_out = -0.577215664901532860606512090082402431 - 1. / x;
        lgam_stirling = lgamma(x);
        p_out = 1.;
        return;
    }
    double p = 1., q = 0.;
    while (x < 6.) {  // cutils.pyx:83-85
        q = fma(q, x, p);
        p *= x;
        x += 1.;
    }
    const double r = fast_rcp_pos(x), lx = log(x), r2 = r * r;  // cutils.pyx:89-92
    double res = lx - .5 * r;
    res -= r2 * ((1. / 12.) - r2 * ((1. / 120.) - r2 * (1. / 252.)));
    psi_out = res - q * fast_rcp_pos(p);
    const double series =
        r * ((1. / 12.) -
             r2 * ((1. / 360.) -
                   r2 * ((1. / 1260.) -
                         r2 * ((1. / 1680.) - r2 * ((1. / 1188.) - r2 * ((691. / 360360.) - r2 * (1. / 156.)))))));
    lgam_stirling = (x - .5) * lx - x + 0.91893853320467274178 + series;
    p_out = p;
}

__device__ __forceinline__ void psi_lgamma_as103(double x, double& psi_out, double& lgam_out)
{
    double st, p;
    psi_lgamma_as103_split(x, psi_out, st, p);
    lgam_out = (p == 1.) ? st : st - log(p);
}

// Exact digamma for x > 0 (recurrence to x >= 10, then the asymptotic series through x^-14; |err| ~ 1e-16
// relative).  Used only where the reference calls scipy.special.psi (hyper-parameter Newton updates).
__device__ __forceinline__ double psi_exact(double x)
{
    double shift = 0.;
    while (x < 10.) {
        shift -= 1. / x;
        x += 1.;
    }
    const double z = 1. / (x * x);
    double p = 8.33333333333333333333e-2;
    p = p * z - 2.10927960927960927961e-2;
    p = p * z + 7.57575757575757575758e-3;
    p = p * z - 4.16666666666666666667e-3;
    p = p * z + 3.96825396825396825397e-3;
    p = p * z - 8.33333333333333333333e-3;
    p = p * z + 8.33333333333333333333e-2;
    return log(x) - 0.5 / x - z * p + shift;
}

// c / x for a normal positive x, without the IEEE division's special-case path: hardware reciprocal approximation
// y0 (MUFU.RCP64H, relative error e ~ 2^-23) corrected as c*y0*(1 + e + e^2), e = 1 - x*y0 (error e^3 ~ 1e-21, then
// rounding: within ~2 ulp of the correctly rounded quotient; the parity budget on gamma / lambda is 1e-9 relative).
// The division sits on the dependent chain of every sub-batch, and a dependent FP64 operation costs ~30 cycles on this
// part (tools/membench7.cu: a lone warp needs ~750 cycles of chain per sub-batch): this form is 3 levels deep after
// the MUFU (e | c*y0 -> e + e^2 -> fma) where two Newton steps + the multiply are 5, the IEEE division ~9 and a branch.
__device__ __forceinline__ double fast_div_pos(double c, double x)
{
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double e = fma(-x, y, 1.);
    const double r = c * y;
    const double e2 = fma(e, e, e);
    return fma(r, e2, r);
}

// 16-byte global loads of a K-vector slice.  The beta tables are re-read by many documents: keep them on the
// read-only path; streaming data (CSR arrays) goes through L1::no_allocate.
__device__ __forceinline__ double2 ldg2(const double* p) { return __ldg(reinterpret_cast<const double2*>(p)); }

// L2 eviction policy for data that is touched once per pass (CSR/CSC index streams, theta/gamma tables, statistics
// output): evict-first, so that the streams do not push the re-used K-vector tables out of the 126 MB L2.
__device__ __forceinline__ uint64_t l2_evict_first_policy()
{
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}

__device__ __forceinline__ int ld_stream_s32(const int32_t* p, uint64_t pol)
{
    int v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.s32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
    return v;
}

__device__ __forceinline__ double ld_stream_f64(const double* p, uint64_t pol)
{
    double v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol));
    return v;
}

__device__ __forceinline__ void st_stream_f64(double* p, double v, uint64_t pol)
{
    asm volatile("st.global.L1::no_allocate.L2::cache_hint.f64 [%0], %1, %2;" ::"l"(p), "d"(v), "l"(pol) : "memory");
}

// final fixed-order reduction of per-CTA records: out[j] = sum_b rec[b][j]
__global__ void reduce_records_kernel(const double* __restrict__ rec, int nrec, double* __restrict__ out);
int launch_reduce_records(const double* rec, int nrec, double* out, cudaStream_t stream);

inline int grid_for_warps(int64_t nwarps_needed, int warps_per_block, int blocks_per_sm)
{
    int64_t blocks = (nwarps_needed + warps_per_block - 1) / warps_per_block;
    int64_t cap = (int64_t)kSMs * blocks_per_sm;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

}  // namespace ttlda
