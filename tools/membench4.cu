// tools/membench4.cu — does the TMA gather4 instruction (sm_100: cp.async.bulk.tensor.2d ... tile::gather4) feed the
// K-vector gather faster than register-staged LDG?  (development aid)
// One warp = producer and consumer: lane 0 issues ONE gather4 per sub-batch (4 rows of ld doubles -> one smem stage,
// completion on an mbarrier), NSTAGE sub-batches ahead; all lanes then read their 16-byte slices with LDS.128 and run
// the same dot / group-reduce / divide / axpy chain as the product kernels (LEVEL 4) or just sum (LEVEL 0).
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
#include "../tuotuo_b200/csrc/traverse.cuh"
using namespace ttlda;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void tma_gather4(uint32_t dst, const CUtensorMap* map, int col, int r0, int r1, int r2, int r3, uint32_t bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes "
                 "[%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::"r"(dst), "l"(map), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(bar)
                 : "memory");
}

template <int E, int LEVEL, int NSTAGE>
__global__ void __launch_bounds__(128) k(const __grid_constant__ CUtensorMap map, const double* __restrict__ table,
                                         const int* __restrict__ idx, const int* __restrict__ cnts, long n, int ld,
                                         int per_warp, double* out)
{
    constexpr int G = 8, R = 4;
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x & 31, g = lane % G, r = lane / G, warp = threadIdx.x >> 5;
    const uint32_t row_bytes = ld * 8, stage_bytes = R * row_bytes;
    unsigned char* wbase = smem + (size_t)warp * (NSTAGE * stage_bytes + 128);
    const uint32_t stages = smem_u32(wbase), bars = stages + NSTAGE * stage_bytes;
    if (lane == 0) for (int s = 0; s < NSTAGE; ++s) mbar_init(bars + 8 * s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
    const long warp0 = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long nwarps = ((long)gridDim.x * blockDim.x) >> 5;
    const bool last_ok = KSlice<G, E>::last_slice_ok(g, ld);
    KSlice<G, E> t, acc;
    t.load(table + (warp0 % 1000) * ld, g, last_ok);
    acc.zero();
    double sink = 0;
    uint32_t gq = 0;  // sub-batches consumed so far by this warp
    for (long b = warp0 * per_warp; b + per_warp <= n; b += nwarps * per_warp) {
        const int Q = per_warp / R;  // sub-batches in this segment (per_warp multiple of 32)
        // produce(q): indices of items b + q*4 .. +3 are read straight from global by lane 0 (L1/L2 hits)
        auto produce = [&](int q) {
            if (lane == 0) {
                const int4 w = *reinterpret_cast<const int4*>(idx + b + (long)q * 4);
                const uint32_t st = (gq + q) % NSTAGE;
                mbar_expect_tx(bars + 8 * st, stage_bytes);
                tma_gather4(stages + st * stage_bytes, &map, 0, w.x, w.y, w.z, w.w, bars + 8 * st);
            }
        };
        for (int q = 0; q < NSTAGE - 1 && q < Q; ++q) produce(q);
        for (int q = 0; q < Q; ++q) {
            if (q + NSTAGE - 1 < Q) { __syncwarp(); produce(q + NSTAGE - 1); }
            const uint32_t st = (gq + q) % NSTAGE;
            mbar_wait(bars + 8 * st, ((gq + q) / NSTAGE) & 1);
            const double* row = reinterpret_cast<const double*>(wbase + st * stage_bytes + r * row_bytes);
            KSlice<G, E> x;
#pragma unroll
            for (int j = 0; j < E - 1; ++j) x.v[j] = *reinterpret_cast<const double2*>(row + 2 * (g + G * j));
            x.v[E - 1] = last_ok ? *reinterpret_cast<const double2*>(row + 2 * (g + G * (E - 1))) : make_double2(0., 0.);
            if (LEVEL == 0) {
#pragma unroll
                for (int j = 0; j < E; ++j) sink += x.v[j].x + x.v[j].y;
            } else {
                const double p = group_sum<G>(t.dot(x));
                acc.axpy(1.0 / (p + 1e-100), x);
            }
        }
        __syncwarp();
        gq += Q;
    }
#pragma unroll
    for (int j = 0; j < E; ++j) sink += acc.v[j].x + acc.v[j].y;
    if (sink == 12345.678) out[0] = sink;
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int E, int LEVEL, int NSTAGE>
void run(const CUtensorMap& map, const double* table, const int* idx, const int* cnts, long n, int ld, double* out, int ctas)
{
    const size_t smem = 4 * (size_t)(NSTAGE * 4 * ld * 8 + 128);
    cudaFuncSetAttribute(k<E, LEVEL, NSTAGE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int grid = 148 * ctas;
    for (int w = 0; w < 2; ++w) k<E, LEVEL, NSTAGE><<<grid, 128, smem>>>(map, table, idx, cnts, n, ld, 192, out);
    cudaEventRecord(e0);
    const int reps = 4;
    for (int w = 0; w < reps; ++w) k<E, LEVEL, NSTAGE><<<grid, 128, smem>>>(map, table, idx, cnts, n, ld, 192, out);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= reps;
    int nb = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k<E, LEVEL, NSTAGE>, 128, smem);
    printf("gather4 level=%d nstage=%d grid=%dx148 resident ctas/sm=%d : %8.3f ms  %8.1f GB/s   %s\n", LEVEL, NSTAGE, ctas, nb, ms,
           (double)n * ld * 8 / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
}

int main()
{
    const int ld = 100;
    const long rows = 50000;
    const long n = 192L * 148 * 16 * 4 * 22;
    double* out; cudaMalloc(&out, 8);
    double* table; cudaMalloc(&table, rows * ld * 8);
    std::vector<double> ht(rows * ld);
    for (size_t i = 0; i < ht.size(); ++i) ht[i] = 1e-3 + (i % 97) * 1e-5;
    cudaMemcpy(table, ht.data(), ht.size() * 8, cudaMemcpyHostToDevice);
    std::vector<int> h(n), hc(n, 1);
    unsigned long long s = 88172645463325252ULL;
    for (long i = 0; i < n; ++i) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; h[i] = (int)(s % rows); }
    int *idx, *cnts; cudaMalloc(&idx, n * 4); cudaMalloc(&cnts, n * 4);
    cudaMemcpy(idx, h.data(), n * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(cnts, hc.data(), n * 4, cudaMemcpyHostToDevice);

    EncodeFn encode = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t ce = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &qres);
    if (ce != cudaSuccess || !encode) { printf("no cuTensorMapEncodeTiled: %s\n", cudaGetErrorString(ce)); return 1; }
    CUtensorMap map;
    const cuuint64_t gdim[2] = {(cuuint64_t)ld, (cuuint64_t)rows};
    const cuuint64_t gstride[1] = {(cuuint64_t)ld * 8};
    const cuuint32_t box[2] = {(cuuint32_t)ld, 1};
    const cuuint32_t estr[2] = {1, 1};
    CUresult cr = encode(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, table, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("cuTensorMapEncodeTiled -> %d\n", (int)cr);
    if (cr != CUDA_SUCCESS) return 1;
    run<7, 0, 4>(map, table, idx, cnts, n, ld, out, 4);
    printf("sync: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    run<7, 0, 8>(map, table, idx, cnts, n, ld, out, 2);
    run<7, 4, 4>(map, table, idx, cnts, n, ld, out, 4);
    run<7, 4, 8>(map, table, idx, cnts, n, ld, out, 2);
    run<7, 4, 2>(map, table, idx, cnts, n, ld, out, 8);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
