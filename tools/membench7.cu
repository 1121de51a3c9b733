// tools/membench7.cu — is the full inner loop latency-bound (throughput ~ resident warps) or saturating something?
// membench2's level-4 loop (dot, reduce, fast division, axpy) and level 0 (loads only) with occupancy limited to
// 1..3 CTAs/SM through dynamic shared memory.  (development aid)
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "../tuotuo_b200/csrc/ttlda_common.cuh"
#include "../tuotuo_b200/csrc/traverse.cuh"
using namespace ttlda;

// sum over each aligned group of 8 lanes through the FP64 tensor core: one DMMA m8n8k4 against a matrix of ones sums
// the 4 lanes of every quad (A[lane/4][lane%4] = this lane's value, B = 1), one shuffle + add joins the two quads.
__device__ __forceinline__ double group8_sum_mma(double p)
{
    double d0, d1;
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%4, %5};"
                 : "=d"(d0), "=d"(d1)
                 : "d"(p), "d"(1.0), "d"(0.0), "d"(0.0));
    return d0 + __shfl_xor_sync(0xffffffffu, d0, 4);
}

template <int G, int E, int FULL>
__global__ void __launch_bounds__(128, 3) k(const double* __restrict__ table, const int* __restrict__ idx, long n, int ld,
                                           int per_warp, double* out)
{
    constexpr int R = 32 / G;
    const int lane = threadIdx.x & 31, g = lane % G, r = lane / G;
    const long warp0 = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long nwarps = ((long)gridDim.x * blockDim.x) >> 5;
    const bool last_ok = KSlice<G, E>::last_slice_ok(g, ld);
    KSlice<G, E> t, acc;
    t.load(table + (warp0 % 1000) * ld, g, last_ok);
    acc.zero();
    double sink = 0;
    for (long b = warp0 * per_warp; b + per_warp <= n; b += nwarps * per_warp) {
        int nx_w = idx[b + lane];
        KSlice<G, E> rowA;
        rowA.load_at(row_at(table + 2 * g, __shfl_sync(0xffffffffu, nx_w, r), ld), last_ok);
        for (long n0 = b; n0 < b + per_warp; n0 += 32) {
            const int my_w = nx_w;
            const bool more = n0 + 32 < b + per_warp;
            if (more) nx_w = idx[n0 + 32 + lane];
            walk_rows_carried<G, E>(table, ld, my_w, 32, nx_w, more, rowA, g, r, last_ok, [&](int s, const KSlice<G, E>& row) {
                // FULL: 0 loads only | 1 full chain | 2 dot only | 3 dot + reduce | 4 dot + reduce + divide | 5 axpy only
                if (FULL == 0) {
#pragma unroll
                    for (int j = 0; j < E; ++j) sink += row.v[j].x + row.v[j].y;
                } else if (FULL == 1) {
                    const double p = group_sum<G>(t.dot(row, 1e-100 / G));
                    acc.axpy(fast_div_pos(1., p), row);
                } else if (FULL == 2) {
                    sink += t.dot(row);
                } else if (FULL == 3) {
                    sink += group_sum<G>(t.dot(row));
                } else if (FULL == 4) {
                    sink += fast_div_pos(1., group_sum<G>(t.dot(row, 1e-100 / G)));
                } else if (FULL == 5) {
                    acc.axpy(1.5, row);
                } else if (FULL == 6) {
                    sink += group8_sum_mma(t.dot(row));
                } else {
                    const double p = group8_sum_mma(t.dot(row, 1e-100 / G));
                    acc.axpy(fast_div_pos(1., p), row);
                }
            });
        }
    }
#pragma unroll
    for (int j = 0; j < E; ++j) sink += acc.v[j].x + acc.v[j].y;
    if (sink == 12345.678) out[0] = sink;
}

template <int FULL>
void run(int ctas, const double* table, const int* idx, long n, int ld, double* out)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const size_t smem = ctas >= 3 ? 0 : (ctas == 2 ? 100 * 1024 : 200 * 1024);
    cudaFuncSetAttribute(k<8, 7, FULL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int nb = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k<8, 7, FULL>, 128, smem);
    int grid = 148 * 16;
    for (int w = 0; w < 2; ++w) k<8, 7, FULL><<<grid, 128, smem>>>(table, idx, n, ld, 192, out);
    cudaEventRecord(e0);
    const int reps = 4;
    for (int w = 0; w < reps; ++w) k<8, 7, FULL><<<grid, 128, smem>>>(table, idx, n, ld, 192, out);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= reps;
    printf("level=%d ctas/sm=%d (%2d warps) : %8.3f ms  %8.1f GB/s  %7.1f cycles per sub-batch per warp (1.92 GHz)  %s\n", (int)FULL, nb, nb * 4, ms,
           (double)n * ld * 8 / ms / 1e6, ms * 1e-3 * 1.92e9 * (nb * 4 * 148) / ((double)n / 4), cudaGetErrorString(cudaGetLastError()));
}

int main()
{
    const int ld = 100;
    const long n = 192L * 148 * 16 * 4 * 22;
    double* out; cudaMalloc(&out, 8);
    const long rows = 50000;
    double* table; cudaMalloc(&table, rows * ld * 8);
    std::vector<double> ht(rows * ld);
    for (size_t i = 0; i < ht.size(); ++i) ht[i] = 1e-3 + (i % 97) * 1e-5;
    cudaMemcpy(table, ht.data(), ht.size() * 8, cudaMemcpyHostToDevice);
    std::vector<int> h(n);
    unsigned long long s = 88172645463325252ULL;
    std::vector<double> cdf(rows);
    double tot = 0; for (long i = 0; i < rows; ++i) { tot += 1.0 / (i + 1); cdf[i] = tot; }
    for (long i = 0; i < n; ++i) {
        s ^= s << 13; s ^= s >> 7; s ^= s << 17;
        double u = (double)(s >> 11) / 9007199254740992.0 * tot;
        long lo = 0, hi = rows - 1;
        while (lo < hi) { long m = (lo + hi) / 2; if (cdf[m] < u) lo = m + 1; else hi = m; }
        h[i] = (int)((lo * 7919L) % rows);
    }
    int* idx; cudaMalloc(&idx, n * 4);
    cudaMemcpy(idx, h.data(), n * 4, cudaMemcpyHostToDevice);
    for (int c = 1; c <= 3; c += 2) {
        run<0>(c, table, idx, n, ld, out);
        run<2>(c, table, idx, n, ld, out);
        run<3>(c, table, idx, n, ld, out);
        run<4>(c, table, idx, n, ld, out);
        run<5>(c, table, idx, n, ld, out);
        run<1>(c, table, idx, n, ld, out);
        run<6>(c, table, idx, n, ld, out);
        run<7>(c, table, idx, n, ld, out);
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
