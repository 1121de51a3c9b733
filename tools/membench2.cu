// tools/membench2.cu — which part of the per-nonzero work costs the gather its bandwidth?  (development aid)
// Same traversal as the product kernels (group-mapped slices, ping-pong prefetch, 32-index batches), with the
// arithmetic enabled step by step:  LEVEL 0 = loads only, 1 = + dot, 2 = + group reduction, 3 = + division,
// 4 = + axpy (full inner loop).
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "../tuotuo_b200/csrc/traverse.cuh"
using namespace ttlda;

template <int G, int E, int LEVEL, int MINB>
__global__ void __launch_bounds__(128, MINB) k(const double* __restrict__ table, const int* __restrict__ idx,
                                              const int* __restrict__ cnts, long n, int ld, int per_warp, double* out)
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
        int nx_w = idx[b + lane], nx_c = cnts[b + lane];
        for (long n0 = b; n0 < b + per_warp; n0 += 32) {
            const int my_w = nx_w, my_c = nx_c;
            if (n0 + 32 < b + per_warp) { nx_w = idx[n0 + 32 + lane]; nx_c = cnts[n0 + 32 + lane]; }
            walk_rows<G, E, true>(table, ld, my_w, 32, g, r, last_ok, [&](int s, const KSlice<G, E>& row) {
                if (LEVEL == 0) {
#pragma unroll
                    for (int j = 0; j < E; ++j) sink += row.v[j].x + row.v[j].y;
                } else {
                    double p = t.dot(row);
                    if (LEVEL >= 2) p = group_sum<G>(p);
                    double rr = p;
                    if (LEVEL >= 3) rr = (double)__shfl_sync(0xffffffffu, my_c, s * R + r) / (p + 1e-100);
                    if (LEVEL >= 4) acc.axpy(rr, row); else sink += rr;
                }
            });
        }
    }
#pragma unroll
    for (int j = 0; j < E; ++j) sink += acc.v[j].x + acc.v[j].y;
    if (sink == 12345.678) out[0] = sink;
}

template <int G, int E, int LEVEL, int MINB>
void run(const char* name, const double* table, const int* idx, const int* cnts, long n, int ld, double* out)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    int grid = 148 * 16;
    for (int w = 0; w < 2; ++w) k<G, E, LEVEL, MINB><<<grid, 128>>>(table, idx, cnts, n, ld, 192, out);
    cudaEventRecord(e0);
    const int reps = 4;
    for (int w = 0; w < reps; ++w) k<G, E, LEVEL, MINB><<<grid, 128>>>(table, idx, cnts, n, ld, 192, out);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= reps;
    int nb = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k<G, E, LEVEL, MINB>, 128, 0);
    printf("%-26s G=%2d E=%d level=%d minb=%d ctas/sm=%d : %8.3f ms  %8.1f GB/s   %s\n", name, G, E, LEVEL, MINB, nb, ms,
           (double)n * ld * 8 / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
}

int main(int argc, char** argv)
{
    const int ld = argc > 1 ? atoi(argv[1]) : 100;
    const long n = 192L * 148 * 16 * 4 * 22;  // ~40M rows
    double* out; cudaMalloc(&out, 8);
    for (int zipf = 0; zipf < 2; ++zipf) {
        const long rows = 50000;
        double* table; cudaMalloc(&table, rows * ld * 8);
        std::vector<double> ht(rows * ld);
        for (size_t i = 0; i < ht.size(); ++i) ht[i] = 1e-3 + (i % 97) * 1e-5;
        cudaMemcpy(table, ht.data(), ht.size() * 8, cudaMemcpyHostToDevice);
        std::vector<int> h(n), hc(n, 1);
        unsigned long long s = 88172645463325252ULL;
        std::vector<double> cdf(rows);
        double tot = 0; for (long i = 0; i < rows; ++i) { tot += 1.0 / (i + 1); cdf[i] = tot; }
        for (long i = 0; i < n; ++i) {
            s ^= s << 13; s ^= s >> 7; s ^= s << 17;
            if (!zipf) h[i] = (int)(s % rows);
            else {
                double u = (double)(s >> 11) / 9007199254740992.0 * tot;
                long lo = 0, hi = rows - 1;
                while (lo < hi) { long m = (lo + hi) / 2; if (cdf[m] < u) lo = m + 1; else hi = m; }
                h[i] = (int)((lo * 7919L) % rows);  // scatter the hot ids over the table
            }
        }
        int *idx, *cnts; cudaMalloc(&idx, n * 4); cudaMalloc(&cnts, n * 4);
        cudaMemcpy(idx, h.data(), n * 4, cudaMemcpyHostToDevice);
        cudaMemcpy(cnts, hc.data(), n * 4, cudaMemcpyHostToDevice);
        const char* name = zipf ? "zipf rows, 40 MB table" : "uniform rows, 40 MB table";
        run<8, 7, 0, 3>(name, table, idx, cnts, n, ld, out);
        run<8, 7, 1, 3>(name, table, idx, cnts, n, ld, out);
        run<8, 7, 2, 3>(name, table, idx, cnts, n, ld, out);
        run<8, 7, 3, 3>(name, table, idx, cnts, n, ld, out);
        run<8, 7, 4, 3>(name, table, idx, cnts, n, ld, out);
        run<8, 7, 0, 1>(name, table, idx, cnts, n, ld, out);
        run<8, 7, 4, 1>(name, table, idx, cnts, n, ld, out);
        run<16, 4, 0, 4>(name, table, idx, cnts, n, ld, out);
        run<16, 4, 4, 4>(name, table, idx, cnts, n, ld, out);
        cudaFree(idx); cudaFree(cnts); cudaFree(table);
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
