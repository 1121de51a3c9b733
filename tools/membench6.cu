// tools/membench6.cu — can L1 be made to keep the hot rows?  (development aid)
// Zipfian row ids over a 40 MB table; rows of the NHOT most frequent ids are loaded with L1::evict_last, all others
// with L1::no_allocate (predicated pair of loads), against the plain ld.global.nc walk.  Full E-step inner loop.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "../tuotuo_b200/csrc/ttlda_common.cuh"
#include "../tuotuo_b200/csrc/traverse.cuh"
using namespace ttlda;

__device__ __forceinline__ double2 ld_hot(const double* p)
{
    double2 v;
    asm volatile("ld.global.nc.L1::evict_last.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ double2 ld_cold(const double* p)
{
    double2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}

template <int G, int E, int MODE>  // MODE 0: plain, 1: hot/cold hints, 2: everything no_allocate
__device__ __forceinline__ void load_row(KSlice<G, E>& s, const double* row, bool hot, int g, bool last_ok)
{
#pragma unroll
    for (int j = 0; j < E; ++j) {
        const double* p = row + 2 * (g + G * j);
        const bool ok = (j < E - 1) || last_ok;
        if (MODE == 0) s.v[j] = ok ? ldg2(p) : make_double2(0., 0.);
        else if (MODE == 2) s.v[j] = ok ? ld_cold(p) : make_double2(0., 0.);
        else {
            s.v[j] = make_double2(0., 0.);
            if (ok && hot) s.v[j] = ld_hot(p);
            if (ok && !hot) s.v[j] = ld_cold(p);
        }
    }
}

template <int G, int E, int MODE, bool FULL>
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
        for (long n0 = b; n0 < b + per_warp; n0 += 32) {
            const int my_w = nx_w;
            if (n0 + 32 < b + per_warp) nx_w = idx[n0 + 32 + lane];
            KSlice<G, E> A, B;
            auto ld_sub = [&](KSlice<G, E>& dst, int s) {
                const int w = __shfl_sync(0xffffffffu, my_w, s * R + r);
                load_row<G, E, MODE>(dst, table + (long)(w & 0x7fffffff) * ld, w < 0, g, last_ok);
            };
            auto body = [&](const KSlice<G, E>& row) {
                if (!FULL) {
#pragma unroll
                    for (int j = 0; j < E; ++j) sink += row.v[j].x + row.v[j].y;
                } else {
                    const double p = group_sum<G>(t.dot(row));
                    acc.axpy(fast_div_pos(1., p + 1e-100), row);
                }
            };
            ld_sub(A, 0);
            for (int s = 0; s < 8; s += 2) {
                ld_sub(B, s + 1);
                body(A);
                if (s + 2 < 8) ld_sub(A, s + 2);
                body(B);
            }
        }
    }
#pragma unroll
    for (int j = 0; j < E; ++j) sink += acc.v[j].x + acc.v[j].y;
    if (sink == 12345.678) out[0] = sink;
}

template <int MODE, bool FULL>
void run(const char* name, const double* table, const int* idx, long n, int ld, double* out)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    int grid = 148 * 16;
    for (int w = 0; w < 2; ++w) k<8, 7, MODE, FULL><<<grid, 128>>>(table, idx, n, ld, 192, out);
    cudaEventRecord(e0);
    const int reps = 4;
    for (int w = 0; w < reps; ++w) k<8, 7, MODE, FULL><<<grid, 128>>>(table, idx, n, ld, 192, out);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= reps;
    printf("%-28s mode=%d full=%d : %8.3f ms  %8.1f GB/s   %s\n", name, MODE, (int)FULL, ms, (double)n * ld * 8 / ms / 1e6,
           cudaGetErrorString(cudaGetLastError()));
}

int main(int argc, char** argv)
{
    const int ld = 100;
    const int nhot = argc > 1 ? atoi(argv[1]) : 128;
    const long n = 192L * 148 * 16 * 4 * 22;
    double* out; cudaMalloc(&out, 8);
    const long rows = 50000;
    double* table; cudaMalloc(&table, rows * ld * 8);
    std::vector<double> ht(rows * ld);
    for (size_t i = 0; i < ht.size(); ++i) ht[i] = 1e-3 + (i % 97) * 1e-5;
    cudaMemcpy(table, ht.data(), ht.size() * 8, cudaMemcpyHostToDevice);
    std::vector<int> h(n);
    unsigned long long s = 88172645463325252ULL;
    // document-frequency-like law: P(rank) ~ min(1, c / rank) flattened at the head (the top words are in every
    // document once), approximated by 1/(rank + 40)
    std::vector<double> cdf(rows);
    double tot = 0; for (long i = 0; i < rows; ++i) { tot += 1.0 / (i + 40); cdf[i] = tot; }
    long nh = 0;
    for (long i = 0; i < n; ++i) {
        s ^= s << 13; s ^= s >> 7; s ^= s << 17;
        double u = (double)(s >> 11) / 9007199254740992.0 * tot;
        long lo = 0, hi = rows - 1;
        while (lo < hi) { long m = (lo + hi) / 2; if (cdf[m] < u) lo = m + 1; else hi = m; }
        h[i] = (int)((lo * 7919L) % rows);
        if (lo < nhot) { h[i] |= 0x80000000; ++nh; }
    }
    printf("hot share of the stream: %.3f (nhot=%d)\n", (double)nh / n, nhot);
    int* idx; cudaMalloc(&idx, n * 4);
    cudaMemcpy(idx, h.data(), n * 4, cudaMemcpyHostToDevice);
    run<0, false>("plain ld.global.nc", table, idx, n, ld, out);
    run<1, false>("hot evict_last/cold no_alloc", table, idx, n, ld, out);
    run<2, false>("all no_allocate", table, idx, n, ld, out);
    run<0, true>("plain ld.global.nc", table, idx, n, ld, out);
    run<1, true>("hot evict_last/cold no_alloc", table, idx, n, ld, out);
    run<2, true>("all no_allocate", table, idx, n, ld, out);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
