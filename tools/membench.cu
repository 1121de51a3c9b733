// tools/membench.cu — what gather bandwidth does this B200 actually give?  (development aid, not product code)
//   rows of `row_doubles` fp64 gathered from a table of `table_mb` MB by random / sequential row indices,
//   G lanes per row (16-byte loads), results reduced to defeat dead-code elimination.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/membench tools/membench.cu && tools/membench
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <vector>

template <int G, int E, int UNROLL>
__global__ void __launch_bounds__(256) gather_kernel(const double* __restrict__ table, const int* __restrict__ idx,
                                                    long n, int ld, double* out)
{
    constexpr int R = 32 / G;
    const int lane = threadIdx.x & 31, g = lane % G, r = lane / G;
    const long warp0 = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long nwarps = ((long)gridDim.x * blockDim.x) >> 5;
    double acc = 0;
    // each warp handles items [i, i + R*UNROLL) per iteration
    for (long i = warp0 * R * UNROLL; i + R * UNROLL <= n; i += nwarps * R * UNROLL) {
        double2 v[UNROLL][E];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const double* row = table + (long)idx[i + u * R + r] * ld;
#pragma unroll
            for (int j = 0; j < E; ++j) {
                const int k = 2 * (g + G * j);
                v[u][j] = (k < ld) ? __ldg(reinterpret_cast<const double2*>(row + k)) : make_double2(0, 0);
            }
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u)
#pragma unroll
            for (int j = 0; j < E; ++j) acc += v[u][j].x + v[u][j].y;
    }
    if (acc == 12345.678) out[0] = acc;
}

template <int G, int E, int UNROLL>
void run(const char* name, const double* table, const int* idx, long n, int ld, double* out, int blocks_per_sm)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    int grid = 148 * blocks_per_sm;
    for (int w = 0; w < 2; ++w) gather_kernel<G, E, UNROLL><<<grid, 256>>>(table, idx, n, ld, out);
    cudaEventRecord(e0);
    const int reps = 5;
    for (int w = 0; w < reps; ++w) gather_kernel<G, E, UNROLL><<<grid, 256>>>(table, idx, n, ld, out);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    ms /= reps;
    printf("%-44s G=%2d E=%d unroll=%d ctas/sm=%d : %8.3f ms  %8.1f GB/s\n", name, G, E, UNROLL, blocks_per_sm, ms,
           (double)n * ld * 8 / ms / 1e6);
}

int main(int argc, char** argv)
{
    const int ld = 100;
    const long n = 40L * 1000 * 1000;  // gathered rows per launch (32 GB)
    double* out;
    cudaMalloc(&out, 8);
    for (int cfg = 0; cfg < 3; ++cfg) {
        const long rows = cfg == 0 ? 50000 : (cfg == 1 ? 1000000 : 62500);  // 40 MB (L2), 800 MB (HBM), 50 MB window
        double* table;
        cudaMalloc(&table, rows * ld * 8);
        cudaMemset(table, 0, rows * ld * 8);
        std::vector<int> h(n);
        unsigned long long s = 88172645463325252ULL;
        for (long i = 0; i < n; ++i) {
            s ^= s << 13; s ^= s >> 7; s ^= s << 17;
            h[i] = (int)(s % rows);
        }
        int* idx;
        cudaMalloc(&idx, n * 4);
        cudaMemcpy(idx, h.data(), n * 4, cudaMemcpyHostToDevice);
        char name[64];
        snprintf(name, sizeof name, "random rows of 800 B from %ld MB", rows * ld * 8 / 1000000);
        run<8, 7, 1>(name, table, idx, n, ld, out, 8);
        run<8, 7, 2>(name, table, idx, n, ld, out, 8);
        run<8, 7, 4>(name, table, idx, n, ld, out, 4);
        run<16, 4, 2>(name, table, idx, n, ld, out, 8);
        run<16, 4, 4>(name, table, idx, n, ld, out, 8);
        run<32, 2, 4>(name, table, idx, n, ld, out, 8);
        run<32, 2, 8>(name, table, idx, n, ld, out, 8);
        cudaFree(idx);
        cudaFree(table);
    }
    // plain streaming read of 4 GB for reference
    {
        const long rows = 5000000;
        double* table;
        cudaMalloc(&table, rows * ld * 8);
        cudaMemset(table, 0, rows * ld * 8);
        std::vector<int> h(n);
        for (long i = 0; i < n; ++i) h[i] = (int)(i % rows);
        int* idx;
        cudaMalloc(&idx, n * 4);
        cudaMemcpy(idx, h.data(), n * 4, cudaMemcpyHostToDevice);
        run<32, 2, 8>("sequential rows from 4000 MB (HBM stream)", table, idx, n, ld, out, 8);
        run<8, 7, 4>("sequential rows from 4000 MB (HBM stream)", table, idx, n, ld, out, 4);
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
