// tools/membench5.cu — does a deeper row pipeline pay once theta lives in shared memory?  (development aid)
// Full E-step inner loop (dot, group reduction, fast division, axpy) over 800-byte rows of a 40 MB table with
//   NBUF = 2 | 3 register row buffers (prefetch distance 1 | 2), continuous across 32-row batches,
//   TSMEM = theta K-vector in registers | in shared memory (frees 4*E registers for the third buffer).
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "../tuotuo_b200/csrc/ttlda_common.cuh"
#include "../tuotuo_b200/csrc/traverse.cuh"
using namespace ttlda;

template <int G, int E, int NBUF, bool TSMEM, int MINB>
__global__ void __launch_bounds__(128, MINB) k(const double* __restrict__ table, const int* __restrict__ idx,
                                              const int* __restrict__ cnts, long n, int ld, int per_warp, double* out)
{
    constexpr int R = 32 / G, SPB = 32 / R;  // sub-batches per 32-row batch
    extern __shared__ double smem[];
    double* ts = smem + (threadIdx.x >> 5) * (2 * G * E);
    const int lane = threadIdx.x & 31, g = lane % G, r = lane / G;
    const long warp0 = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long nwarps = ((long)gridDim.x * blockDim.x) >> 5;
    const bool last_ok = KSlice<G, E>::last_slice_ok(g, ld);
    KSlice<G, E> t, acc;
    t.load(table + (warp0 % 1000) * ld, g, last_ok);
    if (TSMEM) {
        t.to_smem(ts, g);
        __syncwarp();
    }
    acc.zero();
    double sink = 0;
    for (long b = warp0 * per_warp; b + per_warp <= n; b += nwarps * per_warp) {
        int cur_w = idx[b + lane], cur_c = cnts[b + lane];
        int nxt_w = idx[b + 32 + lane], nxt_c = cnts[b + 32 + lane];
        int curbatch = 0;
        const int Q = per_warp / R;
        auto row_of = [&](int q) {
            const int sel = ((q / SPB) == curbatch) ? cur_w : nxt_w;
            return table + (long)__shfl_sync(0xffffffffu, sel, (q % SPB) * R + r) * ld;
        };
        auto body = [&](int q, const KSlice<G, E>& row) {
            const int c = __shfl_sync(0xffffffffu, cur_c, (q % SPB) * R + r);
            double p;
            if (TSMEM) {
                double p0 = 0, p1 = 0, p2 = 0, p3 = 0;
#pragma unroll
                for (int j = 0; j < E; ++j) {
                    const double2 tv = *reinterpret_cast<const double2*>(ts + 2 * (g + G * j));
                    if (j & 1) { p2 = fma(tv.x, row.v[j].x, p2); p3 = fma(tv.y, row.v[j].y, p3); }
                    else { p0 = fma(tv.x, row.v[j].x, p0); p1 = fma(tv.y, row.v[j].y, p1); }
                }
                p = (p0 + p1) + (p2 + p3);
            } else {
                p = t.dot(row);
            }
            p = group_sum<G>(p);
            acc.axpy(fast_div_pos((double)c, p + 1e-100), row);
            if ((q % SPB) == SPB - 1) {  // batch end: rotate the index registers, fetch two batches ahead
                cur_w = nxt_w; cur_c = nxt_c; ++curbatch;
                const long nb = b + (long)(curbatch + 1) * 32;
                if (nb < b + per_warp) { nxt_w = idx[nb + lane]; nxt_c = cnts[nb + lane]; }
            }
        };
        KSlice<G, E> A, B, C;
        A.load(row_of(0), g, last_ok);
        if (NBUF == 3) {
            B.load(row_of(1), g, last_ok);
            for (int q = 0; q < Q; q += 3) {
                if (q + 2 < Q) C.load(row_of(q + 2), g, last_ok);
                body(q, A);
                if (q + 3 < Q) A.load(row_of(q + 3), g, last_ok);
                body(q + 1, B);
                if (q + 4 < Q) B.load(row_of(q + 4), g, last_ok);
                body(q + 2, C);
            }
        } else {
            for (int q = 0; q < Q; q += 2) {
                if (q + 1 < Q) B.load(row_of(q + 1), g, last_ok);
                body(q, A);
                if (q + 2 < Q) A.load(row_of(q + 2), g, last_ok);
                body(q + 1, B);
            }
        }
    }
#pragma unroll
    for (int j = 0; j < E; ++j) sink += acc.v[j].x + acc.v[j].y;
    if (sink == 12345.678) out[0] = sink;
}

template <int G, int E, int NBUF, bool TSMEM, int MINB>
void run(const char* name, const double* table, const int* idx, const int* cnts, long n, int ld, double* out)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    int grid = 148 * 16;
    const size_t smem = 4 * 2 * G * E * sizeof(double);
    for (int w = 0; w < 2; ++w) k<G, E, NBUF, TSMEM, MINB><<<grid, 128, smem>>>(table, idx, cnts, n, ld, 192, out);
    cudaEventRecord(e0);
    const int reps = 4;
    for (int w = 0; w < reps; ++w) k<G, E, NBUF, TSMEM, MINB><<<grid, 128, smem>>>(table, idx, cnts, n, ld, 192, out);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= reps;
    int nb = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k<G, E, NBUF, TSMEM, MINB>, 128, smem);
    cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, k<G, E, NBUF, TSMEM, MINB>);
    printf("%-22s nbuf=%d tsmem=%d minb=%d regs=%3d ctas/sm=%d : %8.3f ms  %8.1f GB/s   %s\n", name, NBUF, (int)TSMEM,
           MINB, fa.numRegs, nb, ms, (double)n * ld * 8 / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
}

int main(int argc, char** argv)
{
    const int ld = argc > 1 ? atoi(argv[1]) : 100;
    const long n = 192L * 148 * 16 * 4 * 22;  // ~40M rows
    double* out; cudaMalloc(&out, 8);
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
        double u = (double)(s >> 11) / 9007199254740992.0 * tot;
        long lo = 0, hi = rows - 1;
        while (lo < hi) { long m = (lo + hi) / 2; if (cdf[m] < u) lo = m + 1; else hi = m; }
        h[i] = (int)((lo * 7919L) % rows);
    }
    int *idx, *cnts; cudaMalloc(&idx, n * 4); cudaMalloc(&cnts, n * 4);
    cudaMemcpy(idx, h.data(), n * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(cnts, hc.data(), n * 4, cudaMemcpyHostToDevice);
    const char* name = "zipf rows, 40 MB";
    run<8, 7, 2, false, 3>(name, table, idx, cnts, n, ld, out);
    run<8, 7, 2, true, 3>(name, table, idx, cnts, n, ld, out);
    run<8, 7, 2, true, 4>(name, table, idx, cnts, n, ld, out);
    run<8, 7, 3, true, 3>(name, table, idx, cnts, n, ld, out);
    run<8, 7, 3, false, 2>(name, table, idx, cnts, n, ld, out);
    run<8, 7, 3, false, 3>(name, table, idx, cnts, n, ld, out);
    run<8, 7, 3, true, 2>(name, table, idx, cnts, n, ld, out);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
