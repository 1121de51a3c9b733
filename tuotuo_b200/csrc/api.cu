// api.cu — library plumbing: version, error reporting, device check, record reduction, layout transposes.
#include <cstdarg>
#include <cstdio>

#include "ttlda_common.cuh"
#include "tma_gather.cuh"

namespace ttlda {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what)
{
    set_error("%s: CUDA error %d (%s)", what, (int)e, cudaGetErrorString(e));
    return TTLDA_ECUDA;
}

__global__ void reduce_records_kernel(const double* __restrict__ rec, int nrec, double* __restrict__ out)
{
    __shared__ double sm[NPART * 32];
    double v[NPART];
#pragma unroll
    for (int j = 0; j < NPART; ++j) v[j] = 0.;
    for (int b = threadIdx.x; b < nrec; b += blockDim.x) {
#pragma unroll
        for (int j = 0; j < NPART; ++j) v[j] += rec[(size_t)b * NPART + j];
    }
    block_sum<NPART>(v, sm);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int j = 0; j < NPART; ++j) out[j] = v[j];
    }
}

int launch_reduce_records(const double* rec, int nrec, double* out, cudaStream_t stream)
{
    reduce_records_kernel<<<1, 256, 0, stream>>>(rec, nrec, out);
    TTLDA_LAUNCH_CHECK("reduce_records");
    return TTLDA_OK;
}

// dst[c][r] = src[r][c], src [rows][src_ld], dst [cols][dst_ld]; 32x32 tiles through shared memory.
__global__ void transpose_kernel(const double* __restrict__ src, int64_t rows, int64_t cols, int64_t src_ld,
                                 double* __restrict__ dst, int64_t dst_ld)
{
    __shared__ double tile[32][33];
    const int64_t r0 = (int64_t)blockIdx.y * 32, c0 = (int64_t)blockIdx.x * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        int64_t r = r0 + i, c = c0 + threadIdx.x;
        tile[i][threadIdx.x] = (r < rows && c < cols) ? src[r * src_ld + c] : 0.;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        int64_t c = c0 + i, r = r0 + threadIdx.x;  // dst row = c, dst col = r
        if (c < cols && r < dst_ld) dst[c * dst_ld + r] = tile[threadIdx.x][i];  // r >= rows reads zeros (padding)
    }
}

// Tensor map of a [rows][ld] fp64 table for tile::gather4 loads: 2-D, box {ld, 1} (one row per gathered index), no
// swizzle / interleave.  The driver's encoder is reached through the runtime (no libcuda link dependency).
int make_row_gather_map(const double* table, int64_t rows, int64_t ld, CUtensorMap* out)
{
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                 const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                 CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn encode = nullptr;
    if (!encode) {
        cudaDriverEntryPointQueryResult q;
        void* fn = nullptr;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
        if (e != cudaSuccess || !fn) {
            set_error("cuTensorMapEncodeTiled not available (%s)", cudaGetErrorString(e));
            return TTLDA_ECUDA;
        }
        encode = (EncodeFn)fn;
    }
    const cuuint64_t gdim[2] = {(cuuint64_t)ld, (cuuint64_t)(rows > 0 ? rows : 1)};
    const cuuint64_t gstride[1] = {(cuuint64_t)ld * 8};
    const cuuint32_t box[2] = {(cuuint32_t)ld, 1};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(table), gdim, gstride, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d) for a [%lld][%lld] table", (int)r, (long long)rows, (long long)ld);
        return TTLDA_ECUDA;
    }
    return TTLDA_OK;
}

}  // namespace ttlda

using namespace ttlda;

extern "C" {

int ttlda_version(void) { return (0 << 16) | (1 << 8) | 0; }

const char* ttlda_last_error(void) { return g_err; }

int ttlda_check_device(int device)
{
    cudaDeviceProp p;
    cudaError_t e = cudaGetDeviceProperties(&p, device);
    if (e != cudaSuccess) return cuda_fail(e, "cudaGetDeviceProperties");
    if (p.major != 10) {
        set_error("device %d is sm_%d%d; libttlda is built for sm_100a only", device, p.major, p.minor);
        return TTLDA_EARCH;
    }
    return TTLDA_OK;
}

int64_t ttlda_ld_for(int64_t K) { return (K + 3) / 4 * 4; }

int64_t ttlda_reduce_workspace_bytes(void) { return reduce_ws_bytes(); }

int ttlda_transpose_pad_f64(const double* src, int64_t rows, int64_t cols, double* dst, int64_t dst_ld, void* stream)
{
    TTLDA_REQUIRE(src && dst, "NULL pointer");
    TTLDA_REQUIRE(rows >= 0 && cols >= 0 && dst_ld >= rows, "bad shape");
    if (rows == 0 || cols == 0) return TTLDA_OK;
    dim3 grid((unsigned)((cols + 31) / 32), (unsigned)((dst_ld + 31) / 32)), block(32, 8);
    transpose_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(src, rows, cols, cols, dst, dst_ld);
    TTLDA_LAUNCH_CHECK("transpose_pad");
    return TTLDA_OK;
}

int ttlda_transpose_unpad_f64(const double* src, int64_t rows, int64_t cols, int64_t src_ld, double* dst, void* stream)
{
    TTLDA_REQUIRE(src && dst, "NULL pointer");
    TTLDA_REQUIRE(rows >= 0 && cols >= 0 && src_ld >= cols, "bad shape");
    if (rows == 0 || cols == 0) return TTLDA_OK;
    dim3 grid((unsigned)((cols + 31) / 32), (unsigned)((rows + 31) / 32)), block(32, 8);
    transpose_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(src, rows, cols, src_ld, dst, rows);
    TTLDA_LAUNCH_CHECK("transpose_unpad");
    return TTLDA_OK;
}

}  // extern "C"
