// K1<target>: fused UASTC unpack + repack kernels.
// One thread per 16-byte block, 128-bit coalesced loads and stores, constant tables in shared
// memory.  Replaces uastc::Decoder::{transcode,decode_to_rgba} (reference src/uastc.rs:89-165).
#include "uastc_device.cuh"
#include "kernels.h"

namespace b2bu {

__device__ DevTables g_tables;
static const DevTables h_tables =
#include "device_tables_gen.inc"
    ;

cudaError_t upload_tables()
{
    return cudaMemcpyToSymbol(g_tables, &h_tables, sizeof(DevTables));
}

__device__ __forceinline__ void load_tables(DevTables* dst)
{
    static_assert(sizeof(DevTables) % 16 == 0, "DevTables must be a multiple of 16 bytes");
    const uint4* src = reinterpret_cast<const uint4*>(&g_tables);
    uint4* d = reinterpret_cast<uint4*>(dst);
    for (int i = threadIdx.x; i < (int)(sizeof(DevTables) / 16); i += blockDim.x) d[i] = src[i];
    __syncthreads();
}

__device__ __forceinline__ void report_error(unsigned long long* err, uint64_t block_index, uint32_t code)
{
    // first failing block wins (uastc.rs:161-163: the first Err aborts the slice)
    atomicMin(err, (unsigned long long)((block_index << 8) | code));
}

template <int TARGET>
__global__ void __launch_bounds__(256) uastc_transcode_kernel(const uint4* __restrict__ in, void* __restrict__ out,
                                                              uint64_t nblocks, uint32_t blocks_per_row, uint64_t index_base,
                                                              unsigned long long* __restrict__ err)
{
    __shared__ DevTables T;
    load_tables(&T);
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nblocks; i += stride) {
        const uint4 b = __ldg(in + i);
        BlockOut o;
        const uint32_t e = transcode_one<TARGET>(b, T, o);
        if (e != ERR_OK) {
            report_error(err, index_base + i, e);
            o.v = make_uint4(0u, 0u, 0u, 0u); o.etc = make_uint2(0u, 0u);
#pragma unroll
            for (int k = 0; k < 16; k++) o.px[k] = 0u;
        }
        if (TARGET == TGT_RGBA) {
            // uastc.rs:96-106: row-major image, pitch 4*blocks_per_row pixels
            const uint64_t bx = i % blocks_per_row, by = i / blocks_per_row;
            uint4* dst = reinterpret_cast<uint4*>(out) + (by * 4) * blocks_per_row + bx;
#pragma unroll
            for (int y = 0; y < 4; y++) dst[(uint64_t)y * blocks_per_row] = make_uint4(o.px[4 * y], o.px[4 * y + 1], o.px[4 * y + 2], o.px[4 * y + 3]);
        } else if (TARGET == TGT_ETC1) {
            reinterpret_cast<uint2*>(out)[i] = o.etc;
        } else {
            reinterpret_cast<uint4*>(out)[i] = o.v;
        }
    }
}

cudaError_t launch_uastc_transcode(int target, const void* d_in, void* d_out, uint64_t nblocks, uint32_t blocks_per_row,
                                   uint64_t index_base, unsigned long long* d_err, int sm_count, cudaStream_t stream)
{
    if (nblocks == 0) return cudaSuccess;
    const int threads = 256;
    uint64_t want = (nblocks + threads - 1) / threads;
    const uint64_t cap = (uint64_t)sm_count * 8;
    const unsigned grid = (unsigned)(want < cap ? want : cap);
    const uint4* in = reinterpret_cast<const uint4*>(d_in);
    switch (target) {
    case TGT_RGBA: uastc_transcode_kernel<TGT_RGBA><<<grid, threads, 0, stream>>>(in, d_out, nblocks, blocks_per_row, index_base, d_err); break;
    case TGT_ASTC: uastc_transcode_kernel<TGT_ASTC><<<grid, threads, 0, stream>>>(in, d_out, nblocks, blocks_per_row, index_base, d_err); break;
    case TGT_BC7:  uastc_transcode_kernel<TGT_BC7><<<grid, threads, 0, stream>>>(in, d_out, nblocks, blocks_per_row, index_base, d_err); break;
    case TGT_ETC1: uastc_transcode_kernel<TGT_ETC1><<<grid, threads, 0, stream>>>(in, d_out, nblocks, blocks_per_row, index_base, d_err); break;
    case TGT_ETC2: uastc_transcode_kernel<TGT_ETC2><<<grid, threads, 0, stream>>>(in, d_out, nblocks, blocks_per_row, index_base, d_err); break;
    default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

}  // namespace b2bu
