// K1<target>: fused UASTC unpack + repack kernels.
// One thread per 16-byte block, 128-bit coalesced loads and stores, constant tables in shared
// memory.  Replaces uastc::Decoder::{transcode,decode_to_rgba} (reference src/uastc.rs:89-165).
#include "uastc_device.cuh"
#include "kernels.h"

namespace b2bu {

constexpr int kMaxDevicesK = 16;
constexpr uint64_t kSortedMinBlocks = 2048;   // below this the sort cannot pay for itself

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

// shared-memory atomic add with the address space made explicit (the generic form costs ~10 instructions)
__device__ __forceinline__ uint32_t atom_add_shared(uint32_t* p, uint32_t v)
{
    uint32_t old;
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
    asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(a), "r"(v) : "memory");
    return old;
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

// ------------------------------------------------------------------------------------------
// Mode-sorted tile kernel.  A warp that holds 32 consecutive blocks of a real texture sees many
// different UASTC modes, and the mode-specialised code above would then run one mode at a time
// with most lanes idle (measured: 2.4 of 32 lanes active on a shuffled payload).  So a CTA takes a
// tile of TILE consecutive blocks, counting-sorts the block indices by mode in shared memory
// (warp-aggregated with match.any), and its warps then pull 32-block work items that are
// mode-uniform.  Results go back to the block's original slot in a shared staging buffer and
// leave with fully coalesced 128-bit stores.  Bins are padded to 32, so lane occupancy is
// TILE / (TILE + ~16 per mode present).
// ------------------------------------------------------------------------------------------
constexpr int kBins = 20;                       // modes 0..18 + the invalid code (19)

template <int TARGET> struct SortedCfg {
    static constexpr int TILE = TARGET == TGT_RGBA ? 1024 : 2048;
    static constexpr int THREADS = 512;
    static constexpr int PER = TILE / THREADS;
    static constexpr int OB = TARGET == TGT_RGBA ? 64 : TARGET == TGT_ETC1 ? 8 : 16;
    static constexpr int MAXORD = TILE + kBins * 32;
    static constexpr int MAXITEMS = MAXORD / 32;
    // ASTC / BC7 / ETC1 / ETC2 results overwrite the block's own 16-byte input slot (read once by the
    // same thread just before); only RGBA (64 B per block) needs a separate staging buffer
    static constexpr bool IN_PLACE = OB <= 16;
    static constexpr int CTAS_PER_SM = TARGET == TGT_ASTC ? 3 : 2;
    // work-item scheduling inside a tile: ASTC items are short and even, a static round-robin beats the
    // shared counter; the heavier, more uneven targets (BC7, RGBA, ETC) gain from dynamic pulls
    static constexpr bool DYNAMIC = TARGET != TGT_ASTC;
    static constexpr size_t OFF_IN = sizeof(DevTables);
    static constexpr size_t OFF_OUT = IN_PLACE ? OFF_IN : OFF_IN + (size_t)TILE * 16;
    static constexpr size_t OFF_ORDER = OFF_OUT + (size_t)TILE * (IN_PLACE ? 16 : OB);
    static constexpr size_t OFF_IMODE = OFF_ORDER + (size_t)MAXORD * 2;
    static constexpr size_t OFF_CNT = (OFF_IMODE + MAXITEMS + 15) / 16 * 16;
    static constexpr size_t SMEM = OFF_CNT + 4 * (32 + 32 + 4);
};

template <int TARGET>
__global__ void __launch_bounds__(SortedCfg<TARGET>::THREADS, SortedCfg<TARGET>::CTAS_PER_SM)
uastc_sorted_kernel(const uint4* __restrict__ in, void* __restrict__ out, uint64_t nblocks, uint32_t blocks_per_row,
                    uint64_t index_base, unsigned long long* __restrict__ err)
{
    using C = SortedCfg<TARGET>;
    extern __shared__ __align__(16) unsigned char smem[];
    DevTables& T = *reinterpret_cast<DevTables*>(smem);
    uint4* in_s = reinterpret_cast<uint4*>(smem + C::OFF_IN);
    unsigned char* out_s = smem + C::OFF_OUT;
    uint16_t* order = reinterpret_cast<uint16_t*>(smem + C::OFF_ORDER);
    uint8_t* item_mode = smem + C::OFF_IMODE;
    uint32_t* cnt = reinterpret_cast<uint32_t*>(smem + C::OFF_CNT);
    uint32_t* offs = cnt + 32;
    uint32_t* ctl = offs + 32;                   // [0] next work item, [1] number of work items

    const int tid = threadIdx.x, lane = tid & 31;
    if (tid < 32) cnt[tid] = 0;
    load_tables(&T);                             // ends with __syncthreads()

    const uint64_t ntiles = (nblocks + C::TILE - 1) / C::TILE;
    for (uint64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const uint64_t base = tile * C::TILE;
        const uint32_t nt = (uint32_t)(nblocks - base < (uint64_t)C::TILE ? nblocks - base : (uint64_t)C::TILE);

        // ---- A: load the tile, classify, rank inside each mode bin ----
        for (int i = tid; i < C::MAXORD; i += C::THREADS) order[i] = 0xFFFFu;
        uint32_t mymode[C::PER], mypos[C::PER];
        uint4 blk[C::PER];
#pragma unroll
        for (int k = 0; k < C::PER; k++) {                 // all global loads in flight before any use
            const uint32_t idx = tid + k * C::THREADS;
            blk[k] = idx < nt ? __ldg(in + base + idx) : make_uint4(0u, 0u, 0u, 0u);
        }
#pragma unroll
        for (int k = 0; k < C::PER; k++) {
            const uint32_t idx = tid + k * C::THREADS;
            in_s[idx] = blk[k];
            mymode[k] = idx < nt ? (uint32_t)T.mode_lut[blk[k].x & 127u] : 31u;
        }
#pragma unroll
        for (int k = 0; k < C::PER; k++) {
            const uint32_t m = mymode[k];
            const uint32_t peers = __match_any_sync(0xFFFFFFFFu, m);
            const int leader = __ffs(peers) - 1;
            uint32_t p0 = 0;
            if (lane == leader && m < (uint32_t)kBins) p0 = atom_add_shared(&cnt[m], __popc(peers));
            p0 = __shfl_sync(0xFFFFFFFFu, p0, leader);
            mypos[k] = p0 + __popc(peers & ((1u << lane) - 1u));
        }
        __syncthreads();
        // ---- B: bin offsets (each bin padded to a multiple of 32) and the item -> mode map ----
        if (tid < 32) {
            const uint32_t c = tid < kBins ? cnt[tid] : 0u;
            const uint32_t padded = (c + 31u) & ~31u;
            uint32_t incl = padded;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, d); if (lane >= d) incl += v; }
            const uint32_t excl = incl - padded;
            offs[tid] = excl;
            for (uint32_t j = excl >> 5; j < (incl >> 5); j++) item_mode[j] = (uint8_t)tid;
            cnt[tid] = 0;
            if (tid == 31) { ctl[0] = 0; ctl[1] = incl >> 5; }
        }
        __syncthreads();
        // ---- C: scatter block indices into their bins ----
#pragma unroll
        for (int k = 0; k < C::PER; k++)
            if (mymode[k] < (uint32_t)kBins) order[offs[mymode[k]] + mypos[k]] = (uint16_t)(tid + k * C::THREADS);
        __syncthreads();
        // ---- D: warps pull mode-uniform work items ----
        const uint32_t nitems = ctl[1];
        for (uint32_t it = tid >> 5;; it += C::THREADS / 32) {
            uint32_t item = it;
            if (C::DYNAMIC) {
                if (lane == 0) item = atom_add_shared(&ctl[0], 1u);
                item = __shfl_sync(0xFFFFFFFFu, item, 0);
            }
            if (item >= nitems) break;
            const uint32_t mode = item_mode[item];
            const uint32_t idx = order[item * 32 + lane];
            if (idx != 0xFFFFu) {
                const uint4 b = in_s[idx];
                BlockOut o;
                const uint32_t e = transcode_mode<TARGET>(mode, b, T, o);
                if (e != ERR_OK) {
                    report_error(err, index_base + base + idx, e);
                    o.v = make_uint4(0u, 0u, 0u, 0u); o.etc = make_uint2(0u, 0u);
#pragma unroll
                    for (int k = 0; k < 16; k++) o.px[k] = 0u;
                }
                if (TARGET == TGT_RGBA) {
#pragma unroll
                    for (int y = 0; y < 4; y++)
                        reinterpret_cast<uint4*>(out_s)[y * C::TILE + idx] = make_uint4(o.px[4 * y], o.px[4 * y + 1], o.px[4 * y + 2], o.px[4 * y + 3]);
                } else if (TARGET == TGT_ETC1) {
                    reinterpret_cast<uint2*>(out_s)[idx * 2] = o.etc;          // first half of the block's own slot
                } else {
                    reinterpret_cast<uint4*>(out_s)[idx] = o.v;
                }
            }
        }
        __syncthreads();
        // ---- E: coalesced stores ----
        if (TARGET == TGT_RGBA) {
            // uastc.rs:96-106: row-major image, pitch 4*blocks_per_row pixels
            const uint32_t bx0 = (uint32_t)(base % blocks_per_row);
            const uint64_t by0 = base / blocks_per_row;
            uint4* dst = reinterpret_cast<uint4*>(out);
            for (uint32_t i = tid; i < nt; i += C::THREADS) {
                const uint32_t t = bx0 + i;
                const uint64_t by = by0 + t / blocks_per_row;
                const uint32_t bx = t % blocks_per_row;
                uint4* p = dst + (by * 4) * blocks_per_row + bx;
#pragma unroll
                for (int y = 0; y < 4; y++) p[(uint64_t)y * blocks_per_row] = reinterpret_cast<const uint4*>(out_s)[y * C::TILE + i];
            }
        } else if (TARGET == TGT_ETC1) {
            uint2* dst = reinterpret_cast<uint2*>(out) + base;
            for (uint32_t i = tid; i < nt; i += C::THREADS) dst[i] = reinterpret_cast<const uint2*>(out_s)[i * 2];
        } else {
            uint4* dst = reinterpret_cast<uint4*>(out) + base;
            for (uint32_t i = tid; i < nt; i += C::THREADS) dst[i] = reinterpret_cast<const uint4*>(out_s)[i];
        }
        // the next tile's phase A overwrites in_s (== out_s for the in-place targets) and order
        if (C::IN_PLACE) __syncthreads();
    }
}

template <int TARGET>
static cudaError_t launch_sorted(const uint4* in, void* d_out, uint64_t nblocks, uint32_t bpr, uint64_t index_base,
                                 unsigned long long* d_err, int sm_count, cudaStream_t stream)
{
    using C = SortedCfg<TARGET>;
    static bool configured[kMaxDevicesK] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < kMaxDevicesK && !configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(uastc_sorted_kernel<TARGET>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
        if (e != cudaSuccess) return e;
        configured[dev] = true;
    }
    const uint64_t ntiles = (nblocks + C::TILE - 1) / C::TILE;
    const uint64_t cap = (uint64_t)sm_count * C::CTAS_PER_SM;
    const unsigned grid = (unsigned)(ntiles < cap ? ntiles : cap);
    uastc_sorted_kernel<TARGET><<<grid, C::THREADS, C::SMEM, stream>>>(in, d_out, nblocks, bpr, index_base, d_err);
    return cudaGetLastError();
}

cudaError_t launch_uastc_transcode(int target, const void* d_in, void* d_out, uint64_t nblocks, uint32_t blocks_per_row,
                                   uint64_t index_base, unsigned long long* d_err, int sm_count, cudaStream_t stream)
{
    if (nblocks == 0) return cudaSuccess;
    const uint4* in = reinterpret_cast<const uint4*>(d_in);
    if (nblocks >= kSortedMinBlocks) {
        switch (target) {
        case TGT_RGBA: return launch_sorted<TGT_RGBA>(in, d_out, nblocks, blocks_per_row, index_base, d_err, sm_count, stream);
        case TGT_ASTC: return launch_sorted<TGT_ASTC>(in, d_out, nblocks, blocks_per_row, index_base, d_err, sm_count, stream);
        case TGT_BC7:  return launch_sorted<TGT_BC7>(in, d_out, nblocks, blocks_per_row, index_base, d_err, sm_count, stream);
        case TGT_ETC1: return launch_sorted<TGT_ETC1>(in, d_out, nblocks, blocks_per_row, index_base, d_err, sm_count, stream);
        case TGT_ETC2: return launch_sorted<TGT_ETC2>(in, d_out, nblocks, blocks_per_row, index_base, d_err, sm_count, stream);
        default: return cudaErrorInvalidValue;
        }
    }
    // small inputs (single blocks, the tail mips of a chain): plain one-thread-per-block kernel
    const int threads = 256;
    uint64_t want = (nblocks + threads - 1) / threads;
    const uint64_t cap = (uint64_t)sm_count * 8;
    const unsigned grid = (unsigned)(want < cap ? want : cap);
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
