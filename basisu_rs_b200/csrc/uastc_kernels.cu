// K1<target>: fused UASTC unpack + repack kernels.
// One thread per 16-byte block, 128-bit coalesced loads and stores, constant tables in shared
// memory.  Replaces uastc::Decoder::{transcode,decode_to_rgba} (reference src/uastc.rs:89-165).
#include <cstddef>
#include "uastc_device.cuh"
#include "kernels.h"

namespace b2bu {

constexpr int kMaxDevicesK = 16;
constexpr uint64_t kSortedMinBlocks = 2048;   // below this the sort cannot pay for itself

// tile shape of the mode-sorted kernel (overridable for tuning runs: -DB2BU_TILE=... etc.)
#ifndef B2BU_TILE
#define B2BU_TILE 2048
#endif
#ifndef B2BU_THREADS
#define B2BU_THREADS 512
#endif
#ifndef B2BU_CTAS_ASTC
#define B2BU_CTAS_ASTC 3
#endif
#ifndef B2BU_DYN_ASTC
#define B2BU_DYN_ASTC 0
#endif
#ifndef B2BU_CTAS_OTHER
#define B2BU_CTAS_OTHER 2
#endif

__device__ DevTables g_tables;
static const DevTables h_tables =
#include "device_tables_gen.inc"
    ;

cudaError_t upload_tables()
{
    return cudaMemcpyToSymbol(g_tables, &h_tables, sizeof(DevTables));
}

// DevTables is ordered front-end, ASTC, BC7, ETC: a kernel copies only the prefix its target reads.
template <int TARGET> struct TableBytes {
    static constexpr size_t raw = TARGET == TGT_ASTC ? offsetof(DevTables, bc7p2)
                                : TARGET == TGT_BC7 ? offsetof(DevTables, etc1_mod)
                                : TARGET == TGT_RGBA ? offsetof(DevTables, trit_enc) : sizeof(DevTables);
    static constexpr size_t value = (raw + 15) / 16 * 16;
};

__device__ __forceinline__ void load_tables(DevTables* dst, size_t bytes = sizeof(DevTables))
{
    static_assert(sizeof(DevTables) % 16 == 0, "DevTables must be a multiple of 16 bytes");
    const uint4* src = reinterpret_cast<const uint4*>(&g_tables);
    uint4* d = reinterpret_cast<uint4*>(dst);
    for (int i = threadIdx.x; i < (int)(bytes / 16); i += blockDim.x) d[i] = src[i];
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
// processing order of the bins: roughly by decreasing per-block cost (endpoint count, trits/quints, subsets)
__constant__ uint8_t kBinOrder[32] = {3, 9, 16, 4, 7, 2, 12, 10, 11, 13, 14, 6, 0, 18, 5, 1, 17, 15, 8, 19,
                                       31, 31, 31, 31, 31, 31, 31, 31, 31, 31, 31, 31};

// ---- TMA (1-D bulk async copy) + mbarrier helpers, sm_90+ PTX -------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_store_1d(void* gmem_dst, const void* smem_src, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(smem_u32(smem_src)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template <int TARGET> struct SortedCfg {
    static constexpr int TILE = B2BU_TILE;
    static constexpr int THREADS = B2BU_THREADS;
    static constexpr int PER = TILE / THREADS;
    static constexpr int OB = TARGET == TGT_RGBA ? 64 : TARGET == TGT_ETC1 ? 8 : 16;
    // ASTC / BC7 / ETC2 results overwrite the block's own 16-byte input slot (read once by the same
    // thread just before) and leave with one bulk store; ETC1 (8 B) is staged compactly in a second
    // buffer; RGBA (64 B) is written straight to global memory by the lane that decoded it.
    static constexpr bool IN_PLACE = OB == 16;
    static constexpr bool TMA_STORE = OB <= 16;
    static constexpr int CTAS_PER_SM = TARGET == TGT_ASTC ? B2BU_CTAS_ASTC : B2BU_CTAS_OTHER;
    // work-item scheduling inside a tile: ASTC items are short and even, a static round-robin beats the
    // shared counter; the heavier, more uneven targets (BC7, RGBA, ETC) gain from dynamic pulls
    static constexpr bool DYNAMIC = TARGET != TGT_ASTC || B2BU_DYN_ASTC;
    static constexpr int MAXORD = TILE + kBins * 32;
    static constexpr int MAXITEMS = MAXORD / 32;
    static constexpr size_t OFF_IN = (TableBytes<TARGET>::value + 127) / 128 * 128;   // two input buffers (double buffered)
    static constexpr size_t OFF_OUT = OFF_IN + 2 * (size_t)TILE * 16;       // ETC1 staging only
    static constexpr size_t OFF_ORDER = OFF_OUT + (TARGET == TGT_ETC1 ? (size_t)TILE * 8 : 0);
    static constexpr size_t OFF_IMODE = OFF_ORDER + (size_t)MAXORD * 2;
    static constexpr size_t OFF_CNT = (OFF_IMODE + MAXITEMS + 15) / 16 * 16;
    static constexpr size_t OFF_BAR = OFF_CNT + 4 * (32 + 32 + 4);
    static constexpr size_t SMEM = OFF_BAR + 16;
};

template <int TARGET>
__global__ void __launch_bounds__(SortedCfg<TARGET>::THREADS, SortedCfg<TARGET>::CTAS_PER_SM)
uastc_sorted_kernel(const uint4* __restrict__ in, void* __restrict__ out, uint64_t nblocks, uint32_t blocks_per_row,
                    uint64_t index_base, unsigned long long* __restrict__ err)
{
    using C = SortedCfg<TARGET>;
    extern __shared__ __align__(128) unsigned char smem[];
    DevTables& T = *reinterpret_cast<DevTables*>(smem);
    uint4* xbuf = reinterpret_cast<uint4*>(smem + C::OFF_IN);
    unsigned char* out_s = smem + C::OFF_OUT;
    uint16_t* order = reinterpret_cast<uint16_t*>(smem + C::OFF_ORDER);
    uint8_t* item_mode = smem + C::OFF_IMODE;
    uint32_t* cnt = reinterpret_cast<uint32_t*>(smem + C::OFF_CNT);
    uint32_t* offs = cnt + 32;
    uint32_t* ctl = offs + 32;                   // [0] next work item, [1] number of work items
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::OFF_BAR);

    const int tid = threadIdx.x, lane = tid & 31;
    const uint64_t ntiles = (nblocks + C::TILE - 1) / C::TILE;
    auto tile_blocks = [&](uint64_t t) -> uint32_t {
        const uint64_t rem = nblocks - t * C::TILE;
        return (uint32_t)(rem < (uint64_t)C::TILE ? rem : (uint64_t)C::TILE);
    };
    if (tid < 32) cnt[tid] = 0;
    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    load_tables(&T, TableBytes<TARGET>::value);  // ends with __syncthreads()
    if (tid == 0) {                              // prologue: first tile -> buffer 0
        const uint64_t t0 = blockIdx.x;
        const uint32_t bytes = tile_blocks(t0) * 16u;
        mbar_expect_tx(&bars[0], bytes);
        tma_load_1d(xbuf, in + t0 * C::TILE, bytes, &bars[0]);
    }

    uint32_t it = 0;
    for (uint64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, it++) {
        const uint32_t cur = it & 1u;
        uint4* in_s = xbuf + cur * C::TILE;
        const uint64_t base = tile * C::TILE;
        const uint32_t nt = tile_blocks(tile);
        const uint32_t bx0 = TARGET == TGT_RGBA ? (uint32_t)(base % blocks_per_row) : 0u;
        const uint64_t by0 = TARGET == TGT_RGBA ? base / blocks_per_row : 0u;

        // ---- A: wait for the tile, classify, rank inside each mode bin ----
        for (int i = tid; i < C::MAXORD / 8; i += C::THREADS) reinterpret_cast<uint4*>(order)[i] = make_uint4(~0u, ~0u, ~0u, ~0u);
        mbar_wait(&bars[cur], (it >> 1) & 1u);
        uint32_t mymode[C::PER], mypos[C::PER];
#pragma unroll
        for (int k = 0; k < C::PER; k++) {
            const uint32_t idx = tid + k * C::THREADS;
            mymode[k] = idx < nt ? (uint32_t)T.mode_lut[in_s[idx].x & 127u] : 31u;
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
            // bins are laid out heaviest mode first so that the dynamic pulls end with the cheap items
            const uint32_t bin = tid < kBins ? (uint32_t)kBinOrder[tid] : 31u;
            const uint32_t c = tid < kBins ? cnt[bin] : 0u;
            const uint32_t padded = (c + 31u) & ~31u;
            uint32_t incl = padded;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, d); if (lane >= d) incl += v; }
            const uint32_t excl = incl - padded;
            offs[bin] = excl;
            for (uint32_t j = excl >> 5; j < (incl >> 5); j++) item_mode[j] = (uint8_t)bin;
            cnt[bin] = 0;
            if (tid == 31) { ctl[0] = 0; ctl[1] = incl >> 5; }
        }
        __syncthreads();
        // ---- C: scatter block indices into their bins ----
#pragma unroll
        for (int k = 0; k < C::PER; k++)
            if (mymode[k] < (uint32_t)kBins) order[offs[mymode[k]] + mypos[k]] = (uint16_t)(tid + k * C::THREADS);
        // prefetch the next tile into the other buffer while this one is transcoded
        if (tid == 0 && tile + gridDim.x < ntiles) {
            if (C::TMA_STORE) tma_store_wait_read();     // the bulk store that last read that buffer has drained
            const uint64_t tn = tile + gridDim.x;
            const uint32_t bytes = tile_blocks(tn) * 16u;
            mbar_expect_tx(&bars[cur ^ 1u], bytes);
            tma_load_1d(xbuf + (cur ^ 1u) * C::TILE, in + tn * C::TILE, bytes, &bars[cur ^ 1u]);
        } else if (tid == 0 && TARGET == TGT_ETC1) {
            tma_store_wait_read();                        // ETC1 staging buffer is single: previous store must have read it
        }
        __syncthreads();
        // ---- D: warps take mode-uniform work items ----
        const uint32_t nitems = ctl[1];
        for (uint32_t wi = tid >> 5;; wi += C::THREADS / 32) {
            uint32_t item = wi;
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
                    // uastc.rs:96-106: row-major image, pitch 4*blocks_per_row pixels; four 16-byte row stores
                    const uint32_t t = bx0 + idx;                       // 32-bit: bx0 < blocks_per_row, idx < TILE
                    const uint64_t by = by0 + t / blocks_per_row;
                    const uint32_t bx = t % blocks_per_row;
                    uint4* p = reinterpret_cast<uint4*>(out) + (by * 4) * blocks_per_row + bx;
#pragma unroll
                    for (int y = 0; y < 4; y++) p[(uint64_t)y * blocks_per_row] = make_uint4(o.px[4 * y], o.px[4 * y + 1], o.px[4 * y + 2], o.px[4 * y + 3]);
                } else if (TARGET == TGT_ETC1) {
                    reinterpret_cast<uint2*>(out_s)[idx] = o.etc;
                } else {
                    in_s[idx] = o.v;
                }
            }
        }
        // ---- E: one bulk store per tile ----
        if (C::TMA_STORE) {
            fence_async_smem();                          // make the generic-proxy writes visible to the async proxy
            __syncthreads();
            if (tid == 0) {
                if (TARGET == TGT_ETC1) {
                    if ((nt & 1u) == 0u) tma_store_1d(reinterpret_cast<uint2*>(out) + base, out_s, nt * 8u);
                } else {
                    tma_store_1d(reinterpret_cast<uint4*>(out) + base, in_s, nt * 16u);
                }
            }
            if (TARGET == TGT_ETC1 && (nt & 1u)) {       // odd tail: bulk copies need multiples of 16 bytes
                uint2* dst = reinterpret_cast<uint2*>(out) + base;
                for (uint32_t i = tid; i < nt; i += C::THREADS) dst[i] = reinterpret_cast<const uint2*>(out_s)[i];
            }
        } else {
            __syncthreads();                             // RGBA: in_s / order are reused by the next tile
        }
    }
    if (C::TMA_STORE && tid == 0) tma_store_wait_all();
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
