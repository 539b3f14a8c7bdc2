// K1<target>: fused UASTC unpack + repack kernels.
// One thread per 16-byte block, 128-bit coalesced loads and stores, constant tables in shared
// memory.  Replaces uastc::Decoder::{transcode,decode_to_rgba} (reference src/uastc.rs:89-165).
#include <cstddef>
#include "uastc_device.cuh"
#include "ptx_helpers.cuh"
#include "kernels.h"

namespace b2bu {

constexpr int kMaxDevicesK = 16;
#ifndef B2BU_SORT_MIN
#define B2BU_SORT_MIN 2048
#endif
constexpr uint64_t kSortedMinBlocks = B2BU_SORT_MIN;   // below this the sort cannot pay for itself

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
    asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(a), "r"(v));
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
        // uastc.rs:96-106: row-major image, pitch 4*blocks_per_row pixels
        const uint64_t bx = i % blocks_per_row, by = i / blocks_per_row;
        StridedRowSink sink{reinterpret_cast<uint4*>(out) + (by * 4) * blocks_per_row + bx, (uint64_t)blocks_per_row};
        const uint32_t e = transcode_mode_sink<TARGET>(T.mode_lut[b.x & 127u], b, T, o, sink);
        if (e != ERR_OK) {
            report_error(err, index_base + i, e);
            o.v = make_uint4(0u, 0u, 0u, 0u); o.etc = make_uint2(0u, 0u);
            if (TARGET == TGT_RGBA) {
#pragma unroll 1
                for (int y = 0; y < 4; y++) sink.row(y, make_uint4(0u, 0u, 0u, 0u));
            }
        }
        if (TARGET == TGT_ETC1) reinterpret_cast<uint2*>(out)[i] = o.etc;
        else if (TARGET != TGT_RGBA) reinterpret_cast<uint4*>(out)[i] = o.v;
    }
}

// ------------------------------------------------------------------------------------------
// Mode-sorted, warp-specialised, persistent tile pipeline (one CTA per SM).
//
// Mode-specialised code diverges 19 ways inside a warp on real data and its instruction footprint
// (46-190 KB) thrashes the 32 KB L1.5 instruction cache when every warp of an SM runs a different
// mode.  So each CTA walks a contiguous range of blocks in tiles and runs three roles concurrently:
//
//   DMA warp (1 lane)   TMA bulk load of tile k+1 into the free slot, bulk store of finished tiles
//   sorter warps        classify tile k+1 by UASTC mode, counting-sort the block indices into
//                       mode bins padded to 32 (per-warp shared histograms, no CTA barrier)
//   worker warps        pull 32-block, mode-uniform work items of tile k in bin order (heavy modes
//                       first) and run the specialised transcode.  All workers of the SM are inside
//                       the same few modes at any time, which keeps the hot code I-cache resident.
//
// The roles are connected by mbarriers per slot (full -> sorted -> done -> free); there is no
// __syncthreads in the steady state and workers run on from tile k into tile k+1 without waiting
// for each other.  16-byte results overwrite the block's own input slot in shared memory and leave
// with one bulk store per tile; ETC1 (8 B) and RGBA (64 B, stored as four pixel rows per block row)
// are staged in a second buffer so that every global write is a full-line bulk store.
// ------------------------------------------------------------------------------------------
constexpr int kBins = 20;                       // modes 0..18 + the invalid code (19)
// processing order of the bins: roughly by decreasing per-block cost (endpoint count, trits/quints, subsets)
__constant__ uint8_t kBinOrder[32] = {3, 9, 16, 4, 7, 2, 12, 10, 11, 13, 14, 6, 0, 18, 5, 1, 17, 15, 8, 19,
                                       31, 31, 31, 31, 31, 31, 31, 31, 31, 31, 31, 31};

#ifndef B2BU_TILE16
#define B2BU_TILE16 4096
#endif
#ifndef B2BU_TILE_RGBA
#define B2BU_TILE_RGBA 1536     // 64 B of shared memory per block and slot; 1024 -> 1344 -> 1536 blocks: RGBA 152 -> 131 -> 122 us (1664: 124)
#endif
#ifndef B2BU_TILE_ETC1
#define B2BU_TILE_ETC1 B2BU_TILE16
#endif
// RGBA output straight from the worker threads to the image (four 16-byte row segments per block, scattered; the L2
// merges the half-written 32-byte sectors) instead of through a [4][TILE] staging buffer and bulk stores: no 64 B per
// block of staging, so the RGBA tile can be as large as the others.
#ifndef B2BU_RGBA_DIRECT
#define B2BU_RGBA_DIRECT 0
#endif
#ifndef B2BU_TILE_RGBA_DIRECT
#define B2BU_TILE_RGBA_DIRECT 4096
#endif
#ifndef B2BU_SORT_WARPS
#define B2BU_SORT_WARPS 8
#endif
#ifndef B2BU_WORK_WARPS
#define B2BU_WORK_WARPS 23
#endif

#ifdef B2BU_TRACE
// tuning aid (never in the product build): per-CTA clock64 time stamps of the pipeline hand-offs
__device__ unsigned long long g_trace[160][64];
#define TRACE(slot) do { if ((slot) < 64) g_trace[blockIdx.x][(slot)] = clock64(); } while (0)
#define TRACE_ADD(slot, v) do { g_trace[blockIdx.x][(slot)] += (v); } while (0)
#else
#define TRACE(slot) do { } while (0)
#define TRACE_ADD(slot, v) do { } while (0)
#endif

// the DMA lane's waits for a finished tile: sleeping polls (see mbar_wait_backoff); B2BU_DMA_SLEEP_NS = 0 spins
#ifndef B2BU_DMA_SLEEP_NS
#define B2BU_DMA_SLEEP_NS 200
#endif
#if B2BU_DMA_SLEEP_NS > 0
#define B2BU_DMA_WAIT(bar, parity) mbar_wait_backoff((bar), (parity), B2BU_DMA_SLEEP_NS)
#else
#define B2BU_DMA_WAIT(bar, parity) mbar_wait((bar), (parity))
#endif

#ifndef B2BU_SORT_UNIFORM
#define B2BU_SORT_UNIFORM 0     // measured: ASTC shuffled 55 -> 62 us, coherent 64 -> 62 us; off
#endif
// With B2BU_SUBLOAD a tile is loaded as several bulk copies with a barrier each, so that the sorter classifies the first
// part while the rest is still arriving (the load of a 64 KiB tile takes ~3000 cycles when all SMs load at once).  Measured:
// ASTC 56 -> 58 us -- like every other attempt to overlap the stages more, it loses: the SM is short of issue slots, not of
// overlap.  Off.
#ifndef B2BU_SUBLOAD
#define B2BU_SUBLOAD 0
#endif
#ifndef B2BU_STATIC_BC7
#define B2BU_STATIC_BC7 0
#endif
#ifndef B2BU_SORT_SPREAD
#define B2BU_SORT_SPREAD 1
#endif
#ifndef B2BU_CTAS_PER_SM
#define B2BU_CTAS_PER_SM 1      // persistent CTAs per SM; 2 x 512 threads with half-size tiles measured 72 us (ASTC) against 55
#endif
#ifndef B2BU_STAGGER
#define B2BU_STAGGER 0
#endif
#ifndef B2BU_SLOTS
#define B2BU_SLOTS 2
#endif
#ifndef B2BU_SLOTS16
#define B2BU_SLOTS16 B2BU_SLOTS
#endif

template <int TARGET> struct PipeCfg {
    static constexpr int OB = TARGET == TGT_RGBA ? 64 : TARGET == TGT_ETC1 ? 8 : 16;
    static constexpr bool IN_PLACE = OB == 16;
    static constexpr bool DIRECT = TARGET == TGT_RGBA && B2BU_RGBA_DIRECT;     // no staged output at all
    // tile slots in flight (load / sort / work / store are four stages: a slot is busy through all of them)
    static constexpr int NS = IN_PLACE ? B2BU_SLOTS16 : B2BU_SLOTS;
    static constexpr bool DYNAMIC = B2BU_STATIC_BC7 ? (TARGET != TGT_ASTC && TARGET != TGT_BC7) : TARGET != TGT_ASTC;
    static constexpr int TILE = TARGET == TGT_RGBA ? (DIRECT ? B2BU_TILE_RGBA_DIRECT : B2BU_TILE_RGBA) : TARGET == TGT_ETC1 ? B2BU_TILE_ETC1 : B2BU_TILE16;
    static constexpr int SORT_WARPS = B2BU_SORT_WARPS;
    static constexpr int SORT_THREADS = SORT_WARPS * 32;
    static constexpr int WORK_WARPS = B2BU_WORK_WARPS;
    static constexpr int THREADS = 32 * (1 + SORT_WARPS + WORK_WARPS);
    static constexpr int PERS = (TILE + SORT_THREADS - 1) / SORT_THREADS;     // blocks per sorter thread
    static constexpr int SUB = !B2BU_SUBLOAD ? TILE : (TILE % 1024 == 0 ? 1024 : TILE % 512 == 0 ? 512 : TILE);   // blocks per bulk load
    static constexpr int NSUB = TILE / SUB;
    static_assert(SUB % SORT_THREADS == 0 || NSUB == 1, "a sub-load must be whole sorter passes");
    static constexpr int MAXORD = TILE + kBins * 32;
    static constexpr int MAXITEMS = MAXORD / 32;
    static constexpr size_t OFF_IN = (TableBytes<TARGET>::value + 127) / 128 * 128;   // two slots
    static constexpr size_t OFF_OUT = OFF_IN + NS * (size_t)TILE * 16;                 // staging for ETC1 / RGBA, one per slot
    // RGBA stages 48 B per block: pixel row 0 goes where the block's input was (TileRowSink)
    static constexpr size_t OUT_SLOT = (IN_PLACE || DIRECT) ? 0 : TARGET == TGT_RGBA ? (size_t)TILE * 48 : (size_t)TILE * OB;
    static constexpr size_t OFF_ORDER = OFF_OUT + NS * OUT_SLOT;
    static constexpr size_t OFF_INFO = OFF_ORDER + NS * (size_t)MAXORD * 2;
    static constexpr size_t OFF_WCNT = (OFF_INFO + NS * 32 * 4 + 15) / 16 * 16;
    static constexpr size_t OFF_BASE = OFF_WCNT + (size_t)SORT_WARPS * 32 * 4;
    static constexpr size_t OFF_CTL = OFF_BASE + (size_t)SORT_WARPS * 32 * 4;
    static constexpr size_t OFF_BAR = (OFF_CTL + NS * 4 * 4 + 7) / 8 * 8;
    static constexpr size_t SMEM = OFF_BAR + (NSUB + 2) * NS * 8;
    static_assert(SMEM <= 227 * 1024, "tile configuration does not fit shared memory");
    static_assert(THREADS <= 1024, "too many warps");
};

// contiguous, 32-block aligned share of CTA c out of G
// (quot, rem) = (nblocks / G, nblocks % G) come from the host
__device__ __forceinline__ uint64_t cta_range_start(uint64_t nblocks, uint64_t quot, uint32_t rem, uint32_t c, uint32_t G)
{
    if (c >= G) return nblocks;
    return (quot * c + rem * c / G) & ~31ull;
}

template <int TARGET>
__global__ void __launch_bounds__(PipeCfg<TARGET>::THREADS, B2BU_CTAS_PER_SM)
uastc_sorted_kernel(const uint4* __restrict__ in, void* __restrict__ out, uint64_t nblocks, uint32_t blocks_per_row,
                    uint64_t index_base, unsigned long long* __restrict__ err, uint64_t range_quot, uint32_t range_rem)
{
    using C = PipeCfg<TARGET>;
    extern __shared__ __align__(128) unsigned char smem[];
    DevTables& T = *reinterpret_cast<DevTables*>(smem);
    uint4* in_s = reinterpret_cast<uint4*>(smem + C::OFF_IN);                 // [2][TILE]
    unsigned char* out_s = smem + C::OFF_OUT;                                 // [2][OUT_SLOT]
    uint16_t* order = reinterpret_cast<uint16_t*>(smem + C::OFF_ORDER);       // [2][MAXORD]
    uint32_t* bintab = reinterpret_cast<uint32_t*>(smem + C::OFF_INFO);       // [2][32], in bin order: first item | blocks in the bin << 16
    uint32_t* wcnt = reinterpret_cast<uint32_t*>(smem + C::OFF_WCNT);         // [SORT_WARPS][32]
    uint32_t* wbase = reinterpret_cast<uint32_t*>(smem + C::OFF_BASE);        // [SORT_WARPS][32]
    uint32_t* ctl = reinterpret_cast<uint32_t*>(smem + C::OFF_CTL);           // [2][4]: next item, number of items
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::OFF_BAR);          // full[2], sorted[2], done[2]
    uint64_t* bar_full = bars, *bar_sorted = bars + C::NS * C::NSUB, *bar_done = bar_sorted + C::NS;        // full[NS][NSUB]

    // role order by hardware warp id: B2BU_SORT_FIRST = 1 puts the sorter warps at the low ids
#ifndef B2BU_SORT_FIRST
#define B2BU_SORT_FIRST 0
#endif
    const int tid = threadIdx.x, lane = tid & 31;
    const int hwarp = tid >> 5;
    const int warp = B2BU_SORT_FIRST ? (hwarp < C::SORT_WARPS ? hwarp + C::WORK_WARPS : hwarp < C::SORT_WARPS + C::WORK_WARPS ? hwarp - C::SORT_WARPS : hwarp) : hwarp;

    // this CTA's contiguous block range, cut into equal tiles of at most TILE blocks (multiples of 32)
    const uint64_t r0 = cta_range_start(nblocks, range_quot, range_rem, blockIdx.x, gridDim.x);
    const uint64_t r1 = cta_range_start(nblocks, range_quot, range_rem, blockIdx.x + 1, gridDim.x);
    const uint32_t rlen = (uint32_t)(r1 - r0);
    // tile k covers [tile_start(k), tile_start(k + 1)) of the range.  The first two tiles are short (TILE/4, TILE/2)
    // so that the workers start early; the rest of the range is cut into equal tiles of at most TILE blocks.
    // B2BU_STAGGER: CTAs that start together stay in lockstep, so all 148 load at once, then all sort, all work, all store --
    // every load and store runs at 1/148 of the HBM bandwidth and the memory system idles in between.  Different first-tile
    // sizes put the CTAs at different phases of the tile period.
    const uint32_t want0 = (uint32_t)C::TILE / 4 + (B2BU_STAGGER ? (blockIdx.x % (uint32_t)B2BU_STAGGER) * ((uint32_t)C::TILE * 3 / 4 / (uint32_t)(B2BU_STAGGER > 1 ? B2BU_STAGGER - 1 : 1)) & ~31u : 0u);
    const uint32_t a0 = rlen < want0 ? rlen : want0;
    const uint32_t a1 = rlen - a0 < (uint32_t)C::TILE / 2 ? rlen - a0 : (uint32_t)C::TILE / 2;
    const uint32_t rest = rlen - a0 - a1;
    const uint32_t nrest = (rest + C::TILE - 1) / C::TILE;
    const uint32_t tsz = nrest ? (((rest + nrest - 1) / nrest + 31u) & ~31u) : 0u;
    const uint32_t ntiles = (a0 ? 1u : 0u) + (a1 ? 1u : 0u) + nrest;
    auto tile_start = [&](uint32_t k) -> uint32_t {
        if (k == 0) return 0u;
        if (k == 1) return a0;
        const uint32_t o = a0 + a1 + (k - 2) * tsz;
        return o < rlen ? o : rlen;
    };
    auto tile_blocks = [&](uint32_t k) -> uint32_t { return tile_start(k + 1) - tile_start(k); };

    if (tid == 0) {
        for (int i = 0; i < C::NS * C::NSUB; i++) mbar_init(&bar_full[i], 1);
        for (int i = 0; i < C::NS; i++) { mbar_init(&bar_sorted[i], 1); mbar_init(&bar_done[i], C::WORK_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    load_tables(&T, TableBytes<TARGET>::value);  // ends with __syncthreads()
    if (tid == 0) TRACE(60);

    if (warp == C::WORK_WARPS + C::SORT_WARPS) {
        // ================================ DMA warp ================================
        if (lane != 0) return;
        auto store_tile = [&](uint32_t k) {
            const uint32_t s = k % C::NS, nt = tile_blocks(k);
            const uint64_t g0 = r0 + tile_start(k);
            if (C::DIRECT) return;                             // the workers have written the image themselves
            fence_async_smem();
            if (TARGET == TGT_RGBA) {
                // four pixel rows per block row; a tile may span several block rows
                const uint4* src123 = reinterpret_cast<const uint4*>(out_s + s * C::OUT_SLOT);   // pixel rows 1-3: [3][TILE]
                const uint4* src0 = in_s + s * C::TILE;                                            // pixel row 0: in place
                uint32_t i = 0;
                uint64_t by = g0 / blocks_per_row;
                uint32_t bx = (uint32_t)(g0 - by * blocks_per_row);
                while (i < nt) {
                    const uint32_t seg = nt - i < blocks_per_row - bx ? nt - i : blocks_per_row - bx;
#pragma unroll
                    for (int y = 0; y < 4; y++)
                        tma_store_1d_nocommit(reinterpret_cast<uint4*>(out) + (by * 4 + y) * blocks_per_row + bx,
                                              (y == 0 ? src0 : src123 + (y - 1) * C::TILE) + i, seg * 16u);
                    i += seg; bx = 0; by++;
                }
            } else if (TARGET == TGT_ETC1) {
                const uint2* src = reinterpret_cast<const uint2*>(out_s + s * C::OUT_SLOT);
                uint2* dst = reinterpret_cast<uint2*>(out) + g0;
                if (nt >> 1) tma_store_1d_nocommit(dst, src, (nt >> 1) * 16u);
                if (nt & 1u) dst[nt - 1] = src[nt - 1];          // bulk copies move multiples of 16 bytes
            } else {
                tma_store_1d_nocommit(reinterpret_cast<uint4*>(out) + g0, in_s + s * C::TILE, nt * 16u);
            }
            tma_store_commit();
        };
        for (uint32_t k = 0; k < ntiles; k++) {
            const uint32_t s = k % C::NS, u = k / C::NS;
            if (k >= (uint32_t)C::NS) {                        // slot reuse: tile k-NS must be finished and stored
                B2BU_DMA_WAIT(&bar_done[s], (u - 1u) & 1u);
                store_tile(k - C::NS);
                tma_store_wait_read();
            }
            const uint32_t ntk = tile_blocks(k);
#pragma unroll
            for (int q = 0; q < C::NSUB; q++) {                // every part's barrier completes once per tile, empty parts too
                const uint32_t lo = (uint32_t)(q * C::SUB), hi = ntk < lo + (uint32_t)C::SUB ? ntk : lo + (uint32_t)C::SUB;
                const uint32_t bytes = hi > lo ? (hi - lo) * 16u : 0u;
                mbar_expect_tx(&bar_full[s * C::NSUB + q], bytes);
                if (bytes) tma_load_1d(in_s + s * C::TILE + lo, in + r0 + tile_start(k) + lo, bytes, &bar_full[s * C::NSUB + q]);
            }
            do { if (k < 6) TRACE(k * 6 + 0); } while (0);
        }
        for (uint32_t k = ntiles >= (uint32_t)C::NS ? ntiles - C::NS : 0; k < ntiles; k++) {
            B2BU_DMA_WAIT(&bar_done[k % C::NS], (k / C::NS) & 1u);
            store_tile(k);
        }
        tma_store_wait_all();
        TRACE(59);
        return;
    }

    if (warp >= C::WORK_WARPS) {
        // ================================ sorter warps ================================
        const int sw = warp - C::WORK_WARPS, st = sw * 32 + lane;
        // block of sorter thread st inside each run of SORT_THREADS blocks.  Real textures have runs of one mode, and 32
        // consecutive blocks of one mode under one warp instruction are a 32-way same-address atomic: B2BU_SORT_SPREAD
        // puts the lanes of a warp 9 blocks apart (a bijection on 0..255; the 16-byte stride keeps the 4-way bank pattern).
        const int bst = (B2BU_SORT_SPREAD && (C::SORT_THREADS & (C::SORT_THREADS - 1)) == 0) ? ((st * 9) & (C::SORT_THREADS - 1)) : st;
        uint32_t* mycnt = wcnt + sw * 32;
        for (uint32_t k = 0; k < ntiles; k++) {
            const uint32_t s = k % C::NS, u = k / C::NS, nt = tile_blocks(k);
            const uint4* tin = in_s + s * C::TILE;
            mycnt[lane] = 0;
            __syncwarp();
            mbar_wait(&bar_full[s * C::NSUB], u & 1u);
            if (st == 0) do { if (k < 6) TRACE(k * 6 + 1); } while (0);
            // A: classify, rank inside (warp, mode).  Shared-memory latency is hundreds of cycles while the workers
            // keep the LSU busy, so every step is a branch-free pass over all of the thread's blocks (PERS loads /
            // atomics in flight); passes are cut short on the two short start-up tiles (jmax is warp-uniform).
            const int jmax = (int)((nt + C::SORT_THREADS - 1) / C::SORT_THREADS);
#ifdef B2BU_TRACE
#define PH(n) do { if (k == 4 && lane == 0 && (sw == 0 || sw == C::SORT_WARPS - 1)) TRACE(40 + (sw ? 10 : 0) + (n)); } while (0)
#else
#define PH(n) do { } while (0)
#endif
            PH(0);
            uint32_t mr[C::PERS];
#pragma unroll
            for (int j = 0; j < C::PERS; j++) {
                // blocks past the end of a short tile read stale slot contents (always inside the slot); they are discarded below
                mr[j] = 31u;
                if (j > 0 && (j * C::SORT_THREADS) % C::SUB == 0 && j < jmax) mbar_wait(&bar_full[s * C::NSUB + (j * C::SORT_THREADS) / C::SUB], u & 1u);
                if (j < jmax) mr[j] = tin[bst + j * C::SORT_THREADS].x & 127u;
            }
#pragma unroll
            for (int j = 0; j < C::PERS; j++) {
                if (j < jmax) {
                    const uint32_t lut = T.mode_lut[mr[j]];
                    mr[j] = (uint32_t)(bst + j * C::SORT_THREADS) < nt ? lut : 31u;     // 31 = unused bin
                }
            }
            PH(1);
            {
                // rank = old value of the warp's private counter (direct atomics: a MATCH.ANY / ballot + leader
                // scheme measured 2-3x slower because each step waits for the previous atomic's round trip).  Real
                // textures are spatially coherent, and 32 consecutive blocks of one mode would be a 32-way same-address
                // atomic: with B2BU_SORT_UNIFORM a warp whose 32 blocks agree adds 32 once and the shuffles that hand out
                // the ranks run in the second pass -- it costs the mixed case more than it gains on runs, so it is off.
                uint32_t rk[C::PERS];
#if B2BU_SORT_UNIFORM
                uint32_t unim = 0u;
#pragma unroll
                for (int j = 0; j < C::PERS; j++) {
                    if (j < jmax) {
                        const uint32_t m = mr[j];
                        const bool uni = __all_sync(0xFFFFFFFFu, m == __shfl_sync(0xFFFFFFFFu, m, 0));
                        unim |= (uni ? 1u : 0u) << j;
                        rk[j] = 0u;
                        if (!uni || lane == 0) rk[j] = atomicAdd(&mycnt[m], uni ? 32u : 1u);
                    }
                }
#pragma unroll
                for (int j = 0; j < C::PERS; j++) {
                    if (j < jmax) {
                        if ((unim >> j) & 1u) rk[j] = __shfl_sync(0xFFFFFFFFu, rk[j], 0) + (uint32_t)lane;
                        mr[j] |= rk[j] << 8;
                    }
                }
#else
#pragma unroll
                for (int j = 0; j < C::PERS; j++) if (j < jmax) rk[j] = atomicAdd(&mycnt[mr[j]], 1u);
#pragma unroll
                for (int j = 0; j < C::PERS; j++) if (j < jmax) mr[j] |= rk[j] << 8;
#endif
            }
            PH(2);
            named_bar_sync(1, C::SORT_THREADS);
            PH(3);
            // B: bin offsets (each bin padded to a multiple of 32, heaviest mode first).  Every sorter warp computes the
            // scan redundantly from the warp counters (no hand-off, no second barrier); lane l owns bin kBinOrder[l].
            {
                const uint32_t bin = lane < kBins ? (uint32_t)kBinOrder[lane] : 31u;
                uint32_t cw[C::SORT_WARPS], c = 0, mine_base = 0;
#pragma unroll
                for (int w = 0; w < C::SORT_WARPS; w++) cw[w] = wcnt[w * 32 + bin];
#pragma unroll
                for (int w = 0; w < C::SORT_WARPS; w++) { if (w == sw) mine_base = c; c += lane < kBins ? cw[w] : 0u; }
                const uint32_t padded = (c + 31u) & ~31u;
                uint32_t incl = padded;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) { const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, d); if (lane >= d) incl += v; }
                const uint32_t excl = incl - padded;
                wbase[sw * 32 + bin] = mine_base + excl;                        // this warp's first slot in each bin (bin 31 is a dummy)
                if (sw == 0) {
                    bintab[s * 32 + lane] = (excl >> 5) | (c << 16);             // workers: item -> (mode, lanes) without a per-item table
                    if (lane == 31) { ctl[s * 4 + 0] = 0; ctl[s * 4 + 1] = incl >> 5; }
                }
            }
            PH(4);
            __syncwarp();
            PH(5);
            // C: scatter block indices into their bins
            const uint32_t* mybase = wbase + sw * 32;
            uint16_t* ord = order + s * C::MAXORD;
            {
                uint32_t bs[C::PERS];
#pragma unroll
                for (int j = 0; j < C::PERS; j++) if (j < jmax) bs[j] = mybase[mr[j] & 0xFFu];
#pragma unroll
                for (int j = 0; j < C::PERS; j++)
                    if (j < jmax && (mr[j] & 0xFFu) < (uint32_t)kBins) ord[bs[j] + (mr[j] >> 8)] = (uint16_t)(bst + j * C::SORT_THREADS);
            }
            PH(6);
            named_bar_sync(1, C::SORT_THREADS);
            PH(7);
            if (st == 0) { mbar_arrive(&bar_sorted[s]); do { if (k < 6) TRACE(k * 6 + 2); } while (0); }
        }
        return;
    }

    // ================================ worker warps ================================
    for (uint32_t k = 0; k < ntiles; k++) {
        const uint32_t s = k % C::NS, u = k / C::NS;
        uint4* tin = in_s + s * C::TILE;
        unsigned char* tout = out_s + s * C::OUT_SLOT;
        const uint16_t* ord = order + s * C::MAXORD;
        const uint64_t base = r0 + tile_start(k);
#ifdef B2BU_TRACE
        const long long tw0 = clock64();
#endif
        mbar_wait(&bar_sorted[s], u & 1u);
#ifdef B2BU_TRACE
        if (lane == 0) { if (warp == 0) { TRACE_ADD(61, clock64() - tw0); do { if (k < 6) TRACE(k * 6 + 3); } while (0); } if (warp == C::WORK_WARPS - 1) TRACE_ADD(62, clock64() - tw0); }
#endif
#pragma unroll
        for (int q = 0; q < C::NSUB; q++) mbar_wait(&bar_full[s * C::NSUB + q], u & 1u);   // completed long ago: observes the bulk-copied bytes directly
        const uint32_t nitems = ctl[s * 4 + 1];
        const uint32_t bt = bintab[s * 32 + lane];
        // Items are pulled in bin order from a shared counter.  Besides balancing uneven items this keeps every worker of
        // the SM inside the same few modes, i.e. the same few KB of code: dealing the items round-robin instead let the
        // warps drift apart and ran the large-code targets (ETC1/ETC2) 3x slower on instruction-cache misses.  ASTC's
        // code is small enough for the static deal to win (no counter round trip per item).
        for (uint32_t it = (uint32_t)warp;; it += C::WORK_WARPS) {
            uint32_t item = it;
            if (C::DYNAMIC) {
                if (lane == 0) item = atom_add_shared(&ctl[s * 4 + 0], 1u);
                item = __shfl_sync(0xFFFFFFFFu, item, 0);
            }
            if (item >= nitems) break;
            // bin of this item: lanes hold the bins' first items in bin order (non-decreasing; empty bins repeat the value)
            const uint32_t nle = (uint32_t)__popc(__ballot_sync(0xFFFFFFFFu, lane < kBins && (bt & 0xFFFFu) <= item));
            const uint32_t bsel = __shfl_sync(0xFFFFFFFFu, bt, (int)nle - 1);
            const uint32_t mode = kBinOrder[nle - 1u];
            const uint32_t rest = (bsel >> 16) - 32u * (item - (bsel & 0xFFFFu));      // blocks of the bin from this item on
            if ((uint32_t)lane < rest) {
                const uint32_t idx = ord[item * 32 + lane];
                const uint4 b = tin[idx];
                BlockOut o;
                TileRowSink sink{tin + idx, reinterpret_cast<uint4*>(tout) + idx, (uint64_t)C::TILE};
                if (C::DIRECT) {
                    // uastc.rs:96-106: block (bx, by) of the row-major image, pitch 4 * blocks_per_row pixels (a device holds
                    // fewer than 2^32 blocks of 80 bytes)
                    const uint32_t gi = (uint32_t)(base + idx), by = gi / blocks_per_row, bx = gi - by * blocks_per_row;
                    uint4* g = reinterpret_cast<uint4*>(out) + (uint64_t)by * 4u * blocks_per_row + bx;
                    sink = TileRowSink{g, g + blocks_per_row, (uint64_t)blocks_per_row};
                }
#ifdef B2BU_NULL_WORK      // tuning aid: the pipeline without the transcode (blocks are copied)
                const uint32_t e = mode == 19u ? (uint32_t)ERR_MODE : (uint32_t)ERR_OK;
                o.v = b; o.etc = make_uint2(b.x, b.y);
#else
                const uint32_t e = transcode_mode_sink<TARGET>(mode, b, T, o, sink);
#endif
                if (e != ERR_OK) {
                    report_error(err, index_base + base + idx, e);
                    o.v = make_uint4(0u, 0u, 0u, 0u); o.etc = make_uint2(0u, 0u);
                    if (TARGET == TGT_RGBA) {
#pragma unroll 1
                        for (int y = 0; y < 4; y++) sink.row(y, make_uint4(0u, 0u, 0u, 0u));
                    }
                }
                if (TARGET == TGT_ETC1) reinterpret_cast<uint2*>(tout)[idx] = o.etc;
                else if (TARGET != TGT_RGBA) tin[idx] = o.v;
            }
        }
        fence_async_smem();                       // generic-proxy writes -> visible to the bulk store
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar_done[s]);
        if (lane == 0 && warp == 0) do { if (k < 6) TRACE(k * 6 + 4); } while (0);
    }
    if (lane == 0 && warp == 0) TRACE(63);
}

template <int TARGET>
static cudaError_t launch_sorted(const uint4* in, void* d_out, uint64_t nblocks, uint32_t bpr, uint64_t index_base,
                                 unsigned long long* d_err, int sm_count, cudaStream_t stream)
{
    using C = PipeCfg<TARGET>;
    static bool configured[kMaxDevicesK] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < kMaxDevicesK && !configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(uastc_sorted_kernel<TARGET>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
        if (e != cudaSuccess) return e;
        configured[dev] = true;
    }
    // one persistent CTA per SM; fewer when the input is small (at least ~one half tile each)
    const uint64_t want = (nblocks + C::TILE / 2 - 1) / (C::TILE / 2);
    const uint64_t cap = (uint64_t)sm_count * B2BU_CTAS_PER_SM;
    const unsigned grid = (unsigned)(want < cap ? want : cap);
    uastc_sorted_kernel<TARGET><<<grid, C::THREADS, C::SMEM, stream>>>(in, d_out, nblocks, bpr, index_base, d_err, nblocks / grid, (uint32_t)(nblocks % grid));
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

#ifdef B2BU_TRACE
extern "C" __attribute__((visibility("default"))) int b2bu_debug_trace(unsigned long long* dst, int reset)
{
    if (reset) { static unsigned long long z[160][64]; return (int)cudaMemcpyToSymbol(g_trace, z, sizeof z); }
    return (int)cudaMemcpyFromSymbol(dst, g_trace, sizeof(unsigned long long) * 160 * 64);
}
#endif

}  // namespace b2bu
