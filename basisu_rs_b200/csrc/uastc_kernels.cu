// K1<target>: fused UASTC unpack + repack kernels.
// One thread per 16-byte block, 128-bit coalesced loads and stores, constant tables in shared
// memory.  Replaces uastc::Decoder::{transcode,decode_to_rgba} (reference src/uastc.rs:89-165).
#include <atomic>
#include <cstddef>
#include <mutex>
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

// predicated forms (straight-line code: left to the compiler, a conditional atomic or store becomes a branch region per use)
// (each takes the two sides of its condition  a < b  so that the predicate is one compare, not a materialised bool)
__device__ __forceinline__ uint32_t atom_inc_shared(uint32_t saddr)
{
    uint32_t old;
    asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(old) : "r"(saddr));
    return old;
}
__device__ __forceinline__ void sts_u16_if_lt(uint32_t saddr, uint32_t v, uint32_t a, uint32_t b)
{
    asm volatile("{\n.reg .pred p;\nsetp.lt.u32 p, %2, %3;\n@p st.shared.u16 [%0], %1;\n}\n" ::"r"(saddr), "h"((uint16_t)v), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ void cp_async_4_if_lt(uint32_t sdst, const void* gsrc, uint32_t a, uint32_t b)
{
    asm volatile("{\n.reg .pred p;\nsetp.lt.u32 p, %2, %3;\n@p cp.async.ca.shared.global [%0], [%1], 4;\n}\n" ::"r"(sdst), "l"(gsrc), "r"(a), "r"(b) : "memory");
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
// (46-190 KB) thrashes the instruction cache when every warp of an SM runs a different mode.  So each
// CTA walks a contiguous range of blocks in tiles and runs three roles concurrently:
//
//   DMA warp (1 lane)   TMA bulk load of tile k+1 into the free data slot, bulk store of finished tiles
//   sorter warps        classify a tile by UASTC mode and counting-sort its block indices into mode bins
//                       padded to 32 (per-warp shared histograms), then write one descriptor per 32-block
//                       work item.  The mode bytes are read straight from GLOBAL memory, so the sort does
//                       not wait for the tile's bulk load and does not occupy a data slot: it runs up to
//                       kOrderSlots - 1 tiles ahead of the workers and is off their critical path (with
//                       the sort reading the loaded tile, load -> sort -> work of one slot was a serial
//                       chain longer than the work on the other slot: the workers idled 20-37 %).
//   worker warps        pull 32-block, mode-uniform work items of tile k in bin order (heavy modes first)
//                       and run the specialised transcode.  All workers of the SM are inside the same few
//                       modes at any time, which keeps the hot code I-cache resident.
//
// The roles are connected by mbarriers (data slot: full -> done; order slot: sorted -> free); there is
// no __syncthreads in the steady state and workers run on from tile k into tile k+1 without waiting
// for each other.  16-byte results overwrite the block's own input slot in shared memory and leave
// with one bulk store per tile; ETC1 (8 B) and RGBA (64 B, stored as four pixel rows per block row)
// are staged in a second buffer so that every global write is a full-line bulk store.
// ------------------------------------------------------------------------------------------
constexpr int kBins = 20;                       // modes 0..18 + the invalid code (19)
// processing order of the bins: roughly by decreasing per-block cost (endpoint count, trits/quints, subsets)
__constant__ uint8_t kBinOrder[32] = {3, 9, 16, 4, 7, 2, 12, 10, 11, 13, 14, 6, 0, 18, 5, 1, 17, 15, 8, 19,
                                       31, 31, 31, 31, 31, 31, 31, 31, 31, 31, 31, 31};

// tile sizes in blocks, each the best of a 2560..4096 scan on B200 (profiles/tile_scan_r02.txt; the spread over that
// range is 2..6 %).  RGBA stages 48 B and ETC1 8 B of output per block and slot beside the 16 B of input.
#ifndef B2BU_TILE16
#define B2BU_TILE16 4096        // ETC2, and BC7 unless overridden
#endif
#ifndef B2BU_TILE_ASTC
#define B2BU_TILE_ASTC 3584
#endif
#ifndef B2BU_TILE_BC7
#define B2BU_TILE_BC7 B2BU_TILE16
#endif
#ifndef B2BU_TILE_RGBA
#define B2BU_TILE_RGBA 1536
#endif
#ifndef B2BU_TILE_ETC1
#define B2BU_TILE_ETC1 3072
#endif
#ifndef B2BU_SORT_WARPS
#define B2BU_SORT_WARPS 8
#endif
#ifndef B2BU_WORK_WARPS
#define B2BU_WORK_WARPS 23
#endif
#ifndef B2BU_ORDER_SLOTS
#define B2BU_ORDER_SLOTS 2
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

template <int TARGET> struct PipeCfg {
    static constexpr int OB = TARGET == TGT_RGBA ? 64 : TARGET == TGT_ETC1 ? 8 : 16;
    static constexpr bool IN_PLACE = OB == 16;
    static constexpr int NS = 2;                   // data slots: one being worked on, one being stored / loaded
    static constexpr int NO = B2BU_ORDER_SLOTS;    // order slots: how far the sorter may run ahead
    static constexpr bool DYNAMIC = TARGET != TGT_ASTC;
    static constexpr int TILE = TARGET == TGT_RGBA ? B2BU_TILE_RGBA : TARGET == TGT_ETC1 ? B2BU_TILE_ETC1
                              : TARGET == TGT_ASTC ? B2BU_TILE_ASTC : TARGET == TGT_BC7 ? B2BU_TILE_BC7 : B2BU_TILE16;
    static constexpr int SORT_WARPS = B2BU_SORT_WARPS;
    static constexpr int SORT_THREADS = SORT_WARPS * 32;
    static constexpr int WORK_WARPS = B2BU_WORK_WARPS;
    static constexpr int THREADS = 32 * (1 + SORT_WARPS + WORK_WARPS);
    static constexpr int PERS = (TILE + SORT_THREADS - 1) / SORT_THREADS;     // blocks per sorter thread
    static constexpr int MAXORD = TILE + kBins * 32;
    static constexpr int MAXITEMS = MAXORD / 32;
    static constexpr size_t OFF_IN = (TableBytes<TARGET>::value + 127) / 128 * 128;
    static constexpr size_t OFF_OUT = OFF_IN + NS * (size_t)TILE * 16;                 // staging for ETC1 / RGBA, one per slot
    // RGBA stages 48 B per block: pixel row 0 goes where the block's input was (TileRowSink)
    static constexpr size_t OUT_SLOT = IN_PLACE ? 0 : TARGET == TGT_RGBA ? (size_t)TILE * 48 : (size_t)TILE * OB;
    static constexpr size_t OFF_ORDER = OFF_OUT + NS * OUT_SLOT;
    static constexpr size_t OFF_ITEMS = OFF_ORDER + NO * (size_t)MAXORD * 2;
    // first word (it holds the mode code) of every block of the tile being sorted and of the next one, fetched with cp.async
    static constexpr int STAGE = PERS * SORT_THREADS;
    static constexpr size_t OFF_STAGE = (OFF_ITEMS + NO * (size_t)MAXITEMS * 2 + 15) / 16 * 16;
    static constexpr size_t OFF_WCNT = OFF_STAGE + 2 * (size_t)STAGE * 4;
    static constexpr size_t OFF_BASE = OFF_WCNT + (size_t)SORT_WARPS * 32 * 4;
    static constexpr size_t OFF_CTL = OFF_BASE + (size_t)SORT_WARPS * 32 * 4;
    static constexpr int ND = 8;                   // tile descriptors in flight (first block, blocks): sorter -> DMA lane, workers
    static constexpr size_t OFF_DESC = (OFF_CTL + NO * 2 * 4 + 7) / 8 * 8;
    static constexpr size_t OFF_BAR = OFF_DESC + ND * 8 + 8;
    static constexpr size_t SMEM = OFF_BAR + (2 * NS + 2 * NO) * 8;
    static_assert(SMEM <= 227 * 1024, "tile configuration does not fit shared memory");
    static_assert(THREADS <= 1024, "too many warps");
    static_assert((SORT_THREADS & (SORT_THREADS - 1)) == 0, "the sorter's block permutation needs a power of two");
    static_assert(ND >= NO + NS + 2, "a descriptor must outlive the store of its tile");
};

// Tiles are handed out DYNAMICALLY: the CTAs of a launch draw block ranges from one global cursor (sched[0]).  Equal static
// shares left the slowest SM 15 % behind the median (identical work, different instruction-fetch and memory latencies), and
// the launch lasts as long as its slowest CTA.  Tile 0 of every CTA is a fixed short one (TILE/4 blocks at blockIdx * TILE/4:
// no round trip to the cursor in front of the first load, and the workers start early); everything behind the first round is
// drawn from the cursor in full tiles.  (Guided self-scheduling -- tiles that shrink towards the end of the launch so that the
// CTAs finish together -- was measured and lost: a tile's fixed costs, the sort's barriers and the bins' padding to 32 blocks,
// weigh more on small tiles than the imbalance of at most one tile costs.  ASTC: full tiles 58 us, shrinking to a quarter 60,
// to an eighth 66.)

template <int TARGET>
__global__ void __launch_bounds__(PipeCfg<TARGET>::THREADS, 1)
uastc_sorted_kernel(const uint4* __restrict__ in, void* __restrict__ out, uint32_t nblocks, uint32_t blocks_per_row,
                    uint64_t index_base, unsigned long long* __restrict__ err, unsigned int* __restrict__ sched)
{
    using C = PipeCfg<TARGET>;
    extern __shared__ __align__(128) unsigned char smem[];
    DevTables& T = *reinterpret_cast<DevTables*>(smem);
    uint4* in_s = reinterpret_cast<uint4*>(smem + C::OFF_IN);                 // [NS][TILE]
    unsigned char* out_s = smem + C::OFF_OUT;                                 // [NS][OUT_SLOT]
    uint16_t* order = reinterpret_cast<uint16_t*>(smem + C::OFF_ORDER);       // [NO][MAXORD] block index inside the tile, by bin
    uint16_t* items = reinterpret_cast<uint16_t*>(smem + C::OFF_ITEMS);       // [NO][MAXITEMS] mode | blocks of the item << 8
    uint32_t* stage = reinterpret_cast<uint32_t*>(smem + C::OFF_STAGE);       // [2][STAGE]
    uint32_t* wcnt = reinterpret_cast<uint32_t*>(smem + C::OFF_WCNT);         // [SORT_WARPS][32]
    uint32_t* wbase = reinterpret_cast<uint32_t*>(smem + C::OFF_BASE);        // [SORT_WARPS][32]
    uint32_t* ctl = reinterpret_cast<uint32_t*>(smem + C::OFF_CTL);           // [NO][2]: next item, number of items
    uint2* tdesc = reinterpret_cast<uint2*>(smem + C::OFF_DESC);              // [ND] tile k: (first block, blocks); blocks == 0 ends the CTA
    volatile uint32_t* desc_seq = reinterpret_cast<volatile uint32_t*>(smem + C::OFF_DESC + C::ND * 8);   // descriptors published so far
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::OFF_BAR);
    uint64_t* bar_full = bars, *bar_done = bars + C::NS, *bar_sorted = bars + 2 * C::NS, *bar_ofree = bar_sorted + C::NO;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // Programmatic dependent launch: the NEXT kernel of the stream may be scheduled onto an SM as soon as this launch's CTA has
    // left it, and runs its prologue (constant tables -> shared memory, barrier initialisation: nothing an earlier kernel writes)
    // while the slowest CTAs of this launch are still working; griddepcontrol.wait below holds it back until this launch -- and
    // everything before it in the stream -- has completed and flushed, so stream order is what it always was.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (tid == 0) *desc_seq = 0u;
    if (tid == 0) {
        for (int i = 0; i < C::NS; i++) { mbar_init(&bar_full[i], 1); mbar_init(&bar_done[i], C::WORK_WARPS); }
        for (int i = 0; i < C::NO; i++) { mbar_init(&bar_sorted[i], 1); mbar_init(&bar_ofree[i], C::WORK_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    load_tables(&T, TableBytes<TARGET>::value);  // ends with __syncthreads()
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (tid == 0) TRACE(60);
#ifdef B2BU_TRACE
    if (tid == 0) { unsigned long long gt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt)); g_trace[blockIdx.x][57] = gt; }
#endif

    if (warp == C::WORK_WARPS + C::SORT_WARPS) {
        // ================================ DMA warp ================================
        if (lane != 0) return;
        auto store_tile = [&](uint32_t k) {
            const uint32_t s = k % C::NS, nt = tdesc[k % C::ND].y;
            const uint64_t g0 = tdesc[k % C::ND].x;
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
        uint32_t k = 0;
        for (;; k++) {
            while (*desc_seq <= k) asm volatile("nanosleep.u32 100;");          // the sorter publishes tile k's descriptor
            __threadfence_block();
            const uint2 d = tdesc[k % C::ND];
            if (d.y == 0u) break;                              // no more work for this CTA
            const uint32_t s = k % C::NS, u = k / C::NS;
            if (k >= (uint32_t)C::NS) {                        // slot reuse: tile k-NS must be finished and stored
                mbar_wait_backoff(&bar_done[s], (u - 1u) & 1u, 200);
                store_tile(k - C::NS);
                tma_store_wait_read();
            }
            mbar_expect_tx(&bar_full[s], d.y * 16u);
            tma_load_1d(in_s + s * C::TILE, in + d.x, d.y * 16u, &bar_full[s]);
            do { if (k < 6) TRACE(k * 6 + 0); } while (0);
        }
        const uint32_t ntiles = k;
        for (uint32_t j = ntiles >= (uint32_t)C::NS ? ntiles - C::NS : 0; j < ntiles; j++) {
            mbar_wait_backoff(&bar_done[j % C::NS], (j / C::NS) & 1u, 200);
            store_tile(j);
        }
        tma_store_wait_all();
        TRACE(59);
#ifdef B2BU_TRACE
        { unsigned long long gt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt)); g_trace[blockIdx.x][58] = gt; }
#endif
        // the last CTA of the launch to finish rewinds the cursor for the next launch that uses this slot
        if (atomicInc(&sched[1], gridDim.x - 1u) == gridDim.x - 1u) sched[0] = 0u;
        return;
    }

    if (warp >= C::WORK_WARPS) {
        // ================================ sorter warps ================================
        const int sw = warp - C::WORK_WARPS, st = sw * 32 + lane;
        // Block of sorter thread st inside each run of SORT_THREADS blocks.  Real textures have runs of one mode, and 32
        // consecutive blocks of one mode under one warp instruction would be a 32-way same-address atomic: lane pairs stay
        // together (the two blocks of one 32-byte sector), the pairs of a warp are 9 sectors apart (a bijection on 0..SORT_THREADS-1).
        const int bst = ((((st >> 1) * 9) & (C::SORT_THREADS / 2 - 1)) << 1) | (st & 1);
        uint32_t* mycnt = wcnt + sw * 32;
        // The mode words of tile k+1 are fetched (4-byte cp.async, global -> shared, no registers in flight) while tile k is
        // being sorted, so that the global-memory latency is not part of the per-tile chain.  The copies are coalesced (a warp
        // instruction covers 32 consecutive blocks = four whole 128-byte lines: scattered 32-byte sector reads ran the DRAM at a
        // fraction of its bandwidth); the classification reads the staged words back in the spread order (conflict-free:
        // the 16 lane pairs fall on different even banks), which another thread has copied -- hence a barrier after the wait.
        auto blocks_of = [&](uint32_t nt, int first) -> uint32_t { return nt > (uint32_t)first ? (nt - (uint32_t)first + C::SORT_THREADS - 1) / C::SORT_THREADS : 0u; };
        auto prefetch = [&](uint32_t k) {
            const uint2 d = tdesc[k % C::ND];
            const uint32_t myj = blocks_of(d.y, st);          // 0 for the end marker
            const uint4* src = in + d.x + st;
            const uint32_t dst = smem_u32(stage + (k & 1u) * C::STAGE + st);
#pragma unroll
            for (int j = 0; j < C::PERS; j++) cp_async_4_if_lt(dst + j * C::SORT_THREADS * 4, src + j * C::SORT_THREADS, (uint32_t)j, myj);
            cp_async_commit();
        };
        // thread 0 draws tile j's block range from the launch's cursor and publishes it; the other sorter threads may read it
        // after the next named barrier, the DMA lane polls desc_seq, the workers read it behind bar_sorted
        auto draw = [&](uint32_t j) {
            if (st != 0) return;
            uint2 d = make_uint2(0u, 0u);
            constexpr uint32_t T0 = (uint32_t)C::TILE / 4u;
            const uint32_t round0 = gridDim.x * T0;            // blocks of the fixed first round
            if (j == 0u) {
                const uint32_t pos = blockIdx.x * T0;
                if (pos < nblocks) d = make_uint2(pos, nblocks - pos < T0 ? nblocks - pos : T0);
            } else if (round0 < nblocks) {
                const uint32_t cur = *reinterpret_cast<volatile unsigned int*>(&sched[0]);
                // (Holding the sorter back near the end of the launch, so that no CTA sits on drawn tiles when the cursor runs
                // out, halves the spread of the CTAs' finishing times but costs every CTA more than it saves: 60 -> 62 us.)
                if (cur < nblocks - round0) {
                    const uint32_t pos = round0 + atomicAdd(&sched[0], (uint32_t)C::TILE);
                    if (pos < nblocks) d = make_uint2(pos, nblocks - pos < (uint32_t)C::TILE ? nblocks - pos : (uint32_t)C::TILE);
                }
            }
            tdesc[j % C::ND] = d;
            __threadfence_block();
            *desc_seq = j + 1u;
        };
        draw(0);
        named_bar_sync(1, C::SORT_THREADS);
        prefetch(0);
        for (uint32_t k = 0;; k++) {
            const uint32_t o = k % C::NO, nt = tdesc[k % C::ND].y;
            if (nt == 0u) {
                // end marker: hand it to the workers through the same barrier (behind the release of the order slot, so that no
                // worker can still be waiting on the slot's previous phase)
                if (k >= (uint32_t)C::NO) mbar_wait(&bar_ofree[o], (k / C::NO - 1u) & 1u);
                if (st == 0) mbar_arrive(&bar_sorted[o]);
                break;
            }
            draw(k + 1);
            mycnt[lane] = 0;
            __syncwarp();
            cp_async_wait_group_0();                            // this thread's copies of tile k's words have landed
            named_bar_sync(1, C::SORT_THREADS);                 // ... and everybody else's; tile k+1's descriptor is visible
            prefetch(k + 1);
            // A: classify, rank inside (warp, mode).  Every step is a branch-free pass over all of the thread's blocks (PERS
            // loads / atomics in flight); blocks past the end of a short tile are predicated off.
            const uint32_t myj = blocks_of(nt, bst);            // passes j in which block bst + j * SORT_THREADS is inside the tile
            // (the twelve unused counters 20..31 all serve as dummy bins, spread over the lanes: in a short tile -- every CTA's
            // first one -- most of a warp's atomics would otherwise hit ONE address and serialise)
            const uint32_t dummy = (uint32_t)kBins + (uint32_t)lane % (32u - (uint32_t)kBins);
            const uint32_t* src = stage + (k & 1u) * C::STAGE + bst;
            uint32_t mr[C::PERS];
#pragma unroll
            for (int j = 0; j < C::PERS; j++) mr[j] = src[j * C::SORT_THREADS] & 127u;      // stale words past the end: discarded below
#pragma unroll
            for (int j = 0; j < C::PERS; j++) { const uint32_t lut = T.mode_lut[mr[j]]; mr[j] = (uint32_t)j < myj ? lut : dummy; }  // blocks past the end: a dummy bin
            // the order slot must have been consumed (tile k - NO)
            if (k >= (uint32_t)C::NO) mbar_wait(&bar_ofree[o], (k / C::NO - 1u) & 1u);
            if (st == 0) do { if (k < 6) TRACE(k * 6 + 1); } while (0);
            {
                // rank = old value of the warp's private counter (direct atomics: a MATCH.ANY / ballot + leader scheme
                // measured 2-3x slower because each step waits for the previous atomic's round trip)
                const uint32_t cnt_base = smem_u32(mycnt);
                uint32_t rk[C::PERS];
#pragma unroll
                for (int j = 0; j < C::PERS; j++) rk[j] = atom_inc_shared(cnt_base + 4u * mr[j]);   // straight-line: blocks past the end count into the dummy bin
#pragma unroll
                for (int j = 0; j < C::PERS; j++) mr[j] |= rk[j] << 8;
            }
            named_bar_sync(1, C::SORT_THREADS);
            // B: bin offsets (each bin padded to a multiple of 32, heaviest mode first).  Every sorter warp computes the
            // scan redundantly from the warp counters (no hand-off, no second barrier); lane l owns bin kBinOrder[l].
            {
                const uint32_t bin = lane < kBins ? (uint32_t)kBinOrder[lane] : 31u;
                // blocks of this bin counted by every (warp, histogram), summed in that order; this warp's histograms start at mine[]
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
                // one descriptor per work item: the workers read (mode, blocks) with a single broadcast load
                uint16_t* it = items + o * C::MAXITEMS + (excl >> 5);
#pragma unroll 1
                for (uint32_t i = (uint32_t)sw; i * 32u < c; i += C::SORT_WARPS) {
                    const uint32_t left = c - i * 32u;
                    it[i] = (uint16_t)(bin | ((left < 32u ? left : 32u) << 8));
                }
                if (sw == 0 && lane == 31) { ctl[o * 2 + 0] = 0; ctl[o * 2 + 1] = incl >> 5; }
            }
            __syncwarp();
            // C: scatter block indices into their bins
            const uint32_t* mybase = wbase + sw * 32;
            uint16_t* ord = order + o * C::MAXORD;
            {
                uint32_t bs[C::PERS];
#pragma unroll
                for (int j = 0; j < C::PERS; j++) bs[j] = mybase[mr[j] & 0xFFu];
                const uint32_t ord_base = smem_u32(ord);
#pragma unroll
                for (int j = 0; j < C::PERS; j++) sts_u16_if_lt(ord_base + 2u * (bs[j] + (mr[j] >> 8)), (uint32_t)(bst + j * C::SORT_THREADS), (uint32_t)j, myj);
            }
            named_bar_sync(1, C::SORT_THREADS);
            if (st == 0) { mbar_arrive(&bar_sorted[o]); do { if (k < 6) TRACE(k * 6 + 2); } while (0); }
        }
        return;
    }

    // ================================ worker warps ================================
    for (uint32_t k = 0;; k++) {
        const uint32_t s = k % C::NS, o = k % C::NO;
        uint4* tin = in_s + s * C::TILE;
        unsigned char* tout = out_s + s * C::OUT_SLOT;
        const uint16_t* ord = order + o * C::MAXORD;
        const uint16_t* itm = items + o * C::MAXITEMS;
#ifdef B2BU_TRACE
        const long long tw0 = clock64();
#endif
        mbar_wait(&bar_sorted[o], (k / C::NO) & 1u);
        const uint2 dk = tdesc[k % C::ND];
        if (dk.y == 0u) break;                    // end marker
        const uint64_t base = dk.x;
#ifdef B2BU_TRACE
        const long long tw1 = clock64();
#endif
        mbar_wait(&bar_full[s], (k / C::NS) & 1u);
#ifdef B2BU_TRACE
        if (lane == 0 && warp == 0) { TRACE_ADD(61, tw1 - tw0); TRACE_ADD(62, clock64() - tw1); do { if (k < 6) TRACE(k * 6 + 3); } while (0); }
#endif
        const uint32_t nitems = ctl[o * 2 + 1];
        // Items are pulled in bin order from a shared counter.  Besides balancing uneven items this keeps every worker of
        // the SM inside the same few modes, i.e. the same few KB of code: dealing the items round-robin instead let the
        // warps drift apart and ran the large-code targets (ETC1/ETC2) 3x slower on instruction-cache misses.  ASTC's
        // code is small enough for the static deal to win (no counter round trip per item).
        // (Claiming the NEXT item before the current one is processed hides the counter round trip but costs far more at the end
        // of a tile, where a warp that is busy with a long item then sits on an item an idle warp could have taken: RGBA 97 -> 152 us.)
        for (uint32_t it = (uint32_t)warp;; it += C::WORK_WARPS) {
            uint32_t item = it;
            if (C::DYNAMIC) {
                if (lane == 0) item = atom_add_shared(&ctl[o * 2 + 0], 1u);
                item = __shfl_sync(0xFFFFFFFFu, item, 0);
            }
            if (item >= nitems) break;
            // descriptor and block index are loaded side by side (the index of a padding lane is stale, never used): under
            // load a shared-memory access costs ~200 cycles, and an item's descriptor -> index -> block chain is latency
            // nothing else in the warp can hide
            const uint32_t desc = itm[item];
            const uint32_t idx = ord[item * 32 + lane];
            const uint32_t mode = desc & 0xFFu;
            if ((uint32_t)lane < (desc >> 8)) {
                const uint4 b = tin[idx];
                BlockOut o_;
                TileRowSink sink{tin + idx, reinterpret_cast<uint4*>(tout) + idx, (uint64_t)C::TILE};
#ifdef B2BU_NULL_WORK      // tuning aid: the pipeline without the transcode (blocks are copied)
                const uint32_t e = mode == 19u ? (uint32_t)ERR_MODE : (uint32_t)ERR_OK;
                o_.v = b; o_.etc = make_uint2(b.x, b.y);
#else
                const uint32_t e = transcode_mode_sink<TARGET>(mode, b, T, o_, sink);
#endif
                if (e != ERR_OK) {
                    report_error(err, index_base + base + idx, e);
                    o_.v = make_uint4(0u, 0u, 0u, 0u); o_.etc = make_uint2(0u, 0u);
                    if (TARGET == TGT_RGBA) {
#pragma unroll 1
                        for (int y = 0; y < 4; y++) sink.row(y, make_uint4(0u, 0u, 0u, 0u));
                    }
                }
                if (TARGET == TGT_ETC1) reinterpret_cast<uint2*>(tout)[idx] = o_.etc;
                else if (TARGET != TGT_RGBA) tin[idx] = o_.v;
            }
        }
        fence_async_smem();                       // generic-proxy writes -> visible to the bulk store
        __syncwarp();
        if (lane == 0) { mbar_arrive(&bar_done[s]); mbar_arrive(&bar_ofree[o]); }
        if (lane == 0 && warp == 0) do { if (k < 6) TRACE(k * 6 + 4); } while (0);
    }
    if (lane == 0 && warp == 0) TRACE(63);
}

// Cursor slots of the dynamic tile scheduler: {next block, CTAs finished} per launch in flight.  A launch takes the next slot of
// a ring; the last CTA to finish rewinds the slot's cursor, so a slot is reusable as soon as its launch has ended (launches that
// overlap in time -- different streams -- sit in different slots as long as fewer than kSchedSlots of them are in flight).
constexpr unsigned kSchedSlots = 64, kSchedStride = 8;          // 32 bytes apart
static unsigned int* g_sched[kMaxDevicesK] = {};
static std::mutex g_sched_mu;
static std::atomic<unsigned> g_sched_seq{0};

static cudaError_t sched_slot(int dev, unsigned int** slot)
{
    if (dev < 0 || dev >= kMaxDevicesK) return cudaErrorInvalidDevice;
    if (!g_sched[dev]) {
        std::lock_guard<std::mutex> lk(g_sched_mu);
        if (!g_sched[dev]) {
            unsigned int* p = nullptr;
            cudaError_t e = cudaMalloc(&p, kSchedSlots * kSchedStride * sizeof(unsigned int));
            if (e != cudaSuccess) return e;
            e = cudaMemset(p, 0, kSchedSlots * kSchedStride * sizeof(unsigned int));
            if (e != cudaSuccess) { cudaFree(p); return e; }
            g_sched[dev] = p;
        }
    }
    *slot = g_sched[dev] + (g_sched_seq.fetch_add(1, std::memory_order_relaxed) % kSchedSlots) * kSchedStride;
    return cudaSuccess;
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
    // the kernel counts blocks in 32 bits: larger inputs go in pieces (whole block rows for the row-major RGBA image)
    const uint64_t unit = TARGET == TGT_RGBA ? (uint64_t)bpr : 32u;
    const uint64_t piece_max = ((uint64_t)1 << 31) / unit * unit;
    for (uint64_t done = 0; done < nblocks;) {
        const uint64_t n = nblocks - done < piece_max ? nblocks - done : piece_max;
        unsigned int* slot = nullptr;
        cudaError_t e = sched_slot(dev, &slot);
        if (e != cudaSuccess) return e;
        // one persistent CTA per SM; fewer when the input is small (at least ~one first-round tile each)
        const uint64_t want = (n + C::TILE / 4 - 1) / (C::TILE / 4);
        const unsigned grid = (unsigned)(want < (uint64_t)sm_count ? want : (uint64_t)sm_count);
        const size_t ob = (size_t)C::OB;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid);
        cfg.blockDim = dim3(C::THREADS);
        cfg.dynamicSmemBytes = C::SMEM;
        cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;       // see griddepcontrol.* in the kernel
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        e = cudaLaunchKernelEx(&cfg, uastc_sorted_kernel<TARGET>, in + done, static_cast<void*>(static_cast<unsigned char*>(d_out) + done * ob),
                               (uint32_t)n, bpr, (uint64_t)(index_base + done), d_err, slot);
        if (e != cudaSuccess) return e;
        done += n;
    }
    return cudaSuccess;
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
