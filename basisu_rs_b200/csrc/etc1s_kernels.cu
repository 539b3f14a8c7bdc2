// ETC1S / BasisLZ device side (K2 entropy decode, K3 codebook gather).
//
// K2  etc1s_entropy_decode: ONE WARP PER SLICE.  The slice bitstream is an inherently serial chain
//     (reference src/basis_lz/mod.rs:188-458), so the warp runs the chain redundantly on all lanes
//     (identical state, broadcast shared-memory reads -- no shuffles on the critical path) and uses
//     its 32 lanes for everything that is parallel: staging the compressed bytes into a shared
//     window with coalesced loads, pre-loading the previous row's endpoint indices for the
//     "up / up-left" predictors, and flushing the decoded (endpoint, selector) pairs coalesced.
//     The four Huffman models live in shared memory as 10-bit first-level tables; longer codes
//     fall back to the full flat table (huffman.rs:151) in global memory.
// K3  etc1s_gather_etc1 / etc1s_gather_rgba: one thread per block, pure gather
//     (mod.rs:122-146, :163-181).
#include "etc1s_device.h"

namespace b2bu {

// etc.rs:436-445 ETC1 intensity modifier table
__device__ const int16_t kEtc1Mod[32] = {-8, -2, 2, 8, -17, -5, 5, 17, -29, -9, 9, 29, -42, -13, 13, 42,
                                        -60, -18, 18, 60, -80, -24, 24, 80, -106, -33, 33, 106, -183, -47, 47, 183};

constexpr int kL1Bits = 10;
constexpr int kL1Size = 1 << kL1Bits;
constexpr uint32_t kLong = 0xFFFFFFFFu;          // first-level entry: code longer than kL1Bits
constexpr int kRound = 32;                        // blocks decoded between two cooperative phases
constexpr int kHalfBytes = 1024;                  // the compressed-byte ring is refilled in 1 KiB pieces
constexpr int kRingHalves = 4;                    // 4 KiB ring per warp: the piece being read, the next, one in flight
constexpr int kRingWords = kRingHalves * kHalfBytes / 4;
// A round of 32 blocks consumes well under one piece: per block at most one predictor symbol (16 bits) with
// a 4-bit-chunk VLC (40), a delta (16), a selector symbol (16), a run symbol (16) and a 7-bit-chunk VLC (40).
static_assert(kRound * 144 / 8 <= kHalfBytes, "ring piece too small for one round");

struct WarpShared {
    uint32_t ring[kRingWords];
    uint16_t hist[64];                            // selector history when it fits (it always does for real files)
};

struct BitState { uint64_t buf; int avail; uint32_t nextw; };                     // nextw: absolute word index in the slice

__device__ __forceinline__ void bits_ensure32(BitState& s, const uint32_t* ring)
{
    if (s.avail < 32) {
        s.buf |= (uint64_t)ring[s.nextw & (kRingWords - 1)] << s.avail;
        s.avail += 32;
        s.nextw++;
    }
}
__device__ __forceinline__ void bits_skip(BitState& s, uint32_t n) { s.buf >>= n; s.avail -= (int)n; }

// One Huffman model as the warp sees it: 10-bit first-level table in shared memory for the short codes; longer
// codes are resolved by a warp-parallel canonical decode (lane l tests code length l+1 against its `upper`
// bound, a ballot picks the length, a shuffle fetches the symbol index) so that no code needs a global-memory
// round trip on the serial chain.  `flat` is only read for tables that are not valid prefix codes.
struct HuffView {
    const uint32_t* l1;            // shared
    const uint16_t* syms;          // shared (or global when the symbol arrays do not fit)
    const uint32_t* flat;          // global
    uint32_t upper;                // this lane's length bound
    int32_t base;
    uint32_t max_len;
    bool canon;
};

// huffman.rs:186-198 decode_symbol.  Returns the symbol or 0xFFFFFFFF when no code matches.
__device__ __forceinline__ uint32_t huff_decode(BitState& s, const uint32_t* ring, const HuffView& h, int lane)
{
    bits_ensure32(s, ring);
    const uint32_t e = h.l1[(uint32_t)s.buf & (kL1Size - 1)];
    if (e != kLong) {
        const uint32_t len = e & 31u;
        if (len == 0u) return 0xFFFFFFFFu;
        bits_skip(s, len);
        return e >> 5;
    }
    if (h.canon) {
        const uint32_t v = __brev((uint32_t)s.buf) >> 16;                          // next 16 stream bits, first bit most significant
        const uint32_t bal = __ballot_sync(0xFFFFFFFFu, v < h.upper) & 0xFFFFu;
        if (bal == 0u) return 0xFFFFFFFFu;
        const uint32_t len = (uint32_t)__ffs((int)bal);                            // 1..16
        const uint32_t idx = __shfl_sync(0xFFFFFFFFu, (uint32_t)(h.base + (int32_t)(v >> (15 - (lane & 15)))), (int)len - 1);
        bits_skip(s, len);
        return h.syms[idx];
    }
    const uint32_t f = __ldg(h.flat + ((uint32_t)s.buf & ((1u << h.max_len) - 1u)));
    const uint32_t len = f & 31u;
    if (len == 0u) return 0xFFFFFFFFu;
    bits_skip(s, len);
    return f >> 5;
}

// mod.rs:585-608 decode_vlc.  Returns false when the reference would panic (ofs >= 32).
__device__ __forceinline__ bool vlc_decode(BitState& s, const uint32_t* ring, uint32_t chunk_bits, uint32_t& v)
{
    v = 0;
    uint32_t ofs = 0;
    for (;;) {
        bits_ensure32(s, ring);
        const uint32_t c = (uint32_t)s.buf & ((2u << chunk_bits) - 1u);
        bits_skip(s, chunk_bits + 1);
        v |= (c & ((1u << chunk_bits) - 1u)) << ofs;
        ofs += chunk_bits;
        if ((c >> chunk_bits) == 0u) return true;
        if (ofs >= 32u) return false;
    }
}

// One 1 KiB piece of the slice's bytes -> two uint4 per lane (zeros past the end: bitreader.rs:44,55).
__device__ __forceinline__ void piece_load(const uint8_t* __restrict__ data, uint64_t data_len, uint32_t piece, int lane, uint4 (&r)[2])
{
#pragma unroll
    for (int k = 0; k < 2; k++) {
        const uint64_t o = (uint64_t)piece * kHalfBytes + 16u * (uint32_t)(lane + 32 * k);
        if (o + 16 <= data_len) r[k] = __ldg(reinterpret_cast<const uint4*>(data + o));        // slices start 16-byte aligned
        else {
            uint32_t w[4] = {0u, 0u, 0u, 0u};
            for (int b = 0; b < 16; b++) if (o + b < data_len) w[b >> 2] |= (uint32_t)data[o + b] << (8 * (b & 3));
            r[k] = make_uint4(w[0], w[1], w[2], w[3]);
        }
    }
}
__device__ __forceinline__ void piece_store(uint32_t* ring, uint32_t piece, int lane, const uint4 (&r)[2])
{
    uint4* dst = reinterpret_cast<uint4*>(ring + (piece % kRingHalves) * (kHalfBytes / 4));
    dst[lane] = r[0];
    dst[lane + 32] = r[1];
}

// rows_in_smem: the per-slice row state (endpoint indices and predictor bits of the previous row) lives in shared
// memory when the slice is at most `row_cap` blocks wide, else in the global scratch area (very wide slices).
__global__ void __launch_bounds__(512) etc1s_entropy_decode_kernel(Etc1sDecodeParams P, uint32_t row_cap, uint32_t sym_smem_bytes)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint32_t* l1s = reinterpret_cast<uint32_t*>(smem_raw);                      // 4 tables x 1024 entries
    const int nwarps = blockDim.x >> 5;
    uint16_t* syms_s = reinterpret_cast<uint16_t*>(smem_raw + 4 * kL1Size * sizeof(uint32_t));   // sorted symbols (when they fit)
    WarpShared* wsh_all = reinterpret_cast<WarpShared*>(smem_raw + 4 * kL1Size * sizeof(uint32_t) + sym_smem_bytes);
    uint16_t* rows_all = reinterpret_cast<uint16_t*>(wsh_all + nwarps);           // per warp: row_cap endpoint indices + row_cap/2 pred bytes
    for (int i = threadIdx.x; i < 4 * kL1Size; i += blockDim.x) l1s[i] = P.l1[i];
    if (sym_smem_bytes) for (uint32_t i = threadIdx.x; i < (P.sym_ofs[4] + 1u) / 2u; i += blockDim.x)
        reinterpret_cast<uint32_t*>(syms_s)[i] = reinterpret_cast<const uint32_t*>(P.syms)[i];
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t slice = blockIdx.x * nwarps + warp;
    if (slice >= P.num_slices) return;
    WarpShared& W = wsh_all[warp];
    const Etc1sSliceJob job = P.jobs[slice];
    const uint8_t* __restrict__ data = P.data + job.data_ofs;
    uint32_t* __restrict__ out = P.out_idx + job.out_ofs;
    const uint32_t nbx = job.nbx, nby = job.nby;
    const bool rows_in_smem = nbx <= row_cap;
    // previous row: endpoint index per block (u16) and the predictor bits of every 2x2 group's lower half (u8 per 2 blocks)
    uint16_t* rowep = rows_in_smem ? rows_all + (size_t)warp * (row_cap + row_cap / 2 + 8)
                                   : reinterpret_cast<uint16_t*>(P.scratch + job.scratch_ofs);
    uint8_t* predrow = reinterpret_cast<uint8_t*>(rowep + (rows_in_smem ? row_cap : ((nbx + 7u) & ~7u)));
    uint16_t* hist = P.hist_size <= 64u ? W.hist : reinterpret_cast<uint16_t*>(P.scratch + job.scratch_ofs + etc1s_row_state_bytes(nbx));
    const uint32_t num_endpoints = P.num_endpoints, num_selectors = P.num_selectors, hist_size = P.hist_size;
    const uint32_t rle_sym = (hist_size + num_selectors) & 0xFFFFu;               // mod.rs:220-222 (u16 arithmetic)

    HuffView hv[4];
#pragma unroll
    for (int t = 0; t < 4; t++) {
        hv[t].l1 = l1s + t * kL1Size;
        hv[t].syms = (sym_smem_bytes ? syms_s : P.syms) + P.sym_ofs[t];
        hv[t].flat = P.flat[t];
        hv[t].upper = P.canon[t * 32 + (lane & 15)];
        hv[t].base = (int32_t)P.canon[t * 32 + 16 + (lane & 15)];
        hv[t].max_len = P.max_len[t];
        hv[t].canon = (P.canon_ok >> t) & 1u;
    }
    for (uint32_t i = lane; i < hist_size; i += 32) hist[i] = 0;                   // mod.rs:616-621
    // prime the ring with the first three pieces
    uint32_t loaded = 0;                          // pieces [0, loaded) are (or were) in the ring
    for (; loaded < 3; loaded++) { uint4 r[2]; piece_load(data, job.data_len, loaded, lane, r); piece_store(W.ring, loaded, lane, r); }
    __syncwarp();

    BitState bs;
    bs.buf = ((uint64_t)W.ring[1] << 32) | W.ring[0];
    bs.avail = 64;
    bs.nextw = 2;
    uint32_t rover = hist_size / 2, sel_rle = 0, pred_rep = 0, prev_sym = 0, cur = 0, prev_ep = 0;
    uint32_t err = 0, carry_up = 0;           // carry_up: previous row's endpoint index of block x0-1 (its slot is overwritten by then)

    for (uint32_t y = 0; y < nby && !err; y++) {
        for (uint32_t x0 = 0; x0 < nbx && !err; x0 += kRound) {
            const uint32_t nb = nbx - x0 < (uint32_t)kRound ? nbx - x0 : (uint32_t)kRound;
            // ---- cooperative: start fetching the next ring piece if the reader is about to need it ----
            // reader position in pieces; pieces up to pos+2 must be resident before the next round starts
            const uint32_t pos = (bs.nextw * 4u) / kHalfBytes;
            const bool fetch = loaded < pos + 3u;                                 // warp-uniform
            uint4 pre[2];
            if (fetch) piece_load(data, job.data_len, loaded, lane, pre);
            // row above (x0-1 .. x0+31) into registers: lane l holds block x0 + l - 1, lane 0 of the next round's view via shfl
            uint32_t up_mine = carry_up, up_last = 0;
            if (y > 0) {
                if (lane >= 1 && x0 + lane - 1 < nbx) up_mine = rowep[x0 + lane - 1];
                if (x0 + 31 < nbx) up_last = rowep[x0 + 31];
            }
            carry_up = up_last;
            __syncwarp();                          // everyone has read the old row before it is overwritten below
            uint32_t mine = 0;
            // ---- serial chain, executed redundantly by every lane (mod.rs:245-455) ----
            for (uint32_t b = 0; b < nb; b++) {
                const uint32_t x = x0 + b;
                if ((x & 1u) == 0u) {
                    if ((y & 1u) == 0u) {
                        if (pred_rep != 0u) { pred_rep--; cur = prev_sym; }
                        else {
                            const uint32_t s = huff_decode(bs, W.ring, hv[0], lane);
                            if (s == 0xFFFFFFFFu) { err = ETC1S_ERR_HUFFMAN; break; }
                            if (s == 256u) {
                                uint32_t v;
                                if (!vlc_decode(bs, W.ring, 4, v)) { err = ETC1S_ERR_VLC; break; }
                                pred_rep = v + 3u - 1u;
                                cur = prev_sym;
                            } else { cur = s & 0xFFu; prev_sym = cur; }
                        }
                        if (lane == 0) predrow[x >> 1] = (uint8_t)(cur >> 4);
                    } else cur = predrow[x >> 1];
                }
                const uint32_t pred = cur & 3u;
                cur >>= 2;
                uint32_t ep;
                if (pred == 0u) { if (x == 0u) { err = ETC1S_ERR_PREDICTION; break; } ep = prev_ep; }
                else if (pred == 1u) {
                    if (y == 0u) { err = ETC1S_ERR_PREDICTION; break; }
                    ep = b == 31u ? up_last : __shfl_sync(0xFFFFFFFFu, up_mine, b + 1);
                }
                else if (pred == 2u) {
                    if (P.is_video) ep = 0u;                                      // quirk C-5: previous-frame state is always zero
                    else { if (x == 0u || y == 0u) { err = ETC1S_ERR_PREDICTION; break; } ep = __shfl_sync(0xFFFFFFFFu, up_mine, b); }
                } else {
                    const uint32_t d = huff_decode(bs, W.ring, hv[1], lane);
                    if (d == 0xFFFFFFFFu) { err = ETC1S_ERR_HUFFMAN; break; }
                    ep = (d + prev_ep) & 0xFFFFu;
                    if (ep >= num_endpoints) ep = (ep - num_endpoints) & 0xFFFFu;
                }
                prev_ep = ep;
                uint32_t sel;
                if (!P.is_video || pred != 2u) {
                    uint32_t sym;
                    if (sel_rle > 0u) { sel_rle--; sym = num_selectors; }
                    else {
                        sym = huff_decode(bs, W.ring, hv[2], lane);
                        if (sym == 0xFFFFFFFFu) { err = ETC1S_ERR_HUFFMAN; break; }
                        if (sym == rle_sym) {
                            const uint32_t r = huff_decode(bs, W.ring, hv[3], lane);
                            if (r == 0xFFFFFFFFu) { err = ETC1S_ERR_HUFFMAN; break; }
                            uint32_t cnt = 3u + r;
                            if (r == 63u) {
                                uint32_t v;
                                if (!vlc_decode(bs, W.ring, 7, v)) { err = ETC1S_ERR_VLC; break; }
                                cnt = 3u + v;
                            }
                            sel_rle = cnt - 1u;
                            sym = num_selectors;
                        }
                    }
                    if (sym >= num_selectors) {
                        const uint32_t k = sym - num_selectors;
                        if (hist_size == 0u || k >= hist_size) { err = ETC1S_ERR_PREDICTION; break; }     // asserts mod.rs:404,409
                        sel = hist[k];
                        if (k != 0u) { const uint16_t a = hist[k >> 1]; __syncwarp(); if (lane == 0) { hist[k >> 1] = (uint16_t)sel; hist[k] = a; } __syncwarp(); }
                    } else {
                        sel = sym;
                        if (hist_size > 0u) { if (lane == 0) hist[rover] = (uint16_t)sym; rover++; if (rover == hist_size) rover = hist_size / 2; __syncwarp(); }
                    }
                } else sel = 0u;
                if (ep >= num_endpoints || sel >= num_selectors) { err = ETC1S_ERR_RANGE; break; }        // asserts mod.rs:443-444
                if ((uint32_t)lane == b) mine = ep | (sel << 16);
            }
            // ---- cooperative: flush the decoded pairs, update the row state, land the prefetched ring piece ----
            if (!err && (uint32_t)lane < nb) {
                out[(uint64_t)y * nbx + x0 + lane] = mine;
                rowep[x0 + lane] = (uint16_t)mine;
            }
            if (fetch) { piece_store(W.ring, loaded, lane, pre); loaded++; }
            __syncwarp();
        }
    }
    if (lane == 0) P.status[slice] = err;
}

// K3a: mod.rs:163-181 -- ETC1S block = [R5<<3, G5<<3, B5<<3, inten<<5 | inten<<2 | 3, selector etc1 bytes]
__global__ void __launch_bounds__(256) etc1s_gather_etc1_kernel(const uint32_t* __restrict__ idx, uint64_t nblocks,
                                                                const uint32_t* __restrict__ endpoints, const uint32_t* __restrict__ sel_etc1,
                                                                uint2* __restrict__ out)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nblocks; i += stride) {
        const uint32_t v = idx[i];
        const uint32_t e = __ldg(endpoints + (v & 0xFFFFu));                       // inten | r5 << 8 | g5 << 16 | b5 << 24
        const uint32_t inten = e & 0xFFu;
        const uint32_t lo = (((e >> 8) & 0xFFu) << 3 & 0xFFu) | ((((e >> 16) & 0xFFu) << 3 & 0xFFu) << 8) | ((((e >> 24) & 0xFFu) << 3 & 0xFFu) << 16) |
                            ((((inten << 5) | (inten << 2) | 3u) & 0xFFu) << 24);
        out[i] = make_uint2(lo, __ldg(sel_etc1 + (v >> 16)));
    }
}

// K3b: mod.rs:114-151 -- RGBA image, pitch 4*nbx pixels; the optional alpha slice overwrites A with the G of its colour
__global__ void __launch_bounds__(256) etc1s_gather_rgba_kernel(const uint32_t* __restrict__ idx_rgb, const uint32_t* __restrict__ idx_alpha,
                                                                uint32_t nbx, uint64_t nblocks, const uint32_t* __restrict__ endpoints,
                                                                const uint32_t* __restrict__ sel_plain, uint4* __restrict__ out)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nblocks; i += stride) {
        uint32_t colors[4], alphas[4] = {255u, 255u, 255u, 255u};
        uint32_t rows, arows = 0;
        {
            const uint32_t v = idx_rgb[i];
            const uint32_t e = __ldg(endpoints + (v & 0xFFFFu));
            rows = __ldg(sel_plain + (v >> 16));
            const uint32_t inten = e & 7u;
            int base[3];
#pragma unroll
            for (int c = 0; c < 3; c++) { const uint32_t c5 = (e >> (8 + 8 * c)) & 0xFFu; base[c] = (int)(((c5 << 3) | (c5 >> 2)) & 0xFFu); }   // etc.rs:396-406
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int md = __ldg(&kEtc1Mod[inten * 4 + k]);                                                    // etc.rs:420-431
                uint32_t px = 0xFF000000u;
#pragma unroll
                for (int c = 0; c < 3; c++) { int t = base[c] + md; t = t < 0 ? 0 : t > 255 ? 255 : t; px |= (uint32_t)t << (8 * c); }
                colors[k] = px;
            }
        }
        if (idx_alpha) {
            const uint32_t v = idx_alpha[i];
            const uint32_t e = __ldg(endpoints + (v & 0xFFFFu));
            arows = __ldg(sel_plain + (v >> 16));
            const uint32_t inten = e & 7u;
            const uint32_t g5 = (e >> 16) & 0xFFu;
            const int base = (int)(((g5 << 3) | (g5 >> 2)) & 0xFFu);
#pragma unroll
            for (int k = 0; k < 4; k++) { int t = base + __ldg(&kEtc1Mod[inten * 4 + k]); alphas[k] = (uint32_t)(t < 0 ? 0 : t > 255 ? 255 : t); }
        }
        const uint64_t by = i / nbx;
        const uint32_t bx = (uint32_t)(i - by * nbx);
        uint4* p = out + (by * 4) * nbx + bx;
#pragma unroll
        for (int y = 0; y < 4; y++) {
            const uint32_t r = (rows >> (8 * y)) & 0xFFu, ar = (arows >> (8 * y)) & 0xFFu;
            uint32_t px[4];
#pragma unroll
            for (int x = 0; x < 4; x++) {
                const uint32_t c = colors[(r >> (2 * x)) & 3u];
                px[x] = idx_alpha ? ((c & 0x00FFFFFFu) | (alphas[(ar >> (2 * x)) & 3u] << 24)) : c;
            }
            p[(uint64_t)y * nbx] = make_uint4(px[0], px[1], px[2], px[3]);
        }
    }
}

// shared memory of K2: 4 first-level tables + sorted symbols + per warp {ring, history, previous-row state for row_cap blocks}
static size_t etc1s_decode_smem_bytes(int warps, uint32_t row_cap, uint32_t sym_bytes)
{
    return 4 * kL1Size * sizeof(uint32_t) + sym_bytes + (size_t)warps * (sizeof(WarpShared) + ((size_t)row_cap + row_cap / 2 + 8) * 2);
}

cudaError_t launch_etc1s_decode(const Etc1sDecodeParams& P, int warps_per_cta, uint32_t max_nbx, cudaStream_t stream)
{
    if (P.num_slices == 0) return cudaSuccess;
    const size_t limit = 220 * 1024;
    // sorted symbols in shared memory when they fit in 96 KiB (codebooks up to ~24k entries each), else read from global
    uint32_t sym_bytes = ((P.sym_ofs[4] * 2u) + 15u) & ~15u;
    if (sym_bytes > 96 * 1024) sym_bytes = 0;
    // previous-row state in shared memory when it fits, else in the scratch area
    uint32_t row_cap = (max_nbx + 7u) & ~7u;
    if (etc1s_decode_smem_bytes(1, row_cap, sym_bytes) > limit) row_cap = 0;
    while (warps_per_cta > 1 && etc1s_decode_smem_bytes(warps_per_cta, row_cap, sym_bytes) > limit) warps_per_cta >>= 1;
    static bool configured[16] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 16 && !configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(etc1s_entropy_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)limit);
        if (e != cudaSuccess) return e;
        configured[dev] = true;
    }
    const unsigned grid = (P.num_slices + warps_per_cta - 1) / warps_per_cta;
    etc1s_entropy_decode_kernel<<<grid, 32 * warps_per_cta, etc1s_decode_smem_bytes(warps_per_cta, row_cap, sym_bytes), stream>>>(P, row_cap, sym_bytes);
    return cudaGetLastError();
}

cudaError_t launch_etc1s_gather_etc1(const uint32_t* idx, uint64_t nblocks, const uint32_t* endpoints, const uint32_t* sel_etc1, void* out,
                                     int sm_count, cudaStream_t stream)
{
    if (nblocks == 0) return cudaSuccess;
    const uint64_t want = (nblocks + 255) / 256, cap = (uint64_t)sm_count * 16;
    etc1s_gather_etc1_kernel<<<(unsigned)(want < cap ? want : cap), 256, 0, stream>>>(idx, nblocks, endpoints, sel_etc1, reinterpret_cast<uint2*>(out));
    return cudaGetLastError();
}

cudaError_t launch_etc1s_gather_rgba(const uint32_t* idx_rgb, const uint32_t* idx_alpha, uint32_t nbx, uint64_t nblocks, const uint32_t* endpoints,
                                     const uint32_t* sel_plain, void* out, int sm_count, cudaStream_t stream)
{
    if (nblocks == 0) return cudaSuccess;
    const uint64_t want = (nblocks + 255) / 256, cap = (uint64_t)sm_count * 16;
    etc1s_gather_rgba_kernel<<<(unsigned)(want < cap ? want : cap), 256, 0, stream>>>(idx_rgb, idx_alpha, nbx, nblocks, endpoints, sel_plain,
                                                                                     reinterpret_cast<uint4*>(out));
    return cudaGetLastError();
}

}  // namespace b2bu
