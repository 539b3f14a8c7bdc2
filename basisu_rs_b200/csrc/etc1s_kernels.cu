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
constexpr int kWinWords = 256;                    // compressed-byte window per warp (1 KiB)

struct WarpShared {
    uint32_t win[kWinWords];
    uint32_t stage[kRound];
    uint32_t up[kRound + 1];                      // endpoint indices of the previous row, x0-1 .. x0+31
    uint16_t hist[64];                            // selector history when it fits (it always does for real files)
};

struct BitState { uint64_t buf; int avail; uint32_t nextw; };

__device__ __forceinline__ void bits_ensure32(BitState& s, const uint32_t* win)
{
    if (s.avail < 32) {
        const uint32_t w = s.nextw < (uint32_t)kWinWords ? win[s.nextw] : 0u;
        s.buf |= (uint64_t)w << s.avail;
        s.avail += 32;
        s.nextw++;
    }
}
__device__ __forceinline__ void bits_skip(BitState& s, uint32_t n, uint64_t& consumed) { s.buf >>= n; s.avail -= (int)n; consumed += n; }

// huffman.rs:186-198 decode_symbol.  Returns the symbol or 0xFFFFFFFF when no code matches.
__device__ __forceinline__ uint32_t huff_decode(BitState& s, const uint32_t* win, const uint32_t* l1, const uint32_t* __restrict__ flat,
                                                uint32_t max_len, uint64_t& consumed)
{
    bits_ensure32(s, win);
    uint32_t e = l1[(uint32_t)s.buf & (kL1Size - 1)];
    if (e == kLong) e = __ldg(flat + ((uint32_t)s.buf & ((1u << max_len) - 1u)));
    const uint32_t len = e & 31u;
    if (len == 0u) return 0xFFFFFFFFu;
    bits_skip(s, len, consumed);
    return e >> 5;
}

// mod.rs:585-608 decode_vlc.  Returns false when the reference would panic (ofs >= 32).
__device__ __forceinline__ bool vlc_decode(BitState& s, const uint32_t* win, uint32_t chunk_bits, uint32_t& v, uint64_t& consumed)
{
    v = 0;
    uint32_t ofs = 0;
    for (;;) {
        bits_ensure32(s, win);
        const uint32_t c = (uint32_t)s.buf & ((2u << chunk_bits) - 1u);
        bits_skip(s, chunk_bits + 1, consumed);
        v |= (c & ((1u << chunk_bits) - 1u)) << ofs;
        ofs += chunk_bits;
        if ((c >> chunk_bits) == 0u) return true;
        if (ofs >= 32u) return false;
    }
}

__global__ void __launch_bounds__(128) etc1s_entropy_decode_kernel(Etc1sDecodeParams P)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint32_t* l1s = reinterpret_cast<uint32_t*>(smem_raw);                      // 4 tables x 1024 entries
    WarpShared* wsh_all = reinterpret_cast<WarpShared*>(smem_raw + 4 * kL1Size * sizeof(uint32_t));
    for (int i = threadIdx.x; i < 4 * kL1Size; i += blockDim.x) l1s[i] = P.l1[i];
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t slice = blockIdx.x * (blockDim.x >> 5) + warp;
    if (slice >= P.num_slices) return;
    WarpShared& W = wsh_all[warp];
    const Etc1sSliceJob job = P.jobs[slice];
    const uint8_t* __restrict__ data = P.data + job.data_ofs;
    uint32_t* __restrict__ out = P.out_idx + job.out_ofs;
    uint8_t* predrow = P.scratch + job.scratch_ofs;                               // nbx bytes
    uint16_t* hist = P.hist_size <= 64u ? W.hist : reinterpret_cast<uint16_t*>(P.scratch + job.scratch_ofs + ((job.nbx + 15u) & ~15u));
    const uint32_t nbx = job.nbx, nby = job.nby;
    const uint32_t num_endpoints = P.num_endpoints, num_selectors = P.num_selectors, hist_size = P.hist_size;
    const uint32_t rle_sym = (hist_size + num_selectors) & 0xFFFFu;               // mod.rs:220-222 (u16 arithmetic)

    for (uint32_t i = lane; i < hist_size; i += 32) hist[i] = 0;                   // mod.rs:616-621
    __syncwarp();

    uint64_t consumed = 0;                        // bits consumed so far
    uint32_t rover = hist_size / 2, sel_rle = 0, pred_rep = 0, prev_sym = 0, cur = 0, prev_ep = 0;
    uint32_t err = 0;

    for (uint32_t y = 0; y < nby && !err; y++) {
        for (uint32_t x0 = 0; x0 < nbx && !err; x0 += kRound) {
            const uint32_t nb = nbx - x0 < (uint32_t)kRound ? nbx - x0 : (uint32_t)kRound;
            // ---- cooperative: refill the byte window at the current position, fetch the row above ----
            const uint64_t wbyte = (consumed >> 3) & ~3ull;
#pragma unroll
            for (int k = 0; k < kWinWords / 32; k++) {
                const uint64_t o = wbyte + 4ull * (uint32_t)(lane + 32 * k);
                uint32_t w = 0;
                if (o + 4 <= job.data_len) w = (uint32_t)data[o] | ((uint32_t)data[o + 1] << 8) | ((uint32_t)data[o + 2] << 16) | ((uint32_t)data[o + 3] << 24);
                else for (int b = 0; b < 4; b++) if (o + b < job.data_len) w |= (uint32_t)data[o + b] << (8 * b);   // bitreader.rs:44,55: zeros past the end
                W.win[lane + 32 * k] = w;
            }
            if (y > 0) {
                const uint32_t* above = out + (uint64_t)(y - 1) * nbx;
                if (x0 + lane >= 1 && x0 + lane - 1 < nbx) W.up[lane] = __ldcg(above + x0 + lane - 1) & 0xFFFFu;
                if (lane == 0 && x0 + 31 < nbx) W.up[32] = __ldcg(above + x0 + 31) & 0xFFFFu;
            }
            __syncwarp();
            BitState bs;
            {
                const uint32_t rel = (uint32_t)(consumed - wbyte * 8);            // 0..31
                bs.buf = (((uint64_t)W.win[1] << 32) | W.win[0]) >> rel;
                bs.avail = 64 - (int)rel;
                bs.nextw = 2;
            }
            // ---- serial chain, executed redundantly by every lane (mod.rs:245-455) ----
            for (uint32_t b = 0; b < nb; b++) {
                const uint32_t x = x0 + b;
                if ((x & 1u) == 0u) {
                    if ((y & 1u) == 0u) {
                        if (pred_rep != 0u) { pred_rep--; cur = prev_sym; }
                        else {
                            const uint32_t s = huff_decode(bs, W.win, l1s + 0 * kL1Size, P.flat[0], P.max_len[0], consumed);
                            if (s == 0xFFFFFFFFu) { err = ETC1S_ERR_HUFFMAN; break; }
                            if (s == 256u) {
                                uint32_t v;
                                if (!vlc_decode(bs, W.win, 4, v, consumed)) { err = ETC1S_ERR_VLC; break; }
                                pred_rep = v + 3u - 1u;
                                cur = prev_sym;
                            } else { cur = s & 0xFFu; prev_sym = cur; }
                        }
                        predrow[x] = (uint8_t)(cur >> 4);
                    } else cur = predrow[x];
                }
                const uint32_t pred = cur & 3u;
                cur >>= 2;
                uint32_t ep;
                if (pred == 0u) { if (x == 0u) { err = ETC1S_ERR_PREDICTION; break; } ep = prev_ep; }
                else if (pred == 1u) { if (y == 0u) { err = ETC1S_ERR_PREDICTION; break; } ep = W.up[b + 1]; }
                else if (pred == 2u) {
                    if (P.is_video) ep = 0u;                                      // quirk C-5: previous-frame state is always zero
                    else { if (x == 0u || y == 0u) { err = ETC1S_ERR_PREDICTION; break; } ep = W.up[b]; }
                } else {
                    const uint32_t d = huff_decode(bs, W.win, l1s + 1 * kL1Size, P.flat[1], P.max_len[1], consumed);
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
                        sym = huff_decode(bs, W.win, l1s + 2 * kL1Size, P.flat[2], P.max_len[2], consumed);
                        if (sym == 0xFFFFFFFFu) { err = ETC1S_ERR_HUFFMAN; break; }
                        if (sym == rle_sym) {
                            const uint32_t r = huff_decode(bs, W.win, l1s + 3 * kL1Size, P.flat[3], P.max_len[3], consumed);
                            if (r == 0xFFFFFFFFu) { err = ETC1S_ERR_HUFFMAN; break; }
                            uint32_t cnt = 3u + r;
                            if (r == 63u) {
                                uint32_t v;
                                if (!vlc_decode(bs, W.win, 7, v, consumed)) { err = ETC1S_ERR_VLC; break; }
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
                        if (k != 0u) { const uint16_t a = hist[k >> 1]; __syncwarp(); hist[k >> 1] = (uint16_t)sel; hist[k] = a; __syncwarp(); }
                    } else {
                        sel = sym;
                        if (hist_size > 0u) { hist[rover] = (uint16_t)sym; rover++; if (rover == hist_size) rover = hist_size / 2; __syncwarp(); }
                    }
                } else sel = 0u;
                if (ep >= num_endpoints || sel >= num_selectors) { err = ETC1S_ERR_RANGE; break; }        // asserts mod.rs:443-444
                if (lane == 0) W.stage[b] = ep | (sel << 16);
            }
            __syncwarp();
            // ---- cooperative: flush the decoded pairs ----
            if (!err && (uint32_t)lane < nb) out[(uint64_t)y * nbx + x0 + lane] = W.stage[lane];
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

size_t etc1s_decode_smem_bytes(int warps) { return 4 * kL1Size * sizeof(uint32_t) + (size_t)warps * sizeof(WarpShared); }

cudaError_t launch_etc1s_decode(const Etc1sDecodeParams& P, int warps_per_cta, cudaStream_t stream)
{
    if (P.num_slices == 0) return cudaSuccess;
    const unsigned grid = (P.num_slices + warps_per_cta - 1) / warps_per_cta;
    etc1s_entropy_decode_kernel<<<grid, 32 * warps_per_cta, etc1s_decode_smem_bytes(warps_per_cta), stream>>>(P);
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
