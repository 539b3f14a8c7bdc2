// UASTC 4x4 block front-end + the five target back-ends as mode-specialised device code.
//
// Design (B200-first, not a port): the reference walks every block with a bit-at-a-time reader
// (src/bitreader.rs) and byte-at-a-time writers (src/bitwriter.rs).  Here one thread owns one
// 16-byte block held in four registers; every UASTC mode is a template instantiation whose field
// offsets are compile-time constants, so a field read is one funnel shift + mask, the weight
// stream is handled as a whole 64/96-bit word (SWAR), and the output block is assembled in four
// registers and stored with one 128-bit store.  Tables that are indexed by per-thread data live
// in shared memory (DevTables).
//
// Reference behaviour being reproduced (paths relative to /root/reference/src):
//   uastc.rs:237-327 decode_block_to_rgba     target_formats/astc.rs:8-181
//   target_formats/bc7.rs:9-553               target_formats/etc.rs:11-341
#pragma once
#include <cstdint>
#ifndef B2BU_HOST_EMU   // tests/emu compiles this header for the host with shimmed intrinsics
#include <cuda_runtime.h>
#endif
#include "device_tables.h"

namespace b2bu {

enum { FMT_RGB = 0, FMT_RGBA = 1, FMT_LA = 2 };
enum { TGT_RGBA = 0, TGT_ASTC = 1, TGT_BC7 = 2, TGT_ETC1 = 3, TGT_ETC2 = 4 };
enum { ERR_OK = 0, ERR_LEN = 1, ERR_MODE = 2, ERR_PATTERN = 3 };

#define B2BU_DI __device__ __forceinline__

// ------------------------------------------------------------------------------------------
// Mode descriptors (uastc.rs:528-557) and everything derivable from them at compile time.
// ------------------------------------------------------------------------------------------
template <int M> struct MI;
#define B2BU_MODE(M, CS, R, F, WB, PL, SS, FL)                                                   \
    template <> struct MI<M> {                                                                   \
        static constexpr int code_size = CS, range = R, fmt = F, wbits = WB, planes = PL,       \
                             subsets = SS, flags_bits = FL;                                      \
    };
B2BU_MODE(0, 4, 19, FMT_RGB, 4, 1, 1, 15)
B2BU_MODE(1, 6, 20, FMT_RGB, 2, 1, 1, 15)
B2BU_MODE(2, 5, 8, FMT_RGB, 3, 1, 2, 15)
B2BU_MODE(3, 5, 7, FMT_RGB, 2, 1, 3, 15)
B2BU_MODE(4, 5, 12, FMT_RGB, 2, 1, 2, 15)
B2BU_MODE(5, 5, 20, FMT_RGB, 3, 1, 1, 15)
B2BU_MODE(6, 5, 18, FMT_RGB, 2, 2, 1, 15)
B2BU_MODE(7, 5, 12, FMT_RGB, 2, 1, 2, 15)
B2BU_MODE(9, 5, 8, FMT_RGBA, 2, 1, 2, 23)
B2BU_MODE(10, 3, 13, FMT_RGBA, 4, 1, 1, 17)
B2BU_MODE(11, 2, 13, FMT_RGBA, 2, 2, 1, 17)
B2BU_MODE(12, 3, 19, FMT_RGBA, 3, 1, 1, 17)
B2BU_MODE(13, 5, 20, FMT_RGBA, 1, 2, 1, 23)
B2BU_MODE(14, 5, 20, FMT_RGBA, 2, 1, 1, 23)
B2BU_MODE(15, 7, 20, FMT_LA, 4, 1, 1, 23)
B2BU_MODE(16, 6, 20, FMT_LA, 2, 1, 2, 23)
B2BU_MODE(17, 6, 20, FMT_LA, 2, 2, 1, 23)
B2BU_MODE(18, 4, 11, FMT_RGB, 5, 1, 1, 15)
#undef B2BU_MODE

__host__ __device__ constexpr int range_bits(int r) { return r == 7 ? 2 : r == 8 ? 4 : r == 11 ? 5 : r == 12 ? 3 : r == 13 ? 4 : r == 18 ? 5 : r == 19 ? 6 : 8; }
__host__ __device__ constexpr int range_tq(int r) { return (r == 7 || r == 13 || r == 19) ? 3 : (r == 12 || r == 18) ? 5 : 0; }
__host__ __device__ constexpr int trit_tail_bits(int n) { return n == 0 ? 0 : n == 1 ? 2 : n == 2 ? 4 : n == 3 ? 5 : 7; }
__host__ __device__ constexpr int quint_tail_bits(int n) { return n == 0 ? 0 : n == 1 ? 3 : 5; }

template <int M> struct MD : MI<M> {
    using I = MI<M>;
    static constexpr int NC = I::fmt == FMT_RGB ? 3 : I::fmt == FMT_RGBA ? 4 : 2;   // uastc.rs:470
    static constexpr int N = NC * I::subsets * 2;                                   // uastc.rs:478
    static constexpr int B = range_bits(I::range);
    static constexpr int TQ = range_tq(I::range);
    static constexpr int TQBITS = TQ == 3 ? (N / 5) * 8 + trit_tail_bits(N % 5)
                                : TQ == 5 ? (N / 3) * 7 + quint_tail_bits(N % 3) : 0;
    static constexpr int CSB = (I::planes == 2 && I::fmt != FMT_LA) ? 2 : 0;        // uastc.rs:343
    static constexpr int PB = (M == 7 || I::subsets == 2) ? 5 : I::subsets == 3 ? 4 : 0;   // uastc.rs:352
    static constexpr int PCOUNT = M == 7 ? 19 : I::subsets == 2 ? 30 : I::subsets == 3 ? 11 : 1;
    static constexpr int FPOS = I::code_size;                      // transcoding flags
    static constexpr int CPOS = I::code_size + I::flags_bits;      // component selector
    static constexpr int PPOS = CPOS + CSB;                        // pattern index
    static constexpr int EPOS = PPOS + PB;                         // trit / quint groups
    static constexpr int BPOS = EPOS + TQBITS;                     // raw endpoint bits
    static constexpr int WPOS = BPOS + N * B;                      // weights
    static constexpr int WTOT = 16 * I::planes * I::wbits - I::subsets * I::planes;
    static constexpr int WUNI = 16 * I::planes * I::wbits;         // after re-inserting the anchor MSBs
    static_assert(WPOS + WTOT <= 128, "mode does not fit a block");
    static constexpr bool HAS_ALPHA = I::fmt != FMT_RGB;
};

// ------------------------------------------------------------------------------------------
// 128-bit register helpers.  Every position is a compile-time constant after inlining, so the
// selects below fold to nothing; they also stay correct for run-time positions.
// ------------------------------------------------------------------------------------------
B2BU_DI uint32_t wsel(const uint4& v, int i) { return i == 0 ? v.x : i == 1 ? v.y : i == 2 ? v.z : i == 3 ? v.w : 0u; }

B2BU_DI uint32_t mask_lo(int n) { return n >= 32 ? 0xFFFFFFFFu : n <= 0 ? 0u : ((1u << n) - 1u); }

// bits [pos, pos+len) of v, len <= 32; bits past 127 read as zero (bitreader.rs:37-60)
B2BU_DI uint32_t getbits(const uint4& v, int pos, int len)
{
    if (len <= 0) return 0u;
    const int w = pos >> 5, s = pos & 31;
    const uint32_t lo = wsel(v, w), hi = wsel(v, w + 1);
    const uint32_t r = (s == 0) ? lo : (s + len <= 32) ? (lo >> s) : __funnelshift_r(lo, hi, s);
    return (len >= 32 || s + len == 32) ? r : (r & mask_lo(len));
}

B2BU_DI uint4 shr128(const uint4& v, int k)
{
    const int q = k >> 5, r = k & 31;
    uint4 o;
    o.x = r ? __funnelshift_r(wsel(v, q), wsel(v, q + 1), r) : wsel(v, q);
    o.y = r ? __funnelshift_r(wsel(v, q + 1), wsel(v, q + 2), r) : wsel(v, q + 1);
    o.z = r ? __funnelshift_r(wsel(v, q + 2), wsel(v, q + 3), r) : wsel(v, q + 2);
    o.w = r ? (wsel(v, q + 3) >> r) : wsel(v, q + 3);
    return o;
}
B2BU_DI uint4 shl128(const uint4& v, int k)
{
    const int q = k >> 5, r = k & 31;
    uint4 o;
    o.w = r ? __funnelshift_l(wsel(v, 2 - q), wsel(v, 3 - q), r) : wsel(v, 3 - q);
    o.z = r ? __funnelshift_l(wsel(v, 1 - q), wsel(v, 2 - q), r) : wsel(v, 2 - q);
    o.y = r ? __funnelshift_l(wsel(v, 0 - q), wsel(v, 1 - q), r) : wsel(v, 1 - q);
    o.x = r ? (wsel(v, 0 - q) << r) : wsel(v, 0 - q);
    return o;
}
B2BU_DI uint4 or128(const uint4& a, const uint4& b) { return make_uint4(a.x | b.x, a.y | b.y, a.z | b.z, a.w | b.w); }
B2BU_DI uint4 xor128(const uint4& a, const uint4& b) { return make_uint4(a.x ^ b.x, a.y ^ b.y, a.z ^ b.z, a.w ^ b.w); }
B2BU_DI uint4 and128(const uint4& a, const uint4& b) { return make_uint4(a.x & b.x, a.y & b.y, a.z & b.z, a.w & b.w); }
// low n bits set
B2BU_DI uint4 mask128(int n)
{
    return make_uint4(mask_lo(n), mask_lo(n - 32), mask_lo(n - 64), mask_lo(n - 96));
}
B2BU_DI uint4 u64_to_128(uint64_t v) { return make_uint4((uint32_t)v, (uint32_t)(v >> 32), 0u, 0u); }
B2BU_DI uint4 u32_to_128(uint32_t v) { return make_uint4(v, 0u, 0u, 0u); }

// OR `len` (<= 32) bits of val into out at bit position pos (bitwriter.rs:23-51, overflow dropped)
B2BU_DI void putbits(uint4& out, int pos, int len, uint32_t val)
{
    if (len <= 0) return;
    out = or128(out, shl128(u32_to_128(val & mask_lo(len)), pos));
}

// zero bit inserted at position pos: bits >= pos move up by one
template <typename T> B2BU_DI T insert_zero(T v, uint32_t pos)
{
    const T low = v & (((T)1 << pos) - (T)1);
    return low | ((v ^ low) << 1);
}
// bit at position pos removed: bits > pos move down by one
template <typename T> B2BU_DI T delete_bit(T v, uint32_t pos)
{
    const T low = v & (((T)1 << pos) - (T)1);
    return low | ((v >> (pos + 1)) << pos);
}

// ------------------------------------------------------------------------------------------
// Front-end: header fields, endpoints, weights
// ------------------------------------------------------------------------------------------
template <int M> B2BU_DI uint32_t read_compsel(const uint4& b)                 // uastc.rs:343-350
{
    using D = MD<M>;
    if (D::CSB) return getbits(b, D::CPOS, 2);
    return D::planes == 2 ? 3u : 0u;
}
template <int M> B2BU_DI uint32_t read_pattern(const uint4& b) { return getbits(b, MD<M>::PPOS, MD<M>::PB); }

template <int M> B2BU_DI uint32_t pattern_word(const DevTables& T, uint32_t pat)          // uastc.rs:368-376
{
    if (M == 7) return T.pat23[pat];
    if (MD<M>::subsets == 2) return T.pat2[pat];
    if (MD<M>::subsets == 3) return T.pat3[pat];
    return 0u;
}
template <int M> B2BU_DI uint32_t anchor_byte(const DevTables& T, uint32_t pat)            // uastc.rs:378-385
{
    if (M == 7) return T.anc23[pat];
    if (MD<M>::subsets == 2) return T.anc2[pat];
    if (MD<M>::subsets == 3) return T.anc3[pat];
    return 0u;
}

// uastc.rs:616-695 decode_endpoints: digits (trit/quint) and raw bits per value
template <int M> B2BU_DI void unpack_quant(const uint4& b, const DevTables& T, uint32_t (&m)[MD<M>::N], uint32_t (&d)[MD<M>::N])
{
    using D = MD<M>;
    if (D::TQ == 3) {
#pragma unroll
        for (int g = 0; g < (D::N + 4) / 5; g++) {
            const int cnt = (D::N - 5 * g) < 5 ? (D::N - 5 * g) : 5;
            const int nb = cnt == 5 ? 8 : trit_tail_bits(cnt);
            const uint32_t x = T.trit_dec[getbits(b, D::EPOS + 8 * g, nb)];
#pragma unroll
            for (int k = 0; k < cnt; k++) d[5 * g + k] = (x >> (2 * k)) & 3u;
        }
    } else if (D::TQ == 5) {
#pragma unroll
        for (int g = 0; g < (D::N + 2) / 3; g++) {
            const int cnt = (D::N - 3 * g) < 3 ? (D::N - 3 * g) : 3;
            const int nb = cnt == 3 ? 7 : quint_tail_bits(cnt);
            const uint32_t x = T.quint_dec[getbits(b, D::EPOS + 7 * g, nb)];
#pragma unroll
            for (int k = 0; k < cnt; k++) d[3 * g + k] = (x >> (3 * k)) & 7u;
        }
    } else {
#pragma unroll
        for (int i = 0; i < D::N; i++) d[i] = 0u;
    }
#pragma unroll
    for (int i = 0; i < D::N; i++) m[i] = getbits(b, D::BPOS + D::B * i, D::B);
}

// uastc.rs:585-614 unquant_endpoint
template <int M> B2BU_DI uint32_t unquant(const DevTables& T, uint32_t d, uint32_t m)
{
    constexpr int R = MD<M>::range;
    if (R == 20) return m;
    if (R == 8) return m * 17u;
    if (R == 11) return (m << 3) | (m >> 2);
    const uint32_t idx = (d << MD<M>::B) | m;
    if (R == 7) return T.unq7[idx];
    if (R == 12) return T.unq12[idx];
    if (R == 13) return T.unq13[idx];
    if (R == 18) return T.unq18[idx];
    return T.unq19[idx];
}

// uastc.rs:721-740 decode_weights, returned as ONE uniform LSB-first stream: 16*planes fields of
// wbits bits (texel-major, plane-minor) with the anchors' missing MSB re-inserted as zero.
template <int M> B2BU_DI uint4 uniform_weights(const uint4& b, const DevTables& T, uint32_t pat)
{
    using D = MD<M>;
    constexpr int wb = D::wbits;
    uint4 R = and128(shr128(b, D::WPOS), mask128(D::WTOT));
    if (D::subsets == 1) {
        if (D::planes == 1) {
            const uint4 low = and128(R, mask128(wb - 1));
            return or128(low, shl128(shr128(R, wb - 1), wb));
        } else {
            const uint4 f0 = and128(R, mask128(wb - 1));
            const uint4 f1 = shl128(and128(shr128(R, wb - 1), mask128(wb - 1)), wb);
            return or128(or128(f0, f1), shl128(shr128(R, 2 * wb - 2), 2 * wb));
        }
    } else {
        const uint32_t ab = anchor_byte<M>(T, pat);
        const uint32_t a1 = ab & 15u, a2 = ab >> 4;
        if (D::WUNI <= 32) {
            uint32_t v = R.x;
            v = insert_zero<uint32_t>(v, wb - 1);
            v = insert_zero<uint32_t>(v, a1 * wb + wb - 1);
            if (D::subsets == 3) v = insert_zero<uint32_t>(v, a2 * wb + wb - 1);
            return u32_to_128(v);
        } else {
            uint64_t v = (uint64_t)R.x | ((uint64_t)R.y << 32);
            v = insert_zero<uint64_t>(v, wb - 1);
            v = insert_zero<uint64_t>(v, a1 * wb + wb - 1);
            if (D::subsets == 3) v = insert_zero<uint64_t>(v, a2 * wb + wb - 1);
            return u64_to_128(v);
        }
    }
}

// uastc.rs:697-719 unquant_weights: ASTC rule = replicate to 6 bits, +1 above 32
template <int WB> B2BU_DI uint32_t unquant_weight(uint32_t w)
{
    uint32_t r;
    if (WB == 1) r = w * 63u;
    else if (WB == 2) r = w * 21u;
    else if (WB == 3) r = w * 9u;
    else if (WB == 4) r = (w << 2) | (w >> 2);
    else r = (w << 1) | (w >> 4);
    return r + (r >> 5);
}

template <int M> B2BU_DI void unpack_endpoints(const uint4& b, const DevTables& T, uint32_t (&e)[MD<M>::N])
{
    uint32_t m[MD<M>::N], d[MD<M>::N];
    unpack_quant<M>(b, T, m, d);
#pragma unroll
    for (int i = 0; i < MD<M>::N; i++) e[i] = unquant<M>(T, d[i], m[i]);
}

// ------------------------------------------------------------------------------------------
// uastc.rs:237-327 decode_block_to_rgba, split in two so that the expensive half is ONE piece of
// code shared by every mode (instruction-cache footprint: the 18 specialised copies of the texel
// loop were 150 KB and thrashed the 32 KB L1.5 I-cache):
//   canon_front<M>   mode-specialised, small: endpoints -> per-channel multiplier / offset words,
//                    weights -> one byte per texel, already unquantised to 0..64 (SWAR on the
//                    whole weight stream)
//   interp_rows<>    mode-independent: 16 texels of ASTC interpolation from the canonical block
//
// The interpolation (uastc.rs:218-235 with srgb = false) is  ((l*257)*(64-w) + (h*257)*w + 32) >> 14.
// Written as  l*257*64 + 32 + (h-l)*257*w  and scaled by 4 it is ONE multiply-add per channel and texel,
//   r = D*w + L,   D = (h - l) * 1028 (two's complement),   L = l * 65792 + 128,
// whose byte 2 (bits 16-23) is the result: r = 4 * (numerator), numerator >> 14 <= 255.  The multiply-adds
// issue on the fma pipe, which the bit-twiddling rest of the kernel leaves mostly idle; the alu pipe only
// sees the weight-byte extract and three byte permutes per texel.
// ------------------------------------------------------------------------------------------
struct Canon {
    uint32_t d[3][4], l[3][4];     // per subset and output channel R, G, B, A.  Dual plane (always one subset): d[1] is the
                                   // multiplier of the plane-1 weight (non-zero for the selected channel only, where d[0] is zero)
    uint4 w0, w1;                  // unquantised weights of plane 0 / 1, one byte per texel
    uint32_t pw;                   // subset map, 2 bits per texel
};

// ASTC weight unquantisation (uastc.rs:697-719) on four byte lanes at once
template <int WB> B2BU_DI uint32_t unquant_weight4(uint32_t w)
{
    uint32_t r;
    if (WB == 1) return w << 6;                                        // 0 -> 0, 1 -> 63 + 1
    else if (WB == 2) r = w * 21u;
    else if (WB == 3) r = w * 9u;
    else if (WB == 4) r = (w << 2) | ((w >> 2) & 0x03030303u);
    else r = (w << 1) | ((w >> 4) & 0x01010101u);
    return r + ((r >> 5) & 0x01010101u);
}

// four consecutive texels (4*j .. 4*j+3) of plane p of the uniform stream U -> four byte lanes
template <int WB, int PLANES> B2BU_DI uint32_t weight_bytes4(const uint4& U, int j, int p)
{
    uint32_t t;
    if (PLANES == 1) {
        t = getbits(U, 4 * WB * j, 4 * WB);
        if (WB == 2) { t = (t | (t << 12)) & 0x000F000Fu; t = (t | (t << 6)) & 0x03030303u; }
        else if (WB == 3) { t = (t | (t << 10)) & 0x003F003Fu; t = (t | (t << 5)) & 0x07070707u; }
        else if (WB == 4) { t = (t | (t << 8)) & 0x00FF00FFu; t = (t | (t << 4)) & 0x0F0F0F0Fu; }
        else { t = (t & 0x3FFu) | ((t << 6) & 0x03FF0000u); t = (t & 0x001F001Fu) | ((t << 3) & 0x1F001F00u); }   // WB == 5
    } else {
        // texel-major, plane-minor: texel i of plane p sits at bit (2 * i + p) * WB
        t = getbits(U, 8 * WB * j + p * WB, 8 * WB - WB);
        if (WB == 1) { t &= 0x55u; t = (t | (t << 12)) & 0x00050005u; t = (t | (t << 6)) & 0x01010101u; }
        else { t &= 0x3333u; t = (t | (t << 8)) & 0x00330033u; t = (t | (t << 4)) & 0x03030303u; }   // WB == 2
    }
    return unquant_weight4<WB>(t);
}

// multiplier / offset words of one channel with endpoints l, h (see above)
B2BU_DI uint32_t lerp_mul(uint32_t l, uint32_t h) { return (h - l) * 1028u; }
B2BU_DI uint32_t lerp_off(uint32_t l) { return l * 65792u + 128u; }

template <int M> B2BU_DI void canon_front(const uint4& b, const DevTables& T, uint32_t pat, uint32_t compsel, Canon& c)
{
    using D = MD<M>;
    uint32_t e[D::N];
    unpack_endpoints<M>(b, T, e);
    // uastc.rs:176-216 assemble_endpoint_pairs: RGB -> A = 255, LA -> L replicated to R, G, B
#pragma unroll
    for (int s = 0; s < D::subsets; s++) {
        const int o = s * D::NC * 2;
        if (D::fmt == FMT_LA) {
            c.d[s][0] = c.d[s][1] = c.d[s][2] = lerp_mul(e[o], e[o + 1]); c.l[s][0] = c.l[s][1] = c.l[s][2] = lerp_off(e[o]);
            c.d[s][3] = lerp_mul(e[o + 2], e[o + 3]); c.l[s][3] = lerp_off(e[o + 2]);
        } else {
#pragma unroll
            for (int ch = 0; ch < 3; ch++) { c.d[s][ch] = lerp_mul(e[o + 2 * ch], e[o + 2 * ch + 1]); c.l[s][ch] = lerp_off(e[o + 2 * ch]); }
            if (D::fmt == FMT_RGBA) { c.d[s][3] = lerp_mul(e[o + 6], e[o + 7]); c.l[s][3] = lerp_off(e[o + 6]); }
            else { c.d[s][3] = 0u; c.l[s][3] = lerp_off(255u); }
        }
    }
    const uint4 U = uniform_weights<M>(b, T, pat);
    c.w0 = make_uint4(weight_bytes4<D::wbits, D::planes>(U, 0, 0), weight_bytes4<D::wbits, D::planes>(U, 1, 0),
                      weight_bytes4<D::wbits, D::planes>(U, 2, 0), weight_bytes4<D::wbits, D::planes>(U, 3, 0));
    if (D::planes == 2) {
        c.w1 = make_uint4(weight_bytes4<D::wbits, D::planes>(U, 0, 1), weight_bytes4<D::wbits, D::planes>(U, 1, 1),
                          weight_bytes4<D::wbits, D::planes>(U, 2, 1), weight_bytes4<D::wbits, D::planes>(U, 3, 1));
        // the channel `compsel` follows plane 1 (uastc.rs:296-313): its multiplier moves to d[1]
#pragma unroll
        for (int ch = 0; ch < 4; ch++) { const bool sel = compsel == (uint32_t)ch; c.d[1][ch] = sel ? c.d[0][ch] : 0u; c.d[0][ch] = sel ? 0u : c.d[0][ch]; }
    }
    c.pw = pattern_word<M>(T, pat);
}

// r[ch] = d[ch] * w + l[ch] for the four channels when `on` is non-zero: four PREDICATED multiply-adds under one predicate.
// (Left to the compiler, `if (subset == 1) r = ...` becomes a divergent branch per texel.)
B2BU_DI void mad4_if(uint32_t (&r)[4], uint32_t on, const uint32_t (&d)[4], uint32_t w, const uint32_t (&l)[4])
{
#ifdef B2BU_HOST_EMU
    if (on) for (int ch = 0; ch < 4; ch++) r[ch] = d[ch] * w + l[ch];
#else
    asm("{\n.reg .pred p;\nsetp.ne.u32 p, %4, 0;\n"
        "@p mad.lo.u32 %0, %5, %9, %10;\n@p mad.lo.u32 %1, %6, %9, %11;\n@p mad.lo.u32 %2, %7, %9, %12;\n@p mad.lo.u32 %3, %8, %9, %13;\n}\n"
        : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3])
        : "r"(on), "r"(d[0]), "r"(d[1]), "r"(d[2]), "r"(d[3]), "r"(w), "r"(l[0]), "r"(l[1]), "r"(l[2]), "r"(l[3]));
#endif
}

// One row of four texels.  NSUB: subsets (1..3); DUAL: two weight planes (one subset); ALPHA = false: the caller never looks at
// byte 3 of a texel (UASTC -> ETC1), so the alpha multiply-add and its byte permute are left out (byte 3 = 0).
// (Extracting the weight byte with a byte dot product -- IDP.4A issues on the fma pipe, tools/probe_pipes.cu -- instead of a byte
// permute was measured: RGBA 94 -> 97 us, so the permute stays.)
template <int NSUB, bool DUAL, bool ALPHA>
B2BU_DI uint4 interp_row(const Canon& c, uint32_t wr0, uint32_t wr1, uint32_t pwr)
{
    uint32_t px[4];
#pragma unroll
    for (int x = 0; x < 4; x++) {
        const uint32_t w = __byte_perm(wr0, 0u, 0x4440 + x);
        uint32_t r[4];
#pragma unroll
        for (int ch = 0; ch < 4; ch++) r[ch] = c.d[0][ch] * w + c.l[0][ch];
        if (DUAL) {
            const uint32_t v = __byte_perm(wr1, 0u, 0x4440 + x);
#pragma unroll
            for (int ch = 0; ch < 4; ch++) r[ch] += c.d[1][ch] * v;
        }
        // subsets 1 / 2 (the map holds 0..2: bit 0 / bit 1 of the texel's field) overwrite the result of subset 0
        if (NSUB >= 2) mad4_if(r, pwr & (1u << (2 * x)), c.d[1], w, c.l[1]);
        if (NSUB == 3) mad4_if(r, pwr & (2u << (2 * x)), c.d[2], w, c.l[2]);
        // byte 2 of every channel word -> R, G, B, A
        if (ALPHA) px[x] = __byte_perm(__byte_perm(r[0], r[1], 0x0062), __byte_perm(r[2], r[3], 0x0062), 0x5410);
        else px[x] = __byte_perm(__byte_perm(r[0], r[1], 0x3362), r[2], 0x7610);       // byte 3 of a channel word is zero
    }
    return make_uint4(px[0], px[1], px[2], px[3]);
}

// Row sinks: where the four pixel rows of a block go.  ROLLED sinks let the row loop stay a loop
// (4x less code); the register-array sink needs it unrolled.
struct PxArraySink {
    static constexpr bool ROLLED = false;
    uint32_t* px;
    B2BU_DI void row(int y, const uint4& v) { px[4 * y] = v.x; px[4 * y + 1] = v.y; px[4 * y + 2] = v.z; px[4 * y + 3] = v.w; }
};
struct StridedRowSink {                    // rows `stride` uint4 apart (shared staging buffer or the global image)
    static constexpr bool ROLLED = true;
    uint4* p;
    uint64_t stride;
    B2BU_DI void row(int y, const uint4& v) { p[(uint64_t)y * stride] = v; }
};

// Tile staging of the RGBA kernel: pixel row 0 of a block replaces the block's own 16 input bytes (the thread has read them),
// rows 1-3 go to a [3][stride] array -- 48 instead of 64 staged bytes per block, so a tile holds a third more blocks.
struct TileRowSink {
    static constexpr bool ROLLED = true;
    uint4* row0;
    uint4* rows123;
    uint64_t stride;
    B2BU_DI void row(int y, const uint4& v) { if (y == 0) *row0 = v; else rows123[(uint64_t)(y - 1) * stride] = v; }
};

template <int NSUB, bool DUAL, bool ALPHA, class Sink> B2BU_DI void interp_rows(Canon& c, Sink& sink)
{
    if (Sink::ROLLED) {
#pragma unroll 1
        for (int y = 0; y < 4; y++) {
            sink.row(y, interp_row<NSUB, DUAL, ALPHA>(c, c.w0.x, c.w1.x, c.pw));
            c.w0.x = c.w0.y; c.w0.y = c.w0.z; c.w0.z = c.w0.w;
            if (DUAL) { c.w1.x = c.w1.y; c.w1.y = c.w1.z; c.w1.z = c.w1.w; }
            c.pw >>= 8;
        }
    } else {
        sink.row(0, interp_row<NSUB, DUAL, ALPHA>(c, c.w0.x, c.w1.x, c.pw));
        sink.row(1, interp_row<NSUB, DUAL, ALPHA>(c, c.w0.y, c.w1.y, c.pw >> 8));
        sink.row(2, interp_row<NSUB, DUAL, ALPHA>(c, c.w0.z, c.w1.z, c.pw >> 16));
        sink.row(3, interp_row<NSUB, DUAL, ALPHA>(c, c.w0.w, c.w1.w, c.pw >> 24));
    }
}

// ---- SWAR helpers for the weight streams (BC7) ----------------------------------------------
// 8 fields of 2 bits (16 bits) -> 8 nibble-spaced fields (32 bits)
B2BU_DI uint32_t spread2to4(uint32_t x)
{
    x = (x | (x << 8)) & 0x00FF00FFu;
    x = (x | (x << 4)) & 0x0F0F0F0Fu;
    return (x | (x << 2)) & 0x33333333u;
}
// 8 fields of 3 bits (24 bits) -> 8 nibble-spaced fields (32 bits)
B2BU_DI uint32_t spread3to4(uint32_t x)
{
    x = (x & 0x00000FFFu) | ((x & 0x00FFF000u) << 4);
    x = (x & 0x003F003Fu) | ((x & 0x0FC00FC0u) << 2);
    return (x & 0x07070707u) | ((x & 0x38383838u) << 1);
}
// the 8 two-bit fields at nibble positions of x (bits 4i, 4i+1) -> 16 contiguous bits
B2BU_DI uint32_t compress4to2(uint32_t x)
{
    x &= 0x33333333u;
    x = (x | (x >> 2)) & 0x0F0F0F0Fu;
    x = (x | (x >> 4)) & 0x00FF00FFu;
    return (x | (x >> 8)) & 0x0000FFFFu;
}

constexpr uint32_t kTwoSubsetModes = (1u << 2) | (1u << 4) | (1u << 7) | (1u << 9) | (1u << 16);
constexpr uint32_t kDualPlaneModes = (1u << 6) | (1u << 11) | (1u << 13) | (1u << 17);

template <bool ALPHA, class Sink> B2BU_DI void interp_block(uint32_t mode, Canon& c, Sink& sink)
{
    if ((kDualPlaneModes >> mode) & 1u) interp_rows<1, true, ALPHA>(c, sink);
    else if (mode == 3u) interp_rows<3, false, ALPHA>(c, sink);
    else if ((kTwoSubsetModes >> mode) & 1u) interp_rows<2, false, ALPHA>(c, sink);
    else interp_rows<1, false, ALPHA>(c, sink);
}

// void-extent colour (uastc.rs:387-394): bits 5..36
B2BU_DI uint32_t mode8_rgba(const uint4& b) { return getbits(b, 5, 32); }

// ------------------------------------------------------------------------------------------
// ASTC back-end (target_formats/astc.rs:8-181)
// ------------------------------------------------------------------------------------------
B2BU_DI uint4 astc_void_extent(uint32_t rgba)                                   // astc.rs:17-43
{
    const uint32_t r = rgba & 0xFFu, g = (rgba >> 8) & 0xFFu, bl = (rgba >> 16) & 0xFFu, a = rgba >> 24;
    return make_uint4(0xFFFFFDFCu, 0xFFFFFFFFu, (r * 257u) | ((g * 257u) << 16), (bl * 257u) | ((a * 257u) << 16));
}

__host__ __device__ constexpr uint32_t astc_block_mode(int m)                                       // astc.rs:333-354
{
    return m == 0 ? 0x0242 : m == 1 ? 0x0042 : m == 2 ? 0x0853 : m == 3 ? 0x1042 : m == 4 ? 0x0842 : m == 5 ? 0x0053
         : m == 6 ? 0x0442 : m == 7 ? 0x0842 : m == 9 ? 0x0842 : m == 10 ? 0x0242 : m == 11 ? 0x0442 : m == 12 ? 0x0053
         : m == 13 ? 0x0441 : m == 14 ? 0x0042 : m == 15 ? 0x0242 : m == 16 ? 0x0842 : m == 17 ? 0x0442 : m == 18 ? 0x0253 : 0;
}

// Mask with WB ones at every texel of `subset`, from the 2-bit-per-texel subset map.  WB == 2 is
// pure arithmetic on the map; WB == 3 (UASTC mode 2 only) uses the pre-expanded 2-subset table.
B2BU_DI uint32_t subset_mask2(uint32_t pw, uint32_t subset)
{
    const uint32_t lo = pw & 0x55555555u, hi = (pw >> 1) & 0x55555555u;      // subset bit 0 / bit 1 per texel
    const uint32_t sel = subset == 0u ? (~(lo | hi) & 0x55555555u) : subset == 1u ? lo : hi;
    return sel * 3u;
}

template <int M> B2BU_DI uint4 astc_block(const uint4& b, const DevTables& T, uint32_t pat, uint32_t compsel)
{
    using D = MD<M>;
    uint32_t m[D::N], d[D::N];
    unpack_quant<M>(b, T, m, d);

    // astc.rs:55-78: avoid blue contraction -- swap each endpoint pair of a subset whose first
    // endpoint has the larger R+G+B, and invert that subset's weights
    bool inv[3] = {false, false, false};
    if (D::fmt != FMT_LA) {
#pragma unroll
        for (int s = 0; s < D::subsets; s++) {
            const int o = s * D::NC * 2;
            const uint32_t s0 = unquant<M>(T, d[o + 0], m[o + 0]) + unquant<M>(T, d[o + 2], m[o + 2]) + unquant<M>(T, d[o + 4], m[o + 4]);
            const uint32_t s1 = unquant<M>(T, d[o + 1], m[o + 1]) + unquant<M>(T, d[o + 3], m[o + 3]) + unquant<M>(T, d[o + 5], m[o + 5]);
            inv[s] = s0 > s1;
#pragma unroll
            for (int c = 0; c < D::NC; c++) {
                const uint32_t ma = m[o + 2 * c], mb = m[o + 2 * c + 1], da = d[o + 2 * c], db = d[o + 2 * c + 1];
                m[o + 2 * c] = inv[s] ? mb : ma; m[o + 2 * c + 1] = inv[s] ? ma : mb;
                d[o + 2 * c] = inv[s] ? db : da; d[o + 2 * c + 1] = inv[s] ? da : db;
            }
        }
    }

    uint4 out = make_uint4(0u, 0u, 0u, 0u);
    int pos = 0;
    // astc.rs:80-96: block mode, partition seed + "same CEM" bits, CEM
    constexpr uint32_t cem = D::fmt == FMT_RGB ? 8u : D::fmt == FMT_RGBA ? 12u : 4u;
    uint32_t hdr = astc_block_mode(M);
    pos = 13;
    if (D::PB) {
        const uint32_t seed = M == 7 ? T.seed23[pat] : D::subsets == 2 ? T.seed2[pat] : T.seed3[pat];
        hdr |= seed << 13;
        pos += 12;
    }
    hdr |= cem << pos;
    pos += 4;
    out.x = hdr;

    // astc.rs:98-141: BISE stream, raw bits interleaved with the T / Q bits
    if (D::TQ == 5) {
#pragma unroll
        for (int g = 0; g < (D::N + 2) / 3; g++) {
            const int cnt = (D::N - 3 * g) < 3 ? (D::N - 3 * g) : 3;
            uint32_t id = 0;
#pragma unroll
            for (int k = cnt - 1; k >= 0; k--) id = id * 5u + d[3 * g + k];
            const uint32_t q = T.quint_enc[id];
            // 3B + 7 <= 22 bits per group
            uint32_t grp = m[3 * g] | ((q & 7u) << D::B);
            if (cnt > 1) grp |= m[3 * g + 1] << (D::B + 3);
            grp |= ((q >> 3) & 3u) << (2 * D::B + 3);
            if (cnt > 2) grp |= m[3 * g + 2] << (2 * D::B + 5);
            grp |= ((q >> 5) & 3u) << (3 * D::B + 5);
            putbits(out, pos, 3 * D::B + 7, grp);
            pos += 3 * D::B + 7;
        }
    } else if (D::TQ == 3) {
#pragma unroll
        for (int g = 0; g < (D::N + 4) / 5; g++) {
            const int cnt = (D::N - 5 * g) < 5 ? (D::N - 5 * g) : 5;
            uint32_t id = 0;
#pragma unroll
            for (int k = cnt - 1; k >= 0; k--) id = id * 3u + d[5 * g + k];
            const uint32_t t = T.trit_enc[id];
            // 5B + 8 <= 38 bits per group
            uint64_t grp = (uint64_t)m[5 * g] | ((uint64_t)(t & 3u) << D::B);
            if (cnt > 1) grp |= (uint64_t)m[5 * g + 1] << (D::B + 2);
            grp |= (uint64_t)((t >> 2) & 3u) << (2 * D::B + 2);
            if (cnt > 2) grp |= (uint64_t)m[5 * g + 2] << (2 * D::B + 4);
            grp |= (uint64_t)((t >> 4) & 1u) << (3 * D::B + 4);
            if (cnt > 3) grp |= (uint64_t)m[5 * g + 3] << (3 * D::B + 5);
            grp |= (uint64_t)((t >> 5) & 3u) << (4 * D::B + 5);
            if (cnt > 4) grp |= (uint64_t)m[5 * g + 4] << (4 * D::B + 7);
            grp |= (uint64_t)((t >> 7) & 1u) << (5 * D::B + 7);
            out = or128(out, shl128(u64_to_128(grp), pos));
            pos += 5 * D::B + 8;
        }
    } else {
#pragma unroll
        for (int i = 0; i < D::N; i++) { putbits(out, pos, D::B, m[i]); pos += D::B; }
    }

    // astc.rs:143-178: weights fill from bit 127 downward, each bit-reversed == the uniform
    // LSB-first stream reversed as a whole; CCS follows un-reversed
    uint4 U = uniform_weights<M>(b, T, pat);
    if (D::subsets == 1) {
        if (inv[0]) U = xor128(U, mask128(D::WUNI));
    } else {
        if (D::wbits == 3) {
            // mode 2: two subsets, 3-bit weights; pat2w3 = texels of subset 1 expanded to 3 bits each
            const uint64_t m1 = T.pat2w3[pat];
            uint64_t x = 0;
            if (inv[0]) x |= ~m1 & 0xFFFFFFFFFFFFull;
            if (inv[1]) x |= m1;
            U = xor128(U, u64_to_128(x));
        } else {
            const uint32_t pw = pattern_word<M>(T, pat);
            uint32_t x = 0;
#pragma unroll
            for (int s = 0; s < D::subsets; s++)
                if (inv[s]) x |= subset_mask2(pw, (uint32_t)s);
            U.x ^= x;
        }
    }
    const uint4 Wrev = make_uint4(__brev(U.w), __brev(U.z), __brev(U.y), __brev(U.x));
    out = or128(out, Wrev);
    if (D::planes == 2) putbits(out, 128 - D::WUNI - 2, 2, compsel);
    return out;
}

// ------------------------------------------------------------------------------------------
// BC7 back-end (target_formats/bc7.rs:9-553)
// ------------------------------------------------------------------------------------------
__host__ __device__ constexpr int bc7_mode_for(int m)                                               // bc7.rs:582-589
{
    return m == 0 ? 6 : m == 1 ? 3 : m == 2 ? 1 : m == 3 ? 2 : m == 4 ? 3 : m == 5 ? 6 : m == 6 ? 5 : m == 7 ? 2
         : m == 9 ? 7 : m == 10 ? 6 : m == 11 ? 5 : m == 12 ? 6 : m == 13 ? 5 : m == 14 ? 6 : m == 15 ? 6 : m == 16 ? 7
         : m == 17 ? 5 : 6;
}
struct Bc7Info { int pat_bits, color_bits, alpha_bits, weight_bits, planes, subsets, p_bits, sp_bits, channels; };
__host__ __device__ constexpr Bc7Info bc7_info(int bm)                                              // bc7.rs:570-579
{
    return bm == 1 ? Bc7Info{6, 6, 0, 3, 1, 2, 0, 1, 3} : bm == 2 ? Bc7Info{6, 5, 0, 2, 1, 3, 0, 0, 3}
         : bm == 3 ? Bc7Info{6, 7, 0, 2, 1, 2, 1, 0, 3} : bm == 5 ? Bc7Info{0, 7, 8, 2, 2, 1, 0, 0, 4}
         : bm == 6 ? Bc7Info{0, 7, 7, 4, 1, 1, 1, 0, 4} : Bc7Info{6, 5, 5, 2, 1, 2, 1, 0, 4};
}

// bc7.rs:312-375 + :18-59: void-extent -> BC7 mode 6 (lossless unless a channel is 0 and another 255) or mode 5
B2BU_DI uint4 bc7_void_extent(uint32_t rgba, const DevTables& T)
{
    uint32_t c[4] = {rgba & 0xFFu, (rgba >> 8) & 0xFFu, (rgba >> 16) & 0xFFu, rgba >> 24};
    uint32_t err0 = 0, err1 = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) { err0 += c[i] == 255u; err1 += c[i] == 0u; }
    uint4 out = make_uint4(0u, 0u, 0u, 0u);
    if (err0 > 0 && err1 > 0) {
        // mode 5: 6 mode bits, 2 rotation bits, 7-bit colour pairs, 8-bit alpha pair, colour index 1, alpha index 0
        int pos = 8;
        out.x = 1u << 5;
#pragma unroll
        for (int i = 0; i < 3; i++) { putbits(out, pos, 7, T.m5lo[c[i]]); pos += 7; putbits(out, pos, 7, T.m5hi[c[i]]); pos += 7; }
        putbits(out, pos, 8, c[3]); pos += 8; putbits(out, pos, 8, c[3]); pos += 8;
        // colour weights: anchor 1 bit (value 1), then 15 x 2 bits of 01
        putbits(out, pos, 1, 1u); pos += 1;
        putbits(out, pos, 30, 0x15555555u); pos += 30;
        return out;
    }
    const uint32_t p = err1 < err0 ? 1u : 0u;
    int pos = 7;
    out.x = 1u << 6;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const uint32_t idx = c[i] + (p ? 0u : 1u);                               // bc7.rs:1126-1131
        putbits(out, pos, 7, T.m6lo[idx]); pos += 7; putbits(out, pos, 7, T.m6hi[idx]); pos += 7;
    }
    putbits(out, pos, 2, p * 3u); pos += 2;
    // weights: anchor 3 bits (5), then 15 x 4 bits of 5
    putbits(out, pos, 3, 5u); pos += 3;
    putbits(out, pos, 32, 0x55555555u); pos += 32;
    putbits(out, pos, 28, 0x05555555u);
    return out;
}

// bc7.rs:478-553 determine_unique_pbits for 8 total bits (BC7 modes 3 and 6), integer form:
// quantise(v,p) = nearest value of parity p (ties up), clamp; error = [parity(v) != p]
B2BU_DI uint32_t q8(uint32_t v, uint32_t p)
{
    // ((v - p + 1) >> 1) * 2 + p clamped to [p, 254 + p]
    const uint32_t q = (((v + 1u - p) >> 1) << 1) + p;
    return q > 254u + p ? 254u + p : q;
}

// Output block as two 64-bit halves; every position below is a compile-time constant after unrolling.
struct Out128 { uint64_t lo, hi; };
B2BU_DI void emit(Out128& o, int pos, int len, uint32_t v)        // v < 2^len
{
    if (len <= 0) return;
    if (pos + len <= 64) o.lo |= (uint64_t)v << pos;
    else if (pos >= 64) o.hi |= (uint64_t)v << (pos - 64);
    else { o.lo |= (uint64_t)v << pos; o.hi |= (uint64_t)v >> (64 - pos); }
}
B2BU_DI void emit64(Out128& o, int pos, uint64_t v)               // v already limited to the bits that fit
{
    if (pos >= 64) o.hi |= v << (pos - 64);
    else if (pos == 0) o.lo |= v;
    else { o.lo |= v << pos; o.hi |= v >> (64 - pos); }
}

// bc7.rs:9-310 for one (non void-extent) UASTC mode.  Endpoints travel as packed R|G<<8|B<<16|A<<24 words per
// subset endpoint (permutation, anchor swaps and the 8-bit p-bit search work on whole words), weights as one
// 64-bit stream per plane (width conversion, per-subset inversion and anchor-bit removal are SWAR).
template <int M> B2BU_DI uint4 bc7_block(const uint4& b, const DevTables& T, uint32_t pat, uint32_t compsel)
{
    using D = MD<M>;
    constexpr int BM = bc7_mode_for(M);
    constexpr Bc7Info BI = bc7_info(BM);
    constexpr int bb = BI.weight_bits;

    // ---- endpoints per UASTC subset, packed (uastc.rs:176-216).  UASTC mode 2 keeps its 4-bit quantised values
    // (v = 17 k): the shared p-bit LUTs of BC7 mode 1 are indexed by k
    uint32_t plo[3] = {0u, 0u, 0u}, phi[3] = {0u, 0u, 0u};
    {
        uint32_t e[D::N];
        if (BM == 1) { uint32_t d[D::N]; unpack_quant<M>(b, T, e, d); }
        else unpack_endpoints<M>(b, T, e);
#pragma unroll
        for (int s = 0; s < D::subsets; s++) {
            const int o = s * D::NC * 2;
            if (D::fmt == FMT_LA) {
                plo[s] = e[o] * 0x00010101u | (e[o + 2] << 24); phi[s] = e[o + 1] * 0x00010101u | (e[o + 3] << 24);
            } else if (D::fmt == FMT_RGB) {
                plo[s] = e[o] | (e[o + 2] << 8) | (e[o + 4] << 16) | 0xFF000000u; phi[s] = e[o + 1] | (e[o + 3] << 8) | (e[o + 5] << 16) | 0xFF000000u;
            } else {
                plo[s] = e[o] | (e[o + 2] << 8) | (e[o + 4] << 16) | (e[o + 6] << 24); phi[s] = e[o + 1] | (e[o + 3] << 8) | (e[o + 5] << 16) | (e[o + 7] << 24);
            }
        }
    }

    // ---- weights: one bb-bit-per-texel stream per plane (bc7.rs:377-398) ----
    const uint4 U = uniform_weights<M>(b, T, pat);
    uint64_t W0 = 0, W1 = 0;
    if (D::planes == 1) {
        if (D::wbits == bb) W0 = (uint64_t)U.x | ((uint64_t)U.y << 32);
        else if (D::wbits == 2) {                                    // 2 -> 4 bits: x * 5
            W0 = (uint64_t)(spread2to4(U.x & 0xFFFFu) * 5u) | ((uint64_t)(spread2to4(U.x >> 16) * 5u) << 32);
        } else if (D::wbits == 3) {                                  // 3 -> 4 bits: 2x + (x >> 2)
            const uint32_t x0 = spread3to4(U.x & 0xFFFFFFu), x1 = spread3to4(getbits(U, 24, 24));
            W0 = (uint64_t)(2u * x0 + ((x0 >> 2) & 0x11111111u)) | ((uint64_t)(2u * x1 + ((x1 >> 2) & 0x11111111u)) << 32);
        } else {                                                     // 5 -> 4 bits: LUT
#pragma unroll
            for (int i = 0; i < 16; i++) W0 |= (uint64_t)T.w5to4[getbits(U, 5 * i, 5)] << (4 * i);
        }
    } else if (D::wbits == 1) {                                      // texel-major, plane-minor; 1 -> 2 bits: x * 3
        W0 = (uint64_t)((U.x & 0x55555555u) * 3u);
        W1 = (uint64_t)(((U.x >> 1) & 0x55555555u) * 3u);
    } else {                                                         // two planes of 2-bit weights, de-interleaved
        W0 = (uint64_t)(compress4to2(U.x) | (compress4to2(U.y) << 16));
        W1 = (uint64_t)(compress4to2(U.x >> 2) | (compress4to2(U.y >> 2) << 16));
    }

    Out128 out;
    out.lo = 1ull << BM; out.hi = 0;
    int pos = BM + 1;

    uint32_t a1 = 0, a2 = 0;          // BC7 anchors of subsets 1, 2
    uint32_t qlo[3], qhi[3];          // endpoints per BC7 subset
    if (BI.subsets > 1) {
        // bc7.rs:113-198: partition index, subset permutation, anchor-MSB fix-up
        uint32_t info, bpw;
        if (M == 1) { info = 0u | (0u << 8) | (15u << 16) | (2u << 24); bpw = T.bc7pat2[0]; }
        else if (M == 7) { info = T.bc7p23[pat]; bpw = T.bc7pat23[pat]; }
        else if (D::subsets == 2) { info = T.bc7p2[pat]; bpw = T.bc7pat2[pat]; }
        else { info = T.bc7p3[pat]; bpw = T.bc7pat3[pat]; }
        emit(out, pos, BI.pat_bits, info & 63u);
        pos += BI.pat_bits;
        a1 = (info >> 16) & 15u; a2 = (info >> 20) & 15u;
#pragma unroll
        for (int s = 0; s < BI.subsets; s++) {
            const uint32_t src = (info >> (8 + 2 * s)) & 3u;
            uint32_t l = plo[0], h = phi[0];
            if (D::subsets >= 2) { if (src == 1u) { l = plo[1]; h = phi[1]; } }
            if (D::subsets >= 3) { if (src == 2u) { l = plo[2]; h = phi[2]; } }
            qlo[s] = l; qhi[s] = h;
        }
        // a BC7 subset whose anchor weight has its MSB set swaps its endpoints and inverts its weights
        const bool inv0 = (W0 >> (bb - 1)) & 1ull;
        const bool inv1 = (W0 >> (a1 * bb + bb - 1)) & 1ull;
        const bool inv2 = BI.subsets == 3 ? (bool)((W0 >> (a2 * bb + bb - 1)) & 1ull) : false;
        const bool invs[3] = {inv0, inv1, inv2};
#pragma unroll
        for (int s = 0; s < BI.subsets; s++) { const uint32_t l = qlo[s], h = qhi[s]; qlo[s] = invs[s] ? h : l; qhi[s] = invs[s] ? l : h; }
        if (bb == 2) {
            uint32_t x = 0;
#pragma unroll
            for (int s = 0; s < BI.subsets; s++) if (invs[s]) x |= subset_mask2(bpw, (uint32_t)s);
            W0 ^= (uint64_t)x;
        } else {
            // UASTC mode 2 -> BC7 mode 1 (3-bit weights, 2 subsets): the BC7 map is the UASTC map or its complement
            const uint64_t full = 0xFFFFFFFFFFFFull, u1 = T.pat2w3[pat];
            const uint64_t m1 = ((info >> 8) & 3u) == 0u ? u1 : (~u1 & full);
            uint64_t x = 0;
            if (inv0) x |= ~m1 & full;
            if (inv1) x |= m1;
            W0 ^= x;
        }
    } else {
        qlo[0] = plo[0]; qhi[0] = phi[0];
        if (D::planes == 2) {
            // bc7.rs:207-246: rotate the dual-plane channel into alpha.  The "anchor MSB set" branches
            // (:200-205, :208-236) are dead: UASTC anchors carry no MSB and every width LUT keeps them low.
            const uint32_t sel = compsel == 0u ? 0x0213u : compsel == 1u ? 0x1230u : compsel == 2u ? 0x2310u : 0x3210u;
            qlo[0] = __byte_perm(qlo[0], 0u, sel); qhi[0] = __byte_perm(qhi[0], 0u, sel);
            emit(out, pos, 2, (compsel + 1u) & 3u);
            pos += 2;
        }
    }

    // ---- endpoint re-quantisation (bc7.rs:249-273): after this block every byte of q* is the field to write ----
    uint32_t pb[3][2] = {{0u, 0u}, {0u, 0u}, {0u, 0u}};
    if (BI.p_bits && BI.color_bits == 7) {
        // 8 total bits (bc7.rs:478-553 in integers): the error of p is the number of channels whose parity differs from
        // p, so p = 1 iff more than half of the channels are odd; then field = p ? v >> 1 : min((v + 1) >> 1, 127)
        constexpr uint32_t CH1 = BI.channels == 3 ? 0x00010101u : 0x01010101u;
#pragma unroll
        for (int s = 0; s < BI.subsets; s++) {
#pragma unroll
            for (int k = 0; k < 2; k++) {
                const uint32_t x = k ? qhi[s] : qlo[s];
                const uint32_t odd = (uint32_t)__popc(x & CH1);
                const uint32_t p = (2u * odd > (uint32_t)BI.channels) ? 1u : 0u;
                uint32_t q = ((x >> 1) & 0x7F7F7F7Fu) + (p ? 0u : (x & 0x01010101u));
                q -= (q >> 7) & 0x01010101u;
                if (k) qhi[s] = q; else qlo[s] = q;
                pb[s][k] = p;
            }
        }
    } else if (BI.p_bits) {
        // 6 total bits (BC7 mode 7): LUTs of the exact squared error of both p and of the quantised value
#pragma unroll
        for (int s = 0; s < BI.subsets; s++) {
#pragma unroll
            for (int k = 0; k < 2; k++) {
                const uint32_t x = k ? qhi[s] : qlo[s];
                const uint32_t c0 = x & 0xFFu, c1 = (x >> 8) & 0xFFu, c2 = (x >> 16) & 0xFFu, c3 = x >> 24;
                const uint32_t esum = T.pe6p[c0] + T.pe6p[c1] + T.pe6p[c2] + T.pe6p[c3];       // err(p=0) | err(p=1) << 16, each < 2^14
                const uint32_t p = (esum >> 16) < (esum & 0xFFFFu) ? 1u : 0u;
                const uint8_t* tq = T.pq6[p];
                const uint32_t q = (uint32_t)(tq[c0] >> 1) | ((uint32_t)(tq[c1] >> 1) << 8) | ((uint32_t)(tq[c2] >> 1) << 16) | ((uint32_t)(tq[c3] >> 1) << 24);
                if (k) qhi[s] = q; else qlo[s] = q;
                pb[s][k] = p;
            }
        }
    } else if (BI.sp_bits) {
        // BC7 mode 1 <= UASTC mode 2: bytes hold k (endpoint = 17 k); f32 terms from the LUT, summed in the reference's order
#pragma unroll
        for (int s = 0; s < BI.subsets; s++) {
            float err[2];
#pragma unroll
            for (int p = 0; p < 2; p++) {
                float acc = 0.f;
#pragma unroll
                for (int c = 0; c < 3; c++) {
                    const float t = __fadd_rn(__uint_as_float(T.se7_bits[p][(qlo[s] >> (8 * c)) & 0xFFu]), __uint_as_float(T.se7_bits[p][(qhi[s] >> (8 * c)) & 0xFFu]));
                    acc = __fadd_rn(acc, t);
                }
                err[p] = acc;
            }
            const uint32_t p = err[1] < err[0] ? 1u : 0u;
            const uint8_t* tq = T.sq7[p];
            uint32_t l = 0, h = 0;
#pragma unroll
            for (int c = 0; c < 3; c++) {
                l |= (uint32_t)(tq[(qlo[s] >> (8 * c)) & 0xFFu] >> 1) << (8 * c);
                h |= (uint32_t)(tq[(qhi[s] >> (8 * c)) & 0xFFu] >> 1) << (8 * c);
            }
            qlo[s] = l; qhi[s] = h;
            pb[s][0] = p; pb[s][1] = p;
        }
    } else {
        // no p-bits: (e * (2^bits - 1) + 127) / 255 from the LUTs; BC7 mode 5 keeps its 8-bit alpha
        const uint8_t* tq = BI.color_bits == 5 ? T.q5 : T.q7;
#pragma unroll
        for (int s = 0; s < BI.subsets; s++) {
#pragma unroll
            for (int k = 0; k < 2; k++) {
                const uint32_t x = k ? qhi[s] : qlo[s];
                uint32_t q = (uint32_t)tq[x & 0xFFu] | ((uint32_t)tq[(x >> 8) & 0xFFu] << 8) | ((uint32_t)tq[(x >> 16) & 0xFFu] << 16);
                if (BI.channels == 4) q |= x & 0xFF000000u;
                if (k) qhi[s] = q; else qlo[s] = q;
            }
        }
    }

    // ---- bc7.rs:276-307: endpoints channel-major, p-bits, weights ----
#pragma unroll
    for (int c = 0; c < BI.channels; c++) {
        const int nb = c == 3 ? BI.alpha_bits : BI.color_bits;
#pragma unroll
        for (int s = 0; s < BI.subsets; s++) {
            emit(out, pos, nb, (qlo[s] >> (8 * c)) & 0xFFu); pos += nb;
            emit(out, pos, nb, (qhi[s] >> (8 * c)) & 0xFFu); pos += nb;
        }
    }
    if (BI.p_bits) {
#pragma unroll
        for (int s = 0; s < BI.subsets; s++) { emit(out, pos, 2, (pb[s][1] << 1) | pb[s][0]); pos += 2; }
    } else if (BI.sp_bits) {
        emit(out, pos, 2, (pb[1][0] << 1) | pb[0][0]); pos += 2;
    }
    // weights: drop the anchors' MSB (always 0 by now), highest position first
#pragma unroll
    for (int p = 0; p < BI.planes; p++) {
        uint64_t ws = p ? W1 : W0;
        if (BI.subsets == 3) {
            const uint32_t ahi = a1 > a2 ? a1 : a2, alo = a1 > a2 ? a2 : a1;
            ws = delete_bit<uint64_t>(ws, ahi * bb + bb - 1);
            ws = delete_bit<uint64_t>(ws, alo * bb + bb - 1);
        } else if (BI.subsets == 2) {
            ws = delete_bit<uint64_t>(ws, a1 * bb + bb - 1);
        }
        ws = delete_bit<uint64_t>(ws, bb - 1);
        emit64(out, pos, ws);
        pos += 16 * bb - BI.subsets;
    }
    return make_uint4((uint32_t)out.lo, (uint32_t)(out.lo >> 32), (uint32_t)out.hi, (uint32_t)(out.hi >> 32));
}

// ------------------------------------------------------------------------------------------
// ETC1 / ETC2 back-end (target_formats/etc.rs:11-341).  Mode independent once the texels exist.
// ------------------------------------------------------------------------------------------
struct EtcFlags { uint32_t flip, diff, i0, i1, has_bias, bias, etc2tm; };

template <int M> B2BU_DI EtcFlags read_trans_flags(const uint4& b)             // uastc.rs:411-436
{
    using D = MD<M>;
    constexpr bool m10_12 = (M >= 10 && M <= 12);
    EtcFlags f;
    int pos = D::FPOS + 1 + (m10_12 ? 0 : 1);            // bc1h0, bc1h1
    f.flip = getbits(b, pos, 1); pos += 1;
    f.diff = getbits(b, pos, 1); pos += 1;
    f.i0 = getbits(b, pos, 3); pos += 3;
    f.i1 = getbits(b, pos, 3); pos += 3;
    f.has_bias = m10_12 ? 0u : 1u;
    f.bias = m10_12 ? 0u : getbits(b, pos, 5); pos += m10_12 ? 0 : 5;
    f.etc2tm = D::HAS_ALPHA ? getbits(b, pos, 8) : 0u;
    return f;
}

B2BU_DI uint2 etc2_solid_alpha(uint32_t v)                                      // etc.rs:261-275
{
    return make_uint2(v | (0x1Du << 8) | (0x92u << 16) | (0x49u << 24), 0x24u | (0x92u << 8) | (0x49u << 16) | (0x24u << 24));
}

// etc.rs:277-341 write_etc2_alpha_block
B2BU_DI uint2 etc2_alpha_block(const uint32_t (&px)[16], uint32_t etc2tm, const DevTables& T)
{
    if (etc2tm == 0u) return etc2_solid_alpha(255u);
    uint32_t mn = 255u, mx = 0u;
#pragma unroll
    for (int i = 0; i < 16; i++) { const uint32_t a = px[i] >> 24; mn = min(mn, a); mx = max(mx, a); }
    if (mn == mx) return etc2_solid_alpha(mn);
    const uint32_t ti = etc2tm & 15u;
    const int mult = (int)(etc2tm >> 4);
    // centre = round(lerp(min, max, -mod_min/range)), literal f32 with no contraction (etc.rs:301-307)
    const float amt = __uint_as_float(T.eac_amt_bits[ti]), om = __uint_as_float(T.eac_1m_amt_bits[ti]);
    const float l = __fadd_rn(__fmul_rn((float)mn, om), __fmul_rn((float)mx, amt));
    const int center = (int)roundf(l);
    int vals[8];
#pragma unroll
    for (int k = 0; k < 8; k++) { const int v = center + (int)T.eac_mod[ti][k] * mult; vals[k] = v < 0 ? 0 : v > 255 ? 255 : v; }
    // Nearest of the 8 values, lowest index on ties (etc.rs:317-323 min_by_key).  In value order the candidates are
    // k = 3,2,1,0,4,5,6,7 (the table rows are {-m0..-m3, m0-1..m3-1} with m increasing, clamping keeps the order), so the
    // answer is the number of boundaries between neighbours that the texel's alpha passes: below k = 0 a tie goes to the
    // higher value (= lower index), from k = 0 upwards to the lower one.  7 compares per texel instead of
    // 8 x (difference, abs, key, min); tests/test_emu_kernels.py checks the rule exhaustively.
    const int sv[8] = {vals[3], vals[2], vals[1], vals[0], vals[4], vals[5], vals[6], vals[7]};
    int thr[7];
#pragma unroll
    for (int j = 6; j >= 0; j--) {
        const int sum = sv[j] + sv[j + 1];
        const bool same = sv[j] == sv[j + 1];
        // equal neighbours: below k = 0 the boundary is always passed (towards the lower index); above, it is passed together
        // with the next one (a texel only moves past a duplicate when a later, different value is strictly closer)
        thr[j] = j < 3 ? (same ? 0 : (sum + 1) >> 1) : (same ? (j == 6 ? 256 : thr[j < 6 ? j + 1 : 6]) : (sum >> 1) + 1);
    }
    uint64_t sel = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) {
        const int a = (int)(px[i] >> 24);
        uint32_t p = 0;
#pragma unroll
        for (int j = 0; j < 7; j++) p += a >= thr[j] ? 1u : 0u;
        const uint32_t best = (0x76540123u >> (4u * p)) & 7u;
        const int x = i / 4, y = i % 4, id = y * 4 + x;  // etc.rs:325-329 (transposed pixel order)
        sel |= (uint64_t)best << (45 - id * 3);
    }
    // bytes: centre, etc2tm, then selectors big-endian (48 bits)
    const uint32_t s47_40 = (uint32_t)(sel >> 40) & 0xFFu, s39_32 = (uint32_t)(sel >> 32) & 0xFFu;
    const uint32_t lo32 = (uint32_t)sel;
    uint2 r;
    r.x = ((uint32_t)center & 0xFFu) | (etc2tm << 8) | (s47_40 << 16) | (s39_32 << 24);
    r.y = __byte_perm(lo32, 0u, 0x0123);
    return r;
}

// etc.rs:203-259 apply_etc1_bias for one channel.  The reference's 32-way match on the bias index is a table here
// (a switch diverges up to 32 ways inside a warp): per bias, 2 bits (delta + 2) for each (sub-block, channel).
constexpr int etc1_bias_delta_ref(uint32_t bias, int c, bool s1)
{
    return bias == 2 ? (s1 ? 0 : (c == 0 ? -1 : 0)) : bias == 5 ? (s1 ? 0 : (c == 1 ? -1 : 0)) : bias == 6 ? (s1 ? 0 : (c == 2 ? -1 : 0))
         : bias == 7 ? (s1 ? 0 : (c == 0 ? 1 : 0)) : bias == 11 ? (s1 ? 0 : (c == 1 ? 1 : 0)) : bias == 15 ? (s1 ? 0 : (c == 2 ? 1 : 0))
         : bias == 18 ? (s1 ? (c == 0 ? -1 : 0) : 0) : bias == 19 ? (s1 ? (c == 1 ? -1 : 0) : 0) : bias == 20 ? (s1 ? (c == 2 ? -1 : 0) : 0)
         : bias == 21 ? (s1 ? (c == 0 ? 1 : 0) : 0) : bias == 24 ? (s1 ? (c == 1 ? 1 : 0) : 0) : bias == 8 ? (s1 ? (c == 2 ? 1 : 0) : 0)
         : bias == 10 ? -2 : bias == 27 ? (s1 ? 0 : -1) : bias == 28 ? (s1 ? -1 : 1) : bias == 29 ? (s1 ? 1 : 0) : bias == 30 ? (s1 ? -1 : 0)
         : bias == 31 ? (s1 ? 0 : 1) : (int)((bias / (c == 0 ? 1u : c == 1 ? 3u : 9u)) % 3u) - 1;
}
constexpr uint32_t etc1_bias_word(uint32_t bias)
{
    uint32_t w = 0;
    for (int sb = 0; sb < 2; sb++)
        for (int c = 0; c < 3; c++) w |= (uint32_t)(etc1_bias_delta_ref(bias, c, sb == 1) + 2) << (2 * (sb * 3 + c));
    return w;
}
#define B2BU_BW(i) (uint16_t)etc1_bias_word(i)
static __device__ const uint16_t kEtc1BiasLut[32] = {
    B2BU_BW(0), B2BU_BW(1), B2BU_BW(2), B2BU_BW(3), B2BU_BW(4), B2BU_BW(5), B2BU_BW(6), B2BU_BW(7), B2BU_BW(8), B2BU_BW(9), B2BU_BW(10),
    B2BU_BW(11), B2BU_BW(12), B2BU_BW(13), B2BU_BW(14), B2BU_BW(15), B2BU_BW(16), B2BU_BW(17), B2BU_BW(18), B2BU_BW(19), B2BU_BW(20),
    B2BU_BW(21), B2BU_BW(22), B2BU_BW(23), B2BU_BW(24), B2BU_BW(25), B2BU_BW(26), B2BU_BW(27), B2BU_BW(28), B2BU_BW(29), B2BU_BW(30),
    B2BU_BW(31)};
#undef B2BU_BW
// bw: the block's kEtc1BiasLut entry
B2BU_DI uint32_t etc1_apply_bias(uint32_t v0, uint32_t bw, int c, uint32_t limit, uint32_t subblock)
{
    const int delta = (int)((bw >> (2 * ((int)subblock * 3 + c))) & 3u) - 2;
    int v = (int)v0;
    if (v == 0) v += (delta == -2) ? 3 : delta + 1;
    else if (v == (int)limit) v += delta - 1;
    else { v += delta; if (v < 0 || v > (int)limit) v = (v - delta) - delta; }
    return (uint32_t)v;
}

// etc.rs:78-198.  The reference transposes the texels when !flip so that each sub-block is a run
// of 8; here texel (x,y) simply belongs to sub-block (flip ? y>>1 : x>>1) and its selector lands at
// ETC pixel id x*4+y either way (etc.rs:188-194, :363-393).
B2BU_DI uint2 etc1_block(const uint32_t (&px)[16], const EtcFlags& f, const DevTables& T)
{
    // quadrant sums in 16-bit lanes: q[qy][qx]
    // quadrant sums per channel, q[qy][qx][c], as byte dot products (one IDP.4A on the fma pipe per channel and texel)
    uint32_t q[2][2][3] = {{{0u, 0u, 0u}, {0u, 0u, 0u}}, {{0u, 0u, 0u}, {0u, 0u, 0u}}};
#pragma unroll
    for (int i = 0; i < 16; i++) {
        const int x = i & 3, y = i >> 2;
#pragma unroll
        for (int c = 0; c < 3; c++) q[y >> 1][x >> 1][c] = __dp4a(px[i], 1u << (8 * c), q[y >> 1][x >> 1][c]);
    }
    const bool flip = f.flip != 0u;
    const uint32_t limit = f.diff ? 31u : 15u;
    uint32_t c0[3], c1[3];
#pragma unroll
    for (int c = 0; c < 3; c++) {
        // sub-block 0 = TL + (flip ? TR : BL), sub-block 1 = BR + (flip ? BL : TR)
        const uint32_t s0 = q[0][0][c] + (flip ? q[0][1][c] : q[1][0][c]), s1 = q[1][1][c] + (flip ? q[1][0][c] : q[0][1][c]);
        c0[c] = (s0 * limit + 1020u) / 2040u;
        c1[c] = (s1 * limit + 1020u) / 2040u;
    }
    if (f.has_bias) {
        const uint32_t bw = kEtc1BiasLut[f.bias & 31u];
#pragma unroll
        for (int c = 0; c < 3; c++) { c0[c] = etc1_apply_bias(c0[c], bw, c, limit, 0u); c1[c] = etc1_apply_bias(c1[c], bw, c, limit, 1u); }
    }
    uint32_t base0[3], base1[3], hdr = 0;
    if (!f.diff) {
#pragma unroll
        for (int c = 0; c < 3; c++) {
            hdr |= ((c0[c] << 4) | c1[c]) << (8 * c);
            base0[c] = c0[c] * 17u; base1[c] = c1[c] * 17u;
        }
    } else {
#pragma unroll
        for (int c = 0; c < 3; c++) {
            int dlt = (int)c1[c] - (int)c0[c];
            dlt = dlt < -4 ? -4 : dlt > 3 ? 3 : dlt;
            hdr |= ((c0[c] << 3) | ((uint32_t)dlt & 7u)) << (8 * c);
            const uint32_t c1d = (uint32_t)((int)c0[c] + dlt);
            base0[c] = (c0[c] << 3) | (c0[c] >> 2); base1[c] = (c1d << 3) | (c1d >> 2);
        }
    }
    hdr |= ((f.i0 << 5) | (f.i1 << 2) | (f.diff << 1) | f.flip) << 24;

    // luminance thresholds per sub-block (etc.rs:165-177)
    int thr[2][3];
#pragma unroll
    for (int sb = 0; sb < 2; sb++) {
        const uint32_t inten = sb ? f.i1 : f.i0;
        int lum[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int md = T.etc1_mod[inten][k];
            int l = 0;
#pragma unroll
            for (int c = 0; c < 3; c++) {
                int v = (int)(sb ? base1[c] : base0[c]) + md;
                v = v < 0 ? 0 : v > 255 ? 255 : v;
                l += v * (c == 0 ? 108 : c == 1 ? 366 : 38);
            }
            lum[k] = l;
        }
        thr[sb][0] = (lum[0] + lum[1]) / 2; thr[sb][1] = (lum[1] + lum[2]) / 2; thr[sb][2] = (lum[2] + lum[3]) / 2;
    }
    // thresholds for the two quadrants whose sub-block depends on flip: TR (x>=2,y<2) and BL (x<2,y>=2)
    int thr_tr[3], thr_bl[3];
#pragma unroll
    for (int k = 0; k < 3; k++) { thr_tr[k] = flip ? thr[0][k] : thr[1][k]; thr_bl[k] = flip ? thr[1][k] : thr[0][k]; }

    // selector s = #(lum >= t_k); its ETC1 code [3,2,0,1] (etc.rs:433) has msb = (s < 2) = (lum < t1) and
    // lsb = (s == 0 || s == 3) = (lum < t0) || (lum >= t2).  Both are SIGN bits -- of lum - t1, and of (lum - t0) | ~(lum - t2)
    // (all operands are below 2^18) -- and a funnel shift (acc : x) << 1 appends the sign of x to acc in one instruction: the
    // texels are visited in descending bit position, so the two 16-bit planes come out in ETC1's order without any
    // per-texel insert (3 subtractions, which the fma pipe takes, + 1 LOP3 + 2 SHF instead of 3 compares + 4 predicated ORs).
    uint32_t msb = 0u, lsb = 0u;
#pragma unroll
    for (int bitpos = 15; bitpos >= 0; bitpos--) {
        const int pid = bitpos >= 8 ? bitpos - 8 : bitpos + 8;        // byte 1 holds pixels 0..7, byte 0 pixels 8..15
        const int x = pid >> 2, y = pid & 3, i = y * 4 + x;           // ETC pixel id = x * 4 + y (etc.rs:363-393)
        const int qx = x >> 1, qy = y >> 1;
        const int t0 = (qx == qy) ? thr[qx][0] : (qx ? thr_tr[0] : thr_bl[0]);
        const int t1 = (qx == qy) ? thr[qx][1] : (qx ? thr_tr[1] : thr_bl[1]);
        const int t2 = (qx == qy) ? thr[qx][2] : (qx ? thr_tr[2] : thr_bl[2]);
        // 108 R + 366 G + 38 B as two byte dot products (366 = 2 x 183 does not fit a signed byte weight)
        const int lum = __dp4a(px[i], 0x0026B76Cu, __dp4a(px[i], 0x0000B700u, 0u));
        msb = __funnelshift_l((uint32_t)(lum - t1), msb, 1);
        lsb = __funnelshift_l((uint32_t)(lum - t0) | ~(uint32_t)(lum - t2), lsb, 1);
    }
    const uint32_t selbits = (msb & 0xFFFFu) | (lsb << 16);
    return make_uint2(hdr, selbits);
}

// etc.rs:43-76: void-extent -> ETC1 from the stored hints
B2BU_DI uint2 etc1_void_extent(const uint4& b)
{
    const uint32_t d = getbits(b, 37, 1), ii = getbits(b, 38, 3), s = getbits(b, 41, 2);
    const uint32_t r = getbits(b, 43, 5), g = getbits(b, 48, 5), bl = getbits(b, 53, 5);
    uint32_t hdr;
    // u8 arithmetic in the reference: (c << 4 | c) keeps only the low 8 bits (etc.rs:54-56)
    if (!d) hdr = (((r << 4) | r) & 0xFFu) | ((((g << 4) | g) & 0xFFu) << 8) | ((((bl << 4) | bl) & 0xFFu) << 16);
    else hdr = (r << 3) | (g << 11) | (bl << 19);
    hdr |= ((ii << 5) | (ii << 2) | (d << 1)) << 24;
    const uint32_t code = (0x1u << 6 | 0x0u << 4 | 0x2u << 2 | 0x3u) >> (2 * s) & 3u;   // [3,2,0,1][s]
    const uint32_t hi = (code >> 1) ? 0xFFFFu : 0u, lo = (code & 1u) ? 0xFFFFu : 0u;
    return make_uint2(hdr, hi | (lo << 16));
}

// ------------------------------------------------------------------------------------------
// One block, any mode: dispatch on the mode id to the specialised code above.
// ------------------------------------------------------------------------------------------
#define B2BU_FOR_EACH_MODE(X) X(0) X(1) X(2) X(3) X(4) X(5) X(6) X(7) X(9) X(10) X(11) X(12) X(13) X(14) X(15) X(16) X(17) X(18)

// ---- per-target block functions: return false on error ------------------------------------
struct BlockOut { uint4 v; uint2 etc; uint32_t px[16]; };

template <int M> B2BU_DI bool header_ok(const uint4& b, uint32_t& pat, uint32_t& compsel)
{
    compsel = read_compsel<M>(b);
    pat = read_pattern<M>(b);
    return pat < (uint32_t)MD<M>::PCOUNT;          // uastc.rs:360-365
}

// mode = T.mode_lut[low 7 bits] (uastc.rs:329-341); 19 = invalid code.
// RGBA rows go to `sink` (TARGET == TGT_RGBA); the other targets fill o.v / o.etc.
template <int TARGET, class Sink>
B2BU_DI uint32_t transcode_mode_sink(uint32_t mode, const uint4& b, const DevTables& T, BlockOut& o, Sink& sink)
{
    uint32_t pat = 0, compsel = 0;
    if (mode == 8u) {
        const uint32_t c = mode8_rgba(b);
        if (TARGET == TGT_RGBA) {
            const uint4 r = make_uint4(c, c, c, c);
            if (Sink::ROLLED) {
#pragma unroll 1
                for (int y = 0; y < 4; y++) sink.row(y, r);
            } else {
#pragma unroll
                for (int y = 0; y < 4; y++) sink.row(y, r);
            }
        } else if (TARGET == TGT_ASTC) o.v = astc_void_extent(c);
        else if (TARGET == TGT_BC7) o.v = bc7_void_extent(c, T);
        else {
            o.etc = etc1_void_extent(b);
            if (TARGET == TGT_ETC2) { const uint2 a = etc2_solid_alpha(c >> 24); o.v = make_uint4(a.x, a.y, o.etc.x, o.etc.y); }
        }
        return ERR_OK;
    }
    if (TARGET == TGT_ASTC || TARGET == TGT_BC7) {
        switch (mode) {
#define X(M) case M: if (!header_ok<M>(b, pat, compsel)) return ERR_PATTERN; \
                     o.v = (TARGET == TGT_ASTC) ? astc_block<M>(b, T, pat, compsel) : bc7_block<M>(b, T, pat, compsel); break;
            B2BU_FOR_EACH_MODE(X)
#undef X
        default: return ERR_MODE;
        }
        return ERR_OK;
    }
    // RGBA, ETC1, ETC2: mode-specialised front-end, then texel code shared by all modes of a class
    EtcFlags f;
    Canon c;
    switch (mode) {
#define X(M) case M: if (!header_ok<M>(b, pat, compsel)) return ERR_PATTERN; \
                     if (TARGET != TGT_RGBA) f = read_trans_flags<M>(b); \
                     canon_front<M>(b, T, pat, compsel, c); break;
        B2BU_FOR_EACH_MODE(X)
#undef X
    default: return ERR_MODE;
    }
    if (TARGET == TGT_RGBA) {
        interp_block<true>(mode, c, sink);
    } else {
        PxArraySink ps{o.px};
        interp_block<TARGET == TGT_ETC2>(mode, c, ps);                 // ETC1 never reads alpha
        o.etc = etc1_block(o.px, f, T);
        if (TARGET == TGT_ETC2) { const uint2 a = etc2_alpha_block(o.px, f.etc2tm, T); o.v = make_uint4(a.x, a.y, o.etc.x, o.etc.y); }
    }
    return ERR_OK;
}

template <int TARGET>
B2BU_DI uint32_t transcode_mode(uint32_t mode, const uint4& b, const DevTables& T, BlockOut& o)
{
    PxArraySink ps{o.px};
    return transcode_mode_sink<TARGET>(mode, b, T, o, ps);
}

template <int TARGET>
B2BU_DI uint32_t transcode_one(const uint4& b, const DevTables& T, BlockOut& o)
{
    return transcode_mode<TARGET>(T.mode_lut[b.x & 127u], b, T, o);
}

}  // namespace b2bu
