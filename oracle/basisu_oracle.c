/*
 * basisu_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 *
 * A plain-C restatement of the reference crate's scalar algorithm for the hot path
 * (JakubValtar/basisu_rs: UASTC block -> RGBA / ASTC / BC7 / ETC1 / ETC2 and the ETC1S /
 * BasisLZ slice decode).  It follows the reference's structure literally -- bit-at-a-time
 * reader/writers, the same operation order, literal f32 arithmetic -- so that it can serve as
 * the bit-exact checker for the CUDA kernels and as the "port" CPU baseline in bench.py.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  Nothing under basisu_rs_b200/ links or calls it.
 *
 * PINNING: the UASTC half is pinned by the reference's own 608 x 5 known-answer vectors
 * (tests/golden/uastc_kat.bin, from reference tests/block_test_cases/ *.rs files) -- see
 * tests/test_oracle_kat.py.  The ETC1S half has NO reference vector anywhere ("parity
 * unpinned"); it is a careful restatement only.
 *
 * Compile: gcc -O2 -ffp-contract=off -fno-fast-math (x86-64 SSE, no x87) -- see oracle/Makefile.
 * Each function cites the reference file:line it follows (paths relative to the crate root).
 */
#include <stdint.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <pthread.h>

#include "oracle_tables.h"

#define ORC_API __attribute__((visibility("default")))

enum { ORC_OK = 0, ORC_ERR_LEN = 1, ORC_ERR_MODE = 2, ORC_ERR_PATTERN = 3, ORC_ERR_HUFF = 4,
       ORC_ERR_SELECTOR_CB = 5, ORC_ERR_PRED = 6, ORC_ERR_VLC = 7, ORC_ERR_RANGE = 8,
       ORC_ERR_HEADER = 9, ORC_ERR_CRC = 10, ORC_ERR_FORMAT = 11, ORC_ERR_ARG = 12 };

/* src/lib.rs:57-61  mask!(n): n low one-bits, mask(0)=0, mask(32)=all ones */
static inline uint32_t mask32(uint32_t n) { return n >= 32 ? 0xFFFFFFFFu : ((1u << n) - 1u); }

/* ------------------------------------------------------------------------------------------
 * src/bitreader.rs:3-61  BitReaderLsb -- bytes past the end read as 0
 * ---------------------------------------------------------------------------------------- */
typedef struct { const uint8_t *bytes; size_t len; size_t bit_pos; } BitReader;

static void br_init(BitReader *r, const uint8_t *b, size_t len) { r->bytes = b; r->len = len; r->bit_pos = 0; }

static uint32_t br_peek(const BitReader *r, unsigned count)   /* bitreader.rs:37-60 */
{
    size_t byte = r->bit_pos / 8;
    uint32_t result = 0;
    unsigned read = 0;
    {
        unsigned bit = (unsigned)(r->bit_pos % 8);
        uint32_t v = byte < r->len ? r->bytes[byte] : 0;
        result |= v >> bit;
        read += 8 - bit;
        byte += 1;
    }
    for (;;) {
        if (read >= count) return result & mask32(count);
        uint32_t v = byte < r->len ? r->bytes[byte] : 0;
        result |= v << read;
        read += 8;
        byte += 1;
    }
}
static void br_remove(BitReader *r, unsigned count) { r->bit_pos += count; }          /* :33 */
static uint32_t br_read(BitReader *r, unsigned count) { uint32_t v = br_peek(r, count); br_remove(r, count); return v; }

/* ------------------------------------------------------------------------------------------
 * src/bitwriter.rs:3-54  BitWriterLsb  /  :58-116 BitWriterMsbRevBytes
 * OR into a zeroed buffer; bytes outside the buffer are silently dropped.
 * ---------------------------------------------------------------------------------------- */
typedef struct { uint8_t *bytes; size_t len; size_t bit_pos; } BitWriter;

static void bw_init(BitWriter *w, uint8_t *b, size_t len) { w->bytes = b; w->len = len; w->bit_pos = 0; }

static void bw_or_at(uint8_t *bytes, size_t len, size_t bit_pos, unsigned count, uint32_t v)
{
    size_t byte = bit_pos / 8;
    unsigned written = 0;
    {
        unsigned bit = (unsigned)(bit_pos % 8);
        if (byte < len) bytes[byte] |= (uint8_t)(v << bit);
        written += 8 - bit;
        byte += 1;
    }
    for (;;) {
        if (written >= count) return;
        if (byte < len) bytes[byte] |= (uint8_t)(v >> written);
        written += 8;
        byte += 1;
    }
}
static void bw_write(BitWriter *w, unsigned count, uint32_t v)     /* bitwriter.rs:23-51 */
{
    v &= mask32(count);
    size_t pos = w->bit_pos;
    w->bit_pos += count;
    bw_or_at(w->bytes, w->len, pos, count, v);
}

typedef struct { uint8_t *bytes; size_t len; size_t bit_pos; } BitWriterMsb;
static void bwm_init(BitWriterMsb *w, uint8_t *b, size_t len) { w->bytes = b; w->len = len; w->bit_pos = len * 8; }
static void bwm_write(BitWriterMsb *w, unsigned count, uint32_t v)  /* bitwriter.rs:89-113 */
{
    v &= mask32(count);
    w->bit_pos -= count;                       /* wrapping_sub; never wraps for valid modes */
    bw_or_at(w->bytes, w->len, w->bit_pos, count, v);
}
static uint32_t reverse_bits32(uint32_t x)
{
    x = (x >> 16) | (x << 16);
    x = ((x & 0xFF00FF00u) >> 8) | ((x & 0x00FF00FFu) << 8);
    x = ((x & 0xF0F0F0F0u) >> 4) | ((x & 0x0F0F0F0Fu) << 4);
    x = ((x & 0xCCCCCCCCu) >> 2) | ((x & 0x33333333u) << 2);
    x = ((x & 0xAAAAAAAAu) >> 1) | ((x & 0x55555555u) << 1);
    return x;
}
static void bwm_write_rev(BitWriterMsb *w, unsigned count, uint32_t v)   /* bitwriter.rs:79-82 */
{
    /* v.reverse_bits().wrapping_shr(32 - count): shift amount taken mod 32 */
    v = reverse_bits32(v) >> ((32u - count) & 31u);
    bwm_write(w, count, v);
}

/* ------------------------------------------------------------------------------------------
 * src/uastc.rs:443-557  Mode table
 * ---------------------------------------------------------------------------------------- */
enum { FMT_RGB = 0, FMT_RGBA = 1, FMT_LA = 2 };
typedef struct {
    uint8_t id, code_size, endpoint_range_index, format, weight_bits, plane_count, subset_count, trans_flags_bits;
} Mode;

static const Mode MODES[19] = {        /* uastc.rs:528-557 */
    { 0, 4, 19, FMT_RGB,  4, 1, 1, 15 }, { 1, 6, 20, FMT_RGB,  2, 1, 1, 15 },
    { 2, 5,  8, FMT_RGB,  3, 1, 2, 15 }, { 3, 5,  7, FMT_RGB,  2, 1, 3, 15 },
    { 4, 5, 12, FMT_RGB,  2, 1, 2, 15 }, { 5, 5, 20, FMT_RGB,  3, 1, 1, 15 },
    { 6, 5, 18, FMT_RGB,  2, 2, 1, 15 }, { 7, 5, 12, FMT_RGB,  2, 1, 2, 15 },
    { 8, 5,  0, FMT_RGBA, 0, 1, 1,  0 },
    { 9, 5,  8, FMT_RGBA, 2, 1, 2, 23 }, {10, 3, 13, FMT_RGBA, 4, 1, 1, 17 },
    {11, 2, 13, FMT_RGBA, 2, 2, 1, 17 }, {12, 3, 19, FMT_RGBA, 3, 1, 1, 17 },
    {13, 5, 20, FMT_RGBA, 1, 2, 1, 23 }, {14, 5, 20, FMT_RGBA, 2, 1, 1, 23 },
    {15, 7, 20, FMT_LA,   4, 1, 1, 23 }, {16, 6, 20, FMT_LA,   2, 1, 2, 23 },
    {17, 6, 20, FMT_LA,   2, 2, 1, 23 }, {18, 4, 11, FMT_RGB,  5, 1, 1, 15 },
};

static int mode_has_alpha(const Mode *m) { return m->format != FMT_RGB; }                   /* :456 */
static int mode_has_blue(const Mode *m) { return m->format != FMT_LA; }                     /* :463 */
static unsigned mode_channel_count(const Mode *m) { return m->format == FMT_RGB ? 3 : m->format == FMT_RGBA ? 4 : 2; }
static unsigned mode_endpoint_count(const Mode *m) { return mode_channel_count(m) * m->subset_count * 2; }  /* :478 */
static unsigned mode_weight_count(const Mode *m) { return m->plane_count * 16u; }           /* :482 */

/* src/target_formats/astc.rs:299-331  BISE_RANGES */
typedef struct { uint8_t bits, trits, quints; const char *deq_b; uint8_t deq_c; } BiseCounts;
static const BiseCounts BISE_RANGES[21] = {
    {1,0,0,"         ",0}, {0,1,0,"         ",0}, {2,0,0,"         ",0}, {0,0,1,"         ",0},
    {1,1,0,"000000000",204}, {3,0,0,"         ",0}, {1,0,1,"000000000",113}, {2,1,0,"b000b0bb0",93},
    {4,0,0,"         ",0}, {2,0,1,"b0000bb00",54}, {3,1,0,"cb000cbcb",44}, {5,0,0,"         ",0},
    {3,0,1,"cb0000cbc",26}, {4,1,0,"dcb000dcb",22}, {6,0,0,"         ",0}, {4,0,1,"dcb0000dc",13},
    {5,1,0,"edcb000ed",11}, {7,0,0,"         ",0}, {5,0,1,"edcb0000e",6}, {6,1,0,"fedcb000f",5},
    {8,0,0,"         ",0},
};

typedef struct { uint8_t c[4]; } Color32;                                                    /* src/color.rs:5 */
static Color32 color32(uint8_t r, uint8_t g, uint8_t b, uint8_t a) { Color32 c = {{r, g, b, a}}; return c; }

/* uastc.rs:329-341 decode_mode */
static int decode_mode(BitReader *r, const Mode **out)
{
    uint32_t code = br_peek(r, 7);
    unsigned idx = ORC_MODE_LUT[code];
    if (idx >= 19) return ORC_ERR_MODE;            /* "invalid mode index" */
    *out = &MODES[idx];
    br_remove(r, MODES[idx].code_size);
    return ORC_OK;
}

/* uastc.rs:343-350 decode_compsel */
static unsigned decode_compsel(BitReader *r, const Mode *m)
{
    if (m->plane_count == 2 && m->format == FMT_LA) return 3;
    if (m->plane_count == 2) return br_read(r, 2);
    return 0;
}

/* uastc.rs:352-366 decode_pattern_index */
static int decode_pattern_index(BitReader *r, const Mode *m, unsigned *pat)
{
    unsigned idx, count;
    if (m->id == 7) { idx = br_read(r, 5); count = 19; }
    else if (m->subset_count == 1) { *pat = 0; return ORC_OK; }
    else if (m->subset_count == 2) { idx = br_read(r, 5); count = 30; }
    else { idx = br_read(r, 4); count = 11; }
    if (idx < count) { *pat = idx; return ORC_OK; }
    return ORC_ERR_PATTERN;                        /* "block pattern is not valid" */
}

static const uint8_t ZERO16[16] = {0};
/* uastc.rs:368-376 get_pattern */
static const uint8_t *get_pattern(const Mode *m, unsigned pat)
{
    if (m->id == 7) return ORC_PATTERNS_2_3[pat];
    if (m->subset_count == 1) return ZERO16;
    if (m->subset_count == 2) return ORC_PATTERNS_2[pat];
    return ORC_PATTERNS_3[pat];
}
/* uastc.rs:378-385 get_anchor_weight_indices */
static unsigned get_anchors(const Mode *m, unsigned pat, const uint8_t **a)
{
    if (m->id == 7) { *a = ORC_PATTERNS_2_3_ANCHORS[pat]; return 2; }
    if (m->subset_count == 1) { *a = ZERO16; return 1; }
    if (m->subset_count == 2) { *a = ORC_PATTERNS_2_ANCHORS[pat]; return 2; }
    *a = ORC_PATTERNS_3_ANCHORS[pat]; return 3;
}

/* uastc.rs:387-394 decode_mode8_rgba */
static Color32 decode_mode8_rgba(BitReader *r)
{
    uint8_t R = (uint8_t)br_read(r, 8), G = (uint8_t)br_read(r, 8), B = (uint8_t)br_read(r, 8), A = (uint8_t)br_read(r, 8);
    return color32(R, G, B, A);
}

typedef struct { int etc1d; uint8_t etc1i, etc1s, etc1r, etc1g, etc1b; } Mode8Etc1Flags;
/* uastc.rs:400-409 */
static Mode8Etc1Flags decode_mode8_etc1_flags(BitReader *r)
{
    Mode8Etc1Flags f;
    f.etc1d = br_read(r, 1) == 1;
    f.etc1i = (uint8_t)br_read(r, 3);
    f.etc1s = (uint8_t)br_read(r, 2);
    f.etc1r = (uint8_t)br_read(r, 5);
    f.etc1g = (uint8_t)br_read(r, 5);
    f.etc1b = (uint8_t)br_read(r, 5);
    return f;
}

typedef struct { int bc1h0, bc1h1, etc1f, etc1d; uint8_t etc1i0, etc1i1; int has_bias; uint8_t etc1bias, etc2tm; } TransFlags;
/* uastc.rs:411-436 decode_trans_flags */
static TransFlags decode_trans_flags(BitReader *r, const Mode *m)
{
    TransFlags f;
    int m10_12 = m->id >= 10 && m->id <= 12;
    f.bc1h0 = br_read(r, 1) == 1;
    f.bc1h1 = m10_12 ? 0 : (br_read(r, 1) == 1);
    f.etc1f = br_read(r, 1) == 1;
    f.etc1d = br_read(r, 1) == 1;
    f.etc1i0 = (uint8_t)br_read(r, 3);
    f.etc1i1 = (uint8_t)br_read(r, 3);
    f.has_bias = !m10_12;
    f.etc1bias = m10_12 ? 0 : (uint8_t)br_read(r, 5);
    f.etc2tm = mode_has_alpha(m) ? (uint8_t)br_read(r, 8) : 0;
    return f;
}
/* uastc.rs:438-441 */
static void skip_trans_flags(BitReader *r, const Mode *m) { br_remove(r, m->trans_flags_bits); }

typedef struct { uint8_t trit_quint, bits; } QuantEndpoint;                                   /* uastc.rs:579-583 */

/* uastc.rs:585-614 unquant_endpoint */
static uint8_t unquant_endpoint(QuantEndpoint q, unsigned range_index)
{
    const BiseCounts *range = &BISE_RANGES[range_index];
    uint16_t quant_bits = q.bits;
    if (range->trits == 0 && range->quints == 0 && range->bits > 0) {
        uint16_t bits_la = (uint16_t)(quant_bits << (8 - range->bits));
        uint16_t val = 0;
        while (bits_la > 0) { val |= bits_la; bits_la >>= range->bits; }
        return (uint8_t)val;
    } else {
        uint16_t a = (quant_bits & 1) ? 511 : 0;
        uint16_t b = 0;
        for (int j = 0; j < 9; j++) {
            b <<= 1;
            char shift = range->deq_b[j];
            if (shift != '0') b |= (quant_bits >> (shift - 'a')) & 1;
        }
        uint16_t c = range->deq_c;
        uint16_t d = q.trit_quint;
        uint16_t val = (uint16_t)(d * c + b);
        val ^= a;
        return (uint8_t)((a & 0x80) | (val >> 2));
    }
}

/* uastc.rs:616-695 decode_endpoints */
static void decode_endpoints(BitReader *r, unsigned range_index, unsigned value_count, QuantEndpoint out[18])
{
    memset(out, 0, 18 * sizeof(QuantEndpoint));
    const BiseCounts *range = &BISE_RANGES[range_index];
    unsigned bit_count = range->bits;

    if (range->quints > 0) {
        unsigned out_pos = 0;
        for (unsigned g = 0; g < value_count / 3; g++) {
            uint8_t quints = (uint8_t)br_read(r, 7);
            for (int k = 0; k < 3; k++) { out[out_pos].trit_quint = quints % 5; quints /= 5; out_pos++; }
        }
        unsigned remaining = value_count - out_pos;
        if (remaining > 0) {
            unsigned bits_used = remaining == 1 ? 3 : 5;
            uint8_t quints = (uint8_t)br_read(r, bits_used);
            for (unsigned k = 0; k < remaining; k++) { out[out_pos].trit_quint = quints % 5; quints /= 5; out_pos++; }
        }
    }
    if (range->trits > 0) {
        unsigned out_pos = 0;
        for (unsigned g = 0; g < value_count / 5; g++) {
            uint8_t trits = (uint8_t)br_read(r, 8);
            for (int k = 0; k < 5; k++) { out[out_pos].trit_quint = trits % 3; trits /= 3; out_pos++; }
        }
        unsigned remaining = value_count - out_pos;
        if (remaining > 0) {
            static const unsigned used[5] = {0, 2, 4, 5, 7};
            uint8_t trits = (uint8_t)br_read(r, used[remaining]);
            for (unsigned k = 0; k < remaining; k++) { out[out_pos].trit_quint = trits % 3; trits /= 3; out_pos++; }
        }
    }
    if (bit_count > 0)
        for (unsigned i = 0; i < value_count; i++) out[i].bits = (uint8_t)br_read(r, bit_count);
}

/* uastc.rs:697-719 unquant_weights */
static void unquant_weights(uint8_t *w, unsigned n, unsigned weight_bits)
{
    static const uint8_t LUT1[2] = {0, 64};
    static const uint8_t LUT2[4] = {0, 21, 43, 64};
    static const uint8_t LUT3[8] = {0, 9, 18, 27, 37, 46, 55, 64};
    static const uint8_t LUT4[16] = {0, 4, 8, 12, 17, 21, 25, 29, 35, 39, 43, 47, 52, 56, 60, 64};
    static const uint8_t LUT5[32] = {0, 2, 4, 6, 8, 10, 12, 14, 16, 18, 20, 22, 24, 26, 28, 30,
                                     34, 36, 38, 40, 42, 44, 46, 48, 50, 52, 54, 56, 58, 60, 62, 64};
    const uint8_t *lut = weight_bits == 1 ? LUT1 : weight_bits == 2 ? LUT2 : weight_bits == 3 ? LUT3
                       : weight_bits == 4 ? LUT4 : LUT5;
    for (unsigned i = 0; i < n; i++) w[i] = lut[w[i]];
}

/* uastc.rs:721-740 decode_weights: raw weights, texel-major plane-minor; anchors one bit short */
static void decode_weights(BitReader *r, const Mode *m, unsigned pat, uint8_t out[32])
{
    unsigned plane_count = m->plane_count;
    const uint8_t *anchors; unsigned na = get_anchors(m, pat, &anchors);
    uint8_t bits[16];
    for (int i = 0; i < 16; i++) bits[i] = m->weight_bits;
    for (unsigned a = 0; a < na; a++) bits[anchors[a]] = (uint8_t)(m->weight_bits - 1);
    for (unsigned i = 0; i < 16; i++)
        for (unsigned p = 0; p < plane_count; p++)
            out[plane_count * i + p] = (uint8_t)br_read(r, bits[i]);
}

/* uastc.rs:176-216 assemble_endpoint_pairs */
static void assemble_endpoint_pairs(const Mode *m, const uint8_t *eb, unsigned n, Color32 pairs[3][2])
{
    memset(pairs, 0, sizeof(Color32) * 6);
    unsigned chunk = m->format == FMT_RGB ? 6 : m->format == FMT_RGBA ? 8 : 4;
    for (unsigned s = 0; s < 3 && (s + 1) * chunk <= n; s++) {
        const uint8_t *b = eb + s * chunk;
        if (m->format == FMT_RGB) { pairs[s][0] = color32(b[0], b[2], b[4], 0xFF); pairs[s][1] = color32(b[1], b[3], b[5], 0xFF); }
        else if (m->format == FMT_RGBA) { pairs[s][0] = color32(b[0], b[2], b[4], b[6]); pairs[s][1] = color32(b[1], b[3], b[5], b[7]); }
        else { pairs[s][0] = color32(b[0], b[0], b[0], b[2]); pairs[s][1] = color32(b[1], b[1], b[1], b[3]); }
    }
}

/* uastc.rs:218-235 astc_interpolate */
static uint8_t astc_interpolate(uint8_t l8, uint8_t h8, uint8_t w8, int srgb)
{
    uint32_t l = l8, h = h8, w = w8;
    if (srgb) { l = (l << 8) | 0x80; h = (h << 8) | 0x80; }
    else { l = (l << 8) | l; h = (h << 8) | h; }
    uint32_t k = (l * (64 - w) + h * w + 32) >> 6;
    return (uint8_t)(k >> 8);
}

/* uastc.rs:237-327 decode_block_to_rgba */
static int decode_block_to_rgba(const uint8_t bytes[16], Color32 output[16])
{
    BitReader rd; br_init(&rd, bytes, 16);
    const Mode *mode;
    int e = decode_mode(&rd, &mode);
    if (e) return e;
    if (mode->id == 8) {
        Color32 c = decode_mode8_rgba(&rd);
        for (int i = 0; i < 16; i++) output[i] = c;
        return ORC_OK;
    }
    skip_trans_flags(&rd, mode);
    unsigned compsel = decode_compsel(&rd, mode);
    unsigned pat;
    e = decode_pattern_index(&rd, mode, &pat);
    if (e) return e;

    unsigned endpoint_count = mode_endpoint_count(mode);
    unsigned weight_count = mode_weight_count(mode);
    uint8_t endpoints[18] = {0};
    uint8_t weights[32] = {0};
    QuantEndpoint q[18];
    decode_endpoints(&rd, mode->endpoint_range_index, endpoint_count, q);
    for (unsigned i = 0; i < endpoint_count; i++) endpoints[i] = unquant_endpoint(q[i], mode->endpoint_range_index);
    decode_weights(&rd, mode, pat, weights);
    unquant_weights(weights, weight_count, mode->weight_bits);

    const int srgb = 0;
    Color32 pairs[3][2];
    assemble_endpoint_pairs(mode, endpoints, endpoint_count, pairs);
    if (mode->subset_count == 1) {
        const Color32 e0 = pairs[0][0], e1 = pairs[0][1];
        if (mode->plane_count == 1) {
            for (int i = 0; i < 16; i++) {
                uint8_t w = weights[i];
                output[i] = color32(astc_interpolate(e0.c[0], e1.c[0], w, srgb), astc_interpolate(e0.c[1], e1.c[1], w, srgb),
                                    astc_interpolate(e0.c[2], e1.c[2], w, srgb), astc_interpolate(e0.c[3], e1.c[3], w, 0));
            }
        } else {
            for (int i = 0; i < 16; i++) {
                const uint8_t *ws = &weights[2 * i];
                uint8_t wr = compsel == 0 ? ws[1] : ws[0];
                uint8_t wg = compsel == 1 ? ws[1] : ws[0];
                uint8_t wb = compsel == 2 ? ws[1] : ws[0];
                uint8_t wa = compsel == 3 ? ws[1] : ws[0];
                output[i] = color32(astc_interpolate(e0.c[0], e1.c[0], wr, srgb), astc_interpolate(e0.c[1], e1.c[1], wg, srgb),
                                    astc_interpolate(e0.c[2], e1.c[2], wb, srgb), astc_interpolate(e0.c[3], e1.c[3], wa, 0));
            }
        }
    } else {
        const uint8_t *pattern = get_pattern(mode, pat);
        for (int i = 0; i < 16; i++) {
            const Color32 e0 = pairs[pattern[i]][0], e1 = pairs[pattern[i]][1];
            uint8_t w = weights[i];
            output[i] = color32(astc_interpolate(e0.c[0], e1.c[0], w, srgb), astc_interpolate(e0.c[1], e1.c[1], w, srgb),
                                astc_interpolate(e0.c[2], e1.c[2], w, srgb), astc_interpolate(e0.c[3], e1.c[3], w, 0));
        }
    }
    return ORC_OK;
}

/* ------------------------------------------------------------------------------------------
 * src/target_formats/astc.rs:8-181 convert_block_from_uastc (UASTC -> ASTC 4x4)
 * ---------------------------------------------------------------------------------------- */
static int astc_from_uastc(const uint8_t bytes[16], uint8_t output[16])
{
    BitReader rd; br_init(&rd, bytes, 16);
    const Mode *mode;
    int e = decode_mode(&rd, &mode);
    if (e) return e;
    memset(output, 0, 16);
    BitWriter wr; bw_init(&wr, output, 16);

    if (mode->id == 8) {                                            /* astc.rs:17-43 */
        Color32 rgba = decode_mode8_rgba(&rd);
        bw_write(&wr, 12, 0xDFC);
        bw_write(&wr, 20, 0x000FFFFF);
        bw_write(&wr, 32, 0xFFFFFFFFu);
        for (int c = 0; c < 4; c++) { uint16_t v = rgba.c[c]; bw_write(&wr, 16, (uint16_t)(v << 8 | v)); }
        return ORC_OK;
    }
    skip_trans_flags(&rd, mode);
    unsigned compsel = decode_compsel(&rd, mode);
    unsigned pat;
    e = decode_pattern_index(&rd, mode, &pat);
    if (e) return e;
    unsigned endpoint_count = mode_endpoint_count(mode);
    QuantEndpoint q[18];
    decode_endpoints(&rd, mode->endpoint_range_index, endpoint_count, q);

    int invert_subset_weights[3] = {0, 0, 0};
    if (mode_has_blue(mode)) {                                      /* astc.rs:55-78 */
        unsigned per = endpoint_count / mode->subset_count;
        for (unsigned s = 0; s < mode->subset_count; s++) {
            QuantEndpoint *qs = q + s * per;
            uint8_t ep[6] = {0};
            for (unsigned i = 0; i < 6 && i < per; i++) ep[i] = unquant_endpoint(qs[i], mode->endpoint_range_index);
            uint32_t s0 = (uint32_t)ep[0] + ep[2] + ep[4];
            uint32_t s1 = (uint32_t)ep[1] + ep[3] + ep[5];
            if (s0 > s1) {
                invert_subset_weights[s] = 1;
                for (unsigned i = 0; i + 1 < per; i += 2) { QuantEndpoint t = qs[i]; qs[i] = qs[i + 1]; qs[i + 1] = t; }
            }
        }
    }
    /* astc.rs:80-96 block mode, partition, CEM */
    bw_write(&wr, 13, ORC_UASTC_TO_ASTC_BLOCK_MODE_13[mode->id]);
    {
        int have = 1; uint16_t seed = 0;
        if (mode->id == 7) seed = ORC_PATTERNS_2_3_ASTC_INDEX_10[pat];
        else if (mode->subset_count == 1) have = 0;
        else if (mode->subset_count == 2) seed = ORC_PATTERNS_2_ASTC_INDEX_10[pat];
        else seed = ORC_PATTERNS_3_ASTC_INDEX_10[pat];
        if (have) { bw_write(&wr, 10, seed); bw_write(&wr, 2, 0); }
        unsigned cem = mode->format == FMT_RGB ? 8 : mode->format == FMT_RGBA ? 12 : 4;
        bw_write(&wr, 4, cem);
    }
    {   /* astc.rs:98-141 endpoints; iterates all 18 slots (tail slots are zero) */
        const BiseCounts *range = &BISE_RANGES[mode->endpoint_range_index];
        unsigned bit_count = range->bits;
        if (range->quints > 0) {
            for (unsigned base = 0; base < 18; base += 3) {
                unsigned n = 18 - base < 3 ? 18 - base : 3;
                unsigned id = 0;
                for (int k = (int)n - 1; k >= 0; k--) id = id * 5 + q[base + k].trit_quint;
                uint8_t Q = ORC_ASTC_QUINT_ENCODE_LUT[id];
                bw_write(&wr, bit_count, 0 < n ? q[base + 0].bits : 0); bw_write(&wr, 3, Q);
                bw_write(&wr, bit_count, 1 < n ? q[base + 1].bits : 0); bw_write(&wr, 2, Q >> 3);
                bw_write(&wr, bit_count, 2 < n ? q[base + 2].bits : 0); bw_write(&wr, 2, Q >> 5);
            }
        } else if (range->trits > 0) {
            for (unsigned base = 0; base < 18; base += 5) {
                unsigned n = 18 - base < 5 ? 18 - base : 5;
                unsigned id = 0;
                for (int k = (int)n - 1; k >= 0; k--) id = id * 3 + q[base + k].trit_quint;
                uint8_t T = ORC_ASTC_TRIT_ENCODE_LUT[id];
                bw_write(&wr, bit_count, 0 < n ? q[base + 0].bits : 0); bw_write(&wr, 2, T);
                bw_write(&wr, bit_count, 1 < n ? q[base + 1].bits : 0); bw_write(&wr, 2, T >> 2);
                bw_write(&wr, bit_count, 2 < n ? q[base + 2].bits : 0); bw_write(&wr, 1, T >> 4);
                bw_write(&wr, bit_count, 3 < n ? q[base + 3].bits : 0); bw_write(&wr, 2, T >> 5);
                bw_write(&wr, bit_count, 4 < n ? q[base + 4].bits : 0); bw_write(&wr, 1, T >> 7);
            }
        } else {
            for (unsigned i = 0; i < 18; i++) bw_write(&wr, bit_count, q[i].bits);
        }
    }
    {   /* astc.rs:143-178 weights from the top, bit-reversed; then CCS */
        BitWriterMsb wm; bwm_init(&wm, output, 16);
        uint8_t raw[32];
        decode_weights(&rd, mode, pat, raw);
        unsigned wc = mode_weight_count(mode);
        const uint8_t *pattern = get_pattern(mode, pat);
        for (unsigned i = 0; i < wc; i++) {
            unsigned texel = i / mode->plane_count;
            unsigned subset = mode->subset_count == 1 ? 0 : pattern[texel];
            uint8_t w = raw[i];
            if (invert_subset_weights[subset]) w = (uint8_t)~w;
            bwm_write_rev(&wm, mode->weight_bits, w);
        }
        if (mode->plane_count != 1) bwm_write(&wm, 2, compsel);
    }
    return ORC_OK;
}

/* ------------------------------------------------------------------------------------------
 * src/target_formats/bc7.rs
 * ---------------------------------------------------------------------------------------- */
typedef struct { uint8_t id, pat_bits, endpoint_count, color_bits, alpha_bits, weight_bits, plane_count, subset_count, p_bits, sp_bits; } Bc7Mode;
static const Bc7Mode BC7_MODES[8] = {                      /* bc7.rs:570-579 */
    {0, 4, 18, 4, 0, 3, 1, 3, 1, 0}, {1, 6, 12, 6, 0, 3, 1, 2, 0, 1}, {2, 6, 18, 5, 0, 2, 1, 3, 0, 0},
    {3, 6, 12, 7, 0, 2, 1, 2, 1, 0}, {4, 0,  8, 5, 6, 2, 2, 1, 0, 0}, {5, 0,  8, 7, 8, 2, 2, 1, 0, 0},
    {6, 0,  8, 7, 7, 4, 1, 1, 1, 0}, {7, 6, 16, 5, 5, 2, 1, 2, 1, 0},
};

/* bc7.rs:1126-1136 */
static void mode6_optimal_endpoint(uint8_t c, int p_bit, uint8_t *lo, uint8_t *hi)
{
    unsigned i = (unsigned)c + (p_bit ? 0 : 1);
    *lo = ORC_BC7_MODE_6_OPTIMAL_ENDPOINTS[i][0]; *hi = ORC_BC7_MODE_6_OPTIMAL_ENDPOINTS[i][1];
}
static uint32_t mode6_optimal_endpoint_err(uint8_t c, int p_bit) { return ((c == 0 && p_bit) || (c == 255 && !p_bit)) ? 1 : 0; }

/* bc7.rs:312-375 */
static void convert_mode8_to_bc7(Color32 solid, unsigned *mode, Color32 endpoint[2], uint8_t p_bits[2], uint8_t weights[2])
{
    uint32_t err0 = 0, err1 = 0;
    for (int c = 0; c < 4; c++) { err0 += mode6_optimal_endpoint_err(solid.c[c], 0); err1 += mode6_optimal_endpoint_err(solid.c[c], 1); }
    memset(endpoint, 0, sizeof(Color32) * 2); p_bits[0] = p_bits[1] = 0; weights[0] = weights[1] = 0;
    if (err0 > 0 && err1 > 0) {
        *mode = 5;
        for (int c = 0; c < 3; c++) {
            endpoint[0].c[c] = ORC_BC7_MODE_5_OPTIMAL_ENDPOINTS[solid.c[c]][0];
            endpoint[1].c[c] = ORC_BC7_MODE_5_OPTIMAL_ENDPOINTS[solid.c[c]][1];
        }
        endpoint[0].c[3] = solid.c[3]; endpoint[1].c[3] = solid.c[3];
        weights[0] = 1; weights[1] = 0;
    } else {
        *mode = 6;
        int best_p = err1 < err0;
        for (int c = 0; c < 4; c++) mode6_optimal_endpoint(solid.c[c], best_p, &endpoint[0].c[c], &endpoint[1].c[c]);
        p_bits[0] = p_bits[1] = (uint8_t)best_p;
        weights[0] = weights[1] = 5;
    }
}

/* bc7.rs:377-398 */
static void convert_weights_to_bc7(uint8_t w[16], unsigned ub, unsigned bb)
{
    static const uint8_t U1B2[2] = {0, 3};
    static const uint8_t U2B4[4] = {0, 5, 10, 15};
    static const uint8_t U3B4[8] = {0, 2, 4, 6, 9, 11, 13, 15};
    static const uint8_t U5B4[32] = {0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 6, 7, 8, 9, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13, 14, 14, 15, 15};
    const uint8_t *lut;
    if (ub == 1 && bb == 2) lut = U1B2; else if (ub == 2 && bb == 4) lut = U2B4;
    else if (ub == 3 && bb == 4) lut = U3B4; else if (ub == 5 && bb == 4) lut = U5B4;
    else return;   /* a == b */
    for (int i = 0; i < 16; i++) w[i] = lut[w[i]];
}

static int32_t clamp_i32(int32_t v, int32_t lo, int32_t hi) { return v < lo ? lo : v > hi ? hi : v; }

/* bc7.rs:408-475 determine_shared_pbits -- literal f32 */
static void determine_shared_pbits(unsigned total_comps, unsigned comp_bits, Color32 pair[2], uint8_t out[2])
{
    unsigned total_bits = comp_bits + 1;
    int32_t iscalep = (1 << total_bits) - 1;
    float scalep = (float)iscalep;
    float xl[4], xh[4];
    for (int i = 0; i < 4; i++) { xl[i] = (float)pair[0].c[i] / 255.f; xh[i] = (float)pair[1].c[i] / 255.f; }
    memset(pair, 0, sizeof(Color32) * 2);
    float best_err = 1e+9f;
    uint8_t s_bit = 0;
    for (int32_t p = 0; p < 2; p++) {
        Color32 xmin, xmax;
        for (int c = 0; c < 4; c++) {
            xmin.c[c] = (uint8_t)clamp_i32((int32_t)((xl[c] * scalep - (float)p) / 2.f + 0.5f) * 2 + p, p, iscalep - 1 + p);
            xmax.c[c] = (uint8_t)clamp_i32((int32_t)((xh[c] * scalep - (float)p) / 2.f + 0.5f) * 2 + p, p, iscalep - 1 + p);
        }
        Color32 sl, sh;
        for (int i = 0; i < 4; i++) {
            sl.c[i] = (uint8_t)(xmin.c[i] << (8 - total_bits)); sl.c[i] |= sl.c[i] >> total_bits;
            sh.c[i] = (uint8_t)(xmax.c[i] << (8 - total_bits)); sh.c[i] |= sh.c[i] >> total_bits;
        }
        float err = 0.f;
        for (unsigned i = 0; i < total_comps; i++) {
            float a = (float)sl.c[i] / 255.f - xl[i];
            float b = (float)sh.c[i] / 255.f - xh[i];
            float a2 = a * a, b2 = b * b;
            float t = a2 + b2;
            err = err + t;
        }
        if (err < best_err) {
            best_err = err; s_bit = (uint8_t)p;
            for (int j = 0; j < 4; j++) { pair[0].c[j] = xmin.c[j] >> 1; pair[1].c[j] = xmax.c[j] >> 1; }
        }
    }
    out[0] = s_bit; out[1] = s_bit;
}

/* bc7.rs:478-553 determine_unique_pbits -- literal f32 */
static void determine_unique_pbits(unsigned total_comps, unsigned comp_bits, Color32 pair[2], uint8_t out[2])
{
    unsigned total_bits = comp_bits + 1;
    int32_t iscalep = (1 << total_bits) - 1;
    float scalep = (float)iscalep;
    float xl[4], xh[4];
    for (int i = 0; i < 4; i++) { xl[i] = (float)pair[0].c[i] / 255.f; xh[i] = (float)pair[1].c[i] / 255.f; }
    memset(pair, 0, sizeof(Color32) * 2);
    float best_err0 = 1e+9f, best_err1 = 1e+9f;
    out[0] = out[1] = 0;
    for (int32_t p = 0; p < 2; p++) {
        Color32 xmin, xmax;
        for (int c = 0; c < 4; c++) {
            xmin.c[c] = (uint8_t)clamp_i32((int32_t)((xl[c] * scalep - (float)p) / 2.f + 0.5f) * 2 + p, p, iscalep - 1 + p);
            xmax.c[c] = (uint8_t)clamp_i32((int32_t)((xh[c] * scalep - (float)p) / 2.f + 0.5f) * 2 + p, p, iscalep - 1 + p);
        }
        Color32 sl, sh;
        for (int i = 0; i < 4; i++) {
            /* u8 << then wrapping_shr(total_bits): shift amount mod 8 (quirk C-10: 8 -> 0) */
            sl.c[i] = (uint8_t)(xmin.c[i] << (8 - total_bits)); sl.c[i] |= (uint8_t)(sl.c[i] >> (total_bits & 7));
            sh.c[i] = (uint8_t)(xmax.c[i] << (8 - total_bits)); sh.c[i] |= (uint8_t)(sh.c[i] >> (total_bits & 7));
        }
        float err0 = 0.f, err1 = 0.f;
        for (unsigned i = 0; i < total_comps; i++) {
            float a = (float)sl.c[i] - xl[i] * 255.f;
            float b = (float)sh.c[i] - xh[i] * 255.f;
            err0 = err0 + a * a;
            err1 = err1 + b * b;
        }
        if (err0 < best_err0) { best_err0 = err0; out[0] = (uint8_t)p; for (int j = 0; j < 4; j++) pair[0].c[j] = xmin.c[j] >> 1; }
        if (err1 < best_err1) { best_err1 = err1; out[1] = (uint8_t)p; for (int j = 0; j < 4; j++) pair[1].c[j] = xmax.c[j] >> 1; }
    }
}

/* bc7.rs:9-310 convert_block_from_uastc (UASTC -> BC7) */
static int bc7_from_uastc(const uint8_t bytes[16], uint8_t output[16])
{
    enum { ALPHA = 3 };
    BitReader rd; br_init(&rd, bytes, 16);
    const Mode *mode;
    int e = decode_mode(&rd, &mode);
    if (e) return e;
    memset(output, 0, 16);
    BitWriter wr; bw_init(&wr, output, 16);

    if (mode->id == 8) {                                            /* bc7.rs:18-59 */
        Color32 rgba = decode_mode8_rgba(&rd);
        unsigned m; Color32 ep[2]; uint8_t pb[2], wts[2];
        convert_mode8_to_bc7(rgba, &m, ep, pb, wts);
        const Bc7Mode *bm = &BC7_MODES[m];
        bw_write(&wr, m + 1, 1u << m);
        if (m == 5) bw_write(&wr, 2, 0);
        for (int ch = 0; ch < 4; ch++) {
            unsigned bc = ch != ALPHA ? bm->color_bits : bm->alpha_bits;
            bw_write(&wr, bc, ep[0].c[ch]); bw_write(&wr, bc, ep[1].c[ch]);
        }
        if (m == 6) bw_write(&wr, 2, (uint32_t)(pb[1] << 1) | pb[0]);
        for (unsigned pl = 0; pl < bm->plane_count; pl++) {
            bw_write(&wr, bm->weight_bits - 1u, wts[pl]);
            for (int i = 0; i < 15; i++) bw_write(&wr, bm->weight_bits, wts[pl]);
        }
        return ORC_OK;
    }

    unsigned bc7_mode_index = ORC_UASTC_TO_BC7_MODES[mode->id];
    const Bc7Mode *bm = &BC7_MODES[bc7_mode_index];
    skip_trans_flags(&rd, mode);
    unsigned compsel = decode_compsel(&rd, mode);
    unsigned uastc_pat;
    e = decode_pattern_index(&rd, mode, &uastc_pat);
    if (e) return e;

    unsigned bc7_endpoints_per_channel = 2u * bm->subset_count;
    unsigned bc7_channel_count = bm->endpoint_count / bc7_endpoints_per_channel;

    Color32 endpoints[3][2];
    {
        unsigned n = mode_endpoint_count(mode);
        QuantEndpoint q[18];
        decode_endpoints(&rd, mode->endpoint_range_index, n, q);
        uint8_t unq[18] = {0};
        for (unsigned i = 0; i < n; i++) unq[i] = unquant_endpoint(q[i], mode->endpoint_range_index);
        assemble_endpoint_pairs(mode, unq, 18, endpoints);       /* ref passes the whole 18-byte array */
    }

    uint8_t weights[2][16]; memset(weights, 0, sizeof weights);
    {
        uint8_t raw[32];
        decode_weights(&rd, mode, uastc_pat, raw);
        if (mode->plane_count == 1) {
            for (int i = 0; i < 16; i++) weights[0][i] = raw[i];
            convert_weights_to_bc7(weights[0], mode->weight_bits, bm->weight_bits);
        } else {
            for (int i = 0; i < 32; i++) weights[i & 1][i >> 1] = raw[i];
            convert_weights_to_bc7(weights[0], mode->weight_bits, bm->weight_bits);
            convert_weights_to_bc7(weights[1], mode->weight_bits, bm->weight_bits);
        }
    }

    unsigned nsub = bm->subset_count, nplanes = bm->plane_count;
    bw_write(&wr, bc7_mode_index + 1, 1u << bc7_mode_index);

    static const uint8_t ANCHOR0[1] = {0};
    const uint8_t *bc7_anchors = ANCHOR0; unsigned n_anchors = 1;

    if (bm->subset_count != 1) {                                    /* bc7.rs:113-198 */
        unsigned bc7_pat; const uint8_t *pattern; const uint8_t *anchors; unsigned na; uint8_t perm[3]; unsigned nperm;
        if (mode->id == 1) {
            bc7_pat = ORC_PATTERNS_2_BC7_INDEX_INV[0][0];
            pattern = ORC_PATTERNS_2_BC7[uastc_pat];
            anchors = ORC_PATTERNS_2_BC7_ANCHORS[bc7_pat]; na = 2;
            perm[0] = 0; perm[1] = 0; nperm = 2;
        } else if (mode->id == 7) {
            bc7_pat = ORC_PATTERNS_2_3_BC7_INDEX_PERM[uastc_pat][0];
            unsigned p = ORC_PATTERNS_2_3_BC7_INDEX_PERM[uastc_pat][1];
            memcpy(perm, ORC_PATTERNS_2_3_BC7_TO_ASTC_PERMUTATIONS[p], 3); nperm = 3;
            pattern = ORC_PATTERNS_2_3_BC7[uastc_pat];
            anchors = ORC_PATTERNS_3_BC7_ANCHORS[bc7_pat]; na = 3;
        } else if (mode->subset_count == 2) {
            bc7_pat = ORC_PATTERNS_2_BC7_INDEX_INV[uastc_pat][0];
            int inv = ORC_PATTERNS_2_BC7_INDEX_INV[uastc_pat][1];
            pattern = ORC_PATTERNS_2_BC7[uastc_pat];
            anchors = ORC_PATTERNS_2_BC7_ANCHORS[bc7_pat]; na = 2;
            if (inv) { perm[0] = 1; perm[1] = 0; } else { perm[0] = 0; perm[1] = 1; }
            nperm = 2;
        } else {
            bc7_pat = ORC_PATTERNS_3_BC7_INDEX_PERM[uastc_pat][0];
            unsigned p = ORC_PATTERNS_3_BC7_INDEX_PERM[uastc_pat][1];
            memcpy(perm, ORC_PATTERNS_3_BC7_TO_ASTC_PERMUTATIONS[p], 3); nperm = 3;
            pattern = ORC_PATTERNS_3_BC7[uastc_pat];
            anchors = ORC_PATTERNS_3_BC7_ANCHORS[bc7_pat]; na = 3;
        }
        bc7_anchors = anchors; n_anchors = na;
        bw_write(&wr, bm->pat_bits, bc7_pat);
        {   /* permute: dst[X] = src[perm[X]], zip stops at the shorter of perm / dst(3) */
            Color32 permuted[3][2]; memset(permuted, 0, sizeof permuted);
            for (unsigned x = 0; x < nperm && x < 3; x++) { permuted[x][0] = endpoints[perm[x]][0]; permuted[x][1] = endpoints[perm[x]][1]; }
            for (unsigned s = 0; s < nsub; s++) { endpoints[s][0] = permuted[s][0]; endpoints[s][1] = permuted[s][1]; }
        }
        {
            uint32_t weight_mask = mask32(bm->weight_bits);
            uint32_t msb = 1u << (bm->weight_bits - 1);
            int invert_subset[3] = {0, 0, 0};
            for (unsigned a = 0; a < na && a < 3; a++) invert_subset[a] = (weights[0][anchors[a]] & msb) != 0;
            for (unsigned s = 0; s < nsub; s++)
                if (invert_subset[s]) { Color32 t = endpoints[s][0]; endpoints[s][0] = endpoints[s][1]; endpoints[s][1] = t; }
            for (int i = 0; i < 16; i++)
                if (invert_subset[pattern[i]]) weights[0][i] = (uint8_t)(~weights[0][i] & weight_mask);
        }
    } else {                                                        /* bc7.rs:199-247 */
        uint32_t weight_mask = mask32(bm->weight_bits);
        uint32_t msb = 1u << (bm->weight_bits - 1);
        if (mode->plane_count == 1) {
            if (weights[0][0] & msb) {
                Color32 t = endpoints[0][0]; endpoints[0][0] = endpoints[0][1]; endpoints[0][1] = t;
                for (int i = 0; i < 16; i++) weights[0][i] = (uint8_t)(~weights[0][i] & weight_mask);
            }
        } else {
            int inv0 = (weights[0][0] & msb) != 0, inv1 = (weights[1][0] & msb) != 0;
            Color32 *pair = endpoints[0];
            for (int k = 0; k < 2; k++) { uint8_t t = pair[k].c[compsel]; pair[k].c[compsel] = pair[k].c[ALPHA]; pair[k].c[ALPHA] = t; }
            if (inv0) { Color32 t = pair[0]; pair[0] = pair[1]; pair[1] = t; }
            if (inv0 != inv1) { uint8_t t = pair[0].c[ALPHA]; pair[0].c[ALPHA] = pair[1].c[ALPHA]; pair[1].c[ALPHA] = t; }
            int inv[2] = {inv0, inv1};
            for (unsigned pl = 0; pl < nplanes; pl++)
                if (inv[pl]) for (int i = 0; i < 16; i++) weights[pl][i] = (uint8_t)(~weights[pl][i] & weight_mask);
            bw_write(&wr, 2, (compsel + 1) & 3);
            if (bm->id == 4) bw_write(&wr, 1, 0);
        }
    }

    unsigned color_bits = bm->color_bits, alpha_bits = bm->alpha_bits;
    uint8_t p_bits[3][2]; memset(p_bits, 0, sizeof p_bits);
    if (bm->p_bits != 0) {
        for (unsigned s = 0; s < nsub; s++) determine_unique_pbits(bc7_channel_count, bm->color_bits, endpoints[s], p_bits[s]);
    } else if (bm->sp_bits != 0) {
        for (unsigned s = 0; s < nsub; s++) determine_shared_pbits(bc7_channel_count, bm->color_bits, endpoints[s], p_bits[s]);
    } else {
        for (unsigned s = 0; s < nsub; s++)
            for (int k = 0; k < 2; k++) {
                for (int ch = 0; ch < 3; ch++)
                    endpoints[s][k].c[ch] = (uint8_t)(((uint32_t)endpoints[s][k].c[ch] * mask32(color_bits) + 127) / 255);
                endpoints[s][k].c[ALPHA] = (uint8_t)(((uint32_t)endpoints[s][k].c[ALPHA] * mask32(alpha_bits) + 127) / 255);
            }
    }
    for (unsigned ch = 0; ch < bc7_channel_count; ch++) {
        unsigned bc = ch != ALPHA ? color_bits : alpha_bits;
        for (unsigned s = 0; s < nsub; s++) { bw_write(&wr, bc, endpoints[s][0].c[ch]); bw_write(&wr, bc, endpoints[s][1].c[ch]); }
    }
    if (bm->p_bits != 0) {
        for (unsigned s = 0; s < nsub; s++) bw_write(&wr, 2, (uint32_t)(p_bits[s][1] << 1) | p_bits[s][0]);
    } else if (bm->sp_bits != 0) {
        bw_write(&wr, 2, (uint32_t)(p_bits[1][0] << 1) | p_bits[0][0]);
    }
    {
        uint8_t bit_counts[16];
        for (int i = 0; i < 16; i++) bit_counts[i] = bm->weight_bits;
        for (unsigned a = 0; a < n_anchors; a++) bit_counts[bc7_anchors[a]] -= 1;
        for (unsigned pl = 0; pl < nplanes; pl++)
            for (int i = 0; i < 16; i++) bw_write(&wr, bit_counts[i], weights[pl][i]);
    }
    return ORC_OK;
}

/* ------------------------------------------------------------------------------------------
 * src/target_formats/etc.rs
 * ---------------------------------------------------------------------------------------- */
static const uint8_t SELECTOR_ID_TO_ETC1[4] = {3, 2, 0, 1};                                  /* etc.rs:433 */
static const int16_t ETC1_MODIFIERS[8][4] = {                                                 /* etc.rs:436-445 */
    {-8, -2, 2, 8}, {-17, -5, 5, 17}, {-29, -9, 9, 29}, {-42, -13, 13, 42},
    {-60, -18, 18, 60}, {-80, -24, 24, 80}, {-106, -33, 33, 106}, {-183, -47, 47, 183},
};
static const int8_t ETC2_ALPHA_MODIFIERS[16][8] = {                                           /* etc.rs:451-468 */
    {-3, -6, -9, -15, 2, 5, 8, 14}, {-3, -7, -10, -13, 2, 6, 9, 12}, {-2, -5, -8, -13, 1, 4, 7, 12},
    {-2, -4, -6, -13, 1, 3, 5, 12}, {-3, -6, -8, -12, 2, 5, 7, 11}, {-3, -7, -9, -11, 2, 6, 8, 10},
    {-4, -7, -8, -11, 3, 6, 7, 10}, {-3, -5, -8, -11, 2, 4, 7, 10}, {-2, -6, -8, -10, 1, 5, 7, 9},
    {-2, -5, -8, -10, 1, 4, 7, 9}, {-2, -4, -8, -10, 1, 3, 7, 9}, {-2, -5, -7, -10, 1, 4, 6, 9},
    {-3, -4, -7, -10, 2, 3, 6, 9}, {-1, -2, -3, -10, 0, 1, 2, 9}, {-4, -6, -8, -9, 3, 5, 7, 8},
    {-3, -5, -7, -9, 2, 4, 6, 8},
};

typedef struct { uint8_t selectors[4]; uint8_t etc1_bytes[4]; } Selector;                    /* etc.rs:343-350 */

static unsigned selector_get(const Selector *s, unsigned x, unsigned y) { return (s->selectors[y] >> (2 * x)) & 3; }  /* :354 */
static void selector_set(Selector *s, unsigned x, unsigned y, uint8_t val)                    /* etc.rs:363-393 */
{
    unsigned shift = 2 * x;
    s->selectors[y] &= (uint8_t)~(3u << shift);
    s->selectors[y] |= (uint8_t)(val << shift);
    uint8_t mod_id = SELECTOR_ID_TO_ETC1[val];
    unsigned pixel_id = x * 4 + y;
    unsigned ms = 1 - (pixel_id / 8);
    unsigned ls = ms + 2;
    unsigned bit = pixel_id % 8;
    s->etc1_bytes[ls] &= (uint8_t)~(1u << bit);
    s->etc1_bytes[ls] |= (uint8_t)((mod_id % 2) << bit);
    s->etc1_bytes[ms] &= (uint8_t)~(1u << bit);
    s->etc1_bytes[ms] |= (uint8_t)((mod_id / 2) << bit);
}

static Color32 color_5_to_8(Color32 c) { return color32((uint8_t)((c.c[0] << 3) | (c.c[0] >> 2)), (uint8_t)((c.c[1] << 3) | (c.c[1] >> 2)), (uint8_t)((c.c[2] << 3) | (c.c[2] >> 2)), 255); } /* :396 */
static Color32 color_4_to_8(Color32 c) { return color32((uint8_t)((c.c[0] << 4) | c.c[0]), (uint8_t)((c.c[1] << 4) | c.c[1]), (uint8_t)((c.c[2] << 4) | c.c[2]), 255); }              /* :408 */

static uint8_t clamp255(int v) { return (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v); }
/* etc.rs:420-431 */
static void apply_mod_to_base_color(Color32 base, unsigned inten, Color32 colors[4])
{
    for (int i = 0; i < 4; i++) {
        int m = ETC1_MODIFIERS[inten][i];
        colors[i] = color32(clamp255(base.c[0] + m), clamp255(base.c[1] + m), clamp255(base.c[2] + m), 255);
    }
}

/* etc.rs:203-259 apply_etc1_bias */
static Color32 apply_etc1_bias(Color32 block_color, unsigned bias, uint32_t limit, uint32_t subblock)
{
    static const uint8_t S_DIVS[3] = {1, 3, 9};
    for (int c = 0; c < 3; c++) {
        int delta;
        switch (bias) {
        case 2:  delta = subblock == 1 ? 0 : (c == 0 ? -1 : 0); break;
        case 5:  delta = subblock == 1 ? 0 : (c == 1 ? -1 : 0); break;
        case 6:  delta = subblock == 1 ? 0 : (c == 2 ? -1 : 0); break;
        case 7:  delta = subblock == 1 ? 0 : (c == 0 ? 1 : 0); break;
        case 11: delta = subblock == 1 ? 0 : (c == 1 ? 1 : 0); break;
        case 15: delta = subblock == 1 ? 0 : (c == 2 ? 1 : 0); break;
        case 18: delta = subblock == 1 ? (c == 0 ? -1 : 0) : 0; break;
        case 19: delta = subblock == 1 ? (c == 1 ? -1 : 0) : 0; break;
        case 20: delta = subblock == 1 ? (c == 2 ? -1 : 0) : 0; break;
        case 21: delta = subblock == 1 ? (c == 0 ? 1 : 0) : 0; break;
        case 24: delta = subblock == 1 ? (c == 1 ? 1 : 0) : 0; break;
        case 8:  delta = subblock == 1 ? (c == 2 ? 1 : 0) : 0; break;
        case 10: delta = -2; break;
        case 27: delta = subblock == 1 ? 0 : -1; break;
        case 28: delta = subblock == 1 ? -1 : 1; break;
        case 29: delta = subblock == 1 ? 1 : 0; break;
        case 30: delta = subblock == 1 ? -1 : 0; break;
        case 31: delta = subblock == 1 ? 0 : 1; break;
        default: delta = (int)((bias / S_DIVS[c]) % 3) - 1; break;
        }
        int v = block_color.c[c];
        if (v == 0) { if (delta == -2) v += 3; else v += delta + 1; }
        else if (v == (int)limit) v += delta - 1;
        else { v += delta; if (v < 0 || v > (int)limit) v = (v - delta) - delta; }
        block_color.c[c] = (uint8_t)v;
    }
    return block_color;
}

/* etc.rs:261-275 */
static void write_solid_etc2_alpha_block(uint8_t out[8], uint8_t value)
{
    out[0] = value; out[1] = (1 << 4) | 13;
    out[2] = 0x92; out[3] = 0x49; out[4] = 0x24; out[5] = 0x92; out[6] = 0x49; out[7] = 0x24;
}

/* etc.rs:277-341 write_etc2_alpha_block -- literal f32 centre */
static void write_etc2_alpha_block(uint8_t out[8], uint8_t etc2tm, const Color32 rgba[16])
{
    if (etc2tm == 0) { write_solid_etc2_alpha_block(out, 255); return; }
    uint8_t min_alpha = 255, max_alpha = 0;
    for (int i = 0; i < 16; i++) { if (rgba[i].c[3] < min_alpha) min_alpha = rgba[i].c[3]; if (rgba[i].c[3] > max_alpha) max_alpha = rgba[i].c[3]; }
    if (min_alpha == max_alpha) { write_solid_etc2_alpha_block(out, min_alpha); return; }
    unsigned table_index = etc2tm & 15;
    int multiplier = etc2tm >> 4;
    const int8_t *mod_table = ETC2_ALPHA_MODIFIERS[table_index];
    int mod_min = mod_table[3], mod_max = mod_table[7];
    int range = mod_max - mod_min;
    float amt = -((float)mod_min) / (float)range;
    float a = (float)min_alpha, b = (float)max_alpha;
    float t0 = 1.0f - amt;
    float t1 = a * t0;
    float t2 = b * amt;
    float l = t1 + t2;
    int center = (int)roundf(l);
    uint8_t values[8];
    for (int i = 0; i < 8; i++) values[i] = clamp255(center + mod_table[i] * multiplier);
    uint64_t selectors = 0;
    for (int i = 0; i < 16; i++) {
        int al = rgba[i].c[3];
        int best = 0, best_d = abs((int)values[0] - al);
        for (int k = 1; k < 8; k++) { int d = abs((int)values[k] - al); if (d < best_d) { best_d = d; best = k; } }
        int x = i / 4, y = i % 4, id = y * 4 + x;
        selectors |= (uint64_t)best << (45 - id * 3);
    }
    out[0] = (uint8_t)center;
    out[1] = etc2tm;
    for (int k = 0; k < 6; k++) out[2 + k] = (uint8_t)(selectors >> (8 * (5 - k)));   /* to_be_bytes()[2..8] */
}

/* etc.rs:32-201 convert_block_from_uastc; alpha == NULL -> ETC1 only */
static int etc_from_uastc(const uint8_t bytes[16], uint8_t output[8], uint8_t *alpha)
{
    BitReader rd; br_init(&rd, bytes, 16);
    const Mode *mode;
    int e = decode_mode(&rd, &mode);
    if (e) return e;
    memset(output, 0, 8);
    BitWriter wr; bw_init(&wr, output, 8);

    if (mode->id == 8) {                                            /* etc.rs:43-76 */
        if (alpha) { Color32 rgba = decode_mode8_rgba(&rd); write_solid_etc2_alpha_block(alpha, rgba.c[3]); }
        else br_remove(&rd, 32);
        Mode8Etc1Flags f = decode_mode8_etc1_flags(&rd);
        if (!f.etc1d) {
            bw_write(&wr, 8, (uint8_t)(f.etc1r << 4 | f.etc1r)); bw_write(&wr, 8, (uint8_t)(f.etc1g << 4 | f.etc1g)); bw_write(&wr, 8, (uint8_t)(f.etc1b << 4 | f.etc1b));
        } else {
            bw_write(&wr, 8, (uint8_t)(f.etc1r << 3)); bw_write(&wr, 8, (uint8_t)(f.etc1g << 3)); bw_write(&wr, 8, (uint8_t)(f.etc1b << 3));
        }
        bw_write(&wr, 8, (uint8_t)(f.etc1i << 5 | f.etc1i << 2 | (f.etc1d ? 1 : 0) << 1));
        static const uint8_t SEL[4] = {3, 2, 0, 1};
        uint8_t selector = SEL[f.etc1s];
        uint16_t s_lo = selector & 1, s_hi = selector >> 1;
        bw_write(&wr, 16, (uint16_t)(0u - s_hi));
        bw_write(&wr, 16, (uint16_t)(0u - s_lo));
        return ORC_OK;
    }

    TransFlags tf = decode_trans_flags(&rd, mode);
    Color32 rgba[16];
    e = decode_block_to_rgba(bytes, rgba);
    if (e) return e;
    if (alpha) write_etc2_alpha_block(alpha, tf.etc2tm, rgba);

    if (!tf.etc1f)                                                  /* etc.rs:86-95 transpose */
        for (int y = 0; y < 3; y++)
            for (int x = y + 1; x < 4; x++) { Color32 t = rgba[y * 4 + x]; rgba[y * 4 + x] = rgba[x * 4 + y]; rgba[x * 4 + y] = t; }

    unsigned color_bits = !tf.etc1d ? 4 : 5;
    uint32_t limit = mask32(color_bits);
    Color32 avg[2]; memset(avg, 0, sizeof avg);
    for (int sb = 0; sb < 2; sb++) {
        uint16_t sum[4] = {0, 0, 0, 0};
        for (int i = 0; i < 8; i++) for (int c = 0; c < 4; c++) sum[c] = (uint16_t)(sum[c] + rgba[sb * 8 + i].c[c]);
        for (int c = 0; c < 3; c++) avg[sb].c[c] = (uint8_t)(((uint32_t)sum[c] * limit + 1020) / (8 * 255));
    }
    Color32 c0 = avg[0], c1 = avg[1];
    if (tf.has_bias) { c0 = apply_etc1_bias(avg[0], tf.etc1bias, limit, 0); c1 = apply_etc1_bias(avg[1], tf.etc1bias, limit, 1); }

    Color32 block_colors[2][4];
    if (!tf.etc1d) {
        bw_write(&wr, 8, (uint8_t)(c0.c[0] << 4 | c1.c[0])); bw_write(&wr, 8, (uint8_t)(c0.c[1] << 4 | c1.c[1])); bw_write(&wr, 8, (uint8_t)(c0.c[2] << 4 | c1.c[2]));
        apply_mod_to_base_color(color_4_to_8(c0), tf.etc1i0, block_colors[0]);
        apply_mod_to_base_color(color_4_to_8(c1), tf.etc1i1, block_colors[1]);
    } else {
        int16_t d[3];
        for (int c = 0; c < 3; c++) d[c] = (int16_t)clamp_i32((int)c1.c[c] - (int)c0.c[c], -4, 3);
        for (int c = 0; c < 3; c++) bw_write(&wr, 8, (uint8_t)(c0.c[c] << 3 | (uint8_t)(d[c] & 7)));
        Color32 c1d = color32((uint8_t)(c0.c[0] + d[0]), (uint8_t)(c0.c[1] + d[1]), (uint8_t)(c0.c[2] + d[2]), 255);
        apply_mod_to_base_color(color_5_to_8(c0), tf.etc1i0, block_colors[0]);
        apply_mod_to_base_color(color_5_to_8(c1d), tf.etc1i1, block_colors[1]);
    }
    bw_write(&wr, 8, (uint8_t)(tf.etc1i0 << 5 | tf.etc1i1 << 2 | (tf.etc1d ? 1 : 0) << 1 | (tf.etc1f ? 1 : 0)));

    Selector selector; memset(&selector, 0, sizeof selector);
    static const int LUM[3] = {108, 366, 38};
    for (int sb = 0; sb < 2; sb++) {
        int block_lums[4];
        for (int k = 0; k < 4; k++) block_lums[k] = block_colors[sb][k].c[0] * LUM[0] + block_colors[sb][k].c[1] * LUM[1] + block_colors[sb][k].c[2] * LUM[2];
        int l01 = (block_lums[0] + block_lums[1]) / 2, l12 = (block_lums[1] + block_lums[2]) / 2, l23 = (block_lums[2] + block_lums[3]) / 2;
        for (int i = 0; i < 8; i++) {
            const Color32 *c = &rgba[sb * 8 + i];
            int lum = c->c[0] * LUM[0] + c->c[1] * LUM[1] + c->c[2] * LUM[2];
            uint8_t sel = (uint8_t)((lum >= l01) + (lum >= l12) + (lum >= l23));
            unsigned x = i & 3, y = 2 * sb + (i >> 2);
            if (tf.etc1f) selector_set(&selector, x, y, sel); else selector_set(&selector, y, x, sel);
        }
    }
    uint32_t sb32 = (uint32_t)selector.etc1_bytes[0] | (uint32_t)selector.etc1_bytes[1] << 8 | (uint32_t)selector.etc1_bytes[2] << 16 | (uint32_t)selector.etc1_bytes[3] << 24;
    bw_write(&wr, 32, sb32);
    return ORC_OK;
}

/* ==========================================================================================
 * Exported block-level API (mirrors src/lib.rs:29-53)
 * ======================================================================================== */
ORC_API int orc_unpack_uastc_block_to_rgba(const uint8_t in[16], uint32_t out[16])
{
    Color32 px[16];
    int e = decode_block_to_rgba(in, px);
    if (e) return e;
    for (int i = 0; i < 16; i++)
        out[i] = (uint32_t)px[i].c[0] | (uint32_t)px[i].c[1] << 8 | (uint32_t)px[i].c[2] << 16 | (uint32_t)px[i].c[3] << 24;
    return ORC_OK;
}
ORC_API int orc_transcode_uastc_block_to_astc(const uint8_t in[16], uint8_t out[16]) { return astc_from_uastc(in, out); }
ORC_API int orc_transcode_uastc_block_to_bc7(const uint8_t in[16], uint8_t out[16]) { return bc7_from_uastc(in, out); }
ORC_API int orc_transcode_uastc_block_to_etc1(const uint8_t in[16], uint8_t out[8]) { return etc_from_uastc(in, out, NULL); }
ORC_API int orc_transcode_uastc_block_to_etc2(const uint8_t in[16], uint8_t out[16])
{
    memset(out, 0, 16);
    return etc_from_uastc(in, out + 8, out);      /* etc.rs:19-30: alpha = first 8 bytes */
}

/* target ids shared with include/b2bu.h: 0 RGBA, 1 ASTC, 2 BC7, 3 ETC1, 4 ETC2 */
static size_t out_block_bytes(int fmt) { return fmt == 0 ? 64 : fmt == 3 ? 8 : 16; }

/* src/uastc.rs:112-165 Decoder::transcode / transcode_into: first error aborts; *bad = block index */
static int transcode_range(int fmt, const uint8_t *data, size_t first, size_t last, uint8_t *out, size_t *bad)
{
    for (size_t i = first; i < last; i++) {
        int e;
        const uint8_t *b = data + 16 * i;
        switch (fmt) {
        case 1: e = astc_from_uastc(b, out + 16 * i); break;
        case 2: e = bc7_from_uastc(b, out + 16 * i); break;
        case 3: e = etc_from_uastc(b, out + 8 * i, NULL); break;
        case 4: memset(out + 16 * i, 0, 16); e = etc_from_uastc(b, out + 16 * i + 8, out + 16 * i); break;
        default: return ORC_ERR_ARG;
        }
        if (e) { if (bad) *bad = i; return e; }
    }
    return ORC_OK;
}
/* src/uastc.rs:89-110 Decoder::decode_to_rgba: row-major image, pitch 4*blocks_per_row pixels */
static int rgba_range(const uint8_t *data, size_t first, size_t last, size_t bpr, uint32_t *out, size_t *bad)
{
    size_t stride = 4 * bpr;
    for (size_t i = first; i < last; i++) {
        uint32_t px[16];
        int e = orc_unpack_uastc_block_to_rgba(data + 16 * i, px);
        if (e) { if (bad) *bad = i; return e; }
        size_t bx = i % bpr, by = i / bpr;
        for (int y = 0; y < 4; y++) memcpy(out + (4 * by + (size_t)y) * stride + 4 * bx, px + 4 * y, 16);
    }
    return ORC_OK;
}

typedef struct { int fmt; const uint8_t *data; size_t first, last, bpr; uint8_t *out; int err; size_t bad; } Job;
static void *job_main(void *p)
{
    Job *j = (Job *)p;
    j->bad = (size_t)-1;
    if (j->fmt == 0) j->err = rgba_range(j->data, j->first, j->last, j->bpr, (uint32_t *)j->out, &j->bad);
    else j->err = transcode_range(j->fmt, j->data, j->first, j->last, j->out, &j->bad);
    return NULL;
}

/* Slice-level entry (single or multi threaded; static block partition).  fmt 0 needs blocks_per_row.
 * Returns the error of the lowest failing block (reference semantics: first Err aborts). */
ORC_API int orc_uastc_transcode_slice(int fmt, const uint8_t *data, size_t nbytes, size_t blocks_per_row,
                                      uint8_t *out, int threads, size_t *first_bad_block)
{
    if (nbytes % 16 != 0) return ORC_ERR_LEN;      /* uastc.rs:55-56 */
    size_t n = nbytes / 16;
    if (fmt == 0 && blocks_per_row == 0) return n ? ORC_ERR_ARG : ORC_OK;
    if (threads < 1) threads = 1;
    if ((size_t)threads > n) threads = n ? (int)n : 1;
    Job *jobs = (Job *)calloc((size_t)threads, sizeof(Job));
    pthread_t *th = (pthread_t *)calloc((size_t)threads, sizeof(pthread_t));
    for (int t = 0; t < threads; t++) {
        jobs[t].fmt = fmt; jobs[t].data = data; jobs[t].bpr = blocks_per_row; jobs[t].out = out;
        jobs[t].first = n * (size_t)t / (size_t)threads; jobs[t].last = n * (size_t)(t + 1) / (size_t)threads;
        if (threads == 1) job_main(&jobs[t]); else pthread_create(&th[t], NULL, job_main, &jobs[t]);
    }
    int err = ORC_OK; size_t bad = (size_t)-1;
    for (int t = 0; t < threads; t++) {
        if (threads > 1) pthread_join(th[t], NULL);
        if (jobs[t].err && err == ORC_OK) { err = jobs[t].err; bad = jobs[t].bad; }
    }
    free(jobs); free(th);
    if (first_bad_block) *first_bad_block = bad;
    (void)out_block_bytes;
    return err;
}

/* unit-level hooks used by tests (bit reader/writer sweeps mirror bitreader.rs:63-100, bitwriter.rs:118-225) */
ORC_API uint32_t orc_bitreader_read_at(const uint8_t *bytes, size_t len, size_t offset, unsigned count)
{
    BitReader r; br_init(&r, bytes, len); br_remove(&r, (unsigned)offset); return br_read(&r, count);
}
ORC_API void orc_bitwriter_lsb(uint8_t *bytes, size_t len, size_t offset, unsigned count, uint32_t v)
{
    BitWriter w; bw_init(&w, bytes, len); w.bit_pos = offset; bw_write(&w, count, v);
}
ORC_API void orc_bitwriter_msb(uint8_t *bytes, size_t len, size_t offset_from_top, unsigned count, uint32_t v, int rev)
{
    BitWriterMsb w; bwm_init(&w, bytes, len); w.bit_pos -= offset_from_top;
    if (rev) bwm_write_rev(&w, count, v); else bwm_write(&w, count, v);
}
ORC_API uint8_t orc_unquant_endpoint(unsigned trit_quint, unsigned bits, unsigned range)
{
    QuantEndpoint q = {(uint8_t)trit_quint, (uint8_t)bits}; return unquant_endpoint(q, range);
}
/* exposes the literal-f32 p-bit searches so tests can sweep their domains against the kernels' integer forms */
ORC_API void orc_bc7_unique_pbits(unsigned comps, unsigned comp_bits, const uint8_t lo[4], const uint8_t hi[4], uint8_t out_lo[4], uint8_t out_hi[4], uint8_t p[2])
{
    Color32 pair[2]; memcpy(pair[0].c, lo, 4); memcpy(pair[1].c, hi, 4);
    determine_unique_pbits(comps, comp_bits, pair, p);
    memcpy(out_lo, pair[0].c, 4); memcpy(out_hi, pair[1].c, 4);
}
ORC_API void orc_bc7_shared_pbits(unsigned comps, unsigned comp_bits, const uint8_t lo[4], const uint8_t hi[4], uint8_t out_lo[4], uint8_t out_hi[4], uint8_t p[2])
{
    Color32 pair[2]; memcpy(pair[0].c, lo, 4); memcpy(pair[1].c, hi, 4);
    determine_shared_pbits(comps, comp_bits, pair, p);
    memcpy(out_lo, pair[0].c, 4); memcpy(out_hi, pair[1].c, 4);
}

#include "basisu_oracle_etc1s.inc"
#include "etc1s_encoder.inc"
