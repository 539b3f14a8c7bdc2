/*
 * INDEPENDENT WITNESSES (test infrastructure): decoders for the TARGET formats, written from the public format
 * specifications -- Khronos Data Format Specification (ASTC LDR profile: block modes, integer sequence encoding,
 * endpoint / weight unquantisation, partition hash; BPTC/BC7 modes 5 and 6; ETC1; ETC2 EAC alpha) -- and NOT from the
 * reference crate or from this repo's oracle.  They decode what the transcoders produce, so that a transcoder bug and an
 * oracle bug made from the same misreading of the reference cannot cancel: e.g. UASTC -> ASTC is lossless, hence
 * astc_decode(transcode(block)) must equal unpack_to_rgba(block) texel for texel.
 *
 * Nothing here is linked into libb2bu.so or used by bench.py; tests/ only.
 */
#include <stdint.h>
#include <string.h>

#define EXPORT __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------------------------------
 * ASTC, 2D LDR, 4x4 footprint (weight grid == footprint, so no infill), decode to UNORM8 (top 8 bits of the 16-bit result)
 * ------------------------------------------------------------------------------------------------------------------ */
static uint32_t bits128(const uint8_t* b, int pos, int n)          /* LSB-first */
{
    uint32_t v = 0;
    for (int i = 0; i < n; i++) {
        const int p = pos + i;
        if (p >= 0 && p < 128) v |= (uint32_t)((b[p >> 3] >> (p & 7)) & 1) << i;
    }
    return v;
}

typedef struct { int trits, quints, bits; } ise_range;

/* quantisation levels in increasing order: 2,3,4,5,6,8,10,12,16,20,24,32,40,48,64,80,96,128,160,192,256 */
static const ise_range kRanges[21] = {{0, 0, 1}, {1, 0, 0}, {0, 0, 2}, {0, 1, 0}, {1, 0, 1}, {0, 0, 3}, {0, 1, 1}, {1, 0, 2}, {0, 0, 4}, {0, 1, 2}, {1, 0, 3},
                                      {0, 0, 5}, {0, 1, 3}, {1, 0, 4}, {0, 0, 6}, {0, 1, 4}, {1, 0, 5}, {0, 0, 7}, {0, 1, 5}, {1, 0, 6}, {0, 0, 8}};

static int ise_bits(ise_range r, int n)
{
    int total = n * r.bits;
    if (r.trits) total += (8 * n + 4) / 5;
    if (r.quints) total += (7 * n + 2) / 3;
    return total;
}

/* Integer sequence decode: n values from an LSB-first bit source.  get(ctx, pos, n) returns stream bits. */
typedef uint32_t (*bitsrc)(const void* ctx, int pos, int n);

static void ise_decode(bitsrc get, const void* ctx, int start, ise_range r, int n, int* out_bits, int* out_tq)
{
    int pos = start;
    const int total = ise_bits(r, n);
    const int end = start + total;
    if (r.trits) {
        for (int i = 0; i < n; i += 5) {
            /* m0 T[1:0] m1 T[3:2] m2 T[4] m3 T[6:5] m4 T[7]; a truncated last block simply lacks its upper part (reads as 0) */
            static const int tb[5] = {2, 2, 1, 2, 1};
            uint32_t T = 0, m[5] = {0, 0, 0, 0, 0};
            int tpos = 0;
            for (int k = 0; k < 5; k++) {
                if (i + k < n) {
                    int nb = r.bits; if (pos + nb > end) nb = end - pos; if (nb < 0) nb = 0;
                    m[k] = get(ctx, pos, nb); pos += nb;
                    int nt = tb[k]; if (pos + nt > end) nt = end - pos; if (nt < 0) nt = 0;
                    T |= get(ctx, pos, nt) << tpos; pos += nt;
                }
                tpos += tb[k];
            }
            int t[5];
            uint32_t C;
            if (((T >> 2) & 7) == 7) { C = (((T >> 5) & 7) << 2) | (T & 3); t[4] = 2; t[3] = 2; }
            else {
                C = T & 31;
                if (((T >> 5) & 3) == 3) { t[4] = 2; t[3] = (T >> 7) & 1; }
                else { t[4] = (T >> 7) & 1; t[3] = (T >> 5) & 3; }
            }
            if ((C & 3) == 3) { t[2] = 2; t[1] = (C >> 4) & 1; t[0] = (((C >> 3) & 1) << 1) | (((C >> 2) & 1) & ~((C >> 3) & 1)); }
            else if (((C >> 2) & 3) == 3) { t[2] = 2; t[1] = 2; t[0] = C & 3; }
            else { t[2] = (C >> 4) & 1; t[1] = (C >> 2) & 3; t[0] = (((C >> 1) & 1) << 1) | ((C & 1) & ~((C >> 1) & 1)); }
            for (int k = 0; k < 5 && i + k < n; k++) { out_bits[i + k] = (int)m[k]; out_tq[i + k] = t[k]; }
        }
    } else if (r.quints) {
        for (int i = 0; i < n; i += 3) {
            /* m0 Q[2:0] m1 Q[4:3] m2 Q[6:5] */
            static const int qb[3] = {3, 2, 2};
            uint32_t Q = 0, m[3] = {0, 0, 0};
            int qpos = 0;
            for (int k = 0; k < 3; k++) {
                if (i + k < n) {
                    int nb = r.bits; if (pos + nb > end) nb = end - pos; if (nb < 0) nb = 0;
                    m[k] = get(ctx, pos, nb); pos += nb;
                    int nq = qb[k]; if (pos + nq > end) nq = end - pos; if (nq < 0) nq = 0;
                    Q |= get(ctx, pos, nq) << qpos; pos += nq;
                }
                qpos += qb[k];
            }
            int q[3];
            if (((Q >> 1) & 3) == 3 && ((Q >> 5) & 3) == 0) {
                const uint32_t q0b = Q & 1;
                q[2] = (int)((q0b << 2) | ((((Q >> 4) & 1) & ~q0b) << 1) | (((Q >> 3) & 1) & ~q0b));
                q[1] = 4; q[0] = 4;
            } else {
                uint32_t C;
                if (((Q >> 1) & 3) == 3) { q[2] = 4; C = (((Q >> 3) & 3) << 3) | ((~(Q >> 5) & 3) << 1) | (Q & 1); }
                else { q[2] = (Q >> 5) & 3; C = Q & 31; }
                if ((C & 7) == 5) { q[1] = 4; q[0] = (C >> 3) & 3; }
                else { q[1] = (C >> 3) & 3; q[0] = C & 7; }
            }
            for (int k = 0; k < 3 && i + k < n; k++) { out_bits[i + k] = (int)m[k]; out_tq[i + k] = q[k]; }
        }
    } else {
        for (int i = 0; i < n; i++) { out_bits[i] = (int)get(ctx, pos, r.bits); out_tq[i] = 0; pos += r.bits; }
    }
}

static uint32_t src_forward(const void* ctx, int pos, int n) { return bits128((const uint8_t*)ctx, pos, n); }
/* weights are stored from bit 127 downwards: stream bit i is block bit 127 - i */
static uint32_t src_reverse(const void* ctx, int pos, int n)
{
    const uint8_t* b = (const uint8_t*)ctx;
    uint32_t v = 0;
    for (int i = 0; i < n; i++) {
        const int p = 127 - (pos + i);
        if (p >= 0 && p < 128) v |= (uint32_t)((b[p >> 3] >> (p & 7)) & 1) << i;
    }
    return v;
}

static int bit(int v, int i) { return (v >> i) & 1; }

/* colour endpoint unquantisation to 0..255 */
static int unquant_color(ise_range r, int m, int d)
{
    if (!r.trits && !r.quints) {
        /* bit replication */
        int v = 0, have = 0;
        while (have < 8) { v = (v << r.bits) | m; have += r.bits; }
        return (v >> (have - 8)) & 0xFF;
    }
    const int a = m & 1, b = bit(m, 1), c = bit(m, 2), dd = bit(m, 3), e = bit(m, 4), f = bit(m, 5);
    int A = a ? 0x1FF : 0, B = 0, C = 0;
    if (r.trits) {
        switch (r.bits) {
        case 1: B = 0; C = 204; break;
        case 2: B = (b << 8) | (b << 4) | (b << 2) | (b << 1); C = 93; break;
        case 3: B = (c << 8) | (b << 7) | (c << 3) | (b << 2) | (c << 1) | b; C = 44; break;
        case 4: B = (dd << 8) | (c << 7) | (b << 6) | (dd << 2) | (c << 1) | b; C = 22; break;
        case 5: B = (e << 8) | (dd << 7) | (c << 6) | (b << 5) | (e << 1) | dd; C = 11; break;
        default: B = (f << 8) | (e << 7) | (dd << 6) | (c << 5) | (b << 4) | f; C = 5; break;
        }
    } else {
        switch (r.bits) {
        case 1: B = 0; C = 113; break;
        case 2: B = (b << 8) | (b << 3) | (b << 2); C = 54; break;
        case 3: B = (c << 8) | (b << 7) | (c << 2) | (b << 1) | c; C = 26; break;
        case 4: B = (dd << 8) | (c << 7) | (b << 6) | (dd << 1) | c; C = 13; break;
        default: B = (e << 8) | (dd << 7) | (c << 6) | (b << 5) | e; C = 6; break;
        }
    }
    int T = d * C + B;
    T ^= A;
    return (A & 0x80) | (T >> 2);
}

/* weight unquantisation to 0..64 */
static int unquant_weight(ise_range r, int m, int d)
{
    int w;
    if (!r.trits && !r.quints) {
        int v = 0, have = 0;
        while (have < 6) { v = (v << r.bits) | m; have += r.bits; }
        w = (v >> (have - 6)) & 63;
    } else if (r.bits == 0) {
        static const int t3[3] = {0, 32, 63}, q5[5] = {0, 16, 32, 47, 63};
        w = r.trits ? t3[d] : q5[d];
    } else {
        const int a = m & 1, b = bit(m, 1), c = bit(m, 2);
        int A = a ? 0x7F : 0, B = 0, C = 0;
        if (r.trits) {
            if (r.bits == 1) { B = 0; C = 50; }
            else if (r.bits == 2) { B = (b << 6) | (b << 2) | b; C = 23; }
            else { B = (c << 6) | (b << 5) | (c << 1) | b; C = 11; }
        } else {
            if (r.bits == 1) { B = 0; C = 28; }
            else { B = (b << 6) | (b << 1); C = 13; }
        }
        int T = d * C + B;
        T ^= A;
        w = (A & 0x20) | (T >> 2);
    }
    if (w > 32) w += 1;
    return w;
}

static uint32_t hash52(uint32_t p)
{
    p ^= p >> 15; p -= p << 17; p += p << 7; p += p << 4; p ^= p >> 5; p += p << 16; p ^= p >> 7; p ^= p >> 3; p ^= p << 6; p ^= p >> 17;
    return p;
}

static int select_partition(int seed, int x, int y, int z, int partitioncount, int small_block)
{
    if (small_block) { x <<= 1; y <<= 1; z <<= 1; }
    seed += (partitioncount - 1) * 1024;
    const uint32_t rnum = hash52((uint32_t)seed);
    uint8_t s[12];
    s[0] = rnum & 0xF; s[1] = (rnum >> 4) & 0xF; s[2] = (rnum >> 8) & 0xF; s[3] = (rnum >> 12) & 0xF; s[4] = (rnum >> 16) & 0xF; s[5] = (rnum >> 20) & 0xF;
    s[6] = (rnum >> 24) & 0xF; s[7] = (rnum >> 28) & 0xF; s[8] = (rnum >> 18) & 0xF; s[9] = (rnum >> 22) & 0xF; s[10] = (rnum >> 26) & 0xF;
    s[11] = ((rnum >> 30) | (rnum << 2)) & 0xF;
    for (int i = 0; i < 12; i++) s[i] = (uint8_t)(s[i] * s[i]);
    int sh1, sh2, sh3;
    if (seed & 1) { sh1 = (seed & 2) ? 4 : 5; sh2 = (partitioncount == 3) ? 6 : 5; }
    else { sh1 = (partitioncount == 3) ? 6 : 5; sh2 = (seed & 2) ? 4 : 5; }
    sh3 = (seed & 0x10) ? sh1 : sh2;
    s[0] >>= sh1; s[1] >>= sh2; s[2] >>= sh1; s[3] >>= sh2; s[4] >>= sh1; s[5] >>= sh2; s[6] >>= sh1; s[7] >>= sh2;
    s[8] >>= sh3; s[9] >>= sh3; s[10] >>= sh3; s[11] >>= sh3;
    int a = s[0] * x + s[1] * y + s[10] * z + (int)(rnum >> 14);
    int b = s[2] * x + s[3] * y + s[11] * z + (int)(rnum >> 10);
    int c = s[4] * x + s[5] * y + s[8] * z + (int)(rnum >> 6);
    int d = s[6] * x + s[7] * y + s[9] * z + (int)(rnum >> 2);
    a &= 0x3F; b &= 0x3F; c &= 0x3F; d &= 0x3F;
    if (partitioncount < 4) d = 0;
    if (partitioncount < 3) c = 0;
    if (a >= b && a >= c && a >= d) return 0;
    if (b >= c && b >= d) return 1;
    if (c >= d) return 2;
    return 3;
}

static void blue_contract(int* r, int* g, int* b) { *r = (*r + *b) >> 1; *g = (*g + *b) >> 1; }

/* Decodes one 4x4 LDR block to 16 RGBA8 texels (raster order).  Returns 0, or a negative code for anything outside the
 * subset of ASTC that a UASTC transcode can produce (other footprints, HDR endpoint modes, multiple CEM classes ...). */
EXPORT int spec_astc_decode_4x4(const uint8_t blk[16], uint8_t out[64])
{
    const int mode = (int)bits128(blk, 0, 11);
    if ((mode & 0x1FF) == 0x1FC) {
        /* void extent: bit 9 = HDR flag, bits 10-11 reserved (ones), 4 x 13-bit extents, then R,G,B,A as UNORM16 */
        if (mode & 0x200) return -2;
        for (int t = 0; t < 16; t++)
            for (int c = 0; c < 4; c++) out[4 * t + c] = (uint8_t)(bits128(blk, 64 + 16 * c, 16) >> 8);
        return 0;
    }
    if ((mode & 0xF) == 0) return -3;                       /* reserved */
    int R, W, H;
    if ((mode & 3) != 0) {
        R = (bit(mode, 4)) | (bit(mode, 0) << 1) | (bit(mode, 1) << 2);
        const int A = (mode >> 5) & 3, B = (mode >> 7) & 3;
        switch ((mode >> 2) & 3) {
        case 0: W = B + 4; H = A + 2; break;
        case 1: W = B + 8; H = A + 2; break;
        case 2: W = A + 2; H = B + 8; break;
        default: if (bit(mode, 8)) { W = (B & 1) + 2; H = A + 2; } else { W = A + 2; H = (B & 1) + 6; } break;
        }
    } else return -4;                                       /* the 12-wide / 6x10 layouts never describe a 4x4 grid */
    if (W != 4 || H != 4) return -5;
    const int dual = bit(mode, 10), hp = bit(mode, 9);
    if (R < 2) return -6;
    static const int wr_lo[8] = {-1, -1, 0, 1, 2, 3, 4, 5}, wr_hi[8] = {-1, -1, 6, 7, 8, 9, 10, 11};
    const ise_range wr = kRanges[hp ? wr_hi[R] : wr_lo[R]];
    const int nweights = 16 * (dual ? 2 : 1);
    const int wbits = ise_bits(wr, nweights);
    if (nweights > 64 || wbits < 24 || wbits > 96) return -7;

    const int nparts = (int)bits128(blk, 11, 2) + 1;
    if (dual && nparts == 4) return -8;
    int cem, seed = 0, cpos;
    if (nparts == 1) { cem = (int)bits128(blk, 13, 4); cpos = 17; }
    else {
        seed = (int)bits128(blk, 13, 10);
        const int cemfield = (int)bits128(blk, 23, 6);
        if ((cemfield & 3) != 0) return -9;                 /* partitions with different endpoint modes */
        cem = cemfield >> 2;
        cpos = 29;
    }
    if (cem != 4 && cem != 8 && cem != 12) return -10;      /* LA direct, RGB direct, RGBA direct */
    const int nvals = nparts * (2 * (cem / 4 + 1));
    const int avail = 128 - wbits - cpos - (dual ? 2 : 0);
    int ri = -1;
    for (int i = 20; i >= 0; i--) if (ise_bits(kRanges[i], nvals) <= avail) { ri = i; break; }
    if (ri < 4) return -11;                                 /* fewer than 6 levels is illegal for colours */
    const ise_range cr = kRanges[ri];
    int cb[18], ct[18], ep[18];
    ise_decode(src_forward, blk, cpos, cr, nvals, cb, ct);
    for (int i = 0; i < nvals; i++) ep[i] = unquant_color(cr, cb[i], ct[i]);

    int wb[32], wt[32], wq[32];
    ise_decode(src_reverse, blk, 0, wr, nweights, wb, wt);
    for (int i = 0; i < nweights; i++) wq[i] = unquant_weight(wr, wb[i], wt[i]);
    const int ccs = dual ? (int)bits128(blk, 128 - wbits - 2, 2) : -1;

    int e0[4][4], e1[4][4];                                 /* per partition RGBA */
    const int per = nvals / nparts;
    for (int p = 0; p < nparts; p++) {
        const int* v = ep + p * per;
        if (cem == 4) {
            for (int c = 0; c < 3; c++) { e0[p][c] = v[0]; e1[p][c] = v[1]; }
            e0[p][3] = v[2]; e1[p][3] = v[3];
        } else {
            const int s0 = v[0] + v[2] + v[4], s1 = v[1] + v[3] + v[5];
            int a0 = 255, a1 = 255;
            if (cem == 12) { a0 = v[6]; a1 = v[7]; }
            if (s1 >= s0) {
                e0[p][0] = v[0]; e0[p][1] = v[2]; e0[p][2] = v[4]; e0[p][3] = a0;
                e1[p][0] = v[1]; e1[p][1] = v[3]; e1[p][2] = v[5]; e1[p][3] = a1;
            } else {
                e0[p][0] = v[1]; e0[p][1] = v[3]; e0[p][2] = v[5]; e0[p][3] = a1;
                e1[p][0] = v[0]; e1[p][1] = v[2]; e1[p][2] = v[4]; e1[p][3] = a0;
                blue_contract(&e0[p][0], &e0[p][1], &e0[p][2]);
                blue_contract(&e1[p][0], &e1[p][1], &e1[p][2]);
            }
        }
    }
    for (int y = 0; y < 4; y++)
        for (int x = 0; x < 4; x++) {
            const int t = y * 4 + x;
            const int p = nparts == 1 ? 0 : select_partition(seed, x, y, 0, nparts, 1);
            for (int c = 0; c < 4; c++) {
                const int w = dual ? wq[2 * t + (c == ccs ? 1 : 0)] : wq[t];
                const int c0 = (e0[p][c] << 8) | e0[p][c], c1 = (e1[p][c] << 8) | e1[p][c];
                const int v = (c0 * (64 - w) + c1 * w + 32) >> 6;
                out[4 * t + c] = (uint8_t)(v >> 8);
            }
        }
    return 0;
}

/* the partition (subset) index of every texel of a 4x4 block for a 10-bit seed -- exposed so that a test can compare the
 * hash with the pattern tables */
EXPORT void spec_astc_partition_map(int seed, int nparts, uint8_t out[16])
{
    for (int y = 0; y < 4; y++)
        for (int x = 0; x < 4; x++) out[y * 4 + x] = (uint8_t)select_partition(seed, x, y, 0, nparts, 1);
}

/* ------------------------------------------------------------------------------------------------------------------
 * ETC1 (OpenGL ES: OES_compressed_ETC1_RGB8_texture; KDFS "ETC1").  ETC2's T / H / planar modes are reported, not decoded.
 * ------------------------------------------------------------------------------------------------------------------ */
static int clamp255(int v) { return v < 0 ? 0 : v > 255 ? 255 : v; }

EXPORT int spec_etc1_decode(const uint8_t b[8], uint8_t out[64])
{
    static const int mods[8][2] = {{2, 8}, {5, 17}, {9, 29}, {13, 42}, {18, 60}, {24, 80}, {33, 106}, {47, 183}};
    const int diff = (b[3] >> 1) & 1, flip = b[3] & 1;
    const int tab[2] = {(b[3] >> 5) & 7, (b[3] >> 2) & 7};
    int base[2][3];
    for (int c = 0; c < 3; c++) {
        if (!diff) {
            const int c1 = b[c] >> 4, c2 = b[c] & 15;
            base[0][c] = c1 * 17; base[1][c] = c2 * 17;
        } else {
            const int c1 = b[c] >> 3;
            int d = b[c] & 7; if (d >= 4) d -= 8;
            const int c2 = c1 + d;
            if (c2 < 0 || c2 > 31) return -1;               /* ETC2 mode, not ETC1 */
            base[0][c] = (c1 << 3) | (c1 >> 2); base[1][c] = (c2 << 3) | (c2 >> 2);
        }
    }
    const uint32_t msb = ((uint32_t)b[4] << 8) | b[5], lsb = ((uint32_t)b[6] << 8) | b[7];
    for (int x = 0; x < 4; x++)
        for (int y = 0; y < 4; y++) {
            const int i = x * 4 + y;                        /* pixel index: column-major */
            const int sb = flip ? (y >= 2) : (x >= 2);
            const int hi = (msb >> i) & 1, lo = (lsb >> i) & 1;
            const int mag = mods[tab[sb]][lo];
            const int m = hi ? -mag : mag;
            uint8_t* px = out + 4 * (y * 4 + x);
            for (int c = 0; c < 3; c++) px[c] = (uint8_t)clamp255(base[sb][c] + m);
            px[3] = 255;
        }
    return 0;
}

/* ETC2 EAC alpha block (the first 8 bytes of an ETC2 RGBA8 block) -> 16 alpha values in raster order */
EXPORT void spec_eac_alpha_decode(const uint8_t b[8], uint8_t out[16])
{
    static const int tbl[16][8] = {{-3, -6, -9, -15, 2, 5, 8, 14}, {-3, -7, -10, -13, 2, 6, 9, 12}, {-2, -5, -8, -13, 1, 4, 7, 12}, {-2, -4, -6, -13, 1, 3, 5, 12},
                                   {-3, -6, -8, -12, 2, 5, 7, 11}, {-3, -7, -9, -11, 2, 6, 8, 10}, {-4, -7, -8, -11, 3, 6, 7, 10}, {-3, -5, -8, -11, 2, 4, 7, 10},
                                   {-2, -6, -8, -10, 1, 5, 7, 9}, {-2, -5, -8, -10, 1, 4, 7, 9}, {-2, -4, -8, -10, 1, 3, 7, 9}, {-2, -5, -7, -10, 1, 4, 6, 9},
                                   {-3, -4, -7, -10, 2, 3, 6, 9}, {-1, -2, -3, -10, 0, 1, 2, 9}, {-4, -6, -8, -9, 3, 5, 7, 8}, {-3, -5, -7, -9, 2, 4, 6, 8}};
    const int base = b[0], mult = b[1] >> 4, ti = b[1] & 15;
    uint64_t sel = 0;
    for (int i = 2; i < 8; i++) sel = (sel << 8) | b[i];
    for (int i = 0; i < 16; i++) {
        const int idx = (int)((sel >> (45 - 3 * i)) & 7);   /* pixel i: x = i / 4, y = i % 4 */
        const int x = i / 4, y = i % 4;
        out[y * 4 + x] = (uint8_t)clamp255(base + tbl[ti][idx] * mult);
    }
}

/* ------------------------------------------------------------------------------------------------------------------
 * BC7 (BPTC), the single-subset modes a UASTC transcode produces: mode 5 (7.7.7 + 8 bit alpha, two index sets) and mode 6
 * (7.7.7.7 + p-bit, 4-bit indices).  Partitioned modes return -1 (their partition tables are spec constants this file
 * does not restate).
 * ------------------------------------------------------------------------------------------------------------------ */
EXPORT int spec_bc7_decode_mode56(const uint8_t blk[16], uint8_t out[64])
{
    static const int w2[4] = {0, 21, 43, 64};
    static const int w4[16] = {0, 4, 9, 13, 17, 21, 26, 30, 34, 38, 43, 47, 51, 55, 60, 64};
    int mode = 0;
    while (mode < 8 && !bit(blk[0], mode)) mode++;
    int pos = mode + 1;
    if (mode == 6) {
        int e[2][4];
        for (int c = 0; c < 4; c++) for (int k = 0; k < 2; k++) { e[k][c] = (int)bits128(blk, pos, 7); pos += 7; }
        for (int k = 0; k < 2; k++) { const int p = (int)bits128(blk, pos, 1); pos += 1; for (int c = 0; c < 4; c++) e[k][c] = (e[k][c] << 1) | p; }
        for (int t = 0; t < 16; t++) {
            const int nb = t == 0 ? 3 : 4;
            const int w = w4[bits128(blk, pos, nb)]; pos += nb;
            for (int c = 0; c < 4; c++) out[4 * t + c] = (uint8_t)(((64 - w) * e[0][c] + w * e[1][c] + 32) >> 6);
        }
        return 0;
    }
    if (mode == 5) {
        const int rot = (int)bits128(blk, pos, 2); pos += 2;
        int e[2][4];
        for (int c = 0; c < 3; c++) for (int k = 0; k < 2; k++) { const int v = (int)bits128(blk, pos, 7); pos += 7; e[k][c] = (v << 1) | (v >> 6); }
        for (int k = 0; k < 2; k++) { e[k][3] = (int)bits128(blk, pos, 8); pos += 8; }
        int ci[16], ai[16];
        for (int t = 0; t < 16; t++) { const int nb = t == 0 ? 1 : 2; ci[t] = (int)bits128(blk, pos, nb); pos += nb; }
        for (int t = 0; t < 16; t++) { const int nb = t == 0 ? 1 : 2; ai[t] = (int)bits128(blk, pos, nb); pos += nb; }
        for (int t = 0; t < 16; t++) {
            int px[4];
            for (int c = 0; c < 3; c++) px[c] = ((64 - w2[ci[t]]) * e[0][c] + w2[ci[t]] * e[1][c] + 32) >> 6;
            px[3] = ((64 - w2[ai[t]]) * e[0][3] + w2[ai[t]] * e[1][3] + 32) >> 6;
            if (rot) { const int tmp = px[3]; px[3] = px[rot - 1]; px[rot - 1] = tmp; }
            for (int c = 0; c < 4; c++) out[4 * t + c] = (uint8_t)px[c];
        }
        return 0;
    }
    return -1;
}

EXPORT int spec_bc7_mode(const uint8_t blk[16])
{
    int mode = 0;
    while (mode < 8 && !bit(blk[0], mode)) mode++;
    return mode;
}
