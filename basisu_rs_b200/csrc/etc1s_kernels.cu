// ETC1S / BasisLZ device side (K2 entropy decode, K3 codebook gather).
//
// K2  etc1s_entropy_decode: the slice bitstream is one serial chain (reference src/basis_lz/mod.rs:188-458),
//     but only the BIT POSITION is inherently serial: which table is read next depends on the predictor bits
//     and the two run counters, never on the endpoint values or on the selector history.  So every slice is
//     decoded by a pipeline of warps connected by a token ring in shared memory:
//
//       tokenizer warp   walks the bitstream and emits one token per block: predictor (2 bits), the endpoint
//                        delta symbol, the raw selector symbol (codebook index, history reference or run
//                        repeat).  Two forms.  Narrow (more than one slice per SM, video): the chain per symbol
//                        is  AND -> LDS (first-level table in shared memory) -> funnel shift, all lanes redundant
//                        on identical state.  Wide (at most one slice per SM): helper warps decode EVERY bit
//                        position against the hot models ahead of the tokenizer (one lane per position) and leave
//                        code size and symbol per position in a shared ring, delta + selector symbol combined,
//                        so that a block is one  LDS.U8 -> IADD  link of the chain (see the speculation ring below).
//       resolver warp    turns 32 tokens at a time into (endpoint, selector) pairs: the endpoint predictors
//                        (left / up / up-left / delta) are a segmented scan over the warp (composition of
//                        "set to c" and "add d mod n"); the approximate-move-to-front selector history is the
//                        one serial loop and runs over the round's history references only (inserts are a
//                        scatter whose slots follow from a prefix count); then one coalesced store of the 32
//                        index words and the row state for the next row.
//
//     Errors keep the reference's order: every check has a key (block, phase) and the slice status is the code
//     of the smallest key over both warps.
// K3  etc1s_gather_etc1 / etc1s_gather_rgba: one thread per block, pure gather
//     (mod.rs:122-146, :163-181).
#include "etc1s_device.h"
#include "ptx_helpers.cuh"

namespace b2bu {

// etc.rs:436-445 ETC1 intensity modifier table
__device__ const int16_t kEtc1Mod[32] = {-8, -2, 2, 8, -17, -5, 5, 17, -29, -9, 9, 29, -42, -13, 13, 42,
                                        -60, -18, 18, 60, -80, -24, 24, 80, -106, -33, 33, 106, -183, -47, 47, 183};

constexpr int kRound = 32;                        // blocks per token round
constexpr int kTokRounds = 4;                     // depth of the token ring
constexpr int kHalfBytes = 1024;                  // the compressed-byte ring is refilled in 1 KiB pieces
constexpr int kRingHalves = 4;                    // 4 KiB ring per slice: the piece being read, the next, one in flight
constexpr int kRingWords = kRingHalves * kHalfBytes / 4;
// A round of 32 blocks consumes well under one piece: per block at most one predictor symbol (16 bits) with
// a 4-bit-chunk VLC (40), a delta (16), a selector symbol (16), a run symbol (16) and a 7-bit-chunk VLC (40).
static_assert(kRound * 144 / 8 <= kHalfBytes, "ring piece too small for one round");

// token word A: selector symbol | predictor << 16 | flags; word B: endpoint delta symbol
constexpr uint32_t kTokSkipSel = 1u << 18;        // texture video, predictor 2: no selector symbol (mod.rs:366)
constexpr uint32_t kTokErr = 1u << 31;            // tokenizer stopped here: code in bits 24-27, phase in bits 28-30
// order of the checks inside one block (mod.rs:245-455): predictor symbol (1), predictor asserts (2), delta symbol (3),
// selector symbols (4), history asserts (5), final range asserts (6)
enum { PH_PRED_SYM = 1, PH_PRED_CHECK = 2, PH_DELTA = 3, PH_SELECTOR = 4, PH_HISTORY = 5, PH_RANGE = 6 };

struct alignas(16) PipeShared {
    // compressed bytes of the slice (tokenizer): four 1 KiB pieces, and the piece in slot 0 a second time behind them, so
    // that the reader's address only has to wrap once per round (a round consumes less than one piece)
    uint32_t ring[kRingWords + kHalfBytes / 4];
    uint2 tok[kTokRounds][kRound];                // tokenizer -> resolver
    uint16_t hist[64];                            // selector history when it fits (it always does for real files)
    uint2 hdesc[kRound];                          // resolver: shared addresses of the entries of the round's history references
    uint16_t hres[kRound];                        //           ... and what they read
    uint64_t bar_full[kTokRounds], bar_empty[kTokRounds];
    uint32_t abort;                               // set by the resolver when it has found an error
    uint32_t taken;                               // rounds the resolver has taken out of the token ring (wide pipelines read this instead of bar_empty)
};

// ---- speculation ring ("wide" pipelines: at most two slices per SM, so there are warps to spare) ----
// The only thing the tokenizer's chain really needs from a table read is the CODE LENGTH at the current bit position.
// Helper warps decode every bit position of the stream against the three hot models ahead of the tokenizer -- one lane per
// position, no dependence between positions -- and leave (length, symbol) per position and model in shared memory.  The
// tokenizer's link becomes  LDS.U8 [q + table] -> IADD  (measured: 28-34 cycles against 44-62 for LDS -> SHF -> LOP3 ->
// IMAD, tools/probe_chase.cu), with the symbol fetched by a second load that nothing waits for.
constexpr int kSpecPos = 4096;                    // bit positions in the ring
constexpr int kSpecChunk = 256;                   // positions per helper work item
constexpr int kSpecChunks = kSpecPos / kSpecChunk;
constexpr int kSpecMirrorChunks = 5;              // the ring's first chunks are repeated behind it: the reader wraps at sync points only
constexpr int kSpecMirror = kSpecMirrorChunks * kSpecChunk;
constexpr int kSpecAhead = 6;                     // chunks that must be complete from the reader's chunk on at a sync point
// between two sync points (one per round, and after every slow pair) the fast path reads at most 16 pairs = 16 * (16 + 2 * 32)
// = 1280 bits on from where it is
static_assert(255 + 1280 < kSpecAhead * kSpecChunk && 1280 <= kSpecMirror && kSpecAhead < kSpecChunks, "speculation ring sync distance");
// entry: symbol in the low half, flags in byte 2, code size * 4 in byte 3 (0 for anything the fast path does not take)
constexpr uint32_t kSpecSpecial = 1u << 16;       // no code, run symbol, symbol above 16 bits
constexpr uint32_t kSpecDelta0 = 1u << 22;        // predictor symbols only: the first / second block of the pair has predictor 3
constexpr uint32_t kSpecDelta1 = 1u << 23;
constexpr uint32_t kSpecTableBytes = (kSpecPos + kSpecMirror) * 4u;   // distance between the three tables of a ring
// The four tables of a ring, per bit position p:
//   0  endpoint_pred symbol at p            2  selector symbol at p
//   1  delta_endpoint symbol at p FOLLOWED BY the selector symbol behind it: delta symbol | flags | both code sizes (x 4)
//   3  that selector symbol | the delta's own code size (x 4) in byte 3 (for blocks inside a selector run)
// so a block costs the tokenizer ONE link whether it has a delta symbol or not.
struct alignas(16) SpecShared {
    uint32_t ent[4][kSpecPos + kSpecMirror];
    uint32_t stage[kEtc1sHelpers > 0 ? kEtc1sHelpers : 1][kSpecChunk + 32];   // a helper's selector entries of its chunk and the 32 positions behind it
    uint32_t done[kSpecChunks];                   // chunk c complete: done[c % kSpecChunks] == c + 1
    uint32_t consumed;                            // the reader is at or beyond this chunk
    uint32_t raw_loaded;                          // bytes of the slice that have reached the compressed-byte ring
    uint32_t stop, pad;
};

// 64 buffered stream bits (hi:lo, `avail` valid, zeros above), the next ring word already loaded, and `pre`: the low
// word as it was before the last refill.  Reads take at most 16 bits and every read is followed by a refill to >= 32
// bits, so `pre` always holds >= 16 valid bits and the next table index can be formed from it without waiting for the
// refill (which then sits beside the dependent chain, not on it).
// x: bits still buffered in its low 8 bits (<= 64), garbage above; raddr: shared address of the ring word after nw
struct BitState { uint32_t lo, hi, pre, x, nw, raddr; };

// Drops `e & 31` (<= 16) bits and refills; bits 5-7 of e must be zero (callers strip flags / symbol bits that sit there).
// Written as predicated PTX so that the refill is straight-line code (the compiler's version of the same C++ is a branch
// with a convergence barrier around it on every symbol) and kept to 11 instructions: the bit count is updated by
// subtracting the whole table entry (the symbol in bits 8+ only disturbs bits the count does not use), the shifts take
// their amounts modulo 32 straight from that register, and the ring address is a plain running pointer.
__device__ __forceinline__ void bits_consume(BitState& s, uint32_t e)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        ".reg .u32 sh, t;\n"
        "shf.r.wrap.b32 %0, %0, %1, %6;\n"                   // lo = (hi:lo) >> len
        "shf.r.wrap.b32 %1, %1, 0, %6;\n"                    // hi >>= len
        "sub.u32 %3, %3, %6;\n"                              // count -= len
        "mov.b32 %2, %0;\n"                                  // pre = lo
        "and.b32 t, %3, 0xE0;\n"
        "setp.eq.u32 p, t, 0;\n"                             // count < 32 (it is >= 16 here)
        "shf.l.wrap.b32 sh, 0, %4, %3;\n"                    // nw << count
        "@p or.b32 %0, %0, sh;\n"
        "@p shf.l.wrap.b32 %1, %4, 0, %3;\n"                 // hi = nw >> (32 - count)
        "@p add.u32 %3, %3, 32;\n"
        "@p ld.shared.u32 %4, [%5];\n"
        "@p add.u32 %5, %5, 4;\n"
        "}\n"
        : "+r"(s.lo), "+r"(s.hi), "=r"(s.pre), "+r"(s.x), "+r"(s.nw), "+r"(s.raddr)
        : "r"(e)
        : "memory");
}

// Out-of-line slow paths (by value: the bit state must stay in registers on the fast path).
struct SlowSym { BitState bs; uint32_t sym; };
// huffman.rs:186-198 for a first-level entry flagged `special`: a run symbol (already consumed) or a code longer than the
// first-level table, which is looked up in the reference's flat table in global memory.  sym = 0xFFFFFFFF: no code matches.
__device__ __noinline__ SlowSym huff_slow(BitState bs, uint32_t e, const uint32_t* __restrict__ flat, uint32_t max_len)
{
    SlowSym r;
    if ((e & 31u) != 0u) { r.bs = bs; r.sym = e >> 8; return r; }
    const uint32_t f = __ldg(flat + (bs.lo & ((1u << max_len) - 1u)));
    if ((f & 31u) == 0u) { r.bs = bs; r.sym = 0xFFFFFFFFu; return r; }
    bits_consume(bs, f & 31u);
    r.bs = bs; r.sym = f >> 5;
    return r;
}

// mod.rs:585-608 decode_vlc.  sym = the value, or 0xFFFFFFFF when the reference would panic (ofs >= 32).
__device__ __noinline__ SlowSym vlc_decode(BitState bs, uint32_t chunk_bits)
{
    SlowSym r;
    uint32_t v = 0, ofs = 0;
    for (;;) {
        const uint32_t c = bs.lo & ((2u << chunk_bits) - 1u);
        bits_consume(bs, chunk_bits + 1u);
        v |= (c & ((1u << chunk_bits) - 1u)) << ofs;
        ofs += chunk_bits;
        if ((c >> chunk_bits) == 0u) break;
        if (ofs >= 32u) { r.bs = bs; r.sym = 0xFFFFFFFFu; return r; }
    }
    r.bs = bs; r.sym = v;
    return r;
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
    if (piece % kRingHalves == 0u) {                          // slot 0 is mirrored behind the ring
        uint4* mir = reinterpret_cast<uint4*>(ring + kRingWords);
        mir[lane] = r[0];
        mir[lane + 32] = r[1];
    }
}

#ifdef B2BU_K2_TRACE
// tuning aid (never in the product build): per-slice cycle / event counters of the two stages
__device__ unsigned long long g_k2trace[64][16];
__device__ unsigned long long g_k2trace2[64][2];
#define K2T(slot, v) do { if (lane == 0 && trace_slice < 64u) g_k2trace[trace_slice][(slot)] += (unsigned long long)(v); } while (0)
#define K2T_DECL(...) __VA_ARGS__
#else
#define K2T(slot, v) do { } while (0)
#define K2T_DECL(...)
#endif

__device__ __forceinline__ uint32_t ld_volatile_shared(const uint32_t* p) { return *reinterpret_cast<const volatile uint32_t*>(p); }

// ---------------------------------------------------------------------------------------------------
// Stage 1: tokenizer warp.  predrow: one 64-bit word per 32 blocks of the row (2 predictor bits per block), written on
// even rows for the odd row below (mod.rs:286-296).
//
// A lone warp pays the full latency of everything it issues: a shared-memory load and a taken branch cost about the
// same (~30 cycles), an ALU instruction ~4.  So the hot loop (fast_pair) is straight-line predicated code, two blocks
// per iteration, with exactly one rarely-taken branch: the table reads a block does not need (no delta symbol, inside
// a run) are predicated off instead of branched around, and everything unusual (a code longer than the first-level
// table, a run symbol, an invalid code) only ORs a flag into `spec`; a flagged pair is thrown away and decoded again
// from the saved state by pair_slow, which follows the reference's control flow literally.
// ---------------------------------------------------------------------------------------------------
struct TokState { BitState bs; uint32_t sel_rle, pred_rep, prev_sym, cur, terr; };   // terr: code | phase << 4 when the stream is bad

struct TokConsts {
    uint32_t t0, t1, t2, t3;          // shared byte addresses of the first-level tables
    uint32_t m0, m1, m2, m3;          // index masks
    uint32_t ring_base, num_selectors, rle_sym, is_video;
};

__device__ __forceinline__ uint32_t lds_u32(uint32_t addr)
{
    uint32_t e;
    asm("ld.shared.u32 %0, [%1];" : "=r"(e) : "r"(addr));
    return e;
}
// table read that is skipped (reads as 0: consumes nothing) when `on` is zero
__device__ __forceinline__ uint32_t lds_u32_if(uint32_t addr, uint32_t on)
{
    uint32_t e;
    asm("{\n.reg .pred p;\nsetp.ne.u32 p, %2, 0;\nmov.u32 %0, 0;\n@p ld.shared.u32 %0, [%1];\n}\n" : "=r"(e) : "r"(addr), "r"(on));
    return e;
}

// The reference's control flow for one pair of blocks (or the single last block of an odd-width row), mod.rs:259-426.
// even_row: decode the 2x2 group's predictor symbol (else st.cur already holds the pair's 4 predictor bits).
__device__ __noinline__ TokState pair_slow(TokState st, const TokConsts K, const Etc1sDecodeParams& P, uint2* tk, uint32_t nblk, uint32_t even_row)
{
    BitState bs = st.bs;
    uint32_t terr = 0, cur = st.cur;
    if (even_row) {
        if (st.pred_rep != 0u) { st.pred_rep--; cur = st.prev_sym; }
        else {
            const uint32_t e = lds_u32(K.t0 + ((bs.pre & K.m0) << 2));
            bits_consume(bs, e & ~kL1Special);
            cur = e >> 8;
            if (e & kL1Special) {
                SlowSym r = huff_slow(bs, e, P.flat[0], P.max_len[0]);
                bs = r.bs; cur = r.sym;
                if (cur == 0xFFFFFFFFu) terr = ETC1S_ERR_HUFFMAN | (PH_PRED_SYM << 4);
                else if (cur == 256u) {                                           // mod.rs:268-275: run of the previous symbol
                    r = vlc_decode(bs, 4);
                    bs = r.bs;
                    if (r.sym == 0xFFFFFFFFu) terr = ETC1S_ERR_VLC | (PH_PRED_SYM << 4);
                    st.pred_rep = r.sym + 3u - 1u;
                    cur = st.prev_sym;
                }
            }
            if (terr) tk[0] = make_uint2(kTokErr | ((terr & 15u) << 24) | ((terr >> 4) << 28), 0u);
            cur &= 0xFFu;
            st.prev_sym = cur;
        }
    }
    for (uint32_t j = 0; j < nblk && !terr; j++) {
        const uint32_t pred = (cur >> (2u * j)) & 3u;
        uint32_t d = 0u, tokA = pred << 16;
        if (pred == 3u) {                                                         // mod.rs:340-353: DPCM delta symbol
            const uint32_t e = lds_u32(K.t1 + ((bs.pre & K.m1) << 2));
            bits_consume(bs, e & ~kL1Special);
            d = e >> 8;
            if (e & kL1Special) {
                const SlowSym r = huff_slow(bs, e, P.flat[1], P.max_len[1]);
                bs = r.bs; d = r.sym;
                if (d == 0xFFFFFFFFu) terr = ETC1S_ERR_HUFFMAN | (PH_DELTA << 4);
            }
        }
        if (K.is_video && pred == 2u) tokA |= kTokSkipSel;
        else if (terr) { }
        else if (st.sel_rle > 0u) { st.sel_rle--; tokA |= K.num_selectors; }      // mod.rs:370-372
        else {
            const uint32_t e = lds_u32(K.t2 + ((bs.pre & K.m2) << 2));
            bits_consume(bs, e & ~kL1Special);
            uint32_t sym = e >> 8;
            if (e & kL1Special) {
                SlowSym r = huff_slow(bs, e, P.flat[2], P.max_len[2]);
                bs = r.bs; sym = r.sym;
                if (sym == 0xFFFFFFFFu) terr = ETC1S_ERR_HUFFMAN | (PH_SELECTOR << 4);
                else if (sym == K.rle_sym) {                                      // mod.rs:378-396
                    const uint32_t e3 = lds_u32(K.t3 + ((bs.pre & K.m3) << 2));
                    bits_consume(bs, e3 & ~kL1Special);
                    uint32_t cnt = e3 >> 8;
                    if (e3 & kL1Special) { r = huff_slow(bs, e3, P.flat[3], P.max_len[3]); bs = r.bs; cnt = r.sym; }
                    if (cnt == 0xFFFFFFFFu) terr = ETC1S_ERR_HUFFMAN | (PH_SELECTOR << 4);
                    else {
                        if (cnt == 63u) {
                            r = vlc_decode(bs, 7);
                            bs = r.bs; cnt = r.sym;
                            if (cnt == 0xFFFFFFFFu) terr = ETC1S_ERR_VLC | (PH_SELECTOR << 4);
                        }
                        st.sel_rle = 3u + cnt - 1u;
                        sym = K.num_selectors;
                    }
                }
            }
            tokA |= sym & 0xFFFFu;
        }
        if (terr) tokA = (tokA & 0x30000u) | kTokErr | ((terr & 15u) << 24) | ((terr >> 4) << 28);
        tk[j] = make_uint2(tokA, d);
    }
    st.bs = bs; st.cur = cur; st.terr = terr;
    return st;
}

// Fast path for NP consecutive pairs of blocks of a non-video slice; returns false (state untouched, no tokens written)
// when one of them needs pair_slow.  EVEN: the pairs start with their 2x2 group's predictor symbol and `grp` receives
// the NP symbols (8 bits each, first pair lowest); else `grp` supplies 4 predictor bits per pair.
template <bool EVEN, int NP>
__device__ __forceinline__ bool fast_pairs(TokState& st, const TokConsts& K, uint2* tk, uint32_t& grp)
{
    BitState bs = st.bs;
    uint32_t spec = 0u, sel_rle = st.sel_rle, pred_rep = st.pred_rep, prev_sym = st.prev_sym, syms = 0u;
    uint2 t[2 * NP];
#pragma unroll
    for (int q = 0; q < NP; q++) {
        uint32_t cur;
        if (EVEN) {
            const uint32_t rep = pred_rep != 0u ? 1u : 0u;
            const uint32_t e0 = lds_u32_if(K.t0 + ((bs.pre & K.m0) << 2), rep ^ 1u);
            bits_consume(bs, e0);
            spec |= e0;
            cur = rep ? prev_sym : ((e0 >> 8) & 0xFFu);
            pred_rep -= rep;
            prev_sym = cur;
            syms |= cur << (8 * q);
        } else cur = (grp >> (4 * q)) & 15u;
#pragma unroll
        for (int j = 0; j < 2; j++) {
            const uint32_t pred = (cur >> (2 * j)) & 3u;
            const uint32_t e1 = lds_u32_if(K.t1 + ((bs.pre & K.m1) << 2), pred == 3u ? 1u : 0u);
            bits_consume(bs, e1);
            const uint32_t run = sel_rle != 0u ? 1u : 0u;
            const uint32_t e2 = lds_u32_if(K.t2 + ((bs.pre & K.m2) << 2), run ^ 1u);
            bits_consume(bs, e2);
            spec |= e1 | e2;
            sel_rle -= run;
            t[2 * q + j] = make_uint2((run ? K.num_selectors : (e2 >> 8)) | (pred << 16), e1 >> 8);
        }
    }
    if (spec & kL1Special) return false;
#pragma unroll
    for (int i = 0; i < 2 * NP; i += 2) *reinterpret_cast<uint4*>(tk + i) = make_uint4(t[i].x, t[i].y, t[i + 1].x, t[i + 1].y);
    st.bs = bs; st.sel_rle = sel_rle; st.pred_rep = pred_rep; st.prev_sym = prev_sym;
    if (EVEN) grp = syms;
    return true;
}

static __device__ void etc1s_tokenize(const Etc1sDecodeParams& P, const Etc1sSliceJob& job, PipeShared& W, const uint32_t* l1s,
                                      unsigned long long* predrow, int lane, uint32_t trace_slice)
{
    const uint8_t* __restrict__ data = P.data + job.data_ofs;
    const uint32_t nbx = job.nbx, nby = job.nby;
    // (opaque(): the values are pinned in registers; left alone, ptxas re-derives them from the kernel parameters and the
    // CTA's shared window -- S2UR / LDCU -- in front of every table read, which a lone warp pays in full)
    auto opaque = [](uint32_t v) -> uint32_t { uint32_t r; asm volatile("mov.u32 %0, %1;" : "=r"(r) : "r"(v)); return r; };
    const uint32_t l1_base = smem_u32(l1s);
    TokConsts K;
    K.t0 = opaque(l1_base + 4u * P.l1_ofs[0]); K.t1 = opaque(l1_base + 4u * P.l1_ofs[1]);
    K.t2 = opaque(l1_base + 4u * P.l1_ofs[2]); K.t3 = opaque(l1_base + 4u * P.l1_ofs[3]);
    K.m0 = opaque((1u << P.l1_bits[0]) - 1u); K.m1 = opaque((1u << P.l1_bits[1]) - 1u);
    K.m2 = opaque((1u << P.l1_bits[2]) - 1u); K.m3 = opaque((1u << P.l1_bits[3]) - 1u);
    K.num_selectors = opaque(P.num_selectors & 0xFFFFu);
    K.rle_sym = opaque((P.hist_size + K.num_selectors) & 0xFFFFu);                // mod.rs:220-222 (u16 arithmetic)
    K.is_video = opaque(P.is_video);
    uint32_t* ring = W.ring;
    K.ring_base = opaque(smem_u32(ring));
    const bool fast_ok = P.is_video == 0u;

    // prime the ring with the first three pieces
    uint32_t loaded = 0;                          // pieces [0, loaded) are (or were) in the ring
    for (; loaded < 3; loaded++) { uint4 r[2]; piece_load(data, job.data_len, loaded, lane, r); piece_store(ring, loaded, lane, r); }
    __syncwarp();

    TokState st;
    st.bs.lo = ring[0]; st.bs.hi = ring[1]; st.bs.pre = st.bs.lo; st.bs.x = 64u; st.bs.nw = ring[2]; st.bs.raddr = K.ring_base + 12u;
    uint32_t lap_bytes = 0;                       // slice byte offset of the ring's first slot in the reader's current lap
    st.sel_rle = 0; st.pred_rep = 0; st.prev_sym = 0; st.cur = 0; st.terr = 0;
    uint32_t round = 0;
    K2T_DECL(const long long tt0 = clock64(); uint32_t n_slow = 0; uint32_t n_sym = 0; long long t_wait = 0;)

    for (uint32_t y = 0; y < nby; y++) {
        const bool even = (y & 1u) == 0u;
        for (uint32_t x0 = 0; x0 < nbx; x0 += kRound, round++) {
            const uint32_t nb = nbx - x0 < (uint32_t)kRound ? nbx - x0 : (uint32_t)kRound;
            const uint32_t slot = round % kTokRounds, use = round / kTokRounds;
            if (round >= (uint32_t)kTokRounds) {                                  // the resolver must have read the slot's previous tokens
                K2T_DECL(const long long w0 = clock64();)
                while (!mbar_try_wait_once(&W.bar_empty[slot], (use - 1u) & 1u))
                    if (ld_volatile_shared(&W.abort)) return;
                K2T_DECL(t_wait += clock64() - w0;)
            }
            // start fetching the next ring piece if the reader is about to need it: pieces up to pos+2 must be resident
            // before the next round starts
            if (st.bs.raddr >= K.ring_base + (uint32_t)kRingWords * 4u) { st.bs.raddr -= (uint32_t)kRingWords * 4u; lap_bytes += (uint32_t)kRingWords * 4u; }
            const uint32_t pos = (lap_bytes + (st.bs.raddr - K.ring_base)) / kHalfBytes;
            const bool fetch = loaded < pos + 3u;                                 // warp-uniform
            uint4 pre[2];
            if (fetch) piece_load(data, job.data_len, loaded, lane, pre);

            uint2* tk = W.tok[slot];
            unsigned long long nextp = 0ull;
            unsigned long long curp = even ? 0ull : predrow[x0 >> 5];
            const uint32_t npairs = (nb + 1u) >> 1;
            constexpr int NP = 2;                                                 // pairs per fast-path step
            if (fast_ok && nb == (uint32_t)kRound) {
                // full round: 8 straight-line steps of 4 blocks; a step that meets anything unusual is redone pair by pair
                if (even) {
                    for (uint32_t q = 0; q < (uint32_t)kRound / 2; q += NP) {
                        uint32_t grp = 0u;
                        if (!fast_pairs<true, NP>(st, K, tk + 2u * q, grp)) {
                            K2T_DECL(n_slow++;)
                            for (int i = 0; i < NP && !st.terr; i++) { st = pair_slow(st, K, P, tk + 2u * (q + i), 2u, 1u); grp |= st.cur << (8 * i); }
                            if (st.terr) break;
                        }
#pragma unroll
                        for (int i = 0; i < NP; i++) nextp = (nextp >> 4) | ((unsigned long long)((grp >> (8 * i + 4)) & 15u) << 60);
                    }
                    if (lane == 0) predrow[x0 >> 5] = nextp;
                } else {
                    for (uint32_t q = 0; q < (uint32_t)kRound / 2; q += NP) {
                        uint32_t grp = (uint32_t)curp & ((1u << (4 * NP)) - 1u);
                        curp >>= 4 * NP;
                        if (!fast_pairs<false, NP>(st, K, tk + 2u * q, grp)) {
                            K2T_DECL(n_slow++;)
                            for (int i = 0; i < NP && !st.terr; i++) { st.cur = (grp >> (4 * i)) & 15u; st = pair_slow(st, K, P, tk + 2u * (q + i), 2u, 0u); }
                            if (st.terr) break;
                        }
                    }
                }
            } else if (even) {
                for (uint32_t q = 0; q < npairs; q++) {
                    st = pair_slow(st, K, P, tk + 2u * q, 2u * q + 1u < nb ? 2u : 1u, 1u);
                    if (st.terr) break;
                    nextp = (nextp >> 4) | ((unsigned long long)(st.cur >> 4) << 60);    // pair q ends up at bits 4q .. 4q+3
                }
                if (lane == 0) predrow[x0 >> 5] = nextp >> (4u * (16u - npairs));
            } else {
                for (uint32_t q = 0; q < npairs; q++) {
                    st.cur = (uint32_t)curp & 15u;
                    curp >>= 4;
                    st = pair_slow(st, K, P, tk + 2u * q, 2u * q + 1u < nb ? 2u : 1u, 0u);
                    if (st.terr) break;
                }
            }
            K2T_DECL(n_sym += nb;)
            if (fetch) { piece_store(ring, loaded, lane, pre); loaded++; }
            __syncwarp();                                                         // tokens, ring piece and predictor row are visible to the warp
            if (lane == 0) mbar_arrive(&W.bar_full[slot]);                        // release: the resolver's wait acquires
            if (st.terr) return;
        }
    }
    K2T(0, clock64() - tt0); K2T(1, t_wait); K2T(2, n_sym); K2T(3, n_slow);
}

// ---------------------------------------------------------------------------------------------------
// Wide pipelines, stage 0: helper warps fill the speculation ring.  Helper h of H takes chunks h, h + H, ...
// ---------------------------------------------------------------------------------------------------
// First-level entry -> speculation entry, in two halves so that a helper can have all the global lookups of a chunk in flight
// at once.  A code longer than the first-level table (size 0, flagged) is looked up in the reference's flat table in global
// memory HERE, off the tokenizer's chain: the helpers have the time, so in wide pipelines only run symbols and invalid codes
// are left to the slow path.  v: the 32 stream bits from this position on.
__device__ __forceinline__ uint32_t spec_flat_lookup(uint32_t e, uint32_t v, const uint32_t* __restrict__ flat, uint32_t max_len)
{
    const bool need = (e & kL1Special) != 0u && (e & 31u) == 0u;            // huffman.rs:186-198
    return need ? (__ldg(flat + (v & ((1u << max_len) - 1u))) | 0x80000000u) : 0u;        // (symbols have 16 bits: bit 31 is free)
}
__device__ __forceinline__ uint32_t spec_entry(uint32_t e, uint32_t f, uint32_t run_sym, bool is_pred)
{
    uint32_t sym = e >> 8, len = e & 31u;
    if ((e & kL1Special) != 0u) {
        if (len != 0u) return kSpecSpecial;                                   // a run symbol
        sym = (f & 0x7FFFFFFFu) >> 5; len = f & 31u;
        if (len == 0u || sym == run_sym) return kSpecSpecial;
    }
    if (len == 0u || sym > 0xFFFFu) return kSpecSpecial;
    uint32_t r = (len << 26) | sym;
    if (is_pred) r |= ((sym & 3u) == 3u ? kSpecDelta0 : 0u) | ((sym & 12u) == 12u ? kSpecDelta1 : 0u);
    return r;
}

static __device__ void etc1s_speculate(const Etc1sDecodeParams& P, PipeShared& W, SpecShared& S, const uint32_t* l1s, uint32_t helper, uint32_t helpers, int lane)
{
    const uint32_t* t0 = l1s + P.l1_ofs[0];
    const uint32_t* t1 = l1s + P.l1_ofs[1];
    const uint32_t* t2 = l1s + P.l1_ofs[2];
    const uint32_t m0 = (1u << P.l1_bits[0]) - 1u, m1 = (1u << P.l1_bits[1]) - 1u, m2 = (1u << P.l1_bits[2]) - 1u;
    const uint32_t rle_sym = (P.hist_size + (P.num_selectors & 0xFFFFu)) & 0xFFFFu;       // mod.rs:220-222 (u16 arithmetic)
    volatile uint32_t* consumed = &S.consumed;
    volatile uint32_t* raw_loaded = &S.raw_loaded;
    volatile uint32_t* stop = &S.stop;
    for (uint32_t c = helper;; c += helpers) {
        // the ring slot must be free (the reader is past chunk c - kSpecChunks) and the chunk's bytes (and the word after) resident
        const uint32_t need_bytes = c * (kSpecChunk / 8) + kSpecChunk / 8 + 8u;
        while (!(c < *consumed + (uint32_t)kSpecChunks && need_bytes <= *raw_loaded)) {
            if (*stop) return;
            __nanosleep(64);
        }
        asm volatile("" ::: "memory");
        const uint32_t slot = c % kSpecChunks;
        uint32_t* stage = S.stage[helper];
        constexpr int G = kSpecChunk / 32;
        uint32_t ep[G], ed[G], es[G + 1], vv[G + 1];
#pragma unroll
        for (int g = 0; g <= G; g++) {                                                           // first-level entries
            const uint32_t w = (c * (uint32_t)G + (uint32_t)g) % (uint32_t)kRingWords;           // (the ring's first piece is mirrored behind it)
            const uint32_t v = __funnelshift_r(W.ring[w], W.ring[w + 1u], (uint32_t)lane);
            vv[g] = v;
            es[g] = t2[v & m2];
            if (g < G) { ep[g] = t0[v & m0]; ed[g] = t1[v & m1]; }
        }
        {
            uint32_t fp[G], fd[G], fs[G + 1];
#pragma unroll
            for (int g = 0; g <= G; g++) {                                                       // the long codes' lookups, all in flight together
                fs[g] = spec_flat_lookup(es[g], vv[g], P.flat[2], P.max_len[2]);
                if (g < G) { fp[g] = spec_flat_lookup(ep[g], vv[g], P.flat[0], P.max_len[0]); fd[g] = spec_flat_lookup(ed[g], vv[g], P.flat[1], P.max_len[1]); }
            }
#pragma unroll
            for (int g = 0; g <= G; g++) {
                stage[g * 32 + lane] = es[g] = spec_entry(es[g], fs[g], rle_sym, false);
                if (g < G) { ep[g] = spec_entry(ep[g], fp[g], 256u, true); ed[g] = spec_entry(ed[g], fd[g], 0xFFFFFFFFu, false); }
            }
        }
        __syncwarp();
#pragma unroll
        for (int g = 0; g < kSpecChunk / 32; g++) {
            const uint32_t i = slot * kSpecChunk + (uint32_t)g * 32u + (uint32_t)lane;
            const uint32_t d = ed[g];
            const uint32_t es2 = stage[g * 32 + lane + (d >> 26)];                               // the selector symbol behind this delta symbol
            const bool bad = ((d | es2) & kSpecSpecial) != 0u;
            const uint32_t da = bad ? ((d & 0xFFFFu) | kSpecSpecial) : ((d & 0xFFFFu) | ((d & 0xFF000000u) + (es2 & 0xFF000000u)));
            const uint32_t db = (es2 & 0xFFFFu) | (d & 0xFF000000u);
            S.ent[0][i] = ep[g]; S.ent[1][i] = da; S.ent[2][i] = es[g]; S.ent[3][i] = db;
            if (slot < (uint32_t)kSpecMirrorChunks) { S.ent[0][i + kSpecPos] = ep[g]; S.ent[1][i + kSpecPos] = da; S.ent[2][i + kSpecPos] = es[g]; S.ent[3][i + kSpecPos] = db; }
        }
        __syncwarp();                                                                            // the stage is reused by the next chunk
        __threadfence_block();
        __syncwarp();
        if (lane == 0) *reinterpret_cast<volatile uint32_t*>(&S.done[slot]) = c + 1u;
    }
}

// One predicated read of a speculation entry: the code size (x 4) for the chain and the whole word for the token.
__device__ __forceinline__ void spec_ld(uint32_t addr, uint32_t on, uint32_t& len4, uint32_t& e)
{
    asm volatile("{\n.reg .pred p;\nsetp.ne.u32 p, %3, 0;\nmov.u32 %0, 0;\nmov.u32 %1, 0;\n@p ld.shared.u8 %0, [%2+3];\n@p ld.shared.u32 %1, [%2];\n}\n"
                 : "=&r"(len4), "=&r"(e) : "r"(addr), "r"(on) : "memory");
}
// ... and of a predictor symbol's entry: when it is skipped (inside a repeat run) `e` keeps the previous symbol's word
__device__ __forceinline__ void spec_ld_keep(uint32_t addr, uint32_t on, uint32_t& len4, uint32_t& e)
{
    asm volatile("{\n.reg .pred p;\nsetp.ne.u32 p, %3, 0;\nmov.u32 %0, 0;\n@p ld.shared.u8 %0, [%2+3];\n@p ld.shared.u32 %1, [%2];\n}\n"
                 : "=&r"(len4), "+r"(e) : "r"(addr), "r"(on) : "memory");
}

struct WideState { uint32_t q, sel_rle, pred_rep, prev_e; };      // q: 4 * position in the speculation ring; prev_e: entry of the last predictor symbol
struct WideConsts { uint32_t ent_p, num_selectors; };             // ent_p: shared address of the ring's first table

// The wide form of fast_pairs: same tokens, same "anything unusual -> false, nothing written".
template <bool EVEN, int NP>
__device__ __forceinline__ bool fast_pairs_wide(WideState& st, const WideConsts& K, uint2* tk, uint32_t& grp)
{
    uint32_t a = st.q + K.ent_p, spec = 0u, sel_rle = st.sel_rle, pred_rep = st.pred_rep, prev_e = st.prev_e, syms = 0u;
    uint2 t[2 * NP];
#pragma unroll
    for (int p = 0; p < NP; p++) {
        uint32_t cur, d0, d1;
        if (EVEN) {
            const uint32_t rep = pred_rep != 0u ? 1u : 0u;
            uint32_t l0;
            spec_ld_keep(a, rep ^ 1u, l0, prev_e);
            a += l0;
            spec |= prev_e;
            cur = prev_e & 0xFFu;
            d0 = prev_e & kSpecDelta0; d1 = prev_e & kSpecDelta1;
            pred_rep -= rep;
            syms |= cur << (8 * p);
        } else {
            cur = (grp >> (4 * p)) & 15u;
            d0 = (cur & 3u) == 3u ? 1u : 0u; d1 = (cur & 12u) == 12u ? 1u : 0u;
        }
#pragma unroll
        for (int j = 0; j < 2; j++) {
            uint32_t l1, e1, l2, e2;
            asm volatile("{\n.reg .pred p;\nsetp.ne.u32 p, %3, 0;\nmov.u32 %0, 0;\nmov.u32 %1, 0;\n@p ld.shared.u8 %0, [%2+%4];\n@p ld.shared.u32 %1, [%2+%5];\n}\n"
                         : "=&r"(l1), "=&r"(e1) : "r"(a), "r"(j ? d1 : d0), "n"(3u * kSpecTableBytes + 3u), "n"(kSpecTableBytes) : "memory");
            a += l1;
            const uint32_t run = sel_rle != 0u ? 1u : 0u;
            spec_ld(a + 2u * kSpecTableBytes, run ^ 1u, l2, e2);
            a += l2;
            spec |= e1 | e2;
            sel_rle -= run;
            t[2 * p + j] = make_uint2((run ? K.num_selectors : (e2 & 0xFFFFu)) | (((cur >> (2 * j)) & 3u) << 16), e1 & 0xFFFFu);
        }
    }
    if (spec & kSpecSpecial) return false;
#pragma unroll
    for (int i = 0; i < 2 * NP; i += 2) *reinterpret_cast<uint4*>(tk + i) = make_uint4(t[i].x, t[i].y, t[i + 1].x, t[i + 1].y);
    st.q = a - K.ent_p; st.sel_rle = sel_rle; st.pred_rep = pred_rep; st.prev_e = prev_e;
    if (EVEN) grp = syms;
    return true;
}

// The same for kLeanPairs pairs while no run is active (runs only start in the slow path): every selector and predictor
// symbol is read, so only the delta read is conditional -- and that one is predicated together with its add, nothing has
// to be zeroed.  One running address; the three tables sit at constant distances (immediates of the loads), so a link is
// LDS.U8 -> IADD -> LDS.U8.  (A lone warp issues slowly: the instruction count matters here as much as the chain.)
constexpr int kLeanPairs = 4;
// SELRUN: a selector run covers all eight blocks (no selector symbols; a delta symbol's own size comes from table 3);
// PREDREP (even rows): a predictor repeat run covers all four pairs (no predictor symbols).  Flat image regions are made of these.
template <bool EVEN, bool SELRUN = false, bool PREDREP = false>
__device__ __forceinline__ bool lean_pairs(WideState& st, const WideConsts& K, uint2* tk, uint32_t& grp)
{
    uint32_t a = st.q + K.ent_p, spec = 0u, prev_e = st.prev_e, syms = 0u, e1 = 0u;
    uint2 t[2 * kLeanPairs];
#pragma unroll
    for (int p = 0; p < kLeanPairs; p++) {
        uint32_t cur, d0, d1;
        if (EVEN) {
            if (!PREDREP) {
                uint32_t l0;
                asm volatile("ld.shared.u8 %0, [%2+3];\nld.shared.u32 %1, [%2];\n" : "=&r"(l0), "=&r"(prev_e) : "r"(a) : "memory");
                a += l0;
                spec |= prev_e;
            }
            cur = prev_e & 0xFFu;
            d0 = prev_e & kSpecDelta0; d1 = prev_e & kSpecDelta1;
            syms |= cur << (8 * p);
        } else {
            cur = (grp >> (4 * p)) & 15u;
            d0 = (cur & 3u) == 3u ? 1u : 0u; d1 = (cur & 12u) == 12u ? 1u : 0u;
        }
#pragma unroll
        for (int j = 0; j < 2; j++) {
            if (SELRUN) {
                asm volatile("{\n.reg .pred p;\n.reg .u32 l;\nsetp.ne.u32 p, %3, 0;\n@p ld.shared.u8 l, [%0+%4];\n@p ld.shared.u32 %1, [%0+%5];\n@p add.u32 %0, %0, l;\n@p or.b32 %2, %2, %1;\n}\n"
                             : "+r"(a), "+r"(e1), "+r"(spec) : "r"(j ? d1 : d0), "n"(3u * kSpecTableBytes + 3u), "n"(kSpecTableBytes) : "memory");
                t[2 * p + j] = make_uint2(K.num_selectors | (((cur >> (2 * j)) & 3u) << 16), e1 & 0xFFFFu);
            } else {
                uint32_t l, e2;
                // with a delta symbol: table 1 (both symbols' sizes, the delta symbol) + table 3 (the selector symbol behind it);
                // without: table 2.  Exactly one of the two groups runs, so `l` and `e2` are always written -- but ptxas takes two
                // complementary predicated writes for a partial definition and keeps an "old value" alive (spills, in a kernel of
                // this size), hence the explicit zero in front.
                asm volatile("{\n.reg .pred p;\n.reg .b64 zz;\nsetp.ne.u32 p, %5, 0;\nmov.b64 zz, 0;\nmov.b64 {%0, %1}, zz;\n"
                             "@p ld.shared.u8 %0, [%4+%6];\n@p ld.shared.u32 %2, [%4+%7];\n@p ld.shared.u16 %1, [%4+%8];\n"
                             "@!p ld.shared.u8 %0, [%4+%9];\n@!p ld.shared.u32 %1, [%4+%10];\n"
                             "@p or.b32 %3, %3, %2;\n}\n"
                             : "=&r"(l), "=&r"(e2), "+r"(e1), "+r"(spec)
                             : "r"(a), "r"(j ? d1 : d0), "n"(kSpecTableBytes + 3u), "n"(kSpecTableBytes), "n"(3u * kSpecTableBytes),
                               "n"(2u * kSpecTableBytes + 3u), "n"(2u * kSpecTableBytes) : "memory");
                a += l;
                spec |= e2;                                                           // (the u16 load from table 3 brings no flag bits)
                t[2 * p + j] = make_uint2((e2 & 0xFFFFu) | (((cur >> (2 * j)) & 3u) << 16), e1 & 0xFFFFu);
            }
        }
    }
    if (spec & kSpecSpecial) return false;
#pragma unroll
    for (int i = 0; i < 2 * kLeanPairs; i += 2) *reinterpret_cast<uint4*>(tk + i) = make_uint4(t[i].x, t[i].y, t[i + 1].x, t[i + 1].y);
    st.q = a - K.ent_p; st.prev_e = prev_e;
    if (SELRUN) st.sel_rle -= 2u * kLeanPairs;
    if (EVEN && PREDREP) st.pred_rep -= (uint32_t)kLeanPairs;
    if (EVEN) grp = syms;
    return true;
}

// A whole round of lean steps with the state in registers throughout.  Returns the number of steps done (kRound / 2 /
// kLeanPairs = all); the step it stops in front of has left no trace.  clo / chi: this row's predictor bits (odd rows);
// nlo / nhi: the bits for the row below (even rows), pair i at bits 4 i .. 4 i + 3.
template <bool EVEN>
__device__ __forceinline__ int lean_round(WideState& st, const WideConsts& K, uint2* tk, uint32_t clo, uint32_t chi, uint32_t& nlo, uint32_t& nhi)
{
    constexpr int kSteps = kRound / 2 / kLeanPairs;
#pragma unroll
    for (int sidx = 0; sidx < kSteps; sidx++) {
        uint32_t grp = ((sidx < kSteps / 2 ? clo : chi) >> (16 * (sidx % (kSteps / 2)))) & 0xFFFFu;
        if (!lean_pairs<EVEN>(st, K, tk + 2 * kLeanPairs * sidx, grp)) return sidx;
        if (EVEN) {                                                                   // the high nibble of each pair's symbol belongs to the row below
            uint32_t n = (grp >> 4) & 0x0F0F0F0Fu;
            n = (n | (n >> 4)) & 0x00FF00FFu;
            n = (n | (n >> 8)) & 0xFFFFu;
            if (sidx < kSteps / 2) nlo |= n << (16 * (sidx % (kSteps / 2))); else nhi |= n << (16 * (sidx % (kSteps / 2)));
        }
    }
    return kSteps;
}

// Wide pipelines, stage 1.  The position is (base_chunk * kSpecChunk + q / 4) in bits; the bit buffer of the narrow tokenizer
// only exists while pair_slow runs (rebuilt from the compressed-byte ring at the current position, and turned back into a
// position).  Everything off the fast path is out of line and gets / returns the state BY VALUE, so that the fast path's
// state never has an address (a lone warp cannot hide a single local-memory access).
#ifdef B2BU_K2_TRACE
struct WideCtx { WideState ws; uint32_t base_chunk, done_upto, terr, gone, aux, aux2; };
#else
struct WideCtx { WideState ws; uint32_t base_chunk, done_upto, terr, gone, aux; };
#endif   // done_upto: chunks [0, done_upto) are known complete

__device__ __forceinline__ void st_shared_lane0(void* p, uint32_t v, int lane)    // no divergent region: a lone warp pays ~50 cycles for BSSY / BSYNC
{
    asm volatile("{\n.reg .pred p;\nsetp.eq.u32 p, %2, 0;\n@p st.volatile.shared.u32 [%0], %1;\n}\n" ::"r"(smem_u32(p)), "r"(v), "r"(lane) : "memory");
}

// wrap, tell the helpers where the reader is, and wait until the fast path's reach from here is covered by complete chunks
__device__ __forceinline__ void wide_sync(WideCtx& c, SpecShared& S, PipeShared& W, int lane)
{
    while (c.ws.q >= (uint32_t)kSpecPos * 4u) { c.ws.q -= (uint32_t)kSpecPos * 4u; c.base_chunk += (uint32_t)kSpecChunks; }
    const uint32_t here = c.base_chunk + (c.ws.q >> 2) / (uint32_t)kSpecChunk;
    st_shared_lane0(&S.consumed, here, lane);                                          // nothing before this chunk is looked at again
    const uint32_t need = here + (uint32_t)kSpecAhead;
    while (c.done_upto < need) {
        K2T_DECL(const long long w0 = clock64();)
        while (ld_volatile_shared(&S.done[c.done_upto % kSpecChunks]) != c.done_upto + 1u)
            if (ld_volatile_shared(&W.abort)) { c.gone = 1u; return; }
        K2T_DECL(c.aux2 += (uint32_t)(clock64() - w0);)
        c.done_upto++;
    }
    asm volatile("" ::: "memory");
}

// the reference's control flow for one pair, from and back to the position; aux returns the pair's predictor bits
__device__ __noinline__ WideCtx wide_slow_pair(WideCtx c, const TokConsts K, const Etc1sDecodeParams& P, const uint32_t* ring, uint2* tk, uint32_t nblk,
                                               uint32_t even_row, uint32_t cur)
{
    const uint32_t pos = (c.base_chunk % (uint32_t)(kRingWords * 32 / kSpecChunk)) * (uint32_t)kSpecChunk + (c.ws.q >> 2);     // modulo the byte ring
    const uint32_t w = (pos >> 5) % (uint32_t)kRingWords, sh = pos & 31u;
    TokState st;
    st.bs.lo = __funnelshift_r(ring[w], ring[w + 1u], sh);
    st.bs.hi = ring[w + 1u] >> sh;
    st.bs.pre = st.bs.lo; st.bs.x = 64u - sh; st.bs.nw = ring[w + 2u]; st.bs.raddr = K.ring_base + 4u * (w + 3u);
    const uint32_t raddr0 = st.bs.raddr, x0 = st.bs.x;
    st.sel_rle = c.ws.sel_rle; st.pred_rep = c.ws.pred_rep; st.prev_sym = c.ws.prev_e & 0xFFu; st.cur = cur; st.terr = 0u;
    st = pair_slow(st, K, P, tk, nblk, even_row);
    const uint32_t used = ((st.bs.raddr - raddr0) << 3) + x0 - (st.bs.x & 0xFFu);
    c.ws.q += used << 2;
    c.ws.sel_rle = st.sel_rle; c.ws.pred_rep = st.pred_rep;
    const uint32_t ps = st.prev_sym & 0xFFu;
    c.ws.prev_e = ps | ((ps & 3u) == 3u ? kSpecDelta0 : 0u) | ((ps & 12u) == 12u ? kSpecDelta1 : 0u);
    c.terr = st.terr;
    c.aux = st.cur;
    return c;
}

// the general form of a lean step: two fast steps of two pairs, each falling back to the slow path pair by pair.
// grp4 in: the step's predictor bits (odd rows); aux out: the four predictor symbols (even rows)
__device__ __noinline__ WideCtx wide_general_step(WideCtx c, const TokConsts K, const WideConsts KW, const Etc1sDecodeParams& P, SpecShared& S, PipeShared& W,
                                                  uint2* tk4, uint32_t even, uint32_t grp4, int lane)
{
    constexpr int NP = 2;
    uint32_t out = 0u;
    for (uint32_t h = 0; h < (uint32_t)kLeanPairs && !c.terr && !c.gone; h += NP) {
        uint32_t grp = even ? 0u : (grp4 >> (4u * h)) & ((1u << (4 * NP)) - 1u);
        const uint32_t g0 = grp;
        if (!(even ? fast_pairs_wide<true, NP>(c.ws, KW, tk4 + 2u * h, grp) : fast_pairs_wide<false, NP>(c.ws, KW, tk4 + 2u * h, grp))) {
            grp = 0u;
            for (int i = 0; i < NP && !c.terr && !c.gone; i++) {
                c = wide_slow_pair(c, K, P, W.ring, tk4 + 2u * (h + i), 2u, even, even ? 0u : (g0 >> (4 * i)) & 15u);
                grp |= c.aux << (8 * i);
                wide_sync(c, S, W, lane);                                             // a slow pair may have read run lengths: re-establish the reach
            }
        }
        out |= grp << (8u * h);
    }
    c.aux = out;
    return c;
}

static __device__ void etc1s_tokenize_wide(const Etc1sDecodeParams& P, const Etc1sSliceJob& job, PipeShared& W, SpecShared& S, const uint32_t* l1s,
                                           unsigned long long* predrow, int lane, uint32_t trace_slice)
{
    const uint8_t* __restrict__ data = P.data + job.data_ofs;
    const uint32_t nbx = job.nbx, nby = job.nby;
    auto opaque = [](uint32_t v) -> uint32_t { uint32_t r; asm volatile("mov.u32 %0, %1;" : "=r"(r) : "r"(v)); return r; };
    const uint32_t l1_base = smem_u32(l1s);
    TokConsts K;
    K.t0 = opaque(l1_base + 4u * P.l1_ofs[0]); K.t1 = opaque(l1_base + 4u * P.l1_ofs[1]);
    K.t2 = opaque(l1_base + 4u * P.l1_ofs[2]); K.t3 = opaque(l1_base + 4u * P.l1_ofs[3]);
    K.m0 = opaque((1u << P.l1_bits[0]) - 1u); K.m1 = opaque((1u << P.l1_bits[1]) - 1u);
    K.m2 = opaque((1u << P.l1_bits[2]) - 1u); K.m3 = opaque((1u << P.l1_bits[3]) - 1u);
    K.num_selectors = opaque(P.num_selectors & 0xFFFFu);
    K.rle_sym = opaque((P.hist_size + K.num_selectors) & 0xFFFFu);
    K.is_video = opaque(P.is_video);
    uint32_t* ring = W.ring;
    K.ring_base = opaque(smem_u32(ring));
    WideConsts KW;
    KW.ent_p = opaque(smem_u32(S.ent[0]));
    KW.num_selectors = K.num_selectors;

    uint32_t loaded = 0;
    for (; loaded < 3; loaded++) { uint4 r[2]; piece_load(data, job.data_len, loaded, lane, r); piece_store(ring, loaded, lane, r); }
    __threadfence_block();
    __syncwarp();
    st_shared_lane0(&S.raw_loaded, loaded * (uint32_t)kHalfBytes, lane);

    WideCtx c;
    c.ws.q = 0u; c.ws.sel_rle = 0u; c.ws.pred_rep = 0u; c.ws.prev_e = 0u;
    c.base_chunk = 0u; c.done_upto = 0u; c.terr = 0u; c.gone = 0u; c.aux = 0u;
    K2T_DECL(c.aux2 = 0u;)
    uint32_t round = 0;
    uint32_t free_upto = (uint32_t)kTokRounds;    // token slots of rounds [0, free_upto) are known to be free
    K2T_DECL(const long long tt0 = clock64(); uint32_t n_slow = 0; uint32_t n_sym = 0; long long t_wait = 0; uint32_t n_lean = 0;)

    for (uint32_t y = 0; y < nby; y++) {
        const bool even = (y & 1u) == 0u;
        for (uint32_t x0 = 0; x0 < nbx; x0 += kRound, round++) {
            const uint32_t nb = nbx - x0 < (uint32_t)kRound ? nbx - x0 : (uint32_t)kRound;
            const uint32_t slot = round % kTokRounds;
            if (round >= free_upto) {                                               // the resolver counts the rounds it has taken out of the ring
                K2T_DECL(const long long w0 = clock64();)
                uint32_t taken;
                while ((taken = ld_volatile_shared(&W.taken)) + (uint32_t)kTokRounds <= round)
                    if (ld_volatile_shared(&W.abort)) return;
                free_upto = taken + (uint32_t)kTokRounds;
                K2T_DECL(t_wait += clock64() - w0;)
            }
            wide_sync(c, S, W, lane);
            if (c.gone) return;
            K2T_DECL(t_wait += c.aux2; c.aux2 = 0u;)
            const uint32_t piece = (c.base_chunk + (c.ws.q >> 2) / (uint32_t)kSpecChunk) / (uint32_t)(8 * kHalfBytes / kSpecChunk);
            const bool fetch = loaded < piece + 3u;                                 // warp-uniform
            uint4 pre[2];
            if (fetch) piece_load(data, job.data_len, loaded, lane, pre);

            uint2* tk = W.tok[slot];
            const uint32_t npairs = (nb + 1u) >> 1;
            if (nb == (uint32_t)kRound) {
                static_assert(kLeanPairs == 4 && kRound == 32, "lean_round's packing of the predictor bits");
                constexpr int kSteps = kRound / 2 / kLeanPairs;
                uint32_t nlo = 0u, nhi = 0u;                                          // next row's predictor bits: pair i at bits 4 i .. 4 i + 3
                uint32_t clo = 0u, chi = 0u;
                if (!even) { const uint2 c2 = *reinterpret_cast<const uint2*>(predrow + (x0 >> 5)); clo = c2.x; chi = c2.y; }
                int sidx = 0;
                if ((c.ws.sel_rle | c.ws.pred_rep) == 0u) sidx = even ? lean_round<true>(c.ws, KW, tk, clo, chi, nlo, nhi) : lean_round<false>(c.ws, KW, tk, clo, chi, nlo, nhi);
                K2T_DECL(n_lean += (uint32_t)sidx;)
#pragma unroll 1
                for (; sidx < kSteps; sidx++) {                                       // runs are active, or something unusual is ahead
                    uint32_t grp4 = ((sidx < kSteps / 2 ? clo : chi) >> (16 * (sidx % (kSteps / 2)))) & 0xFFFFu;
                    uint2* tk4 = tk + 2 * kLeanPairs * sidx;
                    const bool sr = c.ws.sel_rle >= 2u * kLeanPairs, s0 = c.ws.sel_rle == 0u;
                    const bool pr = c.ws.pred_rep >= (uint32_t)kLeanPairs, p0 = c.ws.pred_rep == 0u;
                    bool ok = false;
                    if (even) {
                        if ((sr || s0) && (pr || p0))
                            ok = sr ? (pr ? lean_pairs<true, true, true>(c.ws, KW, tk4, grp4) : lean_pairs<true, true, false>(c.ws, KW, tk4, grp4))
                                    : (pr ? lean_pairs<true, false, true>(c.ws, KW, tk4, grp4) : lean_pairs<true, false, false>(c.ws, KW, tk4, grp4));
                    } else if (sr || s0) ok = sr ? lean_pairs<false, true, false>(c.ws, KW, tk4, grp4) : lean_pairs<false, false, false>(c.ws, KW, tk4, grp4);
                    if (!ok) {                                                        // the general form
                        K2T_DECL(n_slow++;)
                        c = wide_general_step(c, K, KW, P, S, W, tk4, even ? 1u : 0u, grp4, lane);
                        if (c.gone) return;
                        if (c.terr) break;
                        grp4 = c.aux;
                    }
                    if (even) {
                        uint32_t n = (grp4 >> 4) & 0x0F0F0F0Fu;
                        n = (n | (n >> 4)) & 0x00FF00FFu;
                        n = (n | (n >> 8)) & 0xFFFFu;
                        if (sidx < kSteps / 2) nlo |= n << (16 * (sidx % (kSteps / 2))); else nhi |= n << (16 * (sidx % (kSteps / 2)));
                    }
                }
                if (even) asm volatile("{\n.reg .pred p;\nsetp.eq.u32 p, %3, 0;\n@p st.v2.u32 [%0], {%1, %2};\n}\n" ::"l"(predrow + (x0 >> 5)), "r"(nlo), "r"(nhi), "r"(lane) : "memory");
            } else if (even) {
                unsigned long long nextp = 0ull;
                for (uint32_t q = 0; q < npairs; q++) {
                    c = wide_slow_pair(c, K, P, ring, tk + 2u * q, 2u * q + 1u < nb ? 2u : 1u, 1u, 0u);
                    if (c.terr) break;
                    nextp = (nextp >> 4) | ((unsigned long long)(c.aux >> 4) << 60);  // pair q ends up at bits 4q .. 4q+3
                }
                if (lane == 0) predrow[x0 >> 5] = nextp >> (4u * (16u - npairs));
            } else {
                unsigned long long curp = predrow[x0 >> 5];
                for (uint32_t q = 0; q < npairs; q++) {
                    c = wide_slow_pair(c, K, P, ring, tk + 2u * q, 2u * q + 1u < nb ? 2u : 1u, 0u, (uint32_t)curp & 15u);
                    curp >>= 4;
                    if (c.terr) break;
                }
            }
            K2T_DECL(n_sym += nb;)
            if (fetch) { piece_store(ring, loaded, lane, pre); loaded++; __threadfence_block(); }
            __syncwarp();
            if (fetch) st_shared_lane0(&S.raw_loaded, loaded * (uint32_t)kHalfBytes, lane);
            asm volatile("{\n.reg .pred p;\nsetp.eq.u32 p, %1, 0;\n@p mbarrier.arrive.shared::cta.b64 _, [%0];\n}\n" ::"r"(smem_u32(&W.bar_full[slot])), "r"(lane) : "memory");
            if (c.terr) return;
        }
    }
    K2T(0, clock64() - tt0); K2T(1, t_wait); K2T(2, n_sym); K2T(3, n_slow); K2T(15, n_lean);
    K2T_DECL(if (lane == 0 && trace_slice < 64u) { uint32_t smid; asm("mov.u32 %0, %%smid;" : "=r"(smid)); g_k2trace2[trace_slice][0] = smid; g_k2trace2[trace_slice][1] = (unsigned long long)tt0; })
}

// ---------------------------------------------------------------------------------------------------
// Stage 2: resolver warp.  rowep: endpoint index of every block of the previous row.
// ---------------------------------------------------------------------------------------------------
static __device__ void etc1s_resolve(const Etc1sDecodeParams& P, const Etc1sSliceJob& job, PipeShared& W, uint16_t* rowep, uint16_t* hist,
                                     uint32_t* __restrict__ out, uint32_t* status, int lane, uint32_t trace_slice)
{
    const uint32_t nbx = job.nbx, nby = job.nby;
    const uint32_t num_endpoints = P.num_endpoints, num_selectors = P.num_selectors, hist_size = P.hist_size;
    const bool is_video = P.is_video != 0u;
    // the scan needs (d + prev) mod n == the reference's 16-bit wrap-and-subtract: true when both terms are < n <= 32768
    const bool scan_ok_n = num_endpoints <= 32768u;
    for (uint32_t i = lane; i < hist_size; i += 32) hist[i] = 0;                   // mod.rs:616-621
    __syncwarp();
    uint32_t rover = hist_size / 2, prev_ep = 0, carry_up = 0, round = 0;
    K2T_DECL(const long long tt0 = clock64(); long long t_wait = 0; uint32_t n_hit = 0; uint32_t n_serial = 0; long long t_seg[4] = {0, 0, 0, 0};)

    for (uint32_t y = 0; y < nby; y++) {
        for (uint32_t x0 = 0; x0 < nbx; x0 += kRound, round++) {
            const uint32_t nb = nbx - x0 < (uint32_t)kRound ? nbx - x0 : (uint32_t)kRound;
            const uint32_t slot = round % kTokRounds, use = round / kTokRounds;
            const uint32_t x = x0 + lane;
            // previous row (independent of the tokens: loaded while waiting)
            uint32_t up = 0, upl = carry_up, up_last = 0;
            if (y > 0u) {
                if (x < nbx) up = rowep[x];
                if (lane > 0 && x - 1u < nbx) upl = rowep[x - 1u];
                if (x0 + 31u < nbx) up_last = rowep[x0 + 31u];
            }
            carry_up = up_last;
            K2T_DECL(const long long w0 = clock64();)
            mbar_wait(&W.bar_full[slot], use & 1u);
            K2T_DECL(t_wait += clock64() - w0; const long long s0 = clock64();)
            uint2 tok = make_uint2(0u, 0u);
            if ((uint32_t)lane < nb) tok = W.tok[slot][lane];
            __syncwarp();
            if (lane == 0) { mbar_arrive(&W.bar_empty[slot]); *reinterpret_cast<volatile uint32_t*>(&W.taken) = round + 1u; }

            const uint32_t errmask = __ballot_sync(0xFFFFFFFFu, (tok.x & kTokErr) != 0u);
            const uint32_t nv = errmask ? (uint32_t)__ffs((int)errmask) - 1u : nb;  // lanes < nv hold complete tokens
            const uint32_t pred = (tok.x >> 16) & 3u, d = tok.y;
            uint32_t key = 0xFFFFFFFFu;                                              // (block << 3 | phase) << 8 | code of this lane's first error
            if (errmask && (uint32_t)lane == nv) key = (((uint32_t)lane << 3 | ((tok.x >> 28) & 7u)) << 8) | ((tok.x >> 24) & 15u);
            const bool has_pred = (uint32_t)lane < nv || (errmask && (uint32_t)lane == nv && ((tok.x >> 28) & 7u) > (uint32_t)PH_PRED_CHECK);
            if (has_pred) {                                                          // asserts mod.rs:304-339
                const bool bad = (pred == 0u && x == 0u) || (pred == 1u && y == 0u) || (pred == 2u && !is_video && (x == 0u || y == 0u));
                if (bad) { const uint32_t k2 = (((uint32_t)lane << 3 | PH_PRED_CHECK) << 8) | ETC1S_ERR_PREDICTION; key = k2 < key ? k2 : key; }
            }

            // ---- endpoints: f_b(prev) = const c (up / up-left / video 0), prev (left) or (prev + d) mod n ----
            K2T_DECL(const long long s1 = clock64(); t_seg[0] += s1 - s0;)
            uint32_t ep;
            const bool live = (uint32_t)lane < nv;
            const bool scan_ok = scan_ok_n && !__any_sync(0xFFFFFFFFu, live && pred == 3u && d >= num_endpoints);
            if (scan_ok) {
                uint32_t f = 0u;                                                     // bit 31: constant; low bits: value
                if (live) f = pred == 1u ? (0x80000000u | up) : pred == 2u ? (0x80000000u | (is_video ? 0u : upl)) : pred == 3u ? d : 0u;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t g = __shfl_up_sync(0xFFFFFFFFu, f, o);
                    if (lane >= o && !(f >> 31)) {
                        uint32_t v = (g & 0x7FFFFFFFu) + f;
                        if (v >= num_endpoints) v -= num_endpoints;
                        f = (g & 0x80000000u) | v;
                    }
                }
                ep = f & 0x7FFFFFFFu;
                if (!(f >> 31)) { ep += prev_ep; if (ep >= num_endpoints) ep -= num_endpoints; }
            } else {
                ep = 0;
                K2T_DECL(n_serial++;)
                uint32_t pe = prev_ep;
                for (uint32_t b = 0; b < nv; b++) {                                  // the reference's own arithmetic, block by block
                    const uint32_t pb = __shfl_sync(0xFFFFFFFFu, pred, b), db = __shfl_sync(0xFFFFFFFFu, d, b);
                    const uint32_t ub = __shfl_sync(0xFFFFFFFFu, up, b), ulb = __shfl_sync(0xFFFFFFFFu, upl, b);
                    uint32_t e;
                    if (pb == 0u) e = pe;
                    else if (pb == 1u) e = ub;
                    else if (pb == 2u) e = is_video ? 0u : ulb;
                    else { e = (db + pe) & 0xFFFFu; if (e >= num_endpoints) e = (e - num_endpoints) & 0xFFFFu; }
                    pe = e;
                    if ((uint32_t)lane == b) ep = e;
                }
            }
            if (nv > 0u) prev_ep = __shfl_sync(0xFFFFFFFFu, ep, nv - 1u);

            // ---- selectors: the history buffer is the serial part (mod.rs:399-426, :610-640) ----
            K2T_DECL(const long long s2 = clock64(); t_seg[1] += s2 - s1;)
            uint32_t sel = 0u;
            const uint32_t sym_l = tok.x & 0xFFFFu;
            const bool hit_l = sym_l >= num_selectors;
            // a round without real history references (index 0 -- the run repeat -- swaps with itself and entry 0 is outside the
            // rover's range [size / 2, size) when size >= 2): the inserts are a scatter, everything else reads entry 0
            if (!is_video && hist_size >= 2u && hist_size <= 64u && nv == (uint32_t)kRound &&
                __ballot_sync(0xFFFFFFFFu, hit_l && sym_l != num_selectors) == 0u) {
                uint16_t* hs = W.hist;
                const uint32_t ins = __ballot_sync(0xFFFFFFFFu, !hit_l);
                const uint32_t n_ins = (uint32_t)__popc(ins), r = (uint32_t)__popc(ins & ((1u << lane) - 1u));
                const uint32_t half = hist_size / 2u, period = hist_size - half;
                const uint32_t first = hs[0];
                __syncwarp();
                if (!hit_l && r + period >= n_ins) hs[half + (rover - half + r) % period] = (uint16_t)sym_l;    // the last store to a slot wins
                rover = half + (rover - half + n_ins) % period;
                sel = hit_l ? first : sym_l;
            } else if (!is_video && hist_size == 64u && nv == (uint32_t)kRound &&
                       __ballot_sync(0xFFFFFFFFu, hit_l && sym_l - num_selectors >= hist_size) == 0u) {
                // Full round with history references (none out of range, mod.rs:409), the usual 64-entry buffer.  Only the
                // references are serial: an insert's slot follows from the number of inserts before it, 32 inserts never meet in
                // the rover's 32 slots, and no reference reads a slot before the inserts ahead of it have landed if the inserts
                // are stored segment by segment -- the segment in front of reference j just before its loads.  A lone warp pays
                // 5-6 cycles per instruction whatever it is, so the loop runs over the references only (compacted through
                // shared memory) and holds nine instructions: a reference is "swap entries k and k / 2".
                uint16_t* hs = W.hist;
                const uint32_t k_l = sym_l - num_selectors;
                const uint32_t any_k1 = __ballot_sync(0xFFFFFFFFu, hit_l && k_l == 1u);          // only index 1 changes entry 0
                const bool ser_l = hit_l && (k_l != 0u || any_k1 != 0u);
                const uint32_t ser = __ballot_sync(0xFFFFFFFFu, ser_l), ins = __ballot_sync(0xFFFFFFFFu, !hit_l);
                const uint32_t lt = (1u << lane) - 1u;
                const uint32_t n_ser = (uint32_t)__popc(ser), seg_l = (uint32_t)__popc(ser & lt);
                const uint32_t n_ins = (uint32_t)__popc(ins), slot_l = 32u + ((rover - 32u + (uint32_t)__popc(ins & lt)) & 31u);
                const uint32_t hs_base = smem_u32(hs);
                const uint32_t first = hs[0];
                if (ser_l) W.hdesc[seg_l] = make_uint2(hs_base + 2u * k_l, hs_base + 2u * (k_l >> 1));
                __syncwarp();
                const uint32_t my_slot = hs_base + 2u * slot_l, hres = smem_u32(W.hres);
#pragma unroll 4
                for (uint32_t j = 0; j < n_ser; j++) {
                    const uint2 dsc = W.hdesc[j];
                    asm volatile("{\n.reg .pred p;\n.reg .u32 v, t;\nsetp.eq.u32 p, %3, %4;\n@p st.shared.u16 [%5], %6;\n"
                                 "ld.shared.u16 v, [%0];\nld.shared.u16 t, [%1];\nst.shared.u16 [%0], t;\nst.shared.u16 [%1], v;\nst.shared.u16 [%2], v;\n}\n"
                                 ::"r"(dsc.x), "r"(dsc.y), "r"(hres + 2u * j), "r"(hit_l ? 0xFFFFFFFFu : seg_l), "r"(j), "r"(my_slot), "r"(sym_l) : "memory");
                }
                if (!hit_l && seg_l == n_ser) hs[slot_l] = (uint16_t)sym_l;
                __syncwarp();
                rover = 32u + ((rover - 32u + n_ins) & 31u);
                sel = ser_l ? W.hres[seg_l] : hit_l ? first : sym_l;
            } else
            for (uint32_t b = 0; b < nv; b++) {
                const uint32_t a = __shfl_sync(0xFFFFFFFFu, tok.x, b);
                const uint32_t sym = a & 0xFFFFu;
                uint32_t s = 0u;
                if (a & kTokSkipSel) s = 0u;                                         // quirk C-5: previous-frame state is always zero
                else if (sym >= num_selectors) {
                    const uint32_t k = sym - num_selectors;
                    if (hist_size == 0u || k >= hist_size) {                         // asserts mod.rs:404,409
                        if ((uint32_t)lane == b) { const uint32_t k5 = ((b << 3 | PH_HISTORY) << 8) | ETC1S_ERR_PREDICTION; key = k5 < key ? k5 : key; }
                        break;
                    }
                    s = hist[k];
                    K2T_DECL(n_hit++;)
                    if (k != 0u) { const uint16_t t = hist[k >> 1]; hist[k >> 1] = (uint16_t)s; hist[k] = t; }     // every lane stores the same values
                } else {
                    s = sym;
                    if (hist_size > 0u) { hist[rover] = (uint16_t)sym; rover++; if (rover == hist_size) rover = hist_size / 2; }
                }
                if ((uint32_t)lane == b) sel = s;
            }
            K2T_DECL(const long long s3 = clock64(); t_seg[2] += s3 - s2;)
            if (live && (ep >= num_endpoints || sel >= num_selectors)) {             // asserts mod.rs:443-444
                const uint32_t k6 = (((uint32_t)lane << 3 | PH_RANGE) << 8) | ETC1S_ERR_RANGE;
                key = k6 < key ? k6 : key;
            }
            if (__any_sync(0xFFFFFFFFu, key != 0xFFFFFFFFu)) {
                const uint32_t first = __reduce_min_sync(0xFFFFFFFFu, key);
                if (lane == 0) { *status = first & 0xFFu; *reinterpret_cast<volatile uint32_t*>(&W.abort) = 1u; }
                return;
            }
            __syncwarp();                                                            // everyone has read the old row before it is overwritten
            if ((uint32_t)lane < nb) {
                out[(uint64_t)y * nbx + x] = ep | (sel << 16);
                rowep[x] = (uint16_t)ep;
            }
            __syncwarp();
            K2T_DECL(t_seg[3] += clock64() - s3;)
        }
    }
    if (lane == 0) *status = 0u;
    K2T(8, t_seg[0]); K2T(9, t_seg[1]); K2T(10, t_seg[2]); K2T(11, t_seg[3]);
    K2T(4, clock64() - tt0); K2T(5, t_wait); K2T(6, n_hit); K2T(7, n_serial);
}

// `pipes` slice pipelines per CTA: warps [0, pipes) tokenize, warps [pipes, 2 pipes) resolve, and with `helpers` > 0 (wide
// pipelines) warps [2 pipes, (2 + helpers) pipes) fill the speculation rings.  The per-slice row state lives in shared memory
// when the slice is at most `row_cap` blocks wide, else in the global scratch area.
__global__ void __launch_bounds__(512) etc1s_entropy_decode_kernel(const __grid_constant__ Etc1sDecodeParams P, uint32_t pipes, uint32_t row_cap, uint32_t helpers)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint32_t* l1s = reinterpret_cast<uint32_t*>(smem_raw);
    const uint32_t l1_words = P.l1_ofs[4];
    PipeShared* pipe_all = reinterpret_cast<PipeShared*>(smem_raw + (((size_t)l1_words * 4 + 15) & ~(size_t)15));
    SpecShared* spec_all = reinterpret_cast<SpecShared*>(pipe_all + pipes);
    unsigned char* rows_all = helpers ? reinterpret_cast<unsigned char*>(spec_all + pipes) : reinterpret_cast<unsigned char*>(pipe_all + pipes);
    for (uint32_t i = threadIdx.x; i < l1_words; i += blockDim.x) l1s[i] = P.l1[i];
    if (threadIdx.x < pipes) {
        PipeShared& W = pipe_all[threadIdx.x];
        for (int r = 0; r < kTokRounds; r++) { mbar_init(&W.bar_full[r], 1); mbar_init(&W.bar_empty[r], 1); }
        W.abort = 0u; W.taken = 0u;
        if (helpers) {
            SpecShared& S = spec_all[threadIdx.x];
            for (int i = 0; i < kSpecChunks; i++) S.done[i] = 0u;
            S.consumed = 0u; S.raw_loaded = 0u; S.stop = 0u;
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const uint32_t warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t pipe = warp % pipes, role = warp / pipes;
    const uint32_t slice = blockIdx.x * pipes + pipe;
    if (slice >= P.num_slices) return;
    PipeShared& W = pipe_all[pipe];
    const Etc1sSliceJob job = P.jobs[slice];
    const size_t row_bytes = etc1s_row_state_bytes(row_cap);
    unsigned char* rows = job.nbx <= row_cap ? rows_all + (size_t)pipe * row_bytes : P.scratch + job.scratch_ofs;
    uint16_t* rowep = reinterpret_cast<uint16_t*>(rows);
    unsigned long long* predrow = reinterpret_cast<unsigned long long*>(rows + ((((size_t)job.nbx + 7u) & ~(size_t)7u) * 2u));
    if (job.nbx <= row_cap) predrow = reinterpret_cast<unsigned long long*>(rows + ((((size_t)row_cap + 7u) & ~(size_t)7u) * 2u));
    uint16_t* hist = P.hist_size <= 64u ? W.hist : reinterpret_cast<uint16_t*>(P.scratch + job.scratch_ofs + etc1s_row_state_bytes(job.nbx));
    if (role == 0u) {
        if (helpers) {
            SpecShared& S = spec_all[pipe];
            etc1s_tokenize_wide(P, job, W, S, l1s, predrow, lane, slice);
            __syncwarp();
            if (lane == 0) *reinterpret_cast<volatile uint32_t*>(&S.stop) = 1u;     // on every way out: the helpers poll it
        } else etc1s_tokenize(P, job, W, l1s, predrow, lane, slice);
    } else if (role == 1u) etc1s_resolve(P, job, W, rowep, hist, P.out_idx + job.out_ofs, P.status + slice, lane, slice);
    else etc1s_speculate(P, W, spec_all[pipe], l1s, role - 2u, helpers, lane);
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

// K3c (EXTENSION, not in the reference -- the definition is oracle/basisu_oracle_etc1s.inc etc1s_emit_bc1): ETC1S -> BC1.
// The selector word (one byte per row, 2 bits per x) already has BC1's index layout, so the per-texel step is a 4-entry
// remap of 2-bit fields done on the whole word at once.
__global__ void __launch_bounds__(256) etc1s_gather_bc1_kernel(const uint32_t* __restrict__ idx, uint64_t nblocks,
                                                               const uint32_t* __restrict__ endpoints, const uint32_t* __restrict__ sel_plain,
                                                               uint2* __restrict__ out)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nblocks; i += stride) {
        const uint32_t v = idx[i];
        const uint32_t e = __ldg(endpoints + (v & 0xFFFFu));                       // inten | r5 << 8 | g5 << 16 | b5 << 24
        const uint32_t rows = __ldg(sel_plain + (v >> 16));
        const uint32_t inten = e & 7u;
        int C[4][3];
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const uint32_t c5 = (e >> (8 + 8 * c)) & 0xFFu;
            const int base = (int)(((c5 << 3) | (c5 >> 2)) & 0xFFu);               // etc.rs:396-406
#pragma unroll
            for (int k = 0; k < 4; k++) { int t = base + __ldg(&kEtc1Mod[inten * 4 + k]); C[k][c] = t < 0 ? 0 : t > 255 ? 255 : t; }
        }
        // fields equal to k, as a 01 pattern per 2-bit field
        const uint32_t L = rows & 0x55555555u, H = (rows >> 1) & 0x55555555u;
        const uint32_t eq[4] = {~H & ~L & 0x55555555u, ~H & L, H & ~L, H & L};
        const int lo = eq[0] ? 0 : eq[1] ? 1 : eq[2] ? 2 : 3;
        const int hi = eq[3] ? 3 : eq[2] ? 2 : eq[1] ? 1 : 0;
        uint32_t w[2];
#pragma unroll
        for (int t = 0; t < 2; t++) {
            const int k = t ? hi : lo;
            int r = C[0][0], g = C[0][1], b = C[0][2];
#pragma unroll
            for (int q = 1; q < 4; q++) if (k == q) { r = C[q][0]; g = C[q][1]; b = C[q][2]; }
            w[t] = (((uint32_t)r * 31u + 127u) / 255u) << 11 | (((uint32_t)g * 63u + 127u) / 255u) << 5 | (((uint32_t)b * 31u + 127u) / 255u);
        }
        const uint32_t c0 = w[0] > w[1] ? w[0] : w[1], c1 = w[0] > w[1] ? w[1] : w[0];
        uint32_t bits = 0u;
        if (c0 != c1) {
            int pal[4][3];
#pragma unroll
            for (int t = 0; t < 2; t++) {
                const uint32_t u = t ? c1 : c0, r5 = u >> 11, g6 = (u >> 5) & 63u, b5 = u & 31u;
                pal[t][0] = (int)((r5 << 3) | (r5 >> 2)); pal[t][1] = (int)((g6 << 2) | (g6 >> 4)); pal[t][2] = (int)((b5 << 3) | (b5 >> 2));
            }
#pragma unroll
            for (int c = 0; c < 3; c++) { pal[2][c] = (2 * pal[0][c] + pal[1][c]) / 3; pal[3][c] = (pal[0][c] + 2 * pal[1][c]) / 3; }
#pragma unroll
            for (int k = 0; k < 4; k++) {
                int best = 0x7fffffff;
                uint32_t m = 0;
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    int dist = 0;
#pragma unroll
                    for (int c = 0; c < 3; c++) { const int t = C[k][c] - pal[j][c]; dist += t * t; }
                    if (dist < best) { best = dist; m = (uint32_t)j; }
                }
                bits |= eq[k] * m;                                                 // m <= 3: no carry out of a 2-bit field
            }
        }
        out[i] = make_uint2(c0 | (c1 << 16), bits);
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

// shared memory of K2: first-level tables + per pipeline {PipeShared, previous-row state for row_cap blocks}
static size_t etc1s_decode_smem_bytes(uint32_t l1_words, int pipes, uint32_t row_cap, int helpers)
{
    return (((size_t)l1_words * 4 + 15) & ~(size_t)15) + (size_t)pipes * (sizeof(PipeShared) + (helpers ? sizeof(SpecShared) : 0) + etc1s_row_state_bytes(row_cap));
}

constexpr size_t kK2SmemLimit = 224 * 1024;

Etc1sDecodePlan plan_etc1s_decode(uint32_t num_slices, uint32_t max_nbx, int sm_count, const uint32_t l1_words[kEtc1sTableSets], bool is_video)
{
    Etc1sDecodePlan plan;
    int pipes = (int)((num_slices + (uint32_t)sm_count - 1u) / (uint32_t)(sm_count > 0 ? sm_count : 1));
    if (pipes < 1) pipes = 1;
    if (pipes > 8) pipes = 8;                                 // 512 threads: the tokenizer wants more than 64 registers
    uint32_t row_cap = (max_nbx + 31u) & ~31u;
    // wide pipelines (helper warps + speculation ring) when there is at most one slice per SM (two rings do not fit beside
    // the tables); texture video has no fast path
    int helpers = (pipes == 1 && !is_video) ? kEtc1sHelpers : 0;
    if (helpers && etc1s_decode_smem_bytes(l1_words[0], pipes, row_cap, helpers) > kK2SmemLimit) helpers = 0;
    // the largest table set that fits beside `pipes` pipelines with their row state in shared memory; failing that, the
    // smallest set with the row state in global scratch and as many pipelines as fit
    int set = kEtc1sTableSets - 1;
    while (set > 0 && etc1s_decode_smem_bytes(l1_words[set], pipes, row_cap, helpers) > kK2SmemLimit) set--;
    if (etc1s_decode_smem_bytes(l1_words[set], 1, row_cap, helpers) > kK2SmemLimit) row_cap = 0;       // very wide slices
    while (pipes > 1 && etc1s_decode_smem_bytes(l1_words[set], pipes, row_cap, helpers) > kK2SmemLimit) pipes--;
    plan.pipes = pipes;
    plan.row_cap = row_cap;
    plan.table_set = set;
    plan.helpers = helpers;
    return plan;
}

cudaError_t launch_etc1s_decode(const Etc1sDecodeParams& P, const Etc1sDecodePlan& plan, cudaStream_t stream)
{
    if (P.num_slices == 0) return cudaSuccess;
    const size_t smem = etc1s_decode_smem_bytes(P.l1_ofs[4], plan.pipes, plan.row_cap, plan.helpers);
    if (smem > kK2SmemLimit) return cudaErrorInvalidValue;                                     // the host sizes the tables to fit
    static bool configured[16] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 16 && !configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(etc1s_entropy_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kK2SmemLimit);
        if (e != cudaSuccess) return e;
        configured[dev] = true;
    }
    const unsigned grid = (P.num_slices + plan.pipes - 1) / plan.pipes;
    etc1s_entropy_decode_kernel<<<grid, 32 * (2 + plan.helpers) * plan.pipes, smem, stream>>>(P, (uint32_t)plan.pipes, plan.row_cap, (uint32_t)plan.helpers);
    return cudaGetLastError();
}

#ifdef B2BU_K2_TRACE
extern "C" __attribute__((visibility("default"))) int b2bu_debug_k2_trace(unsigned long long* dst, int reset)
{
    if (reset == 1) { static unsigned long long z[64][16]; return (int)cudaMemcpyToSymbol(g_k2trace, z, sizeof z); }
    if (reset == 2) return (int)cudaMemcpyFromSymbol(dst, g_k2trace2, sizeof(unsigned long long) * 64 * 2);
    return (int)cudaMemcpyFromSymbol(dst, g_k2trace, sizeof(unsigned long long) * 64 * 16);
}
#endif

cudaError_t launch_etc1s_gather_etc1(const uint32_t* idx, uint64_t nblocks, const uint32_t* endpoints, const uint32_t* sel_etc1, void* out,
                                     int sm_count, cudaStream_t stream)
{
    if (nblocks == 0) return cudaSuccess;
    const uint64_t want = (nblocks + 255) / 256, cap = (uint64_t)sm_count * 16;
    etc1s_gather_etc1_kernel<<<(unsigned)(want < cap ? want : cap), 256, 0, stream>>>(idx, nblocks, endpoints, sel_etc1, reinterpret_cast<uint2*>(out));
    return cudaGetLastError();
}

cudaError_t launch_etc1s_gather_bc1(const uint32_t* idx, uint64_t nblocks, const uint32_t* endpoints, const uint32_t* sel_plain, void* out,
                                    int sm_count, cudaStream_t stream)
{
    if (nblocks == 0) return cudaSuccess;
    const uint64_t want = (nblocks + 255) / 256, cap = (uint64_t)sm_count * 16;
    etc1s_gather_bc1_kernel<<<(unsigned)(want < cap ? want : cap), 256, 0, stream>>>(idx, nblocks, endpoints, sel_plain, reinterpret_cast<uint2*>(out));
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
