// ETC1S / BasisLZ host side of the C ABI.
//
// Once per file (basis_lz::Decoder::new, reference src/basis_lz/mod.rs:64-95): decode the endpoint
// and selector codebooks and the four slice Huffman models on the host, exactly as the reference
// does (they are small and serial), then upload them: codebooks as flat u32 arrays, each Huffman
// model as a 10-bit first-level table for shared memory plus the reference's full flat table.
// Per call: upload the slice bitstreams, run K2 (one pipeline of warps per slice) and K3 (gather), copy back.
#include <algorithm>
#include <cstring>
#include <memory>
#include <vector>

#include "../../include/b2bu.h"
#include "etc1s_device.h"
#include "etc1s_host.h"
#include "host_internal.h"

namespace b2bu {

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return cuda_fail(e_, #call); } while (0)

// Shared memory given to the four first-level Huffman tables of K2 (the rest holds the per-slice pipelines).  Three sets
// are kept per file: a wide one for launches with one slice per SM (a lone pipeline leaves ~210 KB free), a medium one that
// still fits beside 8 pipelines of ordinary width (codes that miss the first level cost a global-memory lookup and a redone
// pair: 592 slices of 512x512 blocks went from 59 to 34 ms with it), and a narrow one for very wide slices.
constexpr size_t kL1BudgetBytes[kEtc1sTableSets] = {96 * 1024, 160 * 1024, 208 * 1024};

// ---- bit cursor: LSB first, bytes past the end read as zero (src/bitreader.rs:27-60) ----------
struct BitCursor {
    const uint8_t* p; size_t len; uint64_t pos = 0;
    BitCursor(const uint8_t* d, size_t n) : p(d), len(n) {}
    uint32_t peek(unsigned n) const
    {
        uint64_t v = 0;
        const size_t byte = (size_t)(pos >> 3);
        for (int i = 0; i < 8; i++) if (byte + i < len) v |= (uint64_t)p[byte + i] << (8 * i);
        v >>= (pos & 7);
        return n >= 32 ? (uint32_t)v : (uint32_t)(v & ((1ull << n) - 1));
    }
    uint32_t read(unsigned n) { const uint32_t v = peek(n); pos += n; return v; }
};

// ---- Huffman model (src/basis_lz/huffman.rs) ---------------------------------------------------
struct HuffModel {
    std::vector<uint32_t> flat;     // 1 << max_len entries of symbol << 5 | code size (0 = no code), as huffman.rs:151-170 fills them
    unsigned max_len = 0;
    uint32_t count[17] = {0};       // codes per length (count[0] unused)
};

static uint32_t bit_reverse32(uint32_t x)
{
    x = (x >> 16) | (x << 16);
    x = ((x & 0xFF00FF00u) >> 8) | ((x & 0x00FF00FFu) << 8);
    x = ((x & 0xF0F0F0F0u) >> 4) | ((x & 0x0F0F0F0Fu) << 4);
    x = ((x & 0xCCCCCCCCu) >> 2) | ((x & 0x33333333u) << 2);
    return ((x & 0xAAAAAAAAu) >> 1) | ((x & 0x55555555u) << 1);
}

// huffman.rs:133-184 from_sizes: canonical codes in ascending symbol order, bit-reversed, every
// table slot whose low `size` bits equal the code is filled (later symbols overwrite earlier ones)
static int huff_from_sizes(const std::vector<uint8_t>& sizes, HuffModel& m)
{
    uint32_t count[17] = {0}, next[17] = {0};
    unsigned max_len = 0;
    for (uint8_t s : sizes) {
        if (s > 16) return B2BU_ERR_HUFFMAN;
        count[s]++;
        if (s > max_len) max_len = s;
    }
    count[0] = 0;
    uint32_t total = 0;
    for (unsigned b = 1; b <= 16; b++) { total = (total + count[b - 1]) << 1; next[b] = total; }
    m.max_len = max_len;
    m.flat.assign((size_t)1 << max_len, 0u);
    for (size_t sym = 0; sym < sizes.size(); sym++) {
        const unsigned size = sizes[sym];
        if (!size) continue;
        const uint32_t code = (bit_reverse32(next[size]) >> (32 - size)) & 0xFFFFu;
        const uint32_t variants = (1u << (max_len - size)) & 0xFFFFu;              // u16 in the reference
        for (uint32_t fill = 0; fill < variants; fill++) {
            const size_t id = (size_t)((((fill << size) & 0xFFFFu) | code));
            if (id >= m.flat.size()) return B2BU_ERR_HUFFMAN;                       // the reference would panic (index out of bounds)
            m.flat[id] = ((uint32_t)sym << 5) | size;
        }
        next[size]++;
    }
    for (unsigned b = 0; b <= 16; b++) if (next[b] > 65536u) return B2BU_ERR_HUFFMAN;   // "codes don't fit into 16 bits"
    for (unsigned b = 1; b <= 16; b++) m.count[b] = count[b];
    return B2BU_OK;
}

static int huff_decode_host(const HuffModel& m, BitCursor& c, uint32_t& sym)       // huffman.rs:186-198
{
    const uint32_t e = m.flat[c.peek(m.max_len)];
    if ((e & 31u) == 0) return B2BU_ERR_HUFFMAN;
    c.pos += e & 31u;
    sym = e >> 5;
    return B2BU_OK;
}

// huffman.rs:43-118 read_huffman_table
static int read_huffman_table(BitCursor& c, HuffModel& out)
{
    static const uint8_t order[21] = {17, 18, 19, 20, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15, 16};
    const size_t total_used_syms = c.read(14);
    const unsigned num_cl = c.read(5);
    if (num_cl > 21) return B2BU_ERR_HUFFMAN;                                      // the reference would panic
    std::vector<uint8_t> cl_sizes(21, 0);
    for (unsigned i = 0; i < num_cl; i++) cl_sizes[order[i]] = (uint8_t)c.read(3);
    HuffModel cl;
    int st = huff_from_sizes(cl_sizes, cl);
    if (st) return st;
    std::vector<uint8_t> sizes;
    sizes.reserve(total_used_syms + 140);
    while (sizes.size() < total_used_syms) {
        uint32_t s;
        if ((st = huff_decode_host(cl, c, s))) return st;
        if (s <= 16) sizes.push_back((uint8_t)s);
        else if (s == 17) sizes.insert(sizes.end(), 3 + c.read(3), 0);
        else if (s == 18) sizes.insert(sizes.end(), 11 + c.read(7), 0);
        else {
            if (sizes.empty() || sizes.back() == 0) return B2BU_ERR_HUFFMAN;       // huffman.rs:82-91, :98-107
            const uint8_t prev = sizes.back();
            const size_t n = s == 19 ? 3 + c.read(2) : 7 + c.read(7);
            sizes.insert(sizes.end(), n, prev);
        }
    }
    return huff_from_sizes(sizes, out);
}

// First-level table of `bits` bits for shared memory (entry format: etc1s_device.h).  A slot is directly usable iff every
// flat slot that shares its low `bits` bits holds the same code of at most `bits` bits; everything else, and the symbol
// `run_sym` (a run marker the fast path does not handle), is flagged special.
static void build_l1(const HuffModel& m, unsigned bits, uint32_t run_sym, uint32_t* l1)
{
    for (uint32_t i = 0; i < (1u << bits); i++) {
        uint32_t f;
        bool ok = true;
        if (m.max_len <= bits) f = m.flat[i & ((1u << m.max_len) - 1u)];
        else {
            f = m.flat[i];
            for (uint32_t k = 1; k < (1u << (m.max_len - bits)) && ok; k++) ok = m.flat[i | (k << bits)] == f;
            ok = ok && (f & 31u) <= bits;
        }
        if (!ok || (f & 31u) == 0u) { l1[i] = kL1Special; continue; }               // long code or no code: size 0
        const uint32_t sym = f >> 5;
        l1[i] = (sym << 8) | (sym == run_sym ? kL1Special : 0u) | (f & 31u);
    }
}

// First-level widths for the four slice models under a shared-memory budget: start at <= 10 bits and keep widening the
// table whose long codes cost most (Kraft mass of the codes that do not fit x how often the model is read per block).
static void choose_l1_bits(const HuffModel* models, size_t budget_bytes, unsigned* bits)
{
    static const double weight[4] = {0.25, 0.5, 1.0, 0.02};     // predictor symbol per 2x2 group, delta, selector, run length
    auto mass = [&](int t, unsigned L) { double s = 0; for (unsigned b = L + 1; b <= 16; b++) s += models[t].count[b] / double(1u << b); return s * weight[t]; };
    size_t used = 0;
    for (int t = 0; t < 4; t++) { bits[t] = std::min(models[t].max_len, 10u); used += (size_t)4 << bits[t]; }
    for (;;) {
        // widen the table with the largest remaining cost (not the largest marginal gain: a table whose codes are all
        // 12 bits long gains nothing from the step 10 -> 11)
        int best = -1;
        double best_cost = 0;
        for (int t = 0; t < 4; t++) {
            if (bits[t] >= models[t].max_len || bits[t] >= 15u || used + ((size_t)4 << bits[t]) > budget_bytes) continue;
            const double cost = mass(t, bits[t]);
            if (cost > best_cost) { best = t; best_cost = cost; }
        }
        if (best < 0) break;
        used += (size_t)4 << bits[best];
        bits[best]++;
    }
}

}  // namespace b2bu

using namespace b2bu;

struct b2bu_etc1s {
    int device = 0;
    uint32_t num_endpoints = 0, num_selectors = 0, hist_size = 0;
    bool is_video = false;
    uint32_t max_len[4] = {0, 0, 0, 0};
    // device copies
    uint32_t* d_endpoints = nullptr;     // inten | r5 << 8 | g5 << 16 | b5 << 24
    uint32_t* d_sel_plain = nullptr;     // 4 rows, 2 bits per x          (etc.rs:343-361)
    uint32_t* d_sel_etc1 = nullptr;      // ETC1 bit planes               (etc.rs:363-393)
    // narrow / medium / wide set of the four first-level tables, each set back to back
    uint32_t* d_l1[kEtc1sTableSets] = {};
    uint32_t l1_bits[kEtc1sTableSets][4] = {}, l1_ofs[kEtc1sTableSets][5] = {};
    int last_set = 0;                    // table set used by the last call
    uint32_t* d_flat[4] = {nullptr, nullptr, nullptr, nullptr};
    // per-call scratch (grow only)
    void* d_data = nullptr; size_t data_cap = 0;
    void* d_idx = nullptr; size_t idx_cap = 0;
    void* d_out = nullptr; size_t out_cap = 0;
    void* d_scratch = nullptr; size_t scratch_cap = 0;
    void* d_jobs = nullptr; size_t jobs_cap = 0;
    void* d_status = nullptr; size_t status_cap = 0;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};   // phase boundaries of the last call: H2D | K2 | K3 | D2H
    float last_ms[3] = {0.f, 0.f, 0.f};
    uint64_t last_blocks = 0, last_in_bytes = 0, last_out_bytes = 0;
    std::mutex mu;
};

namespace b2bu {

// etc.rs:363-393 Selector::set_selector for a whole selector (4 row bytes) -> ETC1 bit planes
static uint32_t selector_etc1_bytes(const uint8_t rows[4])
{
    static const uint8_t to_etc1[4] = {3, 2, 0, 1};
    uint8_t b[4] = {0, 0, 0, 0};
    for (unsigned y = 0; y < 4; y++)
        for (unsigned x = 0; x < 4; x++) {
            const unsigned mod = to_etc1[(rows[y] >> (2 * x)) & 3];
            const unsigned pixel = x * 4 + y, ms = 1 - pixel / 8, ls = ms + 2, bit = pixel % 8;
            b[ls] |= (uint8_t)((mod & 1) << bit);
            b[ms] |= (uint8_t)((mod >> 1) << bit);
        }
    return (uint32_t)b[0] | ((uint32_t)b[1] << 8) | ((uint32_t)b[2] << 16) | ((uint32_t)b[3] << 24);
}

static int upload(uint32_t** dst, const std::vector<uint32_t>& v)
{
    CK(cudaMalloc(dst, std::max<size_t>(v.size(), 1) * sizeof(uint32_t)));
    if (!v.empty()) CK(cudaMemcpy(*dst, v.data(), v.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
    return B2BU_OK;
}

struct SliceReq { const uint8_t* data; uint64_t len; uint32_t nbx, nby; const uint8_t* d_data; };   // d_data: the same bytes on the device, or null
struct ImageReq { int rgb_slice, alpha_slice; uint64_t out_ofs; };      // indices into the slice list; alpha_slice = -1 if none

// Decodes `slices` with K2 and emits one output per image with K3.  target: B2BU_ETC1 (images = slices)
// or B2BU_RGBA (an image may combine an rgb and an alpha slice).
static int etc1s_run(b2bu_etc1s* h, int target, const std::vector<SliceReq>& slices, const std::vector<ImageReq>& images, uint8_t* out,
                     uint64_t out_total)
{
    DeviceCtx* c;
    int st = get_ctx(&c);
    if (st) return st;
    // the handle's codebooks and scratch live on the device it was opened on: streams of another device's context cannot run on them
    if (c->device != h->device) return B2BU_ERR_ARGUMENT;
    std::lock_guard<std::mutex> lk(h->mu);
    const size_t ns = slices.size();
    if (ns == 0) return B2BU_OK;
    std::vector<Etc1sSliceJob> jobs(ns);
    uint64_t data_total = 0, blocks_total = 0, scratch_total = 0;
    uint32_t max_nbx = 0;
    for (size_t i = 0; i < ns; i++) {
        jobs[i].data_ofs = data_total; jobs[i].data_len = slices[i].len;
        jobs[i].out_ofs = blocks_total; jobs[i].scratch_ofs = scratch_total;
        jobs[i].nbx = slices[i].nbx; jobs[i].nby = slices[i].nby;
        data_total += (slices[i].len + 15) & ~15ull;
        blocks_total += (uint64_t)slices[i].nbx * slices[i].nby;
        scratch_total += etc1s_row_state_bytes(slices[i].nbx) + (h->hist_size > 64 ? ((2ull * h->hist_size + 15) & ~15ull) : 0);
        max_nbx = std::max(max_nbx, slices[i].nbx);
    }
    if ((st = ensure(&h->d_data, &h->data_cap, data_total + 16))) return st;
    if ((st = ensure(&h->d_idx, &h->idx_cap, blocks_total * 4 + 16))) return st;
    if ((st = ensure(&h->d_out, &h->out_cap, out_total + 16))) return st;
    if ((st = ensure(&h->d_scratch, &h->scratch_cap, scratch_total + 16))) return st;
    if ((st = ensure(&h->d_jobs, &h->jobs_cap, ns * sizeof(Etc1sSliceJob)))) return st;
    if ((st = ensure(&h->d_status, &h->status_cap, ns * 4))) return st;
    cudaStream_t s = c->streams[0];
    for (size_t i = 0; i < ns; i++)
        if (slices[i].len) {
            if (slices[i].d_data) CK(cudaMemcpyAsync(static_cast<uint8_t*>(h->d_data) + jobs[i].data_ofs, slices[i].d_data, slices[i].len, cudaMemcpyDeviceToDevice, s));
            else CK(cudaMemcpyAsync(static_cast<uint8_t*>(h->d_data) + jobs[i].data_ofs, slices[i].data, slices[i].len, cudaMemcpyHostToDevice, s));
        }
    CK(cudaMemcpyAsync(h->d_jobs, jobs.data(), ns * sizeof(Etc1sSliceJob), cudaMemcpyHostToDevice, s));

    for (int i = 0; i < 4; i++) if (!h->ev[i]) CK(cudaEventCreate(&h->ev[i]));
    CK(cudaEventRecord(h->ev[0], s));

    Etc1sDecodeParams P;
    P.data = static_cast<const uint8_t*>(h->d_data);
    P.jobs = static_cast<const Etc1sSliceJob*>(h->d_jobs);
    P.num_slices = (uint32_t)ns;
    P.out_idx = static_cast<uint32_t*>(h->d_idx);
    P.scratch = static_cast<uint8_t*>(h->d_scratch);
    uint32_t l1_words[kEtc1sTableSets];
    for (int k = 0; k < kEtc1sTableSets; k++) l1_words[k] = h->l1_ofs[k][4];
    const Etc1sDecodePlan dplan = plan_etc1s_decode((uint32_t)ns, max_nbx, c->sm_count, l1_words, h->is_video);
    const int set = dplan.table_set;
    h->last_set = set;
    P.l1 = h->d_l1[set];
    for (int t = 0; t < 4; t++) { P.flat[t] = h->d_flat[t]; P.max_len[t] = h->max_len[t]; P.l1_bits[t] = h->l1_bits[set][t]; P.l1_ofs[t] = h->l1_ofs[set][t]; }
    P.l1_ofs[4] = h->l1_ofs[set][4];
    P.num_endpoints = h->num_endpoints; P.num_selectors = h->num_selectors; P.hist_size = h->hist_size; P.is_video = h->is_video ? 1u : 0u;
    P.status = static_cast<uint32_t*>(h->d_status);
    { NvtxScope nv("b2bu K2 etc1s_entropy_decode"); CK(launch_etc1s_decode(P, dplan, s)); }
    count_launch(1);
    CK(cudaEventRecord(h->ev[1], s));

    // The verdict of every slice is read before anything is gathered: a resolver that stops at the first error leaves the rest
    // of its slice's index plane unwritten (d_idx is grow-only scratch, never cleared), and K3 would index the codebooks with
    // whatever is there.  A failed call therefore launches no gather and leaves the caller's buffer untouched (the reference
    // returns Err and no image, basis_lz/mod.rs:97-186).  K2 runs for milliseconds; the extra synchronisation is noise.
    std::vector<uint32_t> status(ns);
    CK(cudaMemcpyAsync(status.data(), h->d_status, ns * 4, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    for (size_t i = 0; i < ns; i++) if (status[i]) return (int)status[i];          // first failing slice in file order

    const uint32_t* idx = static_cast<const uint32_t*>(h->d_idx);
    NvtxScope nv3("b2bu K3 etc1s gather + D2H");
    if (target == B2BU_ETC1) {
        // images are the slices in order and ETC1 output does not depend on the slice shape: one gather over everything
        CK(launch_etc1s_gather_etc1(idx, blocks_total, h->d_endpoints, h->d_sel_etc1, h->d_out, c->sm_count, s));
        count_launch(1);
    } else if (target == B2BU_BC1) {
        CK(launch_etc1s_gather_bc1(idx, blocks_total, h->d_endpoints, h->d_sel_plain, h->d_out, c->sm_count, s));
        count_launch(1);
    } else {
        for (const ImageReq& im : images) {
            const Etc1sSliceJob& j = jobs[im.rgb_slice];
            const uint32_t* ia = im.alpha_slice >= 0 ? idx + jobs[im.alpha_slice].out_ofs : nullptr;
            CK(launch_etc1s_gather_rgba(idx + j.out_ofs, ia, j.nbx, (uint64_t)j.nbx * j.nby, h->d_endpoints, h->d_sel_plain,
                                        static_cast<uint8_t*>(h->d_out) + im.out_ofs, c->sm_count, s));
            count_launch(1);
        }
    }
    CK(cudaEventRecord(h->ev[2], s));
    CK(cudaMemcpyAsync(out, h->d_out, out_total, cudaMemcpyDeviceToHost, s));
    CK(cudaEventRecord(h->ev[3], s));
    CK(cudaStreamSynchronize(s));
    for (int i = 0; i < 3; i++) cudaEventElapsedTime(&h->last_ms[i], h->ev[i], h->ev[i + 1]);
    h->last_blocks = blocks_total; h->last_in_bytes = data_total; h->last_out_bytes = out_total;
    return B2BU_OK;
}

static int etc1s_open_impl(uint32_t endpoint_count, uint32_t selector_count, const uint8_t* ep, size_t ep_len, const uint8_t* sel, size_t sel_len,
                           const uint8_t* tab, size_t tab_len, bool is_video, b2bu_etc1s** out)
{
    DeviceCtx* c;
    int st = get_ctx(&c);
    if (st) return st;
    struct Closer { void operator()(b2bu_etc1s* p) const { b2bu_etc1s_close(p); } };        // frees whatever was uploaded so far
    std::unique_ptr<b2bu_etc1s, Closer> h(new b2bu_etc1s);
    h->device = c->device;
    h->num_endpoints = endpoint_count; h->num_selectors = selector_count; h->is_video = is_video;

    // ---- endpoint codebook (mod.rs:461-516) ----
    std::vector<uint32_t> endpoints(endpoint_count);
    {
        BitCursor bc(ep, ep_len);
        HuffModel m0, m1, m2, mi;
        if ((st = read_huffman_table(bc, m0)) || (st = read_huffman_table(bc, m1)) || (st = read_huffman_table(bc, m2)) ||
            (st = read_huffman_table(bc, mi))) return st;
        const bool grayscale = bc.read(1) == 1;
        uint32_t prev[3] = {16, 16, 16}, prev_inten = 0;
        for (uint32_t i = 0; i < endpoint_count; i++) {
            uint32_t d;
            if ((st = huff_decode_host(mi, bc, d))) return st;
            const uint32_t inten = (d + prev_inten) & 7u;
            prev_inten = inten;
            uint32_t col[3];
            for (int ch = 0; ch < (grayscale ? 1 : 3); ch++) {
                const HuffModel& m = prev[ch] <= 9 ? m0 : prev[ch] <= 21 ? m1 : m2;
                if ((st = huff_decode_host(m, bc, d))) return st;
                col[ch] = (prev[ch] + (d & 0xFFu)) & 31u;                        // u8 wrapping_add, then & 31
                prev[ch] = col[ch];
            }
            if (grayscale) col[1] = col[2] = col[0];
            endpoints[i] = inten | (col[0] << 8) | (col[1] << 16) | (col[2] << 24);
        }
    }
    // ---- selector codebook (mod.rs:524-583) ----
    std::vector<uint32_t> sel_plain(selector_count), sel_etc1(selector_count);
    {
        BitCursor bc(sel, sel_len);
        const bool global = bc.read(1) == 1, hybrid = bc.read(1) == 1, raw = bc.read(1) == 1;
        if (global || hybrid) return B2BU_ERR_SELECTOR_CB;
        HuffModel m;
        if (!raw && (st = read_huffman_table(bc, m))) return st;
        uint8_t prev[4] = {0, 0, 0, 0};
        for (uint32_t i = 0; i < selector_count; i++) {
            uint8_t rows[4];
            for (int y = 0; y < 4; y++) {
                if (raw || i == 0) rows[y] = (uint8_t)bc.read(8);
                else {
                    uint32_t d;
                    if ((st = huff_decode_host(m, bc, d))) return st;
                    rows[y] = (uint8_t)((uint8_t)d ^ prev[y]);
                }
                prev[y] = rows[y];
            }
            sel_plain[i] = (uint32_t)rows[0] | ((uint32_t)rows[1] << 8) | ((uint32_t)rows[2] << 16) | ((uint32_t)rows[3] << 24);
            sel_etc1[i] = selector_etc1_bytes(rows);
        }
    }
    // ---- slice models (mod.rs:77-83) ----
    HuffModel models[4];
    {
        BitCursor bc(tab, tab_len);
        for (int t = 0; t < 4; t++) if ((st = read_huffman_table(bc, models[t]))) return st;
        h->hist_size = bc.read(13);
    }
    std::vector<uint32_t> l1[kEtc1sTableSets];
    const uint32_t run_sym[4] = {256u, 0xFFFFFFFFu, (h->hist_size + selector_count) & 0xFFFFu, 0xFFFFFFFFu};   // mod.rs:220-222
    for (int set = 0; set < kEtc1sTableSets; set++) {
        unsigned bits[4];
        choose_l1_bits(models, kL1BudgetBytes[set], bits);
        for (int t = 0; t < 4; t++) {
            h->l1_bits[set][t] = bits[t]; h->l1_ofs[set][t] = (uint32_t)l1[set].size(); h->max_len[t] = models[t].max_len;
            l1[set].resize(l1[set].size() + ((size_t)1 << bits[t]));
            build_l1(models[t], bits[t], run_sym[t], l1[set].data() + h->l1_ofs[set][t]);
        }
        h->l1_ofs[set][4] = (uint32_t)l1[set].size();
    }

    CK(cudaSetDevice(h->device));
    if ((st = upload(&h->d_endpoints, endpoints)) || (st = upload(&h->d_sel_plain, sel_plain)) || (st = upload(&h->d_sel_etc1, sel_etc1)) ||
        (st = upload(&h->d_l1[0], l1[0])) || (st = upload(&h->d_l1[1], l1[1])) || (st = upload(&h->d_l1[2], l1[2]))) return st;
    for (int t = 0; t < 4; t++) if ((st = upload(&h->d_flat[t], models[t].flat))) return st;
    *out = h.release();
    return B2BU_OK;
}

int etc1s_read_file(int target, const uint8_t* buf, size_t len, const b2bu_header& hd, const SliceDesc* descs, const b2bu_image* plan,
                    uint32_t nimg, bool pair, uint8_t* out, const uint8_t* d_file)
{
    // basis.rs:262-300 make_basis_lz_decoder: sections are slices of the file (out of range => the reference panics)
    if ((uint64_t)hd.endpoint_cb_file_ofs + hd.endpoint_cb_file_size > len || (uint64_t)hd.selector_cb_file_ofs + hd.selector_cb_file_size > len ||
        (uint64_t)hd.tables_file_ofs + hd.tables_file_size > len || (uint64_t)hd.extended_file_ofs + hd.extended_file_size > len) return B2BU_ERR_RANGE;
    b2bu_etc1s* h = nullptr;
    // quirk C-1 (basis.rs:289-291): total_selectors is passed as BOTH the endpoint and the selector count
    int st = etc1s_open_impl(hd.total_selectors, hd.total_selectors, buf + hd.endpoint_cb_file_ofs, hd.endpoint_cb_file_size,
                             buf + hd.selector_cb_file_ofs, hd.selector_cb_file_size, buf + hd.tables_file_ofs, hd.tables_file_size,
                             hd.tex_type == 3, &h);
    if (st) return st;
    std::vector<SliceReq> slices;
    std::vector<ImageReq> images;
    uint64_t total = 0;
    for (uint32_t i = 0; i < nimg; i++) {
        const SliceDesc& s = descs[pair ? 2 * i : i];
        ImageReq im;
        im.rgb_slice = (int)slices.size();
        slices.push_back({buf + s.file_ofs, s.file_size, s.num_blocks_x, s.num_blocks_y, d_file ? d_file + s.file_ofs : nullptr});
        im.alpha_slice = -1;
        if (pair) {
            const SliceDesc& a = descs[2 * i + 1];
            im.alpha_slice = (int)slices.size();
            slices.push_back({buf + a.file_ofs, a.file_size, a.num_blocks_x, a.num_blocks_y, d_file ? d_file + a.file_ofs : nullptr});
        }
        im.out_ofs = plan[i].offset;
        images.push_back(im);
        total = plan[i].offset + plan[i].nbytes;
    }
    st = etc1s_run(h, target, slices, images, out, total);
    b2bu_etc1s_close(h);
    return st;
}

}  // namespace b2bu

extern "C" {

int b2bu_etc1s_open(uint32_t endpoint_count, uint32_t selector_count, const uint8_t* endpoint_data, size_t endpoint_len,
                    const uint8_t* selector_data, size_t selector_len, const uint8_t* tables_data, size_t tables_len, int is_video,
                    b2bu_etc1s** handle)
{
    if (!handle || !endpoint_data || !selector_data || !tables_data || endpoint_count > 65535 || selector_count > 65535) return B2BU_ERR_ARGUMENT;
    *handle = nullptr;
    return etc1s_open_impl(endpoint_count, selector_count, endpoint_data, endpoint_len, selector_data, selector_len, tables_data, tables_len,
                           is_video != 0, handle);
}

void b2bu_etc1s_close(b2bu_etc1s* h)
{
    if (!h) return;
    cudaSetDevice(h->device);
    for (int i = 0; i < 4; i++) if (h->ev[i]) cudaEventDestroy(h->ev[i]);
    cudaFree(h->d_endpoints); cudaFree(h->d_sel_plain); cudaFree(h->d_sel_etc1); for (int k = 0; k < kEtc1sTableSets; k++) cudaFree(h->d_l1[k]);
    for (int t = 0; t < 4; t++) cudaFree(h->d_flat[t]);
    cudaFree(h->d_data); cudaFree(h->d_idx); cudaFree(h->d_out); cudaFree(h->d_scratch); cudaFree(h->d_jobs); cudaFree(h->d_status);
    delete h;
}

int b2bu_etc1s_last_timing(b2bu_etc1s* h, float* entropy_ms, float* gather_ms, float* d2h_ms, uint64_t* blocks)
{
    if (!h) return B2BU_ERR_ARGUMENT;
    std::lock_guard<std::mutex> lk(h->mu);
    if (entropy_ms) *entropy_ms = h->last_ms[0];
    if (gather_ms) *gather_ms = h->last_ms[1];
    if (d2h_ms) *d2h_ms = h->last_ms[2];
    if (blocks) *blocks = h->last_blocks;
    return B2BU_OK;
}

int b2bu_etc1s_table_info(b2bu_etc1s* h, uint32_t l1_bits[4], uint32_t max_code_len[4])
{
    if (!h) return B2BU_ERR_ARGUMENT;
    for (int t = 0; t < 4; t++) { if (l1_bits) l1_bits[t] = h->l1_bits[h->last_set][t]; if (max_code_len) max_code_len[t] = h->max_len[t]; }
    return B2BU_OK;
}

int b2bu_etc1s_transcode_slices(b2bu_etc1s* h, int target, uint32_t nbx, uint32_t nby, const uint8_t* data, size_t data_len,
                                const uint64_t* slice_ofs, const uint64_t* slice_len, uint32_t num_slices, uint8_t* out, size_t out_bytes)
{
    if (!h || (target != B2BU_ETC1 && target != B2BU_RGBA && target != B2BU_BC1) || (num_slices && (!data || !slice_ofs || !slice_len || !out)))
        return B2BU_ERR_ARGUMENT;
    const uint64_t per = (uint64_t)nbx * nby * (target == B2BU_RGBA ? 64 : 8);
    if (out_bytes < per * num_slices) return B2BU_ERR_ARGUMENT;
    std::vector<SliceReq> slices(num_slices);
    std::vector<ImageReq> images(num_slices);
    for (uint32_t i = 0; i < num_slices; i++) {
        if (slice_ofs[i] + slice_len[i] > data_len) return B2BU_ERR_RANGE;
        slices[i] = {data + slice_ofs[i], slice_len[i], nbx, nby, nullptr};
        images[i] = {(int)i, -1, per * i};
    }
    return etc1s_run(h, target, slices, images, out, per * num_slices);
}

int b2bu_etc1s_transcode_to_etc1(b2bu_etc1s* h, uint32_t nbx, uint32_t nby, const uint8_t* slice, size_t slice_len, uint8_t* out, size_t out_bytes)
{
    const uint64_t ofs = 0, len = slice_len;
    return b2bu_etc1s_transcode_slices(h, B2BU_ETC1, nbx, nby, slice, slice_len, &ofs, &len, 1, out, out_bytes);
}

int b2bu_etc1s_transcode_to_bc1(b2bu_etc1s* h, uint32_t nbx, uint32_t nby, const uint8_t* slice, size_t slice_len, uint8_t* out, size_t out_bytes)
{
    const uint64_t ofs = 0, len = slice_len;
    return b2bu_etc1s_transcode_slices(h, B2BU_BC1, nbx, nby, slice, slice_len, &ofs, &len, 1, out, out_bytes);
}

int b2bu_etc1s_decode_to_rgba(b2bu_etc1s* h, uint32_t nbx, uint32_t nby, const uint8_t* rgb_slice, size_t rgb_len, const uint8_t* alpha_slice,
                              size_t alpha_len, uint8_t* out, size_t out_bytes)
{
    if (!h || !rgb_slice || !out) return B2BU_ERR_ARGUMENT;
    const uint64_t total = (uint64_t)nbx * nby * 64;
    if (out_bytes < total) return B2BU_ERR_ARGUMENT;
    std::vector<SliceReq> slices;
    slices.push_back({rgb_slice, rgb_len, nbx, nby, nullptr});
    if (alpha_slice) slices.push_back({alpha_slice, alpha_len, nbx, nby, nullptr});
    std::vector<ImageReq> images(1);
    images[0] = {0, alpha_slice ? 1 : -1, 0};
    return etc1s_run(h, B2BU_RGBA, slices, images, out, total);
}

}  // extern "C"
