// Structures shared by the ETC1S host layer (etc1s_host.cu) and kernels (etc1s_kernels.cu).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace b2bu {

constexpr uint32_t kL1Special = 32u;

enum { ETC1S_OK = 0, ETC1S_ERR_HUFFMAN = 4, ETC1S_ERR_PREDICTION = 6, ETC1S_ERR_VLC = 7, ETC1S_ERR_RANGE = 8 };   // == b2bu_status

struct Etc1sSliceJob {
    uint64_t data_ofs;      // byte offset of the slice bitstream inside Etc1sDecodeParams::data
    uint64_t data_len;
    uint64_t out_ofs;       // first block of this slice inside out_idx
    uint64_t scratch_ofs;   // per-slice scratch: previous-row state (etc1s_row_state_bytes) + large history buffers
    uint32_t nbx, nby;
};

struct Etc1sDecodeParams {
    const uint8_t* data;
    const Etc1sSliceJob* jobs;
    uint32_t num_slices;
    uint32_t* out_idx;              // per block: endpoint_index | selector_index << 16
    uint8_t* scratch;
    // four first-level tables back to back (copied to shared memory): table t has 1 << l1_bits[t] entries starting at word
    // l1_ofs[t] (l1_ofs[4] = total words).  Entry = symbol << 8 | special << 5 | code size; `special` (kL1Special) marks
    // everything the fast path does not handle: codes longer than l1_bits (size 0 -> look in `flat`), slots without a
    // code, and the run symbols (endpoint-pred repeat 256, selector-history RLE).
    const uint32_t* l1;
    uint32_t l1_bits[4];
    uint32_t l1_ofs[5];
    const uint32_t* flat[4];        // full flat tables (huffman.rs:151): symbol << 5 | code size, 1 << max_len entries
    uint32_t max_len[4];
    uint32_t num_endpoints, num_selectors, hist_size, is_video;
    uint32_t* status;               // per slice: 0 or an ETC1S_ERR_* code
};

// bytes of per-slice row state in the scratch area: u16 endpoint index per block + u8 predictor bits per 2 blocks (16-byte multiples)
__host__ __device__ inline uint64_t etc1s_row_state_bytes(uint32_t nbx)
{
    return (((uint64_t)nbx + 7u) & ~7ull) * 2u + ((((uint64_t)nbx + 1u) / 2u + 15u) & ~15ull);
}

// Launch shape of K2.  pipes: slice pipelines (one tokenizer warp + one resolver warp each) per CTA -- the slices are spread
// over the SMs first and packed only when there are more slices than SMs; row_cap: widest slice whose previous-row state
// fits in shared memory (wider slices keep it in the scratch area); table_set: the host keeps kEtc1sTableSets sets of
// first-level tables of growing size (see etc1s_host.cu) and the launch takes the largest one that fits beside the pipelines.
// helpers: with at most one slice per SM every pipeline gets kEtc1sHelpers more warps that decode all bit positions ahead
// of the tokenizer (the speculation ring, etc1s_kernels.cu); 0 = the two-warp pipeline.
constexpr int kEtc1sTableSets = 3;
#ifndef B2BU_K2_HELPERS
#define B2BU_K2_HELPERS 6
#endif
constexpr int kEtc1sHelpers = B2BU_K2_HELPERS;
struct Etc1sDecodePlan { int pipes; uint32_t row_cap; int table_set; int helpers; };
Etc1sDecodePlan plan_etc1s_decode(uint32_t num_slices, uint32_t max_nbx, int sm_count, const uint32_t l1_words[kEtc1sTableSets], bool is_video);
cudaError_t launch_etc1s_decode(const Etc1sDecodeParams& P, const Etc1sDecodePlan& plan, cudaStream_t stream);
cudaError_t launch_etc1s_gather_etc1(const uint32_t* idx, uint64_t nblocks, const uint32_t* endpoints, const uint32_t* sel_etc1, void* out,
                                     int sm_count, cudaStream_t stream);
cudaError_t launch_etc1s_gather_bc1(const uint32_t* idx, uint64_t nblocks, const uint32_t* endpoints, const uint32_t* sel_plain, void* out,
                                    int sm_count, cudaStream_t stream);
cudaError_t launch_etc1s_gather_rgba(const uint32_t* idx_rgb, const uint32_t* idx_alpha, uint32_t nbx, uint64_t nblocks, const uint32_t* endpoints,
                                     const uint32_t* sel_plain, void* out, int sm_count, cudaStream_t stream);

}  // namespace b2bu
