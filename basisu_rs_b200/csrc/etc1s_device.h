// Structures shared by the ETC1S host layer (etc1s_host.cu) and kernels (etc1s_kernels.cu).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace b2bu {

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
    const uint32_t* l1;             // 4 first-level tables of 1024 entries: symbol << 5 | code size (0 = no code, ~0 = long code)
    const uint32_t* flat[4];        // full flat tables (huffman.rs:151), same entry format, 1 << max_len entries
    uint32_t max_len[4];
    // long codes (> 10 bits): canonical description per table, 16 x upper then 16 x base (see HuffModel), and the sorted symbols
    const uint32_t* canon;
    const uint16_t* syms;
    uint32_t sym_ofs[5];
    uint32_t canon_ok;              // bit t set: table t is a valid prefix code and `canon` applies; else the flat table is read
    uint32_t num_endpoints, num_selectors, hist_size, is_video;
    uint32_t* status;               // per slice: 0 or an ETC1S_ERR_* code
};

// bytes of per-slice row state in the scratch area: u16 endpoint index per block + u8 predictor bits per 2 blocks (16-byte multiples)
__host__ __device__ inline uint64_t etc1s_row_state_bytes(uint32_t nbx)
{
    return (((uint64_t)nbx + 7u) & ~7ull) * 2u + ((((uint64_t)nbx + 1u) / 2u + 15u) & ~15ull);
}

cudaError_t launch_etc1s_decode(const Etc1sDecodeParams& P, int warps_per_cta, uint32_t max_nbx, cudaStream_t stream);
cudaError_t launch_etc1s_gather_etc1(const uint32_t* idx, uint64_t nblocks, const uint32_t* endpoints, const uint32_t* sel_etc1, void* out,
                                     int sm_count, cudaStream_t stream);
cudaError_t launch_etc1s_gather_rgba(const uint32_t* idx_rgb, const uint32_t* idx_alpha, uint32_t nbx, uint64_t nblocks, const uint32_t* endpoints,
                                     const uint32_t* sel_plain, void* out, int sm_count, cudaStream_t stream);

}  // namespace b2bu
