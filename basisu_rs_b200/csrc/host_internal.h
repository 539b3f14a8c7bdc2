// Host-runtime internals shared by capi.cu, basis_file.cu and etc1s_host.cu.
#pragma once
#include <cstdint>
#include <mutex>
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>          // header-only: the ranges cost nothing unless a tool (nsys, ncu --nvtx) is attached
#include "../../include/b2bu.h"

namespace b2bu {

constexpr int kMaxDevices = 16;
#ifndef B2BU_STREAMS
#define B2BU_STREAMS 4
#endif
#ifndef B2BU_CHUNK_LOG2
#define B2BU_CHUNK_LOG2 19
#endif
constexpr int kStreams = B2BU_STREAMS;           // buffer slots of the host-pointer pipeline; streams[0..2] = H2D, kernels, D2H
constexpr size_t kChunkBlocks = size_t(1) << B2BU_CHUNK_LOG2; // 16 B << 19 = 8 MiB of UASTC per pipeline stage

enum { ERR_MODE_DEV = 2, ERR_PATTERN_DEV = 3 };  // == ERR_MODE / ERR_PATTERN in uastc_device.cuh

struct DeviceCtx {
    std::mutex init_mu, run_mu;
    bool ready = false;
    int device = 0, sm_count = 0;
    cudaStream_t streams[kStreams] = {};
    void* d_in[kStreams] = {};
    size_t in_cap[kStreams] = {};
    void* d_out[kStreams] = {};
    size_t out_cap[kStreams] = {};
    cudaEvent_t ev_h2d[kStreams] = {}, ev_kernel[kStreams] = {}, ev_d2h[kStreams] = {};   // per buffer slot
    unsigned long long* d_err = nullptr;          // one status word, atomicMin'ed by every launch of a call
    unsigned long long* h_err = nullptr;          // pinned
    void* d_file = nullptr;                       // whole .basis file of the current b2bu_read_to call (device CRC + in-place slices)
    size_t file_cap = 0;
    uint32_t* d_crc = nullptr;                    // CRC partial-sum word
    uint32_t* h_crc = nullptr;                    // pinned
};

// basis.rs:520-535
struct SliceDesc { uint32_t image_index, level_index, flags, orig_width, orig_height, num_blocks_x, num_blocks_y, file_ofs, file_size, crc; };

// NVTX range around a stage of the host runtime (ranges around K1 / K2 / K3 and the copies)
struct NvtxScope {
    explicit NvtxScope(const char* name) { nvtxRangePushA(name); }
    ~NvtxScope() { nvtxRangePop(); }
    NvtxScope(const NvtxScope&) = delete;
    NvtxScope& operator=(const NvtxScope&) = delete;
};

int get_ctx(DeviceCtx** out);
int ensure(void** p, size_t* cap, size_t need);
int cuda_fail(cudaError_t e, const char* what);
int decode_status_word(unsigned long long w, uint64_t* first_bad);
void count_launch(uint64_t n);
int uastc_transcode_slices_based(int target, const void* d_blocks, void* d_out, const b2bu_slice_dev* slices, uint32_t num_slices,
                                 uint64_t first_index, void* d_status, cudaStream_t stream);

}  // namespace b2bu
