// Internal launch interface between the C-ABI host layer (capi.cu) and the kernels.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace b2bu {

cudaError_t upload_tables();

// K1<target>: nblocks UASTC blocks at d_in -> d_out.  blocks_per_row is only used by TGT_RGBA.
// d_err: device u64, pre-set to ~0ull; receives min(((index_base + block_index) << 8) | code) over failing blocks.
cudaError_t launch_uastc_transcode(int target, const void* d_in, void* d_out, uint64_t nblocks, uint32_t blocks_per_row,
                                   uint64_t index_base, unsigned long long* d_err, int sm_count, cudaStream_t stream);

}  // namespace b2bu
