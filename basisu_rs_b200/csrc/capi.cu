// C ABI (include/b2bu.h) + host runtime: per-device context, scratch buffers, streams, the
// chunked H2D -> kernel -> D2H pipeline of the host-pointer entry points, error mapping.
// Host-side counterpart of the reference's crate-private slice API (src/uastc.rs:77-165) and of
// its single-block API (src/lib.rs:29-53).
#include <atomic>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>
#include <algorithm>

#include "../../include/b2bu.h"
#include "host_internal.h"
#include "kernels.h"
#include "crc.h"

namespace b2bu {

static thread_local int t_device = 0;
static thread_local char t_cuda_err[256] = "";
static std::atomic<uint64_t> g_launches{0};

static DeviceCtx g_ctx[kMaxDevices];

int cuda_fail(cudaError_t e, const char* what)
{
    snprintf(t_cuda_err, sizeof t_cuda_err, "%s: %s", what, cudaGetErrorString(e));
    return B2BU_ERR_CUDA;
}
#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return cuda_fail(e_, #call); } while (0)

void count_launch(uint64_t n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int get_ctx(DeviceCtx** out)
{
    const int dev = t_device;
    if (dev < 0 || dev >= kMaxDevices) return B2BU_ERR_ARGUMENT;
    DeviceCtx& c = g_ctx[dev];
    std::lock_guard<std::mutex> lk(c.init_mu);
    CK(cudaSetDevice(dev));
    if (!c.ready) {
        cudaDeviceProp prop;
        CK(cudaGetDeviceProperties(&prop, dev));
        c.device = dev;
        c.sm_count = prop.multiProcessorCount;
        CK(upload_tables());
        for (int i = 0; i < kStreams; i++) {
            CK(cudaStreamCreateWithFlags(&c.streams[i], cudaStreamNonBlocking));
            CK(cudaEventCreateWithFlags(&c.ev_h2d[i], cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&c.ev_kernel[i], cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&c.ev_d2h[i], cudaEventDisableTiming));
        }
        CK(cudaMalloc(&c.d_err, sizeof(unsigned long long)));
        CK(cudaMallocHost(&c.h_err, sizeof(unsigned long long)));
        CK(cudaMalloc(&c.d_crc, sizeof(uint32_t)));
        CK(cudaMallocHost(&c.h_crc, sizeof(uint32_t)));
        c.ready = true;
    }
    *out = &c;
    return B2BU_OK;
}

int ensure(void** p, size_t* cap, size_t need)
{
    if (*cap >= need) return B2BU_OK;
    if (*p) { cudaFree(*p); *p = nullptr; *cap = 0; }
    size_t want = need + need / 4 + 4096;
    cudaError_t e = cudaMalloc(p, want);
    if (e != cudaSuccess) { want = need; e = cudaMalloc(p, want); }
    if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc(scratch)");
    *cap = want;
    return B2BU_OK;
}

int decode_status_word(unsigned long long w, uint64_t* first_bad)
{
    if (w == ~0ull) { if (first_bad) *first_bad = ~0ull; return B2BU_OK; }
    if (first_bad) *first_bad = (uint64_t)(w >> 8);
    const unsigned code = (unsigned)(w & 0xFF);
    return code == ERR_MODE_DEV ? B2BU_ERR_MODE : code == ERR_PATTERN_DEV ? B2BU_ERR_PATTERN : B2BU_ERR_CUDA;
}

// Host-pointer UASTC path: chunks of whole block rows, round-robin over kStreams streams so that
// the H2D copy of chunk c+1, the kernel of chunk c and the D2H copy of chunk c-1 overlap.
static int uastc_host_run(int target, const uint8_t* blocks, size_t nblocks, size_t bpr, uint8_t* out, uint64_t* first_bad)
{
    NvtxScope nv("b2bu K1 host pipeline (H2D | uastc_sorted_kernel | D2H)");
    DeviceCtx* c;
    int st = get_ctx(&c);
    if (st) return st;
    std::lock_guard<std::mutex> lk(c->run_mu);
    const size_t ob = b2bu_block_bytes(target);
    size_t chunk = kChunkBlocks;
    if (target == B2BU_RGBA) {                           // whole block rows keep the output contiguous
        chunk = (kChunkBlocks / bpr) * bpr;
        if (chunk == 0) chunk = bpr;
    }
    if (chunk > nblocks) chunk = nblocks;
    for (int s = 0; s < kStreams; s++) {
        if ((st = ensure(&c->d_in[s], &c->in_cap[s], chunk * 16))) return st;
        if ((st = ensure(&c->d_out[s], &c->out_cap[s], chunk * ob))) return st;
    }
    // chunk schedule: the first H2D copy and the last D2H copy cannot overlap with anything, so the pipeline ramps
    // up (chunk/8, /4, /2), runs full-size chunks, and ramps down again (/2, /4, /8)
    const size_t unit = target == B2BU_RGBA ? bpr : 1;              // RGBA chunks are whole block rows
    std::vector<size_t> sched;
    {
        const size_t ramp[3] = {std::max(chunk >> 3, unit) / unit * unit, std::max(chunk >> 2, unit) / unit * unit, std::max(chunk >> 1, unit) / unit * unit};
        const size_t ramp_total = ramp[0] + ramp[1] + ramp[2];
        size_t left = nblocks;
        if (nblocks > 2 * ramp_total) {
            for (int k = 0; k < 3; k++) { sched.push_back(ramp[k]); left -= ramp[k]; }
            while (left > ramp_total) { const size_t n = std::min(chunk, left - ramp_total); sched.push_back(n); left -= n; }
            for (int k = 2; k >= 0 && left; k--) { const size_t n = k ? std::min(ramp[k], left) : left; sched.push_back(n); left -= n; }
        } else {
            while (left) { const size_t n = std::min(ramp[1], left); sched.push_back(n); left -= n; }
        }
    }
    // Three streams, one per engine (H2D copies, kernels, D2H copies), chained per buffer slot with events: the copy
    // engines always have the next transfer queued and no transfer waits behind an unrelated one of the other direction.
    cudaStream_t sH = c->streams[0], sK = c->streams[1], sD = c->streams[2];
    CK(cudaMemsetAsync(c->d_err, 0xFF, sizeof(unsigned long long), sK));
    size_t done = 0;
    int ci = 0;
    for (const size_t n : sched) {
        const int s = ci % kStreams;
        if (ci >= kStreams) CK(cudaStreamWaitEvent(sH, c->ev_kernel[s], 0));      // the kernel that read d_in[s] has finished
        CK(cudaMemcpyAsync(c->d_in[s], blocks + done * 16, n * 16, cudaMemcpyHostToDevice, sH));
        CK(cudaEventRecord(c->ev_h2d[s], sH));
        CK(cudaStreamWaitEvent(sK, c->ev_h2d[s], 0));
        if (ci >= kStreams) CK(cudaStreamWaitEvent(sK, c->ev_d2h[s], 0));         // the copy that read d_out[s] has finished
        CK(launch_uastc_transcode(target, c->d_in[s], c->d_out[s], n, (uint32_t)bpr, done, c->d_err, c->sm_count, sK));
        count_launch(1);
        CK(cudaEventRecord(c->ev_kernel[s], sK));
        CK(cudaStreamWaitEvent(sD, c->ev_kernel[s], 0));
        CK(cudaMemcpyAsync(out + done * ob, c->d_out[s], n * ob, cudaMemcpyDeviceToHost, sD));
        CK(cudaEventRecord(c->ev_d2h[s], sD));
        done += n;
        ci++;
    }
    CK(cudaMemcpyAsync(c->h_err, c->d_err, sizeof(unsigned long long), cudaMemcpyDeviceToHost, sD));   // after the last kernel (sD waited on it)
    CK(cudaStreamSynchronize(sD));
    return decode_status_word(*c->h_err, first_bad);
}

}  // namespace b2bu

using namespace b2bu;

extern "C" {

const char* b2bu_error_string(int status)
{
    switch (status) {
    case B2BU_OK: return "ok";
    case B2BU_ERR_LENGTH: return "data length is not divisible by UASTC block size (16)";
    case B2BU_ERR_MODE: return "invalid mode index";
    case B2BU_ERR_PATTERN: return "block pattern is not valid";
    case B2BU_ERR_HUFFMAN: return "invalid Huffman table or code (see huffman.rs:85-106,177,193)";
    case B2BU_ERR_SELECTOR_CB: return "Global/Hybrid selector codebooks are not supported";
    case B2BU_ERR_PREDICTION: return "malformed ETC1S prediction (reference panics)";
    case B2BU_ERR_VLC: return "VLC value overflow (reference panics)";
    case B2BU_ERR_RANGE: return "section, slice or codebook index out of range (reference panics)";
    case B2BU_ERR_SIG: return "Sig mismatch, not a Basis Universal file";
    case B2BU_ERR_HEADER_SIZE: return "unexpected header size";
    case B2BU_ERR_HEADER_CRC: return "Header CRC16 failed";
    case B2BU_ERR_DATA_CRC: return "Data CRC16 failed";
    case B2BU_ERR_TEX_FORMAT: return "Unknown texture format";
    case B2BU_ERR_UNIMPLEMENTED: return "not implemented for this texture format (reference: unimplemented!())";
    case B2BU_ERR_ALPHA_SLICES: return "alpha slice layout is invalid";
    case B2BU_ERR_SLICE_DESC: return "slice description array is truncated";
    case B2BU_ERR_ARGUMENT: return "invalid argument";
    case B2BU_ERR_CUDA: return "CUDA error";
    case B2BU_ERR_NOMEM: return "out of host memory";
    default: return "unknown status";
    }
}

const char* b2bu_last_cuda_error(void) { return t_cuda_err; }

int b2bu_device_count(int* count)
{
    if (!count) return B2BU_ERR_ARGUMENT;
    CK(cudaGetDeviceCount(count));
    return B2BU_OK;
}

int b2bu_init(int device)
{
    if (device < 0 || device >= kMaxDevices) return B2BU_ERR_ARGUMENT;
    t_device = device;
    DeviceCtx* c;
    return get_ctx(&c);
}

size_t b2bu_block_bytes(int target) { return target == B2BU_RGBA ? 64 : target == B2BU_ETC1 ? 8 : 16; }

void* b2bu_host_alloc(size_t bytes)
{
    void* p = nullptr;
    if (cudaSetDevice(t_device) != cudaSuccess) return nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) return nullptr;
    return p;
}
void b2bu_host_free(void* p) { if (p) cudaFreeHost(p); }

uint64_t b2bu_launch_count(void) { return g_launches.load(); }

int b2bu_uastc_transcode(int target, const uint8_t* blocks, size_t nbytes, uint8_t* out, size_t out_bytes, uint64_t* first_bad_block)
{
    if (first_bad_block) *first_bad_block = ~0ull;
    if (target < B2BU_ASTC || target > B2BU_ETC2) return B2BU_ERR_ARGUMENT;
    if (nbytes % 16 != 0) return B2BU_ERR_LENGTH;                                  // uastc.rs:55-56
    const size_t n = nbytes / 16;
    if (n == 0) return B2BU_OK;
    if (!blocks || !out || out_bytes < n * b2bu_block_bytes(target)) return B2BU_ERR_ARGUMENT;
    return uastc_host_run(target, blocks, n, 1, out, first_bad_block);
}

int b2bu_uastc_decode_rgba(const uint8_t* blocks, size_t nbytes, size_t blocks_per_row, uint32_t* out_pixels,
                           size_t out_pixel_count, uint64_t* first_bad_block)
{
    if (first_bad_block) *first_bad_block = ~0ull;
    if (nbytes % 16 != 0) return B2BU_ERR_LENGTH;
    const size_t n = nbytes / 16;
    if (n == 0) return B2BU_OK;
    if (!blocks || !out_pixels || blocks_per_row == 0 || n % blocks_per_row != 0 || out_pixel_count < n * 16) return B2BU_ERR_ARGUMENT;
    return uastc_host_run(B2BU_RGBA, blocks, n, blocks_per_row, reinterpret_cast<uint8_t*>(out_pixels), first_bad_block);
}

int b2bu_unpack_uastc_block_to_rgba(const uint8_t in[16], uint32_t out[16]) { return b2bu_uastc_decode_rgba(in, 16, 1, out, 16, nullptr); }
int b2bu_transcode_uastc_block_to_astc(const uint8_t in[16], uint8_t out[16]) { return b2bu_uastc_transcode(B2BU_ASTC, in, 16, out, 16, nullptr); }
int b2bu_transcode_uastc_block_to_bc7(const uint8_t in[16], uint8_t out[16]) { return b2bu_uastc_transcode(B2BU_BC7, in, 16, out, 16, nullptr); }
int b2bu_transcode_uastc_block_to_etc1(const uint8_t in[16], uint8_t out[8]) { return b2bu_uastc_transcode(B2BU_ETC1, in, 16, out, 8, nullptr); }
int b2bu_transcode_uastc_block_to_etc2(const uint8_t in[16], uint8_t out[16]) { return b2bu_uastc_transcode(B2BU_ETC2, in, 16, out, 16, nullptr); }

int b2bu_uastc_transcode_dev(int target, const void* d_blocks, size_t nbytes, size_t blocks_per_row, void* d_out,
                             size_t out_bytes, void* d_status, void* stream)
{
    if (target < B2BU_RGBA || target > B2BU_ETC2) return B2BU_ERR_ARGUMENT;
    if (nbytes % 16 != 0) return B2BU_ERR_LENGTH;
    const size_t n = nbytes / 16;
    if (n == 0) return B2BU_OK;
    if (!d_blocks || !d_out || !d_status || out_bytes < n * b2bu_block_bytes(target)) return B2BU_ERR_ARGUMENT;
    if (target == B2BU_RGBA && (blocks_per_row == 0 || n % blocks_per_row != 0)) return B2BU_ERR_ARGUMENT;
    if ((reinterpret_cast<uintptr_t>(d_blocks) | reinterpret_cast<uintptr_t>(d_out)) & 15) return B2BU_ERR_ARGUMENT;
    DeviceCtx* c;
    int st = get_ctx(&c);
    if (st) return st;
    CK(launch_uastc_transcode(target, d_blocks, d_out, n, (uint32_t)(blocks_per_row ? blocks_per_row : 1), 0,
                              reinterpret_cast<unsigned long long*>(d_status), c->sm_count, reinterpret_cast<cudaStream_t>(stream)));
    count_launch(1);
    return B2BU_OK;
}

int b2bu_uastc_transcode_slices_dev(int target, const void* d_blocks, void* d_out, const b2bu_slice_dev* slices, uint32_t num_slices,
                                    void* d_status, void* stream)
{
    return uastc_transcode_slices_based(target, d_blocks, d_out, slices, num_slices, 0, d_status, reinterpret_cast<cudaStream_t>(stream));
}

}  // extern "C"

namespace b2bu {

// b2bu_uastc_transcode_slices_dev with the status-word block numbering starting at first_index
int uastc_transcode_slices_based(int target, const void* d_blocks, void* d_out, const b2bu_slice_dev* slices, uint32_t num_slices,
                                 uint64_t first_index, void* d_status, cudaStream_t stream)
{
    if (target < B2BU_RGBA || target > B2BU_ETC2) return B2BU_ERR_ARGUMENT;
    if (num_slices == 0) return B2BU_OK;
    if (!d_blocks || !d_out || !slices || !d_status) return B2BU_ERR_ARGUMENT;
    if ((reinterpret_cast<uintptr_t>(d_blocks) | reinterpret_cast<uintptr_t>(d_out)) & 15) return B2BU_ERR_ARGUMENT;
    const uint64_t ob = b2bu_block_bytes(target);
    for (uint32_t i = 0; i < num_slices; i++) {
        const b2bu_slice_dev& s = slices[i];
        if ((s.in_ofs & 15) || (s.out_ofs & (target == B2BU_ETC1 ? 7 : 15))) return B2BU_ERR_ARGUMENT;
        if (target == B2BU_RGBA && s.nblocks && (s.blocks_per_row == 0 || s.nblocks % s.blocks_per_row != 0)) return B2BU_ERR_ARGUMENT;
    }
    DeviceCtx* c;
    int st = get_ctx(&c);
    if (st) return st;
    const uint8_t* in = static_cast<const uint8_t*>(d_blocks);
    uint8_t* out = static_cast<uint8_t*>(d_out);
    uint64_t base = first_index;
    for (uint32_t i = 0; i < num_slices;) {
        // a run of slices that is one contiguous block array on both sides.  RGBA output depends on the slice shape, but
        // slices of the same width that follow each other are one taller image (a batch of equally sized textures)
        uint64_t n = slices[i].nblocks;
        uint32_t j = i + 1;
        while (j < num_slices && slices[j].in_ofs == slices[i].in_ofs + n * 16 && slices[j].out_ofs == slices[i].out_ofs + n * ob &&
               (target != B2BU_RGBA || slices[j].blocks_per_row == slices[i].blocks_per_row))
            n += slices[j++].nblocks;
        if (n) {
            uint64_t skip = 0;
            if (target == B2BU_ETC1 && (slices[i].out_ofs & 15)) {
                // 8-byte blocks: a run that starts in the middle of a 16-byte line gives its first block to the small kernel so
                // that the tile kernel's bulk stores stay 16-byte aligned
                skip = 1;
                CK(launch_uastc_transcode(target, in + slices[i].in_ofs, out + slices[i].out_ofs, 1, 1u, base, reinterpret_cast<unsigned long long*>(d_status),
                                          c->sm_count, stream));
                count_launch(1);
            }
            if (n > skip) {
                CK(launch_uastc_transcode(target, in + slices[i].in_ofs + skip * 16, out + slices[i].out_ofs + skip * ob, n - skip,
                                          slices[i].blocks_per_row ? slices[i].blocks_per_row : 1u, base + skip,
                                          reinterpret_cast<unsigned long long*>(d_status), c->sm_count, stream));
                count_launch(1);
            }
        }
        base += n;
        i = j;
    }
    return B2BU_OK;
}

}  // namespace b2bu

extern "C" {

int b2bu_crc16_dev(const void* d_data, size_t len, uint16_t crc, uint16_t* result, void* stream)
{
    if (!result || (len && !d_data)) return B2BU_ERR_ARGUMENT;
    DeviceCtx* c;
    int st = get_ctx(&c);
    if (st) return st;
    std::lock_guard<std::mutex> lk(c->run_mu);
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    CK(cudaMemsetAsync(c->d_crc, 0, sizeof(uint32_t), s));
    CK(launch_crc16_dev(d_data, len, c->d_crc, c->sm_count, s));
    count_launch(len ? 1 : 0);
    CK(cudaMemcpyAsync(c->h_crc, c->d_crc, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    *result = crc16_finish(*c->h_crc, len, crc);
    return B2BU_OK;
}

int b2bu_status_reset_dev(void* d_status, void* stream)
{
    if (!d_status) return B2BU_ERR_ARGUMENT;
    CK(cudaMemsetAsync(d_status, 0xFF, 8, reinterpret_cast<cudaStream_t>(stream)));
    return B2BU_OK;
}

int b2bu_status_read_dev(const void* d_status, void* stream, uint64_t* first_bad_block)
{
    if (!d_status) return B2BU_ERR_ARGUMENT;
    unsigned long long w = 0;
    CK(cudaMemcpyAsync(&w, d_status, 8, cudaMemcpyDeviceToHost, reinterpret_cast<cudaStream_t>(stream)));
    CK(cudaStreamSynchronize(reinterpret_cast<cudaStream_t>(stream)));
    return decode_status_word(w, first_bad_block);
}

}  // extern "C"
