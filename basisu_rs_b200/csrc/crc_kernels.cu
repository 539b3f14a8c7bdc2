// File-level CRC-16 on the device.  The reference verifies a bytewise CRC-16 over the whole payload
// on one CPU core before it decodes anything (src/basis.rs:338-341, :364-372: init 0xFFFF, polynomial 0x1021, final NOT):
// tens of milliseconds for a 64 MiB file in front of kernels that take tens of microseconds.  The register update is
// linear over GF(2), so with r(D) = D(x) * x^16 mod P the remainder of a concatenation is
//     r(A || B) = r(A) * x^(8 |B|) mod P  xor  r(B),
// and the file's CRC is  ~( ~init * x^(8 n)  xor  sum_i r(chunk_i) * x^(8 * bytes after chunk i) )  for any chunking.
// Each thread runs the reference's own byte step over 64 contiguous bytes; lanes, warps and CTAs are combined with
// multiplications by precomputed powers of x, and every CTA folds its share into one word with atomicXor.
#include "crc.h"

namespace b2bu {

// a * b mod P over GF(2), P = x^16 + x^12 + x^5 + 1
__host__ __device__ inline uint32_t crc_mulmod(uint32_t a, uint32_t b)
{
    uint32_t r = 0;
#pragma unroll
    for (int i = 15; i >= 0; i--) {
        r = ((r << 1) ^ ((r & 0x8000u) ? 0x1021u : 0u)) & 0xFFFFu;
        if ((b >> i) & 1u) r ^= a;
    }
    return r;
}

// basis.rs:366-370, one byte, state without the initial / final NOT
__host__ __device__ inline uint32_t crc_step(uint32_t r, uint32_t b)
{
    const uint32_t q = b ^ (r >> 8);
    const uint32_t k = ((q >> 4) ^ q) & 0xFFu;
    return ((r << 8) ^ k ^ (k << 5) ^ (k << 12)) & 0xFFFFu;
}

__device__ __forceinline__ uint32_t crc_word(uint32_t r, uint32_t w)
{
    r = crc_step(r, w & 0xFFu);
    r = crc_step(r, (w >> 8) & 0xFFu);
    r = crc_step(r, (w >> 16) & 0xFFu);
    return crc_step(r, w >> 24);
}

struct CrcPowers {
    uint16_t lane[32];      // x^(8 * 64 * (31 - lane)): bytes between the end of a lane's chunk and the end of its warp's 2 KiB
    uint16_t warp[8];       // x^(8 * 2048 * (7 - warp))
    uint16_t pow2[48];      // x^(2^j)
};

uint16_t crc16_xpow(uint64_t nbits)
{
    uint32_t r = 1, base = 2;
    while (nbits) {
        if (nbits & 1) r = crc_mulmod(r, base);
        base = crc_mulmod(base, base);
        nbits >>= 1;
    }
    return (uint16_t)r;
}

uint16_t crc16_raw_host(const uint8_t* p, size_t n, uint16_t r0)
{
    uint32_t r = r0;
    for (size_t i = 0; i < n; i++) r = crc_step(r, p[i]);
    return (uint16_t)r;
}

uint16_t crc16_shift(uint16_t r, uint64_t nbytes) { return (uint16_t)crc_mulmod(r, crc16_xpow(8 * nbytes)); }

static CrcPowers make_powers()
{
    CrcPowers p;
    for (int l = 0; l < 32; l++) p.lane[l] = crc16_xpow(8ull * 64 * (31 - l));
    for (int w = 0; w < 8; w++) p.warp[w] = crc16_xpow(8ull * 2048 * (7 - w));
    uint32_t b = 2;
    for (int j = 0; j < 48; j++) { p.pow2[j] = (uint16_t)b; b = crc_mulmod(b, b); }
    return p;
}

// data: 16-byte aligned, nchunks chunks of kCrcChunkBytes; bytes_after: bytes of the message that follow the last chunk
__global__ void __launch_bounds__(256) crc16_partial_kernel(const uint4* __restrict__ data, uint64_t nchunks, uint64_t bytes_after,
                                                            const __grid_constant__ CrcPowers pw, uint32_t* __restrict__ acc)
{
    __shared__ uint32_t wsum[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t mine = 0;                                        // thread 0: this CTA's contribution
    for (uint64_t c = blockIdx.x; c < nchunks; c += gridDim.x) {
        const uint4* p = data + c * (kCrcChunkBytes / 16) + (size_t)threadIdx.x * 4;
        uint32_t r = 0;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const uint4 v = __ldg(p + i);
            r = crc_word(crc_word(crc_word(crc_word(r, v.x), v.y), v.z), v.w);
        }
        r = __reduce_xor_sync(0xFFFFFFFFu, crc_mulmod(r, pw.lane[lane]));
        if (lane == 0) wsum[warp] = crc_mulmod(r, pw.warp[warp]);
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t s = 0;
            for (int w = 0; w < 8; w++) s ^= wsum[w];
            const uint64_t after_bits = 8ull * ((nchunks - 1 - c) * kCrcChunkBytes + bytes_after);
            uint32_t x = 1;
            for (int j = 0; j < 48; j++) if ((after_bits >> j) & 1ull) x = crc_mulmod(x, pw.pow2[j]);
            mine ^= crc_mulmod(s, x);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0 && mine) atomicXor(acc, mine);
}

// the unaligned head and the short tail of a message: n < one chunk, 256 threads, bytewise
__global__ void __launch_bounds__(256) crc16_small_kernel(const uint8_t* __restrict__ p, uint32_t n, uint64_t bytes_after,
                                                          const __grid_constant__ CrcPowers pw, uint32_t* __restrict__ acc)
{
    const uint32_t m = (n + 255u) / 256u, start = threadIdx.x * m;
    const uint32_t end = start + m < n ? start + m : n;
    uint32_t v = 0;
    if (start < n) {
        uint32_t r = 0;
        for (uint32_t i = start; i < end; i++) r = crc_step(r, p[i]);
        const uint64_t after_bits = 8ull * ((uint64_t)(n - end) + bytes_after);
        uint32_t x = 1;
        for (int j = 0; j < 48; j++) if ((after_bits >> j) & 1ull) x = crc_mulmod(x, pw.pow2[j]);
        v = crc_mulmod(r, x);
    }
    v = __reduce_xor_sync(0xFFFFFFFFu, v);
    if ((threadIdx.x & 31) == 0 && v) atomicXor(acc, v);
}

static const CrcPowers& powers()
{
    static const CrcPowers pw = make_powers();
    return pw;
}

cudaError_t launch_crc16_dev(const void* d_data, uint64_t len, uint32_t* d_acc, int sm_count, cudaStream_t stream, uint64_t bytes_after)
{
    const uint8_t* p = static_cast<const uint8_t*>(d_data);
    uint64_t head = (16 - (reinterpret_cast<uintptr_t>(p) & 15)) & 15;
    if (head > len) head = len;
    const uint64_t nchunks = (len - head) / kCrcChunkBytes;
    const uint64_t tail = len - head - nchunks * kCrcChunkBytes;
    if (head) crc16_small_kernel<<<1, 256, 0, stream>>>(p, (uint32_t)head, len - head + bytes_after, powers(), d_acc);
    if (nchunks) {
        const uint64_t cap = (uint64_t)sm_count * 8;
        crc16_partial_kernel<<<(unsigned)(nchunks < cap ? nchunks : cap), 256, 0, stream>>>(reinterpret_cast<const uint4*>(p + head), nchunks, tail + bytes_after, powers(), d_acc);
    }
    if (tail) crc16_small_kernel<<<1, 256, 0, stream>>>(p + head + nchunks * kCrcChunkBytes, (uint32_t)tail, bytes_after, powers(), d_acc);
    return cudaGetLastError();
}

uint16_t crc16_finish(uint32_t acc, uint64_t len, uint16_t crc)
{
    // basis.rs:365,371: the running value starts at !crc and the result is inverted
    return (uint16_t)~(crc16_shift((uint16_t)~crc, len) ^ (uint16_t)acc);
}

cudaError_t launch_crc16_partial(const void* d_data, uint64_t nchunks, uint64_t bytes_after, uint32_t* d_acc, int sm_count, cudaStream_t stream)
{
    if (nchunks == 0) return cudaSuccess;
    const uint64_t cap = (uint64_t)sm_count * 8;
    crc16_partial_kernel<<<(unsigned)(nchunks < cap ? nchunks : cap), 256, 0, stream>>>(reinterpret_cast<const uint4*>(d_data), nchunks, bytes_after, powers(), d_acc);
    return cudaGetLastError();
}

}  // namespace b2bu
