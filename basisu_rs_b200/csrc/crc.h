// CRC-16 of the .basis container (reference src/basis.rs:364-372) split into device partial sums + host glue.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

namespace b2bu {

constexpr size_t kCrcChunkBytes = 16384;     // one CTA pass: 256 threads x 64 bytes

// r(D) = D(x) * x^16 mod P continued from state r0 (the reference's loop without the initial / final NOT)
uint16_t crc16_raw_host(const uint8_t* p, size_t n, uint16_t r0);
// r * x^(8 nbytes) mod P: what a remainder becomes when nbytes more bytes follow
uint16_t crc16_shift(uint16_t r, uint64_t nbytes);
uint16_t crc16_xpow(uint64_t nbits);

// XORs into *d_acc (zeroed by the caller) the sum over the nchunks 16 KiB chunks at d_data (16-byte aligned) of
// r(chunk) * x^(8 * bytes between the end of the chunk and the end of the message), bytes_after = message bytes after
// the last chunk.
cudaError_t launch_crc16_partial(const void* d_data, uint64_t nchunks, uint64_t bytes_after, uint32_t* d_acc, int sm_count, cudaStream_t stream);

// Whole message at any alignment: XORs sum_i r(piece_i) * x^(8 * bytes after piece i) over a head / 16 KiB chunks / tail
// split into *d_acc (zeroed by the caller; up to three launches).  crc16_finish turns the sum into basis.rs's crc16(r, crc).
// bytes_after: bytes of the message that follow this piece (a message may be fed piece by piece, in any order).
cudaError_t launch_crc16_dev(const void* d_data, uint64_t len, uint32_t* d_acc, int sm_count, cudaStream_t stream, uint64_t bytes_after = 0);
uint16_t crc16_finish(uint32_t acc, uint64_t len, uint16_t crc);

}  // namespace b2bu
