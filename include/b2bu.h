/*
 * b2bu.h -- C ABI of the B200-native Basis Universal transcoder (libb2bu.so).
 *
 * This is the drop-in boundary for the hot path of JakubValtar/basisu_rs.  The reference is a
 * pure-Rust library with no FFI of its own; each entry point below names the reference item it
 * replaces (paths relative to the reference crate root) and is what a `extern "C"` block in a
 * Rust shim would bind.  Plain pointers and sizes only.
 *
 * Conventions
 *   - Every function returns a b2bu_status code (0 = OK).  b2bu_error_string() maps a code to the
 *     exact message string of the reference's Err(String) where one exists.  Reference panics
 *     (unimplemented!/assert!/slice OOB) are reported as codes: a C ABI must not unwind.
 *   - "first failing block aborts": like uastc.rs:161-163, an error in any block fails the whole
 *     call; *first_bad_block receives the lowest failing block index, outputs are unspecified.
 *   - Host-pointer functions copy host->device, run the CUDA kernels and copy back inside the
 *     call.  *_dev functions take device pointers + a CUDA stream (void* == cudaStream_t) and
 *     only enqueue work; they are what the device-resident throughput is measured on.
 *   - There is no CPU fallback.  Without a usable CUDA device every call returns B2BU_ERR_CUDA.
 *   - Thread-safe; one lazily created context per device (constant tables + scratch buffers).
 */
#ifndef B2BU_H
#define B2BU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif

/* Target formats.  1..4 mirror uastc::TargetTextureFormat (src/uastc.rs:41-47). */
enum b2bu_target {
    B2BU_RGBA = 0, /* unpacked RGBA32, row-major image (uastc.rs:89-110)                  64 B/block */
    B2BU_ASTC = 1, /* ASTC 4x4   (target_formats/astc.rs:8)                               16 B/block */
    B2BU_BC7 = 2,  /* BC7        (target_formats/bc7.rs:9)                                16 B/block */
    B2BU_ETC1 = 3, /* ETC1       (target_formats/etc.rs:11)                                8 B/block */
    B2BU_ETC2 = 4, /* ETC2 RGBA  (target_formats/etc.rs:19) = EAC alpha + ETC1 colour     16 B/block */
    B2BU_UASTC = 5,/* file level only: copy the UASTC payload (basis.rs:175, uastc.rs:85)  16 B/block */
    B2BU_BC1 = 6   /* ETC1S slices only, EXTENSION: the reference has no BC1 (see b2bu_etc1s_transcode_to_bc1) 8 B/block */
};

enum b2bu_status {
    B2BU_OK = 0,
    B2BU_ERR_LENGTH = 1,        /* "data length is not divisible by UASTC block size (16)"  uastc.rs:56 */
    B2BU_ERR_MODE = 2,          /* "invalid mode index"                                      uastc.rs:336 */
    B2BU_ERR_PATTERN = 3,       /* "block pattern is not valid"                              uastc.rs:364 */
    B2BU_ERR_HUFFMAN = 4,       /* huffman.rs:85-106,177,193 messages                                    */
    B2BU_ERR_SELECTOR_CB = 5,   /* "Global/Hybrid selector codebooks are not supported"      mod.rs:532,536 */
    B2BU_ERR_PREDICTION = 6,    /* reference panics: malformed ETC1S prediction              mod.rs:304-339 */
    B2BU_ERR_VLC = 7,           /* reference panics: decode_vlc overflow                     mod.rs:602-604 */
    B2BU_ERR_RANGE = 8,         /* reference panics: slice/section outside the file, index out of codebook */
    B2BU_ERR_SIG = 9,           /* "Sig mismatch, not a Basis Universal file"                basis.rs:309 */
    B2BU_ERR_HEADER_SIZE = 10,  /* "Expected at least 77 byte header..." / "...unexpected header size" :313-327 */
    B2BU_ERR_HEADER_CRC = 11,   /* "Header CRC16 failed"                                     basis.rs:332 */
    B2BU_ERR_DATA_CRC = 12,     /* "Data CRC16 failed"                                       basis.rs:12 */
    B2BU_ERR_TEX_FORMAT = 13,   /* "Unknown texture format"                                  basis.rs:410 */
    B2BU_ERR_UNIMPLEMENTED = 14,/* reference unimplemented!(): ETC1S file to etc2/uastc/astc/bc7 basis.rs:171.. */
    B2BU_ERR_ALPHA_SLICES = 15, /* "File has alpha, but slice count is odd" etc.             basis.rs:19,30,35 */
    B2BU_ERR_SLICE_DESC = 16,   /* "Expected 23 byte slice desc at pos ..."                  basis.rs:350 */
    B2BU_ERR_ARGUMENT = 17,     /* bad argument to this C API (null pointer, short output buffer, bad target) */
    B2BU_ERR_CUDA = 18,         /* CUDA runtime failure or no device: see b2bu_last_cuda_error()           */
    B2BU_ERR_NOMEM = 19         /* host allocation failed inside the library (a C ABI must not unwind)     */
};

const char* b2bu_error_string(int status);
const char* b2bu_last_cuda_error(void);

/* Selects the CUDA device for the calling thread's subsequent calls and creates its context
 * (uploads the constant tables).  Optional: every entry point initialises device 0 lazily. */
int b2bu_init(int device);
int b2bu_device_count(int* count);

/* bytes per output block for a target (64/16/16/8/16/16) */
size_t b2bu_block_bytes(int target);

/* Pinned host memory for inputs/outputs of the host-pointer functions (makes their copies
 * asynchronous and lets them overlap with the kernels).  Plain malloc'd memory also works. */
void* b2bu_host_alloc(size_t bytes);
void b2bu_host_free(void* p);

/* ---- single block API: src/lib.rs:29-53 ---------------------------------------------------- */
int b2bu_unpack_uastc_block_to_rgba(const uint8_t in[16], uint32_t out[16]);       /* lib.rs:29 */
int b2bu_transcode_uastc_block_to_astc(const uint8_t in[16], uint8_t out[16]);     /* lib.rs:33 */
int b2bu_transcode_uastc_block_to_bc7(const uint8_t in[16], uint8_t out[16]);      /* lib.rs:39 */
int b2bu_transcode_uastc_block_to_etc1(const uint8_t in[16], uint8_t out[8]);      /* lib.rs:43 */
int b2bu_transcode_uastc_block_to_etc2(const uint8_t in[16], uint8_t out[16]);     /* lib.rs:49 */

/* ---- slice level: uastc::Decoder (src/uastc.rs:77-165) -------------------------------------- */
/* Decoder::transcode / _transcode_into (uastc.rs:112-145).  target in {ASTC,BC7,ETC1,ETC2}.
 * out_bytes must be >= (nbytes/16) * b2bu_block_bytes(target). */
int b2bu_uastc_transcode(int target, const uint8_t* blocks, size_t nbytes, uint8_t* out, size_t out_bytes,
                         uint64_t* first_bad_block);
/* Decoder::decode_to_rgba (uastc.rs:89-110): nbytes/16 blocks, raster order, blocks_per_row per
 * row -> row-major RGBA8 image with a pitch of 4*blocks_per_row pixels.  nbytes/16 must be a
 * multiple of blocks_per_row.  out_pixels holds 16*(nbytes/16) pixels (R,G,B,A bytes). */
int b2bu_uastc_decode_rgba(const uint8_t* blocks, size_t nbytes, size_t blocks_per_row, uint32_t* out_pixels,
                           size_t out_pixel_count, uint64_t* first_bad_block);

/* Device-resident variants (what bench.py's `value` times).  d_status points to 8 bytes of device
 * memory: reset it with b2bu_status_reset_dev, launch any number of calls on the same stream,
 * then read it back with b2bu_status_read_dev (synchronises the stream). */
int b2bu_uastc_transcode_dev(int target, const void* d_blocks, size_t nbytes, size_t blocks_per_row, void* d_out,
                             size_t out_bytes, void* d_status, void* stream);
/* Several slices of one device buffer in one call -- the mip chain of a texture, the images of a file
 * (basis.rs:145-260 loop over the slice descriptors and call Decoder::transcode once per slice).  in_ofs /
 * out_ofs are byte offsets from d_blocks / d_out (multiples of 16; out_ofs of B2BU_ETC1: of 8, so that
 * slices with odd block counts can be packed back to back); blocks_per_row is only used by
 * B2BU_RGBA.  Slices that follow each other without a gap in both buffers are merged into ONE kernel
 * launch (B2BU_RGBA: if they also have the same blocks_per_row, i.e. form one taller image): a full mip
 * chain is a single launch instead of 14, a batch of equally sized textures is one launch.  In d_status,
 * block indices count through the slices in array order. */
typedef struct b2bu_slice_dev {
    uint64_t in_ofs, out_ofs, nblocks;
    uint32_t blocks_per_row, reserved;
} b2bu_slice_dev;
int b2bu_uastc_transcode_slices_dev(int target, const void* d_blocks, void* d_out, const b2bu_slice_dev* slices,
                                    uint32_t num_slices, void* d_status, void* stream);
int b2bu_status_reset_dev(void* d_status, void* stream);
int b2bu_status_read_dev(const void* d_status, void* stream, uint64_t* first_bad_block);
/* Measurement aid: thread-level integer instruction throughput of the current device in Tops/s, for a stream of alu-pipe
 * instructions (LOP3 / SHF) and for an alu / fma-pipe mix (LOP3 / IMAD).  This is the INT denominator of the roofline
 * (north_star: the slower of bytes / HBM bandwidth and integer ops / INT throughput). */
int b2bu_probe_int_peak(double* alu_tops, double* mixed_tops);
/* number of kernel launches issued by this library in this process (bench.py's gpu_launches) */
uint64_t b2bu_launch_count(void);

/* ---- ETC1S / BasisLZ: basis_lz::Decoder (src/basis_lz/mod.rs:50-186) ------------------------ */
typedef struct b2bu_etc1s b2bu_etc1s;
/* Decoder::new (mod.rs:64-95): decodes both codebooks and the four slice Huffman models on the
 * host (once per file) and uploads them. */
int b2bu_etc1s_open(uint32_t endpoint_count, uint32_t selector_count, const uint8_t* endpoint_data, size_t endpoint_len,
                    const uint8_t* selector_data, size_t selector_len, const uint8_t* tables_data, size_t tables_len,
                    int is_video, b2bu_etc1s** handle);
void b2bu_etc1s_close(b2bu_etc1s* handle);
/* Decoder::transcode_to_etc1 (mod.rs:153-186): out = nbx*nby*8 bytes. */
int b2bu_etc1s_transcode_to_etc1(b2bu_etc1s* h, uint32_t nbx, uint32_t nby, const uint8_t* slice, size_t slice_len,
                                 uint8_t* out, size_t out_bytes);
/* Decoder::decode_to_rgba (mod.rs:97-151): alpha_slice may be NULL.  out = nbx*nby*64 bytes. */
int b2bu_etc1s_decode_to_rgba(b2bu_etc1s* h, uint32_t nbx, uint32_t nby, const uint8_t* rgb_slice, size_t rgb_len,
                              const uint8_t* alpha_slice, size_t alpha_len, uint8_t* out, size_t out_bytes);
/* Batched form used for many slices of one file (one launch per phase): slices are given as
 * offsets into one host buffer; every slice has nbx*nby blocks.  target in {ETC1, RGBA}. */
int b2bu_etc1s_transcode_slices(b2bu_etc1s* h, int target, uint32_t nbx, uint32_t nby, const uint8_t* data, size_t data_len,
                                const uint64_t* slice_ofs, const uint64_t* slice_len, uint32_t num_slices,
                                uint8_t* out, size_t out_bytes);

/* EXTENSION -- not in the reference.  BASELINE.json names ETC1S -> BC1, but basisu_rs has no BC1 code (only the unused
 * bc1h0 / bc1h1 hint bits, uastc.rs:31-32), so there is no reference output to be exact to: the result is DEFINED by
 * oracle/basisu_oracle_etc1s.inc etc1s_emit_bc1 (endpoints = RGB565 of the lowest / highest ETC1S colour the block
 * uses, selectors remapped to the nearest of the four BC1 palette entries) and the kernel reproduces that bit for bit.
 * Also reachable as target B2BU_BC1 of b2bu_etc1s_transcode_slices. */
int b2bu_etc1s_transcode_to_bc1(b2bu_etc1s* h, uint32_t num_blocks_x, uint32_t num_blocks_y, const uint8_t* slice,
                                size_t slice_len, uint8_t* out, size_t out_bytes);

/* Device-side duration of the phases of the last call on this handle (CUDA events on the call's stream):
 * K2 entropy decode (all slices), K3 codebook gather, and the D2H copy of the result.  The phases are
 * reported separately because K2 is a latency-bound serial chain per slice while K3 is HBM-bound. */
int b2bu_etc1s_last_timing(b2bu_etc1s* h, float* entropy_ms, float* gather_ms, float* d2h_ms, uint64_t* blocks);

/* How the four slice Huffman models (huffman.rs:133-184; endpoint predictor, delta endpoint, selector, selector
 * run length) sat in K2's shared memory during the last call on this handle: width in bits of each first-level table
 * (a wide set is used when every slice gets an SM of its own, a narrow one for packed batches) and each model's
 * longest code (codes longer than the first-level width are looked up in the flat table in global memory). */
int b2bu_etc1s_table_info(b2bu_etc1s* h, uint32_t l1_bits[4], uint32_t max_code_len[4]);

/* ---- file level: src/basis.rs:8-260, src/lib.rs:63-79 ---------------------------------------- */
typedef struct b2bu_header {       /* basis.rs:419-454 (all 26 fields, widened to u32) */
    uint32_t sig, ver, header_size, header_crc16, data_size, data_crc16, total_slices, total_images, tex_format,
             flags, tex_type, us_per_frame, reserved, userdata0, userdata1, total_endpoints, endpoint_cb_file_ofs,
             endpoint_cb_file_size, total_selectors, selector_cb_file_ofs, selector_cb_file_size, tables_file_ofs,
             tables_file_size, slice_desc_file_ofs, extended_file_ofs, extended_file_size;
} b2bu_header;

typedef struct b2bu_image {        /* lib.rs:63-68 Image<u8> + where its data sits in `out` */
    uint32_t w, h, stride;
    uint32_t reserved;
    uint64_t offset, nbytes;
} b2bu_image;

/* basis.rs:307-336 read_header (sig, size, header CRC16) */
int b2bu_read_header(const uint8_t* buf, size_t len, b2bu_header* header);
/* basis.rs:364-372 crc16 */
uint16_t b2bu_crc16(const uint8_t* data, size_t len, uint16_t crc);
/* The same CRC over len bytes of DEVICE memory (any alignment), computed by the GPU: the register update is
 * GF(2)-linear, so 16 KiB chunks are reduced independently and combined with powers of x.  Synchronises the
 * stream.  b2bu_read_to uses this path for files of 256 KiB and more (one upload of the whole file, CRC and
 * transcode kernels back to back), so that the file-level API is not bound by a single-core CRC loop. */
int b2bu_crc16_dev(const void* d_data, size_t len, uint16_t crc, uint16_t* result, void* stream);

/* basis.rs:8,92,145,175,204,233 read_to_{rgba,etc1,etc2,uastc,astc,bc7}.
 * Call with out == NULL to size: fills header, images[0..min(n,max_images)) (offset/nbytes/w/h/
 * stride), *num_images and *out_needed without touching the GPU.  Call again with a buffer of
 * at least *out_needed bytes to transcode; image i's bytes are at out + images[i].offset.
 * Data CRC (basis.rs:9-13 checks it before anything else): files below 256 KiB are checked on the host in both
 * calls.  For larger files the check runs on the GPU inside the TRANSCODING call; the sizing call does not read
 * the payload at all, so it cannot report "Data CRC16 failed" -- and, so that the reference's error precedence
 * still holds, it reports no error of the file BODY either (slice table, alpha pairing ...): it returns
 * B2BU_OK with *num_images = 0 and *out_needed = 0, and the transcoding call (any out != NULL, out_cap may be 0)
 * delivers the verdict, CRC first.  Header errors are reported by both calls. */
int b2bu_read_to(int target, const uint8_t* buf, size_t len, b2bu_header* header, b2bu_image* images,
                 uint32_t max_images, uint32_t* num_images, uint8_t* out, uint64_t out_cap, uint64_t* out_needed);

/* b2bu_read_to with options.  B2BU_READ_APPLY_Y_FLIP: the reference parses the header's YFlipped flag (Header::has_y_flipped,
 * basis.rs:467-469) but leaves acting on it to the caller -- its own tests do it in tests/common.rs:284-301 (rgba_rows: the
 * first `h` rows of `stride` bytes, in reverse order when the flag is set).  With this option and target B2BU_RGBA, images of a
 * file whose header has the flag come back with their first `h` pixel rows already in that reversed order (rows past `h`, the
 * padding of the last block row, stay where they are); every other case is b2bu_read_to unchanged. */
enum b2bu_read_flags { B2BU_READ_APPLY_Y_FLIP = 1 };
int b2bu_read_to_flags(int target, const uint8_t* buf, size_t len, b2bu_header* header, b2bu_image* images,
                       uint32_t max_images, uint32_t* num_images, uint8_t* out, uint64_t out_cap, uint64_t* out_needed,
                       uint32_t flags);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* B2BU_H */
