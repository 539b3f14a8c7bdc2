//! `extern "C"` declarations, 1:1 with include/b2bu.h (each item cites the reference item it replaces).
//! SOURCE ONLY: never compiled here (no rustc in the build image).
#![allow(non_camel_case_types, dead_code)]
use core::ffi::{c_char, c_int, c_void};

// enum b2bu_target
pub const B2BU_RGBA: c_int = 0;
pub const B2BU_ASTC: c_int = 1;
pub const B2BU_BC7: c_int = 2;
pub const B2BU_ETC1: c_int = 3;
pub const B2BU_ETC2: c_int = 4;
pub const B2BU_UASTC: c_int = 5;
pub const B2BU_BC1: c_int = 6; // extension: the reference has no BC1
pub const B2BU_READ_APPLY_Y_FLIP: u32 = 1;

/// basis.rs:419-454, all 26 fields widened to u32, in file order
#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct b2bu_header {
    pub sig: u32, pub ver: u32, pub header_size: u32, pub header_crc16: u32, pub data_size: u32, pub data_crc16: u32,
    pub total_slices: u32, pub total_images: u32, pub tex_format: u32, pub flags: u32, pub tex_type: u32, pub us_per_frame: u32,
    pub reserved: u32, pub userdata0: u32, pub userdata1: u32, pub total_endpoints: u32, pub endpoint_cb_file_ofs: u32,
    pub endpoint_cb_file_size: u32, pub total_selectors: u32, pub selector_cb_file_ofs: u32, pub selector_cb_file_size: u32,
    pub tables_file_ofs: u32, pub tables_file_size: u32, pub slice_desc_file_ofs: u32, pub extended_file_ofs: u32,
    pub extended_file_size: u32,
}

/// lib.rs:63-68 Image<u8> + where its bytes sit in the caller's output buffer
#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct b2bu_image { pub w: u32, pub h: u32, pub stride: u32, pub reserved: u32, pub offset: u64, pub nbytes: u64 }

#[repr(C)]
pub struct b2bu_etc1s { _private: [u8; 0] }

#[repr(C)]
#[derive(Clone, Copy)]
pub struct b2bu_slice_dev { pub in_ofs: u64, pub out_ofs: u64, pub nblocks: u64, pub blocks_per_row: u32, pub reserved: u32 }

extern "C" {
    pub fn b2bu_error_string(status: c_int) -> *const c_char;
    pub fn b2bu_last_cuda_error() -> *const c_char;
    pub fn b2bu_init(device: c_int) -> c_int;
    pub fn b2bu_device_count(count: *mut c_int) -> c_int;
    pub fn b2bu_block_bytes(target: c_int) -> usize;
    pub fn b2bu_host_alloc(bytes: usize) -> *mut c_void;
    pub fn b2bu_host_free(p: *mut c_void);
    // lib.rs:29-53
    pub fn b2bu_unpack_uastc_block_to_rgba(inp: *const u8, out: *mut u32) -> c_int;
    pub fn b2bu_transcode_uastc_block_to_astc(inp: *const u8, out: *mut u8) -> c_int;
    pub fn b2bu_transcode_uastc_block_to_bc7(inp: *const u8, out: *mut u8) -> c_int;
    pub fn b2bu_transcode_uastc_block_to_etc1(inp: *const u8, out: *mut u8) -> c_int;
    pub fn b2bu_transcode_uastc_block_to_etc2(inp: *const u8, out: *mut u8) -> c_int;
    // uastc.rs:89-165
    pub fn b2bu_uastc_transcode(target: c_int, blocks: *const u8, nbytes: usize, out: *mut u8, out_bytes: usize, first_bad_block: *mut u64) -> c_int;
    pub fn b2bu_uastc_decode_rgba(blocks: *const u8, nbytes: usize, blocks_per_row: usize, out_px: *mut u32, out_px_count: usize,
                                  first_bad_block: *mut u64) -> c_int;
    pub fn b2bu_uastc_transcode_dev(target: c_int, d_blocks: *const c_void, nbytes: usize, blocks_per_row: usize, d_out: *mut c_void,
                                    out_bytes: usize, d_status: *mut c_void, stream: *mut c_void) -> c_int;
    pub fn b2bu_uastc_transcode_slices_dev(target: c_int, d_blocks: *const c_void, d_out: *mut c_void, slices: *const b2bu_slice_dev,
                                           num_slices: u32, d_status: *mut c_void, stream: *mut c_void) -> c_int;
    pub fn b2bu_status_reset_dev(d_status: *mut c_void, stream: *mut c_void) -> c_int;
    pub fn b2bu_status_read_dev(d_status: *const c_void, stream: *mut c_void, first_bad_block: *mut u64) -> c_int;
    pub fn b2bu_probe_int_peak(alu_tops: *mut f64, mixed_tops: *mut f64) -> c_int;
    pub fn b2bu_launch_count() -> u64;
    // basis_lz/mod.rs:64-186
    pub fn b2bu_etc1s_open(endpoint_count: u32, selector_count: u32, ep: *const u8, ep_len: usize, sel: *const u8, sel_len: usize,
                           tables: *const u8, tables_len: usize, is_video: c_int, handle: *mut *mut b2bu_etc1s) -> c_int;
    pub fn b2bu_etc1s_close(handle: *mut b2bu_etc1s);
    pub fn b2bu_etc1s_transcode_to_etc1(h: *mut b2bu_etc1s, nbx: u32, nby: u32, slice: *const u8, len: usize, out: *mut u8, out_bytes: usize) -> c_int;
    pub fn b2bu_etc1s_decode_to_rgba(h: *mut b2bu_etc1s, nbx: u32, nby: u32, rgb: *const u8, rgb_len: usize, alpha: *const u8, alpha_len: usize,
                                     out: *mut u8, out_bytes: usize) -> c_int;
    pub fn b2bu_etc1s_transcode_slices(h: *mut b2bu_etc1s, target: c_int, nbx: u32, nby: u32, data: *const u8, data_len: usize,
                                       slice_ofs: *const u64, slice_len: *const u64, num_slices: u32, out: *mut u8, out_bytes: usize) -> c_int;
    pub fn b2bu_etc1s_transcode_to_bc1(h: *mut b2bu_etc1s, nbx: u32, nby: u32, slice: *const u8, len: usize, out: *mut u8, out_bytes: usize) -> c_int;
    pub fn b2bu_etc1s_last_timing(h: *mut b2bu_etc1s, entropy_ms: *mut f32, gather_ms: *mut f32, d2h_ms: *mut f32, blocks: *mut u64) -> c_int;
    pub fn b2bu_etc1s_table_info(h: *mut b2bu_etc1s, l1_bits: *mut u32, max_code_len: *mut u32) -> c_int;
    // basis.rs:8-372
    pub fn b2bu_read_header(buf: *const u8, len: usize, header: *mut b2bu_header) -> c_int;
    pub fn b2bu_crc16(data: *const u8, len: usize, crc: u16) -> u16;
    pub fn b2bu_crc16_dev(d_data: *const c_void, len: usize, crc: u16, result: *mut u16, stream: *mut c_void) -> c_int;
    pub fn b2bu_read_to(target: c_int, buf: *const u8, len: usize, header: *mut b2bu_header, images: *mut b2bu_image, max_images: u32,
                        num_images: *mut u32, out: *mut u8, out_cap: u64, out_needed: *mut u64) -> c_int;
    pub fn b2bu_read_to_flags(target: c_int, buf: *const u8, len: usize, header: *mut b2bu_header, images: *mut b2bu_image, max_images: u32,
                              num_images: *mut u32, out: *mut u8, out_cap: u64, out_needed: *mut u64, flags: u32) -> c_int;
}
