//! The public API of JakubValtar/basisu_rs (reference src/lib.rs:20-79, src/basis.rs:8-260) with every body replaced by a
//! call into libb2bu.so: same names, same argument and return types, same `Error = String` messages (b2bu_error_string()
//! returns the reference's own strings).  Differences a caller can observe: the crate is not `no_std` and not
//! `forbid(unsafe_code)` (it links a CUDA library), and reference panics (`unimplemented!()` for ETC1S files to
//! ASTC / BC7 / ETC2 / UASTC, out-of-range slices, malformed ETC1S predictions) come back as `Err` instead of unwinding.
//!
//! SOURCE ONLY: this image has no rustc / cargo; the file has never been compiled.  The tested boundary is the C ABI.
mod ffi;

use core::ffi::{c_int, CStr};

pub type Error = String; // lib.rs:26
pub type Result<T> = core::result::Result<T, Error>;

pub const UASTC_BLOCK_SIZE: usize = 16;
pub const ASTC_BLOCK_SIZE: usize = 16;
pub const BC7_BLOCK_SIZE: usize = 16;
pub const ETC1_BLOCK_SIZE: usize = 8;
pub const ETC2_BLOCK_SIZE: usize = 16;

fn check(status: c_int) -> Result<()> {
    if status == 0 {
        return Ok(());
    }
    let msg = unsafe { CStr::from_ptr(ffi::b2bu_error_string(status)) };
    Err(msg.to_string_lossy().into_owned())
}

// ---- lib.rs:29-53 ---------------------------------------------------------------------------------------------------
pub fn unpack_uastc_block_to_rgba(data: [u8; UASTC_BLOCK_SIZE]) -> Result<[u32; 16]> {
    let mut out = [0u32; 16];
    check(unsafe { ffi::b2bu_unpack_uastc_block_to_rgba(data.as_ptr(), out.as_mut_ptr()) })?;
    Ok(out)
}

pub fn transcode_uastc_block_to_astc(data: [u8; UASTC_BLOCK_SIZE]) -> Result<[u8; ASTC_BLOCK_SIZE]> {
    let mut out = [0u8; ASTC_BLOCK_SIZE];
    check(unsafe { ffi::b2bu_transcode_uastc_block_to_astc(data.as_ptr(), out.as_mut_ptr()) })?;
    Ok(out)
}

pub fn transcode_uastc_block_to_bc7(data: [u8; UASTC_BLOCK_SIZE]) -> Result<[u8; BC7_BLOCK_SIZE]> {
    let mut out = [0u8; BC7_BLOCK_SIZE];
    check(unsafe { ffi::b2bu_transcode_uastc_block_to_bc7(data.as_ptr(), out.as_mut_ptr()) })?;
    Ok(out)
}

pub fn transcode_uastc_block_to_etc1(data: [u8; UASTC_BLOCK_SIZE]) -> Result<[u8; ETC1_BLOCK_SIZE]> {
    let mut out = [0u8; ETC1_BLOCK_SIZE];
    check(unsafe { ffi::b2bu_transcode_uastc_block_to_etc1(data.as_ptr(), out.as_mut_ptr()) })?;
    Ok(out)
}

pub fn transcode_uastc_block_to_etc2(data: [u8; UASTC_BLOCK_SIZE]) -> Result<[u8; ETC2_BLOCK_SIZE]> {
    let mut out = [0u8; ETC2_BLOCK_SIZE];
    check(unsafe { ffi::b2bu_transcode_uastc_block_to_etc2(data.as_ptr(), out.as_mut_ptr()) })?;
    Ok(out)
}

// ---- lib.rs:63-79 ---------------------------------------------------------------------------------------------------
pub struct Image<T> {
    pub w: u32,
    pub h: u32,
    pub stride: u32,
    pub data: Vec<T>,
}

// ---- basis.rs:419-473 -----------------------------------------------------------------------------------------------
#[derive(Clone, Copy, Debug)]
pub struct Header {
    pub sig: u16, pub ver: u16, pub header_size: u16, pub header_crc16: u16, pub data_size: u32, pub data_crc16: u16,
    pub total_slices: u32, pub total_images: u32, pub tex_format: u8, pub flags: u16, pub tex_type: u8, pub us_per_frame: u32,
    pub reserved: u32, pub userdata0: u32, pub userdata1: u32, pub total_endpoints: u16, pub endpoint_cb_file_ofs: u32,
    pub endpoint_cb_file_size: u32, pub total_selectors: u16, pub selector_cb_file_ofs: u32, pub selector_cb_file_size: u32,
    pub tables_file_ofs: u32, pub tables_file_size: u32, pub slice_desc_file_ofs: u32, pub extended_file_ofs: u32,
    pub extended_file_size: u32,
}

impl Header {
    pub fn has_alpha(&self) -> bool { (self.flags & 4) != 0 }      // HeaderFlags::HasAlphaSlices, basis.rs:463-465
    pub fn has_y_flipped(&self) -> bool { (self.flags & 2) != 0 }  // HeaderFlags::YFlipped, basis.rs:467-469

    fn from_c(h: &ffi::b2bu_header) -> Self {
        Header {
            sig: h.sig as u16, ver: h.ver as u16, header_size: h.header_size as u16, header_crc16: h.header_crc16 as u16,
            data_size: h.data_size, data_crc16: h.data_crc16 as u16, total_slices: h.total_slices, total_images: h.total_images,
            tex_format: h.tex_format as u8, flags: h.flags as u16, tex_type: h.tex_type as u8, us_per_frame: h.us_per_frame,
            reserved: h.reserved, userdata0: h.userdata0, userdata1: h.userdata1, total_endpoints: h.total_endpoints as u16,
            endpoint_cb_file_ofs: h.endpoint_cb_file_ofs, endpoint_cb_file_size: h.endpoint_cb_file_size,
            total_selectors: h.total_selectors as u16, selector_cb_file_ofs: h.selector_cb_file_ofs,
            selector_cb_file_size: h.selector_cb_file_size, tables_file_ofs: h.tables_file_ofs, tables_file_size: h.tables_file_size,
            slice_desc_file_ofs: h.slice_desc_file_ofs, extended_file_ofs: h.extended_file_ofs, extended_file_size: h.extended_file_size,
        }
    }
}

/// basis.rs:8,92,145,175,204,233: the size query, then the transcoding call (for files of 256 KiB and more the data CRC and every
/// error of the file body are delivered by the second call, CRC first, as in the reference)
fn read_to(target: c_int, buf: &[u8]) -> Result<(Header, Vec<Image<u8>>)> {
    let mut header = ffi::b2bu_header::default();
    let (mut count, mut need) = (0u32, 0u64);
    check(unsafe {
        ffi::b2bu_read_to(target, buf.as_ptr(), buf.len(), &mut header, core::ptr::null_mut(), 0, &mut count, core::ptr::null_mut(), 0, &mut need)
    })?;
    let mut images = vec![ffi::b2bu_image::default(); count.max(1) as usize];
    let mut out = vec![0u8; need.max(1) as usize];
    check(unsafe {
        ffi::b2bu_read_to(target, buf.as_ptr(), buf.len(), &mut header, images.as_mut_ptr(), count, &mut count, out.as_mut_ptr(), need, &mut need)
    })?;
    let imgs = images[..count as usize]
        .iter()
        .map(|im| Image { w: im.w, h: im.h, stride: im.stride, data: out[im.offset as usize..(im.offset + im.nbytes) as usize].to_vec() })
        .collect();
    Ok((Header::from_c(&header), imgs))
}

pub fn read_to_rgba(buf: &[u8]) -> Result<(Header, Vec<Image<u8>>)> { read_to(ffi::B2BU_RGBA, buf) }
pub fn read_to_etc1(buf: &[u8]) -> Result<Vec<Image<u8>>> { read_to(ffi::B2BU_ETC1, buf).map(|r| r.1) }
pub fn read_to_etc2(buf: &[u8]) -> Result<Vec<Image<u8>>> { read_to(ffi::B2BU_ETC2, buf).map(|r| r.1) }
pub fn read_to_uastc(buf: &[u8]) -> Result<Vec<Image<u8>>> { read_to(ffi::B2BU_UASTC, buf).map(|r| r.1) }
pub fn read_to_astc(buf: &[u8]) -> Result<Vec<Image<u8>>> { read_to(ffi::B2BU_ASTC, buf).map(|r| r.1) }
pub fn read_to_bc7(buf: &[u8]) -> Result<Vec<Image<u8>>> { read_to(ffi::B2BU_BC7, buf).map(|r| r.1) }
