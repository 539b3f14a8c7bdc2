// ETC1S / BasisLZ host side: codebook + Huffman model decode (once per file) and the launch
// logic for K2 (entropy decode) and K3 (codebook gather).  See etc1s_host.cu.
#pragma once
#include <cstdint>
#include "../../include/b2bu.h"
#include "host_internal.h"

namespace b2bu {

// File-level ETC1S path of b2bu_read_to (basis.rs:16-69, :98-123): decodes every slice of the
// file into `out` according to the image plan.  d_file: device copy of the whole file when the caller has uploaded
// it (device CRC path; the slices are then gathered with device-to-device copies on streams[0]), else null.
int etc1s_read_file(int target, const uint8_t* buf, size_t len, const b2bu_header& h, const SliceDesc* descs,
                    const b2bu_image* plan, uint32_t nimg, bool pair, uint8_t* out, const uint8_t* d_file);

}  // namespace b2bu
