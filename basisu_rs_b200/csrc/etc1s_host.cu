// TEMPORARY: ETC1S entry points are wired up after the UASTC path is parity-green on the GPU.
#include "etc1s_host.h"
namespace b2bu {
int etc1s_read_file(int, const uint8_t*, size_t, const b2bu_header&, const SliceDesc*, const b2bu_image*, uint32_t, bool, uint8_t*) { return B2BU_ERR_UNIMPLEMENTED; }
}
extern "C" {
int b2bu_etc1s_open(uint32_t, uint32_t, const uint8_t*, size_t, const uint8_t*, size_t, const uint8_t*, size_t, int, b2bu_etc1s**) { return B2BU_ERR_UNIMPLEMENTED; }
void b2bu_etc1s_close(b2bu_etc1s*) {}
int b2bu_etc1s_transcode_to_etc1(b2bu_etc1s*, uint32_t, uint32_t, const uint8_t*, size_t, uint8_t*, size_t) { return B2BU_ERR_UNIMPLEMENTED; }
int b2bu_etc1s_decode_to_rgba(b2bu_etc1s*, uint32_t, uint32_t, const uint8_t*, size_t, const uint8_t*, size_t, uint8_t*, size_t) { return B2BU_ERR_UNIMPLEMENTED; }
int b2bu_etc1s_transcode_slices(b2bu_etc1s*, int, uint32_t, uint32_t, const uint8_t*, size_t, const uint64_t*, const uint64_t*, uint32_t, uint8_t*, size_t) { return B2BU_ERR_UNIMPLEMENTED; }
}
