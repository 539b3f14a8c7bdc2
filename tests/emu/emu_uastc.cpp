// HOST EMULATION of the device code in basisu_rs_b200/csrc/uastc_device.cuh (TEST INFRASTRUCTURE).
// The CUDA intrinsics the kernels use are shimmed below so that the exact kernel source can be
// exercised on a machine without a GPU (pytest -m "not gpu") against the golden vectors and the
// oracle.  This is never part of the product library and is no CPU fallback: libb2bu.so does not
// contain it.
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <algorithm>

#define B2BU_HOST_EMU 1
#define __device__
#define __host__
#define __forceinline__ inline
struct uint4 { uint32_t x, y, z, w; };
struct uint2 { uint32_t x, y; };
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }
static inline uint2 make_uint2(uint32_t x, uint32_t y) { return uint2{x, y}; }
static inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t s) { return (uint32_t)(((((uint64_t)hi) << 32) | lo) >> (s & 31)); }
static inline uint32_t __funnelshift_l(uint32_t lo, uint32_t hi, uint32_t s) { return (uint32_t)((((((uint64_t)hi) << 32) | lo) << (s & 31)) >> 32); }
static inline uint32_t __brev(uint32_t x)
{
    x = (x >> 16) | (x << 16);
    x = ((x & 0xFF00FF00u) >> 8) | ((x & 0x00FF00FFu) << 8);
    x = ((x & 0xF0F0F0F0u) >> 4) | ((x & 0x0F0F0F0Fu) << 4);
    x = ((x & 0xCCCCCCCCu) >> 2) | ((x & 0x33333333u) << 2);
    return ((x & 0xAAAAAAAAu) >> 1) | ((x & 0x55555555u) << 1);
}
static inline uint32_t __byte_perm(uint32_t a, uint32_t b, uint32_t sel)
{
    const uint64_t v = ((uint64_t)b << 32) | a;
    uint32_t r = 0;
    for (int i = 0; i < 4; i++) r |= (uint32_t)((v >> (8 * ((sel >> (4 * i)) & 7))) & 0xFF) << (8 * i);
    return r;
}
// unsigned bytes of a times unsigned bytes of b (the kernel only uses weights < 256 as u8 x u8), accumulated
static inline uint32_t __dp4a(uint32_t a, uint32_t b, uint32_t c)
{
    for (int i = 0; i < 4; i++) c += ((a >> (8 * i)) & 0xFF) * ((b >> (8 * i)) & 0xFF);
    return c;
}
static inline int __popc(uint32_t x) { return __builtin_popcount(x); }
static inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
static inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
static inline float __uint_as_float(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
using std::min;
using std::max;

#include "../../basisu_rs_b200/csrc/uastc_device.cuh"

using namespace b2bu;

static const DevTables kTables =
#include "../../basisu_rs_b200/csrc/device_tables_gen.inc"
    ;

// Transcodes n blocks exactly as one kernel thread per block would.  Returns 0 or
// ((first_bad_block << 8) | code) like the device status word.
extern "C" __attribute__((visibility("default")))
uint64_t emu_uastc_transcode(int target, const uint8_t* in, size_t n, size_t blocks_per_row, uint8_t* out)
{
    uint64_t status = ~0ull;
    for (size_t i = 0; i < n; i++) {
        uint4 b;
        memcpy(&b, in + 16 * i, 16);
        BlockOut o;
        memset(&o, 0, sizeof o);
        uint32_t e;
        switch (target) {
        case TGT_RGBA: e = transcode_one<TGT_RGBA>(b, kTables, o); break;
        case TGT_ASTC: e = transcode_one<TGT_ASTC>(b, kTables, o); break;
        case TGT_BC7: e = transcode_one<TGT_BC7>(b, kTables, o); break;
        case TGT_ETC1: e = transcode_one<TGT_ETC1>(b, kTables, o); break;
        default: e = transcode_one<TGT_ETC2>(b, kTables, o); break;
        }
        if (e != ERR_OK) { status = std::min<uint64_t>(status, ((uint64_t)i << 8) | e); memset(&o, 0, sizeof o); }
        if (target == TGT_RGBA) {
            const size_t bx = i % blocks_per_row, by = i / blocks_per_row;
            for (int y = 0; y < 4; y++) memcpy(out + (((by * 4 + y) * blocks_per_row + bx) * 16), &o.px[4 * y], 16);
        } else if (target == TGT_ETC1) memcpy(out + 8 * i, &o.etc, 8);
        else memcpy(out + 16 * i, &o.v, 16);
    }
    return status;
}
