// Measurement aid (not product code): issue throughput of the integer instructions the kernels are built from, one kind or a
// fixed mix per kernel, as thread-level ops per clock per SM.  Every instruction is inline PTX on 8 independent chains so that
// the SASS holds exactly the counted instructions (check: cuobjdump -sass tools/bin/probe_pipes | grep -c ...).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/probe_pipes tools/probe_pipes.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

enum { K_LOP3, K_SHF, K_PRMT, K_IMAD, K_IMADHI, K_IMADWIDE, K_DP4A, K_SEL, K_IADD3, K_SHL, K_SHR, K_LOP_IMAD, K_PRMT_IMAD, K_2ALU_1IMAD,
       K_1ALU_2IMAD, K_LOP_IMADHI, K_FFMA, K_LOP_FFMA, K_LOP_IMAD_FFMA, K_BFE, K_POPC, K_BREV, K_SHF_IMAD, K_LOP_DP4A, K_IMAD_DP4A, K_COUNT };
static const char* kNames[K_COUNT] = {"lop3", "shf.l.wrap", "prmt", "mad.lo", "mad.hi", "mad.wide", "dp4a", "selp", "add3", "shl", "shr", "lop3+mad.lo",
                                      "prmt+mad.lo", "2 lop3+mad.lo", "lop3+2 mad.lo", "lop3+mad.hi", "ffma", "lop3+ffma", "lop3+mad.lo+ffma", "bfe", "popc", "brev",
                                      "shf+mad.lo", "lop3+dp4a", "mad.lo+dp4a"};
static const int kOpsPerStep[K_COUNT] = {1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 2, 2, 3, 3, 2, 1, 2, 3, 1, 1, 1, 2, 2, 2};

#define LOP3(x, y, z) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x) : "r"(y), "r"(z))
#define IMAD(x, y, z) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(y), "r"(z))
#define IMADHI(x, y, z) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(y), "r"(z))
#define FFMA(x, y, z) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x) : "f"(y), "f"(z))

template <int KIND> __global__ void __launch_bounds__(1024) probe(uint32_t* out, uint32_t iters, uint32_t seed)
{
    uint32_t a[8];
    float f[8];
    unsigned long long w[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { a[i] = seed + threadIdx.x * 8u + i; f[i] = (float)a[i]; w[i] = a[i]; }
    const uint32_t k1 = seed | 1u, k2 = seed * 3u + 7u;
    const float g1 = 1.0001f, g2 = 0.5f;
    for (uint32_t it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 8; r++) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                if (KIND == K_LOP3) LOP3(a[i], k1, k2);
                if (KIND == K_SHF) asm volatile("shf.l.wrap.b32 %0, %0, %1, 7;" : "+r"(a[i]) : "r"(k1));
                if (KIND == K_PRMT) asm volatile("prmt.b32 %0, %0, %1, 0x3172;" : "+r"(a[i]) : "r"(k1));
                if (KIND == K_IMAD) IMAD(a[i], k1, k2);
                if (KIND == K_IMADHI) IMADHI(a[i], k1, k2);
                if (KIND == K_IMADWIDE) asm volatile("{.reg .b32 lo, hi; mov.b64 {lo, hi}, %0; mad.wide.u32 %0, lo, %1, %0;}" : "+l"(w[i]) : "r"(k1));
                if (KIND == K_DP4A) asm volatile("dp4a.u32.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(k1), "r"(k2));
                if (KIND == K_SEL) asm volatile("{.reg .pred p; setp.ne.u32 p, %1, 0; selp.u32 %0, %0, %2, p;}" : "+r"(a[i]) : "r"(k1), "r"(k2));
                if (KIND == K_IADD3) asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(a[(i + 1) & 7]));
                if (KIND == K_SHL) asm volatile("shl.b32 %0, %0, 3;" : "+r"(a[i]));
                if (KIND == K_SHR) asm volatile("shr.u32 %0, %0, 3;" : "+r"(a[i]));
                if (KIND == K_LOP_IMAD) { LOP3(a[i], k1, k2); IMAD(a[i], k1, k2); }
                if (KIND == K_PRMT_IMAD) { asm volatile("prmt.b32 %0, %0, %1, 0x3172;" : "+r"(a[i]) : "r"(k1)); IMAD(a[i], k1, k2); }
                if (KIND == K_2ALU_1IMAD) { LOP3(a[i], k1, k2); LOP3(a[i], k2, k1); IMAD(a[i], k1, k2); }
                if (KIND == K_1ALU_2IMAD) { LOP3(a[i], k1, k2); IMAD(a[i], k2, k1); IMAD(a[i], k1, k2); }
                if (KIND == K_LOP_IMADHI) { LOP3(a[i], k1, k2); IMADHI(a[i], k1, k2); }
                if (KIND == K_FFMA) FFMA(f[i], g1, g2);
                if (KIND == K_LOP_FFMA) { LOP3(a[i], k1, k2); FFMA(f[i], g1, g2); }
                if (KIND == K_LOP_IMAD_FFMA) { LOP3(a[i], k1, k2); IMAD(a[i], k1, k2); FFMA(f[i], g1, g2); }
                if (KIND == K_BFE) asm volatile("bfe.u32 %0, %0, 5, 9;" : "+r"(a[i]));
                if (KIND == K_POPC) asm volatile("popc.b32 %0, %0;" : "+r"(a[i]));
                if (KIND == K_BREV) asm volatile("brev.b32 %0, %0;" : "+r"(a[i]));
                if (KIND == K_LOP_DP4A) { LOP3(a[i], k1, k2); asm volatile("dp4a.u32.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(k1), "r"(k2)); }
                if (KIND == K_IMAD_DP4A) { IMAD(a[i], k1, k2); asm volatile("dp4a.u32.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(k1), "r"(k2)); }
                if (KIND == K_SHF_IMAD) { asm volatile("shf.l.wrap.b32 %0, %0, %1, 7;" : "+r"(a[i]) : "r"(k1)); IMAD(a[i], k1, k2); }
            }
        }
    }
    uint32_t x = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) x ^= a[i] ^ (uint32_t)w[i] ^ (uint32_t)(w[i] >> 32) ^ __float_as_uint(f[i]);
    if (x == 0x12345u) out[0] = x;
}

template <int KIND> static void run(uint32_t* d, int sms, double ghz)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const uint32_t iters = 2048;
    const unsigned grid = (unsigned)sms * 2;
    double best = 0;
    for (int rep = 0; rep < 4; rep++) {
        cudaEventRecord(e0);
        probe<KIND><<<grid, 1024>>>(d, iters, 12345u + rep);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        const double ops = (double)grid * 1024.0 * iters * 64 * kOpsPerStep[KIND];
        const double tops = ops / (ms * 1e-3) / 1e12;
        if (tops > best) best = tops;
    }
    printf("%-20s %7.2f Tops/s  %6.1f thread-ops/clk/SM at %.3f GHz\n", kNames[KIND], best, best * 1e12 / (sms * ghz * 1e9), ghz);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
}

template <int K> struct All { static void go(uint32_t* d, int sms, double ghz) { run<K>(d, sms, ghz); All<K + 1>::go(d, sms, ghz); } };
template <> struct All<K_COUNT> { static void go(uint32_t*, int, double) {} };

int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    uint32_t* d;
    cudaMalloc(&d, 4);
    printf("%s, %d SMs, max clock %.3f GHz\n", p.name, p.multiProcessorCount, khz * 1e-6);
    All<0>::go(d, p.multiProcessorCount, khz * 1e-6);
    return 0;
}
