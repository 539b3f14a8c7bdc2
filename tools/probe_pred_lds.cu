// Tuning aid: latency of a dependent chain through shared-memory loads for one warp, with the load (a) always executed,
// (b) predicated off, (c) absent -- does a predicated-off LDS still cost its consumer the load latency?  (It does.)
// build here, run on the GPU box:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/probe_pred_lds tools/probe_pred_lds.cu
//                                  gpurun -- ./tools/bin/probe_pred_lds
#include <cstdio>
#include <cuda_runtime.h>
__global__ void probe(unsigned* out, int iters, int mode)
{
    __shared__ unsigned tab[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) tab[i] = (i * 7u + 3u) & 1023u;
    __syncthreads();
    unsigned x = threadIdx.x & 0u, on = mode == 0 ? 1u : 0u;
    const unsigned base = (unsigned)__cvta_generic_to_shared(tab);
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 8; k++) {
            unsigned e = 0;
            if (mode == 3) {                     // a warp-uniform, data-dependent branch around the load (taken about half the time)
                if (x & 1u) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(e) : "r"(base + ((x & 1023u) << 2)));
            } else if (mode == 4) {              // the same branch with nothing inside but an ALU op
                if (x & 1u) asm volatile("add.u32 %0, %1, 5;" : "=r"(e) : "r"(x));
            } else if (mode == 5) {              // load AND its consumer predicated off: does the consumer still wait for the load's scoreboard?
                asm volatile("{\n.reg .pred p;\n.reg .u32 t;\nsetp.ne.u32 p, %2, 0;\n@p ld.shared.u32 t, [%1];\n@p add.u32 %0, %0, t;\n}\n" : "+r"(x) : "r"(base + ((x & 1023u) << 2)), "r"(on));
            } else if (mode < 2) asm volatile("{\n.reg .pred p;\nsetp.ne.u32 p, %2, 0;\nmov.u32 %0, 0;\n@p ld.shared.u32 %0, [%1];\n}\n" : "=r"(e) : "r"(base + ((x & 1023u) << 2)), "r"(on));
            x = (x + e + 1u) & 1023u;           // dependent on the (possibly skipped) load
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) { out[0] = x; out[1] = (unsigned)((t1 - t0) / (iters * 8)); }
}
int main()
{
    unsigned* d; cudaMalloc(&d, 8);
    for (int mode = 0; mode < 6; mode++) {
        probe<<<1, 32>>>(d, 20000, mode);
        unsigned h[2]; cudaMemcpy(h, d, 8, cudaMemcpyDeviceToHost);
        printf("%s: %u cycles per link\n", mode == 0 ? "LDS executed" : mode == 1 ? "LDS predicated off" : mode == 2 ? "no LDS" : mode == 3 ? "LDS under a uniform branch (about half taken)" : mode == 4 ? "uniform branch around an ALU op" : "LDS and its consumer both predicated off", h[1]);
    }
    return 0;
}
