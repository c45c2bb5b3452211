// pipes.cu -- issue-rate micro-benchmarks for the instructions the sphere loop is made of (B200, sm_100a):
// FFMA, FFMA2 (fma.rn.f32x2), MUFU.SQRT, LOP3, VIMNMX3.  Prints warp-instructions per cycle per SM sub-partition.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 4096

__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}

template <int kKind> __global__ void __launch_bounds__(1024) k(float *out, float seed, long long *cycles) {
    float a[8];
    unsigned long long p[8];
    int q[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        a[i] = seed + i + threadIdx.x;
        p[i] = ((unsigned long long)__float_as_uint(a[i]) << 32) | __float_as_uint(a[i] * 0.5f);
        q[i] = __float_as_int(a[i]);
    }
    const unsigned long long m2 = ((unsigned long long)__float_as_uint(0.999f) << 32) | __float_as_uint(0.999f);
    unsigned long long g0; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(g0));
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (kKind == 0) a[i] = fmaf(a[i], 0.999f, seed);
            if (kKind == 11) a[i] = fmaf(a[i], a[(i + 1) & 7], seed);
            if (kKind == 1) p[i] = fma2(p[i], m2, p[(i + 1) & 7]);
            if (kKind == 2) asm volatile("sqrt.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
            if (kKind == 3) asm volatile("lop3.b32 %0, %0, %1, %2, 0xEA;" : "+r"(q[i]) : "r"(q[(i + 1) & 7]), "r"(q[(i + 3) & 7]));
            if (kKind == 4) q[i] = __vimin3_s32(q[i], q[(i + 1) & 7], q[(i + 3) & 7]);
            if (kKind == 5) { a[i] = fmaf(a[i], 0.999f, seed); asm volatile("sqrt.approx.ftz.f32 %0, %0;" : "+f"(a[(i + 4) & 7])); }
            if (kKind == 6) { a[i] = fmaf(a[i], 0.999f, seed); asm volatile("lop3.b32 %0, %0, %1, %2, 0xEA;" : "+r"(q[i]) : "r"(q[(i + 1) & 7]), "r"(q[(i + 3) & 7])); }
            if (kKind == 7) { p[i] = fma2(p[i], m2, p[(i + 1) & 7]); asm volatile("lop3.b32 %0, %0, %1, %2, 0xEA;" : "+r"(q[i]) : "r"(q[(i + 1) & 7]), "r"(q[(i + 3) & 7])); }
            if (kKind == 8) { p[i] = fma2(p[i], m2, p[(i + 1) & 7]); a[i] = fmaf(a[i], 0.999f, seed); }
            if (kKind == 9) { p[i] = fma2(p[i], m2, p[(i + 1) & 7]); asm volatile("lop3.b32 %0, %0, %1, %2, 0xEA;" : "+r"(q[i]) : "r"(q[(i + 1) & 7]), "r"(q[(i + 3) & 7]));
                              if ((i & 3) == 0) asm volatile("sqrt.approx.ftz.f32 %0, %0;" : "+f"(a[i])); }
            if (kKind == 10) { a[i] = fmaf(a[i], 0.999f, seed); a[(i+4)&7] = fmaf(a[(i+4)&7], 0.998f, seed); asm volatile("lop3.b32 %0, %0, %1, %2, 0xEA;" : "+r"(q[i]) : "r"(q[(i + 1) & 7]), "r"(q[(i + 3) & 7])); }
        }
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i] + __uint_as_float((unsigned)p[i]) + __uint_as_float((unsigned)(p[i] >> 32)) + (float)q[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    unsigned long long g1; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(g1));
    if (threadIdx.x == 0 && blockIdx.x == 0) { cycles[0] = t1 - t0; cycles[1] = (long long)(g1 - g0); }
}

template <int kKind> void run(const char *name, int per_iter) {
    float *out; long long *cyc, h[2];
    cudaMalloc(&out, 148 * 1024 * sizeof(float)); cudaMalloc(&cyc, 16);
    k<kKind><<<148, 1024>>>(out, 1.0f, cyc);        // 1 CTA x 32 warps per SM = 8 warps per sub-partition
    cudaDeviceSynchronize();
    k<kKind><<<148, 1024>>>(out, 1.0f, cyc);
    cudaDeviceSynchronize();
    cudaMemcpy(h, cyc, 16, cudaMemcpyDeviceToHost);
    double warp_inst = (double)ITERS * per_iter * 8 /* warps per SMSP */;
    printf("%-28s %8lld cycles %8lld ns (%.0f MHz)  %.3f warp-inst/clk/SMSP  %.3f warp-inst/ns/SMSP\n", name, h[0], h[1],
           1e3 * (double)h[0] / (double)h[1], warp_inst / (double)h[0], warp_inst / (double)h[1]);
    cudaFree(out); cudaFree(cyc);
}

int main() {
    run<0>("FFMA imm", 8);
    run<11>("FFMA reg", 8);
    run<1>("FFMA2 (f32x2)", 8);
    run<2>("MUFU.SQRT", 8);
    run<3>("LOP3", 8);
    run<4>("VIMNMX3", 8);
    run<5>("FFMA + MUFU.SQRT interleaved", 16);
    run<6>("FFMA + LOP3 1:1", 16);
    run<7>("FFMA2 + LOP3 1:1", 16);
    run<8>("FFMA2 + FFMA 1:1", 16);
    run<9>("FFMA2 + LOP3 + MUFU/4", 18);
    run<10>("2 FFMA + LOP3", 24);
    return 0;
}
