// select2.cu -- does tracing TWO rays per thread through the uniform-operand sphere loop pay?  Times the selection
// loop of brute_select_pkc alone (56 spheres in the kernel parameter block) with one ray and with two rays per thread,
// at the occupancy each variant would have in the path kernel.  Prints ns per ray and group of 8 spheres.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef unsigned long long f32x2;
#define DEV __device__ __forceinline__
DEV f32x2 pack2(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
DEV void unpack2(f32x2 v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
DEV f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
DEV f32x2 add2(f32x2 a, f32x2 b) { f32x2 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
DEV float sqrta(float x) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
struct Pk { ulonglong2 q[64]; int n_padded, mask; };
struct Ray { float ox, oy, oz, dx, dy, dz; };
struct RC { f32x2 Dx, Dy, Dz, nod, Bx, By, Bz, noo; };
DEV RC prep(const Ray &r) {
    RC c; const float od = r.ox * r.dx + r.oy * r.dy + r.oz * r.dz, oo = r.ox * r.ox + r.oy * r.oy + r.oz * r.oz;
    c.Dx = pack2(r.dx, r.dx); c.Dy = pack2(r.dy, r.dy); c.Dz = pack2(r.dz, r.dz); c.nod = pack2(-od, -od);
    c.Bx = pack2(2 * r.ox, 2 * r.ox); c.By = pack2(2 * r.oy, 2 * r.oy); c.Bz = pack2(2 * r.oz, 2 * r.oz); c.noo = pack2(-oo, -oo);
    return c;
}
DEV void pairkeys(const ulonglong2 a, const ulonglong2 b, const RC &c, int mask, int i0, int &k0, int &k1) {
    const f32x2 neg1 = pack2(-1.f, -1.f);
    const f32x2 tca = fma2(b.x, c.Dz, fma2(a.y, c.Dy, fma2(a.x, c.Dx, c.nod)));
    const f32x2 nm = fma2(b.x, c.Bz, fma2(a.y, c.By, fma2(a.x, c.Bx, add2(b.y, c.noo))));
    const f32x2 disc = fma2(tca, tca, nm);
    float tc0, tc1, d0, d1; unpack2(tca, tc0, tc1); unpack2(disc, d0, d1);
    const float s0 = sqrta(__int_as_float(__float_as_int(d0) | (__float_as_int(tc0) & (int)0x80000000)));
    const float s1 = sqrta(__int_as_float(__float_as_int(d1) | (__float_as_int(tc1) & (int)0x80000000)));
    float t0, t1; unpack2(fma2(pack2(s0, s1), neg1, tca), t0, t1);
    k0 = (__float_as_int(t0) & mask) | i0; k1 = (__float_as_int(t1) & mask) | (i0 + 1);
}
template <int kRays, int kMinBlocks>
__global__ void __launch_bounds__(256, kMinBlocks) k(const __grid_constant__ Pk pk, const Ray *rays, int iters, int *out) {
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    Ray r[kRays];
#pragma unroll
    for (int q = 0; q < kRays; ++q) r[q] = rays[(tid * kRays + q) & 0xfffff];
    int acc = 0;
    for (int it = 0; it < iters; ++it) {
        RC c[kRays]; int best[kRays];
#pragma unroll
        for (int q = 0; q < kRays; ++q) { c[q] = prep(r[q]); best[q] = 0x7f800000; }
#pragma unroll
        for (int base = 0; base < 64; base += 8) {
            if (base >= pk.n_padded) break;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const ulonglong2 a = pk.q[base + 2 * j], b = pk.q[base + 2 * j + 1];
#pragma unroll
                for (int q = 0; q < kRays; ++q) {
                    int k0, k1; pairkeys(a, b, c[q], pk.mask, base + 2 * j, k0, k1);
                    best[q] = __vimin3_s32(best[q], k0, k1);
                }
            }
        }
#pragma unroll
        for (int q = 0; q < kRays; ++q) {       // make the next iteration depend on the result (like a bounce)
            acc += best[q];
            const float e = 1e-3f * (float)(best[q] & 63);
            r[q].ox += e; r[q].dy -= e * 0.01f;
        }
    }
    out[tid] = acc;
}
template <int kRays, int kMinBlocks> void run(const Pk &pk, const Ray *rays, int *out, int sms) {
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k<kRays, kMinBlocks>, 256, 0);
    cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, k<kRays, kMinBlocks>);
    const int grid = sms * per_sm, iters = 2000 / kRays;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<kRays, kMinBlocks><<<grid, 256>>>(pk, rays, iters, out); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0); k<kRays, kMinBlocks><<<grid, 256>>>(pk, rays, iters, out); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); best = ms < best ? ms : best;
    }
    const double ray_groups = (double)grid * 256 * kRays * iters * (pk.n_padded / 8);
    printf("rays/thread %d  min blocks %d: %3d regs, %d CTAs/SM, %.3f ms, %.4f ns per warp-ray-group/SM-subpartition-equivalent, %.2f G sphere tests/s\n",
           kRays, kMinBlocks, fa.numRegs, per_sm, best, best * 1e6 / (ray_groups / 32 / (sms * 4)), ray_groups * 8 / best / 1e6);
}
int main() {
    Pk pk; pk.n_padded = 56; pk.mask = 0x7fffffc0;
    float *f = reinterpret_cast<float *>(pk.q);
    srand(1);
    for (int i = 0; i < 64 * 4; ++i) f[i] = (float)rand() / RAND_MAX * 8.f - 4.f;
    Ray *h = new Ray[1 << 20], *d; int *out;
    for (int i = 0; i < (1 << 20); ++i) { float x = (float)rand() / RAND_MAX - .5f, y = (float)rand() / RAND_MAX - .5f, z = (float)rand() / RAND_MAX - .5f;
        float l = sqrtf(x * x + y * y + z * z) + 1e-6f; h[i] = Ray{(float)rand() / RAND_MAX * 4 - 2, (float)rand() / RAND_MAX * 4 - 2, (float)rand() / RAND_MAX * 4, x / l, y / l, z / l}; }
    cudaMalloc(&d, sizeof(Ray) << 20); cudaMemcpy(d, h, sizeof(Ray) << 20, cudaMemcpyHostToDevice);
    cudaMalloc(&out, 4 * 148 * 8 * 256 * 4);
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    run<1, 4>(pk, d, out, sms); run<1, 3>(pk, d, out, sms); run<1, 2>(pk, d, out, sms);
    run<2, 4>(pk, d, out, sms); run<2, 3>(pk, d, out, sms); run<2, 2>(pk, d, out, sms); run<2, 1>(pk, d, out, sms);
    run<4, 2>(pk, d, out, sms); run<4, 1>(pk, d, out, sms);
    return 0;
}
