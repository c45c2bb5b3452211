// rt_common.cuh -- shared device-side vocabulary: vectors, math traits, Philox, scene views.
//
// Two instantiations of everything in rt_trace.cuh exist:
//   T = float   the product path ("fast": FMA-contracted, rsqrt/approx-sqrt, winner-only hit point)
//   T = double  the parity build ("exact": compiled with -fmad=false in rt_f64.cu and written in the
//               reference's operation order so it tracks the Python/IEEE-double arithmetic to ~1 ulp)
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>
#include "rt_lbvh.cuh"

namespace rt {

#define RT_DEV __device__ __forceinline__
// RT_CHECKED build (csrc/Makefile target `checked` -> librt_b200_checked.so): device-side bounds / protocol assertions on
// every indexed access of the kernels.  compute-sanitizer is closed on the pool these kernels are developed on, so the
// GPU test suite is also run against this library (RT_B200_LIB=...): a violated assertion aborts the launch with
// cudaErrorAssert and every later call of the test fails loudly.
#ifdef RT_CHECKED
#include <assert.h>
#define RT_ASSERT(cond) assert(cond)
#else
#define RT_ASSERT(cond) ((void)0)
#endif
#define RT_NO_ID_DEV INT32_MIN
#define RT_PATH_MAX_DEPTH 32     /* deepest TraditionalRenderer recursion the path kernel unrolls */

// ------------------------------------------------------------------ math traits
template <typename T> struct M;
template <> struct M<float> {
    static constexpr bool exact = false;
    using v4 = float4;
    // single MUFU each, denormals flushed (no range-fixup code around the MUFU)
    static RT_DEV float sqrt(float x) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
    static RT_DEV float rsqrt(float x) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
    static RT_DEV float acos(float x) { return acosf(fminf(1.f, fmaxf(-1.f, x))); }
    static RT_DEV float fabs(float x) { return fabsf(x); }
    static RT_DEV float rint(float x) { return rintf(x); }
    static RT_DEV float trunc(float x) { return truncf(x); }
    static RT_DEV float tan(float x) { return tanf(x); }
    static RT_DEV void sincos(float x, float *s, float *c) { sincosf(x, s, c); }
    static RT_DEV float inf() { return CUDART_INF_F; }
    static RT_DEV v4 make4(float x, float y, float z, float w) { return make_float4(x, y, z, w); }
};
template <> struct M<double> {
    static constexpr bool exact = true;
    using v4 = double4;
    static RT_DEV double sqrt(double x) { return ::sqrt(x); }
    static RT_DEV double rsqrt(double x) { return 1.0 / ::sqrt(x); }
    static RT_DEV double acos(double x) { return ::acos(x); }
    static RT_DEV double fabs(double x) { return ::fabs(x); }
    static RT_DEV double rint(double x) { return ::rint(x); }
    static RT_DEV double trunc(double x) { return ::trunc(x); }
    static RT_DEV double tan(double x) { return ::tan(x); }
    static RT_DEV void sincos(double x, double *s, double *c) { *s = ::sin(x); *c = ::cos(x); }
    static RT_DEV double inf() { return CUDART_INF; }
    static RT_DEV v4 make4(double x, double y, double z, double w) { return make_double4(x, y, z, w); }
};

// ------------------------------------------------------------------ 3-vectors (vector.py semantics)
template <typename T> struct V3 { T x, y, z; };
template <typename T> RT_DEV V3<T> mk(T x, T y, T z) { V3<T> r; r.x = x; r.y = y; r.z = z; return r; }
template <typename T> RT_DEV V3<T> operator+(V3<T> a, V3<T> b) { return mk<T>(a.x + b.x, a.y + b.y, a.z + b.z); }
template <typename T> RT_DEV V3<T> operator-(V3<T> a, V3<T> b) { return mk<T>(a.x - b.x, a.y - b.y, a.z - b.z); }
template <typename T> RT_DEV V3<T> operator-(V3<T> a) { return mk<T>(-a.x, -a.y, -a.z); }
template <typename T> RT_DEV V3<T> operator*(V3<T> a, T s) { return mk<T>(a.x * s, a.y * s, a.z * s); }
template <typename T> RT_DEV T dot(V3<T> a, V3<T> b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
template <typename T> RT_DEV V3<T> cross(V3<T> a, V3<T> b) {
    return mk<T>(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
template <typename T> RT_DEV T mag(V3<T> a) { return M<T>::sqrt(dot(a, a)); }
// Vector.normalise (vector.py:110-112): exact = three divisions by the magnitude; fast = one rsqrt
template <typename T> RT_DEV V3<T> normalise(V3<T> a) {
    if constexpr (M<T>::exact) { T m = ::sqrt(dot(a, a)); return mk<T>(a.x / m, a.y / m, a.z / m); }
    else { T i = M<T>::rsqrt(dot(a, a)); return a * i; }
}
// Vector.angleBetween (vector.py:61-62).  fast: clamps the cosine (the reference would produce NaN -> raise)
template <typename T> RT_DEV T angle_between(V3<T> a, V3<T> b) {
    if constexpr (M<T>::exact) return ::acos(dot(a, b) / (::sqrt(dot(a, a)) * ::sqrt(dot(b, b))));
    else return M<T>::acos(dot(a, b) * M<T>::rsqrt(dot(a, a) * dot(b, b)));
}
// Vector.reflectInVector (vector.py:64-67).  `self` and `n` are unit vectors on every call site of the hot path;
// exact re-normalises like the reference, fast normalises the result only.
template <typename T> RT_DEV V3<T> reflect(V3<T> self, V3<T> B) {
    if constexpr (M<T>::exact) {
        V3<T> v = normalise(self), n = normalise(B);
        return normalise(v - n * (T(2) * dot(v, n)));
    } else {
        return normalise(self - B * (T(2) * dot(self, B)));
    }
}
// Vector.refractInVector (vector.py:69-92); false = total internal reflection
template <typename T> RT_DEV bool refract(V3<T> self, V3<T> B, T ra, T rb, V3<T> &out) {
    V3<T> v = self, nrm = B;
    if constexpr (M<T>::exact) { v = normalise(self); nrm = normalise(B); }
    T n = ra / rb;
    T cosI = dot(v, nrm);
    if (cosI < T(-1)) cosI = T(-1);
    if (cosI > T(1)) cosI = T(1);
    if (cosI < T(0)) cosI = -cosI;
    T k = T(1) - (n * n) * (T(1) - cosI * cosI);
    if (k < T(0)) return false;
    out = normalise(v * n + nrm * (n * cosI - M<T>::sqrt(k)));
    return true;
}
// Vector.rotate (vector.py:117-127): row vector times R(angle)
template <typename T> RT_DEV V3<T> rotate(V3<T> s, V3<T> ang) {
    T sa, ca, sb, cb, sc, cc;
    M<T>::sincos(ang.x, &sa, &ca); M<T>::sincos(ang.y, &sb, &cb); M<T>::sincos(ang.z, &sc, &cc);
    T r00 = cc * cb * ca - sc * sa, r01 = cc * cb * sa + sc * ca, r02 = -cc * sb;
    T r10 = -sc * cb * ca - cc * sa, r11 = -sc * cb * sa + cc * ca, r12 = sc * sb;
    T r20 = sb * ca, r21 = sb * sa, r22 = cb;
    return mk<T>(s.x * r00 + s.y * r10 + s.z * r20, s.x * r01 + s.y * r11 + s.z * r21,
                 s.x * r02 + s.y * r12 + s.z * r22);
}

// ------------------------------------------------------------------ Philox4x32-10
// Counter-based RNG keyed (pixel, sample, slot>>1); words 2*(slot&1)+{0,1} feed `slot`
// (slot 0 = camera jitter, slot k+1 = diffuse bounce at depth k).  Identical to oracle/rt_oracle.c.
struct Philox4 { uint32_t w[4]; };
RT_DEV Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
        uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
        uint32_t n0 = h1 ^ c1 ^ k0, n2 = h0 ^ c3 ^ k1;
        c0 = n0; c1 = l1; c2 = n2; c3 = l0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    Philox4 o; o.w[0] = c0; o.w[1] = c1; o.w[2] = c2; o.w[3] = c3;
    return o;
}
// the same rounds with the ten round keys precomputed (kernel parameter block: they become constant-bank operands)
RT_DEV Philox4 philox4x32_10_rk(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, const uint32_t (&rk)[20]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
        uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
        uint32_t n0 = h1 ^ c1 ^ rk[2 * r], n2 = h0 ^ c3 ^ rk[2 * r + 1];
        c0 = n0; c1 = l1; c2 = n2; c3 = l0;
    }
    Philox4 o; o.w[0] = c0; o.w[1] = c1; o.w[2] = c2; o.w[3] = c3;
    return o;
}
#define RT_PHILOX_TAG 0x52544232u /* "RTB2" */
template <typename T> RT_DEV T u01(uint32_t w) { return T(w >> 8) * T(1.0 / 16777216.0); }   // 24-bit, exact in f32

// Per-path RNG with a two-slot cache: one Philox call serves two consecutive slots.
struct PathRng {
    uint32_t pixel, sample, k0, k1;
    uint32_t c2, c3;       // cached words 2,3
    uint32_t cached_pair;  // pair index the cache belongs to (0xffffffff = none)
    RT_DEV void begin(uint32_t pix, uint32_t smp, uint32_t key0, uint32_t key1) {
        pixel = pix; sample = smp; k0 = key0; k1 = key1; cached_pair = 0xffffffffu;
    }
    RT_DEV void pair(uint32_t slot, uint32_t &a, uint32_t &b) {
        uint32_t pr = slot >> 1;
        if ((slot & 1u) && cached_pair == pr) { a = c2; b = c3; return; }
        Philox4 o = philox4x32_10(pixel, sample, pr, RT_PHILOX_TAG, k0, k1);
        if (slot & 1u) { a = o.w[2]; b = o.w[3]; }
        else { a = o.w[0]; b = o.w[1]; c2 = o.w[2]; c3 = o.w[3]; cached_pair = pr; }
    }
    RT_DEV void pair_rk(uint32_t slot, uint32_t &a, uint32_t &b, const uint32_t (&rk)[20]) {
        uint32_t pr = slot >> 1;
        if ((slot & 1u) && cached_pair == pr) { a = c2; b = c3; return; }
        Philox4 o = philox4x32_10_rk(pixel, sample, pr, RT_PHILOX_TAG, rk);
        if (slot & 1u) { a = o.w[2]; b = o.w[3]; }
        else { a = o.w[0]; b = o.w[1]; c2 = o.w[2]; c3 = o.w[3]; cached_pair = pr; }
    }
};

// ------------------------------------------------------------------ scene views
// Device-resident flattened scene for one precision.  vec4 packing:
//   sph   cx cy cz r            mat   reflective transparent emitive refractive_index
//   pk    sphere pairs (2j, 2j+1): (cx0 cx1 cy0 cy1), (cz0 cz1 w0 w1) with w = r^2 - |c|^2 (packed selection loop)
//   col   r g b 1/r             g_vec x y z max_angle     g_col r g b strength
//   p_pos x y z max_angle       p_col r g b strength      l_pos x y z -      l_col r g b -
//   lpk   light pairs (2j, 2j+1), RT_LPK_STRIDE = 4 vectors each: 128*(x0 x1 y0 y1), (128*z0 128*z1 R0 R1), (G0 G1 B0 B1),
//         (|A0|^2 |A1|^2 - -) with A = 128 * centre; RGB = colour*0.3*16384
template <typename T> struct SceneDev {
    using v4 = typename M<T>::v4;
    int n, nG, nP, nL;
    const v4 *sph, *pk, *mat, *col;
    const int *ids;
    const v4 *g_vec, *g_col; const int *g_func;
    const v4 *p_pos, *p_col; const int *p_id, *p_func;
    const v4 *l_pos, *l_col, *lpk; const int *l_index;
    const uint8_t *small;
    int key_mask;             // 0x7ffffff8, passed as DATA so that the selection loop's (t & mask) | k stays ONE LOP3
    int key_mask6;            // 0x7fffffc0: the same for brute_select_pkc (6 index bits)
    T bg[3];
    BvhView bvh;              // optional LBVH, see rt_lbvh.cuh
};

// Sphere pairs and light pairs (the FP32 `pk` / `lpk` arrays) of a small scene carried IN THE KERNEL PARAMETER BLOCK:
// the selection loop and the direct-light loop then read them through the constant bank into uniform registers
// (LDCU.128) and FFMA2 takes them as uniform operands, so no vector register, no register-file read port and no
// shared-memory load is spent on warp-uniform scene data.
#define RT_PKC_MAX 64                    /* spheres */
#define RT_LPKC_MAX 32                   /* Algorithm-B light spheres (the `lpk` light-pair array, 4 vectors per pair) */
#define RT_LPK_STRIDE 4                  /* float4 vectors per light pair */
struct PkConst { ulonglong2 q[RT_PKC_MAX]; ulonglong2 l[RT_LPK_STRIDE * RT_LPKC_MAX / 2]; };
struct PkNone {};

// sphere arrays as the tracing functions see them (shared-memory staged or global)
template <typename T> struct SphereView {
    using v4 = typename M<T>::v4;
    int n;
    int n_padded;            // sph[] readable up to here: n rounded up to 8 with NaN-radius spheres (never hit)
    int key_mask, key_mask6; // SceneDev::key_mask, key_mask6
    const v4 *sph, *pk, *mat, *col;
    const int *ids;
    const float4 *cw;        // path kernel kMode 3 only: (cx, cy, cz, w) per sphere, the pair array un-interleaved
};

enum : int { STAT_RAYS = 0, STAT_INTER = 1, STAT_LIGHT = 2, STAT_SMALL = 3, STAT_QUERIES = 4, STAT_SPHERE_TESTS = 5,
             STAT_AABB_TESTS = 6, STAT_DEAD_QUERIES = 7, STAT_COUNT = 8 };

RT_DEV unsigned long long warp_sum(unsigned long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace rt
