// rt_lbvh.cu -- on-device LBVH over the scene's spheres (Karras, "Maximizing Parallelism in the Construction of
// BVHs, Octrees, and k-d Trees", HPG 2012): 30-bit Morton codes of the sphere centres (index appended so keys are
// unique) -> radix sort -> one thread per internal node finds its range and split -> bottom-up AABB refit with one
// atomic arrival counter per node.  Used for the "many small spheres" scenes (SURVEY.md 8d, C4 scaled variant); the
// reference itself only ever loops over every sphere (RL/ray.py:164-166), the hierarchy is a pure culling structure.
#include <cub/device/device_radix_sort.cuh>

#include "rt_lbvh_build.h"

namespace rt {

// order-preserving float <-> int mapping for atomicMin/atomicMax
__device__ __forceinline__ int f2ord(float f) { int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7fffffff; }
__device__ __forceinline__ float ord2f(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

struct BuildHeader {
    int lo[3], hi[3];   // centroid bounds of the regular spheres (ordered-int encoded)
    int n_huge;
};

__global__ void lbvh_init_kernel(BuildHeader *h) {
    for (int k = 0; k < 3; ++k) { h->lo[k] = 0x7fffffff; h->hi[k] = (int)0x80000000; }
    h->n_huge = 0;
}

__global__ void lbvh_bounds_kernel(const float4 *sph, int n, float huge_radius, BuildHeader *h) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    int huge = 0;
    if (i < n) {
        const float4 s = sph[i];
        if (s.w > huge_radius) huge = 1;
        else { lo[0] = hi[0] = s.x; lo[1] = hi[1] = s.y; lo[2] = hi[2] = s.z; }
    }
    for (int o = 16; o > 0; o >>= 1) {
        for (int k = 0; k < 3; ++k) {
            lo[k] = fminf(lo[k], __shfl_down_sync(0xffffffffu, lo[k], o));
            hi[k] = fmaxf(hi[k], __shfl_down_sync(0xffffffffu, hi[k], o));
        }
        huge += __shfl_down_sync(0xffffffffu, huge, o);
    }
    if ((threadIdx.x & 31) == 0) {
        for (int k = 0; k < 3; ++k) {
            if (lo[k] <= hi[k]) { atomicMin(&h->lo[k], f2ord(lo[k])); atomicMax(&h->hi[k], f2ord(hi[k])); }
        }
        if (huge) atomicAdd(&h->n_huge, huge);
    }
}

__device__ __forceinline__ unsigned expand_bits10(unsigned v) {
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}

__global__ void lbvh_morton_kernel(const float4 *sph, int n, float huge_radius, const BuildHeader *h,
                                   unsigned long long *keys) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 s = sph[i];
    if (s.w > huge_radius) { keys[i] = (0xffffffffull << 32) | (unsigned)i; return; }
    const float c[3] = {s.x, s.y, s.z};
    unsigned q[3];
    for (int k = 0; k < 3; ++k) {
        const float lo = ord2f(h->lo[k]), hi = ord2f(h->hi[k]);
        const float ext = hi - lo;
        float u = ext > 0.f ? (c[k] - lo) / ext : 0.f;
        u = fminf(fmaxf(u * 1024.f, 0.f), 1023.f);
        q[k] = (unsigned)u;
    }
    const unsigned code = (expand_bits10(q[0]) << 2) | (expand_bits10(q[1]) << 1) | expand_bits10(q[2]);
    keys[i] = ((unsigned long long)code << 32) | (unsigned)i;
}

__global__ void lbvh_prims_kernel(const unsigned long long *keys, int n, int *prims) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) prims[i] = (int)(unsigned)(keys[i] & 0xffffffffull);
}

__device__ __forceinline__ int delta(const unsigned long long *keys, int n, int i, int j) {
    if (j < 0 || j >= n) return -1;
    return __clzll((long long)(keys[i] ^ keys[j]));      // keys are unique (index in the low word)
}

// one thread per internal node (Karras 2012, section 4)
__global__ void lbvh_hierarchy_kernel(const unsigned long long *keys, int n, float4 *node4, int *parent_internal,
                                      int *parent_leaf) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    const int d = (delta(keys, n, i, i + 1) - delta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
    const int dmin = delta(keys, n, i, i - d);
    int lmax = 2;
    while (delta(keys, n, i, i + lmax * d) > dmin) lmax <<= 1;
    int l = 0;
    for (int t = lmax >> 1; t >= 1; t >>= 1)
        if (delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
    const int j = i + l * d;
    const int dnode = delta(keys, n, i, j);
    int s = 0;
    for (int t = (l + 1) >> 1;; t = (t + 1) >> 1) {
        if (delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
        if (t == 1) break;
    }
    const int gamma = i + s * d + min(d, 0);
    const int first = min(i, j), last = max(i, j);
    const int left = (first == gamma) ? ~gamma : gamma;
    const int right = (last == gamma + 1) ? ~(gamma + 1) : gamma + 1;
    node4[4 * i + 3] = make_float4(__int_as_float(left), __int_as_float(right), 0.f, 0.f);
    if (left < 0) parent_leaf[~left] = i; else parent_internal[left] = i;
    if (right < 0) parent_leaf[~right] = i; else parent_internal[right] = i;
    if (i == 0) parent_internal[0] = -1;
}

// one thread per leaf walks up; the second thread to arrive at a node merges and continues
__global__ void lbvh_refit_kernel(const float4 *sph, const int *prims, int n, float4 *node4, const int *parent_internal,
                                  const int *parent_leaf, int *arrivals) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const float4 s = sph[prims[k]];
    // grow outward: the boxes must never cull a hit the (float or double) sphere test would report
    const float grow = s.w * 1e-5f + 1e-6f * (fabsf(s.x) + fabsf(s.y) + fabsf(s.z) + s.w) + 1e-7f;
    float lo[3] = {s.x - s.w - grow, s.y - s.w - grow, s.z - s.w - grow};
    float hi[3] = {s.x + s.w + grow, s.y + s.w + grow, s.z + s.w + grow};
    int me = ~k, cur = parent_leaf[k];
    while (cur >= 0) {
        const float4 ch = node4[4 * cur + 3];
        const bool is_left = __float_as_int(ch.x) == me;
        volatile float *zrow = reinterpret_cast<volatile float *>(&node4[4 * cur + 2]);
        if (is_left) { __stcg(&node4[4 * cur + 0], make_float4(lo[0], hi[0], lo[1], hi[1])); zrow[0] = lo[2]; zrow[1] = hi[2]; }
        else { __stcg(&node4[4 * cur + 1], make_float4(lo[0], hi[0], lo[1], hi[1])); zrow[2] = lo[2]; zrow[3] = hi[2]; }
        __threadfence();
        if (atomicAdd(&arrivals[cur], 1) == 0) return;
        __threadfence();
        const float4 a = __ldcg(&node4[4 * cur + 0]), b = __ldcg(&node4[4 * cur + 1]);
        const float z0 = zrow[0], z1 = zrow[1], z2 = zrow[2], z3 = zrow[3];
        lo[0] = fminf(a.x, b.x); hi[0] = fmaxf(a.y, b.y);
        lo[1] = fminf(a.z, b.z); hi[1] = fmaxf(a.w, b.w);
        lo[2] = fminf(z0, z2); hi[2] = fmaxf(z1, z3);
        me = cur;
        cur = parent_internal[cur];
    }
}

void lbvh_drop(LbvhStorage &s) {
    if (s.nodes) cudaFree(s.nodes);
    if (s.prims) cudaFree(s.prims);
    s.nodes = s.prims = nullptr;
    s.view = BvhView{0, 0, nullptr, nullptr, 0, nullptr};
}

#define LB(call)                               \
    do {                                       \
        err = (call);                          \
        if (err != cudaSuccess) goto done;     \
    } while (0)

cudaError_t lbvh_build(LbvhStorage &s, const float4 *spheres, int n, float huge_radius, cudaStream_t st) {
    lbvh_drop(s);
    if (n <= 0) return cudaSuccess;
    cudaError_t err = cudaSuccess;
    BuildHeader *hdr = nullptr;
    unsigned long long *keys_in = nullptr, *keys_out = nullptr;
    void *tmp = nullptr;
    int *parent_internal = nullptr, *parent_leaf = nullptr, *arrivals = nullptr, *prims = nullptr;
    float4 *nodes = nullptr;
    size_t tmp_bytes = 0;
    BuildHeader host;
    int n_reg = 0;
    const int tb = 256, gb = (n + tb - 1) / tb;
    LB(cudaMalloc((void **)&hdr, sizeof(BuildHeader)));
    LB(cudaMalloc((void **)&keys_in, sizeof(unsigned long long) * (size_t)n));
    LB(cudaMalloc((void **)&keys_out, sizeof(unsigned long long) * (size_t)n));
    LB(cudaMalloc((void **)&prims, sizeof(int) * (size_t)n));
    lbvh_init_kernel<<<1, 1, 0, st>>>(hdr);
    lbvh_bounds_kernel<<<gb, tb, 0, st>>>(spheres, n, huge_radius, hdr);
    lbvh_morton_kernel<<<gb, tb, 0, st>>>(spheres, n, huge_radius, hdr, keys_in);
    LB(cudaGetLastError());
    LB(cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, keys_in, keys_out, n, 0, 64, st));
    LB(cudaMalloc(&tmp, tmp_bytes ? tmp_bytes : 1));
    LB(cub::DeviceRadixSort::SortKeys(tmp, tmp_bytes, keys_in, keys_out, n, 0, 64, st));
    lbvh_prims_kernel<<<gb, tb, 0, st>>>(keys_out, n, prims);
    LB(cudaGetLastError());
    LB(cudaMemcpyAsync(&host, hdr, sizeof host, cudaMemcpyDeviceToHost, st));
    LB(cudaStreamSynchronize(st));
    n_reg = n - host.n_huge;
    if (n_reg >= 1) {
        if (n_reg >= 2) {
            LB(cudaMalloc((void **)&nodes, sizeof(float4) * 4 * (size_t)(n_reg - 1)));
            LB(cudaMalloc((void **)&parent_internal, sizeof(int) * (size_t)n_reg));
            LB(cudaMalloc((void **)&parent_leaf, sizeof(int) * (size_t)n_reg));
            LB(cudaMalloc((void **)&arrivals, sizeof(int) * (size_t)n_reg));
            LB(cudaMemsetAsync(arrivals, 0, sizeof(int) * (size_t)n_reg, st));
            LB(cudaMemsetAsync(nodes, 0, sizeof(float4) * 4 * (size_t)(n_reg - 1), st));
            const int gr = (n_reg + tb - 1) / tb;
            lbvh_hierarchy_kernel<<<gr, tb, 0, st>>>(keys_out, n_reg, nodes, parent_internal, parent_leaf);
            lbvh_refit_kernel<<<gr, tb, 0, st>>>(spheres, prims, n_reg, nodes, parent_internal, parent_leaf, arrivals);
            LB(cudaGetLastError());
            LB(cudaStreamSynchronize(st));
        }
        s.nodes = nodes; nodes = nullptr;
        s.prims = prims; prims = nullptr;
        s.view.nodes = n_reg;
        s.view.root = n_reg >= 2 ? 0 : ~0;
        s.view.node4 = reinterpret_cast<const float4 *>(s.nodes);
        s.view.prims = reinterpret_cast<const int *>(s.prims);
        s.view.n_huge = host.n_huge;
        s.view.huge = s.view.prims + n_reg;
    }
done:
    if (hdr) cudaFree(hdr);
    if (keys_in) cudaFree(keys_in);
    if (keys_out) cudaFree(keys_out);
    if (tmp) cudaFree(tmp);
    if (parent_internal) cudaFree(parent_internal);
    if (parent_leaf) cudaFree(parent_leaf);
    if (arrivals) cudaFree(arrivals);
    if (prims) cudaFree(prims);
    if (nodes) cudaFree(nodes);
    return err;
}

}  // namespace rt
