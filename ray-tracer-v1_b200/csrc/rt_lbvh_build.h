// rt_lbvh_build.h -- host interface of the on-device LBVH builder (rt_lbvh.cu).
#pragma once
#include "rt_lbvh.cuh"

namespace rt {

struct LbvhStorage {
    BvhView view = {0, 0, nullptr, nullptr, 0, nullptr};
    void *nodes = nullptr;      // float4[4 * (n_reg - 1)]
    void *prims = nullptr;      // int[n]: Morton-sorted regular spheres, then the huge ones
};

// Morton codes -> radix sort -> Karras hierarchy -> bottom-up AABB refit, all on `st`; synchronises once to learn how
// many spheres are "huge" (radius > huge_radius: kept out of the hierarchy, tested brute force).
cudaError_t lbvh_build(LbvhStorage &s, const float4 *spheres, int n, float huge_radius, cudaStream_t st);
void lbvh_drop(LbvhStorage &s);

}  // namespace rt
