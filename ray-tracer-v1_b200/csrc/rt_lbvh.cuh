// rt_lbvh.cuh -- device view of the on-device LBVH (built by rt_lbvh_build, see rt_lbvh.cu).
//
// Layout (Aila-Laine style, one 64-byte record per internal node, both child boxes in the parent):
//   node4[4*i+0] = left  box  (min.x, max.x, min.y, max.y)
//   node4[4*i+1] = right box  (min.x, max.x, min.y, max.y)
//   node4[4*i+2] = (left.min.z, left.max.z, right.min.z, right.max.z)
//   node4[4*i+3] = (left child, right child, -, -) as int bits; child >= 0 internal node, < 0 leaf ~k where
//                  prims[k] is the scene index of the sphere
// Boxes are float for both precisions and grown outward at build time, so they only ever cull; hit/miss and
// distances always come from the sphere test of the chosen precision.
#pragma once
#include <cuda_runtime.h>

namespace rt {

#define RT_BVH_STACK 64

struct BvhView {
    int nodes;              // number of primitives in the hierarchy; 0 = no BVH, brute force
    int root;               // >= 0 internal node, < 0 leaf (single primitive)
    const float4 *node4;
    const int *prims;
    int n_huge;             // spheres kept out of the hierarchy (walls), tested first
    const int *huge;
};

}  // namespace rt
