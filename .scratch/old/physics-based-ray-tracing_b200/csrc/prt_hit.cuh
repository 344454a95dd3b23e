// prt_hit.cuh -- the two ray queries of the path, one ray per lane: closest_hit (== scene.ray_intersect,
// /root/reference/CustomIntegrator.py:146,309) and occluded (the connection ray, :159,324).  Used by the
// acquisition megakernel, the tile path tracer and the query kernels; the wavefront trace kernels
// (prt_wavefront.cu) run the same node / triangle tests with their own warp-level scheduling.
//
// Triangles go through the compressed 8-wide BVH (prt_bvh8.cuh) by default: one 80-byte node step replaces ~3
// dependent BVH2 levels and the traversal stack holds one 8-byte entry per level.  -DPRT_MEGA_BVH8=0 restores the
// binary while-while traversal (traverse_bvh in prt_device.cuh) for A/B measurements; hits are the same watertight
// test either way.
#pragma once
#include "prt_bvh8.cuh"
#include "prt_device.cuh"

#ifndef PRT_MEGA_BVH8
#define PRT_MEGA_BVH8 0
#endif

namespace prt {

// nearest (ANY = false) or any (ANY = true) triangle within [0, tbest]; returns the SORTED (LBVH-order) triangle index
// or -1, like traverse_bvh
template <bool ANY>
__device__ __forceinline__ int traverse_bvh8(const DScene &sc, float3 o, float3 d, float &tbest, float &b1, float &b2) {
    if (sc.n_tris == 0) return -1;
    const RayPre rp = ray_precompute(d);
#if PRT_TRI_ROWS
    const RayRows rr = ray_rows(rp);
#endif
    const Bvh8Ray r8 = bvh8_ray(o, d);
    uint2 gstack[BVH8_STACK];
    int sp = 0, best = -1;
    if (sc.n_small < sc.n_tris) {
#if PRT_TRI_ROWS
        best = test_big_tris<ANY>(sc, rr, o, tbest, b1, b2);
#else
        best = test_big_tris<ANY>(sc, rp, o, tbest, b1, b2);
#endif
        if ((ANY && best >= 0) || sc.n_small == 0) return best;
    }
    uint2 ng = make_uint2(0u, 0x80000000u);      // node group in hand: child base, hit bits | imask (root = slot 7 ^ octinv)
    for (;;) {
        uint2 tg = make_uint2(0u, 0u);
        if (ng.y > 0x00ffffffu) {
            const uint32_t hits = ng.y, imask8 = ng.y & 0xffu;
            const int bit = 31 - __clz(hits);
            ng.y &= ~(1u << bit);
            if (ng.y > 0x00ffffffu && sp < BVH8_STACK) gstack[sp++] = ng;
            const uint32_t slot_index = (uint32_t) (bit - 24) ^ (r8.octinv4 & 0xffu);
            const uint32_t rel = __popc(imask8 & ~(0xffffffffu << slot_index));
            uint32_t child_base, tri_base, imask;
            const uint32_t hm = bvh8_node(sc.nodes8, ng.x + rel, r8, tbest, child_base, tri_base, imask);
            ng = make_uint2(child_base, (hm & 0xff000000u) | imask);
            tg = make_uint2(tri_base, hm & 0x00ffffffu);
        }
        while (tg.y) {
            const int bit = 31 - __clz(tg.y);
            tg.y &= ~(1u << bit);
            const float4 *tv = sc.tri_v8 + 3 * (size_t) (tg.x + (uint32_t) bit);
            const float4 a = ldg4(tv), b = ldg4(tv + 1), c = ldg4(tv + 2);
#if PRT_TRI_ROWS
            if (intersect_tri_rows(rr, o, xyz(a), xyz(b), xyz(c), tbest, b1, b2)) {
#else
            if (intersect_tri_wt(rp, o, xyz(a), xyz(b), xyz(c), tbest, b1, b2)) {
#endif
                best = (int) (__float_as_uint(b.w) >> 2);      // k_bvh8_annotate: (sorted triangle << 2) | shading queue
                if (ANY) return best;
            }
        }
        if (ng.y <= 0x00ffffffu) {
            if (sp > 0) ng = gstack[--sp];
            else return best;
        }
    }
}

template <bool ANY>
__device__ __forceinline__ int traverse_tris(const DScene &sc, float3 o, float3 d, float &tbest, float &b1, float &b2) {
#if PRT_MEGA_BVH8
    return traverse_bvh8<ANY>(sc, o, d, tbest, b1, b2);
#else
    return traverse_bvh<ANY>(sc, o, d, tbest, b1, b2);
#endif
}

// scene.ray_intersect: nearest hit over analytic primitives (staged in shared memory by the caller)
// and the triangle BVH
// TRIS = false compiles the BVH traversal (and its stack) out: kernels specialised for analytic-only scenes
template <bool TRIS = true>
__device__ __forceinline__ bool closest_hit(const DScene &sc, const DPrim *prims, float3 o, float3 d, float tmax, Hit &h) {
    int best = -1;
    float tb = tmax;
    for (int i = 0; i < sc.n_prims; i++) {
        float t = intersect_prim(prims[i], o, d, tb);
        if (t >= 0.0f && (best < 0 || t < tb)) {
            best = i;
            tb = t;
        }
    }
    if (TRIS) {
        float b1 = 0.0f, b2 = 0.0f;
        float tt = tb;
        int tri = traverse_tris<false>(sc, o, d, tt, b1, b2);
        if (tri >= 0 && (best < 0 || tt < tb)) {
            fill_tri_hit(sc, tri, tt, b1, b2, h);
            return true;
        }
    }
    if (best < 0) return false;
    fill_prim_hit(prims[best], best, o, d, tb, h);
    return true;
}

template <bool TRIS = true>
__device__ __forceinline__ bool occluded(const DScene &sc, const DPrim *prims, float3 o, float3 d, float tmax) {
    // any hit: leave at the first one (in the Box scenes the connection ray of a whole warp is blocked by the same
    // primitive, so the exit is warp-uniform)
    for (int i = 0; i < sc.n_prims; i++)
        if (intersect_prim(prims[i], o, d, tmax) >= 0.0f) return true;
    if (!TRIS) return false;
    float b1, b2, tt = tmax;
    return traverse_tris<true>(sc, o, d, tt, b1, b2) >= 0;
}

}  // namespace prt
