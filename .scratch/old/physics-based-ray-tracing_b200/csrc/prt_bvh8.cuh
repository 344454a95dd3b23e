// prt_bvh8.cuh -- compressed 8-wide BVH ("BVH8c"): node layout and the per-lane traversal step.
//
// Why: on the 10 M-triangle scene (BASELINE config 5) the BVH2 closest-hit kernel is bound by the chain of dependent
// node fetches (profiles/r01: 23 GB of DRAM reads per bounce at 18 % of HBM bandwidth, stalls on the node LDG and on
// the local-memory stack).  An 8-wide node with child boxes quantised to 8 bits relative to the parent box (the
// layout of Ylitie, Karras, Laine, "Efficient incoherent ray traversal on GPUs through compressed wide BVHs", HPG 2017)
// is 80 B for eight children instead of 7 x 64 B of BVH2 nodes, cuts the dependent chain by ~3x, needs no distance
// sort (children sit in slots ordered by octant, so `slot ^ ray_octant` is the traversal order) and shrinks the
// traversal stack to one 8-byte entry per level.  It replaces the Embree BVH hidden inside mi.load_dict
// (/root/reference/USMain.py:257) for the light-transport kernels; triangle tests are the same watertight test as
// everywhere else, so hits do not change.
//
// Node = 5 x float4 (80 B):
//   n0 = (p.x, p.y, p.z, bits{ex, ey, ez, imask})     p = box origin, e* = biased exponents of the grid step,
//                                                     imask bit s = slot s holds an inner node
//   n1 = (bits child_base, bits tri_base, bits meta[0..3], bits meta[4..7])
//        meta: inner 0b001_11sss (sss = slot) ; leaf 0bccc_ooooo (ccc = unary triangle count 1/3/7, ooooo = offset of
//        its first triangle from tri_base, <= 23) ; empty 0
//   n2 = (qlo.x[0..3], qlo.x[4..7], qlo.y[0..3], qlo.y[4..7])
//   n3 = (qlo.z[0..3], qlo.z[4..7], qhi.x[0..3], qhi.x[4..7])
//   n4 = (qhi.y[0..3], qhi.y[4..7], qhi.z[0..3], qhi.z[4..7])
// Inner children of a node are stored contiguously from child_base in slot order; the triangles of its leaf children
// contiguously from tri_base (a second copy of the triangle array in this order, + the map back to LBVH order).
#pragma once
#include "prt_device.cuh"

namespace prt {

static constexpr int BVH8_STACK = 40;

struct Bvh8Ray {
    float3 o, inv;        // origin, 1 / direction (direction components clamped away from 0)
    uint32_t octinv4;     // (7 - octant) replicated in 4 bytes
};

__device__ __forceinline__ Bvh8Ray bvh8_ray(float3 o, float3 d) {
    const float eps = 8.2718061e-25f;  // 2^-80
    Bvh8Ray r;
    r.o = o;
    r.inv = mk3(1.0f / (fabsf(d.x) > eps ? d.x : copysignf(eps, d.x)), 1.0f / (fabsf(d.y) > eps ? d.y : copysignf(eps, d.y)),
                1.0f / (fabsf(d.z) > eps ? d.z : copysignf(eps, d.z)));
    const uint32_t oct = (d.x < 0.0f ? 4u : 0u) | (d.y < 0.0f ? 2u : 0u) | (d.z < 0.0f ? 1u : 0u);
    r.octinv4 = (7u - oct) * 0x01010101u;
    return r;
}

__device__ __forceinline__ uint32_t sign_extend_s8x4(uint32_t x) {
    uint32_t r;
    asm("prmt.b32 %0, %1, 0x0, 0x0000BA98;" : "=r"(r) : "r"(x));
    return r;
}

__device__ __forceinline__ float byte_f(uint32_t w, int j) { return (float) ((w >> (8 * j)) & 0xffu); }

// tests the (up to) eight children of node `idx`; returns the hit mask: bits 24..31 = inner children at position
// 24 + (slot ^ octinv), bits 0..23 = triangles (offsets from tri_base); writes child_base / tri_base / imask
// PLAIN: `nodes` is not global memory (a shared-memory copy of the top levels): ordinary loads instead of ld.global.nc
template <bool PLAIN = false>
__device__ __forceinline__ uint32_t bvh8_node(const float4 *__restrict__ nodes, uint32_t idx, const Bvh8Ray &r, float tmax,
                                              uint32_t &child_base, uint32_t &tri_base, uint32_t &imask) {
    const float4 *n = nodes + 5 * (size_t) idx;
    const float4 n0 = PLAIN ? n[0] : ldg4(n), n1 = PLAIN ? n[1] : ldg4(n + 1), n2 = PLAIN ? n[2] : ldg4(n + 2),
                 n3 = PLAIN ? n[3] : ldg4(n + 3), n4 = PLAIN ? n[4] : ldg4(n + 4);
    const uint32_t e = __float_as_uint(n0.w);
    imask = e >> 24;
    child_base = __float_as_uint(n1.x);
    tri_base = __float_as_uint(n1.y);
    const float ax = __uint_as_float((e & 0xffu) << 23) * r.inv.x;
    const float ay = __uint_as_float(((e >> 8) & 0xffu) << 23) * r.inv.y;
    const float az = __uint_as_float(((e >> 16) & 0xffu) << 23) * r.inv.z;
    const float ox = (n0.x - r.o.x) * r.inv.x, oy = (n0.y - r.o.y) * r.inv.y, oz = (n0.z - r.o.z) * r.inv.z;
    uint32_t hitmask = 0;
#pragma unroll
    for (int half = 0; half < 2; half++) {
        const uint32_t meta4 = __float_as_uint(half ? n1.w : n1.z);
        const uint32_t is_inner4 = (meta4 & (meta4 << 1)) & 0x10101010u;
        const uint32_t inner_mask4 = sign_extend_s8x4(is_inner4 << 3);
        const uint32_t bit_index4 = (meta4 ^ (r.octinv4 & inner_mask4)) & 0x1f1f1f1fu;
        const uint32_t child_bits4 = (meta4 >> 5) & 0x07070707u;
        const uint32_t lox = __float_as_uint(half ? n2.y : n2.x), loy = __float_as_uint(half ? n2.w : n2.z);
        const uint32_t loz = __float_as_uint(half ? n3.y : n3.x), hix = __float_as_uint(half ? n3.w : n3.z);
        const uint32_t hiy = __float_as_uint(half ? n4.y : n4.x), hiz = __float_as_uint(half ? n4.w : n4.z);
        const uint32_t nx = r.inv.x < 0.0f ? hix : lox, fx = r.inv.x < 0.0f ? lox : hix;
        const uint32_t ny = r.inv.y < 0.0f ? hiy : loy, fy = r.inv.y < 0.0f ? loy : hiy;
        const uint32_t nz = r.inv.z < 0.0f ? hiz : loz, fz = r.inv.z < 0.0f ? loz : hiz;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const float t0x = fmaf(byte_f(nx, j), ax, ox), t1x = fmaf(byte_f(fx, j), ax, ox);
            const float t0y = fmaf(byte_f(ny, j), ay, oy), t1y = fmaf(byte_f(fy, j), ay, oy);
            const float t0z = fmaf(byte_f(nz, j), az, oz), t1z = fmaf(byte_f(fz, j), az, oz);
            const float cmin = fmaxf(fmaxf(t0x, t0y), fmaxf(t0z, 0.0f));
            const float cmax = fminf(fminf(t1x, t1y), fminf(t1z, tmax));
            if (cmin <= cmax * 1.0000004f) hitmask |= ((child_bits4 >> (8 * j)) & 0xffu) << ((bit_index4 >> (8 * j)) & 0xffu);
        }
    }
    return hitmask;
}

}  // namespace prt
