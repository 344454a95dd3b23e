// prt_bvh8.cu -- GPU build of the compressed 8-wide BVH (layout: prt_bvh8.cuh) from the LBVH's binary radix tree.
//
// Top-down, one kernel launch per level of the wide tree: every thread turns one binary sub-tree root into one wide
// node by repeatedly opening the child with the largest surface area until eight children are reached (sub-trees of
// <= 3 triangles stay closed: they become leaf children), places the children in octant-ordered slots (greedy
// assignment on dot(child centre - parent centre, slot diagonal)), quantises their boxes conservatively to 8 bits on
// a power-of-two grid anchored at the parent box, reserves contiguous storage for its inner children and for the
// triangles of its leaf children with two atomics, and writes the binary roots of its inner children as the tasks of
// the next level.  The host reads one counter per level (~log8 N levels).
#include <cfloat>
#include <cstring>
#include <string>

#include "prt_bvh8.cuh"
#include "prt_internal.h"

namespace prt {

#ifndef PRT_BVH8_FILL
#define PRT_BVH8_FILL 1
#endif
#ifndef PRT_LEAF8
#define PRT_LEAF8 3
#endif
static constexpr int LEAF8 = PRT_LEAF8;   // triangles per leaf child (unary count in 3 bits: 1 .. 3)

struct Child8 {
    int   ref;       // binary-tree ref: >= 0 inner node, < 0 single triangle ~sorted index
    float lo[3], hi[3];
};

__device__ __forceinline__ int sub_count(const int2 *__restrict__ ranges, int ref) {
    if (ref < 0) return 1;
    const int2 r = ranges[ref];
    return r.y - r.x + 1;
}

__global__ void __launch_bounds__(128) k_bvh8_level(int begin, int end, int *__restrict__ src, const float *__restrict__ nodes2,
                                                    const int2 *__restrict__ children, const int2 *__restrict__ ranges,
                                                    const float4 *__restrict__ tri_v_sorted, float4 *__restrict__ nodes8,
                                                    float4 *__restrict__ tri_v8, uint32_t *__restrict__ tri8_sorted,
                                                    int *__restrict__ counters) {
    const int idx = begin + blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= end) return;
    const int b2 = src[idx];
    Child8 c[8];
    int n = 2;
    {
        const float *rec = nodes2 + 16 * (size_t) b2;
        const int2 ch = children[b2];
        c[0].ref = ch.x;
        c[1].ref = ch.y;
#pragma unroll
        for (int k = 0; k < 3; k++) {
            c[0].lo[k] = rec[k]; c[0].hi[k] = rec[3 + k];
            c[1].lo[k] = rec[6 + k]; c[1].hi[k] = rec[9 + k];
        }
    }
    while (n < 8) {
        int best = -1;
        float best_a = -1.0f;
        for (int k = 0; k < n; k++) {
            if (c[k].ref < 0 || sub_count(ranges, c[k].ref) <= LEAF8) continue;
            const float ex = c[k].hi[0] - c[k].lo[0], ey = c[k].hi[1] - c[k].lo[1], ez = c[k].hi[2] - c[k].lo[2];
            const float a = ex * ey + ey * ez + ez * ex;
            if (a > best_a) { best_a = a; best = k; }
        }
        if (best < 0) break;
        const int o = c[best].ref;
        const float *rec = nodes2 + 16 * (size_t) o;
        const int2 ch = children[o];
        c[best].ref = ch.x;
        c[n].ref = ch.y;
#pragma unroll
        for (int k = 0; k < 3; k++) {
            c[best].lo[k] = rec[k]; c[best].hi[k] = rec[3 + k];
            c[n].lo[k] = rec[6 + k]; c[n].hi[k] = rec[9 + k];
        }
        n++;
    }
#if PRT_BVH8_FILL
    // Spare slots go to the leaf children: a bottom node ends up with ~4 children (1 inner + 3 leaves of 2-3 triangles on the
    // 10 M-triangle scene), yet the traversal tests all eight slots anyway.  Splitting the largest multi-triangle leaves until
    // the slots are used gives their triangles tighter boxes of their own -- fewer triangle tests per ray, same node count.
    while (n < 8) {
        int best = -1;
        float best_a = -1.0f;
        for (int k = 0; k < n; k++) {
            if (c[k].ref < 0 || sub_count(ranges, c[k].ref) > LEAF8) continue;
            const float ex = c[k].hi[0] - c[k].lo[0], ey = c[k].hi[1] - c[k].lo[1], ez = c[k].hi[2] - c[k].lo[2];
            const float a = ex * ey + ey * ez + ez * ex;
            if (a > best_a) { best_a = a; best = k; }
        }
        if (best < 0) break;
        const int o = c[best].ref;
        const float *rec = nodes2 + 16 * (size_t) o;
        const int2 ch = children[o];
        c[best].ref = ch.x;
        c[n].ref = ch.y;
#pragma unroll
        for (int k = 0; k < 3; k++) {
            c[best].lo[k] = rec[k]; c[best].hi[k] = rec[3 + k];
            c[n].lo[k] = rec[6 + k]; c[n].hi[k] = rec[9 + k];
        }
        n++;
    }
#endif
    // parent box, grid
    float plo[3] = { FLT_MAX, FLT_MAX, FLT_MAX }, phi[3] = { -FLT_MAX, -FLT_MAX, -FLT_MAX };
    for (int k = 0; k < n; k++)
#pragma unroll
        for (int a = 0; a < 3; a++) {
            plo[a] = fminf(plo[a], c[k].lo[a]);
            phi[a] = fmaxf(phi[a], c[k].hi[a]);
        }
    uint32_t eb[3];
    float inv_step[3];
#pragma unroll
    for (int a = 0; a < 3; a++) {
        const float ext = fmaxf(phi[a] - plo[a], 1e-30f);
        int e;
        frexpf(ext / 255.0f, &e);          // ext / 255 = m * 2^e, m in [0.5, 1)  ->  2^e >= ext / 255
        e = max(min(e, 126), -125);
        while (e < 126 && ceilf((phi[a] - plo[a]) * exp2f((float) -e) + 0.002f) > 255.0f) e++;
        eb[a] = (uint32_t) (e + 127);
        inv_step[a] = exp2f((float) -e);
    }
    // slots: greedy assignment, best (child, slot) pair first
    int slot_of[8], child_in[8];
#pragma unroll
    for (int k = 0; k < 8; k++) { slot_of[k] = -1; child_in[k] = -1; }
    const float pc[3] = { 0.5f * (plo[0] + phi[0]), 0.5f * (plo[1] + phi[1]), 0.5f * (plo[2] + phi[2]) };
    for (int it = 0; it < n; it++) {
        float bestv = -FLT_MAX;
        int bk = -1, bs = -1;
        for (int k = 0; k < n; k++) {
            if (slot_of[k] >= 0) continue;
            const float dx = 0.5f * (c[k].lo[0] + c[k].hi[0]) - pc[0], dy = 0.5f * (c[k].lo[1] + c[k].hi[1]) - pc[1],
                        dz = 0.5f * (c[k].lo[2] + c[k].hi[2]) - pc[2];
            for (int s = 0; s < 8; s++) {
                if (child_in[s] >= 0) continue;
                const float v = ((s & 4) ? dx : -dx) + ((s & 2) ? dy : -dy) + ((s & 1) ? dz : -dz);
                if (v > bestv) { bestv = v; bk = k; bs = s; }
            }
        }
        slot_of[bk] = bs;
        child_in[bs] = bk;
    }
    // classify, reserve storage
    uint32_t imask = 0;
    int n_inner = 0, n_leaf_tris = 0;
    for (int s = 0; s < 8; s++) {
        const int k = child_in[s];
        if (k < 0) continue;
        const int cnt = sub_count(ranges, c[k].ref);
        if (c[k].ref >= 0 && cnt > LEAF8) { imask |= 1u << s; n_inner++; }
        else n_leaf_tris += cnt;
    }
    const int child_base = n_inner ? atomicAdd(&counters[0], n_inner) : 0;
    const int tri_base = n_leaf_tris ? atomicAdd(&counters[1], n_leaf_tris) : 0;
    uint32_t meta[8], q[6][8];
    int rel = 0, off = 0;
    for (int s = 0; s < 8; s++) {
        const int k = child_in[s];
        if (k < 0) {
            meta[s] = 0;
#pragma unroll
            for (int a = 0; a < 3; a++) { q[a][s] = 255u; q[3 + a][s] = 0u; }
            continue;
        }
#pragma unroll
        for (int a = 0; a < 3; a++) {
            const float l = floorf((c[k].lo[a] - plo[a]) * inv_step[a] - 0.001f), h = ceilf((c[k].hi[a] - plo[a]) * inv_step[a] + 0.001f);
            q[a][s] = (uint32_t) fminf(fmaxf(l, 0.0f), 255.0f);
            q[3 + a][s] = (uint32_t) fminf(fmaxf(h, 0.0f), 255.0f);
        }
        if (imask & (1u << s)) {
            meta[s] = (1u << 5) | (24u + (uint32_t) s);
            src[child_base + rel] = c[k].ref;
            rel++;
        } else {
            const int cnt = sub_count(ranges, c[k].ref);
            const int first = c[k].ref < 0 ? ~c[k].ref : ranges[c[k].ref].x;
            meta[s] = (((1u << cnt) - 1u) << 5) | (uint32_t) off;
            for (int j = 0; nodes8 && j < cnt; j++) {
                const size_t dst = (size_t) tri_base + off + j, sidx = (size_t) first + j;
                tri_v8[3 * dst] = tri_v_sorted[3 * sidx];
                tri_v8[3 * dst + 1] = tri_v_sorted[3 * sidx + 1];
                tri_v8[3 * dst + 2] = tri_v_sorted[3 * sidx + 2];
                tri8_sorted[dst] = (uint32_t) sidx;
            }
            off += cnt;
        }
    }
    if (!nodes8) return;     // counting pass: only the task list and the two counters are produced
    auto pack4 = [](const uint32_t *b) { return b[0] | (b[1] << 8) | (b[2] << 16) | (b[3] << 24); };
    float4 *out = nodes8 + BVH8_NODE_F4 * (size_t) idx;
    out[0] = make_float4(plo[0], plo[1], plo[2], __uint_as_float(eb[0] | (eb[1] << 8) | (eb[2] << 16) | (imask << 24)));
    out[1] = make_float4(__int_as_float(child_base), __int_as_float(tri_base), __uint_as_float(pack4(meta)), __uint_as_float(pack4(meta + 4)));
    out[2] = make_float4(__uint_as_float(pack4(q[0])), __uint_as_float(pack4(q[0] + 4)), __uint_as_float(pack4(q[1])), __uint_as_float(pack4(q[1] + 4)));
    out[3] = make_float4(__uint_as_float(pack4(q[2])), __uint_as_float(pack4(q[2] + 4)), __uint_as_float(pack4(q[3])), __uint_as_float(pack4(q[3] + 4)));
    out[4] = make_float4(__uint_as_float(pack4(q[4])), __uint_as_float(pack4(q[4] + 4)), __uint_as_float(pack4(q[5])), __uint_as_float(pack4(q[5] + 4)));
}

// After the per-triangle tables exist (sorted order): stamp every BVH8 triangle with what a hit needs, so that retiring a
// ray costs no dependent loads: v1.w = bits((sorted index << 2) | shading queue of its material)
__global__ void k_bvh8_annotate(uint32_t n, const uint32_t *__restrict__ tri8_sorted, const int4 *__restrict__ tri_info,
                                const DMaterial *__restrict__ mats, float4 *__restrict__ tri_v8) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t sorted = tri8_sorted[i];
    const int kind = mats[tri_info[sorted].z].kind;
    const uint32_t qi = kind == PRT_MAT_DIFFUSE ? 0u : (kind == PRT_MAT_DIELECTRIC ? 1u : 2u);
    tri_v8[3 * (size_t) i + 1].w = __uint_as_float((sorted << 2) | qi);
}

int bvh8_annotate(uint32_t n, const uint32_t *tri8_sorted, const int4 *tri_info, const DMaterial *mats, float4 *tri_v8, cudaStream_t st) {
    if (!n || !tri_v8) return PRT_OK;
    k_bvh8_annotate<<<(n + 255) / 256, 256, 0, st>>>(n, tri8_sorted, tri_info, mats, tri_v8);
    PRT_CUDA(cudaGetLastError());
    return PRT_OK;
}

// n >= 1 triangles.  nodes2 (16 floats / binary node, padded child boxes), children, ranges: the LBVH's arrays.
// On success *out_nodes8 / *out_tri_v8 / *out_tri8_sorted are fresh device allocations owned by the caller.
int build_bvh8(uint32_t n, const float4 *tri_v_sorted, const float *nodes2, const int2 *children, const int2 *ranges,
               float4 **out_nodes8, uint32_t *out_n_nodes8, float4 **out_tri_v8, uint32_t **out_tri8_sorted, int *out_levels,
               cudaStream_t st, uint32_t n_extra) {
    *out_nodes8 = nullptr;
    *out_tri_v8 = nullptr;
    *out_tri8_sorted = nullptr;
    *out_n_nodes8 = 0;
    *out_levels = 0;
    if (n == 0) return PRT_OK;
    PRT_REQUIRE(n_extra == 0 || n >= 2, "build_bvh8: a super root needs a hierarchy below it");
    if (n == 1) {
        // no binary tree exists: one wide node whose slot 0 is a one-triangle leaf covering the node's whole grid
        float4 v[3];
        PRT_CUDA(cudaMemcpyAsync(v, tri_v_sorted, sizeof v, cudaMemcpyDeviceToHost, st));
        PRT_CUDA(cudaStreamSynchronize(st));
        float lo[3] = { fminf(v[0].x, fminf(v[1].x, v[2].x)), fminf(v[0].y, fminf(v[1].y, v[2].y)), fminf(v[0].z, fminf(v[1].z, v[2].z)) };
        float hi[3] = { fmaxf(v[0].x, fmaxf(v[1].x, v[2].x)), fmaxf(v[0].y, fmaxf(v[1].y, v[2].y)), fmaxf(v[0].z, fmaxf(v[1].z, v[2].z)) };
        uint32_t eb[3];
        for (int a = 0; a < 3; a++) {
            const float pad = 4.0f * 1.1920929e-7f * fmaxf(fmaxf(fabsf(lo[a]), fabsf(hi[a])), hi[a] - lo[a]) + 1e-30f;
            lo[a] -= pad;
            hi[a] += pad;
            int e;
            frexpf(fmaxf(hi[a] - lo[a], 1e-30f) / 254.0f, &e);
            e = e < -125 ? -125 : (e > 126 ? 126 : e);
            eb[a] = (uint32_t) (e + 127);
        }
        auto bits = [](uint32_t u) { float f; memcpy(&f, &u, 4); return f; };
        float4 node[5];
        node[0] = make_float4(lo[0], lo[1], lo[2], bits(eb[0] | (eb[1] << 8) | (eb[2] << 16)));
        node[1] = make_float4(bits(0u), bits(0u), bits(1u << 5), bits(0u));
        const uint32_t qlo = 0xffffff00u, qhi = 0x000000ffu;     // slot 0: [0, 255]; empty slots: lo 255 > hi 0
        node[2] = make_float4(bits(qlo), bits(0xffffffffu), bits(qlo), bits(0xffffffffu));
        node[3] = make_float4(bits(qlo), bits(0xffffffffu), bits(qhi), bits(0u));
        node[4] = make_float4(bits(qhi), bits(0u), bits(qhi), bits(0u));
        float4 *nodes8 = nullptr, *tv8 = nullptr;
        uint32_t *map = nullptr;
        PRT_CUDA(cudaMalloc(&nodes8, sizeof(float4) * BVH8_NODE_F4));
        PRT_CUDA(cudaMalloc(&tv8, sizeof v));
        PRT_CUDA(cudaMalloc(&map, sizeof(uint32_t)));
        PRT_CUDA(cudaMemcpy(nodes8, node, sizeof node, cudaMemcpyHostToDevice));
        PRT_CUDA(cudaMemcpy(tv8, v, sizeof v, cudaMemcpyHostToDevice));
        PRT_CUDA(cudaMemset(map, 0, sizeof(uint32_t)));
        *out_nodes8 = nodes8;
        *out_n_nodes8 = 1;
        *out_tri_v8 = tv8;
        *out_tri8_sorted = map;
        *out_levels = 1;
        return PRT_OK;
    }
    // Two passes over the same level-by-level loop: the first only counts (the wide tree has ~n/7 nodes, but the safe
    // a-priori bound is n: an 800 MB scratch array at 10 M triangles whose cudaFree alone cost 0.7 s), the second writes
    // into exactly sized arrays (the node count does not depend on the order in which the atomics hand out storage).
    const size_t cap = (size_t) n;                 // tasks: every wide node opens a distinct binary node with > 3 triangles
    float4 *nodes8 = nullptr, *tri_v8 = nullptr;
    uint32_t *tri8_sorted = nullptr;
    int *src = nullptr, *counters = nullptr;
    PRT_CUDA(cudaMalloc(&src, sizeof(int) * cap));
    PRT_CUDA(cudaMalloc(&counters, sizeof(int) * 2));
    int end = 1, levels = 0;
    for (int pass = 0; pass < 2; pass++) {
        if (pass == 1) {
            PRT_CUDA(cudaMalloc(&nodes8, sizeof(float4) * BVH8_NODE_F4 * (size_t) end));
            PRT_CUDA(cudaMalloc(&tri_v8, sizeof(float4) * 3 * ((size_t) n + n_extra)));
            PRT_CUDA(cudaMalloc(&tri8_sorted, sizeof(uint32_t) * ((size_t) n + n_extra)));
        }
        // n_extra > 0: node 0 and the first n_extra triangle records are left for the caller's super root
        // (bvh8_write_super_root); the hierarchy's own root becomes node 1
        const int first = n_extra ? 1 : 0;
        const int init[2] = { first + 1, (int) n_extra };
        PRT_CUDA(cudaMemcpyAsync(counters, init, sizeof init, cudaMemcpyHostToDevice, st));
        PRT_CUDA(cudaMemsetAsync(src, 0, sizeof(int) * 2, st));   // task `first` = binary root (binary node 0)
        int begin = first;
        end = first + 1;
        levels = first;
        while (begin < end) {
            const int cnt = end - begin;
            k_bvh8_level<<<(cnt + 127) / 128, 128, 0, st>>>(begin, end, src, nodes2, children, ranges, tri_v_sorted, nodes8, tri_v8,
                                                            tri8_sorted, counters);
            int h[2];
            PRT_CUDA(cudaMemcpyAsync(h, counters, sizeof h, cudaMemcpyDeviceToHost, st));
            PRT_CUDA(cudaStreamSynchronize(st));
            begin = end;
            end = h[0];
            levels++;
            if (levels > 256) { set_error("build_bvh8: runaway depth"); return PRT_ERR_STATE; }
        }
        PRT_CUDA(cudaGetLastError());
    }
    cudaFree(src);
    cudaFree(counters);
    if (levels > BVH8_STACK) {      // one stack entry per level at most: deeper trees would drop subtrees silently
        cudaFree(nodes8);
        cudaFree(tri_v8);
        cudaFree(tri8_sorted);
        set_error("build_bvh8: the 8-wide tree is " + std::to_string(levels) + " levels deep; the traversal stack holds " +
                  std::to_string(BVH8_STACK));
        return PRT_ERR_UNSUPPORTED;
    }
    *out_nodes8 = nodes8;
    *out_n_nodes8 = (uint32_t) end;
    *out_tri_v8 = tri_v8;
    *out_tri8_sorted = tri8_sorted;
    *out_levels = levels;
    return PRT_OK;
}

// Super root: triangles whose boxes span much of the scene (the walls of a room around a 10 M-triangle floor) inflate every
// ancestor on their Morton path when they sit in the LBVH -- rays then visit leaves all over the tree (measured on BASELINE
// config 5: 11.3 triangle tests and 13.0 node steps per ray with them in the tree; 4.5 and 9.8 for the floor alone).  They
// are kept out of the LBVH and become LEAF CHILDREN of an extra node 0, whose one inner child is the hierarchy's root
// (node 1): one more node step per ray, and only the oversized triangles whose own box the ray crosses are tested.
//   big_v   host, [n_extra][3], consecutive triangles form the leaf children (ceil(n_extra / 7) each); v1.w already holds
//           the annotation ((sorted index << 2) | shading queue)
//   root_lo / root_hi   box of the hierarchy (node 1)
int bvh8_write_super_root(uint32_t n_small, uint32_t n_extra, const float4 *big_v, const float root_lo[3], const float root_hi[3],
                          float4 *nodes8, float4 *tri_v8, uint32_t *tri8_sorted, cudaStream_t st) {
    PRT_REQUIRE(n_extra >= 1 && n_extra <= 21, "bvh8_write_super_root: 1..21 triangles");
    const int per = (int) (n_extra + 6) / 7, n_leaf = ((int) n_extra + per - 1) / per;
    float lo[8][3], hi[8][3];
    int cnt[8];
    auto pad_box = [](float *l, float *h) {
        for (int a = 0; a < 3; a++) {
            const float pad = 4.0f * 1.1920929e-7f * fmaxf(fmaxf(fabsf(l[a]), fabsf(h[a])), h[a] - l[a]) + 1e-30f;
            l[a] -= pad;
            h[a] += pad;
        }
    };
    for (int a = 0; a < 3; a++) { lo[0][a] = root_lo[a]; hi[0][a] = root_hi[a]; }
    pad_box(lo[0], hi[0]);
    cnt[0] = 0;
    for (int g = 0; g < n_leaf; g++) {
        float *l = lo[1 + g], *h = hi[1 + g];
        for (int a = 0; a < 3; a++) { l[a] = FLT_MAX; h[a] = -FLT_MAX; }
        cnt[1 + g] = 0;
        for (int j = g * per; j < (int) n_extra && j < (g + 1) * per; j++, cnt[1 + g]++)
            for (int c = 0; c < 3; c++) {
                const float4 v = big_v[3 * j + c];
                l[0] = fminf(l[0], v.x); l[1] = fminf(l[1], v.y); l[2] = fminf(l[2], v.z);
                h[0] = fmaxf(h[0], v.x); h[1] = fmaxf(h[1], v.y); h[2] = fmaxf(h[2], v.z);
            }
        pad_box(l, h);
    }
    const int n = 1 + n_leaf;
    float plo[3] = { FLT_MAX, FLT_MAX, FLT_MAX }, phi[3] = { -FLT_MAX, -FLT_MAX, -FLT_MAX };
    for (int k = 0; k < n; k++)
        for (int a = 0; a < 3; a++) { plo[a] = fminf(plo[a], lo[k][a]); phi[a] = fmaxf(phi[a], hi[k][a]); }
    uint32_t eb[3];
    float inv_step[3];
    for (int a = 0; a < 3; a++) {       // same grid rule as k_bvh8_level
        const float ext = fmaxf(phi[a] - plo[a], 1e-30f);
        int e;
        frexpf(ext / 255.0f, &e);
        e = e > 126 ? 126 : (e < -125 ? -125 : e);
        while (e < 126 && ceilf((phi[a] - plo[a]) * exp2f((float) -e) + 0.002f) > 255.0f) e++;
        eb[a] = (uint32_t) (e + 127);
        inv_step[a] = exp2f((float) -e);
    }
    uint32_t meta[8] = { 0 }, q[6][8];
    for (int s = 0; s < 8; s++)
        for (int a = 0; a < 3; a++) { q[a][s] = 255u; q[3 + a][s] = 0u; }
    int off = 0;
    for (int k = 0; k < n; k++) {       // child k sits in slot k: the order among ONE inner child and leaves does not matter
        for (int a = 0; a < 3; a++) {
            const float l = floorf((lo[k][a] - plo[a]) * inv_step[a] - 0.001f), h = ceilf((hi[k][a] - plo[a]) * inv_step[a] + 0.001f);
            q[a][k] = (uint32_t) fminf(fmaxf(l, 0.0f), 255.0f);
            q[3 + a][k] = (uint32_t) fminf(fmaxf(h, 0.0f), 255.0f);
        }
        if (k == 0) meta[k] = (1u << 5) | 24u;
        else {
            meta[k] = (((1u << cnt[k]) - 1u) << 5) | (uint32_t) off;
            off += cnt[k];
        }
    }
    auto bits = [](uint32_t u) { float f; memcpy(&f, &u, 4); return f; };
    auto pack4 = [](const uint32_t *b) { return b[0] | (b[1] << 8) | (b[2] << 16) | (b[3] << 24); };
    float4 node[5];
    node[0] = make_float4(plo[0], plo[1], plo[2], bits(eb[0] | (eb[1] << 8) | (eb[2] << 16) | (1u << 24)));   // imask: slot 0
    node[1] = make_float4(bits(1u), bits(0u), bits(pack4(meta)), bits(pack4(meta + 4)));                        // child_base 1, tri_base 0
    node[2] = make_float4(bits(pack4(q[0])), bits(pack4(q[0] + 4)), bits(pack4(q[1])), bits(pack4(q[1] + 4)));
    node[3] = make_float4(bits(pack4(q[2])), bits(pack4(q[2] + 4)), bits(pack4(q[3])), bits(pack4(q[3] + 4)));
    node[4] = make_float4(bits(pack4(q[4])), bits(pack4(q[4] + 4)), bits(pack4(q[5])), bits(pack4(q[5] + 4)));
    uint32_t map[21];
    for (uint32_t j = 0; j < n_extra; j++) map[j] = n_small + j;
    PRT_CUDA(cudaMemcpyAsync(nodes8, node, sizeof node, cudaMemcpyHostToDevice, st));
    PRT_CUDA(cudaMemcpyAsync(tri_v8, big_v, sizeof(float4) * 3 * n_extra, cudaMemcpyHostToDevice, st));
    PRT_CUDA(cudaMemcpyAsync(tri8_sorted, map, sizeof(uint32_t) * n_extra, cudaMemcpyHostToDevice, st));
    PRT_CUDA(cudaStreamSynchronize(st));     // the sources are host stack / caller memory
    return PRT_OK;
}

}  // namespace prt
