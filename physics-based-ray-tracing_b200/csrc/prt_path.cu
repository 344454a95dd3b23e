// prt_path.cu -- light-transport path tracer (BASELINE config 4).
//
// Replaces mi.render(scene) for the scene /root/reference/scenes/cbox.xml describes: Mitsuba's `path` integrator
// (:5-9, max_depth, rr_depth 5), `perspective` sensor (:11-21), `independent` sampler (:22-24), `hdrfilm` + tent
// filter (:25-31), `diffuse` / `dielectric` / `conductor` BSDFs (:36-54) and the area emitter on the luminaire
// quad (:58-84).  None of that is reference code -- it is the un-vendored Mitsuba wheel -- so the semantics are
// restated from SURVEY.md Appendix C.7 and checked against oracle/orc_pt.inl.
//
// Execution model: one CTA owns a 16 x 16 pixel tile, one thread owns one pixel and walks its sample list with
// in-place path regeneration (a lane whose path ends splats it and starts its next sample in the same loop
// iteration).  The tent-filter splat (2 x 2 pixels) goes to an 18 x 18 RGBW tile in shared memory; only the
// finished tile is added to the global film, so the film sees 4 floats per pixel per launch instead of 16 atomics
// per SAMPLE.  Lanes of a warp form an 8 x 4 pixel block (coherent primary rays).  Per-path state lives in
// registers; the scene (analytic primitives in shared memory, BVH2 + triangles through the read-only path) is
// tiny for this config and stays on chip.
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "prt_hit.cuh"
#include "prt_internal.h"
#include "prt_path.h"

namespace prt {

static constexpr int PT_THREADS = 256;
struct PtCounters {
    unsigned paths, segments, rays, shadow;
};

// one iteration of path.cpp's loop with both ray queries executed in place; returns false when the path is finished
__device__ __forceinline__ bool pt_step(const PtDev &P, const DPrim *prims, PtState &st, PtCounters &cn) {
    Hit h;
    cn.rays++;
    const bool valid = closest_hit<true>(P.sc, prims, st.o, st.d, PRT_INF, h);
    if (valid) cn.segments++;
    ShadowReq sr;
    const bool live = pt_shade(P, st, h, valid, sr);
    if (sr.want) {
        cn.rays++;
        cn.shadow++;
        if (!occluded<true>(P.sc, prims, sr.o, sr.d, sr.tmax)) pt_apply_shadow(st, sr);
    }
    return live;
}

__global__ void __launch_bounds__(PT_THREADS, 2) k_render_path(const PtDev P) {
    __shared__ DPrim sprims[MAX_SMEM_PRIMS];
    __shared__ float4 tile[PT_HALO * PT_HALO];
    const DPrim *prims = sprims;
    if (P.sc.n_prims > MAX_SMEM_PRIMS) prims = P.sc.prims;
    else {
        const float4 *src = reinterpret_cast<const float4 *>(P.sc.prims);
        float4 *dst = reinterpret_cast<float4 *>(sprims);
        for (int i = threadIdx.x; i < P.sc.n_prims * (int) (sizeof(DPrim) / 16); i += blockDim.x) dst[i] = src[i];
    }
    // lane -> pixel of an 8 x 4 block; 2 x 4 warps tile the 16 x 16 CTA tile
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int lx = (warp & 1) * 8 + (lane & 7), ly = (warp >> 1) * 4 + (lane >> 3);
    PtCounters cn = { 0, 0, 0, 0 };
    const int n_tiles = P.tiles_x * P.tiles_y;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const int tx0 = (t % P.tiles_x) * PT_TILE, ty0 = (t / P.tiles_x) * PT_TILE;
        for (int i = threadIdx.x; i < PT_HALO * PT_HALO; i += blockDim.x) tile[i] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        __syncthreads();
        const int x = tx0 + lx, y = ty0 + ly;
        const bool inside = x < P.W && y < P.H;
        PtState st;
        uint32_t j = 0;
        bool live = false;
        for (;;) {
            if (!live && inside && j < P.n_s) {
                pt_init(P, x, y, P.s_offset + j * P.s_stride, st);
                j++;
                cn.paths++;
                live = true;
            }
            // warp-wide exit test = the reconvergence point of every iteration (a per-lane `break` let regenerating and
            // continuing lanes drift apart for good: see the same fix in k_acquire, prt_acquire.cu)
            if (!__any_sync(0xffffffffu, live)) break;
            if (live) {
                live = pt_step(P, prims, st, cn);
                if (!live) pt_splat(P.tent, tile, tx0, ty0, st.px, st.py, st.res);
            }
        }
        __syncthreads();
        pt_flush_tile(P, tile, tx0, ty0);
        __syncthreads();
    }
    if (P.stats) {
        unsigned v[4] = { cn.paths, cn.segments, cn.rays, cn.shadow };
#pragma unroll
        for (int q = 0; q < 4; q++) {
            unsigned xx = v[q];
#pragma unroll
            for (int o = 16; o; o >>= 1) xx += __shfl_xor_sync(0xffffffffu, xx, o);
            if (lane == 0 && xx) atomicAdd(P.stats + q, (unsigned long long) xx);
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Resident-scene kernel: scenes whose geometry fits shared memory (scenes/cbox.xml: 2 analytic spheres + 12 triangles).
//
// For such a scene the wavefront pipeline spends more than half of its step streaming path state and ray records through
// HBM (profiles/r01: shade + generate + film = 53 % of the cbox step, 380 B per segment) around ray queries that touch
// 2 KB of geometry.  Here nothing leaves the SM: the analytic primitives AND the triangles are staged in shared memory
// and tested by brute force (no tree, no stack, uniform control flow across the warp), path state lives in registers,
// and the only global traffic is the finished 16 x 16 film tile.  Unlike k_render_path, work inside a tile is handed out
// dynamically: the tile's (sample, pixel) items are numbered s * 256 + pixel-in-tile (32 consecutive items = one 8 x 4
// pixel block of one sample, so primary rays stay coherent) and a lane whose path ends takes the next item from a
// shared-memory counter, whichever pixel it belongs to -- a thread that drew long paths no longer holds its warp and its
// CTA back (with 16 samples per pixel the fixed pixel -> thread map left ~25 % of the lane-iterations idle).
// Same pt_init / pt_shade, same per-path PCG32 streams: identical paths, identical radiance per path.
// ------------------------------------------------------------------------------------------------------------------
static constexpr int PT_RES_MAX_TRIS = 64;
#ifndef PT_RES_MINB
#define PT_RES_MINB 3
#endif

struct ResScene {
    const DPrim *prims;      // shared memory
    const float4 *tv;        // shared memory: [n_tris][3] vertices, sorted order (v1.w unused here)
    const float4 *bx;        // shared memory: [n_boxes][2] distinct padded triangle boxes (lo, hi) ...
    const unsigned long long *bm;   // ... and the triangles each of them bounds (the two halves of a quad share one)
    int n_prims, n_tris, n_boxes;
};

#ifndef PT_RES_CULL
#define PT_RES_CULL 1               // 1: slab test against every triangle's box, watertight test only where the ray crosses it
#endif

#ifndef PT_RES_STASH
#define PT_RES_STASH 1              // 1: shadow rays parked in a per-warp stash and traced 32 at a time; 0: one query per lane and iteration, mixed kinds
#endif
#ifndef PT_RES_REGEN_MIN
#define PT_RES_REGEN_MIN 6        // idle lanes a warp collects before it splats finished paths and starts new ones
#endif

// nearest hit of the ray (o, d) within [0, tmax] by brute force over the staged scene: analytic primitive `prim` or sorted
// triangle `tri` (at most one of them >= 0 on return)
__device__ __forceinline__ bool res_query(const ResScene &R, float3 o, float3 d, float tmax, float &tb, int &prim, int &tri, float &b1,
                                          float &b2) {
    prim = -1;
    tri = -1;
    tb = tmax;
    for (int i = 0; i < R.n_prims; i++) {
        const float t = intersect_prim(R.prims[i], o, d, tb);
        if (t >= 0.0f && (prim < 0 || t < tb)) { prim = i; tb = t; }
    }
    const RayRows rr = ray_rows(ray_precompute(d));
#if PT_RES_CULL
    // The boxes are tested by all lanes together (one shared-memory broadcast per box, 14 instructions); the watertight
    // test (45) then runs only for a lane's own candidates, in ascending order like the plain loop -- the same triangles
    // give the same hits, ties included.  In a room every ray leaves through ONE wall: 2-3 candidates out of 12.
    const float3 inv = mk3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
    unsigned long long cand = 0ull;
    for (int u = 0; u < R.n_boxes; u++) {
        const float4 lo = R.bx[2 * u], hi = R.bx[2 * u + 1];
        if (box_entry(lo.x, lo.y, lo.z, hi.x, hi.y, hi.z, o, inv, tb) < PRT_INF) cand |= R.bm[u];
    }
    while (cand) {
        const int j = __ffsll((long long) cand) - 1;
        cand &= cand - 1ull;
        const float4 *tv = R.tv + 3 * j;
        if (intersect_tri_rows(rr, o, xyz(tv[0]), xyz(tv[1]), xyz(tv[2]), tb, b1, b2)) tri = j;
    }
#else
    for (int j = 0; j < R.n_tris; j++) {
        const float4 *tv = R.tv + 3 * j;
        if (intersect_tri_rows(rr, o, xyz(tv[0]), xyz(tv[1]), xyz(tv[2]), tb, b1, b2)) tri = j;
    }
#endif
    if (tri >= 0) prim = -1;
    return prim >= 0 || tri >= 0;
}

// One loop, one ray query per lane and iteration.  A lane is idle (mode 0: its path is finished, result waiting to be
// splatted), about to extend its path (mode 1) or about to trace its shadow ray (mode 2).  Both kinds of ray run through
// the SAME brute-force loop in the same iteration, so the query -- two thirds of all instructions -- executes with nearly
// full warps; a first version that traced "closest, then shadow" inside one iteration ran the shadow loop with 15 of 32
// lanes and regenerated / splatted paths 3 to 9 lanes at a time (ncu r02: 16.1 active lanes on average).  Finished lanes
// wait until PT_RES_REGEN_MIN of them can splat and restart together.
__global__ void __launch_bounds__(PT_THREADS, PT_RES_MINB) k_render_resident(const PtDev P) {
    __shared__ DPrim sprims[MAX_SMEM_PRIMS];
    __shared__ float4 stri[3 * PT_RES_MAX_TRIS];
    __shared__ float4 sbox[2 * PT_RES_MAX_TRIS];
    __shared__ unsigned long long sbmask[PT_RES_MAX_TRIS];
    __shared__ int s_nbox;
    __shared__ float4 tile[PT_HALO * PT_HALO];
    __shared__ unsigned s_next;
#if PT_RES_STASH
    __shared__ float s_stash[PT_THREADS / 32][64][12];     // < 32 parked before an E-iteration + at most 32 new ones
#endif
    {
        const float4 *src = reinterpret_cast<const float4 *>(P.sc.prims);
        float4 *dst = reinterpret_cast<float4 *>(sprims);
        for (int i = threadIdx.x; i < P.sc.n_prims * (int) (sizeof(DPrim) / 16); i += blockDim.x) dst[i] = src[i];
        for (int i = threadIdx.x; i < 3 * P.sc.n_tris; i += blockDim.x) stri[i] = P.sc.tri_v[i];
        if (threadIdx.x == 0) {      // distinct triangle boxes, padded by a few ulps (<= 64 triangles: a few microseconds, once per CTA)
            int nb = 0;
            for (int j = 0; j < P.sc.n_tris; j++) {
                const float4 a = P.sc.tri_v[3 * j], b = P.sc.tri_v[3 * j + 1], c = P.sc.tri_v[3 * j + 2];
                float lo[3] = { fminf(a.x, fminf(b.x, c.x)), fminf(a.y, fminf(b.y, c.y)), fminf(a.z, fminf(b.z, c.z)) };
                float hi[3] = { fmaxf(a.x, fmaxf(b.x, c.x)), fmaxf(a.y, fmaxf(b.y, c.y)), fmaxf(a.z, fmaxf(b.z, c.z)) };
                for (int k = 0; k < 3; k++) {
                    const float pad = 8.0f * 1.1920929e-7f * fmaxf(fmaxf(fabsf(lo[k]), fabsf(hi[k])), hi[k] - lo[k]) + 1e-30f;
                    lo[k] -= pad;
                    hi[k] += pad;
                }
                int u = 0;
                for (; u < nb; u++) {
                    const float4 l = sbox[2 * u], h = sbox[2 * u + 1];
                    if (l.x == lo[0] && l.y == lo[1] && l.z == lo[2] && h.x == hi[0] && h.y == hi[1] && h.z == hi[2]) break;
                }
                if (u == nb) {
                    sbox[2 * u] = make_float4(lo[0], lo[1], lo[2], 0.0f);
                    sbox[2 * u + 1] = make_float4(hi[0], hi[1], hi[2], 0.0f);
                    sbmask[u] = 0ull;
                    nb++;
                }
                sbmask[u] |= 1ull << j;
            }
            s_nbox = nb;
        }
    }
    __syncthreads();
    ResScene R;
    R.prims = sprims; R.tv = stri; R.bx = sbox; R.bm = sbmask; R.n_prims = P.sc.n_prims; R.n_tris = P.sc.n_tris; R.n_boxes = s_nbox;
    const int lane = threadIdx.x & 31;
    const unsigned FULLM = 0xffffffffu;
    PtCounters cn = { 0, 0, 0, 0 };
    const int n_tiles = P.tiles_x * P.tiles_y;
    const unsigned n_items = P.n_s * 256u;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const int tx0 = (t % P.tiles_x) * PT_TILE, ty0 = (t / P.tiles_x) * PT_TILE;
        for (int i = threadIdx.x; i < PT_HALO * PT_HALO; i += blockDim.x) tile[i] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        if (threadIdx.x == 0) s_next = 0u;
        __syncthreads();
#if PT_RES_STASH
        // Phase-separated warp iterations.  An E-iteration extends every live path by one segment: closest hit, pt_shade,
        // next ray -- all lanes in the same phase.  The shadow ray a segment asks for is PARKED in a per-warp shared-memory
        // stash together with what it will contribute (NEE radiance x MIS weight) and where (the sample's film position); an
        // S-iteration pops 32 of them, traces them as any-hit queries and splats the unoccluded ones straight into the tile
        // (radiance only: the sample's filter weight is added once, when its path ends).  Film accumulation is linear, so
        // splitting a sample's radiance into its terms changes nothing but the order of the float additions.
        PtState st;
        bool live = false, has_res = false, dry = false;      // dry: warp-uniform
        float(*stash)[12] = s_stash[threadIdx.x >> 5];
        int n_st = 0;                                         // warp-uniform
        for (;;) {
            const unsigned m_live = __ballot_sync(FULLM, live);
            if (n_st >= 32 || (n_st > 0 && m_live == 0u && dry)) {
                const int take = min(n_st, 32);
                if (lane < take) {
                    const float *e = stash[n_st - take + lane];
                    float tb, b1, b2;
                    int prim, tri;
                    cn.rays++;
                    cn.shadow++;
                    if (!res_query(R, mk3(e[0], e[1], e[2]), mk3(e[3], e[4], e[5]), e[6], tb, prim, tri, b1, b2))
                        pt_splat(P.tent, tile, tx0, ty0, e[10], e[11], mk3(e[7], e[8], e[9]), 0.0f);
                }
                n_st -= take;
                __syncwarp();
                continue;
            }
            const unsigned idle = ~m_live;
            if (idle && (__popc(idle) >= PT_RES_REGEN_MIN || m_live == 0u)) {
                if (!live && has_res) {
                    pt_splat(P.tent, tile, tx0, ty0, st.px, st.py, st.res);
                    has_res = false;
                }
                if (!dry) {
                    unsigned base = 0u;
                    const int leader = __ffs(idle) - 1;
                    if (lane == leader) base = atomicAdd(&s_next, (unsigned) __popc(idle));
                    base = __shfl_sync(FULLM, base, leader);
                    if (!live) {
                        const unsigned item = base + __popc(idle & ((1u << lane) - 1u));
                        if (item < n_items) {
                            const unsigned j = item >> 8, r = item & 255u, w = r >> 5, l = r & 31u;
                            const int x = tx0 + (int) ((w & 1u) * 8u + (l & 7u)), y = ty0 + (int) ((w >> 1) * 4u + (l >> 3));
                            if (x < P.W && y < P.H) {
                                pt_init(P, x, y, P.s_offset + j * P.s_stride, st);
                                cn.paths++;
                                live = true;
                            }
                        }
                    }
                    dry = base + (unsigned) __popc(idle) >= n_items;
                }
            }
            if (!__any_sync(FULLM, live)) {
                if (dry && n_st == 0) break;
                continue;
            }
            ShadowReq sr;
            sr.want = false;
            if (live) {
                float tb, b1 = 0.0f, b2 = 0.0f;
                int prim, tri;
                const bool hit = res_query(R, st.o, st.d, PRT_INF, tb, prim, tri, b1, b2);
                cn.rays++;
                Hit h;
                if (tri >= 0) fill_tri_hit(P.sc, tri, tb, b1, b2, h);
                else if (prim >= 0) fill_prim_hit(R.prims[prim], prim, st.o, st.d, tb, h);
                if (hit) cn.segments++;
                live = pt_shade(P, st, h, hit, sr);
                has_res = !live;
            }
            const unsigned m_sh = __ballot_sync(FULLM, sr.want);
            if (sr.want) {
                float *e = stash[n_st + __popc(m_sh & ((1u << lane) - 1u))];
                e[0] = sr.o.x; e[1] = sr.o.y; e[2] = sr.o.z; e[3] = sr.d.x; e[4] = sr.d.y; e[5] = sr.d.z; e[6] = sr.tmax;
                e[7] = sr.c.x * sr.w; e[8] = sr.c.y * sr.w; e[9] = sr.c.z * sr.w; e[10] = st.px; e[11] = st.py;
            }
            n_st += __popc(m_sh);
            __syncwarp();
        }
#else
        PtState st;
        ShadowReq sr;
        sr.want = false;
        int mode = 0;
        bool has_res = false, live_after = false, dry = false;      // dry: warp-uniform
        for (;;) {
            const unsigned idle = __ballot_sync(FULLM, mode == 0);
            if (idle && (__popc(idle) >= PT_RES_REGEN_MIN || (idle == FULLM))) {
                if (mode == 0 && has_res) {
                    pt_splat(P.tent, tile, tx0, ty0, st.px, st.py, st.res);
                    has_res = false;
                }
                if (!dry) {
                    unsigned base = 0u;
                    const int leader = __ffs(idle) - 1;
                    if (lane == leader) base = atomicAdd(&s_next, (unsigned) __popc(idle));
                    base = __shfl_sync(FULLM, base, leader);
                    if (mode == 0) {
                        const unsigned item = base + __popc(idle & ((1u << lane) - 1u));
                        if (item < n_items) {
                            const unsigned j = item >> 8, r = item & 255u, w = r >> 5, l = r & 31u;
                            const int x = tx0 + (int) ((w & 1u) * 8u + (l & 7u)), y = ty0 + (int) ((w >> 1) * 4u + (l >> 3));
                            if (x < P.W && y < P.H) {
                                pt_init(P, x, y, P.s_offset + j * P.s_stride, st);
                                cn.paths++;
                                mode = 1;
                            }
                        }
                    }
                    dry = base + (unsigned) __popc(idle) >= n_items;
                }
            }
            if (!__any_sync(FULLM, mode != 0)) {
                if (dry) break;
                continue;
            }
            if (mode != 0) {
                const bool ext = mode == 1;
                const float3 qo = ext ? st.o : sr.o, qd = ext ? st.d : sr.d;
                float tb, b1 = 0.0f, b2 = 0.0f;
                int prim, tri;
                const bool hit = res_query(R, qo, qd, ext ? PRT_INF : sr.tmax, tb, prim, tri, b1, b2);
                cn.rays++;
                if (ext) {
                    Hit h;
                    if (tri >= 0) fill_tri_hit(P.sc, tri, tb, b1, b2, h);
                    else if (prim >= 0) fill_prim_hit(R.prims[prim], prim, qo, qd, tb, h);
                    if (hit) cn.segments++;
                    live_after = pt_shade(P, st, h, hit, sr);
                    mode = sr.want ? 2 : (live_after ? 1 : 0);
                } else {
                    cn.shadow++;
                    if (!hit) pt_apply_shadow(st, sr);
                    mode = live_after ? 1 : 0;
                }
                has_res = mode == 0;
            }
        }
#endif
        __syncthreads();
        pt_flush_tile(P, tile, tx0, ty0);
        __syncthreads();
    }
    if (P.stats) {
        unsigned v[4] = { cn.paths, cn.segments, cn.rays, cn.shadow };
#pragma unroll
        for (int q = 0; q < 4; q++) {
            unsigned xx = v[q];
#pragma unroll
            for (int o = 16; o; o >>= 1) xx += __shfl_xor_sync(0xffffffffu, xx, o);
            if (lane == 0 && xx) atomicAdd(P.stats + q, (unsigned long long) xx);
        }
    }
}

// hdrfilm develop (scenes/cbox.xml:25-31): rgb = sum(w c) / sum(w); pixels no sample reached stay 0
__global__ void __launch_bounds__(256) k_film_develop(const float4 *__restrict__ film, uint64_t n_pixels, float *__restrict__ rgb) {
    for (uint64_t i = blockIdx.x * (uint64_t) blockDim.x + threadIdx.x; i < n_pixels; i += (uint64_t) gridDim.x * blockDim.x) {
        const float4 f = film[i];
        const float inv = f.w > 0.0f ? 1.0f / fmaxf(f.w, 1e-30f) : 0.0f;
        rgb[3 * i + 0] = f.x * inv;
        rgb[3 * i + 1] = f.y * inv;
        rgb[3 * i + 2] = f.z * inv;
    }
}

static int launch_develop(prt_context *c, const float *film, uint64_t n_pixels, float *rgb, cudaStream_t st) {
    uint64_t grid = (n_pixels + 255) / 256;
    const uint64_t cap = (uint64_t) c->sm_count * 8;
    if (grid > cap) grid = cap;
    if (grid < 1) grid = 1;
    k_film_develop<<<(unsigned) grid, 256, 0, st>>>(reinterpret_cast<const float4 *>(film), n_pixels, rgb);
    PRT_CUDA(cudaGetLastError());
    return PRT_OK;
}

static int fill_pt(prt_scene *s, const prt_render_params *p, uint64_t seed, uint32_t spp_total, uint32_t s_offset, uint32_t s_stride,
                   PtDev &P) {
    PRT_REQUIRE(p->width > 0 && p->height > 0 && p->max_depth >= 0 && spp_total > 0, "render_path: invalid render parameters");
    PRT_REQUIRE((uint64_t) p->width * p->height < (1ull << 31), "render_path: film too large");
    P.kind_mask = 0;
    for (auto &m : s->mats) P.kind_mask |= 1u << (m.kind & 31);
    for (auto &m : s->mats)
        PRT_REQUIRE(m.kind != PRT_MAT_ULTRA, "render_path: ultrasound_bsdf has no light-transport meaning (eval == 0, CustomBSDF.py:177)");
    for (auto &pr : s->prims) {
        const DMaterial &m = s->mats[pr.material];
        if (m.emission[0] > 0 || m.emission[1] > 0 || m.emission[2] > 0) {
            set_error("render_path: area emitters on analytic primitives are not supported (use a mesh)");
            return PRT_ERR_UNSUPPORTED;
        }
    }
    const double *m = p->to_world;
    P.sc = s->view();
    P.T0 = make_float4((float) m[0], (float) m[1], (float) m[2], (float) m[3]);
    P.T1 = make_float4((float) m[4], (float) m[5], (float) m[6], (float) m[7]);
    P.T2 = make_float4((float) m[8], (float) m[9], (float) m[10], (float) m[11]);
    double t = std::tan(p->fov_deg * M_PI / 360.0);
    if (p->width <= p->height) { P.tan_x = (float) t; P.tan_y = (float) (t * p->height / p->width); }
    else { P.tan_y = (float) t; P.tan_x = (float) (t * p->width / p->height); }
    P.near_clip = (float) p->near_clip;
    P.W = p->width; P.H = p->height; P.max_depth = p->max_depth; P.rr_depth = p->rr_depth; P.tent = p->rfilter == 1;
    P.seed = seed;
    P.spp_total = spp_total;
    P.s_offset = s_offset;
    P.s_stride = s_stride ? s_stride : 1;
    P.n_s = s_offset < spp_total ? (uint32_t) (((uint64_t) spp_total - s_offset + P.s_stride - 1) / P.s_stride) : 0;
    P.tiles_x = (p->width + PT_TILE - 1) / PT_TILE;
    P.tiles_y = (p->height + PT_TILE - 1) / PT_TILE;
    P.film = nullptr;
    P.stats = nullptr;
    P.box_lo = make_float3(s->stats.scene_lo[0], s->stats.scene_lo[1], s->stats.scene_lo[2]);
    P.box_hi = make_float3(s->stats.scene_hi[0], s->stats.scene_hi[1], s->stats.scene_hi[2]);
    return PRT_OK;
}

static int launch_pt(prt_context *c, const PtDev &P, cudaStream_t st) {
    // Two execution models.  The wavefront pipeline (prt_wavefront.cu: queues + dynamic ray
    // fetch) is the default: measured on B200 it beats this tile megakernel on both bench scenes (cbox 5.3 vs 3.6
    // Grays/s, 10 M-triangle height field 1.14 vs 0.67), because per-ray traversal length is heavy-tailed and the
    // megakernel idles the lanes of a warp behind its longest ray.  PRT_PT_MODE=mega|wavefront overrides.
    // PRT_PT_MODE=mega|wavefront|resident overrides.  Default: scenes whose geometry fits shared memory (<= 64 analytic
    // primitives, <= 64 triangles) run in the resident-scene kernel; everything else in the wavefront pipeline.
    const char *mode = getenv("PRT_PT_MODE");
    const bool fits = P.sc.n_prims <= MAX_SMEM_PRIMS && P.sc.n_tris <= PT_RES_MAX_TRIS;
    int which = fits ? 2 : 1;                      // 0 tile megakernel, 1 wavefront, 2 resident
    if (mode && mode[0] == 'm') which = 0;
    else if (mode && mode[0] == 'w') which = 1;
    else if (mode && mode[0] == 'r') which = 2;
    if (which == 2 && !fits) { set_error("render_path: PRT_PT_MODE=resident needs <= 64 primitives and <= 64 triangles"); return PRT_ERR_INVALID; }
    // The wavefront enqueues ~4 kernels per bounce for max_depth bounces without ever reading a queue length back; Mitsuba's
    // default max_depth = -1 (unbounded, mapped to 2^20 by the plugin) would be millions of empty launches.  Paths of
    // unbounded depth end by Russian roulette: the per-lane loop of the megakernel handles them naturally.
    if (which == 1 && P.max_depth > 64) which = 0;
    if (which == 1) return launch_wavefront(c, P, st);
    const void *kern = which == 2 ? (const void *) k_render_resident : (const void *) k_render_path;
    int per_sm = 0;
    PRT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, PT_THREADS, 0));
    if (per_sm < 1) per_sm = 1;
    int grid = c->sm_count * per_sm;
    int n_tiles = P.tiles_x * P.tiles_y;
    if (grid > n_tiles) grid = n_tiles;
    {
        ProfScope ps(c, PRT_KC_MEGAKERNEL, st);
        if (which == 2) k_render_resident<<<grid, PT_THREADS, 0, st>>>(P);
        else k_render_path<<<grid, PT_THREADS, 0, st>>>(P);
    }
    PRT_CUDA(cudaGetLastError());
    c->last_launches = 1;
    return PRT_OK;
}

}  // namespace prt

using namespace prt;

extern "C" {

int prt_render_path_dev(prt_scene *s, const prt_render_params *p, uint64_t seed, uint32_t spp_total, uint32_t sample_offset,
                        uint32_t sample_stride, float *film_rgbw_dev, uint64_t *stats_dev, void *stream) {
    PRT_REQUIRE(s && p && film_rgbw_dev, "prt_render_path_dev: null argument");
    if (!s->committed) { set_error("prt_render_path_dev: scene not committed"); return PRT_ERR_STATE; }
    std::lock_guard<std::mutex> lk(s->ctx->mtx);
    PRT_CUDA(cudaSetDevice(s->ctx->device));
    PtDev P;
    int rc = fill_pt(s, p, seed, spp_total, sample_offset, sample_stride, P);
    if (rc) return rc;
    P.film = film_rgbw_dev;
    P.stats = reinterpret_cast<unsigned long long *>(stats_dev);
    return launch_pt(s->ctx, P, (cudaStream_t) stream);
}

int prt_film_develop_dev(prt_context *c, const float *film_rgbw_dev, uint64_t n_pixels, float *rgb_dev, void *stream) {
    PRT_REQUIRE(c && film_rgbw_dev && rgb_dev, "prt_film_develop_dev: null argument");
    std::lock_guard<std::mutex> lk(c->mtx);
    PRT_CUDA(cudaSetDevice(c->device));
    return launch_develop(c, film_rgbw_dev, n_pixels, rgb_dev, (cudaStream_t) stream);
}

// host-buffer render: develop = false -> out is the RGBW film [H][W][4]; true -> the developed image [H][W][3]
static int render_host(prt_scene *s, const prt_render_params *p, uint64_t seed, uint32_t spp_total, uint32_t sample_offset,
                       uint32_t sample_stride, float *film_rgbw, prt_render_stats *stats, bool develop) {
    PRT_REQUIRE(s && p && film_rgbw, "prt_render_path: null argument");
    if (!s->committed) { set_error("prt_render_path: scene not committed"); return PRT_ERR_STATE; }
    std::lock_guard<std::mutex> lk(s->ctx->mtx);
    prt_context *c = s->ctx;
    PRT_CUDA(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    PtDev P;
    int rc = fill_pt(s, p, seed, spp_total, sample_offset, sample_stride, P);
    if (rc) return rc;
    const size_t n = (size_t) p->width * p->height * 4;
    const size_t n_out = develop ? n / 4 * 3 : n;
    rc = ensure_scratch(c, n, develop ? n_out : 0, 0);
    if (rc) return rc;
    ScopedEvents<4> ev;
    PRT_REQUIRE(ev.ok, "cudaEventCreate failed");
    cudaEvent_t e0 = ev.e[0], e1 = ev.e[1], e2 = ev.e[2], e3 = ev.e[3];
    PRT_CUDA(cudaEventRecord(e0, st));
    PRT_CUDA(cudaMemsetAsync(c->acc_dev, 0, sizeof(float) * n, st));
    PRT_CUDA(cudaMemsetAsync(c->stats_dev, 0, sizeof(uint64_t) * 16, st));
    P.film = c->acc_dev;
    P.stats = reinterpret_cast<unsigned long long *>(c->stats_dev);
    PRT_CUDA(cudaEventRecord(e1, st));
    rc = launch_pt(c, P, st);
    if (rc) return rc;
    const float *src = c->acc_dev;
    if (develop) {
        rc = launch_develop(c, c->acc_dev, n / 4, c->aux_dev, st);
        if (rc) return rc;
        c->last_launches++;
        src = c->aux_dev;
    }
    PRT_CUDA(cudaEventRecord(e2, st));
    float *pin = reinterpret_cast<float *>(c->pinned);
    cudaPointerAttributes attr;
    const bool pinned_dst = cudaPointerGetAttributes(&attr, film_rgbw) == cudaSuccess && attr.type == cudaMemoryTypeHost;
    cudaGetLastError();
    if (pinned_dst) pin = film_rgbw;
    PRT_CUDA(cudaMemcpyAsync(pin, src, sizeof(float) * n_out, cudaMemcpyDeviceToHost, st));
    uint64_t hs[16];
    PRT_CUDA(cudaMemcpyAsync(hs, c->stats_dev, sizeof(uint64_t) * 16, cudaMemcpyDeviceToHost, st));
    PRT_CUDA(cudaEventRecord(e3, st));
    PRT_CUDA(cudaStreamSynchronize(st));
    if (!pinned_dst) memcpy(film_rgbw, pin, sizeof(float) * n_out);
    if (getenv("PRT_WF_STATS"))     // counters of a -DWF_STATS=1 build (zero otherwise)
        fprintf(stderr, "wf_stats rays %llu shadow %llu | ext nodes %llu tris %llu iters %llu | sh nodes %llu tris %llu iters %llu\n",
                (unsigned long long) hs[2], (unsigned long long) hs[3], (unsigned long long) hs[4], (unsigned long long) hs[5],
                (unsigned long long) hs[8], (unsigned long long) hs[6], (unsigned long long) hs[7], (unsigned long long) hs[9]);
    if (stats) {
        stats->paths = hs[0]; stats->segments = hs[1]; stats->rays = hs[2]; stats->shadow_rays = hs[3];
        PRT_CUDA(cudaEventElapsedTime(&stats->kernel_ms, e1, e2));
        PRT_CUDA(cudaEventElapsedTime(&stats->total_ms, e0, e3));
        stats->launches = (uint32_t) c->last_launches;
        stats->_pad = 0;
    }
    return PRT_OK;
}

int prt_render_path(prt_scene *s, const prt_render_params *p, uint64_t seed, uint32_t spp_total, uint32_t sample_offset,
                    uint32_t sample_stride, float *film_rgbw, prt_render_stats *stats) {
    return render_host(s, p, seed, spp_total, sample_offset, sample_stride, film_rgbw, stats, false);
}

int prt_render_image(prt_scene *s, const prt_render_params *p, uint64_t seed, uint32_t spp_total, uint32_t sample_offset,
                     uint32_t sample_stride, float *image_rgb, prt_render_stats *stats) {
    return render_host(s, p, seed, spp_total, sample_offset, sample_stride, image_rgb, stats, true);
}

}  // extern "C"
