// prt_wavefront.cu -- wavefront form of the light-transport path tracer (BASELINE
// configs 4 and 5; north star item 3: "ray-gen, BVH traversal, per-material shading queues compacted with warp
// ballot/prefix-sum, shadow-ray and emitter-NEE kernels, film accumulation").
//
// It computes exactly what prt_path.cu's tile megakernel computes -- both run prt_path.h's pt_init / pt_shade on
// the same per-path PCG32 streams (the Mitsuba `path` integrator of /root/reference/scenes/cbox.xml:5-9, restated
// from SURVEY.md Appendix C.7) -- but moves the two ray queries into their own persistent kernels:
//
//   k_wf_generate            camera rays of a batch of samples -> path state (SoA, float4 per field) + extend queue
//   per bounce b:
//     k_wf_trace<false>      closest hit.  Persistent warps pull rays from the queue in chunks; a lane whose ray is
//                            finished retires it and takes the next one, so the heavy-tailed traversal length of an
//                            incoherent ray (profiles/r01_summary.md: 176 node steps in a warp whose rays need 28 on
//                            average) no longer idles the other 31 lanes.  Retiring classifies the hit by material
//                            and appends the path to that material's shading queue (warp ballot + one atomic).
//     k_wf_shade<queue>      one thread per queued path: rebuilds the surface interaction from (t, b1, b2, id), runs
//                            pt_shade, appends the shadow-ray request to the shadow queue and the surviving path to
//                            the next bounce's extend queue (both ballot-compacted)
//     k_wf_trace<true>       shadow rays (any hit), same dynamic fetch; unoccluded ones add their NEE term
//   k_wf_film                one CTA per 16 x 16 pixel tile splats the tile's finished samples into shared memory
//                            (tent filter) and adds the tile to the film
//
// All launches of a batch are enqueued back to back: queue lengths live in device memory (one counter block per
// bounce, zeroed once), so the host never synchronises inside a batch.
#include <cstdlib>

#include "prt_bvh8.cuh"
#include "prt_internal.h"
#include "prt_path.h"

namespace prt {

static constexpr unsigned FULL = 0xffffffffu;
static constexpr int WF_QUEUES = 3;        // shading queues: 0 diffuse, 1 dielectric, 2 everything else (conductor, null)
static constexpr int WF_CSTRIDE = 16;      // ints per bounce in the counter array
static constexpr int WF_CHUNK = 64;        // rays a warp reserves per atomic on the queue head
static constexpr int WF_TRACE_THREADS = 128;
static constexpr int WF_SHADE_THREADS = 256;
enum { C_EXT = 0, C_MAT = 1, C_SH = 4, C_HEAD_EXT = 8, C_HEAD_SH = 12 };

struct WfBuf {
    float4 *S0;      // o.xyz, px
    float4 *S1;      // d.xyz, py
    float4 *S2;      // throughput rgb, eta
    float4 *S3;      // radiance rgb, prev_pdf
    float4 *S4;      // prev_p.xyz, bits(depth | prev_delta << 16)
    uint4  *RNG;     // pcg32 state, inc
    float4 *HIT;     // t, b1, b2, bits(id): id < 0 miss, < n_prims analytic primitive, else n_prims + sorted triangle
    float4 *SH0;     // shadow ray o.xyz, tmax        (indexed by shadow-queue position)
    float4 *SH1;     // shadow ray d.xyz, mis weight
    float4 *SH2;     // contribution rgb, bits(slot)
    uint32_t *q_ext[2];
    uint32_t *q_mat[WF_QUEUES];
    int *cnt;        // [bounces + 1][WF_CSTRIDE]
    uint32_t cap, L, n_layers, j0;
};

__device__ __forceinline__ unsigned lanemask_lt() {
    unsigned m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

// warp-aggregated append; must be reached by all 32 lanes.  Returns the position this lane's entry got (or -1).
__device__ __forceinline__ int wf_reserve(int *counter, bool pred) {
    const unsigned m = __ballot_sync(FULL, pred);
    if (!m) return -1;
    const int leader = __ffs(m) - 1;
    int base = 0;
    if ((int) (threadIdx.x & 31) == leader) base = atomicAdd(counter, __popc(m));
    base = __shfl_sync(FULL, base, leader);
    return pred ? base + __popc(m & lanemask_lt()) : -1;
}

__device__ __forceinline__ void wf_load_state(const WfBuf &B, uint32_t slot, PtState &st) {
    float4 a = B.S0[slot], b = B.S1[slot], c = B.S2[slot], d = B.S3[slot], e = B.S4[slot];
    uint4 r = B.RNG[slot];
    st.o = xyz(a); st.px = a.w;
    st.d = xyz(b); st.py = b.w;
    st.thr = xyz(c); st.eta = c.w;
    st.res = xyz(d); st.prev_pdf = d.w;
    st.prev_p = xyz(e);
    int f = __float_as_int(e.w);
    st.depth = f & 0xffff;
    st.prev_delta = (f >> 16) != 0;
    st.rng.state = ((uint64_t) r.y << 32) | r.x;
    st.rng.inc = ((uint64_t) r.w << 32) | r.z;
}

__device__ __forceinline__ void wf_store_state(const WfBuf &B, uint32_t slot, const PtState &st) {
    B.S0[slot] = make_float4(st.o.x, st.o.y, st.o.z, st.px);
    B.S1[slot] = make_float4(st.d.x, st.d.y, st.d.z, st.py);
    B.S2[slot] = make_float4(st.thr.x, st.thr.y, st.thr.z, st.eta);
    B.S3[slot] = make_float4(st.res.x, st.res.y, st.res.z, st.prev_pdf);
    B.S4[slot] = make_float4(st.prev_p.x, st.prev_p.y, st.prev_p.z, __int_as_float(st.depth | ((int) st.prev_delta << 16)));
    B.RNG[slot] = make_uint4((uint32_t) st.rng.state, (uint32_t) (st.rng.state >> 32), (uint32_t) st.rng.inc, (uint32_t) (st.rng.inc >> 32));
}

// slot -> pixel: a layer (one sample of every pixel) is laid out tile by tile, 256 slots per 16 x 16 tile, and the 32
// consecutive slots of a warp form an 8 x 4 pixel block (coherent camera rays)
__device__ __forceinline__ bool wf_slot_pixel(const PtDev &P, uint32_t r, int &x, int &y) {
    const int tile = (int) (r >> 8), it = (int) (r & 255u), w = it >> 5, lane = it & 31;
    x = (tile % P.tiles_x) * PT_TILE + (w & 1) * 8 + (lane & 7);
    y = (tile / P.tiles_x) * PT_TILE + (w >> 1) * 4 + (lane >> 3);
    return x < P.W && y < P.H;
}

__device__ __forceinline__ void wf_add_stat(const PtDev &P, int which, unsigned v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    if ((threadIdx.x & 31) == 0 && v && P.stats) atomicAdd(P.stats + which, (unsigned long long) v);
}

__global__ void __launch_bounds__(WF_SHADE_THREADS) k_wf_generate(const PtDev P, const WfBuf B) {
    const uint32_t n = B.n_layers * B.L;      // multiple of 256
    unsigned made = 0;
    for (uint32_t s0 = blockIdx.x * blockDim.x; s0 < n; s0 += gridDim.x * blockDim.x) {
        const uint32_t slot = s0 + threadIdx.x;
        const uint32_t layer = slot / B.L, r = slot - layer * B.L;
        int x, y;
        const bool inside = wf_slot_pixel(P, r, x, y);
        if (inside) {
            PtState st;
            pt_init(P, x, y, P.s_offset + (B.j0 + layer) * P.s_stride, st);
            wf_store_state(B, slot, st);
            made++;
        }
        const int q = wf_reserve(B.cnt + C_EXT, inside);
        if (inside) B.q_ext[0][q] = slot;
    }
    wf_add_stat(P, 0, made);
}

// ------------------------------------------------------------------------------------------------------------------
// ray queries with dynamic fetch
// ------------------------------------------------------------------------------------------------------------------
template <bool ANY, bool W8>
__global__ void __launch_bounds__(WF_TRACE_THREADS) k_wf_trace(const PtDev P, const WfBuf B, const int bounce) {
    __shared__ DPrim sprims[MAX_SMEM_PRIMS];
    const DScene &sc = P.sc;
    const DPrim *prims = sprims;
    if (sc.n_prims > MAX_SMEM_PRIMS) prims = sc.prims;
    else {
        const float4 *src = reinterpret_cast<const float4 *>(sc.prims);
        float4 *dst = reinterpret_cast<float4 *>(sprims);
        for (int i = threadIdx.x; i < sc.n_prims * (int) (sizeof(DPrim) / 16); i += blockDim.x) dst[i] = src[i];
    }
    __syncthreads();
    int *C = B.cnt + bounce * WF_CSTRIDE;
    const int n = ANY ? C[C_SH] : C[C_EXT];
    int *head = C + (ANY ? C_HEAD_SH : C_HEAD_EXT);
    const uint32_t *queue = B.q_ext[bounce & 1];
    const int lane = threadIdx.x & 31;
    const int DONE = 0x7fffffff;

    int pool_next = 0, pool_end = 0;     // warp-uniform: queue positions this warp has reserved
    bool dry = false;                    // warp-uniform: the queue is exhausted
    bool has = false;
    uint32_t slot = 0;                   // path slot (closest) / shadow-queue position (any)
    float3 o = mk3(0, 0, 0), d = mk3(0, 0, 1);
    RayPre rp = ray_precompute(d);
    float tbest = 0.0f, b1 = 0.0f, b2 = 0.0f, prim_t = 0.0f;
    int best = -1, best_prim = -1, sp = 0;
    // BVH2 state (W8 == false)
    float3 inv = mk3(0, 0, 0);
    int ref = DONE;
    int   stack_ref[W8 ? 1 : PRT_STACK];
    float stack_t[W8 ? 1 : PRT_STACK];
    // BVH8 state (W8 == true): current node group (child base, hit bits | imask), its triangle group, stack of groups
    Bvh8Ray r8 = bvh8_ray(o, d);
    uint2 ng = make_uint2(0, 0);
    uint2 gstack[W8 ? BVH8_STACK : 1];
    bool busy = false;
    unsigned n_rays = 0, n_valid = 0;

    for (;;) {
        // ---- retire finished rays ----
        const bool fin = has && (W8 ? !busy : ref == DONE);
        if (__any_sync(FULL, fin)) {
            if (!ANY) {
                int qi = -1;
                if (fin) {
                    int id = -1, material = 0;
                    float t = tbest;
                    if (best >= 0 && (best_prim < 0 || tbest < prim_t)) {
                        const int sorted = W8 ? (int) __ldg(sc.tri8_sorted + best) : best;
                        id = sc.n_prims + sorted;
                        material = __ldg(&sc.tri_info[sorted].z);
                    } else if (best_prim >= 0) {
                        id = best_prim;
                        t = prim_t;
                        material = prims[best_prim].material;
                    }
                    B.HIT[slot] = make_float4(t, b1, b2, __int_as_float(id));
                    if (id >= 0) {
                        const int kind = __ldg(&sc.mats[material].kind);
                        qi = kind == PRT_MAT_DIFFUSE ? 0 : (kind == PRT_MAT_DIELECTRIC ? 1 : 2);
                        n_valid++;
                    }
                }
#pragma unroll
                for (int k = 0; k < WF_QUEUES; k++) {
                    const int q = wf_reserve(C + C_MAT + k, qi == k);
                    if (qi == k) B.q_mat[k][q] = slot;
                }
            } else if (fin && best < 0) {
                const float4 c = B.SH2[slot];
                const float w = B.SH1[slot].w;
                const uint32_t ps = (uint32_t) __float_as_int(c.w);
                float4 r = B.S3[ps];
                r.x = fmaf(c.x, w, r.x);
                r.y = fmaf(c.y, w, r.y);
                r.z = fmaf(c.z, w, r.z);
                B.S3[ps] = r;
            }
            if (fin) has = false;
        }
        // ---- refill idle lanes ----
        if (!dry) {
            unsigned need = __ballot_sync(FULL, !has);
            while (need) {
                if (pool_next == pool_end) {
                    int base = 0;
                    if (lane == 0) base = atomicAdd(head, WF_CHUNK);
                    base = __shfl_sync(FULL, base, 0);
                    pool_next = base;
                    pool_end = min(base + WF_CHUNK, n);
                    if (pool_next >= pool_end) {
                        pool_next = pool_end = 0;
                        dry = true;
                        break;
                    }
                }
                const int avail = pool_end - pool_next;
                const int rank = __popc(need & lanemask_lt());
                const bool take = !has && rank < avail;
                if (take) {
                    const int idx = pool_next + rank;
                    if (!ANY) {
                        slot = queue[idx];
                        const float4 a = B.S0[slot], bb = B.S1[slot];
                        o = xyz(a);
                        d = xyz(bb);
                        tbest = PRT_INF;
                    } else {
                        slot = (uint32_t) idx;
                        const float4 a = B.SH0[idx], bb = B.SH1[idx];
                        o = xyz(a);
                        d = xyz(bb);
                        tbest = a.w;
                    }
                    has = true;
                    n_rays++;
                    best = -1;
                    best_prim = -1;
                    prim_t = tbest;
                    bool blocked = false;
                    for (int i = 0; i < sc.n_prims; i++) {
                        const float t = intersect_prim(prims[i], o, d, prim_t);
                        if (t >= 0.0f && (best_prim < 0 || t < prim_t)) {
                            best_prim = i;
                            prim_t = t;
                            if (ANY) blocked = true;
                        }
                    }
                    if (!ANY) tbest = prim_t;
                    rp = ray_precompute(d);
                    sp = 0;
                    const bool go = !(sc.n_tris == 0 || blocked);
                    if (W8) {
                        r8 = bvh8_ray(o, d);
                        ng = make_uint2(0u, 0x80000000u);
                        busy = go;
                    } else {
                        inv = mk3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
                        ref = go ? sc.root_ref : DONE;
                    }
                    if (ANY && blocked) best = 0;
                }
                pool_next += min(avail, __popc(need));
                need = __ballot_sync(FULL, !has);
            }
        }
        if (!__any_sync(FULL, has)) break;

        if (W8) {
            // ---- one wide node per lane and iteration, then the triangles it yielded ----
            if (busy) {
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
                    const uint32_t ti = tg.x + (uint32_t) bit;
                    const float4 *tv = sc.tri_v8 + 3 * (size_t) ti;
                    const float4 a = ldg4(tv), b = ldg4(tv + 1), c = ldg4(tv + 2);
                    if (intersect_tri_wt(rp, o, xyz(a), xyz(b), xyz(c), tbest, b1, b2)) {
                        best = (int) ti;
                        if (ANY) { busy = false; break; }
                    }
                }
                if (busy && ng.y <= 0x00ffffffu) {
                    if (sp > 0) ng = gstack[--sp];
                    else busy = false;
                }
            }
        } else {
#define PRT_POP()                                                        \
    do {                                                                 \
        ref = DONE;                                                      \
        while (sp > 0) {                                                 \
            --sp;                                                        \
            if (stack_t[sp] <= tbest) { ref = stack_ref[sp]; break; }    \
        }                                                                \
    } while (0)
            // ---- inner nodes: every lane descends until it holds a leaf or is done ----
            while ((unsigned) ref < (unsigned) DONE) {
                const float4 *nd = sc.nodes + 4 * (size_t) ref;
                const float4 q0 = ldg4(nd), q1 = ldg4(nd + 1), q2 = ldg4(nd + 2), q3 = ldg4(nd + 3);
                const float tl = box_entry(q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, o, inv, tbest);
                const float tr = box_entry(q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, o, inv, tbest);
                const int rl = __float_as_int(q3.x), rr = __float_as_int(q3.y);
                const bool hl = tl < PRT_INF, hr = tr < PRT_INF;
                if (hl && hr) {
                    const bool lf = tl <= tr;
                    if (sp < PRT_STACK) {
                        stack_ref[sp] = lf ? rr : rl;
                        stack_t[sp] = lf ? tr : tl;
                        sp++;
                    }
                    ref = lf ? rl : rr;
                } else if (hl || hr) {
                    ref = hl ? rl : rr;
                } else {
                    PRT_POP();
                }
            }
            // ---- leaf ----
            if (ref < 0) {
                const int code = ~ref;
                const int first = code >> 2, count = (code & 3) + 1;
                bool stop = false;
                for (int j = 0; j < count; j++) {
                    const float4 *tv = sc.tri_v + 3 * (size_t) (first + j);
                    const float4 a = ldg4(tv), b = ldg4(tv + 1), c = ldg4(tv + 2);
                    if (intersect_tri_wt(rp, o, xyz(a), xyz(b), xyz(c), tbest, b1, b2)) {
                        best = first + j;
                        if (ANY) { stop = true; break; }
                    }
                }
                if (stop) ref = DONE;
                else PRT_POP();
            }
#undef PRT_POP
        }
    }
    wf_add_stat(P, 2, n_rays);
    if (ANY) wf_add_stat(P, 3, n_rays);
    else wf_add_stat(P, 1, n_valid);
}

// ------------------------------------------------------------------------------------------------------------------
// shading, one kernel per material queue
// ------------------------------------------------------------------------------------------------------------------
template <int QI>
__global__ void __launch_bounds__(WF_SHADE_THREADS) k_wf_shade(const PtDev P, const WfBuf B, const int bounce) {
    int *C = B.cnt + bounce * WF_CSTRIDE;
    int *Cn = C + WF_CSTRIDE;
    const int n = C[C_MAT + QI];
    const uint32_t *q = B.q_mat[QI];
    uint32_t *qn = B.q_ext[(bounce + 1) & 1];
    for (int i0 = blockIdx.x * blockDim.x; i0 < n; i0 += gridDim.x * blockDim.x) {
        const int i = i0 + threadIdx.x;
        bool live = false;
        uint32_t slot = 0;
        ShadowReq sr;
        sr.want = false;
        if (i < n) {
            slot = q[i];
            PtState st;
            wf_load_state(B, slot, st);
            const float4 hv = B.HIT[slot];
            const int id = __float_as_int(hv.w);
            Hit h;
            if (id >= P.sc.n_prims) fill_tri_hit(P.sc, id - P.sc.n_prims, hv.x, hv.y, hv.z, h);
            else fill_prim_hit(P.sc.prims[id], id, st.o, st.d, hv.x, h);
            live = pt_shade(P, st, h, true, sr);
            wf_store_state(B, slot, st);
        }
        const int j = wf_reserve(C + C_SH, sr.want);
        if (sr.want) {
            B.SH0[j] = make_float4(sr.o.x, sr.o.y, sr.o.z, sr.tmax);
            B.SH1[j] = make_float4(sr.d.x, sr.d.y, sr.d.z, sr.w);
            B.SH2[j] = make_float4(sr.c.x, sr.c.y, sr.c.z, __int_as_float((int) slot));
        }
        const int e = wf_reserve(Cn + C_EXT, live);
        if (live) qn[e] = slot;
    }
}

// ------------------------------------------------------------------------------------------------------------------
// film: one CTA per pixel tile, all layers of the batch
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_wf_film(const PtDev P, const WfBuf B) {
    __shared__ float4 tile[PT_HALO * PT_HALO];
    const int t = blockIdx.x;
    const int tx0 = (t % P.tiles_x) * PT_TILE, ty0 = (t / P.tiles_x) * PT_TILE;
    for (int i = threadIdx.x; i < PT_HALO * PT_HALO; i += blockDim.x) tile[i] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    __syncthreads();
    const uint32_t r = (uint32_t) t * 256u + threadIdx.x;
    int x, y;
    if (wf_slot_pixel(P, r, x, y)) {
        for (uint32_t layer = 0; layer < B.n_layers; layer++) {
            const uint32_t slot = layer * B.L + r;
            const float px = B.S0[slot].w, py = B.S1[slot].w;
            const float4 res = B.S3[slot];
            pt_splat(P.tent, tile, tx0, ty0, px, py, xyz(res));
        }
    }
    __syncthreads();
    pt_flush_tile(P, tile, tx0, ty0);
}

static int wf_grid(prt_context *c, const void *kernel, int threads, int *grid) {
    int per_sm = 0;
    PRT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, 0));
    if (per_sm < 1) per_sm = 1;
    *grid = c->sm_count * per_sm;
    return PRT_OK;
}

int launch_wavefront(prt_context *c, const PtDev &P, cudaStream_t st) {
    const uint32_t n_tiles = (uint32_t) P.tiles_x * (uint32_t) P.tiles_y;
    const uint64_t L = (uint64_t) n_tiles * 256u;
    uint64_t batch = 1ull << 24;
    if (const char *e = getenv("PRT_WF_BATCH")) {
        long long v = atoll(e);
        if (v > 0) batch = (uint64_t) v;
    }
    uint64_t layers = batch / L;
    if (layers < 1) layers = 1;
    if (layers > P.n_s) layers = P.n_s ? P.n_s : 1;
    const uint64_t cap = layers * L;
    PRT_REQUIRE(cap < (1ull << 31), "render_path (wavefront): batch too large");
    const int bounces = P.max_depth > 1 ? P.max_depth : 1;
    const size_t cnt_bytes = sizeof(int) * WF_CSTRIDE * (size_t) (bounces + 1);
    const size_t per_slot = 16 * 10 + 4 * (2 + WF_QUEUES);
    const size_t need = (size_t) cap * per_slot + ((cnt_bytes + 255) & ~(size_t) 255);
    if (need > c->wf_cap) {
        if (c->wf_dev) cudaFree(c->wf_dev);
        c->wf_dev = nullptr;
        c->wf_cap = 0;
        PRT_CUDA(cudaMalloc(&c->wf_dev, need));
        c->wf_cap = need;
    }
    WfBuf B;
    {
        char *p = reinterpret_cast<char *>(c->wf_dev);
        auto take = [&](size_t bytes) { char *r = p; p += bytes; return r; };
        B.cnt = reinterpret_cast<int *>(take((cnt_bytes + 255) & ~(size_t) 255));
        B.S0 = reinterpret_cast<float4 *>(take(16 * cap));
        B.S1 = reinterpret_cast<float4 *>(take(16 * cap));
        B.S2 = reinterpret_cast<float4 *>(take(16 * cap));
        B.S3 = reinterpret_cast<float4 *>(take(16 * cap));
        B.S4 = reinterpret_cast<float4 *>(take(16 * cap));
        B.RNG = reinterpret_cast<uint4 *>(take(16 * cap));
        B.HIT = reinterpret_cast<float4 *>(take(16 * cap));
        B.SH0 = reinterpret_cast<float4 *>(take(16 * cap));
        B.SH1 = reinterpret_cast<float4 *>(take(16 * cap));
        B.SH2 = reinterpret_cast<float4 *>(take(16 * cap));
        for (int k = 0; k < 2; k++) B.q_ext[k] = reinterpret_cast<uint32_t *>(take(4 * cap));
        for (int k = 0; k < WF_QUEUES; k++) B.q_mat[k] = reinterpret_cast<uint32_t *>(take(4 * cap));
    }
    B.cap = (uint32_t) cap;
    B.L = (uint32_t) L;
    int g_gen = 1, g_ext = 1, g_sh = 1, g_shade[WF_QUEUES] = { 1, 1, 1 };
    int rc;
    if ((rc = wf_grid(c, (const void *) k_wf_generate, WF_SHADE_THREADS, &g_gen))) return rc;
    // BVH8c by default when it exists; PRT_BVH=2 keeps the binary tree (A/B runs, parity tests between the two)
    const char *bsel = getenv("PRT_BVH");
    const bool w8 = P.sc.n_nodes8 > 0 && !(bsel && bsel[0] == '2');
    const void *k_ext = w8 ? (const void *) k_wf_trace<false, true> : (const void *) k_wf_trace<false, false>;
    const void *k_sh = w8 ? (const void *) k_wf_trace<true, true> : (const void *) k_wf_trace<true, false>;
    if ((rc = wf_grid(c, k_ext, WF_TRACE_THREADS, &g_ext))) return rc;
    if ((rc = wf_grid(c, k_sh, WF_TRACE_THREADS, &g_sh))) return rc;
    if ((rc = wf_grid(c, (const void *) k_wf_shade<0>, WF_SHADE_THREADS, &g_shade[0]))) return rc;
    if ((rc = wf_grid(c, (const void *) k_wf_shade<1>, WF_SHADE_THREADS, &g_shade[1]))) return rc;
    if ((rc = wf_grid(c, (const void *) k_wf_shade<2>, WF_SHADE_THREADS, &g_shade[2]))) return rc;
    int launches = 0;
    for (uint64_t j0 = 0; j0 < P.n_s; j0 += layers) {
        B.j0 = (uint32_t) j0;
        B.n_layers = (uint32_t) (P.n_s - j0 < layers ? P.n_s - j0 : layers);
        PRT_CUDA(cudaMemsetAsync(B.cnt, 0, cnt_bytes, st));
        k_wf_generate<<<g_gen, WF_SHADE_THREADS, 0, st>>>(P, B);
        launches++;
        for (int b = 0; b < bounces; b++) {
            if (w8) k_wf_trace<false, true><<<g_ext, WF_TRACE_THREADS, 0, st>>>(P, B, b);
            else k_wf_trace<false, false><<<g_ext, WF_TRACE_THREADS, 0, st>>>(P, B, b);
            launches++;
            if (P.kind_mask & (1u << PRT_MAT_DIFFUSE)) { k_wf_shade<0><<<g_shade[0], WF_SHADE_THREADS, 0, st>>>(P, B, b); launches++; }
            if (P.kind_mask & (1u << PRT_MAT_DIELECTRIC)) { k_wf_shade<1><<<g_shade[1], WF_SHADE_THREADS, 0, st>>>(P, B, b); launches++; }
            if (P.kind_mask & ~((1u << PRT_MAT_DIFFUSE) | (1u << PRT_MAT_DIELECTRIC))) {
                k_wf_shade<2><<<g_shade[2], WF_SHADE_THREADS, 0, st>>>(P, B, b);
                launches++;
            }
            if (b + 1 < P.max_depth && (P.kind_mask & (1u << PRT_MAT_DIFFUSE)) && P.sc.n_emitters > 0) {
                if (w8) k_wf_trace<true, true><<<g_sh, WF_TRACE_THREADS, 0, st>>>(P, B, b);
                else k_wf_trace<true, false><<<g_sh, WF_TRACE_THREADS, 0, st>>>(P, B, b);
                launches++;
            }
        }
        k_wf_film<<<n_tiles, 256, 0, st>>>(P, B);
        launches++;
        PRT_CUDA(cudaGetLastError());
    }
    c->last_launches = launches;
    return PRT_OK;
}

}  // namespace prt
