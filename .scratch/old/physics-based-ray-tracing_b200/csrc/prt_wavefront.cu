// prt_wavefront.cu -- wavefront form of the light-transport path tracer (BASELINE
// configs 4 and 5; north star item 3: "ray-gen, BVH traversal, per-material shading queues compacted with warp
// ballot/prefix-sum, shadow-ray and emitter-NEE kernels, film accumulation").
//
// It computes exactly what prt_path.cu's tile megakernel computes -- both run prt_path.h's pt_init / pt_shade on
// the same per-path PCG32 streams (the Mitsuba `path` integrator of /root/reference/scenes/cbox.xml:5-9, restated
// from SURVEY.md Appendix C.7) -- but moves the two ray queries into their own persistent kernels:
//
//   k_wf_generate            camera rays of a batch of samples -> path state (SoA, float4 per field) + ray records
//   per bounce b:
//     k_wf_trace<false>      closest hit over the compressed 8-wide BVH (prt_bvh8.cuh).  Persistent warps pull ray
//                            records from the queue in chunks; a lane whose ray is finished retires it and takes the
//                            next one, so the heavy-tailed traversal length of an incoherent ray no longer idles the
//                            other 31 lanes.  Retiring appends the path to the shading queue of the material it hit
//                            (warp ballot + one atomic); the triangle record carries that queue and its own index, so
//                            retiring costs no dependent loads.
//     k_wf_shade<queue>      one thread per queued path: rebuilds the surface interaction from (t, b1, b2, id), runs
//                            pt_shade, appends the shadow-ray record to the shadow queue and the next ray's record to
//                            the next bounce's extend queue (both ballot-compacted, so records are dense and a warp's
//                            refill reads consecutive 16-byte words).  Everything a ray needs in the traversal loop
//                            (origin, reciprocal direction, the shear constants of the watertight triangle test) is
//                            precomputed HERE, where all 32 lanes are busy, not in the trace kernel's refill path.
//     k_wf_trace<true>       shadow rays (any hit), same dynamic fetch; unoccluded ones add their NEE term
//   k_wf_film                one CTA per 16 x 16 pixel tile splats the tile's finished samples into shared memory
//                            (tent filter) and adds the tile to the film
//
// All launches of a batch are enqueued back to back: queue lengths live in device memory (one counter block per
// bounce, zeroed once), so the host never synchronises inside a batch.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "prt_bvh8.cuh"
#include "prt_internal.h"
#include "prt_path.h"

namespace prt {

static constexpr unsigned FULL = 0xffffffffu;
static constexpr int WF_QUEUES = 3;        // shading queues: 0 diffuse, 1 dielectric, 2 everything else (conductor, null)
static constexpr int WF_CSTRIDE = 16;      // ints per bounce in the counter array
#ifndef WF_CHUNK
#define WF_CHUNK 64                        // rays a warp reserves per atomic on the queue head
#endif
static constexpr int WF_TRACE_THREADS = 128;
static constexpr int WF_SHADE_THREADS = 256;
#ifndef WF_SHADE_MINB
#define WF_SHADE_MINB 2                   // min resident CTAs per SM the shading kernels are compiled for
#endif
#ifndef WF_COOP_MAX
#define WF_COOP_MAX 16                     // warp-cooperative triangle tests while at most this many lanes hold triangles
#endif
#ifndef WF_REFILL_MIN
#define WF_REFILL_MIN 1                    // idle lanes before a warp goes back to the queue
#endif
#ifndef WF_TRACE_MINB
#define WF_TRACE_MINB 8                   // min resident CTAs per SM the trace kernels are compiled for
#endif
#ifndef WF_QCHUNK
#define WF_QCHUNK 32                      // shading-queue entries a trace warp reserves per atomic (0: one atomic per retire event)
#endif
#ifndef WF_SLOT_SHADE
#define WF_SLOT_SHADE 70                  // a shading kernel walks the path SLOTS in order (picking its material by the per-slot tag)
#endif                                    // when its queue holds more than this percentage of the batch's slots, else the compacted queue; 0: never
#ifndef WF_SSTACK
#define WF_SSTACK 0                       // traversal-stack entries per lane kept in shared memory (the rest spills to local memory)
#endif
#ifndef WF_PREFETCH
#define WF_PREFETCH 0                     // bit 0: ray records of a reserved chunk -> L2; bit 1: next node -> L1; bit 2: triangles -> L1
#endif
static_assert(WF_QCHUNK == 0 || WF_QCHUNK >= 32, "a retire event appends up to 32 entries: one fresh chunk must hold them");
static constexpr uint32_t WF_HOLE = 0xffffffffu;   // unused tail of a warp's reserved queue chunk
// per-bounce counter block.  C_SH is the LAST int of a block and C_EXT the first of the next one, and the array starts one
// int past an 8-byte boundary: the shadow-ray count of bounce b and the extend-ray count of bounce b + 1 form one aligned
// 64-bit word, so a shading warp reserves both of its output ranges with a single atomic.
enum { C_EXT = 0, C_MAT = 1, C_HEAD_EXT = 8, C_HEAD_SH = 12, C_SH = WF_CSTRIDE - 1 };

// The wavefront's own streams (ray records, path state, hit records, queues) are written once and read once, gigabytes
// apart; evict-first hints (ld/st.global.cs) were meant to stop them displacing BVH nodes and triangles from L2.  Measured
// on B200: no gain (height field 1 910 vs 1 932 Mrays/s, cbox 6 540 vs 6 577) -- L2's own replacement already keeps the
// hot upper levels.  Left as a build knob, off.
#ifndef WF_SHADE_KSEL
#define WF_SHADE_KSEL 1                   // shading kernels 0 / 1 are compiled for their one material kind
#endif
#ifndef WF_SHADE_PREFETCH
#define WF_SHADE_PREFETCH 0
#endif
#ifndef WF_STREAM_HINTS
#define WF_STREAM_HINTS 0
#endif
#if WF_STREAM_HINTS
__device__ __forceinline__ float4 ld_stream(const float4 *p) { return __ldcs(p); }
__device__ __forceinline__ void st_stream(float4 *p, float4 v) { __stcs(p, v); }
__device__ __forceinline__ void st_stream(uint32_t *p, uint32_t v) { __stcs(p, v); }
#else
__device__ __forceinline__ float4 ld_stream(const float4 *p) { return *p; }
__device__ __forceinline__ void st_stream(float4 *p, float4 v) { *p = v; }
__device__ __forceinline__ void st_stream(uint32_t *p, uint32_t v) { *p = v; }
#endif

__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

// ray record, 4 x float4, written by the producer at the ray's queue position:
//   r0 = origin xyz, bits(path slot)          r1 = direction xyz, tmax
//   r2 = 1/direction xyz (clamped, bvh8_ray), extra (shadow rays: MIS weight)
//   r3 = Sx, Sy, Sz, bits(kx | ky << 2 | kz << 4)     (ray_precompute: watertight triangle test)
struct WfRays {
    float4 *r0, *r1, *r2, *r3;
    uint32_t *key;   // sort key of the ray (origin cell | direction octant), nullptr when ray sorting is off
};

struct WfBuf {
    // path state: one 128-byte record (= one cache line) per slot, so the shading kernels' random access by slot
    // costs one line instead of seven sectors in seven lines
    //   [0] o.xyz, px   [1] d.xyz, py   [2] throughput rgb, eta   [3] radiance rgb, prev_pdf
    //   [4] prev_p.xyz, bits(depth | prev_delta << 16)   [5] pcg32 state, inc
    //   [6] hit: t, b1, b2, bits(id) (id < 0 miss, < n_prims analytic primitive, else n_prims + sorted triangle)
    float4 *ST;
    WfRays ext[2];   // extend rays of bounce b live in ext[b & 1]
    WfRays sh;       // shadow rays of the current bounce
    float4 *SHC;     // shadow rays: NEE contribution rgb
    uint32_t *q_mat[WF_QUEUES];
    uint8_t *tag;    // [cap] per slot: 1 + shading queue of the hit waiting to be shaded, 0 = nothing to shade
    int *cnt;        // [bounces + 1][WF_CSTRIDE]
    uint32_t cap, L, n_layers, j0;
    // ray sorting (big scenes): key = Morton code of the origin's cell (sort_bits per axis over the scene box) and the
    // direction octant; rays are traced in key order through a permutation of the queue positions
    int sort_bits, sort_mode;            // bits per axis (0 = off); mode 0: cell major, octant minor; 1: octant major
    float3 key_lo, key_scale;
};

// analytic primitives staged in shared memory by the trace kernels (all of them, or none if there are too many)
__host__ __device__ __forceinline__ int sc_n_smem_prims(const DScene &sc) { return sc.n_prims > MAX_SMEM_PRIMS ? 0 : sc.n_prims; }
static size_t wf_trace_smem(const DScene &sc) {
    return (size_t) sc_n_smem_prims(sc) * sizeof(DPrim) + sizeof(uint2) * WF_SSTACK * WF_TRACE_THREADS;
}

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

__device__ __forceinline__ uint32_t wf_part1by2(uint32_t x) {      // 10 bits -> every third bit
    x &= 0x3ffu;
    x = (x | (x << 16)) & 0x030000ffu;
    x = (x | (x << 8)) & 0x0300f00fu;
    x = (x | (x << 4)) & 0x030c30c3u;
    x = (x | (x << 2)) & 0x09249249u;
    return x;
}

__device__ __forceinline__ uint32_t wf_ray_key(const WfBuf &B, float3 o, float3 d) {
    const float top = (float) ((1 << B.sort_bits) - 1);
    const uint32_t ix = (uint32_t) fminf(fmaxf((o.x - B.key_lo.x) * B.key_scale.x, 0.0f), top);
    const uint32_t iy = (uint32_t) fminf(fmaxf((o.y - B.key_lo.y) * B.key_scale.y, 0.0f), top);
    const uint32_t iz = (uint32_t) fminf(fmaxf((o.z - B.key_lo.z) * B.key_scale.z, 0.0f), top);
    const uint32_t cell = (wf_part1by2(ix) << 2) | (wf_part1by2(iy) << 1) | wf_part1by2(iz);
    const uint32_t oct = (d.x < 0.0f ? 4u : 0u) | (d.y < 0.0f ? 2u : 0u) | (d.z < 0.0f ? 1u : 0u);
    return B.sort_mode ? (oct << (3 * B.sort_bits)) | cell : (cell << 3) | oct;
}

__device__ __forceinline__ void wf_write_ray(const WfBuf &B, const WfRays &R, int pos, float3 o, float3 d, float tmax, uint32_t slot, float extra) {
    if (R.key) R.key[pos] = wf_ray_key(B, o, d);
    const RayPre rp = ray_precompute(d);
    const Bvh8Ray r8 = bvh8_ray(o, d);
    st_stream(R.r0 + pos, make_float4(o.x, o.y, o.z, __uint_as_float(slot)));
    st_stream(R.r1 + pos, make_float4(d.x, d.y, d.z, tmax));
    st_stream(R.r2 + pos, make_float4(r8.inv.x, r8.inv.y, r8.inv.z, extra));
    st_stream(R.r3 + pos, make_float4(rp.Sx, rp.Sy, rp.Sz, __int_as_float(rp.kx | (rp.ky << 2) | (rp.kz << 4))));
}

__device__ __forceinline__ void wf_load_state(const WfBuf &B, uint32_t slot, PtState &st) {
    const float4 *rec = B.ST + 8 * (size_t) slot;
    float4 a = ld_stream(rec), b = ld_stream(rec + 1), c = ld_stream(rec + 2), d = ld_stream(rec + 3), e = ld_stream(rec + 4);
    const float4 rr = ld_stream(rec + 5);
    uint4 r = make_uint4(__float_as_uint(rr.x), __float_as_uint(rr.y), __float_as_uint(rr.z), __float_as_uint(rr.w));
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
    float4 *rec = B.ST + 8 * (size_t) slot;
    st_stream(rec, make_float4(st.o.x, st.o.y, st.o.z, st.px));
    st_stream(rec + 1, make_float4(st.d.x, st.d.y, st.d.z, st.py));
    st_stream(rec + 2, make_float4(st.thr.x, st.thr.y, st.thr.z, st.eta));
    st_stream(rec + 3, make_float4(st.res.x, st.res.y, st.res.z, st.prev_pdf));
    st_stream(rec + 4, make_float4(st.prev_p.x, st.prev_p.y, st.prev_p.z, __int_as_float(st.depth | ((int) st.prev_delta << 16))));
    st_stream(rec + 5, make_float4(__uint_as_float((uint32_t) st.rng.state), __uint_as_float((uint32_t) (st.rng.state >> 32)),
                         __uint_as_float((uint32_t) st.rng.inc), __uint_as_float((uint32_t) (st.rng.inc >> 32))));
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
        PtState st;
        if (inside) {
            pt_init(P, x, y, P.s_offset + (B.j0 + layer) * P.s_stride, st);
            wf_store_state(B, slot, st);
            made++;
        }
        const int q = wf_reserve(B.cnt + C_EXT, inside);
        if (inside) wf_write_ray(B, B.ext[0], q, st.o, st.d, PRT_INF, slot, 0.0f);
    }
    wf_add_stat(P, 0, made);
}

// ------------------------------------------------------------------------------------------------------------------
// ray queries with dynamic fetch over the compressed 8-wide BVH
// ------------------------------------------------------------------------------------------------------------------
// triangle `off` of a group: base < 2^31 -> BVH8 leaf order (tri_v8); bit 31 set -> the oversized triangles in tri_v
__device__ __forceinline__ const float4 *wf_tri_ptr(const DScene &sc, uint32_t base, uint32_t off) {
    const float4 *arr = (base >> 31) ? sc.tri_v : sc.tri_v8;
    return arr + 3 * (size_t) ((base & 0x7fffffffu) + off);
}

#ifndef WF_TRI_PEEL
#define WF_TRI_PEEL 0                    // per-lane triangle rounds only while > WF_COOP_MAX lanes have one; tails go to the cooperative test
#endif
#ifndef WF_TRI_DEFER
#define WF_TRI_DEFER 0                   // > 0: a lane keeps the triangles its node step yielded and waits; the warp tests them only in
#endif                                    // full rounds of 32 (ray, triangle) pairs, or when no lane can advance otherwise
#ifndef WF_TOPN
#define WF_TOPN 0                         // first WF_TOPN nodes of the 8-wide BVH (breadth-first: 73 = top three levels) staged in shared memory
#endif
template <bool ANY>
__global__ void __launch_bounds__(WF_TRACE_THREADS, WF_TRACE_MINB) k_wf_trace(const PtDev P, const WfBuf B, const int bounce,
                                                                              const uint32_t *__restrict__ perm) {
    // dynamic shared memory: [analytic primitives (n_prims x 128 B; none for pure mesh scenes, which leaves that much more
    // of the SM's unified array to L1)][optionally the first WF_SSTACK traversal-stack entries of every lane, entry-major].
    // The stack in local memory misses L1 on 89 % of the pops (ncu r01), yet moving it to shared memory bought nothing
    // measurable on B200 (4 / 6 / 8 / 12 entries: 1.85-1.87 vs 1.90 Grays/s without): the L1 capacity it takes away costs
    // as much as the pops it saves.  Default 0.
    extern __shared__ float4 s_dyn[];
    DPrim *sprims = reinterpret_cast<DPrim *>(s_dyn);
    uint2 *sstack = reinterpret_cast<uint2 *>(s_dyn + (sc_n_smem_prims(P.sc) * (int) (sizeof(DPrim) / 16))) + threadIdx.x;
    __shared__ int s_win[WF_TRACE_THREADS];
    __shared__ unsigned s_tmin[WF_TRACE_THREADS];
#if WF_TOPN
    // The top levels are fetched by every ray of every warp.  They hit in L1 anyway (ncu r01: top of the tree is < 6 KB), so
    // staging them is worth one L1 -> shared latency difference per visit, nothing more: measured in profiles/r02_summary.md
    __shared__ float4 s_top[5 * WF_TOPN];
    {
        const int ntop = min(WF_TOPN, P.sc.n_nodes8);
        for (int i = threadIdx.x; i < 5 * ntop; i += blockDim.x) s_top[i] = P.sc.nodes8[i];
    }
#endif
#if WF_QCHUNK
    // warp-private cursors into the shading queues: a warp reserves WF_QCHUNK entries per atomic and fills them over
    // several retire events, so retiring a ray does not wait for a round trip to the L2 atomic unit
    __shared__ int s_qcur[WF_TRACE_THREADS / 32][2 * WF_QUEUES];
    if (!ANY && (threadIdx.x & 31) < 2 * WF_QUEUES) s_qcur[threadIdx.x >> 5][threadIdx.x & 31] = 0;
#endif
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
    const WfRays R = ANY ? B.sh : B.ext[bounce & 1];
    const int lane = threadIdx.x & 31;
    int *sw = s_win + (threadIdx.x & ~31);
    unsigned *stm = s_tmin + (threadIdx.x & ~31);

    int pool_next = 0, pool_end = 0;     // warp-uniform: queue positions this warp has reserved
    bool dry = false;                    // warp-uniform: the queue is exhausted
    bool has = false, busy = false;      // lane holds a ray / its traversal is still running
    bool fresh = false;                  // the ray has not been tested against the oversized triangles yet
    uint32_t slot = 0, qpos = 0;         // path slot, queue position of the ray
    Bvh8Ray r8;
    r8.o = mk3(0, 0, 0); r8.inv = mk3(1, 1, 1); r8.octinv4 = 0;
    RayPre rp;
    rp.kx = 0; rp.ky = 1; rp.kz = 2; rp.Sx = rp.Sy = 0.0f; rp.Sz = 1.0f;
    int kpack = 0;

    float tbest = 0.0f, b1 = 0.0f, b2 = 0.0f, prim_t = 0.0f;
    int best = -1, best_prim = -1, sp = 0;      // best: (sorted triangle << 2) | shading queue, or -1
    uint2 ng = make_uint2(0, 0);                // node group in hand: child base, hit bits | imask
    uint2 tg = make_uint2(0u, 0u);              // triangle group in hand: base, 24-bit mask (consumed by the triangle phase)
    uint2 gstack[BVH8_STACK > WF_SSTACK ? BVH8_STACK - WF_SSTACK : 1];
    unsigned n_rays = 0, n_valid = 0;

    for (;;) {
        // ---- retire finished rays ----
        const bool fin = has && !busy;
        if (__any_sync(FULL, fin)) {
            if (!ANY) {
                int qi = -1;
                if (fin) {
                    int id = -1;
                    float t = tbest;
                    if (best >= 0 && (best_prim < 0 || tbest < prim_t)) {
                        id = sc.n_prims + (best >> 2);
                        qi = best & 3;
                    } else if (best_prim >= 0) {
                        id = best_prim;
                        t = prim_t;
                        const int kind = __ldg(&sc.mats[prims[best_prim].material].kind);
                        qi = kind == PRT_MAT_DIFFUSE ? 0 : (kind == PRT_MAT_DIELECTRIC ? 1 : 2);
                    }
                    st_stream(B.ST + 8 * (size_t) slot + 6, make_float4(t, b1, b2, __int_as_float(id)));
                    if (id >= 0) n_valid++;
                }
#if WF_SLOT_SHADE
                if (fin) B.tag[slot] = (uint8_t) (qi + 1);
#endif
#if WF_QCHUNK
                int *qc = s_qcur[threadIdx.x >> 5];
#pragma unroll
                for (int k = 0; k < WF_QUEUES; k++) {
                    const unsigned m = __ballot_sync(FULL, qi == k);
                    if (!m) continue;
                    const int cnt = __popc(m), rank = __popc(m & lanemask_lt());
                    int next = qc[2 * k], end = qc[2 * k + 1];
                    const int room = end - next;
                    int fresh = 0;
                    if (cnt > room) {      // WF_QCHUNK >= 32 >= cnt: one new chunk always suffices
                        if (lane == 0) fresh = atomicAdd(C + C_MAT + k, WF_QCHUNK);
                        fresh = __shfl_sync(FULL, fresh, 0);
                    }
                    if (qi == k) st_stream(B.q_mat[k] + (rank < room ? next + rank : fresh + (rank - room)), slot);
                    __syncwarp();
                    if (lane == 0) {
                        if (cnt > room) { qc[2 * k] = fresh + (cnt - room); qc[2 * k + 1] = fresh + WF_QCHUNK; }
                        else qc[2 * k] = next + cnt;
                    }
                    __syncwarp();
                }
#else
#pragma unroll
                for (int k = 0; k < WF_QUEUES; k++) {
                    const int q = wf_reserve(C + C_MAT + k, qi == k);
                    if (qi == k) B.q_mat[k][q] = slot;
                }
#endif
            } else if (fin && best < 0 && best_prim < 0) {
                const float4 c = ld_stream(B.SHC + qpos);
                const float w = ld_stream(R.r2 + qpos).w;
                float4 *acc = B.ST + 8 * (size_t) slot + 3;
                float4 r = ld_stream(acc);
                r.x = fmaf(c.x, w, r.x);
                r.y = fmaf(c.y, w, r.y);
                r.z = fmaf(c.z, w, r.z);
                st_stream(acc, r);
            }
            if (fin) has = false;
        }
        // ---- refill idle lanes ----
        if (!dry) {
            unsigned need = __ballot_sync(FULL, !has);
            if (__popc(need) < WF_REFILL_MIN && need != FULL && __any_sync(FULL, busy)) need = 0;
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
#if WF_PREFETCH & 1
                    // the warp consumes this chunk a few rays per iteration: pull its records towards L2 now
                    for (int i = pool_next + lane; i < pool_end; i += 32) {
                        prefetch_l2(R.r0 + i);
                        prefetch_l2(R.r2 + i);
                        prefetch_l2(R.r3 + i);
                        if (sc.n_prims > 0 || ANY) prefetch_l2(R.r1 + i);
                    }
#endif
                }
                const int avail = pool_end - pool_next;
                const int rank = __popc(need & lanemask_lt());
                const bool take = !has && rank < avail;
                if (take) {
                    qpos = perm ? __ldg(perm + pool_next + rank) : (uint32_t) (pool_next + rank);
                    const float4 a = ld_stream(R.r0 + qpos), c2 = ld_stream(R.r2 + qpos), c3 = ld_stream(R.r3 + qpos);
                    r8.o = xyz(a);
                    slot = __float_as_uint(a.w);
                    r8.inv = xyz(c2);
                    const uint32_t oct = (c2.x < 0.0f ? 4u : 0u) | (c2.y < 0.0f ? 2u : 0u) | (c2.z < 0.0f ? 1u : 0u);
                    r8.octinv4 = (7u - oct) * 0x01010101u;
                    rp.Sx = c3.x; rp.Sy = c3.y; rp.Sz = c3.z;
                    kpack = __float_as_int(c3.w);
                    rp.kx = kpack & 3; rp.ky = (kpack >> 2) & 3; rp.kz = (kpack >> 4) & 3;
                    has = true;
                    n_rays++;
                    best = -1;
                    best_prim = -1;
                    tbest = PRT_INF;
                    bool blocked = false;
                    if (sc.n_prims > 0 || ANY) {
                        const float4 bb = ld_stream(R.r1 + qpos);
                        tbest = bb.w;
                        prim_t = tbest;
                        const float3 d = xyz(bb);
                        for (int i = 0; i < sc.n_prims; i++) {
                            const float t = intersect_prim(prims[i], r8.o, d, prim_t);
                            if (t >= 0.0f && (best_prim < 0 || t < prim_t)) {
                                best_prim = i;
                                prim_t = t;
                                if (ANY) blocked = true;
                            }
                        }
                        if (!ANY) tbest = prim_t;
                    }
                    // the scene's oversized triangles (DScene::n_small) are not in the hierarchy: they become this ray's FIRST
                    // triangle group, dealt out with everybody else's leaf triangles in the triangle phase below (a loop at
                    // this point would run with the 2-4 lanes that happen to refill together: measured 81 -> 108 ms)
                    fresh = sc.n_small < sc.n_tris && !blocked;
                    sp = 0;
                    ng = make_uint2(0u, 0x80000000u);
                    busy = !((sc.n_small == 0 && !fresh) || blocked);
                }
                pool_next += min(avail, __popc(need));
                need = __ballot_sync(FULL, !has);
            }
        }
        if (!__any_sync(FULL, has)) break;

        // ---- one wide node per lane and iteration ----
#if WF_TRI_DEFER
        const bool can_step = busy && tg.y == 0u;      // a lane with untested triangles waits for the warp's next triangle round
#else
        const bool can_step = busy;
        tg = make_uint2(0u, 0u);
#endif
        bool stepped = false;
        if (can_step && fresh) {
            stepped = true;
            // bit 31 of the base: the group lives in tri_v (sorted order, behind the n_small triangles of the tree), not in tri_v8
            tg = make_uint2(0x80000000u | (uint32_t) sc.n_small, (1u << (sc.n_tris - sc.n_small)) - 1u);
            fresh = false;
            if (sc.n_small == 0) ng.y = 0u;                  // no tree at all: the traversal ends after this group
        } else if (can_step && ng.y > 0x00ffffffu) {
            stepped = true;
            const uint32_t hits = ng.y, imask8 = ng.y & 0xffu;
            const int bit = 31 - __clz(hits);
            ng.y &= ~(1u << bit);
            if (ng.y > 0x00ffffffu && sp < BVH8_STACK) {
                if (sp < WF_SSTACK) sstack[sp * WF_TRACE_THREADS] = ng;
                else gstack[sp - WF_SSTACK] = ng;
                sp++;
            }
            const uint32_t slot_index = (uint32_t) (bit - 24) ^ (r8.octinv4 & 0xffu);
            const uint32_t rel = __popc(imask8 & ~(0xffffffffu << slot_index));
            uint32_t child_base, tri_base, imask;
#if WF_TOPN
            const uint32_t nidx = ng.x + rel;
            const uint32_t hm = nidx < (uint32_t) WF_TOPN ? bvh8_node<true>(s_top, nidx, r8, tbest, child_base, tri_base, imask)
                                                           : bvh8_node(sc.nodes8, nidx, r8, tbest, child_base, tri_base, imask);
#else
            const uint32_t hm = bvh8_node(sc.nodes8, ng.x + rel, r8, tbest, child_base, tri_base, imask);
#endif
            ng = make_uint2(child_base, (hm & 0xff000000u) | imask);
            tg = make_uint2(tri_base, hm & 0x00ffffffu);
#if WF_PREFETCH & 4
            if (tg.y) {     // the triangle tests below may be dealt to other lanes: start the fetch from the owner now
                const float4 *tv = sc.tri_v8 + 3 * (size_t) (tg.x + (uint32_t) (31 - __clz(tg.y)));
                prefetch_l1(tv);
                prefetch_l1(tv + 2);
            }
#endif
#if WF_PREFETCH & 2
            if (ng.y > 0x00ffffffu) {   // the node this lane visits next: its fetch overlaps the triangle phase
                const int nb = 31 - __clz(ng.y);
                const uint32_t si = (uint32_t) (nb - 24) ^ (r8.octinv4 & 0xffu);
                const float4 *nn = sc.nodes8 + 5 * (size_t) (ng.x + __popc(ng.y & 0xffu & ~(0xffffffffu << si)));
                prefetch_l1(nn);
                prefetch_l1(nn + 4);
            }
#endif
        }
        // ---- the triangles those nodes yielded ----
        unsigned mT = __ballot_sync(FULL, tg.y != 0u);
#if WF_TRI_PEEL
        if (__popc(mT) > WF_COOP_MAX) {
            // most lanes hold triangles (small scenes, coherent rays): every lane tests ONE triangle of its own list per round,
            // for as long as more than WF_COOP_MAX lanes still have one.  The lists differ in length (0 .. 24): running them
            // to exhaustion left 6.7 of 32 lanes active on average (ncu r01, height field), so the long tails are handed to the
            // cooperative branch below, which deals (ray, triangle) pairs out evenly.  The shear rows of the selection-free
            // triangle test are rebuilt here rather than carried through the traversal loop (nine more live registers spill).
#if PRT_TRI_ROWS
            const RayRows rr = ray_rows(rp);
#endif
            do {
                if (tg.y) {
                    const int bit = 31 - __clz(tg.y);
                    tg.y &= ~(1u << bit);
                    const float4 *tv = wf_tri_ptr(sc, tg.x, (uint32_t) bit);
                    const float4 a = ldg4(tv), b = ldg4(tv + 1), c = ldg4(tv + 2);
#if PRT_TRI_ROWS
                    if (intersect_tri_rows(rr, r8.o, xyz(a), xyz(b), xyz(c), tbest, b1, b2)) {
#else
                    if (intersect_tri_wt(rp, r8.o, xyz(a), xyz(b), xyz(c), tbest, b1, b2)) {
#endif
                        best = __float_as_int(b.w);
                        if (ANY) { busy = false; tg.y = 0u; }
                    }
                }
                mT = __ballot_sync(FULL, tg.y != 0u);
            } while (__popc(mT) > WF_COOP_MAX);
        }
        if (mT) {
#else
        if (!WF_TRI_DEFER && __popc(mT) > WF_COOP_MAX) {
            // most lanes hold triangles (small scenes, coherent rays): every lane walks its own list.  The shear rows of
            // the selection-free triangle test (intersect_tri_rows) are rebuilt here, once per list, rather than carried
            // through the traversal loop: nine more live registers spill at 64 (height field 1 875 -> 1 818 Mrays/s), and
            // the cooperative branch below would have to shuffle nine values per pair instead of four
#if PRT_TRI_ROWS
            const RayRows rr = ray_rows(rp);
#endif
            while (tg.y) {
                const int bit = 31 - __clz(tg.y);
                tg.y &= ~(1u << bit);
                const float4 *tv = wf_tri_ptr(sc, tg.x, (uint32_t) bit);
                const float4 a = ldg4(tv), b = ldg4(tv + 1), c = ldg4(tv + 2);
#if PRT_TRI_ROWS
                if (intersect_tri_rows(rr, r8.o, xyz(a), xyz(b), xyz(c), tbest, b1, b2)) {
#else
                if (intersect_tri_wt(rp, r8.o, xyz(a), xyz(b), xyz(c), tbest, b1, b2)) {
#endif
                    best = __float_as_int(b.w);
                    if (ANY) { busy = false; break; }
                }
            }
        } else if (mT) {
#endif
            // A few lanes hold several triangles each and most hold none (the per-lane loop ran with 4.5 of 32 lanes
            // on incoherent rays).  The warp's (ray, triangle) pairs are numbered by a prefix sum and dealt out one
            // per lane: a lane pulls the owning ray through shuffles, tests its triangle, and the closest hit per
            // owner is chosen with two shared-memory atomics (distance first, then the lowest pair).
            const unsigned cnt = __popc(tg.y);
            unsigned incl = cnt;
#pragma unroll
            for (int k = 1; k < 32; k <<= 1) {
                const unsigned v = __shfl_up_sync(FULL, incl, k);
                if (lane >= k) incl += v;
            }
            const unsigned total_all = __shfl_sync(FULL, incl, 31), off = incl - cnt;
#if WF_TRI_DEFER
            // full rounds only while some lane still advanced through the tree this iteration; everything once nobody can
            const bool moving = __any_sync(FULL, stepped);
            const unsigned total = moving ? (total_all >= (unsigned) WF_TRI_DEFER ? (total_all & ~31u) : 0u) : total_all;
#else
            const unsigned total = total_all;
#endif
            for (unsigned base = 0; base < total; base += 32) {
                stm[lane] = 0xffffffffu;
                sw[lane] = 32;
                __syncwarp();
                const unsigned p = base + lane;
                const bool valid = p < total;
                // owner of pair p = the first lane whose inclusive prefix exceeds p: a five-step binary search over the
                // shuffled prefix sums (the first version had every owner write its lane id into up to 24 shared-memory
                // slots, one store per iteration: 10 % of the kernel's instructions, ncu r02a)
                int own = 0;
#pragma unroll
                for (int step = 16; step; step >>= 1) {
                    const unsigned v = __shfl_sync(FULL, incl, (own + step - 1) & 31);
                    if (v <= p) own += step;
                }
                if (!valid) own = lane;
                const unsigned o_off = __shfl_sync(FULL, off, own), o_mask = __shfl_sync(FULL, tg.y, own);
                const unsigned o_base = __shfl_sync(FULL, tg.x, own);
                const float3 ro = mk3(__shfl_sync(FULL, r8.o.x, own), __shfl_sync(FULL, r8.o.y, own), __shfl_sync(FULL, r8.o.z, own));
                RayPre q;
                q.Sx = __shfl_sync(FULL, rp.Sx, own);
                q.Sy = __shfl_sync(FULL, rp.Sy, own);
                q.Sz = __shfl_sync(FULL, rp.Sz, own);
                const int kp = __shfl_sync(FULL, kpack, own);
                q.kx = kp & 3; q.ky = (kp >> 2) & 3; q.kz = (kp >> 4) & 3;
                float ht = __shfl_sync(FULL, tbest, own), hb1 = 0.0f, hb2 = 0.0f;
                int hid = -1;
                bool hit = false;
                if (valid) {
                    unsigned m = o_mask;
                    for (unsigned r = p - o_off; r; r--) m &= m - 1u;
                    const float4 *tv = wf_tri_ptr(sc, o_base, (uint32_t) (__ffs(m) - 1));
                    const float4 a = ldg4(tv), b = ldg4(tv + 1), c = ldg4(tv + 2);
#if PRT_TRI_ROWS
                    // same arithmetic as the per-lane branch and the megakernels (bit-identical hits across back ends);
                    // the rows are rebuilt from the four shuffled words rather than shuffled as nine
                    hit = intersect_tri_rows(ray_rows(q), ro, xyz(a), xyz(b), xyz(c), ht, hb1, hb2);
#else
                    hit = intersect_tri_wt(q, ro, xyz(a), xyz(b), xyz(c), ht, hb1, hb2);
#endif
                    hid = __float_as_int(b.w);
                    ht += 0.0f;
                    if (hit) atomicMin(&stm[own], __float_as_uint(ht));
                }
                __syncwarp();
                if (hit && __float_as_uint(ht) == stm[own]) atomicMin(&sw[own], lane);
                __syncwarp();
                const int wl = sw[lane];
                const int srcl = wl < 32 ? wl : lane;
                const float wt = __shfl_sync(FULL, ht, srcl), wb1 = __shfl_sync(FULL, hb1, srcl), wb2 = __shfl_sync(FULL, hb2, srcl);
                const int wid = __shfl_sync(FULL, hid, srcl);
                if (wl < 32) {
                    tbest = wt;
                    b1 = wb1;
                    b2 = wb2;
                    best = wid;
                    if (ANY) busy = false;
                }
                __syncwarp();
            }
#if WF_TRI_DEFER
            {   // drop the pairs that were tested (the lowest set bits come first in the numbering above)
                const unsigned done = total > off ? min(total - off, cnt) : 0u;
                for (unsigned i = 0; i < done; i++) tg.y &= tg.y - 1u;
                if (!busy) tg.y = 0u;
            }
#endif
        }
#if WF_TRI_DEFER
        if (busy && tg.y == 0u && ng.y <= 0x00ffffffu) {
#else
        if (busy && ng.y <= 0x00ffffffu) {
#endif
            if (sp > 0) {
                --sp;
                ng = sp < WF_SSTACK ? sstack[sp * WF_TRACE_THREADS] : gstack[sp - WF_SSTACK];
            } else busy = false;
        }
    }
#if WF_QCHUNK
    if (!ANY) {     // unused tail of the last reserved chunk of every queue: mark as holes for the shading kernels
        const int *qc = s_qcur[threadIdx.x >> 5];
#pragma unroll
        for (int k = 0; k < WF_QUEUES; k++)
            for (int i = qc[2 * k] + lane; i < qc[2 * k + 1]; i += 32) B.q_mat[k][i] = WF_HOLE;
    }
#endif
    wf_add_stat(P, 2, n_rays);
    if (ANY) wf_add_stat(P, 3, n_rays);
    else wf_add_stat(P, 1, n_valid);
}

// ------------------------------------------------------------------------------------------------------------------
// ray sorting: LSD radix sort of (key, queue position) with 8-bit digits; the number of rays lives in device memory
// (the queue counter), so every kernel is launched for the batch capacity and blocks past the end do nothing
// ------------------------------------------------------------------------------------------------------------------
static constexpr int RQ_THREADS = 256, RQ_ITEMS = 8, RQ_TILE = RQ_THREADS * RQ_ITEMS;

__global__ void __launch_bounds__(RQ_THREADS) k_rq_hist(const uint32_t *__restrict__ keys, const int *__restrict__ n_dev, int shift,
                                                        uint32_t *__restrict__ hist, uint32_t nblocks) {
    __shared__ uint32_t h[256];
    const uint32_t n = (uint32_t) *n_dev;
    const uint32_t base = blockIdx.x * RQ_TILE;
    h[threadIdx.x] = 0;
    __syncthreads();
    if (base < n) {
#pragma unroll
        for (int i = 0; i < RQ_ITEMS; i++) {
            const uint32_t idx = base + i * RQ_THREADS + threadIdx.x;
            if (idx < n) atomicAdd(&h[(keys[idx] >> shift) & 255u], 1u);
        }
        __syncthreads();
    }
    hist[threadIdx.x * nblocks + blockIdx.x] = h[threadIdx.x];
}

// vals_in == nullptr: the value of element i is i (first pass)
__global__ void __launch_bounds__(RQ_THREADS) k_rq_scatter(const uint32_t *__restrict__ keys_in, const uint32_t *__restrict__ vals_in,
                                                           uint32_t *__restrict__ keys_out, uint32_t *__restrict__ vals_out,
                                                           const int *__restrict__ n_dev, int shift, const uint32_t *__restrict__ hist,
                                                           uint32_t nblocks) {
    __shared__ uint32_t wh[RQ_THREADS / 32][256];
    const uint32_t n = (uint32_t) *n_dev;
    if (blockIdx.x * RQ_TILE >= n) return;
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int j = lane; j < 256; j += 32) wh[w][j] = 0;
    __syncwarp();
    const uint32_t base = blockIdx.x * RQ_TILE + w * (RQ_ITEMS * 32);
    uint32_t key[RQ_ITEMS], off[RQ_ITEMS];
#pragma unroll
    for (int i = 0; i < RQ_ITEMS; i++) {
        const uint32_t idx = base + i * 32 + lane;
        const bool ok = idx < n;
        key[i] = ok ? keys_in[idx] : 0xffffffffu;
        const uint32_t digit = ok ? ((key[i] >> shift) & 255u) : 256u;
        const uint32_t peers = __match_any_sync(FULL, digit);
        const uint32_t rank = __popc(peers & ((1u << lane) - 1u));
        const int leader = __ffs(peers) - 1;
        uint32_t pre = 0;
        if (ok && lane == leader) {
            pre = wh[w][digit];
            wh[w][digit] = pre + __popc(peers);
        }
        pre = __shfl_sync(FULL, pre, leader);
        off[i] = pre + rank;
        __syncwarp();
    }
    __syncthreads();
    {
        const uint32_t d = threadIdx.x;
        uint32_t running = hist[d * nblocks + blockIdx.x];
#pragma unroll
        for (int ww = 0; ww < RQ_THREADS / 32; ww++) {
            const uint32_t c = wh[ww][d];
            wh[ww][d] = running;
            running += c;
        }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < RQ_ITEMS; i++) {
        const uint32_t idx = base + i * 32 + lane;
        if (idx < n) {
            const uint32_t pos = wh[w][(key[i] >> shift) & 255u] + off[i];
            keys_out[pos] = key[i];
            vals_out[pos] = vals_in ? vals_in[idx] : idx;
        }
    }
}

struct WfSort {
    uint32_t *keys[2], *vals[2], *hist, *scan;
    uint32_t nblocks;
    int passes;
};

// sorts the first *n_dev entries of `keys`; returns the permutation (sorted rank -> queue position) in *perm
static int wf_sort(const WfSort &S, const uint32_t *keys, const int *n_dev, const uint32_t **perm, cudaStream_t st, int *launches) {
    const uint32_t *kin = keys, *vin = nullptr;
    int cur = 0;
    for (int pass = 0; pass < S.passes; pass++) {
        k_rq_hist<<<S.nblocks, RQ_THREADS, 0, st>>>(kin, n_dev, 8 * pass, S.hist, S.nblocks);
        exclusive_scan_u32(S.hist, 256 * S.nblocks, S.scan, st);
        k_rq_scatter<<<S.nblocks, RQ_THREADS, 0, st>>>(kin, vin, S.keys[cur], S.vals[cur], n_dev, 8 * pass, S.hist, S.nblocks);
        kin = S.keys[cur];
        vin = S.vals[cur];
        cur ^= 1;
        *launches += 5;
    }
    *perm = vin;
    PRT_CUDA(cudaGetLastError());
    return PRT_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// shading, one kernel per material queue
// ------------------------------------------------------------------------------------------------------------------
#ifndef WF_SHADE0_MINB
#define WF_SHADE0_MINB 3                  // the diffuse kernel is compiled without the other materials and fits 80 registers (32 B spill)
#endif
template <int QI>
__global__ void __launch_bounds__(WF_SHADE_THREADS, QI == 0 ? WF_SHADE0_MINB : WF_SHADE_MINB) k_wf_shade(const PtDev P, const WfBuf B, const int bounce) {
    int *C = B.cnt + bounce * WF_CSTRIDE;
    static_assert(C_SH + 1 == WF_CSTRIDE && C_EXT == 0, "shadow count of bounce b must sit right below the extend count of bounce b + 1");
    // Two ways to find this kernel's paths.  The compacted queue (slot ids appended by the trace kernel, ballot +
    // prefix popcount per retire event) costs nothing for sparse materials, but it is filled in RETIRE order, which
    // drifts towards a random permutation of the slots within ~5 bounces: path state is then gathered from scattered
    // DRAM pages and the same kernel runs 2x slower (profiles/r01_summary.md).  When most of the batch is in this
    // queue anyway, walk the SLOTS in order instead and pick this material by its tag: state becomes a sequential
    // stream of 128-byte records and the next bounce's ray queue comes out slot-coherent at warp granularity.
    const int nq = C[C_MAT + QI];
    const uint32_t *q = B.q_mat[QI];
#if WF_SLOT_SHADE
    const int ns = (int) (B.n_layers * B.L);
    const bool by_slot = (long long) nq * 100 > (long long) ns * WF_SLOT_SHADE;
    const int n = by_slot ? ns : nq;
#else
    const bool by_slot = false;
    const int n = nq;
#endif
    const WfRays Rn = B.ext[(bounce + 1) & 1];
    for (int i0 = blockIdx.x * blockDim.x; i0 < n; i0 += gridDim.x * blockDim.x) {
        const int i = i0 + threadIdx.x;
#if WF_SHADE_PREFETCH
        {   // the next iteration's inputs are at known addresses: pull them towards L2 while this one computes.  Measured on
            // B200: nothing gained (cbox shade 22.1 vs 21.6 ms, height field 13.7 vs 13.7) -- off by default
            const long long i_next = (long long) i + (long long) gridDim.x * blockDim.x;
            if (i_next < n) {
                if (by_slot) {
                    const char *rec = reinterpret_cast<const char *>(B.ST + 8 * (size_t) i_next);
                    prefetch_l2(rec); prefetch_l2(rec + 32); prefetch_l2(rec + 64); prefetch_l2(rec + 96);
                    if ((threadIdx.x & 31) == 0) prefetch_l2(B.tag + i_next);
                } else if ((threadIdx.x & 7) == 0) {
                    prefetch_l2(q + i_next);
                }
            }
        }
#endif
        bool live = false;
        uint32_t slot = 0;
        ShadowReq sr;
        sr.want = false;
        PtState st;
        bool mine = false;
        if (i < n) {
            if (by_slot) {
                slot = (uint32_t) i;
                mine = B.tag[i] == (uint8_t) (QI + 1);
            } else {
                slot = q[i];
                mine = slot != WF_HOLE;
            }
        }
        if (mine) {
            wf_load_state(B, slot, st);
            const float4 hv = ld_stream(B.ST + 8 * (size_t) slot + 6);
            const int id = __float_as_int(hv.w);
            Hit h;
            if (id >= P.sc.n_prims) fill_tri_hit(P.sc, id - P.sc.n_prims, hv.x, hv.y, hv.z, h);
            else fill_prim_hit(P.sc.prims[id], id, st.o, st.d, hv.x, h);
#if WF_SHADE_KSEL
            live = pt_shade<QI == 0 ? PRT_MAT_DIFFUSE : (QI == 1 ? PRT_MAT_DIELECTRIC : -1)>(P, st, h, true, sr);
#else
            live = pt_shade(P, st, h, true, sr);
#endif
            wf_store_state(B, slot, st);
#if WF_SLOT_SHADE
            if (!live) B.tag[slot] = 0;      // a live path's tag is rewritten when its next ray retires
#endif
        }
        // one 64-bit atomic per warp for both output queues (two dependent round trips to L2 were 18 % of this kernel's
        // stall samples at its 25 % occupancy)
        const unsigned m_sh = __ballot_sync(FULL, sr.want), m_ex = __ballot_sync(FULL, live);
        unsigned long long base2 = 0ull;
        if (m_sh | m_ex) {
            if ((threadIdx.x & 31) == 0)
                base2 = atomicAdd(reinterpret_cast<unsigned long long *>(C + C_SH),
                                  (unsigned long long) __popc(m_sh) | ((unsigned long long) __popc(m_ex) << 32));
            base2 = __shfl_sync(FULL, base2, 0);
        }
        const int j = (int) (base2 & 0xffffffffull) + __popc(m_sh & lanemask_lt());
        const int e = (int) (base2 >> 32) + __popc(m_ex & lanemask_lt());
        if (sr.want) {
            wf_write_ray(B, B.sh, j, sr.o, sr.d, sr.tmax, slot, sr.w);
            st_stream(B.SHC + j, make_float4(sr.c.x, sr.c.y, sr.c.z, 0.0f));
        }
        if (live) wf_write_ray(B, Rn, e, st.o, st.d, PRT_INF, slot, 0.0f);
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
            const float4 *rec = B.ST + 8 * (size_t) slot;
            const float px = rec[0].w, py = rec[1].w;
            const float4 res = rec[3];
            pt_splat(P.tent, tile, tx0, ty0, px, py, xyz(res));
        }
    }
    __syncthreads();
    pt_flush_tile(P, tile, tx0, ty0);
}

static int wf_grid(prt_context *c, const void *kernel, int threads, int *grid, size_t smem = 0) {
    int per_sm = 0;
    PRT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem));
    if (per_sm < 1) per_sm = 1;
    *grid = c->sm_count * per_sm;
    return PRT_OK;
}

int launch_wavefront(prt_context *c, const PtDev &P, cudaStream_t st) {
    PRT_REQUIRE(P.sc.n_small == 0 || P.sc.n_nodes8 > 0, "render_path (wavefront): the scene has no 8-wide BVH");
    const uint32_t n_tiles = (uint32_t) P.tiles_x * (uint32_t) P.tiles_y;
    const uint64_t L = (uint64_t) n_tiles * 256u;
    uint64_t batch = 1ull << 25;      // 2^25 path slots = 11.7 GB of state + queues (measured: cbox +1 % over 2^24, -8 % at 2^22)
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
    const size_t cnt_bytes = (sizeof(int) * (WF_CSTRIDE * (size_t) (bounces + 1) + 2) + 255) & ~(size_t) 255;
    // ---- run-time knobs of the big-scene path ----
    //   PRT_WF_SORT=<bits per axis>  sort every bounce's rays by (origin cell, direction octant).  PRT_WF_SORT_MODE=1: octant
    //                                major.  PRT_WF_SORT_WHAT: bit 0 extend rays, bit 1 shadow rays (default 3).  OFF by
    //                                default: measured on the 10 M-triangle height field (profiles/r02_summary.md) the sorted
    //                                closest-hit launches are 2 % faster, the shadow launches 8 %, and the sorts themselves
    //                                cost 8 % of the step -- a diffuse bounce's rays share an origin cell, not a destination.
    //   PRT_L2_PERSIST_MB=<MB>       access-policy window over the 8-wide BVH nodes (persisting L2 lines)
    auto env_int = [](const char *name, int dflt) { const char *e = getenv(name); return e && *e ? atoi(e) : dflt; };
    int sort_bits = env_int("PRT_WF_SORT", 0);
    if (sort_bits > 9) sort_bits = 9;                      // 3 x 9 cell bits + 3 octant bits = 30-bit keys
    if (sort_bits < 0 || P.sc.n_tris == 0) sort_bits = 0;
    const int sort_mode = env_int("PRT_WF_SORT_MODE", 0), sort_what = sort_bits ? env_int("PRT_WF_SORT_WHAT", 3) : 0;
    const size_t per_slot = 16 * (8 + 8 + 4 + 1) + 4 * WF_QUEUES + (sort_bits ? 4 * 6 : 0);
    WfSort S;
    S.nblocks = (uint32_t) ((cap + RQ_TILE - 1) / RQ_TILE);
    S.passes = (3 * sort_bits + 3 + 7) / 8;
    const size_t hist_words = 256 * (size_t) S.nblocks, scan_words = hist_words / 2048 + 8192;
    const size_t sort_bytes = sort_bits ? 4 * (hist_words + scan_words) + 256 : 0;
    // every trace warp may leave one partly used chunk per queue behind (holes): room for them on top of `cap` entries
    const size_t q_slack = (size_t) c->sm_count * 64 * (WF_QCHUNK ? WF_QCHUNK : 1);
    const size_t need = (size_t) cap * (per_slot + 1) + 4 * WF_QUEUES * q_slack + cnt_bytes + sort_bytes;
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
        auto take4 = [&]() { return reinterpret_cast<float4 *>(take(16 * cap)); };
        auto take1 = [&]() { return reinterpret_cast<uint32_t *>(take(4 * cap)); };
        B.cnt = reinterpret_cast<int *>(take(cnt_bytes)) + 1;      // odd int offset: see the counter layout above
        B.ST = reinterpret_cast<float4 *>(take(128 * cap));
        for (int k = 0; k < 2; k++) { B.ext[k].r0 = take4(); B.ext[k].r1 = take4(); B.ext[k].r2 = take4(); B.ext[k].r3 = take4(); B.ext[k].key = nullptr; }
        B.sh.r0 = take4(); B.sh.r1 = take4(); B.sh.r2 = take4(); B.sh.r3 = take4(); B.sh.key = nullptr;
        B.SHC = take4();
        for (int k = 0; k < WF_QUEUES; k++) B.q_mat[k] = reinterpret_cast<uint32_t *>(take(4 * (cap + q_slack)));
        B.tag = reinterpret_cast<uint8_t *>(take(cap));
        if (sort_bits) {
            p = reinterpret_cast<char *>(((uintptr_t) p + 255) & ~(uintptr_t) 255);
            // one key array serves both extend-ray buffers: the keys shading writes for bounce b + 1 are consumed by the
            // sort at the start of bounce b + 1, before shading writes the next set
            uint32_t *kext = take1();
            if (sort_what & 1) B.ext[0].key = B.ext[1].key = kext;
            uint32_t *ksh = take1();
            if (sort_what & 2) B.sh.key = ksh;
            S.keys[0] = take1(); S.keys[1] = take1(); S.vals[0] = take1(); S.vals[1] = take1();
            S.hist = reinterpret_cast<uint32_t *>(take(4 * hist_words));
            S.scan = reinterpret_cast<uint32_t *>(take(4 * scan_words));
        }
    }
    B.sort_bits = sort_bits;
    B.sort_mode = sort_mode;
    {
        const float cells = (float) (1 << (sort_bits ? sort_bits : 1));
        const float3 ext = make_float3(P.box_hi.x - P.box_lo.x, P.box_hi.y - P.box_lo.y, P.box_hi.z - P.box_lo.z);
        B.key_lo = P.box_lo;
        B.key_scale = make_float3(ext.x > 0.0f ? cells / ext.x : 0.0f, ext.y > 0.0f ? cells / ext.y : 0.0f, ext.z > 0.0f ? cells / ext.z : 0.0f);
    }
    const int l2_mb = env_int("PRT_L2_PERSIST_MB", 0);
    bool l2_window = false;
    if (l2_mb > 0 && P.sc.n_nodes8 > 0) {
        size_t want = (size_t) l2_mb << 20;
        if (want > (size_t) c->prop.persistingL2CacheMaxSize) want = (size_t) c->prop.persistingL2CacheMaxSize;
        if (want > 0 && cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want) == cudaSuccess) {
            cudaStreamAttrValue av;
            memset(&av, 0, sizeof av);
            size_t bytes = (size_t) P.sc.n_nodes8 * 80;
            if (bytes > (size_t) c->prop.accessPolicyMaxWindowSize) bytes = (size_t) c->prop.accessPolicyMaxWindowSize;
            av.accessPolicyWindow.base_ptr = const_cast<float4 *>(P.sc.nodes8);
            av.accessPolicyWindow.num_bytes = bytes;
            av.accessPolicyWindow.hitRatio = bytes > want ? (float) want / (float) bytes : 1.0f;
            av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
            av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
            l2_window = cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &av) == cudaSuccess;
        }
        cudaGetLastError();
    }
    B.cap = (uint32_t) cap;
    B.L = (uint32_t) L;
    int g_gen = 1, g_ext = 1, g_sh = 1, g_shade[WF_QUEUES] = { 1, 1, 1 };
    int rc;
    if ((rc = wf_grid(c, (const void *) k_wf_generate, WF_SHADE_THREADS, &g_gen))) return rc;
    const size_t trace_smem = wf_trace_smem(P.sc);
    if ((rc = wf_grid(c, (const void *) k_wf_trace<false>, WF_TRACE_THREADS, &g_ext, trace_smem))) return rc;
    if ((rc = wf_grid(c, (const void *) k_wf_trace<true>, WF_TRACE_THREADS, &g_sh, trace_smem))) return rc;
    if ((rc = wf_grid(c, (const void *) k_wf_shade<0>, WF_SHADE_THREADS, &g_shade[0]))) return rc;
    if ((rc = wf_grid(c, (const void *) k_wf_shade<1>, WF_SHADE_THREADS, &g_shade[1]))) return rc;
    if ((rc = wf_grid(c, (const void *) k_wf_shade<2>, WF_SHADE_THREADS, &g_shade[2]))) return rc;
    int launches = 0;
    for (uint64_t j0 = 0; j0 < P.n_s; j0 += layers) {
        B.j0 = (uint32_t) j0;
        B.n_layers = (uint32_t) (P.n_s - j0 < layers ? P.n_s - j0 : layers);
        PRT_CUDA(cudaMemsetAsync(B.cnt - 1, 0, cnt_bytes, st));
#if WF_SLOT_SHADE
        PRT_CUDA(cudaMemsetAsync(B.tag, 0, (size_t) B.n_layers * L, st));
#endif
        {
            ProfScope ps(c, PRT_KC_GENERATE, st);
            k_wf_generate<<<g_gen, WF_SHADE_THREADS, 0, st>>>(P, B);
            launches++;
        }
        for (int b = 0; b < bounces; b++) {
            const uint32_t *perm = nullptr;
            if ((sort_what & 1) && b > 0) {        // camera rays (b == 0) come out of k_wf_generate tile by tile: coherent as they are
                ProfScope ps(c, PRT_KC_OTHER, st);
                const int before = launches;
                if ((rc = wf_sort(S, B.ext[b & 1].key, B.cnt + b * WF_CSTRIDE + C_EXT, &perm, st, &launches))) return rc;
                ps.kernels = launches - before;
            }
            {
                ProfScope ps(c, PRT_KC_TRACE_CLOSEST, st);
                k_wf_trace<false><<<g_ext, WF_TRACE_THREADS, trace_smem, st>>>(P, B, b, perm);
                launches++;
            }
            {
                ProfScope ps(c, PRT_KC_SHADE, st);
                const int before = launches;
                if (P.kind_mask & (1u << PRT_MAT_DIFFUSE)) { k_wf_shade<0><<<g_shade[0], WF_SHADE_THREADS, 0, st>>>(P, B, b); launches++; }
                if (P.kind_mask & (1u << PRT_MAT_DIELECTRIC)) { k_wf_shade<1><<<g_shade[1], WF_SHADE_THREADS, 0, st>>>(P, B, b); launches++; }
                if (P.kind_mask & ~((1u << PRT_MAT_DIFFUSE) | (1u << PRT_MAT_DIELECTRIC))) {
                    k_wf_shade<2><<<g_shade[2], WF_SHADE_THREADS, 0, st>>>(P, B, b);
                    launches++;
                }
                ps.kernels = launches - before;
            }
            if (b + 1 < P.max_depth && (P.kind_mask & (1u << PRT_MAT_DIFFUSE)) && P.sc.n_emitters > 0) {
                perm = nullptr;
                if (sort_what & 2) {
                    ProfScope ps(c, PRT_KC_OTHER, st);
                    const int before = launches;
                    if ((rc = wf_sort(S, B.sh.key, B.cnt + b * WF_CSTRIDE + C_SH, &perm, st, &launches))) return rc;
                    ps.kernels = launches - before;
                }
                ProfScope ps(c, PRT_KC_TRACE_SHADOW, st);
                k_wf_trace<true><<<g_sh, WF_TRACE_THREADS, trace_smem, st>>>(P, B, b, perm);
                launches++;
            }
        }
        {
            ProfScope ps(c, PRT_KC_FILM, st);
            k_wf_film<<<n_tiles, 256, 0, st>>>(P, B);
            launches++;
        }
        PRT_CUDA(cudaGetLastError());
        if (getenv("PRT_WF_DEBUG")) {     // per-bounce queue lengths of this batch (profiling aid; synchronises)
            std::vector<int> h(WF_CSTRIDE * (size_t) (bounces + 1));
            PRT_CUDA(cudaMemcpyAsync(h.data(), B.cnt, sizeof(int) * h.size(), cudaMemcpyDeviceToHost, st));
            PRT_CUDA(cudaStreamSynchronize(st));
            for (int b = 0; b < bounces; b++)
                fprintf(stderr, "[prt wf] batch %u bounce %d: extend rays %d, shadow rays %d, shading queues %d / %d / %d\n", B.j0, b,
                        h[b * WF_CSTRIDE + C_EXT], h[b * WF_CSTRIDE + C_SH], h[b * WF_CSTRIDE + C_MAT], h[b * WF_CSTRIDE + C_MAT + 1],
                        h[b * WF_CSTRIDE + C_MAT + 2]);
        }
    }
    if (l2_window) {
        cudaStreamAttrValue av;
        memset(&av, 0, sizeof av);
        cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &av);      // num_bytes = 0 disables the window
        cudaCtxResetPersistingL2Cache();
    }
    c->last_launches = launches;
    return PRT_OK;
}

}  // namespace prt
