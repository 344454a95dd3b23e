// prt_acquire.cu -- the acquisition hot path as one persistent sm_100a kernel.
//
// Replaces UltraIntegrator.simulate_acquisition (/root/reference/CustomIntegrator.py:60-232, one width-1
// Dr.Jit while_loop per ray) and simulate_acquisition_parallel (:235-405, GIL-bound thread pool).
// One CUDA thread owns one path at a time and keeps its whole state in registers (ray, amp, atten, tof,
// path length, PCG32 stream); when a path ends the lane immediately regenerates the next path of its
// strided list, so warps stay full until the tail.  Lanes of a warp work on consecutive (angle, element)
// pairs: nearly parallel primary rays (coherent traversal) whose deposits land in different
// channel_buf rows (no same-address atomic storms).  Analytic primitives are staged in shared memory;
// deposits are fire-and-forget red.global.add.f32 into the L2-resident channel buffer.
// Canonical path semantics: SURVEY.md Appendix F.  Line tags CI:n / CB:n cite the reference files.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "prt_hit.cuh"
#include "prt_internal.h"

namespace prt {

static constexpr int ACQ_THREADS = 256;
static constexpr int MAX_SMEM_PRIMS = 64;

struct AcqDev {
    DScene sc;
    float4 T0, T1, T2;  // sensor to_world rows
    float3 nT;          // normalize(T * (0,0,1))
    float c, fs, pitch, two_pi_f, att_k, alpha_m, alpha_c, cos_c, cos_m, max_len, n_rays, inv_spp;
    int n_a, n_e, Tn, max_depth;
    unsigned qf;
    const float2 *sincos;  // [n_a] (sin theta, cos theta)
    uint64_t seed;
    uint32_t spp_total, s_offset, s_stride;
    uint64_t n_s, total;   // samples per (a,e) for this call, total paths of this call
    int wae;               // 1: a warp's 32 lanes take 32 samples of ONE (angle, element) (identical primary rays); 0: 32 elements
    int a_first, a_count;  // angle range of this LAUNCH (prt_acquire pipelines one launch per angle with its D2H slice)
    unsigned long long var_mask;   // prt_acquire_variants: materials (bit = id) whose parameter is overridden in this launch
    int var_index;
    float var_value;
    float *buf, *tx;
    unsigned long long *stats;  // {paths, segments, rays, deposits, misses}
};

struct PathState {
    float3 o, d;
    float amp, atten, tof, geo, t0;
    int a, depth;
    Pcg32 rng;
};

__device__ __forceinline__ float elem_x(const AcqDev &P, int e) { return P.pitch * ((float) e - (float) (P.n_e - 1) * 0.5f); }  // CI:84

__device__ __forceinline__ void init_path(const AcqDev &P, uint32_t ae, uint32_t s, PathState &ps) {
    // n_a * n_e fits 32 bits (fill_params): 32-bit division, not the 64-bit call
    const uint32_t qa = ae / (uint32_t) P.n_e;
    int a = (int) qa, e = (int) (ae - qa * (uint32_t) P.n_e);
    float2 sc = __ldg(P.sincos + a);
    float xe = elem_x(P, e);
    ps.t0 = (xe * sc.x) / P.c;                                                  // CI:87
    ps.o = xpoint(P.T0, P.T1, P.T2, mk3(xe, 0.0f, 0.0f));                       // CI:97,103
    ps.d = normalize(xvec(P.T0, P.T1, P.T2, mk3(sc.x, 0.0f, sc.y)));            // CI:98,104
    ps.amp = 1.0f; ps.atten = 1.0f; ps.tof = 0.0f; ps.geo = 0.0f;               // CI:110-114
    ps.a = a; ps.depth = 0;
    ps.rng = path_rng(P.seed, (uint64_t) ae * (uint64_t) P.spp_total + (uint64_t) s);      // RNG contract, SURVEY.md 8(d)
}

// ------------------------------------------------------------------------------------------------------------------
// Echo cache: software combining of deposits in shared memory.
//
// The reference's primary ray depends on (angle, element) only (CustomIntegrator.py:270-273), so all 209 716 samples of a pair
// hit the same point, and its first-segment echo can only land in 64 bins (one per receive element).  One launch therefore
// fires 13.4 M `red.global.add.f32` at a few thousand addresses -- or at a few HUNDRED, for the steered angles whose
// transmit delays line the elements' echoes up in the same time bin: the +-15 degree launches of the headline workload took
// 2.05 ms against 1.06 ms at 0 degrees for the same number of rays (profiles/r02_launches_bench.csv), the L2 serialising
// same-address atomics.  Each CTA keeps a direct-mapped table of (bin, partial sum) in shared memory: a deposit whose bin
// owns its slot is a shared-memory atomic; a bin that finds its slot taken by another goes to global memory as before; the
// table is flushed with one global atomic per occupied slot when the CTA retires.  Sums are re-associated (atomics never
// had an order), nothing else changes.
// ------------------------------------------------------------------------------------------------------------------
#ifndef PRT_ACQ_CACHE
#define PRT_ACQ_CACHE 0                   // slots per CTA (power of two, e.g. 2048; 0 = deposits go straight to global memory -- the default: measured 4 % SLOWER with the table, profiles/r02_summary.md)
#endif
// the mesh kernel's path stash already takes 35 KB of the 48 KB of static shared memory: it gets a quarter-size table
#define ECHO_SLOTS(tris) ((tris) ? PRT_ACQ_CACHE / 4 : PRT_ACQ_CACHE)
#define ECHO_BITS(tris) (__builtin_ctz(PRT_ACQ_CACHE) - ((tris) ? 2 : 0))
struct EchoCache {
    unsigned *key;    // [1 << bits] bin index, 0xffffffff = free
    float *val;
    int bits;
};

__device__ __forceinline__ void echo_add(const AcqDev &P, const EchoCache &ec, size_t flat, float v) {
#if PRT_ACQ_CACHE
    if (flat < 0xffffffffull) {
        const unsigned a = (unsigned) flat;
        const unsigned h = (a * 2654435761u) >> (32 - ec.bits);
        unsigned k = ec.key[h];
        if (k == 0xffffffffu) k = atomicCAS(&ec.key[h], 0xffffffffu, a), k = k == 0xffffffffu ? a : k;
        if (k == a) { atomicAdd(&ec.val[h], v); return; }
    }
#endif
    atomicAdd(P.buf + flat, v);
}

__device__ __forceinline__ void echo_cache_init(const EchoCache &ec) {
#if PRT_ACQ_CACHE
    for (int i = threadIdx.x; i < (1 << ec.bits); i += blockDim.x) { ec.key[i] = 0xffffffffu; ec.val[i] = 0.0f; }
    __syncthreads();
#endif
}

__device__ __forceinline__ void echo_cache_flush(const AcqDev &P, const EchoCache &ec) {
#if PRT_ACQ_CACHE
    __syncthreads();
    for (int i = threadIdx.x; i < (1 << ec.bits); i += blockDim.x) {
        const unsigned k = ec.key[i];
        const float v = ec.val[i];
        if (k != 0xffffffffu && v != 0.0f) atomicAdd(P.buf + k, v);
    }
#endif
}

struct Counters {
    unsigned paths, segments, rays, deposits, misses;
};

// executes ONE segment of the path; returns false when the path terminates.  rec != nullptr records decisions.
template <bool TRIS>
__device__ __forceinline__ bool segment(const AcqDev &P, const DPrim *prims, PathState &ps, Counters &cn, prt_seg_record *rec,
                                        const EchoCache &ec) {
    Hit h;
    cn.rays++;
    if (!closest_hit<TRIS>(P.sc, prims, ps.o, ps.d, PRT_INF, h)) { cn.misses++; return false; }   // CI:146-147 / 309-312
    cn.segments++;
    const float dist = h.t;
    ps.geo += dist;                                                              // CI:209 / 315
    const float tof_here = ps.tof + dist / P.c;                                  // CI:165 / 316
    if (!(P.qf & PRT_QF_TOF_LAST_SEGMENT)) ps.tof = tof_here;
    const float u_recv = ps.rng.next_f32();                                      // CI:153 / 319
    const float s1 = ps.rng.next_f32();                                          // CI:173 / 337
    const float s2 = ps.rng.next_f32();                                          // CI:174 / 337
    const float u_rr = ps.rng.next_f32();                                        // CI:219 / 365
    int recv = min((int) floorf(u_recv * (float) P.n_e), P.n_e - 1);             // CI:154
    const float3 tgt = xpoint(P.T0, P.T1, P.T2, mk3(elem_x(P, recv), 0.0f, 0.0f));   // CI:156-157
    const float3 to_t = tgt - h.p;
    const float dist_recv = sqrtf(dot(to_t, to_t));                              // CI:166 / 329
    const float3 sec = mk3(to_t.x / dist_recv, to_t.y / dist_recv, to_t.z / dist_recv);   // CI:158 / 322
    const float3 so = spawn_origin(h.p, h.ng, sec);
    float vis_tmax = PRT_INF;                                                    // Q1: maxt = inf
    if (P.qf & PRT_QF_CONNECT_TO_TARGET) {
        float3 q = tgt - so;
        vis_tmax = sqrtf(dot(q, q)) * (1.0f - 1e-4f);
    }
    cn.rays++;
    const bool visible = !occluded<TRIS>(P.sc, prims, so, sec, vis_tmax);              // CI:159-160 / 324-325
    ps.atten *= expf((P.att_k * dist) / 8.686f);                                 // CI:162-163 / 328
    const float Ttot = (ps.t0 + tof_here) + dist_recv / P.c;                     // CI:167 / 329
    const float phase = P.two_pi_f * Ttot;                                       // CI:168 / 330
    const float3 md = -ps.d;
    const float3 wi = mk3(dot(md, h.fs), dot(md, h.ft), dot(md, h.ns));          // si.wi
    const DMaterial &mat = P.sc.mats[h.material];
    float mZ = __ldg(&mat.p[0]), mA = __ldg(&mat.p[1]);                          // impedance, roughness (CB:12-18)
    if ((P.var_mask >> (h.material & 63)) & 1ull) {                                             // finite-difference variant (USMain.py:264)
        if (P.var_index == 0) mZ = P.var_value;
        else mA = P.var_value;
    }
    float3 dir; float pdf, a_resp; bool reflect;
    ultra_bsdf_sample(wi, h.ng, h.ns, mZ, mA, s1, s2, dir, pdf, a_resp, reflect);   // CI:175 / 338
    const float cos_theta = dot(h.ns, md);                                       // CI:176 / 340
    ps.amp *= a_resp * cos_theta * fmaxf(pdf, 1e-6f);                            // CI:177 / 341
    const float kf = rintf(Ttot * P.fs);                                         // CI:191 / 351-352 (half-even)
    int k = (int) kf;
    bool in_range = kf >= 0.0f && kf < (float) P.Tn;
    if (P.qf & PRT_QF_CLAMP_TIDX) {                                              // CI:192
        k = !(kf >= 0.0f) ? 0 : (kf > (float) (P.Tn - 1) ? P.Tn - 1 : k);
        in_range = true;
    }
    // The echo's value only matters if it is deposited (CI:197-203 scatter-add under `visible & active`; :353-354):
    // directivity, sin(phase) and the product are skipped for blocked or out-of-range connections (in the Box scenes
    // whole warps are blocked together).
    float press = 0.0f;
    const bool deposit = visible && in_range;
    if (deposit || rec) {
        // CI:124-133: alpha = |acos(dot)|; w_i = 1 (alpha <= alpha_m), linear ramp to 0 at alpha_c, else 0.  acos is
        // monotone, so the two plateaus are decided on the cosine and acosf only runs on the ramp (rare: the aperture
        // subtends a few degrees; the ramp is continuous at both ends, so an ulp-level tie is immaterial)
        const float w_i = directivity_wi(P.nT, sec, P.cos_m, P.cos_c, P.alpha_m, P.alpha_c);
        const float w_o = dot(ps.d, h.ns) / P.n_rays;                            // CI:118,184
        press = ps.atten * ps.amp * (w_i * w_o) * sinf(phase);                   // CI:187 / 348
    }
    if (deposit) {
        if (P.buf) echo_add(P, ec, ((size_t) ps.a * P.n_e + recv) * (size_t) P.Tn + (size_t) k, press * P.inv_spp);   // CI:197-203 / 354
        cn.deposits++;
    }
    ps.d = normalize(dir);                                                       // CI:205-206 / 358-359 (Q9)
    ps.o = spawn_origin(h.p, h.ng, ps.d);
    ps.depth++;                                                                  // CI:210 / 361
    const float prod = ps.atten * ps.amp;
    const float rr = (P.qf & PRT_QF_RR_NO_ABS) ? fminf(prod, 1.0f) : fminf(fabsf(prod), 1.0f);   // CI:220 / 364
    const bool survive = u_rr < rr;                                              // CI:221
    ps.atten = survive ? ps.atten / rr : 0.0f;                                   // CI:224
    if (rec) {
        prt_seg_record &r = rec[ps.depth - 1];
        r.valid = 1; r.prim = h.prim; r.shape = h.shape; r.recv = recv; r.visible = visible; r.reflect = reflect;
        r.k = k; r.survive = survive; r.t = dist; r.total_time = Ttot; r.press = press; r.amp = ps.amp; r.atten = ps.atten;
        r.dir[0] = ps.d.x; r.dir[1] = ps.d.y; r.dir[2] = ps.d.z;
    }
    if (P.qf & PRT_QF_SINGLE_BOUNCE) return false;
    if (!survive || !(dot(ps.d, P.nT) >= P.cos_c)) return false;                 // CI:212-223 (Q11)
    return ps.depth < P.max_depth && ps.geo < P.max_len;                         // CI:141 / 307
}

__device__ __forceinline__ const DPrim *stage_prims(const DScene &sc, DPrim *smem) {
    if (sc.n_prims > MAX_SMEM_PRIMS) return sc.prims;
    const float4 *src = reinterpret_cast<const float4 *>(sc.prims);
    float4 *dst = reinterpret_cast<float4 *>(smem);
    for (int i = threadIdx.x; i < sc.n_prims * (int) (sizeof(DPrim) / 16); i += blockDim.x) dst[i] = src[i];
    __syncthreads();
    return smem;
}

// occupancy targets: the analytic-only specialisation has no traversal stack and fits 4 CTAs/SM (64 regs);
// the BVH one is held at 3 CTAs/SM (80 regs)
// GPRIMS = true: more than MAX_SMEM_PRIMS analytic primitives, read from global memory.  Making that a template
// parameter (instead of a run-time pointer choice) lets the compiler see that `prims` points into shared memory in the
// common case and emit LDS instead of generic loads.
#ifndef PRT_ACQ_WAE
#define PRT_ACQ_WAE 1                     // warp-per-(angle, element) lane map: 0 off, 1 mesh scenes, 2 all scenes
#endif
#ifndef PRT_ACQ_DEFER
#define PRT_ACQ_DEFER 1                   // mesh scenes: park continuing paths and run their segments in separate warp iterations
#endif
#if PRT_ACQ_DEFER
static constexpr int ACQ_STASH_CAP = 64;  // < 32 parked before an iteration + at most 32 new ones
#endif
#ifndef PRT_ACQ_LDS
#define PRT_ACQ_LDS 1
#endif
template <bool TRIS, bool GPRIMS>
__global__ void __launch_bounds__(ACQ_THREADS, TRIS ? 3 : 4) k_acquire(const AcqDev P) {
    __shared__ DPrim sprims[MAX_SMEM_PRIMS];
    constexpr bool TRIS_K = TRIS;
    (void) TRIS_K;
    EchoCache ec;
#if PRT_ACQ_CACHE
    __shared__ unsigned s_ec_key[ECHO_SLOTS(TRIS_K)];
    __shared__ float s_ec_val[ECHO_SLOTS(TRIS_K)];
    ec.key = s_ec_key; ec.val = s_ec_val; ec.bits = ECHO_BITS(TRIS_K);
#else
    ec.key = nullptr; ec.val = nullptr; ec.bits = 0;
#endif
    echo_cache_init(ec);
#if PRT_ACQ_LDS
    const DPrim *prims = GPRIMS ? P.sc.prims : sprims;
    if (!GPRIMS) {
        const float4 *src = reinterpret_cast<const float4 *>(P.sc.prims);
        float4 *dst = reinterpret_cast<float4 *>(sprims);
        for (int i = threadIdx.x; i < P.sc.n_prims * (int) (sizeof(DPrim) / 16); i += blockDim.x) dst[i] = src[i];
        __syncthreads();
    }
#else
    const DPrim *prims = stage_prims(P.sc, sprims);
#endif
    // sample-major launch order: consecutive lanes -> consecutive (angle, element) of this launch's angle range,
    // same sample.  (ae, si) advance incrementally -- no 64-bit divisions in the loop.
    const uint32_t n_ae = (uint32_t) P.a_count * (uint32_t) P.n_e;
    const uint32_t ae0 = (uint32_t) P.a_first * (uint32_t) P.n_e;
    const uint64_t stride = (uint64_t) gridDim.x * blockDim.x;
    const uint64_t j0 = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    // Two lane -> path maps.  Default: consecutive lanes take consecutive (angle, element) pairs at the same sample (their
    // deposits fall into different rows).  P.wae (mesh scenes with >= 64 samples per pair): the 32 lanes of a warp take 32
    // consecutive SAMPLES of one (angle, element) -- the reference's primary ray depends on the element only
    // (CustomIntegrator.py:270-273), so the whole warp traverses the BVH with one and the same ray until the first hit.
    const uint32_t lane_id = threadIdx.x & 31;
    const uint64_t u0 = P.wae ? (j0 >> 5) : j0, du = P.wae ? (stride >> 5) : stride;
    const uint32_t d_ae = (uint32_t) (du % n_ae), d_si = (uint32_t) (du / n_ae);
    uint32_t ae = (uint32_t) (u0 % n_ae);
    uint64_t sq = u0 / n_ae;                       // sample index, or block of 32 samples
    uint64_t si = P.wae ? sq * 32 + lane_id : sq;
    // transmit-delay table, CI:87-94 / 254-257
    if (P.tx && j0 < n_ae) {
        uint32_t g = ae0 + (uint32_t) j0;
        float2 sc = __ldg(P.sincos + (int) (g / (uint32_t) P.n_e));
        P.tx[g] = (elem_x(P, (int) (g % (uint32_t) P.n_e)) * sc.x) / P.c;
    }
    Counters cn = { 0, 0, 0, 0, 0 };
    PathState ps;
#if PRT_ACQ_DEFER
    if (TRIS || PRT_ACQ_DEFER == 2) {      // 2: analytic scenes as well (A/B knob)
        // Mesh scenes: primary segments (32 parallel rays of neighbouring elements: near-identical traversals) and the
        // segments of continuing paths (scattered directions) are not mixed in one warp iteration.  A path that survives
        // its segment is parked in a per-warp shared-memory stash (ballot-compacted, 17 words, odd stride: conflict
        // free); the warp runs a secondary iteration whenever 32 are parked (or its primary work is exhausted).  In the
        // interleaved loop a warp ran 19.4 of 32 lanes (ncu r01): every iteration waited for its few incoherent rays.
        __shared__ float s_stash[ACQ_THREADS / 32][ACQ_STASH_CAP][17];
        float(*stash)[17] = s_stash[threadIdx.x >> 5];
        const int lane = threadIdx.x & 31;
        int n_st = 0;                                  // warp-uniform
        for (;;) {
            const bool prim_left = __any_sync(0xffffffffu, si < P.n_s);
            const bool sec = n_st >= 32 || (!prim_left && n_st > 0);
            if (!sec && !prim_left) break;
            bool have = false;
            if (sec) {
                const int take = min(n_st, 32);
                if (lane < take) {
                    const float *e = stash[n_st - take + lane];
                    ps.o = mk3(e[0], e[1], e[2]); ps.d = mk3(e[3], e[4], e[5]);
                    ps.amp = e[6]; ps.atten = e[7]; ps.tof = e[8]; ps.geo = e[9]; ps.t0 = e[10];
                    ps.a = __float_as_int(e[11]); ps.depth = __float_as_int(e[12]);
                    ps.rng.state = ((uint64_t) __float_as_uint(e[14]) << 32) | __float_as_uint(e[13]);
                    ps.rng.inc = ((uint64_t) __float_as_uint(e[16]) << 32) | __float_as_uint(e[15]);
                    have = true;
                }
                n_st -= take;
                __syncwarp();
            } else if (si < P.n_s) {
                init_path(P, ae0 + ae, P.s_offset + (uint32_t) si * P.s_stride, ps);
                ae += d_ae;
                sq += d_si;
                if (ae >= n_ae) { ae -= n_ae; sq++; }
                si = P.wae ? sq * 32 + lane_id : sq;
                cn.paths++;
                have = P.max_depth > 0;
            }
            const bool cont = have && segment<TRIS>(P, prims, ps, cn, nullptr, ec);
            const unsigned m = __ballot_sync(0xffffffffu, cont);
            if (cont) {
                float *e = stash[n_st + __popc(m & ((1u << lane) - 1u))];
                e[0] = ps.o.x; e[1] = ps.o.y; e[2] = ps.o.z; e[3] = ps.d.x; e[4] = ps.d.y; e[5] = ps.d.z;
                e[6] = ps.amp; e[7] = ps.atten; e[8] = ps.tof; e[9] = ps.geo; e[10] = ps.t0;
                e[11] = __int_as_float(ps.a); e[12] = __int_as_float(ps.depth);
                e[13] = __uint_as_float((uint32_t) ps.rng.state); e[14] = __uint_as_float((uint32_t) (ps.rng.state >> 32));
                e[15] = __uint_as_float((uint32_t) ps.rng.inc); e[16] = __uint_as_float((uint32_t) (ps.rng.inc >> 32));
            }
            n_st += __popc(m);
            __syncwarp();
        }
    } else
#endif
    {
        // Every iteration ends in a warp-wide vote: it is the loop's exit test AND the point where the warp reconverges.
        // The first version let each lane `break` / `continue` on its own; once a few paths of a warp had ended early the
        // lanes that regenerate and the lanes that continue never met again, and the whole segment ran twice per iteration
        // with half the lanes each (ncu r02n: 15.8 of 32 lanes and 2x the instructions on the +-15 degree launches of the
        // headline workload, where 9 % of the paths end after one segment; 31.7 lanes at 0 degrees, where none does).
        bool live = false;
        for (;;) {
            if (!live && si < P.n_s) {
                init_path(P, ae0 + ae, P.s_offset + (uint32_t) si * P.s_stride, ps);
                ae += d_ae;
                sq += d_si;
                if (ae >= n_ae) { ae -= n_ae; sq++; }
                si = P.wae ? sq * 32 + lane_id : sq;
                cn.paths++;
                live = P.max_depth > 0;
            }
            if (!__any_sync(0xffffffffu, live)) {
                if (!__any_sync(0xffffffffu, si < P.n_s)) break;
                continue;
            }
            if (live) live = segment<TRIS>(P, prims, ps, cn, nullptr, ec);
        }
    }
    if (P.buf) echo_cache_flush(P, ec);
    if (P.stats) {
        unsigned v[5] = { cn.paths, cn.segments, cn.rays, cn.deposits, cn.misses };
#pragma unroll
        for (int q = 0; q < 5; q++) {
            unsigned x = v[q];
#pragma unroll
            for (int o = 16; o; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
            if ((threadIdx.x & 31) == 0 && x) atomicAdd(P.stats + q, (unsigned long long) x);
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Mesh scenes, per-lane state machine (PRT_ACQ_SM).
//
// k_acquire<true> runs a path segment as "closest-hit traversal, connection-ray traversal, shading" one after the other
// inside one loop iteration.  Each of the two traversals is a while loop of its own: its lanes leave at different times and
// the warp waits for the slowest (ncu r01 on the ring: the closest-hit loop runs 21 of 32 lanes, the any-hit loop 12).
// Here a lane owns ONE ray query at a time -- the extend ray or the connection ray of its path -- and every loop iteration
// advances every lane's query by a few nodes / one leaf, whichever kind of query it is: the traversal code, 85 % of the
// kernel's instructions, is shared by all lanes at all times.  Lanes whose query is finished wait until PRT_ACQ_SM_BATCH of
// them can be served together: `back` (deposit the echo of a finished connection ray, continue or end the path), `regen`
// (start the next path), `front` (everything between the closest hit and the connection ray: receive element, UltraBSDF,
// amplitude, time bin -- CustomIntegrator.py:153-224) -- the same arithmetic as segment<true>, split at the connection query.
// ------------------------------------------------------------------------------------------------------------------
#ifndef PRT_ACQ_SM_DEFAULT
#define PRT_ACQ_SM_DEFAULT 0
#endif
#ifndef PRT_ACQ_SM_BATCH
#define PRT_ACQ_SM_BATCH 8
#endif
#ifndef PRT_ACQ_SM_NODES
#define PRT_ACQ_SM_NODES 2                // inner-node steps per iteration before the leaf step
#endif

struct Echo {          // what the connection ray decides: deposit `value` into bin `flat` iff it is unoccluded
    long long flat;    // < 0: out of range (nothing to deposit)
    float value;
    bool cont;         // the path goes on after this segment
};

// segment<true> from the closest hit to the connection ray (same statements, same order); returns the connection ray
__device__ __forceinline__ void segment_front(const AcqDev &P, PathState &ps, const Hit &h, Echo &e, float3 &so, float3 &sec, float &vis_tmax) {
    const float dist = h.t;
    ps.geo += dist;                                                              // CI:209 / 315
    const float tof_here = ps.tof + dist / P.c;                                  // CI:165 / 316
    if (!(P.qf & PRT_QF_TOF_LAST_SEGMENT)) ps.tof = tof_here;
    const float u_recv = ps.rng.next_f32();                                      // CI:153 / 319
    const float s1 = ps.rng.next_f32();                                          // CI:173 / 337
    const float s2 = ps.rng.next_f32();                                          // CI:174 / 337
    const float u_rr = ps.rng.next_f32();                                        // CI:219 / 365
    const int recv = min((int) floorf(u_recv * (float) P.n_e), P.n_e - 1);       // CI:154
    const float3 tgt = xpoint(P.T0, P.T1, P.T2, mk3(elem_x(P, recv), 0.0f, 0.0f));   // CI:156-157
    const float3 to_t = tgt - h.p;
    const float dist_recv = sqrtf(dot(to_t, to_t));                              // CI:166 / 329
    sec = mk3(to_t.x / dist_recv, to_t.y / dist_recv, to_t.z / dist_recv);       // CI:158 / 322
    so = spawn_origin(h.p, h.ng, sec);
    vis_tmax = PRT_INF;                                                          // Q1: maxt = inf
    if (P.qf & PRT_QF_CONNECT_TO_TARGET) {
        const float3 q = tgt - so;
        vis_tmax = sqrtf(dot(q, q)) * (1.0f - 1e-4f);
    }
    ps.atten *= expf((P.att_k * dist) / 8.686f);                                 // CI:162-163 / 328
    const float Ttot = (ps.t0 + tof_here) + dist_recv / P.c;                     // CI:167 / 329
    const float phase = P.two_pi_f * Ttot;                                       // CI:168 / 330
    const float3 md = -ps.d;
    const float3 wi = mk3(dot(md, h.fs), dot(md, h.ft), dot(md, h.ns));          // si.wi
    const DMaterial &mat = P.sc.mats[h.material];
    float mZ = __ldg(&mat.p[0]), mA = __ldg(&mat.p[1]);                          // impedance, roughness (CB:12-18)
    if ((P.var_mask >> (h.material & 63)) & 1ull) {                              // finite-difference variant (USMain.py:264)
        if (P.var_index == 0) mZ = P.var_value;
        else mA = P.var_value;
    }
    float3 dir; float pdf, a_resp; bool reflect;
    ultra_bsdf_sample(wi, h.ng, h.ns, mZ, mA, s1, s2, dir, pdf, a_resp, reflect);   // CI:175 / 338
    const float cos_theta = dot(h.ns, md);                                       // CI:176 / 340
    ps.amp *= a_resp * cos_theta * fmaxf(pdf, 1e-6f);                            // CI:177 / 341
    const float kf = rintf(Ttot * P.fs);                                         // CI:191 / 351-352 (half-even)
    int k = (int) kf;
    bool in_range = kf >= 0.0f && kf < (float) P.Tn;
    if (P.qf & PRT_QF_CLAMP_TIDX) {                                              // CI:192
        k = !(kf >= 0.0f) ? 0 : (kf > (float) (P.Tn - 1) ? P.Tn - 1 : k);
        in_range = true;
    }
    const float w_i = directivity_wi(P.nT, sec, P.cos_m, P.cos_c, P.alpha_m, P.alpha_c);
    const float w_o = dot(ps.d, h.ns) / P.n_rays;                                // CI:118,184
    const float press = ps.atten * ps.amp * (w_i * w_o) * sinf(phase);           // CI:187 / 348
    e.flat = in_range ? (long long) (((size_t) ps.a * P.n_e + recv) * (size_t) P.Tn + (size_t) k) : -1;   // CI:197-198
    e.value = press * P.inv_spp;
    ps.d = normalize(dir);                                                       // CI:205-206 / 358-359 (Q9)
    ps.o = spawn_origin(h.p, h.ng, ps.d);
    ps.depth++;                                                                  // CI:210 / 361
    const float prod = ps.atten * ps.amp;
    const float rr = (P.qf & PRT_QF_RR_NO_ABS) ? fminf(prod, 1.0f) : fminf(fabsf(prod), 1.0f);   // CI:220 / 364
    const bool survive = u_rr < rr;                                              // CI:221
    ps.atten = survive ? ps.atten / rr : 0.0f;                                   // CI:224
    e.cont = !(P.qf & PRT_QF_SINGLE_BOUNCE) && survive && dot(ps.d, P.nT) >= P.cos_c &&      // CI:212-223 (Q11)
             ps.depth < P.max_depth && ps.geo < P.max_len;                       // CI:141 / 307
}

template <bool GPRIMS>
__global__ void __launch_bounds__(ACQ_THREADS, 3) k_acquire_sm(const AcqDev P) {
    __shared__ DPrim sprims[MAX_SMEM_PRIMS];
    constexpr bool TRIS_K = false;           // no stash in this kernel: the full table fits
    (void) TRIS_K;
    EchoCache ec;
#if PRT_ACQ_CACHE
    __shared__ unsigned s_ec_key[ECHO_SLOTS(TRIS_K)];
    __shared__ float s_ec_val[ECHO_SLOTS(TRIS_K)];
    ec.key = s_ec_key; ec.val = s_ec_val; ec.bits = ECHO_BITS(TRIS_K);
#else
    ec.key = nullptr; ec.val = nullptr; ec.bits = 0;
#endif
    echo_cache_init(ec);
    const DPrim *prims = GPRIMS ? P.sc.prims : sprims;
    if (!GPRIMS) {
        const float4 *src = reinterpret_cast<const float4 *>(P.sc.prims);
        float4 *dst = reinterpret_cast<float4 *>(sprims);
        for (int i = threadIdx.x; i < P.sc.n_prims * (int) (sizeof(DPrim) / 16); i += blockDim.x) dst[i] = src[i];
        __syncthreads();
    }
    const DScene &sc = P.sc;
    const uint32_t n_ae = (uint32_t) P.a_count * (uint32_t) P.n_e;
    const uint32_t ae0 = (uint32_t) P.a_first * (uint32_t) P.n_e;
    const uint64_t stride = (uint64_t) gridDim.x * blockDim.x;
    const uint64_t j0 = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t d_ae = (uint32_t) (stride % n_ae), d_si = (uint32_t) (stride / n_ae);
    uint32_t ae = (uint32_t) (j0 % n_ae);
    uint64_t si = j0 / n_ae;
    if (P.tx && j0 < n_ae) {                      // transmit-delay table, CI:87-94 / 254-257
        const uint32_t g = ae0 + (uint32_t) j0;
        const float2 scv = __ldg(P.sincos + (int) (g / (uint32_t) P.n_e));
        P.tx[g] = (elem_x(P, (int) (g % (uint32_t) P.n_e)) * scv.x) / P.c;
    }
    Counters cn = { 0, 0, 0, 0, 0 };
    PathState ps;
    Echo echo;
    echo.flat = -1; echo.value = 0.0f; echo.cont = false;
    // the lane's ray query
    const int DONE = 0x7fffffff;
    float3 qo = mk3(0, 0, 0), qinv = mk3(1, 1, 1);
    RayPre rp;
    rp.kx = 0; rp.ky = 1; rp.kz = 2; rp.Sx = rp.Sy = 0.0f; rp.Sz = 1.0f;
    float tbest = 0.0f, b1 = 0.0f, b2 = 0.0f, prim_t = 0.0f;
    int best = -1, best_prim = -1, ref = DONE, sp = 0;
    int stack_ref[PRT_STACK];
    float stack_t[PRT_STACK];
    // 0: no path; 1: extend query running / finished; 2: connection query running / finished
    int kind = 0;
    bool exhausted = false;

    // start a query on (o, d) within [0, tmax]: analytic primitives first (brute force), then the tree below that bound
    auto begin_query = [&](float3 o, float3 d, float tmax, bool any) {
        qo = o;
        qinv = mk3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
        rp = ray_precompute(d);
        best = -1; best_prim = -1; sp = 0;
        tbest = tmax;
        bool blocked = false;
        for (int i = 0; i < sc.n_prims; i++) {
            const float t = intersect_prim(prims[i], o, d, tbest);
            if (t >= 0.0f && (best_prim < 0 || t < tbest)) { best_prim = i; tbest = t; if (any) { blocked = true; break; } }
        }
        prim_t = tbest;
        ref = (blocked || sc.n_small == 0) ? DONE : sc.root_ref;
        cn.rays++;
    };

    for (;;) {
        // ---------------- serve finished lanes, PRT_ACQ_SM_BATCH at a time ----------------
        const bool ready = ref == DONE && !(kind == 0 && exhausted);
        const unsigned mready = __ballot_sync(0xffffffffu, ready);
        const unsigned mtrav = __ballot_sync(0xffffffffu, ref != DONE);
        if (mready && (__popc(mready) >= PRT_ACQ_SM_BATCH || !mtrav)) {
            if (ready && kind == 2) {                                     // back: the connection ray has decided
                const bool visible = best < 0 && best_prim < 0;
                if (visible && echo.flat >= 0) {
                    if (P.buf) echo_add(P, ec, (size_t) echo.flat, echo.value);  // CI:197-203 / 354
                    cn.deposits++;
                }
                kind = echo.cont ? 1 : 0;
                if (echo.cont) begin_query(ps.o, ps.d, PRT_INF, false);
            } else if (ready && kind == 1) {                              // front: the extend ray has its closest hit
                Hit h;
                bool hit = true;
                if (best >= 0 && (best_prim < 0 || tbest < prim_t)) fill_tri_hit(sc, best, tbest, b1, b2, h);
                else if (best_prim >= 0) fill_prim_hit(prims[best_prim], best_prim, ps.o, ps.d, prim_t, h);
                else hit = false;
                if (!hit) {
                    cn.misses++;                                          // CI:146-147 / 309-312
                    kind = 0;
                } else {
                    cn.segments++;
                    float3 so, sec;
                    float vis_tmax;
                    segment_front(P, ps, h, echo, so, sec, vis_tmax);
                    kind = 2;
                    begin_query(so, sec, vis_tmax, true);
                }
            }
            if (ready && kind == 0 && ref == DONE) {                      // regen: next path of this lane
                if (si < P.n_s) {
                    init_path(P, ae0 + ae, P.s_offset + (uint32_t) si * P.s_stride, ps);
                    ae += d_ae;
                    si += d_si;
                    if (ae >= n_ae) { ae -= n_ae; si++; }
                    cn.paths++;
                    if (P.max_depth > 0) {
                        kind = 1;
                        begin_query(ps.o, ps.d, PRT_INF, false);
                    }
                } else {
                    exhausted = true;
                }
            }
        }
        if (!__any_sync(0xffffffffu, ref != DONE)) {
            if (!__any_sync(0xffffffffu, ref == DONE && !(kind == 0 && exhausted))) break;
            continue;
        }
        // ---------------- advance every running query: a few inner nodes, then one leaf ----------------
#pragma unroll
        for (int step = 0; step < PRT_ACQ_SM_NODES; step++) {
            if ((unsigned) ref < (unsigned) DONE) {
                const float4 *n = sc.nodes + 4 * (size_t) ref;
                const float4 q0 = ldg4(n), q1 = ldg4(n + 1), q2 = ldg4(n + 2), q3 = ldg4(n + 3);
                const float tl = box_entry(q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, qo, qinv, tbest);
                const float tr = box_entry(q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, qo, qinv, tbest);
                const int rl = __float_as_int(q3.x), rr = __float_as_int(q3.y);
                const bool hl = tl < PRT_INF, hr = tr < PRT_INF;
                if (hl && hr) {
                    const bool lf = tl <= tr;
                    if (sp < PRT_STACK) { stack_ref[sp] = lf ? rr : rl; stack_t[sp] = lf ? tr : tl; sp++; }
                    ref = lf ? rl : rr;
                } else if (hl || hr) {
                    ref = hl ? rl : rr;
                } else {
                    ref = DONE;
                    while (sp > 0) { --sp; if (stack_t[sp] <= tbest) { ref = stack_ref[sp]; break; } }
                }
            }
        }
        if (ref < 0) {                                                    // a leaf: up to 4 triangles
            const int code = ~ref;
            const int first = code >> 2, count = (code & 3) + 1;
            const RayRows rr = ray_rows(rp);
            bool stop = false;
            for (int j = 0; j < count; j++) {
                const float4 *tv = sc.tri_v + 3 * (size_t) (first + j);
                const float4 a = ldg4(tv), b = ldg4(tv + 1), c = ldg4(tv + 2);
                if (intersect_tri_rows(rr, qo, xyz(a), xyz(b), xyz(c), tbest, b1, b2)) {
                    best = first + j;
                    if (kind == 2) { stop = true; break; }                // any hit
                }
            }
            ref = DONE;
            if (!stop)
                while (sp > 0) { --sp; if (stack_t[sp] <= tbest) { ref = stack_ref[sp]; break; } }
        }
    }
    if (P.buf) echo_cache_flush(P, ec);
    if (P.stats) {
        unsigned v[5] = { cn.paths, cn.segments, cn.rays, cn.deposits, cn.misses };
#pragma unroll
        for (int q = 0; q < 5; q++) {
            unsigned x = v[q];
#pragma unroll
            for (int o = 16; o; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
            if ((threadIdx.x & 31) == 0 && x) atomicAdd(P.stats + q, (unsigned long long) x);
        }
    }
}

__global__ void __launch_bounds__(ACQ_THREADS) k_acquire_trace(const AcqDev P, const uint64_t *__restrict__ path_idx, uint64_t n,
                                                                prt_seg_record *rec) {
    __shared__ DPrim sprims[MAX_SMEM_PRIMS];
    const DPrim *prims = stage_prims(P.sc, sprims);
    uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t path = path_idx[i];
    PathState ps;
    Counters cn = { 0, 0, 0, 0, 0 };
    init_path(P, (uint32_t) (path / P.spp_total), (uint32_t) (path % P.spp_total), ps);
    bool live = P.max_depth > 0;
    EchoCache ec;
    ec.key = nullptr; ec.val = nullptr; ec.bits = 0;         // the decision trace deposits nothing (P.buf == nullptr)
    while (live) live = segment<true>(P, prims, ps, cn, rec + i * (uint64_t) P.max_depth, ec);
}

static int fill_params(prt_scene *s, const prt_acq_params *p, uint64_t seed, uint32_t spp_total, uint32_t s_offset,
                       uint32_t s_stride, AcqDev &P, cudaStream_t st) {
    PRT_REQUIRE(p->n_angles > 0 && p->n_elements > 0 && p->time_samples > 0 && p->max_depth >= 0 && p->angles_deg,
                "acquire: invalid acquisition parameters");
    PRT_REQUIRE(spp_total > 0, "acquire: spp_total must be > 0");
    PRT_REQUIRE((uint64_t) p->n_angles * p->n_elements * (uint64_t) p->time_samples < (1ull << 40), "acquire: channel buffer too large");
    PRT_REQUIRE((uint64_t) p->n_angles * p->n_elements < (1ull << 31), "acquire: more than 2^31 (angle, element) pairs");
    prt_context *c = s->ctx;
    int rc = ensure_scratch(c, 0, 0, (size_t) p->n_angles);
    if (rc) return rc;
    // per-angle (sin, cos) in fp32 exactly as the reference forms them: theta = angle * pi / 180 (CI:78)
    std::vector<float2> sc(p->n_angles);
    for (int a = 0; a < p->n_angles; a++) {
        float theta = (float) p->angles_deg[a] * (float) M_PI / 180.0f;
        sc[a] = make_float2(sinf(theta), cosf(theta));
    }
    const float2 *table = nullptr;
    for (const auto &t : c->angle_tables)
        if (t.host.size() == sc.size() && !memcmp(t.host.data(), sc.data(), sizeof(float2) * sc.size())) { table = t.dev; break; }
    if (!table) {           // first use of this set of angles: one synchronous upload, kept for the life of the context
        if (c->angle_tables.size() >= 64) {
            PRT_CUDA(cudaDeviceSynchronize());
            for (auto &t : c->angle_tables) cudaFree(t.dev);
            c->angle_tables.clear();
        }
        float2 *dev = nullptr;
        PRT_CUDA(cudaMalloc((void **) &dev, sizeof(float2) * sc.size()));
        cudaError_t e = cudaMemcpy(dev, sc.data(), sizeof(float2) * sc.size(), cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaStreamSynchronize(0);    // pageable source: the DMA itself is ordered on the null stream
        if (e != cudaSuccess) { cudaFree(dev); PRT_CUDA(e); }
        c->angle_tables.push_back({sc, dev});
        table = dev;
    }
    (void) st;
    const double *m = p->sensor_to_world;
    P.sc = s->view();
    P.T0 = make_float4((float) m[0], (float) m[1], (float) m[2], (float) m[3]);
    P.T1 = make_float4((float) m[4], (float) m[5], (float) m[6], (float) m[7]);
    P.T2 = make_float4((float) m[8], (float) m[9], (float) m[10], (float) m[11]);
    float nx = P.T0.z, ny = P.T1.z, nz = P.T2.z;  // T * (0,0,1)
    float nl = sqrtf(nx * nx + ny * ny + nz * nz);
    P.nT = make_float3(nx / nl, ny / nl, nz / nl);
    P.c = (float) p->sound_speed;
    P.fs = (float) p->fs;
    P.pitch = (float) p->pitch;
    P.two_pi_f = (float) (2.0 * M_PI * p->frequency);
    P.att_k = (float) (-p->attenuation * p->frequency * 1e-6);
    P.alpha_m = (float) (p->main_beam_deg * M_PI / 180.0);
    P.alpha_c = (float) (p->cutoff_deg * M_PI / 180.0);
    P.cos_c = cosf(P.alpha_c);
    P.cos_m = cosf(P.alpha_m);
    P.max_len = (float) p->max_path_len;
    P.n_rays = (float) (p->n_angles * p->n_elements);
    P.inv_spp = 1.0f / (float) spp_total;
    P.n_a = p->n_angles;
    P.n_e = p->n_elements;
    P.Tn = p->time_samples;
    P.max_depth = p->max_depth;
    P.qf = p->quirk_flags;
    P.sincos = table;
    P.seed = seed;
    P.spp_total = spp_total;
    P.s_offset = s_offset;
    P.s_stride = s_stride ? s_stride : 1;
    P.n_s = s_offset < spp_total ? ((uint64_t) spp_total - s_offset + P.s_stride - 1) / P.s_stride : 0;
    P.total = P.n_s * (uint64_t) p->n_angles * (uint64_t) p->n_elements;
    P.a_first = 0;
    P.a_count = p->n_angles;
    P.wae = 0;
    P.var_mask = 0ull;
    P.var_index = 0;
    P.var_value = 0.0f;
    P.buf = nullptr;
    P.tx = nullptr;
    P.stats = nullptr;
    return PRT_OK;
}

static int launch_acquire(prt_context *c, const AcqDev &P, cudaStream_t st) {
    // persistent grid: a whole number of CTAs per SM (occupancy-derived), never more than the work needs
    const bool tris = P.sc.n_tris > 0;
    int per_sm = 0;
    const bool gp = P.sc.n_prims > MAX_SMEM_PRIMS;
    void (*kern)(const AcqDev) = tris ? (gp ? k_acquire<true, true> : k_acquire<true, false>)
                                      : (gp ? k_acquire<false, true> : k_acquire<false, false>);
    // mesh scenes: the per-lane state machine (PRT_ACQ_SM=0 restores the segment-at-a-time kernel for A/B)
    static const int use_sm = [] { const char *e = getenv("PRT_ACQ_SM"); return e && *e ? atoi(e) : PRT_ACQ_SM_DEFAULT; }();
    const bool sm = tris && use_sm && P.sc.n_small == P.sc.n_tris;
    if (sm) kern = gp ? k_acquire_sm<true> : k_acquire_sm<false>;
    PRT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, ACQ_THREADS, 0));
    if (per_sm < 1) per_sm = 1;
    const uint64_t n_ae_l = (uint64_t) P.a_count * P.n_e;
    AcqDev Q = P;
    Q.wae = !sm && (PRT_ACQ_WAE == 2 || (PRT_ACQ_WAE == 1 && tris)) && P.n_s >= 64;
    uint64_t want = ((Q.wae ? (P.n_s + 31) / 32 * 32 : P.n_s) * n_ae_l + ACQ_THREADS - 1) / ACQ_THREADS;
    uint64_t n_ae_blocks = (n_ae_l + ACQ_THREADS - 1) / ACQ_THREADS;
    if (want < n_ae_blocks) want = n_ae_blocks;  // the tx-delay table is written by the first n_a*n_e threads
    uint64_t grid = (uint64_t) c->sm_count * per_sm;
    if (grid > want) grid = want;
    if (grid < 1) grid = 1;
    {
        ProfScope ps(c, PRT_KC_ACQUIRE, st);
        kern<<<(unsigned) grid, ACQ_THREADS, 0, st>>>(Q);
    }
    PRT_CUDA(cudaGetLastError());
    return PRT_OK;
}

}  // namespace prt

using namespace prt;

extern "C" {

int prt_acquire_dev(prt_scene *s, const prt_acq_params *p, uint64_t seed, uint32_t spp_total, uint32_t sample_offset,
                    uint32_t sample_stride, float *channel_buf_dev, float *tx_delays_dev, uint64_t *stats_dev, void *stream) {
    PRT_REQUIRE(s && p, "prt_acquire_dev: null argument");
    return prt_acquire_dev_angles(s, p, seed, spp_total, sample_offset, sample_stride, 0, p->n_angles, channel_buf_dev, tx_delays_dev,
                                  stats_dev, stream);
}

int prt_acquire_dev_angles(prt_scene *s, const prt_acq_params *p, uint64_t seed, uint32_t spp_total, uint32_t sample_offset,
                           uint32_t sample_stride, int32_t angle_first, int32_t angle_count, float *channel_buf_dev,
                           float *tx_delays_dev, uint64_t *stats_dev, void *stream) {
    PRT_REQUIRE(s && p && channel_buf_dev, "prt_acquire_dev: null argument");
    PRT_REQUIRE(angle_first >= 0 && angle_count >= 0 && angle_first + angle_count <= p->n_angles, "prt_acquire_dev_angles: angle range out of bounds");
    if (angle_count == 0) return PRT_OK;
    if (!s->committed) { set_error("prt_acquire_dev: scene not committed"); return PRT_ERR_STATE; }
    std::lock_guard<std::mutex> lk(s->ctx->mtx);
    PRT_CUDA(cudaSetDevice(s->ctx->device));
    return prt::acquire_enqueue(s, p, seed, spp_total, sample_offset, sample_stride, angle_first, angle_count, channel_buf_dev,
                                tx_delays_dev, stats_dev, (cudaStream_t) stream);
}

}  // extern "C"

namespace prt {
// the launches of prt_acquire_dev_angles; the caller holds the context mutex and has set the device (also used by
// prt_us_render in prt_das.cu, which chains the post-processing kernels behind it on the same stream)
int acquire_enqueue(prt_scene *s, const prt_acq_params *p, uint64_t seed, uint32_t spp_total, uint32_t sample_offset,
                    uint32_t sample_stride, int32_t angle_first, int32_t angle_count, float *channel_buf_dev, float *tx_delays_dev,
                    uint64_t *stats_dev, cudaStream_t st) {
    AcqDev P;
    int rc = fill_params(s, p, seed, spp_total, sample_offset, sample_stride, P, st);
    if (rc) return rc;
    P.buf = channel_buf_dev;
    P.tx = tx_delays_dev;
    P.stats = reinterpret_cast<unsigned long long *>(stats_dev);
    // One launch per steering angle.  With stride = grid * 256 a multiple of n_e, every thread then keeps ONE
    // (angle, element) for all of its paths, so the lanes of a warp stay on 32 neighbouring elements of one angle
    // however their paths regenerate; in a single launch over all angles the lanes drift onto different angles
    // and the warp's rays decohere (measured on B200: ring 42.2 -> 26.5 ms, Sphere_Box intended 13.6 -> 10.3 ms).
    // PRT_ACQ_SPLIT=0 restores the single launch.
    static const bool split = [] { const char *e = getenv("PRT_ACQ_SPLIT"); return !(e && e[0] == '0'); }();
    P.a_first = angle_first;
    P.a_count = angle_count;
    if (!split || angle_count == 1) return launch_acquire(s->ctx, P, st);
    for (int a = angle_first; a < angle_first + angle_count; a++) {
        AcqDev Q = P;
        Q.a_first = a;
        Q.a_count = 1;
        rc = launch_acquire(s->ctx, Q, st);
        if (rc) return rc;
    }
    return PRT_OK;
}
}  // namespace prt

extern "C" {

int prt_acquire(prt_scene *s, const prt_acq_params *p, uint64_t seed, uint32_t spp_total, uint32_t sample_offset,
                uint32_t sample_stride, float *channel_buf, float *tx_delays, prt_acq_stats *stats) {
    PRT_REQUIRE(s && p && channel_buf, "prt_acquire: null argument");
    if (!s->committed) { set_error("prt_acquire: scene not committed"); return PRT_ERR_STATE; }
    std::lock_guard<std::mutex> lk(s->ctx->mtx);
    prt_context *c = s->ctx;
    PRT_CUDA(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    AcqDev P;
    int rc = fill_params(s, p, seed, spp_total, sample_offset, sample_stride, P, st);
    if (rc) return rc;
    const size_t n_buf = (size_t) p->n_angles * p->n_elements * (size_t) p->time_samples;
    const size_t n_tx = (size_t) p->n_angles * p->n_elements;
    rc = ensure_scratch(c, n_buf, n_tx, (size_t) p->n_angles);
    if (rc) return rc;
    ScopedEvents<4> ev;
    PRT_REQUIRE(ev.ok, "cudaEventCreate failed");
    cudaEvent_t e0 = ev.e[0], e1 = ev.e[1], e2 = ev.e[2], e3 = ev.e[3];
    PRT_CUDA(cudaEventRecord(e0, st));
    PRT_CUDA(cudaMemsetAsync(c->acc_dev, 0, sizeof(float) * n_buf, st));
    PRT_CUDA(cudaMemsetAsync(c->stats_dev, 0, sizeof(uint64_t) * 8, st));
    P.buf = c->acc_dev;
    P.tx = c->aux_dev;
    P.stats = reinterpret_cast<unsigned long long *>(c->stats_dev);
    PRT_CUDA(cudaEventRecord(e1, st));
    // Is the destination page-locked (prt_host_alloc / cudaHostRegister)?  Then copy straight into it, one angle
    // slice at a time on the copy stream while the next angle's paths are still being traced.
    cudaPointerAttributes attr;
    bool pinned_dst = cudaPointerGetAttributes(&attr, channel_buf) == cudaSuccess && attr.type == cudaMemoryTypeHost;
    cudaGetLastError();
    float *pin = reinterpret_cast<float *>(c->pinned);
    unsigned launches = 0;
    if (pinned_dst && p->n_angles > 1) {
        const size_t slice = (size_t) p->n_elements * (size_t) p->time_samples;
        for (int a = 0; a < p->n_angles; a++) {
            AcqDev Q = P;
            Q.a_first = a;
            Q.a_count = 1;
            rc = launch_acquire(c, Q, st);
            if (rc) return rc;
            launches++;
            PRT_CUDA(cudaEventRecord(c->slice_done[a & 1], st));
            PRT_CUDA(cudaStreamWaitEvent(c->copy_stream, c->slice_done[a & 1], 0));
            PRT_CUDA(cudaMemcpyAsync(channel_buf + a * slice, c->acc_dev + a * slice, sizeof(float) * slice, cudaMemcpyDeviceToHost,
                                     c->copy_stream));
        }
        PRT_CUDA(cudaEventRecord(e2, st));
        PRT_CUDA(cudaEventRecord(c->slice_done[0], c->copy_stream));
        PRT_CUDA(cudaStreamWaitEvent(st, c->slice_done[0], 0));
    } else {
        rc = launch_acquire(c, P, st);
        if (rc) return rc;
        launches = 1;
        PRT_CUDA(cudaEventRecord(e2, st));
        PRT_CUDA(cudaMemcpyAsync(pinned_dst ? channel_buf : pin, c->acc_dev, sizeof(float) * n_buf, cudaMemcpyDeviceToHost, st));
    }
    PRT_CUDA(cudaMemcpyAsync(pin + n_buf, c->aux_dev, sizeof(float) * n_tx, cudaMemcpyDeviceToHost, st));
    uint64_t hs[8];
    PRT_CUDA(cudaMemcpyAsync(hs, c->stats_dev, sizeof(uint64_t) * 8, cudaMemcpyDeviceToHost, st));
    PRT_CUDA(cudaEventRecord(e3, st));
    PRT_CUDA(cudaStreamSynchronize(st));
    if (!pinned_dst) memcpy(channel_buf, pin, sizeof(float) * n_buf);
    if (tx_delays) memcpy(tx_delays, pin + n_buf, sizeof(float) * n_tx);
    if (stats) {
        stats->paths = hs[0]; stats->segments = hs[1]; stats->rays = hs[2]; stats->deposits = hs[3]; stats->misses = hs[4];
        PRT_CUDA(cudaEventElapsedTime(&stats->kernel_ms, e1, e2));
        PRT_CUDA(cudaEventElapsedTime(&stats->total_ms, e0, e3));
        stats->launches = launches;
        stats->_pad = 0;
    }
    return PRT_OK;
}

// "next" row f2: the driver's finite-difference loop (USMain.py:262-289) runs f(rough) and
// f(rough + eps) as two full acquisitions with a parameter patch in between.  Here all variants of ONE material
// parameter are traced in one call: same seed and per-path PCG32 streams for every variant (common random numbers,
// so the difference of two planes is not buried in Monte-Carlo noise), no parameter upload or BVH touch in between,
// one zero-fill, one result transfer.  The override travels in the kernel parameter block.
int prt_acquire_variants(prt_scene *s, const prt_acq_params *p, uint64_t seed, uint32_t spp_total, uint32_t sample_offset,
                         uint32_t sample_stride, uint64_t material_mask, int param_index, const double *values, uint32_t n_values,
                         float *channel_bufs, float *tx_delays, prt_acq_stats *stats) {
    PRT_REQUIRE(s && p && channel_bufs && values, "prt_acquire_variants: null argument");
    PRT_REQUIRE(n_values >= 1 && n_values <= PRT_MAX_VARIANTS, "prt_acquire_variants: 1..PRT_MAX_VARIANTS values");
    if (!s->committed) { set_error("prt_acquire_variants: scene not committed"); return PRT_ERR_STATE; }
    PRT_REQUIRE(material_mask != 0 && (param_index == 0 || param_index == 1),
                "prt_acquire_variants: empty material mask / parameter index out of range (0 = impedance, 1 = roughness)");
    PRT_REQUIRE(s->mats.size() <= 64, "prt_acquire_variants: more than 64 materials");
    for (size_t m = 0; m < 64; m++)
        if ((material_mask >> m) & 1ull)
            PRT_REQUIRE(m < s->mats.size() && s->mats[m].kind == PRT_MAT_ULTRA, "prt_acquire_variants: mask selects a material that is not an ultrasound_bsdf");
    std::lock_guard<std::mutex> lk(s->ctx->mtx);
    prt_context *c = s->ctx;
    PRT_CUDA(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    AcqDev P;
    int rc = fill_params(s, p, seed, spp_total, sample_offset, sample_stride, P, st);
    if (rc) return rc;
    const size_t n_buf = (size_t) p->n_angles * p->n_elements * (size_t) p->time_samples;
    const size_t n_tx = (size_t) p->n_angles * p->n_elements;
    rc = ensure_scratch(c, n_buf * n_values, n_tx, (size_t) p->n_angles);
    if (rc) return rc;
    ScopedEvents<3> ev;
    PRT_REQUIRE(ev.ok, "cudaEventCreate failed");
    cudaEvent_t e0 = ev.e[0], e1 = ev.e[1], e2 = ev.e[2];
    PRT_CUDA(cudaEventRecord(e0, st));
    PRT_CUDA(cudaMemsetAsync(c->acc_dev, 0, sizeof(float) * n_buf * n_values, st));
    PRT_CUDA(cudaMemsetAsync(c->stats_dev, 0, sizeof(uint64_t) * 8 * PRT_MAX_VARIANTS, st));
    unsigned launches = 0;
    for (uint32_t v = 0; v < n_values; v++) {
        AcqDev Q = P;
        Q.buf = c->acc_dev + v * n_buf;
        Q.tx = v == 0 ? c->aux_dev : nullptr;
        Q.stats = reinterpret_cast<unsigned long long *>(c->stats_dev) + 8 * v;
        Q.var_mask = material_mask;
        Q.var_index = param_index;
        Q.var_value = (float) values[v];
        for (int a = 0; a < p->n_angles; a++) {       // one launch per steering angle, as in prt_acquire_dev
            Q.a_first = a;
            Q.a_count = 1;
            rc = launch_acquire(c, Q, st);
            if (rc) return rc;
            launches++;
        }
    }
    PRT_CUDA(cudaEventRecord(e1, st));
    cudaPointerAttributes attr;
    const bool pinned_dst = cudaPointerGetAttributes(&attr, channel_bufs) == cudaSuccess && attr.type == cudaMemoryTypeHost;
    cudaGetLastError();
    float *pin = reinterpret_cast<float *>(c->pinned);
    PRT_CUDA(cudaMemcpyAsync(pinned_dst ? channel_bufs : pin, c->acc_dev, sizeof(float) * n_buf * n_values, cudaMemcpyDeviceToHost, st));
    PRT_CUDA(cudaMemcpyAsync(pin + n_buf * n_values, c->aux_dev, sizeof(float) * n_tx, cudaMemcpyDeviceToHost, st));
    uint64_t hs[8 * PRT_MAX_VARIANTS];
    PRT_CUDA(cudaMemcpyAsync(hs, c->stats_dev, sizeof hs, cudaMemcpyDeviceToHost, st));
    PRT_CUDA(cudaEventRecord(e2, st));
    PRT_CUDA(cudaStreamSynchronize(st));
    if (!pinned_dst) memcpy(channel_bufs, pin, sizeof(float) * n_buf * n_values);
    if (tx_delays) memcpy(tx_delays, pin + n_buf * n_values, sizeof(float) * n_tx);
    if (stats) {
        float k_ms = 0.0f, t_ms = 0.0f;
        PRT_CUDA(cudaEventElapsedTime(&k_ms, e0, e1));
        PRT_CUDA(cudaEventElapsedTime(&t_ms, e0, e2));
        for (uint32_t v = 0; v < n_values; v++) {
            const uint64_t *h = hs + 8 * v;
            stats[v].paths = h[0]; stats[v].segments = h[1]; stats[v].rays = h[2]; stats[v].deposits = h[3]; stats[v].misses = h[4];
            stats[v].kernel_ms = k_ms; stats[v].total_ms = t_ms;     // of the whole call
            stats[v].launches = launches;
            stats[v]._pad = 0;
        }
    }
    return PRT_OK;
}

int prt_acquire_trace(prt_scene *s, const prt_acq_params *p, uint64_t seed, uint32_t spp_total, const uint64_t *path_idx,
                      uint64_t n, prt_seg_record *rec) {
    PRT_REQUIRE(s && p && path_idx && rec, "prt_acquire_trace: null argument");
    if (!s->committed) { set_error("prt_acquire_trace: scene not committed"); return PRT_ERR_STATE; }
    if (n == 0) return PRT_OK;
    std::lock_guard<std::mutex> lk(s->ctx->mtx);
    prt_context *c = s->ctx;
    PRT_CUDA(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    AcqDev P;
    int rc = fill_params(s, p, seed, spp_total, 0, 1, P, st);
    if (rc) return rc;
    const uint64_t limit = (uint64_t) p->n_angles * p->n_elements * (uint64_t) spp_total;
    for (uint64_t i = 0; i < n; i++) PRT_REQUIRE(path_idx[i] < limit, "prt_acquire_trace: path index out of range");
    uint64_t *idx_d = nullptr;
    prt_seg_record *rec_d = nullptr;
    const size_t nrec = (size_t) n * (size_t) (p->max_depth > 0 ? p->max_depth : 1);
    cudaError_t e = cudaMalloc(&idx_d, sizeof(uint64_t) * n);
    if (e == cudaSuccess) e = cudaMalloc(&rec_d, sizeof(prt_seg_record) * nrec);
    if (e == cudaSuccess) e = cudaMemcpyAsync(idx_d, path_idx, sizeof(uint64_t) * n, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemsetAsync(rec_d, 0, sizeof(prt_seg_record) * nrec, st);
    if (e == cudaSuccess) {
        k_acquire_trace<<<(unsigned) ((n + ACQ_THREADS - 1) / ACQ_THREADS), ACQ_THREADS, 0, st>>>(P, idx_d, n, rec_d);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(rec, rec_d, sizeof(prt_seg_record) * (size_t) n * (size_t) p->max_depth, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(idx_d);
    cudaFree(rec_d);
    PRT_CUDA(e);
    return PRT_OK;
}

}  // extern "C"
