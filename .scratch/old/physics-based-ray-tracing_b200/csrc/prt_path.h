// prt_path.h -- what the two path-tracing back ends share (prt_path.cu: tile megakernel, prt_wavefront.cu:
// wavefront pipeline): the parameter block, the per-path state, camera ray generation and the SHADING half of one
// iteration of Mitsuba's `path` loop.  Both back ends run this code on the same per-path
// PCG32 stream, so a path takes the same decisions and accumulates the same radiance bit for bit in either; they
// differ only in where the two ray queries (closest hit, shadow ray) execute.
#pragma once
#include "prt_internal.h"

namespace prt {

struct PtDev {
    DScene sc;
    float4 T0, T1, T2;           // camera to_world rows
    float tan_x, tan_y, near_clip;
    int W, H, max_depth, rr_depth, tent;
    uint64_t seed;
    uint32_t spp_total, s_offset, s_stride, n_s;
    int tiles_x, tiles_y;
    unsigned kind_mask;          // bit k set iff some material has kind k (PRT_MAT_*)
    float3 box_lo, box_hi;       // bounds of the triangle geometry (ray-sort cells, prt_wavefront.cu)
    float *film;                 // [H][W][4]
    unsigned long long *stats;   // {paths, segments, rays, shadow_rays}
};

static constexpr int PT_TILE = 16;
static constexpr int PT_HALO = PT_TILE + 2;
static constexpr int MAX_SMEM_PRIMS = 64;

struct PtState {
    float3 o, d, thr, res, prev_p;
    float eta, prev_pdf, px, py;
    int depth;
    bool prev_delta;
    Pcg32 rng;
};

// a deferred next-event-estimation term: res += c * w iff the segment (o, d, tmax) is unoccluded
struct ShadowReq {
    bool   want;
    float3 o, d, c;
    float  tmax, w;
};

__device__ __forceinline__ void pt_init(const PtDev &P, int x, int y, uint32_t s, PtState &st) {
    uint64_t path = ((uint64_t) y * (uint64_t) P.W + (uint64_t) x) * (uint64_t) P.spp_total + (uint64_t) s;
    st.rng = path_rng(P.seed, path);
    float jx = st.rng.next_f32(), jy = st.rng.next_f32();
    st.px = (float) x + jx;
    st.py = (float) y + jy;
    float sx = st.px / (float) P.W, sy = st.py / (float) P.H;
    float3 dl = normalize(mk3((1.0f - 2.0f * sx) * P.tan_x, (1.0f - 2.0f * sy) * P.tan_y, 1.0f));
    float3 o = mk3(P.T0.w, P.T1.w, P.T2.w);
    float3 d = xvec(P.T0, P.T1, P.T2, dl);
    float tn = P.near_clip / dl.z;
    st.o = mk3(fmaf(d.x, tn, o.x), fmaf(d.y, tn, o.y), fmaf(d.z, tn, o.z));
    st.d = d;
    st.thr = mk3(1.0f, 1.0f, 1.0f);
    st.res = mk3(0.0f, 0.0f, 0.0f);
    st.prev_p = st.o;
    st.eta = 1.0f;
    st.prev_pdf = 1.0f;
    st.prev_delta = true;
    st.depth = 0;
}

__device__ __forceinline__ void pt_apply_shadow(PtState &st, const ShadowReq &sr) {
    st.res.x = fmaf(sr.c.x, sr.w, st.res.x);
    st.res.y = fmaf(sr.c.y, sr.w, st.res.y);
    st.res.z = fmaf(sr.c.z, sr.w, st.res.z);
}

// Shading half of one iteration of path.cpp's loop, given the closest hit `h` of the ray (st.o, st.d) (`valid` =
// something was hit).  Adds directly seen emission (MIS against emitter sampling at the previous vertex), draws the
// emitter sample and reports it as a shadow-ray request, samples the BSDF, applies Russian roulette and writes the
// next ray into st.  Returns false when the path is finished.
// KSEL >= 0: the caller knows the material kind of every hit it passes (the wavefront's per-material shading kernels), so
// the other materials' code -- and their registers -- drop out at compile time; KSEL < 0: read it from the material.
template <int KSEL = -1>
__device__ __forceinline__ bool pt_shade(const PtDev &P, PtState &st, const Hit &h, bool valid, ShadowReq &sr) {
    sr.want = false;
    const float3 md = -st.d;
    int kind = PRT_MAT_NULL;
    float3 refl_rgb = mk3(0.0f, 0.0f, 0.0f);
    float p0 = 0.0f, p1 = 1.0f;
    if (valid) {
        const DMaterial &m = P.sc.mats[h.material];
        kind = KSEL >= 0 ? KSEL : __ldg(&m.kind);
        p0 = __ldg(&m.p[0]);
        p1 = __ldg(&m.p[1]);
        refl_rgb = mk3(p0, p1, __ldg(&m.p[2]));
        float3 Le = mk3(__ldg(&m.emission[0]), __ldg(&m.emission[1]), __ldg(&m.emission[2]));
        // ---- direct emission, MIS against emitter sampling at the previous vertex ----
        if ((Le.x > 0.0f || Le.y > 0.0f || Le.z > 0.0f) && dot(md, h.ns) > 0.0f) {
            float em_pdf = 0.0f;
            if (!st.prev_delta) {
                float3 dv = h.p - st.prev_p;
                int ei = __ldg(P.sc.shape_emitter + h.shape);
                em_pdf = __ldg(P.sc.em_inv_area + ei) * dot(dv, dv) / fabsf(dot(st.d, h.ns)) / (float) P.sc.n_emitters;
                if (!isfinite(em_pdf)) em_pdf = 0.0f;
            }
            float w = mis_weight(st.prev_pdf, em_pdf);
            st.res.x = fmaf(st.thr.x * Le.x, w, st.res.x);
            st.res.y = fmaf(st.thr.y * Le.y, w, st.res.y);
            st.res.z = fmaf(st.thr.z * Le.z, w, st.res.z);
        }
    }
    if (!(st.depth + 1 < P.max_depth) || !valid) return false;
    const float3 wi = mk3(dot(md, h.fs), dot(md, h.ft), dot(md, h.ns));
    // ---- emitter sampling (smooth BSDFs only) ----
    if (kind == PRT_MAT_DIFFUSE && P.sc.n_emitters > 0) {
        float u1 = st.rng.next_f32(), u2 = st.rng.next_f32();
        float fe = u1 * (float) P.sc.n_emitters;
        int ei = min((int) fe, P.sc.n_emitters - 1);
        u1 = fe - (float) ei;
        int f0 = __ldg(P.sc.em_first + ei), f1 = __ldg(P.sc.em_first + ei + 1);
        float total = __ldg(&P.sc.em_tri[3 * (size_t) (f1 - 1)].w);
        float target = u2 * total;
        int f = f0;
        while (f < f1 - 1 && __ldg(&P.sc.em_tri[3 * (size_t) f].w) < target) f++;
        float lo = f > f0 ? __ldg(&P.sc.em_tri[3 * (size_t) (f - 1)].w) : 0.0f;
        float4 a = ldg4(P.sc.em_tri + 3 * (size_t) f), b = ldg4(P.sc.em_tri + 3 * (size_t) f + 1), c = ldg4(P.sc.em_tri + 3 * (size_t) f + 2);
        u2 = (target - lo) / (a.w - lo);
        float tq = sqrtf(fmaxf(1.0f - u1, 0.0f));
        float b1 = 1.0f - tq, b2 = tq * u2;
        float3 e0 = xyz(b) - xyz(a), e1 = xyz(c) - xyz(a);
        float3 ps = xyz(a) + e0 * b1 + e1 * b2;
        float3 pn = normalize(cross(e0, e1));
        if (c.w != 0.0f) pn = -pn;
        float3 dv = ps - h.p;
        float dist2 = dot(dv, dv);
        float3 dd = dv * (1.0f / sqrtf(dist2));
        float x = dist2 / fabsf(dot(dd, pn));
        float pdf = __ldg(P.sc.em_inv_area + ei) * (isfinite(x) ? x : 0.0f);
        if (dot(dd, pn) < 0.0f && pdf != 0.0f) {
            pdf /= (float) P.sc.n_emitters;
            const float ci = wi.z, co = dot(dd, h.ns);
            // the shadow ray is only needed when the contribution can be non-zero (one-sided diffuse)
            if (ci > 0.0f && co > 0.0f) {
                float3 so = spawn_origin(h.p, h.ng, dd);
                float3 sv = ps - so;
                float sd = sqrtf(dot(sv, sv));
                const DMaterial &me = P.sc.mats[__float_as_int(b.w)];
                float f_cos = co * (1.0f / PRT_PI_F);
                sr.want = true;
                sr.o = so;
                sr.d = sv * (1.0f / sd);
                sr.tmax = sd * (1.0f - 10.0f * PRT_RAY_EPSILON);
                sr.w = mis_weight(pdf, f_cos) * f_cos / pdf;
                sr.c = mk3(st.thr.x * refl_rgb.x * __ldg(&me.emission[0]), st.thr.y * refl_rgb.y * __ldg(&me.emission[1]),
                           st.thr.z * refl_rgb.z * __ldg(&me.emission[2]));
            }
        }
    }
    // ---- BSDF sampling ----
    const float s1 = st.rng.next_f32();
    const float s2x = st.rng.next_f32(), s2y = st.rng.next_f32();
    float3 wo, bw = mk3(0.0f, 0.0f, 0.0f);
    float bs_pdf, bs_eta = 1.0f;
    bool bs_delta = false;
    if (kind == PRT_MAT_DIFFUSE) {
        float dx, dy;
        disk_concentric(s2x, s2y, dx, dy);
        float z = sqrtf(fmaxf(1.0f - dx * dx - dy * dy, 0.0f));
        wo = mk3(dx, dy, z);
        bs_pdf = z * (1.0f / PRT_PI_F);
        if (wi.z > 0.0f && bs_pdf > 0.0f) bw = refl_rgb;
    } else if (kind == PRT_MAT_DIELECTRIC) {
        float ct, eit, eti;
        float r = fresnel_dielectric(wi.z, p0 / p1, ct, eit, eti);
        bs_delta = true;
        if (s1 <= r) {
            wo = mk3(-wi.x, -wi.y, wi.z);
            bs_pdf = r;
            bw = mk3(1.0f, 1.0f, 1.0f);
        } else {
            wo = mk3(-eti * wi.x, -eti * wi.y, ct);
            bs_pdf = 1.0f - r;
            bs_eta = eit;
            bw = mk3(eti * eti, eti * eti, eti * eti);
        }
    } else if (kind == PRT_MAT_CONDUCTOR) {
        bs_delta = true;
        wo = mk3(-wi.x, -wi.y, wi.z);
        bs_pdf = 1.0f;
        if (wi.z > 0.0f) bw = refl_rgb;
    } else {
        return false;
    }
    const float3 wd = h.fs * wo.x + h.ft * wo.y + h.ns * wo.z;
    st.prev_p = h.p;
    st.o = spawn_origin(h.p, h.ng, wd);
    st.d = wd;
    st.thr = mk3(st.thr.x * bw.x, st.thr.y * bw.y, st.thr.z * bw.z);
    st.eta *= bs_eta;
    st.prev_pdf = bs_pdf;
    st.prev_delta = bs_delta;
    st.depth++;
    const float tmax = fmaxf(st.thr.x, fmaxf(st.thr.y, st.thr.z));
    const float rr_prob = fminf(tmax * st.eta * st.eta, 0.95f);
    const float u_rr = st.rng.next_f32();
    const bool rr_active = st.depth >= P.rr_depth;
    if (rr_active) {
        float inv = 1.0f / rr_prob;
        st.thr = st.thr * inv;
    }
    return !((rr_active && !(u_rr < rr_prob)) || tmax == 0.0f);
}

// ImageBlock::put of one finished sample into a shared-memory RGBW tile whose origin is pixel (tx0 - 1, ty0 - 1):
// tent filter of radius 1 (pixel centres at i + 0.5) or box
__device__ __forceinline__ void pt_splat(int tent, float4 *tile, int tx0, int ty0, float px, float py, float3 res) {
    if (!tent) {
        int x = (int) floorf(px) - tx0 + 1, y = (int) floorf(py) - ty0 + 1;
        float *q = reinterpret_cast<float *>(tile + y * PT_HALO + x);
        atomicAdd(q, res.x); atomicAdd(q + 1, res.y); atomicAdd(q + 2, res.z); atomicAdd(q + 3, 1.0f);
        return;
    }
    int x0 = (int) floorf(px - 0.5f), y0 = (int) floorf(py - 0.5f);
#pragma unroll
    for (int dy = 0; dy < 2; dy++)
#pragma unroll
        for (int dx = 0; dx < 2; dx++) {
            int x = x0 + dx, y = y0 + dy;
            float wx = fmaxf(1.0f - fabsf((float) x + 0.5f - px), 0.0f), wy = fmaxf(1.0f - fabsf((float) y + 0.5f - py), 0.0f);
            float w = wx * wy;
            int lx = x - tx0 + 1, ly = y - ty0 + 1;
            if (w > 0.0f && lx >= 0 && ly >= 0 && lx < PT_HALO && ly < PT_HALO) {
                float *q = reinterpret_cast<float *>(tile + ly * PT_HALO + lx);
                atomicAdd(q, res.x * w); atomicAdd(q + 1, res.y * w); atomicAdd(q + 2, res.z * w); atomicAdd(q + 3, w);
            }
        }
}

// finished shared-memory tile -> global film (halo pixels belong to neighbouring tiles as well: atomics)
__device__ __forceinline__ void pt_flush_tile(const PtDev &P, const float4 *tile, int tx0, int ty0) {
    for (int i = threadIdx.x; i < PT_HALO * PT_HALO; i += blockDim.x) {
        int gx = tx0 - 1 + (i % PT_HALO), gy = ty0 - 1 + (i / PT_HALO);
        float4 v = tile[i];
        if (gx >= 0 && gy >= 0 && gx < P.W && gy < P.H && v.w != 0.0f) {
            float *q = P.film + 4 * ((size_t) gy * P.W + gx);
            atomicAdd(q, v.x); atomicAdd(q + 1, v.y); atomicAdd(q + 2, v.z); atomicAdd(q + 3, v.w);
        }
    }
}

// wavefront pipeline: generate -> extend (dynamic ray fetch) -> per-material shade -> shadow, until every path is done
int launch_wavefront(prt_context *c, const PtDev &P, cudaStream_t st);

}  // namespace prt
