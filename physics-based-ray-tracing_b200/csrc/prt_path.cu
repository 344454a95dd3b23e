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

#include "prt_internal.h"

namespace prt {

static constexpr int PT_THREADS = 256;
static constexpr int PT_TILE = 16;
static constexpr int PT_HALO = PT_TILE + 2;
static constexpr int MAX_SMEM_PRIMS = 64;

struct PtDev {
    DScene sc;
    float4 T0, T1, T2;           // camera to_world rows
    float tan_x, tan_y, near_clip;
    int W, H, max_depth, rr_depth, tent;
    uint64_t seed;
    uint32_t spp_total, s_offset, s_stride, n_s;
    int tiles_x, tiles_y;
    float *film;                 // [H][W][4]
    unsigned long long *stats;   // {paths, segments, rays, shadow_rays}
};

__device__ __forceinline__ float mis_weight(float a, float b) {
    a *= a;
    b *= b;
    float w = a / (a + b);
    return isfinite(w) ? w : 0.0f;
}

__device__ __forceinline__ void disk_concentric(float ux, float uy, float &ox, float &oy) {
    float x = fmaf(2.0f, ux, -1.0f), y = fmaf(2.0f, uy, -1.0f);
    if (x == 0.0f && y == 0.0f) { ox = 0.0f; oy = 0.0f; return; }
    bool q = fabsf(x) < fabsf(y);
    float r = q ? y : x, rp = q ? x : y;
    float phi = 0.25f * PRT_PI_F * rp / r;
    if (q) phi = 0.5f * PRT_PI_F - phi;
    float s, c;
    sincosf(phi, &s, &c);
    ox = r * c;
    oy = r * s;
}

// mitsuba fresnel(cos_theta_i, eta)
__device__ __forceinline__ float fresnel_dielectric(float cos_i, float eta, float &cos_t, float &eta_it, float &eta_ti) {
    bool outside = cos_i >= 0.0f;
    float rcp_eta = 1.0f / eta;
    eta_it = outside ? eta : rcp_eta;
    eta_ti = outside ? rcp_eta : eta;
    float ct2 = 1.0f - (1.0f - cos_i * cos_i) * eta_ti * eta_ti;
    float ci = fabsf(cos_i), ct = sqrtf(fmaxf(ct2, 0.0f));
    float a_s = (-eta_it * ct + ci) / (eta_it * ct + ci);
    float a_p = (-eta_it * ci + ct) / (eta_it * ci + ct);
    float r = 0.5f * (a_s * a_s + a_p * a_p);
    if (eta == 1.0f) r = 0.0f;
    else if (ci == 0.0f) r = 1.0f;
    cos_t = cos_i >= 0.0f ? -ct : ct;
    return r;
}

struct PtState {
    float3 o, d, thr, res, prev_p;
    float eta, prev_pdf, px, py;
    int depth;
    bool prev_delta;
    Pcg32 rng;
};

struct PtCounters {
    unsigned paths, segments, rays, shadow;
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

// one iteration of path.cpp's loop; returns false when the path is finished
__device__ __forceinline__ bool pt_step(const PtDev &P, const DPrim *prims, PtState &st, PtCounters &cn) {
    Hit h;
    cn.rays++;
    const bool valid = closest_hit<true>(P.sc, prims, st.o, st.d, PRT_INF, h);
    if (valid) cn.segments++;
    const float3 md = -st.d;
    int kind = PRT_MAT_NULL;
    float3 refl_rgb = mk3(0.0f, 0.0f, 0.0f);
    float p0 = 0.0f, p1 = 1.0f;
    if (valid) {
        const DMaterial &m = P.sc.mats[h.material];
        kind = __ldg(&m.kind);
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
                float3 sdir = sv * (1.0f / sd);
                cn.rays++;
                cn.shadow++;
                if (!occluded<true>(P.sc, prims, so, sdir, sd * (1.0f - 10.0f * PRT_RAY_EPSILON))) {
                    const DMaterial &me = P.sc.mats[__float_as_int(b.w)];
                    float f_cos = co * (1.0f / PRT_PI_F);
                    float w = mis_weight(pdf, f_cos) * f_cos / pdf;
                    st.res.x = fmaf(st.thr.x * refl_rgb.x * __ldg(&me.emission[0]), w, st.res.x);
                    st.res.y = fmaf(st.thr.y * refl_rgb.y * __ldg(&me.emission[1]), w, st.res.y);
                    st.res.z = fmaf(st.thr.z * refl_rgb.z * __ldg(&me.emission[2]), w, st.res.z);
                }
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

__device__ __forceinline__ void splat(const PtDev &P, float4 *tile, int tx0, int ty0, const PtState &st) {
    // ImageBlock::put, tent filter radius 1 (pixel centres at i + 0.5); tile origin (tx0 - 1, ty0 - 1)
    if (!P.tent) {
        int x = (int) floorf(st.px) - tx0 + 1, y = (int) floorf(st.py) - ty0 + 1;
        float *q = reinterpret_cast<float *>(tile + y * PT_HALO + x);
        atomicAdd(q, st.res.x); atomicAdd(q + 1, st.res.y); atomicAdd(q + 2, st.res.z); atomicAdd(q + 3, 1.0f);
        return;
    }
    int x0 = (int) floorf(st.px - 0.5f), y0 = (int) floorf(st.py - 0.5f);
#pragma unroll
    for (int dy = 0; dy < 2; dy++)
#pragma unroll
        for (int dx = 0; dx < 2; dx++) {
            int x = x0 + dx, y = y0 + dy;
            float wx = fmaxf(1.0f - fabsf((float) x + 0.5f - st.px), 0.0f), wy = fmaxf(1.0f - fabsf((float) y + 0.5f - st.py), 0.0f);
            float w = wx * wy;
            int lx = x - tx0 + 1, ly = y - ty0 + 1;
            if (w > 0.0f && lx >= 0 && ly >= 0 && lx < PT_HALO && ly < PT_HALO) {
                float *q = reinterpret_cast<float *>(tile + ly * PT_HALO + lx);
                atomicAdd(q, st.res.x * w); atomicAdd(q + 1, st.res.y * w); atomicAdd(q + 2, st.res.z * w); atomicAdd(q + 3, w);
            }
        }
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
            if (!live) {
                if (!inside || j >= P.n_s) break;
                pt_init(P, x, y, P.s_offset + j * P.s_stride, st);
                j++;
                cn.paths++;
                live = true;
            }
            live = pt_step(P, prims, st, cn);
            if (!live) splat(P, tile, tx0, ty0, st);
        }
        __syncthreads();
        // finished tile -> global film (halo pixels belong to neighbouring tiles as well: atomics)
        for (int i = threadIdx.x; i < PT_HALO * PT_HALO; i += blockDim.x) {
            int gx = tx0 - 1 + (i % PT_HALO), gy = ty0 - 1 + (i / PT_HALO);
            float4 v = tile[i];
            if (gx >= 0 && gy >= 0 && gx < P.W && gy < P.H && v.w != 0.0f) {
                float *q = P.film + 4 * ((size_t) gy * P.W + gx);
                atomicAdd(q, v.x); atomicAdd(q + 1, v.y); atomicAdd(q + 2, v.z); atomicAdd(q + 3, v.w);
            }
        }
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

static int fill_pt(prt_scene *s, const prt_render_params *p, uint64_t seed, uint32_t spp_total, uint32_t s_offset, uint32_t s_stride,
                   PtDev &P) {
    PRT_REQUIRE(p->width > 0 && p->height > 0 && p->max_depth >= 0 && spp_total > 0, "render_path: invalid render parameters");
    PRT_REQUIRE((uint64_t) p->width * p->height < (1ull << 31), "render_path: film too large");
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
    return PRT_OK;
}

static int launch_pt(prt_context *c, const PtDev &P, cudaStream_t st) {
    int per_sm = 0;
    PRT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_render_path, PT_THREADS, 0));
    if (per_sm < 1) per_sm = 1;
    int grid = c->sm_count * per_sm;
    int n_tiles = P.tiles_x * P.tiles_y;
    if (grid > n_tiles) grid = n_tiles;
    k_render_path<<<grid, PT_THREADS, 0, st>>>(P);
    PRT_CUDA(cudaGetLastError());
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

int prt_render_path(prt_scene *s, const prt_render_params *p, uint64_t seed, uint32_t spp_total, uint32_t sample_offset,
                    uint32_t sample_stride, float *film_rgbw, prt_render_stats *stats) {
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
    rc = ensure_scratch(c, n, 0, 0);
    if (rc) return rc;
    cudaEvent_t e0, e1, e2, e3;
    PRT_CUDA(cudaEventCreate(&e0)); PRT_CUDA(cudaEventCreate(&e1)); PRT_CUDA(cudaEventCreate(&e2)); PRT_CUDA(cudaEventCreate(&e3));
    PRT_CUDA(cudaEventRecord(e0, st));
    PRT_CUDA(cudaMemsetAsync(c->acc_dev, 0, sizeof(float) * n, st));
    PRT_CUDA(cudaMemsetAsync(c->stats_dev, 0, sizeof(uint64_t) * 8, st));
    P.film = c->acc_dev;
    P.stats = reinterpret_cast<unsigned long long *>(c->stats_dev);
    PRT_CUDA(cudaEventRecord(e1, st));
    rc = launch_pt(c, P, st);
    if (rc) return rc;
    PRT_CUDA(cudaEventRecord(e2, st));
    float *pin = reinterpret_cast<float *>(c->pinned);
    cudaPointerAttributes attr;
    const bool pinned_dst = cudaPointerGetAttributes(&attr, film_rgbw) == cudaSuccess && attr.type == cudaMemoryTypeHost;
    cudaGetLastError();
    if (pinned_dst) pin = film_rgbw;
    PRT_CUDA(cudaMemcpyAsync(pin, c->acc_dev, sizeof(float) * n, cudaMemcpyDeviceToHost, st));
    uint64_t hs[8];
    PRT_CUDA(cudaMemcpyAsync(hs, c->stats_dev, sizeof(uint64_t) * 8, cudaMemcpyDeviceToHost, st));
    PRT_CUDA(cudaEventRecord(e3, st));
    PRT_CUDA(cudaStreamSynchronize(st));
    if (!pinned_dst) memcpy(film_rgbw, pin, sizeof(float) * n);
    if (stats) {
        stats->paths = hs[0]; stats->segments = hs[1]; stats->rays = hs[2]; stats->shadow_rays = hs[3];
        PRT_CUDA(cudaEventElapsedTime(&stats->kernel_ms, e1, e2));
        PRT_CUDA(cudaEventElapsedTime(&stats->total_ms, e0, e3));
        stats->launches = 1;
        stats->_pad = 0;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaEventDestroy(e2); cudaEventDestroy(e3);
    return PRT_OK;
}

}  // extern "C"
