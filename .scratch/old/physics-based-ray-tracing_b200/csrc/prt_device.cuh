// prt_device.cuh -- device-side building blocks shared by the sm_100a kernels:
// vector math, PCG32 / sample_tea_32 streams, analytic-primitive and watertight-triangle tests,
// short-stack BVH2 traversal, Mitsuba-style surface interaction, and the UltraBSDF sampler.
//
// Reference call sites these replace (under /root/reference):
//   scene.ray_intersect            CustomIntegrator.py:146,159,309,324   -> closest_hit / occluded
//   si.spawn_ray / sh_frame / wi   CustomIntegrator.py:159,176,206       -> spawn_origin / Hit
//   UltraBSDF.sample               CustomBSDF.py:87-175                  -> ultra_bsdf_sample
// Mitsuba semantics per SURVEY.md Appendix C; the canonical path is SURVEY.md Appendix F.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define PRT_RAY_EPSILON 8.94069671630859375e-05f /* 1500 * 2^-24 */
#define PRT_PI_F 3.14159265358979323846f
#define PRT_INF __int_as_float(0x7f800000)

namespace prt {

// ------------------------------------------------------------------------------------------------
// device scene layout (all arrays 16-byte aligned, read through the read-only path)
// ------------------------------------------------------------------------------------------------
struct __align__(16) DPrim {  // 128 B
    float4 w0, w1, w2;        // to_world rows
    float4 o0, o1, o2;        // to_object rows
    float4 aux;               // sphere: centre xyz + radius ; planar: world normal xyz
    int kind, material, flip, shape;
};

struct __align__(16) DMaterial {  // 48 B
    int   kind;
    float p[7];
    float emission[3];
    float pad;
};

// BVH2 node, 64 B = 4 x float4 ("Node = 64 B"):
//   q0 = (L.lo.x, L.lo.y, L.lo.z, L.hi.x)  q1 = (L.hi.y, L.hi.z, R.lo.x, R.lo.y)
//   q2 = (R.lo.z, R.hi.x, R.hi.y, R.hi.z)  q3 = (bits left_ref, bits right_ref, -, -)
// child ref >= 0: internal node index; ref < 0: leaf, ~ref = (first_sorted_triangle << 2) | (count - 1)
struct DScene {
    const DPrim     *prims;
    const DMaterial *mats;
    const float4    *nodes;
    const float4    *tri_v;    // [n_tris][3] world-space vertices in BVH (sorted) order; .w of v0 = bits(orig index)
    const float4    *tri_n;    // [n_tris][3] world-space corner normals, sorted order (valid iff info.w & 1)
    const int4      *tri_info; // sorted order: {orig_index, shape, material, flags (1 = has normals, 2 = flip)}
    int n_prims, n_mats, n_tris, root_ref;
    // Triangles [0, n_small) (sorted order) are in the hierarchy.  [n_small, n_tris) are OVERSIZED triangles (bounding-box
    // area > 1024 x the scene's mean: the ten wall triangles of a closed box around ten million small ones) kept out of it
    // and tested one by one before every traversal: in a Morton-ordered tree such a triangle inflates the box of every
    // ancestor on its path to the full wall, and every ray near that wall then walks the whole chain.  (Embree handles the
    // same case with spatial splits; a handful of brute-force tests is the GPU-cheap equivalent.)  v1.w of such a triangle
    // = bits((sorted index << 2) | shading queue), like the BVH8 copies.
    int n_small;
    // area emitters (light transport only): every emissive mesh shape is one emitter
    const float4    *em_tri;         // [n_em_tris][3]: v0 + running area (w), v1 + bits(material) (w), v2 + flip (w)
    const int       *em_first;       // [n_emitters + 1] ranges into em_tri
    const float     *em_inv_area;    // [n_emitters]
    const int       *shape_emitter;  // [n_shapes] emitter index or -1
    int n_emitters, n_shapes;
    // compressed 8-wide BVH over the same triangles (prt_bvh8.cuh); n_nodes8 == 0: not built
    const float4    *nodes8;         // [n_nodes8][5]
    const float4    *tri_v8;         // [n_tris][3] vertices in BVH8 leaf order
    const uint32_t  *tri8_sorted;    // [n_tris] BVH8 triangle position -> sorted (LBVH) position
    int n_nodes8;
};

// ------------------------------------------------------------------------------------------------
// math
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float3 mk3(float x, float y, float z) { return make_float3(x, y, z); }
__device__ __forceinline__ float3 operator+(float3 a, float3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ float3 operator-(float3 a, float3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ float3 operator-(float3 a) { return mk3(-a.x, -a.y, -a.z); }
__device__ __forceinline__ float3 operator*(float3 a, float s) { return mk3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ float3 operator*(float s, float3 a) { return mk3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ float  dot(float3 a, float3 b) { return fmaf(a.x, b.x, fmaf(a.y, b.y, a.z * b.z)); }
__device__ __forceinline__ float3 cross(float3 a, float3 b) {
    return mk3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
__device__ __forceinline__ float3 normalize(float3 a) {
    float inv = 1.0f / sqrtf(dot(a, a));
    return a * inv;
}
__device__ __forceinline__ float3 xyz(float4 a) { return mk3(a.x, a.y, a.z); }

__device__ __forceinline__ float3 xpoint(float4 r0, float4 r1, float4 r2, float3 p) {
    return mk3(fmaf(r0.x, p.x, fmaf(r0.y, p.y, fmaf(r0.z, p.z, r0.w))),
               fmaf(r1.x, p.x, fmaf(r1.y, p.y, fmaf(r1.z, p.z, r1.w))),
               fmaf(r2.x, p.x, fmaf(r2.y, p.y, fmaf(r2.z, p.z, r2.w))));
}
__device__ __forceinline__ float3 xvec(float4 r0, float4 r1, float4 r2, float3 v) {
    return mk3(fmaf(r0.x, v.x, fmaf(r0.y, v.y, r0.z * v.z)), fmaf(r1.x, v.x, fmaf(r1.y, v.y, r1.z * v.z)),
               fmaf(r2.x, v.x, fmaf(r2.y, v.y, r2.z * v.z)));
}
// normal transform = multiply by the TRANSPOSE of the inverse rows
__device__ __forceinline__ float3 xnormal(float4 i0, float4 i1, float4 i2, float3 n) {
    return mk3(fmaf(i0.x, n.x, fmaf(i1.x, n.y, i2.x * n.z)), fmaf(i0.y, n.x, fmaf(i1.y, n.y, i2.y * n.z)),
               fmaf(i0.z, n.x, fmaf(i1.z, n.y, i2.z * n.z)));
}

// Mitsuba coordinate_system(): Duff et al. 2017
__device__ __forceinline__ void coordinate_system(float3 n, float3 &s, float3 &t) {
    float sign = copysignf(1.0f, n.z);
    float a = -1.0f / (sign + n.z);
    float b = n.x * n.y * a;
    s = mk3((n.x * n.x * a) * sign + 1.0f, b * sign, -n.x * sign);
    t = mk3(b, fmaf(n.y, n.y * a, sign), -n.y);
}

// ------------------------------------------------------------------------------------------------
// RNG: one PCG32 stream per path, seeded as Mitsuba's `independent` sampler seeds a wavefront
// (SURVEY.md C.6, 8(d) RNG contract)
// ------------------------------------------------------------------------------------------------
struct Pcg32 {
    uint64_t state, inc;
    __device__ __forceinline__ uint32_t next_u32() {
        uint64_t old = state;
        state = old * 0x5851f42d4c957f2dULL + inc;
        uint32_t xs = (uint32_t) (((old >> 18u) ^ old) >> 27u);
        uint32_t rot = (uint32_t) (old >> 59u);
        return __funnelshift_r(xs, xs, rot);
    }
    __device__ __forceinline__ float next_f32() { return __uint_as_float((next_u32() >> 9) | 0x3f800000u) - 1.0f; }
    __device__ __forceinline__ void seed(uint64_t initstate, uint64_t initseq) {
        state = 0;
        inc = (initseq << 1u) | 1u;
        next_u32();
        state += initstate;
        next_u32();
    }
};

__device__ __forceinline__ void sample_tea_32(uint32_t v0, uint32_t v1, uint32_t &o0, uint32_t &o1) {
    uint32_t sum = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        sum += 0x9e3779b9u;
        v0 += ((v1 << 4) + 0xa341316cu) ^ (v1 + sum) ^ ((v1 >> 5) + 0xc8013ea4u);
        v1 += ((v0 << 4) + 0xad90777du) ^ (v0 + sum) ^ ((v0 >> 5) + 0x7e95761eu);
    }
    o0 = v0;
    o1 = v1;
}

__device__ __forceinline__ Pcg32 path_rng(uint64_t seed, uint64_t path) {
    uint32_t v0, v1;
    sample_tea_32((uint32_t) seed + (uint32_t) (path >> 32), (uint32_t) path, v0, v1);
    Pcg32 r;
    r.seed((uint64_t) v0, (uint64_t) v1);
    return r;
}

// ------------------------------------------------------------------------------------------------
// hit record == the slice of Mitsuba's SurfaceInteraction3f the path reads
// ------------------------------------------------------------------------------------------------
struct Hit {
    float  t;
    float3 p, ng, ns, fs, ft;
    int    prim, shape, material;
};

// compact closest-hit candidate carried through traversal
struct Cand {
    float t;     // current best distance (== tmax while nothing was hit)
    int   id;    // >= 0: analytic primitive ; <= -2: sorted triangle ~id... ; -1: none
    float b1, b2;
};

__device__ __forceinline__ bool solve_quadratic(float a, float b, float c, float &x0, float &x1) {
    if (a == 0.0f) {
        if (b == 0.0f) return false;
        x0 = x1 = -c / b;
        return true;
    }
    // b^2 - 4ac with the rounding error of the product recovered by FMA (Kahan): the discriminant is then
    // accurate to ~1 ulp even near grazing incidence, where the naive form cancels
    float a4 = 4.0f * a;
    float p = a4 * c;
    float dp = fmaf(a4, c, -p);
    float discrim = fmaf(b, b, -p) - dp;
    if (!(discrim >= 0.0f)) return false;
    float temp = -0.5f * (b + copysignf(sqrtf(discrim), b));
    float x0p = temp / a, x1p = c / temp;
    x0 = fminf(x0p, x1p);
    x1 = fmaxf(x0p, x1p);
    return true;
}

// returns t >= 0 or -1 (cone/cylinder are builder-defined object-space quadrics)
__device__ __forceinline__ float intersect_prim(const DPrim &pr, float3 o, float3 d, float tmax) {
    if (pr.kind == 0) {  // sphere, world space
        // Same roots as Mitsuba's quadratic but evaluated in the cancellation-free form of
        // Haines et al. (Ray Tracing Gems ch. 7): the discriminant is r^2 - |oc - (oc.dn) dn|^2, so fp32 keeps
        // ~1e-6 relative accuracy in t down to |cos| ~ 1e-2 instead of losing it in B^2 - 4AC.
        float3 oc = o - xyz(pr.aux);
        float r = pr.aux.w;
        float inv_len = rsqrtf(dot(d, d));
        float3 dn = d * inv_len;
        float bp = -dot(oc, dn);
        float3 l = mk3(fmaf(bp, dn.x, oc.x), fmaf(bp, dn.y, oc.y), fmaf(bp, dn.z, oc.z));
        float disc = fmaf(r, r, -dot(l, l));
        if (!(disc >= 0.0f)) return -1.0f;
        float c = dot(oc, oc) - r * r;
        float q = bp + copysignf(sqrtf(disc), bp);
        float ta = c / q, tb = q;                  // roots for the normalised direction
        if (q == 0.0f) ta = tb = 0.0f;
        float n0 = fminf(ta, tb) * inv_len, n1 = fmaxf(ta, tb) * inv_len;
        if (!(n0 <= tmax && n1 >= 0.0f)) return -1.0f;
        if (n0 < 0.0f && n1 > tmax) return -1.0f;
        return n0 < 0.0f ? n1 : n0;
    }
    if (pr.kind == 1 || pr.kind == 3) {  // rectangle / disk
        // object-space z row first: most candidates are rejected on t alone, before the x / y rows are transformed
        // (same operations in the same order as xpoint / xvec, so t, lx, ly are bit-identical to the full transform)
        const float olz = fmaf(pr.o2.x, o.x, fmaf(pr.o2.y, o.y, fmaf(pr.o2.z, o.z, pr.o2.w)));
        const float dlz = fmaf(pr.o2.x, d.x, fmaf(pr.o2.y, d.y, pr.o2.z * d.z));
        float t = -olz / dlz;
        if (!(t >= 0.0f && t <= tmax)) return -1.0f;
        const float olx = fmaf(pr.o0.x, o.x, fmaf(pr.o0.y, o.y, fmaf(pr.o0.z, o.z, pr.o0.w)));
        const float oly = fmaf(pr.o1.x, o.x, fmaf(pr.o1.y, o.y, fmaf(pr.o1.z, o.z, pr.o1.w)));
        const float dlx = fmaf(pr.o0.x, d.x, fmaf(pr.o0.y, d.y, pr.o0.z * d.z));
        const float dly = fmaf(pr.o1.x, d.x, fmaf(pr.o1.y, d.y, pr.o1.z * d.z));
        float lx = fmaf(t, dlx, olx), ly = fmaf(t, dly, oly);
        bool in = pr.kind == 1 ? (fabsf(lx) <= 1.0f && fabsf(ly) <= 1.0f) : (lx * lx + ly * ly <= 1.0f);
        return in ? t : -1.0f;
    }
    float3 ol = xpoint(pr.o0, pr.o1, pr.o2, o), dl = xvec(pr.o0, pr.o1, pr.o2, d);
    float A, B, C;
    if (pr.kind == 2) {  // cone x^2+y^2 = (1-z)^2
        float w = 1.0f - ol.z;
        A = dl.x * dl.x + dl.y * dl.y - dl.z * dl.z;
        B = 2.0f * (ol.x * dl.x + ol.y * dl.y + w * dl.z);
        C = ol.x * ol.x + ol.y * ol.y - w * w;
    } else {  // cylinder x^2+y^2 = 1
        A = dl.x * dl.x + dl.y * dl.y;
        B = 2.0f * (ol.x * dl.x + ol.y * dl.y);
        C = ol.x * ol.x + ol.y * ol.y - 1.0f;
    }
    float n0, n1;
    if (!solve_quadratic(A, B, C, n0, n1)) return -1.0f;
    float z0 = fmaf(n0, dl.z, ol.z), z1 = fmaf(n1, dl.z, ol.z);
    if (n0 >= 0.0f && n0 <= tmax && z0 >= 0.0f && z0 <= 1.0f) return n0;
    if (n1 >= 0.0f && n1 <= tmax && z1 >= 0.0f && z1 <= 1.0f) return n1;
    return -1.0f;
}

__device__ __forceinline__ void finish_frame(Hit &h, float3 dp_du) {
    float3 s = dp_du - h.ns * dot(h.ns, dp_du);
    float l2 = dot(s, s);
    if (l2 > 0.0f) {
        h.fs = s * (1.0f / sqrtf(l2));
        h.ft = cross(h.ns, h.fs);
    } else {
        coordinate_system(h.ns, h.fs, h.ft);
    }
}

__device__ __forceinline__ void fill_prim_hit(const DPrim &pr, int index, float3 o, float3 d, float t, Hit &h) {
    h.t = t;
    h.prim = index;
    h.shape = pr.shape;
    h.material = pr.material;
    float3 pw = mk3(fmaf(t, d.x, o.x), fmaf(t, d.y, o.y), fmaf(t, d.z, o.z));
    float3 dp_du;
    if (pr.kind == 0) {
        float3 c = xyz(pr.aux);
        float3 n = normalize(pw - c);
        h.p = mk3(fmaf(n.x, pr.aux.w, c.x), fmaf(n.y, pr.aux.w, c.y), fmaf(n.z, pr.aux.w, c.z));
        float3 loc = xpoint(pr.o0, pr.o1, pr.o2, h.p);
        dp_du = xvec(pr.w0, pr.w1, pr.w2, mk3(-loc.y, loc.x, 0.0f));
        h.ng = n;
    } else if (pr.kind == 1 || pr.kind == 3) {
        float3 ol = xpoint(pr.o0, pr.o1, pr.o2, o), dl = xvec(pr.o0, pr.o1, pr.o2, d);
        float3 loc = mk3(fmaf(t, dl.x, ol.x), fmaf(t, dl.y, ol.y), 0.0f);
        h.p = xpoint(pr.w0, pr.w1, pr.w2, loc);
        h.ng = xyz(pr.aux);
        dp_du = mk3(pr.w0.x, pr.w1.x, pr.w2.x);  // to_world * (2,0,0), scale irrelevant after normalisation
    } else {
        float3 loc = xpoint(pr.o0, pr.o1, pr.o2, pw);
        float3 nl = pr.kind == 2 ? mk3(loc.x, loc.y, 1.0f - loc.z) : mk3(loc.x, loc.y, 0.0f);
        h.p = pw;
        h.ng = normalize(xnormal(pr.o0, pr.o1, pr.o2, nl));
        dp_du = xvec(pr.w0, pr.w1, pr.w2, mk3(-loc.y, loc.x, 0.0f));
    }
    if (pr.flip) h.ng = -h.ng;
    h.ns = h.ng;
    finish_frame(h, dp_du);
}

// ------------------------------------------------------------------------------------------------
// watertight ray/triangle test (Woop, Benthin, Wald 2013).  Products are kept un-contracted so that
// the edge functions of two triangles sharing an edge are evaluated on identical operands.
// ------------------------------------------------------------------------------------------------
struct RayPre {
    int   kx, ky, kz;
    float Sx, Sy, Sz;
};

__device__ __forceinline__ float comp(float3 v, int k) { return k == 0 ? v.x : (k == 1 ? v.y : v.z); }

__device__ __forceinline__ RayPre ray_precompute(float3 d) {
    RayPre r;
    float ax = fabsf(d.x), ay = fabsf(d.y), az = fabsf(d.z);
    r.kz = (ax > ay) ? (ax > az ? 0 : 2) : (ay > az ? 1 : 2);
    r.kx = r.kz == 2 ? 0 : r.kz + 1;
    r.ky = r.kx == 2 ? 0 : r.kx + 1;
    float dz = comp(d, r.kz);
    if (dz < 0.0f) {
        int tmp = r.kx;
        r.kx = r.ky;
        r.ky = tmp;
    }
    r.Sx = comp(d, r.kx) / dz;
    r.Sy = comp(d, r.ky) / dz;
    r.Sz = 1.0f / dz;
    return r;
}

// on a hit with t in [0, tbest]: updates tbest, b1 (weight of v1), b2 (weight of v2), returns true
__device__ __forceinline__ bool intersect_tri_wt(const RayPre &rp, float3 o, float3 v0, float3 v1, float3 v2, float &tbest,
                                                 float &b1, float &b2) {
    float3 A = v0 - o, B = v1 - o, C = v2 - o;
    float Akz = comp(A, rp.kz), Bkz = comp(B, rp.kz), Ckz = comp(C, rp.kz);
    // shear: a function of (vertex, ray) only, so a vertex shared by two triangles maps to the same point whether
    // or not the multiply-add is fused -- fusing it is safe; the EDGE functions below must stay un-fused
    float Ax = fmaf(-rp.Sx, Akz, comp(A, rp.kx)), Ay = fmaf(-rp.Sy, Akz, comp(A, rp.ky));
    float Bx = fmaf(-rp.Sx, Bkz, comp(B, rp.kx)), By = fmaf(-rp.Sy, Bkz, comp(B, rp.ky));
    float Cx = fmaf(-rp.Sx, Ckz, comp(C, rp.kx)), Cy = fmaf(-rp.Sy, Ckz, comp(C, rp.ky));
    float U = __fsub_rn(__fmul_rn(Cx, By), __fmul_rn(Cy, Bx));
    float V = __fsub_rn(__fmul_rn(Ax, Cy), __fmul_rn(Ay, Cx));
    float W = __fsub_rn(__fmul_rn(Bx, Ay), __fmul_rn(By, Ax));
    if (U == 0.0f || V == 0.0f || W == 0.0f) {  // exactly on an edge: decide in double
        double CxBy = (double) Cx * (double) By, CyBx = (double) Cy * (double) Bx;
        U = (float) (CxBy - CyBx);
        double AxCy = (double) Ax * (double) Cy, AyCx = (double) Ay * (double) Cx;
        V = (float) (AxCy - AyCx);
        double BxAy = (double) Bx * (double) Ay, ByAx = (double) By * (double) Ax;
        W = (float) (BxAy - ByAx);
    }
    if ((U < 0.0f || V < 0.0f || W < 0.0f) && (U > 0.0f || V > 0.0f || W > 0.0f)) return false;
    float det = U + V + W;
    if (det == 0.0f) return false;
    float Az = rp.Sz * Akz, Bz = rp.Sz * Bkz, Cz = rp.Sz * Ckz;
    float T = fmaf(U, Az, fmaf(V, Bz, W * Cz));
    float rdet = 1.0f / det;
    float t = T * rdet;
    if (!(t >= 0.0f && t <= tbest)) return false;
    tbest = t;
    b1 = V * rdet;
    b2 = W * rdet;
    return true;
}

// The same test with the permutation folded into three matrix rows per ray:
//   rx = e_kx - Sx e_kz,  ry = e_ky - Sy e_kz,  rz = Sz e_kz          (entries 1, 0 and the shear constant)
// so that (Ax, Ay, Az) = (A.rx, A.ry, A.rz) needs no run-time component selection.  comp() costs two compares and two
// selects per use; nine uses per triangle were 13-15 % of all instructions of the mesh kernels (profiles/r01, ncu joined
// with line info).  Products with 0 and 1 are exact, so each row dot differs from Woop's single fused multiply-add by at
// most one extra rounding -- and it is still a function of (vertex, ray) only: a vertex shared by two triangles maps to
// the same point, which is all watertightness needs.  The edge functions stay un-fused.
struct RayRows {
    float3 rx, ry, rz;
};

__device__ __forceinline__ RayRows ray_rows(const RayPre &p) {
    RayRows r;
    r.rx = mk3((p.kx == 0 ? 1.0f : 0.0f) - (p.kz == 0 ? p.Sx : 0.0f), (p.kx == 1 ? 1.0f : 0.0f) - (p.kz == 1 ? p.Sx : 0.0f),
               (p.kx == 2 ? 1.0f : 0.0f) - (p.kz == 2 ? p.Sx : 0.0f));
    r.ry = mk3((p.ky == 0 ? 1.0f : 0.0f) - (p.kz == 0 ? p.Sy : 0.0f), (p.ky == 1 ? 1.0f : 0.0f) - (p.kz == 1 ? p.Sy : 0.0f),
               (p.ky == 2 ? 1.0f : 0.0f) - (p.kz == 2 ? p.Sy : 0.0f));
    r.rz = mk3(p.kz == 0 ? p.Sz : 0.0f, p.kz == 1 ? p.Sz : 0.0f, p.kz == 2 ? p.Sz : 0.0f);
    return r;
}

__device__ __forceinline__ float row_dot(float3 v, float3 r) { return fmaf(v.x, r.x, fmaf(v.y, r.y, v.z * r.z)); }

__device__ __forceinline__ bool intersect_tri_rows(const RayRows &rr, float3 o, float3 v0, float3 v1, float3 v2, float &tbest,
                                                   float &b1, float &b2) {
    const float3 A = v0 - o, B = v1 - o, C = v2 - o;
    const float Ax = row_dot(A, rr.rx), Ay = row_dot(A, rr.ry);
    const float Bx = row_dot(B, rr.rx), By = row_dot(B, rr.ry);
    const float Cx = row_dot(C, rr.rx), Cy = row_dot(C, rr.ry);
    float U = __fsub_rn(__fmul_rn(Cx, By), __fmul_rn(Cy, Bx));
    float V = __fsub_rn(__fmul_rn(Ax, Cy), __fmul_rn(Ay, Cx));
    float W = __fsub_rn(__fmul_rn(Bx, Ay), __fmul_rn(By, Ax));
    if (U == 0.0f || V == 0.0f || W == 0.0f) {  // exactly on an edge: decide in double
        double CxBy = (double) Cx * (double) By, CyBx = (double) Cy * (double) Bx;
        U = (float) (CxBy - CyBx);
        double AxCy = (double) Ax * (double) Cy, AyCx = (double) Ay * (double) Cx;
        V = (float) (AxCy - AyCx);
        double BxAy = (double) Bx * (double) Ay, ByAx = (double) By * (double) Ax;
        W = (float) (BxAy - ByAx);
    }
    if ((U < 0.0f || V < 0.0f || W < 0.0f) && (U > 0.0f || V > 0.0f || W > 0.0f)) return false;
    const float det = U + V + W;
    if (det == 0.0f) return false;
    const float Az = row_dot(A, rr.rz), Bz = row_dot(B, rr.rz), Cz = row_dot(C, rr.rz);
    const float T = fmaf(U, Az, fmaf(V, Bz, W * Cz));
    const float rdet = 1.0f / det;
    const float t = T * rdet;
    if (!(t >= 0.0f && t <= tbest)) return false;
    tbest = t;
    b1 = V * rdet;
    b2 = W * rdet;
    return true;
}

#ifndef PRT_TRI_ROWS
#define PRT_TRI_ROWS 1      // 1: intersect_tri_rows in every traversal; 0: the component-selecting intersect_tri_wt (A/B)
#endif
#ifndef PRT_ROWS_LATE
#define PRT_ROWS_LATE 0     // binary traversal: 1 rebuilds the rows at every leaf, 0 keeps them in registers across the node loop
#endif

__device__ __forceinline__ float4 ldg4(const float4 *p) { return __ldg(p); }

// slab test against one child box; returns entry distance or +inf on a miss
__device__ __forceinline__ float box_entry(float lox, float loy, float loz, float hix, float hiy, float hiz, float3 o,
                                           float3 inv, float tmax) {
    float t0x = (lox - o.x) * inv.x, t1x = (hix - o.x) * inv.x;
    float t0y = (loy - o.y) * inv.y, t1y = (hiy - o.y) * inv.y;
    float t0z = (loz - o.z) * inv.z, t1z = (hiz - o.z) * inv.z;
    // fminf/fmaxf drop NaNs (0 * inf when the origin lies exactly on a slab of a parallel ray)
    float tn = fmaxf(fmaxf(fminf(t0x, t1x), fminf(t0y, t1y)), fmaxf(fminf(t0z, t1z), 0.0f));
    float tf = fminf(fminf(fmaxf(t0x, t1x), fmaxf(t0y, t1y)), fminf(fmaxf(t0z, t1z), tmax));
    return tn <= tf * 1.0000004f ? tn : PRT_INF;
}

#define PRT_STACK 48

// the oversized triangles of the scene (DScene::n_small), nearest (or any) within [0, tbest]; returns the sorted index or -1
template <bool ANY, typename RAYT>
__device__ __forceinline__ int test_big_tris(const DScene &sc, const RAYT &rt, float3 o, float &tbest, float &b1, float &b2) {
    int best = -1;
    for (int j = sc.n_small; j < sc.n_tris; j++) {
        const float4 *tv = sc.tri_v + 3 * (size_t) j;
        const float4 a = ldg4(tv), b = ldg4(tv + 1), c = ldg4(tv + 2);
#if PRT_TRI_ROWS
        if (intersect_tri_rows(rt, o, xyz(a), xyz(b), xyz(c), tbest, b1, b2)) {
#else
        if (intersect_tri_wt(rt, o, xyz(a), xyz(b), xyz(c), tbest, b1, b2)) {
#endif
            best = j;
            if (ANY) return best;
        }
    }
    return best;
}

// nearest triangle (ANY = false) or any triangle (ANY = true) along the ray within [0, tbest].
// returns the SORTED triangle index or -1; tbest/b1/b2 updated on a hit.
template <bool ANY>
__device__ __forceinline__ int traverse_bvh(const DScene &sc, float3 o, float3 d, float &tbest, float &b1, float &b2) {
    if (sc.n_tris == 0) return -1;
    const RayPre rp = ray_precompute(d);
#if PRT_TRI_ROWS && !PRT_ROWS_LATE
    const RayRows rr = ray_rows(rp);
#endif
    const float3 inv = mk3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
    int   stack_ref[PRT_STACK];
    float stack_t[PRT_STACK];
    int sp = 0, best = -1;
    if (sc.n_small < sc.n_tris) {
#if PRT_TRI_ROWS
        best = test_big_tris<ANY>(sc, ray_rows(rp), o, tbest, b1, b2);
#else
        best = test_big_tris<ANY>(sc, rp, o, tbest, b1, b2);
#endif
        if ((ANY && best >= 0) || sc.n_small == 0) return best;
    }
    int ref = sc.root_ref;
    const int DONE = 0x7fffffff;
    // pop the next entry whose box still starts before the current best hit
#define PRT_POP()                                                        \
    do {                                                                 \
        ref = DONE;                                                      \
        while (sp > 0) {                                                 \
            --sp;                                                        \
            if (stack_t[sp] <= tbest) { ref = stack_ref[sp]; break; }    \
        }                                                                \
    } while (0)
    // "while-while" traversal (Aila & Laine 2009): all lanes first descend inner nodes until each holds a
    // leaf (or is done), then the warp tests triangles together -- node and leaf code never interleave
    for (;;) {
        while ((unsigned) ref < (unsigned) DONE) {
            const float4 *n = sc.nodes + 4 * (size_t) ref;
            float4 q0 = ldg4(n), q1 = ldg4(n + 1), q2 = ldg4(n + 2), q3 = ldg4(n + 3);
            float tl = box_entry(q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, o, inv, tbest);
            float tr = box_entry(q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, o, inv, tbest);
            int rl = __float_as_int(q3.x), rr = __float_as_int(q3.y);
            bool hl = tl < PRT_INF, hr = tr < PRT_INF;
            if (hl && hr) {
                bool lf = tl <= tr;
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
        if (ref == DONE) return best;
        {
            int code = ~ref;
            int first = code >> 2, count = (code & 3) + 1;
#if PRT_TRI_ROWS && PRT_ROWS_LATE
            const RayRows rr = ray_rows(rp);     // rebuilt per leaf instead of living through the node loop
#endif
            for (int j = 0; j < count; j++) {
                const float4 *tv = sc.tri_v + 3 * (size_t) (first + j);
                float4 a = ldg4(tv), b = ldg4(tv + 1), c = ldg4(tv + 2);
#if PRT_TRI_ROWS
                if (intersect_tri_rows(rr, o, xyz(a), xyz(b), xyz(c), tbest, b1, b2)) {
#else
                if (intersect_tri_wt(rp, o, xyz(a), xyz(b), xyz(c), tbest, b1, b2)) {
#endif
                    best = first + j;
                    if (ANY) return best;
                }
            }
        }
        PRT_POP();
    }
#undef PRT_POP
}

__device__ __forceinline__ void fill_tri_hit(const DScene &sc, int sorted_tri, float t, float b1, float b2, Hit &h) {
    const float4 *tv = sc.tri_v + 3 * (size_t) sorted_tri;
    float3 p0 = xyz(ldg4(tv)), p1 = xyz(ldg4(tv + 1)), p2 = xyz(ldg4(tv + 2));
    int4 info = __ldg(sc.tri_info + sorted_tri);
    float b0 = 1.0f - b1 - b2;
    h.t = t;
    h.prim = sc.n_prims + info.x;
    h.shape = info.y;
    h.material = info.z;
    h.p = p0 * b0 + p1 * b1 + p2 * b2;
    h.ng = normalize(cross(p1 - p0, p2 - p0));
    if (info.w & 1) {
        const float4 *tn = sc.tri_n + 3 * (size_t) sorted_tri;
        float3 n0 = xyz(ldg4(tn)), n1 = xyz(ldg4(tn + 1)), n2 = xyz(ldg4(tn + 2));
        h.ns = normalize(n0 * b0 + n1 * b1 + n2 * b2);
    } else {
        h.ns = h.ng;
    }
    if (info.w & 2) {
        h.ng = -h.ng;
        h.ns = -h.ns;
    }
    float3 du, dv;
    coordinate_system(h.ng, du, dv);
    finish_frame(h, du);
}

// si.spawn_ray(d) origin
__device__ __forceinline__ float3 spawn_origin(float3 p, float3 ng, float3 d) {
    float m = fmaxf(fabsf(p.x), fmaxf(fabsf(p.y), fabsf(p.z)));
    float mag = copysignf((1.0f + m) * PRT_RAY_EPSILON, dot(ng, d));
    return mk3(fmaf(ng.x, mag, p.x), fmaf(ng.y, mag, p.y), fmaf(ng.z, mag, p.z));
}

// directivity_weight_i (CustomIntegrator.py:120-135 / 289-304): alpha = |acos(dot(n_T, -sec))|; 1 for alpha <= alpha_m, linear
// ramp to 0 at alpha_c, else 0.  acos is monotone, so the two plateaus are decided on the cosine and acosf only runs on the
// ramp (rare: the aperture subtends a few degrees; the ramp is continuous at both ends, so an ulp-level tie is immaterial).
__device__ __forceinline__ float directivity_wi(float3 nT, float3 sec, float cos_m, float cos_c, float alpha_m, float alpha_c) {
    const float cdt = dot(nT, -sec);
    float w_i = cdt >= cos_m ? 1.0f : 0.0f;
    if (cdt < cos_m && cdt >= cos_c) {
        const float al = fabsf(acosf(cdt));
        w_i = al <= alpha_m ? 1.0f : (al <= alpha_c ? (alpha_c - al) / (alpha_c - alpha_m) : 0.0f);
    }
    return w_i;
}

// UltraBSDF.sample (CustomBSDF.py:87-175, _ggx_sample :30-61, ggx_pdf == 1 :81-82), SURVEY.md Appendix F.
// All quirks Q4-Q9 are kept: scalar sample on the disk diagonal, mixed local/world frames, m flipped
// against wi, "reflection" = wi + 2 (wi.m) m, local components returned as the new world direction.
__device__ __forceinline__ void ultra_bsdf_sample(float3 wi, float3 ng, float3 ns, float Z, float alpha, float s1, float s2,
                                                  float3 &dir, float &pdf, float &amp, bool &reflect) {
    float3 fs, ft;
    coordinate_system(ng, fs, ft);                                   // CB:32
    float3 w = mk3(dot(wi, fs), dot(wi, ft), dot(wi, ng));           // CB:33
    float3 ws = normalize(mk3(alpha * w.x, alpha * w.y, w.z));       // CB:37-38
    float inv = rsqrtf(fmaxf(fmaf(-ws.z, ws.z, 1.0f), 1e-7f));       // CB:41
    float3 T1 = mk3(ws.y * inv, -ws.x * inv, 0.0f);                  // CB:42-44
    float3 T2 = cross(ws, T1);                                       // CB:45
    float r = fmaf(2.0f, s1, -1.0f);                                 // CB:48 concentric disk of (s1, s1): phi = pi/4
    float qx = r * 0.70710678118654752f, qy = qx;
    float S = 0.5f * (1.0f + ws.z);                                  // CB:51
    qy = fmaf(1.0f - S, sqrtf(fmaxf(fmaf(-qx, qx, 1.0f), 0.0f)), S * qy);   // CB:52
    float zz = sqrtf(fmaxf(1.0f - qx * qx - qy * qy, 0.0f));         // CB:55
    float3 ms = T1 * qx + T2 * qy + ws * zz;
    float3 m = normalize(mk3(alpha * ms.x, alpha * ms.y, ms.z));     // CB:56-59
    if (!(dot(m, wi) < 0.0f)) m = -m;                                // CB:100
    float cwm = dot(wi, m);                                          // CB:101
    const float Z1 = Z, Z2 = 1.2f;                                   // CB:104-107 (entering is always False)
    float ratio = Z1 / Z2;                                           // CB:111
    float cTr = fabsf(cwm);                                          // CB:119
    float sq = 1.0f - (ratio * ratio) * (1.0f - cTr * cTr);          // CB:120
    float cTt = sqrtf(fmaxf(sq, 0.0f));                              // CB:121
    float Ar = (Z1 * cTr - Z2 * cTt) / (Z1 * cTr + Z2 * cTt);        // CB:122-123
    float At = 1.0f - Ar;                                            // CB:124
    float3 refl = wi + m * (2.0f * cwm);                             // CB:130
    float3 trans = refl * ratio + m * (ratio * cTr - cTt);           // CB:131
    reflect = (sq < 0.0f) || (s2 < Ar * Ar);                         // CB:137-145
    if (reflect) {
        dir = refl;
        pdf = 1.0f / (4.0f * fabsf(cwm));                            // CB:153-154
        amp = Ar;
    } else {
        float anwi = fabsf(dot(ns, wi));                             // CB:156
        float anwo = fmaxf(fabsf(dot(ns, trans)), 1e-7f);            // CB:157
        dir = trans;
        pdf = (ratio * ratio) * fabsf(dot(trans, m)) / (anwi * anwo);  // CB:158
        amp = At;
    }
}

// ------------------------------------------------------------------------------------------------
// light-transport helpers shared by the megakernel (prt_path.cu) and the wavefront kernels (prt_wavefront.cu):
// Mitsuba's mis_weight, warp::square_to_uniform_disk_concentric and fresnel()
// ------------------------------------------------------------------------------------------------
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

}  // namespace prt
