/*
 * oracle/orc.c -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).  See orc.h for scope and the
 * "parity unpinned" statement.  Scene container, CPU BVH (median split, independent of the GPU LBVH),
 * PCG32 / sample_tea_32, and the two precision instantiations of orc_impl.inl.
 */
#define _GNU_SOURCE
#include "orc.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>
#include <pthread.h>

static inline void orc_atomic_add(double *addr, double v) {
    union { double d; uint64_t u; } old, neu;
    uint64_t *a = (uint64_t *) addr;
    old.u = __atomic_load_n(a, __ATOMIC_RELAXED);
    do { neu.d = old.d + v; } while (!__atomic_compare_exchange_n(a, &old.u, neu.u, 1, __ATOMIC_RELAXED, __ATOMIC_RELAXED));
}

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif
#define ORC_RAY_EPSILON (1500.0 * 5.9604644775390625e-08) /* C.3: RayEpsilon = 1500 * 2^-24 */

typedef struct {
    int    kind, material, flip, shape;
    double to_world[12], to_object[12];
    double radius; /* sphere: |to_world * (1,0,0)| */
} orc_prim;

typedef struct {
    int    kind;
    double p[8];
    double emission[3];
} orc_material;

typedef struct {
    double lo[3], hi[3];
    int    left, right, first, count; /* count > 0 => leaf over tri_order[first .. first+count) */
} orc_node;

struct orc_scene {
    orc_prim     *prims;      int n_prims, cap_prims;
    orc_material *materials;  int n_materials, cap_materials;
    double       *tri_v;      /* [n_tris][3][3] world space */
    double       *tri_n;      /* [n_tris][3][3] world-space corner normals (if tri_has_n) */
    int          *tri_shape, *tri_material;
    unsigned char *tri_has_n, *tri_flip;
    int           n_tris, cap_tris;
    int           n_shapes;
    orc_node     *nodes;      int n_nodes, cap_nodes;
    int          *tri_order;
    int           use_bvh;
    /* area emitters: every emissive mesh shape is one emitter (Mitsuba: uniform pick, then area-weighted face) */
    int           n_emitters;
    int          *shape_emitter;     /* [n_shapes] emitter index or -1 */
    double       *emitter_inv_area;  /* [n_emitters] */
    int          *emitter_first;     /* [n_emitters + 1] range into em_tri / em_cdf */
    int          *em_tri;            /* triangle ids */
    double       *em_cdf;            /* running area within the emitter */
};

/* ---- RNG: PCG32 + sample_tea_32 exactly as Mitsuba's `independent` sampler seeds a wavefront (C.6) ---- */
#define PCG32_MULT 0x5851f42d4c957f2dULL

void orc_sample_tea_32(uint32_t v0, uint32_t v1, int rounds, uint32_t *o0, uint32_t *o1) {
    uint32_t sum = 0;
    for (int i = 0; i < rounds; i++) {
        sum += 0x9e3779b9u;
        v0 += ((v1 << 4) + 0xa341316cu) ^ (v1 + sum) ^ ((v1 >> 5) + 0xc8013ea4u);
        v1 += ((v0 << 4) + 0xad90777du) ^ (v0 + sum) ^ ((v0 >> 5) + 0x7e95761eu);
    }
    *o0 = v0; *o1 = v1;
}

uint32_t orc_pcg32_next_u32(uint64_t *state, uint64_t inc) {
    uint64_t old = *state;
    *state = old * PCG32_MULT + inc;
    uint32_t xs = (uint32_t) (((old >> 18u) ^ old) >> 27u);
    uint32_t rot = (uint32_t) (old >> 59u);
    return (xs >> rot) | (xs << ((~rot + 1u) & 31));
}

void orc_pcg32_seed(uint64_t initstate, uint64_t initseq, uint64_t *state, uint64_t *inc) {
    *state = 0;
    *inc = (initseq << 1u) | 1u;
    orc_pcg32_next_u32(state, *inc);
    *state += initstate;
    orc_pcg32_next_u32(state, *inc);
}

float orc_pcg32_next_f32(uint64_t *state, uint64_t inc) {
    union { uint32_t u; float f; } x;
    x.u = (orc_pcg32_next_u32(state, inc) >> 9) | 0x3f800000u;
    return x.f - 1.0f;
}

/* RNG contract: (v0, v1) = sample_tea_32(seed + hi32(path), lo32(path)), 4 rounds;
 * pcg32.seed(initstate = v0, initseq = v1).  For path < 2^32 this is Mitsuba's sampler.seed(seed, wavefront). */
void orc_path_rng(uint64_t seed, uint64_t path_index, uint64_t *state, uint64_t *inc) {
    uint32_t v0, v1;
    orc_sample_tea_32((uint32_t) seed + (uint32_t) (path_index >> 32), (uint32_t) path_index, 4, &v0, &v1);
    orc_pcg32_seed((uint64_t) v0, (uint64_t) v1, state, inc);
}

/* ---- scene container ---- */
static int invert_affine(const double m[16], double inv12[12]) {
    /* inverse of the upper-left 3x3 + translation (bottom row assumed 0 0 0 1) */
    double a = m[0], b = m[1], c = m[2], d = m[4], e = m[5], f = m[6], g = m[8], h = m[9], i = m[10];
    double det = a * (e * i - f * h) - b * (d * i - f * g) + c * (d * h - e * g);
    if (det == 0.0) return -1;
    double id = 1.0 / det;
    double r[9] = { (e * i - f * h) * id, (c * h - b * i) * id, (b * f - c * e) * id,
                    (f * g - d * i) * id, (a * i - c * g) * id, (c * d - a * f) * id,
                    (d * h - e * g) * id, (b * g - a * h) * id, (a * e - b * d) * id };
    double tx = m[3], ty = m[7], tz = m[11];
    for (int k = 0; k < 3; k++) {
        inv12[4 * k + 0] = r[3 * k + 0]; inv12[4 * k + 1] = r[3 * k + 1]; inv12[4 * k + 2] = r[3 * k + 2];
        inv12[4 * k + 3] = -(r[3 * k + 0] * tx + r[3 * k + 1] * ty + r[3 * k + 2] * tz);
    }
    return 0;
}

orc_scene *orc_scene_create(void) { return (orc_scene *) calloc(1, sizeof(orc_scene)); }

void orc_scene_destroy(orc_scene *s) {
    if (!s) return;
    free(s->prims); free(s->materials); free(s->tri_v); free(s->tri_n); free(s->tri_shape);
    free(s->tri_material); free(s->tri_has_n); free(s->tri_flip); free(s->nodes); free(s->tri_order);
    free(s->shape_emitter); free(s->emitter_inv_area); free(s->emitter_first); free(s->em_tri); free(s->em_cdf);
    free(s);
}

int orc_scene_add_material(orc_scene *s, int kind, const double p[8], const double emission_rgb[3]) {
    if (s->n_materials == s->cap_materials) {
        s->cap_materials = s->cap_materials ? 2 * s->cap_materials : 8;
        s->materials = (orc_material *) realloc(s->materials, sizeof(orc_material) * s->cap_materials);
    }
    orc_material *m = &s->materials[s->n_materials];
    m->kind = kind;
    for (int i = 0; i < 8; i++) m->p[i] = p ? p[i] : 0.0;
    for (int i = 0; i < 3; i++) m->emission[i] = emission_rgb ? emission_rgb[i] : 0.0;
    return s->n_materials++;
}

int orc_scene_set_material_param(orc_scene *s, int material, int index, double value) {
    if (material < 0 || material >= s->n_materials || index < 0 || index >= 8) return -1;
    s->materials[material].p[index] = value;
    return 0;
}

int orc_scene_add_prim(orc_scene *s, int kind, const double to_world[16], int material, int flip) {
    if (kind < ORC_SPHERE || kind > ORC_CYLINDER || material < 0 || material >= s->n_materials) return -1;
    if (s->n_prims == s->cap_prims) {
        s->cap_prims = s->cap_prims ? 2 * s->cap_prims : 8;
        s->prims = (orc_prim *) realloc(s->prims, sizeof(orc_prim) * s->cap_prims);
    }
    orc_prim *p = &s->prims[s->n_prims];
    p->kind = kind; p->material = material; p->flip = flip; p->shape = s->n_shapes;
    memcpy(p->to_world, to_world, sizeof(double) * 12);
    if (invert_affine(to_world, p->to_object)) return -2;
    p->radius = sqrt(to_world[0] * to_world[0] + to_world[4] * to_world[4] + to_world[8] * to_world[8]);
    s->n_prims++;
    return s->n_shapes++;
}

int orc_scene_add_mesh(orc_scene *s, const double *v, uint32_t nv, const double *vn, const uint32_t *idx, uint32_t nt,
                       const double to_world[16], int material, int flip) {
    if (material < 0 || material >= s->n_materials) return -1;
    double inv[12];
    if (invert_affine(to_world, inv)) return -2;
    if (s->n_tris + (int) nt > s->cap_tris) {
        s->cap_tris = s->n_tris + (int) nt;
        s->tri_v = (double *) realloc(s->tri_v, sizeof(double) * 9 * s->cap_tris);
        s->tri_n = (double *) realloc(s->tri_n, sizeof(double) * 9 * s->cap_tris);
        s->tri_shape = (int *) realloc(s->tri_shape, sizeof(int) * s->cap_tris);
        s->tri_material = (int *) realloc(s->tri_material, sizeof(int) * s->cap_tris);
        s->tri_has_n = (unsigned char *) realloc(s->tri_has_n, s->cap_tris);
        s->tri_flip = (unsigned char *) realloc(s->tri_flip, s->cap_tris);
    }
    for (uint32_t t = 0; t < nt; t++) {
        int o = s->n_tris + (int) t;
        for (int c = 0; c < 3; c++) {
            uint32_t vi = idx[3 * t + c];
            if (vi >= nv) return -3;
            const double *p = v + 3 * (size_t) vi;
            for (int r = 0; r < 3; r++)
                s->tri_v[9 * (size_t) o + 3 * c + r] =
                    to_world[4 * r] * p[0] + to_world[4 * r + 1] * p[1] + to_world[4 * r + 2] * p[2] + to_world[4 * r + 3];
            if (vn) {
                const double *n = vn + 3 * (size_t) vi;
                double w[3];
                for (int r = 0; r < 3; r++) w[r] = inv[r] * n[0] + inv[4 + r] * n[1] + inv[8 + r] * n[2];
                double l = sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
                for (int r = 0; r < 3; r++) s->tri_n[9 * (size_t) o + 3 * c + r] = l > 0 ? w[r] / l : 0.0;
            }
        }
        s->tri_shape[o] = s->n_shapes;
        s->tri_material[o] = material;
        s->tri_has_n[o] = vn != NULL;
        s->tri_flip[o] = (unsigned char) flip;
    }
    s->n_tris += (int) nt;
    return s->n_shapes++;
}

int orc_scene_counts(const orc_scene *s, int32_t *n_prims, int32_t *n_tris, int32_t *n_shapes) {
    if (n_prims) *n_prims = s->n_prims;
    if (n_tris) *n_tris = s->n_tris;
    if (n_shapes) *n_shapes = s->n_shapes;
    return 0;
}

/* ---- CPU BVH: top-down, split at the median of the widest centroid axis, leaves of <= 4 triangles ---- */
static const double *g_sort_cent;
static int g_sort_axis;
static int cmp_cent(const void *a, const void *b) {
    double ca = g_sort_cent[3 * (size_t) (*(const int *) a) + g_sort_axis];
    double cb = g_sort_cent[3 * (size_t) (*(const int *) b) + g_sort_axis];
    return (ca > cb) - (ca < cb);
}

static int build_node(orc_scene *s, const double *cent, int first, int count) {
    if (s->n_nodes == s->cap_nodes) {
        s->cap_nodes = s->cap_nodes ? 2 * s->cap_nodes : 64;
        s->nodes = (orc_node *) realloc(s->nodes, sizeof(orc_node) * s->cap_nodes);
    }
    int me = s->n_nodes++;
    double lo[3] = { DBL_MAX, DBL_MAX, DBL_MAX }, hi[3] = { -DBL_MAX, -DBL_MAX, -DBL_MAX };
    double clo[3] = { DBL_MAX, DBL_MAX, DBL_MAX }, chi[3] = { -DBL_MAX, -DBL_MAX, -DBL_MAX };
    for (int j = 0; j < count; j++) {
        int t = s->tri_order[first + j];
        for (int c = 0; c < 3; c++)
            for (int a = 0; a < 3; a++) {
                double x = s->tri_v[9 * (size_t) t + 3 * c + a];
                if (x < lo[a]) lo[a] = x;
                if (x > hi[a]) hi[a] = x;
            }
        for (int a = 0; a < 3; a++) {
            double x = cent[3 * (size_t) t + a];
            if (x < clo[a]) clo[a] = x;
            if (x > chi[a]) chi[a] = x;
        }
    }
    for (int a = 0; a < 3; a++) {
        double pad = 1e-6 * (fabs(lo[a]) + fabs(hi[a]) + (hi[a] - lo[a])) + 1e-12;
        s->nodes[me].lo[a] = lo[a] - pad;
        s->nodes[me].hi[a] = hi[a] + pad;
    }
    if (count <= 4) {
        s->nodes[me].first = first; s->nodes[me].count = count; s->nodes[me].left = s->nodes[me].right = -1;
        return me;
    }
    int axis = 0;
    if (chi[1] - clo[1] > chi[axis] - clo[axis]) axis = 1;
    if (chi[2] - clo[2] > chi[axis] - clo[axis]) axis = 2;
    g_sort_cent = cent; g_sort_axis = axis;
    qsort(s->tri_order + first, (size_t) count, sizeof(int), cmp_cent);
    int half = count / 2;
    int l = build_node(s, cent, first, half);
    int r = build_node(s, cent, first + half, count - half);
    s->nodes[me].left = l; s->nodes[me].right = r; s->nodes[me].first = 0; s->nodes[me].count = 0;
    return me;
}

static void build_emitters(orc_scene *s) {
    free(s->shape_emitter); free(s->emitter_inv_area); free(s->emitter_first); free(s->em_tri); free(s->em_cdf);
    s->shape_emitter = (int *) malloc(sizeof(int) * (size_t) (s->n_shapes + 1));
    s->emitter_inv_area = (double *) calloc((size_t) s->n_shapes + 1, sizeof(double));
    s->emitter_first = (int *) calloc((size_t) s->n_shapes + 2, sizeof(int));
    s->em_tri = (int *) malloc(sizeof(int) * (size_t) (s->n_tris + 1));
    s->em_cdf = (double *) malloc(sizeof(double) * (size_t) (s->n_tris + 1));
    s->n_emitters = 0;
    for (int i = 0; i < s->n_shapes; i++) s->shape_emitter[i] = -1;
    int n = 0;
    for (int t = 0; t < s->n_tris; t++) {
        const orc_material *m = &s->materials[s->tri_material[t]];
        if (!(m->emission[0] > 0 || m->emission[1] > 0 || m->emission[2] > 0)) continue;
        int sh = s->tri_shape[t];
        if (s->shape_emitter[sh] < 0) {          /* triangles of a shape are contiguous */
            s->shape_emitter[sh] = s->n_emitters;
            s->emitter_first[s->n_emitters] = n;
            s->n_emitters++;
        }
        const double *v = s->tri_v + 9 * (size_t) t;
        double e0[3] = { v[3] - v[0], v[4] - v[1], v[5] - v[2] }, e1[3] = { v[6] - v[0], v[7] - v[1], v[8] - v[2] };
        double cx = e0[1] * e1[2] - e0[2] * e1[1], cy = e0[2] * e1[0] - e0[0] * e1[2], cz = e0[0] * e1[1] - e0[1] * e1[0];
        double area = 0.5 * sqrt(cx * cx + cy * cy + cz * cz);
        int first = s->emitter_first[s->n_emitters - 1];
        s->em_tri[n] = t;
        s->em_cdf[n] = (n > first ? s->em_cdf[n - 1] : 0.0) + area;
        n++;
        s->emitter_first[s->n_emitters] = n;
    }
    for (int e = 0; e < s->n_emitters; e++) s->emitter_inv_area[e] = 1.0 / s->em_cdf[s->emitter_first[e + 1] - 1];
}

int orc_scene_commit(orc_scene *s, int use_bvh) {
    free(s->nodes); s->nodes = NULL; s->n_nodes = s->cap_nodes = 0;
    free(s->tri_order); s->tri_order = NULL;
    s->use_bvh = use_bvh;
    build_emitters(s);
    if (!use_bvh || s->n_tris == 0) return 0;
    double *cent = (double *) malloc(sizeof(double) * 3 * (size_t) s->n_tris);
    s->tri_order = (int *) malloc(sizeof(int) * (size_t) s->n_tris);
    for (int t = 0; t < s->n_tris; t++) {
        s->tri_order[t] = t;
        for (int a = 0; a < 3; a++)
            cent[3 * (size_t) t + a] =
                (s->tri_v[9 * (size_t) t + a] + s->tri_v[9 * (size_t) t + 3 + a] + s->tri_v[9 * (size_t) t + 6 + a]) / 3.0;
    }
    build_node(s, cent, 0, s->n_tris);
    free(cent);
    return 0;
}

/* ---- two precision instantiations ---- */
#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)

#define R float
#define FN(n) CAT(n, _f32)
#define RSQRT_FN sqrtf
#define RSIN_FN sinf
#define RCOS_FN cosf
#define REXP_FN expf
#define RACOS_FN acosf
#define RFLOOR_FN floorf
#define RRINT_FN rintf
#define COPYSIGN_FN copysignf
#define RINF ((float) INFINITY)
#define REPS FLT_EPSILON
#include "orc_impl.inl"
#undef R
#undef FN
#undef RSQRT_FN
#undef RSIN_FN
#undef RCOS_FN
#undef REXP_FN
#undef RACOS_FN
#undef RFLOOR_FN
#undef RRINT_FN
#undef COPYSIGN_FN
#undef RINF
#undef REPS

#define R double
#define FN(n) CAT(n, _f64)
#define RSQRT_FN sqrt
#define RSIN_FN sin
#define RCOS_FN cos
#define REXP_FN exp
#define RACOS_FN acos
#define RFLOOR_FN floor
#define RRINT_FN rint
#define COPYSIGN_FN copysign
#define RINF ((double) INFINITY)
#define REPS DBL_EPSILON
#include "orc_impl.inl"
#undef R
#undef FN

/* ---- public wrappers ---- */
#define WRAP_CLOSEST(SUF, RT)                                                                                      \
    for (uint64_t i = 0; i < n; i++) {                                                                             \
        v3##SUF oo = { (RT) o[3 * i], (RT) o[3 * i + 1], (RT) o[3 * i + 2] };                                      \
        v3##SUF dd = { (RT) d[3 * i], (RT) d[3 * i + 1], (RT) d[3 * i + 2] };                                      \
        RT tm = tmax ? (RT) tmax[i] : (RT) INFINITY;                                                               \
        hit##SUF h;                                                                                                \
        int ok = closest##SUF(sc, oo, dd, tm, &h, NULL);                                                           \
        if (t) t[i] = ok ? (double) h.t : INFINITY;                                                                \
        if (prim) prim[i] = ok ? h.prim : -1;                                                                      \
        if (shape) shape[i] = ok ? h.shape : -1;                                                                   \
        if (ok) {                                                                                                  \
            v3##SUF md = { -dd.x, -dd.y, -dd.z };                                                                  \
            if (p)  { p[3 * i] = h.p.x; p[3 * i + 1] = h.p.y; p[3 * i + 2] = h.p.z; }                              \
            if (ng) { ng[3 * i] = h.ng.x; ng[3 * i + 1] = h.ng.y; ng[3 * i + 2] = h.ng.z; }                        \
            if (ns) { ns[3 * i] = h.ns.x; ns[3 * i + 1] = h.ns.y; ns[3 * i + 2] = h.ns.z; }                        \
            if (wi) { wi[3 * i] = dot##SUF(md, h.fs); wi[3 * i + 1] = dot##SUF(md, h.ft);                          \
                      wi[3 * i + 2] = dot##SUF(md, h.ns); }                                                        \
            if (fs) { fs[3 * i] = h.fs.x; fs[3 * i + 1] = h.fs.y; fs[3 * i + 2] = h.fs.z; }                        \
        }                                                                                                          \
    }

int orc_trace_closest_frame(const orc_scene *sc, int prec, const double *o, const double *d, const double *tmax, uint64_t n,
                            double *t, int32_t *prim, int32_t *shape, double *p, double *ng, double *ns, double *wi,
                            double *fs) {
    if (prec == 32) { WRAP_CLOSEST(_f32, float) }
    else if (prec == 64) { WRAP_CLOSEST(_f64, double) }
    else return -1;
    return 0;
}

int orc_trace_closest(const orc_scene *sc, int prec, const double *o, const double *d, const double *tmax, uint64_t n,
                      double *t, int32_t *prim, int32_t *shape, double *p, double *ng, double *ns, double *wi) {
    return orc_trace_closest_frame(sc, prec, o, d, tmax, n, t, prim, shape, p, ng, ns, wi, NULL);
}

int orc_trace_occluded(const orc_scene *sc, int prec, const double *o, const double *d, const double *tmax, uint64_t n,
                       uint8_t *hit) {
    for (uint64_t i = 0; i < n; i++) {
        if (prec == 32) {
            v3_f32 oo = { (float) o[3 * i], (float) o[3 * i + 1], (float) o[3 * i + 2] };
            v3_f32 dd = { (float) d[3 * i], (float) d[3 * i + 1], (float) d[3 * i + 2] };
            hit[i] = (uint8_t) occluded_f32(sc, oo, dd, tmax ? (float) tmax[i] : (float) INFINITY, NULL);
        } else {
            v3_f64 oo = { o[3 * i], o[3 * i + 1], o[3 * i + 2] };
            v3_f64 dd = { d[3 * i], d[3 * i + 1], d[3 * i + 2] };
            hit[i] = (uint8_t) occluded_f64(sc, oo, dd, tmax ? tmax[i] : INFINITY, NULL);
        }
    }
    return 0;
}

int orc_ultra_bsdf_n(int prec, uint64_t n, const double *wi, const double *ng, const double *ns, const double *impedance,
                     const double *roughness, const double *s1, const double *s2, double *dir, double *pdf, double *amp,
                     int32_t *reflect) {
    for (uint64_t i = 0; i < n; i++) {
        int rc = orc_ultra_bsdf(prec, wi + 3 * i, ng + 3 * i, ns + 3 * i, impedance[i], roughness[i], s1[i], s2[i], dir + 3 * i,
                                pdf + i, amp + i, reflect + i);
        if (rc) return rc;
    }
    return 0;
}

int orc_directivity(int prec, const double sensor_to_world[16], const double sec_dir[3], const double ray_dir[3],
                    const double normal[3], double main_beam_deg, double cutoff_deg, double num_rays, double *w_i, double *w_o) {
    orc_acq_params p;
    memset(&p, 0, sizeof p);
    memcpy(p.sensor_to_world, sensor_to_world, sizeof p.sensor_to_world);
    p.main_beam_deg = main_beam_deg; p.cutoff_deg = cutoff_deg; p.n_angles = 1; p.n_elements = 1;
    if (prec == 32) {
        acq_f32 q; acq_setup_f32(&p, &q);
        v3_f32 s = { (float) sec_dir[0], (float) sec_dir[1], (float) sec_dir[2] };
        v3_f32 d = { (float) ray_dir[0], (float) ray_dir[1], (float) ray_dir[2] }, n = { (float) normal[0], (float) normal[1], (float) normal[2] };
        *w_i = directivity_wi_f32(q.nT, s, q.alpha_m, q.alpha_c);
        *w_o = dot_f32(d, n) / (float) num_rays;                     /* CI:117-118 */
    } else if (prec == 64) {
        acq_f64 q; acq_setup_f64(&p, &q);
        v3_f64 s = { sec_dir[0], sec_dir[1], sec_dir[2] }, d = { ray_dir[0], ray_dir[1], ray_dir[2] }, n = { normal[0], normal[1], normal[2] };
        *w_i = directivity_wi_f64(q.nT, s, q.alpha_m, q.alpha_c);
        *w_o = dot_f64(d, n) / num_rays;
    } else return -1;
    return 0;
}

int orc_directivity_n(int prec, uint64_t n, const double *sensor_to_world /*[n][16]*/, const double *sec_dir, const double *ray_dir,
                      const double *normal, const double *main_beam_deg, const double *cutoff_deg, const double *num_rays,
                      double *w_i, double *w_o) {
    for (uint64_t i = 0; i < n; i++) {
        int rc = orc_directivity(prec, sensor_to_world + 16 * i, sec_dir + 3 * i, ray_dir + 3 * i, normal + 3 * i, main_beam_deg[i],
                                 cutoff_deg[i], num_rays[i], w_i + i, w_o + i);
        if (rc) return rc;
    }
    return 0;
}

int orc_ultra_bsdf(int prec, const double wi[3], const double ng[3], const double ns[3], double impedance,
                   double roughness, double s1, double s2, double dir[3], double *pdf, double *amp, int32_t *reflect) {
    int rf;
    if (prec == 32) {
        v3_f32 a = { (float) wi[0], (float) wi[1], (float) wi[2] }, b = { (float) ng[0], (float) ng[1], (float) ng[2] },
               c = { (float) ns[0], (float) ns[1], (float) ns[2] }, o;
        float pf, am;
        ultra_bsdf_f32(a, b, c, (float) impedance, (float) roughness, (float) s1, (float) s2, &o, &pf, &am, &rf);
        dir[0] = o.x; dir[1] = o.y; dir[2] = o.z; *pdf = pf; *amp = am;
    } else {
        v3_f64 a = { wi[0], wi[1], wi[2] }, b = { ng[0], ng[1], ng[2] }, c = { ns[0], ns[1], ns[2] }, o;
        ultra_bsdf_f64(a, b, c, impedance, roughness, s1, s2, &o, pdf, amp, &rf);
        dir[0] = o.x; dir[1] = o.y; dir[2] = o.z;
    }
    *reflect = rf;
    return 0;
}

int orc_acquire(const orc_scene *sc, int prec, const orc_acq_params *p, uint64_t seed, uint32_t spp_total,
                uint32_t s_offset, uint32_t s_stride, double *buf, double *tx, orc_stats *stats, int n_threads) {
    if (prec == 32) return acquire_f32(sc, p, seed, spp_total, s_offset, s_stride, buf, tx, stats, n_threads);
    if (prec == 64) return acquire_f64(sc, p, seed, spp_total, s_offset, s_stride, buf, tx, stats, n_threads);
    return -1;
}

int orc_acquire_trace(const orc_scene *sc, int prec, const orc_acq_params *p, uint64_t seed, uint32_t spp_total,
                      const uint64_t *path_idx, uint64_t n, orc_seg_record *rec) {
    memset(rec, 0, sizeof(orc_seg_record) * n * (size_t) p->max_depth);
    for (uint64_t i = 0; i < n; i++) {
        uint64_t ae = path_idx[i] / spp_total;
        uint32_t s = (uint32_t) (path_idx[i] % spp_total);
        int a = (int) (ae / (uint64_t) p->n_elements), e = (int) (ae % (uint64_t) p->n_elements);
        if (a >= p->n_angles) return -2;
        if (prec == 32) {
            acq_f32 q; acq_setup_f32(p, &q);
            acq_path_f32(sc, &q, p, seed, spp_total, a, e, s, NULL, rec + i * (size_t) p->max_depth, NULL);
        } else {
            acq_f64 q; acq_setup_f64(p, &q);
            acq_path_f64(sc, &q, p, seed, spp_total, a, e, s, NULL, rec + i * (size_t) p->max_depth, NULL);
        }
    }
    return 0;
}

int orc_render_path(const orc_scene *sc, int prec, const orc_render_params *p, uint64_t seed, uint32_t spp_total,
                    uint32_t s_offset, uint32_t s_stride, double *film, orc_stats *stats, uint64_t *shadow_rays, int n_threads) {
    if (prec == 32) return render_path_f32(sc, p, seed, spp_total, s_offset, s_stride, film, stats, shadow_rays, n_threads);
    if (prec == 64) return render_path_f64(sc, p, seed, spp_total, s_offset, s_stride, film, stats, shadow_rays, n_threads);
    return -1;
}
