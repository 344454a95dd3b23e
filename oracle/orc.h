/*
 * oracle/orc.h -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE)
 *
 * Plain-C restatement of the reference's ray-traced acquisition loop
 *   /root/reference/CustomIntegrator.py:60-232   (UltraIntegrator.simulate_acquisition, "D")
 *   /root/reference/CustomIntegrator.py:235-405  (UltraIntegrator.simulate_acquisition_parallel, "P")
 *   /root/reference/CustomBSDF.py:30-61,87-175   (UltraBSDF._ggx_sample / sample)
 * together with the Mitsuba-3 semantics those lines lean on (shape intersection,
 * SurfaceInteraction3f, Frame3f, spawn_ray, PCG32 / sample_tea_32).  Mitsuba 3 and
 * Dr.Jit are un-vendored, un-pinned PyPI dependencies of the reference (no
 * requirements file; .idea/misc.xml + pyc mtimes suggest the 3.6.x era) and are not
 * installed here, so their published algorithms are restated from the Mitsuba 3
 * sources as remembered.
 *
 * PINNING.  The reference ships no tests, golden vectors or fixtures, and Mitsuba cannot be installed here.  The
 * oracle is pinned against
 *   (1) fixtures produced by the reference's OWN Python: tests/golden/make_ref_fixtures.py executes
 *       /root/reference/CustomBSDF.py and CustomIntegrator.py unmodified (on the repository's mitsuba / drjit stand-ins,
 *       ray queries served by this oracle's intersector, uniforms injected from the per-path PCG32 streams) and records
 *       12 000 UltraBSDF.sample calls, 4 000 directivity weights and 32 000 per-segment records of
 *       simulate_acquisition / simulate_acquisition_parallel on ten scenes (tests/golden/ref_*.npz;
 *       tests/test_ref_fixtures.py: every decision of every record reproduced exactly in binary32);
 *   (2) the analytic anchors of SURVEY.md section 8(c) (tests/golden/anchors.json, exact / f64 arithmetic);
 *   (3) an independent pure-Python transliteration (oracle/pyref.py) and an external PCG32 known-answer vector.
 * STILL UNPINNED ("[MEM]"): what the reference obtains from the mitsuba wheel itself -- shape intersection routines,
 * si.spawn_ray's offset, Frame3f / sh_frame / si.wi conventions, square_to_uniform_disk_concentric, TEA seeding, and the
 * whole `path` integrator of orc_pt.inl (no reference code exists for it).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library.  The product (libprt_b200.so) never links or calls it.
 *
 * Every entry point takes a precision switch `prec`: 32 evaluates the path in IEEE
 * binary32 (what Mitsuba's llvm_ad_mono `Float` is), 64 in binary64 (ground truth for
 * the analytic anchors).  Array arguments are always double on the ABI.
 */
#ifndef ORC_H
#define ORC_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* analytic primitive kinds (object-space definitions, Mitsuba conventions) */
enum {
    ORC_SPHERE    = 0, /* unit sphere, centre 0                       (mitsuba `sphere`)    */
    ORC_RECTANGLE = 1, /* [-1,1]^2 in z = 0, normal +z                (mitsuba `rectangle`) */
    ORC_CONE      = 2, /* x^2+y^2 = (1-z)^2, z in [0,1], open base    (builder-defined; MitsubaScenes/Cone_*.xml) */
    ORC_DISK      = 3, /* unit disk in z = 0, normal +z               (mitsuba `disk`)      */
    ORC_CYLINDER  = 4  /* x^2+y^2 = 1, z in [0,1], open               (mitsuba `cylinder` p0=(0,0,0), p1=(0,0,1), r=1) */
};

/* material kinds */
enum {
    ORC_MAT_ULTRA      = 0, /* p[0]=impedance p[1]=roughness      CustomBSDF.py:8-26 */
    ORC_MAT_DIFFUSE    = 1, /* p[0..2]=reflectance rgb                               */
    ORC_MAT_DIELECTRIC = 2, /* p[0]=int_ior p[1]=ext_ior                             */
    ORC_MAT_CONDUCTOR  = 3, /* perfect mirror (material=none); p[0..2]=specular rgb  */
    ORC_MAT_NULL       = 4
};

/* quirk switches.  0 = canonical path (Appendix F). */
enum {
    ORC_QF_CLAMP_TIDX        = 1u << 0, /* A.1#3 D: clamp t_idx to [0,T-1] instead of dropping  CustomIntegrator.py:192 */
    ORC_QF_TOF_LAST_SEGMENT  = 1u << 1, /* A.1#2 D: tof never advances                          CustomIntegrator.py:165,226 */
    ORC_QF_SINGLE_BOUNCE     = 1u << 2, /* A.1#8 P as written: one bounce per ray               CustomIntegrator.py:365-376 */
    ORC_QF_RR_NO_ABS         = 1u << 3, /* A.1#6 D: rr = min(atten*amp, 1) without abs          CustomIntegrator.py:220 */
    ORC_QF_CONNECT_TO_TARGET = 1u << 4  /* not the reference: connection ray stops at the receive element (fixes Q1) */
};

typedef struct {
    int32_t  n_angles, n_elements, time_samples, max_depth;
    double   pitch, fs, sound_speed, frequency, attenuation;
    double   main_beam_deg, cutoff_deg, max_path_len;
    double   sensor_to_world[16]; /* row-major 4x4 */
    uint32_t quirk_flags;
    uint32_t _pad;
    const double *angles_deg;     /* [n_angles] */
} orc_acq_params;

typedef struct {
    uint64_t paths, segments, rays, deposits, misses;
    uint64_t nodes_visited, tris_tested; /* BVH instrumentation (closest + any) */
} orc_stats;

/* one record per executed segment of one path (decision-level parity) */
typedef struct {
    int32_t  valid;      /* 1 if the segment executed (a hit was found) */
    int32_t  prim;       /* global primitive id of the hit */
    int32_t  shape;      /* shape index of the hit */
    int32_t  recv;       /* receive element */
    int32_t  visible;    /* connection un-occluded */
    int32_t  reflect;    /* 1 = reflected, 0 = transmitted */
    int32_t  k;          /* time bin (before range test) */
    int32_t  survive;    /* RR outcome */
    double   t;          /* hit distance */
    double   total_time; /* Ttot */
    double   press;      /* deposit value (before /spp) */
    double   amp, atten; /* state after the segment */
    double   dir[3];     /* new direction */
} orc_seg_record;

typedef struct orc_scene orc_scene;

orc_scene *orc_scene_create(void);
void       orc_scene_destroy(orc_scene *);
/* returns material id */
int orc_scene_add_material(orc_scene *, int kind, const double p[8], const double emission_rgb[3]);
/* returns shape id; to_world row-major 4x4 */
int orc_scene_add_prim(orc_scene *, int kind, const double to_world[16], int material, int flip_normals);
/* v [nv][3] object space, vn [nv][3] or NULL, idx [nt][3]; returns shape id */
int orc_scene_add_mesh(orc_scene *, const double *v, uint32_t nv, const double *vn, const uint32_t *idx,
                       uint32_t nt, const double to_world[16], int material, int flip_normals);
int orc_scene_set_material_param(orc_scene *, int material, int index, double value);
/* builds the CPU BVH over all triangles; use_bvh = 0 keeps brute force */
int orc_scene_commit(orc_scene *, int use_bvh);
int orc_scene_counts(const orc_scene *, int32_t *n_prims, int32_t *n_tris, int32_t *n_shapes);

/* == scene.ray_intersect (CustomIntegrator.py:146,309): nearest hit, tmax given per ray (inf allowed).
 * outputs may be NULL.  prim = -1, t = inf on miss. */
int orc_trace_closest(const orc_scene *, int prec, const double *o, const double *d, const double *tmax, uint64_t n,
                      double *t, int32_t *prim, int32_t *shape, double *p, double *ng, double *ns, double *wi);
/* same, plus the shading-frame tangent fs (si.sh_frame.s; t = n x s) the fixture harness needs to rebuild
 * SurfaceInteraction3f for the reference's own Python (tests/golden/make_ref_fixtures.py) */
int orc_trace_closest_frame(const orc_scene *, int prec, const double *o, const double *d, const double *tmax, uint64_t n,
                            double *t, int32_t *prim, int32_t *shape, double *p, double *ng, double *ns, double *wi,
                            double *fs);
int orc_trace_occluded(const orc_scene *, int prec, const double *o, const double *d, const double *tmax, uint64_t n,
                       uint8_t *hit);

/* == UltraBSDF.sample (CustomBSDF.py:87-175) on explicit inputs.  out: dir[3], pdf, amp, reflect */
int orc_ultra_bsdf(int prec, const double wi[3], const double ng[3], const double ns[3], double impedance,
                   double roughness, double s1, double s2, double dir[3], double *pdf, double *amp, int32_t *reflect);

/* == directivity_weight_i / directivity_weight_o (CustomIntegrator.py:114-135 / 286-304) on explicit inputs */
int orc_directivity(int prec, const double sensor_to_world[16], const double sec_dir[3], const double ray_dir[3],
                    const double normal[3], double main_beam_deg, double cutoff_deg, double num_rays, double *w_i, double *w_o);

/* batched forms of the two (arrays of n) */
int orc_ultra_bsdf_n(int prec, uint64_t n, const double *wi, const double *ng, const double *ns, const double *impedance,
                     const double *roughness, const double *s1, const double *s2, double *dir, double *pdf, double *amp,
                     int32_t *reflect);
int orc_directivity_n(int prec, uint64_t n, const double *sensor_to_world /*[n][16]*/, const double *sec_dir, const double *ray_dir,
                      const double *normal, const double *main_beam_deg, const double *cutoff_deg, const double *num_rays,
                      double *w_i, double *w_o);

/* == simulate_acquisition{,_parallel}.  Samples s = s_offset + j*s_stride < spp_total are traced for
 * every (angle, element); path index i = (a*n_e + e)*spp_total + s seeds PCG32 via sample_tea_32.
 * buf [n_a][n_e][T] is ACCUMULATED into (caller zeroes); tx_delays [n_a][n_e] overwritten. */
int orc_acquire(const orc_scene *, int prec, const orc_acq_params *, uint64_t seed, uint32_t spp_total,
                uint32_t s_offset, uint32_t s_stride, double *buf, double *tx_delays, orc_stats *stats,
                int n_threads);
/* decision trace of selected paths: rec [n][max_depth] */
int orc_acquire_trace(const orc_scene *, int prec, const orc_acq_params *, uint64_t seed, uint32_t spp_total,
                      const uint64_t *path_idx, uint64_t n, orc_seg_record *rec);

/* == mi.render(scene) with the `path` integrator (scenes/cbox.xml:5-32); SURVEY.md Appendix C.7 */
typedef struct {
    double  to_world[16];
    double  fov_deg;            /* along the smaller image axis */
    double  near_clip, far_clip;
    int32_t width, height, max_depth, rr_depth;
    int32_t rfilter;            /* 0 box, 1 tent (radius 1) */
    int32_t _pad;
} orc_render_params;
/* film [H][W][4] = (sum w R, sum w G, sum w B, sum w), ACCUMULATED into (caller zeroes) */
int orc_render_path(const orc_scene *, int prec, const orc_render_params *, uint64_t seed, uint32_t spp_total,
                    uint32_t s_offset, uint32_t s_stride, double *film, orc_stats *stats, uint64_t *shadow_rays,
                    int n_threads);

/* RNG contract */
void     orc_sample_tea_32(uint32_t v0, uint32_t v1, int rounds, uint32_t *o0, uint32_t *o1);
void     orc_pcg32_seed(uint64_t initstate, uint64_t initseq, uint64_t *state, uint64_t *inc);
uint32_t orc_pcg32_next_u32(uint64_t *state, uint64_t inc);
float    orc_pcg32_next_f32(uint64_t *state, uint64_t inc);
void     orc_path_rng(uint64_t seed, uint64_t path_index, uint64_t *state, uint64_t *inc);

#ifdef __cplusplus
}
#endif
#endif
