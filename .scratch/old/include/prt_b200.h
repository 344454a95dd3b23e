/*
 * include/prt_b200.h -- C ABI of libprt_b200.so, the B200 (sm_100a) engine behind the reference's
 * Python plugin surface.  Plain pointers and sizes only; no torch / C++ types.
 *
 * The reference (ReaganCardoza/Physics-Based-Ray-Tracing) has no native interface of its own: its hot
 * path calls Mitsuba 3's Python API.  Each entry point below names the reference call site it replaces
 * (file:line under /root/reference).  INTEGRATION.md shows the ctypes stubs a maintainer adds.
 *
 * Conventions
 *   - every function returns 0 on success, a negative prt_status otherwise; prt_last_error() gives text
 *   - arrays without a `_dev` suffix are caller-owned contiguous HOST buffers; `_dev` are device pointers
 *     on the context's device (e.g. torch tensors' data_ptr()), `stream` is a cudaStream_t (0 = default)
 *   - matrices are row-major 4x4 float64 (converted to fp32 on upload, as Mitsuba's llvm_ad_* `Float`)
 *   - one context per (process, device); calls on one context are serialised by an internal mutex
 *   - there is NO CPU fallback: without a CUDA device every compute call fails with PRT_ERR_CUDA
 */
#ifndef PRT_B200_H
#define PRT_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    PRT_OK = 0,
    PRT_ERR_INVALID = -1,   /* bad argument / handle                       */
    PRT_ERR_CUDA = -2,      /* CUDA runtime error (see prt_last_error)     */
    PRT_ERR_STATE = -3,     /* e.g. tracing an uncommitted scene           */
    PRT_ERR_UNSUPPORTED = -4
} prt_status;

/* analytic primitives, object-space definitions (Mitsuba conventions) */
enum { PRT_SPHERE = 0, PRT_RECTANGLE = 1, PRT_CONE = 2, PRT_DISK = 3, PRT_CYLINDER = 4 };
/* materials */
enum { PRT_MAT_ULTRA = 0, PRT_MAT_DIFFUSE = 1, PRT_MAT_DIELECTRIC = 2, PRT_MAT_CONDUCTOR = 3, PRT_MAT_NULL = 4 };
/* quirk switches of the acquisition loop; 0 = canonical (Appendix F) */
enum {
    PRT_QF_CLAMP_TIDX = 1u << 0,        /* CustomIntegrator.py:192   clamp instead of drop            */
    PRT_QF_TOF_LAST_SEGMENT = 1u << 1,  /* CustomIntegrator.py:165   tof never advances                */
    PRT_QF_SINGLE_BOUNCE = 1u << 2,     /* CustomIntegrator.py:365-376 as written: one bounce per ray  */
    PRT_QF_RR_NO_ABS = 1u << 3,         /* CustomIntegrator.py:220   rr prob without abs               */
    PRT_QF_CONNECT_TO_TARGET = 1u << 4  /* not the reference: connection ray ends at the element      */
};

typedef struct prt_context prt_context;
typedef struct prt_scene prt_scene;

/* ---- context ------------------------------------------------------------------------------- */
const char *prt_last_error(void);                 /* thread-local message of the last failure    */
const char *prt_version(void);
int prt_device_count(int *count);
/* replaces mi.set_variant("cuda_ad_mono") (TestScene.py:3): binds a CUDA device */
int prt_create(int device, prt_context **out);
int prt_destroy(prt_context *);
/* Per-kernel-class device timing for bench.py's roofline leg: between prt_profile_begin and prt_profile_read every
 * kernel group the library enqueues is bracketed by a CUDA event pair ON THE STREAM IT IS LAUNCHED ON
 * (`launches` counts kernels; the shading kernels of one bounce share one event pair).
 * prt_profile_read waits for the recorded work, sums the elapsed times per class and switches profiling off. */
enum { PRT_KC_GENERATE = 0, PRT_KC_TRACE_CLOSEST = 1, PRT_KC_TRACE_SHADOW = 2, PRT_KC_SHADE = 3, PRT_KC_FILM = 4,
       PRT_KC_ACQUIRE = 5, PRT_KC_MEGAKERNEL = 6, PRT_KC_OTHER = 7, PRT_KC_COUNT = 8 };
typedef struct {
    double   ms[8];
    uint32_t launches[8];
} prt_kernel_times;
int prt_profile_begin(prt_context *);
int prt_profile_read(prt_context *, prt_kernel_times *out);
int prt_device_info(prt_context *, int *sm_count, int *cc_major, int *cc_minor, uint64_t *global_mem_bytes);
/* Page-locked host memory for result buffers.  prt_acquire / prt_render_path detect a pinned destination and
 * copy device -> host straight into it (per-angle slices overlapped with the kernel of the next angle);
 * pageable destinations go through an internal pinned staging buffer + memcpy. */
int prt_host_alloc(prt_context *, uint64_t bytes, void **out);
int prt_host_free(prt_context *, void *ptr);

/* ---- scene upload (replaces mi.load_dict / mi.load_file, USMain.py:257; Mitsuba builds Embree here) */
int prt_scene_create(prt_context *, prt_scene **out);
int prt_scene_destroy(prt_scene *);
/* kind = PRT_MAT_*; p[8]: ULTRA {impedance, roughness} (CustomBSDF.py:12-18), DIFFUSE {r,g,b},
 * DIELECTRIC {int_ior, ext_ior}, CONDUCTOR {r,g,b}; emission = area-emitter radiance or NULL */
int prt_scene_add_material(prt_scene *, int kind, const double p[8], const double emission_rgb[3], int *material_id);
/* mi.traverse(scene)[...] = v ; params.update()  (USMain.py:259,264-265): no rebuild */
int prt_scene_set_material_param(prt_scene *, int material_id, int index, double value);
/* params['<shape>.to_world'] = T ; params.update(): moves one shape of a committed scene WITHOUT rebuilding it.  Analytic
 * primitive: its 128-byte record is replaced.  Mesh: the shape's triangles are moved on the device (new * old^-1), the
 * binary tree's boxes are refitted bottom-up over its unchanged topology and the 8-wide BVH is derived again -- no Morton
 * codes, no sort, no hierarchy emission (the reference reaches the same point through
 * mi.traverse(scene) / params.update(), USMain.py:259,264-265, where Mitsuba rebuilds its Embree scene).  Scenes of more
 * than 2^22 triangles keep no topology and are rebuilt. */
int prt_scene_set_shape_transform(prt_scene *, int shape_id, const double to_world[16]);
int prt_scene_add_primitive(prt_scene *, int kind, const double to_world[16], int material_id, int flip_normals,
                            int *shape_id);
/* v [nv][3], vn [nv][3] or NULL, idx [nt][3] (object space) */
int prt_scene_add_mesh(prt_scene *, const double *v, uint32_t nv, const double *vn, const uint32_t *idx, uint32_t nt,
                       const double to_world[16], int material_id, int flip_normals, int *shape_id);

typedef struct {
    uint32_t n_primitives, n_triangles, n_nodes, max_leaf_size;
    float    build_ms;          /* Morton + sort + hierarchy + refit, CUDA-event timed */
    float    sah_cost;          /* SAH cost of the final tree (root area normalised)   */
    float    scene_lo[3], scene_hi[3];
    uint64_t device_bytes;      /* bytes of device memory held by the scene            */
    uint32_t n_nodes8;          /* nodes of the compressed 8-wide BVH (80 B each), 0 = none */
    uint32_t bvh8_levels;
    float    bvh8_build_ms;     /* collapse + quantisation + triangle reorder, CUDA-event timed */
    uint32_t n_oversized;       /* triangles kept out of the hierarchy and tested one by one (bounding-box area > 1024 x the
                                   scene's mean, at most 64): n_nodes == n_triangles - n_oversized - 1 */
} prt_bvh_stats;
/* SoA upload + GPU LBVH build (Morton codes -> radix sort -> Karras hierarchy -> bottom-up refit) */
int prt_scene_commit(prt_scene *, prt_bvh_stats *out /* nullable */);
/* the statistics prt_scene_commit reported, as they stand now (a transform update refreshes build_ms / sah_cost / bounds) */
int prt_scene_get_stats(prt_scene *, prt_bvh_stats *out);

/* ---- ray queries == scene.ray_intersect(ray) (CustomIntegrator.py:146,159,309,324) ------------ */
/* o,d [n][3] f32, tmax [n] or NULL (= inf).  Outputs nullable.  prim = -1, t = inf on a miss.
 * prim: analytic primitives first [0, n_primitives), then n_primitives + triangle index (input order) */
int prt_trace_closest(prt_scene *, const float *o, const float *d, const float *tmax, uint64_t n,
                      float *t, int32_t *prim, int32_t *shape, float *p /*[n][3]*/, float *ng, float *ns, float *wi,
                      float *sh_s /*[n][3] shading-frame tangent s; t = n x s*/);
int prt_trace_occluded(prt_scene *, const float *o, const float *d, const float *tmax, uint64_t n, uint8_t *hit);
/* == UltraBSDF.sample(ctx, si, sample1, sample2) (CustomBSDF.py:87-175) on n explicit interactions */
int prt_ultra_bsdf_sample(prt_context *, uint64_t n, const float *wi, const float *ng, const float *ns,
                          const float *impedance, const float *roughness, const float *s1, const float *s2,
                          float *dir /*[n][3]*/, float *pdf, float *amp, int32_t *reflect);
/* == directivity_weight_i(sec_dir, alpha_m, alpha_c) and directivity_weight_o(ray_dir, n, num_rays), the two nested
 * functions of CustomIntegrator.py:114-135 / 286-304, on n explicit inputs (parity tests against tests/golden/
 * ref_directivity.npz, which holds the outputs of the reference's own bytecode) */
int prt_directivity_weights(prt_context *, uint64_t n, const double sensor_to_world[16], const float *sec_dir /*[n][3]*/,
                            const float *ray_dir /*[n][3]*/, const float *normal /*[n][3]*/, double main_beam_deg,
                            double cutoff_deg, double num_rays, float *w_i, float *w_o);

/* ---- acquisition == UltraIntegrator.simulate_acquisition{,_parallel}(scene) --------------------
 * (CustomIntegrator.py:60-232, 235-405).  Property names / defaults: CustomIntegrator.py:16-42. */
typedef struct {
    int32_t  n_angles, n_elements, time_samples, max_depth;
    double   pitch, fs, sound_speed, frequency, attenuation;
    double   main_beam_deg, cutoff_deg, max_path_len;   /* max_path_len = 0.2 (CustomIntegrator.py:141) */
    double   sensor_to_world[16];                        /* UltraSensor.transform                         */
    uint32_t quirk_flags;
    uint32_t _pad;
    const double *angles_deg;                            /* [n_angles], host                              */
} prt_acq_params;

typedef struct {
    uint64_t paths, segments, rays, deposits, misses;
    float    kernel_ms;     /* device time of the path kernel (CUDA events)   */
    float    total_ms;      /* including zero-fill and copies                 */
    uint32_t launches;      /* kernels launched by this call                  */
    uint32_t _pad;
} prt_acq_stats;

/* Traces samples s = sample_offset + j*sample_stride < spp_total of every (angle, element).
 * Path (a, e, s) draws from PCG32 seeded by sample_tea_32(seed, (a*n_e + e)*spp_total + s).
 * channel_buf [n_a][n_e][T] f32 is OVERWRITTEN with this call's deposits (each scaled by 1/spp_total);
 * tx_delays [n_a][n_e] f32 = x_e sin(theta_a)/c  (CustomIntegrator.py:87,254-257). */
int prt_acquire(prt_scene *, const prt_acq_params *, uint64_t seed, uint32_t spp_total, uint32_t sample_offset,
                uint32_t sample_stride, float *channel_buf, float *tx_delays, prt_acq_stats *stats /*nullable*/);
/* Same, accumulating (+=) into a caller-owned DEVICE buffer on `stream`; asynchronous.  stats_dev:
 * 5 x uint64 {paths, segments, rays, deposits, misses} accumulated with atomics, or NULL. */
int prt_acquire_dev(prt_scene *, const prt_acq_params *, uint64_t seed, uint32_t spp_total, uint32_t sample_offset,
                    uint32_t sample_stride, float *channel_buf_dev, float *tx_delays_dev, uint64_t *stats_dev,
                    void *stream);
/* Same for the steering angles [angle_first, angle_first + angle_count) only (they own disjoint slices of
 * channel_buf): lets a multi-GPU caller start the all-reduce of one angle's slice while the next angle is traced. */
int prt_acquire_dev_angles(prt_scene *, const prt_acq_params *, uint64_t seed, uint32_t spp_total,
                           uint32_t sample_offset, uint32_t sample_stride, int32_t angle_first, int32_t angle_count,
                           float *channel_buf_dev, float *tx_delays_dev, uint64_t *stats_dev, void *stream);

/* decision-level trace of selected paths (parity tests): rec [n][max_depth] */
typedef struct {
    int32_t valid, prim, shape, recv, visible, reflect, k, survive;
    float   t, total_time, press, amp, atten, dir[3];
} prt_seg_record;
/* "next" row f2: the finite-difference loop of the driver (/root/reference/USMain.py:262-289: forward(rough),
 * forward(rough + eps) = params['shape.bsdf.roughness'] = v; params.update(); simulate_acquisition_parallel).
 * Traces n_values (<= PRT_MAX_VARIANTS) acquisitions that differ in ONE ultrasound_bsdf parameter
 * (param_index 0 = impedance, 1 = roughness; CustomBSDF.py:12-18) of the materials in material_mask (bit = material
 * id; the driver's key 'shape.bsdf.roughness' addresses every ultrasound_bsdf at once), with common random numbers
 * (same seed, same per-path PCG32 streams), into channel_bufs [n_values][n_a][n_e][T].  The scene's stored
 * parameter is not modified.  stats: nullable, [n_values]; kernel_ms / total_ms are those of the whole call. */
#define PRT_MAX_VARIANTS 16
int prt_acquire_variants(prt_scene *, const prt_acq_params *, uint64_t seed, uint32_t spp_total, uint32_t sample_offset,
                         uint32_t sample_stride, uint64_t material_mask, int param_index, const double *values,
                         uint32_t n_values, float *channel_bufs, float *tx_delays /*nullable*/,
                         prt_acq_stats *stats /*nullable*/);
int prt_acquire_trace(prt_scene *, const prt_acq_params *, uint64_t seed, uint32_t spp_total,
                      const uint64_t *path_idx, uint64_t n, prt_seg_record *rec);

/* ---- light-transport path tracer == mi.render(scene) with the `path` integrator that
 * scenes/cbox.xml:5-9 names (max_depth, rr_depth 5), perspective sensor (:11-21), independent
 * sampler (:22-24), hdrfilm + tent filter (:25-31).  Mitsuba semantics: SURVEY.md Appendix C.7. */
typedef struct {
    double   to_world[16];
    double   fov_deg;            /* along the SMALLER image axis (cbox.xml:12) */
    double   near_clip, far_clip;
    int32_t  width, height;
    int32_t  max_depth, rr_depth;
    int32_t  rfilter;            /* 0 = box, 1 = tent (radius 1) */
    int32_t  _pad;
} prt_render_params;

typedef struct {
    uint64_t paths, segments, rays, shadow_rays;
    float    kernel_ms, total_ms;
    uint32_t launches, _pad;
} prt_render_stats;

/* film_rgbw [H][W][4] f32 = (sum w*R, sum w*G, sum w*B, sum w), OVERWRITTEN.  Sample s of pixel (x,y)
 * uses PCG32 stream sample_tea_32(seed, (y*W + x)*spp_total + s) (64-bit index folded as documented). */
int prt_render_path(prt_scene *, const prt_render_params *, uint64_t seed, uint32_t spp_total, uint32_t sample_offset,
                    uint32_t sample_stride, float *film_rgbw, prt_render_stats *stats /*nullable*/);
int prt_render_path_dev(prt_scene *, const prt_render_params *, uint64_t seed, uint32_t spp_total,
                        uint32_t sample_offset, uint32_t sample_stride, float *film_rgbw_dev, uint64_t *stats_dev,
                        void *stream);
/* == mi.render(scene) as /root/reference/RayTracingV0.py:49 calls it: the film is developed ON THE DEVICE
 * (hdrfilm: rgb = sum(w c) / sum(w), 0 where no sample landed) and only image_rgb [H][W][3] f32 crosses the
 * bus.  A page-locked destination (prt_host_alloc) is written directly, without a staging copy. */
int prt_render_image(prt_scene *, const prt_render_params *, uint64_t seed, uint32_t spp_total, uint32_t sample_offset,
                     uint32_t sample_stride, float *image_rgb, prt_render_stats *stats /*nullable*/);
/* develop a device-resident RGBW film (e.g. after the all-reduce over sample shards) into rgb_dev [n_pixels][3] */
int prt_film_develop_dev(prt_context *, const float *film_rgbw_dev, uint64_t n_pixels, float *rgb_dev, void *stream);

/* ---- "next" row f1: delay-and-sum beamformer + envelope (replaces ultraspy, USMain.py:129-208) -- */
typedef struct {
    int32_t n_angles, n_elements, time_samples, nx, nz;
    double  fs, sound_speed, pitch, t0;
    double  f_number;            /* 0 = full aperture */
} prt_das_params;
/* channel [n_a][n_e][T], tx_delays [n_a][n_e], angles_deg [n_a], x [nx], z [nz] -> rf [nx][nz] f32 and
 * envelope [nx][nz] f32 (magnitude of the analytic signal along z) */
int prt_das_beamform(prt_context *, const prt_das_params *, const float *channel, const float *tx_delays,
                     const double *angles_deg, const float *x, const float *z, float *rf, float *envelope);
/* envelope of an already beamformed image rf [nx][nz] (DelayAndSum.compute_envelope, USMain.py:208) */
int prt_envelope(prt_context *, const float *rf, int32_t nx, int32_t nz, float *envelope);


/* ---- the driver's whole us_render() (/root/reference/USMain.py:92-224) in one call, channel data resident on the
 * device: acquisition -> optional pulse shaping (sigma = wave_cycles / (4 f)) -> delay-and-sum -> envelope -> log
 * compression (db = 20 log10(env + 1e-12), clipped to [max - dynamic_range_db, max], scaled to [0, 1]).
 * x [nx], z [nz] host; bmode [nz][nx] (the driver's display_image, :224), envelope [nx][nz] (nullable).  Single GPU:
 * with sample shards the channel buffers must be all-reduced before beamforming (use prt_acquire_dev + prt_das_beamform). */
typedef struct {
    int32_t nx, nz;
    double  t0, f_number;        /* as prt_das_params */
    int32_t shape_pulse, _pad;
    double  wave_cycles;         /* CustomIntegrator.py:20 */
    double  dynamic_range_db;    /* 60 in the driver (USMain.py:213) */
} prt_us_render_params;
int prt_us_render(prt_scene *, const prt_acq_params *, uint64_t seed, uint32_t spp_total, uint32_t sample_offset,
                  uint32_t sample_stride, const prt_us_render_params *, const float *x, const float *z, float *bmode,
                  float *envelope /*nullable*/, prt_acq_stats *stats /*nullable*/);
/* us_render() minus the acquisition, on a channel buffer that already lives on the device (e.g. the all-reduced buffer of a
 * sample-sharded multi-GPU acquisition): pulse shaping (optional) -> delay-and-sum -> envelope -> log compression; bmode [nz][nx] and
 * envelope [nx][nz] (nullable) are host buffers.  Synchronises `stream` before returning. */
int prt_us_postprocess_dev(prt_context *, const prt_acq_params *, const prt_us_render_params *, const float *x, const float *z,
                           const float *channel_dev, void *stream, float *bmode, float *envelope /*nullable*/);

/* ---- "next" row f4: pulse shaping (prototype at /root/reference/RayTracingV0.py:185-204, "UltraRay Eq. 14").
 * channel [n_rows][T] of delta echoes -> out [n_rows][T] = zero-phase convolution of every row with
 * h(t) = sin(2 pi fc t) exp(-t^2 / sigma_s^2), truncated at |t| <= 4 sigma_s (<= 1024 samples either side).
 * UltraIntegrator's `wave_cycles` (CustomIntegrator.py:20, unused there) maps to sigma_s = wave_cycles / (4 fc). */
int prt_pulse_shape(prt_context *, const float *channel, uint64_t n_rows, int32_t time_samples, double fs, double fc,
                    double sigma_s, float *out);
int prt_pulse_shape_dev(prt_context *, const float *channel_dev, uint64_t n_rows, int32_t time_samples, double fs, double fc,
                        double sigma_s, float *out_dev, void *stream);

#ifdef __cplusplus
}
#endif
#endif
