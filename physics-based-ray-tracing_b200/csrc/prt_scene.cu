// prt_scene.cu -- context + scene container behind the C ABI: host-side assembly of analytic primitives,
// materials and world-space triangle soup, upload to device-resident float4 SoA buffers, LBVH build.
// Replaces mi.set_variant / mi.load_dict / mi.traverse(...).update() (/root/reference/USMain.py:12,257-265).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <tuple>

#include "prt_internal.h"

namespace prt {

static thread_local std::string g_last_error;

void set_error(const std::string &msg) { g_last_error = msg; }

int cuda_fail(cudaError_t e, const char *what, const char *file, int line) {
    g_last_error = std::string("CUDA error: ") + cudaGetErrorString(e) + " in `" + what + "` at " + file + ":" + std::to_string(line);
    return PRT_ERR_CUDA;
}

static inline float __int_as_float_host(int i) {
    float f;
    memcpy(&f, &i, sizeof f);
    return f;
}

static bool invert_affine(const double m[16], double inv[12]) {
    double a = m[0], b = m[1], c = m[2], d = m[4], e = m[5], f = m[6], g = m[8], h = m[9], i = m[10];
    double det = a * (e * i - f * h) - b * (d * i - f * g) + c * (d * h - e * g);
    if (det == 0.0 || !std::isfinite(det)) return false;
    double id = 1.0 / det;
    double r[9] = { (e * i - f * h) * id, (c * h - b * i) * id, (b * f - c * e) * id, (f * g - d * i) * id, (a * i - c * g) * id,
                    (c * d - a * f) * id, (d * h - e * g) * id, (b * g - a * h) * id, (a * e - b * d) * id };
    for (int k = 0; k < 3; k++) {
        inv[4 * k] = r[3 * k];
        inv[4 * k + 1] = r[3 * k + 1];
        inv[4 * k + 2] = r[3 * k + 2];
        inv[4 * k + 3] = -(r[3 * k] * m[3] + r[3 * k + 1] * m[7] + r[3 * k + 2] * m[11]);
    }
    return true;
}

int scratch_slot(prt_context *ctx, int slot, size_t bytes, void **out) {
    prt_context::Scratch &s = ctx->scratch[slot];
    if (bytes > s.cap) {
        if (s.p) cudaFree(s.p);
        s.p = nullptr;
        s.cap = 0;
        PRT_CUDA(cudaMalloc(&s.p, bytes));
        s.cap = bytes;
    }
    *out = s.p;
    return PRT_OK;
}

int ensure_scratch(prt_context *ctx, size_t acc_floats, size_t aux_floats, size_t n_angles) {
    if (acc_floats > ctx->acc_cap) {
        if (ctx->acc_dev) cudaFree(ctx->acc_dev);
        ctx->acc_dev = nullptr;
        ctx->acc_cap = 0;
        PRT_CUDA(cudaMalloc(&ctx->acc_dev, sizeof(float) * acc_floats));
        ctx->acc_cap = acc_floats;
    }
    if (aux_floats > ctx->aux_cap) {
        if (ctx->aux_dev) cudaFree(ctx->aux_dev);
        ctx->aux_dev = nullptr;
        ctx->aux_cap = 0;
        PRT_CUDA(cudaMalloc(&ctx->aux_dev, sizeof(float) * aux_floats));
        ctx->aux_cap = aux_floats;
    }
    if (n_angles > ctx->angles_cap) {
        if (ctx->angles_dev) cudaFree(ctx->angles_dev);
        ctx->angles_dev = nullptr;
        ctx->angles_cap = 0;
        PRT_CUDA(cudaMalloc(&ctx->angles_dev, sizeof(double) * n_angles));
        ctx->angles_cap = n_angles;
    }
    size_t need = sizeof(float) * (acc_floats + aux_floats);
    if (need > ctx->pinned_cap) {
        if (ctx->pinned) cudaFreeHost(ctx->pinned);
        ctx->pinned = nullptr;
        ctx->pinned_cap = 0;
        PRT_CUDA(cudaMallocHost(&ctx->pinned, need));
        ctx->pinned_cap = need;
    }
    return PRT_OK;
}

}  // namespace prt

using namespace prt;

prt::DScene prt_scene::view() const {
    DScene v;
    v.prims = prims_dev;
    v.mats = mats_dev;
    v.nodes = nodes_dev;
    v.tri_v = tri_v_dev;
    v.tri_n = tri_n_dev;
    v.tri_info = tri_info_dev;
    v.n_prims = (int) prims.size();
    v.n_mats = (int) mats.size();
    v.n_tris = (int) n_tris;
    v.n_small = (int) n_small;
    v.root_ref = root_ref;
    v.em_tri = em_tri_dev;
    v.em_first = em_first_dev;
    v.em_inv_area = em_inv_area_dev;
    v.shape_emitter = shape_emitter_dev;
    v.n_emitters = n_emitters;
    v.n_shapes = n_shapes;
    v.nodes8 = nodes8_dev;
    v.tri_v8 = tri_v8_dev;
    v.tri8_sorted = tri8_sorted_dev;
    v.n_nodes8 = (int) n_nodes8;
    return v;
}

extern "C" {

const char *prt_last_error(void) { return g_last_error.c_str(); }
const char *prt_version(void) { return "prt_b200 0.1.0 (sm_100a)"; }

int prt_device_count(int *count) {
    PRT_REQUIRE(count, "prt_device_count: null output");
    *count = 0;
    PRT_CUDA(cudaGetDeviceCount(count));
    return PRT_OK;
}

int prt_create(int device, prt_context **out) {
    PRT_REQUIRE(out, "prt_create: null output");
    *out = nullptr;
    int n = 0;
    PRT_CUDA(cudaGetDeviceCount(&n));
    if (device < 0 || device >= n) {
        set_error("prt_create: no such CUDA device " + std::to_string(device) + " (" + std::to_string(n) + " visible); there is no CPU fallback");
        return PRT_ERR_CUDA;
    }
    PRT_CUDA(cudaSetDevice(device));
    prt_context *c = new prt_context();
    c->device = device;
    PRT_CUDA(cudaGetDeviceProperties(&c->prop, device));
    c->sm_count = c->prop.multiProcessorCount;
    c->acc_dev = c->aux_dev = nullptr;
    c->acc_cap = c->aux_cap = 0;
    c->angles_dev = nullptr;
    c->angles_cap = 0;
    c->pinned = nullptr;
    c->pinned_cap = 0;
    c->stats_dev = nullptr;
    PRT_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    PRT_CUDA(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    PRT_CUDA(cudaEventCreateWithFlags(&c->slice_done[0], cudaEventDisableTiming));
    PRT_CUDA(cudaEventCreateWithFlags(&c->slice_done[1], cudaEventDisableTiming));
    PRT_CUDA(cudaMalloc(&c->stats_dev, sizeof(uint64_t) * 8 * PRT_MAX_VARIANTS));
    *out = c;
    return PRT_OK;
}

int prt_destroy(prt_context *c) {
    if (!c) return PRT_OK;
    cudaSetDevice(c->device);
    if (c->acc_dev) cudaFree(c->acc_dev);
    if (c->aux_dev) cudaFree(c->aux_dev);
    if (c->angles_dev) cudaFree(c->angles_dev);
    for (auto &t : c->angle_tables) cudaFree(t.dev);
    if (c->stats_dev) cudaFree(c->stats_dev);
    if (c->pinned) cudaFreeHost(c->pinned);
    if (c->wf_dev) cudaFree(c->wf_dev);
    for (auto &sl : c->scratch)
        if (sl.p) cudaFree(sl.p);
    for (auto &pp : c->prof) { cudaEventDestroy(pp.e0); cudaEventDestroy(pp.e1); }
    cudaStreamDestroy(c->stream);
    cudaStreamDestroy(c->copy_stream);
    cudaEventDestroy(c->slice_done[0]);
    cudaEventDestroy(c->slice_done[1]);
    delete c;
    return PRT_OK;
}

int prt_profile_begin(prt_context *c) {
    PRT_REQUIRE(c, "prt_profile_begin: null context");
    std::lock_guard<std::mutex> lk(c->mtx);
    for (auto &pp : c->prof) { cudaEventDestroy(pp.e0); cudaEventDestroy(pp.e1); }
    c->prof.clear();
    c->prof_on = true;
    return PRT_OK;
}

int prt_profile_read(prt_context *c, prt_kernel_times *out) {
    PRT_REQUIRE(c && out, "prt_profile_read: null argument");
    std::lock_guard<std::mutex> lk(c->mtx);
    PRT_CUDA(cudaSetDevice(c->device));
    for (int k = 0; k < PRT_KC_COUNT; k++) { out->ms[k] = 0.0; out->launches[k] = 0; }
    int rc = PRT_OK;
    for (auto &pp : c->prof) {
        float ms = 0.0f;
        cudaError_t e = cudaEventSynchronize(pp.e1);
        if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, pp.e0, pp.e1);
        if (e != cudaSuccess && rc == PRT_OK) rc = cuda_fail(e, "prt_profile_read", __FILE__, __LINE__);
        out->ms[pp.cls] += ms;
        out->launches[pp.cls] += (uint32_t) pp.kernels;
        cudaEventDestroy(pp.e0);
        cudaEventDestroy(pp.e1);
    }
    c->prof.clear();
    c->prof_on = false;
    return rc;
}

int prt_device_info(prt_context *c, int *sm_count, int *cc_major, int *cc_minor, uint64_t *global_mem_bytes) {
    PRT_REQUIRE(c, "prt_device_info: null context");
    if (sm_count) *sm_count = c->sm_count;
    if (cc_major) *cc_major = c->prop.major;
    if (cc_minor) *cc_minor = c->prop.minor;
    if (global_mem_bytes) *global_mem_bytes = (uint64_t) c->prop.totalGlobalMem;
    return PRT_OK;
}

int prt_host_alloc(prt_context *c, uint64_t bytes, void **out) {
    PRT_REQUIRE(c && out && bytes > 0, "prt_host_alloc: invalid argument");
    *out = nullptr;
    PRT_CUDA(cudaSetDevice(c->device));
    PRT_CUDA(cudaHostAlloc(out, (size_t) bytes, cudaHostAllocPortable));
    return PRT_OK;
}

int prt_host_free(prt_context *c, void *ptr) {
    PRT_REQUIRE(c, "prt_host_free: null context");
    if (ptr) PRT_CUDA(cudaFreeHost(ptr));
    return PRT_OK;
}

int prt_scene_create(prt_context *c, prt_scene **out) {
    PRT_REQUIRE(c && out, "prt_scene_create: null argument");
    prt_scene *s = new prt_scene();
    s->ctx = c;
    s->n_shapes = 0;
    s->committed = false;
    s->prims_dev = nullptr;
    s->mats_dev = nullptr;
    s->nodes_dev = s->tri_v_dev = s->tri_n_dev = nullptr;
    s->tri_info_dev = nullptr;
    s->em_tri_dev = nullptr;
    s->em_first_dev = s->shape_emitter_dev = nullptr;
    s->em_inv_area_dev = nullptr;
    s->n_emitters = 0;
    s->n_tris = s->n_nodes = 0;
    s->root_ref = -1;
    s->device_bytes = 0;
    memset(&s->stats, 0, sizeof s->stats);
    *out = s;
    return PRT_OK;
}

static void free_device(prt_scene *s) {
    cudaSetDevice(s->ctx->device);
    free_topology(s->topo);
    if (s->prims_dev) cudaFree(s->prims_dev);
    if (s->mats_dev) cudaFree(s->mats_dev);
    if (s->nodes_dev) cudaFree(s->nodes_dev);
    if (s->tri_v_dev) cudaFree(s->tri_v_dev);
    if (s->tri_n_dev) cudaFree(s->tri_n_dev);
    if (s->tri_info_dev) cudaFree(s->tri_info_dev);
    if (s->nodes8_dev) cudaFree(s->nodes8_dev);
    if (s->tri_v8_dev) cudaFree(s->tri_v8_dev);
    if (s->tri8_sorted_dev) cudaFree(s->tri8_sorted_dev);
    s->nodes8_dev = s->tri_v8_dev = nullptr;
    s->tri8_sorted_dev = nullptr;
    s->n_nodes8 = 0;
    if (s->em_tri_dev) cudaFree(s->em_tri_dev);
    if (s->em_first_dev) cudaFree(s->em_first_dev);
    if (s->shape_emitter_dev) cudaFree(s->shape_emitter_dev);
    if (s->em_inv_area_dev) cudaFree(s->em_inv_area_dev);
    s->em_tri_dev = nullptr;
    s->em_first_dev = s->shape_emitter_dev = nullptr;
    s->em_inv_area_dev = nullptr;
    s->n_emitters = 0;
    s->prims_dev = nullptr;
    s->mats_dev = nullptr;
    s->nodes_dev = s->tri_v_dev = s->tri_n_dev = nullptr;
    s->tri_info_dev = nullptr;
    s->device_bytes = 0;
}

int prt_scene_destroy(prt_scene *s) {
    if (!s) return PRT_OK;
    std::lock_guard<std::mutex> lk(s->ctx->mtx);
    free_device(s);
    delete s;
    return PRT_OK;
}

int prt_scene_add_material(prt_scene *s, int kind, const double p[8], const double emission_rgb[3], int *material_id) {
    PRT_REQUIRE(s, "prt_scene_add_material: null scene");
    PRT_REQUIRE(kind >= PRT_MAT_ULTRA && kind <= PRT_MAT_NULL, "prt_scene_add_material: unknown material kind");
    DMaterial m;
    memset(&m, 0, sizeof m);
    m.kind = kind;
    for (int i = 0; i < 7; i++) m.p[i] = p ? (float) p[i] : 0.0f;
    for (int i = 0; i < 3; i++) m.emission[i] = emission_rgb ? (float) emission_rgb[i] : 0.0f;
    s->mats.push_back(m);
    s->committed = false;
    if (material_id) *material_id = (int) s->mats.size() - 1;
    return PRT_OK;
}

int prt_scene_set_material_param(prt_scene *s, int material_id, int index, double value) {
    PRT_REQUIRE(s, "prt_scene_set_material_param: null scene");
    PRT_REQUIRE(material_id >= 0 && material_id < (int) s->mats.size() && index >= 0 && index < 7,
                "prt_scene_set_material_param: index out of range");
    std::lock_guard<std::mutex> lk(s->ctx->mtx);
    s->mats[material_id].p[index] = (float) value;
    if (s->committed && s->mats_dev) {
        PRT_CUDA(cudaSetDevice(s->ctx->device));
        PRT_CUDA(cudaMemcpy(s->mats_dev + material_id, &s->mats[material_id], sizeof(DMaterial), cudaMemcpyHostToDevice));
    }
    return PRT_OK;
}

int prt_scene_add_primitive(prt_scene *s, int kind, const double to_world[16], int material_id, int flip_normals, int *shape_id) {
    PRT_REQUIRE(s && to_world, "prt_scene_add_primitive: null argument");
    PRT_REQUIRE(kind >= PRT_SPHERE && kind <= PRT_CYLINDER, "prt_scene_add_primitive: unknown primitive kind");
    PRT_REQUIRE(material_id >= 0 && material_id < (int) s->mats.size(), "prt_scene_add_primitive: unknown material");
    double inv[12];
    PRT_REQUIRE(invert_affine(to_world, inv), "prt_scene_add_primitive: singular to_world");
    DPrim p;
    const double *m = to_world;
    p.w0 = make_float4((float) m[0], (float) m[1], (float) m[2], (float) m[3]);
    p.w1 = make_float4((float) m[4], (float) m[5], (float) m[6], (float) m[7]);
    p.w2 = make_float4((float) m[8], (float) m[9], (float) m[10], (float) m[11]);
    p.o0 = make_float4((float) inv[0], (float) inv[1], (float) inv[2], (float) inv[3]);
    p.o1 = make_float4((float) inv[4], (float) inv[5], (float) inv[6], (float) inv[7]);
    p.o2 = make_float4((float) inv[8], (float) inv[9], (float) inv[10], (float) inv[11]);
    if (kind == PRT_SPHERE) {
        // centre = to_world*(0,0,0), radius = |to_world*(1,0,0)|
        double r = std::sqrt(m[0] * m[0] + m[4] * m[4] + m[8] * m[8]);
        p.aux = make_float4((float) m[3], (float) m[7], (float) m[11], (float) r);
    } else {
        // world normal of the object-space +z plane: normalize(inverse-transpose * (0,0,1)), fp32 like the oracle
        float nx = (float) inv[8], ny = (float) inv[9], nz = (float) inv[10];
        float l = std::sqrt(nx * nx + ny * ny + nz * nz);
        p.aux = make_float4(nx / l, ny / l, nz / l, 0.0f);
    }
    p.kind = kind;
    p.material = material_id;
    p.flip = flip_normals ? 1 : 0;
    p.shape = s->n_shapes;
    s->prims.push_back(p);
    s->committed = false;
    if (shape_id) *shape_id = s->n_shapes;
    s->n_shapes++;
    return PRT_OK;
}

int prt_scene_add_mesh(prt_scene *s, const double *v, uint32_t nv, const double *vn, const uint32_t *idx, uint32_t nt,
                       const double to_world[16], int material_id, int flip_normals, int *shape_id) {
    PRT_REQUIRE(s && v && idx && to_world, "prt_scene_add_mesh: null argument");
    PRT_REQUIRE(material_id >= 0 && material_id < (int) s->mats.size(), "prt_scene_add_mesh: unknown material");
    double inv[12];
    PRT_REQUIRE(invert_affine(to_world, inv), "prt_scene_add_mesh: singular to_world");
    HostMesh hm;
    hm.has_n = vn != nullptr;
    hm.shape = s->n_shapes;
    hm.material = material_id;
    hm.flip = flip_normals ? 1 : 0;
    hm.nt = nt;
    memcpy(hm.to_world, to_world, sizeof hm.to_world);
    hm.v.resize((size_t) nt * 9);
    if (vn) hm.n.resize((size_t) nt * 9);
    const double *m = to_world;
    for (uint32_t t = 0; t < nt; t++) {
        for (int c = 0; c < 3; c++) {
            uint32_t vi = idx[3 * (size_t) t + c];
            if (vi >= nv) {
                set_error("prt_scene_add_mesh: vertex index out of range");
                return PRT_ERR_INVALID;
            }
            const double *p = v + 3 * (size_t) vi;
            for (int r = 0; r < 3; r++)
                hm.v[9 * (size_t) t + 3 * c + r] = (float) (m[4 * r] * p[0] + m[4 * r + 1] * p[1] + m[4 * r + 2] * p[2] + m[4 * r + 3]);
            if (vn) {
                const double *n = vn + 3 * (size_t) vi;
                double w[3];
                for (int r = 0; r < 3; r++) w[r] = inv[r] * n[0] + inv[4 + r] * n[1] + inv[8 + r] * n[2];
                double l = std::sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
                for (int r = 0; r < 3; r++) hm.n[9 * (size_t) t + 3 * c + r] = l > 0 ? (float) (w[r] / l) : 0.0f;
            }
        }
    }
    s->meshes.push_back(std::move(hm));
    s->committed = false;
    if (shape_id) *shape_id = s->n_shapes;
    s->n_shapes++;
    return PRT_OK;
}

__global__ void k_gather_aux(const uint32_t *__restrict__ order, uint32_t n, const int4 *__restrict__ info_in,
                             const float4 *__restrict__ n_in, int4 *info_out, float4 *n_out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t src = order[i];
    info_out[i] = info_in[src];
    if (n_in) {
        n_out[3 * (size_t) i] = n_in[3 * (size_t) src];
        n_out[3 * (size_t) i + 1] = n_in[3 * (size_t) src + 1];
        n_out[3 * (size_t) i + 2] = n_in[3 * (size_t) src + 2];
    }
}

}  // extern "C"

// area-emitter tables: emissive mesh shapes, uniform pick over emitters, area-weighted face pick (Mitsuba).  (Re)built from
// the host meshes: at commit, and after a transform change of an emissive shape.
static int build_emitters(prt_scene *s, size_t *bytes_out) {
    if (s->em_tri_dev) cudaFree(s->em_tri_dev);
    if (s->em_first_dev) cudaFree(s->em_first_dev);
    if (s->shape_emitter_dev) cudaFree(s->shape_emitter_dev);
    if (s->em_inv_area_dev) cudaFree(s->em_inv_area_dev);
    s->em_tri_dev = nullptr;
    s->em_first_dev = s->shape_emitter_dev = nullptr;
    s->em_inv_area_dev = nullptr;
    s->n_emitters = 0;
        std::vector<float4> et;
        std::vector<int> first, shape_em((size_t) s->n_shapes + 1, -1);
        std::vector<float> inv_area;
        for (auto &m : s->meshes) {
            const DMaterial &mat = s->mats[m.material];
            if (!(mat.emission[0] > 0 || mat.emission[1] > 0 || mat.emission[2] > 0) || m.nt == 0) continue;
            shape_em[m.shape] = (int) inv_area.size();
            first.push_back((int) (et.size() / 3));
            double run = 0.0;
            for (uint32_t t = 0; t < m.nt; t++) {
                const float *v = &m.v[9 * (size_t) t];
                double e0[3] = { (double) v[3] - v[0], (double) v[4] - v[1], (double) v[5] - v[2] };
                double e1[3] = { (double) v[6] - v[0], (double) v[7] - v[1], (double) v[8] - v[2] };
                double cx = e0[1] * e1[2] - e0[2] * e1[1], cy = e0[2] * e1[0] - e0[0] * e1[2], cz = e0[0] * e1[1] - e0[1] * e1[0];
                run += 0.5 * std::sqrt(cx * cx + cy * cy + cz * cz);
                et.push_back(make_float4(v[0], v[1], v[2], (float) run));
                et.push_back(make_float4(v[3], v[4], v[5], __int_as_float_host(m.material)));
                et.push_back(make_float4(v[6], v[7], v[8], m.flip ? 1.0f : 0.0f));
            }
            inv_area.push_back((float) (1.0 / run));
        }
        first.push_back((int) (et.size() / 3));
        s->n_emitters = (int) inv_area.size();
        PRT_CUDA(cudaMalloc(&s->shape_emitter_dev, sizeof(int) * shape_em.size()));
        PRT_CUDA(cudaMemcpy(s->shape_emitter_dev, shape_em.data(), sizeof(int) * shape_em.size(), cudaMemcpyHostToDevice));
        PRT_CUDA(cudaMalloc(&s->em_first_dev, sizeof(int) * first.size()));
        PRT_CUDA(cudaMemcpy(s->em_first_dev, first.data(), sizeof(int) * first.size(), cudaMemcpyHostToDevice));
        if (s->n_emitters) {
            PRT_CUDA(cudaMalloc(&s->em_tri_dev, sizeof(float4) * et.size()));
            PRT_CUDA(cudaMemcpy(s->em_tri_dev, et.data(), sizeof(float4) * et.size(), cudaMemcpyHostToDevice));
            PRT_CUDA(cudaMalloc(&s->em_inv_area_dev, sizeof(float) * inv_area.size()));
            PRT_CUDA(cudaMemcpy(s->em_inv_area_dev, inv_area.data(), sizeof(float) * inv_area.size(), cudaMemcpyHostToDevice));
        }
        *bytes_out = sizeof(float4) * et.size() + sizeof(int) * (first.size() + shape_em.size());
    return PRT_OK;
}

namespace prt {
struct XformDev { float D[12], Dinv[12]; };
// moves the (sorted) triangles of one shape: vertices by D, corner normals by the inverse transpose; .w words are kept
__global__ void k_xform_shape(uint32_t n, int shape, XformDev X, const int4 *__restrict__ info, float4 *tri_v, float4 *tri_n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || info[i].y != shape) return;
#pragma unroll
    for (int c = 0; c < 3; c++) {
        float4 v = tri_v[3 * (size_t) i + c];
        const float x = v.x, y = v.y, z = v.z;
        v.x = fmaf(X.D[0], x, fmaf(X.D[1], y, fmaf(X.D[2], z, X.D[3])));
        v.y = fmaf(X.D[4], x, fmaf(X.D[5], y, fmaf(X.D[6], z, X.D[7])));
        v.z = fmaf(X.D[8], x, fmaf(X.D[9], y, fmaf(X.D[10], z, X.D[11])));
        tri_v[3 * (size_t) i + c] = v;
        if (tri_n && (info[i].w & 1)) {
            float4 q = tri_n[3 * (size_t) i + c];
            const float a = q.x, b = q.y, d = q.z;
            float wx = X.Dinv[0] * a + X.Dinv[4] * b + X.Dinv[8] * d, wy = X.Dinv[1] * a + X.Dinv[5] * b + X.Dinv[9] * d,
                  wz = X.Dinv[2] * a + X.Dinv[6] * b + X.Dinv[10] * d;
            const float l = sqrtf(wx * wx + wy * wy + wz * wz);
            if (l > 0.0f) { q.x = wx / l; q.y = wy / l; q.z = wz / l; }
            tri_n[3 * (size_t) i + c] = q;
        }
    }
}

}  // namespace prt

extern "C" {

int prt_scene_commit(prt_scene *s, prt_bvh_stats *out) {
    PRT_REQUIRE(s, "prt_scene_commit: null scene");
    std::lock_guard<std::mutex> lk(s->ctx->mtx);
    PRT_CUDA(cudaSetDevice(s->ctx->device));
    free_device(s);
    memset(&s->stats, 0, sizeof s->stats);
    cudaStream_t st = s->ctx->stream;
    size_t bytes = 0;
    if (!s->prims.empty()) {
        PRT_CUDA(cudaMalloc(&s->prims_dev, sizeof(DPrim) * s->prims.size()));
        PRT_CUDA(cudaMemcpy(s->prims_dev, s->prims.data(), sizeof(DPrim) * s->prims.size(), cudaMemcpyHostToDevice));
        bytes += sizeof(DPrim) * s->prims.size();
    }
    if (!s->mats.empty()) {
        PRT_CUDA(cudaMalloc(&s->mats_dev, sizeof(DMaterial) * s->mats.size()));
        PRT_CUDA(cudaMemcpy(s->mats_dev, s->mats.data(), sizeof(DMaterial) * s->mats.size(), cudaMemcpyHostToDevice));
        bytes += sizeof(DMaterial) * s->mats.size();
    }
    uint64_t nt = 0;
    bool any_n = false;
    for (auto &m : s->meshes) {
        nt += m.nt;
        any_n |= m.has_n;
    }
    PRT_REQUIRE(nt < (1ull << 29), "prt_scene_commit: too many triangles (limit 2^29)");
    s->n_tris = (uint32_t) nt;
    s->n_small = 0;
    s->root_ref = -1;
    if (nt) {
        // host staging in input order: float4 vertices, int4 info, float4 normals
        std::vector<float4> hv(nt * 3), hn(any_n ? nt * 3 : 0);
        std::vector<int4> hi(nt);
        size_t o = 0;
        for (auto &m : s->meshes) {
            for (uint32_t t = 0; t < m.nt; t++, o++) {
                for (int c = 0; c < 3; c++) {
                    hv[3 * o + c] = make_float4(m.v[9 * (size_t) t + 3 * c], m.v[9 * (size_t) t + 3 * c + 1], m.v[9 * (size_t) t + 3 * c + 2], 0.0f);
                    if (any_n)
                        hn[3 * o + c] = m.has_n ? make_float4(m.n[9 * (size_t) t + 3 * c], m.n[9 * (size_t) t + 3 * c + 1], m.n[9 * (size_t) t + 3 * c + 2], 0.0f)
                                                : make_float4(0, 0, 0, 0);
                }
                hi[o] = make_int4((int) o, m.shape, m.material, (m.has_n ? 1 : 0) | (m.flip ? 2 : 0));
            }
        }
        // Oversized triangles (bounding-box area > 1024 x the mean; at most 21, largest first) are moved to the END of the
        // staging arrays and kept out of the LBVH (DScene::n_small).  The 8-wide tree hangs them under a super root as leaf
        // children (bvh8_write_super_root: why, and what it measured); the binary traversal of the megakernels tests them
        // ahead of the tree (test_big_tris).  PRT_BIG_TRIS=0 switches the split off.
        float blo[3] = { 3.4e38f, 3.4e38f, 3.4e38f }, bhi[3] = { -3.4e38f, -3.4e38f, -3.4e38f };
        uint64_t n_big = 0;
        {
            std::vector<float> area(nt);
            double sum = 0.0;
            for (uint64_t t = 0; t < nt; t++) {
                float lo[3], hi[3];
                for (int k = 0; k < 3; k++) {
                    const float a = (&hv[3 * t].x)[k], b = (&hv[3 * t + 1].x)[k], c = (&hv[3 * t + 2].x)[k];
                    lo[k] = std::min(a, std::min(b, c));
                    hi[k] = std::max(a, std::max(b, c));
                    blo[k] = std::min(blo[k], lo[k]);
                    bhi[k] = std::max(bhi[k], hi[k]);
                }
                const float ex = hi[0] - lo[0], ey = hi[1] - lo[1], ez = hi[2] - lo[2];
                area[t] = ex * ey + ey * ez + ez * ex;
                sum += area[t];
            }
            const char *e = getenv("PRT_BIG_TRIS");
            const bool enabled = !(e && e[0] == '0') && nt > 64;
            const float limit = (float) (1024.0 * sum / (double) nt);
            std::vector<uint64_t> big;
            if (enabled)
                for (uint64_t t = 0; t < nt; t++)
                    if (area[t] > limit) big.push_back(t);
            if (big.size() > 21) {       // seven leaf children of three triangles under the super root
                std::partial_sort(big.begin(), big.begin() + 21, big.end(), [&](uint64_t a, uint64_t b) { return area[a] > area[b]; });
                big.resize(21);
            }
            {   // consecutive triangles share a leaf child: order them by supporting plane (the two halves of a wall quad
                // have the same box), then by centroid
                auto key = [&](uint64_t t) {
                    const float4 a = hv[3 * t], b = hv[3 * t + 1], c = hv[3 * t + 2];
                    const float ux = b.x - a.x, uy = b.y - a.y, uz = b.z - a.z, vx = c.x - a.x, vy = c.y - a.y, vz = c.z - a.z;
                    const float n[3] = { uy * vz - uz * vy, uz * vx - ux * vz, ux * vy - uy * vx };
                    int ax = fabsf(n[0]) > fabsf(n[1]) ? 0 : 1;
                    if (fabsf(n[2]) > fabsf(n[ax])) ax = 2;
                    const float cen[3] = { (a.x + b.x + c.x) / 3.0f, (a.y + b.y + c.y) / 3.0f, (a.z + b.z + c.z) / 3.0f };
                    return std::make_tuple(ax, cen[ax], cen[(ax + 1) % 3], cen[(ax + 2) % 3], t);
                };
                std::sort(big.begin(), big.end(), [&](uint64_t x, uint64_t y) { return key(x) < key(y); });
            }
            n_big = big.size();
            if (n_big) {
                std::vector<char> is_big(nt, 0);
                for (uint64_t t : big) is_big[t] = 1;
                std::vector<float4> v2(nt * 3), n2(any_n ? nt * 3 : 0);
                std::vector<int4> i2(nt);
                uint64_t w = 0;
                auto move = [&](uint64_t t) {
                    for (int c = 0; c < 3; c++) {
                        v2[3 * w + c] = hv[3 * t + c];
                        if (any_n) n2[3 * w + c] = hn[3 * t + c];
                    }
                    i2[w] = hi[t];          // .x keeps the ORIGINAL triangle index (what prt_trace_closest reports)
                    w++;
                };
                for (uint64_t t = 0; t < nt; t++)
                    if (!is_big[t]) move(t);
                for (uint64_t t : big) move(t);
                hv.swap(v2);
                hn.swap(n2);
                hi.swap(i2);
            }
        }
        const uint64_t n_small = nt - n_big;
        s->n_small = (uint32_t) n_small;
        float4 *v_in = nullptr, *n_in = nullptr;
        int4 *i_in = nullptr;
        uint32_t *order = nullptr;
        struct Staging {       // input-order copies, freed on every way out of this block
            void *p[4] = { nullptr, nullptr, nullptr, nullptr };
            int n = 0;
            cudaError_t alloc(void **out, size_t b) {
                const cudaError_t e = cudaMalloc(out, b);
                if (e == cudaSuccess) p[n++] = *out;
                return e;
            }
            ~Staging() { for (int i = 0; i < n; i++) cudaFree(p[i]); }
        } staging;
        PRT_CUDA(staging.alloc((void **) &v_in, sizeof(float4) * 3 * nt));
        PRT_CUDA(staging.alloc((void **) &i_in, sizeof(int4) * nt));
        PRT_CUDA(staging.alloc((void **) &order, sizeof(uint32_t) * nt));
        PRT_CUDA(cudaMemcpy(v_in, hv.data(), sizeof(float4) * 3 * nt, cudaMemcpyHostToDevice));
        PRT_CUDA(cudaMemcpy(i_in, hi.data(), sizeof(int4) * nt, cudaMemcpyHostToDevice));
        if (any_n) {
            PRT_CUDA(staging.alloc((void **) &n_in, sizeof(float4) * 3 * nt));
            PRT_CUDA(cudaMemcpy(n_in, hn.data(), sizeof(float4) * 3 * nt, cudaMemcpyHostToDevice));
            PRT_CUDA(cudaMalloc(&s->tri_n_dev, sizeof(float4) * 3 * nt));
            bytes += sizeof(float4) * 3 * nt;
        }
        PRT_CUDA(cudaMalloc(&s->tri_v_dev, sizeof(float4) * 3 * nt));
        PRT_CUDA(cudaMalloc(&s->tri_info_dev, sizeof(int4) * nt));
        PRT_CUDA(cudaMalloc(&s->nodes_dev, sizeof(float4) * 4 * (nt > 1 ? nt - 1 : 1)));
        bytes += sizeof(float4) * 3 * nt + sizeof(int4) * nt + sizeof(float4) * 4 * (nt > 1 ? nt - 1 : 1);
        Bvh8Out b8;
        b8.n_extra = (uint32_t) n_big;
        // scenes of up to 2^22 triangles keep the tree's topology (24 B per triangle) so that a transform change can refit
        int rc = build_lbvh(s->ctx, v_in, (uint32_t) n_small, s->tri_v_dev, order, s->nodes_dev, &s->root_ref, &s->stats, st, &b8,
                            n_small <= (1u << 22) ? &s->topo : nullptr);
        if (rc) return rc;
        if (n_big) {       // the oversized triangles keep their staging order behind the sorted ones; k_gather_tris's stamp (v0.w = source index)
            std::vector<uint32_t> tail(n_big);
            for (uint64_t j = 0; j < n_big; j++) {
                tail[j] = (uint32_t) (n_small + j);
                hv[3 * (n_small + j)].w = __int_as_float_host((int) (n_small + j));
            }
            PRT_CUDA(cudaMemcpyAsync(order + n_small, tail.data(), sizeof(uint32_t) * n_big, cudaMemcpyHostToDevice, st));
            PRT_CUDA(cudaMemcpyAsync(s->tri_v_dev + 3 * n_small, hv.data() + 3 * n_small, sizeof(float4) * 3 * n_big, cudaMemcpyHostToDevice, st));
            PRT_CUDA(cudaStreamSynchronize(st));
        }
        s->nodes8_dev = b8.nodes8;          // owned by the scene from here on (freed with it, also on the error paths below)
        s->tri_v8_dev = b8.tri_v8;
        s->tri8_sorted_dev = b8.tri8_sorted;
        if (n_big && b8.n_nodes8) {
            rc = bvh8_write_super_root((uint32_t) n_small, (uint32_t) n_big, hv.data() + 3 * n_small, s->stats.scene_lo, s->stats.scene_hi,
                                       b8.nodes8, b8.tri_v8, b8.tri8_sorted, st);
            if (rc) return rc;
        }
        for (int k = 0; k < 3; k++) {       // bounds of ALL triangles (build_lbvh reported the hierarchy's)
            s->stats.scene_lo[k] = blo[k];
            s->stats.scene_hi[k] = bhi[k];
        }
        s->n_nodes8 = b8.n_nodes8;
        s->bvh8_levels = b8.levels;
        s->bvh8_build_ms = b8.build_ms;
        if (b8.n_nodes8) bytes += sizeof(float4) * BVH8_NODE_F4 * (size_t) b8.n_nodes8 + (sizeof(float4) * 3 + sizeof(uint32_t)) * nt;
        k_gather_aux<<<(unsigned) ((nt + 255) / 256), 256, 0, st>>>(order, (uint32_t) nt, i_in, n_in, s->tri_info_dev, s->tri_n_dev);
        rc = bvh8_annotate((uint32_t) nt, s->tri8_sorted_dev, s->tri_info_dev, s->mats_dev, s->tri_v8_dev, st);
        if (rc) return rc;
        PRT_CUDA(cudaStreamSynchronize(st));
        PRT_CUDA(cudaGetLastError());
        s->n_nodes = s->stats.n_nodes;
    }
    {
        size_t eb = 0;
        int rc = build_emitters(s, &eb);
        if (rc) return rc;
        bytes += eb;
    }
    s->device_bytes = bytes;
    s->stats.n_primitives = (uint32_t) s->prims.size();
    s->stats.n_triangles = s->n_tris;
    s->stats.device_bytes = bytes;
    s->stats.n_nodes8 = s->n_nodes8;
    s->stats.bvh8_levels = (uint32_t) s->bvh8_levels;
    s->stats.bvh8_build_ms = s->bvh8_build_ms;
    s->stats.n_oversized = s->n_tris - s->n_small;
    s->committed = true;
    if (out) *out = s->stats;
    return PRT_OK;
}

int prt_scene_set_shape_transform(prt_scene *s, int shape_id, const double to_world[16]) {
    PRT_REQUIRE(s && to_world, "prt_scene_set_shape_transform: null argument");
    PRT_REQUIRE(shape_id >= 0 && shape_id < s->n_shapes, "prt_scene_set_shape_transform: unknown shape");
    double inv_new[12];
    PRT_REQUIRE(invert_affine(to_world, inv_new), "prt_scene_set_shape_transform: singular to_world");
    std::unique_lock<std::mutex> lk(s->ctx->mtx);
    PRT_CUDA(cudaSetDevice(s->ctx->device));
    for (size_t i = 0; i < s->prims.size(); i++) {
        if (s->prims[i].shape != shape_id) continue;
        // analytic primitive: new rows, one 128-byte upload; no hierarchy involved
        DPrim &p = s->prims[i];
        const int kind = p.kind, material = p.material, flip = p.flip;
        const double *m = to_world;
        p.w0 = make_float4((float) m[0], (float) m[1], (float) m[2], (float) m[3]);
        p.w1 = make_float4((float) m[4], (float) m[5], (float) m[6], (float) m[7]);
        p.w2 = make_float4((float) m[8], (float) m[9], (float) m[10], (float) m[11]);
        p.o0 = make_float4((float) inv_new[0], (float) inv_new[1], (float) inv_new[2], (float) inv_new[3]);
        p.o1 = make_float4((float) inv_new[4], (float) inv_new[5], (float) inv_new[6], (float) inv_new[7]);
        p.o2 = make_float4((float) inv_new[8], (float) inv_new[9], (float) inv_new[10], (float) inv_new[11]);
        if (kind == PRT_SPHERE) {
            double r = std::sqrt(m[0] * m[0] + m[4] * m[4] + m[8] * m[8]);
            p.aux = make_float4((float) m[3], (float) m[7], (float) m[11], (float) r);
        } else {
            float nx = (float) inv_new[8], ny = (float) inv_new[9], nz = (float) inv_new[10];
            float l = std::sqrt(nx * nx + ny * ny + nz * nz);
            p.aux = make_float4(nx / l, ny / l, nz / l, 0.0f);
        }
        p.kind = kind; p.material = material; p.flip = flip; p.shape = shape_id;
        if (s->committed && s->prims_dev)
            PRT_CUDA(cudaMemcpy(s->prims_dev + i, &p, sizeof(DPrim), cudaMemcpyHostToDevice));
        return PRT_OK;
    }
    for (auto &hm : s->meshes) {
        if (hm.shape != shape_id) continue;
        // delta = new * old^-1 moves the world-space vertices the scene holds (host copy and device copy alike)
        double inv_old[12];
        PRT_REQUIRE(invert_affine(hm.to_world, inv_old), "prt_scene_set_shape_transform: singular previous transform");
        double D[12];
        for (int r = 0; r < 3; r++) {
            for (int c = 0; c < 3; c++)
                D[4 * r + c] = to_world[4 * r] * inv_old[c] + to_world[4 * r + 1] * inv_old[4 + c] + to_world[4 * r + 2] * inv_old[8 + c];
            D[4 * r + 3] = to_world[4 * r] * inv_old[3] + to_world[4 * r + 1] * inv_old[7] + to_world[4 * r + 2] * inv_old[11] + to_world[4 * r + 3];
        }
        double D16[16] = { D[0], D[1], D[2], D[3], D[4], D[5], D[6], D[7], D[8], D[9], D[10], D[11], 0, 0, 0, 1 }, Dinv[12];
        PRT_REQUIRE(invert_affine(D16, Dinv), "prt_scene_set_shape_transform: singular delta");
        for (size_t k = 0; k < (size_t) hm.nt * 3; k++) {
            float *v = &hm.v[3 * k];
            const double x = v[0], y = v[1], z = v[2];
            for (int r = 0; r < 3; r++) v[r] = (float) (D[4 * r] * x + D[4 * r + 1] * y + D[4 * r + 2] * z + D[4 * r + 3]);
            if (hm.has_n) {
                float *nn = &hm.n[3 * k];
                const double a = nn[0], b = nn[1], c = nn[2];
                double w[3];
                for (int r = 0; r < 3; r++) w[r] = Dinv[r] * a + Dinv[4 + r] * b + Dinv[8 + r] * c;
                const double l = std::sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
                for (int r = 0; r < 3; r++) nn[r] = l > 0 ? (float) (w[r] / l) : 0.0f;
            }
        }
        memcpy(hm.to_world, to_world, sizeof hm.to_world);
        if (!s->committed) return PRT_OK;
        if (!s->topo.children || s->n_small < 2) {      // no topology kept (> 2^22 triangles): full rebuild
            s->committed = false;
            lk.unlock();
            return prt_scene_commit(s, nullptr);
        }
        cudaStream_t st = s->ctx->stream;
        XformDev X;
        for (int k = 0; k < 12; k++) { X.D[k] = (float) D[k]; X.Dinv[k] = (float) Dinv[k]; }
        k_xform_shape<<<(s->n_tris + 255) / 256, 256, 0, st>>>(s->n_tris, shape_id, X, s->tri_info_dev, s->tri_v_dev, s->tri_n_dev);
        PRT_CUDA(cudaGetLastError());
        if (s->nodes8_dev) cudaFree(s->nodes8_dev);
        if (s->tri_v8_dev) cudaFree(s->tri_v8_dev);
        if (s->tri8_sorted_dev) cudaFree(s->tri8_sorted_dev);
        s->nodes8_dev = s->tri_v8_dev = nullptr;
        s->tri8_sorted_dev = nullptr;
        Bvh8Out b8;
        const uint32_t n_big = s->n_tris - s->n_small;
        b8.n_extra = n_big;
        prt_bvh_stats rs = s->stats;
        int rc = refit_lbvh(s->topo, s->tri_v_dev, s->nodes_dev, &rs, st, &b8);
        if (rc) { s->committed = false; return rc; }
        s->nodes8_dev = b8.nodes8; s->tri_v8_dev = b8.tri_v8; s->tri8_sorted_dev = b8.tri8_sorted;
        s->n_nodes8 = b8.n_nodes8; s->bvh8_levels = b8.levels;
        if (n_big && b8.n_nodes8) {      // the moved oversized triangles (k_xform_shape above) and the refitted root box
            std::vector<float4> bv(3 * (size_t) n_big);
            PRT_CUDA(cudaMemcpyAsync(bv.data(), s->tri_v_dev + 3 * (size_t) s->n_small, sizeof(float4) * 3 * n_big, cudaMemcpyDeviceToHost, st));
            PRT_CUDA(cudaStreamSynchronize(st));
            rc = bvh8_write_super_root(s->n_small, n_big, bv.data(), rs.scene_lo, rs.scene_hi, b8.nodes8, b8.tri_v8, b8.tri8_sorted, st);
            if (rc) return rc;
        }
        rc = bvh8_annotate(s->n_tris, s->tri8_sorted_dev, s->tri_info_dev, s->mats_dev, s->tri_v8_dev, st);
        if (rc) return rc;
        s->stats.build_ms = rs.build_ms;
        s->stats.sah_cost = rs.sah_cost;
        s->stats.n_nodes8 = s->n_nodes8;
        s->stats.bvh8_levels = (uint32_t) s->bvh8_levels;
        {   // bounds of everything, oversized triangles included: from the host copies
            float lo[3] = { 3.4e38f, 3.4e38f, 3.4e38f }, hi[3] = { -3.4e38f, -3.4e38f, -3.4e38f };
            for (auto &m2 : s->meshes)
                for (size_t k = 0; k < (size_t) m2.nt * 3; k++)
                    for (int r = 0; r < 3; r++) { lo[r] = std::min(lo[r], m2.v[3 * k + r]); hi[r] = std::max(hi[r], m2.v[3 * k + r]); }
            for (int r = 0; r < 3; r++) { s->stats.scene_lo[r] = lo[r]; s->stats.scene_hi[r] = hi[r]; }
        }
        const DMaterial &mat = s->mats[hm.material];
        if (mat.emission[0] > 0 || mat.emission[1] > 0 || mat.emission[2] > 0) {
            size_t eb = 0;
            rc = build_emitters(s, &eb);
            if (rc) return rc;
        }
        PRT_CUDA(cudaStreamSynchronize(st));
        return PRT_OK;
    }
    set_error("prt_scene_set_shape_transform: shape not found");
    return PRT_ERR_INVALID;
}

int prt_scene_get_stats(prt_scene *s, prt_bvh_stats *out) {
    PRT_REQUIRE(s && out, "prt_scene_get_stats: null argument");
    if (!s->committed) { set_error("prt_scene_get_stats: scene not committed"); return PRT_ERR_STATE; }
    std::lock_guard<std::mutex> lk(s->ctx->mtx);
    *out = s->stats;
    return PRT_OK;
}

}  // extern "C"
