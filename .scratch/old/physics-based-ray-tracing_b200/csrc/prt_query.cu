// prt_query.cu -- batched ray queries at the scene.ray_intersect boundary
// (/root/reference/CustomIntegrator.py:146,159,309,324) and batched UltraBSDF.sample
// (/root/reference/CustomBSDF.py:87-175).  These are the entry points the parity tests drive.
#include "prt_hit.cuh"
#include "prt_internal.h"

namespace prt {

static constexpr int MAX_SMEM_PRIMS = 64;

__device__ __forceinline__ const DPrim *stage_prims(const DScene &sc, DPrim *smem) {
    // analytic primitives live in shared memory for the whole kernel (<= 8 KB); larger sets stay in global
    if (sc.n_prims > MAX_SMEM_PRIMS) return sc.prims;
    const float4 *src = reinterpret_cast<const float4 *>(sc.prims);
    float4 *dst = reinterpret_cast<float4 *>(smem);
    for (int i = threadIdx.x; i < sc.n_prims * (int) (sizeof(DPrim) / 16); i += blockDim.x) dst[i] = src[i];
    __syncthreads();
    return smem;
}

__global__ void __launch_bounds__(256) k_trace_closest(DScene sc, const float *__restrict__ o, const float *__restrict__ d,
                                                        const float *__restrict__ tmax, uint64_t n, float *t, int32_t *prim,
                                                        int32_t *shape, float *p, float *ng, float *ns, float *wi, float *sh_s) {
    __shared__ DPrim sprims[MAX_SMEM_PRIMS];
    const DPrim *prims = stage_prims(sc, sprims);
    uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float3 oo = mk3(o[3 * i], o[3 * i + 1], o[3 * i + 2]), dd = mk3(d[3 * i], d[3 * i + 1], d[3 * i + 2]);
    Hit h;
    bool ok = closest_hit(sc, prims, oo, dd, tmax ? tmax[i] : PRT_INF, h);
    if (t) t[i] = ok ? h.t : PRT_INF;
    if (prim) prim[i] = ok ? h.prim : -1;
    if (shape) shape[i] = ok ? h.shape : -1;
    if (!ok) return;
    float3 md = -dd;
    if (p) { p[3 * i] = h.p.x; p[3 * i + 1] = h.p.y; p[3 * i + 2] = h.p.z; }
    if (ng) { ng[3 * i] = h.ng.x; ng[3 * i + 1] = h.ng.y; ng[3 * i + 2] = h.ng.z; }
    if (ns) { ns[3 * i] = h.ns.x; ns[3 * i + 1] = h.ns.y; ns[3 * i + 2] = h.ns.z; }
    if (wi) { wi[3 * i] = dot(md, h.fs); wi[3 * i + 1] = dot(md, h.ft); wi[3 * i + 2] = dot(md, h.ns); }
    if (sh_s) { sh_s[3 * i] = h.fs.x; sh_s[3 * i + 1] = h.fs.y; sh_s[3 * i + 2] = h.fs.z; }
}

__global__ void __launch_bounds__(256) k_trace_occluded(DScene sc, const float *__restrict__ o, const float *__restrict__ d,
                                                         const float *__restrict__ tmax, uint64_t n, uint8_t *hit) {
    __shared__ DPrim sprims[MAX_SMEM_PRIMS];
    const DPrim *prims = stage_prims(sc, sprims);
    uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float3 oo = mk3(o[3 * i], o[3 * i + 1], o[3 * i + 2]), dd = mk3(d[3 * i], d[3 * i + 1], d[3 * i + 2]);
    hit[i] = occluded(sc, prims, oo, dd, tmax ? tmax[i] : PRT_INF) ? 1 : 0;
}

__global__ void k_ultra_bsdf(uint64_t n, const float *__restrict__ wi, const float *__restrict__ ng, const float *__restrict__ ns,
                             const float *__restrict__ Z, const float *__restrict__ rough, const float *__restrict__ s1,
                             const float *__restrict__ s2, float *dir, float *pdf, float *amp, int32_t *reflect) {
    uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float3 dd;
    float pf, am;
    bool rf;
    ultra_bsdf_sample(mk3(wi[3 * i], wi[3 * i + 1], wi[3 * i + 2]), mk3(ng[3 * i], ng[3 * i + 1], ng[3 * i + 2]),
                      mk3(ns[3 * i], ns[3 * i + 1], ns[3 * i + 2]), Z[i], rough[i], s1[i], s2[i], dd, pf, am, rf);
    dir[3 * i] = dd.x; dir[3 * i + 1] = dd.y; dir[3 * i + 2] = dd.z;
    pdf[i] = pf;
    amp[i] = am;
    reflect[i] = rf ? 1 : 0;
}

__global__ void k_directivity(uint64_t n, float3 nT, float cos_m, float cos_c, float alpha_m, float alpha_c, float n_rays,
                              const float *__restrict__ sec, const float *__restrict__ rd, const float *__restrict__ nrm,
                              float *w_i, float *w_o) {
    uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    w_i[i] = directivity_wi(nT, mk3(sec[3 * i], sec[3 * i + 1], sec[3 * i + 2]), cos_m, cos_c, alpha_m, alpha_c);
    w_o[i] = dot(mk3(rd[3 * i], rd[3 * i + 1], rd[3 * i + 2]), mk3(nrm[3 * i], nrm[3 * i + 1], nrm[3 * i + 2])) / n_rays;   // CI:117-118
}

// small RAII pool of per-call device buffers
struct DevBufs {
    std::vector<void *> ptrs;
    ~DevBufs() { for (void *p : ptrs) cudaFree(p); }
    template <typename T> int alloc(T **out, size_t count) {
        *out = nullptr;
        PRT_CUDA(cudaMalloc((void **) out, sizeof(T) * (count ? count : 1)));
        ptrs.push_back(*out);
        return PRT_OK;
    }
    template <typename T> int upload(T **out, const T *host, size_t count, cudaStream_t st) {
        int rc = alloc(out, count);
        if (rc) return rc;
        PRT_CUDA(cudaMemcpyAsync(*out, host, sizeof(T) * count, cudaMemcpyHostToDevice, st));
        return PRT_OK;
    }
};

}  // namespace prt

using namespace prt;

#define TRY(x) do { int _rc = (x); if (_rc) return _rc; } while (0)

extern "C" {

int prt_trace_closest(prt_scene *s, const float *o, const float *d, const float *tmax, uint64_t n, float *t, int32_t *prim,
                      int32_t *shape, float *p, float *ng, float *ns, float *wi, float *sh_s) {
    PRT_REQUIRE(s && o && d, "prt_trace_closest: null argument");
    if (!s->committed) { set_error("prt_trace_closest: scene not committed"); return PRT_ERR_STATE; }
    if (n == 0) return PRT_OK;
    std::lock_guard<std::mutex> lk(s->ctx->mtx);
    PRT_CUDA(cudaSetDevice(s->ctx->device));
    cudaStream_t st = s->ctx->stream;
    DevBufs b;
    float *o_d, *d_d, *tm_d = nullptr, *t_d, *p_d, *ng_d, *ns_d, *wi_d, *fs_d;
    int32_t *prim_d, *shape_d;
    TRY(b.upload(&o_d, o, 3 * n, st));
    TRY(b.upload(&d_d, d, 3 * n, st));
    if (tmax) TRY(b.upload(&tm_d, tmax, n, st));
    TRY(b.alloc(&t_d, n)); TRY(b.alloc(&prim_d, n)); TRY(b.alloc(&shape_d, n));
    TRY(b.alloc(&p_d, 3 * n)); TRY(b.alloc(&ng_d, 3 * n)); TRY(b.alloc(&ns_d, 3 * n)); TRY(b.alloc(&wi_d, 3 * n)); TRY(b.alloc(&fs_d, 3 * n));
    PRT_CUDA(cudaMemsetAsync(p_d, 0, sizeof(float) * 3 * n, st));
    PRT_CUDA(cudaMemsetAsync(ng_d, 0, sizeof(float) * 3 * n, st));
    PRT_CUDA(cudaMemsetAsync(ns_d, 0, sizeof(float) * 3 * n, st));
    PRT_CUDA(cudaMemsetAsync(wi_d, 0, sizeof(float) * 3 * n, st));
    PRT_CUDA(cudaMemsetAsync(fs_d, 0, sizeof(float) * 3 * n, st));
    k_trace_closest<<<(unsigned) ((n + 255) / 256), 256, 0, st>>>(s->view(), o_d, d_d, tm_d, n, t_d, prim_d, shape_d, p_d, ng_d, ns_d, wi_d, fs_d);
    PRT_CUDA(cudaGetLastError());
    if (t) PRT_CUDA(cudaMemcpyAsync(t, t_d, sizeof(float) * n, cudaMemcpyDeviceToHost, st));
    if (prim) PRT_CUDA(cudaMemcpyAsync(prim, prim_d, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, st));
    if (shape) PRT_CUDA(cudaMemcpyAsync(shape, shape_d, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, st));
    if (p) PRT_CUDA(cudaMemcpyAsync(p, p_d, sizeof(float) * 3 * n, cudaMemcpyDeviceToHost, st));
    if (ng) PRT_CUDA(cudaMemcpyAsync(ng, ng_d, sizeof(float) * 3 * n, cudaMemcpyDeviceToHost, st));
    if (ns) PRT_CUDA(cudaMemcpyAsync(ns, ns_d, sizeof(float) * 3 * n, cudaMemcpyDeviceToHost, st));
    if (wi) PRT_CUDA(cudaMemcpyAsync(wi, wi_d, sizeof(float) * 3 * n, cudaMemcpyDeviceToHost, st));
    if (sh_s) PRT_CUDA(cudaMemcpyAsync(sh_s, fs_d, sizeof(float) * 3 * n, cudaMemcpyDeviceToHost, st));
    PRT_CUDA(cudaStreamSynchronize(st));
    return PRT_OK;
}

int prt_trace_occluded(prt_scene *s, const float *o, const float *d, const float *tmax, uint64_t n, uint8_t *hit) {
    PRT_REQUIRE(s && o && d && hit, "prt_trace_occluded: null argument");
    if (!s->committed) { set_error("prt_trace_occluded: scene not committed"); return PRT_ERR_STATE; }
    if (n == 0) return PRT_OK;
    std::lock_guard<std::mutex> lk(s->ctx->mtx);
    PRT_CUDA(cudaSetDevice(s->ctx->device));
    cudaStream_t st = s->ctx->stream;
    DevBufs b;
    float *o_d, *d_d, *tm_d = nullptr;
    uint8_t *h_d;
    TRY(b.upload(&o_d, o, 3 * n, st));
    TRY(b.upload(&d_d, d, 3 * n, st));
    if (tmax) TRY(b.upload(&tm_d, tmax, n, st));
    TRY(b.alloc(&h_d, n));
    k_trace_occluded<<<(unsigned) ((n + 255) / 256), 256, 0, st>>>(s->view(), o_d, d_d, tm_d, n, h_d);
    PRT_CUDA(cudaGetLastError());
    PRT_CUDA(cudaMemcpyAsync(hit, h_d, n, cudaMemcpyDeviceToHost, st));
    PRT_CUDA(cudaStreamSynchronize(st));
    return PRT_OK;
}

int prt_ultra_bsdf_sample(prt_context *c, uint64_t n, const float *wi, const float *ng, const float *ns, const float *impedance,
                          const float *roughness, const float *s1, const float *s2, float *dir, float *pdf, float *amp,
                          int32_t *reflect) {
    PRT_REQUIRE(c && wi && ng && ns && impedance && roughness && s1 && s2 && dir && pdf && amp && reflect,
                "prt_ultra_bsdf_sample: null argument");
    if (n == 0) return PRT_OK;
    std::lock_guard<std::mutex> lk(c->mtx);
    PRT_CUDA(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    DevBufs b;
    float *wi_d, *ng_d, *ns_d, *z_d, *r_d, *s1_d, *s2_d, *dir_d, *pdf_d, *amp_d;
    int32_t *rf_d;
    TRY(b.upload(&wi_d, wi, 3 * n, st)); TRY(b.upload(&ng_d, ng, 3 * n, st)); TRY(b.upload(&ns_d, ns, 3 * n, st));
    TRY(b.upload(&z_d, impedance, n, st)); TRY(b.upload(&r_d, roughness, n, st));
    TRY(b.upload(&s1_d, s1, n, st)); TRY(b.upload(&s2_d, s2, n, st));
    TRY(b.alloc(&dir_d, 3 * n)); TRY(b.alloc(&pdf_d, n)); TRY(b.alloc(&amp_d, n)); TRY(b.alloc(&rf_d, n));
    k_ultra_bsdf<<<(unsigned) ((n + 255) / 256), 256, 0, st>>>(n, wi_d, ng_d, ns_d, z_d, r_d, s1_d, s2_d, dir_d, pdf_d, amp_d, rf_d);
    PRT_CUDA(cudaGetLastError());
    PRT_CUDA(cudaMemcpyAsync(dir, dir_d, sizeof(float) * 3 * n, cudaMemcpyDeviceToHost, st));
    PRT_CUDA(cudaMemcpyAsync(pdf, pdf_d, sizeof(float) * n, cudaMemcpyDeviceToHost, st));
    PRT_CUDA(cudaMemcpyAsync(amp, amp_d, sizeof(float) * n, cudaMemcpyDeviceToHost, st));
    PRT_CUDA(cudaMemcpyAsync(reflect, rf_d, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, st));
    PRT_CUDA(cudaStreamSynchronize(st));
    return PRT_OK;
}

int prt_directivity_weights(prt_context *c, uint64_t n, const double sensor_to_world[16], const float *sec_dir, const float *ray_dir,
                            const float *normal, double main_beam_deg, double cutoff_deg, double num_rays, float *w_i, float *w_o) {
    PRT_REQUIRE(c && sensor_to_world && sec_dir && ray_dir && normal && w_i && w_o, "prt_directivity_weights: null argument");
    if (n == 0) return PRT_OK;
    std::lock_guard<std::mutex> lk(c->mtx);
    PRT_CUDA(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    // the same derived constants as fill_params (prt_acquire.cu): n_T = normalize(T (0,0,1)), radians, cosines
    const float nx = (float) sensor_to_world[2], ny = (float) sensor_to_world[6], nz = (float) sensor_to_world[10];
    const float nl = sqrtf(nx * nx + ny * ny + nz * nz);
    const float3 nT = make_float3(nx / nl, ny / nl, nz / nl);
    const float am = (float) (main_beam_deg * M_PI / 180.0), ac = (float) (cutoff_deg * M_PI / 180.0);
    DevBufs b;
    float *s_d, *r_d, *n_d, *wi_d, *wo_d;
    TRY(b.upload(&s_d, sec_dir, 3 * n, st)); TRY(b.upload(&r_d, ray_dir, 3 * n, st)); TRY(b.upload(&n_d, normal, 3 * n, st));
    TRY(b.alloc(&wi_d, n)); TRY(b.alloc(&wo_d, n));
    k_directivity<<<(unsigned) ((n + 255) / 256), 256, 0, st>>>(n, nT, cosf(am), cosf(ac), am, ac, (float) num_rays, s_d, r_d, n_d, wi_d, wo_d);
    PRT_CUDA(cudaGetLastError());
    PRT_CUDA(cudaMemcpyAsync(w_i, wi_d, sizeof(float) * n, cudaMemcpyDeviceToHost, st));
    PRT_CUDA(cudaMemcpyAsync(w_o, wo_d, sizeof(float) * n, cudaMemcpyDeviceToHost, st));
    PRT_CUDA(cudaStreamSynchronize(st));
    return PRT_OK;
}

}  // extern "C"
