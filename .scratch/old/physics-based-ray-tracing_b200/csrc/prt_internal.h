// prt_internal.h -- host-side state behind the C ABI (include/prt_b200.h)
#pragma once
#include <cuda_runtime.h>

#include <mutex>
#include <string>
#include <vector>

#include "../../include/prt_b200.h"
#include "prt_device.cuh"

namespace prt {

void set_error(const std::string &msg);
int cuda_fail(cudaError_t e, const char *what, const char *file, int line);

#define PRT_CUDA(call)                                                       \
    do {                                                                     \
        cudaError_t _e = (call);                                             \
        if (_e != cudaSuccess) return prt::cuda_fail(_e, #call, __FILE__, __LINE__); \
    } while (0)

#define PRT_REQUIRE(cond, msg)               \
    do {                                     \
        if (!(cond)) {                       \
            prt::set_error(msg);             \
            return PRT_ERR_INVALID;          \
        }                                    \
    } while (0)

// N timing events that are destroyed on every exit path (the entry points return early on the first CUDA error)
template <int N>
struct ScopedEvents {
    cudaEvent_t e[N];
    bool ok = true;
    ScopedEvents() {
        for (int i = 0; i < N; i++) e[i] = nullptr;
        for (int i = 0; i < N && ok; i++) ok = cudaEventCreate(&e[i]) == cudaSuccess;
    }
    ~ScopedEvents() {
        for (int i = 0; i < N; i++)
            if (e[i]) cudaEventDestroy(e[i]);
    }
    ScopedEvents(const ScopedEvents &) = delete;
    ScopedEvents &operator=(const ScopedEvents &) = delete;
};

struct HostMesh {
    std::vector<float> v;   // [nt][3][3] world space
    std::vector<float> n;   // [nt][3][3] world-space corner normals (if has_n)
    bool has_n;
    int  shape, material, flip;
    uint32_t nt;
    double to_world[16];    // the transform v / n were built with (prt_scene_set_shape_transform applies new * old^-1)
};

// what a refit-only update needs from the build: the binary tree's topology (prt_bvh.cu keeps it alive for scenes of up to
// 2^22 triangles: 24 B per triangle)
struct LbvhTopology {
    int2 *children = nullptr, *ranges = nullptr;
    int *parent_internal = nullptr, *parent_leaf = nullptr, *visit = nullptr;
    float *root_box = nullptr;       // 8 floats (6 box + sah + pad)
    uint32_t n = 0;
};

}  // namespace prt

struct prt_context {
    int device;
    int sm_count;
    cudaDeviceProp prop;
    std::mutex mtx;
    cudaStream_t stream;          // library-owned stream for the host-buffer entry points
    cudaStream_t copy_stream;     // D2H of finished result slices, overlapped with the next launch
    cudaEvent_t  slice_done[2];
    // scratch for the acquisition / render entry points (grown on demand, reused across calls)
    float    *acc_dev;   size_t acc_cap;      // accumulator (channel_buf / film)
    float    *aux_dev;   size_t aux_cap;      // tx_delays etc.
    uint64_t *stats_dev;                      // 8 x u64
    double   *angles_dev; size_t angles_cap;
    // per-angle (sin, cos) tables of the acquisition, cached by CONTENT: a table is uploaded once, never overwritten while
    // the context lives, so asynchronous launches on different streams can never see each other's angles and the *_dev entry
    // points do no copy / stream synchronisation in steady state
    struct AngleTable { std::vector<float2> host; float2 *dev; };
    std::vector<AngleTable> angle_tables;
    void     *pinned;    size_t pinned_cap;   // pinned staging for D2H of results
    void     *wf_dev = nullptr; size_t wf_cap = 0;   // wavefront path-tracer state / queues (prt_wavefront.cu)
    int       last_launches = 0;              // kernels enqueued by the most recent render call
    // grow-only device scratch slots for the post-processing entry points (prt_das.cu): no cudaMalloc / cudaFree (and the
    // device-wide synchronisation cudaFree implies) per call
    struct Scratch { void *p = nullptr; size_t cap = 0; } scratch[8];
    // optional per-kernel-class timing (prt_profile_begin / prt_profile_read): CUDA event pairs on the launching stream
    bool prof_on = false;
    struct ProfPair { int cls, kernels; cudaEvent_t e0, e1; };
    std::vector<ProfPair> prof;
};

namespace prt {
// brackets the launches of one kernel class with an event pair while profiling is on (bench.py's roofline leg);
// a no-op otherwise
struct ProfScope {
    prt_context *c; cudaStream_t st; cudaEvent_t e0 = nullptr, e1 = nullptr; int cls, kernels = 1;
    ProfScope(prt_context *c_, int cls_, cudaStream_t st_) : c(c_), st(st_), cls(cls_) {
        if (!c->prof_on) return;
        if (cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess) { e0 = e1 = nullptr; return; }
        cudaEventRecord(e0, st);
    }
    ~ProfScope() {
        if (!e0) return;
        cudaEventRecord(e1, st);
        c->prof.push_back({cls, kernels, e0, e1});
    }
};
}  // namespace prt

struct prt_scene {
    prt_context *ctx;
    std::vector<prt::DPrim>     prims;
    std::vector<prt::DMaterial> mats;
    std::vector<prt::HostMesh>  meshes;
    int  n_shapes;
    bool committed;
    // device
    prt::DPrim     *prims_dev;
    prt::DMaterial *mats_dev;
    float4 *nodes_dev, *tri_v_dev, *tri_n_dev;
    float4 *nodes8_dev = nullptr, *tri_v8_dev = nullptr;   // compressed 8-wide BVH (prt_bvh8.cu)
    uint32_t *tri8_sorted_dev = nullptr;
    uint32_t n_nodes8 = 0;
    int bvh8_levels = 0;
    float bvh8_build_ms = 0.0f;
    int4   *tri_info_dev;
    float4 *em_tri_dev;
    int    *em_first_dev, *shape_emitter_dev;
    float  *em_inv_area_dev;
    int     n_emitters;
    uint32_t n_tris, n_nodes;
    prt::LbvhTopology topo;        // kept for prt_scene_set_shape_transform (empty: a transform change rebuilds)
    uint32_t n_small = 0;          // triangles in the hierarchy; the n_tris - n_small oversized ones sit behind them (DScene::n_small)
    int root_ref;
    uint64_t device_bytes;
    prt_bvh_stats stats;
    prt::DScene view() const;
};

namespace prt {
// prt_bvh.cu: builds the LBVH over `n` triangles given in INPUT order.
//   tri_v_in  [n][3] float4 world-space vertices (device), reordered into tri_v_out in sorted order
//   order_out [n] sorted position -> input index
struct Bvh8Out {
    float4 *nodes8;
    float4 *tri_v8;
    uint32_t *tri8_sorted;
    uint32_t n_nodes8;
    int levels;
    float build_ms;
};
// bvh8 != nullptr: also derive the compressed 8-wide BVH from the binary tree (buffers owned by the caller afterwards)
// keep != nullptr: the topology arrays are handed to the caller instead of being freed
int build_lbvh(prt_context *ctx, const float4 *tri_v_in, uint32_t n, float4 *tri_v_out, uint32_t *order_out,
               float4 *nodes_out, int *root_ref, prt_bvh_stats *stats, cudaStream_t stream, Bvh8Out *bvh8 = nullptr,
               LbvhTopology *keep = nullptr);
// new boxes for an unchanged topology after the (sorted) vertices moved: bottom-up refit + padding / child refs, then the
// 8-wide BVH is derived again from the refitted binary tree.  No Morton codes, no sort, no hierarchy emission.
int refit_lbvh(const LbvhTopology &t, const float4 *tri_v_sorted, float4 *nodes, prt_bvh_stats *stats, cudaStream_t st, Bvh8Out *bvh8);
void free_topology(LbvhTopology &t);
int bvh8_annotate(uint32_t n, const uint32_t *tri8_sorted, const int4 *tri_info, const DMaterial *mats, float4 *tri_v8, cudaStream_t st);
int build_bvh8(uint32_t n, const float4 *tri_v_sorted, const float *nodes2, const int2 *children, const int2 *ranges,
               float4 **out_nodes8, uint32_t *out_n_nodes8, float4 **out_tri_v8, uint32_t **out_tri8_sorted, int *out_levels,
               cudaStream_t st);
// exclusive prefix sum of n u32 in place; scratch: n / 2048 + 4096 words (prt_bvh.cu)
int exclusive_scan_u32(uint32_t *data, uint32_t n, uint32_t *scratch, cudaStream_t st);
int ensure_scratch(prt_context *ctx, size_t acc_floats, size_t aux_floats, size_t n_angles);
int scratch_slot(prt_context *ctx, int slot, size_t bytes, void **out);
int acquire_enqueue(prt_scene *s, const prt_acq_params *p, uint64_t seed, uint32_t spp_total, uint32_t sample_offset,
                    uint32_t sample_stride, int32_t angle_first, int32_t angle_count, float *channel_buf_dev, float *tx_delays_dev,
                    uint64_t *stats_dev, cudaStream_t st);   // grow-only; contents undefined after growth
}  // namespace prt
