// prt_bvh.cu -- GPU LBVH build for the triangle soup of a scene (north-star subsystem 2):
//   centroid bounds -> 63-bit Morton codes -> LSD radix sort (own kernels, 8-bit digits, stable
//   warp-match ranking) -> Karras 2012 hierarchy -> bottom-up AABB refit with atomic visit flags ->
//   leaf collapse (subtrees of <= 4 triangles become one leaf) -> SAH cost report.
// Replaces the Embree BVH build hidden inside mi.load_dict (/root/reference/USMain.py:257).
// Node layout: prt_device.cuh (64 B, both child boxes inline).
#include <cfloat>
#include <string>
#include <vector>

#include "prt_internal.h"

namespace prt {

#ifndef PRT_MAX_LEAF
#define PRT_MAX_LEAF 4
#endif
static constexpr int MAX_LEAF = PRT_MAX_LEAF;   // triangles per BVH2 leaf (the leaf ref encodes count - 1 in two bits)

// ---- order-preserving float <-> uint mapping for atomicMin/Max --------------------------------
__device__ __forceinline__ unsigned f2ord(float f) {
    unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

__global__ void k_init_bounds(unsigned *b) {
    if (threadIdx.x < 3) b[threadIdx.x] = 0xffffffffu;      // min
    else if (threadIdx.x < 6) b[threadIdx.x] = 0u;          // max
}

__global__ void k_centroid_bounds(const float4 *__restrict__ tv, uint32_t n, unsigned *bounds) {
    float lo[3] = { FLT_MAX, FLT_MAX, FLT_MAX }, hi[3] = { -FLT_MAX, -FLT_MAX, -FLT_MAX };
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float4 a = tv[3 * (size_t) i], b = tv[3 * (size_t) i + 1], c = tv[3 * (size_t) i + 2];
        float cx = (a.x + b.x + c.x) * (1.0f / 3.0f), cy = (a.y + b.y + c.y) * (1.0f / 3.0f), cz = (a.z + b.z + c.z) * (1.0f / 3.0f);
        lo[0] = fminf(lo[0], cx); lo[1] = fminf(lo[1], cy); lo[2] = fminf(lo[2], cz);
        hi[0] = fmaxf(hi[0], cx); hi[1] = fmaxf(hi[1], cy); hi[2] = fmaxf(hi[2], cz);
    }
#pragma unroll
    for (int a = 0; a < 3; a++) {
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            lo[a] = fminf(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
            hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
        }
    }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int a = 0; a < 3; a++) {
            atomicMin(&bounds[a], f2ord(lo[a]));
            atomicMax(&bounds[3 + a], f2ord(hi[a]));
        }
    }
}

__device__ __forceinline__ uint64_t expand21(uint64_t x) {
    x &= 0x1fffffull;
    x = (x | x << 32) & 0x1f00000000ffffull;
    x = (x | x << 16) & 0x1f0000ff0000ffull;
    x = (x | x << 8) & 0x100f00f00f00f00full;
    x = (x | x << 4) & 0x10c30c30c30c30c3ull;
    x = (x | x << 2) & 0x1249249249249249ull;
    return x;
}

__global__ void k_morton(const float4 *__restrict__ tv, uint32_t n, const unsigned *__restrict__ bounds, uint64_t *keys,
                         uint32_t *vals) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float lo[3], ext[3];
#pragma unroll
    for (int a = 0; a < 3; a++) {
        lo[a] = ord2f(bounds[a]);
        ext[a] = ord2f(bounds[3 + a]) - lo[a];
    }
    float4 a = tv[3 * (size_t) i], b = tv[3 * (size_t) i + 1], c = tv[3 * (size_t) i + 2];
    float cc[3] = { (a.x + b.x + c.x) * (1.0f / 3.0f), (a.y + b.y + c.y) * (1.0f / 3.0f), (a.z + b.z + c.z) * (1.0f / 3.0f) };
    // ONE scale for the three axes (cubic cells).  Normalising every axis by its own extent turns a thin slab of geometry --
    // a height field: 2 x 2 x 0.14 -- into a cube, so a third of the Morton bits split it by HEIGHT into layers that overlap
    // completely in the other two directions (measured: SAH cost 220 instead of 54 on the 10 M-triangle floor).
    const float emax = fmaxf(ext[0], fmaxf(ext[1], ext[2]));
    uint64_t q[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        float u = emax > 0.0f ? (cc[k] - lo[k]) / emax : 0.0f;
        u = fminf(fmaxf(u, 0.0f), 1.0f);
        q[k] = (uint64_t) fminf(u * 2097152.0f, 2097151.0f);
    }
    keys[i] = (expand21(q[0]) << 2) | (expand21(q[1]) << 1) | expand21(q[2]);
    vals[i] = i;
}

// ---- radix sort (u64 key, u32 value), 8 passes of 8 bits -----------------------------------------
static constexpr int RS_THREADS = 256, RS_ITEMS = 8, RS_TILE = RS_THREADS * RS_ITEMS;

__global__ void __launch_bounds__(RS_THREADS) k_rs_hist(const uint64_t *__restrict__ keys, uint32_t n, int shift,
                                                        uint32_t *__restrict__ hist, uint32_t nblocks) {
    __shared__ uint32_t h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    uint32_t base = blockIdx.x * RS_TILE;
#pragma unroll
    for (int i = 0; i < RS_ITEMS; i++) {
        uint32_t idx = base + i * RS_THREADS + threadIdx.x;
        if (idx < n) atomicAdd(&h[(uint32_t) (keys[idx] >> shift) & 255u], 1u);
    }
    __syncthreads();
    hist[threadIdx.x * nblocks + blockIdx.x] = h[threadIdx.x];
}

__global__ void __launch_bounds__(RS_THREADS) k_rs_scatter(const uint64_t *__restrict__ keys_in,
                                                           const uint32_t *__restrict__ vals_in, uint64_t *__restrict__ keys_out,
                                                           uint32_t *__restrict__ vals_out, uint32_t n, int shift,
                                                           const uint32_t *__restrict__ hist, uint32_t nblocks) {
    __shared__ uint32_t wh[RS_THREADS / 32][256];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int j = lane; j < 256; j += 32) wh[w][j] = 0;
    __syncwarp();
    const uint32_t base = blockIdx.x * RS_TILE + w * (RS_ITEMS * 32);
    uint64_t key[RS_ITEMS];
    uint32_t off[RS_ITEMS];
#pragma unroll
    for (int i = 0; i < RS_ITEMS; i++) {
        uint32_t idx = base + i * 32 + lane;
        bool ok = idx < n;
        key[i] = ok ? keys_in[idx] : ~0ull;
        uint32_t digit = ok ? ((uint32_t) (key[i] >> shift) & 255u) : 256u;
        uint32_t peers = __match_any_sync(0xffffffffu, digit);
        uint32_t rank = __popc(peers & ((1u << lane) - 1u));
        int leader = __ffs(peers) - 1;
        uint32_t pre = 0;
        if (ok && lane == leader) {
            pre = wh[w][digit];
            wh[w][digit] = pre + __popc(peers);
        }
        pre = __shfl_sync(0xffffffffu, pre, leader);
        off[i] = pre + rank;
        __syncwarp();
    }
    __syncthreads();
    {
        uint32_t d = threadIdx.x;
        uint32_t running = hist[d * nblocks + blockIdx.x];
#pragma unroll
        for (int ww = 0; ww < RS_THREADS / 32; ww++) {
            uint32_t c = wh[ww][d];
            wh[ww][d] = running;
            running += c;
        }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < RS_ITEMS; i++) {
        uint32_t idx = base + i * 32 + lane;
        if (idx < n) {
            uint32_t digit = (uint32_t) (key[i] >> shift) & 255u;
            uint32_t pos = wh[w][digit] + off[i];
            keys_out[pos] = key[i];
            vals_out[pos] = vals_in[idx];
        }
    }
}

// ---- exclusive scan (u32), 2048 elements per block, recursive over block sums ---------------------
static constexpr int SC_THREADS = 256, SC_ITEMS = 8, SC_TILE = SC_THREADS * SC_ITEMS;

__global__ void __launch_bounds__(SC_THREADS) k_scan_block(uint32_t *data, uint32_t n, uint32_t *sums) {
    __shared__ uint32_t warp_sums[SC_THREADS / 32];
    uint32_t base = blockIdx.x * SC_TILE + threadIdx.x * SC_ITEMS;
    uint32_t v[SC_ITEMS], total = 0;
#pragma unroll
    for (int i = 0; i < SC_ITEMS; i++) {
        v[i] = (base + i < n) ? data[base + i] : 0u;
        total += v[i];
    }
    uint32_t incl = total;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) warp_sums[w] = incl;
    __syncthreads();
    if (w == 0) {
        uint32_t s = lane < SC_THREADS / 32 ? warp_sums[lane] : 0u;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, s, o);
            if (lane >= o) s += t;
        }
        if (lane < SC_THREADS / 32) warp_sums[lane] = s;
    }
    __syncthreads();
    uint32_t excl = incl - total + (w ? warp_sums[w - 1] : 0u);
#pragma unroll
    for (int i = 0; i < SC_ITEMS; i++) {
        if (base + i < n) data[base + i] = excl;
        excl += v[i];
    }
    if (threadIdx.x == SC_THREADS - 1 && sums) sums[blockIdx.x] = excl;
}

__global__ void k_scan_add(uint32_t *data, uint32_t n, const uint32_t *__restrict__ sums) {
    uint32_t i = blockIdx.x * SC_TILE + threadIdx.x;
    uint32_t add = sums[blockIdx.x];
#pragma unroll
    for (int k = 0; k < SC_ITEMS; k++, i += SC_THREADS)
        if (i < n) data[i] += add;
}

static int exclusive_scan(uint32_t *data, uint32_t n, uint32_t *scratch, cudaStream_t st) {
    uint32_t nb = (n + SC_TILE - 1) / SC_TILE;
    if (nb <= 1) {
        k_scan_block<<<1, SC_THREADS, 0, st>>>(data, n, nullptr);
        return 0;
    }
    k_scan_block<<<nb, SC_THREADS, 0, st>>>(data, n, scratch);
    exclusive_scan(scratch, nb, scratch + nb, st);
    k_scan_add<<<nb, SC_THREADS, 0, st>>>(data, n, scratch);
    return 0;
}

int exclusive_scan_u32(uint32_t *data, uint32_t n, uint32_t *scratch, cudaStream_t st) { return exclusive_scan(data, n, scratch, st); }

// ---- Karras 2012 ----------------------------------------------------------------------------------
__device__ __forceinline__ int delta(const uint64_t *__restrict__ keys, int n, int i, int j) {
    if (j < 0 || j >= n) return -1;
    uint64_t a = keys[i], b = keys[j];
    if (a == b) return 64 + __clz((unsigned) i ^ (unsigned) j);
    return __clzll((long long) (a ^ b));
}

// child encoding during the build: >= 0 internal node, < 0 leaf ~sorted_index
__global__ void k_karras(const uint64_t *__restrict__ keys, int n, int2 *children, int2 *ranges, int *parent_internal,
                         int *parent_leaf) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    int d = (delta(keys, n, i, i + 1) - delta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
    int dmin = delta(keys, n, i, i - d);
    int lmax = 2;
    while (delta(keys, n, i, i + lmax * d) > dmin) lmax <<= 1;
    int l = 0;
    for (int t = lmax >> 1; t >= 1; t >>= 1)
        if (delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
    int j = i + l * d;
    int dnode = delta(keys, n, i, j);
    int s = 0, t = l;
    do {
        t = (t + 1) >> 1;
        if (delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
    } while (t > 1);
    int gamma = i + s * d + min(d, 0);
    int lo = min(i, j), hi = max(i, j);
    int left = (lo == gamma) ? ~gamma : gamma;
    int right = (hi == gamma + 1) ? ~(gamma + 1) : gamma + 1;
    children[i] = make_int2(left, right);
    ranges[i] = make_int2(lo, hi);
    if (left >= 0) parent_internal[left] = i; else parent_leaf[~left] = i;
    if (right >= 0) parent_internal[right] = i; else parent_leaf[~right] = i;
    if (i == 0) parent_internal[0] = -1;
}

__global__ void k_gather_tris(const float4 *__restrict__ tv_in, const uint32_t *__restrict__ order, uint32_t n, float4 *tv_out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t src = order[i];
    float4 a = tv_in[3 * (size_t) src], b = tv_in[3 * (size_t) src + 1], c = tv_in[3 * (size_t) src + 2];
    a.w = __uint_as_float(src);
    tv_out[3 * (size_t) i] = a;
    tv_out[3 * (size_t) i + 1] = b;
    tv_out[3 * (size_t) i + 2] = c;
}

// bottom-up refit: every child writes its box into its half of the parent's node record; the second
// arrival (atomic flag) unions both halves and continues upward.
__global__ void k_refit(const float4 *__restrict__ tv_sorted, int n, const int2 *__restrict__ children,
                        const int *__restrict__ parent_internal, const int *__restrict__ parent_leaf, float *nodes /*16 f / node*/,
                        int *visit, float *root_box) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 a = tv_sorted[3 * (size_t) i], b = tv_sorted[3 * (size_t) i + 1], c = tv_sorted[3 * (size_t) i + 2];
    float lo[3] = { fminf(a.x, fminf(b.x, c.x)), fminf(a.y, fminf(b.y, c.y)), fminf(a.z, fminf(b.z, c.z)) };
    float hi[3] = { fmaxf(a.x, fmaxf(b.x, c.x)), fmaxf(a.y, fmaxf(b.y, c.y)), fmaxf(a.z, fmaxf(b.z, c.z)) };
    int me = ~i;
    int p = parent_leaf[i];
    while (p >= 0) {
        int2 ch = children[p];
        float *rec = nodes + 16 * (size_t) p;
        int side = (ch.x == me) ? 0 : 6;
#pragma unroll
        for (int k = 0; k < 3; k++) {
            rec[side + k] = lo[k];
            rec[side + 3 + k] = hi[k];
        }
        __threadfence();
        if (atomicAdd(&visit[p], 1) == 0) return;  // sibling not there yet
        int other = 6 - side;
        volatile float *vr = rec;
#pragma unroll
        for (int k = 0; k < 3; k++) {
            lo[k] = fminf(lo[k], vr[other + k]);
            hi[k] = fmaxf(hi[k], vr[other + 3 + k]);
        }
        me = p;
        p = parent_internal[p];
    }
    if (root_box) {
#pragma unroll
        for (int k = 0; k < 3; k++) {
            root_box[k] = lo[k];
            root_box[3 + k] = hi[k];
        }
    }
}

// final pass: pad boxes by a few ulps (conservative traversal), write child refs with leaf collapse
__global__ void k_emit(int n, const int2 *__restrict__ children, const int2 *__restrict__ ranges, float *nodes) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    float *rec = nodes + 16 * (size_t) i;
#pragma unroll
    for (int s = 0; s < 12; s += 6) {
#pragma unroll
        for (int k = 0; k < 3; k++) {
            float lo = rec[s + k], hi = rec[s + 3 + k];
            float pad = 4.0f * 1.1920929e-7f * fmaxf(fmaxf(fabsf(lo), fabsf(hi)), hi - lo) + 1e-30f;
            rec[s + k] = lo - pad;
            rec[s + 3 + k] = hi + pad;
        }
    }
    int2 ch = children[i];
    int refs[2] = { ch.x, ch.y };
#pragma unroll
    for (int s = 0; s < 2; s++) {
        int c = refs[s];
        if (c < 0) {
            refs[s] = ~((~c) << 2);  // single-triangle leaf
        } else {
            int2 r = ranges[c];
            int cnt = r.y - r.x + 1;
            if (cnt <= MAX_LEAF) refs[s] = ~((r.x << 2) | (cnt - 1));
        }
    }
    rec[12] = __int_as_float(refs[0]);
    rec[13] = __int_as_float(refs[1]);
    rec[14] = 0.0f;
    rec[15] = 0.0f;
}

// SAH cost over the REACHABLE tree: sum_internal A(child boxes)/A(root) * 1.2 + sum_leaves A/A(root) * count
__global__ void k_sah(int n, const float *__restrict__ nodes, const int2 *__restrict__ ranges, const int *__restrict__ parent_internal,
                      const float *__restrict__ root_box, float *out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    float cost = 0.0f;
    if (i < n - 1) {
        // a node is reachable iff neither it nor any ancestor was collapsed, i.e. its own range is > MAX_LEAF
        int2 r = ranges[i];
        bool reachable = (r.y - r.x + 1) > MAX_LEAF;
        if (reachable) {
            const float *rec = nodes + 16 * (size_t) i;
            float rx = root_box[3] - root_box[0], ry = root_box[4] - root_box[1], rz = root_box[5] - root_box[2];
            float ra = fmaxf(2.0f * (rx * ry + ry * rz + rz * rx), 1e-30f);
#pragma unroll
            for (int s = 0; s < 2; s++) {
                float ex = rec[6 * s + 3] - rec[6 * s], ey = rec[6 * s + 4] - rec[6 * s + 1], ez = rec[6 * s + 5] - rec[6 * s + 2];
                float a = 2.0f * (ex * ey + ey * ez + ez * ex) / ra;
                int ref = __float_as_int(rec[12 + s]);
                cost += ref >= 0 ? 1.2f * a : a * (float) (((~ref) & 3) + 1);
            }
        }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) cost += __shfl_xor_sync(0xffffffffu, cost, o);
    if ((threadIdx.x & 31) == 0 && cost != 0.0f) atomicAdd(out, cost);
    (void) parent_internal;
}

// depth of the binary tree = the longest leaf-to-root parent chain (the traversal stack must hold one entry per level)
__global__ void k_depth(int n, const int *__restrict__ parent_internal, const int *__restrict__ parent_leaf, int *max_depth) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int d = 0;
    if (i < n)
        for (int p = parent_leaf[i]; p >= 0; p = parent_internal[p]) d++;
#pragma unroll
    for (int o = 16; o; o >>= 1) d = max(d, __shfl_xor_sync(0xffffffffu, d, o));
    if ((threadIdx.x & 31) == 0 && d) atomicMax(max_depth, d);
}

namespace {
// every temporary of the build: released on EVERY exit path (the PRT_CUDA macro returns on the first error)
struct BuildTemps {
    std::vector<void *> ptrs;
    std::vector<cudaEvent_t> events;
    ~BuildTemps() {
        for (void *p : ptrs) cudaFree(p);
        for (cudaEvent_t e : events) cudaEventDestroy(e);
    }
    template <typename T> cudaError_t alloc(T **out, size_t count) {
        *out = nullptr;
        cudaError_t e = cudaMalloc((void **) out, sizeof(T) * (count ? count : 1));
        if (e == cudaSuccess) ptrs.push_back(*out);
        return e;
    }
    cudaError_t event(cudaEvent_t *e) {
        cudaError_t rc = cudaEventCreate(e);
        if (rc == cudaSuccess) events.push_back(*e);
        return rc;
    }
};
}  // namespace

int build_lbvh(prt_context *ctx, const float4 *tri_v_in, uint32_t n, float4 *tri_v_out, uint32_t *order_out,
               float4 *nodes_out, int *root_ref, prt_bvh_stats *stats, cudaStream_t st, Bvh8Out *bvh8, LbvhTopology *keep) {
    (void) ctx;
    if (keep) *keep = LbvhTopology();
    const uint32_t n_extra = bvh8 ? bvh8->n_extra : 0u;
    if (bvh8) *bvh8 = Bvh8Out{ nullptr, nullptr, nullptr, 0, 0, 0.0f, n_extra };
    if (n == 0) {
        *root_ref = -1;
        return PRT_OK;
    }
    BuildTemps tmp;
    cudaEvent_t e0, e1;
    PRT_CUDA(tmp.event(&e0));
    PRT_CUDA(tmp.event(&e1));
    const uint32_t nb_sort = (n + RS_TILE - 1) / RS_TILE;
    const uint32_t hist_len = 256 * nb_sort;
    uint64_t *keys[2];
    uint32_t *vals[2], *hist, *scan_scratch;
    unsigned *bounds;
    int2 *children, *ranges;
    int *parent_internal, *parent_leaf, *visit, *depth_dev;
    float *root_box, *sah;
    PRT_CUDA(tmp.alloc(&keys[0], n));
    PRT_CUDA(tmp.alloc(&keys[1], n));
    PRT_CUDA(tmp.alloc(&vals[0], n));
    PRT_CUDA(tmp.alloc(&vals[1], n));
    PRT_CUDA(tmp.alloc(&hist, hist_len));
    PRT_CUDA(tmp.alloc(&scan_scratch, hist_len / SC_TILE + 4096));
    PRT_CUDA(tmp.alloc(&bounds, 8));
    PRT_CUDA(tmp.alloc(&children, n));
    PRT_CUDA(tmp.alloc(&ranges, n));
    PRT_CUDA(tmp.alloc(&parent_internal, n));
    PRT_CUDA(tmp.alloc(&parent_leaf, n));
    PRT_CUDA(tmp.alloc(&visit, n));
    PRT_CUDA(tmp.alloc(&root_box, 8));
    PRT_CUDA(tmp.alloc(&depth_dev, 1));
    sah = root_box + 6;
    PRT_CUDA(cudaEventRecord(e0, st));
    const int T = 256;
    const uint32_t nb = (n + T - 1) / T;
    k_init_bounds<<<1, 32, 0, st>>>(bounds);
    k_centroid_bounds<<<min(nb, 148u * 8u), T, 0, st>>>(tri_v_in, n, bounds);
    k_morton<<<nb, T, 0, st>>>(tri_v_in, n, bounds, keys[0], vals[0]);
    int cur = 0;
    for (int pass = 0; pass < 8; pass++) {
        int shift = 8 * pass;
        k_rs_hist<<<nb_sort, RS_THREADS, 0, st>>>(keys[cur], n, shift, hist, nb_sort);
        exclusive_scan(hist, hist_len, scan_scratch, st);
        k_rs_scatter<<<nb_sort, RS_THREADS, 0, st>>>(keys[cur], vals[cur], keys[cur ^ 1], vals[cur ^ 1], n, shift, hist, nb_sort);
        cur ^= 1;
    }
    PRT_CUDA(cudaMemcpyAsync(order_out, vals[cur], sizeof(uint32_t) * n, cudaMemcpyDeviceToDevice, st));
    k_gather_tris<<<nb, T, 0, st>>>(tri_v_in, vals[cur], n, tri_v_out);
    PRT_CUDA(cudaMemsetAsync(root_box, 0, sizeof(float) * 8, st));
    if (n >= 2) {
        PRT_CUDA(cudaMemsetAsync(visit, 0, sizeof(int) * n, st));
        k_karras<<<nb, T, 0, st>>>(keys[cur], (int) n, children, ranges, parent_internal, parent_leaf);
        k_refit<<<nb, T, 0, st>>>(tri_v_out, (int) n, children, parent_internal, parent_leaf, (float *) nodes_out, visit, root_box);
        k_emit<<<nb, T, 0, st>>>((int) n, children, ranges, (float *) nodes_out);
        k_sah<<<nb, T, 0, st>>>((int) n, (const float *) nodes_out, ranges, parent_internal, root_box, sah);
        PRT_CUDA(cudaMemsetAsync(depth_dev, 0, sizeof(int), st));
        k_depth<<<nb, T, 0, st>>>((int) n, parent_internal, parent_leaf, depth_dev);
    }
    PRT_CUDA(cudaEventRecord(e1, st));
    PRT_CUDA(cudaStreamSynchronize(st));
    PRT_CUDA(cudaGetLastError());
    if (n >= 2) {
        // The binary traversal (traverse_bvh) pushes at most one entry per level and silently drops entries beyond PRT_STACK:
        // a deeper tree would lose hits without any error.  63-bit Morton codes + index tie-break bound the depth by
        // 63 + log2(duplicates), which only pathological inputs (tens of thousands of coincident centroids) approach.
        int depth = 0;
        PRT_CUDA(cudaMemcpy(&depth, depth_dev, sizeof(int), cudaMemcpyDeviceToHost));
        if (depth > PRT_STACK) {
            set_error("prt_scene_commit: the LBVH is " + std::to_string(depth) + " levels deep; the traversal stack holds " +
                      std::to_string(PRT_STACK) + " (coincident triangles?)");
            return PRT_ERR_UNSUPPORTED;
        }
    }
    if (bvh8) {
        cudaEvent_t b0, b1;
        PRT_CUDA(tmp.event(&b0));
        PRT_CUDA(tmp.event(&b1));
        PRT_CUDA(cudaEventRecord(b0, st));
        int rc = build_bvh8(n, tri_v_out, (const float *) nodes_out, children, ranges, &bvh8->nodes8, &bvh8->n_nodes8, &bvh8->tri_v8,
                            &bvh8->tri8_sorted, &bvh8->levels, st, n_extra);
        if (rc) return rc;
        PRT_CUDA(cudaEventRecord(b1, st));
        PRT_CUDA(cudaStreamSynchronize(st));
        PRT_CUDA(cudaEventElapsedTime(&bvh8->build_ms, b0, b1));
    }
    *root_ref = (n <= (uint32_t) MAX_LEAF) ? ~(int) ((0u << 2) | (n - 1)) : 0;
    if (stats) {
        float ms = 0.0f, hb[8];
        PRT_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        PRT_CUDA(cudaMemcpy(hb, root_box, sizeof(float) * 8, cudaMemcpyDeviceToHost));
        stats->build_ms = ms;
        stats->sah_cost = hb[6];
        for (int k = 0; k < 3; k++) {
            stats->scene_lo[k] = hb[k];
            stats->scene_hi[k] = hb[3 + k];
        }
        stats->n_nodes = n >= 2 ? n - 1 : 0;
        stats->max_leaf_size = MAX_LEAF;
    }
    if (keep && n >= 2) {       // hand the topology over: take these six out of the temporaries' list
        void *mine[6] = { children, ranges, parent_internal, parent_leaf, visit, root_box };
        for (void *m : mine)
            for (auto it = tmp.ptrs.begin(); it != tmp.ptrs.end(); ++it)
                if (*it == m) { tmp.ptrs.erase(it); break; }
        keep->children = children; keep->ranges = ranges; keep->parent_internal = parent_internal; keep->parent_leaf = parent_leaf;
        keep->visit = visit; keep->root_box = root_box; keep->n = n;
    }
    return PRT_OK;      // ~BuildTemps releases every (other) temporary
}

void free_topology(LbvhTopology &t) {
    cudaFree(t.children); cudaFree(t.ranges); cudaFree(t.parent_internal); cudaFree(t.parent_leaf); cudaFree(t.visit); cudaFree(t.root_box);
    t = LbvhTopology();
}

int refit_lbvh(const LbvhTopology &t, const float4 *tri_v_sorted, float4 *nodes, prt_bvh_stats *stats, cudaStream_t st, Bvh8Out *bvh8) {
    const uint32_t n = t.n;
    const uint32_t n_extra = bvh8 ? bvh8->n_extra : 0u;
    if (bvh8) *bvh8 = Bvh8Out{ nullptr, nullptr, nullptr, 0, 0, 0.0f, n_extra };
    PRT_REQUIRE(n >= 2 && t.children, "refit_lbvh: no topology was kept for this scene");
    BuildTemps tmp;
    cudaEvent_t e0, e1;
    PRT_CUDA(tmp.event(&e0));
    PRT_CUDA(tmp.event(&e1));
    const int T = 256;
    const uint32_t nb = (n + T - 1) / T;
    PRT_CUDA(cudaEventRecord(e0, st));
    PRT_CUDA(cudaMemsetAsync(t.visit, 0, sizeof(int) * n, st));
    PRT_CUDA(cudaMemsetAsync(t.root_box, 0, sizeof(float) * 8, st));
    k_refit<<<nb, T, 0, st>>>(tri_v_sorted, (int) n, t.children, t.parent_internal, t.parent_leaf, (float *) nodes, t.visit, t.root_box);
    k_emit<<<nb, T, 0, st>>>((int) n, t.children, t.ranges, (float *) nodes);
    k_sah<<<nb, T, 0, st>>>((int) n, (const float *) nodes, t.ranges, t.parent_internal, t.root_box, t.root_box + 6);
    PRT_CUDA(cudaGetLastError());
    if (bvh8) {
        int rc = build_bvh8(n, tri_v_sorted, (const float *) nodes, t.children, t.ranges, &bvh8->nodes8, &bvh8->n_nodes8, &bvh8->tri_v8,
                            &bvh8->tri8_sorted, &bvh8->levels, st, n_extra);
        if (rc) return rc;
    }
    PRT_CUDA(cudaEventRecord(e1, st));
    PRT_CUDA(cudaStreamSynchronize(st));
    if (stats) {
        float ms = 0.0f, hb[8];
        PRT_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        PRT_CUDA(cudaMemcpy(hb, t.root_box, sizeof(float) * 8, cudaMemcpyDeviceToHost));
        stats->build_ms = ms;           // the refit's time replaces the build's
        stats->sah_cost = hb[6];
        for (int k = 0; k < 3; k++) { stats->scene_lo[k] = hb[k]; stats->scene_hi[k] = hb[3 + k]; }
    }
    return PRT_OK;
}

}  // namespace prt
