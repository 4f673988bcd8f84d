// Non-overlapping max pooling (kernel = stride, no padding, optional ceil mode), forward and backward.
// Replaces nn.MaxPool2d on the residual branch of ScalogramEncoderBlock (scalogram_model.py:434-441,
// MaxPool2d(kernel_size = stride_pool, ceil_mode = True)) and the main-path poolings (:402-403, :424-425),
// plus their autograd (max_pool2d_with_indices_backward).  No index tensor is kept: the backward pass
// re-derives the arg-max from x (first maximum in row-major window order, the tie rule of the reference's
// ATen kernels), so it streams x and dy once and writes dx once.
#include "common.cuh"

namespace cpc {

struct PoolGeom {
    int H, W, OH, OW, k;
    FastDiv d_ow;
};

// one thread per output window; grid (B*C planes, window chunks of one plane)
__global__ void __launch_bounds__(256) maxpool_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, PoolGeom g) {
    const int plane = blockIdx.x;
    const int idx = blockIdx.y * blockDim.x + threadIdx.x;
    if (idx >= g.OH * g.OW) return;
    int oh, ow;
    g.d_ow.divmod(idx, oh, ow);
    const float* px = x + (size_t)plane * g.H * g.W;
    const int h0 = oh * g.k, w0 = ow * g.k;
    const int h1 = min(h0 + g.k, g.H), w1 = min(w0 + g.k, g.W);
    float m = -INFINITY;
    for (int h = h0; h < h1; ++h)
        for (int w = w0; w < w1; ++w) {
            const float v = __ldg(px + (size_t)h * g.W + w);
            if (v > m || v != v) m = v;
        }
    y[(size_t)plane * g.OH * g.OW + idx] = m;
}

__global__ void __launch_bounds__(256) maxpool_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                         float* __restrict__ dx, PoolGeom g) {
    const int plane = blockIdx.x;
    const int idx = blockIdx.y * blockDim.x + threadIdx.x;
    if (idx >= g.OH * g.OW) return;
    int oh, ow;
    g.d_ow.divmod(idx, oh, ow);
    const float* px = x + (size_t)plane * g.H * g.W;
    float* pdx = dx + (size_t)plane * g.H * g.W;
    const int h0 = oh * g.k, w0 = ow * g.k;
    const int h1 = min(h0 + g.k, g.H), w1 = min(w0 + g.k, g.W);
    float m = -INFINITY;
    int ah = h0, aw = w0;
    for (int h = h0; h < h1; ++h)
        for (int w = w0; w < w1; ++w) {
            const float v = __ldg(px + (size_t)h * g.W + w);
            if (v > m || v != v) { m = v; ah = h; aw = w; }
        }
    const float gval = __ldg(dy + (size_t)plane * g.OH * g.OW + idx);
    for (int h = h0; h < h1; ++h)
        for (int w = w0; w < w1; ++w) pdx[(size_t)h * g.W + w] = (h == ah && w == aw) ? gval : 0.f;
}

// ---- 2 x 2 windows on even-width planes (every pooling of the arch-7 encoder): a window's two rows are one 8-byte
// load each, a thread owns POOL2_PER_THREAD windows 256 apart and issues all their loads before any compare / store
// (the generic kernel's four dependent 4-byte loads per thread ran at ~3 TB/s).
constexpr int POOL2_PER_THREAD = 4;

__device__ __forceinline__ void pool2_argmax(const float2 r0, const float2 r1, bool has_r1, float& m, int& arg) {
    // first maximum in row-major order; NaN propagates like the reference's ATen kernel (v > m || isnan(v))
    m = r0.x; arg = 0;
    if (r0.y > m || r0.y != r0.y) { m = r0.y; arg = 1; }
    if (has_r1) {
        if (r1.x > m || r1.x != r1.x) { m = r1.x; arg = 2; }
        if (r1.y > m || r1.y != r1.y) { m = r1.y; arg = 3; }
    }
}

__global__ void __launch_bounds__(256) maxpool2_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, PoolGeom g) {
    const int plane = blockIdx.x;
    const float* px = x + (size_t)plane * g.H * g.W;
    float* py = y + (size_t)plane * g.OH * g.OW;
    const int n = g.OH * g.OW;
    const int i0 = blockIdx.y * (256 * POOL2_PER_THREAD) + threadIdx.x;
    float2 r0[POOL2_PER_THREAD], r1[POOL2_PER_THREAD];
    bool two[POOL2_PER_THREAD];
#pragma unroll
    for (int u = 0; u < POOL2_PER_THREAD; ++u) {
        const int idx = i0 + u * 256;
        r0[u] = make_float2(0.f, 0.f); r1[u] = r0[u]; two[u] = false;
        if (idx < n) {
            int oh, ow;
            g.d_ow.divmod(idx, oh, ow);
            const float* p0 = px + (size_t)(2 * oh) * g.W + 2 * ow;
            r0[u] = __ldg(reinterpret_cast<const float2*>(p0));
            two[u] = 2 * oh + 1 < g.H;
            if (two[u]) r1[u] = __ldg(reinterpret_cast<const float2*>(p0 + g.W));
        }
    }
#pragma unroll
    for (int u = 0; u < POOL2_PER_THREAD; ++u) {
        const int idx = i0 + u * 256;
        if (idx < n) {
            float m; int arg;
            pool2_argmax(r0[u], r1[u], two[u], m, arg);
            py[idx] = m;
        }
    }
}

__global__ void __launch_bounds__(256) maxpool2_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                          float* __restrict__ dx, PoolGeom g) {
    const int plane = blockIdx.x;
    const float* px = x + (size_t)plane * g.H * g.W;
    float* pdx = dx + (size_t)plane * g.H * g.W;
    const float* pdy = dy + (size_t)plane * g.OH * g.OW;
    const int n = g.OH * g.OW;
    const int i0 = blockIdx.y * (256 * POOL2_PER_THREAD) + threadIdx.x;
    float2 r0[POOL2_PER_THREAD], r1[POOL2_PER_THREAD];
    float gv[POOL2_PER_THREAD];
    int off[POOL2_PER_THREAD];                                   // element offset of the window origin, -1: none
    bool two[POOL2_PER_THREAD];
#pragma unroll
    for (int u = 0; u < POOL2_PER_THREAD; ++u) {
        const int idx = i0 + u * 256;
        r0[u] = make_float2(0.f, 0.f); r1[u] = r0[u]; two[u] = false; off[u] = -1; gv[u] = 0.f;
        if (idx < n) {
            int oh, ow;
            g.d_ow.divmod(idx, oh, ow);
            off[u] = (2 * oh) * g.W + 2 * ow;
            r0[u] = __ldg(reinterpret_cast<const float2*>(px + off[u]));
            two[u] = 2 * oh + 1 < g.H;
            if (two[u]) r1[u] = __ldg(reinterpret_cast<const float2*>(px + off[u] + g.W));
            gv[u] = __ldg(pdy + idx);
        }
    }
#pragma unroll
    for (int u = 0; u < POOL2_PER_THREAD; ++u) {
        if (off[u] >= 0) {
            float m; int arg;
            pool2_argmax(r0[u], r1[u], two[u], m, arg);
            *reinterpret_cast<float2*>(pdx + off[u]) = make_float2(arg == 0 ? gv[u] : 0.f, arg == 1 ? gv[u] : 0.f);
            if (two[u])
                *reinterpret_cast<float2*>(pdx + off[u] + g.W) = make_float2(arg == 2 ? gv[u] : 0.f, arg == 3 ? gv[u] : 0.f);
        }
    }
}

// dx += max-pool gradient (dx already holds another branch's gradient of the same tensor): the window's dx pairs travel
// with its x pairs (all loads of a thread issued up front), the arg-max element is incremented and the pairs are
// written back -- no zero fill, no separate add pass over the tensor.
__global__ void __launch_bounds__(256) maxpool2_bwd_acc_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                              float* __restrict__ dx, PoolGeom g) {
    const int plane = blockIdx.x;
    const float* px = x + (size_t)plane * g.H * g.W;
    float* pdx = dx + (size_t)plane * g.H * g.W;
    const float* pdy = dy + (size_t)plane * g.OH * g.OW;
    const int n = g.OH * g.OW;
    const int i0 = blockIdx.y * (256 * POOL2_PER_THREAD) + threadIdx.x;
    float2 r0[POOL2_PER_THREAD], r1[POOL2_PER_THREAD], d0[POOL2_PER_THREAD], d1[POOL2_PER_THREAD];
    float gv[POOL2_PER_THREAD];
    int off[POOL2_PER_THREAD];
    bool two[POOL2_PER_THREAD];
#pragma unroll
    for (int u = 0; u < POOL2_PER_THREAD; ++u) {
        const int idx = i0 + u * 256;
        r0[u] = make_float2(0.f, 0.f); r1[u] = r0[u]; d0[u] = r0[u]; d1[u] = r0[u]; two[u] = false; off[u] = -1; gv[u] = 0.f;
        if (idx < n) {
            int oh, ow;
            g.d_ow.divmod(idx, oh, ow);
            off[u] = (2 * oh) * g.W + 2 * ow;
            r0[u] = __ldg(reinterpret_cast<const float2*>(px + off[u]));
            d0[u] = *reinterpret_cast<const float2*>(pdx + off[u]);
            two[u] = 2 * oh + 1 < g.H;
            if (two[u]) {
                r1[u] = __ldg(reinterpret_cast<const float2*>(px + off[u] + g.W));
                d1[u] = *reinterpret_cast<const float2*>(pdx + off[u] + g.W);
            }
            gv[u] = __ldg(pdy + idx);
        }
    }
#pragma unroll
    for (int u = 0; u < POOL2_PER_THREAD; ++u) {
        if (off[u] >= 0) {
            float m; int arg;
            pool2_argmax(r0[u], r1[u], two[u], m, arg);
            if (arg < 2) {
                if (arg == 0) d0[u].x += gv[u]; else d0[u].y += gv[u];
                *reinterpret_cast<float2*>(pdx + off[u]) = d0[u];
            } else {
                if (arg == 2) d1[u].x += gv[u]; else d1[u].y += gv[u];
                *reinterpret_cast<float2*>(pdx + off[u] + g.W) = d1[u];
            }
        }
    }
}

__global__ void __launch_bounds__(256) maxpool_bwd_acc_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                             float* __restrict__ dx, PoolGeom g) {
    const int plane = blockIdx.x;
    const int idx = blockIdx.y * blockDim.x + threadIdx.x;
    if (idx >= g.OH * g.OW) return;
    int oh, ow;
    g.d_ow.divmod(idx, oh, ow);
    const float* px = x + (size_t)plane * g.H * g.W;
    const int h0 = oh * g.k, w0 = ow * g.k;
    const int h1 = min(h0 + g.k, g.H), w1 = min(w0 + g.k, g.W);
    float m = -INFINITY;
    int ah = h0, aw = w0;
    for (int h = h0; h < h1; ++h)
        for (int w = w0; w < w1; ++w) {
            const float v = __ldg(px + (size_t)h * g.W + w);
            if (v > m || v != v) { m = v; ah = h; aw = w; }
        }
    dx[(size_t)plane * g.H * g.W + (size_t)ah * g.W + aw] += __ldg(dy + (size_t)plane * g.OH * g.OW + idx);
}

// the 2 x 2 fast path needs every window row to be one aligned 8-byte pair
static bool pool2_ok(const cpc_pool_params* p, const void* a, const void* b) {
    return p->kernel == 2 && p->w_in % 2 == 0 && p->w_out * 2 == p->w_in &&
           ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b)) & 7) == 0 &&
           ((int64_t)p->h_in * p->w_in) % 2 == 0;
}

static int pool_validate(const cpc_pool_params* p) {
    if (!p) return CPC_ERR_NULL;
    if (p->batch <= 0 || p->channels <= 0 || p->h_in <= 0 || p->w_in <= 0 || p->kernel <= 0) return CPC_ERR_BAD_SHAPE;
    const int oh = p->ceil_mode ? ceil_div(p->h_in, p->kernel) : p->h_in / p->kernel;
    const int ow = p->ceil_mode ? ceil_div(p->w_in, p->kernel) : p->w_in / p->kernel;
    if (oh <= 0 || ow <= 0 || oh != p->h_out || ow != p->w_out) return CPC_ERR_BAD_SHAPE;
    if ((int64_t)p->batch * p->channels > (1ll << 30) || (int64_t)p->h_out * p->w_out > 65535ll * 256) return CPC_ERR_BAD_SHAPE;
    return CPC_OK;
}

static PoolGeom pool_geom(const cpc_pool_params* p) {
    PoolGeom g;
    g.H = p->h_in; g.W = p->w_in; g.OH = p->h_out; g.OW = p->w_out; g.k = p->kernel;
    g.d_ow = FastDiv(g.OW);
    return g;
}

}  // namespace cpc

using namespace cpc;

extern "C" int cpc_maxpool_fwd(const float* x, float* y, const cpc_pool_params* p, void* stream) {
    int st = pool_validate(p);
    if (st != CPC_OK) return st;
    if (!x || !y) return CPC_ERR_NULL;
    if ((st = check_device()) != CPC_OK) return st;
    PoolGeom g = pool_geom(p);
    const int planes = p->batch * p->channels;
    if (pool2_ok(p, x, x))
        maxpool2_fwd_kernel<<<dim3(planes, ceil_div(g.OH * g.OW, 256 * POOL2_PER_THREAD)), 256, 0, (cudaStream_t)stream>>>(x, y, g);
    else
        maxpool_fwd_kernel<<<dim3(planes, ceil_div(g.OH * g.OW, 256)), 256, 0, (cudaStream_t)stream>>>(x, y, g);
    CPC_LAUNCH_CHECK();
    count_launch();
    return CPC_OK;
}

extern "C" int cpc_maxpool_bwd(const float* x, const float* dy, float* dx, const cpc_pool_params* p, void* stream) {
    int st = pool_validate(p);
    if (st != CPC_OK) return st;
    if (!x || !dy || !dx) return CPC_ERR_NULL;
    if ((st = check_device()) != CPC_OK) return st;
    PoolGeom g = pool_geom(p);
    const int planes = p->batch * p->channels;
    cudaStream_t s = (cudaStream_t)stream;
    // floor mode can leave trailing rows / columns outside every window: they get zero gradient
    if (g.OH * g.k < g.H || g.OW * g.k < g.W)
        if (cudaMemsetAsync(dx, 0, sizeof(float) * (size_t)planes * g.H * g.W, s) != cudaSuccess) return CPC_ERR_CUDA;
    if (pool2_ok(p, x, dx))
        maxpool2_bwd_kernel<<<dim3(planes, ceil_div(g.OH * g.OW, 256 * POOL2_PER_THREAD)), 256, 0, s>>>(x, dy, dx, g);
    else
        maxpool_bwd_kernel<<<dim3(planes, ceil_div(g.OH * g.OW, 256)), 256, 0, s>>>(x, dy, dx, g);
    CPC_LAUNCH_CHECK();
    count_launch();
    return CPC_OK;
}

extern "C" int cpc_maxpool_bwd_accumulate(const float* x, const float* dy, float* dx, const cpc_pool_params* p, void* stream) {
    int st = pool_validate(p);
    if (st != CPC_OK) return st;
    if (!x || !dy || !dx) return CPC_ERR_NULL;
    if ((st = check_device()) != CPC_OK) return st;
    PoolGeom g = pool_geom(p);
    const int planes = p->batch * p->channels;
    cudaStream_t s = (cudaStream_t)stream;
    if (pool2_ok(p, x, dx))
        maxpool2_bwd_acc_kernel<<<dim3(planes, ceil_div(g.OH * g.OW, 256 * POOL2_PER_THREAD)), 256, 0, s>>>(x, dy, dx, g);
    else
        maxpool_bwd_acc_kernel<<<dim3(planes, ceil_div(g.OH * g.OW, 256)), 256, 0, s>>>(x, dy, dx, g);
    CPC_LAUNCH_CHECK();
    count_launch();
    return CPC_OK;
}
