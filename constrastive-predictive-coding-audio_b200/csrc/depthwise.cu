// Depthwise convolution (groups = channels, multiplier 1, no bias), forward / data gradient / weight gradient.
// Replaces the first half of Conv2dSeparable (scalogram_model.py:532-544: nn.Conv2d(C, C, k, groups=C, bias=False)
// followed by a 1x1 conv, which runs on the implicit-GEMM kernels) and its autograd.  One multiply-add per tap and
// element: a memory-bound kernel; the k-fold reuse of every input element is served by L1 / L2 (a thread block owns
// a 32 x 8 output patch of one plane, so the rows it re-reads per vertical tap stay resident).
#include "common.cuh"

namespace cpc {

struct DwGeom {
    int C, H, W, OH, OW, kh, kw, sh, sw, pt, pl;
};

// grid (ceil(OW / 32), ceil(OH / 8), B * C); block (32, 8)
__global__ void __launch_bounds__(256) dwconv_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                        float* __restrict__ y, DwGeom g) {
    const int plane = blockIdx.z, c = plane % g.C;
    const int ow = blockIdx.x * 32 + threadIdx.x, oh = blockIdx.y * 8 + threadIdx.y;
    if (ow >= g.OW || oh >= g.OH) return;
    const float* px = x + (size_t)plane * g.H * g.W;
    const float* pw = w + (size_t)c * g.kh * g.kw;
    float acc = 0.f;
    for (int i = 0; i < g.kh; ++i) {
        const int h = oh * g.sh + i - g.pt;
        if (h < 0 || h >= g.H) continue;
        for (int j = 0; j < g.kw; ++j) {
            const int ww = ow * g.sw + j - g.pl;
            if (ww < 0 || ww >= g.W) continue;
            acc = fmaf(__ldg(px + (size_t)h * g.W + ww), __ldg(pw + i * g.kw + j), acc);
        }
    }
    y[((size_t)plane * g.OH + oh) * g.OW + ow] = acc;
}

// dx[h, w] = sum over taps (i, j) with (h + pt - i) % sh == 0, (w + pl - j) % sw == 0 of dy[(h+pt-i)/sh, (w+pl-j)/sw] * w[i, j]
__global__ void __launch_bounds__(256) dwconv_dgrad_kernel(const float* __restrict__ dy, const float* __restrict__ w,
                                                          float* __restrict__ dx, DwGeom g) {
    const int plane = blockIdx.z, c = plane % g.C;
    const int ww = blockIdx.x * 32 + threadIdx.x, h = blockIdx.y * 8 + threadIdx.y;
    if (ww >= g.W || h >= g.H) return;
    const float* pdy = dy + (size_t)plane * g.OH * g.OW;
    const float* pw = w + (size_t)c * g.kh * g.kw;
    float acc = 0.f;
    for (int i = 0; i < g.kh; ++i) {
        const int th = h + g.pt - i;
        if (th < 0 || th % g.sh) continue;
        const int oh = th / g.sh;
        if (oh >= g.OH) continue;
        for (int j = 0; j < g.kw; ++j) {
            const int tw = ww + g.pl - j;
            if (tw < 0 || tw % g.sw) continue;
            const int ow = tw / g.sw;
            if (ow >= g.OW) continue;
            acc = fmaf(__ldg(pdy + (size_t)oh * g.OW + ow), __ldg(pw + i * g.kw + j), acc);
        }
    }
    dx[((size_t)plane * g.H + h) * g.W + ww] = acc;
}

// dw[c, i, j] = sum_{b, oh, ow} dy[b, c, oh, ow] * x[b, c, oh*sh + i - pt, ow*sw + j - pl]
// grid (C, B): a block reduces one plane pair for every tap (the pair is re-read per tap out of L1 / L2) and adds its
// partial sums to dw (zeroed by the caller) with one atomic per tap.
__global__ void __launch_bounds__(256) dwconv_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                          float* __restrict__ dw, DwGeom g) {
    __shared__ float red[8];
    const int c = blockIdx.x, b = blockIdx.y;
    const size_t plane = (size_t)b * g.C + c;
    const float* px = x + plane * g.H * g.W;
    const float* pdy = dy + plane * g.OH * g.OW;
    const int n = g.OH * g.OW;
    for (int i = 0; i < g.kh; ++i)
        for (int j = 0; j < g.kw; ++j) {
            float acc = 0.f;
            for (int idx = threadIdx.x; idx < n; idx += blockDim.x) {
                const int oh = idx / g.OW, ow = idx - oh * g.OW;
                const int h = oh * g.sh + i - g.pt, ww = ow * g.sw + j - g.pl;
                if (h >= 0 && h < g.H && ww >= 0 && ww < g.W)
                    acc = fmaf(__ldg(pdy + idx), __ldg(px + (size_t)h * g.W + ww), acc);
            }
            acc = warp_sum(acc);
            if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
            __syncthreads();
            if (threadIdx.x < 32) {
                float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
                v = warp_sum(v);
                if (threadIdx.x == 0) atomicAdd(dw + ((size_t)c * g.kh + i) * g.kw + j, v);
            }
            __syncthreads();
        }
}

static int dw_validate(const cpc_conv_params* p) {
    if (!p) return CPC_ERR_NULL;
    if (p->batch <= 0 || p->c_in <= 0 || p->c_out != p->c_in || p->h_in <= 0 || p->w_in <= 0 || p->h_out <= 0 ||
        p->w_out <= 0 || p->kh <= 0 || p->kw <= 0 || p->stride_h <= 0 || p->stride_w <= 0 || p->pad_top < 0 ||
        p->pad_left < 0)
        return CPC_ERR_BAD_SHAPE;
    // every output reads at least one row / column that starts inside the padded input
    if ((int64_t)(p->h_out - 1) * p->stride_h - p->pad_top >= p->h_in ||
        (int64_t)(p->w_out - 1) * p->stride_w - p->pad_left >= p->w_in)
        return CPC_ERR_BAD_SHAPE;
    if ((int64_t)p->batch * p->c_in > 65535 * 1024ll) return CPC_ERR_BAD_SHAPE;
    return CPC_OK;
}

static DwGeom dw_geom(const cpc_conv_params* p) {
    return DwGeom{p->c_in, p->h_in, p->w_in, p->h_out, p->w_out, p->kh, p->kw, p->stride_h, p->stride_w, p->pad_top, p->pad_left};
}

}  // namespace cpc

using namespace cpc;

extern "C" int cpc_dwconv_fwd(const float* x, const float* w, float* y, const cpc_conv_params* p, void* stream) {
    int st = dw_validate(p);
    if (st != CPC_OK) return st;
    if (!x || !w || !y) return CPC_ERR_NULL;
    if ((st = check_device()) != CPC_OK) return st;
    const DwGeom g = dw_geom(p);
    dim3 grid(ceil_div(g.OW, 32), ceil_div(g.OH, 8), p->batch * g.C);
    dwconv_fwd_kernel<<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>(x, w, y, g);
    CPC_LAUNCH_CHECK();
    count_launch();
    return CPC_OK;
}

extern "C" int cpc_dwconv_dgrad(const float* dy, const float* w, float* dx, const cpc_conv_params* p, void* stream) {
    int st = dw_validate(p);
    if (st != CPC_OK) return st;
    if (!dy || !w || !dx) return CPC_ERR_NULL;
    if ((st = check_device()) != CPC_OK) return st;
    const DwGeom g = dw_geom(p);
    dim3 grid(ceil_div(g.W, 32), ceil_div(g.H, 8), p->batch * g.C);
    dwconv_dgrad_kernel<<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>(dy, w, dx, g);
    CPC_LAUNCH_CHECK();
    count_launch();
    return CPC_OK;
}

extern "C" int cpc_dwconv_wgrad(const float* x, const float* dy, float* dw, const cpc_conv_params* p, void* stream) {
    int st = dw_validate(p);
    if (st != CPC_OK) return st;
    if (!x || !dy || !dw) return CPC_ERR_NULL;
    if ((st = check_device()) != CPC_OK) return st;
    const DwGeom g = dw_geom(p);
    cudaStream_t s = (cudaStream_t)stream;
    if (cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)g.C * g.kh * g.kw, s) != cudaSuccess) return CPC_ERR_CUDA;
    dwconv_wgrad_kernel<<<dim3(g.C, p->batch), 256, 0, s>>>(x, dy, dw, g);
    CPC_LAUNCH_CHECK();
    count_launch();
    return CPC_OK;
}
