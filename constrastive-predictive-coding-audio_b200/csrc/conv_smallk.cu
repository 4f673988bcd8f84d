// Direct convolution kernels for layers whose reduction length K = C_in * kh * kw is tiny (<= 36):
// the first scalogram conv (3x3 stride 2 on the 2-channel scalogram, scalogram_model.py:387-397), the 1x1
// residual conv on the 2-channel pooled scalogram (:442-446) and AudioEncoder layer 0 (C_in = 1, k = 10,
// audio_model.py:30-34).  As GEMMs these have K = 2 .. 18: they are bound by the activation read/write, so
// they run on the CUDA cores with one pass over the data instead of the tensor-core machinery.
//   forward : one thread per output pixel, 32 output channels in registers, weights broadcast from smem.
//   wgrad   : lanes = consecutive pixels (coalesced dy / x reads), each warp owns 4 output channels x K taps
//             in registers, warp-shuffle reduction, one atomicAdd per (co, tap) and block.
#include "common.cuh"

namespace cpc {

struct SmallGeom {
    int B, Cin, H, W, Cout, OH, OW, kh, kw, sh, sw, pt, pl, K, ohow;
    FastDiv d_ow, d_kw, d_khkw;
};

static SmallGeom small_geom(const cpc_conv_params* p) {
    SmallGeom g;
    g.B = p->batch; g.Cin = p->c_in; g.H = p->h_in; g.W = p->w_in; g.Cout = p->c_out; g.OH = p->h_out; g.OW = p->w_out;
    g.kh = p->kh; g.kw = p->kw; g.sh = p->stride_h; g.sw = p->stride_w; g.pt = p->pad_top; g.pl = p->pad_left;
    g.K = g.Cin * g.kh * g.kw; g.ohow = g.OH * g.OW;
    g.d_ow = FastDiv(g.OW); g.d_kw = FastDiv(g.kw); g.d_khkw = FastDiv(g.kh * g.kw);
    return g;
}

// Per-tap offsets relative to the window origin (ci * H * W + i * W + j), -1 for k >= K; computed once per thread.
template <int KT>
__device__ __forceinline__ void small_tap_offsets(const SmallGeom& g, int (&toff)[KT]) {
#pragma unroll
    for (int k = 0; k < KT; ++k) {
        int ci, r, i, j;
        g.d_khkw.divmod(k, ci, r);
        g.d_kw.divmod(r, i, j);
        toff[k] = k < g.K ? (ci * g.H + i) * g.W + j : -1;
    }
}

// gathers the K input taps of output pixel (b, pix) into v[0..KT) (zeros for padding and k >= K).
// `interior`: no window touches padding, so the taps are plain offset loads.
template <int KT>
__device__ __forceinline__ void small_gather(const float* __restrict__ x, const SmallGeom& g, const int (&toff)[KT],
                                             bool interior, int b, int pix, float (&v)[KT]) {
    int oh, ow;
    g.d_ow.divmod(pix, oh, ow);
    const int h0 = oh * g.sh - g.pt, w0 = ow * g.sw - g.pl;
    const float* xo = x + (size_t)b * g.Cin * g.H * g.W + (long)h0 * g.W + w0;
    if (interior) {
#pragma unroll
        for (int k = 0; k < KT; ++k) v[k] = toff[k] >= 0 ? __ldg(xo + toff[k]) : 0.f;
        return;
    }
#pragma unroll
    for (int k = 0; k < KT; ++k) {
        float t = 0.f;
        if (toff[k] >= 0) {
            int ci, r, i, j;
            g.d_khkw.divmod(k, ci, r);
            g.d_kw.divmod(r, i, j);
            if ((unsigned)(h0 + i) < (unsigned)g.H && (unsigned)(w0 + j) < (unsigned)g.W) t = __ldg(xo + toff[k]);
        }
        v[k] = t;
    }
}

// grid (pixel chunks, B, Cout / 32); block 256 threads, SF_PIX pixels per thread 256 apart (ncu: with one pixel per
// thread the kernel was instruction-issue bound at 79 % -- 144 shared-memory weight reads and 64-bit store addressing
// per 576 FMAs; two pixels share every weight read)
constexpr int SF_PIX = 2;
template <int KT>
__global__ void __launch_bounds__(256) small_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                       const float* __restrict__ bias, float* __restrict__ y, SmallGeom g,
                                                       int relu, int interior) {
    __shared__ __align__(16) float ws[KT][32];
    __shared__ float bs[32];
    const int co0 = blockIdx.z * 32;
    for (int idx = threadIdx.x; idx < KT * 32; idx += blockDim.x) {
        const int k = idx >> 5, c = idx & 31;
        ws[k][c] = (k < g.K && co0 + c < g.Cout) ? __ldg(w + (size_t)(co0 + c) * g.K + k) : 0.f;
    }
    if (threadIdx.x < 32) bs[threadIdx.x] = (bias && co0 + threadIdx.x < g.Cout) ? __ldg(bias + co0 + threadIdx.x) : 0.f;
    __syncthreads();
    const int pix0 = blockIdx.x * (256 * SF_PIX) + threadIdx.x;
    if (pix0 >= g.ohow) return;
    const int b = blockIdx.y;
    int toff[KT];
    small_tap_offsets<KT>(g, toff);
    float v[SF_PIX][KT];
    bool live[SF_PIX];
#pragma unroll
    for (int u = 0; u < SF_PIX; ++u) {
        live[u] = pix0 + u * 256 < g.ohow;
        if (live[u]) small_gather<KT>(x, g, toff, interior != 0, b, pix0 + u * 256, v[u]);
        else {
#pragma unroll
            for (int k = 0; k < KT; ++k) v[u][k] = 0.f;
        }
    }
    float acc[SF_PIX][32];
#pragma unroll
    for (int u = 0; u < SF_PIX; ++u)
#pragma unroll
        for (int c = 0; c < 32; ++c) acc[u][c] = bs[c];
#pragma unroll
    for (int k = 0; k < KT; ++k) {
#pragma unroll
        for (int c4 = 0; c4 < 8; ++c4) {
            const float4 wv = *reinterpret_cast<const float4*>(&ws[k][c4 * 4]);
#pragma unroll
            for (int u = 0; u < SF_PIX; ++u) {
                acc[u][c4 * 4 + 0] = fmaf(v[u][k], wv.x, acc[u][c4 * 4 + 0]);
                acc[u][c4 * 4 + 1] = fmaf(v[u][k], wv.y, acc[u][c4 * 4 + 1]);
                acc[u][c4 * 4 + 2] = fmaf(v[u][k], wv.z, acc[u][c4 * 4 + 2]);
                acc[u][c4 * 4 + 3] = fmaf(v[u][k], wv.w, acc[u][c4 * 4 + 3]);
            }
        }
    }
    const int n_ch = min(32, g.Cout - co0);
    float* yo = y + ((size_t)b * g.Cout + co0) * g.ohow + pix0;
    if (n_ch == 32) {
#pragma unroll
        for (int c = 0; c < 32; ++c, yo += g.ohow) {
#pragma unroll
            for (int u = 0; u < SF_PIX; ++u)
                if (live[u]) yo[u * 256] = relu ? fmaxf(acc[u][c], 0.f) : acc[u][c];
        }
    } else {
#pragma unroll
        for (int c = 0; c < 32; ++c, yo += g.ohow) {
            if (c < n_ch) {
#pragma unroll
                for (int u = 0; u < SF_PIX; ++u)
                    if (live[u]) yo[u * 256] = relu ? fmaxf(acc[u][c], 0.f) : acc[u][c];
            }
        }
    }
}

// grid (pixel chunks of SW_ITERS * 32 pixels, B, Cout / 32); block = 32 / CPW warps, warp w owns channels CPW*w .. +CPW-1
// (CPW = 8 while the CPW x KT accumulators fit in registers: every x tap loaded is then reused by 8 channels)
constexpr int SW_ITERS = 64;
template <int KT, int CPW>
__global__ void __launch_bounds__(32 * (32 / CPW)) small_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                                     float* __restrict__ dw, SmallGeom g, int interior) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int toff[KT];
    small_tap_offsets<KT>(g, toff);
    const int b = blockIdx.y;
    const int co = blockIdx.z * 32 + warp * CPW;
    float acc[CPW][KT];
#pragma unroll
    for (int c = 0; c < CPW; ++c)
#pragma unroll
        for (int k = 0; k < KT; ++k) acc[c][k] = 0.f;
    const int p_begin = blockIdx.x * (SW_ITERS * 32);
    const float* dyb = dy + ((size_t)b * g.Cout + co) * g.ohow;
    for (int it = 0; it < SW_ITERS; ++it) {
        const int pix = p_begin + it * 32 + lane;
        if (pix >= g.ohow) break;                                   // later iterations are out of range for every lane >= this one
        float v[KT];
        small_gather<KT>(x, g, toff, interior != 0, b, pix, v);
        float d[CPW];
#pragma unroll
        for (int c = 0; c < CPW; ++c) d[c] = co + c < g.Cout ? __ldg(dyb + (size_t)c * g.ohow + pix) : 0.f;
#pragma unroll
        for (int c = 0; c < CPW; ++c)
#pragma unroll
            for (int k = 0; k < KT; ++k) acc[c][k] = fmaf(d[c], v[k], acc[c][k]);
    }
#pragma unroll
    for (int c = 0; c < CPW; ++c)
#pragma unroll
        for (int k = 0; k < KT; ++k) {
            const float s = warp_sum(acc[c][k]);
            if (lane == 0 && k < g.K && co + c < g.Cout) atomicAdd(dw + (size_t)(co + c) * g.K + k, s);
        }
}

bool smallk_eligible(const cpc_conv_params* p, int which) {
    if (which == 1) return false;                                   // data gradient stays on the generic kernels
    const int64_t k = (int64_t)p->c_in * p->kh * p->kw;
    // C_in >= 16 has a tensor-core path; this family is for the activation-bound first layers
    return k <= 36 && p->c_in < 16 && p->batch <= 65535 && (p->c_out + 31) / 32 <= 65535;
}

template <int KT>
static void small_launch(int which, const float* x, const float* w, const float* bias, const float* dy, float* out,
                         const SmallGeom& g, int relu, cudaStream_t s) {
    // every window inside the input <=> no bounds checks in the gather
    const int interior = g.pt == 0 && g.pl == 0 && (g.OH - 1) * g.sh + g.kh <= g.H && (g.OW - 1) * g.sw + g.kw <= g.W;
    if (which == 0) {
        dim3 grid(ceil_div(g.ohow, 256 * SF_PIX), g.B, ceil_div(g.Cout, 32));
        small_fwd_kernel<KT><<<grid, 256, 0, s>>>(x, w, bias, out, g, relu, interior);
    } else {
        dim3 grid(ceil_div(g.ohow, SW_ITERS * 32), g.B, ceil_div(g.Cout, 32));
        constexpr int CPW = KT <= 18 ? 8 : 4;
        small_wgrad_kernel<KT, CPW><<<grid, 32 * (32 / CPW), 0, s>>>(x, dy, out, g, interior);
    }
}

// which = 0: y = conv(x, w) + bias [relu];  which = 2: dw = wgrad(x, dy) (dw zero-initialised here)
int smallk_launch(int which, const float* x, const float* w, const float* bias, const float* dy, float* out,
                  const cpc_conv_params* p, cudaStream_t s) {
    if (!smallk_eligible(p, which)) return CPC_ERR_UNSUPPORTED;
    SmallGeom g = small_geom(p);
    if (which == 2 && cudaMemsetAsync(out, 0, sizeof(float) * (size_t)g.Cout * g.K, s) != cudaSuccess) return CPC_ERR_CUDA;
    if (g.K <= 2) small_launch<2>(which, x, w, bias, dy, out, g, p->relu, s);
    else if (g.K <= 4) small_launch<4>(which, x, w, bias, dy, out, g, p->relu, s);
    else if (g.K <= 10) small_launch<10>(which, x, w, bias, dy, out, g, p->relu, s);
    else if (g.K <= 18) small_launch<18>(which, x, w, bias, dy, out, g, p->relu, s);
    else if (g.K <= 27) small_launch<27>(which, x, w, bias, dy, out, g, p->relu, s);
    else small_launch<36>(which, x, w, bias, dy, out, g, p->relu, s);
    if (cudaGetLastError() != cudaSuccess) return CPC_ERR_CUDA;
    count_launch();
    return CPC_OK;
}

}  // namespace cpc
