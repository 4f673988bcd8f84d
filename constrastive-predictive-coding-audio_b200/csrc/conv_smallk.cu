// Direct convolution kernels for layers whose reduction length K = C_in * kh * kw is tiny (<= 36):
// the first scalogram conv (3x3 stride 2 on the 2-channel scalogram, scalogram_model.py:387-397), the 1x1
// residual conv on the 2-channel pooled scalogram (:442-446) and AudioEncoder layer 0 (C_in = 1, k = 10,
// audio_model.py:30-34).  As GEMMs these have K = 2 .. 18: they are bound by the activation read/write, so
// they run on the CUDA cores with one pass over the data instead of the tensor-core machinery.
//   forward : one thread per output pixel, 32 output channels in registers, weights broadcast from smem.
//   wgrad   : dw[co][tap] = sum_pixels dy[co][pix] * x[tap-shifted pix] as warp-level tensor-core MMAs
//             (mma.sync m16n8k16: M = 32 output channels, N = taps, reduction over 16 consecutive output pixels of a row),
//             operands split into bf16 hi / lo in registers (three MMAs per product, fp32 accumulate), 24 accumulator
//             registers per thread instead of the 144 of the FMA formulation it replaces (12 % occupancy, 0.48 ms for the
//             first e24 layer; the layer's data is 409 MB = 0.07 ms at the HBM roofline).
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

// ---- weight gradient on warp-level tensor cores ---------------------------------------------------------------------
__device__ __forceinline__ uint32_t bf16x2_rn(float lo_elem, float hi_elem) {       // {hi_elem, lo_elem} -> one b32
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi_elem), "f"(lo_elem));
    return r;
}
// v0 (low half), v1 (high half) -> packed bf16 hi parts and packed bf16 residuals
__device__ __forceinline__ void split_pair(float v0, float v1, uint32_t& hi, uint32_t& lo) {
    hi = bf16x2_rn(v0, v1);
    const float h0 = __uint_as_float(hi << 16), h1 = __uint_as_float(hi & 0xffff0000u);
    lo = bf16x2_rn(v0 - h0, v1 - h1);
}
__device__ __forceinline__ void mma_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// grid (blocks, Cout / 32); 256 threads.  A warp walks over units (item, output row, 16-pixel segment); NT = ceil(K / 8).
template <int NT>
__global__ void __launch_bounds__(256) small_wgrad_mma_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                             float* __restrict__ dw, SmallGeom g, int n_units, int segs) {
    __shared__ int tap_off[NT * 8], tap_i[NT * 8], tap_j[NT * 8];
    __shared__ float red[32][NT * 8 + 1];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int gq = lane >> 2, tq = lane & 3;                       // fragment row group / thread-in-group
    for (int n = threadIdx.x; n < NT * 8; n += blockDim.x) {
        int ci, r, i, j;
        g.d_khkw.divmod(n, ci, r);
        g.d_kw.divmod(r, i, j);
        tap_off[n] = n < g.K ? (ci * g.H + i) * g.W + j : -1;
        tap_i[n] = i;
        tap_j[n] = j;
    }
    for (int idx = threadIdx.x; idx < 32 * (NT * 8 + 1); idx += blockDim.x) (&red[0][0])[idx] = 0.f;
    __syncthreads();
    const int co0 = blockIdx.y * 32;
    float acc[2][NT][4];
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int n = 0; n < NT; ++n)
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[m][n][q] = 0.f;
    const FastDiv d_segs(segs), d_oh(g.OH);
    const int warps_total = gridDim.x * (blockDim.x >> 5);
    for (int unit = blockIdx.x * (blockDim.x >> 5) + warp; unit < n_units; unit += warps_total) {
        int row, seg, b, oh;
        d_segs.divmod(unit, row, seg);
        d_oh.divmod(row, b, oh);
        const int ow0 = seg * 16;
        const int c0 = ow0 + 2 * tq, c1 = c0 + 8;                  // this thread's pixel pairs (c0, c0+1), (c1, c1+1)
        // ---- A: dy[co][pixel], rows co0 + {gq, gq+8, gq+16, gq+24}
        uint32_t a_hi[2][4], a_lo[2][4];
#pragma unroll
        for (int m = 0; m < 2; ++m)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int co = co0 + m * 16 + h * 8 + gq;
                const float* pr = dy + (((size_t)b * g.Cout + co) * g.OH + oh) * g.OW;
                const bool live = co < g.Cout;
                const float v00 = (live && c0 < g.OW) ? __ldg(pr + c0) : 0.f, v01 = (live && c0 + 1 < g.OW) ? __ldg(pr + c0 + 1) : 0.f;
                const float v10 = (live && c1 < g.OW) ? __ldg(pr + c1) : 0.f, v11 = (live && c1 + 1 < g.OW) ? __ldg(pr + c1 + 1) : 0.f;
                split_pair(v00, v01, a_hi[m][h], a_lo[m][h]);              // a0 / a1: columns 2t, 2t+1
                split_pair(v10, v11, a_hi[m][2 + h], a_lo[m][2 + h]);      // a2 / a3: columns 2t+8, 2t+9
            }
        // ---- B: x[tap][pixel] per 8-tap tile, then 2 x 3 MMAs
        const float* xb = x + (size_t)b * g.Cin * g.H * g.W;
#pragma unroll
        for (int n = 0; n < NT; ++n) {
            const int tap = n * 8 + gq;
            const int off = tap_off[tap];
            const int ih = oh * g.sh + tap_i[tap] - g.pt;
            const bool row_ok = off >= 0 && (unsigned)ih < (unsigned)g.H;
            const float* px = xb + (long)(oh * g.sh - g.pt) * g.W + off - g.pl;      // + ow * sw per pixel
            float v[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int ow = (q < 2 ? c0 : c1) + (q & 1);
                const int iw = ow * g.sw + tap_j[tap] - g.pl;
                v[q] = (row_ok && ow < g.OW && (unsigned)iw < (unsigned)g.W) ? __ldg(px + ow * g.sw) : 0.f;
            }
            uint32_t b_hi0, b_lo0, b_hi1, b_lo1;
            split_pair(v[0], v[1], b_hi0, b_lo0);
            split_pair(v[2], v[3], b_hi1, b_lo1);
#pragma unroll
            for (int m = 0; m < 2; ++m) {
                mma_16816(acc[m][n], a_hi[m], b_hi0, b_hi1);
                mma_16816(acc[m][n], a_hi[m], b_lo0, b_lo1);
                mma_16816(acc[m][n], a_lo[m], b_hi0, b_hi1);
            }
        }
    }
    // block reduction in shared memory, then one atomic per (co, tap) and block.  C fragment: c0/c1 row gq cols 2t, 2t+1;
    // c2/c3 row gq + 8
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int n = 0; n < NT; ++n)
#pragma unroll
            for (int q = 0; q < 4; ++q)
                atomicAdd(&red[m * 16 + (q >> 1) * 8 + gq][n * 8 + 2 * tq + (q & 1)], acc[m][n][q]);
    __syncthreads();
    for (int idx = threadIdx.x; idx < 32 * NT * 8; idx += blockDim.x) {
        const int c = idx / (NT * 8), k = idx - c * (NT * 8);
        if (k < g.K && co0 + c < g.Cout) atomicAdd(dw + (size_t)(co0 + c) * g.K + k, red[c][k]);
    }
}

template <int NT>
static void small_wgrad_mma_launch(const float* x, const float* dy, float* dw, const SmallGeom& g, cudaStream_t s) {
    const int segs = ceil_div(g.OW, 16);
    const long units = (long)g.B * g.OH * segs;
    int blocks = (int)((units + 7) / 8);
    if (blocks > 148 * 4) blocks = 148 * 4;
    small_wgrad_mma_kernel<NT><<<dim3(blocks, ceil_div(g.Cout, 32)), 256, 0, s>>>(x, dy, dw, g, (int)units, segs);
}

// ---- data gradient towards a tiny number of input channels ---------------------------------------------------------
// dx[b, ci, h, w] = sum_{co, i, j} dy[b, co, (h + pt - i) / sh, (w + pl - j) / sw] * w[co, ci, i, j]   (taps that divide).
// As a GEMM its N is C_in (1 or 2): the tiled kernels run it at 1/64 of their rate (28 ms for the first e29 layer, whose
// data gradient the gradient penalty needs).  Here: one thread per input pixel and all C_in channels, the whole weight
// tensor in shared memory (broadcast reads), dy read through L1 (a block owns a 32 x 8 patch, so the rows it re-reads per
// vertical tap stay resident).
constexpr int SD_MAX_CIN = 4;
constexpr int SD_MAX_W = 12288;                                   // floats of shared memory for the weights (48 KB)
template <int CI>
__global__ void __launch_bounds__(256) small_dgrad_kernel(const float* __restrict__ dy, const float* __restrict__ w,
                                                         float* __restrict__ dx, SmallGeom g) {
    extern __shared__ float ws[];                                 // [co][ci][i][j] as in global memory
    const int n_w = g.Cout * g.Cin * g.kh * g.kw;
    for (int idx = threadIdx.y * 32 + threadIdx.x; idx < n_w; idx += 256) ws[idx] = __ldg(w + idx);
    __syncthreads();
    const int b = blockIdx.z;
    const int ww = blockIdx.x * 32 + threadIdx.x, h = blockIdx.y * 8 + threadIdx.y;
    if (ww >= g.W || h >= g.H) return;
    float acc[CI];
#pragma unroll
    for (int c = 0; c < CI; ++c) acc[c] = 0.f;
    const float* dyb = dy + (size_t)b * g.Cout * g.ohow;
    const int khkw = g.kh * g.kw;
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
            const float* pd = dyb + (size_t)oh * g.OW + ow;
            const float* pw = ws + i * g.kw + j;
#pragma unroll 4
            for (int co = 0; co < g.Cout; ++co) {
                const float d = __ldg(pd + (size_t)co * g.ohow);
#pragma unroll
                for (int c = 0; c < CI; ++c)
                    if (c < g.Cin) acc[c] = fmaf(d, pw[(co * g.Cin + c) * khkw], acc[c]);
            }
        }
    }
#pragma unroll
    for (int c = 0; c < CI; ++c)
        if (c < g.Cin) dx[(((size_t)b * g.Cin + c) * g.H + h) * g.W + ww] = acc[c];
}

bool small_dgrad_eligible(const cpc_conv_params* p) {
    return p->c_in <= SD_MAX_CIN && (int64_t)p->c_out * p->c_in * p->kh * p->kw <= SD_MAX_W && p->batch <= 65535;
}

int small_dgrad_launch(const float* dy, const float* w, float* dx, const cpc_conv_params* p, cudaStream_t s) {
    if (!small_dgrad_eligible(p)) return CPC_ERR_UNSUPPORTED;
    const SmallGeom g = small_geom(p);
    const dim3 grid(ceil_div(g.W, 32), ceil_div(g.H, 8), g.B), block(32, 8);
    const size_t smem = sizeof(float) * (size_t)g.Cout * g.Cin * g.kh * g.kw;
    if (g.Cin == 1) small_dgrad_kernel<1><<<grid, block, smem, s>>>(dy, w, dx, g);
    else if (g.Cin == 2) small_dgrad_kernel<2><<<grid, block, smem, s>>>(dy, w, dx, g);
    else small_dgrad_kernel<SD_MAX_CIN><<<grid, block, smem, s>>>(dy, w, dx, g);
    if (cudaGetLastError() != cudaSuccess) return CPC_ERR_CUDA;
    count_launch();
    return CPC_OK;
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
    if (which == 2 && !(p->flags & CPC_CONV_FLAG_NO_MMA_SMALL_WGRAD) && (long)g.B * g.OH * ceil_div(g.OW, 16) < (1l << 31)) {
        switch (ceil_div(g.K, 8)) {
            case 1: small_wgrad_mma_launch<1>(x, dy, out, g, s); break;
            case 2: small_wgrad_mma_launch<2>(x, dy, out, g, s); break;
            case 3: small_wgrad_mma_launch<3>(x, dy, out, g, s); break;
            case 4: small_wgrad_mma_launch<4>(x, dy, out, g, s); break;
            default: small_wgrad_mma_launch<5>(x, dy, out, g, s); break;
        }
        if (cudaGetLastError() != cudaSuccess) return CPC_ERR_CUDA;
        count_launch();
        return CPC_OK;
    }
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
