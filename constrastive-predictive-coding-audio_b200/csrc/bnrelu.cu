// Fused BatchNorm2d (batch statistics) + ReLU (+ cropped residual add + ReLU), forward and backward.
// Replaces, for one conv stage of ScalogramEncoderBlock (scalogram_model.py:399-431), the cuDNN
// batch-norm forward/backward, the ReLU, and -- for the second stage of a block -- the centre-cropped
// residual add (scalogram_model.py:451-472) and the ReLU the encoder applies between blocks (:523-527):
//
//     v   = relu_if(relu, gamma * (x - mean) * rstd + beta)
//     out = relu_if(outer_relu, v + residual[:, :, off_h : off_h + H, off_w : off_w + W])
//
// All kernels are HBM-bound streaming passes over NCHW planes: forward = statistics pass (1 read) + apply
// pass (1-2 reads, 1 write); backward = reduction pass (2-3 reads) + gradient pass (2-3 reads, 1-2 writes).
// The inner ReLU mask is recomputed from x; the outer one (ReLU after the residual add) is either recomputed from x and
// the residual, or -- *_mask entry points -- read from a bit mask the forward pass wrote (1 bit per element instead of
// re-reading the residual in both backward passes: 2 x 327 MB for the first block of arch 7).
#include <cuda_bf16.h>
#include "common.cuh"

namespace cpc {

constexpr int BN_THREADS = 256;
constexpr int BN_PER_THREAD = 16;
constexpr int BN_SEG = BN_THREADS * BN_PER_THREAD;
// The backward kernels walk a thread's BN_PER_THREAD elements in batches of BN_BATCH (loads of a batch issued together).
// ncu: with all 16 elements of 2-3 tensors live at once they needed 74-86 registers, 2-3 blocks per SM, 20-35 % of the
// warp slots, and streamed at 3.2-4.4 TB/s, while the 32-register statistics kernel (95 % of the slots) reaches 6.1 TB/s:
// on B200 resident warps, not loads per thread, fill the memory pipeline.
constexpr int BN_BATCH = 4;

struct BnGeom {
    int B, C, H, W, HW;
    int RH, RW, roh, row;          // residual plane geometry; RH == 0 -> none
    int relu, outer_relu;
    FastDiv d_w;
};

static BnGeom bn_geom(const cpc_bn_params* p) {
    BnGeom g;
    g.B = p->batch; g.C = p->channels; g.H = p->height; g.W = p->width; g.HW = g.H * g.W;
    g.RH = p->res_height; g.RW = p->res_width; g.roh = p->res_off_h; g.row = p->res_off_w;
    g.relu = p->relu; g.outer_relu = p->outer_relu;
    g.d_w = FastDiv(g.W);
    return g;
}

// Threads per block of the plane-per-block streaming kernels (a block owns threads * BN_PER_THREAD consecutive
// elements): the plane is cut into as few blocks as 256 threads allow and the block is then shrunk to fit its share, so
// that no block runs mostly empty -- a 34 x 156 plane is 2 x 192 threads (86 % of the slots loaded) instead of one
// full and one 30 % full 256-thread block.
static int bn_block_threads(int HW) {
    const int nseg = ceil_div(HW, BN_SEG);
    const int per = ceil_div(HW, nseg);
    const int t = (ceil_div(per, BN_PER_THREAD) + 31) & ~31;
    return t < 32 ? 32 : (t > BN_THREADS ? BN_THREADS : t);
}

static int bn_validate(const cpc_bn_params* p) {
    if (!p) return CPC_ERR_NULL;
    if (p->batch <= 0 || p->channels <= 0 || p->height <= 0 || p->width <= 0) return CPC_ERR_BAD_SHAPE;
    if ((int64_t)p->height * p->width > (1ll << 30) || (int64_t)p->batch * p->channels > (1ll << 30)) return CPC_ERR_BAD_SHAPE;
    if (p->res_height < 0 || p->res_width < 0) return CPC_ERR_BAD_SHAPE;
    if (p->res_height > 0) {
        if (p->res_off_h < 0 || p->res_off_w < 0 || p->res_off_h + p->height > p->res_height ||
            p->res_off_w + p->width > p->res_width)
            return CPC_ERR_BAD_SHAPE;
        if ((int64_t)p->res_height * p->res_width > (1ll << 30)) return CPC_ERR_BAD_SHAPE;
    }
    if (!(p->eps > 0.f) || p->momentum < 0.f || p->momentum > 1.f) return CPC_ERR_BAD_SHAPE;
    if (p->packed_planes < 0 || p->packed_planes > 2) return CPC_ERR_BAD_SHAPE;
    return CPC_OK;
}

__device__ __forceinline__ double block_sum_d(double v, double* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double r = 0.0;
    if (threadIdx.x < 32) {
        r = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
    }
    return r;
}

// Reduction kernels: a block walks BN_RED_SEGS consecutive segments of its plane with register accumulators and pays
// the block reduction + 2 double atomics once (with one segment per block that tail was ~25 % of a block's life and the
// ragged last segment of a 9828-element plane cost as much as a full one: 3.2 TB/s).
constexpr int BN_RED_SEGS = 4;

// grid (B*C planes, segment groups of one plane).  sums[c*2 + {0,1}] += sum x, sum x^2 (double).
__global__ void __launch_bounds__(BN_THREADS) bn_stats_kernel(const float* __restrict__ x, double* __restrict__ sums,
                                                             int C, int HW) {
    __shared__ double red[BN_THREADS / 32];
    const int plane = blockIdx.x;
    const int c = plane % C;
    const float* px = x + (size_t)plane * HW;
    float s = 0.f, q = 0.f;
    for (int seg = 0; seg < BN_RED_SEGS; ++seg) {
        const int i0 = (blockIdx.y * BN_RED_SEGS + seg) * (int)blockDim.x * BN_PER_THREAD + threadIdx.x;
        if (i0 - (int)threadIdx.x >= HW) break;
        float v[BN_PER_THREAD];
#pragma unroll
        for (int u = 0; u < BN_PER_THREAD; ++u) {
            const int i = i0 + u * (int)blockDim.x;
            v[u] = i < HW ? __ldg(px + i) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < BN_PER_THREAD; ++u) { s += v[u]; q = fmaf(v[u], v[u], q); }
    }
    const double bs = block_sum_d((double)s, red);
    const double bq = block_sum_d((double)q, red);
    if (threadIdx.x == 0) { atomicAdd(sums + 2 * c, bs); atomicAdd(sums + 2 * c + 1, bq); }
}

// One thread per channel: mean / rstd, affine (scale, shift), running statistics.
__global__ void bn_finalize_kernel(const double* __restrict__ sums, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float* __restrict__ running_mean,
                                   float* __restrict__ running_var, float* __restrict__ save_mean,
                                   float* __restrict__ save_rstd, float* __restrict__ affine, int C, double count,
                                   float eps, float momentum, int training) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    float mean, rstd;
    if (training) {
        const double m = sums[2 * c] / count;
        double var = sums[2 * c + 1] / count - m * m;
        if (var < 0.0) var = 0.0;
        mean = (float)m;
        rstd = (float)(1.0 / sqrt(var + (double)eps));
        if (running_mean) running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
        if (running_var) {
            const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
            running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
        }
    } else {
        mean = running_mean[c];
        rstd = rsqrtf(running_var[c] + eps);
    }
    save_mean[c] = mean;
    save_rstd[c] = rstd;
    const float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
    affine[2 * c] = g * rstd;
    affine[2 * c + 1] = b - g * rstd * mean;
}

// All streaming kernels below are two-phase per thread: issue every load of the thread's BN_PER_THREAD elements
// first (independent, predicated), then compute and store -- one load in flight per thread is latency-bound.
template <bool MASK>
__global__ void __launch_bounds__(BN_THREADS) bn_apply_kernel(const float* __restrict__ x, const float* __restrict__ affine,
                                                             const float* __restrict__ res, float* __restrict__ out,
                                                             uint32_t* __restrict__ mask, BnGeom g) {
    const int plane = blockIdx.x;
    const int c = plane % g.C;
    const float sc = __ldg(affine + 2 * c), sh = __ldg(affine + 2 * c + 1);
    const float* px = x + (size_t)plane * g.HW;
    float* po = out + (size_t)plane * g.HW;
    const float* pr = res ? res + (size_t)plane * g.RH * g.RW + (size_t)g.roh * g.RW + g.row : nullptr;
    // the mask variant works in batches: the ballots pin every load of a batch in front of them, and a full batch of
    // 2 x 16 live values would halve the occupancy (the plain variant is scheduled that way by the compiler anyway)
    constexpr int BATCH = MASK ? BN_BATCH : BN_PER_THREAD;
    uint32_t* pm = MASK ? mask + (size_t)plane * ((g.HW + 31) >> 5) : nullptr;
#pragma unroll 1
    for (int half = 0; half < BN_PER_THREAD / BATCH; ++half) {
        const int i0 = (blockIdx.y * BN_PER_THREAD + half * BATCH) * (int)blockDim.x + threadIdx.x;
        float xv[BATCH], rv[BATCH];
#pragma unroll
        for (int u = 0; u < BATCH; ++u) {
            const int i = i0 + u * (int)blockDim.x;
            xv[u] = i < g.HW ? __ldg(px + i) : 0.f;
            rv[u] = 0.f;
            if (pr && i < g.HW) {
                int h, w;
                g.d_w.divmod(i, h, w);
                rv[u] = __ldg(pr + (size_t)h * g.RW + w);
            }
        }
#pragma unroll
        for (int u = 0; u < BATCH; ++u) {
            const int i = i0 + u * (int)blockDim.x;
            float v = fmaf(xv[u], sc, sh);
            if (g.relu) v = fmaxf(v, 0.f);
            if (pr) {
                v += rv[u];
                if constexpr (MASK) {
                    // bit (i & 31) of word i / 32 of this plane: a warp's lanes hold 32 consecutive, 32-aligned elements
                    const unsigned bits = __ballot_sync(0xffffffffu, i < g.HW && v > 0.f);
                    if ((threadIdx.x & 31) == 0 && i < g.HW) pm[i >> 5] = bits;
                }
                if (g.outer_relu) v = fmaxf(v, 0.f);
            }
            if (i < g.HW) po[i] = v;
        }
    }
}

// Gradient entering the normalisation: g2 = dout * [outer mask] * [inner mask]; also returns xhat.
// outer_bit: -1 = recompute the outer mask from x and the residual, 0 / 1 = the bit the forward pass saved.
__device__ __forceinline__ float bn_grad_in(float dout, float xv, float mean, float rstd, float gam, float bet, bool has_res,
                                            float resv, const BnGeom& g, float& xhat, float& g1, int outer_bit = -1) {
    xhat = (xv - mean) * rstd;
    const float y = fmaf(xhat, gam, bet);
    const float v = g.relu ? fmaxf(y, 0.f) : y;
    g1 = dout;
    if (has_res && g.outer_relu && (outer_bit >= 0 ? outer_bit == 0 : !(v + resv > 0.f))) g1 = 0.f;
    return (g.relu && !(y > 0.f)) ? 0.f : g1;
}

// sums2[c*2 + {0,1}] += sum g2, sum g2 * xhat
template <bool MASK>
__global__ void __launch_bounds__(BN_THREADS) bn_bwd_reduce_kernel(const float* __restrict__ dout, const float* __restrict__ x,
                                                                  const float* __restrict__ gamma,
                                                                  const float* __restrict__ beta,
                                                                  const float* __restrict__ save_mean,
                                                                  const float* __restrict__ save_rstd,
                                                                  const float* __restrict__ res,
                                                                  const uint32_t* __restrict__ mask,
                                                                  double* __restrict__ sums2, BnGeom g) {
    __shared__ double red[BN_THREADS / 32];
    const int plane = blockIdx.x;
    const int c = plane % g.C;
    const float mean = __ldg(save_mean + c), rstd = __ldg(save_rstd + c);
    const float gam = gamma ? __ldg(gamma + c) : 1.f, bet = beta ? __ldg(beta + c) : 0.f;
    const float* px = x + (size_t)plane * g.HW;
    const float* pd = dout + (size_t)plane * g.HW;
    const float* pr = res ? res + (size_t)plane * g.RH * g.RW + (size_t)g.roh * g.RW + g.row : nullptr;
    const bool need_res = !MASK && pr != nullptr && g.outer_relu;
    const uint32_t* pm = MASK ? mask + (size_t)plane * ((g.HW + 31) >> 5) : nullptr;
    float s = 0.f, q = 0.f;
    for (int seg = 0; seg < BN_RED_SEGS; ++seg) {
        const int i0 = (blockIdx.y * BN_RED_SEGS + seg) * (int)blockDim.x * BN_PER_THREAD + threadIdx.x;
        if (i0 - (int)threadIdx.x >= g.HW) break;
#pragma unroll 1
        for (int part = 0; part < BN_PER_THREAD / BN_BATCH; ++part) {
            const int ib = i0 + part * BN_BATCH * (int)blockDim.x;
            if (ib - (int)threadIdx.x >= g.HW) break;
            float xv[BN_BATCH], dv[BN_BATCH], rv[MASK ? 1 : BN_BATCH];
            uint32_t ob = 0;                                         // MASK: bit u = saved outer-ReLU bit of element u
#pragma unroll
            for (int u = 0; u < BN_BATCH; ++u) {
                const int i = ib + u * (int)blockDim.x;
                const bool in = i < g.HW;
                xv[u] = in ? __ldg(px + i) : 0.f;
                dv[u] = in ? __ldg(pd + i) : 0.f;
                if constexpr (MASK) {
                    if (in) ob |= ((__ldg(pm + (i >> 5)) >> (i & 31)) & 1u) << u;
                    continue;
                }
                rv[MASK ? 0 : u] = 0.f;
                if (need_res && in) {
                    int h, w;
                    g.d_w.divmod(i, h, w);
                    rv[MASK ? 0 : u] = __ldg(pr + (size_t)h * g.RW + w);
                }
            }
#pragma unroll
            for (int u = 0; u < BN_BATCH; ++u) {
                const int i = ib + u * (int)blockDim.x;
                if (i < g.HW) {
                    float xhat, g1;
                    const float g2 = bn_grad_in(dv[u], xv[u], mean, rstd, gam, bet, pr != nullptr,
                                                MASK ? 0.f : rv[MASK ? 0 : u], g, xhat, g1, MASK ? (int)((ob >> u) & 1u) : -1);
                    s += g2;
                    q = fmaf(g2, xhat, q);
                }
            }
        }
    }
    const double bs = block_sum_d((double)s, red);
    const double bq = block_sum_d((double)q, red);
    if (threadIdx.x == 0) { atomicAdd(sums2 + 2 * c, bs); atomicAdd(sums2 + 2 * c + 1, bq); }
}

// dx = gamma * rstd * (g2 - mean(g2) - xhat * mean(g2 * xhat))   [training]
// dx = gamma * rstd * g2                                           [running statistics]
// d_res (inside the crop) = g1;  dgamma / dbeta written by the first block of each channel of item 0.
template <bool MASK>
__global__ void __launch_bounds__(BN_THREADS) bn_bwd_apply_kernel(const float* __restrict__ dout, const float* __restrict__ x,
                                                                 const float* __restrict__ gamma,
                                                                 const float* __restrict__ beta,
                                                                 const float* __restrict__ save_mean,
                                                                 const float* __restrict__ save_rstd,
                                                                 const float* __restrict__ res,
                                                                 const uint32_t* __restrict__ mask,
                                                                 const double* __restrict__ sums2, float* __restrict__ dx,
                                                                 float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                                 float* __restrict__ d_res, BnGeom g, double count,
                                                                 int training) {
    const int plane = blockIdx.x;
    const int c = plane % g.C;
    const float mean = __ldg(save_mean + c), rstd = __ldg(save_rstd + c);
    const float gam = gamma ? __ldg(gamma + c) : 1.f, bet = beta ? __ldg(beta + c) : 0.f;
    const double sg = sums2[2 * c], sgx = sums2[2 * c + 1];
    if (plane < g.C && blockIdx.y == 0 && threadIdx.x == 0) {
        if (dbeta) dbeta[c] = (float)sg;
        if (dgamma) dgamma[c] = (float)sgx;
    }
    const float m1 = training ? (float)(sg / count) : 0.f;
    const float m2 = training ? (float)(sgx / count) : 0.f;
    const float k = gam * rstd;
    const float* px = x + (size_t)plane * g.HW;
    const float* pd = dout + (size_t)plane * g.HW;
    float* pdx = dx + (size_t)plane * g.HW;
    const size_t roff = (size_t)plane * g.RH * g.RW + (size_t)g.roh * g.RW + g.row;
    const float* pr = res ? res + roff : nullptr;
    float* pdr = d_res ? d_res + roff : nullptr;
    const int i0 = blockIdx.y * (int)blockDim.x * BN_PER_THREAD + threadIdx.x;
    const bool need_res = !MASK && pr != nullptr && g.outer_relu;
    const uint32_t* pm = MASK ? mask + (size_t)plane * ((g.HW + 31) >> 5) : nullptr;
#pragma unroll 1
    for (int part = 0; part < BN_PER_THREAD / BN_BATCH; ++part) {
        const int ib = i0 + part * BN_BATCH * (int)blockDim.x;
        if (ib - (int)threadIdx.x >= g.HW) break;
        float xv[BN_BATCH], dv[BN_BATCH], rv[MASK ? 1 : BN_BATCH];
        int ro[MASK ? 1 : BN_BATCH];                                 // offset inside the residual plane (MASK: recomputed)
        uint32_t ob = 0;                                             // MASK: bit u = saved outer-ReLU bit of element u
#pragma unroll
        for (int u = 0; u < BN_BATCH; ++u) {
            const int i = ib + u * (int)blockDim.x;
            const bool in = i < g.HW;
            xv[u] = in ? __ldg(px + i) : 0.f;
            dv[u] = in ? __ldg(pd + i) : 0.f;
            if constexpr (MASK) {
                if (in) ob |= ((__ldg(pm + (i >> 5)) >> (i & 31)) & 1u) << u;
                continue;
            }
            rv[MASK ? 0 : u] = 0.f;
            ro[MASK ? 0 : u] = 0;
            if ((pr || pdr) && in) {
                int h, w;
                g.d_w.divmod(i, h, w);
                ro[MASK ? 0 : u] = h * g.RW + w;
                if (need_res) rv[MASK ? 0 : u] = __ldg(pr + ro[MASK ? 0 : u]);
            }
        }
#pragma unroll
        for (int u = 0; u < BN_BATCH; ++u) {
            const int i = ib + u * (int)blockDim.x;
            if (i < g.HW) {
                float xhat, g1;
                const float g2 = bn_grad_in(dv[u], xv[u], mean, rstd, gam, bet, pr != nullptr, MASK ? 0.f : rv[MASK ? 0 : u], g,
                                            xhat, g1, MASK ? (int)((ob >> u) & 1u) : -1);
                pdx[i] = k * (g2 - m1 - xhat * m2);
                if (pdr) {
                    if constexpr (MASK) {
                        int h, w;
                        g.d_w.divmod(i, h, w);
                        pdr[h * g.RW + w] = g1;
                    } else {
                        pdr[ro[MASK ? 0 : u]] = g1;
                    }
                }
            }
        }
    }
}

// ---- small planes -------------------------------------------------------------------------------------------------
// With H*W well below one segment (the 2 x 77 and 16 x 77 planes of the deepest block) the plane-per-block kernels above
// launch B*C nearly empty blocks that each pay a block reduction and two atomics: 0.2 ms for a 10 MB tensor.  Here a
// block owns BN_SMALL_CHUNK consecutive elements of ONE CHANNEL across the batch (element e = item * HW + i).
constexpr int BN_SMALL_CHUNK = BN_THREADS * 32;

struct SmallLoc { bool in; size_t off; int i, plane; };
__device__ __forceinline__ SmallLoc small_locate(const BnGeom& g, const FastDiv& d_hw, int c, int e, int n) {
    SmallLoc l;
    l.in = e < n;
    int b = 0;
    l.i = 0;
    if (l.in) d_hw.divmod(e, b, l.i);
    l.plane = b * g.C + c;
    l.off = (size_t)l.plane * g.HW + l.i;
    return l;
}

__global__ void __launch_bounds__(BN_THREADS) bn_small_stats_kernel(const float* __restrict__ x, double* __restrict__ sums,
                                                                   BnGeom g, FastDiv d_hw) {
    __shared__ double red[BN_THREADS / 32];
    const int c = blockIdx.x, n = g.B * g.HW;
    float s = 0.f, q = 0.f;
    const int e0 = blockIdx.y * BN_SMALL_CHUNK;
#pragma unroll 8
    for (int e = e0 + threadIdx.x; e < min(n, e0 + BN_SMALL_CHUNK); e += BN_THREADS) {
        const SmallLoc l = small_locate(g, d_hw, c, e, n);
        const float v = __ldg(x + l.off);
        s += v; q = fmaf(v, v, q);
    }
    const double bs = block_sum_d((double)s, red);
    const double bq = block_sum_d((double)q, red);
    if (threadIdx.x == 0) { atomicAdd(sums + 2 * c, bs); atomicAdd(sums + 2 * c + 1, bq); }
}

__global__ void __launch_bounds__(BN_THREADS) bn_small_apply_kernel(const float* __restrict__ x, const float* __restrict__ affine,
                                                                   const float* __restrict__ res, float* __restrict__ out,
                                                                   BnGeom g, FastDiv d_hw) {
    const int c = blockIdx.x, n = g.B * g.HW;
    const float sc = __ldg(affine + 2 * c), sh = __ldg(affine + 2 * c + 1);
    const int e0 = blockIdx.y * BN_SMALL_CHUNK;
#pragma unroll 8
    for (int e = e0 + threadIdx.x; e < min(n, e0 + BN_SMALL_CHUNK); e += BN_THREADS) {
        const SmallLoc l = small_locate(g, d_hw, c, e, n);
        float v = fmaf(__ldg(x + l.off), sc, sh);
        if (g.relu) v = fmaxf(v, 0.f);
        if (res) {
            int h, w;
            g.d_w.divmod(l.i, h, w);
            v += __ldg(res + (size_t)l.plane * g.RH * g.RW + (size_t)(g.roh + h) * g.RW + g.row + w);
            if (g.outer_relu) v = fmaxf(v, 0.f);
        }
        out[l.off] = v;
    }
}

__global__ void __launch_bounds__(BN_THREADS) bn_small_bwd_reduce_kernel(
    const float* __restrict__ dout, const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
    const float* __restrict__ save_mean, const float* __restrict__ save_rstd, const float* __restrict__ res,
    double* __restrict__ sums2, BnGeom g, FastDiv d_hw) {
    __shared__ double red[BN_THREADS / 32];
    const int c = blockIdx.x, n = g.B * g.HW;
    const float mean = __ldg(save_mean + c), rstd = __ldg(save_rstd + c);
    const float gam = gamma ? __ldg(gamma + c) : 1.f, bet = beta ? __ldg(beta + c) : 0.f;
    const bool need_res = res != nullptr && g.outer_relu;
    float s = 0.f, q = 0.f;
    const int e0 = blockIdx.y * BN_SMALL_CHUNK;
#pragma unroll 8
    for (int e = e0 + threadIdx.x; e < min(n, e0 + BN_SMALL_CHUNK); e += BN_THREADS) {
        const SmallLoc l = small_locate(g, d_hw, c, e, n);
        float rv = 0.f;
        if (need_res) {
            int h, w;
            g.d_w.divmod(l.i, h, w);
            rv = __ldg(res + (size_t)l.plane * g.RH * g.RW + (size_t)(g.roh + h) * g.RW + g.row + w);
        }
        float xhat, g1;
        const float g2 = bn_grad_in(__ldg(dout + l.off), __ldg(x + l.off), mean, rstd, gam, bet, res != nullptr, rv, g, xhat, g1);
        s += g2;
        q = fmaf(g2, xhat, q);
    }
    const double bs = block_sum_d((double)s, red);
    const double bq = block_sum_d((double)q, red);
    if (threadIdx.x == 0) { atomicAdd(sums2 + 2 * c, bs); atomicAdd(sums2 + 2 * c + 1, bq); }
}

__global__ void __launch_bounds__(BN_THREADS) bn_small_bwd_apply_kernel(
    const float* __restrict__ dout, const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
    const float* __restrict__ save_mean, const float* __restrict__ save_rstd, const float* __restrict__ res,
    const double* __restrict__ sums2, float* __restrict__ dx, float* __restrict__ dgamma, float* __restrict__ dbeta,
    float* __restrict__ d_res, BnGeom g, FastDiv d_hw, double count, int training) {
    const int c = blockIdx.x, n = g.B * g.HW;
    const float mean = __ldg(save_mean + c), rstd = __ldg(save_rstd + c);
    const float gam = gamma ? __ldg(gamma + c) : 1.f, bet = beta ? __ldg(beta + c) : 0.f;
    const double sg = sums2[2 * c], sgx = sums2[2 * c + 1];
    if (blockIdx.y == 0 && threadIdx.x == 0) {
        if (dbeta) dbeta[c] = (float)sg;
        if (dgamma) dgamma[c] = (float)sgx;
    }
    const float m1 = training ? (float)(sg / count) : 0.f;
    const float m2 = training ? (float)(sgx / count) : 0.f;
    const float k = gam * rstd;
    const bool need_res = res != nullptr && g.outer_relu;
    const int e0 = blockIdx.y * BN_SMALL_CHUNK;
#pragma unroll 8
    for (int e = e0 + threadIdx.x; e < min(n, e0 + BN_SMALL_CHUNK); e += BN_THREADS) {
        const SmallLoc l = small_locate(g, d_hw, c, e, n);
        size_t ro = 0;
        float rv = 0.f;
        if (res || d_res) {
            int h, w;
            g.d_w.divmod(l.i, h, w);
            ro = (size_t)l.plane * g.RH * g.RW + (size_t)(g.roh + h) * g.RW + g.row + w;
            if (need_res) rv = __ldg(res + ro);
        }
        float xhat, g1;
        const float g2 = bn_grad_in(__ldg(dout + l.off), __ldg(x + l.off), mean, rstd, gam, bet, res != nullptr, rv, g, xhat, g1);
        dx[l.off] = k * (g2 - m1 - xhat * m2);
        if (d_res) d_res[ro] = g1;
    }
}

// planes below this many elements take the per-channel kernels
constexpr int BN_SMALL_HW = 2048;

// ---- packed-output variants ---------------------------------------------------------------------------------------
// bf16 hi / lo planes of a conv operand ([plane][B*C*H rows][Wp], Wp = W rounded up to 8, pad columns zero).  Used when
// the only consumer of an activation / gradient is a tensor-core conv: the producer writes the operand form directly
// and the fp32 tensor plus its packing pass disappear.  W must be even: a thread owns PAIRS of horizontally adjacent
// elements (8-byte loads, 4-byte bf16x2 stores; 2-byte stores ran the kernel 25 % slower than the fp32 version).
struct PackedOut {
    __nv_bfloat16* base;
    long plane_stride;       // elements between the hi and the lo plane; 0: hi plane only (bf16 operand mode)
    int Wp;
};
constexpr int BN_PAIRS = BN_PER_THREAD / 2;

__device__ __forceinline__ void packed_store2(const PackedOut& po, const BnGeom& g, int plane, int h, int w, float a, float b) {
    __nv_bfloat16* q = po.base + ((long)plane * g.H + h) * po.Wp + w;
    const __nv_bfloat162 hi = __floats2bfloat162_rn(a, b);
    const __nv_bfloat162 lo = __floats2bfloat162_rn(a - __low2float(hi), b - __high2float(hi));
    *reinterpret_cast<__nv_bfloat162*>(q) = hi;
    if (po.plane_stride) *reinterpret_cast<__nv_bfloat162*>(q + po.plane_stride) = lo;
    if (w + 2 == g.W) {                                          // last pair of the row: zero the pad columns
        const __nv_bfloat162 z = __floats2bfloat162_rn(0.f, 0.f);
        for (int k = 2; k < po.Wp - w; k += 2) {
            *reinterpret_cast<__nv_bfloat162*>(q + k) = z;
            if (po.plane_stride) *reinterpret_cast<__nv_bfloat162*>(q + po.plane_stride + k) = z;
        }
    }
}

// grid (B*C planes, pair segments); no residual
__global__ void __launch_bounds__(BN_THREADS) bn_apply_packed_kernel(const float* __restrict__ x, const float* __restrict__ affine,
                                                                    PackedOut pk, BnGeom g, FastDiv d_w2) {
    const int plane = blockIdx.x;
    const int c = plane % g.C;
    const float sc = __ldg(affine + 2 * c), sh = __ldg(affine + 2 * c + 1);
    const float2* px = reinterpret_cast<const float2*>(x + (size_t)plane * g.HW);
    const int n_pairs = g.HW >> 1;
    const int j0 = blockIdx.y * (int)blockDim.x * BN_PAIRS + threadIdx.x;
    constexpr int PB = BN_BATCH / 2;                                 // pairs per load batch
#pragma unroll 1
    for (int part = 0; part < BN_PAIRS / PB; ++part) {
        const int jb = j0 + part * PB * (int)blockDim.x;
        if (jb - (int)threadIdx.x >= n_pairs) break;
        float2 xv[PB];
#pragma unroll
        for (int u = 0; u < PB; ++u) {
            const int j = jb + u * (int)blockDim.x;
            xv[u] = j < n_pairs ? __ldg(px + j) : make_float2(0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < PB; ++u) {
            const int j = jb + u * (int)blockDim.x;
            if (j < n_pairs) {
                float a = fmaf(xv[u].x, sc, sh), b = fmaf(xv[u].y, sc, sh);
                if (g.relu) { a = fmaxf(a, 0.f); b = fmaxf(b, 0.f); }
                int h, wp;
                d_w2.divmod(j, h, wp);
                packed_store2(pk, g, plane, h, 2 * wp, a, b);
            }
        }
    }
}

template <bool MASK>
__global__ void __launch_bounds__(BN_THREADS) bn_bwd_apply_packed_kernel(
    const float* __restrict__ dout, const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
    const float* __restrict__ save_mean, const float* __restrict__ save_rstd, const float* __restrict__ res,
    const uint32_t* __restrict__ mask, const double* __restrict__ sums2, PackedOut pk, float* __restrict__ dx_sum,
    float* __restrict__ dgamma,
    float* __restrict__ dbeta, float* __restrict__ d_res, BnGeom g, FastDiv d_w2, double count, int training) {
    const int plane = blockIdx.x;
    const int c = plane % g.C;
    const float mean = __ldg(save_mean + c), rstd = __ldg(save_rstd + c);
    const float gam = gamma ? __ldg(gamma + c) : 1.f, bet = beta ? __ldg(beta + c) : 0.f;
    const double sg = sums2[2 * c], sgx = sums2[2 * c + 1];
    if (plane < g.C && blockIdx.y == 0 && threadIdx.x == 0) {
        if (dbeta) dbeta[c] = (float)sg;
        if (dgamma) dgamma[c] = (float)sgx;
    }
    const float m1 = training ? (float)(sg / count) : 0.f;
    const float m2 = training ? (float)(sgx / count) : 0.f;
    const float k = gam * rstd;
    const float2* px = reinterpret_cast<const float2*>(x + (size_t)plane * g.HW);
    const float2* pd = reinterpret_cast<const float2*>(dout + (size_t)plane * g.HW);
    const size_t roff = (size_t)plane * g.RH * g.RW + (size_t)g.roh * g.RW + g.row;
    const float* pr = res ? res + roff : nullptr;
    float* pdr = d_res ? d_res + roff : nullptr;
    const int n_pairs = g.HW >> 1;
    const int j0 = blockIdx.y * (int)blockDim.x * BN_PAIRS + threadIdx.x;
    const bool need_res = !MASK && pr != nullptr && g.outer_relu;
    const uint32_t* pm = MASK ? mask + (size_t)plane * ((g.HW + 31) >> 5) : nullptr;
    constexpr int PB = BN_BATCH / 2;                                 // pairs per load batch
    float total = 0.f;
#pragma unroll 1
    for (int part = 0; part < BN_PAIRS / PB; ++part) {
        const int jb = j0 + part * PB * (int)blockDim.x;
        if (jb - (int)threadIdx.x >= n_pairs) break;
        float2 xv[PB], dv[PB], rv[MASK ? 1 : PB];
        int hh[PB], ww[PB];
        uint32_t ob = 0;                                             // MASK: bits 2u, 2u + 1 = saved bits of pair u
#pragma unroll
        for (int u = 0; u < PB; ++u) {
            const int j = jb + u * (int)blockDim.x;
            const bool in = j < n_pairs;
            xv[u] = in ? __ldg(px + j) : make_float2(0.f, 0.f);
            dv[u] = in ? __ldg(pd + j) : make_float2(0.f, 0.f);
            if constexpr (!MASK) rv[MASK ? 0 : u] = make_float2(0.f, 0.f);
            hh[u] = 0; ww[u] = 0;
            if constexpr (MASK) {
                if (in) ob |= ((__ldg(pm + (j >> 4)) >> ((2 * j) & 31)) & 3u) << (2 * u);   // both bits of the pair
            }
            if (in) {
                int wp;
                d_w2.divmod(j, hh[u], wp);
                ww[u] = 2 * wp;
                if (need_res) {
                    const float* q = pr + (size_t)hh[u] * g.RW + ww[u];
                    rv[MASK ? 0 : u] = make_float2(__ldg(q), __ldg(q + 1));
                }
            }
        }
#pragma unroll
        for (int u = 0; u < PB; ++u) {
            const int j = jb + u * (int)blockDim.x;
            if (j < n_pairs) {
                float xh0, xh1, g10, g11;
                const float2 r2 = MASK ? make_float2(0.f, 0.f) : rv[MASK ? 0 : u];
                const float g20 = bn_grad_in(dv[u].x, xv[u].x, mean, rstd, gam, bet, pr != nullptr, r2.x, g, xh0, g10,
                                             MASK ? (int)((ob >> (2 * u)) & 1u) : -1);
                const float g21 = bn_grad_in(dv[u].y, xv[u].y, mean, rstd, gam, bet, pr != nullptr, r2.y, g, xh1, g11,
                                             MASK ? (int)((ob >> (2 * u + 1)) & 1u) : -1);
                const float d0 = k * (g20 - m1 - xh0 * m2), d1 = k * (g21 - m1 - xh1 * m2);
                packed_store2(pk, g, plane, hh[u], ww[u], d0, d1);
                total += d0 + d1;
                if (pdr) {
                    float* q = pdr + (size_t)hh[u] * g.RW + ww[u];
                    q[0] = g10; q[1] = g11;
                }
            }
        }
    }
    if (dx_sum) {
        // bias gradient of the conv in front of this batch norm = per-channel sum of dx (analytically zero in training
        // mode: what the reference accumulates there is rounding noise, and so is this)
        __shared__ float red[BN_THREADS / 32];
        total = warp_sum(total);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = total;
        __syncthreads();
        if (threadIdx.x < 32) {
            float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
            v = warp_sum(v);
            if (threadIdx.x == 0) atomicAdd(dx_sum + c, v);
        }
    }
}

}  // namespace cpc

using namespace cpc;

// workspace: [C*2 doubles: sums][C*2 floats: affine]
extern "C" size_t cpc_bn_relu_workspace_bytes(const cpc_bn_params* p) {
    if (bn_validate(p) != CPC_OK) return 0;
    return align_up(sizeof(double) * 2 * (size_t)p->channels, 256) + align_up(sizeof(float) * 2 * (size_t)p->channels, 256);
}

// Bytes of the outer-ReLU bit mask of (p): 0 when there is nothing to save (no residual / no outer ReLU) or when the
// small-plane kernels serve this shape (they recompute).
extern "C" size_t cpc_bn_mask_bytes(const cpc_bn_params* p) {
    if (bn_validate(p) != CPC_OK || p->res_height == 0 || !p->outer_relu) return 0;
    const int64_t hw = (int64_t)p->height * p->width;
    if (hw < BN_SMALL_HW) return 0;
    return sizeof(uint32_t) * (size_t)p->batch * p->channels * ((hw + 31) / 32);
}

static int bn_fwd_impl(const float* x, const float* gamma, const float* beta, float* running_mean, float* running_var,
                       const float* residual, float* out, void* packed_out, float* save_mean, float* save_rstd,
                       const cpc_bn_params* p, void* workspace, size_t workspace_bytes, void* stream, void* relu_mask = nullptr) {
    int st = bn_validate(p);
    if (relu_mask && cpc_bn_mask_bytes(p) == 0) return CPC_ERR_UNSUPPORTED;
    if (st != CPC_OK) return st;
    if (!x || (!out && !packed_out) || !save_mean || !save_rstd) return CPC_ERR_NULL;
    if (!p->training && (!running_mean || !running_var)) return CPC_ERR_NULL;
    if ((p->res_height > 0) != (residual != nullptr)) return CPC_ERR_NULL;
    if (packed_out && (p->width % 2 != 0 || residual)) return CPC_ERR_UNSUPPORTED;
    if (packed_out && ((reinterpret_cast<uintptr_t>(packed_out) & 15) != 0 || (reinterpret_cast<uintptr_t>(x) & 7) != 0))
        return CPC_ERR_ALIGNMENT;
    const size_t need = cpc_bn_relu_workspace_bytes(p);
    if (!workspace || workspace_bytes < need) return CPC_ERR_WORKSPACE;
    if ((reinterpret_cast<uintptr_t>(workspace) & 7) != 0) return CPC_ERR_ALIGNMENT;
    if ((st = check_device()) != CPC_OK) return st;
    cudaStream_t s = (cudaStream_t)stream;
    BnGeom g = bn_geom(p);
    double* sums = reinterpret_cast<double*>(workspace);
    float* affine = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + align_up(sizeof(double) * 2 * (size_t)g.C, 256));
    const int nt = bn_block_threads(g.HW);
    const dim3 grid(g.B * g.C, ceil_div(g.HW, nt * BN_PER_THREAD));
    int launches = 2;
    const bool small = g.HW < BN_SMALL_HW && !packed_out && (int64_t)g.B * g.HW < 65535ll * BN_SMALL_CHUNK;
    const dim3 sgrid(g.C, ceil_div(g.B * g.HW, BN_SMALL_CHUNK));
    if (p->training) {
        if (cudaMemsetAsync(sums, 0, sizeof(double) * 2 * (size_t)g.C, s) != cudaSuccess) return CPC_ERR_CUDA;
        if (small) {
            bn_small_stats_kernel<<<sgrid, BN_THREADS, 0, s>>>(x, sums, g, FastDiv(g.HW));
        } else {
            const dim3 rgrid(g.B * g.C, ceil_div(g.HW, nt * BN_PER_THREAD * BN_RED_SEGS));
            bn_stats_kernel<<<rgrid, nt, 0, s>>>(x, sums, g.C, g.HW);
        }
        CPC_LAUNCH_CHECK();
        ++launches;
    }
    bn_finalize_kernel<<<ceil_div(g.C, 128), 128, 0, s>>>(sums, gamma, beta, running_mean, running_var, save_mean, save_rstd,
                                                         affine, g.C, (double)g.B * g.HW, p->eps, p->momentum, p->training);
    CPC_LAUNCH_CHECK();
    if (packed_out) {
        const int Wp = (g.W + 7) & ~7;
        PackedOut pk{reinterpret_cast<__nv_bfloat16*>(packed_out), p->packed_planes == 1 ? 0l : (long)g.B * g.C * g.H * Wp, Wp};
        const dim3 pgrid(g.B * g.C, ceil_div(g.HW / 2, nt * BN_PAIRS));
        bn_apply_packed_kernel<<<pgrid, nt, 0, s>>>(x, affine, pk, g, FastDiv(g.W / 2));
    } else if (small) {
        bn_small_apply_kernel<<<sgrid, BN_THREADS, 0, s>>>(x, affine, residual, out, g, FastDiv(g.HW));
    } else {
        if (relu_mask)
            bn_apply_kernel<true><<<grid, nt, 0, s>>>(x, affine, residual, out, reinterpret_cast<uint32_t*>(relu_mask), g);
        else
            bn_apply_kernel<false><<<grid, nt, 0, s>>>(x, affine, residual, out, nullptr, g);
    }
    CPC_LAUNCH_CHECK();
    count_launch(launches);
    return CPC_OK;
}

extern "C" int cpc_bn_relu_fwd(const float* x, const float* gamma, const float* beta, float* running_mean,
                               float* running_var, const float* residual, float* out, float* save_mean, float* save_rstd,
                               const cpc_bn_params* p, void* workspace, size_t workspace_bytes, void* stream) {
    if (!out) return CPC_ERR_NULL;
    return bn_fwd_impl(x, gamma, beta, running_mean, running_var, residual, out, nullptr, save_mean, save_rstd, p, workspace,
                       workspace_bytes, stream);
}

extern "C" int cpc_bn_relu_fwd_mask(const float* x, const float* gamma, const float* beta, float* running_mean,
                                    float* running_var, const float* residual, float* out, float* save_mean,
                                    float* save_rstd, void* relu_mask, const cpc_bn_params* p, void* workspace,
                                    size_t workspace_bytes, void* stream) {
    if (!out) return CPC_ERR_NULL;
    return bn_fwd_impl(x, gamma, beta, running_mean, running_var, residual, out, nullptr, save_mean, save_rstd, p, workspace,
                       workspace_bytes, stream, relu_mask);
}

extern "C" int cpc_bn_relu_fwd_packed(const float* x, const float* gamma, const float* beta, float* running_mean,
                                      float* running_var, const float* residual, void* packed_out, float* save_mean,
                                      float* save_rstd, const cpc_bn_params* p, void* workspace, size_t workspace_bytes,
                                      void* stream) {
    if (!packed_out) return CPC_ERR_NULL;
    return bn_fwd_impl(x, gamma, beta, running_mean, running_var, residual, nullptr, packed_out, save_mean, save_rstd, p,
                       workspace, workspace_bytes, stream);
}

extern "C" size_t cpc_bn_packed_bytes(const cpc_bn_params* p) {
    if (bn_validate(p) != CPC_OK) return 0;
    const size_t Wp = (size_t)((p->width + 7) & ~7);
    return (p->packed_planes == 1 ? 1 : 2) * (size_t)p->batch * p->channels * p->height * Wp * sizeof(__nv_bfloat16);
}

static int bn_bwd_impl(const float* dout, const float* x, const float* gamma, const float* beta, const float* save_mean,
                       const float* save_rstd, const float* residual, float* dx, void* packed_dx, float* dx_sum,
                       float* dgamma, float* dbeta, float* d_residual, const cpc_bn_params* p, void* workspace,
                       size_t workspace_bytes, void* stream, const void* relu_mask = nullptr) {
    int st = bn_validate(p);
    if (st != CPC_OK) return st;
    if (relu_mask && cpc_bn_mask_bytes(p) == 0) return CPC_ERR_UNSUPPORTED;
    const uint32_t* mask = reinterpret_cast<const uint32_t*>(relu_mask);
    if (!dout || !x || !save_mean || !save_rstd || (!dx && !packed_dx)) return CPC_ERR_NULL;
    if ((p->res_height > 0) != (residual != nullptr)) return CPC_ERR_NULL;
    if (d_residual && !residual) return CPC_ERR_NULL;
    if (packed_dx && p->width % 2 != 0) return CPC_ERR_UNSUPPORTED;
    if (packed_dx && ((reinterpret_cast<uintptr_t>(packed_dx) & 15) != 0 ||
                      ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dout)) & 7) != 0))
        return CPC_ERR_ALIGNMENT;
    const size_t need = cpc_bn_relu_workspace_bytes(p);
    if (!workspace || workspace_bytes < need) return CPC_ERR_WORKSPACE;
    if ((reinterpret_cast<uintptr_t>(workspace) & 7) != 0) return CPC_ERR_ALIGNMENT;
    if ((st = check_device()) != CPC_OK) return st;
    cudaStream_t s = (cudaStream_t)stream;
    BnGeom g = bn_geom(p);
    double* sums2 = reinterpret_cast<double*>(workspace);
    if (cudaMemsetAsync(sums2, 0, sizeof(double) * 2 * (size_t)g.C, s) != cudaSuccess) return CPC_ERR_CUDA;
    if (packed_dx && dx_sum && cudaMemsetAsync(dx_sum, 0, sizeof(float) * (size_t)g.C, s) != cudaSuccess) return CPC_ERR_CUDA;
    const bool crop = g.RH != g.H || g.RW != g.W;
    if (d_residual && crop &&
        cudaMemsetAsync(d_residual, 0, sizeof(float) * (size_t)g.B * g.C * g.RH * g.RW, s) != cudaSuccess)
        return CPC_ERR_CUDA;
    const int nt = bn_block_threads(g.HW);
    const dim3 grid(g.B * g.C, ceil_div(g.HW, nt * BN_PER_THREAD));
    const bool small = g.HW < BN_SMALL_HW && !packed_dx && (int64_t)g.B * g.HW < 65535ll * BN_SMALL_CHUNK;
    const dim3 sgrid(g.C, ceil_div(g.B * g.HW, BN_SMALL_CHUNK));
    if (small) {
        bn_small_bwd_reduce_kernel<<<sgrid, BN_THREADS, 0, s>>>(dout, x, gamma, beta, save_mean, save_rstd, residual, sums2, g,
                                                               FastDiv(g.HW));
    } else {
        const dim3 rgrid(g.B * g.C, ceil_div(g.HW, nt * BN_PER_THREAD * BN_RED_SEGS));
        if (mask)
            bn_bwd_reduce_kernel<true><<<rgrid, nt, 0, s>>>(dout, x, gamma, beta, save_mean, save_rstd, residual, mask, sums2, g);
        else
            bn_bwd_reduce_kernel<false><<<rgrid, nt, 0, s>>>(dout, x, gamma, beta, save_mean, save_rstd, residual, mask, sums2, g);
    }
    CPC_LAUNCH_CHECK();
    if (packed_dx) {
        const int Wp = (g.W + 7) & ~7;
        PackedOut pk{reinterpret_cast<__nv_bfloat16*>(packed_dx), p->packed_planes == 1 ? 0l : (long)g.B * g.C * g.H * Wp, Wp};
        const dim3 pgrid(g.B * g.C, ceil_div(g.HW / 2, nt * BN_PAIRS));
        auto* kern = mask ? bn_bwd_apply_packed_kernel<true> : bn_bwd_apply_packed_kernel<false>;
        kern<<<pgrid, nt, 0, s>>>(dout, x, gamma, beta, save_mean, save_rstd, residual, mask, sums2, pk, dx_sum, dgamma,
                                         dbeta, d_residual, g, FastDiv(g.W / 2), (double)g.B * g.HW, p->training);
    } else if (small) {
        bn_small_bwd_apply_kernel<<<sgrid, BN_THREADS, 0, s>>>(dout, x, gamma, beta, save_mean, save_rstd, residual, sums2, dx,
                                                              dgamma, dbeta, d_residual, g, FastDiv(g.HW), (double)g.B * g.HW,
                                                              p->training);
    } else {
        auto* kern = mask ? bn_bwd_apply_kernel<true> : bn_bwd_apply_kernel<false>;
        kern<<<grid, nt, 0, s>>>(dout, x, gamma, beta, save_mean, save_rstd, residual, mask, sums2, dx, dgamma, dbeta,
                                        d_residual, g, (double)g.B * g.HW, p->training);
    }
    CPC_LAUNCH_CHECK();
    count_launch(2);
    return CPC_OK;
}

extern "C" int cpc_bn_relu_bwd(const float* dout, const float* x, const float* gamma, const float* beta,
                               const float* save_mean, const float* save_rstd, const float* residual, float* dx,
                               float* dgamma, float* dbeta, float* d_residual, const cpc_bn_params* p, void* workspace,
                               size_t workspace_bytes, void* stream) {
    if (!dx) return CPC_ERR_NULL;
    return bn_bwd_impl(dout, x, gamma, beta, save_mean, save_rstd, residual, dx, nullptr, nullptr, dgamma, dbeta, d_residual,
                       p, workspace, workspace_bytes, stream);
}

extern "C" int cpc_bn_relu_bwd_packed(const float* dout, const float* x, const float* gamma, const float* beta,
                                      const float* save_mean, const float* save_rstd, const float* residual,
                                      void* packed_dx, float* dx_sum, float* dgamma, float* dbeta, float* d_residual,
                                      const cpc_bn_params* p, void* workspace, size_t workspace_bytes, void* stream) {
    if (!packed_dx) return CPC_ERR_NULL;
    return bn_bwd_impl(dout, x, gamma, beta, save_mean, save_rstd, residual, nullptr, packed_dx, dx_sum, dgamma, dbeta,
                       d_residual, p, workspace, workspace_bytes, stream);
}

extern "C" int cpc_bn_relu_bwd_mask(const float* dout, const float* x, const float* gamma, const float* beta,
                                    const float* save_mean, const float* save_rstd, const float* residual,
                                    const void* relu_mask, float* dx, float* dgamma, float* dbeta, float* d_residual,
                                    const cpc_bn_params* p, void* workspace, size_t workspace_bytes, void* stream) {
    if (!dx) return CPC_ERR_NULL;
    return bn_bwd_impl(dout, x, gamma, beta, save_mean, save_rstd, residual, dx, nullptr, nullptr, dgamma, dbeta, d_residual,
                       p, workspace, workspace_bytes, stream, relu_mask);
}

extern "C" int cpc_bn_relu_bwd_packed_mask(const float* dout, const float* x, const float* gamma, const float* beta,
                                           const float* save_mean, const float* save_rstd, const float* residual,
                                           const void* relu_mask, void* packed_dx, float* dx_sum, float* dgamma,
                                           float* dbeta, float* d_residual, const cpc_bn_params* p, void* workspace,
                                           size_t workspace_bytes, void* stream) {
    if (!packed_dx) return CPC_ERR_NULL;
    return bn_bwd_impl(dout, x, gamma, beta, save_mean, save_rstd, residual, nullptr, packed_dx, dx_sum, dgamma, dbeta,
                       d_residual, p, workspace, workspace_bytes, stream, relu_mask);
}
