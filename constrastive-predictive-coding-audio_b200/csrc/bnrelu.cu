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
// Masks are recomputed from x (and the residual), so nothing but x, mean and rstd is saved for backward.
#include "common.cuh"

namespace cpc {

constexpr int BN_THREADS = 256;
constexpr int BN_PER_THREAD = 16;
constexpr int BN_SEG = BN_THREADS * BN_PER_THREAD;

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
        r = threadIdx.x < (BN_THREADS / 32) ? red[threadIdx.x] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
    }
    return r;
}

// grid (B*C planes, segments of one plane).  sums[c*2 + {0,1}] += sum x, sum x^2 (double).
__global__ void __launch_bounds__(BN_THREADS) bn_stats_kernel(const float* __restrict__ x, double* __restrict__ sums,
                                                             int C, int HW) {
    __shared__ double red[BN_THREADS / 32];
    const int plane = blockIdx.x;
    const int c = plane % C;
    const float* px = x + (size_t)plane * HW;
    const int i0 = blockIdx.y * BN_SEG + threadIdx.x;
    float s = 0.f, q = 0.f;
#pragma unroll
    for (int u = 0; u < BN_PER_THREAD; ++u) {
        const int i = i0 + u * BN_THREADS;
        if (i < HW) { const float v = __ldg(px + i); s += v; q = fmaf(v, v, q); }
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
__global__ void __launch_bounds__(BN_THREADS) bn_apply_kernel(const float* __restrict__ x, const float* __restrict__ affine,
                                                             const float* __restrict__ res, float* __restrict__ out,
                                                             BnGeom g) {
    const int plane = blockIdx.x;
    const int c = plane % g.C;
    const float sc = __ldg(affine + 2 * c), sh = __ldg(affine + 2 * c + 1);
    const float* px = x + (size_t)plane * g.HW;
    float* po = out + (size_t)plane * g.HW;
    const float* pr = res ? res + (size_t)plane * g.RH * g.RW + (size_t)g.roh * g.RW + g.row : nullptr;
    const int i0 = blockIdx.y * BN_SEG + threadIdx.x;
    float xv[BN_PER_THREAD], rv[BN_PER_THREAD];
#pragma unroll
    for (int u = 0; u < BN_PER_THREAD; ++u) {
        const int i = i0 + u * BN_THREADS;
        xv[u] = i < g.HW ? __ldg(px + i) : 0.f;
        rv[u] = 0.f;
        if (pr && i < g.HW) {
            int h, w;
            g.d_w.divmod(i, h, w);
            rv[u] = __ldg(pr + (size_t)h * g.RW + w);
        }
    }
#pragma unroll
    for (int u = 0; u < BN_PER_THREAD; ++u) {
        const int i = i0 + u * BN_THREADS;
        float v = fmaf(xv[u], sc, sh);
        if (g.relu) v = fmaxf(v, 0.f);
        if (pr) {
            v += rv[u];
            if (g.outer_relu) v = fmaxf(v, 0.f);
        }
        if (i < g.HW) po[i] = v;
    }
}

// Gradient entering the normalisation: g2 = dout * [outer mask] * [inner mask]; also returns xhat.
__device__ __forceinline__ float bn_grad_in(float dout, float xv, float mean, float rstd, float gam, float bet, bool has_res,
                                            float resv, const BnGeom& g, float& xhat, float& g1) {
    xhat = (xv - mean) * rstd;
    const float y = fmaf(xhat, gam, bet);
    const float v = g.relu ? fmaxf(y, 0.f) : y;
    g1 = dout;
    if (has_res && g.outer_relu && !(v + resv > 0.f)) g1 = 0.f;
    return (g.relu && !(y > 0.f)) ? 0.f : g1;
}

// sums2[c*2 + {0,1}] += sum g2, sum g2 * xhat
__global__ void __launch_bounds__(BN_THREADS) bn_bwd_reduce_kernel(const float* __restrict__ dout, const float* __restrict__ x,
                                                                  const float* __restrict__ gamma,
                                                                  const float* __restrict__ beta,
                                                                  const float* __restrict__ save_mean,
                                                                  const float* __restrict__ save_rstd,
                                                                  const float* __restrict__ res, double* __restrict__ sums2,
                                                                  BnGeom g) {
    __shared__ double red[BN_THREADS / 32];
    const int plane = blockIdx.x;
    const int c = plane % g.C;
    const float mean = __ldg(save_mean + c), rstd = __ldg(save_rstd + c);
    const float gam = gamma ? __ldg(gamma + c) : 1.f, bet = beta ? __ldg(beta + c) : 0.f;
    const float* px = x + (size_t)plane * g.HW;
    const float* pd = dout + (size_t)plane * g.HW;
    const float* pr = res ? res + (size_t)plane * g.RH * g.RW + (size_t)g.roh * g.RW + g.row : nullptr;
    const int i0 = blockIdx.y * BN_SEG + threadIdx.x;
    float xv[BN_PER_THREAD], dv[BN_PER_THREAD], rv[BN_PER_THREAD];
    const bool need_res = pr != nullptr && g.outer_relu;
#pragma unroll
    for (int u = 0; u < BN_PER_THREAD; ++u) {
        const int i = i0 + u * BN_THREADS;
        const bool in = i < g.HW;
        xv[u] = in ? __ldg(px + i) : 0.f;
        dv[u] = in ? __ldg(pd + i) : 0.f;
        rv[u] = 0.f;
        if (need_res && in) {
            int h, w;
            g.d_w.divmod(i, h, w);
            rv[u] = __ldg(pr + (size_t)h * g.RW + w);
        }
    }
    float s = 0.f, q = 0.f;
#pragma unroll
    for (int u = 0; u < BN_PER_THREAD; ++u) {
        const int i = i0 + u * BN_THREADS;
        if (i < g.HW) {
            float xhat, g1;
            const float g2 = bn_grad_in(dv[u], xv[u], mean, rstd, gam, bet, pr != nullptr, rv[u], g, xhat, g1);
            s += g2;
            q = fmaf(g2, xhat, q);
        }
    }
    const double bs = block_sum_d((double)s, red);
    const double bq = block_sum_d((double)q, red);
    if (threadIdx.x == 0) { atomicAdd(sums2 + 2 * c, bs); atomicAdd(sums2 + 2 * c + 1, bq); }
}

// dx = gamma * rstd * (g2 - mean(g2) - xhat * mean(g2 * xhat))   [training]
// dx = gamma * rstd * g2                                           [running statistics]
// d_res (inside the crop) = g1;  dgamma / dbeta written by the first block of each channel of item 0.
__global__ void __launch_bounds__(BN_THREADS) bn_bwd_apply_kernel(const float* __restrict__ dout, const float* __restrict__ x,
                                                                 const float* __restrict__ gamma,
                                                                 const float* __restrict__ beta,
                                                                 const float* __restrict__ save_mean,
                                                                 const float* __restrict__ save_rstd,
                                                                 const float* __restrict__ res,
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
    const int i0 = blockIdx.y * BN_SEG + threadIdx.x;
    float xv[BN_PER_THREAD], dv[BN_PER_THREAD], rv[BN_PER_THREAD];
    int ro[BN_PER_THREAD];                                           // offset inside the residual plane
    const bool need_res = pr != nullptr && g.outer_relu;
#pragma unroll
    for (int u = 0; u < BN_PER_THREAD; ++u) {
        const int i = i0 + u * BN_THREADS;
        const bool in = i < g.HW;
        xv[u] = in ? __ldg(px + i) : 0.f;
        dv[u] = in ? __ldg(pd + i) : 0.f;
        rv[u] = 0.f;
        ro[u] = 0;
        if ((pr || pdr) && in) {
            int h, w;
            g.d_w.divmod(i, h, w);
            ro[u] = h * g.RW + w;
            if (need_res) rv[u] = __ldg(pr + ro[u]);
        }
    }
#pragma unroll
    for (int u = 0; u < BN_PER_THREAD; ++u) {
        const int i = i0 + u * BN_THREADS;
        if (i < g.HW) {
            float xhat, g1;
            const float g2 = bn_grad_in(dv[u], xv[u], mean, rstd, gam, bet, pr != nullptr, rv[u], g, xhat, g1);
            pdx[i] = k * (g2 - m1 - xhat * m2);
            if (pdr) pdr[ro[u]] = g1;
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

extern "C" int cpc_bn_relu_fwd(const float* x, const float* gamma, const float* beta, float* running_mean,
                               float* running_var, const float* residual, float* out, float* save_mean, float* save_rstd,
                               const cpc_bn_params* p, void* workspace, size_t workspace_bytes, void* stream) {
    int st = bn_validate(p);
    if (st != CPC_OK) return st;
    if (!x || !out || !save_mean || !save_rstd) return CPC_ERR_NULL;
    if (!p->training && (!running_mean || !running_var)) return CPC_ERR_NULL;
    if ((p->res_height > 0) != (residual != nullptr)) return CPC_ERR_NULL;
    const size_t need = cpc_bn_relu_workspace_bytes(p);
    if (!workspace || workspace_bytes < need) return CPC_ERR_WORKSPACE;
    if ((reinterpret_cast<uintptr_t>(workspace) & 7) != 0) return CPC_ERR_ALIGNMENT;
    if ((st = check_device()) != CPC_OK) return st;
    cudaStream_t s = (cudaStream_t)stream;
    BnGeom g = bn_geom(p);
    double* sums = reinterpret_cast<double*>(workspace);
    float* affine = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + align_up(sizeof(double) * 2 * (size_t)g.C, 256));
    const dim3 grid(g.B * g.C, ceil_div(g.HW, BN_SEG));
    int launches = 2;
    if (p->training) {
        if (cudaMemsetAsync(sums, 0, sizeof(double) * 2 * (size_t)g.C, s) != cudaSuccess) return CPC_ERR_CUDA;
        bn_stats_kernel<<<grid, BN_THREADS, 0, s>>>(x, sums, g.C, g.HW);
        CPC_LAUNCH_CHECK();
        ++launches;
    }
    bn_finalize_kernel<<<ceil_div(g.C, 128), 128, 0, s>>>(sums, gamma, beta, running_mean, running_var, save_mean, save_rstd,
                                                         affine, g.C, (double)g.B * g.HW, p->eps, p->momentum, p->training);
    CPC_LAUNCH_CHECK();
    bn_apply_kernel<<<grid, BN_THREADS, 0, s>>>(x, affine, residual, out, g);
    CPC_LAUNCH_CHECK();
    count_launch(launches);
    return CPC_OK;
}

extern "C" int cpc_bn_relu_bwd(const float* dout, const float* x, const float* gamma, const float* beta,
                               const float* save_mean, const float* save_rstd, const float* residual, float* dx,
                               float* dgamma, float* dbeta, float* d_residual, const cpc_bn_params* p, void* workspace,
                               size_t workspace_bytes, void* stream) {
    int st = bn_validate(p);
    if (st != CPC_OK) return st;
    if (!dout || !x || !save_mean || !save_rstd || !dx) return CPC_ERR_NULL;
    if ((p->res_height > 0) != (residual != nullptr)) return CPC_ERR_NULL;
    if (d_residual && !residual) return CPC_ERR_NULL;
    const size_t need = cpc_bn_relu_workspace_bytes(p);
    if (!workspace || workspace_bytes < need) return CPC_ERR_WORKSPACE;
    if ((reinterpret_cast<uintptr_t>(workspace) & 7) != 0) return CPC_ERR_ALIGNMENT;
    if ((st = check_device()) != CPC_OK) return st;
    cudaStream_t s = (cudaStream_t)stream;
    BnGeom g = bn_geom(p);
    double* sums2 = reinterpret_cast<double*>(workspace);
    if (cudaMemsetAsync(sums2, 0, sizeof(double) * 2 * (size_t)g.C, s) != cudaSuccess) return CPC_ERR_CUDA;
    const bool crop = g.RH != g.H || g.RW != g.W;
    if (d_residual && crop &&
        cudaMemsetAsync(d_residual, 0, sizeof(float) * (size_t)g.B * g.C * g.RH * g.RW, s) != cudaSuccess)
        return CPC_ERR_CUDA;
    const dim3 grid(g.B * g.C, ceil_div(g.HW, BN_SEG));
    bn_bwd_reduce_kernel<<<grid, BN_THREADS, 0, s>>>(dout, x, gamma, beta, save_mean, save_rstd, residual, sums2, g);
    CPC_LAUNCH_CHECK();
    bn_bwd_apply_kernel<<<grid, BN_THREADS, 0, s>>>(dout, x, gamma, beta, save_mean, save_rstd, residual, sums2, dx, dgamma,
                                                   dbeta, d_residual, g, (double)g.B * g.HW, p->training);
    CPC_LAUNCH_CHECK();
    count_launch(2);
    return CPC_OK;
}
