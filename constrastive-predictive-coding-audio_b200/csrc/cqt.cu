// Constant-Q front end: complex filterbank correlation + fused log-power / phase-difference / pooling.
// Replaces CQT.forward (constant_q_transform.py:161-172) and PreprocessingModule.forward
// (scalogram_model.py:75-102).
//
// Stage 1: all octave groups in one launch.  Group g is a GEMM  [B*T frames] x [2 n_g channels] over
//          K_g taps; frame t of group g reads x[b, off_g + hop*t + n], off_g = (K_0 - K_g)/2, so every
//          group is centred on the same sample and none reads the last sample of the item.
// Stage 2: one elementwise pass turning (re, im) into the scalogram the trainer consumes.
#include "common.cuh"

namespace cpc {

struct CqtGroups {
    int n_groups;
    int ksize[CPC_CQT_MAX_GROUPS], lo[CPC_CQT_MAX_GROUPS], hi[CPC_CQT_MAX_GROUPS], off[CPC_CQT_MAX_GROUPS];
    long long woff[CPC_CQT_MAX_GROUPS];
};

struct FrameRows {   // rows: (b, t); k: tap
    static constexpr bool kFast = true;
    const float* x; int M, K, T, hop, pitch, off; FastDiv d_t;
    __device__ __forceinline__ float load(int m, int k) const {
        if (m >= M || k >= K) return 0.f;
        int b, t;
        d_t.divmod(m, b, t);
        return __ldg(x + (size_t)b * pitch + off + (size_t)t * hop + k);
    }
};
struct FilterRows {
    static constexpr bool kFast = true;
    const float* w; int N, K;
    __device__ __forceinline__ float load(int n, int k) const {
        if (n >= N || k >= K) return 0.f;
        return __ldg(w + (size_t)n * K + k);
    }
};

// grid (frame tiles, groups); complex out (B, F, T, 2)
// cplx holds bins [f0, f0 + F) only: (B, F, T, 2); groups [g_first, g_first + gridDim.y) are computed
__global__ void __launch_bounds__(TILE_THREADS) cqt_filterbank_kernel(const float* __restrict__ x,
                                                                     const float* __restrict__ weights,
                                                                     float* __restrict__ cplx, CqtGroups gr, int B, int T,
                                                                     int F, int hop, int pitch, int g_first, int f0) {
    __shared__ TileSmem sm;
    const int g = blockIdx.y + g_first;
    const int ng = gr.hi[g] - gr.lo[g];
    const int M = B * T;
    FrameRows la{x, M, gr.ksize[g], T, hop, pitch, gr.off[g], FastDiv(T)};
    FilterRows lb{weights + gr.woff[g], 2 * ng, gr.ksize[g]};
    const int row0 = blockIdx.x * TILE;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    for (int col0 = 0; col0 < 2 * ng; col0 += TILE) {
        float acc[4][4] = {};
        tile_gemm(la, lb, row0, col0, 0, la.K, acc, sm);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int m = row0 + tx * 4 + i;
            if (m >= M) continue;
            int b, t;
            la.d_t.divmod(m, b, t);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int n = col0 + ty * 4 + j;
                if (n >= 2 * ng) continue;
                const int part = n >= ng;
                const int bin = gr.lo[g] + (part ? n - ng : n) - f0;
                cplx[(((size_t)b * F + bin) * T + t) * 2 + part] = acc[i][j];
            }
        }
    }
}

// One thread per output element (b, f, to) for the Fs bins [f0, f0 + Fs) held by cplx (B, Fs, T, 2).
// mode 1: out (B,1,F,To); mode 2: out (B,2,F,To).
__global__ void __launch_bounds__(256) cqt_scalogram_kernel(const float* __restrict__ cplx,
                                                           const float* __restrict__ phase_fixed,
                                                           const float* __restrict__ phase_scale, float* __restrict__ out,
                                                           int B, int F, int T, int To, int mode, int pool, float eps,
                                                           float log_offset, float norm, float power, int f0, int Fs) {
    const long total = (long)B * Fs * To;
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int to = (int)(idx % To);
    const long bf = idx / To;
    const int fl = (int)(bf % Fs);
    const int f = f0 + fl;
    const int b = (int)(bf / Fs);
    const float2* z = reinterpret_cast<const float2*>(cplx) + ((size_t)b * Fs + fl) * T;
    const float kPi = 3.14159265358979323846f;
    float amp = -INFINITY, ph = -INFINITY;
    for (int q = 0; q < pool; ++q) {
        const int tt = to * pool + q;           // index in the un-pooled output time axis
        if (mode == CPC_CQT_LOGPOW) {
            const float2 c = __ldg(z + tt);
            // (sqrt(re^2+im^2))^2 as the reference computes it (constant_q_transform.py:43, scalogram_model.py:88)
            const float a = sqrtf(c.x * c.x + c.y * c.y);
            amp = fmaxf(amp, logf(a * a + eps) + log_offset);
        } else {
            const float2 c1 = __ldg(z + tt + 1), c0 = __ldg(z + tt);
            const float a = sqrtf(c1.x * c1.x + c1.y * c1.y);
            amp = fmaxf(amp, logf(a * a + eps) + log_offset);
            float pd = atan2f(c1.y, c1.x) - atan2f(c0.y, c0.x) + __ldg(phase_fixed + f);
            if (pd > kPi) pd -= 2.f * kPi;
            if (pd < -kPi) pd += 2.f * kPi;
            ph = fmaxf(ph, pd * __ldg(phase_scale + f));
        }
    }
    amp *= norm;
    if (power != 1.f) amp = powf(amp, power);
    if (mode == CPC_CQT_LOGPOW) {
        out[((size_t)b * F + f) * To + to] = amp;
    } else {
        ph *= norm;
        if (power != 1.f) ph = powf(ph, power);
        out[(((size_t)b * 2 + 0) * F + f) * To + to] = amp;
        out[(((size_t)b * 2 + 1) * F + f) * To + to] = ph;
    }
}

static int cqt_validate(const cpc_cqt_params* p) {
    if (!p) return CPC_ERR_NULL;
    if (p->batch <= 0 || p->n_samples <= 0 || p->x_pitch < p->n_samples || p->n_bins <= 0 || p->hop <= 0 ||
        p->n_frames <= 0 || p->n_groups <= 0 || p->n_groups > CPC_CQT_MAX_GROUPS)
        return CPC_ERR_BAD_SHAPE;
    if (p->mode < 0 || p->mode > 2 || (p->pool_t != 1 && p->pool_t != 2)) return CPC_ERR_BAD_SHAPE;
    const int k0 = p->kernel_size[0];
    if ((int64_t)(p->n_frames - 1) * p->hop + k0 > (int64_t)p->n_samples - 1) return CPC_ERR_BAD_SHAPE;
    int next = 0;
    for (int g = 0; g < p->n_groups; ++g) {
        if (p->kernel_size[g] <= 0 || p->kernel_size[g] > k0 || ((k0 - p->kernel_size[g]) & 1)) return CPC_ERR_BAD_SHAPE;
        if (p->bin_lo[g] != next || p->bin_hi[g] <= p->bin_lo[g]) return CPC_ERR_BAD_SHAPE;
        next = p->bin_hi[g];
    }
    if (next != p->n_bins) return CPC_ERR_BAD_SHAPE;
    if ((int64_t)p->batch * p->n_frames > (1ll << 31) - 1) return CPC_ERR_BAD_SHAPE;
    const int t_eff = p->mode == CPC_CQT_LOGPOW_PHASE ? p->n_frames - 1 : p->n_frames;
    if (p->mode != CPC_CQT_COMPLEX && t_eff / p->pool_t <= 0) return CPC_ERR_BAD_SHAPE;
    return CPC_OK;
}

}  // namespace cpc

using namespace cpc;

// cqt_umma.cu
namespace cpc {
bool cqt_umma_eligible(const cpc_cqt_params* p);
int cqt_umma_tensor_groups(const cpc_cqt_params* p);
size_t cqt_umma_workspace(const cpc_cqt_params* p);
size_t cqt_umma_packed_filter_bytes(const cpc_cqt_params* p);
int cqt_umma_pack_filters(const float* weights, void* packed, const cpc_cqt_params* p, cudaStream_t s);
int cqt_umma_launch(const float* x, const float* weights, const void* packed_filters, const float* phase_fixed,
                    const float* phase_scale, float* out, const cpc_cqt_params* p, void* workspace, size_t workspace_bytes,
                    cudaStream_t s);
}

// flags & CPC_CQT_FLAG_NO_TENSOR keeps every group on the CUDA-core kernels (A/B switch for tests)
static bool tensor_cqt(const cpc_cqt_params* p) {
    if (p->flags & CPC_CQT_FLAG_NO_TENSOR) return false;
    return cqt_umma_eligible(p);
}

// bytes of the complex intermediate the CUDA-core path needs for groups [g_first, n_groups)
static size_t simt_cplx_bytes(const cpc_cqt_params* p, int g_first) {
    if (p->mode == CPC_CQT_COMPLEX || g_first >= p->n_groups) return 0;
    const int fs = p->n_bins - p->bin_lo[g_first];
    return align_up(sizeof(float) * 2 * (size_t)p->batch * fs * p->n_frames, 1024);
}

extern "C" size_t cpc_cqt_workspace_bytes(const cpc_cqt_params* p) {
    if (!p || cqt_validate(p) != CPC_OK) return 0;
    if (tensor_cqt(p)) return cqt_umma_workspace(p) + simt_cplx_bytes(p, cqt_umma_tensor_groups(p)) + 1024;
    return simt_cplx_bytes(p, 0);
}

extern "C" size_t cpc_cqt_packed_filter_bytes(const cpc_cqt_params* p) {
    if (!p || cqt_validate(p) != CPC_OK || !tensor_cqt(p)) return 0;
    return cqt_umma_packed_filter_bytes(p);
}

extern "C" int cpc_cqt_pack_filters(const float* weights, void* packed, const cpc_cqt_params* p, void* stream) {
    int st = cqt_validate(p);
    if (st != CPC_OK) return st;
    if (!weights || !packed) return CPC_ERR_NULL;
    if (!tensor_cqt(p)) return CPC_ERR_UNSUPPORTED;
    if ((st = check_device()) != CPC_OK) return st;
    return cqt_umma_pack_filters(weights, packed, p, (cudaStream_t)stream);
}

extern "C" int cpc_cqt_fwd_ex(const float* x, const float* weights, const void* packed_filters, const float* phase_fixed,
                              const float* phase_scale, float* out, const cpc_cqt_params* p, void* workspace,
                              size_t workspace_bytes, void* stream) {
    int st = cqt_validate(p);
    if (st != CPC_OK) return st;
    if (!x || !weights || !out) return CPC_ERR_NULL;
    if (p->mode == CPC_CQT_LOGPOW_PHASE && (!phase_fixed || !phase_scale)) return CPC_ERR_NULL;
    const size_t need = cpc_cqt_workspace_bytes(p);
    if (need && (!workspace || workspace_bytes < need)) return CPC_ERR_WORKSPACE;
    if ((st = check_device()) != CPC_OK) return st;
    cudaStream_t s = (cudaStream_t)stream;
    CqtGroups gr;
    gr.n_groups = p->n_groups;
    for (int g = 0; g < p->n_groups; ++g) {
        gr.ksize[g] = p->kernel_size[g];
        gr.lo[g] = p->bin_lo[g];
        gr.hi[g] = p->bin_hi[g];
        gr.off[g] = (p->kernel_size[0] - p->kernel_size[g]) / 2;
        gr.woff[g] = p->weight_offset[g];
    }
    int g_first = 0;
    uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
    if (tensor_cqt(p)) {
        // tensor cores, written in the final format
        const size_t uw = cqt_umma_workspace(p);
        st = cqt_umma_launch(x, weights, packed_filters, phase_fixed, phase_scale, out, p, ws, uw, s);
        if (st != CPC_OK) return st;
        g_first = cqt_umma_tensor_groups(p);
        ws = reinterpret_cast<uint8_t*>(align_up(reinterpret_cast<uintptr_t>(ws) + uw, 1024));
        if (g_first >= p->n_groups) return CPC_OK;
    }
    // remaining groups: CUDA-core filterbank + elementwise pass over their bins
    const int f0 = p->bin_lo[g_first], fs = p->n_bins - f0;
    const bool direct = p->mode == CPC_CQT_COMPLEX;
    float* cplx = direct ? out : reinterpret_cast<float*>(ws);
    const int M = p->batch * p->n_frames;
    dim3 grid(ceil_div(M, TILE), p->n_groups - g_first);
    cqt_filterbank_kernel<<<grid, TILE_THREADS, 0, s>>>(x, weights, cplx, gr, p->batch, p->n_frames,
                                                       direct ? p->n_bins : fs, p->hop, p->x_pitch, g_first, direct ? 0 : f0);
    CPC_LAUNCH_CHECK();
    count_launch();
    if (!direct) {
        const int t_eff = p->mode == CPC_CQT_LOGPOW_PHASE ? p->n_frames - 1 : p->n_frames;
        const int to = t_eff / p->pool_t;
        const long total = (long)p->batch * fs * to;
        cqt_scalogram_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(
            cplx, phase_fixed, phase_scale, out, p->batch, p->n_bins, p->n_frames, to, p->mode, p->pool_t, p->eps,
            p->log_offset, p->norm, p->power, f0, fs);
        CPC_LAUNCH_CHECK();
        count_launch();
    }
    return CPC_OK;
}

extern "C" int cpc_cqt_fwd(const float* x, const float* weights, const float* phase_fixed, const float* phase_scale,
                           float* out, const cpc_cqt_params* p, void* workspace, size_t workspace_bytes, void* stream) {
    return cpc_cqt_fwd_ex(x, weights, nullptr, phase_fixed, phase_scale, out, p, workspace, workspace_bytes, stream);
}
