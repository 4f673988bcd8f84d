// Strided conv2d (conv1d = h 1) forward / dgrad / wgrad as implicit GEMM, NCHW fp32.
// Replaces nn.Conv1d / nn.Conv2d (+ZeroPad2d) of the reference encoders (audio_model.py:30-44,
// scalogram_model.py:155-201,387-431) and their autograd.
//
// This file is the fp32 CUDA-core implementation: the im2col gather is done on the fly by loader
// functors feeding a 64x64 register-tiled GEMM (common.cuh).  The tcgen05 implicit-GEMM path for
// channel counts >= 16 lives in conv_umma.cu and is selected by the dispatcher in this file.
#include "common.cuh"

#include <cuda_bf16.h>
#include <cstdlib>

namespace cpc {

// conv_umma.cu
size_t umma_conv_workspace(const cpc_conv_params* p, int which);
bool umma_conv_eligible(const cpc_conv_params* p, int which);
int umma_conv_launch(const float* in, const float* w, const float* bias, float* out, const cpc_conv_params* p, int which,
                     const void* pre, void* workspace, size_t workspace_bytes, cudaStream_t s);
int pack_split_launch(const float* x, __nv_bfloat16* out, long rows, int W, int Wp, int planes, int nrep, int w_mul,
                      int rep_mul, int w_off, cudaStream_t s);
int pack_split_sum_launch(const float* x, __nv_bfloat16* out, float* sums, long rows, int W, int Wp, int planes, int H, int C,
                          int nrep, cudaStream_t s);
int umma_dy_replicas(const cpc_conv_params* p);                  // conv_umma.cu

size_t umma_wgrad_workspace(const cpc_conv_params* p);
bool umma_wgrad_eligible(const cpc_conv_params* p);
int umma_wgrad_launch(const float* x, const float* dy, float* dw, const cpc_conv_params* p, const void* pre_x,
                      const void* pre_dy, void* workspace, size_t workspace_bytes, cudaStream_t s);

// conv_tall.cu: row-streaming kernels for kh x 1, stride-1, 32 -> 32 channel convolutions
size_t tall_conv_workspace(const cpc_conv_params* p, int which);
bool tall_conv_eligible(const cpc_conv_params* p, int which);
int tall_conv_launch(const float* in, const float* w, const float* bias, float* out, const cpc_conv_params* p, int which,
                     const void* pre, void* workspace, size_t workspace_bytes, cudaStream_t s);
int tall_wgrad_launch(const float* x, const float* dy, float* dw, const cpc_conv_params* p, const void* pre_x,
                      const void* pre_dy, void* workspace, size_t workspace_bytes, cudaStream_t s);

// conv_smallk.cu: direct kernels for C_in * kh * kw <= 36 (forward and weight gradient)
bool smallk_eligible(const cpc_conv_params* p, int which);
int smallk_launch(int which, const float* x, const float* w, const float* bias, const float* dy, float* out,
                  const cpc_conv_params* p, cudaStream_t s);
bool small_dgrad_eligible(const cpc_conv_params* p);
int small_dgrad_launch(const float* dy, const float* w, float* dx, const cpc_conv_params* p, cudaStream_t s);
static bool smallk_path(const cpc_conv_params* p, int which) {
    if (p->flags & CPC_CONV_FLAG_NO_SMALLK) return false;       // A/B switch: keep tiny-K convs on the tiled kernels
    return smallk_eligible(p, which);
}

// conv_tall128.cu: row-streaming kernels for kh x 1, stride-1 convolutions with 128 output channels
size_t tall128_workspace(const cpc_conv_params* p, int which);
bool tall128_eligible(const cpc_conv_params* p, int which);
int tall128_conv_launch(const float* in, const float* w, const float* bias, float* out, const cpc_conv_params* p, int which,
                        const void* pre, void* workspace, size_t workspace_bytes, cudaStream_t s);
int tall128_wgrad_launch(const float* x, const float* dy, float* dw, const cpc_conv_params* p, const void* pre_x,
                         const void* pre_dy, void* workspace, size_t workspace_bytes, cudaStream_t s);

// Debug switches (tests use them to A/B kernel families on one shape): CPC_FORCE_CUDA_CORE_CONV=1 selects the
// fp32 CUDA-core kernels, flags & CPC_CONV_FLAG_NO_TALL keeps tall convolutions on the generic tcgen05 kernel.
// Returns 0 (not a tall conv), 1 (conv_tall.cu, 32 -> 32 channels) or 2 (conv_tall128.cu).
static int tall_path(const cpc_conv_params* p, int which) {
    if (p->flags & (CPC_CONV_FLAG_CUDA_CORE | CPC_CONV_FLAG_NO_TALL)) return 0;
    if (tall_conv_eligible(p, which)) return 1;
    if (tall128_eligible(p, which)) return 2;
    return 0;
}

static bool tensor_core_path(const cpc_conv_params* p, int which) {
    if (p->flags & CPC_CONV_FLAG_CUDA_CORE) return false;
    return which == 2 ? umma_wgrad_eligible(p) : umma_conv_eligible(p, which);
}

struct ConvGeom {
    int B, Cin, H, W, Cout, OH, OW, kh, kw, sh, sw, pt, pl;
    int khkw, ohow, hw;
    FastDiv d_ow, d_ohow, d_kw, d_khkw, d_w, d_hw, d_sh, d_sw;
};

static ConvGeom make_geom(const cpc_conv_params* p) {
    ConvGeom g;
    g.B = p->batch; g.Cin = p->c_in; g.H = p->h_in; g.W = p->w_in;
    g.Cout = p->c_out; g.OH = p->h_out; g.OW = p->w_out;
    g.kh = p->kh; g.kw = p->kw; g.sh = p->stride_h; g.sw = p->stride_w;
    g.pt = p->pad_top; g.pl = p->pad_left;
    g.khkw = g.kh * g.kw; g.ohow = g.OH * g.OW; g.hw = g.H * g.W;
    g.d_ow = FastDiv(g.OW); g.d_ohow = FastDiv(g.ohow); g.d_kw = FastDiv(g.kw); g.d_khkw = FastDiv(g.khkw);
    g.d_w = FastDiv(g.W); g.d_hw = FastDiv(g.hw); g.d_sh = FastDiv(g.sh); g.d_sw = FastDiv(g.sw);
    return g;
}

static int validate(const cpc_conv_params* p) {
    if (!p) return CPC_ERR_NULL;
    if (p->batch <= 0 || p->c_in <= 0 || p->h_in <= 0 || p->w_in <= 0 || p->c_out <= 0 || p->h_out <= 0 ||
        p->w_out <= 0 || p->kh <= 0 || p->kw <= 0 || p->stride_h <= 0 || p->stride_w <= 0 || p->pad_top < 0 ||
        p->pad_left < 0)
        return CPC_ERR_BAD_SHAPE;
    // bottom / right zero padding is implied by (h_out, w_out) and may be arbitrarily large; sizes must fit
    // int32 indexing
    const int64_t lim = (1ll << 31) - 1;
    if ((int64_t)p->batch * p->h_out * p->w_out > lim || (int64_t)p->batch * p->h_in * p->w_in > lim ||
        (int64_t)p->c_in * p->kh * p->kw > lim || (int64_t)p->c_out * p->kh * p->kw > lim)
        return CPC_ERR_BAD_SHAPE;
    return CPC_OK;
}

// ---- loaders ------------------------------------------------------------------------------------
struct FwdA {   // rows: output pixels (b, oh, ow); k: (ci, i, j)
    static constexpr bool kFast = false;
    const float* x; ConvGeom g; int M, K;
    __device__ __forceinline__ float load(int m, int k) const {
        if (m >= M || k >= K) return 0.f;
        int b, pix, oh, ow, ci, r, i, j;
        g.d_ohow.divmod(m, b, pix);
        g.d_ow.divmod(pix, oh, ow);
        g.d_khkw.divmod(k, ci, r);
        g.d_kw.divmod(r, i, j);
        const int h = oh * g.sh - g.pt + i, w = ow * g.sw - g.pl + j;
        if ((unsigned)h >= (unsigned)g.H || (unsigned)w >= (unsigned)g.W) return 0.f;
        return __ldg(x + ((size_t)(b * g.Cin + ci) * g.H + h) * g.W + w);
    }
};
struct DenseRows {   // rows of a row-major (rows, K) matrix
    static constexpr bool kFast = true;
    const float* w; int N, K;
    __device__ __forceinline__ float load(int n, int k) const {
        if (n >= N || k >= K) return 0.f;
        return __ldg(w + (size_t)n * K + k);
    }
};
struct DgradA {   // rows: input pixels (b, ih, iw); k: (co, i, j)
    static constexpr bool kFast = false;
    const float* dy; ConvGeom g; int M, K;
    __device__ __forceinline__ float load(int m, int k) const {
        if (m >= M || k >= K) return 0.f;
        int b, pix, ih, iw, co, r, i, j;
        g.d_hw.divmod(m, b, pix);
        g.d_w.divmod(pix, ih, iw);
        g.d_khkw.divmod(k, co, r);
        g.d_kw.divmod(r, i, j);
        const int th = ih + g.pt - i, tw = iw + g.pl - j;
        if (th < 0 || tw < 0) return 0.f;
        const int oh = g.d_sh.div(th), ow = g.d_sw.div(tw);
        if (oh * g.sh != th || ow * g.sw != tw || oh >= g.OH || ow >= g.OW) return 0.f;
        return __ldg(dy + ((size_t)(b * g.Cout + co) * g.OH + oh) * g.OW + ow);
    }
};
struct DgradB {   // rows: ci; k: (co, i, j) -> w[co, ci, i, j]
    static constexpr bool kFast = true;
    const float* w; ConvGeom g; int N, K;
    __device__ __forceinline__ float load(int n, int k) const {
        if (n >= N || k >= K) return 0.f;
        int co, r;
        g.d_khkw.divmod(k, co, r);
        return __ldg(w + ((size_t)co * g.Cin + n) * g.khkw + r);
    }
};
struct WgradA {   // rows: co; k: output pixels (b, oh, ow)
    static constexpr bool kFast = true;
    const float* dy; ConvGeom g; int M, K;
    __device__ __forceinline__ float load(int m, int k) const {
        if (m >= M || k >= K) return 0.f;
        int b, pix;
        g.d_ohow.divmod(k, b, pix);
        return __ldg(dy + (size_t)(b * g.Cout + m) * g.ohow + pix);
    }
};
struct WgradB {   // rows: (ci, i, j); k: output pixels (b, oh, ow)
    static constexpr bool kFast = true;
    const float* x; ConvGeom g; int N, K;
    __device__ __forceinline__ float load(int n, int k) const {
        if (n >= N || k >= K) return 0.f;
        int b, pix, oh, ow, ci, r, i, j;
        g.d_ohow.divmod(k, b, pix);
        g.d_ow.divmod(pix, oh, ow);
        g.d_khkw.divmod(n, ci, r);
        g.d_kw.divmod(r, i, j);
        const int h = oh * g.sh - g.pt + i, w = ow * g.sw - g.pl + j;
        if ((unsigned)h >= (unsigned)g.H || (unsigned)w >= (unsigned)g.W) return 0.f;
        return __ldg(x + ((size_t)(b * g.Cin + ci) * g.H + h) * g.W + w);
    }
};

// ---- kernels ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TILE_THREADS) conv_fwd_kernel(FwdA la, DenseRows lb, const float* __restrict__ bias,
                                                               float* __restrict__ y, int relu) {
    __shared__ TileSmem sm;
    const int row0 = blockIdx.x * TILE, col0 = blockIdx.y * TILE;
    float acc[4][4] = {};
    tile_gemm(la, lb, row0, col0, 0, la.K, acc, sm);
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const ConvGeom& g = la.g;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = row0 + tx * 4 + i;
        if (m >= la.M) continue;
        int b, pix;
        g.d_ohow.divmod(m, b, pix);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = col0 + ty * 4 + j;
            if (n >= lb.N) continue;
            float v = acc[i][j] + (bias ? __ldg(bias + n) : 0.f);
            if (relu) v = fmaxf(v, 0.f);
            y[(size_t)(b * g.Cout + n) * g.ohow + pix] = v;
        }
    }
}

__global__ void __launch_bounds__(TILE_THREADS) conv_dgrad_kernel(DgradA la, DgradB lb, float* __restrict__ dx) {
    __shared__ TileSmem sm;
    const int row0 = blockIdx.x * TILE, col0 = blockIdx.y * TILE;
    float acc[4][4] = {};
    tile_gemm(la, lb, row0, col0, 0, la.K, acc, sm);
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const ConvGeom& g = la.g;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = row0 + tx * 4 + i;
        if (m >= la.M) continue;
        int b, pix;
        g.d_hw.divmod(m, b, pix);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = col0 + ty * 4 + j;
            if (n >= lb.N) continue;
            dx[(size_t)(b * g.Cin + n) * g.hw + pix] = acc[i][j];
        }
    }
}

__global__ void __launch_bounds__(TILE_THREADS) conv_wgrad_kernel(WgradA la, WgradB lb, float* __restrict__ dw,
                                                                 int k_per_split) {
    __shared__ TileSmem sm;
    const int row0 = blockIdx.x * TILE, col0 = blockIdx.y * TILE;
    const int k_begin = blockIdx.z * k_per_split;
    const int k_end = min(la.K, k_begin + k_per_split);
    float acc[4][4] = {};
    tile_gemm(la, lb, row0, col0, k_begin, k_end, acc, sm);
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = row0 + tx * 4 + i;
        if (m >= la.M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = col0 + ty * 4 + j;
            if (n >= lb.N) continue;
            atomicAdd(dw + (size_t)m * lb.N + n, acc[i][j]);
        }
    }
}

// dbias[co] = sum_{b, pix} dy[b, co, pix]; grid (segments of one plane, B*Cout), one float atomic per block
constexpr int DB_THREADS = 256;
constexpr int DB_PER_THREAD = 16;
__global__ void __launch_bounds__(DB_THREADS) conv_dbias_kernel(const float* __restrict__ dy, float* __restrict__ dbias,
                                                               int Cout, int ohow) {
    const int plane = blockIdx.x;
    const float* p = dy + (size_t)plane * ohow;
    const int i0 = blockIdx.y * (DB_THREADS * DB_PER_THREAD) + threadIdx.x;
    float s = 0.f;
#pragma unroll
    for (int u = 0; u < DB_PER_THREAD; ++u) {
        const int i = i0 + u * DB_THREADS;
        if (i < ohow) s += __ldg(p + i);
    }
    __shared__ float red[8];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        float v = threadIdx.x < 8 ? red[threadIdx.x] : 0.f;
        v = warp_sum(v);
        if (threadIdx.x == 0) atomicAdd(dbias + plane % Cout, v);
    }
}

}  // namespace cpc

using namespace cpc;

extern "C" size_t cpc_conv_workspace_bytes(const cpc_conv_params* p, int which) {
    if (validate(p) != CPC_OK) return 0;
    if ((which == 0 || which == 2) && smallk_path(p, which)) return 0;
    if (which >= 0 && which <= 2 && tall_path(p, which))
        return tall_path(p, which) == 1 ? tall_conv_workspace(p, which) : tall128_workspace(p, which);
    if ((which == 0 || which == 1) && tensor_core_path(p, which)) return umma_conv_workspace(p, which);
    if (which == 2 && tensor_core_path(p, 2)) return umma_wgrad_workspace(p);
    return 0;
}

// ---- caller-packed operands ---------------------------------------------------------------------------
// Which kernel family serves (p, which): 0 tiled CUDA-core, 1 direct small-K, 2 tall 32ch, 3 tall 128, 4 generic tcgen05
static int conv_family(const cpc_conv_params* p, int which) {
    if ((which == 0 || which == 2) && smallk_path(p, which)) return 1;
    if (which == 1 && !(p->flags & CPC_CONV_FLAG_NO_SMALLK) && small_dgrad_eligible(p)) return 1;
    if (const int tp = tall_path(p, which)) return tp == 1 ? 2 : 3;
    return tensor_core_path(p, which) ? 4 : 0;
}

extern "C" int cpc_conv_kernel_family(const cpc_conv_params* p, int which) {
    if (validate(p) != CPC_OK || which < 0 || which > 2) return -1;
    return conv_family(p, which);
}

// replicas of the canonical caller-packed dy: 2 when the generic strided data gradient reads shifted copies
static int dy_replicas(const cpc_conv_params* p) { return conv_family(p, 1) == 4 ? umma_dy_replicas(p) : 1; }

extern "C" size_t cpc_conv_packed_bytes(const cpc_conv_params* p, int operand) {
    if (validate(p) != CPC_OK) return 0;
    const size_t planes = p->precision == 1 ? 1 : 2;
    const size_t Wp = (size_t)((p->w_out + 7) & ~7);
    if (operand == 0) {
        if (conv_family(p, 0) < 2 && conv_family(p, 2) < 2) return 0;
        return align_up(planes * p->kw * p->batch * p->c_in * p->h_in * Wp * 2, 1024);
    }
    if (operand == 1) {
        if (conv_family(p, 2) < 2) return 0;
        const size_t nrep = dy_replicas(p);
        const size_t Wd = (size_t)((p->w_out + (nrep == 2 ? 1 : 0) + 7) & ~7);
        return align_up(planes * nrep * p->batch * p->c_out * p->h_out * Wd * 2, 1024);
    }
    return 0;
}

extern "C" int cpc_conv_pack(const float* src, void* packed, const cpc_conv_params* p, int operand, void* stream) {
    int st = validate(p);
    if (st != CPC_OK) return st;
    if (!src || !packed) return CPC_ERR_NULL;
    if (cpc_conv_packed_bytes(p, operand) == 0) return CPC_ERR_UNSUPPORTED;
    if ((reinterpret_cast<uintptr_t>(packed) & 15) != 0) return CPC_ERR_ALIGNMENT;   // TMA global base
    if ((st = check_device()) != CPC_OK) return st;
    const int planes = p->precision == 1 ? 1 : 2;
    const int Wp = (p->w_out + 7) & ~7;
    __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(packed);
    if (operand == 0)
        st = pack_split_launch(src, out, (long)p->batch * p->c_in * p->h_in, p->w_in, Wp, planes, p->kw, p->stride_w, 1,
                               -p->pad_left, (cudaStream_t)stream);
    else {
        const int nrep = dy_replicas(p);
        const int Wd = (p->w_out + (nrep == 2 ? 1 : 0) + 7) & ~7;
        st = pack_split_launch(src, out, (long)p->batch * p->c_out * p->h_out, p->w_out, Wd, planes, nrep, 1, -1, 0,
                               (cudaStream_t)stream);
    }
    if (st == CPC_OK) count_launch();
    return st;
}

extern "C" int cpc_conv_pack_dy(const float* dy, void* packed, float* dbias, const cpc_conv_params* p, void* stream) {
    if (!dbias) return cpc_conv_pack(dy, packed, p, 1, stream);
    int st = validate(p);
    if (st != CPC_OK) return st;
    if (!dy || !packed) return CPC_ERR_NULL;
    if (cpc_conv_packed_bytes(p, 1) == 0) return CPC_ERR_UNSUPPORTED;
    if ((reinterpret_cast<uintptr_t>(packed) & 15) != 0) return CPC_ERR_ALIGNMENT;
    if ((st = check_device()) != CPC_OK) return st;
    const int planes = p->precision == 1 ? 1 : 2;
    const int nrep = dy_replicas(p);
    const int Wp = (p->w_out + (nrep == 2 ? 1 : 0) + 7) & ~7;
    st = pack_split_sum_launch(dy, reinterpret_cast<__nv_bfloat16*>(packed), dbias, (long)p->batch * p->c_out * p->h_out,
                               p->w_out, Wp, planes, p->h_out, p->c_out, nrep, (cudaStream_t)stream);
    if (st == CPC_OK) count_launch();
    return st;
}

extern "C" int cpc_conv_fwd(const float* x, const float* w, const float* bias, float* y, const cpc_conv_params* p,
                            void* workspace, size_t workspace_bytes, void* stream) {
    return cpc_conv_fwd_ex(x, w, bias, y, p, nullptr, workspace, workspace_bytes, stream);
}
extern "C" int cpc_conv_dgrad(const float* dy, const float* w, float* dx, const cpc_conv_params* p, void* workspace,
                              size_t workspace_bytes, void* stream) {
    return cpc_conv_dgrad_ex(dy, w, dx, p, nullptr, workspace, workspace_bytes, stream);
}
extern "C" int cpc_conv_wgrad(const float* x, const float* dy, float* dw, float* dbias, const cpc_conv_params* p,
                              void* workspace, size_t workspace_bytes, void* stream) {
    return cpc_conv_wgrad_ex(x, dy, dw, dbias, p, nullptr, nullptr, workspace, workspace_bytes, stream);
}

extern "C" int cpc_conv_fwd_ex(const float* x, const float* w, const float* bias, float* y, const cpc_conv_params* p,
                               const void* packed_x, void* workspace, size_t workspace_bytes, void* stream) {
    int st = validate(p);
    if (st != CPC_OK) return st;
    {
        // the fp32 operand may be omitted when its packed form is given and the row-streaming kernels (which read only
        // the packed form) serve this configuration
        const int fam = conv_family(p, 0);
        const bool packed_only = !x && packed_x && (fam == 2 || fam == 3);
        if ((!x && !packed_only) || !w || !y) return CPC_ERR_NULL;
    }
    if ((st = check_device()) != CPC_OK) return st;
    cudaStream_t s = (cudaStream_t)stream;
    switch (conv_family(p, 0)) {
        case 1: return smallk_launch(0, x, w, bias, nullptr, y, p, s);
        case 2: return tall_conv_launch(x, w, bias, y, p, 0, packed_x, workspace, workspace_bytes, s);
        case 3: return tall128_conv_launch(x, w, bias, y, p, 0, packed_x, workspace, workspace_bytes, s);
        case 4: return umma_conv_launch(x, w, bias, y, p, 0, packed_x, workspace, workspace_bytes, s);
        default: break;
    }
    ConvGeom g = make_geom(p);
    const int M = g.B * g.ohow, N = g.Cout, K = g.Cin * g.khkw;
    FwdA la{x, g, M, K};
    DenseRows lb{w, N, K};
    dim3 grid(ceil_div(M, TILE), ceil_div(N, TILE));
    conv_fwd_kernel<<<grid, TILE_THREADS, 0, s>>>(la, lb, bias, y, p->relu);
    CPC_LAUNCH_CHECK();
    count_launch();
    return CPC_OK;
}

extern "C" int cpc_conv_dgrad_ex(const float* dy, const float* w, float* dx, const cpc_conv_params* p, const void* packed_dy,
                                 void* workspace, size_t workspace_bytes, void* stream) {
    int st = validate(p);
    if (st != CPC_OK) return st;
    {
        const int fam = conv_family(p, 1);
        const bool packed_only = !dy && packed_dy && (fam == 2 || fam == 3);
        if ((!dy && !packed_only) || !w || !dx) return CPC_ERR_NULL;
    }
    if ((st = check_device()) != CPC_OK) return st;
    cudaStream_t s = (cudaStream_t)stream;
    switch (conv_family(p, 1)) {
        case 1: return small_dgrad_launch(dy, w, dx, p, s);
        case 2: return tall_conv_launch(dy, w, nullptr, dx, p, 1, packed_dy, workspace, workspace_bytes, s);
        case 3: return tall128_conv_launch(dy, w, nullptr, dx, p, 1, packed_dy, workspace, workspace_bytes, s);
        case 4: return umma_conv_launch(dy, w, nullptr, dx, p, 1, packed_dy, workspace, workspace_bytes, s);
        default: break;
    }
    ConvGeom g = make_geom(p);
    const int M = g.B * g.hw, N = g.Cin, K = g.Cout * g.khkw;
    DgradA la{dy, g, M, K};
    DgradB lb{w, g, N, K};
    dim3 grid(ceil_div(M, TILE), ceil_div(N, TILE));
    conv_dgrad_kernel<<<grid, TILE_THREADS, 0, s>>>(la, lb, dx);
    CPC_LAUNCH_CHECK();
    count_launch();
    return CPC_OK;
}

static int launch_dbias(const float* dy, float* dbias, const ConvGeom& g, cudaStream_t s) {
    if (cudaMemsetAsync(dbias, 0, sizeof(float) * (size_t)g.Cout, s) != cudaSuccess) return CPC_ERR_CUDA;
    const dim3 grid(g.B * g.Cout, ceil_div(g.ohow, DB_THREADS * DB_PER_THREAD));
    conv_dbias_kernel<<<grid, DB_THREADS, 0, s>>>(dy, dbias, g.Cout, g.ohow);
    CPC_LAUNCH_CHECK();
    count_launch();
    return CPC_OK;
}

extern "C" int cpc_conv_wgrad_ex(const float* x, const float* dy, float* dw, float* dbias, const cpc_conv_params* p,
                                 const void* packed_x, const void* packed_dy, void* workspace, size_t workspace_bytes,
                                 void* stream) {
    int st = validate(p);
    if (st != CPC_OK) return st;
    const int fam = conv_family(p, 2);
    {
        const bool tall = fam == 2 || fam == 3;
        if ((!x && !(packed_x && tall)) || (!dy && !(packed_dy && tall && !dbias)) || !dw) return CPC_ERR_NULL;
    }
    if ((st = check_device()) != CPC_OK) return st;
    ConvGeom g = make_geom(p);
    cudaStream_t s = (cudaStream_t)stream;
    if (fam != 0) {
        if (fam == 1) st = smallk_launch(2, x, nullptr, nullptr, dy, dw, p, s);
        else if (fam == 2) st = tall_wgrad_launch(x, dy, dw, p, packed_x, packed_dy, workspace, workspace_bytes, s);
        else if (fam == 3) st = tall128_wgrad_launch(x, dy, dw, p, packed_x, packed_dy, workspace, workspace_bytes, s);
        else st = umma_wgrad_launch(x, dy, dw, p, packed_x, packed_dy, workspace, workspace_bytes, s);
        if (st != CPC_OK) return st;
        return dbias ? launch_dbias(dy, dbias, g, s) : CPC_OK;
    }
    const int M = g.Cout, N = g.Cin * g.khkw, K = g.B * g.ohow;
    if (cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)M * N, s) != cudaSuccess) return CPC_ERR_CUDA;
    const int tiles = ceil_div(M, TILE) * ceil_div(N, TILE);
    int splits = ceil_div(148 * 4, tiles);
    const int max_splits = ceil_div(K, TILE_K * 8);
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    int k_per_split = ceil_div(ceil_div(K, splits), TILE_K) * TILE_K;
    splits = ceil_div(K, k_per_split);
    WgradA la{dy, g, M, K};
    WgradB lb{x, g, N, K};
    dim3 grid(ceil_div(M, TILE), ceil_div(N, TILE), splits);
    conv_wgrad_kernel<<<grid, TILE_THREADS, 0, s>>>(la, lb, dw, k_per_split);
    CPC_LAUNCH_CHECK();
    count_launch();
    if (dbias) return launch_dbias(dy, dbias, g, s);
    return CPC_OK;
}
