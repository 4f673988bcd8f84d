// Stride-1 convolution forward / data-gradient as an implicit GEMM on the 5th-gen tensor cores.
//
//   D[pixel, n] = sum_{tap, c} A[pixel + tap, c] * Wp[n, (tap, c)]
//
// * Activations stay in PyTorch's NCHW order.  Along w the pixels of one channel are contiguous, so an
//   activation tile is an "MN-major" UMMA operand: a TMA box (64 pixels x 1 row x KC channels) lands in
//   shared memory as KC rows of 128 B, exactly the SWIZZLE_128B MN-major canonical layout.  The conv tap
//   only shifts the box coordinates; zero padding (incl. the folded ZeroPad2d) is TMA out-of-bounds fill.
// * fp32-faithful mode: every fp32 operand is split into bf16 hi + lo planes (x = hi + lo + O(2^-17 x))
//   by a packing pass, and each K chunk issues three MMA groups hi*hi + hi*lo + lo*hi into the same fp32
//   TMEM accumulator (error ~1e-5 relative, vs 1e-3 for single bf16 and ~5e-4 for TF32).
//   bf16 mode loads and multiplies only the hi planes.
// * Persistent CTAs, warp-specialised: warp 0 TMA producer, warp 1 MMA issuer (one elected lane),
//   warp 2 TMEM allocator, warps 4-7 epilogue (TMEM -> registers -> bias/ReLU -> coalesced NCHW stores).
//   Two TMEM accumulator buffers let the epilogue of tile t overlap the MMAs of tile t+1.
// * The data gradient of a stride-1 conv is the same kernel run on dy with flipped taps and swapped
//   channel roles (weights re-packed accordingly).
#include "common.cuh"
#include "umma.cuh"

namespace cpc {
using namespace umma;

constexpr int UM_THREADS = 256;
constexpr int ATOM = 64;                 // pixels per MN-major swizzle atom (128 B of bf16)
constexpr int KCHUNK = 64;               // K elements per pipeline stage
constexpr int A_PLANE_BYTES = 2 * KCHUNK * 128;   // two 64-pixel atoms x 64 k-rows x 128 B = 16 KB
constexpr int SMEM_LIMIT = 227 * 1024;

struct UmmaConv {
    int n_atoms, OH, OW, AW;             // output geometry (atoms = 64-pixel runs along w)
    int n_rows_out;                      // output channels (GEMM N total)
    int n_tile, n_ntiles;
    int kh, kw, ntaps, tpc, cin_eff, cin_chunks, n_chunks;
    int pt, pl;
    int planes;                          // 2 = fp32-faithful split, 1 = bf16
    int relu, stages;
    const float* bias;
    float* y;
};

// ---- packing kernels --------------------------------------------------------------------------------
// fp32 (rows, W) -> bf16 (planes, nshift, rows, Wp); Wp % 8 == 0; one thread per 8 output columns.
// Replica s holds x[.., w + s - pad_left] (zero outside [0, W)): the horizontal tap offset is baked in here
// because a TMA box must start on a 16-byte boundary of the innermost dimension.
__global__ void __launch_bounds__(256) pack_split_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out,
                                                        long rows, int W, int Wp, int planes, int nshift, int pad_left) {
    const int groups = Wp >> 3;
    const long total = rows * groups;
    const long plane_stride = (long)nshift * rows * Wp;
    for (long g = (long)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (long)gridDim.x * blockDim.x) {
        const long row = g / groups;
        const int w0 = (int)(g - row * groups) << 3;
        const float* src = x + row * W;
        for (int s = 0; s < nshift; ++s) {
            __align__(16) __nv_bfloat16 hi[8];
            __align__(16) __nv_bfloat16 lo[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int w = w0 + i + s - pad_left;
                const float v = (w >= 0 && w < W) ? __ldg(src + w) : 0.f;
                hi[i] = __float2bfloat16_rn(v);
                lo[i] = __float2bfloat16_rn(v - __bfloat162float(hi[i]));
            }
            const long o = ((long)s * rows + row) * Wp + w0;
            *reinterpret_cast<uint4*>(out + o) = *reinterpret_cast<const uint4*>(hi);
            if (planes == 2) *reinterpret_cast<uint4*>(out + plane_stride + o) = *reinterpret_cast<const uint4*>(lo);
        }
    }
}

// weights (Cout, Cin, kh, kw) fp32 -> bf16 (planes, n_chunks, Nrows, 64), K-major rows of one K chunk.
// transpose_flip = 0: rows n = co, k = (tap, ci)              (forward)
//                = 1: rows n = ci, k = (flipped tap, co)      (data gradient)
__global__ void __launch_bounds__(256) pack_weights_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out,
                                                          int Cout, int Cin, int kh, int kw, int n_rows, int k_ch,
                                                          int cin_eff, int cin_chunks, int tpc, int n_chunks, int planes,
                                                          int transpose_flip) {
    const long total = (long)n_chunks * n_rows * KCHUNK;
    const int ntaps = kh * kw;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const int kk = (int)(idx % KCHUNK);
        const long t = idx / KCHUNK;
        const int n = (int)(t % n_rows);
        const int chunk = (int)(t / n_rows);
        int tap, c;
        if (cin_eff == KCHUNK) { tap = chunk / cin_chunks; c = (chunk - tap * cin_chunks) * KCHUNK + kk; }
        else                   { tap = chunk * tpc + kk / cin_eff; c = kk % cin_eff; }
        float v = 0.f;
        if (tap < ntaps && c < k_ch) {
            int i = tap / kw, j = tap - i * kw;
            if (transpose_flip) { i = kh - 1 - i; j = kw - 1 - j; v = __ldg(w + (((size_t)c * Cin + n) * kh + i) * kw + j); }
            else                { v = __ldg(w + (((size_t)n * Cin + c) * kh + i) * kw + j); }
        }
        const __nv_bfloat16 hi = __float2bfloat16_rn(v);
        out[idx] = hi;
        if (planes == 2) out[total + idx] = __float2bfloat16_rn(v - __bfloat162float(hi));
    }
}

// ---- the GEMM kernel ----------------------------------------------------------------------------------
struct __align__(8) UmmaBarriers {
    uint64_t full[8], empty[8], acc_full[2], acc_empty[2];
    uint32_t tmem_base;
};

__global__ void __launch_bounds__(UM_THREADS, 1) umma_conv_kernel(const __grid_constant__ CUtensorMap tmap_a,
                                                                 const __grid_constant__ CUtensorMap tmap_b,
                                                                 const UmmaConv p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int b_plane_bytes = p.n_tile * 128;
    const int stage_bytes = p.planes * (A_PLANE_BYTES + b_plane_bytes);
    UmmaBarriers* bars = reinterpret_cast<UmmaBarriers*>(smem + (size_t)p.stages * stage_bytes);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_mtiles = (p.n_atoms + 1) >> 1;
    const int n_tiles = n_mtiles * p.n_ntiles;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmap_a);
        prefetch_tmap(&tmap_b);
        for (int s = 0; s < p.stages; ++s) { mbar_init(&bars->full[s], 1); mbar_init(&bars->empty[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&bars->acc_full[b], 1); mbar_init(&bars->acc_empty[b], 4); }
        fence_barrier_init();
    }
    if (warp == 2) { tmem_alloc(&bars->tmem_base, 512); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                const int mtile = tile / p.n_ntiles, ntile = tile - mtile * p.n_ntiles;
                int ab[2], aoh[2], aw0[2];
                for (int a = 0; a < 2; ++a) {
                    const int atom = mtile * 2 + a;            // beyond n_atoms -> batch index OOB -> zero fill
                    const int row = atom / p.AW;
                    aw0[a] = (atom - row * p.AW) * ATOM;
                    ab[a] = row / p.OH;
                    aoh[a] = row - ab[a] * p.OH;
                }
                for (int q = 0; q < p.n_chunks; ++q) {
                    mbar_wait(&bars->empty[stage], phase ^ 1);
                    uint8_t* st = smem + (size_t)stage * stage_bytes;
                    mbar_expect_tx(&bars->full[stage], (uint32_t)stage_bytes);
                    for (int pl = 0; pl < p.planes; ++pl) {
                        uint8_t* a_dst = st + pl * A_PLANE_BYTES;
                        for (int a = 0; a < 2; ++a) {
                            if (p.cin_eff == KCHUNK) {
                                const int tap = q / p.cin_chunks, c0 = (q - tap * p.cin_chunks) * KCHUNK;
                                const int i = tap / p.kw, j = tap - i * p.kw;
                                tma_load_5d(a_dst + a * (KCHUNK * 128), &tmap_a, &bars->full[stage], aw0[a],
                                            aoh[a] + i - p.pt, c0, ab[a], pl * p.kw + j);
                            } else {
                                for (int t = 0; t < p.tpc; ++t) {
                                    int tap = q * p.tpc + t;
                                    if (tap >= p.ntaps) tap = 0;        // phantom tap: its packed weights are zero
                                    const int i = tap / p.kw, j = tap - i * p.kw;
                                    tma_load_5d(a_dst + a * (KCHUNK * 128) + t * p.cin_eff * 128, &tmap_a, &bars->full[stage],
                                                aw0[a], aoh[a] + i - p.pt, 0, ab[a], pl * p.kw + j);
                                }
                            }
                        }
                        tma_load_4d(st + p.planes * A_PLANE_BYTES + pl * b_plane_bytes, &tmap_b, &bars->full[stage], 0,
                                    ntile * p.n_tile, q, pl);
                    }
                    if (++stage == p.stages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            const uint32_t idesc = make_idesc_bf16(128, p.n_tile, /*A MN-major*/ 1, /*B K-major*/ 0);
            int stage = 0; uint32_t phase = 0;
            int buf = 0; uint32_t acc_phase = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                mbar_wait(&bars->acc_empty[buf], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)buf * 256;
                for (int q = 0; q < p.n_chunks; ++q) {
                    mbar_wait(&bars->full[stage], phase);
                    tc_fence_after();
                    const uint32_t st = smem_u32(smem + (size_t)stage * stage_bytes);
                    const uint32_t b0 = st + p.planes * A_PLANE_BYTES;
                    const int ncombo = p.planes == 2 ? 3 : 1;
                    for (int cb = 0; cb < ncombo; ++cb) {
                        const uint32_t a_addr = st + (cb == 2 ? A_PLANE_BYTES : 0);          // (hi,hi) (hi,lo) (lo,hi)
                        const uint32_t b_addr = b0 + (cb == 1 ? b_plane_bytes : 0);
#pragma unroll
                        for (int k = 0; k < KCHUNK / 16; ++k) {
                            const uint64_t ad = make_smem_desc(a_addr + k * (16 * 128), KCHUNK * 128, 1024);
                            const uint64_t bd = make_smem_desc(b_addr + k * 32, 16, 1024);
                            mma_bf16(d_tmem, ad, bd, idesc, (q | cb | k) != 0);
                        }
                    }
                    tc_commit(&bars->empty[stage]);           // frees the smem stage once these MMAs retire
                    if (++stage == p.stages) { stage = 0; phase ^= 1; }
                }
                tc_commit(&bars->acc_full[buf]);               // accumulator complete -> epilogue
                if (++buf == 2) { buf = 0; acc_phase ^= 1; }
            }
        }
    } else if (warp >= 4) {
        // ===== epilogue: TMEM -> registers -> bias / ReLU -> NCHW fp32 =====
        const int ew = warp & 3;                               // TMEM lane quarter this warp may access
        const int r = ew * 32 + lane;                          // tile row = pixel
        int buf = 0; uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const int mtile = tile / p.n_ntiles, ntile = tile - mtile * p.n_ntiles;
            const int atom = mtile * 2 + (r >> 6);
            const int row = atom / p.AW;
            const int ow = (atom - row * p.AW) * ATOM + (r & 63);
            const int b = row / p.OH, oh = row - b * p.OH;
            const bool valid = atom < p.n_atoms && ow < p.OW;
            mbar_wait(&bars->acc_full[buf], acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)buf * 256;
            float* ybase = p.y + ((size_t)b * p.n_rows_out * p.OH + oh) * p.OW + ow;
            const size_t chan_stride = (size_t)p.OH * p.OW;
            for (int c0 = 0; c0 < p.n_tile; c0 += 32) {
                uint32_t v[32];
                tmem_ld32(taddr + c0, v);
                tmem_ld_wait();
                if (valid) {
#pragma unroll
                    for (int c = 0; c < 32; ++c) {
                        const int n = ntile * p.n_tile + c0 + c;
                        float f = __uint_as_float(v[c]);
                        if (p.bias) f += __ldg(p.bias + n);
                        if (p.relu) f = fmaxf(f, 0.f);
                        ybase[(size_t)n * chan_stride] = f;
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars->acc_empty[buf]);
            if (++buf == 2) { buf = 0; acc_phase ^= 1; }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// ---- host side ----------------------------------------------------------------------------------------
struct UmmaPlan {
    bool ok;
    int in_ch, out_ch, H, W, Wp, OH, OW;      // GEMM view: input (B,in_ch,H,W) -> output (B,out_ch,OH,OW)
    int kh, kw, pt, pl;
    int cin_eff, cin_chunks, tpc, n_chunks, n_tile, n_ntiles, planes, stages;
    size_t act_bytes, w_bytes, smem_bytes;
};

static UmmaPlan make_plan(const cpc_conv_params* p, int which) {
    UmmaPlan u{};
    u.ok = false;
    if (p->stride_h != 1 || p->stride_w != 1) return u;
    if (which == 0) {
        u.in_ch = p->c_in; u.out_ch = p->c_out; u.H = p->h_in; u.W = p->w_in; u.OH = p->h_out; u.OW = p->w_out;
        u.pt = p->pad_top; u.pl = p->pad_left;
    } else {
        u.in_ch = p->c_out; u.out_ch = p->c_in; u.H = p->h_out; u.W = p->w_out; u.OH = p->h_in; u.OW = p->w_in;
        u.pt = p->kh - 1 - p->pad_top; u.pl = p->kw - 1 - p->pad_left;
        if (u.pt < 0 || u.pl < 0) return u;
    }
    u.kh = p->kh; u.kw = p->kw;
    const int ci = u.in_ch, co = u.out_ch;
    if (!(ci == 16 || ci == 32 || ci % 64 == 0)) return u;
    if (co % 32 != 0 || !(co <= 128 || co % 128 == 0)) return u;
    if ((long)p->batch * u.OH * ((u.OW + ATOM - 1) / ATOM) > (1l << 30)) return u;
    u.cin_eff = ci >= KCHUNK ? KCHUNK : ci;
    u.cin_chunks = ci >= KCHUNK ? ci / KCHUNK : 1;
    u.tpc = KCHUNK / u.cin_eff;
    const int ntaps = u.kh * u.kw;
    u.n_chunks = ci >= KCHUNK ? ntaps * u.cin_chunks : (ntaps + u.tpc - 1) / u.tpc;
    u.n_tile = co <= 128 ? co : 128;
    u.n_ntiles = co / u.n_tile;
    u.planes = p->precision == 1 ? 1 : 2;
    u.Wp = (u.W + 7) & ~7;
    const int stage_bytes = u.planes * (A_PLANE_BYTES + u.n_tile * 128);
    u.stages = (SMEM_LIMIT - 2048) / stage_bytes;
    if (u.stages > 8) u.stages = 8;
    if (u.stages < 2) return u;
    u.smem_bytes = (size_t)u.stages * stage_bytes + sizeof(UmmaBarriers) + 1024;
    u.act_bytes = align_up((size_t)u.planes * u.kw * p->batch * ci * u.H * u.Wp * 2, 1024);
    u.w_bytes = align_up((size_t)u.planes * u.n_chunks * co * KCHUNK * 2, 1024);
    u.ok = true;
    return u;
}

size_t umma_conv_workspace(const cpc_conv_params* p, int which) {
    UmmaPlan u = make_plan(p, which);
    return u.ok ? u.act_bytes + u.w_bytes + 1024 : 0;
}

bool umma_conv_eligible(const cpc_conv_params* p, int which) { return make_plan(p, which).ok; }

// which = 0: y = conv(x, w) + bias ; which = 1: dx = conv_transpose(dy, w).   `in` is x or dy, `out` is y or dx.
int umma_conv_launch(const float* in, const float* w, const float* bias, float* out, const cpc_conv_params* p, int which,
                     void* workspace, size_t workspace_bytes, cudaStream_t s) {
    UmmaPlan u = make_plan(p, which);
    if (!u.ok) return CPC_ERR_UNSUPPORTED;
    if (!workspace || workspace_bytes < u.act_bytes + u.w_bytes + 1024) return CPC_ERR_WORKSPACE;
    uint8_t* ws = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(workspace) + 1023) & ~(uintptr_t)1023);
    __nv_bfloat16* act = reinterpret_cast<__nv_bfloat16*>(ws);
    __nv_bfloat16* wp = reinterpret_cast<__nv_bfloat16*>(ws + u.act_bytes);
    const int B = p->batch;
    const long rows = (long)B * u.in_ch * u.H;
    {
        const long groups = rows * (u.Wp / 8);
        int blocks = (int)((groups + 255) / 256);
        if (blocks > 148 * 16) blocks = 148 * 16;
        pack_split_kernel<<<blocks, 256, 0, s>>>(in, act, rows, u.W, u.Wp, u.planes, u.kw, u.pl);
        CPC_LAUNCH_CHECK();
        const long wtotal = (long)u.n_chunks * u.out_ch * KCHUNK;
        int wblocks = (int)((wtotal + 255) / 256);
        if (wblocks > 148 * 8) wblocks = 148 * 8;
        pack_weights_kernel<<<wblocks, 256, 0, s>>>(w, wp, p->c_out, p->c_in, p->kh, p->kw, u.out_ch, u.in_ch, u.cin_eff,
                                                    u.cin_chunks, u.tpc, u.n_chunks, u.planes, which);
        CPC_LAUNCH_CHECK();
    }
    CUtensorMap ta, tb;
    {
        // replica / plane index is the outermost dimension; OW + pad columns of a replica may be addressed
        const uint64_t dims[5] = {(uint64_t)u.Wp, (uint64_t)u.H, (uint64_t)u.in_ch, (uint64_t)B,
                                  (uint64_t)u.planes * u.kw};
        const uint64_t row_b = (uint64_t)u.Wp * 2;
        const uint64_t strides[4] = {row_b, row_b * u.H, row_b * u.H * u.in_ch, row_b * u.H * u.in_ch * B};
        const uint32_t box[5] = {ATOM, 1, (uint32_t)u.cin_eff, 1, 1};
        if (!make_tmap_bf16(&ta, act, 5, dims, strides, box)) return CPC_ERR_CUDA;
        const uint64_t wd[4] = {KCHUNK, (uint64_t)u.out_ch, (uint64_t)u.n_chunks, (uint64_t)u.planes};
        const uint64_t wsr[3] = {KCHUNK * 2, (uint64_t)KCHUNK * 2 * u.out_ch, (uint64_t)KCHUNK * 2 * u.out_ch * u.n_chunks};
        const uint32_t wbox[4] = {KCHUNK, (uint32_t)u.n_tile, 1, 1};
        if (!make_tmap_bf16(&tb, wp, 4, wd, wsr, wbox)) return CPC_ERR_CUDA;
    }
    UmmaConv k{};
    k.AW = (u.OW + ATOM - 1) / ATOM;
    k.OH = u.OH; k.OW = u.OW;
    k.n_atoms = B * u.OH * k.AW;
    k.n_rows_out = u.out_ch; k.n_tile = u.n_tile; k.n_ntiles = u.n_ntiles;
    k.kh = u.kh; k.kw = u.kw; k.ntaps = u.kh * u.kw; k.tpc = u.tpc; k.cin_eff = u.cin_eff; k.cin_chunks = u.cin_chunks;
    k.n_chunks = u.n_chunks; k.pt = u.pt; k.pl = u.pl; k.planes = u.planes;
    k.relu = which == 0 ? p->relu : 0;
    k.stages = u.stages;
    k.bias = which == 0 ? bias : nullptr;
    k.y = out;
    if (cudaFuncSetAttribute(umma_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT) != cudaSuccess)
        return CPC_ERR_CUDA;
    const int n_tiles = ((k.n_atoms + 1) / 2) * k.n_ntiles;
    int sms = 148;
    int grid = n_tiles < sms ? n_tiles : sms;
    umma_conv_kernel<<<grid, UM_THREADS, u.smem_bytes, s>>>(ta, tb, k);
    CPC_LAUNCH_CHECK();
    count_launch(3);
    return CPC_OK;
}

}  // namespace cpc
