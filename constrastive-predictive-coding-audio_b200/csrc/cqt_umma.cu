// Constant-Q filterbank on the tcgen05 tensor cores, with the log-power / phase-difference epilogue fused.
// Replaces CQT.forward (constant_q_transform.py:161-172) + PreprocessingModule.forward
// (scalogram_model.py:75-102) for every octave group whose kernel is at least 128 taps long.
//
//   out[b, f, t] = sum_n x[b, off_g + hop*t + n] * W_g[f, n]          (correlation, real | imag channels)
//
// GEMM view per group: M = 128 frames of one item, N = 2*n_g channels (padded to NP), K = K_g taps.
// The audio is viewed as a matrix X2[s][r] = x[hop*s + r]; frame t, taps [hop*m + 64*h, +64) is row t + m of
// X2, column half h.  So ALL the A operands of a 128-frame tile are row-shifted windows of one shared-memory
// slab of 256 X2 rows, which is loaded ONCE per tile (128 KB with both bf16 planes): the K loop only moves the
// descriptor start address by m rows (SWIZZLE_128B K-major with the descriptor's base-offset field), while
// the filter chunks (B operand) stream through a 4-stage TMA ring.  One tile accumulates every group of its
// "set" into separate TMEM columns, then the epilogue turns (re, im) into the scalogram in registers:
//   complex (B,F,T,2) | log-power (B,1,F,T) | log-power + unwrapped phase difference (B,2,F,T-1),
// with coalesced stores along t.  fp32-faithful arithmetic = bf16 hi/lo split, 3 MMA groups.
// Tiles (set, item, frame block) are handed out through an atomic counter in descending cost order.
#include "common.cuh"
#include "umma.cuh"

#include <cstdlib>

namespace cpc {
using namespace umma;

constexpr int CQ_THREADS = 384;                        // 4 control warps + 2 x 4 epilogue warps
constexpr int CQ_ACC_COLS = 256;                       // TMEM columns per accumulator buffer (two buffers)
constexpr int CQ_SLAB_ROWS = 256;
constexpr int CQ_SLAB_PLANE = CQ_SLAB_ROWS * 128;        // 32 KB: one column half, one plane
constexpr int CQ_BSTAGES = 5;
constexpr int CQ_MAX_SETS = 8;

struct CqGroup {
    int K, off, n_g, bin_lo, n_chunks, w_row0;           // w_row0: first row of the group in the packed filters
};

struct CqtUmma {
    int B, T, F, hop, halves, n_blocks, fpb;             // fpb: new frames per tile (127 in phase mode, else 128)
    int NP;                                              // padded channel count per group (multiple of 16, <= 64)
    int n_sets, set_first[CQ_MAX_SETS], set_count[CQ_MAX_SETS];
    int n_tiles;
    CqGroup g[CPC_CQT_MAX_GROUPS];
    int mode, To;
    float eps, log_offset, norm, power;
    const float* phase_fixed;
    const float* phase_scale;
    float* out;
    int* counter;
    int base_offset;
};

struct __align__(8) CqBarriers {
    uint64_t bfull[CQ_BSTAGES], bempty[CQ_BSTAGES], slab_full, slab_empty, acc_full[2], acc_empty[2], sfull[2], sempty[2];
    uint32_t tmem_base;
    int tile_id[2];
    float exch[2][2][4][32];                         // [epilogue warp set][parity][warp][bin]
};

// x (B, pitch) fp32 -> bf16 [plane][b][S*hop] (zeros past the item)
__global__ void __launch_bounds__(256) cqt_pack_audio_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out,
                                                            int B, int L, int pitch, int Lp) {
    const long groups = (long)B * (Lp >> 3);
    const long plane = (long)B * Lp;
    for (long gi = (long)blockIdx.x * blockDim.x + threadIdx.x; gi < groups; gi += (long)gridDim.x * blockDim.x) {
        const int b = (int)(gi / (Lp >> 3));
        const int i0 = (int)(gi - (long)b * (Lp >> 3)) << 3;
        __align__(16) __nv_bfloat16 hi[8];
        __align__(16) __nv_bfloat16 lo[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float v = (i0 + i < L) ? __ldg(x + (size_t)b * pitch + i0 + i) : 0.f;
            hi[i] = __float2bfloat16_rn(v);
            lo[i] = __float2bfloat16_rn(v - __bfloat162float(hi[i]));
        }
        const long o = (long)b * Lp + i0;
        *reinterpret_cast<uint4*>(out + o) = *reinterpret_cast<const uint4*>(hi);
        *reinterpret_cast<uint4*>(out + plane + o) = *reinterpret_cast<const uint4*>(lo);
    }
}

// filters: group block (2*n_g, K_g) fp32 row-major [real bins; imag bins] -> bf16 rows of 64 taps:
//   row = w_row0 + chunk * 2*NP + plane * NP + n
__global__ void __launch_bounds__(256) cqt_pack_filters_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out,
                                                              int K, int n2, int NP, int w_row0) {
    const int n_chunks = K >> 6;
    const long total = (long)n_chunks * NP * 64;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const int kk = (int)(idx & 63);
        const int n = (int)((idx >> 6) % NP);
        const int c = (int)((idx >> 6) / NP);
        // rows [0, n_g) real bins, rows [NP/2, NP/2 + n_g) imaginary bins, everything else zero
        const int ng = n2 >> 1, hp = NP >> 1;
        const int src = n < hp ? (n < ng ? n : -1) : (n - hp < ng ? ng + n - hp : -1);
        const float v = src >= 0 ? __ldg(w + (size_t)src * K + c * 64 + kk) : 0.f;
        const __nv_bfloat16 hi = __float2bfloat16_rn(v);
        const size_t row = (size_t)w_row0 + (size_t)c * 2 * NP + n;
        out[row * 64 + kk] = hi;
        out[(row + NP) * 64 + kk] = __float2bfloat16_rn(v - __bfloat162float(hi));
    }
}

__device__ __forceinline__ uint64_t cq_desc(uint32_t addr, int use_base_offset) {
    // K-major SWIZZLE_128B; the window may start on any 128-byte row of the slab
    uint64_t d = make_smem_desc(addr, 16, 1024);
    if (use_base_offset) d |= (uint64_t)((addr >> 7) & 7) << 49;
    return d;
}

__device__ __forceinline__ void cq_tile(const CqtUmma& p, int tile, int& set, int& b, int& blk) {
    const int per = p.B * p.n_blocks;
    set = tile / per;
    const int r = tile - set * per;
    b = r / p.n_blocks;
    blk = r - b * p.n_blocks;
}

__device__ __forceinline__ void epi_bar(int id) { asm volatile("bar.sync %0, 128;" ::"r"(id) : "memory"); }

// atan2 to ~2e-7 rad: octant reduction + the cephes atanf kernel (|z| <= tan(pi/8)), fast divisions.
__device__ __forceinline__ float fast_atan2(float y, float x) {
    const float ax = fabsf(x), ay = fabsf(y);
    const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
    float t = mx > 0.f ? __fdividef(mn, mx) : 0.f;               // in [0, 1]
    float base = 0.f;
    if (t > 0.41421356f) { base = 0.78539816f; t = __fdividef(t - 1.f, t + 1.f); }
    const float z = t * t;
    float r = fmaf(fmaf(fmaf(fmaf(8.05374449538e-2f, z, -1.38776856032e-1f), z, 1.99777106478e-1f), z, -3.33329491539e-1f) * z, t, t);
    r += base;
    if (ay > ax) r = 1.57079632679f - r;
    if (x < 0.f) r = 3.14159265359f - r;
    return y < 0.f ? -r : r;
}

__global__ void __launch_bounds__(CQ_THREADS, 1) cqt_umma_kernel(const __grid_constant__ CUtensorMap tmap_x,
                                                                const __grid_constant__ CUtensorMap tmap_w,
                                                                const CqtUmma p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* slab = smem;                                            // [half][plane][256 rows][128 B]
    const int slab_bytes = p.halves * 2 * CQ_SLAB_PLANE;
    uint8_t* b_ring = smem + slab_bytes;
    const int b_stage = 2 * p.NP * 128;
    CqBarriers* bars = reinterpret_cast<CqBarriers*>(b_ring + CQ_BSTAGES * b_stage);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmap_x);
        prefetch_tmap(&tmap_w);
        for (int s = 0; s < CQ_BSTAGES; ++s) { mbar_init(&bars->bfull[s], 1); mbar_init(&bars->bempty[s], 1); }
        mbar_init(&bars->slab_full, 1);
        mbar_init(&bars->slab_empty, 1);
        for (int s = 0; s < 2; ++s) {
            mbar_init(&bars->acc_full[s], 1);
            mbar_init(&bars->acc_empty[s], 8);
            mbar_init(&bars->sfull[s], 1);
            mbar_init(&bars->sempty[s], 9);
        }
        fence_barrier_init();
    }
    if (warp == 2) { tmem_alloc(&bars->tmem_base, 512); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;

    if (warp == 0) {
        // ===== scheduler + TMA producer =====
        if (lane == 0) {
            uint32_t bn = 0, tn = 0;
            for (uint32_t i = 0;; ++i) {
                const int slot = i & 1;
                mbar_wait(&bars->sempty[slot], ((i >> 1) & 1) ^ 1);
                int tile = atomicAdd(p.counter, 1);
                if (tile >= p.n_tiles) tile = -1;
                bars->tile_id[slot] = tile;
                mbar_arrive(&bars->sfull[slot]);
                if (tile < 0) break;
                int set, b, blk;
                cq_tile(p, tile, set, b, blk);
                mbar_wait(&bars->slab_empty, (tn & 1) ^ 1);
                mbar_expect_tx(&bars->slab_full, (uint32_t)slab_bytes);
                for (int h = 0; h < p.halves; ++h)
                    tma_load_4d(slab + h * 2 * CQ_SLAB_PLANE, &tmap_x, &bars->slab_full, h * 64, blk * p.fpb, b, 0);
                ++tn;
                for (int gi = 0; gi < p.set_count[set]; ++gi) {
                    const CqGroup& g = p.g[p.set_first[set] + gi];
                    for (int c = 0; c < g.n_chunks; ++c, ++bn) {
                        const int stage = bn % CQ_BSTAGES;
                        mbar_wait(&bars->bempty[stage], ((bn / CQ_BSTAGES) & 1) ^ 1);
                        mbar_expect_tx(&bars->bfull[stage], (uint32_t)b_stage);
                        tma_load_3d(b_ring + stage * b_stage, &tmap_w, &bars->bfull[stage], 0, g.w_row0 + c * 2 * p.NP, 0);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: warp-uniform loop, one elected lane issues.  An N = 64 MMA occupies the tensor pipe for
        // only 32 cycles, so the issue path is kept to one 64-bit add per descriptor =====
        {
            const uint32_t idesc = make_idesc_bf16(128, p.NP, 0, 0);
            const uint32_t slab_addr = smem_u32(slab);
            const uint32_t hop_shift = p.hop == 64 ? 6 : 7;                  // hop is 64 or 128
            uint32_t bn = 0, tn = 0;
            for (uint32_t i = 0;; ++i) {
                const int slot = i & 1;
                mbar_wait(&bars->sfull[slot], (i >> 1) & 1);
                const int tile = *reinterpret_cast<volatile int*>(&bars->tile_id[slot]);
                __syncwarp();
                if (lane == 0) mbar_arrive(&bars->sempty[slot]);
                if (tile < 0) break;
                int set, b, blk;
                cq_tile(p, tile, set, b, blk);
                const uint32_t buf = tn & 1;
                mbar_wait(&bars->acc_empty[buf], ((tn >> 1) & 1) ^ 1);
                mbar_wait(&bars->slab_full, tn & 1);
                tc_fence_after();
                for (int gi = 0; gi < p.set_count[set]; ++gi) {
                    const CqGroup& g = p.g[p.set_first[set] + gi];
                    const uint32_t d_tmem = tmem_base + buf * CQ_ACC_COLS + (uint32_t)(gi * p.NP);
                    for (int c = 0; c < g.n_chunks; ++c, ++bn) {
                        const int stage = bn % CQ_BSTAGES;
                        mbar_wait(&bars->bfull[stage], (bn / CQ_BSTAGES) & 1);
                        tc_fence_after();
                        const uint32_t so = (uint32_t)(g.off + c * 64);        // sample offset of this tap chunk
                        const uint32_t m = so >> hop_shift, half = (so & (p.hop - 1)) >> 6;
                        const uint32_t a_hi = slab_addr + half * 2 * CQ_SLAB_PLANE + m * 128;
                        const uint32_t b_hi = smem_u32(b_ring + stage * b_stage);
                        if (elect_one()) {
                            const uint64_t a_d_hi = cq_desc(a_hi, p.base_offset), a_d_lo = cq_desc(a_hi + CQ_SLAB_PLANE, p.base_offset);
                            const uint64_t b_d_hi = make_smem_desc(b_hi, 16, 1024), b_d_lo = make_smem_desc(b_hi + p.NP * 128, 16, 1024);
#pragma unroll
                            for (int cb = 0; cb < 3; ++cb) {                   // (hi,hi) (hi,lo) (lo,hi)
                                const uint64_t ad = cb == 2 ? a_d_lo : a_d_hi, bd = cb == 1 ? b_d_lo : b_d_hi;
#pragma unroll
                                for (int k = 0; k < 4; ++k)                    // +32 B per K step = +2 in the address field
                                    mma_bf16(d_tmem, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc, (uint32_t)(c | cb | k));
                            }
                            tc_commit(&bars->bempty[stage]);
                        }
                        __syncwarp();
                    }
                }
                if (elect_one()) {
                    tc_commit(&bars->slab_empty);
                    tc_commit(&bars->acc_full[buf]);
                }
                __syncwarp();
                ++tn;
            }
        }
    } else if (warp >= 4) {
        // ===== epilogue: (re, im) -> output format =====
        const int ew = warp & 3;                                               // TMEM lane quarter
        const int eset = (warp - 4) >> 2;                                      // warp set 0 / 1: even / odd groups
        const int r = ew * 32 + lane;
        const float kPi = 3.14159265358979323846f;
        uint32_t tn = 0, xb = 0;
        for (uint32_t i = 0;; ++i) {
            const int slot = i & 1;
            mbar_wait(&bars->sfull[slot], (i >> 1) & 1);
            const int tile = *reinterpret_cast<volatile int*>(&bars->tile_id[slot]);
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars->sempty[slot]);
            if (tile < 0) break;
            int set, b, blk;
            cq_tile(p, tile, set, b, blk);
            const int t = blk * p.fpb + r;                                     // frame of this thread
            const uint32_t buf = tn & 1;
            mbar_wait(&bars->acc_full[buf], (tn >> 1) & 1);
            tc_fence_after();
            const uint32_t lane_base = tmem_base + ((uint32_t)(ew * 32) << 16) + buf * CQ_ACC_COLS;
            for (int gi = eset; gi < p.set_count[set]; gi += 2) {
                const CqGroup& g = p.g[p.set_first[set] + gi];
                for (int j0 = 0; j0 < g.n_g; j0 += 32) {
                    uint32_t re[32], im[32];
                    tmem_ld32(lane_base + (uint32_t)(gi * p.NP + j0), re);
                    tmem_ld32(lane_base + (uint32_t)(gi * p.NP + (p.NP >> 1) + j0), im);
                    tmem_ld_wait();
                    const int nb = min(32, g.n_g - j0);
                    if (p.mode == CPC_CQT_COMPLEX) {
                        if (t < p.T) {
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                if (j < nb) {
                                    float2* o = reinterpret_cast<float2*>(p.out) + ((size_t)b * p.F + g.bin_lo + j0 + j) * p.T + t;
                                    *o = make_float2(__uint_as_float(re[j]), __uint_as_float(im[j]));
                                }
                        }
                    } else if (p.mode == CPC_CQT_LOGPOW) {
                        if (t < p.T) {
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                if (j < nb) {
                                    const float x = __uint_as_float(re[j]), y = __uint_as_float(im[j]);
                                    float amp = (__logf(fmaf(x, x, y * y) + p.eps) + p.log_offset) * p.norm;
                                    if (p.power != 1.f) amp = powf(amp, p.power);
                                    p.out[((size_t)b * p.F + g.bin_lo + j0 + j) * p.To + t] = amp;
                                }
                        }
                    } else {
                        // phase difference needs frame t - 1: previous lane, or lane 31 of the previous warp
                        float ph[32];
#pragma unroll
                        for (int j = 0; j < 32; ++j) ph[j] = fast_atan2(__uint_as_float(im[j]), __uint_as_float(re[j]));
                        float* ex = &bars->exch[eset][xb & 1][0][0];
                        if (lane == 31) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) ex[ew * 32 + j] = ph[j];
                        }
                        epi_bar(1 + eset);
                        ++xb;
                        const bool emit = r >= 1 && t < p.T;
                        const int to = t - 1;
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            float prev = __shfl_up_sync(0xffffffffu, ph[j], 1);
                            if (lane == 0 && ew > 0) prev = ex[(ew - 1) * 32 + j];
                            if (emit && j < nb) {
                                const int f = g.bin_lo + j0 + j;
                                const float x = __uint_as_float(re[j]), y = __uint_as_float(im[j]);
                                float amp = (__logf(fmaf(x, x, y * y) + p.eps) + p.log_offset) * p.norm;
                                float pd = ph[j] - prev + __ldg(p.phase_fixed + f);
                                if (pd > kPi) pd -= 2.f * kPi;
                                if (pd < -kPi) pd += 2.f * kPi;
                                float phv = pd * __ldg(p.phase_scale + f) * p.norm;
                                if (p.power != 1.f) { amp = powf(amp, p.power); phv = powf(phv, p.power); }
                                p.out[(((size_t)b * 2 + 0) * p.F + f) * p.To + to] = amp;
                                p.out[(((size_t)b * 2 + 1) * p.F + f) * p.To + to] = phv;
                            }
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars->acc_empty[buf]);
            ++tn;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// ---- host side --------------------------------------------------------------------------------------
struct CqtUmmaPlan {
    bool ok;
    int n_tensor_groups;          // groups [0, n) run here, the rest (K < 128) on the CUDA-core kernel
    int NP, halves, S, Lp;
    size_t audio_bytes, filter_bytes, total_rows;
};

static CqtUmmaPlan cqt_umma_plan(const cpc_cqt_params* p) {
    CqtUmmaPlan u{};
    u.ok = false;
    if (p->hop != 64 && p->hop != 128) return u;
    if (p->pool_t != 1 && p->mode != CPC_CQT_COMPLEX) return u;       // pooled modes keep the two-kernel path
    const int k0 = p->kernel_size[0];
    if (k0 < 128 || k0 / p->hop > 128) return u;                      // slab holds 128 frames + 127 row shifts
    int n = 0, np = 16;
    while (n < p->n_groups && p->kernel_size[n] >= 128) {
        const int n2 = 2 * (p->bin_hi[n] - p->bin_lo[n]);
        if (n2 > 64) return u;
        const int need = 2 * (((n2 >> 1) + 7) & ~7);                 // real | imag halves, each a multiple of 8 columns
        np = need > np ? need : np;
        ++n;
    }
    if (n == 0) return u;
    u.NP = (np + 15) & ~15;
    u.n_tensor_groups = n;
    u.halves = p->hop / 64;
    u.S = (p->n_samples + p->hop - 1) / p->hop;
    u.Lp = u.S * p->hop;
    u.audio_bytes = align_up((size_t)2 * p->batch * u.Lp * 2, 1024);
    size_t rows = 0;
    for (int g = 0; g < n; ++g) rows += (size_t)(p->kernel_size[g] / 64) * 2 * u.NP;
    u.total_rows = rows;
    u.filter_bytes = align_up(rows * 128, 1024);
    u.ok = true;
    return u;
}

bool cqt_umma_eligible(const cpc_cqt_params* p) { return cqt_umma_plan(p).ok; }
int cqt_umma_tensor_groups(const cpc_cqt_params* p) { return cqt_umma_plan(p).n_tensor_groups; }
size_t cqt_umma_workspace(const cpc_cqt_params* p) {
    CqtUmmaPlan u = cqt_umma_plan(p);
    return u.ok ? u.audio_bytes + u.filter_bytes + 256 + 1024 : 0;
}

// Computes groups [0, n_tensor_groups) straight into `out` in the requested mode.
int cqt_umma_launch(const float* x, const float* weights, const float* phase_fixed, const float* phase_scale, float* out,
                    const cpc_cqt_params* p, void* workspace, size_t workspace_bytes, cudaStream_t s) {
    CqtUmmaPlan u = cqt_umma_plan(p);
    if (!u.ok) return CPC_ERR_UNSUPPORTED;
    if (!workspace || workspace_bytes < cqt_umma_workspace(p)) return CPC_ERR_WORKSPACE;
    uint8_t* ws = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(workspace) + 1023) & ~(uintptr_t)1023);
    __nv_bfloat16* xp = reinterpret_cast<__nv_bfloat16*>(ws);
    __nv_bfloat16* wp = reinterpret_cast<__nv_bfloat16*>(ws + u.audio_bytes);
    int* counter = reinterpret_cast<int*>(ws + u.audio_bytes + u.filter_bytes);
    if (cudaMemsetAsync(counter, 0, sizeof(int), s) != cudaSuccess) return CPC_ERR_CUDA;
    {
        const long groups = (long)p->batch * (u.Lp / 8);
        int blocks = (int)((groups + 255) / 256);
        if (blocks > 148 * 8) blocks = 148 * 8;
        cqt_pack_audio_kernel<<<blocks, 256, 0, s>>>(x, xp, p->batch, p->n_samples, p->x_pitch, u.Lp);
        CPC_LAUNCH_CHECK();
    }
    CqtUmma k{};
    int row0 = 0;
    for (int g = 0; g < u.n_tensor_groups; ++g) {
        CqGroup& q = k.g[g];
        q.K = p->kernel_size[g];
        q.off = (p->kernel_size[0] - q.K) / 2;
        q.n_g = p->bin_hi[g] - p->bin_lo[g];
        q.bin_lo = p->bin_lo[g];
        q.n_chunks = q.K / 64;
        q.w_row0 = row0;
        const long total = (long)q.n_chunks * u.NP * 64;
        int blocks = (int)((total + 255) / 256);
        if (blocks > 148 * 4) blocks = 148 * 4;
        cqt_pack_filters_kernel<<<blocks, 256, 0, s>>>(weights + p->weight_offset[g], wp, q.K, 2 * q.n_g, u.NP, row0);
        CPC_LAUNCH_CHECK();
        row0 += q.n_chunks * 2 * u.NP;
    }
    // sets: consecutive groups, each at most 512 / NP accumulators, greedily balanced towards equal tap counts
    {
        long total = 0;
        for (int g = 0; g < u.n_tensor_groups; ++g) total += k.g[g].K;
        const int max_per_set = CQ_ACC_COLS / u.NP;
        const long target = (total + 2) / 3;
        int g = 0;
        k.n_sets = 0;
        while (g < u.n_tensor_groups) {
            int cnt = 0;
            long acc = 0;
            const bool last_slot = k.n_sets == CQ_MAX_SETS - 1;
            while (g + cnt < u.n_tensor_groups && cnt < max_per_set && (cnt == 0 || last_slot || acc + k.g[g + cnt].K <= target)) {
                acc += k.g[g + cnt].K;
                ++cnt;
            }
            if (last_slot && g + cnt < u.n_tensor_groups) return CPC_ERR_UNSUPPORTED;
            k.set_first[k.n_sets] = g;
            k.set_count[k.n_sets] = cnt;
            ++k.n_sets;
            g += cnt;
        }
    }
    CUtensorMap tx, tw;
    {
        const uint64_t dims[4] = {(uint64_t)p->hop, (uint64_t)u.S, (uint64_t)p->batch, 2};
        const uint64_t strides[3] = {(uint64_t)p->hop * 2, (uint64_t)u.Lp * 2, (uint64_t)u.Lp * 2 * p->batch};
        const uint32_t box[4] = {64, CQ_SLAB_ROWS, 1, 2};
        if (!make_tmap_bf16(&tx, xp, 4, dims, strides, box)) return CPC_ERR_CUDA;
        const uint64_t wd[3] = {64, (uint64_t)u.total_rows, 1};
        const uint64_t wsr[2] = {128, 128 * (uint64_t)u.total_rows};
        const uint32_t wbox[3] = {64, (uint32_t)(2 * u.NP), 1};
        if (!make_tmap_bf16(&tw, wp, 3, wd, wsr, wbox)) return CPC_ERR_CUDA;
    }
    k.B = p->batch; k.T = p->n_frames; k.F = p->n_bins; k.hop = p->hop; k.halves = u.halves; k.NP = u.NP;
    k.mode = p->mode;
    k.fpb = p->mode == CPC_CQT_LOGPOW_PHASE ? 127 : 128;
    const int t_out = p->mode == CPC_CQT_LOGPOW_PHASE ? p->n_frames - 1 : p->n_frames;
    k.To = t_out;
    k.n_blocks = ceil_div(t_out, k.fpb);
    k.n_tiles = k.n_sets * k.B * k.n_blocks;
    k.eps = p->eps; k.log_offset = p->log_offset; k.norm = p->norm; k.power = p->power;
    { const char* e = std::getenv("CPC_CQT_BASE_OFFSET"); k.base_offset = (e && e[0] == '1') ? 1 : 0; }
    k.phase_fixed = phase_fixed; k.phase_scale = phase_scale; k.out = out; k.counter = counter;
    const size_t smem_bytes = (size_t)u.halves * 2 * CQ_SLAB_PLANE + (size_t)CQ_BSTAGES * 2 * u.NP * 128 + sizeof(CqBarriers) + 1024;
    if (cudaFuncSetAttribute(cqt_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes) != cudaSuccess)
        return CPC_ERR_CUDA;
    const int grid = k.n_tiles < 148 ? k.n_tiles : 148;
    cqt_umma_kernel<<<grid, CQ_THREADS, smem_bytes, s>>>(tx, tw, k);
    CPC_LAUNCH_CHECK();
    count_launch(2 + u.n_tensor_groups);
    return CPC_OK;
}

}  // namespace cpc
