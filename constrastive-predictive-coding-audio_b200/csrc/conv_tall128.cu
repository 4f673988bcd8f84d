// Row-streaming tcgen05 kernels for tall pitch convolutions with 128 output channels
// (kh x 1, stride 1; scalogram_model.py:387-431 -- the 30x1, 128 -> 128 conv of scalogram_resnet_architecture_7
// block 1), forward, data gradient and weight gradient.  Same idea as conv_tall.cu (an input row is fetched
// from L2 once per tile and reused by every output row it contributes to), different shape regime:
//
// Forward / data gradient (tall128_conv_kernel):
//   out[b, n, r, w] = sum_{c, i} src[b, c, r - P + i, w] * Wt[n, c, i],   n < 128,  C = 64 * NQ channels
//   * tile = 128 pixels x R = 4 output rows, one 128-column fp32 accumulator per row (512 TMEM columns).
//   * channel-chunk-major: for each chunk q of 64 channels, step j streams input row j (MN-major A operand,
//     32 KB with both bf16 planes) and weight tap (j, q) (K-major B operand, 32 KB) and issues, for every
//     output row rho whose tap j - rho is inside the kernel, D[rho] += A_j * B_{j - rho}.  The last R taps
//     stay resident in a 5-slot ring, so activations and weights are each read once per tile and chunk.
//   * tiles come from an atomic counter (conv_tall.cu explains why); the weight-packing kernel zeroes it.
//   * bf16 operand mode (cpc_conv_params.precision = 1): hi planes only (half the bytes per step), one product.
// Weight gradient (tall128_wgrad_kernel):
//   dW[co, ci, i] = sum_{b, r, w} dy[b, co, r, w] * x[b, ci, r - P + i, w]
//   * K = 64-pixel chunks, both operands K-major: A = one x row (128 ci x 64 px), B = one dy row.
//   * a CTA owns NT = 4 consecutive taps (4 x 128 TMEM columns) and a share of the pixel atoms; per dy row r
//     it streams dy row r and x row r - P + i0 + NT - 1; the NT newest x rows stay resident.
#include "common.cuh"
#include "umma.cuh"

namespace cpc {
using namespace umma;

int pack_split_launch(const float* x, __nv_bfloat16* out, long rows, int W, int Wp, int planes, int nrep, int w_mul,
                      int rep_mul, int w_off, cudaStream_t s);   // conv_umma.cu

constexpr int T8_N = 128;
constexpr int T8_R = 4;                      // output rows per tile
constexpr int T8_S = 2;                      // activation stages
constexpr int T8_T = 5;                      // tap ring (>= R - 1 + S)
constexpr int T8_TILE = 2 * 128 * 128;       // 32 KB: both planes of a 128 x 64 operand tile
constexpr int T8_THREADS = 256;
constexpr int T8_SMEM = (T8_S + T8_T) * T8_TILE + 1024 + 256;

struct Tall128Conv {
    int B, H_src, H_out, W, AW, kh, P, NQ;
    int n_units, n_pairs, n_rtiles, n_tiles;
    int relu;
    int planes;                              // 2: bf16 hi/lo (fp32-faithful), 1: hi plane only (bf16 operand mode)
    const float* bias;
    float* out;                              // (B, 128, H_out, W)
    int* counter;                            // tile scheduler (workspace; zero at launch)
};

struct __align__(8) Tall128Barriers {
    uint64_t full[T8_S], empty[T8_S], acc_full, acc_empty;
    uint64_t sfull[2], sempty[2];            // tile-id queue (umma.cuh: sched_push / sched_pop)
    int tile_id[2];
    uint32_t tmem_base;
};

// weights (Cout, Cin, kh, 1) fp32 -> bf16 [tap][q][plane][n (128)][64 c]
//   flip_swap = 0: n = co, c = ci, tap = i;   1: n = ci, c = co, tap = kh - 1 - i
__global__ void __launch_bounds__(256) tall128_pack_weights_kernel(const float* __restrict__ w,
                                                                  __nv_bfloat16* __restrict__ out, int Cin, int kh, int C,
                                                                  int flip_swap, int* __restrict__ counter) {
    if (blockIdx.x == 0 && threadIdx.x == 0) *counter = 0;        // tile scheduler of the conv kernel that follows
    const int NQ = C >> 6;
    const long total = (long)kh * NQ * T8_N * 64;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const int kk = (int)(idx & 63);
        const int n = (int)((idx >> 6) & (T8_N - 1));
        const int tq = (int)(idx >> 13);
        const int q = tq % NQ, tap = tq / NQ;
        const int c = q * 64 + kk;
        const int i = flip_swap ? kh - 1 - tap : tap;
        const int co = flip_swap ? c : n, ci = flip_swap ? n : c;
        const float v = __ldg(w + ((size_t)co * Cin + ci) * kh + i);
        const __nv_bfloat16 hi = __float2bfloat16_rn(v);
        const size_t o = ((size_t)tq * 2 * T8_N + n) * 64 + kk;
        out[o] = hi;
        out[o + (size_t)T8_N * 64] = __float2bfloat16_rn(v - __bfloat162float(hi));
    }
}

__device__ __forceinline__ void t8_tile_rows(const Tall128Conv& p, int tile, int& pair, int& r0, int& j_lo, int& j_hi) {
    pair = tile / p.n_rtiles;
    r0 = (tile - pair * p.n_rtiles) * T8_R;
    j_lo = max(0, p.P - r0);
    j_hi = min(T8_R + p.kh - 1, p.H_src + p.P - r0);
    if (j_hi < j_lo) j_hi = j_lo;
}

__global__ void __launch_bounds__(T8_THREADS, 1) tall128_conv_kernel(const __grid_constant__ CUtensorMap tmap_a,
                                                                    const __grid_constant__ CUtensorMap tmap_w,
                                                                    const Tall128Conv p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* a_ring = smem;                                  // stage: [atom][plane][64 ch][128 B]
    uint8_t* w_ring = smem + T8_S * T8_TILE;                 // slot:  [plane][128 n][128 B]
    Tall128Barriers* bars = reinterpret_cast<Tall128Barriers*>(w_ring + T8_T * T8_TILE);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmap_a);
        prefetch_tmap(&tmap_w);
        for (int s = 0; s < T8_S; ++s) { mbar_init(&bars->full[s], 1); mbar_init(&bars->empty[s], 1); }
        mbar_init(&bars->acc_full, 1);
        mbar_init(&bars->acc_empty, 4);
        for (int s = 0; s < 2; ++s) { mbar_init(&bars->sfull[s], 1); mbar_init(&bars->sempty[s], 5); }
        fence_barrier_init();
    }
    if (warp == 2) { tmem_alloc(&bars->tmem_base, 512); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;

    if (warp == 0) {
        if (lane == 0) {
            uint32_t v = 0;
            const uint32_t op_bytes = p.planes == 2 ? T8_TILE : T8_TILE / 2;   // one operand tile: both planes or hi only
            for (uint32_t ti = 0;; ++ti) {
                const int tile = sched_push(bars, ti, p.counter, p.n_tiles);
                if (tile < 0) break;
                int pair, r0, j_lo, j_hi;
                t8_tile_rows(p, tile, pair, r0, j_lo, j_hi);
                if (j_hi == j_lo) continue;
                int ab[2], aw0[2];
                for (int a = 0; a < 2; ++a) {
                    const int u = pair * 2 + a;
                    ab[a] = u / p.AW;
                    aw0[a] = (u - ab[a] * p.AW) * 64;
                }
                for (int q = 0; q < p.NQ; ++q)
                    for (int j = j_lo - (T8_R - 1); j < j_hi; ++j, ++v) {
                        const int stage = v % T8_S;
                        mbar_wait(&bars->empty[stage], ((v / T8_S) & 1) ^ 1);
                        const bool with_row = j >= j_lo;
                        mbar_expect_tx(&bars->full[stage], op_bytes * (with_row ? 2 : 1));
                        // tap j of chunk q; j outside [0, kh) addresses outside the tensor -> zero fill
                        const int tq = (j < 0 || j >= p.kh) ? -1 : j * p.NQ + q;
                        tma_load_3d(w_ring + (v % T8_T) * T8_TILE, &tmap_w, &bars->full[stage], 0, 0, tq);
                        if (with_row) {
                            uint8_t* st = a_ring + stage * T8_TILE;
                            for (int a = 0; a < 2; ++a)
                                tma_load_5d(st + a * (T8_TILE / 2), &tmap_a, &bars->full[stage], aw0[a], r0 - p.P + j, q * 64,
                                            ab[a], 0);
                        }
                    }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = make_idesc_bf16(128, T8_N, /*A MN-major*/ 1, /*B K-major*/ 0);
            const uint32_t w_base = smem_u32(w_ring);
            const int n_cb = p.planes == 2 ? 3 : 1;
            uint32_t v = 0, acc_phase = 0;
            for (uint32_t ti = 0;; ++ti) {
                const int tile = sched_pop_thread(bars, ti);          // this role is a single thread
                if (tile < 0) break;
                int pair, r0, j_lo, j_hi;
                t8_tile_rows(p, tile, pair, r0, j_lo, j_hi);
                if (j_hi == j_lo) continue;
                mbar_wait(&bars->acc_empty, acc_phase ^ 1);
                tc_fence_after();
                uint32_t started = 0;
                for (int q = 0; q < p.NQ; ++q)
                    for (int j = j_lo - (T8_R - 1); j < j_hi; ++j, ++v) {
                        const int stage = v % T8_S;
                        mbar_wait(&bars->full[stage], (v / T8_S) & 1);
                        tc_fence_after();
                        if (j >= j_lo) {
                            // one descriptor per operand tile; planes and K steps are constant increments of its 16-byte
                            // address field (the issuing thread's instruction count per MMA bounds this kernel)
                            const uint64_t a_d0 = make_smem_desc(smem_u32(a_ring + stage * T8_TILE), T8_TILE / 2, 1024);
                            for (int rho = 0; rho < T8_R; ++rho) {
                                const int tap = j - rho;
                                if (tap < 0 || tap >= p.kh) continue;
                                const uint64_t b_d0 = make_smem_desc(w_base + ((v - rho) % T8_T) * T8_TILE, 16, 1024);
                                const uint32_t d_tmem = tmem_base + (uint32_t)rho * T8_N;
#pragma unroll
                                for (int cb = 0; cb < 3; ++cb) {               // (hi,hi) (hi,lo) (lo,hi)
                                    if (cb >= n_cb) break;
                                    const uint64_t a_d = a_d0 + (uint64_t)(cb == 2 ? (64 * 128) >> 4 : 0);
                                    const uint64_t b_d = b_d0 + (uint64_t)(cb == 1 ? (T8_N * 128) >> 4 : 0);
#pragma unroll
                                    for (int k = 0; k < 4; ++k)
                                        mma_bf16(d_tmem, a_d + (uint64_t)(k * ((16 * 128) >> 4)), b_d + (uint64_t)(k * 2), idesc,
                                                 ((started >> rho) & 1u) | (uint32_t)(cb | k));
                                }
                                started |= 1u << rho;
                            }
                        }
                        tc_commit(&bars->empty[stage]);
                    }
                tc_commit(&bars->acc_full);
                acc_phase ^= 1;
            }
        }
    } else if (warp >= 4) {
        const int ew = warp & 3;
        const int r = ew * 32 + lane;
        uint32_t acc_phase = 0;
        for (uint32_t ti = 0;; ++ti) {
            const int tile = sched_pop(bars, ti, lane);
            if (tile < 0) break;
            int pair, r0, j_lo, j_hi;
            t8_tile_rows(p, tile, pair, r0, j_lo, j_hi);
            const int u = pair * 2 + (r >> 6);
            const int b = u / p.AW;
            const int pw = (u - b * p.AW) * 64 + (r & 63);
            const bool valid = u < p.n_units && pw < p.W;
            const bool empty_tile = j_hi == j_lo;
            if (!empty_tile) {
                mbar_wait(&bars->acc_full, acc_phase);
                tc_fence_after();
            }
            const size_t chan_stride = (size_t)p.H_out * p.W;
            for (int rho = 0; rho < T8_R; ++rho) {
                const int row = r0 + rho;
                if (row >= p.H_out) break;
                // the accumulator of this row was written iff some input row inside the source meets a tap
                const bool written = !empty_tile && max(j_lo, rho) < min(j_hi, rho + p.kh);
                float* o = p.out + ((size_t)b * T8_N * p.H_out + row) * p.W + pw;
                for (int c0 = 0; c0 < T8_N; c0 += 32) {
                    uint32_t v[32];
                    if (written) {
                        tmem_ld32(tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(rho * T8_N + c0), v);
                        tmem_ld_wait();
                    } else {
#pragma unroll
                        for (int c = 0; c < 32; ++c) v[c] = 0u;
                    }
                    if (valid) {
#pragma unroll
                        for (int c = 0; c < 32; ++c) {
                            float f = __uint_as_float(v[c]);
                            if (p.bias) f += __ldg(p.bias + c0 + c);
                            if (p.relu) f = fmaxf(f, 0.f);
                            o[(size_t)(c0 + c) * chan_stride] = f;
                        }
                    }
                }
            }
            if (!empty_tile) {
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bars->acc_empty);
                acc_phase ^= 1;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// ---- weight gradient --------------------------------------------------------------------------------------
constexpr int T8W_NT = 4;                    // taps per CTA
constexpr int T8W_XS = 5;                    // x row ring (NT resident + 1 in flight)
constexpr int T8W_YS = 2;
constexpr int T8W_SMEM = (T8W_XS + T8W_YS) * T8_TILE + 1024 + 256;

struct Tall128Wgrad {
    int B, H_src, H_out, W, AW, kh, P;
    int n_units, n_splits, n_tgroups;
    int planes;
    float* dw;                               // (128, 128, kh) fp32, zero-initialised
};

struct __align__(8) Tall128WgradBarriers {
    uint64_t xfull[T8W_XS], xempty[T8W_XS], yfull[T8W_YS], yempty[T8W_YS], acc_full;
    uint32_t tmem_base;
};

__global__ void __launch_bounds__(T8_THREADS, 1) tall128_wgrad_kernel(const __grid_constant__ CUtensorMap tmap_x,
                                                                     const __grid_constant__ CUtensorMap tmap_dy,
                                                                     const Tall128Wgrad p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* x_ring = smem;                                  // slot: [plane][128 ci][128 B]
    uint8_t* y_ring = smem + T8W_XS * T8_TILE;
    Tall128WgradBarriers* bars = reinterpret_cast<Tall128WgradBarriers*>(y_ring + T8W_YS * T8_TILE);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tg = blockIdx.x % p.n_tgroups, split = blockIdx.x / p.n_tgroups;
    const int i0 = tg * T8W_NT;                              // taps i0 .. i0 + NT - 1
    // dy row r pairs with x rows h = r - P + i0 + k, k < NT.  x row sequence index n <-> h = n - P + i0:
    // dy row r uses n = r .. r + NT - 1; per unit n runs over [0, H_out + NT - 1).
    const int n_rows = p.H_out + T8W_NT - 1;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmap_x);
        prefetch_tmap(&tmap_dy);
        for (int s = 0; s < T8W_XS; ++s) { mbar_init(&bars->xfull[s], 1); mbar_init(&bars->xempty[s], 1); }
        for (int s = 0; s < T8W_YS; ++s) { mbar_init(&bars->yfull[s], 1); mbar_init(&bars->yempty[s], 1); }
        mbar_init(&bars->acc_full, 1);
        fence_barrier_init();
    }
    if (warp == 2) { tmem_alloc(&bars->tmem_base, 512); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;

    if (warp == 0) {
        if (lane == 0) {
            uint32_t xn = 0, yn = 0;
            const uint32_t op_bytes = p.planes == 2 ? T8_TILE : T8_TILE / 2;
            for (int u = split; u < p.n_units; u += p.n_splits) {
                const int b = u / p.AW, w0 = (u - b * p.AW) * 64;
                for (int n = 0; n < n_rows; ++n) {
                    {   // x row n
                        const int slot = xn % T8W_XS;
                        mbar_wait(&bars->xempty[slot], ((xn / T8W_XS) & 1) ^ 1);
                        mbar_expect_tx(&bars->xfull[slot], op_bytes);
                        tma_load_5d(x_ring + slot * T8_TILE, &tmap_x, &bars->xfull[slot], w0, n - p.P + i0, 0, b, 0);
                        ++xn;
                    }
                    const int r = n - (T8W_NT - 1);                           // dy row whose x window is now complete
                    if (r >= 0) {
                        const int slot = yn % T8W_YS;
                        mbar_wait(&bars->yempty[slot], ((yn / T8W_YS) & 1) ^ 1);
                        mbar_expect_tx(&bars->yfull[slot], op_bytes);
                        tma_load_5d(y_ring + slot * T8_TILE, &tmap_dy, &bars->yfull[slot], w0, r, 0, b, 0);
                        ++yn;
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = make_idesc_bf16(128, T8_N, 0, 0);
            const int n_cb = p.planes == 2 ? 3 : 1;
            uint32_t xn = 0, yn = 0, started = 0;
            for (int u = split; u < p.n_units; u += p.n_splits) {
                for (int n = 0; n < n_rows; ++n, ++xn) {
                    mbar_wait(&bars->xfull[xn % T8W_XS], (xn / T8W_XS) & 1);
                    const int r = n - (T8W_NT - 1);
                    if (r >= 0) {
                        const int yslot = yn % T8W_YS;
                        mbar_wait(&bars->yfull[yslot], (yn / T8W_YS) & 1);
                        tc_fence_after();
                        const uint64_t y_d0 = make_smem_desc(smem_u32(y_ring + yslot * T8_TILE), 16, 1024);
                        for (int k = 0; k < T8W_NT; ++k) {
                            const int h = r - p.P + i0 + k;                   // x row paired with dy row r for tap i0 + k
                            if (i0 + k >= p.kh || h < 0 || h >= p.H_src) continue;
                            const uint64_t x_d0 =
                                make_smem_desc(smem_u32(x_ring + ((xn - (T8W_NT - 1) + k) % T8W_XS) * T8_TILE), 16, 1024);
                            const uint32_t d_tmem = tmem_base + (uint32_t)k * T8_N;
#pragma unroll
                            for (int cb = 0; cb < 3; ++cb) {                   // (x hi, dy hi) (x hi, dy lo) (x lo, dy hi)
                                if (cb >= n_cb) break;
                                const uint64_t a_d = x_d0 + (uint64_t)(cb == 2 ? (128 * 128) >> 4 : 0);
                                const uint64_t b_d = y_d0 + (uint64_t)(cb == 1 ? (128 * 128) >> 4 : 0);
#pragma unroll
                                for (int ks = 0; ks < 4; ++ks)
                                    mma_bf16(d_tmem, a_d + (uint64_t)(ks * 2), b_d + (uint64_t)(ks * 2), idesc,
                                             ((started >> k) & 1u) | (uint32_t)(cb | ks));
                            }
                            started |= 1u << k;
                        }
                        // x row n - (NT - 1) leaves the window after this step
                        tc_commit(&bars->xempty[(xn - (T8W_NT - 1)) % T8W_XS]);
                        tc_commit(&bars->yempty[yslot]);
                        ++yn;
                    }
                }
                // the last NT - 1 x rows of the unit were never released by a step
                for (int k = 1; k < T8W_NT; ++k) tc_commit(&bars->xempty[(xn - T8W_NT + k) % T8W_XS]);
            }
            tc_commit(&bars->acc_full);
        }
    } else if (warp >= 4) {
        const int ew = warp & 3;
        const int ci = ew * 32 + lane;
        mbar_wait(&bars->acc_full, 0);
        tc_fence_after();
        for (int k = 0; k < T8W_NT; ++k) {
            const int tap = i0 + k;
            if (tap >= p.kh) continue;
            // touched iff some dy row r in [0, H_out) has its x row r - P + tap inside the source
            const int r_lo = max(0, p.P - tap), r_hi = min(p.H_out, p.H_src + p.P - tap);
            if (r_hi <= r_lo) continue;
            for (int n0 = 0; n0 < T8_N; n0 += 32) {
                uint32_t v[32];
                tmem_ld32(tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(k * T8_N + n0), v);
                tmem_ld_wait();
#pragma unroll
                for (int c = 0; c < 32; ++c)
                    atomicAdd(p.dw + ((size_t)(n0 + c) * T8_N + ci) * p.kh + tap, __uint_as_float(v[c]));
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// ---- host side --------------------------------------------------------------------------------------------
static bool t8_common(const cpc_conv_params* p) {
    return (p->precision == 0 || p->precision == 1) && p->kw == 1 && p->stride_h == 1 && p->stride_w == 1 &&
           p->pad_left == 0 && p->kh >= 4 &&
           p->w_out == p->w_in && p->pad_top < p->kh && (int64_t)p->batch * ((p->w_in + 63) / 64) < (1 << 28);
}
bool tall128_eligible(const cpc_conv_params* p, int which) {
    if (!t8_common(p)) return false;
    if (which == 0) return p->c_out == T8_N && p->c_in % 64 == 0 && p->c_in <= 512;
    if (which == 1) return p->c_in == T8_N && p->c_out % 64 == 0 && p->c_out <= 512;
    return p->c_in == T8_N && p->c_out == T8_N;
}

static int t8_planes(const cpc_conv_params* p) { return p->precision == 1 ? 1 : 2; }
static size_t t8_act_bytes(int B, int C, int H, int W, int planes) {
    const int Wp = (W + 7) & ~7;
    return align_up((size_t)planes * B * C * H * Wp * 2, 1024);
}
static size_t t8_w_bytes(int kh, int C) { return align_up((size_t)kh * (C / 64) * 2 * T8_N * 64 * 2, 1024); }

size_t tall128_workspace(const cpc_conv_params* p, int which) {
    if (!tall128_eligible(p, which)) return 0;
    const int planes = t8_planes(p);
    if (which == 2)
        return t8_act_bytes(p->batch, p->c_in, p->h_in, p->w_in, planes) +
               t8_act_bytes(p->batch, p->c_out, p->h_out, p->w_out, planes) + 1024;
    const int C = which == 0 ? p->c_in : p->c_out, H = which == 0 ? p->h_in : p->h_out;
    return t8_act_bytes(p->batch, C, H, p->w_in, planes) + t8_w_bytes(p->kh, C) + 256 + 1024;   // + tile counter
}

static bool t8_act_tmap(CUtensorMap* t, const void* base, int B, int C, int H, int Wp, int box_c, int planes) {
    const uint64_t rb = (uint64_t)Wp * 2;
    const uint64_t dims[5] = {(uint64_t)Wp, (uint64_t)H, (uint64_t)C, (uint64_t)B, (uint64_t)planes};
    const uint64_t strides[4] = {rb, rb * H, rb * H * C, rb * H * C * B};
    const uint32_t box[5] = {64, 1, (uint32_t)box_c, 1, (uint32_t)planes};
    return make_tmap_bf16(t, base, 5, dims, strides, box);
}

int tall128_conv_launch(const float* in, const float* w, const float* bias, float* out, const cpc_conv_params* p, int which,
                        const void* pre, void* workspace, size_t workspace_bytes, cudaStream_t s) {
    if (which > 1 || !tall128_eligible(p, which)) return CPC_ERR_UNSUPPORTED;
    if (!workspace || workspace_bytes < tall128_workspace(p, which)) return CPC_ERR_WORKSPACE;
    const int B = p->batch, W = p->w_in, Wp = (W + 7) & ~7;
    const int C = which == 0 ? p->c_in : p->c_out;
    const int H_src = which == 0 ? p->h_in : p->h_out;
    const int H_out = which == 0 ? p->h_out : p->h_in;
    const int P = which == 0 ? p->pad_top : p->kh - 1 - p->pad_top;
    const int planes = t8_planes(p);
    uint8_t* ws = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(workspace) + 1023) & ~(uintptr_t)1023);
    const __nv_bfloat16* act = pre ? reinterpret_cast<const __nv_bfloat16*>(pre) : reinterpret_cast<__nv_bfloat16*>(ws);
    __nv_bfloat16* wp = reinterpret_cast<__nv_bfloat16*>(ws + t8_act_bytes(B, C, H_src, W, planes));
    int* counter = reinterpret_cast<int*>(reinterpret_cast<uint8_t*>(wp) + t8_w_bytes(p->kh, C));
    int st = pre ? CPC_OK
                 : pack_split_launch(in, reinterpret_cast<__nv_bfloat16*>(ws), (long)B * C * H_src, W, Wp, planes, 1, 1, 1, 0, s);
    if (st != CPC_OK) return st;
    {
        const long total = (long)p->kh * (C / 64) * T8_N * 64;
        int blocks = (int)((total + 255) / 256);
        if (blocks > 148 * 8) blocks = 148 * 8;
        tall128_pack_weights_kernel<<<blocks, 256, 0, s>>>(w, wp, p->c_in, p->kh, C, which, counter);
        CPC_LAUNCH_CHECK();
    }
    CUtensorMap ta, tw;
    if (!t8_act_tmap(&ta, act, B, C, H_src, Wp, 64, planes)) return CPC_ERR_CUDA;
    {
        const uint64_t n_tq = (uint64_t)p->kh * (C / 64);
        const uint64_t dims[3] = {64, 2 * T8_N, n_tq};
        const uint64_t strides[2] = {128, 128 * 2 * T8_N};
        const uint32_t box[3] = {64, (uint32_t)(planes * T8_N), 1};      // bf16 operand mode: the hi rows of a tap only
        if (!make_tmap_bf16(&tw, wp, 3, dims, strides, box)) return CPC_ERR_CUDA;
    }
    Tall128Conv k{};
    k.B = B; k.H_src = H_src; k.H_out = H_out; k.W = W; k.AW = (W + 63) / 64; k.kh = p->kh; k.P = P; k.NQ = C / 64;
    k.n_units = B * k.AW; k.n_pairs = (k.n_units + 1) / 2; k.n_rtiles = ceil_div(H_out, T8_R);
    k.n_tiles = k.n_pairs * k.n_rtiles;
    k.relu = which == 0 ? p->relu : 0; k.bias = which == 0 ? bias : nullptr; k.out = out;
    k.planes = planes;
    k.counter = counter;
    if (cudaFuncSetAttribute(tall128_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, T8_SMEM) != cudaSuccess)
        return CPC_ERR_CUDA;
    const int grid = k.n_tiles < 148 ? k.n_tiles : 148;
    tall128_conv_kernel<<<grid, T8_THREADS, T8_SMEM, s>>>(ta, tw, k);
    CPC_LAUNCH_CHECK();
    count_launch(3);
    return CPC_OK;
}

int tall128_wgrad_launch(const float* x, const float* dy, float* dw, const cpc_conv_params* p, const void* pre_x,
                         const void* pre_dy, void* workspace, size_t workspace_bytes, cudaStream_t s) {
    if (!tall128_eligible(p, 2)) return CPC_ERR_UNSUPPORTED;
    if (!workspace || workspace_bytes < tall128_workspace(p, 2)) return CPC_ERR_WORKSPACE;
    const int B = p->batch, W = p->w_in, Wp = (W + 7) & ~7;
    const int planes = t8_planes(p);
    const size_t x_bytes = t8_act_bytes(B, T8_N, p->h_in, W, planes);
    uint8_t* ws = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(workspace) + 1023) & ~(uintptr_t)1023);
    const __nv_bfloat16* xp = pre_x ? reinterpret_cast<const __nv_bfloat16*>(pre_x) : reinterpret_cast<__nv_bfloat16*>(ws);
    const __nv_bfloat16* dyp = pre_dy ? reinterpret_cast<const __nv_bfloat16*>(pre_dy)
                                      : reinterpret_cast<__nv_bfloat16*>(ws + x_bytes);
    int st = pre_x ? CPC_OK
                   : pack_split_launch(x, reinterpret_cast<__nv_bfloat16*>(ws), (long)B * T8_N * p->h_in, W, Wp, planes, 1, 1, 1, 0, s);
    if (st != CPC_OK) return st;
    st = pre_dy ? CPC_OK : pack_split_launch(dy, reinterpret_cast<__nv_bfloat16*>(ws + x_bytes), (long)B * T8_N * p->h_out, W, Wp,
                                             planes, 1, 1, 1, 0, s);
    if (st != CPC_OK) return st;
    CUtensorMap tx, tdy;
    if (!t8_act_tmap(&tx, xp, B, T8_N, p->h_in, Wp, T8_N, planes)) return CPC_ERR_CUDA;
    if (!t8_act_tmap(&tdy, dyp, B, T8_N, p->h_out, Wp, T8_N, planes)) return CPC_ERR_CUDA;
    Tall128Wgrad k{};
    k.B = B; k.H_src = p->h_in; k.H_out = p->h_out; k.W = W; k.AW = (W + 63) / 64; k.kh = p->kh; k.P = p->pad_top;
    k.n_units = B * k.AW;
    k.planes = planes;
    k.n_tgroups = ceil_div(p->kh, T8W_NT);
    k.n_splits = 148 / k.n_tgroups;
    if (k.n_splits < 1) k.n_splits = 1;
    if (k.n_splits > k.n_units) k.n_splits = k.n_units;
    k.dw = dw;
    if (cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)T8_N * T8_N * p->kh, s) != cudaSuccess) return CPC_ERR_CUDA;
    if (cudaFuncSetAttribute(tall128_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, T8W_SMEM) != cudaSuccess)
        return CPC_ERR_CUDA;
    tall128_wgrad_kernel<<<k.n_tgroups * k.n_splits, T8_THREADS, T8W_SMEM, s>>>(tx, tdy, k);
    CPC_LAUNCH_CHECK();
    count_launch(3);
    return CPC_OK;
}

}  // namespace cpc
