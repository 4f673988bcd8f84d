// Row-streaming tcgen05 kernels for the "tall" pitch convolutions of the scalogram encoders
// (kernel kh x 1, stride 1, 32 -> 32 channels; scalogram_model.py:387-431, e.g. the 64x1 conv of
// scalogram_resnet_architecture_7 block 0), forward, data gradient and weight gradient.
//
// Why: in the generic implicit-GEMM kernel (conv_umma.cu) a kh x 1 conv re-fetches its activation tile
// from L2 once per vertical tap (kh = 64 -> 26 GB of L2->SM traffic per pass at B = 64), which bounds it
// at ~50 TFLOP/s.  Here an input row is brought into shared memory ONCE per tile and multiplied against
// every output row it contributes to.
//
// Forward / data gradient (tall_conv_kernel):
//   out[b, n, r, w] = sum_{c, i} src[b, c, r - P + i, w] * Wt[n, c, i]           (P = zero rows on top)
//   * tile = 128 pixels (two 64-pixel atoms along w) x R = 16 output rows; the 16 x 32 = 512 fp32
//     accumulator columns fill TMEM exactly (column = rho * 32 + n).
//   * step j streams input row j of the tile (MN-major A operand: K rows = [hi c0..31 | lo c0..31]) and
//     weight tap j into a ring of tap slots stored in DESCENDING tap order, so the B operand of the
//     output-row group {4g .. 4g+3} -- taps {j-4g, .., j-4g-3} stacked along N = 128 -- is one contiguous
//     K-major block: a full-rate 128 x 128 x 16 MMA instead of four 128 x 32 x 16 ones.
//   * fp32-faithful arithmetic = bf16 hi/lo split, products hi*hi + hi*lo + lo*hi (both planes of a
//     weight row share one 128-byte swizzle row: [hi c0..31 | lo c0..31]).
//   * the data gradient is the same kernel on dy with channel roles swapped and taps flipped.
//   * tiles are handed out by an atomic counter (zeroed by the weight-packing kernel that precedes the launch): the cost
//     of a tile depends on how many of its taps fall into the zero padding (31 ... 94 steps in arch-7 block 0), a static
//     round-robin gave some CTAs only cheap and others only expensive tiles (83 % balance), and a CTA that starts late
//     because another kernel (an NCCL all-reduce on the communication stream) holds its SM simply takes fewer tiles.
//   * bf16 operand mode (cpc_conv_params.precision = 1): the activation operand has its hi plane only (half the
//     bytes per row) and one product, hi*hi, is issued instead of three.
// Weight gradient (tall_wgrad_kernel):
//   dW[co, ci, i] = sum_{b, r, w} dy[b, co, r, w] * x[b, ci, r - P + i, w]
//   * both operands are K-major over 64-pixel chunks; M = (ci, 4 x rows), N = (co, 4 dy rows), so one
//     128 x 128 accumulator block holds the 16 (x row, dy row) pairs of a row-group pair, i.e. taps
//     base + delta - rho with base = h0 - r0 + P.  A CTA owns 3 consecutive bases (384 TMEM columns) and
//     a share of the pixel atoms; x row groups slide through a 4-slot ring and are used by 3 steps each.
//   * flush: red.global.add of every valid (tap) entry into the zero-initialised dW.
#include "common.cuh"
#include "umma.cuh"

namespace cpc {
using namespace umma;

int pack_split_launch(const float* x, __nv_bfloat16* out, long rows, int W, int Wp, int planes, int nrep, int w_mul,
                      int rep_mul, int w_off, cudaStream_t s);   // conv_umma.cu

constexpr int TL_C = 32;                     // channels on both sides
constexpr int TL_R = 16;                     // output rows per tile
constexpr int TL_S = 5;                      // activation-row stages
constexpr int TL_T = 20;                     // tap ring slots (multiple of 4, >= R - 1 + S)
constexpr int TL_SLOT = TL_C * 128;          // 4 KB: 32 n-rows x (hi 32 ch | lo 32 ch)
constexpr int TL_ASTAGE = 2 * 64 * 128;      // 16 KB: 2 atoms x 64 K-rows x 64 pixels
constexpr int TL_THREADS = 256;
constexpr int TL_SMEM = TL_S * TL_ASTAGE + (TL_T + 3) * TL_SLOT + 1024 + 256;

struct TallConv {
    int B, H_src, H_out, W, AW;              // AW = 64-pixel atoms per row
    int kh, P;                               // taps, zero rows above the source
    int n_units, n_pairs, n_rtiles, n_tiles;
    int relu;
    int planes;                              // 2: bf16 hi/lo (fp32-faithful), 1: hi plane only (bf16 operand mode)
    int rt_rev;                              // row tiles of a pixel pair are handed out bottom-up (expensive ones first)
    const float* bias;
    float* out;                              // (B, 32, H_out, W) fp32
    int* counter;                            // tile scheduler (workspace; zero at launch)
};

struct __align__(8) TallBarriers {
    uint64_t full[TL_S], empty[TL_S], acc_full, acc_empty;
    uint64_t sfull[2], sempty[2];            // tile-id queue: scheduler (producer lane) -> MMA warp + 4 epilogue warps
    int tile_id[2];
    uint32_t tmem_base;
};

// weights (Cout, Cin, kh, 1) fp32 -> bf16 [tap][n][hi c0..31 | lo c0..31]
//   flip_swap = 0: n = co, c = ci, tap = i            (forward)
//   flip_swap = 1: n = ci, c = co, tap = kh - 1 - i   (data gradient)
__global__ void __launch_bounds__(256) tall_pack_weights_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out,
                                                               int kh, int flip_swap, int* __restrict__ counter) {
    if (blockIdx.x == 0 && threadIdx.x == 0) *counter = 0;        // tile scheduler of the conv kernel that follows
    const int total = kh * TL_C * TL_C;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const int c = idx % TL_C;
        const int n = (idx / TL_C) % TL_C;
        const int tap = idx / (TL_C * TL_C);
        const int i = flip_swap ? kh - 1 - tap : tap;
        const int co = flip_swap ? c : n, ci = flip_swap ? n : c;
        const float v = __ldg(w + ((size_t)co * TL_C + ci) * kh + i);
        const __nv_bfloat16 hi = __float2bfloat16_rn(v);
        __nv_bfloat16* row = out + ((size_t)tap * TL_C + n) * 64;
        row[c] = hi;
        row[32 + c] = __float2bfloat16_rn(v - __bfloat162float(hi));
    }
}

__device__ __forceinline__ void tall_tile_rows(const TallConv& p, int tile, int& pair, int& r0, int& j_lo, int& j_hi) {
    pair = tile / p.n_rtiles;
    const int rt = tile - pair * p.n_rtiles;
    r0 = (p.rt_rev ? p.n_rtiles - 1 - rt : rt) * TL_R;
    j_lo = max(0, p.P - r0);
    j_hi = min(TL_R + p.kh - 1, p.H_src + p.P - r0);
    if (j_hi < j_lo) j_hi = j_lo;
}

__global__ void __launch_bounds__(TL_THREADS, 1) tall_conv_kernel(const __grid_constant__ CUtensorMap tmap_a,
                                                                 const __grid_constant__ CUtensorMap tmap_w,
                                                                 const TallConv p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* a_ring = smem;
    uint8_t* w_ring = smem + TL_S * TL_ASTAGE;
    TallBarriers* bars = reinterpret_cast<TallBarriers*>(w_ring + (TL_T + 3) * TL_SLOT);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmap_a);
        prefetch_tmap(&tmap_w);
        for (int s = 0; s < TL_S; ++s) { mbar_init(&bars->full[s], 1); mbar_init(&bars->empty[s], 1); }
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
        // ===== tile scheduler + TMA producer: one step = (weight tap j, optionally input row j) =====
        if (lane == 0) {
            uint32_t v = 0;                                   // running step counter -> stage and tap slot
            for (uint32_t ti = 0;; ++ti) {
                const int tile = sched_push(bars, ti, p.counter, p.n_tiles);
                if (tile < 0) break;
                int pair, r0, j_lo, j_hi;
                tall_tile_rows(p, tile, pair, r0, j_lo, j_hi);
                if (j_hi == j_lo) continue;                   // no input row inside the source: bias-only tile
                const uint32_t a_bytes = p.planes == 2 ? TL_ASTAGE : TL_ASTAGE / 2;   // hi plane only: rows 0..31 of an atom
                int ab[2], aw0[2];
                for (int a = 0; a < 2; ++a) {
                    const int u = pair * 2 + a;               // unit >= n_units -> batch index out of range -> zeros
                    ab[a] = u / p.AW;
                    aw0[a] = (u - ab[a] * p.AW) * 64;
                }
                for (int j = j_lo - (TL_R - 1); j < j_hi; ++j, ++v) {
                    const int stage = v % TL_S;
                    const uint32_t phase = (v / TL_S) & 1;
                    mbar_wait(&bars->empty[stage], phase ^ 1);
                    const int slot = (TL_T - (int)(v % TL_T)) % TL_T;
                    const bool with_row = j >= j_lo;
                    const uint32_t bytes = TL_SLOT * (slot < 3 ? 2 : 1) + (with_row ? a_bytes : 0);
                    mbar_expect_tx(&bars->full[stage], bytes);
                    tma_load_3d(w_ring + slot * TL_SLOT, &tmap_w, &bars->full[stage], 0, 0, j);   // j outside [0,kh) -> zeros
                    if (slot < 3) tma_load_3d(w_ring + (TL_T + slot) * TL_SLOT, &tmap_w, &bars->full[stage], 0, 0, j);
                    if (with_row) {
                        uint8_t* st = a_ring + stage * TL_ASTAGE;
                        for (int a = 0; a < 2; ++a)
                            tma_load_5d(st + a * (64 * 128), &tmap_a, &bars->full[stage], aw0[a], r0 - p.P + j, 0, ab[a], 0);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: warp-uniform loop, one elected lane issues; descriptors advance by constant adds =====
        {
            const uint32_t idesc = make_idesc_bf16(128, 128, /*A MN-major*/ 1, /*B K-major*/ 0);
            const uint32_t w_base = smem_u32(w_ring);
            uint32_t v = 0, acc_phase = 0;
            for (uint32_t ti = 0;; ++ti) {
                const int tile = sched_pop(bars, ti, lane);
                if (tile < 0) break;
                int pair, r0, j_lo, j_hi;
                tall_tile_rows(p, tile, pair, r0, j_lo, j_hi);
                if (j_hi == j_lo) continue;
                mbar_wait(&bars->acc_empty, acc_phase ^ 1);
                tc_fence_after();
                for (int j = j_lo - (TL_R - 1); j < j_hi; ++j, ++v) {
                    const int stage = v % TL_S;
                    const uint32_t phase = (v / TL_S) & 1;
                    mbar_wait(&bars->full[stage], phase);
                    tc_fence_after();
                    if (elect_one()) {
                        if (j >= j_lo) {
                            // A (MN-major): hi plane rows 0..31, lo plane rows 32..63, 16 K-rows (2 KB) per K step
                            const uint64_t a_hi = make_smem_desc(smem_u32(a_ring + stage * TL_ASTAGE), 64 * 128, 1024);
                            const uint64_t a_lo = a_hi + (uint64_t)((32 * 128) >> 4);
                            const uint32_t accum0 = j != j_lo ? 1u : 0u;
                            const int vslot = (int)(v % TL_T);
                            for (int g = 0; g < TL_R / 4; ++g) {
                                const int t_hi = j - 4 * g;                    // newest tap of the group (row 4g)
                                const bool live = t_hi >= 0 && t_hi - 3 < p.kh;    // any tap inside [0, kh)
                                if (j != j_lo && !live) continue;              // first step initialises every group
                                int start = 4 * g + TL_T - vslot;              // slot of tap j - 4g
                                if (start >= TL_T) start -= TL_T;
                                // B (K-major): [hi c0..31 | lo c0..31] per 128-byte row, 32 B per K step
                                const uint64_t b_hi = make_smem_desc(w_base + start * TL_SLOT, 16, 1024);
                                const uint64_t b_lo = b_hi + (uint64_t)(64 >> 4);
                                const uint32_t d_tmem = tmem_base + (uint32_t)g * 128;
                                mma_bf16(d_tmem, a_hi, b_hi, idesc, accum0);                                   // hi * hi
                                mma_bf16(d_tmem, a_hi + (uint64_t)((16 * 128) >> 4), b_hi + 2, idesc, 1u);
                                if (p.planes == 1) continue;                                                   // bf16 operand mode
                                mma_bf16(d_tmem, a_hi, b_lo, idesc, 1u);                                       // hi * lo
                                mma_bf16(d_tmem, a_hi + (uint64_t)((16 * 128) >> 4), b_lo + 2, idesc, 1u);
                                mma_bf16(d_tmem, a_lo, b_hi, idesc, 1u);                                       // lo * hi
                                mma_bf16(d_tmem, a_lo + (uint64_t)((16 * 128) >> 4), b_hi + 2, idesc, 1u);
                            }
                        }
                        tc_commit(&bars->empty[stage]);
                    }
                    __syncwarp();
                }
                if (elect_one()) tc_commit(&bars->acc_full);
                __syncwarp();
                acc_phase ^= 1;
            }
        }
    } else if (warp >= 4) {
        // ===== epilogue: TMEM -> registers -> bias / ReLU -> NCHW fp32 =====
        const int ew = warp & 3;
        const int r = ew * 32 + lane;                                          // tile pixel
        uint32_t acc_phase = 0;
        for (uint32_t ti = 0;; ++ti) {
            const int tile = sched_pop(bars, ti, lane);
            if (tile < 0) break;
            int pair, r0, j_lo, j_hi;
            tall_tile_rows(p, tile, pair, r0, j_lo, j_hi);
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
            for (int rho = 0; rho < TL_R; ++rho) {
                const int row = r0 + rho;
                if (row >= p.H_out) break;
                uint32_t v[32];
                if (!empty_tile) {
                    tmem_ld32(tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)rho * 32, v);
                    tmem_ld_wait();
                } else {
#pragma unroll
                    for (int c = 0; c < 32; ++c) v[c] = 0u;
                }
                if (valid) {
                    float* o = p.out + ((size_t)b * TL_C * p.H_out + row) * p.W + pw;
#pragma unroll
                    for (int c = 0; c < 32; ++c) {
                        float f = __uint_as_float(v[c]);
                        if (p.bias) f += __ldg(p.bias + c);
                        if (p.relu) f = fmaxf(f, 0.f);
                        o[(size_t)c * chan_stride] = f;
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

// ---- weight gradient ------------------------------------------------------------------------------------
constexpr int TW_GROUP = 2 * 128 * 128;      // 32 KB: 2 planes x (32 ch x 4 rows) x 64 pixels
constexpr int TW_XS = 4;                     // x row-group ring
constexpr int TW_YS = 2;                     // dy row-group ring
constexpr int TW_NB = 3;                     // tap bases (accumulator blocks) per CTA
constexpr int TW_SMEM = (TW_XS + TW_YS) * TW_GROUP + 1024 + 256;

struct TallWgrad {
    int B, H_src, H_out, W, AW, kh, P;
    int n_units, n_splits, n_dgroups;
    int n_blocks;                            // tap blocks that hold at least one tap inside [0, kh)
    int balanced;                            // 1: group g owns CTAs cta_start[g] .. cta_start[g + 1] (work-proportional)
    int cta_start[kMaxSplitGroups + 1];
    int d_min;                               // block d covers taps 4*d + P + delta - rho
    int Q;                                   // dy row groups
    int planes;                              // 2: hi/lo planes, 1: hi plane only (bf16 operand mode)
    float* dw;                               // (32, 32, kh) fp32, zero-initialised
};

struct __align__(8) TallWgradBarriers {
    uint64_t xfull[TW_XS], xempty[TW_XS], yfull[TW_YS], yempty[TW_YS], acc_full;
    uint32_t tmem_base;
};

__device__ __forceinline__ int floor_div4(int a) { return a >= 0 ? a >> 2 : -((3 - a) >> 2); }

__global__ void __launch_bounds__(TL_THREADS, 1) tall_wgrad_kernel(const __grid_constant__ CUtensorMap tmap_x,
                                                                  const __grid_constant__ CUtensorMap tmap_dy,
                                                                  const TallWgrad p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* x_ring = smem;
    uint8_t* y_ring = smem + TW_XS * TW_GROUP;
    TallWgradBarriers* bars = reinterpret_cast<TallWgradBarriers*>(y_ring + TW_YS * TW_GROUP);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // accumulator group and share of the pixel atoms: groups whose taps mostly pair with rows outside the source do less
    // work per atom and get fewer CTAs (host: balance_group_ctas)
    int dg, split, n_splits;
    if (p.balanced) {
        dg = 0;
        while (dg + 1 < p.n_dgroups && (int)blockIdx.x >= p.cta_start[dg + 1]) ++dg;
        split = (int)blockIdx.x - p.cta_start[dg];
        n_splits = p.cta_start[dg + 1] - p.cta_start[dg];
    } else {
        dg = blockIdx.x % p.n_dgroups;
        split = blockIdx.x / p.n_dgroups;
        n_splits = p.n_splits;
    }
    const int d0 = p.d_min + dg * TW_NB;     // x group index = dy group index + d
    const int nb_live = min(TW_NB, p.n_blocks - dg * TW_NB);     // blocks of this group that hold a tap inside [0, kh)

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmap_x);
        prefetch_tmap(&tmap_dy);
        for (int s = 0; s < TW_XS; ++s) { mbar_init(&bars->xfull[s], 1); mbar_init(&bars->xempty[s], 1); }
        for (int s = 0; s < TW_YS; ++s) { mbar_init(&bars->yfull[s], 1); mbar_init(&bars->yempty[s], 1); }
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
            const uint32_t g_bytes = p.planes == 2 ? TW_GROUP : TW_GROUP / 2;
            auto load_x = [&](int b, int w0, int n) {            // x row group d0 + n of the current atom
                const int slot = xn % TW_XS;
                mbar_wait(&bars->xempty[slot], ((xn / TW_XS) & 1) ^ 1);
                mbar_expect_tx(&bars->xfull[slot], g_bytes);
                tma_load_5d(x_ring + slot * TW_GROUP, &tmap_x, &bars->xfull[slot], w0, 4 * (d0 + n), 0, b, 0);
                ++xn;
            };
            for (int u = split; u < p.n_units; u += n_splits) {
                const int b = u / p.AW, w0 = (u - b * p.AW) * 64;
                load_x(b, w0, 0);
                load_x(b, w0, 1);
                for (int q = 0; q < p.Q; ++q) {
                    load_x(b, w0, q + 2);
                    const int slot = yn % TW_YS;
                    mbar_wait(&bars->yempty[slot], ((yn / TW_YS) & 1) ^ 1);
                    mbar_expect_tx(&bars->yfull[slot], g_bytes);
                    tma_load_5d(y_ring + slot * TW_GROUP, &tmap_dy, &bars->yfull[slot], w0, 4 * q, 0, b, 0);
                    ++yn;
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = make_idesc_bf16(128, 128, 0, 0);
            const int n_xgroups = (p.H_src + 3) >> 2;
            const int n_cb = p.planes == 2 ? 3 : 1;
            uint32_t xn = 0, yn = 0, started = 0;
            for (int u = split; u < p.n_units; u += n_splits) {
                // x groups n = 0, 1 of this atom are consumed together with n = 2 at step 0
                for (int q = 0; q < p.Q; ++q) {
                    // wait for the newest x group of this step (and, at q = 0, the two before it)
                    const int first = q == 0 ? 0 : 2;
                    for (int k = first; k < 3; ++k) {
                        const uint32_t n = xn + k;
                        mbar_wait(&bars->xfull[n % TW_XS], (n / TW_XS) & 1);
                    }
                    const int yslot = yn % TW_YS;
                    mbar_wait(&bars->yfull[yslot], (yn / TW_YS) & 1);
                    tc_fence_after();
                    const uint64_t y_d0 = make_smem_desc(smem_u32(y_ring + yslot * TW_GROUP), 16, 1024);
                    for (int k = 0; k < nb_live; ++k) {
                        const int pg = q + d0 + k;                            // x row group index
                        if (pg < 0 || pg >= n_xgroups) continue;              // rows outside the source: zeros
                        const uint64_t x_d0 = make_smem_desc(smem_u32(x_ring + ((xn + k) % TW_XS) * TW_GROUP), 16, 1024);
                        const uint32_t d_tmem = tmem_base + (uint32_t)k * 128;
#pragma unroll
                        for (int cb = 0; cb < 3; ++cb) {                       // (x hi, dy hi) (x hi, dy lo) (x lo, dy hi)
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
                    tc_commit(&bars->xempty[xn % TW_XS]);                     // oldest x group is done after this step
                    tc_commit(&bars->yempty[yslot]);
                    ++xn;
                    ++yn;
                }
                // the last two x groups of the atom were loaded but belong to no further step
                tc_commit(&bars->xempty[xn % TW_XS]); ++xn;
                tc_commit(&bars->xempty[xn % TW_XS]); ++xn;
            }
            tc_commit(&bars->acc_full);
        }
    } else if (warp >= 4) {
        const int ew = warp & 3;
        const int m = ew * 32 + lane;                                          // accumulator row -> (ci, delta)
        const int ci = m >> 2, delta = m & 3;
        mbar_wait(&bars->acc_full, 0);
        tc_fence_after();
        const int n_xgroups = (p.H_src + 3) >> 2;
        for (int k = 0; k < nb_live; ++k) {
            // the block was touched iff some dy group q in [0, Q) pairs with an x group inside the source
            const int q_lo = max(0, -(d0 + k)), q_hi = min(p.Q, n_xgroups - (d0 + k));
            if (q_hi <= q_lo) continue;
            const int base = 4 * (d0 + k) + p.P;
            for (int n0 = 0; n0 < 128; n0 += 32) {
                uint32_t v[32];
                tmem_ld32(tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(k * 128 + n0), v);
                tmem_ld_wait();
#pragma unroll
                for (int c = 0; c < 32; ++c) {
                    const int n = n0 + c;
                    const int co = n >> 2, rho = n & 3;
                    const int tap = base + delta - rho;
                    if (tap >= 0 && tap < p.kh)
                        atomicAdd(p.dw + ((size_t)co * TL_C + ci) * p.kh + tap, __uint_as_float(v[c]));
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// ---- host side ------------------------------------------------------------------------------------------
static bool tall_shape_ok(const cpc_conv_params* p) {
    return (p->precision == 0 || p->precision == 1) && p->kw == 1 && p->stride_h == 1 && p->stride_w == 1 &&
           p->pad_left == 0 && p->c_in == TL_C && p->c_out == TL_C && p->kh >= 8 && p->w_out == p->w_in && p->pad_top < p->kh &&
           (int64_t)p->batch * ((p->w_in + 63) / 64) < (1 << 28);
}
bool tall_conv_eligible(const cpc_conv_params* p, int which) { (void)which; return tall_shape_ok(p); }

static int tall_planes(const cpc_conv_params* p) { return p->precision == 1 ? 1 : 2; }
static size_t tall_act_bytes(int B, int H, int W, int planes) {
    const int Wp = (W + 7) & ~7;
    return align_up((size_t)planes * B * TL_C * H * Wp * 2, 1024);
}
static size_t tall_w_bytes(int kh) { return align_up((size_t)kh * TL_C * 64 * 2, 1024); }

size_t tall_conv_workspace(const cpc_conv_params* p, int which) {
    if (!tall_shape_ok(p)) return 0;
    const int planes = tall_planes(p);
    if (which == 2)
        return tall_act_bytes(p->batch, p->h_in, p->w_in, planes) + tall_act_bytes(p->batch, p->h_out, p->w_out, planes) + 1024;
    const int H = which == 0 ? p->h_in : p->h_out;
    return tall_act_bytes(p->batch, H, p->w_in, planes) + tall_w_bytes(p->kh) + 256 + 1024;   // + tile counter
}

static bool tall_act_tmap(CUtensorMap* t, const void* base, int B, int H, int Wp, int box_rows, int planes) {
    const uint64_t rb = (uint64_t)Wp * 2;
    const uint64_t dims[5] = {(uint64_t)Wp, (uint64_t)H, (uint64_t)TL_C, (uint64_t)B, (uint64_t)planes};
    const uint64_t strides[4] = {rb, rb * H, rb * H * TL_C, rb * H * TL_C * B};
    const uint32_t box[5] = {64, (uint32_t)box_rows, (uint32_t)TL_C, 1, (uint32_t)planes};
    return make_tmap_bf16(t, base, 5, dims, strides, box);
}

// which = 0: y = conv(x, w) + bias [relu];  which = 1: dx = conv_transpose(dy, w)
int tall_conv_launch(const float* in, const float* w, const float* bias, float* out, const cpc_conv_params* p, int which,
                     const void* pre, void* workspace, size_t workspace_bytes, cudaStream_t s) {
    if (!tall_shape_ok(p)) return CPC_ERR_UNSUPPORTED;
    if (!workspace || workspace_bytes < tall_conv_workspace(p, which)) return CPC_ERR_WORKSPACE;
    const int B = p->batch, W = p->w_in, Wp = (W + 7) & ~7;
    const int H_src = which == 0 ? p->h_in : p->h_out;
    const int H_out = which == 0 ? p->h_out : p->h_in;
    const int P = which == 0 ? p->pad_top : p->kh - 1 - p->pad_top;
    const int planes = tall_planes(p);
    uint8_t* ws = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(workspace) + 1023) & ~(uintptr_t)1023);
    const __nv_bfloat16* act = pre ? reinterpret_cast<const __nv_bfloat16*>(pre) : reinterpret_cast<__nv_bfloat16*>(ws);
    __nv_bfloat16* wp = reinterpret_cast<__nv_bfloat16*>(ws + tall_act_bytes(B, H_src, W, planes));
    int st = pre ? CPC_OK
                 : pack_split_launch(in, reinterpret_cast<__nv_bfloat16*>(ws), (long)B * TL_C * H_src, W, Wp, planes, 1, 1, 1, 0, s);
    if (st != CPC_OK) return st;
    int* counter = reinterpret_cast<int*>(reinterpret_cast<uint8_t*>(wp) + tall_w_bytes(p->kh));
    tall_pack_weights_kernel<<<ceil_div(p->kh * TL_C * TL_C, 256), 256, 0, s>>>(w, wp, p->kh, which, counter);
    CPC_LAUNCH_CHECK();
    CUtensorMap ta, tw;
    if (!tall_act_tmap(&ta, act, B, H_src, Wp, 1, planes)) return CPC_ERR_CUDA;
    {
        const uint64_t dims[3] = {64, (uint64_t)TL_C, (uint64_t)p->kh};
        const uint64_t strides[2] = {128, 128 * TL_C};
        const uint32_t box[3] = {64, (uint32_t)TL_C, 1};
        if (!make_tmap_bf16(&tw, wp, 3, dims, strides, box)) return CPC_ERR_CUDA;
    }
    TallConv k{};
    k.B = B; k.H_src = H_src; k.H_out = H_out; k.W = W; k.AW = (W + 63) / 64;
    k.kh = p->kh; k.P = P;
    k.n_units = B * k.AW; k.n_pairs = (k.n_units + 1) / 2; k.n_rtiles = ceil_div(H_out, TL_R);
    k.n_tiles = k.n_pairs * k.n_rtiles;
    k.relu = which == 0 ? p->relu : 0; k.bias = which == 0 ? bias : nullptr; k.out = out;
    k.planes = planes;
    k.counter = counter;
    {
        // steps of the first and the last row tile of a pair (tall_tile_rows): the cheaper end goes last, so that the
        // tiles still running when the counter runs out are short ones
        auto steps = [&](int r0) {
            const int j_lo = P - r0 > 0 ? P - r0 : 0;
            const int j_hi = TL_R + p->kh - 1 < H_src + P - r0 ? TL_R + p->kh - 1 : H_src + P - r0;
            return j_hi > j_lo ? j_hi - j_lo + TL_R - 1 : 0;
        };
        k.rt_rev = steps(0) < steps((k.n_rtiles - 1) * TL_R) ? 1 : 0;
    }
    if (cudaFuncSetAttribute(tall_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TL_SMEM) != cudaSuccess)
        return CPC_ERR_CUDA;
    const int grid = k.n_tiles < 148 ? k.n_tiles : 148;
    tall_conv_kernel<<<grid, TL_THREADS, TL_SMEM, s>>>(ta, tw, k);
    CPC_LAUNCH_CHECK();
    count_launch(3);
    return CPC_OK;
}

int tall_wgrad_launch(const float* x, const float* dy, float* dw, const cpc_conv_params* p, const void* pre_x,
                      const void* pre_dy, void* workspace, size_t workspace_bytes, cudaStream_t s) {
    if (!tall_shape_ok(p)) return CPC_ERR_UNSUPPORTED;
    if (!workspace || workspace_bytes < tall_conv_workspace(p, 2)) return CPC_ERR_WORKSPACE;
    const int B = p->batch, W = p->w_in, Wp = (W + 7) & ~7;
    const int planes = tall_planes(p);
    const size_t x_bytes = tall_act_bytes(B, p->h_in, W, planes);
    uint8_t* ws = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(workspace) + 1023) & ~(uintptr_t)1023);
    const __nv_bfloat16* xp = pre_x ? reinterpret_cast<const __nv_bfloat16*>(pre_x) : reinterpret_cast<__nv_bfloat16*>(ws);
    const __nv_bfloat16* dyp = pre_dy ? reinterpret_cast<const __nv_bfloat16*>(pre_dy)
                                      : reinterpret_cast<__nv_bfloat16*>(ws + x_bytes);
    int st = pre_x ? CPC_OK
                   : pack_split_launch(x, reinterpret_cast<__nv_bfloat16*>(ws), (long)B * TL_C * p->h_in, W, Wp, planes, 1, 1, 1, 0, s);
    if (st != CPC_OK) return st;
    st = pre_dy ? CPC_OK : pack_split_launch(dy, reinterpret_cast<__nv_bfloat16*>(ws + x_bytes), (long)B * TL_C * p->h_out, W, Wp,
                                             planes, 1, 1, 1, 0, s);
    if (st != CPC_OK) return st;
    CUtensorMap tx, tdy;
    if (!tall_act_tmap(&tx, xp, B, p->h_in, Wp, 4, planes)) return CPC_ERR_CUDA;
    if (!tall_act_tmap(&tdy, dyp, B, p->h_out, Wp, 4, planes)) return CPC_ERR_CUDA;
    TallWgrad k{};
    k.B = B; k.H_src = p->h_in; k.H_out = p->h_out; k.W = W; k.AW = (W + 63) / 64; k.kh = p->kh; k.P = p->pad_top;
    k.n_units = B * k.AW;
    k.Q = (p->h_out + 3) / 4;
    k.planes = planes;
    // blocks d with some tap 4d + P + delta - rho in [0, kh):  4d + P + 3 >= 0  and  4d + P - 3 <= kh - 1
    const int d_lo = -((k.P + 3) / 4);                                  // ceil((-3 - P) / 4)
    int d_hi = p->kh + 2 - k.P;                                          // floor((kh + 2 - P) / 4)
    d_hi = d_hi >= 0 ? d_hi / 4 : -((3 - d_hi) / 4);
    const int n_blocks = d_hi - d_lo + 1;
    k.d_min = d_lo;
    k.n_dgroups = ceil_div(n_blocks, TW_NB);
    k.n_blocks = n_blocks;
    k.n_splits = 148 / k.n_dgroups;
    if (k.n_splits < 1) k.n_splits = 1;
    if (k.n_splits > k.n_units) k.n_splits = k.n_units;
    int n_ctas = k.n_dgroups * k.n_splits;
    {
        // cost of one pixel atom for group g, in MMA blocks: (dy group q, x group q + d) pairs inside the source over its
        // live blocks, but never less than the time its Q ring steps take to load (~2 block times per step)
        int cost[kMaxSplitGroups];
        const int n_xg = (p->h_in + 3) / 4;
        bool ok = k.n_dgroups <= kMaxSplitGroups;
        for (int g = 0; ok && g < k.n_dgroups; ++g) {
            int c = 0;
            for (int b = 0; b < TW_NB && g * TW_NB + b < n_blocks; ++b) {
                const int d = d_lo + g * TW_NB + b;
                const int q_lo = -d > 0 ? -d : 0, q_hi = k.Q < n_xg - d ? k.Q : n_xg - d;
                c += q_hi > q_lo ? q_hi - q_lo : 0;
            }
            cost[g] = c > 2 * k.Q ? c : 2 * k.Q;
        }
        const int total = ok ? balance_group_ctas(cost, k.n_dgroups, k.n_units, 148, k.cta_start) : 0;
        k.balanced = total > 0 ? 1 : 0;
        if (k.balanced) n_ctas = total;
    }
    k.dw = dw;
    if (cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)TL_C * TL_C * p->kh, s) != cudaSuccess) return CPC_ERR_CUDA;
    if (cudaFuncSetAttribute(tall_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TW_SMEM) != cudaSuccess)
        return CPC_ERR_CUDA;
    tall_wgrad_kernel<<<n_ctas, TL_THREADS, TW_SMEM, s>>>(tx, tdy, k);
    CPC_LAUNCH_CHECK();
    count_launch(3);
    return CPC_OK;
}

}  // namespace cpc
