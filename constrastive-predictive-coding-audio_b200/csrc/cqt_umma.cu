// Constant-Q filterbank on the tcgen05 tensor cores, with the log-power / phase-difference epilogue fused.
// Replaces CQT.forward (constant_q_transform.py:161-172) + PreprocessingModule.forward
// (scalogram_model.py:75-102) for every octave group whose kernel is at least 128 taps long.
//
//   out[b, f, t] = sum_n x[b, off_g + hop*t + n] * W_g[f, n]          (correlation, real | imag channels)
//
// GEMM view per group: M = 128 frames of one item, N = 2*n_g channels (padded to NP), K = K_g taps.
// The audio is viewed as a matrix X2[s][r] = x[hop*s + r]; frame t, taps [hop*m + 64*h, +64) is row t + m of
// X2, column half h.  So ALL the A operands of a 128-frame tile are row-shifted windows of one shared-memory
// slab of 256 X2 rows, which is loaded ONCE per tile (128 KB with both planes): the K loop only moves the
// descriptor start address by m rows (SWIZZLE_128B K-major), while the filter chunks (B operand) stream through
// a TMA ring.  One tile accumulates every group of its "set" into separate TMEM columns, then the epilogue turns
// (re, im) into the scalogram in registers:
//   complex (B,F,T,2) | log-power (B,1,F,T) | log-power + unwrapped phase difference (B,2,F,T-1),
// with coalesced stores along t.  Tiles (set, item, frame block) are handed out through an atomic counter in
// descending cost order.
//
// fp32-exact arithmetic on 16-bit tensor-core operands.  Both operands are split into TWO fp16 planes after an exact
// power-of-two scaling that puts the largest magnitude of the item (audio) / of the bin (filter) at 2^14:
//     v * 2^s = hi + lo + r,   hi = fp16(v * 2^s),  lo = fp16(v * 2^s - hi),   |r| <= max(2^-23 |v * 2^s|, 2^-25)
// i.e. 22 mantissa bits wherever |v| >= 2^-17 of the item's / bin's peak and an absolute error of 2^-39 of the peak
// below that (bf16 hi/lo planes only carry 16 bits: the 1e-5 relative error of that scheme is what log / atan2 of
// near-silent cells amplified beyond the 1e-3 parity bound).  Products hi*hi, hi*lo, lo*hi are exact in the fp32
// accumulator; lo*lo (2^-22) is dropped; the epilogue multiplies by 2^-(s_item + s_bin).
// The three products of a K step are TWO MMAs: the filter planes are stacked along N, [W_hi | W_lo], so that
// X_hi * [W_hi | W_lo] fetches the audio window once for both (these MMAs are bound by the shared-memory operand
// fetch, not by the tensor pipe), then X_lo * W_hi accumulates onto the W_lo columns; the epilogue adds the two column
// blocks.  Short groups (K < 512, 2 % of the work) use three N = NP MMAs into one column block instead, which lets
// several of them share a tile.  Long groups accumulate in rounds of 16 chunks whose partial sums are added in registers
// (see CQ_ROUND_CHUNKS: the tensor core's accumulator truncates); rounds alternate between two TMEM buffers, which also
// overlaps the epilogue of a tile with the MMAs of the next.  64-tap chunks in which every filter of the group is zero (the centre-padding of
// constant_q_transform.py:132-140: 25 % of the longest group) are found while packing the filters and skipped.
#include "common.cuh"
#include "umma.cuh"

#include <cuda_fp16.h>

namespace cpc {
using namespace umma;

constexpr int CQ_THREADS = 384;                        // 4 control warps + 2 x 4 epilogue warps
constexpr int CQ_ACC_COLS = 256;                       // TMEM columns per accumulator buffer (two buffers)
constexpr int CQ_SLAB_ROWS = 256;
constexpr int CQ_SLAB_PLANE = CQ_SLAB_ROWS * 128;        // 32 KB: one column half, one plane
constexpr int CQ_BSTAGES = 5;
constexpr int CQ_MAX_SETS = 8;
constexpr int CQ_WIDE_MIN_K = 512;                     // groups at least this long stack [W_hi | W_lo] along N
// The tensor core adds into its fp32 accumulator with truncation (measured: every accumulating MMA shrinks the running sum
// by ~2^-26 of its magnitude, coherently, so the relative error of a group grows linearly with its number of K steps:
// 1.05e-5 after the 768 steps of the 16384-tap group against 1.8e-6 for an fp32 FMA chain -- and being a coherent shrink
// per bin it shifts the log-power scalogram by a per-bin constant, which hurts more than random noise of that size).
// Sets of at most two (wide) groups therefore accumulate in ROUNDS of CQ_ROUND_CHUNKS 64-tap chunks (16: measured 0.64 ms /
// 1.3e-6 on the longest group; 8: 0.70 ms / 7e-7; one round: 0.58 ms / 1.05e-5): a round starts a
// fresh accumulator in one of the two TMEM buffers, and while the tensor core works on the next round in the other
// buffer the epilogue warps pull the finished partial sums into registers and add them there (fp32, round to nearest).
constexpr int CQ_ROUND_CHUNKS = 16;

struct CqGroup {
    int K, off, n_g, bin_lo, n_chunks, w_row0;           // w_row0: first row of the group in the packed filters
    int wide, col;                                       // wide: [main | correction] column blocks; col: first TMEM column
};

// Per-call device bookkeeping (workspace): the tile scheduler's counter.
struct CqMeta {
    int counter;
};
// Lives at the end of the packed filter blob: written once by the filter packing kernels, read by every forward call.
struct CqFilterMeta {
    int live_lo[CPC_CQT_MAX_GROUPS];                     // first / one-past-last tap (in the group's padded tap axis) at
    int live_hi[CPC_CQT_MAX_GROUPS];                     //   which any filter of the group is non-zero
};

struct CqtUmma {
    int B, T, F, hop, halves, n_blocks, fpb;             // fpb: new frames per tile (127 in phase mode, else 128)
    int NP;                                              // padded channel count per group (multiple of 16, <= 64)
    int n_sets, set_first[CQ_MAX_SETS], set_count[CQ_MAX_SETS];
    int set_round[CQ_MAX_SETS];                          // chunks per group and round (CQ_ROUND_CHUNKS, or "all" = 1 << 20)
    int n_tiles;
    CqGroup g[CPC_CQT_MAX_GROUPS];
    int n_groups;
    int mode, To;
    int half;                                            // CPC_CQT_FLAG_HALF_OPERANDS: hi planes only, one product
    float eps, log_offset, norm, power;
    const float* phase_fixed;
    const float* phase_scale;
    const float* inv_item;                               // (B)  2^-s of the item's audio scaling
    const float* inv_bin;                                // (F)  2^-s of the bin's filter scaling
    float* out;
    CqMeta* meta;
    const CqFilterMeta* fmeta;
};

struct __align__(8) CqBarriers {
    uint64_t bfull[CQ_BSTAGES], bempty[CQ_BSTAGES], slab_full, slab_empty, acc_full[2], acc_empty[2], sfull[2], sempty[2];
    uint32_t tmem_base;
    int tile_id[2];
    int c_lo[CPC_CQT_MAX_GROUPS], c_hi[CPC_CQT_MAX_GROUPS];   // live 64-tap chunk range of every group
    float exch[2][2][4][32];                         // [epilogue warp set][parity][warp][bin]
};

// 2^s with s clamped so that the result and its reciprocal are normal floats
__device__ __forceinline__ float pow2i(int s) {
    s = s < -100 ? -100 : (s > 100 ? 100 : s);
    return __int_as_float((s + 127) << 23);
}
// power-of-two scale that puts a magnitude with float bits `bits` into [2^14, 2^15)
__device__ __forceinline__ int scale_exponent(uint32_t bits) { return 14 - ((int)((bits >> 23) & 0xff) - 127); }

__global__ void cqt_init_meta_kernel(CqMeta* meta, uint32_t* item_max, int B) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) meta->counter = 0;
    if (i < B) item_max[i] = 0u;
}
__global__ void cqt_init_filter_meta_kernel(CqFilterMeta* fm) {
    const int i = threadIdx.x;
    if (i < CPC_CQT_MAX_GROUPS) { fm->live_lo[i] = 0x7fffffff; fm->live_hi[i] = 0; }
}

// largest |x| of every item (float bits of non-negative values order like unsigned integers)
__global__ void __launch_bounds__(256) cqt_item_max_kernel(const float* __restrict__ x, uint32_t* __restrict__ item_max, int L,
                                                          int pitch) {
    const int b = blockIdx.y;
    const float* row = x + (size_t)b * pitch;
    uint32_t m = 0;
    const int stride = gridDim.x * blockDim.x;
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    if (((pitch & 3) == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0)) {
        const float4* row4 = reinterpret_cast<const float4*>(row);
        const int n4 = L >> 2;
        for (int q = tid; q < n4; q += stride) {
            const float4 v = __ldg(row4 + q);
            m = max(max(m, __float_as_uint(fabsf(v.x))), __float_as_uint(fabsf(v.y)));
            m = max(max(m, __float_as_uint(fabsf(v.z))), __float_as_uint(fabsf(v.w)));
        }
        const int tail = (n4 << 2) + tid;                  // the (< 4) trailing samples, by the first threads
        if (tail < L) m = max(m, __float_as_uint(fabsf(__ldg(row + tail))));
    } else {
        for (int i = tid; i < L; i += stride) m = max(m, __float_as_uint(fabsf(__ldg(row + i))));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m) atomicMax(item_max + b, m);
}

// x (B, pitch) fp32 -> fp16 [plane][b][S*hop] of x * 2^s_b (zeros past the item)
__global__ void __launch_bounds__(256) cqt_pack_audio_kernel(const float* __restrict__ x, __half* __restrict__ out,
                                                            const uint32_t* __restrict__ item_max, float* __restrict__ inv_item,
                                                            int B, int L, int pitch, int Lp) {
    const long groups = (long)B * (Lp >> 3);
    const long plane = (long)B * Lp;
    for (long gi = (long)blockIdx.x * blockDim.x + threadIdx.x; gi < groups; gi += (long)gridDim.x * blockDim.x) {
        const int b = (int)(gi / (Lp >> 3));
        const int i0 = (int)(gi - (long)b * (Lp >> 3)) << 3;
        const int se = scale_exponent(__ldg(item_max + b));
        const float scale = pow2i(se);
        if (i0 == 0) inv_item[b] = pow2i(-se);
        __align__(16) __half hi[8];
        __align__(16) __half lo[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float v = (i0 + i < L) ? __ldg(x + (size_t)b * pitch + i0 + i) * scale : 0.f;
            hi[i] = __float2half_rn(v);
            lo[i] = __float2half_rn(v - __half2float(hi[i]));
        }
        const long o = (long)b * Lp + i0;
        *reinterpret_cast<uint4*>(out + o) = *reinterpret_cast<const uint4*>(hi);
        *reinterpret_cast<uint4*>(out + plane + o) = *reinterpret_cast<const uint4*>(lo);
    }
}

// filters: group block (2*n_g, K_src) fp32 row-major [real bins; imag bins] -> fp16 rows of 64 taps:
//   row = w_row0 + chunk * 2*NP + plane * NP + n,   n < NP/2: real part of bin n, n >= NP/2: imaginary part of bin n - NP/2
// The group's tap axis is K >= K_src taps long with the source filters centred in it (`pad` zero taps on either side: a
// 64-tap group becomes a 128-tap group so that its window starts on a 64-sample boundary of the slab).
// One block per bin slot: pass 1 finds the bin's peak magnitude (-> its power-of-two scale) and its non-zero tap range,
// pass 2 writes the two planes of both rows.  Slots beyond the group's bins write zero rows.
__global__ void __launch_bounds__(256) cqt_pack_filters_kernel(const float* __restrict__ w, __half* __restrict__ out,
                                                              float* __restrict__ inv_bin, CqFilterMeta* fm, int g, int K,
                                                              int K_src, int pad, int ng, int NP, int w_row0, int bin_lo) {
    __shared__ uint32_t s_max;
    __shared__ int s_lo, s_hi;
    const int j = blockIdx.x;                              // bin slot, < NP / 2
    const bool real_bin = j < ng;
    if (threadIdx.x == 0) { s_max = 0u; s_lo = 0x7fffffff; s_hi = 0; }
    __syncthreads();
    if (real_bin) {
        uint32_t m = 0;
        int lo = 0x7fffffff, hi = 0;
        for (int k = threadIdx.x; k < K_src; k += blockDim.x) {
            const float re = __ldg(w + (size_t)j * K_src + k), im = __ldg(w + (size_t)(ng + j) * K_src + k);
            const uint32_t a = max(__float_as_uint(fabsf(re)), __float_as_uint(fabsf(im)));
            if (a) { m = max(m, a); lo = min(lo, k + pad); hi = max(hi, k + pad + 1); }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
            lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
            hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
        }
        if ((threadIdx.x & 31) == 0) { atomicMax(&s_max, m); atomicMin(&s_lo, lo); atomicMax(&s_hi, hi); }
    }
    __syncthreads();
    const int se = scale_exponent(s_max);
    const float scale = pow2i(se);
    if (real_bin && threadIdx.x == 0) {
        inv_bin[bin_lo + j] = pow2i(-se);
        if (s_hi > s_lo) { atomicMin(&fm->live_lo[g], s_lo); atomicMax(&fm->live_hi[g], s_hi); }
    }
    const int hp = NP >> 1;
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
        const int c = k >> 6, kk = k & 63;
        const int ks = k - pad;
        const bool live = real_bin && ks >= 0 && ks < K_src;
#pragma unroll
        for (int part = 0; part < 2; ++part) {
            const float v = live ? __ldg(w + (size_t)(part * ng + j) * K_src + ks) * scale : 0.f;
            const __half hi = __float2half_rn(v);
            const size_t row = (size_t)w_row0 + (size_t)c * 2 * NP + part * hp + j;
            out[row * 64 + kk] = hi;
            out[(row + NP) * 64 + kk] = __float2half_rn(v - __half2float(hi));
        }
    }
}

// Instruction descriptor for kind::f16 with FP16 A/B (format code 0) and F32 accumulate, both operands K-major.
__host__ __device__ inline uint32_t make_idesc_f16(int m, int n) {
    uint32_t d = 0;
    d |= 1u << 4;                                // D format F32
    d |= (uint32_t)(n >> 3) << 17;
    d |= (uint32_t)(m >> 4) << 24;
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
    // one-plane mode: the tensor maps' boxes cover the hi plane / the W_hi rows only (the lo halves of the slab and of a
    // ring stage stay unused)
    const uint32_t slab_tx = p.half ? (uint32_t)slab_bytes / 2 : (uint32_t)slab_bytes;
    const uint32_t b_tx = p.half ? (uint32_t)b_stage / 2 : (uint32_t)b_stage;
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
    if (warp == 3 && lane < p.n_groups) {
        // live chunk range of the group (at least one chunk, so that every accumulator column is written)
        const int n_chunks = p.g[lane].n_chunks;
        int lo = p.fmeta->live_lo[lane] >> 6, hi = (p.fmeta->live_hi[lane] + 63) >> 6;
        if (hi <= lo) { lo = 0; hi = 1; }
        lo = lo < 0 ? 0 : (lo > n_chunks - 1 ? n_chunks - 1 : lo);
        hi = hi > n_chunks ? n_chunks : (hi < lo + 1 ? lo + 1 : hi);
        bars->c_lo[lane] = lo;
        bars->c_hi[lane] = hi;
    }
    if (warp == 2) { tmem_alloc(&bars->tmem_base, 512); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;

    // rounds of a tile: every group of the set contributes its live chunks [c_lo + r*R, c_lo + (r+1)*R) to round r
    auto set_rounds = [&](int set) {
        const int R = p.set_round[set];
        int n = 1;
        for (int gi = 0; gi < p.set_count[set]; ++gi) {
            const int gidx = p.set_first[set] + gi;
            const int live = bars->c_hi[gidx] - bars->c_lo[gidx];
            n = max(n, (live + R - 1) / R);
        }
        return n;
    };

    if (warp == 0) {
        // ===== scheduler + TMA producer =====
        if (lane == 0) {
            uint32_t bn = 0, tn = 0;
            for (uint32_t i = 0;; ++i) {
                const int slot = i & 1;
                mbar_wait(&bars->sempty[slot], ((i >> 1) & 1) ^ 1);
                int tile = atomicAdd(&p.meta->counter, 1);
                if (tile >= p.n_tiles) tile = -1;
                bars->tile_id[slot] = tile;
                mbar_arrive(&bars->sfull[slot]);
                if (tile < 0) break;
                int set, b, blk;
                cq_tile(p, tile, set, b, blk);
                mbar_wait(&bars->slab_empty, (tn & 1) ^ 1);
                mbar_expect_tx(&bars->slab_full, slab_tx);
                for (int h = 0; h < p.halves; ++h)
                    tma_load_4d(slab + h * 2 * CQ_SLAB_PLANE, &tmap_x, &bars->slab_full, h * 64, blk * p.fpb, b, 0);
                ++tn;
                const int R = p.set_round[set], n_rounds = set_rounds(set);
                for (int r = 0; r < n_rounds; ++r)
                    for (int gi = 0; gi < p.set_count[set]; ++gi) {
                        const int gidx = p.set_first[set] + gi;
                        const CqGroup& g = p.g[gidx];
                        const int c0 = bars->c_lo[gidx] + r * R, c1 = min(bars->c_hi[gidx], c0 + R);
                        for (int c = c0; c < c1; ++c, ++bn) {
                            const int stage = bn % CQ_BSTAGES;
                            mbar_wait(&bars->bempty[stage], ((bn / CQ_BSTAGES) & 1) ^ 1);
                            mbar_expect_tx(&bars->bfull[stage], b_tx);
                            tma_load_3d(b_ring + stage * b_stage, &tmap_w, &bars->bfull[stage], 0, g.w_row0 + c * 2 * p.NP, 0);
                        }
                    }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: warp-uniform loop, one elected lane issues =====
        {
            const uint32_t idesc_np = make_idesc_f16(128, p.NP), idesc_2np = make_idesc_f16(128, 2 * p.NP);
            const uint32_t slab_addr = smem_u32(slab);
            const uint32_t hop_shift = p.hop == 64 ? 6 : 7;                  // hop is 64 or 128
            uint32_t bn = 0, tn = 0, rc = 0;                                 // rc: rounds issued so far (buffer = rc & 1)
            for (uint32_t i = 0;; ++i) {
                const int slot = i & 1;
                mbar_wait(&bars->sfull[slot], (i >> 1) & 1);
                const int tile = *reinterpret_cast<volatile int*>(&bars->tile_id[slot]);
                __syncwarp();
                if (lane == 0) mbar_arrive(&bars->sempty[slot]);
                if (tile < 0) break;
                int set, b, blk;
                cq_tile(p, tile, set, b, blk);
                const int R = p.set_round[set], n_rounds = set_rounds(set);
                for (int r = 0; r < n_rounds; ++r, ++rc) {
                    const uint32_t buf = rc & 1;
                    mbar_wait(&bars->acc_empty[buf], ((rc >> 1) & 1) ^ 1);
                    if (r == 0) mbar_wait(&bars->slab_full, tn & 1);
                    tc_fence_after();
                    for (int gi = 0; gi < p.set_count[set]; ++gi) {
                        const int gidx = p.set_first[set] + gi;
                        const CqGroup& g = p.g[gidx];
                        const uint32_t d_tmem = tmem_base + buf * CQ_ACC_COLS + (uint32_t)g.col;
                        const int c0 = bars->c_lo[gidx] + r * R, c1 = min(bars->c_hi[gidx], c0 + R);
                        for (int c = c0; c < c1; ++c, ++bn) {
                            const int stage = bn % CQ_BSTAGES;
                            mbar_wait(&bars->bfull[stage], (bn / CQ_BSTAGES) & 1);
                            tc_fence_after();
                            const uint32_t so = (uint32_t)(g.off + c * 64);    // sample offset of this tap chunk
                            const uint32_t m = so >> hop_shift, half = (so & (p.hop - 1)) >> 6;
                            const uint32_t a_hi = slab_addr + half * 2 * CQ_SLAB_PLANE + m * 128;
                            const uint32_t b_hi = smem_u32(b_ring + stage * b_stage);
                            if (elect_one()) {
                                // K-major SWIZZLE_128B; the audio window may start on any 128-byte row of the slab (the
                                // swizzle follows absolute address bits, so no descriptor base offset is needed)
                                const uint64_t a_d_hi = make_smem_desc(a_hi, 16, 1024), a_d_lo = make_smem_desc(a_hi + CQ_SLAB_PLANE, 16, 1024);
                                const uint64_t b_d_hi = make_smem_desc(b_hi, 16, 1024), b_d_lo = make_smem_desc(b_hi + p.NP * 128, 16, 1024);
                                const uint32_t first = (uint32_t)(c - c0);     // 0 on the first chunk of the round: overwrite
                                if (p.half) {
                                    // one-plane mode (bf16 operand mode of the model): X_hi * W_hi into the main block
#pragma unroll
                                    for (int k = 0; k < 4; ++k)
                                        mma_bf16(d_tmem, a_d_hi + (uint64_t)(2 * k), b_d_hi + (uint64_t)(2 * k), idesc_np, first | (uint32_t)k);
                                } else if (g.wide) {
#pragma unroll
                                    for (int k = 0; k < 4; ++k) {              // +32 B per K step = +2 in the address field
                                        // X_hi * [W_hi | W_lo] -> both column blocks; X_lo * W_hi -> onto the second block
                                        mma_bf16(d_tmem, a_d_hi + (uint64_t)(2 * k), b_d_hi + (uint64_t)(2 * k), idesc_2np, first | (uint32_t)k);
                                        mma_bf16(d_tmem + (uint32_t)p.NP, a_d_lo + (uint64_t)(2 * k), b_d_hi + (uint64_t)(2 * k), idesc_np, 1u);
                                    }
                                } else {
#pragma unroll
                                    for (int cb = 0; cb < 3; ++cb) {           // (hi,hi) (hi,lo) (lo,hi) into one column block
                                        const uint64_t ad = cb == 2 ? a_d_lo : a_d_hi, bd = cb == 1 ? b_d_lo : b_d_hi;
#pragma unroll
                                        for (int k = 0; k < 4; ++k)
                                            mma_bf16(d_tmem, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc_np, first | (uint32_t)(cb | k));
                                    }
                                }
                                tc_commit(&bars->bempty[stage]);
                            }
                            __syncwarp();
                        }
                    }
                    if (elect_one()) {
                        if (r == n_rounds - 1) tc_commit(&bars->slab_empty);
                        tc_commit(&bars->acc_full[buf]);
                    }
                    __syncwarp();
                }
                ++tn;
            }
        }
    } else if (warp >= 4) {
        // ===== epilogue: partial sums of every round -> registers; after the last round (re, im) -> output format =====
        const int ew = warp & 3;                                               // TMEM lane quarter
        const int eset = (warp - 4) >> 2;                                      // warp set 0 / 1: even / odd groups
        const int row = ew * 32 + lane;
        const float kPi = 3.14159265358979323846f;
        uint32_t rc = 0, xb = 0;

        // adds 32 accumulator columns into dst (assigns when `assign`)
        auto pull_part = [&](uint32_t col, uint32_t (&dst)[32], bool assign) {
            if (assign) {
                tmem_ld32(col, dst);
                tmem_ld_wait();
            } else {
                uint32_t cr[32];
                tmem_ld32(col, cr);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j) dst[j] = __float_as_uint(__uint_as_float(dst[j]) + __uint_as_float(cr[j]));
            }
        };
        // adds the accumulator block(s) of group g in buffer `buf` into re / im (assigns when `fresh`)
        auto pull = [&](const CqGroup& g, uint32_t buf, int j0, uint32_t (&re)[32], uint32_t (&im)[32], bool fresh) {
            const uint32_t base = tmem_base + ((uint32_t)(ew * 32) << 16) + buf * CQ_ACC_COLS + (uint32_t)(g.col + j0);
            pull_part(base, re, fresh);                                       // main block: real | imaginary columns
            pull_part(base + (uint32_t)(p.NP >> 1), im, fresh);
            if (g.wide && !p.half) {                                          // the hi*lo + lo*hi correction block
                pull_part(base + (uint32_t)p.NP, re, false);
                pull_part(base + (uint32_t)(p.NP + (p.NP >> 1)), im, false);
            }
        };

        // (re, im) of bins [g.bin_lo + j0, +32) at this thread's frame -> output
        auto emit = [&](const CqGroup& g, int j0, uint32_t (&re)[32], uint32_t (&im)[32], int b, int t, float inv_item) {
            const int nb = min(32, g.n_g - j0);
            // undo the two power-of-two scalings (exact)
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const float sc = __ldg(p.inv_bin + g.bin_lo + j0 + (j < nb ? j : 0));
                re[j] = __float_as_uint(__uint_as_float(re[j]) * inv_item * sc);
                im[j] = __float_as_uint(__uint_as_float(im[j]) * inv_item * sc);
            }
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
                            float amp = (logf(fmaf(x, x, y * y) + p.eps) + p.log_offset) * p.norm;
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
                const bool put = row >= 1 && t < p.T;
                const int to = t - 1;
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    float prev = __shfl_up_sync(0xffffffffu, ph[j], 1);
                    if (lane == 0 && ew > 0) prev = ex[(ew - 1) * 32 + j];
                    if (put && j < nb) {
                        const int f = g.bin_lo + j0 + j;
                        const float x = __uint_as_float(re[j]), y = __uint_as_float(im[j]);
                        float amp = (logf(fmaf(x, x, y * y) + p.eps) + p.log_offset) * p.norm;
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
        };

        for (uint32_t i = 0;; ++i) {
            const int slot = i & 1;
            mbar_wait(&bars->sfull[slot], (i >> 1) & 1);
            const int tile = *reinterpret_cast<volatile int*>(&bars->tile_id[slot]);
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars->sempty[slot]);
            if (tile < 0) break;
            int set, b, blk;
            cq_tile(p, tile, set, b, blk);
            const int t = blk * p.fpb + row;                                   // frame of this thread
            const float inv_item = __ldg(p.inv_item + b);
            const int R = p.set_round[set], n_rounds = set_rounds(set);
            if (n_rounds == 1) {
                // the whole tap range sits in one buffer: any number of groups, handled alternately by the two warp sets
                const uint32_t buf = rc & 1;
                mbar_wait(&bars->acc_full[buf], (rc >> 1) & 1);
                tc_fence_after();
                for (int gi = eset; gi < p.set_count[set]; gi += 2) {
                    const CqGroup& g = p.g[p.set_first[set] + gi];
                    for (int j0 = 0; j0 < g.n_g; j0 += 32) {
                        uint32_t re[32], im[32];
                        pull(g, buf, j0, re, im, true);
                        emit(g, j0, re, im, b, t, inv_item);
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bars->acc_empty[buf]);
                ++rc;
            } else {
                // several rounds (host side: such sets hold at most two groups of at most 32 bins, one per warp set): the
                // partial sums of a round are added in registers and its buffer is released at once
                const bool mine = eset < p.set_count[set];
                const int gidx = p.set_first[set] + (mine ? eset : 0);
                const CqGroup& g = p.g[gidx];
                const int c_lo = bars->c_lo[gidx], c_hi = bars->c_hi[gidx];
                uint32_t re[32], im[32];
                bool fresh = true;
                for (int r = 0; r < n_rounds; ++r, ++rc) {
                    const uint32_t buf = rc & 1;
                    mbar_wait(&bars->acc_full[buf], (rc >> 1) & 1);
                    tc_fence_after();
                    if (mine && c_lo + r * R < c_hi) {
                        pull(g, buf, 0, re, im, fresh);
                        fresh = false;
                    }
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&bars->acc_empty[buf]);
                }
                if (mine) emit(g, 0, re, im, b, t, inv_item);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// ---- host side --------------------------------------------------------------------------------------
struct CqtUmmaPlan {
    bool ok;
    int n_tensor_groups;          // groups [0, n) run here, the rest on the CUDA-core kernel
    int NP, halves, S, Lp;
    size_t audio_bytes, filter_bytes, fmeta_bytes, meta_bytes, total_rows;
};

// tap count of group g on the tensor path: groups shorter than 128 taps are zero-padded (centred) to 128
static inline int cq_eff_k(int k) { return k < 128 ? 128 : k; }

static CqtUmmaPlan cqt_umma_plan(const cpc_cqt_params* p) {
    CqtUmmaPlan u{};
    u.ok = false;
    if (p->hop != 64 && p->hop != 128) return u;
    if (p->pool_t != 1 && p->mode != CPC_CQT_COMPLEX) return u;       // pooled modes keep the two-kernel path
    const int k0 = p->kernel_size[0];
    if (k0 < 128 || k0 / p->hop > 128) return u;                      // slab holds 128 frames + 127 row shifts
    int n = 0, np = 16;
    while (n < p->n_groups) {
        const int ke = cq_eff_k(p->kernel_size[n]);
        // the group's window must start on a 64-sample boundary of the slab and consist of whole 64-tap chunks
        if (ke > k0 || (ke & 63) || (((k0 - ke) / 2) & 63) || ((ke - p->kernel_size[n]) & 1)) break;
        const int n2 = 2 * (p->bin_hi[n] - p->bin_lo[n]);
        if (n2 > 64) break;
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
    for (int g = 0; g < n; ++g) rows += (size_t)(cq_eff_k(p->kernel_size[g]) / 64) * 2 * u.NP;
    u.total_rows = rows;
    u.filter_bytes = align_up(rows * 128, 1024);
    u.fmeta_bytes = align_up(sizeof(CqFilterMeta) + (size_t)p->n_bins * 4, 1024);     // CqFilterMeta | inv_bin (F f32)
    u.meta_bytes = align_up(sizeof(CqMeta) + 16 + (size_t)p->batch * 8, 1024);        // CqMeta | item_max | inv_item
    u.ok = true;
    return u;
}

bool cqt_umma_eligible(const cpc_cqt_params* p) { return cqt_umma_plan(p).ok; }
int cqt_umma_tensor_groups(const cpc_cqt_params* p) { return cqt_umma_plan(p).n_tensor_groups; }
size_t cqt_umma_packed_filter_bytes(const cpc_cqt_params* p) {
    CqtUmmaPlan u = cqt_umma_plan(p);
    return u.ok ? u.filter_bytes + u.fmeta_bytes + 1024 : 0;
}
size_t cqt_umma_workspace(const cpc_cqt_params* p) {
    CqtUmmaPlan u = cqt_umma_plan(p);
    return u.ok ? u.audio_bytes + u.meta_bytes + 1024 + cqt_umma_packed_filter_bytes(p) : 0;
}

static void cq_fill_groups(const cpc_cqt_params* p, const CqtUmmaPlan& u, CqtUmma& k) {
    int row0 = 0;
    for (int g = 0; g < u.n_tensor_groups; ++g) {
        CqGroup& q = k.g[g];
        q.K = cq_eff_k(p->kernel_size[g]);
        q.off = (p->kernel_size[0] - q.K) / 2;
        q.n_g = p->bin_hi[g] - p->bin_lo[g];
        q.bin_lo = p->bin_lo[g];
        q.n_chunks = q.K / 64;
        q.w_row0 = row0;
        q.wide = (q.K >= CQ_WIDE_MIN_K && 2 * u.NP <= CQ_ACC_COLS) ? 1 : 0;
        row0 += q.n_chunks * 2 * u.NP;
    }
}

// Packs the filterbank of groups [0, n_tensor_groups) into `packed` (cqt_umma_packed_filter_bytes): fp16 hi / lo rows,
// per-bin scales and the live tap range of every group.  The result only depends on the weights: callers keep it.
int cqt_umma_pack_filters(const float* weights, void* packed, const cpc_cqt_params* p, cudaStream_t s) {
    CqtUmmaPlan u = cqt_umma_plan(p);
    if (!u.ok) return CPC_ERR_UNSUPPORTED;
    uint8_t* pk = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(packed) + 1023) & ~(uintptr_t)1023);
    __half* wp = reinterpret_cast<__half*>(pk);
    CqFilterMeta* fm = reinterpret_cast<CqFilterMeta*>(pk + u.filter_bytes);
    float* inv_bin = reinterpret_cast<float*>(pk + u.filter_bytes + sizeof(CqFilterMeta));
    CqtUmma k{};
    cq_fill_groups(p, u, k);
    cqt_init_filter_meta_kernel<<<1, 32, 0, s>>>(fm);
    CPC_LAUNCH_CHECK();
    for (int g = 0; g < u.n_tensor_groups; ++g) {
        const CqGroup& q = k.g[g];
        const int k_src = p->kernel_size[g];
        cqt_pack_filters_kernel<<<u.NP / 2, 256, 0, s>>>(weights + p->weight_offset[g], wp, inv_bin, fm, g, q.K, k_src,
                                                        (q.K - k_src) / 2, q.n_g, u.NP, q.w_row0, q.bin_lo);
        CPC_LAUNCH_CHECK();
    }
    count_launch(1 + u.n_tensor_groups);
    return CPC_OK;
}

// Computes groups [0, n_tensor_groups) straight into `out` in the requested mode.  `packed_filters` (optional): the
// caller's copy made by cqt_umma_pack_filters; without it the filters are packed into the workspace on every call.
int cqt_umma_launch(const float* x, const float* weights, const void* packed_filters, const float* phase_fixed,
                    const float* phase_scale, float* out, const cpc_cqt_params* p, void* workspace, size_t workspace_bytes,
                    cudaStream_t s) {
    CqtUmmaPlan u = cqt_umma_plan(p);
    if (!u.ok) return CPC_ERR_UNSUPPORTED;
    if (!workspace || workspace_bytes < cqt_umma_workspace(p)) return CPC_ERR_WORKSPACE;
    uint8_t* ws = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(workspace) + 1023) & ~(uintptr_t)1023);
    __half* xp = reinterpret_cast<__half*>(ws);
    uint8_t* mb = ws + u.audio_bytes;
    CqMeta* meta = reinterpret_cast<CqMeta*>(mb);
    uint32_t* item_max = reinterpret_cast<uint32_t*>(mb + 16);
    float* inv_item = reinterpret_cast<float*>(item_max + p->batch);
    if (!packed_filters) {
        void* own = ws + u.audio_bytes + u.meta_bytes;
        const int st = cqt_umma_pack_filters(weights, own, p, s);
        if (st != CPC_OK) return st;
        packed_filters = own;
    }
    const uint8_t* pk = reinterpret_cast<const uint8_t*>((reinterpret_cast<uintptr_t>(packed_filters) + 1023) & ~(uintptr_t)1023);
    const __half* wp = reinterpret_cast<const __half*>(pk);
    const CqFilterMeta* fm = reinterpret_cast<const CqFilterMeta*>(pk + u.filter_bytes);
    const float* inv_bin = reinterpret_cast<const float*>(pk + u.filter_bytes + sizeof(CqFilterMeta));
    {
        cqt_init_meta_kernel<<<ceil_div(p->batch, 256), 256, 0, s>>>(meta, item_max, p->batch);
        CPC_LAUNCH_CHECK();
        int bx = ceil_div(p->n_samples, 256 * 4 * 8);                 // ~8 float4 loads per thread
        if (bx < 1) bx = 1;
        cqt_item_max_kernel<<<dim3(bx, p->batch), 256, 0, s>>>(x, item_max, p->n_samples, p->x_pitch);
        CPC_LAUNCH_CHECK();
        const long groups = (long)p->batch * (u.Lp / 8);
        int blocks = (int)((groups + 255) / 256);
        if (blocks > 148 * 8) blocks = 148 * 8;
        cqt_pack_audio_kernel<<<blocks, 256, 0, s>>>(x, xp, item_max, inv_item, p->batch, p->n_samples, p->x_pitch, u.Lp);
        CPC_LAUNCH_CHECK();
    }
    CqtUmma k{};
    cq_fill_groups(p, u, k);
    // sets: consecutive groups whose accumulator columns fit one TMEM buffer, greedily balanced towards equal tap counts.
    // Sets of one or two wide groups accumulate in rounds (one group per epilogue warp set keeps its partial sums in
    // registers); a wide group never shares a set with more than one other group.
    {
        long total = 0;
        for (int g = 0; g < u.n_tensor_groups; ++g) total += k.g[g].K;
        const long target = (total + 2) / 3;
        int g = 0;
        k.n_sets = 0;
        while (g < u.n_tensor_groups) {
            int cnt = 0, cols = 0;
            long acc = 0;
            bool any_wide = false;
            const bool last_slot = k.n_sets == CQ_MAX_SETS - 1;
            while (g + cnt < u.n_tensor_groups) {
                const CqGroup& q = k.g[g + cnt];
                const int need = q.wide ? 2 * u.NP : u.NP;
                if (cols + need > CQ_ACC_COLS) break;
                if (cnt > 0 && !last_slot && acc + q.K > target) break;
                if ((any_wide || q.wide) && cnt >= 2) break;               // rounds: at most two groups
                if (any_wide && !q.wide) break;                            // keep round sets homogeneous
                k.g[g + cnt].col = cols;
                cols += need;
                acc += q.K;
                any_wide = any_wide || q.wide;
                ++cnt;
            }
            if (cnt == 0 || (last_slot && g + cnt < u.n_tensor_groups)) return CPC_ERR_UNSUPPORTED;
            k.set_first[k.n_sets] = g;
            k.set_count[k.n_sets] = cnt;
            const int round_override = (p->flags >> 8) & 0xff;              // experiments: flags bits 8..15 = chunks per round
            k.set_round[k.n_sets] = any_wide ? (round_override ? round_override : CQ_ROUND_CHUNKS) : (1 << 20);
            ++k.n_sets;
            g += cnt;
        }
    }
    CUtensorMap tx, tw;
    {
        // 16-bit elements: the tensor maps only move bytes, so the bf16 encoder serves the fp16 planes
        const uint64_t dims[4] = {(uint64_t)p->hop, (uint64_t)u.S, (uint64_t)p->batch, 2};
        const uint64_t strides[3] = {(uint64_t)p->hop * 2, (uint64_t)u.Lp * 2, (uint64_t)u.Lp * 2 * p->batch};
        const uint32_t half = (p->flags & CPC_CQT_FLAG_HALF_OPERANDS) ? 1u : 0u;
        const uint32_t box[4] = {64, CQ_SLAB_ROWS, 1, half ? 1u : 2u};
        if (!make_tmap_bf16(&tx, xp, 4, dims, strides, box)) return CPC_ERR_CUDA;
        const uint64_t wd[3] = {64, (uint64_t)u.total_rows, 1};
        const uint64_t wsr[2] = {128, 128 * (uint64_t)u.total_rows};
        const uint32_t wbox[3] = {64, (uint32_t)((half ? 1 : 2) * u.NP), 1};       // [W_hi | W_lo] rows of a chunk, or W_hi
        if (!make_tmap_bf16(&tw, wp, 3, wd, wsr, wbox)) return CPC_ERR_CUDA;
    }
    k.B = p->batch; k.T = p->n_frames; k.F = p->n_bins; k.hop = p->hop; k.halves = u.halves; k.NP = u.NP;
    k.n_groups = u.n_tensor_groups;
    k.mode = p->mode;
    k.half = (p->flags & CPC_CQT_FLAG_HALF_OPERANDS) ? 1 : 0;
    k.fpb = p->mode == CPC_CQT_LOGPOW_PHASE ? 127 : 128;
    const int t_out = p->mode == CPC_CQT_LOGPOW_PHASE ? p->n_frames - 1 : p->n_frames;
    k.To = t_out;
    k.n_blocks = ceil_div(t_out, k.fpb);
    k.n_tiles = k.n_sets * k.B * k.n_blocks;
    k.eps = p->eps; k.log_offset = p->log_offset; k.norm = p->norm; k.power = p->power;
    k.phase_fixed = phase_fixed; k.phase_scale = phase_scale; k.out = out; k.meta = meta; k.fmeta = fm;
    k.inv_item = inv_item; k.inv_bin = inv_bin;
    const size_t smem_bytes = (size_t)u.halves * 2 * CQ_SLAB_PLANE + (size_t)CQ_BSTAGES * 2 * u.NP * 128 + sizeof(CqBarriers) + 1024;
    if (cudaFuncSetAttribute(cqt_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes) != cudaSuccess)
        return CPC_ERR_CUDA;
    const int grid = k.n_tiles < 148 ? k.n_tiles : 148;
    cqt_umma_kernel<<<grid, CQ_THREADS, smem_bytes, s>>>(tx, tw, k);
    CPC_LAUNCH_CHECK();
    count_launch(4);
    return CPC_OK;
}

}  // namespace cpc
