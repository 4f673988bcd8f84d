// Strided convolution forward / data-gradient / weight-gradient as implicit GEMMs on the 5th-gen tensor cores.
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
// * A TMA box must start on a 16-byte boundary of the innermost (w) dimension, so everything horizontal
//   -- tap offset j, stride s_w, left padding -- is baked into "replicas" by the packing pass:
//   replica_r[.., w'] = src[.., w_mul*w' + rep_mul*r + w_off].  Vertical taps/strides are plain TMA row
//   coordinates  h = h_mul*oh + tap_h_mul*i + h_off.
// * The data gradient runs the same kernel on dy with swapped channel roles, once per output parity class
//   (h+pt mod s_h, w+pl mod s_w): inside a class it is a stride-1 correlation with the sub-kernel
//   w[:, :, rh + s_h*i', rw + s_w*j'], and the epilogue scatters to h = s_h*a + rh - pt.
#include <cstdlib>
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
    int n_atoms, PH, PW, AW;             // GEMM pixel domain (PH x PW per item; atoms = 64-pixel runs along w)
    int n_rows_out;                      // output channels (GEMM N total)
    int n_tile, n_ntiles;
    int th, tw, ntaps, tpc, cin_eff, cin_chunks, n_chunks;   // tap grid th x tw
    int h_mul, tap_h_mul, h_off;         // source row = h_mul*ph + tap_h_mul*i + h_off
    int out_H, out_W, oh_mul, oh_off, ow_mul, ow_off;        // output pixel = (oh_mul*ph + oh_off, ow_mul*pw + ow_off)
    int planes;                          // 2 = fp32-faithful split, 1 = bf16
    int nrep;                            // replicas per plane in the packed operand (>= tw)
    int srcH;                            // rows of the source tensor (taps landing outside are all-zero)
    int w_resident;                      // 1: all packed weight chunks live in shared memory for the whole kernel
    int fuse;                            // > 0: columns are (parity class, channel) with `fuse` channels per class (32)
    int relu, stages;
    const float* bias;
    float* y;
};

// ---- packing kernels --------------------------------------------------------------------------------
// fp32 (rows, W) -> bf16 (planes, nrep, rows, Wp); Wp % 8 == 0; one thread per 8 output columns.
// replica r, column w' holds src[w_mul*w' + rep_mul*r + w_off] (zero outside [0, W)).
__global__ void __launch_bounds__(256) pack_split_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out,
                                                        long rows, int W, int Wp, int planes, int nrep, int w_mul,
                                                        int rep_mul, int w_off, int vec_ok) {
    const int groups = Wp >> 3;
    const long total = rows * groups;
    const long plane_stride = (long)nrep * rows * Wp;
    for (long g = (long)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (long)gridDim.x * blockDim.x) {
        const long row = g / groups;
        const int w0 = (int)(g - row * groups) << 3;
        const float* src = x + row * W;
        if (w_mul == 2 && rep_mul == 1 && nrep == 3) {
            // 3x3 / stride-2 forward operand: the three replicas of 8 output columns read ONE window of 17 consecutive
            // source columns (replica r, column i <- window[2 i + r]): load and split it once instead of three strided
            // gathers (the generic loop below ran at 3.5 TB/s against 5.7 for the plain pack)
            const int wb = 2 * w0 + w_off;
            if (wb >= 0 && wb + 17 <= W) {
                float win[17];
                if (vec_ok && !(wb & 1)) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float2 t = __ldg(reinterpret_cast<const float2*>(src + wb) + i);
                        win[2 * i] = t.x; win[2 * i + 1] = t.y;
                    }
                    win[16] = __ldg(src + wb + 16);
                } else {
#pragma unroll
                    for (int i = 0; i < 17; ++i) win[i] = __ldg(src + wb + i);
                }
                __nv_bfloat16 whi[17], wlo[17];
#pragma unroll
                for (int i = 0; i < 17; ++i) {
                    whi[i] = __float2bfloat16_rn(win[i]);
                    wlo[i] = __float2bfloat16_rn(win[i] - __bfloat162float(whi[i]));
                }
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    __align__(16) __nv_bfloat16 hi[8];
                    __align__(16) __nv_bfloat16 lo[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) { hi[i] = whi[2 * i + r]; lo[i] = wlo[2 * i + r]; }
                    const long o = ((long)r * rows + row) * Wp + w0;
                    *reinterpret_cast<uint4*>(out + o) = *reinterpret_cast<const uint4*>(hi);
                    if (planes == 2) *reinterpret_cast<uint4*>(out + plane_stride + o) = *reinterpret_cast<const uint4*>(lo);
                }
                continue;
            }
        }
        for (int r = 0; r < nrep; ++r) {
            __align__(16) __nv_bfloat16 hi[8];
            __align__(16) __nv_bfloat16 lo[8];
            const int wb = w_mul * w0 + rep_mul * r + w_off;             // source column of output column w0
            float v[8];
            if (vec_ok && w_mul == 1 && !(wb & 1) && wb >= 0 && wb + 8 <= W) {
                // unit stride, 8-byte aligned, fully inside the row: four 8-byte loads (half the L1 wavefronts)
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float2 t = __ldg(reinterpret_cast<const float2*>(src + wb) + i);
                    v[2 * i] = t.x; v[2 * i + 1] = t.y;
                }
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int w = wb + w_mul * i;
                    v[i] = (w >= 0 && w < W) ? __ldg(src + w) : 0.f;
                }
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                hi[i] = __float2bfloat16_rn(v[i]);
                lo[i] = __float2bfloat16_rn(v[i] - __bfloat162float(hi[i]));
            }
            const long o = ((long)r * rows + row) * Wp + w0;
            *reinterpret_cast<uint4*>(out + o) = *reinterpret_cast<const uint4*>(hi);
            if (planes == 2) *reinterpret_cast<uint4*>(out + plane_stride + o) = *reinterpret_cast<const uint4*>(lo);
        }
    }
}

// Plain pack of dy (one replica, unit stride) that also reduces the bias gradient: rows are (item, channel, h), so
// sums[channel] += sum of the row.  dy is read from HBM exactly once for both results (the separate bias-gradient pass
// re-read every dy: 0.32 ms of the e24 step).  Per loop iteration a block covers 256 consecutive 8-column groups, i.e. a
// few consecutive rows: per-row bins in shared memory, then one thread folds the bins channel by channel.
constexpr int PACK_BINS = 264;
__global__ void __launch_bounds__(256) pack_split_sum_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out,
                                                            float* __restrict__ sums, long rows, int W, int Wp, int planes,
                                                            FastDiv d_h, FastDiv d_groups, int C, int vec_ok, int nrep) {
    // bins[parity][plane - first plane of the block]: a block's 256 consecutive groups span a few consecutive rows, i.e.
    // one or two (item, channel) planes; lanes of one plane are a contiguous lane range, summed by a segmented warp scan
    __shared__ float bins[2][PACK_BINS];
    const int groups = Wp >> 3, lane = threadIdx.x & 31;
    const long total = rows * groups;
    const long rep_stride = rows * (long)Wp;                      // replica 1 (column w' = dy[w' - 1]) follows replica 0
    const long plane_stride = nrep * rep_stride;
    for (int t = threadIdx.x; t < 2 * PACK_BINS; t += 256) (&bins[0][0])[t] = 0.f;
    __syncthreads();
    const long n_iter = (total + (long)gridDim.x * 256 - 1) / ((long)gridDim.x * 256);
    for (long it = 0; it < n_iter; ++it) {
        float* bin = bins[it & 1];
        const long g0 = (it * gridDim.x + blockIdx.x) * 256;          // first group of this block in this iteration
        const long g = g0 + threadIdx.x;
        const int plane_first = d_h.div(d_groups.div((int)min(g0, total - 1)));
        int plane = -1;
        float sum = 0.f;
        if (g < total) {
            int row, gi;
            d_groups.divmod((int)g, row, gi);
            const int w0 = gi << 3;
            plane = d_h.div(row);
            const float* src = x + (long)row * W;
            float v[8];
            if (vec_ok && w0 + 8 <= W) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float2 t = __ldg(reinterpret_cast<const float2*>(src + w0) + i);
                    v[2 * i] = t.x; v[2 * i + 1] = t.y;
                }
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = (w0 + i < W) ? __ldg(src + w0 + i) : 0.f;
            }
            __align__(16) __nv_bfloat16 hi[8];
            __align__(16) __nv_bfloat16 lo[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                sum += v[i];
                hi[i] = __float2bfloat16_rn(v[i]);
                lo[i] = __float2bfloat16_rn(v[i] - __bfloat162float(hi[i]));
            }
            const long o = (long)row * Wp + w0;
            *reinterpret_cast<uint4*>(out + o) = *reinterpret_cast<const uint4*>(hi);
            if (planes == 2) *reinterpret_cast<uint4*>(out + plane_stride + o) = *reinterpret_cast<const uint4*>(lo);
            if (nrep == 2) {
                // the shifted replica of the same 8 columns: {dy[w0 - 1], v[0..6]}
                const float vp = (w0 > 0 && w0 - 1 < W) ? __ldg(src + w0 - 1) : 0.f;
                __align__(16) __nv_bfloat16 h1[8];
                __align__(16) __nv_bfloat16 l1[8];
                h1[0] = __float2bfloat16_rn(vp);
                l1[0] = __float2bfloat16_rn(vp - __bfloat162float(h1[0]));
#pragma unroll
                for (int i = 1; i < 8; ++i) { h1[i] = hi[i - 1]; l1[i] = lo[i - 1]; }
                *reinterpret_cast<uint4*>(out + rep_stride + o) = *reinterpret_cast<const uint4*>(h1);
                if (planes == 2) *reinterpret_cast<uint4*>(out + plane_stride + rep_stride + o) = *reinterpret_cast<const uint4*>(l1);
            }
        }
        // segmented inclusive scan over lanes with equal plane (planes are non-decreasing across lanes)
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const float t = __shfl_up_sync(0xffffffffu, sum, o);
            const int q = __shfl_up_sync(0xffffffffu, plane, o);
            if (lane >= o && q == plane) sum += t;
        }
        const int next = __shfl_down_sync(0xffffffffu, plane, 1);
        if (plane >= 0 && (lane == 31 || next != plane)) atomicAdd(&bin[plane - plane_first], sum);
        __syncthreads();
        if (g0 < total) {
            const int plane_last = d_h.div(d_groups.div((int)min(g0 + 255, total - 1)));
            for (int t = threadIdx.x; t <= plane_last - plane_first; t += 256) {
                atomicAdd(sums + (plane_first + t) % C, bin[t]);
                bin[t] = 0.f;                                    // ready for iteration it + 2 (a barrier lies in between)
            }
        }
    }
}

int pack_split_sum_launch(const float* x, __nv_bfloat16* out, float* sums, long rows, int W, int Wp, int planes, int H, int C,
                          int nrep, cudaStream_t s) {
    if (cudaMemsetAsync(sums, 0, sizeof(float) * (size_t)C, s) != cudaSuccess) return CPC_ERR_CUDA;
    const long groups = rows * (Wp / 8);
    int blocks = (int)((groups + 255) / 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (blocks < 1) blocks = 1;
    const int vec_ok = (W % 2 == 0) && (reinterpret_cast<uintptr_t>(x) % 8 == 0);
    if (groups >= (1l << 31)) return CPC_ERR_BAD_SHAPE;
    pack_split_sum_kernel<<<blocks, 256, 0, s>>>(x, out, sums, rows, W, Wp, planes, FastDiv(H), FastDiv(Wp / 8), C, vec_ok, nrep);
    return cudaGetLastError() == cudaSuccess ? CPC_OK : CPC_ERR_CUDA;
}

// host launcher shared with conv_tall.cu / conv_tall128.cu
int pack_split_launch(const float* x, __nv_bfloat16* out, long rows, int W, int Wp, int planes, int nrep, int w_mul,
                      int rep_mul, int w_off, cudaStream_t s) {
    const long groups = rows * (Wp / 8);
    int blocks = (int)((groups + 255) / 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (blocks < 1) blocks = 1;
    // float2 loads need every row start 8-byte aligned
    const int vec_ok = (W % 2 == 0) && (reinterpret_cast<uintptr_t>(x) % 8 == 0);
    pack_split_kernel<<<blocks, 256, 0, s>>>(x, out, rows, W, Wp, planes, nrep, w_mul, rep_mul, w_off, vec_ok);
    return cudaGetLastError() == cudaSuccess ? CPC_OK : CPC_ERR_CUDA;
}

// weights (Cout, Cin, kh, kw) fp32 -> bf16 (planes, n_chunks, Nrows, 64), K-major rows of one K chunk.
// GEMM tap (i', j') of a (th x tw) tap grid reads source tap (i_off + i_mul*i', j_off + j_mul*j').
// swap = 0: rows n = co, k channel = ci (forward);  swap = 1: rows n = ci, k channel = co (data gradient).
// fuse = c > 0 (data gradient of a stride-2 conv, all four parity classes at once): rows n = (class, ci) with c
// input channels per class, class = 2*rh + rw reads source tap (rh + 2*i', rw + 2*j'); taps outside the kernel are zero.
struct TapMap { int th, tw, i_off, i_mul, j_off, j_mul, swap, fuse; };

__global__ void __launch_bounds__(256) pack_weights_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out,
                                                          int Cin, int kh, int kw, int n_rows, int k_ch, int cin_eff,
                                                          int cin_chunks, int tpc, int n_chunks, int planes, TapMap tm) {
    const long total = (long)n_chunks * n_rows * KCHUNK;
    const int ntaps = tm.th * tm.tw;
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
            const int ti = tap / tm.tw, tj = tap - ti * tm.tw;
            if (tm.fuse) {
                const int cls = n / tm.fuse, ci = n - cls * tm.fuse;
                const int i = (cls >> 1) + 2 * ti, j = (cls & 1) + 2 * tj;
                if (i < kh && j < kw) v = __ldg(w + (((size_t)c * Cin + ci) * kh + i) * kw + j);
            } else {
                const int i = tm.i_off + tm.i_mul * ti, j = tm.j_off + tm.j_mul * tj;
                const size_t co = tm.swap ? c : n, ci = tm.swap ? n : c;
                v = __ldg(w + ((co * Cin + ci) * kh + i) * kw + j);
            }
        }
        const __nv_bfloat16 hi = __float2bfloat16_rn(v);
        out[idx] = hi;
        if (planes == 2) out[total + idx] = __float2bfloat16_rn(v - __bfloat162float(hi));
    }
}

// ---- the GEMM kernel ----------------------------------------------------------------------------------
// A K chunk contributes nothing when every tap in it reads a source row outside [0, srcH) for both atoms of the
// tile (vertical zero padding, or the empty parts of a data gradient).  The source row of vertical tap i is
// row + i * tap_h_mul with tap_h_mul = +-1, so the live taps of an atom are ONE contiguous range of i: the range is
// computed once per tile (ncu: testing the 60 chunks of the 15x1 data gradient one by one cost the single producer /
// issuer warps 40 k cycles per tile for 8 live chunks) and only its chunks are visited.  An empty range still runs
// chunk 0, whose zero-filled loads initialise the accumulator.  Producer and MMA issuer evaluate the same function.
__device__ __forceinline__ void umma_live_chunks(const UmmaConv& p, int row0, int row1, int& q_begin, int& q_end) {
    int i_lo = p.th, i_hi = -1;
    const int rows[2] = {row0, row1};
#pragma unroll
    for (int a = 0; a < 2; ++a) {
        const int r = rows[a];
        int lo, hi;
        if (p.tap_h_mul > 0) { lo = max(0, -r); hi = min(p.th - 1, p.srcH - 1 - r); }
        else                 { lo = max(0, r - (p.srcH - 1)); hi = min(p.th - 1, r); }
        if (lo <= hi) { i_lo = min(i_lo, lo); i_hi = max(i_hi, hi); }
    }
    if (i_hi < i_lo) { q_begin = 0; q_end = 1; return; }
    const int t_lo = i_lo * p.tw, t_hi = (i_hi + 1) * p.tw;          // taps [t_lo, t_hi)
    if (p.cin_eff == KCHUNK) { q_begin = t_lo * p.cin_chunks; q_end = t_hi * p.cin_chunks; }
    else { q_begin = t_lo / p.tpc; q_end = (t_hi + p.tpc - 1) / p.tpc; }
}

struct __align__(8) UmmaBarriers {
    uint64_t full[8], empty[8], acc_full[2], acc_empty[2], wfull;
    uint32_t tmem_base;
};

__global__ void __launch_bounds__(UM_THREADS, 1) umma_conv_kernel(const __grid_constant__ CUtensorMap tmap_a,
                                                                 const __grid_constant__ CUtensorMap tmap_b,
                                                                 const UmmaConv p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int b_plane_bytes = p.n_tile * 128;
    // weights-resident mode: [all chunks: planes x n_tile x 128 B][A-only stages]; else [stages of A + B]
    const int w_chunk_bytes = p.planes * b_plane_bytes;
    const int w_res_bytes = p.w_resident ? p.n_chunks * w_chunk_bytes : 0;
    const int stage_bytes = p.planes * A_PLANE_BYTES + (p.w_resident ? 0 : w_chunk_bytes);
    uint8_t* w_res = smem;
    uint8_t* stages_base = smem + w_res_bytes;
    UmmaBarriers* bars = reinterpret_cast<UmmaBarriers*>(stages_base + (size_t)p.stages * stage_bytes);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_mtiles = (p.n_atoms + 1) >> 1;
    const int n_tiles = n_mtiles * p.n_ntiles;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmap_a);
        prefetch_tmap(&tmap_b);
        for (int s = 0; s < p.stages; ++s) { mbar_init(&bars->full[s], 1); mbar_init(&bars->empty[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&bars->acc_full[b], 1); mbar_init(&bars->acc_empty[b], 4); }
        mbar_init(&bars->wfull, 1);
        fence_barrier_init();
    }
    if (warp == 2) { tmem_alloc(&bars->tmem_base, 512); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;

    if (warp == 0) {
        // ===== TMA producer: the whole warp runs the loop in lockstep and the loads of one stage are issued by
        // different lanes (one load each), so the per-load index arithmetic runs in parallel instead of on one thread =====
        if (p.w_resident) {                                     // every tile of this CTA uses the same (only) N tile
            if (lane == 0) mbar_expect_tx(&bars->wfull, (uint32_t)w_res_bytes);
            __syncwarp();
            for (int l = lane; l < p.n_chunks * p.planes; l += 32) {
                const int q = l / p.planes, pl = l - q * p.planes;
                tma_load_4d(w_res + q * w_chunk_bytes + pl * b_plane_bytes, &tmap_b, &bars->wfull, 0, 0, q, pl);
            }
        }
        // Each lane owns ONE load of a stage -- (plane, atom, tap-in-chunk) of the activations or one weight plane --
        // decoded once here; inside the chunk loop the (tap row, tap column, channel chunk) of chunk q advance
        // incrementally.  The loop must stay short: a single warp retires roughly one dependent instruction per
        // 6-8 cycles and a chunk's MMAs last only ~770 cycles (ncu: runtime integer divisions here made the producer
        // the bottleneck of every launch of this kernel).
        const int taps_per_load = p.cin_eff == KCHUNK ? 1 : p.tpc;
        const int n_a_loads = p.planes * 2 * taps_per_load;     // <= 2 * 2 * 4 = 16
        const int n_loads = n_a_loads + (p.w_resident ? 0 : p.planes);
        const bool is_a = lane < n_a_loads, is_b = !is_a && lane < n_loads;
        const int my_t = is_a ? lane % taps_per_load : 0;
        const int my_a = is_a ? (lane / taps_per_load) & 1 : 0;
        const int my_pl = is_a ? lane / (2 * taps_per_load) : lane - n_a_loads;
        const uint32_t my_off = is_a ? (uint32_t)(my_pl * A_PLANE_BYTES + my_a * (KCHUNK * 128) + my_t * p.cin_eff * 128)
                                     : (uint32_t)(p.planes * A_PLANE_BYTES + my_pl * b_plane_bytes);
        const int my_rep = my_pl * p.nrep;
        int stage = 0; uint32_t phase = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const int mtile = tile / p.n_ntiles, ntile = tile - mtile * p.n_ntiles;
            int aoh[2], my_b = 0, my_w0 = 0, my_row = 0;
            for (int a = 0; a < 2; ++a) {
                const int atom = mtile * 2 + a;                 // beyond n_atoms -> batch index OOB -> zero fill
                const int row = atom / p.AW;
                const int b = row / p.PH;
                aoh[a] = (row - b * p.PH) * p.h_mul + p.h_off;
                if (a == my_a) { my_b = b; my_w0 = (atom - row * p.AW) * ATOM; my_row = aoh[a]; }
            }
            int q_begin, q_end;
            umma_live_chunks(p, aoh[0], aoh[1], q_begin, q_end);
            // chunk q_begin: full-chunk mode -> (tap, channel chunk cc); shared-chunk mode -> this lane's tap q*tpc + t
            int ti, tj, cc = 0;
            {
                const int tap0 = p.cin_eff == KCHUNK ? q_begin / p.cin_chunks : q_begin * p.tpc + my_t;
                if (p.cin_eff == KCHUNK) cc = q_begin - tap0 * p.cin_chunks;
                ti = tap0 / p.tw; tj = tap0 - ti * p.tw;
            }
            for (int q = q_begin; q < q_end; ++q) {
                if (lane == 0) {
                    mbar_wait(&bars->empty[stage], phase ^ 1);
                    mbar_expect_tx(&bars->full[stage], (uint32_t)stage_bytes);
                }
                __syncwarp();
                uint8_t* dst = stages_base + (size_t)stage * stage_bytes + my_off;
                if (is_a) {
                    // a phantom tap (beyond the kernel, last shared chunk) re-reads tap 0: its packed weights are zero
                    const bool phantom = ti >= p.th;
                    tma_load_5d(dst, &tmap_a, &bars->full[stage], my_w0, my_row + (phantom ? 0 : ti * p.tap_h_mul),
                                cc * KCHUNK, my_b, my_rep + (phantom ? 0 : tj));
                } else if (is_b) {
                    tma_load_4d(dst, &tmap_b, &bars->full[stage], 0, ntile * p.n_tile, q, my_pl);
                }
                if (p.cin_eff == KCHUNK) {
                    if (++cc == p.cin_chunks) { cc = 0; if (++tj == p.tw) { tj = 0; ++ti; } }
                } else {
                    tj += p.tpc;
                    while (tj >= p.tw) { tj -= p.tw; ++ti; }
                }
                if (++stage == p.stages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: warp-uniform loop, one elected lane issues =====
        {
            const uint32_t idesc = make_idesc_bf16(128, p.n_tile, /*A MN-major*/ 1, /*B K-major*/ 0);
            int stage = 0; uint32_t phase = 0;
            int buf = 0; uint32_t acc_phase = 0;
            if (p.w_resident) { mbar_wait(&bars->wfull, 0); tc_fence_after(); }
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                mbar_wait(&bars->acc_empty[buf], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)buf * 256;
                int aoh[2];
                {
                    const int mtile = tile / p.n_ntiles;
                    for (int a = 0; a < 2; ++a) {
                        const int row = (mtile * 2 + a) / p.AW;
                        aoh[a] = (row - (row / p.PH) * p.PH) * p.h_mul + p.h_off;
                    }
                }
                int q_begin, q_end;
                umma_live_chunks(p, aoh[0], aoh[1], q_begin, q_end);
                for (int q = q_begin; q < q_end; ++q) {
                    mbar_wait(&bars->full[stage], phase);
                    tc_fence_after();
                    const uint32_t st = smem_u32(stages_base + (size_t)stage * stage_bytes);
                    const uint32_t b0 = p.w_resident ? smem_u32(w_res + q * w_chunk_bytes) : st + p.planes * A_PLANE_BYTES;
                    const int ncombo = p.planes == 2 ? 3 : 1;
                    if (elect_one()) {
                        for (int cb = 0; cb < ncombo; ++cb) {
                            const uint32_t a_addr = st + (cb == 2 ? A_PLANE_BYTES : 0);      // (hi,hi) (hi,lo) (lo,hi)
                            const uint32_t b_addr = b0 + (cb == 1 ? b_plane_bytes : 0);
                            // descriptors advance by a constant per K step: +16 K-rows of 128 B (A, MN-major), +32 B (B)
                            const uint64_t ad0 = make_smem_desc(a_addr, KCHUNK * 128, 1024);
                            const uint64_t bd0 = make_smem_desc(b_addr, 16, 1024);
#pragma unroll
                            for (int k = 0; k < KCHUNK / 16; ++k)
                                mma_bf16(d_tmem, ad0 + (uint64_t)(k * ((16 * 128) >> 4)), bd0 + (uint64_t)(k * (32 >> 4)), idesc,
                                         ((q - q_begin) | cb | k) != 0);
                        }
                        tc_commit(&bars->empty[stage]);       // frees the smem stage once these MMAs retire
                    }
                    __syncwarp();
                    if (++stage == p.stages) { stage = 0; phase ^= 1; }
                }
                if (elect_one()) tc_commit(&bars->acc_full[buf]);              // accumulator complete -> epilogue
                __syncwarp();
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
            const int pw = (atom - row * p.AW) * ATOM + (r & 63);
            const int b = row / p.PH;
            const int oh = (row - b * p.PH) * p.oh_mul + p.oh_off, ow = pw * p.ow_mul + p.ow_off;
            const bool valid = atom < p.n_atoms && pw < p.PW && oh < p.out_H && ow < p.out_W;
            mbar_wait(&bars->acc_full[buf], acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)buf * 256;
            const size_t chan_stride = (size_t)p.out_H * p.out_W;
            if (p.fuse) {
                // class-fused stride-2 data gradient: pixel (a, e) of the class domain owns dx[2a + rh, 2e + rw] for the
                // four classes; the rw = 0 / 1 values of one channel are stored back to back, so a warp fills whole
                // 32-byte sectors (the per-class kernels left half-written sectors to be merged in DRAM)
                const int a = row - b * p.PH;
                const bool in_dom = atom < p.n_atoms && pw < p.PW;
                float* yb = p.y + (size_t)b * p.fuse * chan_stride;
#pragma unroll 1
                for (int rh = 0; rh < 2; ++rh) {
                    uint32_t v0[32], v1[32];
                    tmem_ld32(taddr + (2 * rh) * 32, v0);
                    tmem_ld32(taddr + (2 * rh + 1) * 32, v1);
                    tmem_ld_wait();
                    const int h = 2 * a + rh, w0 = 2 * pw;
                    if (in_dom && h < p.out_H) {
                        float* yo = yb + (size_t)h * p.out_W + w0;
                        const bool ok0 = w0 < p.out_W, ok1 = w0 + 1 < p.out_W;
#pragma unroll
                        for (int c = 0; c < 32; ++c) {
                            if (ok0) yo[(size_t)c * chan_stride] = __uint_as_float(v0[c]);
                            if (ok1) yo[(size_t)c * chan_stride + 1] = __uint_as_float(v1[c]);
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bars->acc_empty[buf]);
                if (++buf == 2) { buf = 0; acc_phase ^= 1; }
                continue;
            }
            float* ybase = p.y + ((size_t)b * p.n_rows_out * p.out_H + oh) * p.out_W + ow;
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
// One launch of umma_conv_kernel: out[b, n, (oh,ow)(ph,pw)] = sum_{taps, c} src[b, c, row(ph,i), col(pw,j)] * w
struct ConvProblem {
    const float* src; int in_ch, srcH, srcW;          // operand read through TMA (x, or dy for the data gradient)
    int out_ch;
    TapMap tm;                                         // tap grid + where each tap sits in the (kh, kw) kernel
    int PH, PW;                                        // pixel domain of the GEMM rows
    int h_mul, tap_h_mul, h_off;                       // source row    = h_mul*ph + tap_h_mul*i + h_off
    int w_mul, rep_mul, w_off;                         // source column = w_mul*pw + rep_mul*j + w_off (baked into replicas)
    int out_H, out_W, oh_mul, oh_off, ow_mul, ow_off;  // destination pixel
    // operand already packed by the caller / a sibling problem: [plane][pre_nrep][rows][pre_Wp], replica r as above
    const __nv_bfloat16* pre; int pre_Wp, pre_nrep;
    int fuse;                                          // channels per parity class of a class-fused data gradient (0: off)
};

struct UmmaPlan {
    bool ok;
    int Wp, cin_eff, cin_chunks, tpc, n_chunks, n_tile, n_ntiles, planes, stages, w_resident;
    size_t act_bytes, w_bytes, smem_bytes;
};

static UmmaPlan plan_problem(const ConvProblem& c, int batch, int precision) {
    UmmaPlan u{};
    u.ok = false;
    const int ci = c.in_ch, co = c.out_ch;
    if (!(ci == 16 || ci == 32 || ci % 64 == 0)) return u;
    if (co % 32 != 0 || !(co <= 128 || co % 128 == 0)) return u;
    if (c.PH <= 0 || c.PW <= 0 || c.tm.th <= 0 || c.tm.tw <= 0) return u;
    if ((long)batch * c.PH * ((c.PW + ATOM - 1) / ATOM) > (1l << 30)) return u;
    u.cin_eff = ci >= KCHUNK ? KCHUNK : ci;
    u.cin_chunks = ci >= KCHUNK ? ci / KCHUNK : 1;
    u.tpc = KCHUNK / u.cin_eff;
    const int ntaps = c.tm.th * c.tm.tw;
    u.n_chunks = ci >= KCHUNK ? ntaps * u.cin_chunks : (ntaps + u.tpc - 1) / u.tpc;
    u.n_tile = co <= 128 ? co : 128;
    u.n_ntiles = co / u.n_tile;
    u.planes = precision == 1 ? 1 : 2;
    u.Wp = c.pre ? c.pre_Wp : (c.PW + 7) & ~7;         // replicas are addressed by GEMM pixel column
    int stage_bytes = u.planes * (A_PLANE_BYTES + u.n_tile * 128);
    // all weight chunks fit next to >= 3 activation-only stages: keep them resident (every tile re-reads them otherwise)
    const long w_all = (long)u.n_chunks * u.planes * u.n_tile * 128;
    u.w_resident = u.n_ntiles == 1 && w_all + 3L * u.planes * A_PLANE_BYTES <= SMEM_LIMIT - 2048;
    long avail = SMEM_LIMIT - 2048;
    if (u.w_resident) { stage_bytes = u.planes * A_PLANE_BYTES; avail -= w_all; }
    u.stages = (int)(avail / stage_bytes);
    if (u.stages > 8) u.stages = 8;
    if (u.stages < 2) return u;
    u.smem_bytes = (size_t)(u.w_resident ? w_all : 0) + (size_t)u.stages * stage_bytes + sizeof(UmmaBarriers) + 1024;
    u.act_bytes = c.pre ? 0 : align_up((size_t)u.planes * c.tm.tw * batch * ci * c.srcH * u.Wp * 2, 1024);
    u.w_bytes = align_up((size_t)u.planes * u.n_chunks * co * KCHUNK * 2, 1024);
    u.ok = true;
    return u;
}

static int run_problem(const ConvProblem& c, const float* w, const float* bias, float* out, const cpc_conv_params* p,
                       int relu, void* workspace, size_t workspace_bytes, cudaStream_t s) {
    const int B = p->batch;
    UmmaPlan u = plan_problem(c, B, p->precision);
    if (!u.ok) return CPC_ERR_UNSUPPORTED;
    if (!workspace || workspace_bytes < u.act_bytes + u.w_bytes + 1024) return CPC_ERR_WORKSPACE;
    uint8_t* ws = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(workspace) + 1023) & ~(uintptr_t)1023);
    const __nv_bfloat16* act = c.pre ? c.pre : reinterpret_cast<__nv_bfloat16*>(ws);
    __nv_bfloat16* wp = reinterpret_cast<__nv_bfloat16*>(ws + u.act_bytes);
    const long rows = (long)B * c.in_ch * c.srcH;
    const int nrep = c.pre ? c.pre_nrep : c.tm.tw;
    {
        if (!c.pre && pack_split_launch(c.src, reinterpret_cast<__nv_bfloat16*>(ws), rows, c.srcW, u.Wp, u.planes, c.tm.tw,
                                        c.w_mul, c.rep_mul, c.w_off, s) != CPC_OK)
            return CPC_ERR_CUDA;
        const long wtotal = (long)u.n_chunks * c.out_ch * KCHUNK;
        int wblocks = (int)((wtotal + 255) / 256);
        if (wblocks > 148 * 8) wblocks = 148 * 8;
        pack_weights_kernel<<<wblocks, 256, 0, s>>>(w, wp, p->c_in, p->kh, p->kw, c.out_ch, c.in_ch, u.cin_eff,
                                                    u.cin_chunks, u.tpc, u.n_chunks, u.planes, c.tm);
        CPC_LAUNCH_CHECK();
    }
    CUtensorMap ta, tb;
    {
        // replica / plane index is the outermost dimension
        const uint64_t dims[5] = {(uint64_t)u.Wp, (uint64_t)c.srcH, (uint64_t)c.in_ch, (uint64_t)B,
                                  (uint64_t)u.planes * nrep};
        const uint64_t row_b = (uint64_t)u.Wp * 2;
        const uint64_t strides[4] = {row_b, row_b * c.srcH, row_b * c.srcH * c.in_ch, row_b * c.srcH * c.in_ch * B};
        const uint32_t box[5] = {ATOM, 1, (uint32_t)u.cin_eff, 1, 1};
        if (!make_tmap_bf16(&ta, act, 5, dims, strides, box)) return CPC_ERR_CUDA;
        const uint64_t wd[4] = {KCHUNK, (uint64_t)c.out_ch, (uint64_t)u.n_chunks, (uint64_t)u.planes};
        const uint64_t wsr[3] = {KCHUNK * 2, (uint64_t)KCHUNK * 2 * c.out_ch, (uint64_t)KCHUNK * 2 * c.out_ch * u.n_chunks};
        const uint32_t wbox[4] = {KCHUNK, (uint32_t)u.n_tile, 1, 1};
        if (!make_tmap_bf16(&tb, wp, 4, wd, wsr, wbox)) return CPC_ERR_CUDA;
    }
    UmmaConv k{};
    k.AW = (c.PW + ATOM - 1) / ATOM;
    k.PH = c.PH; k.PW = c.PW;
    k.n_atoms = B * c.PH * k.AW;
    k.n_rows_out = c.out_ch; k.n_tile = u.n_tile; k.n_ntiles = u.n_ntiles;
    k.th = c.tm.th; k.tw = c.tm.tw; k.ntaps = c.tm.th * c.tm.tw; k.tpc = u.tpc; k.cin_eff = u.cin_eff;
    k.cin_chunks = u.cin_chunks; k.n_chunks = u.n_chunks;
    k.h_mul = c.h_mul; k.tap_h_mul = c.tap_h_mul; k.h_off = c.h_off;
    k.out_H = c.out_H; k.out_W = c.out_W; k.oh_mul = c.oh_mul; k.oh_off = c.oh_off; k.ow_mul = c.ow_mul; k.ow_off = c.ow_off;
    k.fuse = c.fuse;
    k.planes = u.planes; k.nrep = nrep; k.w_resident = u.w_resident; k.relu = relu; k.stages = u.stages; k.bias = bias; k.y = out; k.srcH = c.srcH;
    if (cudaFuncSetAttribute(umma_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT) != cudaSuccess)
        return CPC_ERR_CUDA;
    const int n_tiles = ((k.n_atoms + 1) / 2) * k.n_ntiles;
    const int grid = n_tiles < 148 ? n_tiles : 148;
    umma_conv_kernel<<<grid, UM_THREADS, u.smem_bytes, s>>>(ta, tb, k);
    CPC_LAUNCH_CHECK();
    count_launch(c.pre ? 2 : 3);
    return CPC_OK;
}

static ConvProblem fwd_problem(const float* x, const cpc_conv_params* p) {
    ConvProblem c{};
    c.src = x; c.in_ch = p->c_in; c.srcH = p->h_in; c.srcW = p->w_in; c.out_ch = p->c_out;
    c.tm = TapMap{p->kh, p->kw, 0, 1, 0, 1, 0};
    c.PH = p->h_out; c.PW = p->w_out;
    c.h_mul = p->stride_h; c.tap_h_mul = 1; c.h_off = -p->pad_top;
    c.w_mul = p->stride_w; c.rep_mul = 1; c.w_off = -p->pad_left;
    c.out_H = p->h_out; c.out_W = p->w_out; c.oh_mul = 1; c.oh_off = 0; c.ow_mul = 1; c.ow_off = 0;
    return c;
}

// Data gradient, parity class (rh, rw): dx[h, w] with (h + pt) % sh == rh, (w + pl) % sw == rw.
//   h = sh*a + rh - pt, taps i = rh + sh*i' read dy row a - i'.  Returns false when the class has no pixels.
static bool dgrad_problem(const float* dy, const cpc_conv_params* p, int rh, int rw, ConvProblem& c, bool& has_taps) {
    const int sh = p->stride_h, sw = p->stride_w;
    c = ConvProblem{};
    c.src = dy; c.in_ch = p->c_out; c.srcH = p->h_out; c.srcW = p->w_out; c.out_ch = p->c_in;
    const int th = rh < p->kh ? (p->kh - rh + sh - 1) / sh : 0;
    const int tw = rw < p->kw ? (p->kw - rw + sw - 1) / sw : 0;
    has_taps = th > 0 && tw > 0;
    c.tm = TapMap{th, tw, rh, sh, rw, sw, 1};
    const int a_min = p->pad_top > rh ? (p->pad_top - rh + sh - 1) / sh : 0;
    const int e_min = p->pad_left > rw ? (p->pad_left - rw + sw - 1) / sw : 0;
    const int a_top = p->h_in - 1 + p->pad_top - rh, e_top = p->w_in - 1 + p->pad_left - rw;
    if (a_top < 0 || e_top < 0) return false;
    c.PH = a_top / sh - a_min + 1; c.PW = e_top / sw - e_min + 1;
    if (c.PH <= 0 || c.PW <= 0) return false;
    c.h_mul = 1; c.tap_h_mul = -1; c.h_off = a_min;
    c.w_mul = 1; c.rep_mul = -1; c.w_off = e_min;
    c.out_H = p->h_in; c.out_W = p->w_in;
    c.oh_mul = sh; c.oh_off = sh * a_min + rh - p->pad_top;
    c.ow_mul = sw; c.ow_off = sw * e_min + rw - p->pad_left;
    return true;
}

static bool channels_ok(int ci, int co) {
    return (ci == 16 || ci == 32 || ci % 64 == 0) && co % 32 == 0 && (co <= 128 || co % 128 == 0);
}

bool umma_conv_eligible(const cpc_conv_params* p, int which) {
    if (p->stride_h > 8 || p->stride_w > 8) return false;
    if (which == 0) return plan_problem(fwd_problem(nullptr, p), p->batch, p->precision).ok;
    if (!channels_ok(p->c_out, p->c_in)) return false;
    // every non-empty class with taps must be plannable
    for (int rh = 0; rh < p->stride_h; ++rh)
        for (int rw = 0; rw < p->stride_w; ++rw) {
            ConvProblem c; bool taps;
            if (!dgrad_problem(nullptr, p, rh, rw, c, taps) || !taps) continue;
            if (!plan_problem(c, p->batch, p->precision).ok) return false;
        }
    return true;
}

// The data-gradient classes read dy through replicas  column w' -> dy[w' - r + w_off].  When every class has the
// same w_off the replicas are packed ONCE (max tw replicas, widest pitch) and shared by all classes.
struct DgradShare {
    bool common;            // all classes agree on (w_mul = 1, w_off)
    int n_classes, max_tw, max_pw, w_off;
    size_t act_bytes;       // shared activation planes
    size_t max_w_bytes;     // largest packed weight block of any class
    size_t max_single;      // largest (act + w) of any class when packed per class
};

static DgradShare dgrad_share(const cpc_conv_params* p) {
    DgradShare d{};
    d.common = true;
    bool first = true;
    for (int rh = 0; rh < p->stride_h; ++rh)
        for (int rw = 0; rw < p->stride_w; ++rw) {
            ConvProblem c; bool taps;
            if (!dgrad_problem(nullptr, p, rh, rw, c, taps) || !taps) continue;
            UmmaPlan u = plan_problem(c, p->batch, p->precision);
            ++d.n_classes;
            if (first) { d.w_off = c.w_off; first = false; }
            if (c.w_off != d.w_off || c.w_mul != 1) d.common = false;
            if (c.tm.tw > d.max_tw) d.max_tw = c.tm.tw;
            if (c.PW > d.max_pw) d.max_pw = c.PW;
            if (u.w_bytes > d.max_w_bytes) d.max_w_bytes = u.w_bytes;
            if (u.act_bytes + u.w_bytes > d.max_single) d.max_single = u.act_bytes + u.w_bytes;
        }
    const int planes = p->precision == 1 ? 1 : 2;
    const int Wp = (d.max_pw + 7) & ~7;
    d.act_bytes = align_up((size_t)planes * d.max_tw * p->batch * p->c_out * p->h_out * Wp * 2, 1024);
    return d;
}

// Class-fused data gradient of a stride-2 conv (all four output parity classes in ONE launch): the GEMM pixel domain
// is (a, e) = (h / 2, w / 2), the tap grid the 2 x 2 shifts (i', j') of dy, and the N = 4 * C_in = 128 accumulator
// columns are (class, channel); weight rows of taps a class does not have are zero.  Every dy tile is fetched once
// for all classes (the per-class launches fetch it 9 / 4 times) and dx is written in full sectors.
static bool fused_dgrad_problem(const float* dy, const cpc_conv_params* p, ConvProblem& c) {
    if (p->flags & CPC_CONV_FLAG_NO_FUSED_DGRAD) return false;
    if (p->stride_h != 2 || p->stride_w != 2 || p->pad_top != 0 || p->pad_left != 0) return false;
    if (p->kh < 2 || p->kh > 4 || p->kw < 2 || p->kw > 4 || p->c_in != 32) return false;
    c = ConvProblem{};
    c.src = dy; c.in_ch = p->c_out; c.srcH = p->h_out; c.srcW = p->w_out; c.out_ch = 4 * p->c_in;
    c.tm = TapMap{2, 2, 0, 2, 0, 2, 1, p->c_in};
    c.PH = (p->h_in + 1) / 2; c.PW = (p->w_in + 1) / 2;
    c.h_mul = 1; c.tap_h_mul = -1; c.h_off = 0;
    c.w_mul = 1; c.rep_mul = -1; c.w_off = 0;
    c.out_H = p->h_in; c.out_W = p->w_in; c.oh_mul = 2; c.oh_off = 0; c.ow_mul = 2; c.ow_off = 0;
    c.fuse = p->c_in;
    return plan_problem(c, p->batch, p->precision).ok;
}

// Canonical caller-packed dy (cpc_conv_pack operand 1): one plain replica, or -- when the data gradient of this strided
// conv reads dy through two column-shifted replicas (w' -> dy[w' - r]) -- both of them, [plane][2][rows][round8(w_out+1)],
// so that ONE packing pass serves the weight gradient (replica 0) and every data-gradient class.
int umma_dy_replicas(const cpc_conv_params* p) {
    if (p->stride_h == 1 && p->stride_w == 1) return 1;
    if (!umma_conv_eligible(p, 1)) return 1;
    ConvProblem f;
    if (fused_dgrad_problem(nullptr, p, f)) return 2;
    const DgradShare d = dgrad_share(p);
    return (d.common && d.n_classes > 1 && d.max_tw == 2 && d.w_off == 0) ? 2 : 1;
}
static int dy_packed_wp(const cpc_conv_params* p, int nrep) { return (p->w_out + (nrep == 2 ? 1 : 0) + 7) & ~7; }

size_t umma_conv_workspace(const cpc_conv_params* p, int which) {
    if (!umma_conv_eligible(p, which)) return 0;
    size_t need = 0;
    if (which == 0) {
        UmmaPlan u = plan_problem(fwd_problem(nullptr, p), p->batch, p->precision);
        need = u.act_bytes + u.w_bytes;
    } else {
        const DgradShare d = dgrad_share(p);
        need = d.max_single;
        if (d.common && d.act_bytes + d.max_w_bytes > need) need = d.act_bytes + d.max_w_bytes;
        ConvProblem f;
        if (fused_dgrad_problem(nullptr, p, f)) {
            UmmaPlan u = plan_problem(f, p->batch, p->precision);
            if (u.act_bytes + u.w_bytes > need) need = u.act_bytes + u.w_bytes;
        }
    }
    return need + 2048;
}

// Canonical caller-packed layouts (cpc_conv_pack):
//   x : [plane][kw][B*Cin*H][round8(w_out)],  replica r column w' = x[stride_w * w' + r - pad_left]
//   dy: [plane][1][B*Cout*OH][round8(w_out)]  (plain)
// which = 0: y = conv(x, w) + bias ; which = 1: dx = conv_transpose(dy, w).   `in` is x or dy, `out` is y or dx;
// `pre` is the caller-packed copy of `in` (or NULL).
int umma_conv_launch(const float* in, const float* w, const float* bias, float* out, const cpc_conv_params* p, int which,
                     const void* pre, void* workspace, size_t workspace_bytes, cudaStream_t s) {
    const __nv_bfloat16* prep = reinterpret_cast<const __nv_bfloat16*>(pre);
    if (which == 0) {
        ConvProblem c = fwd_problem(in, p);
        if (prep) { c.pre = prep; c.pre_Wp = (p->w_out + 7) & ~7; c.pre_nrep = p->kw; }
        return run_problem(c, w, bias, out, p, p->relu, workspace, workspace_bytes, s);
    }
    const int pre_rep = prep ? umma_dy_replicas(p) : 1;          // layout of the caller-packed dy
    {
        ConvProblem f;
        if (fused_dgrad_problem(in, p, f)) {
            if (prep && pre_rep == 2) { f.pre = prep; f.pre_Wp = dy_packed_wp(p, 2); f.pre_nrep = 2; }
            return run_problem(f, w, nullptr, out, p, 0, workspace, workspace_bytes, s);
        }
    }
    bool need_zero = false;
    for (int rh = 0; rh < p->stride_h; ++rh)
        for (int rw = 0; rw < p->stride_w; ++rw) {
            ConvProblem c; bool taps;
            if (dgrad_problem(in, p, rh, rw, c, taps) && !taps) need_zero = true;
        }
    if (need_zero &&
        cudaMemsetAsync(out, 0, sizeof(float) * (size_t)p->batch * p->c_in * p->h_in * p->w_in, s) != cudaSuccess)
        return CPC_ERR_CUDA;
    const DgradShare d = dgrad_share(p);
    const int pre_Wp = (p->w_out + 7) & ~7;
    // caller-packed plain dy serves every class iff no class needs a shifted replica
    const bool use_pre = prep && pre_rep == 1 && d.common && d.max_tw == 1 && d.w_off == 0 && d.max_pw <= pre_Wp;
    const bool use_pre2 = prep && pre_rep == 2 && d.common && d.max_tw <= 2 && d.w_off == 0;
    const __nv_bfloat16* shared = nullptr;
    int shared_Wp = 0, shared_nrep = 0;
    uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
    size_t ws_bytes = workspace_bytes;
    if (use_pre) {
        shared = prep; shared_Wp = pre_Wp; shared_nrep = 1;
    } else if (use_pre2) {
        shared = prep; shared_Wp = dy_packed_wp(p, 2); shared_nrep = 2;
    } else if (d.common && d.n_classes > 1) {
        if (!workspace || workspace_bytes < d.act_bytes + d.max_w_bytes + 2048) return CPC_ERR_WORKSPACE;
        uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(workspace) + 1023) & ~(uintptr_t)1023);
        shared_Wp = (d.max_pw + 7) & ~7;
        shared_nrep = d.max_tw;
        const int planes = p->precision == 1 ? 1 : 2;
        if (pack_split_launch(in, reinterpret_cast<__nv_bfloat16*>(base), (long)p->batch * p->c_out * p->h_out, p->w_out,
                              shared_Wp, planes, shared_nrep, 1, -1, d.w_off, s) != CPC_OK)
            return CPC_ERR_CUDA;
        count_launch();
        shared = reinterpret_cast<const __nv_bfloat16*>(base);
        ws = base + d.act_bytes;
        ws_bytes = workspace_bytes - (size_t)(ws - reinterpret_cast<uint8_t*>(workspace));
    }
    for (int rh = 0; rh < p->stride_h; ++rh)
        for (int rw = 0; rw < p->stride_w; ++rw) {
            ConvProblem c; bool taps;
            if (!dgrad_problem(in, p, rh, rw, c, taps) || !taps) continue;
            if (shared) { c.pre = shared; c.pre_Wp = shared_Wp; c.pre_nrep = shared_nrep; }
            const int st = run_problem(c, w, nullptr, out, p, 0, ws, ws_bytes, s);
            if (st != CPC_OK) return st;
        }
    return CPC_OK;
}


// =====================================================================================================
// Weight gradient of a (strided) conv on the tensor cores.
//
//   dW[co, ci, i, j] = sum_{b, oh, ow} dy[b, co, oh, ow] * x[b, ci, sh*oh + i - pt, sw*ow + j - pl]
//
// The reduction runs over pixels, which are contiguous in NCHW for BOTH operands, so both are K-major
// SWIZZLE_128B tiles straight out of TMA: one K chunk = 64 consecutive pixels of one output row.
//   A tile (M = 128 rows) = CB input channels x TH vertical taps of x, one TMA box (64 w, TH h, CB c);
//   B tile (N rows)       = N output channels of dy,                   one TMA box (64 w, 1 h, N c).
// A CTA owns one (x-row tile, co tile, horizontal tap j) output tile and a slice of the pixel range
// (split-K); the fp32 TMEM accumulator is flushed with red.global.add into the zero-initialised dW.
// =====================================================================================================
struct UmmaWgrad {
    int OH, AW, B;                        // pixel chunks: (b, oh, w-atom)
    int n_chunks_total, chunks_per_split, n_splits;
    int n_xtiles, n_ntiles, n_tile;
    int CB, TH, cin, cout, kh, kw;        // M tile = CB channels x TH taps
    int c_tiles, h_tiles;                 // x tiles = c_tiles * h_tiles
    int pt, sh, planes, stages;
    float* dw;
};

__global__ void __launch_bounds__(UM_THREADS, 1) umma_wgrad_kernel(const __grid_constant__ CUtensorMap tmap_x,
                                                                  const __grid_constant__ CUtensorMap tmap_dy,
                                                                  const UmmaWgrad p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    constexpr int A_BYTES = 128 * 128;                           // 128 rows x 64 pixels bf16
    const int b_bytes = p.n_tile * 128;
    const int stage_bytes = p.planes * (A_BYTES + b_bytes);
    UmmaBarriers* bars = reinterpret_cast<UmmaBarriers*>(smem + (size_t)p.stages * stage_bytes);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // tile decode: blockIdx.x = ((j * n_xtiles + xtile) * n_ntiles + ntile), blockIdx.y = split
    int t = blockIdx.x;
    const int ntile = t % p.n_ntiles; t /= p.n_ntiles;
    const int xtile = t % p.n_xtiles;
    const int j = t / p.n_xtiles;
    const int ctile = xtile / p.h_tiles, htile = xtile - ctile * p.h_tiles;
    const int c0 = ctile * p.CB, i0 = htile * p.TH;
    const int q_begin = blockIdx.y * p.chunks_per_split;
    const int q_end = min(p.n_chunks_total, q_begin + p.chunks_per_split);
    const int nq = q_end - q_begin;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmap_x);
        prefetch_tmap(&tmap_dy);
        for (int s = 0; s < p.stages; ++s) { mbar_init(&bars->full[s], 1); mbar_init(&bars->empty[s], 1); }
        mbar_init(&bars->acc_full[0], 1);
        fence_barrier_init();
    }
    if (warp == 2) { tmem_alloc(&bars->tmem_base, 256); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;

    if (warp == 0) {
        // whole warp in lockstep; lanes 0 .. 2*planes-1 issue one TMA load each (x / dy, per plane)
        int stage = 0; uint32_t phase = 0;
        // (item, output row, 64-pixel atom) of chunk q advance incrementally: no divisions in the loop
        int wa, b, oh;
        {
            const int row = q_begin / p.AW;
            wa = q_begin - row * p.AW;
            b = row / p.OH; oh = row - b * p.OH;
        }
        for (int q = q_begin; q < q_end; ++q) {
            const int w0 = wa * ATOM;
            if (lane == 0) {
                mbar_wait(&bars->empty[stage], phase ^ 1);
                mbar_expect_tx(&bars->full[stage], (uint32_t)stage_bytes);
            }
            __syncwarp();
            uint8_t* st = smem + (size_t)stage * stage_bytes;
            if (lane < 2 * p.planes) {
                const int pl = lane >> 1;
                if ((lane & 1) == 0)
                    tma_load_5d(st + pl * A_BYTES, &tmap_x, &bars->full[stage], w0, oh * p.sh + i0 - p.pt, c0, b, pl * p.kw + j);
                else
                    tma_load_5d(st + p.planes * A_BYTES + pl * b_bytes, &tmap_dy, &bars->full[stage], w0, oh, ntile * p.n_tile,
                                b, pl);
            }
            if (++wa == p.AW) { wa = 0; if (++oh == p.OH) { oh = 0; ++b; } }
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = make_idesc_bf16(128, p.n_tile, 0, 0);
            int stage = 0; uint32_t phase = 0;
            for (int q = 0; q < nq; ++q) {
                mbar_wait(&bars->full[stage], phase);
                tc_fence_after();
                const uint32_t st = smem_u32(smem + (size_t)stage * stage_bytes);
                const uint32_t b0 = st + p.planes * A_BYTES;
                const int ncombo = p.planes == 2 ? 3 : 1;
                for (int cb = 0; cb < ncombo; ++cb) {
                    const uint64_t a_d = make_smem_desc(st + (cb == 2 ? A_BYTES : 0), 16, 1024);
                    const uint64_t b_d = make_smem_desc(b0 + (cb == 1 ? b_bytes : 0), 16, 1024);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        mma_bf16(tmem_base, a_d + (uint64_t)(k * 2), b_d + (uint64_t)(k * 2), idesc, (q | cb | k) != 0);
                }
                tc_commit(&bars->empty[stage]);
                if (++stage == p.stages) { stage = 0; phase ^= 1; }
            }
            tc_commit(&bars->acc_full[0]);
        }
    } else if (warp >= 4) {
        const int ew = warp & 3;
        const int m = ew * 32 + lane;                            // accumulator row -> (channel, tap)
        const int ci = c0 + m / p.TH, i = i0 + m % p.TH;
        const bool valid = nq > 0 && ci < p.cin && i < p.kh;
        mbar_wait(&bars->acc_full[0], 0);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16);
        for (int n0 = 0; n0 < p.n_tile; n0 += 32) {
            uint32_t v[32];
            tmem_ld32(taddr + n0, v);
            tmem_ld_wait();
            if (valid) {
#pragma unroll
                for (int c = 0; c < 32; ++c) {
                    const int co = ntile * p.n_tile + n0 + c;
                    atomicAdd(p.dw + (((size_t)co * p.cin + ci) * p.kh + i) * p.kw + j, __uint_as_float(v[c]));
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) { tc_fence_after(); tmem_dealloc(tmem_base, 256); }
}

struct WgradPlan {
    bool ok;
    int CB, TH, c_tiles, h_tiles, n_tile, n_ntiles, planes, stages, Wp_x, Wp_dy, AW;
    size_t x_bytes, dy_bytes, smem_bytes;
};

static WgradPlan make_wgrad_plan(const cpc_conv_params* p) {
    WgradPlan u{};
    u.ok = false;
    if (p->stride_h > 8 || p->stride_w > 8) return u;
    const int ci = p->c_in, co = p->c_out;
    if (!(ci == 16 || ci == 32 || ci == 64 || ci % 128 == 0)) return u;
    if (co % 32 != 0 || !(co <= 256 || co % 256 == 0)) return u;
    u.CB = ci < 128 ? ci : 128;
    u.TH = 128 / u.CB;
    u.c_tiles = ci / u.CB;
    u.h_tiles = (p->kh + u.TH - 1) / u.TH;
    u.n_tile = co <= 256 ? co : 256;
    u.n_ntiles = co / u.n_tile;
    u.planes = p->precision == 1 ? 1 : 2;
    u.Wp_x = (p->w_out + 7) & ~7;                      // replicas are addressed by output column
    u.Wp_dy = (p->w_out + 7) & ~7;
    u.AW = (p->w_out + ATOM - 1) / ATOM;
    if ((long)p->batch * p->h_out * u.AW > (1l << 30)) return u;
    const int stage_bytes = u.planes * (128 * 128 + u.n_tile * 128);
    u.stages = (SMEM_LIMIT - 2048) / stage_bytes;
    if (u.stages > 8) u.stages = 8;
    if (u.stages < 2) return u;
    u.smem_bytes = (size_t)u.stages * stage_bytes + sizeof(UmmaBarriers) + 1024;
    u.x_bytes = align_up((size_t)u.planes * p->kw * p->batch * ci * p->h_in * u.Wp_x * 2, 1024);
    u.dy_bytes = align_up((size_t)u.planes * p->batch * co * p->h_out * u.Wp_dy * 2, 1024);
    u.ok = true;
    return u;
}

size_t umma_wgrad_workspace(const cpc_conv_params* p) {
    WgradPlan u = make_wgrad_plan(p);
    return u.ok ? u.x_bytes + u.dy_bytes + 1024 : 0;
}
bool umma_wgrad_eligible(const cpc_conv_params* p) { return make_wgrad_plan(p).ok; }

int umma_wgrad_launch(const float* x, const float* dy, float* dw, const cpc_conv_params* p, const void* pre_x,
                      const void* pre_dy, void* workspace, size_t workspace_bytes, cudaStream_t s) {
    WgradPlan u = make_wgrad_plan(p);
    if (!u.ok) return CPC_ERR_UNSUPPORTED;
    if (!workspace || workspace_bytes < u.x_bytes + u.dy_bytes + 1024) return CPC_ERR_WORKSPACE;
    uint8_t* ws = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(workspace) + 1023) & ~(uintptr_t)1023);
    const __nv_bfloat16* xp = pre_x ? reinterpret_cast<const __nv_bfloat16*>(pre_x) : reinterpret_cast<__nv_bfloat16*>(ws);
    const __nv_bfloat16* dyp =
        pre_dy ? reinterpret_cast<const __nv_bfloat16*>(pre_dy) : reinterpret_cast<__nv_bfloat16*>(ws + u.x_bytes);
    const int B = p->batch;
    {
        const long rows_x = (long)B * p->c_in * p->h_in;
        if (!pre_x && pack_split_launch(x, reinterpret_cast<__nv_bfloat16*>(ws), rows_x, p->w_in, u.Wp_x, u.planes, p->kw,
                                        p->stride_w, 1, -p->pad_left, s) != CPC_OK)
            return CPC_ERR_CUDA;
        const long rows_dy = (long)B * p->c_out * p->h_out;
        if (!pre_dy && pack_split_launch(dy, reinterpret_cast<__nv_bfloat16*>(ws + u.x_bytes), rows_dy, p->w_out, u.Wp_dy,
                                         u.planes, 1, 1, 0, 0, s) != CPC_OK)
            return CPC_ERR_CUDA;
    }
    CUtensorMap tx, tdy;
    {
        const uint64_t rb = (uint64_t)u.Wp_x * 2;
        const uint64_t dims[5] = {(uint64_t)u.Wp_x, (uint64_t)p->h_in, (uint64_t)p->c_in, (uint64_t)B,
                                  (uint64_t)u.planes * p->kw};
        const uint64_t strides[4] = {rb, rb * p->h_in, rb * p->h_in * p->c_in, rb * p->h_in * p->c_in * B};
        const uint32_t box[5] = {ATOM, (uint32_t)u.TH, (uint32_t)u.CB, 1, 1};
        if (!make_tmap_bf16(&tx, xp, 5, dims, strides, box)) return CPC_ERR_CUDA;
        // caller-packed dy may carry the data gradient's second replica: wider pitch, planes twice as far apart
        const int dy_rep = pre_dy ? umma_dy_replicas(p) : 1;
        const uint64_t rd = (uint64_t)(pre_dy ? dy_packed_wp(p, dy_rep) : u.Wp_dy) * 2;
        const uint64_t ddims[5] = {(uint64_t)p->w_out, (uint64_t)p->h_out, (uint64_t)p->c_out, (uint64_t)B,
                                   (uint64_t)u.planes};
        const uint64_t dstr[4] = {rd, rd * p->h_out, rd * p->h_out * p->c_out, rd * p->h_out * p->c_out * B * dy_rep};
        const uint32_t dbox[5] = {ATOM, 1, (uint32_t)u.n_tile, 1, 1};
        if (!make_tmap_bf16(&tdy, dyp, 5, ddims, dstr, dbox)) return CPC_ERR_CUDA;
    }
    UmmaWgrad k{};
    k.OH = p->h_out; k.AW = u.AW; k.B = B;
    k.n_chunks_total = B * p->h_out * u.AW;
    k.n_xtiles = u.c_tiles * u.h_tiles; k.n_ntiles = u.n_ntiles; k.n_tile = u.n_tile;
    k.CB = u.CB; k.TH = u.TH; k.cin = p->c_in; k.cout = p->c_out; k.kh = p->kh; k.kw = p->kw;
    k.c_tiles = u.c_tiles; k.h_tiles = u.h_tiles;
    k.pt = p->pad_top; k.sh = p->stride_h; k.planes = u.planes; k.stages = u.stages;
    k.dw = dw;
    const int tiles = k.n_xtiles * k.n_ntiles * p->kw;
    // one CTA per SM (shared memory): tiles * splits must not exceed the 148 SMs or a few CTAs form a second wave
    // (ncu: 153 / 150 CTAs doubled the duration of the three large weight gradients)
    int splits = 148 / tiles;
    const int max_splits = (k.n_chunks_total + 15) / 16;
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    k.chunks_per_split = (k.n_chunks_total + splits - 1) / splits;
    k.n_splits = (k.n_chunks_total + k.chunks_per_split - 1) / k.chunks_per_split;
    if (cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)p->c_out * p->c_in * p->kh * p->kw, s) != cudaSuccess)
        return CPC_ERR_CUDA;
    if (cudaFuncSetAttribute(umma_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT) != cudaSuccess)
        return CPC_ERR_CUDA;
    umma_wgrad_kernel<<<dim3(tiles, k.n_splits), UM_THREADS, u.smem_bytes, s>>>(tx, tdy, k);
    CPC_LAUNCH_CHECK();
    count_launch(3);
    return CPC_OK;
}

}  // namespace cpc
