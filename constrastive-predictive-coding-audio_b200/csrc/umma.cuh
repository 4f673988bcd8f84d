// sm_100a building blocks: mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (MMA / TMEM) PTX wrappers,
// shared-memory matrix descriptors and the host-side tensor-map encoder.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

namespace cpc {
namespace umma {

// ---- mbarrier -------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a mis-programmed pipeline traps (kernel error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    for (uint32_t spin = 0; !mbar_try_wait(bar, parity); ++spin) {
        if (spin > (1u << 27)) __trap();
    }
}

// One lane of a fully converged warp.  Single-thread roles (MMA issue, TMA issue) run their loops with the WHOLE
// warp and elect a lane only for the issuing instruction: the loop state then stays warp-uniform, so ptxas keeps
// descriptors / addresses in uniform registers instead of converting per-thread values (R2UR + retry loops).
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}

// ---- TMA ------------------------------------------------------------------------------------------
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* m) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_5d(void* smem, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3,
                                            int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(smem_u32(smem)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3),
        "r"(c4)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(smem)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* smem, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(smem)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}

// ---- tcgen05 / TMEM -------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols)
                 : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], bf16 inputs, fp32 accumulate.
__device__ __forceinline__ void mma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- descriptors ----------------------------------------------------------------------------------
// Shared-memory matrix descriptor (sm_100 "version 1"), SWIZZLE_128B.  Offsets are in bytes, 16-byte units
// are encoded.  K-major : rows of 128 B (64 bf16 of K), 8-row groups `sbo` apart, lbo unused.
//               MN-major: rows of 128 B (64 bf16 of M/N), 8 K-rows per group `sbo` apart, the next 64 M/N
//                         elements `lbo` apart.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;                      // descriptor version 1 (Blackwell)
    d |= (uint64_t)2 << 61;                      // SWIZZLE_128B
    return d;
}
// Instruction descriptor for kind::f16 with BF16 A/B and F32 accumulate.
__host__ __device__ inline uint32_t make_idesc_bf16(int m, int n, int a_mn_major, int b_mn_major) {
    uint32_t d = 0;
    d |= 1u << 4;                                // D format F32
    d |= 1u << 7;                                // A format BF16
    d |= 1u << 10;                               // B format BF16
    d |= (uint32_t)(a_mn_major & 1) << 15;
    d |= (uint32_t)(b_mn_major & 1) << 16;
    d |= (uint32_t)(n >> 3) << 17;
    d |= (uint32_t)(m >> 4) << 24;
    return d;
}

// ---- host: tensor maps ------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// bf16 tensor, dims[0] contiguous; strides_bytes[i] is the stride of dims[i+1]; 128B swizzle; OOB -> 0.
inline bool make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                           const uint32_t* box) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) return false;
    cuuint64_t gd[5], gs[4];
    cuuint32_t bx[5], es[5];
    for (int i = 0; i < rank; ++i) { gd[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
    for (int i = 0; i + 1 < rank; ++i) gs[i] = strides_bytes[i];
    return fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gd, gs, bx, es,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// ---- CTAs per accumulator group of a split-K weight-gradient kernel -------------------------------------------------
// A CTA owns one group of accumulators (taps) for the whole launch and a share of the `n_units` reduction units; the
// cost of one unit differs between groups (taps that pair with rows outside the source are skipped).  Hands `total`
// CTAs to the groups so that the largest per-CTA cost ceil(n_units / ctas[g]) * cost[g] is as small as possible
// (greedy: the next CTA goes to the currently slowest group); start[g] .. start[g + 1] are group g's CTAs.
constexpr int kMaxSplitGroups = 48;
inline int balance_group_ctas(const int* cost, int n_groups, int n_units, int total, int* start) {
    int ctas[kMaxSplitGroups];
    if (n_groups > kMaxSplitGroups || n_groups > total) return 0;
    for (int g = 0; g < n_groups; ++g) ctas[g] = 1;
    for (int used = n_groups; used < total; ++used) {
        int worst = -1;
        long worst_cost = -1;
        for (int g = 0; g < n_groups; ++g) {
            if (ctas[g] >= n_units) continue;                     // one unit per CTA already
            const long c = (long)((n_units + ctas[g] - 1) / ctas[g]) * (cost[g] > 0 ? cost[g] : 1);
            if (c > worst_cost) { worst_cost = c; worst = g; }
        }
        if (worst < 0) break;
        ++ctas[worst];
    }
    start[0] = 0;
    for (int g = 0; g < n_groups; ++g) start[g + 1] = start[g] + ctas[g];
    return start[n_groups];
}

// ---- dynamic tile scheduling of a persistent kernel -----------------------------------------------------------------
// A two-entry queue in shared memory: `sfull[2]` (count 1), `sempty[2]` (count = number of consumer warps), `tile_id[2]`.
// One scheduler thread takes tile numbers from a global counter that is zero at launch; every consumer warp reads each
// entry.  -1 ends the kernel.
// Scheduler side (one thread): next tile from the global counter, -1 when the work is used up.
template <class Bars>
__device__ __forceinline__ int sched_push(Bars* bars, uint32_t i, int* counter, int n_tiles) {
    const int slot = i & 1;
    mbar_wait(&bars->sempty[slot], ((i >> 1) & 1) ^ 1);
    int tile = atomicAdd(counter, 1);
    if (tile >= n_tiles) tile = -1;
    bars->tile_id[slot] = tile;
    mbar_arrive(&bars->sfull[slot]);
    return tile;
}
// Consumer side (a whole warp): reads entry i and releases its slot.
template <class Bars>
__device__ __forceinline__ int sched_pop(Bars* bars, uint32_t i, int lane) {
    const int slot = i & 1;
    mbar_wait(&bars->sfull[slot], (i >> 1) & 1);
    const int tile = *reinterpret_cast<volatile int*>(&bars->tile_id[slot]);
    __syncwarp();
    if (lane == 0) mbar_arrive(&bars->sempty[slot]);
    return tile;
}
// The same for a consumer role that is a single thread.
template <class Bars>
__device__ __forceinline__ int sched_pop_thread(Bars* bars, uint32_t i) {
    const int slot = i & 1;
    mbar_wait(&bars->sfull[slot], (i >> 1) & 1);
    const int tile = *reinterpret_cast<volatile int*>(&bars->tile_id[slot]);
    mbar_arrive(&bars->sempty[slot]);
    return tile;
}

}  // namespace umma
}  // namespace cpc
