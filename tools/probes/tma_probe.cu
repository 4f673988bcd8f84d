// Build: nvcc -gencode arch=compute_100a,code=sm_100a -o tma_probe tma_probe.cu -ldl ; run each case in its own process: ./tma_probe <case index>
// Probe: where does TMA put a box whose inner extent is narrower than the swizzle span?
// Tensor: bf16 [H=8][W=64], value = h*64 + w.  Box {bw, bh}; smem dumped as element indices (after un-swizzling
// is NOT applied: raw smem order), so the placement rule can be read off directly.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <dlfcn.h>

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__global__ void probe(const __grid_constant__ CUtensorMap tm, float* out, int bytes, int c0, int c1) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __nv_bfloat16* s = (__nv_bfloat16*)smem;
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) s[i] = __float2bfloat16(-1.f);
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar);
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
        asm volatile("fence.mbarrier_init.release.cluster;");
        asm volatile("fence.proxy.async.shared::cta;");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes));
        uint32_t d = (uint32_t)__cvta_generic_to_shared(smem);
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                     ::"r"(d), "l"(&tm), "r"(b), "r"(c0), "r"(c1) : "memory");
        uint32_t ok = 0;
        while (!ok) {
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
                         : "=r"(ok) : "r"(b) : "memory");
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) out[i] = __bfloat162float(s[i]);
}

int main(int argc, char** argv) {
    const int only = argc > 1 ? atoi(argv[1]) : -1;
    void* h = dlopen("libcuda.so.1", RTLD_NOW);
    EncodeFn enc = (EncodeFn)dlsym(h, "cuTensorMapEncodeTiled");
    const int H = 8, W = 64;
    __nv_bfloat16 host[H * W];
    for (int i = 0; i < H * W; ++i) host[i] = __float2bfloat16((float)i);
    __nv_bfloat16* dev; cudaMalloc(&dev, sizeof(host)); cudaMemcpy(dev, host, sizeof(host), cudaMemcpyHostToDevice);
    float* out; cudaMalloc(&out, 2048 * 4);
    struct Case { const char* name; CUtensorMapSwizzle sw; int bw, bh, es1, c0, c1; } cases[] = {
        {"SW128 box 32x2", CU_TENSOR_MAP_SWIZZLE_128B, 32, 2, 1, 0, 0},
        {"SW128 box 16x4", CU_TENSOR_MAP_SWIZZLE_128B, 16, 4, 1, 0, 0},
        {"SW64  box 32x2", CU_TENSOR_MAP_SWIZZLE_64B, 32, 2, 1, 0, 0},
        {"SW128 box 32x3 elemstride 2 (rows 1,3)", CU_TENSOR_MAP_SWIZZLE_128B, 32, 3, 2, 3, 1},
        {"SW128 box 64x3 elemstride 2 (rows 0,2)", CU_TENSOR_MAP_SWIZZLE_128B, 64, 3, 2, 0, 0},
        {"SW128 box 32x1 start col 3 row 2 (unaligned)", CU_TENSOR_MAP_SWIZZLE_128B, 32, 1, 1, 3, 2},
        {"SW128 box 64x4 elemstride 2 at row 0 (rows 0,2)", CU_TENSOR_MAP_SWIZZLE_128B, 64, 4, 2, 0, 0},
        {"SW128 box 64x2 elemstride 2 at row 1 (row 1)", CU_TENSOR_MAP_SWIZZLE_128B, 64, 2, 2, 0, 1},
    };
    int idx = -1;
    for (auto& c : cases) {
        ++idx;
        if (only >= 0 && idx != only) continue;
        CUtensorMap tm;
        cuuint64_t gd[2] = {W, H}; cuuint64_t gs[1] = {W * 2};
        cuuint32_t bx[2] = {(cuuint32_t)c.bw, (cuuint32_t)c.bh}; cuuint32_t es[2] = {1, (cuuint32_t)c.es1};
        CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dev, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, c.sw,
                         CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("== %s: encode rc=%d\n", c.name, (int)r);
        if (r != CUDA_SUCCESS) continue;
        const int rows = (c.bh + c.es1 - 1) / c.es1;
        probe<<<1, 128, 8192>>>(tm, out, c.bw * rows * 2, c.c0, c.c1);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("   kernel error %s\n", cudaGetErrorString(e)); return 1; }
        float hostout[2048]; cudaMemcpy(hostout, out, sizeof(hostout), cudaMemcpyDeviceToHost);
        for (int row = 0; row < 6; ++row) {          // 6 smem rows of 128 B (64 elements), printed as 8 chunks of 16 B: first element
            printf("   smem+%4d:", row * 128);
            for (int ch = 0; ch < 8; ++ch) printf(" %5.0f", hostout[row * 64 + ch * 8]);
            printf("\n");
        }
    }
    return 0;
}
