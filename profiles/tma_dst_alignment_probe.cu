// TMA probe: does a tiled cp.async.bulk.tensor load accept a shared-memory destination that is only 32- or 16-byte
// aligned (needed to give the z-planes of a staged field box a bank-skewed pitch)?  One 16 x 12 x 1 box per plane,
// plane p stored at base + p * (16*12 + skew) floats.  Usage: ./a.out <skew_floats>
// nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 profiles/tma_dst_alignment_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>
__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void k(const __grid_constant__ CUtensorMap map, float *out, int skew, int planes) {
    extern __shared__ unsigned char raw[];
    float *t = (float *)(raw + ((128u - (s32(raw) & 127u)) & 127u));
    __shared__ __align__(8) uint64_t bar;
    const int pitch = 16 * 12 + skew;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bar)), "r"((uint32_t)(planes * 16 * 12 * 4)) : "memory");
        for (int p = 0; p < planes; p++)
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                         ::"r"(s32(t + p * pitch)), "l"(&map), "r"(4), "r"(-2), "r"(3 + p), "r"(s32(&bar)) : "memory");
    }
    __syncthreads();
    uint32_t done = 0;
    while (!done) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0,1,0,p; }" : "=r"(done) : "r"(s32(&bar)) : "memory");
    for (int i = threadIdx.x; i < planes * pitch; i += blockDim.x) out[i] = t[i];
}
typedef CUresult (*Fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main(int argc, char **argv) {
    int skew = argc > 1 ? atoi(argv[1]) : 8, planes = 4;
    size_t n = 64 * 64 * 64;
    std::vector<float> h(n);
    for (size_t i = 0; i < n; i++) h[i] = (float)i;
    float *d; cudaMalloc(&d, n * 4); cudaMemcpy(d, h.data(), n * 4, cudaMemcpyHostToDevice);
    void *p; cudaDriverEntryPointQueryResult q; cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    CUtensorMap map; memset(&map, 0, sizeof(map));
    cuuint64_t dims[3] = {64, 64, 64}, str[2] = {64 * 4, 64 * 64 * 4};
    cuuint32_t box[3] = {16, 12, 1}, es[3] = {1, 1, 1};
    printf("encode %d\n", (int)((Fn)p)(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE));
    const int pitch = 16 * 12 + skew;
    float *o; cudaMalloc(&o, planes * pitch * 4); cudaMemset(o, 0, planes * pitch * 4);
    k<<<1, 128, planes * pitch * 4 + 256>>>(map, o, skew, planes);
    cudaError_t e = cudaDeviceSynchronize();
    printf("skew %d floats (%d-byte plane alignment): %s\n", skew, (pitch * 4) % 128 == 0 ? 128 : ((pitch * 4) % 32 == 0 ? 32 : 16), cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    std::vector<float> ho(planes * pitch); cudaMemcpy(ho.data(), o, ho.size() * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int pl = 0; pl < planes; pl++) for (int y = 0; y < 12; y++) for (int x = 0; x < 16; x++) {
        int gy = y - 2, gz = 3 + pl, gx = 4 + x;
        float want = gy < 0 ? 0.0f : (float)(gx + 64 * (gy + 64 * gz));
        if (ho[pl * pitch + y * 16 + x] != want) bad++;
    }
    printf("mismatches: %d\n", bad);
    return bad != 0;
}
