// minimal TMA probe: variants selected by argv
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
template<int RANK>
__global__ void k(const __grid_constant__ CUtensorMap tmap, float* out, int n, int x, int y, int z, unsigned bytes) {
    extern __shared__ __align__(128) unsigned char smem[];
    float* tile = (float*)smem;
    uint64_t* bar = (uint64_t*)(smem + 32768);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
        if (RANK == 3)
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                ::"r"(smem_u32(tile)), "l"(&tmap), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(z) : "memory");
        else
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                ::"r"(smem_u32(tile)), "l"(&tmap), "r"(smem_u32(bar)), "r"(x), "r"(y) : "memory");
    }
    asm volatile("{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(smem_u32(bar)) : "memory");
    for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = tile[i];
}
typedef CUresult (*Enc)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main(int argc, char** argv) {
    int rank = atoi(argv[1]), bw = atoi(argv[2]), bh = atoi(argv[3]), bq = atoi(argv[4]), x = atoi(argv[5]), y = atoi(argv[6]);
    int w = 128, h = 64, q = 40;
    float* d; cudaMalloc(&d, (size_t)w*h*q*4);
    float* hbuf = (float*)malloc((size_t)w*h*q*4);
    for (int i = 0; i < w*h*q; i++) hbuf[i] = (float)i;
    cudaMemcpy(d, hbuf, (size_t)w*h*q*4, cudaMemcpyHostToDevice);
    void* p = nullptr; cudaDriverEntryPointQueryResult qr;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr);
    printf("entry %d %d %p\n", (int)e, (int)qr, p);
    Enc enc = (Enc)p;
    CUtensorMap tm;
    cuuint64_t gdim[3] = {(cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)q};
    cuuint64_t gstr[2] = {(cuuint64_t)w*4, (cuuint64_t)w*h*4};
    cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bq};
    cuuint32_t es[3] = {1,1,1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, rank, d, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode %d\n", (int)r);
    int n = bw*bh*(rank==3?bq:1);
    float* o; cudaMalloc(&o, n*4);
    cudaFuncSetAttribute(k<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000);
    cudaFuncSetAttribute(k<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000);
    if (rank==3) k<3><<<1,128,40000>>>(tm, o, n, x, y, 1, n*4); else k<2><<<1,128,40000>>>(tm, o, n, x, y, 0, n*4);
    e = cudaDeviceSynchronize();
    printf("run: %s\n", cudaGetErrorString(e));
    if (e == cudaSuccess) { float* ho=(float*)malloc(n*4); cudaMemcpy(ho,o,n*4,cudaMemcpyDeviceToHost); printf("vals %g %g %g %g .. %g\n", ho[0],ho[1],ho[2],ho[bw],ho[n-1]); }
    return 0;
}
