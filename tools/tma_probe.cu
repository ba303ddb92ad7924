// tma_probe.cu -- development probe: which TMA box shapes / data types work on this part.
// nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o /tmp/tma_probe tools/tma_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int BX, int BY, int BZ>
__global__ void probe(const __grid_constant__ CUtensorMap map, int c0, int c1, int c2, double* out, int mode) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) unsigned long long bar;
    const uint32_t b = smem_u32(&bar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(BX * BY * BZ * 8) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                     ::"r"(smem_u32(smem)), "l"(reinterpret_cast<unsigned long long>(&map)), "r"(c0), "r"(c1), "r"(c2), "r"(b) : "memory");
    }
    asm volatile(
        "{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}\n"
        ::"r"(b), "r"(0) : "memory");
    const double* s = reinterpret_cast<const double*>(smem);
    for (int i = threadIdx.x; i < BX * BY * BZ; i += blockDim.x) out[i] = s[i];
}

template <int BX, int BY, int BZ>
int run(EncodeTiledFn enc, const char* name, int nx, int ny, int nz, int c0, int c1, int c2) {
    size_t n = (size_t)nx * ny * nz;
    std::vector<double> h(n);
    for (size_t i = 0; i < n; ++i) h[i] = (double)i + 1.0;
    double *d, *o;
    cudaMalloc(&d, n * 8); cudaMalloc(&o, BX * BY * BZ * 8);
    cudaMemcpy(d, h.data(), n * 8, cudaMemcpyHostToDevice);
    CUtensorMap map;
    cuuint64_t dims[3] = {(cuuint64_t)nx, (cuuint64_t)ny, (cuuint64_t)nz};
    cuuint64_t strides[2] = {(cuuint64_t)nx * 8, (cuuint64_t)nx * ny * 8};
    cuuint32_t box[3] = {BX, BY, BZ};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("%s: encode failed %d\n", name, (int)r); return 1; }
    auto k = probe<BX, BY, BZ>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100000);
    k<<<1, 128, BX * BY * BZ * 8 + 1024>>>(map, c0, c1, c2, o, 0);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: KERNEL FAILED: %s\n", name, cudaGetErrorString(e)); return 2; }
    std::vector<double> res(BX * BY * BZ);
    cudaMemcpy(res.data(), o, res.size() * 8, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int z = 0; z < BZ; ++z) for (int y = 0; y < BY; ++y) for (int x = 0; x < BX; ++x) {
        long gx = c0 + x, gy = c1 + y, gz = c2 + z;
        double want = (gx < 0 || gx >= nx || gy < 0 || gy >= ny || gz < 0 || gz >= nz) ? 0.0
                      : (double)((gz * ny + gy) * nx + gx) + 1.0;
        if (res[(z * BY + y) * BX + x] != want) ++bad;
    }
    printf("%s: box {%d,%d,%d} tensor {%d,%d,%d} at (%d,%d,%d): %s (%d mismatches)\n", name, BX, BY, BZ, nx, ny, nz,
           c0, c1, c2, bad ? "WRONG" : "ok", bad);
    cudaFree(d); cudaFree(o);
    return bad ? 3 : 0;
}

int main(int argc, char** argv) {
    int which = argc > 1 ? atoi(argv[1]) : 0;
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaFree(0);
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p) { printf("no entry point\n"); return 1; }
    EncodeTiledFn enc = (EncodeTiledFn)p;
    int rc = 0;
    // each case in its own process slot: argv selects one (a failed kernel poisons the context)
    switch (which) {
    case 0: rc = run<64, 8, 1>(enc, "aligned64x8", 1000, 1000, 4, 64, 8, 1); break;
    case 1: rc = run<64, 8, 5>(enc, "A-tile", 1000, 1000, 10, 128, 16, 5); break;
    case 2: rc = run<66, 10, 1>(enc, "w66-odd-start", 1000, 1000, 4, 127, 7, 2); break;
    case 3: rc = run<66, 10, 1>(enc, "w66-even-start", 1000, 1000, 4, 128, 7, 2); break;
    case 4: rc = run<64, 8, 1>(enc, "w64-odd-start", 1000, 1000, 4, 127, 7, 2); break;
    case 5: rc = run<68, 10, 1>(enc, "w68-even-start", 1000, 1000, 4, 126, 7, 2); break;
    case 6: rc = run<68, 10, 1>(enc, "w68-neg-start", 1000, 1000, 4, -2, -1, 0); break;
    case 7: rc = run<68, 10, 1>(enc, "w68-oob-pos", 1000, 1000, 4, 958, 995, 3); break;
    case 8: rc = run<68, 10, 1>(enc, "w68-small-tensor", 12, 12, 3, -2, -1, 1); break;
    case 9: rc = run<64, 8, 5>(enc, "A-small-tensor", 12, 12, 10, 0, 0, 5); break;
    case 10: rc = run<72, 10, 1>(enc, "w72-even", 1000, 1000, 4, 124, 7, 2); break;
    case 11: rc = run<80, 10, 1>(enc, "w80-even", 1000, 1000, 4, 120, 7, 2); break;
    case 12: rc = run<64, 10, 1>(enc, "w64-y-odd", 1000, 1000, 4, 126, 7, 2); break;
    case 13: rc = run<128, 10, 1>(enc, "w128-even", 1000, 1000, 4, 120, 7, 2); break;
    }
    return rc;
}
