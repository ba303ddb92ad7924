// microbench.cu -- development probe: the per-SM rates that bound the SG apply kernel on this part
// (FP64 FMA issue, shared-memory and shuffle bandwidth, L2->SM bandwidth via LDG and via TMA bulk
// copies, TMEM load/store bandwidth).  Not part of the product.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench tools/microbench.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---------------------------------------------------------------- DFMA
template <int ILP>
__global__ void k_dfma(double* out, int iters, double a, double b) {
    double x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x + i;
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < ILP; ++i) x[i] = fma(x[i], a, b);
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i];
    if (s == 1.2345) out[0] = s;
}

// ---------------------------------------------------------------- LDS (+ optional SHFL / DFMA mix)
// MODE 0: LDS.128 only; 1: LDS.64 only; 2: LDS.128 + 2 SHFL.32 per load; 3: SHFL only (4 per iter);
// 4: LDS.128 + 4 DFMA; 5: LDS.128 + 8 DFMA; 6: LDS.128 + 2 DFMA
template <int MODE>
__global__ void k_lds(double* out, int iters, long long* clk) {
    extern __shared__ __align__(16) double sm[];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = i;
    __syncthreads();
    double acc0 = 0, acc1 = 0, acc2 = 1, acc3 = 2, acc4 = 3, acc5 = 4, acc6 = 5, acc7 = 6;
    int idx = threadIdx.x * 2;
    unsigned u = threadIdx.x;
    long long t0 = clock64();
#pragma unroll 4
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0 || MODE == 2 || MODE >= 4) {
            double2 v = *reinterpret_cast<const double2*>(sm + ((idx + it * 64) & 4094));
            acc0 += v.x; acc1 += v.y;
        }
        if (MODE == 1) {
            double v = sm[((idx >> 1) + it * 32) & 4095];
            acc0 += v;
        }
        if (MODE == 7) {   // lanes 2i, 2i+1 read the same double: 16 distinct addresses (128 B) per warp
            double v = sm[((threadIdx.x >> 1) + it * 16) & 4095];
            acc0 += v;
        }
        if (MODE == 8) {   // lanes i, i+16 read the same double (pairs across the half-warps)
            double v = sm[((threadIdx.x & 15) + 16 * (threadIdx.x >> 5) + it * 16) & 4095];
            acc0 += v;
        }
        if (MODE == 2) {
            u = __shfl_xor_sync(0xffffffffu, u, 1) + 1;
            u = __shfl_xor_sync(0xffffffffu, u, 2) + 1;
        }
        if (MODE == 3) {
            u = __shfl_xor_sync(0xffffffffu, u, 1) + 1;
            u = __shfl_xor_sync(0xffffffffu, u, 2) + 1;
            u = __shfl_xor_sync(0xffffffffu, u, 4) + 1;
            u = __shfl_xor_sync(0xffffffffu, u, 8) + 1;
        }
        if (MODE == 4 || MODE == 5 || MODE == 6) {
            acc2 = fma(acc2, 1.0000001, acc0); acc3 = fma(acc3, 1.0000001, acc1);
            if (MODE != 6) { acc4 = fma(acc4, 1.0000001, acc0); acc5 = fma(acc5, 1.0000001, acc1); }
            if (MODE == 5) {
                acc6 = fma(acc6, 1.0000001, acc0); acc7 = fma(acc7, 1.0000001, acc1);
                acc2 = fma(acc2, 1.0000002, acc1); acc3 = fma(acc3, 1.0000002, acc0);
            }
        }
    }
    long long t1 = clock64();
    double s = acc0 + acc1 + acc2 + acc3 + acc4 + acc5 + acc6 + acc7 + (double)u;
    if (s == 1.2345) out[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) clk[0] = t1 - t0;
}

// ---------------------------------------------------------------- L2 / HBM read through LDG.128
__global__ void k_ldg(const int4* __restrict__ buf, size_t n16, int reps, int* out) {
    int4 acc = make_int4(0, 0, 0, 0);
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (int r = 0; r < reps; ++r) {
        size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll 8
        for (; i < n16; i += stride) {
            int4 v = __ldcg(buf + i);
            acc.x ^= v.x; acc.y ^= v.y; acc.z ^= v.z; acc.w ^= v.w;
        }
    }
    if ((acc.x ^ acc.y ^ acc.z ^ acc.w) == 0x12345678) out[0] = acc.x;
}

// ---------------------------------------------------------------- L2 read through TMA bulk copies
template <int STAGES, int CHUNK>
__global__ void k_bulk(const unsigned char* __restrict__ buf, size_t nchunks, int reps, int* out) {
    extern __shared__ __align__(128) unsigned char smb[];
    __shared__ __align__(8) unsigned long long bars[STAGES];
    if (threadIdx.x == 0) {
        for (int i = 0; i < STAGES; ++i)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bars[i])), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    // one thread: keep STAGES copies in flight, never read the data (pure L2->SM transfer rate)
    size_t total = 0;
    for (size_t c = blockIdx.x; c < nchunks; c += gridDim.x) ++total;
    total *= reps;
    size_t issued = 0, done = 0;
    size_t c = blockIdx.x;
    uint32_t phase[STAGES];
    for (int i = 0; i < STAGES; ++i) phase[i] = 0;
    while (done < total) {
        while (issued < total && issued - done < STAGES) {
            int s = issued % STAGES;
            uint32_t b = smem_u32(&bars[s]);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(CHUNK) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(smem_u32(smb + (size_t)s * CHUNK)), "l"(buf + c * CHUNK), "r"(CHUNK), "r"(b) : "memory");
            ++issued;
            c += gridDim.x;
            if (c >= nchunks) c = blockIdx.x;
        }
        int s = done % STAGES;
        uint32_t b = smem_u32(&bars[s]);
        asm volatile(
            "{\n.reg .pred P1;\nLW:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DN;\nbra LW;\nDN:\n}\n"
            ::"r"(b), "r"(phase[s]) : "memory");
        phase[s] ^= 1;
        ++done;
    }
    if (smb[0] == 77 && smb[1] == 99) out[0] = 1;
}

// ---------------------------------------------------------------- TMEM ld / st
// each warp owns its 32 TMEM lanes; MODE 0: ld only, 1: st only, 2: ld + DFMA + st (accumulate)
template <int MODE>
__global__ void k_tmem(double* out, int iters, long long* clk) {
    __shared__ uint32_t tmem_base_s;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t base = tmem_base_s;
    // lane field = bits 31..16; this warp's lanes start at 32 * (warp % 4); warps 4..7 use columns 256..
    const uint32_t addr0 = base + (((uint32_t)(warp & 3) * 32u) << 16) + (uint32_t)(warp >> 2) * 256u;
    uint32_t r[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) r[i] = threadIdx.x + i;
    // initialise the columns this warp touches
    for (int c = 0; c < 256; c += 16) {
        asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                     ::"r"(addr0 + c), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
                       "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    double acc = 0;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        const uint32_t a = addr0 + (uint32_t)((it * 16) & 255);
        if (MODE == 0 || MODE == 2) {
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                         : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                           "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                         : "r"(a) : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        }
        if (MODE == 2) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                double d = __hiloint2double(r[2 * i + 1], r[2 * i]);
                d = fma(d, 1.0000001, 0.5);
                r[2 * i] = __double2loint(d); r[2 * i + 1] = __double2hiint(d);
            }
        }
        if (MODE == 0) acc += (double)r[0];
        if (MODE == 1 || MODE == 2) {
            asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                         ::"r"(a), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
                           "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
            if (MODE == 2) asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        }
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    long long t1 = clock64();
    if (acc == 1.2345) out[0] = acc + r[3];
    if (threadIdx.x == 0 && blockIdx.x == 0) clk[0] = t1 - t0;
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(512));
}

// single-warp dependent chains: MODE 0 = ld,wait (latency of a load), 1 = ld,wait,fma,st,wait::st (full RMW round trip),
// 2 = ld,wait,fma,st without wait::st (same address every iteration)
template <int MODE>
__global__ void k_tmem_lat(double* out, int iters, long long* clk) {
    __shared__ uint32_t tmem_base_s;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t base = tmem_base_s;
    const uint32_t a = base + (((uint32_t)(warp & 3) * 32u) << 16);
    uint32_t r[8];
    for (int i = 0; i < 8; ++i) r[i] = 0;
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 ::"r"(a), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(a + (r[0] & 0)) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (MODE >= 1) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                double d = __hiloint2double(r[2 * i + 1], r[2 * i]);
                d = fma(d, 1.0, 1.0);
                r[2 * i] = __double2loint(d); r[2 * i + 1] = __double2hiint(d);
            }
            asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                         ::"r"(a), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
            if (MODE == 1) asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        }
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    long long t1 = clock64();
    // read back: with MODE 1/2 every iteration added 1.0 -> value must equal iters if ld-after-st is ordered
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(a) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    if (threadIdx.x == 0 && blockIdx.x == 0) { clk[0] = t1 - t0; out[0] = __hiloint2double(r[1], r[0]); }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(512));
}

static float timeit(void (*launch)(void*), void* arg, int reps = 3) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    launch(arg); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int i = 0; i < reps; ++i) {
        cudaEventRecord(a); launch(arg); cudaEventRecord(b); CK(cudaEventSynchronize(b));
        float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
    }
    return best;
}

int main(int argc, char** argv) {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    const int nsm = p.multiProcessorCount;
    printf("device %s, %d SMs, L2 %d MB\n", p.name, nsm, p.l2CacheSize >> 20);
    double* dout; CK(cudaMalloc(&dout, 1024));
    long long* dclk; CK(cudaMalloc(&dclk, 64));
    int* iout; CK(cudaMalloc(&iout, 64));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms;

    // ---- DFMA
    for (int warps : {4, 8, 16, 32}) {
        const int iters = 20000;
        k_dfma<8><<<nsm, warps * 32>>>(dout, 100, 1.0000001, 0.5); CK(cudaDeviceSynchronize());
        cudaEventRecord(e0); k_dfma<8><<<nsm, warps * 32>>>(dout, iters, 1.0000001, 0.5); cudaEventRecord(e1);
        CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);
        double fl = 2.0 * 8 * iters * warps * 32.0 * nsm;
        printf("DFMA  ILP8  %2d warps/SM: %.2f TFLOP/s (%.3f ms)\n", warps, fl / ms * 1e-9, ms);
    }
    {
        auto runilp = [&](auto kern, int ilp, int warps) {
            const int iters = 20000;
            kern<<<nsm, warps * 32>>>(dout, 100, 1.0000001, 0.5); CK(cudaDeviceSynchronize());
            cudaEventRecord(e0); kern<<<nsm, warps * 32>>>(dout, iters, 1.0000001, 0.5); cudaEventRecord(e1);
            CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);
            double fl = 2.0 * ilp * iters * warps * 32.0 * nsm;
            printf("DFMA  ILP%d  %2d warps/SM: %.2f TFLOP/s\n", ilp, warps, fl / ms * 1e-9);
        };
        for (int warps : {4, 8}) {
            runilp(k_dfma<1>, 1, warps); runilp(k_dfma<2>, 2, warps); runilp(k_dfma<3>, 3, warps);
            runilp(k_dfma<4>, 4, warps); runilp(k_dfma<6>, 6, warps); runilp(k_dfma<12>, 12, warps);
        }
    }
    // ---- LDS / SHFL
    {
        const int iters = 20000;
        auto run = [&](auto kern, const char* name, double bytes_per_iter_thread, double shfl_per_iter, double dfma_per_iter, int warps) {
            kern<<<nsm, warps * 32, 32768>>>(dout, 100, dclk); CK(cudaDeviceSynchronize());
            cudaEventRecord(e0); kern<<<nsm, warps * 32, 32768>>>(dout, iters, dclk); cudaEventRecord(e1);
            CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);
            long long clk; CK(cudaMemcpy(&clk, dclk, 8, cudaMemcpyDeviceToHost));
            double per_clk = (double)iters * warps * 32 / (double)clk;   // thread-iterations per clk per SM
            printf("%-28s %2d warps: %.3f ms, %lld clk, LDS %.1f B/clk/SM, SHFL %.1f lanes/clk/SM, DFMA %.1f /clk/SM\n", name, warps, ms, clk,
                   per_clk * bytes_per_iter_thread, per_clk * shfl_per_iter, per_clk * dfma_per_iter);
        };
        for (int warps : {8, 16, 32}) {
            run(k_lds<0>, "LDS.128", 16, 0, 0, warps);
            run(k_lds<1>, "LDS.64", 8, 0, 0, warps);
            run(k_lds<7>, "LDS.64 lanes paired (2i,2i+1)", 8, 0, 0, warps);
            run(k_lds<8>, "LDS.64 lanes paired (i,i+16)", 8, 0, 0, warps);
            run(k_lds<2>, "LDS.128 + 2 SHFL", 16, 2, 0, warps);
            run(k_lds<3>, "4 SHFL", 0, 4, 0, warps);
            run(k_lds<6>, "LDS.128 + 2 DFMA", 16, 0, 2, warps);
            run(k_lds<4>, "LDS.128 + 4 DFMA", 16, 0, 4, warps);
            run(k_lds<5>, "LDS.128 + 8 DFMA", 16, 0, 8, warps);
        }
    }
    // ---- LDG from L2 / HBM
    {
        size_t big = (size_t)4 << 30;
        int4* buf; CK(cudaMalloc(&buf, big)); CK(cudaMemset(buf, 1, big));
        for (size_t mb : {16, 32, 64, 96, 4096}) {
            size_t n16 = (mb << 20) / 16;
            int reps = mb >= 4096 ? 2 : (int)(8192 / mb);
            for (int bps : {2, 4, 8}) {
                k_ldg<<<nsm * bps, 512>>>(buf, n16, 1, iout); CK(cudaDeviceSynchronize());
                cudaEventRecord(e0); k_ldg<<<nsm * bps, 512>>>(buf, n16, reps, iout); cudaEventRecord(e1);
                CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);
                printf("LDG.128 %4zu MB x%4d, %d CTA/SM x512: %.1f GB/s\n", mb, reps, bps, (double)(mb << 20) * reps / ms * 1e-6);
            }
        }
        // ---- TMA bulk copies
        for (size_t mb : {32, 64, 4096}) {
            int reps = mb >= 4096 ? 2 : (int)(8192 / mb);
            {
                constexpr int ST = 8, CH = 8192;
                auto k = k_bulk<ST, CH>;
                CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, ST * CH));
                size_t nch = (mb << 20) / CH;
                for (int bps : {1, 2, 3}) {
                    k<<<nsm * bps, 32, ST * CH>>>((unsigned char*)buf, nch, 1, iout); CK(cudaDeviceSynchronize());
                    cudaEventRecord(e0); k<<<nsm * bps, 32, ST * CH>>>((unsigned char*)buf, nch, reps, iout); cudaEventRecord(e1);
                    CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);
                    printf("BULK 8KBx8  %4zu MB x%4d, %d CTA/SM: %.1f GB/s\n", mb, reps, bps, (double)(mb << 20) * reps / ms * 1e-6);
                }
            }
            {
                constexpr int ST = 16, CH = 4096;
                auto k = k_bulk<ST, CH>;
                CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, ST * CH));
                size_t nch = (mb << 20) / CH;
                for (int bps : {1, 2, 3}) {
                    k<<<nsm * bps, 32, ST * CH>>>((unsigned char*)buf, nch, 1, iout); CK(cudaDeviceSynchronize());
                    cudaEventRecord(e0); k<<<nsm * bps, 32, ST * CH>>>((unsigned char*)buf, nch, reps, iout); cudaEventRecord(e1);
                    CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);
                    printf("BULK 4KBx16 %4zu MB x%4d, %d CTA/SM: %.1f GB/s\n", mb, reps, bps, (double)(mb << 20) * reps / ms * 1e-6);
                }
            }
        }
        cudaFree(buf);
    }
    // ---- TMEM
    {
        const int iters = 20000;
        auto run = [&](auto kern, const char* name, double bytes, int warps) {
            kern<<<nsm, warps * 32>>>(dout, 100, dclk); CK(cudaDeviceSynchronize());
            cudaEventRecord(e0); kern<<<nsm, warps * 32>>>(dout, iters, dclk); cudaEventRecord(e1);
            CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);
            long long clk; CK(cudaMemcpy(&clk, dclk, 8, cudaMemcpyDeviceToHost));
            printf("%-22s %d warps: %.3f ms, %lld clk, %.1f B/clk/SM (each way)\n", name, warps, ms, clk,
                   (double)iters * warps * 32 * bytes / (double)clk);
        };
        for (int warps : {4, 8}) {
            run(k_tmem<0>, "TMEM ld x16 (+wait)", 64, warps);
            run(k_tmem<1>, "TMEM st x16", 64, warps);
            run(k_tmem<2>, "TMEM ld+8DFMA+st", 64, warps);
        }
    }
    {
        const int iters = 10000;
        auto run = [&](auto kern, const char* name, int warps) {
            kern<<<1, warps * 32>>>(dout, iters, dclk); CK(cudaDeviceSynchronize());
            long long clk; double v; CK(cudaMemcpy(&clk, dclk, 8, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&v, dout, 8, cudaMemcpyDeviceToHost));
            printf("%-40s %d warps: %.1f clk/iter, final value %.1f (expect %d when ordered)\n", name, warps, (double)clk / iters, v, iters);
        };
        run(k_tmem_lat<0>, "TMEM ld.x8 + wait (latency)", 1);
        run(k_tmem_lat<1>, "TMEM ld,wait,fma,st,wait::st", 1);
        run(k_tmem_lat<2>, "TMEM ld,wait,fma,st (no wait::st)", 1);
        run(k_tmem_lat<1>, "TMEM ld,wait,fma,st,wait::st", 4);
        run(k_tmem_lat<2>, "TMEM ld,wait,fma,st (no wait::st)", 4);
    }
    printf("done\n");
    return 0;
}
