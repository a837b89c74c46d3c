// tools/fp64_probe.cu -- FP64 issue-rate experiments on B200 (not part of the product).
// Measures the non-fused DMUL/DADD rate of the K=4 Poisson-binomial update with and without the
// companion instructions of the real kernel (PRMT address + shared-memory table lookup), at several
// occupancies.  Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o build/fp64_probe tools/fp64_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int VARIANT, int CHAINS, int K = 4>
__global__ void probe(int iters, double *sink, const uint32_t *words)
{
    extern __shared__ __align__(16) uint8_t smem[];
    // 64 KB table at a 64 KB aligned shared address
    uint32_t base = (uint32_t)__cvta_generic_to_shared(smem);
    uint32_t lut = (base + 0xFFFFu) & ~0xFFFFu;
    double2 *t = reinterpret_cast<double2 *>(smem + (lut - base));
    for (int i = threadIdx.x; i < 256 * 16; i += blockDim.x) {
        double p = 1e-4 * (1 + (i >> 4) % 40);
        t[i] = make_double2(1.0 - p, p);
    }
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t lut_lane = lut + ((VARIANT == 3) ? lane * 8 : (lane & 15) * 16);
    double P[CHAINS][K];
#pragma unroll
    for (int c = 0; c < CHAINS; c++)
#pragma unroll
        for (int j = 0; j < K; j++) P[c][j] = 1.0 / (1.0 + threadIdx.x + c + j);
    uint32_t w = words[threadIdx.x & 255];
    double q = 0.9999, e = 1e-4;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int c = 0; c < CHAINS; c++) {
#pragma unroll
            for (int b = 0; b < 4; b++) {
                if (VARIANT == 1) {            // + PRMT only (result feeds nothing expensive)
                    uint32_t addr = __byte_perm(w, lut_lane, 0x7604u | (b << 4));
                    w ^= addr & 0x100u;
                } else if (VARIANT == 2) {     // + PRMT + LDS.128 (q, e)
                    uint32_t addr = __byte_perm(w, lut_lane, 0x7604u | (b << 4));
                    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(q), "=d"(e) : "r"(addr));
                } else if (VARIANT == 3) {     // + PRMT + LDS.64 (p) + DSUB
                    uint32_t addr = __byte_perm(w, lut_lane, 0x7604u | (b << 4));
                    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(e) : "r"(addr));
                    q = __dsub_rn(1.0, e);
                }
#pragma unroll
                for (int j = K - 1; j >= 1; j--) P[c][j] = __dadd_rn(__dmul_rn(q, P[c][j]), __dmul_rn(e, P[c][j - 1]));
                P[c][0] = __dmul_rn(q, P[c][0]);
            }
        }
        w = w * 1664525u + 1013904223u;
        w &= 0x27272727u;   // keep "qualities" < 40
    }
    double s = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; c++)
#pragma unroll
        for (int j = 0; j < K; j++) s += P[c][j];
    if (s == 123.456) sink[0] = s;
}

template <int VARIANT, int CHAINS, int K = 4>
void run(const char *name, int threads, int blocks_per_sm, int sms, double *sink, uint32_t *words)
{
    const int iters = 20000 / CHAINS;
    const int smem = 2 * 65536 + 1024;
    cudaFuncSetAttribute(probe<VARIANT, CHAINS, K>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    probe<VARIANT, CHAINS, K><<<sms * blocks_per_sm, threads, smem>>>(iters / 10, sink, words);
    cudaEventRecord(a);
    probe<VARIANT, CHAINS, K><<<sms * blocks_per_sm, threads, smem>>>(iters, sink, words);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    cudaError_t err = cudaGetLastError();
    double fp64 = (double)sms * blocks_per_sm * threads * (double)iters * CHAINS * 4 * (3 * K - 2 + (VARIANT == 3 ? 1 : 0));
    double bases = (double)sms * blocks_per_sm * threads * (double)iters * CHAINS * 4;
    printf("K=%d %-28s warps/SM %2d chains %d : %7.3f ms  %6.2f TFP64op/s  %6.2f Gbase/s  (%s)\n", K, name,
           threads / 32 * blocks_per_sm, CHAINS, ms, fp64 / ms / 1e9, bases / ms / 1e6, cudaGetErrorString(err));
}

int main()
{
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    int sms = prop.multiProcessorCount;
    double *sink; uint32_t *words;
    cudaMalloc(&sink, 64);
    cudaMalloc(&words, 1024);
    uint32_t h[256];
    for (int i = 0; i < 256; i++) h[i] = (i * 2654435761u) & 0x27272727u;
    cudaMemcpy(words, h, 1024, cudaMemcpyHostToDevice);
    printf("%s, %d SMs\n", prop.name, sms);
    for (int threads : {256, 512, 768, 1024}) {
        run<0, 1>("fp64 only", threads, 1, sms, sink, words);
        run<1, 1>("fp64 + PRMT", threads, 1, sms, sink, words);
        run<2, 1>("fp64 + PRMT + LDS.128", threads, 1, sms, sink, words);
        run<3, 1>("fp64 + PRMT + LDS.64 + DSUB", threads, 1, sms, sink, words);
    }
    run<0, 2>("fp64 only", 512, 1, sms, sink, words);
    run<2, 2>("fp64 + PRMT + LDS.128", 512, 1, sms, sink, words);
    run<3, 2>("fp64 + PRMT + LDS.64 + DSUB", 512, 1, sms, sink, words);
    run<2, 2>("fp64 + PRMT + LDS.128", 256, 1, sms, sink, words);
    // the two-entry sweep of the cascade (K = 2)
    for (int threads : {512, 768, 1024}) {
        run<0, 1, 2>("fp64 only", threads, 1, sms, sink, words);
        run<3, 1, 2>("fp64 + PRMT + LDS.64 + DSUB", threads, 1, sms, sink, words);
        run<2, 1, 2>("fp64 + PRMT + LDS.128", threads, 1, sms, sink, words);
    }
    run<0, 2, 2>("fp64 only", 512, 1, sms, sink, words);
    run<3, 2, 2>("fp64 + PRMT + LDS.64 + DSUB", 512, 1, sms, sink, words);
    run<3, 4, 2>("fp64 + PRMT + LDS.64 + DSUB", 512, 1, sms, sink, words);
    run<2, 2, 2>("fp64 + PRMT + LDS.128", 512, 1, sms, sink, words);
    return 0;
}
