// Integer-pipe micro-benchmark for the roofline denominator (the MAC32 peak is not
// in MEASURED_PEAKS.json).  Measures, on all SMs:
//   imad      : independent 32-bit IMAD (lo) chains
//   imad_wide : independent mad.wide.u32 (IMAD.WIDE) chains
//   cc_pair   : mad.lo.cc / madc.hi.cc carry chains (what field.cuh emits -> IMAD.WIDE.X)
//   fp_mul / fr_mul : back-to-back Montgomery products (achieved MAC32/s at 300 / 136 per mul)
// Prints one JSON object.
#include <cstdio>
#include <cuda_runtime.h>
#include "../bellman_mpc_b200/csrc/field.cuh"
using namespace bmpc;

#define ITERS 4096
#define CHAINS 8

__global__ void k_imad(uint32_t* out, uint32_t a, uint32_t b) {
    uint32_t x[CHAINS];
    for (int j = 0; j < CHAINS; j++) x[j] = threadIdx.x + j;
    for (int i = 0; i < ITERS; i++)
#pragma unroll
        for (int j = 0; j < CHAINS; j++) x[j] = x[j] * a + b;
    uint32_t s = 0;
    for (int j = 0; j < CHAINS; j++) s ^= x[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_imad_wide(uint64_t* out, uint32_t a, uint32_t b) {
    uint64_t x[CHAINS];
    for (int j = 0; j < CHAINS; j++) x[j] = threadIdx.x + j;
    for (int i = 0; i < ITERS; i++)
#pragma unroll
        for (int j = 0; j < CHAINS; j++)
            asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(x[j]) : "r"((uint32_t)x[j] ^ a), "r"(b));
    uint64_t s = 0;
    for (int j = 0; j < CHAINS; j++) s ^= x[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_cc_pair(uint32_t* out, uint32_t a, uint32_t b) {
    uint32_t lo[CHAINS], hi[CHAINS];
    for (int j = 0; j < CHAINS; j++) { lo[j] = threadIdx.x + j; hi[j] = j; }
    for (int i = 0; i < ITERS; i++) {
        lo[0] = mad_lo_cc(a, b, lo[0]);
        hi[0] = madc_hi_cc(a, b, hi[0]);
#pragma unroll
        for (int j = 1; j < CHAINS; j++) {
            lo[j] = madc_lo_cc(a + j, b, lo[j]);
            hi[j] = madc_hi_cc(a + j, b, hi[j]);
        }
        hi[CHAINS - 1] = addc(hi[CHAINS - 1], 0);
    }
    uint32_t s = 0;
    for (int j = 0; j < CHAINS; j++) s ^= lo[j] ^ hi[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <class F, int NCH>
__global__ void k_field_mul(F* out, const F* in, int iters) {
    F x[NCH];
    F y = in[1];
    for (int j = 0; j < NCH; j++) x[j] = in[0 + (j & 1)];
    for (int i = 0; i < iters; i++)
#pragma unroll
        for (int j = 0; j < NCH; j++) x[j] = x[j] * y;
    F s = x[0];
    for (int j = 1; j < NCH; j++) s = s + x[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class K>
static double time_ms(K launch) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch(); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; r++) {
        cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    int sms = prop.multiProcessorCount;
    int blocks = sms * 8, threads = 256;
    size_t nthr = (size_t)blocks * threads;
    void* out; cudaMalloc(&out, nthr * 64);
    Fp* fpin; cudaMalloc(&fpin, 2 * sizeof(Fp));
    Fp h[2]; h[0] = Fp::one(); h[1] = Fp::r2();
    cudaMemcpy(fpin, h, sizeof(h), cudaMemcpyHostToDevice);
    Fr* frin; cudaMalloc(&frin, 2 * sizeof(Fr));
    Fr g[2]; g[0] = Fr::one(); g[1] = Fr::r2();
    cudaMemcpy(frin, g, sizeof(g), cudaMemcpyHostToDevice);
    double t_imad = time_ms([&] { k_imad<<<blocks, threads>>>((uint32_t*)out, 3, 5); });
    double t_wide = time_ms([&] { k_imad_wide<<<blocks, threads>>>((uint64_t*)out, 3, 5); });
    double t_cc = time_ms([&] { k_cc_pair<<<blocks, threads>>>((uint32_t*)out, 3, 5); });
    const int FI = 256;
    double t_fp1 = time_ms([&] { k_field_mul<Fp, 1><<<blocks, 128>>>((Fp*)out, fpin, FI); });
    double t_fp2 = time_ms([&] { k_field_mul<Fp, 2><<<blocks, 128>>>((Fp*)out, fpin, FI); });
    double t_fr2 = time_ms([&] { k_field_mul<Fr, 2><<<blocks, 256>>>((Fr*)out, frin, FI); });
    double ops = (double)nthr * ITERS * CHAINS;
    double fp1 = (double)blocks * 128 * FI * 1 / (t_fp1 * 1e-3);
    double fp2 = (double)blocks * 128 * FI * 2 / (t_fp2 * 1e-3);
    double fr2 = (double)blocks * 256 * FI * 2 / (t_fr2 * 1e-3);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("{\"sms\": %d, \"clock_khz\": %d, \"imad_per_s\": %.4g, \"imad_wide_per_s\": %.4g, "
           "\"cc_pair_mac32_per_s\": %.4g, \"fp_mul_per_s_1chain\": %.4g, \"fp_mul_per_s_2chain\": %.4g, "
           "\"fr_mul_per_s_2chain\": %.4g, \"fp_mac32_per_s\": %.4g, \"fr_mac32_per_s\": %.4g}\n",
           sms, clk, ops / (t_imad * 1e-3), ops / (t_wide * 1e-3), ops / (t_cc * 1e-3), fp1, fp2, fr2,
           (fp1 > fp2 ? fp1 : fp2) * 300, fr2 * 136);
    return 0;
}
