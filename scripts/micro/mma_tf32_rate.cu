// Microbenchmark: issue rate of mma.sync.m16n8k8 TF32 (legacy warp-level MMA) on sm_100a, per SM.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mma_rate scripts/micro/mma_tf32_rate.cu && /tmp/mma_rate
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void mma_tf32(float (&d)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

template <int NACC>
__global__ void k(float *out, int iters) {
    float acc[NACC][4];
    unsigned a[4], b[2];
    for (int i = 0; i < 4; ++i) a[i] = __float_as_uint(1.0f + threadIdx.x * 1e-3f + i);
    for (int i = 0; i < 2; ++i) b[i] = __float_as_uint(0.5f + threadIdx.x * 1e-3f + i);
    for (int j = 0; j < NACC; ++j)
        for (int i = 0; i < 4; ++i) acc[j][i] = 0.f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < NACC; ++j) mma_tf32(acc[j], a, b);
    }
    float s = 0.f;
    for (int j = 0; j < NACC; ++j)
        for (int i = 0; i < 4; ++i) s += acc[j][i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
void run(int warps_per_sm) {
    int sms = 148, iters = 4096;
    float *out;
    cudaMalloc(&out, sizeof(float) * sms * warps_per_sm * 32);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k<NACC><<<sms, warps_per_sm * 32>>>(out, 16);
    cudaEventRecord(e0);
    k<NACC><<<sms, warps_per_sm * 32>>>(out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double mmas = (double)sms * warps_per_sm * iters * NACC;
    int clk_khz;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    double clks = ms * 1e-3 * clk_khz * 1e3;
    printf("acc=%d warps/SM=%2d: %.3f ms, %.3f MMA/clk/SM (%.1f clk per MMA per SM), %.1f TFLOP/s tf32\n", NACC, warps_per_sm, ms,
           mmas / sms / clks, clks / (mmas / sms), mmas * 2 * 16 * 8 * 8 / (ms * 1e-3) / 1e12);
    cudaFree(out);
}

int main() {
    run<1>(4); run<4>(4); run<8>(4); run<4>(8); run<4>(16); run<8>(16); run<4>(24);
    return 0;
}
