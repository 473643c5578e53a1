// camera.cu -- row a7 of SURVEY.md section 8: N interpolated camera sub-exposure deltas from the
// two se(3) vectors the MoveModel heads emit (see camera_math.cuh).  O(N) work: one thread per
// sub-exposure; the reference spends ~100 tiny torch/pypose launches on it per render call.
#include "camera_math.cuh"
#include "common.cuh"

namespace d4 {

__device__ __forceinline__ float interp_u(int n, int N) {
    // torch.linspace(0, 1, N)[n] in fp32: start + n * step for the first half, end - (N-1-n) * step after
    if (N <= 1) return 0.f;
    const float step = 1.0f / (float)(N - 1);
    return (n < N / 2) ? (float)n * step : 1.0f - (float)(N - 1 - n) * step;
}

__global__ void camera_interp_fwd_kernel(const float *__restrict__ start6, const float *__restrict__ end6, int N,
                                         float *__restrict__ RTs) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    float s[6], e[6], Rt[12];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        s[i] = __ldg(start6 + i);
        e[i] = __ldg(end6 + i);
    }
    camera_interp_one<float>(s, e, interp_u(n, N), Rt);
#pragma unroll
    for (int i = 0; i < 12; ++i) RTs[12 * n + i] = Rt[i];
}

__global__ void camera_interp_bwd_kernel(const float *__restrict__ start6, const float *__restrict__ end6, int N,
                                         const float *__restrict__ v_RTs, float *__restrict__ v_start6,
                                         float *__restrict__ v_end6) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    typedef Dual<12> DU;
    DU s[6], e[6], Rt[12];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        s[i].v = __ldg(start6 + i);
        e[i].v = __ldg(end6 + i);
#pragma unroll
        for (int j = 0; j < 12; ++j) {
            s[i].d[j] = (j == i) ? 1.f : 0.f;
            e[i].d[j] = (j == 6 + i) ? 1.f : 0.f;
        }
    }
    camera_interp_one<DU>(s, e, interp_u(n, N), Rt);
    float g[12];
#pragma unroll
    for (int j = 0; j < 12; ++j) g[j] = 0.f;
#pragma unroll
    for (int i = 0; i < 12; ++i) {
        const float v = __ldg(v_RTs + 12 * n + i);
#pragma unroll
        for (int j = 0; j < 12; ++j) g[j] = fmaf(v, Rt[i].d[j], g[j]);
    }
#pragma unroll
    for (int j = 0; j < 6; ++j) {
        atomicAdd(v_start6 + j, g[j]);
        atomicAdd(v_end6 + j, g[6 + j]);
    }
}

}  // namespace d4

using namespace d4;

extern "C" int d4_camera_interp_fwd(const float *start6, const float *end6, int N, float *RTs, d4_stream_t stream) {
    D4_CHECK_ARG(start6 && end6 && RTs && N >= 1, "d4_camera_interp_fwd: bad arguments");
    camera_interp_fwd_kernel<<<cdiv(N, 32), 32, 0, as_stream(stream)>>>(start6, end6, N, RTs);
    D4_CHECK_LAUNCH("d4_camera_interp_fwd");
    return 0;
}

extern "C" int d4_camera_interp_bwd(const float *start6, const float *end6, int N, const float *v_RTs,
                                    float *v_start6, float *v_end6, d4_stream_t stream) {
    D4_CHECK_ARG(start6 && end6 && v_RTs && v_start6 && v_end6 && N >= 1, "d4_camera_interp_bwd: bad arguments");
    camera_interp_bwd_kernel<<<cdiv(N, 32), 32, 0, as_stream(stream)>>>(start6, end6, N, v_RTs, v_start6, v_end6);
    D4_CHECK_LAUNCH("d4_camera_interp_bwd");
    return 0;
}
