// correlation.cu -- row f4 of SURVEY.md section 8: the PWC-Net cost volume of the reference's AlignedLoss front-end
// (flow3d/loss_utils.py:161-189 -> flow3d/models/pwcnet.py:179,187 -> FunctionCorrelation).  Replaces the two CuPy
// RawKernels the reference launches per call (flow3d/models/external/pwcnet/correlation/correlation.py):
//   kernel_Correlation_rearrange (:8-33)     NCHW -> zero-padded NHWC copies of both inputs (2 launches, 2 x C*H*W*4 B
//                                            written and re-read)
//   kernel_Correlation_updateOutput (:35-103) one 32-thread block per output PIXEL, 81 displacements one after the
//                                            other, each with a shared-memory reduction and two block barriers
// by one kernel on the NCHW inputs: a CTA owns a 32 x 8 pixel tile, stages the first image's tile and the second
// image's (32 + 8) x (8 + 8) halo tile in shared memory eight channels at a time (zero outside the image, as the
// reference's padding), and every thread keeps the 81 sums of its pixel in registers.  No padded copies, no
// reductions, no barriers besides the two per channel chunk; out[b, (dy+4)*9 + (dx+4), y, x] = <f1[:, y, x],
// f2[:, y+dy, x+dx]> / C.  The backward (correlation_bwd_kernel below) restates kernel_Correlation_updateGradFirst /
// updateGradSecond (:105-233); the reference itself never runs it (PWC-Net sits under torch.no_grad,
// loss_utils.py:171-172), it completes the module for callers that do differentiate through the cost volume.
#include "common.cuh"

namespace d4 {

constexpr int kCorrTW = 32, kCorrTH = 8, kCorrR = 4, kCorrCK = 8;
constexpr int kCorrHW = kCorrTW + 2 * kCorrR, kCorrHH = kCorrTH + 2 * kCorrR;  // halo tile 40 x 16
constexpr int kCorrD = 2 * kCorrR + 1;                                          // 9 displacements per axis

__global__ void __launch_bounds__(kCorrTW *kCorrTH)
correlation_fwd_kernel(const float *__restrict__ first, const float *__restrict__ second, int C, int H, int W,
                       float *__restrict__ out) {
    __shared__ float s_first[kCorrCK][kCorrTH][kCorrTW];
    __shared__ float s_second[kCorrCK][kCorrHH][kCorrHW + 1];
    const int b = blockIdx.z;
    const int x0 = blockIdx.x * kCorrTW, y0 = blockIdx.y * kCorrTH;
    const int tx = threadIdx.x % kCorrTW, ty = threadIdx.x / kCorrTW;
    const int x = x0 + tx, y = y0 + ty;
    const int64_t plane = (int64_t)H * W;
    const float *f1 = first + (int64_t)b * C * plane, *f2 = second + (int64_t)b * C * plane;
    float acc[kCorrD * kCorrD];
#pragma unroll
    for (int k = 0; k < kCorrD * kCorrD; ++k) acc[k] = 0.f;
    for (int c0 = 0; c0 < C; c0 += kCorrCK) {
        const int nc = min(kCorrCK, C - c0);
        // stage: first tile (nc x 8 x 32) and second halo tile (nc x 16 x 40), zero outside the image
        for (int e = threadIdx.x; e < kCorrCK * kCorrTH * kCorrTW; e += kCorrTW * kCorrTH) {
            const int c = e / (kCorrTH * kCorrTW), r = e - c * (kCorrTH * kCorrTW);
            const int yy = y0 + r / kCorrTW, xx = x0 + r % kCorrTW;
            s_first[c][r / kCorrTW][r % kCorrTW] =
                (c < nc && yy < H && xx < W) ? __ldg(f1 + (int64_t)(c0 + c) * plane + (int64_t)yy * W + xx) : 0.f;
        }
        for (int e = threadIdx.x; e < kCorrCK * kCorrHH * kCorrHW; e += kCorrTW * kCorrTH) {
            const int c = e / (kCorrHH * kCorrHW), r = e - c * (kCorrHH * kCorrHW);
            const int hy = r / kCorrHW, hx = r - hy * kCorrHW;
            const int yy = y0 + hy - kCorrR, xx = x0 + hx - kCorrR;
            s_second[c][hy][hx] = (c < nc && yy >= 0 && yy < H && xx >= 0 && xx < W)
                                      ? __ldg(f2 + (int64_t)(c0 + c) * plane + (int64_t)yy * W + xx)
                                      : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int c = 0; c < kCorrCK; ++c) {
            const float a = s_first[c][ty][tx];
#pragma unroll
            for (int dy = 0; dy < kCorrD; ++dy) {
#pragma unroll
                for (int dx = 0; dx < kCorrD; ++dx) acc[dy * kCorrD + dx] = fmaf(a, s_second[c][ty + dy][tx + dx], acc[dy * kCorrD + dx]);
            }
        }
        __syncthreads();
    }
    if (x < W && y < H) {
        const float inv = 1.0f / (float)C;
        float *o = out + (int64_t)b * kCorrD * kCorrD * plane + (int64_t)y * W + x;
#pragma unroll
        for (int k = 0; k < kCorrD * kCorrD; ++k) o[(int64_t)k * plane] = acc[k] * inv;
    }
}

// Backward of the cost volume: kernel_Correlation_updateGradFirst (correlation.py:105-167) and
// kernel_Correlation_updateGradSecond (:169-233), which the reference launches once per SAMPLE (:341-381), in one
// launch for the whole batch.  One thread per (b, c, y, x), x fastest -- the reference's mapping -- and its summation
// order (dy outer, dx inner, fused multiply-adds, one division by C at the end):
//   gradFirst [b,c,y,x] = (1/C) sum_{dy,dx} gradOutput[b, op, y, x]         * second[b, c, y+dy, x+dx]
//   gradSecond[b,c,y,x] = (1/C) sum_{dy,dx} gradOutput[b, op, y-dy, x-dx]   * first [b, c, y-dy, x-dx]
// with op = (dy+4)*9 + (dx+4) and zero outside the image (the reference reads its zero-padded copies; a displacement
// that leaves the image is skipped in gradSecond and multiplies a zero in gradFirst, exactly as there).
// Reads are coalesced along x and served by L1 / L2 (every value is reused by the 81 displacements of its neighbours).
__global__ void __launch_bounds__(256)
correlation_bwd_kernel(const float *__restrict__ first, const float *__restrict__ second,
                       const float *__restrict__ grad_out, int C, int H, int W, float *__restrict__ grad_first,
                       float *__restrict__ grad_second) {
    const int64_t plane = (int64_t)H * W;
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // (c, y, x) of sample b
    if (e >= C * plane) return;
    const int b = blockIdx.z;
    const int c = (int)(e / plane);
    const int r = (int)(e - c * plane);
    const int y = r / W, x = r - y * W;
    const float *f1 = first + ((int64_t)b * C + c) * plane, *f2 = second + ((int64_t)b * C + c) * plane;
    const float *g = grad_out + (int64_t)b * kCorrD * kCorrD * plane;
    float s1 = 0.f, s2 = 0.f;
    for (int dy = -kCorrR; dy <= kCorrR; ++dy) {
        for (int dx = -kCorrR; dx <= kCorrR; ++dx) {
            const int op = (dy + kCorrR) * kCorrD + (dx + kCorrR);
            const float *gp = g + (int64_t)op * plane;
            if (grad_first) {
                const int yy = y + dy, xx = x + dx;
                const float bot1 = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? __ldg(f2 + (int64_t)yy * W + xx) : 0.f;
                s1 = fmaf(__ldg(gp + r), bot1, s1);
            }
            if (grad_second) {
                const int yy = y - dy, xx = x - dx;
                if (yy >= 0 && yy < H && xx >= 0 && xx < W)
                    s2 = fmaf(__ldg(gp + (int64_t)yy * W + xx), __ldg(f1 + (int64_t)yy * W + xx), s2);
            }
        }
    }
    if (grad_first) grad_first[((int64_t)b * C + c) * plane + r] = s1 / (float)C;
    if (grad_second) grad_second[((int64_t)b * C + c) * plane + r] = s2 / (float)C;
}

}  // namespace d4

using namespace d4;

extern "C" int d4_correlation_fwd(const float *first, const float *second, int B, int C, int H, int W, float *out,
                                  d4_stream_t stream) {
    D4_CHECK_ARG(B >= 0 && C >= 1 && H >= 1 && W >= 1 && B <= 65535, "d4_correlation_fwd: bad sizes");
    if (B == 0) return 0;
    D4_CHECK_ARG(first && second && out, "d4_correlation_fwd: null pointer");
    dim3 grid(cdiv(W, kCorrTW), cdiv(H, kCorrTH), B);
    correlation_fwd_kernel<<<grid, kCorrTW * kCorrTH, 0, as_stream(stream)>>>(first, second, C, H, W, out);
    D4_CHECK_LAUNCH("d4_correlation_fwd");
    return 0;
}

extern "C" int d4_correlation_bwd(const float *first, const float *second, const float *grad_out, int B, int C, int H,
                                  int W, float *grad_first, float *grad_second, d4_stream_t stream) {
    D4_CHECK_ARG(B >= 0 && C >= 1 && H >= 1 && W >= 1 && B <= 65535, "d4_correlation_bwd: bad sizes");
    if (B == 0 || (!grad_first && !grad_second)) return 0;
    D4_CHECK_ARG(first && second && grad_out, "d4_correlation_bwd: null pointer");
    dim3 grid(cdiv((int64_t)C * H * W, 256), 1, B);
    correlation_bwd_kernel<<<grid, 256, 0, as_stream(stream)>>>(first, second, grad_out, C, H, W, grad_first, grad_second);
    D4_CHECK_LAUNCH("d4_correlation_bwd");
    return 0;
}
