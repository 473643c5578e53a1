// combine.cu -- row a13 of SURVEY.md section 8: the N-way combine of the
// sub-exposure renders (flow3d/scene_model.py:386-397).  The reference builds
// three torch.stack copies of the N images and reduces each (mean / max / min):
// ~4 N P D floats of HBM traffic.  Here every input element is read exactly once
// and the blurry image written once: N*P*(D+1)*4 B read + P*(D+1)*4 B written.
#include "common.cuh"

namespace d4 {

// element e in [0, P*D): channel d = e % D; e in [P*D, P*D + P): alpha
__global__ void __launch_bounds__(256)
combine_fwd_kernel(const float *__restrict__ imgs, const float *__restrict__ alphas, int N, int64_t P, int D,
                   int max_ch, int min_ch, int ref_quirk, float *__restrict__ out_img,
                   float *__restrict__ out_alpha) {
    const int64_t PD = P * D;
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= PD + P) return;
    const bool is_alpha = e >= PD;
    const float *src = is_alpha ? alphas + (e - PD) : imgs + e;
    const int64_t stride = is_alpha ? P : PD;
    const int d = is_alpha ? -1 : (int)(e % D);
    const bool is_max = (d >= 0 && d == max_ch), is_min = (d >= 0 && d == min_ch);
    float sum = 0.f, mx = -INFINITY, mn = INFINITY;
    const int n_ext = ref_quirk ? N - 1 : N;  // extrema over r_0..r_{N-2} (+ mean) in quirk mode
#pragma unroll 4
    for (int n = 0; n < N; ++n) {
        const float v = __ldg(src + n * stride);
        sum += v;
        if (n < n_ext) {
            mx = fmaxf(mx, v);
            mn = fminf(mn, v);
        }
    }
    const float mean = sum / (float)N;
    float r = mean;
    if (is_max) r = ref_quirk ? fmaxf(mx, mean) : mx;
    if (is_min) r = ref_quirk ? fminf(mn, mean) : mn;
    if (is_alpha) out_alpha[e - PD] = mean;
    else out_img[e] = r;
}

__global__ void __launch_bounds__(256)
combine_bwd_kernel(const float *__restrict__ imgs, int N, int64_t P, int D, int max_ch, int min_ch, int ref_quirk,
                   const float *__restrict__ v_out_img, const float *__restrict__ v_out_alpha,
                   float *__restrict__ v_imgs, float *__restrict__ v_alphas) {
    const int64_t PD = P * D;
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= PD + P) return;
    const float invN = 1.0f / (float)N;
    if (e >= PD) {
        const float v = __ldg(v_out_alpha + (e - PD)) * invN;
        for (int n = 0; n < N; ++n) v_alphas[n * P + (e - PD)] = v;
        return;
    }
    const int d = (int)(e % D);
    const float v = __ldg(v_out_img + e);
    const bool is_max = d == max_ch, is_min = d == min_ch;
    if (!is_max && !is_min) {
        const float vn = v * invN;
        for (int n = 0; n < N; ++n) v_imgs[n * PD + e] = vn;
        return;
    }
    // arg-extremum (first occurrence) over r_0..r_{n_ext-1} [, mean]
    const int n_ext = ref_quirk ? N - 1 : N;
    float sum = 0.f, best = is_max ? -INFINITY : INFINITY;
    int arg = -1;
    for (int n = 0; n < N; ++n) {
        const float x = __ldg(imgs + n * PD + e);
        sum += x;
        if (n < n_ext && (is_max ? x > best : x < best)) {
            best = x;
            arg = n;
        }
    }
    const float mean = sum / (float)N;
    const bool mean_wins = ref_quirk && (arg < 0 || (is_max ? mean > best : mean < best));
    for (int n = 0; n < N; ++n) v_imgs[n * PD + e] = mean_wins ? v * invN : (n == arg ? v : 0.f);
}

// Row f3: densification statistics (flow3d/trainer.py:953-990, Trainer._prepare_control_step).
// The reference runs, per sub-exposure render, ~10 launches (where / clone / scale / norm / 2x index_add /
// index_select / maximum / index_put) over [G]; here one thread per Gaussian walks all N renders:
// N*G*12 B read, no atomics, same accumulation order as the reference's loop over ii.
__global__ void __launch_bounds__(256)
densify_stats_kernel(const float *__restrict__ v_means2d, const int32_t *__restrict__ radii, int N, int G, float sx,
                     float sy, float inv_max_wh, float *__restrict__ grad_norm_acc, int64_t *__restrict__ vis_count,
                     float *__restrict__ max_radii) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= G) return;
    float acc = grad_norm_acc[g];
    int64_t cnt = vis_count[g];
    float mr = max_radii ? max_radii[g] : 0.f;
    for (int n = 0; n < N; ++n) {
        const int64_t idx = (int64_t)n * G + g;
        const int32_t r = __ldg(radii + idx);
        if (r <= 0) continue;
        const float2 v = __ldg(reinterpret_cast<const float2 *>(v_means2d) + idx);
        const float gx = v.x * sx, gy = v.y * sy;
        acc += sqrtf(gx * gx + gy * gy);
        cnt += 1;
        mr = fmaxf(mr, (float)r * inv_max_wh);
    }
    grad_norm_acc[g] = acc;
    vis_count[g] = cnt;
    if (max_radii) max_radii[g] = mr;
}

}  // namespace d4

using namespace d4;

extern "C" int d4_densify_stats(const float *v_means2d, const int32_t *radii, int N, int G, float sx, float sy,
                                float inv_max_wh, float *grad_norm_acc, int64_t *vis_count, float *max_radii,
                                d4_stream_t stream) {
    D4_CHECK_ARG(N >= 1 && G >= 0 && grad_norm_acc && vis_count, "d4_densify_stats: bad arguments");
    if (G == 0) return 0;
    D4_CHECK_ARG(v_means2d && radii && ((uintptr_t)v_means2d & 7) == 0, "d4_densify_stats: null/unaligned pointer");
    densify_stats_kernel<<<cdiv(G, 256), 256, 0, as_stream(stream)>>>(v_means2d, radii, N, G, sx, sy, inv_max_wh,
                                                                     grad_norm_acc, vis_count, max_radii);
    D4_CHECK_LAUNCH("d4_densify_stats");
    return 0;
}

extern "C" int d4_combine_fwd(const float *imgs, const float *alphas, int N, int64_t P, int D, int max_ch, int min_ch,
                              int ref_quirk, float *out_img, float *out_alpha, d4_stream_t stream) {
    D4_CHECK_ARG(imgs && alphas && out_img && out_alpha && N >= 1 && P >= 0 && D >= 1, "d4_combine_fwd: bad arguments");
    if (P == 0) return 0;
    combine_fwd_kernel<<<cdiv(P * D + P, 256), 256, 0, as_stream(stream)>>>(imgs, alphas, N, P, D, max_ch, min_ch,
                                                                           ref_quirk, out_img, out_alpha);
    D4_CHECK_LAUNCH("d4_combine_fwd");
    return 0;
}

extern "C" int d4_combine_bwd(const float *imgs, int N, int64_t P, int D, int max_ch, int min_ch, int ref_quirk,
                              const float *v_out_img, const float *v_out_alpha, float *v_imgs, float *v_alphas,
                              d4_stream_t stream) {
    D4_CHECK_ARG(imgs && v_out_img && v_out_alpha && v_imgs && v_alphas && N >= 1 && P >= 0 && D >= 1,
                 "d4_combine_bwd: bad arguments");
    if (P == 0) return 0;
    combine_bwd_kernel<<<cdiv(P * D + P, 256), 256, 0, as_stream(stream)>>>(imgs, N, P, D, max_ch, min_ch, ref_quirk,
                                                                           v_out_img, v_out_alpha, v_imgs, v_alphas);
    D4_CHECK_LAUNCH("d4_combine_bwd");
    return 0;
}
