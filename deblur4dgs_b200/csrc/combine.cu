// combine.cu -- row a13 of SURVEY.md section 8: the N-way combine of the
// sub-exposure renders (flow3d/scene_model.py:386-397).  The reference builds
// three torch.stack copies of the N images and reduces each (mean / max / min):
// ~4 N P D floats of HBM traffic.  Here every input element is read exactly once
// and the blurry image written once: N*P*(D+1)*4 B read + P*(D+1)*4 B written.
#include "common.cuh"

namespace d4 {

// One thread owns VEC consecutive floats of the image ([0, P*D), channel d = e % D) or of the alpha plane
// ([P*D, P*D + P)); VEC = 4 (LDG.128 / STG.128) whenever P*D and P are multiples of 4, else 1.
template <int VEC>
struct Pack {
    float v[VEC];
};
template <int VEC>
__device__ __forceinline__ Pack<VEC> load_pack(const float *p) {
    Pack<VEC> r;
    if constexpr (VEC == 4) {
        const float4 x = __ldg(reinterpret_cast<const float4 *>(p));
        r.v[0] = x.x, r.v[1] = x.y, r.v[2] = x.z, r.v[3] = x.w;
    } else {
        r.v[0] = __ldg(p);
    }
    return r;
}
template <int VEC>
__device__ __forceinline__ void store_pack(float *p, const Pack<VEC> &r) {
    if constexpr (VEC == 4) *reinterpret_cast<float4 *>(p) = make_float4(r.v[0], r.v[1], r.v[2], r.v[3]);
    else *p = r.v[0];
}

// Winner code of the max / min channel of one pixel, written by the forward for the backward:
// n in [0, N) = sub-exposure holding the (first) extremum, kMeanWins = the mean itself (ref_quirk only).
constexpr uint8_t kMeanWins = 255;

template <int VEC>
__global__ void __launch_bounds__(256)
combine_fwd_kernel(const float *__restrict__ imgs, const float *__restrict__ alphas, int N, int64_t P, int D,
                   int max_ch, int min_ch, int ref_quirk, float *__restrict__ out_img,
                   float *__restrict__ out_alpha, uint8_t *__restrict__ arg_max, uint8_t *__restrict__ arg_min) {
    const int64_t PD = P * D;
    const int64_t e = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * VEC;
    if (e >= PD + P) return;
    const bool is_alpha = e >= PD;
    const float *src = is_alpha ? alphas + (e - PD) : imgs + e;
    const int64_t stride = is_alpha ? P : PD;
    const int d0 = is_alpha ? -1 : (int)(e % D);
    const int n_ext = ref_quirk ? N - 1 : N;  // extrema over r_0..r_{N-2} (+ mean) in quirk mode
    Pack<VEC> sum, mx, mn;
#pragma unroll
    for (int k = 0; k < VEC; ++k) sum.v[k] = 0.f, mx.v[k] = -INFINITY, mn.v[k] = INFINITY;
#pragma unroll 4
    for (int n = 0; n < N; ++n) {
        const Pack<VEC> x = load_pack<VEC>(src + n * stride);
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
            sum.v[k] += x.v[k];
            if (n < n_ext) {
                mx.v[k] = fmaxf(mx.v[k], x.v[k]);
                mn.v[k] = fminf(mn.v[k], x.v[k]);
            }
        }
    }
    Pack<VEC> r;
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
        const float mean = sum.v[k] / (float)N;
        const int d = (d0 + k) % D;
        r.v[k] = mean;
        const bool is_max = !is_alpha && d == max_ch, is_min = !is_alpha && d == min_ch;
        if (is_max || is_min) {
            const float best = is_max ? mx.v[k] : mn.v[k];
            const bool mean_wins = ref_quirk && (n_ext == 0 || (is_max ? mean > best : mean < best));
            r.v[k] = mean_wins ? mean : best;
            uint8_t *amap = is_max ? arg_max : arg_min;
            if (amap) {
                // first sub-exposure holding the extremum, as torch.max / min(dim) report it (the N values of this
                // one channel were just read: the re-read is served by L1 / L2).  Measured alternatives at c3:
                // tracking the arg inside the loop 0.189 ms, all N packs kept in registers 0.224 ms, this 0.184 ms
                // (0.128 ms without any winner map -- which would cost the backward 0.2 ms of re-reading).
                int arg = 0;
                for (int n = n_ext - 1; n >= 0; --n)
                    if (__ldg(src + n * stride + k) == best) arg = n;
                amap[(e + k) / D] = mean_wins ? kMeanWins : (uint8_t)arg;
            }
        }
    }
    store_pack<VEC>(is_alpha ? out_alpha + (e - PD) : out_img + e, r);
}

template <int VEC>
__global__ void __launch_bounds__(256)
combine_bwd_kernel(const uint8_t *__restrict__ arg_max, const uint8_t *__restrict__ arg_min, int N, int64_t P, int D,
                   int max_ch, int min_ch, const float *__restrict__ v_out_img, const float *__restrict__ v_out_alpha,
                   float *__restrict__ v_imgs, float *__restrict__ v_alphas) {
    const int64_t PD = P * D;
    const int64_t e = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * VEC;
    if (e >= PD + P) return;
    const float invN = 1.0f / (float)N;
    if (e >= PD) {
        Pack<VEC> v = load_pack<VEC>(v_out_alpha + (e - PD));
#pragma unroll
        for (int k = 0; k < VEC; ++k) v.v[k] *= invN;
        for (int n = 0; n < N; ++n) store_pack<VEC>(v_alphas + n * P + (e - PD), v);
        return;
    }
    const int d0 = (int)(e % D);
    const Pack<VEC> v = load_pack<VEC>(v_out_img + e);
    // route: -1 = mean channel (v / N to every sub-exposure), else the winner code of the max / min channel
    int route[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
        const int d = (d0 + k) % D;
        route[k] = -1;
        if (d == max_ch) route[k] = arg_max[(e + k) / D];
        if (d == min_ch) route[k] = arg_min[(e + k) / D];
        if (route[k] == kMeanWins) route[k] = -1;
    }
    for (int n = 0; n < N; ++n) {
        Pack<VEC> o;
#pragma unroll
        for (int k = 0; k < VEC; ++k) o.v[k] = route[k] < 0 ? v.v[k] * invN : (n == route[k] ? v.v[k] : 0.f);
        store_pack<VEC>(v_imgs + n * PD + e, o);
    }
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
                              int ref_quirk, float *out_img, float *out_alpha, uint8_t *arg_max, uint8_t *arg_min,
                              d4_stream_t stream) {
    D4_CHECK_ARG(imgs && alphas && out_img && out_alpha && N >= 1 && N < 255 && P >= 0 && D >= 1,
                 "d4_combine_fwd: bad arguments");
    if (P == 0) return 0;
    const bool vec4 = (P % 4 == 0) && (((uintptr_t)imgs | (uintptr_t)alphas | (uintptr_t)out_img | (uintptr_t)out_alpha) & 15) == 0;
    if (vec4)
        combine_fwd_kernel<4><<<cdiv((P * D + P) / 4, 256), 256, 0, as_stream(stream)>>>(
            imgs, alphas, N, P, D, max_ch, min_ch, ref_quirk, out_img, out_alpha, arg_max, arg_min);
    else
        combine_fwd_kernel<1><<<cdiv(P * D + P, 256), 256, 0, as_stream(stream)>>>(
            imgs, alphas, N, P, D, max_ch, min_ch, ref_quirk, out_img, out_alpha, arg_max, arg_min);
    D4_CHECK_LAUNCH("d4_combine_fwd");
    return 0;
}

extern "C" int d4_combine_bwd(const uint8_t *arg_max, const uint8_t *arg_min, int N, int64_t P, int D, int max_ch,
                              int min_ch, const float *v_out_img, const float *v_out_alpha, float *v_imgs,
                              float *v_alphas, d4_stream_t stream) {
    D4_CHECK_ARG(v_out_img && v_out_alpha && v_imgs && v_alphas && N >= 1 && P >= 0 && D >= 1,
                 "d4_combine_bwd: bad arguments");
    D4_CHECK_ARG((max_ch < 0 || max_ch >= D || arg_max) && (min_ch < 0 || min_ch >= D || arg_min),
                 "d4_combine_bwd: the max / min channel needs the winner map written by d4_combine_fwd");
    if (P == 0) return 0;
    const bool vec4 = (P % 4 == 0) && (((uintptr_t)v_out_img | (uintptr_t)v_out_alpha | (uintptr_t)v_imgs | (uintptr_t)v_alphas) & 15) == 0;
    if (vec4)
        combine_bwd_kernel<4><<<cdiv((P * D + P) / 4, 256), 256, 0, as_stream(stream)>>>(
            arg_max, arg_min, N, P, D, max_ch, min_ch, v_out_img, v_out_alpha, v_imgs, v_alphas);
    else
        combine_bwd_kernel<1><<<cdiv(P * D + P, 256), 256, 0, as_stream(stream)>>>(
            arg_max, arg_min, N, P, D, max_ch, min_ch, v_out_img, v_out_alpha, v_imgs, v_alphas);
    D4_CHECK_LAUNCH("d4_combine_bwd");
    return 0;
}
