// assemble.cu -- row f1 of SURVEY.md section 8 ("next"): the per-render-call glue in front of the
// rasterizer.  The reference recomputes it with ~10 elementwise / cat launches per render() call:
//   GaussianParams activations  scale = exp, opacity = sigmoid, colour = sigmoid   (flow3d/params.py:39-43,70-84)
//   fg | bg concatenation                                                            (flow3d/scene_model.py:122-143)
//   feature vector  colors_override = [rgb | fg-mask | track channels]              (flow3d/scene_model.py:205-289)
// Here: one pass, 1 thread per Gaussian, raw parameters read once, activated arrays written once.
#include "common.cuh"

namespace d4 {

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

__global__ void __launch_bounds__(256)
assemble_fwd_kernel(const float *__restrict__ fg_scales, const float *__restrict__ bg_scales,
                    const float *__restrict__ fg_opac, const float *__restrict__ bg_opac,
                    const float *__restrict__ fg_colors, const float *__restrict__ bg_colors,
                    const float *__restrict__ extra, int Gf, int Gb, int E, int with_mask,
                    float *__restrict__ scales, float *__restrict__ opac, float *__restrict__ colors) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    const int G = Gf + Gb;
    if (g >= G) return;
    const bool fg = g < Gf;
    const int l = fg ? g : g - Gf;
    const float *sp = (fg ? fg_scales : bg_scales) + 3LL * l;
    const float *cp = (fg ? fg_colors : bg_colors) + 3LL * l;
    const float o = fg ? fg_opac[l] : bg_opac[l];
    const int D0 = 3 + (with_mask ? 1 : 0) + E;
    float *so = scales + 3LL * g, *co = colors + (int64_t)g * D0;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        so[k] = expf(__ldg(sp + k));
        co[k] = sigmoidf_(__ldg(cp + k));
    }
    opac[g] = sigmoidf_(o);
    int k0 = 3;
    if (with_mask) co[k0++] = fg ? 1.0f : 0.0f;
    for (int k = 0; k < E; ++k) co[k0 + k] = __ldg(extra + (int64_t)g * E + k);
}

__global__ void __launch_bounds__(256)
assemble_bwd_kernel(const float *__restrict__ scales, const float *__restrict__ opac, const float *__restrict__ colors,
                    const float *__restrict__ v_scales, const float *__restrict__ v_opac,
                    const float *__restrict__ v_colors, int Gf, int Gb, int E, int with_mask,
                    float *__restrict__ v_fg_scales, float *__restrict__ v_bg_scales, float *__restrict__ v_fg_opac,
                    float *__restrict__ v_bg_opac, float *__restrict__ v_fg_colors, float *__restrict__ v_bg_colors,
                    float *__restrict__ v_extra) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    const int G = Gf + Gb;
    if (g >= G) return;
    const bool fg = g < Gf;
    const int l = fg ? g : g - Gf;
    const int D0 = 3 + (with_mask ? 1 : 0) + E;
    float *vs = (fg ? v_fg_scales : v_bg_scales) + 3LL * l;
    float *vcl = (fg ? v_fg_colors : v_bg_colors) + 3LL * l;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        vs[k] = __ldg(v_scales + 3LL * g + k) * __ldg(scales + 3LL * g + k);  // d exp = exp
        const float c = __ldg(colors + (int64_t)g * D0 + k);
        vcl[k] = __ldg(v_colors + (int64_t)g * D0 + k) * c * (1.0f - c);       // d sigmoid = s (1 - s)
    }
    const float o = __ldg(opac + g);
    (fg ? v_fg_opac : v_bg_opac)[l] = __ldg(v_opac + g) * o * (1.0f - o);
    if (v_extra) {
        const int k0 = 3 + (with_mask ? 1 : 0);
        for (int k = 0; k < E; ++k) v_extra[(int64_t)g * E + k] = __ldg(v_colors + (int64_t)g * D0 + k0 + k);
    }
}

}  // namespace d4

using namespace d4;

extern "C" int d4_assemble_fwd(const float *fg_scales, const float *bg_scales, const float *fg_opac,
                               const float *bg_opac, const float *fg_colors, const float *bg_colors,
                               const float *extra, int Gf, int Gb, int E, int with_mask, float *scales, float *opac,
                               float *colors, d4_stream_t stream) {
    D4_CHECK_ARG(Gf >= 0 && Gb >= 0 && E >= 0, "d4_assemble_fwd: bad sizes");
    if (Gf + Gb == 0) return 0;
    D4_CHECK_ARG(scales && opac && colors && (Gf == 0 || (fg_scales && fg_opac && fg_colors)) &&
                     (Gb == 0 || (bg_scales && bg_opac && bg_colors)) && (E == 0 || extra),
                 "d4_assemble_fwd: null pointer");
    assemble_fwd_kernel<<<cdiv(Gf + Gb, 256), 256, 0, as_stream(stream)>>>(fg_scales, bg_scales, fg_opac, bg_opac,
                                                                           fg_colors, bg_colors, extra, Gf, Gb, E,
                                                                           with_mask, scales, opac, colors);
    D4_CHECK_LAUNCH("d4_assemble_fwd");
    return 0;
}

extern "C" int d4_assemble_bwd(const float *scales, const float *opac, const float *colors, const float *v_scales,
                               const float *v_opac, const float *v_colors, int Gf, int Gb, int E, int with_mask,
                               float *v_fg_scales, float *v_bg_scales, float *v_fg_opac, float *v_bg_opac,
                               float *v_fg_colors, float *v_bg_colors, float *v_extra, d4_stream_t stream) {
    D4_CHECK_ARG(Gf >= 0 && Gb >= 0 && E >= 0, "d4_assemble_bwd: bad sizes");
    if (Gf + Gb == 0) return 0;
    D4_CHECK_ARG(scales && opac && colors && v_scales && v_opac && v_colors &&
                     (Gf == 0 || (v_fg_scales && v_fg_opac && v_fg_colors)) &&
                     (Gb == 0 || (v_bg_scales && v_bg_opac && v_bg_colors)),
                 "d4_assemble_bwd: null pointer");
    assemble_bwd_kernel<<<cdiv(Gf + Gb, 256), 256, 0, as_stream(stream)>>>(
        scales, opac, colors, v_scales, v_opac, v_colors, Gf, Gb, E, with_mask, v_fg_scales, v_bg_scales, v_fg_opac,
        v_bg_opac, v_fg_colors, v_bg_colors, v_extra);
    D4_CHECK_LAUNCH("d4_assemble_bwd");
    return 0;
}
