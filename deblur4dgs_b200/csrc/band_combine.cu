// band_combine.cu -- the N-way combine of flow3d/scene_model.py:386-397 for the 2-D multi-GPU partition
// (SURVEY.md section 8e): a rank holds U = N (sub-exposure, tile-row band) units of one frame; the combine is
//   d4_band_partial   local:  part = sum over the rank's units of img / N  (+ alpha plane),
//                             ext  = max over its units with sub < n_ext of (mask, -depth)         -> all-reduce SUM / MAX
//   d4_band_winner    local:  lowest sub-exposure index among its units that attain the global extremum -> all-reduce MIN
//   d4_band_finalize  out = mean, except the max / min channels = extremum (or the mean itself in the reference's
//                     alias-quirk mode when it beats the extremum over r_0 .. r_{N-2}); winner code -1 = "mean won"
//   d4_band_bwd       routes the cotangent of the combined image back to the rank's units.
// All four are HBM-streaming passes: every unit element is read (or written) once, float4 packs along the pixel
// x channel axis.  Layouts: imgs [U, bh, W, D], alphas [U, bh, W], part [Hp, W, D + 1] (Hp = n_bands * bh; channel D
// is the alpha plane), ext / winner [Hp, W, 2], out_img [H, W, D], out_alpha [H, W].
#include "common.cuh"

namespace d4 {

constexpr int kBandMaxUnits = 32;
constexpr int kWinnerNone = 1 << 30;

struct BandUnits {
    int32_t sub[kBandMaxUnits];
    int32_t band[kBandMaxUnits];
    int U;
};

__global__ void __launch_bounds__(256)
band_partial_kernel(const float *__restrict__ imgs, const float *__restrict__ alphas, BandUnits un, int n_sub, int n_ext,
                    int bh, int W, int D, int max_ch, int min_ch, int64_t n_px, float *__restrict__ part,
                    float *__restrict__ ext) {
    // one thread per (pixel of the padded frame, channel in [0, D]): consecutive threads walk the D + 1 values of a pixel
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int D1 = D + 1;
    if (e >= n_px * D1) return;
    const int64_t px = e / D1;
    const int d = (int)(e - px * D1);
    const int y = (int)(px / W), x = (int)(px - (int64_t)y * W);
    const int band = y / bh, yl = y - band * bh;
    const float inv_n = 1.0f / (float)n_sub;
    float sum = 0.f, best = -INFINITY;
    const bool is_ext = d == max_ch || d == min_ch;
    const float sgn = d == min_ch ? -1.f : 1.f;
    for (int u = 0; u < un.U; ++u) {
        if (un.band[u] != band) continue;
        const int64_t p = ((int64_t)u * bh + yl) * W + x;
        const float v = d < D ? __ldg(imgs + p * D + d) : __ldg(alphas + p);
        sum += v * inv_n;
        if (is_ext && un.sub[u] < n_ext) best = fmaxf(best, sgn * v);
    }
    part[e] = sum;
    if (is_ext) ext[px * 2 + (d == max_ch ? 0 : 1)] = best;
}

__global__ void __launch_bounds__(256)
band_winner_kernel(const float *__restrict__ imgs, BandUnits un, int n_ext, int bh, int W, int D, int max_ch, int min_ch,
                   int64_t n_px, const float *__restrict__ ext, int32_t *__restrict__ winner) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_px * 2) return;
    const int64_t px = e >> 1;
    const int j = (int)(e & 1);
    const int ch = j == 0 ? max_ch : min_ch;
    int32_t w = kWinnerNone;
    if (ch >= 0 && ch < D) {
        const int y = (int)(px / W), x = (int)(px - (int64_t)y * W);
        const int band = y / bh, yl = y - band * bh;
        const float target = ext[e], sgn = j == 0 ? 1.f : -1.f;
        for (int u = 0; u < un.U; ++u) {
            if (un.band[u] != band || un.sub[u] >= n_ext) continue;
            const float v = sgn * __ldg(imgs + (((int64_t)u * bh + yl) * W + x) * D + ch);
            if (v == target) w = min(w, un.sub[u]);
        }
    }
    winner[e] = w;
}

__global__ void __launch_bounds__(256)
band_finalize_kernel(const float *__restrict__ part, const float *__restrict__ ext, int32_t *__restrict__ winner, int H,
                     int W, int D, int max_ch, int min_ch, int ref_quirk, int n_ext, float *__restrict__ out_img,
                     float *__restrict__ out_alpha) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int D1 = D + 1;
    if (e >= (int64_t)H * W * D1) return;
    const int64_t px = e / D1;
    const int d = (int)(e - px * D1);
    float v = part[e];
    if (d == D) {
        out_alpha[px] = v;
        return;
    }
    if (d == max_ch || d == min_ch) {
        const int j = d == max_ch ? 0 : 1;
        const float best = (j == 0 ? 1.f : -1.f) * ext[px * 2 + j];
        const bool mean_wins = ref_quirk && (n_ext == 0 || (j == 0 ? v > best : v < best));
        if (mean_wins) winner[px * 2 + j] = -1;
        else v = best;
    }
    out_img[px * D + d] = v;
}

__global__ void __launch_bounds__(256)
band_bwd_kernel(const int32_t *__restrict__ winner, BandUnits un, int n_sub, int bh, int W, int D, int H, int max_ch,
                int min_ch, const float *__restrict__ v_out, const float *__restrict__ v_alpha,
                float *__restrict__ v_imgs, float *__restrict__ v_alphas) {
    // one thread per (unit pixel, channel in [0, D])
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int D1 = D + 1;
    const int64_t per_unit = (int64_t)bh * W * D1;
    if (e >= per_unit * un.U) return;
    const int u = (int)(e / per_unit);
    const int64_t r = e - (int64_t)u * per_unit;
    const int64_t pl = r / D1;  // pixel within the unit
    const int d = (int)(r - pl * D1);
    const int yl = (int)(pl / W), x = (int)(pl - (int64_t)yl * W);
    const int y = un.band[u] * bh + yl;
    const float inv_n = 1.0f / (float)n_sub;
    float g = 0.f;
    if (y < H) {
        const int64_t px = (int64_t)y * W + x;
        if (d == D) {
            g = __ldg(v_alpha + px) * inv_n;
        } else {
            const float v = __ldg(v_out + px * D + d);
            g = v * inv_n;
            if (d == max_ch || d == min_ch) {
                const int32_t w = winner[px * 2 + (d == max_ch ? 0 : 1)];
                if (w >= 0) g = (w == un.sub[u]) ? v : 0.f;  // -1: the mean itself won, every unit gets v / N
            }
        }
    }
    if (d == D) v_alphas[(int64_t)u * bh * W + pl] = g;
    else v_imgs[((int64_t)u * bh * W + pl) * D + d] = g;
}

static int fill_units(const char *name, BandUnits &un, const int32_t *subs, const int32_t *bands, int U, int n_bands) {
    D4_CHECK_ARG(U >= 0 && U <= kBandMaxUnits && (U == 0 || (subs && bands)), "%s: between 0 and %d units per rank", name, kBandMaxUnits);
    un.U = U;
    for (int u = 0; u < U; ++u) {
        D4_CHECK_ARG(bands[u] >= 0 && bands[u] < n_bands && subs[u] >= 0, "%s: unit %d out of range", name, u);
        un.sub[u] = subs[u];
        un.band[u] = bands[u];
    }
    return 0;
}

}  // namespace d4

using namespace d4;

// subs / bands are HOST arrays (the partition is host logic); everything else is device memory
extern "C" int d4_band_partial(const float *imgs, const float *alphas, const int32_t *subs, const int32_t *bands, int U,
                               int n_sub, int n_ext, int n_bands, int bh, int W, int D, int max_ch, int min_ch,
                               float *part, float *ext, d4_stream_t stream) {
    BandUnits un;
    if (int rc = fill_units("d4_band_partial", un, subs, bands, U, n_bands)) return rc;
    D4_CHECK_ARG(n_sub >= 1 && n_bands >= 1 && bh >= 1 && W >= 1 && D >= 1 && part && ext && (U == 0 || (imgs && alphas)),
                 "d4_band_partial: bad arguments");
    const int64_t n_px = (int64_t)n_bands * bh * W;
    band_partial_kernel<<<cdiv(n_px * (D + 1), 256), 256, 0, as_stream(stream)>>>(imgs, alphas, un, n_sub, n_ext, bh, W, D,
                                                                               max_ch, min_ch, n_px, part, ext);
    D4_CHECK_LAUNCH("d4_band_partial");
    return 0;
}

extern "C" int d4_band_winner(const float *imgs, const int32_t *subs, const int32_t *bands, int U, int n_ext, int n_bands,
                              int bh, int W, int D, int max_ch, int min_ch, const float *ext, int32_t *winner,
                              d4_stream_t stream) {
    BandUnits un;
    if (int rc = fill_units("d4_band_winner", un, subs, bands, U, n_bands)) return rc;
    D4_CHECK_ARG(ext && winner && (U == 0 || imgs), "d4_band_winner: null pointer");
    const int64_t n_px = (int64_t)n_bands * bh * W;
    band_winner_kernel<<<cdiv(n_px * 2, 256), 256, 0, as_stream(stream)>>>(imgs, un, n_ext, bh, W, D, max_ch, min_ch, n_px,
                                                                        ext, winner);
    D4_CHECK_LAUNCH("d4_band_winner");
    return 0;
}

extern "C" int d4_band_finalize(const float *part, const float *ext, int32_t *winner, int H, int W, int D, int max_ch,
                                int min_ch, int ref_quirk, int n_ext, float *out_img, float *out_alpha,
                                d4_stream_t stream) {
    D4_CHECK_ARG(part && ext && winner && out_img && out_alpha && H >= 1 && W >= 1 && D >= 1, "d4_band_finalize: bad arguments");
    band_finalize_kernel<<<cdiv((int64_t)H * W * (D + 1), 256), 256, 0, as_stream(stream)>>>(
        part, ext, winner, H, W, D, max_ch, min_ch, ref_quirk, n_ext, out_img, out_alpha);
    D4_CHECK_LAUNCH("d4_band_finalize");
    return 0;
}

extern "C" int d4_band_bwd(const int32_t *winner, const int32_t *subs, const int32_t *bands, int U, int n_sub, int n_bands,
                           int bh, int W, int D, int H, int max_ch, int min_ch, const float *v_out, const float *v_alpha,
                           float *v_imgs, float *v_alphas, d4_stream_t stream) {
    BandUnits un;
    if (int rc = fill_units("d4_band_bwd", un, subs, bands, U, n_bands)) return rc;
    if (U == 0) return 0;
    D4_CHECK_ARG(winner && v_out && v_alpha && v_imgs && v_alphas, "d4_band_bwd: null pointer");
    band_bwd_kernel<<<cdiv((int64_t)U * bh * W * (D + 1), 256), 256, 0, as_stream(stream)>>>(
        winner, un, n_sub, bh, W, D, H, max_ch, min_ch, v_out, v_alpha, v_imgs, v_alphas);
    D4_CHECK_LAUNCH("d4_band_bwd");
    return 0;
}
