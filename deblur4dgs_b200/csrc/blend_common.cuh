// blend_common.cuh -- helpers shared by the blend translation units (blend.cu, blend_bwd_gp.cu):
// tile / batch constants, the argument block, the per-thread pixel mapping and the staging of one
// batch of Gaussians into shared memory (geometry pre-scaled to base 2, per-warp reach masks).
#pragma once
#include "common.cuh"

namespace d4 {

constexpr int kTile = 16;
constexpr int kBlendThreads = kTile * kTile;
constexpr int kBatch = kBlendThreads;

template <int D>
struct BlendCfg {
    static constexpr int DS = (D + 3) / 4 * 4;  // smem colour stride (float4 aligned)
    static constexpr int DP = D | 1;            // odd stride for conflict-free per-pixel staging
    static constexpr int V = D + 6;             // per-Gaussian gradient values
};

struct BlendArgs {
    const float *means2d, *conics, *opacities, *colors, *depths, *backgrounds;
    int64_t colors_cs;
    int C, G, D0, width, height, tile_w, tile_h;
    const int32_t *tile_offsets, *flatten_ids;
    int64_t n_isects;
    int normalize_depth;
    // optional [n_isects] bytes, written by the forward and read by the backward: bit w of entry i is set iff some
    // pixel of warp w's 8x4 block passed the alpha test for intersection i in the forward.  The backward uses it in
    // place of the (conservative, geometric) reach mask: same arithmetic in both directions, so no contributing
    // pixel is lost, and the ~24 % of visits whose Gaussian touches a block's bounding box but no pixel disappear.
    uint8_t *hit_masks;
};

// pixel owned by this thread: warp w -> 8x4 block (w&1, w>>1), lane -> (lane&7, lane>>3)
__device__ __forceinline__ void pixel_of_thread(int tid, int &lx, int &ly) {
    int w = tid >> 5, lane = tid & 31;
    lx = (w & 1) * 8 + (lane & 7);
    ly = (w >> 1) * 4 + (lane >> 3);
}

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
constexpr float kLog2_255 = 7.994353436858858f;

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// Reach mask of one Gaussian for the tile whose first pixel is (tile_x0, tile_y0): bit w set iff the Gaussian can
// reach alpha >= 1/255 on some pixel of warp w's 8x4 block.  The ellipse {sigma <= ln(255 o)} has the bounding box
// |dx| <= sqrt(2 tau cov_xx), |dy| <= sqrt(2 tau cov_yy); the test is conservative (slack for rounding), so dropping
// a Gaussian with an empty mask never changes a result.  Same arithmetic as stage_store below.
__device__ __forceinline__ uint32_t reach_mask_of(float x, float y, float L, float ca, float cb, float cc, int tile_x0,
                                                  int tile_y0) {
    uint32_t mask = 0u;
    const float tau = (L + kLog2_255) * kLn2;  // ln(255 * opacity)
    const float det = ca * cc - cb * cb;
    if (!(det > 0.f) || !(ca > 0.f) || !(cc > 0.f)) {
        mask = 0xffu;  // degenerate conic: no culling
    } else if (tau >= 0.f) {
        const float k = 2.0f * tau / det;
        const float ex = sqrtf(k * cc) * 1.0001f + 1e-3f;
        const float ey = sqrtf(k * ca) * 1.0001f + 1e-3f;
        const float rx = x - (float)tile_x0, ry = y - (float)tile_y0;
#pragma unroll
        for (int w = 0; w < 8; ++w) {
            const float x0 = (float)((w & 1) * 8) + 0.5f, y0 = (float)((w >> 1) * 4) + 0.5f;
            const bool hit = (rx >= x0 - ex) && (rx <= x0 + 7.0f + ex) && (ry >= y0 - ey) && (ry <= y0 + 3.0f + ey);
            mask |= hit ? (1u << w) : 0u;
        }
    }
    return mask;
}

// Stage one batch: thread tr loads Gaussian `idx` (if in range) into slot tr.
// Shared-memory record per Gaussian (exponent pre-scaled to base 2 so that the per-pair
// evaluation is 5 FMA-pipe ops + one MUFU.EX2):
//   s_geom  = (x, y, L = log2(opacity), flatten id)
//   s_conic = (A', B', C', 1/opacity)  with  A' = -a/2 log2e, B' = -b log2e, C' = -c/2 log2e
//   => opacity * exp(-sigma) = exp2(L + A' dx^2 + B' dx dy + C' dy^2)
//   s_mask  = bit w set iff the Gaussian can reach alpha >= 1/255 on some pixel of warp w's 8x4
//             block: the ellipse {sigma <= ln(255 o)} has the bounding box
//             |dx| <= sqrt(2 tau cov_xx), |dy| <= sqrt(2 tau cov_yy); the test is conservative
//             (slack for rounding), so skipping never changes a result.
// Two-step staging (grouped backward): stage_load issues the global loads of a slot into registers, stage_store
// publishes the record later -- the loads of the next batch fly while the warp waits at the batch barrier.
template <int D, bool kStageColors = true>
struct StagedRec {
    int32_t g;  // flatten id, -1 when the slot is out of range
    float x, y, op, ca, cb, cc;
    float col[kStageColors ? BlendCfg<D>::DS : 1];
};

template <int D, bool kStageColors = true>
__device__ __forceinline__ void stage_load(StagedRec<D, kStageColors> &r, const BlendArgs &a, int c, int64_t idx,
                                           bool in_range) {
    constexpr int DS = BlendCfg<D>::DS;
    r.g = -1;
    r.x = r.y = r.op = r.ca = r.cb = r.cc = 0.f;
    if constexpr (kStageColors) {
#pragma unroll
        for (int k = 0; k < DS; ++k) r.col[k] = 0.f;  // pad lanes feed the packed fp32x2 path: keep them finite
    }
    if (in_range) {
        const int32_t g = __ldg(a.flatten_ids + idx);
        const int32_t gl = g - c * a.G;
        r.g = g;
        const float2 xy = __ldg(reinterpret_cast<const float2 *>(a.means2d) + g);
        r.x = xy.x, r.y = xy.y;
        r.op = __ldg(a.opacities + gl);
        const float *cp = a.conics + 3LL * g;
        r.ca = __ldg(cp), r.cb = __ldg(cp + 1), r.cc = __ldg(cp + 2);
        if constexpr (kStageColors) {
            const float *col = a.colors + c * a.colors_cs + (int64_t)gl * a.D0;
            const int d0 = a.depths ? D - 1 : D;
            if ((d0 & 3) == 0) {
#pragma unroll
                for (int k = 0; k < D / 4; ++k) {
                    if (4 * k < d0) {
                        const float4 v = __ldg(reinterpret_cast<const float4 *>(col) + k);
                        r.col[4 * k] = v.x, r.col[4 * k + 1] = v.y, r.col[4 * k + 2] = v.z, r.col[4 * k + 3] = v.w;
                    }
                }
            } else {
#pragma unroll
                for (int k = 0; k < D; ++k)
                    if (k < d0) r.col[k] = __ldg(col + k);
            }
            if (a.depths) r.col[D - 1] = __ldg(a.depths + g);
        }
    }
}

template <int D, bool kStageColors = true>
__device__ __forceinline__ void stage_store(const StagedRec<D, kStageColors> &r, int tr, int tile_x0, int tile_y0,
                                            float4 *s_geom, float4 *s_conic, float *s_col, uint32_t *s_mask,
                                            int given_mask = -1) {
    constexpr int DS = BlendCfg<D>::DS;
    if (r.g < 0) {
        s_mask[tr] = 0u;
        return;
    }
    const float L = __log2f(r.op);
    s_geom[tr] = make_float4(r.x, r.y, L, __int_as_float(r.g));
    s_conic[tr] = make_float4(-0.5f * kLog2e * r.ca, -kLog2e * r.cb, -0.5f * kLog2e * r.cc, 1.0f / r.op);
    // per-warp reach mask (or the forward's hit mask when the caller has one)
    uint32_t mask = 0u;
    const float tau = (L + kLog2_255) * kLn2;  // ln(255 * opacity)
    const float det = r.ca * r.cc - r.cb * r.cb;
    if (given_mask >= 0) {
        mask = (uint32_t)given_mask;
    } else if (!(det > 0.f) || !(r.ca > 0.f) || !(r.cc > 0.f)) {
        mask = 0xffu;  // degenerate conic: no culling
    } else if (tau >= 0.f) {
        const float k = 2.0f * tau / det;
        const float ex = sqrtf(k * r.cc) * 1.0001f + 1e-3f;
        const float ey = sqrtf(k * r.ca) * 1.0001f + 1e-3f;
        const float rx = r.x - (float)tile_x0, ry = r.y - (float)tile_y0;
#pragma unroll
        for (int w = 0; w < 8; ++w) {
            const float x0 = (float)((w & 1) * 8) + 0.5f, y0 = (float)((w >> 1) * 4) + 0.5f;
            const bool hit = (rx >= x0 - ex) && (rx <= x0 + 7.0f + ex) && (ry >= y0 - ey) && (ry <= y0 + 3.0f + ey);
            mask |= hit ? (1u << w) : 0u;
        }
    }
    s_mask[tr] = mask;
    if constexpr (kStageColors) {
        float *dst = s_col + tr * DS;
#pragma unroll
        for (int k4 = 0; k4 < DS / 4; ++k4)
            *reinterpret_cast<float4 *>(dst + 4 * k4) = make_float4(r.col[4 * k4], r.col[4 * k4 + 1], r.col[4 * k4 + 2], r.col[4 * k4 + 3]);
    }
}

// one-step staging (forward and shuffle backward): global loads go straight to shared memory
template <int D, bool kStageColors = true>
__device__ __forceinline__ void stage_gaussian(const BlendArgs &a, int c, int64_t idx, bool in_range, int tr,
                                               int tile_x0, int tile_y0, float4 *s_geom, float4 *s_conic,
                                               float *s_col, uint32_t *s_mask) {
    constexpr int DS = BlendCfg<D>::DS;
    if (!in_range) {
        s_mask[tr] = 0u;
        return;
    }
    int32_t g = __ldg(a.flatten_ids + idx);
    int32_t gl = g - c * a.G;
    float2 xy = __ldg(reinterpret_cast<const float2 *>(a.means2d) + g);
    float opac = __ldg(a.opacities + gl);
    const float *cp = a.conics + 3LL * g;
    const float ca = __ldg(cp), cb = __ldg(cp + 1), cc = __ldg(cp + 2);
    const float L = __log2f(opac);
    s_geom[tr] = make_float4(xy.x, xy.y, L, __int_as_float(g));
    s_conic[tr] = make_float4(-0.5f * kLog2e * ca, -kLog2e * cb, -0.5f * kLog2e * cc, 1.0f / opac);
    // per-warp reach mask
    uint32_t mask = 0u;
    const float tau = (L + kLog2_255) * kLn2;  // ln(255 * opacity)
    const float det = ca * cc - cb * cb;
    if (!(det > 0.f) || !(ca > 0.f) || !(cc > 0.f)) {
        mask = 0xffu;  // degenerate conic: no culling
    } else if (tau >= 0.f) {
        const float k = 2.0f * tau / det;
        const float ex = sqrtf(k * cc) * 1.0001f + 1e-3f;
        const float ey = sqrtf(k * ca) * 1.0001f + 1e-3f;
        const float rx = xy.x - (float)tile_x0, ry = xy.y - (float)tile_y0;
#pragma unroll
        for (int w = 0; w < 8; ++w) {
            const float x0 = (float)((w & 1) * 8) + 0.5f, y0 = (float)((w >> 1) * 4) + 0.5f;
            const bool hit = (rx >= x0 - ex) && (rx <= x0 + 7.0f + ex) && (ry >= y0 - ey) && (ry <= y0 + 3.0f + ey);
            mask |= hit ? (1u << w) : 0u;
        }
    }
    s_mask[tr] = mask;
    if constexpr (!kStageColors) return;
    const float *col = a.colors + c * a.colors_cs + (int64_t)gl * a.D0;
    float *dst = s_col + tr * DS;
    const int d0 = a.depths ? D - 1 : D;
    if ((d0 & 3) == 0) {
#pragma unroll
        for (int k = 0; k < D / 4; ++k) {
            if (4 * k < d0) {
                float4 v = __ldg(reinterpret_cast<const float4 *>(col) + k);
                *reinterpret_cast<float4 *>(dst + 4 * k) = v;
            }
        }
    } else {
#pragma unroll
        for (int k = 0; k < D; ++k)
            if (k < d0) dst[k] = __ldg(col + k);
    }
    if (a.depths) dst[D - 1] = __ldg(a.depths + g);
#pragma unroll
    for (int k = D; k < DS; ++k) dst[k] = 0.f;  // pad lanes feed the packed fp32x2 path: keep them finite
}

#define D4_FOR_EACH_D(X) X(1) X(2) X(3) X(4) X(5) X(6) X(7) X(8) X(9) X(16) X(17) X(32) X(33)

// blend_bwd_gp.cu: grouped backward (lane = pixel recurrence, lane = Gaussian accumulation).
// Returns 0 when launched, -1 when D is not built, 1 on a CUDA configuration error.
int launch_blend_bwd_gp(int D, const BlendArgs &a, const float *render_alphas, const int32_t *last_ids,
                        const float *acc_depth, const float *v_render_colors, const float *v_render_alphas,
                        float *v_means2d, float *v_conics, float *v_colors, float *v_opacities, float *v_depths,
                        cudaStream_t stream);

}  // namespace d4
