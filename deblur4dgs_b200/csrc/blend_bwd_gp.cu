// blend_bwd_gp.cu -- row a11 of SURVEY.md section 8: backward of the per-tile alpha compositing
// (gsplat rasterize_to_pixels bwd + ED normalisation backward), "grouped" formulation.
//
// The per-pixel recurrence (T, S) is inherently serial over the Gaussians of a tile, but the
// D+6 per-Gaussian sums over pixels are not.  The shuffle kernel in blend.cu pays ~100 issue slots
// per (warp, Gaussian) for a 32-lane butterfly over those D+6 values; here the two are separated:
//
//   phase 1 (lane = pixel):   walk the warp's hit list back to front; per contributing Gaussian
//       only the two pair scalars  fac = alpha*T  and  v_sigma = dL/dsigma  are produced and
//       parked in a per-warp [16 Gaussians x 32 pixels] shared-memory tile (2 STS);
//   phase 2 (lane = Gaussian): once 16 Gaussians are parked, lane (g, half) sweeps 16 of the 32
//       pixels for ITS Gaussian and accumulates all D+6 sums privately in registers --
//         v_colors[g]  = sum_p fac[g][p] * v_out[p][:]          (packed FFMA2, v_out broadcast from smem)
//         conic / xy / opacity sums = second moments of v_sigma[g][p] about the Gaussian centre
//       -- then one 16-lane exchange joins the two halves and the totals go to the per-CTA
//       accumulator.  No per-Gaussian warp reduction at all: ~35 issue slots per Gaussian.
//
// Everything else (one CTA per (camera, tile), 8x4 pixel block per warp, reach masks, base-2
// exponent, CTA-level accumulator flushed once per (tile, Gaussian), overlap of the flush with the
// staging of the next batch) is shared with blend.cu.
#include <limits.h>

#include <type_traits>

#include "blend_common.cuh"

namespace d4 {

constexpr int kGrpStride = 33;  // 32 pixels + 1: conflict-free for the phase-1 row writes and the phase-2 reads

// GR = Gaussians parked per warp before one phase-2 sweep (8 or 16); 32 / GR lanes share one Gaussian
template <int D, int B, int GR>
struct GpCfg {
    static constexpr int DS = BlendCfg<D>::DS;
    static constexpr int V = BlendCfg<D>::V;
    static constexpr int VS = V | 1;  // odd accumulator stride
    static constexpr size_t smem_bytes() {
        return sizeof(float4) * 2 * B + sizeof(float) * B * (DS + VS) + sizeof(float) * kBlendThreads * DS +
               sizeof(float) * (kBlendThreads / 32) * 2 * GR * kGrpStride;
    }
};

template <int D, int B, int MINB, int U, int GR>
__global__ void __launch_bounds__(kBlendThreads, MINB)
blend_bwd_gp_kernel(BlendArgs a, const float *__restrict__ render_alphas, const int32_t *__restrict__ last_ids,
                    const float *__restrict__ acc_depth, const float *__restrict__ v_render_colors,
                    const float *__restrict__ v_render_alphas, float *__restrict__ v_means2d,
                    float *__restrict__ v_conics, float *__restrict__ v_colors, float *__restrict__ v_opacities,
                    float *__restrict__ v_depths) {
    using Cfg = GpCfg<D, B, GR>;
    constexpr int DS = Cfg::DS, V = Cfg::V, VS = Cfg::VS;
    static_assert(GR == 8 || GR == 16, "phase-2 group is 8 or 16 Gaussians");
    constexpr int NW = kBlendThreads / 32;
    static_assert(B % 32 == 0 && B <= kBlendThreads, "batch must be a multiple of the warp size and <= 256");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4 *s_geom = reinterpret_cast<float4 *>(smem_raw);
    float4 *s_conic = s_geom + B;
    float *s_col = reinterpret_cast<float *>(s_conic + B);
    float *s_vout = s_col + B * DS;                 // [256 pixels][DS], pixel index == thread index
    float *s_acc = s_vout + kBlendThreads * DS;     // [B][VS]
    float *s_tiles = s_acc + B * VS;                // [8 warps][2][GR][kGrpStride]
    __shared__ int32_t s_max[NW];
    __shared__ uint32_t s_mask[B];
    __shared__ int32_t s_slot[NW][GR];
    __shared__ int32_t s_gid[2][B];  // flatten ids of the batch being processed / being flushed

    const int n_tiles = a.tile_w * a.tile_h;
    const int ct = blockIdx.x;
    const int c = ct / n_tiles;
    const int tile = ct - c * n_tiles;
    const int ty = tile / a.tile_w, tx = tile - ty * a.tile_w;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    int lx, ly;
    pixel_of_thread(tid, lx, ly);
    const int j = tx * kTile + lx, i = ty * kTile + ly;
    const bool inside = (i < a.height) && (j < a.width);
    const float px = (float)j + 0.5f, py = (float)i + 0.5f;
    const int64_t pid = ((int64_t)c * a.height + i) * a.width + j;

    const int64_t range_start = a.tile_offsets[ct];
    int64_t range_end = (ct == a.C * n_tiles - 1) ? a.n_isects : (int64_t)a.tile_offsets[ct + 1];
    if (range_end <= range_start) return;  // uniform for the CTA

    // ---- per-pixel state
    constexpr int D2 = (D + 1) / 2;
    float2 v2[D2];  // v_out as fp32x2 pairs (pad lane zero)
    float T_final = 1.f, v_ra = 0.f, bgdot = 0.f;
    int32_t bin_final = -1;
    {
        float v_out[D];
        if (inside) {
            const float alpha_px = render_alphas[pid];
            T_final = 1.0f - alpha_px;
            bin_final = last_ids[pid];
            v_ra = v_render_alphas[pid];
#pragma unroll
            for (int k = 0; k < D; ++k) v_out[k] = __ldg(v_render_colors + pid * D + k);
            if (a.normalize_depth) {
                const float ac = fmaxf(alpha_px, 1e-10f);
                const float vd = v_out[D - 1];
                v_out[D - 1] = vd / ac;
                if (alpha_px > 1e-10f) v_ra += -vd * acc_depth[pid] / (ac * ac);
            }
        } else {
#pragma unroll
            for (int k = 0; k < D; ++k) v_out[k] = 0.f;
        }
        if (a.backgrounds) {
            const int d0 = a.depths ? D - 1 : D;
#pragma unroll
            for (int k = 0; k < D; ++k)
                if (k < d0) bgdot = fmaf(__ldg(a.backgrounds + (int64_t)c * a.D0 + k), v_out[k], bgdot);
        }
#pragma unroll
        for (int k2 = 0; k2 < D2; ++k2) v2[k2] = make_float2(v_out[2 * k2], (2 * k2 + 1 < D) ? v_out[2 * k2 + 1] : 0.f);
        // phase 2 reads every pixel's v_out as a warp-wide broadcast
        float *vo = s_vout + tid * DS;
#pragma unroll
        for (int k = 0; k < DS; ++k) vo[k] = k < D ? v_out[k] : 0.f;
    }
    // constant part of dL/dalpha_i * (1 - alpha_i):  T_final * (v_alpha_out - bg.v_out)
    const float tail = T_final * (v_ra - bgdot);
    float T = T_final;
    float S = 0.f;  // sum_{j>i} <c_j, v_out> alpha_j T_j

    // nothing behind the last contributing Gaussian of any pixel of the CTA matters
    const int32_t warp_bin_final = __reduce_max_sync(0xffffffffu, bin_final);
    if (lane == 0) s_max[w] = warp_bin_final;
    for (int e = tid; e < B * VS; e += kBlendThreads) s_acc[e] = 0.f;
    __syncthreads();
    int32_t block_bin_final = s_max[0];
#pragma unroll
    for (int k = 1; k < NW; ++k) block_bin_final = max(block_bin_final, s_max[k]);
    range_end = min(range_end, (int64_t)block_bin_final + 1);
    if (range_end <= range_start) return;
    const int num_batches = (int)((range_end - range_start + B - 1) / B);

    float *s_fac = s_tiles + w * 2 * GR * kGrpStride;
    float *s_vs = s_fac + GR * kGrpStride;
    const float bx0 = (float)(tx * kTile + (w & 1) * 8) + 0.5f;
    int nb = 0;  // Gaussians parked in the warp's tile (warp-uniform)

    // ---- phase 2: GS Gaussians per sweep; lane (pg, part) sweeps 32 / PARTS = GS pixels for Gaussian row pg,
    // pixels [part*GS, part*GS + GS) of the warp's 8x4 block (pixel p sits at x = p & 7, y = p >> 3)
    auto sweep_group = [&](auto gs_tag) {
        constexpr int GS = decltype(gs_tag)::value;
        constexpr int PARTS = 32 / GS;
        constexpr int NHP = (V + PARTS - 1) / PARTS;  // values owned by each part after the exchange
        __syncwarp();
        const int pg = lane & (GS - 1), part = lane / GS;
        const float by0 = (float)(ty * kTile + (w >> 1) * 4 + (part * GS) / 8) + 0.5f;
        const float *p2_vo = s_vout + (w * 32 + part * GS) * DS;
        const bool rowok = pg < nb;
        const int t = rowok ? s_slot[w][pg] : 0;
        const float4 g0 = s_geom[t], cn = s_conic[t];
        const float Xl = g0.x - bx0, Yl = g0.y - by0;
        float2 acc[DS / 2];
#pragma unroll
        for (int k = 0; k < DS / 2; ++k) acc[k] = make_float2(0.f, 0.f);
        float axx = 0.f, axy = 0.f, ayy = 0.f, ax = 0.f, ay = 0.f, a0 = 0.f;
        const float *fr = s_fac + pg * kGrpStride + part * GS;
        const float *vr = s_vs + pg * kGrpStride + part * GS;
#pragma unroll
        for (int q = 0; q < GS; ++q) {
            const float fac = fr[q], vs = vr[q];
            const float2 f2 = make_float2(fac, fac);
#pragma unroll
            for (int k4 = 0; k4 < DS / 4; ++k4) {
                const float4 v = *reinterpret_cast<const float4 *>(p2_vo + q * DS + 4 * k4);
                acc[2 * k4] = __ffma2_rn(f2, make_float2(v.x, v.y), acc[2 * k4]);
                acc[2 * k4 + 1] = __ffma2_rn(f2, make_float2(v.z, v.w), acc[2 * k4 + 1]);
            }
            const float dx = Xl - (float)(q & 7), dy = Yl - (float)(q >> 3);
            const float t1 = vs * dx, t2 = vs * dy;
            axx = fmaf(t1, dx, axx);
            axy = fmaf(t1, dy, axy);
            ayy = fmaf(t2, dy, ayy);
            ax += t1;
            ay += t2;
            a0 += vs;
        }
        // the D+6 values of this part (all linear in the moments, so the parts simply add)
        //   conic (a, b, c) = (-2A', -B', -2C') / log2e ;  cn.w = 1 / opacity
        float r[PARTS * NHP];
#pragma unroll
        for (int k = 0; k < D; ++k) r[k] = (k & 1) ? acc[k >> 1].y : acc[k >> 1].x;
        const float ka = cn.x * (-2.0f / kLog2e), kb = cn.y * (-1.0f / kLog2e), kc = cn.z * (-2.0f / kLog2e);
        r[D + 0] = 0.5f * axx;
        r[D + 1] = axy;
        r[D + 2] = 0.5f * ayy;
        r[D + 3] = fmaf(ka, ax, kb * ay);
        r[D + 4] = fmaf(kb, ax, kc * ay);
        r[D + 5] = -cn.w * a0;
#pragma unroll
        for (int k = V; k < PARTS * NHP; ++k) r[k] = 0.f;
        // transposing exchange between the parts: lane (pg, part) ends up owning values [part*NHP, part*NHP + NHP)
        {
            int n = PARTS * NHP;
#pragma unroll
            for (int off = 16; off >= GS; off >>= 1) {
                n >>= 1;
                const bool upper = (lane & off) != 0;
#pragma unroll
                for (int k = 0; k < n; ++k) {
                    const float send = upper ? r[k] : r[k + n];
                    const float keep = upper ? r[k + n] : r[k];
                    r[k] = keep + __shfl_xor_sync(0xffffffffu, send, off);
                }
            }
        }
        float *dst = s_acc + t * VS + part * NHP;
#pragma unroll
        for (int k = 0; k < NHP; ++k)
            if (rowok && r[k] != 0.f && (part * NHP + k < V)) atomicAdd(dst + k, r[k]);
        __syncwarp();
        nb = 0;
    };

    // flush of one batch's CTA-level sums: one global atomic per non-zero (Gaussian, value); zeroes as it goes.
    // Warp w takes slots w, w + 8, ...; lane k < V owns value k, so its destination array / stride are fixed.
    auto flush_acc = [&](int n_slots, const int32_t *gids) {
        static_assert(V <= 64, "flush: at most two values per lane");
        float *fl_base = nullptr;
        int fl_stride = 0;
        bool fl_local = false;  // indexed by the camera-local Gaussian id (colours, opacity) instead of the flat id
        {
            const int d0 = a.depths ? D - 1 : D;
            const int k = lane;
            if (k < d0) { fl_base = v_colors + c * a.colors_cs + k; fl_stride = a.D0; fl_local = true; }
            else if (k < D) { fl_base = v_depths; fl_stride = 1; }
            else if (k < D + 3) { fl_base = v_conics + (k - D); fl_stride = 3; }
            else if (k < D + 5) { fl_base = v_means2d + (k - D - 3); fl_stride = 2; }
            else if (k < V) { fl_base = v_opacities; fl_stride = 1; fl_local = true; }
        }
        for (int t = w; t < n_slots; t += NW) {
            const int32_t g = gids[t];
            if (lane < V) {
                const float val = s_acc[t * VS + lane];
                if (val != 0.f) {
                    s_acc[t * VS + lane] = 0.f;
                    atomicAdd(fl_base + (int64_t)(fl_local ? g - c * a.G : g) * fl_stride, val);
                }
            }
            if constexpr (V > 32) {  // D = 32 / 33: values 32.. are owned a second time by lanes 0..V-33
                const int k = 32 + lane;
                if (k < V) {
                    const float val = s_acc[t * VS + k];
                    if (val != 0.f) {
                        s_acc[t * VS + k] = 0.f;
                        const int d0 = a.depths ? D - 1 : D;
                        const int32_t gl = g - c * a.G;
                        float *dst;
                        if (k < d0) dst = v_colors + c * a.colors_cs + (int64_t)gl * a.D0 + k;
                        else if (k < D) dst = v_depths + g;
                        else if (k < D + 3) dst = v_conics + 3LL * g + (k - D);
                        else if (k < D + 5) dst = v_means2d + 2LL * g + (k - D - 3);
                        else dst = v_opacities + gl;
                        atomicAdd(dst, val);
                    }
                }
            }
        }
    };
    int prev_size = 0;

    for (int b = 0; b < num_batches; ++b) {
        // (barrier C of the previous iteration has passed: every warp is done with batch b-1)
        const int64_t batch_end = range_end - 1 - (int64_t)B * b;  // slot 0 = furthest back
        const int batch_size = (int)min((int64_t)B, batch_end + 1 - range_start);
        if (tid < B) {
            const bool in_range = batch_end - tid >= range_start;
            stage_gaussian<D>(a, c, batch_end - tid, in_range, tid, tx * kTile, ty * kTile, s_geom, s_conic, s_col,
                              s_mask);
            s_gid[b & 1][tid] = in_range ? __ldg(a.flatten_ids + (batch_end - tid)) : 0;
        }
        flush_acc(prev_size, s_gid[(b & 1) ^ 1]);  // overlaps the staging loads of this batch
        prev_size = batch_size;
        __syncthreads();  // barrier B: staging visible, accumulators clean

        // slot t holds intersection batch_end - t; this pixel takes part from slot t_px on, the warp from t0 on
        const int t_px = (int)min((int64_t)INT_MAX, batch_end - (int64_t)bin_final);
        const int t0 = (int)max((int64_t)0, batch_end - (int64_t)warp_bin_final);

        // warp-uniform iterator over this warp's hit list (slots whose reach mask has bit w set)
        int chunk = (t0 >> 5) - 1;
        uint32_t bits = 0u;
        auto next_hit = [&]() -> int {
            while (bits == 0u) {
                ++chunk;
                if (chunk * 32 >= batch_size) return -1;
                bits = __ballot_sync(0xffffffffu, (s_mask[chunk * 32 + lane] >> w) & 1u);
                if (chunk == (t0 >> 5)) bits &= ~((1u << (t0 & 31)) - 1u);  // slots behind the warp's last contributor
            }
            const int t = chunk * 32 + __ffs(bits) - 1;
            bits &= bits - 1;
            return t;
        };

        // ---- phase 1, U hits per trip.  Everything that does not depend on the running (T, S) -- exponent,
        // validity, <c_g, v_out> -- is evaluated for all U hits first as straight-line independent code (ILP);
        // the recurrence itself is then one FMUL (T) and one FFMA (S) deep per hit.
        bool more = true;
        for (;;) {
            int tu[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                tu[u] = more ? next_hit() : -1;
                more = tu[u] >= 0;
            }
            const bool have = tu[0] >= 0;
            if (nb > 0 && (!have || nb > GR - U)) {  // make room for U rows / drain at the batch end
                if (GR == 16 && nb <= 8) sweep_group(std::integral_constant<int, 8>{});  // short drain: 4 lanes per Gaussian
                else sweep_group(std::integral_constant<int, GR>{});
            }
            if (!have) break;
            float al[U], ar[U], sd[U];
            uint32_t cm = 0u;  // hits with at least one contributing pixel in this warp
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int t = max(tu[u], 0);
                const float4 g0 = s_geom[t], cn = s_conic[t];
                const float dx = g0.x - px, dy = g0.y - py;
                const float power = fmaf(cn.z * dy, dy, fmaf(fmaf(cn.y, dy, cn.x * dx), dx, g0.z));
                const float araw = ex2_approx(power);  // opacity * exp(-sigma)
                const float alpha = fminf(kAlphaMax, araw);
                const bool valid = tu[u] >= 0 && tu[u] >= t_px && power <= g0.z && alpha >= kAlphaMin;
                al[u] = valid ? alpha : 0.f;
                ar[u] = (valid && araw <= kAlphaMax) ? araw : 0.f;  // dL/dsigma is zero where alpha was clamped
                cm |= __any_sync(0xffffffffu, valid) ? (1u << u) : 0u;
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                // s = <c_g, v_out>, four independent partial sums
                const float *cp = s_col + max(tu[u], 0) * DS;
                float2 sa = make_float2(0.f, 0.f), sb = make_float2(0.f, 0.f);
#pragma unroll
                for (int k4 = 0; k4 < DS / 4; ++k4) {  // pad lanes of v2 are zero
                    const float4 cv = *reinterpret_cast<const float4 *>(cp + 4 * k4);
                    if (2 * k4 < D2) sa = __ffma2_rn(make_float2(cv.x, cv.y), v2[2 * k4], sa);
                    if (2 * k4 + 1 < D2) sb = __ffma2_rn(make_float2(cv.z, cv.w), v2[2 * k4 + 1], sb);
                }
                sd[u] = (sa.x + sa.y) + (sb.x + sb.y);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                // alpha == 0 (pixel not taking part): ra = 1, T and S unchanged, fac = v_sigma = 0
                const float ra = __fdividef(1.0f, 1.0f - al[u]);  // alpha <= 0.999: MUFU.RCP is within 1 ulp here
                T *= ra;
                const float fac = al[u] * T;
                const float v_alpha = sd[u] * T - (S - tail) * ra;
                S = fmaf(sd[u], fac, S);
                const float vs = ar[u] != 0.f ? -ar[u] * v_alpha : 0.f;
                if ((cm >> u) & 1u) {  // warp-uniform: park the row
                    s_fac[nb * kGrpStride + lane] = fac;
                    s_vs[nb * kGrpStride + lane] = vs;
                    if (lane == 0) s_slot[w][nb] = tu[u];
                    ++nb;
                }
            }
        }
        __syncthreads();  // barrier C: every warp is done with batch b
    }
    flush_acc(prev_size, s_gid[(num_batches & 1) ^ 1]);
}

// Largest batch (multiple of 32, <= 256) whose shared memory still lets 3 CTAs share an SM
// (228 KB per SM, 1 KB reserved per CTA, ~12 B / slot + 0.6 KB of static arrays).
template <int D, int GR>
constexpr int batch_for_3_ctas() {
    int best = 32;
    constexpr size_t per_slot = sizeof(float4) * 2 + sizeof(float) * (BlendCfg<D>::DS + (BlendCfg<D>::V | 1)) + 12;
    constexpr size_t fixed = sizeof(float) * kBlendThreads * BlendCfg<D>::DS +
                             sizeof(float) * (kBlendThreads / 32) * 2 * GR * kGrpStride + 640;
    constexpr size_t budget = (228 * 1024) / 3 - 1024;
    for (int b = 32; b <= 256; b += 32)
        if (fixed + per_slot * b <= budget) best = b;
    return best;
}

template <int D, int B, int MINB, int U, int GR>
static int launch_gp(const BlendArgs &a, const float *ra, const int32_t *li, const float *ad, const float *vrc,
                     const float *vra, float *vm, float *vc, float *vcol, float *vo, float *vd, cudaStream_t st) {
    constexpr size_t smem = GpCfg<D, B, GR>::smem_bytes();
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(blend_bwd_gp_kernel<D, B, MINB, U, GR>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)smem) != cudaSuccess)
            return 1;
        configured = true;
    }
    const int grid = a.C * a.tile_w * a.tile_h;
    blend_bwd_gp_kernel<D, B, MINB, U, GR><<<grid, kBlendThreads, smem, st>>>(a, ra, li, ad, vrc, vra, vm, vc, vcol, vo, vd);
    return 0;
}

// launch configurations <batch, CTAs / SM, hits per phase-1 trip, phase-2 group> (D <= 17; wider D runs one fixed shape).
// Default: the largest batch that keeps 3 CTAs / SM (96 at D = 17, 256 at D <= 8).
// MEASURED at c3 (D = 17, B200, profiles/r01d_bwd_gp.md): <96,3,4,16> 7.73 ms (default), <128,3,4,8> 8.31, <64,4,4,8> 8.49,
// <256,2,4,16> 8.67, <96,3,8,16> 8.80, <64,3,4,8> 8.98, <256,2,8,16> 9.66; the shuffle kernel of blend.cu 8.32 ms.
int launch_blend_bwd_gp(int D, int cfg, const BlendArgs &a, const float *ra, const int32_t *li, const float *ad,
                        const float *vrc, const float *vra, float *vm, float *vc, float *vcol, float *vo, float *vd,
                        cudaStream_t st) {
#define GP_ARGS a, ra, li, ad, vrc, vra, vm, vc, vcol, vo, vd, st
    switch (D) {
#define X(n)                                                                           \
    case n:                                                                            \
        if constexpr (n <= 17) {                                                       \
            if (cfg == 1) return launch_gp<n, 128, 3, 4, 8>(GP_ARGS);                  \
            return launch_gp<n, batch_for_3_ctas<n, 16>(), 3, 4, 16>(GP_ARGS);         \
        } else {                                                                       \
            return launch_gp<n, 128, 1, 4, 16>(GP_ARGS);                               \
        }
        D4_FOR_EACH_D(X)
#undef X
#undef GP_ARGS
        default:
            return -1;
    }
}

}  // namespace d4
