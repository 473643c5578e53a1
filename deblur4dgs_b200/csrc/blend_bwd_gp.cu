// blend_bwd_gp.cu -- row a11 of SURVEY.md section 8: backward of the per-tile alpha compositing
// (gsplat rasterize_to_pixels bwd + ED normalisation backward), "grouped" formulation.
//
// The per-pixel recurrence (T, S) is inherently serial over the Gaussians of a tile, but the
// D+6 per-Gaussian sums over pixels are not.  The shuffle kernel in blend.cu pays ~100 issue slots
// per (warp, Gaussian) for a 32-lane butterfly over those D+6 values; here the two are separated:
//
//   phase 1 (lane = pixel):   walk the warp's hit list back to front, U hits per trip.  Everything
//       that does not depend on the running (T, S) -- exponent, validity vote, <c_g, v_out> -- is
//       straight-line independent code for the U hits (ILP); the recurrence is then one FMUL (T) and
//       one FFMA (S) deep per hit.  Per contributing Gaussian only the two pair scalars
//       fac = alpha*T and v_sigma = dL/dsigma are produced and parked in a per-warp
//       [GR Gaussians x 32 pixels] shared-memory tile (2 STS) together with a copy of the Gaussian's
//       geometry record, so a parked row is self-contained and survives the staging of later batches;
//   phase 2 (lane = Gaussian): once the tile cannot take another trip, lane (g, part) sweeps
//       32 / PARTS pixels for ITS Gaussian and accumulates all D+6 sums privately in registers --
//         v_colors[g]  = sum_p fac[g][p] * v_out[p][:]          (packed FFMA2, v_out broadcast from smem)
//         conic / xy / opacity sums = second moments of v_sigma[g][p] about the Gaussian centre
//       -- then one exchange joins the parts and each lane issues its share of the D+6 global
//       reductions (RED.ADD.F32): one per (warp, Gaussian, value).  No per-Gaussian warp reduction,
//       no CTA-level accumulator, no flush.
//
// Staging of the tile's Gaussians (geometry pre-scaled to base 2, colours, per-warp reach masks) is
// shared with blend.cu; with NBUF = 2 it is double-buffered: batch b+1 is staged by a rotating set of
// warps while the others already work on batch b, and there is ONE barrier per batch.
//
// The forward can hand over per-intersection HIT MASKS (BlendArgs::hit_masks: which of the tile's 8 pixel blocks
// passed the alpha test); they are staged in place of the geometric reach masks, so a warp only visits Gaussians
// that really contribute to one of its pixels.
//
// MEASURED at c3 (D = 17, 9 x 1.49 M intersections, B200; profiles/r01d_*): shuffle kernel 8.32 ms
// (7.7e9 warp instructions, issue slots 80 % busy); this kernel 6.19 ms with hit masks (4.3e9 warp instructions,
// issue slots 60 % busy, ~20 % of the samples at the per-batch barrier), 6.81 ms with reach masks; the launch
// shapes that were tried are listed at launch_blend_bwd_gp below and in profiles/r01d_bwd_experiments.md.
#include <limits.h>

#include <type_traits>

#include "blend_common.cuh"

namespace d4 {

__device__ __forceinline__ float rcp_approx(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// (fac, v_sigma) park tile row stride: 32 pixels + 1, conflict-free for the phase-1 row writes and the phase-2
// reads.  (An XOR swizzle without the pad column frees 1 KB but costs a LOP3 per sweep step: measured slower.)
constexpr int kGrpStride = 33;

// B = Gaussians per staged batch, NBUF = staging buffers, GR = rows of the per-warp park tile (8 or 16)
template <int D, int B, int NBUF, int GR>
struct GpCfg {
    static constexpr int DS = BlendCfg<D>::DS;
    static constexpr int V = BlendCfg<D>::V;
    static constexpr int NW = kBlendThreads / 32;
    static constexpr size_t smem_bytes() {
        return NBUF * (sizeof(float4) * 2 * B + sizeof(float) * B * DS + sizeof(uint32_t) * B)  // staged batches
               + sizeof(float) * kBlendThreads * DS                                             // v_out of the tile
               + sizeof(float) * NW * 2 * GR * kGrpStride                                       // (fac, v_sigma) tiles
               + sizeof(float4) * NW * GR * 2                                                   // parked geometry
               + NW * B;                                                                        // per-warp hit lists
    }
};

template <int D, int B, int NBUF, int MINB, int U, int GR, bool HL>
__global__ void __launch_bounds__(kBlendThreads, MINB)
blend_bwd_gp_kernel(BlendArgs a, const float *__restrict__ render_alphas, const int32_t *__restrict__ last_ids,
                    const float *__restrict__ acc_depth, const float *__restrict__ v_render_colors,
                    const float *__restrict__ v_render_alphas, float *__restrict__ v_means2d,
                    float *__restrict__ v_conics, float *__restrict__ v_colors, float *__restrict__ v_opacities,
                    float *__restrict__ v_depths) {
    using Cfg = GpCfg<D, B, NBUF, GR>;
    constexpr int DS = Cfg::DS, V = Cfg::V, NW = Cfg::NW;
    static_assert(B % 32 == 0 && B <= kBlendThreads, "batch must be a multiple of the warp size and <= 256");
    static_assert(GR == 8 || GR == 16, "park tile holds 8 or 16 Gaussians");
    static_assert(NBUF == 1 || NBUF == 2, "single- or double-buffered staging");
    static_assert(U == 4 && U <= GR / 2, "a trip is four hits (one packed word of the hit list)");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4 *s_geom_all = reinterpret_cast<float4 *>(smem_raw);              // [NBUF][B]
    float4 *s_conic_all = s_geom_all + NBUF * B;                            // [NBUF][B]
    float4 *s_rows_all = s_conic_all + NBUF * B;                            // [NW][GR][2]  parked (geom, conic)
    float *s_col_all = reinterpret_cast<float *>(s_rows_all + NW * GR * 2);  // [NBUF][B][DS]
    float *s_vout = s_col_all + NBUF * B * DS;                              // [256 pixels][DS], pixel == thread
    float *s_tiles = s_vout + kBlendThreads * DS;                           // [NW][2][GR][kGrpStride]
    uint32_t *s_mask_all = reinterpret_cast<uint32_t *>(s_tiles + NW * 2 * GR * kGrpStride);  // [NBUF][B]
    uint8_t *s_hits_all = reinterpret_cast<uint8_t *>(s_mask_all + NBUF * B);                 // [NW][B]
    __shared__ int32_t s_max[NW];

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
    __syncthreads();
    int32_t block_bin_final = s_max[0];
#pragma unroll
    for (int k = 1; k < NW; ++k) block_bin_final = max(block_bin_final, s_max[k]);
    range_end = min(range_end, (int64_t)block_bin_final + 1);
    if (range_end <= range_start) return;
    const int num_batches = (int)((range_end - range_start + B - 1) / B);

    float *s_fac = s_tiles + w * 2 * GR * kGrpStride;
    float *s_vs = s_fac + GR * kGrpStride;
    float4 *s_rows = s_rows_all + w * GR * 2;
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
        const bool rowok = pg < nb;  // rows >= nb hold stale data: computed on, never stored
        const float4 g0 = s_rows[2 * pg], cn = s_rows[2 * pg + 1];
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
        // one global reduction per (warp, Gaussian, value)
        const int d0 = a.depths ? D - 1 : D;
        const int32_t g = __float_as_int(g0.w);
        const int32_t gl = g - c * a.G;
#pragma unroll
        for (int k = 0; k < NHP; ++k) {
            const int idx = part * NHP + k;
            float *dst;
            if (idx < d0) dst = v_colors + c * a.colors_cs + (int64_t)gl * a.D0 + idx;
            else if (idx < D) dst = v_depths + g;
            else if (idx < D + 3) dst = v_conics + 3LL * g + (idx - D);
            else if (idx < D + 5) dst = v_means2d + 2LL * g + (idx - D - 3);
            else dst = v_opacities + gl;
            if (rowok && r[k] != 0.f && idx < V) atomicAdd(dst, r[k]);
        }
        __syncwarp();
        nb = 0;
    };

    // staging of batch nbt into buffer buf; with double buffering the B/32 staging warps rotate with the batch
    // index so that no warp carries the extra work every time
    auto stage_batch = [&](int nbt, int buf) {
        constexpr int SW = B / 32;
        const int rel = NBUF == 2 ? ((w - nbt * SW) & (NW - 1)) : w;
        if (rel < SW) {
            const int tr = rel * 32 + lane;
            const int64_t bend = range_end - 1 - (int64_t)B * nbt;  // slot 0 = furthest back
            const bool in_range = bend - tr >= range_start;
            StagedRec<D> r0;
            stage_load<D>(r0, a, c, bend - tr, in_range);
            const int hm = (a.hit_masks && in_range) ? (int)a.hit_masks[bend - tr] : -1;
            stage_store<D>(r0, tr, tx * kTile, ty * kTile, s_geom_all + buf * B, s_conic_all + buf * B,
                           s_col_all + buf * B * DS, s_mask_all + buf * B, hm);
        }
    };

    stage_batch(0, 0);
    __syncthreads();

    for (int b = 0; b < num_batches; ++b) {
        const int buf = NBUF == 2 ? (b & 1) : 0;
        if (NBUF == 2 && b + 1 < num_batches) stage_batch(b + 1, buf ^ 1);  // overlaps this batch's work
        const float4 *s_geom = s_geom_all + buf * B;
        const float4 *s_conic = s_conic_all + buf * B;
        const float *s_col = s_col_all + buf * B * DS;
        const uint32_t *s_mask = s_mask_all + buf * B;
        const int64_t batch_end = range_end - 1 - (int64_t)B * b;
        const int batch_size = (int)min((int64_t)B, batch_end + 1 - range_start);

        // slot t holds intersection batch_end - t; this pixel takes part from slot t_px on, the warp from t0 on
        const int t_px = (int)min((int64_t)INT_MAX, batch_end - (int64_t)bin_final);
        const int t0 = (int)max((int64_t)0, batch_end - (int64_t)warp_bin_final);

        // ---- hit list of this warp in the batch: slots whose reach mask has bit w set, from the warp's last
        // contributor (t0) on.  HL: compacted up front to one byte per hit, a trip then fetches its four slots
        // with one LDS (fewer instructions; pays off when registers are not the limit, D <= 9).  Otherwise a
        // warp-uniform bit iterator walks the masks on the fly.
        uint8_t *s_hits = s_hits_all + w * B;
        int n_hits = 0;
        int chunk = (t0 >> 5) - 1;
        uint32_t bits = 0u;
        bool more = true;
        if constexpr (HL) {
            for (int ch = t0 >> 5; ch * 32 < batch_size; ++ch) {
                const int t = ch * 32 + lane;
                const bool mine = ((s_mask[t] >> w) & 1u) != 0u && t >= t0;
                const uint32_t bb = __ballot_sync(0xffffffffu, mine);
                if (mine) s_hits[n_hits + __popc(bb & ((1u << lane) - 1u))] = (uint8_t)t;
                n_hits += __popc(bb);
            }
            __syncwarp();
        }
        auto next_hit = [&]() -> int {
            while (bits == 0u) {
                ++chunk;
                if (chunk * 32 >= batch_size) return -1;
                bits = __ballot_sync(0xffffffffu, (s_mask[chunk * 32 + lane] >> w) & 1u);
                if (chunk == (t0 >> 5)) bits &= ~((1u << (t0 & 31)) - 1u);
            }
            const int t = chunk * 32 + __ffs(bits) - 1;
            bits &= bits - 1;
            return t;
        };

        // ---- phase 1, U = 4 hits per trip
        for (int h = 0;; h += U) {
            int tu[U];
            if constexpr (HL) {
                if (h >= n_hits) break;
                const uint32_t packed = *reinterpret_cast<const uint32_t *>(s_hits + h);  // bytes past n_hits: masked
#pragma unroll
                for (int u = 0; u < U; ++u) tu[u] = h + u < n_hits ? (int)((packed >> (8 * u)) & 0xffu) : -1;
            } else {
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    tu[u] = more ? next_hit() : -1;
                    more = tu[u] >= 0;
                }
                if (tu[0] < 0) break;
            }
            if (nb > GR - U) sweep_group(std::integral_constant<int, GR>{});  // make room for U rows
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
                const float ra = rcp_approx(1.0f - al[u]);  // 1 - alpha in [0.001, 1]: MUFU.RCP is within 1 ulp here
                T *= ra;
                const float fac = al[u] * T;
                const float v_alpha = sd[u] * T - (S - tail) * ra;
                S = fmaf(sd[u], fac, S);
                const float vs = ar[u] != 0.f ? -ar[u] * v_alpha : 0.f;
                if ((cm >> u) & 1u) {  // warp-uniform: park the row with a copy of its geometry record
                    s_fac[nb * kGrpStride + lane] = fac;
                    s_vs[nb * kGrpStride + lane] = vs;
                    if (lane < 8) {
                        const float *src = reinterpret_cast<const float *>(lane < 4 ? s_geom + tu[u] : s_conic + tu[u]);
                        reinterpret_cast<float *>(s_rows + 2 * nb)[lane] = src[lane & 3];
                    }
                    ++nb;
                }
            }
        }
        if constexpr (NBUF == 1) {
            // single buffer: this thread's slot of the next batch is fetched into registers BEFORE the barrier --
            // the loads fly while the warp waits for the slowest warp of the tile -- and published after it
            StagedRec<D> rec;
            const bool stager = b + 1 < num_batches && tid < B;
            const int64_t bend = range_end - 1 - (int64_t)B * (b + 1);
            const bool in_next = stager && bend - tid >= range_start;
            stage_load<D>(rec, a, c, bend - tid, in_next);
            const int hm = (a.hit_masks && in_next) ? (int)a.hit_masks[bend - tid] : -1;
            __syncthreads();  // every warp is done with this batch's buffer
            if (b + 1 < num_batches) {
                if (stager) stage_store<D>(rec, tid, tx * kTile, ty * kTile, s_geom_all, s_conic_all, s_col_all, s_mask_all, hm);
                __syncthreads();
            }
        } else {
            __syncthreads();  // every warp is done with this batch's buffer and the next batch is staged
        }
    }
    // drain what is still parked
    if (nb > 0) {
        if (GR == 16 && nb <= 8) sweep_group(std::integral_constant<int, 8>{});  // 4 lanes per Gaussian
        else sweep_group(std::integral_constant<int, GR>{});
    }
}

template <int D, int B, int NBUF, int MINB, int U, int GR, bool HL = (D <= 9)>
static int launch_gp(const BlendArgs &a, const float *ra, const int32_t *li, const float *ad, const float *vrc,
                     const float *vra, float *vm, float *vc, float *vcol, float *vo, float *vd, cudaStream_t st) {
    constexpr size_t smem = GpCfg<D, B, NBUF, GR>::smem_bytes();
    // the opt-in is per device: set it on every launch (a per-process flag would miss a second GPU)
    if (cudaFuncSetAttribute(blend_bwd_gp_kernel<D, B, NBUF, MINB, U, GR, HL>,
                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        return 1;
    const int grid = a.C * a.tile_w * a.tile_h;
    blend_bwd_gp_kernel<D, B, NBUF, MINB, U, GR, HL><<<grid, kBlendThreads, smem, st>>>(a, ra, li, ad, vrc, vra, vm, vc, vcol, vo, vd);
    return 0;
}

// Largest batch (multiple of 32, <= 256) whose shared memory still lets NCTA CTAs share an SM
// (228 KB per SM, 1 KB reserved per CTA, small static arrays).
template <int D, int NBUF, int GR, int NCTA>
constexpr int batch_for() {
    constexpr size_t budget = (228 * 1024) / NCTA - 1024 - 64;
    int best = 32;
    for (int b = 32; b <= 256; b += 32) {
        const size_t per_slot = NBUF * (sizeof(float4) * 2 + sizeof(float) * BlendCfg<D>::DS + sizeof(uint32_t)) + kBlendThreads / 32;
        const size_t fixed = sizeof(float) * kBlendThreads * BlendCfg<D>::DS +
                             sizeof(float) * (kBlendThreads / 32) * 2 * GR * kGrpStride +
                             sizeof(float4) * (kBlendThreads / 32) * GR * 2;
        if (fixed + per_slot * b <= budget) best = b;
    }
    return best;
}

// Launch shape: single-buffered staging, the largest batch that keeps 3 CTAs on an SM, 16-row park tile.
// MEASURED at c3 (B200, N = 9 x 1.49 M intersections; blend_bwd ms, round 1): D = 17: this shape 7.04, double-buffered
// staging with one barrier per batch 7.53, 8-row park tile with 4 CTAs / SM 8.31, 8-row tile with 256-slot batches
// 7.70, compacted hit list 7.35; shuffle kernel (blend.cu) 8.32.  D = 5: 4.22 (compacted hit list; 4.54 without).
// Larger batches matter more than one barrier less, 16-row sweeps more than occupancy (the other shapes were
// removed in round 2; profiles/r01d_bwd_experiments.md keeps the numbers).
int launch_blend_bwd_gp(int D, const BlendArgs &a, const float *ra, const int32_t *li, const float *ad,
                        const float *vrc, const float *vra, float *vm, float *vc, float *vcol, float *vo, float *vd,
                        cudaStream_t st) {
#define GP_ARGS a, ra, li, ad, vrc, vra, vm, vc, vcol, vo, vd, st
    switch (D) {
#define X(n)                                                                                          \
    case n:                                                                                           \
        if constexpr (n <= 17) {                                                                      \
            return launch_gp<n, batch_for<n, 1, 16, 3>(), 1, 3, 4, 16>(GP_ARGS);                      \
        } else {                                                                                      \
            return launch_gp<n, 128, 1, 1, 4, 16>(GP_ARGS);                                           \
        }
        D4_FOR_EACH_D(X)
#undef X
#undef GP_ARGS
        default:
            return -1;
    }
}

}  // namespace d4
