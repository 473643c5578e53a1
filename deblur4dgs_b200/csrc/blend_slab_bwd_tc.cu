// blend_slab_bwd_tc.cu -- row a11 of SURVEY.md section 8, tensor-core formulation of blend_slab_bwd.cu for the
// 16-colour(+depth) records of the benchmark configuration (gsplat rasterize_to_pixels bwd + ED normalisation
// backward, call site flow3d/scene_model.py:360-373).
//
// Ring, hit words and the per-warp 16-row queue are those of blend_slab_bwd.cu.  What changes is where the two
// contractions of a queued 16-row group run:
//   phase 1  S[16 rows x 32 pixels]  = C[16 x 16 colours] . V^T[16 x 32]       (s = <c_g, v_out> of every pair)
//   phase 2  G[16 rows x 16 colours] = F[16 x 32 pixels]  . V[32 x 16]         (v_colors, F = alpha * T)
//            g[16 rows x 8]          = (F | VS)[16 x 32]  . Q[32 x 8]          (v_depth and the six moments of
//                                                                               v_sigma: Q = vd, 1, x, y, x^2, xy, y^2)
// go through mma.sync.m16n8k8 (TF32 operands, fp32 accumulate) with the 3xTF32 split a = a_hi + a_lo, b = b_hi + b_lo
// (a_lo b_hi + a_hi b_lo + a_hi b_hi: the product is exact to ~2^-22, i.e. fp32-grade; the pixel-coordinate columns
// of Q are small integers and exact in TF32).  The SIMT kernel spends per group 256 instructions / 128 shared-memory
// wavefronts on S and ~500 / ~180 on G, g; here the operands are fetched once per fragment (conflict-free LDS.32 from
// the same tiles) and the (T, S) recurrence of phase 1 -- the only serial part -- stays lane = pixel on the fp32 pipe.
// The gradients leave as vectorised reductions (red.global.add.v2.f32) straight from the accumulator fragments.
#include <limits.h>

#include "slab.cuh"

namespace d4 {

namespace {

constexpr int kD0 = 16;    // colour channels of this specialisation
constexpr int kRows = 16;  // rows of the per-warp queue == M of the MMAs
constexpr int kPark = 36;  // row stride of the (fac | S) and v_sigma tiles: A-fragment loads hit bank 4 g + t

__device__ __forceinline__ float rcp_approx_t(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// D += A(16x8, row) . B(8x8, col), TF32 operands.  Fragments (g = lane >> 2, t = lane & 3):
//   a0 (g, t)  a1 (g + 8, t)  a2 (g, t + 4)  a3 (g + 8, t + 4);   b0 (k = t, n = g)  b1 (k = t + 4, n = g)
//   d0 (g, 2t) d1 (g, 2t + 1) d2 (g + 8, 2t) d3 (g + 8, 2t + 1)
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// x = hi + lo with hi carrying the 10 mantissa bits TF32 keeps (lo is exact in fp32; the MMA truncates it to TF32)
__device__ __forceinline__ void split_tf32(float x, uint32_t &hi, uint32_t &lo) {
    hi = __float_as_uint(x) & 0xffffe000u;
    lo = __float_as_uint(x - __uint_as_float(hi));
}
__device__ __forceinline__ void red_add_v2(float *p, float x, float y) {
    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(x), "f"(y) : "memory");
}

// cotangent tile of the CTA: pixel `pl` of warp w keeps its 16 colour cotangents at (w * 32 + pl) * 16, channel ch at
// position ch ^ vsw(pl).  With this key both B-fragment walks -- phase 1 (k = channel, n = pixel) and phase 2
// (k = pixel, n = channel) -- touch 32 distinct banks per load.
__device__ __forceinline__ int vsw(int pl) { return ((pl >> 1) & 1) * 8 + ((pl >> 2) & 1) * 4; }

template <bool DEPTH>
struct TcCfg {
    static constexpr int NW = kSlabConsumers;
    __host__ __device__ static constexpr size_t warp_bytes() {
        return sizeof(float4) * kRows * 2 + sizeof(float) * kRows * kD0 + sizeof(float) * 2 * kRows * kPark +
               sizeof(int32_t) * kRows;
    }
    static constexpr size_t fixed_bytes() {
        return sizeof(float) * kBlendThreads * kD0 + (DEPTH ? sizeof(float) * kBlendThreads : 0) + NW * warp_bytes() + 128;
    }
    static constexpr size_t stage_bytes() { return (size_t)kSlabChunk * (32 + 4 * kD0); }
#ifndef D4_TC_BWD_CTAS
#define D4_TC_BWD_CTAS 3
#endif
    static constexpr int kCtas = D4_TC_BWD_CTAS;
    static constexpr int stages() {
        const size_t budget = (228 * 1024) / kCtas - 1024 - 64;
        int s = (int)((budget - fixed_bytes()) / stage_bytes());
        return s > 8 ? 8 : s;
    }
    static constexpr int kStages = stages();
    static_assert(kStages >= 2, "the ring needs two stages");
    static constexpr size_t smem_bytes() { return fixed_bytes() + kStages * stage_bytes(); }
};

}  // namespace

// TC1: phase 1's <c_g, v_out> on the tensor cores as well (false: per-lane FFMA2 dot products as in the SIMT kernel)
template <bool DEPTH, bool TC1>
__global__ void __launch_bounds__(kSlabThreads, (TcCfg<DEPTH>::kCtas))
blend_bwd_slab_tc_kernel(SlabArgs a, const float *__restrict__ render_alphas, const int32_t *__restrict__ last_ids,
                         const float *__restrict__ acc_depth, const float *__restrict__ v_render_colors,
                         const float *__restrict__ v_render_alphas, float *__restrict__ v_means2d,
                         float *__restrict__ v_conics, float *__restrict__ v_colors, float *__restrict__ v_opacities,
                         float *__restrict__ v_depths) {
    using Cfg = TcCfg<DEPTH>;
    constexpr int D0 = kD0, D = D0 + (DEPTH ? 1 : 0), NW = Cfg::NW, S = Cfg::kStages, CH = kSlabChunk;
    constexpr int U = 4, GR = kRows, PK = kPark;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float4 *s_rec = reinterpret_cast<float4 *>(smem_raw);                                 // [S][CH][2]
    float *s_col = reinterpret_cast<float *>(s_rec + S * CH * 2);                         // [S][CH][D0]
    float *s_vout = s_col + S * CH * D0;                                                  // [256 pixels][D0], swizzled
    float *s_vd = s_vout + kBlendThreads * D0;                                            // [256] depth cotangent
    unsigned char *s_warp_all = reinterpret_cast<unsigned char *>(s_vd + (DEPTH ? kBlendThreads : 0));
    uint64_t *s_full = reinterpret_cast<uint64_t *>(s_warp_all + NW * Cfg::warp_bytes());  // [S]
    uint64_t *s_empty = s_full + S;                                                       // [S]
    __shared__ int32_t s_max[NW];

    const int n_tiles = a.tile_w * a.tile_h;
    const int ct = blockIdx.x;
    const int c = ct / n_tiles;
    const int tile = ct - c * n_tiles;
    const int ty = tile / a.tile_w, tx = tile - ty * a.tile_w;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const bool producer = w == NW;

    const int32_t seg_start = a.tile_offsets[ct];
    const int32_t cnt = a.rec_counts[ct];
    if (cnt <= 0) return;  // uniform for the CTA

    int lx = 0, ly = 0;
    if (!producer) pixel_of_thread(tid, lx, ly);
    const int j = tx * kTile + lx, i = ty * kTile + ly;
    const bool inside = !producer && (i < a.height) && (j < a.width);
    const float px = (float)j + 0.5f, py = (float)i + 0.5f;
    const int64_t pid = ((int64_t)c * a.height + i) * a.width + j;

    // ---- step 1 of the prologue: where the stream starts.  Only last_ids is needed for that, so it is fetched alone
    // and the CTA meets once; the producer then streams the first stages WHILE the pixels fetch their cotangents
    int32_t bin_final = -1;
    if (inside) bin_final = last_ids[pid];
    const int32_t warp_bin_final = __reduce_max_sync(0xffffffffu, bin_final);
    if (!producer && lane == 0) s_max[w] = warp_bin_final;
    if (tid == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(s_full + s, 1 + 32);
            mbar_init(s_empty + s, NW);
        }
        mbar_init_fence();
    }
    __syncthreads();
    // nothing behind the last contributing record of any pixel of the CTA matters
    int32_t block_bin_final = s_max[0];
#pragma unroll
    for (int k = 1; k < NW; ++k) block_bin_final = max(block_bin_final, s_max[k]);
    const int rel_hi = block_bin_final - seg_start;  // last needed record, relative to the tile's run
    if (rel_hi < 0) return;                          // uniform for the CTA
    const int c_hi = rel_hi / CH;                    // chunks c_hi .. 0 are streamed, in this order

    if (producer) {
        // ------------------------------------------------------------------------------------ producer warp
        int stage = 0, phase = 0;
        auto load_idm = [&](int k) -> uint32_t {
            const int n_valid = min(CH, rel_hi + 1 - k * CH);
            return lane < n_valid
                       ? __ldg(reinterpret_cast<const uint32_t *>(a.recs + 2 * ((int64_t)seg_start + k * CH + lane)) + 3)
                       : 0u;
        };
        uint32_t idm_next = load_idm(c_hi);
        for (int k = c_hi; k >= 0; --k) {
            if (c_hi - k >= S) mbar_wait(s_empty + stage, phase ^ 1);
            const uint32_t idm = idm_next;
            if (k > 0) idm_next = load_idm(k - 1);
            const int n_valid = min(CH, rel_hi + 1 - k * CH);
            slab_issue_stage<D0, true>(a, c, (int64_t)seg_start + (int64_t)k * CH, n_valid, idm, s_rec + stage * CH * 2,
                                       s_col + stage * CH * D0, s_full + stage, lane);
            if (++stage == S) stage = 0, phase ^= 1;
        }
        return;  // every consumer waits for every stage: no copy is in flight when the CTA retires
    }

    // ---- step 2 (consumer warps): per-pixel state.  A warp only ever reads the cotangent rows of its OWN 32 pixels
    // (s_vw, s_vd + w * 32), so a warp barrier orders these stores -- no second CTA barrier
    constexpr int D2 = D0 / 2;
    [[maybe_unused]] float2 v2[TC1 ? 1 : D2];  // colour cotangent as fp32x2 pairs (SIMT phase 1 only)
    float vd = 0.f;                            // depth cotangent (after the ED normalisation backward)
    float T_final = 1.f, v_ra = 0.f, bgdot = 0.f;
    {
        float v_out[D];
        if (inside) {
            const float alpha_px = render_alphas[pid];
            T_final = 1.0f - alpha_px;
            v_ra = v_render_alphas[pid];
#pragma unroll
            for (int k = 0; k < D; ++k) v_out[k] = __ldg(v_render_colors + pid * D + k);
            if constexpr (DEPTH) {
                if (a.normalize_depth) {
                    const float ac = fmaxf(alpha_px, 1e-10f);
                    const float vdd = v_out[D - 1];
                    v_out[D - 1] = vdd / ac;
                    if (alpha_px > 1e-10f) v_ra += -vdd * acc_depth[pid] / (ac * ac);
                }
            }
        } else {
#pragma unroll
            for (int k = 0; k < D; ++k) v_out[k] = 0.f;
        }
        if (a.backgrounds) {
#pragma unroll
            for (int k = 0; k < D0; ++k) bgdot = fmaf(__ldg(a.backgrounds + (int64_t)c * D0 + k), v_out[k], bgdot);
        }
        if constexpr (!TC1) {
#pragma unroll
            for (int k2 = 0; k2 < D2; ++k2) v2[k2] = make_float2(v_out[2 * k2], v_out[2 * k2 + 1]);
        }
        if constexpr (DEPTH) vd = v_out[D - 1];
        // the B operand of both contractions: 16-byte piece q of this pixel's row at piece position q ^ (vsw >> 2)
        float *vo = s_vout + tid * D0;
        const int pk = vsw(lane) >> 2;
#pragma unroll
        for (int k4 = 0; k4 < D0 / 4; ++k4)
            *reinterpret_cast<float4 *>(vo + 4 * (k4 ^ pk)) =
                make_float4(v_out[4 * k4], v_out[4 * k4 + 1], v_out[4 * k4 + 2], v_out[4 * k4 + 3]);
        if constexpr (DEPTH) s_vd[tid] = vd;
    }
    __syncwarp();

    // ---------------------------------------------------------------------------------------- consumer warps
    // constant part of dL/dalpha_i * (1 - alpha_i):  T_final * (v_alpha_out - bg.v_out)
    const float tail = T_final * (v_ra - bgdot);
    float T = T_final;
    float Sacc = 0.f;  // sum_{j>i} <c_j, v_out> alpha_j T_j

    unsigned char *s_warp = s_warp_all + w * Cfg::warp_bytes();
    float4 *s_qrec = reinterpret_cast<float4 *>(s_warp);                  // [GR][2]   queued records
    float *s_qcol = reinterpret_cast<float *>(s_qrec + GR * 2);           // [GR][D0]  queued colour rows (swizzled)
    float *s_fac = s_qcol + GR * D0;                                      // [GR][PK]  S, then fac = alpha * T
    float *s_vs = s_fac + GR * PK;                                        // [GR][PK]  v_sigma, then the moment table
    int32_t *s_qidx = reinterpret_cast<int32_t *>(s_vs + GR * PK);        // [GR]      record indices
    const float *s_vw = s_vout + w * 32 * D0;                             // the warp's 32 cotangent rows
    const int fg = lane >> 2, ft = lane & 3;                              // fragment coordinates of this lane
    int nb = 0;  // rows queued (warp-uniform)

    // ---- phase 1 over the nb queued rows (lane = pixel), U rows per trip
    auto phase1 = [&]() {
        __syncwarp();
        if constexpr (TC1) {
            // S = C . V^T: 2 k-steps (8 channels each) x 4 n-tiles (8 pixels each) x 3 (split) MMAs
            float sacc[4][4];
#pragma unroll
            for (int nt = 0; nt < 4; ++nt)
#pragma unroll
                for (int e = 0; e < 4; ++e) sacc[nt][e] = 0.f;
            const int key = (fg >> 1) & 3;  // slab_key<16>(fg) == slab_key<16>(fg + 8)
            const float *ca = s_qcol + fg * D0 + ft;
            const int vk = vsw(fg);  // pixel 8 nt + fg of the warp: the key does not depend on nt
            const float *vb = s_vw + fg * D0;
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
                uint32_t ah[4], al[4];
                const int p0 = 4 * ((2 * ks) ^ key), p1 = 4 * ((2 * ks + 1) ^ key);
                split_tf32(ca[p0], ah[0], al[0]);
                split_tf32(ca[8 * D0 + p0], ah[1], al[1]);
                split_tf32(ca[p1], ah[2], al[2]);
                split_tf32(ca[8 * D0 + p1], ah[3], al[3]);
                const int c0 = (8 * ks + ft) ^ vk, c1 = (8 * ks + 4 + ft) ^ vk;
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) {
                    uint32_t bh0, bl0, bh1, bl1;
                    split_tf32(vb[nt * 8 * D0 + c0], bh0, bl0);
                    split_tf32(vb[nt * 8 * D0 + c1], bh1, bl1);
                    mma_tf32(sacc[nt], al, bh0, bh1);
                    mma_tf32(sacc[nt], ah, bl0, bl1);
                    mma_tf32(sacc[nt], ah, bh0, bh1);
                }
            }
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                *reinterpret_cast<float2 *>(s_fac + fg * PK + 8 * nt + 2 * ft) = make_float2(sacc[nt][0], sacc[nt][1]);
                *reinterpret_cast<float2 *>(s_fac + (fg + 8) * PK + 8 * nt + 2 * ft) = make_float2(sacc[nt][2], sacc[nt][3]);
            }
            __syncwarp();
        }
#pragma unroll 1
        for (int r8 = 0; r8 < nb; r8 += 8) {  // eight rows per iteration: the colour swizzle keys are static
#pragma unroll
            for (int h = 0; h < 8; h += U) {
                const int r0 = r8 + h;
                float al[U], ar[U], sd[U];
                int4 qi = *reinterpret_cast<const int4 *>(s_qidx + r0);
                const int32_t qidx[U] = {qi.x, qi.y, qi.z, qi.w};
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int r = r0 + u;
                    const float4 g0 = s_qrec[2 * r], cn = s_qrec[2 * r + 1];
                    const float dx = g0.x - px, dy = g0.y - py;
                    const float power = fmaf(cn.z * dy, dy, fmaf(fmaf(cn.y, dy, cn.x * dx), dx, g0.z));
                    const float araw = ex2_approx(power);  // opacity * exp(-sigma)
                    const float alpha = fminf(kAlphaMax, araw);
                    const bool valid = qidx[u] <= bin_final && power <= g0.z && alpha >= kAlphaMin;
                    al[u] = valid ? alpha : 0.f;
                    ar[u] = (valid && araw <= kAlphaMax) ? araw : 0.f;  // dL/dsigma is zero where alpha was clamped
                    float s;
                    if constexpr (TC1) {
                        s = s_fac[r * PK + lane];
                    } else {
                        const float *cp = s_qcol + r * D0;
                        constexpr int PPSm = D0 / 4 - 1;
                        const int key = slab_key<D0>(h + u);  // == slab_key(r): r8 is a multiple of 8
                        float2 sa = make_float2(0.f, 0.f), sb = make_float2(0.f, 0.f);
#pragma unroll
                        for (int k4 = 0; k4 < D0 / 4; ++k4) {
                            const float4 cv = *reinterpret_cast<const float4 *>(cp + 4 * ((k4 ^ key) & PPSm));
                            sa = __ffma2_rn(make_float2(cv.x, cv.y), v2[TC1 ? 0 : 2 * k4], sa);
                            sb = __ffma2_rn(make_float2(cv.z, cv.w), v2[TC1 ? 0 : 2 * k4 + 1], sb);
                        }
                        s = (sa.x + sa.y) + (sb.x + sb.y);
                    }
                    if constexpr (DEPTH) s = fmaf(cn.w, vd, s);
                    sd[u] = s;
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    // alpha == 0 (pixel not taking part): ra = 1, T and S unchanged, fac = v_sigma = 0
                    const float ra = rcp_approx_t(1.0f - al[u]);  // 1 - alpha in [0.001, 1]: MUFU.RCP is within 1 ulp here
                    T *= ra;
                    const float fac = al[u] * T;
                    const float v_alpha = sd[u] * T - (Sacc - tail) * ra;
                    Sacc = fmaf(sd[u], fac, Sacc);
                    const float vs = ar[u] != 0.f ? -ar[u] * v_alpha : 0.f;
                    s_fac[(r0 + u) * PK + lane] = fac;
                    s_vs[(r0 + u) * PK + lane] = vs;
                }
            }
        }
    };

    // ---- phase 2: the per-row sums over the warp's 32 pixels (pixel p sits at x = p & 7, y = p >> 3 of the 8x4 block)
    // B of the third n-tile, column n = fg: 0 depth cotangent, 1..6 the monomials 1, x, y, x^2, xy, y^2, 7 zero --
    // value at pixel (x = ft + 4 jj, y = ks) is q0[jj] + ks q1[jj] + ks^2 q2
    float q0[2], q1[2];
#pragma unroll
    for (int jj = 0; jj < 2; ++jj) {
        const float x = (float)(ft + 4 * jj);
        q0[jj] = fg == 1 ? 1.f : (fg == 2 ? x : (fg == 4 ? x * x : 0.f));
        q1[jj] = fg == 3 ? 1.f : (fg == 5 ? x : 0.f);
    }
    const float q2 = fg == 6 ? 1.f : 0.f;
    const float bx0 = (float)(tx * kTile + (w & 1) * 8) + 0.5f;
    const float by0 = (float)(ty * kTile + (w >> 1) * 4) + 0.5f;

    auto phase2 = [&]() {
        __syncwarp();
        float acc[2][4], accd[4], accm[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[0][e] = acc[1][e] = accd[e] = accm[e] = 0.f;
        const float *fa = s_fac + fg * PK + ft, *va = s_vs + fg * PK + ft;
        // colour B fragments: b0 = V[pixel 8 ks + ft][channel 8 nt + fg], b1 = V[pixel 8 ks + ft + 4][same]
        const int h8 = 8 * (ft >> 1);
        const float *vb0 = s_vw + ft * D0 + fg, *vb1 = s_vw + (ft + 4) * D0 + (fg ^ 4);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            uint32_t ah[4], al[4];
            split_tf32(fa[8 * ks], ah[0], al[0]);
            split_tf32(fa[8 * ks + 8 * PK], ah[1], al[1]);
            split_tf32(fa[8 * ks + 4], ah[2], al[2]);
            split_tf32(fa[8 * ks + 8 * PK + 4], ah[3], al[3]);
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) {
                const int off = ks * 8 * D0 + ((8 * nt) ^ h8);
                uint32_t bh0, bl0, bh1, bl1;
                split_tf32(vb0[off], bh0, bl0);
                split_tf32(vb1[off], bh1, bl1);
                mma_tf32(acc[nt], al, bh0, bh1);
                mma_tf32(acc[nt], ah, bl0, bl1);
                mma_tf32(acc[nt], ah, bh0, bh1);
            }
            float b0 = fmaf((float)(ks * ks), q2, fmaf((float)ks, q1[0], q0[0]));
            float b1 = fmaf((float)(ks * ks), q2, fmaf((float)ks, q1[1], q0[1]));
            if constexpr (DEPTH) {
                if (fg == 0) {
                    b0 = s_vd[w * 32 + 8 * ks + ft];
                    b1 = s_vd[w * 32 + 8 * ks + ft + 4];
                }
            }
            uint32_t bh0, bl0, bh1, bl1;
            split_tf32(b0, bh0, bl0);
            split_tf32(b1, bh1, bl1);
            if constexpr (DEPTH) {
                mma_tf32(accd, al, bh0, bh1);
                mma_tf32(accd, ah, bl0, bl1);
                mma_tf32(accd, ah, bh0, bh1);
            }
            uint32_t vh[4], vl[4];
            split_tf32(va[8 * ks], vh[0], vl[0]);
            split_tf32(va[8 * ks + 8 * PK], vh[1], vl[1]);
            split_tf32(va[8 * ks + 4], vh[2], vl[2]);
            split_tf32(va[8 * ks + 8 * PK + 4], vh[3], vl[3]);
            mma_tf32(accm, vl, bh0, bh1);  // the monomial columns are exact in TF32: no b_lo term
            mma_tf32(accm, vh, bh0, bh1);
        }
        __syncwarp();  // every lane is done with the fac / v_sigma tiles
        // per-row table [v_depth, m0, mx, my, mxx, mxy, myy, -] in place of the v_sigma tile
        {
            const bool dcol = DEPTH && ft == 0;
            *reinterpret_cast<float2 *>(s_vs + fg * PK + 2 * ft) = make_float2(dcol ? accd[0] : accm[0], accm[1]);
            *reinterpret_cast<float2 *>(s_vs + (fg + 8) * PK + 2 * ft) = make_float2(dcol ? accd[2] : accm[2], accm[3]);
        }
        // v_colors straight from the accumulators: rows fg and fg + 8, channels 8 nt + 2 ft, + 1
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int row = fg + 8 * half;
            if (row < nb) {
                const uint32_t gl = reinterpret_cast<const uint32_t *>(s_qrec + 2 * row)[3] & kRecIdMask;
                float *vcol = v_colors + c * a.colors_cs + (int64_t)gl * D0 + 2 * ft;
#pragma unroll
                for (int nt = 0; nt < 2; ++nt) {
                    const float x = acc[nt][2 * half], y = acc[nt][2 * half + 1];
                    if (x != 0.f || y != 0.f) red_add_v2(vcol + 8 * nt, x, y);
                }
            }
        }
        __syncwarp();
        if (lane < nb) {  // lane = row: geometry gradients from the moments of v_sigma about the block origin
            const float4 g0 = s_qrec[2 * lane], cn = s_qrec[2 * lane + 1];
            const float4 ma = *reinterpret_cast<const float4 *>(s_vs + lane * PK);
            const float4 mb = *reinterpret_cast<const float4 *>(s_vs + lane * PK + 4);
            const float m0 = ma.y, mx = ma.z, my = ma.w, mxx = mb.x, mxy = mb.y, myy = mb.z;
            const float X = g0.x - bx0, Y = g0.y - by0;  // dx = X - x, dy = Y - y on pixel (x, y) of the block
            const float ax = fmaf(X, m0, -mx), ay = fmaf(Y, m0, -my);         // sum vs dx, sum vs dy
            const float axx = fmaf(X, ax, fmaf(-X, mx, mxx));                 // sum vs dx^2
            const float axy = fmaf(X, ay, fmaf(-Y, mx, mxy));                 // sum vs dx dy
            const float ayy = fmaf(Y, ay, fmaf(-Y, my, myy));                 // sum vs dy^2
            //   conic (a, b, c) = (-2A', -B', -2C') / log2e ;  1 / opacity = exp2(-L)
            const float ka = cn.x * (-2.0f / kLog2e), kb = cn.y * (-1.0f / kLog2e), kc = cn.z * (-2.0f / kLog2e);
            const int32_t gl = (int32_t)(__float_as_uint(g0.w) & kRecIdMask);
            const int64_t g = (int64_t)c * a.G + gl;
            if constexpr (DEPTH) {
                if (ma.x != 0.f) atomicAdd(v_depths + g, ma.x);
            }
            if (axx != 0.f) atomicAdd(v_conics + 3LL * g, 0.5f * axx);
            if (axy != 0.f) atomicAdd(v_conics + 3LL * g + 1, axy);
            if (ayy != 0.f) atomicAdd(v_conics + 3LL * g + 2, 0.5f * ayy);
            const float gx = fmaf(ka, ax, kb * ay), gy = fmaf(kb, ax, kc * ay);
            if (gx != 0.f || gy != 0.f) red_add_v2(v_means2d + 2LL * g, gx, gy);
            if (m0 != 0.f) atomicAdd(v_opacities + gl, -ex2_approx(-g0.z) * m0);
        }
        __syncwarp();
        nb = 0;
    };

    // ---- the ring, consumer side: chunk c_hi first
    const int kw_hi = warp_bin_final - seg_start;       // last record of the warp, relative (< 0: none)
    const int k_warp_hi = kw_hi < 0 ? -1 : kw_hi / CH;  // chunks above hold no hit words written for this warp
    const int64_t hb_base = ((int64_t)(seg_start >> 5) + ct) * NW + w;
    // Hit words of this warp for 32 chunks at a time: lane i holds the word of chunk k_top - i.  Bits behind the warp's
    // last contributing record are dropped (see blend_slab_bwd.cu).
    auto load_words = [&](int k_top) -> uint32_t {
        const int kk = k_top - lane;
        if (kk < 0 || kk > k_warp_hi) return 0u;
        uint32_t wd = __ldg(a.hit_bits + hb_base + (int64_t)kk * NW);
        if (kk == k_warp_hi) wd &= 0xffffffffu >> (31 - (kw_hi & 31));
        return wd;
    };
    int stage = 0, phase = 0;
    int k = c_hi + 1;      // chunk being drained (none yet)
    int k_top = c_hi;      // chunk whose hit word lane 0 holds
    uint32_t words = load_words(k_top);
    bool holding = false;  // the ring stage of chunk k is still in use
    uint32_t bits = 0u;    // hits of chunk k not yet queued
    for (;;) {
        // ---- fill the queue: lane L owns record L of the chunk; hits are queued back to front (highest record first)
        while (nb < GR) {  // warp-uniform
            if (bits == 0u) {
                if (holding) {  // chunk drained: give the stage back
                    __syncwarp();
                    if (lane == 0) mbar_arrive(s_empty + stage);
                    if (++stage == S) stage = 0, phase ^= 1;
                    holding = false;
                }
                if (k == 0) break;
                --k;
                if (k_top - k == 32) {
                    k_top = k;
                    words = load_words(k_top);
                }
                bits = __shfl_sync(0xffffffffu, words, k_top - k);
                mbar_wait(s_full + stage, phase);
                holding = true;
                continue;
            }
            const float4 *recs = s_rec + stage * CH * 2;
            const float *cols = s_col + stage * CH * D0;
            const bool hit = (bits >> lane) & 1u;
            const int row = nb + __popc(bits & ~((2u << lane) - 1u));
            const bool take = hit && row < GR;
            if (take) {
                // the two halves of the record in the order that keeps a quarter-warp on distinct banks
                const int h0 = (lane >> 2) & 1;
                const float4 ra0 = recs[2 * lane + h0], ra1 = recs[2 * lane + (h0 ^ 1)];
                s_qrec[2 * row + h0] = ra0;
                s_qrec[2 * row + (h0 ^ 1)] = ra1;
                constexpr int PPSm = D0 / 4 - 1;
                const int ks = slab_key<D0>(lane), kr = slab_key<D0>(row);
#pragma unroll
                for (int k4 = 0; k4 < D0 / 4; ++k4)  // logical piece k4: swizzled by slot in the stage, by row in the queue
                    *reinterpret_cast<float4 *>(s_qcol + row * D0 + 4 * ((k4 ^ kr) & PPSm)) =
                        *reinterpret_cast<const float4 *>(cols + lane * D0 + 4 * ((k4 ^ ks) & PPSm));
                s_qidx[row] = seg_start + k * CH + lane;
            }
            const uint32_t taken = __ballot_sync(0xffffffffu, take);
            bits &= ~taken;
            nb += __popc(taken);
        }
        if (nb == 0) break;  // the stream is exhausted and nothing is queued
        if (nb < GR && lane >= nb && lane < GR) {
            // last, partial group: make the unused rows inert (finite zero records and colours, an index no pixel
            // reaches) -- phase 1 evaluates rows up to the next multiple of 8, the MMAs all sixteen
            s_qrec[2 * lane] = make_float4(0.f, 0.f, 0.f, 0.f);
            s_qrec[2 * lane + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int k4 = 0; k4 < D0 / 4; ++k4) *reinterpret_cast<float4 *>(s_qcol + lane * D0 + 4 * k4) = make_float4(0.f, 0.f, 0.f, 0.f);
            s_qidx[lane] = INT_MAX;
        }
        phase1();
        phase2();
    }
}

template <bool DEPTH, bool TC1>
static int launch_tc(const SlabArgs &a, const float *ra, const int32_t *li, const float *ad, const float *vrc,
                     const float *vra, float *vm, float *vc, float *vcol, float *vo, float *vd, cudaStream_t st) {
    constexpr size_t smem = TcCfg<DEPTH>::smem_bytes();
    if (cudaFuncSetAttribute(blend_bwd_slab_tc_kernel<DEPTH, TC1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)smem) != cudaSuccess)
        return 1;
    const int grid = a.C * a.tile_w * a.tile_h;
    blend_bwd_slab_tc_kernel<DEPTH, TC1><<<grid, kSlabThreads, smem, st>>>(a, ra, li, ad, vrc, vra, vm, vc, vcol, vo, vd);
    return 0;
}

// variant: 1 = phase 2 on the tensor cores, 2 = both phases.  Returns -1 when (D0, alignment) is not served here.
int launch_blend_bwd_slab_tc(int variant, int D0, bool depth, const SlabArgs &a, const float *ra, const int32_t *li,
                             const float *ad, const float *vrc, const float *vra, float *vm, float *vc, float *vcol,
                             float *vo, float *vd, cudaStream_t st) {
    if (D0 != kD0) return -1;
    // vectorised reductions: 8-byte aligned colour / means2d gradient rows
    if ((reinterpret_cast<uintptr_t>(vcol) & 7u) || (reinterpret_cast<uintptr_t>(vm) & 7u) || (a.colors_cs & 1)) return -1;
    if (variant == 2)
        return depth ? launch_tc<true, true>(a, ra, li, ad, vrc, vra, vm, vc, vcol, vo, vd, st)
                     : launch_tc<false, true>(a, ra, li, ad, vrc, vra, vm, vc, vcol, vo, vd, st);
    return depth ? launch_tc<true, false>(a, ra, li, ad, vrc, vra, vm, vc, vcol, vo, vd, st)
                 : launch_tc<false, false>(a, ra, li, ad, vrc, vra, vm, vc, vcol, vo, vd, st);
}

}  // namespace d4
