// blend_slab_bwd.cu -- row a11 of SURVEY.md section 8: backward of the per-tile alpha compositing (gsplat
// rasterize_to_pixels bwd + ED normalisation backward) over the packed record slabs of slab.cuh.
//
// Arithmetic: the grouped formulation of blend_bwd_gp.cu --
//   phase 1 (lane = pixel):    walk the warp's hits back to front, four per trip; exponent, validity and <c_g, v_out>
//       of the four hits are independent straight-line code, the (T, S) recurrence is one FMUL + one FFMA deep per
//       hit; per hit only fac = alpha*T and v_sigma = dL/dsigma are parked as one row of a per-warp
//       [16 Gaussians x 32 pixels] shared-memory tile;
//   phase 2 (lane = Gaussian): lane (g, half) sweeps 16 pixels for ITS Gaussian, accumulates the D+6 sums privately
//       (v_colors = sum_p fac * v_out[p][:], conic / xy / opacity sums as second moments of v_sigma about the centre),
//       joins the halves with one exchange and issues one global RED.ADD per (warp, Gaussian, value).
// Data movement -- what changed against blend_bwd_gp.cu.  The tile's records and colour rows arrive through a ring of
// shared-memory stages filled by a PRODUCER warp (cp.async.bulk + 16-byte cp.async, mbarrier completion) in
// back-to-front order; the eight consumer warps synchronise with the ring only (full / empty mbarriers), never with
// each other.  Which records a warp visits comes from the forward's per-(chunk, warp) hit words
// (SlabArgs::hit_bits): one uniform 4-byte load per 32 records, prefetched a chunk ahead.  A warp does not work on
// the ring in place: lane L copies record L of the chunk (if it is a hit) with its colour row into the warp's own
// 16-row queue -- 12 shared-memory instructions per chunk, however many hits -- and gives the stage back at once; phases
// 1 and 2 then run over the 16 queued rows at static addresses.  Fast warps are never held back by slow ones beyond the
// depth of the ring, and a stage is held for a few dozen cycles instead of for the whole evaluation.
#include <limits.h>

#include "slab.cuh"

namespace d4 {

__device__ __forceinline__ float rcp_approx_s(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

constexpr int kParkStride = 33;  // (fac, v_sigma) park tile row stride: 32 pixels + 1 (conflict-free both ways)
constexpr int kRows = 16;        // rows of the per-warp queue == rows of the park tile

template <int D0, bool DEPTH>
struct BwdSlabCfg {
    static constexpr int D = D0 + (DEPTH ? 1 : 0);
    static constexpr int V = D + 6;
    static constexpr int NW = kSlabConsumers;
    // per consumer warp: queued records + colour rows + record indices, (fac, v_sigma) park tile
    static constexpr size_t warp_bytes() {
        return sizeof(float4) * kRows * 2 + sizeof(float) * kRows * D0 + sizeof(int32_t) * kRows +
               sizeof(float) * 2 * kRows * kParkStride;
    }
    static constexpr size_t fixed_bytes() {
        return sizeof(float) * kBlendThreads * D0 + (DEPTH ? sizeof(float) * kBlendThreads : 0)  // v_out of the tile
               + NW * warp_bytes() + 128;                                                        // + mbarriers
    }
    static constexpr size_t stage_bytes() { return (size_t)kSlabChunk * (32 + 4 * D0); }
    // ring depth: what still lets kCtas CTAs share an SM (228 KB, 1 KB reserved per CTA)
    static constexpr int stages_for(int ncta) {
        const size_t budget = (228 * 1024) / ncta - 1024 - 64;
        int s = (int)((budget - fixed_bytes()) / stage_bytes());
        return s > 8 ? 8 : s;
    }
    static constexpr int kCtas = D0 <= 16 ? 3 : 2;
    static constexpr int kStages = stages_for(kCtas);
    static_assert(kStages >= 2, "the ring needs two stages");
    static constexpr size_t smem_bytes() { return fixed_bytes() + kStages * stage_bytes(); }
};

template <int D0, bool DEPTH>
__global__ void __launch_bounds__(kSlabThreads, (BwdSlabCfg<D0, DEPTH>::kCtas))
blend_bwd_slab_kernel(SlabArgs a, const float *__restrict__ render_alphas, const int32_t *__restrict__ last_ids,
                      const float *__restrict__ acc_depth, const float *__restrict__ v_render_colors,
                      const float *__restrict__ v_render_alphas, float *__restrict__ v_means2d,
                      float *__restrict__ v_conics, float *__restrict__ v_colors, float *__restrict__ v_opacities,
                      float *__restrict__ v_depths) {
    using Cfg = BwdSlabCfg<D0, DEPTH>;
    constexpr int D = Cfg::D, V = Cfg::V, NW = Cfg::NW, S = Cfg::kStages, CH = kSlabChunk;
    constexpr int U = 4, GR = kRows;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float4 *s_rec = reinterpret_cast<float4 *>(smem_raw);                        // [S][CH][2]
    float *s_col = reinterpret_cast<float *>(s_rec + S * CH * 2);                // [S][CH][D0]
    float *s_vout = s_col + S * CH * D0;                                         // [256 pixels][D0], pixel == thread
    float *s_vd = s_vout + kBlendThreads * D0;                                   // [256] depth cotangent (DEPTH)
    float4 *s_qrec_all = reinterpret_cast<float4 *>(s_vd + (DEPTH ? kBlendThreads : 0));  // [NW][GR][2]  queued records
    float *s_qcol_all = reinterpret_cast<float *>(s_qrec_all + NW * GR * 2);     // [NW][GR][D0]  queued colour rows
    float *s_tiles = s_qcol_all + NW * GR * D0;                                  // [NW][2][GR][kParkStride]
    int32_t *s_qidx_all = reinterpret_cast<int32_t *>(s_tiles + NW * 2 * GR * kParkStride);  // [NW][GR] record indices
    uint64_t *s_full = reinterpret_cast<uint64_t *>(s_qidx_all + NW * GR);       // [S]
    uint64_t *s_empty = s_full + S;                                              // [S]
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

    // ---- per-pixel state
    constexpr int D2 = D0 / 2;
    float2 v2[D2];  // colour cotangent as fp32x2 pairs
    float vd = 0.f;  // depth cotangent (after the ED normalisation backward)
    float T_final = 1.f, v_ra = 0.f, bgdot = 0.f;
    int32_t bin_final = -1;
    if (!producer) {
        float v_out[D];
        if (inside) {
            const float alpha_px = render_alphas[pid];
            T_final = 1.0f - alpha_px;
            bin_final = last_ids[pid];
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
#pragma unroll
        for (int k2 = 0; k2 < D2; ++k2) v2[k2] = make_float2(v_out[2 * k2], v_out[2 * k2 + 1]);
        if constexpr (DEPTH) vd = v_out[D - 1];
        // phase 2 reads every pixel's cotangent as a warp-wide broadcast
        float *vo = s_vout + tid * D0;
#pragma unroll
        for (int k4 = 0; k4 < D0 / 4; ++k4)
            *reinterpret_cast<float4 *>(vo + 4 * k4) = make_float4(v_out[4 * k4], v_out[4 * k4 + 1], v_out[4 * k4 + 2], v_out[4 * k4 + 3]);
        if constexpr (DEPTH) s_vd[tid] = vd;
    }
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

    // ---------------------------------------------------------------------------------------- consumer warps
    // constant part of dL/dalpha_i * (1 - alpha_i):  T_final * (v_alpha_out - bg.v_out)
    const float tail = T_final * (v_ra - bgdot);
    float T = T_final;
    float Sacc = 0.f;  // sum_{j>i} <c_j, v_out> alpha_j T_j

    float4 *s_qrec = s_qrec_all + w * GR * 2;
    float *s_qcol = s_qcol_all + w * GR * D0;
    int32_t *s_qidx = s_qidx_all + w * GR;
    float *s_fac = s_tiles + w * 2 * GR * kParkStride;
    float *s_vs = s_fac + GR * kParkStride;
    const float bx0 = (float)(tx * kTile + (w & 1) * 8) + 0.5f;
    int nb = 0;  // rows queued (warp-uniform)

    // ---- phase 1 over the nb queued rows (lane = pixel), U rows per trip
    auto phase1 = [&]() {
        __syncwarp();
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
                    // s = <c_g, v_out>, independent partial sums
                    const float *cp = s_qcol + r * D0;
                    constexpr int PPSm = D0 / 4 - 1;
                    const int key = slab_key<D0>(h + u);  // == slab_key(r): r8 is a multiple of 8
                    float2 sa = make_float2(0.f, 0.f), sb = make_float2(0.f, 0.f);
#pragma unroll
                    for (int k4 = 0; k4 < D0 / 4; ++k4) {
                        const float4 cv = *reinterpret_cast<const float4 *>(cp + 4 * ((k4 ^ key) & PPSm));
                        sa = __ffma2_rn(make_float2(cv.x, cv.y), v2[2 * k4], sa);
                        sb = __ffma2_rn(make_float2(cv.z, cv.w), v2[2 * k4 + 1], sb);
                    }
                    float s = (sa.x + sa.y) + (sb.x + sb.y);
                    if constexpr (DEPTH) s = fmaf(cn.w, vd, s);
                    sd[u] = s;
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    // alpha == 0 (pixel not taking part): ra = 1, T and S unchanged, fac = v_sigma = 0
                    const float ra = rcp_approx_s(1.0f - al[u]);  // 1 - alpha in [0.001, 1]: MUFU.RCP is within 1 ulp here
                    T *= ra;
                    const float fac = al[u] * T;
                    const float v_alpha = sd[u] * T - (Sacc - tail) * ra;
                    Sacc = fmaf(sd[u], fac, Sacc);
                    const float vs = ar[u] != 0.f ? -ar[u] * v_alpha : 0.f;
                    s_fac[(r0 + u) * kParkStride + lane] = fac;
                    s_vs[(r0 + u) * kParkStride + lane] = vs;
                }
            }
        }
    };

    // ---- phase 2: lane (pg, part) sweeps 16 pixels for queued row pg, pixels [part*16, part*16 + 16) of the warp's
    // 8x4 block (pixel p sits at x = p & 7, y = p >> 3)
    auto phase2 = [&]() {
        constexpr int GS = GR;
        static_assert(GS == 16, "two half-warps per row");
        __syncwarp();
        const int pg = lane & (GS - 1), part = lane / GS;
        const float by0 = (float)(ty * kTile + (w >> 1) * 4 + (part * GS) / 8) + 0.5f;
        const float *p2_vo = s_vout + (w * 32 + part * GS) * D0;
        [[maybe_unused]] const float *p2_vd = s_vd + w * 32 + part * GS;
        const bool rowok = pg < nb;  // rows >= nb hold stale data: computed on, never stored
        const float4 g0 = s_qrec[2 * pg], cn = s_qrec[2 * pg + 1];
        const float Xl = g0.x - bx0, Yl = g0.y - by0;
        float2 acc[D2];
#pragma unroll
        for (int k = 0; k < D2; ++k) acc[k] = make_float2(0.f, 0.f);
        [[maybe_unused]] float accd = 0.f;
        float axx = 0.f, axy = 0.f, ayy = 0.f, ax = 0.f, ay = 0.f, a0 = 0.f;
        const float *fr = s_fac + pg * kParkStride + part * GS;
        const float *vr = s_vs + pg * kParkStride + part * GS;
#pragma unroll
        for (int q = 0; q < GS; ++q) {
            const float fac = fr[q], vs = vr[q];
            const float2 f2 = make_float2(fac, fac);
#pragma unroll
            for (int k4 = 0; k4 < D0 / 4; ++k4) {
                const float4 v = *reinterpret_cast<const float4 *>(p2_vo + q * D0 + 4 * k4);
                acc[2 * k4] = __ffma2_rn(f2, make_float2(v.x, v.y), acc[2 * k4]);
                acc[2 * k4 + 1] = __ffma2_rn(f2, make_float2(v.z, v.w), acc[2 * k4 + 1]);
            }
            if constexpr (DEPTH) accd = fmaf(fac, p2_vd[q], accd);
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
        //   conic (a, b, c) = (-2A', -B', -2C') / log2e ;  1 / opacity = exp2(-L)
        float r[V];
#pragma unroll
        for (int k = 0; k < D0; ++k) r[k] = (k & 1) ? acc[k >> 1].y : acc[k >> 1].x;
        if constexpr (DEPTH) r[D - 1] = accd;
        const float ka = cn.x * (-2.0f / kLog2e), kb = cn.y * (-1.0f / kLog2e), kc = cn.z * (-2.0f / kLog2e);
        r[D + 0] = 0.5f * axx;
        r[D + 1] = axy;
        r[D + 2] = 0.5f * ayy;
        r[D + 3] = fmaf(ka, ax, kb * ay);
        r[D + 4] = fmaf(kb, ax, kc * ay);
        r[D + 5] = -ex2_approx(-g0.z) * a0;
        // join the two halves of every row in part 0, which then issues one global reduction per (warp, Gaussian, value)
        // at addresses that are static per value (no per-lane target selection)
#pragma unroll
        for (int k = 0; k < V; ++k) r[k] += __shfl_xor_sync(0xffffffffu, r[k], GS);
        if (part == 0 && rowok) {
            const int32_t gl = (int32_t)(__float_as_uint(g0.w) & kRecIdMask);
            const int64_t g = (int64_t)c * a.G + gl;
            float *vcol = v_colors + c * a.colors_cs + (int64_t)gl * D0;
#pragma unroll
            for (int k = 0; k < D0; ++k)
                if (r[k] != 0.f) atomicAdd(vcol + k, r[k]);
            if constexpr (DEPTH) {
                if (r[D - 1] != 0.f) atomicAdd(v_depths + g, r[D - 1]);
            }
#pragma unroll
            for (int k = 0; k < 3; ++k)
                if (r[D + k] != 0.f) atomicAdd(v_conics + 3LL * g + k, r[D + k]);
#pragma unroll
            for (int k = 0; k < 2; ++k)
                if (r[D + 3 + k] != 0.f) atomicAdd(v_means2d + 2LL * g + k, r[D + 3 + k]);
            if (r[D + 5] != 0.f) atomicAdd(v_opacities + gl, r[D + 5]);
        }
        __syncwarp();
        nb = 0;
    };

    // ---- the ring, consumer side: chunk c_hi first
    const int kw_hi = warp_bin_final - seg_start;       // last record of the warp, relative (< 0: none)
    const int k_warp_hi = kw_hi < 0 ? -1 : kw_hi / CH;  // chunks above hold no hit words written for this warp
    const int64_t hb_base = ((int64_t)(seg_start >> 5) + ct) * NW + w;
    // Hit words of this warp for 32 chunks at a time: lane i holds the word of chunk k_top - i (one strided load per
    // 32 chunks instead of a dependent global load per chunk).  Bits behind the warp's last contributing record are
    // dropped: the forward also marks a record on which its last pixel saturated (and stopped), and such a record may
    // not even be streamed.
    auto load_words = [&](int k_top) -> uint32_t {
        const int kk = k_top - lane;
        if (kk < 0 || kk > k_warp_hi) return 0u;
        uint32_t wd = __ldg(a.hit_bits + hb_base + (int64_t)kk * NW);
        if (kk == k_warp_hi) wd &= 0xffffffffu >> (31 - (kw_hi & 31));
        return wd;
    };
    int stage = 0, phase = 0;
    int k = c_hi + 1;         // chunk being drained (none yet)
    int k_top = c_hi;         // chunk whose hit word lane 0 holds
    uint32_t words = load_words(k_top);
    bool holding = false;     // the ring stage of chunk k is still in use
    uint32_t bits = 0u;       // hits of chunk k not yet queued
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
            // last, partial group: the rows up to the next multiple of U are evaluated too -- make them inert
            // (finite zero records and colours, an index no pixel reaches)
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

template <int D0, bool DEPTH>
static int launch_bwd_slab(const SlabArgs &a, const float *ra, const int32_t *li, const float *ad, const float *vrc,
                           const float *vra, float *vm, float *vc, float *vcol, float *vo, float *vd, cudaStream_t st) {
    constexpr size_t smem = BwdSlabCfg<D0, DEPTH>::smem_bytes();
    if (cudaFuncSetAttribute(blend_bwd_slab_kernel<D0, DEPTH>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)smem) != cudaSuccess)
        return 1;
    const int grid = a.C * a.tile_w * a.tile_h;
    blend_bwd_slab_kernel<D0, DEPTH><<<grid, kSlabThreads, smem, st>>>(a, ra, li, ad, vrc, vra, vm, vc, vcol, vo, vd);
    return 0;
}

int launch_blend_bwd_slab(int D0, bool depth, const SlabArgs &a, const float *ra, const int32_t *li, const float *ad,
                          const float *vrc, const float *vra, float *vm, float *vc, float *vcol, float *vo, float *vd,
                          cudaStream_t st) {
#define X(n)                                                                                   \
    case n:                                                                                    \
        return depth ? launch_bwd_slab<n, true>(a, ra, li, ad, vrc, vra, vm, vc, vcol, vo, vd, st)  \
                     : launch_bwd_slab<n, false>(a, ra, li, ad, vrc, vra, vm, vc, vcol, vo, vd, st);
    switch (D0) {
        X(4) X(8) X(16) X(32)
        default: return -1;
    }
#undef X
}

}  // namespace d4

using namespace d4;

static int blend_bwd_slab_any(const char *name, int variant, const void *recs, const int32_t *tile_offsets,
                              const int32_t *rec_counts, const float *colors, int64_t colors_cam_stride,
                              const float *backgrounds, int C, int G, int D0, int with_depth, int width, int height,
                              int tile_size, int tile_w, int tile_h, int normalize_depth, const float *render_alphas,
                              const int32_t *last_ids, const float *acc_depth, const float *v_render_colors,
                              const float *v_render_alphas, const uint32_t *hit_bits, float *v_means2d, float *v_conics,
                              float *v_colors, float *v_opacities, float *v_depths, d4_stream_t stream) {
    SlabArgs a{(const float4 *)recs, tile_offsets, rec_counts, colors, colors_cam_stride, backgrounds,
               const_cast<uint32_t *>(hit_bits), C, G, width, height, tile_w, tile_h, normalize_depth};
    if (int rc = check_slab_args(name, a, D0, tile_size)) return rc;
    D4_CHECK_ARG(render_alphas && last_ids && v_render_colors && v_render_alphas && hit_bits && v_means2d && v_conics &&
                     v_opacities && v_colors,
                 "%s: null pointer", name);
    D4_CHECK_ARG(!with_depth || v_depths, "%s: v_depths required with the depth channel", name);
    D4_CHECK_ARG(!normalize_depth || (with_depth && acc_depth), "%s: normalize_depth needs the depth channel and acc_depth", name);
    D4_CHECK_ARG(variant >= 0 && variant <= 2, "%s: variant must be 0 (fp32 pipe), 1 or 2 (tensor cores)", name);
    int rc = -1;
    if (variant > 0)  // served for the 16-colour records with 8-byte aligned gradient rows, otherwise falls through
        rc = launch_blend_bwd_slab_tc(variant, D0, with_depth != 0, a, render_alphas, last_ids, acc_depth, v_render_colors,
                                      v_render_alphas, v_means2d, v_conics, v_colors, v_opacities, v_depths,
                                      as_stream(stream));
    if (rc < 0)
        rc = launch_blend_bwd_slab(D0, with_depth != 0, a, render_alphas, last_ids, acc_depth, v_render_colors,
                                   v_render_alphas, v_means2d, v_conics, v_colors, v_opacities, v_depths,
                                   as_stream(stream));
    if (rc != 0) {
        set_error("%s: %s", name, rc < 0 ? "channel count not built" : "kernel configuration failed");
        return rc < 0 ? 2 : 1;
    }
    D4_CHECK_LAUNCH(name);
    return 0;
}

extern "C" int d4_blend_bwd_slab(const void *recs, const int32_t *tile_offsets, const int32_t *rec_counts,
                                 const float *colors, int64_t colors_cam_stride, const float *backgrounds, int C, int G,
                                 int D0, int with_depth, int width, int height, int tile_size, int tile_w, int tile_h,
                                 int normalize_depth, const float *render_alphas, const int32_t *last_ids,
                                 const float *acc_depth, const float *v_render_colors, const float *v_render_alphas,
                                 const uint32_t *hit_bits, float *v_means2d, float *v_conics, float *v_colors,
                                 float *v_opacities, float *v_depths, d4_stream_t stream) {
    return blend_bwd_slab_any("d4_blend_bwd_slab", D4_BLEND_BWD_DEFAULT_VARIANT, recs, tile_offsets, rec_counts, colors,
                              colors_cam_stride, backgrounds, C, G, D0, with_depth, width, height, tile_size, tile_w,
                              tile_h, normalize_depth, render_alphas, last_ids, acc_depth, v_render_colors,
                              v_render_alphas, hit_bits, v_means2d, v_conics, v_colors, v_opacities, v_depths, stream);
}

extern "C" int d4_blend_bwd_slab_default_variant(void) { return D4_BLEND_BWD_DEFAULT_VARIANT; }

extern "C" int d4_blend_bwd_slab_variant(const void *recs, const int32_t *tile_offsets, const int32_t *rec_counts,
                                         const float *colors, int64_t colors_cam_stride, const float *backgrounds,
                                         int C, int G, int D0, int with_depth, int width, int height, int tile_size,
                                         int tile_w, int tile_h, int normalize_depth, const float *render_alphas,
                                         const int32_t *last_ids, const float *acc_depth, const float *v_render_colors,
                                         const float *v_render_alphas, const uint32_t *hit_bits, float *v_means2d,
                                         float *v_conics, float *v_colors, float *v_opacities, float *v_depths,
                                         int variant, d4_stream_t stream) {
    return blend_bwd_slab_any("d4_blend_bwd_slab_variant", variant, recs, tile_offsets, rec_counts, colors,
                              colors_cam_stride, backgrounds, C, G, D0, with_depth, width, height, tile_size, tile_w,
                              tile_h, normalize_depth, render_alphas, last_ids, acc_depth, v_render_colors,
                              v_render_alphas, hit_bits, v_means2d, v_conics, v_colors, v_opacities, v_depths, stream);
}
