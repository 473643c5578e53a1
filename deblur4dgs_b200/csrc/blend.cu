// blend.cu -- rows a10 / a11 of SURVEY.md section 8: per-tile alpha compositing.
//   d4_blend_fwd : gsplat rasterize_to_pixels fwd (+ depth channel, + ED normalisation)
//   d4_blend_bwd : gsplat rasterize_to_pixels bwd (+ ED normalisation backward)
//
// One CTA per (camera, 16x16 tile), 256 threads = 256 pixels; each WARP owns a
// compact 8x4 pixel block (not a 16x2 strip) so that a Gaussian's footprint
// touches as few warps as possible and whole warps skip it after one vote.
// Gaussians of the tile are staged 256 at a time into shared memory
// (xy+opacity, conic, D colour channels, id); the per-pair loop reads them as
// warp-wide broadcasts (blend_common.cuh).
//
// Backward: the default is the grouped kernel of blend_bwd_gp.cu.  This file keeps
// the warp-butterfly formulation (D4_BWD=shfl): per (warp, Gaussian) the D+6 partial
// sums of the 32 pixels are reduced with a TRANSPOSING butterfly (31 shuffles for up
// to 32 values instead of 5 per value), added to a per-CTA shared-memory accumulator
// and flushed to HBM once per (tile, Gaussian).  Both carry the per-pixel recurrences
// as scalars:
//   s_i = <c_i, v_out>,  S = sum_{j>i} s_j alpha_j T_j,
//   dL/dalpha_i = T_i s_i - (S + T_final (bg.v_out - v_alpha_out)) / (1 - alpha_i)
// which is algebraically gsplat's per-channel buffer[] form with D fewer
// registers and D fewer FMAs per pair.
//
// These kernels are bound by fp32 issue, MUFU.EX2 and latency chains, not by
// HBM: algorithmic bytes per (pixel, Gaussian) pair are ~0.4.
#include "blend_common.cuh"

namespace d4 {

// ----------------------------------------------------------------------------- forward
// kMasks: also record, per intersection, which of the tile's 8 pixel blocks passed the alpha test (BlendArgs::hit_masks)
template <int D, bool kMasks>
__global__ void __launch_bounds__(kBlendThreads, (D <= 9 ? 5 : (D <= 17 ? 4 : 1)))
blend_fwd_kernel(BlendArgs a, float *__restrict__ render_colors, float *__restrict__ render_alphas,
                 int32_t *__restrict__ last_ids, float *__restrict__ acc_depth) {
    constexpr int DS = BlendCfg<D>::DS;
    constexpr int DP = BlendCfg<D>::DP;
    __shared__ float4 s_geom[kBatch];
    __shared__ float4 s_conic[kBatch];
    __shared__ __align__(16) float s_col[kBatch * (DS > DP ? DS : DP)];
    __shared__ uint32_t s_mask[kBatch];
    __shared__ uint32_t s_hitw[kMasks ? kBlendThreads / 32 : 1][kBatch / 32];  // per warp: slots of the batch past the alpha test

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
    const int64_t range_end = (ct == a.C * n_tiles - 1) ? a.n_isects : (int64_t)a.tile_offsets[ct + 1];
    const int num_batches = (int)((range_end - range_start + kBatch - 1) / kBatch);
    // hit masks of a finished batch: thread t gathers bit (t & 31) of word (t >> 5) of every warp
    auto write_hit_masks = [&](int64_t start, int size) {
        if (tid < size) {
            uint32_t m = 0u;
#pragma unroll
            for (int ww = 0; ww < kBlendThreads / 32; ++ww) m |= ((s_hitw[ww][tid >> 5] >> (tid & 31)) & 1u) << ww;
            a.hit_masks[start + tid] = (uint8_t)m;
        }
    };
    int prev_b = 0, prev_size = 0;  // last processed batch whose masks are still to be written

    float T = 1.0f;
    int32_t cur_idx = 0;
    bool done = !inside;
    // accumulators as fp32x2 pairs: Blackwell's FFMA2 retires two fp32 FMAs per issue slot, and this
    // kernel is bound by issue slots
    constexpr int D2 = (D + 1) / 2;
    float2 out2[D2];
#pragma unroll
    for (int k = 0; k < D2; ++k) out2[k] = make_float2(0.f, 0.f);

    // this thread's slot of batch b is fetched into registers at the end of batch b-1: the loads fly while the
    // warp waits at the batch barrier for the slowest warp of the tile
    StagedRec<D> rec;
    stage_load<D>(rec, a, c, range_start + tid, num_batches > 0 && range_start + tid < range_end);
    for (int b = 0; b < num_batches; ++b) {
        const bool all_done = __syncthreads_count(done) >= kBlendThreads;
        if constexpr (kMasks) {
            if (prev_size > 0) write_hit_masks(range_start + (int64_t)kBatch * prev_b, prev_size);
            prev_size = 0;
        }
        if (all_done) break;
        const int64_t batch_start = range_start + (int64_t)kBatch * b;
        stage_store<D>(rec, tid, tx * kTile, ty * kTile, s_geom, s_conic, s_col, s_mask);
        __syncthreads();
        if constexpr (kMasks) {
            if (lane < kBatch / 32) s_hitw[w][lane] = 0u;
            __syncwarp();
        }
        const int batch_size = (int)min((int64_t)kBatch, range_end - batch_start);
        bool warp_done = __all_sync(0xffffffffu, done);
        uint32_t hitbits = 0u;  // slots of the current 32-slot chunk with at least one pixel past the alpha test
        // one Gaussian of the hit list for this pixel; returns true when the whole warp is finished
        auto composite = [&](int t, float power, float L) -> bool {
            const float alpha = fminf(kAlphaMax, ex2_approx(power));
            const bool valid = !done && power <= L && alpha >= kAlphaMin;
            if (!__any_sync(0xffffffffu, valid)) return false;
            if constexpr (kMasks) hitbits |= 1u << (t & 31);
            if (valid) {
                const float next_T = T * (1.0f - alpha);
                if (next_T <= kTMin) {
                    done = true;
                } else {
                    const float vis = alpha * T;
                    const float2 vis2 = make_float2(vis, vis);
                    const float *cp = s_col + t * DS;
#pragma unroll
                    for (int k4 = 0; k4 < DS / 4; ++k4) {  // DS = D rounded up to 4: the pad lanes are never stored
                        const float4 cv = *reinterpret_cast<const float4 *>(cp + 4 * k4);
                        if (2 * k4 < D2) out2[2 * k4] = __ffma2_rn(make_float2(cv.x, cv.y), vis2, out2[2 * k4]);
                        if (2 * k4 + 1 < D2) out2[2 * k4 + 1] = __ffma2_rn(make_float2(cv.z, cv.w), vis2, out2[2 * k4 + 1]);
                    }
                    cur_idx = (int32_t)(batch_start + t);
                    T = next_T;
                }
            }
            return __all_sync(0xffffffffu, done);
        };
        for (int chunk = 0; chunk * 32 < batch_size && !warp_done; ++chunk) {
            uint32_t bits = __ballot_sync(0xffffffffu, (s_mask[chunk * 32 + lane] >> w) & 1u);
            while (bits) {
                // two hits per trip: both exponents are evaluated before either is composited (ILP for the
                // LDS -> FMA -> MUFU chain, which is what the issue slots were waiting on)
                const int ta = chunk * 32 + __ffs(bits) - 1;
                bits &= bits - 1;
                const bool has_b = bits != 0u;
                const int tb = has_b ? chunk * 32 + __ffs(bits) - 1 : ta;
                bits &= bits - 1;  // no-op when bits == 0
                const float4 ga = s_geom[ta], ca = s_conic[ta];
                const float4 gb = s_geom[tb], cb = s_conic[tb];
                const float dxa = ga.x - px, dya = ga.y - py, dxb = gb.x - px, dyb = gb.y - py;
                const float pa = fmaf(ca.z * dya, dya, fmaf(fmaf(ca.y, dya, ca.x * dxa), dxa, ga.z));
                const float pb = fmaf(cb.z * dyb, dyb, fmaf(fmaf(cb.y, dyb, cb.x * dxb), dxb, gb.z));
                if (composite(ta, pa, ga.z)) { warp_done = true; break; }
                if (has_b && composite(tb, pb, gb.z)) { warp_done = true; break; }
            }
            if constexpr (kMasks) {
                if (lane == 0) s_hitw[w][chunk] = hitbits;
                hitbits = 0u;
            }
        }
        if constexpr (kMasks) {
            prev_b = b;
            prev_size = batch_size;
        }
        {
            const int64_t next = batch_start + kBatch + tid;
            stage_load<D>(rec, a, c, next, b + 1 < num_batches && next < range_end);
        }
    }

    if constexpr (kMasks) {
        if (prev_size > 0) {  // masks of the last batch that was processed (prev_size is uniform for the CTA)
            __syncthreads();
            write_hit_masks(range_start + (int64_t)kBatch * prev_b, prev_size);
        }
    }

    // epilogue: background, ED normalisation, coalesced store through shared memory
    float out[D];
#pragma unroll
    for (int k = 0; k < D; ++k) out[k] = (k & 1) ? out2[k >> 1].y : out2[k >> 1].x;
    const float alpha_out = 1.0f - T;
    if (a.backgrounds) {
        const int d0 = a.depths ? D - 1 : D;
#pragma unroll
        for (int k = 0; k < D; ++k)
            if (k < d0) out[k] = fmaf(T, __ldg(a.backgrounds + (int64_t)c * a.D0 + k), out[k]);
    }
    if (inside) {
        render_alphas[pid] = alpha_out;
        last_ids[pid] = cur_idx;
        if (a.normalize_depth) {
            acc_depth[pid] = out[D - 1];
            out[D - 1] = out[D - 1] / fmaxf(alpha_out, 1e-10f);
        }
    }
    __syncthreads();  // everyone is done reading s_col
    {
        float *dst = s_col + (ly * kTile + lx) * DP;
#pragma unroll
        for (int k = 0; k < D; ++k) dst[k] = out[k];
    }
    __syncthreads();
    // every tile row is one contiguous run of 16*D floats in the channels-last image: thread `tid` owns column
    // positions tid, tid + 256, ... of that run (pixel / channel split computed once, not per element)
    constexpr int row_elems = kTile * D;
    constexpr int kCols = (row_elems + kBlendThreads - 1) / kBlendThreads;
    int src_off[kCols];
    bool col_ok[kCols];
#pragma unroll
    for (int q = 0; q < kCols; ++q) {
        const int col = tid + q * kBlendThreads;
        const int pxl = col / D, k = col - pxl * D;
        src_off[q] = pxl * DP + k;
        col_ok[q] = col < row_elems && (tx * kTile + pxl) < a.width;
    }
    const int rows = min(kTile, a.height - ty * kTile);
    float *dst_row = render_colors + (((int64_t)c * a.height + ty * kTile) * a.width + tx * kTile) * D + tid;
    for (int r = 0; r < rows; ++r) {
#pragma unroll
        for (int q = 0; q < kCols; ++q)
            if (col_ok[q]) dst_row[q * kBlendThreads] = s_col[r * kTile * DP + src_off[q]];
        dst_row += (int64_t)a.width * D;
    }
}

// ----------------------------------------------------------------------------- backward
// Transposing butterfly: v[0..NV) per lane -> v[0] = sum over the warp of value (lane & (NV-1)).
template <int NV>
__device__ __forceinline__ void warp_transpose_reduce(float (&v)[NV], int lane) {
#pragma unroll
    for (int h = NV / 2; h >= 1; h >>= 1) {
        const bool upper = (lane & h) != 0;
#pragma unroll
        for (int i = 0; i < h; ++i) {
            const float lo = v[i], hi = v[i + h];
            const float send = upper ? lo : hi;
            const float keep = upper ? hi : lo;
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, h);
        }
    }
#pragma unroll
    for (int o = NV; o < 32; o <<= 1) v[0] += __shfl_xor_sync(0xffffffffu, v[0], o);
}

template <int V>
struct RedWidth {
    static constexpr int value = V <= 8 ? 8 : (V <= 16 ? 16 : 32);
};

constexpr int kBatchB = 256;  // Gaussians staged per batch

template <int D>
struct BwdCfg {
    static constexpr int VSH = D + 6;  // values reduced per (warp, Gaussian)
    static constexpr int RW = RedWidth<VSH>::value;
    static constexpr size_t smem_bytes() {
        return sizeof(float4) * 2 * kBatchB + sizeof(float) * kBatchB * (BlendCfg<D>::DS + (BlendCfg<D>::V | 1));
    }
};

template <int D>
__global__ void __launch_bounds__(kBlendThreads, 4)
blend_bwd_kernel(BlendArgs a, const float *__restrict__ render_alphas, const int32_t *__restrict__ last_ids,
                 const float *__restrict__ acc_depth, const float *__restrict__ v_render_colors,
                 const float *__restrict__ v_render_alphas, float *__restrict__ v_means2d,
                 float *__restrict__ v_conics, float *__restrict__ v_colors, float *__restrict__ v_opacities,
                 float *__restrict__ v_depths) {
    constexpr int DS = BlendCfg<D>::DS;
    constexpr int V = BlendCfg<D>::V;
    constexpr int VS = V | 1;  // odd accumulator stride
    constexpr int VSH = BwdCfg<D>::VSH, RW = BwdCfg<D>::RW;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4 *s_geom = reinterpret_cast<float4 *>(smem_raw);
    float4 *s_conic = s_geom + kBatchB;
    float *s_col = reinterpret_cast<float *>(s_conic + kBatchB);
    float *s_acc = s_col + kBatchB * DS;
    __shared__ int32_t s_max[kBlendThreads / 32];
    __shared__ uint32_t s_mask[kBatchB];
    __shared__ int32_t s_gid[2][kBatchB];  // flatten ids of the batch being processed / being flushed

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

    // per-pixel state
    float v_out[D];
    float T_final = 1.f, v_ra = 0.f;
    int32_t bin_final = -1;
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
    float bgdot = 0.f;
    if (a.backgrounds) {
        const int d0 = a.depths ? D - 1 : D;
#pragma unroll
        for (int k = 0; k < D; ++k)
            if (k < d0) bgdot = fmaf(__ldg(a.backgrounds + (int64_t)c * a.D0 + k), v_out[k], bgdot);
    }
    // v_out as fp32x2 pairs for the packed FFMA2 / FMUL2 path (pad lane zero; smem colour pads are finite)
    constexpr int D2 = (D + 1) / 2;
    float2 v2[D2];
#pragma unroll
    for (int k2 = 0; k2 < D2; ++k2) v2[k2] = make_float2(v_out[2 * k2], (2 * k2 + 1 < D) ? v_out[2 * k2 + 1] : 0.f);
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
    for (int k = 1; k < kBlendThreads / 32; ++k) block_bin_final = max(block_bin_final, s_max[k]);
    range_end = min(range_end, (int64_t)block_bin_final + 1);
    if (range_end <= range_start) return;
    const int num_batches = (int)((range_end - range_start + kBatchB - 1) / kBatchB);


    // flush of one batch's CTA-level sums: one global atomic per non-zero (Gaussian, value); zeroes as it goes
    auto flush_acc = [&](int n_slots, const int32_t *gids) {
        const int d0 = a.depths ? D - 1 : D;
        for (int e = tid; e < n_slots * V; e += kBlendThreads) {
            const int t = e / V, k = e - t * V;
            const float val = s_acc[t * VS + k];
            if (val == 0.f) continue;
            s_acc[t * VS + k] = 0.f;
            const int32_t g = gids[t];
            const int32_t gl = g - c * a.G;
            float *dst;
            if (k < d0) dst = v_colors + c * a.colors_cs + (int64_t)gl * a.D0 + k;
            else if (k < D) dst = v_depths + g;
            else if (k < D + 3) dst = v_conics + 3LL * g + (k - D);
            else if (k < D + 5) dst = v_means2d + 2LL * g + (k - D - 3);
            else dst = v_opacities + gl;
            atomicAdd(dst, val);
        }
    };
    for (int e = tid; e < kBatchB * VS; e += kBlendThreads) s_acc[e] = 0.f;
    int prev_size = 0;

    for (int b = 0; b < num_batches; ++b) {
        // (barrier C of the previous iteration has passed: every warp is done with batch b-1)
        const int64_t batch_end = range_end - 1 - (int64_t)kBatchB * b;  // slot 0 = furthest back
        const int batch_size = (int)min((int64_t)kBatchB, batch_end + 1 - range_start);
        if (tid < kBatchB) {
            const bool in_range = batch_end - tid >= range_start;
            stage_gaussian<D>(a, c, batch_end - tid, in_range, tid, tx * kTile, ty * kTile, s_geom, s_conic, s_col,
                              s_mask);
            s_gid[b & 1][tid] = in_range ? __ldg(a.flatten_ids + (batch_end - tid)) : 0;
        }
        flush_acc(prev_size, s_gid[(b & 1) ^ 1]);  // overlaps the staging loads of this batch
        prev_size = batch_size;
        __syncthreads();  // barrier B: staging visible, accumulators clean

        const int t0 = (int)max((int64_t)0, batch_end - (int64_t)warp_bin_final);
        for (int chunk = t0 >> 5; chunk * 32 < batch_size; ++chunk) {
            uint32_t bits = __ballot_sync(0xffffffffu, (s_mask[chunk * 32 + lane] >> w) & 1u);
            if (chunk == (t0 >> 5)) bits &= ~((1u << (t0 & 31)) - 1u);  // slots behind the warp's last contributor
            while (bits) {
                const int t = chunk * 32 + __ffs(bits) - 1;
                bits &= bits - 1;
                const float4 g0 = s_geom[t];
                const float4 cn = s_conic[t];
                const float dx = g0.x - px, dy = g0.y - py;
                const float power = fmaf(cn.z * dy, dy, fmaf(fmaf(cn.y, dy, cn.x * dx), dx, g0.z));
                const float araw = ex2_approx(power);  // opacity * exp(-sigma)
                const float alpha = fminf(kAlphaMax, araw);
                const bool valid = (batch_end - t <= (int64_t)bin_final) && power <= g0.z && alpha >= kAlphaMin;
                if (!__any_sync(0xffffffffu, valid)) continue;

                constexpr int RSZ = VSH <= 8 ? 8 : (VSH <= 16 ? 16 : (VSH <= 24 ? 24 : (VSH <= 32 ? 32 : 40)));
                float r[RSZ];
#pragma unroll
                for (int k = 0; k < RSZ; ++k) r[k] = 0.f;
                float fac = 0.f;
                if (valid) {
                    const float ra = __fdividef(1.0f, 1.0f - alpha);  // alpha <= 0.999: MUFU.RCP is within 1 ulp here
                    T *= ra;
                    fac = alpha * T;
                    // s = <c_g, v_out>, four independent partial sums (short dependency chain)
                    const float *cp = s_col + t * DS;
                    float2 sa = make_float2(0.f, 0.f), sb = make_float2(0.f, 0.f);
#pragma unroll
                    for (int k4 = 0; k4 < DS / 4; ++k4) {  // pad lanes of v2 are zero
                        const float4 cv = *reinterpret_cast<const float4 *>(cp + 4 * k4);
                        if (2 * k4 < D2) sa = __ffma2_rn(make_float2(cv.x, cv.y), v2[2 * k4], sa);
                        if (2 * k4 + 1 < D2) sb = __ffma2_rn(make_float2(cv.z, cv.w), v2[2 * k4 + 1], sb);
                    }
                    const float s = (sa.x + sa.y) + (sb.x + sb.y);
                    const float2 fac2 = make_float2(fac, fac);
#pragma unroll
                    for (int k2 = 0; k2 < D2; ++k2) {
                        const float2 p = __fmul2_rn(fac2, v2[k2]);
                        r[2 * k2] = p.x;
                        if (2 * k2 + 1 < D) r[2 * k2 + 1] = p.y;
                    }
                    const float v_alpha = s * T - (S - tail) * ra;
                    S = fmaf(s, fac, S);
                    if (araw <= kAlphaMax) {
                        // v_sigma = -araw * v_alpha; conic terms in the primed (base-2) scale:
                        //   a dx + b dy = -(2 A' dx + B' dy) / log2e
                        const float v_sigma = -araw * v_alpha;
                        const float vs2 = v_sigma * (-1.0f / kLog2e);
                        r[D + 0] = 0.5f * v_sigma * dx * dx;
                        r[D + 1] = v_sigma * dx * dy;
                        r[D + 2] = 0.5f * v_sigma * dy * dy;
                        r[D + 3] = vs2 * fmaf(2.0f * cn.x, dx, cn.y * dy);
                        r[D + 4] = vs2 * fmaf(cn.y, dx, 2.0f * cn.z * dy);
                        r[D + 5] = araw * cn.w * v_alpha;  // exp(-sigma) * v_alpha
                    }
                }
                // lane j < VSH ends up owning value j
                float mine;
                if constexpr (VSH > 16 && VSH <= 24) {
                    // 16 + 8 split: 16 + 9 shuffles instead of 31 for a padded 32-wide butterfly
                    float ra16[16], rb8[8];
#pragma unroll
                    for (int k = 0; k < 16; ++k) ra16[k] = r[k];
#pragma unroll
                    for (int k = 0; k < 8; ++k) rb8[k] = r[16 + k];
                    warp_transpose_reduce<16>(ra16, lane);
                    warp_transpose_reduce<8>(rb8, lane);
                    mine = lane < 16 ? ra16[0] : rb8[0];
                } else if constexpr (VSH > 32) {
                    static_assert(VSH <= 40, "too many shuffle-reduced values");
                    float ra32[32], rb8[8];
#pragma unroll
                    for (int k = 0; k < 32; ++k) ra32[k] = r[k];
#pragma unroll
                    for (int k = 0; k < 8; ++k) rb8[k] = r[32 + k];
                    warp_transpose_reduce<32>(ra32, lane);
                    warp_transpose_reduce<8>(rb8, lane);
                    mine = ra32[0];
                    if (lane < VSH - 32) {  // values 32.. are owned a second time by lanes 0..VSH-33
                        const int j = 32 + lane;
                        const int idx2 = j;
                        atomicAdd(&s_acc[t * VS + idx2], rb8[0]);
                    }
                } else {
                    warp_transpose_reduce<RW>(r, lane);
                    mine = r[0];
                }
                if (lane < (VSH < 32 ? VSH : 32)) {
                    const int idx = lane;
                    atomicAdd(&s_acc[t * VS + idx], mine);
                }
            }
        }
        __syncthreads();  // barrier C: every warp is done with batch b
    }
    flush_acc(prev_size, s_gid[(num_batches & 1) ^ 1]);
}

// Two backward formulations: the grouped kernel of blend_bwd_gp.cu (default) and the shuffle kernel of this file (kept
// as the A/B reference and as the second implementation the parity tests cross-check); the caller selects with the
// bwd_mode argument of d4_blend_bwd.  Two tensor-core formulations of the colour-gradient contraction (3xTF32
// mma.sync) were measured slower than both and removed: profiles/r01c_bwd_tc.md.
// ----------------------------------------------------------------------------- dispatch
template <int D>
static int launch_fwd(const BlendArgs &a, float *rc, float *ra, int32_t *li, float *ad, cudaStream_t st) {
    int grid = a.C * a.tile_w * a.tile_h;
    if (a.hit_masks) blend_fwd_kernel<D, true><<<grid, kBlendThreads, 0, st>>>(a, rc, ra, li, ad);
    else blend_fwd_kernel<D, false><<<grid, kBlendThreads, 0, st>>>(a, rc, ra, li, ad);
    return 0;
}
template <int D>
static int launch_bwd(int bwd_mode, const BlendArgs &a, const float *ra, const int32_t *li, const float *ad,
                      const float *vrc, const float *vra, float *vm, float *vc, float *vcol, float *vo, float *vd,
                      cudaStream_t st) {
    int grid = a.C * a.tile_w * a.tile_h;
    if (bwd_mode == 0) {
        const int rc = launch_blend_bwd_gp(D, a, ra, li, ad, vrc, vra, vm, vc, vcol, vo, vd, st);
        if (rc >= 0) return rc;
    }
    constexpr size_t smem = BwdCfg<D>::smem_bytes();
    // the opt-in is per device: set it on every launch (a per-process flag would miss a second GPU)
    if (cudaFuncSetAttribute(blend_bwd_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        return 1;
    blend_bwd_kernel<D><<<grid, kBlendThreads, smem, st>>>(a, ra, li, ad, vrc, vra, vm, vc, vcol, vo, vd);
    return 0;
}

}  // namespace d4

using namespace d4;

static int check_blend_args(const char *name, const BlendArgs &a, int tile_size) {
    D4_CHECK_ARG(tile_size == kTile, "%s: only tile_size 16 is built", name);
    D4_CHECK_ARG(a.C >= 1 && a.G >= 0 && a.D0 >= 0 && a.width > 0 && a.height > 0, "%s: bad sizes", name);
    D4_CHECK_ARG(a.tile_w == (a.width + kTile - 1) / kTile && a.tile_h == (a.height + kTile - 1) / kTile,
                 "%s: tile grid does not match the image size", name);
    D4_CHECK_ARG(a.tile_offsets && (a.flatten_ids || a.n_isects == 0), "%s: null pointer", name);
    D4_CHECK_ARG(a.n_isects == 0 || (a.means2d && a.conics && a.opacities && (a.colors || a.D0 == 0)),
                 "%s: null Gaussian arrays with a non-empty intersection list", name);
    D4_CHECK_ARG(((uintptr_t)a.means2d & 7) == 0, "%s: means2d must be 8-byte aligned", name);
    if ((a.D0 & 3) == 0 && a.D0 > 0)
        D4_CHECK_ARG(((uintptr_t)a.colors & 15) == 0 && (a.colors_cs & 3) == 0, "%s: colors must be 16-byte aligned", name);
    return 0;
}

extern "C" int d4_blend_fwd(const float *means2d, const float *conics, const float *opacities,
                            const float *colors, int64_t colors_cam_stride, const float *depths,
                            const float *backgrounds, int C, int G, int D0, int width, int height,
                            int tile_size, int tile_w, int tile_h, const int32_t *tile_offsets,
                            const int32_t *flatten_ids, int64_t n_isects, int normalize_depth,
                            float *render_colors, float *render_alphas, int32_t *last_ids, float *acc_depth,
                            uint8_t *hit_masks, d4_stream_t stream) {
    BlendArgs a{means2d, conics, opacities, colors, depths, backgrounds, colors_cam_stride, C, G, D0, width,
                height, tile_w, tile_h, tile_offsets, flatten_ids, n_isects, normalize_depth, hit_masks};
    if (int rc = check_blend_args("d4_blend_fwd", a, tile_size)) return rc;
    D4_CHECK_ARG(render_colors && render_alphas && last_ids, "d4_blend_fwd: null output");
    D4_CHECK_ARG(!normalize_depth || (depths && acc_depth), "d4_blend_fwd: normalize_depth needs depths and acc_depth");
    const int D = D0 + (depths ? 1 : 0);
    switch (D) {
#define X(n)                                                                                                     \
    case n:                                                                                                      \
        if (launch_fwd<n>(a, render_colors, render_alphas, last_ids, acc_depth, as_stream(stream)) != 0) {       \
            set_error("d4_blend_fwd: kernel configuration failed");                                              \
            return 1;                                                                                            \
        }                                                                                                        \
        break;
        D4_FOR_EACH_D(X)
#undef X
        default:
            set_error("d4_blend_fwd: channel count D=%d not built (pad to one of 1-9,16,17,32,33)", D);
            return 2;
    }
    D4_CHECK_LAUNCH("d4_blend_fwd");
    return 0;
}

extern "C" int d4_blend_bwd(const float *means2d, const float *conics, const float *opacities,
                            const float *colors, int64_t colors_cam_stride, const float *depths,
                            const float *backgrounds, int C, int G, int D0, int width, int height,
                            int tile_size, int tile_w, int tile_h, const int32_t *tile_offsets,
                            const int32_t *flatten_ids, int64_t n_isects, int normalize_depth,
                            const float *render_alphas, const int32_t *last_ids, const float *acc_depth,
                            const float *v_render_colors, const float *v_render_alphas, float *v_means2d,
                            float *v_conics, float *v_colors, float *v_opacities, float *v_depths,
                            const uint8_t *hit_masks, int bwd_mode, d4_stream_t stream) {
    BlendArgs a{means2d, conics, opacities, colors, depths, backgrounds, colors_cam_stride, C, G, D0, width,
                height, tile_w, tile_h, tile_offsets, flatten_ids, n_isects, normalize_depth,
                const_cast<uint8_t *>(hit_masks)};
    D4_CHECK_ARG(bwd_mode == 0 || bwd_mode == 1, "d4_blend_bwd: bwd_mode is 0 (grouped) or 1 (shuffle)");
    if (int rc = check_blend_args("d4_blend_bwd", a, tile_size)) return rc;
    D4_CHECK_ARG(render_alphas && last_ids && v_render_colors && v_render_alphas && v_means2d && v_conics &&
                     v_opacities && (v_colors || D0 == 0),
                 "d4_blend_bwd: null pointer");
    D4_CHECK_ARG(!depths || v_depths, "d4_blend_bwd: v_depths required when depths is given");
    D4_CHECK_ARG(!normalize_depth || (depths && acc_depth), "d4_blend_bwd: normalize_depth needs depths and acc_depth");
    if (n_isects == 0) return 0;
    const int D = D0 + (depths ? 1 : 0);
    switch (D) {
#define X(n)                                                                                                   \
    case n:                                                                                                    \
        if (launch_bwd<n>(bwd_mode, a, render_alphas, last_ids, acc_depth, v_render_colors, v_render_alphas,   \
                          v_means2d, v_conics, v_colors, v_opacities, v_depths, as_stream(stream)) != 0) {     \
            set_error("d4_blend_bwd: kernel configuration failed");                                            \
            return 1;                                                                                          \
        }                                                                                                      \
        break;
        D4_FOR_EACH_D(X)
#undef X
        default:
            set_error("d4_blend_bwd: channel count D=%d not built (pad to one of 1-9,16,17,32,33)", D);
            return 2;
    }
    D4_CHECK_LAUNCH("d4_blend_bwd");
    return 0;
}
