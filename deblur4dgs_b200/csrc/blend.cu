// blend.cu -- rows a10 / a11 of SURVEY.md section 8: per-tile alpha compositing.
//   d4_blend_fwd : gsplat rasterize_to_pixels fwd (+ depth channel, + ED normalisation)
//   d4_blend_bwd : gsplat rasterize_to_pixels bwd (+ ED normalisation backward)
//
// One CTA per (camera, 16x16 tile), 256 threads = 256 pixels; each WARP owns a
// compact 8x4 pixel block (not a 16x2 strip) so that a Gaussian's footprint
// touches as few warps as possible and whole warps skip it after one vote.
// Gaussians of the tile are staged 256 at a time into shared memory
// (xy+opacity, conic, D colour channels, id); the per-pair loop reads them as
// warp-wide broadcasts.
//
// Backward, per (warp, Gaussian): the D+6 partial sums of the 32 pixels are
// reduced with a TRANSPOSING butterfly (31 shuffles for up to 32 values instead
// of 5 per value), added to a per-CTA shared-memory accumulator and flushed to
// HBM once per (tile, Gaussian) -- one global atomic per value per tile instead
// of one per warp.  The per-pixel recurrences are carried as scalars:
//   s_i = <c_i, v_out>,  S = sum_{j>i} s_j alpha_j T_j,
//   dL/dalpha_i = T_i s_i - (S + T_final (bg.v_out - v_alpha_out)) / (1 - alpha_i)
// which is algebraically gsplat's per-channel buffer[] form with D fewer
// registers and D fewer FMAs per pair.
//
// These kernels are bound by fp32 issue + MUFU.EX2 + shuffle throughput, not by
// HBM: algorithmic bytes per (pixel, Gaussian) pair are ~0.4.
#include "blend_common.cuh"

namespace d4 {

// ----------------------------------------------------------------------------- forward
template <int D>
__global__ void __launch_bounds__(kBlendThreads)
blend_fwd_kernel(BlendArgs a, float *__restrict__ render_colors, float *__restrict__ render_alphas,
                 int32_t *__restrict__ last_ids, float *__restrict__ acc_depth) {
    constexpr int DS = BlendCfg<D>::DS;
    constexpr int DP = BlendCfg<D>::DP;
    __shared__ float4 s_geom[kBatch];
    __shared__ float4 s_conic[kBatch];
    __shared__ __align__(16) float s_col[kBatch * (DS > DP ? DS : DP)];
    __shared__ uint32_t s_mask[kBatch];

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

    float T = 1.0f;
    int32_t cur_idx = 0;
    bool done = !inside;
    // accumulators as fp32x2 pairs: Blackwell's FFMA2 retires two fp32 FMAs per issue slot, and this
    // kernel is bound by issue slots
    constexpr int D2 = (D + 1) / 2;
    float2 out2[D2];
#pragma unroll
    for (int k = 0; k < D2; ++k) out2[k] = make_float2(0.f, 0.f);

    for (int b = 0; b < num_batches; ++b) {
        if (__syncthreads_count(done) >= kBlendThreads) break;
        const int64_t batch_start = range_start + (int64_t)kBatch * b;
        stage_gaussian<D>(a, c, batch_start + tid, batch_start + tid < range_end, tid, tx * kTile, ty * kTile, s_geom,
                          s_conic, s_col, s_mask);
        __syncthreads();
        const int batch_size = (int)min((int64_t)kBatch, range_end - batch_start);
        bool warp_done = __all_sync(0xffffffffu, done);
        // one Gaussian of the hit list for this pixel; returns true when the whole warp is finished
        auto composite = [&](int t, float power, float L) -> bool {
            const float alpha = fminf(kAlphaMax, ex2_approx(power));
            const bool valid = !done && power <= L && alpha >= kAlphaMin;
            if (!__any_sync(0xffffffffu, valid)) return false;
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

// ---- tensor-core path for the colour gradients ----------------------------------------------
// v_colors[g][ch] = sum_pixels fac[g][p] * v_out[p][ch] is a genuine dense contraction per warp:
// [16 Gaussians x 32 pixels] x [32 pixels x 8*NT channels].  It runs on the tensor cores as
// m16n8k8 TF32 MMAs with BOTH operands split into hi + lo TF32 parts (3 MMAs per product:
// hi*hi + lo*hi + hi*lo, fp32 accumulate), which keeps ~2^-21 relative accuracy -- plain TF32
// (2^-11) would break the 1e-4 parity.  tcgen05 is not applicable: its operands are CTA-wide
// shared-memory tiles issued by one thread, these are per-warp 16x32 fragments.
__device__ __forceinline__ void tf32_split(float x, uint32_t &hi, uint32_t &lo) {
    // hi = x truncated to TF32 (one LOP3; cvt.rna.tf32 costs ~5 SASS instructions on sm_100), lo = x - hi
    // exactly.  |lo| <= 2^-10 |x| and is itself truncated by the MMA: total error ~2^-20 per product.
    hi = __float_as_uint(x) & 0xffffe000u;
    lo = __float_as_uint(x - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

constexpr int kFacRows = 16;    // Gaussians buffered per warp before one MMA flush
constexpr int kFacStride = 36;  // 32 pixels + 4 pad: conflict-free A-fragment loads

// Backward shared-memory budget at D = 17, 128 Gaussians per batch: geometry 4 KB + colours 10 KB +
// accumulators 11.75 KB + slot ids 1 KB + MMA buffers (v_out 16 KB swizzled, fac 18 KB) = 61 KB
// -> 3 CTAs (24 warps) per SM.  Measured alternatives (profiles/): 256-Gaussian batches fit only
// 2 CTAs/SM and are latency-bound (issue slots 54 % busy); reading the colours from global memory
// instead of staging them frees 20 KB but puts an L1 round trip on the T/S recurrence (slower).
//
// MEASURED at c3 (D = 17, 9 x 1.49 M intersections): the tensor-core path executes 41 % fewer
// instructions than the all-shuffle reduction (6.2e9 vs 7.9e9 warp instructions) but is NOT faster:
// 9.2-10.4 ms vs 8.8 ms.  Its shared-memory footprint caps occupancy, the per-warp flush adds
// LDS / split / CAS-atomic latency, and the kernel is then bound by barrier imbalance and
// short-scoreboard stalls rather than by issue slots.  It is therefore compiled out by default
// (kUseMmaBwd = false) and kept for the next round (needs operands kept in registers).
constexpr bool kUseMmaBwd = false;
constexpr int kBatchB = kUseMmaBwd ? 128 : 256;

template <int D>
struct BwdCfg {
    static constexpr int NT = kUseMmaBwd ? D / 8 : 0;  // n-tiles of 8 channels handled by the MMA path
    static constexpr int DM = NT * 8;                // channels [0, DM) -> tensor cores
    static constexpr int VSH = D - DM + 6;           // values still reduced with shuffles
    static constexpr int RW = RedWidth<VSH>::value;
    // s_vout row stride: DM == 16 uses an XOR swizzle (stride 16, conflict-free); otherwise a padded
    // stride == 8 or 24 (mod 32) so that the 4 x 8 B-fragment addresses of a warp hit 32 banks
    static constexpr bool SWZ = (DM == 16);
    static constexpr int VOS = NT > 0 ? (SWZ ? 16 : ((DM % 16 == 8) ? DM + 16 : DM + 8)) : 0;
    static constexpr size_t smem_bytes() {
        return sizeof(float4) * 2 * kBatchB + sizeof(float) * kBatchB * (BlendCfg<D>::DS + (BlendCfg<D>::V | 1)) +
               (NT > 0 ? sizeof(float) * (kBlendThreads / 32) * (32 * VOS + kFacRows * kFacStride) : 0);
    }
    static __device__ __forceinline__ int vout_idx(int px, int ch) {
        return SWZ ? px * 16 + (ch ^ (((px >> 1) & 1) << 3)) : px * VOS + ch;
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
    constexpr int NT = BwdCfg<D>::NT, DM = BwdCfg<D>::DM, VSH = BwdCfg<D>::VSH, RW = BwdCfg<D>::RW;
    constexpr int VOS = BwdCfg<D>::VOS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4 *s_geom = reinterpret_cast<float4 *>(smem_raw);
    float4 *s_conic = s_geom + kBatchB;
    float *s_col = reinterpret_cast<float *>(s_conic + kBatchB);
    float *s_acc = s_col + kBatchB * DS;
    float *s_vout_all = s_acc + kBatchB * VS;                               // [8 warps][32 px][VOS]
    float *s_fac_all = s_vout_all + (kBlendThreads / 32) * 32 * VOS;       // [8 warps][16][36]
    __shared__ int32_t s_max[kBlendThreads / 32];
    __shared__ uint32_t s_mask[kBatchB];
    __shared__ int32_t s_slot[kBlendThreads / 32][kFacRows];
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
    float *s_vout = s_vout_all + w * 32 * VOS;
    float *s_fac = s_fac_all + w * kFacRows * kFacStride;
    if constexpr (NT > 0) {
        // B operand of the MMA path: this warp's 32 x DM block of v_out (zero-padded to VOS columns)
#pragma unroll
        for (int k = 0; k < DM; ++k) s_vout[BwdCfg<D>::vout_idx(lane, k)] = v_out[k];
    }
    __syncthreads();
    int32_t block_bin_final = s_max[0];
#pragma unroll
    for (int k = 1; k < kBlendThreads / 32; ++k) block_bin_final = max(block_bin_final, s_max[k]);
    range_end = min(range_end, (int64_t)block_bin_final + 1);
    if (range_end <= range_start) return;
    const int num_batches = (int)((range_end - range_start + kBatchB - 1) / kBatchB);

    const int gid = lane >> 2, tig = lane & 3;
    int nb = 0;  // Gaussians buffered in s_fac (warp-uniform)

    // one MMA flush: s_acc[slot[row]][ch] += sum_p s_fac[row][p] * s_vout[p][ch], rows < nb
    auto flush_colors = [&]() {
        if constexpr (NT > 0) {
            __syncwarp();
            float acc[NT][4];
#pragma unroll
            for (int n = 0; n < NT; ++n)
#pragma unroll
                for (int q = 0; q < 4; ++q) acc[n][q] = 0.f;
            const bool r0ok = gid < nb, r1ok = gid + 8 < nb;
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                const int c0 = ks * 8 + tig;
                float af[4];
                af[0] = r0ok ? s_fac[gid * kFacStride + c0] : 0.f;
                af[1] = r1ok ? s_fac[(gid + 8) * kFacStride + c0] : 0.f;
                af[2] = r0ok ? s_fac[gid * kFacStride + c0 + 4] : 0.f;
                af[3] = r1ok ? s_fac[(gid + 8) * kFacStride + c0 + 4] : 0.f;
                uint32_t ah[4], al[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) tf32_split(af[q], ah[q], al[q]);
#pragma unroll
                for (int n = 0; n < NT; ++n) {
                    uint32_t bh0, bl0, bh1, bl1;
                    tf32_split(s_vout[BwdCfg<D>::vout_idx(c0, n * 8 + gid)], bh0, bl0);
                    tf32_split(s_vout[BwdCfg<D>::vout_idx(c0 + 4, n * 8 + gid)], bh1, bl1);
                    mma_tf32(acc[n], ah, bh0, bh1);
                    mma_tf32(acc[n], al, bh0, bh1);
                    mma_tf32(acc[n], ah, bl0, bl1);
                }
            }
            const int slot0 = r0ok ? s_slot[w][gid] : 0, slot1 = r1ok ? s_slot[w][gid + 8] : 0;
#pragma unroll
            for (int n = 0; n < NT; ++n) {
                const int ch = n * 8 + tig * 2;
                if (r0ok) {
                    if (acc[n][0] != 0.f) atomicAdd(&s_acc[slot0 * VS + ch], acc[n][0]);
                    if (acc[n][1] != 0.f) atomicAdd(&s_acc[slot0 * VS + ch + 1], acc[n][1]);
                }
                if (r1ok) {
                    if (acc[n][2] != 0.f) atomicAdd(&s_acc[slot1 * VS + ch], acc[n][2]);
                    if (acc[n][3] != 0.f) atomicAdd(&s_acc[slot1 * VS + ch + 1], acc[n][3]);
                }
            }
            __syncwarp();
            nb = 0;
        }
    };

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
                    if constexpr (DM == 0) {
                        const float2 fac2 = make_float2(fac, fac);
#pragma unroll
                        for (int k2 = 0; k2 < D2; ++k2) {
                            const float2 p = __fmul2_rn(fac2, v2[k2]);
                            r[2 * k2] = p.x;
                            if (2 * k2 + 1 < D) r[2 * k2 + 1] = p.y;
                        }
                    } else {
#pragma unroll
                        for (int k = DM; k < D; ++k) r[k - DM] = fac * v_out[k];  // channels not on the MMA path
                    }
                    const float v_alpha = s * T - (S - tail) * ra;
                    S = fmaf(s, fac, S);
                    if (araw <= kAlphaMax) {
                        // v_sigma = -araw * v_alpha; conic terms in the primed (base-2) scale:
                        //   a dx + b dy = -(2 A' dx + B' dy) / log2e
                        const float v_sigma = -araw * v_alpha;
                        const float vs2 = v_sigma * (-1.0f / kLog2e);
                        r[D - DM + 0] = 0.5f * v_sigma * dx * dx;
                        r[D - DM + 1] = v_sigma * dx * dy;
                        r[D - DM + 2] = 0.5f * v_sigma * dy * dy;
                        r[D - DM + 3] = vs2 * fmaf(2.0f * cn.x, dx, cn.y * dy);
                        r[D - DM + 4] = vs2 * fmaf(cn.y, dx, 2.0f * cn.z * dy);
                        r[D - DM + 5] = araw * cn.w * v_alpha;  // exp(-sigma) * v_alpha
                    }
                }
                if constexpr (NT > 0) {
                    // buffer this Gaussian's per-pixel weights for the tensor-core flush
                    s_fac[nb * kFacStride + lane] = fac;
                    if (lane == 0) s_slot[w][nb] = t;
                    ++nb;
                }
                // shuffle path: the remaining VSH values; lane j < VSH ends up owning value j
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
                        const int idx2 = j < D - DM ? DM + j : D + (j - (D - DM));
                        atomicAdd(&s_acc[t * VS + idx2], rb8[0]);
                    }
                } else {
                    warp_transpose_reduce<RW>(r, lane);
                    mine = r[0];
                }
                if (lane < (VSH < 32 ? VSH : 32)) {
                    const int idx = lane < D - DM ? DM + lane : D + (lane - (D - DM));
                    atomicAdd(&s_acc[t * VS + idx], mine);
                }
                if constexpr (NT > 0) {
                    if (nb == kFacRows) flush_colors();
                }
            }
        }
        if (nb > 0) flush_colors();
        __syncthreads();  // barrier C: every warp is done with batch b
    }
    flush_acc(prev_size, s_gid[(num_batches & 1) ^ 1]);
}

// ----------------------------------------------------------------------------- backward, tensor-core v2
// D = 16 / 17 (the Deblur4DGS dynamic pass).  Both per-warp contractions over the colour channels run on the
// tensor cores with the CONSTANT operand (this warp's 32 x 16 block of v_out) held in registers in the
// two fragment layouts they need, so no shared-memory operand buffer and no per-lane v_out[] array:
//   phase A  S[px x g]   = Vout[px x ch] * C^T[ch x g]      (the <c_g, v_out> dots of 16 Gaussians at once)
//   phase B  the sequential per-pixel recurrence over those 16 Gaussians (alpha, T, S, geometry terms),
//            reading s from / writing fac = alpha*T into one 16 x 32 shared-memory tile per warp
//   phase C  Vc^T[ch x g] = Vout^T[ch x px] * Fac^T[px x g]  (colour gradients of the 16 Gaussians)
// m16n8k8 TF32 MMAs, both operands split hi + lo (3 MMAs per product, ~2^-21 relative accuracy).
constexpr int kSfStride = 36;  // 32 pixels + 4: conflict-free for all fragment accesses below
// Default of the D4_BWD_TC switch.  MEASURED at c3 (profiles/r01c_bwd_tc.md): 6.6-7.0e9 warp instructions
// instead of 7.6e9, contributing path 108 instead of ~195 instructions, parity green -- but 9.7-10.5 ms vs
// 8.3 ms for the shuffle kernel: the 32 operand registers limit it to 3 CTAs/SM (80 regs, some spills), every
// group of 16 Gaussians pays ~350-480 instructions of split / fragment / atomic overhead, and issue slots are
// only 61 % busy (barrier + short-scoreboard stalls).  Off by default; D4_BWD_TC=1 enables it for A/B runs.
constexpr int kDefaultTcBwd = 0;
constexpr int kDefaultBwdMode = 1;  // 1 = grouped backward (blend_bwd_gp.cu), 0 = shuffle kernel of this file
constexpr int kDefaultGpCfg = 0;

template <int D>
struct TcCfg {
    static constexpr int V = BlendCfg<D>::V, VS = V | 1, DS = BlendCfg<D>::DS;
    static constexpr size_t smem_bytes() {
        return sizeof(float4) * 2 * kBatchB + sizeof(float) * kBatchB * (DS + VS) +
               sizeof(float) * (kBlendThreads / 32) * 16 * kSfStride;
    }
};

__device__ __forceinline__ void split4(const float (&x)[4], uint32_t (&hi)[4], uint32_t (&lo)[4]) {
#pragma unroll
    for (int q = 0; q < 4; ++q) tf32_split(x[q], hi[q], lo[q]);
}

template <int D>
__global__ void __launch_bounds__(kBlendThreads, 3)
blend_bwd_tc_kernel(BlendArgs a, const float *__restrict__ render_alphas, const int32_t *__restrict__ last_ids,
                    const float *__restrict__ acc_depth, const float *__restrict__ v_render_colors,
                    const float *__restrict__ v_render_alphas, float *__restrict__ v_means2d,
                    float *__restrict__ v_conics, float *__restrict__ v_colors, float *__restrict__ v_opacities,
                    float *__restrict__ v_depths) {
    static_assert(D == 16 || D == 17, "tensor-core backward is built for 16 colour channels (+ depth)");
    constexpr int DS = TcCfg<D>::DS, V = TcCfg<D>::V, VS = TcCfg<D>::VS;
    constexpr int VSH = D - 16 + 6;  // shuffle-reduced values: [depth channel,] 3 conic, 2 xy, 1 opacity
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4 *s_geom = reinterpret_cast<float4 *>(smem_raw);
    float4 *s_conic = s_geom + kBatchB;
    float *s_col = reinterpret_cast<float *>(s_conic + kBatchB);
    float *s_acc = s_col + kBatchB * DS;
    float *s_sf_all = s_acc + kBatchB * VS;
    __shared__ int32_t s_max[kBlendThreads / 32];
    __shared__ uint32_t s_mask[kBatchB];
    __shared__ int32_t s_gid[2][kBatchB];

    const int n_tiles = a.tile_w * a.tile_h;
    const int ct = blockIdx.x;
    const int c = ct / n_tiles;
    const int tile = ct - c * n_tiles;
    const int ty = tile / a.tile_w, tx = tile - ty * a.tile_w;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int gid = lane >> 2, tig = lane & 3;
    int lx, ly;
    pixel_of_thread(tid, lx, ly);
    const int j = tx * kTile + lx, i = ty * kTile + ly;
    const bool inside = (i < a.height) && (j < a.width);
    const float px = (float)j + 0.5f, py = (float)i + 0.5f;
    const int64_t pid = ((int64_t)c * a.height + i) * a.width + j;

    const int64_t range_start = a.tile_offsets[ct];
    int64_t range_end = (ct == a.C * n_tiles - 1) ? a.n_isects : (int64_t)a.tile_offsets[ct + 1];
    if (range_end <= range_start) return;

    float *s_sf = s_sf_all + w * 16 * kSfStride;
    float T_final = 1.f, v_ra = 0.f, v_depth = 0.f, bgdot = 0.f;
    int32_t bin_final = -1;
    float va[2][2][4], vc[4][4];
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
        if constexpr (D == 17) v_depth = v_out[16];
        // transpose this warp's 32 x 16 block of v_out through shared memory into the two fragment layouts
#pragma unroll
        for (int k = 0; k < 16; ++k) s_sf[lane * 17 + k] = v_out[k];
        __syncwarp();
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
                va[mt][ks][0] = s_sf[(mt * 16 + gid) * 17 + ks * 8 + tig];
                va[mt][ks][1] = s_sf[(mt * 16 + gid + 8) * 17 + ks * 8 + tig];
                va[mt][ks][2] = s_sf[(mt * 16 + gid) * 17 + ks * 8 + tig + 4];
                va[mt][ks][3] = s_sf[(mt * 16 + gid + 8) * 17 + ks * 8 + tig + 4];
            }
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            vc[ks][0] = s_sf[(ks * 8 + tig) * 17 + gid];
            vc[ks][1] = s_sf[(ks * 8 + tig) * 17 + gid + 8];
            vc[ks][2] = s_sf[(ks * 8 + tig + 4) * 17 + gid];
            vc[ks][3] = s_sf[(ks * 8 + tig + 4) * 17 + gid + 8];
        }
        __syncwarp();
    }
    const float tail = T_final * (v_ra - bgdot);
    float T = T_final;
    float S = 0.f;

    const int32_t warp_bin_final = __reduce_max_sync(0xffffffffu, bin_final);
    if (lane == 0) s_max[w] = warp_bin_final;
    __syncthreads();
    int32_t block_bin_final = s_max[0];
#pragma unroll
    for (int k = 1; k < kBlendThreads / 32; ++k) block_bin_final = max(block_bin_final, s_max[k]);
    range_end = min(range_end, (int64_t)block_bin_final + 1);
    if (range_end <= range_start) return;
    const int num_batches = (int)((range_end - range_start + kBatchB - 1) / kBatchB);

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
        const int64_t batch_end = range_end - 1 - (int64_t)kBatchB * b;
        const int batch_size = (int)min((int64_t)kBatchB, batch_end + 1 - range_start);
        if (tid < kBatchB) {
            const bool in_range = batch_end - tid >= range_start;
            stage_gaussian<D>(a, c, batch_end - tid, in_range, tid, tx * kTile, ty * kTile, s_geom, s_conic, s_col,
                              s_mask);
            s_gid[b & 1][tid] = in_range ? __ldg(a.flatten_ids + (batch_end - tid)) : 0;
        }
        flush_acc(prev_size, s_gid[(b & 1) ^ 1]);
        prev_size = batch_size;
        __syncthreads();  // barrier B

        const int t0 = (int)max((int64_t)0, batch_end - (int64_t)warp_bin_final);
        for (int sc = t0 >> 7; sc * 128 < batch_size; ++sc) {
            // hit masks of the four 32-slot chunks of this 128-slot super-chunk
            uint32_t m[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int base = sc * 128 + q * 32;
                uint32_t bits = 0u;
                if (base < batch_size) bits = __ballot_sync(0xffffffffu, (s_mask[base + lane] >> w) & 1u);
                if (base + 31 < t0) bits = 0u;
                else if (base < t0) bits &= ~((1u << (t0 - base)) - 1u);
                m[q] = bits;
            }
            const int c1 = __popc(m[0]), c2 = c1 + __popc(m[1]), c3 = c2 + __popc(m[2]), total = c3 + __popc(m[3]);
            for (int grp = 0; grp * 16 < total; ++grp) {
                const int ng = min(16, total - grp * 16);
                // lane jj < ng owns the slot of the (grp*16 + jj)-th hit
                int tl = 0;
                {
                    const int r = grp * 16 + (lane & 15);
                    if (r < total) {
                        int q, rr;
                        if (r < c1) { q = 0; rr = r; }
                        else if (r < c2) { q = 1; rr = r - c1; }
                        else if (r < c3) { q = 2; rr = r - c2; }
                        else { q = 3; rr = r - c3; }
                        const uint32_t mq = q == 0 ? m[0] : (q == 1 ? m[1] : (q == 2 ? m[2] : m[3]));
                        tl = sc * 128 + q * 32 + (int)__fns(mq, 0, rr + 1);
                    }
                }
                // ---------------- phase A: S[px][g] = sum_ch v_out[px][ch] * color[g][ch] -----------------
#pragma unroll
                for (int mt = 0; mt < 2; ++mt) {
                    uint32_t ah[2][4], al[2][4];
                    split4(va[mt][0], ah[0], al[0]);
                    split4(va[mt][1], ah[1], al[1]);
#pragma unroll
                    for (int nt = 0; nt < 2; ++nt) {
                        const int tg = __shfl_sync(0xffffffffu, tl, nt * 8 + gid);
                        const float *cp = s_col + tg * DS + tig;
                        float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                        for (int ks = 0; ks < 2; ++ks) {
                            uint32_t bh0, bl0, bh1, bl1;
                            tf32_split(cp[ks * 8], bh0, bl0);
                            tf32_split(cp[ks * 8 + 4], bh1, bl1);
                            mma_tf32(acc, ah[ks], bh0, bh1);
                            mma_tf32(acc, al[ks], bh0, bh1);
                            mma_tf32(acc, ah[ks], bl0, bl1);
                        }
                        float *dst = s_sf + (nt * 8 + 2 * tig) * kSfStride + mt * 16 + gid;
                        dst[0] = acc[0];
                        dst[kSfStride] = acc[1];
                        dst[8] = acc[2];
                        dst[kSfStride + 8] = acc[3];
                    }
                }
                __syncwarp();
                // ---------------- phase B: per-pixel recurrence over the group's Gaussians -----------------
                for (int gi = 0; gi < ng; ++gi) {
                    const int t = __shfl_sync(0xffffffffu, tl, gi);
                    const float4 g0 = s_geom[t];
                    const float4 cn = s_conic[t];
                    const float dx = g0.x - px, dy = g0.y - py;
                    const float power = fmaf(cn.z * dy, dy, fmaf(fmaf(cn.y, dy, cn.x * dx), dx, g0.z));
                    const float araw = ex2_approx(power);
                    const float alpha = fminf(kAlphaMax, araw);
                    const bool valid = (batch_end - t <= (int64_t)bin_final) && power <= g0.z && alpha >= kAlphaMin;
                    float *sf = s_sf + gi * kSfStride + lane;
                    if (!__any_sync(0xffffffffu, valid)) {
                        *sf = 0.f;
                        continue;
                    }
                    float r[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k) r[k] = 0.f;
                    float fac = 0.f;
                    if (valid) {
                        float s = *sf;
                        if constexpr (D == 17) s = fmaf(s_col[t * DS + 16], v_depth, s);
                        const float ra = __fdividef(1.0f, 1.0f - alpha);
                        T *= ra;
                        fac = alpha * T;
                        const float v_alpha = s * T - (S - tail) * ra;
                        S = fmaf(s, fac, S);
                        if constexpr (D == 17) r[0] = fac * v_depth;
                        if (araw <= kAlphaMax) {
                            const float v_sigma = -araw * v_alpha;
                            const float vs2 = v_sigma * (-1.0f / kLog2e);
                            r[D - 16 + 0] = 0.5f * v_sigma * dx * dx;
                            r[D - 16 + 1] = v_sigma * dx * dy;
                            r[D - 16 + 2] = 0.5f * v_sigma * dy * dy;
                            r[D - 16 + 3] = vs2 * fmaf(2.0f * cn.x, dx, cn.y * dy);
                            r[D - 16 + 4] = vs2 * fmaf(cn.y, dx, 2.0f * cn.z * dy);
                            r[D - 16 + 5] = araw * cn.w * v_alpha;
                        }
                    }
                    *sf = fac;
                    warp_transpose_reduce<8>(r, lane);
                    if (lane < VSH) atomicAdd(&s_acc[t * VS + 16 + lane], r[0]);  // values 16.. are contiguous in s_acc
                }
                __syncwarp();
                // ---------------- phase C: Vc[ch][g] = sum_px v_out[px][ch] * fac[g][px] --------------------
                {
                    float acc[2][4];
#pragma unroll
                    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                        for (int q = 0; q < 4; ++q) acc[nt][q] = 0.f;
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        uint32_t ah[4], al[4];
                        split4(vc[ks], ah, al);
#pragma unroll
                        for (int nt = 0; nt < 2; ++nt) {
                            const float *fp = s_sf + (nt * 8 + gid) * kSfStride + ks * 8 + tig;
                            uint32_t bh0, bl0, bh1, bl1;
                            tf32_split(fp[0], bh0, bl0);
                            tf32_split(fp[4], bh1, bl1);
                            mma_tf32(acc[nt], ah, bh0, bh1);
                            mma_tf32(acc[nt], al, bh0, bh1);
                            mma_tf32(acc[nt], ah, bl0, bl1);
                        }
                    }
#pragma unroll
                    for (int nt = 0; nt < 2; ++nt) {
                        const int gA = nt * 8 + 2 * tig;
                        const int tA = __shfl_sync(0xffffffffu, tl, gA), tB = __shfl_sync(0xffffffffu, tl, gA + 1);
                        if (gA < ng) {
                            if (acc[nt][0] != 0.f) atomicAdd(&s_acc[tA * VS + gid], acc[nt][0]);
                            if (acc[nt][2] != 0.f) atomicAdd(&s_acc[tA * VS + gid + 8], acc[nt][2]);
                        }
                        if (gA + 1 < ng) {
                            if (acc[nt][1] != 0.f) atomicAdd(&s_acc[tB * VS + gid], acc[nt][1]);
                            if (acc[nt][3] != 0.f) atomicAdd(&s_acc[tB * VS + gid + 8], acc[nt][3]);
                        }
                    }
                }
                __syncwarp();
            }
        }
        __syncthreads();  // barrier C
    }
    flush_acc(prev_size, s_gid[(num_batches & 1) ^ 1]);
}

// ----------------------------------------------------------------------------- dispatch
template <int D>
static int launch_fwd(const BlendArgs &a, float *rc, float *ra, int32_t *li, float *ad, cudaStream_t st) {
    int grid = a.C * a.tile_w * a.tile_h;
    blend_fwd_kernel<D><<<grid, kBlendThreads, 0, st>>>(a, rc, ra, li, ad);
    return 0;
}
static bool use_tc_bwd() {
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("D4_BWD_TC");
        v = e ? atoi(e) : kDefaultTcBwd;
    }
    return v != 0;
}
// D4_BWD selects the backward formulation: "gp" = grouped (blend_bwd_gp.cu), "shfl" = warp-butterfly kernel below.
// Read on every call so that one process can A/B the two (tests, scripts/ab_blend_bwd.py).
static int bwd_mode() {
    const char *e = getenv("D4_BWD");
    if (!e) return kDefaultBwdMode;
    return (e[0] == 's') ? 0 : 1;
}
static int bwd_gp_cfg() {
    const char *e = getenv("D4_BWD_GP_CFG");
    return e ? atoi(e) : kDefaultGpCfg;
}

template <int D>
static int launch_bwd(const BlendArgs &a, const float *ra, const int32_t *li, const float *ad, const float *vrc,
                      const float *vra, float *vm, float *vc, float *vcol, float *vo, float *vd, cudaStream_t st) {
    int grid = a.C * a.tile_w * a.tile_h;
    if (bwd_mode() == 1 && !use_tc_bwd()) {
        const int rc = launch_blend_bwd_gp(D, bwd_gp_cfg(), a, ra, li, ad, vrc, vra, vm, vc, vcol, vo, vd, st);
        if (rc >= 0) return rc;
    }
    if constexpr (D == 16 || D == 17) {
        if (use_tc_bwd()) {
            constexpr size_t smem_tc = TcCfg<D>::smem_bytes();
            static bool configured_tc = false;
            if (!configured_tc) {
                if (cudaFuncSetAttribute(blend_bwd_tc_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem_tc) != cudaSuccess)
                    return 1;
                configured_tc = true;
            }
            blend_bwd_tc_kernel<D><<<grid, kBlendThreads, smem_tc, st>>>(a, ra, li, ad, vrc, vra, vm, vc, vcol, vo, vd);
            return 0;
        }
    }
    constexpr size_t smem = BwdCfg<D>::smem_bytes();
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(blend_bwd_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return 1;
        configured = true;
    }
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
                            d4_stream_t stream) {
    BlendArgs a{means2d, conics, opacities, colors, depths, backgrounds, colors_cam_stride, C, G, D0, width,
                height, tile_w, tile_h, tile_offsets, flatten_ids, n_isects, normalize_depth};
    if (int rc = check_blend_args("d4_blend_fwd", a, tile_size)) return rc;
    D4_CHECK_ARG(render_colors && render_alphas && last_ids, "d4_blend_fwd: null output");
    D4_CHECK_ARG(!normalize_depth || (depths && acc_depth), "d4_blend_fwd: normalize_depth needs depths and acc_depth");
    const int D = D0 + (depths ? 1 : 0);
    switch (D) {
#define X(n) case n: launch_fwd<n>(a, render_colors, render_alphas, last_ids, acc_depth, as_stream(stream)); break;
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
                            d4_stream_t stream) {
    BlendArgs a{means2d, conics, opacities, colors, depths, backgrounds, colors_cam_stride, C, G, D0, width,
                height, tile_w, tile_h, tile_offsets, flatten_ids, n_isects, normalize_depth};
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
        launch_bwd<n>(a, render_alphas, last_ids, acc_depth, v_render_colors, v_render_alphas, v_means2d,      \
                      v_conics, v_colors, v_opacities, v_depths, as_stream(stream));                           \
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
