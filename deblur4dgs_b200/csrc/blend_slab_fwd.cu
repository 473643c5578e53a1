// blend_slab_fwd.cu -- row a10 of SURVEY.md section 8: forward alpha compositing (gsplat rasterize_to_pixels fwd +
// depth channel + ED normalisation) over the packed per-tile record slabs of slab.cuh.
//
// One CTA per (camera, 16x16 tile): eight consumer warps (one compact 8x4 pixel block each) and one producer warp.
// The producer streams the tile's records through a ring of kStages shared-memory stages (cp.async.bulk for the
// records, 16-byte cp.async for the colour rows, both completing on the stage's "full" mbarrier); the consumers never
// meet at a CTA barrier inside the main loop -- a warp whose pixels are all saturated just keeps releasing stages
// until the producer sees that all eight are done and stops streaming.
//
// Per pair the arithmetic is that of blend.cu (base-2 exponent from pre-scaled conics: 5 FMA-pipe ops + MUFU.EX2,
// packed FFMA2 colour accumulation), so the two forward paths agree to the last bit of every decision.
#include "slab.cuh"

namespace d4 {

template <int D0, bool DEPTH>
struct FwdSlabCfg {
    static constexpr int D = D0 + (DEPTH ? 1 : 0);
    static constexpr int DP = D | 1;  // odd stride of the epilogue transpose buffer
    static constexpr int kStages = D0 <= 8 ? 8 : (D0 <= 16 ? 6 : 3);
    __host__ __device__ static constexpr size_t ring_bytes() { return (size_t)kStages * kSlabChunk * (32 + 4 * D0); }
    __host__ __device__ static constexpr size_t epi_bytes() { return sizeof(float) * kBlendThreads * DP; }
    static constexpr size_t smem_bytes() {
        return (ring_bytes() > epi_bytes() ? ring_bytes() : epi_bytes()) + 2 * kStages * sizeof(uint64_t) + 16;
    }
};

template <int D0, bool DEPTH, bool kMasks>
__global__ void __launch_bounds__(kSlabThreads, (D0 <= 16 ? 4 : 2))
blend_fwd_slab_kernel(SlabArgs a, float *__restrict__ render_colors, float *__restrict__ render_alphas,
                      int32_t *__restrict__ last_ids, float *__restrict__ acc_depth) {
    using Cfg = FwdSlabCfg<D0, DEPTH>;
    constexpr int D = Cfg::D, DP = Cfg::DP, S = Cfg::kStages, CH = kSlabChunk;
    constexpr int D2 = D0 / 2;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float4 *s_rec = reinterpret_cast<float4 *>(smem_raw);                  // [S][CH][2]
    float *s_col = reinterpret_cast<float *>(s_rec + S * CH * 2);          // [S][CH][D0]
    constexpr size_t data_bytes = Cfg::ring_bytes() > Cfg::epi_bytes() ? Cfg::ring_bytes() : Cfg::epi_bytes();
    uint64_t *s_full = reinterpret_cast<uint64_t *>(smem_raw + ((data_bytes + 15) & ~(size_t)15));  // [S]
    uint64_t *s_empty = s_full + S;                                                                 // [S]
    __shared__ int s_ndone;  // consumer warps with all pixels saturated (or outside the image)
    const uint32_t s_rec_addr = smem_u32(s_rec), s_col_addr = smem_u32(s_col);

    const int n_tiles = a.tile_w * a.tile_h;
    const int ct = blockIdx.x;
    const int c = ct / n_tiles;
    const int tile = ct - c * n_tiles;
    const int ty = tile / a.tile_w, tx = tile - ty * a.tile_w;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const bool producer = w == kSlabConsumers;

    const int32_t seg_start = a.tile_offsets[ct];
    const int32_t cnt = a.rec_counts[ct];
    const int n_chunks = (cnt + CH - 1) / CH;

    if (tid == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(s_full + s, 1 + 32);          // expect_tx arrival + one cp.async arrival per producer lane
            mbar_init(s_empty + s, kSlabConsumers);  // one arrival per consumer warp
        }
        s_ndone = 0;
        mbar_init_fence();
    }
    __syncthreads();

    int lx = 0, ly = 0;
    if (!producer) pixel_of_thread(tid, lx, ly);
    const int j = tx * kTile + lx, i = ty * kTile + ly;
    const bool inside = !producer && (i < a.height) && (j < a.width);
    const float px = (float)j + 0.5f, py = (float)i + 0.5f;
    const int64_t pid = ((int64_t)c * a.height + i) * a.width + j;

    float T = 1.0f;
    int32_t cur_idx = -1;
    float2 out2[D2];
#pragma unroll
    for (int k = 0; k < D2; ++k) out2[k] = make_float2(0.f, 0.f);
    float outd = 0.f;

    if (producer) {
        // ------------------------------------------------------------------------------------ producer warp
        int stage = 0, phase = 0, issued = 0;
        uint32_t idm_next = 0u;
        if (n_chunks > 0 && lane < cnt)
            idm_next = __ldg(reinterpret_cast<const uint32_t *>(a.recs + 2 * ((int64_t)seg_start + lane)) + 3);
        for (int k = 0; k < n_chunks; ++k) {
            bool stop = ld_volatile_s32(&s_ndone) >= kSlabConsumers;
            if (!stop && k >= S) {
                while (!mbar_try_wait(s_empty + stage, phase ^ 1)) {
                    if (ld_volatile_s32(&s_ndone) >= kSlabConsumers) {
                        stop = true;
                        break;
                    }
                }
            }
            if (stop) break;
            const int64_t first = (int64_t)seg_start + (int64_t)k * CH;
            const int n_valid = min(CH, cnt - k * CH);
            const uint32_t idm = idm_next;
            idm_next = 0u;
            if (k + 1 < n_chunks && (k + 1) * CH + lane < cnt)
                idm_next = __ldg(reinterpret_cast<const uint32_t *>(a.recs + 2 * (first + CH + lane)) + 3);
            slab_issue_stage<D0, false>(a, c, first, n_valid, idm, s_rec + stage * CH * 2, s_col + stage * CH * D0,
                                 s_full + stage, lane);
            issued = k + 1;
            if (++stage == S) stage = 0, phase ^= 1;
        }
        // every copy that was issued must have landed before the ring is reused by the epilogue / the CTA exits
        for (int k = max(0, issued - S); k < issued; ++k) mbar_wait(s_full + (k % S), (k / S) & 1);
    } else {
        // ------------------------------------------------------------------------------------ consumer warps
        bool done = !inside;
        bool warp_done = __all_sync(0xffffffffu, done);
        if (warp_done && lane == 0) atomicAdd(&s_ndone, 1);
        const int64_t hb_base = ((int64_t)(seg_start >> 5) + ct) * kSlabConsumers + w;
        int stage = 0, phase = 0;
        for (int k = 0; k < n_chunks; ++k) {
            if (!warp_done) {
                mbar_wait(s_full + stage, phase);
            } else {
                bool quit = false;
                while (!mbar_try_wait(s_full + stage, phase)) {
                    if (ld_volatile_s32(&s_ndone) >= kSlabConsumers) {
                        quit = true;
                        break;
                    }
                }
                if (quit) break;
            }
            if (!warp_done) {
                // shared-window addresses of the stage (explicit: see lds128)
                const uint32_t recs = s_rec_addr + (uint32_t)stage * (CH * 32);
                const uint32_t cols = s_col_addr + (uint32_t)stage * (CH * D0 * 4);
                const int n_valid = min(CH, cnt - k * CH);
                const uint32_t idm = lane < n_valid ? lds32u(recs + lane * 32 + 12) : 0u;
                uint32_t bits = __ballot_sync(0xffffffffu, (idm >> (24 + w)) & 1u);
                uint32_t hitbits = 0u;
                const int32_t idx0 = seg_start + k * CH;
                // one record of the hit list for this pixel.  Whether the whole warp has finished is asked once per chunk,
                // not per record (a vote + branch per composite was 12 % of the kernel's instructions): a finished warp
                // walks the rest of its chunk with valid == false everywhere, i.e. an exponent and a vote per record
                auto composite = [&](int t, float power, float L, float depth) {
                    const float alpha = fminf(kAlphaMax, ex2_approx(power));
                    const bool valid = !done && power <= L && alpha >= kAlphaMin;
                    if (!__any_sync(0xffffffffu, valid)) return;
                    if constexpr (kMasks) hitbits |= 1u << t;
                    if (valid) {
                        const float next_T = T * (1.0f - alpha);
                        if (next_T <= kTMin) {
                            done = true;
                        } else {
                            const float vis = alpha * T;
                            const float2 vis2 = make_float2(vis, vis);
                            const uint32_t cp = cols + (uint32_t)t * (D0 * 4);
#pragma unroll
                            for (int k4 = 0; k4 < D0 / 4; ++k4) {
                                const float4 cv = lds128(cp + 16 * k4);
                                out2[2 * k4] = __ffma2_rn(make_float2(cv.x, cv.y), vis2, out2[2 * k4]);
                                out2[2 * k4 + 1] = __ffma2_rn(make_float2(cv.z, cv.w), vis2, out2[2 * k4 + 1]);
                            }
                            if constexpr (DEPTH) outd = fmaf(depth, vis, outd);
                            cur_idx = idx0 + t;
                            T = next_T;
                        }
                    }
                };
                while (bits) {
                    // two hits per trip: both exponents are evaluated before either is composited
                    const int ta = __ffs(bits) - 1;
                    bits &= bits - 1;
                    const bool has_b = bits != 0u;
                    const int tb = has_b ? __ffs(bits) - 1 : ta;
                    bits &= bits - 1;  // no-op when bits == 0
                    const float4 ga = lds128(recs + ta * 32), ca = lds128(recs + ta * 32 + 16);
                    const float4 gb = lds128(recs + tb * 32), cb = lds128(recs + tb * 32 + 16);
                    const float dxa = ga.x - px, dya = ga.y - py, dxb = gb.x - px, dyb = gb.y - py;
                    const float pa = fmaf(ca.z * dya, dya, fmaf(fmaf(ca.y, dya, ca.x * dxa), dxa, ga.z));
                    const float pb = fmaf(cb.z * dyb, dyb, fmaf(fmaf(cb.y, dyb, cb.x * dxb), dxb, gb.z));
                    composite(ta, pa, ga.z, ca.w);
                    if (has_b) composite(tb, pb, gb.z, cb.w);
                }
                warp_done = __all_sync(0xffffffffu, done);
                if constexpr (kMasks) {
                    if (lane == 0) a.hit_bits[hb_base + (int64_t)k * kSlabConsumers] = hitbits;
                }
                if (warp_done && lane == 0) atomicAdd(&s_ndone, 1);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(s_empty + stage);
            if (++stage == S) stage = 0, phase ^= 1;
        }
    }
    __syncthreads();  // the ring is idle: all issued copies have landed, all consumers have left the main loop

    // epilogue: background, ED normalisation, coalesced store through shared memory
    float *s_out = reinterpret_cast<float *>(smem_raw);
    if (!producer) {
        float out[D];
#pragma unroll
        for (int k = 0; k < D0; ++k) out[k] = (k & 1) ? out2[k >> 1].y : out2[k >> 1].x;
        if constexpr (DEPTH) out[D - 1] = outd;
        const float alpha_out = 1.0f - T;
        if (a.backgrounds) {
#pragma unroll
            for (int k = 0; k < D0; ++k) out[k] = fmaf(T, __ldg(a.backgrounds + (int64_t)c * D0 + k), out[k]);
        }
        if (inside) {
            render_alphas[pid] = alpha_out;
            last_ids[pid] = cur_idx;
            if constexpr (DEPTH) {
                if (a.normalize_depth) {
                    acc_depth[pid] = out[D - 1];
                    out[D - 1] = out[D - 1] / fmaxf(alpha_out, 1e-10f);
                }
            }
        }
        float *dst = s_out + (ly * kTile + lx) * DP;
#pragma unroll
        for (int k = 0; k < D; ++k) dst[k] = out[k];
    }
    __syncthreads();
    if (!producer) {
        // every tile row is one contiguous run of 16*D floats in the channels-last image
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
                if (col_ok[q]) dst_row[q * kBlendThreads] = s_out[r * kTile * DP + src_off[q]];
            dst_row += (int64_t)a.width * D;
        }
    }
}

template <int D0, bool DEPTH>
static int launch_fwd_slab(bool masks, const SlabArgs &a, float *rc, float *ra, int32_t *li, float *ad, cudaStream_t st) {
    constexpr size_t smem = FwdSlabCfg<D0, DEPTH>::smem_bytes();
    const int grid = a.C * a.tile_w * a.tile_h;
    // the opt-in is per device and cheap: set it on every launch (no per-process "configured" flag)
    if (masks) {
        if (cudaFuncSetAttribute(blend_fwd_slab_kernel<D0, DEPTH, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)smem) != cudaSuccess)
            return 1;
        blend_fwd_slab_kernel<D0, DEPTH, true><<<grid, kSlabThreads, smem, st>>>(a, rc, ra, li, ad);
    } else {
        if (cudaFuncSetAttribute(blend_fwd_slab_kernel<D0, DEPTH, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)smem) != cudaSuccess)
            return 1;
        blend_fwd_slab_kernel<D0, DEPTH, false><<<grid, kSlabThreads, smem, st>>>(a, rc, ra, li, ad);
    }
    return 0;
}

int launch_blend_fwd_slab(int D0, bool depth, bool masks, const SlabArgs &a, float *rc, float *ra, int32_t *li,
                          float *ad, cudaStream_t st) {
#define X(n)                                                                              \
    case n:                                                                               \
        return depth ? launch_fwd_slab<n, true>(masks, a, rc, ra, li, ad, st)             \
                     : launch_fwd_slab<n, false>(masks, a, rc, ra, li, ad, st);
    switch (D0) {
        X(4) X(8) X(16) X(32)
        default: return -1;
    }
#undef X
}

}  // namespace d4

using namespace d4;

int d4::check_slab_args(const char *name, const SlabArgs &a, int D0, int tile_size) {
    D4_CHECK_ARG(tile_size == kTile, "%s: only tile_size 16 is built", name);
    D4_CHECK_ARG(a.C >= 1 && a.G >= 0 && a.width > 0 && a.height > 0, "%s: bad sizes", name);
    D4_CHECK_ARG(a.tile_w == (a.width + kTile - 1) / kTile && a.tile_h == (a.height + kTile - 1) / kTile,
                 "%s: tile grid does not match the image size", name);
    D4_CHECK_ARG(a.recs && a.tile_offsets && a.rec_counts && (a.colors || a.G == 0), "%s: null pointer", name);
    D4_CHECK_ARG(D0 == 4 || D0 == 8 || D0 == 16 || D0 == 32, "%s: colour width %d not built (pad to 4, 8, 16 or 32)", name, D0);
    D4_CHECK_ARG(((uintptr_t)a.recs & 15) == 0 && ((uintptr_t)a.colors & 15) == 0 && (a.colors_cs & 3) == 0,
                 "%s: records and colours must be 16-byte aligned", name);
    D4_CHECK_ARG(a.G < (1 << 24), "%s: packed records hold 24-bit Gaussian ids (G = %d)", name, a.G);
    return 0;
}

static int blend_fwd_slab_any(const char *name, int variant, const void *recs, const int32_t *tile_offsets,
                              const int32_t *rec_counts, const float *colors, int64_t colors_cam_stride,
                              const float *backgrounds, int C, int G, int D0, int with_depth, int width, int height,
                              int tile_size, int tile_w, int tile_h, int normalize_depth, float *render_colors,
                              float *render_alphas, int32_t *last_ids, float *acc_depth, uint32_t *hit_bits,
                              d4_stream_t stream) {
    SlabArgs a{(const float4 *)recs, tile_offsets, rec_counts, colors, colors_cam_stride, backgrounds, hit_bits,
               C, G, width, height, tile_w, tile_h, normalize_depth};
    if (int rc = check_slab_args(name, a, D0, tile_size)) return rc;
    D4_CHECK_ARG(render_colors && render_alphas && last_ids, "%s: null output", name);
    D4_CHECK_ARG(!normalize_depth || (with_depth && acc_depth), "%s: normalize_depth needs the depth channel and acc_depth", name);
    D4_CHECK_ARG(variant >= 0 && variant <= 1, "%s: variant must be 0 (fp32 pipe) or 1 (tensor cores)", name);
    int rc = -1;
    if (variant == 1)  // served for the 16-colour records, otherwise falls through
        rc = launch_blend_fwd_slab_tc(D0, with_depth != 0, hit_bits != nullptr, a, render_colors, render_alphas, last_ids,
                                      acc_depth, as_stream(stream));
    if (rc < 0)
        rc = launch_blend_fwd_slab(D0, with_depth != 0, hit_bits != nullptr, a, render_colors, render_alphas, last_ids,
                                   acc_depth, as_stream(stream));
    if (rc != 0) {
        set_error("%s: %s", name, rc < 0 ? "channel count not built" : "kernel configuration failed");
        return rc < 0 ? 2 : 1;
    }
    D4_CHECK_LAUNCH(name);
    return 0;
}

extern "C" int d4_blend_fwd_slab(const void *recs, const int32_t *tile_offsets, const int32_t *rec_counts,
                                 const float *colors, int64_t colors_cam_stride, const float *backgrounds, int C, int G,
                                 int D0, int with_depth, int width, int height, int tile_size, int tile_w, int tile_h,
                                 int normalize_depth, float *render_colors, float *render_alphas, int32_t *last_ids,
                                 float *acc_depth, uint32_t *hit_bits, d4_stream_t stream) {
    return blend_fwd_slab_any("d4_blend_fwd_slab", D4_BLEND_FWD_DEFAULT_VARIANT, recs, tile_offsets, rec_counts, colors,
                              colors_cam_stride, backgrounds, C, G, D0, with_depth, width, height, tile_size, tile_w,
                              tile_h, normalize_depth, render_colors, render_alphas, last_ids, acc_depth, hit_bits, stream);
}

extern "C" int d4_blend_fwd_slab_default_variant(void) { return D4_BLEND_FWD_DEFAULT_VARIANT; }

extern "C" int d4_blend_fwd_slab_variant(const void *recs, const int32_t *tile_offsets, const int32_t *rec_counts,
                                         const float *colors, int64_t colors_cam_stride, const float *backgrounds,
                                         int C, int G, int D0, int with_depth, int width, int height, int tile_size,
                                         int tile_w, int tile_h, int normalize_depth, float *render_colors,
                                         float *render_alphas, int32_t *last_ids, float *acc_depth, uint32_t *hit_bits,
                                         int variant, d4_stream_t stream) {
    return blend_fwd_slab_any("d4_blend_fwd_slab_variant", variant, recs, tile_offsets, rec_counts, colors,
                              colors_cam_stride, backgrounds, C, G, D0, with_depth, width, height, tile_size, tile_w,
                              tile_h, normalize_depth, render_colors, render_alphas, last_ids, acc_depth, hit_bits, stream);
}
