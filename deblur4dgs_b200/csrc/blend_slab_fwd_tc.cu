// blend_slab_fwd_tc.cu -- row a10 of SURVEY.md section 8, queue + tensor-core formulation of blend_slab_fwd.cu for the
// 16-colour(+depth) records of the benchmark configuration (gsplat rasterize_to_pixels fwd + depth channel + ED
// normalisation, call site flow3d/scene_model.py:360-373).
//
// blend_slab_fwd.cu composites a record the moment a warp meets it: per visited (warp, record) it re-reads the
// record and the 16 colours from shared memory for 32 pixels (6 LDS.128) and spends 8 packed FMAs per lane on
// `out += alpha T colour`.  Here the colour accumulation of 16 records at a time is ONE small GEMM per warp,
//     O[32 pixels x 16 colours] += W^T[32 x 16 rows] . C[16 rows x 16 colours],      W[row][pixel] = alpha T
// on mma.sync.m16n8k8 (TF32 operands, fp32 accumulate, 3xTF32 split => fp32-grade), with O resident in the
// accumulator fragments for the whole tile.  What stays on the fp32 pipe, lane = pixel, is exactly the serial part:
// exponent, the alpha / transmittance decisions (same arithmetic, bit for bit, as blend_slab_fwd.cu and as the
// backward kernels -- the hit words and last_ids they consume are decided here), T and the depth channel.
// Data movement is the backward's: producer warp + mbarrier ring, lane L copies record L of a chunk (if the record's
// reach mask has the warp's bit) with its colour row into the warp's own 16-row queue and gives the stage back.
#include "slab.cuh"

namespace d4 {

namespace {

constexpr int kD0 = 16;    // colour channels of this specialisation
constexpr int kRows = 16;  // rows of the per-warp queue == K of one accumulation
constexpr int kWs = 40;    // row stride of the weight tile: A-fragment loads hit bank 8 t + g, row stores are linear

__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void split_tf32(float x, uint32_t &hi, uint32_t &lo) {
    hi = __float_as_uint(x) & 0xffffe000u;
    lo = __float_as_uint(x - __uint_as_float(hi));
}

// queued colour rows: 16-byte piece q of row r sits at piece position q ^ qkey(r); the B-fragment loads
// (rows 8 ks + t (+4), channel 8 nt + g) then touch 32 distinct banks
__device__ __forceinline__ int qkey(int row) { return ((row >> 1) & 1) * 2; }

template <bool DEPTH>
struct FwdTcCfg {
    static constexpr int D = kD0 + (DEPTH ? 1 : 0);
    static constexpr int DP = D | 1;  // odd stride of the epilogue transpose buffer
    static constexpr int NW = kSlabConsumers;
    static constexpr int kCtas = 3;
    __host__ __device__ static constexpr size_t warp_bytes() {
        return sizeof(float4) * kRows * 2 + sizeof(float) * kRows * kD0 + sizeof(float) * kRows * kWs +
               sizeof(int32_t) * kRows;
    }
    __host__ __device__ static constexpr size_t stage_bytes() { return (size_t)kSlabChunk * (32 + 4 * kD0); }
    static constexpr int kStages = 6;
    __host__ __device__ static constexpr size_t work_bytes() { return NW * warp_bytes() + kStages * stage_bytes(); }
    __host__ __device__ static constexpr size_t epi_bytes() { return sizeof(float) * kBlendThreads * DP; }
    static constexpr size_t smem_bytes() {
        return (work_bytes() > epi_bytes() ? work_bytes() : epi_bytes()) + 2 * kStages * sizeof(uint64_t) + 16;
    }
};

}  // namespace

template <bool DEPTH, bool kMasks>
__global__ void __launch_bounds__(kSlabThreads, (FwdTcCfg<DEPTH>::kCtas))
blend_fwd_slab_tc_kernel(SlabArgs a, float *__restrict__ render_colors, float *__restrict__ render_alphas,
                         int32_t *__restrict__ last_ids, float *__restrict__ acc_depth) {
    using Cfg = FwdTcCfg<DEPTH>;
    constexpr int D0 = kD0, D = Cfg::D, DP = Cfg::DP, S = Cfg::kStages, CH = kSlabChunk, NW = Cfg::NW;
    constexpr int GR = kRows, WS = kWs, U = 4;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float4 *s_rec = reinterpret_cast<float4 *>(smem_raw);                              // [S][CH][2]
    float *s_col = reinterpret_cast<float *>(s_rec + S * CH * 2);                      // [S][CH][D0] (swizzled rows)
    unsigned char *s_warp_all = reinterpret_cast<unsigned char *>(s_col + S * CH * D0);
    constexpr size_t data_bytes = Cfg::work_bytes() > Cfg::epi_bytes() ? Cfg::work_bytes() : Cfg::epi_bytes();
    uint64_t *s_full = reinterpret_cast<uint64_t *>(smem_raw + ((data_bytes + 15) & ~(size_t)15));  // [S]
    uint64_t *s_empty = s_full + S;                                                                 // [S]
    __shared__ int s_ndone;  // consumer warps with all pixels saturated (or outside the image)

    const int n_tiles = a.tile_w * a.tile_h;
    const int ct = blockIdx.x;
    const int c = ct / n_tiles;
    const int tile = ct - c * n_tiles;
    const int ty = tile / a.tile_w, tx = tile - ty * a.tile_w;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const bool producer = w == NW;

    const int32_t seg_start = a.tile_offsets[ct];
    const int32_t cnt = a.rec_counts[ct];
    const int n_chunks = (cnt + CH - 1) / CH;

    if (tid == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(s_full + s, 1 + 32);  // expect_tx arrival + one cp.async arrival per producer lane
            mbar_init(s_empty + s, NW);     // one arrival per consumer warp
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
    float outd = 0.f;
    // O[32 pixels x 16 colours] of the warp in accumulator fragments: acc[mt][nt] = pixels 16 mt + g (+8), colours
    // 8 nt + 2 t (+1)
    float acc[2][2][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[mt][nt][e] = 0.f;

    if (producer) {
        // ------------------------------------------------------------------------------------ producer warp
        int stage = 0, phase = 0, issued = 0;
        uint32_t idm_next = 0u;
        if (n_chunks > 0 && lane < cnt)
            idm_next = __ldg(reinterpret_cast<const uint32_t *>(a.recs + 2 * ((int64_t)seg_start + lane)) + 3);
        for (int k = 0; k < n_chunks; ++k) {
            bool stop = ld_volatile_s32(&s_ndone) >= NW;
            if (!stop && k >= S) {
                while (!mbar_try_wait(s_empty + stage, phase ^ 1)) {
                    if (ld_volatile_s32(&s_ndone) >= NW) {
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
            slab_issue_stage<D0, true>(a, c, first, n_valid, idm, s_rec + stage * CH * 2, s_col + stage * CH * D0,
                                       s_full + stage, lane);
            issued = k + 1;
            if (++stage == S) stage = 0, phase ^= 1;
        }
        // every copy that was issued must have landed before the ring is reused by the epilogue / the CTA exits
        for (int k = max(0, issued - S); k < issued; ++k) mbar_wait(s_full + (k % S), (k / S) & 1);
    } else {
        // ------------------------------------------------------------------------------------ consumer warps
        unsigned char *s_warp = s_warp_all + w * Cfg::warp_bytes();
        float4 *s_qrec = reinterpret_cast<float4 *>(s_warp);          // [GR][2]   queued records
        float *s_qcol = reinterpret_cast<float *>(s_qrec + GR * 2);   // [GR][D0]  queued colour rows (qkey swizzle)
        float *s_w = s_qcol + GR * D0;                                // [GR][WS]  alpha * T per (row, pixel)
        int32_t *s_qidx = reinterpret_cast<int32_t *>(s_w + GR * WS);  // [GR]      record indices
        const int fg = lane >> 2, ft = lane & 3;
        const int64_t hb_base = ((int64_t)(seg_start >> 5) + ct) * NW + w;

        bool done = !inside;
        bool warp_done = __all_sync(0xffffffffu, done);
        if (warp_done && lane == 0) atomicAdd(&s_ndone, 1);

        int stage = 0, phase = 0;
        int k = -1;            // chunk being drained
        bool holding = false;  // the ring stage of chunk k is still in use
        bool exhausted = n_chunks == 0;
        uint32_t bits = 0u;    // records of chunk k that reach this warp's block and are not queued yet
        int nb = 0;            // rows queued (warp-uniform)

        auto release = [&]() {
            __syncwarp();
            if (lane == 0) mbar_arrive(s_empty + stage);
            if (++stage == S) stage = 0, phase ^= 1;
            holding = false;
        };

        while (!warp_done) {
            // ---- fill the queue front to back: lane L owns record L of the chunk
            while (nb < GR) {  // warp-uniform
                if (bits == 0u) {
                    if (holding) release();
                    if (k + 1 >= n_chunks) {
                        exhausted = true;
                        break;
                    }
                    ++k;
                    mbar_wait(s_full + stage, phase);
                    holding = true;
                    const int n_valid = min(CH, cnt - k * CH);
                    const uint32_t idm =
                        lane < n_valid ? reinterpret_cast<const uint32_t *>(s_rec + (stage * CH + lane) * 2)[3] : 0u;
                    bits = __ballot_sync(0xffffffffu, (idm >> (24 + w)) & 1u);
                    // the word of every chunk this warp meets starts at zero; hits are OR-ed in after their evaluation
                    if constexpr (kMasks) {
                        if (lane == 0) a.hit_bits[hb_base + (int64_t)k * NW] = 0u;
                    }
                    continue;
                }
                const float4 *recs = s_rec + stage * CH * 2;
                const float *cols = s_col + stage * CH * D0;
                const bool hit = (bits >> lane) & 1u;
                const int row = nb + __popc(bits & ((1u << lane) - 1u));
                const bool take = hit && row < GR;
                if (take) {
                    const int h0 = (lane >> 2) & 1;  // halves in the order that keeps a quarter-warp on distinct banks
                    const float4 ra0 = recs[2 * lane + h0], ra1 = recs[2 * lane + (h0 ^ 1)];
                    s_qrec[2 * row + h0] = ra0;
                    s_qrec[2 * row + (h0 ^ 1)] = ra1;
                    const int ks = slab_key<D0>(lane), kr = qkey(row);
#pragma unroll
                    for (int k4 = 0; k4 < D0 / 4; ++k4)  // logical piece k4: swizzled by slot in the stage, by row in the queue
                        *reinterpret_cast<float4 *>(s_qcol + row * D0 + 4 * (k4 ^ kr)) =
                            *reinterpret_cast<const float4 *>(cols + lane * D0 + 4 * ((k4 ^ ks) & 3));
                    s_qidx[row] = seg_start + k * CH + lane;
                }
                const uint32_t taken = __ballot_sync(0xffffffffu, take);
                bits &= ~taken;
                nb += __popc(taken);
            }
            if (nb == 0) break;  // the stream is exhausted and nothing is queued
            if (nb < GR && lane >= nb && lane < GR) {
                // last, partial group: inert rows (an exponent of -1e30 never passes the alpha test; zero colours)
                s_qrec[2 * lane] = make_float4(0.f, 0.f, -1e30f, 0.f);
                s_qrec[2 * lane + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int k4 = 0; k4 < D0 / 4; ++k4) *reinterpret_cast<float4 *>(s_qcol + lane * D0 + 4 * k4) = make_float4(0.f, 0.f, 0.f, 0.f);
                s_qidx[lane] = -1;
            }
            __syncwarp();

            // ---- the serial part, lane = pixel: U rows per trip, all sixteen rows (inert ones park w = 0)
            uint32_t hitrows = 0u;
#pragma unroll 1
            for (int r0 = 0; r0 < GR; r0 += U) {
                float pw[U], lz[U], dep[U];
                const int4 qi = *reinterpret_cast<const int4 *>(s_qidx + r0);
                const int32_t qidx[U] = {qi.x, qi.y, qi.z, qi.w};
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const float4 g0 = s_qrec[2 * (r0 + u)], cn = s_qrec[2 * (r0 + u) + 1];
                    const float dx = g0.x - px, dy = g0.y - py;
                    pw[u] = fmaf(cn.z * dy, dy, fmaf(fmaf(cn.y, dy, cn.x * dx), dx, g0.z));
                    lz[u] = g0.z;
                    dep[u] = cn.w;
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const float alpha = fminf(kAlphaMax, ex2_approx(pw[u]));
                    const bool valid = !done && pw[u] <= lz[u] && alpha >= kAlphaMin;
                    if constexpr (kMasks) hitrows |= __any_sync(0xffffffffu, valid) ? (1u << (r0 + u)) : 0u;
                    const float next_T = T * (1.0f - alpha);
                    const bool stop = valid && next_T <= kTMin;  // saturates BEFORE this record is included
                    const bool take = valid && !stop;
                    const float vis = take ? alpha * T : 0.f;
                    done = done || stop;
                    T = take ? next_T : T;
                    cur_idx = take ? qidx[u] : cur_idx;
                    if constexpr (DEPTH) outd = fmaf(dep[u], vis, outd);
                    s_w[(r0 + u) * WS + lane] = vis;
                }
            }
            __syncwarp();

            // ---- O += W^T . C on the tensor pipe: A = W^T (m = pixel, k = row), B = C (k = row, n = colour)
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
                uint32_t bh[2][2], bl[2][2];
#pragma unroll
                for (int nt = 0; nt < 2; ++nt) {
                    // rows 8 ks + ft and + 4, channel 8 nt + fg, at position ch ^ (8 * bit1(row)) = ch ^ 8 (ft >> 1)
                    const int pos = ((8 * nt) ^ (8 * (ft >> 1))) + fg;
                    split_tf32(s_qcol[(8 * ks + ft) * D0 + pos], bh[nt][0], bl[nt][0]);
                    split_tf32(s_qcol[(8 * ks + ft + 4) * D0 + pos], bh[nt][1], bl[nt][1]);
                }
#pragma unroll
                for (int mt = 0; mt < 2; ++mt) {
                    uint32_t ah[4], al[4];
                    const float *wp = s_w + (8 * ks + ft) * WS + 16 * mt + fg;
                    split_tf32(wp[0], ah[0], al[0]);           // (pixel g,     row t)
                    split_tf32(wp[8], ah[1], al[1]);           // (pixel g + 8, row t)
                    split_tf32(wp[4 * WS], ah[2], al[2]);      // (pixel g,     row t + 4)
                    split_tf32(wp[4 * WS + 8], ah[3], al[3]);  // (pixel g + 8, row t + 4)
#pragma unroll
                    for (int nt = 0; nt < 2; ++nt) {
                        mma_tf32(acc[mt][nt], al, bh[nt][0], bh[nt][1]);
                        mma_tf32(acc[mt][nt], ah, bl[nt][0], bl[nt][1]);
                        mma_tf32(acc[mt][nt], ah, bh[nt][0], bh[nt][1]);
                    }
                }
            }
            // ---- hit words: lane r reports row r
            if constexpr (kMasks) {
                if (lane < nb && ((hitrows >> lane) & 1u)) {
                    const int rel = s_qidx[lane] - seg_start;
                    atomicOr(a.hit_bits + hb_base + (int64_t)(rel >> 5) * NW, 1u << (rel & 31));
                }
            }
            __syncwarp();
            nb = 0;
            warp_done = __all_sync(0xffffffffu, done);
            if (warp_done) {
                if (lane == 0) atomicAdd(&s_ndone, 1);
                break;
            }
            if (exhausted) break;
        }
        // ---- a warp that has finished keeps returning the remaining stages until the producer stops streaming
        if (holding) release();
        if (!exhausted) {
            for (++k; k < n_chunks; ++k) {
                bool quit = false;
                while (!mbar_try_wait(s_full + stage, phase)) {
                    if (ld_volatile_s32(&s_ndone) >= NW) {
                        quit = true;
                        break;
                    }
                }
                if (quit) break;
                holding = true;
                release();
            }
        }
    }
    __syncthreads();  // the ring is idle: all issued copies have landed, all consumers have left the main loop

    // epilogue: background, ED normalisation, coalesced store through shared memory
    float *s_out = reinterpret_cast<float *>(smem_raw);
    if (!producer) {
        const int fg = lane >> 2, ft = lane & 3;
        const float alpha_out = 1.0f - T;
        if (inside) {
            render_alphas[pid] = alpha_out;
            last_ids[pid] = cur_idx;
        }
        if constexpr (DEPTH) {
            float od = outd;
            if (inside && a.normalize_depth) {
                acc_depth[pid] = od;
                od = od / fmaxf(alpha_out, 1e-10f);
            }
            s_out[(ly * kTile + lx) * DP + D0] = od;
        }
        // colours from the accumulator fragments: pixel p = 16 mt + 8 half + g of the warp, channels 8 nt + 2 t (+1)
        float bgv[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
        if (a.backgrounds) {
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) {
                bgv[nt][0] = __ldg(a.backgrounds + (int64_t)c * D0 + 8 * nt + 2 * ft);
                bgv[nt][1] = __ldg(a.backgrounds + (int64_t)c * D0 + 8 * nt + 2 * ft + 1);
            }
        }
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const int p = 16 * mt + 8 * half + fg;
                const float Tp = __shfl_sync(0xffffffffu, T, p);
                int plx, ply;
                pixel_of_thread(w * 32 + p, plx, ply);
                float *dst = s_out + (ply * kTile + plx) * DP + 2 * ft;
#pragma unroll
                for (int nt = 0; nt < 2; ++nt) {
                    dst[8 * nt] = fmaf(Tp, bgv[nt][0], acc[mt][nt][2 * half]);
                    dst[8 * nt + 1] = fmaf(Tp, bgv[nt][1], acc[mt][nt][2 * half + 1]);
                }
            }
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
            const int pxl = col / D, kk = col - pxl * D;
            src_off[q] = pxl * DP + kk;
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

template <bool DEPTH, bool kMasks>
static int launch_fwd_tc(const SlabArgs &a, float *rc, float *ra, int32_t *li, float *ad, cudaStream_t st) {
    constexpr size_t smem = FwdTcCfg<DEPTH>::smem_bytes();
    if (cudaFuncSetAttribute(blend_fwd_slab_tc_kernel<DEPTH, kMasks>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)smem) != cudaSuccess)
        return 1;
    const int grid = a.C * a.tile_w * a.tile_h;
    blend_fwd_slab_tc_kernel<DEPTH, kMasks><<<grid, kSlabThreads, smem, st>>>(a, rc, ra, li, ad);
    return 0;
}

// Returns -1 when the call is not served here (D0 != 16).
int launch_blend_fwd_slab_tc(int D0, bool depth, bool masks, const SlabArgs &a, float *rc, float *ra, int32_t *li,
                             float *ad, cudaStream_t st) {
    if (D0 != kD0) return -1;
    if (depth) return masks ? launch_fwd_tc<true, true>(a, rc, ra, li, ad, st) : launch_fwd_tc<true, false>(a, rc, ra, li, ad, st);
    return masks ? launch_fwd_tc<false, true>(a, rc, ra, li, ad, st) : launch_fwd_tc<false, false>(a, rc, ra, li, ad, st);
}

}  // namespace d4
