// slab.cuh -- the "slab" data path of the blend kernels (rows a10 / a11 of SURVEY.md section 8).
//
// After tile binning every (camera, tile) owns a contiguous, depth-sorted run of PACKED 32-byte intersection
// records (d4_isect_pack / the tail of tile_sort_kernel):
//     rec[0] = (x, y, L = log2(opacity), idm)      idm = local Gaussian id (24 bits) | reach mask << 24
//     rec[1] = (A', B', C', depth)                 conic pre-scaled to base 2: A' = -a/2 log2e, B' = -b log2e, ...
//   => opacity * exp(-sigma) = exp2(L + A' dx^2 + B' dx dy + C' dy^2)
// Records whose reach mask is empty (the ellipse {alpha >= 1/255} misses all eight 8x4 pixel blocks of the tile)
// are dropped at packing time: the run of tile t is [tile_offsets[t], tile_offsets[t] + rec_counts[t]).
//
// The blend kernels move a tile's run through a ring of shared-memory stages (32 records each):
//   * a PRODUCER warp streams the records with one bulk async copy per stage (cp.async.bulk -> SASS UBLKCP, completion
//     counted in bytes on the stage's "full" mbarrier) and gathers the colour rows of the 32 Gaussians with 16-byte
//     cp.async (LDGSTS) that arrive on the same mbarrier -- no register staging, no CTA-wide barrier;
//   * the eight CONSUMER warps (one 8x4 pixel block each) wait on "full", walk the hits of their block and arrive on
//     the stage's "empty" mbarrier; warps of a tile may drift apart by the depth of the ring.
#pragma once
#include "blend_common.cuh"

namespace d4 {

constexpr int kSlabChunk = 32;                      // records per ring stage
constexpr int kSlabConsumers = kBlendThreads / 32;  // 8 warps, one 8x4 pixel block each
constexpr int kSlabThreads = kBlendThreads + 32;    // + the producer warp
constexpr uint32_t kRecIdMask = 0x00ffffffu;        // local Gaussian id bits of rec[0].w

struct SlabArgs {
    const float4 *recs;           // [n_isects][2] packed records (see above)
    const int32_t *tile_offsets;  // [C * tiles] first record of every (camera, tile)
    const int32_t *rec_counts;    // [C * tiles] records kept for the tile
    const float *colors;          // [G, D0] or [C, G, D0], D0 a multiple of 4 and 16-byte aligned
    int64_t colors_cs;            // elements between cameras (0 = shared)
    const float *backgrounds;     // [C, D0] or null
    // per (tile chunk of 32 records, consumer warp): bit i set iff record i of the chunk passed the alpha test on
    // some pixel of the warp's block in the forward.  Word index = ((tile_offsets[t] >> 5) + t + chunk) * 8 + warp.
    uint32_t *hit_bits;
    int C, G, width, height, tile_w, tile_h;
    int normalize_depth;
};

// words of SlabArgs::hit_bits for n_isects intersections in n_segments (camera, tile) runs
static inline size_t slab_hit_words(int64_t n_isects, int64_t n_segments) {
    return (size_t)((n_isects >> 5) + n_segments + 1) * kSlabConsumers;
}

// ---------------------------------------------------------------------------------------------- PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0u;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
// global -> shared bulk async copy (TMA engine, 1-D); bytes a multiple of 16, both addresses 16-byte aligned
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void cp_async16(void *dst_smem, const void *src_gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src_gmem) : "memory");
}
// arrive on `bar` once all cp.async issued so far by this thread have landed (does not bump the pending count)
__device__ __forceinline__ void cp_async_arrive_noinc(uint64_t *bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ int ld_volatile_s32(const int *p) { return *reinterpret_cast<const volatile int *>(p); }
// shared-memory loads at explicit 32-bit shared-window addresses: keeps the hot loops free of the generic-address
// bookkeeping (cluster-window base, S2R) the compiler otherwise re-derives per load under register pressure
__device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t lds32u(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}

// Backward: colour rows sit swizzled in the ring stages and in the per-warp queues: the 16-byte piece q of the
// row of slot s is stored at piece position q ^ slab_key<D0>(s).  Eight consecutive slots then touch eight different
// 16-byte bank groups per piece, so the per-lane LDS.128 / STS.128 of the backward's queue copy are conflict-free.
template <int D0>
__host__ __device__ __forceinline__ constexpr int slab_key(int s) {
    constexpr int PPS = D0 / 4;
    return PPS == 1 ? 0 : (PPS == 2 ? ((s >> 2) & 1) : (PPS == 4 ? ((s >> 1) & 3) : (s & 7)));
}

// ---------------------------------------------------------------------------------------------- producer
// One ring stage: records by bulk copy, colour rows by 16-byte async gathers.  `idm` = rec[0].w of record
// first + lane (0 beyond n_valid).  full barrier: 1 (expect_tx) + 32 (cp.async arrivals) pending arrivals.
template <int D0, bool kSwizzle>
__device__ __forceinline__ void slab_issue_stage(const SlabArgs &a, int c, int64_t first, int n_valid, uint32_t idm,
                                                 float4 *s_rec_stage, float *s_col_stage, uint64_t *full, int lane) {
    constexpr int PPS = D0 / 4;  // 16-byte pieces per colour row
    static_assert(D0 % 4 == 0 && D0 >= 4, "slab path stages colour rows in 16-byte pieces");
    if (lane == 0) {
        const uint32_t bytes = (uint32_t)n_valid * 32u;
        mbar_arrive_expect_tx(full, bytes);
        bulk_g2s(s_rec_stage, a.recs + 2 * first, bytes, full);
    }
    const float *cbase = a.colors + (int64_t)c * a.colors_cs;
#pragma unroll
    for (int p = 0; p < PPS; ++p) {
        const int e = p * 32 + lane;
        const int slot = e / PPS, q = e - slot * PPS;
        const uint32_t id = __shfl_sync(0xffffffffu, idm, slot) & kRecIdMask;
        if (slot < n_valid)
            cp_async16(s_col_stage + slot * D0 + (kSwizzle ? ((q ^ slab_key<D0>(slot)) & (PPS - 1)) : q) * 4,
                       cbase + (int64_t)id * D0 + q * 4);
    }
    cp_async_arrive_noinc(full);
}

int check_slab_args(const char *name, const SlabArgs &a, int D0, int tile_size);

// entry points of the slab translation units (blend_slab_fwd.cu / blend_slab_bwd.cu)
// return 0 when launched, -1 when (D0, depth) is not built, 1 on a CUDA configuration error
int launch_blend_fwd_slab(int D0, bool depth, bool masks, const SlabArgs &a, float *render_colors, float *render_alphas,
                          int32_t *last_ids, float *acc_depth, cudaStream_t st);
int launch_blend_bwd_slab(int D0, bool depth, const SlabArgs &a, const float *render_alphas, const int32_t *last_ids,
                          const float *acc_depth, const float *v_render_colors, const float *v_render_alphas,
                          float *v_means2d, float *v_conics, float *v_colors, float *v_opacities, float *v_depths,
                          cudaStream_t st);
// blend_slab_bwd_tc.cu: the same backward with the per-group contractions on mma.sync (3xTF32); variant 1 = the
// gradient sums (phase 2), 2 = also <c_g, v_out> (phase 1).  Returns -1 when the call is not served by it
// (D0 != 16, or gradient rows that are not 8-byte aligned).
int launch_blend_bwd_slab_tc(int variant, int D0, bool depth, const SlabArgs &a, const float *render_alphas,
                             const int32_t *last_ids, const float *acc_depth, const float *v_render_colors,
                             const float *v_render_alphas, float *v_means2d, float *v_conics, float *v_colors,
                             float *v_opacities, float *v_depths, cudaStream_t st);
// blend_slab_fwd_tc.cu: the forward with the colour accumulation of 16 queued records at a time on mma.sync
// (O += W^T . C, 3xTF32).  Returns -1 when the call is not served by it (D0 != 16).
int launch_blend_fwd_slab_tc(int D0, bool depth, bool masks, const SlabArgs &a, float *render_colors,
                             float *render_alphas, int32_t *last_ids, float *acc_depth, cudaStream_t st);
// what d4_blend_fwd_slab runs (d4_blend_fwd_slab_variant selects explicitly)
#ifndef D4_BLEND_FWD_DEFAULT_VARIANT
#define D4_BLEND_FWD_DEFAULT_VARIANT 0
#endif
// what d4_blend_bwd_slab runs (d4_blend_bwd_slab_variant selects explicitly)
#ifndef D4_BLEND_BWD_DEFAULT_VARIANT
#define D4_BLEND_BWD_DEFAULT_VARIANT 2
#endif

}  // namespace d4
