// binning.cu -- row a9 of SURVEY.md section 8: the integer half of tile binning.
//   d4_exclusive_scan_i32 : cumsum of tiles_per_gauss   (torch.cumsum in gsplat.isect_tiles)
//   d4_sort_pairs_u64     : stable LSD radix sort        (cub::DeviceRadixSort in gsplat)
//   d4_tile_offsets       : gsplat.isect_offset_encode
//
// All three are HBM-bound integer passes.  Layout: keys u64 / values u32 as two
// separate streams (SoA), 2048 pairs per CTA, 8 bits per pass.  Per pass:
//   (1) per-CTA digit histogram            -> hist[digit][cta]      (digit-major)
//   (2) 256 CTAs scan their digit's row    -> exclusive per-CTA offsets + digit totals
//   (3) stable scatter: warp-level multi-split with __match_any_sync gives each
//       key its rank among equal digits; CTA-level digit bases come from (2).
// No decoupled look-back / spin-waiting anywhere: every kernel is a plain
// bulk-synchronous pass, so a scheduling surprise cannot hang the device.
#include <stdlib.h>

#include "slab.cuh"

namespace d4 {

// ----------------------------------------------------------------------------- scan
constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;

__device__ __forceinline__ int64_t block_sum_i64(int64_t v, int64_t *smem /*[8]*/) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) smem[w] = v;
    __syncthreads();
    int64_t t = 0;
#pragma unroll
    for (int i = 0; i < kScanThreads / 32; ++i) t += smem[i];
    return t;
}

__global__ void __launch_bounds__(kScanThreads)
scan_block_sums_kernel(const int32_t *__restrict__ in, int64_t n, int64_t *__restrict__ block_sums,
                       int32_t *__restrict__ block_max /* nullable */) {
    __shared__ int64_t sm[kScanThreads / 32];
    __shared__ int32_t smax[kScanThreads / 32];
    int64_t base = (int64_t)blockIdx.x * kScanTile;
    int64_t acc = 0;
    int32_t mx = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        int64_t j = base + (int64_t)i * kScanThreads + threadIdx.x;
        if (j < n) {
            const int32_t v = in[j];
            acc += v;
            mx = max(mx, v);
        }
    }
    int64_t t = block_sum_i64(acc, sm);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = t;
    if (block_max) {
        mx = __reduce_max_sync(0xffffffffu, mx);
        if ((threadIdx.x & 31) == 0) smax[threadIdx.x >> 5] = mx;
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int i = 1; i < kScanThreads / 32; ++i) mx = max(mx, smax[i]);
            block_max[blockIdx.x] = mx;
        }
    }
}

__global__ void __launch_bounds__(kScanThreads)
scan_apply_kernel(const int32_t *__restrict__ in, int64_t n, const int64_t *__restrict__ block_sums,
                  int32_t *__restrict__ out, int64_t *__restrict__ total, const int32_t *__restrict__ block_max,
                  int64_t *__restrict__ max_out /* nullable: written by the last block */) {
    __shared__ int64_t sm[kScanThreads / 32];
    __shared__ int32_t warp_tot[kScanThreads / 32];
    // prefix of all earlier CTAs (each CTA re-reduces the short block_sums array)
    int64_t acc = 0;
    for (int j = threadIdx.x; j < (int)blockIdx.x; j += kScanThreads) acc += block_sums[j];
    int64_t prefix = block_sum_i64(acc, sm);
    // local exclusive scan: thread t owns items [t*8, t*8+8) of the tile
    int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
    int32_t v[kScanItems];
    int32_t tsum = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        v[i] = (base + i < n) ? in[base + i] : 0;
        tsum += v[i];
    }
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int32_t incl = tsum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) warp_tot[w] = incl;
    __syncthreads();
    int32_t woff = 0;
#pragma unroll
    for (int i = 0; i < kScanThreads / 32; ++i)
        if (i < w) woff += warp_tot[i];
    int64_t run = prefix + woff + (incl - tsum);
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        if (base + i < n) out[base + i] = (int32_t)run;
        run += v[i];
    }
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == kScanThreads - 1) {
        *total = run;
        if (max_out) {
            int32_t mx = 0;
            for (int j = 0; j < (int)gridDim.x; ++j) mx = max(mx, block_max[j]);
            *max_out = mx;
        }
    }
}

// ----------------------------------------------------------------------------- radix sort
constexpr int kSortThreads = 256;
constexpr int kSortItems = 8;
constexpr int kSortTile = kSortThreads * kSortItems;  // 2048 pairs per CTA
constexpr int kRadix = 256;

__global__ void __launch_bounds__(kSortThreads)
sort_hist_kernel(const uint64_t *__restrict__ keys, int64_t n, int shift, uint32_t mask, int nblocks,
                 uint32_t *__restrict__ hist /*[256][nblocks]*/) {
    __shared__ uint32_t h[kRadix];
    h[threadIdx.x] = 0;
    __syncthreads();
    int64_t base = (int64_t)blockIdx.x * kSortTile;
#pragma unroll
    for (int i = 0; i < kSortItems; ++i) {
        int64_t j = base + (int64_t)i * kSortThreads + threadIdx.x;
        if (j < n) atomicAdd(&h[(uint32_t)(keys[j] >> shift) & mask], 1u);
    }
    __syncthreads();
    hist[(int64_t)threadIdx.x * nblocks + blockIdx.x] = h[threadIdx.x];
}

// one CTA per digit: exclusive scan of that digit's per-CTA counts (in place) + digit total
__global__ void __launch_bounds__(kSortThreads)
sort_scan_kernel(uint32_t *__restrict__ hist, int nblocks, uint32_t *__restrict__ digit_total) {
    __shared__ uint32_t warp_tot[kSortThreads / 32];
    __shared__ uint32_t carry_s;
    uint32_t *row = hist + (int64_t)blockIdx.x * nblocks;
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int start = 0; start < nblocks; start += kSortThreads) {
        int j = start + threadIdx.x;
        uint32_t v = (j < nblocks) ? row[j] : 0u;
        uint32_t incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) warp_tot[w] = incl;
        __syncthreads();
        uint32_t woff = 0, chunk = 0;
#pragma unroll
        for (int i = 0; i < kSortThreads / 32; ++i) {
            uint32_t t = warp_tot[i];
            if (i < w) woff += t;
            chunk += t;
        }
        uint32_t carry = carry_s;
        if (j < nblocks) row[j] = carry + woff + incl - v;
        __syncthreads();
        if (threadIdx.x == 0) carry_s = carry + chunk;
        __syncthreads();
    }
    if (threadIdx.x == 0) digit_total[blockIdx.x] = carry_s;
}

__global__ void __launch_bounds__(kSortThreads)
sort_scatter_kernel(const uint64_t *__restrict__ keys_in, const uint32_t *__restrict__ vals_in,
                    uint64_t *__restrict__ keys_out, uint32_t *__restrict__ vals_out, int64_t n,
                    int shift, uint32_t mask, int nblocks, const uint32_t *__restrict__ hist,
                    const uint32_t *__restrict__ digit_total) {
    __shared__ uint32_t warp_cnt[kSortThreads / 32][kRadix];
    __shared__ uint32_t digit_off[kRadix];
    __shared__ uint32_t warp_tot[kSortThreads / 32];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
#pragma unroll
    for (int i = 0; i < kSortThreads / 32; ++i) warp_cnt[i][tid] = 0;

    // global base of digit `tid`: exclusive scan of the 256 digit totals
    {
        uint32_t v = digit_total[tid];
        uint32_t incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) warp_tot[w] = incl;
        __syncthreads();
        uint32_t woff = 0;
#pragma unroll
        for (int i = 0; i < kSortThreads / 32; ++i)
            if (i < w) woff += warp_tot[i];
        digit_off[tid] = woff + incl - v + hist[(int64_t)tid * nblocks + blockIdx.x];
    }
    __syncthreads();

    // each warp owns a contiguous run of 256 pairs, item i of lane l at run + i*32 + l:
    // processing items in order i = 0..7 visits the run in memory order => stable.
    const int64_t wbase = (int64_t)blockIdx.x * kSortTile + (int64_t)w * (32 * kSortItems);
    uint64_t key[kSortItems];
    uint32_t val[kSortItems];
    uint32_t rank[kSortItems];
    const uint32_t lt_mask = (1u << lane) - 1u;
#pragma unroll
    for (int i = 0; i < kSortItems; ++i) {
        int64_t j = wbase + i * 32 + lane;
        bool valid = j < n;
        key[i] = valid ? keys_in[j] : 0ull;
        val[i] = valid ? vals_in[j] : 0u;
    }
#pragma unroll
    for (int i = 0; i < kSortItems; ++i) {
        int64_t j = wbase + i * 32 + lane;
        bool valid = j < n;
        uint32_t d = (uint32_t)(key[i] >> shift) & mask;
        uint32_t d_eff = valid ? d : (kRadix + lane);  // out-of-range lanes match nobody
        uint32_t peers = __match_any_sync(0xffffffffu, d_eff);
        uint32_t before = __popc(peers & lt_mask);
        uint32_t base = valid ? warp_cnt[w][d] : 0u;
        __syncwarp();
        if (valid && before == 0) warp_cnt[w][d] = base + __popc(peers);
        __syncwarp();
        rank[i] = base + before;
    }
    __syncthreads();
    // turn per-warp counts into exclusive offsets over warps (digit = tid)
    {
        uint32_t run = 0;
#pragma unroll
        for (int i = 0; i < kSortThreads / 32; ++i) {
            uint32_t t = warp_cnt[i][tid];
            warp_cnt[i][tid] = run;
            run += t;
        }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < kSortItems; ++i) {
        int64_t j = wbase + i * 32 + lane;
        if (j < n) {
            uint32_t d = (uint32_t)(key[i] >> shift) & mask;
            uint32_t pos = digit_off[d] + warp_cnt[w][d] + rank[i];
            keys_out[pos] = key[i];
            vals_out[pos] = val[i];
        }
    }
}

// ----------------------------------------------------------------------------- tile offsets
__global__ void __launch_bounds__(256)
tile_offsets_kernel(const int64_t *__restrict__ ids, int64_t n, int n_tiles, int tile_n_bits,
                    int64_t total, int32_t *__restrict__ offsets) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n) return;
    const int64_t tmask = (1LL << tile_n_bits) - 1;
    int64_t hi = ids[idx] >> 32;
    int64_t cur = (hi >> tile_n_bits) * n_tiles + (hi & tmask);
    if (idx == 0)
        for (int64_t t = 0; t <= cur && t < total; ++t) offsets[t] = 0;
    int64_t nxt = total;
    if (idx + 1 < n) {
        int64_t h2 = ids[idx + 1] >> 32;
        nxt = (h2 >> tile_n_bits) * n_tiles + (h2 & tmask);
    }
    for (int64_t t = cur + 1; t <= nxt && t < total; ++t) offsets[t] = (int32_t)(idx + 1);
}

// ----------------------------------------------------------------------------- tile-bucketed binning
// B200-first alternative to "emit + 6-pass global radix sort" (gsplat's structure): the sort key is
// (camera, tile | depth), and every (camera, tile) segment is small (hundreds to a few thousand
// entries), so
//   (1) count intersections per (camera, tile)      -> exclusive scan == isect_offsets directly,
//   (2) emit every intersection straight into its tile's segment (unordered, atomic cursor),
//   (3) one CTA per tile sorts its segment in SHARED MEMORY by the unique 64-bit key
//       (depth bits << 32 | flatten id) with a bitonic network and writes the final lists.
// ~20-28 B of HBM traffic per intersection instead of 6 passes x 36 B.  The result is bit-identical
// to the stable radix sort: ties in depth are ordered by flatten id, which is emission order.
constexpr int kTileSortMaxCap = 16384;  // segment capacity of the shared-memory sort (128 KB of keys; > 4096 keys
                                        // need the dynamic shared memory opt-in and leave one CTA per SM)

__device__ __forceinline__ void tile_rect_dev(float m2x, float m2y, int32_t radius, int tile_size, int tile_w,
                                              int tile_h, int &x0, int &y0, int &x1, int &y1) {
    // identical arithmetic to project_math.cuh::tile_rect (exact: divisions by the tile size, floor / ceil)
    const float ts = (float)tile_size;
    const float tr = __fdiv_rn((float)radius, ts);
    const float tx = __fdiv_rn(m2x, ts), ty = __fdiv_rn(m2y, ts);
    const float fw = (float)tile_w, fh = (float)tile_h;
    x0 = (int)fminf(fmaxf(floorf(__fsub_rn(tx, tr)), 0.f), fw);
    y0 = (int)fminf(fmaxf(floorf(__fsub_rn(ty, tr)), 0.f), fh);
    x1 = (int)fminf(fmaxf(ceilf(__fadd_rn(tx, tr)), 0.f), fw);
    y1 = (int)fminf(fmaxf(ceilf(__fadd_rn(ty, tr)), 0.f), fh);
}

// kLanesPerGauss threads share one Gaussian and stride over its tile rectangle: the tile counts per Gaussian are
// heavy-tailed (a few large Gaussians cover hundreds of tiles), and the emit loop is a chain of returning atomics
// (measured at c3: 1 lane 0.327 / 0.355 ms for count / emit, 4 lanes 0.184 / 0.248, 8 lanes the same, 16 lanes
// 0.268 / 0.314 -- with 4 the count runs at the L2 atomic rate, ~70 G atomics/s)
constexpr int kLanesPerGauss = 4;

__global__ void __launch_bounds__(256)
tile_count_kernel(const float *__restrict__ means2d, const int32_t *__restrict__ radii, int C, int G, int tile_size,
                  int tile_w, int tile_h, int32_t *__restrict__ tile_counts) {
    const int64_t tg = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t idx = tg / kLanesPerGauss;
    const int sub = (int)(tg - idx * kLanesPerGauss);
    if (idx >= (int64_t)C * G) return;
    const int32_t r = radii[idx];
    if (r <= 0) return;
    const float2 m2 = __ldg(reinterpret_cast<const float2 *>(means2d) + idx);
    int x0, y0, x1, y1;
    tile_rect_dev(m2.x, m2.y, r, tile_size, tile_w, tile_h, x0, y0, x1, y1);
    int32_t *cnt = tile_counts + (idx / G) * (int64_t)tile_w * tile_h;
    const int wx = x1 - x0, n = wx * (y1 - y0);
    for (int q = sub; q < n; q += kLanesPerGauss) {
        const int i = q / wx, j = q - i * wx;
        atomicAdd(cnt + (y0 + i) * tile_w + x0 + j, 1);
    }
}

__global__ void __launch_bounds__(256)
bucket_emit_kernel(const float *__restrict__ means2d, const int32_t *__restrict__ radii,
                   const float *__restrict__ depths, int C, int G, int tile_size, int tile_w, int tile_h,
                   const int32_t *__restrict__ tile_offsets, int32_t *__restrict__ cursors,
                   uint64_t *__restrict__ bucket_keys, int64_t capacity, int bucket_stride) {
    // bucket_stride > 0: fixed-stride buckets (tile t owns bucket_keys[t * bucket_stride ...], no offsets needed: the
    // count pass and the scan before the emit are skipped; `cursors` doubles as the per-tile counts)
    const int64_t tg = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t idx = tg / kLanesPerGauss;
    const int sub = (int)(tg - idx * kLanesPerGauss);
    if (idx >= (int64_t)C * G) return;
    const int32_t r = radii[idx];
    if (r <= 0) return;
    const float2 m2 = __ldg(reinterpret_cast<const float2 *>(means2d) + idx);
    int x0, y0, x1, y1;
    tile_rect_dev(m2.x, m2.y, r, tile_size, tile_w, tile_h, x0, y0, x1, y1);
    const int64_t cbase = (idx / G) * (int64_t)tile_w * tile_h;
    const uint64_t key = ((uint64_t)(uint32_t)__float_as_int(depths[idx]) << 32) | (uint32_t)idx;
    const int wx = x1 - x0, n = wx * (y1 - y0);
    for (int q = sub; q < n; q += kLanesPerGauss) {
        const int i = q / wx, j = q - i * wx;
        const int64_t t = cbase + (y0 + i) * tile_w + x0 + j;
        const int32_t slot = atomicAdd(cursors + t, 1);
        if (bucket_stride > 0) {
            if (slot < bucket_stride) bucket_keys[t * bucket_stride + slot] = key;
        } else {
            const int32_t pos = tile_offsets[t] + slot;
            if (pos < capacity) bucket_keys[pos] = key;  // beyond the caller's capacity: dropped, the overflow is flagged by the sort
        }
    }
}

// ----------------------------------------------------------------------------- packed intersection records
// Input of the slab blend kernels (slab.cuh): per intersection one 32-byte record with everything the per-pixel
// evaluation needs, in the tile's depth order, records with an empty reach mask dropped.  One CTA per (camera, tile).
struct PackArgs {
    const float *means2d, *conics, *opacities, *depths;  // [C,G,2], [C,G,3], [G], [C,G] (depths may be null)
    int G, tile_w, tile_size;
    float4 *recs;         // [n_isects][2]
    int32_t *rec_counts;  // [C * tiles]
};

template <typename IdAt>
__device__ __forceinline__ void pack_segment(const PackArgs &p, IdAt id_at, int n, int32_t start, int64_t seg,
                                             int n_tiles, int32_t *s_warp /*[blockDim / 32]*/) {
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, n_warps = blockDim.x >> 5;
    const int c = (int)(seg / n_tiles), tile = (int)(seg - (int64_t)c * n_tiles);
    const int ty = tile / p.tile_w, tx = tile - ty * p.tile_w;
    int run = 0;
    for (int base = 0; base < n; base += blockDim.x) {
        const int i = base + tid;
        float4 r0 = make_float4(0.f, 0.f, 0.f, 0.f), r1 = r0;
        bool keep = false;
        if (i < n) {
            const int32_t g = id_at(i);
            const int32_t gl = g - c * p.G;
            const float2 xy = __ldg(reinterpret_cast<const float2 *>(p.means2d) + g);
            const float op = __ldg(p.opacities + gl);
            const float *cp = p.conics + 3LL * g;
            const float ca = __ldg(cp), cb = __ldg(cp + 1), cc = __ldg(cp + 2);
            const float L = __log2f(op);
            const uint32_t mask = reach_mask_of(xy.x, xy.y, L, ca, cb, cc, tx * p.tile_size, ty * p.tile_size);
            keep = mask != 0u;
            r0 = make_float4(xy.x, xy.y, L, __uint_as_float((uint32_t)gl | (mask << 24)));
            r1 = make_float4(-0.5f * kLog2e * ca, -kLog2e * cb, -0.5f * kLog2e * cc, p.depths ? __ldg(p.depths + g) : 0.f);
        }
        const uint32_t bal = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) s_warp[w] = __popc(bal);
        __syncthreads();
        int woff = 0, total = 0;
        for (int k = 0; k < n_warps; ++k) {
            const int t = s_warp[k];
            woff += k < w ? t : 0;
            total += t;
        }
        if (keep) {
            float4 *dst = p.recs + 2 * ((int64_t)start + run + woff + __popc(bal & ((1u << lane) - 1u)));
            dst[0] = r0;
            dst[1] = r1;
        }
        run += total;
        __syncthreads();
    }
    if (tid == 0) p.rec_counts[seg] = run;
}

// The run of segment `seg` in the sorted lists and whether the tile sort could take it (capacity mode, see
// tile_sort_kernel): shared by the sort and by the stand-alone packing pass that follows it for long lists.
__device__ __forceinline__ bool segment_fits(const int32_t *__restrict__ tile_offsets, int64_t n_isects,
                                             const int64_t *__restrict__ n_isects_dev, int64_t capacity, int sort_capacity,
                                             int bucket_stride, int64_t n_segments, int64_t seg, int64_t &start64,
                                             int64_t &end64) {
    if (n_isects_dev) n_isects = *n_isects_dev;
    start64 = tile_offsets[seg];
    end64 = (seg == n_segments - 1) ? n_isects : (int64_t)tile_offsets[seg + 1];
    return end64 <= capacity && end64 - start64 <= sort_capacity && (bucket_stride == 0 || end64 - start64 <= bucket_stride);
}

__global__ void __launch_bounds__(256)
isect_pack_kernel(PackArgs p, const int32_t *__restrict__ tile_offsets, const int32_t *__restrict__ flatten_ids,
                  int64_t n_isects, int64_t n_segments, int n_tiles) {
    __shared__ int32_t s_warp[8];
    const int64_t seg = blockIdx.x;
    const int32_t start = tile_offsets[seg];
    const int32_t end = (seg == n_segments - 1) ? (int32_t)n_isects : tile_offsets[seg + 1];
    pack_segment(p, [&](int i) { return __ldg(flatten_ids + start + i); }, end - start, start, seg, n_tiles, s_warp);
}

// packing pass after a tile sort that did not pack (long lists): same segment logic as tile_sort_kernel, but at eight
// CTAs per SM -- the gathers of a batch are latency-bound and the sort's shared-memory footprint leaves 1-3 CTAs
__global__ void __launch_bounds__(256)
isect_pack_after_sort_kernel(PackArgs p, const int32_t *__restrict__ tile_offsets, const int32_t *__restrict__ flatten_ids,
                             int64_t n_isects, const int64_t *__restrict__ n_isects_dev, int64_t capacity,
                             int sort_capacity, int bucket_stride, int64_t n_segments, int n_tiles) {
    __shared__ int32_t s_warp[8];
    const int64_t seg = blockIdx.x;
    int64_t start64, end64;
    const bool fits = segment_fits(tile_offsets, n_isects, n_isects_dev, capacity, sort_capacity, bucket_stride, n_segments,
                                   seg, start64, end64);
    const int32_t start = (int32_t)start64;
    const int n = fits ? (int)(end64 - start64) : 0;
    pack_segment(p, [&](int i) { return __ldg(flatten_ids + start + i); }, n, start, seg, n_tiles, s_warp);
}

__global__ void __launch_bounds__(256, 8)  // 32 registers: eight CTAs per SM for the short lists of the benchmark scene
tile_sort_kernel(const uint64_t *__restrict__ bucket_keys, const int32_t *__restrict__ tile_offsets, int64_t n_isects,
                 const int64_t *__restrict__ n_isects_dev, int64_t capacity, int sort_capacity,
                 int64_t *__restrict__ overflow, int64_t n_segments, int n_tiles, int tile_n_bits,
                 int64_t *__restrict__ isect_ids, int32_t *__restrict__ flatten_ids, PackArgs pack,
                 const int32_t *__restrict__ bucket_counts, int bucket_stride) {
    extern __shared__ uint64_t s_keys[];
    __shared__ int32_t s_warp[8];
    const int64_t seg = blockIdx.x;
    // capacity mode (n_isects_dev != null): the intersection count lives on the device, the buffers hold `capacity`
    // entries and the shared-memory sort `sort_capacity` keys; a tile that does not fit is dropped (no records) and
    // *overflow is raised for the host to see later -- no device -> host sync inside the step
    int64_t start64, end64;
    const bool fits = segment_fits(tile_offsets, n_isects, n_isects_dev, capacity, sort_capacity, bucket_stride, n_segments,
                                   seg, start64, end64);
    // fixed-stride buckets: the unsorted keys of the segment sit at bucket_keys[seg * bucket_stride ...]
    const uint64_t *src_keys = bucket_stride > 0 ? bucket_keys + seg * (int64_t)bucket_stride : bucket_keys + start64;
    (void)bucket_counts;
    if (!fits && overflow && threadIdx.x == 0) *overflow = 1;
    const int32_t start = (int32_t)start64;
    const int n = fits ? (int)(end64 - start64) : 0;
    if (n <= 0) {
        if (pack.recs && threadIdx.x == 0) pack.rec_counts[seg] = 0;
        return;
    }
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
    const int64_t cam = seg / n_tiles, tile = seg - cam * n_tiles;
    const int64_t hi_bits = (cam << (32 + tile_n_bits)) | (tile << 32);
    auto write_out = [&](auto key_at) {
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const uint64_t key = key_at(i);
            flatten_ids[start + i] = (int32_t)(uint32_t)key;
            isect_ids[start + i] = hi_bits | (int64_t)(key >> 32);
        }
        // packed records of the tile for the slab blend kernels, straight from the sorted keys in shared memory
        if (pack.recs) pack_segment(pack, [&](int i) { return (int32_t)(uint32_t)key_at(i); }, n, start, seg, n_tiles, s_warp);
    };
    auto kp = [](int k) { return k ^ ((k & 16) - ((k & 16) >> 4)); };  // slot of key k (see below)

    // ---- long lists (stress configurations: thousands of entries per tile): 64-key runs sorted by one warp each, then
    // log2(n / 64) merge-path levels between two halves of the buffer.  The bitonic network costs n log^2 n / 4
    // compare-exchanges (78 stages at 4096 keys), the merges n log n element moves: measured at c5 x2 (170 M
    // intersections, 3640 per tile on average) the network was 13.2 of the step's 26.1 ms.
    const int n64 = (n + 63) & ~63;
    // every thread produces E consecutive outputs per merge level; E is a power of two, so a range never straddles a pair
    int eshift = 3;
    while ((blockDim.x << eshift) < n64) ++eshift;
    const int E = 1 << eshift;
    // merged runs are stored with one slot of padding per E keys: thread t writes from slot t (E + 1), an odd stride, so
    // the 32 lanes of a warp hit 16 distinct bank pairs (unpadded they would all hit the same one)
    auto pm = [&](int i) { return i + (i >> eshift); };
    const int nbuf = pm(n64) + 1;
    if (n > 1024 && 2 * nbuf <= sort_capacity) {
        uint64_t *buf0 = s_keys, *buf1 = s_keys + nbuf;
        for (int i = threadIdx.x; i < n64; i += blockDim.x) buf0[kp(i)] = i < n ? src_keys[i] : ~0ull;
        __syncthreads();
        for (int base = w * 64; base < n64; base += n_warps * 64) {  // ascending 64-key runs, warp-local
            for (int k = 2; k <= 64; k <<= 1)
                for (int jj = k >> 1; jj > 0; jj >>= 1) {
                    const int lo = base + (((lane & ~(jj - 1)) << 1) | (lane & (jj - 1)));
                    const int plo = kp(lo), phi = kp(lo | jj);
                    const uint64_t a = buf0[plo], b = buf0[phi];
                    const bool asc = k == 64 || (lo & k) == 0;
                    if ((a > b) == asc) {
                        buf0[plo] = b;
                        buf0[phi] = a;
                    }
                    __syncwarp();
                }
        }
        __syncthreads();
        const uint64_t *src = buf0;
        uint64_t *dst = buf1;
        bool swz = true;  // the runs of the first level sit in the mirrored layout of the warp-local sort
        for (int run = 64; run < n64; run <<= 1) {
            const int o0 = threadIdx.x * E;
            if (o0 < n64) {
                const int pbase = o0 & ~(2 * run - 1);
                const int a_len = min(run, n64 - pbase), b_len = min(run, n64 - pbase - a_len);
                const int a0 = pbase, b0 = pbase + a_len, d = o0 - pbase;
                auto at = [&](int i) { return src[swz ? kp(i) : pm(i)]; };
                // merge path: how many of the first d outputs come from run A (keys are unique)
                int lo = max(0, d - b_len), hi = min(d, a_len);
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    if (at(a0 + mid) <= at(b0 + d - 1 - mid)) lo = mid + 1;
                    else hi = mid;
                }
                int i = lo, j = d - lo;
                uint64_t a = i < a_len ? at(a0 + i) : ~0ull, b = j < b_len ? at(b0 + j) : ~0ull;
                const int cnt = min(E, a_len + b_len - d);
                for (int t = 0; t < cnt; ++t) {
                    const bool take_a = a <= b;
                    dst[pm(o0 + t)] = take_a ? a : b;
                    if (take_a) {
                        ++i;
                        a = i < a_len ? at(a0 + i) : ~0ull;
                    } else {
                        ++j;
                        b = j < b_len ? at(b0 + j) : ~0ull;
                    }
                }
            }
            __syncthreads();
            const uint64_t *t2 = src;
            src = dst;
            dst = const_cast<uint64_t *>(t2);
            swz = false;
        }
        const uint64_t *sorted = src;
        const bool sorted_swz = swz;  // n64 == 64 cannot happen here (n > 1024), kept for clarity
        write_out([&](int i) { return sorted[sorted_swz ? kp(i) : pm(i)]; });
        return;
    }

    int n_pad = 1;
    while (n_pad < n) n_pad <<= 1;
    // Key k sits at slot kp(k): within every second 16-key row (128 bytes = all 32 banks) the columns are mirrored.
    // A stage with partner distance jj < 16 touches the same 8 of 16 columns in each of the 4 rows of a 64-key chunk
    // -- 4 wavefronts for 32 x 8 bytes where 2 are needed; mirrored odd rows use the complementary columns (ncu before:
    // 38 % of the kernel's shared-memory wavefronts were bank conflicts at 93 % LSU data-pipe utilisation).
    for (int i = threadIdx.x; i < n_pad; i += blockDim.x) s_keys[kp(i)] = i < n ? src_keys[i] : ~0ull;
    __syncthreads();
    // Bitonic network.  Stages with partner distance j <= 32 only exchange within aligned 64-key chunks: chunk q is
    // owned by warp q % 8 for the whole sort, so those stages need a warp barrier only.  Block barriers remain
    // around the stages with j >= 64 (6 of the 45 stages at 512 keys).
    auto cmpx = [&](int lo, int hi, int k) {  // logical positions lo < hi
        const int plo = kp(lo), phi = kp(hi);
        const uint64_t a = s_keys[plo], b = s_keys[phi];
        const bool asc = (lo & k) == 0;
        if ((a > b) == asc) {
            s_keys[plo] = b;
            s_keys[phi] = a;
        }
    };
    for (int k = 2; k <= n_pad; k <<= 1) {
        int j = k >> 1;
        if (j >= 64) {
            __syncthreads();  // the warp-local tails of the previous k are complete
            for (; j >= 64; j >>= 1) {
                for (int i = threadIdx.x; i < (n_pad >> 1); i += blockDim.x) {
                    const int lo = ((i & ~(j - 1)) << 1) | (i & (j - 1));
                    cmpx(lo, lo | j, k);
                }
                __syncthreads();
            }
        }
        for (int base = w * 64; base < n_pad; base += n_warps * 64) {
            for (int jj = j; jj > 0; jj >>= 1) {
                const int lo = base + (((lane & ~(jj - 1)) << 1) | (lane & (jj - 1)));
                const int hi = lo | jj;
                if (hi < n_pad) cmpx(lo, hi, k);
                __syncwarp();
            }
        }
    }
    __syncthreads();
    write_out([&](int i) { return s_keys[kp(i)]; });
}

}  // namespace d4

using namespace d4;

extern "C" int d4_tile_sort_capacity(void) { return kTileSortMaxCap; }

extern "C" int d4_tile_count(const float *means2d, const int32_t *radii, int C, int G, int tile_size, int tile_w,
                             int tile_h, int32_t *tile_counts, d4_stream_t stream) {
    D4_CHECK_ARG(C >= 1 && G >= 0 && tile_counts, "d4_tile_count: bad arguments");
    if (G == 0) return 0;
    D4_CHECK_ARG(means2d && radii, "d4_tile_count: null pointer");
    tile_count_kernel<<<cdiv((int64_t)C * G * kLanesPerGauss, 256), 256, 0, as_stream(stream)>>>(
        means2d, radii, C, G, tile_size, tile_w, tile_h, tile_counts);
    D4_CHECK_LAUNCH("d4_tile_count");
    return 0;
}

extern "C" int d4_bucket_emit(const float *means2d, const int32_t *radii, const float *depths, int C, int G,
                              int tile_size, int tile_w, int tile_h, const int32_t *tile_offsets, int32_t *cursors,
                              uint64_t *bucket_keys, int64_t capacity, int bucket_stride, d4_stream_t stream) {
    D4_CHECK_ARG(capacity >= 0 && capacity < (1LL << 31) && bucket_stride >= 0, "d4_bucket_emit: bad capacity");
    D4_CHECK_ARG(C >= 1 && G >= 0 && (int64_t)C * G < (1LL << 32), "d4_bucket_emit: bad sizes");
    if (G == 0) return 0;
    D4_CHECK_ARG(means2d && radii && depths && (tile_offsets || bucket_stride > 0) && cursors && bucket_keys, "d4_bucket_emit: null pointer");
    bucket_emit_kernel<<<cdiv((int64_t)C * G * kLanesPerGauss, 256), 256, 0, as_stream(stream)>>>(
        means2d, radii, depths, C, G, tile_size, tile_w, tile_h, tile_offsets, cursors, bucket_keys, capacity, bucket_stride);
    D4_CHECK_LAUNCH("d4_bucket_emit");
    return 0;
}

static int check_pack(const char *name, const float *means2d, const float *conics, const float *opacities, int G,
                      int tile_size, float4 *recs, int32_t *rec_counts) {
    D4_CHECK_ARG(means2d && conics && opacities && recs && rec_counts, "%s: null pointer", name);
    D4_CHECK_ARG(G < (1 << 24), "%s: packed records hold 24-bit Gaussian ids (G = %d)", name, G);
    D4_CHECK_ARG(tile_size == kTile, "%s: only tile_size 16 is built", name);
    D4_CHECK_ARG(((uintptr_t)recs & 15) == 0 && ((uintptr_t)means2d & 7) == 0, "%s: misaligned pointer", name);
    return 0;
}

extern "C" int d4_isect_pack(const float *means2d, const float *conics, const float *opacities, const float *depths,
                             int C, int G, int tile_size, int tile_w, int tile_h, const int32_t *tile_offsets,
                             const int32_t *flatten_ids, int64_t n_isects, void *recs, int32_t *rec_counts,
                             d4_stream_t stream) {
    D4_CHECK_ARG(C >= 1 && tile_w >= 1 && tile_h >= 1 && n_isects >= 0 && tile_offsets && rec_counts, "d4_isect_pack: bad arguments");
    if (n_isects == 0) {  // nothing to pack (the Gaussian arrays may be empty / null)
        cudaMemsetAsync(rec_counts, 0, sizeof(int32_t) * (size_t)C * tile_w * tile_h, as_stream(stream));
        return 0;
    }
    if (int rc = check_pack("d4_isect_pack", means2d, conics, opacities, G, tile_size, (float4 *)recs, rec_counts)) return rc;
    D4_CHECK_ARG(flatten_ids || n_isects == 0, "d4_isect_pack: null pointer");
    const int64_t n_seg = (int64_t)C * tile_w * tile_h;
    PackArgs p{means2d, conics, opacities, depths, G, tile_w, tile_size, (float4 *)recs, rec_counts};
    isect_pack_kernel<<<(unsigned)n_seg, 256, 0, as_stream(stream)>>>(p, tile_offsets, flatten_ids, n_isects, n_seg,
                                                                      tile_w * tile_h);
    D4_CHECK_LAUNCH("d4_isect_pack");
    return 0;
}

extern "C" size_t d4_slab_hit_words(int64_t n_isects, int64_t n_segments) { return slab_hit_words(n_isects, n_segments); }

static int launch_tile_sort(const char *name, const uint64_t *bucket_keys, const int32_t *tile_offsets, int64_t n_isects,
                            const int64_t *n_isects_dev, int64_t capacity, int sort_capacity, int64_t *overflow, int C,
                            int tile_w, int tile_h, int64_t *isect_ids, int32_t *flatten_ids, const PackArgs &p,
                            cudaStream_t st, int bucket_stride = 0) {
    int n_pad = 1;
    while (n_pad < sort_capacity) n_pad <<= 1;
    // long lists are merge-sorted between two halves of the buffer (tile_sort_kernel): twice the keys while three CTAs
    // still share an SM (64 KB each) -- the packing tail of the kernel is latency-bound and needs the occupancy (measured
    // at c5 x2: doubling 64 KB to 128 KB made the kernel slower, 13.2 -> 16.0 ms, although every tile took the merge path)
    if (n_pad > 2048 && n_pad <= 4096) n_pad *= 2;
    const size_t smem = sizeof(uint64_t) * (size_t)n_pad;
    if (smem > 48 * 1024 &&
        cudaFuncSetAttribute(tile_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
        set_error("%s: cannot reserve %zu bytes of shared memory for the tile sort", name, smem);
        return 1;
    }
    const int64_t n_seg = (int64_t)C * tile_w * tile_h;
    // long lists: the records are packed by a second kernel at full occupancy (c5 x2: 9.3 ms fused)
    const bool split_pack = p.recs != nullptr && n_pad > 2048;
    PackArgs p_sort = p;
    if (split_pack) p_sort.recs = nullptr;
    tile_sort_kernel<<<(unsigned)n_seg, 256, smem, st>>>(bucket_keys, tile_offsets, n_isects, n_isects_dev, capacity, n_pad,
                                                        overflow, n_seg, tile_w * tile_h, d4_tile_n_bits(tile_w * tile_h),
                                                        isect_ids, flatten_ids, p_sort, nullptr, bucket_stride);
    D4_CHECK_LAUNCH(name);
    if (split_pack) {
        isect_pack_after_sort_kernel<<<(unsigned)n_seg, 256, 0, st>>>(p, tile_offsets, flatten_ids, n_isects, n_isects_dev,
                                                                     capacity, n_pad, bucket_stride, n_seg, tile_w * tile_h);
        D4_CHECK_LAUNCH(name);
    }
    return 0;
}

extern "C" int d4_tile_sort_pack(const uint64_t *bucket_keys, const int32_t *tile_offsets, int64_t n_isects, int C,
                                 int tile_w, int tile_h, int max_count, int64_t *isect_ids, int32_t *flatten_ids,
                                 const float *means2d, const float *conics, const float *opacities,
                                 const float *depths, int G, int tile_size, void *recs, int32_t *rec_counts,
                                 d4_stream_t stream) {
    D4_CHECK_ARG(C >= 1 && tile_w >= 1 && tile_h >= 1 && n_isects >= 0, "d4_tile_sort_pack: bad arguments");
    D4_CHECK_ARG(max_count <= kTileSortMaxCap, "d4_tile_sort_pack: a tile holds %d intersections, capacity is %d "
                                               "(use d4_isect_emit + d4_sort_pairs_u64 + d4_isect_pack)", max_count, kTileSortMaxCap);
    if (int rc = check_pack("d4_tile_sort_pack", means2d, conics, opacities, G, tile_size, (float4 *)recs, rec_counts)) return rc;
    D4_CHECK_ARG(tile_offsets, "d4_tile_sort_pack: null pointer");
    const int64_t n_seg = (int64_t)C * tile_w * tile_h;
    if (n_isects == 0) {
        cudaMemsetAsync(rec_counts, 0, sizeof(int32_t) * n_seg, as_stream(stream));
        return 0;
    }
    D4_CHECK_ARG(bucket_keys && isect_ids && flatten_ids, "d4_tile_sort_pack: null pointer");
    PackArgs p{means2d, conics, opacities, depths, G, tile_w, tile_size, (float4 *)recs, rec_counts};
    return launch_tile_sort("d4_tile_sort_pack", bucket_keys, tile_offsets, n_isects, nullptr, n_isects, max_count, nullptr, C,
                            tile_w, tile_h, isect_ids, flatten_ids, p, as_stream(stream));
}

extern "C" int d4_tile_sort_pack_cap(const uint64_t *bucket_keys, const int32_t *tile_offsets, const int64_t *bin_stats,
                                     int64_t capacity, int sort_capacity, int C, int tile_w, int tile_h,
                                     int64_t *isect_ids, int32_t *flatten_ids, const float *means2d, const float *conics,
                                     const float *opacities, const float *depths, int G, int tile_size, void *recs,
                                     int32_t *rec_counts, int64_t *overflow, int bucket_stride, d4_stream_t stream) {
    D4_CHECK_ARG(C >= 1 && tile_w >= 1 && tile_h >= 1 && capacity >= 1 && capacity < (1LL << 31) && sort_capacity >= 1 &&
                     sort_capacity <= kTileSortMaxCap,
                 "d4_tile_sort_pack_cap: bad arguments (sort capacity <= %d)", kTileSortMaxCap);
    if (int rc = check_pack("d4_tile_sort_pack_cap", means2d, conics, opacities, G, tile_size, (float4 *)recs, rec_counts)) return rc;
    D4_CHECK_ARG(bucket_keys && tile_offsets && bin_stats && isect_ids && flatten_ids && overflow,
                 "d4_tile_sort_pack_cap: null pointer");
    PackArgs p{means2d, conics, opacities, depths, G, tile_w, tile_size, (float4 *)recs, rec_counts};
    D4_CHECK_ARG(bucket_stride >= 0, "d4_tile_sort_pack_cap: bad bucket stride");
    return launch_tile_sort("d4_tile_sort_pack_cap", bucket_keys, tile_offsets, 0, bin_stats, capacity, sort_capacity, overflow,
                            C, tile_w, tile_h, isect_ids, flatten_ids, p, as_stream(stream), bucket_stride);
}

extern "C" int d4_tile_sort(const uint64_t *bucket_keys, const int32_t *tile_offsets, int64_t n_isects, int C,
                            int tile_w, int tile_h, int max_count, int64_t *isect_ids, int32_t *flatten_ids,
                            d4_stream_t stream) {
    D4_CHECK_ARG(C >= 1 && tile_w >= 1 && tile_h >= 1 && n_isects >= 0, "d4_tile_sort: bad arguments");
    D4_CHECK_ARG(max_count <= kTileSortMaxCap, "d4_tile_sort: a tile holds %d intersections, capacity is %d "
                                               "(use d4_isect_emit + d4_sort_pairs_u64)", max_count, kTileSortMaxCap);
    if (n_isects == 0) return 0;
    D4_CHECK_ARG(bucket_keys && tile_offsets && isect_ids && flatten_ids, "d4_tile_sort: null pointer");
    return launch_tile_sort("d4_tile_sort", bucket_keys, tile_offsets, n_isects, nullptr, n_isects, max_count, nullptr, C,
                            tile_w, tile_h, isect_ids, flatten_ids, PackArgs{}, as_stream(stream));
}

extern "C" size_t d4_scan_workspace_bytes(int64_t n) {
    const size_t nb = (size_t)cdiv(n > 0 ? n : 1, kScanTile);
    return sizeof(int64_t) * (nb + 1) + sizeof(int32_t) * (nb + 2);  // block sums + block maxima
}

static int scan_i32(const char *name, const int32_t *in, int64_t n, int32_t *out_exclusive, int64_t *total, int64_t *max_out,
                    void *workspace, size_t workspace_bytes, cudaStream_t st) {
    D4_CHECK_ARG(n >= 0 && total, "%s: bad arguments", name);
    if (n == 0) {
        cudaMemsetAsync(total, 0, sizeof(int64_t), st);
        if (max_out) cudaMemsetAsync(max_out, 0, sizeof(int64_t), st);
        return 0;
    }
    D4_CHECK_ARG(in && out_exclusive && workspace && workspace_bytes >= d4_scan_workspace_bytes(n),
                 "%s: null pointer or workspace too small", name);
    int nb = cdiv(n, kScanTile);
    int64_t *bs = reinterpret_cast<int64_t *>(workspace);
    int32_t *bm = reinterpret_cast<int32_t *>(bs + nb + 1);
    scan_block_sums_kernel<<<nb, kScanThreads, 0, st>>>(in, n, bs, max_out ? bm : nullptr);
    scan_apply_kernel<<<nb, kScanThreads, 0, st>>>(in, n, bs, out_exclusive, total, bm, max_out);
    D4_CHECK_LAUNCH(name);
    return 0;
}

extern "C" int d4_exclusive_scan_i32(const int32_t *in, int64_t n, int32_t *out_exclusive, int64_t *total,
                                     void *workspace, size_t workspace_bytes, d4_stream_t stream) {
    return scan_i32("d4_exclusive_scan_i32", in, n, out_exclusive, total, nullptr, workspace, workspace_bytes, as_stream(stream));
}

extern "C" int d4_scan_counts(const int32_t *counts, int64_t n, int32_t *offsets, int64_t *stats, void *workspace,
                              size_t workspace_bytes, d4_stream_t stream) {
    D4_CHECK_ARG(stats, "d4_scan_counts: null pointer");
    return scan_i32("d4_scan_counts", counts, n, offsets, stats, stats + 1, workspace, workspace_bytes, as_stream(stream));
}

extern "C" int d4_tile_sort_capacity_max(void) { return kTileSortMaxCap; }

extern "C" size_t d4_sort_workspace_bytes(int64_t n) {
    int nb = cdiv(n > 0 ? n : 1, kSortTile);
    return sizeof(uint32_t) * ((size_t)kRadix * nb + kRadix);
}

extern "C" int d4_sort_pairs_u64(uint64_t *keys_a, uint32_t *vals_a, uint64_t *keys_b, uint32_t *vals_b,
                                 int64_t n, int begin_bit, int end_bit, void *workspace,
                                 size_t workspace_bytes, int *result_in_b, d4_stream_t stream) {
    D4_CHECK_ARG(n >= 0 && n < (1LL << 31) && begin_bit >= 0 && end_bit <= 64 && begin_bit <= end_bit && result_in_b,
                 "d4_sort_pairs_u64: bad arguments");
    *result_in_b = 0;
    if (n <= 1 || begin_bit == end_bit) return 0;
    D4_CHECK_ARG(keys_a && vals_a && keys_b && vals_b && workspace && workspace_bytes >= d4_sort_workspace_bytes(n),
                 "d4_sort_pairs_u64: null pointer or workspace too small");
    int nb = cdiv(n, kSortTile);
    uint32_t *hist = reinterpret_cast<uint32_t *>(workspace);
    uint32_t *dtot = hist + (size_t)kRadix * nb;
    uint64_t *ki = keys_a, *ko = keys_b;
    uint32_t *vi = vals_a, *vo = vals_b;
    int flips = 0;
    for (int shift = begin_bit; shift < end_bit; shift += 8) {
        int bits = end_bit - shift < 8 ? end_bit - shift : 8;
        uint32_t mask = (1u << bits) - 1u;
        sort_hist_kernel<<<nb, kSortThreads, 0, as_stream(stream)>>>(ki, n, shift, mask, nb, hist);
        sort_scan_kernel<<<kRadix, kSortThreads, 0, as_stream(stream)>>>(hist, nb, dtot);
        sort_scatter_kernel<<<nb, kSortThreads, 0, as_stream(stream)>>>(ki, vi, ko, vo, n, shift, mask, nb, hist, dtot);
        uint64_t *tk = ki; ki = ko; ko = tk;
        uint32_t *tv = vi; vi = vo; vo = tv;
        ++flips;
    }
    D4_CHECK_LAUNCH("d4_sort_pairs_u64");
    *result_in_b = flips & 1;
    return 0;
}

extern "C" int d4_tile_offsets(const int64_t *isect_ids_sorted, int64_t n_isects, int C, int tile_w, int tile_h,
                               int32_t *offsets, d4_stream_t stream) {
    D4_CHECK_ARG(offsets && C >= 1 && tile_w >= 1 && tile_h >= 1 && n_isects >= 0, "d4_tile_offsets: bad arguments");
    int64_t total = (int64_t)C * tile_w * tile_h;
    if (n_isects == 0) {
        cudaMemsetAsync(offsets, 0, sizeof(int32_t) * total, as_stream(stream));
        return 0;
    }
    D4_CHECK_ARG(isect_ids_sorted, "d4_tile_offsets: null pointer");
    int tb = d4_tile_n_bits(tile_w * tile_h);
    tile_offsets_kernel<<<cdiv(n_isects, 256), 256, 0, as_stream(stream)>>>(isect_ids_sorted, n_isects,
                                                                           tile_w * tile_h, tb, total, offsets);
    D4_CHECK_LAUNCH("d4_tile_offsets");
    return 0;
}
