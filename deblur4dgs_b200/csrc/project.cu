// project.cu -- rows a8 / a9(emit) / a12 of SURVEY.md section 8:
//   d4_project_fwd  : gsplat fully_fused_projection fwd + isect_tiles pass 1
//   d4_project_bwd  : gsplat fully_fused_projection bwd
//   d4_isect_emit   : gsplat isect_tiles pass 2 (key/value emission)
// Compiled with -fmad=false (see project_math.cuh): these kernels are HBM-bound
// streaming passes over the Gaussian SoA (68 B in+out per Gaussian forward), so
// the un-fused multiplies cost nothing measurable.
#include "common.cuh"
#include "project_math.cuh"

namespace d4 {

constexpr int kProjThreads = 256;

__global__ void __launch_bounds__(kProjThreads)
project_fwd_kernel(const float *__restrict__ means, int64_t means_cs, const float *__restrict__ quats,
                   int64_t quats_cs, const float *__restrict__ scales,
                   const float *__restrict__ viewmats, int64_t vm_cs, const float *__restrict__ Ks,
                   int64_t k_cs, int C, int G, int width, int height, float eps2d, float near_plane,
                   float far_plane, float radius_clip, int tile_size, int tile_w, int tile_h,
                   int32_t *__restrict__ radii, float *__restrict__ means2d,
                   float *__restrict__ depths, float *__restrict__ conics,
                   int32_t *__restrict__ tiles_per_gauss, const int32_t *__restrict__ cam_row0, int window_height) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)C * G) return;
    int c = (int)(idx / G);
    int g = (int)(idx - (int64_t)c * G);
    const float *V = viewmats + c * vm_cs;
    const float *K = Ks + c * k_cs;
    float Vl[12], Kl[6];
#pragma unroll
    for (int i = 0; i < 12; ++i) Vl[i] = __ldg(V + i);
#pragma unroll
    for (int i = 0; i < 6; ++i) Kl[i] = __ldg(K + i);
    const float *mp = means + c * means_cs + 3LL * g;
    float m[3] = {__ldg(mp), __ldg(mp + 1), __ldg(mp + 2)};
    float4 q4 = __ldg(reinterpret_cast<const float4 *>(quats + c * quats_cs) + g);
    float q[4] = {q4.x, q4.y, q4.z, q4.w};
    const float *sp = scales + 3LL * g;
    float s[3] = {__ldg(sp), __ldg(sp + 1), __ldg(sp + 2)};
    ProjOut o = project_one(m, q, s, Vl, Kl, width, height, eps2d, near_plane, far_plane, radius_clip);
    if (cam_row0) {
        // row window of this camera (multi-GPU tile-row bands): the projection itself is that of the FULL image
        // (same tan-fov clamp, same visibility); the window only moves the origin of the rows -- an exact fp32
        // subtraction of a small integer -- and drops what cannot touch its rows
        o.m2y = __fsub_rn(o.m2y, (float)cam_row0[c]);
        const float r = (float)o.radius;
        if (o.radius > 0 && (__fadd_rn(o.m2y, r) <= 0.f || __fsub_rn(o.m2y, r) >= (float)window_height)) o.radius = 0;
    }
    radii[idx] = o.radius;
    reinterpret_cast<float2 *>(means2d)[idx] = make_float2(o.m2x, o.m2y);
    depths[idx] = o.depth;
    conics[3 * idx] = o.ca;
    conics[3 * idx + 1] = o.cb;
    conics[3 * idx + 2] = o.cc;
    if (tiles_per_gauss) {
        int n = 0;
        if (o.radius > 0) {
            int x0, y0, x1, y1;
            tile_rect(o.m2x, o.m2y, o.radius, tile_size, tile_w, tile_h, &x0, &y0, &x1, &y1);
            n = (y1 - y0) * (x1 - x0);
        }
        tiles_per_gauss[idx] = n;
    }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__global__ void __launch_bounds__(kProjThreads)
project_bwd_kernel(const float *__restrict__ means, int64_t means_cs, const float *__restrict__ quats,
                   int64_t quats_cs, const float *__restrict__ scales,
                   const float *__restrict__ viewmats, int64_t vm_cs, const float *__restrict__ Ks,
                   int64_t k_cs, int C, int G, int width, int height,
                   const int32_t *__restrict__ radii, const float *__restrict__ conics,
                   const float *__restrict__ v_means2d, const float *__restrict__ v_depths,
                   const float *__restrict__ v_conics, float *__restrict__ v_means,
                   float *__restrict__ v_quats, float *__restrict__ v_scales,
                   float *__restrict__ v_viewmats) {
    // grid: x over Gaussians, y over cameras (so that a block never straddles cameras)
    int c = blockIdx.y;
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    bool active = g < G;
    int64_t idx = (int64_t)c * G + (active ? g : 0);
    if (active) active = radii[idx] > 0;
    ProjGrad pg;
#pragma unroll
    for (int i = 0; i < 9; ++i) pg.v_R[i] = 0.f;
#pragma unroll
    for (int i = 0; i < 3; ++i) pg.v_t[i] = 0.f;
    if (active) {
        const float *V = viewmats + c * vm_cs;
        const float *K = Ks + c * k_cs;
        float Vl[12], Kl[6];
#pragma unroll
        for (int i = 0; i < 12; ++i) Vl[i] = __ldg(V + i);
#pragma unroll
        for (int i = 0; i < 6; ++i) Kl[i] = __ldg(K + i);
        const float *mp = means + c * means_cs + 3LL * g;
        float m[3] = {__ldg(mp), __ldg(mp + 1), __ldg(mp + 2)};
        float4 q4 = __ldg(reinterpret_cast<const float4 *>(quats + c * quats_cs) + g);
        float q[4] = {q4.x, q4.y, q4.z, q4.w};
        const float *sp = scales + 3LL * g;
        float s[3] = {__ldg(sp), __ldg(sp + 1), __ldg(sp + 2)};
        float2 vm2 = __ldg(reinterpret_cast<const float2 *>(v_means2d) + idx);
        project_one_bwd(m, q, s, Vl, Kl, width, height, __ldg(conics + 3 * idx),
                        __ldg(conics + 3 * idx + 1), __ldg(conics + 3 * idx + 2), vm2.x, vm2.y,
                        __ldg(v_depths + idx), __ldg(v_conics + 3 * idx), __ldg(v_conics + 3 * idx + 1),
                        __ldg(v_conics + 3 * idx + 2), &pg);
        float *om = v_means + c * means_cs + 3LL * g;
        float *oq = v_quats + c * quats_cs + 4LL * g;
        float *os = v_scales + 3LL * g;
        if (C == 1 || means_cs != 0) {
            om[0] = pg.v_mean[0]; om[1] = pg.v_mean[1]; om[2] = pg.v_mean[2];
        } else {
            atomicAdd(om, pg.v_mean[0]); atomicAdd(om + 1, pg.v_mean[1]); atomicAdd(om + 2, pg.v_mean[2]);
        }
        if (C == 1 || quats_cs != 0) {
            *reinterpret_cast<float4 *>(oq) = make_float4(pg.v_quat[0], pg.v_quat[1], pg.v_quat[2], pg.v_quat[3]);
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) atomicAdd(oq + j, pg.v_quat[j]);
        }
        if (C == 1) {
            os[0] = pg.v_scale[0]; os[1] = pg.v_scale[1]; os[2] = pg.v_scale[2];
        } else {
#pragma unroll
            for (int j = 0; j < 3; ++j) atomicAdd(os + j, pg.v_scale[j]);
        }
    }
    if (v_viewmats) {
        // block reduction of the 12 view-matrix gradient entries, one atomic set per block
        __shared__ float red[kProjThreads / 32][12];
        int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
        for (int i = 0; i < 9; ++i) {
            float v = warp_sum(pg.v_R[i]);
            if (lane == 0) red[w][i] = v;
        }
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            float v = warp_sum(pg.v_t[i]);
            if (lane == 0) red[w][9 + i] = v;
        }
        __syncthreads();
        if (threadIdx.x < 12) {
            float acc = 0.f;
#pragma unroll
            for (int k = 0; k < kProjThreads / 32; ++k) acc += red[k][threadIdx.x];
            int i = threadIdx.x;
            // entries 0..8 -> rotation R[i/3][i%3] at viewmat[4*(i/3) + i%3]; 9..11 -> translation column
            int pos = (i < 9) ? (4 * (i / 3) + (i % 3)) : (4 * (i - 9) + 3);
            if (acc != 0.f) atomicAdd(v_viewmats + 16LL * c + pos, acc);
        }
    }
}

__global__ void __launch_bounds__(kProjThreads)
isect_emit_kernel(const float *__restrict__ means2d, const int32_t *__restrict__ radii,
                  const float *__restrict__ depths, const int32_t *__restrict__ cum_excl, int C, int G,
                  int tile_size, int tile_w, int tile_h, int tile_n_bits,
                  int64_t *__restrict__ isect_ids, int32_t *__restrict__ flatten_ids) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)C * G) return;
    int32_t r = radii[idx];
    if (r <= 0) return;
    float2 m2 = __ldg(reinterpret_cast<const float2 *>(means2d) + idx);
    int x0, y0, x1, y1;
    tile_rect(m2.x, m2.y, r, tile_size, tile_w, tile_h, &x0, &y0, &x1, &y1);
    int64_t cid = idx / G;
    int64_t cid_enc = cid << (32 + tile_n_bits);
    int64_t depth_enc = (int64_t)__float_as_int(depths[idx]);
    int64_t cur = cum_excl[idx];
    for (int i = y0; i < y1; ++i)
        for (int j = x0; j < x1; ++j) {
            int64_t tile_id = (int64_t)i * tile_w + j;
            isect_ids[cur] = cid_enc | (tile_id << 32) | depth_enc;
            flatten_ids[cur] = (int32_t)idx;
            ++cur;
        }
}

}  // namespace d4

using namespace d4;

extern "C" int d4_tile_n_bits(int n_tiles) {
    int b = 0;
    while ((1LL << (b + 1)) <= n_tiles) ++b;
    return b + 1;
}

extern "C" int d4_project_fwd(const float *means, int64_t means_cam_stride, const float *quats,
                              int64_t quats_cam_stride, const float *scales, const float *viewmats,
                              int64_t viewmat_cam_stride, const float *Ks, int64_t k_cam_stride, int C,
                              int G, int width, int height, float eps2d, float near_plane,
                              float far_plane, float radius_clip, int tile_size, int tile_w, int tile_h,
                              int32_t *radii, float *means2d, float *depths, float *conics,
                              int32_t *tiles_per_gauss, const int32_t *cam_row0, int window_height,
                              d4_stream_t stream) {
    D4_CHECK_ARG(C >= 1 && G >= 0 && width > 0 && height > 0 && tile_size > 0, "d4_project_fwd: bad sizes");
    D4_CHECK_ARG(!cam_row0 || (window_height > 0 && tile_h == (window_height + tile_size - 1) / tile_size),
                 "d4_project_fwd: a row window needs window_height > 0 and tile_h of the window");
    if (G == 0) return 0;
    D4_CHECK_ARG(means && quats && scales && viewmats && Ks && radii && means2d && depths && conics,
                 "d4_project_fwd: null pointer");
    D4_CHECK_ARG(((uintptr_t)quats & 15) == 0 && ((uintptr_t)means2d & 7) == 0 && (quats_cam_stride % 4) == 0,
                 "d4_project_fwd: quats must be 16-byte and means2d 8-byte aligned");
    if (G == 0) return 0;
    int64_t n = (int64_t)C * G;
    project_fwd_kernel<<<cdiv(n, kProjThreads), kProjThreads, 0, as_stream(stream)>>>(
        means, means_cam_stride, quats, quats_cam_stride, scales, viewmats, viewmat_cam_stride, Ks,
        k_cam_stride, C, G, width, height, eps2d, near_plane, far_plane, radius_clip, tile_size, tile_w,
        tile_h, radii, means2d, depths, conics, tiles_per_gauss, cam_row0, window_height);
    D4_CHECK_LAUNCH("d4_project_fwd");
    return 0;
}

extern "C" int d4_project_bwd(const float *means, int64_t means_cam_stride, const float *quats,
                              int64_t quats_cam_stride, const float *scales, const float *viewmats,
                              int64_t viewmat_cam_stride, const float *Ks, int64_t k_cam_stride, int C,
                              int G, int width, int height, float eps2d, const int32_t *radii,
                              const float *conics, const float *v_means2d, const float *v_depths,
                              const float *v_conics, float *v_means, float *v_quats, float *v_scales,
                              float *v_viewmats, d4_stream_t stream) {
    (void)eps2d;
    D4_CHECK_ARG(C >= 1 && C <= 65535 && G >= 0, "d4_project_bwd: bad sizes");
    if (G == 0) return 0;
    D4_CHECK_ARG(means && quats && scales && viewmats && Ks && radii && conics && v_means2d && v_depths &&
                     v_conics && v_means && v_quats && v_scales,
                 "d4_project_bwd: null pointer");
    D4_CHECK_ARG(((uintptr_t)quats & 15) == 0 && ((uintptr_t)v_quats & 15) == 0 && ((uintptr_t)v_means2d & 7) == 0 &&
                     (quats_cam_stride % 4) == 0,
                 "d4_project_bwd: alignment");
    if (G == 0) return 0;
    dim3 grid(cdiv(G, kProjThreads), C);
    project_bwd_kernel<<<grid, kProjThreads, 0, as_stream(stream)>>>(
        means, means_cam_stride, quats, quats_cam_stride, scales, viewmats, viewmat_cam_stride, Ks,
        k_cam_stride, C, G, width, height, radii, conics, v_means2d, v_depths, v_conics, v_means, v_quats,
        v_scales, v_viewmats);
    D4_CHECK_LAUNCH("d4_project_bwd");
    return 0;
}

extern "C" int d4_isect_emit(const float *means2d, const int32_t *radii, const float *depths,
                             const int32_t *cum_tiles_exclusive, int C, int G, int tile_size, int tile_w,
                             int tile_h, int64_t *isect_ids, int32_t *flatten_ids, d4_stream_t stream) {
    if ((int64_t)C * G == 0) return 0;
    D4_CHECK_ARG(means2d && radii && depths && cum_tiles_exclusive, "d4_isect_emit: null pointer");
    int tb = d4_tile_n_bits(tile_w * tile_h);
    isect_emit_kernel<<<cdiv((int64_t)C * G, kProjThreads), kProjThreads, 0, as_stream(stream)>>>(
        means2d, radii, depths, cum_tiles_exclusive, C, G, tile_size, tile_w, tile_h, tb, isect_ids,
        flatten_ids);
    D4_CHECK_LAUNCH("d4_isect_emit");
    return 0;
}
