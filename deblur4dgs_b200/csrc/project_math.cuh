// project_math.cuh -- per-Gaussian projection arithmetic (SURVEY.md rows a8, a9
// first pass, a12), shared by the kernels in project.cu.
//
// Everything here is plain IEEE fp32 in a FIXED evaluation order: project.cu is
// compiled with -fmad=false so that no multiply-add is contracted, sqrt and
// division are the correctly rounded ones.  That makes radii, tile rectangles
// and depth-key bits reproducible bit for bit by any other strict-fp32
// implementation of the same expressions (the parity tests rely on this).
//
// The functions are __host__ __device__ so that tests/host_harness can run the
// very same arithmetic on the CPU without a GPU.
#pragma once
#include <math.h>
#include <stdint.h>

#ifdef __CUDACC__
#define D4_HD __host__ __device__ __forceinline__
#else
#define D4_HD static inline
#endif

namespace d4 {

struct ProjOut {
    int32_t radius;  // 0 = culled
    float m2x, m2y, depth, ca, cb, cc;
};

// R (row-major 3x3) of the normalised quaternion q = (w,x,y,z); also returns
// the normalised quaternion and 1/|q| for the backward pass.
D4_HD void quat_to_rotmat(const float q[4], float R[9], float qn[4], float *inv_norm) {
    float w = q[0], x = q[1], y = q[2], z = q[3];
    float n2 = x * x + y * y + z * z + w * w;
    float inv = 1.0f / sqrtf(n2);
    x = x * inv; y = y * inv; z = z * inv; w = w * inv;
    float x2 = x * x, y2 = y * y, z2 = z * z;
    float xy = x * y, xz = x * z, yz = y * z;
    float wx = w * x, wy = w * y, wz = w * z;
    R[0] = 1.0f - 2.0f * (y2 + z2); R[1] = 2.0f * (xy - wz);        R[2] = 2.0f * (xz + wy);
    R[3] = 2.0f * (xy + wz);        R[4] = 1.0f - 2.0f * (x2 + z2); R[5] = 2.0f * (yz - wx);
    R[6] = 2.0f * (xz - wy);        R[7] = 2.0f * (yz + wx);        R[8] = 1.0f - 2.0f * (x2 + y2);
    qn[0] = w; qn[1] = x; qn[2] = y; qn[3] = z;
    *inv_norm = inv;
}

// world covariance, 6 unique entries (00 01 02 11 12 22), cov = (R S)(R S)^T
D4_HD void covar_from_RS(const float R[9], const float s[3], float M[9], float cv[6]) {
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) M[3 * i + j] = R[3 * i + j] * s[j];
    cv[0] = M[0] * M[0] + M[1] * M[1] + M[2] * M[2];
    cv[1] = M[0] * M[3] + M[1] * M[4] + M[2] * M[5];
    cv[2] = M[0] * M[6] + M[1] * M[7] + M[2] * M[8];
    cv[3] = M[3] * M[3] + M[4] * M[4] + M[5] * M[5];
    cv[4] = M[3] * M[6] + M[4] * M[7] + M[5] * M[8];
    cv[5] = M[6] * M[6] + M[7] * M[7] + M[8] * M[8];
}

// cc = Rv cov Rv^T (6 unique)
D4_HD void covar_to_cam(const float Rv[9], const float cv[6], float cc[6]) {
    const float S[9] = {cv[0], cv[1], cv[2], cv[1], cv[3], cv[4], cv[2], cv[4], cv[5]};
    float A[9];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            A[3 * i + j] = Rv[3 * i + 0] * S[0 + j] + Rv[3 * i + 1] * S[3 + j] + Rv[3 * i + 2] * S[6 + j];
    cc[0] = A[0] * Rv[0] + A[1] * Rv[1] + A[2] * Rv[2];
    cc[1] = A[0] * Rv[3] + A[1] * Rv[4] + A[2] * Rv[5];
    cc[2] = A[0] * Rv[6] + A[1] * Rv[7] + A[2] * Rv[8];
    cc[3] = A[3] * Rv[3] + A[4] * Rv[4] + A[5] * Rv[5];
    cc[4] = A[3] * Rv[6] + A[4] * Rv[7] + A[5] * Rv[8];
    cc[5] = A[6] * Rv[6] + A[7] * Rv[7] + A[8] * Rv[8];
}

// Forward projection of one Gaussian (gsplat fully_fused_projection_fwd).
// V = row-major 4x4 world->camera, K = row-major 3x3 intrinsics.
D4_HD ProjOut project_one(const float m[3], const float q[4], const float s[3], const float *V,
                          const float *K, int width, int height, float eps2d, float near_plane,
                          float far_plane, float radius_clip) {
    ProjOut o;
    o.radius = 0; o.m2x = 0.f; o.m2y = 0.f; o.depth = 0.f; o.ca = 0.f; o.cb = 0.f; o.cc = 0.f;
    const float Rv[9] = {V[0], V[1], V[2], V[4], V[5], V[6], V[8], V[9], V[10]};
    float x = Rv[0] * m[0] + Rv[1] * m[1] + Rv[2] * m[2] + V[3];
    float y = Rv[3] * m[0] + Rv[4] * m[1] + Rv[5] * m[2] + V[7];
    float z = Rv[6] * m[0] + Rv[7] * m[1] + Rv[8] * m[2] + V[11];
    if (z < near_plane || z > far_plane) return o;

    float R[9], qn[4], inv, M[9], cv[6], cc[6];
    quat_to_rotmat(q, R, qn, &inv);
    covar_from_RS(R, s, M, cv);
    covar_to_cam(Rv, cv, cc);

    float fx = K[0], fy = K[4], cx = K[2], cy = K[5];
    float tan_fovx = 0.5f * (float)width / fx;
    float tan_fovy = 0.5f * (float)height / fy;
    float lim_x = 1.3f * tan_fovx, lim_y = 1.3f * tan_fovy;
    float rz = 1.0f / z;
    float rz2 = rz * rz;
    float tx = z * fminf(lim_x, fmaxf(-lim_x, x * rz));
    float ty = z * fminf(lim_y, fmaxf(-lim_y, y * rz));
    float J00 = fx * rz, J02 = -fx * tx * rz2;
    float J11 = fy * rz, J12 = -fy * ty * rz2;
    float v0x = cc[0] * J00 + cc[2] * J02;
    float v0z = cc[2] * J00 + cc[5] * J02;
    float v1x = cc[1] * J11 + cc[2] * J12;
    float v1y = cc[3] * J11 + cc[4] * J12;
    float v1z = cc[4] * J11 + cc[5] * J12;
    float c00 = J00 * v0x + J02 * v0z;
    float c01 = J00 * v1x + J02 * v1z;
    float c11 = J11 * v1y + J12 * v1z;
    float m2x = fx * x * rz + cx;
    float m2y = fy * y * rz + cy;

    c00 = c00 + eps2d;
    c11 = c11 + eps2d;
    float det = c00 * c11 - c01 * c01;
    if (det <= 0.f) return o;
    float invdet = 1.0f / det;

    float b = 0.5f * (c00 + c11);
    float v1 = b + sqrtf(fmaxf(0.01f, b * b - det));
    float radius = ceilf(3.0f * sqrtf(v1));
    if (radius <= radius_clip) return o;
    if (m2x + radius <= 0.f || m2x - radius >= (float)width || m2y + radius <= 0.f ||
        m2y - radius >= (float)height)
        return o;
    o.radius = (int32_t)radius;
    o.m2x = m2x; o.m2y = m2y; o.depth = z;
    o.ca = c11 * invdet; o.cb = -c01 * invdet; o.cc = c00 * invdet;
    return o;
}

// tile rectangle [x0,x1) x [y0,y1) covered by the square means2d +- radius
// (gsplat isect_tiles; the float->uint conversion saturates negatives to 0)
D4_HD void tile_rect(float m2x, float m2y, int32_t radius, int tile_size, int tile_w, int tile_h,
                     int *x0, int *y0, int *x1, int *y1) {
    float ts = (float)tile_size;
    float tr = (float)radius / ts;
    float tx = m2x / ts, ty = m2y / ts;
    float fx0 = floorf(tx - tr), fy0 = floorf(ty - tr);
    float fx1 = ceilf(tx + tr), fy1 = ceilf(ty + tr);
    float fw = (float)tile_w, fh = (float)tile_h;
    *x0 = (int)fminf(fmaxf(fx0, 0.f), fw);
    *y0 = (int)fminf(fmaxf(fy0, 0.f), fh);
    *x1 = (int)fminf(fmaxf(fx1, 0.f), fw);
    *y1 = (int)fminf(fmaxf(fy1, 0.f), fh);
}

struct ProjGrad {
    float v_mean[3], v_quat[4], v_scale[3];
    float v_R[9], v_t[3];  // contribution to d/d viewmat (rotation part, translation)
};

// Backward of project_one for a non-culled Gaussian (gsplat
// fully_fused_projection_bwd: inverse_vjp, persp_proj_vjp,
// pos/covar_world_to_cam_vjp, quat_scale_to_covar_vjp).
D4_HD void project_one_bwd(const float m[3], const float q[4], const float s[3], const float *V,
                           const float *K, int width, int height, float ca, float cb, float cd,
                           float v_m2x, float v_m2y, float v_depth, float v_ca, float v_cb,
                           float v_cd, ProjGrad *o) {
    const float Rv[9] = {V[0], V[1], V[2], V[4], V[5], V[6], V[8], V[9], V[10]};
    float fx = K[0], fy = K[4];
    float x = Rv[0] * m[0] + Rv[1] * m[1] + Rv[2] * m[2] + V[3];
    float y = Rv[3] * m[0] + Rv[4] * m[1] + Rv[5] * m[2] + V[7];
    float z = Rv[6] * m[0] + Rv[7] * m[1] + Rv[8] * m[2] + V[11];
    float R[9], qn[4], inv, M[9], cv[6], cc[6];
    quat_to_rotmat(q, R, qn, &inv);
    covar_from_RS(R, s, M, cv);
    covar_to_cam(Rv, cv, cc);

    // conic = inverse(cov2d): v_cov2d = -conic * v_conic_sym * conic
    float va = v_ca, vb = 0.5f * v_cb, vd = v_cd;
    float p00 = ca * va + cb * vb, p01 = ca * vb + cb * vd;
    float p10 = cb * va + cd * vb, p11 = cb * vb + cd * vd;
    float Gm[4];
    Gm[0] = -(p00 * ca + p01 * cb);
    Gm[1] = -(p00 * cb + p01 * cd);
    Gm[2] = -(p10 * ca + p11 * cb);
    Gm[3] = -(p10 * cb + p11 * cd);

    float tan_fovx = 0.5f * (float)width / fx;
    float tan_fovy = 0.5f * (float)height / fy;
    float lim_x = 1.3f * tan_fovx, lim_y = 1.3f * tan_fovy;
    float rz = 1.0f / z, rz2 = rz * rz, rz3 = rz2 * rz;
    float tx = z * fminf(lim_x, fmaxf(-lim_x, x * rz));
    float ty = z * fminf(lim_y, fmaxf(-lim_y, y * rz));
    const float J[6] = {fx * rz, 0.f, -fx * tx * rz2, 0.f, fy * rz, -fy * ty * rz2};
    const float S[9] = {cc[0], cc[1], cc[2], cc[1], cc[3], cc[4], cc[2], cc[4], cc[5]};
    float GJ[6], v_cc[9], JS[6], vJ[6];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) GJ[3 * i + j] = Gm[2 * i] * J[j] + Gm[2 * i + 1] * J[3 + j];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) v_cc[3 * i + j] = J[i] * GJ[j] + J[3 + i] * GJ[3 + j];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            JS[3 * i + j] = J[3 * i] * S[j] + J[3 * i + 1] * S[3 + j] + J[3 * i + 2] * S[6 + j];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            vJ[3 * i + j] = (Gm[2 * i] + Gm[i]) * JS[j] + (Gm[2 * i + 1] + Gm[2 + i]) * JS[3 + j];
    float vmc[3];
    vmc[0] = fx * rz * v_m2x;
    vmc[1] = fy * rz * v_m2y;
    vmc[2] = -(fx * x * v_m2x + fy * y * v_m2y) * rz2;
    if (x * rz <= lim_x && x * rz >= -lim_x) vmc[0] += -fx * rz2 * vJ[2];
    else vmc[2] += -fx * rz3 * vJ[2] * tx;
    if (y * rz <= lim_y && y * rz >= -lim_y) vmc[1] += -fy * rz2 * vJ[5];
    else vmc[2] += -fy * rz3 * vJ[5] * ty;
    vmc[2] += -fx * rz2 * vJ[0] - fy * rz2 * vJ[4] + 2.f * fx * tx * rz3 * vJ[2] +
              2.f * fy * ty * rz3 * vJ[5];
    vmc[2] += v_depth;

    // mc = Rv m + t
#pragma unroll
    for (int j = 0; j < 3; ++j) o->v_mean[j] = Rv[j] * vmc[0] + Rv[3 + j] * vmc[1] + Rv[6 + j] * vmc[2];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
        for (int j = 0; j < 3; ++j) o->v_R[3 * i + j] = vmc[i] * m[j];
        o->v_t[i] = vmc[i];
    }
    // cc = Rv cov Rv^T: v_cov = Rv^T v_cc Rv ; v_Rv += (v_cc + v_cc^T) Rv cov
    const float W3[9] = {cv[0], cv[1], cv[2], cv[1], cv[3], cv[4], cv[2], cv[4], cv[5]};
    float T1[9], v_cov[9], RW[9];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            T1[3 * i + j] = v_cc[3 * i] * Rv[j] + v_cc[3 * i + 1] * Rv[3 + j] + v_cc[3 * i + 2] * Rv[6 + j];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            v_cov[3 * i + j] = Rv[i] * T1[j] + Rv[3 + i] * T1[3 + j] + Rv[6 + i] * T1[6 + j];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            RW[3 * i + j] = Rv[3 * i] * W3[j] + Rv[3 * i + 1] * W3[3 + j] + Rv[3 * i + 2] * W3[6 + j];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            float acc = 0.f;
#pragma unroll
            for (int k = 0; k < 3; ++k) acc += (v_cc[3 * i + k] + v_cc[3 * k + i]) * RW[3 * k + j];
            o->v_R[3 * i + j] += acc;
        }
    // cov = M M^T, M = R S
    float vM[9], vR[9];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            float acc = 0.f;
#pragma unroll
            for (int k = 0; k < 3; ++k) acc += (v_cov[3 * i + k] + v_cov[3 * k + i]) * M[3 * k + j];
            vM[3 * i + j] = acc;
        }
#pragma unroll
    for (int j = 0; j < 3; ++j) o->v_scale[j] = R[j] * vM[j] + R[3 + j] * vM[3 + j] + R[6 + j] * vM[6 + j];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) vR[3 * i + j] = vM[3 * i + j] * s[j];
    float qw = qn[0], qx = qn[1], qy = qn[2], qz = qn[3];
    float vqn[4];
    vqn[0] = 2.f * (qx * (vR[7] - vR[5]) + qy * (vR[2] - vR[6]) + qz * (vR[3] - vR[1]));
    vqn[1] = 2.f * (-2.f * qx * (vR[4] + vR[8]) + qy * (vR[1] + vR[3]) + qz * (vR[2] + vR[6]) + qw * (vR[7] - vR[5]));
    vqn[2] = 2.f * (qx * (vR[1] + vR[3]) - 2.f * qy * (vR[0] + vR[8]) + qz * (vR[5] + vR[7]) + qw * (vR[2] - vR[6]));
    vqn[3] = 2.f * (qx * (vR[2] + vR[6]) + qy * (vR[5] + vR[7]) - 2.f * qz * (vR[0] + vR[4]) + qw * (vR[3] - vR[1]));
    float dotp = vqn[0] * qw + vqn[1] * qx + vqn[2] * qy + vqn[3] * qz;
    o->v_quat[0] = (vqn[0] - dotp * qw) * inv;
    o->v_quat[1] = (vqn[1] - dotp * qx) * inv;
    o->v_quat[2] = (vqn[2] - dotp * qy) * inv;
    o->v_quat[3] = (vqn[3] - dotp * qz) * inv;
}

}  // namespace d4
