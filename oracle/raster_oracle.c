/*
 * oracle/raster_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C, fp32, OpenMP over Gaussians / tiles) of the
 * rasterizer the reference calls at flow3d/scene_model.py:360-373:
 * gsplat==1.1.1 `rasterization(..., packed=False)` (requirements.txt:137).
 *
 * PARITY UNPINNED: gsplat is a CUDA-only, un-vendored third-party dependency
 * that is absent from /root/reference and cannot be installed here, and the
 * reference holds no golden vectors for this path.  The
 * arithmetic below restates the published gsplat v1.1.x algorithm
 * (fully_fused_projection, isect_tiles, isect_offset_encode,
 * rasterize_to_pixels fwd/bwd).  It is pinned
 * internally only: its forward is cross-checked against an independent dense
 * pure-torch restatement (oracle/raster_torch.py) and its hand-derived
 * backward against torch autograd of that restatement (tests/test_oracle.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.
 *
 * Floating point: strict fp32, no FMA contraction (build with
 * -ffp-contract=off, no -ffast-math).  The projection spells out the
 * evaluation order term by term; the CUDA projection kernel is compiled with
 * -fmad=false and follows the same order so that radii / tile rectangles /
 * depth-key bits can be compared bit-exactly.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_API __attribute__((visibility("default")))

ORC_API int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

ORC_API void orc_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* ------------------------------------------------------------------------ */
/* projection forward (gsplat fully_fused_projection_fwd, packed=False)      */
/* ------------------------------------------------------------------------ */

static inline void quat_to_rotmat(const float *q, float R[9]) {
    /* q = (w, x, y, z), un-normalised; normalised here */
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
}

/* world covariance (6 unique, order 00 01 02 11 12 22) from quat + scale */
static inline void quat_scale_to_covar(const float *q, const float *s, float cv[6]) {
    float R[9];
    quat_to_rotmat(q, R);
    float M[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) M[3 * i + j] = R[3 * i + j] * s[j];
    cv[0] = M[0] * M[0] + M[1] * M[1] + M[2] * M[2];
    cv[1] = M[0] * M[3] + M[1] * M[4] + M[2] * M[5];
    cv[2] = M[0] * M[6] + M[1] * M[7] + M[2] * M[8];
    cv[3] = M[3] * M[3] + M[4] * M[4] + M[5] * M[5];
    cv[4] = M[3] * M[6] + M[4] * M[7] + M[5] * M[8];
    cv[5] = M[6] * M[6] + M[7] * M[7] + M[8] * M[8];
}

/* cc = Rv * cov * Rv^T (6 unique); cov given as 6 unique */
static inline void covar_world_to_cam(const float Rv[9], const float cv[6], float cc[6]) {
    float S[9] = {cv[0], cv[1], cv[2], cv[1], cv[3], cv[4], cv[2], cv[4], cv[5]};
    float A[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            A[3 * i + j] = Rv[3 * i + 0] * S[0 + j] + Rv[3 * i + 1] * S[3 + j] + Rv[3 * i + 2] * S[6 + j];
    cc[0] = A[0] * Rv[0] + A[1] * Rv[1] + A[2] * Rv[2];
    cc[1] = A[0] * Rv[3] + A[1] * Rv[4] + A[2] * Rv[5];
    cc[2] = A[0] * Rv[6] + A[1] * Rv[7] + A[2] * Rv[8];
    cc[3] = A[3] * Rv[3] + A[4] * Rv[4] + A[5] * Rv[5];
    cc[4] = A[3] * Rv[6] + A[4] * Rv[7] + A[5] * Rv[8];
    cc[5] = A[6] * Rv[6] + A[7] * Rv[7] + A[8] * Rv[8];
}

/*
 * means: [Cm, G, 3] with camera stride means_cs (0 => shared by all cameras)
 * quats: [Cq, G, 4] with camera stride quats_cs; scales [G,3]
 * viewmats [C,4,4] row-major, Ks [C,3,3] row-major (stride 0 allowed via vm_cs/k_cs)
 * outputs: radii i32 [C,G], means2d [C,G,2], depths [C,G], conics [C,G,3]
 */
ORC_API void orc_project_fwd(const float *means, long means_cs, const float *quats, long quats_cs,
                             const float *scales, const float *viewmats, long vm_cs, const float *Ks,
                             long k_cs, int C, int G, int width, int height, float eps2d,
                             float near_plane, float far_plane, float radius_clip, int32_t *radii,
                             float *means2d, float *depths, float *conics) {
#pragma omp parallel for schedule(static)
    for (long idx = 0; idx < (long)C * G; ++idx) {
        int c = (int)(idx / G), g = (int)(idx % G);
        const float *V = viewmats + c * vm_cs;
        const float *K = Ks + c * k_cs;
        const float *m = means + c * means_cs + 3L * g;
        const float *q = quats + c * quats_cs + 4L * g;
        const float *s = scales + 3L * g;
        radii[idx] = 0;
        means2d[2 * idx] = 0.f; means2d[2 * idx + 1] = 0.f;
        depths[idx] = 0.f;
        conics[3 * idx] = conics[3 * idx + 1] = conics[3 * idx + 2] = 0.f;

        float Rv[9] = {V[0], V[1], V[2], V[4], V[5], V[6], V[8], V[9], V[10]};
        float x = Rv[0] * m[0] + Rv[1] * m[1] + Rv[2] * m[2] + V[3];
        float y = Rv[3] * m[0] + Rv[4] * m[1] + Rv[5] * m[2] + V[7];
        float z = Rv[6] * m[0] + Rv[7] * m[1] + Rv[8] * m[2] + V[11];
        if (z < near_plane || z > far_plane) continue;

        float cv[6], cc[6];
        quat_scale_to_covar(q, s, cv);
        covar_world_to_cam(Rv, cv, cc);

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
        /* v0 = cc * (J00,0,J02)^T ; v1 = cc * (0,J11,J12)^T */
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
        if (det <= 0.f) continue;
        float invdet = 1.0f / det;
        float ca = c11 * invdet, cb = -c01 * invdet, ccn = c00 * invdet;

        float b = 0.5f * (c00 + c11);
        float v1 = b + sqrtf(fmaxf(0.01f, b * b - det));
        float radius = ceilf(3.0f * sqrtf(v1));
        if (radius <= radius_clip) continue;
        if (m2x + radius <= 0.f || m2x - radius >= (float)width || m2y + radius <= 0.f ||
            m2y - radius >= (float)height)
            continue;

        radii[idx] = (int32_t)radius;
        means2d[2 * idx] = m2x; means2d[2 * idx + 1] = m2y;
        depths[idx] = z;
        conics[3 * idx] = ca; conics[3 * idx + 1] = cb; conics[3 * idx + 2] = ccn;
    }
}

/* ------------------------------------------------------------------------ */
/* projection backward (gsplat fully_fused_projection_bwd)                    */
/* ------------------------------------------------------------------------ */

/*
 * v_means [Cm,G,3] / v_quats [Cq,G,4] use the same camera strides as the
 * inputs; with stride 0 contributions of all cameras are summed.  v_scales
 * [G,3] sums over cameras.  v_viewmats [C,4,4] (may be NULL).  All outputs
 * must be zero-initialised by the caller.  Accumulation in double so that the
 * oracle carries no summation-order noise.
 */
ORC_API void orc_project_bwd(const float *means, long means_cs, const float *quats, long quats_cs,
                             const float *scales, const float *viewmats, long vm_cs, const float *Ks,
                             long k_cs, int C, int G, int width, int height, float eps2d,
                             const int32_t *radii, const float *conics, const float *v_means2d,
                             const float *v_depths, const float *v_conics, double *v_means,
                             double *v_quats, double *v_scales, double *v_viewmats) {
    (void)eps2d;
    for (int c = 0; c < C; ++c) {
        const float *V = viewmats + c * vm_cs;
        const float *K = Ks + c * k_cs;
        float Rv[9] = {V[0], V[1], V[2], V[4], V[5], V[6], V[8], V[9], V[10]};
        double vR_acc[9] = {0}, vt_acc[3] = {0};
#pragma omp parallel
        {
            double vR_loc[9] = {0}, vt_loc[3] = {0};
#pragma omp for schedule(static)
            for (int g = 0; g < G; ++g) {
                long idx = (long)c * G + g;
                if (radii[idx] <= 0) continue;
                const float *m = means + c * means_cs + 3L * g;
                const float *q = quats + c * quats_cs + 4L * g;
                const float *s = scales + 3L * g;
                float fx = K[0], fy = K[4];
                float x = Rv[0] * m[0] + Rv[1] * m[1] + Rv[2] * m[2] + V[3];
                float y = Rv[3] * m[0] + Rv[4] * m[1] + Rv[5] * m[2] + V[7];
                float z = Rv[6] * m[0] + Rv[7] * m[1] + Rv[8] * m[2] + V[11];
                float cv[6], cc[6];
                quat_scale_to_covar(q, s, cv);
                covar_world_to_cam(Rv, cv, cc);

                /* conic = inverse(cov2d_blur): v_cov2d = -Minv * v_Minv * Minv */
                float a = conics[3 * idx], b = conics[3 * idx + 1], d = conics[3 * idx + 2];
                float va = v_conics[3 * idx], vb = 0.5f * v_conics[3 * idx + 1], vd = v_conics[3 * idx + 2];
                /* P = Minv * vM (2x2), Q = P * Minv */
                float p00 = a * va + b * vb, p01 = a * vb + b * vd;
                float p10 = b * va + d * vb, p11 = b * vb + d * vd;
                float g00 = -(p00 * a + p01 * b);
                float g01 = -(p00 * b + p01 * d);
                float g10 = -(p10 * a + p11 * b);
                float g11 = -(p10 * b + p11 * d);
                /* symmetric by construction; keep general form */

                /* persp_proj_vjp */
                float tan_fovx = 0.5f * (float)width / fx;
                float tan_fovy = 0.5f * (float)height / fy;
                float lim_x = 1.3f * tan_fovx, lim_y = 1.3f * tan_fovy;
                float rz = 1.0f / z, rz2 = rz * rz, rz3 = rz2 * rz;
                float tx = z * fminf(lim_x, fmaxf(-lim_x, x * rz));
                float ty = z * fminf(lim_y, fmaxf(-lim_y, y * rz));
                float J[6] = {fx * rz, 0.f, -fx * tx * rz2, 0.f, fy * rz, -fy * ty * rz2}; /* 2x3 */
                float S[9] = {cc[0], cc[1], cc[2], cc[1], cc[3], cc[4], cc[2], cc[4], cc[5]};
                float Gm[4] = {g00, g01, g10, g11};
                /* v_cc = J^T G J (3x3) */
                float GJ[6];
                for (int i = 0; i < 2; ++i)
                    for (int j = 0; j < 3; ++j) GJ[3 * i + j] = Gm[2 * i] * J[j] + Gm[2 * i + 1] * J[3 + j];
                float v_cc[9];
                for (int i = 0; i < 3; ++i)
                    for (int j = 0; j < 3; ++j) v_cc[3 * i + j] = J[i] * GJ[j] + J[3 + i] * GJ[3 + j];
                /* v_J = G J S^T + G^T J S (2x3), S symmetric */
                float JS[6];
                for (int i = 0; i < 2; ++i)
                    for (int j = 0; j < 3; ++j)
                        JS[3 * i + j] = J[3 * i] * S[j] + J[3 * i + 1] * S[3 + j] + J[3 * i + 2] * S[6 + j];
                float vJ[6];
                for (int i = 0; i < 2; ++i)
                    for (int j = 0; j < 3; ++j)
                        vJ[3 * i + j] = (Gm[2 * i] + Gm[i]) * JS[j] + (Gm[2 * i + 1] + Gm[2 + i]) * JS[3 + j];
                float vm2x = v_means2d[2 * idx], vm2y = v_means2d[2 * idx + 1];
                float vmc[3];
                vmc[0] = fx * rz * vm2x;
                vmc[1] = fy * rz * vm2y;
                vmc[2] = -(fx * x * vm2x + fy * y * vm2y) * rz2;
                if (x * rz <= lim_x && x * rz >= -lim_x) vmc[0] += -fx * rz2 * vJ[2];
                else vmc[2] += -fx * rz3 * vJ[2] * tx;
                if (y * rz <= lim_y && y * rz >= -lim_y) vmc[1] += -fy * rz2 * vJ[5];
                else vmc[2] += -fy * rz3 * vJ[5] * ty;
                vmc[2] += -fx * rz2 * vJ[0] - fy * rz2 * vJ[4] + 2.f * fx * tx * rz3 * vJ[2] +
                          2.f * fy * ty * rz3 * vJ[5];
                vmc[2] += v_depths[idx];

                /* pos_world_to_cam_vjp: mc = Rv m + t */
                float vm[3];
                for (int j = 0; j < 3; ++j) vm[j] = Rv[j] * vmc[0] + Rv[3 + j] * vmc[1] + Rv[6 + j] * vmc[2];
                for (int i = 0; i < 3; ++i) {
                    for (int j = 0; j < 3; ++j) vR_loc[3 * i + j] += (double)vmc[i] * m[j];
                    vt_loc[i] += vmc[i];
                }
                /* covar_world_to_cam_vjp: cc = Rv cov Rv^T
                   v_cov = Rv^T v_cc Rv ; v_Rv = v_cc Rv cov^T + v_cc^T Rv cov */
                float W3[9] = {cv[0], cv[1], cv[2], cv[1], cv[3], cv[4], cv[2], cv[4], cv[5]};
                float T1[9], v_cov[9];
                for (int i = 0; i < 3; ++i)
                    for (int j = 0; j < 3; ++j)
                        T1[3 * i + j] = v_cc[3 * i] * Rv[j] + v_cc[3 * i + 1] * Rv[3 + j] + v_cc[3 * i + 2] * Rv[6 + j];
                for (int i = 0; i < 3; ++i)
                    for (int j = 0; j < 3; ++j)
                        v_cov[3 * i + j] = Rv[i] * T1[j] + Rv[3 + i] * T1[3 + j] + Rv[6 + i] * T1[6 + j];
                float RW[9];
                for (int i = 0; i < 3; ++i)
                    for (int j = 0; j < 3; ++j)
                        RW[3 * i + j] = Rv[3 * i] * W3[j] + Rv[3 * i + 1] * W3[3 + j] + Rv[3 * i + 2] * W3[6 + j];
                for (int i = 0; i < 3; ++i)
                    for (int j = 0; j < 3; ++j) {
                        float acc = 0.f;
                        for (int k = 0; k < 3; ++k) acc += (v_cc[3 * i + k] + v_cc[3 * k + i]) * RW[3 * k + j];
                        vR_loc[3 * i + j] += acc;
                    }

                /* quat_scale_to_covar_vjp: cov = M M^T, M = R S */
                float R[9];
                quat_to_rotmat(q, R);
                float M[9];
                for (int i = 0; i < 3; ++i)
                    for (int j = 0; j < 3; ++j) M[3 * i + j] = R[3 * i + j] * s[j];
                float vM[9];
                for (int i = 0; i < 3; ++i)
                    for (int j = 0; j < 3; ++j) {
                        float acc = 0.f;
                        for (int k = 0; k < 3; ++k) acc += (v_cov[3 * i + k] + v_cov[3 * k + i]) * M[3 * k + j];
                        vM[3 * i + j] = acc;
                    }
                float vs[3], vR[9];
                for (int j = 0; j < 3; ++j) vs[j] = R[j] * vM[j] + R[3 + j] * vM[3 + j] + R[6 + j] * vM[6 + j];
                for (int i = 0; i < 3; ++i)
                    for (int j = 0; j < 3; ++j) vR[3 * i + j] = vM[3 * i + j] * s[j];
                /* quat (normalised) from vR */
                float qw = q[0], qx = q[1], qy = q[2], qz = q[3];
                float inv = 1.0f / sqrtf(qx * qx + qy * qy + qz * qz + qw * qw);
                qw *= inv; qx *= inv; qy *= inv; qz *= inv;
                float vqn[4];
                vqn[0] = 2.f * (qx * (vR[7] - vR[5]) + qy * (vR[2] - vR[6]) + qz * (vR[3] - vR[1]));
                vqn[1] = 2.f * (-2.f * qx * (vR[4] + vR[8]) + qy * (vR[1] + vR[3]) + qz * (vR[2] + vR[6]) + qw * (vR[7] - vR[5]));
                vqn[2] = 2.f * (qx * (vR[1] + vR[3]) - 2.f * qy * (vR[0] + vR[8]) + qz * (vR[5] + vR[7]) + qw * (vR[2] - vR[6]));
                vqn[3] = 2.f * (qx * (vR[2] + vR[6]) + qy * (vR[5] + vR[7]) - 2.f * qz * (vR[0] + vR[4]) + qw * (vR[3] - vR[1]));
                float dotp = vqn[0] * qw + vqn[1] * qx + vqn[2] * qy + vqn[3] * qz;
                float vq[4] = {(vqn[0] - dotp * qw) * inv, (vqn[1] - dotp * qx) * inv,
                               (vqn[2] - dotp * qy) * inv, (vqn[3] - dotp * qz) * inv};

                double *om = v_means + c * means_cs + 3L * g;
                double *oq = v_quats + c * quats_cs + 4L * g;
                double *os = v_scales + 3L * g;
                if (means_cs == 0 && C > 1) {
                    for (int j = 0; j < 3; ++j) {
#pragma omp atomic
                        om[j] += vm[j];
                    }
                } else
                    for (int j = 0; j < 3; ++j) om[j] += vm[j];
                if (quats_cs == 0 && C > 1) {
                    for (int j = 0; j < 4; ++j) {
#pragma omp atomic
                        oq[j] += vq[j];
                    }
                } else
                    for (int j = 0; j < 4; ++j) oq[j] += vq[j];
                for (int j = 0; j < 3; ++j) os[j] += vs[j]; /* cameras are serial, g unique per thread */
            }
#pragma omp critical
            {
                for (int i = 0; i < 9; ++i) vR_acc[i] += vR_loc[i];
                for (int i = 0; i < 3; ++i) vt_acc[i] += vt_loc[i];
            }
        }
        if (v_viewmats) {
            double *o = v_viewmats + 16L * c;
            for (int i = 0; i < 3; ++i) {
                for (int j = 0; j < 3; ++j) o[4 * i + j] += vR_acc[3 * i + j];
                o[4 * i + 3] += vt_acc[i];
            }
        }
    }
}

/* ------------------------------------------------------------------------ */
/* tile intersection (gsplat isect_tiles), sort, offsets                      */
/* ------------------------------------------------------------------------ */

static inline void tile_rect(const float *m2, int32_t radius, int tile_size, int tw, int th,
                             int *x0, int *y0, int *x1, int *y1) {
    float ts = (float)tile_size;
    float tr = (float)radius / ts;
    float tx = m2[0] / ts, ty = m2[1] / ts;
    float fx0 = floorf(tx - tr), fy0 = floorf(ty - tr);
    float fx1 = ceilf(tx + tr), fy1 = ceilf(ty + tr);
    /* (uint32_t) conversion of a negative float saturates to 0 on the GPU */
    long a;
    a = fx0 < 0.f ? 0 : (long)fx0; *x0 = (int)(a > tw ? tw : a);
    a = fy0 < 0.f ? 0 : (long)fy0; *y0 = (int)(a > th ? th : a);
    a = fx1 < 0.f ? 0 : (long)fx1; *x1 = (int)(a > tw ? tw : a);
    a = fy1 < 0.f ? 0 : (long)fy1; *y1 = (int)(a > th ? th : a);
}

/* pass 1: tiles_per_gauss [C,G] i32; returns total number of intersections */
ORC_API int64_t orc_isect_count(const float *means2d, const int32_t *radii, int C, int G,
                                int tile_size, int tw, int th, int32_t *tiles_per_gauss) {
    int64_t total = 0;
    for (long idx = 0; idx < (long)C * G; ++idx) {
        int32_t n = 0;
        if (radii[idx] > 0) {
            int x0, y0, x1, y1;
            tile_rect(means2d + 2 * idx, radii[idx], tile_size, tw, th, &x0, &y0, &x1, &y1);
            n = (y1 - y0) * (x1 - x0);
        }
        tiles_per_gauss[idx] = n;
        total += n;
    }
    return total;
}

ORC_API int orc_tile_n_bits(int n_tiles) {
    /* (uint32_t)floor(log2(n_tiles)) + 1 */
    int b = 0;
    while ((1L << (b + 1)) <= n_tiles) ++b;
    return b + 1;
}

/* pass 2: emit unsorted isect_ids (i64) / flatten_ids (i32) in (c,g,tile row-major) order */
ORC_API void orc_isect_emit(const float *means2d, const int32_t *radii, const float *depths, int C,
                            int G, int tile_size, int tw, int th, int64_t *isect_ids,
                            int32_t *flatten_ids) {
    int tile_n_bits = orc_tile_n_bits(tw * th);
    int64_t cur = 0;
    for (long idx = 0; idx < (long)C * G; ++idx) {
        if (radii[idx] <= 0) continue;
        int x0, y0, x1, y1;
        tile_rect(means2d + 2 * idx, radii[idx], tile_size, tw, th, &x0, &y0, &x1, &y1);
        int64_t cid = idx / G;
        int64_t cid_enc = cid << (32 + tile_n_bits);
        int32_t dbits;
        memcpy(&dbits, depths + idx, 4);
        int64_t depth_enc = (int64_t)dbits;
        for (int i = y0; i < y1; ++i)
            for (int j = x0; j < x1; ++j) {
                int64_t tile_id = (int64_t)i * tw + j;
                isect_ids[cur] = cid_enc | (tile_id << 32) | depth_enc;
                flatten_ids[cur] = (int32_t)idx;
                ++cur;
            }
    }
}

/* stable LSD radix sort of (i64 key, i32 val) pairs, 16-bit digits, ascending */
ORC_API void orc_sort_pairs(int64_t *keys, int32_t *vals, int64_t n, int end_bit) {
    if (n <= 1) return;
    int64_t *k2 = (int64_t *)malloc(sizeof(int64_t) * (size_t)n);
    int32_t *v2 = (int32_t *)malloc(sizeof(int32_t) * (size_t)n);
    size_t *cnt = (size_t *)malloc(sizeof(size_t) * 65537);
    int64_t *ka = keys, *kb = k2;
    int32_t *va = vals, *vb = v2;
    for (int shift = 0; shift < end_bit; shift += 16) {
        memset(cnt, 0, sizeof(size_t) * 65537);
        for (int64_t i = 0; i < n; ++i) cnt[(((uint64_t)ka[i]) >> shift & 0xFFFF) + 1]++;
        for (int d = 0; d < 65536; ++d) cnt[d + 1] += cnt[d];
        for (int64_t i = 0; i < n; ++i) {
            size_t p = cnt[((uint64_t)ka[i]) >> shift & 0xFFFF]++;
            kb[p] = ka[i];
            vb[p] = va[i];
        }
        int64_t *tk = ka; ka = kb; kb = tk;
        int32_t *tv = va; va = vb; vb = tv;
    }
    if (ka != keys) {
        memcpy(keys, ka, sizeof(int64_t) * (size_t)n);
        memcpy(vals, va, sizeof(int32_t) * (size_t)n);
    }
    free(k2); free(v2); free(cnt);
}

/* isect_offset_encode: offsets [C, th, tw] i32 = first sorted index of each (cam,tile) */
ORC_API void orc_tile_offsets(const int64_t *isect_ids_sorted, int64_t n_isects, int C, int tw,
                              int th, int32_t *offsets) {
    int n_tiles = tw * th;
    int tile_n_bits = orc_tile_n_bits(n_tiles);
    int64_t total = (int64_t)C * n_tiles;
    int64_t next = 0; /* next (cam,tile) linear id whose offset is not yet written */
    for (int64_t i = 0; i < n_isects; ++i) {
        int64_t hi = isect_ids_sorted[i] >> 32;
        int64_t cam = hi >> tile_n_bits;
        int64_t tile = hi & ((1L << tile_n_bits) - 1);
        int64_t lin = cam * n_tiles + tile;
        while (next <= lin && next < total) offsets[next++] = (int32_t)i;
    }
    while (next < total) offsets[next++] = (int32_t)n_isects;
}

/* ------------------------------------------------------------------------ */
/* blend forward (gsplat rasterize_to_pixels_fwd)                             */
/* ------------------------------------------------------------------------ */

#define ALPHA_MIN (1.0f / 255.0f)
#define ALPHA_MAX 0.999f
#define T_MIN 1e-4f
/* relative half-width of the "knife-edge" band around the two thresholds */
#define EDGE_REL 2e-5f

/*
 * means2d [C*G,2], conics [C*G,3], opacities [C*G], colors [C*G, D] (already
 * expanded per camera, as gsplat's python wrapper does), backgrounds [C,D] or
 * NULL.  Outputs: render_colors [C,H,W,D], render_alphas [C,H,W], last_ids
 * [C,H,W] i32, edge [C,H,W] u8 (1 where some threshold decision of this pixel
 * was within EDGE_REL of flipping -- such pixels may legitimately differ from
 * an implementation whose exp() differs in the last ulps; may be NULL).
 */
ORC_API void orc_blend_fwd(const float *means2d, const float *conics, const float *opacities,
                           const float *colors, const float *backgrounds, int C, int D, int width,
                           int height, int tile_size, int tw, int th, const int32_t *tile_offsets,
                           const int32_t *flatten_ids, int64_t n_isects, float *render_colors,
                           float *render_alphas, int32_t *last_ids, uint8_t *edge) {
    long n_tiles = (long)tw * th;
#pragma omp parallel for schedule(dynamic, 4)
    for (long ct = 0; ct < C * n_tiles; ++ct) {
        int c = (int)(ct / n_tiles);
        int tile = (int)(ct % n_tiles);
        int ty = tile / tw, tx = tile % tw;
        int64_t start = tile_offsets[ct];
        int64_t end = (ct == C * n_tiles - 1) ? n_isects : tile_offsets[ct + 1];
        float pix[64];
        for (int iy = 0; iy < tile_size; ++iy)
            for (int ix = 0; ix < tile_size; ++ix) {
                int i = ty * tile_size + iy, j = tx * tile_size + ix;
                if (i >= height || j >= width) continue;
                float px = (float)j + 0.5f, py = (float)i + 0.5f;
                float T = 1.0f;
                int32_t cur = 0;
                uint8_t e = 0;
                float *out = (D <= 64) ? pix : (float *)malloc(sizeof(float) * D);
                for (int k = 0; k < D; ++k) out[k] = 0.f;
                for (int64_t idx = start; idx < end; ++idx) {
                    int32_t g = flatten_ids[idx];
                    float dx = means2d[2L * g] - px, dy = means2d[2L * g + 1] - py;
                    float ca = conics[3L * g], cb = conics[3L * g + 1], cd = conics[3L * g + 2];
                    float sigma = 0.5f * (ca * dx * dx + cd * dy * dy) + cb * dx * dy;
                    float opac = opacities[g];
                    float alpha = fminf(ALPHA_MAX, opac * expf(-sigma));
                    if (fabsf(alpha - ALPHA_MIN) <= EDGE_REL * ALPHA_MIN * fmaxf(1.f, fabsf(sigma))) e = 1;
                    if (sigma < 0.f || alpha < ALPHA_MIN) continue;
                    float next_T = T * (1.0f - alpha);
                    if (fabsf(next_T - T_MIN) <= 16.f * EDGE_REL * T_MIN) e = 1;
                    if (next_T <= T_MIN) break;
                    float vis = alpha * T;
                    const float *cp = colors + (long)g * D;
                    for (int k = 0; k < D; ++k) out[k] += cp[k] * vis;
                    cur = (int32_t)idx;
                    T = next_T;
                }
                long pid = ((long)c * height + i) * width + j;
                render_alphas[pid] = 1.0f - T;
                for (int k = 0; k < D; ++k)
                    render_colors[pid * D + k] = backgrounds ? out[k] + T * backgrounds[(long)c * D + k] : out[k];
                last_ids[pid] = cur;
                if (edge) edge[pid] = e;
                if (out != pix) free(out);
            }
    }
}

/*
 * Per-intersection hit masks of the forward (the product's BlendArgs::hit_masks): bit w of hit_masks[idx] is set
 * iff some pixel of the w-th 8x4 pixel block of the tile (w = (iy / 4) * 2 + ix / 8) reaches intersection idx
 * before it has terminated and passes the alpha test there (the Gaussian that saturates a pixel counts: it is
 * tested, not composited).  edge_isect[idx] = 1 where one of those decisions was within the knife-edge band.
 * Same per-pixel walk as orc_blend_fwd; tile_size must be 16.  Both outputs zero-initialised by the caller.
 */
ORC_API void orc_hit_masks(const float *means2d, const float *conics, const float *opacities, int C, int width,
                           int height, int tile_size, int tw, int th, const int32_t *tile_offsets,
                           const int32_t *flatten_ids, int64_t n_isects, uint8_t *hit_masks, uint8_t *edge_isect) {
    long n_tiles = (long)tw * th;
#pragma omp parallel for schedule(dynamic, 4)
    for (long ct = 0; ct < C * n_tiles; ++ct) {
        int tile = (int)(ct % n_tiles);
        int ty = tile / tw, tx = tile % tw;
        int64_t start = tile_offsets[ct];
        int64_t end = (ct == C * n_tiles - 1) ? n_isects : tile_offsets[ct + 1];
        for (int iy = 0; iy < tile_size; ++iy)
            for (int ix = 0; ix < tile_size; ++ix) {
                int i = ty * tile_size + iy, j = tx * tile_size + ix;
                if (i >= height || j >= width) continue;
                const uint8_t bit = (uint8_t)(1u << ((iy / 4) * 2 + ix / 8));
                float px = (float)j + 0.5f, py = (float)i + 0.5f;
                float T = 1.0f;
                for (int64_t idx = start; idx < end; ++idx) {
                    int32_t g = flatten_ids[idx];
                    float dx = means2d[2L * g] - px, dy = means2d[2L * g + 1] - py;
                    float ca = conics[3L * g], cb = conics[3L * g + 1], cd = conics[3L * g + 2];
                    float sigma = 0.5f * (ca * dx * dx + cd * dy * dy) + cb * dx * dy;
                    float alpha = fminf(ALPHA_MAX, opacities[g] * expf(-sigma));
                    if (fabsf(alpha - ALPHA_MIN) <= EDGE_REL * ALPHA_MIN * fmaxf(1.f, fabsf(sigma))) edge_isect[idx] = 1;
                    if (sigma < 0.f || alpha < ALPHA_MIN) continue;
                    hit_masks[idx] |= bit; /* one tile per thread: no race */
                    float next_T = T * (1.0f - alpha);
                    if (fabsf(next_T - T_MIN) <= 16.f * EDGE_REL * T_MIN) {
                        /* the termination decision is on the knife edge: everything behind it may differ */
                        for (int64_t r = idx; r < end; ++r) edge_isect[r] = 1;
                    }
                    if (next_T <= T_MIN) break;
                    T = next_T;
                }
            }
    }
}

/* ------------------------------------------------------------------------ */
/* blend backward (gsplat rasterize_to_pixels_bwd)                            */
/* ------------------------------------------------------------------------ */

/*
 * Per-pixel arithmetic in fp32 exactly as gsplat's kernel; the per-Gaussian
 * sums over pixels (atomicAdd in gsplat, order undefined) are accumulated in
 * double: v_means2d [C*G,2], v_conics [C*G,3], v_colors [C*G,D], v_opacities
 * [C*G] -- all zero-initialised by the caller.
 */
ORC_API void orc_blend_bwd(const float *means2d, const float *conics, const float *opacities,
                           const float *colors, const float *backgrounds, int C, int G, int D,
                           int width, int height, int tile_size, int tw, int th,
                           const int32_t *tile_offsets, const int32_t *flatten_ids,
                           int64_t n_isects, const float *render_alphas, const int32_t *last_ids,
                           const float *v_render_colors, const float *v_render_alphas,
                           double *v_means2d, double *v_conics, double *v_colors,
                           double *v_opacities) {
    (void)G;
    long n_tiles = (long)tw * th;
#pragma omp parallel for schedule(dynamic, 4)
    for (long ct = 0; ct < C * n_tiles; ++ct) {
        int c = (int)(ct / n_tiles);
        int tile = (int)(ct % n_tiles);
        int ty = tile / tw, tx = tile % tw;
        int64_t start = tile_offsets[ct];
        int64_t end = (ct == C * n_tiles - 1) ? n_isects : tile_offsets[ct + 1];
        if (end <= start) continue;
        int64_t nt = end - start;
        int V = D + 6; /* per-Gaussian local sums: D colors, 3 conic, 2 xy, 1 opacity */
        float *buffer = (float *)malloc(sizeof(float) * D);
        double *loc = (double *)calloc((size_t)nt * V, sizeof(double));
        for (int iy = 0; iy < tile_size; ++iy)
            for (int ix = 0; ix < tile_size; ++ix) {
                int i = ty * tile_size + iy, j = tx * tile_size + ix;
                if (i >= height || j >= width) continue;
                long pid = ((long)c * height + i) * width + j;
                float px = (float)j + 0.5f, py = (float)i + 0.5f;
                float T_final = 1.0f - render_alphas[pid];
                float T = T_final;
                for (int k = 0; k < D; ++k) buffer[k] = 0.f;
                int32_t bin_final = last_ids[pid];
                const float *vrc = v_render_colors + pid * D;
                float vra = v_render_alphas[pid];
                float bgdot = 0.f;
                if (backgrounds)
                    for (int k = 0; k < D; ++k) bgdot += backgrounds[(long)c * D + k] * vrc[k];
                for (int64_t idx = (bin_final < end - 1 ? bin_final : end - 1); idx >= start; --idx) {
                    int32_t g = flatten_ids[idx];
                    float dx = means2d[2L * g] - px, dy = means2d[2L * g + 1] - py;
                    float ca = conics[3L * g], cb = conics[3L * g + 1], cd = conics[3L * g + 2];
                    float sigma = 0.5f * (ca * dx * dx + cd * dy * dy) + cb * dx * dy;
                    float opac = opacities[g];
                    float vis = expf(-sigma);
                    float alpha = fminf(ALPHA_MAX, opac * vis);
                    if (sigma < 0.f || alpha < ALPHA_MIN) continue;
                    double *acc = loc + (idx - start) * V;
                    float ra = 1.0f / (1.0f - alpha);
                    T *= ra;
                    float fac = alpha * T;
                    const float *cp = colors + (long)g * D;
                    float v_alpha = 0.f;
                    for (int k = 0; k < D; ++k) {
                        acc[k] += (double)(fac * vrc[k]);
                        v_alpha += (cp[k] * T - buffer[k] * ra) * vrc[k];
                    }
                    v_alpha += T_final * ra * vra;
                    if (backgrounds) v_alpha += -T_final * ra * bgdot;
                    if (opac * vis <= ALPHA_MAX) {
                        float v_sigma = -opac * vis * v_alpha;
                        acc[D + 0] += (double)(0.5f * v_sigma * dx * dx);
                        acc[D + 1] += (double)(v_sigma * dx * dy);
                        acc[D + 2] += (double)(0.5f * v_sigma * dy * dy);
                        acc[D + 3] += (double)(v_sigma * (ca * dx + cb * dy));
                        acc[D + 4] += (double)(v_sigma * (cb * dx + cd * dy));
                        acc[D + 5] += (double)(vis * v_alpha);
                    }
                    for (int k = 0; k < D; ++k) buffer[k] += cp[k] * fac;
                }
            }
        /* flush the tile-local sums (one atomic per value per (tile, Gaussian)) */
        for (int64_t r = 0; r < nt; ++r) {
            int32_t g = flatten_ids[start + r];
            const double *acc = loc + r * V;
            for (int k = 0; k < D; ++k) {
#pragma omp atomic
                v_colors[(long)g * D + k] += acc[k];
            }
            for (int k = 0; k < 3; ++k) {
#pragma omp atomic
                v_conics[3L * g + k] += acc[D + k];
            }
            for (int k = 0; k < 2; ++k) {
#pragma omp atomic
                v_means2d[2L * g + k] += acc[D + 3 + k];
            }
#pragma omp atomic
            v_opacities[g] += acc[D + 5];
        }
        free(loc);
        free(buffer);
    }
}
