// deform.cu -- rows a1-a6 of SURVEY.md section 8, fused over all N sub-exposures:
// softmax(motion_coefs) -> K-basis blend at floor/ceil frames -> time lerp ->
// Gram-Schmidt -> R mu + t, quat(R) (x) q -> camera sub-exposure transform.
// The reference spends ~40 elementwise/einsum launches per sub-exposure on this
// (params.py:142-180, scene_model.py:76-120, 352-353); here the canonical
// parameters are read ONCE for all N timestamps (4*(K+7) B per fg Gaussian) and
// only the N deformed (mean, quat) pairs are written (28 B each): HBM-bound.
//
// Backward uses forward-mode dual numbers for the per-Gaussian SE(3) part (see
// deform_math.cuh) and reduces the basis / time / camera gradients over
// Gaussians with a transposing warp butterfly -> shared memory -> one global
// atomic per CTA and value.
#include "common.cuh"
#include "deform_math.cuh"

namespace d4 {

constexpr int kDefThreads = 128;
constexpr int kMaxK = 64;

struct FramePair {
    int pre, nxt;
    float w;
};

__device__ __forceinline__ FramePair frame_pair(float t, int T) {
    // params.py:152-153,173: indices clamped, weight from the clamped floor
    FramePair f;
    float lo = fminf(fmaxf(floorf(t), 0.f), (float)(T - 1));
    float hi = fminf(fmaxf(ceilf(t), 0.f), (float)(T - 1));
    f.pre = (int)lo;
    f.nxt = (int)hi;
    f.w = t - (float)f.pre;
    return f;
}

// blended (transl 3, rot6d 6) at one frame: sum_k c_k * base[k, frame]
__device__ __forceinline__ void blend_bases(const float *s_coef, int K, int T, int frame,
                                            const float *__restrict__ rots,
                                            const float *__restrict__ transls, float out[9]) {
#pragma unroll
    for (int j = 0; j < 9; ++j) out[j] = 0.f;
    for (int k = 0; k < K; ++k) {
        const float ck = s_coef[k * kDefThreads];
        const float *tp = transls + ((int64_t)k * T + frame) * 3;
        const float *rp = rots + ((int64_t)k * T + frame) * 6;
#pragma unroll
        for (int j = 0; j < 3; ++j) out[j] += ck * __ldg(tp + j);
#pragma unroll
        for (int j = 0; j < 6; ++j) out[3 + j] += ck * __ldg(rp + j);
    }
}

// softmax over K raw logits of one Gaussian, result in s_coef[k * kDefThreads]
__device__ __forceinline__ void softmax_coefs(const float *__restrict__ raw, int K, float *s_coef) {
    float mx = -INFINITY;
    for (int k = 0; k < K; ++k) {
        float v = __ldg(raw + k);
        s_coef[k * kDefThreads] = v;
        mx = fmaxf(mx, v);
    }
    float sum = 0.f;
    for (int k = 0; k < K; ++k) {
        float e = expf(s_coef[k * kDefThreads] - mx);
        s_coef[k * kDefThreads] = e;
        sum += e;
    }
    for (int k = 0; k < K; ++k) s_coef[k * kDefThreads] = s_coef[k * kDefThreads] / sum;
}

__device__ __forceinline__ void apply_rt(const float *__restrict__ RT, const float m[3], float o[3]) {
    if (RT == nullptr) {
        o[0] = m[0]; o[1] = m[1]; o[2] = m[2];
        return;
    }
#pragma unroll
    for (int i = 0; i < 3; ++i)
        o[i] = __ldg(RT + 4 * i) * m[0] + __ldg(RT + 4 * i + 1) * m[1] + __ldg(RT + 4 * i + 2) * m[2] + __ldg(RT + 4 * i + 3);
}

__global__ void __launch_bounds__(kDefThreads)
deform_fg_fwd_kernel(const float *__restrict__ fg_means, const float *__restrict__ fg_quats,
                     const float *__restrict__ coefs_raw, const float *__restrict__ rots,
                     const float *__restrict__ transls, const float *__restrict__ times,
                     const float *__restrict__ RTs, int Gf, int G, int K, int T, int N,
                     float *__restrict__ out_means, float *__restrict__ out_quats) {
    extern __shared__ float s_coef_all[];  // [K][kDefThreads]
    int g = blockIdx.x * kDefThreads + threadIdx.x;
    if (g >= Gf) return;
    float *s_coef = s_coef_all + threadIdx.x;
    softmax_coefs(coefs_raw + (int64_t)g * K, K, s_coef);
    float mu[3] = {__ldg(fg_means + 3LL * g), __ldg(fg_means + 3LL * g + 1), __ldg(fg_means + 3LL * g + 2)};
    float4 q4 = __ldg(reinterpret_cast<const float4 *>(fg_quats) + g);
    float q[4] = {q4.x, q4.y, q4.z, q4.w};
    for (int n = 0; n < N; ++n) {
        FramePair f = frame_pair(__ldg(times + n), T);
        float bp[9], bn[9], bl[9];
        blend_bases(s_coef, K, T, f.pre, rots, transls, bp);
        blend_bases(s_coef, K, T, f.nxt, rots, transls, bn);
#pragma unroll
        for (int j = 0; j < 9; ++j) bl[j] = (1.0f - f.w) * bp[j] + f.w * bn[j];
        float om[3], oq[4], oc[3];
        deform_point<float>(bl, bl + 3, mu, q, om, oq);
        apply_rt(RTs ? RTs + 12LL * n : nullptr, om, oc);
        float *mo = out_means + ((int64_t)n * G + g) * 3;
        mo[0] = oc[0]; mo[1] = oc[1]; mo[2] = oc[2];
        reinterpret_cast<float4 *>(out_quats)[(int64_t)n * G + g] = make_float4(oq[0], oq[1], oq[2], oq[3]);
    }
}

__global__ void __launch_bounds__(256)
deform_bg_fwd_kernel(const float *__restrict__ bg_means, const float *__restrict__ bg_quats,
                     const float *__restrict__ RTs, int Gf, int Gb, int G, int N,
                     float *__restrict__ out_means, float *__restrict__ out_quats) {
    int g = blockIdx.x * 256 + threadIdx.x;
    if (g >= Gb) return;
    float mu[3] = {__ldg(bg_means + 3LL * g), __ldg(bg_means + 3LL * g + 1), __ldg(bg_means + 3LL * g + 2)};
    float4 q = __ldg(reinterpret_cast<const float4 *>(bg_quats) + g);
    float nrm = fmaxf(sqrtf(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w), 1e-12f);
    float4 qh = make_float4(q.x / nrm, q.y / nrm, q.z / nrm, q.w / nrm);
    for (int n = 0; n < N; ++n) {
        float oc[3];
        apply_rt(RTs ? RTs + 12LL * n : nullptr, mu, oc);
        float *mo = out_means + ((int64_t)n * G + Gf + g) * 3;
        mo[0] = oc[0]; mo[1] = oc[1]; mo[2] = oc[2];
        reinterpret_cast<float4 *>(out_quats)[(int64_t)n * G + Gf + g] = qh;
    }
}

// ------------------------------------------------------------------------------- backward
template <int NV>
__device__ __forceinline__ void warp_transpose_reduce_d(float (&v)[NV], int lane) {
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

// v_out_means'' -> v_mu' = R_rt^T v, and the 12 camera-delta gradient entries (+1 spare slot
// used by the fg kernel for the time gradient), reduced over the CTA into s_red[16]
__device__ __forceinline__ void rt_backward(const float *__restrict__ RT, const float vm[3], const float mprime[3],
                                            float vmu[3], float r16[16]) {
    if (RT == nullptr) {
        vmu[0] = vm[0]; vmu[1] = vm[1]; vmu[2] = vm[2];
        return;
    }
#pragma unroll
    for (int j = 0; j < 3; ++j) vmu[j] = __ldg(RT + j) * vm[0] + __ldg(RT + 4 + j) * vm[1] + __ldg(RT + 8 + j) * vm[2];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
        for (int j = 0; j < 3; ++j) r16[4 * i + j] = vm[i] * mprime[j];
        r16[4 * i + 3] = vm[i];
    }
}

constexpr int kDefStride = kDefThreads + 1;  // row stride of the backward's [K][threads] / [9][threads] tiles: a thread's own
                                             // column and a (k, j) thread's walk along a row are both bank-conflict free

__global__ void __launch_bounds__(kDefThreads)
deform_fg_bwd_kernel(const float *__restrict__ fg_means, const float *__restrict__ fg_quats,
                     const float *__restrict__ coefs_raw, const float *__restrict__ rots,
                     const float *__restrict__ transls, const float *__restrict__ times,
                     const float *__restrict__ RTs, int Gf, int G, int K, int T, int N,
                     const float *__restrict__ v_out_means, const float *__restrict__ v_out_quats,
                     float *__restrict__ v_fg_means, float *__restrict__ v_fg_quats,
                     float *__restrict__ v_coefs_raw, float *__restrict__ v_rots,
                     float *__restrict__ v_transls, float *__restrict__ v_times, float *__restrict__ v_RTs) {
    // Per Gaussian and timestamp: hand-derived reverse-mode VJP of deform_point (deform_math.cuh).  The basis gradient
    // v_B[k][j] = sum_g c_gk * v_blend_gj is a [K x 128] x [128 x 9] product per CTA and timestamp: the coefficients
    // already sit in shared memory, the 9 blend cotangents are parked next to them, and K*9 threads each walk one
    // (k, j) row pair -- no warp shuffles, one global atomic per (CTA, timestamp, k, j, frame).
    extern __shared__ float smem[];
    float *s_coef_all = smem;                          // [K][kDefStride] softmaxed coefficients
    float *s_vcoef_all = smem + K * kDefStride;        // [K][kDefStride] dL/dcoef
    float *s_vb = smem + 2 * K * kDefStride;           // [9][kDefStride] blend cotangents of the current timestamp
    float *s_red = s_vb + 9 * kDefStride;              // [16]     12 camera entries + time gradient
    const int tid = threadIdx.x, lane = tid & 31;
    const int g = blockIdx.x * kDefThreads + tid;
    const bool active = g < Gf;
    float *s_coef = s_coef_all + tid;
    float *s_vcoef = s_vcoef_all + tid;
    float mu[3] = {0.f, 0.f, 0.f}, q[4] = {1.f, 0.f, 0.f, 0.f};
    if (active) {
        // softmax over K raw logits (as softmax_coefs, with the padded stride)
        const float *raw = coefs_raw + (int64_t)g * K;
        float mx = -INFINITY;
        for (int k = 0; k < K; ++k) {
            const float v = __ldg(raw + k);
            s_coef[k * kDefStride] = v;
            mx = fmaxf(mx, v);
        }
        float sum = 0.f;
        for (int k = 0; k < K; ++k) {
            const float e = expf(s_coef[k * kDefStride] - mx);
            s_coef[k * kDefStride] = e;
            sum += e;
        }
        for (int k = 0; k < K; ++k) s_coef[k * kDefStride] = s_coef[k * kDefStride] / sum;
        mu[0] = __ldg(fg_means + 3LL * g); mu[1] = __ldg(fg_means + 3LL * g + 1); mu[2] = __ldg(fg_means + 3LL * g + 2);
        float4 q4 = __ldg(reinterpret_cast<const float4 *>(fg_quats) + g);
        q[0] = q4.x; q[1] = q4.y; q[2] = q4.z; q[3] = q4.w;
    } else {
        for (int k = 0; k < K; ++k) s_coef[k * kDefStride] = 0.f;
    }
    for (int k = 0; k < K; ++k) s_vcoef[k * kDefStride] = 0.f;
    float v_mu[3] = {0.f, 0.f, 0.f}, v_q[4] = {0.f, 0.f, 0.f, 0.f};

    for (int n = 0; n < N; ++n) {
        if (tid < 16) s_red[tid] = 0.f;
        const FramePair f = frame_pair(__ldg(times + n), T);
        const float *RT = RTs ? RTs + 12LL * n : nullptr;
        float vb[9];
        float r16[16];
#pragma unroll
        for (int j = 0; j < 9; ++j) vb[j] = 0.f;
#pragma unroll
        for (int j = 0; j < 16; ++j) r16[j] = 0.f;
        if (active) {
            // blended (transl 3, rot6d 6) at the two frames
            float bp[9], bn[9], bl[9];
#pragma unroll
            for (int j = 0; j < 9; ++j) bp[j] = bn[j] = 0.f;
            for (int k = 0; k < K; ++k) {
                const float ck = s_coef[k * kDefStride];
                const float *tp = transls + ((int64_t)k * T + f.pre) * 3, *tn = transls + ((int64_t)k * T + f.nxt) * 3;
                const float *rp = rots + ((int64_t)k * T + f.pre) * 6, *rn = rots + ((int64_t)k * T + f.nxt) * 6;
#pragma unroll
                for (int j = 0; j < 3; ++j) bp[j] += ck * __ldg(tp + j), bn[j] += ck * __ldg(tn + j);
#pragma unroll
                for (int j = 0; j < 6; ++j) bp[3 + j] += ck * __ldg(rp + j), bn[3 + j] += ck * __ldg(rn + j);
            }
#pragma unroll
            for (int j = 0; j < 9; ++j) bl[j] = (1.0f - f.w) * bp[j] + f.w * bn[j];
            const float *vmp = v_out_means + ((int64_t)n * G + g) * 3;
            const float vm[3] = {__ldg(vmp), __ldg(vmp + 1), __ldg(vmp + 2)};
            const float4 vq4 = __ldg(reinterpret_cast<const float4 *>(v_out_quats) + (int64_t)n * G + g);
            const float vq[4] = {vq4.x, vq4.y, vq4.z, vq4.w};
            // camera delta backward: v_mu' = R_rt^T v
            float vmu[3] = {vm[0], vm[1], vm[2]};
            if (RT != nullptr) {
#pragma unroll
                for (int j = 0; j < 3; ++j) vmu[j] = __ldg(RT + j) * vm[0] + __ldg(RT + 4 + j) * vm[1] + __ldg(RT + 8 + j) * vm[2];
            }
            float grad[16], mprime[3];
            deform_point_vjp(bl, bl + 3, mu, q, vmu, vq, mprime, grad);
            if (RT != nullptr) {
#pragma unroll
                for (int i = 0; i < 3; ++i) {
#pragma unroll
                    for (int j = 0; j < 3; ++j) r16[4 * i + j] = vm[i] * mprime[j];
                    r16[4 * i + 3] = vm[i];
                }
            }
#pragma unroll
            for (int j = 0; j < 9; ++j) vb[j] = grad[j];
#pragma unroll
            for (int j = 0; j < 3; ++j) v_mu[j] += grad[9 + j];
#pragma unroll
            for (int j = 0; j < 4; ++j) v_q[j] += grad[12 + j];
            // time gradient: d/dw of the lerp
            float vw = 0.f;
#pragma unroll
            for (int j = 0; j < 9; ++j) vw += (bn[j] - bp[j]) * vb[j];
            r16[15] = vw;
            // dL/dcoef_k += sum_j lerp(base_pre, base_next)[k][j] * vb_j
            for (int k = 0; k < K; ++k) {
                const float *tp = transls + ((int64_t)k * T + f.pre) * 3, *tn = transls + ((int64_t)k * T + f.nxt) * 3;
                const float *rp = rots + ((int64_t)k * T + f.pre) * 6, *rn = rots + ((int64_t)k * T + f.nxt) * 6;
                float acc = 0.f;
#pragma unroll
                for (int j = 0; j < 3; ++j) acc += ((1.0f - f.w) * __ldg(tp + j) + f.w * __ldg(tn + j)) * vb[j];
#pragma unroll
                for (int j = 0; j < 6; ++j) acc += ((1.0f - f.w) * __ldg(rp + j) + f.w * __ldg(rn + j)) * vb[3 + j];
                s_vcoef[k * kDefStride] += acc;
            }
        }
#pragma unroll
        for (int j = 0; j < 9; ++j) s_vb[j * kDefStride + tid] = vb[j];
        __syncthreads();  // s_vb complete, s_red zeroed
        // camera-delta and time gradients: transposing butterfly over the warp, then one shared atomic per value
        warp_transpose_reduce_d<16>(r16, lane);
        if (lane < 16 && r16[0] != 0.f) atomicAdd(&s_red[lane], r16[0]);
        // basis gradients: thread (k, j) sums c_gk * vb_gj over the CTA's Gaussians
        for (int e = tid; e < K * 9; e += kDefThreads) {
            const int k = e / 9, jj = e - 9 * k;
            const float *cr = s_coef_all + k * kDefStride, *vr = s_vb + jj * kDefStride;
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll 8
            for (int t = 0; t < kDefThreads; t += 4) {
                a0 = fmaf(cr[t], vr[t], a0);
                a1 = fmaf(cr[t + 1], vr[t + 1], a1);
                a2 = fmaf(cr[t + 2], vr[t + 2], a2);
                a3 = fmaf(cr[t + 3], vr[t + 3], a3);
            }
            const float val = (a0 + a1) + (a2 + a3);
            if (val == 0.f) continue;
            // bases at the floor frame get (1 - w), at the ceil frame w
            if (jj < 3) {
                atomicAdd(v_transls + ((int64_t)k * T + f.pre) * 3 + jj, (1.0f - f.w) * val);
                atomicAdd(v_transls + ((int64_t)k * T + f.nxt) * 3 + jj, f.w * val);
            } else {
                atomicAdd(v_rots + ((int64_t)k * T + f.pre) * 6 + (jj - 3), (1.0f - f.w) * val);
                atomicAdd(v_rots + ((int64_t)k * T + f.nxt) * 6 + (jj - 3), f.w * val);
            }
        }
        __syncthreads();  // s_red complete; every (k, j) thread is done with s_vb
        if (tid < 12 && v_RTs && s_red[tid] != 0.f) atomicAdd(v_RTs + 12LL * n + tid, s_red[tid]);
        if (tid == 15 && s_red[15] != 0.f) atomicAdd(v_times + n, s_red[15]);
        __syncthreads();  // s_red read before the next timestamp zeroes it
    }
    if (active) {
        v_fg_means[3LL * g] = v_mu[0]; v_fg_means[3LL * g + 1] = v_mu[1]; v_fg_means[3LL * g + 2] = v_mu[2];
        reinterpret_cast<float4 *>(v_fg_quats)[g] = make_float4(v_q[0], v_q[1], v_q[2], v_q[3]);
        // softmax backward: v_raw_k = c_k (v_c_k - sum_m c_m v_c_m)
        float dotp = 0.f;
        for (int k = 0; k < K; ++k) dotp += s_coef[k * kDefStride] * s_vcoef[k * kDefStride];
        for (int k = 0; k < K; ++k)
            v_coefs_raw[(int64_t)g * K + k] = s_coef[k * kDefStride] * (s_vcoef[k * kDefStride] - dotp);
    }
}

__global__ void __launch_bounds__(kDefThreads)
deform_bg_bwd_kernel(const float *__restrict__ bg_means, const float *__restrict__ bg_quats,
                     const float *__restrict__ RTs, int Gf, int Gb, int G, int N,
                     const float *__restrict__ v_out_means, const float *__restrict__ v_out_quats,
                     float *__restrict__ v_bg_means, float *__restrict__ v_bg_quats, float *__restrict__ v_RTs) {
    __shared__ float s_red[16];
    const int tid = threadIdx.x, lane = tid & 31;
    const int g = blockIdx.x * kDefThreads + tid;
    const bool active = g < Gb;
    float mu[3] = {0.f, 0.f, 0.f};
    float4 q = make_float4(1.f, 0.f, 0.f, 0.f);
    if (active) {
        mu[0] = __ldg(bg_means + 3LL * g); mu[1] = __ldg(bg_means + 3LL * g + 1); mu[2] = __ldg(bg_means + 3LL * g + 2);
        q = __ldg(reinterpret_cast<const float4 *>(bg_quats) + g);
    }
    float v_mu[3] = {0.f, 0.f, 0.f};
    float4 vqh = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int n = 0; n < N; ++n) {
        if (tid < 16) s_red[tid] = 0.f;
        __syncthreads();
        float r16[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) r16[j] = 0.f;
        if (active) {
            const float *vmp = v_out_means + ((int64_t)n * G + Gf + g) * 3;
            const float vm[3] = {__ldg(vmp), __ldg(vmp + 1), __ldg(vmp + 2)};
            float vmu[3];
            rt_backward(RTs ? RTs + 12LL * n : nullptr, vm, mu, vmu, r16);
            v_mu[0] += vmu[0]; v_mu[1] += vmu[1]; v_mu[2] += vmu[2];
            const float4 vq = __ldg(reinterpret_cast<const float4 *>(v_out_quats) + (int64_t)n * G + Gf + g);
            vqh.x += vq.x; vqh.y += vq.y; vqh.z += vq.z; vqh.w += vq.w;
        }
        if (v_RTs && RTs) {
            warp_transpose_reduce_d<16>(r16, lane);
            if (lane < 12 && r16[0] != 0.f) atomicAdd(&s_red[lane], r16[0]);
            __syncthreads();
            if (tid < 12 && s_red[tid] != 0.f) atomicAdd(v_RTs + 12LL * n + tid, s_red[tid]);
        }
        __syncthreads();
    }
    if (active) {
        v_bg_means[3LL * g] = v_mu[0]; v_bg_means[3LL * g + 1] = v_mu[1]; v_bg_means[3LL * g + 2] = v_mu[2];
        // F.normalize backward: q_hat = q / max(|q|, eps)
        float nrm = sqrtf(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
        float4 out;
        if (nrm >= 1e-12f) {
            float inv = 1.0f / nrm;
            float4 qh = make_float4(q.x * inv, q.y * inv, q.z * inv, q.w * inv);
            float dotp = vqh.x * qh.x + vqh.y * qh.y + vqh.z * qh.z + vqh.w * qh.w;
            out = make_float4((vqh.x - dotp * qh.x) * inv, (vqh.y - dotp * qh.y) * inv, (vqh.z - dotp * qh.z) * inv,
                              (vqh.w - dotp * qh.w) * inv);
        } else {
            out = make_float4(vqh.x * 1e12f, vqh.y * 1e12f, vqh.z * 1e12f, vqh.w * 1e12f);
        }
        reinterpret_cast<float4 *>(v_bg_quats)[g] = out;
    }
}

// ------------------------------------------------------------------------------- compute_transforms
// MotionBases.compute_transforms (params.py:142-180) as a stand-alone op for its other callers
// (trainer.py:478,485,701; init_utils.py:322): coefs are ALREADY softmaxed, out [G,B,3,4] = [R | t].
__global__ void __launch_bounds__(kDefThreads)
transforms_fwd_kernel(const float *__restrict__ coefs, const float *__restrict__ rots,
                      const float *__restrict__ transls, const float *__restrict__ ts, int G, int K, int T, int B,
                      float *__restrict__ out) {
    extern __shared__ float s_coef_all[];
    const int g = blockIdx.x * kDefThreads + threadIdx.x;
    if (g >= G) return;
    float *s_coef = s_coef_all + threadIdx.x;
    for (int k = 0; k < K; ++k) s_coef[k * kDefThreads] = __ldg(coefs + (int64_t)g * K + k);
    for (int b = 0; b < B; ++b) {
        const FramePair f = frame_pair(__ldg(ts + b), T);
        float bp[9], bn[9], bl[9];
        blend_bases(s_coef, K, T, f.pre, rots, transls, bp);
        blend_bases(s_coef, K, T, f.nxt, rots, transls, bn);
#pragma unroll
        for (int j = 0; j < 9; ++j) bl[j] = (1.0f - f.w) * bp[j] + f.w * bn[j];
        float x[3], y[3], z[3];
        rot6d_to_cols<float>(bl + 3, x, y, z);
        float4 *o = reinterpret_cast<float4 *>(out + ((int64_t)g * B + b) * 12);
        o[0] = make_float4(x[0], y[0], z[0], bl[0]);
        o[1] = make_float4(x[1], y[1], z[1], bl[1]);
        o[2] = make_float4(x[2], y[2], z[2], bl[2]);
    }
}

__global__ void __launch_bounds__(kDefThreads)
transforms_bwd_kernel(const float *__restrict__ coefs, const float *__restrict__ rots,
                      const float *__restrict__ transls, const float *__restrict__ ts, int G, int K, int T, int B,
                      const float *__restrict__ v_out, float *__restrict__ v_coefs, float *__restrict__ v_rots,
                      float *__restrict__ v_transls, float *__restrict__ v_ts) {
    extern __shared__ float smem[];
    float *s_coef_all = smem;                     // [K][kDefThreads]
    float *s_vcoef_all = smem + K * kDefThreads;  // [K][kDefThreads]
    float *s_vB = smem + 2 * K * kDefThreads;     // [K][9]
    float *s_red = s_vB + K * 9;                  // [16], slot 15 = time gradient
    const int tid = threadIdx.x, lane = tid & 31;
    const int g = blockIdx.x * kDefThreads + tid;
    const bool active = g < G;
    float *s_coef = s_coef_all + tid, *s_vcoef = s_vcoef_all + tid;
    for (int k = 0; k < K; ++k) {
        s_coef[k * kDefThreads] = active ? __ldg(coefs + (int64_t)g * K + k) : 0.f;
        s_vcoef[k * kDefThreads] = 0.f;
    }
    for (int b = 0; b < B; ++b) {
        for (int e = tid; e < K * 9 + 16; e += kDefThreads) s_vB[e] = 0.f;
        __syncthreads();
        const FramePair f = frame_pair(__ldg(ts + b), T);
        float vb[9], r16[16];
#pragma unroll
        for (int j = 0; j < 9; ++j) vb[j] = 0.f;
#pragma unroll
        for (int j = 0; j < 16; ++j) r16[j] = 0.f;
        if (active) {
            float bp[9], bn[9], bl[9];
            blend_bases(s_coef, K, T, f.pre, rots, transls, bp);
            blend_bases(s_coef, K, T, f.nxt, rots, transls, bn);
#pragma unroll
            for (int j = 0; j < 9; ++j) bl[j] = (1.0f - f.w) * bp[j] + f.w * bn[j];
            const float4 *vo = reinterpret_cast<const float4 *>(v_out + ((int64_t)g * B + b) * 12);
            const float4 v0 = __ldg(vo), v1 = __ldg(vo + 1), v2 = __ldg(vo + 2);
            const float vx[3] = {v0.x, v1.x, v2.x}, vy[3] = {v0.y, v1.y, v2.y}, vz[3] = {v0.z, v1.z, v2.z};
            vb[0] = v0.w; vb[1] = v1.w; vb[2] = v2.w;
            typedef Dual<6> DU;
            DU r6[6], x[3], y[3], z[3];
#pragma unroll
            for (int a = 0; a < 6; ++a) {
                r6[a].v = bl[3 + a];
#pragma unroll
                for (int c = 0; c < 6; ++c) r6[a].d[c] = (a == c) ? 1.f : 0.f;
            }
            rot6d_to_cols<DU>(r6, x, y, z);
#pragma unroll
            for (int a = 0; a < 6; ++a) {
                float acc = 0.f;
#pragma unroll
                for (int i = 0; i < 3; ++i) acc += vx[i] * x[i].d[a] + vy[i] * y[i].d[a] + vz[i] * z[i].d[a];
                vb[3 + a] = acc;
            }
            float vw = 0.f;
#pragma unroll
            for (int j = 0; j < 9; ++j) vw += (bn[j] - bp[j]) * vb[j];
            r16[15] = vw;
            for (int k = 0; k < K; ++k) {
                const float *tp = transls + ((int64_t)k * T + f.pre) * 3, *tn = transls + ((int64_t)k * T + f.nxt) * 3;
                const float *rp = rots + ((int64_t)k * T + f.pre) * 6, *rn = rots + ((int64_t)k * T + f.nxt) * 6;
                float acc = 0.f;
#pragma unroll
                for (int j = 0; j < 3; ++j) acc += ((1.0f - f.w) * __ldg(tp + j) + f.w * __ldg(tn + j)) * vb[j];
#pragma unroll
                for (int j = 0; j < 6; ++j) acc += ((1.0f - f.w) * __ldg(rp + j) + f.w * __ldg(rn + j)) * vb[3 + j];
                s_vcoef[k * kDefThreads] += acc;
            }
        }
        warp_transpose_reduce_d<16>(r16, lane);
        if (lane == 15 && r16[0] != 0.f) atomicAdd(&s_red[15], r16[0]);
        for (int k = 0; k < K; ++k) {
            const float ck = s_coef[k * kDefThreads];
            float r[16];
#pragma unroll
            for (int j = 0; j < 9; ++j) r[j] = ck * vb[j];
#pragma unroll
            for (int j = 9; j < 16; ++j) r[j] = 0.f;
            warp_transpose_reduce_d<16>(r, lane);
            if (lane < 9 && r[0] != 0.f) atomicAdd(&s_vB[k * 9 + lane], r[0]);
        }
        __syncthreads();
        for (int e = tid; e < K * 9; e += kDefThreads) {
            const float val = s_vB[e];
            if (val == 0.f) continue;
            const int k = e / 9, jj = e - 9 * k;
            if (jj < 3) {
                atomicAdd(v_transls + ((int64_t)k * T + f.pre) * 3 + jj, (1.0f - f.w) * val);
                atomicAdd(v_transls + ((int64_t)k * T + f.nxt) * 3 + jj, f.w * val);
            } else {
                atomicAdd(v_rots + ((int64_t)k * T + f.pre) * 6 + (jj - 3), (1.0f - f.w) * val);
                atomicAdd(v_rots + ((int64_t)k * T + f.nxt) * 6 + (jj - 3), f.w * val);
            }
        }
        if (tid == 15 && s_red[15] != 0.f) atomicAdd(v_ts + b, s_red[15]);
        __syncthreads();
    }
    if (active)
        for (int k = 0; k < K; ++k) v_coefs[(int64_t)g * K + k] = s_vcoef[k * kDefThreads];
}

}  // namespace d4

using namespace d4;

extern "C" int d4_compute_transforms_fwd(const float *coefs, const float *rots, const float *transls, const float *ts,
                                         int G, int K, int T, int B, float *out, d4_stream_t stream) {
    D4_CHECK_ARG(G >= 0 && K >= 1 && K <= kMaxK && T >= 1 && B >= 1, "d4_compute_transforms_fwd: bad sizes");
    if (G == 0) return 0;
    D4_CHECK_ARG(coefs && rots && transls && ts && out && ((uintptr_t)out & 15) == 0,
                 "d4_compute_transforms_fwd: null/unaligned pointer");
    size_t smem = sizeof(float) * K * kDefThreads;
    transforms_fwd_kernel<<<cdiv(G, kDefThreads), kDefThreads, smem, as_stream(stream)>>>(coefs, rots, transls, ts, G, K,
                                                                                         T, B, out);
    D4_CHECK_LAUNCH("d4_compute_transforms_fwd");
    return 0;
}

extern "C" int d4_compute_transforms_bwd(const float *coefs, const float *rots, const float *transls, const float *ts,
                                         int G, int K, int T, int B, const float *v_out, float *v_coefs, float *v_rots,
                                         float *v_transls, float *v_ts, d4_stream_t stream) {
    D4_CHECK_ARG(G >= 0 && K >= 1 && K <= kMaxK && T >= 1 && B >= 1, "d4_compute_transforms_bwd: bad sizes");
    if (G == 0) return 0;
    D4_CHECK_ARG(coefs && rots && transls && ts && v_out && v_coefs && v_rots && v_transls && v_ts &&
                     ((uintptr_t)v_out & 15) == 0,
                 "d4_compute_transforms_bwd: null/unaligned pointer");
    size_t smem = sizeof(float) * (2 * K * kDefThreads + K * 9 + 16);
    if (smem > 48 * 1024)
        cudaFuncSetAttribute(transforms_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    transforms_bwd_kernel<<<cdiv(G, kDefThreads), kDefThreads, smem, as_stream(stream)>>>(
        coefs, rots, transls, ts, G, K, T, B, v_out, v_coefs, v_rots, v_transls, v_ts);
    D4_CHECK_LAUNCH("d4_compute_transforms_bwd");
    return 0;
}

static int check_deform(const char *name, int Gf, int Gb, int K, int T, int N) {
    D4_CHECK_ARG(Gf >= 0 && Gb >= 0 && N >= 1 && T >= 1, "%s: bad sizes", name);
    D4_CHECK_ARG(Gf == 0 || (K >= 1 && K <= kMaxK), "%s: K must be in [1,%d]", name, kMaxK);
    return 0;
}

extern "C" int d4_deform_fwd(const float *fg_means, const float *fg_quats, const float *motion_coefs,
                             const float *bg_means, const float *bg_quats, const float *rots,
                             const float *transls, const float *times, const float *RTs, int Gf, int Gb, int K,
                             int T, int N, float *out_means, float *out_quats, d4_stream_t stream) {
    if (int rc = check_deform("d4_deform_fwd", Gf, Gb, K, T, N)) return rc;
    D4_CHECK_ARG(out_means && out_quats && ((uintptr_t)out_quats & 15) == 0, "d4_deform_fwd: bad outputs");
    const int G = Gf + Gb;
    if (Gf > 0) {
        D4_CHECK_ARG(fg_means && fg_quats && motion_coefs && rots && transls && times && ((uintptr_t)fg_quats & 15) == 0,
                     "d4_deform_fwd: null/unaligned fg input");
        size_t smem = sizeof(float) * K * kDefThreads;
        if (smem > 48 * 1024)
            cudaFuncSetAttribute(deform_fg_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        deform_fg_fwd_kernel<<<cdiv(Gf, kDefThreads), kDefThreads, smem, as_stream(stream)>>>(
            fg_means, fg_quats, motion_coefs, rots, transls, times, RTs, Gf, G, K, T, N, out_means, out_quats);
    }
    if (Gb > 0) {
        D4_CHECK_ARG(bg_means && bg_quats && ((uintptr_t)bg_quats & 15) == 0, "d4_deform_fwd: null/unaligned bg input");
        deform_bg_fwd_kernel<<<cdiv(Gb, 256), 256, 0, as_stream(stream)>>>(bg_means, bg_quats, RTs, Gf, Gb, G, N,
                                                                            out_means, out_quats);
    }
    D4_CHECK_LAUNCH("d4_deform_fwd");
    return 0;
}

extern "C" int d4_deform_bwd(const float *fg_means, const float *fg_quats, const float *motion_coefs,
                             const float *bg_means, const float *bg_quats, const float *rots,
                             const float *transls, const float *times, const float *RTs, int Gf, int Gb, int K,
                             int T, int N, const float *v_out_means, const float *v_out_quats, float *v_fg_means,
                             float *v_fg_quats, float *v_motion_coefs, float *v_bg_means, float *v_bg_quats,
                             float *v_rots, float *v_transls, float *v_times, float *v_RTs, d4_stream_t stream) {
    if (int rc = check_deform("d4_deform_bwd", Gf, Gb, K, T, N)) return rc;
    D4_CHECK_ARG(v_out_means && v_out_quats && ((uintptr_t)v_out_quats & 15) == 0, "d4_deform_bwd: bad cotangents");
    const int G = Gf + Gb;
    if (Gf > 0) {
        D4_CHECK_ARG(fg_means && fg_quats && motion_coefs && rots && transls && times && v_fg_means && v_fg_quats &&
                         v_motion_coefs && v_rots && v_transls && v_times && ((uintptr_t)fg_quats & 15) == 0 &&
                         ((uintptr_t)v_fg_quats & 15) == 0,
                     "d4_deform_bwd: null/unaligned fg pointer");
        size_t smem = sizeof(float) * ((2 * K + 9) * kDefStride + 16);
        if (smem > 48 * 1024)
            cudaFuncSetAttribute(deform_fg_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        deform_fg_bwd_kernel<<<cdiv(Gf, kDefThreads), kDefThreads, smem, as_stream(stream)>>>(
            fg_means, fg_quats, motion_coefs, rots, transls, times, RTs, Gf, G, K, T, N, v_out_means, v_out_quats,
            v_fg_means, v_fg_quats, v_motion_coefs, v_rots, v_transls, v_times, v_RTs);
    }
    if (Gb > 0) {
        D4_CHECK_ARG(bg_means && bg_quats && v_bg_means && v_bg_quats && ((uintptr_t)bg_quats & 15) == 0 &&
                         ((uintptr_t)v_bg_quats & 15) == 0,
                     "d4_deform_bwd: null/unaligned bg pointer");
        deform_bg_bwd_kernel<<<cdiv(Gb, kDefThreads), kDefThreads, 0, as_stream(stream)>>>(
            bg_means, bg_quats, RTs, Gf, Gb, G, N, v_out_means, v_out_quats, v_bg_means, v_bg_quats, v_RTs);
    }
    D4_CHECK_LAUNCH("d4_deform_bwd");
    return 0;
}
