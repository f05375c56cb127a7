// Device helpers shared by the SIMT (pe_simt.cu) and tensor-core (pe_tc.cu) engines: tanh-layer jet chain rule,
// its adjoint, and the residual / seed stage of each loss kind.  Equations: SURVEY.md Appendix A.1/A.2;
// reference lines cited in pe_simt.cu.
#pragma once
#include "pe_common.cuh"

namespace pe_dev {

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float2 ld2(const float* p) { return *reinterpret_cast<const float2*>(p); }

template <int K> struct StreamTraits {
    static constexpr int KF = (K == 5) ? 3 : (K - 1);   // number of plain first-derivative streams (1..KF)
    static constexpr bool TT = (K == 5);                // stream 4 = second time derivative of stream 3
};

// branch-free tanh: odd Taylor polynomial to x^11 below |x| = 0.3, 1 - 2/(1 + 2^(2|x| log2 e)) above (MUFU.EX2 + MUFU.RCP).
// max relative error 2.6e-7 in exact fp32 arithmetic (+ ~1 ulp of the two approx units): fp32-class, and unlike tanhf() it
// has no data-dependent branches (a warp of collocation points always straddles tanhf's range split).
__device__ __forceinline__ float tanh_branchfree(float x) {
    const float ax = fabsf(x);
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(ax * 2.885390081777927f));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.0f));
    const float big = copysignf(fmaf(-2.0f, r, 1.0f), x);
    const float x2 = x * x;
    float p = -1382.0f / 155925.0f;
    p = fmaf(p, x2, 62.0f / 2835.0f);
    p = fmaf(p, x2, -17.0f / 315.0f);
    p = fmaf(p, x2, 2.0f / 15.0f);
    p = fmaf(p, x2, -1.0f / 3.0f);
    const float small = fmaf(x * x2, p, x);
    return ax < 0.3f ? small : big;
}

// ---- tanh layer, forward jets
template <int K, bool FAST_TANH = false>
__device__ __forceinline__ void act_fwd(float (&z)[K], float bias) {
    float a = FAST_TANH ? tanh_branchfree(z[0] + bias) : tanhf(z[0] + bias);
    float s = fmaf(-a, a, 1.f);
    float zt = (K == 5) ? z[3] : 0.f;
    z[0] = a;
#pragma unroll
    for (int k = 1; k <= StreamTraits<K>::KF; ++k) z[k] = s * z[k];
    if (K == 5) z[4] = fmaf(s, z[4], -2.f * a * z[3] * zt);      // a_tt = s z_tt - 2 a s z_t^2   (z[3] already = s z_t)
}

// ---- tanh layer, adjoint.  A = stashed outputs (a, a_x, ..), ab = adjoints of the outputs;
// returns adjoints of the pre-activations in ab.
template <int K>
__device__ __forceinline__ void act_bwd(float (&ab)[K], const float (&A)[K]) {
    float a = A[0];
    float s = fmaf(-a, a, 1.f);
    float acc = 0.f;
#pragma unroll
    for (int k = 1; k <= StreamTraits<K>::KF; ++k) acc = fmaf(A[k], ab[k], acc);     // s*z_k = A_k
    float zv = s * ab[0] - 2.f * a * acc;
    if (K == 5) {
        float inv_s = (s > 0.f) ? (1.f / s) : 0.f;
        float zt = A[3] * inv_s;                       // z_t
        float sztt = fmaf(2.f * a * A[3], zt, A[4]);   // s*z_tt = a_tt + 2 a s z_t^2
        zv = fmaf(-2.f * a * sztt, ab[4], zv);
        zv = fmaf(-2.f * fmaf(-3.f * a, a, 1.f) * A[3] * zt, ab[4], zv);
        float zb3 = fmaf(s, ab[3], -4.f * a * A[3] * ab[4]);
        ab[4] = s * ab[4];
        ab[3] = zb3;
        ab[1] = s * ab[1]; ab[2] = s * ab[2];
    } else {
#pragma unroll
        for (int k = 1; k <= StreamTraits<K>::KF; ++k) ab[k] = s * ab[k];
    }
    ab[0] = zv;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---- residual stage for one point: Y[k][o] -> loss partials + seeds Ybar[k][o] (in place)
template <int K>
__device__ __forceinline__ void residual_stage(float (&Y)[K][PE_UJ], const pe_term_desc& T, const float* __restrict__ aux_row,
                                               const float* __restrict__ row, bool valid, float inv_n,
                                               float (&tsum)[PE_MAX_TERMS]) {
    const float sc = valid ? 2.f * inv_n : 0.f;
    if (K == 5 && T.kind == PE_RES_F5) {
        // composite u = P + D*N  (plate:382-387), D,P jets precomputed (frozen nets)
        float D[5][5];
        if (T.aux_k) {
#pragma unroll
            for (int o = 0; o < 5; ++o) {
                float N[5], Pj[5];
#pragma unroll
                for (int k = 0; k < 5; ++k) { N[k] = Y[k][o]; D[k][o] = aux_row[k * 5 + o]; Pj[k] = aux_row[25 + k * 5 + o]; }
                Y[0][o] = fmaf(D[0][o], N[0], Pj[0]);
#pragma unroll
                for (int k = 1; k < 4; ++k) Y[k][o] = Pj[k] + D[k][o] * N[0] + D[0][o] * N[k];
                Y[4][o] = Pj[4] + D[4][o] * N[0] + 2.f * D[3][o] * N[3] + D[0][o] * N[4];
            }
        }
        const float E = T.E, mu = T.mu, rho = T.rho;
        const float c11 = E / (1.f - mu * mu), c12 = E * mu / (1.f - mu * mu), G = E / (2.f * (1.f + mu));
        float e11 = Y[1][0], e22 = Y[2][1], e12 = Y[2][0] + Y[1][1];
        float f_s11 = Y[0][2] - (c11 * e11 + c12 * e22);
        float f_s22 = Y[0][3] - (c12 * e11 + c11 * e22);
        float f_s12 = Y[0][4] - G * e12;
        float f_u = Y[1][2] + Y[2][4] - rho * Y[4][0];
        float f_v = Y[2][3] + Y[1][4] - rho * Y[4][1];
        if (valid) {
            tsum[0] += f_u * f_u + f_v * f_v;
            tsum[1] += f_s11 * f_s11 + f_s22 * f_s22 + f_s12 * f_s12;
        }
        float bu = sc * T.w[0] * f_u, bv = sc * T.w[0] * f_v;
        float b11 = sc * T.w[1] * f_s11, b22 = sc * T.w[1] * f_s22, b12 = sc * T.w[1] * f_s12;
#pragma unroll
        for (int k = 0; k < 5; ++k)
#pragma unroll
            for (int o = 0; o < PE_UJ; ++o) Y[k][o] = 0.f;
        Y[0][2] = b11; Y[0][3] = b22; Y[0][4] = b12;
        Y[1][0] = -(c11 * b11 + c12 * b22);
        Y[2][1] = -(c12 * b11 + c11 * b22);
        Y[2][0] = -G * b12; Y[1][1] = -G * b12;
        Y[1][2] = bu; Y[2][4] = bu; Y[4][0] = -rho * bu;
        Y[2][3] = bv; Y[1][4] = bv; Y[4][1] = -rho * bv;
        if (T.aux_k) {   // adjoint of the composite (linear in N jets)
#pragma unroll
            for (int o = 0; o < 5; ++o) {
                float u0 = Y[0][o], u1 = Y[1][o], u2 = Y[2][o], u3 = Y[3][o], u4 = Y[4][o];
                Y[0][o] = D[0][o] * u0 + D[1][o] * u1 + D[2][o] * u2 + D[3][o] * u3 + D[4][o] * u4;
                Y[1][o] = D[0][o] * u1;
                Y[2][o] = D[0][o] * u2;
                Y[3][o] = D[0][o] * u3 + 2.f * D[3][o] * u4;
                Y[4][o] = D[0][o] * u4;
            }
        }
    } else if (K == 4 && T.kind == PE_RES_F7) {
        const float E = T.E, mu = T.mu, rho = T.rho;
        const float coef = E / ((1.f + mu) * (1.f - 2.f * mu));
        const float c11 = coef * (1.f - mu), c12 = coef * mu, G = E / (2.f * (1.f + mu));
        float e11 = Y[1][0], e22 = Y[2][1], e12 = Y[2][0] + Y[1][1];
        float f_s11 = Y[0][4] - (c11 * e11 + c12 * e22);
        float f_s22 = Y[0][5] - (c12 * e11 + c11 * e22);
        float f_s12 = Y[0][6] - G * e12;
        float f_ut = Y[3][0] - Y[0][2];
        float f_vt = Y[3][1] - Y[0][3];
        float f_u = Y[1][4] + Y[2][6] - rho * Y[3][2];
        float f_v = Y[2][5] + Y[1][6] - rho * Y[3][3];
        if (valid) {
            tsum[0] += f_u * f_u + f_v * f_v + f_ut * f_ut + f_vt * f_vt;
            tsum[1] += f_s11 * f_s11 + f_s22 * f_s22 + f_s12 * f_s12;
        }
        float bu = sc * T.w[0] * f_u, bv = sc * T.w[0] * f_v, but = sc * T.w[0] * f_ut, bvt = sc * T.w[0] * f_vt;
        float b11 = sc * T.w[1] * f_s11, b22 = sc * T.w[1] * f_s22, b12 = sc * T.w[1] * f_s12;
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int o = 0; o < PE_UJ; ++o) Y[k][o] = 0.f;
        Y[0][4] = b11; Y[0][5] = b22; Y[0][6] = b12;
        Y[1][0] = -(c11 * b11 + c12 * b22);
        Y[2][1] = -(c12 * b11 + c11 * b22);
        Y[2][0] = -G * b12; Y[1][1] = -G * b12;
        Y[3][0] = but; Y[0][2] = -but;
        Y[3][1] = bvt; Y[0][3] = -bvt;
        Y[1][4] = bu; Y[2][6] = bu; Y[3][2] = -rho * bu;
        Y[2][5] = bv; Y[1][6] = bv; Y[3][3] = -rho * bv;
    } else if (K == 1 && T.kind == PE_RES_TRACTION) {
        float D0[5];
        if (T.aux_k) {
#pragma unroll
            for (int o = 0; o < 5; ++o) { D0[o] = aux_row[o]; Y[0][o] = fmaf(D0[o], Y[0][o], aux_row[5 + o]); }
        }
        float nx = -row[0] / T.hole_r, ny = -row[1] / T.hole_r;
        float tx = Y[0][2] * nx + Y[0][4] * ny;
        float ty = Y[0][4] * nx + Y[0][3] * ny;
        if (valid) tsum[0] += tx * tx + ty * ty;
        float btx = sc * T.w[0] * tx, bty = sc * T.w[0] * ty;
#pragma unroll
        for (int o = 0; o < PE_UJ; ++o) Y[0][o] = 0.f;
        Y[0][2] = btx * nx; Y[0][3] = bty * ny; Y[0][4] = btx * ny + bty * nx;
        if (T.aux_k) {
#pragma unroll
            for (int o = 0; o < 5; ++o) Y[0][o] *= D0[o];
        }
    } else if ((K == 1 && T.kind == PE_RES_COLS) || (K == 2 && T.kind == PE_RES_DT)) {
        constexpr int ks = (K == 2) ? 1 : 0;
        float seed[PE_UJ];
#pragma unroll
        for (int o = 0; o < PE_UJ; ++o) seed[o] = 0.f;
#pragma unroll
        for (int c = 0; c < PE_MAX_COLS; ++c) {
            if (c < T.ncols) {
                const int col = T.col[c];
                float tg = (T.tgt[c] >= 0) ? row[T.tgt[c]] : 0.f;
                float val = 0.f;
#pragma unroll
                for (int o = 0; o < PE_UJ; ++o) if (o == col) val = Y[ks][o];
                float r = val - tg;
                if (valid) tsum[c] += r * r;
                float b = sc * T.w[c] * r;
#pragma unroll
                for (int o = 0; o < PE_UJ; ++o) if (o == col) seed[o] += b;
            }
        }
#pragma unroll
        for (int k = 0; k < K; ++k)
#pragma unroll
            for (int o = 0; o < PE_UJ; ++o) Y[k][o] = (k == ks) ? seed[o] : 0.f;
    }
}


}  // namespace pe_dev
