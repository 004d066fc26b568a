// plb_collide.cuh -- per-node arithmetic of the fluidLB step, in registers.
//
// Every expression keeps the operand order of the reference's numba kernels so
// that a build with -fmad=false (libplb_strict.so) is bit-identical to the
// reference on BGK paths; the production build lets ptxas contract a*b+c into
// DFMA, which is the only licensed difference (<= 1e-12 relative).
//
// Integer lattice velocities are folded at compile time: the reference
// multiplies by float(cx[k]) in {-1, 0, 1}; x*1 == x, x*(-1) == -x and adding
// a signed zero are exact, so dropping those operations changes no bit (only,
// possibly, the sign of an exact zero).
#pragma once
#include "plb_internal.h"

namespace plb {

__device__ constexpr int d_cx[Q] = PLB_CX_LIST;
__device__ constexpr int d_cy[Q] = PLB_CY_LIST;
__device__ constexpr int d_inv[Q] = PLB_INV_LIST;

// float(c) * v for c in {-1, 0, 1}
template <int C>
__device__ __forceinline__ double mulc(double v)
{
    if constexpr (C == 0) return 0.0;
    else if constexpr (C == 1) return v;
    else return -v;
}

// float(cx)*a + float(cy)*b  (e.g. cu = cx[k]*ux + cy[k]*uy,
// pylabolt/parallel/cpu/collision_kernels.py:36)
template <int K>
__device__ __forceinline__ double cdot(double a, double b)
{
    constexpr int cx = d_cx[K], cy = d_cy[K];
    if constexpr (cx == 0 && cy == 0) return 0.0;
    else if constexpr (cy == 0) return mulc<cx>(a);
    else if constexpr (cx == 0) return mulc<cy>(b);
    else return mulc<cx>(a) + mulc<cy>(b);
}

struct Moments {
    double rho, ux, uy, fx, fy;
};

// Phases 2-4 of Solver.single_time_step (pylabolt/solvers/fluidLB.py:211-224):
//   rho = sum_k f_k, k = 0..8 in order   (cpu/compute_fields_kernels.py:24-28)
//   F   = rho * g                        (cpu/force_field_kernels.py:23-25)
//   u   = (sum_k c_k f_k) * inv + 0.5 * F * inv, inv = 1/(rho + eps)
//                                        (cpu/compute_fields_kernels.py:55-65)
// The gravity force is computed even when forcing is None (g = 0 then).
__device__ __forceinline__ Moments moments(const KParams &p, const double f[Q])
{
    Moments m;
    double rho = f[0];
    rho += f[1]; rho += f[2]; rho += f[3]; rho += f[4];
    rho += f[5]; rho += f[6]; rho += f[7]; rho += f[8];
    double sx = f[1];
    sx -= f[3]; sx += f[5]; sx -= f[6]; sx -= f[7]; sx += f[8];
    double sy = f[2];
    sy -= f[4]; sy += f[5]; sy += f[6]; sy -= f[7]; sy -= f[8];
    m.rho = rho;
    m.fx = rho * p.gx;
    m.fy = rho * p.gy;
    double inv = 1.0 / (rho + p.eps);
    m.ux = sx * inv + 0.5 * m.fx * inv;
    m.uy = sy * inv + 0.5 * m.fy * inv;
    return m;
}

// Second-order equilibrium of direction K,
// pylabolt/parallel/cpu/equilibrium_kernels.py:26-35:
//   w_k * rho * (1 + inv_cs_2*cu + 0.5*inv_cs_4*cu*cu - 0.5*inv_cs_2*u2)
template <int K>
__device__ __forceinline__ double feq(const KParams &p, const Moments &m,
                                      double u2)
{
    const double wrho = p.w[K] * m.rho;
    const double c = 0.5 * p.inv_cs_2 * u2;
    if constexpr (K == 0) {
        return wrho * (1.0 - c);
    } else {
        const double cu = cdot<K>(m.ux, m.uy);
        return wrho * (1.0 + p.inv_cs_2 * cu + 0.5 * p.inv_cs_4 * cu * cu - c);
    }
}

// Guo source of direction K without the (1 - omega/2) prefactor.
//   guo_linear       cpu/collision_kernels.py:90-92
//   guo_second_order cpu/collision_kernels.py:142-148
template <int K, int FORCING>
__device__ __forceinline__ double guo(const KParams &p, const Moments &m)
{
    constexpr int cx = d_cx[K], cy = d_cy[K];
    if constexpr (FORCING == 1) {
        if constexpr (K == 0) return 0.0;
        else return p.w[K] * cdot<K>(m.fx, m.fy) * p.inv_cs_2;
    } else {
        const double cu = cdot<K>(m.ux, m.uy);
        double const_x = (double(cx) - m.ux) * p.inv_cs_2;
        double const_y = (double(cy) - m.uy) * p.inv_cs_2;
        if constexpr (cx != 0) const_x = const_x + mulc<cx>(cu) * p.inv_cs_4;
        if constexpr (cy != 0) const_y = const_y + mulc<cy>(cu) * p.inv_cs_4;
        return p.w[K] * (const_x * m.fx + const_y * m.fy);
    }
}

// BGK relaxation of direction K, cpu/collision_kernels.py:45, :93-97, :149-153
template <int K, int FORCING>
__device__ __forceinline__ double bgk(const KParams &p, const Moments &m,
                                      double u2, double fk)
{
    const double e = feq<K>(p, m, u2);
    if constexpr (FORCING == 0) {
        return (1.0 - p.omega) * fk + p.omega * e;
    } else if constexpr (FORCING == 1 && K == 0) {
        return (1.0 - p.omega) * fk + p.omega * e;   // source is an exact zero
    } else {
        return (1.0 - p.omega) * fk + p.omega * e +
               (1.0 - 0.5 * p.omega) * guo<K, FORCING>(p, m);
    }
}

template <int FORCING, int K = 0>
__device__ __forceinline__ void bgk_all(const KParams &p, const Moments &m,
                                        double u2, const double f[Q],
                                        double g[Q])
{
    g[K] = bgk<K, FORCING>(p, m, u2, f[K]);
    if constexpr (K + 1 < Q) bgk_all<FORCING, K + 1>(p, m, u2, f, g);
}

template <int K = 0>
__device__ __forceinline__ void feq_all(const KParams &p, const Moments &m,
                                        double u2, double e[Q])
{
    e[K] = feq<K>(p, m, u2);
    if constexpr (K + 1 < Q) feq_all<K + 1>(p, m, u2, e);
}

template <int FORCING, int K = 0>
__device__ __forceinline__ void guo_all(const KParams &p, const Moments &m,
                                        double phi[Q])
{
    phi[K] = guo<K, FORCING>(p, m);
    if constexpr (K + 1 < Q) guo_all<FORCING, K + 1>(p, m, phi);
}

// M v for the Lallemand-Luo matrix of base/collision_operator.py:147-157
// (rows rho, e, eps, jx, qx, jy, qy, pxx, pxy), unrolled with shared partial
// sums -- a 9x9 contraction per node stays in registers, never a GEMM.
__device__ __forceinline__ void mrt_forward(const double v[Q], double m[Q])
{
    const double a = (v[1] + v[3]) + (v[2] + v[4]);   // axis sum
    const double b = (v[5] + v[7]) + (v[6] + v[8]);   // diagonal sum
    const double ax = v[1] - v[3], ay = v[2] - v[4];
    const double dx = (v[5] - v[6]) + (v[8] - v[7]);
    const double dy = (v[5] + v[6]) - (v[7] + v[8]);
    m[0] = v[0] + a + b;
    m[1] = -4.0 * v[0] - a + 2.0 * b;
    m[2] = 4.0 * v[0] - 2.0 * a + b;
    m[3] = ax + dx;
    m[4] = dx - 2.0 * ax;
    m[5] = ay + dy;
    m[6] = dy - 2.0 * ay;
    m[7] = (v[1] + v[3]) - (v[2] + v[4]);
    m[8] = (v[5] + v[7]) - (v[6] + v[8]);
}

// M^T c (the rows of M are orthogonal, inv(M) = M^T diag(1/|row|^2); the
// caller has already divided c by the squared row norms).
__device__ __forceinline__ void mrt_backward(const double c[Q], double v[Q])
{
    v[0] = c[0] - 4.0 * c[1] + 4.0 * c[2];
    const double ax = c[0] - c[1] - 2.0 * c[2];   // common part of k = 1..4
    const double dg = c[0] + 2.0 * c[1] + c[2];   // common part of k = 5..8
    const double jx = c[3] - 2.0 * c[4], jy = c[5] - 2.0 * c[6];
    const double qx = c[3] + c[4], qy = c[5] + c[6];
    v[1] = ax + jx + c[7];
    v[3] = ax - jx + c[7];
    v[2] = ax + jy - c[7];
    v[4] = ax - jy - c[7];
    v[5] = dg + qx + qy + c[8];
    v[6] = dg - qx + qy - c[8];
    v[7] = dg - qx - qy + c[8];
    v[8] = dg + qx - qy - c[8];
}

// MRT -- OUR DEFINITION (no upstream kernel):
//   g = f - Minv diag(S) M (f - feq) + Minv (I - diag(S)/2) M Phi
template <int FORCING>
__device__ __forceinline__ void mrt_all(const KParams &p, const Moments &m,
                                        double u2, const double f[Q],
                                        double g[Q])
{
    constexpr double inv_norm2[Q] = {1.0 / 9,  1.0 / 36, 1.0 / 36,
                                     1.0 / 6,  1.0 / 12, 1.0 / 6,
                                     1.0 / 12, 1.0 / 4,  1.0 / 4};
    double e[Q], fneq[Q], mom[Q], c[Q];
    feq_all(p, m, u2, e);
#pragma unroll
    for (int k = 0; k < Q; ++k) fneq[k] = f[k] - e[k];
    mrt_forward(fneq, mom);
#pragma unroll
    for (int r = 0; r < Q; ++r) c[r] = -(p.s[r] * inv_norm2[r]) * mom[r];
    if constexpr (FORCING != 0) {
        double phi[Q], mphi[Q];
        guo_all<FORCING>(p, m, phi);
        mrt_forward(phi, mphi);
#pragma unroll
        for (int r = 0; r < Q; ++r)
            c[r] += ((1.0 - 0.5 * p.s[r]) * inv_norm2[r]) * mphi[r];
    }
    double dv[Q];
    mrt_backward(c, dv);
#pragma unroll
    for (int k = 0; k < Q; ++k) g[k] = f[k] + dv[k];
}

// Guo source for the MRT paths.  MRT has no upstream arithmetic to preserve,
// so the second-order bracket is evaluated in the cheaper, algebraically
// identical form  w_k * (inv_cs_2 * (c_k.F - u.F) + inv_cs_4 * (c_k.u)(c_k.F)).
template <int K, int FORCING>
__device__ __forceinline__ double guo_mrt(const KParams &p, const Moments &m,
                                          double uF)
{
    if constexpr (FORCING == 1) {
        return guo<K, 1>(p, m);
    } else if constexpr (K == 0) {
        return -(p.w[0] * p.inv_cs_2) * uF;
    } else {
        const double cF = cdot<K>(m.fx, m.fy);
        const double cu = cdot<K>(m.ux, m.uy);
        return p.w[K] * (p.inv_cs_2 * (cF - uF) + p.inv_cs_4 * (cu * cF));
    }
}

template <int FORCING, int K = 0>
__device__ __forceinline__ void guo_mrt_all(const KParams &p, const Moments &m,
                                            double uF, double phi[Q])
{
    phi[K] = guo_mrt<K, FORCING>(p, m, uF);
    if constexpr (K + 1 < Q) guo_mrt_all<FORCING, K + 1>(p, m, uF, phi);
}

// MRT with the reference's rates S = (1, 1, 1, 1, 1, 1, 1, s7, s8)
// (base/collision_operator.py:159-163).  With P the projector on the two
// stress moments (rows 7, 8 of M, squared norm 4):
//   g = feq + Phi/2 + Minv diag(0..0, 1-s7, 1-s8) M (f - feq + Phi/2)
// so only two non-equilibrium moments are ever formed:
//   A = row7 . h,  B = row8 . h,  h = f - feq + Phi/2
//   g_k = feq_k + Phi_k/2 + (1-s7)/4 * row7_k * A + (1-s8)/4 * row8_k * B
// Algebraically identical to mrt_all() for these rates (checked against the
// oracle's full-matrix form to 1e-12).
// feq and Phi/2 of a direction K and of its opposite inv(K) share their even
// part (c -> -c flips only the terms that are odd in c):
//   feq_{K,inv}  = w rho ((1 - u2/(2cs2) + (c.u)^2/(2cs4)) +- (c.u)/cs2)
//   Phi_{K,inv}  = w ((c.u)(c.F)/cs4 - u.F/cs2 +- (c.F)/cs2)
// MRT only (our definition); the BGK paths keep the reference's operand order.
template <int K, int FORCING>
__device__ __forceinline__ void mrt_pair(const KParams &p, const Moments &m,
                                         double base, double uF, double e[Q],
                                         double half_phi[Q])
{
    constexpr int KI = d_inv[K];
    const double cu = cdot<K>(m.ux, m.uy);
    const double wrho = p.w[K] * m.rho;
    const double even = base + (0.5 * p.inv_cs_4 * cu) * cu;
    const double odd = p.inv_cs_2 * cu;
    e[K] = wrho * (even + odd);
    e[KI] = wrho * (even - odd);
    if constexpr (FORCING != 0) {
        const double cF = cdot<K>(m.fx, m.fy);
        const double hw = 0.5 * p.w[K];
        const double r = p.inv_cs_2 * cF;
        if constexpr (FORCING == 1) {
            half_phi[K] = hw * r;
            half_phi[KI] = -(hw * r);
        } else {
            const double s = p.inv_cs_4 * (cu * cF) - p.inv_cs_2 * uF;
            half_phi[K] = hw * (s + r);
            half_phi[KI] = hw * (s - r);
        }
    }
}

template <int FORCING>
__device__ __forceinline__ void mrt_reduced_all(const KParams &p,
                                                const Moments &m, double u2,
                                                const double f[Q], double g[Q])
{
    double e[Q], half_phi[Q];
    const double base = 1.0 - 0.5 * p.inv_cs_2 * u2;
    const double uF = m.ux * m.fx + m.uy * m.fy;
    e[0] = (p.w[0] * m.rho) * base;
    half_phi[0] = (FORCING == 2) ? -(0.5 * p.w[0] * p.inv_cs_2) * uF : 0.0;
    mrt_pair<1, FORCING>(p, m, base, uF, e, half_phi);
    mrt_pair<2, FORCING>(p, m, base, uF, e, half_phi);
    mrt_pair<5, FORCING>(p, m, base, uF, e, half_phi);
    mrt_pair<8, FORCING>(p, m, base, uF, e, half_phi);
    if constexpr (FORCING != 0) {
#pragma unroll
        for (int k = 0; k < Q; ++k) {
            g[k] = e[k] + half_phi[k];      // feq + Phi/2
            e[k] = e[k] - half_phi[k];      // so that h = f - e
        }
    } else {
#pragma unroll
        for (int k = 0; k < Q; ++k) g[k] = e[k];
    }
    const double A = ((f[1] - e[1]) + (f[3] - e[3])) -
                     ((f[2] - e[2]) + (f[4] - e[4]));
    const double B = ((f[5] - e[5]) + (f[7] - e[7])) -
                     ((f[6] - e[6]) + (f[8] - e[8]));
    const double a = (0.25 * (1.0 - p.s[7])) * A;
    const double b = (0.25 * (1.0 - p.s[8])) * B;
    g[1] += a; g[3] += a; g[2] -= a; g[4] -= a;
    g[5] += b; g[7] += b; g[6] -= b; g[8] -= b;
}

// Phases 2-5 for one node: moments, then post-collision populations g.
// COLL: 0 BGK, 1 MRT with nine free rates, 2 MRT with the reference's rates.
template <int COLL, int FORCING>
__device__ __forceinline__ Moments collide(const KParams &p, const double f[Q],
                                           double g[Q])
{
    const Moments m = moments(p, f);
    const double u2 = m.ux * m.ux + m.uy * m.uy;
    if constexpr (COLL == 0) bgk_all<FORCING>(p, m, u2, f, g);
    else if constexpr (COLL == 1) mrt_all<FORCING>(p, m, u2, f, g);
    else mrt_reduced_all<FORCING>(p, m, u2, f, g);
    return m;
}

}  // namespace plb
