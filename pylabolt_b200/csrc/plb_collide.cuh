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
#include <cmath>

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

// 1 / x for the MRT paths (OUR DEFINITION, not bit-compared with upstream
// code): the hardware seed (MUFU.RCP64H, ~2^-20) and two Newton steps -- four
// DFMA, an error of an ulp, and none of the range checks and the slow-path
// call an IEEE division carries (x = rho + eps is of order one).  The BGK
// paths, which are pinned to the reference bit for bit, keep the division.
#ifndef PLB_MRT_RCP_NEWTON
#define PLB_MRT_RCP_NEWTON 1
#endif
// PLB_MRT_PAIR_FMA3=1 forms a direction pair with three FMAs instead of two
// adds and two FMAs (one fp64 instruction less per pair, +1 % GLUPS) -- and
// doubles the distance to the oracle on the channel benchmark (1.62e-12
// against 7.2e-13 relative after 400 steps, bench.py's in-run parity on the
// B200): off.  The Newton reciprocal changes that distance in no digit.
#ifndef PLB_MRT_PAIR_FMA3
#define PLB_MRT_PAIR_FMA3 0
#endif
__device__ __forceinline__ double rcp_newton(double x)
{
#if defined(PLB_EMU_RUNTIME) || !PLB_MRT_RCP_NEWTON
    return 1.0 / x;
#else
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    return fma(r, e, r);
#endif
}

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

// MRT with nine free rates -- OUR DEFINITION (no upstream kernel; SURVEY.md
// App. A.2):
//   g = f - Minv diag(S) M (f - feq) + Minv (I - diag(S)/2) M Phi
// evaluated in MOMENT space: m = M f is formed once (it also yields rho and
// the momentum), the moments of feq and of the Guo source Phi are low-order
// polynomials in (rho, u, F) with host-computed coefficients (MrtMoments), and
//   g = f + Minv [ S (m_eq - m) + (I - S/2) m_Phi ].
// ~165 fp64 operations per node with second-order Guo forcing; the
// population-space form (nine feq, nine Phi, two forward transforms) needed
// 304 and was fp64-bound.  The 9 x 9 transforms stay unrolled in registers.
template <int FORCING>
__device__ __forceinline__ Moments collide_mrt_moments(const KParams &p,
                                                       const double f[Q],
                                                       double g[Q])
{
    const MrtMoments &t = p.mrtm;
    double mom[Q], c[Q];
    mrt_forward(f, mom);
    Moments m;
    m.rho = mom[0];
    m.fx = m.rho * p.gx;
    m.fy = m.rho * p.gy;
    const double inv = rcp_newton(m.rho + p.eps);
    // u = (sum_k c_k f_k + F/2) / rho, cpu/compute_fields_kernels.py:55-65
    m.ux = (mom[3] + 0.5 * m.fx) * inv;
    m.uy = (mom[5] + 0.5 * m.fy) * inv;
    const double ux = m.ux, uy = m.uy, rho = m.rho;
    const double uxx = ux * ux, uyy = uy * uy;
    const double jx = rho * ux, jy = rho * uy;
#pragma unroll
    for (int r = 0; r < 3; ++r)
        c[r] = t.sn[r] * (rho * (t.A[r] + t.Qx[r] * uxx + t.Qy[r] * uyy) - mom[r]);
    c[3] = t.sn[3] * (t.B[0] * jx - mom[3]);
    c[4] = t.sn[4] * (t.B[1] * jx - mom[4]);
    c[5] = t.sn[5] * (t.B[2] * jy - mom[5]);
    c[6] = t.sn[6] * (t.B[3] * jy - mom[6]);
    c[7] = t.sn[7] * (rho * (t.Q7x * uxx + t.Q7y * uyy) - mom[7]);
    c[8] = t.sn[8] * (t.Q8 * (jx * uy) - mom[8]);
    if constexpr (FORCING == 1) {
        // linear Guo source w_k (c_k . F) / cs^2: momentum-like moments only
        c[3] += t.hn[3] * (t.B[0] * m.fx);
        c[4] += t.hn[4] * (t.B[1] * m.fx);
        c[5] += t.hn[5] * (t.B[2] * m.fy);
        c[6] += t.hn[6] * (t.B[3] * m.fy);
    } else if constexpr (FORCING == 2) {
        const double xfx = ux * m.fx, yfy = uy * m.fy;
        const double uf = xfx + yfy;
#pragma unroll
        for (int r = 0; r < 3; ++r)
            c[r] += t.hn[r] * (t.Gx[r] * xfx + t.Gy[r] * yfy - t.GA[r] * uf);
        c[3] += t.hn[3] * (t.B[0] * m.fx);
        c[4] += t.hn[4] * (t.B[1] * m.fx);
        c[5] += t.hn[5] * (t.B[2] * m.fy);
        c[6] += t.hn[6] * (t.B[3] * m.fy);
        c[7] += t.hn[7] * (2.0 * (t.Q7x * xfx + t.Q7y * yfy));
        c[8] += t.hn[8] * (t.Q8 * (ux * m.fy + uy * m.fx));
    }
    double dv[Q];
    mrt_backward(c, dv);
#pragma unroll
    for (int k = 0; k < Q; ++k) g[k] = f[k] + dv[k];
    return m;
}

// MRT with the reference's rates S = (1, 1, 1, 1, 1, 1, 1, s7, s8)
// (base/collision_operator.py:159-163) -- OUR DEFINITION like mrt_all(), and
// algebraically identical to it for these rates (checked against the oracle's
// full-matrix form to 1e-12).  With P_r the projector on stress moment r
// (rows 7, 8 of M, squared norm 4):
//   g = feq + Phi/2 + sum_r (1 - s_r) P_r h,      h = f - feq + Phi/2
// so only the two stress moments of h are ever formed,
//   A = row7 . h,  B = row8 . h,   g_k = (feq + Phi/2)_k + qa row7_k A + qb row8_k B
// and the stress moments of feq and Phi are known in closed form (isotropy of
// the D2Q9 weights; F = rho g is the reference's only force,
// cpu/force_field_kernels.py:23-25):
//   row7 . feq = k7 rho (ux^2 - uy^2)        row7 . Phi/2 = k7 rho (ux gx - uy gy)
//   row8 . feq = k8 rho ux uy                row8 . Phi/2 = k8 rho (ux gy + uy gx)/2
// (second-order Guo; the linear Guo source has no stress moment), with
// k7 = w_1/cs^4 and k8 = 4 w_5/cs^4 (both 1 up to the rounding of the
// reference's cs).  feq + Phi/2 of a direction K and of its opposite share
// their even part (c -> -c flips only the terms that are odd in c):
//   (feq + Phi/2)_{K, inv K} = w rho (E +- O)
//   E = 1 - (u.u + u.g)/(2 cs^2) + (c.u)(c.u + c.g)/(2 cs^4)
//   O = (c.u + c.g/2)/cs^2
// 84 fp64 operations per node with second-order Guo forcing (the generic
// nine-rate transform needs 304), no per-direction temporaries kept live.
template <int K, int FORCING>
__device__ __forceinline__ void stress_pair(const KParams &p, double ux,
                                            double uy, double wrho, double base,
                                            double proj, double g[Q])
{
    constexpr int KI = d_inv[K];
    constexpr int slot = (K == 1) ? 0 : (K == 2) ? 1 : (K == 5) ? 2 : 3;
    static_assert(K == 1 || K == 2 || K == 5 || K == 8, "pair representative");
    const double cu = cdot<K>(ux, uy);
    double even, odd;
    if constexpr (FORCING == 2) {
        even = cu * (p.mrt.k4 * cu + p.mrt.k4cg[slot]) + base;
    } else {
        even = (p.mrt.k4 * cu) * cu + base;
    }
    if constexpr (FORCING != 0) odd = p.inv_cs_2 * cu + p.mrt.k2cg[slot];
    else odd = p.inv_cs_2 * cu;
#if PLB_MRT_PAIR_FMA3
    // wrho (even +- odd) + proj as three fused multiply-adds
    const double centre = fma(wrho, even, proj);
    g[K] = fma(wrho, odd, centre);
    g[KI] = fma(-wrho, odd, centre);
#else
    g[K] = wrho * (even + odd) + proj;
    g[KI] = wrho * (even - odd) + proj;
#endif
}

template <int FORCING>
__device__ __forceinline__ Moments collide_mrt_stress(const KParams &p,
                                                      const double f[Q],
                                                      double g[Q])
{
    // pair sums / differences shared by rho, the momentum and the stresses
    const double p13 = f[1] + f[3], p24 = f[2] + f[4];
    const double p57 = f[5] + f[7], p68 = f[6] + f[8];
    const double d13 = f[1] - f[3], d24 = f[2] - f[4];
    const double d57 = f[5] - f[7], d86 = f[8] - f[6];
    Moments m;
    m.rho = ((f[0] + p13) + p24) + (p57 + p68);
    m.fx = m.rho * p.gx;
    m.fy = m.rho * p.gy;
    const double inv = rcp_newton(m.rho + p.eps);
    // u = (sum_k c_k f_k + F/2) / rho, cpu/compute_fields_kernels.py:55-65
    m.ux = (((d13 + d57) + d86) + m.rho * p.mrt.hgx) * inv;
    m.uy = (((d24 + d57) - d86) + m.rho * p.mrt.hgy) * inv;
    const double ux = m.ux, uy = m.uy;

    double base, sa, sb;
    if constexpr (FORCING == 2) {
        base = 1.0 - p.mrt.k2 * ((ux * ux + uy * uy) + (ux * p.gx + uy * p.gy));
        sa = ux * (ux - p.gx) - uy * (uy - p.gy);
        sb = ux * (uy - p.mrt.hgy) - uy * p.mrt.hgx;
    } else {
        base = 1.0 - p.mrt.k2 * (ux * ux + uy * uy);
        sa = ux * ux - uy * uy;
        sb = ux * uy;
    }
    const double A = (p13 - p24) - (p.mrt.k7 * m.rho) * sa;
    const double B = (p57 - p68) - (p.mrt.k8 * m.rho) * sb;
    const double a = p.mrt.qa * A, b = p.mrt.qb * B;

    const double w1rho = p.w[1] * m.rho, w5rho = p.w[5] * m.rho;
    g[0] = (p.w[0] * m.rho) * base;
    stress_pair<1, FORCING>(p, ux, uy, w1rho, base, a, g);
    stress_pair<2, FORCING>(p, ux, uy, w1rho, base, -a, g);
    stress_pair<5, FORCING>(p, ux, uy, w5rho, base, b, g);
    stress_pair<8, FORCING>(p, ux, uy, w5rho, base, -b, g);
    return m;
}

// Phases 2-5 for one node: moments, then post-collision populations g.
// COLL: 0 BGK, 1 MRT with nine free rates, 2 MRT with the reference's rates.
template <int COLL, int FORCING>
__device__ __forceinline__ Moments collide(const KParams &p, const double f[Q],
                                           double g[Q])
{
    if constexpr (COLL == 2) {
        return collide_mrt_stress<FORCING>(p, f, g);
    } else if constexpr (COLL == 1) {
        return collide_mrt_moments<FORCING>(p, f, g);
    } else {
        const Moments m = moments(p, f);
        const double u2 = m.ux * m.ux + m.uy * m.uy;
        bgk_all<FORCING>(p, m, u2, f, g);
        return m;
    }
}

}  // namespace plb
