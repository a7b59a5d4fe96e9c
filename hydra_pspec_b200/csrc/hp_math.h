// hp_math.h -- scalar math shared by device kernels and (for CPU unit tests) host code.
//
// Everything here is HP_HD (host+device) and header-only so that tests/ can compile the very
// same source with g++ (csrc/hp_math_host.cpp) and check it against scipy without a GPU.
//
//  * philox4x32-10 counter RNG, 53-bit uniforms, Box-Muller normals, Marsaglia-Tsang gamma
//  * regularised upper incomplete gamma Q(a,x)   (scipy.special.gammaincc; used by
//    hydra_pspec/pspec.py:51 through invgamma.cdf)
//  * inversion sampler on a log grid            (hydra_pspec/pspec.py:11-64)
//  * scalar model of the reference's preconditioned CG (hydra_pspec/pspec.py:228)
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define HP_HD __host__ __device__ __forceinline__
#else
#define HP_HD inline
#endif

namespace hp {

// ------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al. 2011).  counter = 128 bit, key = 64 bit.
struct u32x4 { uint32_t x, y, z, w; };

HP_HD void philox_mulhilo(uint32_t a, uint32_t b, uint32_t& hi, uint32_t& lo) {
#if defined(__CUDA_ARCH__)
    lo = a * b;
    hi = __umulhi(a, b);
#else
    uint64_t p = (uint64_t)a * (uint64_t)b;
    lo = (uint32_t)p;
    hi = (uint32_t)(p >> 32);
#endif
}

HP_HD u32x4 philox4x32_10(u32x4 ctr, uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0, lo0, hi1, lo1;
        philox_mulhilo(M0, ctr.x, hi0, lo0);
        philox_mulhilo(M1, ctr.z, hi1, lo1);
        u32x4 n;
        n.x = hi1 ^ ctr.y ^ k0;
        n.y = lo1;
        n.z = hi0 ^ ctr.w ^ k1;
        n.w = lo0;
        ctr = n;
        k0 += W0;
        k1 += W1;
    }
    return ctr;
}

// [0,1) with 53 random bits
HP_HD double u53(uint32_t hi, uint32_t lo) {
    return ((double)(hi >> 5) * 67108864.0 + (double)(lo >> 6)) * (1.0 / 9007199254740992.0);
}

// two independent N(0,1) from one philox block (Box-Muller in double precision)
HP_HD void normal_pair(u32x4 r, double& n0, double& n1) {
    double u1 = 1.0 - u53(r.x, r.y);  // (0,1]
    double u2 = u53(r.z, r.w);        // [0,1)
    double rad = sqrt(-2.0 * log(u1));
    double s, c;
#if defined(__CUDA_ARCH__)
    sincospi(2.0 * u2, &s, &c);
#else
    s = sin(6.283185307179586476925286766559 * u2);
    c = cos(6.283185307179586476925286766559 * u2);
#endif
    n0 = rad * c;
    n1 = rad * s;
}

// Two independent N(0,1) with the transcendental part in single precision.  Used for the O(N T)
// fluctuation draws of the GCR step, where the FP64 pipe is the bottleneck (it is shared with DMMA):
// a draw is a random number, its low-order bits carry no information, and a 2^-24 relative
// discretisation of the variate is ~5 orders of magnitude below any Monte-Carlo error.  The radius
// keeps the full 53-bit range of u1 (ln u1 = ln m + e ln 2 with u1 = m 2^e split exactly), so the
// tails reach 8.5 sigma as in the double-precision version.
HP_HD void normal_pair_fast(u32x4 r, double& n0, double& n1) {
    double u1 = 1.0 - u53(r.x, r.y);  // (0,1]
    int e;
    float m = (float)frexp(u1, &e);   // u1 = m 2^e, m in [0.5, 1]
    float u2 = (float)(r.z >> 8) * (1.0f / 16777216.0f) + (float)(r.w >> 8) * (1.0f / 16777216.0f / 16777216.0f);
    float lnu = logf(m) + (float)e * 0.69314718055994530942f;
    float rad = sqrtf(-2.0f * lnu);
    float s, c;
#if defined(__CUDA_ARCH__)
    sincospif(2.0f * u2, &s, &c);
#else
    s = sinf(6.283185307179586f * u2);
    c = cosf(6.283185307179586f * u2);
#endif
    n0 = (double)(rad * c);
    n1 = (double)(rad * s);
}

// Gamma(alpha, 1) for alpha >= 1 (Marsaglia & Tsang 2000).  Consumes philox blocks
// (c0, c1, sub, c3) with sub = 0, 1, 2, ...
HP_HD double gamma_mt(double alpha, uint32_t c0, uint32_t c1, uint32_t c3, uint32_t k0, uint32_t k1) {
    const double d = alpha - 1.0 / 3.0;
    const double c = 1.0 / sqrt(9.0 * d);
    for (uint32_t sub = 0; sub < 4096u; ++sub) {
        u32x4 ctr;
        ctr.x = c0; ctr.y = c1; ctr.z = sub; ctr.w = c3;
        u32x4 r = philox4x32_10(ctr, k0, k1);
        double x, dummy;
        normal_pair(r, x, dummy);
        double v = 1.0 + c * x;
        if (v <= 0.0) continue;
        v = v * v * v;
        ctr.z = sub | 0x80000000u;
        u32x4 r2 = philox4x32_10(ctr, k0, k1);
        double u = 1.0 - u53(r2.x, r2.y);
        if (log(u) < 0.5 * x * x + d - d * v + d * log(v)) return d * v;
    }
    return d;  // unreachable in practice (acceptance > 95 %)
}

// ------------------------------------------------------------------------------------------
// Regularised incomplete gamma functions.
//
// prefactor  x^a e^-x / Gamma(a+1) = exp(a (log1p(mu) - mu)) / (sqrt(2 pi a) Gamma*(a)),
// mu = (x-a)/a, with Gamma*(a) the Stirling correction.  This keeps full relative accuracy
// near the transition x ~ a for the large shape parameters used here (a = Ntimes).
HP_HD double gamma_star(double a) {
    // Gamma(a) = sqrt(2 pi / a) (a/e)^a Gamma*(a);  asymptotic series, good to 1e-16 for a >= 12
    double r = 1.0 / a, r2 = r * r;
    return 1.0 + r * (1.0 / 12.0) + r2 * (1.0 / 288.0) - r2 * r * (139.0 / 51840.0)
         - r2 * r2 * (571.0 / 2488320.0) + r2 * r2 * r * (163879.0 / 209018880.0)
         + r2 * r2 * r2 * (5246819.0 / 75246796800.0) - r2 * r2 * r2 * r * (534703531.0 / 902961561600.0);
}

HP_HD double igam_prefactor(double a, double x) {
    // x^a e^-x / Gamma(a+1)
    if (a >= 12.0) {
        double mu = (x - a) / a;
        double e = a * (log1p(mu) - mu);
        if (e < -745.0) return 0.0;
        return exp(e) / (sqrt(6.283185307179586476925286766559 * a) * gamma_star(a));
    }
    double e = a * log(x) - x - lgamma(a + 1.0);
    if (e < -745.0) return 0.0;
    return exp(e);
}

// Q(a, x) = Gamma(a, x) / Gamma(a), a > 0, x >= 0
HP_HD double igamc(double a, double x) {
    if (!(x > 0.0)) return 1.0;
    if (isinf(x)) return 0.0;
    const double eps = 1.1102230246251565e-16;
    double pref = igam_prefactor(a, x);
    if (x < a + 1.0) {
        // P by power series, Q = 1 - P
        double term = 1.0, sum = 1.0, ap = a;
        for (int n = 0; n < 20000; ++n) {
            ap += 1.0;
            term *= x / ap;
            sum += term;
            if (term < sum * eps * 0.25) break;
        }
        double p = pref * sum;
        return p < 1.0 ? 1.0 - p : 0.0;
    }
    // Q by continued fraction (modified Lentz), Q = pref * a * CF ... written with
    // x^a e^-x / Gamma(a) = pref * a
    const double tiny = 1e-300;
    double b = x + 1.0 - a;
    double c = 1.0 / tiny;
    double d = 1.0 / b;
    double h = d;
    for (int i = 1; i < 20000; ++i) {
        double an = -(double)i * ((double)i - a);
        b += 2.0;
        d = an * d + b;
        if (fabs(d) < tiny) d = tiny;
        c = b + an / c;
        if (fabs(c) < tiny) c = tiny;
        d = 1.0 / d;
        double del = d * c;
        h *= del;
        if (fabs(del - 1.0) < eps) break;
    }
    return pref * a * h;
}

// ------------------------------------------------------------------------------------------
// Grid of the reference's inversion sampler (pspec.py:50): np.logspace(log10(lo), log10(hi), n)
HP_HD double invsamp_grid_x(double log10_lo, double log10_hi, int j, int ngrid) {
    // numpy linspace: start + j*step, last point pinned to stop
    double step = (log10_hi - log10_lo) / (double)(ngrid - 1);
    double y = (j == ngrid - 1) ? log10_hi : log10_lo + (double)j * step;
    return pow(10.0, y);
}

// Serial reference implementation of the whole sampler (the device kernel evaluates the CDF
// grid in parallel and then calls invsamp_from_cdf with the same arrays).
//   cdf[] is overwritten:  normalised, then compacted to its strictly increasing subsequence.
HP_HD double invsamp_from_cdf(double* cdf, double* xg, int ngrid, double u) {
    // cdf -= min ; cdf /= max  (pspec.py:52-53)
    double mn = cdf[0], mx;
    for (int j = 1; j < ngrid; ++j) mn = cdf[j] < mn ? cdf[j] : mn;
    for (int j = 0; j < ngrid; ++j) cdf[j] -= mn;
    mx = cdf[0];
    for (int j = 1; j < ngrid; ++j) mx = cdf[j] > mx ? cdf[j] : mx;
    for (int j = 0; j < ngrid; ++j) cdf[j] /= mx;
    // np.unique(cdf, return_index=True) for a non-decreasing sequence: keep first occurrences
    // (Q(a, beta/x) is non-decreasing in x; a sample that violates monotonicity by an ulp is
    //  merged into its predecessor, which moves the interpolant by less than that ulp).
    int m = 0;
    for (int j = 0; j < ngrid; ++j) {
        if (m == 0 || cdf[j] > cdf[m - 1]) { cdf[m] = cdf[j]; xg[m] = xg[j]; ++m; }
    }
    if (m < 2) return NAN;  // reference: interp1d raises / returns nan on a degenerate grid
    // interp1d(kind='linear'): u in [cdf[0], cdf[m-1]] = [0, 1]
    int lo = 0, hi = m - 1;
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (cdf[mid] <= u) lo = mid; else hi = mid;
    }
    double slope = (xg[hi] - xg[lo]) / (cdf[hi] - cdf[lo]);
    return slope * (u - cdf[lo]) + xg[lo];
}

// ------------------------------------------------------------------------------------------
// Scalar model of scipy.sparse.linalg.cg(A, b, rtol=1e-8, atol=1e-6, M=pinv(A)) as the
// reference calls it (pspec.py:228).  Because M A = I, every iterate is a multiple of
// x* = A^-1 b:  x = xi x*, r = rho b, p = pi x*.  Returns xi (complex) given c = b^H x* and
// ||b||.  See oracle/hydra_oracle.py:cg_theta for the derivation.
struct cplx { double re, im; };
HP_HD cplx cmul(cplx a, cplx b) { cplx r; r.re = a.re * b.re - a.im * b.im; r.im = a.re * b.im + a.im * b.re; return r; }
HP_HD cplx cdiv(cplx a, cplx b) {
    // Smith's algorithm (what C / numpy use for complex128 division)
    cplx r;
    if (fabs(b.re) >= fabs(b.im)) {
        double t = b.im / b.re, den = b.re + b.im * t;
        r.re = (a.re + a.im * t) / den;
        r.im = (a.im - a.re * t) / den;
    } else {
        double t = b.re / b.im, den = b.re * t + b.im;
        r.re = (a.re * t + a.im) / den;
        r.im = (a.im * t - a.re) / den;
    }
    return r;
}

HP_HD cplx cg_theta(cplx c, double bnorm, double rtol, double atol, int maxiter) {
    cplx xi; xi.re = 0.0; xi.im = 0.0;
    if (bnorm == 0.0) return xi;
    double tol = atol > rtol * bnorm ? atol : rtol * bnorm;
    cplx rho; rho.re = 1.0; rho.im = 0.0;
    cplx pi; pi.re = 0.0; pi.im = 0.0;
    cplx rho_prev; rho_prev.re = 1.0; rho_prev.im = 0.0;
    cplx cc; cc.re = c.re; cc.im = -c.im;
    for (int it = 0; it < maxiter; ++it) {
        double arho = hypot(rho.re, rho.im);
        if (arho * bnorm < tol) break;
        double ar2 = arho * arho;
        cplx rho_cur; rho_cur.re = ar2 * c.re; rho_cur.im = ar2 * c.im;
        if (it == 0) {
            pi = rho;
        } else {
            cplx beta = cdiv(rho_cur, rho_prev);
            cplx bp = cmul(beta, pi);
            pi.re = rho.re + bp.re;
            pi.im = rho.im + bp.im;
        }
        double ap = hypot(pi.re, pi.im);
        double ap2 = ap * ap;
        cplx den; den.re = ap2 * cc.re; den.im = ap2 * cc.im;
        cplx alpha = cdiv(rho_cur, den);
        cplx step = cmul(alpha, pi);
        xi.re += step.re; xi.im += step.im;
        rho.re -= step.re; rho.im -= step.im;
        rho_prev = rho_cur;
    }
    return xi;
}

}  // namespace hp
