// hp_diag.cuh -- 32x32 complex diagonal block: Cholesky factor and its triangular inverse by ONE warp.
//
// The block lives in shared memory (planar re / im, row stride kLdBlk).  The first version used all
// 512 threads with two __syncthreads per column (128 block-wide barriers per diagonal block, ~50 us:
// more than the DMMA time of the whole block column).  Here a single warp runs a left-looking
// factorisation (lane = row; row c is a broadcast read, row `lane` a conflict-free strided one) and
// then the inverse (lane = column of V; every lane only re-reads its own column), with __syncwarp as
// the only synchronisation: ~1000 shared-memory loads + ~1000 complex FMAs per lane, ~10 us.
#pragma once
#include "hp_kernels.cuh"

namespace hp {

// In place: lower triangle of (Ar, Ai) := L (upper part zeroed), (Vr, Vi) := L^-1 (upper part zeroed).
// Returns true if a pivot was not positive.  Call with all 32 lanes of one warp.
__device__ __forceinline__ bool diag_chol_inverse_warp(double* Ar, double* Ai, double* Vr, double* Vi) {
    const int lane = threadIdx.x & 31;
    bool bad = false;
    double* rowr = Ar + lane * kLdBlk;
    double* rowi = Ai + lane * kLdBlk;
    for (int c = 0; c < 32; ++c) {
        const double* cr = Ar + c * kLdBlk;
        const double* ci = Ai + c * kLdBlk;
        double s0r = 0.0, s0i = 0.0, s1r = 0.0, s1i = 0.0;
        int p = 0;
        for (; p + 1 < c; p += 2) {
            const double lr0 = rowr[p], li0 = rowi[p], xr0 = cr[p], xi0 = ci[p];
            const double lr1 = rowr[p + 1], li1 = rowi[p + 1], xr1 = cr[p + 1], xi1 = ci[p + 1];
            s0r += lr0 * xr0 + li0 * xi0; s0i += li0 * xr0 - lr0 * xi0;
            s1r += lr1 * xr1 + li1 * xi1; s1i += li1 * xr1 - lr1 * xi1;
        }
        if (p < c) {
            const double lr0 = rowr[p], li0 = rowi[p], xr0 = cr[p], xi0 = ci[p];
            s0r += lr0 * xr0 + li0 * xi0; s0i += li0 * xr0 - lr0 * xi0;
        }
        const double xr = rowr[c] - (s0r + s1r), xi = rowi[c] - (s0i + s1i);
        const double piv = __shfl_sync(0xffffffffu, xr, c);
        if (!(piv > 0.0)) bad = true;
        const double d = sqrt(piv), inv = 1.0 / d;
        __syncwarp();  // every lane has read column c of its row and the old row c
        if (lane == c) { rowr[c] = d; rowi[c] = 0.0; }
        else if (lane > c) { rowr[c] = xr * inv; rowi[c] = xi * inv; }
        else { rowr[c] = 0.0; rowi[c] = 0.0; }
        __syncwarp();
    }
    // V = L^-1, lane = column:  V[r][lane] = (delta - sum_{p<r} L[r][p] V[p][lane]) / L[r][r]
    for (int r = 0; r < 32; ++r) {
        const double* lr_ = Ar + r * kLdBlk;
        const double* li_ = Ai + r * kLdBlk;
        double s0r = 0.0, s0i = 0.0, s1r = 0.0, s1i = 0.0;
        int p = 0;
        for (; p + 1 < r; p += 2) {
            const double lr0 = lr_[p], li0 = li_[p], vr0 = Vr[p * kLdBlk + lane], vi0 = Vi[p * kLdBlk + lane];
            const double lr1 = lr_[p + 1], li1 = li_[p + 1], vr1 = Vr[(p + 1) * kLdBlk + lane], vi1 = Vi[(p + 1) * kLdBlk + lane];
            s0r += lr0 * vr0 - li0 * vi0; s0i += lr0 * vi0 + li0 * vr0;
            s1r += lr1 * vr1 - li1 * vi1; s1i += lr1 * vi1 + li1 * vr1;
        }
        if (p < r) {
            const double lr0 = lr_[p], li0 = li_[p], vr0 = Vr[p * kLdBlk + lane], vi0 = Vi[p * kLdBlk + lane];
            s0r += lr0 * vr0 - li0 * vi0; s0i += lr0 * vi0 + li0 * vr0;
        }
        const double dinv = 1.0 / lr_[r];
        double vr = 0.0, vi = 0.0;
        if (lane < r) { vr = -(s0r + s1r) * dinv; vi = -(s0i + s1i) * dinv; }
        else if (lane == r) { vr = dinv; }
        Vr[r * kLdBlk + lane] = vr;
        Vi[r * kLdBlk + lane] = vi;
    }
    __syncwarp();
    return bad;
}

}  // namespace hp
