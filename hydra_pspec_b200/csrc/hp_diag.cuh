// hp_diag.cuh -- 32x32 complex diagonal block: Cholesky factor and its triangular inverse by TWO warps.
//
// The block lives in shared memory (planar re / im, row stride kLdBlk).  The first version used all
// 512 threads with two __syncthreads per column (128 block-wide barriers per diagonal block).  Here
// warp 0 runs a left-looking factorisation (lane = row; row c is a broadcast read, row `lane` a
// conflict-free strided one) and warp 1 follows one column behind with the inverse (lane = column of
// V; a lane only re-reads its own column, and row r of L is final once warp 0 has finished column
// r).  The two warps meet at a 64-thread named barrier once per column; nothing else in the CTA
// synchronises.  Loads are 16-byte (two columns per step) with two independent accumulator pairs.
#pragma once
#include "hp_kernels.cuh"

namespace hp {

constexpr int kDiagBarrier = 8;  // named barrier id shared by the two warps (ids 1..3 are k_solve's)

__device__ __forceinline__ void diag_pair_sync() { asm volatile("bar.sync %0, 64;" ::"n"(kDiagBarrier) : "memory"); }

// In place: lower triangle of (Ar, Ai) := L (upper part zeroed), (Vr, Vi) := L^-1 (upper part zeroed).
// Call with warps 0 and 1 of the CTA (all 64 threads).  Returns (on warp 0) true if a pivot was not positive.
__device__ __forceinline__ bool diag_chol_inverse_2warps(double* Ar, double* Ai, double* Vr, double* Vi) {
    const int lane = threadIdx.x & 31;
    const int which = (threadIdx.x >> 5) & 1;
    bool bad = false;
    if (which == 0) {
        double* rowr = Ar + lane * kLdBlk;
        double* rowi = Ai + lane * kLdBlk;
        for (int c = 0; c < 32; ++c) {
            const double* cr = Ar + c * kLdBlk;
            const double* ci = Ai + c * kLdBlk;
            double s0r = 0.0, s0i = 0.0, s1r = 0.0, s1i = 0.0;
            int p = 0;
            for (; p + 1 < c; p += 2) {
                const double2 lr = *reinterpret_cast<const double2*>(rowr + p), li = *reinterpret_cast<const double2*>(rowi + p);
                const double2 xr = *reinterpret_cast<const double2*>(cr + p), xi = *reinterpret_cast<const double2*>(ci + p);
                s0r += lr.x * xr.x + li.x * xi.x; s0i += li.x * xr.x - lr.x * xi.x;
                s1r += lr.y * xr.y + li.y * xi.y; s1i += li.y * xr.y - lr.y * xi.y;
            }
            if (p < c) {
                const double lr0 = rowr[p], li0 = rowi[p], xr0 = cr[p], xi0 = ci[p];
                s0r += lr0 * xr0 + li0 * xi0; s0i += li0 * xr0 - lr0 * xi0;
            }
            const double xr = rowr[c] - (s0r + s1r), xi = rowi[c] - (s0i + s1i);
            const double piv = __shfl_sync(0xffffffffu, xr, c);
            if (!(piv > 0.0)) bad = true;
            const double inv = rsqrt(piv), d = piv * inv;
            if (lane == c) { rowr[c] = d; rowi[c] = 0.0; }
            else if (lane > c) { rowr[c] = xr * inv; rowi[c] = xi * inv; }
            else { rowr[c] = 0.0; rowi[c] = 0.0; }
            __syncwarp();
            diag_pair_sync();  // column c of L (hence row c up to its diagonal) is final: warp 1 may take row c
        }
    } else {
        // V = L^-1, lane = column:  V[r][lane] = (delta - sum_{p<r} L[r][p] V[p][lane]) / L[r][r]
        for (int r = 0; r < 32; ++r) {
            diag_pair_sync();
            const double* lr_ = Ar + r * kLdBlk;
            const double* li_ = Ai + r * kLdBlk;
            double s0r = 0.0, s0i = 0.0, s1r = 0.0, s1i = 0.0;
            int p = 0;
            for (; p + 1 < r; p += 2) {
                const double2 lr = *reinterpret_cast<const double2*>(lr_ + p), li = *reinterpret_cast<const double2*>(li_ + p);
                const double vr0 = Vr[p * kLdBlk + lane], vi0 = Vi[p * kLdBlk + lane];
                const double vr1 = Vr[(p + 1) * kLdBlk + lane], vi1 = Vi[(p + 1) * kLdBlk + lane];
                s0r += lr.x * vr0 - li.x * vi0; s0i += lr.x * vi0 + li.x * vr0;
                s1r += lr.y * vr1 - li.y * vi1; s1i += lr.y * vi1 + li.y * vr1;
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
    }
    return bad;
}

}  // namespace hp
