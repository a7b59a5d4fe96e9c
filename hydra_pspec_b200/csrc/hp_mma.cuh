// hp_mma.cuh -- complex128 block products on the FP64 tensor pipe (DMMA.8x8x4, sm_100a).
//
// B200's FP64 tensor instruction is mma.sync.m8n8k4.f64 (larger PTX shapes lower to sequences
// of DMMA.8x8x4 -- see profiles/microbench/fp64_peak.cu: 37.1 TFLOP/s measured vs 33.5 for
// DFMA).  A complex product is four real ones on planar (re / im) operands held in shared
// memory:
//      Cr += Ar.Br + (-sa sb Ai).Bi        Ci += (sb Ar).Bi + (sa Ai).Br
// with sa, sb = -1 when the operand is conjugated.
//
// Fragment layout of m8n8k4 (PTX ISA, "Matrix fragments for mma.m8n8k4 with .f64"):
//      g = lane >> 2, q = lane & 3
//      A (8x4, row):  a  = A[g][q]
//      B (4x8, col):  b  = B[q][g]
//      C (8x8):       c0 = C[g][2q], c1 = C[g][2q+1]
//
// Shared-memory operand tiles are addressed through two patterns, both conflict-free for a
// leading dimension ld == 4 (mod 16) doubles:
//      'N':  elem(i, k) at base[i * ld + k]      (k contiguous)
//      'T':  elem(i, k) at base[k * ld + i]      (stored transposed)
#pragma once
#include <cuda_runtime.h>

namespace hp {

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// Accumulate a (8*WI) x (8*WJ) complex warp tile:  C += op(A) (8WI x K) . op(B) (K x 8WJ)
//   A tile origin = element (row 0, k 0) of this warp's rows;  B tile origin = (k 0, col 0).
//   AT / BT : operand stored transposed in shared memory ('T' pattern above)
//   AC / BC : operand conjugated
//   K multiple of 4.
template <int WI, int WJ, bool AT, bool AC, bool BT, bool BC>
__device__ __forceinline__ void warp_zgemm(double (&cr)[WI][WJ][2], double (&ci)[WI][WJ][2],
                                           const double* __restrict__ Ar, const double* __restrict__ Ai, int lda,
                                           const double* __restrict__ Br, const double* __restrict__ Bi, int ldb,
                                           int K) {
    const int lane = threadIdx.x & 31;
    const int g = lane >> 2, q = lane & 3;
    const double sa = AC ? -1.0 : 1.0, sb = BC ? -1.0 : 1.0;
#pragma unroll 2
    for (int kk = 0; kk < K; kk += 4) {
        double ar[WI], ai_sa[WI], ai_neg[WI], ar_sb[WI];
        double br[WJ], bi[WJ];
#pragma unroll
        for (int i = 0; i < WI; ++i) {
            const int off = AT ? (kk + q) * lda + (8 * i + g) : (8 * i + g) * lda + (kk + q);
            double xr = Ar[off], xi = Ai[off];
            ar[i] = xr;
            ar_sb[i] = sb * xr;
            ai_sa[i] = sa * xi;
            ai_neg[i] = -sa * sb * xi;
        }
#pragma unroll
        for (int j = 0; j < WJ; ++j) {
            const int off = BT ? (8 * j + g) * ldb + (kk + q) : (kk + q) * ldb + (8 * j + g);
            br[j] = Br[off];
            bi[j] = Bi[off];
        }
#pragma unroll
        for (int i = 0; i < WI; ++i)
#pragma unroll
            for (int j = 0; j < WJ; ++j) {
                dmma884(cr[i][j][0], cr[i][j][1], ar[i], br[j]);
                dmma884(ci[i][j][0], ci[i][j][1], ar_sb[i], bi[j]);
                dmma884(cr[i][j][0], cr[i][j][1], ai_neg[i], bi[j]);
                dmma884(ci[i][j][0], ci[i][j][1], ai_sa[i], br[j]);
            }
    }
}

// Same product with three real DMMAs per complex MAC instead of four (the "3M" scheme, as in k_solve):
//      P1 += Ar.Br,   P2 += Ai.Bi,   P3 += (Ar + sa Ai).(Br + sb Bi)
//      C  += (P1 - sa sb P2) + i (P3 - P1 - sa sb P2)
// The operand sums cost two DADDs per fragment; the FP64 tensor pipe sees 25 % fewer instructions.  Normwise the
// rounding error is that of the ordinary product.  The caller keeps P[3] across its K loop and calls
// warp_zgemm3m_finish once.
template <int WI, int WJ, bool AT, bool AC, bool BT, bool BC>
__device__ __forceinline__ void warp_zgemm3m(double (&P)[3][WI][WJ][2], const double* __restrict__ Ar,
                                             const double* __restrict__ Ai, int lda, const double* __restrict__ Br,
                                             const double* __restrict__ Bi, int ldb, int K) {
    const int lane = threadIdx.x & 31;
    const int g = lane >> 2, q = lane & 3;
#pragma unroll 2
    for (int kk = 0; kk < K; kk += 4) {
        double ar[WI], ai[WI], as[WI];
        double br[WJ], bi[WJ], bs[WJ];
#pragma unroll
        for (int i = 0; i < WI; ++i) {
            const int off = AT ? (kk + q) * lda + (8 * i + g) : (8 * i + g) * lda + (kk + q);
            ar[i] = Ar[off]; ai[i] = Ai[off];
            as[i] = AC ? ar[i] - ai[i] : ar[i] + ai[i];
        }
#pragma unroll
        for (int j = 0; j < WJ; ++j) {
            const int off = BT ? (8 * j + g) * ldb + (kk + q) : (kk + q) * ldb + (8 * j + g);
            br[j] = Br[off]; bi[j] = Bi[off];
            bs[j] = BC ? br[j] - bi[j] : br[j] + bi[j];
        }
#pragma unroll
        for (int i = 0; i < WI; ++i)
#pragma unroll
            for (int j = 0; j < WJ; ++j) {
                dmma884(P[0][i][j][0], P[0][i][j][1], ar[i], br[j]);
                dmma884(P[1][i][j][0], P[1][i][j][1], ai[i], bi[j]);
                dmma884(P[2][i][j][0], P[2][i][j][1], as[i], bs[j]);
            }
    }
}
template <int WI, int WJ>
__device__ __forceinline__ void warp_zero3m(double (&P)[3][WI][WJ][2]) {
#pragma unroll
    for (int p = 0; p < 3; ++p)
#pragma unroll
        for (int i = 0; i < WI; ++i)
#pragma unroll
            for (int j = 0; j < WJ; ++j) P[p][i][j][0] = P[p][i][j][1] = 0.0;
}
// cr + i ci = the accumulated product;  AC / BC as in the accumulation calls
template <int WI, int WJ, bool AC, bool BC>
__device__ __forceinline__ void warp_zgemm3m_finish(const double (&P)[3][WI][WJ][2], double (&cr)[WI][WJ][2],
                                                    double (&ci)[WI][WJ][2]) {
    constexpr double s = (AC != BC) ? -1.0 : 1.0;  // sa * sb
#pragma unroll
    for (int i = 0; i < WI; ++i)
#pragma unroll
        for (int j = 0; j < WJ; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                cr[i][j][e] = P[0][i][j][e] - s * P[1][i][j][e];
                ci[i][j][e] = P[2][i][j][e] - P[0][i][j][e] - s * P[1][i][j][e];
            }
}

template <int WI, int WJ>
__device__ __forceinline__ void warp_zero(double (&cr)[WI][WJ][2], double (&ci)[WI][WJ][2]) {
#pragma unroll
    for (int i = 0; i < WI; ++i)
#pragma unroll
        for (int j = 0; j < WJ; ++j) {
            cr[i][j][0] = cr[i][j][1] = 0.0;
            ci[i][j][0] = ci[i][j][1] = 0.0;
        }
}

}  // namespace hp
