// hp_solve.cu -- k_solve: two-sided triangular solve of the factored GCR system for all times.
//
// Replaces the per-time preconditioned CG of the reference (gcr_fgmodes_1d, pspec.py:151-235):
// with M = L L^H from k_chol, every time sample is a right-hand side of the same system.
//
// CTA = (tile of 16 times, baseline).  The 16-column solution tile stays in shared memory for the
// whole forward (L) and backward (L^H) block substitution; the 32x32 blocks of L stream from L2
// through a 3-stage ring filled by TMA bulk copies (cp.async.bulk + mbarrier, one producer warp),
// and eight consumer warps run the block products on the FP64 tensor pipe (DMMA.8x8x4), each on an
// 8x8 complex tile with four independent accumulators.
//
// Right-hand sides are built on the fly (never stored):
//     injected draws :  r = lam * Rfix + wa                      (Rfix carries B^H N^-1/2 omega_b)
//     Philox         :  r = lam * Rfix, and xi ~ CN(0, I) is added to y = L^-1 r before the
//                       backward pass.  Since cov(lam B^H N^-1/2 omega_b + omega_a) = M = L L^H, this
//                       is the same distribution as drawing omega_a, omega_b (pspec.py:215-222) and
//                       needs no transform of the noise realisation.
#include "hp_kernels.cuh"
#include "hp_math.h"
#include "hp_mma.cuh"

namespace hp {

namespace {

constexpr int kLdX = 20;       // 16 + 4 doubles, == 4 mod 16
constexpr int kStages = 3;
constexpr int kConsumers = 256;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "LAB_WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra LAB_DONE_%=;\n\t"
        "bra LAB_WAIT_%=;\n\t"
        "LAB_DONE_%=:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// acc[0] += Ar.Br, acc[1] += Ai.Bi, acc[2] += Ar.Bi, acc[3] += Ai.Br  over k in [k0, k1) (multiples of 4).
//   A element (row g, k): AT ? A[k * lda + g] : A[g * lda + k]   (pointers already offset to the warp's rows)
//   B element (k, col g): B[k * ldb + g]
template <bool AT>
__device__ __forceinline__ void tile_mma(double (&acc)[4][2], const double* __restrict__ Ar, const double* __restrict__ Ai,
                                         int lda, const double* __restrict__ Br, const double* __restrict__ Bi, int ldb,
                                         int k0, int k1) {
    const int lane = threadIdx.x & 31, g = lane >> 2, q = lane & 3;
    const int aoff = AT ? q * lda + g : g * lda + q;
    const int astep = AT ? 4 * lda : 4;
    const double* pa_r = Ar + aoff + (AT ? k0 * lda : k0);
    const double* pa_i = Ai + aoff + (AT ? k0 * lda : k0);
    const double* pb_r = Br + (k0 + q) * ldb + g;
    const double* pb_i = Bi + (k0 + q) * ldb + g;
    double ar = *pa_r, ai = *pa_i, br = *pb_r, bi = *pb_i;
#pragma unroll 4
    for (int kk = k0; kk < k1; kk += 4) {
        double nar = 0.0, nai = 0.0, nbr = 0.0, nbi = 0.0;
        if (kk + 4 < k1) {
            pa_r += astep; pa_i += astep; pb_r += 4 * ldb; pb_i += 4 * ldb;
            nar = *pa_r; nai = *pa_i; nbr = *pb_r; nbi = *pb_i;
        }
        dmma884(acc[0][0], acc[0][1], ar, br);
        dmma884(acc[1][0], acc[1][1], ai, bi);
        dmma884(acc[2][0], acc[2][1], ar, bi);
        dmma884(acc[3][0], acc[3][1], ai, br);
        ar = nar; ai = nai; br = nbr; bi = nbi;
    }
}

struct RhsCtx {
    const double* Rfix; const double* eta; const double* wa; const double* lam;
    int Np, n, N, T, t0;
};

// right-hand side element (system row `row`, tile column `col`)
__device__ __forceinline__ void rhs_elem(const RhsCtx& c, int row, int col, double& vr, double& vi) {
    int t = c.t0 + col;
    vr = 0.0; vi = 0.0;
    if (t >= c.T || row >= c.N) return;
    size_t off = 2 * ((size_t)t * c.Np + row);
    double2 x = *reinterpret_cast<const double2*>(c.Rfix + off);
    if (c.eta) { double2 e = *reinterpret_cast<const double2*>(c.eta + off); x.x += e.x; x.y += e.y; }
    double l = c.lam[row];
    vr = l * x.x; vi = l * x.y;
    if (c.wa && row < c.n) { double2 wv = *reinterpret_cast<const double2*>(c.wa + off); vr += wv.x; vi += wv.y; }
}

}  // namespace

size_t solve_smem_bytes(int nblk) {
    size_t Np = (size_t)nblk * 32;
    size_t d = 2 * Np * kLdX               // X tile planes
             + (size_t)kStages * kLBlkDoubles  // ring of L blocks
             + 2 * 32 * kLdX               // Z
             + 16 * 16 * 3 + 32            // reductions + theta
             + 2 * kStages + 2;            // mbarriers
    return d * sizeof(double);
}

__global__ void __launch_bounds__(288) k_solve(SolveArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int nblk = a.nblk, Np = nblk * 32;
    double* ring = reinterpret_cast<double*>(smem_raw);          // [kStages][2304], 16-byte aligned for TMA
    double* Xr = ring + (size_t)kStages * kLBlkDoubles;
    double* Xi = Xr + (size_t)Np * kLdX;
    double* Zr = Xi + (size_t)Np * kLdX;
    double* Zi = Zr + 32 * kLdX;
    double* red = Zi + 32 * kLdX;                                 // 16*16*3
    double* theta = red + 16 * 16 * 3;                            // 32
    uint64_t* full = reinterpret_cast<uint64_t*>(theta + 32);     // [kStages]
    uint64_t* empty = full + kStages;                             // [kStages]

    const int sys = blockIdx.y, tile = blockIdx.x;
    const double* Lp = a.Lp + (size_t)sys * tri_blocks(nblk) * kLBlkDoubles;
    const double* Linvp = a.Linvp + (size_t)sys * nblk * kLBlkDoubles;
    const double* lam = a.lam + (size_t)sys * Np;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == 8) {
        // ------------------------------------------------------------------ producer warp
        if (lane == 0) {
            // pull this tile's right-hand-side rows into L2 ahead of the consumers' row-by-row loads
            for (int t = 0; t < kTT; ++t) {
                if (tile * kTT + t >= a.T) break;
                const double* rrow = a.Rfix + 2 * (((size_t)sys * a.Tp + (size_t)tile * kTT + t) * Np);
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(rrow), "r"((uint32_t)(Np * 16)) : "memory");
                if (a.wa) {
                    const double* wrow = a.wa + 2 * (((size_t)sys * a.Tp + (size_t)tile * kTT + t) * Np);
                    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(wrow), "r"((uint32_t)(Np * 16)) : "memory");
                }
            }
            uint32_t item = 0;
            auto push = [&](const double* src) {
                const uint32_t s = item % kStages, k = item / kStages;
                if (k > 0) mbar_wait(&empty[s], (k - 1) & 1);
                mbar_arrive_expect_tx(&full[s], kLBlkDoubles * 8);
                bulk_g2s(ring + (size_t)s * kLBlkDoubles, src, kLBlkDoubles * 8, &full[s]);
                ++item;
            };
            for (int i = 0; i < nblk; ++i) {
                for (int j = 0; j < i; ++j) push(Lp + blk_index(i, j) * kLBlkDoubles);
                push(Linvp + (size_t)i * kLBlkDoubles);
            }
            for (int i = nblk - 1; i >= 0; --i) {
                for (int j = i + 1; j < nblk; ++j) push(Lp + blk_index(j, i) * kLBlkDoubles);
                push(Linvp + (size_t)i * kLBlkDoubles);
            }
        }
        return;
    }

    // ---------------------------------------------------------------------- consumer warps
    RhsCtx rc;
    rc.Rfix = a.Rfix + 2 * (size_t)sys * a.Tp * Np;
    rc.eta = a.eta ? a.eta + 2 * (size_t)sys * a.Tp * Np : nullptr;
    rc.wa = a.wa ? a.wa + 2 * (size_t)sys * a.Tp * Np : nullptr;
    rc.lam = lam; rc.Np = Np; rc.n = a.n; rc.N = a.N; rc.T = a.T; rc.t0 = tile * kTT;
    const uint32_t chain = a.chain_ids ? (uint32_t)a.chain_ids[sys] : (uint32_t)sys;
    const int g = lane >> 2, q = lane & 3;
    const int ti = warp >> 1, tj = warp & 1;  // warp tile: rows 8 ti.., cols 8 tj..
    uint32_t item = 0;
    auto stage_ptr = [&](uint32_t s) { return ring + (size_t)s * kLBlkDoubles; };
    auto acquire = [&]() -> const double* {
        const uint32_t s = item % kStages, k = item / kStages;
        mbar_wait(&full[s], k & 1);
        return stage_ptr(s);
    };
    auto release = [&]() {
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[item % kStages]);
        ++item;
    };

    // ------------------------------------------------------------------ forward:  L Y = R
    double rn[2], in_[2];  // right-hand side of the next block row (prefetched one row ahead)
#pragma unroll
    for (int e = 0; e < 2; ++e) rhs_elem(rc, 8 * ti + g, 8 * tj + 2 * q + e, rn[e], in_[e]);
    for (int i = 0; i < nblk; ++i) {
        double rr[2] = {rn[0], rn[1]}, ri[2] = {in_[0], in_[1]};
        if (i + 1 < nblk) {
#pragma unroll
            for (int e = 0; e < 2; ++e) rhs_elem(rc, 32 * (i + 1) + 8 * ti + g, 8 * tj + 2 * q + e, rn[e], in_[e]);
        }
        double acc[4][2] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};
        for (int j = 0; j < i; ++j) {
            const double* blk = acquire();
            tile_mma<false>(acc, blk + 8 * ti * kLdBlk, blk + kLPlane + 8 * ti * kLdBlk, kLdBlk,
                            Xr + (size_t)32 * j * kLdX + 8 * tj, Xi + (size_t)32 * j * kLdX + 8 * tj, kLdX, 0, 32);
            release();
        }
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            int r = 8 * ti + g, c = 8 * tj + 2 * q + e;
            Zr[r * kLdX + c] = rr[e] - (acc[0][e] - acc[1][e]);
            Zi[r * kLdX + c] = ri[e] - (acc[2][e] + acc[3][e]);
        }
        consumer_sync();
        {
            const double* blk = acquire();
            double y[4][2] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};
            // Y_i = Linv_ii . Z ; Linv lower triangular: k < 8 (ti + 1)
            tile_mma<false>(y, blk + 8 * ti * kLdBlk, blk + kLPlane + 8 * ti * kLdBlk, kLdBlk, Zr + 8 * tj, Zi + 8 * tj, kLdX,
                            0, 8 * (ti + 1));
            release();
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                int row = 32 * i + 8 * ti + g, c = 8 * tj + 2 * q + e;
                double vr = y[0][e] - y[1][e], vi = y[2][e] + y[3][e];
                if (a.philox_wa && row < a.N && rc.t0 + c < a.T) {
                    // y += xi, xi ~ CN(0, 1): fluctuation term of the constrained realisation
                    u32x4 ctr; ctr.x = (uint32_t)row; ctr.y = (uint32_t)(rc.t0 + c); ctr.z = a.iter; ctr.w = chain;
                    double n0, n1;
                    normal_pair(philox4x32_10(ctr, a.key0, a.key1 ^ 0xA5A5A5A5u), n0, n1);
                    vr += n0 * 0.70710678118654752440; vi += n1 * 0.70710678118654752440;
                }
                Xr[(size_t)row * kLdX + c] = vr;
                Xi[(size_t)row * kLdX + c] = vi;
            }
        }
        consumer_sync();
    }

    // ------------------------------------------------------------------ backward:  L^H X = Y
    for (int i = nblk - 1; i >= 0; --i) {
        double acc[4][2] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};
        for (int j = i + 1; j < nblk; ++j) {
            const double* blk = acquire();
            // acc += L_ji^H . X_j :  A element (r, k) = conj(L_ji[k][r])
            tile_mma<true>(acc, blk + 8 * ti, blk + kLPlane + 8 * ti, kLdBlk, Xr + (size_t)32 * j * kLdX + 8 * tj,
                           Xi + (size_t)32 * j * kLdX + 8 * tj, kLdX, 0, 32);
            release();
        }
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            int r = 8 * ti + g, c = 8 * tj + 2 * q + e;
            // conj(A) B:  re = rr + ii, im = ri - ir
            Zr[r * kLdX + c] = Xr[(size_t)(32 * i + r) * kLdX + c] - (acc[0][e] + acc[1][e]);
            Zi[r * kLdX + c] = Xi[(size_t)(32 * i + r) * kLdX + c] - (acc[2][e] - acc[3][e]);
        }
        consumer_sync();
        {
            const double* blk = acquire();
            double y[4][2] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};
            // X_i = Linv_ii^H . Z ; Linv^H upper triangular: k >= 8 ti
            tile_mma<true>(y, blk + 8 * ti, blk + kLPlane + 8 * ti, kLdBlk, Zr + 8 * tj, Zi + 8 * tj, kLdX, 8 * ti, 32);
            release();
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                int row = 32 * i + 8 * ti + g, c = 8 * tj + 2 * q + e;
                Xr[(size_t)row * kLdX + c] = y[0][e] + y[1][e];
                Xi[(size_t)row * kLdX + c] = y[2][e] - y[3][e];
            }
        }
        consumer_sync();
    }

    // ------------------------------------------------------------------ epilogue
    if (a.cg_compat) {
        // c = b^H x*, ||b||^2 in the reference's (unwhitened) variables:
        //   weights lam^2 on the signal rows, 1 on the foreground rows ("CG model").
        int col = tid & 15, rg = tid >> 4;
        double sre = 0.0, sim = 0.0, sb = 0.0;
        for (int row = rg; row < a.N; row += 16) {
            double vr, vi;
            rhs_elem(rc, row, col, vr, vi);
            double w = row < a.n ? lam[row] * lam[row] : 1.0;
            double xr = Xr[(size_t)row * kLdX + col], xi = Xi[(size_t)row * kLdX + col];
            sre += w * (vr * xr + vi * xi);  // conj(R) X
            sim += w * (vr * xi - vi * xr);
            sb += w * (vr * vr + vi * vi);
        }
        red[(rg * 16 + col) * 3 + 0] = sre; red[(rg * 16 + col) * 3 + 1] = sim; red[(rg * 16 + col) * 3 + 2] = sb;
        consumer_sync();
        if (tid < 16) {
            double cre = 0.0, cim = 0.0, b2 = 0.0;
            for (int r2 = 0; r2 < 16; ++r2) {
                cre += red[(r2 * 16 + tid) * 3 + 0]; cim += red[(r2 * 16 + tid) * 3 + 1]; b2 += red[(r2 * 16 + tid) * 3 + 2];
            }
            cplx c; c.re = cre; c.im = cim;
            cplx th = cg_theta(c, sqrt(b2), 1e-8, 1e-6, 100000);
            theta[2 * tid] = th.re; theta[2 * tid + 1] = th.im;
        }
        consumer_sync();
        for (int e = tid; e < Np * 16; e += kConsumers) {
            int row = e >> 4, col2 = e & 15;
            double xr = Xr[(size_t)row * kLdX + col2], xi = Xi[(size_t)row * kLdX + col2];
            double tr = theta[2 * col2], tim = theta[2 * col2 + 1];
            Xr[(size_t)row * kLdX + col2] = tr * xr - tim * xi;
            Xi[(size_t)row * kLdX + col2] = tr * xi + tim * xr;
        }
        consumer_sync();
    }
    double* Xg = a.X + 2 * ((size_t)sys * a.Tp + (size_t)tile * kTT) * Np;
    double* Sg = a.Ssc ? a.Ssc + 2 * ((size_t)sys * a.Tp + (size_t)tile * kTT) * a.n : nullptr;
    for (int t = 0; t < kTT; ++t) {
        for (int row = tid; row < Np; row += kConsumers) {
            double xr = Xr[(size_t)row * kLdX + t], xi = Xi[(size_t)row * kLdX + t];
            *reinterpret_cast<double2*>(Xg + 2 * ((size_t)t * Np + row)) = make_double2(xr, xi);
            if (Sg && row < a.n) {
                double l = lam[row];
                *reinterpret_cast<double2*>(Sg + 2 * ((size_t)t * a.n + row)) = make_double2(l * xr, l * xi);
            }
        }
    }
    double* Pp = a.Ppart + ((size_t)sys * a.ntiles + tile) * a.n;
    for (int row = tid; row < a.n; row += kConsumers) {
        double s = 0.0;
#pragma unroll
        for (int t = 0; t < kTT; ++t) {
            double xr = Xr[(size_t)row * kLdX + t], xi = Xi[(size_t)row * kLdX + t];
            s += xr * xr + xi * xi;
        }
        Pp[row] = s;
    }
}

void launch_solve(const SolveArgs& a, cudaStream_t st) {
    size_t smem = solve_smem_bytes(a.nblk);
    static size_t attr_smem = 0;
    if (smem > attr_smem) {
        cudaFuncSetAttribute(k_solve, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr_smem = smem;
    }
    k_solve<<<dim3(a.ntiles, a.nsys), 288, smem, st>>>(a);
}

}  // namespace hp
