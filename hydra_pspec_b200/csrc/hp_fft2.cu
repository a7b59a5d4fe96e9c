// hp_fft2.cu -- k_post_fft2: the fused delay-transform kernel with register-resident FFTs.
//
// Same work as k_post_fft (hp_fft.cu; reference: gibbs_step_fgmodes, pspec.py:442-485, and the dense Fourier operator of
// utils.py:14-40 / pspec.py:91-95): per (baseline, time)  s = U^H (lam * ytilde),  model = s + F f,  residual, chi^2,
// ln-posterior partial sums, and |U (w s)|^2 partial sums.  ncu on the round-1 kernel: 81 % L1/TEX throughput, 32 % of the
// shared-memory wavefronts were bank-conflict replays of the five Stockham passes, 13 % DMMA pipe; this version
//
//   * keeps a whole FFT in the registers of one warp (E = Nfreqs / 32 points per lane; radices that divide E, so a lane owns
//     whole butterflies): between two passes the points cross shared memory once (E 16-byte stores + E 16-byte loads per
//     lane, index i stored at i + i / 8: conflict-free for the stride-Ns Stockham stores), with warp-level synchronisation
//     only; 384 points = 6 x 4 x 4 x 4 -> three crossings instead of five passes of load-compute-store with CTA barriers;
//   * reads ytilde straight from global into the first pass and leaves the last pass in registers (signal_cr goes from there
//     to global); the second transform runs its passes in the reverse radix order;
//   * computes F f on the tensor pipe (3M) and finishes the residual in the accumulator layout: the foreground model never
//     goes to shared memory (one buffer instead of two: three CTAs per SM instead of two).
//
// Taken when Nfreqs is a multiple of 32 with E in {4, 8, 12} and a plan of at most four passes exists (128, 256, 384, ...);
// everything else (and the general-basis first iteration) stays on k_post_fft.
#include "hp_kernels.cuh"
#include "hp_math.h"
#include "hp_mma.cuh"
#include <cstdlib>

namespace hp {

namespace {

constexpr int kTP2 = 8;   // times per CTA = warps per CTA
#ifndef HP_FFT2_CTAS
#define HP_FFT2_CTAS 2     // resident CTAs per SM the register budget is sized for
#endif

__device__ __forceinline__ double2 cmul(double2 a, double2 b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ double2 mul_mi(double2 a) { return make_double2(a.y, -a.x); }   // a * (-i)

// in-place forward DFT of R points:  a[q] <- sum_r a[r] exp(-2 pi i q r / R)
template <int R>
__device__ __forceinline__ void butterfly(double2* a);
template <>
__device__ __forceinline__ void butterfly<2>(double2* a) {
    const double2 t = a[1];
    a[1] = csub(a[0], t);
    a[0] = cadd(a[0], t);
}
template <>
__device__ __forceinline__ void butterfly<3>(double2* a) {
    const double S3 = 0.86602540378443864676;
    const double2 t1 = cadd(a[1], a[2]);
    const double2 t2 = make_double2(a[0].x - 0.5 * t1.x, a[0].y - 0.5 * t1.y);
    const double2 t3 = make_double2(S3 * (a[1].y - a[2].y), -S3 * (a[1].x - a[2].x));   // -i sqrt(3)/2 (a1 - a2)
    a[0] = cadd(a[0], t1);
    a[1] = cadd(t2, t3);
    a[2] = csub(t2, t3);
}
template <>
__device__ __forceinline__ void butterfly<4>(double2* a) {
    const double2 b0 = cadd(a[0], a[2]), b1 = csub(a[0], a[2]), b2 = cadd(a[1], a[3]), b3 = mul_mi(csub(a[1], a[3]));
    a[0] = cadd(b0, b2); a[1] = cadd(b1, b3); a[2] = csub(b0, b2); a[3] = csub(b1, b3);
}
template <>
__device__ __forceinline__ void butterfly<6>(double2* a) {
    // 6 = 2 x 3 (Good-Thomas: no twiddles): three-point DFTs of (a0, a2, a4) and (a3, a5, a1), then two-point combinations
    double2 e[3] = {a[0], a[2], a[4]}, o[3] = {a[3], a[5], a[1]};
    butterfly<3>(e);
    butterfly<3>(o);
    // X[q] = E[q mod 3] + (-1)^q O[q mod 3]
    a[0] = cadd(e[0], o[0]); a[3] = csub(e[0], o[0]);
    a[4] = cadd(e[1], o[1]); a[1] = csub(e[1], o[1]);
    a[2] = cadd(e[2], o[2]); a[5] = csub(e[2], o[2]);
}
template <>
__device__ __forceinline__ void butterfly<8>(double2* a) {
    const double H = 0.70710678118654752440;
    double2 e[4] = {a[0], a[2], a[4], a[6]}, o[4] = {a[1], a[3], a[5], a[7]};
    butterfly<4>(e);
    butterfly<4>(o);
    // twiddles exp(-2 pi i q / 8), q = 0..3:  1,  (1 - i)/sqrt2,  -i,  (-1 - i)/sqrt2
    const double2 o1 = make_double2(H * (o[1].x + o[1].y), H * (o[1].y - o[1].x));
    const double2 o2 = mul_mi(o[2]);
    const double2 o3 = make_double2(H * (o[3].y - o[3].x), -H * (o[3].x + o[3].y));
    a[0] = cadd(e[0], o[0]); a[4] = csub(e[0], o[0]);
    a[1] = cadd(e[1], o1);   a[5] = csub(e[1], o1);
    a[2] = cadd(e[2], o2);   a[6] = csub(e[2], o2);
    a[3] = cadd(e[3], o3);   a[7] = csub(e[3], o3);
}

__device__ __forceinline__ int phys(int i) { return i + (i >> 3); }

// One Stockham pass of one warp's FFT.  v[u * R + r] = input r of butterfly j = lane + 32 u (index j + r * nb, nb = n / R).
//   kLoad : fetch the inputs from `row` (shared memory, index i at phys(i)); else they are in v already
//   kStore: write the outputs to `row` (index jq Ns R + k + r Ns; k = j mod Ns, jq = j / Ns) between two __syncwarp;
//           else they stay in v as output r of butterfly j = index j + r * Ns (only the last pass: Ns R = n)
template <int E, int R, bool kLoad, bool kStore>
__device__ __forceinline__ void fft_pass(double2 (&v)[E], double2* row, int n, int Ns, uint32_t magic, const double2* __restrict__ tw,
                                         int lane) {
    constexpr int NB = E / R;   // butterflies per lane
    const int nb = n / R;
    if (kLoad) {
#pragma unroll
        for (int u = 0; u < NB; ++u)
#pragma unroll
            for (int r = 0; r < R; ++r) v[u * R + r] = row[phys(lane + 32 * u + r * nb)];
    }
    const int tstep = nb / Ns;   // n / (Ns R)
#pragma unroll
    for (int u = 0; u < NB; ++u) {
        const int j = lane + 32 * u;
        const int jq = Ns > 1 ? (int)__umulhi((uint32_t)j, magic) : j;   // j / Ns
        const int k = j - jq * Ns;
        if (Ns > 1 && k) {
#pragma unroll
            for (int r = 1; r < R; ++r) v[u * R + r] = cmul(v[u * R + r], tw[r * k * tstep]);
        }
        butterfly<R>(&v[u * R]);
    }
    if (kStore) {
        __syncwarp();   // every lane has its inputs in registers: the row may be overwritten
#pragma unroll
        for (int u = 0; u < NB; ++u) {
            const int j = lane + 32 * u;
            const int jq = Ns > 1 ? (int)__umulhi((uint32_t)j, magic) : j;
            const int k = j - jq * Ns;
#pragma unroll
            for (int r = 0; r < R; ++r) row[phys(jq * Ns * R + k + r * Ns)] = v[u * R + r];
        }
        __syncwarp();
    }
}

// pass p of a plan, radix chosen at run time among the divisors of E
template <int E, bool kLoad, bool kStore>
__device__ __forceinline__ void fft_pass_any(int R, double2 (&v)[E], double2* row, int n, int Ns, uint32_t magic,
                                             const double2* __restrict__ tw, int lane) {
    if (R == 4) { if constexpr (E % 4 == 0) fft_pass<E, 4, kLoad, kStore>(v, row, n, Ns, magic, tw, lane); }
    else if (R == 8) { if constexpr (E % 8 == 0) fft_pass<E, 8, kLoad, kStore>(v, row, n, Ns, magic, tw, lane); }
    else if (R == 6) { if constexpr (E % 6 == 0) fft_pass<E, 6, kLoad, kStore>(v, row, n, Ns, magic, tw, lane); }
    else if (R == 2) { if constexpr (E % 2 == 0) fft_pass<E, 2, kLoad, kStore>(v, row, n, Ns, magic, tw, lane); }
    else if (R == 3) { if constexpr (E % 3 == 0) fft_pass<E, 3, kLoad, kStore>(v, row, n, Ns, magic, tw, lane); }
}

}  // namespace

// radices (each dividing E = n / 32, at most four passes) for the register-resident FFT; false: use k_post_fft
bool make_fft2_plan(int n, FftPlan* fwd, FftPlan* rev) {
    if (n % 32 != 0) return false;
    const int E = n / 32;
    if (E != 4 && E != 8 && E != 12) return false;
    int rad[8], nf = 0, rem = n;
    const int cand[5] = {8, 6, 4, 3, 2};
    while (rem > 1 && nf < 8) {
        int pick = 0;
        for (int c : cand)
            if (E % c == 0 && rem % c == 0) { pick = c; break; }
        if (!pick) return false;
        rad[nf++] = pick;
        rem /= pick;
    }
    if (rem != 1 || nf < 2 || nf > 4) return false;
    // forward plan: smallest radix last would leave few outputs per butterfly in registers; order as found (largest first)
    for (int dir = 0; dir < 2; ++dir) {
        FftPlan* p = dir ? rev : fwd;
        p->n = n; p->nf = nf;
        int Ns = 1;
        for (int i = 0; i < nf; ++i) {
            p->radix[i] = dir ? rad[nf - 1 - i] : rad[i];
            p->magic[i] = Ns > 1 ? (uint32_t)((0x100000000ull + (uint64_t)Ns - 1) / (uint64_t)Ns) : 0u;
            Ns *= p->radix[i];
        }
    }
    return true;
}

size_t postfft2_smem_bytes(int n, int m) {
    const int mk = ((m + 3) / 4) * 4;
    const int ld = n + n / 8 + 1;   // padded row: index i at i + i / 8, row stride == 1 (mod 8) sixteen-byte units
    return sizeof(double2) * ((size_t)kTP2 * ld + n + (size_t)kTP2 * (mk + 1)) + 80 * sizeof(double);
}

template <int E>
__global__ void __launch_bounds__(32 * kTP2, HP_FFT2_CTAS) k_post_fft2(PostFftArgs a, FftPlan prev) {
    constexpr int kThreads = 32 * kTP2;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const FftPlan& pfwd = a.plan;
    const int n = pfwd.n, m = a.m, mk = ((m + 3) / 4) * 4, ldf = mk + 1, ld = n + n / 8 + 1;
    double2* A = reinterpret_cast<double2*>(smem_raw);          // [8][ld]: one FFT row per warp; later s in frequency space
    double2* tw = A + (size_t)kTP2 * ld;                         // [n]
    double2* fs = tw + n;                                        // [8][ldf] foreground amplitudes of the tile's times
    double* red = reinterpret_cast<double*>(fs + (size_t)kTP2 * ldf);   // [8 warps][8 times]
    const int sys = blockIdx.y, tile = blockIdx.x, t0 = tile * kTP2;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const double* X = a.X + 2 * ((size_t)sys * a.Tp + t0) * a.Np;
    const double* lam = a.lam + (size_t)sys * a.Np;
    const double* w = a.w + (size_t)sys * a.w_bs + (size_t)t0 * a.w_ts;
    const double* nd = a.ninvd + (size_t)sys * n;
    const double* wd = a.wd + 2 * ((size_t)sys * a.Tp + t0) * n;
    double* Sf = a.Sf + 2 * ((size_t)sys * a.sf_bs + (size_t)t0 * n);
    const double2* twg = reinterpret_cast<const double2*>(a.tw);
    const double rsn = rsqrt((double)n);

    {   // the data rows are consumed in the residual phase: start them on their way from DRAM to L2 now
        const char* wdp = reinterpret_cast<const char*>(wd);
        const size_t bytes = (size_t)min(kTP2, a.T - t0) * n * 16;
        for (size_t off = (size_t)tid * 128; off < bytes; off += kThreads * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(wdp + off));
    }
    // ---- first pass inputs straight from global:  conj(lam * ytilde) * (-1)^k  (U^H a = conj(U conj(a)); n is even)
    const int t = warp;                    // this warp's time
    const bool live = t0 + t < a.T;
    double2 v[E];
    {
        const int R0 = pfwd.radix[0], nb0 = n / R0;
        const double* xr = X + 2 * (size_t)t * a.Np;
#pragma unroll
        for (int e = 0; e < E; ++e) {
            // element e = u * R0 + r  <->  index (lane + 32 u) + r * nb0
            const int u = e / R0, r = e - u * R0;   // (R0 is warp-uniform; E / R0 butterflies per lane)
            const int k = lane + 32 * u + r * nb0;
            double2 y = make_double2(0.0, 0.0);
            double l = 0.0;
            if (live) { y = *reinterpret_cast<const double2*>(xr + 2 * k); l = lam[k]; }
            const double sg = (k & 1) ? -l : l;
            v[e] = make_double2(sg * y.x, -sg * y.y);
        }
    }
    for (int j = tid; j < n; j += kThreads) tw[j] = twg[j];
    for (int e = tid; e < kTP2 * ldf; e += kThreads) {
        const int tt = e / ldf, j = e - tt * ldf;
        double2 f = make_double2(0.0, 0.0);
        if (j < m && t0 + tt < a.T) f = *reinterpret_cast<const double2*>(X + 2 * ((size_t)tt * a.Np + n + j));
        fs[e] = f;
        if (j < m && t0 + tt < a.T && a.fg_out)
            *reinterpret_cast<double2*>(a.fg_out + (size_t)sys * a.fg_bs + 2 * ((size_t)(t0 + tt) * m + j)) = f;
    }
    __syncthreads();   // twiddles and f in shared memory
    double2* row = A + (size_t)t * ld;
    // ---- FFT 1 (forward plan): first pass from registers, middle passes through the row, last pass stays in registers
    {
        const int nf = pfwd.nf;
        int Ns = 1;
        fft_pass_any<E, false, true>(pfwd.radix[0], v, row, n, Ns, pfwd.magic[0], tw, lane);
        Ns *= pfwd.radix[0];
        for (int p = 1; p < nf - 1; ++p) {
            fft_pass_any<E, true, true>(pfwd.radix[p], v, row, n, Ns, pfwd.magic[p], tw, lane);
            Ns *= pfwd.radix[p];
        }
        fft_pass_any<E, true, false>(pfwd.radix[nf - 1], v, row, n, Ns, pfwd.magic[nf - 1], tw, lane);
        // outputs: element u * R + r = index x = lane + 32 u + r * Ns (Ns = n / R).  s = conj(out * (-1)^x * c0) / sqrt(n)
        const int R = pfwd.radix[nf - 1];
        const double2 c0 = tw[(int)(((long long)(n / 2) * (n / 2)) % n)];   // exp(-2 pi i h^2 / n), h = n / 2
        __syncwarp();   // the last pass has read the row
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const int u = e / R, r = e - u * R;
            const int x = lane + 32 * u + r * Ns;
            double2 o = cmul(v[e], c0);
            const double sg = (x & 1) ? -rsn : rsn;
            o = make_double2(sg * o.x, -sg * o.y);
            row[phys(x)] = o;
            if (live) *reinterpret_cast<double2*>(Sf + 2 * ((size_t)t * n + x)) = o;
        }
    }
    __syncthreads();   // s of all eight times is in A
    // ---- foreground model F f on the tensor pipe (3M), finished in the accumulator layout:
    //      lane (g, q) holds (F f)[time g][x = 8 ct + 2 q + e]; residual, chi^2, ln-posterior partials, masked signal
    {
        const double* Ft = a.Ft + 2 * (size_t)sys * m * n;
        const int nct = n / 8;
        const int g = lane >> 2, q = lane & 3;
        const bool tlive = t0 + g < a.T;
        const double* wrow = w + (size_t)g * a.w_ts;
        double lnp = 0.0;
        double2 bnx[4];
        auto load_b = [&](int ct, int kc, double2 (&b)[4]) {
            const int x = 8 * ct + g;
#pragma unroll
            for (int s4 = 0; s4 < 4; ++s4) {
                const int k = kc + 4 * s4 + q;
                b[s4] = (k < m && ct < nct) ? *reinterpret_cast<const double2*>(Ft + 2 * ((size_t)k * n + x)) : make_double2(0.0, 0.0);
            }
        };
        load_b(warp, 0, bnx);
        for (int ct = warp; ct < nct; ct += kTP2) {
            double p1[2] = {0.0, 0.0}, p2[2] = {0.0, 0.0}, p3[2] = {0.0, 0.0};
            // this tile's data and mask (32 contiguous bytes per lane), in flight during the products
            const int x0 = 8 * ct + 2 * q;
            double2 d0 = make_double2(0.0, 0.0), d1 = d0;
            if (tlive) {
                d0 = *reinterpret_cast<const double2*>(wd + 2 * ((size_t)g * n + x0));
                d1 = *reinterpret_cast<const double2*>(wd + 2 * ((size_t)g * n + x0 + 1));
            }
            for (int kc = 0; kc < mk; kc += 16) {   // chunks of four k-steps: A fragments from shared memory, B one chunk ahead
                double2 bc[4];
#pragma unroll
                for (int s4 = 0; s4 < 4; ++s4) bc[s4] = bnx[s4];
                if (kc + 16 < mk) load_b(ct, kc + 16, bnx); else load_b(ct + kTP2, 0, bnx);
#pragma unroll
                for (int s4 = 0; s4 < 4; ++s4) {
                    const int k = kc + 4 * s4 + q;
                    const double2 af = (k < mk) ? fs[(size_t)g * ldf + k] : make_double2(0.0, 0.0);
                    dmma884(p1[0], p1[1], af.x, bc[s4].x);
                    dmma884(p2[0], p2[1], af.y, bc[s4].y);
                    dmma884(p3[0], p3[1], af.x + af.y, bc[s4].x + bc[s4].y);
                }
            }
            const double2 s0 = A[(size_t)g * ld + phys(x0)], s1 = A[(size_t)g * ld + phys(x0 + 1)];
            const double wx0 = (a.w_ts == 0 || tlive) ? wrow[x0] : 0.0, wx1 = (a.w_ts == 0 || tlive) ? wrow[x0 + 1] : 0.0;
            const double n0 = nd[x0], n1 = nd[x0 + 1];
            const double rr0 = d0.x - s0.x - (p1[0] - p2[0]), ri0 = d0.y - s0.y - (p3[0] - p1[0] - p2[0]);
            const double rr1 = d1.x - s1.x - (p1[1] - p2[1]), ri1 = d1.y - s1.y - (p3[1] - p1[1] - p2[1]);
            const double q0 = rr0 * rr0 + ri0 * ri0, q1 = rr1 * rr1 + ri1 * ri1;
            if (tlive) {
                if (a.chisq_out)
                    *reinterpret_cast<double2*>(a.chisq_out + (size_t)sys * a.chisq_bs + (size_t)(t0 + g) * n + x0) = make_double2(q0 * n0, q1 * n1);
                lnp += wx0 * n0 * q0 + wx1 * n1 * q1;
                if (a.Rm) {
                    double* rm = a.Rm + 2 * (((size_t)sys * a.Tp + t0 + g) * n + x0);
                    *reinterpret_cast<double2*>(rm) = make_double2(wx0 * rr0, wx0 * ri0);
                    *reinterpret_cast<double2*>(rm + 2) = make_double2(wx1 * rr1, wx1 * ri1);
                }
            }
            if (a.Empart) {
                // input of the second transform, in place: w s (-1)^x  (x0 is even)
                A[(size_t)g * ld + phys(x0)] = make_double2(wx0 * s0.x, wx0 * s0.y);
                A[(size_t)g * ld + phys(x0 + 1)] = make_double2(-wx1 * s1.x, -wx1 * s1.y);
            }
        }
        lnp += __shfl_xor_sync(0xffffffffu, lnp, 1);
        lnp += __shfl_xor_sync(0xffffffffu, lnp, 2);
        if (q == 0) red[warp * kTP2 + g] = lnp;
    }
    __syncthreads();
    if (tid < kTP2) {
        double s = 0.0;
        for (int wv = 0; wv < kTP2; ++wv) s += red[wv * kTP2 + tid];
        if (t0 + tid < a.Tp) a.lnp1[(size_t)sys * a.Tp + t0 + tid] = t0 + tid < a.T ? s : 0.0;
    }
    // ---- |U (w s)|^2 summed over the tile's times (second term of ln_post, pspec.py:479-483): reverse plan
    if (a.Empart) {
        const int nf = prev.nf;
        int Ns = 1;
        for (int p = 0; p < nf - 1; ++p) {
            fft_pass_any<E, true, true>(prev.radix[p], v, row, n, Ns, prev.magic[p], tw, lane);
            Ns *= prev.radix[p];
        }
        fft_pass_any<E, true, false>(prev.radix[nf - 1], v, row, n, Ns, prev.magic[nf - 1], tw, lane);
        const int R = prev.radix[nf - 1];
        __syncwarp();   // the last pass has read the row: park |.|^2 in it (doubles, index k)
        double* erow = reinterpret_cast<double*>(row);
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const int u = e / R, r = e - u * R;
            erow[lane + 32 * u + r * Ns] = live ? v[e].x * v[e].x + v[e].y * v[e].y : 0.0;
        }
        __syncthreads();
        double* Ep = a.Empart + ((size_t)sys * gridDim.x + tile) * n;
        for (int k = tid; k < n; k += kThreads) {
            double acc = 0.0;
#pragma unroll
            for (int tt = 0; tt < kTP2; ++tt) acc += reinterpret_cast<const double*>(A + (size_t)tt * ld)[k];
            Ep[k] = acc / (double)n;
        }
    }
}

// true: the launch was taken by k_post_fft2
bool launch_post_fft2(const PostFftArgs& a, const FftPlan& fwd, const FftPlan& rev, cudaStream_t st) {
    const int n = fwd.n, E = n / 32;
    const size_t smem = postfft2_smem_bytes(n, a.m);
    static size_t attr_dev[kMaxDev][3] = {{0}};
    const int slot = E == 4 ? 0 : (E == 8 ? 1 : 2);
    size_t& attr = attr_dev[current_device_slot()][slot];
    PostFftArgs b = a;
    b.plan = fwd;
    const dim3 grid((a.T + kTP2 - 1) / kTP2, a.nsys);
    if (E == 4) {
        if (smem > attr) { cudaFuncSetAttribute(k_post_fft2<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); attr = smem; }
        k_post_fft2<4><<<grid, 32 * kTP2, smem, st>>>(b, rev);
    } else if (E == 8) {
        if (smem > attr) { cudaFuncSetAttribute(k_post_fft2<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); attr = smem; }
        k_post_fft2<8><<<grid, 32 * kTP2, smem, st>>>(b, rev);
    } else if (E == 12) {
        if (smem > attr) { cudaFuncSetAttribute(k_post_fft2<12>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); attr = smem; }
        k_post_fft2<12><<<grid, 32 * kTP2, smem, st>>>(b, rev);
    } else {
        return false;
    }
    return true;
}

}  // namespace hp
