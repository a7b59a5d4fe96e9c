// hp_solve2.cu -- k_solve2: persistent, register-blocked GCR solve  x = W^H (W r + xi)  for all times of all baselines.
//
// Replaces the per-time preconditioned CG of the reference (gcr_fgmodes_1d, pspec.py:151-235) like k_solve (hp_solve.cu),
// with the structure the ncu counters of round 1 asked for (profiles/r1_v9_summary.md: 4 LDS per 3 DMMA, math-pipe throttle
// from the 3M operand sums, staging + epilogue + ring refill exposed once per tile because nothing else fits on the SM):
//
//   * one persistent CTA per SM walks the (baseline, 16-time tile) list; the W-block ring never drains between tiles;
//   * eight consumer warps, each a 16 x 16 complex register tile (WI = WJ = 2): one A/B fragment pair feeds four DMMA
//     triplets, i.e. 8 LDS + 4 DADD per 12 DMMA instead of 16 + 8;
//   * the block row (32 x 16) is two such tiles (row halves) x four K slices; k-steps are dealt round-robin to the slices so
//     that triangular diagonal blocks stay balanced; the partial sums meet in a double-buffered exchange area and each warp
//     finishes one 8 x 8 quarter -- lazily, after the first block of the NEXT row, so no warp waits at the end of a row;
//   * right-hand sides are stored in the tile's own shared-memory layout (k_rhs_tile), so a block row is one 8 KiB TMA bulk
//     copy; the rows of the next tile are fetched into the rows the second pass has finished with (a second producer warp),
//     and pass 1 consumes its first block row in the order the rows arrive;
//   * pass 2 never writes shared memory: x goes from the accumulators to global, sum_t |x|^2 with it.
//
// Pass 1 (y = W1 r + xi, in place, block rows downwards, diagonal block last) uses W1 = W diag(lam) when the stored
// right-hand sides are unscaled (Philox mode: r = lam * Rfix), so no per-iteration pass over Rfix exists anywhere.
#include "hp_kernels.cuh"
#include "hp_math.h"
#include "hp_mma.cuh"
#include "hp_async.cuh"
#include <cstdlib>

namespace hp {

using namespace async;

namespace {

constexpr int kCW = 8;                        // consumer warps: 2 row halves x 4 K slices
constexpr int kS2Threads = 32 * (kCW + 2);    // + W producer warp + right-hand-side producer warp
constexpr int kMaxStages2 = 6;
constexpr int kMaxBlk2 = 20;
constexpr int kRowDoubles = 2 * 32 * kTT;     // one block row of the tile: [plane][32 rows][16 columns, XOR-swizzled]
constexpr int kExchDoubles = 2 * 4 * 3 * 128; // per parity: [row half][destination quarter][source slot][e][lane][re, im]

__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

struct Frag { double ar[2], ai[2], br[2], bi[2]; };

// A operand: pass 1  A[row][k] = blk[row * 36 + k];  pass 2  A[row][k] = conj(W[k][row]) = blk[k * 36 + row]
template <bool kP2>
__device__ __forceinline__ void load_frag(Frag& f, const double* __restrict__ a, const double* __restrict__ b, int bo0, int bo1) {
    constexpr int wistep = kP2 ? 8 : 8 * kLdBlk;
    f.ar[0] = a[0]; f.ai[0] = a[kLPlane];
    f.ar[1] = a[wistep]; f.ai[1] = a[kLPlane + wistep];
    f.br[0] = b[bo0]; f.bi[0] = b[bo0 + 32 * kTT];
    f.br[1] = b[bo1]; f.bi[1] = b[bo1 + 32 * kTT];
}
// 3M:  P0 += Ar.Br,  P1 += Ai.Bi,  P2 += (Ar +- Ai).(Br + Bi)
template <bool kP2>
__device__ __forceinline__ void mma_frag(double (&P)[3][2][2][2], const Frag& f) {
    double as[2], bs[2];
#pragma unroll
    for (int wi = 0; wi < 2; ++wi) as[wi] = kP2 ? f.ar[wi] - f.ai[wi] : f.ar[wi] + f.ai[wi];
#pragma unroll
    for (int wj = 0; wj < 2; ++wj) bs[wj] = f.br[wj] + f.bi[wj];
#pragma unroll
    for (int wi = 0; wi < 2; ++wi)
#pragma unroll
        for (int wj = 0; wj < 2; ++wj) dmma884(P[0][wi][wj][0], P[0][wi][wj][1], f.ar[wi], f.br[wj]);
#pragma unroll
    for (int wi = 0; wi < 2; ++wi)
#pragma unroll
        for (int wj = 0; wj < 2; ++wj) dmma884(P[1][wi][wj][0], P[1][wi][wj][1], f.ai[wi], f.bi[wj]);
#pragma unroll
    for (int wi = 0; wi < 2; ++wi)
#pragma unroll
        for (int wj = 0; wj < 2; ++wj) dmma884(P[2][wi][wj][0], P[2][wi][wj][1], as[wi], bs[wj]);
}

}  // namespace

static size_t solve2_smem_bytes(int nblk, int stages) {
    return sizeof(double) * ((size_t)stages * kLBlkDoubles + (size_t)nblk * kRowDoubles + 2 * kExchDoubles) +
           sizeof(uint64_t) * (2 * kMaxStages2 + 2 * kMaxBlk2 + 2) + 128;
}
int solve2_stages(int nblk, size_t max_smem) {
    if (nblk > kMaxBlk2) return 0;
    int s = kMaxStages2;
    while (s >= 2 && solve2_smem_bytes(nblk, s) > max_smem) --s;
    return s >= 2 ? s : 0;
}

// Phase timers (HP_S2_TIMERS=1 selects the instrumented instantiation; profiles/scripts/solve2_timers.py): clock64 deltas of
// lane 0 of every consumer warp, summed over warps and CTAs.
//  [0] waiting for a W block  [1] waiting for right-hand-side rows  [2] waiting for the exchange barrier  [3] Philox draws
//  [4] hand-over  [5] collect + finish (without [2])  [6] pass 1  [7] pass 2  [8] consumer barrier between the passes
//  [9] blocks acquired  [10] early probes that succeeded  [11] whole consumer loop
__device__ unsigned long long g_solve2_cycles[16];

template <bool kTimers>
__global__ void __launch_bounds__(kS2Threads, 1) k_solve2(Solve2Args a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int nblk = a.nblk, Np = nblk * 32, S = a.stages;
    double* ring = reinterpret_cast<double*>(smem_raw);                  // [S][2304]
    double* tileS = ring + (size_t)S * kLBlkDoubles;                     // [nblk][2][32][16]
    double* exch = tileS + (size_t)nblk * kRowDoubles;                   // [2][kExchDoubles]
    uint64_t* full = reinterpret_cast<uint64_t*>(exch + 2 * kExchDoubles);
    uint64_t* empty = full + kMaxStages2;
    uint64_t* rhs_full = empty + kMaxStages2;                            // [nblk] block row j of the tile has landed
    uint64_t* rowfree = rhs_full + kMaxBlk2;                             // [nblk] pass 2 is done with block row j
    uint64_t* exbar = rowfree + kMaxBlk2;                                // [2] partial sums of a block row are in exch[parity]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int total = a.ntiles * a.nsys;
    const size_t tri = tri_blocks(nblk);

    if (tid == 0) {
        for (int s = 0; s < S; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], kCW); }
        for (int j = 0; j < nblk; ++j) { mbar_init(&rhs_full[j], 1); mbar_init(&rowfree[j], kCW); }
        mbar_init(&exbar[0], kCW);
        mbar_init(&exbar[1], kCW);
        mbar_fence_init();
    }
    __syncthreads();

    if (warp == kCW) {
        // ------------------------------------------------------------------ W producer
        if (lane == 0) {
            uint32_t s = 0, round = 0;
            auto push = [&](const double* src) {
                if (round > 0) mbar_wait(&empty[s], (round - 1) & 1);
                mbar_arrive_expect_tx(&full[s], kLBlkDoubles * 8);
                bulk_g2s(ring + (size_t)s * kLBlkDoubles, src, kLBlkDoubles * 8, &full[s]);
                if (++s == (uint32_t)S) { s = 0; ++round; }
            };
            for (int w = blockIdx.x; w < total; w += gridDim.x) {
                const int sys = w / a.ntiles;
                const double* W1 = a.W1 + (size_t)sys * tri * kLBlkDoubles;
                const double* W2 = a.W2 + (size_t)sys * tri * kLBlkDoubles;
                for (int i = nblk - 1; i >= 0; --i)
                    for (int j = 0; j <= i; ++j) push(W1 + blk_index(i, j) * kLBlkDoubles);
                for (int i = 0; i < nblk; ++i)
                    for (int j = i; j < nblk; ++j) push(W2 + blk_index(j, i) * kLBlkDoubles);
            }
        }
        return;
    }
    if (warp == kCW + 1) {
        // ------------------------------------------------------------------ right-hand-side producer
        if (lane == 0) {
            uint32_t it = 0;
            for (int w = blockIdx.x; w < total; w += gridDim.x, ++it) {
                const double* src = a.Rt + (size_t)w * nblk * kRowDoubles;
                for (int j = 0; j < nblk; ++j) {
                    if (it > 0) mbar_wait(&rowfree[j], (it - 1) & 1);   // the previous tile's pass 2 no longer reads row j
                    mbar_arrive_expect_tx(&rhs_full[j], kRowDoubles * 8);
                    bulk_g2s(tileS + (size_t)j * kRowDoubles, src + (size_t)j * kRowDoubles, kRowDoubles * 8, &rhs_full[j]);
                }
            }
        }
        return;
    }

    // ---------------------------------------------------------------------- consumer warps
    const int g = lane >> 2, q = lane & 3;
    const int tg = warp >> 2, ks = warp & 3;    // row half of the block row, K slice (k-steps ks and ks + 4 of every block)
    const int mwi = ks >> 1, mwj = ks & 1;      // the 8 x 8 quarter of the 16 x 16 tile this warp finishes
    // per-lane fragment offsets of k-step ks (k-step ks + 4: + 16 columns of A / + 16 rows of A^T / + 16 rows of the tile)
    const int aoff1 = (16 * tg + g) * kLdBlk + 4 * ks + q;
    const int aoff2 = (4 * ks + q) * kLdBlk + 16 * tg + g;
    const int bo0 = (4 * ks + q) * kTT + (g ^ (q << 2));
    const int bo1 = (4 * ks + q) * kTT + ((8 + g) ^ (q << 2));
    // the two elements (e = 0, 1) of the quarter this lane finishes: row `frow` of the block row, columns fcol, fcol + 1
    const int frow = 16 * tg + 8 * mwi + g, fcol = 8 * mwj + 2 * q;

    uint32_t slot = 0, parity = 0;
    bool ready = false, ready_next = false;
    uint32_t rowctr = 0, it = 0;
    long long tacc[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
#define S2_T0() long long t0_ = 0; if (kTimers) t0_ = clock64()
#define S2_T1(idx) do { if (kTimers) tacc[idx] += clock64() - t0_; } while (0)
    long long tloop0 = 0;
    if (kTimers) tloop0 = clock64();
    auto acquire = [&]() -> const double* {
        if (kTimers) { tacc[9] += 1; tacc[10] += ready ? 1 : 0; }
        if (!ready) { S2_T0(); mbar_wait(&full[slot], parity); S2_T1(0); }
        return ring + (size_t)slot * kLBlkDoubles;
    };
    auto probe_next = [&]() {
        uint32_t ns = slot + 1, np = parity;
        if (ns == (uint32_t)S) { ns = 0; np ^= 1; }
        ready_next = mbar_test(&full[ns], np);
    };
    auto release = [&]() {
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[slot]);
        if (++slot == (uint32_t)S) { slot = 0; parity ^= 1; }
        ready = ready_next;
    };

    double P[3][2][2][2];
    auto zeroP = [&]() {
#pragma unroll
        for (int p = 0; p < 3; ++p)
#pragma unroll
            for (int wi = 0; wi < 2; ++wi)
#pragma unroll
                for (int wj = 0; wj < 2; ++wj) P[p][wi][wj][0] = P[p][wi][wj][1] = 0.0;
    };
    // end of a block row: this warp's partial sums of the three foreign quarters go to the exchange buffer, its own quarter
    // stays in (mre, mim)
    double mre[2], mim[2];
    auto hand_over = [&](bool p2) {
        S2_T0();
        double* buf = exch + (size_t)(rowctr & 1) * kExchDoubles + (size_t)tg * 4 * 3 * 128;
#pragma unroll
        for (int qd = 0; qd < 4; ++qd) {
            const int wi = qd >> 1, wj = qd & 1;
            double re[2], im[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const double p0 = P[0][wi][wj][e], p1 = P[1][wi][wj][e], p2v = P[2][wi][wj][e];
                re[e] = p2 ? p0 + p1 : p0 - p1;              // conj(A) B : A B
                im[e] = p2 ? p2v - p0 + p1 : p2v - p0 - p1;
            }
            if (qd == ks) {
                mre[0] = re[0]; mim[0] = im[0]; mre[1] = re[1]; mim[1] = im[1];
            } else {
                double* d = buf + (size_t)(qd * 3 + ((ks - qd - 1) & 3)) * 128 + 2 * lane;
                *reinterpret_cast<double2*>(d) = make_double2(re[0], im[0]);
                *reinterpret_cast<double2*>(d + 64) = make_double2(re[1], im[1]);
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&exbar[rowctr & 1]);
        S2_T1(4);
    };
    // ... and, once every warp has handed over, the three foreign contributions to this warp's quarter are added
    auto collect = [&](uint32_t rc) {
        { S2_T0(); mbar_wait(&exbar[rc & 1], (rc >> 1) & 1); S2_T1(2); }
        const double* buf = exch + (size_t)(rc & 1) * kExchDoubles + (size_t)(tg * 4 + ks) * 3 * 128 + 2 * lane;
#pragma unroll
        for (int sl = 0; sl < 3; ++sl) {
            const double2 v0 = *reinterpret_cast<const double2*>(buf + sl * 128);
            const double2 v1 = *reinterpret_cast<const double2*>(buf + sl * 128 + 64);
            mre[0] += v0.x; mim[0] += v0.y; mre[1] += v1.x; mim[1] += v1.y;
        }
    };

    for (int w = blockIdx.x; w < total; w += gridDim.x, ++it) {
        const int sys = w / a.ntiles, tile = w - sys * a.ntiles, t0 = tile * kTT;
        const uint32_t chain = a.chain_ids ? (uint32_t)a.chain_ids[sys] : (uint32_t)(a.chain0 + sys);

        // ------------------------------------------------------------------ pass 1:  y = W1 r (+ xi), in place
        int pend = -1;                 // block row whose quarter is still to be finished
        uint32_t pend_rc = 0;
        double xi[2][2] = {{0.0, 0.0}, {0.0, 0.0}}, xi_pend[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
        long long tpass = 0;
        if (kTimers) tpass = clock64();
        auto finish1 = [&]() {
            S2_T0();
            collect(pend_rc);
            double* trow = tileS + (size_t)pend * kRowDoubles + frow * kTT;
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int c = (fcol + e) ^ ((frow & 3) << 2);
                trow[c] = mre[e] + xi_pend[e][0];
                trow[c + 32 * kTT] = mim[e] + xi_pend[e][1];
            }
            S2_T1(5);
        };
        for (int i = nblk - 1; i >= 0; --i) {
            zeroP();
            // fluctuation draws of the two elements this lane finishes: independent of the products, issued first so that
            // their integer / MUFU work fills issue slots between the row's DMMAs
            if (a.philox) {
                S2_T0();
                const int row = 32 * i + frow;
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    xi[e][0] = xi[e][1] = 0.0;
                    if (row < a.N && t0 + fcol + e < a.T) {
                        u32x4 ctr; ctr.x = (uint32_t)row; ctr.y = (uint32_t)(t0 + fcol + e); ctr.z = a.iter; ctr.w = chain;
                        normal_pair_fast(philox4x32_10(ctr, a.key0, a.key1 ^ 0xA5A5A5A5u), xi[e][0], xi[e][1]);
                        xi[e][0] *= 0.70710678118654752440; xi[e][1] *= 0.70710678118654752440;
                    }
                }
                S2_T1(3);
            }
            for (int j = 0; j <= i; ++j) {
                if (i == nblk - 1) { S2_T0(); mbar_wait(&rhs_full[j], it & 1); S2_T1(1); }   // first block row of a tile: rows in arrival order
                const double* blk = acquire();
                probe_next();
                const double* bt = tileS + (size_t)j * kRowDoubles;
                // diagonal block (lower triangular): rows of half tg need k < 16 (tg + 1)
                const bool second = (j < i) || tg == 1;
                Frag f0, f1;
                load_frag<false>(f0, blk + aoff1, bt, bo0, bo1);
                if (second) load_frag<false>(f1, blk + aoff1 + 16, bt + 16 * kTT, bo0, bo1);
                mma_frag<false>(P, f0);
                if (second) mma_frag<false>(P, f1);
                release();
                if (j == 0 && pend >= 0) { finish1(); pend = -1; }
            }
            hand_over(false);
            pend = i; pend_rc = rowctr++;
#pragma unroll
            for (int e = 0; e < 2; ++e) { xi_pend[e][0] = xi[e][0]; xi_pend[e][1] = xi[e][1]; }
        }
        finish1();
        { S2_T0(); consumer_sync(); S2_T1(8); }   // y complete and visible to all consumer warps
        if (kTimers) { const long long now = clock64(); tacc[6] += now - tpass; tpass = now; }

        // ------------------------------------------------------------------ pass 2:  x = W2^H y, straight to global
        pend = -1;
        double* Xg = a.X + 2 * ((size_t)sys * a.Tp + t0) * Np;
        double* Pp = a.Ppart + ((size_t)sys * 2 * a.ntiles + 2 * tile + mwj) * a.n;
        auto finish2 = [&]() {
            S2_T0();
            collect(pend_rc);
            const int row = 32 * pend + frow;
            double p = 0.0;
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                *reinterpret_cast<double2*>(Xg + 2 * ((size_t)(fcol + e) * Np + row)) = make_double2(mre[e], mim[e]);
                p += mre[e] * mre[e] + mim[e] * mim[e];
            }
            p += __shfl_xor_sync(0xffffffffu, p, 1);
            p += __shfl_xor_sync(0xffffffffu, p, 2);
            if (q == 0 && row < a.n) Pp[row] = p;
            S2_T1(5);
        };
        for (int i = 0; i < nblk; ++i) {
            zeroP();
            for (int j = i; j < nblk; ++j) {
                const double* blk = acquire();
                probe_next();
                const double* bt = tileS + (size_t)j * kRowDoubles;
                // diagonal block W_ii^H (upper triangular): rows of half tg need k >= 16 tg
                const bool first = (j > i) || tg == 0;
                Frag f0, f1;
                if (first) load_frag<true>(f0, blk + aoff2, bt, bo0, bo1);
                load_frag<true>(f1, blk + aoff2 + 16 * kLdBlk, bt + 16 * kTT, bo0, bo1);
                if (first) mma_frag<true>(P, f0);
                mma_frag<true>(P, f1);
                release();
                if (j == i) {
                    // y_i is not read again: the next tile's block row i may land on it
                    if (lane == 0) mbar_arrive(&rowfree[i]);
                    if (pend >= 0) { finish2(); pend = -1; }
                }
            }
            hand_over(true);
            pend = i; pend_rc = rowctr++;
        }
        finish2();
        if (kTimers) tacc[7] += clock64() - tpass;
    }
    if (kTimers) {
        tacc[11] = clock64() - tloop0;
        tacc[5] -= tacc[2];   // the exchange-barrier wait is nested in the finish timers
        if (lane == 0)
            for (int i = 0; i < 12; ++i) atomicAdd(&g_solve2_cycles[i], (unsigned long long)tacc[i]);
    }
#undef S2_T0
#undef S2_T1
}

extern "C" void hp_solve2_timers(unsigned long long* out, int reset) {
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out, g_solve2_cycles, sizeof(unsigned long long) * 16);
    if (reset) { unsigned long long z[16] = {0}; cudaMemcpyToSymbol(g_solve2_cycles, z, sizeof(z)); }
}

void launch_solve2(const Solve2Args& a_in, cudaStream_t st) {
    Solve2Args a = a_in;
    static int max_smem = 0, num_sm = 0;
    if (!max_smem) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
        cudaDeviceGetAttribute(&num_sm, cudaDevAttrMultiProcessorCount, dev);
    }
    a.stages = solve2_stages(a.nblk, (size_t)max_smem);
    if (a.stages == 0) return;   // excluded by solve2_stages (the engine never launches this)
    const size_t smem = solve2_smem_bytes(a.nblk, a.stages);
    static size_t attr_dev[kMaxDev] = {0};
    static int timers = -1, stage_cap = 0;
    if (timers < 0) {
        const char* ev = getenv("HP_S2_TIMERS");
        timers = (ev && ev[0] == '1') ? 1 : 0;
        const char* sc = getenv("HP_S2_STAGES");   // experiments: cap the ring depth
        stage_cap = sc ? atoi(sc) : 0;
    }
    if (stage_cap >= 2 && a.stages > stage_cap) a.stages = stage_cap;
    size_t& attr_smem = attr_dev[current_device_slot()];
    if (smem > attr_smem) {
        cudaFuncSetAttribute(k_solve2<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(k_solve2<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr_smem = smem;
    }
    const int total = a.ntiles * a.nsys;
    int grid = a.grid_limit > 0 ? a.grid_limit : num_sm;
    if (grid > total) grid = total;
    if (timers) k_solve2<true><<<grid, kS2Threads, smem, st>>>(a);
    else k_solve2<false><<<grid, kS2Threads, smem, st>>>(a);
}

// ---------------------------------------------------------------------------------------------------------------------
// k_rhs_tile: right-hand sides in the shared-memory layout of k_solve2's tile,
//     Rt[sys][tile][block row][plane][r][c ^ 4 (r & 3)] = (lam_row) Rfix[sys][16 tile + c][32 j + r] (+ wa)
// Built once per chain load when the chain runs unscaled right-hand sides (Philox mode), per iteration otherwise.
__global__ void __launch_bounds__(256) k_rhs_tile(double* __restrict__ Rt, const double* __restrict__ Rfix, const double* __restrict__ wa,
                                                  const double* __restrict__ lam, int nblk, int n, int N, int T, int Tp) {
    const int Np = nblk * 32;
    const int tile = blockIdx.x, sys = blockIdx.y, t0 = tile * kTT;
    const int ntiles = gridDim.x;
    const double* R = Rfix + 2 * ((size_t)sys * Tp + t0) * Np;
    const double* Wa = wa ? wa + 2 * ((size_t)sys * Tp + t0) * Np : nullptr;
    const double* lm = lam ? lam + (size_t)sys * Np : nullptr;
    double* out = Rt + ((size_t)sys * ntiles + tile) * nblk * kRowDoubles;
    // thread -> (column c, row): consecutive threads read consecutive rows of one time (coalesced 16-byte loads)
    for (int e = threadIdx.x; e < Np * kTT; e += blockDim.x) {
        const int c = e / Np, row = e - c * Np;
        double vr = 0.0, vi = 0.0;
        if (t0 + c < T && row < N) {
            const double2 x = *reinterpret_cast<const double2*>(R + 2 * ((size_t)c * Np + row));
            const double l = lm ? lm[row] : 1.0;
            vr = l * x.x; vi = l * x.y;
            if (Wa && row < n) { const double2 y = *reinterpret_cast<const double2*>(Wa + 2 * ((size_t)c * Np + row)); vr += y.x; vi += y.y; }
        }
        const int j = row >> 5, r = row & 31;
        double* o = out + (size_t)j * kRowDoubles + r * kTT + (c ^ ((r & 3) << 2));
        o[0] = vr;
        o[32 * kTT] = vi;
    }
}
void launch_rhs_tile(double* Rt, const double* Rfix, const double* wa, const double* lam, int nblk, int n, int N, int T, int Tp,
                     int ntiles, int nsys, cudaStream_t st) {
    k_rhs_tile<<<dim3(ntiles, nsys), 256, 0, st>>>(Rt, Rfix, wa, lam, nblk, n, N, T, Tp);
}

// ---------------------------------------------------------------------------------------------------------------------
// k_cg_scale: the scalar model of the reference's truncated CG (pspec.py:228; DESIGN.md section 1, hp_math.h: cg_theta).
// Every column of the exact solution is multiplied by theta(c, ||b||) with
//     c = sum_k lam_k^2 conj(r_k) x_k + sum_j conj(r_j) x_j,   ||b||^2 = sum_k lam_k^2 |r_k|^2 + sum_j |r_j|^2
// in the whitened variables (r = lam * Rfix + wa).  One CTA per (time, baseline); works on the global solution, so it
// serves every solve path (k_solve2, the dense-product solve of large N).
__global__ void __launch_bounds__(128) k_cg_scale(double* __restrict__ X, const double* __restrict__ Rfix, const double* __restrict__ wa,
                                                  const double* __restrict__ lam, int n, int N, int Np, int T, int Tp) {
    const int t = blockIdx.x, sys = blockIdx.y;
    if (t >= T) return;
    const size_t off = 2 * ((size_t)sys * Tp + t) * Np;
    double* x = X + off;
    const double* R = Rfix + off;
    const double* Wa = wa ? wa + off : nullptr;
    const double* lm = lam + (size_t)sys * Np;
    double sre = 0.0, sim = 0.0, sb = 0.0;
    for (int row = threadIdx.x; row < N; row += blockDim.x) {
        const double l = lm[row];
        double vr = l * R[2 * row], vi = l * R[2 * row + 1];
        if (Wa && row < n) { vr += Wa[2 * row]; vi += Wa[2 * row + 1]; }
        const double wgt = row < n ? l * l : 1.0;
        const double xr = x[2 * row], xim = x[2 * row + 1];
        sre += wgt * (vr * xr + vi * xim);   // conj(r) x
        sim += wgt * (vr * xim - vi * xr);
        sb += wgt * (vr * vr + vi * vi);
    }
    __shared__ double red[3][4];
    __shared__ double th[2];
    for (int o = 16; o > 0; o >>= 1) {
        sre += __shfl_down_sync(0xffffffffu, sre, o);
        sim += __shfl_down_sync(0xffffffffu, sim, o);
        sb += __shfl_down_sync(0xffffffffu, sb, o);
    }
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = sre; red[1][threadIdx.x >> 5] = sim; red[2][threadIdx.x >> 5] = sb; }
    __syncthreads();
    if (threadIdx.x == 0) {
        cplx c; c.re = red[0][0] + red[0][1] + red[0][2] + red[0][3]; c.im = red[1][0] + red[1][1] + red[1][2] + red[1][3];
        const double b2 = red[2][0] + red[2][1] + red[2][2] + red[2][3];
        const cplx v = cg_theta(c, sqrt(b2), 1e-8, 1e-6, 100000);
        th[0] = v.re; th[1] = v.im;
    }
    __syncthreads();
    const double tr = th[0], ti = th[1];
    for (int row = threadIdx.x; row < Np; row += blockDim.x) {
        const double xr = x[2 * row], xim = x[2 * row + 1];
        x[2 * row] = tr * xr - ti * xim;
        x[2 * row + 1] = tr * xim + ti * xr;
    }
}
void launch_cg_scale(double* X, const double* Rfix, const double* wa, const double* lam, int n, int N, int Np, int T, int Tp,
                     int nsys, cudaStream_t st) {
    k_cg_scale<<<dim3(T, nsys), 128, 0, st>>>(X, Rfix, wa, lam, n, N, Np, T, Tp);
}

}  // namespace hp
