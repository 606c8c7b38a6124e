// LAUUM with the gradient reduction in its epilogue (K3 + K1g fused).
//
// Reference arithmetic replaced: the per-hyper-parameter solve / trace expressions of
// _emulatoroptimise.py:345-372 and :450-487, which collapse to
//     grad_k = 1/2 sum_ij T^k_ij W_ij,   W = A^-1 - U U^T,   U = [A^-1 H K^-T | sqrt(f) A^-1 (y - H beta)],
// with T^k built from E_ij = exp(-sum_k ((x_ik - x_jk)/delta_k)^2)  (_emulatorkernels.py:53-71, :126-144).
//
// Round 1 formed A^-1 = L^-T L^-1 with the generic DMMA GEMM, stored its lower tiles (8 n^2 / 2 bytes per item) and
// read them back in a separate reduction kernel that also recomputed E and the U.U^T dot products on the FP64 ALU
// (3.15 ms per 32-item step, 0.32 of the FP64 roof, long-scoreboard bound on the A^-1 loads).  Here a 128x128 tile of
// W never leaves the registers of the CTA that computed it:
//   * the k loop runs over L^-1 (k >= i, as LAUUM) and then over the NR columns of U, whose tiles come from a
//     pre-negated k-major copy (-U^T as the A operand, U^T as the B operand), so W = A^-1 - U U^T falls out of the
//     tensor pipe and the 2 x 18 FMA per entry of the old dot products disappear;
//   * the epilogue stages the scaled input tiles X_i / delta, X_j / delta in the (now free) pipeline stages, reads
//     E_ij = exp(-D_ij) for the 64 entries a lane holds from the copy the covariance build kept (the FP64 pipe is the
//     bottleneck of this path and HBM is idle: 8 bytes per entry are cheaper than 2d + 22 FP64 instructions), and
//     reduces sum W E Delta_k^2 (k < d), sum W E, sum_diag W, sum_diag W r into one partial row per tile, which
//     grad_finalize_kernel adds up in a fixed order.
// A^-1 is not written at all on this path (the sensitivity code builds it on demand with the generic kernel).
#include "gpe_gemm.cuh"
#include "gpe_kernels.cuh"

#include <algorithm>
#include <cstdlib>

namespace gpe {

struct LauumGradP {
    const double* Li;      // [B][np][np]  L^-1 (lower; upper blocks zero)
    long long sL;
    int np, n, d;
    int ku;                // k-tiles of the U part (ceil(nu / 16))
    const double* Ut;      // [B][NR][np]   U^T
    const double* nUt;     // [B][NR][np]  -U^T
    long long sU;
    const double* X;       // [n][d]
    const double* r;       // [n] or null
    const double* winv;    // [B][d]  1 / delta
    const double* E;       // [B][np][np]  exp(-D) as written by the covariance build (lower 64x64 tiles)
    long long sE;
    double* part;          // [B][ntile][d + 3]
    int ntile;
    double* Wout;          // non-null: store the W tile (strides sL / np) and leave the reduction to another kernel
    int order;             // tile order: 0 lower-triangle rows, 1 the generic kernel's column groups (GPE_LG_ORDER)
};

constexpr int LG_LDX = 128 + 4;     // row stride of the k-major X tiles in the epilogue

__device__ __forceinline__ void tri_decode_lg(int t, int& ti, int& tj) {
    int r = (int)((sqrtf(8.0f * (float)t + 1.0f) - 1.0f) * 0.5f);
    while ((r + 1) * (r + 2) / 2 <= t) r++;
    while (r * (r + 1) / 2 > t) r--;
    ti = r;
    tj = t - r * (r + 1) / 2;
}

__global__ void __launch_bounds__(WS_THREADS, 1) lauum_grad_kernel(LauumGradP p) {
    constexpr int BM = 128, BN = 128, WMW = 4, WNW = WS_CONSUMERS / WMW;
    constexpr int WTM = BM / WMW, WTN = BN / WNW, FM = WTM / 8, FN = WTN / 8;       // warp tile 32 x 64: FM = 4, FN = 8
    constexpr int A_EL = TileShape<BM, false>::ELEMS, B_EL = TileShape<BN, false>::ELEMS;
    constexpr int A_LD = TileShape<BM, false>::LD, B_LD = TileShape<BN, false>::LD;
    extern __shared__ __align__(16) double smem[];
    __shared__ __align__(8) uint64_t full_bar[GEMM_STAGES], empty_bar[GEMM_STAGES];

    // lower-triangle tiles, row by row: tile t = ti (ti + 1) / 2 + tj.  Row ti has the k range [ti * 128, np):
    // the launch starts on its longest tiles and ends on its shortest
    int ti, tj;
    if (p.order == 0) {
        tri_decode_lg((int)blockIdx.x, ti, tj);
    } else {
        // the generic kernel's order for row-triangular launches: column blocks in groups of 16, rows ascending inside a group
        const int T = p.np / BM, id = (int)blockIdx.x;
        const int full = (T / WS_COLGROUP) * WS_COLGROUP * T;
        int g0, gsz, r;
        if (id < full) { g0 = (id / (WS_COLGROUP * T)) * WS_COLGROUP; gsz = WS_COLGROUP; r = id % (WS_COLGROUP * T); }
        else { g0 = (T / WS_COLGROUP) * WS_COLGROUP; gsz = T - g0; r = id - full; }
        ti = r / gsz;
        tj = g0 + (r - ti * gsz);
        if (ti < tj) return;
    }
    const int tile_id = ti * (ti + 1) / 2 + tj;
    const int b = blockIdx.z;
    const int m0 = ti * BM, n0 = tj * BN;
    const int kbeg = m0;
    const int KT0 = (p.np - kbeg) / GEMM_BK, KT = KT0 + p.ku;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < GEMM_STAGES; s++) {
            mbar_init(&full_bar[s], 32);
            mbar_init(&empty_bar[s], WS_CONSUMERS);
        }
    }
    __syncthreads();

    double* As = smem;
    double* Bs = smem + GEMM_STAGES * A_EL;

    if (warp == WS_CONSUMERS) {
        // ===== producer warp: L^-1 tiles (both operands: rows k, columns m0.. / n0..), then the U tiles =====
        const double* Lb = p.Li + (size_t)b * p.sL;
        const double* Ag = Lb + (size_t)kbeg * p.np + m0;
        const double* Bg = Lb + (size_t)kbeg * p.np + n0;
        const double* Au = p.nUt + (size_t)b * p.sU + m0;
        const double* Bu = p.Ut + (size_t)b * p.sU + n0;
        const size_t kstep = (size_t)GEMM_BK * p.np;
        for (int kt = 0; kt < KT; kt++) {
            const int s = kt % GEMM_STAGES;
            if (kt >= GEMM_STAGES) mbar_wait(&empty_bar[s], ((kt / GEMM_STAGES) - 1) & 1);
            if (kt < KT0) {
                load_tile_warp<BM, false>(As + s * A_EL, Ag + kt * kstep, p.np, lane);
                load_tile_warp<BN, false>(Bs + s * B_EL, Bg + kt * kstep, p.np, lane);
            } else {
                load_tile_warp<BM, false>(As + s * A_EL, Au + (kt - KT0) * kstep, p.np, lane);
                load_tile_warp<BN, false>(Bs + s * B_EL, Bu + (kt - KT0) * kstep, p.np, lane);
            }
            cp_async_mbar_arrive_noinc(&full_bar[s]);
            // the epilogue reads this tile's 128 x 128 block of E (written by the covariance build a whole factorisation
            // ago: it comes from HBM).  Ask L2 for it while the consumers still have a few k-tiles to go.
            if (p.Wout == nullptr && (kt == KT - GEMM_STAGES - 4 || (KT < GEMM_STAGES + 5 && kt == 0))) {
                const double* Eg = p.E + (size_t)b * p.sE + (size_t)m0 * p.np + n0;
#pragma unroll 4
                for (int q = 0; q < 32; q++) {                 // 128 rows x 8 lines of 128 bytes
                    const int li = q * 32 + lane;
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(Eg + (size_t)(li >> 3) * p.np + (li & 7) * 16));
                }
            }
        }
        cp_async_wait<0>();
        return;
    }

    // ===== consumer warps: 4 x 2 grid, warps 0..3 rows 0..3 of column 0, warps 4..7 rows 3..0 of column 1 =====
    const int wrow = (warp < 4 ? warp : 7 - warp), wcol = warp / 4;
    const int wm0 = wrow * WTM, wn0 = wcol * WTN;
    const int fr = lane >> 2, fc = lane & 3;
    // inside the diagonal block of L^-1 (k in [m0, m0 + 128)) rows k < m see only zeros of the A operand
    const int wk_lo = wm0 / GEMM_BK;
    double acc[FM][FN][2];
#pragma unroll
    for (int i = 0; i < FM; i++)
#pragma unroll
        for (int j = 0; j < FN; j++) acc[i][j][0] = acc[i][j][1] = 0.0;
    const int a_off = fc * A_LD + wm0 + fr;
    const int b_off = fc * B_LD + wn0 + fr;
    constexpr int a_kk = 4 * A_LD, b_kk = 4 * B_LD;

    // Diagonal tiles (ti == tj): only entries on or below the diagonal are used.  With the 4 x 2 warp grid the warps of
    // rows 0, 1 in column 1 have nothing to do and the warps (row 0, column 0), (row 2, column 1) need only their left four
    // fragment columns: per SM sub-partition that is 48 / 48 / 32 / 32 DMMAs per k4 step instead of 64 -- the tile costs
    // 3/4 (diagonal tiles are 8.8 % of a LAUUM launch at n = 4096).  Warp-uniform, decided outside the k loop.
    int jn = FN;
    if (ti == tj) jn = (wcol == 0) ? (wrow == 0 ? FN / 2 : FN) : (wrow == 3 ? FN : (wrow == 2 ? FN / 2 : 0));
    for (int kt = 0; kt < KT; kt++) {
        const int s = kt % GEMM_STAGES;
        mbar_wait(&full_bar[s], (kt / GEMM_STAGES) & 1);
        const double* at = As + s * A_EL + a_off;
        const double* bt = Bs + s * B_EL + b_off;
        if (kt >= wk_lo) {
            if (jn == FN) {
#pragma unroll
                for (int kk = 0; kk < GEMM_BK / 4; kk++) {
                    double af[FM], bf[FN];
#pragma unroll
                    for (int i = 0; i < FM; i++) af[i] = at[kk * a_kk + i * 8];
#pragma unroll
                    for (int j = 0; j < FN; j++) bf[j] = bt[kk * b_kk + j * 8];
#pragma unroll
                    for (int i = 0; i < FM; i++)
#pragma unroll
                        for (int j = 0; j < FN; j++) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
                }
            } else if (jn == FN / 2) {
#pragma unroll
                for (int kk = 0; kk < GEMM_BK / 4; kk++) {
                    double af[FM], bf[FN / 2];
#pragma unroll
                    for (int i = 0; i < FM; i++) af[i] = at[kk * a_kk + i * 8];
#pragma unroll
                    for (int j = 0; j < FN / 2; j++) bf[j] = bt[kk * b_kk + j * 8];
#pragma unroll
                    for (int i = 0; i < FM; i++)
#pragma unroll
                        for (int j = 0; j < FN / 2; j++) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[s]);
    }

    if (p.Wout != nullptr) {                         // split route: W = A^-1 - U U^T to global memory, 16-byte stores
        double* Cg = p.Wout + (size_t)b * p.sL + (size_t)(m0 + wm0 + fr) * p.np + n0 + wn0 + 2 * fc;
#pragma unroll
        for (int i = 0; i < FM; i++)
#pragma unroll
            for (int j = 0; j < FN; j++)
                *reinterpret_cast<double2*>(Cg + (size_t)(8 * i) * p.np + 8 * j) = make_double2(acc[i][j][0], acc[i][j][1]);
        return;
    }
    // ===== epilogue: acc = W tile.  The pipeline stages are free: stage the E tile and the scaled input tiles there. =====
    const int d = p.d, n = p.n;
    named_bar_sync(2, WS_CONSUMERS * 32);           // every consumer is past its last stage read
    double* Es = smem;                              // [128][LG_LDX]  exp(-D) of this tile (from the covariance build's copy);
                                                    //                later the per-lane partial sums [d][256]
    double* Xi = Es + 128 * LG_LDX;                 // [d][LG_LDX]  rows m0 .. m0 + 127, k-major, scaled by 1 / delta
    double* Xj = Xi + (size_t)d * LG_LDX;           // [d][LG_LDX]  rows n0 ..
    double* red = Xj + (size_t)d * LG_LDX;          // [8][3]
    {
        // thread -> 16-byte chunk (tid & 63) of rows (tid >> 6) + 4 it
        const double* src = p.E + (size_t)b * p.sE + (size_t)(m0 + (tid >> 6)) * p.np + n0 + 2 * (tid & 63);
        double* dst = Es + (tid >> 6) * LG_LDX + 2 * (tid & 63);
        const size_t sstep = (size_t)4 * p.np;
#pragma unroll 4
        for (int it = 0; it < 32; it++) {
            cp_async16(dst, src);
            src += sstep;
            dst += 4 * LG_LDX;
        }
        cp_async_commit();
        const double* __restrict__ w = p.winv + (size_t)b * d;
        const double* __restrict__ Xg = p.X;
        const int row = tid >> 1, gi = m0 + row, gj = n0 + row;
        for (int k = tid & 1; k < d; k += 2) {
            const double wk = __ldg(w + k);
            const double vi = (gi < n) ? __ldg(Xg + (size_t)gi * d + k) : 0.0;
            const double vj = (gj < n) ? __ldg(Xg + (size_t)gj * d + k) : 0.0;
            Xi[k * LG_LDX + row] = vi * wk;
            Xj[k * LG_LDX + row] = vj * wk;
        }
        cp_async_wait<0>();
    }
    named_bar_sync(2, WS_CONSUMERS * 32);

    const int row0 = wm0 + fr, col0 = wn0 + 2 * fc;          // lane's entries: rows row0 + 8 i, columns col0 + 8 j + {0, 1}
    double sE = 0.0, sD = 0.0, sDr = 0.0;
    // pass 1: acc <- t = 2 W E for the pairs below the diagonal inside the n x n matrix, 0 elsewhere; the diagonal sums
    // on the way.  Tiles strictly below the diagonal and inside the matrix -- all but a few -- need no case analysis.
    if (ti > tj && m0 + BM <= n) {
#pragma unroll
        for (int j = 0; j < FN; j++)
#pragma unroll
            for (int i = 0; i < FM; i++) {
                const double2 ev = *reinterpret_cast<const double2*>(Es + (row0 + 8 * i) * LG_LDX + col0 + 8 * j);
                const double t0 = 2.0 * acc[i][j][0] * ev.x, t1 = 2.0 * acc[i][j][1] * ev.y;
                acc[i][j][0] = t0;
                acc[i][j][1] = t1;
                sE += t0 + t1;
            }
    } else {
        const bool diag_tile = (ti == tj);
#pragma unroll
        for (int j = 0; j < FN; j++)
#pragma unroll
            for (int i = 0; i < FM; i++) {
                const int gi = m0 + row0 + 8 * i;
                const double2 ev = *reinterpret_cast<const double2*>(Es + (row0 + 8 * i) * LG_LDX + col0 + 8 * j);
#pragma unroll
                for (int e = 0; e < 2; e++) {
                    const int gj = n0 + col0 + 8 * j + e;
                    const double wv = acc[i][j][e];
                    double tv = 2.0 * wv * (e ? ev.y : ev.x);
                    bool below = gi < n;                        // gj < gi < n off the diagonal tiles
                    if (diag_tile) {
                        below = below && gj < gi;               // each pair once, from below; selected, not multiplied by 0:
                        if (gi == gj && gi < n) {               // the copy of E holds the lower 64x64 tiles only
                            sD += wv;
                            if (p.r != nullptr) sDr = fma(wv, p.r[gi], sDr);
                        }
                    }
                    tv = below ? tv : 0.0;
                    acc[i][j][e] = tv;
                    sE += tv;
                }
            }
    }
    sE = warp_sum(sE); sD = warp_sum(sD); sDr = warp_sum(sDr);
    if (lane == 0) { red[warp * 3] = sE; red[warp * 3 + 1] = sD; red[warp * 3 + 2] = sDr; }
    // pass 2: per dimension, sum t * Delta_k^2 over the lane's 64 entries.  With all 64 t's in registers (128 of the 168
    // a thread of a 9-warp CTA can have) the compiler has no room to overlap the sub -> mul -> fma triples and every
    // instruction waits for the one before (ncu: stall_wait on each of them, 44 % of the FP64 issue rate).  So the
    // right half of the warp tile (fragment columns 4..7) is parked in shared memory -- each lane overwrites the E
    // entries it has just consumed, no other lane touches them -- and the two halves are reduced one after the other
    // with eight triples in flight.  The lane's partial sums go to shared memory; cross-lane sums happen once at the end.
#pragma unroll
    for (int j = FN / 2; j < FN; j++)
#pragma unroll
        for (int i = 0; i < FM; i++)
            *reinterpret_cast<double2*>(Es + (row0 + 8 * i) * LG_LDX + col0 + 8 * j) = make_double2(acc[i][j][0], acc[i][j][1]);
    double* gpart = Xj + (size_t)d * LG_LDX + 8 * 3;      // [d][256], behind the reduction rows
    auto half_pass = [&](const double (&t)[FM][FN / 2][2], int jbase, bool add) {
        for (int k = 0; k < d; k++) {
            double xi[FM];
#pragma unroll
            for (int i = 0; i < FM; i++) xi[i] = Xi[k * LG_LDX + row0 + 8 * i];
            double g[FM][2];
#pragma unroll
            for (int i = 0; i < FM; i++) g[i][0] = g[i][1] = 0.0;
#pragma unroll
            for (int j = 0; j < FN / 2; j++) {
                const double2 xj = *reinterpret_cast<const double2*>(Xj + k * LG_LDX + col0 + 8 * (jbase + j));
                double d0[FM], d1[FM], m0[FM], m1[FM];
#pragma unroll
                for (int i = 0; i < FM; i++) { d0[i] = xi[i] - xj.x; d1[i] = xi[i] - xj.y; }
#pragma unroll
                for (int i = 0; i < FM; i++) { m0[i] = t[i][j][0] * d0[i]; m1[i] = t[i][j][1] * d1[i]; }
#pragma unroll
                for (int i = 0; i < FM; i++) { g[i][0] = fma(m0[i], d0[i], g[i][0]); g[i][1] = fma(m1[i], d1[i], g[i][1]); }
            }
            const double gs = ((g[0][0] + g[0][1]) + (g[1][0] + g[1][1])) + ((g[2][0] + g[2][1]) + (g[3][0] + g[3][1]));
            if (add) gpart[k * 256 + tid] += gs;
            else gpart[k * 256 + tid] = gs;
        }
    };
    {
        double t[FM][FN / 2][2];
#pragma unroll
        for (int i = 0; i < FM; i++)
#pragma unroll
            for (int j = 0; j < FN / 2; j++) { t[i][j][0] = acc[i][j][0]; t[i][j][1] = acc[i][j][1]; }
        half_pass(t, 0, false);
#pragma unroll
        for (int i = 0; i < FM; i++)
#pragma unroll
            for (int j = 0; j < FN / 2; j++) {
                const double2 v = *reinterpret_cast<const double2*>(Es + (row0 + 8 * i) * LG_LDX + col0 + 8 * (j + FN / 2));
                t[i][j][0] = v.x;
                t[i][j][1] = v.y;
            }
        half_pass(t, FN / 2, true);
    }
    named_bar_sync(2, WS_CONSUMERS * 32);
    const int nv = d + 3;
    double* pout = p.part + ((size_t)b * p.ntile + tile_id) * nv;
    // dimension k: 16 threads add 16 lane partials each, then a 16-lane butterfly -- a fixed order
    for (int k0 = 0; k0 < d; k0 += 16) {
        const int k = k0 + (tid >> 4), part = tid & 15;
        double sacc = 0.0;
        if (k < d) {
#pragma unroll
            for (int q = 0; q < 16; q++) sacc += gpart[k * 256 + part * 16 + q];
        }
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) sacc += __shfl_xor_sync(0xffffffffu, sacc, o);
        if (k < d && part == 0) pout[k] = sacc;
    }
    if (tid < 3) {
        double sacc = 0.0;
#pragma unroll
        for (int w = 0; w < WS_CONSUMERS; w++) sacc += red[w * 3 + tid];
        pout[d + tid] = sacc;
    }
}


// ------------------------------------------------------------------------------------------------------------------
// Two CTAs per SM: the same kernel on 128x64 tiles with 4 math warps (4x1 grid, warp tile 32x64 as above) + 1 producer warp
// and 3 stages.  The epilogue of one CTA -- barrier, E-tile and X-tile loads, the scalar FP64 passes -- then overlaps the
// other CTA's DMMA main loop instead of leaving the tensor pipe idle (one-CTA kernel: 0.9 ms of waits per 32-item launch), and
// the idle warps of diagonal tiles give their pipe slots to the neighbour.  Shared memory: max(3 stages = 77 KB, epilogue =
// E tile 70 KB + X tiles (d = 16: 26 KB)) per CTA.
constexpr int LG2_CONS = 4;
constexpr int LG2_THREADS = (LG2_CONS + 1) * 32;
constexpr int LG2_STAGES = 3;
constexpr int LG2_LDE = 64 + 4;        // row stride of the 128 x 64 E tile and of the k-major X_j tile

__global__ void __launch_bounds__(LG2_THREADS, 2) lauum_grad64_kernel(LauumGradP p) {
    constexpr int BM = 128, BN = 64, FM = 4, FN = 8;               // per warp: 32 rows x 64 columns
    constexpr int A_EL = TileShape<BM, false>::ELEMS, B_EL = TileShape<BN, false>::ELEMS;
    constexpr int A_LD = TileShape<BM, false>::LD, B_LD = TileShape<BN, false>::LD;
    extern __shared__ __align__(16) double smem[];
    __shared__ __align__(8) uint64_t full_bar[LG2_STAGES], empty_bar[LG2_STAGES];

    // lower-triangle tiles in 64-column units, row by row: row ti holds tj = 0 .. 2 ti + 1; tile id t = ti (ti + 1) + tj
    int ti, tj;
    {
        const int t = (int)blockIdx.x;
        int r = (int)((sqrtf(4.0f * (float)t + 1.0f) - 1.0f) * 0.5f);
        while ((r + 1) * (r + 2) <= t) r++;
        while (r * (r + 1) > t) r--;
        ti = r;
        tj = t - r * (r + 1);
    }
    const int b = blockIdx.z;
    const int m0 = ti * BM, n0 = tj * BN;
    const int kbeg = m0;
    const int KT0 = (p.np - kbeg) / GEMM_BK, KT = KT0 + p.ku;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < LG2_STAGES; s++) {
            mbar_init(&full_bar[s], 32);
            mbar_init(&empty_bar[s], LG2_CONS);
        }
    }
    __syncthreads();

    double* As = smem;
    double* Bs = smem + LG2_STAGES * A_EL;

    if (warp == LG2_CONS) {
        const double* Lb = p.Li + (size_t)b * p.sL;
        const double* Ag = Lb + (size_t)kbeg * p.np + m0;
        const double* Bg = Lb + (size_t)kbeg * p.np + n0;
        const double* Au = p.nUt + (size_t)b * p.sU + m0;
        const double* Bu = p.Ut + (size_t)b * p.sU + n0;
        const size_t kstep = (size_t)GEMM_BK * p.np;
        for (int kt = 0; kt < KT; kt++) {
            const int s = kt % LG2_STAGES;
            if (kt >= LG2_STAGES) mbar_wait(&empty_bar[s], ((kt / LG2_STAGES) - 1) & 1);
            if (kt < KT0) {
                load_tile_warp<BM, false>(As + s * A_EL, Ag + kt * kstep, p.np, lane);
                load_tile_warp<BN, false>(Bs + s * B_EL, Bg + kt * kstep, p.np, lane);
            } else {
                load_tile_warp<BM, false>(As + s * A_EL, Au + (kt - KT0) * kstep, p.np, lane);
                load_tile_warp<BN, false>(Bs + s * B_EL, Bu + (kt - KT0) * kstep, p.np, lane);
            }
            cp_async_mbar_arrive_noinc(&full_bar[s]);
            if (p.Wout == nullptr && (kt == KT - LG2_STAGES - 4 || (KT < LG2_STAGES + 5 && kt == 0))) {
                const double* Eg = p.E + (size_t)b * p.sE + (size_t)m0 * p.np + n0;
#pragma unroll 4
                for (int q = 0; q < 16; q++) {                 // 128 rows x 4 lines of 128 bytes
                    const int li = q * 32 + lane;
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(Eg + (size_t)(li >> 2) * p.np + (li & 3) * 16));
                }
            }
        }
        cp_async_wait<0>();
        return;
    }

    const int wm0 = warp * 32;
    const int fr = lane >> 2, fc = lane & 3;
    const int wk_lo = wm0 / GEMM_BK;
    double acc[FM][FN][2];
#pragma unroll
    for (int i = 0; i < FM; i++)
#pragma unroll
        for (int j = 0; j < FN; j++) acc[i][j][0] = acc[i][j][1] = 0.0;
    const int a_off = fc * A_LD + wm0 + fr;
    const int b_off = fc * B_LD + fr;
    constexpr int a_kk = 4 * A_LD, b_kk = 4 * B_LD;
    // fragment columns that touch the lower triangle: all of them below the diagonal band; on it, a warp whose rows end
    // before the tile's columns begin has nothing to do, one whose rows end inside the left half needs that half only
    int jn = FN;
    {
        const int row_end = m0 + wm0 + 32;            // one past the warp's last row
        if (row_end <= n0) jn = 0;
        else if (row_end <= n0 + 32) jn = FN / 2;
    }
    for (int kt = 0; kt < KT; kt++) {
        const int s = kt % LG2_STAGES;
        mbar_wait(&full_bar[s], (kt / LG2_STAGES) & 1);
        const double* at = As + s * A_EL + a_off;
        const double* bt = Bs + s * B_EL + b_off;
        if (kt >= wk_lo) {
            if (jn == FN) {
#pragma unroll
                for (int kk = 0; kk < GEMM_BK / 4; kk++) {
                    double af[FM], bf[FN];
#pragma unroll
                    for (int i = 0; i < FM; i++) af[i] = at[kk * a_kk + i * 8];
#pragma unroll
                    for (int j = 0; j < FN; j++) bf[j] = bt[kk * b_kk + j * 8];
#pragma unroll
                    for (int i = 0; i < FM; i++)
#pragma unroll
                        for (int j = 0; j < FN; j++) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
                }
            } else if (jn == FN / 2) {
#pragma unroll
                for (int kk = 0; kk < GEMM_BK / 4; kk++) {
                    double af[FM], bf[FN / 2];
#pragma unroll
                    for (int i = 0; i < FM; i++) af[i] = at[kk * a_kk + i * 8];
#pragma unroll
                    for (int j = 0; j < FN / 2; j++) bf[j] = bt[kk * b_kk + j * 8];
#pragma unroll
                    for (int i = 0; i < FM; i++)
#pragma unroll
                        for (int j = 0; j < FN / 2; j++) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[s]);
    }

    if (p.Wout != nullptr) {
        double* Cg = p.Wout + (size_t)b * p.sL + (size_t)(m0 + wm0 + fr) * p.np + n0 + 2 * fc;
#pragma unroll
        for (int i = 0; i < FM; i++)
#pragma unroll
            for (int j = 0; j < FN; j++)
                *reinterpret_cast<double2*>(Cg + (size_t)(8 * i) * p.np + 8 * j) = make_double2(acc[i][j][0], acc[i][j][1]);
        return;
    }
    // ===== epilogue (as in the one-CTA kernel, 128 threads) =====
    const int d = p.d, n = p.n;
    named_bar_sync(2, LG2_CONS * 32);
    double* Es = smem;                              // [128][LG2_LDE]; later the per-lane partial sums [d][128]
    double* Xi = Es + (128 * LG2_LDE > 128 * d ? 128 * LG2_LDE : 128 * d);   // [d][LG_LDX]
    double* Xj = Xi + (size_t)d * LG_LDX;           // [d][LG2_LDE]
    double* red = Xj + (size_t)d * LG2_LDE;         // [4][3]
    {
        const double* src = p.E + (size_t)b * p.sE + (size_t)(m0 + (tid >> 5)) * p.np + n0 + 2 * (tid & 31);
        double* dst = Es + (tid >> 5) * LG2_LDE + 2 * (tid & 31);
        const size_t sstep = (size_t)4 * p.np;
#pragma unroll 4
        for (int it = 0; it < 32; it++) {
            cp_async16(dst, src);
            src += sstep;
            dst += 4 * LG2_LDE;
        }
        cp_async_commit();
        const double* __restrict__ w = p.winv + (size_t)b * d;
        const double* __restrict__ Xg = p.X;
        const int gi = m0 + tid, gj = n0 + tid;
        for (int k = 0; k < d; k++) {
            const double wk = __ldg(w + k);
            Xi[k * LG_LDX + tid] = (gi < n) ? __ldg(Xg + (size_t)gi * d + k) * wk : 0.0;
            if (tid < BN) Xj[k * LG2_LDE + tid] = (gj < n) ? __ldg(Xg + (size_t)gj * d + k) * wk : 0.0;
        }
        cp_async_wait<0>();
    }
    named_bar_sync(2, LG2_CONS * 32);

    const int row0 = wm0 + fr, col0 = 2 * fc;
    double sE = 0.0, sD = 0.0, sDr = 0.0;
    if (n0 + BN <= m0 && m0 + BM <= n) {            // strictly below the diagonal band and inside the matrix
#pragma unroll
        for (int j = 0; j < FN; j++)
#pragma unroll
            for (int i = 0; i < FM; i++) {
                const double2 ev = *reinterpret_cast<const double2*>(Es + (row0 + 8 * i) * LG2_LDE + col0 + 8 * j);
                const double t0 = 2.0 * acc[i][j][0] * ev.x, t1 = 2.0 * acc[i][j][1] * ev.y;
                acc[i][j][0] = t0;
                acc[i][j][1] = t1;
                sE += t0 + t1;
            }
    } else {
#pragma unroll
        for (int j = 0; j < FN; j++)
#pragma unroll
            for (int i = 0; i < FM; i++) {
                const int gi = m0 + row0 + 8 * i;
                const double2 ev = *reinterpret_cast<const double2*>(Es + (row0 + 8 * i) * LG2_LDE + col0 + 8 * j);
#pragma unroll
                for (int e = 0; e < 2; e++) {
                    const int gj = n0 + col0 + 8 * j + e;
                    const double wv = acc[i][j][e];
                    double tv = 2.0 * wv * (e ? ev.y : ev.x);
                    const bool below = gi < n && gj < gi;       // each pair once, from below; selected, not multiplied by 0
                    if (gi == gj && gi < n) {
                        sD += wv;
                        if (p.r != nullptr) sDr = fma(wv, p.r[gi], sDr);
                    }
                    tv = below ? tv : 0.0;
                    acc[i][j][e] = tv;
                    sE += tv;
                }
            }
    }
    sE = warp_sum(sE); sD = warp_sum(sD); sDr = warp_sum(sDr);
    if (lane == 0) { red[warp * 3] = sE; red[warp * 3 + 1] = sD; red[warp * 3 + 2] = sDr; }
    named_bar_sync(2, LG2_CONS * 32);               // the E tile is consumed: its memory now takes the partial sums
    double* gpart = Es;                             // [d][128]
    for (int k = 0; k < d; k++) {
        double xi[FM];
#pragma unroll
        for (int i = 0; i < FM; i++) xi[i] = Xi[k * LG_LDX + row0 + 8 * i];
        double g[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
#pragma unroll
        for (int j = 0; j < FN; j++) {
            const double2 xj = *reinterpret_cast<const double2*>(Xj + k * LG2_LDE + col0 + 8 * j);
#pragma unroll
            for (int i = 0; i < FM; i++) {
                const double d0 = xi[i] - xj.x, d1 = xi[i] - xj.y;
                g[i & 1][0] = fma(acc[i][j][0] * d0, d0, g[i & 1][0]);
                g[i & 1][1] = fma(acc[i][j][1] * d1, d1, g[i & 1][1]);
            }
        }
        gpart[k * 128 + tid] = (g[0][0] + g[0][1]) + (g[1][0] + g[1][1]);
    }
    named_bar_sync(2, LG2_CONS * 32);
    const int nv = d + 3;
    double* pout = p.part + ((size_t)b * p.ntile + blockIdx.x) * nv;
    for (int k0 = 0; k0 < d; k0 += 16) {            // dimension k: 8 threads add 16 lane partials each, then an 8-lane butterfly
        const int k = k0 + (tid >> 3), part = tid & 7;
        double sacc = 0.0;
        if (k < d) {
#pragma unroll
            for (int q = 0; q < 16; q++) sacc += gpart[k * 128 + part * 16 + q];
        }
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) sacc += __shfl_xor_sync(0xffffffffu, sacc, o);
        if (k < d && part == 0) pout[k] = sacc;
    }
    if (tid < 3) {
        double sacc = 0.0;
#pragma unroll
        for (int w = 0; w < LG2_CONS; w++) sacc += red[w * 3 + tid];
        pout[d + tid] = sacc;
    }
}

static size_t lauum_grad64_smem(int d) {
    const size_t stages = (size_t)LG2_STAGES * (TileShape<128, false>::ELEMS + TileShape<64, false>::ELEMS) * sizeof(double);
    const size_t epi = ((size_t)std::max(128 * LG2_LDE, 128 * d) + (size_t)d * LG_LDX + (size_t)d * LG2_LDE + 4 * 3) * sizeof(double);
    return std::max(stages, epi);
}

// -U^T and U^T, k-major: Ut[b][c][i] = U[b][i][c]
__global__ void __launch_bounds__(256) ut_kernel(const double* __restrict__ U, int np, double* __restrict__ Ut, double* __restrict__ nUt) {
    __shared__ double T[32][NR + 1];
    const int b = blockIdx.y, i0 = blockIdx.x * 32;
    const double* Ub = U + ((size_t)b * np + i0) * NR;
    for (int e = threadIdx.x; e < 32 * NR; e += 256) T[e / NR][e % NR] = Ub[e];
    __syncthreads();
    for (int e = threadIdx.x; e < 32 * NR; e += 256) {
        const int c = e / 32, i = e % 32;
        const double v = T[i][c];
        const size_t o = ((size_t)b * NR + c) * np + i0 + i;
        Ut[o] = v;
        nUt[o] = -v;
    }
}

static size_t lauum_grad_smem(int d) {
    // the epilogue's E tile, two k-major input tiles and the reduction rows reuse the pipeline stages
    const size_t epi = ((size_t)128 * LG_LDX + (size_t)2 * d * LG_LDX + 8 * 3 + (size_t)256 * d) * sizeof(double);
    return std::max(epi, gemm_smem_bytes<128, 128, false, false>());
}

bool lauum_grad_supported(int d) { return lauum_grad_smem(d) <= 227 * 1024; }

cudaError_t launch_lauum_grad(const double* Li, long long sL, int np, int n, int d, int nu, const double* U, double* Ut, double* nUt,
                              const double* X, const double* r, const double* winv, const double* E, long long sE, double* part,
                              int B, cudaStream_t st, double* Wout) {
    if (np % 128 || (Wout == nullptr && !lauum_grad_supported(d))) return cudaErrorInvalidValue;
    ut_kernel<<<dim3(np / 32, B), 256, 0, st>>>(U, np, Ut, nUt);
    LauumGradP p;
    p.Li = Li; p.sL = sL; p.np = np; p.n = n; p.d = d;
    p.ku = (nu + GEMM_BK - 1) / GEMM_BK;
    p.Ut = Ut; p.nUt = nUt; p.sU = (long long)NR * np;
    p.X = X; p.r = r; p.winv = winv; p.E = E; p.sE = sE; p.part = part; p.Wout = Wout;
    const int T = np / 128;
    p.ntile = T * (T + 1) / 2;
    static int order = -1, wide64 = -1;
    if (order < 0) {
        const char* e = getenv("GPE_LG_ORDER");
        order = e ? atoi(e) : 0;
        e = getenv("GPE_LG64");                       // two-CTAs-per-SM 128x64 variant (A/B knob)
        wide64 = e ? atoi(e) : 0;
    }
    p.order = order;
    if (wide64 && 2 * lauum_grad64_smem(d) <= 220 * 1024) {
        const size_t smem = lauum_grad64_smem(d);
        static SmemOptIn optin64;
        if (cudaError_t e = optin64.ensure(lauum_grad64_kernel, smem); e != cudaSuccess) return e;
        static bool carve = false;
        if (!carve) {
            cudaFuncSetAttribute(lauum_grad64_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            carve = true;
        }
        p.ntile = T * (T + 1);
        lauum_grad64_kernel<<<dim3(p.ntile, 1, B), LG2_THREADS, smem, st>>>(p);
        return cudaGetLastError();
    }
    const size_t smem = Wout ? gemm_smem_bytes<128, 128, false, false>() : lauum_grad_smem(d);
    static SmemOptIn optin;
    if (cudaError_t e = optin.ensure(lauum_grad_kernel, smem); e != cudaSuccess) return e;
    lauum_grad_kernel<<<dim3(order ? T * T : p.ntile, 1, B), WS_THREADS, smem, st>>>(p);
    return cudaGetLastError();
}

int lauum_grad_ntiles(int npad) {
    const int T = npad / 128;
    static int wide64 = -1;
    if (wide64 < 0) {
        const char* e = getenv("GPE_LG64");
        wide64 = e ? atoi(e) : 0;
    }
    return wide64 ? T * (T + 1) : T * (T + 1) / 2;       // partial rows per item: one per 128x64 or per 128x128 lower tile
}

}  // namespace gpe
