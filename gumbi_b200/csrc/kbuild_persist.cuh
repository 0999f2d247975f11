// K-build v6: persistent strip kernel for the single-term models (pymc/GP.py:410,462,561,569):
//   K_ij = eta^2 k(|u_i - u_j|) * prod_f B_f[c_f(i), c_f(j)]        u = x / ls,  f <= 2 Coregion factors
// with the augmentation of the training matrix (noise + jitter on the diagonal, y^T in row N, identity padding) exactly as
// kbuild_dmma_kernel<TRAIN> writes it.  Additive / Linear models stay on kbuild_dmma_kernel.
//
// Structure: persistent CTAs (grid = SMs x resident CTAs) pull (row tile, strip of column tiles) items off an atomic counter;
// column-side data (features, squared norms, Coregion levels) arrive through a 3-stage cp.async ring filled by ONE rotating producer
// warp; the row side of an item lives in registers as ready-made DMMA A fragments pre-scaled by -2c; r^2 = s_i + s_j + DMMA (PyMC's
// own expanded form, Stationary.square_dist); exp() = 2048-entry 2^(j/2048) table in shared memory (eta^2 folded in) + cubic.
//
// What bounds it on B200 (round-2 measurements, profiles/r02ka..ke_micro_kbuild.log): an SM sub-partition DISPATCHES like a single-issue
// core on which a DFMA / DMUL / DADD takes 2 slots and a DMMA.8x8x4 ~14, everything else 1 -- ablations (no stores / no lookup / no
// DMMA / no evaluation) add up term by term, and neither more resident warps, more registers (interleaved entries), a
// conflict-free 16-copy table, 128-byte store rows nor fewer barriers moved the time by more than 2 %.  Per warp-entry (32 matrix
// entries): fp64 slots 2 (1 + 7) + 14 = 30 (ExpQuad, d = 8) / 2 (1 + 5 + 7 + 3) + 14 = 46 (Matern-5/2); HBM time at 6468 GB/s is 45
// slots.  So the only lever left is the count of OTHER instructions, which v4 -> v6 took from 33 to ~17 per entry:
//   * fp64 constants are KERNEL PARAMETERS (constant-bank / register operands of the DFMA itself) -- as literals ptxas re-materialised
//     them with two moves per use;
//   * clip(r^2, 0) (+ 1e-12 under the Matern square root) and the cap that keeps the exp argument reduction valid are two integer
//     min/max on the high word, in place;
//   * the table stores each entry with its high word biased by -(j << 9), so that  hi + (n << 9)  (one IMAD) applies the binary
//     exponent n >> 11 and cancels the index bits; the underflow flush is tested once per group of four entries;
//   * one producer warp issues a tile's copies (a lane = 16 bytes of every row) instead of 256 threads re-deriving slot addresses;
//     output row pointers are set up once per item; the tile is evaluated in two column halves (8 live accumulators).
// Measured, N = 32768 (same box, profiles/r02ke_micro_kbuild.log): ExpQuad d = 8 1.100 -> 0.90 ms (0.74 of the HBM peak), Matern-5/2
// 1.313 -> 1.11 ms (0.60), 2-output ICM d = 4 1.055 -> 0.89 ms (0.75).
// Accuracy: table exact to 0.5 ulp, |reduced argument| <= ln2/4096 so the cubic truncates at 2^-58 relative; the one-step
// argument reduction carries |x| * 2^-54 -- the same size as the rounding of x itself.  Entrywise gate vs the oracle: 5e-12.
#pragma once
#include <cstring>
#include "kbuild.cuh"

namespace gb2 {

constexpr int KB4_TAB = 2048;       // entries of the exp table: 2^(j/2048) (a conflict-free 16-copy 256-entry table + quartic measured 8-13 % slower,
                                    // profiles/r02g_micro_kbuild.log)
constexpr int KB4_TAB_LOG2 = 11;
constexpr int KB4_STAGES = 3;
constexpr int KB4_TS = 68;          // shared row stride (doubles) of a staged column tile: conflict-free 4x8 DMMA B fragments
constexpr int KB4_MAXCG = 2;

__device__ double g_exp2_tab2k[KB4_TAB];   // 2^(j/2048), filled by the host at gb2_create

struct KB4Args {
    const double* Fi; int64_t stride_i; int64_t n_i;     // row side: feature table of the term (row 0 = first scaled coordinate)
    const double* Fj; int64_t stride_j; int64_t n_j;     // column side
    const int* Ci; const int* Cj;                        // category tables (row f = Coregion factor f of this term), or nullptr
    const double* Btab;
    const double* y; double* out; int64_t ld;
    int n_row_tiles, n_col_tiles, strip;                 // strip = column tiles per work item
    int own_stride, own_rank, compact;
    int* ctr;                                            // [0] next item, [1] CTAs finished (self-resetting)
    // constants of the per-entry code, per kind (zs = 1/2 for ExpQuad: exp(-z/2), 1 for the Matern family: exp(-w)); kernel parameters
    // so that they are constant-bank operands of the instruction that uses them
    double cA;        // -zs * 2048 / ln 2
    double cR;        // ln 2 / (2048 zs)
    double q3;        // -zs^3 / 6
    double c13;       // 1/3   (Matern-5/2 polynomial)
    double c375;      // 3/8   (square-root correction)
    int hi_zmin;      // high word of c * 1e-12 (Matern: clip(r^2, 0) + 1e-12 of euclidean_dist) or 0 (ExpQuad: clip(r^2, 0))
    int hi_zmax;      // high word of the cap on z: exp argument -1000 (2000 for ExpQuad, 1e6 under the Matern square root)
};

// Per kind: the accumulator holds  z = c r^2 (+ c 1e-12)  with c = kb4_scale(kind)  (row side pre-scaled by -2c, norms by c)
__host__ __device__ inline double kb4_scale(int kind) {
    switch (kind) {
        case GB2_EXPQUAD: return 1.0;
        case GB2_MATERN52: return 5.0;     // w = sqrt(5) r
        case GB2_MATERN32: return 3.0;     // w = sqrt(3) r
        case GB2_MATERN12: return 1.0;     // w = r
        default: return 0.25;              // Exponential: exp(-r/2), w = r/2
    }
}

inline void kb4_set_constants(KB4Args& a, int kind) {
    const long double LOG2E = 1.442695040888963407359924681001892137L, LN2 = 0.693147180559945309417232121458176568L;
    const double zs = kind == GB2_EXPQUAD ? 0.5 : 1.0;
    a.cA = -(double)(LOG2E * (long double)(zs * KB4_TAB));      // zs * 2048 is a power of two: correctly rounded log2(e), scaled exactly
    a.cR = (double)(LN2 / (long double)(zs * KB4_TAB));
    a.q3 = -zs * zs * zs / 6.0;
    a.c13 = 1.0 / 3.0;
    a.c375 = 0.375;
    const double zmin = kind == GB2_EXPQUAD ? 0.0 : kb4_scale(kind) * 1e-12;
    int64_t bits;
    memcpy(&bits, &zmin, 8);
    a.hi_zmin = (int)(bits >> 32);
    const double zmax = kind == GB2_EXPQUAD ? 2000.0 : 1.0e6;
    memcpy(&bits, &zmax, 8);
    a.hi_zmax = (int)(bits >> 32);
}

// Per-entry arithmetic, written for GROUPS of G entries, statement by statement across the group (kb4_eval).
//   exp: eta^2 * exp(-x * zs), table eta^2 * 2^(j/2048), j = n mod 2048, whose entries carry the high word biased by -(j << 9):
//   hi + (n << 9) = hi(eta^2 2^(j/2048)) + ((n >> 11) << 20), i.e. times 2^(n >> 11) .  rr = x + n cR = -(reduced argument) / zs,
//   cubic of exp(-zs rr) - 1 with coefficients -zs, zs^2 / 2, -zs^3 / 6.
//   Huge scaled distances: z is clipped from above (high word, same instruction pair as the clip from below) where the exp argument
//   reaches -1000, so that n = round(...) never wraps; any result whose exponent field underflows -- every clipped argument for
//   eta^2 < 1e125, or a tiny eta^2 on top of a tiny exp -- is flushed to an exact 0 (the oracle's exp underflows to 0 at -745).
//   clip(r^2, 0, inf) of Stationary.square_dist (+ the 1e-12 of euclidean_dist, already inside z, for the Matern family) is ONE signed
//   integer max on the high word (z >= 0 orders like its bit pattern, a negative z has the sign bit set): the expanded form can
//   come out negative by ~1e-16 |u|^2 (duplicated points, the diagonal) -- NaN under the square root once that exceeds the epsilon.
//   A clipped value keeps its low word: c 1e-12 (1 + < 2^-20), or a denormal for ExpQuad -- both far below the rounding of r^2.
//   sqrt(a) for a normal positive a: MUFU.RSQ64H seed (2^-22) + one third-order correction, 5 fp64 operations, residual ~2^-67.
template <int KIND, int G>
__device__ __forceinline__ void kb4_eval(int kind_rt, double* v, const unsigned char* __restrict__ tab, const KB4Args& a) {
    constexpr int SHIFT = 20 - KB4_TAB_LOG2;
    const int kind = KIND >= 0 ? KIND : kind_rt;
    const bool eq = kind == GB2_EXPQUAD;
    constexpr double MAGIC = 6755399441055744.0;            // 1.5 * 2^52
    const double h2 = eq ? 0.125 : 0.5, m1 = eq ? -0.5 : -1.0;          // zs^2 / 2, -zs
    // every statement runs over the G entries of the group before the next one starts: G independent dependency chains, interleaved
    // in program order (a warp issues in order; one chain alone leaves the fp64 pipe idle for the 8.5 cycles of every DFMA)
    double x[G], rr[G], t[G];
    int n[G];
    int2 T[G];
#pragma unroll
    for (int q = 0; q < G; q++) {
        // clip on the high word, in place (the low word keeps its register)
        x[q] = v[q];
        asm("{\n\t.reg .b32 lo, hi;\n\tmov.b64 {lo, hi}, %0;\n\tmax.s32 hi, hi, %1;\n\tmin.s32 hi, hi, %2;\n\tmov.b64 %0, {lo, hi};\n\t}" : "+d"(x[q]) : "r"(a.hi_zmin), "r"(a.hi_zmax));
    }
    if (!eq) {
        double y[G], e[G];
#pragma unroll
        for (int q = 0; q < G; q++) asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y[q]) : "d"(x[q]));
#pragma unroll
        for (int q = 0; q < G; q++) t[q] = x[q] * y[q];
#pragma unroll
        for (int q = 0; q < G; q++) e[q] = fma(-t[q], y[q], 1.0);
#pragma unroll
        for (int q = 0; q < G; q++) y[q] = fma(a.c375, e[q], 0.5);
#pragma unroll
        for (int q = 0; q < G; q++) e[q] = t[q] * e[q];
#pragma unroll
        for (int q = 0; q < G; q++) x[q] = fma(e[q], y[q], t[q]);
    }
#pragma unroll
    for (int q = 0; q < G; q++) t[q] = fma(x[q], a.cA, MAGIC);
#pragma unroll
    for (int q = 0; q < G; q++) {
        n[q] = __double2loint(t[q]);                         // round(-zs x 2048 / ln2) <= 0
        T[q] = *reinterpret_cast<const int2*>(tab + ((n[q] << 3) & ((KB4_TAB - 1) << 3)));
    }
#pragma unroll
    for (int q = 0; q < G; q++) t[q] = t[q] - MAGIC;
#pragma unroll
    for (int q = 0; q < G; q++) rr[q] = fma(t[q], a.cR, x[q]);       // |rr| <= ln2 / (4096 zs)
#pragma unroll
    for (int q = 0; q < G; q++) t[q] = fma(rr[q], a.q3, h2);
#pragma unroll
    for (int q = 0; q < G; q++) t[q] = fma(t[q], rr[q], m1);
#pragma unroll
    for (int q = 0; q < G; q++) rr[q] = rr[q] * t[q];                // exp(-zs rr) - 1
    if (kind == GB2_MATERN52) {
#pragma unroll
        for (int q = 0; q < G; q++) t[q] = fma(a.c13, x[q], 1.0);
#pragma unroll
        for (int q = 0; q < G; q++) t[q] = fma(t[q], x[q], 1.0);     // 1 + w + w^2/3 = 1 + sqrt5 r + 5/3 r^2
    } else if (kind == GB2_MATERN32) {
#pragma unroll
        for (int q = 0; q < G; q++) t[q] = 1.0 + x[q];
    }
    // binary exponent; an exponent field that underflows (exp < ~2^-1022 / eta^2, which includes every clipped-from-above argument)
    // flushes the entry to an exact 0 -- tested once for the group, the per-entry selects sit on the rarely taken path
    int hmin = 0x7fffffff;
#pragma unroll
    for (int q = 0; q < G; q++) {
        T[q].y += n[q] << SHIFT;
        hmin = min(hmin, T[q].y);
    }
    if (__builtin_expect(hmin < 0x00100000, 0)) {
#pragma unroll
        for (int q = 0; q < G; q++)
            if (T[q].y < 0x00100000) T[q] = make_int2(0, 0);
    }
#pragma unroll
    for (int q = 0; q < G; q++) {
        const double Ts = __hiloint2double(T[q].y, T[q].x);
        rr[q] = fma(Ts, rr[q], Ts);
    }
    if (kind == GB2_MATERN52 || kind == GB2_MATERN32) {
#pragma unroll
        for (int q = 0; q < G; q++) v[q] = rr[q] * t[q];
    } else {
#pragma unroll
        for (int q = 0; q < G; q++) v[q] = rr[q];                    // ExpQuad, Matern12, Exponential
    }
}

__device__ __forceinline__ void kb4_cp_async16(unsigned smem_dst, const void* gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_dst), "l"(gmem_src));
}

// KS = DMMA k-steps (d <= 4 KS), NCG = Coregion factors of the term
template <bool TRAIN, int KIND, int KS, int NCG, int OCC>
__global__ void __launch_bounds__(KB_THREADS, OCC)
kbuild_persist_kernel(KParams kp, KB4Args a) {
    extern __shared__ __align__(16) unsigned char kb_smem[];
    constexpr int ROWS = 4 * KS + 1;                   // staged rows of a column tile: 4 KS features (rows d .. 4 KS - 1 stay zero) + the squared norms
    constexpr int TILE_D = ROWS * KB4_TS;              // doubles per staged column tile
    constexpr int NCGX = NCG > 0 ? NCG : 1;
    const unsigned char* sTab = kb_smem;                                        // [2048] x 8 bytes, high words biased (kb4_eval)
    double* sB = reinterpret_cast<double*>(kb_smem) + KB4_TAB;                  // [STAGES][ROWS][TS]
    double* sBt = sB + KB4_STAGES * TILE_D;                                     // [NCG][P*P <= 64] Coregion tables
    int* sCj = reinterpret_cast<int*>(sBt + NCGX * GB2_MAX_P * GB2_MAX_P);      // [STAGES][NCG][64]
    __shared__ int s_item;

    const TermDev& T = kp.t[0];
    const int d = T.d;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t4 = lane & 3;
    const int r0 = (warp >> 1) * 16, c0 = (warp & 1) * 32;
    const int kind_rt = T.kind;
    const double csc = kb4_scale(KIND >= 0 ? KIND : kind_rt);
    const double ceps = (KIND >= 0 ? KIND : kind_rt) == GB2_EXPQUAD ? 0.0 : csc * 1e-12;
    const double* Fit = a.Fi + (int64_t)T.feat_off * a.stride_i;
    const double* Fjt = a.Fj + (int64_t)T.feat_off * a.stride_j;

    // one-time per CTA: exp table scaled by eta^2 (biased high words), zeroed stages, Coregion tables
    for (int e = tid; e < KB4_TAB; e += KB_THREADS) {
        const double v = T.eta2 * g_exp2_tab2k[e];
        reinterpret_cast<int2*>(kb_smem)[e] = make_int2(__double2loint(v), __double2hiint(v) - (e << (20 - KB4_TAB_LOG2)));
    }
    for (int e = tid; e < KB4_STAGES * TILE_D; e += KB_THREADS) sB[e] = 0.0;
    if (NCG > 0)
        for (int e = tid; e < NCG * GB2_MAX_P * GB2_MAX_P; e += KB_THREADS) {
            const int f = e / (GB2_MAX_P * GB2_MAX_P), q = e % (GB2_MAX_P * GB2_MAX_P);
            sBt[e] = q < T.cg_P[f] * T.cg_P[f] ? a.Btab[T.cg_Boff[f] + q] : 0.0;
        }

    // Prefetch of one ring stage (a column tile: d feature rows + the squared norms of 64 points, + Coregion levels) by ONE warp: a lane
    // copies 16 bytes of every row.  (v5 spread the chunks over all 256 threads, and every thread re-derived its slot addresses per
    // tile: 9 % of all instructions.)  The producer role rotates over the warps with the tile counter.
    const unsigned sB_s = (unsigned)__cvta_generic_to_shared(sB), sCj_s = (unsigned)__cvta_generic_to_shared(sCj);
    auto prefetch = [&](int jt, int stage) {
        const double* src = Fjt + (int64_t)jt * KB_T + lane * 2;
        const unsigned dst = sB_s + (unsigned)((stage * TILE_D + lane * 2) * 8);
#pragma unroll
        for (int k = 0; k < 4 * KS; k++)
            if (k < d) kb4_cp_async16(dst + k * KB4_TS * 8, src + (int64_t)k * a.stride_j);
        kb4_cp_async16(dst + 4 * KS * KB4_TS * 8, src + (int64_t)d * a.stride_j);
        if (NCG > 0 && lane < NCG * 16) {
            const int f = lane >> 4, ch = lane & 15;
            kb4_cp_async16(sCj_s + (unsigned)(((stage * NCGX + f) * KB_T + ch * 4) * 4), a.Cj + (int64_t)T.cg_cat[f] * a.stride_j + (int64_t)jt * KB_T + ch * 4);
        }
    };

    // work items: TRAIN -- row tile bi has strips 0 .. bi / strip (lower triangle); groups of `strip` row tiles share a strip count
    const int strip = a.strip;
    int n_items;
    if (TRAIN) {
        const int ng = (a.n_row_tiles + strip - 1) / strip;          // row groups; group q (rows q*strip ..) has q + 1 strips per row
        const int full = ng - 1;
        n_items = strip * (full * (full + 1) / 2) + (a.n_row_tiles - full * strip) * ng;
    } else {
        n_items = a.n_row_tiles * ((a.n_col_tiles + strip - 1) / strip);
    }
    const int n_full_cols = (int)(a.n_j / KB_T);             // column tiles that lie completely inside the data

    for (;;) {
        __syncthreads();                                   // everybody is done with the previous item's stages (and with s_item)
        if (tid == 0) s_item = atomicAdd(a.ctr, 1);
        __syncthreads();
        const int w = s_item;
        if (w >= n_items) break;
        int bi, sidx;
        if (TRAIN) {
            // group q = rows [q strip, (q+1) strip) with q + 1 strips each; strip * q (q + 1) / 2 items precede it
            int q = (int)((sqrt(8.0 * (double)w / strip + 1.0) - 1.0) * 0.5);
            while (strip * ((q + 1) * (q + 2) / 2) <= w) q++;
            while (q > 0 && strip * (q * (q + 1) / 2) > w) q--;
            const int rem = w - strip * (q * (q + 1) / 2);
            bi = q * strip + rem / (q + 1);
            sidx = rem % (q + 1);
            if (a.own_stride > 1 && ((bi * KB_T) / TILE) % a.own_stride != a.own_rank) continue;   // another rank's row block
        } else {
            const int ns = (a.n_col_tiles + strip - 1) / strip;
            bi = w / ns;
            sidx = w % ns;
        }
        const int jt0 = sidx * strip;
        int jt1 = jt0 + strip < a.n_col_tiles ? jt0 + strip : a.n_col_tiles;
        if (TRAIN && jt1 > bi + 1) jt1 = bi + 1;
        const int64_t i0 = (int64_t)bi * KB_T;

        // column ring: two tiles in flight before the first one is consumed (producer of tile jt0 + s = warp s mod 8)
        if (warp == 0) prefetch(jt0, 0);
        if (warp == 1 && jt0 + 1 < jt1) prefetch(jt0 + 1, 1);
        asm volatile("cp.async.commit_group;\n" ::);

        // row side of the item as DMMA A fragments: lane (g, t4) holds feature 4 ks + t4 of rows r0 + g and r0 + 8 + g, scaled by -2c
        double af[2][KS], si[2];
        int rowoff[NCGX][2];
#pragma unroll
        for (int mi = 0; mi < 2; mi++) {
            const int64_t gi = i0 + r0 + mi * 8 + g;
#pragma unroll
            for (int ks = 0; ks < KS; ks++) {
                const int k = ks * 4 + t4;
                af[mi][ks] = k < d ? -2.0 * csc * Fit[(int64_t)k * a.stride_i + gi] : 0.0;
            }
            si[mi] = fma(csc, Fit[(int64_t)d * a.stride_i + gi], ceps);
#pragma unroll
            for (int f = 0; f < NCG; f++) rowoff[f][mi] = f * GB2_MAX_P * GB2_MAX_P + a.Ci[(int64_t)T.cg_cat[f] * a.stride_i + gi] * T.cg_P[f];
        }
        // output rows r0 + g and r0 + 8 + g of this lane (the same 128-row block: one row mapping), at this lane's first column of tile 0
        double* drow = a.out + kb_out_row(i0 + r0 + g, a.own_stride, a.own_rank, TRAIN ? a.compact : 0) * a.ld + c0 + 2 * t4;
        const int64_t ld8 = 8 * a.ld;
        const bool row_full = i0 + KB_T <= a.n_i;

        int st = 0, pf = 2;
        for (int jt = jt0; jt < jt1; jt++) {
            // the group that filled this stage was committed one or two tiles ago by its producer warp; every other warp has nothing pending
            asm volatile("cp.async.wait_group 0;\n" ::);
            __syncthreads();
            if (jt + 2 < jt1 && warp == ((jt + 2 - jt0) & 7)) prefetch(jt + 2, pf);
            asm volatile("cp.async.commit_group;\n" ::);
            pf = pf == KB4_STAGES - 1 ? 0 : pf + 1;
            const double* cB = sB + st * TILE_D;
            const int* cC = sCj + st * NCGX * KB_T;
            st = st == KB4_STAGES - 1 ? 0 : st + 1;
            const int64_t j0 = (int64_t)jt * KB_T;
            const bool interior = row_full && jt < n_full_cols && (!TRAIN || bi != jt);
            double* dtile = drow + j0;
            // two column halves of this warp's 16 x 32 patch: 8 live accumulators per lane each
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int cb = c0 + h * 16;
                double acc[2][4];                      // [nj][2 mi + e]
#pragma unroll
                for (int nj = 0; nj < 2; nj++) {
                    const double2 sj = *reinterpret_cast<const double2*>(cB + 4 * KS * KB4_TS + cb + nj * 8 + 2 * t4);
#pragma unroll
                    for (int mi = 0; mi < 2; mi++) {
                        acc[nj][2 * mi + 0] = fma(csc, sj.x, si[mi]);
                        acc[nj][2 * mi + 1] = fma(csc, sj.y, si[mi]);
                    }
                }
#pragma unroll
                for (int ks = 0; ks < KS; ks++) {
                    const double* pb = cB + (ks * 4 + t4) * KB4_TS + cb + g;
#pragma unroll
                    for (int nj = 0; nj < 2; nj++) {
                        const double b = pb[nj * 8];
                        kb_dmma(acc[nj][0], acc[nj][1], af[0][ks], b);
                        kb_dmma(acc[nj][2], acc[nj][3], af[1][ks], b);
                    }
                }
#pragma unroll
                for (int nj = 0; nj < 2; nj++) {
                    kb4_eval<KIND, 4>(kind_rt, acc[nj], sTab, a);
                    if (NCG > 0) {
#pragma unroll
                        for (int f = 0; f < NCG; f++) {
                            const int2 c2 = *reinterpret_cast<const int2*>(cC + f * KB_T + cb + nj * 8 + 2 * t4);
#pragma unroll
                            for (int mi = 0; mi < 2; mi++) {
                                acc[nj][2 * mi + 0] *= sBt[rowoff[f][mi] + c2.x];
                                acc[nj][2 * mi + 1] *= sBt[rowoff[f][mi] + c2.y];
                            }
                        }
                    }
                }
                if (interior) {
#pragma unroll
                    for (int mi = 0; mi < 2; mi++)
#pragma unroll
                        for (int nj = 0; nj < 2; nj++)
                            *reinterpret_cast<double2*>(dtile + mi * ld8 + h * 16 + nj * 8) = make_double2(acc[nj][2 * mi], acc[nj][2 * mi + 1]);
                    continue;
                }
                // boundary tiles (the diagonal tile of a training row, ragged edges): augmentation and padding per entry
#pragma unroll
                for (int mi = 0; mi < 2; mi++) {
                    const int64_t gi = i0 + r0 + mi * 8 + g;
#pragma unroll
                    for (int nj = 0; nj < 2; nj++) {
                        double o[2];
#pragma unroll
                        for (int e = 0; e < 2; e++) {
                            const int64_t gj = j0 + cb + nj * 8 + 2 * t4 + e;
                            double v = acc[nj][2 * mi + e];
                            if (TRAIN) {
                                if (gi < a.n_i && gj < a.n_j) {
                                    if (gi == gj) {
                                        double nz = kp.sigma2;
                                        if (kp.noise_cat >= 0) {
                                            const int c = a.Ci[(int64_t)kp.noise_cat * a.stride_i + gi];
                                            nz *= __ldg(a.Btab + kp.noise_Boff + c * kp.noise_P + c);
                                        }
                                        v += nz + kp.jitter;
                                    }
                                } else if (gi == a.n_i && gj < a.n_j) {
                                    v = a.y[gj];
                                } else {
                                    v = (gi == gj) ? 1.0 : 0.0;
                                }
                            } else {
                                if (gi >= a.n_i || gj >= a.n_j) v = 0.0;
                            }
                            o[e] = v;
                        }
                        *reinterpret_cast<double2*>(dtile + mi * ld8 + h * 16 + nj * 8) = make_double2(o[0], o[1]);
                    }
                }
            }
        }
    }
    // self-resetting counters: the last CTA out leaves both at zero for the next launch on this handle
    if (tid == 0) {
        __threadfence();
        if (atomicAdd(a.ctr + 1, 1) == (int)gridDim.x - 1) { a.ctr[0] = 0; a.ctr[1] = 0; __threadfence(); }
    }
}

template <int KS, int NCG>
constexpr size_t kb4_smem_bytes() {
    return (size_t)(KB4_TAB + KB4_STAGES * (4 * KS + 1) * KB4_TS + (NCG > 0 ? NCG : 1) * GB2_MAX_P * GB2_MAX_P) * sizeof(double) +
           (size_t)KB4_STAGES * (NCG > 0 ? NCG : 1) * KB_T * sizeof(int);
}

// models the persistent kernel covers: one term, no Linear part, <= 2 Coregion factors, any stationary kind
inline bool kb4_eligible(const KParams& kp, bool train, int compact) {
    return kp.n_terms == 1 && kp.t[0].n_lin == 0 && kp.t[0].n_coreg <= KB4_MAXCG && kp.t[0].d >= 1 && kp.t[0].d <= 16 && (train || !compact);
}

constexpr int KB4_OCC = 4;   // resident CTAs per SM the register allocation is bounded for (tools/micro_kbuild.cu times 2, 3, 4)

template <bool TRAIN, int KIND, int KS, int NCG, int OCC = KB4_OCC>
inline void kb4_launch_one(cudaStream_t s, int n_sm, const KParams& kp, const KB4Args& a) {
    static bool configured = false;   // per instantiation; cudaFuncSetAttribute is idempotent, the flag only saves the call
    if (!configured) {
        cudaFuncSetAttribute(kbuild_persist_kernel<TRAIN, KIND, KS, NCG, OCC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kb4_smem_bytes<KS, NCG>());
        configured = true;
    }
    kbuild_persist_kernel<TRAIN, KIND, KS, NCG, OCC><<<n_sm * OCC, KB_THREADS, kb4_smem_bytes<KS, NCG>(), s>>>(kp, a);
}

template <bool TRAIN, int KIND>
inline void kb4_launch_kind(cudaStream_t s, int n_sm, const KParams& kp, const KB4Args& a) {
    const int ks = (kp.t[0].d + 3) / 4, ncg = kp.t[0].n_coreg;
#define GB2_KB4(KS_, NCG_) if (ks == KS_ && ncg == NCG_) return kb4_launch_one<TRAIN, KIND, KS_, NCG_>(s, n_sm, kp, a)
    GB2_KB4(1, 0); GB2_KB4(2, 0); GB2_KB4(3, 0); GB2_KB4(4, 0);
    GB2_KB4(1, 1); GB2_KB4(2, 1); GB2_KB4(3, 1); GB2_KB4(4, 1);
    GB2_KB4(1, 2); GB2_KB4(2, 2); GB2_KB4(3, 2); GB2_KB4(4, 2);
#undef GB2_KB4
}

template <bool TRAIN>
inline void kbuild_persist_launch(cudaStream_t s, int n_sm, const KParams& kp, const double* Btab, const double* Fi, const int* Ci, int64_t stride_i,
                                  int64_t n_i, const double* Fj, const int* Cj, int64_t stride_j, int64_t n_j, int n_row_tiles, int n_col_tiles,
                                  const double* y, double* out, int64_t ld, int own_stride, int own_rank, int compact, int* ctr) {
    KB4Args a{};
    a.Fi = Fi; a.stride_i = stride_i; a.n_i = n_i; a.Fj = Fj; a.stride_j = stride_j; a.n_j = n_j; a.Ci = Ci; a.Cj = Cj; a.Btab = Btab;
    a.y = y; a.out = out; a.ld = ld; a.n_row_tiles = n_row_tiles; a.n_col_tiles = n_col_tiles;
    a.own_stride = own_stride; a.own_rank = own_rank; a.compact = compact; a.ctr = ctr;
    // strip length: long strips amortise the per-item row set-up, short ones balance small problems (>= ~16 items per CTA)
    const double tiles = TRAIN ? 0.5 * n_row_tiles * (double)(n_row_tiles + 1) / (own_stride > 1 ? own_stride : 1) : (double)n_row_tiles * n_col_tiles;
    int strip = (int)(tiles / (16.0 * n_sm * KB4_OCC));
    a.strip = strip < 1 ? 1 : (strip > 8 ? 8 : strip);
    kb4_set_constants(a, kp.t[0].kind);
    if (kp.t[0].kind == GB2_EXPQUAD) kb4_launch_kind<TRAIN, GB2_EXPQUAD>(s, n_sm, kp, a);
    else if (kp.t[0].kind == GB2_MATERN52) kb4_launch_kind<TRAIN, GB2_MATERN52>(s, n_sm, kp, a);
    else kb4_launch_kind<TRAIN, -1>(s, n_sm, kp, a);
}

}  // namespace gb2
