// K-build v5: persistent strip kernel for the single-term models (pymc/GP.py:410,462,561,569):
//   K_ij = eta^2 k(|u_i - u_j|) * prod_f B_f[c_f(i), c_f(j)]        u = x / ls,  f <= 2 Coregion factors
// with the augmentation of the training matrix (noise + jitter on the diagonal, y^T in row N, identity padding) exactly as
// kbuild_dmma_kernel<TRAIN> writes it.  Additive / Linear models stay on kbuild_dmma_kernel.
//
// Structure (round 2, first version): persistent CTAs (grid = SMs x resident CTAs) pull (row tile, strip of column tiles) items off an
// atomic counter; column-side data (features, squared norms, Coregion levels) arrive through a 3-stage cp.async ring; the row side
// of an item lives in registers as ready-made DMMA A fragments pre-scaled by -2c; r^2 = s_i + s_j + DMMA (PyMC's own expanded
// form, Stationary.square_dist); exp() = 2048-entry 2^(j/2048) table in shared memory (eta^2 folded in) + cubic.
//
// Why v5 (ncu of v4 at C4, profiles/r02t_ncu_kbuild_matern_summary.txt + its SASS): 50 warp instructions per entry, of which only 17
// use the fp64 pipe (16 scalar + 1 DMMA = 48 of its cycles) -- ISSUE SLOTS (50 cycles per warp-entry) bounded the kernel as much as the
// fp64 pipe did, and each ran at 55 %.  The other 33: literal fp64 constants re-materialised with two moves per use (64
// registers), a 4-instruction integer clip, 10 integer instructions around the table lookup, and ~12 per entry of per-tile
// overhead (64-bit address arithmetic of the prefetch and of the stores, done per tile).  Here:
//   * fp64 constants are KERNEL PARAMETERS (constant-bank operands of the DFMA itself, no moves);
//   * clip(r^2, 0) (+ 1e-12 under the Matern square root) is ONE signed integer max on the high word;
//   * the table stores each entry with its high word biased by -(j << 9), so that  hi + (n << 9)  (one LEA) applies the binary
//     exponent n >> 11 and cancels the index bits -- no mask / shift / add chain;
//   * per-thread prefetch slots (source pointer, shared address) and the two output row pointers are set up once per CTA / item,
//     a tile costs one 64-bit add each; the tile is evaluated in two column halves (8 live accumulators, not 16) so that the
//     64-register budget leaves room to interleave independent entries.
// fp64 slots per entry are unchanged: 1 + d + 7 (ExpQuad), 1 + d + 5 (sqrt) + 7 + 3 (Matern-5/2).
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
}

// eta^2 * exp(-z * zs) for 0 <= z, through  n = round(-z * zs * 2048 / ln 2)  and the table  eta^2 * 2^(j/2048), j = n mod 2048, whose
// entries carry the high word biased by -(j << 9):  hi + (n << 9) = hi(eta^2 2^(j/2048)) + ((n >> 11) << 20), i.e. times 2^(n >> 11).
//   rr = z + n cR = -(reduced argument) / zs, cubic of exp(-zs rr) - 1 with coefficients -zs, zs^2 / 2, -zs^3 / 6.
// Arguments with zs z >= 693 (exp < 2^-1000 ~ 1e-301) return an exact 0: decided on the HIGH WORD OF z (z >= 0 orders like its bit
// pattern), because for huge scaled distances (z > ~1e6) the low word of t -- n -- wraps around and must not be consulted.  A
// result whose exponent field would underflow (tiny eta^2 on top of a tiny exp) is flushed to 0 as well.
template <int ZS2>   // ZS2 = 2 * zs
__device__ __forceinline__ double kb4_exp(double z, const unsigned char* __restrict__ tab, const KB4Args& a) {
    constexpr double zs = 0.5 * ZS2;
    constexpr double MAGIC = 6755399441055744.0;            // 1.5 * 2^52
    const double t = fma(z, a.cA, MAGIC);
    const int n = __double2loint(t);                         // round(-zs z 2048 / ln2) <= 0
    const double kf = t - MAGIC;
    const double rr = fma(kf, a.cR, z);                     // |rr| <= ln2 / (4096 zs)
    const double p1 = fma(rr, a.q3, 0.5 * zs * zs);
    const double p2 = fma(p1, rr, -zs);
    const double m = rr * p2;                                // exp(-zs rr) - 1
    const int2 T = *reinterpret_cast<const int2*>(tab + ((n << 3) & ((KB4_TAB - 1) << 3)));
    const int hi = T.y + (n << 9);
    constexpr int HI_ZMAX = ZS2 == 1 ? 0x4095A800 /* 1386.0 */ : 0x4085A800 /* 693.0 */;
    const bool tiny = hi < 0x00100000 || __double2hiint(z) >= HI_ZMAX;
    const double Ts = __hiloint2double(tiny ? 0 : hi, tiny ? 0 : T.x);
    return fma(Ts, m, Ts);
}

// sqrt(a) for a normal positive a: MUFU.RSQ64H seed (2^-22) + one third-order correction, 5 fp64 operations, residual ~2^-67
__device__ __forceinline__ double kb4_sqrt(double a, double c375) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
    const double t = a * y;
    const double e = fma(-t, y, 1.0);
    const double p = fma(c375, e, 0.5);
    const double te = t * e;
    return fma(te, p, t);
}

template <int KIND>
__device__ __forceinline__ double kb4_value(int kind_rt, double z, const unsigned char* __restrict__ tab, const KB4Args& a) {
    const int kind = KIND >= 0 ? KIND : kind_rt;
    // clip(r^2, 0, inf) of Stationary.square_dist (+ the 1e-12 of euclidean_dist, already inside z, for the Matern family) as ONE signed
    // integer max on the high word (z >= 0 orders like its bit pattern, a negative z has the sign bit set): the expanded form can
    // come out negative by ~1e-16 |u|^2 (duplicated points, the diagonal) -- NaN under the square root once that exceeds the epsilon.
    // A clipped value keeps its low word: c 1e-12 (1 + < 2^-20), or a denormal for ExpQuad -- both far below the rounding of r^2
    z = __hiloint2double(max(__double2hiint(z), a.hi_zmin), __double2loint(z));
    if (kind == GB2_EXPQUAD) return kb4_exp<1>(z, tab, a);
    const double w = kb4_sqrt(z, a.c375);
    const double e = kb4_exp<2>(w, tab, a);
    if (kind == GB2_MATERN52) return e * fma(fma(a.c13, w, 1.0), w, 1.0);   // 1 + w + w^2/3 = 1 + sqrt5 r + 5/3 r^2
    if (kind == GB2_MATERN32) return e * (1.0 + w);
    return e;                                         // Matern12, Exponential
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
    constexpr int STAGE_D = ROWS * KB4_TS;             // doubles per stage
    constexpr int NSLOT = (ROWS + 7) / 8;              // 16-byte prefetch chunks per thread and tile: 32 chunks per row, 8 rows per pass
    const unsigned char* sTab = kb_smem;                                        // [2048] x 8 bytes, high words biased (kb4_exp)
    double* sB = reinterpret_cast<double*>(kb_smem) + KB4_TAB;                  // [STAGES][ROWS][TS]
    double* sBt = sB + KB4_STAGES * STAGE_D;                                    // [NCG][P*P <= 64] Coregion tables
    int* sCj = reinterpret_cast<int*>(sBt + (NCG > 0 ? NCG : 1) * GB2_MAX_P * GB2_MAX_P);   // [STAGES][NCG][64]
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
        reinterpret_cast<int2*>(kb_smem)[e] = make_int2(__double2loint(v), __double2hiint(v) - (e << 9));
    }
    for (int e = tid; e < KB4_STAGES * STAGE_D; e += KB_THREADS) sB[e] = 0.0;
    if (NCG > 0)
        for (int e = tid; e < NCG * GB2_MAX_P * GB2_MAX_P; e += KB_THREADS) {
            const int f = e / (GB2_MAX_P * GB2_MAX_P), q = e % (GB2_MAX_P * GB2_MAX_P);
            sBt[e] = q < T.cg_P[f] * T.cg_P[f] ? a.Btab[T.cg_Boff[f] + q] : 0.0;
        }

    // this thread's prefetch slots: slot s = chunk (tid & 31) of table row k = 8 s + (tid >> 5); row d (the squared norms) lands in stage row 4 KS
    const double* pf_src[NSLOT];
    unsigned pf_dst[NSLOT];
    bool pf_on[NSLOT];
#pragma unroll
    for (int s = 0; s < NSLOT; s++) {
        const int k = 8 * s + (tid >> 5), ch = tid & 31;
        pf_on[s] = k <= d;
        pf_src[s] = Fjt + (int64_t)(k <= d ? k : 0) * a.stride_j + ch * 2;
        pf_dst[s] = (unsigned)__cvta_generic_to_shared(sB + (k < d ? k : 4 * KS) * KB4_TS + ch * 2);
    }
    const int* pf_csrc = nullptr;
    unsigned pf_cdst = 0;
    if (NCG > 0 && tid < NCG * 16) {
        const int f = tid >> 4, ch = tid & 15;
        pf_csrc = a.Cj + (int64_t)T.cg_cat[f] * a.stride_j + ch * 4;
        pf_cdst = (unsigned)__cvta_generic_to_shared(sCj + f * KB_T + ch * 4);
    }
    auto prefetch = [&](int jt, int stage) {
        const int64_t j0 = (int64_t)jt * KB_T;
#pragma unroll
        for (int s = 0; s < NSLOT; s++)
            if (pf_on[s]) kb4_cp_async16(pf_dst[s] + stage * (STAGE_D * 8), pf_src[s] + j0);
        if (NCG > 0 && tid < NCG * 16) kb4_cp_async16(pf_cdst + stage * (NCG * KB_T * 4), pf_csrc + j0);
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

        // column ring: two tiles in flight before the first one is consumed
        prefetch(jt0, 0);
        asm volatile("cp.async.commit_group;\n" ::);
        if (jt0 + 1 < jt1) prefetch(jt0 + 1, 1);
        asm volatile("cp.async.commit_group;\n" ::);

        // row side of the item as DMMA A fragments: lane (g, t4) holds feature 4 ks + t4 of rows r0 + g and r0 + 8 + g, scaled by -2c
        double af[2][KS], si[2];
        int rowoff[NCG > 0 ? NCG : 1][2];
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
            asm volatile("cp.async.wait_group 1;\n" ::);
            __syncthreads();
            if (jt + 2 < jt1) prefetch(jt + 2, pf);
            asm volatile("cp.async.commit_group;\n" ::);
            pf = pf == KB4_STAGES - 1 ? 0 : pf + 1;
            const double* cB = sB + st * STAGE_D;
            const int* cC = sCj + st * (NCG > 0 ? NCG : 1) * KB_T;
            st = st == KB4_STAGES - 1 ? 0 : st + 1;
            const int64_t j0 = (int64_t)jt * KB_T;
            const bool interior = row_full && jt < n_full_cols && (!TRAIN || bi != jt);
            double* dtile = drow + j0;
            // two column halves of this warp's 16 x 32 patch: 8 live accumulators per lane each
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int cb = c0 + h * 16;
                double acc[2][2][2];
#pragma unroll
                for (int nj = 0; nj < 2; nj++) {
                    const double2 sj = *reinterpret_cast<const double2*>(cB + 4 * KS * KB4_TS + cb + nj * 8 + 2 * t4);
#pragma unroll
                    for (int mi = 0; mi < 2; mi++) {
                        acc[mi][nj][0] = fma(csc, sj.x, si[mi]);
                        acc[mi][nj][1] = fma(csc, sj.y, si[mi]);
                    }
                }
#pragma unroll
                for (int ks = 0; ks < KS; ks++) {
                    const double* pb = cB + (ks * 4 + t4) * KB4_TS + cb + g;
#pragma unroll
                    for (int nj = 0; nj < 2; nj++) {
                        const double b = pb[nj * 8];
                        kb_dmma(acc[0][nj][0], acc[0][nj][1], af[0][ks], b);
                        kb_dmma(acc[1][nj][0], acc[1][nj][1], af[1][ks], b);
                    }
                }
#pragma unroll
                for (int nj = 0; nj < 2; nj++) {
                    int cj[NCG > 0 ? NCG : 1][2];
#pragma unroll
                    for (int f = 0; f < NCG; f++) {
                        const int2 c2 = *reinterpret_cast<const int2*>(cC + f * KB_T + cb + nj * 8 + 2 * t4);
                        cj[f][0] = c2.x; cj[f][1] = c2.y;
                    }
#pragma unroll
                    for (int mi = 0; mi < 2; mi++)
#pragma unroll
                        for (int e = 0; e < 2; e++) {
                            double v = kb4_value<KIND>(kind_rt, acc[mi][nj][e], sTab, a);
#pragma unroll
                            for (int f = 0; f < NCG; f++) v *= sBt[rowoff[f][mi] + cj[f][e]];
                            acc[mi][nj][e] = v;
                        }
                }
                if (interior) {
#pragma unroll
                    for (int mi = 0; mi < 2; mi++)
#pragma unroll
                        for (int nj = 0; nj < 2; nj++)
                            *reinterpret_cast<double2*>(dtile + mi * ld8 + h * 16 + nj * 8) = make_double2(acc[mi][nj][0], acc[mi][nj][1]);
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
                            double v = acc[mi][nj][e];
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
        asm volatile("cp.async.wait_group 0;\n" ::);
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
