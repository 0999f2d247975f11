// Frozen copy of the round-2 (call t) persistent K-build kernel with renamed symbols: the comparison arm of tools/micro_kbuild.cu only.
// K-build v4: persistent strip kernel for the single-term models (pymc/GP.py:410,462,561,569):
//   K_ij = eta^2 k(|u_i - u_j|) * prod_f B_f[c_f(i), c_f(j)]        u = x / ls,  f <= 2 Coregion factors
// with the augmentation of the training matrix (noise + jitter on the diagonal, y^T in row N, identity padding) exactly as
// kbuild_dmma_kernel<TRAIN> writes it.  Additive / Linear models stay on kbuild_dmma_kernel.
//
// Why a new kernel (round-1 ncu of kbuild_strip_kernel, profiles/r01_ncu_kbuild_tf32_summary.txt + r01h source hot spots): 19 % of
// the launch was tail (one CTA per strip, triangular work), 20 % of the samples sat on the per-tile synchronous load of the
// column norms and the barrier behind it, and the per-entry code needed ~24 fp64 issue slots (the pipe that bounds this kernel
// on B200: DMMA and DFMA share the 64 lanes/clk/SM fp64 pipe).  Here:
//   * persistent CTAs (grid = SMs x resident CTAs) pull (row tile, strip of column tiles) items off an atomic counter -- no tail
//     beyond one item, no launch-shape dependence on the triangle;
//   * column-side data (features, squared norms, Coregion levels) arrive through a 3-stage cp.async ring, nothing synchronous
//     inside the tile loop; the row side of an item lives in registers as ready-made DMMA A fragments;
//   * fp64 slots per entry: the squared distance is  s_i + s_j  (one DADD, accumulator initialisation)  +  DMMA over d real
//     features with the row side pre-scaled by -2 (PyMC's own expanded form, Stationary.square_dist), no augmented k-step;
//     exp() uses a 2048-entry 2^(j/2048) table in shared memory and a cubic (7 fp64 operations, was 12), its range test
//     runs on the integer pipe.  ExpQuad: 1 + d + 7 slots (16 at d = 8, was 24); Matern-5/2: 1 + d + 5 (sqrt) + 7 + 3 = 24 (was ~34).
// Accuracy: table exact to 0.5 ulp, |reduced argument| <= ln2/4096 so the cubic truncates at 2^-58 relative; the one-step
// argument reduction carries |x| * 2^-54 -- the same size as the rounding of x itself.  Entrywise gate vs the oracle: 5e-12.
#pragma once
#include "../gumbi_b200/csrc/kbuild.cuh"

namespace gb2 {

constexpr int KB4O_TAB = 2048;       // entries of the exp table: 2^(j/2048)
constexpr int KB4O_REP = 1;          // copies of every entry.  Measured (profiles/r02g_micro_kbuild.log): a 256-entry table with 16 copies (one
                                    // per lane of a half-warp: no bank conflicts on the lookup, quartic instead of cubic) is 8-13 % SLOWER than
                                    // this single copy with its 2-3-way conflicts -- the extra fp64 operation costs more than the conflicts
constexpr int KB4O_TAB_LOG2 = 11;
constexpr int KB4O_STAGES = 3;
constexpr int KB4O_TS = 68;          // shared row stride (doubles) of a staged column tile: conflict-free 4x8 DMMA B fragments
constexpr int KB4O_MAXCG = 2;

__device__ double g_exp2_tab2k_r02t[KB4O_TAB];   // 2^(j/2048), filled by the host at gb2_create

struct KB4OArgs {
    const double* Fi; int64_t stride_i; int64_t n_i;     // row side: feature table of the term (row 0 = first scaled coordinate)
    const double* Fj; int64_t stride_j; int64_t n_j;     // column side
    const int* Ci; const int* Cj;                        // category tables (row f = Coregion factor f of this term), or nullptr
    const double* Btab;
    const double* y; double* out; int64_t ld;
    int n_row_tiles, n_col_tiles, strip;                 // strip = column tiles per work item
    int own_stride, own_rank, compact;
    int* ctr;                                            // [0] next item, [1] CTAs finished (self-resetting)
};

// eta^2 * exp(-z * zs) for z >= ~0, through  n = round(-z * zs * 2048 / ln 2):  tab = eta^2 * 2^(j/2048).
//   cA = -zs * 2048 / ln2,  cR = ln2 / (2048 zs)  ->  rr = z + n cR = -(reduced argument) / zs,  and with q(rr) the cubic of
//   exp(-zs rr) - 1:  k1 = -zs, k2 = zs^2 / 2, k3 = -zs^3 / 6.
template <int ZS2>   // ZS2 = 2 * zs: 1 (ExpQuad, z = r^2, exp(-z/2)) or 2 (Matern family, z = w, exp(-z))
__device__ __forceinline__ double kb4o_exp(double z, const double* __restrict__ tab) {
    constexpr double zs = 0.5 * ZS2;
    constexpr double LN2 = 0.693147180559945309417232121458176568;
    constexpr double LOG2E_TAB = (double)KB4O_TAB / LN2;    // 2048 / ln 2   (constant-folded by the host compiler, correctly rounded)
    constexpr double LN2_TAB = LN2 / (double)KB4O_TAB;      // ln 2 / 2048
    constexpr double MAGIC = 6755399441055744.0;            // 1.5 * 2^52
    const double t = fma(z, -zs * LOG2E_TAB, MAGIC);
    const int n = __double2loint(t);                         // round(-zs z 2048 / ln2) <= 0
    const double kf = t - MAGIC;
    const double rr = fma(kf, LN2_TAB / zs, z);             // |rr| <= ln2 / (4096 zs)
    const double q1 = fma(rr, -zs * zs * zs / 6.0, 0.5 * zs * zs);
    const double q2 = fma(q1, rr, -zs);
    const double m = rr * q2;                                // exp(-zs rr) - 1
    const double T = tab[(n & (KB4O_TAB - 1)) * KB4O_REP];
    const double res = fma(T, m, T);
    // 2^(n >> 11) by exponent arithmetic.  Arguments with zs z >= 693 (exp < 2^-1000 ~ 1e-301) return an exact 0: decided on the HIGH
    // WORD OF z (integer pipe; z >= 0 orders like its bit pattern, a rounding-negative z has the sign bit set and compares below),
    // because for huge scaled distances (z > ~1e6) the low word of t -- n -- wraps around and must not be consulted.  A result whose
    // exponent field would underflow (tiny eta^2 on top of a tiny exp) is flushed to 0 as well.
    constexpr int HI_ZMAX = ZS2 == 1 ? 0x4095A800 /* 1386.0 */ : 0x4085A800 /* 693.0 */;
    const int hi = __double2hiint(res) + ((n >> KB4O_TAB_LOG2) << 20);
    const bool tiny = __double2hiint(z) >= HI_ZMAX || hi < 0x00100000;
    return __hiloint2double(tiny ? 0 : hi, tiny ? 0 : __double2loint(res));
}

// sqrt(a) for a normal positive a: MUFU.RSQ64H seed (2^-22) + one third-order correction, 5 fp64 operations, residual ~2^-67
__device__ __forceinline__ double kb4o_sqrt(double a) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
    const double t = a * y;
    const double e = fma(-t, y, 1.0);
    const double p = fma(0.375, e, 0.5);
    const double te = t * e;
    return fma(te, p, t);
}

// Per kind: the accumulator holds  z = c r^2 (+ c 1e-12)  with c = kb4o_scale(kind)  (row side pre-scaled by -2c, norms by c)
__host__ __device__ inline double kb4o_scale(int kind) {
    switch (kind) {
        case GB2_EXPQUAD: return 1.0;
        case GB2_MATERN52: return 5.0;     // w = sqrt(5) r
        case GB2_MATERN32: return 3.0;     // w = sqrt(3) r
        case GB2_MATERN12: return 1.0;     // w = r
        default: return 0.25;              // Exponential: exp(-r/2), w = r/2
    }
}

template <int KIND>
__device__ __forceinline__ double kb4o_value(int kind_rt, double z, double zmin, const double* __restrict__ tab) {
    const int kind = KIND >= 0 ? KIND : kind_rt;
    // clip(r^2, 0, inf) of Stationary.square_dist, on the integer pipe (doubles order like their bit patterns as signed 64-bit
    // integers when the right-hand side is >= 0): the expanded form can come out negative by ~1e-16 |u|^2 (duplicated points, the
    // diagonal) -- harmless for exp at ordinary scales, but NaN under the Matern square root once it exceeds the 1e-12 epsilon
    if (kind == GB2_EXPQUAD) return kb4o_exp<1>(__double2hiint(z) < 0 ? 0.0 : z, tab);
    z = __double_as_longlong(z) < __double_as_longlong(zmin) ? zmin : z;   // zmin = c * 1e-12: clip(r^2, 0) + 1e-12 (euclidean_dist)
    const double w = kb4o_sqrt(z);
    const double e = kb4o_exp<2>(w, tab);
    if (kind == GB2_MATERN52) return e * fma(fma(1.0 / 3.0, w, 1.0), w, 1.0);   // 1 + w + w^2/3 = 1 + sqrt5 r + 5/3 r^2
    if (kind == GB2_MATERN32) return e * (1.0 + w);
    return e;                                         // Matern12, Exponential
}

// KS = DMMA k-steps (d <= 4 KS), NCG = Coregion factors of the term
template <bool TRAIN, int KIND, int KS, int NCG, int OCC>
__global__ void __launch_bounds__(KB_THREADS, OCC)
kbuild_persist_r02t_kernel(KParams kp, KB4OArgs a) {
    extern __shared__ __align__(16) unsigned char kb_smem[];
    double* sTab = reinterpret_cast<double*>(kb_smem);                         // [2048] x KB4O_REP
    double* sB = sTab + KB4O_TAB * KB4O_REP;                                                // [STAGES][4 KS][TS]
    double* sS = sB + KB4O_STAGES * 4 * KS * KB4O_TS;                             // [STAGES][64]  column squared norms (raw)
    double* sBt = sS + KB4O_STAGES * KB_T;                                       // [NCG][P*P <= 64] Coregion tables
    int* sCj = reinterpret_cast<int*>(sBt + (NCG > 0 ? NCG : 1) * GB2_MAX_P * GB2_MAX_P);   // [STAGES][NCG][64]
    __shared__ int s_item;

    const TermDev& T = kp.t[0];
    const int d = T.d;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t4 = lane & 3;
    const int r0 = (warp >> 1) * 16, c0 = (warp & 1) * 32;
    const int kind_rt = T.kind;
    const double csc = kb4o_scale(KIND >= 0 ? KIND : kind_rt);
    const double ceps = (KIND >= 0 ? KIND : kind_rt) == GB2_EXPQUAD ? 0.0 : csc * 1e-12;
    const double* Fit = a.Fi + (int64_t)T.feat_off * a.stride_i;
    const double* Fjt = a.Fj + (int64_t)T.feat_off * a.stride_j;

    // one-time per CTA: exp table scaled by eta^2, zero rows of the staged tiles (features d .. 4 KS - 1), Coregion tables
    for (int e = tid; e < KB4O_TAB * KB4O_REP; e += KB_THREADS) sTab[e] = T.eta2 * g_exp2_tab2k_r02t[e / KB4O_REP];
    const double* tabl = sTab + (lane & (KB4O_REP - 1));      // this lane's copy of the table
    for (int e = tid; e < KB4O_STAGES * 4 * KS * KB4O_TS; e += KB_THREADS) sB[e] = 0.0;
    if (NCG > 0)
        for (int e = tid; e < NCG * GB2_MAX_P * GB2_MAX_P; e += KB_THREADS) {
            const int f = e / (GB2_MAX_P * GB2_MAX_P), q = e % (GB2_MAX_P * GB2_MAX_P);
            sBt[e] = q < T.cg_P[f] * T.cg_P[f] ? a.Btab[T.cg_Boff[f] + q] : 0.0;
        }

    auto prefetch = [&](int jt, int stage) {
        const int64_t j0 = (int64_t)jt * KB_T;
        double* dB = sB + stage * 4 * KS * KB4O_TS;
        for (int c = tid; c < (d + 1) * 32; c += KB_THREADS) {       // d feature rows + the squared norms, 32 chunks of 16 bytes each
            const int k = c >> 5, ch = c & 31;
            double* dst = k < d ? dB + k * KB4O_TS + ch * 2 : sS + stage * KB_T + ch * 2;
            kb_cp_async16(dst, Fjt + (int64_t)k * a.stride_j + j0 + ch * 2);
        }
        if (NCG > 0 && tid < NCG * 16) {
            const int f = tid >> 4, ch = tid & 15;
            kb_cp_async16(sCj + (stage * NCG + f) * KB_T + ch * 4, a.Cj + (int64_t)T.cg_cat[f] * a.stride_j + j0 + ch * 4);
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

        for (int jt = jt0; jt < jt1; jt++) {
            const int stage = (jt - jt0) % KB4O_STAGES;
            asm volatile("cp.async.wait_group 1;\n" ::);
            __syncthreads();
            if (jt + 2 < jt1) prefetch(jt + 2, (jt + 2 - jt0) % KB4O_STAGES);
            asm volatile("cp.async.commit_group;\n" ::);
            const double* cB = sB + stage * 4 * KS * KB4O_TS;
            const double* cS = sS + stage * KB_T;
            double acc[2][4][2];
#pragma unroll
            for (int ni = 0; ni < 4; ni++) {
                const double2 sj = *reinterpret_cast<const double2*>(cS + c0 + ni * 8 + 2 * t4);
#pragma unroll
                for (int mi = 0; mi < 2; mi++) {
                    acc[mi][ni][0] = fma(csc, sj.x, si[mi]);
                    acc[mi][ni][1] = fma(csc, sj.y, si[mi]);
                }
            }
#pragma unroll
            for (int ks = 0; ks < KS; ks++) {
                const double* pb = cB + (ks * 4 + t4) * KB4O_TS + c0 + g;
#pragma unroll
                for (int ni = 0; ni < 4; ni++) {
                    const double b = pb[ni * 8];
                    kb_dmma(acc[0][ni][0], acc[0][ni][1], af[0][ks], b);
                    kb_dmma(acc[1][ni][0], acc[1][ni][1], af[1][ks], b);
                }
            }
            const int64_t j0 = (int64_t)jt * KB_T;
            const bool interior = (!TRAIN || bi != jt) && i0 + KB_T <= a.n_i && j0 + KB_T <= a.n_j;
#pragma unroll
            for (int ni = 0; ni < 4; ni++) {
                int cj[NCG > 0 ? NCG : 1][2];
#pragma unroll
                for (int f = 0; f < NCG; f++) {
                    const int2 c2 = *reinterpret_cast<const int2*>(sCj + (stage * NCG + f) * KB_T + c0 + ni * 8 + 2 * t4);
                    cj[f][0] = c2.x; cj[f][1] = c2.y;
                }
#pragma unroll
                for (int mi = 0; mi < 2; mi++)
#pragma unroll
                    for (int e = 0; e < 2; e++) {
                        double v = kb4o_value<KIND>(kind_rt, acc[mi][ni][e], ceps, tabl);
#pragma unroll
                        for (int f = 0; f < NCG; f++) v *= sBt[rowoff[f][mi] + cj[f][e]];
                        acc[mi][ni][e] = v;
                    }
            }
            if (interior) {
#pragma unroll
                for (int mi = 0; mi < 2; mi++) {
                    double* dst = a.out + kb_out_row(i0 + r0 + mi * 8 + g, a.own_stride, a.own_rank, TRAIN ? a.compact : 0) * a.ld + j0 + c0 + 2 * t4;
#pragma unroll
                    for (int ni = 0; ni < 4; ni++) *reinterpret_cast<double2*>(dst + ni * 8) = make_double2(acc[mi][ni][0], acc[mi][ni][1]);
                }
                continue;
            }
#pragma unroll
            for (int mi = 0; mi < 2; mi++) {
                const int64_t gi = i0 + r0 + mi * 8 + g;
#pragma unroll
                for (int ni = 0; ni < 4; ni++) {
                    double o[2];
#pragma unroll
                    for (int e = 0; e < 2; e++) {
                        const int64_t gj = j0 + c0 + ni * 8 + 2 * t4 + e;
                        double v = acc[mi][ni][e];
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
                    *reinterpret_cast<double2*>(a.out + kb_out_row(gi, a.own_stride, a.own_rank, TRAIN ? a.compact : 0) * a.ld + j0 + c0 + ni * 8 + 2 * t4) =
                        make_double2(o[0], o[1]);
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
constexpr size_t kb4o_smem_bytes() {
    return (size_t)(KB4O_TAB * KB4O_REP + KB4O_STAGES * 4 * KS * KB4O_TS + KB4O_STAGES * KB_T + (NCG > 0 ? NCG : 1) * GB2_MAX_P * GB2_MAX_P) * sizeof(double) +
           (size_t)KB4O_STAGES * (NCG > 0 ? NCG : 1) * KB_T * sizeof(int);
}

// models the persistent kernel covers: one term, no Linear part, <= 2 Coregion factors, any stationary kind
inline bool kb4o_eligible(const KParams& kp, bool train, int compact) {
    return kp.n_terms == 1 && kp.t[0].n_lin == 0 && kp.t[0].n_coreg <= KB4O_MAXCG && kp.t[0].d >= 1 && kp.t[0].d <= 16 && (train || !compact);
}

constexpr int KB4O_OCC = 4;   // resident CTAs per SM the register allocation is bounded for (tools/micro_kbuild.cu times 2, 3, 4)

template <bool TRAIN, int KIND, int KS, int NCG, int OCC = KB4O_OCC>
inline void kb4o_launch_one(cudaStream_t s, int n_sm, const KParams& kp, const KB4OArgs& a) {
    static bool configured = false;   // per instantiation; cudaFuncSetAttribute is idempotent, the flag only saves the call
    if (!configured) {
        cudaFuncSetAttribute(kbuild_persist_r02t_kernel<TRAIN, KIND, KS, NCG, OCC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kb4o_smem_bytes<KS, NCG>());
        configured = true;
    }
    kbuild_persist_r02t_kernel<TRAIN, KIND, KS, NCG, OCC><<<n_sm * OCC, KB_THREADS, kb4o_smem_bytes<KS, NCG>(), s>>>(kp, a);
}

template <bool TRAIN, int KIND>
inline void kb4o_launch_kind(cudaStream_t s, int n_sm, const KParams& kp, const KB4OArgs& a) {
    const int ks = (kp.t[0].d + 3) / 4, ncg = kp.t[0].n_coreg;
#define GB2_KB4O(KS_, NCG_) if (ks == KS_ && ncg == NCG_) return kb4o_launch_one<TRAIN, KIND, KS_, NCG_>(s, n_sm, kp, a)
    GB2_KB4O(1, 0); GB2_KB4O(2, 0); GB2_KB4O(3, 0); GB2_KB4O(4, 0);
    GB2_KB4O(1, 1); GB2_KB4O(2, 1); GB2_KB4O(3, 1); GB2_KB4O(4, 1);
    GB2_KB4O(1, 2); GB2_KB4O(2, 2); GB2_KB4O(3, 2); GB2_KB4O(4, 2);
#undef GB2_KB4O
}

template <bool TRAIN>
inline void kbuild_persist_r02t_launch(cudaStream_t s, int n_sm, const KParams& kp, const double* Btab, const double* Fi, const int* Ci, int64_t stride_i,
                                  int64_t n_i, const double* Fj, const int* Cj, int64_t stride_j, int64_t n_j, int n_row_tiles, int n_col_tiles,
                                  const double* y, double* out, int64_t ld, int own_stride, int own_rank, int compact, int* ctr) {
    KB4OArgs a{};
    a.Fi = Fi; a.stride_i = stride_i; a.n_i = n_i; a.Fj = Fj; a.stride_j = stride_j; a.n_j = n_j; a.Ci = Ci; a.Cj = Cj; a.Btab = Btab;
    a.y = y; a.out = out; a.ld = ld; a.n_row_tiles = n_row_tiles; a.n_col_tiles = n_col_tiles;
    a.own_stride = own_stride; a.own_rank = own_rank; a.compact = compact; a.ctr = ctr;
    // strip length: long strips amortise the per-item row set-up, short ones balance small problems (>= ~16 items per CTA)
    const double tiles = TRAIN ? 0.5 * n_row_tiles * (double)(n_row_tiles + 1) / (own_stride > 1 ? own_stride : 1) : (double)n_row_tiles * n_col_tiles;
    int strip = (int)(tiles / (16.0 * n_sm * KB4O_OCC));
    a.strip = strip < 1 ? 1 : (strip > 8 ? 8 : strip);
    if (kp.t[0].kind == GB2_EXPQUAD) kb4o_launch_kind<TRAIN, GB2_EXPQUAD>(s, n_sm, kp, a);
    else if (kp.t[0].kind == GB2_MATERN52) kb4o_launch_kind<TRAIN, GB2_MATERN52>(s, n_sm, kp, a);
    else kb4o_launch_kind<TRAIN, -1>(s, n_sm, kp, a);
}

}  // namespace gb2
