// Covariance-matrix build kernels.
//
//   prep_features : X (n, D_in) row-major  ->  feature table F (n_feat, stride) and category table C.
//   kbuild_kernel<true>  : lower-triangle 64x64 tiles of the augmented training matrix
//                          A = [ K(X,X) + (sigma^2 B_noise[p,p] + jitter) I   . ]
//                              [ y^T                                         1 ]   (+ identity padding)
//   kbuild_kernel<false> : At[m, i] = K(x*_m, x_i)   (rows = prediction points, cols = training points)
//
// Arithmetic follows pm.gp.cov.* as called from gumbi/regression/pymc/GP.py:410,453,462,561,569:
//   r2 = clip(|xi|^2 + |xj|^2 - 2 xi.xj, 0)  with x already divided by ls   (Stationary.square_dist)
//   ExpQuad exp(-r2/2); Matern use r = sqrt(r2 + 1e-12)                      (Stationary.euclidean_dist)
// HBM-bound by design: every entry is written exactly once with 32-byte-per-thread row-contiguous stores;
// the per-point features of a tile (<= 2 x 64 x n_feat doubles) are staged once in shared memory.
#pragma once
#include "gb2_internal.cuh"

namespace gb2 {

__global__ void prep_features(const double* __restrict__ X, int64_t n, int64_t stride, PrepParams pp,
                              double* __restrict__ F, int* __restrict__ C, int* __restrict__ bad) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= stride) return;
    const bool live = i < n;
    const double* x = X + i * pp.D_in;
    for (int t = 0; t < pp.n_terms; t++) {
        const int off = pp.feat_off[t], d = pp.d[t];
        double sq = 0.0;
        for (int k = 0; k < d; k++) {
            double v = live ? x[pp.cont_idx[t][k]] * pp.inv_ls[t][k] : 0.0;
            F[(int64_t)(off + k) * stride + i] = v;
            sq += v * v;
        }
        F[(int64_t)(off + d) * stride + i] = sq;
        for (int l = 0; l < pp.n_lin[t]; l++)
            F[(int64_t)(off + d + 1 + l) * stride + i] = live ? x[pp.lin_idx[t][l]] - pp.c[t][l] : 0.0;
    }
    for (int f = 0; f < pp.n_cat; f++) {
        int ci = 0;
        if (live) {
            double v = x[pp.cat_col[f]];
            ci = (int)v;  // Coregion: pt.cast(X, "int32") truncates
            if (!(v == v) || ci < 0 || ci >= pp.cat_P[f]) { atomicExch(bad, 1); ci = 0; }
        }
        C[(int64_t)f * stride + i] = ci;
    }
}

__device__ __forceinline__ double stationary(int kind, double r2) {
    if (kind == GB2_EXPQUAD) return exp(-0.5 * r2);
    const double r = sqrt(r2 + 1e-12);
    switch (kind) {
        case GB2_MATERN52: {
            const double s5 = 2.23606797749978969641;
            return (1.0 + s5 * r + (5.0 / 3.0) * (r * r)) * exp(-s5 * r);
        }
        case GB2_MATERN32: {
            const double s3 = 1.73205080756887729353;
            return (1.0 + s3 * r) * exp(-s3 * r);
        }
        case GB2_MATERN12: return exp(-r);
        default: return exp(-0.5 * r);  // GB2_EXPONENTIAL
    }
}

constexpr int KB_T = 64;        // tile edge
constexpr int KB_THREADS = 256; // 16 x 16 threads, 4 x 4 entries each

// TRAIN: Fi/Ci and Fj/Cj are the same tables (stride_i == stride_j == Np), lower tiles only, augmentation applied.
// !TRAIN: i indexes prediction points (rows, n_i = M), j indexes training points (cols, n_j = N).
template <bool TRAIN>
__global__ void __launch_bounds__(KB_THREADS)
kbuild_kernel(KParams kp, const double* __restrict__ Btab,
              const double* __restrict__ Fi, const int* __restrict__ Ci, int64_t stride_i, int64_t n_i,
              const double* __restrict__ Fj, const int* __restrict__ Cj, int64_t stride_j, int64_t n_j,
              const double* __restrict__ y, double* __restrict__ out, int64_t ld, int own_stride, int own_rank) {
    const int bi = blockIdx.y, bj = blockIdx.x;
    if (TRAIN && bj > bi) return;
    // multi-GPU row-block sharding: this rank builds only the 128-row blocks it owns (block-cyclic)
    if (TRAIN && own_stride > 1 && ((bi * KB_T) / TILE) % own_stride != own_rank) return;
    extern __shared__ __align__(16) unsigned char kb_smem[];
    const int nf = kp.n_feat, nc = kp.n_cat;
    double* sFi = reinterpret_cast<double*>(kb_smem);
    double* sFj = sFi + nf * KB_T;
    int* sCi = reinterpret_cast<int*>(sFj + nf * KB_T);
    int* sCj = sCi + nc * KB_T;
    const int64_t i0 = (int64_t)bi * KB_T, j0 = (int64_t)bj * KB_T;
    for (int e = threadIdx.x; e < nf * KB_T; e += KB_THREADS) {
        int r = e / KB_T, p = e % KB_T;
        sFi[e] = Fi[(int64_t)r * stride_i + i0 + p];
        sFj[e] = Fj[(int64_t)r * stride_j + j0 + p];
    }
    for (int e = threadIdx.x; e < nc * KB_T; e += KB_THREADS) {
        int r = e / KB_T, p = e % KB_T;
        sCi[e] = Ci[(int64_t)r * stride_i + i0 + p];
        sCj[e] = Cj[(int64_t)r * stride_j + j0 + p];
    }
    __syncthreads();
    const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
    double val[4][4];
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
        for (int b = 0; b < 4; b++) val[a][b] = 0.0;

    for (int t = 0; t < kp.n_terms; t++) {
        const TermDev& T = kp.t[t];
        const double* fi = sFi + T.feat_off * KB_T;
        const double* fj = sFj + T.feat_off * KB_T;
        double acc[4][4];
#pragma unroll
        for (int a = 0; a < 4; a++)
#pragma unroll
            for (int b = 0; b < 4; b++) acc[a][b] = 0.0;
        for (int k = 0; k < T.d; k++) {
            double xa[4];
#pragma unroll
            for (int a = 0; a < 4; a++) xa[a] = fi[k * KB_T + ty + 16 * a];
            const double2 b01 = *reinterpret_cast<const double2*>(fj + k * KB_T + tx * 4);
            const double2 b23 = *reinterpret_cast<const double2*>(fj + k * KB_T + tx * 4 + 2);
            const double xb[4] = {b01.x, b01.y, b23.x, b23.y};
#pragma unroll
            for (int a = 0; a < 4; a++)
#pragma unroll
                for (int b = 0; b < 4; b++) acc[a][b] = fma(xa[a], xb[b], acc[a][b]);
        }
        {
            double sa[4], sb[4];
#pragma unroll
            for (int a = 0; a < 4; a++) sa[a] = fi[T.d * KB_T + ty + 16 * a];
#pragma unroll
            for (int b = 0; b < 4; b++) sb[b] = fj[T.d * KB_T + tx * 4 + b];
#pragma unroll
            for (int a = 0; a < 4; a++)
#pragma unroll
                for (int b = 0; b < 4; b++) {
                    double r2 = fmax(fma(-2.0, acc[a][b], sa[a] + sb[b]), 0.0);
                    acc[a][b] = T.eta2 * stationary(T.kind, r2);
                }
        }
        if (T.n_lin > 0) {
            double lin[4][4];
#pragma unroll
            for (int a = 0; a < 4; a++)
#pragma unroll
                for (int b = 0; b < 4; b++) lin[a][b] = 0.0;
            for (int l = 0; l < T.n_lin; l++) {
                const int row = (T.d + 1 + l) * KB_T;
                double xa[4], xb[4];
#pragma unroll
                for (int a = 0; a < 4; a++) xa[a] = fi[row + ty + 16 * a];
#pragma unroll
                for (int b = 0; b < 4; b++) xb[b] = fj[row + tx * 4 + b];
#pragma unroll
                for (int a = 0; a < 4; a++)
#pragma unroll
                    for (int b = 0; b < 4; b++) lin[a][b] = fma(xa[a], xb[b], lin[a][b]);
            }
#pragma unroll
            for (int a = 0; a < 4; a++)
#pragma unroll
                for (int b = 0; b < 4; b++) acc[a][b] = fma(T.tau, lin[a][b], acc[a][b]);
        }
        for (int f = 0; f < T.n_coreg; f++) {
            const int* ci = sCi + T.cg_cat[f] * KB_T;
            const int* cj = sCj + T.cg_cat[f] * KB_T;
            const double* B = Btab + T.cg_Boff[f];
            const int P = T.cg_P[f];
#pragma unroll
            for (int a = 0; a < 4; a++) {
                const int ca = ci[ty + 16 * a] * P;
#pragma unroll
                for (int b = 0; b < 4; b++) acc[a][b] *= __ldg(B + ca + cj[tx * 4 + b]);
            }
        }
#pragma unroll
        for (int a = 0; a < 4; a++)
#pragma unroll
            for (int b = 0; b < 4; b++) val[a][b] += acc[a][b];
    }

#pragma unroll
    for (int a = 0; a < 4; a++) {
        const int64_t gi = i0 + ty + 16 * a;
        double o[4];
#pragma unroll
        for (int b = 0; b < 4; b++) {
            const int64_t gj = j0 + tx * 4 + b;
            double v = val[a][b];
            if (TRAIN) {
                if (gi < n_i && gj < n_j) {
                    if (gi == gj) {
                        double nz = kp.sigma2;
                        if (kp.noise_cat >= 0) {
                            const int c = sCi[kp.noise_cat * KB_T + ty + 16 * a];
                            nz *= __ldg(Btab + kp.noise_Boff + c * kp.noise_P + c);
                        }
                        v += nz + kp.jitter;
                    }
                } else if (gi == n_i && gj < n_j) {
                    v = y[gj];            // augmented row: forward substitution of y rides along the factorisation
                } else {
                    v = (gi == gj) ? 1.0 : 0.0;
                }
            } else {
                if (gi >= n_i || gj >= n_j) v = 0.0;
            }
            o[b] = v;
        }
        double* dst = out + gi * ld + j0 + tx * 4;
        *reinterpret_cast<double2*>(dst) = make_double2(o[0], o[1]);
        *reinterpret_cast<double2*>(dst + 2) = make_double2(o[2], o[3]);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// K-build v2: the Gram part of the squared distance runs on the fp64 tensor pipe (DMMA m8n8k4) and the exponential is a
// table-driven fp64 exp (64-entry 2^(j/64) table in shared memory + degree-5 polynomial, 10 fp64 instructions instead of
// the ~25 of the library exp), so that the fp64 vector pipe -- which bounded v1 at 0.20 of the HBM roofline (ncu: fp64
// pipe 27 % active, issue 37 %, DRAM 13 %) -- has roughly a third of the work per entry.
//
// Per term the tile edges are staged as augmented feature rows (K-dim ka = round_up(d + 2, 4)):
//     row side   [ u_0 .. u_{d-1},  -|u|^2/2,  1,        0.. ]        u = x / ls
//     col side   [ u_0 .. u_{d-1},  1,        -|u|^2/2,  0.. ]
// so that one DMMA chain delivers  u_i.u_j - |u_i|^2/2 - |u_j|^2/2 = -r^2/2  (the expanded form PyMC uses, GP.py:410 ->
// Stationary.square_dist) directly as the exponent of ExpQuad.
// Tile 64 x 64, 8 warps; warp w owns rows (w>>1)*16..+16 and columns (w&1)*32..+32 = 2 x 4 DMMA blocks; a thread holds
// (row g, columns 2t, 2t+1) of each block and stores it as one 16-byte row-contiguous piece.
// ---------------------------------------------------------------------------------------------------------------
constexpr int KB2_TS = 68;   // shared row stride (doubles): 68 = 4 mod 16 makes the 8x4 / 4x8 DMMA fragment loads conflict-free

__device__ double g_exp2_tab[64];   // 2^(j/64), filled by the host at gb2_create

// exp(x) for x <= ~0; exactly 0 below x = -700 (exp < 1e-304).  |relative error| ~ 2e-16.
__device__ __forceinline__ double exp_tab(double x, const double* __restrict__ tab) {
    const double t = fma(x, 92.33248261689366, 6755399441055744.0);   // 64/ln2, 1.5*2^52: low word of t = round(64 x / ln 2)
    const int n = __double2loint(t);
    const double kf = t - 6755399441055744.0;
    double r = fma(kf, -0.01083042469326756, x);                       // ln2/64 split hi (32-bit mantissa) + lo
    r = fma(kf, -2.9815858269852933e-12, r);
    const double r2 = r * r;
    double q = fma(r, 1.0 / 120.0, 1.0 / 24.0);
    q = fma(q, r, 1.0 / 6.0);
    q = fma(q, r, 0.5);
    const double p = fma(q, r2, r);                                     // e^r - 1, |r| <= ln2/128
    const double T = tab[n & 63];
    const double res = fma(T, p, T);
    const double sc = __hiloint2double(__double2hiint(res) + ((n >> 6) << 20), __double2loint(res));
    return x < -700.0 ? 0.0 : sc;
}

// value of the stationary kernel from x = -r^2/2
__device__ __forceinline__ double stationary_x(int kind, double x, const double* __restrict__ tab) {
    if (kind == GB2_EXPQUAD) return exp_tab(fmin(x, 0.0), tab);
    const double r = sqrt(fmax(-2.0 * x, 0.0) + 1e-12);
    switch (kind) {
        case GB2_MATERN52: {
            const double s5 = 2.23606797749978969641;
            return (1.0 + s5 * r + (5.0 / 3.0) * (r * r)) * exp_tab(-s5 * r, tab);
        }
        case GB2_MATERN32: {
            const double s3 = 1.73205080756887729353;
            return (1.0 + s3 * r) * exp_tab(-s3 * r, tab);
        }
        case GB2_MATERN12: return exp_tab(-r, tab);
        default: return exp_tab(-0.5 * r, tab);  // GB2_EXPONENTIAL
    }
}

__device__ __forceinline__ void kb_dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__host__ __device__ inline int kb2_ka(int d) { return (d + 2 + 3) / 4 * 4; }

// Row of the output buffer that holds global row gi: identity, or -- storage-sharded multi-GPU mode -- the position inside the
// contiguous run of 128-row blocks this rank owns (block b -> local block (b - rank) / stride).
__device__ __forceinline__ int64_t kb_out_row(int64_t gi, int own_stride, int own_rank, int compact) {
    if (!compact) return gi;
    return ((gi / TILE - own_rank) / own_stride) * TILE + gi % TILE;
}

// KIND >= 0 with SIMPLE: compile-time specialisation for the common model (one term, that stationary kernel, no Linear, no
// Coregion) -- the per-entry code then has no kind switch and no term / factor loops.  KIND = -1: everything at run time.
template <bool TRAIN, int KIND, bool SIMPLE>
__global__ void __launch_bounds__(KB_THREADS)
kbuild_dmma_kernel(KParams kp, const double* __restrict__ Btab,
                   const double* __restrict__ Fi, const int* __restrict__ Ci, int64_t stride_i, int64_t n_i,
                   const double* __restrict__ Fj, const int* __restrict__ Cj, int64_t stride_j, int64_t n_j,
                   const double* __restrict__ y, double* __restrict__ out, int64_t ld, int own_stride, int own_rank, int compact) {
    const int bi = blockIdx.y, bj = blockIdx.x;
    if (TRAIN && bj > bi) return;
    if (TRAIN && own_stride > 1 && ((bi * KB_T) / TILE) % own_stride != own_rank) return;
    // K(X*, X) in the storage-sharded mode: this rank builds only the column blocks (training points) it owns, stored contiguously
    if (!TRAIN && compact && ((bj * KB_T) / TILE) % own_stride != own_rank) return;
    extern __shared__ __align__(16) unsigned char kb_smem[];
    // total augmented / linear rows over the terms
    int ka_tot = 0, nl_tot = 0;
    for (int t = 0; t < kp.n_terms; t++) { ka_tot += kb2_ka(kp.t[t].d); nl_tot += kp.t[t].n_lin; }
    const int nc = kp.n_cat;
    double* sA = reinterpret_cast<double*>(kb_smem);          // [ka_tot][TS] row side
    double* sB = sA + ka_tot * KB2_TS;                         // [ka_tot][TS] column side
    double* sLi = sB + ka_tot * KB2_TS;                        // [nl_tot][TS]
    double* sLj = sLi + nl_tot * KB2_TS;                       // [nl_tot][TS]
    double* sTab = sLj + nl_tot * KB2_TS;                      // [64]
    int* sCi = reinterpret_cast<int*>(sTab + 64);              // [nc][64]
    int* sCj = sCi + (nc > 0 ? nc : 1) * KB_T;
    const int64_t i0 = (int64_t)bi * KB_T, j0 = (int64_t)bj * KB_T;
    const int64_t jc0 = (!TRAIN && compact) ? kb_out_row(j0, own_stride, own_rank, 1) : j0;   // output column of the tile's first column
    if (threadIdx.x < 64) sTab[threadIdx.x] = g_exp2_tab[threadIdx.x];
    {
        int aoff = 0, loff = 0;
        for (int t = 0; t < kp.n_terms; t++) {
            const TermDev& T = kp.t[t];
            const int d = T.d, ka = kb2_ka(d);
            for (int e = threadIdx.x; e < ka * KB_T; e += KB_THREADS) {
                const int k = e / KB_T, p = e % KB_T;
                double a, b;
                if (k < d) {
                    a = Fi[(int64_t)(T.feat_off + k) * stride_i + i0 + p];
                    b = Fj[(int64_t)(T.feat_off + k) * stride_j + j0 + p];
                } else if (k == d) {
                    a = -0.5 * Fi[(int64_t)(T.feat_off + d) * stride_i + i0 + p];
                    b = 1.0;
                } else if (k == d + 1) {
                    a = 1.0;
                    b = -0.5 * Fj[(int64_t)(T.feat_off + d) * stride_j + j0 + p];
                } else {
                    a = b = 0.0;
                }
                sA[(aoff + k) * KB2_TS + p] = a;
                sB[(aoff + k) * KB2_TS + p] = b;
            }
            for (int e = threadIdx.x; e < T.n_lin * KB_T; e += KB_THREADS) {
                const int l = e / KB_T, p = e % KB_T;
                sLi[(loff + l) * KB2_TS + p] = Fi[(int64_t)(T.feat_off + d + 1 + l) * stride_i + i0 + p];
                sLj[(loff + l) * KB2_TS + p] = Fj[(int64_t)(T.feat_off + d + 1 + l) * stride_j + j0 + p];
            }
            aoff += ka; loff += T.n_lin;
        }
    }
    for (int e = threadIdx.x; e < nc * KB_T; e += KB_THREADS) {
        const int r = e / KB_T, p = e % KB_T;
        sCi[e] = Ci[(int64_t)r * stride_i + i0 + p];
        sCj[e] = Cj[(int64_t)r * stride_j + j0 + p];
    }
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t4 = lane & 3;
    const int r0 = (warp >> 1) * 16, c0 = (warp & 1) * 32;
    double val[2][4][2];
#pragma unroll
    for (int mi = 0; mi < 2; mi++)
#pragma unroll
        for (int ni = 0; ni < 4; ni++) val[mi][ni][0] = val[mi][ni][1] = 0.0;

    int aoff = 0, loff = 0;
    const int n_terms = SIMPLE ? 1 : kp.n_terms;
    for (int t = 0; t < n_terms; t++) {
        const TermDev& T = kp.t[t];
        const int ka = kb2_ka(T.d);
        const int kind = KIND >= 0 ? KIND : T.kind;
        double acc[2][4][2];
#pragma unroll
        for (int mi = 0; mi < 2; mi++)
#pragma unroll
            for (int ni = 0; ni < 4; ni++) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
        for (int kk = 0; kk < ka; kk += 4) {
            const double* pa = sA + (aoff + kk + t4) * KB2_TS + r0 + g;
            const double* pb = sB + (aoff + kk + t4) * KB2_TS + c0 + g;
            const double a0 = pa[0], a1 = pa[8];
            const double b0 = pb[0], b1 = pb[8], b2 = pb[16], b3 = pb[24];
            kb_dmma(acc[0][0][0], acc[0][0][1], a0, b0); kb_dmma(acc[0][1][0], acc[0][1][1], a0, b1);
            kb_dmma(acc[0][2][0], acc[0][2][1], a0, b2); kb_dmma(acc[0][3][0], acc[0][3][1], a0, b3);
            kb_dmma(acc[1][0][0], acc[1][0][1], a1, b0); kb_dmma(acc[1][1][0], acc[1][1][1], a1, b1);
            kb_dmma(acc[1][2][0], acc[1][2][1], a1, b2); kb_dmma(acc[1][3][0], acc[1][3][1], a1, b3);
        }
#pragma unroll
        for (int mi = 0; mi < 2; mi++)
#pragma unroll
            for (int ni = 0; ni < 4; ni++)
#pragma unroll
                for (int e = 0; e < 2; e++) {
                    double v = T.eta2 * stationary_x(kind, acc[mi][ni][e], sTab);
                    if (SIMPLE) { val[mi][ni][e] = v; continue; }
                    const int pr = r0 + mi * 8 + g, pc = c0 + ni * 8 + 2 * t4 + e;
                    if (T.n_lin > 0) {
                        double lin = 0.0;
                        for (int l = 0; l < T.n_lin; l++) lin = fma(sLi[(loff + l) * KB2_TS + pr], sLj[(loff + l) * KB2_TS + pc], lin);
                        v = fma(T.tau, lin, v);
                    }
                    for (int f = 0; f < T.n_coreg; f++)
                        v *= __ldg(Btab + T.cg_Boff[f] + sCi[T.cg_cat[f] * KB_T + pr] * T.cg_P[f] + sCj[T.cg_cat[f] * KB_T + pc]);
                    val[mi][ni][e] += v;
                }
        aoff += ka; loff += T.n_lin;
    }

    // interior tiles (no diagonal, no augmented row, no padding): plain stores
    if ((!TRAIN || bi != bj) && i0 + KB_T <= n_i && j0 + KB_T <= n_j) {
#pragma unroll
        for (int mi = 0; mi < 2; mi++) {
            double* dst = out + kb_out_row(i0 + r0 + mi * 8 + g, own_stride, own_rank, TRAIN ? compact : 0) * ld + jc0 + c0 + 2 * t4;
#pragma unroll
            for (int ni = 0; ni < 4; ni++) *reinterpret_cast<double2*>(dst + ni * 8) = make_double2(val[mi][ni][0], val[mi][ni][1]);
        }
        return;
    }
#pragma unroll
    for (int mi = 0; mi < 2; mi++) {
        const int pr = r0 + mi * 8 + g;
        const int64_t gi = i0 + pr;
#pragma unroll
        for (int ni = 0; ni < 4; ni++) {
            double o[2];
#pragma unroll
            for (int e = 0; e < 2; e++) {
                const int64_t gj = j0 + c0 + ni * 8 + 2 * t4 + e;
                double v = val[mi][ni][e];
                if (TRAIN) {
                    if (gi < n_i && gj < n_j) {
                        if (gi == gj) {
                            double nz = kp.sigma2;
                            if (kp.noise_cat >= 0) {
                                const int c = sCi[kp.noise_cat * KB_T + pr];
                                nz *= __ldg(Btab + kp.noise_Boff + c * kp.noise_P + c);
                            }
                            v += nz + kp.jitter;
                        }
                    } else if (gi == n_i && gj < n_j) {
                        v = y[gj];
                    } else {
                        v = (gi == gj) ? 1.0 : 0.0;
                    }
                } else {
                    if (gi >= n_i || gj >= n_j) v = 0.0;
                }
                o[e] = v;
            }
            *reinterpret_cast<double2*>(out + kb_out_row(gi, own_stride, own_rank, TRAIN ? compact : 0) * ld + jc0 + c0 + ni * 8 + 2 * t4) =
                make_double2(o[0], o[1]);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// K-build v3 (the common model: one stationary term, no Linear, no Coregion): a CTA keeps its 64 row points resident and walks
// a strip of up to KB3_JG column tiles; the column-side features of tile j+1 are prefetched with cp.async into the other
// half of a double buffer while tile j is computed, so the global-load latency that v2 paid once per 64x64 tile
// (ncu: 25 % of the samples in the tile prologue + barrier) is off the critical path.  Same DMMA Gram + table exp as v2, with
// eta^2 folded into the exp table and the range clamp reduced to one compare + select.
// ---------------------------------------------------------------------------------------------------------------
constexpr int KB3_JG = 8;

__device__ __forceinline__ void kb_cp_async16(void* smem_dst, const void* gmem_src) {
    unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gmem_src));
}

// eta^2 * exp(x), x <= ~0; tab = eta^2 * 2^(j/64).  x < -700 returns 0 (exp < 1e-304).
__device__ __forceinline__ double exp_tab_scaled(double x, const double* __restrict__ tab) {
    const double t = fma(x, 92.33248261689366, 6755399441055744.0);
    const int n = __double2loint(t);
    const double kf = t - 6755399441055744.0;
    double r = fma(kf, -0.01083042469326756, x);
    r = fma(kf, -2.9815858269852933e-12, r);
    const double r2 = r * r;
    double q = fma(r, 1.0 / 120.0, 1.0 / 24.0);
    q = fma(q, r, 1.0 / 6.0);
    q = fma(q, r, 0.5);
    const double p = fma(q, r2, r);
    const double T = tab[n & 63];
    const double res = fma(T, p, T);
    const double sc = __hiloint2double(__double2hiint(res) + ((n >> 6) << 20), __double2loint(res));
    return x < -700.0 ? 0.0 : sc;
}

// OCC = resident CTAs per SM the register allocation is bounded for (3: 78 registers, 4: 64 registers, a few bytes of spill)
template <bool TRAIN, int KIND, int OCC>
__global__ void __launch_bounds__(KB_THREADS, OCC)
kbuild_strip_kernel(KParams kp, const double* __restrict__ Fi, int64_t stride_i, int64_t n_i, const double* __restrict__ Fj,
                    int64_t stride_j, int64_t n_j, int n_col_tiles, const double* __restrict__ y, double* __restrict__ out, int64_t ld,
                    int own_stride, int own_rank, int compact) {
    const int bi = blockIdx.y;
    const int jt0 = blockIdx.x * KB3_JG;
    if (TRAIN && jt0 > bi) return;
    if (TRAIN && own_stride > 1 && ((bi * KB_T) / TILE) % own_stride != own_rank) return;
    int jt1 = jt0 + KB3_JG < n_col_tiles ? jt0 + KB3_JG : n_col_tiles;
    if (TRAIN && jt1 > bi + 1) jt1 = bi + 1;
    extern __shared__ __align__(16) unsigned char kb_smem[];
    const TermDev& T = kp.t[0];
    const int d = T.d, ka = kb2_ka(d);
    // (Tried and measured slower on B200, 0.117 vs 0.102 ms at C2: for d % 4 == 0, dropping the augmented DMMA k-step in favour of
    // two DADDs per entry and testing the exp range on the integer pipe.  The augmented form stays.)
    const int kdot = ka;
    double* sA = reinterpret_cast<double*>(kb_smem);      // [ka][TS]
    double* sB = sA + ka * KB2_TS;                         // [2][ka][TS]
    double* sTab = sB + 2 * ka * KB2_TS;                   // [64]
    const int tid = threadIdx.x;
    const int64_t i0 = (int64_t)bi * KB_T;
    const double* Fjt = Fj + (int64_t)T.feat_off * stride_j;

    auto prefetch = [&](int jt, int buf) {
        double* dst = sB + buf * ka * KB2_TS;
        const int64_t j0 = (int64_t)jt * KB_T;
        for (int c = tid; c < d * 32; c += KB_THREADS) {     // d rows x 32 chunks of 16 bytes
            const int k = c >> 5, ch = c & 31;
            kb_cp_async16(dst + k * KB2_TS + ch * 2, Fjt + (int64_t)k * stride_j + j0 + ch * 2);
        }
        if (tid < KB_T) dst[(d + 1) * KB2_TS + tid] = -0.5 * Fjt[(int64_t)d * stride_j + j0 + tid];
        asm volatile("cp.async.commit_group;\n" ::);
    };

    // one-time: exp table (scaled by eta^2), constant rows, row-side features
    if (tid < 64) sTab[tid] = T.eta2 * g_exp2_tab[tid];
    for (int e = tid; e < 2 * KB_T; e += KB_THREADS) {
        const int buf = e / KB_T, p = e % KB_T;
        sB[buf * ka * KB2_TS + d * KB2_TS + p] = 1.0;
        for (int k = d + 2; k < ka; k++) sB[buf * ka * KB2_TS + k * KB2_TS + p] = 0.0;
    }
    {
        const int p = tid & 63;
        for (int k = tid >> 6; k < ka; k += KB_THREADS / 64) {
            double a;
            if (k < d) a = Fi[(int64_t)(T.feat_off + k) * stride_i + i0 + p];
            else if (k == d) a = -0.5 * Fi[(int64_t)(T.feat_off + d) * stride_i + i0 + p];
            else a = (k == d + 1) ? 1.0 : 0.0;
            sA[k * KB2_TS + p] = a;
        }
    }
    prefetch(jt0, 0);

    const int warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t4 = lane & 3;
    const int r0 = (warp >> 1) * 16, c0 = (warp & 1) * 32;
    for (int jt = jt0; jt < jt1; jt++) {
        const int buf = (jt - jt0) & 1;
        asm volatile("cp.async.wait_group 0;\n" ::);
        __syncthreads();
        if (jt + 1 < jt1) prefetch(jt + 1, buf ^ 1);
        const double* cB = sB + buf * ka * KB2_TS;
        double acc[2][4][2];
#pragma unroll
        for (int mi = 0; mi < 2; mi++)
#pragma unroll
            for (int ni = 0; ni < 4; ni++) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
        for (int kk = 0; kk < kdot; kk += 4) {
            const double* pa = sA + (kk + t4) * KB2_TS + r0 + g;
            const double* pb = cB + (kk + t4) * KB2_TS + c0 + g;
            const double a0 = pa[0], a1 = pa[8];
            const double b0 = pb[0], b1 = pb[8], b2 = pb[16], b3 = pb[24];
            kb_dmma(acc[0][0][0], acc[0][0][1], a0, b0); kb_dmma(acc[0][1][0], acc[0][1][1], a0, b1);
            kb_dmma(acc[0][2][0], acc[0][2][1], a0, b2); kb_dmma(acc[0][3][0], acc[0][3][1], a0, b3);
            kb_dmma(acc[1][0][0], acc[1][0][1], a1, b0); kb_dmma(acc[1][1][0], acc[1][1][1], a1, b1);
            kb_dmma(acc[1][2][0], acc[1][2][1], a1, b2); kb_dmma(acc[1][3][0], acc[1][3][1], a1, b3);
        }
#pragma unroll
        for (int mi = 0; mi < 2; mi++)
#pragma unroll
            for (int ni = 0; ni < 4; ni++)
#pragma unroll
                for (int e = 0; e < 2; e++) {
                    const double x = acc[mi][ni][e];
                    if (KIND == GB2_EXPQUAD) {
                        acc[mi][ni][e] = exp_tab_scaled(x, sTab);
                    } else {  // GB2_MATERN52
                        const double r = sqrt(fmax(-2.0 * x, 0.0) + 1e-12);
                        const double s5 = 2.23606797749978969641;
                        acc[mi][ni][e] = (1.0 + s5 * r + (5.0 / 3.0) * (r * r)) * exp_tab_scaled(-s5 * r, sTab);
                    }
                }
        const int64_t j0 = (int64_t)jt * KB_T;
        if ((!TRAIN || bi != jt) && i0 + KB_T <= n_i && j0 + KB_T <= n_j) {
#pragma unroll
            for (int mi = 0; mi < 2; mi++) {
                double* dst = out + kb_out_row(i0 + r0 + mi * 8 + g, own_stride, own_rank, TRAIN ? compact : 0) * ld + j0 + c0 + 2 * t4;
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
                        if (gi < n_i && gj < n_j) {
                            if (gi == gj) v += kp.sigma2 + kp.jitter;     // SIMPLE: no noise Coregion (it needs an output column)
                        } else if (gi == n_i && gj < n_j) {
                            v = y[gj];
                        } else {
                            v = (gi == gj) ? 1.0 : 0.0;
                        }
                    } else {
                        if (gi >= n_i || gj >= n_j) v = 0.0;
                    }
                    o[e] = v;
                }
                *reinterpret_cast<double2*>(out + kb_out_row(gi, own_stride, own_rank, TRAIN ? compact : 0) * ld + j0 + c0 + ni * 8 + 2 * t4) =
                    make_double2(o[0], o[1]);
            }
        }
    }
}

inline size_t kbuild_strip_smem_bytes(const KParams& kp) { return (size_t)(3 * kb2_ka(kp.t[0].d) * KB2_TS + 64) * sizeof(double); }

// persistent kernel for the single-term models (kbuild_persist.cuh, included at the end of this header's users)
inline bool kb4_eligible(const KParams& kp, bool train, int compact);
template <bool TRAIN>
inline void kbuild_persist_launch(cudaStream_t s, int n_sm, const KParams& kp, const double* Btab, const double* Fi, const int* Ci, int64_t stride_i,
                                  int64_t n_i, const double* Fj, const int* Cj, int64_t stride_j, int64_t n_j, int n_row_tiles, int n_col_tiles,
                                  const double* y, double* out, int64_t ld, int own_stride, int own_rank, int compact, int* ctr);

// host-side dispatch over the specialisations
template <bool TRAIN>
inline void kbuild_dmma_launch(cudaStream_t s, dim3 grid, size_t smem, const KParams& kp, const double* Btab, const double* Fi, const int* Ci,
                               int64_t stride_i, int64_t n_i, const double* Fj, const int* Cj, int64_t stride_j, int64_t n_j, const double* y,
                               double* out, int64_t ld, int own_stride, int own_rank, int compact = 0, int occ = 4, int* ctr = nullptr, int n_sm = 0) {
    // single-term models without a Linear part: persistent strip kernel (kbuild_persist.cuh); ctr == nullptr keeps the round-1 kernels (ablation)
    if (ctr && kb4_eligible(kp, TRAIN, compact)) {
        kbuild_persist_launch<TRAIN>(s, n_sm, kp, Btab, Fi, Ci, stride_i, n_i, Fj, Cj, stride_j, n_j, (int)grid.y, (int)grid.x, y, out, ld, own_stride,
                                     own_rank, compact, ctr);
        return;
    }
    // (the strip kernel's prefetch pipeline walks consecutive column tiles; the column-sharded K* build uses the per-tile kernel)
    const bool simple = kp.n_terms == 1 && kp.t[0].n_lin == 0 && kp.t[0].n_coreg == 0 && kp.noise_cat < 0 && (TRAIN || !compact);
    if (simple && (kp.t[0].kind == GB2_EXPQUAD || kp.t[0].kind == GB2_MATERN52)) {
        // grid: x = strips of KB3_JG column tiles, y = row tiles
        dim3 sgrid((grid.x + KB3_JG - 1) / KB3_JG, grid.y);
        const size_t ssm = kbuild_strip_smem_bytes(kp);
#define GB2_STRIP(KIND, OCC)                                                                                                          \
    kbuild_strip_kernel<TRAIN, KIND, OCC><<<sgrid, KB_THREADS, ssm, s>>>(kp, Fi, stride_i, n_i, Fj, stride_j, n_j, (int)grid.x, y, out, ld, \
                                                                         own_stride, own_rank, compact)
        if (kp.t[0].kind == GB2_EXPQUAD) { if (occ == 3) GB2_STRIP(GB2_EXPQUAD, 3); else GB2_STRIP(GB2_EXPQUAD, 4); }
        else { if (occ == 3) GB2_STRIP(GB2_MATERN52, 3); else GB2_STRIP(GB2_MATERN52, 4); }
#undef GB2_STRIP
        return;
    }
#define GB2_KB_LAUNCH(KIND, SIMPLE)                                                                                             \
    kbuild_dmma_kernel<TRAIN, KIND, SIMPLE><<<grid, KB_THREADS, smem, s>>>(kp, Btab, Fi, Ci, stride_i, n_i, Fj, Cj, stride_j, n_j, y, out, ld, \
                                                                           own_stride, own_rank, compact)
    GB2_KB_LAUNCH(-1, false);
#undef GB2_KB_LAUNCH
}

template <bool TRAIN>
inline cudaError_t kbuild_dmma_configure() {
    cudaError_t e;
    if ((e = cudaFuncSetAttribute(kbuild_strip_kernel<TRAIN, GB2_EXPQUAD, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(kbuild_strip_kernel<TRAIN, GB2_MATERN52, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(kbuild_strip_kernel<TRAIN, GB2_EXPQUAD, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(kbuild_strip_kernel<TRAIN, GB2_MATERN52, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024)) != cudaSuccess) return e;
    return cudaFuncSetAttribute(kbuild_dmma_kernel<TRAIN, -1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
}

inline size_t kbuild_dmma_smem_bytes(const KParams& kp) {
    int ka_tot = 0, nl_tot = 0;
    for (int t = 0; t < kp.n_terms; t++) { ka_tot += kb2_ka(kp.t[t].d); nl_tot += kp.t[t].n_lin; }
    return (size_t)(2 * ka_tot + 2 * nl_tot) * KB2_TS * sizeof(double) + 64 * sizeof(double) +
           (size_t)2 * (kp.n_cat > 0 ? kp.n_cat : 1) * KB_T * sizeof(int);
}

inline size_t kbuild_smem_bytes(const KParams& kp) {
    return (size_t)2 * kp.n_feat * KB_T * sizeof(double) + (size_t)2 * (kp.n_cat > 0 ? kp.n_cat : 1) * KB_T * sizeof(int);
}

// kss[m] = cov_total.diag(x*_m), nz[m] = noise.diag(x*_m)   (Marginal._build_conditional, diag=True branch)
__device__ __forceinline__ void point_diag(const KParams& kp, const double* __restrict__ Btab,
                                           const double* __restrict__ F, const int* __restrict__ C,
                                           int64_t stride, int64_t m, double& kss, double& nz) {
    kss = 0.0;
    for (int t = 0; t < kp.n_terms; t++) {
        const TermDev& T = kp.t[t];
        double kc = T.eta2;  // Stationary.diag == 1
        if (T.n_lin > 0) {
            double s = 0.0;
            for (int l = 0; l < T.n_lin; l++) {
                double v = F[(int64_t)(T.feat_off + T.d + 1 + l) * stride + m];
                s = fma(v, v, s);
            }
            kc = fma(T.tau, s, kc);
        }
        for (int f = 0; f < T.n_coreg; f++) {
            int c = C[(int64_t)T.cg_cat[f] * stride + m];
            kc *= Btab[T.cg_Boff[f] + c * T.cg_P[f] + c];
        }
        kss += kc;
    }
    nz = kp.sigma2;
    if (kp.noise_cat >= 0) {
        int c = C[(int64_t)kp.noise_cat * stride + m];
        nz *= Btab[kp.noise_Boff + c * kp.noise_P + c];
    }
}

}  // namespace gb2
