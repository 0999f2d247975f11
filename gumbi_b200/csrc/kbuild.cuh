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
