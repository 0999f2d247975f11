// Gradient of the marginal log-likelihood w.r.t. every hyper-parameter of the covariance.
//
// pm.find_MAP (gumbi/regression/pymc/GP.py:809-811) differentiates MvNormal.logp through Cholesky by reverse mode; the
// closed form of that gradient is
//     d logp / d theta = 1/2 sum_ij G_ij dK_ij/dtheta,     G = alpha alpha^T - K^-1,   alpha = K^-1 y = L^-T v.
// Steps (all on device, enqueued by gb2_mll_grad in gb2_abi.cu):
//   1. W = L^-T as rows: the predict solve (trsm_rec) applied to the identity.  Because row N of the augmented factor
//      holds v^T, column N of the solved identity is -alpha: extract_alpha_kernel takes it and zeroes the column.
//   2. S = W W^T (lower tiles, DMMA dgemm_nt) = K^-1.
//   3. mll_grad_kernel: one pass over the lower 64x64 tiles of G = alpha alpha^T - S, recomputing K_ij and its partial
//      derivatives from the per-point feature tables exactly as kbuild_kernel does; per-parameter partial sums are
//      reduced in registers -> warp shuffles -> one atomicAdd(double) per CTA and parameter.
// HBM-bound pass: reads 8 N(N+1)/2 bytes of S once.
#pragma once
#include "kbuild.cuh"

namespace gb2 {

// flat gradient layout (doubles) -- mirrored in include/gumbi_b200.h and gumbi_b200/_lib.py
constexpr int GR_LS = 0;
constexpr int GR_ETA = GB2_MAX_D;
constexpr int GR_C = GB2_MAX_D + 1;
constexpr int GR_TAU = GB2_MAX_D + 1 + GB2_MAX_LIN;
constexpr int GR_B = GB2_MAX_D + 2 + GB2_MAX_LIN;
constexpr int GR_TERM = GR_B + GB2_MAX_COREG * GB2_MAX_P * GB2_MAX_P;
constexpr int GR_SIGMA = GB2_MAX_TERMS * GR_TERM;
constexpr int GR_NOISE_B = GR_SIGMA + 1;
constexpr int GR_LEN = GR_NOISE_B + GB2_MAX_P * GB2_MAX_P;
static_assert(GR_LEN == GB2_GRAD_LEN, "gradient layout out of sync with include/gumbi_b200.h");

__global__ void set_identity_kernel(double* __restrict__ A, int64_t n, int64_t ld) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * (ld / 2)) return;
    const int64_t r = idx / (ld / 2), c = (idx % (ld / 2)) * 2;
    *reinterpret_cast<double2*>(A + r * ld + c) = make_double2(c == r ? 1.0 : 0.0, c + 1 == r ? 1.0 : 0.0);
}

// alpha[m] = -W[m, N];  W[m, N] = 0   (see step 1 above); rows m >= N of alpha are zeroed
__global__ void extract_alpha_kernel(double* __restrict__ W, int64_t ld, int64_t N, int64_t Np, double* __restrict__ alpha) {
    const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= Np) return;
    alpha[m] = m < N ? -W[m * ld + N] : 0.0;
    W[m * ld + N] = 0.0;
}

// value and "h(r) = -(dk/dr)/r" of the stationary kernels, so that dk/dls_k = h * (u_ik - u_jk)^2 / ls_k with u = x/ls
__device__ __forceinline__ void stationary_vh(int kind, double r2, double& k, double& hh) {
    if (kind == GB2_EXPQUAD) { k = exp(-0.5 * r2); hh = k; return; }
    const double r = sqrt(r2 + 1e-12);
    switch (kind) {
        case GB2_MATERN52: {
            const double s5 = 2.23606797749978969641;
            const double e = exp(-s5 * r);
            k = (1.0 + s5 * r + (5.0 / 3.0) * (r * r)) * e;
            hh = (5.0 / 3.0) * (1.0 + s5 * r) * e;
            return;
        }
        case GB2_MATERN32: {
            const double s3 = 1.73205080756887729353;
            const double e = exp(-s3 * r);
            k = (1.0 + s3 * r) * e;
            hh = 3.0 * e;
            return;
        }
        case GB2_MATERN12: k = exp(-r); hh = k / r; return;
        default: k = exp(-0.5 * r); hh = 0.5 * k / r; return;  // GB2_EXPONENTIAL
    }
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Block-wide sum of one value per thread, added to *dst by thread 0.  red: >= 8 doubles of shared scratch.
__device__ __forceinline__ void block_accumulate(double v, double* red, double* dst) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < KB_THREADS / 32; w++) s += red[w];
        if (s != 0.0) atomicAdd(dst, s);
    }
}

__global__ void __launch_bounds__(KB_THREADS)
mll_grad_kernel(KParams kp, const double* __restrict__ Btab, const double* __restrict__ F, const int* __restrict__ C,
                int64_t stride, int64_t n, const double* __restrict__ S, int64_t lds, const double* __restrict__ alpha,
                double* __restrict__ grad) {
    const int bi = blockIdx.y, bj = blockIdx.x;
    if (bj > bi) return;
    extern __shared__ __align__(16) unsigned char kb_smem[];
    const int nf = kp.n_feat, nc = kp.n_cat;
    double* sFi = reinterpret_cast<double*>(kb_smem);
    double* sFj = sFi + nf * KB_T;
    double* sBin = sFj + nf * KB_T;                       // GB2_MAX_P^2 bins (one Coregion factor at a time)
    double* red = sBin + GB2_MAX_P * GB2_MAX_P;           // 8
    int* sCi = reinterpret_cast<int*>(red + 8);
    int* sCj = sCi + (nc > 0 ? nc : 1) * KB_T;
    const int64_t i0 = (int64_t)bi * KB_T, j0 = (int64_t)bj * KB_T;
    for (int e = threadIdx.x; e < nf * KB_T; e += KB_THREADS) {
        const int r = e / KB_T, p = e % KB_T;
        sFi[e] = F[(int64_t)r * stride + i0 + p];
        sFj[e] = F[(int64_t)r * stride + j0 + p];
    }
    for (int e = threadIdx.x; e < nc * KB_T; e += KB_THREADS) {
        const int r = e / KB_T, p = e % KB_T;
        sCi[e] = C[(int64_t)r * stride + i0 + p];
        sCj[e] = C[(int64_t)r * stride + j0 + p];
    }
    const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
    // G_ij * symmetry weight (2 below the diagonal, 1 on it, 0 outside the live lower triangle)
    double g[4][4];
#pragma unroll
    for (int a = 0; a < 4; a++) {
        const int64_t gi = i0 + ty + 16 * a;
        const double ai = gi < n ? alpha[gi] : 0.0;
        const double2 s01 = *reinterpret_cast<const double2*>(S + gi * lds + j0 + tx * 4);
        const double2 s23 = *reinterpret_cast<const double2*>(S + gi * lds + j0 + tx * 4 + 2);
        const double sv[4] = {s01.x, s01.y, s23.x, s23.y};
#pragma unroll
        for (int b = 0; b < 4; b++) {
            const int64_t gj = j0 + tx * 4 + b;
            const bool live = gi < n && gj <= gi;
            const double aj = live ? alpha[gj] : 0.0;
            g[a][b] = live ? (gi == gj ? 1.0 : 2.0) * (ai * aj - sv[b]) : 0.0;
        }
    }
    __syncthreads();

    // ---- noise: K_ii += sigma^2 Bn[c,c] + jitter
    if (bi == bj) {
        double gs = 0.0;
        const bool hetero = kp.noise_cat >= 0;
        if (hetero) {
            for (int e = threadIdx.x; e < kp.noise_P * kp.noise_P; e += KB_THREADS) sBin[e] = 0.0;
            __syncthreads();
        }
#pragma unroll
        for (int a = 0; a < 4; a++)
#pragma unroll
            for (int b = 0; b < 4; b++)
                if (ty + 16 * a == tx * 4 + b && g[a][b] != 0.0) {
                    double bn = 1.0;
                    if (hetero) {
                        const int c = sCi[kp.noise_cat * KB_T + ty + 16 * a];
                        bn = __ldg(Btab + kp.noise_Boff + c * kp.noise_P + c);
                        atomicAdd(sBin + c * kp.noise_P + c, g[a][b] * kp.sigma2);
                    }
                    gs += g[a][b] * bn;
                }
        block_accumulate(gs, red, grad + GR_SIGMA);   // host multiplies by 2 sigma
        if (hetero) {
            __syncthreads();
            for (int e = threadIdx.x; e < kp.noise_P * kp.noise_P; e += KB_THREADS)
                if (sBin[e] != 0.0) atomicAdd(grad + GR_NOISE_B + e, sBin[e]);
            __syncthreads();
        }
    }

    for (int t = 0; t < kp.n_terms; t++) {
        const TermDev& T = kp.t[t];
        const double* fi = sFi + T.feat_off * KB_T;
        const double* fj = sFj + T.feat_off * KB_T;
        double* gt = grad + t * GR_TERM;
        // r^2 by the same expanded form as the forward build
        double acc[4][4];
#pragma unroll
        for (int a = 0; a < 4; a++)
#pragma unroll
            for (int b = 0; b < 4; b++) acc[a][b] = 0.0;
        for (int k = 0; k < T.d; k++) {
            double xa[4], xb[4];
#pragma unroll
            for (int a = 0; a < 4; a++) xa[a] = fi[k * KB_T + ty + 16 * a];
#pragma unroll
            for (int b = 0; b < 4; b++) xb[b] = fj[k * KB_T + tx * 4 + b];
#pragma unroll
            for (int a = 0; a < 4; a++)
#pragma unroll
                for (int b = 0; b < 4; b++) acc[a][b] = fma(xa[a], xb[b], acc[a][b]);
        }
        double kv[4][4], hv[4][4], lin[4][4], cb[4][4];
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
                    const double r2 = fmax(fma(-2.0, acc[a][b], sa[a] + sb[b]), 0.0);
                    stationary_vh(T.kind, r2, kv[a][b], hv[a][b]);
                    lin[a][b] = 0.0;
                    cb[a][b] = 1.0;
                }
        }
        for (int l = 0; l < T.n_lin; l++) {
            const int row = (T.d + 1 + l) * KB_T;
#pragma unroll
            for (int a = 0; a < 4; a++)
#pragma unroll
                for (int b = 0; b < 4; b++) lin[a][b] = fma(fi[row + ty + 16 * a], fj[row + tx * 4 + b], lin[a][b]);
        }
        for (int f = 0; f < T.n_coreg; f++) {
            const int* ci = sCi + T.cg_cat[f] * KB_T;
            const int* cj = sCj + T.cg_cat[f] * KB_T;
            const double* B = Btab + T.cg_Boff[f];
#pragma unroll
            for (int a = 0; a < 4; a++)
#pragma unroll
                for (int b = 0; b < 4; b++) cb[a][b] *= __ldg(B + ci[ty + 16 * a] * T.cg_P[f] + cj[tx * 4 + b]);
        }
        // ---- eta, tau
        {
            double ge = 0.0, gta = 0.0;
#pragma unroll
            for (int a = 0; a < 4; a++)
#pragma unroll
                for (int b = 0; b < 4; b++) {
                    const double w = g[a][b] * cb[a][b];
                    ge = fma(w, kv[a][b], ge);
                    gta = fma(w, lin[a][b], gta);
                }
            block_accumulate(ge, red, gt + GR_ETA);            // host multiplies by 2 eta
            if (T.n_lin > 0) block_accumulate(gta, red, gt + GR_TAU);
        }
        // ---- lengthscales: sum w eta^2 h (u_ik - u_jk)^2   (host divides by ls_k)
        for (int k = 0; k < T.d; k++) {
            double xa[4], xb[4], gl = 0.0;
#pragma unroll
            for (int a = 0; a < 4; a++) xa[a] = fi[k * KB_T + ty + 16 * a];
#pragma unroll
            for (int b = 0; b < 4; b++) xb[b] = fj[k * KB_T + tx * 4 + b];
#pragma unroll
            for (int a = 0; a < 4; a++)
#pragma unroll
                for (int b = 0; b < 4; b++) {
                    const double df = xa[a] - xb[b];
                    gl = fma(g[a][b] * cb[a][b] * hv[a][b], df * df, gl);
                }
            block_accumulate(gl * T.eta2, red, gt + GR_LS + k);
        }
        // ---- linear offsets: d/dc_l [tau (xi-c)(xj-c)] = -tau ((xi-c) + (xj-c))
        for (int l = 0; l < T.n_lin; l++) {
            const int row = (T.d + 1 + l) * KB_T;
            double gc = 0.0;
#pragma unroll
            for (int a = 0; a < 4; a++)
#pragma unroll
                for (int b = 0; b < 4; b++) gc = fma(g[a][b] * cb[a][b], fi[row + ty + 16 * a] + fj[row + tx * 4 + b], gc);
            block_accumulate(-T.tau * gc, red, gt + GR_C + l);
        }
        // ---- Coregion tables: d/dB_f[p,q] = sum_{ci=p,cj=q} G_ij base_ij prod_{f' != f} B_f'
        for (int f = 0; f < T.n_coreg; f++) {
            const int P = T.cg_P[f];
            const int* ci = sCi + T.cg_cat[f] * KB_T;
            const int* cj = sCj + T.cg_cat[f] * KB_T;
            // fast path: the tile's row levels and column levels are uniform (stacked-by-output data): one bin pair
            const int ci0 = ci[0], cj0 = cj[0];
            bool uni = true;
            for (int e = threadIdx.x; e < KB_T; e += KB_THREADS) uni = uni && ci[e] == ci0 && cj[e] == cj0;
            uni = __syncthreads_and(uni);
            if (!uni) {
                for (int e = threadIdx.x; e < P * P; e += KB_THREADS) sBin[e] = 0.0;
                __syncthreads();
            }
            double gb_lo = 0.0, gb_di = 0.0;  // below-diagonal entries (feed both (p,q) and (q,p)), diagonal entries
#pragma unroll
            for (int a = 0; a < 4; a++)
#pragma unroll
                for (int b = 0; b < 4; b++) {
                    if (g[a][b] == 0.0) continue;
                    double rest = 1.0;
                    for (int f2 = 0; f2 < T.n_coreg; f2++)
                        if (f2 != f)
                            rest *= __ldg(Btab + T.cg_Boff[f2] + sCi[T.cg_cat[f2] * KB_T + ty + 16 * a] * T.cg_P[f2] +
                                          sCj[T.cg_cat[f2] * KB_T + tx * 4 + b]);
                    const bool diag = (i0 + ty + 16 * a) == (j0 + tx * 4 + b);
                    // g carries weight 2 below the diagonal: split it over the two mirrored bins
                    const double v = g[a][b] * (diag ? 1.0 : 0.5) * rest * fma(T.eta2, kv[a][b], T.tau * lin[a][b]);
                    if (uni) {
                        if (diag) gb_di += v; else gb_lo += v;
                    } else {
                        const int p = ci[ty + 16 * a], q = cj[tx * 4 + b];
                        atomicAdd(sBin + p * P + q, v);
                        if (!diag) atomicAdd(sBin + q * P + p, v);
                    }
                }
            if (uni) {
                block_accumulate(gb_lo + gb_di, red, gt + GR_B + f * GB2_MAX_P * GB2_MAX_P + ci0 * P + cj0);
                block_accumulate(gb_lo, red, gt + GR_B + f * GB2_MAX_P * GB2_MAX_P + cj0 * P + ci0);
            } else {
                __syncthreads();
                for (int e = threadIdx.x; e < P * P; e += KB_THREADS)
                    if (sBin[e] != 0.0) atomicAdd(gt + GR_B + f * GB2_MAX_P * GB2_MAX_P + e, sBin[e]);
                __syncthreads();
            }
        }
    }
}

inline size_t mll_grad_smem_bytes(const KParams& kp) {
    return kbuild_smem_bytes(kp) + (size_t)(GB2_MAX_P * GB2_MAX_P + 8) * sizeof(double);
}

}  // namespace gb2
