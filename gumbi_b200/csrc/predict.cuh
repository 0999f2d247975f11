// Posterior mean / variance over a block of prediction points (Marginal._build_conditional):
//   A = L^-1 K(X,X*),  mu = A^T v,  var = k** - colsum(A*A) (+ noise)
// Stored transposed: At (Mp x Np) row-major, one row per prediction point, so that every operand of the solve is
// K-contiguous for dgemm.cuh.  The solve  At <- At L^-T  is a recursive blocked TRSM: leaves multiply by the
// inverted 128x128 diagonal blocks (computed during the factorisation), everything else is one large DMMA GEMM.
#pragma once
#include <algorithm>
#include "cholesky.cuh"
#include "kbuild.cuh"

namespace gb2 {

// tri: the right-hand side is the identity (At starts as I, rows = columns), whose solution W = L^-T is upper triangular as
// rows: row block m is zero left of column block m, so every step only touches the rows above its last column block.
inline void trsm_rec(cudaStream_t s, const double* L, int64_t ld, const double* Dinv, double* At, int64_t ldt,
                     int64_t Mp, int c0, int c1, int& launches, bool tri = false) {
    if (c1 - c0 == 1) {
        double* X = At + (int64_t)c0 * TILE;
        const int64_t rows = tri ? std::min<int64_t>(Mp, (int64_t)(c0 + 1) * TILE) : Mp;
        dgemm_nt_launch<64, 128, GM_SET>(s, X, ldt, Dinv + (int64_t)c0 * TILE * TILE, TILE, X, ldt, rows, TILE, TILE, 0, 0, 0);
        launches++;
        return;
    }
    const int mid = c0 + (c1 - c0 + 1) / 2;
    trsm_rec(s, L, ld, Dinv, At, ldt, Mp, c0, mid, launches, tri);
    const int64_t rows = tri ? std::min<int64_t>(Mp, (int64_t)mid * TILE) : Mp;
    dgemm_sub_launch(s, At + (int64_t)c0 * TILE, ldt, L + (int64_t)mid * TILE * ld + (int64_t)c0 * TILE, ld,
                                     At + (int64_t)mid * TILE, ldt, rows, (int64_t)(c1 - mid) * TILE, (mid - c0) * TILE, 0, 0, 0);
    launches++;
    trsm_rec(s, L, ld, Dinv, At, ldt, Mp, mid, c1, launches, tri);
}

// GB2_TF32 variant: sub-solves of up to `leaf` column blocks stay in fp64 (DMMA); every larger off-diagonal product of the
// recursion runs as a tcgen05 split-TF32 GEMM on the tf32 hi/lo copies of L (mLhi/mLlo) and of the already solved columns of
// At (mAthi/mAtlo, refreshed after each leaf).  n_total = number of column blocks of the whole solve (no split needed for the
// last columns, nothing to their right consumes them).
inline void trsm_rec_tf32(gb2_handle* h, cudaStream_t s, int64_t Mp, int c0, int c1, int n_total, int leaf, int& launches) {
    const int64_t Np = h->Np;
    if (c1 - c0 <= leaf) {
        trsm_rec(s, h->dA, Np, h->dDinv, h->dAt, Np, Mp, c0, c1, launches);
        if (c1 < n_total) {
            const int64_t cols = (int64_t)(c1 - c0) * TILE;
            tc::split_tf32_kernel<<<(unsigned)((Mp * cols / 2 + 255) / 256), 256, 0, s>>>(h->dAt + (int64_t)c0 * TILE, Np, Mp, cols,
                                                                                         h->dAthi + (int64_t)c0 * TILE,
                                                                                         h->dAtlo + (int64_t)c0 * TILE, Np);
            launches++;
        }
        return;
    }
    const int half = (c1 - c0 + 1) / 2;
    const int mid = c0 + (half + leaf - 1) / leaf * leaf;
    trsm_rec_tf32(h, s, Mp, c0, mid, n_total, leaf, launches);
    tc::GemmArgs g{};
    g.C = h->dAt; g.ldc = Np;
    g.n_bi = (int)(Mp / TILE); g.n_bj = c1 - mid;
    g.rb_first = 0; g.rb_stride = 1; g.rb_local_first = -1; g.cblk0 = mid; g.lower = 0;
    g.a_k0 = c0 * TILE; g.b_row0 = mid * TILE; g.b_k0 = c0 * TILE;
    tc::gemm_tf32x3_launch(s, h->n_sm, h->mAthi, h->mAtlo, h->mLhi, h->mLlo, g, (mid - c0) * TILE, launches);
    trsm_rec_tf32(h, s, Mp, mid, c1, n_total, leaf, launches);
}

// At[:, n .. width) = 0: the solve leaves garbage in the augmented (y) column and the padding columns of the last block, which
// the diag reduction skips by its loop bound but a dense At At^T product would pick up.
__global__ void mask_columns_kernel(double* __restrict__ At, int64_t ldt, int64_t rows, int64_t n, int64_t width) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    for (int64_t c = n; c < width; c++) At[r * ldt + c] = 0.0;
}

// cov[m][m] += noise.diag(x*_m)   (pred_noise=True, full-covariance prediction)
__global__ void add_noise_diag_kernel(KParams kp, const double* __restrict__ Btab, const double* __restrict__ Fs, const int* __restrict__ Cs,
                                      int64_t stride_s, int64_t M, double* __restrict__ cov, int64_t ldc) {
    const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    double kss, nz;
    point_diag(kp, Btab, Fs, Cs, stride_s, m, kss, nz);
    cov[m * ldc + m] += nz;
}

// One warp per prediction point: mean = sum_i At[m,i] v[i], var = kss - sum_i At[m,i]^2 (+ noise diag).
__global__ void __launch_bounds__(256)
posterior_reduce_kernel(KParams kp, const double* __restrict__ Btab, const double* __restrict__ Fs,
                        const int* __restrict__ Cs, int64_t stride_s, const double* __restrict__ At, int64_t ldt,
                        const double* __restrict__ v, int64_t n, int64_t M, int pred_noise,
                        double* __restrict__ mean, double* __restrict__ var) {
    const int64_t m = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (m >= M) return;
    const int lane = threadIdx.x & 31;
    const double* row = At + m * ldt;
    double mu = 0.0, ss = 0.0;
    for (int64_t i = 2 * lane; i < n; i += 64) {  // n is even-aligned by construction of ldt; tail handled below
        if (i + 1 < n) {
            const double2 a = *reinterpret_cast<const double2*>(row + i);
            const double2 w = *reinterpret_cast<const double2*>(v + i);
            mu = fma(a.x, w.x, mu); ss = fma(a.x, a.x, ss);
            mu = fma(a.y, w.y, mu); ss = fma(a.y, a.y, ss);
        } else {
            const double a = row[i];
            mu = fma(a, v[i], mu); ss = fma(a, a, ss);
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        mu += __shfl_xor_sync(0xffffffffu, mu, o);
        ss += __shfl_xor_sync(0xffffffffu, ss, o);
    }
    if (lane == 0) {
        double kss, nz;
        point_diag(kp, Btab, Fs, Cs, stride_s, m, kss, nz);
        mean[m] = mu;
        var[m] = (kss - ss) + (pred_noise ? nz : 0.0);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Storage-sharded prediction: L is distributed by row blocks, so the solve  At <- At L^-T  is distributed by COLUMN blocks of
// At (rank r holds At[:, j] for j % G == r, contiguously) and runs right-looking, one 128-column block per step:
//   owner(j):  X_j = At[:, j] inv(L_jj)^T   -- the same fused kernel as the factor panel: stored locally and pushed into slot
//              j % ring_slots of every rank's ring (Mp x 128, ld 128), counters bumped
//   others:    wait for the counter
//   all:       At[:, i] -= X_j L[i, j]^T  for the owned column blocks i > j   (L[i, j] is local: row block i, column block j)
// then every rank reduces its own columns (mean and sum-of-squares partials) and the partials are all-gathered and summed.
// All ranks pass the SAME prediction points; all ranks return the full result.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
posterior_partial_kernel(const double* __restrict__ At, int64_t ldt, const double* __restrict__ v, int64_t n, int64_t M, int64_t ncols_loc,
                         int G, int me, double* __restrict__ part, int64_t pstride) {
    const int64_t m = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (m >= M) return;
    const int lane = threadIdx.x & 31;
    const double* row = At + m * ldt;
    double mu = 0.0, ss = 0.0;
    for (int64_t lc = lane; lc < ncols_loc; lc += 32) {
        const int64_t i = ((lc / TILE) * G + me) * TILE + lc % TILE;   // global training index of local column lc
        if (i < n) {
            const double a = row[lc];
            mu = fma(a, v[i], mu);
            ss = fma(a, a, ss);
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        mu += __shfl_xor_sync(0xffffffffu, mu, o);
        ss += __shfl_xor_sync(0xffffffffu, ss, o);
    }
    if (lane == 0) { part[m] = mu; part[pstride + m] = ss; }
}

__global__ void posterior_final_kernel(KParams kp, const double* __restrict__ Btab, const double* __restrict__ Fs, const int* __restrict__ Cs,
                                       int64_t stride_s, const double* __restrict__ all, int G, int64_t pstride, int64_t M, int pred_noise,
                                       double* __restrict__ mean, double* __restrict__ var) {
    const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    double mu = 0.0, ss = 0.0;
    for (int r = 0; r < G; r++) { mu += all[(int64_t)r * 2 * pstride + m]; ss += all[(int64_t)r * 2 * pstride + pstride + m]; }
    double kss, nz;
    point_diag(kp, Btab, Fs, Cs, stride_s, m, kss, nz);
    mean[m] = mu;
    var[m] = (kss - ss) + (pred_noise ? nz : 0.0);
}

}  // namespace gb2
