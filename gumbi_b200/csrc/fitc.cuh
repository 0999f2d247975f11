// Sparse FITC approximation on the device: pm.gp.MarginalSparse(approx="FITC") as gumbi builds it for
// sparse=True (gumbi/regression/pymc/GP.py:571-578, :585-602), restated from PyMC's MarginalApprox
// (_build_marginal_likelihood_loglik / _build_conditional):
//   Luu = chol(Kuu + jitter I),  A = Luu^-1 Kuf,  Lambda = clip(diag(Kff) - colsum(A*A), 0) + sigma^2,
//   L_B = chol(I + A Lambda^-1 A^T),  c = L_B^-1 A Lambda^-1 y
//   logp = -N/2 log 2pi - 1/2 sum log Lambda - sum log diag(L_B) - 1/2 (y^T Lambda^-1 y - c^T c)
//   predict: As = Luu^-1 Kus, mu = (L_B^-1 As)^T c, var = kss - colsum(As*As) + colsum((L_B^-1 As)^2) [+ sigma^2]
//
// Everything O(N m^2) runs through the kernels of the exact path, used on two small inner systems:
//   * the inducing system (handle `fitc_u`, training set = Xu, noise 0): its factorisation is Luu, and an exact-GP "prediction"
//     at the TRAINING points X leaves A^T = K(X,Xu) Luu^-T in its solve panel and returns diag(Kff) - colsum(A*A) as the
//     "variance" (K*-build kernel, recursive DMMA TRSM, posterior_reduce_kernel);
//   * the B system (handle `fitc_b`): S = [A Lambda^-1/2 ; y^T Lambda^-1/2] is written K-contiguous by the kernel below, one
//     DMMA GEMM S S^T gives [A Lambda^-1 A^T, A Lambda^-1 y] in the augmented layout of the exact path (row m = right-hand side),
//     and the blocked Cholesky turns it into L_B with c = L_B^-1 A Lambda^-1 y in row m -- the same trick that gives v = L^-1 y.
#pragma once
#include "predict.cuh"

namespace gb2 {

// lam[i] = max(var[i], 0) + sigma2 where var = diag(Kff) - colsum(A*A) from the inducing system;
// scal[0] = sum_i log lam[i], scal[1] = sum_i y[i]^2 / lam[i]   (one block, fixed summation order)
__global__ void __launch_bounds__(1024)
fitc_lambda_kernel(const double* __restrict__ var, const double* __restrict__ y, int64_t N, double sigma2, double* __restrict__ lam,
                   double* __restrict__ scal) {
    __shared__ double s0[32], s1[32];
    double a = 0.0, b = 0.0;
    for (int64_t i = threadIdx.x; i < N; i += blockDim.x) {
        const double l = fmax(var[i], 0.0) + sigma2;
        lam[i] = l;
        a += log(l);
        b += y[i] * y[i] / l;
    }
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
    }
    if ((threadIdx.x & 31) == 0) { s0[threadIdx.x >> 5] = a; s1[threadIdx.x >> 5] = b; }
    __syncthreads();
    if (threadIdx.x < 32) {
        a = s0[threadIdx.x]; b = s1[threadIdx.x];
        for (int o = 16; o > 0; o >>= 1) {
            a += __shfl_xor_sync(0xffffffffu, a, o);
            b += __shfl_xor_sync(0xffffffffu, b, o);
        }
        if (threadIdx.x == 0) { scal[0] = a; scal[1] = b; }
    }
}

// St (rows_b x Nk, row stride Nk):  St[j][i] = At[i][j] / sqrt(lam[i])  (j < m),  St[m][i] = y[i] / sqrt(lam[i]),  0 for the
// other rows and for i >= N.  32x32 tiles through shared memory: reads of At and writes of St are both coalesced.
__global__ void __launch_bounds__(256)
fitc_scale_transpose_kernel(const double* __restrict__ At, int64_t ldt, const double* __restrict__ lam, const double* __restrict__ y,
                            int64_t N, int64_t m, double* __restrict__ St, int64_t Nk) {
    __shared__ double tile[32][33];
    const int64_t i0 = (int64_t)blockIdx.x * 32, j0 = (int64_t)blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int r = ty; r < 32; r += 8) {           // r: point within the tile, tx: column j within the tile
        const int64_t i = i0 + r, j = j0 + tx;
        double v = 0.0;
        if (i < N) {
            const double w = rsqrt(lam[i]);
            if (j < m) v = At[i * ldt + j] * w;
            else if (j == m) v = y[i] * w;
        }
        tile[r][tx] = v;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) St[(j0 + r) * Nk + i0 + tx] = tile[tx][r];
}

// after the S S^T product: A[i][i] += 1 for i < m (B = I + A Lambda^-1 A^T); unit diagonal on the augmented row and the padding
__global__ void fitc_fix_diag_kernel(double* __restrict__ A, int64_t ld, int64_t m, int64_t Np) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Np) return;
    if (i < m) A[i * ld + i] += 1.0;
    else A[i * ld + i] = 1.0;
}

// One warp per prediction point: mean = sum_j Ct[p][j] c[j],  var = var1[p] + sum_j Ct[p][j]^2 + add,  Ct = As^T L_B^-T
__global__ void __launch_bounds__(256)
fitc_reduce_kernel(const double* __restrict__ Ct, int64_t ld, const double* __restrict__ c, int64_t m, int64_t M,
                   const double* __restrict__ var1, double add, double* __restrict__ mean, double* __restrict__ var) {
    const int64_t p = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (p >= M) return;
    const int lane = threadIdx.x & 31;
    const double* row = Ct + p * ld;
    double mu = 0.0, ss = 0.0;
    for (int64_t j = lane; j < m; j += 32) {
        const double a = row[j];
        mu = fma(a, c[j], mu);
        ss = fma(a, a, ss);
    }
    for (int o = 16; o > 0; o >>= 1) {
        mu += __shfl_xor_sync(0xffffffffu, mu, o);
        ss += __shfl_xor_sync(0xffffffffu, ss, o);
    }
    if (lane == 0) { mean[p] = mu; var[p] = (var1[p] + ss) + add; }
}

}  // namespace gb2
