// Blocked jittered Cholesky of the augmented training matrix (L = chol(K + Knoise + 1e-6 I)).
//
// Right-looking, block size 128, with a one-step look-ahead on a second (high-priority) stream:
//   step k:   [panel stream]  potrf_diag(k)  : 128x128 diagonal block factor + its inverse (one CTA, shared memory)
//             [panel stream]  L[i,k] = A[i,k] * inv(L_kk)^T  for all row blocks below (DMMA, in place)
//             [main stream]   A[:,k+1] -= L[:,k] L[k+1,k]^T   (next panel first -> releases step k+1 on the panel stream)
//             [main stream]   A[i,j]   -= L[i,k] L[j,k]^T      for the rest of the trailing lower triangle
// The trailing update is the N^3/3 tensor-core contraction (dgemm.cuh); the diagonal panel is the custom kernel below.
// Because row N of the augmented matrix holds y^T, the panel solves turn it into v^T = (L^-1 y)^T for free.
#pragma once
#include "dgemm.cuh"

namespace gb2 {

constexpr int PD_THREADS = 512;
constexpr int PD_LD = TILE + 1;  // padded shared row
constexpr size_t PD_SMEM = (size_t)TILE * PD_LD * sizeof(double) + TILE * sizeof(double);

// Factor the diagonal block starting at global index g0 (in place, lower), then invert it into Dinv (dense 128x128,
// strict upper triangle zero).  Columns with global index >= n_real (the y row and the identity padding) get pivot 1.
// A non-positive pivot records info = global column + 1 (first one wins) and is replaced by 1 so that the rest of the
// pipeline stays finite; the host turns info into LinAlgError.
__global__ void __launch_bounds__(PD_THREADS, 1)
potrf_diag_kernel(double* __restrict__ A, int64_t ld, int64_t g0, int64_t n_real, double* __restrict__ Dinv,
                  int* __restrict__ info) {
    extern __shared__ __align__(16) unsigned char pd_smem[];
    double* S = reinterpret_cast<double*>(pd_smem);
    double* col = S + TILE * PD_LD;
    const int tid = threadIdx.x;
    double* Ab = A + g0 * ld + g0;
    for (int e = tid; e < TILE * TILE; e += PD_THREADS) {
        int r = e >> 7, c = e & 127;
        S[r * PD_LD + c] = (c <= r) ? Ab[(int64_t)r * ld + c] : 0.0;
    }
    __syncthreads();

    for (int j = 0; j < TILE; j++) {
        double p = S[j * PD_LD + j];
        if (g0 + j >= n_real) {
            p = 1.0;
        } else if (!(p > 0.0)) {
            if (tid == 0) atomicCAS(info, 0, (int)(g0 + j + 1));
            p = 1.0;
        }
        const double ljj = sqrt(p);
        const double inv = 1.0 / ljj;
        __syncthreads();  // everyone has read the pivot before it is overwritten
        if (tid == 0) S[j * PD_LD + j] = ljj;
        for (int r = j + 1 + tid; r < TILE; r += PD_THREADS) S[r * PD_LD + j] *= inv;
        __syncthreads();
        const int n = TILE - 1 - j;
        for (int e = tid; e < n * n; e += PD_THREADS) {
            const int rr = e / n, cc = e - rr * n;
            if (cc <= rr) {
                const int r = j + 1 + rr, c = j + 1 + cc;
                S[r * PD_LD + c] = fma(-S[r * PD_LD + j], S[c * PD_LD + j], S[r * PD_LD + c]);
            }
        }
        __syncthreads();
    }
    for (int e = tid; e < TILE * TILE; e += PD_THREADS) {
        int r = e >> 7, c = e & 127;
        if (c <= r) Ab[(int64_t)r * ld + c] = S[r * PD_LD + c];
    }
    __syncthreads();

    // In-place inversion of the lower-triangular block, column by column from the right (LAPACK dtrti2 order):
    //   X[j][j] = 1/L[j][j];   X[r][j] = -X[j][j] * sum_{k=j+1..r} X[r][k] L[k][j]
    // 4 threads share a row r and split the dot product; shuffle-reduced.
    for (int j = TILE - 1; j >= 0; j--) {
        for (int r = j + tid; r < TILE; r += PD_THREADS) col[r] = S[r * PD_LD + j];
        __syncthreads();
        const double xjj = 1.0 / col[j];
        const int r = j + 1 + (tid >> 2), part = tid & 3;
        double s = 0.0;
        if (r < TILE) {
            for (int k = j + 1 + part; k <= r; k += 4) s = fma(S[r * PD_LD + k], col[k], s);
        }
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        if (r < TILE && part == 0) S[r * PD_LD + j] = -s * xjj;
        if (tid == 0) S[j * PD_LD + j] = xjj;
        __syncthreads();
    }
    for (int e = tid; e < TILE * TILE; e += PD_THREADS) {
        int r = e >> 7, c = e & 127;
        Dinv[e] = (c <= r) ? S[r * PD_LD + c] : 0.0;
    }
}

// sum_{i<n} log A[i][i]  and  sum_{i<n} A[n][i]^2  (log-determinant half and |v|^2) -> scal[0], scal[1]
__global__ void mll_terms_kernel(const double* __restrict__ A, int64_t ld, int64_t n, double* __restrict__ scal) {
    __shared__ double s0[32], s1[32];
    double a = 0.0, b = 0.0;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
        a += log(A[i * ld + i]);
        const double v = A[n * ld + i];
        b = fma(v, v, b);
    }
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
    }
    if ((threadIdx.x & 31) == 0) { s0[threadIdx.x >> 5] = a; s1[threadIdx.x >> 5] = b; }
    __syncthreads();
    if (threadIdx.x < 32) {
        const int nw = blockDim.x >> 5;
        a = threadIdx.x < nw ? s0[threadIdx.x] : 0.0;
        b = threadIdx.x < nw ? s1[threadIdx.x] : 0.0;
        for (int o = 16; o > 0; o >>= 1) {
            a += __shfl_xor_sync(0xffffffffu, a, o);
            b += __shfl_xor_sync(0xffffffffu, b, o);
        }
        if (threadIdx.x == 0) { scal[0] = a; scal[1] = b; }
    }
}

inline cudaError_t cholesky_configure() {
    cudaError_t e;
    if ((e = cudaFuncSetAttribute(potrf_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PD_SMEM)) != cudaSuccess) return e;
    if ((e = dgemm_nt_configure<128, 64, GM_SUB>()) != cudaSuccess) return e;
    if ((e = dgemm_nt_configure<64, 128, GM_SET>()) != cudaSuccess) return e;
    return cudaSuccess;
}

// Enqueue the whole factorisation of h->dA (Np x Np).  Returns the number of kernel launches enqueued.
inline int cholesky_enqueue(gb2_handle* h) {
    const int64_t Np = h->Np, ld = h->Np;
    const int nb = (int)(Np / TILE);
    double* A = h->dA;
    int launches = 0;
    cudaStream_t sm = h->s_main, sp = h->opt_lookahead ? h->s_panel : h->s_main;
    const bool two = h->opt_lookahead != 0;
    // event pool: per step one "panel done" and one "next column updated"
    while ((int)h->ev_pool.size() < 2 * nb + 2) {
        cudaEvent_t e;
        cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
        h->ev_pool.push_back(e);
    }
    if (two) {  // panel stream starts after everything already queued on main (the K build)
        cudaEventRecord(h->ev_pool[2 * nb], sm);
        cudaStreamWaitEvent(sp, h->ev_pool[2 * nb], 0);
    }
    for (int k = 0; k < nb; k++) {
        const int64_t g0 = (int64_t)k * TILE;
        const int64_t below = Np - g0 - TILE;
        potrf_diag_kernel<<<1, PD_THREADS, PD_SMEM, sp>>>(A, ld, g0, h->N, h->dDinv + (int64_t)k * TILE * TILE, h->dInfo);
        launches++;
        if (below > 0) {
            double* panel = A + (g0 + TILE) * ld + g0;  // rows below the diagonal block, columns of block k
            dgemm_nt_launch<64, 128, GM_SET>(sp, panel, ld, h->dDinv + (int64_t)k * TILE * TILE, TILE, panel, ld, below, TILE,
                                             TILE, 0, 0, 0);
            launches++;
            if (two) {
                cudaEventRecord(h->ev_pool[2 * k], sp);
                cudaStreamWaitEvent(sm, h->ev_pool[2 * k], 0);
            }
            // next panel column first
            double* C1 = A + (g0 + TILE) * ld + (g0 + TILE);
            dgemm_nt_launch<128, 64, GM_SUB>(sm, panel, ld, panel, ld, C1, ld, below, TILE, TILE, 1, g0 + TILE, g0 + TILE);
            launches++;
            if (two) {
                cudaEventRecord(h->ev_pool[2 * k + 1], sm);
                cudaStreamWaitEvent(sp, h->ev_pool[2 * k + 1], 0);
            }
            if (below > TILE) {
                // rest of the trailing matrix: rows from g0+2T, columns from g0+2T
                const double* Arows = panel + (int64_t)TILE * ld;  // L[i,k], i >= k+2
                const double* Brows = panel + (int64_t)TILE * ld;  // L[j,k], j >= k+2
                double* C2 = A + (g0 + 2 * TILE) * ld + (g0 + 2 * TILE);
                dgemm_nt_launch<128, 64, GM_SUB>(sm, Arows, ld, Brows, ld, C2, ld, below - TILE, below - TILE, TILE, 1,
                                                 g0 + 2 * TILE, g0 + 2 * TILE);
                launches++;
            }
        }
    }
    if (two) {  // join
        cudaEventRecord(h->ev_pool[2 * nb + 1], sp);
        cudaStreamWaitEvent(sm, h->ev_pool[2 * nb + 1], 0);
    }
    mll_terms_kernel<<<1, 1024, 0, sm>>>(A, ld, h->N, h->dScal);
    launches++;
    return launches;
}

}  // namespace gb2
