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

// ---------------------------------------------------------------------------------------------------------------
// Diagonal-panel kernel: factor one 128x128 diagonal block and invert the factor, entirely in shared memory.
//
// The block is held as 10 lower 32x32 sub-blocks (row stride 36 doubles: rows are 16-byte aligned and the 8x4 DMMA
// fragment loads are bank-conflict free).  Work is organised as in a textbook right-looking
// blocked Cholesky with inner block 32:
//   potrf32   one warp, lane r owns row r in registers, column broadcast through a double-buffered shared column
//   trsm32    one warp per sub-block below, lane r owns row r: x <- x L_pp^-T by forward substitution (axpy form)
//   inv32     one warp, lane c owns column c of inv(L_pp)
//   update    C_ij -= L_ip L_jp^T as 32x32x32 block products on DMMA, split in 8-column units over all 16 warps
// and then the inverse of the 128x128 factor is assembled from the four 32x32 inverses with block products
//   X_ij = -X_ii ( sum_{k=j..i-1} L_ik X_kj ).
// Columns with global index >= n_real (the y row of the augmented system and the identity padding) get pivot 1.
// A non-positive pivot records info = global column + 1 (first one wins) and is replaced by 1 so that the rest of the
// pipeline stays finite; the host turns info into LinAlgError.
// ---------------------------------------------------------------------------------------------------------------
constexpr int PD_THREADS = 512;
constexpr int PB = 32;                 // inner block
constexpr int PB_LD = 36;              // row stride of a sub-block, doubles (72 words = 8 mod 32: conflict-free DMMA fragments)
constexpr int PB_SZ = PB * PB_LD;
constexpr int PD_NBLK = 10;            // lower sub-blocks of a 4x4 block grid
// Lb (10 sub-blocks of the factor) + Xt (10 sub-blocks of the inverse, stored transposed) + Xd (4 diagonal inverses)
constexpr size_t PD_SMEM = (size_t)(2 * PD_NBLK + 4) * PB_SZ * sizeof(double) + (TILE + 2 * PB) * sizeof(double);

__device__ __forceinline__ int pd_blk(int i, int j) { return i * (i + 1) / 2 + j; }

// Branch-free 1/sqrt(d): MUFU.RSQ64H seed (rsqrt.approx.ftz.f64, ~2^-22) + two Newton steps in fp64.  The library
// rsqrt() carries a predicated slow-path CALL that pins it in program order; this version is straight-line code that
// ptxas interleaves with the rank-1 update of the previous column.  d must be a normal positive number (pivots are).
__device__ __forceinline__ double pd_rsqrt(double d) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
#pragma unroll
    for (int it = 0; it < 2; it++) {
        const double e = fma(-d * y, y, 1.0);
        y = fma(0.5 * y, e, y);
    }
    return y;
}

// One unit of a block product on DMMA: a 32 x 8 slab  C[0..31][c0..c0+7] (op)= sum_k A[r][k] * Bt[c][k]
// (both operands K-contiguous, row stride PB_LD), optionally followed by a second product accumulated in registers.
//   OP 0: C -= P   1: C = P   2: C += P   3: C = -P          (P = the product(s))
//   TR false: C is a normal sub-block  C[r*LD + c];  TR true: C is stored transposed  C[c*LD + r]
//   G != nullptr: the slab is also written to global memory at G[r*ldg + c]  (row-major)
template <int OP, bool TR>
__device__ __forceinline__ void pd_unit(double* C, const double* A, const double* Bt, const double* A2, const double* Bt2,
                                        int c0, int lane, double* G = nullptr, int ldg = 0) {
    const int g = lane >> 2, t = lane & 3;
    double acc[4][2];
#pragma unroll
    for (int mi = 0; mi < 4; mi++) acc[mi][0] = acc[mi][1] = 0.0;
    const double* pa = A + g * PB_LD + t;
    const double* pb = Bt + (c0 + g) * PB_LD + t;
#pragma unroll
    for (int kk = 0; kk < PB / 4; kk++) {
        const double bv = pb[kk * 4];
#pragma unroll
        for (int mi = 0; mi < 4; mi++) dmma884(acc[mi][0], acc[mi][1], pa[mi * 8 * PB_LD + kk * 4], bv);
    }
    if (A2) {
        pa = A2 + g * PB_LD + t;
        pb = Bt2 + (c0 + g) * PB_LD + t;
#pragma unroll
        for (int kk = 0; kk < PB / 4; kk++) {
            const double bv = pb[kk * 4];
#pragma unroll
            for (int mi = 0; mi < 4; mi++) dmma884(acc[mi][0], acc[mi][1], pa[mi * 8 * PB_LD + kk * 4], bv);
        }
    }
    __syncwarp();  // C may alias Bt (in-place stages): every lane has finished reading before any lane writes
#pragma unroll
    for (int mi = 0; mi < 4; mi++) {
        const int r = mi * 8 + g, c = c0 + 2 * t;
        double v0, v1;
        if (OP == 1) { v0 = acc[mi][0]; v1 = acc[mi][1]; }
        else if (OP == 3) { v0 = -acc[mi][0]; v1 = -acc[mi][1]; }
        else {
            if (TR) { v0 = C[c * PB_LD + r]; v1 = C[(c + 1) * PB_LD + r]; }
            else { const double2 o = *reinterpret_cast<const double2*>(C + r * PB_LD + c); v0 = o.x; v1 = o.y; }
            if (OP == 0) { v0 -= acc[mi][0]; v1 -= acc[mi][1]; }
            else { v0 += acc[mi][0]; v1 += acc[mi][1]; }
        }
        if (TR) { C[c * PB_LD + r] = v0; C[(c + 1) * PB_LD + r] = v1; }
        else *reinterpret_cast<double2*>(C + r * PB_LD + c) = make_double2(v0, v1);
        if (G) *reinterpret_cast<double2*>(G + (int64_t)r * ldg + c) = make_double2(v0, v1);
    }
}

__global__ void __launch_bounds__(PD_THREADS, 1)
potrf_diag_kernel(double* __restrict__ A, int64_t ld, int64_t g0, int64_t n_real, double* __restrict__ Dinv,
                  int* __restrict__ info, long long* __restrict__ clk = nullptr) {
    extern __shared__ __align__(16) unsigned char pd_smem[];
    int clk_i = 0;
#define PD_CLK() do { if (clk && threadIdx.x == 0) clk[clk_i++] = clock64(); } while (0)
    PD_CLK();
    double* Lb = reinterpret_cast<double*>(pd_smem);     // 10 sub-blocks of the factor, L_ij[r][k]
    double* Xt = Lb + PD_NBLK * PB_SZ;                    // 10 sub-blocks of inv(L), transposed: Xt_ij[c][r] = X_ij[r][c]
    double* Xd = Xt + PD_NBLK * PB_SZ;                    // the 4 diagonal sub-blocks of inv(L), not transposed
    double* rdiag = Xd + 4 * PB_SZ;                       // 1 / L[j][j]
    double* colbuf = rdiag + TILE;                        // 2 x 32: double-buffered column broadcast of potrf32
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = PD_THREADS / 32;
    double* Ab = A + g0 * ld + g0;

    // ---- load the lower sub-blocks: one warp per (sub-block, row) = 256 contiguous bytes, 20 independent loads in flight
    // per thread; strict upper parts of the diagonal sub-blocks are zeroed
    {
        double v[PD_NBLK * PB / NW];
#pragma unroll
        for (int q = 0; q < PD_NBLK * PB / NW; q++) {
            const int pr = warp + q * NW, b = pr >> 5, r = pr & 31;
            const int bi = (b >= 6) ? 3 : (b >= 3) ? 2 : (b >= 1) ? 1 : 0;
            const int bj = b - bi * (bi + 1) / 2;
            v[q] = (bi != bj || lane <= r) ? Ab[(int64_t)(bi * PB + r) * ld + bj * PB + lane] : 0.0;
        }
#pragma unroll
        for (int q = 0; q < PD_NBLK * PB / NW; q++) {
            const int pr = warp + q * NW;
            Lb[(pr >> 5) * PB_SZ + (pr & 31) * PB_LD + lane] = v[q];
        }
    }
    __syncthreads();
    PD_CLK();

    for (int p = 0; p < 4; p++) {
        double* Lpp = Lb + pd_blk(p, p) * PB_SZ;
        // ---- potrf32: warp 0, lane r owns row r.  The next pivot's 1/sqrt is started before the column broadcast of
        // the current column is consumed, so its latency hides behind the rank-1 update.
        if (warp == 0) {
            double a[PB];
            const double2* ar = reinterpret_cast<const double2*>(Lpp + lane * PB_LD);
#pragma unroll
            for (int q = 0; q < PB / 2; q++) { const double2 v = ar[q]; a[2 * q] = v.x; a[2 * q + 1] = v.y; }
            // pivots of the y row / identity padding are 1; a non-positive pivot is replaced by 1 and remembered
            // (branch-free, so that ptxas can overlap the 1/sqrt chain with the rank-1 update of the previous column)
            int first_bad = PB;
            auto pivot = [&](double d, int c) {
                const bool pad = g0 + p * PB + c >= n_real;
                const bool bad = !pad && !(d > 0.0);
                first_bad = (bad && c < first_bad) ? c : first_bad;
                return (pad || bad) ? 1.0 : d;
            };
            double d = pivot(__shfl_sync(0xffffffffu, a[0], 0), 0);
            double rs = pd_rsqrt(d);
#pragma unroll
            for (int c = 0; c < PB; c++) {
                double* cb = colbuf + (c & 1) * PB;
                a[c] = (lane == c) ? d * rs : a[c] * rs;
                if (lane == c) rdiag[p * PB + c] = rs;
                cb[lane] = a[c];
                __syncwarp();
                double rs_next = 0.0, d_next = 0.0;
                if (c + 1 < PB) {
                    // lane c+1 owns both numbers the next pivot needs
                    const double tmp = fma(-a[c], a[c], a[c + 1]);
                    d_next = pivot(__shfl_sync(0xffffffffu, tmp, c + 1), c + 1);
                    rs_next = pd_rsqrt(d_next);
                }
#pragma unroll
                for (int c2 = (c + 1) & ~1; c2 < PB; c2 += 2) {
                    const double2 l = *reinterpret_cast<const double2*>(cb + c2);
                    if (c2 > c) a[c2] = fma(-a[c], l.x, a[c2]);
                    a[c2 + 1] = fma(-a[c], l.y, a[c2 + 1]);
                }
                d = d_next;
                rs = rs_next;
            }
            if (first_bad < PB && lane == 0) atomicCAS(info, 0, (int)(g0 + p * PB + first_bad + 1));
            double* wr = Lpp + lane * PB_LD;
#pragma unroll
            for (int c = 0; c < PB; c++) wr[c] = (c <= lane) ? a[c] : 0.0;
        }
        __syncthreads();
        PD_CLK();

        // ---- inv32 (warp 0: column `lane` of inv(L_pp)) || trsm32 (warps 1..3-p: rows of the sub-blocks below)
        if (warp == 0) {
            double b[PB];
#pragma unroll
            for (int k = 0; k < PB; k++) b[k] = (k == lane) ? 1.0 : 0.0;
#pragma unroll
            for (int k = 0; k < PB; k++) {
                b[k] *= rdiag[p * PB + k];
#pragma unroll
                for (int r = k + 1; r < PB; r++) b[r] = fma(-Lpp[r * PB_LD + k], b[k], b[r]);
            }
            double* X = Xd + p * PB_SZ;
            double* XT = Xt + pd_blk(p, p) * PB_SZ;
            double* G = Dinv + (int64_t)(p * PB) * TILE + p * PB;
#pragma unroll
            for (int r = 0; r < PB; r++) {       // X[r][c = lane]; exact zeros above the diagonal
                X[r * PB_LD + lane] = b[r];
                G[r * TILE + lane] = b[r];
            }
            double2* wt = reinterpret_cast<double2*>(XT + lane * PB_LD);
#pragma unroll
            for (int q = 0; q < PB / 2; q++) wt[q] = make_double2(b[2 * q], b[2 * q + 1]);
        } else if (warp <= 3 - p) {
            const int i = p + warp;
            double* Lip = Lb + pd_blk(i, p) * PB_SZ;
            double x[PB];
            const double2* ar = reinterpret_cast<const double2*>(Lip + lane * PB_LD);
#pragma unroll
            for (int q = 0; q < PB / 2; q++) { const double2 v = ar[q]; x[2 * q] = v.x; x[2 * q + 1] = v.y; }
#pragma unroll
            for (int k = 0; k < PB; k++) {
                x[k] *= rdiag[p * PB + k];
#pragma unroll
                for (int c = k + 1; c < PB; c++) x[c] = fma(-Lpp[c * PB_LD + k], x[k], x[c]);
            }
            double2* wr = reinterpret_cast<double2*>(Lip + lane * PB_LD);
#pragma unroll
            for (int q = 0; q < PB / 2; q++) wr[q] = make_double2(x[2 * q], x[2 * q + 1]);
        }
        __syncthreads();
        PD_CLK();

        // ---- trailing update inside the block: C_ij -= L_ip L_jp^T for p < j <= i, 4 column units per product
        {
            const int m = 3 - p;                      // sub-blocks below
            const int n_units = m * (m + 1) / 2 * 4;
            for (int u = warp; u < n_units; u += NW) {
                const int op = u >> 2, chunk = u & 3;
                const int ii = (op >= 3) ? 2 : (op >= 1) ? 1 : 0;
                const int jj = op - ii * (ii + 1) / 2;
                const int i = p + 1 + ii, j = p + 1 + jj;
                pd_unit<0, false>(Lb + pd_blk(i, j) * PB_SZ, Lb + pd_blk(i, p) * PB_SZ, Lb + pd_blk(j, p) * PB_SZ, nullptr,
                                  nullptr, chunk * 8, lane);
            }
        }
        __syncthreads();
        PD_CLK();
    }

    // ---- factor back to global (lower triangle only); overlaps with stage A below (Lb is read-only from here on)
#pragma unroll
    for (int q = 0; q < PD_NBLK * PB / NW; q++) {
        const int pr = warp + q * NW, b = pr >> 5, r = pr & 31;
        const int bi = (b >= 6) ? 3 : (b >= 3) ? 2 : (b >= 1) ? 1 : 0;
        const int bj = b - bi * (bi + 1) / 2;
        if (bi != bj || lane <= r) Ab[(int64_t)(bi * PB + r) * ld + bj * PB + lane] = Lb[b * PB_SZ + r * PB_LD + lane];
    }
    PD_CLK();

    // ---- assemble inv(L):  X_ij = -X_ii T_ij,  T_ij = sum_{k=j..i-1} L_ik X_kj.  T_ij lives (transposed) where X_ij
    // will go; a unit only ever touches its own 8 columns of a sub-block, so the in-place steps are race-free.
    auto XT = [&](int i, int j) { return Xt + pd_blk(i, j) * PB_SZ; };
    auto L = [&](int i, int j) { return Lb + pd_blk(i, j) * PB_SZ; };
    auto GD = [&](int i, int j) { return Dinv + (int64_t)(i * PB) * TILE + j * PB; };
    // stage A: T_ij = L_ij X_jj for all i > j (6 products, 24 units)
    for (int u = warp; u < 24; u += NW) {
        const int op = u >> 2, chunk = u & 3;
        const int i = op < 1 ? 1 : (op < 3 ? 2 : 3);
        const int j = op - (i == 1 ? 0 : (i == 2 ? 1 : 3));
        pd_unit<1, true>(XT(i, j), L(i, j), XT(j, j), nullptr, nullptr, chunk * 8, lane);
    }
    __syncthreads();
    // stage B: X_{j+1,j} = -X_{j+1,j+1} T_{j+1,j}
    for (int u = warp; u < 12; u += NW) {
        const int j = u >> 2, chunk = u & 3;
        pd_unit<3, true>(XT(j + 1, j), Xd + (j + 1) * PB_SZ, XT(j + 1, j), nullptr, nullptr, chunk * 8, lane, GD(j + 1, j), TILE);
    }
    __syncthreads();
    // stage C1: T_{j+2,j} += L_{j+2,j+1} X_{j+1,j}
    for (int u = warp; u < 8; u += NW) {
        const int j = u >> 2, chunk = u & 3;
        pd_unit<2, true>(XT(j + 2, j), L(j + 2, j + 1), XT(j + 1, j), nullptr, nullptr, chunk * 8, lane);
    }
    __syncthreads();
    // stage C2: X_{j+2,j} = -X_{j+2,j+2} T_{j+2,j}
    for (int u = warp; u < 8; u += NW) {
        const int j = u >> 2, chunk = u & 3;
        pd_unit<3, true>(XT(j + 2, j), Xd + (j + 2) * PB_SZ, XT(j + 2, j), nullptr, nullptr, chunk * 8, lane, GD(j + 2, j), TILE);
    }
    __syncthreads();
    // stage D: T_30 += L_31 X_10 + L_32 X_20 ; X_30 = -X_33 T_30
    if (warp < 4) {
        pd_unit<2, true>(XT(3, 0), L(3, 1), XT(1, 0), L(3, 2), XT(2, 0), warp * 8, lane);
        __syncwarp();
        pd_unit<3, true>(XT(3, 0), Xd + 3 * PB_SZ, XT(3, 0), nullptr, nullptr, warp * 8, lane, GD(3, 0), TILE);
    }
    PD_CLK();
#undef PD_CLK
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
