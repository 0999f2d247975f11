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
#include "tf32gemm.cuh"

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
// and the inverse of the 128x128 factor is assembled from the four 32x32 inverses with block products
//   X_ij = -X_ii ( sum_{k=j..i-1} L_ik X_kj ).
// The assembly and the write-back of the factor do NOT follow the factorisation (round 1: 12 k + 3 k of the kernel's 67 k cycles): every
// product is issued as soon as its operands are final, on the warps that idle while warp 0 runs a 32-column pivot chain (6.7 k cycles,
// 15 idle warps) or the 32x32 substitutions (3.5 k cycles) -- see the schedule at the side-work lambdas.  Only  X_3j = -X_33 T_3j
// (one round of 12 independent units) remains behind the last diagonal inverse.
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
    // one third-order step instead of two Newton steps: with e = 1 - d y^2 (|e| ~ 2^-22),  d^-1/2 = y (1 + e/2 + 3e^2/8 + O(e^3)),
    // residual 5e^3/16 ~ 2^-67.  Dependent chain: 4 fp64 operations instead of 6 -- this sits on the per-column critical path.
    const double t = d * y;
    const double e = fma(-t, y, 1.0);
    const double p = fma(0.375, e, 0.5);
    const double ye = y * e;
    return fma(ye, p, y);
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
                  int* __restrict__ info, double* __restrict__ Lpack = nullptr, PushArgs sig = PushArgs{},
                  long long* __restrict__ clk = nullptr) {
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

    // ---- side work: inverse assembly and factor write-back, spread over the warps the factorisation leaves idle --------------------------
    //   T_ij (transposed) lives where X_ij will go; a unit only ever touches its own 8 columns of a sub-block, so in-place steps are race-free.
    //   A(j):  T_ij  = L_ij X_jj            (i > j)        operands final after the substitutions of step j
    //   B(j):  X_{j+1,j} = -X_{j+1,j+1} T_{j+1,j}          after the diagonal inverse of step j+1
    //   C1(j): T_{j+2,j} += L_{j+2,j+1} X_{j+1,j}          after B(j)
    //   C2(j): X_{j+2,j} = -X_{j+2,j+2} T_{j+2,j}          after C1(j) and the diagonal inverse of step j+2
    //   D1:    T_30 += L_31 X_10 + L_32 X_20               after B(0), C2(0)
    //   last:  X_3j = -X_33 T_3j                           after the diagonal inverse of step 3
    // Schedule (sub-phases are the __syncthreads-separated parts of a step: pivot chain | substitutions | block update):
    //   step 1 pivot chain : A(0)                  step 1 substitutions: write-back of column block 0
    //   step 2 pivot chain : A(1), B(0)            step 2 substitutions: C1(0), write-back of column block 1      step 2 block update: C2(0), B(1)
    //   step 3 pivot chain : C1(1), D1, A(2)       step 3 substitutions: write-back of column block 2 and of block (3,3)
    // Side work next to a pivot chain or a substitution runs only on the 12 warps of the other three sub-partitions (warp & 3 != 0): units
    // on warps 4, 8, 12 share warp 0's dispatch port and stretched the pivot chains from 6.1 k to 8.0 k cycles (profiles/r02pa_micro_potrf.log).
    auto XT = [&](int i, int j) { return Xt + pd_blk(i, j) * PB_SZ; };
    auto L = [&](int i, int j) { return Lb + pd_blk(i, j) * PB_SZ; };
    auto GD = [&](int i, int j) { return Dinv + (int64_t)(i * PB) * TILE + j * PB; };
    auto unitA = [&](int j, int u) { const int i = j + 1 + (u >> 2); pd_unit<1, true>(XT(i, j), L(i, j), XT(j, j), nullptr, nullptr, (u & 3) * 8, lane); };
    auto unitB = [&](int j, int u) { pd_unit<3, true>(XT(j + 1, j), Xd + (j + 1) * PB_SZ, XT(j + 1, j), nullptr, nullptr, u * 8, lane, GD(j + 1, j), TILE); };
    auto unitC1 = [&](int j, int u) { pd_unit<2, true>(XT(j + 2, j), L(j + 2, j + 1), XT(j + 1, j), nullptr, nullptr, u * 8, lane); };
    auto unitC2 = [&](int j, int u) { pd_unit<3, true>(XT(j + 2, j), Xd + (j + 2) * PB_SZ, XT(j + 2, j), nullptr, nullptr, u * 8, lane, GD(j + 2, j), TILE); };
    const bool side = (warp & 3) != 0;
    const int sidx = (warp >> 2) * 3 + (warp & 3) - 1;          // 0 .. 11 over the side warps
    // write-back of the final sub-blocks (i, j), i = i_first .. 3, of column block j (lower triangle only) by workers k0 .. k0 + nw - 1 (k = this worker)
    auto store_col = [&](int j, int i_first, int k, int nw) {
        for (int pr = k; pr < (4 - i_first) * PB; pr += nw) {
            const int i = i_first + (pr >> 5), r = pr & 31;
            if (i != j || lane <= r) {
                const double v = Lb[pd_blk(i, j) * PB_SZ + r * PB_LD + lane];
                Ab[(int64_t)(i * PB + r) * ld + j * PB + lane] = v;
                if (Lpack) Lpack[(i * PB + r) * TILE + j * PB + lane] = v;   // contiguous copy for the multi-GPU broadcast
            }
        }
    };

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
        } else if (side && p == 1) {
            unitA(0, sidx);
        } else if (side && p == 2) {
            if (sidx < 8) unitA(1, sidx); else unitB(0, sidx - 8);
        } else if (side && p == 3) {
            if (sidx < 4) unitC1(1, sidx);
            else if (sidx < 8) pd_unit<2, true>(XT(3, 0), L(3, 1), XT(1, 0), L(3, 2), XT(2, 0), (sidx - 4) * 8, lane);   // D1
            else unitA(2, sidx - 8);
        }
        __syncthreads();
        PD_CLK();

        // ---- inv32 (warp 0: column `lane` of inv(L_pp)) || trsm32 (warps 1..3-p: rows of the sub-blocks below) || side work
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
            double* XTp = Xt + pd_blk(p, p) * PB_SZ;
            double* G = Dinv + (int64_t)(p * PB) * TILE + p * PB;
#pragma unroll
            for (int r = 0; r < PB; r++) {       // X[r][c = lane]; exact zeros above the diagonal
                X[r * PB_LD + lane] = b[r];
                G[r * TILE + lane] = b[r];
            }
            double2* wt = reinterpret_cast<double2*>(XTp + lane * PB_LD);
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
        } else if (side && p == 1) {
            if (sidx >= 3) store_col(0, 0, sidx - 3, 9);
        } else if (side && p == 2) {
            if (sidx >= 3 && sidx < 7) unitC1(0, sidx - 3); else if (sidx >= 7) store_col(1, 1, sidx - 7, 5);
        } else if (side && p == 3) {
            if (sidx < 8) store_col(2, 2, sidx, 8); else store_col(3, 3, sidx - 8, 4);      // block (3,3): final since the pivot chain of this step
        }
        __syncthreads();
        PD_CLK();

        // ---- trailing update inside the block: C_ij -= L_ip L_jp^T for p < j <= i, 4 column units per product (|| side work at step 2)
        if (p < 3) {
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
            if (p == 2 && warp >= 4 && warp < 12) {   // the update itself occupies warps 0..3 here
                if (warp <= 7) unitC2(0, warp - 4); else unitB(1, warp - 8);
            }
            __syncthreads();
            PD_CLK();
        }
    }

    // ---- what depends on the last diagonal inverse: X_32 = -X_33 T_32, X_31 = -X_33 T_31, X_30 = -X_33 T_30
    if (warp < 4) unitB(2, warp);
    else if (warp < 8) unitC2(1, warp - 4);
    else if (warp < 12) pd_unit<3, true>(XT(3, 0), Xd + 3 * PB_SZ, XT(3, 0), nullptr, nullptr, (warp - 8) * 8, lane, GD(3, 0), TILE);
    PD_CLK();
#undef PD_CLK
    if (sig.n_peers > 0) {
        // multi-GPU, peer-memory exchange: L_kk / inv(L_kk) are complete in this GPU's memory; tell every peer (they pull)
        __threadfence_system();
        __syncthreads();
        if (tid == 0)
            for (int pr = 0; pr < sig.n_peers; pr++) atomicAdd_system(sig.peerFlag[pr], 1u);
    }
}

// Spin (bounded) until a counter in local memory, bumped by peer GPUs over NVLink, reaches `expected`.
__device__ __forceinline__ void wait_counter(const unsigned* flag, unsigned expected) {
    const volatile unsigned* f = flag;
    const long long t0 = clock64();
    while (*f < expected) {
        if (clock64() - t0 > (1LL << 33)) __trap();   // ~4 s: a protocol bug must end as an error, not as a hung GPU
    }
    __threadfence_system();
}
__global__ void wait_counter_kernel(const unsigned* flag, unsigned expected) {
    if (threadIdx.x == 0) wait_counter(flag, expected);
}

// Device-side barrier over the peers at the start of a factorisation: the one-sided panel pushes of this factorisation must
// not land in a peer's factor while that peer still reads the previous one (predict / gradient run on the same stream as this
// kernel, so "my stream reached this point" means "I am done with the old factor").  Counters are cumulative.
__global__ void p2p_barrier_kernel(PushArgs peers, const unsigned* own, unsigned expected) {
    if (threadIdx.x == 0) {
        __threadfence_system();
        for (int pr = 0; pr < peers.n_peers; pr++) atomicAdd_system(peers.peerFlag[pr], 1u);
        wait_counter(own, expected);
    }
}

// Non-owners of block step k: wait for the owner's announcement, then pull inv(L_kk) and L_kk (2 x 128 KB) from the owner's
// memory through the peer mapping into the local copies, and L_kk into the local factor.  grid = 32 CTAs x 256 threads.
__global__ void __launch_bounds__(256)
pull_diag_kernel(const unsigned* flag, unsigned expected, const double* __restrict__ ownerDinv, const double* __restrict__ ownerLpack,
                 double* __restrict__ Dinv, double* __restrict__ Lpack, double* __restrict__ Adiag, int64_t ld) {
    if (threadIdx.x == 0) wait_counter(flag, expected);
    __syncthreads();
    const int per = TILE * TILE / 2 / gridDim.x;   // double2 elements per CTA
    for (int e = blockIdx.x * per + threadIdx.x; e < (blockIdx.x + 1) * per; e += 256) {
        const double2 d = reinterpret_cast<const double2*>(ownerDinv)[e];
        const double2 l = reinterpret_cast<const double2*>(ownerLpack)[e];
        reinterpret_cast<double2*>(Dinv)[e] = d;
        reinterpret_cast<double2*>(Lpack)[e] = l;
        if (Adiag) {   // replicated storage: the factor copy of every rank also gets the diagonal block
            const int r = (e * 2) / TILE, c = (e * 2) % TILE;
            *reinterpret_cast<double2*>(Adiag + (int64_t)r * ld + c) = l;
        }
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
    if ((e = dgemm_nt_configure<64, 128, GM_SUB, GM_BK, GM_STAGES, 32, 64>()) != cudaSuccess) return e;
    if ((e = dgemm_nt_configure<64, 128, GM_SET>()) != cudaSuccess) return e;
    if ((e = dgemm_nt_configure<64, 128, GM_SET_PUSH>()) != cudaSuccess) return e;
    return cudaSuccess;
}

// 128x128 block copies between the factor storage (ld) and a packed buffer of contiguous 128x128 blocks.
// blockIdx.x = block slot, blockIdx.y = 16-row slice.  dir 0: A -> packed, 1: packed -> A.
// Slot s of rank `r` (owner-major order of an allgather) is global row block first[r] + s*G of column block k.
__global__ void __launch_bounds__(256)
panel_pack_kernel(double* __restrict__ A, int64_t ld, int64_t col0, double* __restrict__ packed, int dir, int G, int me, int nb, int k,
                  int slots_per_rank, int only_rank) {
    const int slot = blockIdx.x;
    int r, li;
    if (only_rank >= 0) { r = only_rank; li = slot; }              // pack: my own blocks, slots 0..cnt-1
    else { r = slot / slots_per_rank; li = slot % slots_per_rank; if (r == me) return; }
    const int first = k + 1 + (((r - (k + 1)) % G) + G) % G;
    const int blk = first + li * G;
    if (blk >= nb) return;
    double* a = A + (int64_t)blk * TILE * ld + col0;
    double* p = packed + ((int64_t)(only_rank >= 0 ? li : slot)) * TILE * TILE;
    const int row0 = blockIdx.y * 8;
    for (int e = threadIdx.x; e < 8 * (TILE / 2); e += 256) {
        const int rr = row0 + e / (TILE / 2), c = (e % (TILE / 2)) * 2;
        if (dir == 0) *reinterpret_cast<double2*>(p + rr * TILE + c) = *reinterpret_cast<const double2*>(a + (int64_t)rr * ld + c);
        else *reinterpret_cast<double2*>(a + (int64_t)rr * ld + c) = *reinterpret_cast<const double2*>(p + rr * TILE + c);
    }
}

// Minimal NCCL surface, bound at run time (dlopen) so that the library loads on machines without NCCL and never pulls a
// second copy next to the one torch.distributed already loaded.
struct NcclId { char bytes[128]; };   // ncclUniqueId
struct NcclApi {
    void* lib = nullptr;
    int (*GetUniqueId)(NcclId*) = nullptr;
    int (*CommInitRank)(void**, int, NcclId, int) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    int (*Broadcast)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};
constexpr int NCCL_FLOAT64 = 8;  // ncclDataType_t ncclFloat64 / ncclDouble

// Timeline instrumentation (set_option("trace", 1); off by default): one-thread kernels that write %globaltimer (ns) in stream
// order.  A stamp needs no shared memory and one thread, so it is scheduled the moment its stream reaches it: the stamp BEFORE a
// kernel is the time that kernel became eligible, the stamp AFTER it the time it completed.  Slots per block step k (6k + ...):
//   0 panel stream: diagonal kernel eligible   1 diagonal kernel done   2 panel solve done   3 next-column update done
//   4 main stream: bulk trailing update eligible   5 bulk trailing update (and the fused predict rows) done
constexpr int TRACE_SLOTS = 6;
__global__ void stamp_kernel(unsigned long long* slot) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    *slot = t;
}
inline void trace_stamp(gb2_handle* h, cudaStream_t s, int k, int slot) {
    if (h->dTrace) stamp_kernel<<<1, 1, 0, s>>>(h->dTrace + (int64_t)k * TRACE_SLOTS + slot);
}

inline cudaEvent_t pool_event(gb2_handle* h, int idx) {
    while ((int)h->ev_pool.size() <= idx) {
        cudaEvent_t e;
        cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
        h->ev_pool.push_back(e);
    }
    return h->ev_pool[idx];
}

// Block steps k0 <= k < k1 of the right-looking factorisation, with every trailing update restricted to column blocks
// < col_limit (col_limit = nb: the plain algorithm; col_limit = k1: factor a block-column panel only and leave the rest of the
// trailing matrix to the caller -- the GB2_TF32 driver below applies that part with tcgen05).
inline int factor_steps(gb2_handle* h, int k0, int k1, int col_limit) {
    const int64_t Np = h->Np, ld = h->Np;
    const int nb = (int)(Np / TILE);
    const int G = h->world, me = h->rank;
    const NcclApi* nc = h->nccl;
    double* A = h->dA;
    int launches = 0;
    cudaStream_t sm = h->s_main, sp = (h->opt_lookahead || G > 1) ? h->s_panel : h->s_main;
    const bool two = sp != sm;
    if (G > 1 && h->p2p_ready && k0 == 0) {
        // one barrier per factorisation, queued behind this rank's previous use of the factor
        PushArgs peers{};
        const size_t slot = (size_t)4 * h->p2p_nbmax;
        for (int r = 0, q = 0; r < G; r++)
            if (r != me) peers.peerFlag[q++] = h->peerFlags[r] + slot;
        peers.n_peers = G - 1;
        h->p2p_epoch++;
        p2p_barrier_kernel<<<1, 32, 0, sm>>>(peers, h->dFlags + slot, (unsigned)(h->p2p_epoch * (G - 1)));
        launches++;
    }
    if (two) {  // the panel stream starts after everything already queued on main (the K build / the previous trailing update)
        cudaEvent_t e = pool_event(h, 3 * nb + 2);
        cudaEventRecord(e, sm);
        cudaStreamWaitEvent(sp, e, 0);
    }
    auto first_owned_after = [&](int k, int r) { return k + 1 + (((r - (k + 1)) % G) + G) % G; };   // smallest block > k owned by r
    auto count_from = [&](int first) { return first < nb ? (nb - first + G - 1) / G : 0; };
    for (int k = k0; k < k1; k++) {
        const int64_t g0 = (int64_t)k * TILE;
        const int64_t below = Np - g0 - TILE;
        const int owner = k % G;
        double* Dk = h->dDinv + (int64_t)k * TILE * TILE;
        double* Lk = G > 1 ? h->dLpack + (int64_t)k * TILE * TILE : nullptr;
        const bool p2p = G > 1 && h->p2p_ready;
        // counters of this block step in the current parity buffer: [0] = diagonal block announced, [1] = panel tiles landed
        const size_t fl = ((size_t)h->p2p_parity * 2) * h->p2p_nbmax + k;
        cudaStream_t sd = sp;
        trace_stamp(h, sd, k, 0);
        if (owner == me) {
            PushArgs sig{};
            if (p2p) {
                for (int r = 0, q = 0; r < G; r++)
                    if (r != me) sig.peerFlag[q++] = h->peerFlags[r] + fl;
                sig.n_peers = G - 1;
            }
            potrf_diag_kernel<<<1, PD_THREADS, PD_SMEM, sd>>>(A, ld, g0, h->N, Dk, h->dInfo, Lk, sig);
            launches++;
            trace_stamp(h, sd, k, 1);
        } else if (p2p) {
            pull_diag_kernel<<<32, 256, 0, sp>>>(h->dFlags + fl, 1u, h->peerDinv[owner] + (int64_t)k * TILE * TILE,
                                                 h->peerLpack[owner] + (int64_t)k * TILE * TILE, Dk, Lk, A + g0 * ld + g0, ld);
            launches++;
        }
        if (G > 1 && !p2p) {
            nc->GroupStart();
            nc->Broadcast(Dk, Dk, (size_t)TILE * TILE, NCCL_FLOAT64, owner, h->comm, sp);
            nc->Broadcast(Lk, Lk, (size_t)TILE * TILE, NCCL_FLOAT64, owner, h->comm, sp);
            nc->GroupEnd();
            launches++;
            if (owner != me)
                cudaMemcpy2DAsync(A + g0 * ld + g0, ld * sizeof(double), Lk, TILE * sizeof(double), TILE * sizeof(double), TILE,
                                  cudaMemcpyDeviceToDevice, sp);
        }
        if (below > 0) {
            const int f1 = first_owned_after(k, me), c1 = count_from(f1);         // owned blocks > k
            double* colk = A + g0;                                                   // column block k, row 0
            bool defer_wait = false;
            if (p2p) {
                // panel solve fused with its exchange: tiles are stored locally and into every peer's factor over NVLink
                if (c1 > 0) {
                    PushArgs push{};
                    for (int r = 0, q = 0; r < G; r++)
                        if (r != me) { push.peerC[q] = h->peerA[r] + g0; push.peerFlag[q] = h->peerFlags[r] + fl + h->p2p_nbmax; q++; }
                    push.n_peers = G - 1;
                    dgemm_nt_launch<64, 128, GM_SET_PUSH>(sp, colk, ld, Dk, TILE, colk, ld, (int64_t)c1 * TILE, TILE, TILE, 0, 0, 0, f1, G, &push);
                    launches++;
                }
                // every other rank pushes (its owned blocks below k) x (128/64 row tiles) CTAs' worth of tiles into this GPU.
                // The rank that owns block k+1 does not need them for the chain: its next-column update reads only rows it owns (and
                // L[k+1,k], which it owns too), and diag(k+1) only that update -- so there the wait moves off the panel stream and gates
                // just the bulk update on the main stream (defer_wait): the chain  diag(k) -> pull -> own panel rows -> next column ->
                // diag(k+1)  no longer includes the slowest peer's push.
                const unsigned expected = (unsigned)(2 * ((nb - (k + 1)) - c1));
                defer_wait = expected > 0 && two && h->opt_chain_on_panel && h->opt_defer_wait && (k + 1) % G == me && k + 1 < nb;
                if (expected > 0 && !defer_wait) {
                    wait_counter_kernel<<<1, 32, 0, sp>>>(h->dFlags + fl + h->p2p_nbmax, expected);
                    launches++;
                }
            } else if (c1 > 0) {
                dgemm_nt_launch<64, 128, GM_SET>(sp, colk, ld, Dk, TILE, colk, ld, (int64_t)c1 * TILE, TILE, TILE, 0, 0, 0, f1, G);
                launches++;
            }
            if (G > 1 && !p2p) {
                const int cmax = count_from(k + 1);                                  // most blocks any rank owns below k
                if (c1 > 0) {
                    panel_pack_kernel<<<dim3(c1, TILE / 8), 256, 0, sp>>>(A, ld, g0, h->dSend, 0, G, me, nb, k, cmax, me);
                    launches++;
                }
                nc->AllGather(h->dSend, h->dRecv, (size_t)cmax * TILE * TILE, NCCL_FLOAT64, h->comm, sp);
                panel_pack_kernel<<<dim3(cmax * G, TILE / 8), 256, 0, sp>>>(A, ld, g0, h->dRecv, 1, G, me, nb, k, cmax, -1);
                launches += 2;
            }
            // Stream choreography (opt_chain_on_panel, default): the chain  diag -> panel solve -> next-column update -> next diag
            // stays on the panel stream in program order; only the bulk update goes to the main stream.  Two event edges per
            // step remain (panel complete -> main; bulk update k-1 done -> next-column update k, normally long satisfied), and
            // neither sits between two kernels of the chain.  (The older schedule, next-column update on the main stream, put two
            // cross-stream hops of ~10 us each on the chain of every block step.)
            trace_stamp(h, sp, k, 2);
            const bool chain = two && h->opt_chain_on_panel;
            if (two) {
                cudaEvent_t e = pool_event(h, 2 * k);
                cudaEventRecord(e, sp);
                cudaStreamWaitEvent(sm, e, 0);
            }
            if (defer_wait) {   // the peers' tiles of column k: needed by the bulk update only (see above)
                wait_counter_kernel<<<1, 32, 0, sm>>>(h->dFlags + fl + h->p2p_nbmax, (unsigned)(2 * ((nb - (k + 1)) - c1)));
                launches++;
            }
            if (chain && k > k0) cudaStreamWaitEvent(sp, pool_event(h, 2 * (k - 1) + 1), 0);   // bulk update k-1 touched column k+1
            // next panel column first: A[i, k+1] -= L[i,k] L[k+1,k]^T for owned i >= k+1
            if (c1 > 0 && k + 1 < col_limit) {
                dgemm_nt_launch<128, 64, GM_SUB>(chain ? sp : sm, colk, ld, A + (g0 + TILE) * ld + g0, ld, A + g0 + TILE, ld, (int64_t)c1 * TILE,
                                                 TILE, TILE, 1, 0, g0 + TILE, f1, G);
                launches++;
            }
            if (chain) trace_stamp(h, sp, k, 3);
            if (two && !chain) {
                cudaEvent_t e = pool_event(h, 2 * k + 1);
                cudaEventRecord(e, sm);
                cudaStreamWaitEvent(sp, e, 0);
            }
            trace_stamp(h, sm, k, 4);
            if (k + 2 < col_limit) {
                // rest of the trailing matrix: owned rows >= k+2, column blocks k+2 .. col_limit-1
                const int f2 = first_owned_after(k + 1, me), c2 = count_from(f2);
                if (c2 > 0) {
                    dgemm_nt_launch<128, 64, GM_SUB>(sm, colk, ld, A + (g0 + 2 * TILE) * ld + g0, ld, A + g0 + 2 * TILE, ld,
                                                     (int64_t)c2 * TILE, (int64_t)(col_limit - (k + 2)) * TILE, TILE, 1, 0,
                                                     g0 + 2 * TILE, f2, G, nullptr, -1, h->opt_bulk_persistent);
                    launches++;
                }
            }
            if (chain) cudaEventRecord(pool_event(h, 2 * k + 1), sm);   // bulk update k done (or nothing to do)
        }
        if (h->ext_rows > 0 && k < h->ext_ncols) {
            // Fused cold predict: the prediction points are extra rows E (ext_rows x Np, = K(X*,X) on entry) below the factor.  Step k
            // gives them what it gives every row below the diagonal -- E[:,k] <- E[:,k] inv(L_kk)^T, then E[:,j] -= E[:,k] L[j,k]^T for
            // the column blocks j > k -- so that E ends as K(X*,X) L^-T, the A^T of the posterior.  These launches sit on the main
            // stream behind the bulk update (never on the diag -> panel -> next-column chain) and fill the SM slots that the
            // shrinking trailing matrix leaves idle.
            if (below <= 0 && two) {   // last block step: nothing else made the main stream wait for this diagonal block
                cudaEvent_t e = pool_event(h, 2 * k);
                cudaEventRecord(e, sp);
                cudaStreamWaitEvent(sm, e, 0);
            }
            // Column blocks are grouped w at a time (opt_fused_group): inside a group the next column receives a narrow left-looking
            // update from the group's earlier columns, and the bulk right-looking update of everything right of the group is ONE
            // GEMM of depth w*128 when the group closes -- w times fewer read-modify-write passes over E, w times the depth per launch.
            const int w = h->opt_fused_group, gw = (k / w) * w, pcol = k - gw;   // group start, position inside the group
            double* Eg = h->ext_At + (int64_t)gw * TILE;
            double* Ek = h->ext_At + g0;
            if (pcol > 0) {
                dgemm_nt_launch<128, 64, GM_SUB>(sm, Eg, h->ext_ld, A + g0 * ld + (int64_t)gw * TILE, ld, Ek, h->ext_ld, h->ext_rows, TILE,
                                                 pcol * TILE, 0, 0, 0);
                launches++;
            }
            dgemm_nt_launch<64, 128, GM_SET>(sm, Ek, h->ext_ld, Dk, TILE, Ek, h->ext_ld, h->ext_rows, TILE, TILE, 0, 0, 0);
            launches++;
            if ((pcol == w - 1 || k == h->ext_ncols - 1) && k + 1 < h->ext_ncols) {
                dgemm_nt_launch<128, 64, GM_SUB>(sm, Eg, h->ext_ld, A + (g0 + TILE) * ld + (int64_t)gw * TILE, ld, Ek + TILE, h->ext_ld, h->ext_rows,
                                                 (int64_t)(h->ext_ncols - (k + 1)) * TILE, (pcol + 1) * TILE, 0, 0, 0);
                launches++;
            }
        }
        trace_stamp(h, sm, k, 5);
    }
    if (two) {  // join
        cudaEvent_t e = pool_event(h, 3 * nb + 3);
        cudaEventRecord(e, sp);
        cudaStreamWaitEvent(sm, e, 0);
    }
    return launches;
}

// ---------------------------------------------------------------------------------------------------------------
// Storage-sharded factorisation (N beyond one GPU's HBM): rank r keeps only the row blocks i with i % G == r, contiguously.
// Same block steps as factor_steps; what changes is where the panel lives.  Column block k of L is needed by everybody as
// the B operand of the trailing update, but only for the duration of step k: it is pushed (by the fused panel-solve kernel,
// over NVLink, and into the sender's own copy too) into slot k % ring_slots of a ring of (Np x 128) panel buffers in GLOBAL row
// order.  No credit protocol is needed: a rank can only push panel k+2 after its own trailing update k, and panel k+3 cannot
// start before every rank has pushed panel k+2, so at most panels k, k+1, k+2 are live while a rank still reads panel k.
// fp64 kernels only (the tf32 panels are not wired into this mode yet).
// ---------------------------------------------------------------------------------------------------------------
__global__ void mll_terms_compact_kernel(const double* __restrict__ A, int64_t ld, int64_t n, int G, int me, double* __restrict__ part) {
    __shared__ double s0[32], s1[32];
    double a = 0.0, b = 0.0;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x)
        if ((i / TILE) % G == me) a += log(A[(((i / TILE) - me) / G * TILE + i % TILE) * ld + i]);
    if ((n / TILE) % G == me) {
        const double* vrow = A + (((n / TILE) - me) / G * TILE + n % TILE) * ld;
        for (int64_t i = threadIdx.x; i < n; i += blockDim.x) b = fma(vrow[i], vrow[i], b);
    }
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
    }
    if ((threadIdx.x & 31) == 0) { s0[threadIdx.x >> 5] = a; s1[threadIdx.x >> 5] = b; }
    __syncthreads();
    if (threadIdx.x == 0) {
        a = b = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); w++) { a += s0[w]; b += s1[w]; }
        part[0] = a; part[1] = b;
    }
}
__global__ void sum_pairs_kernel(const double* __restrict__ all, int G, double* __restrict__ out) {
    if (threadIdx.x < 2) {
        double s = 0.0;
        for (int r = 0; r < G; r++) s += all[2 * r + threadIdx.x];
        out[threadIdx.x] = s;
    }
}

// k0 <= k < k1, trailing updates restricted to column blocks < col_limit (as factor_steps).  split_c0 >= 0 (GB2_TF32): as soon as
// panel k is complete in its ring slot, its rows below block col_limit are split into the tf32 hi/lo panel buffers (column block
// k - split_c0), because the ring slot is recycled a few steps later while the tcgen05 update needs the whole 8-block panel.
inline int factor_steps_compact(gb2_handle* h, int k0, int k1, int col_limit, int split_c0 = -1) {
    const int64_t Np = h->Np, ld = h->Np;
    const int nb = (int)(Np / TILE);
    const int G = h->world, me = h->rank;
    double* A = h->dA;   // (nloc * 128, Np): local block li = global block li * G + me
    int launches = 0;
    cudaStream_t sm = h->s_main, sp = h->s_panel;
    if (k0 == 0) {   // one barrier per factorisation
        PushArgs peers{};
        const size_t slot = (size_t)4 * h->p2p_nbmax;
        for (int r = 0, q = 0; r < G; r++)
            if (r != me) peers.peerFlag[q++] = h->peerFlags[r] + slot;
        peers.n_peers = G - 1;
        h->p2p_epoch++;
        p2p_barrier_kernel<<<1, 32, 0, sm>>>(peers, h->dFlags + slot, (unsigned)(h->p2p_epoch * (G - 1)));
        launches++;
    }
    {   // the panel stream starts after everything queued on main (K build / barrier / previous tcgen05 update)
        cudaEvent_t e = pool_event(h, 3 * nb + 2);
        cudaEventRecord(e, sm);
        cudaStreamWaitEvent(sp, e, 0);
    }
    auto first_owned_after = [&](int k, int r) { return k + 1 + (((r - (k + 1)) % G) + G) % G; };
    auto count_from = [&](int first) { return first < nb ? (nb - first + G - 1) / G : 0; };
    for (int k = k0; k < k1; k++) {
        const int64_t g0 = (int64_t)k * TILE;
        const int64_t below = Np - g0 - TILE;
        const int owner = k % G;
        double* Dk = h->dDinv + (int64_t)k * TILE * TILE;
        double* Lk = h->dLpack + (int64_t)k * TILE * TILE;
        const size_t fl = ((size_t)h->p2p_parity * 2) * h->p2p_nbmax + k;
        const int64_t ring_off = (int64_t)(k % h->ring_slots) * h->ring_slot_elems;
        if (owner == me) {
            PushArgs sig{};
            for (int r = 0, q = 0; r < G; r++)
                if (r != me) sig.peerFlag[q++] = h->peerFlags[r] + fl;
            sig.n_peers = G - 1;
            const int64_t li = (k - me) / G;
            // the kernel addresses the block as A + g0*ld + g0: shift the base so that this lands on local row block li
            potrf_diag_kernel<<<1, PD_THREADS, PD_SMEM, sp>>>(A + (li * TILE - g0) * ld, ld, g0, h->N, Dk, h->dInfo, Lk, sig);
        } else {
            pull_diag_kernel<<<32, 256, 0, sp>>>(h->dFlags + fl, 1u, h->peerDinv[owner] + (int64_t)k * TILE * TILE,
                                                 h->peerLpack[owner] + (int64_t)k * TILE * TILE, Dk, Lk, nullptr, ld);
        }
        launches++;
        if (below <= 0) break;
        const int f1 = first_owned_after(k, me), c1 = count_from(f1);
        const int l1 = (f1 - me) / G;
        double* colk = A + g0;                        // column block k of the local rows
        const double* Pk = h->dRing + ring_off;       // panel k, global row order, ld = 128
        if (c1 > 0) {
            PushArgs push{};
            int q = 0;
            for (int r = 0; r < G; r++) {
                push.peerC[q] = h->peerRing[r] + ring_off;                                   // includes this rank's own ring
                push.peerFlag[q] = r == me ? nullptr : h->peerFlags[r] + fl + h->p2p_nbmax;
                q++;
            }
            push.n_peers = G;
            push.ld = TILE;
            dgemm_nt_launch<64, 128, GM_SET_PUSH>(sp, colk, ld, Dk, TILE, colk, ld, (int64_t)c1 * TILE, TILE, TILE, 0, 0, 0, f1, G, &push, l1);
            launches++;
        }
        const unsigned expected = (unsigned)(2 * ((nb - (k + 1)) - c1));
        if (expected > 0) {
            wait_counter_kernel<<<1, 32, 0, sp>>>(h->dFlags + fl + h->p2p_nbmax, expected);
            launches++;
        }
        if (split_c0 >= 0 && col_limit < nb) {
            const int64_t rows = Np - (int64_t)col_limit * TILE, pld = (int64_t)h->tf32_nb() * TILE;
            tc::split_tf32_kernel<<<(unsigned)((rows * TILE / 2 + 255) / 256), 256, 0, sp>>>(
                Pk + (int64_t)col_limit * TILE * TILE, TILE, rows, TILE, h->dPhi + (int64_t)col_limit * TILE * pld + (int64_t)(k - split_c0) * TILE,
                h->dPlo + (int64_t)col_limit * TILE * pld + (int64_t)(k - split_c0) * TILE, pld);
            launches++;
        }
        const bool chain = h->opt_chain_on_panel != 0;   // see factor_steps: the next-column update stays on the panel stream
        cudaEvent_t e0 = pool_event(h, 2 * k);
        cudaEventRecord(e0, sp);
        cudaStreamWaitEvent(sm, e0, 0);
        if (chain && k > k0) cudaStreamWaitEvent(sp, pool_event(h, 2 * (k - 1) + 1), 0);
        if (c1 > 0 && k + 1 < col_limit) {   // next column first
            dgemm_nt_launch<128, 64, GM_SUB>(chain ? sp : sm, colk, ld, Pk + (g0 + TILE) * TILE, TILE, A + g0 + TILE, ld, (int64_t)c1 * TILE, TILE,
                                             TILE, 1, 0, g0 + TILE, f1, G, nullptr, l1);
            launches++;
        }
        if (!chain) {
            cudaEvent_t e1 = pool_event(h, 2 * k + 1);
            cudaEventRecord(e1, sm);
            cudaStreamWaitEvent(sp, e1, 0);
        }
        if (k + 2 < col_limit) {
            const int f2 = first_owned_after(k + 1, me), c2 = count_from(f2);
            if (c2 > 0) {
                dgemm_nt_launch<128, 64, GM_SUB>(sm, colk, ld, Pk + (g0 + 2 * TILE) * TILE, TILE, A + g0 + 2 * TILE, ld, (int64_t)c2 * TILE,
                                                 (int64_t)(col_limit - (k + 2)) * TILE, TILE, 1, 0, g0 + 2 * TILE, f2, G, nullptr, (f2 - me) / G);
                launches++;
            }
        }
        if (chain) cudaEventRecord(pool_event(h, 2 * k + 1), sm);   // bulk update k done
    }
    cudaEvent_t e = pool_event(h, 3 * nb + 3);
    cudaEventRecord(e, sp);
    cudaStreamWaitEvent(sm, e, 0);
    return launches;
}

// Enqueue the whole factorisation of h->dA (Np x Np).  Returns the number of kernel launches enqueued.
//
// Multi-GPU (h->world > 1): 128-row blocks are owned block-cyclically (block i -> rank i % world).  Every rank
// keeps the full-size buffer but builds/updates only the rows it owns; per block step
//   owner:      potrf_diag(k)                                    -> L_kk, inv(L_kk)
//   all:        ncclBroadcast(inv(L_kk), L_kk)                    (256 KB)
//   all:        L[i,k] = A[i,k] inv(L_kk)^T  for owned i > k      (local rows of the panel)
//   all:        pack -> ncclAllGather -> unpack                   (every rank now holds the whole panel column k)
//   all:        trailing update of the owned rows (next column first = look-ahead), exactly as on one GPU.
// When the loop ends every rank holds the complete factor (each panel was gathered everywhere), so predict() runs locally
// on whatever slice of the prediction grid the rank is given.
//
// GB2_TF32: block columns are grouped in panels of h->tf32_nb() blocks (auto: 8 = 1024 columns from Np >= 8192, else 4 = 512 columns).  A panel is
// factored in fp64 by factor_steps (DMMA), split into tf32 hi/lo pairs, and the whole trailing matrix is updated by ONE
// tcgen05 split-TF32 SYRK of depth 512 (tf32gemm.cuh) -- 8x fewer passes over the trailing matrix than the 128-wide steps.
inline int cholesky_enqueue(gb2_handle* h) {
    const int64_t Np = h->Np, ld = h->Np;
    const int nb = (int)(Np / TILE);
    int launches = 0;
    const int pw = h->tf32_nb();
    if (h->compact) {
        if (h->precision == GB2_TF32 && nb > pw) {
            // tf32 panels on the storage-sharded layout: fp64 panel [c0, c1) (its column blocks are split to tf32 as they complete in
            // the ring), then one tcgen05 update of the owned row blocks >= c1 with the whole panel as the B operand
            const int G = h->world, me = h->rank;
            for (int c0 = 0; c0 < nb; c0 += pw) {
                const int c1 = c0 + pw < nb ? c0 + pw : nb;
                launches += factor_steps_compact(h, c0, c1, c1, c0);
                if (c1 >= nb) break;
                const int f = c1 + (((me - c1) % G) + G) % G;
                const int cnt = f < nb ? (nb - f + G - 1) / G : 0;
                tc::GemmArgs g{};
                g.C = h->dA; g.ldc = ld;
                g.n_bi = cnt; g.n_bj = nb - c1;
                g.rb_first = f; g.rb_stride = G; g.rb_local_first = (f - me) / G; g.cblk0 = c1; g.lower = 1;
                g.a_k0 = 0; g.b_row0 = c1 * TILE; g.b_k0 = 0;
                tc::gemm_tf32x3_launch(h->s_main, h->n_sm, h->mPhi, h->mPlo, h->mPhi, h->mPlo, g, (c1 - c0) * TILE, launches);
            }
        } else {
            launches += factor_steps_compact(h, 0, nb, nb);
        }
        // log-determinant / |v|^2: per-rank partial sums over the owned rows, all-gathered and summed; v broadcast to everybody
        const int G = h->world, me = h->rank;
        double* part = h->dScal + 2;                        // [2] mine, gathered into dPart
        mll_terms_compact_kernel<<<1, 1024, 0, h->s_main>>>(h->dA, ld, h->N, G, me, part);
        h->nccl->AllGather(part, h->dPart, 2, NCCL_FLOAT64, h->comm, h->s_main);
        sum_pairs_kernel<<<1, 32, 0, h->s_main>>>(h->dPart, G, h->dScal);
        const int vb = (int)(h->N / TILE), vowner = vb % G;
        if (vowner == me)
            cudaMemcpyAsync(h->dV, h->dA + (((int64_t)(vb - me) / G) * TILE + h->N % TILE) * ld, (size_t)Np * sizeof(double),
                            cudaMemcpyDeviceToDevice, h->s_main);
        h->nccl->Broadcast(h->dV, h->dV, (size_t)Np, NCCL_FLOAT64, vowner, h->comm, h->s_main);
        return launches + 4;
    }
    if (h->precision == GB2_TF32 && nb > pw) {
        cudaStream_t sm = h->s_main;
        const int G = h->world, me = h->rank;
        for (int c0 = 0; c0 < nb; c0 += pw) {
            const int c1 = c0 + pw < nb ? c0 + pw : nb;
            launches += factor_steps(h, c0, c1, c1);
            if (c1 >= nb) break;
            // every rank holds the whole factored panel (it was all-gathered block step by block step): split all of it ...
            const int64_t rows = Np - (int64_t)c1 * TILE, cols = (int64_t)(c1 - c0) * TILE;
            const int64_t pld = (int64_t)pw * TILE;
            tc::split_tf32_kernel<<<(unsigned)((rows * cols / 2 + 255) / 256), 256, 0, sm>>>(
                h->dA + (int64_t)c1 * TILE * ld + (int64_t)c0 * TILE, ld, rows, cols, h->dPhi + (int64_t)c1 * TILE * pld,
                h->dPlo + (int64_t)c1 * TILE * pld, pld);
            launches++;
            // ... and update the row blocks this rank owns (all of them on one GPU)
            const int f = c1 + (((me - c1) % G) + G) % G;   // first owned block >= c1
            const int cnt = f < nb ? (nb - f + G - 1) / G : 0;
            tc::GemmArgs g{};
            g.C = h->dA; g.ldc = ld;
            g.n_bi = cnt; g.n_bj = nb - c1;
            g.rb_first = f; g.rb_stride = G; g.rb_local_first = -1; g.cblk0 = c1; g.lower = 1;
            g.a_k0 = 0; g.b_row0 = c1 * TILE; g.b_k0 = 0;
            tc::gemm_tf32x3_launch(sm, h->n_sm, h->mPhi, h->mPlo, h->mPhi, h->mPlo, g, (int)cols, launches);
        }
    } else if (h->fp64_panel() > 1 && nb > h->fp64_panel() && h->ext_rows == 0) {   // (the storage-sharded mode returned above)
        // Two-level blocking (option "fp64_panel" = pw; default: auto, see gb2_handle::fp64_panel).  The plain algorithm applies every 128-column panel to the
        // whole trailing matrix: N/128 read-modify-write passes of depth 128 (87 % tensor-pipe activity, prologue/epilogue bound).
        // Here pw column blocks are factored as one panel (factor_steps restricted to the panel's own columns, exactly what the
        // tf32 path does), then applied to the trailing matrix by ONE update of depth pw*128 -- split in two so that the next panel
        // can start early:  (1) the next panel's columns, on the main stream;  (2) everything right of them, on a second bulk
        // stream, overlapping the next panel's factorisation.  (1) of group g waits for (2) of group g-1, which touched its columns.
        cudaStream_t sm = h->s_main;
        if (!h->s_bulk2) {
            int lo_p, hi_p;
            cudaDeviceGetStreamPriorityRange(&lo_p, &hi_p);
            cudaStreamCreateWithPriority(&h->s_bulk2, cudaStreamNonBlocking, lo_p);
        }
        cudaStream_t sb = h->s_bulk2;
        const int pw2 = h->fp64_panel(), ev0 = 6 * nb + 32;
        double* A = h->dA;
        // multi-GPU (replicated storage): every rank holds the whole factored panel (its tiles were pushed / all-gathered block step by
        // block step) and updates the row blocks it owns -- first owned block >= c and their count, as in the tf32 branch above
        const int G = h->world, me = h->rank;
        auto first_owned = [&](int c) { return c + (((me - c) % G) + G) % G; };
        auto count_owned = [&](int f) { return f < nb ? (nb - f + G - 1) / G : 0; };
        int g = 0;
        for (int c0 = 0; c0 < nb; c0 += pw2, g++) {
            const int c1 = c0 + pw2 < nb ? c0 + pw2 : nb;
            launches += factor_steps(h, c0, c1, c1);
            if (c1 >= nb) break;
            const int c2 = c1 + pw2 < nb ? c1 + pw2 : nb;
            const int kd = (c1 - c0) * TILE;
            const double* Apan = A + (int64_t)c0 * TILE;                  // the panel's columns; rows are addressed through rb_first
            cudaEvent_t eP = pool_event(h, ev0 + 2 * g);
            cudaEventRecord(eP, sm);                                     // panel g complete (factor_steps joined its panel stream)
            if (g > 0) cudaStreamWaitEvent(sm, pool_event(h, ev0 + 2 * (g - 1) + 1), 0);
            const int f1 = first_owned(c1), n1 = count_owned(f1);
            if (n1 > 0) {
                dgemm_sub_launch(sm, Apan, ld, A + (int64_t)c1 * TILE * ld + (int64_t)c0 * TILE, ld, A + (int64_t)c1 * TILE, ld, (int64_t)n1 * TILE,
                                 (int64_t)(c2 - c1) * TILE, kd, 1, 0, (int64_t)c1 * TILE, f1, G, h->opt_bulk_persistent);
                launches++;
            }
            cudaStreamWaitEvent(sb, eP, 0);
            const int f2 = first_owned(c2), n2 = count_owned(f2);
            if (c2 < nb && n2 > 0) {
                dgemm_sub_launch(sb, Apan, ld, A + (int64_t)c2 * TILE * ld + (int64_t)c0 * TILE, ld, A + (int64_t)c2 * TILE, ld, (int64_t)n2 * TILE,
                                 (int64_t)(nb - c2) * TILE, kd, 1, 0, (int64_t)c2 * TILE, f2, G, h->opt_bulk_persistent);
                launches++;
            }
            cudaEventRecord(pool_event(h, ev0 + 2 * g + 1), sb);
        }
        cudaEvent_t eJ = pool_event(h, ev0 + 2 * g + 2);
        cudaEventRecord(eJ, sb);
        cudaStreamWaitEvent(sm, eJ, 0);
    } else {
        launches += factor_steps(h, 0, nb, nb);
    }
    mll_terms_kernel<<<1, 1024, 0, h->s_main>>>(h->dA, ld, h->N, h->dScal);
    launches++;
    return launches;
}

}  // namespace gb2
