// fp64 tensor-core contraction  C (-)= A * B^T  on DMMA (mma.sync.m8n8k4.f64 -> SASS DMMA.8x8x4).
//
// tcgen05.mma has no fp64 kind, so the fp64 dense contractions of the path -- the Cholesky
// trailing SYRK/GEMM update, the panel TRSM (as a product with the inverted diagonal block) and the
// predict solve L^-1 K(X,X*) -- all run through this one kernel.  Measured B200 ceiling for this
// instruction: 37.0 TFLOP/s (tools/micro_fp64.cu; cuBLAS DGEMM reaches 35.5).
//
// Both operands are "K-contiguous": A is (rows x k) row-major, B is (cols x k) row-major, which is what
// a row-major lower-triangular factor gives for L_ik L_jk^T without any transposition.
// CTA tile BM x BN (128x64 or 64x128), 8 warps of 32x32, BK = 16, 3-stage cp.async ring.
// Shared rows are padded to 20 doubles (160 B): the 8x4 DMMA fragment then reads 16 lanes x 8 B from 32
// distinct banks per half-warp (row stride = 8 banks mod 32), i.e. conflict-free LDS.64.
// All extents are multiples of the tile (the host pads N and M to 128), so the main loop has no bounds checks.
#pragma once
#include "gb2_internal.cuh"

namespace gb2 {

constexpr int GM_BK = 16;      // product configuration; the kernel is also parametrised over BK / stages / warp tile so that
constexpr int GM_STAGES = 3;   // tools/micro_dgemm.cu can time variants (BK=32 x 2 stages, 32x64 warp tiles) without touching it
constexpr int GM_THREADS = 256;
__host__ __device__ constexpr int gm_lds(int bk) { return bk + 4; }   // padded shared row, doubles: row stride = 8 banks mod 32 for BK = 16, 32

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

enum { GM_SUB = 0, GM_SET = 1, GM_SET_PUSH = 2 };

// GM_SET_PUSH (multi-GPU panel solve fused with its exchange): every output tile is also stored, at the same offset, into the
// factor buffers of the peer GPUs (NVLink peer mappings obtained with cudaIpcOpenMemHandle), and each CTA then bumps a
// counter in every peer's memory so that the peer knows when the whole panel has landed (fence.sys + red.release.sys).
constexpr int GB2_MAX_PEERS = 8;           // push targets: up to 7 peers + (storage-sharded mode) this GPU's own panel ring
struct PushArgs {
    double* peerC[GB2_MAX_PEERS];        // target base: element (global row r, column c of the launch) lives at peerC + r * ld + c
    unsigned* peerFlag[GB2_MAX_PEERS];   // counter to bump in the target's memory (nullptr: no signal, e.g. the local ring)
    int n_peers;
    int64_t ld;                          // row stride of the targets (0: same as the local C)
};

template <int BM, int BN, int BK = GM_BK, int STAGES = GM_STAGES>
constexpr size_t dgemm_smem_bytes() { return (size_t)STAGES * (BM + BN) * gm_lds(BK) * sizeof(double); }

// C[bi*BM.., bj*BN..] (MODE==GM_SUB: -=, GM_SET: =) sum_k A[bi*BM + r, k] * B[bj*BN + c, k],  k < kdepth.
// lower_only: skip tiles lying strictly above the diagonal of the global matrix, where tile (0,0) sits at
// global (row_off, col_off).  In GM_SET mode C may alias A (in-place right-multiplication of a row panel):
// each CTA owns complete rows and has consumed all of its A rows before the epilogue stores.
//
// Row-block sharding (multi-GPU factorisation): the rows a launch covers may be every `rb_stride`-th 128-row
// block starting at block `rb_first` (block-cyclic ownership); row tile bi then sits at row
//   ((rb_first + (bi / (128/BM)) * rb_stride) * 128 + (bi % (128/BM)) * BM   relative to the A / C base pointers.
// rb_first = 0, rb_stride = 1 is the dense case.
template <int BM, int BN, int MODE, int BK = GM_BK, int STAGES = GM_STAGES, int WM = 32, int WN = 32, int MINB = 2>
__global__ void __launch_bounds__((BM / WM) * (BN / WN) * 32, MINB)
dgemm_nt_kernel(const double* A, int64_t lda, const double* __restrict__ B, int64_t ldb, double* C, int64_t ldc,
                int kdepth, int lower_only, int64_t row_off, int64_t col_off, int rb_first, int rb_stride, PushArgs push,
                int rb_local_first) {
    constexpr int THREADS = (BM / WM) * (BN / WN) * 32, LDS = gm_lds(BK), MI = WM / 8, NI = WN / 8, CH = BK / 2;
    static_assert(BM % WM == 0 && BN % WN == 0 && WM % 8 == 0 && WN % 8 == 0 && BK % 4 == 0, "warp tiles of 8x8x4 DMMA fragments");
    static_assert(TILE % BM == 0, "row tiles must not straddle 128-row blocks");
    const int bi = blockIdx.x, bj = blockIdx.y;
    constexpr int TPB = TILE / BM;
    const int64_t grow = ((int64_t)rb_first + (int64_t)(bi / TPB) * rb_stride) * TILE + (int64_t)(bi % TPB) * BM;
    // storage-sharded mode: the rows this rank owns are stored contiguously (local block rb_local_first + bi / TPB); `grow` stays
    // the global row (triangle predicate, push targets), `lrow` addresses A and C
    const int64_t lrow = rb_local_first >= 0 ? ((int64_t)rb_local_first + (int64_t)(bi / TPB)) * TILE + (int64_t)(bi % TPB) * BM : grow;
    if (lower_only && col_off + (int64_t)bj * BN > row_off + grow + (BM - 1)) return;

    extern __shared__ __align__(16) unsigned char gm_smem[];
    double* sA = reinterpret_cast<double*>(gm_smem);
    double* sB = sA + STAGES * BM * LDS;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp % (BM / WM), wn = warp / (BM / WM);
    const int g = lane >> 2, t = lane & 3;

    const double* Ag = A + lrow * lda;
    const double* Bg = B + (int64_t)bj * BN * ldb;

    auto load_stage = [&](int stage, int kt) {
        const int k0 = kt * BK;
        double* dA = sA + stage * BM * LDS;
        double* dB = sB + stage * BN * LDS;
#pragma unroll
        for (int c = tid; c < BM * CH; c += THREADS) {
            const int r = c / CH, ch = c % CH;
            cp_async16(dA + r * LDS + ch * 2, Ag + (int64_t)r * lda + k0 + ch * 2);
        }
#pragma unroll
        for (int c = tid; c < BN * CH; c += THREADS) {
            const int r = c / CH, ch = c % CH;
            cp_async16(dB + r * LDS + ch * 2, Bg + (int64_t)r * ldb + k0 + ch * 2);
        }
    };

    double acc[MI][NI][2];
#pragma unroll
    for (int mi = 0; mi < MI; mi++)
#pragma unroll
        for (int ni = 0; ni < NI; ni++) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;

    const int nk = kdepth / BK;
    // lower_only == 2: both operands are rows of an upper-triangular matrix (W = L^-T: W[r][k] = 0 for k < r), so the product
    // of this tile only has terms from k >= max(first row of the A tile, first row of the B tile)
    const int kt0 = lower_only == 2 ? (int)(((grow > (int64_t)bj * BN) ? grow : (int64_t)bj * BN) / BK) : 0;
#pragma unroll
    for (int s = 0; s < STAGES - 1; s++) {
        if (kt0 + s < nk) load_stage(s, kt0 + s);
        cp_async_commit();
    }
    for (int kt = kt0; kt < nk; kt++) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        const int nxt = kt + STAGES - 1;
        if (nxt < nk) load_stage((nxt - kt0) % STAGES, nxt);
        cp_async_commit();
        const int st = (kt - kt0) % STAGES;
        const double* cA = sA + st * BM * LDS + (wm * WM + g) * LDS + t;
        const double* cB = sB + st * BN * LDS + (wn * WN + g) * LDS + t;
#pragma unroll
        for (int kk = 0; kk < BK / 4; kk++) {
            double a[MI], b[NI];
#pragma unroll
            for (int mi = 0; mi < MI; mi++) a[mi] = cA[mi * 8 * LDS + kk * 4];
#pragma unroll
            for (int ni = 0; ni < NI; ni++) b[ni] = cB[ni * 8 * LDS + kk * 4];
#pragma unroll
            for (int mi = 0; mi < MI; mi++)
#pragma unroll
                for (int ni = 0; ni < NI; ni++) dmma884(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
        }
    }
    cp_async_wait<0>();

    double* Cg = C + (lrow + wm * WM + g) * ldc + (int64_t)bj * BN + wn * WN + 2 * t;
#pragma unroll
    for (int mi = 0; mi < MI; mi++)
#pragma unroll
        for (int ni = 0; ni < NI; ni++) {
            double2* p = reinterpret_cast<double2*>(Cg + (int64_t)mi * 8 * ldc + ni * 8);
            if (MODE == GM_SUB) {
                double2 c = *p;
                c.x -= acc[mi][ni][0];
                c.y -= acc[mi][ni][1];
                *p = c;
            } else {
                *p = make_double2(acc[mi][ni][0], acc[mi][ni][1]);
            }
        }
    if (MODE == GM_SET_PUSH) {
        const int64_t pld = push.ld ? push.ld : ldc;
        const int64_t off = (grow + wm * WM + g) * pld + (int64_t)bj * BN + wn * WN + 2 * t;
        for (int pr = 0; pr < push.n_peers; pr++) {
            double* Pg = push.peerC[pr] + off;
#pragma unroll
            for (int mi = 0; mi < MI; mi++)
#pragma unroll
                for (int ni = 0; ni < NI; ni++)
                    *reinterpret_cast<double2*>(Pg + (int64_t)mi * 8 * pld + ni * 8) = make_double2(acc[mi][ni][0], acc[mi][ni][1]);
        }
        __threadfence_system();          // this thread's peer stores are performed system-wide ...
        __syncthreads();                 // ... for every thread of the CTA, before the CTA announces its tile
        if (tid == 0)
            for (int pr = 0; pr < push.n_peers; pr++)
                if (push.peerFlag[pr]) atomicAdd_system(push.peerFlag[pr], 1u);
    }
}

template <int BM, int BN, int MODE, int BK = GM_BK, int STAGES = GM_STAGES, int WM = 32, int WN = 32, int MINB = 2>
inline cudaError_t dgemm_nt_configure() {
    return cudaFuncSetAttribute(dgemm_nt_kernel<BM, BN, MODE, BK, STAGES, WM, WN, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)dgemm_smem_bytes<BM, BN, BK, STAGES>());
}

// rows x cols output, both multiples of the tile.
template <int BM, int BN, int MODE, int BK = GM_BK, int STAGES = GM_STAGES, int WM = 32, int WN = 32, int MINB = 2>
inline void dgemm_nt_launch(cudaStream_t s, const double* A, int64_t lda, const double* B, int64_t ldb, double* C,
                            int64_t ldc, int64_t rows, int64_t cols, int kdepth, int lower_only, int64_t row_off,
                            int64_t col_off, int rb_first = 0, int rb_stride = 1, const PushArgs* push = nullptr,
                            int rb_local_first = -1) {
    if (rows <= 0 || cols <= 0 || kdepth <= 0) return;
    dim3 grid((unsigned)(rows / BM), (unsigned)(cols / BN));
    PushArgs pa{};
    if (push) pa = *push;
    dgemm_nt_kernel<BM, BN, MODE, BK, STAGES, WM, WN, MINB><<<grid, (BM / WM) * (BN / WN) * 32, dgemm_smem_bytes<BM, BN, BK, STAGES>(), s>>>(
        A, lda, B, ldb, C, ldc, kdepth, lower_only, row_off, col_off, rb_first, rb_stride, pa, rb_local_first);
}

}  // namespace gb2
