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
#include <mutex>
#include <unordered_map>
#include "gb2_internal.cuh"
#include "tf32gemm.cuh"   // mbarrier / TMA helpers (tc::mbar_*, tc::tma_load_2d, tc::encode_tiled_fn)

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

// ---------------------------------------------------------------------------------------------------------------
// TMA-staged variant (the product path; the cp.async kernel above stays as fallback for operands outside a registered
// allocation and as the ablation, set_option("dgemm_tma", 0)).
//
// Same CTA tile (128x64 / 64x128, 8 warps of 32x32, DMMA m8n8k4), same epilogue.  What changes is operand delivery:
//   * one elected thread issues two cp.async.bulk.tensor.2d per k-tile (A: BM rows x 16 doubles, B: BN rows x 16 doubles =
//     128-byte rows, hardware 128-byte swizzle) into a 4-stage ring; completion is tracked by an mbarrier per stage
//     (complete_tx), consumption by a second mbarrier per stage that every warp arrives on -- no __syncthreads and no
//     LDGSTS address arithmetic in the compute warps, and a warp never waits for another warp, only for data;
//   * fragments are read with conflict-free LDS.64 straight out of the swizzled tile: DMMA step j of a k-tile takes the four
//     k indices  8 (t >> 1) + 2 j + (t & 1)  (t = lane % 4) for BOTH operands -- a permutation of the summation order inside
//     the k-tile -- so that the 16 lanes of a half-warp (4 rows x 4 t) hit the 16 distinct 8-byte slots of a 128-byte line:
//     physical 16-byte chunk = (4 (t >> 1) + j) ^ (row & 7), 8-byte half = t & 1.
// Operands are addressed through CUtensorMaps of the ALLOCATION they live in (cuMemGetAddressRange + a cache keyed by
// base / row stride / box rows); the launch passes element coordinates.
// ---------------------------------------------------------------------------------------------------------------
constexpr int GT_STAGES = 4;
constexpr int GT_BK = 16;

template <int BM, int BN>
constexpr size_t dgemm_tma_smem_bytes() { return (size_t)GT_STAGES * (BM + BN) * 128 + 1024 /*alignment slack*/ + 128 /*barriers*/; }

template <int BM, int BN, int MODE, bool FENCE = true, int MINB = 2>
__global__ void __launch_bounds__(256, MINB)
dgemm_tma_kernel(const __grid_constant__ CUtensorMap mA, const __grid_constant__ CUtensorMap mB, int a_row0, int a_col0, int b_row0, int b_col0,
                 double* C, int64_t ldc, int kdepth, int lower_only, int64_t row_off, int64_t col_off, int rb_first, int rb_stride,
                 PushArgs push, int rb_local_first, int n_bi, int n_bj, int xprefetch) {
    constexpr int WM = 32, WN = 32, MI = WM / 8, NI = WN / 8;
    constexpr int STAGE_BYTES = (BM + BN) * 128;
    constexpr int TPB = TILE / BM;
    static_assert(TILE % BM == 0 && (BM / WM) * (BN / WN) == 8, "8 warps of 32x32");

    extern __shared__ unsigned char gt_smem_raw[];
    unsigned char* base = reinterpret_cast<unsigned char*>(((uintptr_t)gt_smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* full = reinterpret_cast<uint64_t*>(base + GT_STAGES * STAGE_BYTES);   // [STAGES] TMA bytes landed
    uint64_t* empty = full + GT_STAGES;                                              // [STAGES] all 8 warps done reading

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp % (BM / WM), wn = warp / (BM / WM);
    const int g = lane >> 2, t = lane & 3;

    if (tid == 0) {
        for (int s = 0; s < GT_STAGES; s++) { tc::mbar_init(full + s, 1); tc::mbar_init(empty + s, 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const int nk = kdepth / GT_BK;
    const int n_tiles = n_bi * n_bj;
    // tile id -> rows / first k-tile; false: the tile lies above the diagonal (lower_only) and is skipped
    auto tile_info = [&](int tile, int& bj, int64_t& grow, int64_t& lrow, int& kt0) -> bool {
        const int bi = tile % n_bi;
        bj = tile / n_bi;
        grow = ((int64_t)rb_first + (int64_t)(bi / TPB) * rb_stride) * TILE + (int64_t)(bi % TPB) * BM;
        lrow = rb_local_first >= 0 ? ((int64_t)rb_local_first + (int64_t)(bi / TPB)) * TILE + (int64_t)(bi % TPB) * BM : grow;
        if (lower_only && col_off + (int64_t)bj * BN > row_off + grow + (BM - 1)) return false;
        // lower_only == 2: both operands are rows of an upper-triangular matrix (W = L^-T), terms start at k = max(first rows)
        kt0 = lower_only == 2 ? (int)(((grow > (int64_t)bj * BN) ? grow : (int64_t)bj * BN) / GT_BK) : 0;
        return true;
    };

    // ---- producer state (thread 0 only): the ring runs ahead of the consumers ACROSS tiles, so the operands of the next tile
    // stream in while this tile's epilogue reads and writes C
    int cur_tile = blockIdx.x;   // the consumers' current tile (xprefetch == 0: the producer never opens a later one)
    int p_tile = blockIdx.x, p_it = 0, p_nit = 0, p_arow = 0, p_brow = 0, p_kt0 = 0;
    uint32_t p_g = 0;        // k-tiles issued so far
    bool p_open = false;     // p_tile's coordinates are loaded
    auto produce = [&](uint32_t upto) {   // issue k-tiles until p_g == upto or the CTA's tiles are exhausted
        while (p_g < upto) {
            if (!p_open) {
                int bj, kt0; int64_t grow, lrow;
                while (p_tile < n_tiles && !tile_info(p_tile, bj, grow, lrow, kt0)) p_tile += gridDim.x;
                if (p_tile >= n_tiles) return;
                if (!xprefetch && p_tile > cur_tile) return;
                p_arow = a_row0 + (int)lrow; p_brow = b_row0 + bj * BN; p_kt0 = kt0; p_nit = nk - kt0; p_it = 0;
                p_open = true;
                if (p_nit <= 0) { p_open = false; p_tile += gridDim.x; continue; }
            }
            const int st = (int)(p_g % GT_STAGES);
            if (p_g >= (uint32_t)GT_STAGES) tc::mbar_wait(empty + st, (p_g / GT_STAGES - 1) & 1u);
            unsigned char* dst = base + st * STAGE_BYTES;
            tc::mbar_expect_tx(full + st, STAGE_BYTES);
            tc::tma_load_2d(dst, &mA, a_col0 + (p_kt0 + p_it) * GT_BK, p_arow, full + st);
            tc::tma_load_2d(dst + BM * 128, &mB, b_col0 + (p_kt0 + p_it) * GT_BK, p_brow, full + st);
            p_g++;
            if (++p_it == p_nit) { p_open = false; p_tile += gridDim.x; }
        }
    };

    // byte offset of this lane's fragment element inside a 128-byte row, per DMMA step j: chunk (4 (t>>1) + j) ^ g, half t & 1
    const uint32_t off0 = (uint32_t)((((4 * (t >> 1)) ^ g) << 4) | ((t & 1) << 3));
    const uint32_t a_base = tc::smem_u32(base) + (uint32_t)((wm * WM + g) * 128);
    const uint32_t b_base = tc::smem_u32(base) + (uint32_t)(BM * 128 + (wn * WN + g) * 128);

    uint32_t c_g = 0;        // k-tiles consumed so far (same sequence as the producer's)
    if (tid == 0) produce(GT_STAGES - 1);
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        int bj, kt0; int64_t grow, lrow;
        if (!tile_info(tile, bj, grow, lrow, kt0)) continue;
        cur_tile = tile;
        const int n_it = nk - kt0;
        double acc[MI][NI][2];
#pragma unroll
        for (int mi = 0; mi < MI; mi++)
#pragma unroll
            for (int ni = 0; ni < NI; ni++) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;

        for (int it = 0; it < n_it; it++, c_g++) {
            if (tid == 0) produce(c_g + GT_STAGES);
            const int st = (int)(c_g % GT_STAGES);
            tc::mbar_wait(full + st, (c_g / GT_STAGES) & 1u);
            const uint32_t sa = a_base + (uint32_t)(st * STAGE_BYTES), sb = b_base + (uint32_t)(st * STAGE_BYTES);
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const uint32_t o = off0 ^ (uint32_t)(j << 4);
                double a[MI], b[NI];
#pragma unroll
                for (int mi = 0; mi < MI; mi++) asm volatile("ld.shared.f64 %0, [%1];" : "=d"(a[mi]) : "r"(sa + o + mi * 1024));
#pragma unroll
                for (int ni = 0; ni < NI; ni++) asm volatile("ld.shared.f64 %0, [%1];" : "=d"(b[ni]) : "r"(sb + o + ni * 1024));
#pragma unroll
                for (int mi = 0; mi < MI; mi++)
#pragma unroll
                    for (int ni = 0; ni < NI; ni++) dmma884(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
            }
            // Release of the stage.  The fragments were read with ordinary (generic-proxy) shared loads; the stage will be refilled by a
            // TMA write (async proxy).  Without this fence ptxas schedules the arrive right behind the ISSUE of the last LDS, ahead of
            // the DMMAs that consume it (cuobjdump of the round-2 build), so the refill could overtake loads still in flight: one warp
            // then multiplied a few 16-byte chunks of the NEXT k-tile -- about one wrong 8 x 32 patch per 10^5 tiles, enough to break
            // every large factorisation (profiles/r02e..r02q_*diag*.log, tools/micro_dgemm pipeline).  fence.proxy.async orders this
            // thread's generic accesses before the async-proxy accesses that follow in the release -> acquire chain.
            if (FENCE) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // FENCE = false: the round-2 race, compiled as a control only
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(empty + st);
        }

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
        if (!xprefetch) __syncthreads();   // diagnostic mode: tiles of a CTA are fully serialised
    }
}

// ---- host side: tensor maps of whole allocations, cached -----------------------------------------------------------------
typedef CUresult (*MemGetAddressRangeFn)(CUdeviceptr*, size_t*, CUdeviceptr);
inline MemGetAddressRangeFn mem_range_fn() {
    static MemGetAddressRangeFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<MemGetAddressRangeFn>(p);
    }
    return fn;
}

struct TmaOperand { CUtensorMap map; int row0, col0; };

// set_option("dgemm_tma", mask): process-wide switch (ablation), see dgemm_tma_try_launch; 1 -> 7, 0 = cp.async-staged kernel everywhere
inline int g_dgemm_tma = 7;
inline int g_dgemm_fence = 1;         // set_option("dgemm_fence", 0/1): 0 drops the generic->async proxy fence at the release of a stage
                                      // (reproduces the round-2 corruption; diagnostic only)
inline int g_dgemm_persistent = 1;   // set_option("dgemm_persistent", 0/1/2): 0 = one CTA per tile (ablation), 2 = persistent without cross-tile prefetch
inline int tma_sm_count() {
    static int n = 0;
    if (!n) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev); if (n <= 0) n = 148; }
    return n;
}

// Tensor map (fp64, box = 16 columns x box_rows rows, 128-byte swizzle) of the allocation holding `ptr`, viewed as a row-major
// matrix with row stride `ld`; (row0, col0) = position of `ptr` inside it.  False if the operand cannot go through TMA.
inline bool tma_operand(const double* ptr, int64_t ld, int box_rows, TmaOperand& out) {
    if (!g_dgemm_tma || ld % 2 != 0 || ld < GT_BK) return false;
    MemGetAddressRangeFn range = mem_range_fn();
    tc::EncodeTiledFn encode = tc::encode_tiled_fn();
    if (!range || !encode) return false;
    CUdeviceptr basep = 0; size_t size = 0;
    if (range(&basep, &size, (CUdeviceptr)(uintptr_t)ptr) != CUDA_SUCCESS || (basep & 15) != 0) return false;
    const int64_t off = (int64_t)(((uintptr_t)ptr - (uintptr_t)basep) / sizeof(double));
    const int64_t rows = (int64_t)(size / sizeof(double)) / ld;
    if (rows < 1 || off / ld > 0x7fffffff) return false;
    struct Key { uintptr_t base; size_t size; int64_t ld; int box; bool operator==(const Key& o) const { return base == o.base && size == o.size && ld == o.ld && box == o.box; } };
    struct Hash { size_t operator()(const Key& k) const { return (size_t)k.base * 1315423911u ^ k.size ^ ((size_t)k.ld << 7) ^ (size_t)k.box; } };
    static std::unordered_map<Key, CUtensorMap, Hash> cache;
    static std::mutex mu;
    const Key key{(uintptr_t)basep, size, ld, box_rows};
    {
        std::lock_guard<std::mutex> lk(mu);
        auto itc = cache.find(key);
        if (itc == cache.end()) {
            CUtensorMap m;
            cuuint64_t dims[2] = {(cuuint64_t)ld, (cuuint64_t)rows};
            cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(double)};
            cuuint32_t box[2] = {(cuuint32_t)GT_BK, (cuuint32_t)box_rows};
            cuuint32_t estr[2] = {1, 1};
            if (encode(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void*)(uintptr_t)basep, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
                return false;
            if (cache.size() > 4096) cache.clear();   // handles come and go (tests): bound the cache
            itc = cache.emplace(key, m).first;
        }
        out.map = itc->second;
    }
    out.row0 = (int)(off / ld);
    out.col0 = (int)(off % ld);
    return true;
}

template <int BM, int BN, int MODE>
inline bool dgemm_tma_try_launch(cudaStream_t s, const double* A, int64_t lda, const double* B, int64_t ldb, double* C, int64_t ldc, int64_t rows,
                                 int64_t cols, int kdepth, int lower_only, int64_t row_off, int64_t col_off, int rb_first, int rb_stride,
                                 const PushArgs& pa, int rb_local_first, int persistent) {
    TmaOperand oa, ob;
    // g_dgemm_tma is a bit mask (diagnostic): 1 = the in-place SET products (panel solves, solve leaves), 2 = SUB updates one
    // 128-column block wide (next-column updates), 4 = all other SUB updates
    const int usage = MODE != GM_SUB ? 1 : (cols <= TILE ? 2 : 4);
    if (!(g_dgemm_tma & usage)) return false;
    if (kdepth % GT_BK != 0 || !tma_operand(A, lda, BM, oa) || !tma_operand(B, ldb, BN, ob)) return false;
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(dgemm_tma_kernel<BM, BN, MODE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dgemm_tma_smem_bytes<BM, BN>()) != cudaSuccess ||
            cudaFuncSetAttribute(dgemm_tma_kernel<BM, BN, MODE, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dgemm_tma_smem_bytes<BM, BN>()) != cudaSuccess)
            return false;
        configured = true;
    }
    // persistent: at most two CTAs per SM walk the tile list (row tile fastest, so concurrent CTAs share the B tile in L2).  Callers whose
    // launch overlaps a latency-critical chain on another stream (the bulk updates of the factorisation) pass persistent = 0: a
    // persistent grid holds every SM until it ends, one CTA per tile lets the chain's kernels in as tiles retire
    const int n_bi = (int)(rows / BM), n_bj = (int)(cols / BN);
    const int64_t n_tiles = (int64_t)n_bi * n_bj;
    const int slots = 2 * tma_sm_count();
    // persistent: 0 = one CTA per tile, 1 = two CTAs per SM, > 1 = that many CTAs (a grid that leaves some SMs to other streams)
    const int64_t cap = persistent > 1 ? persistent : slots;
    const unsigned grid = (unsigned)(g_dgemm_persistent && persistent && n_tiles > cap ? cap : n_tiles);
    const int xprefetch = g_dgemm_persistent == 2 ? 0 : 1;
    if (g_dgemm_fence)
        dgemm_tma_kernel<BM, BN, MODE, true><<<grid, 256, dgemm_tma_smem_bytes<BM, BN>(), s>>>(oa.map, ob.map, oa.row0, oa.col0, ob.row0, ob.col0, C, ldc, kdepth,
                                                                                              lower_only, row_off, col_off, rb_first, rb_stride, pa,
                                                                                              rb_local_first, n_bi, n_bj, xprefetch);
    else   // control: the kernel WITHOUT the stage-release proxy fence (set_option("dgemm_fence", 0)), see DESIGN.md section 7a
        dgemm_tma_kernel<BM, BN, MODE, false><<<grid, 256, dgemm_tma_smem_bytes<BM, BN>(), s>>>(oa.map, ob.map, oa.row0, oa.col0, ob.row0, ob.col0, C, ldc, kdepth,
                                                                                               lower_only, row_off, col_off, rb_first, rb_stride, pa,
                                                                                               rb_local_first, n_bi, n_bj, xprefetch);
    return true;
}

// rows x cols output, both multiples of the tile.
template <int BM, int BN, int MODE, int BK = GM_BK, int STAGES = GM_STAGES, int WM = 32, int WN = 32, int MINB = 2>
inline void dgemm_nt_launch(cudaStream_t s, const double* A, int64_t lda, const double* B, int64_t ldb, double* C,
                            int64_t ldc, int64_t rows, int64_t cols, int kdepth, int lower_only, int64_t row_off,
                            int64_t col_off, int rb_first = 0, int rb_stride = 1, const PushArgs* push = nullptr,
                            int rb_local_first = -1, int persistent = 1) {
    if (rows <= 0 || cols <= 0 || kdepth <= 0) return;
    dim3 grid((unsigned)(rows / BM), (unsigned)(cols / BN));
    PushArgs pa{};
    if (push) pa = *push;
    if constexpr (BK == GM_BK && STAGES == GM_STAGES && WM == 32 && WN == 32 && MINB == 2) {   // the product instantiations
        if (dgemm_tma_try_launch<BM, BN, MODE>(s, A, lda, B, ldb, C, ldc, rows, cols, kdepth, lower_only, row_off, col_off, rb_first, rb_stride, pa,
                                               rb_local_first, persistent))
            return;
    }
    dgemm_nt_kernel<BM, BN, MODE, BK, STAGES, WM, WN, MINB><<<grid, (BM / WM) * (BN / WN) * 32, dgemm_smem_bytes<BM, BN, BK, STAGES>(), s>>>(
        A, lda, B, ldb, C, ldc, kdepth, lower_only, row_off, col_off, rb_first, rb_stride, pa, rb_local_first);
}

// C -= A B^T for the DEEP updates (two-level Cholesky, upper levels of the predict solve): from depth g_dgemm_deep on, the same kernel
// body runs with 64x128 CTA tiles and four warps of 32x64 -- 0.375 instead of 0.5 shared-memory fragment loads per DMMA, the tile
// shape of the library's own fp64 kernel.  Same k order per output entry, so results are bit-identical to the 128x64 configuration
// (tools/micro_dgemm: 0 mismatching entries); measured there +2.7 % at depth >= 4096, -5 % at depth 128, hence the threshold.
inline int g_dgemm_deep = 512;   // set_option("dgemm_deep", depth): 0 = never (ablation)
inline void dgemm_sub_launch(cudaStream_t s, const double* A, int64_t lda, const double* B, int64_t ldb, double* C, int64_t ldc, int64_t rows,
                             int64_t cols, int kdepth, int lower_only, int64_t row_off, int64_t col_off, int rb_first = 0, int rb_stride = 1,
                             int persistent = 1) {
    if (g_dgemm_deep > 0 && kdepth >= g_dgemm_deep && !(g_dgemm_tma & 4))
        dgemm_nt_launch<64, 128, GM_SUB, GM_BK, GM_STAGES, 32, 64>(s, A, lda, B, ldb, C, ldc, rows, cols, kdepth, lower_only, row_off, col_off, rb_first, rb_stride);
    else
        dgemm_nt_launch<128, 64, GM_SUB>(s, A, lda, B, ldb, C, ldc, rows, cols, kdepth, lower_only, row_off, col_off, rb_first, rb_stride, nullptr, -1,
                                         persistent);
}

}  // namespace gb2
