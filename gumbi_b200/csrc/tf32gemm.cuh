// Split-TF32 ("3xTF32") tensor-core contraction on tcgen05:  C(fp64) -= A * B^T  with fp32 TMEM accumulators.
//
// GB2_TF32 precision mode of the two dense contractions of the path: the Cholesky trailing
// SYRK update and the predict solve's GEMMs.  Plain tf32 (10-bit mantissa) cannot hold rtol 1e-2 on the posterior
// variance through a solve with cond(L) ~ 1e3, so every fp64 operand a is split once into two tf32 numbers
//     hi = tf32(a),   lo = tf32(a - hi)            (split_tf32_kernel; residual ~ 2^-22 |a|)
// and the product is accumulated as  hi*hi' + hi*lo' + lo*hi'  (three tcgen05.mma.kind::tf32 per k-step, lo*lo' ~ 2^-22
// dropped).  Accumulation is fp32 in TMEM over at most TC_MAX_K columns per launch; the epilogue converts to fp64 and
// subtracts from the fp64 matrix in HBM, so long sums are carried in fp64 across launches.
//
// Structure (one CTA per SM, persistent over 128 x 128 or 128 x 256 output tiles, 320 threads):
//   warp 0   : TMA producer   -- cp.async.bulk.tensor.2d of the four operand tiles (A_hi, A_lo, B_hi, B_lo; 128 rows x
//                                32 fp32 = 128-byte swizzled rows) into a 3-stage shared-memory ring (64 KB / stage)
//   warp 1   : MMA issuer     -- one elected lane issues 4 k-steps x 3 tcgen05.mma (M=128, N=128, K=8) per stage into one
//                                of two 128-column TMEM accumulators; tcgen05.commit releases the stage / publishes the tile
//   warps 2-9: epilogue       -- tcgen05.ld (32 lanes x 32 columns), transpose through shared memory, fp32 -> fp64, coalesced
//                                C -= acc with the next chunk's loads already in flight, release the accumulator
// Synchronisation is mbarrier-only (full/empty per stage, tmem_full/tmem_empty per accumulator).
#pragma once
#include <cuda.h>
#include "gb2_internal.cuh"

namespace gb2 {
namespace tc {

constexpr int TM = 128, TN = 128, TK = 32;       // output tile; K columns per stage (32 fp32 = one 128-byte swizzle row)
// NB = number of 128-column B tiles a CTA tile spans (output tile 128 x 128*NB).  NB = 2 halves the L2->SM operand traffic per
// flop (the A tile is staged once for 256 columns): at NB = 1 the kernel is capped by L2 bandwidth at ~50 % tensor-pipe
// utilisation (ncu, profiles/r01e), which is what a 3xTF32 product with hi+lo operands costs on 128 x 128 tiles.
constexpr int TILE_BYTES = TM * TK * 4;          // 16 KB per operand tile
constexpr int EPI_WARPS = 8;                     // two epilogue warps per TMEM lane quadrant (each takes every other 32-column chunk)
constexpr int THREADS = 64 + 32 * EPI_WARPS;
constexpr int EPI_LD = 33;                       // padded row of the per-warp 32 x 32 fp32 transpose buffer (conflict-free both ways)
constexpr int EPI_BYTES = EPI_WARPS * 32 * EPI_LD * 4;   // one buffer per epilogue warp
template <int NB> struct Cfg {
    static constexpr int STAGES = NB == 1 ? 3 : 2;
    static constexpr int STAGE_BYTES = (2 + 2 * NB) * TILE_BYTES;     // A_hi, A_lo, NB x (B_hi, B_lo)
    static constexpr int TMEM_COLS = 2 * NB * TN;                     // two accumulators of NB x 128 fp32 columns
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*alignment slack*/ + 256 /*barriers*/ + EPI_BYTES;
};
constexpr int TC_MAX_K = 1024;                   // fp32 accumulation depth per launch

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Bounded wait: a protocol bug must surface as a trapped kernel (an error code at the C ABI), never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    for (uint32_t spin = 0; spin < (1u << 26); spin++) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
        if (ok) return;
    }
    __trap();
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
// 32 lanes x 32 consecutive fp32 columns of the accumulator -> 32 registers per thread (thread = TMEM lane = tile row)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, "
        "%26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major, 128-byte-swizzled operand tile: rows are 128 B apart, 8-row swizzle atoms 1024 B apart (SBO = 64 x 16 B),
// LBO unused for swizzled K-major (1), descriptor version 1 (sm_100), layout type 2 = SWIZZLE_128B.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)64 << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// kind::tf32, D = F32, A = B = TF32, both K-major, N = 128, M = 128
constexpr uint32_t IDESC_TF32_128x128 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);

struct GemmArgs {
    double* C; int64_t ldc;          // C points at global (row 0, column 0); tile (bi, bj) covers row block rb(bi), column block cblk0 + bj
    int n_bi, n_bj;                  // tile grid: n_bi row tiles of 128, n_bj = number of 128-column blocks (CTA tiles span NB of them)
    int rb_first, rb_stride;         // row block of tile row bi = rb_first + bi * rb_stride  (block-cyclic row ownership; dense: stride 1)
    int rb_local_first;              // >= 0: C holds only the owned row blocks, contiguously: tile row bi is local block rb_local_first + bi
                                     // (storage-sharded multi-GPU mode; A rows / the triangle predicate keep using the global block)
    int cblk0;                       // first column block
    int lower;                       // 1: skip tiles whose column block lies above their row block (trailing SYRK)
    int K;                           // multiple of TK, <= TC_MAX_K
    int a_k0;                        // first K column inside the A tensor maps; A rows = row block * 128
    int b_row0, b_k0;                // B rows = b_row0 + bj * 128; first K column inside the B tensor maps
};

constexpr int RASTER_GROUP = 16;     // tile rows per raster band: concurrently running CTAs share ~16 A and ~9 B operand panels in L2

// t -> (bi, bj), banded column-major order inside bands of RASTER_GROUP tile rows; false if the tile is skipped (lower mode).
// bj is returned in units of 128-column blocks (first block of the CTA tile); nh = number of live 128-column halves (1..NB).
template <int NB>
__device__ __forceinline__ bool tile_coords(const GemmArgs& g, int t, int& bi, int& bj, int& nh) {
    const int n_sj = (g.n_bj + NB - 1) / NB;
    const int band_sz = RASTER_GROUP * n_sj;
    const int band = t / band_sz, r = t % band_sz;
    const int rows_in_band = min(RASTER_GROUP, g.n_bi - band * RASTER_GROUP);
    bi = band * RASTER_GROUP + r % rows_in_band;
    const int sj = r / rows_in_band;
    if (sj >= n_sj) return false;   // the last band may be short
    bj = sj * NB;
    int last = min(bj + NB, g.n_bj) - 1;                                           // last 128-column block inside the matrix
    if (g.lower) last = min(last, g.rb_first + bi * g.rb_stride - g.cblk0);        // ... and not above the diagonal
    nh = last - bj + 1;
    return nh >= 1;
}

template <int NB>
__global__ void __launch_bounds__(THREADS, 1)
gemm_tf32x3_kernel(const __grid_constant__ CUtensorMap mAhi, const __grid_constant__ CUtensorMap mAlo,
                   const __grid_constant__ CUtensorMap mBhi, const __grid_constant__ CUtensorMap mBlo, GemmArgs g) {
    extern __shared__ unsigned char tc_smem_raw[];
    // operand stages need 1024-byte alignment (128B swizzle atoms)
    unsigned char* base = reinterpret_cast<unsigned char*>(((uintptr_t)tc_smem_raw + 1023) & ~(uintptr_t)1023);
    constexpr int STAGES = Cfg<NB>::STAGES, STAGE_BYTES = Cfg<NB>::STAGE_BYTES, TMEM_COLS = Cfg<NB>::TMEM_COLS;
    uint64_t* bars = reinterpret_cast<uint64_t*>(base + STAGES * STAGE_BYTES);
    uint64_t* full = bars;                  // [STAGES]  TMA bytes landed
    uint64_t* empty = bars + STAGES;        // [STAGES]  MMAs that read the stage have completed
    uint64_t* tfull = bars + 2 * STAGES;    // [2]       accumulator complete
    uint64_t* tempty = bars + 2 * STAGES + 2;  // [2]    accumulator drained by the epilogue
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
    float* epi = reinterpret_cast<float*>(base + STAGES * STAGE_BYTES + 256);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_tiles = ((g.n_bi + RASTER_GROUP - 1) / RASTER_GROUP) * RASTER_GROUP * ((g.n_bj + NB - 1) / NB);   // raster slots (some are skipped)
    const int nk = g.K / TK;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < STAGES; s++) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
        for (int b = 0; b < 2; b++) { mbar_init(tfull + b, 1); mbar_init(tempty + b, EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
                int bi, bj, nh;
                if (!tile_coords<NB>(g, t, bi, bj, nh)) continue;
                const int arow = (g.rb_first + bi * g.rb_stride) * TM, brow = g.b_row0 + bj * TN;
                for (int kb = 0; kb < nk; kb++) {
                    mbar_wait(empty + stage, phase ^ 1);
                    unsigned char* st = base + stage * STAGE_BYTES;
                    mbar_expect_tx(full + stage, (2 + 2 * nh) * TILE_BYTES);
                    tma_load_2d(st, &mAhi, g.a_k0 + kb * TK, arow, full + stage);
                    tma_load_2d(st + TILE_BYTES, &mAlo, g.a_k0 + kb * TK, arow, full + stage);
                    for (int hh = 0; hh < nh; hh++) {
                        tma_load_2d(st + (2 + 2 * hh) * TILE_BYTES, &mBhi, g.b_k0 + kb * TK, brow + hh * TN, full + stage);
                        tma_load_2d(st + (3 + 2 * hh) * TILE_BYTES, &mBlo, g.b_k0 + kb * TK, brow + hh * TN, full + stage);
                    }
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer: one thread issues every tcgen05.mma and the commits that track them =================
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            int it = 0;
            for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
                int bi_, bj_, nh;
                if (!tile_coords<NB>(g, t, bi_, bj_, nh)) continue;
                const int buf = it & 1;
                const uint32_t tphase = (uint32_t)(it >> 1) & 1;
                it++;
                mbar_wait(tempty + buf, tphase ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)(buf * NB * TN);
                for (int kb = 0; kb < nk; kb++) {
                    mbar_wait(full + stage, phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(base + stage * STAGE_BYTES);
                    const uint64_t dAhi = make_smem_desc(sa), dAlo = make_smem_desc(sa + TILE_BYTES);
                    for (int hh = 0; hh < nh; hh++) {
                        const uint64_t dBhi = make_smem_desc(sa + (2 + 2 * hh) * TILE_BYTES), dBlo = make_smem_desc(sa + (3 + 2 * hh) * TILE_BYTES);
                        const uint32_t td = tmem_d + (uint32_t)(hh * TN);
#pragma unroll
                        for (int k = 0; k < TK / 8; k++) {
                            // advancing 8 tf32 = 32 bytes along K inside the swizzle row: +2 in the 16-byte address field
                            const uint64_t o = (uint64_t)(k * 2);
                            umma_tf32(td, dAlo + o, dBhi + o, IDESC_TF32_128x128, (kb | k) != 0);   // small terms first
                            umma_tf32(td, dAhi + o, dBlo + o, IDESC_TF32_128x128, 1);
                            umma_tf32(td, dAhi + o, dBhi + o, IDESC_TF32_128x128, 1);
                        }
                    }
                    umma_commit(empty + stage);                 // frees the stage once these MMAs have read it
                    if (kb == nk - 1) umma_commit(tfull + buf); // accumulator complete
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else {
        // ================= epilogue (warps 2..9) =================
        const int q = warp & 3;  // TMEM lane quadrant this warp may access
        int it = 0;
        for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
            int bi, bj, nh;
            if (!tile_coords<NB>(g, t, bi, bj, nh)) continue;
            const int buf = it & 1;
            const uint32_t tphase = (uint32_t)(it >> 1) & 1;
            it++;
            mbar_wait(tfull + buf, tphase);
            tc_fence_after();
            // TMEM hands every thread one ROW (32 consecutive columns); the fp64 read-modify-write of C wants one row spread
            // over the warp (lane = column, 256 contiguous bytes per request).  Transpose through a padded 32 x 32 buffer.
            // The C traffic of a depth-512 update is as large as its operand traffic, and with one chunk per warp in flight the
            // kernel was bound by the epilogue's memory-level parallelism (4 warps x 8 KB per SM): 8 warps, and the loads of a
            // warp's next chunk are issued before the current chunk is transposed and stored (2 x 8 KB per warp in flight).
            float* S = epi + (warp - 2) * 32 * EPI_LD;
            const int half = (warp - 2) >> 2;                      // which of the two warps of this lane quadrant
            const int64_t crow_blk = g.rb_local_first >= 0 ? (int64_t)(g.rb_local_first + bi) : (int64_t)(g.rb_first + bi * g.rb_stride);
            double* cbase = g.C + (crow_blk * TM + q * 32) * g.ldc + (int64_t)(g.cblk0 + bj) * TN + lane;
            const int n_chunks = nh * (TN / 32);
            double o[32], o2[32];
            if (half < n_chunks) {
#pragma unroll
                for (int u = 0; u < 32; u++) o[u] = cbase[half * 32 + (int64_t)u * g.ldc];
            }
#pragma unroll 1
            for (int ch = half; ch < n_chunks; ch += 2) {
                const int c0 = ch * 32;
                double* cp = cbase + c0;
                const bool more = ch + 2 < n_chunks;
                if (more) {
#pragma unroll
                    for (int u = 0; u < 32; u++) o2[u] = cp[64 + (int64_t)u * g.ldc];
                }
                float v[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * NB * TN + c0), v);
#pragma unroll
                for (int c = 0; c < 32; c++) S[lane * EPI_LD + c] = v[c];
                __syncwarp();
#pragma unroll
                for (int u = 0; u < 32; u++) cp[(int64_t)u * g.ldc] = o[u] - (double)S[u * EPI_LD + lane];
                __syncwarp();
                if (more) {
#pragma unroll
                    for (int u = 0; u < 32; u++) o[u] = o2[u];
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty + buf);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

// fp64 (rows x cols, ld) -> hi/lo tf32-valued fp32 arrays (rows x cols, ldo)
__device__ __forceinline__ float to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
__global__ void split_tf32_kernel(const double* __restrict__ A, int64_t ld, int64_t rows, int64_t cols, float* __restrict__ hi,
                                  float* __restrict__ lo, int64_t ldo) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t c2 = cols / 2;
    if (idx >= rows * c2) return;
    const int64_t r = idx / c2, c = (idx % c2) * 2;
    const double2 a = *reinterpret_cast<const double2*>(A + r * ld + c);
    const float h0 = to_tf32((float)a.x), h1 = to_tf32((float)a.y);
    const float l0 = to_tf32((float)(a.x - (double)h0)), l1 = to_tf32((float)(a.y - (double)h1));
    *reinterpret_cast<float2*>(hi + r * ldo + c) = make_float2(h0, h1);
    *reinterpret_cast<float2*>(lo + r * ldo + c) = make_float2(l0, l1);
}

// ---- host side -------------------------------------------------------------------------------------------------
// 2-D fp32 tensor (rows x cols, row stride ld floats), box = 32 columns x 128 rows, 128-byte swizzle
// cuTensorMapEncodeTiled is a driver-API symbol: it is resolved through the runtime (cudaGetDriverEntryPoint) so that the
// library carries no link-time dependency on libcuda.so.1 and still loads (and fails loudly at gb2_create) on a GPU-less box.
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

inline CUresult make_tmap(CUtensorMap* map, const float* ptr, uint64_t rows, uint64_t cols, uint64_t ld) {
    EncodeTiledFn cuTensorMapEncodeTiled = encode_tiled_fn();
    if (!cuTensorMapEncodeTiled) return CUDA_ERROR_NOT_FOUND;
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {ld * sizeof(float)};
    cuuint32_t box[2] = {(cuuint32_t)TK, (cuuint32_t)TM};
    cuuint32_t estr[2] = {1, 1};
    return cuTensorMapEncodeTiled(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), dims, strides, box, estr,
                                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
}

inline cudaError_t gemm_tf32x3_configure() {
    cudaError_t e = cudaFuncSetAttribute(gemm_tf32x3_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<1>::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(gemm_tf32x3_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<2>::SMEM_BYTES);
}

// C -= A B^T over the tile grid, K split into launches of <= TC_MAX_K columns (fp64 carry between launches).
inline void gemm_tf32x3_launch(cudaStream_t s, int n_sm, const CUtensorMap& mAhi, const CUtensorMap& mAlo, const CUtensorMap& mBhi,
                               const CUtensorMap& mBlo, GemmArgs g, int K_total, int& launches, int nb_tile = 2) {
    // 128 x 256 CTA tiles whenever there are enough of them to fill the machine, else 128 x 128
    const int NBsel = (nb_tile == 2 && g.n_bi * ((g.n_bj + 1) / 2) >= n_sm) ? 2 : 1;
    const int n_tiles = g.n_bi * ((g.n_bj + NBsel - 1) / NBsel);
    if (n_tiles <= 0 || K_total <= 0) return;
    const int grid = n_tiles < n_sm ? n_tiles : n_sm;
    const int a_k0 = g.a_k0, b_k0 = g.b_k0;
    for (int k0 = 0; k0 < K_total; k0 += TC_MAX_K) {
        g.K = K_total - k0 < TC_MAX_K ? K_total - k0 : TC_MAX_K;
        g.a_k0 = a_k0 + k0;
        g.b_k0 = b_k0 + k0;
        if (NBsel == 2) gemm_tf32x3_kernel<2><<<grid, THREADS, Cfg<2>::SMEM_BYTES, s>>>(mAhi, mAlo, mBhi, mBlo, g);
        else gemm_tf32x3_kernel<1><<<grid, THREADS, Cfg<1>::SMEM_BYTES, s>>>(mAhi, mAlo, mBhi, mBlo, g);
        launches++;
    }
}

}  // namespace tc
}  // namespace gb2
