// Internal declarations shared by the kernels of the gumbi_b200 core (one translation unit).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include "../../include/gumbi_b200.h"

namespace gb2 {

constexpr int TILE = 128;  // matrix padding granule and Cholesky block size (rows/cols)

// ---------------------------------------------------------------------------------------------
// Device-side description of the covariance function.  Passed by value as a kernel parameter.
// Per point i the "feature table" F (doubles, one row per feature, point index contiguous) holds,
// for every term: d scaled coordinates x/ls, the squared norm of those, and n_lin centred linear
// coordinates x-c.  The "category table" C (int32) holds one row per distinct Coregion column.
// ---------------------------------------------------------------------------------------------
struct TermDev {
    int kind, d, n_lin, n_coreg;
    int feat_off;                   // first feature row of this term
    int cg_cat[GB2_MAX_COREG];      // category-table row used by Coregion factor f
    int cg_P[GB2_MAX_COREG];
    int cg_Boff[GB2_MAX_COREG];     // offset (doubles) of its P*P table in Btab
    double eta2, tau;
};

struct KParams {
    int n_terms, n_feat, n_cat;
    int noise_cat, noise_P, noise_Boff;   // noise_cat < 0: homoskedastic
    double sigma2, jitter;
    TermDev t[GB2_MAX_TERMS];
};

// Host-built recipe for the feature-prep kernel (column gathers + scaling).
struct PrepParams {
    int n_terms, n_cat, D_in;
    int d[GB2_MAX_TERMS], n_lin[GB2_MAX_TERMS], feat_off[GB2_MAX_TERMS];
    int cont_idx[GB2_MAX_TERMS][GB2_MAX_D];
    double inv_ls[GB2_MAX_TERMS][GB2_MAX_D];
    int lin_idx[GB2_MAX_TERMS][GB2_MAX_LIN];
    double c[GB2_MAX_TERMS][GB2_MAX_LIN];
    int cat_col[GB2_MAX_TERMS * GB2_MAX_COREG + 1];
    int cat_P[GB2_MAX_TERMS * GB2_MAX_COREG + 1];
};

inline int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }

struct NcclApi;

}  // namespace gb2

struct gb2_handle {
    int device = 0;
    int precision = GB2_FP64;
    cudaStream_t s_main = nullptr, s_panel = nullptr;
    std::string err;

    // training set
    int64_t N = 0, Np = 0;   // Np = round_up(N+1, TILE): row N carries y (augmented system), rest is identity padding
    int D_in = 0;
    double* dX = nullptr;    // (N, D_in) row-major
    double* dy = nullptr;    // (N)
    int64_t X_cap = 0, y_cap = 0;
    bool have_train = false, have_kernel = false, factorized = false;

    // kernel description
    gb2::KParams kp{};
    gb2::PrepParams pp{};
    double* dBtab = nullptr;
    int btab_len = 0;

    // derived per-train-point tables
    double* dF = nullptr;    // (n_feat, Np)
    int* dC = nullptr;       // (n_cat, Np)
    int64_t F_cap = 0, C_cap = 0;

    // factor storage
    double* dA = nullptr;    // (Np, Np) row-major; lower triangle holds K then L; row N holds y then v
    int64_t A_cap = 0;       // allocated Np
    double* dDinv = nullptr; // (Np/TILE, TILE, TILE) inverses of the diagonal blocks of L
    int* dInfo = nullptr;    // first failing pivot (1-based), 0 if none
    double* dScal = nullptr; // [0]=sum log L_ii, [1]=|v|^2

    // predict scratch
    double* dXs = nullptr; int64_t Xs_cap = 0;          // (M, D_in)
    double* dFs = nullptr; int* dCs = nullptr; int64_t Fs_cap = 0, Cs_cap = 0;
    double* dAt = nullptr; int64_t At_cap = 0;           // (Mp, Np) rows = test points
    double* dMean = nullptr; double* dVar = nullptr; int64_t out_cap = 0;
    double* dCov = nullptr; int64_t cov_cap = 0;         // (Mp, Mp) full posterior covariance (gb2_predict_full)

    // MLL-gradient scratch: W = L^-T (Np x Np), S = K^-1 (Np x Np), alpha (Np), flat gradient
    double* dW = nullptr; double* dS = nullptr; int64_t G_cap = 0;
    double* dAlpha = nullptr; int64_t alpha_cap = 0;
    uint64_t factor_count = 0, alpha_for = 0;   // dAlpha belongs to factorisation number alpha_for (0 = none)
    double* dGrad = nullptr;
    double eta_host[GB2_MAX_TERMS] = {0, 0, 0, 0};
    double sigma_host = 0.0;

    // multi-GPU row-block sharding of the factorisation (gb2_dist_init); world == 1: single GPU
    int rank = 0, world = 1;
    const gb2::NcclApi* nccl = nullptr;
    void* comm = nullptr;       // ncclComm_t
    double* dLpack = nullptr;   // (Np/TILE, TILE, TILE) contiguous copies of the diagonal blocks of L (broadcast payload)
    double* dSend = nullptr; double* dRecv = nullptr; int64_t xch_cap = 0;   // panel allgather staging
    // peer-memory exchange (NVLink): IPC mappings of the peers' factor / diagonal-block buffers and of their counters
    int opt_p2p = 1;
    bool p2p_ready = false;
    double* peerA[8] = {}; double* peerDinv[8] = {}; double* peerLpack[8] = {}; unsigned* peerFlags[8] = {};
    unsigned* dFlags = nullptr;      // [2 parities][2 kinds][p2p_nbmax] counters bumped by the peers
    int64_t p2p_nbmax = 0;
    int p2p_parity = 0;
    int x_parity = 0;
    int64_t p2p_epoch = 0;
    char* dIpcXch = nullptr;
    // storage-sharded mode (set_option("shard_storage", 1)): every rank keeps only the 128-row blocks of the factor it owns
    // (contiguously: local block li <-> global block li * world + rank) plus a ring of panel buffers the peers push into
    int opt_shard_storage = 0;
    bool compact = false;            // state of the current allocation / factorisation
    int64_t nloc = 0;                // 128-row blocks held locally (same allocation size on every rank: ceil(nb / world))
    double* dRing = nullptr; double* peerRing[8] = {};
    int ring_slots = 0; int64_t ring_slot_elems = 0;
    double* dV = nullptr; int64_t V_cap = 0;           // v = L^-1 y replicated on every rank (broadcast after the factorisation)
    double* dPart = nullptr; int64_t part_cap = 0;     // per-rank partial sums of the posterior reduction, all-gathered

    // GB2_TF32: tf32 hi/lo splits + their TMA descriptors (tf32gemm.cuh)
    int n_sm = 148;
    float* dPhi = nullptr; float* dPlo = nullptr; int64_t P_cap = 0;       // (Np, opt_tf32_nb*TILE) current factor panel
    float* dLhi = nullptr; float* dLlo = nullptr; int64_t Lsplit_cap = 0;  // (Np, Np) the factor, for the predict solve
    float* dAthi = nullptr; float* dAtlo = nullptr; int64_t Atsplit_cap = 0;  // (chunk, Np) solved prediction panel
    CUtensorMap mPhi{}, mPlo{}, mLhi{}, mLlo{}, mAthi{}, mAtlo{};
    bool L_split_valid = false;

    // options
    int opt_tf32_nb = 0;     // GB2_TF32 factor-panel width in 128-column blocks; 0 = auto (8 for N >= 8192, else 4)
    int opt_tf32_leaf = 4;   // GB2_TF32 solve: sub-solves up to this many blocks stay on the fp64 kernels
    int tf32_nb() const { return opt_tf32_nb > 0 ? opt_tf32_nb : (Np >= 8192 ? 8 : 4); }
    int opt_kbuild_v1 = 0;
    int opt_kbuild_persist = 1;   // single-term models: persistent strip kernel (kbuild_persist.cuh); 0 = round-1 kernels (ablation)
    int* dKbCtr = nullptr;        // its work counters (self-resetting)
    int opt_kbuild_occ = 4;
    int opt_chain_on_panel = 1;   // Cholesky: keep the next-column update on the panel stream (no cross-stream hop on the chain)
    int opt_lookahead = 1;
    int opt_defer_wait = 1;        // multi-GPU: on the owner of block k+1 the wait for the peers' tiles of column k gates only the bulk update
    int opt_bulk_persistent = 0;   // bulk trailing updates of the factorisation: one CTA per tile (0), persistent grid of 2 CTAs per SM (1), or a
                                   // persistent grid of this many CTAs (> 1: leaves SMs to the chain's kernels), see dgemm_tma_try_launch
    // fused cold predict (gb2_factorize_predict): prediction points ride along as ext_rows extra rows of the factor (row-major, ld ext_ld)
    double* ext_At = nullptr; int64_t ext_rows = 0, ext_ld = 0; int ext_ncols = 0;
    // "trace" option: %globaltimer stamps around the kernels of every block step of factor_steps (6 per step), see gb2_get_trace
    unsigned long long* dTrace = nullptr; int64_t trace_cap = 0; int trace_steps = 0;
    // two-level blocking of the fp64 factorisation (single GPU): panels of opt_fp64_panel column blocks, one deep DMMA update per panel
    // -1 = auto: 16 from Np >= 16384 (measured, profiles/r02j/r02k_bench_c4_*: Cholesky at N = 32768 432.6 ms plain, 405 / 390 / 383 / 380 ms
    // with panels of 2 / 4 / 8 / 16), plain below (N = 8192: 9.21 ms plain, 9.44 / 9.53 with 2 / 4)
    int opt_fp64_panel = -1; cudaStream_t s_bulk2 = nullptr;
    int fp64_panel() const { return opt_fp64_panel >= 0 ? opt_fp64_panel : (Np >= 16384 ? 16 : 0); }
    int opt_fused_group = 4;     // fused cold predict: column blocks per bulk update of the prediction rows (1, 2, 4, 8)
    int opt_solve_streams = 4;   // fp64 predict solve: split the prediction rows over this many concurrent streams (wave-tail filling);
                                 // measured (profiles/r02j_bench_*): C2 solve 24.28 / 22.97 / 22.83 ms, C4 345.3 / 339.0 / 336.9 ms with 1 / 2 / 4
    cudaStream_t s_aux[3] = {nullptr, nullptr, nullptr};
    cudaEvent_t ev_fork = nullptr, ev_join[3] = {nullptr, nullptr, nullptr};

    // sparse FITC approximation (gb2_fitc_*, fitc.cuh): the inducing-point system and the m x m B system are two inner single-GPU
    // fp64 handles; kernel_host = the description last given to gb2_set_kernel (the inner systems get copies with sigma = 0)
    gb2_handle* fitc_u = nullptr; gb2_handle* fitc_b = nullptr;
    gb2_kernel kernel_host{};
    std::vector<double> kernel_Bstore;
    double* dFitc = nullptr; int64_t fitc_cap = 0;    // [Lambda | scratch mean | scratch var], 3 x max(Np, Mp)
    double* dSt = nullptr; int64_t St_cap = 0;        // S = [A Lambda^-1/2 ; y^T Lambda^-1/2], (round_up(m+1,128) x round_up(N,128))
    double* dFitcScal = nullptr;                      // [0] sum log Lambda, [1] y^T Lambda^-1 y
    int64_t fitc_m = 0; bool fitc_ready = false;

    // timing
    cudaEvent_t ev[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev_panel[2] = {nullptr, nullptr}, ev_col[2] = {nullptr, nullptr};
    std::vector<cudaEvent_t> ev_pool;
    cudaEvent_t ev_mark[4] = {nullptr, nullptr, nullptr, nullptr};
    double timings[GB2_N_TIMINGS] = {0, 0, 0, 0, 0, 0, 0, 0};
    int64_t launches = 0;
};

#define GB2_CUDA(h, call)                                                                         \
    do {                                                                                          \
        cudaError_t e__ = (call);                                                                 \
        if (e__ != cudaSuccess) {                                                                 \
            (h)->err = std::string(#call) + ": " + cudaGetErrorString(e__) + " (" + __FILE__ +   \
                       ":" + std::to_string(__LINE__) + ")";                                     \
            return -100 - (int)e__;                                                               \
        }                                                                                         \
    } while (0)

#define GB2_ARG(h, cond, msg)                                                                     \
    do {                                                                                          \
        if (!(cond)) {                                                                            \
            (h)->err = std::string("invalid argument: ") + (msg);                                \
            return -1;                                                                            \
        }                                                                                         \
    } while (0)
