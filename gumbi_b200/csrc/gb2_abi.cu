// extern "C" boundary of the gumbi_b200 core (declared in include/gumbi_b200.h).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -shared -Xcompiler -fPIC gb2_abi.cu -o libgumbi_b200.so
#include "predict.cuh"
#include "mllgrad.cuh"
#include "kbuild_persist.cuh"
#include "fitc.cuh"

#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <new>

using namespace gb2;

static std::string g_create_err;

namespace {

template <typename T>
int ensure(gb2_handle* h, T*& p, int64_t& cap, int64_t need) {
    if (need <= cap && p) return 0;
    if (p) GB2_CUDA(h, cudaFree(p));
    p = nullptr; cap = 0;
    GB2_CUDA(h, cudaMalloc(&p, (size_t)need * sizeof(T)));
    cap = need;
    return 0;
}

int set_train_common(gb2_handle* h, const double* X, int64_t N, int32_t D_in, const double* y, cudaMemcpyKind kind) {
    GB2_ARG(h, X && y, "X and y must be non-null");
    GB2_ARG(h, N >= 1, "N must be >= 1");
    GB2_ARG(h, D_in >= 1 && D_in <= 64, "D_in must be in [1, 64]");
    GB2_CUDA(h, cudaSetDevice(h->device));
    const int64_t Np = round_up(N + 1, TILE);
    // buffers are re-used across calls and only ever grow: cudaFree / cudaMalloc are device-wide synchronisation points and get
    // far more expensive once peer mappings exist (the 4-GPU end-to-end arm lost ~250 ms per step with per-call re-allocation)
    int rc;
    if ((rc = ensure(h, h->dX, h->X_cap, N * (int64_t)D_in))) return rc;
    if ((rc = ensure(h, h->dy, h->y_cap, N))) return rc;
    GB2_CUDA(h, cudaMemcpyAsync(h->dX, X, (size_t)N * D_in * sizeof(double), kind, h->s_main));
    GB2_CUDA(h, cudaMemcpyAsync(h->dy, y, (size_t)N * sizeof(double), kind, h->s_main));
    GB2_CUDA(h, cudaStreamSynchronize(h->s_main));
    h->N = N; h->Np = Np; h->D_in = D_in;
    h->have_train = true; h->factorized = false; h->fitc_ready = false;
    return 0;
}

// (re)build PrepParams/KParams feature offsets; needs D_in for index validation
int validate_against_train(gb2_handle* h) {
    const PrepParams& pp = h->pp;
    for (int t = 0; t < pp.n_terms; t++) {
        for (int k = 0; k < pp.d[t]; k++) GB2_ARG(h, pp.cont_idx[t][k] >= 0 && pp.cont_idx[t][k] < h->D_in, "cont_idx out of range for D_in");
        for (int l = 0; l < pp.n_lin[t]; l++) GB2_ARG(h, pp.lin_idx[t][l] >= 0 && pp.lin_idx[t][l] < h->D_in, "lin_idx out of range for D_in");
    }
    for (int f = 0; f < pp.n_cat; f++) GB2_ARG(h, pp.cat_col[f] >= 0 && pp.cat_col[f] < h->D_in, "coregion column out of range for D_in");
    return 0;
}

int make_tmap_checked(gb2_handle* h, CUtensorMap* map, const float* ptr, int64_t rows, int64_t cols, int64_t ld) {
    const CUresult r = tc::make_tmap(map, ptr, (uint64_t)rows, (uint64_t)cols, (uint64_t)ld);
    if (r != CUDA_SUCCESS) {
        h->err = "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")";
        return -300 - (int)r;
    }
    return 0;
}

double ms_between(cudaEvent_t a, cudaEvent_t b) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, a, b);
    return (double)ms;
}

// Storage-sharded prediction (see predict.cuh).  Collective: every rank calls it with the same points.
int predict_compact(gb2_handle* h, const double* Xs, int64_t M, int32_t pred_noise, double* mean, double* var, bool dev) {
    const int64_t Np = h->Np, N = h->N;
    const int G = h->world, me = h->rank;
    const int nb = (int)(Np / TILE);
    const int ncols = (int)((N + TILE - 1) / TILE);              // column blocks that carry training points
    const int64_t chunk = std::min<int64_t>(16384, round_up(M, TILE));   // ring slots hold up to 16384 x 128
    const int64_t ldt = h->nloc * TILE;                            // local columns of At
    cudaStream_t s = h->s_main;
    int rc;
    if ((rc = ensure(h, h->dAt, h->At_cap, chunk * ldt))) return rc;
    if ((rc = ensure(h, h->dFs, h->Fs_cap, (int64_t)std::max(1, h->kp.n_feat) * chunk))) return rc;
    if ((rc = ensure(h, h->dCs, h->Cs_cap, (int64_t)std::max(1, h->kp.n_cat) * chunk))) return rc;
    const double* dXs_all = Xs;
    double* dmean_all = mean; double* dvar_all = var;
    if (!dev) {
        if ((rc = ensure(h, h->dXs, h->Xs_cap, M * h->D_in))) return rc;
        int64_t oc = h->out_cap;
        if ((rc = ensure(h, h->dMean, oc, M))) return rc;
        if ((rc = ensure(h, h->dVar, h->out_cap, M))) return rc;
        GB2_CUDA(h, cudaMemcpyAsync(h->dXs, Xs, (size_t)M * h->D_in * sizeof(double), cudaMemcpyHostToDevice, s));
        dXs_all = h->dXs; dmean_all = h->dMean; dvar_all = h->dVar;
    }
    auto first_owned_after = [&](int k, int r) { return k + 1 + (((r - (k + 1)) % G) + G) % G; };
    auto count_from = [&](int first) { return first < nb ? (nb - first + G - 1) / G : 0; };
    int launches = 0;
    double t_ks = 0, t_solve = 0, t_red = 0;
    GB2_CUDA(h, cudaMemsetAsync(h->dInfo + 1, 0, sizeof(int), s));
    for (int64_t m0 = 0; m0 < M; m0 += chunk) {
        const int64_t Mc = std::min(chunk, M - m0);
        const int64_t Mp = round_up(Mc, TILE);
        GB2_CUDA(h, cudaEventRecord(h->ev[0], s));
        prep_features<<<(unsigned)((Mp + 255) / 256), 256, 0, s>>>(dXs_all + m0 * h->D_in, Mc, Mp, h->pp, h->dFs, h->dCs, h->dInfo + 1);
        dim3 grid((unsigned)(Np / KB_T), (unsigned)(Mp / KB_T));
        kbuild_dmma_launch<false>(s, grid, kbuild_dmma_smem_bytes(h->kp), h->kp, h->dBtab, h->dFs, h->dCs, Mp, Mc, h->dF, h->dC, Np, N, nullptr,
                                  h->dAt, ldt, G, me, 1, 4, h->opt_kbuild_persist ? h->dKbCtr : nullptr, h->n_sm);
        GB2_CUDA(h, cudaEventRecord(h->ev[1], s));
        launches += 2;
        // Step counters of this chunk: two buffers behind the factorisation's counters; the one used now was cleared one chunk
        // ago, the other one (used by the previous chunk) is cleared for the next.  No barrier is needed: every chunk -- and the
        // factorisation -- ends with a collective that a rank only enters after its last push and its last ring read.
        h->x_parity ^= 1;
        unsigned* xbase = h->dFlags + (size_t)4 * h->p2p_nbmax + 4;
        const size_t xoff = (size_t)h->x_parity * h->p2p_nbmax;
        unsigned* cnt = xbase + xoff;                                                    // [j] = solved panel j has landed
        GB2_CUDA(h, cudaMemsetAsync(xbase + (size_t)(h->x_parity ^ 1) * h->p2p_nbmax, 0, (size_t)h->p2p_nbmax * sizeof(unsigned), s));
        for (int j = 0; j < ncols; j++) {
            const int owner = j % G;
            const int64_t ring_off = (int64_t)(j % h->ring_slots) * h->ring_slot_elems;
            double* Xj = h->dRing + ring_off;                                            // (Mp x 128), ld 128
            if (owner == me) {
                const int64_t lj = (j - me) / G;
                double* X = h->dAt + lj * TILE;
                PushArgs push{};
                for (int r = 0; r < G; r++) {
                    push.peerC[r] = h->peerRing[r] + ring_off;
                    push.peerFlag[r] = r == me ? nullptr : h->peerFlags[r] + (size_t)4 * h->p2p_nbmax + 4 + xoff + j;
                }
                push.n_peers = G;
                push.ld = TILE;
                dgemm_nt_launch<64, 128, GM_SET_PUSH>(s, X, ldt, h->dDinv + (int64_t)j * TILE * TILE, TILE, X, ldt, Mp, TILE, TILE, 0, 0, 0, 0, 1,
                                                      &push, -1);
            } else {
                wait_counter_kernel<<<1, 32, 0, s>>>(cnt + j, (unsigned)(Mp / 64));
            }
            launches++;
            const int f = first_owned_after(j, me), c = count_from(f);
            if (c > 0) {
                const int64_t lf = (f - me) / G;
                dgemm_nt_launch<128, 64, GM_SUB>(s, Xj, TILE, h->dA + lf * TILE * Np + (int64_t)j * TILE, Np, h->dAt + lf * TILE, ldt, Mp,
                                                 (int64_t)c * TILE, TILE, 0, 0, 0);
                launches++;
            }
        }
        GB2_CUDA(h, cudaEventRecord(h->ev[2], s));
        // partial reduction over the owned columns, all-gather, final
        const int64_t pstride = Mp;
        double* mine = h->dPart;                                   // [2 * Mp] mine, then [G][2 * Mp] gathered
        double* all = h->dPart + 2 * pstride;
        posterior_partial_kernel<<<(unsigned)((Mc + 7) / 8), 256, 0, s>>>(h->dAt, ldt, h->dV, N, Mc, ldt, G, me, mine, pstride);
        const int nrc = h->nccl->AllGather(mine, all, (size_t)2 * pstride, NCCL_FLOAT64, h->comm, s);
        if (nrc != 0) { h->err = std::string("ncclAllGather: ") + h->nccl->GetErrorString(nrc); return -200 - nrc; }
        posterior_final_kernel<<<(unsigned)((Mc + 255) / 256), 256, 0, s>>>(h->kp, h->dBtab, h->dFs, h->dCs, Mp, all, G, pstride, Mc, pred_noise,
                                                                           dmean_all + m0, dvar_all + m0);
        launches += 3;
        GB2_CUDA(h, cudaEventRecord(h->ev[3], s));
        GB2_CUDA(h, cudaGetLastError());
        GB2_CUDA(h, cudaStreamSynchronize(s));
        t_ks += ms_between(h->ev[0], h->ev[1]);
        t_solve += ms_between(h->ev[1], h->ev[2]);
        t_red += ms_between(h->ev[2], h->ev[3]);
    }
    if (!dev) {
        GB2_CUDA(h, cudaMemcpyAsync(mean, h->dMean, (size_t)M * sizeof(double), cudaMemcpyDeviceToHost, s));
        GB2_CUDA(h, cudaMemcpyAsync(var, h->dVar, (size_t)M * sizeof(double), cudaMemcpyDeviceToHost, s));
    }
    int bad = 0;
    GB2_CUDA(h, cudaMemcpyAsync(&bad, h->dInfo + 1, sizeof(int), cudaMemcpyDeviceToHost, s));
    GB2_CUDA(h, cudaStreamSynchronize(s));
    h->timings[3] = t_ks; h->timings[4] = t_solve; h->timings[5] = t_red; h->timings[7] = launches;
    GB2_ARG(h, bad == 0, "a Coregion column of Xs holds a level index outside [0, P)");
    return 0;
}

int predict_common(gb2_handle* h, const double* Xs, int64_t M, int32_t pred_noise, double* mean, double* var, bool dev) {
    GB2_ARG(h, h->factorized, "gb2_predict called before a successful gb2_factorize");
    GB2_ARG(h, Xs && mean && var, "null pointer");
    GB2_ARG(h, M >= 0, "M must be >= 0");
    if (M == 0) return 0;
    GB2_CUDA(h, cudaSetDevice(h->device));
    if (h->compact) return predict_compact(h, Xs, M, pred_noise, mean, var, dev);
    const int64_t Np = h->Np, N = h->N;
    // chunk the prediction points so that the (chunk x Np) solve panel stays within ~8 GiB
    int64_t chunk = std::max<int64_t>(TILE, ((int64_t)8 << 30) / (Np * (int64_t)sizeof(double)) / TILE * TILE);
    chunk = std::min<int64_t>(chunk, round_up(M, TILE));
    const int ncols = (int)((N + TILE - 1) / TILE);  // column blocks that carry training points
    int rc;
    if ((rc = ensure(h, h->dAt, h->At_cap, chunk * Np))) return rc;
    const bool tf32 = h->precision == GB2_TF32 && ncols > h->opt_tf32_leaf;
    if (tf32) {
        // tf32 hi/lo copies: the factor (once per factorisation) and the solved panel (per chunk)
        if (h->Lsplit_cap < Np * Np) {
            if (h->dLhi) GB2_CUDA(h, cudaFree(h->dLhi));
            if (h->dLlo) GB2_CUDA(h, cudaFree(h->dLlo));
            h->dLhi = h->dLlo = nullptr; h->Lsplit_cap = 0;
            GB2_CUDA(h, cudaMalloc(&h->dLhi, (size_t)Np * Np * sizeof(float)));
            GB2_CUDA(h, cudaMalloc(&h->dLlo, (size_t)Np * Np * sizeof(float)));
            h->Lsplit_cap = Np * Np;
            h->L_split_valid = false;
        }
        if (!h->L_split_valid) {
            if ((rc = make_tmap_checked(h, &h->mLhi, h->dLhi, Np, Np, Np))) return rc;
            if ((rc = make_tmap_checked(h, &h->mLlo, h->dLlo, Np, Np, Np))) return rc;
            tc::split_tf32_kernel<<<(unsigned)((Np * Np / 2 + 255) / 256), 256, 0, h->s_main>>>(h->dA, Np, Np, Np, h->dLhi, h->dLlo, Np);
            h->L_split_valid = true;
        }
        if (h->Atsplit_cap < chunk * Np) {
            if (h->dAthi) GB2_CUDA(h, cudaFree(h->dAthi));
            if (h->dAtlo) GB2_CUDA(h, cudaFree(h->dAtlo));
            h->dAthi = h->dAtlo = nullptr; h->Atsplit_cap = 0;
            GB2_CUDA(h, cudaMalloc(&h->dAthi, (size_t)chunk * Np * sizeof(float)));
            GB2_CUDA(h, cudaMalloc(&h->dAtlo, (size_t)chunk * Np * sizeof(float)));
            h->Atsplit_cap = chunk * Np;
        }
        if ((rc = make_tmap_checked(h, &h->mAthi, h->dAthi, chunk, Np, Np))) return rc;
        if ((rc = make_tmap_checked(h, &h->mAtlo, h->dAtlo, chunk, Np, Np))) return rc;
    }
    if ((rc = ensure(h, h->dFs, h->Fs_cap, (int64_t)std::max(1, h->kp.n_feat) * chunk))) return rc;
    if ((rc = ensure(h, h->dCs, h->Cs_cap, (int64_t)std::max(1, h->kp.n_cat) * chunk))) return rc;
    const double* dXs_all = Xs;
    double* dmean_all = mean; double* dvar_all = var;
    if (!dev) {
        if ((rc = ensure(h, h->dXs, h->Xs_cap, M * h->D_in))) return rc;
        int64_t oc = h->out_cap;
        if ((rc = ensure(h, h->dMean, oc, M))) return rc;
        if ((rc = ensure(h, h->dVar, h->out_cap, M))) return rc;
        GB2_CUDA(h, cudaMemcpyAsync(h->dXs, Xs, (size_t)M * h->D_in * sizeof(double), cudaMemcpyHostToDevice, h->s_main));
        dXs_all = h->dXs; dmean_all = h->dMean; dvar_all = h->dVar;
    }
    cudaStream_t s = h->s_main;
    double t_prep = 0, t_ks = 0, t_solve = 0, t_red = 0;
    int launches = 0;
    GB2_CUDA(h, cudaMemsetAsync(h->dInfo + 1, 0, sizeof(int), s));
    for (int64_t m0 = 0; m0 < M; m0 += chunk) {
        const int64_t Mc = std::min(chunk, M - m0);
        const int64_t Mp = round_up(Mc, TILE);
        GB2_CUDA(h, cudaEventRecord(h->ev[0], s));
        prep_features<<<(unsigned)((Mp + 255) / 256), 256, 0, s>>>(dXs_all + m0 * h->D_in, Mc, Mp, h->pp, h->dFs, h->dCs, h->dInfo + 1);
        GB2_CUDA(h, cudaEventRecord(h->ev[1], s));
        dim3 grid((unsigned)(Np / KB_T), (unsigned)(Mp / KB_T));
        if (h->opt_kbuild_v1)
            kbuild_kernel<false><<<grid, KB_THREADS, kbuild_smem_bytes(h->kp), s>>>(h->kp, h->dBtab, h->dFs, h->dCs, Mp, Mc, h->dF, h->dC, Np, N,
                                                                                     nullptr, h->dAt, Np, 1, 0);
        else
            kbuild_dmma_launch<false>(s, grid, kbuild_dmma_smem_bytes(h->kp), h->kp, h->dBtab, h->dFs, h->dCs, Mp, Mc, h->dF, h->dC, Np, N, nullptr,
                                      h->dAt, Np, 1, 0, 0, h->opt_kbuild_occ, h->opt_kbuild_persist ? h->dKbCtr : nullptr, h->n_sm);
        GB2_CUDA(h, cudaEventRecord(h->ev[2], s));
        launches += 2;
        const int n_str = (int)std::min<int64_t>(h->opt_solve_streams, Mp / TILE);
        if (tf32) trsm_rec_tf32(h, s, Mp, 0, ncols, ncols, h->opt_tf32_leaf, launches);
        else if (n_str <= 1) trsm_rec(s, h->dA, Np, h->dDinv, h->dAt, Np, Mp, 0, ncols, launches);
        else {
            // The rows of At (prediction points) are independent: run the recursion on n_str row slabs in concurrent streams, so
            // that the partial last wave of one slab's GEMM is filled by CTAs of another slab's (79 row tiles x 4 column tiles
            // on 296 CTA slots leaves the second wave 7 % full when the slabs run one after the other).
            if (!h->ev_fork) GB2_CUDA(h, cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
            GB2_CUDA(h, cudaEventRecord(h->ev_fork, s));
            const int64_t nt = Mp / TILE;
            for (int r = 0; r < n_str; r++) {
                const int64_t t0 = nt * r / n_str, t1 = nt * (r + 1) / n_str;
                cudaStream_t sr = s;
                if (r > 0) {
                    if (!h->s_aux[r - 1]) {
                        int lo_p, hi_p;
                        GB2_CUDA(h, cudaDeviceGetStreamPriorityRange(&lo_p, &hi_p));
                        GB2_CUDA(h, cudaStreamCreateWithPriority(&h->s_aux[r - 1], cudaStreamNonBlocking, lo_p));
                        GB2_CUDA(h, cudaEventCreateWithFlags(&h->ev_join[r - 1], cudaEventDisableTiming));
                    }
                    sr = h->s_aux[r - 1];
                    GB2_CUDA(h, cudaStreamWaitEvent(sr, h->ev_fork, 0));
                }
                trsm_rec(sr, h->dA, Np, h->dDinv, h->dAt + t0 * TILE * Np, Np, (t1 - t0) * TILE, 0, ncols, launches);
                if (r > 0) {
                    GB2_CUDA(h, cudaEventRecord(h->ev_join[r - 1], sr));
                    GB2_CUDA(h, cudaStreamWaitEvent(s, h->ev_join[r - 1], 0));
                }
            }
        }
        GB2_CUDA(h, cudaEventRecord(h->ev[3], s));
        posterior_reduce_kernel<<<(unsigned)((Mc + 7) / 8), 256, 0, s>>>(h->kp, h->dBtab, h->dFs, h->dCs, Mp, h->dAt, Np, h->dA + N * Np, N, Mc,
                                                                        pred_noise, dmean_all + m0, dvar_all + m0);
        launches++;
        GB2_CUDA(h, cudaEventRecord(h->ev[4], s));
        GB2_CUDA(h, cudaGetLastError());
        GB2_CUDA(h, cudaStreamSynchronize(s));
        t_prep += ms_between(h->ev[0], h->ev[1]);
        t_ks += ms_between(h->ev[1], h->ev[2]);
        t_solve += ms_between(h->ev[2], h->ev[3]);
        t_red += ms_between(h->ev[3], h->ev[4]);
    }
    if (!dev) {
        GB2_CUDA(h, cudaMemcpyAsync(mean, h->dMean, (size_t)M * sizeof(double), cudaMemcpyDeviceToHost, s));
        GB2_CUDA(h, cudaMemcpyAsync(var, h->dVar, (size_t)M * sizeof(double), cudaMemcpyDeviceToHost, s));
    }
    int bad = 0;
    GB2_CUDA(h, cudaMemcpyAsync(&bad, h->dInfo + 1, sizeof(int), cudaMemcpyDeviceToHost, s));
    GB2_CUDA(h, cudaStreamSynchronize(s));
    h->timings[3] = t_prep + t_ks; h->timings[4] = t_solve; h->timings[5] = t_red; h->timings[7] = launches;
    GB2_ARG(h, bad == 0, "a Coregion column of Xs holds a level index outside [0, P)");
    return 0;
}

}  // namespace

// driver-API entry point through the runtime (no link-time dependency on libcuda)
template <typename F>
static F driver_fn(const char* name) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) return nullptr;
    return reinterpret_cast<F>(p);
}

extern "C" {

static void p2p_close(gb2_handle* h);

int gb2_abi_version(void) { return GB2_ABI_VERSION; }

const char* gb2_last_error(const gb2_handle* h) { return h ? h->err.c_str() : g_create_err.c_str(); }

int gb2_create(gb2_handle** out, int device, int precision) {
    if (!out) { g_create_err = "invalid argument: out is null"; return -1; }
    *out = nullptr;
    if (precision != GB2_FP64 && precision != GB2_TF32) { g_create_err = "invalid argument: precision"; return -1; }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        g_create_err = std::string("no CUDA device available (") + cudaGetErrorString(e) + "); the gumbi_b200 core has no CPU fallback";
        return -2;
    }
    if (device < 0 || device >= ndev) { g_create_err = "invalid argument: device ordinal out of range"; return -1; }
    gb2_handle* h = new (std::nothrow) gb2_handle();
    if (!h) { g_create_err = "out of host memory"; return -3; }
    h->device = device; h->precision = precision;
    auto fail = [&](cudaError_t ce, const char* what) {
        g_create_err = std::string(what) + ": " + cudaGetErrorString(ce);
        delete h;
        return -100 - (int)ce;
    };
    if ((e = cudaSetDevice(device)) != cudaSuccess) return fail(e, "cudaSetDevice");
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) return fail(e, "cudaGetDeviceProperties");
    if (prop.major < 10) {
        g_create_err = "this library is built for sm_100a (B200) only; found compute capability " + std::to_string(prop.major) + "." + std::to_string(prop.minor);
        delete h;
        return -4;
    }
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    if ((e = cudaStreamCreateWithPriority(&h->s_main, cudaStreamNonBlocking, lo)) != cudaSuccess) return fail(e, "cudaStreamCreate");
    if ((e = cudaStreamCreateWithPriority(&h->s_panel, cudaStreamNonBlocking, hi)) != cudaSuccess) return fail(e, "cudaStreamCreate");
    for (auto& ev : h->ev) if ((e = cudaEventCreate(&ev)) != cudaSuccess) return fail(e, "cudaEventCreate");
    if ((e = cudaMalloc(&h->dInfo, 4 * sizeof(int))) != cudaSuccess) return fail(e, "cudaMalloc");
    if ((e = cudaMalloc(&h->dScal, 4 * sizeof(double))) != cudaSuccess) return fail(e, "cudaMalloc");
    if ((e = cholesky_configure()) != cudaSuccess) return fail(e, "cudaFuncSetAttribute");
    if ((e = tc::gemm_tf32x3_configure()) != cudaSuccess) return fail(e, "cudaFuncSetAttribute");
    h->n_sm = prop.multiProcessorCount;
    if ((e = cudaFuncSetAttribute(kbuild_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024)) != cudaSuccess) return fail(e, "cudaFuncSetAttribute");
    if ((e = cudaFuncSetAttribute(kbuild_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024)) != cudaSuccess) return fail(e, "cudaFuncSetAttribute");
    if ((e = kbuild_dmma_configure<true>()) != cudaSuccess) return fail(e, "cudaFuncSetAttribute");
    if ((e = kbuild_dmma_configure<false>()) != cudaSuccess) return fail(e, "cudaFuncSetAttribute");
    {
        double tab[64];
        for (int j = 0; j < 64; j++) tab[j] = std::exp2((double)j / 64.0);
        if ((e = cudaMemcpyToSymbol(g_exp2_tab, tab, sizeof(tab))) != cudaSuccess) return fail(e, "cudaMemcpyToSymbol");
        std::vector<double> tab2(KB4_TAB);
        for (int j = 0; j < KB4_TAB; j++) tab2[j] = std::exp2((double)j / KB4_TAB);
        if ((e = cudaMemcpyToSymbol(g_exp2_tab2k, tab2.data(), KB4_TAB * sizeof(double))) != cudaSuccess) return fail(e, "cudaMemcpyToSymbol");
    }
    // work counters of the persistent K-build kernel (self-resetting; zeroed once here)
    if ((e = cudaMalloc(&h->dKbCtr, 4 * sizeof(int))) != cudaSuccess) return fail(e, "cudaMalloc");
    if ((e = cudaMemset(h->dKbCtr, 0, 4 * sizeof(int))) != cudaSuccess) return fail(e, "cudaMemset");
    if ((e = cudaFuncSetAttribute(mll_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 164 * 1024)) != cudaSuccess) return fail(e, "cudaFuncSetAttribute");
    if ((e = dgemm_nt_configure<128, 64, GM_SET>()) != cudaSuccess) return fail(e, "cudaFuncSetAttribute");
    // GB2_OPTS="name=value,name=value": ablation switches applied to every handle of the process (measurement scripts); the
    // same names as gb2_set_option
    if (const char* env = getenv("GB2_OPTS")) {
        std::string sopt(env);
        size_t pos = 0;
        while (pos < sopt.size()) {
            size_t end = sopt.find(',', pos);
            if (end == std::string::npos) end = sopt.size();
            const std::string kv = sopt.substr(pos, end - pos);
            const size_t eq = kv.find('=');
            if (eq != std::string::npos && gb2_set_option(h, kv.substr(0, eq).c_str(), atoi(kv.c_str() + eq + 1)) != 0) {
                g_create_err = "GB2_OPTS: " + h->err;
                gb2_destroy(h);
                return -1;
            }
            pos = end + 1;
        }
    }
    *out = h;
    return 0;
}

int gb2_destroy(gb2_handle* h) {
    if (!h) return 0;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    p2p_close(h);
    if (h->comm && h->nccl) { h->nccl->CommDestroy(h->comm); h->comm = nullptr; }
    cudaFree(h->dLpack); cudaFree(h->dSend); cudaFree(h->dRecv); cudaFree(h->dFlags); cudaFree(h->dIpcXch);
    cudaFree(h->dRing); cudaFree(h->dV); cudaFree(h->dPart); cudaFree(h->dCov);
    cudaFree(h->dX); cudaFree(h->dy); cudaFree(h->dBtab); cudaFree(h->dF); cudaFree(h->dC); cudaFree(h->dA);
    cudaFree(h->dDinv); cudaFree(h->dInfo); cudaFree(h->dKbCtr); cudaFree(h->dScal); cudaFree(h->dXs); cudaFree(h->dFs); cudaFree(h->dCs);
    cudaFree(h->dAt); cudaFree(h->dMean); cudaFree(h->dVar);
    cudaFree(h->dW); cudaFree(h->dS); cudaFree(h->dAlpha); cudaFree(h->dGrad);
    cudaFree(h->dPhi); cudaFree(h->dPlo); cudaFree(h->dLhi); cudaFree(h->dLlo); cudaFree(h->dAthi); cudaFree(h->dAtlo);
    for (auto ev : h->ev) if (ev) cudaEventDestroy(ev);
    for (auto ev : h->ev_pool) cudaEventDestroy(ev);
    for (auto ev : h->ev_mark) if (ev) cudaEventDestroy(ev);
    if (h->fitc_u) gb2_destroy(h->fitc_u);
    if (h->fitc_b) gb2_destroy(h->fitc_b);
    cudaSetDevice(h->device);
    cudaFree(h->dFitc); cudaFree(h->dSt); cudaFree(h->dFitcScal);
    if (h->dTrace) cudaFree(h->dTrace);
    if (h->s_bulk2) cudaStreamDestroy(h->s_bulk2);
    for (auto st : h->s_aux) if (st) cudaStreamDestroy(st);
    for (auto ev : h->ev_join) if (ev) cudaEventDestroy(ev);
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    if (h->s_main) cudaStreamDestroy(h->s_main);
    if (h->s_panel) cudaStreamDestroy(h->s_panel);
    delete h;
    return 0;
}

int gb2_set_train(gb2_handle* h, const double* X, int64_t N, int32_t D_in, const double* y) {
    if (!h) return -1;
    return set_train_common(h, X, N, D_in, y, cudaMemcpyHostToDevice);
}

int gb2_set_train_dev(gb2_handle* h, const double* dX, int64_t N, int32_t D_in, const double* dy) {
    if (!h) return -1;
    return set_train_common(h, dX, N, D_in, dy, cudaMemcpyDeviceToDevice);
}

int gb2_set_kernel(gb2_handle* h, const gb2_kernel* k) {
    if (!h) return -1;
    GB2_ARG(h, k != nullptr, "kernel is null");
    GB2_ARG(h, k->n_terms >= 1 && k->n_terms <= GB2_MAX_TERMS, "n_terms must be in [1, GB2_MAX_TERMS]");
    GB2_ARG(h, std::isfinite(k->sigma), "sigma must be finite");
    GB2_ARG(h, std::isfinite(k->jitter) && k->jitter >= 0, "jitter must be finite and >= 0");
    KParams kp{};
    PrepParams pp{};
    std::vector<double> btab;
    kp.n_terms = pp.n_terms = k->n_terms;
    int n_feat = 0, n_cat = 0;
    auto cat_row = [&](int col, int P) -> int {
        for (int f = 0; f < n_cat; f++)
            if (pp.cat_col[f] == col) { pp.cat_P[f] = std::min(pp.cat_P[f], P); return f; }
        pp.cat_col[n_cat] = col; pp.cat_P[n_cat] = P;
        return n_cat++;
    };
    for (int t = 0; t < k->n_terms; t++) {
        const gb2_term& T = k->terms[t];
        TermDev& D = kp.t[t];
        GB2_ARG(h, T.kind >= GB2_EXPQUAD && T.kind <= GB2_EXPONENTIAL, "unknown continuous kernel kind");
        GB2_ARG(h, T.d >= 0 && T.d <= GB2_MAX_D, "d must be in [0, GB2_MAX_D]");
        GB2_ARG(h, T.n_lin >= 0 && T.n_lin <= GB2_MAX_LIN, "n_lin must be in [0, GB2_MAX_LIN]");
        GB2_ARG(h, T.n_coreg >= 0 && T.n_coreg <= GB2_MAX_COREG, "n_coreg must be in [0, GB2_MAX_COREG]");
        GB2_ARG(h, std::isfinite(T.eta) && std::isfinite(T.tau), "eta/tau must be finite");
        D.kind = T.kind; D.d = T.d; D.n_lin = T.n_lin; D.n_coreg = T.n_coreg;
        D.eta2 = T.eta * T.eta; D.tau = T.n_lin > 0 ? T.tau : 0.0;
        D.feat_off = n_feat;
        pp.d[t] = T.d; pp.n_lin[t] = T.n_lin; pp.feat_off[t] = n_feat;
        for (int i = 0; i < T.d; i++) {
            GB2_ARG(h, T.ls[i] > 0 && std::isfinite(T.ls[i]), "lengthscales must be positive and finite");
            GB2_ARG(h, T.cont_idx[i] >= 0, "negative cont_idx");
            pp.cont_idx[t][i] = T.cont_idx[i];
            pp.inv_ls[t][i] = 1.0 / T.ls[i];
        }
        for (int l = 0; l < T.n_lin; l++) {
            GB2_ARG(h, T.lin_idx[l] >= 0 && std::isfinite(T.c[l]), "bad linear kernel parameters");
            pp.lin_idx[t][l] = T.lin_idx[l];
            pp.c[t][l] = T.c[l];
        }
        n_feat += T.d + 1 + T.n_lin;
        for (int f = 0; f < T.n_coreg; f++) {
            GB2_ARG(h, T.coreg_P[f] >= 1 && T.coreg_P[f] <= GB2_MAX_P && T.coreg_B[f] && T.coreg_col[f] >= 0, "bad Coregion factor");
            D.cg_cat[f] = cat_row(T.coreg_col[f], T.coreg_P[f]);
            D.cg_P[f] = T.coreg_P[f];
            D.cg_Boff[f] = (int)btab.size();
            btab.insert(btab.end(), T.coreg_B[f], T.coreg_B[f] + T.coreg_P[f] * T.coreg_P[f]);
        }
    }
    kp.sigma2 = k->sigma * k->sigma;
    kp.jitter = k->jitter;
    kp.noise_cat = -1;
    if (k->noise_col >= 0) {
        GB2_ARG(h, k->noise_P >= 1 && k->noise_P <= GB2_MAX_P && k->noise_B, "bad noise Coregion");
        kp.noise_cat = cat_row(k->noise_col, k->noise_P);
        kp.noise_P = k->noise_P;
        kp.noise_Boff = (int)btab.size();
        btab.insert(btab.end(), k->noise_B, k->noise_B + k->noise_P * k->noise_P);
    }
    for (double b : btab) GB2_ARG(h, std::isfinite(b), "Coregion table holds a non-finite value");
    kp.n_feat = n_feat; kp.n_cat = n_cat;
    pp.n_cat = n_cat;
    GB2_ARG(h, kbuild_smem_bytes(kp) <= 160 * 1024 && kbuild_dmma_smem_bytes(kp) <= 160 * 1024, "too many features per point for the K-build tile");
    GB2_CUDA(h, cudaSetDevice(h->device));
    if (btab.empty()) btab.push_back(1.0);
    if ((int)btab.size() > h->btab_len) {
        if (h->dBtab) GB2_CUDA(h, cudaFree(h->dBtab));
        h->dBtab = nullptr;
        GB2_CUDA(h, cudaMalloc(&h->dBtab, btab.size() * sizeof(double)));
        h->btab_len = (int)btab.size();
    }
    GB2_CUDA(h, cudaMemcpyAsync(h->dBtab, btab.data(), btab.size() * sizeof(double), cudaMemcpyHostToDevice, h->s_main));
    GB2_CUDA(h, cudaStreamSynchronize(h->s_main));
    h->kp = kp; h->pp = pp;
    for (int t = 0; t < k->n_terms; t++) h->eta_host[t] = k->terms[t].eta;
    h->sigma_host = k->sigma;
    {   // deep copy of the description (the caller's Coregion tables need not outlive this call): the FITC inner systems re-use it
        h->kernel_host = *k;
        size_t total = 0;
        for (int t = 0; t < k->n_terms; t++)
            for (int c = 0; c < k->terms[t].n_coreg; c++) total += (size_t)k->terms[t].coreg_P[c] * k->terms[t].coreg_P[c];
        h->kernel_Bstore.assign(total, 0.0);
        size_t off = 0;
        for (int t = 0; t < k->n_terms; t++)
            for (int c = 0; c < k->terms[t].n_coreg; c++) {
                const size_t n = (size_t)k->terms[t].coreg_P[c] * k->terms[t].coreg_P[c];
                std::copy(k->terms[t].coreg_B[c], k->terms[t].coreg_B[c] + n, h->kernel_Bstore.begin() + off);
                h->kernel_host.terms[t].coreg_B[c] = h->kernel_Bstore.data() + off;
                off += n;
            }
        h->kernel_host.noise_B = nullptr;    // only consulted through noise_col (FITC refuses a noise Coregion)
    }
    h->have_kernel = true; h->factorized = false; h->fitc_ready = false;
    return 0;
}

// ---- peer-memory exchange setup (multi-GPU): IPC handles of the factor / diagonal-block / counter buffers are all-gathered
// once per (re)allocation and opened on every peer; afterwards the factorisation's data path is NVLink loads/stores only.
static void p2p_close(gb2_handle* h) {
    for (int r = 0; r < 8; r++) {
        if (r == h->rank) continue;
        if (h->peerA[r]) cudaIpcCloseMemHandle(h->peerA[r]);
        if (h->peerDinv[r]) cudaIpcCloseMemHandle(h->peerDinv[r]);
        if (h->peerLpack[r]) cudaIpcCloseMemHandle(h->peerLpack[r]);
        if (h->peerFlags[r]) cudaIpcCloseMemHandle(h->peerFlags[r]);
        if (h->peerRing[r]) cudaIpcCloseMemHandle(h->peerRing[r]);
        h->peerA[r] = h->peerDinv[r] = h->peerLpack[r] = h->peerRing[r] = nullptr;
        h->peerFlags[r] = nullptr;
    }
    h->p2p_ready = false;
}

// tiny collective used as a barrier / agreement: every rank contributes one int, returns the minimum
static int dist_min_int(gb2_handle* h, int mine, int* out) {
    int* d = reinterpret_cast<int*>(h->dIpcXch);
    GB2_CUDA(h, cudaMemcpyAsync(d + 64, &mine, sizeof(int), cudaMemcpyHostToDevice, h->s_main));
    const int rc = h->nccl->AllGather(d + 64, d, sizeof(int), 0 /*ncclInt8*/, h->comm, h->s_main);
    if (rc != 0) { h->err = std::string("ncclAllGather: ") + h->nccl->GetErrorString(rc); return -200 - rc; }
    int all[8];
    GB2_CUDA(h, cudaMemcpyAsync(all, d, h->world * sizeof(int), cudaMemcpyDeviceToHost, h->s_main));
    GB2_CUDA(h, cudaStreamSynchronize(h->s_main));
    int m = all[0];
    for (int r = 1; r < h->world; r++) m = std::min(m, all[r]);
    *out = m;
    return 0;
}

static int p2p_setup(gb2_handle* h) {
    const int G = h->world, me = h->rank;
    struct Pack { cudaIpcMemHandle_t a, dinv, lpack, flags, ring; };
    static_assert(sizeof(Pack) == 320, "five 64-byte IPC handles");
    Pack mine{};
    int ok = 1;
    if (cudaIpcGetMemHandle(&mine.a, h->dA) != cudaSuccess || cudaIpcGetMemHandle(&mine.dinv, h->dDinv) != cudaSuccess ||
        cudaIpcGetMemHandle(&mine.lpack, h->dLpack) != cudaSuccess || cudaIpcGetMemHandle(&mine.flags, h->dFlags) != cudaSuccess ||
        (h->dRing && cudaIpcGetMemHandle(&mine.ring, h->dRing) != cudaSuccess)) {
        ok = 0;
        cudaGetLastError();
    }
    char* d = h->dIpcXch;   // [0, 8*320) gathered handles, [8*320, +320) mine
    GB2_CUDA(h, cudaMemcpyAsync(d + 8 * 320, &mine, sizeof(Pack), cudaMemcpyHostToDevice, h->s_main));
    int rc = h->nccl->AllGather(d + 8 * 320, d, sizeof(Pack), 0 /*ncclInt8*/, h->comm, h->s_main);
    if (rc != 0) { h->err = std::string("ncclAllGather: ") + h->nccl->GetErrorString(rc); return -200 - rc; }
    Pack all[8];
    GB2_CUDA(h, cudaMemcpyAsync(all, d, (size_t)G * sizeof(Pack), cudaMemcpyDeviceToHost, h->s_main));
    GB2_CUDA(h, cudaStreamSynchronize(h->s_main));
    for (int r = 0; r < G && ok; r++) {
        if (r == me) {
            h->peerA[r] = h->dA; h->peerDinv[r] = h->dDinv; h->peerLpack[r] = h->dLpack; h->peerFlags[r] = h->dFlags; h->peerRing[r] = h->dRing;
            continue;
        }
        void *pa = nullptr, *pd = nullptr, *pl = nullptr, *pf = nullptr, *pg = nullptr;
        if (cudaIpcOpenMemHandle(&pa, all[r].a, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess ||
            cudaIpcOpenMemHandle(&pd, all[r].dinv, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess ||
            cudaIpcOpenMemHandle(&pl, all[r].lpack, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess ||
            cudaIpcOpenMemHandle(&pf, all[r].flags, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess ||
            (h->dRing && cudaIpcOpenMemHandle(&pg, all[r].ring, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess)) {
            ok = 0;
            cudaGetLastError();
        }
        h->peerA[r] = (double*)pa; h->peerDinv[r] = (double*)pd; h->peerLpack[r] = (double*)pl; h->peerFlags[r] = (unsigned*)pf;
        h->peerRing[r] = (double*)pg;
    }
    int all_ok = 0;
    if ((rc = dist_min_int(h, ok, &all_ok))) return rc;
    if (!all_ok) {   // some rank cannot map its peers (no P2P / IPC in this environment): everyone uses the NCCL exchange
        p2p_close(h);
        h->opt_p2p = 0;
        return 0;
    }
    h->p2p_ready = true;
    return 0;
}

static int build_K(gb2_handle* h, int& launches) {
    const int64_t Np = h->Np, N = h->N;
    cudaStream_t s = h->s_main;
    h->pp.D_in = h->D_in;
    int rc;
    if ((rc = validate_against_train(h))) return rc;
    if ((rc = ensure(h, h->dF, h->F_cap, (int64_t)std::max(1, h->kp.n_feat) * Np))) return rc;
    if ((rc = ensure(h, h->dC, h->C_cap, (int64_t)std::max(1, h->kp.n_cat) * Np))) return rc;
    if (h->world > 1 && !h->dIpcXch) GB2_CUDA(h, cudaMalloc(&h->dIpcXch, 16 * 320));
    const bool want_compact = h->world > 1 && h->opt_shard_storage;
    if (want_compact != h->compact) { h->A_cap = 0; h->xch_cap = 0; }   // storage layout changes: everything is re-allocated
    if (h->world > 1 && h->p2p_ready && (h->A_cap < Np || h->xch_cap < Np)) {
        // buffers are about to be re-allocated: every rank first drops its mappings of the peers' old buffers
        p2p_close(h);
        int dummy;
        if ((rc = dist_min_int(h, 1, &dummy))) return rc;
    }
    if (h->A_cap < Np) {
        if (h->dA) GB2_CUDA(h, cudaFree(h->dA));
        if (h->dDinv) GB2_CUDA(h, cudaFree(h->dDinv));
        h->dA = nullptr; h->dDinv = nullptr; h->A_cap = 0;
        h->compact = want_compact;
        h->nloc = want_compact ? (Np / TILE + h->world - 1) / h->world : Np / TILE;   // same allocation on every rank
        GB2_CUDA(h, cudaMalloc(&h->dA, (size_t)h->nloc * TILE * Np * sizeof(double)));
        GB2_CUDA(h, cudaMalloc(&h->dDinv, (size_t)Np * TILE * sizeof(double)));
        // the diagonal-panel kernel writes the lower block-triangle of each inverse only; the rest stays zero
        GB2_CUDA(h, cudaMemsetAsync(h->dDinv, 0, (size_t)Np * TILE * sizeof(double), s));
        h->A_cap = Np;
    }
    if (h->precision == GB2_TF32) {
        const int64_t pld = (int64_t)h->tf32_nb() * TILE;
        if (h->P_cap < Np * pld) {
            if (h->dPhi) GB2_CUDA(h, cudaFree(h->dPhi));
            if (h->dPlo) GB2_CUDA(h, cudaFree(h->dPlo));
            h->dPhi = h->dPlo = nullptr; h->P_cap = 0;
            GB2_CUDA(h, cudaMalloc(&h->dPhi, (size_t)Np * pld * sizeof(float)));
            GB2_CUDA(h, cudaMalloc(&h->dPlo, (size_t)Np * pld * sizeof(float)));
            h->P_cap = Np * pld;
        }
        int rc2;
        if ((rc2 = make_tmap_checked(h, &h->mPhi, h->dPhi, Np, pld, pld))) return rc2;
        if ((rc2 = make_tmap_checked(h, &h->mPlo, h->dPlo, Np, pld, pld))) return rc2;
    }
    if (h->world > 1 && h->xch_cap < Np) {
        // broadcast payload for the diagonal blocks + staging of the per-step panel allgather (<= ceil(nb/world) blocks per rank)
        if (h->dLpack) GB2_CUDA(h, cudaFree(h->dLpack));
        if (h->dSend) GB2_CUDA(h, cudaFree(h->dSend));
        if (h->dRecv) GB2_CUDA(h, cudaFree(h->dRecv));
        h->dLpack = h->dSend = h->dRecv = nullptr; h->xch_cap = 0;
        const int64_t nb = Np / TILE, per_rank = (nb + h->world - 1) / h->world;
        GB2_CUDA(h, cudaMalloc(&h->dLpack, (size_t)Np * TILE * sizeof(double)));
        GB2_CUDA(h, cudaMemsetAsync(h->dLpack, 0, (size_t)Np * TILE * sizeof(double), s));
        GB2_CUDA(h, cudaMalloc(&h->dSend, (size_t)per_rank * TILE * TILE * sizeof(double)));
        GB2_CUDA(h, cudaMalloc(&h->dRecv, (size_t)per_rank * h->world * TILE * TILE * sizeof(double)));
        GB2_CUDA(h, cudaMemsetAsync(h->dSend, 0, (size_t)per_rank * TILE * TILE * sizeof(double), s));
        if (h->dFlags) GB2_CUDA(h, cudaFree(h->dFlags));
        h->p2p_nbmax = nb;
        // [2 parities][2 kinds][nb] factorisation counters, 4 barrier words, [2 parities][nb] prediction-step counters
        GB2_CUDA(h, cudaMalloc(&h->dFlags, ((size_t)6 * nb + 4) * sizeof(unsigned)));
        GB2_CUDA(h, cudaMemsetAsync(h->dFlags, 0, ((size_t)6 * nb + 4) * sizeof(unsigned), s));
        h->p2p_parity = 0;
        h->x_parity = 0;
        h->p2p_epoch = 0;
        if (h->dRing) { GB2_CUDA(h, cudaFree(h->dRing)); h->dRing = nullptr; }
        if (want_compact) {
            // panel ring (factorisation: >= 3 live panels) doubling as the ring of solved prediction panels (a rank may lag
            // world-1 steps behind there): slots of max(Np, 16384) x 128 doubles
            h->ring_slots = std::max(4, h->world + 2);
            h->ring_slot_elems = std::max<int64_t>(Np, 16384) * TILE;
            GB2_CUDA(h, cudaMalloc(&h->dRing, (size_t)h->ring_slots * h->ring_slot_elems * sizeof(double)));
            if ((rc = ensure(h, h->dV, h->V_cap, Np))) return rc;
        }
        if ((rc = ensure(h, h->dPart, h->part_cap, std::max<int64_t>(64, (int64_t)2 * 16384 * (h->world + 1))))) return rc;
        h->xch_cap = Np;
    }
    if (h->world > 1 && h->opt_p2p && !h->p2p_ready) {
        GB2_CUDA(h, cudaStreamSynchronize(s));
        if ((rc = p2p_setup(h))) return rc;
    }
    GB2_ARG(h, !h->compact || h->p2p_ready, "shard_storage needs the NVLink peer exchange (CUDA IPC / P2P unavailable or p2p=0)");
    GB2_ARG(h, !h->compact || !h->opt_kbuild_v1, "shard_storage is not available with the kbuild_v1 ablation");
    if (h->world > 1 && h->p2p_ready) {
        // counters: this factorisation uses parity p; the other parity (used by the previous one, fully consumed) is cleared for
        // the next one now -- no peer can be that far ahead, every factorisation needs every rank's panels
        h->p2p_parity ^= 1;
        GB2_CUDA(h, cudaMemsetAsync(h->dFlags + (size_t)(h->p2p_parity ^ 1) * 2 * h->p2p_nbmax, 0, (size_t)2 * h->p2p_nbmax * sizeof(unsigned), s));
    }
    GB2_CUDA(h, cudaMemsetAsync(h->dInfo, 0, 4 * sizeof(int), s));
    GB2_CUDA(h, cudaEventRecord(h->ev[0], s));
    prep_features<<<(unsigned)((Np + 255) / 256), 256, 0, s>>>(h->dX, N, Np, h->pp, h->dF, h->dC, h->dInfo + 1);
    GB2_CUDA(h, cudaEventRecord(h->ev[1], s));
    dim3 grid((unsigned)(Np / KB_T), (unsigned)(Np / KB_T));
    if (h->opt_kbuild_v1)
        kbuild_kernel<true><<<grid, KB_THREADS, kbuild_smem_bytes(h->kp), s>>>(h->kp, h->dBtab, h->dF, h->dC, Np, N, h->dF, h->dC, Np, N, h->dy,
                                                                                h->dA, Np, h->world, h->rank);
    else
        kbuild_dmma_launch<true>(s, grid, kbuild_dmma_smem_bytes(h->kp), h->kp, h->dBtab, h->dF, h->dC, Np, N, h->dF, h->dC, Np, N, h->dy, h->dA, Np,
                                 h->world, h->rank, h->compact ? 1 : 0, h->opt_kbuild_occ, h->opt_kbuild_persist ? h->dKbCtr : nullptr, h->n_sm);
    GB2_CUDA(h, cudaEventRecord(h->ev[2], s));
    launches += 2;
    return 0;
}

// Prediction points of a fused cold predict (gb2_factorize_predict); M == 0: plain factorisation.
struct ExtPredict { const double* Xs = nullptr; int64_t M = 0; int32_t pred_noise = 0; double* mean = nullptr; double* var = nullptr; bool dev = false;
                    // fill: called after the K-build to overwrite the matrix that gets factorised (FITC B system)
                    int (*fill)(gb2_handle*, void*) = nullptr; void* fill_arg = nullptr; };

static int factorize_impl(gb2_handle* h, const ExtPredict& x) {
    GB2_ARG(h, h->have_train, "gb2_factorize: no training data (call gb2_set_train)");
    GB2_ARG(h, h->have_kernel, "gb2_factorize: no kernel (call gb2_set_kernel)");
    GB2_CUDA(h, cudaSetDevice(h->device));
    h->factorized = false;
    h->L_split_valid = false;
    int launches = 0, rc;
    if ((rc = build_K(h, launches))) return rc;
    if (x.fill && (rc = x.fill(h, x.fill_arg))) return rc;
    const int64_t Np = h->Np, N = h->N, Mp = round_up(x.M, TILE);
    cudaStream_t s = h->s_main;
    if (x.M > 0) {
        GB2_ARG(h, !h->compact, "gb2_factorize_predict is not available in the storage-sharded mode");
        GB2_ARG(h, h->precision == GB2_FP64, "gb2_factorize_predict is fp64 only");
        GB2_ARG(h, Mp * Np * (int64_t)sizeof(double) <= ((int64_t)8 << 30), "too many prediction points for one fused pass (use gb2_factorize + gb2_predict)");
        if ((rc = ensure(h, h->dAt, h->At_cap, Mp * Np))) return rc;
        if ((rc = ensure(h, h->dFs, h->Fs_cap, (int64_t)std::max(1, h->kp.n_feat) * Mp))) return rc;
        if ((rc = ensure(h, h->dCs, h->Cs_cap, (int64_t)std::max(1, h->kp.n_cat) * Mp))) return rc;
        const double* dXs = x.Xs;
        if (!x.dev) {
            if ((rc = ensure(h, h->dXs, h->Xs_cap, x.M * h->D_in))) return rc;
            int64_t oc = h->out_cap;
            if ((rc = ensure(h, h->dMean, oc, x.M))) return rc;
            if ((rc = ensure(h, h->dVar, h->out_cap, x.M))) return rc;
            GB2_CUDA(h, cudaMemcpyAsync(h->dXs, x.Xs, (size_t)x.M * h->D_in * sizeof(double), cudaMemcpyHostToDevice, s));
            dXs = h->dXs;
        }
        // K(X*, X) as rows below the factor, exactly as gb2_predict builds it
        prep_features<<<(unsigned)((Mp + 255) / 256), 256, 0, s>>>(dXs, x.M, Mp, h->pp, h->dFs, h->dCs, h->dInfo + 1);
        dim3 grid((unsigned)(Np / KB_T), (unsigned)(Mp / KB_T));
        kbuild_dmma_launch<false>(s, grid, kbuild_dmma_smem_bytes(h->kp), h->kp, h->dBtab, h->dFs, h->dCs, Mp, x.M, h->dF, h->dC, Np, N, nullptr,
                                  h->dAt, Np, 1, 0, 0, h->opt_kbuild_occ, h->opt_kbuild_persist ? h->dKbCtr : nullptr, h->n_sm);
        GB2_CUDA(h, cudaEventRecord(h->ev[2], s));   // "K build" now covers K and K*
        launches += 2;
        h->ext_At = h->dAt; h->ext_rows = Mp; h->ext_ld = Np; h->ext_ncols = (int)((N + TILE - 1) / TILE);
    }
    launches += cholesky_enqueue(h);
    h->ext_At = nullptr; h->ext_rows = 0;
    GB2_CUDA(h, cudaEventRecord(h->ev[3], s));
    if (x.M > 0) {
        posterior_reduce_kernel<<<(unsigned)((x.M + 7) / 8), 256, 0, s>>>(h->kp, h->dBtab, h->dFs, h->dCs, Mp, h->dAt, Np, h->dA + N * Np, N, x.M,
                                                                         x.pred_noise, x.dev ? x.mean : h->dMean, x.dev ? x.var : h->dVar);
        launches++;
        GB2_CUDA(h, cudaEventRecord(h->ev[4], s));
        if (!x.dev) {
            GB2_CUDA(h, cudaMemcpyAsync(x.mean, h->dMean, (size_t)x.M * sizeof(double), cudaMemcpyDeviceToHost, s));
            GB2_CUDA(h, cudaMemcpyAsync(x.var, h->dVar, (size_t)x.M * sizeof(double), cudaMemcpyDeviceToHost, s));
        }
    }
    int info[2] = {0, 0};
    GB2_CUDA(h, cudaMemcpyAsync(info, h->dInfo, 2 * sizeof(int), cudaMemcpyDeviceToHost, s));
    GB2_CUDA(h, cudaGetLastError());
    GB2_CUDA(h, cudaStreamSynchronize(s));
    h->timings[0] = ms_between(h->ev[0], h->ev[1]);
    h->timings[1] = ms_between(h->ev[1], h->ev[2]);
    h->timings[2] = ms_between(h->ev[2], h->ev[3]);
    h->timings[6] = launches;
    if (x.M > 0) { h->timings[3] = 0; h->timings[4] = 0; h->timings[5] = ms_between(h->ev[3], h->ev[4]); h->timings[7] = 0; }
    if (h->world > 1) {
        // the first failing pivot is only known to the rank that owns its diagonal block: agree on one verdict, otherwise some ranks
        // would raise while the others walk into the next collective.  Encoded so that "smallest positive pivot" is a minimum.
        // one integer per rank: 1 = bad level index, pivot + 1 = first failing pivot, INT_MAX = fine
        const int mine = info[1] != 0 ? 1 : (info[0] > 0 ? info[0] + 1 : 0x7fffffff);
        int agreed = 0, rcq;
        if ((rcq = dist_min_int(h, mine, &agreed))) return rcq;
        info[1] = agreed == 1 ? 1 : 0;
        info[0] = (agreed > 1 && agreed != 0x7fffffff) ? agreed - 1 : 0;
    }
    GB2_ARG(h, info[1] == 0, "a Coregion column of X (or Xs) holds a level index outside [0, P)");
    if (info[0] != 0) {
        h->err = "matrix is not positive definite: leading minor of order " + std::to_string(info[0]);
        return info[0];
    }
    h->factorized = true;
    h->factor_count++;
    return 0;
}

int gb2_factorize(gb2_handle* h) {
    if (!h) return -1;
    return factorize_impl(h, ExtPredict{});
}

// One reference predict call in one pass (PymcGP.predict rebuilds K, re-factorises and solves on every call,
// gumbi/regression/pymc/GP.py:845-847): the prediction points are appended as extra rows of the factor, so the solve
// A^T = K(X*,X) L^-T is carried by the factorisation's own panel solves and right-looking updates (cholesky.cuh, factor_steps).
// Leaves the handle factorised exactly as gb2_factorize does.  Multi-GPU (replicated storage): collective, every rank passes its
// own points.
int gb2_factorize_predict(gb2_handle* h, const double* Xs, int64_t M, int32_t pred_noise, double* mean, double* var) {
    if (!h) return -1;
    GB2_ARG(h, Xs && mean && var, "null pointer");
    GB2_ARG(h, M >= 1, "M must be >= 1");
    ExtPredict x;
    x.Xs = Xs; x.M = M; x.pred_noise = pred_noise; x.mean = mean; x.var = var;
    return factorize_impl(h, x);
}

int gb2_factorize_predict_dev(gb2_handle* h, const double* dXs, int64_t M, int32_t pred_noise, double* dmean, double* dvar) {
    if (!h) return -1;
    GB2_ARG(h, dXs && dmean && dvar, "null pointer");
    GB2_ARG(h, M >= 1, "M must be >= 1");
    ExtPredict x;
    x.Xs = dXs; x.M = M; x.pred_noise = pred_noise; x.mean = dmean; x.var = dvar; x.dev = true;
    return factorize_impl(h, x);
}

int gb2_mll(gb2_handle* h, double* out) {
    if (!h) return -1;
    GB2_ARG(h, out, "out is null");
    GB2_ARG(h, h->factorized, "gb2_mll called before a successful gb2_factorize");
    GB2_CUDA(h, cudaSetDevice(h->device));
    double sc[2];
    GB2_CUDA(h, cudaMemcpy(sc, h->dScal, 2 * sizeof(double), cudaMemcpyDeviceToHost));
    *out = -0.5 * (double)h->N * 1.8378770664093454836 - sc[0] - 0.5 * sc[1];
    return 0;
}

int gb2_mll_grad(gb2_handle* h, double* mll_out, double* grad_out) {
    if (!h) return -1;
    GB2_ARG(h, mll_out && grad_out, "null pointer");
    GB2_ARG(h, h->factorized, "gb2_mll_grad called before a successful gb2_factorize");
    GB2_ARG(h, !h->compact, "gb2_mll_grad is not available in the storage-sharded mode yet");
    GB2_CUDA(h, cudaSetDevice(h->device));
    const int64_t Np = h->Np, N = h->N;
    const int nb = (int)(Np / TILE);
    cudaStream_t s = h->s_main;
    if (h->G_cap < Np) {
        if (h->dW) GB2_CUDA(h, cudaFree(h->dW));
        if (h->dS) GB2_CUDA(h, cudaFree(h->dS));
        h->dW = h->dS = nullptr; h->G_cap = 0;
        GB2_CUDA(h, cudaMalloc(&h->dW, (size_t)Np * Np * sizeof(double)));
        GB2_CUDA(h, cudaMalloc(&h->dS, (size_t)Np * Np * sizeof(double)));
        h->G_cap = Np;
    }
    int rc;
    if ((rc = ensure(h, h->dAlpha, h->alpha_cap, Np))) return rc;
    if (!h->dGrad) GB2_CUDA(h, cudaMalloc(&h->dGrad, GR_LEN * sizeof(double)));
    int launches = 0;
    // 1. W = L^-T (rows), alpha from the augmented column
    set_identity_kernel<<<(unsigned)((Np * (Np / 2) + 255) / 256), 256, 0, s>>>(h->dW, Np, Np);
    trsm_rec(s, h->dA, Np, h->dDinv, h->dW, Np, Np, 0, nb, launches, /*tri=*/true);
    extract_alpha_kernel<<<(unsigned)((Np + 255) / 256), 256, 0, s>>>(h->dW, Np, N, Np, h->dAlpha);
    // 2. S = W W^T = K^-1 (lower tiles)
    dgemm_nt_launch<128, 64, GM_SET>(s, h->dW, Np, h->dW, Np, h->dS, Np, Np, Np, (int)Np, /*lower, triangular k-range*/ 2, 0, 0);
    // 3. one pass over G = alpha alpha^T - S
    GB2_CUDA(h, cudaMemsetAsync(h->dGrad, 0, GR_LEN * sizeof(double), s));
    dim3 grid((unsigned)(Np / KB_T), (unsigned)(Np / KB_T));
    mll_grad_kernel<<<grid, KB_THREADS, mll_grad_smem_bytes(h->kp), s>>>(h->kp, h->dBtab, h->dF, h->dC, Np, N, h->dS, Np, h->dAlpha, h->dGrad);
    GB2_CUDA(h, cudaGetLastError());
    GB2_CUDA(h, cudaMemcpyAsync(grad_out, h->dGrad, GR_LEN * sizeof(double), cudaMemcpyDeviceToHost, s));
    double sc[2];
    GB2_CUDA(h, cudaMemcpyAsync(sc, h->dScal, 2 * sizeof(double), cudaMemcpyDeviceToHost, s));
    GB2_CUDA(h, cudaStreamSynchronize(s));
    *mll_out = -0.5 * (double)N * 1.8378770664093454836 - sc[0] - 0.5 * sc[1];
    // host-side constant factors: the kernel accumulated sum_ij G_ij * (dK_ij/dtheta without these factors)
    for (int t = 0; t < GB2_MAX_TERMS; t++) {
        double* g = grad_out + t * GR_TERM;
        if (t >= h->kp.n_terms) continue;
        for (int k = 0; k < h->pp.d[t]; k++) g[GR_LS + k] *= 0.5 * h->pp.inv_ls[t][k];
        g[GR_ETA] *= 0.5 * 2.0 * h->eta_host[t];
        for (int l = 0; l < GB2_MAX_LIN; l++) g[GR_C + l] *= 0.5;
        g[GR_TAU] *= 0.5;
        for (int e = 0; e < GB2_MAX_COREG * GB2_MAX_P * GB2_MAX_P; e++) g[GR_B + e] *= 0.5;
    }
    grad_out[GR_SIGMA] *= 0.5 * 2.0 * h->sigma_host;
    for (int e = 0; e < GB2_MAX_P * GB2_MAX_P; e++) grad_out[GR_NOISE_B + e] *= 0.5;
    h->timings[6] = launches + 5;
    h->alpha_for = h->factor_count;
    return 0;
}

int gb2_predict(gb2_handle* h, const double* Xs, int64_t M, int32_t pred_noise, double* mean, double* var) {
    if (!h) return -1;
    return predict_common(h, Xs, M, pred_noise, mean, var, false);
}

// Posterior mean and FULL covariance: what gp.conditional(name, Xnew) parameterises (GP.py:913-914):
//   mu = A^T v,  cov = K(X*,X*) - A^T A  (+ noise on the diagonal if pred_noise), A = L^-1 K(X,X*).
// cov_out is (M, M) row-major, symmetric (both triangles filled).  M is limited by the (M x M) + (M x Np) device buffers.
int gb2_predict_full(gb2_handle* h, const double* Xs, int64_t M, int32_t pred_noise, double* mean, double* cov_out) {
    if (!h) return -1;
    GB2_ARG(h, h->factorized, "gb2_predict_full called before a successful gb2_factorize");
    GB2_ARG(h, !h->compact, "gb2_predict_full is not available in the storage-sharded mode");
    GB2_ARG(h, Xs && mean && cov_out, "null pointer");
    GB2_ARG(h, M >= 1 && M <= 32768, "M must be in [1, 32768] for the full-covariance prediction");
    GB2_CUDA(h, cudaSetDevice(h->device));
    const int64_t Np = h->Np, N = h->N, Mp = round_up(M, TILE);
    const int ncols = (int)((N + TILE - 1) / TILE);
    cudaStream_t s = h->s_main;
    int rc;
    if ((rc = ensure(h, h->dAt, h->At_cap, Mp * Np))) return rc;
    if ((rc = ensure(h, h->dFs, h->Fs_cap, (int64_t)std::max(1, h->kp.n_feat) * Mp))) return rc;
    if ((rc = ensure(h, h->dCs, h->Cs_cap, (int64_t)std::max(1, h->kp.n_cat) * Mp))) return rc;
    if ((rc = ensure(h, h->dXs, h->Xs_cap, M * h->D_in))) return rc;
    if ((rc = ensure(h, h->dCov, h->cov_cap, Mp * Mp))) return rc;
    int64_t oc = h->out_cap;
    if ((rc = ensure(h, h->dMean, oc, M))) return rc;
    if ((rc = ensure(h, h->dVar, h->out_cap, M))) return rc;
    GB2_CUDA(h, cudaMemcpyAsync(h->dXs, Xs, (size_t)M * h->D_in * sizeof(double), cudaMemcpyHostToDevice, s));
    GB2_CUDA(h, cudaMemsetAsync(h->dInfo + 1, 0, sizeof(int), s));
    int launches = 0;
    prep_features<<<(unsigned)((Mp + 255) / 256), 256, 0, s>>>(h->dXs, M, Mp, h->pp, h->dFs, h->dCs, h->dInfo + 1);
    // A^T (rows = prediction points), exactly as gb2_predict
    dim3 grid((unsigned)(Np / KB_T), (unsigned)(Mp / KB_T));
    kbuild_dmma_launch<false>(s, grid, kbuild_dmma_smem_bytes(h->kp), h->kp, h->dBtab, h->dFs, h->dCs, Mp, M, h->dF, h->dC, Np, N, nullptr, h->dAt, Np, 1, 0, 0, 4, h->opt_kbuild_persist ? h->dKbCtr : nullptr, h->n_sm);
    trsm_rec(s, h->dA, Np, h->dDinv, h->dAt, Np, Mp, 0, ncols, launches);
    // K(X*, X*) (all tiles) and cov -= At At^T over the training columns only (depth = ncols * 128: the augmented / padding
    // columns right of N inside the last block are zero in At by construction of K(X*,X) and stay zero through the solve)
    dim3 g2((unsigned)(Mp / KB_T), (unsigned)(Mp / KB_T));
    kbuild_dmma_launch<false>(s, g2, kbuild_dmma_smem_bytes(h->kp), h->kp, h->dBtab, h->dFs, h->dCs, Mp, M, h->dFs, h->dCs, Mp, M, nullptr, h->dCov, Mp, 1, 0, 0, 4, h->opt_kbuild_persist ? h->dKbCtr : nullptr, h->n_sm);
    mask_columns_kernel<<<(unsigned)((Mp + 255) / 256), 256, 0, s>>>(h->dAt, Np, Mp, N, (int64_t)ncols * TILE);
    dgemm_nt_launch<128, 64, GM_SUB>(s, h->dAt, Np, h->dAt, Np, h->dCov, Mp, Mp, Mp, ncols * TILE, 0, 0, 0);
    posterior_reduce_kernel<<<(unsigned)((M + 7) / 8), 256, 0, s>>>(h->kp, h->dBtab, h->dFs, h->dCs, Mp, h->dAt, Np, h->dA + N * Np, N, M, pred_noise,
                                                                   h->dMean, h->dVar);
    if (pred_noise) add_noise_diag_kernel<<<(unsigned)((M + 255) / 256), 256, 0, s>>>(h->kp, h->dBtab, h->dFs, h->dCs, Mp, M, h->dCov, Mp);
    GB2_CUDA(h, cudaGetLastError());
    GB2_CUDA(h, cudaMemcpyAsync(mean, h->dMean, (size_t)M * sizeof(double), cudaMemcpyDeviceToHost, s));
    GB2_CUDA(h, cudaMemcpy2DAsync(cov_out, M * sizeof(double), h->dCov, Mp * sizeof(double), M * sizeof(double), M, cudaMemcpyDeviceToHost, s));
    int bad = 0;
    GB2_CUDA(h, cudaMemcpyAsync(&bad, h->dInfo + 1, sizeof(int), cudaMemcpyDeviceToHost, s));
    GB2_CUDA(h, cudaStreamSynchronize(s));
    GB2_ARG(h, bad == 0, "a Coregion column of Xs holds a level index outside [0, P)");
    return 0;
}

// ---- sparse FITC approximation (fitc.cuh) -------------------------------------------------------------------------------------
static int fitc_fill_B(gb2_handle* hb, void* arg) {
    gb2_handle* h = static_cast<gb2_handle*>(arg);
    const int64_t Npb = hb->Np, Nk = round_up(h->N, TILE);
    cudaStream_t s = hb->s_main;
    // [B - I, b ; b^T, .] = S S^T over the training points (lower tiles), in the augmented layout the Cholesky expects
    dgemm_nt_launch<128, 64, GM_SET>(s, h->dSt, Nk, h->dSt, Nk, hb->dA, Npb, Npb, Npb, (int)Nk, /*lower*/ 1, 0, 0);
    fitc_fix_diag_kernel<<<(unsigned)((Npb + 255) / 256), 256, 0, s>>>(hb->dA, Npb, h->fitc_m, Npb);
    GB2_CUDA(hb, cudaGetLastError());
    return 0;
}

static int fitc_inner(gb2_handle* h, gb2_handle*& inner) {
    if (inner) return 0;
    if (gb2_create(&inner, h->device, GB2_FP64) != 0) { h->err = "FITC: " + g_create_err; inner = nullptr; return -1; }
    return 0;
}

// Factorise the FITC approximation for the training set / kernel of `h` with inducing points Xu (m x D_in, host).
// Returns 0, > 0 (the inducing system Kuu + jitter I is not positive definite: first failing pivot) or < 0 (error).
// Replaces MarginalApprox.marginal_likelihood(X, Xu, y, sigma) as built at gumbi/regression/pymc/GP.py:571-578.
int gb2_fitc_factorize(gb2_handle* h, const double* Xu, int64_t m) {
    if (!h) return -1;
    GB2_ARG(h, Xu, "Xu is null");
    GB2_ARG(h, h->have_train && h->have_kernel, "gb2_fitc_factorize needs training data and a kernel");
    GB2_ARG(h, h->world == 1, "the FITC approximation is single-GPU");
    GB2_ARG(h, h->kernel_host.noise_col < 0, "the FITC approximation takes a scalar noise sigma (GP.py:573-577): no noise Coregion");
    GB2_ARG(h, m >= 1 && m <= 16384, "the number of inducing points must be in [1, 16384]");
    const int64_t N = h->N, Npu = round_up(m + 1, TILE), Nk = round_up(N, TILE);
    GB2_ARG(h, Nk * Npu * (int64_t)sizeof(double) <= ((int64_t)8 << 30), "N x n_u too large for one FITC pass (K(X,Xu) Luu^-T must fit in 8 GiB)");
    h->fitc_ready = false;
    int rc;
    if ((rc = fitc_inner(h, h->fitc_u)) || (rc = fitc_inner(h, h->fitc_b))) return rc;
    gb2_handle *hu = h->fitc_u, *hb = h->fitc_b;
    auto inner_fail = [&](gb2_handle* in, const char* what, int code) { h->err = std::string("FITC ") + what + ": " + in->err; return code; };
    std::vector<double> zeros((size_t)m, 0.0);
    gb2_kernel ku = h->kernel_host;
    ku.noise_col = -1; ku.noise_P = 0;
    ku.sigma = 0.0;                      // stabilize(Kuu): jitter only
    if ((rc = gb2_set_train(hu, Xu, m, h->D_in, zeros.data()))) return inner_fail(hu, "inducing system", rc);
    if ((rc = gb2_set_kernel(hu, &ku))) return inner_fail(hu, "inducing system", rc);
    if ((rc = gb2_factorize(hu))) return inner_fail(hu, "inducing system (Kuu + jitter I)", rc);
    GB2_CUDA(h, cudaSetDevice(h->device));
    if ((rc = ensure(h, h->dFitc, h->fitc_cap, 3 * Nk))) return rc;
    if ((rc = ensure(h, h->dSt, h->St_cap, Npu * Nk))) return rc;
    if (!h->dFitcScal) GB2_CUDA(h, cudaMalloc(&h->dFitcScal, 2 * sizeof(double)));
    double* dLam = h->dFitc; double* dTmp = h->dFitc + Nk; double* dVar1 = h->dFitc + 2 * Nk;
    // A^T = K(X,Xu) Luu^-T  (rows = training points) and diag(Kff) - colsum(A*A): an exact-GP prediction of the inducing system at X
    if ((rc = predict_common(hu, h->dX, N, 0, dTmp, dVar1, true))) return inner_fail(hu, "K(X,Xu) solve", rc);
    cudaStream_t s = hu->s_main;
    const double sigma2 = h->kernel_host.sigma * h->kernel_host.sigma;
    fitc_lambda_kernel<<<1, 1024, 0, s>>>(dVar1, h->dy, N, sigma2, dLam, h->dFitcScal);
    fitc_scale_transpose_kernel<<<dim3((unsigned)(Nk / 32), (unsigned)(Npu / 32)), 256, 0, s>>>(hu->dAt, Npu, dLam, h->dy, N, m, h->dSt, Nk);
    GB2_CUDA(h, cudaGetLastError());
    GB2_CUDA(h, cudaStreamSynchronize(s));
    // B system: same shape as the inducing system; its K-build only sizes the buffers, fitc_fill_B overwrites the matrix
    h->fitc_m = m;
    if ((rc = gb2_set_train(hb, Xu, m, h->D_in, zeros.data()))) return inner_fail(hb, "B system", rc);
    if ((rc = gb2_set_kernel(hb, &ku))) return inner_fail(hb, "B system", rc);
    ExtPredict x;
    x.fill = fitc_fill_B; x.fill_arg = h;
    if ((rc = factorize_impl(hb, x))) return inner_fail(hb, "B system (I + A Lambda^-1 A^T)", rc);
    h->fitc_ready = true;
    return 0;
}

// logp of MarginalApprox(approx="FITC").marginal_likelihood at the current kernel (needs gb2_fitc_factorize).
int gb2_fitc_mll(gb2_handle* h, double* out) {
    if (!h) return -1;
    GB2_ARG(h, out, "out is null");
    GB2_ARG(h, h->fitc_ready, "gb2_fitc_mll called before a successful gb2_fitc_factorize");
    GB2_CUDA(h, cudaSetDevice(h->device));
    double sc[2], fs[2];
    GB2_CUDA(h, cudaMemcpy(sc, h->fitc_b->dScal, 2 * sizeof(double), cudaMemcpyDeviceToHost));    // sum log diag(L_B), c^T c
    GB2_CUDA(h, cudaMemcpy(fs, h->dFitcScal, 2 * sizeof(double), cudaMemcpyDeviceToHost));        // sum log Lambda, y^T Lambda^-1 y
    *out = -0.5 * (double)h->N * 1.8378770664093454836 - 0.5 * fs[0] - sc[0] - 0.5 * (fs[1] - sc[1]);
    return 0;
}

// Posterior mean / variance of the FITC conditional (MarginalApprox._build_conditional, diag=True) at M host points.
int gb2_fitc_predict(gb2_handle* h, const double* Xs, int64_t M, int32_t pred_noise, double* mean, double* var) {
    if (!h) return -1;
    GB2_ARG(h, h->fitc_ready, "gb2_fitc_predict called before a successful gb2_fitc_factorize");
    GB2_ARG(h, Xs && mean && var, "null pointer");
    GB2_ARG(h, M >= 0, "M must be >= 0");
    if (M == 0) return 0;
    GB2_CUDA(h, cudaSetDevice(h->device));
    gb2_handle *hu = h->fitc_u, *hb = h->fitc_b;
    const int64_t m = h->fitc_m, Npu = hu->Np;
    const int ncols_b = (int)((m + TILE - 1) / TILE);
    const int64_t chunk = std::min<int64_t>(round_up(M, TILE), std::max<int64_t>(TILE, ((int64_t)4 << 30) / (Npu * (int64_t)sizeof(double)) / TILE * TILE));
    int rc;
    if ((rc = ensure(h, h->dXs, h->Xs_cap, M * h->D_in))) return rc;
    int64_t oc = h->out_cap;
    if ((rc = ensure(h, h->dMean, oc, M))) return rc;
    if ((rc = ensure(h, h->dVar, h->out_cap, M))) return rc;
    if ((rc = ensure(h, h->dFitc, h->fitc_cap, std::max<int64_t>(3 * round_up(h->N, TILE), 3 * chunk)))) return rc;   // Lambda is no longer needed
    cudaStream_t s = hu->s_main;
    GB2_CUDA(h, cudaMemcpyAsync(h->dXs, Xs, (size_t)M * h->D_in * sizeof(double), cudaMemcpyHostToDevice, s));
    GB2_CUDA(h, cudaStreamSynchronize(s));
    const double add = pred_noise ? h->kernel_host.sigma * h->kernel_host.sigma : 0.0;
    double* dTmp = h->dFitc; double* dVar1 = h->dFitc + chunk;
    int launches = 0;
    for (int64_t m0 = 0; m0 < M; m0 += chunk) {
        const int64_t Mc = std::min(chunk, M - m0), Mp = round_up(Mc, TILE);
        // As^T = K(X*,Xu) Luu^-T in the inducing system's solve panel, var1 = kss - colsum(As*As)
        if ((rc = predict_common(hu, h->dXs + m0 * h->D_in, Mc, 0, dTmp, dVar1, true))) { h->err = "FITC K(X*,Xu) solve: " + hu->err; return rc; }
        mask_columns_kernel<<<(unsigned)((Mp + 255) / 256), 256, 0, s>>>(hu->dAt, Npu, Mp, m, Npu);
        trsm_rec(s, hb->dA, Npu, hb->dDinv, hu->dAt, Npu, Mp, 0, ncols_b, launches);      // C^T = As^T L_B^-T
        fitc_reduce_kernel<<<(unsigned)((Mc + 7) / 8), 256, 0, s>>>(hu->dAt, Npu, hb->dA + m * Npu, m, Mc, dVar1, add, h->dMean + m0, h->dVar + m0);
        GB2_CUDA(h, cudaGetLastError());
    }
    GB2_CUDA(h, cudaMemcpyAsync(mean, h->dMean, (size_t)M * sizeof(double), cudaMemcpyDeviceToHost, s));
    GB2_CUDA(h, cudaMemcpyAsync(var, h->dVar, (size_t)M * sizeof(double), cudaMemcpyDeviceToHost, s));
    GB2_CUDA(h, cudaStreamSynchronize(s));
    return 0;
}

int gb2_predict_dev(gb2_handle* h, const double* dXs, int64_t M, int32_t pred_noise, double* dmean, double* dvar) {
    if (!h) return -1;
    return predict_common(h, dXs, M, pred_noise, dmean, dvar, true);
}

int gb2_get_K(gb2_handle* h, double* K_out) {
    if (!h) return -1;
    GB2_ARG(h, K_out, "null pointer");
    GB2_ARG(h, h->have_train && h->have_kernel, "gb2_get_K needs training data and a kernel");
    GB2_ARG(h, !(h->world > 1 && h->opt_shard_storage), "gb2_get_K is a single-GPU test hook");
    GB2_CUDA(h, cudaSetDevice(h->device));
    int launches = 0, rc;
    h->factorized = false;  // the factor storage is overwritten
    if ((rc = build_K(h, launches))) return rc;
    const int64_t N = h->N, Np = h->Np;
    GB2_CUDA(h, cudaMemcpy2DAsync(K_out, N * sizeof(double), h->dA, Np * sizeof(double), N * sizeof(double), N, cudaMemcpyDeviceToHost, h->s_main));
    GB2_CUDA(h, cudaStreamSynchronize(h->s_main));
    for (int64_t i = 0; i < N; i++)
        for (int64_t j = i + 1; j < N; j++) K_out[i * N + j] = K_out[j * N + i];
    return 0;
}

int gb2_get_L(gb2_handle* h, double* L_out) {
    if (!h) return -1;
    GB2_ARG(h, L_out, "null pointer");
    GB2_ARG(h, h->factorized, "gb2_get_L called before a successful gb2_factorize");
    GB2_CUDA(h, cudaSetDevice(h->device));
    if (h->compact) {   // storage-sharded: this rank's row blocks, zeros elsewhere (the caller sums over ranks)
        const int64_t N = h->N, Np = h->Np;
        memset(L_out, 0, (size_t)N * N * sizeof(double));
        for (int64_t b = h->rank; b * TILE < N; b += h->world) {
            const int64_t rows = std::min<int64_t>(TILE, N - b * TILE);
            GB2_CUDA(h, cudaMemcpy2D(L_out + b * TILE * N, N * sizeof(double), h->dA + ((b - h->rank) / h->world) * TILE * Np, Np * sizeof(double),
                                     N * sizeof(double), rows, cudaMemcpyDeviceToHost));
        }
        for (int64_t i = 0; i < N; i++)
            for (int64_t j = i + 1; j < N; j++) L_out[i * N + j] = 0.0;
        return 0;
    }
    const int64_t N = h->N, Np = h->Np;
    GB2_CUDA(h, cudaMemcpy2D(L_out, N * sizeof(double), h->dA, Np * sizeof(double), N * sizeof(double), N, cudaMemcpyDeviceToHost));
    for (int64_t i = 0; i < N; i++)
        for (int64_t j = i + 1; j < N; j++) L_out[i * N + j] = 0.0;
    return 0;
}

int gb2_get_v(gb2_handle* h, double* v_out) {
    if (!h) return -1;
    GB2_ARG(h, v_out, "null pointer");
    GB2_ARG(h, h->factorized, "gb2_get_v called before a successful gb2_factorize");
    GB2_CUDA(h, cudaSetDevice(h->device));
    if (h->compact) {
        GB2_CUDA(h, cudaMemcpy(v_out, h->dV, (size_t)h->N * sizeof(double), cudaMemcpyDeviceToHost));
        return 0;
    }
    GB2_CUDA(h, cudaMemcpy(v_out, h->dA + h->N * h->Np, (size_t)h->N * sizeof(double), cudaMemcpyDeviceToHost));
    return 0;
}

int gb2_get_alpha(gb2_handle* h, double* alpha_out) {
    if (!h) return -1;
    GB2_ARG(h, alpha_out, "null pointer");
    GB2_ARG(h, h->factorized && h->alpha_for != 0 && h->alpha_for == h->factor_count,
            "gb2_get_alpha needs a gb2_mll_grad call on the current factorisation");
    GB2_CUDA(h, cudaSetDevice(h->device));
    GB2_CUDA(h, cudaMemcpy(alpha_out, h->dAlpha, (size_t)h->N * sizeof(double), cudaMemcpyDeviceToHost));
    return 0;
}

int gb2_get_trace(gb2_handle* h, uint64_t* out, int64_t n) {
    if (!h) return -1;
    GB2_ARG(h, out && n >= 0, "invalid argument");
    GB2_ARG(h, h->dTrace, "gb2_get_trace: set_option(\"trace\", 1) first");
    GB2_ARG(h, n <= h->trace_cap, "gb2_get_trace: at most 6 * 4096 stamps");
    GB2_CUDA(h, cudaSetDevice(h->device));
    GB2_CUDA(h, cudaMemcpy(out, h->dTrace, (size_t)n * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    return 0;
}

int gb2_get_timings(gb2_handle* h, double* out) {
    if (!h || !out) return -1;
    for (int i = 0; i < GB2_N_TIMINGS; i++) out[i] = h->timings[i];
    return 0;
}

int gb2_mark(gb2_handle* h, int slot) {
    if (!h) return -1;
    GB2_ARG(h, slot >= 0 && slot < 4, "mark slot must be in [0, 4)");
    GB2_CUDA(h, cudaSetDevice(h->device));
    if (!h->ev_mark[slot]) GB2_CUDA(h, cudaEventCreate(&h->ev_mark[slot]));
    GB2_CUDA(h, cudaEventRecord(h->ev_mark[slot], h->s_main));
    return 0;
}

int gb2_elapsed_ms(gb2_handle* h, int a, int b, double* ms) {
    if (!h) return -1;
    GB2_ARG(h, ms && a >= 0 && a < 4 && b >= 0 && b < 4 && h->ev_mark[a] && h->ev_mark[b], "bad marks");
    GB2_CUDA(h, cudaSetDevice(h->device));
    GB2_CUDA(h, cudaEventSynchronize(h->ev_mark[b]));
    float f = 0.f;
    GB2_CUDA(h, cudaEventElapsedTime(&f, h->ev_mark[a], h->ev_mark[b]));
    *ms = (double)f;
    return 0;
}

// ---- multi-GPU: NCCL bound at run time ---------------------------------------------------------------------------
static NcclApi g_nccl;
static std::string g_nccl_err;

static const NcclApi* load_nccl() {
    if (g_nccl.lib) return &g_nccl;
    // prefer a copy that is already mapped (torch.distributed's), then the usual sonames
    void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
    if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) { g_nccl_err = std::string("cannot load libnccl.so.2: ") + dlerror(); return nullptr; }
    NcclApi a;
    a.lib = lib;
#define GB2_SYM(field, name)                                                                          \
    *(void**)(&a.field) = dlsym(lib, name);                                                           \
    if (!a.field) { g_nccl_err = std::string("libnccl has no symbol ") + name; return nullptr; }
    GB2_SYM(GetUniqueId, "ncclGetUniqueId")
    GB2_SYM(CommInitRank, "ncclCommInitRank")
    GB2_SYM(CommDestroy, "ncclCommDestroy")
    GB2_SYM(Broadcast, "ncclBroadcast")
    GB2_SYM(AllGather, "ncclAllGather")
    GB2_SYM(GroupStart, "ncclGroupStart")
    GB2_SYM(GroupEnd, "ncclGroupEnd")
    GB2_SYM(GetErrorString, "ncclGetErrorString")
#undef GB2_SYM
    g_nccl = a;
    return &g_nccl;
}

int gb2_nccl_unique_id(char* out128) {
    if (!out128) { g_create_err = "invalid argument: out128 is null"; return -1; }
    const NcclApi* nc = load_nccl();
    if (!nc) { g_create_err = g_nccl_err; return -5; }
    NcclId id;
    const int rc = nc->GetUniqueId(&id);
    if (rc != 0) { g_create_err = std::string("ncclGetUniqueId: ") + nc->GetErrorString(rc); return -200 - rc; }
    memcpy(out128, id.bytes, sizeof(id.bytes));
    return 0;
}

int gb2_dist_init(gb2_handle* h, int rank, int world, const char* unique_id128) {
    if (!h) return -1;
    GB2_ARG(h, world >= 1 && rank >= 0 && rank < world, "gb2_dist_init: need 0 <= rank < world");
    GB2_ARG(h, !h->comm, "gb2_dist_init: already initialised");
    if (world == 1) { h->rank = 0; h->world = 1; return 0; }
    GB2_ARG(h, unique_id128 != nullptr, "gb2_dist_init: unique id is null");
    const NcclApi* nc = load_nccl();
    if (!nc) { h->err = g_nccl_err; return -5; }
    GB2_CUDA(h, cudaSetDevice(h->device));
    NcclId id;
    memcpy(id.bytes, unique_id128, sizeof(id.bytes));
    void* comm = nullptr;
    const int rc = nc->CommInitRank(&comm, world, id, rank);
    if (rc != 0) { h->err = std::string("ncclCommInitRank: ") + nc->GetErrorString(rc); return -200 - rc; }
    h->nccl = nc; h->comm = comm; h->rank = rank; h->world = world;
    h->factorized = false;
    return 0;
}

int gb2_dist_allgather_dev(gb2_handle* h, const double* dsend, double* drecv, int64_t count) {
    if (!h) return -1;
    GB2_ARG(h, dsend && drecv && count >= 0, "gb2_dist_allgather_dev: bad arguments");
    GB2_CUDA(h, cudaSetDevice(h->device));
    if (h->world == 1) {
        if (dsend != drecv) GB2_CUDA(h, cudaMemcpyAsync(drecv, dsend, (size_t)count * sizeof(double), cudaMemcpyDeviceToDevice, h->s_main));
    } else {
        const int rc = h->nccl->AllGather(dsend, drecv, (size_t)count, NCCL_FLOAT64, h->comm, h->s_main);
        if (rc != 0) { h->err = std::string("ncclAllGather: ") + h->nccl->GetErrorString(rc); return -200 - rc; }
    }
    GB2_CUDA(h, cudaStreamSynchronize(h->s_main));
    return 0;
}

int gb2_dist_finalize(gb2_handle* h) {
    if (!h) return -1;
    if (h->comm) {
        cudaSetDevice(h->device);
        cudaDeviceSynchronize();
        p2p_close(h);
        h->nccl->CommDestroy(h->comm);
        h->comm = nullptr;
    }
    h->rank = 0; h->world = 1;
    h->factorized = false;
    return 0;
}

int gb2_set_option(gb2_handle* h, const char* name, int value) {
    if (!h || !name) return -1;
    if (!strcmp(name, "lookahead")) { h->opt_lookahead = value ? 1 : 0; return 0; }
    if (!strcmp(name, "tf32_leaf")) {
        GB2_ARG(h, value >= 1 && value <= 16, "tf32_leaf must be in [1, 16]");
        h->opt_tf32_leaf = value;
        return 0;
    }
    if (!strcmp(name, "shard_storage")) {   // 1: every rank stores only the row blocks it owns (N beyond one GPU's HBM); collective choice
        h->opt_shard_storage = value ? 1 : 0;
        h->factorized = false;
        return 0;
    }
    if (!strcmp(name, "p2p")) {   // 1: panel exchange through NVLink peer mappings (default), 0: NCCL broadcast + all-gather
        GB2_ARG(h, !h->p2p_ready || value, "p2p cannot be switched off once the peer mappings are in use");
        h->opt_p2p = value ? 1 : 0;
        return 0;
    }
    if (!strcmp(name, "solve_streams")) {   // fp64 predict solve: row slabs of the prediction points in concurrent streams
        GB2_ARG(h, value >= 1 && value <= 4, "solve_streams must be in [1, 4]");
        h->opt_solve_streams = value;
        return 0;
    }
    if (!strcmp(name, "fp64_panel")) {   // two-level blocking of the fp64 factorisation: column blocks per panel (0/1 = plain algorithm)
        GB2_ARG(h, value >= -1 && value <= 16, "fp64_panel must be in [-1, 16] (-1 = auto)");
        h->opt_fp64_panel = value;
        h->factorized = false;
        return 0;
    }
    if (!strcmp(name, "trace")) {   // timeline stamps around the kernels of every block step (gb2_get_trace); measurement aid
        if (!value) { if (h->dTrace) cudaFree(h->dTrace); h->dTrace = nullptr; h->trace_cap = 0; return 0; }
        const int64_t want = (int64_t)TRACE_SLOTS * 4096;     // up to 4096 block steps (N <= 524k)
        if (!h->dTrace) { GB2_CUDA(h, cudaMalloc(&h->dTrace, want * sizeof(unsigned long long))); h->trace_cap = want; }
        GB2_CUDA(h, cudaMemset(h->dTrace, 0, want * sizeof(unsigned long long)));
        return 0;
    }
    if (!strcmp(name, "fused_group")) {   // gb2_factorize_predict: column blocks per bulk update of the prediction rows
        GB2_ARG(h, value == 1 || value == 2 || value == 4 || value == 8, "fused_group must be 1, 2, 4 or 8");
        h->opt_fused_group = value;
        return 0;
    }
    if (!strcmp(name, "chain_on_panel")) { h->opt_chain_on_panel = value ? 1 : 0; return 0; }   // ablation: Cholesky stream choreography
    if (!strcmp(name, "kbuild_occ")) {   // register bound of the strip K-build: 3 or 4 resident CTAs per SM
        GB2_ARG(h, value == 3 || value == 4, "kbuild_occ must be 3 or 4");
        h->opt_kbuild_occ = value;
        return 0;
    }
    if (!strcmp(name, "dgemm_tma")) { g_dgemm_tma = value == 1 ? 7 : (value & 7); h->factorized = false; return 0; }   // ablation (process-wide): 0 = cp.async-staged fp64 GEMM
    if (!strcmp(name, "defer_wait")) { h->opt_defer_wait = value ? 1 : 0; return 0; }   // ablation (multi-GPU chain)
    if (!strcmp(name, "bulk_persistent")) { GB2_ARG(h, value >= 0 && value <= 4096, "bulk_persistent must be in [0, 4096]"); h->opt_bulk_persistent = value; return 0; }   // ablation: persistent grids for the factorisation's bulk updates
    if (!strcmp(name, "dgemm_deep")) { GB2_ARG(h, value >= 0, "dgemm_deep must be >= 0"); g_dgemm_deep = value; return 0; }   // ablation (process-wide)
    if (!strcmp(name, "dgemm_fence")) { g_dgemm_fence = value ? 1 : 0; return 0; }   // diagnostic (process-wide)
    if (!strcmp(name, "dgemm_persistent")) { g_dgemm_persistent = value; return 0; }   // ablation (process-wide): 0 = one CTA per output tile
    if (!strcmp(name, "kbuild_persist")) { h->opt_kbuild_persist = value ? 1 : 0; h->factorized = false; return 0; }   // ablation: 0 = round-1 strip / per-tile kernels
    if (!strcmp(name, "kbuild_v1")) { h->opt_kbuild_v1 = value ? 1 : 0; h->factorized = false; return 0; }   // ablation: scalar-FMA + libm exp K-build
    if (!strcmp(name, "tf32_nb")) {   // panel width of the GB2_TF32 factorisation / leaf width of its solve, in 128-column blocks
        GB2_ARG(h, value >= 0 && value <= 16, "tf32_nb must be in [0, 16] (0 = auto)");
        h->opt_tf32_nb = value; h->P_cap = 0; h->factorized = false;
        if (h->dPhi) { cudaFree(h->dPhi); h->dPhi = nullptr; }
        if (h->dPlo) { cudaFree(h->dPlo); h->dPlo = nullptr; }
        return 0;
    }
    h->err = std::string("unknown option: ") + name;
    return -1;
}

}  // extern "C"
