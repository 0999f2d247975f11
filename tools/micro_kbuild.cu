// Stand-alone timing + accuracy check of the K-build kernels (strip kernel of round 1 vs the persistent kernel), one GPU.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --expt-relaxed-constexpr tools/micro_kbuild.cu -o tools/micro_kbuild
//   ./tools/micro_kbuild [N=32768]
// Prints, per (kind, d, P): ms and algorithmic GB/s (8 N (N+1) / 2 bytes) of each kernel, max relative deviation between the two
// over the whole lower triangle, and max relative error of sampled entries against a long-double host evaluation of the
// reference formula (Stationary.square_dist expanded form, clipped; Matern: sqrt(r2 + 1e-12)).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>
#include "../gumbi_b200/csrc/kbuild_persist.cuh"
#include "kbuild_persist_r02t.cuh"   // the v4 kernel of round-2 call t, frozen: same-box comparison arm

using namespace gb2;
static bool g_one = false;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

__global__ void maxdiff_kernel(const double* A, const double* B, int64_t n, int64_t ld, double* out) {
    double m = 0;
    for (int64_t r = blockIdx.x; r < n; r += gridDim.x)
        for (int64_t c = threadIdx.x; c <= r; c += blockDim.x) {
            const double a = A[r * ld + c], b = B[r * ld + c];
            const double dlt = fabs(a - b) / fmax(fabs(b), 1e-300);
            if (dlt > m) m = dlt;
        }
    for (int o = 16; o; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<unsigned long long*>(out), (unsigned long long)__double_as_longlong(m));
}

// write-only ceilings: (a) plain streaming fill of the whole square, (b) the K-build's own store pattern (lower-triangle 64x64 tiles,
// 8 warps x (8 rows x 64 B) per store instruction, persistent CTAs over strips) with the arithmetic removed
__global__ void fill_kernel(double2* p, size_t n2) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (size_t)gridDim.x * blockDim.x) p[i] = make_double2(1.0, 2.0);
}
__global__ void __launch_bounds__(256) tile_fill_kernel(double* out, int64_t ld, int n_tiles, int strip, int* ctr) {
    __shared__ int s_item;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t4 = lane & 3;
    const int r0 = (warp >> 1) * 16, c0 = (warp & 1) * 32;
    const int ng = (n_tiles + strip - 1) / strip, full = ng - 1;
    const int n_items = strip * (full * (full + 1) / 2) + (n_tiles - full * strip) * ng;
    for (;;) {
        __syncthreads();
        if (tid == 0) s_item = atomicAdd(ctr, 1);
        __syncthreads();
        const int w = s_item;
        if (w >= n_items) break;
        int q = (int)((sqrt(8.0 * (double)w / strip + 1.0) - 1.0) * 0.5);
        while (strip * ((q + 1) * (q + 2) / 2) <= w) q++;
        while (q > 0 && strip * (q * (q + 1) / 2) > w) q--;
        const int rem = w - strip * (q * (q + 1) / 2);
        const int bi = q * strip + rem / (q + 1), sidx = rem % (q + 1);
        const int jt0 = sidx * strip;
        int jt1 = jt0 + strip < n_tiles ? jt0 + strip : n_tiles;
        if (jt1 > bi + 1) jt1 = bi + 1;
        for (int jt = jt0; jt < jt1; jt++)
            for (int mi = 0; mi < 2; mi++) {
                double* dst = out + ((int64_t)bi * 64 + r0 + mi * 8 + g) * ld + (int64_t)jt * 64 + c0 + 2 * t4;
                for (int ni = 0; ni < 4; ni++) *reinterpret_cast<double2*>(dst + ni * 8) = make_double2((double)w, (double)jt);
            }
    }
    if (tid == 0) { __threadfence(); if (atomicAdd(ctr + 1, 1) == (int)gridDim.x - 1) { ctr[0] = 0; ctr[1] = 0; } }
}

static void write_ceilings(int64_t Np, int n_sm) {
    double* d; int* ctr;
    CK(cudaMalloc(&d, (size_t)Np * Np * 8)); CK(cudaMalloc(&ctr, 16)); CK(cudaMemset(ctr, 0, 16));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float ms;
    const double sq = (double)Np * Np * 8, tri = 8.0 * Np * (Np + 64) / 2;
    for (int it = 0; it < 2; it++) { CK(cudaEventRecord(e0)); CK(cudaMemsetAsync(d, 0, (size_t)Np * Np * 8)); CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize()); }
    CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("write-only ceilings, %lld x %lld fp64: cudaMemset %.3f ms %.0f GB/s", (long long)Np, (long long)Np, ms, sq / ms / 1e6);
    for (int it = 0; it < 2; it++) { CK(cudaEventRecord(e0)); fill_kernel<<<n_sm * 8, 256>>>((double2*)d, (size_t)Np * Np / 2); CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize()); }
    CK(cudaEventElapsedTime(&ms, e0, e1));
    printf(" | streaming STG.128 fill %.3f ms %.0f GB/s", ms, sq / ms / 1e6);
    for (int occ : {2, 4, 8}) {
        for (int it = 0; it < 2; it++) { CK(cudaEventRecord(e0)); tile_fill_kernel<<<n_sm * occ, 256>>>(d, Np, (int)(Np / 64), 8, ctr); CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize()); }
        CK(cudaEventElapsedTime(&ms, e0, e1));
        printf(" | K-build store pattern (lower tiles, %d CTAs/SM) %.3f ms %.0f GB/s", occ, ms, tri / ms / 1e6);
    }
    printf("\n");
    cudaFree(d); cudaFree(ctr);
}

static long double ref_entry(int kind, const std::vector<double>& X, int D_in, int d, const double* ls, int64_t i, int64_t j) {
    long double si = 0, sj = 0, dot = 0;
    for (int k = 0; k < d; k++) {
        const double ui = X[i * D_in + k] * (1.0 / ls[k]), uj = X[j * D_in + k] * (1.0 / ls[k]);   // Stationary: X * (1/ls) in fp64
        si += (long double)ui * ui; sj += (long double)uj * uj; dot += (long double)ui * uj;
    }
    long double r2 = si + sj - 2 * dot;
    if (r2 < 0) r2 = 0;
    if (kind == GB2_EXPQUAD) return expl(-0.5L * r2);
    const long double r = sqrtl(r2 + 1e-12L);
    if (kind == GB2_MATERN52) return (1 + sqrtl(5.0L) * r + 5.0L / 3 * r * r) * expl(-sqrtl(5.0L) * r);
    if (kind == GB2_MATERN32) return (1 + sqrtl(3.0L) * r) * expl(-sqrtl(3.0L) * r);
    if (kind == GB2_MATERN12) return expl(-r);
    return expl(-0.5L * r);
}

template <int KIND, int KS, int NCG, int OCC>
static float time_persist(const KParams& kp, KB4Args a, int n_sm, int strip, int reps) {
    CK(cudaFuncSetAttribute(kbuild_persist_kernel<true, KIND, KS, NCG, OCC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kb4_smem_bytes<KS, NCG>()));
    a.strip = strip;
    kb4_set_constants(a, kp.t[0].kind);
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int i = 0; i < 2; i++) kbuild_persist_kernel<true, KIND, KS, NCG, OCC><<<n_sm * OCC, KB_THREADS, kb4_smem_bytes<KS, NCG>()>>>(kp, a);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    for (int i = 0; i < reps; i++) kbuild_persist_kernel<true, KIND, KS, NCG, OCC><<<n_sm * OCC, KB_THREADS, kb4_smem_bytes<KS, NCG>()>>>(kp, a);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    return ms / reps;
}

template <int KIND, int KS, int NCG>
static float time_persist_r02t(const KParams& kp, const KB4Args& a, int n_sm, int strip, int reps) {
    KB4OArgs o{};
    o.Fi = a.Fi; o.stride_i = a.stride_i; o.n_i = a.n_i; o.Fj = a.Fj; o.stride_j = a.stride_j; o.n_j = a.n_j; o.Ci = a.Ci; o.Cj = a.Cj; o.Btab = a.Btab;
    o.y = a.y; o.out = a.out; o.ld = a.ld; o.n_row_tiles = a.n_row_tiles; o.n_col_tiles = a.n_col_tiles; o.strip = strip;
    o.own_stride = 1; o.own_rank = 0; o.compact = 0; o.ctr = a.ctr;
    CK(cudaFuncSetAttribute(kbuild_persist_r02t_kernel<true, KIND, KS, NCG, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kb4o_smem_bytes<KS, NCG>()));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int i = 0; i < 2; i++) kbuild_persist_r02t_kernel<true, KIND, KS, NCG, 4><<<n_sm * 4, KB_THREADS, kb4o_smem_bytes<KS, NCG>()>>>(kp, o);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    for (int i = 0; i < reps; i++) kbuild_persist_r02t_kernel<true, KIND, KS, NCG, 4><<<n_sm * 4, KB_THREADS, kb4o_smem_bytes<KS, NCG>()>>>(kp, o);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    return ms / reps;
}

template <int KIND, int KS, int NCG>
static void run_case(int64_t n, int d, int P, int n_sm) {
    const int64_t N = n * P, Np = round_up(N + 1, TILE);
    const int D_in = d + (P > 1 ? 1 : 0);
    std::mt19937_64 rng(2021);
    std::normal_distribution<double> nd;
    std::vector<double> X((size_t)N * D_in), y(N), ls(d);
    for (int k = 0; k < d; k++) ls[k] = (1.0 + 0.25 * k) * std::sqrt((double)d);
    for (int64_t i = 0; i < n; i++)
        for (int k = 0; k < d; k++) {
            const double v = nd(rng);
            for (int p = 0; p < P; p++) X[(size_t)(p * n + i) * D_in + k] = v;
        }
    if (P > 1) for (int p = 0; p < P; p++) for (int64_t i = 0; i < n; i++) X[(size_t)(p * n + i) * D_in + d] = p;
    for (auto& v : y) v = nd(rng);
    KParams kp{}; PrepParams pp{};
    kp.n_terms = pp.n_terms = 1; pp.D_in = D_in;
    kp.t[0].kind = KIND; kp.t[0].d = pp.d[0] = d; kp.t[0].n_lin = pp.n_lin[0] = 0; kp.t[0].feat_off = pp.feat_off[0] = 0;
    kp.t[0].eta2 = 1.3; kp.t[0].tau = 0; kp.n_feat = d + 1; kp.sigma2 = 0.01; kp.jitter = 1e-6; kp.noise_cat = -1;
    for (int k = 0; k < d; k++) { pp.cont_idx[0][k] = k; pp.inv_ls[0][k] = 1.0 / ls[k]; }
    std::vector<double> B;
    if (P > 1) {
        kp.n_cat = pp.n_cat = 1; pp.cat_col[0] = d; pp.cat_P[0] = P;
        kp.t[0].n_coreg = 1; kp.t[0].cg_cat[0] = 0; kp.t[0].cg_P[0] = P; kp.t[0].cg_Boff[0] = 0;
        B.resize(P * P);
        std::vector<double> W(P * 2);
        for (auto& v : W) v = nd(rng);
        for (int p = 0; p < P; p++) for (int q = 0; q < P; q++) B[p * P + q] = W[2 * p] * W[2 * q] + W[2 * p + 1] * W[2 * q + 1] + (p == q ? 1.0 : 0.0);
    }
    double *dX, *dy, *dF, *dA, *dB2, *dBtab = nullptr, *dMax;
    int *dC, *dBad, *dCtr;
    CK(cudaMalloc(&dX, X.size() * 8)); CK(cudaMalloc(&dy, y.size() * 8)); CK(cudaMalloc(&dF, (size_t)(d + 1) * Np * 8));
    CK(cudaMalloc(&dC, (size_t)Np * 4)); CK(cudaMalloc(&dBad, 4)); CK(cudaMalloc(&dCtr, 16)); CK(cudaMalloc(&dMax, 8));
    CK(cudaMalloc(&dA, (size_t)Np * Np * 8)); CK(cudaMalloc(&dB2, (size_t)Np * Np * 8));
    CK(cudaMemset(dCtr, 0, 16)); CK(cudaMemset(dBad, 0, 4)); CK(cudaMemset(dMax, 0, 8));
    CK(cudaMemcpy(dX, X.data(), X.size() * 8, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dy, y.data(), y.size() * 8, cudaMemcpyHostToDevice));
    if (P > 1) { CK(cudaMalloc(&dBtab, B.size() * 8)); CK(cudaMemcpy(dBtab, B.data(), B.size() * 8, cudaMemcpyHostToDevice)); }
    prep_features<<<(unsigned)((Np + 255) / 256), 256>>>(dX, N, Np, pp, dF, dC, dBad);
    CK(cudaDeviceSynchronize());

    // round-1 kernels (strip kernel for the plain model, generic DMMA kernel with Coregion) into dB2
    dim3 grid((unsigned)(Np / KB_T), (unsigned)(Np / KB_T));
    CK(kbuild_dmma_configure<true>());
    const int reps = N >= 16384 ? 5 : 20;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int i = 0; i < 2; i++)
        kbuild_dmma_launch<true>(0, grid, kbuild_dmma_smem_bytes(kp), kp, dBtab, dF, dC, Np, N, dF, dC, Np, N, dy, dB2, Np, 1, 0, 0, 4);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    for (int i = 0; i < reps; i++)
        kbuild_dmma_launch<true>(0, grid, kbuild_dmma_smem_bytes(kp), kp, dBtab, dF, dC, Np, N, dF, dC, Np, N, dy, dB2, Np, 1, 0, 0, 4);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms_old;
    CK(cudaEventElapsedTime(&ms_old, e0, e1));
    ms_old /= reps;
    const double bytes = 8.0 * N * (N + 1) / 2 + 8.0 * N * D_in;
    printf("kind %d d %d P %d N %lld | round-1 kernel %.3f ms %.0f GB/s\n", KIND, d, P, (long long)N, ms_old, bytes / ms_old / 1e6);

    KB4Args a{};
    a.Fi = dF; a.stride_i = Np; a.n_i = N; a.Fj = dF; a.stride_j = Np; a.n_j = N; a.Ci = dC; a.Cj = dC; a.Btab = dBtab; a.y = dy; a.out = dA; a.ld = Np;
    a.n_row_tiles = a.n_col_tiles = (int)(Np / KB_T); a.own_stride = 1; a.own_rank = 0; a.compact = 0; a.ctr = dCtr;
    if (!g_one) {
        const float mo = time_persist_r02t<KIND, KS, NCG>(kp, a, n_sm, 8, reps);
        printf("   v4 (round-2 call t) persistent kernel, strip 8 occ4: %.3f ms %.0f GB/s\n", mo, bytes / mo / 1e6);
    }
    if (g_one) {   // profiling mode: the product configuration only
        const float m4 = time_persist<KIND, KS, NCG, KB4_OCC>(kp, a, n_sm, 8, 3);
        printf("   persistent strip 8 occ%d %.3f ms %.0f GB/s\n", KB4_OCC, m4, bytes / m4 / 1e6);
    } else
    for (int strip : {4, 8, 16}) {
        const float m2 = time_persist<KIND, KS, NCG, 2>(kp, a, n_sm, strip, reps);
        const float m3 = time_persist<KIND, KS, NCG, 3>(kp, a, n_sm, strip, reps);
        const float m4 = time_persist<KIND, KS, NCG, 4>(kp, a, n_sm, strip, reps);
        printf("   v6 persistent strip %2d : occ2 %.3f ms %.0f GB/s | occ3 %.3f ms %.0f GB/s | occ4 %.3f ms %.0f GB/s\n", strip, m2, bytes / m2 / 1e6, m3,
               bytes / m3 / 1e6, m4, bytes / m4 / 1e6);
    }
    // accuracy: new vs round-1 over the whole lower triangle (incl. the y row), sampled entries vs long double
    maxdiff_kernel<<<1024, 256>>>(dA, dB2, N + 1, Np, dMax);
    double md;
    CK(cudaMemcpy(&md, dMax, 8, cudaMemcpyDeviceToHost));
    double worst_new = 0, worst_old = 0;
    std::uniform_int_distribution<int64_t> ui(0, N - 1);
    for (int s = 0; s < 4000; s++) {
        int64_t i = ui(rng), j = ui(rng);
        if (s < 400) j = i;                 // diagonal entries
        else if (s < 800) j = std::max<int64_t>(0, i - (s & 63));   // near-diagonal
        if (j > i) std::swap(i, j);
        double vn, vo;
        CK(cudaMemcpy(&vn, dA + i * Np + j, 8, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(&vo, dB2 + i * Np + j, 8, cudaMemcpyDeviceToHost));
        long double r = 1.3L * ref_entry(KIND, X, D_in, d, ls.data(), i, j);
        if (P > 1) r *= B[(int)X[i * D_in + d] * P + (int)X[j * D_in + d]];
        if (i == j) r += 0.01L + 1e-6L;
        const double den = std::max((double)fabsl(r), 1e-300);
        worst_new = std::max(worst_new, (double)fabsl(vn - r) / den);
        worst_old = std::max(worst_old, (double)fabsl(vo - r) / den);
    }
    printf("   max rel dev new vs round-1 (lower triangle) %.3e | sampled vs long double: new %.3e  round-1 %.3e\n", md, worst_new, worst_old);
    cudaFree(dX); cudaFree(dy); cudaFree(dF); cudaFree(dC); cudaFree(dBad); cudaFree(dCtr); cudaFree(dMax); cudaFree(dA); cudaFree(dB2);
    if (dBtab) cudaFree(dBtab);
}

int main(int argc, char** argv) {
    const int64_t N = argc > 1 ? atoll(argv[1]) : 32768;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    {
        double tab[64];
        for (int j = 0; j < 64; j++) tab[j] = std::exp2((double)j / 64.0);
        CK(cudaMemcpyToSymbol(g_exp2_tab, tab, sizeof(tab)));
        std::vector<double> t2(KB4_TAB);
        for (int j = 0; j < KB4_TAB; j++) t2[j] = std::exp2((double)j / KB4_TAB);
        CK(cudaMemcpyToSymbol(g_exp2_tab2k, t2.data(), KB4_TAB * sizeof(double)));
        CK(cudaMemcpyToSymbol(g_exp2_tab2k_r02t, t2.data(), KB4_TAB * sizeof(double)));
    }
    g_one = argc > 2;
    const int n_sm = prop.multiProcessorCount;
    printf("%s, %d SMs\n", prop.name, n_sm);
    if (!g_one) write_ceilings(round_up(N + 1, TILE), n_sm);
    run_case<GB2_EXPQUAD, 2, 0>(N, 8, 1, n_sm);
    run_case<GB2_MATERN52, 2, 0>(N, 8, 1, n_sm);
    if (g_one) return 0;
    run_case<GB2_EXPQUAD, 1, 1>(N / 2, 4, 2, n_sm);
    run_case<GB2_MATERN32, 2, 0>(N / 4, 5, 1, n_sm);
    run_case<GB2_EXPQUAD, 2, 0>(8192, 8, 1, n_sm);
    return 0;
}
