// Stand-alone rates of the fp64 DMMA GEMM (gumbi_b200/csrc/dgemm.cuh) on the shapes the predict solve and the Cholesky trailing
// update launch, against the wave count of each launch (dev tool; not part of the product path).  Answers: how far is the kernel
// itself from the cuBLAS DGEMM figure on a full-wave problem, and how much do partial last waves cost on the recursion's shapes.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I gumbi_b200/csrc -o tools/micro_dgemm tools/micro_dgemm.cu
#include "dgemm.cuh"
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
using namespace gb2;

template <int BM, int BN, int MODE, int BK = GM_BK, int STAGES = GM_STAGES, int WM = 32, int WN = 32, int MINB = 2>
static double run(const double* A, const double* B, double* C, int64_t ld, int64_t rows, int64_t cols, int k, int lower, int reps) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    if (dgemm_nt_configure<BM, BN, MODE, BK, STAGES, WM, WN, MINB>() != cudaSuccess) { printf("configure failed\n"); exit(1); }
    for (int i = 0; i < 2; i++) dgemm_nt_launch<BM, BN, MODE, BK, STAGES, WM, WN, MINB>(0, A, ld, B, ld, C, ld, rows, cols, k, lower, 0, 0);
    cudaEventRecord(e0);
    for (int i = 0; i < reps; i++) dgemm_nt_launch<BM, BN, MODE, BK, STAGES, WM, WN, MINB>(0, A, ld, B, ld, C, ld, rows, cols, k, lower, 0, 0);
    cudaEventRecord(e1);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); exit(1); }
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms / reps;
}

// "check" mode: exactness of the persistent TMA path in the regime of the factorisation's trailing update (k = 128: 8 k-tiles per
// output tile, lower-triangular tile list, many tiles per CTA), repeated; data are multiples of 2^-16 so every sum is exact and the
// result must equal the cp.async kernel's bit for bit.  Small enough to run under compute-sanitizer --tool racecheck.
static int check_mode(int reps) {
    const int64_t ld = 4096, R = 4096 + 512;
    double *A, *B, *C;
    cudaMalloc(&A, R * ld * 8); cudaMalloc(&B, R * ld * 8); cudaMalloc(&C, R * ld * 8);
    std::vector<double> h((size_t)R * ld);
    for (size_t i = 0; i < h.size(); i++) h[i] = (double)((i * 2654435761u) >> 8 & 0xffff) / 65536.0 - 0.5;
    cudaMemcpy(A, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(B, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
    struct PS { int64_t r, c; int k; int lower; } ps[] = {{3072, 3072, 128, 1}, {4096, 2048, 128, 0}, {4096, 4096, 128, 1}, {4096, 3072, 256, 0}, {2048, 4096, 1024, 0}};
    std::vector<double> ref((size_t)R * ld), got((size_t)R * ld);
    int rc = 0;
    for (int mode : {1, 2}) {
        for (auto q : ps) {
            g_dgemm_tma = 0;
            cudaMemset(C, 0, (size_t)R * ld * 8);
            dgemm_nt_configure<128, 64, GM_SUB>();
            dgemm_nt_launch<128, 64, GM_SUB>(0, A + 128 * ld + 256, ld, B + 256 * ld + 512, ld, C, ld, q.r, q.c, q.k, q.lower, 0, 0);
            cudaDeviceSynchronize();
            cudaMemcpy(ref.data(), C, (size_t)q.r * ld * 8, cudaMemcpyDeviceToHost);
            size_t bad_total = 0; int bad_runs = 0;
            for (int rep = 0; rep < reps; rep++) {
                g_dgemm_tma = 7; g_dgemm_persistent = mode;
                cudaMemset(C, 0, (size_t)R * ld * 8);
                dgemm_nt_launch<128, 64, GM_SUB>(0, A + 128 * ld + 256, ld, B + 256 * ld + 512, ld, C, ld, q.r, q.c, q.k, q.lower, 0, 0);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(e)); return 2; }
                cudaMemcpy(got.data(), C, (size_t)q.r * ld * 8, cudaMemcpyDeviceToHost);
                size_t bad = 0;
                for (int64_t i = 0; i < q.r; i++)
                    for (int64_t j = 0; j < q.c; j++) bad += got[i * ld + j] != ref[i * ld + j];
                bad_total += bad; bad_runs += bad != 0;
            }
            printf("check persistent=%d TMA SUB %lld x %lld x %d lower=%d: %zu mismatching entries, %d of %d runs wrong\n", mode, (long long)q.r,
                   (long long)q.c, q.k, q.lower, bad_total, bad_runs, reps);
            rc |= bad_total != 0;
        }
        // in-place leaf X <- X Dinv^T, 64x128 tiles, k = 128, 24576 rows = 384 tiles (> 296 slots), X inside a narrow matrix (ld 512)
        {
            const int64_t r = 24576, ld2 = 512;
            double *X, *X0;
            cudaMalloc(&X, r * ld2 * 8); cudaMalloc(&X0, r * ld2 * 8);
            cudaMemcpy(X0, A, r * ld2 * 8, cudaMemcpyDeviceToDevice);
            std::vector<double> r2((size_t)r * ld2), g2((size_t)r * ld2);
            size_t bad_total = 0; int bad_runs = 0;
            for (int rep = 0; rep <= reps; rep++) {
                g_dgemm_tma = rep == 0 ? 0 : 7; g_dgemm_persistent = mode;
                cudaMemcpy(X, X0, r * ld2 * 8, cudaMemcpyDeviceToDevice);
                dgemm_nt_configure<64, 128, GM_SET>();
                dgemm_nt_launch<64, 128, GM_SET>(0, X + 128, ld2, B + 5 * 128 * ld, ld, X + 128, ld2, r, 128, 128, 0, 0, 0);
                cudaDeviceSynchronize();
                cudaMemcpy((rep == 0 ? r2 : g2).data(), X, r * ld2 * 8, cudaMemcpyDeviceToHost);
                if (rep == 0) continue;
                size_t bad = 0;
                for (size_t i = 0; i < r2.size(); i++) bad += g2[i] != r2[i];
                bad_total += bad; bad_runs += bad != 0;
            }
            printf("check persistent=%d TMA in-place SET %lld x 128 x 128: %zu mismatching entries, %d of %d runs wrong\n", mode, (long long)r, bad_total,
                   bad_runs, reps);
            rc |= bad_total != 0;
            cudaFree(X); cudaFree(X0);
        }
    }
    return rc;
}

// "pipeline" mode: the bulk trailing update exactly as the factorisation issues it (same pointer arithmetic into ONE matrix, rb_first,
// col_off, lower tile list), in three settings:  static = operands untouched between launches;  fresh = the panel (column block k) is
// re-written by a copy kernel right before each update (same stream);  concurrent = a second stream keeps re-writing column block
// k+1 (which the update neither reads nor writes) while the update runs.  TMA-staged result vs cp.async result, exact data.
__global__ void copy_cols_kernel(const double* __restrict__ src, double* __restrict__ dst, int64_t ld, int64_t rows, int64_t c0, int ncols) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < rows * ncols; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / ncols, c = c0 + i % ncols;
        dst[r * ld + c] = src[r * ld + c];
    }
}
__global__ void empty_kernel() {}
static int pipeline_mode(int reps) {
    const int64_t nb = 40, Np = nb * 128, ld = Np;            // 5120 x 5120, row stride not a power of two
    double *M0, *M, *Mref;
    cudaMalloc(&M0, Np * ld * 8); cudaMalloc(&M, Np * ld * 8); cudaMalloc(&Mref, Np * ld * 8);
    std::vector<double> h((size_t)Np * ld), ref((size_t)Np * ld), got((size_t)Np * ld);
    for (size_t i = 0; i < h.size(); i++) h[i] = (double)((i * 2654435761u) >> 8 & 0xffff) / 65536.0 - 0.5;
    cudaMemcpy(M0, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
    cudaStream_t s1, s2;
    cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking); cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking);
    dgemm_nt_configure<128, 64, GM_SUB>();
    int rc = 0;
    auto update = [&](cudaStream_t s, double* A, int k) {
        const int64_t g0 = (int64_t)k * 128;
        const int c2 = (int)(nb - (k + 2));
        dgemm_nt_launch<128, 64, GM_SUB>(s, A + g0, ld, A + (g0 + 256) * ld + g0, ld, A + g0 + 256, ld, (int64_t)c2 * 128, (int64_t)c2 * 128, 128, 1, 0,
                                         g0 + 256, k + 2, 1);
    };
    const char* names[] = {"static operands", "panel re-written by a copy kernel right before the update", "neighbour column re-written concurrently",
                           "EMPTY kernel right before the update", "copy kernel writes ANOTHER matrix right before", "copy kernel, stream sync, update",
                           "panel re-written right before, one CTA per tile", "panel re-written right before, persistent without cross-tile prefetch",
                           "panel re-written right before (repeat)",
                           "copy kernel writes ANOTHER matrix right before, WITHOUT the stage fence (control)", "panel re-written right before, WITHOUT the stage fence (control)"};
    for (int setting = (reps > 12 ? 4 : 0); setting < 11; setting++) {
        if (reps > 12 && (setting == 6 || setting == 7 || setting == 8)) continue;
        size_t bad_total = 0; int bad_runs = 0, runs = 0; bool shown = false;
        g_dgemm_persistent = setting == 6 ? 0 : (setting == 7 ? 2 : 1);
        g_dgemm_fence = setting >= 9 ? 0 : 1;     // settings 9, 10: WITHOUT the stage-release proxy fence (control: the round-2 bug)
        for (int k : {0, 1, 3, 7, 12}) {
            g_dgemm_tma = 0;
            cudaMemcpy(Mref, M0, Np * ld * 8, cudaMemcpyDeviceToDevice);
            cudaDeviceSynchronize();
            update(s1, Mref, k);
            cudaStreamSynchronize(s1);
            cudaMemcpy(ref.data(), Mref, Np * ld * 8, cudaMemcpyDeviceToHost);
            for (int rep = 0; rep < reps; rep++) {
                g_dgemm_tma = 7;
                cudaMemcpy(M, M0, Np * ld * 8, cudaMemcpyDeviceToDevice);
                cudaDeviceSynchronize();
                if (setting == 1 || (setting >= 5 && setting != 9)) copy_cols_kernel<<<592, 256, 0, s1>>>(M0, M, ld, Np, (int64_t)k * 128, 128);
                if (setting == 2)
                    for (int q = 0; q < 40; q++) copy_cols_kernel<<<16, 256, 0, s2>>>(M0, M, ld, Np, (int64_t)(k + 1) * 128, 128);
                if (setting == 3) empty_kernel<<<592, 256, 0, s1>>>();
                if (setting == 4 || setting == 9) copy_cols_kernel<<<592, 256, 0, s1>>>(M0, Mref, ld, Np, (int64_t)k * 128, 128);
                if (setting == 5) cudaStreamSynchronize(s1);
                update(s1, M, k);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(e)); return 2; }
                cudaMemcpy(got.data(), M, Np * ld * 8, cudaMemcpyDeviceToHost);
                size_t bad = 0;
                for (size_t i = 0; i < got.size(); i++) bad += got[i] != ref[i];
                if (bad && !shown) {
                    shown = true;
                    int64_t rmin = Np, rmax = -1, cmin = Np, cmax = -1; int pr = 0;
                    for (size_t i = 0; i < got.size(); i++)
                        if (got[i] != ref[i]) {
                            const int64_t r = i / ld, c = i % ld;
                            rmin = r < rmin ? r : rmin; rmax = r > rmax ? r : rmax; cmin = c < cmin ? c : cmin; cmax = c > cmax ? c : cmax;
                            if (pr++ < 4) printf("      k=%d (%lld,%lld): got %.10f ref %.10f start %.10f\n", k, (long long)r, (long long)c, got[i], ref[i], h[i]);
                        }
                    printf("      first wrong run: %zu entries, rows %lld..%lld, cols %lld..%lld (tile origin col %lld)\n", bad, (long long)rmin, (long long)rmax,
                           (long long)cmin, (long long)cmax, (long long)((k + 2) * 128));
                }
                bad_total += bad; bad_runs += bad != 0; runs++;
            }
        }
        printf("pipeline check (%s): %zu mismatching entries, %d of %d runs wrong\n", names[setting], bad_total, bad_runs, runs);
        rc |= bad_total != 0;
    }
    g_dgemm_persistent = 1; g_dgemm_fence = 1;
    // two consecutive updates k, k+1 on one stream without host synchronisation in between (the second reads what the first wrote)
    {
        size_t bad_total = 0; int bad_runs = 0;
        for (int rep = 0; rep <= reps; rep++) {
            g_dgemm_tma = rep == 0 ? 0 : 7;
            double* T = rep == 0 ? Mref : M;
            cudaMemcpy(T, M0, Np * ld * 8, cudaMemcpyDeviceToDevice);
            cudaDeviceSynchronize();
            for (int k = 0; k < 6; k++) update(s1, T, k);
            cudaDeviceSynchronize();
            cudaMemcpy((rep == 0 ? ref : got).data(), T, Np * ld * 8, cudaMemcpyDeviceToHost);
            if (rep == 0) continue;
            size_t bad = 0;
            for (size_t i = 0; i < got.size(); i++) bad += fabs(got[i] - ref[i]) > 1e-9 * (1.0 + fabs(ref[i]));
            bad_total += bad; bad_runs += bad != 0;
        }
        printf("pipeline check (6 back-to-back updates, tolerance 1e-9): %zu mismatching entries, %d of %d runs wrong\n", bad_total, bad_runs, reps);
        rc |= bad_total != 0;
    }
    return rc;
}

int main(int argc, char** argv) {
    if (argc > 1 && !strcmp(argv[1], "check")) return check_mode(argc > 2 ? atoi(argv[2]) : 20);
    if (argc > 1 && !strcmp(argv[1], "pipeline")) return pipeline_mode(argc > 2 ? atoi(argv[2]) : 6);
    const int64_t ld = 16384, R = 16384;
    double *A, *B, *C;
    cudaMalloc(&A, R * ld * 8); cudaMalloc(&B, R * ld * 8); cudaMalloc(&C, R * ld * 8);
    std::vector<double> h((size_t)R * ld);
    for (size_t i = 0; i < h.size(); i++) h[i] = (double)((i * 2654435761u) >> 8 & 0xffff) / 65536.0 - 0.5;
    cudaMemcpy(A, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(B, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
    cudaMemset(C, 0, h.size() * 8);
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int slots = 2 * sms;
    printf("%-34s %9s %9s %8s %8s\n", "shape (rows x cols x k)", "ms", "TFLOP/s", "CTAs", "waves");
    struct S { int64_t r, c; int k; } shapes[] = {
        {8192, 8192, 8192}, {16384, 16384, 4096},                      // full-wave references
        {9472, 4096, 4096}, {10112, 4096, 4096},                      // 74 vs 79 row tiles (74 * 4 = 296 = one exact wave per 256 columns)
        {10112, 2048, 2048}, {10112, 1024, 1024}, {10112, 512, 512}, {10112, 256, 256}, {10112, 128, 128},
        {9472, 256, 256}, {5120, 256, 256}, {2560, 256, 256},
        // right-looking update of prediction rows appended below the factor (k = one 128-column panel per launch): the rate that
        // decides whether fusing the predict solve into the factorisation's trailing updates would pay
        {10112, 8192, 128}, {10112, 4096, 128}, {10112, 1024, 128}, {10112, 8192, 256}};
    for (int tma : {1, 0}) {
    g_dgemm_tma = tma ? 7 : 0;
    printf("---- operand staging: %s\n", tma ? "TMA (cp.async.bulk.tensor.2d + mbarrier ring)" : "cp.async + __syncthreads (round 1)");
    for (auto s : shapes) {
        const double ms = run<128, 64, GM_SUB>(A, B, C, ld, s.r, s.c, s.k, 0, 10);
        const double ctas = (double)(s.r / 128) * (s.c / 64);
        char name[64];
        snprintf(name, sizeof name, "SUB 128x64  %lld x %lld x %d", (long long)s.r, (long long)s.c, s.k);
        printf("%-34s %9.3f %9.2f %8.0f %8.2f\n", name, ms, 2.0 * s.r * s.c * s.k / ms / 1e9, ctas, ctas / slots);
    }
    for (int64_t r : {10112LL, 9472LL, 5120LL}) {   // the leaf: X <- X * Dinv^T, 64x128 tiles, k = 128
        const double ms = run<64, 128, GM_SET>(A, B, C, ld, r, 128, 128, 0, 20);
        char name[64];
        snprintf(name, sizeof name, "SET 64x128  %lld x 128 x 128", (long long)r);
        printf("%-34s %9.3f %9.2f %8.0f %8.2f\n", name, ms, 2.0 * r * 128 * 128 / ms / 1e9, (double)(r / 64), (double)(r / 64) / slots);
    }
    // lower-triangular trailing update of the Cholesky (SYRK shape): m x m blocks of 128, k = 128
    for (int m : {64, 32, 16, 8}) {
        const double ms = run<128, 64, GM_SUB>(A, A, C, ld, (int64_t)m * 128, (int64_t)m * 128, 128, 1, 20);
        const double ctas = (double)m * (m + 1);   // 128x64 tiles on or below the diagonal (2 per 128x128 block, diagonal blocks whole)
        char name[64];
        snprintf(name, sizeof name, "SYRK lower  %d blocks, k=128", m);
        printf("%-34s %9.3f %9.2f %8.0f %8.2f\n", name, ms, (double)m * (m + 1) / 2 * 2.0 * 128 * 128 * 128 / ms / 1e9, ctas, ctas / slots);
    }
    }   // tma
    {   // TMA-staged kernel against the cp.async kernel on the same product: same arithmetic, permuted summation order inside a k-tile
        const int64_t r = 2048, c = 1024; const int k = 1024;
        std::vector<double> ref((size_t)r * ld), got((size_t)r * ld);
        double worst = 0, scale = 0;
        for (int tma : {0, 1}) {
            g_dgemm_tma = tma ? 7 : 0;
            cudaMemset(C, 0, (size_t)r * ld * 8);
            run<128, 64, GM_SUB>(A + 3 * 128 * ld + 256, B + 128 * ld + 512, C, ld, r, c, k, 0, 0);   // operands at an offset inside their allocations
            run<64, 128, GM_SET>(A + 256, B + 5 * 128 * ld, C + 1024, ld, r, 128, 128, 0, 0);
            cudaMemcpy((tma ? got : ref).data(), C, (size_t)r * ld * 8, cudaMemcpyDeviceToHost);
        }
        for (int64_t i = 0; i < r; i++)
            for (int64_t j = 0; j < c + 128; j++) { worst = fmax(worst, fabs(got[i * ld + j] - ref[i * ld + j])); scale = fmax(scale, fabs(ref[i * ld + j])); }
        printf("check TMA-staged vs cp.async-staged kernel: max |diff| %.3e at scale %.3e (k = %d: rounding-level, different summation order)\n", worst, scale, k);
        g_dgemm_tma = 0;
        cudaMemset(C, 0, (size_t)R * ld * 8);
    }
    {   // the PERSISTENT path of the TMA kernel (more tiles than CTA slots: cross-tile prefetch), repeated to expose races; the test
        // data are multiples of 2^-16, so every product and partial sum is exact and the two kernels must agree bit for bit
        struct PS { int64_t r, c; int k; int lower; } ps[] = {{8192, 4096, 512, 0}, {12288, 2048, 128, 0}, {8192, 8192, 128, 1}, {4096, 4096, 2048, 0}};
        std::vector<double> ref, got;
        for (auto q : ps) {
            ref.assign((size_t)q.r * ld, 0.0); got.assign((size_t)q.r * ld, 0.0);
            g_dgemm_tma = 0;
            cudaMemset(C, 0, (size_t)q.r * ld * 8);
            run<128, 64, GM_SUB>(A + 128 * ld + 256, B + 256 * ld + 512, C, ld, q.r, q.c, q.k, q.lower, 0);
            cudaMemcpy(ref.data(), C, (size_t)q.r * ld * 8, cudaMemcpyDeviceToHost);
            size_t bad_total = 0;
            for (int rep = 0; rep < 5; rep++) {
                g_dgemm_tma = 7;
                cudaMemset(C, 0, (size_t)q.r * ld * 8);
                run<128, 64, GM_SUB>(A + 128 * ld + 256, B + 256 * ld + 512, C, ld, q.r, q.c, q.k, q.lower, 0);
                cudaMemcpy(got.data(), C, (size_t)q.r * ld * 8, cudaMemcpyDeviceToHost);
                size_t bad = 0;
                for (int64_t i = 0; i < q.r; i++)
                    for (int64_t j = 0; j < q.c; j++) bad += got[i * ld + j] != ref[i * ld + j];
                bad_total += bad;
            }
            printf("check persistent TMA SUB %lld x %lld x %d lower=%d: mismatching entries over 5 runs: %zu\n", (long long)q.r, (long long)q.c, q.k, q.lower, bad_total);
        }
        // in-place leaf X <- X Dinv^T (64x128 tiles, k = 128) on 16384 rows = 256 tiles
        {
            const int64_t r = 16384;
            ref.assign((size_t)r * ld, 0.0); got.assign((size_t)r * ld, 0.0);
            size_t bad_total = 0;
            for (int rep = 0; rep < 6; rep++) {
                g_dgemm_tma = rep == 0 ? 0 : 7;
                cudaMemcpy(C, A, (size_t)r * ld * 8, cudaMemcpyDeviceToDevice);
                dgemm_nt_configure<64, 128, GM_SET>();
                dgemm_nt_launch<64, 128, GM_SET>(0, C + 384, ld, B + 5 * 128 * ld, ld, C + 384, ld, r, 128, 128, 0, 0, 0);   // ONE in-place pass (exact)
                if (cudaDeviceSynchronize() != cudaSuccess) { printf("in-place SET failed\n"); exit(1); }
                cudaMemcpy((rep == 0 ? ref : got).data(), C, (size_t)r * ld * 8, cudaMemcpyDeviceToHost);
                if (rep == 0) continue;
                size_t bad = 0;
                for (int64_t i = 0; i < r; i++)
                    for (int64_t j = 0; j < 1024; j++) bad += got[i * ld + j] != ref[i * ld + j];
                bad_total += bad;
            }
            printf("check persistent TMA in-place SET 16384 x 128 x 128: mismatching entries over 5 runs: %zu\n", bad_total);
        }
        g_dgemm_tma = 0;
        cudaMemset(C, 0, (size_t)R * ld * 8);
    }
    // kernel variants on a full-wave shape and on the k = 128 update shape (same arithmetic, different staging / warp tiling):
    //   BK=32 x 2 stages: half the barriers per flop;  32x64 warp tiles, 4 warps per 64x128 CTA: 0.375 instead of 0.5 LDS per DMMA
    //   (what the cuBLAS kernel of the same tile, cutlass_80_tensorop_d884gemm_64x128_16x3, uses);  128x128 CTA, 1 per SM
    struct V { const char* name; double (*fn)(const double*, const double*, double*, int64_t, int64_t, int64_t, int, int, int); int bm, bn; };
    V variants[] = {
        {"128x64 BK16x3 w32x32 (product)", run<128, 64, GM_SUB>, 128, 64},
        {"128x64 BK32x2 w32x32", run<128, 64, GM_SUB, 32, 2>, 128, 64},
        {"64x128 BK16x3 w32x64 (4 warps)", run<64, 128, GM_SUB, 16, 3, 32, 64>, 64, 128},
        {"128x128 BK16x3 w64x32 (1 CTA/SM)", run<128, 128, GM_SUB, 16, 3, 64, 32, 1>, 128, 128},
        {"128x128 BK16x4 w32x64 (1 CTA/SM)", run<128, 128, GM_SUB, 16, 4, 32, 64, 1>, 128, 128}};
    {   // every variant accumulates k in the same order: results must be bit-identical to the product configuration
        const int64_t r = 1024, c = 1024; const int k = 512;
        std::vector<double> ref((size_t)r * ld), got((size_t)r * ld);
        for (auto& v : variants) {
            cudaMemset(C, 0, (size_t)r * ld * 8);
            v.fn(A, B, C, ld, r, c, k, 0, 0);   // reps = 0: the two warm-up launches only (C = -2 A B^T)
            cudaMemcpy(got.data(), C, (size_t)r * ld * 8, cudaMemcpyDeviceToHost);
            if (&v == &variants[0]) ref = got;
            size_t bad = 0;
            for (int64_t i = 0; i < r; i++)
                for (int64_t j = 0; j < c; j++) bad += got[i * ld + j] != ref[i * ld + j];
            printf("check %-36s mismatching entries vs product: %zu  (C[0,0] = %.6f)\n", v.name, bad, got[0]);
        }
        cudaMemset(C, 0, (size_t)R * ld * 8);
    }
    struct S2 { int64_t r, c; int k; } vs[] = {{8192, 8192, 8192}, {10112, 4096, 4096}, {10112, 8192, 128}, {10112, 256, 256}};
    printf("\n%-36s", "variant \\ TFLOP/s at shape");
    for (auto s : vs) printf(" %6lldx%lldx%d", (long long)s.r, (long long)s.c, s.k);
    printf("\n");
    for (auto& v : variants) {
        printf("%-36s", v.name);
        for (auto s : vs) {
            const double ms = v.fn(A, B, C, ld, s.r, s.c, s.k, 0, 6);
            printf(" %18.2f", 2.0 * s.r * s.c * s.k / ms / 1e9);
        }
        printf("\n");
    }
    return 0;
}
