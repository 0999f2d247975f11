// Stand-alone rates of the fp64 DMMA GEMM (gumbi_b200/csrc/dgemm.cuh) on the shapes the predict solve and the Cholesky trailing
// update launch, against the wave count of each launch (dev tool; not part of the product path).  Answers: how far is the kernel
// itself from the cuBLAS DGEMM figure on a full-wave problem, and how much do partial last waves cost on the recursion's shapes.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I gumbi_b200/csrc -o tools/micro_dgemm tools/micro_dgemm.cu
#include "dgemm.cuh"
#include <cstdio>
#include <vector>
using namespace gb2;

template <int BM, int BN, int MODE>
static double run(const double* A, const double* B, double* C, int64_t ld, int64_t rows, int64_t cols, int k, int lower, int reps) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 2; i++) dgemm_nt_launch<BM, BN, MODE>(0, A, ld, B, ld, C, ld, rows, cols, k, lower, 0, 0);
    cudaEventRecord(e0);
    for (int i = 0; i < reps; i++) dgemm_nt_launch<BM, BN, MODE>(0, A, ld, B, ld, C, ld, rows, cols, k, lower, 0, 0);
    cudaEventRecord(e1);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); exit(1); }
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms / reps;
}

int main() {
    const int64_t ld = 16384, R = 16384;
    double *A, *B, *C;
    cudaMalloc(&A, R * ld * 8); cudaMalloc(&B, R * ld * 8); cudaMalloc(&C, R * ld * 8);
    std::vector<double> h((size_t)R * ld);
    for (size_t i = 0; i < h.size(); i++) h[i] = (double)((i * 2654435761u) >> 8 & 0xffff) / 65536.0 - 0.5;
    cudaMemcpy(A, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(B, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
    cudaMemset(C, 0, h.size() * 8);
    dgemm_nt_configure<128, 64, GM_SUB>(); dgemm_nt_configure<64, 128, GM_SET>();
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
    return 0;
}
