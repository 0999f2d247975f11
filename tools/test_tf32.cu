// Stand-alone check of the tcgen05 split-TF32 GEMM (dev tool): C -= A B^T against an fp64 host reference, plus timing.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo tools/test_tf32.cu -o tools/test_tf32 -lcuda
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <random>
#include "../gumbi_b200/csrc/tf32gemm.cuh"
using namespace gb2;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 2; } } while (0)

int run(int M, int N, int K, int lower, bool check, int nbt = 2) {
    std::mt19937_64 rng(1234);
    std::normal_distribution<double> nd(0.0, 1.0);
    std::vector<double> A((size_t)M * K), B((size_t)N * K), C((size_t)M * N), C0;
    for (auto& v : A) v = nd(rng);
    for (auto& v : B) v = nd(rng);
    for (auto& v : C) v = nd(rng);
    C0 = C;
    double *dA, *dB, *dC; float *Ahi, *Alo, *Bhi, *Blo;
    CK(cudaMalloc(&dA, A.size() * 8)); CK(cudaMalloc(&dB, B.size() * 8)); CK(cudaMalloc(&dC, C.size() * 8));
    CK(cudaMalloc(&Ahi, A.size() * 4)); CK(cudaMalloc(&Alo, A.size() * 4)); CK(cudaMalloc(&Bhi, B.size() * 4)); CK(cudaMalloc(&Blo, B.size() * 4));
    CK(cudaMemcpy(dA, A.data(), A.size() * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, B.data(), B.size() * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dC, C.data(), C.size() * 8, cudaMemcpyHostToDevice));
    tc::split_tf32_kernel<<<(unsigned)(((size_t)M * K / 2 + 255) / 256), 256>>>(dA, K, M, K, Ahi, Alo, K);
    tc::split_tf32_kernel<<<(unsigned)(((size_t)N * K / 2 + 255) / 256), 256>>>(dB, K, N, K, Bhi, Blo, K);
    CK(cudaDeviceSynchronize());
    CUtensorMap mAhi, mAlo, mBhi, mBlo;
    CUresult r;
    if ((r = tc::make_tmap(&mAhi, Ahi, M, K, K)) || (r = tc::make_tmap(&mAlo, Alo, M, K, K)) || (r = tc::make_tmap(&mBhi, Bhi, N, K, K)) ||
        (r = tc::make_tmap(&mBlo, Blo, N, K, K))) { printf("cuTensorMapEncodeTiled failed: %d\n", (int)r); return 3; }
    CK(tc::gemm_tf32x3_configure());
    int n_sm = 0; CK(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, 0));
    tc::GemmArgs g{}; g.C = dC; g.ldc = N; g.n_bi = M / 128; g.n_bj = N / 128; g.lower = lower; g.rb_first = 0; g.rb_stride = 1; g.rb_local_first = -1; g.cblk0 = 0;
    g.a_k0 = g.b_row0 = g.b_k0 = 0;
    int launches = 0;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    tc::gemm_tf32x3_launch(0, n_sm, mAhi, mAlo, mBhi, mBlo, g, K, launches, nbt);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    int rc = 0;
    if (check) {
        CK(cudaMemcpy(C.data(), dC, C.size() * 8, cudaMemcpyDeviceToHost));
        double maxerr = 0, maxref = 0; long bad = 0;
        for (int i = 0; i < M; i++)
            for (int j = 0; j < N; j++) {
                const bool live = !lower || (j / 128 <= i / 128);
                double ref = C0[(size_t)i * N + j];
                if (live) { double s = 0; for (int k = 0; k < K; k++) s += A[(size_t)i * K + k] * B[(size_t)j * K + k]; ref -= s; }
                const double err = fabs(C[(size_t)i * N + j] - ref);
                if (err > maxerr) maxerr = err;
                if (fabs(ref) > maxref) maxref = fabs(ref);
                if (err > 1e-3 * sqrt((double)K)) { if (bad < 5) printf("  bad (%d,%d): got %.9g want %.9g\n", i, j, C[(size_t)i * N + j], ref); bad++; }
            }
        printf("check nbt=%d M=%d N=%d K=%d lower=%d launches=%d: max|err|=%.3e (max|ref|=%.3e, rel %.2e) bad=%ld -> %s\n", nbt, M, N, K, lower, launches,
               maxerr, maxref, maxerr / maxref, bad, bad ? "FAIL" : "ok");
        rc = bad ? 1 : 0;
    } else {
        const int reps = 5;
        cudaEventRecord(e0);
        for (int i = 0; i < reps; i++) tc::gemm_tf32x3_launch(0, n_sm, mAhi, mAlo, mBhi, mBlo, g, K, launches, nbt);
        cudaEventRecord(e1);
        CK(cudaDeviceSynchronize());
        float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= reps;
        const double tiles = lower ? (double)g.n_bi * (g.n_bi + 1) / 2 : (double)g.n_bi * g.n_bj;
        const double flop = 2.0 * tiles * 128 * 128 * K;
        printf("time  nbt=%d M=%d N=%d K=%d lower=%d: %.3f ms  %.1f TFLOP/s fp64-equivalent (x3 tf32 MMA = %.1f TFLOP/s tensor)\n", nbt, M, N, K, lower, ms,
               flop / ms / 1e9, 3 * flop / ms / 1e9);
    }
    cudaFree(dA); cudaFree(dB); cudaFree(dC); cudaFree(Ahi); cudaFree(Alo); cudaFree(Bhi); cudaFree(Blo);
    return rc;
}

int main() {
    int rc = 0;
    for (int nbt = 1; nbt <= 2; nbt++) {
        rc |= run(128, 128, 32, 0, true, nbt);
        rc |= run(256, 384, 512, 0, true, nbt);
        rc |= run(1024, 1024, 512, 1, true, nbt);
        rc |= run(2432, 2432, 256, 1, true, nbt);            // odd number of 128-column blocks (19): ragged last CTA tile
        rc |= run(2560, 2432, 2048 + 1024, 0, true, nbt);    // K split over 3 launches, 19 column blocks
    }
    if (rc) { printf("CORRECTNESS FAILED\n"); return rc; }
    for (int nbt = 1; nbt <= 2; nbt++) {
        run(8192, 8192, 512, 1, false, nbt);
        run(32768, 32768, 512, 1, false, nbt);
        run(10112, 8192, 1024, 0, false, nbt);
        run(10112, 16384, 4096, 0, false, nbt);
    }
    return 0;
}
