// Phase clocks of the diagonal-panel kernel on one SPD 128x128 block (dev tool; not part of the product path).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I gumbi_b200/csrc -o tools/micro_potrf tools/micro_potrf.cu
#include "cholesky.cuh"
#include <cmath>
#include <vector>
using namespace gb2;
int main() {
    const int n = TILE;
    std::vector<double> A(n * n);
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) A[i * n + j] = exp(-0.5 * (i - j) * (i - j) / 400.0) + (i == j ? 0.01 : 0.0);
    double *dA, *dD; int* dInfo; long long* dClk;
    cudaMalloc(&dA, n * n * 8); cudaMalloc(&dD, n * n * 8); cudaMalloc(&dInfo, 4); cudaMalloc(&dClk, 64 * 8);
    cudaMemset(dInfo, 0, 4); cudaMemset(dD, 0, n * n * 8);
    cudaFuncSetAttribute(potrf_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PD_SMEM);
    long long clk[64];
    for (int rep = 0; rep < 3; rep++) {
        cudaMemcpy(dA, A.data(), n * n * 8, cudaMemcpyHostToDevice);
        cudaMemset(dClk, 0, 64 * 8);
        potrf_diag_kernel<<<1, PD_THREADS, PD_SMEM>>>(dA, n, 0, n, dD, dInfo, nullptr, PushArgs{}, dClk);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        cudaMemcpy(clk, dClk, 64 * 8, cudaMemcpyDeviceToHost);
        printf("rep %d:", rep);
        for (int i = 1; i < 64 && clk[i]; i++) printf(" %lld", clk[i] - clk[i - 1]);
        printf("  total %lld\n", clk[0] ? 0LL : 0LL);
    }
    // check L L^T = A and Dinv L = I
    std::vector<double> L(n * n), D(n * n);
    cudaMemcpy(L.data(), dA, n * n * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(D.data(), dD, n * n * 8, cudaMemcpyDeviceToHost);
    double e1 = 0, e2 = 0;
    for (int i = 0; i < n; i++)
        for (int j = 0; j <= i; j++) {
            double s = 0, t = 0;
            for (int k = 0; k <= j; k++) s += L[i * n + k] * L[j * n + k];
            for (int k = j; k <= i; k++) t += D[i * n + k] * L[k * n + j];
            e1 = fmax(e1, fabs(s - A[i * n + j]));
            e2 = fmax(e2, fabs(t - (i == j ? 1.0 : 0.0)));
        }
    printf("max|LL^T-A| = %.3e  max|inv(L) L - I| = %.3e\n", e1, e2);

    return 0;
}
