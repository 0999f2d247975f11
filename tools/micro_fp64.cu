// Micro-benchmarks that size the fp64 design on B200 (sm_100a):
//   * DFMA issue rate, DMMA (mma.sync m8n8k4 f64) rate, both together
//   * fp64 exp: libm exp() vs the table-free polynomial used by the K-build
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o micro_fp64 micro_fp64.cu
// Prints one JSON object per line.  Not part of the product path.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int NACC>
__global__ void k_dfma(double* out, int iters, double a, double b) {
    double acc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; i++) acc[i] = threadIdx.x * 1e-9 + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; i++) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void k_dmma(double* out, int iters, double a, double b) {
    double c0[NACC], c1[NACC];
#pragma unroll
    for (int i = 0; i < NACC; i++) { c0[i] = i; c1[i] = -i; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++) dmma884(c0[i], c1[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; i++) s += c0[i] + c1[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// half the warps run DFMA, half DMMA
template <int NACC>
__global__ void k_mixed(double* out, int iters, double a, double b) {
    int warp = threadIdx.x >> 5;
    double s = 0;
    if (warp & 1) {
        double acc[NACC];
#pragma unroll
        for (int i = 0; i < NACC; i++) acc[i] = threadIdx.x * 1e-9 + i;
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int i = 0; i < NACC; i++) acc[i] = fma(acc[i], a, b);
        }
#pragma unroll
        for (int i = 0; i < NACC; i++) s += acc[i];
    } else {
        double c0[NACC], c1[NACC];
#pragma unroll
        for (int i = 0; i < NACC; i++) { c0[i] = i; c1[i] = -i; }
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int i = 0; i < NACC; i++) dmma884(c0[i], c1[i], a, b);
        }
#pragma unroll
        for (int i = 0; i < NACC; i++) s += c0[i] + c1[i];
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// exp(-t), t>=0, Cody-Waite + degree-12 Taylor/minimax-ish polynomial (same as kbuild.cuh)
__device__ __forceinline__ double exp_neg_poly(double t) {
    const double L2E = 1.4426950408889634074;
    const double LN2_HI = 6.93147180369123816490e-01, LN2_LO = 1.90821492927058770002e-10;
    double x = -t;
    double kf = fma(x, L2E, 6755399441055744.0);   // round-to-nearest via magic add
    int k = __double2loint(kf);
    kf -= 6755399441055744.0;
    double r = fma(-kf, LN2_HI, x);
    r = fma(-kf, LN2_LO, r);
    double p = 2.0876756987868099e-09;
    p = fma(p, r, 2.5052108385441720e-08);
    p = fma(p, r, 2.7557319223985888e-07);
    p = fma(p, r, 2.7557319223985893e-06);
    p = fma(p, r, 2.4801587301587302e-05);
    p = fma(p, r, 1.9841269841269841e-04);
    p = fma(p, r, 1.3888888888888889e-03);
    p = fma(p, r, 8.3333333333333332e-03);
    p = fma(p, r, 4.1666666666666664e-02);
    p = fma(p, r, 1.6666666666666666e-01);
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    // scale by 2^k (k in [-1075, 0]); flush below 2^-1000 to 0
    if (k < -1000) return 0.0;
    int hi = __double2hiint(p) + (k << 20);
    return __hiloint2double(hi, __double2loint(p));
}

template <int MODE>
__global__ void k_exp(double* out, const double* in, int n, int reps) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    double t = in[i % n];
    double s = 0;
    for (int r = 0; r < reps; r++) {
        double v = (MODE == 0) ? exp(-t) : exp_neg_poly(t);
        s += v; t += 1e-3;
    }
    out[i] = s;
}

__global__ void k_exp_err(const double* in, int n, double* maxrel) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double a = exp(-in[i]), b = exp_neg_poly(in[i]);
    double rel = fabs(a - b) / fmax(a, 1e-300);
    atomicMax((unsigned long long*)maxrel, __double_as_longlong(rel));
}

// dependent-chain latencies with one warp (clock64 deltas / iters)
__global__ void k_latency(long long* out, int iters, double a, double b) {
    double x = threadIdx.x * 1e-9;
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) x = fma(x, a, b);
    long long t1 = clock64();
    double c0 = x, c1 = -x;
    for (int i = 0; i < iters; i++) dmma884(c0, c1, a, b);
    long long t2 = clock64();
    double y = x;
    for (int i = 0; i < iters; i++) y = rsqrt(y + 1.5);
    long long t3 = clock64();
    double z = y;
    for (int i = 0; i < iters; i++) z = __shfl_sync(0xffffffffu, z, (threadIdx.x + 1) & 31);
    long long t4 = clock64();
    float f = (float)z;
    for (int i = 0; i < iters; i++) f = rsqrtf(f + 1.5f);
    long long t5 = clock64();
    double w = z;
    for (int i = 0; i < iters; i++) w = (double)(float)w + 1.0;
    long long t6 = clock64();
    // independent DFMA issue rate from one warp: 8 accumulators
    double acc[8];
#pragma unroll
    for (int i = 0; i < 8; i++) acc[i] = i + x;
    long long t7 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) acc[i] = fma(acc[i], a, b);
    }
    long long t8 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += acc[i];
    if (threadIdx.x == 0) {
        out[0] = t1 - t0; out[1] = t2 - t1; out[2] = t3 - t2; out[3] = t4 - t3; out[4] = t5 - t4; out[5] = t6 - t5; out[6] = t8 - t7;
        out[7] = (long long)(c0 + c1 + s + f + w);
    }
}

template <typename F>
float time_ms(F f, int reps = 5) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int i = 0; i < reps; i++) {
        cudaEventRecord(e0); f(); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int sms = p.multiProcessorCount;
    printf("{\"device\":\"%s\",\"sms\":%d,\"clock_khz\":%d}\n", p.name, sms, p.clockRate);
    double* out; CK(cudaMalloc(&out, sizeof(double) * sms * 8 * 1024));
    {
        long long* lat; CK(cudaMalloc(&lat, 64)); long long hl[8];
        k_latency<<<1, 32>>>(lat, 2048, 1.0000001, 1e-9); CK(cudaDeviceSynchronize());
        k_latency<<<1, 32>>>(lat, 2048, 1.0000001, 1e-9); CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(hl, lat, 64, cudaMemcpyDeviceToHost));
        printf("{\"bench\":\"latency_cycles\",\"dfma\":%.1f,\"dmma884\":%.1f,\"rsqrt_f64\":%.1f,\"shfl64\":%.1f,\"rsqrt_f32\":%.1f,\"cvt_roundtrip_add\":%.1f,\"dfma_issue_1warp\":%.2f}\n",
               hl[0] / 2048.0, hl[1] / 2048.0, hl[2] / 2048.0, hl[3] / 2048.0, hl[4] / 2048.0, hl[5] / 2048.0, hl[6] / 2048.0 / 8);
    }
    const int iters = 4096;
    for (int warps : {4, 8, 16, 32}) {
        int threads = warps * 32, blocks = sms * 2;
        if (threads > 1024) { threads = 1024; }
        double nth = (double)blocks * threads;
        float ms = time_ms([&] { k_dfma<8><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
        printf("{\"bench\":\"dfma\",\"warps_per_cta\":%d,\"ctas\":%d,\"tflops\":%.2f}\n", warps, blocks, 2.0 * nth * iters * 8 / ms / 1e9);
        ms = time_ms([&] { k_dmma<8><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
        printf("{\"bench\":\"dmma884\",\"warps_per_cta\":%d,\"ctas\":%d,\"tflops\":%.2f}\n", warps, blocks, 2.0 * (nth / 32) * iters * 8 * 256 / ms / 1e9);
        ms = time_ms([&] { k_dmma<16><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
        printf("{\"bench\":\"dmma884_acc16\",\"warps_per_cta\":%d,\"ctas\":%d,\"tflops\":%.2f}\n", warps, blocks, 2.0 * (nth / 32) * iters * 16 * 256 / ms / 1e9);
        ms = time_ms([&] { k_mixed<8><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
        double fl = 2.0 * (nth / 2) * iters * 8 + 2.0 * (nth / 64) * iters * 8 * 256;
        printf("{\"bench\":\"mixed_dfma_dmma\",\"warps_per_cta\":%d,\"ctas\":%d,\"tflops\":%.2f}\n", warps, blocks, fl / ms / 1e9);
    }
    // exp
    int n = 1 << 20; double* in; CK(cudaMalloc(&in, n * sizeof(double)));
    double* h = (double*)malloc(n * sizeof(double));
    for (int i = 0; i < n; i++) h[i] = 60.0 * rand() / RAND_MAX;
    h[0] = 0; h[1] = 1e-12; h[2] = 700.0; h[3] = 745.0;
    CK(cudaMemcpy(in, h, n * sizeof(double), cudaMemcpyHostToDevice));
    double* mr; CK(cudaMalloc(&mr, 8)); CK(cudaMemset(mr, 0, 8));
    k_exp_err<<<n / 256, 256>>>(in, n, mr);
    double hmr; CK(cudaMemcpy(&hmr, mr, 8, cudaMemcpyDeviceToHost));
    printf("{\"bench\":\"exp_poly_maxrel_vs_libm\",\"value\":%.3e}\n", hmr);
    int reps = 256; int blocks = sms * 8, threads = 256;
    double* out2; CK(cudaMalloc(&out2, sizeof(double) * blocks * threads));
    float ms0 = time_ms([&] { k_exp<0><<<blocks, threads>>>(out2, in, n, reps); });
    float ms1 = time_ms([&] { k_exp<1><<<blocks, threads>>>(out2, in, n, reps); });
    double ne = (double)blocks * threads * reps;
    printf("{\"bench\":\"exp_libm\",\"gexp_per_s\":%.1f}\n", ne / ms0 / 1e6);
    printf("{\"bench\":\"exp_poly\",\"gexp_per_s\":%.1f}\n", ne / ms1 / 1e6);
    return 0;
}
