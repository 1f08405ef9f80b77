// fp64_peak.cu — DFMA / DADD / DMUL issue-rate micro-benchmark (measure the FP64
// ceiling on the actual B200 before quoting a fraction of it).  Prints one JSON line.
#include <cuda_runtime.h>
#include <cstdio>

template <int MODE>
__global__ void k_fp64(double* out, int iters, double a, double b) {
    double x0 = threadIdx.x * 1e-3, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; i++) {
        if (MODE == 0) {
            x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
            x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
        } else if (MODE == 1) {
            x0 = __dadd_rn(x0, b); x1 = __dadd_rn(x1, b); x2 = __dadd_rn(x2, b); x3 = __dadd_rn(x3, b);
            x4 = __dadd_rn(x4, b); x5 = __dadd_rn(x5, b); x6 = __dadd_rn(x6, b); x7 = __dadd_rn(x7, b);
        } else {
            x0 = __dmul_rn(x0, a); x1 = __dmul_rn(x1, a); x2 = __dmul_rn(x2, a); x3 = __dmul_rn(x3, a);
            x4 = __dmul_rn(x4, a); x5 = __dmul_rn(x5, a); x6 = __dmul_rn(x6, a); x7 = __dmul_rn(x7, a);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

template <int MODE>
static double run(int sms, double* out) {
    const int iters = 4096, threads = 1024, blocks = sms * 2;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k_fp64<MODE><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; r++) {
        cudaEventRecord(e0);
        k_fp64<MODE><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    double ops = (double)blocks * threads * iters * 8.0;
    return ops / (best * 1e-3);  // thread-instructions per second
}

int main() {
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, 0) != cudaSuccess) {
        printf("{\"error\": \"no device\"}\n");
        return 1;
    }
    double* out;
    cudaMalloc(&out, sizeof(double) * p.multiProcessorCount * 2 * 1024);
    double fma_rate = run<0>(p.multiProcessorCount, out);
    double add_rate = run<1>(p.multiProcessorCount, out);
    double mul_rate = run<2>(p.multiProcessorCount, out);
    int clk = 0;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"cc\": \"%d.%d\", \"clock_khz\": %d, \"dfma_per_s\": %.4e, \"dadd_per_s\": %.4e, "
           "\"dmul_per_s\": %.4e, \"fp64_tflops_fma\": %.2f, \"l2_bytes\": %d, \"smem_per_sm\": %zu}\n",
           p.name, p.multiProcessorCount, p.major, p.minor, clk, fma_rate, add_rate, mul_rate, 2.0 * fma_rate / 1e12,
           p.l2CacheSize, p.sharedMemPerMultiprocessor);
    return 0;
}
