// Micro-benchmark: FP64 issue rates on sm_100a (B200).  Measures the roofs every
// "fraction of FP64 roofline" statement in DESIGN.md/bench.py is quoted against:
//   (1) DMMA  mma.sync.m8n8k4.f64      (2) DMMA m16n8k16.f64 (PTX sm_90+ shape)
//   (3) DFMA  scalar fma.rn.f64        (4) a device-to-device copy (HBM roof cross-check)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ub_fp64 ub_fp64.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

template <int NACC>
__global__ void k_dmma884(double* out, int iters, double a0, double b0) {
    double c[NACC][2];
#pragma unroll
    for (int i = 0; i < NACC; i++) { c[i][0] = 0.0; c[i][1] = 0.0; }
    double a = a0 + threadIdx.x * 1e-9, b = b0;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++) {
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void k_dmma16816(double* out, int iters, double a0, double b0) {
    double c[NACC][4];
#pragma unroll
    for (int i = 0; i < NACC; i++) { c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.0; }
    double a[8], b[4];
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = a0 + threadIdx.x * 1e-9 + i;
#pragma unroll
    for (int i = 0; i < 4; i++) b[i] = b0 + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++) {
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
                         : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                         : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                           "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void k_dfma(double* out, int iters, double a0, double b0) {
    double c[NACC];
#pragma unroll
    for (int i = 0; i < NACC; i++) c[i] = i;
    double a = a0 + threadIdx.x * 1e-9, b = b0;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++) c[i] = fma(a, c[i], b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; i++) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_exp(double* out, int iters, double a0) {
    double x = a0 + threadIdx.x * 1e-6, s = 0;
    for (int it = 0; it < iters; it++) { s += exp(-x); x += 1e-7; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_copy(const double2* __restrict__ in, double2* __restrict__ out, size_t n) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += st) out[i] = in[i];
}

template <class F> float timeit(F f, int reps = 5) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    f(); f(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; r++) {
        CK(cudaEventRecord(e0)); f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int sms = p.multiProcessorCount;
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz\": %d,\n", p.name, sms, p.clockRate);
    double* out; CK(cudaMalloc(&out, sizeof(double) * sms * 32 * 1024));
    const int iters = 20000;
    // sweep warps per SM for DMMA m8n8k4: blocks = sms * bps, threads = tpb
    int cfgs[][2] = {{1, 128}, {1, 256}, {1, 512}, {2, 512}, {1, 1024}, {2, 1024}};
    for (auto& c : cfgs) {
        int blocks = sms * c[0], tpb = c[1];
        float ms = timeit([&] { k_dmma884<8><<<blocks, tpb>>>(out, iters, 1.0, 1e-3); });
        double fl = 2.0 * 8 * 8 * 4 * 8.0 * iters * (double)blocks * (tpb / 32);
        printf(" \"dmma884_acc8_b%d_t%d_tflops\": %.3f,\n", c[0], tpb, fl / ms * 1e-9);
    }
    {
        int blocks = sms * 2, tpb = 512;
        float ms = timeit([&] { k_dmma884<32><<<blocks, tpb>>>(out, iters / 4, 1.0, 1e-3); });
        double fl = 2.0 * 8 * 8 * 4 * 32.0 * (iters / 4) * (double)blocks * (tpb / 32);
        printf(" \"dmma884_acc32_b2_t512_tflops\": %.3f,\n", fl / ms * 1e-9);
        ms = timeit([&] { k_dmma884<2><<<blocks, tpb>>>(out, iters, 1.0, 1e-3); });
        fl = 2.0 * 8 * 8 * 4 * 2.0 * iters * (double)blocks * (tpb / 32);
        printf(" \"dmma884_acc2_b2_t512_tflops\": %.3f,\n", fl / ms * 1e-9);
        ms = timeit([&] { k_dmma884<1><<<sms, 32>>>(out, iters, 1.0, 1e-3); });
        printf(" \"dmma884_dependent_latency_ns\": %.3f,\n", ms * 1e6 / iters);
        ms = timeit([&] { k_dmma16816<8><<<blocks, tpb>>>(out, iters / 4, 1.0, 1e-3); });
        fl = 2.0 * 16 * 8 * 16 * 8.0 * (iters / 4) * (double)blocks * (tpb / 32);
        printf(" \"dmma16816_acc8_b2_t512_tflops\": %.3f,\n", fl / ms * 1e-9);
    }
    for (auto& c : cfgs) {
        int blocks = sms * c[0], tpb = c[1];
        float ms = timeit([&] { k_dfma<16><<<blocks, tpb>>>(out, iters, 1.0000001, 1e-9); });
        double fl = 2.0 * 16.0 * iters * (double)blocks * tpb;
        printf(" \"dfma_acc16_b%d_t%d_tflops\": %.3f,\n", c[0], tpb, fl / ms * 1e-9);
    }
    {
        int blocks = sms * 2, tpb = 1024, it = 2000;
        float ms = timeit([&] { k_exp<<<blocks, tpb>>>(out, it, 0.5); });
        printf(" \"exp_f64_Gexp_per_s\": %.3f,\n", (double)it * blocks * tpb / ms * 1e-6);
    }
    {
        size_t n = (size_t)1 << 27;  // 2 GiB per buffer as double2
        double2 *a, *b; CK(cudaMalloc(&a, n * 16)); CK(cudaMalloc(&b, n * 16));
        CK(cudaMemset(a, 1, n * 16));
        float ms = timeit([&] { k_copy<<<sms * 16, 512>>>(a, b, n); });
        printf(" \"copy_gbs\": %.1f,\n", 2.0 * n * 16 / ms * 1e-6);
        ms = timeit([&] { CK(cudaMemsetAsync(b, 0, n * 16)); });
        printf(" \"memset_gbs\": %.1f\n", 1.0 * n * 16 / ms * 1e-6);
    }
    printf("}\n");
    return 0;
}
