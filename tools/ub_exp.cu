// Accuracy and throughput of gpe_exp (gp_emu_uqsa_b200/csrc/gpe_common.cuh) against CUDA's exp(double).
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -Igp_emu_uqsa_b200/csrc tools/ub_exp.cu -o tools/ub_exp.bin
//   tools/ub_exp.bin
// Arguments: -(u^2) * 745 for u uniform in [0,1) (dense near 0, reaching the denormal range), plus a sweep of
// exact edge values.  Reports the largest distance in ulps between the two and against a long-double host value.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "gpe_common.cuh"

__global__ void both(const double* x, double* a, double* b, size_t n) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < n) { a[i] = gpe::gpe_exp(x[i]); b[i] = exp(x[i]); }
}
template <int WHICH>
__global__ void rate(double* out, int iters) {
    double x = -1e-3 * (threadIdx.x + 1), acc = 0.0;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 8; u++) { acc += WHICH ? gpe::gpe_exp(x) : exp(x); x -= 0.37; if (x < -600.0) x += 599.0; }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
static long long ulps(double a, double b) {
    long long ia, ib;
    memcpy(&ia, &a, 8); memcpy(&ib, &b, 8);
    return llabs(ia - ib);
}
int main() {
    const size_t n = (size_t)1 << 26;
    std::vector<double> hx(n), ha(n), hb(n);
    unsigned long long s = 88172645463325252ull;
    for (size_t i = 0; i < n; i++) {
        s ^= s << 13; s ^= s >> 7; s ^= s << 17;
        double u = (double)(s >> 11) / 9007199254740992.0;
        hx[i] = -(u * u) * 745.0;
    }
    const double edge[] = {0.0, -0.0, -1e-300, -1e-17, -0.34657359027997264, -0.3465735902799727, -0.69314718055994531, -1.0,
                           -707.9999, -708.0, -708.3964185322641, -744.0, -745.2, -1000.0, 1.0, 700.0, 709.7, 710.0};
    for (size_t i = 0; i < sizeof edge / sizeof *edge; i++) hx[i] = edge[i];
    double *dx, *da, *db;
    cudaMalloc(&dx, n * 8); cudaMalloc(&da, n * 8); cudaMalloc(&db, n * 8);
    cudaMemcpy(dx, hx.data(), n * 8, cudaMemcpyHostToDevice);
    both<<<(unsigned)((n + 255) / 256), 256>>>(dx, da, db, n);
    cudaMemcpy(ha.data(), da, n * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(hb.data(), db, n * 8, cudaMemcpyDeviceToHost);
    long long worst = 0, worst_ref_a = 0, worst_ref_b = 0, differ = 0;
    for (size_t i = 0; i < n; i++) {
        long long d = ulps(ha[i], hb[i]);
        if (d) differ++;
        if (d > worst) worst = d;
        if ((i & 63) == 0) {                 // long-double host value on a subsample
            double ref = (double)expl((long double)hx[i]);
            long long ra = ulps(ha[i], ref), rb = ulps(hb[i], ref);
            if (ra > worst_ref_a) worst_ref_a = ra;
            if (rb > worst_ref_b) worst_ref_b = rb;
        }
    }
    printf("{\"n\": %zu, \"max_ulp_gpe_vs_cuda\": %lld, \"fraction_differing\": %.4f, \"max_ulp_gpe_vs_host_long_double\": %lld, "
           "\"max_ulp_cuda_vs_host_long_double\": %lld", n, worst, (double)differ / n, worst_ref_a, worst_ref_b);
    double* out; cudaMalloc(&out, 148 * 8 * 256 * 8);
    for (int which = 0; which < 2; which++) {
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        for (int rep = 0; rep < 2; rep++) {
            cudaEventRecord(e0);
            if (which) rate<1><<<148 * 8, 256>>>(out, 2000); else rate<0><<<148 * 8, 256>>>(out, 2000);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
        }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf(", \"%s_Gexp_per_s\": %.1f", which ? "gpe_exp" : "cuda_exp", 148.0 * 8 * 256 * 2000 * 8 / ms * 1e-6);
    }
    printf("}\n");
    return 0;
}
