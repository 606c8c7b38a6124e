// Where the 128x128 leaf (Cholesky + triangular inverse in shared memory) spends its cycles.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -DGPE_LEAF_TIMING -Iinclude -Igp_emu_uqsa_b200/csrc \
//        tools/leaf_phases.cu -o tools/leaf_phases.bin && tools/leaf_phases.bin
// Includes the product's kernel source with the phase hooks compiled in (thread 0 adds clock64() differences
// between barriers to g_leaf_cyc[]); the library itself is built without them.
#include <cstdio>
#include <vector>
#include "../gp_emu_uqsa_b200/csrc/gpe_kernels.cu"

int main() {
    const int n = 128, B = 1, reps = 20;
    std::vector<double> A((size_t)n * n), M((size_t)n * n);
    unsigned long long s = 88172645463325252ull;
    for (auto& v : M) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; v = (double)(s >> 11) / 9007199254740992.0 - 0.5; }
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) {
            double acc = (i == j) ? 8.0 : 0.0;
            for (int k = 0; k < n; k++) acc += M[i * n + k] * M[j * n + k];
            A[i * n + j] = acc;
        }
    double *dA, *dL, *dF, *dld; int* dst;
    cudaMalloc(&dA, sizeof(double) * n * n); cudaMalloc(&dL, sizeof(double) * n * n); cudaMalloc(&dF, sizeof(double) * n * n);
    cudaMalloc(&dld, 8); cudaMalloc(&dst, 4); cudaMemset(dst, 0, 4);
    cudaMemcpy(dA, A.data(), sizeof(double) * n * n, cudaMemcpyHostToDevice);
    gpe::launch_leaf(dA, dL, n, 0, 0, 0, dld, 1, dst, B, 0, dF);      // warm-up
    cudaDeviceSynchronize();
    unsigned long long zero[16] = {0};
    cudaMemcpyToSymbol(gpe::g_leaf_cyc, zero, sizeof zero);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    for (int r = 0; r < reps; r++) gpe::launch_leaf(dA, dL, n, 0, 0, 0, dld, 1, dst, B, 0, dF);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    unsigned long long cyc[16];
    cudaMemcpyFromSymbol(cyc, gpe::g_leaf_cyc, sizeof cyc);
    const char* name[9] = {"load", "potrf: first 8x8 diagonal factor", "potrf: panel solve", "potrf: trailing update || next diagonal factor",
                           "log-det + status", "trtri: 8x8 diagonal inverses", "trtri: T = L21 X11", "trtri: X21 = -X22 T", "store"};
    double tot = 0;
    for (int i = 0; i < 9; i++) tot += (double)cyc[i] / reps;
    printf("leaf: %.1f us per launch (events, back-to-back launches of one block); clock64 total %.0f cycles\n", ms * 1e3 / reps, tot);
    for (int i = 0; i < 9; i++) printf("  %-40s %8.0f cycles  %5.1f %%\n", name[i], (double)cyc[i] / reps, 100.0 * cyc[i] / reps / tot);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
