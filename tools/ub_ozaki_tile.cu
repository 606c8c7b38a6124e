// Second building block for DESIGN.md section 11 item 1: one FP64 GEMM tile C[128 x 64] = A[128 x 64] * B[64 x 64]^T
// computed entirely with INT8 tensor-core instructions (Ozaki scheme), inside one CTA:
//   1. every thread scales its row of A (and of B) by a power of two and cuts it into S = 9 signed 7-bit slices, written
//      straight into the canonical K-major core-matrix layout tcgen05 reads from shared memory;
//   2. for each scale group g = t + u (smallest first) one thread issues the tcgen05.mma kind::i8 products of all slice
//      pairs (t, u) of the group into one INT32 TMEM accumulator (exact), commits to an mbarrier;
//   3. all threads read their accumulator row with tcgen05.ld and add 2^(-7 (g + 2)) * value into FP64 registers;
//   4. rows / columns are scaled back.
// The result is compared with a long-double host product and with the plain FP64 product's own rounding error.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -Igp_emu_uqsa_b200/csrc tools/ub_ozaki_tile.cu -o tools/ub_ozaki_tile.bin
//   timeout 20 tools/ub_ozaki_tile.bin
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <vector>
#include "gpe_common.cuh"

constexpr int M = 128, N = 64, K = 64, KI = 32, NK = K / KI, S = 9, BITS = 7;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
constexpr uint32_t IDESC = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
__device__ __forceinline__ int canon(int r, int k, int R) {
    return (k / KI) * (R * KI) + ((k % KI) / 16) * (R * 16) + (r / 8) * 128 + (r % 8) * 16 + (k % 16);
}

// scale a row by 2^-e (|x| 2^-e < 1/2), cut it into S slices of BITS bits; returns e
__device__ int slice_row(const double* __restrict__ x, int row, int R, uint8_t* slices /* [S][R*K] */) {
    double amax = 0.0;
    for (int k = 0; k < K; k++) amax = fmax(amax, fabs(x[k]));
    int e = 0;
    if (amax > 0.0) { (void)frexp(amax, &e); e += 1; }
    for (int k = 0; k < K; k++) {
        double r = scalbn(x[k], -e);
        const int o = canon(row, k, R);
#pragma unroll
        for (int t = 0; t < S; t++) {
            r *= (double)(1 << BITS);
            const double q = trunc(r);
            r -= q;
            slices[(size_t)t * R * K + o] = (uint8_t)(int8_t)(int)q;
        }
    }
    return e;
}

__global__ void __launch_bounds__(128, 1) ozaki_tile_kernel(const double* __restrict__ A, const double* __restrict__ B,
                                                            double* __restrict__ C) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sA = smem;                            // [S][M*K]
    uint8_t* sB = smem + (size_t)S * M * K;        // [S][N*K]
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    __shared__ int eB[N];
    const int tid = threadIdx.x, warp = tid >> 5;
    const int eA = slice_row(A + (size_t)tid * K, tid, M, sA);
    if (tid < N) eB[tid] = slice_row(B + (size_t)tid * K, tid, N, sB);
    if (tid == 0) gpe::mbar_init(&bar, 1);
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base_s)), "n"(N));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tmem = tmem_base_s;

    double acc[N];
#pragma unroll
    for (int j = 0; j < N; j++) acc[j] = 0.0;
    uint32_t parity = 0;
    for (int g = S - 1; g >= 0; g--) {             // scale groups, smallest contributions first
        if (tid == 0) {
            bool first = true;
            for (int t = 0; t <= g; t++) {
                const int u = g - t;
                for (int kb = 0; kb < NK; kb++) {
                    const uint64_t da = umma_desc(smem_u32(sA + (size_t)t * M * K + kb * M * KI), M * 16, 128);
                    const uint64_t db = umma_desc(smem_u32(sB + (size_t)u * N * K + kb * N * KI), N * 16, 128);
                    const uint32_t accf = first ? 0u : 1u;
                    first = false;
                    asm volatile(
                        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}\n" ::"r"(tmem),
                        "l"(da), "l"(db), "r"(IDESC), "r"(accf), "r"(0u), "r"(0u), "r"(0u), "r"(0u)
                        : "memory");
                }
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(&bar)) : "memory");
        }
        gpe::mbar_wait(&bar, parity);
        parity ^= 1u;
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        const double scale = scalbn(1.0, -BITS * (g + 2));
#pragma unroll
        for (int c = 0; c < N; c += 8) {
            uint32_t v[8];
            const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c;
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
                         : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                         : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
            for (int j = 0; j < 8; j++) acc[c + j] = fma((double)(int32_t)v[j], scale, acc[c + j]);
        }
        asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
        __syncthreads();                          // everyone has read the accumulator before the next group overwrites it
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    }
#pragma unroll
    for (int j = 0; j < N; j++) C[(size_t)tid * N + j] = scalbn(acc[j], eA + eB[j]);
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "n"(N));
}

int main() {
    std::vector<double> a((size_t)M * K), b((size_t)N * K), c((size_t)M * N);
    unsigned long long s = 88172645463325252ull;
    auto rnd = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return (double)(s >> 11) / 9007199254740992.0 - 0.5; };
    for (int i = 0; i < M; i++)                     // rows with very different magnitudes and entries spread over 12 decades
        for (int k = 0; k < K; k++) a[(size_t)i * K + k] = rnd() * std::pow(10.0, (i % 7) - 3) * std::pow(10.0, -12.0 * std::fabs(rnd()));
    for (int j = 0; j < N; j++)
        for (int k = 0; k < K; k++) b[(size_t)j * K + k] = rnd() * std::pow(10.0, (j % 5) - 2);
    double *dA, *dB, *dC;
    cudaMalloc(&dA, a.size() * 8); cudaMalloc(&dB, b.size() * 8); cudaMalloc(&dC, c.size() * 8);
    cudaMemcpy(dA, a.data(), a.size() * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, b.data(), b.size() * 8, cudaMemcpyHostToDevice);
    const size_t smem = (size_t)S * (M + N) * K;
    cudaFuncSetAttribute(ozaki_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    ozaki_tile_kernel<<<1, 128, smem>>>(dA, dB, dC);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("{\"error\": \"%s\"}\n", cudaGetErrorString(e)); return 1; }
    cudaMemcpy(c.data(), dC, c.size() * 8, cudaMemcpyDeviceToHost);
    double worst_oz = 0.0, worst_f64 = 0.0;         // errors relative to |a_i|^T |b_j|, the componentwise GEMM bound
    for (int i = 0; i < M; i++)
        for (int j = 0; j < N; j++) {
            long double ref = 0.0L, mag = 0.0L;
            double f64 = 0.0;
            for (int k = 0; k < K; k++) {
                ref += (long double)a[(size_t)i * K + k] * (long double)b[(size_t)j * K + k];
                mag += fabsl((long double)a[(size_t)i * K + k] * (long double)b[(size_t)j * K + k]);
                f64 = std::fma(a[(size_t)i * K + k], b[(size_t)j * K + k], f64);
            }
            worst_oz = std::fmax(worst_oz, (double)(fabsl((long double)c[(size_t)i * N + j] - ref) / mag));
            worst_f64 = std::fmax(worst_f64, (double)(fabsl((long double)f64 - ref) / mag));
        }
    printf("{\"tile\": \"%dx%dx%d\", \"slices\": %d, \"int8_mma_per_tile\": %d, \"max_err_ozaki_rel_abs_product\": %.3e, "
           "\"max_err_plain_fp64_rel_abs_product\": %.3e}\n", M, N, K, S, S * (S + 1) / 2 * NK, worst_oz, worst_f64);
    return 0;
}
