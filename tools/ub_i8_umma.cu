// Building block for DESIGN.md section 11 item 1 (FP64 emulation on the INT8 tensor cores): a hand-written
// tcgen05.mma kind::i8 with operands in shared memory (K-major, no swizzle, canonical 8 x 16-byte core matrices),
// the INT32 accumulator in TMEM, read back with tcgen05.ld -- checked bit for bit against a host INT32 product, then
// timed on all SMs.  Stand-alone (no TMA yet: tiles are staged with ordinary stores + fence.proxy.async).
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -Igp_emu_uqsa_b200/csrc tools/ub_i8_umma.cu -o tools/ub_i8_umma.bin
//   timeout 20 tools/ub_i8_umma.bin
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "gpe_common.cuh"

constexpr int M = 128;        // UMMA M (rows of D = TMEM lanes)
constexpr int N = 256;        // UMMA N (columns of D = TMEM columns, one 32-bit column per n)
constexpr int KI = 32;        // K of one kind::i8 instruction
constexpr int NK = 4;         // k blocks staged in shared memory (K = 128)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start address >> 4 in [0,14), leading byte offset >> 4
// in [16,30), stride byte offset >> 4 in [32,46), version = 1 in [46,48), layout type (0 = no swizzle) in [61,64)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
// instruction descriptor (cute::UMMA::InstrDescriptor): c_format S32 = 2 at [4,6), a/b format INT8 = 1 at [7,10) /
// [10,13), both operands K-major (bits 15, 16 = 0), N >> 3 at [17,23), M >> 4 at [24,29)
constexpr uint32_t IDESC = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);

// element (row r, k) of an R x (NK*32) K-major operand in the canonical layout:
// [k block][16-byte k chunk (2)][8-row group][row in group (8)][16 bytes]  ->  SBO = 128 B, LBO = R * 16 B
__host__ __device__ inline int canon(int r, int k, int R) {
    return (k / KI) * (R * KI) + ((k % KI) / 16) * (R * 16) + (r / 8) * 128 + (r % 8) * 16 + (k % 16);
}

__global__ void __launch_bounds__(128, 1) umma_i8_kernel(const int8_t* __restrict__ A, const int8_t* __restrict__ B,
                                                         int32_t* __restrict__ D, int reps, int write_out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sA = smem;                       // M x 128 bytes, canonical layout
    uint8_t* sB = smem + M * NK * KI;         // N x 128 bytes
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int e = tid; e < M * NK * KI; e += 128) sA[e] = (uint8_t)A[e];          // inputs are stored canonical already
    for (int e = tid; e < N * NK * KI; e += 128) sB[e] = (uint8_t)B[e];
    if (tid == 0) gpe::mbar_init(&bar, 1);
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");               // generic-proxy stores -> async proxy
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base_s)), "n"(N));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tmem = tmem_base_s;

    if (tid == 0) {                           // one thread issues every MMA of the CTA
        for (int r = 0; r < reps; r++) {
#pragma unroll
            for (int kb = 0; kb < NK; kb++) {
                const uint64_t da = umma_desc(smem_u32(sA + kb * M * KI), M * 16, 128);
                const uint64_t db = umma_desc(smem_u32(sB + kb * N * KI), N * 16, 128);
                const uint32_t acc = (r > 0 || kb > 0) ? 1u : 0u;
                asm volatile(
                    "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                    "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}\n" ::"r"(tmem),
                    "l"(da), "l"(db), "r"(IDESC), "r"(acc), "r"(0u), "r"(0u), "r"(0u), "r"(0u)
                    : "memory");
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(&bar)) : "memory");
    }
    gpe::mbar_wait(&bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    // D: TMEM lane = row, column = n.  Warp w reads lanes 32w .. 32w+31, eight columns per instruction.
    const int row = tid;
    for (int c = 0; c < N; c += 8) {
        uint32_t v[8];
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                     : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
        if (write_out) {
#pragma unroll
            for (int j = 0; j < 8; j++) D[((size_t)blockIdx.x * M + row) * N + c + j] = (int32_t)v[j];
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "n"(N));
}

int main() {
    const int K = NK * KI;
    std::vector<int8_t> a((size_t)M * K), b((size_t)N * K), ac((size_t)M * K), bc((size_t)N * K);
    unsigned long long s = 88172645463325252ull;
    auto rnd = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return (int8_t)((int)(s >> 33) % 127 - 63); };
    for (auto& v : a) v = rnd();
    for (auto& v : b) v = rnd();
    for (int r = 0; r < M; r++) for (int k = 0; k < K; k++) ac[canon(r, k, M)] = a[(size_t)r * K + k];
    for (int r = 0; r < N; r++) for (int k = 0; k < K; k++) bc[canon(r, k, N)] = b[(size_t)r * K + k];
    int8_t *dA, *dB; int32_t* dD;
    const int nblk = 148;
    cudaMalloc(&dA, ac.size()); cudaMalloc(&dB, bc.size()); cudaMalloc(&dD, sizeof(int32_t) * (size_t)nblk * M * N);
    cudaMemcpy(dA, ac.data(), ac.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(dB, bc.data(), bc.size(), cudaMemcpyHostToDevice);
    const size_t smem = (size_t)(M + N) * K;
    cudaFuncSetAttribute(umma_i8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    umma_i8_kernel<<<1, 128, smem>>>(dA, dB, dD, 1, 1);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("{\"error\": \"%s\"}\n", cudaGetErrorString(e)); return 1; }
    std::vector<int32_t> d((size_t)M * N);
    cudaMemcpy(d.data(), dD, sizeof(int32_t) * d.size(), cudaMemcpyDeviceToHost);
    long long bad = 0;
    for (int i = 0; i < M; i++)
        for (int j = 0; j < N; j++) {
            int32_t ref = 0;
            for (int k = 0; k < K; k++) ref += (int32_t)a[(size_t)i * K + k] * (int32_t)b[(size_t)j * K + k];
            if (ref != d[(size_t)i * N + j]) { if (bad < 5) printf("mismatch (%d,%d): got %d want %d\n", i, j, d[(size_t)i * N + j], ref); bad++; }
        }
    printf("{\"shape\": \"%dx%dx%d int8 -> int32\", \"mismatches\": %lld", M, N, K, bad);
    if (bad == 0) {
        const int reps = 20000;
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        for (int blocks : {1, nblk}) {
            umma_i8_kernel<<<blocks, 128, smem>>>(dA, dB, dD, reps, 0);
            cudaEventRecord(e0);
            umma_i8_kernel<<<blocks, 128, smem>>>(dA, dB, dD, reps, 0);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            printf(", \"int8_TOPs_%d_ctas\": %.1f", blocks, 2.0 * M * N * K * (double)reps * blocks / ms * 1e-9);
        }
    }
    printf("}\n");
    e = cudaGetLastError();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    return bad != 0;
}
