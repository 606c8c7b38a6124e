// Issue rates of the integer / FP32 instructions the residue conversion and the CRT pass are made of (development aid):
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/ub_intops.bin tools/ub_intops.cu && tools/ub_intops.bin
// Each kernel runs ITER x 8 independent chains of one instruction per thread; the result is lane-instructions per clock per SM.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int ITER = 4096;

#define CHAINS8(...)                       \
    _Pragma("unroll 4") for (int it = 0; it < ITER; it++) { \
        _Pragma("unroll") for (int c = 0; c < 8; c++) { __VA_ARGS__ } \
    }

template <int OP>
__global__ void __launch_bounds__(256) k(uint32_t* out, uint32_t a, uint32_t b, long long* clk, uint32_t zero) {
    uint32_t x[8];
    float f[8];
    double dd[8];
    unsigned long long w[8];
#pragma unroll
    for (int c = 0; c < 8; c++) { x[c] = threadIdx.x * 8 + c + a; f[c] = (float)x[c]; dd[c] = (double)x[c]; w[c] = x[c]; }
    const float fa = __uint_as_float(a), fb = __uint_as_float(b);
    long long t0 = clock64();
    if (OP == 0) CHAINS8(asm volatile("dp4a.u32.u32 %0, %0, %1, %2;" : "+r"(x[c]) : "r"(a), "r"(b));)
    if (OP == 1) CHAINS8(asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[c]) : "r"(a), "r"(b));)
    if (OP == 2) CHAINS8(asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(x[c]) : "r"(a));)
    if (OP == 3) CHAINS8(asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[c]) : "r"(x[c]), "r"(a));)
    if (OP == 4) CHAINS8(asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[c]) : "f"(fa), "f"(fb));)
    if (OP == 5) CHAINS8(asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(f[c]) : "f"(fa));)
    if (OP == 6) CHAINS8(asm volatile("prmt.b32 %0, %0, %1, 0x5410;" : "+r"(x[c]) : "r"(a));)
    if (OP == 7) CHAINS8(asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[c]) : "r"(a), "r"(b));)
    if (OP == 8) CHAINS8(asm volatile("fma.rn.f64 %0, %0, %1, %1;" : "+d"(dd[c]) : "d"((double)fa));)
    if (OP == 9) CHAINS8(asm volatile("{ .reg .s64 t; cvt.rzi.s64.f64 t, %0; cvt.u32.u64 %1, t; }" : "+d"(dd[c]), "+r"(x[c]));)
    if (OP == 10) CHAINS8(asm volatile("fma.rm.f32 %0, %0, %1, %2;" : "+f"(f[c]) : "f"(fa), "f"(fb));)
    if (OP == 11) CHAINS8(asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(x[c]) : "r"(a), "r"(b));)
    if (OP == 12) CHAINS8(asm volatile("shf.r.wrap.b32 %0, %0, %1, %2;" : "+r"(x[c]) : "r"(a), "r"(b));)
    if (OP == 13) CHAINS8(asm volatile("dp2a.lo.u32.u32 %0, %0, %1, %2;" : "+r"(x[c]) : "r"(a), "r"(b));)
    if (OP == 14) CHAINS8(asm volatile("mul.f64 %0, %0, %1;" : "+d"(dd[c]) : "d"((double)fa));)
    if (OP == 15) CHAINS8(asm volatile("cvt.rn.f32.u32 %0, %1;" : "=f"(f[c]) : "r"(x[c])); asm volatile("mov.b32 %0, %1;" : "=r"(x[c]) : "f"(f[c]));)
    if (OP == 16) CHAINS8(asm volatile("{ .reg .b32 lo, hi; mov.b64 {lo, hi}, %0; mad.wide.u32 %0, lo, %1, %0; }" : "+l"(w[c]) : "r"(a));)
    // the residue step of the conversion as it is: two IDP.4A, IMAD.HI, IMAD (t -> t mod p)
    if (OP == 17) CHAINS8(uint32_t t = __dp4a(x[c], a, b); t = __dp4a(x[c] ^ 0x5a5a5a5au, b, t);
                          x[c] = __umulhi(t, 0x01010102u) * (0u - 255u) + t;)
    // the same with the quotient taken from the upper half of a 32 x 32 -> 64 product whose lower half is kept alive
    if (OP == 18) CHAINS8(uint32_t t = __dp4a(x[c], a, b); t = __dp4a(x[c] ^ 0x5a5a5a5au, b, t);
                          uint32_t lo, hi;
                          asm volatile("{ .reg .b64 ww; mul.wide.u32 ww, %2, %3; mov.b64 {%0, %1}, ww; }" : "=r"(lo), "=r"(hi) : "r"(t), "r"(0x01010102u));
                          x[c] = hi * (0u - 255u) + (t | (lo & zero));)
    long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int c = 0; c < 8; c++) s += x[c] + __float_as_uint(f[c]) + (uint32_t)__double2hiint(dd[c]) + (uint32_t)w[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

template <int OP>
static void run(const char* name, int ctas_per_sm) {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int grid = sms * ctas_per_sm;
    uint32_t* out;
    long long* clk;
    cudaMalloc(&out, (size_t)grid * 256 * 4);
    cudaMalloc(&clk, grid * sizeof(long long));
    k<OP><<<grid, 256>>>(out, 0x01020304u, 0x05060708u, clk, 0u);
    cudaDeviceSynchronize();
    k<OP><<<grid, 256>>>(out, 0x01020304u, 0x05060708u, clk, 0u);
    cudaDeviceSynchronize();
    long long h[2048];
    cudaMemcpy(h, clk, grid * sizeof(long long), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < grid; i++) avg += (double)h[i];
    avg /= grid;
    // all CTAs of an SM run concurrently: lanes per clock per SM = ctas_per_sm * 256 threads * ITER * 8 / cycles
    printf("{\"op\": \"%s\", \"ctas_per_sm\": %d, \"lane_instr_per_clk_per_sm\": %.1f, \"err\": \"%s\"}\n", name, ctas_per_sm,
           (double)ctas_per_sm * 256.0 * ITER * 8.0 / avg, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out);
    cudaFree(clk);
}

int main() {
    for (int c : {2, 4}) {
        run<0>("dp4a.u32.u32 (IDP.4A)", c);
        run<13>("dp2a.lo.u32.u32 (IDP.2A)", c);
        run<1>("mad.lo.u32 (IMAD)", c);
        run<2>("mul.hi.u32 (IMAD.HI)", c);
        run<11>("mad.hi.u32 (IMAD.HI with addend)", c);
        run<3>("mad.wide.u32 (IMAD.WIDE)", c);
        run<4>("fma.rn.f32 (FFMA)", c);
        run<10>("fma.rm.f32 (FFMA.RM)", c);
        run<5>("add.rn.f32 (FADD)", c);
        run<6>("prmt.b32 (PRMT)", c);
        run<7>("lop3.b32 (LOP3)", c);
        run<12>("shf.r.wrap (SHF)", c);
        run<8>("fma.rn.f64 (DFMA)", c);
        run<14>("mul.f64 (DMUL)", c);
        run<9>("cvt.rzi.s64.f64 (F2I.S64.F64)", c);
        run<15>("cvt.rn.f32.u32 (I2F)", c);
        run<16>("mad.wide.u32 on a changing operand (IMAD.WIDE)", c);
        run<17>("residue step: 2 IDP.4A + IMAD.HI + IMAD (per step)", c);
        run<18>("residue step: 2 IDP.4A + IMAD.WIDE + LOP3 + IMAD (per step)", c);
    }
    return 0;
}
