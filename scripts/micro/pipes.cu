// Instruction-throughput probe for the pipes the feature kernels live on (run on the GPU box):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu && ./pipes
// Each test runs ITER x UNROLL copies of one instruction on NCH independent chains per thread and prints warp
// instructions per clock per SM at 8 / 16 / 32 resident warps per SM.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define ITER 2048

template <int OP>
__global__ void probe(float* out, int n_iter, float seed) {
    float f[8];
    double d[8];
    unsigned long long u[8];
    __shared__ double sm[2048];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        f[i] = seed + i + threadIdx.x;
        d[i] = seed + i + threadIdx.x;
        u[i] = ((unsigned long long)__float_as_uint(seed + i) << 32) | __float_as_uint(seed + threadIdx.x);
    }
    if (threadIdx.x < 2048) sm[threadIdx.x % 2048] = seed;
    __syncthreads();
    const float a = seed * 0.5f, b = seed * 0.25f;
    const double da = seed * 0.5, db = seed * 0.25;
    const unsigned long long ua = ((unsigned long long)__float_as_uint(a) << 32) | __float_as_uint(b);
    uint32_t saddr = (uint32_t)__cvta_generic_to_shared(sm) + (threadIdx.x & 31) * 16;
    for (int it = 0; it < n_iter; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (OP == 0) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(a), "f"(b));
                if (OP == 1) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(u[i]) : "l"(ua), "l"(ua));
                if (OP == 2) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(da), "d"(db));
                if (OP == 3) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(d[i]) : "d"(da));
                if (OP == 4) asm volatile("mul.rn.f64 %0, %0, %1;" : "+d"(d[i]) : "d"(da));
                if (OP == 5) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(u[i]) : "l"(ua));
                if (OP == 6) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(a));
                if (OP == 7) { float2 v; asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(saddr + (i & 3) * 512)); f[i] += v.x; }
                if (OP == 8) { double2 v; asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(saddr + (i & 3) * 512)); d[i] += v.x; }
                if (OP == 9) asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(saddr + (i & 3) * 512), "d"(d[i]), "d"(d[(i + 1) & 7]) : "memory");
                if (OP == 10) f[i] = __shfl_xor_sync(0xffffffffu, f[i], 1 + (i & 3));
                if (OP == 11) { float t; asm volatile("cvt.rn.f32.f64 %0, %1;" : "=f"(t) : "d"(d[i])); f[i] += t; }
                if (OP == 12) { double t; asm volatile("cvt.f64.f32 %0, %1;" : "=d"(t) : "f"(f[i])); d[i] += t; }
                if (OP == 13) asm volatile("lg2.approx.f32 %0, %0;" : "+f"(f[i]));
                if (OP == 14) {   // FFMA2 + FFMA interleaved (one each)
                    asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(u[i]) : "l"(ua), "l"(ua));
                    asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(a), "f"(b));
                }
                if (OP == 15) {   // DFMA + FFMA interleaved
                    asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(da), "d"(db));
                    asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(a), "f"(b));
                }
                if (OP == 16) {   // DFMA + integer add interleaved
                    asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(da), "d"(db));
                    asm volatile("add.u64 %0, %0, %1;" : "+l"(u[i]) : "l"(ua));
                }
                if (OP == 17) asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(f[i]));   // single-register-operand form
            }
        }
    }
    float acc = 0.0f;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc += f[i] + (float)d[i] + (float)(u[i] & 0xffff);
    if (acc == 12345.678f) out[0] = acc;
}

template <int OP>
void run(const char* name, int per_iter_mult) {
    float* out;
    cudaMalloc(&out, 4);
    cudaEvent_t s, e;
    cudaEventCreate(&s);
    cudaEventCreate(&e);
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int clk_khz = 0;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    printf("%-28s", name);
    for (int warps = 8; warps <= 32; warps *= 2) {
        probe<OP><<<sms, warps * 32>>>(out, 64, 1.0f);
        cudaEventRecord(s);
        probe<OP><<<sms, warps * 32>>>(out, ITER, 1.0f);
        cudaEventRecord(e);
        cudaEventSynchronize(e);
        float ms = 0;
        cudaEventElapsedTime(&ms, s, e);
        const double instr = (double)ITER * 32 * per_iter_mult * warps;          // warp instructions per SM
        const double clocks = ms * 1e-3 * clk_khz * 1e3;
        printf("  %2d warps: %6.3f /clk/SM", warps, instr / clocks);
    }
    printf("\n");
    cudaFree(out);
}

int main() {
    int clk_khz = 0;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    printf("clock attr %d kHz (rates below assume the SMs ran at it)\n", clk_khz);
    run<0>("FFMA (3 reg)", 1);
    run<17>("FFMA (1 reg)", 1);
    run<1>("FFMA2", 1);
    run<5>("FMUL2", 1);
    run<6>("FADD", 1);
    run<14>("FFMA2+FFMA pairs", 2);
    run<2>("DFMA", 1);
    run<3>("DADD", 1);
    run<4>("DMUL", 1);
    run<15>("DFMA+FFMA pairs", 2);
    run<16>("DFMA+IADD64 pairs", 2);
    // (LDS / F2F probes: their loop-invariant operands get merged by ptxas, the printed rates are not meaningful)
    run<9>("STS.128", 1);
    run<10>("SHFL", 1);
    run<13>("MUFU.LG2", 1);
    return 0;
}
