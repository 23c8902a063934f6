// Microbenchmark (sm_100a): issue cost of packed f32x2 vs scalar FP32 arithmetic, bf16 packing and MUFU on one SM sub-partition.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o build/fma_pipe scripts/micro/fma_pipe.cu && ./build/fma_pipe
// Each variant runs ITER iterations of 16 independent operations per thread; warps per SM and the clock are reported so the
// result reads as "SM clocks per warp instruction per sub-partition".
#include <cstdio>
#include <cuda_runtime.h>
constexpr int ITER = 4096;
template <int MODE> __global__ void k(float* out, long long* clk, float a, float b) {
  float2 x[8]; float y[16];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = make_float2(a + i, b - i);
#pragma unroll
  for (int i = 0; i < 16; ++i) y[i] = a * i + b;
  const float2 m = make_float2(a, a), c = make_float2(b, b);
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITER; ++it) {
    if (MODE == 0) {          // 16 scalar FFMA (register operands)
#pragma unroll
      for (int i = 0; i < 16; ++i) y[i] = fmaf(y[i], a, b);
    } else if (MODE == 1) {   // 8 packed FFMA2 (= 16 lane-FMAs)
#pragma unroll
      for (int i = 0; i < 8; ++i) x[i] = __ffma2_rn(x[i], m, c);
    } else if (MODE == 2) {   // 8 packed + 16 scalar, independent
#pragma unroll
      for (int i = 0; i < 8; ++i) { x[i] = __ffma2_rn(x[i], m, c); y[2 * i] = fmaf(y[2 * i], a, b); y[2 * i + 1] = fmaf(y[2 * i + 1], a, b); }
    } else if (MODE == 3) {   // 8 packed FMUL2
#pragma unroll
      for (int i = 0; i < 8; ++i) x[i] = __fmul2_rn(x[i], m);
    } else if (MODE == 4) {   // 8 bf16x2 packs + 8 unpack shifts
#pragma unroll
      for (int i = 0; i < 8; ++i) { unsigned r; asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(x[i].y), "f"(x[i].x)); x[i].x = __uint_as_float(r << 16); }
    } else if (MODE == 5) {   // 16 MUFU.EX2
#pragma unroll
      for (int i = 0; i < 16; ++i) asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y[i]) : "f"(y[i]));
    } else if (MODE == 6) {   // 16 scalar FFMA with an immediate multiplicand
#pragma unroll
      for (int i = 0; i < 16; ++i) y[i] = fmaf(y[i], 1.0009765625f, b);
    } else if (MODE == 7) {   // 16 LOP3 (ALU pipe) + 16 scalar FFMA
#pragma unroll
      for (int i = 0; i < 16; ++i) { y[i] = fmaf(y[i], a, b); unsigned u = __float_as_uint(x[i & 7].x); u = (u & 0x7fffffffu) | 0x00400000u; x[i & 7].x = __uint_as_float(u); }
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x[i].x + x[i].y;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += y[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}
template <int MODE> void run(const char* name, int ninstr, int warps) {
  float* out; long long* clk;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&clk, 148 * 8);
  k<MODE><<<148, warps * 32>>>(out, clk, 1.0001f, 0.5f);
  k<MODE><<<148, warps * 32>>>(out, clk, 1.0001f, 0.5f);
  long long h[148];
  cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
  const double per = avg / ((double)ITER * ninstr * (warps / 4.0));
  printf("%-52s %2d warps/SM: %.2f clk per warp instruction per sub-partition\n", name, warps, per);
  cudaFree(out); cudaFree(clk);
}
int main() {
  for (int w : {4, 8, 16}) {
    if (w == 4) {
      run<0>("FFMA scalar (reg, reg, reg)", 16, 4); run<1>("FFMA2 packed", 8, 4); run<2>("8 FFMA2 + 16 FFMA interleaved (per instruction)", 24, 4);
      run<3>("FMUL2 packed", 8, 4); run<4>("F2FP.BF16 pack + shift (per pair of instructions)", 8, 4); run<5>("MUFU.EX2", 16, 4);
      run<6>("FFMA scalar, immediate operand", 16, 4); run<7>("FFMA + 2 LOP3 interleaved (per FFMA)", 16, 4);
    } else if (w == 8) {
      run<0>("FFMA scalar (reg, reg, reg)", 16, 8); run<1>("FFMA2 packed", 8, 8); run<2>("8 FFMA2 + 16 FFMA interleaved (per instruction)", 24, 8);
      run<3>("FMUL2 packed", 8, 8); run<4>("F2FP.BF16 pack + shift (per pair of instructions)", 8, 8); run<5>("MUFU.EX2", 16, 8);
      run<6>("FFMA scalar, immediate operand", 16, 8); run<7>("FFMA + 2 LOP3 interleaved (per FFMA)", 16, 8);
    } else {
      run<0>("FFMA scalar (reg, reg, reg)", 16, 16); run<1>("FFMA2 packed", 8, 16); run<2>("8 FFMA2 + 16 FFMA interleaved (per instruction)", 24, 16);
      run<3>("FMUL2 packed", 8, 16); run<4>("F2FP.BF16 pack + shift (per pair of instructions)", 8, 16); run<5>("MUFU.EX2", 16, 16);
      run<6>("FFMA scalar, immediate operand", 16, 16); run<7>("FFMA + 2 LOP3 interleaved (per FFMA)", 16, 16);
    }
  }
  return 0;
}
