// tc_ptx.h — PTX wrappers shared by the tcgen05/TMA kernels of this library (logistic_tc.cu, gauss_tc.cu):
// mbarriers, TMA loads, tcgen05 fences / commit / MMA issue (batched asm blocks), TMEM loads and stores,
// shared-memory matrix descriptors (128-byte swizzle) and the host-side tensor-map encoder.
// Include inside `namespace bn { namespace { ... } }` of a .cu file (everything here has internal linkage).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>

constexpr int CHUNK_BYTES = 128 * 128;  // 128 rows x 64 bf16 (one SW128 box)

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// try_wait suspends the thread in hardware until the phase completes or the time hint
// expires, so waiting warps do not burn issue slots or clog the MIO queue with polls
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t a = smem_u32(bar);
  uint32_t done;
  do {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(done)
        : "r"(a), "r"(parity), "r"(20000u)
        : "memory");
  } while (!done);
}
// non-blocking phase test
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(done)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return done != 0;
}
__device__ __forceinline__ void tma_load_2d(const void* tmap, void* dst, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
// L2 prefetch of a tile a few blocks ahead of its load, so the load itself never waits on HBM
__device__ __forceinline__ void tma_prefetch_2d(const void* tmap, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];" ::"l"(tmap), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void bulk_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// fire-and-forget add of four consecutive floats in global memory (16-byte aligned)
__device__ __forceinline__ void red_add_f32x4(float4* p, const float4 v) {
  asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
// one elected lane of a converged warp (the form ptxas turns into ELECT + a predicated instruction)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem] · B[smem]
__device__ __forceinline__ void mma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc)
      : "memory");
}
// D[tmem] (+)= A[tmem] · B[smem]
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc)
      : "memory");
}

// ---- batched MMA issue.  One asm block = one elect.sync + up to four tcgen05.mma, so the per-instruction
// issue cost (predicate conversion, descriptor moves into uniform registers) is paid once per block;
// measured: with one elect per MMA the issuing warp needed ~70 clk per instruction, more than the 56-64 clk
// the tensor pipe needs to execute it, and was the bottleneck of the whole kernel.
#define BN_MMA_SS(ACC) "mov.b64 ra, {al, %2};\n\tmov.b64 rb, {bl, %4};\n\t@pe tcgen05.mma.cta_group::1.kind::f16 [%0], ra, rb, %5, " ACC ";\n\t"
#define BN_MMA_TS(ACC) "mov.b64 rb, {bl, %3};\n\t@pe tcgen05.mma.cta_group::1.kind::f16 [%0], [ta], rb, %4, " ACC ";\n\t"
#define BN_SS_HEAD "{\n\t.reg .pred pe, pa, pt;\n\t.reg .b64 ra, rb;\n\t.reg .b32 al, bl;\n\t" \
                   "elect.sync _|pe, 0xffffffff;\n\tsetp.ne.b32 pa, %6, 0;\n\tsetp.eq.b32 pt, 0, 0;\n\tmov.b32 al, %1;\n\tmov.b32 bl, %3;\n\t"
#define BN_SS_STEP "add.u32 al, al, 2;\n\tadd.u32 bl, bl, 2;\n\t"
#define BN_SS_ARGS ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(acc_first) : "memory"
// GEMM1: N4 consecutive K steps inside one 64-column chunk.  Only the 14-bit start-address field of a
// descriptor changes (+32 B = 2 units; shared-memory addresses stay below 2^18, so no carry leaves the
// field): the descriptors travel as 32-bit halves and the high halves are loop constants.
template <int N4>
__device__ __forceinline__ void mma_ss_run(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                           uint32_t idesc, uint32_t acc_first) {
  static_assert(N4 >= 1 && N4 <= 4, "1..4 K steps per chunk");
  if constexpr (N4 == 1)
    asm volatile(BN_SS_HEAD BN_MMA_SS("pa") "}\n" BN_SS_ARGS);
  else if constexpr (N4 == 2)
    asm volatile(BN_SS_HEAD BN_MMA_SS("pa") BN_SS_STEP BN_MMA_SS("pt") "}\n" BN_SS_ARGS);
  else if constexpr (N4 == 3)
    asm volatile(BN_SS_HEAD BN_MMA_SS("pa") BN_SS_STEP BN_MMA_SS("pt") BN_SS_STEP BN_MMA_SS("pt") "}\n" BN_SS_ARGS);
  else
    asm volatile(BN_SS_HEAD BN_MMA_SS("pa") BN_SS_STEP BN_MMA_SS("pt") BN_SS_STEP BN_MMA_SS("pt") BN_SS_STEP BN_MMA_SS("pt") "}\n"
                 BN_SS_ARGS);
}
// GEMM2: the four K steps (16 data rows each) of one 64-row half block for one residual term.  A advances
// 8 TMEM columns, then 24 to the next 32-row chunk (its hi or lo half), B advances 2048 B = 128 units.
__device__ __forceinline__ void mma_ts_run4(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                            uint32_t acc_first) {
  asm volatile("{\n\t.reg .pred pe, pa, pt;\n\t.reg .b64 rb;\n\t.reg .b32 ta, bl;\n\t"
               "elect.sync _|pe, 0xffffffff;\n\tsetp.ne.b32 pa, %5, 0;\n\tsetp.eq.b32 pt, 0, 0;\n\t"
               "mov.b32 ta, %1;\n\tmov.b32 bl, %2;\n\t" BN_MMA_TS("pa")
               "add.u32 ta, ta, 8;\n\tadd.u32 bl, bl, 128;\n\t" BN_MMA_TS("pt")
               "add.u32 ta, ta, 24;\n\tadd.u32 bl, bl, 128;\n\t" BN_MMA_TS("pt")
               "add.u32 ta, ta, 8;\n\tadd.u32 bl, bl, 128;\n\t" BN_MMA_TS("pt") "}\n"
               ::"r"(d_tmem), "r"(a_tmem), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(acc_first) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// shared-memory matrix descriptor, 128-byte swizzle (layout_type 2), descriptor version 1
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fffu);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// K-major operand tile [128 rows][64 k] (+k-chunks 16 KB apart): 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t tile, int kk) {
  return make_desc(tile + (uint32_t)(kk >> 2) * CHUNK_BYTES + (uint32_t)(kk & 3) * 32u, 16u, 1024u);
}
// MN-major operand: the same tile read as [K = rows][MN = columns]; 64-column chunks
// are 16 KB apart (LBO), 8-row groups 1024 B apart (SBO); one MMA consumes 16 rows
__device__ __forceinline__ uint64_t desc_mnmajor(uint32_t tile, int kk) {
  return make_desc(tile + (uint32_t)kk * 2048u, (uint32_t)CHUNK_BYTES, 1024u);
}

__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lg2_approx(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
// two floats -> packed bf16x2 (lo = a, hi = b), round to nearest even
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}

// two floats -> packed f16x2 (lo = a, hi = b), round to nearest even (subnormal results are kept)
__device__ __forceinline__ uint32_t pack_f16(float a, float b) {
  uint32_t r;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}

// bf16 matrix [rows][cols] row-major, box = 128 rows x 64 columns, 128-byte swizzle
bool encode_map(void* out, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows = 128) {
  EncodeTiledFn fn = get_encode();
  if (!fn) return false;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {cols * 2};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(reinterpret_cast<CUtensorMap*>(out), CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims,
                  strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

