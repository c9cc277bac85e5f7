// pcb200 — shared device helpers: sm_100a PTX wrappers (mbarrier, tcgen05, TMEM), UMMA
// descriptors for the no-swizzle canonical layouts, bf16 pack/unpack, error plumbing.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

namespace pcb {

// Kernel function attributes (dynamic shared-memory opt-in, carveout) are PER DEVICE: a process-wide `static bool`
// would configure only the first GPU a process launches on, and >48 KB launches on a second GPU would fail with
// invalid-argument.  DevFlag is a drop-in for such a flag with one atomic slot per device ordinal (configuring twice
// from two threads is idempotent and harmless).
struct DevFlag {
  std::atomic<unsigned char> f[64];
  DevFlag() { for (auto& x : f) x.store(0); }
  static int dev() { int d = 0; if (cudaGetDevice(&d) != cudaSuccess) { cudaGetLastError(); d = 0; } return (d >= 0 && d < 64) ? d : 0; }
  operator bool() const { return f[dev()].load(std::memory_order_acquire) != 0; }
  DevFlag& operator=(bool v) { f[dev()].store(v ? 1 : 0, std::memory_order_release); return *this; }
};

// ----------------------------------------------------------------------------- error plumbing
void set_error(const char* fmt, ...);
void count_launch();
#define PCB_CHECK_ARG(cond, ...)                                   \
  do {                                                             \
    if (!(cond)) {                                                 \
      pcb::set_error(__VA_ARGS__);                                 \
      return PCB_ERR_INVALID;                                      \
    }                                                              \
  } while (0)
#define PCB_CHECK_LAUNCH(what)                                     \
  do {                                                             \
    pcb::count_launch();                                           \
    cudaError_t e__ = cudaGetLastError();                          \
    if (e__ != cudaSuccess) {                                      \
      pcb::set_error("%s: %s", what, cudaGetErrorString(e__));     \
      return PCB_ERR_CUDA;                                         \
    }                                                              \
  } while (0)

// ----------------------------------------------------------------------------- small utils
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ float bf16_lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t v) { return __uint_as_float(v & 0xffff0000u); }
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float round_bf16(float v) {
  return __bfloat162float(__float2bfloat16_rn(v));
}
__device__ __forceinline__ void unpack8(const uint4& v, float* f) {
  f[0] = bf16_lo(v.x); f[1] = bf16_hi(v.x); f[2] = bf16_lo(v.y); f[3] = bf16_hi(v.y);
  f[4] = bf16_lo(v.z); f[5] = bf16_hi(v.z); f[6] = bf16_lo(v.w); f[7] = bf16_hi(v.w);
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  uint4 v;
  v.x = pack_bf16(f[0], f[1]); v.y = pack_bf16(f[2], f[3]);
  v.z = pack_bf16(f[4], f[5]); v.w = pack_bf16(f[6], f[7]);
  return v;
}
__device__ __forceinline__ uint4 ldg_nc(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
// ---- packed fp32x2 arithmetic (Blackwell FFMA2/FMUL2/FADD2: two fp32 lanes per instruction)
__device__ __forceinline__ uint64_t pk2(float a, float b) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void upk2(uint64_t v, float& a, float& b) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ float tanh_mufu(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// GELU(x) = x*Phi(x) with Phi(x) ~= 0.5*(1 + tanh(x*(c0 + c1 s + c2 s^2))), s = min(x^2, 64).
// Minimax fit of the 3-term odd polynomial against the exact erf form on [-8, 8]:
// |GELU err| < 3.0e-5, |GELU' err| < 9.5e-5 (tools/fit_gelu.py); the MUFU tanh.approx.f32 adds a
// relative 2^-11 on tanh.  Both are far below the bf16 rounding (2^-9) applied to the result.
#define PCB_GELU_C0 0.797482750f
#define PCB_GELU_C1 0.0369853532f
#define PCB_GELU_C2 (-3.46672670e-4f)
__device__ __forceinline__ void gelu_fast2p(uint64_t x, float& y0, float& y1);
__device__ __forceinline__ void gelu_fast2(float x0, float x1, float& y0, float& y1) {
  gelu_fast2p(pk2(x0, x1), y0, y1);
}
__device__ __forceinline__ void gelu_fast2p(uint64_t x, float& y0, float& y1) {
  float s0, s1;
  upk2(mul2(x, x), s0, s1);
  const uint64_t s = pk2(fminf(s0, 64.f), fminf(s1, 64.f));
  uint64_t p = fma2(pk2(PCB_GELU_C2, PCB_GELU_C2), s, pk2(PCB_GELU_C1, PCB_GELU_C1));
  p = fma2(p, s, pk2(PCB_GELU_C0, PCB_GELU_C0));
  float u0, u1;
  upk2(mul2(x, p), u0, u1);
  const uint64_t t = pk2(tanh_mufu(u0), tanh_mufu(u1));
  const uint64_t h = mul2(x, pk2(0.5f, 0.5f));
  upk2(fma2(h, t, h), y0, y1);
}
// value and derivative (backward): GELU' = 0.5(1+t) + 0.5 x (1-t^2) (c0 + 3 c1 s + 5 c2 s^2)
__device__ __forceinline__ void gelu_fast_vg(float x, float& val, float& grad) {
  const float s = fminf(x * x, 64.f);
  const float p = fmaf(fmaf(PCB_GELU_C2, s, PCB_GELU_C1), s, PCB_GELU_C0);
  const float dp = fmaf(fmaf(5.f * PCB_GELU_C2, s, 3.f * PCB_GELU_C1), s, PCB_GELU_C0);
  const float t = tanh_mufu(x * p);
  const float h = 0.5f * x;
  val = fmaf(h, t, h);
  const float cdf = fmaf(0.5f, t, 0.5f);
  grad = fmaf(h * fmaf(-t, t, 1.f), (x * x < 64.f) ? dp : 0.f, cdf);
}

// packed pair version: (x0,x1) -> values and derivatives, ~10 issue slots per element
__device__ __forceinline__ void gelu_fast_vg2(uint64_t x, uint64_t& val, uint64_t& grad) {
  float s0, s1;
  upk2(mul2(x, x), s0, s1);
  const uint64_t s = pk2(fminf(s0, 64.f), fminf(s1, 64.f));
  const uint64_t p = fma2(fma2(pk2(PCB_GELU_C2, PCB_GELU_C2), s, pk2(PCB_GELU_C1, PCB_GELU_C1)), s, pk2(PCB_GELU_C0, PCB_GELU_C0));
  const uint64_t dp = fma2(fma2(pk2(5.f * PCB_GELU_C2, 5.f * PCB_GELU_C2), s, pk2(3.f * PCB_GELU_C1, 3.f * PCB_GELU_C1)), s,
                           pk2(PCB_GELU_C0, PCB_GELU_C0));
  float u0, u1;
  upk2(mul2(x, p), u0, u1);
  const float t0 = tanh_mufu(u0), t1 = tanh_mufu(u1);
  const uint64_t t = pk2(t0, t1), nt = pk2(-t0, -t1);
  const uint64_t h = mul2(x, pk2(0.5f, 0.5f));
  val = fma2(h, t, h);
  const uint64_t cdf = fma2(pk2(0.5f, 0.5f), t, pk2(0.5f, 0.5f));
  const uint64_t sech2 = fma2(nt, t, pk2(1.f, 1.f));          // 1 - t^2 (exactly 0 once tanh saturates)
  grad = fma2(mul2(h, sech2), dp, cdf);
}

// exact (erf) GELU and its derivative, fp32
__device__ __forceinline__ float gelu_f(float x) {
  return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f));
}
__device__ __forceinline__ float gelu_grad_f(float x) {
  const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752f));
  const float pdf = 0.3989422804014327f * __expf(-0.5f * x * x);
  return cdf + x * pdf;
}

// Staging copy with memory-level parallelism: issue up to B independent 128-bit global loads per
// thread, THEN run the (convert +) shared-memory stores — a plain load->store loop serialises on
// DRAM latency because the compiler cannot hoist loads across the aliasing shared stores.
template <int B, typename LoadF, typename StoreF>
__device__ __forceinline__ void staged_copy(int total, int tid, int nthreads, LoadF load, StoreF store) {
  for (int q0 = tid; q0 < total; q0 += nthreads * B) {
    uint4 v[B];
#pragma unroll
    for (int b = 0; b < B; ++b) {
      const int q = q0 + b * nthreads;
      if (q < total) v[b] = load(q);
    }
#pragma unroll
    for (int b = 0; b < B; ++b) {
      const int q = q0 + b * nthreads;
      if (q < total) store(q, v[b]);
    }
  }
}

// 16 per-lane values -> column totals over the 32 lanes with 16 shuffles (recursive halving).
// Afterwards lane l holds the total of column  8*b4 + 4*b3 + 2*b2 + b1  (bits of l) in v[0].
__device__ __forceinline__ void warp_colsum16(float* v, int lane) {
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const float send = (lane & 16) ? v[k] : v[k + 8], keep = (lane & 16) ? v[k + 8] : v[k];
    v[k] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float send = (lane & 8) ? v[k] : v[k + 4], keep = (lane & 8) ? v[k + 4] : v[k];
    v[k] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const float send = (lane & 4) ? v[k] : v[k + 2], keep = (lane & 4) ? v[k + 2] : v[k];
    v[k] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
  {
    const float send = (lane & 2) ? v[0] : v[1], keep = (lane & 2) ? v[1] : v[0];
    v[0] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  }
  v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
}
__device__ __forceinline__ int colsum16_col(int lane) {
  return ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
}

// ----------------------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(addr), "r"(parity) : "memory");
  } while (!done);
}
// non-blocking probe of a phase parity (polling issuers that serve several independent pipelines)
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return done != 0;
}
// ----------------------------------------------------------------------------- TMA (cp.async.bulk.tensor)
// One thread arms the mbarrier with the byte count of the box and issues the bulk tensor load; the copy
// engine writes the box densely ([c4][c3][c2][c1][c0] order, innermost = channels) into shared memory,
// zero-filling every element whose coordinate falls outside the tensor (that is the stencil halo padding),
// and completes the transaction on the mbarrier.  Waiters use mbar_wait on the phase parity.
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_5d(void* smem_dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, int c4,
                                            uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4),
        "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

// Host: tensor map of a channels-last bf16 activation [N, D, H, W, C] with a [1, bd, bh, bw, 32-channel] box.
// cuTensorMapEncodeTiled is resolved through the runtime (no link-time dependency on libcuda).
inline bool make_brick_tensor_map(CUtensorMap* map, const void* base, int64_t N, int64_t D, int64_t H, int64_t W, int64_t C,
                                  int bd, int bh, int bw) {
  static PFN_cuTensorMapEncodeTiled_v12000 encode = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
    else
      cudaGetLastError();
  }
  if (encode == nullptr) return false;
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || C % 8 != 0) return false;
  const cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)N};
  const cuuint64_t strides[4] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2, (cuuint64_t)D * H * W * C * 2};
  const cuuint32_t box[5] = {32, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bd, 1};
  const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  if (bw > 256 || bh > 256 || bd > 256) return false;
  return encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), dims, strides, box, estr,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
// generic-proxy smem writes -> visible to the async proxy (UMMA operand reads)
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ----------------------------------------------------------------------------- tcgen05 / TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
               ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
               ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]; issued by ONE thread
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// UMMA shared-memory descriptor, SWIZZLE_NONE ("interleave") canonical layouts (cute
// mma_traits_sm100.hpp make_umma_desc): 8x16B core matrices.
//   K-major  operand [rows x K]: elem(r,k) at (r/8)*SBO + (k/8)*LBO + (r%8)*16 + (k%8)*2
//   MN-major operand [mn x K]  : elem(m,k) at (m/8)*SBO + (k/8)*LBO + (k%8)*16 + (m%8)*2
// One tcgen05.mma (bf16) consumes K=16, i.e. two core matrices LBO apart.
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFFu);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version 1 (Blackwell)
  return d;                // base_offset 0, lbo_mode 0, layout_type SWIZZLE_NONE (0)
}
// instruction descriptor, kind::f16, bf16 x bf16 -> fp32, M=128
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) |
         ((uint32_t)b_mn_major << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__host__ __device__ constexpr uint32_t tmem_cols_pow2(uint32_t n) {
  return n <= 32 ? 32 : n <= 64 ? 64 : n <= 128 ? 128 : n <= 256 ? 256 : 512;
}

// deep-level (C >= 128) MLP data gradient as two launches of the column-split GEMM (csrc/deep_mlp.cu);
// returns 1 when it ran, 0 when the shape is not served (caller uses mlp_bwd_kernel), <0 on error.
int mlp_bwd_deep(const void* y, const double* stats, const float* gamma, const float* beta, const void* w2, const float* b2,
                 const void* w3t, const void* w2t, const void* dout, void* hact, void* dh, void* dyhat, double* gstats,
                 int64_t N, const int64_t y_size[3], int64_t C, int64_t H, int64_t Co, int mode, void* stream);

}  // namespace pcb
