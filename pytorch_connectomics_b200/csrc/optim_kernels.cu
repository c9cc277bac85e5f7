// pcb200 — optimizer-side fusion over the flat arenas: gradient-norm clip + AdamW with the reference's
// per-parameter groups + EMA of the weights in ONE pass over (param, grad, exp_avg, exp_avg_sq, ema).
// Reference: connectomics/training/optimization/build.py:88-113 (param groups: norm layers / biases get their own weight
// decay and lr), Lightning gradient_clip_val (torch.nn.utils.clip_grad_norm_: coef = min(1, max_norm/(norm+1e-6))),
// connectomics/training/lightning/callbacks.py:869-907 (ema = ema*decay + param*(1-decay) after every optimizer step).
// Pure HBM streaming: 5 x 4 B read + 4 x 4 B written per parameter, 128-bit accesses, grid = a multiple of 148 CTAs.
#include "../../include/pcb200.h"
#include "pcb_common.cuh"

namespace pcb {

__global__ void __launch_bounds__(256) grad_sumsq_kernel(const float* __restrict__ g, int64_t n, double* __restrict__ out) {
  double acc = 0.0;
  const int64_t n4 = n >> 2;
  const float4* g4 = reinterpret_cast<const float4*>(g);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 v = __ldg(g4 + i);
    acc += (double)v.x * v.x + (double)v.y * v.y + (double)v.z * v.z + (double)v.w * v.w;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0)
    for (int64_t i = n4 << 2; i < n; ++i) acc += (double)g[i] * g[i];
  for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  __shared__ double s[8];
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += s[w];
    atomicAdd(out, t);
  }
}

// seg_active[i] = 1 iff segment i's gradient has a non-zero element.  A parameter that received no gradient at all (unused
// deep-supervision heads, MedNeXt's dummy_tensor: its slice of the arena stays exactly zero) is SKIPPED — torch.optim
// skips parameters whose .grad is None, which is what DDP(find_unused_parameters=True) leaves for them: no weight
// decay, no moment decay, no EMA drift relative to the reference.
__global__ void __launch_bounds__(256) seg_active_kernel(const float* __restrict__ g, const int64_t* __restrict__ seg_end,
                                                         int nseg, int* __restrict__ active) {
  const int seg = blockIdx.x;
  if (seg >= nseg) return;
  const int64_t lo = seg == 0 ? 0 : seg_end[seg - 1], hi = seg_end[seg];
  int any = 0;
  for (int64_t i = lo + threadIdx.x; i < hi && !any; i += blockDim.x) any |= (g[i] != 0.f);
  any = __syncthreads_or(any);
  if (threadIdx.x == 0) active[seg] = any;
}

struct AdamArgs {
  float* p; const float* g; float* m; float* v; float* ema;
  const int* active;
  int64_t n;
  const int64_t* seg_end; const float* seg_lr; const float* seg_wd; int nseg;
  float beta1, beta2, eps, max_norm, grad_scale, ema_decay;
  float* step;                  // device scalar: number of optimizer steps taken so far (incremented by the kernel's CTA 0 peer)
  const double* sumsq;          // device scalar: sum of squares of the (unscaled) gradient arena, or nullptr (no clip)
};

__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, float* ema, float lr, float wd, float b1,
                                         float b2, float eps, float bc1, float bc2_sqrt, float ema_decay) {
  // torch.optim.AdamW single-tensor update, op for op (no FMA contraction across torch's op boundaries)
  p = __fmul_rn(p, 1.0f - lr * wd);
  m = __fadd_rn(m, __fmul_rn(__fsub_rn(g, m), 1.0f - b1));                 // exp_avg.lerp_(grad, 1 - beta1)
  v = __fadd_rn(__fmul_rn(v, b2), __fmul_rn(__fmul_rn(g, g), 1.0f - b2));  // mul_(beta2).addcmul_(g, g, 1 - beta2)
  const float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(v), bc2_sqrt), eps);
  p = __fadd_rn(p, __fmul_rn(-(lr / bc1), __fdiv_rn(m, denom)));           // addcdiv_(exp_avg, denom, value=-step_size)
  if (ema != nullptr) *ema = __fadd_rn(__fmul_rn(*ema, ema_decay), __fmul_rn(p, 1.0f - ema_decay));
}

__global__ void __launch_bounds__(256) adamw_kernel(AdamArgs a) {
  extern __shared__ int64_t s_end[];                      // [nseg] ends, then lr / wd floats, then active flags
  float* s_lr = reinterpret_cast<float*>(s_end + a.nseg);
  float* s_wd = s_lr + a.nseg;
  int* s_act = reinterpret_cast<int*>(s_wd + a.nseg);
  for (int i = threadIdx.x; i < a.nseg; i += blockDim.x) {
    s_end[i] = a.seg_end[i]; s_lr[i] = a.seg_lr[i]; s_wd[i] = a.seg_wd[i]; s_act[i] = a.active ? a.active[i] : 1;
  }
  __syncthreads();
  const float t = a.step[0] + 1.0f;                       // this step's index (the host-side wrapper bumps a.step afterwards)
  const float bc1 = 1.0f - (float)pow((double)a.beta1, (double)t);
  const float bc2_sqrt = (float)sqrt(1.0 - pow((double)a.beta2, (double)t));
  float gs = a.grad_scale;
  if (a.sumsq != nullptr && a.max_norm > 0.f) {
    const float total = (float)(sqrt(a.sumsq[0]) * (double)a.grad_scale);   // norm of the scaled (averaged) gradient
    const float coef = fminf(a.max_norm / (total + 1e-6f), 1.0f);
    gs *= coef;
  }
  const int64_t n4 = (a.n + 3) >> 2;
  for (int64_t i4 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i4 < n4; i4 += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i0 = i4 << 2;
    int lo = 0, hi = a.nseg - 1;                           // first segment whose end > i0
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (s_end[mid] > i0) hi = mid; else lo = mid + 1; }
    int seg = lo;
    if (i0 + 4 <= a.n && s_end[seg] >= i0 + 4) {           // fast path: 4 elements of one segment, 128-bit accesses
      if (!s_act[seg]) continue;
      float4 p = reinterpret_cast<float4*>(a.p)[i4], m = reinterpret_cast<float4*>(a.m)[i4], v = reinterpret_cast<float4*>(a.v)[i4];
      const float4 g = __ldg(reinterpret_cast<const float4*>(a.g) + i4);
      float4 e = a.ema ? reinterpret_cast<float4*>(a.ema)[i4] : make_float4(0, 0, 0, 0);
      const float lr = s_lr[seg], wd = s_wd[seg];
      adam_one(p.x, __fmul_rn(g.x, gs), m.x, v.x, a.ema ? &e.x : nullptr, lr, wd, a.beta1, a.beta2, a.eps, bc1, bc2_sqrt, a.ema_decay);
      adam_one(p.y, __fmul_rn(g.y, gs), m.y, v.y, a.ema ? &e.y : nullptr, lr, wd, a.beta1, a.beta2, a.eps, bc1, bc2_sqrt, a.ema_decay);
      adam_one(p.z, __fmul_rn(g.z, gs), m.z, v.z, a.ema ? &e.z : nullptr, lr, wd, a.beta1, a.beta2, a.eps, bc1, bc2_sqrt, a.ema_decay);
      adam_one(p.w, __fmul_rn(g.w, gs), m.w, v.w, a.ema ? &e.w : nullptr, lr, wd, a.beta1, a.beta2, a.eps, bc1, bc2_sqrt, a.ema_decay);
      reinterpret_cast<float4*>(a.p)[i4] = p; reinterpret_cast<float4*>(a.m)[i4] = m; reinterpret_cast<float4*>(a.v)[i4] = v;
      if (a.ema) reinterpret_cast<float4*>(a.ema)[i4] = e;
    } else {
      for (int64_t i = i0; i < i0 + 4 && i < a.n; ++i) {
        while (s_end[seg] <= i) ++seg;
        if (!s_act[seg]) continue;
        adam_one(a.p[i], __fmul_rn(a.g[i], gs), a.m[i], a.v[i], a.ema ? a.ema + i : nullptr, s_lr[seg], s_wd[seg], a.beta1,
                 a.beta2, a.eps, bc1, bc2_sqrt, a.ema_decay);
      }
    }
  }
}

__global__ void bump_step_kernel(float* step) {
  if (threadIdx.x == 0 && blockIdx.x == 0) step[0] += 1.0f;
}

}  // namespace pcb

using namespace pcb;

extern "C" int pcb_grad_sumsq(const float* grad, int64_t n, double* out, void* stream) {
  PCB_CHECK_ARG(grad && out && n > 0, "pcb_grad_sumsq: bad argument");
  PCB_CHECK_ARG((reinterpret_cast<uintptr_t>(grad) & 15) == 0, "pcb_grad_sumsq: the arena must be 16-byte aligned");
  int64_t blocks = (n / 4 + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  grad_sumsq_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(grad, n, out);
  PCB_CHECK_LAUNCH("pcb_grad_sumsq");
  return PCB_OK;
}

extern "C" int pcb_adamw_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, float* ema, int64_t n,
                              const int64_t* seg_end, const float* seg_lr, const float* seg_wd, int nseg, float beta1,
                              float beta2, float eps, float* step, const double* grad_sumsq, float max_norm,
                              float grad_scale, float ema_decay, int32_t* seg_active, void* stream) {
  PCB_CHECK_ARG(param && grad && exp_avg && exp_avg_sq && seg_end && seg_lr && seg_wd && step && n > 0 && nseg > 0,
                "pcb_adamw_step: null argument");
  PCB_CHECK_ARG(nseg <= 4096, "pcb_adamw_step: at most 4096 parameter segments (got %d)", nseg);
  const uintptr_t al = reinterpret_cast<uintptr_t>(param) | reinterpret_cast<uintptr_t>(grad) |
                       reinterpret_cast<uintptr_t>(exp_avg) | reinterpret_cast<uintptr_t>(exp_avg_sq) |
                       reinterpret_cast<uintptr_t>(ema);
  PCB_CHECK_ARG((al & 15) == 0, "pcb_adamw_step: the arenas must be 16-byte aligned");
  AdamArgs a{param, grad, exp_avg, exp_avg_sq, ema, seg_active, n, seg_end, seg_lr, seg_wd, nseg, beta1, beta2, eps, max_norm,
             grad_scale, ema_decay, step, grad_sumsq};
  int64_t blocks = ((n + 3) / 4 + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  cudaStream_t st = (cudaStream_t)stream;
  if (seg_active != nullptr) {
    seg_active_kernel<<<(unsigned)nseg, 256, 0, st>>>(grad, seg_end, nseg, seg_active);
    PCB_CHECK_LAUNCH("pcb_adamw_step(active)");
  }
  adamw_kernel<<<(unsigned)blocks, 256, (size_t)nseg * (sizeof(int64_t) + 2 * sizeof(float) + sizeof(int)), st>>>(a);
  PCB_CHECK_LAUNCH("pcb_adamw_step");
  bump_step_kernel<<<1, 32, 0, st>>>(step);
  PCB_CHECK_LAUNCH("pcb_adamw_step(step)");
  return PCB_OK;
}
