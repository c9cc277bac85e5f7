// pcb200 — test-time-augmentation views and the streaming ensemble.
//
// Replaces, for fully valid channels (no affinity channel moves), the per-view tensor ops of
// connectomics/inference/tta.py:691-771 (_run_ensemble: torch.flip + torch.rot90 on the input, the network,
// tta_affinity.py:364-369 invert_view: rot90(-k) then flip), tta.py:312-402 (apply_preprocessing: per-channel
// sigmoid / scale_sigmoid / tanh, channel selection, cast to the output dtype) and
// tta_ensemble.py:94-110 (_add_full_channels: first view copies, "mean" is the running average
// cur += (inc - cur) / (n + 1), "min" / "max" elementwise).  Two kernels:
//   tta_view_kernel  out = rot90(flip(x, axes), k, plane)            (gather; the view handed to the network)
//   tta_fold_kernel  acc <- fold(acc, act(inverse_view(pred))[selected channels])   — the inverse view is only an
//                    index map, so un-rotating, un-flipping, the activation, the channel select, the dtype cast and the
//                    ensemble update are ONE pass over the accumulator instead of six tensor ops per view.
// Arithmetic follows torch's elementwise kernels (compute in fp32, round to the tensor dtype after every op), so the
// fp32 path is bit-identical to the reference expression and the half paths round exactly where torch rounds.
#include "../../include/pcb200.h"
#include <stdlib.h>
#include <string.h>

#include "pcb_common.cuh"

namespace pcb {

struct TtaGeom {
  int s[3];        // spatial size of the tensor being WRITTEN (output of the kernel's index space)
  int in[3];       // spatial size of the tensor being READ
  int flip[3];     // flip flags per spatial axis
  int ra, rb, k;   // rotation plane (spatial axes, ra < 0: none) and number of quarter turns (0..3)
};

// source coordinate of torch.rot90(x, k, (ra, rb))[o] in x (sizes `in`): k=1: x.flip(rb).transpose; k=2: both flips;
// k=3: x.flip(ra).transpose
__device__ __forceinline__ void rot_src(const TtaGeom& g, int k, const int* in_size, int* c) {
  if (g.ra < 0 || k == 0) return;
  const int oa = c[g.ra], ob = c[g.rb];
  if (k == 1) { c[g.ra] = ob; c[g.rb] = in_size[g.rb] - 1 - oa; }
  else if (k == 2) { c[g.ra] = in_size[g.ra] - 1 - oa; c[g.rb] = in_size[g.rb] - 1 - ob; }
  else { c[g.ra] = in_size[g.ra] - 1 - ob; c[g.rb] = oa; }
}

template <typename T> struct TtaNum;
template <> struct TtaNum<float> {
  static __device__ __forceinline__ float ld(const float* p) { return *p; }
  static __device__ __forceinline__ float rnd(float v) { return v; }
  static __device__ __forceinline__ void st(float* p, float v) { *p = v; }
};
template <> struct TtaNum<__half> {
  static __device__ __forceinline__ float ld(const __half* p) { return __half2float(*p); }
  static __device__ __forceinline__ float rnd(float v) { return __half2float(__float2half_rn(v)); }
  static __device__ __forceinline__ void st(__half* p, float v) { *p = __float2half_rn(v); }
};
template <> struct TtaNum<__nv_bfloat16> {
  static __device__ __forceinline__ float ld(const __nv_bfloat16* p) { return __bfloat162float(*p); }
  static __device__ __forceinline__ float rnd(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }
  static __device__ __forceinline__ void st(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
};

// out[plane, o] = x[plane, flip(rot_src(o))]   — planes = N*C
template <typename T>
__global__ void __launch_bounds__(256) tta_view_kernel(const T* __restrict__ x, T* __restrict__ out, TtaGeom g, int64_t planes) {
  const int64_t vol = (int64_t)g.s[0] * g.s[1] * g.s[2];
  const int64_t total = planes * vol;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = i / vol;
    int64_t r = i - p * vol;
    int c[3];
    c[2] = (int)(r % g.s[2]); r /= g.s[2];
    c[1] = (int)(r % g.s[1]);
    c[0] = (int)(r / g.s[1]);
    rot_src(g, g.k, g.in, c);                       // coordinate in flip(x): same sizes as x
#pragma unroll
    for (int a = 0; a < 3; ++a) if (g.flip[a]) c[a] = g.in[a] - 1 - c[a];
    out[i] = x[p * vol + ((int64_t)c[0] * g.in[1] + c[1]) * g.in[2] + c[2]];
  }
}

constexpr int TTA_MAXC = 64;
struct TtaFold {
  TtaGeom g;               // g.s = canonical (accumulator) size, g.in = prediction (view frame) size, g.k = inverse turns
  int N, Cacc, Cpred;
  int n_prev;              // views folded so far
  int src[TTA_MAXC];       // prediction channel feeding accumulator channel c (channel selection o affinity channel move)
  int mode[TTA_MAXC];      // 0 mean, 1 min, 2 max, 3 skip (channel is aggregated through the partial path)
  int act[TTA_MAXC];       // 0 none, 1 sigmoid, 2 scale_sigmoid, 3 tanh, 4 softmax over the members sm_*[sm_off .. sm_off+sm_len)
  float scale[TTA_MAXC];
  // ---- affinity-aware inversion (tta_affinity.py:376-391): canonical[c][p] = spatial[src][p - shift], zero / invalid outside
  int shift[TTA_MAXC][3];
  int any_shift;
  // ---- softmax groups (tta.py:368-370): members are (prediction channel, shift) pairs of the canonical tensor
  int sm_off[TTA_MAXC], sm_len[TTA_MAXC];
  int sm_src[TTA_MAXC];
  int sm_shift[TTA_MAXC][3];
  // ---- validity-aware aggregation of partial channels (tta_ensemble.py:121-162)
  int part[TTA_MAXC];      // index into stats / counts, -1 = fully valid channel
  int pmode[TTA_MAXC];     // ensemble mode of a partial channel (mode[] holds 3 for them)
  int Cpart;
  int mean_as_sum;         // distributed view sharding: "mean" accumulates the plain sum (tta_ensemble.py:92-93)
};

// canonical coordinate c (size g.s) -> offset of the source voxel in the view-frame prediction, after the roll shift `sh`
// of an affinity channel; false when the shifted source falls outside (the wrapped face: zero value, invalid)
__device__ __forceinline__ bool tta_src_offset(const TtaGeom& g, const int* c, const int* sh, int64_t& off) {
  int q[3] = {c[0] - sh[0], c[1] - sh[1], c[2] - sh[2]};
#pragma unroll
  for (int a = 0; a < 3; ++a) if (q[a] < 0 || q[a] >= g.s[a]) return false;
  // canonical = flip(rot90(pred, -k)): undo the flip (sizes of the canonical frame), then the rotation
#pragma unroll
  for (int a = 0; a < 3; ++a) if (g.flip[a]) q[a] = g.s[a] - 1 - q[a];
  rot_src(g, g.k, g.in, q);
  off = ((int64_t)q[0] * g.in[1] + q[1]) * g.in[2] + q[2];
  return true;
}

// TP: prediction dtype, TA: accumulator (output) dtype, TC: count dtype of the partial channels
template <typename TP, typename TA, typename TC>
__global__ void __launch_bounds__(256) tta_fold_kernel(const TP* __restrict__ pred, TA* __restrict__ acc, float* __restrict__ stats,
                                                       TC* __restrict__ counts, const uint8_t* __restrict__ vmask, TtaFold f) {
  const TtaGeom& g = f.g;
  const int64_t vol = (int64_t)g.s[0] * g.s[1] * g.s[2];
  const int64_t pvol = (int64_t)g.in[0] * g.in[1] * g.in[2];
  const int64_t total = (int64_t)f.N * f.Cacc * vol;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t pc = i / vol;
    int64_t r = i - pc * vol;
    const int64_t vox = r;
    const int ch = (int)(pc % f.Cacc);
    const int64_t n = pc / f.Cacc;
    const int pj = f.part[ch];
    if (pj < 0 && f.mode[ch] == 3) continue;
    int c[3];
    c[2] = (int)(r % g.s[2]); r /= g.s[2];
    c[1] = (int)(r % g.s[1]);
    c[0] = (int)(r / g.s[1]);
    int64_t off;
    bool valid = tta_src_offset(g, c, f.shift[ch], off);
    float v = valid ? TtaNum<TP>::ld(pred + (n * f.Cpred + f.src[ch]) * pvol + off) : 0.f;
    // activation in the prediction dtype (torch in-place ops on the network output), then the cast to the output dtype
    const int act = f.act[ch];
    if (act == 1) v = TtaNum<TP>::rnd(1.0f / (1.0f + expf(-v)));
    else if (act == 2) { v = TtaNum<TP>::rnd(v * f.scale[ch]); v = TtaNum<TP>::rnd(1.0f / (1.0f + expf(-v))); }
    else if (act == 3) v = TtaNum<TP>::rnd(tanhf(v));
    else if (act == 4) {
      // softmax over the group's canonical channels at this voxel (fp32 internally, one rounding: torch.softmax)
      float mx = -INFINITY;
      for (int m = f.sm_off[ch]; m < f.sm_off[ch] + f.sm_len[ch]; ++m) {
        int64_t o2;
        const float u = tta_src_offset(g, c, f.sm_shift[m], o2) ? TtaNum<TP>::ld(pred + (n * f.Cpred + f.sm_src[m]) * pvol + o2) : 0.f;
        mx = fmaxf(mx, u);
      }
      float den = 0.f;
      for (int m = f.sm_off[ch]; m < f.sm_off[ch] + f.sm_len[ch]; ++m) {
        int64_t o2;
        const float u = tta_src_offset(g, c, f.sm_shift[m], o2) ? TtaNum<TP>::ld(pred + (n * f.Cpred + f.sm_src[m]) * pvol + o2) : 0.f;
        den += expf(u - mx);
      }
      v = TtaNum<TP>::rnd(expf(v - mx) / den);
    }
    v = TtaNum<TA>::rnd(v);
    if (pj >= 0) {
      // partial channel: statistics in fp32, per-voxel contribution counts; invalid voxels contribute nothing
      if (vmask != nullptr) valid = valid && vmask[(int64_t)pj * vol + vox] != 0;
      if (!valid) continue;
      const int64_t si = (n * f.Cpart + pj) * vol + vox;
      const float cur = stats[si];
      const int pm = f.pmode[ch];
      stats[si] = pm == 0 ? __fadd_rn(cur, v) : (pm == 1 ? fminf(cur, v) : fmaxf(cur, v));
      counts[si] = (TC)(counts[si] + 1);
      continue;
    }
    if (f.n_prev > 0) {
      const float cur = TtaNum<TA>::ld(acc + i);
      const int mode = f.mode[ch];
      if (mode == 0) {
        if (f.mean_as_sum) v = TtaNum<TA>::rnd(__fadd_rn(cur, v));
        else {
          const float delta = TtaNum<TA>::rnd(v - cur);
          const float q = TtaNum<TA>::rnd(__fdiv_rn(delta, (float)(f.n_prev + 1)));
          v = __fadd_rn(cur, q);
        }
      } else if (mode == 1) v = fminf(cur, v);
      else v = fmaxf(cur, v);
    }
    TtaNum<TA>::st(acc + i, v);
  }
}

// invert_view (tta_affinity.py:350-393) as one gather: out[n, c] = canonical channel c of the view prediction
template <typename T>
__global__ void __launch_bounds__(256) tta_unview_kernel(const T* __restrict__ pred, T* __restrict__ out, TtaFold f) {
  const TtaGeom& g = f.g;
  const int64_t vol = (int64_t)g.s[0] * g.s[1] * g.s[2];
  const int64_t pvol = (int64_t)g.in[0] * g.in[1] * g.in[2];
  const int64_t total = (int64_t)f.N * f.Cacc * vol;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t pc = i / vol;
    int64_t r = i - pc * vol;
    const int ch = (int)(pc % f.Cacc);
    const int64_t n = pc / f.Cacc;
    int c[3];
    c[2] = (int)(r % g.s[2]); r /= g.s[2];
    c[1] = (int)(r % g.s[1]);
    c[0] = (int)(r / g.s[1]);
    int64_t off;
    T v = T(0.f);
    if (tta_src_offset(g, c, f.shift[ch], off)) v = pred[(n * f.Cpred + f.src[ch]) * pvol + off];
    out[i] = v;
  }
}

// tta_ensemble.py:187-211 finalize: result[:, channel] = statistics (/ counts for "mean") cast to the output dtype; the
// smallest flat index with zero coverage is reported through *first_zero (the host raises like the reference)
template <typename TA, typename TC>
__global__ void __launch_bounds__(256) tta_finalize_kernel(const float* __restrict__ stats, const TC* __restrict__ counts,
                                                           TA* __restrict__ acc, TtaFold f, int64_t vol,
                                                           unsigned long long* __restrict__ first_zero) {
  const int64_t total = (int64_t)f.N * f.Cpart * vol;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t pc = i / vol, vox = i - pc * vol;
    const int pj = (int)(pc % f.Cpart);
    const int64_t n = pc / f.Cpart;
    const int ch = f.src[pj];                 // accumulator channel of partial channel pj
    const TC cnt = counts[i];
    if (cnt == 0) { atomicMin(first_zero, (unsigned long long)i); continue; }
    float v = stats[i];
    if (f.pmode[pj] == 0) v = __fdiv_rn(v, (float)cnt);
    TtaNum<TA>::st(acc + (n * f.Cacc + ch) * vol + vox, v);
  }
}

static bool geom_ok(const int64_t s[3]) { return s && s[0] > 0 && s[1] > 0 && s[2] > 0 && s[0] * s[1] * s[2] < (1ll << 40); }

}  // namespace pcb

using namespace pcb;

static int tta_grid(int64_t total) {
  int64_t b = (total + 255) / 256;
  if (b > 148 * 16) b = 148 * 16;
  return (int)(b < 1 ? 1 : b);
}

extern "C" int pcb_tta_view(const void* x, void* out, int dtype, int64_t planes, const int64_t in_size[3], int flip_mask,
                            int rot_a, int rot_b, int k, void* stream) {
  PCB_CHECK_ARG(x && out && geom_ok(in_size) && planes > 0, "pcb_tta_view: bad argument");
  PCB_CHECK_ARG(k >= 0 && k <= 3, "pcb_tta_view: k must be in 0..3 (got %d)", k);
  PCB_CHECK_ARG((rot_a < 0 && rot_b < 0) || (rot_a >= 0 && rot_a < 3 && rot_b >= 0 && rot_b < 3 && rot_a != rot_b),
                "pcb_tta_view: bad rotation plane (%d, %d)", rot_a, rot_b);
  TtaGeom g;
  for (int a = 0; a < 3; ++a) { g.in[a] = (int)in_size[a]; g.s[a] = (int)in_size[a]; g.flip[a] = (flip_mask >> a) & 1; }
  g.ra = rot_a; g.rb = rot_b; g.k = rot_a < 0 ? 0 : k;
  if (g.ra >= 0 && (g.k & 1)) { g.s[g.ra] = g.in[g.rb]; g.s[g.rb] = g.in[g.ra]; }
  const int64_t total = planes * in_size[0] * in_size[1] * in_size[2];
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == PCB_F32) tta_view_kernel<float><<<tta_grid(total), 256, 0, st>>>((const float*)x, (float*)out, g, planes);
  else if (dtype == PCB_F16) tta_view_kernel<__half><<<tta_grid(total), 256, 0, st>>>((const __half*)x, (__half*)out, g, planes);
  else if (dtype == PCB_BF16) tta_view_kernel<__nv_bfloat16><<<tta_grid(total), 256, 0, st>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)out, g, planes);
  else { set_error("pcb_tta_view: bad dtype %d", dtype); return PCB_ERR_INVALID; }
  PCB_CHECK_LAUNCH("pcb_tta_view");
  return PCB_OK;
}

static int tta_fill_geom(TtaFold& f, const int64_t acc_size[3], int flip_mask, int rot_a, int rot_b, int k) {
  TtaGeom& g = f.g;
  for (int a = 0; a < 3; ++a) { g.s[a] = (int)acc_size[a]; g.in[a] = (int)acc_size[a]; g.flip[a] = (flip_mask >> a) & 1; }
  g.ra = rot_a; g.rb = rot_b;
  const int kf = rot_a < 0 ? 0 : k;
  g.k = (4 - kf) & 3;                                   // invert_view rotates by -k
  if (g.ra >= 0 && (kf & 1)) { g.in[g.ra] = g.s[g.rb]; g.in[g.rb] = g.s[g.ra]; }   // the view frame has the plane dims swapped
  return 0;
}

extern "C" int pcb_tta_fold_ex(const void* pred, int pred_dtype, void* acc, int acc_dtype, int64_t N, int64_t Cpred, int64_t Cacc,
                               const int64_t acc_size[3], int flip_mask, int rot_a, int rot_b, int k, const int* src_channel,
                               const int* mode, const int* act, const float* act_scale, const int* shift, const int* sm_off,
                               const int* sm_len, const int* sm_src, const int* sm_shift, int n_members, const int* part,
                               const int* part_mode, int64_t Cpart, float* stats, void* counts, int count_dtype,
                               const void* valid_mask, int mean_as_sum, int n_prev, void* stream) {
  PCB_CHECK_ARG(pred && acc && geom_ok(acc_size) && N > 0 && src_channel && mode && act && act_scale, "pcb_tta_fold: bad argument");
  PCB_CHECK_ARG(Cacc > 0 && Cacc <= TTA_MAXC && Cpred > 0, "pcb_tta_fold: at most %d output channels (got %lld)", TTA_MAXC, (long long)Cacc);
  PCB_CHECK_ARG(k >= 0 && k <= 3 && n_prev >= 0, "pcb_tta_fold: bad k / n_prev");
  PCB_CHECK_ARG((rot_a < 0 && rot_b < 0) || (rot_a >= 0 && rot_a < 3 && rot_b >= 0 && rot_b < 3 && rot_a != rot_b),
                "pcb_tta_fold: bad rotation plane (%d, %d)", rot_a, rot_b);
  PCB_CHECK_ARG(n_members >= 0 && n_members <= TTA_MAXC && (n_members == 0 || (sm_off && sm_len && sm_src)),
                "pcb_tta_fold: bad softmax group description (%d members)", n_members);
  PCB_CHECK_ARG(Cpart >= 0 && Cpart <= Cacc && (Cpart == 0 || (part && part_mode && stats && counts && (count_dtype == 0 || count_dtype == 1))),
                "pcb_tta_fold: bad partial-channel description");
  TtaFold f;
  memset(&f, 0, sizeof(f));
  tta_fill_geom(f, acc_size, flip_mask, rot_a, rot_b, k);
  f.N = (int)N; f.Cacc = (int)Cacc; f.Cpred = (int)Cpred; f.n_prev = n_prev; f.Cpart = (int)Cpart; f.mean_as_sum = mean_as_sum ? 1 : 0;
  for (int c = 0; c < Cacc; ++c) {
    PCB_CHECK_ARG(src_channel[c] >= 0 && src_channel[c] < Cpred, "pcb_tta_fold: source channel %d out of range", src_channel[c]);
    PCB_CHECK_ARG(mode[c] >= 0 && mode[c] <= 3 && act[c] >= 0 && act[c] <= 4, "pcb_tta_fold: bad mode / activation code");
    f.src[c] = src_channel[c]; f.mode[c] = mode[c]; f.act[c] = act[c]; f.scale[c] = act_scale[c];
    f.part[c] = -1;
    if (shift) for (int a = 0; a < 3; ++a) { f.shift[c][a] = shift[c * 3 + a]; if (shift[c * 3 + a]) f.any_shift = 1; }
    if (act[c] == 4) {
      PCB_CHECK_ARG(sm_off && sm_len && sm_off[c] >= 0 && sm_len[c] >= 1 && sm_off[c] + sm_len[c] <= n_members,
                    "pcb_tta_fold: channel %d has a softmax activation but no group", c);
      f.sm_off[c] = sm_off[c]; f.sm_len[c] = sm_len[c];
    }
    if (Cpart > 0 && part[c] >= 0) {
      PCB_CHECK_ARG(part[c] < Cpart && part_mode[c] >= 0 && part_mode[c] <= 2, "pcb_tta_fold: bad partial index / mode of channel %d", c);
      f.part[c] = part[c]; f.pmode[c] = part_mode[c]; f.mode[c] = 3;
    }
  }
  for (int m = 0; m < n_members; ++m) {
    PCB_CHECK_ARG(sm_src[m] >= 0 && sm_src[m] < Cpred, "pcb_tta_fold: softmax member channel %d out of range", sm_src[m]);
    f.sm_src[m] = sm_src[m];
    if (sm_shift) for (int a = 0; a < 3; ++a) f.sm_shift[m][a] = sm_shift[m * 3 + a];
  }
  const int64_t total = N * Cacc * acc_size[0] * acc_size[1] * acc_size[2];
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = tta_grid(total);
  const uint8_t* vm = (const uint8_t*)valid_mask;
#define PCB_FOLD(TP, TA)                                                                                          \
  do {                                                                                                            \
    if (count_dtype == 1) tta_fold_kernel<TP, TA, int16_t><<<grid, 256, 0, st>>>((const TP*)pred, (TA*)acc, stats, (int16_t*)counts, vm, f); \
    else tta_fold_kernel<TP, TA, uint8_t><<<grid, 256, 0, st>>>((const TP*)pred, (TA*)acc, stats, (uint8_t*)counts, vm, f);                  \
  } while (0)
#define PCB_FOLD_P(TP)                                                          \
  do {                                                                          \
    if (acc_dtype == PCB_F32) PCB_FOLD(TP, float);                              \
    else if (acc_dtype == PCB_F16) PCB_FOLD(TP, __half);                        \
    else if (acc_dtype == PCB_BF16) PCB_FOLD(TP, __nv_bfloat16);                \
    else { set_error("pcb_tta_fold: bad dtype %d", acc_dtype); return PCB_ERR_INVALID; } \
  } while (0)
  if (pred_dtype == PCB_F32) PCB_FOLD_P(float);
  else if (pred_dtype == PCB_F16) PCB_FOLD_P(__half);
  else if (pred_dtype == PCB_BF16) PCB_FOLD_P(__nv_bfloat16);
  else { set_error("pcb_tta_fold: bad dtype %d", pred_dtype); return PCB_ERR_INVALID; }
#undef PCB_FOLD_P
#undef PCB_FOLD
  PCB_CHECK_LAUNCH("pcb_tta_fold");
  return PCB_OK;
}

extern "C" int pcb_tta_fold(const void* pred, int pred_dtype, void* acc, int acc_dtype, int64_t N, int64_t Cpred, int64_t Cacc,
                            const int64_t acc_size[3], int flip_mask, int rot_a, int rot_b, int k, const int* src_channel,
                            const int* mode, const int* act, const float* act_scale, int n_prev, void* stream) {
  if (mode && Cacc > 0 && Cacc <= TTA_MAXC)
    for (int c = 0; c < Cacc; ++c) PCB_CHECK_ARG(mode[c] >= 0 && mode[c] <= 2 && act && act[c] >= 0 && act[c] <= 3, "pcb_tta_fold: bad mode / activation code");
  return pcb_tta_fold_ex(pred, pred_dtype, acc, acc_dtype, N, Cpred, Cacc, acc_size, flip_mask, rot_a, rot_b, k, src_channel, mode, act,
                         act_scale, nullptr, nullptr, nullptr, nullptr, nullptr, 0, nullptr, nullptr, 0, nullptr, nullptr, 0, nullptr, 0,
                         n_prev, stream);
}

extern "C" int pcb_tta_unview(const void* pred, void* out, int dtype, int64_t N, int64_t Cpred, int64_t Cout, const int64_t out_size[3],
                              int flip_mask, int rot_a, int rot_b, int k, const int* src_channel, const int* shift, void* stream) {
  PCB_CHECK_ARG(pred && out && geom_ok(out_size) && N > 0 && Cpred > 0, "pcb_tta_unview: bad argument");
  PCB_CHECK_ARG(Cout > 0 && Cout <= TTA_MAXC, "pcb_tta_unview: at most %d channels (got %lld)", TTA_MAXC, (long long)Cout);
  PCB_CHECK_ARG(k >= 0 && k <= 3, "pcb_tta_unview: k must be in 0..3 (got %d)", k);
  PCB_CHECK_ARG((rot_a < 0 && rot_b < 0) || (rot_a >= 0 && rot_a < 3 && rot_b >= 0 && rot_b < 3 && rot_a != rot_b),
                "pcb_tta_unview: bad rotation plane (%d, %d)", rot_a, rot_b);
  TtaFold f;
  memset(&f, 0, sizeof(f));
  tta_fill_geom(f, out_size, flip_mask, rot_a, rot_b, k);
  f.N = (int)N; f.Cacc = (int)Cout; f.Cpred = (int)Cpred;
  for (int c = 0; c < Cout; ++c) {
    const int sc = src_channel ? src_channel[c] : c;
    PCB_CHECK_ARG(sc >= 0 && sc < Cpred, "pcb_tta_unview: source channel %d out of range", sc);
    f.src[c] = sc;
    if (shift) for (int a = 0; a < 3; ++a) f.shift[c][a] = shift[c * 3 + a];
  }
  const int64_t total = N * Cout * out_size[0] * out_size[1] * out_size[2];
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = tta_grid(total);
  if (dtype == PCB_F32) tta_unview_kernel<float><<<grid, 256, 0, st>>>((const float*)pred, (float*)out, f);
  else if (dtype == PCB_F16) tta_unview_kernel<__half><<<grid, 256, 0, st>>>((const __half*)pred, (__half*)out, f);
  else if (dtype == PCB_BF16) tta_unview_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)pred, (__nv_bfloat16*)out, f);
  else { set_error("pcb_tta_unview: bad dtype %d", dtype); return PCB_ERR_INVALID; }
  PCB_CHECK_LAUNCH("pcb_tta_unview");
  return PCB_OK;
}

extern "C" int pcb_tta_finalize_partial(const float* stats, const void* counts, int count_dtype, void* acc, int acc_dtype, int64_t N,
                                        int64_t Cacc, int64_t Cpart, const int* part_channel, const int* part_mode, int64_t nvox,
                                        void* first_zero, void* stream) {
  PCB_CHECK_ARG(stats && counts && acc && first_zero && part_channel && part_mode && N > 0 && nvox > 0, "pcb_tta_finalize_partial: bad argument");
  PCB_CHECK_ARG(Cpart > 0 && Cpart <= Cacc && Cacc <= TTA_MAXC && (count_dtype == 0 || count_dtype == 1), "pcb_tta_finalize_partial: bad channel counts");
  TtaFold f;
  memset(&f, 0, sizeof(f));
  f.N = (int)N; f.Cacc = (int)Cacc; f.Cpart = (int)Cpart;
  for (int j = 0; j < Cpart; ++j) {
    PCB_CHECK_ARG(part_channel[j] >= 0 && part_channel[j] < Cacc && part_mode[j] >= 0 && part_mode[j] <= 2, "pcb_tta_finalize_partial: bad channel / mode");
    f.src[j] = part_channel[j]; f.pmode[j] = part_mode[j];
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = tta_grid(N * Cpart * nvox);
  unsigned long long* fz = (unsigned long long*)first_zero;
#define PCB_FIN(TA)                                                                                                              \
  do {                                                                                                                           \
    if (count_dtype == 1) tta_finalize_kernel<TA, int16_t><<<grid, 256, 0, st>>>(stats, (const int16_t*)counts, (TA*)acc, f, nvox, fz); \
    else tta_finalize_kernel<TA, uint8_t><<<grid, 256, 0, st>>>(stats, (const uint8_t*)counts, (TA*)acc, f, nvox, fz);                  \
  } while (0)
  if (acc_dtype == PCB_F32) PCB_FIN(float);
  else if (acc_dtype == PCB_F16) PCB_FIN(__half);
  else if (acc_dtype == PCB_BF16) PCB_FIN(__nv_bfloat16);
  else { set_error("pcb_tta_finalize_partial: bad dtype %d", acc_dtype); return PCB_ERR_INVALID; }
#undef PCB_FIN
  PCB_CHECK_LAUNCH("pcb_tta_finalize_partial");
  return PCB_OK;
}
