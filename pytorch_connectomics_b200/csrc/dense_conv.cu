// pcb200 — dense 3-D convolution path for the MONAI U-Net (`monai_unet`, BASELINE config 1).
//
//   conv_igemm_kernel    Conv3d / ConvTranspose3d as an implicit GEMM on tcgen05: per 128-voxel output tile the
//                        K loop runs over (tap, input-channel chunk); the A operand of every step is the tile's
//                        input rows for that tap, GATHERED from the channels-last activation (zero rows outside
//                        the volume / on the wrong parity of a transposed conv) straight into the K-major
//                        canonical shared-memory layout; B is the tap's [Cout x Cin] weight slab; fp32 accumulation
//                        in TMEM over all taps; epilogue adds the bias.  The same kernel computes data gradients
//                        (conv <-> transposed conv with repacked weights).
//   channel_stats_kernel per-channel sum / sum of squares over N*V (BatchNorm batch statistics), float64.
//   bn_act_kernel        y = PReLU(x*scale + shift)   (BatchNorm folded to a per-channel affine) elementwise.
//   bn_act_bwd_kernel    dz = dy * PReLU'(z) (bf16), S1 = sum dz, S2 = sum dz*xhat, dslope = sum dy*min(z,0).
// Channel counts are padded to multiples of 16 by the host (padded channels carry zeros).
#include "../../include/pcb200.h"
#include "pcb_common.cuh"

namespace pcb {

struct ConvArgs {
  const uint4* x; const uint4* w; const float* bias; uint4* out;
  int D, H, W;          // input spatial size
  int Do, Ho, Wo;       // output spatial size
  int Ci, Co;           // padded channel counts
  int k, stride, pad;   // kernel size, stride, padding
  int transposed;       // 0: src = o*stride + t - pad ; 1: src = (o + pad - t)/stride when divisible
  int KC, NT;
  int64_t Vout, Vin;
};

__global__ void __launch_bounds__(128) conv_igemm_kernel(ConvArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5;
  const int n = blockIdx.y, nt = blockIdx.z;
  const int tile0 = blockIdx.x * 128;
  uint8_t* sA = smem;                           // [128 x KC]
  uint8_t* sW = sA + 128 * a.KC * 2;            // [NT x KC]
  int* sOz = reinterpret_cast<int*>(sW + a.NT * a.KC * 2);   // [128] output coordinates of the tile rows
  int* sOy = sOz + 128;
  int* sOx = sOy + 128;
  int* sSrc = sOx + 128;                        // [128] gathered input row of the current tap (or -1)
  uint64_t* bar = reinterpret_cast<uint64_t*>(sSrc + 128);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  const uint32_t tmem_cols = tmem_cols_pow2(a.NT);
  if (warp == 0) tmem_alloc(tmem_slot, tmem_cols);
  if (tid == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  {
    const int ov = tile0 + tid;
    if (ov < (int)a.Vout) {
      sOx[tid] = ov % a.Wo; sOy[tid] = (ov / a.Wo) % a.Ho; sOz[tid] = ov / (a.Wo * a.Ho);
    } else {
      sOz[tid] = -1000000; sOy[tid] = 0; sOx[tid] = 0;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t acc = *tmem_slot;
  const uint32_t idesc = umma_idesc_bf16(128, a.NT, 0, 0);
  const int kc8 = a.KC >> 3, ci8 = a.Ci >> 3;
  const int ntap = a.k * a.k * a.k;
  const uint4* xn = a.x + (int64_t)n * a.Vin * ci8;
  uint32_t ph = 0;
  bool first = true;
  for (int tap = 0; tap < ntap; ++tap) {
    const int tz = tap / (a.k * a.k), ty = (tap / a.k) % a.k, tx = tap % a.k;
    {   // source row of every tile row for this tap
      const int oz = sOz[tid], oy = sOy[tid], ox = sOx[tid];
      int iz, iy, ix;
      bool ok = oz >= 0;
      if (!a.transposed) {
        iz = oz * a.stride + tz - a.pad; iy = oy * a.stride + ty - a.pad; ix = ox * a.stride + tx - a.pad;
      } else {
        const int qz = oz + a.pad - tz, qy = oy + a.pad - ty, qx = ox + a.pad - tx;
        ok = ok && qz >= 0 && qy >= 0 && qx >= 0 && (qz % a.stride == 0) && (qy % a.stride == 0) && (qx % a.stride == 0);
        iz = qz / a.stride; iy = qy / a.stride; ix = qx / a.stride;
      }
      ok = ok && iz >= 0 && iz < a.D && iy >= 0 && iy < a.H && ix >= 0 && ix < a.W;
      sSrc[tid] = ok ? (iz * a.H + iy) * a.W + ix : -1;
    }
    __syncthreads();
    for (int kc = 0; kc < a.Ci / a.KC; ++kc) {
      if (!first) { mbar_wait(bar, ph); ph ^= 1; }   // previous MMAs have consumed sA / sW
      first = false;
      const uint32_t sbo = kc8 * 128;
      staged_copy<8>(128 * kc8, tid, 128,
          [&](int q) {
            const int r = q >> __ffs(kc8) - 1, c8 = q & (kc8 - 1);
            const int src = sSrc[r];
            return src >= 0 ? __ldg(xn + (int64_t)src * ci8 + kc * kc8 + c8) : make_uint4(0, 0, 0, 0);
          },
          [&](int q, const uint4& v) {
            const int r = q >> __ffs(kc8) - 1, c8 = q & (kc8 - 1);
            *reinterpret_cast<uint4*>(sA + (r >> 3) * sbo + c8 * 128 + (r & 7) * 16) = v;
          });
      const uint4* wsrc = a.w + ((int64_t)tap * a.Co + nt * a.NT) * ci8 + kc * kc8;
      staged_copy<8>(a.NT * kc8, tid, 128,
          [&](int q) { const int r = q >> __ffs(kc8) - 1, c8 = q & (kc8 - 1); return __ldg(wsrc + (int64_t)r * ci8 + c8); },
          [&](int q, const uint4& v) {
            const int r = q >> __ffs(kc8) - 1, c8 = q & (kc8 - 1);
            *reinterpret_cast<uint4*>(sW + (r >> 3) * sbo + c8 * 128 + (r & 7) * 16) = v;
          });
      fence_proxy_async_smem();
      __syncthreads();
      if (tid == 0) {
        tc_fence_after();
        const uint64_t ad = umma_desc(smem_u32(sA), 128, kc8 * 128), bd = umma_desc(smem_u32(sW), 128, kc8 * 128);
        for (int k = 0; k < a.KC / 16; ++k)
          umma_bf16(acc, ad + (uint64_t)(k * 16), bd + (uint64_t)(k * 16), idesc, (tap > 0 || kc > 0 || k > 0) ? 1u : 0u);
        tc_commit(bar);
      }
    }
    __syncthreads();   // sSrc is rewritten for the next tap
  }
  mbar_wait(bar, ph);
  tc_fence_after();
  {
    const int r = tile0 + tid;
    const uint32_t trow = acc + ((uint32_t)(warp * 32) << 16);
    const int64_t orow = ((int64_t)n * a.Vout + r) * (a.Co >> 3) + nt * (a.NT >> 3);
    for (int c16 = 0; c16 < a.NT / 16; ++c16) {
      uint32_t v[16];
      tmem_ld16(trow + c16 * 16, v);
      tmem_ld_wait();
      if (r >= (int)a.Vout) continue;
      float o[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) o[j] = __uint_as_float(v[j]) + (a.bias ? __ldg(a.bias + nt * a.NT + c16 * 16 + j) : 0.f);
      a.out[orow + c16 * 2] = pack8(o);
      a.out[orow + c16 * 2 + 1] = pack8(o + 8);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(acc, tmem_cols);
}

// ---------------------------------------------------------------------------- BatchNorm / PReLU (channels-last)
__global__ void __launch_bounds__(256) channel_stats_kernel(const uint4* __restrict__ x, double* __restrict__ stats, int C,
                                                            int64_t rows) {
  extern __shared__ double s_st[];   // [2*C]
  const int CH = C >> 3;
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) s_st[i] = 0.0;
  __syncthreads();
  const int64_t items = rows * CH, stride = (int64_t)gridDim.x * blockDim.x;
  float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, q[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  int cc_fixed = -1;
  int cnt = 0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < items; i += stride) {
    const int cc = (int)(i % CH);
    if ((cc_fixed >= 0 && cc != cc_fixed) || cnt == 64) {   // flush: channel change or keep fp32 partials short
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        atomicAdd(&s_st[cc_fixed * 8 + j], (double)s[j]); atomicAdd(&s_st[C + cc_fixed * 8 + j], (double)q[j]);
        s[j] = 0.f; q[j] = 0.f;
      }
      cnt = 0;
    }
    cc_fixed = cc;
    float f[8];
    unpack8(__ldg(x + i), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) { s[j] += f[j]; q[j] = fmaf(f[j], f[j], q[j]); }
    ++cnt;
  }
  if (cc_fixed >= 0) {
#pragma unroll
    for (int j = 0; j < 8; ++j) { atomicAdd(&s_st[cc_fixed * 8 + j], (double)s[j]); atomicAdd(&s_st[C + cc_fixed * 8 + j], (double)q[j]); }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) atomicAdd(&stats[i], s_st[i]);
}

__global__ void __launch_bounds__(256) bn_act_kernel(const uint4* __restrict__ x, const float* __restrict__ scale,
                                                     const float* __restrict__ shift, const float* __restrict__ slope_p,
                                                     uint4* __restrict__ out, int C, int64_t rows) {
  const int CH = C >> 3;
  const float slope = __ldg(slope_p);
  const int64_t items = rows * CH;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < items; i += (int64_t)gridDim.x * blockDim.x) {
    const int c0 = (int)(i % CH) * 8;
    float f[8];
    unpack8(__ldg(x + i), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float z = fmaf(f[j], __ldg(scale + c0 + j), __ldg(shift + c0 + j));
      f[j] = z > 0.f ? z : slope * z;
    }
    out[i] = pack8(f);
  }
}

// dz = dy * PReLU'(z); red[0..C) += dz ; red[C..2C) += dz * xhat ; red[2C] += dy * min(z, 0)
__global__ void __launch_bounds__(256) bn_act_bwd_kernel(const uint4* __restrict__ dy, const uint4* __restrict__ x,
                                                         const float* __restrict__ scale, const float* __restrict__ shift,
                                                         const float* __restrict__ mean, const float* __restrict__ rstd,
                                                         const float* __restrict__ slope_p, uint4* __restrict__ dz,
                                                         double* __restrict__ red, int C, int64_t rows) {
  extern __shared__ double s_r[];   // [2*C + 1]
  const int CH = C >> 3;
  const float slope = __ldg(slope_p);
  for (int i = threadIdx.x; i < 2 * C + 1; i += blockDim.x) s_r[i] = 0.0;
  __syncthreads();
  const int64_t items = rows * CH, stride = (int64_t)gridDim.x * blockDim.x;
  float s1[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, s2[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  float da = 0.f;
  int cc_fixed = -1, cnt = 0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < items; i += stride) {
    const int cc = (int)(i % CH);
    if ((cc_fixed >= 0 && cc != cc_fixed) || cnt == 64) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        atomicAdd(&s_r[cc_fixed * 8 + j], (double)s1[j]); atomicAdd(&s_r[C + cc_fixed * 8 + j], (double)s2[j]);
        s1[j] = 0.f; s2[j] = 0.f;
      }
      cnt = 0;
    }
    cc_fixed = cc;
    float g[8], f[8], o[8];
    unpack8(__ldg(dy + i), g);
    unpack8(__ldg(x + i), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = cc * 8 + j;
      const float z = fmaf(f[j], __ldg(scale + c), __ldg(shift + c));
      const float d = round_bf16(z > 0.f ? g[j] : slope * g[j]);
      if (z <= 0.f) da = fmaf(g[j], z, da);
      o[j] = d;
      s1[j] += d;
      s2[j] = fmaf(d, (f[j] - __ldg(mean + c)) * __ldg(rstd + c), s2[j]);
    }
    dz[i] = pack8(o);
    ++cnt;
  }
  if (cc_fixed >= 0) {
#pragma unroll
    for (int j = 0; j < 8; ++j) { atomicAdd(&s_r[cc_fixed * 8 + j], (double)s1[j]); atomicAdd(&s_r[C + cc_fixed * 8 + j], (double)s2[j]); }
  }
  atomicAdd(&s_r[2 * C], (double)da);
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C + 1; i += blockDim.x) atomicAdd(&red[i], s_r[i]);
}

static inline int pick_chunk_c(int64_t n, int cap) {
  for (int c = cap; c >= 16; c >>= 1)
    if (n % c == 0) return c;
  return 0;
}

}  // namespace pcb

using namespace pcb;

extern "C" int pcb_conv_fwd(const void* x, const void* w, const float* bias, void* out, int64_t N, const int64_t in_size[3],
                            const int64_t out_size[3], int64_t Ci, int64_t Co, int k, int stride, int pad, int transposed,
                            void* stream) {
  PCB_CHECK_ARG(x && w && out && in_size && out_size, "pcb_conv_fwd: null argument");
  PCB_CHECK_ARG(Ci % 16 == 0 && Co % 16 == 0 && Ci > 0 && Co > 0, "pcb_conv_fwd: channel counts must be padded to multiples of 16");
  PCB_CHECK_ARG((k == 1 || k == 3) && (stride == 1 || stride == 2) && pad >= 0 && pad <= 1, "pcb_conv_fwd: unsupported k/stride/pad");
  PCB_CHECK_ARG(N > 0 && N <= 65535, "pcb_conv_fwd: bad batch");
  ConvArgs a;
  a.x = (const uint4*)x; a.w = (const uint4*)w; a.bias = bias; a.out = (uint4*)out;
  a.D = (int)in_size[0]; a.H = (int)in_size[1]; a.W = (int)in_size[2];
  a.Do = (int)out_size[0]; a.Ho = (int)out_size[1]; a.Wo = (int)out_size[2];
  a.Ci = (int)Ci; a.Co = (int)Co; a.k = k; a.stride = stride; a.pad = pad; a.transposed = transposed;
  a.KC = pick_chunk_c(Ci, 128);
  a.NT = Co <= 256 ? (int)Co : 256;
  PCB_CHECK_ARG(Co % a.NT == 0, "pcb_conv_fwd: Co must be <= 256 or a multiple of 256");
  a.Vout = (int64_t)a.Do * a.Ho * a.Wo; a.Vin = (int64_t)a.D * a.H * a.W;
  PCB_CHECK_ARG(a.Vout < (1ll << 30) && a.Vin < (1ll << 30), "pcb_conv_fwd: volume too large");
  const size_t smem = (size_t)128 * a.KC * 2 + (size_t)a.NT * a.KC * 2 + 4 * 128 * sizeof(int) + 32;
  static DevFlag configured;
  if (!configured) {
    if (cudaFuncSetAttribute(conv_igemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) {
      set_error("pcb_conv_fwd: cudaFuncSetAttribute failed"); return PCB_ERR_CUDA;
    }
    configured = true;
  }
  dim3 grid((unsigned)((a.Vout + 127) / 128), (unsigned)N, (unsigned)(Co / a.NT));
  conv_igemm_kernel<<<grid, 128, smem, (cudaStream_t)stream>>>(a);
  PCB_CHECK_LAUNCH("pcb_conv_fwd");
  return PCB_OK;
}

extern "C" int pcb_channel_stats(const void* x, double* stats, int64_t C, int64_t rows, void* stream) {
  PCB_CHECK_ARG(x && stats && C % 8 == 0 && C > 0 && rows > 0, "pcb_channel_stats: bad argument");
  int blocks = (int)((rows * (C / 8) + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  channel_stats_kernel<<<blocks, 256, 2 * C * sizeof(double), (cudaStream_t)stream>>>((const uint4*)x, stats, (int)C, rows);
  PCB_CHECK_LAUNCH("pcb_channel_stats");
  return PCB_OK;
}

extern "C" int pcb_bn_act_fwd(const void* x, const float* scale, const float* shift, const float* slope, void* out, int64_t C,
                              int64_t rows, void* stream) {
  PCB_CHECK_ARG(x && scale && shift && slope && out && C % 8 == 0 && C > 0 && rows > 0, "pcb_bn_act_fwd: bad argument");
  int blocks = (int)((rows * (C / 8) + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  bn_act_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>((const uint4*)x, scale, shift, slope, (uint4*)out, (int)C, rows);
  PCB_CHECK_LAUNCH("pcb_bn_act_fwd");
  return PCB_OK;
}

extern "C" int pcb_bn_act_bwd(const void* dy, const void* x, const float* scale, const float* shift, const float* mean,
                              const float* rstd, const float* slope, void* dz, double* red, int64_t C, int64_t rows, void* stream) {
  PCB_CHECK_ARG(dy && x && scale && shift && mean && rstd && slope && dz && red && C % 8 == 0 && C > 0 && rows > 0, "pcb_bn_act_bwd: bad argument");
  int blocks = (int)((rows * (C / 8) + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  bn_act_bwd_kernel<<<blocks, 256, (2 * C + 1) * sizeof(double), (cudaStream_t)stream>>>(
      (const uint4*)dy, (const uint4*)x, scale, shift, mean, rstd, slope, (uint4*)dz, red, (int)C, rows);
  PCB_CHECK_LAUNCH("pcb_bn_act_bwd");
  return PCB_OK;
}
