// pcb200 — MedNeXt forward kernels for sm_100a.
//
// Data layout in HBM: activations are channels-last bf16 [N, D, H, W, C]; one voxel's channel
// vector is a contiguous C*2-byte row, so every global access below is a 128-bit load/store of 8
// channels.  GroupNorm(num_groups=C) statistics are per-(n,c) sum / sum-of-squares in float64.
//
//   stem_kernel      1x1x1 conv Cin->C, NCDHW(any dtype) -> NDHWC bf16          (HBM-bound)
//   dwconv_kernel    depthwise k^3 stencil (same / stride-2 / transposed) + bias + GN statistics
//                    (HBM-bound; fp32 FMA on CUDA cores)
//   mlp_kernel       GN-apply -> GEMM1 (C->H) -> +b2, GELU -> GEMM2 (H->Co) [+ res-conv GEMM]
//                    -> +b3 (+residual/skip): tcgen05.mma (kind::f16, bf16 in / fp32 accum in
//                    TMEM), operands staged in shared memory in the no-swizzle canonical K-major
//                    layout, the expanded [128 x H] tile lives only in TMEM/shared memory.
//   head_kernel      OutBlock 1x1x1 (C -> ncls), NDHWC bf16 -> NCDHW(any dtype)  (HBM-bound)
#include "../../include/pcb200.h"
#include "pcb_common.cuh"

namespace pcb {

// ============================================================================ stem
template <typename TIn>
__global__ void __launch_bounds__(256) stem_kernel(const TIn* __restrict__ x, const float* __restrict__ w,
                                                   const float* __restrict__ b, uint4* __restrict__ out,
                                                   int64_t N, int Cin, int C, int64_t V) {
  const int CH = C >> 3;
  const int64_t total = N * V * CH;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int cc = (int)(i % CH);
    const int64_t v = (i / CH) % V, n = i / (CH * V);
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = b[cc * 8 + j];
    for (int ci = 0; ci < Cin; ++ci) {
      const float xv = (float)x[(n * Cin + ci) * V + v];
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = fmaf(xv, w[(cc * 8 + j) * Cin + ci], acc[j]);
    }
    out[i] = pack8(acc);
  }
}

// ============================================================================ depthwise conv
constexpr int DW_XB = 4;  // outputs per thread along W

struct DwArgs {
  int D, H, W, Do, Ho, Wo, C;
  int add_mode;  // 0 none, 1 add[o] (same index), 2 add[(o/2)] where every coordinate of o is even
  int a1, a2;    // H, W of the compact `add` tensor (add_mode 2)
};

template <int K, int MODE>
__global__ void __launch_bounds__(256) dwconv_kernel(const uint4* __restrict__ x, const float* __restrict__ w,
                                                     const float* __restrict__ bias, uint4* __restrict__ y,
                                                     double* __restrict__ stats, const uint4* __restrict__ add,
                                                     DwArgs a) {
  extern __shared__ double s_stats[];  // [2*C] float64: sums stay order-independent to ~1e-16
  constexpr int P = K / 2;
  constexpr int S = (MODE == PCB_DW_DOWN) ? 2 : 1;
  const int C = a.C, CH = C >> 3;
  const int n = blockIdx.y;
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) s_stats[i] = 0.0;
  __syncthreads();

  const int nstrip = (a.Wo + DW_XB - 1) / DW_XB;
  const int64_t items = (int64_t)a.Do * a.Ho * nstrip * CH;
  const int64_t item = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const bool active = item < items;
  float ssum[8], ssq[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { ssum[j] = 0.f; ssq[j] = 0.f; }
  const int cc = (int)(item % CH);
  if (active) {
    int64_t t = item / CH;
    const int xs = (int)(t % nstrip); t /= nstrip;
    const int oy = (int)(t % a.Ho);
    const int oz = (int)(t / a.Ho);
    const int ox0 = xs * DW_XB;
    float acc[DW_XB][8];
#pragma unroll
    for (int j = 0; j < DW_XB; ++j)
#pragma unroll
      for (int c = 0; c < 8; ++c) acc[j][c] = 0.f;
    const uint4* xn = x + (int64_t)n * a.D * a.H * a.W * CH;

    if (MODE != PCB_DW_UP) {
      constexpr int NIN = S * (DW_XB - 1) + K;
#pragma unroll 1
      for (int dz = 0; dz < K; ++dz) {
        const int iz = oz * S + dz - P;
        if (iz < 0 || iz >= a.D) continue;
#pragma unroll 1
        for (int dy = 0; dy < K; ++dy) {
          const int iy = oy * S + dy - P;
          if (iy < 0 || iy >= a.H) continue;
          const uint4* row = xn + ((int64_t)iz * a.H + iy) * a.W * CH + cc;
          float wv[K][8];
#pragma unroll
          for (int dx = 0; dx < K; ++dx) {
            const float4* wp = reinterpret_cast<const float4*>(w + ((dz * K + dy) * K + dx) * C + cc * 8);
            const float4 w0 = __ldg(wp), w1 = __ldg(wp + 1);
            wv[dx][0] = w0.x; wv[dx][1] = w0.y; wv[dx][2] = w0.z; wv[dx][3] = w0.w;
            wv[dx][4] = w1.x; wv[dx][5] = w1.y; wv[dx][6] = w1.z; wv[dx][7] = w1.w;
          }
#pragma unroll
          for (int i = 0; i < NIN; ++i) {
            const int ix = ox0 * S + i - P;
            if (ix < 0 || ix >= a.W) continue;
            float f[8];
            unpack8(__ldg(row + (int64_t)ix * CH), f);
#pragma unroll
            for (int dx = 0; dx < K; ++dx) {
              if ((i - dx) >= 0 && (i - dx) % S == 0 && (i - dx) / S < DW_XB) {
                const int j = (i - dx) / S;
#pragma unroll
                for (int c = 0; c < 8; ++c) acc[j][c] = fmaf(f[c], wv[dx][c], acc[j][c]);
              }
            }
          }
        }
      }
    } else {
      // ConvTranspose3d(stride 2, padding P): out[o] += in[i] * w[k] with o = 2 i - P + k
#pragma unroll 1
      for (int kz = 0; kz < K; ++kz) {
        const int tz = oz + P - kz;
        if (tz < 0 || (tz & 1) || (tz >> 1) >= a.D) continue;
#pragma unroll 1
        for (int ky = 0; ky < K; ++ky) {
          const int ty = oy + P - ky;
          if (ty < 0 || (ty & 1) || (ty >> 1) >= a.H) continue;
          const uint4* row = xn + ((int64_t)(tz >> 1) * a.H + (ty >> 1)) * a.W * CH + cc;
#pragma unroll
          for (int kx = 0; kx < K; ++kx) {
            const float4* wp = reinterpret_cast<const float4*>(w + ((kz * K + ky) * K + kx) * C + cc * 8);
            const float4 w0 = __ldg(wp), w1 = __ldg(wp + 1);
            const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
            for (int j = 0; j < DW_XB; ++j) {
              const int tx = ox0 + j + P - kx;
              if (tx < 0 || (tx & 1) || (tx >> 1) >= a.W) continue;
              float f[8];
              unpack8(__ldg(row + (int64_t)(tx >> 1) * CH), f);
#pragma unroll
              for (int c = 0; c < 8; ++c) acc[j][c] = fmaf(f[c], wv[c], acc[j][c]);
            }
          }
        }
      }
    }
    float bv[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (bias != nullptr) {
      const float4* bp = reinterpret_cast<const float4*>(bias + cc * 8);
      const float4 b0 = __ldg(bp), b1 = __ldg(bp + 1);
      bv[0] = b0.x; bv[1] = b0.y; bv[2] = b0.z; bv[3] = b0.w; bv[4] = b1.x; bv[5] = b1.y; bv[6] = b1.z; bv[7] = b1.w;
    }
    uint4* yrow = y + (((int64_t)n * a.Do + oz) * a.Ho + oy) * a.Wo * CH + cc;
#pragma unroll
    for (int j = 0; j < DW_XB; ++j) {
      if (ox0 + j >= a.Wo) continue;
      float o[8];
      if (a.add_mode == 1) {
        float f[8];
        unpack8(__ldg(add + ((((int64_t)n * a.Do + oz) * a.Ho + oy) * a.Wo + ox0 + j) * CH + cc), f);
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[j][c] += f[c];
      } else if (a.add_mode == 2) {
        if (!((oz | oy | (ox0 + j)) & 1)) {
          float f[8];
          unpack8(__ldg(add + ((((int64_t)n * ((a.Do + 1) >> 1) + (oz >> 1)) * a.a1 + (oy >> 1)) * a.a2 + ((ox0 + j) >> 1)) * CH + cc), f);
#pragma unroll
          for (int c = 0; c < 8; ++c) acc[j][c] += f[c];
        }
      }
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        o[c] = round_bf16(acc[j][c] + bv[c]);
        ssum[c] += o[c];
        ssq[c] = fmaf(o[c], o[c], ssq[c]);
      }
      yrow[(int64_t)(ox0 + j) * CH] = pack8(o);
    }
  }
  // per-channel partial statistics: warp shuffle over lanes sharing a channel chunk, then shared
  // atomics, then one float64 atomic per channel per CTA
  if (stats == nullptr) return;   // uniform across the grid (backward-data use)
  const bool shuffle_ok = (CH <= 32) && ((32 % CH) == 0);
  if (shuffle_ok) {
    for (int off = 16; off >= CH; off >>= 1) {
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        ssum[c] += __shfl_xor_sync(0xffffffffu, ssum[c], off);
        ssq[c] += __shfl_xor_sync(0xffffffffu, ssq[c], off);
      }
    }
  }
  if (active && (!shuffle_ok || (threadIdx.x & 31) < CH)) {
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      atomicAdd(&s_stats[cc * 8 + c], (double)ssum[c]);
      atomicAdd(&s_stats[C + cc * 8 + c], (double)ssq[c]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x)
    atomicAdd(&stats[(int64_t)n * 2 * C + i], s_stats[i]);
}

// ============================================================================ fused MLP (tcgen05)
struct MlpArgs {
  const uint4* y; const double* stats; const float* gamma; const float* beta;
  const uint4* w2; const float* b2; const uint4* w3; const float* b3;
  const uint4* res; const uint4* xs; const uint4* wr; const float* br; uint4* out;
  int o0, o1, o2;        // output spatial size
  int x0, x1, x2;        // xs (res-conv source) spatial size
  int C, H, Co, Cr;      // channels: in, hidden, out, res-conv in
  int KC, N1, CoT, KCr;  // GEMM1 K chunk, hidden chunk, out-channel tile, res-GEMM K chunk
  int mode;
  int64_t Vy, Vout, Vin; // voxels per sample of y / out / xs
  float inv_count;       // 1 / Vy
};

// copy a [rows x kc] bf16 block (row pitch `pitch8` in uint4 units) into the K-major canonical layout
__device__ __forceinline__ void stage_weights(uint8_t* dst, const uint4* __restrict__ src, int rows, int kc8,
                                              int64_t pitch8, int tid) {
  const uint32_t sbo = kc8 * 128;
  for (int q = tid; q < rows * kc8; q += 128) {
    const int r = q / kc8, c8 = q - r * kc8;
    const uint4 v = __ldg(src + (int64_t)r * pitch8 + c8);
    *reinterpret_cast<uint4*>(dst + (r >> 3) * sbo + c8 * 128 + (r & 7) * 16) = v;
  }
}

__global__ void __launch_bounds__(128) mlp_kernel(MlpArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5;
  const int n = blockIdx.y, cot = blockIdx.z;
  const int64_t tile0 = (int64_t)blockIdx.x * 128;

  // ---- shared memory carve-up
  const int KA = a.KC > a.KCr ? a.KC : a.KCr;
  const int KW3 = a.N1 > a.KCr ? a.N1 : a.KCr;
  uint8_t* sA = smem;                               // [128 x KA]  bf16
  uint8_t* sW2 = sA + 128 * KA * 2;                 // [N1 x KC]
  uint8_t* sH = sW2 + a.N1 * a.KC * 2;              // [128 x N1]
  uint8_t* sW3 = sH + 128 * a.N1 * 2;               // [CoT x max(N1,KCr)]
  float* sScale = reinterpret_cast<float*>(sW3 + a.CoT * KW3 * 2);  // [C]
  float* sShift = sScale + a.C;                     // [C]
  int64_t* sRowY = reinterpret_cast<int64_t*>(sShift + a.C);        // [128] source row in y (or -1)
  int64_t* sRowX = sRowY + 128;                     // [128] source row in xs (or -1)
  uint64_t* bar = reinterpret_cast<uint64_t*>(sRowX + 128);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);

  const uint32_t tmem_cols = tmem_cols_pow2(a.N1 + a.CoT);
  if (warp == 0) tmem_alloc(tmem_slot, tmem_cols);
  if (tid == 0) { mbar_init(bar, 1); fence_mbar_init(); }

  // GroupNorm(num_groups=C) affine from the float64 statistics: yhat = y*scale + shift
  for (int c = tid; c < a.C; c += 128) {
    const double s = a.stats[(int64_t)n * 2 * a.C + c], q = a.stats[(int64_t)n * 2 * a.C + a.C + c];
    const double mean = s * (double)a.inv_count;
    double var = q * (double)a.inv_count - mean * mean;
    if (var < 0.0) var = 0.0;
    const float rstd = (float)(1.0 / sqrt(var + 1e-5));
    const float g = a.gamma[c] * rstd;
    sScale[c] = g;
    sShift[c] = a.beta[c] - (float)mean * g;
  }
  // per-row source indices
  {
    const int64_t ov = tile0 + tid;
    int64_t ry = -1, rx = -1;
    if (ov < a.Vout) {
      if (a.mode == PCB_DW_UP) {
        const int ox = (int)(ov % a.o2), oy = (int)((ov / a.o2) % a.o1), oz = (int)(ov / ((int64_t)a.o2 * a.o1));
        if (ox >= 1 && oy >= 1 && oz >= 1) {
          ry = ((int64_t)(oz - 1) * (a.o1 - 1) + (oy - 1)) * (a.o2 - 1) + (ox - 1);
          if (!((ox - 1) & 1) && !((oy - 1) & 1) && !((oz - 1) & 1))
            rx = ((int64_t)((oz - 1) >> 1) * a.x1 + ((oy - 1) >> 1)) * a.x2 + ((ox - 1) >> 1);
        }
      } else if (a.mode == PCB_DW_DOWN) {
        const int ox = (int)(ov % a.o2), oy = (int)((ov / a.o2) % a.o1), oz = (int)(ov / ((int64_t)a.o2 * a.o1));
        ry = ov;
        rx = ((int64_t)(2 * oz) * a.x1 + 2 * oy) * a.x2 + 2 * ox;
      } else {
        ry = ov;
      }
    }
    sRowY[tid] = ry;
    sRowX[tid] = rx;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t acc1 = tmem_base, acc2 = tmem_base + a.N1;
  uint32_t ph = 0;

  const int nkc = a.C / a.KC, nhc = a.H / a.N1;
  const int kc8 = a.KC >> 3;
  const uint32_t idesc1 = umma_idesc_bf16(128, a.N1, 0, 0), idesc2 = umma_idesc_bf16(128, a.CoT, 0, 0);
  const uint4* yn = a.y + (int64_t)n * a.Vy * (a.C >> 3);

  for (int hc = 0; hc < nhc; ++hc) {
    // -------- GEMM1: acc1[128 x N1] = norm(Y)[128 x C] * W2[hc]^T
    for (int kc = 0; kc < nkc; ++kc) {
      if (!(nkc == 1 && hc > 0)) {  // single K chunk: the normalised A tile stays resident
        const uint32_t sbo = kc8 * 128;
        for (int q = tid; q < 128 * kc8; q += 128) {
          const int r = q / kc8, c8 = q - r * kc8;
          const int64_t ry = sRowY[r];
          uint4 v = make_uint4(0, 0, 0, 0);
          if (ry >= 0) {
            float f[8];
            unpack8(__ldg(yn + ry * (a.C >> 3) + kc * kc8 + c8), f);
            const int c0 = kc * a.KC + c8 * 8;
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] = fmaf(f[j], sScale[c0 + j], sShift[c0 + j]);
            v = pack8(f);
          }
          *reinterpret_cast<uint4*>(sA + (r >> 3) * sbo + c8 * 128 + (r & 7) * 16) = v;
        }
      }
      stage_weights(sW2, a.w2 + (int64_t)hc * a.N1 * (a.C >> 3) + kc * kc8, a.N1, kc8, a.C >> 3, tid);
      fence_proxy_async_smem();
      __syncthreads();
      if (tid == 0) {
        tc_fence_after();
        const uint64_t ad = umma_desc(smem_u32(sA), 128, kc8 * 128);
        const uint64_t bd = umma_desc(smem_u32(sW2), 128, kc8 * 128);
        for (int k = 0; k < a.KC / 16; ++k)
          umma_bf16(acc1, ad + (uint64_t)(k * 16), bd + (uint64_t)(k * 16), idesc1, (kc > 0 || k > 0) ? 1u : 0u);
        tc_commit(bar);
      }
      mbar_wait(bar, ph);
      ph ^= 1;
    }
    tc_fence_after();
    // -------- stage W3[:, hc] while the accumulator is drained
    const int n18 = a.N1 >> 3;
    stage_weights(sW3, a.w3 + (int64_t)cot * a.CoT * (a.H >> 3) + hc * n18, a.CoT, n18, a.H >> 3, tid);
    // -------- epilogue 1: +b2, GELU, bf16 -> sH (K-major A operand of GEMM2)
    {
      const int r = tid;
      const uint32_t trow = acc1 + ((uint32_t)(warp * 32) << 16);
      const uint32_t sbo = n18 * 128;
      uint8_t* dst = sH + (r >> 3) * sbo + (r & 7) * 16;
      for (int c16 = 0; c16 < a.N1 / 16; ++c16) {
        uint32_t v[16];
        tmem_ld16(trow + c16 * 16, v);
        tmem_ld_wait();
        float g[16];
        const float* bp = a.b2 + hc * a.N1 + c16 * 16;
#pragma unroll
        for (int j = 0; j < 16; ++j) g[j] = gelu_f(__uint_as_float(v[j]) + __ldg(bp + j));
        *reinterpret_cast<uint4*>(dst + (c16 * 2) * 128) = pack8(g);
        *reinterpret_cast<uint4*>(dst + (c16 * 2 + 1) * 128) = pack8(g + 8);
      }
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    // -------- GEMM2: acc2[128 x CoT] += H[128 x N1] * W3[cot, hc]^T
    if (tid == 0) {
      tc_fence_after();
      const uint64_t ad = umma_desc(smem_u32(sH), 128, n18 * 128);
      const uint64_t bd = umma_desc(smem_u32(sW3), 128, n18 * 128);
      for (int k = 0; k < a.N1 / 16; ++k)
        umma_bf16(acc2, ad + (uint64_t)(k * 16), bd + (uint64_t)(k * 16), idesc2, (hc > 0 || k > 0) ? 1u : 0u);
      tc_commit(bar);
    }
    mbar_wait(bar, ph);
    ph ^= 1;
  }
  // -------- residual 1x1 conv of the resampling blocks as one more GEMM into acc2
  if (a.wr != nullptr) {
    const int kr8 = a.KCr >> 3;
    const uint4* xn = a.xs + (int64_t)n * a.Vin * (a.Cr >> 3);
    for (int kc = 0; kc < a.Cr / a.KCr; ++kc) {
      const uint32_t sbo = kr8 * 128;
      for (int q = tid; q < 128 * kr8; q += 128) {
        const int r = q / kr8, c8 = q - r * kr8;
        const int64_t rx = sRowX[r];
        uint4 v = make_uint4(0, 0, 0, 0);
        if (rx >= 0) v = __ldg(xn + rx * (a.Cr >> 3) + kc * kr8 + c8);
        *reinterpret_cast<uint4*>(sA + (r >> 3) * sbo + c8 * 128 + (r & 7) * 16) = v;
      }
      stage_weights(sW3, a.wr + (int64_t)cot * a.CoT * (a.Cr >> 3) + kc * kr8, a.CoT, kr8, a.Cr >> 3, tid);
      fence_proxy_async_smem();
      __syncthreads();
      if (tid == 0) {
        tc_fence_after();
        const uint64_t ad = umma_desc(smem_u32(sA), 128, kr8 * 128);
        const uint64_t bd = umma_desc(smem_u32(sW3), 128, kr8 * 128);
        for (int k = 0; k < a.KCr / 16; ++k)
          umma_bf16(acc2, ad + (uint64_t)(k * 16), bd + (uint64_t)(k * 16), idesc2, 1u);
        tc_commit(bar);
      }
      mbar_wait(bar, ph);
      ph ^= 1;
    }
  }
  tc_fence_after();
  // -------- final epilogue: +b3 (+br) (+residual / skip) -> bf16 -> HBM
  {
    const int64_t ov = tile0 + tid;
    const uint32_t trow = acc2 + ((uint32_t)(warp * 32) << 16);
    const bool in_range = ov < a.Vout;
    const bool valid = in_range && sRowY[tid] >= 0;
    const int64_t orow = ((int64_t)n * a.Vout + ov) * (a.Co >> 3) + cot * (a.CoT >> 3);
    for (int c16 = 0; c16 < a.CoT / 16; ++c16) {
      uint32_t v[16];
      tmem_ld16(trow + c16 * 16, v);   // warp-collective: executed by every lane
      tmem_ld_wait();
      if (!in_range) continue;
      float o[16];
      const int co = cot * a.CoT + c16 * 16;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        float t = 0.f;
        if (valid) {
          t = __uint_as_float(v[j]) + __ldg(a.b3 + co + j);
          if (a.br != nullptr) t += __ldg(a.br + co + j);
        }
        o[j] = t;
      }
      if (a.res != nullptr) {
        float f[8];
        unpack8(__ldg(a.res + orow + c16 * 2), f);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] += f[j];
        unpack8(__ldg(a.res + orow + c16 * 2 + 1), f);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[8 + j] += f[j];
      }
      a.out[orow + c16 * 2] = pack8(o);
      a.out[orow + c16 * 2 + 1] = pack8(o + 8);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, tmem_cols);
}

// ============================================================================ head (OutBlock)
template <typename TOut>
__global__ void __launch_bounds__(256) head_kernel(const uint4* __restrict__ x, const float* __restrict__ w,
                                                   const float* __restrict__ b, TOut* __restrict__ out,
                                                   int64_t N, int C, int ncls, int64_t V) {
  extern __shared__ float s_w[];  // [ncls][C] + [ncls]
  for (int i = threadIdx.x; i < C * ncls; i += blockDim.x) {
    const int c = i / ncls, k = i - c * ncls;   // w is [C, ncls]
    s_w[k * C + c] = w[i];
  }
  for (int i = threadIdx.x; i < ncls; i += blockDim.x) s_w[ncls * C + i] = b[i];
  __syncthreads();
  const int CH = C >> 3;
  const int64_t total = N * V;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t n = i / V, v = i - n * V;
    for (int k0 = 0; k0 < ncls; k0 += 8) {
      float acc[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] = (k0 + k < ncls) ? s_w[ncls * C + k0 + k] : 0.f;
      for (int c8 = 0; c8 < CH; ++c8) {
        float f[8];
        unpack8(__ldg(x + i * CH + c8), f);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          if (k0 + k < ncls) {
            const float* wk = s_w + (k0 + k) * C + c8 * 8;
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[k] = fmaf(f[j], wk[j], acc[k]);
          }
        }
      }
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (k0 + k < ncls) out[(n * ncls + k0 + k) * V + v] = (TOut)acc[k];
    }
  }
}

static inline int pick_chunk(int64_t n, int cap) {
  for (int c = cap; c >= 16; c >>= 1)
    if (n % c == 0) return c;
  return 0;
}

}  // namespace pcb

using namespace pcb;

extern "C" int pcb_stem_fwd(const void* x, int in_dtype, const float* w, const float* b, void* out, int64_t N,
                            int64_t Cin, int64_t C, int64_t nvox, void* stream) {
  PCB_CHECK_ARG(x && w && b && out, "pcb_stem_fwd: null argument");
  PCB_CHECK_ARG(C > 0 && C % 8 == 0 && Cin > 0 && N > 0 && nvox > 0, "pcb_stem_fwd: C must be a positive multiple of 8 (got %lld)", (long long)C);
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t total = N * nvox * (C / 8);
  const int grid = (int)((total + 255) / 256 > 148 * 32 ? 148 * 32 : (total + 255) / 256);
  uint4* o = (uint4*)out;
  if (in_dtype == PCB_F32) stem_kernel<float><<<grid, 256, 0, st>>>((const float*)x, w, b, o, N, (int)Cin, (int)C, nvox);
  else if (in_dtype == PCB_F16) stem_kernel<__half><<<grid, 256, 0, st>>>((const __half*)x, w, b, o, N, (int)Cin, (int)C, nvox);
  else if (in_dtype == PCB_BF16) stem_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)x, w, b, o, N, (int)Cin, (int)C, nvox);
  else { set_error("pcb_stem_fwd: bad dtype %d", in_dtype); return PCB_ERR_INVALID; }
  PCB_CHECK_LAUNCH("pcb_stem_fwd");
  return PCB_OK;
}

template <int K>
static void launch_dw(int mode, dim3 grid, size_t smem, cudaStream_t st, const uint4* x, const float* w, const float* b,
                      uint4* y, double* stats, const uint4* add, DwArgs a) {
  if (mode == PCB_DW_SAME) dwconv_kernel<K, PCB_DW_SAME><<<grid, 256, smem, st>>>(x, w, b, y, stats, add, a);
  else if (mode == PCB_DW_DOWN) dwconv_kernel<K, PCB_DW_DOWN><<<grid, 256, smem, st>>>(x, w, b, y, stats, add, a);
  else dwconv_kernel<K, PCB_DW_UP><<<grid, 256, smem, st>>>(x, w, b, y, stats, add, a);
}

static int dwconv_launch(const void* x, const float* w, const float* b, void* y, double* stats, const void* add,
                         int add_mode, int64_t N, const int64_t in_size[3], const int64_t* out_size, int64_t C, int k,
                         int mode, void* stream, const char* what) {
  PCB_CHECK_ARG(x && w && y && in_size, "%s: null argument", what);
  PCB_CHECK_ARG(k == 3 || k == 5 || k == 7, "MedNeXt kernel_size must be 3, 5, or 7. Got: %d", k);
  PCB_CHECK_ARG(mode >= PCB_DW_SAME && mode <= PCB_DW_UP, "%s: bad mode %d", what, mode);
  PCB_CHECK_ARG(C > 0 && C % 8 == 0 && C <= 4096, "%s: C must be a multiple of 8 (got %lld)", what, (long long)C);
  PCB_CHECK_ARG(N > 0 && N <= 65535, "%s: bad batch %lld", what, (long long)N);
  DwArgs a;
  a.D = (int)in_size[0]; a.H = (int)in_size[1]; a.W = (int)in_size[2]; a.C = (int)C;
  a.add_mode = add ? add_mode : 0; a.a1 = a.a2 = 0;
  const int p = k / 2;
  if (out_size) { a.Do = (int)out_size[0]; a.Ho = (int)out_size[1]; a.Wo = (int)out_size[2]; }
  else if (mode == PCB_DW_SAME) { a.Do = a.D; a.Ho = a.H; a.Wo = a.W; }
  else if (mode == PCB_DW_DOWN) { a.Do = (a.D + 2 * p - k) / 2 + 1; a.Ho = (a.H + 2 * p - k) / 2 + 1; a.Wo = (a.W + 2 * p - k) / 2 + 1; }
  else { a.Do = (a.D - 1) * 2 - 2 * p + k; a.Ho = (a.H - 1) * 2 - 2 * p + k; a.Wo = (a.W - 1) * 2 - 2 * p + k; }
  PCB_CHECK_ARG(a.Do > 0 && a.Ho > 0 && a.Wo > 0, "%s: empty output", what);
  if (a.add_mode == 2) { a.a1 = (a.Ho + 1) >> 1; a.a2 = (a.Wo + 1) >> 1; }
  const int64_t items = (int64_t)a.Do * a.Ho * ((a.Wo + DW_XB - 1) / DW_XB) * (C / 8);
  dim3 grid((unsigned)((items + 255) / 256), (unsigned)N);
  const size_t smem = 2 * C * sizeof(double);
  cudaStream_t st = (cudaStream_t)stream;
  if (k == 3) launch_dw<3>(mode, grid, smem, st, (const uint4*)x, w, b, (uint4*)y, stats, (const uint4*)add, a);
  else if (k == 5) launch_dw<5>(mode, grid, smem, st, (const uint4*)x, w, b, (uint4*)y, stats, (const uint4*)add, a);
  else launch_dw<7>(mode, grid, smem, st, (const uint4*)x, w, b, (uint4*)y, stats, (const uint4*)add, a);
  PCB_CHECK_LAUNCH(what);
  return PCB_OK;
}

extern "C" int pcb_dwconv_fwd(const void* x, const float* w, const float* b, void* y, double* stats, int64_t N,
                              const int64_t in_size[3], int64_t C, int k, int mode, void* stream) {
  PCB_CHECK_ARG(b && stats, "pcb_dwconv_fwd: null argument");
  return dwconv_launch(x, w, b, y, stats, nullptr, 0, N, in_size, nullptr, C, k, mode, stream, "pcb_dwconv_fwd");
}

extern "C" int pcb_dwconv_bwd_data(const void* dy, const float* w, const void* add, int add_mode, void* dx, int64_t N,
                                   const int64_t dy_size[3], const int64_t dx_size[3], int64_t C, int k, int fwd_mode,
                                   void* stream) {
  PCB_CHECK_ARG(dx_size, "pcb_dwconv_bwd_data: null argument");
  // gradient of a SAME conv is a SAME conv with the flipped taps (caller flips); of a stride-2 conv the
  // transposed conv; of the transposed conv the stride-2 conv — all with the forward stencil kernel.
  const int mode = fwd_mode == PCB_DW_SAME ? PCB_DW_SAME : (fwd_mode == PCB_DW_DOWN ? PCB_DW_UP : PCB_DW_DOWN);
  return dwconv_launch(dy, w, nullptr, dx, nullptr, add, add_mode, N, dy_size, dx_size, C, k, mode, stream,
                       "pcb_dwconv_bwd_data");
}

extern "C" int pcb_mlp_fwd(const void* y, const double* stats, const float* gamma, const float* beta, const void* w2,
                           const float* b2, const void* w3, const float* b3, const void* res, const void* xs,
                           const void* wr, const float* br, void* out, int64_t N, const int64_t out_size[3],
                           const int64_t xs_size[3], int64_t C, int64_t H, int64_t Co, int64_t Cr, int mode,
                           void* stream) {
  PCB_CHECK_ARG(y && stats && gamma && beta && w2 && b2 && w3 && b3 && out && out_size, "pcb_mlp_fwd: null argument");
  PCB_CHECK_ARG(mode >= PCB_DW_SAME && mode <= PCB_DW_UP, "pcb_mlp_fwd: bad mode %d", mode);
  PCB_CHECK_ARG(C % 16 == 0 && H % 16 == 0 && Co % 16 == 0 && C > 0 && H > 0 && Co > 0,
                "pcb_mlp_fwd: channel counts must be multiples of 16 (C=%lld H=%lld Co=%lld)", (long long)C, (long long)H, (long long)Co);
  PCB_CHECK_ARG((wr == nullptr) || (xs && br && xs_size && Cr > 0 && Cr % 16 == 0 && mode != PCB_DW_SAME),
                "pcb_mlp_fwd: bad res-conv arguments");
  PCB_CHECK_ARG(N > 0 && N <= 65535, "pcb_mlp_fwd: bad batch");
  MlpArgs a;
  a.y = (const uint4*)y; a.stats = stats; a.gamma = gamma; a.beta = beta; a.w2 = (const uint4*)w2; a.b2 = b2;
  a.w3 = (const uint4*)w3; a.b3 = b3; a.res = (const uint4*)res; a.xs = (const uint4*)xs; a.wr = (const uint4*)wr;
  a.br = wr ? br : nullptr; a.out = (uint4*)out;
  a.o0 = (int)out_size[0]; a.o1 = (int)out_size[1]; a.o2 = (int)out_size[2];
  a.C = (int)C; a.H = (int)H; a.Co = (int)Co; a.Cr = wr ? (int)Cr : 0; a.mode = mode;
  a.Vout = (int64_t)a.o0 * a.o1 * a.o2;
  a.x0 = a.x1 = a.x2 = 0;
  if (xs_size) { a.x0 = (int)xs_size[0]; a.x1 = (int)xs_size[1]; a.x2 = (int)xs_size[2]; }
  if (mode == PCB_DW_UP) {
    PCB_CHECK_ARG(a.o0 >= 2 && a.o1 >= 2 && a.o2 >= 2, "pcb_mlp_fwd: UP output too small");
    a.Vy = (int64_t)(a.o0 - 1) * (a.o1 - 1) * (a.o2 - 1);
    if (wr) PCB_CHECK_ARG(a.x0 * 2 == a.o0 && a.x1 * 2 == a.o1 && a.x2 * 2 == a.o2, "pcb_mlp_fwd: UP needs out_size == 2*xs_size");
  } else {
    a.Vy = a.Vout;
    if (wr) PCB_CHECK_ARG((a.x0 - 1) / 2 + 1 == a.o0 && (a.x1 - 1) / 2 + 1 == a.o1 && (a.x2 - 1) / 2 + 1 == a.o2,
                          "pcb_mlp_fwd: DOWN needs out_size == (xs_size-1)/2+1");
  }
  a.Vin = (int64_t)a.x0 * a.x1 * a.x2;
  a.KC = pick_chunk(C, 128);
  a.N1 = pick_chunk(H, 128);
  a.CoT = Co <= 256 ? (int)Co : 256;
  PCB_CHECK_ARG(Co % a.CoT == 0, "pcb_mlp_fwd: Co=%lld must be <=256 or a multiple of 256", (long long)Co);
  a.KCr = wr ? pick_chunk(Cr, 128) : 0;
  a.inv_count = (float)(1.0 / (double)a.Vy);
  const int KA = a.KC > a.KCr ? a.KC : a.KCr, KW3 = a.N1 > a.KCr ? a.N1 : a.KCr;
  const size_t smem = (size_t)128 * KA * 2 + (size_t)a.N1 * a.KC * 2 + (size_t)128 * a.N1 * 2 + (size_t)a.CoT * KW3 * 2 +
                      (size_t)2 * C * sizeof(float) + 256 * sizeof(int64_t) + 16;
  PCB_CHECK_ARG(smem <= 227 * 1024, "pcb_mlp_fwd: tile needs %zu B shared memory", smem);
  static size_t configured = 0;
  if (smem > configured) {
    if (cudaFuncSetAttribute(mlp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024)) != cudaSuccess) {
      set_error("pcb_mlp_fwd: cudaFuncSetAttribute failed: %s", cudaGetErrorString(cudaGetLastError()));
      return PCB_ERR_CUDA;
    }
    configured = 227 * 1024;
  }
  dim3 grid((unsigned)((a.Vout + 127) / 128), (unsigned)N, (unsigned)(Co / a.CoT));
  mlp_kernel<<<grid, 128, smem, (cudaStream_t)stream>>>(a);
  PCB_CHECK_LAUNCH("pcb_mlp_fwd");
  return PCB_OK;
}

extern "C" int pcb_head_fwd(const void* x, const float* w, const float* b, void* out, int out_dtype, int64_t N,
                            int64_t C, int64_t ncls, int64_t nvox, void* stream) {
  PCB_CHECK_ARG(x && w && b && out, "pcb_head_fwd: null argument");
  PCB_CHECK_ARG(C > 0 && C % 8 == 0 && ncls > 0 && N > 0 && nvox > 0, "pcb_head_fwd: bad shape");
  const size_t smem = (size_t)(C * ncls + ncls) * sizeof(float);
  PCB_CHECK_ARG(smem <= 48 * 1024, "pcb_head_fwd: C*ncls too large (%lld x %lld)", (long long)C, (long long)ncls);
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t total = N * nvox;
  const int grid = (int)((total + 255) / 256 > 148 * 16 ? 148 * 16 : (total + 255) / 256);
  if (out_dtype == PCB_F32) head_kernel<float><<<grid, 256, smem, st>>>((const uint4*)x, w, b, (float*)out, N, (int)C, (int)ncls, nvox);
  else if (out_dtype == PCB_F16) head_kernel<__half><<<grid, 256, smem, st>>>((const uint4*)x, w, b, (__half*)out, N, (int)C, (int)ncls, nvox);
  else if (out_dtype == PCB_BF16) head_kernel<__nv_bfloat16><<<grid, 256, smem, st>>>((const uint4*)x, w, b, (__nv_bfloat16*)out, N, (int)C, (int)ncls, nvox);
  else { set_error("pcb_head_fwd: bad dtype %d", out_dtype); return PCB_ERR_INVALID; }
  PCB_CHECK_LAUNCH("pcb_head_fwd");
  return PCB_OK;
}
