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
#include <stdlib.h>
#include <string.h>

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

// Streaming variant (C/8 a power of two, Cin <= 4): the thread's 8-channel chunk is fixed for the whole grid-stride
// loop, so its weight rows and bias sit in registers and the voxel index is a shift — one coalesced 128-bit store per
// item and no 64-bit divisions (the generic kernel above spends most of its instructions on index math and weight loads).
template <typename TIn, int CIN>
__global__ void __launch_bounds__(256) stem_stream_kernel(const TIn* __restrict__ x, const float* __restrict__ w,
                                                          const float* __restrict__ b, uint4* __restrict__ out,
                                                          int C, int64_t V, int ch_shift) {
  const int CH = C >> 3, n = blockIdx.y;
  const int64_t stride = (int64_t)gridDim.x * 256;          // multiple of CH
  const int64_t first = (int64_t)blockIdx.x * 256 + threadIdx.x;
  const int cc = (int)(first & (CH - 1));
  float wr[CIN][8], br[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    br[j] = __ldg(b + cc * 8 + j);
#pragma unroll
    for (int ci = 0; ci < CIN; ++ci) wr[ci][j] = __ldg(w + (cc * 8 + j) * CIN + ci);
  }
  const TIn* xn = x + (int64_t)n * CIN * V;
  uint4* on = out + (int64_t)n * V * CH;
  const int64_t items = V * CH;
  for (int64_t i = first; i < items; i += stride) {
    const int64_t v = i >> ch_shift;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = br[j];
#pragma unroll
    for (int ci = 0; ci < CIN; ++ci) {
      const float xv = (float)xn[(int64_t)ci * V + v];
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = fmaf(xv, wr[ci][j], acc[j]);
    }
    on[i] = pack8(acc);
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

// ---------------------------------------------------------------------------- transposed stride-2 stencil, k = 3
// ConvTranspose3d(k=3, stride 2, padding 1): o = 2 i - 1 + k per axis, so an EVEN output coordinate 2i has the single
// tap k=1 on input i and an ODD one 2i+1 has k=2 on input i and k=0 on input i+1.  One thread owns UP_XB input
// positions along W (2*UP_XB outputs) of one output row for 8 channels: it walks the 1, 2 or 4 contributing input
// rows, loads UP_XB+1 input vectors per row once and applies exactly the taps that exist (27/8 FMAs per output on
// average instead of 27 predicated tap tests).  Epilogue (bias, fused add, GN statistics) as dwconv_kernel.
template <int UP_XB>
__global__ void __launch_bounds__(256, 2) dwconv_up3_kernel(const uint4* __restrict__ x, const float* __restrict__ w,
                                                         const float* __restrict__ bias, uint4* __restrict__ y,
                                                         double* __restrict__ stats, const uint4* __restrict__ add,
                                                         DwArgs a) {
  extern __shared__ double s_stats[];  // [2*C]
  const int C = a.C, CH = C >> 3;
  const int n = blockIdx.y;
  if (stats != nullptr) {
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) s_stats[i] = 0.0;
    __syncthreads();
  }
  const int nstrip = (a.Wo + 2 * UP_XB - 1) / (2 * UP_XB);
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
    const int i0 = xs * UP_XB;
    float acc[2 * UP_XB][8];
#pragma unroll
    for (int j = 0; j < 2 * UP_XB; ++j)
#pragma unroll
      for (int c = 0; c < 8; ++c) acc[j][c] = 0.f;
    const uint4* xn = x + (int64_t)n * a.D * a.H * a.W * CH + cc;
    // contributing (input row, tap) pairs per axis
    // first pair: (o>>1, k = 1 for even o, 2 for odd o); second pair (odd o only): ((o+1)>>1, k = 0)
    const int iz0 = oz >> 1, kz0 = 1 + (oz & 1), iy0 = oy >> 1, ky0 = 1 + (oy & 1);
    const int nz = iz0 >= a.D ? 0 : (((oz & 1) && iz0 + 1 < a.D) ? 2 : 1);
    const int ny = iy0 >= a.H ? 0 : (((oy & 1) && iy0 + 1 < a.H) ? 2 : 1);
#pragma unroll 1
    for (int zt = 0; zt < nz; ++zt) {
      const int iz = iz0 + zt, kz = zt ? 0 : kz0;
#pragma unroll 1
      for (int yt = 0; yt < ny; ++yt) {
        const int iy = iy0 + yt, ky = yt ? 0 : ky0;
        const uint4* row = xn + ((int64_t)iz * a.H + iy) * a.W * CH;
        uint4 v[UP_XB + 1];
#pragma unroll
        for (int j = 0; j <= UP_XB; ++j)
          v[j] = (i0 + j < a.W) ? __ldg(row + (int64_t)(i0 + j) * CH) : make_uint4(0, 0, 0, 0);
        float wv[3][8];
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const float4* wp = reinterpret_cast<const float4*>(w + ((kz * 3 + ky) * 3 + kx) * C + cc * 8);
          const float4 w0 = __ldg(wp), w1 = __ldg(wp + 1);
          wv[kx][0] = w0.x; wv[kx][1] = w0.y; wv[kx][2] = w0.z; wv[kx][3] = w0.w;
          wv[kx][4] = w1.x; wv[kx][5] = w1.y; wv[kx][6] = w1.z; wv[kx][7] = w1.w;
        }
        float f0[8], f1[8];
        unpack8(v[0], f0);
#pragma unroll
        for (int j = 0; j < UP_XB; ++j) {
          unpack8(v[j + 1], f1);
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            acc[2 * j][c] = fmaf(f0[c], wv[1][c], acc[2 * j][c]);
            acc[2 * j + 1][c] = fmaf(f1[c], wv[0][c], fmaf(f0[c], wv[2][c], acc[2 * j + 1][c]));
          }
#pragma unroll
          for (int c = 0; c < 8; ++c) f0[c] = f1[c];
        }
      }
    }
    float bv[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (bias != nullptr) {
      const float4* bp = reinterpret_cast<const float4*>(bias + cc * 8);
      const float4 b0 = __ldg(bp), b1 = __ldg(bp + 1);
      bv[0] = b0.x; bv[1] = b0.y; bv[2] = b0.z; bv[3] = b0.w; bv[4] = b1.x; bv[5] = b1.y; bv[6] = b1.z; bv[7] = b1.w;
    }
    const int ox0 = 2 * i0;
    uint4* yrow = y + (((int64_t)n * a.Do + oz) * a.Ho + oy) * a.Wo * CH + cc;
#pragma unroll
    for (int j = 0; j < 2 * UP_XB; ++j) {
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

// ---------------------------------------------------------------------------- tiled transposed stride-2 stencil, k = 3
// Same arithmetic as dwconv_up3_kernel (o = 2 i - 1 + k: an even output has one tap per axis, an odd one two), but the
// coarse input brick of a CTA ((2+1) x (4+1) x (16+1) voxels x 32 channels for a 4 x 8 x 32 output tile) is staged in shared
// memory once and every thread item owns 8 consecutive outputs along W for 8 channels.  The untiled kernel gathers its 1-4
// input rows straight from L2 inside two dependent loops (round 2: 0.38 ms per 160^3 x 64 ch sample = 1.35 TB/s of writes,
// latency-bound at 16 warps per SM); here the gathers are shared-memory reads and the kernel streams its output.
constexpr int UT_Z = 2, UT_Y = 4, UT_X = 16;                       // coarse tile (outputs: 4 x 8 x 32)
constexpr int UT_BZ = UT_Z + 1, UT_BY = UT_Y + 1, UT_BX = UT_X + 1;

__global__ void __launch_bounds__(256, 2) dwconv_up3_tiled_kernel(const uint4* __restrict__ x, const float* __restrict__ w,
                                                               const float* __restrict__ bias, uint4* __restrict__ y,
                                                               double* __restrict__ stats, const uint4* __restrict__ add,
                                                               DwArgs a, int tiles_y, int tiles_x) {
  __shared__ __align__(16) uint4 s_in[UT_BZ * UT_BY * UT_BX * 4];   // [bz][by][bx][4 chunks of 8 channels]
  __shared__ __align__(16) float s_w[27 * 32];                      // taps of this 32-channel group
  __shared__ double s_stats[64];                                    // [2][32]
  const int C = a.C, CH = C >> 3, tid = threadIdx.x;
  const int ncg = C >> 5, cg = blockIdx.x % ncg;
  int t = blockIdx.x / ncg;
  const int n = blockIdx.y;
  const int tx = t % tiles_x; t /= tiles_x;
  const int ty = t % tiles_y, tz = t / tiles_y;
  const int z0 = tz * UT_Z, y0 = ty * UT_Y, x0 = tx * UT_X;        // coarse origin of the tile
  for (int i = tid; i < 27 * 32; i += 256) s_w[i] = w[(i >> 5) * C + cg * 32 + (i & 31)];
  if (tid < 64) s_stats[tid] = 0.0;
  const uint4* xn = x + (int64_t)n * a.D * a.H * a.W * CH + cg * 4;
  staged_copy<4>(UT_BZ * UT_BY * UT_BX * 4, tid, 256,
      [&](int q) {
        const int c4 = q & 3, v = q >> 2;
        const int bx = v % UT_BX, by = (v / UT_BX) % UT_BY, bz = v / (UT_BX * UT_BY);
        const int gz = z0 + bz, gy = y0 + by, gx = x0 + bx;
        if (gz >= a.D || gy >= a.H || gx >= a.W) return make_uint4(0, 0, 0, 0);
        return __ldg(xn + (((int64_t)gz * a.H + gy) * a.W + gx) * CH + c4);
      },
      [&](int q, const uint4& v4) { s_in[q] = v4; });
  __syncthreads();
  float ssum[8], ssq[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { ssum[j] = 0.f; ssq[j] = 0.f; }
  const int cc = tid & 3;                                           // 8-channel chunk of the group (fixed per thread)
#pragma unroll 1
  for (int item = tid; item < 4 * 8 * 4 * 4; item += 256) {
    const int xseg = (item >> 2) & 3, row = item >> 4;              // 4 segments of 8 outputs; 32 output rows (loz, loy)
    // rows are enumerated by (z, y) parity class so that the two rows of a warp run the same 1 / 2 / 4 tap-row loop (no
    // divergence), and a thread's two items (row, row + 16) pair the classes (even,even)+(odd,odd) / (even,odd)+(odd,even):
    // 1 + 4 and 2 + 2 tap rows — balanced across the CTA's warps
    const int cls = row >> 3, kk = row & 7;
    const int pz = cls >> 1, py = (cls ^ (cls >> 1)) & 1;           // class order 00, 01, 11, 10
    const int loz = 2 * (kk >> 2) + pz, loy = 2 * (kk & 3) + py;
    const int oz = 2 * z0 + loz, oy = 2 * y0 + loy, ox0 = 2 * x0 + 8 * xseg;
    if (oz >= a.Do || oy >= a.Ho || ox0 >= a.Wo) continue;
    uint64_t acc[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[j][c] = 0ull;
    // contributing (input row, tap) pairs per axis: (o >> 1, k = 1 + (o & 1)) and, for odd o, ((o >> 1) + 1, k = 0)
    const int lz0 = loz >> 1, kz0 = 1 + (loz & 1), ly0 = loy >> 1, ky0 = 1 + (loy & 1);
    const int nz = 1 + (loz & 1), ny = 1 + (loy & 1);               // rows beyond the tensor are zero in the brick
    const int li0 = 4 * xseg;                                       // first coarse x of this segment inside the brick
#pragma unroll 1
    for (int zt = 0; zt < nz; ++zt) {
      const int lz = lz0 + zt, kz = zt ? 0 : kz0;
#pragma unroll 1
      for (int yt = 0; yt < ny; ++yt) {
        const int ly = ly0 + yt, ky = yt ? 0 : ky0;
        const uint4* rowp = s_in + ((lz * UT_BY + ly) * UT_BX + li0) * 4 + cc;
        const float* wp = s_w + ((kz * 3 + ky) * 3) * 32 + cc * 8;
        uint64_t w0[4], w1[4], w2[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          w0[c] = pk2(wp[2 * c], wp[2 * c + 1]);
          w1[c] = pk2(wp[32 + 2 * c], wp[32 + 2 * c + 1]);
          w2[c] = pk2(wp[64 + 2 * c], wp[64 + 2 * c + 1]);
        }
        uint4 v4 = rowp[0];
        uint64_t f0[4], f1[4];
        f0[0] = pk2(bf16_lo(v4.x), bf16_hi(v4.x)); f0[1] = pk2(bf16_lo(v4.y), bf16_hi(v4.y));
        f0[2] = pk2(bf16_lo(v4.z), bf16_hi(v4.z)); f0[3] = pk2(bf16_lo(v4.w), bf16_hi(v4.w));
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          v4 = rowp[(j + 1) * 4];
          f1[0] = pk2(bf16_lo(v4.x), bf16_hi(v4.x)); f1[1] = pk2(bf16_lo(v4.y), bf16_hi(v4.y));
          f1[2] = pk2(bf16_lo(v4.z), bf16_hi(v4.z)); f1[3] = pk2(bf16_lo(v4.w), bf16_hi(v4.w));
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            acc[2 * j][c] = fma2(f0[c], w1[c], acc[2 * j][c]);
            acc[2 * j + 1][c] = fma2(f1[c], w0[c], fma2(f0[c], w2[c], acc[2 * j + 1][c]));
            f0[c] = f1[c];
          }
        }
      }
    }
    float bv[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (bias != nullptr) {
      const float4* bp = reinterpret_cast<const float4*>(bias + cg * 32 + cc * 8);
      const float4 b0 = __ldg(bp), b1 = __ldg(bp + 1);
      bv[0] = b0.x; bv[1] = b0.y; bv[2] = b0.z; bv[3] = b0.w; bv[4] = b1.x; bv[5] = b1.y; bv[6] = b1.z; bv[7] = b1.w;
    }
    uint4* yrow = y + (((int64_t)n * a.Do + oz) * a.Ho + oy) * a.Wo * CH + cg * 4 + cc;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (ox0 + j >= a.Wo) continue;
      float av[8];
#pragma unroll
      for (int c = 0; c < 4; ++c) upk2(acc[j][c], av[2 * c], av[2 * c + 1]);
      if (a.add_mode == 1) {
        float f[8];
        unpack8(__ldg(add + ((((int64_t)n * a.Do + oz) * a.Ho + oy) * a.Wo + ox0 + j) * CH + cg * 4 + cc), f);
#pragma unroll
        for (int c = 0; c < 8; ++c) av[c] += f[c];
      } else if (a.add_mode == 2) {
        if (!((oz | oy | (ox0 + j)) & 1)) {
          float f[8];
          unpack8(__ldg(add + ((((int64_t)n * ((a.Do + 1) >> 1) + (oz >> 1)) * a.a1 + (oy >> 1)) * a.a2 + ((ox0 + j) >> 1)) * CH + cg * 4 + cc), f);
#pragma unroll
          for (int c = 0; c < 8; ++c) av[c] += f[c];
        }
      }
      float o[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        o[c] = round_bf16(av[c] + bv[c]);
        ssum[c] += o[c];
        ssq[c] = fmaf(o[c], o[c], ssq[c]);
      }
      yrow[(int64_t)(ox0 + j) * CH] = pack8(o);
    }
  }
  if (stats == nullptr) return;   // uniform across the grid (backward-data use)
  // lanes with the same chunk (lane & 3) hold the same channels
#pragma unroll
  for (int off = 16; off >= 4; off >>= 1)
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      ssum[c] += __shfl_xor_sync(0xffffffffu, ssum[c], off);
      ssq[c] += __shfl_xor_sync(0xffffffffu, ssq[c], off);
    }
  if ((tid & 31) < 4) {
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      atomicAdd(&s_stats[cc * 8 + c], (double)ssum[c]);
      atomicAdd(&s_stats[32 + cc * 8 + c], (double)ssq[c]);
    }
  }
  __syncthreads();
  if (tid < 64) atomicAdd(&stats[(int64_t)n * 2 * C + (tid >> 5) * C + cg * 32 + (tid & 31)], s_stats[tid]);
}

static bool launch_dw_up3_tiled(cudaStream_t st, const uint4* x, const float* w, const float* b, uint4* y, double* stats,
                                const uint4* add, const DwArgs& a, int64_t N) {
  static const bool off = getenv("PCB_NO_UP3_TILED") != nullptr;
  if (off || a.Do > 2 * a.D || a.Ho > 2 * a.H || a.Wo > 2 * a.W) return false;
  const int tz = (a.Do + 2 * UT_Z - 1) / (2 * UT_Z), ty = (a.Ho + 2 * UT_Y - 1) / (2 * UT_Y), tx = (a.Wo + 2 * UT_X - 1) / (2 * UT_X);
  const int64_t nb = (int64_t)tz * ty * tx * (a.C / 32);
  if (nb >= (1ll << 31) || N > 65535) return false;
  dwconv_up3_tiled_kernel<<<dim3((unsigned)nb, (unsigned)N), 256, 0, st>>>(x, w, b, y, stats, add, a, ty, tx);
  return true;
}

// ---------------------------------------------------------------------------- tiled SAME-mode stencil
// One CTA = one 4x8x16 output brick x 32 channels.  The (4+2P)x(8+2P)x(16+2P) input brick is staged in
// shared memory with batched, fully coalesced 128-bit loads (zero fill outside the volume); every thread
// owns 8 consecutive outputs along W for 8 channels (64 fp32 accumulators, packed FFMA2), so each staged
// input column is read once per (dz,dy) row and reused by up to K taps x 8 outputs.  Row pitch is padded by
// one voxel (64 B) so a quarter-warp's LDS.128 touches 8 distinct 16 B bank groups.
constexpr int DT_Z = 4, DT_Y = 8, DT_X = 16, DT_XB = 8;

// One 4x8x16 output brick x 32 channels from the staged input brick: every thread owns 8 consecutive outputs along W
// for 8 channels (64 fp32 accumulators as 32 packed pairs); adds bias (+ fused `add`), rounds to bf16, stores, and
// returns the thread's partial GroupNorm sums of the rounded outputs.
template <int K>
__device__ __forceinline__ void dw_same_brick(const uint4* __restrict__ s_in, const float* __restrict__ s_w,
                                              const float* __restrict__ bias, const uint4* __restrict__ add,
                                              uint4* __restrict__ y, const DwArgs& a, int tid, int n, int cg, int z0, int y0,
                                              int x0, float* ssum, float* ssq) {
  constexpr int P = K / 2;
  constexpr int BY = DT_Y + 2 * P, BX = DT_X + 2 * P, PITCH = BX + 1;
  const int CH = a.C >> 3;
  const int cc = tid & 3, ly = (tid >> 2) & 7, xb = (tid >> 5) & 1, lz = tid >> 6;
  uint64_t acc[DT_XB][4];
#pragma unroll
  for (int j = 0; j < DT_XB; ++j)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[j][c] = 0ull;
#pragma unroll 1
  for (int dz = 0; dz < K; ++dz) {
#pragma unroll 1
    for (int dy = 0; dy < K; ++dy) {
      uint64_t wv[K][4];
#pragma unroll
      for (int dx = 0; dx < K; ++dx) {
        const float4* wp = reinterpret_cast<const float4*>(s_w + ((dz * K + dy) * K + dx) * 32 + cc * 8);
        const float4 w0 = wp[0], w1 = wp[1];
        wv[dx][0] = pk2(w0.x, w0.y); wv[dx][1] = pk2(w0.z, w0.w); wv[dx][2] = pk2(w1.x, w1.y); wv[dx][3] = pk2(w1.z, w1.w);
      }
      const uint4* row = s_in + (((lz + dz) * BY + (ly + dy)) * PITCH + xb * DT_XB) * 4 + cc;
#pragma unroll
      for (int i = 0; i < DT_XB + K - 1; ++i) {
        const uint4 v4 = row[i * 4];
        uint64_t f[4];
        f[0] = pk2(bf16_lo(v4.x), bf16_hi(v4.x)); f[1] = pk2(bf16_lo(v4.y), bf16_hi(v4.y));
        f[2] = pk2(bf16_lo(v4.z), bf16_hi(v4.z)); f[3] = pk2(bf16_lo(v4.w), bf16_hi(v4.w));
#pragma unroll
        for (int dx = 0; dx < K; ++dx) {
          const int j = i - dx;
          if (j >= 0 && j < DT_XB) {
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[j][c] = fma2(f[c], wv[dx][c], acc[j][c]);
          }
        }
      }
    }
  }
  float bv[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (bias != nullptr) {
    const float4* bp = reinterpret_cast<const float4*>(bias + cg * 32 + cc * 8);
    const float4 b0 = __ldg(bp), b1 = __ldg(bp + 1);
    bv[0] = b0.x; bv[1] = b0.y; bv[2] = b0.z; bv[3] = b0.w; bv[4] = b1.x; bv[5] = b1.y; bv[6] = b1.z; bv[7] = b1.w;
  }
#pragma unroll
  for (int c = 0; c < 8; ++c) { ssum[c] = 0.f; ssq[c] = 0.f; }
  const int oz = z0 + lz, oy = y0 + ly;
  if (oz < a.D && oy < a.H) {
    const int64_t rowoff = ((((int64_t)n * a.D + oz) * a.H + oy) * a.W) * CH + cg * 4 + cc;
    uint4 addv[DT_XB];
    if (add != nullptr) {
#pragma unroll
      for (int j = 0; j < DT_XB; ++j) {
        const int ox = x0 + xb * DT_XB + j;
        addv[j] = ox < a.W ? __ldg(add + rowoff + (int64_t)ox * CH) : make_uint4(0, 0, 0, 0);
      }
    }
#pragma unroll
    for (int j = 0; j < DT_XB; ++j) {
      const int ox = x0 + xb * DT_XB + j;
      if (ox >= a.W) continue;
      float o[8];
#pragma unroll
      for (int c = 0; c < 4; ++c) upk2(acc[j][c], o[2 * c], o[2 * c + 1]);
      if (add != nullptr) {
        float f[8];
        unpack8(addv[j], f);
#pragma unroll
        for (int c = 0; c < 8; ++c) o[c] += f[c];
      }
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        o[c] = round_bf16(o[c] + bv[c]);
        ssum[c] += o[c];
        ssq[c] = fmaf(o[c], o[c], ssq[c]);
      }
      y[rowoff + (int64_t)ox * CH] = pack8(o);
    }
  }
}

template <int K>
__global__ void __launch_bounds__(256, 2) dwconv_same_tiled_kernel(const uint4* __restrict__ x, const float* __restrict__ w,
                                                                const float* __restrict__ bias, uint4* __restrict__ y,
                                                                double* __restrict__ stats, const uint4* __restrict__ add,
                                                                DwArgs a, int tiles_y, int tiles_x,
                                                                const __grid_constant__ CUtensorMap tmap, int use_tma) {
  constexpr int P = K / 2;
  constexpr int BZ = DT_Z + 2 * P, BY = DT_Y + 2 * P, BX = DT_X + 2 * P, PITCH = BX + 1;   // voxels
  extern __shared__ __align__(128) uint8_t dsm[];
  uint4* s_in = reinterpret_cast<uint4*>(dsm);                       // [BZ][BY][PITCH][4 chunks]
  float* s_w = reinterpret_cast<float*>(s_in + BZ * BY * PITCH * 4); // [K^3][32]
  double* s_stats = reinterpret_cast<double*>(s_w + K * K * K * 32); // [64]
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_stats + 64);       // TMA completion barrier
  const int tid = threadIdx.x;
  const int CH = a.C >> 3;
  const int ncg = a.C >> 5;
  const int cg = blockIdx.x % ncg;           // 32-channel group, FASTEST in launch order: the groups of one brick read
  const int n = blockIdx.z;                  // the two/four 64 B halves of the same 128 B lines while they are hot in L2
  int t = blockIdx.x / ncg;
  const int tx = t % tiles_x; t /= tiles_x;
  const int ty = t % tiles_y;
  const int tz = t / tiles_y;
  const int z0 = tz * DT_Z, y0 = ty * DT_Y, x0 = tx * DT_X;

  if (use_tma) {
    // the whole input brick (halo included, zero-filled outside the volume) is one bulk tensor copy: box
    // [1][BZ][BY][PITCH][32 ch] of the [N][D][H][W][C] tensor at (n, z0-P, y0-P, x0-P, cg*32)
    if (tid == 0) {
      mbar_init(s_bar, 1);
      fence_mbar_init();
      fence_proxy_async_smem();
      mbar_arrive_expect_tx(s_bar, (uint32_t)(BZ * BY * PITCH * 64));
      tma_load_5d(s_in, &tmap, cg * 32, x0 - P, y0 - P, z0 - P, n, s_bar);
    }
  }
  for (int i = tid; i < K * K * K * 32; i += 256) s_w[i] = w[(i >> 5) * a.C + cg * 32 + (i & 31)];
  if (tid < 64) s_stats[tid] = 0.0;
  const uint4* xn = x + (int64_t)n * a.D * a.H * a.W * CH + cg * 4;
  if (use_tma) {
    __syncthreads();            // barrier init visible to every waiter
    mbar_wait(s_bar, 0);
  } else
  staged_copy<((BZ * BY * BX * 4 + 255) / 256 <= 17 ? (BZ * BY * BX * 4 + 255) / 256 : 8)>(BZ * BY * BX * 4, tid, 256,   // whole brick in flight at once
      [&](int q) {
        const int cc = q & 3, v = q >> 2;
        const int bx = v % BX, by = (v / BX) % BY, bz = v / (BX * BY);
        const int gz = z0 + bz - P, gy = y0 + by - P, gx = x0 + bx - P;
        if (gz < 0 || gz >= a.D || gy < 0 || gy >= a.H || gx < 0 || gx >= a.W) return make_uint4(0, 0, 0, 0);
        return __ldg(xn + (((int64_t)gz * a.H + gy) * a.W + gx) * CH + cc);
      },
      [&](int q, const uint4& v4) {
        const int cc = q & 3, v = q >> 2;
        const int bx = v % BX, by = (v / BX) % BY, bz = v / (BX * BY);
        s_in[((bz * BY + by) * PITCH + bx) * 4 + cc] = v4;
      });
  __syncthreads();

  const int cc = tid & 3;
  float ssum[8], ssq[8];
  dw_same_brick<K>(s_in, s_w, bias, add, y, a, tid, n, cg, z0, y0, x0, ssum, ssq);
  if (stats == nullptr) return;
  // lanes sharing a channel chunk differ in bits 2..4 (ly low bits) -> butterfly, then shared/global f64 atomics
#pragma unroll
  for (int off = 4; off <= 16; off <<= 1) {
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      ssum[c] += __shfl_xor_sync(0xffffffffu, ssum[c], off);
      ssq[c] += __shfl_xor_sync(0xffffffffu, ssq[c], off);
    }
  }
  if ((tid & 31) < 4) {
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      atomicAdd(&s_stats[cc * 8 + c], (double)ssum[c]);
      atomicAdd(&s_stats[32 + cc * 8 + c], (double)ssq[c]);
    }
  }
  __syncthreads();
  if (tid < 64) {
    const int which = tid >> 5, c = tid & 31;
    atomicAdd(&stats[(int64_t)n * 2 * a.C + which * a.C + cg * 32 + c], s_stats[tid]);
  }
}

// Persistent variant (TMA only): one CTA per SM loops over the bricks of one (sample, 32-channel group) with TWO input
// stages — the bulk tensor copy of brick i+1 is in flight while brick i is computed, weights are staged once, and the
// GroupNorm partial sums leave the CTA once (64 f64 atomics per CTA instead of per brick).
template <int K>
__global__ void __launch_bounds__(256, 1) dwconv_same_persist_kernel(const float* __restrict__ w, const float* __restrict__ bias,
                                                                     uint4* __restrict__ y, double* __restrict__ stats,
                                                                     const uint4* __restrict__ add, DwArgs a, int tiles_y,
                                                                     int tiles_x, int nbricks,
                                                                     const __grid_constant__ CUtensorMap tmap) {
  constexpr int P = K / 2;
  constexpr int BZ = DT_Z + 2 * P, BY = DT_Y + 2 * P, BX = DT_X + 2 * P, PITCH = BX + 1;   // voxels
  constexpr int STAGE = BZ * BY * PITCH * 4;                                                // uint4 per stage
  extern __shared__ __align__(128) uint8_t dsm[];
  uint4* s_in = reinterpret_cast<uint4*>(dsm);                       // [2][BZ][BY][PITCH][4 chunks]
  float* s_w = reinterpret_cast<float*>(s_in + 2 * STAGE);           // [K^3][32]
  double* s_stats = reinterpret_cast<double*>(s_w + K * K * K * 32); // [64]
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_stats + 64);       // [2] TMA completion barriers
  const int tid = threadIdx.x;
  const int cg = blockIdx.y, n = blockIdx.z;
  auto issue = [&](int b, int stage) {
    int t = b;
    const int tx = t % tiles_x; t /= tiles_x;
    const int ty = t % tiles_y, tz = t / tiles_y;
    mbar_arrive_expect_tx(&s_bar[stage], (uint32_t)(STAGE * 16));
    tma_load_5d(s_in + stage * STAGE, &tmap, cg * 32, tx * DT_X - P, ty * DT_Y - P, tz * DT_Z - P, n, &s_bar[stage]);
  };
  if (tid == 0) {
    mbar_init(&s_bar[0], 1);
    mbar_init(&s_bar[1], 1);
    fence_mbar_init();
    fence_proxy_async_smem();
    if ((int)blockIdx.x < nbricks) issue(blockIdx.x, 0);
  }
  for (int i = tid; i < K * K * K * 32; i += 256) s_w[i] = w[(i >> 5) * a.C + cg * 32 + (i & 31)];
  if (tid < 64) s_stats[tid] = 0.0;
  __syncthreads();
  const int cc = tid & 3;
  uint32_t phase0 = 0, phase1 = 0;
  int stage = 0;
  for (int b = blockIdx.x; b < nbricks; b += gridDim.x) {
    // stage^1 was fully consumed before the __syncthreads that ended the previous iteration
    if (tid == 0 && b + (int)gridDim.x < nbricks) issue(b + gridDim.x, stage ^ 1);
    if (stage == 0) { mbar_wait(&s_bar[0], phase0); phase0 ^= 1; }
    else { mbar_wait(&s_bar[1], phase1); phase1 ^= 1; }
    int t = b;
    const int tx = t % tiles_x; t /= tiles_x;
    const int ty = t % tiles_y, tz = t / tiles_y;
    float ssum[8], ssq[8];
    dw_same_brick<K>(s_in + stage * STAGE, s_w, bias, add, y, a, tid, n, cg, tz * DT_Z, ty * DT_Y, tx * DT_X, ssum, ssq);
    if (stats != nullptr) {
#pragma unroll
      for (int off = 4; off <= 16; off <<= 1) {
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          ssum[c] += __shfl_xor_sync(0xffffffffu, ssum[c], off);
          ssq[c] += __shfl_xor_sync(0xffffffffu, ssq[c], off);
        }
      }
      if ((tid & 31) < 4) {
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          atomicAdd(&s_stats[cc * 8 + c], (double)ssum[c]);
          atomicAdd(&s_stats[32 + cc * 8 + c], (double)ssq[c]);
        }
      }
    }
    __syncthreads();
    stage ^= 1;
  }
  if (stats != nullptr && tid < 64) {
    const int which = tid >> 5, c = tid & 31;
    atomicAdd(&stats[(int64_t)n * 2 * a.C + which * a.C + cg * 32 + c], s_stats[tid]);
  }
}

// ---------------------------------------------------------------------------- tiled stride-2 stencil, k = 3
// Conv3d(k=3, stride 2, padding 1, groups=C) (MedNeXtDownBlock.conv1; also the data gradient of the transposed conv of
// MedNeXtUpBlock).  One CTA = a 2x4x16 output tile x 32 channels; its 5x9x33 input brick (origin 2*o0-1, zero-filled
// outside the volume) arrives as ONE bulk tensor copy, so 95 KB per CTA are in flight instead of a few dependent
// 128-bit loads per thread (the untiled kernel is latency-bound: ncu long_scoreboard 12 warps/issue, 14 % of DRAM peak).
// Thread = (8-channel chunk, ox, oy) and both oz of the tile: 45 LDS.128 feed 54 taps x 8 channels.
constexpr int DN_Z = 2, DN_Y = 4, DN_X = 16;
constexpr int DN_BZ = 2 * DN_Z + 1, DN_BY = 2 * DN_Y + 1, DN_BX = 2 * DN_X + 1;

__global__ void __launch_bounds__(256, 2) dwconv_down3_tiled_kernel(const float* __restrict__ w, const float* __restrict__ bias,
                                                                    uint4* __restrict__ y, double* __restrict__ stats,
                                                                    const uint4* __restrict__ add, DwArgs a, int tiles_y,
                                                                    int tiles_x, const __grid_constant__ CUtensorMap tmap) {
  extern __shared__ __align__(128) uint8_t dsm[];
  uint4* s_in = reinterpret_cast<uint4*>(dsm);                               // [BZ][BY][BX][4 chunks]
  float* s_w = reinterpret_cast<float*>(s_in + DN_BZ * DN_BY * DN_BX * 4);   // [27][32]
  double* s_stats = reinterpret_cast<double*>(s_w + 27 * 32);                // [64]
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_stats + 64);
  const int tid = threadIdx.x, CH = a.C >> 3;
  const int ncg = a.C >> 5;
  const int cg = blockIdx.x % ncg, n = blockIdx.z;      // channel group fastest (see dwconv_same_tiled_kernel)
  int t = blockIdx.x / ncg;
  const int tx = t % tiles_x; t /= tiles_x;
  const int ty = t % tiles_y, tz = t / tiles_y;
  const int z0 = tz * DN_Z, y0 = ty * DN_Y, x0 = tx * DN_X;
  if (tid == 0) {
    mbar_init(s_bar, 1);
    fence_mbar_init();
    fence_proxy_async_smem();
    mbar_arrive_expect_tx(s_bar, (uint32_t)(DN_BZ * DN_BY * DN_BX * 64));
    tma_load_5d(s_in, &tmap, cg * 32, 2 * x0 - 1, 2 * y0 - 1, 2 * z0 - 1, n, s_bar);
  }
  for (int i = tid; i < 27 * 32; i += 256) s_w[i] = w[(i >> 5) * a.C + cg * 32 + (i & 31)];
  if (tid < 64) s_stats[tid] = 0.0;
  __syncthreads();
  mbar_wait(s_bar, 0);

  const int cc = tid & 3, lx = (tid >> 2) & 15, ly = tid >> 6;
  uint64_t acc[DN_Z][4];
#pragma unroll
  for (int z = 0; z < DN_Z; ++z)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[z][c] = 0ull;
#pragma unroll 1
  for (int pz = 0; pz < DN_BZ; ++pz) {        // input plane pz serves output lz with tap dz = pz - 2 lz in [0, 3)
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
      const uint4* row = s_in + ((pz * DN_BY + 2 * ly + dy) * DN_BX + 2 * lx) * 4 + cc;
      uint64_t f[3][4];
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        const uint4 v4 = row[dx * 4];
        f[dx][0] = pk2(bf16_lo(v4.x), bf16_hi(v4.x)); f[dx][1] = pk2(bf16_lo(v4.y), bf16_hi(v4.y));
        f[dx][2] = pk2(bf16_lo(v4.z), bf16_hi(v4.z)); f[dx][3] = pk2(bf16_lo(v4.w), bf16_hi(v4.w));
      }
#pragma unroll
      for (int lz = 0; lz < DN_Z; ++lz) {
        const int dz = pz - 2 * lz;
        if (dz < 0 || dz > 2) continue;
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
          const float4* wp = reinterpret_cast<const float4*>(s_w + ((dz * 3 + dy) * 3 + dx) * 32 + cc * 8);
          const float4 w0 = wp[0], w1 = wp[1];
          acc[lz][0] = fma2(f[dx][0], pk2(w0.x, w0.y), acc[lz][0]);
          acc[lz][1] = fma2(f[dx][1], pk2(w0.z, w0.w), acc[lz][1]);
          acc[lz][2] = fma2(f[dx][2], pk2(w1.x, w1.y), acc[lz][2]);
          acc[lz][3] = fma2(f[dx][3], pk2(w1.z, w1.w), acc[lz][3]);
        }
      }
    }
  }
  float bv[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (bias != nullptr) {
    const float4* bp = reinterpret_cast<const float4*>(bias + cg * 32 + cc * 8);
    const float4 b0 = __ldg(bp), b1 = __ldg(bp + 1);
    bv[0] = b0.x; bv[1] = b0.y; bv[2] = b0.z; bv[3] = b0.w; bv[4] = b1.x; bv[5] = b1.y; bv[6] = b1.z; bv[7] = b1.w;
  }
  float ssum[8], ssq[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) { ssum[c] = 0.f; ssq[c] = 0.f; }
  const int oy = y0 + ly, ox = x0 + lx;
#pragma unroll
  for (int lz = 0; lz < DN_Z; ++lz) {
    const int oz = z0 + lz;
    if (oz >= a.Do || oy >= a.Ho || ox >= a.Wo) continue;
    const int64_t off = ((((int64_t)n * a.Do + oz) * a.Ho + oy) * a.Wo + ox) * CH + cg * 4 + cc;
    float o[8];
#pragma unroll
    for (int c = 0; c < 4; ++c) upk2(acc[lz][c], o[2 * c], o[2 * c + 1]);
    if (add != nullptr) {
      float f[8];
      unpack8(__ldg(add + off), f);
#pragma unroll
      for (int c = 0; c < 8; ++c) o[c] += f[c];
    }
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      o[c] = round_bf16(o[c] + bv[c]);
      ssum[c] += o[c];
      ssq[c] = fmaf(o[c], o[c], ssq[c]);
    }
    y[off] = pack8(o);
  }
  if (stats == nullptr) return;
#pragma unroll
  for (int off = 4; off <= 16; off <<= 1) {
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      ssum[c] += __shfl_xor_sync(0xffffffffu, ssum[c], off);
      ssq[c] += __shfl_xor_sync(0xffffffffu, ssq[c], off);
    }
  }
  if ((tid & 31) < 4) {
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      atomicAdd(&s_stats[cc * 8 + c], (double)ssum[c]);
      atomicAdd(&s_stats[32 + cc * 8 + c], (double)ssq[c]);
    }
  }
  __syncthreads();
  if (tid < 64) {
    const int which = tid >> 5, c = tid & 31;
    atomicAdd(&stats[(int64_t)n * 2 * a.C + which * a.C + cg * 32 + c], s_stats[tid]);
  }
}

static bool launch_dw_down3_tiled(cudaStream_t st, const uint4* x, const float* w, const float* b, uint4* y, double* stats,
                                  const uint4* add, DwArgs a, int64_t N) {
  static const bool off = getenv("PCB_NO_DOWN3") != nullptr || getenv("PCB_NO_TMA") != nullptr;
  if (off) return false;
  const size_t smem = (size_t)DN_BZ * DN_BY * DN_BX * 64 + 27 * 32 * 4 + 64 * 8 + 16;
  CUtensorMap tmap;
  memset(&tmap, 0, sizeof(tmap));
  if (!make_brick_tensor_map(&tmap, x, N, a.D, a.H, a.W, a.C, DN_BZ, DN_BY, DN_BX)) return false;
  static DevFlag configured;
  if (!configured) {
    cudaFuncSetAttribute(dwconv_down3_tiled_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (cudaFuncSetAttribute(dwconv_down3_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) {
      cudaGetLastError();
      return false;
    }
    configured = true;
  }
  const int tz = (a.Do + DN_Z - 1) / DN_Z, ty = (a.Ho + DN_Y - 1) / DN_Y, tx = (a.Wo + DN_X - 1) / DN_X;
  if ((int64_t)tz * ty * tx * (a.C / 32) >= (1ll << 31) || N > 65535) return false;
  dim3 grid((unsigned)(tz * ty * tx * (a.C / 32)), 1u, (unsigned)N);
  dwconv_down3_tiled_kernel<<<grid, 256, smem, st>>>(w, b, y, stats, add, a, ty, tx, tmap);
  return true;
}

template <int K>
static bool launch_dw_tiled(cudaStream_t st, const uint4* x, const float* w, const float* b, uint4* y, double* stats,
                            const uint4* add, DwArgs a, int64_t N) {
  constexpr int P = K / 2;
  const size_t smem = (size_t)(DT_Z + 2 * P) * (DT_Y + 2 * P) * (DT_X + 2 * P + 1) * 64 + (size_t)K * K * K * 32 * 4 + 64 * 8 + 16;
  if (smem > 227 * 1024) return false;
  CUtensorMap tmap;
  memset(&tmap, 0, sizeof(tmap));
  static const bool no_tma = getenv("PCB_NO_TMA") != nullptr;
  const int use_tma = !no_tma && make_brick_tensor_map(&tmap, x, N, a.D, a.H, a.W, a.C, DT_Z + 2 * P, DT_Y + 2 * P, DT_X + 2 * P + 1);
  static DevFlag configured;
  if (!configured) {
    cudaFuncSetAttribute(dwconv_same_tiled_kernel<K>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (cudaFuncSetAttribute(dwconv_same_tiled_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) {
      cudaGetLastError();
      return false;
    }
    configured = true;
  }
  const int tz = (a.D + DT_Z - 1) / DT_Z, ty = (a.H + DT_Y - 1) / DT_Y, tx = (a.W + DT_X - 1) / DT_X;
  // measured on B200 (level 0, batch 4): the persistent double-buffered variant runs 8 warps/SM and loses to two
  // resident non-persistent CTAs (374 vs 335 us per launch, the brick compute is ALU/FMA-issue bound) -> opt-in only
  static const bool no_persist = getenv("PCB_DW_PERSIST") == nullptr;
  const size_t smem_p = (size_t)2 * (DT_Z + 2 * P) * (DT_Y + 2 * P) * (DT_X + 2 * P + 1) * 64 + (size_t)K * K * K * 32 * 4 + 64 * 8 + 32;
  if (use_tma && !no_persist && smem_p <= 227 * 1024) {
    static DevFlag configured_p;
    if (!configured_p) {
      cudaFuncSetAttribute(dwconv_same_persist_kernel<K>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
      configured_p = cudaFuncSetAttribute(dwconv_same_persist_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) == cudaSuccess;
      if (!configured_p) cudaGetLastError();
    }
    if (configured_p) {
      const int nbricks = tz * ty * tx;
      const int64_t groups = (int64_t)(a.C / 32) * N;
      int ctas = (int)(148 / groups);                 // one resident CTA per SM, whole grid in a single wave
      if (ctas < 1) ctas = 1;
      if (ctas > nbricks) ctas = nbricks;
      dim3 grid((unsigned)ctas, (unsigned)(a.C / 32), (unsigned)N);
      dwconv_same_persist_kernel<K><<<grid, 256, smem_p, st>>>(w, b, y, stats, add, a, ty, tx, nbricks, tmap);
      return true;
    }
  }
  if ((int64_t)tz * ty * tx * (a.C / 32) >= (1ll << 31)) return false;
  dim3 grid((unsigned)(tz * ty * tx * (a.C / 32)), 1u, (unsigned)N);
  dwconv_same_tiled_kernel<K><<<grid, 256, smem, st>>>(x, w, b, y, stats, add, a, ty, tx, tmap, use_tma);
  return true;
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
  staged_copy<8>(rows * kc8, tid, 128,
      [&](int q) { const int r = q >> __ffs(kc8) - 1, c8 = q & (kc8 - 1); return __ldg(src + (int64_t)r * pitch8 + c8); },
      [&](int q, const uint4& v) {
        const int r = q >> __ffs(kc8) - 1, c8 = q & (kc8 - 1);
        *reinterpret_cast<uint4*>(dst + (r >> 3) * sbo + c8 * 128 + (r & 7) * 16) = v;
      });
}

__global__ void __launch_bounds__(128, 4) mlp_kernel(MlpArgs a) {
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
        staged_copy<8>(128 * kc8, tid, 128,
            [&](int q) {
              const int r = q >> __ffs(kc8) - 1, c8 = q & (kc8 - 1);
              const int64_t ry = sRowY[r];
              return ry >= 0 ? __ldg(yn + ry * (a.C >> 3) + kc * kc8 + c8) : make_uint4(0, 0, 0, 0);
            },
            [&](int q, const uint4& raw) {
              const int r = q >> __ffs(kc8) - 1, c8 = q & (kc8 - 1);
              uint4 v = make_uint4(0, 0, 0, 0);
              if (sRowY[r] >= 0) {
                float f[8];
                unpack8(raw, f);
                const int c0 = kc * a.KC + c8 * 8;
#pragma unroll
                for (int j = 0; j < 8; ++j) f[j] = fmaf(f[j], sScale[c0 + j], sShift[c0 + j]);
                v = pack8(f);
              }
              *reinterpret_cast<uint4*>(sA + (r >> 3) * sbo + c8 * 128 + (r & 7) * 16) = v;
            });
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
        const float4* bp = reinterpret_cast<const float4*>(a.b2 + hc * a.N1 + c16 * 16);
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) {
          const float4 b = __ldg(bp + j4);
          gelu_fast2(__uint_as_float(v[4 * j4]) + b.x, __uint_as_float(v[4 * j4 + 1]) + b.y, g[4 * j4], g[4 * j4 + 1]);
          gelu_fast2(__uint_as_float(v[4 * j4 + 2]) + b.z, __uint_as_float(v[4 * j4 + 3]) + b.w, g[4 * j4 + 2], g[4 * j4 + 3]);
        }
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
      staged_copy<8>(128 * kr8, tid, 128,
          [&](int q) {
            const int r = q >> __ffs(kr8) - 1, c8 = q & (kr8 - 1);
            const int64_t rx = sRowX[r];
            return rx >= 0 ? __ldg(xn + rx * (a.Cr >> 3) + kc * kr8 + c8) : make_uint4(0, 0, 0, 0);
          },
          [&](int q, const uint4& v) {
            const int r = q >> __ffs(kr8) - 1, c8 = q & (kr8 - 1);
            *reinterpret_cast<uint4*>(sA + (r >> 3) * sbo + c8 * 128 + (r & 7) * 16) = v;
          });
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

// ============================================================================ persistent fused MLP (levels 0/1)
// Same math as mlp_kernel, restructured for the HBM-bound top levels (C <= 64):
//   * persistent: one CTA per SM loops over 128-voxel tiles; W2 / W3 / Wr and the GroupNorm affine of every
//     sample stay resident in shared memory for the whole kernel;
//   * warp-specialised: 4 loader warps (batched 128-bit loads of the next tile, GN-apply, K-major staging),
//     1 MMA warp (one thread issues tcgen05.mma), 2 x 4 epilogue warps that alternate tiles
//     (TMEM -> +b2 -> GELU -> bf16 -> shared;  TMEM -> +b3 (+br) + residual -> HBM);
//   * TMEM accumulators and the A / H shared-memory tiles are double buffered; every hand-off is an
//     mbarrier (tcgen05.commit on the MMA side), there is no __syncthreads in the tile loop.
struct MlpFusedArgs {
  MlpArgs m;
  int N;             // samples
  int nst;           // A-tile stages (4: one loader warp per stage; 2: two warps per stage)
  int nsx;           // xs-tile stages (<= nst)
  int64_t tps;       // tiles per sample
  int64_t ntiles;    // N * tps
  int fast;          // warp-per-tile register loader (C, Cr in {32, 64}; nst == 4)
  int ld16;          // opt-in (PCB_FWD_LD16=1): 16 instead of 8 loads in flight per loader lane
  uint32_t dm2, dm1; // magic multipliers of the exact division by o2 / o1 (n < 2^31):  n / d == (n * dm) >> ds
  int ds2, ds1;
  uint32_t dmt;      // ... and by tps (tile -> sample)
  int dst;
  int nb2;           // acc2 TMEM buffers per epilogue group: 2 = epilogue 2 of tile k-1 runs after epilogue 1 of tile k
};

// exact n / d for 0 <= n < 2^31 with m = ceil(2^(31+l) / d), l = ceil(log2 d), shift = 31 + l (host: mf_magic)
__device__ __forceinline__ int mf_fdiv(int n, uint32_t m, int sh) {
  return (int)(((uint64_t)(uint32_t)n * (uint64_t)m) >> sh);
}

constexpr int MF_LOAD_WARPS = 4, MF_EPI_WARPS = 8, MF_MMA_WARPS = 2, MF_THREADS = 32 * (MF_LOAD_WARPS + MF_EPI_WARPS + MF_MMA_WARPS);

__device__ __forceinline__ void mf_row_sources(const MlpArgs& a, int ov, int& ry, int& rx) {
  ry = -1; rx = -1;
  if (ov >= (int)a.Vout) return;
  if (a.mode == PCB_DW_UP) {
    const int ox = ov % a.o2, t = ov / a.o2, oy = t % a.o1, oz = t / a.o1;
    if (ox >= 1 && oy >= 1 && oz >= 1) {
      ry = ((oz - 1) * (a.o1 - 1) + (oy - 1)) * (a.o2 - 1) + (ox - 1);
      if (!((ox - 1) & 1) && !((oy - 1) & 1) && !((oz - 1) & 1))
        rx = (((oz - 1) >> 1) * a.x1 + ((oy - 1) >> 1)) * a.x2 + ((ox - 1) >> 1);
    }
  } else if (a.mode == PCB_DW_DOWN) {
    const int ox = ov % a.o2, t = ov / a.o2, oy = t % a.o1, oz = t / a.o1;
    ry = ov;
    rx = ((2 * oz) * a.x1 + 2 * oy) * a.x2 + 2 * ox;
  } else {
    ry = ov;
  }
}

// row sources of output voxel ov with the two divisions done by multiplication (same result as mf_row_sources)
__device__ __forceinline__ void mf_row_sources_fast(const MlpFusedArgs& fa, int ov, int& ry, int& rx) {
  const MlpArgs& a = fa.m;
  ry = -1; rx = -1;
  if (ov >= (int)a.Vout) return;
  if (a.mode == PCB_DW_SAME) { ry = ov; return; }
  const int t = mf_fdiv(ov, fa.dm2, fa.ds2), ox = ov - t * a.o2;
  const int oz = mf_fdiv(t, fa.dm1, fa.ds1), oy = t - oz * a.o1;
  if (a.mode == PCB_DW_UP) {
    if (ox >= 1 && oy >= 1 && oz >= 1) {
      ry = ((oz - 1) * (a.o1 - 1) + (oy - 1)) * (a.o2 - 1) + (ox - 1);
      if (!(((ox - 1) | (oy - 1) | (oz - 1)) & 1))
        rx = (((oz - 1) >> 1) * a.x1 + ((oy - 1) >> 1)) * a.x2 + ((ox - 1) >> 1);
    }
  } else {
    ry = ov;
    rx = ((2 * oz) * a.x1 + 2 * oy) * a.x2 + 2 * ox;
  }
}

// One warp stages a [128 x 8*C8N] bf16 tile into the K-major canonical layout.  lane = (chunk slot, row inside
// the 8-row core matrix): the 8 lanes of a quarter-warp fill one 128-byte core matrix per 128-bit shared store
// (conflict-free) and every global request covers whole 32-byte sectors.  The lane's channels are fixed, so the
// GroupNorm affine (NORM) stays in registers for the tile.  TAB: row sources come from `tab` (-1 = zero row),
// otherwise rows row0 .. row0+nvalid-1 are read in order.  8 loads are in flight per lane before the first store.
// LD = loads in flight per lane before the first store.  With 4 loader warps x 32 lanes x 8 x 16 B = 16 KB in flight per SM
// the level-0 kernel sits exactly at the latency-bandwidth product it was measured at (2.4 TB/s ~ 148 SMs x 16 KB / 1 us);
// LD = 16 doubles the bytes in flight (opt-in until measured).
template <int C8N, bool NORM, bool TAB, int LD = 8>
__device__ __forceinline__ void mf_stage_tile(uint8_t* __restrict__ dst, const uint4* __restrict__ src,
                                              const int* __restrict__ tab, int row0, int nvalid,
                                              const float* __restrict__ sc, const float* __restrict__ sh, int lane) {
  static_assert(C8N == 4 || C8N == 8, "C8N");
  static_assert(LD == 8 || LD == 16, "LD");
  constexpr int J = C8N / 4;      // chunks per lane per row group
  constexpr int RGB = LD / J;     // row groups per batch of LD loads
  const int rl = lane & 7, cs = lane >> 3;
  uint64_t ps[J][4], pt[J][4];
  if (NORM) {
#pragma unroll
    for (int j = 0; j < J; ++j) {
      const float4* sp = reinterpret_cast<const float4*>(sc + (cs + 4 * j) * 8);
      const float4* tp = reinterpret_cast<const float4*>(sh + (cs + 4 * j) * 8);
      const float4 s0 = sp[0], s1 = sp[1], t0 = tp[0], t1 = tp[1];
      ps[j][0] = pk2(s0.x, s0.y); ps[j][1] = pk2(s0.z, s0.w); ps[j][2] = pk2(s1.x, s1.y); ps[j][3] = pk2(s1.z, s1.w);
      pt[j][0] = pk2(t0.x, t0.y); pt[j][1] = pk2(t0.z, t0.w); pt[j][2] = pk2(t1.x, t1.y); pt[j][3] = pk2(t1.z, t1.w);
    }
  }
  uint8_t* dl = dst + cs * 128 + rl * 16;
#pragma unroll 1
  for (int b = 0; b < 16 / RGB; ++b) {
    uint4 v[LD];
    uint32_t ok = 0;
#pragma unroll
    for (int k = 0; k < LD; ++k) {
      const int rg = b * RGB + k / J, j = k % J;
      const int r = rg * 8 + rl;
      int ry;
      if (TAB) ry = tab[r]; else ry = r < nvalid ? row0 + r : -1;
      v[k] = make_uint4(0, 0, 0, 0);
      if (ry >= 0) { v[k] = __ldg(src + (int64_t)ry * C8N + (cs + 4 * j)); ok |= 1u << k; }
    }
#pragma unroll
    for (int k = 0; k < LD; ++k) {
      const int rg = b * RGB + k / J, j = k % J;
      uint4 o = v[k];
      if (NORM) {
        const bool live = (ok >> k) & 1u;
        float a0, a1, a2, a3, a4, a5, a6, a7;
        upk2(fma2(pk2(bf16_lo(o.x), bf16_hi(o.x)), ps[j][0], pt[j][0]), a0, a1);
        upk2(fma2(pk2(bf16_lo(o.y), bf16_hi(o.y)), ps[j][1], pt[j][1]), a2, a3);
        upk2(fma2(pk2(bf16_lo(o.z), bf16_hi(o.z)), ps[j][2], pt[j][2]), a4, a5);
        upk2(fma2(pk2(bf16_lo(o.w), bf16_hi(o.w)), ps[j][3], pt[j][3]), a6, a7);
        o.x = live ? pack_bf16(a0, a1) : 0u; o.y = live ? pack_bf16(a2, a3) : 0u;
        o.z = live ? pack_bf16(a4, a5) : 0u; o.w = live ? pack_bf16(a6, a7) : 0u;
      }
      *reinterpret_cast<uint4*>(dl + rg * (C8N * 128) + j * 512) = o;
    }
  }
}

template <bool LD16>   // LD16: opt-in loader depth (separate instantiation: the default kernel's code is unchanged)
__global__ void __launch_bounds__(MF_THREADS, 1) mlp_fused_kernel(MlpFusedArgs fa) {
  const MlpArgs& a = fa.m;
  extern __shared__ __align__(128) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int c8n = a.C >> 3, h8n = a.H >> 3, r8n = a.Cr >> 3;
  const bool has_rc = a.wr != nullptr;
  const int NST = fa.nst;   // A-tile stages (4 or 2)
  const int NSX = fa.nsx;   // res-conv source (xs) tile stages: NST, or 2 when shared memory is short
  // ---- shared memory carve-up (all operand tiles in the K-major no-swizzle canonical layout)
  uint8_t* sW2 = smem;                                   // [H x C]
  uint8_t* sW3 = sW2 + a.H * a.C * 2;                    // [Co x H]
  uint8_t* sWr = sW3 + a.Co * a.H * 2;                   // [Co x Cr]
  uint8_t* sA = sWr + (has_rc ? a.Co * a.Cr * 2 : 0);    // NST x [128 x C]
  uint8_t* sX = sA + NST * 128 * a.C * 2;                // NSX x [128 x Cr]
  uint8_t* sH = sX + (has_rc ? NSX * 128 * a.Cr * 2 : 0);   // 2 x [128 x H]
  float* sScale = reinterpret_cast<float*>(sH + 2 * 128 * a.H * 2);   // [N][C]
  float* sShift = sScale + fa.N * a.C;                   // [N][C]
  float* sB2 = sShift + fa.N * a.C;                      // [H]
  float* sB3 = sB2 + a.H;                                // [Co]  (b3 + br)
  int* sRow = reinterpret_cast<int*>(sB3 + a.Co);        // [4 loader warps][2][128] row sources of the tile in flight
  uint64_t* bars = reinterpret_cast<uint64_t*>(sRow + 4 * 2 * 128);
  uint64_t* a_full = bars;          // [4] loaders -> MMA
  uint64_t* a_empty = bars + 4;     // [4] MMA -> loaders
  uint64_t* acc1_full = bars + 8;   // [2] MMA -> epilogue
  uint64_t* h_full = bars + 10;     // [2] epilogue -> MMA
  uint64_t* acc2_full = bars + 12;  // [2 groups][2 buffers] MMA -> epilogue
  uint64_t* acc2_empty = bars + 16; // [2 groups][2 buffers] epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);
  const int NB2 = fa.nb2;           // acc2 buffers per epilogue group

  const uint32_t tmem_cols = tmem_cols_pow2(2 * a.H + 2 * NB2 * a.Co);
  if (warp == MF_LOAD_WARPS + MF_EPI_WARPS) tmem_alloc(tmem_slot, tmem_cols);
  if (tid == 0) {
    for (int i = 0; i < 4; ++i) { mbar_init(&a_full[i], 32 * (MF_LOAD_WARPS / NST)); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc1_full[i], 1); mbar_init(&h_full[i], 128); }
    for (int i = 0; i < 4; ++i) { mbar_init(&acc2_full[i], 1); mbar_init(&acc2_empty[i], 128); }
    fence_mbar_init();
  }
  // resident weights, biases, GroupNorm affine of every sample
  {
    const uint32_t sbo2 = c8n * 128;
    for (int q = tid; q < a.H * c8n; q += MF_THREADS) {
      const int r = q / c8n, c8 = q - r * c8n;
      *reinterpret_cast<uint4*>(sW2 + (r >> 3) * sbo2 + c8 * 128 + (r & 7) * 16) = __ldg(a.w2 + (int64_t)r * c8n + c8);
    }
    const uint32_t sbo3 = h8n * 128;
    for (int q = tid; q < a.Co * h8n; q += MF_THREADS) {
      const int r = q / h8n, c8 = q - r * h8n;
      *reinterpret_cast<uint4*>(sW3 + (r >> 3) * sbo3 + c8 * 128 + (r & 7) * 16) = __ldg(a.w3 + (int64_t)r * h8n + c8);
    }
    if (has_rc) {
      const uint32_t sbor = r8n * 128;
      for (int q = tid; q < a.Co * r8n; q += MF_THREADS) {
        const int r = q / r8n, c8 = q - r * r8n;
        *reinterpret_cast<uint4*>(sWr + (r >> 3) * sbor + c8 * 128 + (r & 7) * 16) = __ldg(a.wr + (int64_t)r * r8n + c8);
      }
    }
    for (int i = tid; i < a.H; i += MF_THREADS) sB2[i] = a.b2[i];
    for (int i = tid; i < a.Co; i += MF_THREADS) sB3[i] = a.b3[i] + (has_rc ? a.br[i] : 0.f);
    for (int i = tid; i < fa.N * a.C; i += MF_THREADS) {
      const int n = i / a.C, c = i - n * a.C;
      const double sm = a.stats[(int64_t)n * 2 * a.C + c], q = a.stats[(int64_t)n * 2 * a.C + a.C + c];
      const double mean = sm * (double)a.inv_count;
      double var = q * (double)a.inv_count - mean * mean;
      if (var < 0.0) var = 0.0;
      const float g = a.gamma[c] * (float)(1.0 / sqrt(var + 1e-5));
      sScale[i] = g;
      sShift[i] = a.beta[c] - (float)mean * g;
    }
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // TMEM columns: acc1[g] at g*H, acc2[g][b] at 2H + (g*NB2 + b)*Co

  if (warp < MF_LOAD_WARPS) {
    // ===================================================================== loaders
    // NST == 4: every warp owns the tiles  it == warp (mod 4)  and stage `warp` (4 tiles of loads in flight
    // per SM); NST == 2: warp pairs {0,1} / {2,3} own the even / odd tiles.
    const int wpg = MF_LOAD_WARPS / NST;             // warps per loader group (1 or 2)
    const int grp = warp / wpg;                      // loader group = stage
    const int lt = (warp - grp * wpg) * 32 + lane;   // thread index inside the group
    const int nthr = wpg * 32;
    const uint32_t sboA = c8n * 128, sboX = r8n * 128;
    const int csh = __ffs(c8n) - 1, rsh = r8n > 0 ? __ffs(r8n) - 1 : 0;
    const bool cpow2 = (c8n & (c8n - 1)) == 0, rpow2 = (r8n & (r8n - 1)) == 0;
    int* rowY = sRow + grp * 256;
    int* rowX = rowY + 128;
    int64_t it = grp;
    if (fa.fast) {
      // ---- warp-per-tile register loader (NST == 4): warp w owns stage w and the tiles  it == w (mod 4)
      const bool tabY = a.mode == PCB_DW_UP;
      int uses = 0;
      for (int64_t g = blockIdx.x + (int64_t)warp * gridDim.x; g < fa.ntiles; g += 4ll * gridDim.x, it += 4, ++uses) {
        if (uses >= 1) mbar_wait(&a_empty[warp], (uint32_t)((uses - 1) & 1));
        const int n = (int)(g / fa.tps);
        const int tile0 = (int)((g - (int64_t)n * fa.tps) * 128);
        const int nvalid = min(128, (int)a.Vout - tile0);
        if (tabY || has_rc) {
          __syncwarp();
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            int ry, rx;
            mf_row_sources_fast(fa, tile0 + lane + 32 * k, ry, rx);
            rowY[lane + 32 * k] = ry; rowX[lane + 32 * k] = rx;
          }
          __syncwarp();
        }
        const uint4* yn = a.y + (int64_t)n * a.Vy * c8n;
        const float* sc = sScale + n * a.C;
        const float* sh = sShift + n * a.C;
        uint8_t* dA = sA + warp * 128 * a.C * 2;
        if (c8n == 8) {
          if (tabY) mf_stage_tile<8, true, true>(dA, yn, rowY, tile0, nvalid, sc, sh, lane);
          else if (LD16) mf_stage_tile<8, true, false, 16>(dA, yn, rowY, tile0, nvalid, sc, sh, lane);
          else mf_stage_tile<8, true, false>(dA, yn, rowY, tile0, nvalid, sc, sh, lane);
        } else {
          if (tabY) mf_stage_tile<4, true, true>(dA, yn, rowY, tile0, nvalid, sc, sh, lane);
          else if (LD16) mf_stage_tile<4, true, false, 16>(dA, yn, rowY, tile0, nvalid, sc, sh, lane);
          else mf_stage_tile<4, true, false>(dA, yn, rowY, tile0, nvalid, sc, sh, lane);
        }
        if (has_rc) {
          const uint4* xn = a.xs + (int64_t)n * a.Vin * r8n;
          if (NSX < 4 && it >= NSX) {                 // see the generic loader below
            const int64_t jt = it - NSX;
            mbar_wait(&a_empty[jt & 3], (uint32_t)((jt >> 2) & 1));
          }
          uint8_t* dX = sX + (int)(it % NSX) * 128 * a.Cr * 2;
          if (r8n == 8) mf_stage_tile<8, false, true>(dX, xn, rowX, 0, 0, nullptr, nullptr, lane);
          else mf_stage_tile<4, false, true>(dX, xn, rowX, 0, 0, nullptr, nullptr, lane);
        }
        fence_proxy_async_smem();
        mbar_arrive(&a_full[warp]);
      }
    } else
    for (int64_t g = blockIdx.x + (int64_t)grp * gridDim.x; g < fa.ntiles; g += (int64_t)NST * gridDim.x, it += NST) {
      const int64_t use = it / NST;                  // how many times this stage has been filled before
      if (use >= 1) mbar_wait(&a_empty[grp], (uint32_t)((use - 1) & 1));
      const int n = (int)(g / fa.tps);
      const int tile0 = (int)((g - (int64_t)n * fa.tps) * 128);
      if (wpg == 1) __syncwarp(); else asm volatile("bar.sync %0, 64;" ::"r"(1 + grp) : "memory");   // table reuse
      for (int r = lt; r < 128; r += nthr) {
        int ry, rx;
        mf_row_sources(a, tile0 + r, ry, rx);
        rowY[r] = ry; rowX[r] = rx;
      }
      if (wpg == 1) __syncwarp(); else asm volatile("bar.sync %0, 64;" ::"r"(1 + grp) : "memory");
      const uint4* yn = a.y + (int64_t)n * a.Vy * c8n;
      const float* sc = sScale + n * a.C;
      const float* sh = sShift + n * a.C;
      uint8_t* dA = sA + grp * 128 * a.C * 2;
      staged_copy<8>(128 * c8n, lt, nthr,
          [&](int q) {
            const int r = cpow2 ? (q >> csh) : (q / c8n), c8 = q - r * c8n;
            const int ry = rowY[r];
            return ry >= 0 ? __ldg(yn + (int64_t)ry * c8n + c8) : make_uint4(0, 0, 0, 0);
          },
          [&](int q, const uint4& raw) {
            const int r = cpow2 ? (q >> csh) : (q / c8n), c8 = q - r * c8n;
            uint4 v = make_uint4(0, 0, 0, 0);
            if (rowY[r] >= 0) {
              const float4* scp = reinterpret_cast<const float4*>(sc + c8 * 8);
              const float4* shp = reinterpret_cast<const float4*>(sh + c8 * 8);
              const float4 s0 = scp[0], s1 = scp[1], t0 = shp[0], t1 = shp[1];
              float a0, a1, a2, a3, a4, a5, a6, a7;
              upk2(fma2(pk2(bf16_lo(raw.x), bf16_hi(raw.x)), pk2(s0.x, s0.y), pk2(t0.x, t0.y)), a0, a1);
              upk2(fma2(pk2(bf16_lo(raw.y), bf16_hi(raw.y)), pk2(s0.z, s0.w), pk2(t0.z, t0.w)), a2, a3);
              upk2(fma2(pk2(bf16_lo(raw.z), bf16_hi(raw.z)), pk2(s1.x, s1.y), pk2(t1.x, t1.y)), a4, a5);
              upk2(fma2(pk2(bf16_lo(raw.w), bf16_hi(raw.w)), pk2(s1.z, s1.w), pk2(t1.z, t1.w)), a6, a7);
              v.x = pack_bf16(a0, a1); v.y = pack_bf16(a2, a3); v.z = pack_bf16(a4, a5); v.w = pack_bf16(a6, a7);
            }
            *reinterpret_cast<uint4*>(dA + (r >> 3) * sboA + c8 * 128 + (r & 7) * 16) = v;
          });
      if (has_rc) {
        const uint4* xn = a.xs + (int64_t)n * a.Vin * r8n;
        const int sx = (int)(it % NSX);
        // with fewer xs stages than A stages the slot was last used NSX tiles ago: its res-GEMM (second half of
        // that tile, committed on that tile's a_empty barrier) must have retired before it is overwritten
        if (NSX < NST && it >= NSX) {
          const int64_t jt = it - NSX;
          mbar_wait(&a_empty[jt % NST], (uint32_t)((jt / NST) & 1));
        }
        uint8_t* dX = sX + sx * 128 * a.Cr * 2;
        staged_copy<8>(128 * r8n, lt, nthr,
            [&](int q) {
              const int r = rpow2 ? (q >> rsh) : (q / r8n), c8 = q - r * r8n;
              const int rx = rowX[r];
              return rx >= 0 ? __ldg(xn + (int64_t)rx * r8n + c8) : make_uint4(0, 0, 0, 0);
            },
            [&](int q, const uint4& v) {
              const int r = rpow2 ? (q >> rsh) : (q / r8n), c8 = q - r * r8n;
              *reinterpret_cast<uint4*>(dX + (r >> 3) * sboX + c8 * 128 + (r & 7) * 16) = v;
            });
      }
      fence_proxy_async_smem();
      mbar_arrive(&a_full[grp]);
    }
  } else if (warp >= MF_LOAD_WARPS + MF_EPI_WARPS) {
    // ===================================================================== MMA issuers (one thread per epilogue group)
    // The two epilogue groups are independent pipelines (group g owns the CTA's tiles it == g mod 2, its own acc1 / sH / acc2
    // buffers and barriers), each served by its own issuing thread with blocking mbarrier waits, so a group never waits behind
    // a hand-off of the other one.  (A single issuer that polls both groups with mbarrier.test_wait was measured 23 % slower
    // than the in-order issuer it replaced: test_wait costs ~150 cycles per probe.)  Per group and local tile k:
    //   GEMM1(k): a_full(tile)                                  -> acc1[g]        (acc1[g] is free: h_full(k-1) was seen)
    //   GEMM2(k): h_full[g](k), acc2_empty[g][b] of tile k-NB2  -> acc2[g][b]     (b = k mod NB2)
    const int g = warp - (MF_LOAD_WARPS + MF_EPI_WARPS);
    if (lane == 0) {
      const uint32_t idesc1 = umma_idesc_bf16(128, a.H, 0, 0), idesc2 = umma_idesc_bf16(128, a.Co, 0, 0);
      const uint64_t dW2 = umma_desc(smem_u32(sW2), 128, c8n * 128);
      const uint64_t dW3 = umma_desc(smem_u32(sW3), 128, h8n * 128);
      const uint64_t dWr = has_rc ? umma_desc(smem_u32(sWr), 128, r8n * 128) : 0;
      int64_t k = 0;
      for (int64_t gt = blockIdx.x + (int64_t)g * gridDim.x; gt < fa.ntiles; gt += 2 * (int64_t)gridDim.x, ++k) {
        const int64_t it = 2 * k + g;
        const int s = (int)(it % NST);
        mbar_wait(&a_full[s], (uint32_t)((it / NST) & 1));
        tc_fence_after();
        const uint32_t acc1 = tmem_base + g * a.H;
        const uint64_t dA = umma_desc(smem_u32(sA + s * 128 * a.C * 2), 128, c8n * 128);
        for (int q = 0; q < a.C / 16; ++q) umma_bf16(acc1, dA + (uint64_t)(q * 16), dW2 + (uint64_t)(q * 16), idesc1, q > 0 ? 1u : 0u);
        tc_commit(&acc1_full[g]);
        if (!has_rc) tc_commit(&a_empty[s]);
        mbar_wait(&h_full[g], (uint32_t)(k & 1));
        const int b = NB2 == 2 ? (int)(k & 1) : 0;
        const int64_t prev = k - NB2;                 // the tile whose epilogue 2 last read acc2[g][b]
        if (prev >= 0) mbar_wait(&acc2_empty[g * 2 + b], (uint32_t)((prev / NB2) & 1));
        tc_fence_after();
        const uint32_t acc2 = tmem_base + 2 * a.H + (g * NB2 + b) * a.Co;
        const uint64_t dH = umma_desc(smem_u32(sH + g * 128 * a.H * 2), 128, h8n * 128);
        for (int q = 0; q < a.H / 16; ++q) umma_bf16(acc2, dH + (uint64_t)(q * 16), dW3 + (uint64_t)(q * 16), idesc2, q > 0 ? 1u : 0u);
        if (has_rc) {
          const uint64_t dX = umma_desc(smem_u32(sX + (int)(it % NSX) * 128 * a.Cr * 2), 128, r8n * 128);
          for (int q = 0; q < a.Cr / 16; ++q) umma_bf16(acc2, dX + (uint64_t)(q * 16), dWr + (uint64_t)(q * 16), idesc2, 1u);
        }
        tc_commit(&acc2_full[g * 2 + b]);
        if (has_rc) tc_commit(&a_empty[s]);
      }
    }
  } else {
    // ===================================================================== epilogue warpgroups
    // Group eg owns the tiles it == eg (mod 2).  With two acc2 buffers (NB2 == 2) the order per local tile k is
    //   epilogue 1 (k)  ->  h_full  ->  epilogue 2 (k-1)
    // so the GEMM2 of tile k (and the GEMM1 of k+1 behind it) runs while this group still has the output pass of tile k-1 to
    // do: the group never idles on acc2_full (ncu, round 2: 38 % of the epilogue warps' samples sat in that wait).
    const int eg = (warp - MF_LOAD_WARPS) >> 2;          // 0 / 1: which tiles (it & 1) this group owns
    const int wq = warp & 3;                             // TMEM lane quarter this warp may access
    const int row = wq * 32 + lane;
    const uint32_t lane_off = (uint32_t)(wq * 32) << 16;
    const uint32_t sboH = h8n * 128;
    const int co8 = a.Co >> 3;
    const bool prefetch = a.res != nullptr && co8 <= 8;
    const bool pipe = NB2 == 2;
    // ---- epilogue 1: acc1 -> +b2 -> GELU -> bf16 -> sH[eg]
    auto epi1 = [&]() {
      const uint32_t trow = tmem_base + eg * a.H + lane_off;
      uint8_t* dst = sH + eg * 128 * a.H * 2 + (row >> 3) * sboH + (row & 7) * 16;
      for (int c16 = 0; c16 < a.H / 16; ++c16) {
        uint32_t v[16];
        tmem_ld16(trow + c16 * 16, v);
        tmem_ld_wait();
        float gl[16];
        const float4* bp = reinterpret_cast<const float4*>(sB2 + c16 * 16);
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) {
          const float4 b = bp[j4];
          gelu_fast2p(add2(pk2(__uint_as_float(v[4 * j4]), __uint_as_float(v[4 * j4 + 1])), pk2(b.x, b.y)), gl[4 * j4], gl[4 * j4 + 1]);
          gelu_fast2p(add2(pk2(__uint_as_float(v[4 * j4 + 2]), __uint_as_float(v[4 * j4 + 3])), pk2(b.z, b.w)), gl[4 * j4 + 2], gl[4 * j4 + 3]);
        }
        *reinterpret_cast<uint4*>(dst + (c16 * 2) * 128) = pack8(gl);
        *reinterpret_cast<uint4*>(dst + (c16 * 2 + 1) * 128) = pack8(gl + 8);
      }
    };
    // ---- epilogue 2: acc2[eg][b] -> +b3 (+br) (+residual / skip) -> bf16 -> HBM
    auto epi2 = [&](int b, int64_t orow, bool in_range, bool valid, const uint4 (&rpre)[8]) {
      const uint32_t trow = tmem_base + 2 * a.H + (eg * NB2 + b) * a.Co + lane_off;
#pragma unroll 1
      for (int c16 = 0; c16 < a.Co / 16; ++c16) {
        uint32_t v[16];
        tmem_ld16(trow + c16 * 16, v);
        uint4 r0 = make_uint4(0, 0, 0, 0), r1 = make_uint4(0, 0, 0, 0);
        if (prefetch) {
          // constant-index selection keeps rpre in registers
#pragma unroll
          for (int c = 0; c < 4; ++c) if (c == c16) { r0 = rpre[2 * c]; r1 = rpre[2 * c + 1]; }
        } else if (a.res != nullptr && in_range) {
          r0 = __ldg(a.res + orow + c16 * 2); r1 = __ldg(a.res + orow + c16 * 2 + 1);
        }
        tmem_ld_wait();
        if (!in_range) continue;
        const float4* b3p = reinterpret_cast<const float4*>(sB3 + c16 * 16);
        const uint32_t rw[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
        uint32_t ow[8];
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) {
          const float4 bb = b3p[j4];
          uint64_t lo = pk2(bf16_lo(rw[2 * j4]), bf16_hi(rw[2 * j4])), hi = pk2(bf16_lo(rw[2 * j4 + 1]), bf16_hi(rw[2 * j4 + 1]));
          if (valid) {
            lo = add2(lo, add2(pk2(__uint_as_float(v[4 * j4]), __uint_as_float(v[4 * j4 + 1])), pk2(bb.x, bb.y)));
            hi = add2(hi, add2(pk2(__uint_as_float(v[4 * j4 + 2]), __uint_as_float(v[4 * j4 + 3])), pk2(bb.z, bb.w)));
          }
          float e0, e1, e2, e3;
          upk2(lo, e0, e1);
          upk2(hi, e2, e3);
          ow[2 * j4] = pack_bf16(e0, e1);
          ow[2 * j4 + 1] = pack_bf16(e2, e3);
        }
        a.out[orow + c16 * 2] = make_uint4(ow[0], ow[1], ow[2], ow[3]);
        a.out[orow + c16 * 2 + 1] = make_uint4(ow[4], ow[5], ow[6], ow[7]);
      }
    };
    auto load_res = [&](uint4 (&rpre)[8], int64_t orow, bool in_range) {
#pragma unroll
      for (int c = 0; c < 8; ++c) rpre[c] = (c < co8 && in_range) ? __ldg(a.res + orow + c) : make_uint4(0, 0, 0, 0);
    };
    int64_t k = 0;
    int64_t p_orow = 0;                                  // row of the tile whose epilogue 2 is pending (pipe)
    bool p_in = false, p_valid = false;
    for (int64_t g = blockIdx.x + (int64_t)eg * gridDim.x; g < fa.ntiles; g += 2 * (int64_t)gridDim.x, ++k) {
      const int n = mf_fdiv((int)g, fa.dmt, fa.dst);
      const int ovi = ((int)g - n * (int)fa.tps) * 128 + row;
      int ry, rx;
      mf_row_sources_fast(fa, ovi, ry, rx);
      const bool in_range = ovi < (int)a.Vout, valid = in_range && ry >= 0;
      const int64_t orow = ((int64_t)n * a.Vout + ovi) * co8;
      uint4 rpre[8];                                     // residual / skip row, in flight during epilogue 1
      if (prefetch) {
        if (!pipe) load_res(rpre, orow, in_range);
        else if (k >= 1) load_res(rpre, p_orow, p_in);
      }
      mbar_wait(&acc1_full[eg], (uint32_t)(k & 1));
      // sH[eg] is read by GEMM2(k-1): it has retired when acc2_full of tile k-1 completed (issued an epilogue-2 pass ago)
      if (pipe && k >= 1) mbar_wait(&acc2_full[eg * 2 + (int)((k - 1) & 1)], (uint32_t)(((k - 1) >> 1) & 1));
      tc_fence_after();
      epi1();
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(&h_full[eg]);
      if (pipe) {
        if (k >= 1) {
          epi2((int)((k - 1) & 1), p_orow, p_in, p_valid, rpre);
          tc_fence_before();
          mbar_arrive(&acc2_empty[eg * 2 + (int)((k - 1) & 1)]);
        }
        p_orow = orow; p_in = in_range; p_valid = valid;
      } else {
        mbar_wait(&acc2_full[eg * 2], (uint32_t)(k & 1));
        tc_fence_after();
        epi2(0, orow, in_range, valid, rpre);
        tc_fence_before();
        mbar_arrive(&acc2_empty[eg * 2]);
      }
    }
    if (pipe && k >= 1) {                                // drain: epilogue 2 of the group's last tile
      uint4 rpre[8];
      if (prefetch) load_res(rpre, p_orow, p_in);
      mbar_wait(&acc2_full[eg * 2 + (int)((k - 1) & 1)], (uint32_t)(((k - 1) >> 1) & 1));
      tc_fence_after();
      epi2((int)((k - 1) & 1), p_orow, p_in, p_valid, rpre);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == MF_LOAD_WARPS + MF_EPI_WARPS) tmem_dealloc(tmem_base, tmem_cols);
}

// magic numbers of mf_fdiv: l = ceil(log2 d), m = ceil(2^(31+l) / d) (< 2^32), shift = 31 + l
static void mf_magic(uint32_t d, uint32_t& m, int& sh) {
  int l = 0;
  while ((1ull << l) < d) ++l;
  const unsigned __int128 p = (unsigned __int128)1 << (31 + l);
  m = (uint32_t)((p + d - 1) / d);
  sh = 31 + l;
}

static size_t mlp_fused_smem(const MlpArgs& a, int N, int nst, int nsx) {
  const bool rc = a.wr != nullptr;
  return (size_t)a.H * a.C * 2 + (size_t)a.Co * a.H * 2 + (rc ? (size_t)a.Co * a.Cr * 2 : 0) + (size_t)nst * 128 * a.C * 2 +
         (rc ? (size_t)nsx * 128 * a.Cr * 2 : 0) + (size_t)2 * 128 * a.H * 2 + (size_t)2 * N * a.C * 4 + (size_t)(a.H + a.Co) * 4 +
         4 * 2 * 128 * 4 + 20 * 8 + 16 + 128;
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
  {
    const int ch = (int)(C >> 3);
    static const bool no_stream = getenv("PCB_NO_STEM_STREAM") != nullptr;
    if (!no_stream && Cin <= 4 && ch <= 256 && (ch & (ch - 1)) == 0 && N <= 65535 && in_dtype >= PCB_F32 && in_dtype <= PCB_BF16) {
      int sh = 0;
      while ((1 << sh) < ch) ++sh;
      int64_t nb = (nvox * ch + 256 * 4 - 1) / (256 * 4);
      if (nb > 148 * 8) nb = 148 * 8;
      if (nb < 1) nb = 1;
      dim3 g2((unsigned)nb, (unsigned)N);
#define PCB_STEM(T, K) stem_stream_kernel<T, K><<<g2, 256, 0, st>>>((const T*)x, w, b, o, (int)C, nvox, sh)
#define PCB_STEM_T(T)                                                                     \
      do {                                                                                \
        if (Cin == 1) PCB_STEM(T, 1); else if (Cin == 2) PCB_STEM(T, 2);                  \
        else if (Cin == 3) PCB_STEM(T, 3); else PCB_STEM(T, 4);                           \
      } while (0)
      if (in_dtype == PCB_F32) PCB_STEM_T(float);
      else if (in_dtype == PCB_F16) PCB_STEM_T(__half);
      else PCB_STEM_T(__nv_bfloat16);
#undef PCB_STEM_T
#undef PCB_STEM
      PCB_CHECK_LAUNCH("pcb_stem_fwd");
      return PCB_OK;
    }
  }
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
  cudaStream_t st = (cudaStream_t)stream;
  if (mode == PCB_DW_SAME && C % 32 == 0 && a.add_mode != 2 && a.Do == a.D && a.Ho == a.H && a.Wo == a.W &&
      (int64_t)((a.D + DT_Z - 1) / DT_Z) * ((a.H + DT_Y - 1) / DT_Y) * ((a.W + DT_X - 1) / DT_X) < (1ll << 31) && N <= 65535) {
    bool ok = false;
    if (k == 3) ok = launch_dw_tiled<3>(st, (const uint4*)x, w, b, (uint4*)y, stats, (const uint4*)add, a, N);
    else if (k == 5) ok = launch_dw_tiled<5>(st, (const uint4*)x, w, b, (uint4*)y, stats, (const uint4*)add, a, N);
    else ok = launch_dw_tiled<7>(st, (const uint4*)x, w, b, (uint4*)y, stats, (const uint4*)add, a, N);
    if (ok) { PCB_CHECK_LAUNCH(what); return PCB_OK; }
  }
  if (mode == PCB_DW_DOWN && k == 3 && C % 32 == 0 && a.add_mode != 2 && a.Do == (a.D - 1) / 2 + 1 && a.Ho == (a.H - 1) / 2 + 1 &&
      a.Wo == (a.W - 1) / 2 + 1) {
    if (launch_dw_down3_tiled(st, (const uint4*)x, w, b, (uint4*)y, stats, (const uint4*)add, a, N)) {
      PCB_CHECK_LAUNCH(what);
      return PCB_OK;
    }
  }
  const size_t smem = 2 * C * sizeof(double);
  static const bool no_up3 = getenv("PCB_NO_UP3") != nullptr;
  if (mode == PCB_DW_UP && k == 3 && !no_up3 && C % 32 == 0 &&
      launch_dw_up3_tiled(st, (const uint4*)x, w, b, (uint4*)y, stats, (const uint4*)add, a, N)) {
    PCB_CHECK_LAUNCH(what);
    return PCB_OK;
  }
  if (mode == PCB_DW_UP && k == 3 && !no_up3) {
    static const int up_xb = getenv("PCB_UP3_XB") ? atoi(getenv("PCB_UP3_XB")) : 4;
    const int ow = up_xb == 2 ? 4 : 8;   // outputs per thread along W
    const int64_t items_up = (int64_t)a.Do * a.Ho * ((a.Wo + ow - 1) / ow) * (C / 8);
    dim3 grid_up((unsigned)((items_up + 255) / 256), (unsigned)N);
    if (up_xb == 2) dwconv_up3_kernel<2><<<grid_up, 256, smem, st>>>((const uint4*)x, w, b, (uint4*)y, stats, (const uint4*)add, a);
    else dwconv_up3_kernel<4><<<grid_up, 256, smem, st>>>((const uint4*)x, w, b, (uint4*)y, stats, (const uint4*)add, a);
    PCB_CHECK_LAUNCH(what);
    return PCB_OK;
  }
  const int64_t items = (int64_t)a.Do * a.Ho * ((a.Wo + DW_XB - 1) / DW_XB) * (C / 8);
  dim3 grid((unsigned)((items + 255) / 256), (unsigned)N);
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
  static DevFlag configured;
  if (!configured) {
    cudaFuncSetAttribute(mlp_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (cudaFuncSetAttribute(mlp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024)) != cudaSuccess) {
      set_error("pcb_mlp_fwd: cudaFuncSetAttribute failed: %s", cudaGetErrorString(cudaGetLastError()));
      return PCB_ERR_CUDA;
    }
    configured = true;
  }
  // top levels: persistent warp-specialised kernel (weights resident, single K / hidden chunk)
  {
    int nst = 4, nsx = 4;
    size_t fsm = mlp_fused_smem(a, (int)N, nst, nsx);
    if (fsm > 227 * 1024) { nsx = 2; fsm = mlp_fused_smem(a, (int)N, nst, nsx); }     // 4 A stages, 2 xs stages
    if (fsm > 227 * 1024) { nst = 2; fsm = mlp_fused_smem(a, (int)N, nst, nsx); }
    const bool pow2 = ((C & (C - 1)) == 0) && ((H & (H - 1)) == 0 || true);
    if (getenv("PCB_NO_FUSED") == nullptr && C <= 64 && H <= 256 && Co <= 128 && (!wr || Cr <= 64) && 2 * (H + Co) <= 512 &&
        N <= 8 && fsm <= 227 * 1024 && pow2 && a.Vout < (1ll << 30) && a.Vin < (1ll << 30)) {
      static DevFlag fconf;
      if (!fconf) {
        cudaFuncSetAttribute(mlp_fused_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        cudaFuncSetAttribute(mlp_fused_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (cudaFuncSetAttribute(mlp_fused_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess ||
            cudaFuncSetAttribute(mlp_fused_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) {
          set_error("pcb_mlp_fwd: cudaFuncSetAttribute(fused) failed"); return PCB_ERR_CUDA;
        }
        fconf = true;
      }
      MlpFusedArgs fa;
      fa.m = a; fa.N = (int)N; fa.nst = nst; fa.nsx = nsx; fa.tps = (a.Vout + 127) / 128; fa.ntiles = fa.tps * N;
      fa.fast = nst == 4 && (C == 32 || C == 64) && (!wr || Cr == 32 || Cr == 64) && getenv("PCB_OLD_LOADER") == nullptr;
      { const char* e16 = getenv("PCB_FWD_LD16"); fa.ld16 = (e16 && e16[0] == '1') ? 1 : 0; }
      mf_magic((uint32_t)(a.o2 > 0 ? a.o2 : 1), fa.dm2, fa.ds2);
      mf_magic((uint32_t)(a.o1 > 0 ? a.o1 : 1), fa.dm1, fa.ds1);
      mf_magic((uint32_t)fa.tps, fa.dmt, fa.dst);
      { const char* np = getenv("PCB_FWD_NOPIPE"); fa.nb2 = (2 * H + 4 * Co <= 512 && !(np && np[0] == '1')) ? 2 : 1; }
      int ctas = 148;
      if (fa.ntiles < ctas) ctas = (int)fa.ntiles;
      if (fa.ld16 && fa.fast) mlp_fused_kernel<true><<<ctas, MF_THREADS, fsm, (cudaStream_t)stream>>>(fa);
      else mlp_fused_kernel<false><<<ctas, MF_THREADS, fsm, (cudaStream_t)stream>>>(fa);
      PCB_CHECK_LAUNCH("pcb_mlp_fwd(fused)");
      return PCB_OK;
    }
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
