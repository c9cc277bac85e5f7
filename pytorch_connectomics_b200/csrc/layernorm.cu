// pcb200 — channels-first LayerNorm of the MedNeXt blocks (norm_type = "layer").
//
// Replaces upstream nnunet_mednext blocks.py::LayerNorm(data_format="channels_first") as built for
// cfg.model.mednext.norm = "layer" (connectomics/models/architectures/mednext_models.py:449-476): per voxel,
//   u = mean_c(y), s = mean_c((y - u)^2), yhat = (y - u) / sqrt(s + 1e-5) * weight[c] + bias[c].
// Activations are channels-last bf16 [rows, C], so a voxel's channel vector is one contiguous row: LPR lanes own a
// row (8 channels = one 128-bit load per lane and chunk), the two reductions are warp shuffles, nothing is staged.
// The normalised tensor feeds the fused MLP kernels with identity GroupNorm constants.  Backward:
//   dy = rstd * (g*w - mean_c(g*w) - xhat * mean_c(g*w*xhat)),  dweight[c] += sum_v g*xhat,  dbias[c] += sum_v g.
#include "../../include/pcb200.h"
#include <stdlib.h>

#include "pcb_common.cuh"

namespace pcb {

constexpr int LN_MAXCH = 8;    // chunks of 8 channels per lane (C <= 32 lanes * 8 * 8 = 2048)

// lanes-per-row: C/8 (a power of two) up to a full warp; wider rows take several chunks per lane
static int ln_lanes_per_row(int64_t C) {
  const int ch = (int)(C >> 3);
  if (ch <= 0 || (ch & (ch - 1)) != 0) return 0;      // unsupported channel count
  return ch < 32 ? ch : 32;
}

template <bool BWD, int CPL>
__global__ void __launch_bounds__(256) layernorm_kernel(const uint4* __restrict__ y, const uint4* __restrict__ g,
                                                        const float* __restrict__ weight, const float* __restrict__ bias,
                                                        uint4* __restrict__ out, double* __restrict__ dweight,
                                                        double* __restrict__ dbias, int C, int64_t rows, int lpr) {
  extern __shared__ double s_red[];     // BWD: [2][C]
  const int CH = C >> 3;
  constexpr int cpl = CPL;              // chunks of 8 channels per lane
  const int tid = threadIdx.x, lane = tid & 31;
  const int sub = lane % lpr;           // lane's position inside its row group
  const int rpw = 32 / lpr;             // rows per warp
  if (BWD) {
    for (int i = tid; i < 2 * C; i += 256) s_red[i] = 0.0;
    __syncthreads();
  }
  const int64_t warp_global = ((int64_t)blockIdx.x * 256 + tid) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * 256) >> 5;
  const float inv_c = 1.0f / (float)C;
  float gw_acc[CPL][8], gb_acc[CPL][8];   // BWD partial column sums (only the first cpl entries are live)
  if (BWD) {
#pragma unroll
    for (int q = 0; q < CPL; ++q)
#pragma unroll
      for (int j = 0; j < 8; ++j) { gw_acc[q][j] = 0.f; gb_acc[q][j] = 0.f; }
  }
  for (int64_t r0 = warp_global * rpw; r0 < rows; r0 += nwarps * rpw) {
    const int64_t r = r0 + lane / lpr;
    const bool ok = r < rows;
    const uint4* yr = y + r * CH;
    float s1 = 0.f;
    float v[CPL][8];
#pragma unroll
    for (int q = 0; q < CPL; ++q) {
      if (q < cpl) {
        if (ok) unpack8(ldg_nc(yr + sub + q * lpr), v[q]);
        else {
#pragma unroll
          for (int j = 0; j < 8; ++j) v[q][j] = 0.f;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) s1 += v[q][j];
      }
    }
    for (int off = 1; off < lpr; off <<= 1) s1 += __shfl_xor_sync(0xffffffffu, s1, off);
    const float mean = s1 * inv_c;
    float s2 = 0.f;
#pragma unroll
    for (int q = 0; q < CPL; ++q) {
      if (q < cpl) {
#pragma unroll
        for (int j = 0; j < 8; ++j) { v[q][j] -= mean; s2 = fmaf(v[q][j], v[q][j], s2); }
      }
    }
    for (int off = 1; off < lpr; off <<= 1) s2 += __shfl_xor_sync(0xffffffffu, s2, off);
    const float rstd = rsqrtf(s2 * inv_c + 1e-5f);
    if (!BWD) {
#pragma unroll
      for (int q = 0; q < CPL; ++q) {
        if (q < cpl) {
          const int c0 = (sub + q * lpr) * 8;
          float o[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = fmaf(v[q][j] * rstd, __ldg(weight + c0 + j), __ldg(bias + c0 + j));
          if (ok) out[r * CH + sub + q * lpr] = pack8(o);
        }
      }
    } else {
      // a = sum_c g*w, b = sum_c g*w*xhat
      float gv[CPL][8];
      float a = 0.f, b = 0.f;
#pragma unroll
      for (int q = 0; q < CPL; ++q) {
        if (q < cpl) {
          const int c0 = (sub + q * lpr) * 8;
          if (ok) unpack8(ldg_nc(g + r * CH + sub + q * lpr), gv[q]);
          else {
#pragma unroll
            for (int j = 0; j < 8; ++j) gv[q][j] = 0.f;
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float xh = v[q][j] * rstd;
            gw_acc[q][j] = fmaf(gv[q][j], xh, gw_acc[q][j]);
            gb_acc[q][j] += gv[q][j];
            const float gwv = gv[q][j] * __ldg(weight + c0 + j);
            gv[q][j] = gwv;
            v[q][j] = xh;
            a += gwv;
            b = fmaf(gwv, xh, b);
          }
        }
      }
      for (int off = 1; off < lpr; off <<= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, off);
        b += __shfl_xor_sync(0xffffffffu, b, off);
      }
      a *= inv_c; b *= inv_c;
#pragma unroll
      for (int q = 0; q < CPL; ++q) {
        if (q < cpl) {
          float o[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = rstd * (gv[q][j] - a - v[q][j] * b);
          if (ok) out[r * CH + sub + q * lpr] = pack8(o);
        }
      }
    }
  }
  if (BWD) {
    // lanes with the same `sub` own the same channels: butterfly over the row bits, then shared / global f64
#pragma unroll
    for (int q = 0; q < CPL; ++q) {
      if (q < cpl) {
        for (int off = lpr; off < 32; off <<= 1) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            gw_acc[q][j] += __shfl_xor_sync(0xffffffffu, gw_acc[q][j], off);
            gb_acc[q][j] += __shfl_xor_sync(0xffffffffu, gb_acc[q][j], off);
          }
        }
        if (lane < lpr) {
          const int c0 = (sub + q * lpr) * 8;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            atomicAdd(&s_red[c0 + j], (double)gw_acc[q][j]);
            atomicAdd(&s_red[C + c0 + j], (double)gb_acc[q][j]);
          }
        }
      }
    }
    __syncthreads();
    for (int i = tid; i < C; i += 256) {
      atomicAdd(&dweight[i], s_red[i]);
      atomicAdd(&dbias[i], s_red[C + i]);
    }
  }
}

}  // namespace pcb

using namespace pcb;

static int ln_grid(int64_t rows, int lpr) {
  const int64_t rows_per_cta = (int64_t)8 * (32 / lpr);
  int64_t nb = (rows + rows_per_cta - 1) / rows_per_cta;
  if (nb > 148 * 8) nb = 148 * 8;
  return (int)(nb < 1 ? 1 : nb);
}

extern "C" int pcb_layernorm_fwd(const void* y, const float* weight, const float* bias, void* out, int64_t C, int64_t rows,
                                 void* stream) {
  PCB_CHECK_ARG(y && weight && bias && out, "pcb_layernorm_fwd: null argument");
  PCB_CHECK_ARG(C > 0 && C % 8 == 0 && rows > 0, "pcb_layernorm_fwd: C must be a positive multiple of 8 (got %lld)", (long long)C);
  const int lpr = ln_lanes_per_row(C);
  const int cpl = lpr > 0 ? (int)(C >> 3) / lpr : 0;
  if (lpr == 0 || cpl > LN_MAXCH) {
    set_error("pcb_layernorm_fwd: C=%lld is not supported (C/8 must be a power of two, C <= %d)", (long long)C, 32 * 8 * LN_MAXCH);
    return PCB_ERR_UNSUPPORTED;
  }
#define PCB_LN_FWD(K) layernorm_kernel<false, K><<<ln_grid(rows, lpr), 256, 0, (cudaStream_t)stream>>>( \
      (const uint4*)y, nullptr, weight, bias, (uint4*)out, nullptr, nullptr, (int)C, rows, lpr)
  if (cpl == 1) PCB_LN_FWD(1); else if (cpl == 2) PCB_LN_FWD(2); else if (cpl == 4) PCB_LN_FWD(4); else PCB_LN_FWD(8);
#undef PCB_LN_FWD
  PCB_CHECK_LAUNCH("pcb_layernorm_fwd");
  return PCB_OK;
}

extern "C" int pcb_layernorm_bwd(const void* g, const void* y, const float* weight, void* dy, double* dweight, double* dbias,
                                 int64_t C, int64_t rows, void* stream) {
  PCB_CHECK_ARG(g && y && weight && dy && dweight && dbias, "pcb_layernorm_bwd: null argument");
  PCB_CHECK_ARG(C > 0 && C % 8 == 0 && rows > 0, "pcb_layernorm_bwd: C must be a positive multiple of 8 (got %lld)", (long long)C);
  const int lpr = ln_lanes_per_row(C);
  const int cpl = lpr > 0 ? (int)(C >> 3) / lpr : 0;
  if (lpr == 0 || cpl > LN_MAXCH) {
    set_error("pcb_layernorm_bwd: C=%lld is not supported (C/8 must be a power of two, C <= %d)", (long long)C, 32 * 8 * LN_MAXCH);
    return PCB_ERR_UNSUPPORTED;
  }
#define PCB_LN_BWD(K) layernorm_kernel<true, K><<<ln_grid(rows, lpr), 256, 2 * C * sizeof(double), (cudaStream_t)stream>>>( \
      (const uint4*)y, (const uint4*)g, weight, nullptr, (uint4*)dy, dweight, dbias, (int)C, rows, lpr)
  if (cpl == 1) PCB_LN_BWD(1); else if (cpl == 2) PCB_LN_BWD(2); else if (cpl == 4) PCB_LN_BWD(4); else PCB_LN_BWD(8);
#undef PCB_LN_BWD
  PCB_CHECK_LAUNCH("pcb_layernorm_bwd");
  return PCB_OK;
}
