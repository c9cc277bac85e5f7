// pcb200 — MedNeXt backward kernels for sm_100a (training path).
//
//   mlp_bwd_kernel   per 128-voxel tile, all on tcgen05 with fp32 accumulators in TMEM:
//                      Hpre = norm(y) W2^T ; dG = dOut W3 ; h = Hpre + b2 ;
//                      Hact = GELU(h) ; dh = dG * GELU'(h) ; dYhat = dh W2
//                    writes Hact, dh (bf16, consumed by the weight-gradient GEMMs), dYhat (bf16) and the
//                    two per-(n,c) GroupNorm-backward sums  S1 = sum g, S2 = sum g*xhat  (float64).
//   tn_gemm_kernel   weight gradients  D[m,n] = sum_v A[v,m] B[v,n]  (reduction over voxels) as a
//                    persistent split-K tcgen05 GEMM with MN-major operands (the [voxel, channel]
//                    tiles are fed untransposed), optional GroupNorm-apply on B, optional all-ones
//                    column (bias gradients), partial sums per CTA + deterministic second stage.
//   pw_kernel        row-gather 1x1 conv  out[r,:] = A[map(r),:] W^T (+b)   (res-conv data gradients,
//                    task-head projections)
//   gn_dy_kernel     dy = rstd*gamma*(g - S1/V - xhat*S2/V)  + per-channel sum(dy) (= conv1 bias grad)
//   dw_wgrad_kernel  depthwise weight gradient  dW[c,tap] = sum_v center[v,c] * neigh[S v - P + tap, c]
//   head_bwd / stem_bwd kernels (CUDA cores; tiny channel counts on one side)
#include "../../include/pcb200.h"
#include <stdlib.h>
#include <string.h>

#include "pcb_common.cuh"

namespace pcb {

enum RowMap { MAP_IDENT = 0, MAP_PLUS1 = 1, MAP_TIMES2 = 2, MAP_TIMES2P1 = 3 };
constexpr int DW_XB_WG = 4;   // centre voxels per thread along W in the depthwise weight-gradient kernel

// source row for tile voxel `v` (linear index in a box of size d0 x d1 x d2) in a tensor whose
// spatial size is (s1, s2 trailing dims): identity / +1 on every axis / *2 / *2+1.
__device__ __forceinline__ int64_t map_row(int kind, int64_t v, int d1, int d2, int s1, int s2) {
  if (kind == MAP_IDENT) return v;
  const int x = (int)(v % d2), y = (int)((v / d2) % d1), z = (int)(v / ((int64_t)d2 * d1));
  if (kind == MAP_PLUS1) return ((int64_t)(z + 1) * s1 + (y + 1)) * s2 + (x + 1);
  if (kind == MAP_TIMES2) return ((int64_t)(2 * z) * s1 + 2 * y) * s2 + 2 * x;
  return ((int64_t)(2 * z + 1) * s1 + (2 * y + 1)) * s2 + (2 * x + 1);
}

__device__ __forceinline__ void stage_rows_k(uint8_t* dst, const uint4* __restrict__ src, int rows, int kc8,
                                             int64_t pitch8, int tid, int nthreads = 128) {
  const uint32_t sbo = kc8 * 128;
  const bool p2 = (kc8 & (kc8 - 1)) == 0;          // kc8 = 12 for H = 96 (MedNeXt-L level 0): plain division
  const int sh = __ffs(kc8) - 1;
  staged_copy<8>(rows * kc8, tid, nthreads,
      [&](int q) { const int r = p2 ? (q >> sh) : (q / kc8), c8 = q - r * kc8; return __ldg(src + (int64_t)r * pitch8 + c8); },
      [&](int q, const uint4& v) {
        const int r = p2 ? (q >> sh) : (q / kc8), c8 = q - r * kc8;
        *reinterpret_cast<uint4*>(dst + (r >> 3) * sbo + c8 * 128 + (r & 7) * 16) = v;
      });
}

// ============================================================================ fused MLP backward (dgrad)
struct MlpBwdArgs {
  const uint4* y; const double* stats; const float* gamma; const float* beta;
  const uint4* w2; const float* b2; const uint4* w3t; const uint4* w2t; const uint4* dout;
  uint4* hact; uint4* dh; uint4* dyhat; double* gstats;
  int y1, y2;            // trailing spatial dims of y
  int o1, o2;            // trailing spatial dims of dout
  int C, H, Co;
  int KC, KCo, N1, Ct;
  int mode;
  int64_t Vy, Vout;
  float inv_count;
};

__global__ void __launch_bounds__(128) mlp_bwd_kernel(MlpBwdArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n = blockIdx.y, ct = blockIdx.z;
  const int64_t tile0 = (int64_t)blockIdx.x * 128;

  uint8_t* sA = smem;                                // [128 x KC]   norm(y) chunk
  uint8_t* sD = sA + 128 * a.KC * 2;                 // [128 x KCo]  dOut chunk
  uint8_t* sW2 = sD + 128 * a.KCo * 2;               // [N1 x KC]
  uint8_t* sW3t = sW2 + a.N1 * a.KC * 2;             // [N1 x KCo]
  uint8_t* sDh = sW3t + a.N1 * a.KCo * 2;            // [128 x N1]
  uint8_t* sW2t = sDh + 128 * a.N1 * 2;              // [Ct x N1]
  float* sScale = reinterpret_cast<float*>(sW2t + a.Ct * a.N1 * 2);   // [C] gamma*rstd
  float* sShift = sScale + a.C;                      // [C] beta - mean*gamma*rstd
  float* sRstd = sShift + a.C;                       // [C]
  float* sMR = sRstd + a.C;                          // [C] mean*rstd
  double* sG = reinterpret_cast<double*>(sMR + a.C); // [2*Ct] S1,S2 partials
  int64_t* sRowO = reinterpret_cast<int64_t*>(sG + 2 * a.Ct);   // [128] row in dout (or -1)
  uint64_t* bar = reinterpret_cast<uint64_t*>(sRowO + 128);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);

  const uint32_t tmem_cols = tmem_cols_pow2(2 * a.N1 + a.Ct);
  if (warp == 0) tmem_alloc(tmem_slot, tmem_cols);
  if (tid == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  for (int c = tid; c < a.C; c += 128) {
    const double s = a.stats[(int64_t)n * 2 * a.C + c], q = a.stats[(int64_t)n * 2 * a.C + a.C + c];
    const double mean = s * (double)a.inv_count;
    double var = q * (double)a.inv_count - mean * mean;
    if (var < 0.0) var = 0.0;
    const float rstd = (float)(1.0 / sqrt(var + 1e-5));
    const float g = a.gamma[c] * rstd;
    sScale[c] = g; sShift[c] = a.beta[c] - (float)mean * g; sRstd[c] = rstd; sMR[c] = (float)mean * rstd;
  }
  for (int i = tid; i < 2 * a.Ct; i += 128) sG[i] = 0.0;
  {
    const int64_t p = tile0 + tid;
    sRowO[tid] = p < a.Vy ? map_row(a.mode == PCB_DW_UP ? MAP_PLUS1 : MAP_IDENT, p, a.y1, a.y2, a.o1, a.o2) : -1;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t acc1 = tmem_base, accG = tmem_base + a.N1, accD = tmem_base + 2 * a.N1;
  uint32_t ph = 0;
  const int nkc = a.C / a.KC, nko = a.Co / a.KCo, nhc = a.H / a.N1;
  const int kc8 = a.KC >> 3, ko8 = a.KCo >> 3, n18 = a.N1 >> 3;
  const uint32_t idescH = umma_idesc_bf16(128, a.N1, 0, 0), idescD = umma_idesc_bf16(128, a.Ct, 0, 0);
  const uint4* yn = a.y + (int64_t)n * a.Vy * (a.C >> 3);
  const uint4* dn = a.dout + (int64_t)n * a.Vout * (a.Co >> 3);
  const int64_t prow = tile0 + tid;
  const bool row_ok = prow < a.Vy;

  for (int hc = 0; hc < nhc; ++hc) {
    // ---- Hpre = norm(y) * W2[hc]^T
    for (int kc = 0; kc < nkc; ++kc) {
      if (!(nkc == 1 && hc > 0)) {
        const uint32_t sbo = kc8 * 128;
        staged_copy<8>(128 * kc8, tid, 128,
            [&](int q) {
              const int r = q >> __ffs(kc8) - 1, c8 = q & (kc8 - 1);
              return tile0 + r < a.Vy ? __ldg(yn + (tile0 + r) * (a.C >> 3) + kc * kc8 + c8) : make_uint4(0, 0, 0, 0);
            },
            [&](int q, const uint4& raw) {
              const int r = q >> __ffs(kc8) - 1, c8 = q & (kc8 - 1);
              uint4 v = make_uint4(0, 0, 0, 0);
              if (tile0 + r < a.Vy) {
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
      stage_rows_k(sW2, a.w2 + (int64_t)hc * a.N1 * (a.C >> 3) + kc * kc8, a.N1, kc8, a.C >> 3, tid);
      fence_proxy_async_smem();
      __syncthreads();
      if (tid == 0) {
        tc_fence_after();
        const uint64_t ad = umma_desc(smem_u32(sA), 128, kc8 * 128), bd = umma_desc(smem_u32(sW2), 128, kc8 * 128);
        for (int k = 0; k < a.KC / 16; ++k)
          umma_bf16(acc1, ad + (uint64_t)(k * 16), bd + (uint64_t)(k * 16), idescH, (kc > 0 || k > 0) ? 1u : 0u);
        tc_commit(bar);
      }
      mbar_wait(bar, ph); ph ^= 1;
    }
    // ---- dG = dOut * W3[:, hc]   (B operand = W3^T rows hc*N1.., K = Co)
    for (int kc = 0; kc < nko; ++kc) {
      if (!(nko == 1 && hc > 0)) {
        const uint32_t sbo = ko8 * 128;
        staged_copy<8>(128 * ko8, tid, 128,
            [&](int q) {
              const int r = q >> __ffs(ko8) - 1, c8 = q & (ko8 - 1);
              const int64_t ro = sRowO[r];
              return ro >= 0 ? __ldg(dn + ro * (a.Co >> 3) + kc * ko8 + c8) : make_uint4(0, 0, 0, 0);
            },
            [&](int q, const uint4& v) {
              const int r = q >> __ffs(ko8) - 1, c8 = q & (ko8 - 1);
              *reinterpret_cast<uint4*>(sD + (r >> 3) * sbo + c8 * 128 + (r & 7) * 16) = v;
            });
      }
      stage_rows_k(sW3t, a.w3t + (int64_t)hc * a.N1 * (a.Co >> 3) + kc * ko8, a.N1, ko8, a.Co >> 3, tid);
      fence_proxy_async_smem();
      __syncthreads();
      if (tid == 0) {
        tc_fence_after();
        const uint64_t ad = umma_desc(smem_u32(sD), 128, ko8 * 128), bd = umma_desc(smem_u32(sW3t), 128, ko8 * 128);
        for (int k = 0; k < a.KCo / 16; ++k)
          umma_bf16(accG, ad + (uint64_t)(k * 16), bd + (uint64_t)(k * 16), idescH, (kc > 0 || k > 0) ? 1u : 0u);
        tc_commit(bar);
      }
      mbar_wait(bar, ph); ph ^= 1;
    }
    tc_fence_after();
    // ---- W2^T[ct tile, hc] for the dYhat GEMM
    stage_rows_k(sW2t, a.w2t + (int64_t)ct * a.Ct * (a.H >> 3) + hc * n18, a.Ct, n18, a.H >> 3, tid);
    // ---- epilogue: Hact = GELU(h), dh = dG * GELU'(h)
    {
      const uint32_t t1 = acc1 + ((uint32_t)(warp * 32) << 16), tg = accG + ((uint32_t)(warp * 32) << 16);
      const uint32_t sbo = n18 * 128;
      uint8_t* dst = sDh + (tid >> 3) * sbo + (tid & 7) * 16;
      const int64_t grow = ((int64_t)n * a.Vy + prow) * (a.H >> 3) + hc * n18;
      for (int c16 = 0; c16 < a.N1 / 16; ++c16) {
        uint32_t v1[16], vg[16];
        tmem_ld16(t1 + c16 * 16, v1);
        tmem_ld16(tg + c16 * 16, vg);
        tmem_ld_wait();
        float ha[16], dhv[16];
        const float* bp = a.b2 + hc * a.N1 + c16 * 16;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float h = __uint_as_float(v1[j]) + __ldg(bp + j);
          float gr;
          gelu_fast_vg(h, ha[j], gr);
          dhv[j] = __uint_as_float(vg[j]) * gr;
        }
        const uint4 d0 = pack8(dhv), d1 = pack8(dhv + 8);
        *reinterpret_cast<uint4*>(dst + (c16 * 2) * 128) = d0;
        *reinterpret_cast<uint4*>(dst + (c16 * 2 + 1) * 128) = d1;
        if (row_ok && ct == 0) {
          a.hact[grow + c16 * 2] = pack8(ha);
          a.hact[grow + c16 * 2 + 1] = pack8(ha + 8);
          a.dh[grow + c16 * 2] = d0;
          a.dh[grow + c16 * 2 + 1] = d1;
        }
      }
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    // ---- dYhat[:, ct tile] += dh * W2[hc, ct tile]
    if (tid == 0) {
      tc_fence_after();
      const uint64_t ad = umma_desc(smem_u32(sDh), 128, n18 * 128), bd = umma_desc(smem_u32(sW2t), 128, n18 * 128);
      for (int k = 0; k < a.N1 / 16; ++k)
        umma_bf16(accD, ad + (uint64_t)(k * 16), bd + (uint64_t)(k * 16), idescD, (hc > 0 || k > 0) ? 1u : 0u);
      tc_commit(bar);
    }
    mbar_wait(bar, ph); ph ^= 1;
  }
  tc_fence_after();
  // ---- final epilogue: g = dYhat -> bf16; S1 += g, S2 += g * xhat
  {
    const uint32_t td = accD + ((uint32_t)(warp * 32) << 16);
    const int64_t yrow = ((int64_t)n * a.Vy + prow) * (a.C >> 3) + ct * (a.Ct >> 3);
    for (int c16 = 0; c16 < a.Ct / 16; ++c16) {
      uint32_t v[16];
      tmem_ld16(td + c16 * 16, v);
      tmem_ld_wait();
      float g[16], gx[16];
      const int c0 = ct * a.Ct + c16 * 16;
      if (row_ok) {
        float yv[16];
        unpack8(__ldg(a.y + yrow + c16 * 2), yv);
        unpack8(__ldg(a.y + yrow + c16 * 2 + 1), yv + 8);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          g[j] = round_bf16(__uint_as_float(v[j]));
          gx[j] = g[j] * fmaf(yv[j], sRstd[c0 + j], -sMR[c0 + j]);
        }
        a.dyhat[yrow + c16 * 2] = pack8(g);
        a.dyhat[yrow + c16 * 2 + 1] = pack8(g + 8);
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) { g[j] = 0.f; gx[j] = 0.f; }
      }
      warp_colsum16(g, lane);
      warp_colsum16(gx, lane);
      if (!(lane & 1)) {
        const int col = c16 * 16 + colsum16_col(lane);
        atomicAdd(&sG[col], (double)g[0]);
        atomicAdd(&sG[a.Ct + col], (double)gx[0]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  for (int i = tid; i < a.Ct; i += 128) {
    atomicAdd(&a.gstats[(int64_t)n * 2 * a.C + ct * a.Ct + i], sG[i]);
    atomicAdd(&a.gstats[(int64_t)n * 2 * a.C + a.C + ct * a.Ct + i], sG[a.Ct + i]);
  }
  if (warp == 0) tmem_dealloc(tmem_base, tmem_cols);
}

// ============================================================================ persistent fused MLP backward
// Levels 0/1 (3H + C + Co <= 512 TMEM columns): one kernel per block does the data gradient AND both
// pointwise weight gradients.  Per 128-voxel tile:
//   G1 Hpre = Yhat W2^T      G2 dG = dOut W3          (K-major operands, fp32 in TMEM)
//   E1 Hact = GELU(Hpre+b2), dh = dG*GELU'            (bf16 tiles in shared memory only)
//   G3 dYhat = dh W2         -> E2: bf16 to HBM + GroupNorm-backward sums
//   G4 dW3^T[h,co] += Hact^T dOut     G5 dW2^T[c,h] += Yhat^T dh
// G4/G5 read the SAME shared tiles through MN-major descriptors (LBO/SBO swapped), reduce over the 128
// voxels of the tile, and keep accumulating into TMEM across all tiles of the persistent CTA.  The A tiles
// of G5 carries one extra 8x16 B core matrix per row group holding [1,0,..,0]: in the MN-major view that is
// an all-ones M-row, so row C of dW2^T is db2 = sum dh — that bias
// gradient comes out of the tensor cores for free; db3 = sum dOut is a column sum of the staged dOut tile.
// Hact / dh never touch HBM, no split-K GEMM launches.
struct MlpBwdFusedArgs {
  int split;                    // ws2, one accumulator set: both epilogue groups share every tile (column split)
  MlpBwdArgs m;
  float* part3;      // [P][129][Co]  partial dW3^T rows 0..H-1, row 128 = partial db3
  float* part2;      // [P][128][H]   partial dW2^T (+ row C = db2)
  int N;
  int64_t tps, ntiles;
  uint32_t dm2, dm1; int ds2, ds1;   // exact division by y2 / y1 (mlp_bwd_ws2_kernel, UP-mode dOut row map)
  int ld16;                          // mlp_bwd_ws2_kernel: 16 loads in flight per loader lane (PCB_BWD_LD16=1)
};

constexpr int BF_PF = 8;          // prefetch depth: (C/8 + Co/8) * 128 / 256 <= 8 chunks per thread
// NT threads = NT/128 warpgroups: every warpgroup covers all 128 TMEM lanes (warp w reads lane quarter w%4), the
// warpgroups split the columns.  NT = 256 where TMEM lets two CTAs share an SM (<= 256 columns), NT = 512 where only
// one fits (up_0, down_0, level 1: 16 warps instead of 8 on the SM; measured IPC of the 8-warp CTA: 1.1).

template <int NT>
__global__ void __launch_bounds__(NT, NT == 256 ? 2 : 1) mlp_bwd_fused_kernel(MlpBwdFusedArgs fa) {
  constexpr int BF_THREADS = NT, NPART = NT / 128;
  const MlpBwdArgs& a = fa.m;
  extern __shared__ __align__(128) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row = tid & 127, part = tid >> 7, wq = warp & 3;   // tile row, column part (warpgroup), TMEM lane quarter
  const int c8n = a.C >> 3, h8n = a.H >> 3, o8n = a.Co >> 3;
  const uint32_t pitchA = (c8n + 1) * 128, pitchH = h8n * 128, pitchD = o8n * 128, pitchDh = h8n * 128;
  uint8_t* sW2 = smem;                                 // [H x C]   K-major (B of G1)
  uint8_t* sW3t = sW2 + a.H * a.C * 2;                 // [H x Co]  K-major (B of G2)
  uint8_t* sW2t = sW3t + a.H * a.Co * 2;               // [C x H]   K-major (B of G3)
  uint8_t* sA = sW2t + a.C * a.H * 2;                  // [128 x C]  + ones core matrix per row group
  uint8_t* sD = sA + 16 * pitchA;                      // [128 x Co]
  uint8_t* sH = sD + 16 * pitchD;                      // [128 x H]
  uint8_t* sDh = sH + 16 * pitchH;                     // [128 x H]
  uint8_t* sY = sDh + 16 * pitchDh;                    // [128][C*2+16] raw y rows (x-hat in E2)
  uint8_t* sTail = sY + 128 * (a.C * 2 + 16);          // 2 KB of finite padding read by the M>valid rows
  float* sScale = reinterpret_cast<float*>(sTail + 2048);   // [N][C] gamma*rstd
  float* sShift = sScale + fa.N * a.C;                 // [N][C]
  float* sRstd = sShift + fa.N * a.C;                  // [N][C]
  float* sMR = sRstd + fa.N * a.C;                     // [N][C] mean*rstd
  float* sB2 = sMR + fa.N * a.C;                       // [H]
  double* sG = reinterpret_cast<double*>(sB2 + a.H);   // [2*C] S1,S2 partials of the current sample
  int* sRowO = reinterpret_cast<int*>(sG + 2 * a.C);   // [128]
  uint64_t* bar = reinterpret_cast<uint64_t*>(sRowO + 128);   // [0] G1+G2 / G3 done, [1] G4+G5 done
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 2);

  const uint32_t ncols = 3 * a.H + a.C + a.Co;
  const uint32_t tmem_cols = tmem_cols_pow2(ncols);
  if (warp == 0) tmem_alloc(tmem_slot, tmem_cols);
  if (tid == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); fence_mbar_init(); }
  // resident weights (K-major), biases, per-sample GroupNorm constants, ones core matrices, finite tail
  stage_rows_k(sW2, a.w2, a.H, c8n, c8n, tid, BF_THREADS);
  stage_rows_k(sW3t, a.w3t, a.H, o8n, o8n, tid, BF_THREADS);
  stage_rows_k(sW2t, a.w2t, a.C, h8n, h8n, tid, BF_THREADS);
  for (int i = tid; i < a.H; i += BF_THREADS) sB2[i] = a.b2[i];
  for (int i = tid; i < fa.N * a.C; i += BF_THREADS) {
    const int n = i / a.C, c = i - n * a.C;
    const double sm = a.stats[(int64_t)n * 2 * a.C + c], q = a.stats[(int64_t)n * 2 * a.C + a.C + c];
    const double mean = sm * (double)a.inv_count;
    double var = q * (double)a.inv_count - mean * mean;
    if (var < 0.0) var = 0.0;
    const float rstd = (float)(1.0 / sqrt(var + 1e-5));
    const float g = a.gamma[c] * rstd;
    sScale[i] = g; sShift[i] = a.beta[c] - (float)mean * g; sRstd[i] = rstd; sMR[i] = (float)mean * rstd;
  }
  {
    const uint4 one = make_uint4(0x3F80u, 0, 0, 0);   // bf16 [1,0,0,0,0,0,0,0]
    // 16 row groups x 8 voxels: ones core matrix of sA (chunk index c8n)
    if (tid < 128) {
      *reinterpret_cast<uint4*>(sA + (tid >> 3) * pitchA + c8n * 128 + (tid & 7) * 16) = one;
      *reinterpret_cast<uint4*>(sTail + tid * 16) = make_uint4(0, 0, 0, 0);
    }
  }
  for (int i = tid; i < 2 * a.C; i += BF_THREADS) sG[i] = 0.0;
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t acc1 = tmem_base, accG = tmem_base + a.H, accD = tmem_base + 2 * a.H;
  const uint32_t accW3 = accD + a.C, accW2 = accW3 + a.Co;
  const uint32_t idescH = umma_idesc_bf16(128, a.H, 0, 0), idescD = umma_idesc_bf16(128, a.C, 0, 0);
  const uint32_t idescW3 = umma_idesc_bf16(128, a.Co, 1, 1), idescW2 = umma_idesc_bf16(128, a.H, 1, 1);
  uint32_t ph0 = 0, ph1 = 0;
  int cur_n = -1;
  int64_t it = 0;
  float db3acc = 0.f;   // thread t < Co: running sum_v dOut[v, t]
  const int csh = __ffs(c8n) - 1, osh = __ffs(o8n) - 1;
  const int nchunk_y = 128 * c8n, nchunk_o = 128 * o8n;
  const int ypitch = a.C * 2 + 16;
  uint4 pre[BF_PF];     // next tile's y / dOut chunks, in flight while the current tile is processed
  auto prefetch = [&](int64_t gt) {
    const int pn = (int)(gt / fa.tps);
    const int pt0 = (int)((gt - (int64_t)pn * fa.tps) * 128);
    const uint4* pyn = a.y + (int64_t)pn * a.Vy * c8n;
    const uint4* pdn = a.dout + (int64_t)pn * a.Vout * o8n;
#pragma unroll
    for (int b = 0; b < BF_PF; ++b) {
      const int q = tid + b * BF_THREADS;
      pre[b] = make_uint4(0, 0, 0, 0);
      if (q < nchunk_y) {
        const int r = q >> csh, c8 = q & (c8n - 1);
        if (pt0 + r < (int)a.Vy) pre[b] = __ldg(pyn + (int64_t)(pt0 + r) * c8n + c8);
      } else if (q < nchunk_y + nchunk_o) {
        const int qo = q - nchunk_y, r = qo >> osh, c8 = qo & (o8n - 1);
        if (pt0 + r < (int)a.Vy) {
          const int64_t ro = map_row(a.mode == PCB_DW_UP ? MAP_PLUS1 : MAP_IDENT, pt0 + r, a.y1, a.y2, a.o1, a.o2);
          pre[b] = __ldg(pdn + ro * o8n + c8);
        }
      }
    }
  };
  if ((int64_t)blockIdx.x < fa.ntiles) prefetch(blockIdx.x);
  for (int64_t g = blockIdx.x; g < fa.ntiles; g += gridDim.x, ++it) {
    const int n = (int)(g / fa.tps);
    const int tile0 = (int)((g - (int64_t)n * fa.tps) * 128);
    if (n != cur_n) {   // flush the GroupNorm-backward partial sums of the previous sample
      if (cur_n >= 0) {
        __syncthreads();
        for (int i = tid; i < a.C; i += BF_THREADS) {
          atomicAdd(&a.gstats[(int64_t)cur_n * 2 * a.C + i], sG[i]);
          atomicAdd(&a.gstats[(int64_t)cur_n * 2 * a.C + a.C + i], sG[a.C + i]);
          sG[i] = 0.0; sG[a.C + i] = 0.0;
        }
        __syncthreads();
      }
      cur_n = n;
    }
    const float* sc = sScale + n * a.C; const float* sh = sShift + n * a.C;
    if (it >= 1) { mbar_wait(&bar[1], ph1); ph1 ^= 1; }   // G4/G5 of the previous tile have consumed the tiles
    __syncthreads();                                      // every thread is done with E2's reads of sY
    // the tile's global loads were issued one iteration ago (software prefetch): convert + stage them now
#pragma unroll
    for (int b = 0; b < BF_PF; ++b) {
      const int q = tid + b * BF_THREADS;
      if (q < nchunk_y) {
        const int r = q >> csh, c8 = q & (c8n - 1);
        uint4 v = make_uint4(0, 0, 0, 0);
        if (tile0 + r < (int)a.Vy) {
          float f[8];
          unpack8(pre[b], f);
#pragma unroll
          for (int jj = 0; jj < 8; ++jj) f[jj] = fmaf(f[jj], sc[c8 * 8 + jj], sh[c8 * 8 + jj]);
          v = pack8(f);
        }
        *reinterpret_cast<uint4*>(sA + (r >> 3) * pitchA + c8 * 128 + (r & 7) * 16) = v;
        *reinterpret_cast<uint4*>(sY + r * ypitch + c8 * 16) = pre[b];
      } else if (q < nchunk_y + nchunk_o) {
        const int qo = q - nchunk_y, r = qo >> osh, c8 = qo & (o8n - 1);
        *reinterpret_cast<uint4*>(sD + (r >> 3) * pitchD + c8 * 128 + (r & 7) * 16) = pre[b];
      }
    }
    fence_proxy_async_smem();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      const uint64_t dA = umma_desc(smem_u32(sA), 128, pitchA), dW2 = umma_desc(smem_u32(sW2), 128, c8n * 128);
      for (int k = 0; k < a.C / 16; ++k) umma_bf16(acc1, dA + (uint64_t)(k * 16), dW2 + (uint64_t)(k * 16), idescH, k > 0 ? 1u : 0u);
      const uint64_t dD = umma_desc(smem_u32(sD), 128, pitchD), dW3 = umma_desc(smem_u32(sW3t), 128, o8n * 128);
      for (int k = 0; k < a.Co / 16; ++k) umma_bf16(accG, dD + (uint64_t)(k * 16), dW3 + (uint64_t)(k * 16), idescH, k > 0 ? 1u : 0u);
      tc_commit(&bar[0]);
    }
    if (tid < a.Co) {   // conv3 bias gradient: column sums of the staged dOut tile (overlaps the MMAs)
      const uint8_t* col = sD + (tid >> 3) * 128 + (tid & 7) * 2;
      float sacc = 0.f;
#pragma unroll 8
      for (int r = 0; r < 128; ++r)
        sacc += __uint_as_float((uint32_t)(*reinterpret_cast<const uint16_t*>(col + (r >> 3) * pitchD + (r & 7) * 16)) << 16);
      db3acc += sacc;
    }
    mbar_wait(&bar[0], ph0); ph0 ^= 1;
    tc_fence_after();
    // ---- E1: Hact -> sH, dh -> sDh   (warpgroup `part` handles its share of the hidden columns)
    {
      const uint32_t t1 = acc1 + ((uint32_t)(wq * 32) << 16), tg = accG + ((uint32_t)(wq * 32) << 16);
      uint8_t* dH = sH + (row >> 3) * pitchH + (row & 7) * 16;
      uint8_t* dDh = sDh + (row >> 3) * pitchDh + (row & 7) * 16;
      const int nch = a.H / 16, c_lo = (part * nch) / NPART, c_hi = ((part + 1) * nch) / NPART;
      for (int c16 = c_lo; c16 < c_hi; ++c16) {
        uint32_t v1[16], vg[16];
        tmem_ld16(t1 + c16 * 16, v1);
        tmem_ld16(tg + c16 * 16, vg);
        tmem_ld_wait();
        uint32_t hw[8], dw[8];
        const float4* bp = reinterpret_cast<const float4*>(sB2 + c16 * 16);
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) {
          const float4 b = bp[j4];
          uint64_t va, ga, vb, gb;
          gelu_fast_vg2(add2(pk2(__uint_as_float(v1[4 * j4]), __uint_as_float(v1[4 * j4 + 1])), pk2(b.x, b.y)), va, ga);
          gelu_fast_vg2(add2(pk2(__uint_as_float(v1[4 * j4 + 2]), __uint_as_float(v1[4 * j4 + 3])), pk2(b.z, b.w)), vb, gb);
          ga = mul2(ga, pk2(__uint_as_float(vg[4 * j4]), __uint_as_float(vg[4 * j4 + 1])));
          gb = mul2(gb, pk2(__uint_as_float(vg[4 * j4 + 2]), __uint_as_float(vg[4 * j4 + 3])));
          float e0, e1;
          upk2(va, e0, e1); hw[2 * j4] = pack_bf16(e0, e1);
          upk2(vb, e0, e1); hw[2 * j4 + 1] = pack_bf16(e0, e1);
          upk2(ga, e0, e1); dw[2 * j4] = pack_bf16(e0, e1);
          upk2(gb, e0, e1); dw[2 * j4 + 1] = pack_bf16(e0, e1);
        }
        *reinterpret_cast<uint4*>(dH + (c16 * 2) * 128) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
        *reinterpret_cast<uint4*>(dH + (c16 * 2 + 1) * 128) = make_uint4(hw[4], hw[5], hw[6], hw[7]);
        *reinterpret_cast<uint4*>(dDh + (c16 * 2) * 128) = make_uint4(dw[0], dw[1], dw[2], dw[3]);
        *reinterpret_cast<uint4*>(dDh + (c16 * 2 + 1) * 128) = make_uint4(dw[4], dw[5], dw[6], dw[7]);
      }
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      // G3: dYhat = dh W2
      const uint64_t dDhK = umma_desc(smem_u32(sDh), 128, pitchDh), dW2t = umma_desc(smem_u32(sW2t), 128, h8n * 128);
      for (int k = 0; k < a.H / 16; ++k) umma_bf16(accD, dDhK + (uint64_t)(k * 16), dW2t + (uint64_t)(k * 16), idescD, k > 0 ? 1u : 0u);
      tc_commit(&bar[0]);
      // G4: dW3^T[h (+ones row), co] += Hact^T dOut ; G5: dW2^T[c (+ones row), h] += Yhat^T dh   (K = 128 voxels)
      const uint64_t aH = umma_desc(smem_u32(sH), pitchH, 128), bD = umma_desc(smem_u32(sD), pitchD, 128);
      const uint64_t aA = umma_desc(smem_u32(sA), pitchA, 128), bDh = umma_desc(smem_u32(sDh), pitchDh, 128);
      for (int k = 0; k < 8; ++k)
        umma_bf16(accW3, aH + (uint64_t)(k * 2 * (pitchH >> 4)), bD + (uint64_t)(k * 2 * (pitchD >> 4)), idescW3, (it > 0 || k > 0) ? 1u : 0u);
      for (int k = 0; k < 8; ++k)
        umma_bf16(accW2, aA + (uint64_t)(k * 2 * (pitchA >> 4)), bDh + (uint64_t)(k * 2 * (pitchDh >> 4)), idescW2, (it > 0 || k > 0) ? 1u : 0u);
      tc_commit(&bar[1]);
    }
    if (g + gridDim.x < fa.ntiles) prefetch(g + gridDim.x);   // next tile's loads fly during E2
    mbar_wait(&bar[0], ph0); ph0 ^= 1;
    tc_fence_after();
    // ---- E2: g = dYhat -> bf16 ; S1 += g ; S2 += g*xhat
    {
      const int prow = tile0 + row;
      const bool row_ok = prow < (int)a.Vy;
      const uint32_t td = accD + ((uint32_t)(wq * 32) << 16);
      const int ncd = a.C / 16, d_lo = (part * ncd) / NPART, d_hi = ((part + 1) * ncd) / NPART;
      const int64_t yrow = ((int64_t)n * a.Vy + prow) * c8n;
      const float* rs = sRstd + n * a.C; const float* mr = sMR + n * a.C;
      for (int c16 = d_lo; c16 < d_hi; ++c16) {
        uint32_t v[16];
        tmem_ld16(td + c16 * 16, v);
        const uint4 y0 = *reinterpret_cast<const uint4*>(sY + row * ypitch + c16 * 32);
        const uint4 y1 = *reinterpret_cast<const uint4*>(sY + row * ypitch + c16 * 32 + 16);
        tmem_ld_wait();
        float gq[16], gx[16];
        if (row_ok) {
          float yv[16];
          unpack8(y0, yv);
          unpack8(y1, yv + 8);
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            gq[j] = round_bf16(__uint_as_float(v[j]));
            gx[j] = gq[j] * fmaf(yv[j], rs[c16 * 16 + j], -mr[c16 * 16 + j]);
          }
          a.dyhat[yrow + c16 * 2] = pack8(gq);
          a.dyhat[yrow + c16 * 2 + 1] = pack8(gq + 8);
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) { gq[j] = 0.f; gx[j] = 0.f; }
        }
        warp_colsum16(gq, lane);
        warp_colsum16(gx, lane);
        if (!(lane & 1)) {
          const int col = c16 * 16 + colsum16_col(lane);
          atomicAdd(&sG[col], (double)gq[0]);
          atomicAdd(&sG[a.C + col], (double)gx[0]);
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (cur_n >= 0) {
    for (int i = tid; i < a.C; i += BF_THREADS) {
      atomicAdd(&a.gstats[(int64_t)cur_n * 2 * a.C + i], sG[i]);
      atomicAdd(&a.gstats[(int64_t)cur_n * 2 * a.C + a.C + i], sG[a.C + i]);
    }
  }
  if (it >= 1) mbar_wait(&bar[1], ph1);
  tc_fence_after();
  // ---- weight-gradient partials of this CTA: dW3^T rows 0..H (row H = db3), dW2^T rows 0..C (row C = db2)
  {
    const uint32_t lo = (uint32_t)(wq * 32) << 16;
    float* p3 = fa.part3 + ((int64_t)blockIdx.x * 129 + row) * a.Co;
    if (tid < a.Co) fa.part3[((int64_t)blockIdx.x * 129 + 128) * a.Co + tid] = db3acc;
    for (int c16 = part; c16 < a.Co / 16; c16 += NPART) {
      uint32_t v[16];
      tmem_ld16(accW3 + lo + c16 * 16, v);
      tmem_ld_wait();
      if (row < a.H) {
#pragma unroll
        for (int j = 0; j < 16; ++j) p3[c16 * 16 + j] = it > 0 ? __uint_as_float(v[j]) : 0.f;
      }
    }
    float* p2 = fa.part2 + ((int64_t)blockIdx.x * 128 + row) * a.H;
    for (int c16 = part; c16 < a.H / 16; c16 += NPART) {
      uint32_t v[16];
      tmem_ld16(accW2 + lo + c16 * 16, v);
      tmem_ld_wait();
      if (row <= a.C) {
#pragma unroll
        for (int j = 0; j < 16; ++j) p2[c16 * 16 + j] = it > 0 ? __uint_as_float(v[j]) : 0.f;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, tmem_cols);
}

// ============================================================================ warp-specialised fused MLP backward
// OPT-IN (PCB_BWD_WS=1): same math, operands and partial-sum layout as mlp_bwd_fused_kernel, reorganised the way the forward
// kernel is so that no role ever waits on a block-wide barrier (the 256-thread kernel spends its time in `wait` / `barrier`
// stalls: five __syncthreads per tile and every MMA latency exposed).  One CTA per SM, 13 warps:
//   4 loader warps   warp w owns stage w and the tiles it == w (mod 4): y -> GroupNorm-apply -> sA[w] (K-major, + the
//                    all-ones core matrix per row group), dOut -> sD[w]
//   1 MMA thread     G1 Hpre -> acc1[it&1], G2 dG -> accG[it&1]; one tile behind: G3 dYhat -> accD, G4 / G5 weight gradients
//   2 x 4 epilogue   group eg owns the tiles it == eg (mod 2): E1 (GELU / GELU' -> sH[eg], sDh[eg]) and E2 (dYhat -> HBM,
//                    GroupNorm-backward sums) of its tile; while it waits for G3 the other group runs its E1
// TMEM: 2 x (acc1 + accG) + 2 x accD + accW3 + accW2 = 4H + 2C + Co + H columns (416 at level 0).  Level-0 shape only
// (C = 32, Co = 32, H = 64, SAME / DOWN rows): the level-1 and up_0 shapes do not fit double-buffered accumulators.
constexpr int WS_LOAD = 4, WS_EPI = 8, WS_THREADS = 32 * (WS_LOAD + WS_EPI + 1);
constexpr int WS0_THREADS = 32 * (WS_LOAD + WS_EPI + 3);   // level-0 kernel: three MMA issuer warps (two groups + weight gradients)

// One warp stages a [128 x 32] bf16 tile (4 chunks of 16 B per row) into the K-major canonical layout with row-group
// pitch `pitch` (same lane mapping as mf_stage_tile in mednext_fwd.cu): 8 loads in flight per lane, conflict-free stores.
template <bool NORM>
__device__ __forceinline__ void ws_stage_tile32(uint8_t* __restrict__ dst, uint32_t pitch, const uint4* __restrict__ src,
                                                int row0, int nvalid, const float* __restrict__ sc,
                                                const float* __restrict__ sh, int lane) {
  const int rl = lane & 7, cs = lane >> 3;      // row inside the core matrix, chunk
  uint64_t ps[4], pt[4];
  if (NORM) {
    const float4* sp = reinterpret_cast<const float4*>(sc + cs * 8);
    const float4* tp = reinterpret_cast<const float4*>(sh + cs * 8);
    const float4 s0 = sp[0], s1 = sp[1], t0 = tp[0], t1 = tp[1];
    ps[0] = pk2(s0.x, s0.y); ps[1] = pk2(s0.z, s0.w); ps[2] = pk2(s1.x, s1.y); ps[3] = pk2(s1.z, s1.w);
    pt[0] = pk2(t0.x, t0.y); pt[1] = pk2(t0.z, t0.w); pt[2] = pk2(t1.x, t1.y); pt[3] = pk2(t1.z, t1.w);
  }
  uint8_t* dl = dst + cs * 128 + rl * 16;
#pragma unroll 1
  for (int b = 0; b < 2; ++b) {
    uint4 v[8];
    uint32_t ok = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int r = (b * 8 + k) * 8 + rl;
      v[k] = make_uint4(0, 0, 0, 0);
      if (r < nvalid) { v[k] = __ldg(src + (int64_t)(row0 + r) * 4 + cs); ok |= 1u << k; }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      uint4 o = v[k];
      if (NORM) {
        const bool live = (ok >> k) & 1u;
        float a0, a1, a2, a3, a4, a5, a6, a7;
        upk2(fma2(pk2(bf16_lo(o.x), bf16_hi(o.x)), ps[0], pt[0]), a0, a1);
        upk2(fma2(pk2(bf16_lo(o.y), bf16_hi(o.y)), ps[1], pt[1]), a2, a3);
        upk2(fma2(pk2(bf16_lo(o.z), bf16_hi(o.z)), ps[2], pt[2]), a4, a5);
        upk2(fma2(pk2(bf16_lo(o.w), bf16_hi(o.w)), ps[3], pt[3]), a6, a7);
        o.x = live ? pack_bf16(a0, a1) : 0u; o.y = live ? pack_bf16(a2, a3) : 0u;
        o.z = live ? pack_bf16(a4, a5) : 0u; o.w = live ? pack_bf16(a6, a7) : 0u;
      }
      *reinterpret_cast<uint4*>(dl + (b * 8 + k) * pitch) = o;
    }
  }
}

__global__ void __launch_bounds__(WS0_THREADS, 1) mlp_bwd_ws_kernel(MlpBwdFusedArgs fa) {
  const MlpBwdArgs& a = fa.m;
  extern __shared__ __align__(128) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int c8n = a.C >> 3, h8n = a.H >> 3, o8n = a.Co >> 3;      // 4, 8, 4
  const uint32_t pitchA = (c8n + 1) * 128, pitchD = o8n * 128, pitchH = h8n * 128;
  const uint32_t stageA = 16 * pitchA, stageD = 16 * pitchD, stageH = 16 * pitchH;
  uint8_t* sW2 = smem;                                 // [H x C]   K-major (B of G1)
  uint8_t* sW3t = sW2 + a.H * a.C * 2;                 // [H x Co]  K-major (B of G2)
  uint8_t* sW2t = sW3t + a.H * a.Co * 2;               // [C x H]   K-major (B of G3)
  uint8_t* sA = sW2t + a.C * a.H * 2;                  // 4 x [128 x C] + ones core matrix per row group
  uint8_t* sD = sA + 4 * stageA;                       // 4 x [128 x Co]
  uint8_t* sH = sD + 4 * stageD;                       // 2 x [128 x H]
  uint8_t* sDh = sH + 2 * stageH;                      // 2 x [128 x H]
  uint8_t* sTail = sDh + 2 * stageH;                   // 2 KB finite padding behind the last MN-major operand
  float* sScale = reinterpret_cast<float*>(sTail + 2048);   // [N][C] gamma*rstd
  float* sShift = sScale + fa.N * a.C;                 // [N][C]
  float* sRstd = sShift + fa.N * a.C;                  // [N][C]
  float* sMR = sRstd + fa.N * a.C;                     // [N][C] mean*rstd
  float* sB2 = sMR + fa.N * a.C;                       // [H]
  float* sDb3 = sB2 + a.H;                             // [2][Co] per-group conv3 bias-gradient partials
  double* sG = reinterpret_cast<double*>(sDb3 + 2 * a.Co);  // [N][2C] S1, S2 per sample
  uint64_t* bars = reinterpret_cast<uint64_t*>(sG + fa.N * 2 * a.C);
  uint64_t* a_full = bars;          // [4] loaders -> MMA / epilogue
  uint64_t* a_empty = bars + 4;     // [4] MMA (G4/G5 retired) -> loaders
  uint64_t* hp_full = bars + 8;     // [2] MMA (G1/G2) -> E1
  uint64_t* e1_done = bars + 10;    // [2] E1 -> MMA
  uint64_t* d_full = bars + 12;     // [2 groups][2 buffers] MMA (G3) -> E2
  uint64_t* d_empty = bars + 16;    // [2 groups][2 buffers] E2 -> MMA
  uint64_t* h_free = bars + 20;     // [2] MMA (G3/G4/G5 retired) -> E1
  uint64_t* w_done = bars + 22;     // every MMA of the CTA retired
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 23);

  const uint32_t tmem_cols = tmem_cols_pow2((uint32_t)(5 * a.H + 4 * a.C + a.Co));
  if (warp == WS_LOAD + WS_EPI) tmem_alloc(tmem_slot, tmem_cols);
  if (tid == 0) {
    for (int i = 0; i < 4; ++i) { mbar_init(&a_full[i], 32); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&hp_full[i], 1); mbar_init(&e1_done[i], 128); mbar_init(&h_free[i], 1); }
    for (int i = 0; i < 4; ++i) { mbar_init(&d_full[i], 1); mbar_init(&d_empty[i], 128); }
    mbar_init(w_done, 1);
    fence_mbar_init();
  }
  stage_rows_k(sW2, a.w2, a.H, c8n, c8n, tid, WS0_THREADS);
  stage_rows_k(sW3t, a.w3t, a.H, o8n, o8n, tid, WS0_THREADS);
  stage_rows_k(sW2t, a.w2t, a.C, h8n, h8n, tid, WS0_THREADS);
  for (int i = tid; i < a.H; i += WS0_THREADS) sB2[i] = a.b2[i];
  for (int i = tid; i < fa.N * a.C; i += WS0_THREADS) {
    const int n = i / a.C, c = i - n * a.C;
    const double sm = a.stats[(int64_t)n * 2 * a.C + c], q = a.stats[(int64_t)n * 2 * a.C + a.C + c];
    const double mean = sm * (double)a.inv_count;
    double var = q * (double)a.inv_count - mean * mean;
    if (var < 0.0) var = 0.0;
    const float rstd = (float)(1.0 / sqrt(var + 1e-5));
    const float g = a.gamma[c] * rstd;
    sScale[i] = g; sShift[i] = a.beta[c] - (float)mean * g; sRstd[i] = rstd; sMR[i] = (float)mean * rstd;
  }
  for (int i = tid; i < fa.N * 2 * a.C; i += WS0_THREADS) sG[i] = 0.0;
  for (int i = tid; i < 2 * a.Co; i += WS0_THREADS) sDb3[i] = 0.f;
  // all-ones core matrix (chunk index c8n) of every row group of every A stage; finite tail
  for (int i = tid; i < 4 * 128; i += WS0_THREADS) {
    const int st = i >> 7, r = i & 127;
    *reinterpret_cast<uint4*>(sA + st * stageA + (r >> 3) * pitchA + c8n * 128 + (r & 7) * 16) = make_uint4(0x3F80u, 0, 0, 0);
  }
  for (int i = tid; i < 128; i += WS0_THREADS) *reinterpret_cast<uint4*>(sTail + i * 16) = make_uint4(0, 0, 0, 0);
  // the MN-major operand reads run past the valid channel groups: keep every operand byte finite from the start
  for (uint32_t i = tid * 16; i < 4 * stageD + 4 * stageH; i += WS0_THREADS * 16) *reinterpret_cast<uint4*>(sD + i) = make_uint4(0, 0, 0, 0);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // TMEM columns: acc1[g] g*H | accG[g] 2H + g*H | accD[g][b] 4H + (2g+b)*C | accW3 4H+4C | accW2 4H+4C+Co
  const uint32_t colD = 4 * a.H, colW3 = colD + 4 * a.C, colW2 = colW3 + a.Co;

  if (warp < WS_LOAD) {
    // ===================================================================== loaders
    int uses = 0;
    for (int64_t g = blockIdx.x + (int64_t)warp * gridDim.x; g < fa.ntiles; g += 4ll * gridDim.x, ++uses) {
      if (uses >= 1) mbar_wait(&a_empty[warp], (uint32_t)((uses - 1) & 1));
      const int n = (int)(g / fa.tps);
      const int tile0 = (int)((g - (int64_t)n * fa.tps) * 128);
      const int nvalid = min(128, (int)a.Vy - tile0);
      ws_stage_tile32<true>(sA + warp * stageA, pitchA, a.y + (int64_t)n * a.Vy * c8n, tile0, nvalid, sScale + n * a.C,
                            sShift + n * a.C, lane);
      ws_stage_tile32<false>(sD + warp * stageD, pitchD, a.dout + (int64_t)n * a.Vout * o8n, tile0, nvalid, nullptr, nullptr, lane);
      fence_proxy_async_smem();
      mbar_arrive(&a_full[warp]);
    }
  } else if (warp >= WS_LOAD + WS_EPI) {
    // ===================================================================== MMA issuers: one thread per role, blocking waits
    // The two epilogue groups are independent pipelines (group g: tiles it == g mod 2, own acc1 / accG / accD[2] / sH / sDh).
    //   issuer g (warps 12, 13), local tile k (it = 2k + g, operand stage s = it & 3, accD buffer b = k & 1):
    //     a_full[s]                              -> G1 Hpre -> acc1[g], G2 dG -> accG[g]     (hp_full[g])
    //     e1_done[g](k), d_empty[g][b] of k - 2  -> G3 dYhat -> accD[g][b]                   (d_full[g][b])
    //   weight-gradient issuer (warp 14), tiles in CTA order: e1_done[g](k) -> G4 dW3 +=, G5 dW2 += into the shared
    //     accumulators (one issuing thread: the MMAs retire in order)                        (a_empty[s], h_free[g])
    // (A single polling issuer — mbarrier.test_wait, ~150 cycles per probe — was measured slower than the in-order one.)
    const int role = warp - (WS_LOAD + WS_EPI);
    if (lane == 0 && role < 2) {
      const int g = role;
      const uint32_t idescH = umma_idesc_bf16(128, a.H, 0, 0), idescD = umma_idesc_bf16(128, a.C, 0, 0);
      const uint64_t dW2 = umma_desc(smem_u32(sW2), 128, c8n * 128), dW3 = umma_desc(smem_u32(sW3t), 128, o8n * 128);
      const uint64_t dW2t = umma_desc(smem_u32(sW2t), 128, h8n * 128);
      int64_t k = 0;
      for (int64_t gt = blockIdx.x + (int64_t)g * gridDim.x; gt < fa.ntiles; gt += 2ll * gridDim.x, ++k) {
        const int64_t it = 2 * k + g;
        const int s = (int)(it & 3), b = (int)(k & 1);
        mbar_wait(&a_full[s], (uint32_t)((it >> 2) & 1));
        tc_fence_after();
        const uint64_t dA = umma_desc(smem_u32(sA + s * stageA), 128, pitchA), dD = umma_desc(smem_u32(sD + s * stageD), 128, pitchD);
        for (int q = 0; q < a.C / 16; ++q)
          umma_bf16(tmem_base + g * a.H, dA + (uint64_t)(q * 16), dW2 + (uint64_t)(q * 16), idescH, q > 0 ? 1u : 0u);
        for (int q = 0; q < a.Co / 16; ++q)
          umma_bf16(tmem_base + 2 * a.H + g * a.H, dD + (uint64_t)(q * 16), dW3 + (uint64_t)(q * 16), idescH, q > 0 ? 1u : 0u);
        tc_commit(&hp_full[g]);
        mbar_wait(&e1_done[g], (uint32_t)(k & 1));
        if (k >= 2) mbar_wait(&d_empty[g * 2 + b], (uint32_t)(((k >> 1) - 1) & 1));
        tc_fence_after();
        const uint64_t dDhK = umma_desc(smem_u32(sDh + g * stageH), 128, pitchH);
        for (int q = 0; q < a.H / 16; ++q)
          umma_bf16(tmem_base + colD + (2 * g + b) * a.C, dDhK + (uint64_t)(q * 16), dW2t + (uint64_t)(q * 16), idescD, q > 0 ? 1u : 0u);
        tc_commit(&d_full[g * 2 + b]);
      }
    } else if (lane == 0 && role == 2) {
      const uint32_t idescW3 = umma_idesc_bf16(128, a.Co, 1, 1), idescW2 = umma_idesc_bf16(128, a.H, 1, 1);
      int64_t it = 0;
      for (int64_t gt = blockIdx.x; gt < fa.ntiles; gt += gridDim.x, ++it) {
        const int g = (int)(it & 1), s = (int)(it & 3);
        const int64_t k = it >> 1;
        mbar_wait(&a_full[s], (uint32_t)((it >> 2) & 1));      // the loader's sA / sD writes (G1 / G2 retired before E1 started)
        mbar_wait(&e1_done[g], (uint32_t)(k & 1));             // sH / sDh of this tile
        tc_fence_after();
        const uint64_t aH = umma_desc(smem_u32(sH + g * stageH), pitchH, 128), bD = umma_desc(smem_u32(sD + s * stageD), pitchD, 128);
        const uint64_t aA = umma_desc(smem_u32(sA + s * stageA), pitchA, 128), bDh = umma_desc(smem_u32(sDh + g * stageH), pitchH, 128);
        for (int q = 0; q < 8; ++q)
          umma_bf16(tmem_base + colW3, aH + (uint64_t)(q * 2 * (pitchH >> 4)), bD + (uint64_t)(q * 2 * (pitchD >> 4)), idescW3,
                    (it > 0 || q > 0) ? 1u : 0u);
        for (int q = 0; q < 8; ++q)
          umma_bf16(tmem_base + colW2, aA + (uint64_t)(q * 2 * (pitchA >> 4)), bDh + (uint64_t)(q * 2 * (pitchH >> 4)), idescW2,
                    (it > 0 || q > 0) ? 1u : 0u);
        tc_commit(&a_empty[s]);
        tc_commit(&h_free[g]);
      }
      tc_commit(w_done);
    }
  } else {
    // ===================================================================== epilogue groups
    // Order per local tile k:  E1(k) -> e1_done -> E2(k-1).  G3 of tile k (and G1/G2 of k+1 behind it) run while the group
    // still has the dYhat pass of tile k-1 to do, so it never idles on d_full (ncu, round 2: 22 % of all samples sat there).
    const int eg = (warp - WS_LOAD) >> 2;
    const int wq = warp & 3, row = wq * 32 + lane;
    const uint32_t lane_off = (uint32_t)(wq * 32) << 16;
    float db3acc = 0.f;        // wq == 0: running sum_v dOut[v, lane] over this group's tiles
    // ---- E2 of one tile: g = dYhat -> bf16 -> HBM ; S1 += g ; S2 += g * xhat
    auto epi2 = [&](int b, int n, int prow, const uint4 (&ypre)[4]) {
      const bool row_ok = prow < (int)a.Vy;
      const int64_t yrow = ((int64_t)n * a.Vy + prow) * c8n;
      const uint32_t td = tmem_base + colD + (2 * eg + b) * a.C + lane_off;
      const float* rs = sRstd + n * a.C; const float* mr = sMR + n * a.C;
      double* sGn = sG + n * 2 * a.C;
#pragma unroll
      for (int c16 = 0; c16 < 2; ++c16) {
        uint32_t v[16];
        tmem_ld16(td + c16 * 16, v);
        tmem_ld_wait();
        float gq[16], gx[16];
        if (row_ok) {
          float yv[16];
          unpack8(ypre[2 * c16], yv);
          unpack8(ypre[2 * c16 + 1], yv + 8);
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            gq[j] = round_bf16(__uint_as_float(v[j]));
            gx[j] = gq[j] * fmaf(yv[j], rs[c16 * 16 + j], -mr[c16 * 16 + j]);
          }
          a.dyhat[yrow + c16 * 2] = pack8(gq);
          a.dyhat[yrow + c16 * 2 + 1] = pack8(gq + 8);
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) { gq[j] = 0.f; gx[j] = 0.f; }
        }
        warp_colsum16(gq, lane);
        warp_colsum16(gx, lane);
        if (!(lane & 1)) {
          const int col = c16 * 16 + colsum16_col(lane);
          atomicAdd(&sGn[col], (double)gq[0]);
          atomicAdd(&sGn[a.C + col], (double)gx[0]);
        }
      }
    };
    auto load_y = [&](uint4 (&ypre)[4], int n, int prow) {   // the row's y (L2-resident: the loaders just read it)
      const bool row_ok = prow < (int)a.Vy;
      const int64_t yrow = ((int64_t)n * a.Vy + prow) * c8n;
#pragma unroll
      for (int c = 0; c < 4; ++c) ypre[c] = row_ok ? __ldg(a.y + yrow + c) : make_uint4(0, 0, 0, 0);
    };
    int64_t k = 0;
    int p_n = 0, p_prow = 0;                              // tile whose E2 is pending
    for (int64_t g = blockIdx.x + (int64_t)eg * gridDim.x; g < fa.ntiles; g += 2ll * gridDim.x, ++k) {
      const int64_t it = 2 * k + eg;
      const int s = (int)(it & 3);
      const int n = (int)(g / fa.tps);
      const int tile0 = (int)((g - (int64_t)n * fa.tps) * 128);
      const int prow = tile0 + row;
      uint4 ypre[4];
      if (k >= 1) load_y(ypre, p_n, p_prow);              // in flight during E1
      mbar_wait(&a_full[s], (uint32_t)((it >> 2) & 1));      // this group reads sD[s] itself (conv3 bias gradient)
      if (wq == 0 && lane < a.Co) {
        const uint8_t* col = sD + s * stageD + (lane >> 3) * 128 + (lane & 7) * 2;
        float sacc = 0.f;
#pragma unroll 8
        for (int r = 0; r < 128; ++r)
          sacc += __uint_as_float((uint32_t)(*reinterpret_cast<const uint16_t*>(col + (r >> 3) * pitchD + (r & 7) * 16)) << 16);
        db3acc += sacc;
      }
      mbar_wait(&hp_full[eg], (uint32_t)(k & 1));
      // sH / sDh[eg] are read by G3 / G4 / G5 of tile k-1 (issued an E2 pass ago)
      if (k >= 1) {
        mbar_wait(&h_free[eg], (uint32_t)((k - 1) & 1));                                               // G4 / G5 of tile k-1
        mbar_wait(&d_full[eg * 2 + (int)((k - 1) & 1)], (uint32_t)(((k - 1) >> 1) & 1));           // G3 of tile k-1 (another issuer)
      }
      tc_fence_after();
      // ---- E1: Hact -> sH[eg], dh -> sDh[eg]
      {
        const uint32_t t1 = tmem_base + eg * a.H + lane_off, tg = tmem_base + 2 * a.H + eg * a.H + lane_off;
        uint8_t* dH = sH + eg * stageH + (row >> 3) * pitchH + (row & 7) * 16;
        uint8_t* dDh = sDh + eg * stageH + (row >> 3) * pitchH + (row & 7) * 16;
#pragma unroll 1
        for (int c16 = 0; c16 < a.H / 16; ++c16) {
          uint32_t v1[16], vg[16];
          tmem_ld16(t1 + c16 * 16, v1);
          tmem_ld16(tg + c16 * 16, vg);
          tmem_ld_wait();
          uint32_t hw[8], dw[8];
          const float4* bp = reinterpret_cast<const float4*>(sB2 + c16 * 16);
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            const float4 b = bp[j4];
            uint64_t va, ga, vb, gb;
            gelu_fast_vg2(add2(pk2(__uint_as_float(v1[4 * j4]), __uint_as_float(v1[4 * j4 + 1])), pk2(b.x, b.y)), va, ga);
            gelu_fast_vg2(add2(pk2(__uint_as_float(v1[4 * j4 + 2]), __uint_as_float(v1[4 * j4 + 3])), pk2(b.z, b.w)), vb, gb);
            ga = mul2(ga, pk2(__uint_as_float(vg[4 * j4]), __uint_as_float(vg[4 * j4 + 1])));
            gb = mul2(gb, pk2(__uint_as_float(vg[4 * j4 + 2]), __uint_as_float(vg[4 * j4 + 3])));
            float e0, e1;
            upk2(va, e0, e1); hw[2 * j4] = pack_bf16(e0, e1);
            upk2(vb, e0, e1); hw[2 * j4 + 1] = pack_bf16(e0, e1);
            upk2(ga, e0, e1); dw[2 * j4] = pack_bf16(e0, e1);
            upk2(gb, e0, e1); dw[2 * j4 + 1] = pack_bf16(e0, e1);
          }
          *reinterpret_cast<uint4*>(dH + (c16 * 2) * 128) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
          *reinterpret_cast<uint4*>(dH + (c16 * 2 + 1) * 128) = make_uint4(hw[4], hw[5], hw[6], hw[7]);
          *reinterpret_cast<uint4*>(dDh + (c16 * 2) * 128) = make_uint4(dw[0], dw[1], dw[2], dw[3]);
          *reinterpret_cast<uint4*>(dDh + (c16 * 2 + 1) * 128) = make_uint4(dw[4], dw[5], dw[6], dw[7]);
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(&e1_done[eg]);
      if (k >= 1) {
        const int b = (int)((k - 1) & 1);
        mbar_wait(&d_full[eg * 2 + b], (uint32_t)(((k - 1) >> 1) & 1));   // already complete (waited before E1)
        tc_fence_after();
        epi2(b, p_n, p_prow, ypre);
        tc_fence_before();
        mbar_arrive(&d_empty[eg * 2 + b]);
      }
      p_n = n; p_prow = prow;
    }
    if (k >= 1) {                                           // drain: E2 of the group's last tile
      uint4 ypre[4];
      load_y(ypre, p_n, p_prow);
      const int b = (int)((k - 1) & 1);
      mbar_wait(&d_full[eg * 2 + b], (uint32_t)(((k - 1) >> 1) & 1));
      tc_fence_after();
      epi2(b, p_n, p_prow, ypre);
    }
    if (wq == 0 && lane < a.Co) sDb3[eg * a.Co + lane] = db3acc;
  }
  tc_fence_before();
  __syncthreads();
  // ---- weight-gradient partials of this CTA (every CTA owns >= 1 tile: the grid never exceeds the tile count)
  if (warp >= WS_LOAD && warp < WS_LOAD + 4) {
    mbar_wait(w_done, 0);
    tc_fence_after();
    const int wq = warp & 3, row = wq * 32 + lane;
    const uint32_t lo = (uint32_t)(wq * 32) << 16;
    float* p3 = fa.part3 + ((int64_t)blockIdx.x * 129 + row) * a.Co;
    if (row < a.Co) fa.part3[((int64_t)blockIdx.x * 129 + 128) * a.Co + row] = sDb3[row] + sDb3[a.Co + row];
    for (int c16 = 0; c16 < a.Co / 16; ++c16) {
      uint32_t v[16];
      tmem_ld16(tmem_base + colW3 + lo + c16 * 16, v);
      tmem_ld_wait();
      if (row < a.H) {
#pragma unroll
        for (int j = 0; j < 16; ++j) p3[c16 * 16 + j] = __uint_as_float(v[j]);
      }
    }
    float* p2 = fa.part2 + ((int64_t)blockIdx.x * 128 + row) * a.H;
    for (int c16 = 0; c16 < a.H / 16; ++c16) {
      uint32_t v[16];
      tmem_ld16(tmem_base + colW2 + lo + c16 * 16, v);
      tmem_ld_wait();
      if (row <= a.C) {
#pragma unroll
        for (int j = 0; j < 16; ++j) p2[c16 * 16 + j] = __uint_as_float(v[j]);
      }
    }
  }
  for (int i = tid; i < fa.N * 2 * a.C; i += WS0_THREADS) {
    const double v = sG[i];
    if (v != 0.0) atomicAdd(&a.gstats[i], v);       // sG is [N][2C] exactly like gstats
  }
  tc_fence_before();
  __syncthreads();
  if (warp == WS_LOAD + WS_EPI) tmem_dealloc(tmem_base, tmem_cols);
}

// ---------------------------------------------------------------------------- generalised warp-specialised backward
// OPT-IN (PCB_BWD_WS=2): the role split of mlp_bwd_ws_kernel for every shape the fused backward serves.
//   C8N = C/8 in {4, 8};  NB = accumulator / Hact / dh buffers: 2 where 2(2H + C) + Co + H <= 512 TMEM columns (level 0),
//   else 1 (level 1, up_0, down_0: 2H + C + Co + H <= 512) — with one buffer G1/G2 of tile i+1 still overlap G3/E2 of tile i
//   (the MMA thread issues them as soon as E1 of tile i has drained the accumulators).
//   Operand stages: 4 (NB = 2) or 2 (NB = 1; shared memory), one loader warp per tile either way.
//   UP mode: dOut rows through a per-tile table (p -> p + 1 on every axis, divisions by multiplication).
template <int C8N, bool NORM, bool TAB, int LD = 8>
__device__ __forceinline__ void ws2_stage_tile(uint8_t* __restrict__ dst, uint32_t pitch, const uint4* __restrict__ src,
                                               const int* __restrict__ tab, int row0, int nvalid, const float* __restrict__ sc,
                                               const float* __restrict__ sh, int lane) {
  static_assert(C8N == 4 || C8N == 8, "C8N");
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
      *reinterpret_cast<uint4*>(dl + rg * pitch + j * 512) = o;
    }
  }
}

__device__ __forceinline__ int ws2_fdiv(int n, uint32_t m, int sh) { return (int)(((uint64_t)(uint32_t)n * (uint64_t)m) >> sh); }

template <int C8N, int NB, int NST>   // NST operand stages (2..4): with 4 every loader warp has a tile in flight
__global__ void __launch_bounds__(WS_THREADS, 1) mlp_bwd_ws2_kernel(MlpBwdFusedArgs fa) {
  // One accumulator set (NB == 1): nothing of tile i+1 can overlap E1 of tile i, so BOTH epilogue groups work on the SAME tile
  // — each takes half of the hidden columns in E1 and half of the channel columns in E2 (a warp reads its own TMEM lane
  // quarter; columns are free) — instead of alternating tiles with one group idle: per-tile latency of E1 + E2 halves
  // (opt-in, PCB_BWD_SPLIT=1: measured SLOWER than alternating groups once the operand ring has 4 stages — the shapes were loader-bound).
  const bool SPLIT = NB == 1 && fa.split != 0;
  const MlpBwdArgs& a = fa.m;
  extern __shared__ __align__(128) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int c8n = C8N, h8n = a.H >> 3, o8n = a.Co >> 3;
  const uint32_t pitchA = (c8n + 1) * 128, pitchD = o8n * 128, pitchH = h8n * 128;
  const uint32_t stageA = 16 * pitchA, stageD = 16 * pitchD, stageH = 16 * pitchH;
  uint8_t* sW2 = smem;                                 // [H x C]   K-major (B of G1)
  uint8_t* sW3t = sW2 + a.H * a.C * 2;                 // [H x Co]  K-major (B of G2)
  uint8_t* sW2t = sW3t + a.H * a.Co * 2;               // [C x H]   K-major (B of G3)
  uint8_t* sA = sW2t + a.C * a.H * 2;                  // NST x ([128 x C] + ones core matrix per row group)
  uint8_t* sD = sA + NST * stageA;                     // NST x [128 x Co]
  uint8_t* sH = sD + NST * stageD;                     // NB x [128 x H]
  uint8_t* sDh = sH + NB * stageH;                     // NB x [128 x H]
  uint8_t* sTail = sDh + NB * stageH;                  // 2 KB finite padding behind the last MN-major operand
  float* sScale = reinterpret_cast<float*>(sTail + 2048);   // [N][C] gamma*rstd
  float* sShift = sScale + fa.N * a.C;                 // [N][C]
  float* sRstd = sShift + fa.N * a.C;                  // [N][C]
  float* sMR = sRstd + fa.N * a.C;                     // [N][C] mean*rstd
  float* sB2 = sMR + fa.N * a.C;                       // [H]
  float* sDb3 = sB2 + a.H;                             // [2][Co]
  int* sRow = reinterpret_cast<int*>(sDb3 + 2 * a.Co); // [4 loader warps][128] dOut row of the tile in flight (UP mode only)
  double* sG = reinterpret_cast<double*>(sRow + (a.mode == PCB_DW_UP ? WS_LOAD * 128 : 0));   // [N][2C]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sG + fa.N * 2 * a.C);
  // Barriers are indexed by the TILE (it & 3 for the loader hand-offs, it & 1 for the accumulator hand-offs) and the DATA
  // by it % NST / it % NB: every barrier then has exactly one waiting role that sees each of its completions, whatever
  // NST / NB are (a waiter that skipped a completion would mis-read the phase parity).
  uint64_t* a_full = bars;          // [4] loader warp (it & 3) -> MMA / epilogue
  uint64_t* a_empty = bars + 4;     // [4] MMA (G4/G5 of tile it retired) -> the loader of tile it + NST
  uint64_t* hp_full = bars + 8;     // [2] MMA (G1/G2) -> E1
  uint64_t* e1_done = bars + 10;    // [2] E1 -> MMA
  uint64_t* d_full = bars + 12;     // [2] MMA (G3) -> E2
  uint64_t* d_empty = bars + 14;    // [2] E2 -> MMA
  uint64_t* h_free = bars + 16;     // [2] MMA (G3/G4/G5 retired) -> E1 of tile it + NB
  uint64_t* w_done = bars + 18;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 19);

  const uint32_t tmem_cols = tmem_cols_pow2((uint32_t)(NB * (2 * a.H + a.C) + a.Co + a.H));
  if (warp == WS_LOAD + WS_EPI) tmem_alloc(tmem_slot, tmem_cols);
  if (tid == 0) {
    for (int i = 0; i < 4; ++i) { mbar_init(&a_full[i], 32); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&hp_full[i], 1); mbar_init(&e1_done[i], SPLIT ? 256 : 128); mbar_init(&d_full[i], 1);
      mbar_init(&d_empty[i], SPLIT ? 256 : 128);
      mbar_init(&h_free[i], 1);
    }
    mbar_init(w_done, 1);
    fence_mbar_init();
  }
  stage_rows_k(sW2, a.w2, a.H, c8n, c8n, tid, WS_THREADS);
  stage_rows_k(sW3t, a.w3t, a.H, o8n, o8n, tid, WS_THREADS);
  stage_rows_k(sW2t, a.w2t, a.C, h8n, h8n, tid, WS_THREADS);
  for (int i = tid; i < a.H; i += WS_THREADS) sB2[i] = a.b2[i];
  for (int i = tid; i < fa.N * a.C; i += WS_THREADS) {
    const int n = i / a.C, c = i - n * a.C;
    const double sm = a.stats[(int64_t)n * 2 * a.C + c], q = a.stats[(int64_t)n * 2 * a.C + a.C + c];
    const double mean = sm * (double)a.inv_count;
    double var = q * (double)a.inv_count - mean * mean;
    if (var < 0.0) var = 0.0;
    const float rstd = (float)(1.0 / sqrt(var + 1e-5));
    const float g = a.gamma[c] * rstd;
    sScale[i] = g; sShift[i] = a.beta[c] - (float)mean * g; sRstd[i] = rstd; sMR[i] = (float)mean * rstd;
  }
  for (int i = tid; i < fa.N * 2 * a.C; i += WS_THREADS) sG[i] = 0.0;
  for (int i = tid; i < 2 * a.Co; i += WS_THREADS) sDb3[i] = 0.f;
  // every operand byte finite from the start (MN-major reads run past the valid channel groups), then the ones matrices
  for (uint32_t i = tid * 16; i < NST * (stageA + stageD) + 2 * NB * stageH + 2048; i += WS_THREADS * 16)
    *reinterpret_cast<uint4*>(sA + i) = make_uint4(0, 0, 0, 0);
  __syncthreads();
  for (int i = tid; i < NST * 128; i += WS_THREADS) {
    const int st = i >> 7, r = i & 127;
    *reinterpret_cast<uint4*>(sA + st * stageA + (r >> 3) * pitchA + c8n * 128 + (r & 7) * 16) = make_uint4(0x3F80u, 0, 0, 0);
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // TMEM columns: acc1[b] b*H | accG[b] NB*H + b*H | accD[b] 2*NB*H + b*C | accW3 | accW2
  const uint32_t colG = NB * a.H, colD = 2 * NB * a.H, colW3 = colD + NB * a.C, colW2 = colW3 + a.Co;
  const bool up = a.mode == PCB_DW_UP;

  if (warp < WS_LOAD) {
    // ===================================================================== loaders: warp w owns the tiles it == w (mod 4)
    int* rowO = sRow + warp * 128;
    for (int64_t it = warp, g = blockIdx.x + (int64_t)warp * gridDim.x; g < fa.ntiles; g += 4ll * gridDim.x, it += 4) {
      const int s = (int)(it % NST);
      const int64_t prev = it - NST;                    // the tile that used this stage last
      if (prev >= 0) mbar_wait(&a_empty[prev & 3], (uint32_t)((prev >> 2) & 1));
      const int n = (int)(g / fa.tps);
      const int tile0 = (int)((g - (int64_t)n * fa.tps) * 128);
      const int nvalid = min(128, (int)a.Vy - tile0);
      if (fa.ld16)
        ws2_stage_tile<C8N, true, false, 16>(sA + s * stageA, pitchA, a.y + (int64_t)n * a.Vy * c8n, nullptr, tile0, nvalid,
                                             sScale + n * a.C, sShift + n * a.C, lane);
      else
        ws2_stage_tile<C8N, true, false>(sA + s * stageA, pitchA, a.y + (int64_t)n * a.Vy * c8n, nullptr, tile0, nvalid,
                                         sScale + n * a.C, sShift + n * a.C, lane);
      const uint4* dn = a.dout + (int64_t)n * a.Vout * o8n;
      if (up) {
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int r = lane + 32 * k, p = tile0 + r;
          int ro = -1;
          if (r < nvalid) {
            const int t = ws2_fdiv(p, fa.dm2, fa.ds2), x = p - t * a.y2;
            const int z = ws2_fdiv(t, fa.dm1, fa.ds1), y = t - z * a.y1;
            ro = ((z + 1) * a.o1 + (y + 1)) * a.o2 + (x + 1);
          }
          rowO[r] = ro;
        }
        __syncwarp();
        if (o8n == 8) ws2_stage_tile<8, false, true>(sD + s * stageD, pitchD, dn, rowO, 0, 0, nullptr, nullptr, lane);
        else ws2_stage_tile<4, false, true>(sD + s * stageD, pitchD, dn, rowO, 0, 0, nullptr, nullptr, lane);
      } else {
        if (o8n == 8) ws2_stage_tile<8, false, false>(sD + s * stageD, pitchD, dn, nullptr, tile0, nvalid, nullptr, nullptr, lane);
        else ws2_stage_tile<4, false, false>(sD + s * stageD, pitchD, dn, nullptr, tile0, nvalid, nullptr, nullptr, lane);
      }
      fence_proxy_async_smem();
      mbar_arrive(&a_full[warp]);                        // warp == it & 3
    }
  } else if (warp == WS_LOAD + WS_EPI) {
    // ===================================================================== MMA issuer (one thread)
    if (lane == 0) {
      const uint32_t idescH = umma_idesc_bf16(128, a.H, 0, 0), idescD = umma_idesc_bf16(128, a.C, 0, 0);
      const uint32_t idescW3 = umma_idesc_bf16(128, a.Co, 1, 1), idescW2 = umma_idesc_bf16(128, a.H, 1, 1);
      const uint64_t dW2 = umma_desc(smem_u32(sW2), 128, c8n * 128), dW3 = umma_desc(smem_u32(sW3t), 128, o8n * 128);
      const uint64_t dW2t = umma_desc(smem_u32(sW2t), 128, h8n * 128);
      auto second_half = [&](int64_t j) {     // G3 + weight-gradient GEMMs of local tile j
        const int bj = (int)(j % NB), sj = (int)(j % NST), qj = (int)(j & 1);
        mbar_wait(&e1_done[qj], (uint32_t)((j >> 1) & 1));
        const int64_t jj = j - NB;                      // the tile whose dYhat sat in accD[bj] before
        if (jj >= 0) mbar_wait(&d_empty[jj & 1], (uint32_t)((jj >> 1) & 1));
        tc_fence_after();
        const uint64_t dDhK = umma_desc(smem_u32(sDh + bj * stageH), 128, pitchH);
        for (int k = 0; k < a.H / 16; ++k)
          umma_bf16(tmem_base + colD + bj * a.C, dDhK + (uint64_t)(k * 16), dW2t + (uint64_t)(k * 16), idescD, k > 0 ? 1u : 0u);
        tc_commit(&d_full[qj]);
        const uint64_t aH = umma_desc(smem_u32(sH + bj * stageH), pitchH, 128), bD = umma_desc(smem_u32(sD + sj * stageD), pitchD, 128);
        const uint64_t aA = umma_desc(smem_u32(sA + sj * stageA), pitchA, 128), bDh = umma_desc(smem_u32(sDh + bj * stageH), pitchH, 128);
        for (int k = 0; k < 8; ++k)
          umma_bf16(tmem_base + colW3, aH + (uint64_t)(k * 2 * (pitchH >> 4)), bD + (uint64_t)(k * 2 * (pitchD >> 4)), idescW3,
                    (j > 0 || k > 0) ? 1u : 0u);
        for (int k = 0; k < 8; ++k)
          umma_bf16(tmem_base + colW2, aA + (uint64_t)(k * 2 * (pitchA >> 4)), bDh + (uint64_t)(k * 2 * (pitchH >> 4)), idescW2,
                    (j > 0 || k > 0) ? 1u : 0u);
        tc_commit(&a_empty[j & 3]);
        tc_commit(&h_free[qj]);
      };
      int64_t it = 0;
      for (int64_t g = blockIdx.x; g < fa.ntiles; g += gridDim.x, ++it) {
        const int s = (int)(it % NST), b = (int)(it % NB);
        mbar_wait(&a_full[it & 3], (uint32_t)((it >> 2) & 1));
        // one accumulator buffer: E1 of the previous tile must have drained acc1 / accG before they are overwritten,
        // so its second half (which waits for exactly that) is issued first; with two buffers it trails by one tile
        if (NB == 1 && it >= 1) second_half(it - 1);
        tc_fence_after();
        const uint64_t dA = umma_desc(smem_u32(sA + s * stageA), 128, pitchA), dD = umma_desc(smem_u32(sD + s * stageD), 128, pitchD);
        for (int k = 0; k < a.C / 16; ++k)
          umma_bf16(tmem_base + b * a.H, dA + (uint64_t)(k * 16), dW2 + (uint64_t)(k * 16), idescH, k > 0 ? 1u : 0u);
        for (int k = 0; k < a.Co / 16; ++k)
          umma_bf16(tmem_base + colG + b * a.H, dD + (uint64_t)(k * 16), dW3 + (uint64_t)(k * 16), idescH, k > 0 ? 1u : 0u);
        tc_commit(&hp_full[it & 1]);
        if (NB == 2 && it >= 1) second_half(it - 1);
      }
      if (it >= 1) second_half(it - 1);
      tc_commit(w_done);
    }
  } else {
    // ===================================================================== epilogue groups (tiles it == eg mod 2)
    const int eg = (warp - WS_LOAD) >> 2;
    const int wq = warp & 3, row = wq * 32 + lane;
    const uint32_t lane_off = (uint32_t)(wq * 32) << 16;
    float db3acc = 0.f;        // row < Co: running sum_v dOut[v, row] over this group's tiles
    // alternating: group eg owns the tiles it == eg (mod 2); SPLIT: both groups visit every tile, barriers indexed by it & 1
    const int tstep = SPLIT ? 1 : 2;
    int64_t it = SPLIT ? 0 : eg;
    for (int64_t g = blockIdx.x + it * gridDim.x; g < fa.ntiles; g += (int64_t)tstep * gridDim.x, it += tstep) {
      const int b = (int)(it % NB), s = (int)(it % NST);
      const int q = (int)(it & 1);                         // barrier index of this tile (== eg when alternating)
      const uint32_t par = (uint32_t)((it >> 1) & 1);      // every waiter sees every completion of the barriers it waits on
      const int n = (int)(g / fa.tps);
      const int tile0 = (int)((g - (int64_t)n * fa.tps) * 128);
      const int prow = tile0 + row;
      const bool row_ok = prow < (int)a.Vy;
      mbar_wait(&a_full[it & 3], (uint32_t)((it >> 2) & 1)); // this group reads sD[s] itself (conv3 bias gradient)
      if (row < a.Co && (!SPLIT || eg == 0)) {
        const uint8_t* col = sD + s * stageD + (row >> 3) * 128 + (row & 7) * 2;
        float sacc = 0.f;
#pragma unroll 8
        for (int r = 0; r < 128; ++r)
          sacc += __uint_as_float((uint32_t)(*reinterpret_cast<const uint16_t*>(col + (r >> 3) * pitchD + (r & 7) * 16)) << 16);
        db3acc += sacc;
      }
      mbar_wait(&hp_full[q], par);
      {
        const int64_t pj = it - NB;                       // the tile whose Hact / dh sat in sH[b] / sDh[b] before
        if (pj >= 0) mbar_wait(&h_free[pj & 1], (uint32_t)((pj >> 1) & 1));
      }
      tc_fence_after();
      // ---- E1: Hact -> sH[b], dh -> sDh[b]
      {
        const uint32_t t1 = tmem_base + b * a.H + lane_off, tg = tmem_base + colG + b * a.H + lane_off;
        uint8_t* dH = sH + b * stageH + (row >> 3) * pitchH + (row & 7) * 16;
        uint8_t* dDh = sDh + b * stageH + (row >> 3) * pitchH + (row & 7) * 16;
        const int h16 = a.H / 16, e1_lo = SPLIT ? eg * (h16 >> 1) : 0, e1_hi = SPLIT ? (eg ? h16 : (h16 >> 1)) : h16;
#pragma unroll 1
        for (int c16 = e1_lo; c16 < e1_hi; ++c16) {
          uint32_t v1[16], vg[16];
          tmem_ld16(t1 + c16 * 16, v1);
          tmem_ld16(tg + c16 * 16, vg);
          tmem_ld_wait();
          uint32_t hw[8], dw[8];
          const float4* bp = reinterpret_cast<const float4*>(sB2 + c16 * 16);
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            const float4 bb = bp[j4];
            uint64_t va, ga, vb, gb;
            gelu_fast_vg2(add2(pk2(__uint_as_float(v1[4 * j4]), __uint_as_float(v1[4 * j4 + 1])), pk2(bb.x, bb.y)), va, ga);
            gelu_fast_vg2(add2(pk2(__uint_as_float(v1[4 * j4 + 2]), __uint_as_float(v1[4 * j4 + 3])), pk2(bb.z, bb.w)), vb, gb);
            ga = mul2(ga, pk2(__uint_as_float(vg[4 * j4]), __uint_as_float(vg[4 * j4 + 1])));
            gb = mul2(gb, pk2(__uint_as_float(vg[4 * j4 + 2]), __uint_as_float(vg[4 * j4 + 3])));
            float e0, e1;
            upk2(va, e0, e1); hw[2 * j4] = pack_bf16(e0, e1);
            upk2(vb, e0, e1); hw[2 * j4 + 1] = pack_bf16(e0, e1);
            upk2(ga, e0, e1); dw[2 * j4] = pack_bf16(e0, e1);
            upk2(gb, e0, e1); dw[2 * j4 + 1] = pack_bf16(e0, e1);
          }
          *reinterpret_cast<uint4*>(dH + (c16 * 2) * 128) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
          *reinterpret_cast<uint4*>(dH + (c16 * 2 + 1) * 128) = make_uint4(hw[4], hw[5], hw[6], hw[7]);
          *reinterpret_cast<uint4*>(dDh + (c16 * 2) * 128) = make_uint4(dw[0], dw[1], dw[2], dw[3]);
          *reinterpret_cast<uint4*>(dDh + (c16 * 2 + 1) * 128) = make_uint4(dw[4], dw[5], dw[6], dw[7]);
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(&e1_done[q]);
      // ---- E2: g = dYhat -> bf16 -> HBM ; S1 += g ; S2 += g * xhat
      const int64_t yrow = ((int64_t)n * a.Vy + prow) * c8n;
      mbar_wait(&d_full[q], par);
      tc_fence_after();
      {
        const uint32_t td = tmem_base + colD + b * a.C + lane_off;
        const float* rs = sRstd + n * a.C; const float* mr = sMR + n * a.C;
        double* sGn = sG + n * 2 * a.C;
        constexpr int C16N = C8N / 2;
        const int e2_lo = SPLIT ? eg * (C16N >> 1) : 0, e2_hi = SPLIT ? (eg ? C16N : (C16N >> 1)) : C16N;
#pragma unroll 1
        for (int c16 = e2_lo; c16 < e2_hi; ++c16) {
          uint32_t v[16];
          tmem_ld16(td + c16 * 16, v);
          uint4 y0 = make_uint4(0, 0, 0, 0), y1 = make_uint4(0, 0, 0, 0);
          if (row_ok) { y0 = __ldg(a.y + yrow + c16 * 2); y1 = __ldg(a.y + yrow + c16 * 2 + 1); }   // L2-resident
          tmem_ld_wait();
          float gq[16], gx[16];
          if (row_ok) {
            float yv[16];
            unpack8(y0, yv);
            unpack8(y1, yv + 8);
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              gq[j] = round_bf16(__uint_as_float(v[j]));
              gx[j] = gq[j] * fmaf(yv[j], rs[c16 * 16 + j], -mr[c16 * 16 + j]);
            }
            a.dyhat[yrow + c16 * 2] = pack8(gq);
            a.dyhat[yrow + c16 * 2 + 1] = pack8(gq + 8);
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) { gq[j] = 0.f; gx[j] = 0.f; }
          }
          warp_colsum16(gq, lane);
          warp_colsum16(gx, lane);
          if (!(lane & 1)) {
            const int col = c16 * 16 + colsum16_col(lane);
            atomicAdd(&sGn[col], (double)gq[0]);
            atomicAdd(&sGn[a.C + col], (double)gx[0]);
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&d_empty[q]);
    }
    if (row < a.Co) sDb3[eg * a.Co + row] = db3acc;
  }
  tc_fence_before();
  __syncthreads();
  // ---- weight-gradient partials of this CTA (every CTA owns >= 1 tile: the grid never exceeds the tile count)
  if (warp >= WS_LOAD && warp < WS_LOAD + 4) {
    mbar_wait(w_done, 0);
    tc_fence_after();
    const int wq = warp & 3, row = wq * 32 + lane;
    const uint32_t lo = (uint32_t)(wq * 32) << 16;
    float* p3 = fa.part3 + ((int64_t)blockIdx.x * 129 + row) * a.Co;
    if (row < a.Co) fa.part3[((int64_t)blockIdx.x * 129 + 128) * a.Co + row] = sDb3[row] + sDb3[a.Co + row];
    for (int c16 = 0; c16 < a.Co / 16; ++c16) {
      uint32_t v[16];
      tmem_ld16(tmem_base + colW3 + lo + c16 * 16, v);
      tmem_ld_wait();
      if (row < a.H) {
#pragma unroll
        for (int j = 0; j < 16; ++j) p3[c16 * 16 + j] = __uint_as_float(v[j]);
      }
    }
    float* p2 = fa.part2 + ((int64_t)blockIdx.x * 128 + row) * a.H;
    for (int c16 = 0; c16 < a.H / 16; ++c16) {
      uint32_t v[16];
      tmem_ld16(tmem_base + colW2 + lo + c16 * 16, v);
      tmem_ld_wait();
      if (row <= a.C) {
#pragma unroll
        for (int j = 0; j < 16; ++j) p2[c16 * 16 + j] = __uint_as_float(v[j]);
      }
    }
  }
  for (int i = tid; i < fa.N * 2 * a.C; i += WS_THREADS) {
    const double v = sG[i];
    if (v != 0.0) atomicAdd(&a.gstats[i], v);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == WS_LOAD + WS_EPI) tmem_dealloc(tmem_base, tmem_cols);
}

static size_t mlp_bwd_ws2_smem(int C, int H, int Co, int N, int NB, int NST, bool up = true) {
  return (size_t)H * C * 2 + (size_t)H * Co * 2 + (size_t)C * H * 2 + (size_t)NST * 16 * (C / 8 + 1) * 128 +
         (size_t)NST * 16 * (Co / 8) * 128 + (size_t)2 * NB * 16 * (H / 8) * 128 + 2048 + (size_t)4 * N * C * 4 + (size_t)H * 4 +
         (size_t)2 * Co * 4 + (up ? (size_t)WS_LOAD * 128 * 4 : 0) + (size_t)N * 2 * C * 8 + 19 * 8 + 16 + 128;
}

static void ws2_magic(uint32_t d, uint32_t& m, int& sh) {
  int l = 0;
  while ((1ull << l) < d) ++l;
  const unsigned __int128 p = (unsigned __int128)1 << (31 + l);
  m = (uint32_t)((p + d - 1) / d);
  sh = 31 + l;
}

static size_t mlp_bwd_ws_smem(int C, int H, int Co, int N) {
  return (size_t)H * C * 2 + (size_t)H * Co * 2 + (size_t)C * H * 2 + (size_t)4 * 16 * (C / 8 + 1) * 128 + (size_t)4 * 16 * (Co / 8) * 128 +
         (size_t)4 * 16 * (H / 8) * 128 + 2048 + (size_t)4 * N * C * 4 + (size_t)H * 4 + (size_t)2 * Co * 4 + (size_t)N * 2 * C * 8 +
         23 * 8 + 16 + 128;
}

static size_t mlp_bwd_fused_smem(int C, int H, int Co, int N) {
  return (size_t)H * C * 2 + (size_t)H * Co * 2 + (size_t)C * H * 2 + (size_t)16 * (C / 8 + 1) * 128 + (size_t)16 * (Co / 8) * 128 +
         (size_t)16 * (H / 8) * 128 + (size_t)16 * (H / 8) * 128 + (size_t)128 * (C * 2 + 16) + 2048 + (size_t)4 * N * C * 4 + (size_t)H * 4 +
         (size_t)2 * C * 8 + 128 * 4 + 2 * 8 + 16 + 128;
}

// ============================================================================ TN (split-K) wgrad GEMM
struct TnArgs {
  const uint4* A; const uint4* B;
  float* part;                  // [P][Mtot][Ncols_tot] fp32 partial sums
  const double* stats; const float* gamma; const float* beta;   // GroupNorm-apply on B (nullable)
  int64_t a_pitch8, b_pitch8;   // row pitch (uint4 units)
  int64_t a_sample8, b_sample8; // per-sample stride (uint4 units)
  int Ma, Nb;                   // valid A columns (M of D), B columns (N of D without the ones block)
  int ones;                     // append an all-ones column block (16 wide) to B chunk 0
  int mapA, mapB;
  int tz, ty, tx, tstride, tpad, bs0;   // conv-tap row map for B (mapB 4: conv, 5: transposed conv)
  int d1, d2;                   // iteration box trailing dims (tile voxel -> coordinates)
  int as1, as2, bs1, bs2;       // trailing spatial dims of the A / B tensors (for the row maps)
  int64_t V;                    // voxels per sample in the iteration box
  int N;                        // samples
  int Mtot, Ncols_tot;
  float inv_count;
};

constexpr int TN_NCHUNK = 128;

__global__ void __launch_bounds__(128) tn_gemm_kernel(TnArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5;
  const int mt = blockIdx.y, nc = blockIdx.z;
  const int m0 = mt * 128;
  const int mvalid = min(128, a.Ma - m0), m8n = mvalid >> 3;
  const int n0 = nc * TN_NCHUNK;
  const int nb = min(TN_NCHUNK, a.Nb - n0), nb8 = nb >> 3;
  const bool ones = a.ones && nc == 0;
  const bool m8pow2 = (m8n & (m8n - 1)) == 0, n8pow2 = (nb8 & (nb8 - 1)) == 0;
  const int m8sh = __ffs(m8n) - 1, n8sh = __ffs(nb8) - 1;
  const int ncols = nb + (ones ? 16 : 0);
  // MN-major canonical layout: 16 B chunk (8 channels of voxel v) of channel-group g at g*SBO + v*16.
  // SBO padded so a quarter-warp's 8 stores land on 8 distinct 16 B bank groups.
  const uint32_t padA = m8n >= 8 ? 16 : (m8n >= 4 ? 32 : 64), padB = nb8 >= 8 ? 16 : (nb8 >= 4 ? 32 : 64);
  const uint32_t sboA = 2048 + padA, sboB = 2048 + padB;
  const uint32_t stageA = 16 * sboA, stageB = ((TN_NCHUNK >> 3) + 2) * sboB;
  uint8_t* sA = smem;                 // single stage: several CTAs per SM overlap load / MMA phases
  uint8_t* sB = sA + stageA;
  float* sScale = reinterpret_cast<float*>(sB + stageB);   // [nb]
  float* sShift = sScale + TN_NCHUNK;
  uint64_t* bar = reinterpret_cast<uint64_t*>(sShift + TN_NCHUNK);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);

  const uint32_t tmem_cols = tmem_cols_pow2(ncols);
  if (warp == 0) tmem_alloc(tmem_slot, tmem_cols);
  if (tid == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  // zero the stage once: channel groups beyond the valid ones stay zero for the whole kernel
  for (uint32_t i = tid * 16; i < stageA + stageB; i += 128 * 16) *reinterpret_cast<uint4*>(smem + i) = make_uint4(0, 0, 0, 0);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t acc = *tmem_slot;
  const uint32_t idesc = umma_idesc_bf16(128, ncols, 1, 1);

  const int64_t tps = (a.V + 127) / 128;           // tiles per sample
  const int64_t ntiles = tps * a.N;
  int cur_n = -1;
  uint32_t it = 0;
  uint32_t ph = 0;
  for (int64_t g = blockIdx.x; g < ntiles; g += gridDim.x, ++it) {
    const int n = (int)(g / tps);
    const int64_t v0 = (g - (int64_t)n * tps) * 128;
    if (it >= 1) { mbar_wait(bar, ph); ph ^= 1; }   // previous tile's MMAs have consumed the stage
    if (a.stats != nullptr && n != cur_n) {   // GroupNorm affine of this sample for the B channels
      __syncthreads();
      for (int c = tid; c < nb; c += 128) {
        const int cg = n0 + c;
        const double sm = a.stats[(int64_t)n * 2 * a.Nb + cg], q = a.stats[(int64_t)n * 2 * a.Nb + a.Nb + cg];
        const double mean = sm * (double)a.inv_count;
        double var = q * (double)a.inv_count - mean * mean;
        if (var < 0.0) var = 0.0;
        const float gg = a.gamma[cg] * (float)(1.0 / sqrt(var + 1e-5));
        sScale[c] = gg; sShift[c] = a.beta[cg] - (float)mean * gg;
      }
      __syncthreads();
    }
    cur_n = n;
    const uint4* An = a.A + (int64_t)n * a.a_sample8;
    const uint4* Bn = a.B + (int64_t)n * a.b_sample8;
    staged_copy<16>(128 * m8n, tid, 128,   // 16 loads in flight per thread: the kernel is latency-bound at 2 CTAs/SM
       
        [&](int q) {
          const int v = m8pow2 ? (q >> m8sh) : (q / m8n), g8 = q - v * m8n;
          if (v0 + v >= a.V) return make_uint4(0, 0, 0, 0);
          const int64_t r = map_row(a.mapA, v0 + v, a.d1, a.d2, a.as1, a.as2);
          return __ldg(An + r * a.a_pitch8 + (m0 >> 3) + g8);
        },
        [&](int q, const uint4& val) {
          const int v = m8pow2 ? (q >> m8sh) : (q / m8n), g8 = q - v * m8n;
          *reinterpret_cast<uint4*>(sA + g8 * sboA + v * 16) = val;
        });
    staged_copy<16>(128 * nb8, tid, 128,
        [&](int q) {
          const int v = n8pow2 ? (q >> n8sh) : (q / nb8), g8 = q - v * nb8;
          if (v0 + v >= a.V) return make_uint4(0, 0, 0, 0);
          int64_t r;
          if (a.mapB >= 4) {   // dense-conv weight gradient: B row = input voxel feeding output voxel v through this tap
            const int64_t ov = v0 + v;
            const int ox = (int)(ov % a.d2), oy = (int)((ov / a.d2) % a.d1), oz = (int)(ov / ((int64_t)a.d2 * a.d1));
            int iz, iy, ix;
            bool ok = true;
            if (a.mapB == 4) {
              iz = oz * a.tstride + a.tz - a.tpad; iy = oy * a.tstride + a.ty - a.tpad; ix = ox * a.tstride + a.tx - a.tpad;
            } else {
              const int qz = oz + a.tpad - a.tz, qy = oy + a.tpad - a.ty, qx = ox + a.tpad - a.tx;
              ok = qz >= 0 && qy >= 0 && qx >= 0 && (qz % a.tstride == 0) && (qy % a.tstride == 0) && (qx % a.tstride == 0);
              iz = qz / a.tstride; iy = qy / a.tstride; ix = qx / a.tstride;
            }
            ok = ok && iz >= 0 && iz < a.bs0 && iy >= 0 && iy < a.bs1 && ix >= 0 && ix < a.bs2;
            if (!ok) return make_uint4(0, 0, 0, 0);
            r = ((int64_t)iz * a.bs1 + iy) * a.bs2 + ix;
          } else {
            r = map_row(a.mapB, v0 + v, a.d1, a.d2, a.bs1, a.bs2);
          }
          return __ldg(Bn + r * a.b_pitch8 + (n0 >> 3) + g8);
        },
        [&](int q, const uint4& raw) {
          const int v = n8pow2 ? (q >> n8sh) : (q / nb8), g8 = q - v * nb8;
          uint4 val = raw;
          if (a.stats != nullptr && v0 + v < a.V) {
            float f[8];
            unpack8(raw, f);
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] = fmaf(f[j], sScale[g8 * 8 + j], sShift[g8 * 8 + j]);
            val = pack8(f);
          }
          *reinterpret_cast<uint4*>(sB + g8 * sboB + v * 16) = val;
        });
    if (ones) {   // column nb = 1.0 for valid voxels (bf16 0x3F80), the other 15 columns zero
      const int v = tid;
      *reinterpret_cast<uint4*>(sB + nb8 * sboB + v * 16) = make_uint4((v0 + v < a.V) ? 0x3F80u : 0u, 0, 0, 0);
    }
    fence_proxy_async_smem();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      const uint64_t ad = umma_desc(smem_u32(sA), 128, sboA), bd = umma_desc(smem_u32(sB), 128, sboB);
      for (int k = 0; k < 8; ++k)   // 128 voxels = 8 x K16; K-groups of 8 voxels are LBO = 128 B apart
        umma_bf16(acc, ad + (uint64_t)(k * 16), bd + (uint64_t)(k * 16), idesc, (it > 0 || k > 0) ? 1u : 0u);
      tc_commit(bar);
    }
  }
  if (it > 0) mbar_wait(bar, ph);   // drain
  tc_fence_after();
  {
    const uint32_t trow = acc + ((uint32_t)(warp * 32) << 16);
    float* prow = a.part + ((int64_t)blockIdx.x * a.Mtot + m0 + tid) * a.Ncols_tot;
    for (int c16 = 0; c16 < ncols / 16; ++c16) {
      uint32_t v[16];
      tmem_ld16(trow + c16 * 16, v);
      tmem_ld_wait();
      if (it == 0) {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = 0u;   // this CTA had no tile: contribute zeros
      }
      if (tid < mvalid) {
        const int col0 = (c16 * 16 < nb) ? (n0 + c16 * 16) : a.Nb;   // ones block lives after all Nb columns
        float4* dst = reinterpret_cast<float4*>(prow + col0);
        dst[0] = make_float4(__uint_as_float(v[0]), __uint_as_float(v[1]), __uint_as_float(v[2]), __uint_as_float(v[3]));
        dst[1] = make_float4(__uint_as_float(v[4]), __uint_as_float(v[5]), __uint_as_float(v[6]), __uint_as_float(v[7]));
        dst[2] = make_float4(__uint_as_float(v[8]), __uint_as_float(v[9]), __uint_as_float(v[10]), __uint_as_float(v[11]));
        dst[3] = make_float4(__uint_as_float(v[12]), __uint_as_float(v[13]), __uint_as_float(v[14]), __uint_as_float(v[15]));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(acc, tmem_cols);
}

// ---------------------------------------------------------------------------- warp-specialised split-K wgrad GEMM
// Same operands, row maps, GroupNorm-apply and partial layout as tn_gemm_kernel, restructured (round 2: the single-stage kernel ran
// the up_1 weight gradients at 1.45 TB/s — ncu: 14 k SASS lines of unrolled row-map arithmetic with 64-bit divisions
// (instruction-cache misses, `no_inst` 15 %), a third of the resident CTAs parked in tcgen05.alloc because only two fit in
// TMEM, 10 warps per SM):
//   * persistent, one CTA per SM; 4 loader warps -> NS-stage ring -> 1 MMA-issuing thread -> loaders write the partials;
//   * operands go global -> shared with cp.async (16 B, zero-fill for rows outside the box): no register staging, every copy of
//     a tile (64 KB) is in flight at once and the next tile's copies are issued before the previous tile is handed to the MMA
//     thread; only a B operand that needs the GroupNorm affine passes through registers;
//   * the source row of every tile voxel is computed ONCE per tile (one voxel per loader thread, 32-bit magic divisions) into a
//     shared table instead of once per 16-byte chunk with 64-bit divisions;
//   * MT (1 or 2) row tiles of dW per CTA share the staged B tile (M = 256: B is read once instead of twice).
struct TnWsArgs {
  TnArgs t;
  int MT;                       // row tiles (128 rows of dW each) per CTA
  int NS;                       // stages
  uint32_t dm2, dm1;            // magic multipliers of the exact division by d2 / d1
  int ds2, ds1;
};
constexpr int TNW_LOAD_WARPS = 8, TNW_LOADERS = 32 * TNW_LOAD_WARPS, TNW_THREADS = 32 * (TNW_LOAD_WARPS + 1);   // ncu (4 loader warps): the
// loaders' own instruction issue bounded the kernel (13.5 k warp-instructions per tile on 4 warps, 8 % warps active)

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
  const uint32_t n = valid ? 16u : 0u;          // src-size 0: the 16 destination bytes are zero-filled, src is not read
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ int tnw_fdiv(int n, uint32_t m, int sh) { return (int)(((uint64_t)(uint32_t)n * (uint64_t)m) >> sh); }

__global__ void __launch_bounds__(TNW_THREADS, 1) tn_gemm_ws_kernel(TnWsArgs w) {
  const TnArgs& a = w.t;
  extern __shared__ __align__(128) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5;
  const int MT = w.MT, NS = w.NS;
  const int m0 = blockIdx.y * MT * 128;
  const int mrows = min(MT * 128, a.Ma - m0);            // valid rows of dW handled by this CTA
  const int m8n = mrows >> 3;
  const int n0 = blockIdx.z * TN_NCHUNK;
  const int nb = min(TN_NCHUNK, a.Nb - n0), nb8 = nb >> 3;
  const bool ones = a.ones && blockIdx.z == 0;
  const int ncols = nb + (ones ? 16 : 0);
  // MN-major canonical layout: 16 B chunk (8 channels of voxel v) of channel-group g at g*SBO + v*16; the SBO padding spreads
  // the copies of one voxel's channel groups over distinct banks
  const uint32_t sbo = 2048 + 16;
  const uint32_t bytesA = (uint32_t)(MT * 16) * sbo, bytesB = (uint32_t)((TN_NCHUNK >> 3) + 2) * sbo;
  const uint32_t stage = bytesA + bytesB;
  uint8_t* sStage = smem;                                                      // NS x (A | B)
  int* sRowA = reinterpret_cast<int*>(smem + (size_t)NS * stage);              // [NS][128] source row of A (-1: zero row)
  int* sRowB = sRowA + NS * 128;                                               // [NS][128]
  float* sScale = reinterpret_cast<float*>(sRowB + NS * 128);                  // [N][nb] GroupNorm affine of the B channels
  float* sShift = sScale + (a.stats ? a.N * TN_NCHUNK : 0);
  uint64_t* full = reinterpret_cast<uint64_t*>(sShift + (a.stats ? a.N * TN_NCHUNK : 0));
  uint64_t* empty = full + 4;
  uint64_t* done = empty + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);

  const uint32_t tmem_cols = tmem_cols_pow2((uint32_t)(MT * ncols));
  if (warp == TNW_LOAD_WARPS) tmem_alloc(tmem_slot, tmem_cols);
  if (tid == 0) {
    for (int i = 0; i < NS; ++i) { mbar_init(&full[i], 32 * TNW_LOAD_WARPS); mbar_init(&empty[i], 1); }
    mbar_init(done, 1);
    fence_mbar_init();
  }
  // zero every stage once: channel groups beyond the valid ones stay zero for the whole kernel
  for (uint32_t i = tid * 16; i < (uint32_t)NS * stage; i += TNW_THREADS * 16) *reinterpret_cast<uint4*>(smem + i) = make_uint4(0, 0, 0, 0);
  if (a.stats != nullptr) {
    for (int i = tid; i < a.N * nb; i += TNW_THREADS) {
      const int n = i / nb, c = i - n * nb, cg = n0 + c;
      const double sm = a.stats[(int64_t)n * 2 * a.Nb + cg], q = a.stats[(int64_t)n * 2 * a.Nb + a.Nb + cg];
      const double mean = sm * (double)a.inv_count;
      double var = q * (double)a.inv_count - mean * mean;
      if (var < 0.0) var = 0.0;
      const float gg = a.gamma[cg] * (float)(1.0 / sqrt(var + 1e-5));
      sScale[n * TN_NCHUNK + c] = gg; sShift[n * TN_NCHUNK + c] = a.beta[cg] - (float)mean * gg;
    }
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t acc = *tmem_slot;
  const int tps = (int)((a.V + 127) / 128);           // tiles per sample
  const int64_t ntiles = (int64_t)tps * a.N;

  if (warp < TNW_LOAD_WARPS) {
    // ===================================================================== loaders
    const bool m8pow2 = (m8n & (m8n - 1)) == 0, n8pow2 = (nb8 & (nb8 - 1)) == 0;
    const int m8sh = __ffs(m8n) - 1, n8sh = __ffs(nb8) - 1;
    int64_t it = 0;
    for (int64_t g = blockIdx.x; g < ntiles; g += gridDim.x, ++it) {
      const int s = (int)(it % NS);
      if (it >= NS) mbar_wait(&empty[s], (uint32_t)((it / NS - 1) & 1));
      const int n = (int)(g / tps);
      const int v0 = ((int)(g - (int64_t)n * tps)) * 128;
      int* rowA = sRowA + s * 128;
      int* rowB = sRowB + s * 128;
      if (tid < 128) {   // source rows of voxel v0 + tid (element offsets of the row start, in uint4 units, fit 32 bits)
        const int v = v0 + tid;
        int ra = -1, rb = -1;
        if (v < (int)a.V) {
          if (a.mapA == MAP_IDENT && a.mapB == MAP_IDENT) { ra = v; rb = v; }
          else {
            const int t = tnw_fdiv(v, w.dm2, w.ds2), x = v - t * a.d2;
            const int z = tnw_fdiv(t, w.dm1, w.ds1), y = t - z * a.d1;
            auto mapped = [&](int kind, int s1, int s2) -> int {
              if (kind == MAP_IDENT) return v;
              if (kind == MAP_PLUS1) return ((z + 1) * s1 + (y + 1)) * s2 + (x + 1);
              if (kind == MAP_TIMES2) return ((2 * z) * s1 + 2 * y) * s2 + 2 * x;
              return ((2 * z + 1) * s1 + (2 * y + 1)) * s2 + (2 * x + 1);
            };
            ra = mapped(a.mapA, a.as1, a.as2);
            if (a.mapB >= 4) {   // dense-conv weight gradient: B row = input voxel feeding output voxel v through this tap
              int iz, iy, ix;
              bool ok = true;
              if (a.mapB == 4) {
                iz = z * a.tstride + a.tz - a.tpad; iy = y * a.tstride + a.ty - a.tpad; ix = x * a.tstride + a.tx - a.tpad;
              } else {
                const int qz = z + a.tpad - a.tz, qy = y + a.tpad - a.ty, qx = x + a.tpad - a.tx;
                ok = qz >= 0 && qy >= 0 && qx >= 0 && (qz % a.tstride == 0) && (qy % a.tstride == 0) && (qx % a.tstride == 0);
                iz = qz / a.tstride; iy = qy / a.tstride; ix = qx / a.tstride;
              }
              ok = ok && iz >= 0 && iz < a.bs0 && iy >= 0 && iy < a.bs1 && ix >= 0 && ix < a.bs2;
              rb = ok ? (iz * a.bs1 + iy) * a.bs2 + ix : -1;
            } else {
              rb = mapped(a.mapB, a.bs1, a.bs2);
            }
          }
        }
        rowA[tid] = ra < 0 ? -1 : ra * (int)a.a_pitch8;
        rowB[tid] = rb < 0 ? -1 : rb * (int)a.b_pitch8;
      }
      asm volatile("bar.sync 1, %0;" ::"n"(TNW_LOADERS) : "memory");     // the row tables of this stage (loader warps only)
      uint8_t* sA = sStage + (size_t)s * stage;
      uint8_t* sB = sA + bytesA;
      const uint4* An = a.A + (int64_t)n * a.a_sample8 + (m0 >> 3);
      const uint4* Bn = a.B + (int64_t)n * a.b_sample8 + (n0 >> 3);
      if (a.stats != nullptr) {
        // B through registers: GroupNorm affine of this sample on the way in
        const float* sc = sScale + n * TN_NCHUNK;
        const float* sh = sShift + n * TN_NCHUNK;
        staged_copy<8>(128 * nb8, tid, TNW_LOADERS,
            [&](int q) {
              const int v = n8pow2 ? (q >> n8sh) : (q / nb8), g8 = q - v * nb8;
              const int r = rowB[v];
              return r >= 0 ? __ldg(Bn + (r + g8)) : make_uint4(0, 0, 0, 0);
            },
            [&](int q, const uint4& raw) {
              const int v = n8pow2 ? (q >> n8sh) : (q / nb8), g8 = q - v * nb8;
              uint4 val = raw;
              if (rowB[v] >= 0) {
                float f[8];
                unpack8(raw, f);
#pragma unroll
                for (int j = 0; j < 8; ++j) f[j] = fmaf(f[j], sc[g8 * 8 + j], sh[g8 * 8 + j]);
                val = pack8(f);
              }
              *reinterpret_cast<uint4*>(sB + g8 * sbo + v * 16) = val;
            });
      } else {
        const uint32_t sBu = smem_u32(sB);
#pragma unroll 4
        for (int q = tid; q < 128 * nb8; q += TNW_LOADERS) {
          const int v = n8pow2 ? (q >> n8sh) : (q / nb8), g8 = q - v * nb8;
          const int r = rowB[v];
          cp_async16(sBu + g8 * sbo + v * 16, Bn + (max(r, 0) + g8), r >= 0);
        }
      }
      if (ones && tid < 128) *reinterpret_cast<uint4*>(sB + nb8 * sbo + tid * 16) = make_uint4((v0 + tid < (int)a.V) ? 0x3F80u : 0u, 0, 0, 0);
      {
        const uint32_t sAu = smem_u32(sA);
#pragma unroll 4
        for (int q = tid; q < 128 * m8n; q += TNW_LOADERS) {
          const int v = m8pow2 ? (q >> m8sh) : (q / m8n), g8 = q - v * m8n;
          const int r = rowA[v];
          cp_async16(sAu + g8 * sbo + v * 16, An + (max(r, 0) + g8), r >= 0);
        }
      }
      // hand-off, one tile behind: the copies of tile it stay in flight while those of tile it-1 are drained; cp.async writes
      // through the generic proxy, so the writer fences before it arrives (the MMA reads through the async proxy)
      cp_async_commit();
      if (it >= 1) {
        cp_async_wait<1>();
        fence_proxy_async_smem();
        mbar_arrive(&full[(int)((it - 1) % NS)]);
      }
    }
    if (it >= 1) {
      cp_async_wait<0>();
      fence_proxy_async_smem();
      mbar_arrive(&full[(int)((it - 1) % NS)]);
    }
  } else if (tid == TNW_LOAD_WARPS * 32) {
    // ===================================================================== MMA issuer (one thread)
    const uint32_t idesc = umma_idesc_bf16(128, ncols, 1, 1);
    int64_t it = 0;
    for (int64_t g = blockIdx.x; g < ntiles; g += gridDim.x, ++it) {
      const int s = (int)(it % NS);
      mbar_wait(&full[s], (uint32_t)((it / NS) & 1));
      tc_fence_after();
      const uint32_t sA = smem_u32(sStage + (size_t)s * stage), sB = sA + bytesA;
      const uint64_t bd = umma_desc(sB, 128, sbo);
      for (int mt = 0; mt < MT; ++mt) {
        if (mt * 128 >= mrows) break;
        const uint64_t ad = umma_desc(sA + (uint32_t)(mt * 16) * sbo, 128, sbo);
        for (int k = 0; k < 8; ++k)   // 128 voxels = 8 x K16; K-groups of 8 voxels are LBO = 128 B apart
          umma_bf16(acc + mt * ncols, ad + (uint64_t)(k * 16), bd + (uint64_t)(k * 16), idesc, (it > 0 || k > 0) ? 1u : 0u);
      }
      tc_commit(&empty[s]);
    }
    tc_commit(done);
  }
  // ---- partial sums of this CTA (warps 0-3: one TMEM lane quarter each)
  if (warp < 4) {
    mbar_wait(done, 0);
    tc_fence_after();
    for (int mt = 0; mt < MT; ++mt) {
      const int row = mt * 128 + tid;
      if (mt * 128 >= mrows) break;
      const uint32_t trow = acc + mt * ncols + ((uint32_t)(warp * 32) << 16);
      float* prow = a.part + ((int64_t)blockIdx.x * a.Mtot + m0 + row) * a.Ncols_tot;
      for (int c16 = 0; c16 < ncols / 16; ++c16) {
        uint32_t v[16];
        tmem_ld16(trow + c16 * 16, v);
        tmem_ld_wait();
        if (row < mrows) {
          const int col0 = (c16 * 16 < nb) ? (n0 + c16 * 16) : a.Nb;   // ones block lives after all Nb columns
          float4* dst = reinterpret_cast<float4*>(prow + col0);
          dst[0] = make_float4(__uint_as_float(v[0]), __uint_as_float(v[1]), __uint_as_float(v[2]), __uint_as_float(v[3]));
          dst[1] = make_float4(__uint_as_float(v[4]), __uint_as_float(v[5]), __uint_as_float(v[6]), __uint_as_float(v[7]));
          dst[2] = make_float4(__uint_as_float(v[8]), __uint_as_float(v[9]), __uint_as_float(v[10]), __uint_as_float(v[11]));
          dst[3] = make_float4(__uint_as_float(v[12]), __uint_as_float(v[13]), __uint_as_float(v[14]), __uint_as_float(v[15]));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == TNW_LOAD_WARPS) tmem_dealloc(acc, tmem_cols);
}

// second stage: dW[m*ldm + n*ldn] = sum_p part[p][m][n] (n < Nw) ; db[m] = sum_p part[p][m][Nw]
// 256 threads = 32 outputs x 8 partial-index lanes; fixed summation order -> run-to-run deterministic.
__global__ void __launch_bounds__(256) reduce_partials_kernel(const float* __restrict__ part, int P, int M, int Mtot,
                                                              int Ncols_tot, int Nw, float* __restrict__ dW, int64_t ldm,
                                                              int64_t ldn, float* __restrict__ db) {
  __shared__ float s_red[8][33];
  const int ncol = Nw + (db ? 1 : 0);
  const int64_t total = (int64_t)M * ncol;
  const int lane_o = threadIdx.x & 31, pg = threadIdx.x >> 5;
  for (int64_t base = (int64_t)blockIdx.x * 32; base < total; base += (int64_t)gridDim.x * 32) {
    const int64_t i = base + lane_o;
    float s = 0.f;
    int m = 0, nn = 0;
    if (i < total) {
      m = (int)(i / ncol); nn = (int)(i - (int64_t)m * ncol);
      for (int p = pg; p < P; p += 8) s += part[((int64_t)p * Mtot + m) * Ncols_tot + nn];
    }
    s_red[pg][lane_o] = s;
    __syncthreads();
    if (pg == 0 && i < total) {
      float t = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) t += s_red[k][lane_o];
      if (nn < Nw) dW[m * ldm + nn * ldn] = t;
      else db[m] = t;
    }
    __syncthreads();
  }
}

// ============================================================================ row-gather 1x1 conv (K-major)
struct PwArgs {
  const uint4* A; const uint4* W; const float* bias; uint4* out;
  int K, Nw, KC, NT;
  int map, d1, d2, s1, s2;
  int64_t Vout, Vin;
};

__global__ void __launch_bounds__(128) pw_kernel(PwArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5;
  const int n = blockIdx.y, nt = blockIdx.z;
  const int64_t tile0 = (int64_t)blockIdx.x * 128;
  uint8_t* sA = smem;                       // [128 x KC]
  uint8_t* sW = sA + 128 * a.KC * 2;        // [NT x KC]
  uint64_t* bar = reinterpret_cast<uint64_t*>(sW + a.NT * a.KC * 2);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  const uint32_t tmem_cols = tmem_cols_pow2(a.NT);
  if (warp == 0) tmem_alloc(tmem_slot, tmem_cols);
  if (tid == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t acc = *tmem_slot;
  const uint32_t idesc = umma_idesc_bf16(128, a.NT, 0, 0);
  const int kc8 = a.KC >> 3;
  uint32_t ph = 0;
  const uint4* An = a.A + (int64_t)n * a.Vin * (a.K >> 3);
  for (int kc = 0; kc < a.K / a.KC; ++kc) {
    const uint32_t sbo = kc8 * 128;
    staged_copy<8>(128 * kc8, tid, 128,
        [&](int q) {
          const int r = q >> __ffs(kc8) - 1, c8 = q & (kc8 - 1);
          if (tile0 + r >= a.Vout) return make_uint4(0, 0, 0, 0);
          return __ldg(An + map_row(a.map, tile0 + r, a.d1, a.d2, a.s1, a.s2) * (a.K >> 3) + kc * kc8 + c8);
        },
        [&](int q, const uint4& v) {
          const int r = q >> __ffs(kc8) - 1, c8 = q & (kc8 - 1);
          *reinterpret_cast<uint4*>(sA + (r >> 3) * sbo + c8 * 128 + (r & 7) * 16) = v;
        });
    stage_rows_k(sW, a.W + (int64_t)nt * a.NT * (a.K >> 3) + kc * kc8, a.NT, kc8, a.K >> 3, tid);
    fence_proxy_async_smem();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      const uint64_t ad = umma_desc(smem_u32(sA), 128, kc8 * 128), bd = umma_desc(smem_u32(sW), 128, kc8 * 128);
      for (int k = 0; k < a.KC / 16; ++k)
        umma_bf16(acc, ad + (uint64_t)(k * 16), bd + (uint64_t)(k * 16), idesc, (kc > 0 || k > 0) ? 1u : 0u);
      tc_commit(bar);
    }
    mbar_wait(bar, ph); ph ^= 1;
  }
  tc_fence_after();
  {
    const int64_t r = tile0 + tid;
    const uint32_t trow = acc + ((uint32_t)(warp * 32) << 16);
    const int64_t orow = ((int64_t)n * a.Vout + r) * (a.Nw >> 3) + nt * (a.NT >> 3);
    for (int c16 = 0; c16 < a.NT / 16; ++c16) {
      uint32_t v[16];
      tmem_ld16(trow + c16 * 16, v);
      tmem_ld_wait();
      if (r >= a.Vout) continue;
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

// ============================================================================ GroupNorm backward -> dy
// dy = rstd*gamma*(g - S1/V - xhat*S2/V), xhat = (y-mean)*rstd ; dsum[c] += sum dy (conv1 bias grad)
__global__ void __launch_bounds__(256) gn_dy_kernel(const uint4* __restrict__ g, const uint4* __restrict__ y,
                                                    const double* __restrict__ stats, const double* __restrict__ gstats,
                                                    const float* __restrict__ gamma, uint4* __restrict__ dy,
                                                    double* __restrict__ dsum, int C, int64_t V, float inv_count) {
  // dy = A*g + B*y + D per channel with A = gamma*rstd, B = -A*rstd*S2/V, D = A*(mean*rstd*S2/V - S1/V): the three
  // coefficients are formed once per CTA in f64 and (power-of-two C/8) held in registers by the thread that owns the
  // 8-channel chunk, so the streaming loop is 2 loads, 16 FMAs and 1 store per 8 elements.
  extern __shared__ double s_sum[];   // [C] doubles, then 3*[C] floats
  float* s_k = reinterpret_cast<float*>(s_sum + C);
  const int CH = C >> 3, n = blockIdx.y;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    s_sum[c] = 0.0;
    const double sm = stats[(int64_t)n * 2 * C + c], q = stats[(int64_t)n * 2 * C + C + c];
    const double mean = sm * (double)inv_count;
    double var = q * (double)inv_count - mean * mean;
    if (var < 0.0) var = 0.0;
    const double rstd = (double)(float)(1.0 / sqrt(var + 1e-5));
    const double A = (double)gamma[c] * rstd;
    const double k1 = gstats[(int64_t)n * 2 * C + c] * (double)inv_count;
    const double k2 = gstats[(int64_t)n * 2 * C + C + c] * (double)inv_count;
    s_k[c] = (float)A; s_k[C + c] = (float)(-A * rstd * k2); s_k[2 * C + c] = (float)(A * (mean * rstd * k2 - k1));
  }
  __syncthreads();
  const int64_t items = V * CH;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t first = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const uint4* gn = g + (int64_t)n * items;
  const uint4* yn = y + (int64_t)n * items;
  uint4* dn = dy + (int64_t)n * items;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (stride % CH == 0) {            // the thread's channel chunk never changes
    const int cc = (int)(first % CH);
    float ka[8], kb[8], kd[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { ka[j] = s_k[cc * 8 + j]; kb[j] = s_k[C + cc * 8 + j]; kd[j] = s_k[2 * C + cc * 8 + j]; }
    for (int64_t i = first; i < items; i += stride) {
      float gv[8], yv[8], o[8];
      unpack8(ldg_nc(gn + i), gv);
      unpack8(ldg_nc(yn + i), yv);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        o[j] = round_bf16(fmaf(ka[j], gv[j], fmaf(kb[j], yv[j], kd[j])));
        acc[j] += o[j];
      }
      dn[i] = pack8(o);
    }
    if (first < items) {
#pragma unroll
      for (int j = 0; j < 8; ++j) atomicAdd(&s_sum[cc * 8 + j], (double)acc[j]);
    }
  } else {
    for (int64_t i = first; i < items; i += stride) {
      const int cc = (int)(i % CH);
      float gv[8], yv[8], o[8];
      unpack8(ldg_nc(gn + i), gv);
      unpack8(ldg_nc(yn + i), yv);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = cc * 8 + j;
        o[j] = round_bf16(fmaf(s_k[c], gv[j], fmaf(s_k[C + c], yv[j], s_k[2 * C + c])));
        atomicAdd(&s_sum[c], (double)o[j]);
      }
      dn[i] = pack8(o);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += blockDim.x) atomicAdd(&dsum[i], s_sum[i]);
}

// ============================================================================ depthwise weight gradient
// dW[tap][c] += sum_v center[v,c] * neigh[S*v - P + tap, c]; one (dz,dy) tap row per blockIdx.y
struct DwWgArgs {
  int c0, c1, c2;   // center spatial size
  int n0, n1, n2;   // neighbour spatial size
  int C, S;
};

template <int K>
__global__ void __launch_bounds__(256, K == 3 ? 3 : 1) dw_wgrad_kernel(const uint4* __restrict__ center, const uint4* __restrict__ neigh,
                                                       double* __restrict__ dW, DwWgArgs a) {
  extern __shared__ double s_acc[];   // [K][C]
  constexpr int P = K / 2;
  const int C = a.C, CH = C >> 3;
  // tap row FASTEST in the launch order: the K*K CTAs that walk the same voxels with different (dz,dy) are co-resident,
  // so center / neighbour rows are fetched from DRAM once and re-read through L2 (with the tap row on blockIdx.y every
  // tap row streamed the whole tensor again: ncu 1.67 GB of DRAM traffic for 0.58 GB of operands)
  const int taprow = blockIdx.x % (K * K), sblk = blockIdx.x / (K * K), nsblk = gridDim.x / (K * K);
  const int dz = taprow / K, dyy = taprow % K;
  const int n = blockIdx.z;
  for (int i = threadIdx.x; i < K * C; i += blockDim.x) s_acc[i] = 0.0;
  __syncthreads();
  const int nstrip = (a.c2 + DW_XB_WG - 1) / DW_XB_WG;
  const int64_t items = (int64_t)a.c0 * a.c1 * nstrip * CH;
  float acc[K][8];
#pragma unroll
  for (int k = 0; k < K; ++k)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[k][j] = 0.f;
  const int64_t stride = (int64_t)nsblk * blockDim.x;
  const uint4* cn = center + (int64_t)n * a.c0 * a.c1 * a.c2 * CH;
  const uint4* nn = neigh + (int64_t)n * a.n0 * a.n1 * a.n2 * CH;
  int cc_fixed = -1;
  for (int64_t item = sblk * (int64_t)blockDim.x + threadIdx.x; item < items; item += stride) {
    const int cc = (int)(item % CH);
    if (cc_fixed >= 0 && cc != cc_fixed) {
#pragma unroll
      for (int k = 0; k < K; ++k)
#pragma unroll
        for (int j = 0; j < 8; ++j) { atomicAdd(&s_acc[k * C + cc_fixed * 8 + j], (double)acc[k][j]); acc[k][j] = 0.f; }
    }
    cc_fixed = cc;
    int64_t t = item / CH;
    const int xs = (int)(t % nstrip); t /= nstrip;
    const int cy = (int)(t % a.c1), cz = (int)(t / a.c1);
    const int iz = cz * a.S - P + dz, iy = cy * a.S - P + dyy;
    if (iz < 0 || iz >= a.n0 || iy < 0 || iy >= a.n1) continue;
    const uint4* crow = cn + ((int64_t)cz * a.c1 + cy) * a.c2 * CH + cc;
    const uint4* nrow = nn + ((int64_t)iz * a.n1 + iy) * a.n2 * CH + cc;
    // every load of the strip is issued before the first FMA (branch-free, zero fill outside the rows): the kernel is
    // latency-bound otherwise (ncu: long_scoreboard 17 warps per issue at 48 % occupancy)
    constexpr int NIN = 2 * (DW_XB_WG - 1) + K;          // neighbour vectors a strip can touch at stride 2
    const int cx0 = xs * DW_XB_WG;
    const int nin = a.S * (DW_XB_WG - 1) + K;
    const int ix0 = cx0 * a.S - P;
    uint4 cv4[DW_XB_WG], nv4[NIN];
#pragma unroll
    for (int j = 0; j < DW_XB_WG; ++j)
      cv4[j] = (cx0 + j < a.c2) ? __ldg(crow + (int64_t)(cx0 + j) * CH) : make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int i = 0; i < NIN; ++i) {
      const int ix = ix0 + i;
      nv4[i] = (i < nin && ix >= 0 && ix < a.n2) ? __ldg(nrow + (int64_t)ix * CH) : make_uint4(0, 0, 0, 0);
    }
    if (a.S == 1) {
#pragma unroll
      for (int j = 0; j < DW_XB_WG; ++j) {
        float cv[8];
        unpack8(cv4[j], cv);
#pragma unroll
        for (int k = 0; k < K; ++k) {
          if (j + k < NIN) {
            float nv[8];
            unpack8(nv4[j + k], nv);
#pragma unroll
            for (int c = 0; c < 8; ++c) acc[k][c] = fmaf(cv[c], nv[c], acc[k][c]);
          }
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < DW_XB_WG; ++j) {
        float cv[8];
        unpack8(cv4[j], cv);
#pragma unroll
        for (int k = 0; k < K; ++k) {
          if (2 * j + k < NIN) {
            float nv[8];
            unpack8(nv4[2 * j + k], nv);
#pragma unroll
            for (int c = 0; c < 8; ++c) acc[k][c] = fmaf(cv[c], nv[c], acc[k][c]);
          }
        }
      }
    }
  }
  if (cc_fixed >= 0) {
#pragma unroll
    for (int k = 0; k < K; ++k)
#pragma unroll
      for (int j = 0; j < 8; ++j) atomicAdd(&s_acc[k * C + cc_fixed * 8 + j], (double)acc[k][j]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < K * C; i += blockDim.x)
    atomicAdd(&dW[(int64_t)((dz * K + dyy) * K) * C + i], s_acc[i]);
}

// ---------------------------------------------------------------------------- tiled SAME-mode dw wgrad
// Persistent CTAs loop over 4x8x16 bricks (32 channels): the dy brick and the x brick (+halo) are staged in
// shared memory once, thread (tap row (dz,dy), channel chunk, row partition) slides along W keeping the K
// x-taps in registers, so every staged voxel is read ~K^2/partitions times from shared memory instead of
// K^2 times from L2.  Accumulators live in registers across bricks; one f64 reduction per CTA at the end.
constexpr int WT_Z = 4, WT_Y = 8, WT_X = 16;

template <int K>
__global__ void __launch_bounds__(256, 2) dw_wgrad_same_tiled_kernel(const uint4* __restrict__ dy, const uint4* __restrict__ x,
                                                                  double* __restrict__ dW, int D, int H, int W, int C,
                                                                  int tiles_y, int tiles_x, int nbricks, int N,
                                                                  const __grid_constant__ CUtensorMap tmap_x,
                                                                  const __grid_constant__ CUtensorMap tmap_dy, int use_tma) {
  constexpr int P = K / 2;
  constexpr int BZ = WT_Z + 2 * P, BY = WT_Y + 2 * P, BX = WT_X + 2 * P, PITCH = BX + 1;
  constexpr int NPART = 256 / (K * K * 4);      // row partitions (threads beyond K*K*4*NPART idle)
  extern __shared__ __align__(128) uint8_t dsm[];
  uint4* s_x = reinterpret_cast<uint4*>(dsm);                      // [BZ][BY][PITCH][4]
  uint4* s_dy = s_x + BZ * BY * PITCH * 4;                         // [WT_Z][WT_Y][WT_X][4]
  double* s_red = reinterpret_cast<double*>(s_dy + WT_Z * WT_Y * WT_X * 4);   // [K^3][32]
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_red + K * K * K * 32);      // TMA completion barrier
  // channel group fastest in launch order: the groups of one brick are co-resident and share its 128 B lines in L2
  const int ncg = C >> 5, cg = blockIdx.x % ncg, bx0 = blockIdx.x / ncg, nbx = gridDim.x / ncg;
  const int tid = threadIdx.x, CH = C >> 3;
  const int cc = tid & 3, tr = (tid >> 2) % (K * K), part = (tid >> 2) / (K * K);
  const int dz = tr / K, dyy = tr % K;
  const bool worker = part < NPART;
  uint64_t acc[K][4];
#pragma unroll
  for (int k = 0; k < K; ++k)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[k][c] = 0ull;
  for (int i = tid; i < K * K * K * 32; i += 256) s_red[i] = 0.0;
  if (use_tma && tid == 0) { mbar_init(s_bar, 1); fence_mbar_init(); fence_proxy_async_smem(); }
  uint32_t tma_phase = 0;

  for (int b = bx0; b < nbricks * N; b += nbx) {
    const int n = b / nbricks;
    int t = b - n * nbricks;
    const int tx = t % tiles_x; t /= tiles_x;
    const int ty = t % tiles_y, tz = t / tiles_y;
    const int z0 = tz * WT_Z, y0 = ty * WT_Y, x0 = tx * WT_X;
    const uint4* xn = x + (int64_t)n * D * H * W * CH + cg * 4;
    const uint4* dn = dy + (int64_t)n * D * H * W * CH + cg * 4;
    __syncthreads();   // previous brick fully consumed
    if (use_tma) {
      // both bricks arrive as bulk tensor copies (zero fill outside the volume) on one mbarrier
      if (tid == 0) {
        mbar_arrive_expect_tx(s_bar, (uint32_t)(BZ * BY * PITCH * 64 + WT_Z * WT_Y * WT_X * 64));
        tma_load_5d(s_x, &tmap_x, cg * 32, x0 - P, y0 - P, z0 - P, n, s_bar);
        tma_load_5d(s_dy, &tmap_dy, cg * 32, x0, y0, z0, n, s_bar);
      }
      mbar_wait(s_bar, tma_phase);
      tma_phase ^= 1;
    } else {
    staged_copy<((BZ * BY * BX * 4 + 255) / 256 <= 17 ? (BZ * BY * BX * 4 + 255) / 256 : 8)>(BZ * BY * BX * 4, tid, 256,
        [&](int q) {
          const int c4 = q & 3, v = q >> 2;
          const int bx = v % BX, by = (v / BX) % BY, bz = v / (BX * BY);
          const int gz = z0 + bz - P, gy = y0 + by - P, gx = x0 + bx - P;
          if (gz < 0 || gz >= D || gy < 0 || gy >= H || gx < 0 || gx >= W) return make_uint4(0, 0, 0, 0);
          return __ldg(xn + (((int64_t)gz * H + gy) * W + gx) * CH + c4);
        },
        [&](int q, const uint4& v4) {
          const int c4 = q & 3, v = q >> 2;
          const int bx = v % BX, by = (v / BX) % BY, bz = v / (BX * BY);
          s_x[((bz * BY + by) * PITCH + bx) * 4 + c4] = v4;
        });
    staged_copy<8>(WT_Z * WT_Y * WT_X * 4, tid, 256,
        [&](int q) {
          const int c4 = q & 3, v = q >> 2;
          const int bx = v % WT_X, by = (v / WT_X) % WT_Y, bz = v / (WT_X * WT_Y);
          const int gz = z0 + bz, gy = y0 + by, gx = x0 + bx;
          if (gz >= D || gy >= H || gx >= W) return make_uint4(0, 0, 0, 0);
          return __ldg(dn + (((int64_t)gz * H + gy) * W + gx) * CH + c4);
        },
        [&](int q, const uint4& v4) { s_dy[q] = v4; });
    }
    __syncthreads();
    if (worker) {
      for (int r = part; r < WT_Z * WT_Y; r += NPART) {     // (z,y) rows of the brick owned by this partition
        const int lz = r / WT_Y, ly = r % WT_Y;
        const uint4* xrow = s_x + (((lz + dz) * BY + (ly + dyy)) * PITCH) * 4 + cc;
        const uint4* drow = s_dy + ((lz * WT_Y + ly) * WT_X) * 4 + cc;
        uint64_t win[K][4];                                  // sliding window of K x-taps
#pragma unroll
        for (int k = 0; k < K - 1; ++k) {
          const uint4 v4 = xrow[k * 4];
          win[k + 1][0] = pk2(bf16_lo(v4.x), bf16_hi(v4.x)); win[k + 1][1] = pk2(bf16_lo(v4.y), bf16_hi(v4.y));
          win[k + 1][2] = pk2(bf16_lo(v4.z), bf16_hi(v4.z)); win[k + 1][3] = pk2(bf16_lo(v4.w), bf16_hi(v4.w));
        }
#pragma unroll 4
        for (int xx = 0; xx < WT_X; ++xx) {
#pragma unroll
          for (int k = 0; k < K - 1; ++k)
#pragma unroll
            for (int c = 0; c < 4; ++c) win[k][c] = win[k + 1][c];
          const uint4 v4 = xrow[(xx + K - 1) * 4];
          win[K - 1][0] = pk2(bf16_lo(v4.x), bf16_hi(v4.x)); win[K - 1][1] = pk2(bf16_lo(v4.y), bf16_hi(v4.y));
          win[K - 1][2] = pk2(bf16_lo(v4.z), bf16_hi(v4.z)); win[K - 1][3] = pk2(bf16_lo(v4.w), bf16_hi(v4.w));
          const uint4 d4 = drow[xx * 4];
          uint64_t d[4];
          d[0] = pk2(bf16_lo(d4.x), bf16_hi(d4.x)); d[1] = pk2(bf16_lo(d4.y), bf16_hi(d4.y));
          d[2] = pk2(bf16_lo(d4.z), bf16_hi(d4.z)); d[3] = pk2(bf16_lo(d4.w), bf16_hi(d4.w));
#pragma unroll
          for (int k = 0; k < K; ++k)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[k][c] = fma2(d[c], win[k][c], acc[k][c]);
        }
      }
    }
  }
  __syncthreads();
  if (worker) {
#pragma unroll
    for (int k = 0; k < K; ++k)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float a0, a1;
        upk2(acc[k][c], a0, a1);
        atomicAdd(&s_red[((dz * K + dyy) * K + k) * 32 + cc * 8 + 2 * c], (double)a0);
        atomicAdd(&s_red[((dz * K + dyy) * K + k) * 32 + cc * 8 + 2 * c + 1], (double)a1);
      }
  }
  __syncthreads();
  for (int i = tid; i < K * K * K * 32; i += 256) atomicAdd(&dW[(int64_t)(i >> 5) * C + cg * 32 + (i & 31)], s_red[i]);
}

template <int K>
static bool launch_dw_wgrad_tiled(cudaStream_t st, const uint4* dy, const uint4* x, double* dW, int D, int H, int W, int C, int N) {
  constexpr int P = K / 2;
  const size_t smem = (size_t)(WT_Z + 2 * P) * (WT_Y + 2 * P) * (WT_X + 2 * P + 1) * 64 + (size_t)WT_Z * WT_Y * WT_X * 64 +
                      (size_t)K * K * K * 32 * 8 + 16;
  if (smem > 227 * 1024) return false;
  CUtensorMap tmx, tmd;
  memset(&tmx, 0, sizeof(tmx)); memset(&tmd, 0, sizeof(tmd));
  static const bool no_tma = getenv("PCB_NO_TMA") != nullptr;
  const int use_tma = !no_tma && make_brick_tensor_map(&tmx, x, N, D, H, W, C, WT_Z + 2 * P, WT_Y + 2 * P, WT_X + 2 * P + 1) &&
                      make_brick_tensor_map(&tmd, dy, N, D, H, W, C, WT_Z, WT_Y, WT_X);
  static DevFlag configured;
  if (!configured) {
    cudaFuncSetAttribute(dw_wgrad_same_tiled_kernel<K>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (cudaFuncSetAttribute(dw_wgrad_same_tiled_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) {
      cudaGetLastError();
      return false;
    }
    configured = true;
  }
  const int tz = (D + WT_Z - 1) / WT_Z, ty = (H + WT_Y - 1) / WT_Y, tx = (W + WT_X - 1) / WT_X;
  const int64_t nb = (int64_t)tz * ty * tx;
  if (nb * N >= (1ll << 31)) return false;
  const int ncg = C / 32;
  int ctas = (148 * 2) / ncg;                 // two resident CTAs per SM over all channel groups
  if (ctas < 1) ctas = 1;
  if (nb * N < ctas) ctas = (int)(nb * N);
  dim3 grid((unsigned)(ctas * ncg));
  dw_wgrad_same_tiled_kernel<K><<<grid, 256, smem, st>>>(dy, x, dW, D, H, W, C, ty, tx, (int)nb, N, tmx, tmd, use_tma);
  return true;
}

// ---------------------------------------------------------------------------- tiled stride-2 depthwise weight gradient (k = 3)
// dW[tap][c] += sum_v center[v, c] * fine[2 v - 1 + tap, c]: the weight gradient of the stride-2 conv (center = dY on the coarse
// grid, fine = x) and of the transposed stride-2 conv (center = x on the coarse grid, fine = dY).  Same decomposition as
// dw_wgrad_same_tiled_kernel: a 2x7x8 coarse tile and its 5x15x17 fine brick (x 32 channels) are staged once; thread = (tap row
// (dz,dy), 8-channel chunk, row partition) walks the coarse rows along W with the three x-taps of the fine row in registers (two
// new fine vectors per coarse voxel).  The untiled kernel launched one grid per (dz,dy) tap row, so both tensors were re-streamed
// from L2 nine times (up_0: 21 GB of L2 traffic for 2.3 GB of operands, 1.26 ms per batch-4 step).
constexpr int W2_Z = 2, W2_Y = 7, W2_X = 8;
constexpr int W2_BZ = 2 * W2_Z + 1, W2_BY = 2 * W2_Y + 1, W2_BX = 2 * W2_X + 1;

__global__ void __launch_bounds__(256, 2) dw_wgrad_s2_tiled_kernel(const uint4* __restrict__ center, const uint4* __restrict__ fine,
                                                                double* __restrict__ dW, int c0, int c1, int c2, int f0, int f1,
                                                                int f2, int C, int tiles_y, int tiles_x, int nbricks, int N) {
  constexpr int K = 3, NPART = 7;
  extern __shared__ __align__(128) uint8_t dsm[];
  uint4* s_f = reinterpret_cast<uint4*>(dsm);                               // [BZ][BY][BX][4]
  uint4* s_c = s_f + W2_BZ * W2_BY * W2_BX * 4;                             // [Z][Y][X][4]
  double* s_red = reinterpret_cast<double*>(s_c + W2_Z * W2_Y * W2_X * 4);  // [27][32]
  const int ncg = C >> 5, cg = blockIdx.x % ncg, bx0 = blockIdx.x / ncg, nbx = gridDim.x / ncg;
  const int tid = threadIdx.x, CH = C >> 3;
  const int cc = tid & 3, tr = (tid >> 2) % (K * K), part = (tid >> 2) / (K * K);
  const int dz = tr / K, dyy = tr % K;
  const bool worker = part < NPART;
  uint64_t acc[K][4];
#pragma unroll
  for (int k = 0; k < K; ++k)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[k][c] = 0ull;
  for (int i = tid; i < K * K * K * 32; i += 256) s_red[i] = 0.0;
  for (int b = bx0; b < nbricks * N; b += nbx) {
    const int n = b / nbricks;
    int t = b - n * nbricks;
    const int tx = t % tiles_x; t /= tiles_x;
    const int ty = t % tiles_y, tz = t / tiles_y;
    const int z0 = tz * W2_Z, y0 = ty * W2_Y, x0 = tx * W2_X;
    const uint4* fn = fine + (int64_t)n * f0 * f1 * f2 * CH + cg * 4;
    const uint4* cn = center + (int64_t)n * c0 * c1 * c2 * CH + cg * 4;
    __syncthreads();   // previous brick fully consumed
    staged_copy<10>(W2_BZ * W2_BY * W2_BX * 4, tid, 256,
        [&](int q) {
          const int c4 = q & 3, v = q >> 2;
          const int bx = v % W2_BX, by = (v / W2_BX) % W2_BY, bz = v / (W2_BX * W2_BY);
          const int gz = 2 * z0 - 1 + bz, gy = 2 * y0 - 1 + by, gx = 2 * x0 - 1 + bx;
          if (gz < 0 || gz >= f0 || gy < 0 || gy >= f1 || gx < 0 || gx >= f2) return make_uint4(0, 0, 0, 0);
          return __ldg(fn + (((int64_t)gz * f1 + gy) * f2 + gx) * CH + c4);
        },
        [&](int q, const uint4& v4) { s_f[q] = v4; });
    staged_copy<2>(W2_Z * W2_Y * W2_X * 4, tid, 256,
        [&](int q) {
          const int c4 = q & 3, v = q >> 2;
          const int bx = v % W2_X, by = (v / W2_X) % W2_Y, bz = v / (W2_X * W2_Y);
          const int gz = z0 + bz, gy = y0 + by, gx = x0 + bx;
          if (gz >= c0 || gy >= c1 || gx >= c2) return make_uint4(0, 0, 0, 0);
          return __ldg(cn + (((int64_t)gz * c1 + gy) * c2 + gx) * CH + c4);
        },
        [&](int q, const uint4& v4) { s_c[q] = v4; });
    __syncthreads();
    if (worker) {
      for (int r = part; r < W2_Z * W2_Y; r += NPART) {     // coarse (z,y) rows owned by this partition
        const int lz = r / W2_Y, ly = r % W2_Y;
        const uint4* frow = s_f + (((2 * lz + dz) * W2_BY + (2 * ly + dyy)) * W2_BX) * 4 + cc;
        const uint4* crow = s_c + ((lz * W2_Y + ly) * W2_X) * 4 + cc;
        uint64_t w0[4], w1[4], w2[4];
        {
          const uint4 v4 = frow[0];
          w2[0] = pk2(bf16_lo(v4.x), bf16_hi(v4.x)); w2[1] = pk2(bf16_lo(v4.y), bf16_hi(v4.y));
          w2[2] = pk2(bf16_lo(v4.z), bf16_hi(v4.z)); w2[3] = pk2(bf16_lo(v4.w), bf16_hi(v4.w));
        }
#pragma unroll 2
        for (int xx = 0; xx < W2_X; ++xx) {
#pragma unroll
          for (int c = 0; c < 4; ++c) w0[c] = w2[c];
          const uint4 a4 = frow[(2 * xx + 1) * 4], b4 = frow[(2 * xx + 2) * 4];
          w1[0] = pk2(bf16_lo(a4.x), bf16_hi(a4.x)); w1[1] = pk2(bf16_lo(a4.y), bf16_hi(a4.y));
          w1[2] = pk2(bf16_lo(a4.z), bf16_hi(a4.z)); w1[3] = pk2(bf16_lo(a4.w), bf16_hi(a4.w));
          w2[0] = pk2(bf16_lo(b4.x), bf16_hi(b4.x)); w2[1] = pk2(bf16_lo(b4.y), bf16_hi(b4.y));
          w2[2] = pk2(bf16_lo(b4.z), bf16_hi(b4.z)); w2[3] = pk2(bf16_lo(b4.w), bf16_hi(b4.w));
          const uint4 d4 = crow[xx * 4];
          uint64_t d[4];
          d[0] = pk2(bf16_lo(d4.x), bf16_hi(d4.x)); d[1] = pk2(bf16_lo(d4.y), bf16_hi(d4.y));
          d[2] = pk2(bf16_lo(d4.z), bf16_hi(d4.z)); d[3] = pk2(bf16_lo(d4.w), bf16_hi(d4.w));
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            acc[0][c] = fma2(d[c], w0[c], acc[0][c]);
            acc[1][c] = fma2(d[c], w1[c], acc[1][c]);
            acc[2][c] = fma2(d[c], w2[c], acc[2][c]);
          }
        }
      }
    }
  }
  __syncthreads();
  if (worker) {
#pragma unroll
    for (int k = 0; k < K; ++k)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float a0, a1;
        upk2(acc[k][c], a0, a1);
        atomicAdd(&s_red[((dz * K + dyy) * K + k) * 32 + cc * 8 + 2 * c], (double)a0);
        atomicAdd(&s_red[((dz * K + dyy) * K + k) * 32 + cc * 8 + 2 * c + 1], (double)a1);
      }
  }
  __syncthreads();
  for (int i = tid; i < K * K * K * 32; i += 256) atomicAdd(&dW[(int64_t)(i >> 5) * C + cg * 32 + (i & 31)], s_red[i]);
}

static bool launch_dw_wgrad_s2_tiled(cudaStream_t st, const uint4* center, const uint4* fine, double* dW, const int64_t c_size[3],
                                     const int64_t f_size[3], int C, int N) {
  static const bool off = getenv("PCB_NO_WG2") != nullptr;
  if (off) return false;
  const size_t smem = (size_t)W2_BZ * W2_BY * W2_BX * 64 + (size_t)W2_Z * W2_Y * W2_X * 64 + (size_t)27 * 32 * 8 + 16;
  static DevFlag configured;
  if (!configured) {
    cudaFuncSetAttribute(dw_wgrad_s2_tiled_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (cudaFuncSetAttribute(dw_wgrad_s2_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) {
      cudaGetLastError();
      return false;
    }
    configured = true;
  }
  const int tz = (int)((c_size[0] + W2_Z - 1) / W2_Z), ty = (int)((c_size[1] + W2_Y - 1) / W2_Y), tx = (int)((c_size[2] + W2_X - 1) / W2_X);
  const int64_t nb = (int64_t)tz * ty * tx;
  if (nb * N >= (1ll << 31)) return false;
  const int ncg = C / 32;
  int ctas = (148 * 2) / ncg;                 // two resident CTAs per SM over all channel groups
  if (ctas < 1) ctas = 1;
  if (nb * N < ctas) ctas = (int)(nb * N);
  dw_wgrad_s2_tiled_kernel<<<(unsigned)(ctas * ncg), 256, smem, st>>>(center, fine, dW, (int)c_size[0], (int)c_size[1], (int)c_size[2],
                                                                     (int)f_size[0], (int)f_size[1], (int)f_size[2], C, ty, tx, (int)nb, N);
  return true;
}

// ============================================================================ head / stem backward
// head: out[n,k,v] = sum_c x[v,c] w[c,k] + b[k].   dX[v,c] = sum_k dO[k,v] w[c,k];
//       dW[c,k] += sum_v x[v,c] dO[k,v]; db[k] += sum_v dO[k,v]     (float64 accumulators)
template <typename TO>
__global__ void __launch_bounds__(128) head_bwd_kernel(const TO* __restrict__ dO, const uint4* __restrict__ x,
                                                       const float* __restrict__ w, uint4* __restrict__ dX,
                                                       double* __restrict__ dW, double* __restrict__ db,
                                                       int C, int ncls, int64_t V, int TV) {
  extern __shared__ float sm[];
  float* s_w = sm;                      // [C][ncls]
  float* s_do = s_w + C * ncls;         // [ncls][TV]
  float* s_x = s_do + ncls * TV;        // [TV][C+1]
  const int n = blockIdx.y, tid = threadIdx.x, CH = C >> 3;
  for (int i = tid; i < C * ncls; i += 128) s_w[i] = w[i];
  const int npair = C * ncls;
  // each thread owns pairs (c,k) = tid, tid+128, ... ; accumulate over the CTA's tiles
  float wacc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};   // supports C*ncls <= 1024
  float bacc = 0.f;
  for (int64_t t0 = (int64_t)blockIdx.x * TV; t0 < V; t0 += (int64_t)gridDim.x * TV) {
    __syncthreads();
    const int64_t v = t0 + tid;
    const bool mine = tid < TV && v < V;
    if (tid < TV) {
      for (int k = 0; k < ncls; ++k) s_do[k * TV + tid] = mine ? (float)dO[((int64_t)n * ncls + k) * V + v] : 0.f;
      for (int c8 = 0; c8 < CH; ++c8) {
        float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (mine) unpack8(__ldg(x + ((int64_t)n * V + v) * CH + c8), f);
#pragma unroll
        for (int j = 0; j < 8; ++j) s_x[tid * (C + 1) + c8 * 8 + j] = f[j];
      }
    }
    __syncthreads();
    if (mine) {
      for (int c8 = 0; c8 < CH; ++c8) {
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float s = 0.f;
          for (int k = 0; k < ncls; ++k) s = fmaf(s_do[k * TV + tid], s_w[(c8 * 8 + j) * ncls + k], s);
          o[j] = s;
        }
        dX[((int64_t)n * V + v) * CH + c8] = pack8(o);
      }
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int pr = tid + u * 128;
      if (pr < npair) {
        const int c = pr / ncls, k = pr - c * ncls;
        float s = 0.f;
        for (int vv = 0; vv < TV; ++vv) s = fmaf(s_x[vv * (C + 1) + c], s_do[k * TV + vv], s);
        wacc[u] += s;
      }
    }
    if (tid < ncls) {
      float s = 0.f;
      for (int vv = 0; vv < TV; ++vv) s += s_do[tid * TV + vv];
      bacc += s;
    }
  }
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    const int pr = tid + u * 128;
    if (pr < npair) atomicAdd(&dW[pr], (double)wacc[u]);
  }
  if (tid < ncls) atomicAdd(&db[tid], (double)bacc);
}

// Streaming variant for the full-resolution head (few classes, C/8 a power of two <= 32): thread = (voxel, 8-channel chunk)
// with the chunk fixed per thread, so its weight rows and its dW partials live in registers; x is read and dX written
// exactly once with fully coalesced 128-bit accesses; one warp-shuffle + shared f64 reduction per CTA at the end.
template <typename TO, int NCLS>
__global__ void __launch_bounds__(256) head_bwd_stream_kernel(const TO* __restrict__ dO, const uint4* __restrict__ x,
                                                              const float* __restrict__ w, uint4* __restrict__ dX,
                                                              double* __restrict__ dW, double* __restrict__ db,
                                                              int C, int64_t V) {
  extern __shared__ double s_red[];      // [C*NCLS + NCLS]
  const int CH = C >> 3, n = blockIdx.y, tid = threadIdx.x;
  for (int i = tid; i < C * NCLS + NCLS; i += 256) s_red[i] = 0.0;
  __syncthreads();
  const int64_t T = (int64_t)gridDim.x * 256;            // multiple of CH (host guarantees) -> fixed chunk per thread
  const int64_t first = (int64_t)blockIdx.x * 256 + tid;
  const int cc = (int)(first % CH);
  float wr[8][NCLS], wacc[8][NCLS], bacc[NCLS];
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int k = 0; k < NCLS; ++k) { wr[j][k] = __ldg(w + (cc * 8 + j) * NCLS + k); wacc[j][k] = 0.f; }
#pragma unroll
  for (int k = 0; k < NCLS; ++k) bacc[k] = 0.f;
  const uint4* xn = x + (int64_t)n * V * CH;
  uint4* dxn = dX + (int64_t)n * V * CH;
  const TO* don = dO + (int64_t)n * NCLS * V;
  for (int64_t it = first; it < V * CH; it += T) {
    const int64_t v = it / CH;
    float g[NCLS], f[8], o[8];
#pragma unroll
    for (int k = 0; k < NCLS; ++k) g[k] = (float)don[(int64_t)k * V + v];
    unpack8(ldg_nc(xn + it), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float sacc = 0.f;
#pragma unroll
      for (int k = 0; k < NCLS; ++k) { sacc = fmaf(g[k], wr[j][k], sacc); wacc[j][k] = fmaf(f[j], g[k], wacc[j][k]); }
      o[j] = sacc;
    }
    dxn[it] = pack8(o);
    if (cc == 0) {
#pragma unroll
      for (int k = 0; k < NCLS; ++k) bacc[k] += g[k];
    }
  }
  // lanes with equal (lane % CH) share the chunk: butterfly over the higher lane bits
  for (int off = 16; off >= CH; off >>= 1) {
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
      for (int k = 0; k < NCLS; ++k) wacc[j][k] += __shfl_xor_sync(0xffffffffu, wacc[j][k], off);
#pragma unroll
    for (int k = 0; k < NCLS; ++k) bacc[k] += __shfl_xor_sync(0xffffffffu, bacc[k], off);
  }
  if ((tid & 31) < CH) {
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
      for (int k = 0; k < NCLS; ++k) atomicAdd(&s_red[(cc * 8 + j) * NCLS + k], (double)wacc[j][k]);
    if (cc == 0) {
#pragma unroll
      for (int k = 0; k < NCLS; ++k) atomicAdd(&s_red[C * NCLS + k], (double)bacc[k]);
    }
  }
  __syncthreads();
  for (int i = tid; i < C * NCLS; i += 256) atomicAdd(&dW[i], s_red[i]);
  if (tid < NCLS) atomicAdd(&db[tid], s_red[C * NCLS + tid]);
}

// stem: out[v,c] = sum_ci x[ci,v] w[c,ci] + b[c].  dW[c,ci] += sum_v g[v,c] x[ci,v]; db[c] += sum_v g[v,c]
template <typename TIn>
__global__ void __launch_bounds__(256) stem_bwd_kernel(const uint4* __restrict__ g, const TIn* __restrict__ x,
                                                       double* __restrict__ dW, double* __restrict__ db, int Cin,
                                                       int C, int64_t V) {
  extern __shared__ double s_acc[];   // [C*(Cin+1)]
  const int CH = C >> 3, n = blockIdx.y;
  for (int i = threadIdx.x; i < C * (Cin + 1); i += blockDim.x) s_acc[i] = 0.0;
  __syncthreads();
  const int64_t items = V * CH;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int ci = 0; ci <= Cin; ++ci) {   // ci == Cin: bias column (x = 1)
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    int cc_fixed = -1;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < items; i += stride) {
      const int cc = (int)(i % CH);
      const int64_t v = i / CH;
      if (cc_fixed >= 0 && cc != cc_fixed) {
#pragma unroll
        for (int j = 0; j < 8; ++j) { atomicAdd(&s_acc[(cc_fixed * 8 + j) * (Cin + 1) + ci], (double)acc[j]); acc[j] = 0.f; }
      }
      cc_fixed = cc;
      float gv[8];
      unpack8(__ldg(g + (int64_t)n * items + i), gv);
      const float xv = ci < Cin ? (float)x[((int64_t)n * Cin + ci) * V + v] : 1.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = fmaf(gv[j], xv, acc[j]);
    }
    if (cc_fixed >= 0) {
#pragma unroll
      for (int j = 0; j < 8; ++j) atomicAdd(&s_acc[(cc_fixed * 8 + j) * (Cin + 1) + ci], (double)acc[j]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C * (Cin + 1); i += blockDim.x) {
    const int c = i / (Cin + 1), ci = i - c * (Cin + 1);
    if (ci < Cin) atomicAdd(&dW[c * Cin + ci], s_acc[i]);
    else atomicAdd(&db[c], s_acc[i]);
  }
}

// Streaming variant (C/8 a power of two, Cin <= 4): ONE pass over g with the thread's 8-channel chunk fixed, all Cin+1
// partial columns (weights + bias) in registers; the generic kernel above walks g once per input channel and bias.
template <typename TIn, int CIN>
__global__ void __launch_bounds__(256) stem_bwd_stream_kernel(const uint4* __restrict__ g, const TIn* __restrict__ x,
                                                              double* __restrict__ dW, double* __restrict__ db, int C,
                                                              int64_t V, int ch_shift) {
  extern __shared__ double s_acc[];   // [C*(CIN+1)]
  const int CH = C >> 3, n = blockIdx.y, tid = threadIdx.x;
  for (int i = tid; i < C * (CIN + 1); i += 256) s_acc[i] = 0.0;
  __syncthreads();
  const int64_t items = V * CH;
  const int64_t stride = (int64_t)gridDim.x * 256;          // multiple of CH
  const int64_t first = (int64_t)blockIdx.x * 256 + tid;
  const int cc = (int)(first & (CH - 1));
  float acc[CIN + 1][8];
#pragma unroll
  for (int ci = 0; ci <= CIN; ++ci)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[ci][j] = 0.f;
  const uint4* gn = g + (int64_t)n * items;
  const TIn* xn = x + (int64_t)n * CIN * V;
  for (int64_t i = first; i < items; i += stride) {
    const int64_t v = i >> ch_shift;
    float gv[8];
    unpack8(ldg_nc(gn + i), gv);
#pragma unroll
    for (int ci = 0; ci < CIN; ++ci) {
      const float xv = (float)xn[(int64_t)ci * V + v];
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[ci][j] = fmaf(gv[j], xv, acc[ci][j]);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[CIN][j] += gv[j];
  }
  // lanes with equal (lane % CH) own the same chunk
  for (int off = 16; off >= CH; off >>= 1) {
#pragma unroll
    for (int ci = 0; ci <= CIN; ++ci)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[ci][j] += __shfl_xor_sync(0xffffffffu, acc[ci][j], off);
  }
  if ((tid & 31) < CH || CH > 32) {
#pragma unroll
    for (int ci = 0; ci <= CIN; ++ci)
#pragma unroll
      for (int j = 0; j < 8; ++j) atomicAdd(&s_acc[(cc * 8 + j) * (CIN + 1) + ci], (double)acc[ci][j]);
  }
  __syncthreads();
  for (int i = tid; i < C * (CIN + 1); i += 256) {
    const int c = i / (CIN + 1), ci = i - c * (CIN + 1);
    if (ci < CIN) atomicAdd(&dW[c * CIN + ci], s_acc[i]);
    else atomicAdd(&db[c], s_acc[i]);
  }
}

static inline int pick_chunk_b(int64_t n, int cap) {
  for (int c = cap; c >= 16; c >>= 1)
    if (n % c == 0) return c;
  return 0;
}

}  // namespace pcb

using namespace pcb;

extern "C" int pcb_mlp_bwd(const void* y, const double* stats, const float* gamma, const float* beta, const void* w2,
                           const float* b2, const void* w3t, const void* w2t, const void* dout, void* hact, void* dh,
                           void* dyhat, double* gstats, int64_t N, const int64_t y_size[3], int64_t C, int64_t H,
                           int64_t Co, int mode, void* stream) {
  PCB_CHECK_ARG(y && stats && gamma && beta && w2 && b2 && w3t && w2t && dout && hact && dh && dyhat && gstats && y_size,
                "pcb_mlp_bwd: null argument");
  PCB_CHECK_ARG(C % 16 == 0 && H % 16 == 0 && Co % 16 == 0, "pcb_mlp_bwd: channel counts must be multiples of 16");
  PCB_CHECK_ARG(N > 0 && N <= 65535, "pcb_mlp_bwd: bad batch");
  PCB_CHECK_ARG(mode >= PCB_DW_SAME && mode <= PCB_DW_UP, "pcb_mlp_bwd: bad mode %d", mode);
  {
    const int r = mlp_bwd_deep(y, stats, gamma, beta, w2, b2, w3t, w2t, dout, hact, dh, dyhat, gstats, N, y_size, C, H, Co, mode, stream);
    if (r != 0) return r > 0 ? PCB_OK : r;
  }
  MlpBwdArgs a;
  a.y = (const uint4*)y; a.stats = stats; a.gamma = gamma; a.beta = beta; a.w2 = (const uint4*)w2; a.b2 = b2;
  a.w3t = (const uint4*)w3t; a.w2t = (const uint4*)w2t; a.dout = (const uint4*)dout; a.hact = (uint4*)hact;
  a.dh = (uint4*)dh; a.dyhat = (uint4*)dyhat; a.gstats = gstats;
  a.y1 = (int)y_size[1]; a.y2 = (int)y_size[2];
  a.C = (int)C; a.H = (int)H; a.Co = (int)Co; a.mode = mode;
  a.Vy = y_size[0] * y_size[1] * y_size[2];
  if (mode == PCB_DW_UP) { a.o1 = a.y1 + 1; a.o2 = a.y2 + 1; a.Vout = (y_size[0] + 1) * (int64_t)a.o1 * a.o2; }
  else { a.o1 = a.y1; a.o2 = a.y2; a.Vout = a.Vy; }
  a.KC = pick_chunk_b(C, 128); a.KCo = pick_chunk_b(Co, 128); a.N1 = pick_chunk_b(H, 64);
  a.Ct = C <= 256 ? (int)C : 256;
  PCB_CHECK_ARG(C % a.Ct == 0, "pcb_mlp_bwd: C=%lld must be <=256 or a multiple of 256", (long long)C);
  a.inv_count = (float)(1.0 / (double)a.Vy);
  const size_t smem = (size_t)128 * a.KC * 2 + (size_t)128 * a.KCo * 2 + (size_t)a.N1 * a.KC * 2 + (size_t)a.N1 * a.KCo * 2 +
                      (size_t)128 * a.N1 * 2 + (size_t)a.Ct * a.N1 * 2 + (size_t)4 * C * sizeof(float) +
                      (size_t)2 * a.Ct * sizeof(double) + 128 * sizeof(int64_t) + 16;
  PCB_CHECK_ARG(smem <= 227 * 1024, "pcb_mlp_bwd: tile needs %zu B shared memory", smem);
  static DevFlag configured;
  if (!configured) {
    cudaFuncSetAttribute(mlp_bwd_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (cudaFuncSetAttribute(mlp_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) {
      set_error("pcb_mlp_bwd: cudaFuncSetAttribute failed"); return PCB_ERR_CUDA;
    }
    configured = true;
  }
  dim3 grid((unsigned)((a.Vy + 127) / 128), (unsigned)N, (unsigned)(C / a.Ct));
  mlp_bwd_kernel<<<grid, 128, smem, (cudaStream_t)stream>>>(a);
  PCB_CHECK_LAUNCH("pcb_mlp_bwd");
  return PCB_OK;
}


static bool mlp_bwd_fused_ok(int64_t C, int64_t H, int64_t Co, int64_t N, int64_t Vy, int64_t Vout) {
  auto pow2 = [](int64_t v) { return v > 0 && (v & (v - 1)) == 0; };
  // a hidden width that is not a power of two (MedNeXt-L level 0: H = 96) is served by mlp_bwd_ws2_kernel only
  const bool ws2_shape = (C == 32 || C == 64) && (Co == 32 || Co == 64) && N <= 8;
  return getenv("PCB_NO_FUSED_BWD") == nullptr && pow2(C) && pow2(Co) && (pow2(H) || ws2_shape) && C <= 64 && Co <= 128 && H < 128 + 1 && H % 16 == 0 &&
         3 * H + C + Co <= 512 && (C + Co) / 8 * 128 <= 8 * 256 && N <= 8 && Vy < (1ll << 30) && Vout < (1ll << 30) &&
         mlp_bwd_fused_smem((int)C, (int)H, (int)Co, (int)N) <= 227 * 1024;
}

extern "C" int pcb_mlp_bwd_fused_supported(int64_t C, int64_t H, int64_t Co, int64_t N, const int64_t y_size[3], int mode) {
  const int64_t Vy = y_size[0] * y_size[1] * y_size[2];
  const int64_t Vout = mode == PCB_DW_UP ? (y_size[0] + 1) * (y_size[1] + 1) * (y_size[2] + 1) : Vy;
  return mlp_bwd_fused_ok(C, H, Co, N, Vy, Vout) ? 1 : 0;
}

static int mlp_bwd_fused_ctas(int64_t ntiles, uint32_t tmem_cols) {
  int64_t p = 148 * (tmem_cols <= 256 ? 2 : 1);
  if (p > ntiles) p = ntiles;
  return (int)p;
}

extern "C" int64_t pcb_mlp_bwd_fused_workspace_floats(int64_t C, int64_t H, int64_t Co, int64_t N, const int64_t y_size[3]) {
  const int64_t ntiles = (y_size[0] * y_size[1] * y_size[2] + 127) / 128 * N;
  const int P = mlp_bwd_fused_ctas(ntiles, tmem_cols_pow2((uint32_t)(3 * H + C + Co)));
  return (int64_t)P * (129 * Co + 128 * H);
}

extern "C" int pcb_mlp_bwd_fused(const void* y, const double* stats, const float* gamma, const float* beta, const void* w2,
                                 const float* b2, const void* w3t, const void* w2t, const void* dout, void* dyhat,
                                 double* gstats, float* workspace, float* dW3, float* db3, float* dW2, float* db2, int64_t N,
                                 const int64_t y_size[3], int64_t C, int64_t H, int64_t Co, int mode, void* stream) {
  PCB_CHECK_ARG(y && stats && gamma && beta && w2 && b2 && w3t && w2t && dout && dyhat && gstats && workspace && dW3 && db3 &&
                dW2 && db2 && y_size, "pcb_mlp_bwd_fused: null argument");
  MlpBwdFusedArgs fa;
  memset(&fa, 0, sizeof(fa));
  MlpBwdArgs& a = fa.m;
  a.y = (const uint4*)y; a.stats = stats; a.gamma = gamma; a.beta = beta; a.w2 = (const uint4*)w2; a.b2 = b2;
  a.w3t = (const uint4*)w3t; a.w2t = (const uint4*)w2t; a.dout = (const uint4*)dout; a.hact = nullptr; a.dh = nullptr;
  a.dyhat = (uint4*)dyhat; a.gstats = gstats;
  a.y1 = (int)y_size[1]; a.y2 = (int)y_size[2];
  a.C = (int)C; a.H = (int)H; a.Co = (int)Co; a.mode = mode;
  a.Vy = y_size[0] * y_size[1] * y_size[2];
  if (mode == PCB_DW_UP) { a.o1 = a.y1 + 1; a.o2 = a.y2 + 1; a.Vout = (y_size[0] + 1) * (int64_t)a.o1 * a.o2; }
  else { a.o1 = a.y1; a.o2 = a.y2; a.Vout = a.Vy; }
  PCB_CHECK_ARG(mlp_bwd_fused_ok(C, H, Co, N, a.Vy, a.Vout), "pcb_mlp_bwd_fused: unsupported shape C=%lld H=%lld Co=%lld", (long long)C, (long long)H, (long long)Co);
  a.KC = (int)C; a.KCo = (int)Co; a.N1 = (int)H; a.Ct = (int)C;
  a.inv_count = (float)(1.0 / (double)a.Vy);
  fa.N = (int)N; fa.tps = (a.Vy + 127) / 128; fa.ntiles = fa.tps * N;
  const uint32_t tcols = tmem_cols_pow2((uint32_t)(3 * H + C + Co));
  const int P = mlp_bwd_fused_ctas(fa.ntiles, tcols);
  fa.part3 = workspace; fa.part2 = workspace + (int64_t)P * 129 * Co;
  const size_t smem = mlp_bwd_fused_smem((int)C, (int)H, (int)Co, (int)N);
  static DevFlag configured;
  if (!configured) {
    cudaFuncSetAttribute(mlp_bwd_fused_kernel<256>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(mlp_bwd_fused_kernel<512>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (cudaFuncSetAttribute(mlp_bwd_fused_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess ||
        cudaFuncSetAttribute(mlp_bwd_fused_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) {
      set_error("pcb_mlp_bwd_fused: cudaFuncSetAttribute failed"); return PCB_ERR_CUDA;
    }
    configured = true;
  }
  cudaStream_t st = (cudaStream_t)stream;
  // generalised warp-specialised variant: C in {32, 64}, Co in {32, 64}, every mode
  {
    // default policy (measured round 2, 2 x 160^3: level-0 SAME 1.226 ms ws vs 1.334 ws2 vs 1.536 fused; down_0 0.231 ws2 vs
    // 0.367; up_0 3.14 ws2 vs 3.60; level 1 0.506 ws2 vs 0.518): the level-0 SAME shape runs mlp_bwd_ws_kernel, every other
    // fused-backward shape mlp_bwd_ws2_kernel.  PCB_BWD_WS=0 forces the non-specialised kernel, =1 / =2 one variant.
    const char* ws_env = getenv("PCB_BWD_WS");
    const bool l0_same = C == 32 && Co == 32 && H == 64 && mode != PCB_DW_UP;
    const bool want_ws2 = ws_env ? ws_env[0] == '2' : !l0_same;
    if (want_ws2 && (C == 32 || C == 64) && (Co == 32 || Co == 64) && H % 16 == 0 && H <= 128 && N <= 8) {
      const int NB = (2 * (2 * H + C) + Co + H <= 512) ? 2 : 1;
      // operand stages: 4 with double-buffered accumulators; with one accumulator set as many of 4 / 3 / 2 as shared memory
      // allows (round 2: the NB = 1 shapes were bound by their loaders — two stages keep two of the four loader warps busy and
      // 8 KB in flight per SM; splitting E1 / E2 over both epilogue groups alone changed nothing)
      int nst = 4;
      { const char* e = getenv("PCB_BWD_NST"); if (e && e[0] >= '2' && e[0] <= '4') nst = e[0] - '0'; }
      if (NB == 2) nst = 4;
      while (nst > 2 && mlp_bwd_ws2_smem((int)C, (int)H, (int)Co, (int)N, NB, nst, mode == PCB_DW_UP) > 227 * 1024) --nst;
      const size_t smem_ws = mlp_bwd_ws2_smem((int)C, (int)H, (int)Co, (int)N, NB, nst, mode == PCB_DW_UP);
      if ((int64_t)NB * (2 * H + C) + Co + H <= 512 && smem_ws <= 227 * 1024) {
        auto conf = [&](const void* fn) {
          cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
          if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) { cudaGetLastError(); return false; }
          return true;
        };
        const int Pw = (int)(fa.ntiles < 148 ? fa.ntiles : 148);     // <= P: the caller's workspace is large enough
        fa.part3 = workspace; fa.part2 = workspace + (int64_t)Pw * 129 * Co;
        ws2_magic((uint32_t)a.y2, fa.dm2, fa.ds2);
        ws2_magic((uint32_t)a.y1, fa.dm1, fa.ds1);
        { const char* e16 = getenv("PCB_BWD_LD16"); fa.ld16 = (e16 && e16[0] == '1') ? 1 : 0; }
        // column split of E1 / E2 over both groups: opt-in — with 4 operand stages the alternating scheme is faster
        // (up_0 2.34 vs 2.82 ms, level 1 0.385 vs 0.432 ms at 2 x 160^3)
        { const char* sp = getenv("PCB_BWD_SPLIT"); fa.split = (sp && sp[0] == '1') ? 1 : 0; }
        bool launched = false;
#define PCB_WS2(C8, NBV, NSV)                                                                                      \
        if (!launched && C == 8 * C8 && NB == NBV && nst == NSV && conf((const void*)mlp_bwd_ws2_kernel<C8, NBV, NSV>)) {   \
          mlp_bwd_ws2_kernel<C8, NBV, NSV><<<Pw, WS_THREADS, smem_ws, st>>>(fa); launched = true;                   \
        }
        PCB_WS2(4, 2, 4) PCB_WS2(4, 1, 4) PCB_WS2(4, 1, 3) PCB_WS2(4, 1, 2) PCB_WS2(8, 1, 4) PCB_WS2(8, 1, 3) PCB_WS2(8, 1, 2)
#undef PCB_WS2
        if (launched) {
          PCB_CHECK_LAUNCH("pcb_mlp_bwd_fused(ws2)");
          reduce_partials_kernel<<<(unsigned)((H * Co + 31) / 32), 256, 0, st>>>(fa.part3, Pw, (int)H, 129, (int)Co, (int)Co, dW3, 1, H, nullptr);
          reduce_partials_kernel<<<(unsigned)((Co + 31) / 32), 256, 0, st>>>(fa.part3 + 128 * Co, Pw, 1, 129, (int)Co, (int)Co, db3, 0, 1, nullptr);
          reduce_partials_kernel<<<(unsigned)((C * H + 31) / 32), 256, 0, st>>>(fa.part2, Pw, (int)C, 128, (int)H, (int)H, dW2, 1, C, nullptr);
          reduce_partials_kernel<<<(unsigned)((H + 31) / 32), 256, 0, st>>>(fa.part2 + C * H, Pw, 1, 128, (int)H, (int)H, db2, 0, 1, nullptr);
          PCB_CHECK_LAUNCH("pcb_mlp_bwd_fused(ws2 reduce)");
          return PCB_OK;
        }
        fa.part3 = workspace; fa.part2 = workspace + (int64_t)P * 129 * Co;   // not launched: back to the default layout
      }
    }
  }
  PCB_CHECK_ARG((H & (H - 1)) == 0, "pcb_mlp_bwd_fused: H=%lld needs the generalised warp-specialised kernel (PCB_BWD_WS must not be 0 / 1)", (long long)H);
  // warp-specialised variant for the level-0 shape (see mlp_bwd_ws_kernel); same workspace layout, P = grid size
  {
    const char* ws_env = getenv("PCB_BWD_WS");
    if ((ws_env ? ws_env[0] == '1' : true) && C == 32 && Co == 32 && H == 64 && mode != PCB_DW_UP && N <= 8 &&
        mlp_bwd_ws_smem((int)C, (int)H, (int)Co, (int)N) <= 227 * 1024) {
      static DevFlag conf_ws;
      if (!conf_ws) {
        cudaFuncSetAttribute(mlp_bwd_ws_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        conf_ws = cudaFuncSetAttribute(mlp_bwd_ws_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) == cudaSuccess;
        if (!conf_ws) cudaGetLastError();
      }
      if (conf_ws) {
        const int Pw = (int)(fa.ntiles < 148 ? fa.ntiles : 148);     // <= P: the caller's workspace is large enough
        fa.part3 = workspace; fa.part2 = workspace + (int64_t)Pw * 129 * Co;
        mlp_bwd_ws_kernel<<<Pw, WS0_THREADS, mlp_bwd_ws_smem((int)C, (int)H, (int)Co, (int)N), st>>>(fa);
        PCB_CHECK_LAUNCH("pcb_mlp_bwd_fused(ws)");
        reduce_partials_kernel<<<(unsigned)((H * Co + 31) / 32), 256, 0, st>>>(fa.part3, Pw, (int)H, 129, (int)Co, (int)Co, dW3, 1, H, nullptr);
        reduce_partials_kernel<<<(unsigned)((Co + 31) / 32), 256, 0, st>>>(fa.part3 + 128 * Co, Pw, 1, 129, (int)Co, (int)Co, db3, 0, 1, nullptr);
        reduce_partials_kernel<<<(unsigned)((C * H + 31) / 32), 256, 0, st>>>(fa.part2, Pw, (int)C, 128, (int)H, (int)H, dW2, 1, C, nullptr);
        reduce_partials_kernel<<<(unsigned)((H + 31) / 32), 256, 0, st>>>(fa.part2 + C * H, Pw, 1, 128, (int)H, (int)H, db2, 0, 1, nullptr);
        PCB_CHECK_LAUNCH("pcb_mlp_bwd_fused(ws reduce)");
        return PCB_OK;
      }
    }
  }
  // one CTA per SM (TMEM > 256 columns): four warpgroups; H >= 64 and C >= 64 keep every warpgroup busy in E1 / E2
  const char* nt_env = getenv("PCB_BWD_NT512");
  const bool wide = tcols > 256 && H / 16 >= 4 && !(nt_env && nt_env[0] == '0');
  if (wide) mlp_bwd_fused_kernel<512><<<P, 512, smem, st>>>(fa);
  else mlp_bwd_fused_kernel<256><<<P, 256, smem, st>>>(fa);
  PCB_CHECK_LAUNCH("pcb_mlp_bwd_fused");
  // second stage: dW3[co,h] = sum_p part3[p][h][co] ; db3[co] = sum_p part3[p][H][co]   (and the same for W2)
  reduce_partials_kernel<<<(unsigned)((H * Co + 31) / 32), 256, 0, st>>>(fa.part3, P, (int)H, 129, (int)Co, (int)Co, dW3, 1, H, nullptr);
  reduce_partials_kernel<<<(unsigned)((Co + 31) / 32), 256, 0, st>>>(fa.part3 + 128 * Co, P, 1, 129, (int)Co, (int)Co, db3, 0, 1, nullptr);
  reduce_partials_kernel<<<(unsigned)((C * H + 31) / 32), 256, 0, st>>>(fa.part2, P, (int)C, 128, (int)H, (int)H, dW2, 1, C, nullptr);
  reduce_partials_kernel<<<(unsigned)((H + 31) / 32), 256, 0, st>>>(fa.part2 + C * H, P, 1, 128, (int)H, (int)H, db2, 0, 1, nullptr);
  PCB_CHECK_LAUNCH("pcb_mlp_bwd_fused(reduce)");
  return PCB_OK;
}

static inline bool tn_use_ws() { const char* e = getenv("PCB_TN_WS"); return !(e && e[0] == '0'); }

static inline int tn_num_ctas(int64_t ntiles, int64_t Mtot, int64_t Nc) {
  int64_t p = tn_use_ws() ? 148 : 148 * 3;              // persistent: one CTA per SM (the single-stage kernel: three)
  const int64_t cap = (int64_t)(96 << 20) / (Mtot * Nc * 4);   // keep the partial-sum workspace <= 96 MB
  if (p > cap) p = cap < 1 ? 1 : cap;
  if (p > ntiles) p = ntiles;
  return (int)p;
}

extern "C" int64_t pcb_tn_workspace_floats(int64_t Ma, int64_t Nb, int ones, int64_t N, const int64_t box[3]) {
  const int64_t ntiles = (box[0] * box[1] * box[2] + 127) / 128 * N;
  const int64_t Mtot = (Ma + 127) / 128 * 128, Nc = Nb + (ones ? 16 : 0);
  return (int64_t)tn_num_ctas(ntiles, Mtot, Nc) * Mtot * Nc;
}

static int tn_gemm_impl(const void* A, const void* B, const double* stats, const float* gamma, const float* beta,
                        float* workspace, float* dW, int64_t ldm, int64_t ldn, float* db, int64_t N,
                        const int64_t box[3], int mapA, const int64_t a_size[3], int64_t a_cols, int64_t Ma,
                        int mapB, const int64_t b_size[3], int64_t Nb, int ones, const int* tap, void* stream);

extern "C" int pcb_tn_gemm(const void* A, const void* B, const double* stats, const float* gamma, const float* beta,
                           float* workspace, float* dW, int64_t ldm, int64_t ldn, float* db, int64_t N,
                           const int64_t box[3], int mapA, const int64_t a_size[3], int64_t a_cols, int64_t Ma,
                           int mapB, const int64_t b_size[3], int64_t Nb, int ones, void* stream) {
  PCB_CHECK_ARG(mapB >= 0 && mapB <= 3, "pcb_tn_gemm: bad mapB");
  return tn_gemm_impl(A, B, stats, gamma, beta, workspace, dW, ldm, ldn, db, N, box, mapA, a_size, a_cols, Ma, mapB, b_size, Nb,
                      ones, nullptr, stream);
}

extern "C" int pcb_conv_wgrad_tap(const void* dy, const void* x, float* workspace, float* dW, int64_t ldm, int64_t ldn, float* db,
                                  int64_t N, const int64_t out_size[3], const int64_t in_size[3], int64_t Co, int64_t Ci,
                                  const int tap[3], int stride, int pad, int transposed, void* stream) {
  PCB_CHECK_ARG(tap, "pcb_conv_wgrad_tap: null tap");
  const int t[5] = {tap[0], tap[1], tap[2], stride, pad};
  return tn_gemm_impl(dy, x, nullptr, nullptr, nullptr, workspace, dW, ldm, ldn, db, N, out_size, MAP_IDENT, out_size, Co, Co,
                      transposed ? 5 : 4, in_size, Ci, db ? 1 : 0, t, stream);
}

static int tn_gemm_impl(const void* A, const void* B, const double* stats, const float* gamma, const float* beta,
                        float* workspace, float* dW, int64_t ldm, int64_t ldn, float* db, int64_t N,
                        const int64_t box[3], int mapA, const int64_t a_size[3], int64_t a_cols, int64_t Ma,
                        int mapB, const int64_t b_size[3], int64_t Nb, int ones, const int* tap, void* stream) {
  PCB_CHECK_ARG(A && B && workspace && dW && box && a_size && b_size, "pcb_tn_gemm: null argument");
  PCB_CHECK_ARG(Ma % 16 == 0 && Nb % 16 == 0 && Ma > 0 && Nb > 0 && a_cols >= Ma && a_cols % 8 == 0,
                "pcb_tn_gemm: Ma/Nb must be positive multiples of 16");
  PCB_CHECK_ARG((db == nullptr) == (ones == 0), "pcb_tn_gemm: db and ones must come together");
  TnArgs a;
  a.A = (const uint4*)A; a.B = (const uint4*)B; a.part = workspace; a.stats = stats; a.gamma = gamma; a.beta = beta;
  a.a_pitch8 = a_cols / 8; a.b_pitch8 = Nb / 8;
  const int64_t Va = a_size[0] * a_size[1] * a_size[2], Vb = b_size[0] * b_size[1] * b_size[2];
  a.a_sample8 = Va * a.a_pitch8; a.b_sample8 = Vb * a.b_pitch8;
  a.Ma = (int)Ma; a.Nb = (int)Nb; a.ones = ones; a.mapA = mapA; a.mapB = mapB;
  a.d1 = (int)box[1]; a.d2 = (int)box[2];
  a.as1 = (int)a_size[1]; a.as2 = (int)a_size[2]; a.bs1 = (int)b_size[1]; a.bs2 = (int)b_size[2];
  a.bs0 = (int)b_size[0];
  a.tz = a.ty = a.tx = 0; a.tstride = 1; a.tpad = 0;
  if (tap) { a.tz = tap[0]; a.ty = tap[1]; a.tx = tap[2]; a.tstride = tap[3]; a.tpad = tap[4]; }
  a.V = box[0] * box[1] * box[2]; a.N = (int)N;
  a.Mtot = (int)((Ma + 127) / 128 * 128); a.Ncols_tot = (int)(Nb + (ones ? 16 : 0));
  a.inv_count = (float)(1.0 / (double)Vb);
  const int64_t ntiles = (a.V + 127) / 128 * N;
  const int P = tn_num_ctas(ntiles, a.Mtot, a.Ncols_tot);
  const int mt = a.Mtot / 128, ncn = (int)((Nb + TN_NCHUNK - 1) / TN_NCHUNK);
  const size_t smem = (size_t)16 * (2048 + 64) + (size_t)((TN_NCHUNK >> 3) + 2) * (2048 + 64) +
                      2 * TN_NCHUNK * sizeof(float) + 64;
  static DevFlag configured;
  if (!configured) {
    cudaFuncSetAttribute(tn_gemm_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (cudaFuncSetAttribute(tn_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) {
      set_error("pcb_tn_gemm: cudaFuncSetAttribute failed"); return PCB_ERR_CUDA;
    }
    configured = true;
  }
  cudaStream_t st = (cudaStream_t)stream;
  bool launched = false;
  if (tn_use_ws() && a.V < (1ll << 30) && a.a_sample8 < (1ll << 31) && a.b_sample8 < (1ll << 31)) {
    TnWsArgs w;
    w.t = a;
    const int ncols_max = (int)((Nb < TN_NCHUNK ? Nb : TN_NCHUNK) + (ones ? 16 : 0));
    w.MT = (mt >= 2 && 2 * ncols_max <= 512 && ntiles >= 2 * P) ? 2 : 1;      // share the B tile between two row tiles of dW
    const size_t stage = (size_t)(w.MT * 16 + (TN_NCHUNK >> 3) + 2) * (2048 + 16);
    const size_t fixed = (size_t)(stats ? 2 * N * TN_NCHUNK * sizeof(float) : 0) + 9 * 8 + 16 + 128;
    int ns = 4;
    while (ns > 1 && (size_t)ns * (stage + 2 * 128 * sizeof(int)) + fixed > 227 * 1024) --ns;
    w.NS = ns;
    ws2_magic((uint32_t)(a.d2 > 0 ? a.d2 : 1), w.dm2, w.ds2);
    ws2_magic((uint32_t)(a.d1 > 0 ? a.d1 : 1), w.dm1, w.ds1);
    const size_t smem_ws = (size_t)ns * (stage + 2 * 128 * sizeof(int)) + fixed;
    static DevFlag conf_ws;
    if (!conf_ws) {
      cudaFuncSetAttribute(tn_gemm_ws_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
      conf_ws = cudaFuncSetAttribute(tn_gemm_ws_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) == cudaSuccess;
      if (!conf_ws) cudaGetLastError();
    }
    if (conf_ws && ns >= 2 && smem_ws <= 227 * 1024) {   // the hand-off trails by one tile: at least two stages
      dim3 grid((unsigned)P, (unsigned)((mt + w.MT - 1) / w.MT), (unsigned)ncn);
      tn_gemm_ws_kernel<<<grid, TNW_THREADS, smem_ws, st>>>(w);
      PCB_CHECK_LAUNCH("pcb_tn_gemm(ws)");
      launched = true;
    }
  }
  if (!launched) {
    dim3 grid((unsigned)P, (unsigned)mt, (unsigned)ncn);
    tn_gemm_kernel<<<grid, 128, smem, st>>>(a);
    PCB_CHECK_LAUNCH("pcb_tn_gemm");
  }
  const int64_t total = Ma * (Nb + (db ? 1 : 0));
  reduce_partials_kernel<<<(unsigned)((total + 31) / 32), 256, 0, st>>>(workspace, P, (int)Ma, a.Mtot, a.Ncols_tot,
                                                                         (int)Nb, dW, ldm, ldn, db);
  PCB_CHECK_LAUNCH("pcb_tn_gemm(reduce)");
  return PCB_OK;
}

extern "C" int pcb_pw_fwd(const void* A, const void* W, const float* bias, void* out, int64_t N, const int64_t out_box[3],
                          int map, const int64_t a_size[3], int64_t K, int64_t Nw, void* stream) {
  PCB_CHECK_ARG(A && W && out && out_box && a_size, "pcb_pw_fwd: null argument");
  PCB_CHECK_ARG(K % 16 == 0 && Nw % 16 == 0 && K > 0 && Nw > 0, "pcb_pw_fwd: K and N must be multiples of 16");
  PwArgs a;
  a.A = (const uint4*)A; a.W = (const uint4*)W; a.bias = bias; a.out = (uint4*)out;
  a.K = (int)K; a.Nw = (int)Nw; a.KC = pick_chunk_b(K, 128); a.NT = Nw <= 256 ? (int)Nw : 256;
  PCB_CHECK_ARG(Nw % a.NT == 0, "pcb_pw_fwd: N must be <=256 or a multiple of 256");
  a.map = map; a.d1 = (int)out_box[1]; a.d2 = (int)out_box[2]; a.s1 = (int)a_size[1]; a.s2 = (int)a_size[2];
  a.Vout = out_box[0] * out_box[1] * out_box[2]; a.Vin = a_size[0] * a_size[1] * a_size[2];
  const size_t smem = (size_t)128 * a.KC * 2 + (size_t)a.NT * a.KC * 2 + 32;
  static DevFlag configured;
  if (!configured) {
    if (cudaFuncSetAttribute(pw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) {
      set_error("pcb_pw_fwd: cudaFuncSetAttribute failed"); return PCB_ERR_CUDA;
    }
    configured = true;
  }
  dim3 grid((unsigned)((a.Vout + 127) / 128), (unsigned)N, (unsigned)(Nw / a.NT));
  pw_kernel<<<grid, 128, smem, (cudaStream_t)stream>>>(a);
  PCB_CHECK_LAUNCH("pcb_pw_fwd");
  return PCB_OK;
}

static int gn_bwd_impl(const void* g, const void* y, const double* stats, const double* gstats, const float* gamma,
                       void* dy, double* dsum, int64_t N, int64_t C, int64_t V, double count, void* stream);

extern "C" int pcb_gn_bwd(const void* g, const void* y, const double* stats, const double* gstats, const float* gamma,
                          void* dy, double* dsum, int64_t N, int64_t C, int64_t V, void* stream) {
  return gn_bwd_impl(g, y, stats, gstats, gamma, dy, dsum, N, C, V, (double)V, stream);
}

/* BatchNorm (batch statistics) backward: same formula with statistics pooled over the batch; `stats` / `gstats`
 * hold the pooled sums replicated for every sample ([N,2,C]) and the element count is N*V. */
extern "C" int pcb_bn_bwd(const void* g, const void* x, const double* stats, const double* gstats, const float* gamma,
                          void* dx, double* dsum, int64_t N, int64_t C, int64_t V, void* stream) {
  return gn_bwd_impl(g, x, stats, gstats, gamma, dx, dsum, N, C, V, (double)(N * V), stream);
}

static int gn_bwd_impl(const void* g, const void* y, const double* stats, const double* gstats, const float* gamma,
                       void* dy, double* dsum, int64_t N, int64_t C, int64_t V, double count, void* stream) {
  PCB_CHECK_ARG(g && y && stats && gstats && gamma && dy && dsum, "pcb_gn_bwd: null argument");
  PCB_CHECK_ARG(C % 8 == 0 && C > 0 && V > 0 && N > 0 && N <= 65535, "pcb_gn_bwd: bad shape");
  const int64_t items = V * (C / 8);
  int64_t blocks = (items + 256 * 4 - 1) / (256 * 4);      // >= 4 chunks per thread amortise the per-CTA coefficient setup
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  dim3 grid((unsigned)blocks, (unsigned)N);
  gn_dy_kernel<<<grid, 256, C * sizeof(double) + 3 * C * sizeof(float), (cudaStream_t)stream>>>((const uint4*)g, (const uint4*)y, stats, gstats, gamma,
                                                                     (uint4*)dy, dsum, (int)C, V, (float)(1.0 / count));
  PCB_CHECK_LAUNCH("pcb_gn_bwd");
  return PCB_OK;
}

extern "C" int pcb_dwconv_wgrad(const void* center, const void* neigh, double* dW, int64_t N, const int64_t c_size[3],
                                const int64_t n_size[3], int64_t C, int k, int stride, void* stream) {
  PCB_CHECK_ARG(center && neigh && dW && c_size && n_size, "pcb_dwconv_wgrad: null argument");
  PCB_CHECK_ARG(k == 3 || k == 5 || k == 7, "MedNeXt kernel_size must be 3, 5, or 7. Got: %d", k);
  PCB_CHECK_ARG(C % 8 == 0 && C > 0 && (stride == 1 || stride == 2) && N > 0 && N <= 65535, "pcb_dwconv_wgrad: bad shape");
  cudaStream_t st0 = (cudaStream_t)stream;
  if (stride == 1 && C % 32 == 0 && c_size[0] == n_size[0] && c_size[1] == n_size[1] && c_size[2] == n_size[2]) {
    bool ok = false;
    if (k == 3) ok = launch_dw_wgrad_tiled<3>(st0, (const uint4*)center, (const uint4*)neigh, dW, (int)c_size[0], (int)c_size[1], (int)c_size[2], (int)C, (int)N);
    else if (k == 5) ok = launch_dw_wgrad_tiled<5>(st0, (const uint4*)center, (const uint4*)neigh, dW, (int)c_size[0], (int)c_size[1], (int)c_size[2], (int)C, (int)N);
    else ok = launch_dw_wgrad_tiled<7>(st0, (const uint4*)center, (const uint4*)neigh, dW, (int)c_size[0], (int)c_size[1], (int)c_size[2], (int)C, (int)N);
    if (ok) { PCB_CHECK_LAUNCH("pcb_dwconv_wgrad(tiled)"); return PCB_OK; }
  }
  if (stride == 2 && k == 3 && C % 32 == 0 &&
      launch_dw_wgrad_s2_tiled(st0, (const uint4*)center, (const uint4*)neigh, dW, c_size, n_size, (int)C, (int)N)) {
    PCB_CHECK_LAUNCH("pcb_dwconv_wgrad(stride-2 tiled)");
    return PCB_OK;
  }
  DwWgArgs a{(int)c_size[0], (int)c_size[1], (int)c_size[2], (int)n_size[0], (int)n_size[1], (int)n_size[2], (int)C, stride};
  const int64_t items = (int64_t)a.c0 * a.c1 * ((a.c2 + DW_XB_WG - 1) / DW_XB_WG) * (C / 8);
  int blocks = (int)((items + 255) / 256);
  const int cap = 148 * 2;
  if (blocks > cap) blocks = cap;
  dim3 grid((unsigned)(blocks * k * k), 1u, (unsigned)N);
  const size_t smem = (size_t)k * C * sizeof(double);
  cudaStream_t st = (cudaStream_t)stream;
  if (k == 3) dw_wgrad_kernel<3><<<grid, 256, smem, st>>>((const uint4*)center, (const uint4*)neigh, dW, a);
  else if (k == 5) dw_wgrad_kernel<5><<<grid, 256, smem, st>>>((const uint4*)center, (const uint4*)neigh, dW, a);
  else dw_wgrad_kernel<7><<<grid, 256, smem, st>>>((const uint4*)center, (const uint4*)neigh, dW, a);
  PCB_CHECK_LAUNCH("pcb_dwconv_wgrad");
  return PCB_OK;
}

extern "C" int pcb_head_bwd(const void* dout, int dtype, const void* x, const float* w, void* dx, double* dW, double* db,
                            int64_t N, int64_t C, int64_t ncls, int64_t nvox, void* stream) {
  PCB_CHECK_ARG(dout && x && w && dx && dW && db, "pcb_head_bwd: null argument");
  PCB_CHECK_ARG(C % 8 == 0 && C > 0 && ncls > 0 && C * ncls <= 1024 && ncls <= 128, "pcb_head_bwd: C*ncls must be <= 1024");
  cudaStream_t st = (cudaStream_t)stream;
  {
    // streaming kernel: few classes, C/8 a power of two that divides the warp; the grid keeps each thread's chunk fixed
    const int ch = (int)(C >> 3);
    static const bool no_stream = getenv("PCB_NO_HEAD_STREAM") != nullptr;
    if (!no_stream && ncls <= 4 && ch <= 32 && (ch & (ch - 1)) == 0 && nvox * ch >= 256) {
      int64_t nb = (nvox * ch + 255) / 256;
      if (nb > 148 * 8) nb = 148 * 8;
      dim3 grid_s((unsigned)nb, (unsigned)N);
      const size_t smem_s = (size_t)(C * ncls + ncls) * sizeof(double);
#define PCB_HEAD_STREAM(T, K) head_bwd_stream_kernel<T, K><<<grid_s, 256, smem_s, st>>>((const T*)dout, (const uint4*)x, w, (uint4*)dx, dW, db, (int)C, nvox)
#define PCB_HEAD_STREAM_T(T)                                                             \
      do {                                                                               \
        if (ncls == 1) PCB_HEAD_STREAM(T, 1); else if (ncls == 2) PCB_HEAD_STREAM(T, 2); \
        else if (ncls == 3) PCB_HEAD_STREAM(T, 3); else PCB_HEAD_STREAM(T, 4);           \
      } while (0)
      if (dtype == PCB_F32) PCB_HEAD_STREAM_T(float);
      else if (dtype == PCB_F16) PCB_HEAD_STREAM_T(__half);
      else if (dtype == PCB_BF16) PCB_HEAD_STREAM_T(__nv_bfloat16);
      else { set_error("pcb_head_bwd: bad dtype %d", dtype); return PCB_ERR_INVALID; }
#undef PCB_HEAD_STREAM_T
#undef PCB_HEAD_STREAM
      PCB_CHECK_LAUNCH("pcb_head_bwd");
      return PCB_OK;
    }
  }
  const int TV = C <= 128 ? 128 : 32;
  const size_t smem = (size_t)(C * ncls + ncls * TV + TV * (C + 1)) * sizeof(float);
  PCB_CHECK_ARG(smem <= 200 * 1024, "pcb_head_bwd: shared memory");
  int blocks = (int)((nvox + TV - 1) / TV);
  if (blocks > 148 * 4) blocks = 148 * 4;
  dim3 grid((unsigned)blocks, (unsigned)N);
#define PCB_HEAD_BWD(T)                                                                                                    \
  do {                                                                                                                     \
    cudaFuncSetAttribute(head_bwd_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);                     \
    head_bwd_kernel<T><<<grid, 128, smem, st>>>((const T*)dout, (const uint4*)x, w, (uint4*)dx, dW, db, (int)C, (int)ncls, nvox, TV); \
  } while (0)
  if (dtype == PCB_F32) PCB_HEAD_BWD(float);
  else if (dtype == PCB_F16) PCB_HEAD_BWD(__half);
  else if (dtype == PCB_BF16) PCB_HEAD_BWD(__nv_bfloat16);
  else { set_error("pcb_head_bwd: bad dtype %d", dtype); return PCB_ERR_INVALID; }
#undef PCB_HEAD_BWD
  PCB_CHECK_LAUNCH("pcb_head_bwd");
  return PCB_OK;
}

extern "C" int pcb_stem_bwd(const void* g, const void* x, int in_dtype, double* dW, double* db, int64_t N, int64_t Cin,
                            int64_t C, int64_t nvox, void* stream) {
  PCB_CHECK_ARG(g && x && dW && db, "pcb_stem_bwd: null argument");
  PCB_CHECK_ARG(C % 8 == 0 && C > 0 && Cin > 0 && Cin <= 16, "pcb_stem_bwd: Cin must be <= 16");
  const int64_t items = nvox * (C / 8);
  int blocks = (int)((items + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  dim3 grid((unsigned)blocks, (unsigned)N);
  const size_t smem = (size_t)C * (Cin + 1) * sizeof(double);
  cudaStream_t st = (cudaStream_t)stream;
  {
    const int ch = (int)(C >> 3);
    static const bool no_stream = getenv("PCB_NO_STEM_STREAM") != nullptr;
    if (!no_stream && Cin <= 4 && ch <= 32 && (ch & (ch - 1)) == 0 && N <= 65535 && in_dtype >= PCB_F32 && in_dtype <= PCB_BF16) {
      int sh = 0;
      while ((1 << sh) < ch) ++sh;
      int64_t nb = (items + 256 * 8 - 1) / (256 * 8);
      if (nb > 148 * 6) nb = 148 * 6;
      if (nb < 1) nb = 1;
      dim3 g2((unsigned)nb, (unsigned)N);
#define PCB_STEMB(T, K) stem_bwd_stream_kernel<T, K><<<g2, 256, (size_t)C * (K + 1) * sizeof(double), st>>>((const uint4*)g, (const T*)x, dW, db, (int)C, nvox, sh)
#define PCB_STEMB_T(T)                                                                        \
      do {                                                                                    \
        if (Cin == 1) PCB_STEMB(T, 1); else if (Cin == 2) PCB_STEMB(T, 2);                    \
        else if (Cin == 3) PCB_STEMB(T, 3); else PCB_STEMB(T, 4);                             \
      } while (0)
      if (in_dtype == PCB_F32) PCB_STEMB_T(float);
      else if (in_dtype == PCB_F16) PCB_STEMB_T(__half);
      else PCB_STEMB_T(__nv_bfloat16);
#undef PCB_STEMB_T
#undef PCB_STEMB
      PCB_CHECK_LAUNCH("pcb_stem_bwd");
      return PCB_OK;
    }
  }
  if (in_dtype == PCB_F32) stem_bwd_kernel<float><<<grid, 256, smem, st>>>((const uint4*)g, (const float*)x, dW, db, (int)Cin, (int)C, nvox);
  else if (in_dtype == PCB_F16) stem_bwd_kernel<__half><<<grid, 256, smem, st>>>((const uint4*)g, (const __half*)x, dW, db, (int)Cin, (int)C, nvox);
  else if (in_dtype == PCB_BF16) stem_bwd_kernel<__nv_bfloat16><<<grid, 256, smem, st>>>((const uint4*)g, (const __nv_bfloat16*)x, dW, db, (int)Cin, (int)C, nvox);
  else { set_error("pcb_stem_bwd: bad dtype %d", in_dtype); return PCB_ERR_INVALID; }
  PCB_CHECK_LAUNCH("pcb_stem_bwd");
  return PCB_OK;
}
