// pcb200 — MedNeXt conv2 -> GELU -> conv3 for the deep levels (C >= 128) as two launches of one
// warp-specialised persistent tcgen05 GEMM.
//
// Replaces, for C > 64, the torch.nn modules `nnunet_mednext` builds for MedNeXtBlock.conv2/act/conv3 (+ the
// res_conv of the resampling blocks) behind connectomics/models/architectures/mednext_models.py:374-380.
// At these levels the expanded activation is small (L2-resident) while the weights no longer fit next to the
// tiles in shared memory, so the block is NOT fused across the hidden dimension: splitting it over output
// columns gives (rows/128) x (N/BN) independent tiles — enough CTAs to fill 148 SMs even at 10^3 voxels.
//
//   launch 1:  Hact[M, H]  = GELU( GN(y)[M, C] W2^T + b2 )
//   launch 2:  out [M, Co] = Hact W3^T + b3  (+ xs[rx] Wr^T + br)  (+ residual / skip), zero on padded rows
//
// Kernel: 4 loader warps (one [128 x 64] A chunk + one [BN x 64] weight chunk per stage, GroupNorm applied in
// registers, K-major no-swizzle canonical layout), 1 MMA warp (tcgen05.mma, fp32 accumulate in TMEM, two
// accumulator buffers), 2 x 4 epilogue warps alternating tiles (TMEM -> bias/GELU/residual -> bf16 -> HBM).
// Every hand-off is an mbarrier; there is no __syncthreads in the tile loop.
#include "../../include/pcb200.h"
#include <stdlib.h>
#include <string.h>

#include "pcb_common.cuh"

namespace pcb {

struct GemmArgs {
  const uint4* a;  const uint4* a2;      // A segment 1 rows [N][Va][K1], segment 2 rows [N][Va2][K2] (nullable)
  const uint4* b;  const uint4* b2;      // weights [Nout][K1], [Nout][K2], K-major bf16
  const double* stats; const float* gamma; const float* beta; float inv_count;   // GroupNorm on segment 1 (nullable)
  const float* bias; const float* bias2; // [Nout]; bias2 nullable
  const uint4* res;                      // [N][Vout][Nout] bf16 added to every in-range row (nullable)
  uint4* out;                            // [N][Vout][Nout] bf16
  int K1, K2, Nout, BN;
  int gelu;                              // epilogue activation
  int map1;                              // 0: segment-1 row = output row, 1: row source of `mode`
  int mask;                              // 1: rows without a source (UP padding) produce 0 (+ residual)
  int mode;                              // PCB_DW_*: output row -> (y row, xs row)
  int o1, o2, x1, x2;
  int64_t Va, Va2, Vout;
  int N, S;                              // samples, pipeline stages
  int64_t tps;                           // 128-row tiles per sample
  int ntn;                               // column tiles
  int64_t ntiles;                        // N * tps * ntn
  uint32_t dm2, dm1; int ds2, ds1;       // exact division by o2 / o1 (see mf_fdiv in mednext_fwd.cu)
  // ---- backward (data-gradient) variants
  int bres;                              // weights resident: the CTA keeps ONE column tile, all K chunks of B staged once
  int dual;                              // segment 2 feeds a SECOND accumulator: out = GELU(acc1 + bias), out2 = acc2 * GELU'(acc1 + bias)
  uint4* out2;                           // [N][Vout][Nout] bf16 (dual)
  int map2;                              // segment-2 row source: 0 per `mode`, 1 output row, 2 output row + 1 on every axis (o1, o2 = box dims)
  int async_ld;                          // operands without a GroupNorm affine go global -> shared by cp.async (PCB_GW_ASYNC=0: registers)
  double* gst;                           // GroupNorm-backward sums [N][2][Nout] f64 (+=): S1 = sum g, S2 = sum g * xhat with xhat from
                                         // `res` (= y) and `stats`; the bf16-rounded g is what gets stored and summed
};

constexpr int GW_LOAD = 4, GW_EPI = 8, GW_THREADS = 32 * (GW_LOAD + GW_EPI + 1);

__device__ __forceinline__ int gw_fdiv(int n, uint32_t m, int sh) {
  return (int)(((uint64_t)(uint32_t)n * (uint64_t)m) >> sh);
}

__device__ __forceinline__ void gw_row_sources(const GemmArgs& a, int ov, int& ry, int& rx) {
  ry = -1; rx = -1;
  if (ov >= (int)a.Vout) return;
  if (a.map2 == 1) { ry = ov; rx = ov; return; }
  if (a.map2 == 2) {
    const int t = gw_fdiv(ov, a.dm2, a.ds2), ox = ov - t * a.o2;
    const int oz = gw_fdiv(t, a.dm1, a.ds1), oy = t - oz * a.o1;
    ry = ov;
    rx = ((oz + 1) * (a.o1 + 1) + (oy + 1)) * (a.o2 + 1) + (ox + 1);
    return;
  }
  if (a.mode == PCB_DW_SAME) { ry = ov; return; }
  const int t = gw_fdiv(ov, a.dm2, a.ds2), ox = ov - t * a.o2;
  const int oz = gw_fdiv(t, a.dm1, a.ds1), oy = t - oz * a.o1;
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

// One warp stages 128 rows x 64 bf16 columns (8 chunks of 16 B per row; row pitch `pitch8` chunks in global
// memory) into the K-major canonical layout ((r>>3)*1024 + c8*128 + (r&7)*16).  Same lane mapping as
// mf_stage_tile: a quarter-warp fills one 128-byte core matrix per shared store; the lane's two channel
// chunks are fixed, so the GroupNorm affine stays in registers.  TAB: row sources from `tab` (-1 = zero row).
template <bool NORM, bool TAB>
__device__ __forceinline__ void gw_stage_k64(uint8_t* __restrict__ dst, const uint4* __restrict__ src, int64_t pitch8,
                                             const int* __restrict__ tab, int row0, int nvalid,
                                             const float* __restrict__ sc, const float* __restrict__ sh, int lane) {
  const int rl = lane & 7, cs = lane >> 3;
  uint64_t ps[2][4], pt[2][4];
  if (NORM) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const float4* sp = reinterpret_cast<const float4*>(sc + (cs + 4 * j) * 8);
      const float4* tp = reinterpret_cast<const float4*>(sh + (cs + 4 * j) * 8);
      const float4 s0 = sp[0], s1 = sp[1], t0 = tp[0], t1 = tp[1];
      ps[j][0] = pk2(s0.x, s0.y); ps[j][1] = pk2(s0.z, s0.w); ps[j][2] = pk2(s1.x, s1.y); ps[j][3] = pk2(s1.z, s1.w);
      pt[j][0] = pk2(t0.x, t0.y); pt[j][1] = pk2(t0.z, t0.w); pt[j][2] = pk2(t1.x, t1.y); pt[j][3] = pk2(t1.z, t1.w);
    }
  }
  uint8_t* dl = dst + cs * 128 + rl * 16;
#pragma unroll 1
  for (int b = 0; b < 4; ++b) {
    uint4 v[8];
    uint32_t ok = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int rg = b * 4 + (k >> 1), j = k & 1;
      const int r = rg * 8 + rl;
      int ry;
      if (TAB) ry = tab[r]; else ry = r < nvalid ? row0 + r : -1;
      v[k] = make_uint4(0, 0, 0, 0);
      if (ry >= 0) { v[k] = __ldg(src + (int64_t)ry * pitch8 + (cs + 4 * j)); ok |= 1u << k; }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int rg = b * 4 + (k >> 1), j = k & 1;
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
      *reinterpret_cast<uint4*>(dl + rg * 1024 + j * 512) = o;
    }
  }
}

// The same tile with cp.async (operands that need no GroupNorm affine: hidden activations, dOut, dh, weights): 32 copies of
// 16 B per lane, the whole 16 KB chunk in flight at once instead of four 4 KB register batches (round 2: the deep GEMMs were
// bound by the bytes their four loader warps kept in flight).  Rows without a source are zero-filled (src-size 0).  The caller
// waits for the group and fences the async proxy before it arrives on the stage barrier.
template <bool TAB>
__device__ __forceinline__ void gw_stage_k64_async(uint8_t* __restrict__ dst, const uint4* __restrict__ src, int64_t pitch8,
                                                   const int* __restrict__ tab, int row0, int nvalid, int lane) {
  const int rl = lane & 7, cs = lane >> 3;
  const uint32_t dl = smem_u32(dst) + cs * 128 + rl * 16;
#pragma unroll 4
  for (int rg = 0; rg < 16; ++rg) {
    const int r = rg * 8 + rl;
    int ry;
    if (TAB) ry = tab[r]; else ry = r < nvalid ? row0 + r : -1;
    const uint4* p = src + (int64_t)(ry >= 0 ? ry : 0) * pitch8 + cs;
    const uint32_t n = ry >= 0 ? 16u : 0u;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dl + rg * 1024), "l"(p), "r"(n) : "memory");
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dl + rg * 1024 + 512), "l"(p + 4), "r"(n) : "memory");
  }
}
__device__ __forceinline__ void gw_async_join() {
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}

__global__ void __launch_bounds__(GW_THREADS, 1) gemm_ws_kernel(GemmArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int S = a.S, BN = a.BN;
  const int kch1 = a.K1 >> 6, kch = kch1 + (a.K2 >> 6);
  const bool gstat = a.gst != nullptr;
  const bool norm = a.stats != nullptr && !gstat;
  const int nsc = norm ? a.N * a.K1 : (gstat ? a.N * a.Nout : 0);   // per-(sample, channel) GroupNorm constants
  const int accw = a.dual ? 2 * BN : BN;                             // TMEM columns per accumulator buffer
  uint8_t* sA = smem;                                        // S x [128 x 64]
  const int bstage = (BN > 128 ? BN : 128) * 128;            // a loader call always writes 128 rows
  uint8_t* sB = sA + S * 16384;                              // S x [max(BN,128) x 64]
  float* sScale = reinterpret_cast<float*>(sB + (a.bres ? kch : S) * bstage); // [N][K1]
  float* sShift = sScale + nsc;
  int* sRow = reinterpret_cast<int*>(sShift + nsc);   // [4 loader warps][2][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sRow + GW_LOAD * 2 * 128);
  uint64_t* full = bars;             // [S]  loaders -> MMA
  uint64_t* empty = bars + 4;        // [S]  MMA -> loaders
  uint64_t* acc_full = bars + 8;     // [2]  MMA -> epilogue
  uint64_t* acc_empty = bars + 10;   // [2]  epilogue -> MMA
  uint64_t* b_full = bars + 12;      // resident weights staged (4 loader warps)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);

  const uint32_t tmem_cols = tmem_cols_pow2(2 * accw);
  if (warp == GW_LOAD + GW_EPI) tmem_alloc(tmem_slot, tmem_cols);
  if (tid == 0) {
    for (int i = 0; i < 4; ++i) { mbar_init(&full[i], 32); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 128); }
    mbar_init(b_full, GW_LOAD);
    fence_mbar_init();
  }
  if (norm) {
    for (int i = tid; i < a.N * a.K1; i += GW_THREADS) {
      const int n = i / a.K1, c = i - n * a.K1;
      const double sm = a.stats[(int64_t)n * 2 * a.K1 + c], q = a.stats[(int64_t)n * 2 * a.K1 + a.K1 + c];
      const double mean = sm * (double)a.inv_count;
      double var = q * (double)a.inv_count - mean * mean;
      if (var < 0.0) var = 0.0;
      const float g = a.gamma[c] * (float)(1.0 / sqrt(var + 1e-5));
      sScale[i] = g;
      sShift[i] = a.beta[c] - (float)mean * g;
    }
  } else if (gstat) {   // xhat = y * rstd - mean * rstd
    for (int i = tid; i < nsc; i += GW_THREADS) {
      const int n = i / a.Nout, c = i - n * a.Nout;
      const double sm = a.stats[(int64_t)n * 2 * a.Nout + c], q = a.stats[(int64_t)n * 2 * a.Nout + a.Nout + c];
      const double mean = sm * (double)a.inv_count;
      double var = q * (double)a.inv_count - mean * mean;
      if (var < 0.0) var = 0.0;
      const float rstd = (float)(1.0 / sqrt(var + 1e-5));
      sScale[i] = rstd;
      sShift[i] = (float)mean * rstd;
    }
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // this CTA's tiles: t = blockIdx.x + i * gridDim.x;  column tile fastest so that CTAs running side by side
  // share the A rows through L2
  // resident weights: column tile = blockIdx.x % ntn for the CTA's whole life, row tiles strided by gridDim.x / ntn
  const int64_t mtiles = a.tps * a.N;
  const int nrc = a.bres ? (int)(gridDim.x / a.ntn) : 0;
  const int rowcta = a.bres ? (int)(blockIdx.x / a.ntn) : 0, nt_fixed = a.bres ? (int)(blockIdx.x % a.ntn) : 0;
  const int64_t ntl = a.bres ? (rowcta < mtiles ? (mtiles - rowcta + nrc - 1) / nrc : 0)
                             : ((int64_t)blockIdx.x < a.ntiles ? (a.ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0);

  if (warp < GW_LOAD) {
    // ===================================================================== loaders
    int* rowY = sRow + warp * 256;
    int* rowX = rowY + 128;
    int64_t cached_mt = -1;
    const bool need_tab = (a.map1 && a.mode == PCB_DW_UP) || a.K2 > 0;
    const int64_t nchunks = ntl * kch;
    if (a.bres) {   // the CTA's weight column tile, every K chunk, once
      if (ntl > 0) {
        for (int kc = warp; kc < kch; kc += GW_LOAD) {
          const bool seg1 = kc < kch1;
          const int64_t bp8 = (seg1 ? a.K1 : a.K2) >> 3;
          const uint4* bsrc = (seg1 ? a.b : a.b2) + (int64_t)nt_fixed * BN * bp8 + (seg1 ? kc : kc - kch1) * 8;
          for (int h = 0; h < BN; h += 128) {
            if (a.async_ld) gw_stage_k64_async<false>(sB + kc * bstage + h * 128, bsrc + (int64_t)h * bp8, bp8, nullptr, 0, min(128, BN - h), lane);
            else gw_stage_k64<false, false>(sB + kc * bstage + h * 128, bsrc + (int64_t)h * bp8, bp8, nullptr, 0, min(128, BN - h),
                                            nullptr, nullptr, lane);
          }
        }
        if (a.async_ld) gw_async_join();
        fence_proxy_async_smem();
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(b_full);
    }
    // A stage is filled by ONE warp for the whole kernel (warp w <-> stage w; warps >= S stay idle): a warp that skipped a
    // completion of `empty[s]` would mis-read its phase parity and overwrite a stage the MMA has not consumed — with S = 3
    // (MedNeXt-L up_1: ten resident weight chunks leave room for three A stages) the old q += 4 walk did exactly that and hung.
    for (int64_t q = warp; warp < S && q < nchunks; q += S) {
      const int64_t i = q / kch;
      const int kc = (int)(q - i * kch);
      const int64_t t = blockIdx.x + i * gridDim.x;
      const int64_t mt = a.bres ? rowcta + i * nrc : t / a.ntn;
      const int nt = a.bres ? nt_fixed : (int)(t - mt * a.ntn);
      const int n = (int)(mt / a.tps);
      const int tile0 = (int)((mt - (int64_t)n * a.tps) * 128);
      const int s = (int)(q % S);
      const int64_t u = q / S;
      if (u >= 1) mbar_wait(&empty[s], (uint32_t)((u - 1) & 1));
      if (need_tab && mt != cached_mt) {
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          int ry, rx;
          gw_row_sources(a, tile0 + lane + 32 * k, ry, rx);
          rowY[lane + 32 * k] = ry; rowX[lane + 32 * k] = rx;
        }
        __syncwarp();
        cached_mt = mt;
      }
      const int nvalid = min(128, (int)a.Vout - tile0);
      uint8_t* dA = sA + s * 16384;
      uint8_t* dB = sB + s * bstage;
      const uint4* bsrc;
      int64_t bp8;
      if (kc < kch1) {
        const int64_t p8 = a.K1 >> 3;
        const uint4* src = a.a + (int64_t)n * a.Va * p8 + kc * 8;
        const float* sc = sScale + n * a.K1 + kc * 64;
        const float* sh = sShift + n * a.K1 + kc * 64;
        const bool tab = a.map1 && a.mode == PCB_DW_UP;
        if (norm) {
          if (tab) gw_stage_k64<true, true>(dA, src, p8, rowY, tile0, nvalid, sc, sh, lane);
          else gw_stage_k64<true, false>(dA, src, p8, rowY, tile0, nvalid, sc, sh, lane);
        } else if (a.async_ld) {
          if (tab) gw_stage_k64_async<true>(dA, src, p8, rowY, tile0, nvalid, lane);
          else gw_stage_k64_async<false>(dA, src, p8, rowY, tile0, nvalid, lane);
        } else {
          if (tab) gw_stage_k64<false, true>(dA, src, p8, rowY, tile0, nvalid, sc, sh, lane);
          else gw_stage_k64<false, false>(dA, src, p8, rowY, tile0, nvalid, sc, sh, lane);
        }
        bp8 = p8;
        bsrc = a.b + (int64_t)nt * BN * p8 + kc * 8;
      } else {
        const int64_t p8 = a.K2 >> 3;
        const uint4* src = a.a2 + (int64_t)n * a.Va2 * p8 + (kc - kch1) * 8;
        if (a.async_ld) gw_stage_k64_async<true>(dA, src, p8, rowX, 0, 0, lane);
        else gw_stage_k64<false, true>(dA, src, p8, rowX, 0, 0, nullptr, nullptr, lane);
        bp8 = p8;
        bsrc = a.b2 + (int64_t)nt * BN * p8 + (kc - kch1) * 8;
      }
      if (!a.bres)
        for (int h = 0; h < BN; h += 128) {
          if (a.async_ld) gw_stage_k64_async<false>(dB + h * 128, bsrc + (int64_t)h * bp8, bp8, nullptr, 0, min(128, BN - h), lane);
          else gw_stage_k64<false, false>(dB + h * 128, bsrc + (int64_t)h * bp8, bp8, nullptr, 0, min(128, BN - h), nullptr, nullptr, lane);
        }
      if (a.async_ld) gw_async_join();             // no-op for a chunk that only went through registers
      fence_proxy_async_smem();
      mbar_arrive(&full[s]);
    }
  } else if (warp == GW_LOAD + GW_EPI) {
    // ===================================================================== MMA issuer (one thread)
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(128, BN, 0, 0);
      int64_t q = 0;
      if (a.bres && ntl > 0) { mbar_wait(b_full, 0); tc_fence_after(); }
      for (int64_t i = 0; i < ntl; ++i) {
        const int g = (int)(i & 1);
        if (i >= 2) mbar_wait(&acc_empty[g], (uint32_t)(((i >> 1) - 1) & 1));
        tc_fence_after();
        const uint32_t acc0 = tmem_base + g * accw;
        for (int kc = 0; kc < kch; ++kc, ++q) {
          const bool second = a.dual && kc >= kch1;
          const uint32_t acc = second ? acc0 + BN : acc0;
          const int kfirst = second ? kch1 : 0;
          const int s = (int)(q % S);
          mbar_wait(&full[s], (uint32_t)((q / S) & 1));
          tc_fence_after();
          const uint64_t dA = umma_desc(smem_u32(sA + s * 16384), 128, 1024);
          const uint64_t dB = umma_desc(smem_u32(sB + (a.bres ? kc : s) * bstage), 128, 1024);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(acc, dA + (uint64_t)(k * 16), dB + (uint64_t)(k * 16), idesc, (kc > kfirst || k > 0) ? 1u : 0u);
          tc_commit(&empty[s]);
        }
        tc_commit(&acc_full[g]);
      }
    }
  } else {
    // ===================================================================== epilogue warpgroups
    const int eg = (warp - GW_LOAD) >> 2;
    const int wq = warp & 3;
    const int row = wq * 32 + lane;
    const uint32_t lane_off = (uint32_t)(wq * 32) << 16;
    const int n8 = a.Nout >> 3;
    for (int64_t i = eg; i < ntl; i += 2) {
      const int64_t t = blockIdx.x + i * gridDim.x;
      const int64_t mt = a.bres ? rowcta + i * nrc : t / a.ntn;
      const int nt = a.bres ? nt_fixed : (int)(t - mt * a.ntn);
      const int n = (int)(mt / a.tps);
      const int ovi = (int)((mt - (int64_t)n * a.tps) * 128) + row;
      const bool in_range = ovi < (int)a.Vout;
      bool valid = in_range;
      if (a.mask && in_range && a.mode == PCB_DW_UP) {
        int ry, rx;
        gw_row_sources(a, ovi, ry, rx);
        valid = ry >= 0;
      }
      const int64_t orow = ((int64_t)n * a.Vout + ovi) * n8 + (int64_t)nt * (BN >> 3);
      const int col0 = nt * BN;
      mbar_wait(&acc_full[eg], (uint32_t)((i >> 1) & 1));
      tc_fence_after();
      const uint32_t trow = tmem_base + eg * accw + lane_off;
      if (a.dual) {
        // Hact = GELU(h), dh = dG * GELU'(h) with h = acc1 + bias (both bf16 to HBM for the weight-gradient GEMMs)
#pragma unroll 1
        for (int c16 = 0; c16 < BN / 16; ++c16) {
          uint32_t v1[16], v2[16];
          tmem_ld16(trow + c16 * 16, v1);
          tmem_ld16(trow + BN + c16 * 16, v2);
          tmem_ld_wait();
          if (!in_range) continue;
          const float4* bp = reinterpret_cast<const float4*>(a.bias + col0 + c16 * 16);
          float ha[16], dv[16];
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            const float4 b = __ldg(bp + j4);
            uint64_t val, grad;
            gelu_fast_vg2(add2(pk2(__uint_as_float(v1[4 * j4]), __uint_as_float(v1[4 * j4 + 1])), pk2(b.x, b.y)), val, grad);
            upk2(val, ha[4 * j4], ha[4 * j4 + 1]);
            upk2(mul2(grad, pk2(__uint_as_float(v2[4 * j4]), __uint_as_float(v2[4 * j4 + 1]))), dv[4 * j4], dv[4 * j4 + 1]);
            gelu_fast_vg2(add2(pk2(__uint_as_float(v1[4 * j4 + 2]), __uint_as_float(v1[4 * j4 + 3])), pk2(b.z, b.w)), val, grad);
            upk2(val, ha[4 * j4 + 2], ha[4 * j4 + 3]);
            upk2(mul2(grad, pk2(__uint_as_float(v2[4 * j4 + 2]), __uint_as_float(v2[4 * j4 + 3]))), dv[4 * j4 + 2], dv[4 * j4 + 3]);
          }
          a.out[orow + c16 * 2] = pack8(ha);
          a.out[orow + c16 * 2 + 1] = pack8(ha + 8);
          a.out2[orow + c16 * 2] = pack8(dv);
          a.out2[orow + c16 * 2 + 1] = pack8(dv + 8);
        }
      } else if (gstat) {
        // g = bf16(acc) -> HBM; S1 += g, S2 += g * xhat (column sums over the tile rows, f64 atomics per warp)
        const float* rs = sScale + n * a.Nout + col0;
        const float* mr = sShift + n * a.Nout + col0;
        double* gs = a.gst + (int64_t)n * 2 * a.Nout + col0;
#pragma unroll 1
        for (int c16 = 0; c16 < BN / 16; ++c16) {
          uint32_t v[16];
          tmem_ld16(trow + c16 * 16, v);
          uint4 r0 = make_uint4(0, 0, 0, 0), r1 = make_uint4(0, 0, 0, 0);
          if (in_range) { r0 = __ldg(a.res + orow + c16 * 2); r1 = __ldg(a.res + orow + c16 * 2 + 1); }
          tmem_ld_wait();
          float g[16], gx[16];
          if (in_range) {
            float yv[16];
            unpack8(r0, yv);
            unpack8(r1, yv + 8);
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              g[j] = round_bf16(__uint_as_float(v[j]));
              gx[j] = g[j] * fmaf(yv[j], rs[c16 * 16 + j], -mr[c16 * 16 + j]);
            }
            a.out[orow + c16 * 2] = pack8(g);
            a.out[orow + c16 * 2 + 1] = pack8(g + 8);
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) { g[j] = 0.f; gx[j] = 0.f; }
          }
          warp_colsum16(g, lane);
          warp_colsum16(gx, lane);
          if (!(lane & 1)) {
            const int col = c16 * 16 + colsum16_col(lane);
            atomicAdd(gs + col, (double)g[0]);
            atomicAdd(gs + a.Nout + col, (double)gx[0]);
          }
        }
      } else
#pragma unroll 1
      for (int c16 = 0; c16 < BN / 16; ++c16) {
        uint32_t v[16];
        tmem_ld16(trow + c16 * 16, v);
        uint4 r0 = make_uint4(0, 0, 0, 0), r1 = make_uint4(0, 0, 0, 0);
        if (a.res != nullptr && in_range) { r0 = __ldg(a.res + orow + c16 * 2); r1 = __ldg(a.res + orow + c16 * 2 + 1); }
        tmem_ld_wait();
        if (!in_range) continue;
        const float4* bp = reinterpret_cast<const float4*>(a.bias + col0 + c16 * 16);
        float o[16];
        if (a.gelu) {
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            const float4 b = __ldg(bp + j4);
            gelu_fast2p(add2(pk2(__uint_as_float(v[4 * j4]), __uint_as_float(v[4 * j4 + 1])), pk2(b.x, b.y)), o[4 * j4], o[4 * j4 + 1]);
            gelu_fast2p(add2(pk2(__uint_as_float(v[4 * j4 + 2]), __uint_as_float(v[4 * j4 + 3])), pk2(b.z, b.w)), o[4 * j4 + 2], o[4 * j4 + 3]);
          }
        } else {
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            float4 b = __ldg(bp + j4);
            if (a.bias2 != nullptr) {
              const float4 b2 = __ldg(reinterpret_cast<const float4*>(a.bias2 + col0 + c16 * 16) + j4);
              b.x += b2.x; b.y += b2.y; b.z += b2.z; b.w += b2.w;
            }
            o[4 * j4] = valid ? __uint_as_float(v[4 * j4]) + b.x : 0.f;
            o[4 * j4 + 1] = valid ? __uint_as_float(v[4 * j4 + 1]) + b.y : 0.f;
            o[4 * j4 + 2] = valid ? __uint_as_float(v[4 * j4 + 2]) + b.z : 0.f;
            o[4 * j4 + 3] = valid ? __uint_as_float(v[4 * j4 + 3]) + b.w : 0.f;
          }
          if (a.res != nullptr) {
            float f[8];
            unpack8(r0, f);
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] += f[j];
            unpack8(r1, f);
#pragma unroll
            for (int j = 0; j < 8; ++j) o[8 + j] += f[j];
          }
        }
        a.out[orow + c16 * 2] = pack8(o);
        a.out[orow + c16 * 2 + 1] = pack8(o + 8);
      }
      tc_fence_before();
      mbar_arrive(&acc_empty[eg]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == GW_LOAD + GW_EPI) tmem_dealloc(tmem_base, tmem_cols);
}

static void gw_magic(uint32_t d, uint32_t& m, int& sh) {
  int l = 0;
  while ((1ull << l) < d) ++l;
  const unsigned __int128 p = (unsigned __int128)1 << (31 + l);
  m = (uint32_t)((p + d - 1) / d);
  sh = 31 + l;
}

// picks BN / stages, launches; returns false when the shape is not supported (caller falls back)
static bool gw_launch(GemmArgs a, cudaStream_t st) {
  if (a.K1 % 64 != 0 || a.K2 % 64 != 0 || a.Nout % 64 != 0 || a.N > 8 || a.Vout >= (1ll << 30)) return false;
  a.tps = (a.Vout + 127) / 128;
  const int64_t mtiles = a.tps * a.N;
  // widest column tile that still yields >= 2 tiles per SM, else the narrowest legal one
  int bn = a.dual ? 128 : 256;   // dual: two accumulators x two buffers must fit the 512 TMEM columns
  while (bn > 64 && (a.Nout % bn != 0 || mtiles * (a.Nout / bn) < 2 * 148)) bn >>= 1;
  if (a.Nout % bn != 0) return false;
  a.BN = bn;
  a.ntn = a.Nout / bn;
  a.ntiles = mtiles * a.ntn;
  const size_t fixed = (a.stats ? (size_t)2 * a.N * (a.gst ? a.Nout : a.K1) * 4 : 0) + GW_LOAD * 2 * 128 * 4 + 13 * 8 + 16 + 128;
  const size_t bstage = (size_t)(bn > 128 ? bn : 128) * 128;
  const int kch = (a.K1 + a.K2) >> 6;
  // resident weights when every K chunk of the column tile fits next to >= 3 A stages and each CTA sees >= 3 row tiles
  static const bool no_bres = getenv("PCB_NO_BRES") != nullptr;
  int S = 4;
  size_t dyn = 0;
  a.bres = 0;
  if (!no_bres && a.ntn <= 148 && mtiles >= 3 * (148 / a.ntn) && fixed + kch * bstage + 3 * 16384 <= 227 * 1024) {
    a.bres = 1;
    while (S > 3 && fixed + kch * bstage + (size_t)S * 16384 > 227 * 1024) --S;
    dyn = fixed + kch * bstage + (size_t)S * 16384;
  } else {
    const size_t stage = 16384 + bstage;
    while (S > 2 && fixed + S * stage > 227 * 1024) --S;
    if (fixed + S * stage > 227 * 1024) return false;
    dyn = fixed + S * stage;
  }
  a.S = S;
  { const char* e = getenv("PCB_GW_ASYNC"); a.async_ld = (e && e[0] == '0') ? 0 : 1; }
  gw_magic((uint32_t)(a.o2 > 0 ? a.o2 : 1), a.dm2, a.ds2);
  gw_magic((uint32_t)(a.o1 > 0 ? a.o1 : 1), a.dm1, a.ds1);
  static DevFlag conf;
  if (!conf) {
    cudaFuncSetAttribute(gemm_ws_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (cudaFuncSetAttribute(gemm_ws_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) {
      cudaGetLastError();
      return false;
    }
    conf = true;
  }
  int ctas = 148;
  if (a.ntiles < ctas) ctas = (int)a.ntiles;
  if (a.bres) ctas = (148 / a.ntn) * a.ntn;     // every CTA owns one column tile; mtiles >= 3 row tiles per CTA (checked above)
  gemm_ws_kernel<<<ctas, GW_THREADS, dyn, st>>>(a);
  return true;
}

// Data gradient of conv2 -> GELU -> conv3 for the deep levels (what autograd does for MedNeXtBlock under
// connectomics/training/lightning/model.py:863-910), rows in y space:
//   launch 1 (dual accumulators):  Hpre = GN(y) W2^T,  dG = dOut[row map] W3   ->  Hact = GELU(Hpre + b2),  dh = dG * GELU'(Hpre + b2)
//   launch 2 (stats epilogue)   :  dYhat = dh W2  ->  bf16 + GroupNorm-backward sums
int mlp_bwd_deep(const void* y, const double* stats, const float* gamma, const float* beta, const void* w2, const float* b2,
                 const void* w3t, const void* w2t, const void* dout, void* hact, void* dh, void* dyhat, double* gstats,
                 int64_t N, const int64_t y_size[3], int64_t C, int64_t H, int64_t Co, int mode, void* stream) {
  if (getenv("PCB_NO_DEEP_BWD") != nullptr || getenv("PCB_NO_DEEP") != nullptr) return 0;
  if (C < 128 || C % 64 != 0 || H % 64 != 0 || Co % 64 != 0 || N < 1 || N > 8) return 0;
  const int64_t Vy = y_size[0] * y_size[1] * y_size[2];
  if (Vy <= 0 || (y_size[0] + 1) * (y_size[1] + 1) * (y_size[2] + 1) >= (1ll << 30)) return 0;
  GemmArgs g;
  memset(&g, 0, sizeof(g));
  g.o1 = (int)y_size[1]; g.o2 = (int)y_size[2];
  g.Vout = Vy; g.mode = PCB_DW_SAME; g.N = (int)N;
  g.inv_count = (float)(1.0 / (double)Vy);
  GemmArgs g1 = g;
  g1.a = (const uint4*)y; g1.K1 = (int)C; g1.Va = Vy; g1.b = (const uint4*)w2; g1.Nout = (int)H;
  g1.stats = stats; g1.gamma = gamma; g1.beta = beta; g1.bias = b2;
  g1.a2 = (const uint4*)dout; g1.K2 = (int)Co; g1.b2 = (const uint4*)w3t;
  g1.Va2 = mode == PCB_DW_UP ? (y_size[0] + 1) * (y_size[1] + 1) * (y_size[2] + 1) : Vy;
  g1.map2 = mode == PCB_DW_UP ? 2 : 1;
  g1.dual = 1; g1.out = (uint4*)hact; g1.out2 = (uint4*)dh;
  if (!gw_launch(g1, (cudaStream_t)stream)) return 0;
  count_launch();
  GemmArgs g2 = g;
  g2.a = (const uint4*)dh; g2.K1 = (int)H; g2.Va = Vy; g2.b = (const uint4*)w2t; g2.Nout = (int)C;
  g2.stats = stats; g2.res = (const uint4*)y; g2.out = (uint4*)dyhat; g2.gst = gstats;
  if (!gw_launch(g2, (cudaStream_t)stream)) { set_error("pcb_mlp_bwd(deep): launch 2 rejected the shape"); return PCB_ERR_INVALID; }
  count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("pcb_mlp_bwd(deep): %s", cudaGetErrorString(e)); return PCB_ERR_CUDA; }
  return 1;
}

}  // namespace pcb

using namespace pcb;

// Workspace (bytes) pcb_mlp_fwd_deep needs for the expanded activation; 0 = shape not served by the deep path
// (callers use pcb_mlp_fwd).
extern "C" int64_t pcb_mlp_fwd_deep_workspace(int64_t N, const int64_t out_size[3], int64_t C, int64_t H, int64_t Co,
                                              int64_t Cr) {
  if (!out_size || getenv("PCB_NO_DEEP") != nullptr) return 0;
  if (C < 128 || C % 64 != 0 || H % 64 != 0 || Co % 64 != 0 || (Cr > 0 && Cr % 64 != 0) || N < 1 || N > 8) return 0;
  const int64_t v = out_size[0] * out_size[1] * out_size[2];
  if (v <= 0 || v >= (1ll << 30)) return 0;
  return N * v * H * 2;
}

// Same contract as pcb_mlp_fwd (include/pcb200.h) plus `hact`: [N][Vout][H] bf16 workspace that receives
// GELU(conv2(GN(y))) — kept by the caller for the backward pass.
extern "C" int pcb_mlp_fwd_deep(const void* y, const double* stats, const float* gamma, const float* beta, const void* w2,
                                const float* b2, const void* w3, const float* b3, const void* res, const void* xs,
                                const void* wr, const float* br, void* out, void* hact, int64_t N,
                                const int64_t out_size[3], const int64_t xs_size[3], int64_t C, int64_t H, int64_t Co,
                                int64_t Cr, int mode, void* stream) {
  PCB_CHECK_ARG(y && stats && gamma && beta && w2 && b2 && w3 && b3 && out && hact && out_size, "pcb_mlp_fwd_deep: null argument");
  PCB_CHECK_ARG(mode >= PCB_DW_SAME && mode <= PCB_DW_UP, "pcb_mlp_fwd_deep: bad mode %d", mode);
  PCB_CHECK_ARG((wr == nullptr) || (xs && br && xs_size && Cr > 0 && mode != PCB_DW_SAME), "pcb_mlp_fwd_deep: bad res-conv arguments");
  PCB_CHECK_ARG(pcb_mlp_fwd_deep_workspace(N, out_size, C, H, Co, wr ? Cr : 0) > 0,
                "pcb_mlp_fwd_deep: unsupported shape (C=%lld H=%lld Co=%lld N=%lld)", (long long)C, (long long)H, (long long)Co, (long long)N);
  GemmArgs g;
  memset(&g, 0, sizeof(g));
  const int o0 = (int)out_size[0];
  g.o1 = (int)out_size[1]; g.o2 = (int)out_size[2];
  g.Vout = (int64_t)o0 * g.o1 * g.o2;
  g.mode = mode; g.N = (int)N;
  int64_t Vy = g.Vout, Vin = 0;
  if (mode == PCB_DW_UP) {
    PCB_CHECK_ARG(o0 >= 2 && g.o1 >= 2 && g.o2 >= 2, "pcb_mlp_fwd_deep: UP output too small");
    Vy = (int64_t)(o0 - 1) * (g.o1 - 1) * (g.o2 - 1);
  }
  if (wr) {
    g.x1 = (int)xs_size[1]; g.x2 = (int)xs_size[2];
    Vin = xs_size[0] * xs_size[1] * xs_size[2];
    if (mode == PCB_DW_UP)
      PCB_CHECK_ARG(xs_size[0] * 2 == o0 && g.x1 * 2 == g.o1 && g.x2 * 2 == g.o2, "pcb_mlp_fwd_deep: UP needs out_size == 2*xs_size");
    else
      PCB_CHECK_ARG((xs_size[0] - 1) / 2 + 1 == o0 && (g.x1 - 1) / 2 + 1 == g.o1 && (g.x2 - 1) / 2 + 1 == g.o2,
                    "pcb_mlp_fwd_deep: DOWN needs out_size == ceil(xs_size/2)");
  }
  // launch 1: Hact = GELU(GN(y) W2^T + b2), rows in output-row space
  GemmArgs g1 = g;
  g1.a = (const uint4*)y; g1.K1 = (int)C; g1.Va = Vy; g1.b = (const uint4*)w2; g1.Nout = (int)H;
  g1.stats = stats; g1.gamma = gamma; g1.beta = beta; g1.inv_count = (float)(1.0 / (double)Vy);
  g1.bias = b2; g1.out = (uint4*)hact; g1.gelu = 1; g1.map1 = 1; g1.mask = 0;
  if (!gw_launch(g1, (cudaStream_t)stream)) { set_error("pcb_mlp_fwd_deep: launch 1 rejected the shape"); return PCB_ERR_INVALID; }
  PCB_CHECK_LAUNCH("pcb_mlp_fwd_deep(conv2)");
  // launch 2: out = Hact W3^T + b3 (+ xs Wr^T + br) (+ res)
  GemmArgs g2 = g;
  g2.a = (const uint4*)hact; g2.K1 = (int)H; g2.Va = g.Vout; g2.b = (const uint4*)w3; g2.Nout = (int)Co;
  g2.bias = b3; g2.res = (const uint4*)res; g2.out = (uint4*)out; g2.gelu = 0; g2.map1 = 0; g2.mask = 1;
  if (wr) { g2.a2 = (const uint4*)xs; g2.K2 = (int)Cr; g2.Va2 = Vin; g2.b2 = (const uint4*)wr; g2.bias2 = br; }
  if (!gw_launch(g2, (cudaStream_t)stream)) { set_error("pcb_mlp_fwd_deep: launch 2 rejected the shape"); return PCB_ERR_INVALID; }
  PCB_CHECK_LAUNCH("pcb_mlp_fwd_deep(conv3)");
  return PCB_OK;
}
