/* pcb200 — C ABI of the B200-native hot path for PyTorch Connectomics.
 *
 * The reference (PytorchConnectomics/pytorch_connectomics) is pure Python; it has no FFI.  Its
 * "operator API" for this path is (1) the architecture registry / nn.Module forward contract and
 * (2) the sliding-window engine callable.  This header is the boundary a reference maintainer
 * would bind (ctypes stub in INTEGRATION.md): plain C, device/host pointers + sizes, every call
 * enqueues on the caller's cudaStream_t and never synchronises.  Each entry point cites the
 * reference code it replaces (paths relative to the reference repo root).
 *
 * Conventions
 *   - return 0 (PCB_OK) or a negative pcb_status; pcb_last_error() gives a thread-local message.
 *   - the caller owns every buffer; the library owns nothing but a few __constant__ tables.
 *   - activations inside the network are channels-last bf16:  [N, D, H, W, C]  ("NDHWC").
 *   - tensors crossing the reference API are NCDHW in the caller's dtype (pcb_dtype).
 *   - stream is a cudaStream_t passed as void*.
 */
#ifndef PCB200_H
#define PCB200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum { PCB_OK = 0, PCB_ERR_INVALID = -1, PCB_ERR_CUDA = -2, PCB_ERR_UNSUPPORTED = -3 } pcb_status;
typedef enum { PCB_F32 = 0, PCB_F16 = 1, PCB_BF16 = 2 } pcb_dtype;
typedef enum { PCB_BLEND_CONSTANT = 0, PCB_BLEND_BUMP = 1, PCB_BLEND_DISTANCE = 2 } pcb_blend;
typedef enum { PCB_PAD_CONSTANT = 0, PCB_PAD_REFLECT = 1, PCB_PAD_REPLICATE = 2, PCB_PAD_CIRCULAR = 3 } pcb_pad;
typedef enum { PCB_GRID_EAGER = 0, PCB_GRID_LAZY = 1, PCB_GRID_LAZY_SNAP = 2 } pcb_grid;
/* depthwise-conv flavour of a MedNeXt block (upstream nnunet_mednext blocks.py) */
typedef enum { PCB_DW_SAME = 0, PCB_DW_DOWN = 1, PCB_DW_UP = 2 } pcb_dw_mode;

const char* pcb_last_error(void);
int pcb_version(void);
/* number of pcb200 kernels launched by this process so far (bench.py reports the delta) */
int64_t pcb_launch_count(void);
/* 1 when the library was built for sm_100a and the current device is cc 10.x */
int pcb_device_ok(void);

/* ------------------------------------------------------------------ sliding window: host INT logic
 * connectomics/inference/window.py:57-89  compute_scan_interval (Python round() = half-even) */
int pcb_sw_scan_interval(const int64_t image[3], const int64_t roi[3], const double overlap[3],
                         int64_t interval_out[3]);
/* window.py:92-134 dense_patch_slices (PCB_GRID_EAGER) and lazy.py:269-365 window offsets
 * (PCB_GRID_LAZY / _SNAP, region filter `off < stop && off+roi > start`; pass region=NULL for all).
 * Writes up to `capacity` window starts (z,y,x triples, z-major product order) and the total count. */
int pcb_sw_plan(int grid_kind, const int64_t image[3], const int64_t roi[3], const double overlap[3],
                const int64_t region[6], int64_t* starts_out, int64_t capacity, int64_t* count_out);

/* ------------------------------------------------------------------ sliding window: device kernels
 * window.py:137-243  importance map built in `dtype` arithmetic (bump / constant / distance).
 * roi is left-padded with 1s to 3-D; `ndim` says how many trailing axes are real. */
int pcb_sw_importance_map(int blend, const int64_t roi[3], int ndim, int dtype, double min_value,
                          void* map_out, void* stream);
/* window.py:464-527 _extract_padded_patch_batch: gather n windows [n,C,roi] from vol [1,C,D,H,W];
 * `starts` is a HOST array of n (z,y,x) triples (passed to the kernel as launch parameters: no device allocation, copy
 * or synchronisation — the call is CUDA-graph capturable); pad per window relative to the in-image crop. */
int pcb_sw_extract(const void* vol, int dtype, int64_t C, const int64_t image[3], const int64_t roi[3],
                   const int64_t* starts, int64_t n, int pad_mode, double cval, void* out, void* stream);
/* window.py:648-655 _accumulate for ONE window: value[:, lo:hi] += pred[plo:phi]*map ; weight += map
 * (all in `dtype` arithmetic, mul then add, no FMA — bit-identical to the torch expression).
 * pred is [Cout, roi] ; boxes allow the clipped sub-box form of lazy.py:1216-1227. */
int pcb_sw_accumulate(const void* pred, const void* map, void* value, void* weight, int dtype,
                      int64_t Cout, const int64_t roi[3], const int64_t out_size[3],
                      const int64_t pred_lo[3], const int64_t out_lo[3], const int64_t box[3],
                      void* stream);
/* window.py:641-655: the `for idx ...: _accumulate(...)` loop over one network batch — n FULL windows (pred [n,Cout,roi],
 * `starts` = HOST array of n (z,y,x) triples inside the accumulator) in ONE launch per 16 windows, bit-identical to
 * calling pcb_sw_accumulate once per window in list order (each output voxel is owned by one thread, which applies the
 * covering windows in order). */
int pcb_sw_accumulate_batch(const void* pred, const void* map, void* value, void* weight, int dtype, int64_t Cout,
                            const int64_t roi[3], const int64_t out_size[3], const int64_t* starts, int64_t n,
                            void* stream);
/* window.py:275-294 normalize_weighted_accumulator: value /= clamp_min(weight, 1e-4) (in place). */
int pcb_sw_normalize(void* value, const void* weight, int dtype, int64_t Cout, int64_t nvox, void* stream);

/* ------------------------------------------------------------------ test-time augmentation
 * connectomics/inference/tta.py:706-714: out = rot90(flip(x, flip axes), k, (rot_a, rot_b)) for x [planes, in_size],
 * spatial axes 0..2, flip_mask bit a = flip axis a, rot_a < 0 = no rotation; odd k swaps the plane's dims in `out`. */
int pcb_tta_view(const void* x, void* out, int dtype, int64_t planes, const int64_t in_size[3], int flip_mask,
                 int rot_a, int rot_b, int k, void* stream);
/* One view folded into the ensemble accumulator acc [N, Cacc, acc_size]: the view is inverted by index map
 * (tta_affinity.py:364-369 invert_view: rot90(-k) then flip), channel c reads prediction channel src_channel[c]
 * (tta.py:404-413 select_channel), the activation act[c] (0 none, 1 sigmoid, 2 sigmoid(act_scale*x), 3 tanh; tta.py:312-402)
 * is applied in the prediction dtype, the value is cast to the accumulator dtype and folded with mode[c] (0 mean:
 * cur += (inc - cur) / (n_prev + 1), 1 min, 2 max; n_prev == 0 copies) — tta_ensemble.py:94-110.  pred is
 * [N, Cpred, view-frame size]; the arrays are HOST arrays of Cacc entries (Cacc <= 64). */
int pcb_tta_fold(const void* pred, int pred_dtype, void* acc, int acc_dtype, int64_t N, int64_t Cpred, int64_t Cacc,
                 const int64_t acc_size[3], int flip_mask, int rot_a, int rot_b, int k, const int* src_channel,
                 const int* mode, const int* act, const float* act_scale, int n_prev, void* stream);

/* pcb_tta_fold with the rest of the reference's view inversion and aggregation (same arrays, plus):
 *  - shift[Cacc*3] (or NULL): affinity-aware inversion, tta_affinity.py:376-391 — accumulator channel c reads
 *    spatial[src_channel[c]] displaced by the roll shift, canonical[p] = spatial[p - shift]; voxels whose source falls
 *    outside are zero and INVALID (tta_affinity.py:101-117 valid_slices_for_shift);
 *  - act[c] == 4: softmax over the canonical channels listed in sm_src / sm_shift [sm_off[c], sm_off[c]+sm_len[c])
 *    (prediction channel + shift of every member; tta.py:368-370);
 *  - part[c] >= 0: channel c is a PARTIAL channel (tta_ensemble.py:121-162): its value goes to stats[N, Cpart, vol] (fp32:
 *    sum / min / max per part_mode[c]) and counts[N, Cpart, vol] (count_dtype 0 = uint8, 1 = int16) only where it is valid
 *    (inside the shift's box and, when valid_mask [Cpart, vol] is given, where the mask byte is non-zero); mode[c] is ignored;
 *  - mean_as_sum: "mean" channels accumulate the plain sum (distributed view sharding, tta_ensemble.py:92-93). */
int pcb_tta_fold_ex(const void* pred, int pred_dtype, void* acc, int acc_dtype, int64_t N, int64_t Cpred, int64_t Cacc,
                    const int64_t acc_size[3], int flip_mask, int rot_a, int rot_b, int k, const int* src_channel,
                    const int* mode, const int* act, const float* act_scale, const int* shift, const int* sm_off,
                    const int* sm_len, const int* sm_src, const int* sm_shift, int n_members, const int* part,
                    const int* part_mode, int64_t Cpart, float* stats, void* counts, int count_dtype, const void* valid_mask,
                    int mean_as_sum, int n_prev, void* stream);
/* tta_affinity.py:350-393 invert_view as one gather: out [N, Cout, out_size] (same dtype as pred) = rot90(-k) then flip of
 * pred [N, Cpred, view frame], channel c taken from src_channel[c] (NULL: identity) and displaced by shift[c*3..] (NULL: none);
 * wrapped faces are zero. */
int pcb_tta_unview(const void* pred, void* out, int dtype, int64_t N, int64_t Cpred, int64_t Cout, const int64_t out_size[3],
                   int flip_mask, int rot_a, int rot_b, int k, const int* src_channel, const int* shift, void* stream);
/* tta_ensemble.py:187-211 finalize for the partial channels: acc[:, part_channel[j]] = stats[:, j] (/ counts for mode 0 = mean)
 * cast to the accumulator dtype.  *first_zero (device uint64, initialised by the caller to UINT64_MAX) receives the smallest flat
 * index of stats with zero coverage — the caller raises the reference's RuntimeError when it moved. */
int pcb_tta_finalize_partial(const float* stats, const void* counts, int count_dtype, void* acc, int acc_dtype, int64_t N,
                             int64_t Cacc, int64_t Cpart, const int* part_channel, const int* part_mode, int64_t nvox,
                             void* first_zero, void* stream);

/* ------------------------------------------------------------------ MedNeXt forward ops
 * (upstream nnunet_mednext blocks.py / MedNextV1.py as built by
 *  connectomics/models/architectures/mednext_models.py:374-380,479)
 *
 * stem: Conv3d(Cin, C, k=1).  x NCDHW in `in_dtype`  ->  out NDHWC bf16. w [C,Cin] f32, b [C] f32. */
int pcb_stem_fwd(const void* x, int in_dtype, const float* w, const float* b, void* out,
                 int64_t N, int64_t Cin, int64_t C, int64_t nvox, void* stream);
/* conv1 of a block: depthwise k^3 conv (SAME: stride 1 pad k/2; DOWN: stride 2 pad k/2;
 * UP: ConvTranspose stride 2 pad k/2 -> spatial 2s-1), + bias, bf16 output, and the per-(n,c)
 * sum / sum-of-squares of the rounded output for GroupNorm(num_groups=C) — accumulated into
 * `stats` [N, 2, C] float64 (sums, then sums of squares) which the caller zeroes.  w is [k^3, C] f32 (tap-major), b [C] f32. */
int pcb_dwconv_fwd(const void* x, const float* w, const float* b, void* y, double* stats,
                   int64_t N, const int64_t in_size[3], int64_t C, int k, int mode, void* stream);
/* norm -> conv2 (1x1, C->H) -> GELU -> conv3 (1x1, H->Co) [+ residual | + res_conv(x)] fused on
 * tcgen05: two back-to-back GEMMs per 128-voxel tile, the expanded tensor never leaves the SM.
 *   y      [N, Vy, C]  bf16   dw-conv output;   stats [N,2,C] f64 (sum, sumsq over Vy voxels)
 *   gamma,beta [C] f32 (GroupNorm affine, eps 1e-5);  w2 [H,C] bf16, b2 [H] f32; w3 [Co,H] bf16, b3 [Co] f32
 *   mode SAME: out[v] = mlp(y[v]) + (res ? res[v] : 0)                        (do_res)
 *   mode DOWN: out[v] = mlp(y[v]) + (wr ? wr*xs[2v] + br : 0)                 (res_conv stride 2)
 *   mode UP  : out over (2s)^3; o with any coord 0 -> res[o] (skip) only; else
 *              mlp(y[o-1]) + (wr ? (o-1 all even ? wr*xs[(o-1)/2] : 0) + br : 0) + (res ? res[o] : 0)
 *   xs [N, Vin, Cr] bf16 with spatial size xs_size, wr [Co,Cr] bf16, br [Co] f32 ; out [N, Vout, Co] bf16. */
int pcb_mlp_fwd(const void* y, const double* stats, const float* gamma, const float* beta,
                const void* w2, const float* b2, const void* w3, const float* b3,
                const void* res, const void* xs, const void* wr, const float* br, void* out,
                int64_t N, const int64_t out_size[3], const int64_t xs_size[3], int64_t C, int64_t H,
                int64_t Co, int64_t Cr, int mode, void* stream);
/* OutBlock ConvTranspose3d(C, ncls, k=1): x NDHWC bf16 -> out NCDHW in `out_dtype`.
 * w [C, ncls] f32 (ConvTranspose layout), b [ncls] f32. */
int pcb_head_fwd(const void* x, const float* w, const float* b, void* out, int out_dtype,
                 int64_t N, int64_t C, int64_t ncls, int64_t nvox, void* stream);

/* gradient of pcb_dwconv_fwd w.r.t. its input, with the forward stencil kernel in the dual mode
 * (SAME<-SAME with flipped taps supplied by the caller, DOWN<-UP, UP<-DOWN); optional fused add:
 * add_mode 1: dx[o] += add[o]; add_mode 2: dx[o] += add[o/2] where every coordinate of o is even. */
int pcb_dwconv_bwd_data(const void* dy, const float* w, const void* add, int add_mode, void* dx, int64_t N,
                        const int64_t dy_size[3], const int64_t dx_size[3], int64_t C, int k, int fwd_mode,
                        void* stream);

/* ------------------------------------------------------------------ MedNeXt backward ops (training:
 * what autograd + cuDNN do for the reference under training/lightning/model.py:863-910)
 *
 * fused dgrad of pcb_mlp_fwd on tcgen05: recomputes Hpre = norm(y) W2^T, dG = dOut W3, writes
 * Hact = GELU(Hpre+b2) and dh = dG*GELU'(.) (bf16 [N,Vy,H]), dYhat = dh W2 (bf16 [N,Vy,C]) and the
 * GroupNorm-backward sums gstats [N,2,C] f64 (sum g, sum g*xhat; caller zeroes).
 * w3t = conv3.weight^T [H,Co] bf16, w2t = conv2.weight^T [C,H] bf16.  y_size = spatial size of y;
 * mode UP reads dOut at o = p+1 (dOut spatial = y_size+1). */
int pcb_mlp_bwd(const void* y, const double* stats, const float* gamma, const float* beta, const void* w2,
                const float* b2, const void* w3t, const void* w2t, const void* dout, void* hact, void* dh,
                void* dyhat, double* gstats, int64_t N, const int64_t y_size[3], int64_t C, int64_t H,
                int64_t Co, int mode, void* stream);
/* Deep levels (C >= 128): conv2 -> GELU -> conv3 as two launches of one warp-specialised tcgen05 GEMM
 * (csrc/deep_mlp.cu) — the expanded activation goes through `hact` ([N][Vout][H] bf16, caller-owned, kept for
 * the backward pass) instead of shared memory, so the work splits over output columns as well as rows.
 * Same argument meaning as pcb_mlp_fwd.  pcb_mlp_fwd_deep_workspace returns the bytes `hact` needs, or 0 when
 * the shape is not served by this path (use pcb_mlp_fwd).  Replaces the same reference modules as pcb_mlp_fwd
 * (MedNeXtBlock.conv2/act/conv3 + res_conv, built at mednext_models.py:374-380). */
int64_t pcb_mlp_fwd_deep_workspace(int64_t N, const int64_t out_size[3], int64_t C, int64_t H, int64_t Co, int64_t Cr);
int pcb_mlp_fwd_deep(const void* y, const double* stats, const float* gamma, const float* beta, const void* w2,
                     const float* b2, const void* w3, const float* b3, const void* res, const void* xs,
                     const void* wr, const float* br, void* out, void* hact, int64_t N, const int64_t out_size[3],
                     const int64_t xs_size[3], int64_t C, int64_t H, int64_t Co, int64_t Cr, int mode, void* stream);

/* levels 0/1 (3H+C+Co <= 512): persistent kernel doing pcb_mlp_bwd AND both pointwise weight gradients with
 * Hact/dh kept on chip (TMEM-resident wgrad accumulators, bias grads via an all-ones MN-major row).
 * dW3 [Co,H], db3 [Co], dW2 [H,C], db2 [H] fp32 are overwritten; workspace from ..._workspace_floats. */
int pcb_mlp_bwd_fused_supported(int64_t C, int64_t H, int64_t Co, int64_t N, const int64_t y_size[3], int mode);
int64_t pcb_mlp_bwd_fused_workspace_floats(int64_t C, int64_t H, int64_t Co, int64_t N, const int64_t y_size[3]);
int pcb_mlp_bwd_fused(const void* y, const double* stats, const float* gamma, const float* beta, const void* w2,
                      const float* b2, const void* w3t, const void* w2t, const void* dout, void* dyhat,
                      double* gstats, float* workspace, float* dW3, float* db3, float* dW2, float* db2, int64_t N,
                      const int64_t y_size[3], int64_t C, int64_t H, int64_t Co, int mode, void* stream);
/* weight gradients: dW[m*ldm + n*ldn] = sum_{sample, v in box} A[mapA(v), m] * B[mapB(v), n]
 * (m < Ma, n < Nb), db[m] = sum A[mapA(v), m] when `ones`; split-K persistent tcgen05 GEMM with
 * MN-major operands + deterministic second-stage reduction over `workspace`
 * (pcb_tn_workspace_floats floats).  Row maps: 0 identity, 1 v+1 per axis, 2 2v, 3 2v+1 (into a
 * tensor of spatial size a_size / b_size).  stats/gamma/beta != NULL applies GroupNorm to B. */
int64_t pcb_tn_workspace_floats(int64_t Ma, int64_t Nb, int ones, int64_t N, const int64_t box[3]);
int pcb_tn_gemm(const void* A, const void* B, const double* stats, const float* gamma, const float* beta,
                float* workspace, float* dW, int64_t ldm, int64_t ldn, float* db, int64_t N,
                const int64_t box[3], int mapA, const int64_t a_size[3], int64_t a_cols, int64_t Ma,
                int mapB, const int64_t b_size[3], int64_t Nb, int ones, void* stream);
/* row-gather 1x1 conv: out[n, r, :] = A[n, map(r), :K] W[Nw,K]^T (+bias), bf16 in/out (res-conv data
 * gradients; MedNeXtTaskHead.input_projection, mednext_models.py:169-173). */
int pcb_pw_fwd(const void* A, const void* W, const float* bias, void* out, int64_t N, const int64_t out_box[3],
               int map, const int64_t a_size[3], int64_t K, int64_t Nw, void* stream);
/* GroupNorm(num_groups=C) backward: dy = rstd*gamma*(g - S1/V - xhat*S2/V) (bf16) and dsum[c] += sum dy. */
int pcb_gn_bwd(const void* g, const void* y, const double* stats, const double* gstats, const float* gamma,
               void* dy, double* dsum, int64_t N, int64_t C, int64_t V, void* stream);
/* depthwise weight gradient dW[tap, c] += sum_v center[v,c] * neigh[stride*v - k/2 + tap, c]  (f64 [k^3, C]). */
int pcb_dwconv_wgrad(const void* center, const void* neigh, double* dW, int64_t N, const int64_t c_size[3],
                     const int64_t n_size[3], int64_t C, int k, int stride, void* stream);
/* OutBlock backward: dX (NDHWC bf16), dW [C,ncls] f64 +=, db [ncls] f64 += ; dout NCDHW in `dtype`. */
int pcb_head_bwd(const void* dout, int dtype, const void* x, const float* w, void* dx, double* dW, double* db,
                 int64_t N, int64_t C, int64_t ncls, int64_t nvox, void* stream);
/* stem backward: dW [C,Cin] f64 +=, db [C] f64 += from g (NDHWC bf16) and the NCDHW input. */
int pcb_stem_bwd(const void* g, const void* x, int in_dtype, double* dW, double* db, int64_t N, int64_t Cin,
                 int64_t C, int64_t nvox, void* stream);

/* channels-first LayerNorm of the MedNeXt blocks, cfg.model.mednext.norm = "layer" (upstream nnunet_mednext
 * blocks.py::LayerNorm(data_format="channels_first"), selected at mednext_models.py:449-476): per row (voxel) of a
 * channels-last bf16 [rows, C] tensor, out = (y - mean_c) / sqrt(var_c + 1e-5) * weight + bias.  C/8 must be a power of
 * two (PCB_ERR_UNSUPPORTED otherwise).  The normalised tensor feeds pcb_mlp_fwd with identity GroupNorm constants. */
int pcb_layernorm_fwd(const void* y, const float* weight, const float* bias, void* out, int64_t C, int64_t rows,
                      void* stream);
/* its backward: dy = rstd*(g*w - mean_c(g*w) - xhat*mean_c(g*w*xhat)) (bf16); dweight[c] += sum g*xhat, dbias[c] += sum g
 * (f64, caller zeroes). */
int pcb_layernorm_bwd(const void* g, const void* y, const float* weight, void* dy, double* dweight, double* dbias,
                      int64_t C, int64_t rows, void* stream);

/* ------------------------------------------------------------------ dense conv path (MONAI UNet, `monai_unet`)
 * (monai.networks.blocks.{Convolution,ResidualUnit,ADN} as built by
 *  connectomics/models/architectures/monai_models.py:235-248)
 *
 * Conv3d (transposed=0: src = o*stride + t - pad) / ConvTranspose3d (transposed=1: src = (o + pad - t)/stride) as an
 * implicit GEMM on tcgen05.  x NDHWC bf16 [N,in_size,Ci], w bf16 [k^3][Co][Ci] (tap-major, K-major rows),
 * bias f32 [Co] or NULL, out NDHWC bf16 [N,out_size,Co].  Ci, Co padded to multiples of 16.  The same entry
 * point computes data gradients with repacked weights (conv <-> transposed conv). */
int pcb_conv_fwd(const void* x, const void* w, const float* bias, void* out, int64_t N, const int64_t in_size[3],
                 const int64_t out_size[3], int64_t Ci, int64_t Co, int k, int stride, int pad, int transposed,
                 void* stream);
/* weight gradient of one tap: dW[co*ldm + ci*ldn] = sum_o dy[o,co] * x[src(o,tap),ci]; db[co] = sum dy (optional). */
int pcb_conv_wgrad_tap(const void* dy, const void* x, float* workspace, float* dW, int64_t ldm, int64_t ldn, float* db,
                       int64_t N, const int64_t out_size[3], const int64_t in_size[3], int64_t Co, int64_t Ci,
                       const int tap[3], int stride, int pad, int transposed, void* stream);
/* per-channel sum / sum of squares over `rows` channels-last bf16 rows (BatchNorm batch statistics), f64 [2,C] +=. */
int pcb_channel_stats(const void* x, double* stats, int64_t C, int64_t rows, void* stream);
/* ADN "NA": y = PReLU(x*scale + shift) with BatchNorm folded into the per-channel affine; slope = device float. */
int pcb_bn_act_fwd(const void* x, const float* scale, const float* shift, const float* slope, void* out, int64_t C,
                   int64_t rows, void* stream);
/* dz = dy*PReLU'(z) (bf16); red[0..C) += dz, red[C..2C) += dz*xhat, red[2C] += dy*min(z,0)  (f64, caller zeroes). */
int pcb_bn_act_bwd(const void* dy, const void* x, const float* scale, const float* shift, const float* mean,
                   const float* rstd, const float* slope, void* dz, double* red, int64_t C, int64_t rows, void* stream);
/* BatchNorm backward with batch statistics (pooled sums replicated per sample in stats/gstats [N,2,C]). */
int pcb_bn_bwd(const void* g, const void* x, const double* stats, const double* gstats, const float* gamma,
               void* dx, double* dsum, int64_t N, int64_t C, int64_t V, void* stream);

/* ------------------------------------------------------------------ network- and plan-level entry points (inference)
 * A pcb_net is the forward plan of a MedNeXt — nnunet_mednext MedNextV1.py::MedNeXt.forward as built by
 * connectomics/models/architectures/mednext_models.py:374-380: stem -> blocks (SAME / DOWN / UP with encoder skips) ->
 * OutBlock heads.  The library owns the plan; the CALLER owns the weights (device pointers in kernel layout, below) and
 * the activation arena, and keeps both alive while the plan is used.  One pcb_net_forward call enqueues the whole
 * network on the caller's stream (no host work between kernels, CUDA-graph capturable). */
typedef struct pcb_net pcb_net;
typedef struct {
  int32_t kind;       /* pcb_dw_mode of conv1 */
  int32_t C, H, Co;   /* channels in, hidden (exp_r*C), out */
  int32_t k;          /* depthwise kernel size 3/5/7 */
  int32_t do_res;     /* SAME blocks: out += x */
  int32_t norm;       /* 0 = GroupNorm(num_groups=C) (the only kind the native plan runs) */
  int32_t skip_from;  /* UP blocks: index of the block whose OUTPUT is added as the encoder skip, -1 = none */
  const float* w1;    /* conv1.weight as [k^3, C] f32 (tap-major) */
  const float* b1;    /* conv1.bias [C] */
  const float* gamma; /* norm.weight [C] */
  const float* beta;  /* norm.bias [C] */
  const void* w2;     /* conv2.weight as bf16 [H, C] */
  const float* b2;    /* [H] */
  const void* w3;     /* conv3.weight as bf16 [Co, H] */
  const float* b3;    /* [Co] */
  const void* wr;     /* res_conv.weight as bf16 [Co, C] (DOWN / UP with do_res_up_down) or NULL */
  const float* br;    /* res_conv.bias [Co] or NULL */
} pcb_block_desc;
typedef struct {
  int32_t from_block; /* index of the block whose output feeds this OutBlock; -1 = the stem output */
  int32_t ncls;
  const float* w;     /* conv_out.weight as [C, ncls] f32 (ConvTranspose layout) */
  const float* b;     /* [ncls] */
} pcb_head_desc;
int pcb_net_create(int32_t in_channels, int32_t n_channels, const float* stem_w, const float* stem_b,
                   const pcb_block_desc* blocks, int32_t nblocks, const pcb_head_desc* heads, int32_t nheads, pcb_net** out);
void pcb_net_destroy(pcb_net* net);
int32_t pcb_net_in_channels(const pcb_net* net);
int32_t pcb_net_num_heads(const pcb_net* net);
int32_t pcb_net_head_channels(const pcb_net* net, int32_t head);
/* bytes of activation arena one forward of N samples of spatial `size` needs (liveness-planned: the peak of the live
 * set, not the sum over layers); -1 on error. */
int64_t pcb_net_workspace_bytes(const pcb_net* net, int64_t N, const int64_t size[3]);
/* x NCDHW [N, Cin, size] in `in_dtype` -> outs[h] NCDHW [N, ncls_h, size_h] in `out_dtype` for every head whose
 * outs[h] != NULL (deep-supervision heads read coarser levels: size_h = size / 2^level).  workspace: 256-byte aligned. */
int pcb_net_forward(pcb_net* net, const void* x, int in_dtype, int64_t N, const int64_t size[3], void* const* outs,
                    int out_dtype, void* workspace, int64_t ws_bytes, void* stream);
/* The sliding-window tile loop of connectomics/inference/window.py:563-683 for a pcb_net, entirely in the library:
 * for each batch of `sw_batch` (<= 16) windows of the HOST list `starts` (n (z,y,x) triples, all inside acc_size):
 * crop + pad from vol [1, Cin, image] -> pcb_net_forward -> value[:, box] += pred*map, weight[box] += map in list order
 * (pcb_sw_accumulate_batch semantics: bit-identical to the sequential loop).  value [ncls, acc_size] / weight [acc_size]
 * in `acc_dtype` are NOT zeroed and NOT normalised here (callers carry partial planes across slabs / ranks and call
 * pcb_sw_normalize).  use_graph: the batch body is captured once into a CUDA graph (window starts come from a device
 * table through a device cursor) and replayed — one graph launch per batch.  map: pcb_sw_importance_map in acc_dtype. */
int64_t pcb_sw_run_workspace_bytes(const pcb_net* net, int head, const int64_t roi[3], int sw_batch, int64_t nstarts,
                                   int vol_dtype, int acc_dtype);
int pcb_sw_run(pcb_net* net, int head, const void* vol, int vol_dtype, const int64_t image[3], const int64_t roi[3],
               int pad_mode, double cval, int sw_batch, const int64_t* starts, int64_t nstarts, const void* map,
               void* value, void* weight, int acc_dtype, const int64_t acc_size[3], void* workspace, int64_t ws_bytes,
               int use_graph, void* stream);

/* ------------------------------------------------------------------ optimizer-side fusion over the flat arenas
 * sum of squares of a gradient arena (f64 device scalar, caller zeroes) — torch.nn.utils.clip_grad_norm_'s total norm
 * (Lightning gradient_clip_val, tutorials: 1.0). */
int pcb_grad_sumsq(const float* grad, int64_t n, double* out, void* stream);
/* ONE pass: [clip by global norm] + AdamW + [EMA] over flat fp32 arenas of n elements.  Segment i covers elements
 * [seg_end[i-1], seg_end[i]) with its own lr / weight decay (training/optimization/build.py:88-113: norm layers and
 * biases form their own groups); seg_* are DEVICE arrays.  grad is used as grad*grad_scale (1/world: the all-reduced SUM
 * becomes the mean) * clip coefficient min(1, max_norm / (sqrt(*grad_sumsq)*grad_scale + 1e-6)) when grad_sumsq != NULL
 * and max_norm > 0.  `step` is a device float holding the number of steps taken so far; it is incremented on the stream
 * (CUDA-graph capturable, like torch's capturable=True).  ema (or NULL): ema = ema*decay + param*(1-decay) after the
 * update (training/lightning/callbacks.py:869-907).  seg_active (device scratch, nseg int32, or NULL): segments whose
 * gradient slice is entirely zero are skipped, as torch.optim skips parameters whose .grad is None (what DDP with
 * find_unused_parameters leaves for unused heads).  Update formula = torch.optim.AdamW, op for op. */
int pcb_adamw_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, float* ema, int64_t n,
                   const int64_t* seg_end, const float* seg_lr, const float* seg_wd, int nseg, float beta1, float beta2,
                   float eps, float* step, const double* grad_sumsq, float max_norm, float grad_scale, float ema_decay,
                   int32_t* seg_active, void* stream);

/* ------------------------------------------------------------------ exchange steps
 * The two places where the path exchanges data between ranks, for hosts that do not have torch.distributed:
 *   - the gradient mean of the data-parallel step — Lightning DDPStrategy, connectomics/training/lightning/trainer.py:231-256
 *     (the Python package does the same exchange through torch.distributed over the flat arena, training/ddp.py);
 *   - the overlap planes between z-neighbours of the sharded sliding-window engine (inference/sharded.py), which replaces the
 *     full-accumulator reduce-to-root of connectomics/inference/lazy_distributed.py:78-169.
 * NCCL is bound with dlopen at the first call (the copy already loaded in the process wins); no NCCL -> PCB_ERR_UNSUPPORTED.
 * A pcb_comm belongs to the CUDA device that was current in pcb_comm_init; one communicator per rank process. */
#define PCB_COMM_ID_BYTES 128
typedef struct pcb_comm pcb_comm;
/* rank 0 creates the id and hands the 128 bytes to every rank (MPI, a file, torch.distributed.broadcast_object_list ...) */
int pcb_comm_unique_id(void* id_out);
int pcb_comm_init(const void* unique_id, int rank, int world, pcb_comm** out);   /* collective over all `world` ranks */
void pcb_comm_destroy(pcb_comm* comm);
int pcb_comm_rank(const pcb_comm* comm);
int pcb_comm_world(const pcb_comm* comm);
int pcb_comm_nccl_version(void);          /* NCCL_VERSION_CODE of the bound library, -1 when none */
/* arena[0..numel) = scale * SUM over ranks (in place, on `stream`): scale = 1/world is DDP's gradient mean; scale = 1 leaves
 * the SUM (pcb_adamw_step folds 1/world into its grad_scale).  world == 1: only the scale is applied. */
int pcb_grad_allreduce(pcb_comm* comm, void* arena, int64_t numel, int dtype, float scale, void* stream);
/* One grouped ncclSend/ncclRecv exchange: message i of the send list goes to rank send_peer[i], message j of the receive
 * list comes from recv_peer[j] (a message = value planes followed by weight planes of one face, packed by the caller;
 * the caller adds what it received: own + neighbour, the association documented in inference/sharded.py). */
int pcb_sw_exchange_overlap(pcb_comm* comm, int nsend, const void* const* send_bufs, const int64_t* send_numel,
                            const int* send_peer, int nrecv, void* const* recv_bufs, const int64_t* recv_numel,
                            const int* recv_peer, int dtype, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PCB200_H */
