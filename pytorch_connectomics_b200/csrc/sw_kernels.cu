// pcb200 — sliding-window tiled inference: integer grid logic (host) and the crop/pad,
// overlap-add and normalise kernels (device).  Behaviour follows the reference's
// connectomics/inference/window.py and the grid part of inference/lazy.py (cited per function
// in include/pcb200.h).  All of this is HBM-bound streaming work: 128-bit-friendly contiguous
// x-rows, one pass per tensor, no atomics (windows are accumulated in launch order, so fp sums
// associate exactly like the reference's sequential `+=`).
#include <math.h>
#include <string.h>
#include <vector>

#include "../../include/pcb200.h"
#include "pcb_common.cuh"

namespace pcb {

// ---- dtype-generic "compute in fp32, round to T after every op" arithmetic (torch CPU/CUDA
//      elementwise semantics for f16/bf16)
template <typename T> struct DT;
template <> struct DT<float> {
  __host__ __device__ static float ld(float v) { return v; }
  __host__ __device__ static float st(float v) { return v; }
};
template <> struct DT<__half> {
  __host__ __device__ static float ld(__half v) { return __half2float(v); }
  __host__ __device__ static __half st(float v) { return __float2half_rn(v); }
};
template <> struct DT<__nv_bfloat16> {
  __host__ __device__ static float ld(__nv_bfloat16 v) { return __bfloat162float(v); }
  __host__ __device__ static __nv_bfloat16 st(float v) { return __float2bfloat16_rn(v); }
};
template <typename T> __host__ __device__ inline float rnd(float v) { return DT<T>::ld(DT<T>::st(v)); }

static inline double clamp01(double o) { return o < 0.0 ? 0.0 : (o > 0.99 ? 0.99 : o); }

// window.py:107-118
static void axis_starts_eager(int64_t img, int64_t roi, int64_t stride, std::vector<int64_t>& out) {
  out.clear();
  if (img <= roi) { out.push_back(0); return; }
  if (stride < 1) stride = 1;
  for (int64_t s = 0; s <= img - roi; s += stride) out.push_back(s);
  if (out.back() != img - roi) out.push_back(img - roi);
}
// lazy.py:269-286 (_snap_offsets with border_pad = roi - stride)
static void axis_starts_lazy(int64_t img, int64_t roi, int64_t stride, std::vector<int64_t>& out) {
  out.clear();
  if (img <= roi) { out.push_back(0); return; }
  if (stride < 1) stride = 1;
  int64_t bp = roi - stride; if (bp < 0) bp = 0;
  const int64_t lo = -bp, hi = img - roi + bp;
  for (int64_t s = lo; s <= hi; s += stride) out.push_back(s);
  if (out.empty() || out.back() != hi) out.push_back(hi);
}

template <typename T>
static void host_axis_kernel_bump(int64_t n, std::vector<float>& k) {
  // window.py:178-187 evaluated op-by-op in dtype T
  const float tiny = sizeof(T) == 4 ? 1.17549435e-38f : (std::is_same<T, __half>::value ? 6.103515625e-05f : 1.17549435e-38f);
  k.resize(n);
  float mx = -INFINITY;
  const float denom_n = (float)((double)n + 1.0);
  for (int64_t i = 0; i < n; ++i) {
    float idx = rnd<T>((float)i);
    float u = rnd<T>(idx + 1.0f);
    u = rnd<T>(u / denom_n);
    u = rnd<T>(u * 2.0f);
    u = rnd<T>(u - 1.0f);
    float d = rnd<T>(1.0f - rnd<T>(u * u));
    if (d < tiny) d = tiny;
    float e = rnd<T>(expf(rnd<T>(-1.0f / d)));
    k[i] = e;
    if (e > mx) mx = e;
  }
  if (mx < tiny) mx = tiny;
  for (int64_t i = 0; i < n; ++i) k[i] = rnd<T>(k[i] / mx);
}

template <typename T>
__global__ void imap_kernel(T* __restrict__ out, const float* __restrict__ k0, const float* __restrict__ k1,
                            const float* __restrict__ k2, int64_t n0, int64_t n1, int64_t n2, int blend,
                            float tiny, float min_value) {
  const int64_t total = n0 * n1 * n2;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t x = i % n2, y = (i / n2) % n1, z = i / (n1 * n2);
    float v;
    if (blend == PCB_BLEND_CONSTANT) {
      v = 1.0f;
      if (min_value > 0.f) v = fmaxf(v, min_value);
    } else if (blend == PCB_BLEND_BUMP) {
      v = rnd<T>(rnd<T>(k0[z] * k1[y]) * k2[x]);
      v = fmaxf(v, tiny);
      if (min_value > 0.f) v = fmaxf(v, min_value);
    } else {  // distance transform: min over axes of min(i+1, n-i)  (window.py:234-243)
      v = fminf(fminf(k0[z], k1[y]), k2[x]);
    }
    out[i] = DT<T>::st(v);
  }
}

struct ExtractArgs {
  int64_t C, D, H, W, r0, r1, r2;
  float cval;
  int mode;
};

__device__ __forceinline__ int64_t pad_index(int64_t p, int64_t lo, int64_t hi, int mode) {
  // coordinate p outside the in-image crop [lo,hi) -> source index (relative to the crop, like
  // F.pad applied to the cropped tensor, window.py:492-522)
  if (p >= lo && p < hi) return p;
  if (mode == PCB_PAD_REFLECT) return p < lo ? lo + (lo - p) : (hi - 1) - (p - (hi - 1));
  if (mode == PCB_PAD_REPLICATE) return p < lo ? lo : hi - 1;
  if (mode == PCB_PAD_CIRCULAR) { int64_t n = hi - lo; int64_t m = (p - lo) % n; if (m < 0) m += n; return lo + m; }
  return -1;
}

// Up to SW_MAXB windows per launch.  Window starts come either from the kernel parameters (host-known batch) or from a
// device-resident table indexed through a device cursor (CUDA-graph replay: the same graph serves every batch; the
// cursor is advanced by advance_cursor_kernel at the end of the captured body).  A start with z <= SW_SKIP marks a
// padding slot of the last, partial batch.
constexpr int SW_MAXB = 16;
constexpr int64_t SW_SKIP = -(1ll << 40);
struct WinList {
  int64_t s[SW_MAXB][3];
  const int64_t* table;      // device table of (z,y,x) starts, or nullptr
  const int64_t* cursor;     // device scalar: index of the batch's first window in `table`
  int64_t total;             // windows in `table`
  int n;
};
__device__ __forceinline__ bool win_start(const WinList& wl, int w, int64_t& z, int64_t& y, int64_t& x) {
  if (wl.table == nullptr) { z = wl.s[w][0]; y = wl.s[w][1]; x = wl.s[w][2]; return z > SW_SKIP; }
  const int64_t i = wl.cursor[0] + w;
  if (i >= wl.total) return false;
  z = wl.table[3 * i]; y = wl.table[3 * i + 1]; x = wl.table[3 * i + 2];
  return z > SW_SKIP;
}

template <typename T>
__global__ void extract_kernel(const T* __restrict__ vol, T* __restrict__ out, WinList wl, ExtractArgs a) {
  const int64_t w = blockIdx.y;  // window in batch
  int64_t s0, s1, s2;
  if (!win_start(wl, (int)w, s0, s1, s2)) return;
  const int64_t lo0 = max((int64_t)0, s0), hi0 = min(a.D, s0 + a.r0);
  const int64_t lo1 = max((int64_t)0, s1), hi1 = min(a.H, s1 + a.r1);
  const int64_t lo2 = max((int64_t)0, s2), hi2 = min(a.W, s2 + a.r2);
  // per-window fallback to constant when a reflect/circular pad >= the cropped dim (window.py:509-518)
  int mode = a.mode;
  if (mode == PCB_PAD_REFLECT || mode == PCB_PAD_CIRCULAR) {
    const int64_t st[3] = {s0, s1, s2}, rr[3] = {a.r0, a.r1, a.r2}, im[3] = {a.D, a.H, a.W};
    const int64_t lo[3] = {lo0, lo1, lo2}, hi[3] = {hi0, hi1, hi2};
#pragma unroll
    for (int ax = 0; ax < 3; ++ax) {
      const int64_t before = st[ax] < 0 ? -st[ax] : 0, e = st[ax] + rr[ax], after = e > im[ax] ? e - im[ax] : 0;
      if (before >= hi[ax] - lo[ax] || after >= hi[ax] - lo[ax]) mode = PCB_PAD_CONSTANT;
    }
  }
  const int64_t per = a.C * a.r0 * a.r1 * a.r2;
  const T cv = DT<T>::st(a.cval);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < per; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t x = i % a.r2, y = (i / a.r2) % a.r1, z = (i / (a.r2 * a.r1)) % a.r0, c = i / (a.r2 * a.r1 * a.r0);
    const int64_t pz = pad_index(s0 + z, lo0, hi0, mode), py = pad_index(s1 + y, lo1, hi1, mode),
                  px = pad_index(s2 + x, lo2, hi2, mode);
    T v = cv;
    if (pz >= 0 && py >= 0 && px >= 0) v = vol[((c * a.D + pz) * a.H + py) * a.W + px];
    out[w * per + i] = v;
  }
}

__global__ void advance_cursor_kernel(int64_t* cursor, int64_t by) {
  if (threadIdx.x == 0 && blockIdx.x == 0) cursor[0] += by;
}

struct AccArgs {
  int64_t Cout, r0, r1, r2, o0, o1, o2, p0, p1, p2, q0, q1, q2, b0, b1, b2;
};

template <typename T>
__global__ void accumulate_kernel(const T* __restrict__ pred, const T* __restrict__ map, T* __restrict__ value,
                                  T* __restrict__ weight, AccArgs a) {
  const int64_t nbox = a.b0 * a.b1 * a.b2;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nbox; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t x = i % a.b2, y = (i / a.b2) % a.b1, z = i / (a.b2 * a.b1);
    const int64_t pi = ((a.p0 + z) * a.r1 + (a.p1 + y)) * a.r2 + (a.p2 + x);   // index in the window
    const int64_t oi = ((a.q0 + z) * a.o1 + (a.q1 + y)) * a.o2 + (a.q2 + x);   // index in the output
    const float w = DT<T>::ld(map[pi]);
    weight[oi] = DT<T>::st(__fadd_rn(DT<T>::ld(weight[oi]), w));
    const int64_t pstride = a.r0 * a.r1 * a.r2, ostride = a.o0 * a.o1 * a.o2;
    for (int64_t c = 0; c < a.Cout; ++c) {
      const float prod = rnd<T>(__fmul_rn(DT<T>::ld(pred[c * pstride + pi]), w));
      value[c * ostride + oi] = DT<T>::st(__fadd_rn(DT<T>::ld(value[c * ostride + oi]), prod));
    }
  }
}

// A batch of FULL windows in one launch, bit-identical to accumulating them one after the other in list order: the
// thread of (window w, voxel i) owns output voxel o = start_w + i for the WHOLE batch iff w is the first window of the
// batch that covers o; it then applies every covering window w' >= w in order (value += pred*map, weight += map; mul then
// add, no FMA).  Each output voxel is touched by exactly one thread: no atomics, no race, the reference's association.
struct AccBatchArgs {
  int64_t Cout, r0, r1, r2, o0, o1, o2;
};
template <typename T>
__global__ void accumulate_batch_kernel(const T* __restrict__ pred, const T* __restrict__ map, T* __restrict__ value,
                                        T* __restrict__ weight, WinList wl, AccBatchArgs a) {
  __shared__ int64_t sst[SW_MAXB][3];
  __shared__ int sok[SW_MAXB];
  if (threadIdx.x < wl.n) {
    int64_t z, y, x;
    const bool ok = win_start(wl, threadIdx.x, z, y, x);
    sok[threadIdx.x] = ok ? 1 : 0;
    sst[threadIdx.x][0] = z; sst[threadIdx.x][1] = y; sst[threadIdx.x][2] = x;
  }
  __syncthreads();
  const int w = blockIdx.y;
  if (!sok[w]) return;
  const int64_t per = a.r0 * a.r1 * a.r2, ostride = a.o0 * a.o1 * a.o2;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < per; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t x = i % a.r2, y = (i / a.r2) % a.r1, z = i / (a.r2 * a.r1);
    const int64_t oz = sst[w][0] + z, oy = sst[w][1] + y, ox = sst[w][2] + x;
    if (oz < 0 || oz >= a.o0 || oy < 0 || oy >= a.o1 || ox < 0 || ox >= a.o2) continue;
    bool first = true;
    for (int u = 0; u < w; ++u) {
      if (!sok[u]) continue;
      const int64_t dz = oz - sst[u][0], dy = oy - sst[u][1], dx = ox - sst[u][2];
      if (dz >= 0 && dz < a.r0 && dy >= 0 && dy < a.r1 && dx >= 0 && dx < a.r2) { first = false; break; }
    }
    if (!first) continue;
    const int64_t oi = (oz * a.o1 + oy) * a.o2 + ox;
    float wacc = DT<T>::ld(weight[oi]);
    for (int u = w; u < wl.n; ++u) {
      if (!sok[u]) continue;
      const int64_t dz = oz - sst[u][0], dy = oy - sst[u][1], dx = ox - sst[u][2];
      if (dz < 0 || dz >= a.r0 || dy < 0 || dy >= a.r1 || dx < 0 || dx >= a.r2) continue;
      const int64_t pi = (dz * a.r1 + dy) * a.r2 + dx;
      const float m = DT<T>::ld(map[pi]);
      wacc = rnd<T>(__fadd_rn(wacc, m));
      const T* pw = pred + (int64_t)u * a.Cout * per + pi;
      for (int64_t c = 0; c < a.Cout; ++c) {
        const float prod = rnd<T>(__fmul_rn(DT<T>::ld(pw[c * per]), m));
        value[c * ostride + oi] = DT<T>::st(__fadd_rn(DT<T>::ld(value[c * ostride + oi]), prod));
      }
    }
    weight[oi] = DT<T>::st(wacc);
  }
}

template <typename T>
__global__ void normalize_kernel(T* __restrict__ value, const T* __restrict__ weight, int64_t Cout, int64_t nvox,
                                 float floor_) {
  const int64_t total = Cout * nvox;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const float d = rnd<T>(fmaxf(DT<T>::ld(weight[i % nvox]), floor_));
    value[i] = DT<T>::st(__fdiv_rn(DT<T>::ld(value[i]), d));
  }
}

static inline int grid_for(int64_t n, int threads = 256) {
  int64_t b = (n + threads - 1) / threads;
  const int64_t cap = 148 * 16;  // a few waves of resident CTAs per SM; grid-stride covers the rest
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace pcb

using namespace pcb;

extern "C" int pcb_sw_scan_interval(const int64_t image[3], const int64_t roi[3], const double overlap[3],
                                    int64_t out[3]) {
  PCB_CHECK_ARG(image && roi && overlap && out, "pcb_sw_scan_interval: null argument");
  for (int a = 0; a < 3; ++a) {
    PCB_CHECK_ARG(roi[a] > 0, "roi_size must contain positive values");
    if (image[a] <= roi[a]) { out[a] = image[a]; continue; }
    // Python: max(1, int(round(roi * (1 - ov))))  — round() is round-half-even == nearbyint
    const double v = nearbyint((double)roi[a] * (1.0 - clamp01(overlap[a])));
    out[a] = v < 1.0 ? 1 : (int64_t)v;
  }
  return PCB_OK;
}

extern "C" int pcb_sw_plan(int grid_kind, const int64_t image[3], const int64_t roi[3], const double overlap[3],
                           const int64_t region[6], int64_t* starts_out, int64_t capacity, int64_t* count_out) {
  PCB_CHECK_ARG(image && roi && overlap && count_out, "pcb_sw_plan: null argument");
  PCB_CHECK_ARG(grid_kind >= PCB_GRID_EAGER && grid_kind <= PCB_GRID_LAZY_SNAP, "pcb_sw_plan: bad grid kind %d", grid_kind);
  std::vector<int64_t> ax[3];
  int64_t iv[3];
  if (grid_kind == PCB_GRID_LAZY_SNAP) {
    for (int a = 0; a < 3; ++a) {  // lazy.py:311-314: int(roi * (1 - overlap)) truncation, overlap NOT clamped
      const double v = (double)roi[a] * (1.0 - overlap[a]);
      iv[a] = (int64_t)v; if (iv[a] < 1) iv[a] = 1;
    }
  } else {
    int rc = pcb_sw_scan_interval(image, roi, overlap, iv);
    if (rc) return rc;
  }
  for (int a = 0; a < 3; ++a) {
    if (grid_kind == PCB_GRID_EAGER) axis_starts_eager(image[a], roi[a], iv[a], ax[a]);
    else axis_starts_lazy(image[a], roi[a], iv[a], ax[a]);
    if (region) {  // lazy.py:351-358
      std::vector<int64_t> keep;
      for (int64_t o : ax[a]) if (o < region[3 + a] && o + roi[a] > region[a]) keep.push_back(o);
      ax[a].swap(keep);
    }
  }
  const int64_t total = (int64_t)ax[0].size() * (int64_t)ax[1].size() * (int64_t)ax[2].size();
  *count_out = total;
  if (starts_out) {
    int64_t i = 0;
    for (int64_t z : ax[0]) for (int64_t y : ax[1]) for (int64_t x : ax[2]) {
      if (i >= capacity) return PCB_OK;
      starts_out[3 * i] = z; starts_out[3 * i + 1] = y; starts_out[3 * i + 2] = x; ++i;
    }
  }
  return PCB_OK;
}

template <typename T>
static int imap_impl(int blend, const int64_t roi[3], int ndim, double min_value, void* out, cudaStream_t st) {
  std::vector<float> k[3];
  const bool is_half = std::is_same<T, __half>::value;
  const float tiny = is_half ? 6.103515625e-05f : 1.17549435e-38f;
  for (int a = 0; a < 3; ++a) {
    if (blend == PCB_BLEND_BUMP) host_axis_kernel_bump<T>(roi[a], k[a]);
    else {
      k[a].resize(roi[a]);
      if (a < 3 - ndim) { for (auto& v : k[a]) v = INFINITY; continue; }  // padded leading axis
      for (int64_t i = 0; i < roi[a]; ++i) {
        const float c = rnd<T>((float)i);
        k[a][i] = fminf(rnd<T>(c + 1.0f), rnd<T>(rnd<T>((float)roi[a]) - c));
      }
    }
  }
  float* dk = nullptr;
  const int64_t tot = roi[0] + roi[1] + roi[2];
  if (cudaMallocAsync(&dk, tot * sizeof(float), st) != cudaSuccess) { set_error("imap: cudaMallocAsync failed"); return PCB_ERR_CUDA; }
  if (cudaMemcpyAsync(dk, k[0].data(), roi[0] * 4, cudaMemcpyHostToDevice, st) != cudaSuccess ||
      cudaMemcpyAsync(dk + roi[0], k[1].data(), roi[1] * 4, cudaMemcpyHostToDevice, st) != cudaSuccess ||
      cudaMemcpyAsync(dk + roi[0] + roi[1], k[2].data(), roi[2] * 4, cudaMemcpyHostToDevice, st) != cudaSuccess ||
      cudaStreamSynchronize(st) != cudaSuccess) {  // host staging vectors die at return (one-off setup call, not the tile loop)
    set_error("imap: staging the axis kernels failed: %s", cudaGetErrorString(cudaGetLastError()));
    cudaFreeAsync(dk, st);
    return PCB_ERR_CUDA;
  }
  const int64_t n = roi[0] * roi[1] * roi[2];
  const float mv = min_value > 0 ? rnd<T>((float)min_value) : 0.f;
  imap_kernel<T><<<grid_for(n), 256, 0, st>>>((T*)out, dk, dk + roi[0], dk + roi[0] + roi[1], roi[0], roi[1], roi[2],
                                               blend, tiny, mv);
  cudaFreeAsync(dk, st);
  PCB_CHECK_LAUNCH("pcb_sw_importance_map");
  return PCB_OK;
}

extern "C" int pcb_sw_importance_map(int blend, const int64_t roi[3], int ndim, int dtype, double min_value,
                                     void* map_out, void* stream) {
  PCB_CHECK_ARG(roi && map_out, "pcb_sw_importance_map: null argument");
  PCB_CHECK_ARG(roi[0] > 0 && roi[1] > 0 && roi[2] > 0, "roi_size must contain positive values");
  PCB_CHECK_ARG(blend >= 0 && blend <= 2, "unsupported blending mode %d", blend);
  PCB_CHECK_ARG(ndim >= 1 && ndim <= 3, "ndim must be 1..3");
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == PCB_F32) return imap_impl<float>(blend, roi, ndim, min_value, map_out, st);
  if (dtype == PCB_F16) return imap_impl<__half>(blend, roi, ndim, min_value, map_out, st);
  if (dtype == PCB_BF16) return imap_impl<__nv_bfloat16>(blend, roi, ndim, min_value, map_out, st);
  set_error("pcb_sw_importance_map: bad dtype %d", dtype);
  return PCB_ERR_INVALID;
}

static int extract_launch(const void* vol, int dtype, int64_t C, const int64_t image[3], const int64_t roi[3],
                          const WinList& wl, int pad_mode, double cval, void* out, cudaStream_t st) {
  ExtractArgs a{C, image[0], image[1], image[2], roi[0], roi[1], roi[2], (float)cval, pad_mode};
  const int64_t per = C * roi[0] * roi[1] * roi[2];
  dim3 grid(grid_for(per), (unsigned)wl.n);
  if (dtype == PCB_F32) extract_kernel<float><<<grid, 256, 0, st>>>((const float*)vol, (float*)out, wl, a);
  else if (dtype == PCB_F16) extract_kernel<__half><<<grid, 256, 0, st>>>((const __half*)vol, (__half*)out, wl, a);
  else if (dtype == PCB_BF16) extract_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)vol, (__nv_bfloat16*)out, wl, a);
  else { set_error("pcb_sw_extract: bad dtype %d", dtype); return PCB_ERR_INVALID; }
  PCB_CHECK_LAUNCH("pcb_sw_extract");
  return PCB_OK;
}

static inline size_t dtype_size(int dtype) { return dtype == PCB_F32 ? 4 : 2; }

extern "C" int pcb_sw_extract(const void* vol, int dtype, int64_t C, const int64_t image[3], const int64_t roi[3],
                              const int64_t* starts, int64_t n, int pad_mode, double cval, void* out, void* stream) {
  PCB_CHECK_ARG(vol && image && roi && starts && out, "pcb_sw_extract: null argument");
  PCB_CHECK_ARG(n > 0 && n <= 65535, "pcb_sw_extract: batch of %lld windows unsupported", (long long)n);
  PCB_CHECK_ARG(pad_mode >= 0 && pad_mode <= 3, "pcb_sw_extract: bad padding mode %d", pad_mode);
  PCB_CHECK_ARG(dtype >= PCB_F32 && dtype <= PCB_BF16, "pcb_sw_extract: bad dtype %d", dtype);
  // window starts travel as kernel parameters (SW_MAXB per launch): no device allocation, no host->device copy and no
  // synchronisation in the tile loop — the host can enqueue ahead and the call is CUDA-graph capturable
  const int64_t per = C * roi[0] * roi[1] * roi[2];
  for (int64_t b0 = 0; b0 < n; b0 += SW_MAXB) {
    WinList wl;
    memset(&wl, 0, sizeof(wl));
    wl.n = (int)(n - b0 < SW_MAXB ? n - b0 : SW_MAXB);
    for (int w = 0; w < wl.n; ++w)
      for (int ax = 0; ax < 3; ++ax) wl.s[w][ax] = starts[3 * (b0 + w) + ax];
    int rc = extract_launch(vol, dtype, C, image, roi, wl, pad_mode, cval, (char*)out + (size_t)b0 * per * dtype_size(dtype),
                            (cudaStream_t)stream);
    if (rc) return rc;
  }
  return PCB_OK;
}

extern "C" int pcb_sw_accumulate(const void* pred, const void* map, void* value, void* weight, int dtype,
                                 int64_t Cout, const int64_t roi[3], const int64_t out_size[3],
                                 const int64_t pred_lo[3], const int64_t out_lo[3], const int64_t box[3],
                                 void* stream) {
  PCB_CHECK_ARG(pred && map && value && weight && roi && out_size && pred_lo && out_lo && box, "pcb_sw_accumulate: null argument");
  for (int a = 0; a < 3; ++a) {
    PCB_CHECK_ARG(box[a] > 0 && pred_lo[a] >= 0 && pred_lo[a] + box[a] <= roi[a], "pcb_sw_accumulate: window box outside the ROI");
    PCB_CHECK_ARG(out_lo[a] >= 0 && out_lo[a] + box[a] <= out_size[a], "pcb_sw_accumulate: box outside the accumulator");
  }
  cudaStream_t st = (cudaStream_t)stream;
  AccArgs a{Cout, roi[0], roi[1], roi[2], out_size[0], out_size[1], out_size[2], pred_lo[0], pred_lo[1], pred_lo[2],
            out_lo[0], out_lo[1], out_lo[2], box[0], box[1], box[2]};
  const int64_t nbox = box[0] * box[1] * box[2];
  if (dtype == PCB_F32) accumulate_kernel<float><<<grid_for(nbox), 256, 0, st>>>((const float*)pred, (const float*)map, (float*)value, (float*)weight, a);
  else if (dtype == PCB_F16) accumulate_kernel<__half><<<grid_for(nbox), 256, 0, st>>>((const __half*)pred, (const __half*)map, (__half*)value, (__half*)weight, a);
  else if (dtype == PCB_BF16) accumulate_kernel<__nv_bfloat16><<<grid_for(nbox), 256, 0, st>>>((const __nv_bfloat16*)pred, (const __nv_bfloat16*)map, (__nv_bfloat16*)value, (__nv_bfloat16*)weight, a);
  else { set_error("pcb_sw_accumulate: bad dtype %d", dtype); return PCB_ERR_INVALID; }
  PCB_CHECK_LAUNCH("pcb_sw_accumulate");
  return PCB_OK;
}

static int accumulate_batch_launch(const void* pred, const void* map, void* value, void* weight, int dtype, int64_t Cout,
                                   const int64_t roi[3], const int64_t out_size[3], const WinList& wl, cudaStream_t st) {
  AccBatchArgs a{Cout, roi[0], roi[1], roi[2], out_size[0], out_size[1], out_size[2]};
  dim3 grid(grid_for(roi[0] * roi[1] * roi[2]), (unsigned)wl.n);
  if (dtype == PCB_F32) accumulate_batch_kernel<float><<<grid, 256, 0, st>>>((const float*)pred, (const float*)map, (float*)value, (float*)weight, wl, a);
  else if (dtype == PCB_F16) accumulate_batch_kernel<__half><<<grid, 256, 0, st>>>((const __half*)pred, (const __half*)map, (__half*)value, (__half*)weight, wl, a);
  else if (dtype == PCB_BF16) accumulate_batch_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)pred, (const __nv_bfloat16*)map, (__nv_bfloat16*)value, (__nv_bfloat16*)weight, wl, a);
  else { set_error("pcb_sw_accumulate_batch: bad dtype %d", dtype); return PCB_ERR_INVALID; }
  PCB_CHECK_LAUNCH("pcb_sw_accumulate_batch");
  return PCB_OK;
}

extern "C" int pcb_sw_accumulate_batch(const void* pred, const void* map, void* value, void* weight, int dtype, int64_t Cout,
                                       const int64_t roi[3], const int64_t out_size[3], const int64_t* starts, int64_t n,
                                       void* stream) {
  PCB_CHECK_ARG(pred && map && value && weight && roi && out_size && starts, "pcb_sw_accumulate_batch: null argument");
  PCB_CHECK_ARG(n > 0 && Cout > 0, "pcb_sw_accumulate_batch: empty batch");
  PCB_CHECK_ARG(dtype >= PCB_F32 && dtype <= PCB_BF16, "pcb_sw_accumulate_batch: bad dtype %d", dtype);
  for (int64_t w = 0; w < n; ++w)
    for (int a = 0; a < 3; ++a)
      PCB_CHECK_ARG(starts[3 * w + a] >= 0 && starts[3 * w + a] + roi[a] <= out_size[a],
                    "pcb_sw_accumulate_batch: window %lld outside the accumulator", (long long)w);
  const int64_t per = Cout * roi[0] * roi[1] * roi[2];
  for (int64_t b0 = 0; b0 < n; b0 += SW_MAXB) {       // launches run in stream order, so list order is preserved
    WinList wl;
    memset(&wl, 0, sizeof(wl));
    wl.n = (int)(n - b0 < SW_MAXB ? n - b0 : SW_MAXB);
    for (int w = 0; w < wl.n; ++w)
      for (int ax = 0; ax < 3; ++ax) wl.s[w][ax] = starts[3 * (b0 + w) + ax];
    int rc = accumulate_batch_launch((const char*)pred + (size_t)b0 * per * dtype_size(dtype), map, value, weight, dtype, Cout,
                                     roi, out_size, wl, (cudaStream_t)stream);
    if (rc) return rc;
  }
  return PCB_OK;
}

extern "C" int pcb_sw_normalize(void* value, const void* weight, int dtype, int64_t Cout, int64_t nvox, void* stream) {
  PCB_CHECK_ARG(value && weight && Cout > 0 && nvox > 0, "pcb_sw_normalize: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t n = Cout * nvox;
  if (dtype == PCB_F32) normalize_kernel<float><<<grid_for(n), 256, 0, st>>>((float*)value, (const float*)weight, Cout, nvox, 1.0e-4f);
  else if (dtype == PCB_F16) normalize_kernel<__half><<<grid_for(n), 256, 0, st>>>((__half*)value, (const __half*)weight, Cout, nvox, 1.0e-4f);
  else if (dtype == PCB_BF16) normalize_kernel<__nv_bfloat16><<<grid_for(n), 256, 0, st>>>((__nv_bfloat16*)value, (const __nv_bfloat16*)weight, Cout, nvox, 1.0e-4f);
  else { set_error("pcb_sw_normalize: bad dtype %d", dtype); return PCB_ERR_INVALID; }
  PCB_CHECK_LAUNCH("pcb_sw_normalize");
  return PCB_OK;
}

// ---------------------------------------------------------------------------- the native tile loop (pcb_sw_run)
// window.py:563-683 for a pcb_net: for every batch of windows  crop+pad -> network -> overlap-add blend, all enqueued
// by the library.  The batch body reads its window starts from a device table through a device cursor, so ONE captured
// CUDA graph serves every batch (use_graph): per batch the host does a single cudaGraphLaunch instead of ~50 kernel
// launches.  Windows are blended in list order (pcb_sw_accumulate_batch semantics) => same fp association as the
// reference's sequential loop.
extern "C" void** pcb_net_graph_slot(pcb_net* net, uint64_t** key_out);
extern "C" int32_t pcb_net_in_channels(const pcb_net* net);
extern "C" int32_t pcb_net_head_channels(const pcb_net* net, int32_t head);
extern "C" int32_t pcb_net_num_heads(const pcb_net* net);

namespace {
struct SwRunLayout { int64_t table, cursor, batch, pred, net, total; int64_t padded; };
SwRunLayout sw_run_layout(const pcb_net* net, int head, const int64_t roi[3], int sw_batch, int64_t nstarts, int vol_dtype,
                          int acc_dtype) {
  SwRunLayout l;
  auto al = [](int64_t v) { return (v + 255) & ~(int64_t)255; };
  const int64_t per = roi[0] * roi[1] * roi[2];
  l.padded = (nstarts + sw_batch - 1) / sw_batch * sw_batch;
  int64_t off = 0;
  l.table = off; off += al(l.padded * 3 * (int64_t)sizeof(int64_t));
  l.cursor = off; off += 256;
  l.batch = off; off += al((int64_t)sw_batch * pcb_net_in_channels(net) * per * (int64_t)dtype_size(vol_dtype));
  l.pred = off; off += al((int64_t)sw_batch * pcb_net_head_channels(net, head) * per * (int64_t)dtype_size(acc_dtype));
  l.net = off;
  const int64_t nw = pcb_net_workspace_bytes(net, sw_batch, roi);
  l.total = nw < 0 ? -1 : off + nw;
  return l;
}
}  // namespace

extern "C" int64_t pcb_sw_run_workspace_bytes(const pcb_net* net, int head, const int64_t roi[3], int sw_batch, int64_t nstarts,
                                              int vol_dtype, int acc_dtype) {
  if (!net || !roi || sw_batch < 1 || sw_batch > SW_MAXB || nstarts < 1 || head < 0 || head >= pcb_net_num_heads(net)) return -1;
  return sw_run_layout(net, head, roi, sw_batch, nstarts, vol_dtype, acc_dtype).total;
}

extern "C" int pcb_sw_run(pcb_net* net, int head, const void* vol, int vol_dtype, const int64_t image[3], const int64_t roi[3],
                          int pad_mode, double cval, int sw_batch, const int64_t* starts, int64_t nstarts, const void* map,
                          void* value, void* weight, int acc_dtype, const int64_t acc_size[3], void* workspace,
                          int64_t ws_bytes, int use_graph, void* stream) {
  PCB_CHECK_ARG(net && vol && image && roi && starts && map && value && weight && acc_size && workspace, "pcb_sw_run: null argument");
  PCB_CHECK_ARG(head >= 0 && head < pcb_net_num_heads(net), "pcb_sw_run: bad head %d", head);
  PCB_CHECK_ARG(sw_batch >= 1 && sw_batch <= SW_MAXB, "pcb_sw_run: sw_batch must be 1..%d (got %d)", SW_MAXB, sw_batch);
  PCB_CHECK_ARG(nstarts >= 1, "pcb_sw_run: no windows");
  PCB_CHECK_ARG(pad_mode >= 0 && pad_mode <= 3, "pcb_sw_run: bad padding mode %d", pad_mode);
  PCB_CHECK_ARG(vol_dtype >= PCB_F32 && vol_dtype <= PCB_BF16 && acc_dtype >= PCB_F32 && acc_dtype <= PCB_BF16, "pcb_sw_run: bad dtype");
  PCB_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "pcb_sw_run: workspace must be 256-byte aligned");
  for (int64_t w = 0; w < nstarts; ++w)
    for (int a = 0; a < 3; ++a)
      PCB_CHECK_ARG(starts[3 * w + a] >= 0 && starts[3 * w + a] + roi[a] <= acc_size[a],
                    "pcb_sw_run: window %lld outside the accumulator", (long long)w);
  const SwRunLayout l = sw_run_layout(net, head, roi, sw_batch, nstarts, vol_dtype, acc_dtype);
  PCB_CHECK_ARG(l.total > 0 && l.total <= ws_bytes, "pcb_sw_run: workspace of %lld bytes needed, %lld given", (long long)l.total,
                (long long)ws_bytes);
  cudaStream_t st = (cudaStream_t)stream;
  char* ws = (char*)workspace;
  int64_t* d_table = (int64_t*)(ws + l.table);
  int64_t* d_cursor = (int64_t*)(ws + l.cursor);
  void* d_batch = ws + l.batch;
  void* d_pred = ws + l.pred;
  const int Cin = pcb_net_in_channels(net), ncls = pcb_net_head_channels(net, head), nheads = pcb_net_num_heads(net);
  // starts table (padded to whole batches with skip sentinels) + cursor = 0
  std::vector<int64_t> table((size_t)l.padded * 3, SW_SKIP);
  memcpy(table.data(), starts, (size_t)nstarts * 3 * sizeof(int64_t));
  if (cudaMemcpyAsync(d_table, table.data(), table.size() * sizeof(int64_t), cudaMemcpyHostToDevice, st) != cudaSuccess ||
      cudaMemsetAsync(d_cursor, 0, 256, st) != cudaSuccess ||
      cudaMemsetAsync(d_batch, 0, (size_t)(l.pred - l.batch), st) != cudaSuccess ||
      cudaStreamSynchronize(st) != cudaSuccess) {      // `table` is a host temporary (once per call, not per batch)
    set_error("pcb_sw_run: staging the window table failed: %s", cudaGetErrorString(cudaGetLastError()));
    return PCB_ERR_CUDA;
  }
  std::vector<void*> outs((size_t)nheads, nullptr);
  outs[head] = d_pred;
  WinList wl;
  memset(&wl, 0, sizeof(wl));
  wl.table = d_table; wl.cursor = d_cursor; wl.total = l.padded; wl.n = sw_batch;
  auto body = [&](cudaStream_t s) -> int {
    int rc = extract_launch(vol, vol_dtype, Cin, image, roi, wl, pad_mode, cval, d_batch, s);
    if (rc) return rc;
    rc = pcb_net_forward(net, d_batch, vol_dtype, sw_batch, roi, outs.data(), acc_dtype, ws + l.net, ws_bytes - l.net, s);
    if (rc) return rc;
    rc = accumulate_batch_launch(d_pred, map, value, weight, acc_dtype, ncls, roi, acc_size, wl, s);
    if (rc) return rc;
    advance_cursor_kernel<<<1, 32, 0, s>>>(d_cursor, (int64_t)sw_batch);
    PCB_CHECK_LAUNCH("pcb_sw_run(cursor)");
    return PCB_OK;
  };
  const int64_t nbatches = l.padded / sw_batch;
  int rc = body(st);                                 // first batch eagerly: configures kernels, resolves driver entry points
  if (rc) return rc;
  int64_t done = 1;
  if (use_graph && nbatches - done >= 2) {
    uint64_t* key = nullptr;
    void** slot = pcb_net_graph_slot(net, &key);
    const uint64_t want[12] = {(uint64_t)vol, (uint64_t)value, (uint64_t)weight, (uint64_t)map, (uint64_t)workspace,
                               (uint64_t)(image[0] * 1000003 + image[1] * 1009 + image[2]),
                               (uint64_t)(roi[0] * 1000003 + roi[1] * 1009 + roi[2]),
                               (uint64_t)(acc_size[0] * 1000003 + acc_size[1] * 1009 + acc_size[2]),
                               (uint64_t)(vol_dtype * 16 + acc_dtype + 256 * pad_mode + 4096 * sw_batch + 65536 * head),
                               (uint64_t)l.padded, (uint64_t)(int64_t)(cval * 1e6), (uint64_t)(uintptr_t)st};
    cudaGraphExec_t exec = (cudaGraphExec_t)*slot;
    if (exec != nullptr && memcmp(key, want, sizeof(want)) != 0) { cudaGraphExecDestroy(exec); exec = nullptr; *slot = nullptr; }
    if (exec == nullptr) {
      // the body is recorded on a private stream (the caller's may be the legacy default stream, which cannot capture);
      // nothing executes during capture, and the instantiated graph is launched on the caller's stream
      cudaGraph_t graph = nullptr;
      cudaStream_t cs = nullptr;
      if (cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking) != cudaSuccess ||
          cudaStreamBeginCapture(cs, cudaStreamCaptureModeRelaxed) != cudaSuccess) {
        if (cs) cudaStreamDestroy(cs);
        set_error("pcb_sw_run: cudaStreamBeginCapture failed: %s", cudaGetErrorString(cudaGetLastError())); return PCB_ERR_CUDA;
      }
      rc = body(cs);
      const cudaError_t ce = cudaStreamEndCapture(cs, &graph);
      cudaStreamDestroy(cs);
      if (rc || ce != cudaSuccess || graph == nullptr) {
        if (graph) cudaGraphDestroy(graph);
        if (!rc) { set_error("pcb_sw_run: graph capture failed: %s", cudaGetErrorString(ce)); rc = PCB_ERR_CUDA; }
        cudaGetLastError();
        return rc;
      }
      if (cudaGraphInstantiate(&exec, graph, 0) != cudaSuccess) {
        cudaGraphDestroy(graph);
        set_error("pcb_sw_run: cudaGraphInstantiate failed: %s", cudaGetErrorString(cudaGetLastError())); return PCB_ERR_CUDA;
      }
      cudaGraphDestroy(graph);
      *slot = exec;
      memcpy(key, want, sizeof(want));
    }
    for (; done < nbatches; ++done) {
      if (cudaGraphLaunch(exec, st) != cudaSuccess) {
        set_error("pcb_sw_run: cudaGraphLaunch failed: %s", cudaGetErrorString(cudaGetLastError())); return PCB_ERR_CUDA;
      }
      count_launch();
    }
  }
  for (; done < nbatches; ++done) {
    rc = body(st);
    if (rc) return rc;
  }
  return PCB_OK;
}
