// pcb200 — network-level inference runtime: a MedNeXt forward plan walked entirely in native code.
//
// Reference: nnunet_mednext MedNextV1.py::MedNeXt.forward as built by connectomics/models/architectures/
// mednext_models.py:374-380 (stem -> encoder stages / down blocks -> bottleneck -> up blocks with encoder skips ->
// decoder stages -> OutBlock heads).  The Python modules (architectures/mednext.py) stay the parameter containers and
// the training path; for inference they hand the library a flat description of the blocks with device pointers to the
// kernel-layout weights, and one pcb_net_forward call enqueues the whole network (~45 kernels for MedNeXt-S) on the
// caller's stream: no Python between kernels, no allocator traffic (activations live in ONE caller-owned arena carved
// by a liveness plan), capturable as a CUDA graph (pcb_sw_run does that for the tile loop).
#include <string.h>

#include <algorithm>
#include <vector>

#include "../../include/pcb200.h"
#include "pcb_common.cuh"

struct pcb_net {
  int cin = 0, c0 = 0;
  const float* stem_w = nullptr;
  const float* stem_b = nullptr;
  std::vector<pcb_block_desc> blocks;
  std::vector<pcb_head_desc> heads;
  // CUDA-graph cache of pcb_sw_run's batch body (owned here so that it dies with the plan)
  void* graph_exec = nullptr;
  uint64_t graph_key[12] = {0};
};

namespace {

inline int64_t align256(int64_t v) { return (v + 255) & ~(int64_t)255; }

// First-fit arena with coalescing free list: activations are allocated when produced and released after their last
// consumer, so the footprint is the peak of the live set, not the sum over layers.
struct Arena {
  struct Seg { int64_t off, size; };
  std::vector<Seg> free_;
  int64_t top = 0, peak = 0;
  int64_t alloc(int64_t bytes) {
    bytes = align256(bytes > 0 ? bytes : 1);
    for (size_t i = 0; i < free_.size(); ++i) {
      if (free_[i].size >= bytes) {
        const int64_t off = free_[i].off;
        free_[i].off += bytes; free_[i].size -= bytes;
        if (free_[i].size == 0) free_.erase(free_.begin() + i);
        return off;
      }
    }
    const int64_t off = top;
    top += bytes;
    if (top > peak) peak = top;
    return off;
  }
  void release(int64_t off, int64_t bytes) {
    bytes = align256(bytes > 0 ? bytes : 1);
    free_.push_back({off, bytes});
    std::sort(free_.begin(), free_.end(), [](const Seg& a, const Seg& b) { return a.off < b.off; });
    for (size_t i = 0; i + 1 < free_.size();) {
      if (free_[i].off + free_[i].size == free_[i + 1].off) { free_[i].size += free_[i + 1].size; free_.erase(free_.begin() + i + 1); }
      else ++i;
    }
    if (!free_.empty() && free_.back().off + free_.back().size == top) { top = free_.back().off; free_.pop_back(); }
  }
};

struct Tensor { int64_t off = -1, bytes = 0; int refs = 0; int64_t size[3] = {0, 0, 0}; int C = 0; };

inline void out_sizes(int mode, int k, const int64_t in[3], int64_t y[3], int64_t o[3]) {
  const int p = k / 2;
  for (int a = 0; a < 3; ++a) {
    if (mode == PCB_DW_SAME) { y[a] = in[a]; o[a] = in[a]; }
    else if (mode == PCB_DW_DOWN) { y[a] = (in[a] + 2 * p - k) / 2 + 1; o[a] = y[a]; }
    else { y[a] = (in[a] - 1) * 2 - 2 * p + k; o[a] = y[a] + 1; }
  }
}
inline int64_t vox(const int64_t s[3]) { return s[0] * s[1] * s[2]; }

// One walk of the plan.  dry == true: only the arena simulation (workspace size); otherwise kernels are enqueued.
int walk(const pcb_net* net, bool dry, const void* x, int in_dtype, int64_t N, const int64_t size[3], void* const* outs,
         int out_dtype, char* ws, int64_t ws_bytes, void* stream, int64_t* peak_out) {
  using pcb::set_error;
  const int nb = (int)net->blocks.size();
  Arena arena;
  // GroupNorm statistics of every block in one region: a single memset per forward
  int64_t stats_bytes = 0;
  std::vector<int64_t> stats_off(nb);
  for (int i = 0; i < nb; ++i) { stats_off[i] = stats_bytes; stats_bytes += align256(N * 2 * net->blocks[i].C * (int64_t)sizeof(double)); }
  const int64_t stats_base = arena.alloc(stats_bytes);
  if (!dry) {
    if (arena.peak > ws_bytes) { set_error("pcb_net_forward: workspace too small"); return PCB_ERR_INVALID; }
    if (cudaMemsetAsync(ws + stats_base, 0, (size_t)stats_bytes, (cudaStream_t)stream) != cudaSuccess) {
      set_error("pcb_net_forward: cudaMemsetAsync failed: %s", cudaGetErrorString(cudaGetLastError())); return PCB_ERR_CUDA;
    }
  }
  std::vector<Tensor> t(nb + 1);                   // t[0] = stem output, t[i + 1] = output of block i
  // consumers: next block, encoder-skip consumers, heads
  t[0].refs = nb > 0 ? 1 : 0;
  for (int i = 0; i < nb; ++i) {
    t[i + 1].refs = (i + 1 < nb) ? 1 : 0;
    const int sf = net->blocks[i].skip_from;
    if (net->blocks[i].kind == PCB_DW_UP && sf >= 0) {
      if (sf >= i) { set_error("pcb_net: block %d takes its skip from a later block %d", i, sf); return PCB_ERR_INVALID; }
      t[sf + 1].refs++;
    }
  }
  for (size_t h = 0; h < net->heads.size(); ++h) {
    const int fb = net->heads[h].from_block;
    if (fb < -1 || fb >= nb) { set_error("pcb_net: head %d reads block %d of %d", (int)h, fb, nb); return PCB_ERR_INVALID; }
    if (outs == nullptr || outs[h] != nullptr || dry) t[fb + 1].refs++;
  }
  auto release = [&](Tensor& tt) { if (--tt.refs <= 0 && tt.off >= 0) { arena.release(tt.off, tt.bytes); tt.off = -1; } };
  auto run_heads = [&](int produced) -> int {       // heads reading tensor `produced` (index into t)
    for (size_t h = 0; h < net->heads.size(); ++h) {
      if (net->heads[h].from_block + 1 != produced) continue;
      if (!dry && outs[h] == nullptr) continue;
      if (!dry) {
        int rc = pcb_head_fwd(ws + t[produced].off, net->heads[h].w, net->heads[h].b, outs[h], out_dtype, N, t[produced].C,
                              net->heads[h].ncls, vox(t[produced].size), stream);
        if (rc) return rc;
      }
      release(t[produced]);
    }
    return PCB_OK;
  };

  for (int a = 0; a < 3; ++a) t[0].size[a] = size[a];
  t[0].C = net->c0;
  t[0].bytes = N * vox(size) * net->c0 * 2;
  t[0].off = arena.alloc(t[0].bytes);
  if (!dry) {
    if (arena.peak > ws_bytes) { set_error("pcb_net_forward: workspace too small"); return PCB_ERR_INVALID; }
    int rc = pcb_stem_fwd(x, in_dtype, net->stem_w, net->stem_b, ws + t[0].off, N, net->cin, net->c0, vox(size), stream);
    if (rc) return rc;
  }
  { int rc = run_heads(0); if (rc) return rc; }

  for (int i = 0; i < nb; ++i) {
    const pcb_block_desc& b = net->blocks[i];
    Tensor& in = t[i];
    if (in.C != b.C) { set_error("pcb_net: block %d expects %d channels, its input has %d", i, b.C, in.C); return PCB_ERR_INVALID; }
    int64_t ys[3], os[3];
    out_sizes(b.kind, b.k, in.size, ys, os);
    const int64_t ybytes = N * vox(ys) * b.C * 2;
    const int64_t yoff = arena.alloc(ybytes);
    const bool has_rc = b.wr != nullptr;
    const int64_t deep = pcb_mlp_fwd_deep_workspace(N, os, b.C, b.H, b.Co, has_rc ? b.C : 0);
    const int64_t hoff = deep > 0 ? arena.alloc(deep) : -1;
    Tensor& out = t[i + 1];
    for (int a = 0; a < 3; ++a) out.size[a] = os[a];
    out.C = b.Co;
    out.bytes = N * vox(os) * b.Co * 2;
    out.off = arena.alloc(out.bytes);
    const Tensor* skip = (b.kind == PCB_DW_UP && b.skip_from >= 0) ? &t[b.skip_from + 1] : nullptr;
    if (skip) {
      if (skip->off < 0) { set_error("pcb_net: skip tensor of block %d was released early", i); return PCB_ERR_INVALID; }
      if (skip->C != b.Co || skip->size[0] != os[0] || skip->size[1] != os[1] || skip->size[2] != os[2]) {
        set_error("pcb_net: skip of block %d has shape [%lld,%lld,%lld,%d], the up block produces [%lld,%lld,%lld,%d]", i,
                  (long long)skip->size[0], (long long)skip->size[1], (long long)skip->size[2], skip->C, (long long)os[0],
                  (long long)os[1], (long long)os[2], b.Co);
        return PCB_ERR_INVALID;
      }
    }
    if (!dry) {
      if (arena.peak > ws_bytes) { set_error("pcb_net_forward: workspace too small"); return PCB_ERR_INVALID; }
      double* stats = reinterpret_cast<double*>(ws + stats_base + stats_off[i]);
      int rc = pcb_dwconv_fwd(ws + in.off, b.w1, b.b1, ws + yoff, stats, N, in.size, b.C, b.k, b.kind, stream);
      if (rc) return rc;
      const void* res = nullptr;
      if (b.kind == PCB_DW_SAME && b.do_res) res = ws + in.off;
      else if (skip) res = ws + skip->off;
      const void* xs = has_rc ? (const void*)(ws + in.off) : nullptr;
      if (deep > 0)
        rc = pcb_mlp_fwd_deep(ws + yoff, stats, b.gamma, b.beta, b.w2, b.b2, b.w3, b.b3, res, xs, b.wr, b.br, ws + out.off,
                              ws + hoff, N, os, in.size, b.C, b.H, b.Co, has_rc ? b.C : 0, b.kind, stream);
      else
        rc = pcb_mlp_fwd(ws + yoff, stats, b.gamma, b.beta, b.w2, b.b2, b.w3, b.b3, res, xs, b.wr, b.br, ws + out.off, N, os,
                         in.size, b.C, b.H, b.Co, has_rc ? b.C : 0, b.kind, stream);
      if (rc) return rc;
    }
    arena.release(yoff, ybytes);
    if (hoff >= 0) arena.release(hoff, deep);
    release(in);
    if (skip) release(t[b.skip_from + 1]);
    int rc = run_heads(i + 1);
    if (rc) return rc;
    if (out.refs <= 0 && out.off >= 0) { arena.release(out.off, out.bytes); out.off = -1; }
  }
  if (peak_out) *peak_out = arena.peak;
  return PCB_OK;
}

}  // namespace

extern "C" int pcb_net_create(int32_t in_channels, int32_t n_channels, const float* stem_w, const float* stem_b,
                              const pcb_block_desc* blocks, int32_t nblocks, const pcb_head_desc* heads, int32_t nheads,
                              pcb_net** out) {
  PCB_CHECK_ARG(out && stem_w && stem_b && (blocks || nblocks == 0) && heads && nheads > 0, "pcb_net_create: null argument");
  PCB_CHECK_ARG(in_channels > 0 && n_channels > 0 && n_channels % 16 == 0, "pcb_net_create: base_channels must be a multiple of 16");
  for (int i = 0; i < nblocks; ++i) {
    const pcb_block_desc& b = blocks[i];
    PCB_CHECK_ARG(b.kind >= PCB_DW_SAME && b.kind <= PCB_DW_UP, "pcb_net_create: block %d has bad kind %d", i, b.kind);
    PCB_CHECK_ARG(b.norm == 0, "pcb_net_create: block %d: only GroupNorm(num_groups=C) blocks run on the native plan", i);
    PCB_CHECK_ARG(b.w1 && b.b1 && b.gamma && b.beta && b.w2 && b.b2 && b.w3 && b.b3 && (b.wr == nullptr) == (b.br == nullptr),
                  "pcb_net_create: block %d has null weights", i);
    PCB_CHECK_ARG(b.C > 0 && b.H > 0 && b.Co > 0 && (b.k == 3 || b.k == 5 || b.k == 7), "pcb_net_create: block %d has bad shape", i);
  }
  for (int h = 0; h < nheads; ++h)
    PCB_CHECK_ARG(heads[h].w && heads[h].b && heads[h].ncls > 0, "pcb_net_create: head %d is incomplete", h);
  pcb_net* n = new pcb_net();
  n->cin = in_channels; n->c0 = n_channels; n->stem_w = stem_w; n->stem_b = stem_b;
  n->blocks.assign(blocks, blocks + nblocks);
  n->heads.assign(heads, heads + nheads);
  *out = n;
  return PCB_OK;
}

extern "C" void pcb_net_destroy(pcb_net* net) {
  if (!net) return;
  if (net->graph_exec) cudaGraphExecDestroy((cudaGraphExec_t)net->graph_exec);
  delete net;
}

extern "C" int32_t pcb_net_in_channels(const pcb_net* net) { return net ? net->cin : 0; }
extern "C" int32_t pcb_net_head_channels(const pcb_net* net, int32_t head) {
  return (net && head >= 0 && head < (int)net->heads.size()) ? net->heads[head].ncls : 0;
}
extern "C" int32_t pcb_net_num_heads(const pcb_net* net) { return net ? (int32_t)net->heads.size() : 0; }

extern "C" int64_t pcb_net_workspace_bytes(const pcb_net* net, int64_t N, const int64_t size[3]) {
  if (!net || !size || N <= 0) return -1;
  int64_t peak = 0;
  if (walk(net, true, nullptr, 0, N, size, nullptr, 0, nullptr, 0, nullptr, &peak) != PCB_OK) return -1;
  return peak;
}

extern "C" int pcb_net_forward(pcb_net* net, const void* x, int in_dtype, int64_t N, const int64_t size[3],
                               void* const* outs, int out_dtype, void* workspace, int64_t ws_bytes, void* stream) {
  PCB_CHECK_ARG(net && x && size && outs && workspace, "pcb_net_forward: null argument");
  PCB_CHECK_ARG(N > 0 && size[0] > 0 && size[1] > 0 && size[2] > 0, "pcb_net_forward: empty input");
  PCB_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "pcb_net_forward: workspace must be 256-byte aligned");
  return walk(net, false, x, in_dtype, N, size, outs, out_dtype, (char*)workspace, ws_bytes, stream, nullptr);
}

// pcb_sw_run's graph cache lives in the plan (see sw_kernels.cu)
extern "C" void** pcb_net_graph_slot(pcb_net* net, uint64_t** key_out) {
  if (key_out) *key_out = net->graph_key;
  return &net->graph_exec;
}
