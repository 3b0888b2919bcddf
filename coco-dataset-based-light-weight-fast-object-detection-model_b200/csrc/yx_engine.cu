// Engine: executes a planned op list (built by the Python graph builder) on one stream.
// Everything is resolved at create time — tensor maps, tile shapes, launch geometry — so a run is
// a fixed sequence of kernel launches that can be replayed as a CUDA graph (bs1 latency).
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <mutex>
#include <string>
#include <vector>

#include "yx_internal.h"

namespace yx {

static thread_local std::string g_last_error;

void set_error(const std::string& msg) { g_last_error = msg; }

int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
  char buf[512];
  snprintf(buf, sizeof buf, "CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file, line, what);
  g_last_error = buf;
  return YX_ERR_CUDA;
}

struct Step {
  yx_op op;
  ConvPlan conv;  // valid when op.kind == YX_OP_CONV
  double flops = 0, bytes = 0;
  SparseWeights sp;      // 2:4-packed weights when the layer's mask is compliant (sp_ok)
  bool sp_ok = false;
  const SparseWeights* spw() const { return sp_ok ? &sp : nullptr; }
};

}  // namespace yx

struct yx_engine {
  std::vector<yx::Step> steps;
  void* arena = nullptr;
  size_t arena_bytes = 0;
  const void* weights = nullptr;
  const void* biases = nullptr;
  int in_h = 0, in_w = 0, batch = 0;
  int num_sms = 148;
  cudaGraphExec_t graph_exec = nullptr;
  cudaStream_t graph_stream = nullptr;  // private stream used only to capture the graph
  bool tuned = false;
  std::vector<void*> owned;   // device memory the engine allocated itself (2:4-packed weights and metadata)
  std::vector<std::string> tune_mismatches;  // YX_TUNE_CHECK: candidates whose output differed from the default shape
  // ---- lanes: small batches (bs1 latency, 8 images per GPU) leave most SMs idle in the deep layers, and the head's four
  // pyramid levels (and its cls / reg branches) do not depend on each other: ops are spread over a few streams, ordered by
  // events derived from the arena byte ranges every op reads and writes (RAW, WAR and WAW, so the arena's live-range reuse
  // stays safe).  Lane 0 is the caller's stream.
  static constexpr int kLanes = 4;
  bool lanes_on = false;
  std::vector<int> lane;                    // per step
  std::vector<std::vector<int>> waits;      // per step: steps on OTHER lanes whose completion events it waits for
  std::vector<char> needs_event;            // per step: another lane waits for it
  std::vector<cudaEvent_t> events;          // per step (created when needs_event), + fork + one join event per side lane
  cudaStream_t lane_streams[kLanes] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t fork_event = nullptr, join_events[kLanes] = {nullptr, nullptr, nullptr, nullptr};
  int lanes_used = 1;
};

using namespace yx;

static bool view_ok(const yx_view& v, size_t arena_bytes) {
  if (v.n < 1 || v.h < 1 || v.w < 1 || v.c < 1 || v.pitch < v.c || v.offset < 0) return false;
  if (v.nstride < (int64_t)v.h * v.w * v.pitch - (v.pitch - v.c)) return false;
  const size_t end = (size_t)v.offset + ((size_t)(v.n - 1) * v.nstride + ((size_t)v.h * v.w - 1) * v.pitch + v.c) * 2;
  return end <= arena_bytes;
}

// launches a conv plan; the image-fed stem gets the caller's image bound at launch time
static int launch_conv(const ConvPlan& plan, const void* image, int image_dtype, float in_scale, float in_shift, cudaStream_t st,
                       bool no_pdl = false) {
  if (no_pdl && !plan.p.img_fused) {
    ConvPlan pl = plan;
    pl.no_pdl = 1;
    return conv_launch(pl, st);
  }
  if (!plan.p.img_fused) return conv_launch(plan, st);
  ConvPlan pl = plan;
  int rc = conv_bind_image(&pl, image, image_dtype, in_scale, in_shift);
  return rc != YX_OK ? rc : conv_launch(pl, st);
}

static int run_step(yx_engine* e, const Step& s, const void* image, int image_dtype, float in_scale, float in_shift,
                    cudaStream_t st, bool no_pdl = false) {
  switch (s.op.kind) {
    case YX_OP_CONV: return launch_conv(s.conv, image, image_dtype, in_scale, in_shift, st, no_pdl);
    case YX_OP_S2D:
      return s2d_launch(image, image_dtype, s.op.aux, e->batch, e->in_h, e->in_w, in_scale, in_shift, e->arena, s.op.dst, st);
    case YX_OP_SPP: return spp_launch(e->arena, s.op.src, s.op.dst, st);
    case YX_OP_UPSAMPLE: return upsample_launch(e->arena, s.op.src, s.op.dst, st);
    case YX_OP_DWCONV: return dwconv_launch(e->arena, s.op, e->weights, e->biases, st);
    default: set_error("unknown op kind"); return YX_ERR_INVALID;
  }
}

// In-place residual convs (dst == res, stored with a TMA reduce-add) are not idempotent.  Tuning and profiling launch
// them repeatedly with plain stores instead (same tiles, same traffic up to the L2 read-modify-write) and restore the
// destination from a snapshot before the one real launch.
struct DstSnapshot {
  void* copy = nullptr;
  uint8_t* first = nullptr;
  size_t bytes = 0;
  int take(yx_engine* e, const yx_view& v, cudaStream_t st) {
    first = static_cast<uint8_t*>(e->arena) + v.offset;
    bytes = ((size_t)(v.n - 1) * v.nstride + ((size_t)v.h * v.w - 1) * v.pitch + v.c) * 2;
    YX_CUDA(cudaMalloc(&copy, bytes));
    YX_CUDA(cudaMemcpyAsync(copy, first, bytes, cudaMemcpyDeviceToDevice, st));
    return YX_OK;
  }
  int restore(cudaStream_t st) {
    YX_CUDA(cudaMemcpyAsync(first, copy, bytes, cudaMemcpyDeviceToDevice, st));
    return YX_OK;
  }
  void release(cudaStream_t st) {
    if (copy) { cudaStreamSynchronize(st); cudaFree(copy); copy = nullptr; }
  }
};

// ---- lanes -------------------------------------------------------------------------------------------------------
namespace {
struct ByteView { int64_t off, nstride, len; int n; };   // n intervals [off + i*nstride, + len), bytes
ByteView bytes_of(const yx_view& v) {
  return {v.offset, v.nstride * 2, (((int64_t)v.h * v.w - 1) * v.pitch + v.c) * 2, v.n};
}
bool overlap(const ByteView& a, const ByteView& b) {
  const int64_t a_end = a.off + (a.n - 1) * a.nstride + a.len, b_end = b.off + (b.n - 1) * b.nstride + b.len;
  if (a_end <= b.off || b_end <= a.off) return false;                      // bounding ranges
  for (int i = 0; i < a.n; ++i) {                                          // per-image intervals (level windows of [B,A,C])
    const int64_t a0 = a.off + i * a.nstride, a1 = a0 + a.len;
    for (int j = 0; j < b.n; ++j) {
      const int64_t b0 = b.off + j * b.nstride, b1 = b0 + b.len;
      if (a0 < b1 && b0 < a1) return true;
    }
  }
  return false;
}
struct Access { std::vector<ByteView> r, w; };
Access access_of(const yx_op& op) {
  Access a;
  a.w.push_back(bytes_of(op.dst));
  const bool reads_src = op.kind != YX_OP_S2D && !(op.kind == YX_OP_CONV && (op.aux & 4));
  if (reads_src) a.r.push_back(bytes_of(op.src));
  if (op.kind == YX_OP_CONV && op.res.c > 0) a.r.push_back(bytes_of(op.res));
  if (op.kind == YX_OP_CONV && op.up.c > 0) a.r.push_back(bytes_of(op.up));
  return a;
}
bool any_overlap(const std::vector<ByteView>& x, const std::vector<ByteView>& y) {
  for (const ByteView& a : x)
    for (const ByteView& b : y)
      if (overlap(a, b)) return true;
  return false;
}
}  // namespace

// Assigns every step a lane and the cross-lane events it has to wait for.  Greedy list scheduling in program order: a step
// continues the lane whose LAST step it depends on (the latest such dependency), otherwise it takes a lane nothing has used
// yet, otherwise the least recently used one; dependencies already ordered before it (same lane, or implied by an earlier
// wait: `seen`) need no event.
static void plan_lanes(yx_engine* e) {
  const int n = (int)e->steps.size(), K = yx_engine::kLanes;
  e->lane.assign(n, 0);
  e->waits.assign(n, {});
  e->needs_event.assign(n, 0);
  std::vector<Access> acc(n);
  for (int i = 0; i < n; ++i) acc[i] = access_of(e->steps[i].op);
  std::vector<int> last(K, -1);                        // last step on each lane
  std::vector<std::vector<int>> seen(K, std::vector<int>(K, -1));   // seen[L][M]: latest step of lane M ordered before lane L's next step
  std::vector<std::vector<int>> snap(n);               // seen[] of a step's lane right after the step (for transitive ordering)
  last[0] = 0;                                         // step 0 (reads the caller's image) runs on the caller's stream, before the fork
  for (int k = 0; k < K; ++k) seen[k][0] = 0;          // ... so every lane is ordered after it
  snap[0] = seen[0];
  for (int i = 1; i < n; ++i) {
    std::vector<int> deps;
    for (int j = i - 1; j >= 0; --j)
      if (any_overlap(acc[j].w, acc[i].r) || any_overlap(acc[j].r, acc[i].w) || any_overlap(acc[j].w, acc[i].w)) deps.push_back(j);
    int L = -1;
    for (int d : deps) {                               // deps are in decreasing order: the latest dependency first
      const int ld = e->lane[d];
      if (last[ld] == d) { L = ld; break; }
    }
    if (L < 0)
      for (int k = 0; k < K && L < 0; ++k)
        if (last[k] < 0) L = k;
    if (L < 0) {
      L = 0;
      for (int k = 1; k < K; ++k)
        if (last[k] < last[L]) L = k;
    }
    e->lane[i] = L;
    for (int d : deps) {
      const int ld = e->lane[d];
      if (ld == L || seen[L][ld] >= d) continue;       // same lane, or already ordered by an earlier wait
      e->waits[i].push_back(d);
      e->needs_event[d] = 1;
      seen[L][ld] = d;
      for (int k = 0; k < K; ++k) seen[L][k] = std::max(seen[L][k], snap[d][k]);
    }
    last[L] = i;
    seen[L][L] = i;
    snap[i] = seen[L];
  }
  e->lanes_used = 1;
  for (int i = 0; i < n; ++i) e->lanes_used = std::max(e->lanes_used, e->lane[i] + 1);
  if (getenv("YX_LANES_VERBOSE"))
    for (int i = 0; i < n; ++i) {
      fprintf(stderr, "lanes: step %3d lane %d%s waits", i, e->lane[i], e->needs_event[i] ? " (event)" : "");
      for (int d : e->waits[i]) fprintf(stderr, " %d", d);
      fprintf(stderr, "\n");
    }
}

static int lanes_init(yx_engine* e) {
  plan_lanes(e);
  e->events.assign(e->steps.size(), nullptr);
  for (size_t i = 0; i < e->steps.size(); ++i)
    if (e->needs_event[i]) YX_CUDA(cudaEventCreateWithFlags(&e->events[i], cudaEventDisableTiming));
  YX_CUDA(cudaEventCreateWithFlags(&e->fork_event, cudaEventDisableTiming));
  for (int k = 1; k < e->lanes_used; ++k) {
    YX_CUDA(cudaStreamCreateWithFlags(&e->lane_streams[k], cudaStreamNonBlocking));
    YX_CUDA(cudaEventCreateWithFlags(&e->join_events[k], cudaEventDisableTiming));
  }
  return YX_OK;
}

// Steps [first, n) over the lanes; `st` is lane 0.  Works under stream capture too (the side streams join the capture through
// the fork event and rejoin `st` before it returns).
static int run_lanes(yx_engine* e, size_t first, const void* image, int image_dtype, float in_scale, float in_shift, cudaStream_t st) {
  YX_CUDA(cudaEventRecord(e->fork_event, st));
  for (int k = 1; k < e->lanes_used; ++k) YX_CUDA(cudaStreamWaitEvent(e->lane_streams[k], e->fork_event, 0));
  int rc = YX_OK;
  bool started[yx_engine::kLanes] = {true, false, false, false};
  for (size_t i = first; i < e->steps.size() && rc == YX_OK; ++i) {
    cudaStream_t s = e->lane[i] == 0 ? st : e->lane_streams[e->lane[i]];
    bool waited = !started[e->lane[i]];   // (first kernel of a side lane: it follows the fork's event wait, not a kernel)
    started[e->lane[i]] = true;
    for (int d : e->waits[i])
      if ((size_t)d >= first) { YX_CUDA(cudaStreamWaitEvent(s, e->events[d], 0)); waited = true; }
    rc = run_step(e, e->steps[i], image, image_dtype, in_scale, in_shift, s, waited);
    if (rc == YX_OK && e->needs_event[i]) YX_CUDA(cudaEventRecord(e->events[i], s));
  }
  for (int k = 1; k < e->lanes_used; ++k) {   // always rejoin (also after an error: a capture must not be left forked)
    cudaEventRecord(e->join_events[k], e->lane_streams[k]);
    cudaStreamWaitEvent(st, e->join_events[k], 0);
  }
  return rc;
}

extern "C" const char* yx_last_error(void) { return g_last_error.c_str(); }
extern "C" int yx_abi_version(void) { return YX_ABI_VERSION; }

extern "C" int yx_engine_create(const yx_op* ops, int n_ops, void* arena, size_t arena_bytes, const void* weights,
                                size_t weights_bytes, const void* biases, size_t bias_bytes, int in_h, int in_w,
                                int batch, yx_engine** out) {
  YX_REQUIRE(ops && n_ops > 0 && arena && weights && biases && out, "null argument");
  YX_REQUIRE(((uintptr_t)arena % 1024) == 0 && ((uintptr_t)weights % 256) == 0 && ((uintptr_t)biases % 256) == 0,
             "arena must be 1024-byte aligned, weights/biases 256-byte aligned");
  YX_REQUIRE(ops[0].kind == YX_OP_S2D || (ops[0].kind == YX_OP_CONV && (ops[0].aux & 4)),
             "the first op must read the image: the space-to-depth op or the image-fed stem conv");
  for (int i = 1; i < n_ops; ++i)
    YX_REQUIRE(!(ops[i].kind == YX_OP_CONV && (ops[i].aux & 4)), "only the first op may be fed from the image");
  int dev = 0, sms = 148;
  YX_CUDA(cudaGetDevice(&dev));
  YX_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  int cc_major = 0;
  YX_CUDA(cudaDeviceGetAttribute(&cc_major, cudaDevAttrComputeCapabilityMajor, dev));
  if (cc_major != 10) {
    set_error("yolox_b200 requires an sm_100 (Blackwell B200) device; there is no fallback path");
    return YX_ERR_UNSUPPORTED;
  }
  auto* e = new yx_engine();
  e->arena = arena; e->arena_bytes = arena_bytes; e->weights = weights; e->biases = biases;
  e->in_h = in_h; e->in_w = in_w; e->batch = batch; e->num_sms = sms;
  e->steps.resize(n_ops);
  for (int i = 0; i < n_ops; ++i) {
    Step& s = e->steps[i];
    s.op = ops[i];
    const yx_op& op = s.op;
    char where[64];
    snprintf(where, sizeof where, "op %d: ", i);
    bool ok = view_ok(op.dst, arena_bytes);
    if (op.kind != YX_OP_S2D && !(op.kind == YX_OP_CONV && (op.aux & 4))) ok = ok && view_ok(op.src, arena_bytes);
    if (op.kind == YX_OP_CONV && op.res.c > 0) ok = ok && view_ok(op.res, arena_bytes);
    if (op.kind == YX_OP_CONV && op.up.c > 0) ok = ok && view_ok(op.up, arena_bytes);
    int rc = YX_OK;
    if (!ok) {
      set_error(std::string(where) + "view outside the arena");
      rc = YX_ERR_INVALID;
    } else if (op.kind == YX_OP_CONV) {
      const size_t wend = (size_t)op.w_offset + (size_t)op.cout_pad * op.ksize * ((op.aux & 1) ? 1 : op.ksize) * op.cin_pad * 2;
      const size_t bend = (size_t)op.b_offset + (size_t)op.cout_pad * 4;
      if (wend > weights_bytes || bend > bias_bytes || op.b_offset % 16 != 0) {
        set_error(std::string(where) + "weight/bias range outside the blobs");
        rc = YX_ERR_INVALID;
      } else {
        // 2:4-compliant masks (BASELINE config 3's second mask set; 01_mask_generator-style masks never are): pack the
        // weights for the sparse tensor-core variant; whether a layer then RUNS sparse is the tuner's decision (YX_SPARSE)
        const bool sparse_off = getenv("YX_SPARSE") && strcmp(getenv("YX_SPARSE"), "0") == 0;
        if (!sparse_off && sparse_shape_ok(op)) {
          const int taps = op.ksize * op.ksize, cout128 = round_up(op.cout_pad, 128);
          const size_t wc_bytes = (size_t)cout128 * taps * (op.cin_pad / 2) * 2;
          const size_t meta_bytes = (size_t)(cout128 / 128) * taps * (op.cin_pad / 32) * 128 * 4;
          void *wc = nullptr, *meta = nullptr;
          if (!e->owned.size()) {   // 4-byte counter shared by all packing launches
            void* scratch = nullptr;
            if (cudaMalloc(&scratch, 256) != cudaSuccess) rc = cuda_fail(cudaGetLastError(), "cudaMalloc", __FILE__, __LINE__);
            else e->owned.push_back(scratch);
          }
          if (rc == YX_OK && (cudaMalloc(&wc, wc_bytes) != cudaSuccess || cudaMalloc(&meta, meta_bytes) != cudaSuccess)) {
            rc = cuda_fail(cudaGetLastError(), "cudaMalloc (2:4-packed weights)", __FILE__, __LINE__);
            if (wc) cudaFree(wc);
          }
          if (rc == YX_OK) {
            int compliant = 0;
            rc = sparse_pack(static_cast<const uint8_t*>(weights) + op.w_offset, op.cout_pad, taps, op.cin_pad, wc, meta,
                             static_cast<int*>(e->owned[0]), &compliant, nullptr);
            if (rc == YX_OK && compliant) {
              s.sp.wc = wc; s.sp.meta = static_cast<const uint32_t*>(meta); s.sp_ok = true;
              e->owned.push_back(wc); e->owned.push_back(meta);
            } else {
              cudaFree(wc); cudaFree(meta);
            }
          }
        }
        if (rc == YX_OK) rc = conv_plan(op, arena, weights, biases, sms, nullptr, &s.conv, s.spw());
        s.flops = s.conv.flops; s.bytes = s.conv.bytes;
        if (rc != YX_OK) set_error(std::string(where) + g_last_error);
      }
    } else {
      const double dpx = (double)op.dst.n * op.dst.h * op.dst.w;
      if (op.kind == YX_OP_S2D) s.bytes = dpx * (12.0 * 2 /*image read, fp16*/ + 16.0 * 2);
      else if (op.kind == YX_OP_SPP) s.bytes = dpx * (op.src.c * 2.0 + op.dst.c * 2.0);
      else if (op.kind == YX_OP_UPSAMPLE) s.bytes = dpx * op.dst.c * 2.0 * 1.25;
      else if (op.kind == YX_OP_DWCONV) {
        s.bytes = 2.0 * ((double)op.src.n * op.src.h * op.src.w * op.src.c + dpx * op.dst.c);
        s.flops = 2.0 * dpx * op.dst.c * op.ksize * op.ksize;
      }
    }
    if (rc != YX_OK) { yx_engine_destroy(e); return rc; }
  }
  {  // lanes: on for small batches (YX_LANES=1 forces them on for any batch, 0 turns them off)
    const char* le = getenv("YX_LANES");
    const char* mb = getenv("YX_LANES_MAX_BATCH");   // (experiments; the arena planner reads the same variable)
    e->lanes_on = le ? atoi(le) != 0 : batch <= (mb ? atoi(mb) : 8);
    if (e->lanes_on) {
      int rc = lanes_init(e);
      if (rc != YX_OK) { yx_engine_destroy(e); return rc; }
      if (e->lanes_used < 2) e->lanes_on = false;
    }
  }
  *out = e;
  return YX_OK;
}

extern "C" void yx_engine_destroy(yx_engine* e) {
  if (!e) return;
  for (cudaEvent_t ev : e->events) if (ev) cudaEventDestroy(ev);
  if (e->fork_event) cudaEventDestroy(e->fork_event);
  for (int k = 1; k < yx_engine::kLanes; ++k) {
    if (e->join_events[k]) cudaEventDestroy(e->join_events[k]);
    if (e->lane_streams[k]) cudaStreamDestroy(e->lane_streams[k]);
  }
  if (e->graph_exec) cudaGraphExecDestroy(e->graph_exec);
  if (e->graph_stream) cudaStreamDestroy(e->graph_stream);
  for (void* ptr : e->owned) cudaFree(ptr);
  delete e;
}

extern "C" int yx_engine_num_launches(const yx_engine* e) { return e ? (int)e->steps.size() : 0; }

extern "C" int yx_engine_run(yx_engine* e, const void* image, int image_dtype, float in_scale, float in_shift,
                             int use_graph, void* stream) {
  YX_REQUIRE(e && image, "null argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int rc = run_step(e, e->steps[0], image, image_dtype, in_scale, in_shift, st);
  if (rc) return rc;
  if (!use_graph) {
    if (e->lanes_on) return run_lanes(e, 1, image, image_dtype, in_scale, in_shift, st);
    for (size_t i = 1; i < e->steps.size(); ++i)
      if ((rc = run_step(e, e->steps[i], image, image_dtype, in_scale, in_shift, st)) != YX_OK) return rc;
    return YX_OK;
  }
  if (!e->graph_exec) {
    // capture ops 1..n-1 (none of them touches caller memory, so the graph is replayable as is)
    // Captured on a private stream: the caller's stream may be the legacy default stream, which
    // cannot be captured.  Capture executes nothing; the instantiated graph is launched on `st`.
    cudaGraph_t graph = nullptr;
    if (!e->graph_stream) YX_CUDA(cudaStreamCreateWithFlags(&e->graph_stream, cudaStreamNonBlocking));
    cudaStream_t cs = e->graph_stream;
    YX_CUDA(cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal));
    if (e->lanes_on) rc = run_lanes(e, 1, image, image_dtype, in_scale, in_shift, cs);
    else
      for (size_t i = 1; i < e->steps.size() && rc == YX_OK; ++i)
        rc = run_step(e, e->steps[i], image, image_dtype, in_scale, in_shift, cs);
    cudaError_t ce = cudaStreamEndCapture(cs, &graph);
    if (rc != YX_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
    YX_CUDA(ce);
    ce = cudaGraphInstantiate(&e->graph_exec, graph, 0);
    cudaGraphDestroy(graph);
    YX_CUDA(ce);
  }
  YX_CUDA(cudaGraphLaunch(e->graph_exec, st));
  return YX_OK;
}

extern "C" int yx_engine_run_ops(yx_engine* e, const void* image, int image_dtype, float in_scale, float in_shift,
                                 int first, int count, void* stream) {
  YX_REQUIRE(e && image && first >= 0 && count >= 0 && first + count <= (int)e->steps.size(), "bad op range");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  for (int i = first; i < first + count; ++i) {
    int rc = run_step(e, e->steps[i], image, image_dtype, in_scale, in_shift, st);
    if (rc) return rc;
  }
  return YX_OK;
}


// Per-layer launch-shape selection (the engine's cudnn.benchmark): runs the network ONCE in order on `image`, and for
// every conv times each candidate shape (generic / halo, N tile, CTAs per SM, epilogue groups, staging buffers) on
// the layer's real inputs, keeping the fastest.  All candidates compute the same function, so the arena holds a valid
// forward result afterwards.  Host-synchronising; call it once after create (yx_engine_run works without it, with the
// heuristic shapes).
extern "C" int yx_engine_tune(yx_engine* e, const void* image, int image_dtype, float in_scale, float in_shift, int iters,
                              void* stream) {
  YX_REQUIRE(e && image && iters >= 1, "bad tune arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  cudaEvent_t ev0, ev1;
  YX_CUDA(cudaEventCreate(&ev0));
  YX_CUDA(cudaEventCreate(&ev1));
  int rc = YX_OK;
  const bool verbose = getenv("YX_TUNE_VERBOSE") != nullptr;
  const bool check = getenv("YX_TUNE_CHECK") != nullptr;
  e->tune_mismatches.clear();
  std::vector<ConvTune> cands;
  for (size_t i = 0; i < e->steps.size() && rc == YX_OK; ++i) {
    Step& s = e->steps[i];
    if (s.op.kind != YX_OP_CONV) {
      rc = run_step(e, s, image, image_dtype, in_scale, in_shift, st);
      continue;
    }
    conv_candidates(s.op, &cands, s.sp_ok);
    float best_ms = 1e30f;
    ConvPlan best = s.conv;
    const bool inplace = s.conv.p.has_res == 2;
    DstSnapshot snap;
    if (inplace && (rc = snap.take(e, s.op.dst, st)) != YX_OK) break;
    // YX_TUNE_CHECK=1 (tests): every candidate must reproduce the default shape's output (same function, different
    // fp32 accumulation order => at most a few fp16 steps apart); a candidate that does not is reported and rejected.
    void* check_ref = nullptr;
    unsigned int* check_bits = nullptr;
    if (check) {
      const size_t n_el = (size_t)s.op.dst.n * s.op.dst.h * s.op.dst.w * s.op.dst.c;
      ConvPlan ref_plan = s.conv;
      ref_plan.store_only = inplace ? 1 : 0;
      if (cudaMalloc(&check_ref, n_el * 2) != cudaSuccess || cudaMalloc(&check_bits, 8) != cudaSuccess) {
        rc = cuda_fail(cudaGetLastError(), "tune check alloc", __FILE__, __LINE__);
        break;
      }
      if ((rc = launch_conv(ref_plan, image, image_dtype, in_scale, in_shift, st)) != YX_OK ||
          (rc = view_gather(e->arena, s.op.dst, check_ref, st)) != YX_OK) break;
    }
    for (const ConvTune& t : cands) {
      ConvPlan pl;
      if (conv_plan(s.op, e->arena, e->weights, e->biases, e->num_sms, &t, &pl, s.spw()) != YX_OK) continue;  // shape does not fit
      pl.store_only = inplace ? 1 : 0;
      if ((rc = launch_conv(pl, image, image_dtype, in_scale, in_shift, st)) != YX_OK) break;  // warm-up (also sets the smem attribute)
      if (check) {
        unsigned int bits = 0;
        // a conv that adds a separately loaded residual rounds f(x) to fp16 before the add: a one-step difference of f(x)
        // can be many steps of a small x + f(x), so the step is measured at the activations' O(1) magnitude there
        const float floor_mag = (s.conv.p.has_res == 1) ? 4.0f : 0.0625f;
        if ((rc = view_max_diff(e->arena, s.op.dst, check_ref, floor_mag, check_bits, st)) != YX_OK) break;
        cudaMemcpyAsync(&bits, check_bits, 4, cudaMemcpyDeviceToHost, st);
        cudaStreamSynchronize(st);
        float d;
        memcpy(&d, &bits, 4);
        if (d > 2.0f) {   // same products, other fp32 summation order: at most two fp16 rounding steps apart
          char msg[400];
          snprintf(msg, sizeof msg, "tune check: op %zu candidate '%s' differs from the default shape by %g fp16 steps", i, pl.desc, d);
          fprintf(stderr, "%s\n", msg);
          e->tune_mismatches.push_back(msg);
          continue;
        }
      }
      float ms_min = 1e30f;
      for (int k = 0; k < iters && rc == YX_OK; ++k) {
        cudaEventRecord(ev0, st);
        rc = launch_conv(pl, image, image_dtype, in_scale, in_shift, st);
        cudaEventRecord(ev1, st);
        if (cudaEventSynchronize(ev1) != cudaSuccess) { rc = cuda_fail(cudaGetLastError(), "tune sync", __FILE__, __LINE__); break; }
        float ms = 0;
        cudaEventElapsedTime(&ms, ev0, ev1);
        ms_min = ms < ms_min ? ms : ms_min;
      }
      if (rc != YX_OK) break;
      if (verbose) fprintf(stderr, "  tune op %zu  %-90s %.4f ms\n", i, pl.desc, ms_min);
      if (ms_min < best_ms) { best_ms = ms_min; best = pl; }
    }
    if (check_ref) cudaFree(check_ref);
    if (check_bits) cudaFree(check_bits);
    if (rc != YX_OK) { snap.release(st); break; }
    best.store_only = 0;
    s.conv = best;
    if (verbose) fprintf(stderr, "tune op %zu -> %s  %.4f ms\n", i, best.desc, best_ms);
    if (inplace) rc = snap.restore(st);
    if (rc == YX_OK) rc = launch_conv(s.conv, image, image_dtype, in_scale, in_shift, st);  // leave the chosen variant's output in the arena
    snap.release(st);
  }
  cudaEventDestroy(ev0);
  cudaEventDestroy(ev1);
  if (rc == YX_OK && e->graph_exec) {  // a graph captured with the old shapes is stale
    cudaGraphExecDestroy(e->graph_exec);
    e->graph_exec = nullptr;
  }
  if (rc == YX_OK) YX_CUDA(cudaStreamSynchronize(st));
  e->tuned = rc == YX_OK;
  return rc;
}

static ConvTune tune_from_abi(const yx_conv_tune& a) {
  ConvTune t;
  memset(&t, 0, sizeof t);
  t.variant = a.variant; t.bn = a.n_tile; t.ctas = a.ctas_per_sm; t.mh = a.halves;
  t.epi_groups = a.epilogue_groups; t.stage_bufs = a.staging_buffers; t.w3 = a.second_producer;
  t.no_resident = a.no_resident_weights; t.pair = a.cta_pair; t.sparse = a.sparse; t.epi_alt = a.epilogue_alternate;
  return t;
}
static yx_conv_tune tune_to_abi(const ConvTune& t) {
  yx_conv_tune a;
  memset(&a, 0, sizeof a);
  a.variant = t.variant; a.n_tile = t.bn; a.ctas_per_sm = t.ctas; a.halves = t.mh;
  a.epilogue_groups = t.epi_groups; a.staging_buffers = t.stage_bufs; a.second_producer = t.w3;
  a.no_resident_weights = t.no_resident; a.cta_pair = t.pair; a.sparse = t.sparse; a.epilogue_alternate = t.epi_alt;
  return a;
}

extern "C" int yx_engine_get_tune(const yx_engine* e, int i, yx_conv_tune* out_host) {
  YX_REQUIRE(e && out_host && i >= 0 && i < (int)e->steps.size(), "bad op index");
  YX_REQUIRE(e->steps[i].op.kind == YX_OP_CONV, "not a conv op");
  *out_host = tune_to_abi(e->steps[i].conv.tune);
  return YX_OK;
}

extern "C" int yx_engine_set_tune(yx_engine* e, int i, const yx_conv_tune* tune_host) {
  YX_REQUIRE(e && tune_host && i >= 0 && i < (int)e->steps.size(), "bad op index");
  Step& s = e->steps[i];
  YX_REQUIRE(s.op.kind == YX_OP_CONV, "not a conv op");
  const ConvTune t = tune_from_abi(*tune_host);
  ConvPlan pl;
  int rc = conv_plan(s.op, e->arena, e->weights, e->biases, e->num_sms, &t, &pl, s.spw());
  if (rc != YX_OK) return rc;
  s.conv = pl;
  if (e->graph_exec) {  // a graph captured with the old shape is stale
    cudaGraphExecDestroy(e->graph_exec);
    e->graph_exec = nullptr;
  }
  return YX_OK;
}

extern "C" int yx_engine_tune_mismatches(const yx_engine* e, char* buf_host, int buf_len) {
  YX_REQUIRE(e && buf_host && buf_len > 0, "bad argument");
  std::string all;
  for (const std::string& m : e->tune_mismatches) all += m + "\n";
  snprintf(buf_host, buf_len, "%s", all.c_str());
  return (int)e->tune_mismatches.size();
}

extern "C" int yx_engine_op_desc(const yx_engine* e, int i, char* buf_host, int buf_len) {
  YX_REQUIRE(e && buf_host && buf_len > 0 && i >= 0 && i < (int)e->steps.size(), "bad op index");
  const Step& s = e->steps[i];
  static const char* kinds[] = {"conv", "s2d", "spp", "upsample", "dwconv"};
  if (s.op.kind == YX_OP_CONV) snprintf(buf_host, buf_len, "conv k%d s%d %s%d->%d @%dx%d: %s", s.op.ksize, s.op.stride,
                                        s.op.up.c > 0 ? "up2x+" : "", s.op.src.c + s.op.up.c, s.op.dst.c, s.op.dst.h, s.op.dst.w, s.conv.desc);
  else snprintf(buf_host, buf_len, "%s", s.op.kind >= 0 && s.op.kind <= 4 ? kinds[s.op.kind] : "?");
  return YX_OK;
}

extern "C" int yx_engine_op_sparse_ok(const yx_engine* e, int i) {
  if (!e || i < 0 || i >= (int)e->steps.size()) return 0;
  return (e->steps[i].op.kind == YX_OP_CONV && e->steps[i].sp_ok) ? 1 : 0;
}

extern "C" int yx_engine_profile(yx_engine* e, const void* image, int image_dtype, int iters, void* stream,
                                 float* ms_host, double* flops_host, double* bytes_host, int n_ops) {
  YX_REQUIRE(e && image && ms_host && n_ops == (int)e->steps.size() && iters >= 1, "bad profile arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  cudaEvent_t ev0, ev1;
  YX_CUDA(cudaEventCreate(&ev0));
  YX_CUDA(cudaEventCreate(&ev1));
  int rc = YX_OK;
  for (int i = 0; i < n_ops && rc == YX_OK; ++i) {
    const Step& s = e->steps[i];
    // every op is idempotent on the arena except the in-place residual convs, which are timed with plain stores and
    // then re-run once for real on the restored destination
    const bool inplace = s.op.kind == YX_OP_CONV && s.conv.p.has_res == 2;
    DstSnapshot snap;
    Step timed = s;
    if (inplace) {
      if ((rc = snap.take(e, s.op.dst, st)) != YX_OK) break;
      timed.conv.store_only = 1;
    }
    rc = run_step(e, timed, image, image_dtype, 1.0f, 0.0f, st);  // warm
    if (rc) { snap.release(st); break; }
    cudaEventRecord(ev0, st);
    for (int k = 0; k < iters && rc == YX_OK; ++k) rc = run_step(e, timed, image, image_dtype, 1.0f, 0.0f, st);
    cudaEventRecord(ev1, st);
    if (inplace) {
      if (rc == YX_OK) rc = snap.restore(st);
      if (rc == YX_OK) rc = run_step(e, s, image, image_dtype, 1.0f, 0.0f, st);
    }
    if (cudaEventSynchronize(ev1) != cudaSuccess) { rc = cuda_fail(cudaGetLastError(), "profile sync", __FILE__, __LINE__); break; }
    float ms = 0;
    cudaEventElapsedTime(&ms, ev0, ev1);
    snap.release(st);
    ms_host[i] = ms / iters;
    if (flops_host) flops_host[i] = s.flops;
    if (bytes_host) bytes_host[i] = s.bytes;
  }
  cudaEventDestroy(ev0);
  cudaEventDestroy(ev1);
  return rc;
}

extern "C" int yx_conv2d_ex(const yx_op* op, void* base, const void* weights, const void* biases, const yx_conv_tune* tune,
                            void* stream) {
  YX_REQUIRE(op && base && weights && biases, "null argument");
  YX_REQUIRE(op->kind == YX_OP_CONV, "yx_conv2d expects a YX_OP_CONV op");
  int dev = 0, sms = 148;
  YX_CUDA(cudaGetDevice(&dev));
  YX_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  ConvPlan plan;
  ConvTune t;
  if (tune) t = tune_from_abi(*tune);
  if (tune && tune->sparse) {
    // stand-alone sparse conv (parity tests): pack the 2:4 weights for this one call, run, wait, release
    YX_REQUIRE(sparse_shape_ok(*op), "sparse conv: unsupported geometry (cin % 32, no fused upsample / row-packed stem, metadata <= 256 columns)");
    const int taps = op->ksize * op->ksize, cout128 = round_up(op->cout_pad, 128);
    void *wc = nullptr, *meta = nullptr, *scratch = nullptr;
    YX_CUDA(cudaMalloc(&wc, (size_t)cout128 * taps * (op->cin_pad / 2) * 2));
    YX_CUDA(cudaMalloc(&meta, (size_t)(cout128 / 128) * taps * (op->cin_pad / 32) * 128 * 4));
    YX_CUDA(cudaMalloc(&scratch, 256));
    int compliant = 0;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int rc = sparse_pack(static_cast<const uint8_t*>(weights) + op->w_offset, op->cout_pad, taps, op->cin_pad, wc, meta,
                         static_cast<int*>(scratch), &compliant, st);
    if (rc == YX_OK && !compliant) { set_error("sparse conv: the weights are not 2:4-compliant along the input channels"); rc = YX_ERR_INVALID; }
    SparseWeights spw;
    spw.wc = wc; spw.meta = static_cast<const uint32_t*>(meta);
    if (rc == YX_OK) rc = conv_plan(*op, base, weights, biases, sms, &t, &plan, &spw);
    if (rc == YX_OK) rc = conv_launch(plan, st);
    cudaStreamSynchronize(st);
    cudaFree(wc); cudaFree(meta); cudaFree(scratch);
    return rc;
  }
  int rc = conv_plan(*op, base, weights, biases, sms, tune ? &t : nullptr, &plan);
  if (rc) return rc;
  if (getenv("YX_CONV_TRACE")) {  // diagnostics: print the per-tile timeline of CTA 0 (cycles)
    long long* d = nullptr;
    long long h[32 * 8 + 8];
    YX_CUDA(cudaMalloc(&d, sizeof h));
    YX_CUDA(cudaMemset(d, 0, sizeof h));
    plan.p.trace = d;
    rc = conv_launch(plan, static_cast<cudaStream_t>(stream));
    YX_CUDA(cudaDeviceSynchronize());
    YX_CUDA(cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost));
    cudaFree(d);
    fprintf(stderr, "trace: %s TH %d TW %d tiles %d x %d  (cycles rel. to the first MMA)\n"
                    " tile  prodA_done  mma_start  mma_commit  epi_tfull  epi_staged  store_issue\n",
            plan.desc, plan.p.TH, plan.p.TW, plan.p.n_tiles_m, plan.p.n_tiles_n);
    long long t0 = h[1] ? h[1] : h[0];
    for (int t2 = 0; t2 < 32 && (h[t2 * 8 + 3] || t2 == 0); ++t2)
      fprintf(stderr, " %4d %10lld %10lld %10lld %10lld %10lld %10lld\n", t2, h[t2 * 8 + 0] - t0, h[t2 * 8 + 1] - t0,
              h[t2 * 8 + 2] - t0, h[t2 * 8 + 3] - t0, h[t2 * 8 + 4] - t0, h[t2 * 8 + 5] - t0);
    const long long* w = h + 256;
    const double tot = w[7] > 0 ? (double)w[7] : 1.0;
    fprintf(stderr, "blocked cycles of CTA 0 (%% of %lld): A-producer on empty %.0f%% | B-producer on empty %.0f%% | MMA on fullA %.0f%% "
                    "fullB %.0f%% tmem-empty %.0f%% | epilogue on tmem-full %.0f%% staging/barrier %.0f%%\n",
            w[7], 100.0 * w[0] / tot, 100.0 * w[1] / tot, 100.0 * w[2] / tot, 100.0 * w[3] / tot, 100.0 * w[4] / tot,
            100.0 * w[5] / tot, 100.0 * w[6] / tot);
    return rc;
  }
  return conv_launch(plan, static_cast<cudaStream_t>(stream));
}

extern "C" int yx_conv2d(const yx_op* op, void* base, const void* weights, const void* biases, void* stream) {
  return yx_conv2d_ex(op, base, weights, biases, nullptr, stream);
}
