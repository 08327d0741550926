// One C call per residual cell of the conv encoder / decoder: pvae_cell_fwd / pvae_cell_bwd enqueue every kernel of a
// cell (weight normalisation, implicit-GEMM convolutions, squeeze-excitation, skip branch; in the backward also the
// weight / bias / lognorm gradients) back to back from C++, instead of one Python -> ctypes round trip per kernel.
// The kernels and their order are exactly those of the per-kernel path (poissonvae_b200/conv.py with
// NATIVE_CELLS = False), so both paths produce bit-identical results (tests/test_conv_gpu.py).
//
// Reference semantics: Cell.forward base/common.py:187-192 with ConvLayer.forward :233-242, SELayer :138-142,
// FactorizedReduce / StrideReduce :75-126, the up-sampling skip :35-47.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/pvae_b200.h"
#include "pvae_host.h"

namespace pvae {

static inline int pad32(int c) { return (c + 31) / 32 * 32; }
static inline long long align64(long long n) { return (n + 63) / 64 * 64; }

struct Bump {
  float* base;
  long long used = 0;
  explicit Bump(float* b) : base(b) {}
  float* take(long long n) {
    float* p = base ? base + used : nullptr;
    used += align64(n);
    return p;
  }
};

struct Geometry {
  int N, H, W, OH, OW, CiP, CoP, Co, hd, stride, ks, groups;
  bool dec, up, down;
  long long in_px, out_px;
};

static Geometry geometry(const pvae_cell& c) {
  Geometry g;
  g.N = c.N, g.H = c.H, g.W = c.W;
  g.dec = c.kind == PVAE_CELL_NORMAL_DEC || c.kind == PVAE_CELL_UP_DEC;
  g.up = c.kind == PVAE_CELL_UP_DEC;
  g.down = c.kind == PVAE_CELL_DOWN_FACTORIZED || c.kind == PVAE_CELL_DOWN_STRIDE;
  g.stride = g.down ? 2 : 1;
  g.OH = g.up ? 2 * c.H : (g.down ? (c.H + 2 - 3) / 2 + 1 : c.H);
  g.OW = g.up ? 2 * c.W : (g.down ? (c.W + 2 - 3) / 2 + 1 : c.W);
  g.CiP = pad32(c.Ci), g.CoP = pad32(c.Co), g.Co = c.Co;
  g.hd = std::max(c.Co / 16, 4);  // SELayer base/common.py:132
  g.ks = c.kind == PVAE_CELL_DOWN_FACTORIZED ? 2 : 1;
  g.groups = c.kind == PVAE_CELL_DOWN_FACTORIZED ? 4 : 1;
  g.in_px = static_cast<long long>(c.N) * c.H * c.W;
  g.out_px = static_cast<long long>(c.N) * g.OH * g.OW;
  return g;
}

// tensors kept from the forward for the backward
struct Saved {
  float *x_act, *wf1, *wd1, *bp1, *wf2, *wd2, *bp2, *wfs, *wds, *bps, *c1, *a2, *c2, *skip, *au, *pool, *hid, *gate;
  long long total;
};

static Saved carve_saved(const pvae_cell& c, float* base, bool own_act) {
  const Geometry g = geometry(c);
  Bump b(base);
  Saved s = {};
  s.x_act = own_act ? b.take(g.in_px * g.CiP) : nullptr;
  // packed filters carry their TF32 low parts as a second block of rows (pvae_conv_weight_prep with_lo = 1): 2x
  s.wf1 = b.take(18ll * g.CoP * g.CiP), s.wd1 = b.take(18ll * g.CoP * g.CiP), s.bp1 = b.take(g.CoP);
  if (!g.dec) {
    s.wf2 = b.take(18ll * g.CoP * g.CoP), s.wd2 = b.take(18ll * g.CoP * g.CoP), s.bp2 = b.take(g.CoP);
    s.c1 = b.take(g.out_px * g.CoP), s.a2 = b.take(g.out_px * g.CoP);
  }
  s.c2 = b.take(g.out_px * g.CoP);  // the convolution output the gate multiplies (decoder: the only one)
  if (g.down) {
    s.wfs = b.take(2ll * g.ks * g.ks * g.CoP * g.CiP), s.wds = b.take(2ll * g.ks * g.ks * g.CoP * g.CiP);
    s.bps = b.take(g.CoP);
    s.skip = b.take(g.out_px * g.CoP);
  }
  if (g.up) {
    s.wfs = b.take(2ll * g.CoP * g.CiP), s.wds = b.take(2ll * g.CoP * g.CiP), s.bps = b.take(g.CoP);
    s.skip = b.take(g.in_px * g.CoP);  // 1x1 convolution on the small map
    s.au = b.take(g.out_px * g.CiP);
  }
  s.pool = b.take(static_cast<long long>(g.N) * g.Co), s.hid = b.take(static_cast<long long>(g.N) * g.hd);
  s.gate = b.take(static_cast<long long>(g.N) * g.Co);
  s.total = b.used;
  return s;
}

struct Scratch {
  float *g_c2, *g_c1, *g_a, *g_au, *t, *g_s, *d2, *d1, *ws, *rws;
  long long total;
};

static Scratch carve_scratch(const pvae_cell& c, float* base) {
  const Geometry g = geometry(c);
  Bump b(base);
  Scratch s = {};
  s.g_c2 = b.take(g.out_px * g.CoP);
  if (!g.dec) s.g_c1 = b.take(g.out_px * g.CoP);
  if (g.down) s.g_a = b.take(g.in_px * g.CiP);
  if (g.up) s.g_au = b.take(g.out_px * g.CiP), s.t = b.take(g.in_px * g.CiP), s.g_s = b.take(g.in_px * g.CoP);
  s.d2 = b.take(static_cast<long long>(g.N) * g.Co), s.d1 = b.take(static_cast<long long>(g.N) * g.hd);
  long long w = std::max(pvae_conv_wgrad_ws_floats(g.CiP, g.CoP, 9), pvae_conv_wgrad_ws_floats(g.CoP, g.CoP, 9));
  w = std::max<long long>(w, pvae_conv_wgrad_ws_floats(g.CiP, g.CoP, g.ks * g.ks));
  s.ws = b.take(w);
  s.rws = b.take(pvae_reduce_ws_floats(2ll * g.Co * g.hd + g.Co + g.hd));
  s.total = b.used;
  return s;
}

// Side stream for the weight-gradient work of a cell's backward (off the g_y -> g_x critical path); one per device.
struct CellSide {
  cudaStream_t stream = nullptr;
  cudaEvent_t fork[3] = {nullptr, nullptr, nullptr}, join = nullptr;
};
static CellSide* cell_side() {
  static CellSide per_device[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  CellSide& s = per_device[dev];
  if (!s.stream) {
    if (cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
    for (auto& e : s.fork)
      if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&s.join, cudaEventDisableTiming) != cudaSuccess) return nullptr;
  }
  return &s;
}
static int g_cell_overlap = 1;

// the side stream waits for everything enqueued on `main` so far
static int fork_to(CellSide* side, int i, void* main_stream) {
  if (!side) return PVAE_OK;
  if (int rc = cuda_ok("fork", cudaEventRecord(side->fork[i], static_cast<cudaStream_t>(main_stream)))) return rc;
  return cuda_ok("fork", cudaStreamWaitEvent(side->stream, side->fork[i], 0));
}
static int join_from(CellSide* side, void* main_stream) {
  if (!side) return PVAE_OK;
  if (int rc = cuda_ok("join", cudaEventRecord(side->join, side->stream))) return rc;
  return cuda_ok("join", cudaStreamWaitEvent(static_cast<cudaStream_t>(main_stream), side->join, 0));
}

struct Taps {
  int n = 0;
  int dy[16], dx[16], wi[16];
  void add(int y, int x, int w) { dy[n] = y, dx[n] = x, wi[n] = w, ++n; }
};

// forward convolution of an NHWC map (channel counts already padded)
static int conv_forward(const float* in, const float* wf, const float* bias_p, float* out, float* out_act, int N, int IH, int IW,
                        int CiP, int CoP, int k, int stride, int pad, int precise, void* s) {
  const int OH = (IH + 2 * pad - k) / stride + 1, OW = (IW + 2 * pad - k) / stride + 1;
  Taps t;
  for (int r = 0; r < k; ++r)
    for (int q = 0; q < k; ++q) t.add(r - pad, q - pad, r * k + q);
  return pvae_conv_igemm(in, wf, out, out_act, bias_p, nullptr, nullptr, nullptr, N, IH, IW, CiP, CoP, OH, OW, OH, OW, 1, 0, 0, stride,
                         t.n, t.dy, t.dx, t.wi, k * k, 1, precise, s);
}

// data gradient with the fused epilogue g_in = (conv^T(g_out) + add_pre) * silu'(dsilu_src) + add_post
static int conv_dgrad(const float* g_out, const float* wd, float* g_in, int N, int IH, int IW, int CiP, int OH, int OW, int CoP, int k,
                      int stride, int pad, const float* add_pre, const float* dsilu_src, const float* add_post, int precise, void* s) {
  if (stride == 1) {
    Taps t;
    for (int r = 0; r < k; ++r)
      for (int q = 0; q < k; ++q) t.add(pad - r, pad - q, r * k + q);
    return pvae_conv_igemm(g_out, wd, g_in, nullptr, nullptr, add_pre, dsilu_src, add_post, N, OH, OW, CoP, CiP, IH, IW, IH, IW, 1, 0, 0, 1,
                           t.n, t.dy, t.dx, t.wi, k * k, 1, precise, s);
  }
  for (int ph = 0; ph < 2; ++ph)  // one launch per parity class of the input grid
    for (int pw = 0; pw < 2; ++pw) {
      const int qh = (IH - ph + 1) / 2, qw = (IW - pw + 1) / 2;
      if (qh <= 0 || qw <= 0) continue;
      Taps t;
      for (int r = 0; r < k; ++r)
        for (int q = 0; q < k; ++q)
          if (((ph + pad - r) % 2 == 0) && ((pw + pad - q) % 2 == 0)) {
            // floor division of possibly negative even numbers
            const int a = ph + pad - r, b2 = pw + pad - q;
            t.add(a >= 0 ? a / 2 : -((-a) / 2), b2 >= 0 ? b2 / 2 : -((-b2) / 2), r * k + q);
          }
      if (int rc = pvae_conv_igemm(g_out, wd, g_in, nullptr, nullptr, add_pre, dsilu_src, add_post, N, OH, OW, CoP, CiP, qh, qw, IH, IW,
                                   2, ph, pw, 1, t.n, t.dy, t.dx, t.wi, k * k, 1, precise, s))
        return rc;
    }
  return PVAE_OK;
}

static int conv_wgrad(const float* in, const float* g_out, const float* weight, const float* lognorm, float* g_w, float* g_ln, float* g_b,
                      float* ws, int N, int IH, int IW, int CiP, int CoP, int OH, int OW, int k, int stride, int pad, int groups,
                      int Ci, int Co, int precise, void* s) {
  Taps t;
  for (int r = 0; r < k; ++r)
    for (int q = 0; q < k; ++q) t.add(r - pad, q - pad, r * k + q);
  return pvae_conv_wgrad(in, g_out, weight, lognorm, g_w, g_ln, g_b, ws, N, IH, IW, CiP, CoP, OH, OW, stride, t.n, t.dy, t.dx, t.wi, groups,
                         Ci, Co, precise, s);
}

static int check_cell(const pvae_cell* c, const char* fn) {
  if (!c) return fail(PVAE_ERR_INVALID, "%s: null cell", fn);
  if (c->kind < 0 || c->kind > PVAE_CELL_UP_DEC || c->N < 1 || c->H < 1 || c->W < 1 || c->Ci < 4 || c->Co < 4 || c->Ci % 4 || c->Co % 4)
    return fail(PVAE_ERR_INVALID, "%s: bad cell geometry", fn);
  return PVAE_OK;
}

}  // namespace pvae

using namespace pvae;

#define TRY(expr)                  \
  do {                             \
    if (int rc_ = (expr)) return rc_; \
  } while (0)

extern "C" int64_t pvae_cell_saved_floats(const pvae_cell* c, int own_act) {
  if (check_cell(c, "pvae_cell_saved_floats")) return -1;
  return carve_saved(*c, nullptr, own_act != 0).total;
}

extern "C" int64_t pvae_cell_scratch_floats(const pvae_cell* c) {
  if (check_cell(c, "pvae_cell_scratch_floats")) return -1;
  return carve_scratch(*c, nullptr).total;
}

extern "C" int pvae_cell_fwd(const pvae_cell* c, const float* x, const float* x_act, float* y, float* y_act, float* saved, void* stream) {
  TRY(check_cell(c, "pvae_cell_fwd"));
  if (!x || !y || !y_act || !saved) return fail(PVAE_ERR_INVALID, "pvae_cell_fwd: null pointer");
  const Geometry g = geometry(*c);
  const Saved s = carve_saved(*c, saved, x_act == nullptr);
  if (x_act == nullptr) {
    TRY(pvae_silu_fwd(x, s.x_act, g.in_px * g.CiP, stream));
    x_act = s.x_act;
  }
  TRY(pvae_conv_weight_prep(c->w1, c->ln1, c->b1, s.wf1, s.wd1, s.bp1, c->Co, c->Ci, g.CoP, g.CiP, 9, 1, 1, stream));
  if (!g.dec) {
    TRY(pvae_conv_weight_prep(c->w2, c->ln2, c->b2, s.wf2, s.wd2, s.bp2, c->Co, c->Co, g.CoP, g.CoP, 9, 1, 1, stream));
    TRY(conv_forward(x_act, s.wf1, s.bp1, s.c1, s.a2, g.N, g.H, g.W, g.CiP, g.CoP, 3, g.stride, 1, c->precise, stream));
    TRY(conv_forward(s.a2, s.wf2, s.bp2, s.c2, nullptr, g.N, g.OH, g.OW, g.CoP, g.CoP, 3, 1, 1, c->precise, stream));
    const float* skip = x;
    if (g.down) {
      TRY(pvae_conv_weight_prep(c->skip_w, c->skip_ln, c->skip_b, s.wfs, s.wds, s.bps, c->Co, c->Ci, g.CoP, g.CiP, g.ks * g.ks, g.groups,
                                1, stream));
      TRY(conv_forward(x_act, s.wfs, s.bps, s.skip, nullptr, g.N, g.H, g.W, g.CiP, g.CoP, g.ks, 2, 0, c->precise, stream));
      skip = s.skip;
    }
    return pvae_se_fwd(s.c2, skip, c->se_w1, c->se_b1, c->se_w2, c->se_b2, c->eps_res, y, y_act, s.pool, s.hid, s.gate, g.N, g.OH, g.OW,
                       g.CoP, c->Co, g.hd, 0, stream);
  }
  if (g.up) {
    TRY(pvae_upsample_fwd(x_act, s.au, g.N, g.H, g.W, g.OH, g.OW, g.CiP, stream));
    TRY(conv_forward(s.au, s.wf1, s.bp1, s.c2, nullptr, g.N, g.OH, g.OW, g.CiP, g.CoP, 3, 1, 1, c->precise, stream));
    TRY(pvae_conv_weight_prep(c->skip_w, c->skip_ln, c->skip_b, s.wfs, s.wds, s.bps, c->Co, c->Ci, g.CoP, g.CiP, 1, 1, 1, stream));
    TRY(conv_forward(x, s.wfs, s.bps, s.skip, nullptr, g.N, g.H, g.W, g.CiP, g.CoP, 1, 1, 0, c->precise, stream));
    return pvae_se_fwd(s.c2, s.skip, c->se_w1, c->se_b1, c->se_w2, c->se_b2, c->eps_res, y, y_act, s.pool, s.hid, s.gate, g.N, g.OH, g.OW,
                       g.CoP, c->Co, g.hd, 1, stream);
  }
  TRY(conv_forward(x_act, s.wf1, s.bp1, s.c2, nullptr, g.N, g.H, g.W, g.CiP, g.CoP, 3, 1, 1, c->precise, stream));
  return pvae_se_fwd(s.c2, x, c->se_w1, c->se_b1, c->se_w2, c->se_b2, c->eps_res, y, y_act, s.pool, s.hid, s.gate, g.N, g.OH, g.OW, g.CoP,
                     c->Co, g.hd, 0, stream);
}

extern "C" int pvae_cell_bwd(const pvae_cell* c, const float* x, const float* x_act, const float* g_y, float* g_x, float* saved,
                             float* scratch, void* stream) {
  TRY(check_cell(c, "pvae_cell_bwd"));
  if (!x || !g_y || !g_x || !saved || !scratch) return fail(PVAE_ERR_INVALID, "pvae_cell_bwd: null pointer");
  if (!c->g_w1 || !c->g_ln1 || !c->g_b1 || !c->g_se_w1 || !c->g_se_b1 || !c->g_se_w2 || !c->g_se_b2)
    return fail(PVAE_ERR_INVALID, "pvae_cell_bwd: null gradient pointer");
  const Geometry g = geometry(*c);
  const Saved s = carve_saved(*c, saved, x_act == nullptr);
  const Scratch w = carve_scratch(*c, scratch);
  if (x_act == nullptr) x_act = s.x_act;
  const int pr = c->precise;
  // Critical path (caller's stream): se_bwd -> dgrad(s) -> g_x.  Everything that only produces parameter gradients
  // (SE weights, wgrad + split reduction + weight-norm backward of every convolution) goes to the side stream as soon as
  // its inputs exist, and is joined before returning.  Same kernels either way: results do not depend on the overlap.
  CellSide* side = g_cell_overlap ? cell_side() : nullptr;
  void* s2 = side ? static_cast<void*>(side->stream) : stream;
  TRY(pvae_se_bwd(g_y, s.c2, s.gate, s.hid, c->se_w1, c->se_w2, c->eps_res, w.g_c2, w.d2, w.d1, g.N, g.OH * g.OW, g.CoP, c->Co, g.hd, stream));
  TRY(fork_to(side, 0, stream));
  TRY(pvae_se_param_grads(w.d2, s.hid, w.d1, s.pool, c->g_se_w2, c->g_se_b2, c->g_se_w1, c->g_se_b1, w.rws, g.N, c->Co, g.hd, s2));
  if (!g.dec) {
    if (!c->g_w2 || !c->g_ln2 || !c->g_b2) return fail(PVAE_ERR_INVALID, "pvae_cell_bwd: null gradient pointer");
    if (g.down && (!c->g_skip_w || !c->g_skip_ln || !c->g_skip_b)) return fail(PVAE_ERR_INVALID, "pvae_cell_bwd: null skip gradient pointer");
    TRY(conv_wgrad(s.a2, w.g_c2, c->w2, c->ln2, c->g_w2, c->g_ln2, c->g_b2, w.ws, g.N, g.OH, g.OW, g.CoP, g.CoP, g.OH, g.OW, 3, 1, 1, 1, c->Co,
                   c->Co, pr, s2));
    if (g.down)
      TRY(conv_wgrad(x_act, g_y, c->skip_w, c->skip_ln, c->g_skip_w, c->g_skip_ln, c->g_skip_b, w.ws, g.N, g.H, g.W, g.CiP, g.CoP, g.OH, g.OW,
                     g.ks, 2, 0, g.groups, c->Ci, c->Co, pr, s2));
    TRY(conv_dgrad(w.g_c2, s.wd2, w.g_c1, g.N, g.OH, g.OW, g.CoP, g.OH, g.OW, g.CoP, 3, 1, 1, nullptr, s.c1, nullptr, pr, stream));
    TRY(fork_to(side, 1, stream));
    TRY(conv_wgrad(x_act, w.g_c1, c->w1, c->ln1, c->g_w1, c->g_ln1, c->g_b1, w.ws, g.N, g.H, g.W, g.CiP, g.CoP, g.OH, g.OW, 3, g.stride, 1, 1,
                   c->Ci, c->Co, pr, s2));
    if (!g.down) {
      TRY(conv_dgrad(w.g_c1, s.wd1, g_x, g.N, g.H, g.W, g.CiP, g.OH, g.OW, g.CoP, 3, 1, 1, nullptr, x, g_y, pr, stream));
    } else {
      TRY(conv_dgrad(w.g_c1, s.wd1, w.g_a, g.N, g.H, g.W, g.CiP, g.OH, g.OW, g.CoP, 3, 2, 1, nullptr, nullptr, nullptr, pr, stream));
      TRY(conv_dgrad(g_y, s.wds, g_x, g.N, g.H, g.W, g.CiP, g.OH, g.OW, g.CoP, g.ks, 2, 0, w.g_a, x, nullptr, pr, stream));
    }
    return join_from(side, stream);
  }
  if (g.up) {
    if (!c->g_skip_w || !c->g_skip_ln || !c->g_skip_b) return fail(PVAE_ERR_INVALID, "pvae_cell_bwd: null skip gradient pointer");
    TRY(conv_wgrad(s.au, w.g_c2, c->w1, c->ln1, c->g_w1, c->g_ln1, c->g_b1, w.ws, g.N, g.OH, g.OW, g.CiP, g.CoP, g.OH, g.OW, 3, 1, 1, 1, c->Ci,
                   c->Co, pr, s2));
    TRY(pvae_upsample_bwd(g_y, nullptr, w.g_s, g.N, g.H, g.W, g.OH, g.OW, g.CoP, stream));  // skip branch, small map
    TRY(fork_to(side, 1, stream));
    TRY(conv_wgrad(x, w.g_s, c->skip_w, c->skip_ln, c->g_skip_w, c->g_skip_ln, c->g_skip_b, w.ws, g.N, g.H, g.W, g.CiP, g.CoP, g.H, g.W, 1, 1,
                   0, 1, c->Ci, c->Co, pr, s2));
    TRY(conv_dgrad(w.g_c2, s.wd1, w.g_au, g.N, g.OH, g.OW, g.CiP, g.OH, g.OW, g.CoP, 3, 1, 1, nullptr, nullptr, nullptr, pr, stream));
    TRY(pvae_upsample_bwd(w.g_au, x, w.t, g.N, g.H, g.W, g.OH, g.OW, g.CiP, stream));      // through SiLU -> up-sampling
    TRY(conv_dgrad(w.g_s, s.wds, g_x, g.N, g.H, g.W, g.CiP, g.H, g.W, g.CoP, 1, 1, 0, nullptr, nullptr, w.t, pr, stream));
    return join_from(side, stream);
  }
  TRY(conv_wgrad(x_act, w.g_c2, c->w1, c->ln1, c->g_w1, c->g_ln1, c->g_b1, w.ws, g.N, g.H, g.W, g.CiP, g.CoP, g.OH, g.OW, 3, 1, 1, 1, c->Ci,
                 c->Co, pr, s2));
  TRY(conv_dgrad(w.g_c2, s.wd1, g_x, g.N, g.H, g.W, g.CiP, g.OH, g.OW, g.CoP, 3, 1, 1, nullptr, x, g_y, pr, stream));
  return join_from(side, stream);
}

extern "C" int pvae_set_cell_overlap(int on) {
  g_cell_overlap = on != 0;
  return PVAE_OK;
}
