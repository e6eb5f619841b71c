// densebox_b200 — the network engine: owns the HBM layout of one DenseBox{,LM,LMLOC} replica inside a caller-provided
// workspace and sequences the kernels for forward (DenseBox.py:180-228 / :412-473 / :674-738), the fused loss and
// the hand-written backward (the autograd graph of DenseBox.py:2925), plus the fused SGD step (:2926).
//
// HBM layout (all activations NHWC bf16, 64-byte aligned regions carved from the workspace):
//   col0[N,H,W/2,64] (27 taps + constant 1 of two adjacent pixels per row) -> a11,a12[N,H,W,64] -> p1 -> a21,a22[.,128] -> p2 -> a31,a32[.,256]
//   fusion[N,H/4,W/4,768] = [ upsample(conv4_4) : 512 | conv3_4 : 256 ]   (torch.cat is free: producers write here)
//   p3 -> a41..a44[.,512];  hd[N,H/4,W/4,512*heads] (post-dropout), head_out fp32 [N,H/4,W/4,16|32]
//   refine: rp[.,H/8,W/8,64] -> r1 -> r2 -> rup[N,H/4,W/4,64] -> rf_out fp32 [.,16]
//   one gradient buffer per activation (d_*), parameters as flat fp32 master / grad / momentum + bf16 GEMM copies.
#include "dbx_common.h"
#include <stdio.h>
#include <string.h>
#include <string>
#include <vector>

namespace dbx {

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
static int round_up(int v, int a) { return (v + a - 1) / a * a; }

struct Group {  // one GEMM weight matrix, K-major [rows][ld]
  std::string name;
  int rows = 0, ld = 0, T = 1, cin_pad = 0, kpad = 0;
  size_t w_off = 0, b_off = 0, wd_off = 0;
  bool need_dgrad = false;
};
struct Param {  // one torch tensor pair (weight, bias) placed inside a group
  std::string name;
  int cout, cin, R, S;
  int grp;
  long rowK, kK;
  int cin_pad;
};

struct Buf { std::string name; size_t off, bytes; };

#define DBX_TRY(expr) do { int _rc = (expr); if (_rc) return _rc; } while (0)
#define DBX_K(tag, fl, expr) do { prof_begin(tag, fl, st); int _rc = (expr); prof_end(st); if (_rc) return _rc; } while (0)


struct Net {
  int variant, N, H, W, train;
  // conv1_1 "pairs" layout (default): col0 holds 32 channels per pixel, i.e. one 128-byte GEMM row = two adjacent
  // pixels; conv1_1 is the 128 x 64 matrix [[W 0] [0 W]] producing a11 viewed as [N,H,W/2,128].  Halves the HBM
  // traffic of col0 (im2col write, conv1_1 read, wgrad read) and the constant-1 tap yields the bias gradient inside
  // the weight gradient.  DBX_CONV1_PAIRS=0 at creation: the 64-channel layout (A/B, parity reference).
  bool pairs = true;
  int nh, HC;            // heads, head_out channels
  int ch_start[5];       // rows of conv5_2 group owned by head h
  std::vector<Group> groups;
  std::vector<Param> params;
  std::vector<Buf> bufs;
  char* ws = nullptr;
  size_t ws_bytes = 0;
  size_t flat_n = 0;     // fp32 elements in the flat parameter buffer (weights then biases)
  size_t mid_off = 0, tail_off = 0, bias_off = 0;  // bucket boundaries of the flat buffers (see build())
  size_t wd_n = 0;       // bf16 elements in the dgrad-layout buffer
  bool forward_done = false, loss_done = false;
  int sgd_steps = 0;
  // ---- per-launch accounting: every kernel launch of the engine goes through DBX_K (name, algorithmic FLOPs)
  struct Rec { std::string tag; double flops; cudaEvent_t e0, e1; };
  std::vector<Rec> recs;
  bool profiling = false;
  long long launches = 0;
  void prof_begin(const char* tag, double flops, cudaStream_t st) {
    ++launches;
    if (!profiling) return;
    Rec r; r.tag = tag; r.flops = flops;
    cudaEventCreate(&r.e0); cudaEventCreate(&r.e1);
    cudaEventRecord(r.e0, st);
    recs.push_back(r);
  }
  void prof_end(cudaStream_t st) { if (profiling) cudaEventRecord(recs.back().e1, st); }
  void prof_clear() { for (auto& r : recs) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); } recs.clear(); }
  double macs_of(const char* grp) const {  // algorithmic (unpadded) multiply-adds per output pixel
    const int gid = group_id(grp);
    double m = 0;
    for (auto& p : params) if (p.grp == gid) m += (double)p.cout * p.cin * p.R * p.S;
    return m;
  }
  static double pixels(const Act& a) { return (double)a.N * a.H * a.W; }

  // ---- layout
  size_t cursor = 0;
  size_t add_buf(const char* name, size_t bytes) {
    cursor = align_up(cursor, 256);
    bufs.push_back({name, cursor, bytes});
    cursor += bytes;
    return bufs.back().off;
  }
  void* buf(const char* name, size_t* bytes = nullptr) const {
    for (auto& b : bufs)
      if (b.name == name) { if (bytes) *bytes = b.bytes; return ws ? ws + b.off : (void*)b.off; }
    return nullptr;
  }
  int group_id(const char* name) const {
    for (size_t i = 0; i < groups.size(); ++i) if (groups[i].name == name) return (int)i;
    return -1;
  }
  int param_id(const char* name) const {
    for (size_t i = 0; i < params.size(); ++i) if (params[i].name == name) return (int)i;
    return -1;
  }
  Act act(const char* name, int h, int w, int C, int cs = 0, int coff = 0) const {
    Act a; a.ptr = buf(name); a.N = N; a.H = h; a.W = w; a.C = C; a.cs = cs ? cs : C; a.coff = coff; return a;
  }
  float* W32() const { return (float*)buf("w32"); }
  float* G32() const { return (float*)buf("g32"); }
  float* V32() const { return (float*)buf("v32"); }
  __nv_bfloat16* WK() const { return (__nv_bfloat16*)buf("wk"); }
  __nv_bfloat16* WD() const { return (__nv_bfloat16*)buf("wd"); }

  void add_group(const char* name, int rows, int T, int cin_pad, bool dgrad) {
    Group g; g.name = name; g.rows = rows; g.T = T; g.cin_pad = cin_pad; g.ld = T * cin_pad;
    g.kpad = round_up(rows, 64); g.need_dgrad = dgrad;
    groups.push_back(g);
  }
  void add_param(const char* name, int cout, int cin, int R, int S, const char* grp, long rowK, long kK, int cin_pad) {
    params.push_back({name, cout, cin, R, S, group_id(grp), rowK, kK, cin_pad});
  }

  void build(int variant_, int N_, int H_, int W_, int train_) {
    variant = variant_; N = N_; H = H_; W = W_; train = train_;
    { const char* e = ab_env("DBX_CONV1_PAIRS"); pairs = !(e && e[0] == '0') && (W % 2 == 0); }
    nh = variant == 0 ? 2 : (variant == 1 ? 3 : 4);
    HC = variant == 2 ? 32 : 16;
    const int starts[5] = {0, 1, 5, 9, 17};
    for (int i = 0; i < 5; ++i) ch_start[i] = starts[i];
    // ---- GEMM groups
    add_group("conv1_1", pairs ? 128 : 64, 1, 64, false);
    add_group("conv1_2", 64, 9, 64, true);
    add_group("conv2_1", 128, 9, 64, true);
    add_group("conv2_2", 128, 9, 128, true);
    add_group("conv3_1", 256, 9, 128, true);
    add_group("conv3_2", 256, 9, 256, true);
    add_group("conv3_4", 256, 9, 256, true);
    add_group("conv4_1", 512, 9, 256, true);
    add_group("conv4_2", 512, 9, 512, true);
    add_group("conv4_3", 512, 9, 512, true);
    add_group("conv4_4", 512, 9, 512, true);
    add_group("heads1", 512 * nh, 1, 768, true);
    add_group("heads2", HC, 1, 512 * nh, true);
    if (variant >= 1) {
      add_group("conv6_1_det", 64, 9, 64, true);
      add_group("conv6_2_det", 64, 25, 64, true);
      add_group("conv6_3_det", 16, 1, 64, true);
    }
    // Flat parameter order = completion order of the backward stages, so that every gradient bucket is one contiguous
    // range: [0, mid_off) filters of conv4_1 .. heads (+ refine), complete after stage 0; [mid_off, tail_off) filters
    // of the conv3 block (stage 1); [tail_off, flat_n) filters of conv1_1 .. conv2_2 and EVERY bias (stage 2).  A
    // data-parallel caller all-reduces bucket k while stage k + 1 computes; only the last, 1 MB bucket is exposed.
    size_t off = 0, wd = 0;
    const size_t g4 = (size_t)group_id("conv4_1"), g3 = (size_t)group_id("conv3_1");
    for (size_t i = g4; i < groups.size(); ++i) { groups[i].w_off = off; off += align_up((size_t)groups[i].rows * groups[i].ld, 64); }
    mid_off = off;
    for (size_t i = g3; i < g4; ++i) { groups[i].w_off = off; off += align_up((size_t)groups[i].rows * groups[i].ld, 64); }
    tail_off = off;
    for (size_t i = 0; i < g3; ++i) { groups[i].w_off = off; off += align_up((size_t)groups[i].rows * groups[i].ld, 64); }
    bias_off = off;
    for (auto& g : groups) { g.b_off = off; off += align_up((size_t)g.rows, 64); }
    flat_n = off;
    for (auto& g : groups)
      if (g.need_dgrad) { g.wd_off = wd; wd += align_up((size_t)g.cin_pad * g.T * g.kpad, 64); }
    wd_n = wd;
    // ---- torch tensors -> placement
    add_param("conv1_1", 64, 3, 3, 3, "conv1_1", 0, 0, 3);
    add_param("conv1_2", 64, 64, 3, 3, "conv1_2", 0, 0, 64);
    add_param("conv2_1", 128, 64, 3, 3, "conv2_1", 0, 0, 64);
    add_param("conv2_2", 128, 128, 3, 3, "conv2_2", 0, 0, 128);
    add_param("conv3_1", 256, 128, 3, 3, "conv3_1", 0, 0, 128);
    add_param("conv3_2", 256, 256, 3, 3, "conv3_2", 0, 0, 256);
    add_param("conv3_4", 256, 256, 3, 3, "conv3_4", 0, 0, 256);
    add_param("conv4_1", 512, 256, 3, 3, "conv4_1", 0, 0, 256);
    add_param("conv4_2", 512, 512, 3, 3, "conv4_2", 0, 0, 512);
    add_param("conv4_3", 512, 512, 3, 3, "conv4_3", 0, 0, 512);
    add_param("conv4_4", 512, 512, 3, 3, "conv4_4", 0, 0, 512);
    const char* hn[4] = {"det", "loc", "landmark", "lmloc"};
    const int hc[4] = {1, 4, 4, 8};
    for (int h = 0; h < nh; ++h) {
      add_param((std::string("conv5_1_") + hn[h]).c_str(), 512, 768, 1, 1, "heads1", 512L * h, 0, 768);
      add_param((std::string("conv5_2_") + hn[h]).c_str(), hc[h], 512, 1, 1, "heads2", ch_start[h], 512L * h, 512 * nh);
    }
    if (variant >= 1) {
      add_param("conv6_1_det", 64, 5, 3, 3, "conv6_1_det", 0, 0, 64);
      add_param("conv6_2_det", 64, 64, 5, 5, "conv6_2_det", 0, 0, 64);
      add_param("conv6_3_det", 1, 64, 1, 1, "conv6_3_det", 0, 0, 64);
    }
    // ---- workspace regions
    const size_t px1 = (size_t)N * H * W, px2 = px1 / 4, px4 = px1 / 16, px8 = px1 / 64;
    add_buf("w32", flat_n * 4);
    add_buf("wk", flat_n * 2);
    add_buf("wd", wd_n * 2);
    add_buf("rng", 256);      // {seed, offset} (2 x u64) of the in-epilogue dropout draw
    add_buf("scalars", 256);  // [0] loss f32, [1] counter u32, [2..3] info i32, [16..] loss_partial
    add_buf("loss_partial", (size_t)N * 4);
    add_buf("col0", px1 * (pairs ? 32 : 64) * 2);
    add_buf("a11", px1 * 64 * 2); add_buf("a12", px1 * 64 * 2); add_buf("p1", px2 * 64 * 2);
    add_buf("a21", px2 * 128 * 2); add_buf("a22", px2 * 128 * 2); add_buf("p2", px4 * 128 * 2);
    add_buf("a31", px4 * 256 * 2); add_buf("a32", px4 * 256 * 2); add_buf("fusion", px4 * 768 * 2);
    add_buf("p3", px8 * 256 * 2);
    add_buf("a41", px8 * 512 * 2); add_buf("a42", px8 * 512 * 2); add_buf("a43", px8 * 512 * 2);
    add_buf("a44", px8 * 512 * 2);
    add_buf("hd", px4 * 512 * nh * 2);
    add_buf("head_out", px4 * HC * 4);
    add_buf("wfold", (size_t)HC * 768 * 2);   // inference: conv5_2 o conv5_1 as one [HC][768] matrix (heads_fold)
    add_buf("bfold", 256);
    const int h4 = H / 4, w4 = W / 4, h8 = H / 8, w8 = W / 8;
    if (variant >= 1) {
      add_buf("rp", px8 * 64 * 2);
      add_buf("r1", (size_t)N * (h8 - 2) * (w8 - 2) * 64 * 2);
      add_buf("r2", (size_t)N * (h8 - 6) * (w8 - 6) * 64 * 2);
      add_buf("rup", px4 * 64 * 2);
      add_buf("rf_out", px4 * 16 * 4);
    }
    if (train) {
      add_buf("g32", flat_n * 4);
      add_buf("g11p", 128 * 64 * 4);  // conv1_1 weight gradient of one backward pass in the pairs layout, before the fold
      add_buf("v32", flat_n * 4);
      add_buf("drop", px4 * 512 * nh * 2);
      add_buf("pi1", px2 * 8 * 2);    // arg-max maps of pool1 / pool2: u16 per (pooled pixel, 8-channel vector)
      add_buf("pi2", px4 * 16 * 2);
      add_buf("d_head", px4 * 64 * 2);
      add_buf("d_hd", px4 * 512 * nh * 2);
      add_buf("d_fusion", px4 * 768 * 2);
      add_buf("d_a44", px8 * 512 * 2); add_buf("d_a43", px8 * 512 * 2); add_buf("d_a42", px8 * 512 * 2);
      add_buf("d_a41", px8 * 512 * 2); add_buf("d_p3", px8 * 256 * 2);
      add_buf("d_a34", px4 * 256 * 2); add_buf("d_a32", px4 * 256 * 2); add_buf("d_a31", px4 * 256 * 2);
      add_buf("d_p2", px4 * 128 * 2);
      add_buf("d_a22", px2 * 128 * 2); add_buf("d_a21", px2 * 128 * 2); add_buf("d_p1", px2 * 64 * 2);
      add_buf("d_a12", px1 * 64 * 2); add_buf("d_a11", px1 * 64 * 2);
      if (variant >= 1) {
        add_buf("d_rf", px4 * 64 * 2);
        add_buf("d_rup", px4 * 64 * 2);
        add_buf("d_r2", (size_t)N * (h8 - 6) * (w8 - 6) * 64 * 2);
        add_buf("d_r1", (size_t)N * (h8 - 2) * (w8 - 2) * 64 * 2);
        add_buf("d_rp", px8 * 64 * 2);
      }
    }
    (void)h4; (void)w4;
    ws_bytes = align_up(cursor, 256);
  }

  const __nv_bfloat16* wk_of(const char* g) const { return WK() + groups[group_id(g)].w_off; }
  const __nv_bfloat16* wd_of(const char* g) const { return WD() + groups[group_id(g)].wd_off; }
  const float* bias_of(const char* g) const { return W32() + groups[group_id(g)].b_off; }
  float* gw_of(const char* g) const { return G32() + groups[group_id(g)].w_off; }
  float* gb_of(const char* g) const { return G32() + groups[group_id(g)].b_off; }

  // ---- one-time initialisation of the workspace (zero padding lanes, counters)
  int init(cudaStream_t st) {
    cudaError_t e = cudaMemsetAsync(ws, 0, ws_bytes, st);
    return (int)e;
  }

  // ---- parameters
  int set_param(const char* name, int is_bias, const float* src, long s_co, long s_ci, long s_r, long s_s,
                cudaStream_t st) {
    const int id = param_id(name);
    if (id < 0 || !src) return DBX_ERR_ARG;
    const Param& p = params[id];
    const Group& g = groups[p.grp];
    const bool dup = pairs && std::string(name) == "conv1_1";  // second diagonal block: rows 64.., columns 32..
    if (is_bias) {
      cudaError_t e = cudaMemcpy2DAsync(W32() + g.b_off + p.rowK, 4, src, (size_t)s_co * 4, 4, p.cout,
                                        cudaMemcpyDeviceToDevice, st);
      if (e == cudaSuccess && dup)
        e = cudaMemcpy2DAsync(W32() + g.b_off + 64, 4, src, (size_t)s_co * 4, 4, p.cout, cudaMemcpyDeviceToDevice, st);
      return (int)e;
    }
    int rc = pack_weights(src, p.cout, p.cin, p.R, p.S, s_co, s_ci, s_r, s_s, WK() + g.w_off, g.ld, p.rowK, p.kK,
                          p.cin_pad, nullptr, 0, 0, 0, 0, W32() + g.w_off, st);
    if (!rc && dup)
      rc = pack_weights(src, p.cout, p.cin, p.R, p.S, s_co, s_ci, s_r, s_s, WK() + g.w_off, g.ld, 64, 32, p.cin_pad,
                        nullptr, 0, 0, 0, 0, W32() + g.w_off, st);
    return rc;
  }
  int get_tensor(const char* name, int is_bias, int which, float* dst, long s_co, long s_ci, long s_r, long s_s,
                 cudaStream_t st) {
    const int id = param_id(name);
    if (id < 0 || !dst) return DBX_ERR_ARG;
    if (which == 1 && !train) return DBX_ERR_STATE;
    const Param& p = params[id];
    const Group& g = groups[p.grp];
    const float* base = which == 0 ? W32() : G32();
    if (is_bias) {
      cudaError_t e = cudaMemcpy2DAsync(dst, (size_t)s_co * 4, base + g.b_off + p.rowK, 4, 4, p.cout,
                                        cudaMemcpyDeviceToDevice, st);
      return (int)e;
    }
    return unpack_weights(base + g.w_off, g.ld, p.rowK, p.kK, p.cin_pad, dst, p.cout, p.cin, p.R, p.S, s_co, s_ci,
                          s_r, s_s, st);
  }
  // every listed (weight, bias) pair in one launch: mode 0 = torch -> w32 + wk, 1 = w32 -> torch, 2 = g32 -> torch
  int xfer_params(int mode, int n, const char* const* names, void* const* w_ptrs, const long* w_strides,
                  void* const* b_ptrs, const long* b_strides, cudaStream_t st) {
    if (mode < 0 || mode > 2 || n < 1 || n > ParamXfer::kMax || !names || !w_ptrs || !w_strides || !b_ptrs || !b_strides)
      return DBX_ERR_ARG;
    if (mode == 2 && !train) return DBX_ERR_STATE;
    ParamXfer t{};
    long long elems = 0;
    for (int k = 0; k < n; ++k) {
      const int id = names[k] ? param_id(names[k]) : -1;
      if (id < 0 || !w_ptrs[k] || !b_ptrs[k]) return DBX_ERR_ARG;
      const Param& p = params[id];
      const Group& g = groups[p.grp];
      ParamXfer::E& e = t.e[k];
      e.w = (float*)w_ptrs[k]; e.s_co = w_strides[4 * k]; e.s_ci = w_strides[4 * k + 1];
      e.s_r = w_strides[4 * k + 2]; e.s_s = w_strides[4 * k + 3];
      e.b = (float*)b_ptrs[k]; e.s_b = b_strides[k];
      e.co = p.cout; e.ci = p.cin; e.R = p.R; e.S = p.S;
      e.w_off = (long long)g.w_off; e.b_off = (long long)g.b_off + p.rowK;
      e.ld = g.ld; e.rowK = p.rowK; e.kK = p.kK; e.cin_pad = p.cin_pad;
      e.dup = (pairs && p.name == "conv1_1") ? 1 : 0;
      e.elem0 = elems;
      elems += (long long)p.cout * p.cin * p.R * p.S + p.cout;
    }
    t.n = n; t.total = elems;
    DBX_K("params_xfer", 0.0, params_xfer(t, mode == 0 ? 0 : 1, mode == 2 ? G32() : W32(), WK(), st));
    return DBX_OK;
  }
  // head maps as the NCHW fp32 tensors the reference modules return (any pointer may be null)
  int get_outputs(float* score, float* loc, float* lm, float* lmloc, float* rf, cudaStream_t st) {
    if (!forward_done) return DBX_ERR_STATE;
    if ((lm && variant < 1) || (rf && variant < 1) || (lmloc && variant < 2)) return DBX_ERR_ARG;
    DBX_K("heads_to_nchw", 0.0, heads_to_nchw((const float*)buf("head_out"), HC, variant >= 1 ? (const float*)buf("rf_out") : nullptr,
                                              16, N, (H / 4) * (W / 4), score, loc, lm, lmloc, rf, st));
    return DBX_OK;
  }
  // d(loss)/d(outputs) as autograd delivers them (fp32 NCHW, null = zero) -> "d_head" / "d_rf"
  int set_output_grads(const float* g_score, const float* g_loc, const float* g_lm, const float* g_lmloc,
                       const float* g_rf, cudaStream_t st) {
    if (!train) return DBX_ERR_STATE;
    if ((g_lm && variant < 1) || (g_rf && variant < 1) || (g_lmloc && variant < 2)) return DBX_ERR_ARG;
    DBX_K("nchw_to_head_grads", 0.0, nchw_to_head_grads(g_score, g_loc, g_lm, g_lmloc, g_rf, N, (H / 4) * (W / 4), buf("d_head"),
                                                        variant >= 1 ? buf("d_rf") : nullptr, st));
    return DBX_OK;
  }
  // bf16 dgrad-layout copies of every filter (call after the weights changed, before backward)
  bool dgrad_stale = false;   // set by sgd(): the flipped filters are rebuilt on the side stream during forward
  bool dgrad_pending = false; // transposes in flight on the side stream
  int refresh_dgrad(cudaStream_t st) {
    dgrad_stale = false;
    TransposeBatch tb{};
    int blocks = 0;
    for (auto& g : groups) {
      if (!g.need_dgrad) continue;
      if (tb.n == TransposeBatch::kMax) return DBX_ERR_ARG;
      TransposeBatch::G& e = tb.g[tb.n++];
      e.wk_off = (long long)g.w_off; e.wd_off = (long long)g.wd_off;
      e.rows = g.rows; e.T = g.T; e.cin_pad = g.cin_pad; e.kpad = g.kpad; e.block0 = blocks;
      blocks += ((g.cin_pad + 31) / 32) * ((g.rows + 31) / 32) * g.T;
    }
    DBX_K("transpose_dgrad", 0.0, transpose_dgrad_multi(WK(), WD(), tb, blocks, st));  // every filter in one launch
    return DBX_OK;
  }

  // ---- forward
  int conv(const Act& x, const char* grp, int R, int pad, const Act& out, bool relu, const void* aux, int aux_cs,
           int aux_mode, bool fp32, int block_n, cudaStream_t st) {
    ConvEpilogue e;
    e.bias = bias_of(grp); e.relu = relu ? 1 : 0; e.aux = aux; e.aux_cs = aux_cs; e.aux_mode = aux_mode;
    e.out_fp32 = fp32 ? 1 : 0;
    if (aux_mode && x.C * R * R <= 1024) e.epi_bufs = 4;  // short K with a mask (heads1): 3 stages + 4 boxes
    DBX_K((std::string("fprop:") + grp).c_str(), 2.0 * pixels(out) * macs_of(grp),
          conv_fprop(x, wk_of(grp), R, R, pad, out, e, block_n, st));
    return DBX_OK;
  }
  // conv + bias + ReLU + 2x2 max-pool in one launch (ConvEpilogue::pool_out): `pooled` and the arg-max map are
  // written, the full-resolution activation `full` only when the launch has to fall back to the stand-alone pool.
  int conv_pool(const Act& x, const char* grp, const Act& full, const Act& pooled, void* idx, cudaStream_t st,
                bool keep_full = false) {
    ConvEpilogue e;
    e.bias = bias_of(grp); e.relu = 1;
    e.pool_out = pooled.ptr; e.pool_cs = pooled.cs; e.pool_coff = pooled.coff; e.pool_idx = idx;
    e.pool_keep_full = keep_full ? 1 : 0;
    DBX_K((std::string("fprop:") + grp).c_str(), 2.0 * pixels(full) * macs_of(grp),
          conv_fprop(x, wk_of(grp), 3, 3, 1, full, e, 0, st));
    return DBX_OK;
  }
  // bias_below: the group whose bias gradient is the column sum of dx (dx = dZ of that layer): produced by this
  // launch's epilogue (ConvEpilogue::colsum) instead of a separate pass over dx; the matching wgrad() call then
  // passes bias_done = true.
  int dgrad(const Act& dy, const char* grp, int R, int pad, const Act& dx, const Act* relu_y, cudaStream_t st,
            const char* bias_below = nullptr) {
    ConvEpilogue e;
    if (relu_y) { e.aux = relu_y->ptr; e.aux_cs = relu_y->cs; e.aux_coff = relu_y->coff; e.aux_mode = 1; }
    if (bias_below) e.colsum = gb_of(bias_below);
    DBX_K((std::string("dgrad:") + grp).c_str(), 2.0 * pixels(dy) * macs_of(grp),
          conv_fprop(dy, wd_of(grp), R, R, R - 1 - pad, dx, e, 0, st));
    return DBX_OK;
  }
  // Weight and bias gradients run on a side stream.  The critical chain of the backward pass is dgrad -> pool /
  // upsample backward -> dgrad ...; wgrad(l) only needs dZ(l).  With wgrads queued on a second stream the HBM-bound
  // element-wise kernels of the chain (and the colsum bias gradients, 8 KB smem) co-reside with a persistent tensor
  // kernel instead of leaving the tensor cores idle.
  cudaStream_t side = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  bool side_used = false;
  int wgrad(const Act& x, const Act& dy, const char* grp, int R, int pad, cudaStream_t st, bool bias_done = false) {
    if (side && !profiling) {
      DBX_TRY((int)cudaEventRecord(ev_fork, st));
      DBX_TRY((int)cudaStreamWaitEvent(side, ev_fork, 0));
      launches += bias_done ? 1 : 2;
      if (!bias_done) DBX_TRY(colsum(dy, gb_of(grp), side));
      DBX_TRY(conv_wgrad(x, dy, R, R, pad, gw_of(grp), 0, side));
      side_used = true;
      return DBX_OK;
    }
    if (!bias_done) DBX_K((std::string("colsum:") + grp).c_str(), 0.0, colsum(dy, gb_of(grp), st));
    DBX_K((std::string("wgrad:") + grp).c_str(), 2.0 * pixels(dy) * macs_of(grp),
          conv_wgrad(x, dy, R, R, pad, gw_of(grp), 0, st));
    return DBX_OK;
  }
  int join_side(cudaStream_t st) {
    if (!side_used) return DBX_OK;
    side_used = false;
    DBX_TRY((int)cudaEventRecord(ev_join, side));
    return (int)cudaStreamWaitEvent(st, ev_join, 0);
  }


  // dropout_mode: 0 = eval (identity); 1 = Philox draw inside the conv epilogues from (seed, offset); 2 = multiply
  // by the {0,2} bf16 mask the caller wrote into `drop`; 3 = like 1 but (seed, offset) are read from the `rng`
  // region as the caller left it (CUDA-graph replays: update the region between replays).
  unsigned long long rng_host[2] = {0, 0};
  // x_u8 != null: the batch as decoded image bytes (uint8 NHWC) + the 3 x 256 ToTensor/Normalize table (see
  // im2col3x3_c3_u8); otherwise x = fp32 NCHW, already normalised (the reference's forward argument).
  int forward(const float* x, int dropout_mode, unsigned long long seed, unsigned long long offset, cudaStream_t st,
              const unsigned char* x_u8 = nullptr, const float* lut = nullptr) {
    if (!x && !(x_u8 && lut)) return DBX_ERR_ARG;
    if (dropout_mode < 0 || dropout_mode > 3) return DBX_ERR_ARG;
    if (dropout_mode && !train) return DBX_ERR_STATE;
    const int h2 = H / 2, w2 = W / 2, h4 = H / 4, w4 = W / 4, h8 = H / 8, w8 = W / 8;
    Act col0 = act("col0", H, W, 64), a11 = act("a11", H, W, 64), a12 = act("a12", H, W, 64);
    Act p1 = act("p1", h2, w2, 64), a21 = act("a21", h2, w2, 128), a22 = act("a22", h2, w2, 128);
    Act p2 = act("p2", h4, w4, 128), a31 = act("a31", h4, w4, 256), a32 = act("a32", h4, w4, 256);
    Act fus = act("fusion", h4, w4, 768), fus_up = act("fusion", h4, w4, 512, 768, 0);
    Act a34 = act("fusion", h4, w4, 256, 768, 512);
    Act p3 = act("p3", h8, w8, 256), a41 = act("a41", h8, w8, 512), a42 = act("a42", h8, w8, 512);
    Act a43 = act("a43", h8, w8, 512), a44 = act("a44", h8, w8, 512);
    Act hd = act("hd", h4, w4, 512 * nh), ho = act("head_out", h4, w4, HC);
    if (dgrad_stale && side && !profiling) {  // overlap the filter re-layout with the forward pass
      DBX_TRY((int)cudaEventRecord(ev_fork, st));
      DBX_TRY((int)cudaStreamWaitEvent(side, ev_fork, 0));
      DBX_TRY(refresh_dgrad(side));
      DBX_TRY((int)cudaEventRecord(ev_join, side));
      dgrad_pending = true;
    }
    if (x_u8) DBX_K("im2col", 0.0, im2col3x3_c3_u8(x_u8, lut, col0.ptr, N, H, W, pairs ? 2 : 0, st));
    else DBX_K("im2col", 0.0, im2col3x3_c3(x, col0.ptr, N, H, W, pairs ? 2 : 0, st));
    if (pairs) {
      Act col0p = act("col0", H, W / 2, 64), a11p = act("a11", H, W / 2, 128);
      ConvEpilogue e;
      e.bias = bias_of("conv1_1"); e.relu = 1;
      DBX_K("fprop:conv1_1", 2.0 * pixels(a11) * macs_of("conv1_1"), conv_fprop(col0p, wk_of("conv1_1"), 1, 1, 0, a11p, e, 0, st));
    } else {
      DBX_TRY(conv(col0, "conv1_1", 1, 0, a11, true, nullptr, 0, 0, false, 0, st));
    }
    // pool1 / pool2 ride in the epilogues of conv1_2 / conv2_2: a12 / a22 never reach HBM (nothing else reads them:
    // the backward pass works from the pooled maps + arg-max maps).  DBX_POOL_FUSE=0 / DBX_POOL_IDX=0: separate kernels.
    bool pool_fuse = true;
    { const char* e1 = ab_env("DBX_POOL_FUSE"); const char* e2 = ab_env("DBX_POOL_IDX");
      if ((e1 && e1[0] == '0') || (e2 && e2[0] == '0')) pool_fuse = false; }
    if (pool_fuse) {
      DBX_TRY(conv_pool(a11, "conv1_2", a12, p1, train ? buf("pi1") : nullptr, st));
    } else {
      DBX_TRY(conv(a11, "conv1_2", 3, 1, a12, true, nullptr, 0, 0, false, 0, st));
      DBX_K("pool_fwd", 0.0, maxpool2x2_fwd(a12, p1, st, train ? buf("pi1") : nullptr));
    }
    DBX_TRY(conv(p1, "conv2_1", 3, 1, a21, true, nullptr, 0, 0, false, 0, st));
    if (pool_fuse) {
      DBX_TRY(conv_pool(a21, "conv2_2", a22, p2, train ? buf("pi2") : nullptr, st));
    } else {
      DBX_TRY(conv(a21, "conv2_2", 3, 1, a22, true, nullptr, 0, 0, false, 0, st));
      DBX_K("pool_fwd", 0.0, maxpool2x2_fwd(a22, p2, st, train ? buf("pi2") : nullptr));
    }
    DBX_TRY(conv(p2, "conv3_1", 3, 1, a31, true, nullptr, 0, 0, false, 0, st));
    DBX_TRY(conv(a31, "conv3_2", 3, 1, a32, true, nullptr, 0, 0, false, 0, st));
    if (pool_fuse) {  // conv3_3 skipped (:193-195); conv3_4 feeds the fusion buffer AND pool3
      DBX_TRY(conv_pool(a32, "conv3_4", a34, p3, nullptr, st, true));
    } else {
      DBX_TRY(conv(a32, "conv3_4", 3, 1, a34, true, nullptr, 0, 0, false, 0, st));
      DBX_K("pool_fwd", 0.0, maxpool2x2_fwd(a34, p3, st));
    }
    DBX_TRY(conv(p3, "conv4_1", 3, 1, a41, true, nullptr, 0, 0, false, 0, st));
    DBX_TRY(conv(a41, "conv4_2", 3, 1, a42, true, nullptr, 0, 0, false, 0, st));
    DBX_TRY(conv(a42, "conv4_3", 3, 1, a43, true, nullptr, 0, 0, false, 0, st));
    DBX_TRY(conv(a43, "conv4_4", 3, 1, a44, true, nullptr, 0, 0, false, 0, st));
    DBX_K("upsample_fwd", 0.0, upsample_bilinear_fwd(a44, fus_up, st));
    if (dropout_mode == 1) {
      rng_host[0] = seed; rng_host[1] = offset;
      DBX_TRY((int)cudaMemcpyAsync(buf("rng"), rng_host, 16, cudaMemcpyHostToDevice, st));
    }
    // Inference nets without dropout: the reference puts nothing but Dropout between conv5_1_* and conv5_2_*
    // (DenseBox.py:158-178), so the two 1x1 convolutions are ONE 768 -> HC matrix; the 512-channel hidden maps are
    // never computed, let alone stored.  Folded from the fp32 masters on every forward (a few microseconds).
    bool fold = !train && dropout_mode == 0;
    { const char* e = ab_env("DBX_HEADS_FOLD"); if (e && e[0] == '0') fold = false; }
    if (fold) {
      const Group& g1 = groups[group_id("heads1")];
      const Group& g2 = groups[group_id("heads2")];
      DBX_K("heads_fold", 0.0, heads_fold(W32() + g1.w_off, g1.ld, bias_of("heads1"), W32() + g2.w_off, g2.ld,
                                          bias_of("heads2"), ch_start, nh, HC, buf("wfold"), (float*)buf("bfold"), st));
      ConvEpilogue e;
      e.bias = (const float*)buf("bfold"); e.out_fp32 = 1;
      DBX_K("fprop:heads_folded", 2.0 * pixels(ho) * 768.0 * HC, conv_fprop(fus, buf("wfold"), 1, 1, 0, ho, e, 0, st));
    } else {
      ConvEpilogue e;
      e.bias = bias_of("heads1");
      if (dropout_mode == 2) { e.aux = buf("drop"); e.aux_cs = 512 * nh; e.aux_mode = 2; e.epi_bufs = 4; }
      else if (dropout_mode) { e.aux_mode = 3; e.rng = (const unsigned long long*)buf("rng"); e.rng_channels = 512 * nh; }
      DBX_K("fprop:heads1", 2.0 * pixels(hd) * macs_of("heads1"), conv_fprop(fus, wk_of("heads1"), 1, 1, 0, hd, e, 0, st));
    }
    if (!fold) DBX_TRY(conv(hd, "heads2", 1, 0, ho, false, nullptr, 0, 0, true, 0, st));
    if (variant >= 1) {
      Act rp = act("rp", h8, w8, 64), r1 = act("r1", h8 - 2, w8 - 2, 64), r2 = act("r2", h8 - 6, w8 - 6, 64);
      Act rup = act("rup", h4, w4, 64), rf = act("rf_out", h4, w4, 16);
      DBX_K("refine_pool_pack", 0.0, refine_pool_pack((const float*)ho.ptr, HC, rp.ptr, N, h4, w4, st));
      DBX_TRY(conv(rp, "conv6_1_det", 3, 0, r1, false, nullptr, 0, 0, false, 0, st));
      DBX_TRY(conv(r1, "conv6_2_det", 5, 0, r2, false, nullptr, 0, 0, false, 0, st));
      DBX_K("upsample_fwd", 0.0, upsample_bilinear_fwd(r2, rup, st));
      DBX_TRY(conv(rup, "conv6_3_det", 1, 0, rf, false, nullptr, 0, 0, true, 0, st));
    }
    drop_mode = dropout_mode == 2 ? 2 : (dropout_mode ? 3 : 0);
    forward_done = true;
    return DBX_OK;
  }
  const unsigned long long* count_slots = nullptr; int count_world = 0;  // data parallel: see count_exchange
  int drop_mode = 0;  // how the last forward dropped: 0 none, 2 mask buffer, 3 Philox in place

  // ---- loss (+ gradients w.r.t. the head outputs)
  int loss(const float* bbox, const float* vertices, const float* labels, const long long* rand_idx, int rand_stride,
           const long long* lm_rand_idx, float lambda_loc, float lambda_det, float lambda_lm, int global_pos,
           int global_batch, const int* global_pos_ptr, int clamp_lm, float* d_head_f32, float* d_rf_f32,
           unsigned char* mask_out, unsigned char* lm_mask_out, cudaStream_t st) {
    if (!forward_done) return DBX_ERR_STATE;
    if (H != 240 || W != 240) return DBX_ERR_ARG;  // the reference loss is defined on 60x60 maps only (:1379)
    LossParams p{};
    p.head = (const float*)buf("head_out"); p.HC = HC;
    p.rf = variant >= 1 ? (const float*)buf("rf_out") : nullptr; p.RC = 16;
    p.bbox = bbox; p.vertices = vertices; p.labels = labels;
    p.rand_idx = rand_idx; p.rand_stride = rand_stride; p.lm_rand_idx = lm_rand_idx;
    p.variant = variant; p.lambda_loc = lambda_loc; p.lambda_det = lambda_det; p.lambda_lm = lambda_lm;
    p.global_pos = global_pos; p.global_batch = global_batch; p.global_pos_ptr = global_pos_ptr;
    p.count_slots = count_slots; p.count_world = count_world;
    p.clamp_lm = clamp_lm; p.B = N;
    float* sc = (float*)buf("scalars");
    p.loss = sc; p.counter = (unsigned int*)(sc + 1); p.info = (int*)(sc + 2);
    p.loss_partial = (float*)buf("loss_partial");
    p.d_head = train ? (__nv_bfloat16*)buf("d_head") : nullptr;
    p.d_rf = (train && variant >= 1) ? (__nv_bfloat16*)buf("d_rf") : nullptr;
    p.d_head_f32 = d_head_f32; p.d_rf_f32 = d_rf_f32; p.mask_out = mask_out; p.lm_mask_out = lm_mask_out;
    DBX_K("loss", 0.0, loss_fwd_bwd(p, st));
    loss_done = true;
    return DBX_OK;
  }

  // ---- backward: d_head / d_rf (bf16, written by loss() or by the caller) -> parameter gradients in g32 (+=)
  // stage: -1 = everything; 0 = refine + heads + conv4 block; 1 = conv3 block; 2 = conv2 + conv1 blocks.  After
  // stage k the gradients of bucket k are complete (see grad_bucket): a data-parallel caller all-reduces bucket k
  // while stage k + 1 still computes.
  int backward(cudaStream_t st, int stage = -1) {
    if (!train || !forward_done) return DBX_ERR_STATE;
    if (stage < -1 || stage > 2) return DBX_ERR_ARG;
    if (dgrad_pending) { DBX_TRY((int)cudaStreamWaitEvent(st, ev_join, 0)); dgrad_pending = false; }
    if (dgrad_stale) DBX_TRY(refresh_dgrad(st));
    // bias gradients come out of the epilogue of the launch that produces dZ (DBX_FUSE_BIAS=0: stand-alone colsum)
    bool fuse = true;
    { const char* e = ab_env("DBX_FUSE_BIAS"); if (e && e[0] == '0') fuse = false; }
    bool pool_idx = true;
    { const char* e = ab_env("DBX_POOL_IDX"); if (e && e[0] == '0') pool_idx = false; }
    // The two short-K data gradients (conv2_2, conv1_2) are paced by their epilogue: the column sums cost them
    // +0.047 / +0.05 ms (measured, with or without shared-memory atomics) against 0.027 / 0.05 ms for the stand-alone
    // pass on the side stream, so their bias gradients stay with wgrad().  DBX_FUSE_BIAS_SHORT=1 fuses them too.
    bool fuse_short = false;
    { const char* e = ab_env("DBX_FUSE_BIAS_SHORT"); if (e && e[0] == '1') fuse_short = fuse; }
    const int h2 = H / 2, w2 = W / 2, h4 = H / 4, w4 = W / 4, h8 = H / 8, w8 = W / 8;
    Act col0 = act("col0", H, W, 64), a11 = act("a11", H, W, 64), a12 = act("a12", H, W, 64);
    Act p1 = act("p1", h2, w2, 64), a21 = act("a21", h2, w2, 128), a22 = act("a22", h2, w2, 128);
    Act p2 = act("p2", h4, w4, 128), a31 = act("a31", h4, w4, 256), a32 = act("a32", h4, w4, 256);
    Act fus = act("fusion", h4, w4, 768), a34 = act("fusion", h4, w4, 256, 768, 512);
    Act p3 = act("p3", h8, w8, 256), a41 = act("a41", h8, w8, 512), a42 = act("a42", h8, w8, 512);
    Act a43 = act("a43", h8, w8, 512), a44 = act("a44", h8, w8, 512);
    Act hd = act("hd", h4, w4, 512 * nh);
    Act d_head64 = act("d_head", h4, w4, 64), d_headC = act("d_head", h4, w4, HC, 64, 0);
    Act d_hd = act("d_hd", h4, w4, 512 * nh), d_fus = act("d_fusion", h4, w4, 768);
    Act d_fus_up = act("d_fusion", h4, w4, 512, 768, 0), d_fus_34 = act("d_fusion", h4, w4, 256, 768, 512);
    Act d_a44 = act("d_a44", h8, w8, 512), d_a43 = act("d_a43", h8, w8, 512), d_a42 = act("d_a42", h8, w8, 512);
    Act d_a41 = act("d_a41", h8, w8, 512), d_p3 = act("d_p3", h8, w8, 256);
    Act d_a34 = act("d_a34", h4, w4, 256), d_a32 = act("d_a32", h4, w4, 256), d_a31 = act("d_a31", h4, w4, 256);
    Act d_p2 = act("d_p2", h4, w4, 128), d_a22 = act("d_a22", h2, w2, 128), d_a21 = act("d_a21", h2, w2, 128);
    Act d_p1 = act("d_p1", h2, w2, 64), d_a12 = act("d_a12", H, W, 64), d_a11 = act("d_a11", H, W, 64);

    bool heads1_bias_done = false;
    if (stage <= 0) {
    if (variant >= 1) {
      Act rp = act("rp", h8, w8, 64), r1 = act("r1", h8 - 2, w8 - 2, 64);
      Act rup = act("rup", h4, w4, 64);
      Act d_rf64 = act("d_rf", h4, w4, 64), d_rf16 = act("d_rf", h4, w4, 16, 64, 0);
      Act d_rup = act("d_rup", h4, w4, 64), d_r2 = act("d_r2", h8 - 6, w8 - 6, 64);
      Act d_r1 = act("d_r1", h8 - 2, w8 - 2, 64), d_rp = act("d_rp", h8, w8, 64);
      DBX_TRY(wgrad(rup, d_rf16, "conv6_3_det", 1, 0, st));
      DBX_TRY(dgrad(d_rf64, "conv6_3_det", 1, 0, d_rup, nullptr, st));
      DBX_K("upsample_bwd", 0.0, upsample_bilinear_bwd(d_rup, nullptr, d_r2, st));
      DBX_TRY(wgrad(r1, d_r2, "conv6_2_det", 5, 0, st));
      DBX_TRY(dgrad(d_r2, "conv6_2_det", 5, 0, d_r1, nullptr, st));
      DBX_TRY(wgrad(rp, d_r1, "conv6_1_det", 3, 0, st));
      DBX_TRY(dgrad(d_r1, "conv6_1_det", 3, 0, d_rp, nullptr, st));
      DBX_K("refine_pool_bwd", 0.0, refine_pool_bwd((const float*)buf("head_out"), HC, d_rp.ptr, d_head64.ptr, N, h4, w4, st));
    }
    // heads
    DBX_TRY(wgrad(hd, d_headC, "heads2", 1, 0, st));
    {
      const Group& g = groups[group_id("heads2")];
      if (side && !profiling) { ++launches; DBX_TRY(blockdiag_mask(G32() + g.w_off, g.rows, g.ld, nh, ch_start, side)); }
      else DBX_K("blockdiag_mask", 0.0, blockdiag_mask(G32() + g.w_off, g.rows, g.ld, nh, ch_start, st));
      // The conv5_2 data gradient is a 236 MB store with <= 8 products per element: the streaming kernel (heads2_dgrad,
      // which also yields the bias gradient of conv5_1) replaces the K = 64 tensor-core launch + the stand-alone column
      // sums of d_hd.  DBX_HEADS2_DGRAD=0 (A/B, parity cross-check in tests/test_gpu_fused_bias.py): tensor-core path.
      bool direct = true;
      { const char* e = ab_env("DBX_HEADS2_DGRAD"); if (e) direct = e[0] == '1'; }
      if (direct) {
        heads1_bias_done = fuse;
        DBX_K("dgrad:heads2", 2.0 * pixels(d_head64) * macs_of("heads2"),
              heads2_dgrad(d_head64.ptr, wd_of("heads2"), d_hd.ptr, (size_t)N * h4 * w4, 512 * nh, nh, ch_start,
                           drop_mode, buf("drop"), (const unsigned long long*)buf("rng"),
                           fuse ? gb_of("heads1") : nullptr, st));
      } else {
        ConvEpilogue e;
        if (drop_mode == 2) { e.aux = buf("drop"); e.aux_cs = 512 * nh; e.aux_mode = 2; e.epi_bufs = 8; }
        else if (drop_mode == 3) { e.aux_mode = 3; e.rng = (const unsigned long long*)buf("rng"); e.rng_channels = 512 * nh; }
        DBX_K("dgrad:heads2", 2.0 * pixels(d_head64) * macs_of("heads2"),
              conv_fprop(d_head64, wd_of("heads2"), 1, 1, 0, d_hd, e, 0, st));
      }
    }
    DBX_TRY(wgrad(fus, d_hd, "heads1", 1, 0, st, heads1_bias_done));
    DBX_TRY(dgrad(d_hd, "heads1", 1, 0, d_fus, nullptr, st));
    // conv4 block
    DBX_K("upsample_bwd", 0.0, upsample_bilinear_bwd(d_fus_up, &a44, d_a44, st, fuse ? gb_of("conv4_4") : nullptr));
    DBX_TRY(wgrad(a43, d_a44, "conv4_4", 3, 1, st, fuse));
    DBX_TRY(dgrad(d_a44, "conv4_4", 3, 1, d_a43, &a43, st, fuse ? "conv4_3" : nullptr));
    DBX_TRY(wgrad(a42, d_a43, "conv4_3", 3, 1, st, fuse));
    DBX_TRY(dgrad(d_a43, "conv4_3", 3, 1, d_a42, &a42, st, fuse ? "conv4_2" : nullptr));
    DBX_TRY(wgrad(a41, d_a42, "conv4_2", 3, 1, st, fuse));
    DBX_TRY(dgrad(d_a42, "conv4_2", 3, 1, d_a41, &a41, st, fuse ? "conv4_1" : nullptr));
    DBX_TRY(wgrad(p3, d_a41, "conv4_1", 3, 1, st, fuse));
    DBX_TRY(dgrad(d_a41, "conv4_1", 3, 1, d_p3, nullptr, st));
    if (stage == 0) return join_side(st);
    }
    if (stage == -1 || stage == 1) {
    // conv3 block: pool3 backward + the concat branch of conv3_4, then ReLU mask
    DBX_K("pool_bwd", 0.0, maxpool2x2_bwd(a34, d_p3, &d_fus_34, d_a34, st, fuse ? gb_of("conv3_4") : nullptr));
    DBX_TRY(wgrad(a32, d_a34, "conv3_4", 3, 1, st, fuse));
    DBX_TRY(dgrad(d_a34, "conv3_4", 3, 1, d_a32, &a32, st, fuse ? "conv3_2" : nullptr));
    DBX_TRY(wgrad(a31, d_a32, "conv3_2", 3, 1, st, fuse));
    DBX_TRY(dgrad(d_a32, "conv3_2", 3, 1, d_a31, &a31, st, fuse ? "conv3_1" : nullptr));
    DBX_TRY(wgrad(p2, d_a31, "conv3_1", 3, 1, st, fuse));
    DBX_TRY(dgrad(d_a31, "conv3_1", 3, 1, d_p2, nullptr, st));
    if (stage == 1) return join_side(st);
    }
    // conv2 block
    // pool2 / pool1 backward read the pooled map + the 2-bit arg-max map of the forward pass instead of a22 / a12
    if (pool_idx) DBX_K("pool_bwd", 0.0, maxpool2x2_bwd_idx(p2, d_p2, buf("pi2"), d_a22, st, fuse ? gb_of("conv2_2") : nullptr));
    else DBX_K("pool_bwd", 0.0, maxpool2x2_bwd(a22, d_p2, nullptr, d_a22, st, fuse ? gb_of("conv2_2") : nullptr));
    DBX_TRY(wgrad(a21, d_a22, "conv2_2", 3, 1, st, fuse));
    DBX_TRY(dgrad(d_a22, "conv2_2", 3, 1, d_a21, &a21, st, fuse_short ? "conv2_1" : nullptr));
    DBX_TRY(wgrad(p1, d_a21, "conv2_1", 3, 1, st, fuse_short));
    DBX_TRY(dgrad(d_a21, "conv2_1", 3, 1, d_p1, nullptr, st));
    // conv1 block
    if (pool_idx) DBX_K("pool_bwd", 0.0, maxpool2x2_bwd_idx(p1, d_p1, buf("pi1"), d_a12, st, fuse ? gb_of("conv1_2") : nullptr));
    else DBX_K("pool_bwd", 0.0, maxpool2x2_bwd(a12, d_p1, nullptr, d_a12, st, fuse ? gb_of("conv1_2") : nullptr));
    DBX_TRY(wgrad(a11, d_a12, "conv1_2", 3, 1, st, fuse));
    DBX_TRY(dgrad(d_a12, "conv1_2", 3, 1, d_a11, &a11, st, (fuse_short && !pairs) ? "conv1_1" : nullptr));
    if (pairs) {
      // weight + bias gradient of conv1_1 in the pairs layout: one wgrad into a scratch matrix, then the fold
      Act col0p = act("col0", H, W / 2, 64), d_a11p = act("d_a11", H, W / 2, 128);
      float* scratch = (float*)buf("g11p");
      cudaStream_t ws = (side && !profiling) ? side : st;
      if (ws == side) {
        DBX_TRY((int)cudaEventRecord(ev_fork, st));
        DBX_TRY((int)cudaStreamWaitEvent(side, ev_fork, 0));
        side_used = true;
      }
      DBX_TRY((int)cudaMemsetAsync(scratch, 0, 128 * 64 * 4, ws));
      if (ws == side) {
        launches += 2;
        DBX_TRY(conv_wgrad(col0p, d_a11p, 1, 1, 0, scratch, 0, ws));
        DBX_TRY(conv1_1_fold_pairs(scratch, gw_of("conv1_1"), gb_of("conv1_1"), ws));
      } else {
        DBX_K("wgrad:conv1_1", 2.0 * pixels(d_a11) * macs_of("conv1_1"), conv_wgrad(col0p, d_a11p, 1, 1, 0, scratch, 0, ws));
        DBX_K("fold:conv1_1", 0.0, conv1_1_fold_pairs(scratch, gw_of("conv1_1"), gb_of("conv1_1"), ws));
      }
    } else {
      DBX_TRY(wgrad(col0, d_a11, "conv1_1", 1, 0, st, fuse_short));
    }
    return join_side(st);
  }

  int zero_grad(cudaStream_t st) {
    if (!train) return DBX_ERR_STATE;
    return (int)cudaMemsetAsync(G32(), 0, flat_n * 4, st);
  }

  // part: -1 = every parameter; 0 = gradient buckets 0 and 1 ([0, tail_off)); 1 = the last bucket ([tail_off, flat_n)).
  // A data-parallel caller updates part 0 while the all-reduce of the last bucket is still in flight, then part 1.
  int sgd(float lr, float momentum, float wd, cudaStream_t st, int part = -1) {
    if (!train) return DBX_ERR_STATE;
    if (part < -1 || part > 1) return DBX_ERR_ARG;
    const size_t lo = part == 1 ? tail_off : 0, hi = part == 0 ? tail_off : flat_n;
    DBX_K("sgd", 0.0, sgd_step(W32() + lo, G32() + lo, V32() + lo, WK() + lo, hi - lo, lr, momentum, wd,
                               sgd_steps == 0 ? 1 : 0, 1, st));
    if (part != 0) { ++sgd_steps; dgrad_stale = true; }
    return DBX_OK;
  }
};

}  // namespace dbx

// ------------------------------------------------------------------------------------------------ C ABI
using namespace dbx;
#include "../../include/densebox_b200.h"

extern "C" {

int dbx_net_workspace_bytes(int variant, int N, int H, int W, int train, size_t* bytes) {
  if (!bytes || variant < 0 || variant > 2 || N <= 0 || H <= 0 || W <= 0 || H % 8 || W % 8) return DBX_ERR_ARG;
  if (variant >= 1 && (H / 8 < 7 || W / 8 < 7)) return DBX_ERR_ARG;
  Net n;
  n.build(variant, N, H, W, train);
  *bytes = n.ws_bytes;
  return DBX_OK;
}

int dbx_net_create(int variant, int N, int H, int W, int train, void* workspace, size_t bytes, void* stream,
                   void** handle) {
  size_t need = 0;
  int rc = dbx_net_workspace_bytes(variant, N, H, W, train, &need);
  if (rc) return rc;
  if (!workspace || !handle || ((uintptr_t)workspace & 255)) return DBX_ERR_ARG;
  if (bytes < need) return DBX_ERR_WORKSPACE;
  Net* n = new Net();
  n->build(variant, N, H, W, train);
  n->ws = (char*)workspace;
  rc = n->init((cudaStream_t)stream);
  if (!rc && train) {
    rc = (int)cudaStreamCreateWithFlags(&n->side, cudaStreamNonBlocking);
    if (!rc) rc = (int)cudaEventCreateWithFlags(&n->ev_fork, cudaEventDisableTiming);
    if (!rc) rc = (int)cudaEventCreateWithFlags(&n->ev_join, cudaEventDisableTiming);
  }
  if (rc) { delete n; return rc; }
  *handle = n;
  return DBX_OK;
}

int dbx_net_destroy(void* handle) {
  Net* n = (Net*)handle;
  if (n) {
    if (n->ev_fork) cudaEventDestroy(n->ev_fork);
    if (n->ev_join) cudaEventDestroy(n->ev_join);
    if (n->side) cudaStreamDestroy(n->side);
    n->prof_clear();
    delete n;
  }
  return DBX_OK;
}

int dbx_net_buffer(void* handle, const char* name, void** ptr, size_t* bytes) {
  if (!handle || !name || !ptr) return DBX_ERR_ARG;
  void* p = ((Net*)handle)->buf(name, bytes);
  if (!p) return DBX_ERR_ARG;
  *ptr = p;
  return DBX_OK;
}

int dbx_net_head_channels(void* handle) { return handle ? ((Net*)handle)->HC : DBX_ERR_ARG; }
long long dbx_net_param_elems(void* handle) { return handle ? (long long)((Net*)handle)->flat_n : DBX_ERR_ARG; }

int dbx_net_set_param(void* handle, const char* name, int is_bias, const float* src, long s_co, long s_ci, long s_r,
                      long s_s, void* stream) {
  if (!handle) return DBX_ERR_ARG;
  return ((Net*)handle)->set_param(name, is_bias, src, s_co, s_ci, s_r, s_s, (cudaStream_t)stream);
}
int dbx_net_get_param(void* handle, const char* name, int is_bias, float* dst, long s_co, long s_ci, long s_r,
                      long s_s, void* stream) {
  if (!handle) return DBX_ERR_ARG;
  return ((Net*)handle)->get_tensor(name, is_bias, 0, dst, s_co, s_ci, s_r, s_s, (cudaStream_t)stream);
}
int dbx_net_get_grad(void* handle, const char* name, int is_bias, float* dst, long s_co, long s_ci, long s_r,
                     long s_s, void* stream) {
  if (!handle) return DBX_ERR_ARG;
  return ((Net*)handle)->get_tensor(name, is_bias, 1, dst, s_co, s_ci, s_r, s_s, (cudaStream_t)stream);
}
int dbx_net_xfer_params(void* handle, int mode, int n, const char* const* names, void* const* w_ptrs,
                        const long* w_strides, void* const* b_ptrs, const long* b_strides, void* stream) {
  if (!handle) return DBX_ERR_ARG;
  return ((Net*)handle)->xfer_params(mode, n, names, w_ptrs, w_strides, b_ptrs, b_strides, (cudaStream_t)stream);
}
int dbx_net_get_outputs(void* handle, float* score, float* loc, float* lm, float* lmloc, float* rf, void* stream) {
  if (!handle) return DBX_ERR_ARG;
  return ((Net*)handle)->get_outputs(score, loc, lm, lmloc, rf, (cudaStream_t)stream);
}
int dbx_net_set_output_grads(void* handle, const float* g_score, const float* g_loc, const float* g_lm,
                             const float* g_lmloc, const float* g_rf, void* stream) {
  if (!handle) return DBX_ERR_ARG;
  return ((Net*)handle)->set_output_grads(g_score, g_loc, g_lm, g_lmloc, g_rf, (cudaStream_t)stream);
}
int dbx_net_refresh_dgrad(void* handle, void* stream) {
  if (!handle) return DBX_ERR_ARG;
  return ((Net*)handle)->refresh_dgrad((cudaStream_t)stream);
}
int dbx_net_forward(void* handle, const float* x, int dropout_mode, unsigned long long seed,
                    unsigned long long offset, void* stream) {
  if (!handle) return DBX_ERR_ARG;
  return ((Net*)handle)->forward(x, dropout_mode, seed, offset, (cudaStream_t)stream);
}
int dbx_net_forward_u8(void* handle, const unsigned char* x_u8, const float* table, int dropout_mode,
                       unsigned long long seed, unsigned long long offset, void* stream) {
  if (!handle || !x_u8 || !table) return DBX_ERR_ARG;
  return ((Net*)handle)->forward(nullptr, dropout_mode, seed, offset, (cudaStream_t)stream, x_u8, table);
}
int dbx_net_loss(void* handle, const float* bbox, const float* vertices, const float* labels,
                 const long long* rand_idx, int rand_stride, const long long* lm_rand_idx, float lambda_loc,
                 float lambda_det, float lambda_lm, int global_pos, int global_batch, const int* global_pos_ptr,
                 int clamp_lm, float* d_head_f32, float* d_rf_f32, unsigned char* mask_out,
                 unsigned char* lm_mask_out, void* stream) {
  if (!handle) return DBX_ERR_ARG;
  return ((Net*)handle)->loss(bbox, vertices, labels, rand_idx, rand_stride, lm_rand_idx, lambda_loc, lambda_det,
                              lambda_lm, global_pos, global_batch, global_pos_ptr, clamp_lm, d_head_f32, d_rf_f32,
                              mask_out, lm_mask_out, (cudaStream_t)stream);
}
int dbx_net_backward(void* handle, void* stream) {
  if (!handle) return DBX_ERR_ARG;
  return ((Net*)handle)->backward((cudaStream_t)stream);
}
int dbx_net_backward_stage(void* handle, int stage, void* stream) {
  if (!handle || stage < 0 || stage > 2) return DBX_ERR_ARG;
  return ((Net*)handle)->backward((cudaStream_t)stream, stage);
}
int dbx_net_join(void* handle, void* stream) {
  if (!handle) return DBX_ERR_ARG;
  Net* n = (Net*)handle;
  if (n->dgrad_pending) {
    n->dgrad_pending = false;
    return (int)cudaStreamWaitEvent((cudaStream_t)stream, n->ev_join, 0);
  }
  return DBX_OK;
}
int dbx_net_set_count_slots(void* handle, const void* slots, int world) {
  if (!handle || (slots && (world < 1 || world > 16))) return DBX_ERR_ARG;
  Net* n = (Net*)handle;
  n->count_slots = (const unsigned long long*)slots; n->count_world = slots ? world : 0;
  return DBX_OK;
}
int dbx_net_grad_bucket(void* handle, int bucket, long long* first, long long* count) {
  if (!handle || !first || !count) return DBX_ERR_ARG;
  Net* n = (Net*)handle;
  if (!n->train) return DBX_ERR_STATE;
  const long long mid = (long long)n->mid_off, tail = (long long)n->tail_off;
  switch (bucket) {
    case 0: *first = 0; *count = mid; return DBX_OK;                              // conv4_1 .. heads (+ refine) filters
    case 1: *first = mid; *count = tail - mid; return DBX_OK;                     // conv3 block filters
    case 2: *first = tail; *count = (long long)n->flat_n - tail; return DBX_OK;   // conv1/conv2 filters + every bias
    default: return DBX_ERR_ARG;
  }
}
int dbx_net_zero_grad(void* handle, void* stream) {
  if (!handle) return DBX_ERR_ARG;
  return ((Net*)handle)->zero_grad((cudaStream_t)stream);
}
int dbx_net_profile(void* handle, int enable) {
  if (!handle) return DBX_ERR_ARG;
  Net* n = (Net*)handle;
  n->prof_clear();
  n->profiling = enable != 0;
  return DBX_OK;
}
long long dbx_net_launch_count(void* handle) { return handle ? ((Net*)handle)->launches : -1; }
int dbx_net_profile_count(void* handle) { return handle ? (int)((Net*)handle)->recs.size() : DBX_ERR_ARG; }
int dbx_net_profile_get(void* handle, int i, char* tag, int tag_bytes, double* flops, float* ms) {
  if (!handle || !tag || !flops || !ms) return DBX_ERR_ARG;
  Net* n = (Net*)handle;
  if (i < 0 || i >= (int)n->recs.size()) return DBX_ERR_ARG;
  const Net::Rec& r = n->recs[i];
  snprintf(tag, tag_bytes, "%s", r.tag.c_str());
  *flops = r.flops;
  return (int)cudaEventElapsedTime(ms, r.e0, r.e1);
}
int dbx_net_sgd_step(void* handle, float lr, float momentum, float weight_decay, void* stream) {
  if (!handle) return DBX_ERR_ARG;
  return ((Net*)handle)->sgd(lr, momentum, weight_decay, (cudaStream_t)stream);
}
int dbx_net_sgd_step_part(void* handle, int part, float lr, float momentum, float weight_decay, void* stream) {
  if (!handle || part < 0 || part > 1) return DBX_ERR_ARG;
  return ((Net*)handle)->sgd(lr, momentum, weight_decay, (cudaStream_t)stream, part);
}

}  // extern "C"
