// densebox_b200 — host-side common definitions shared by all translation units.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

namespace dbx {

// Error codes returned through the C ABI (0 = ok, >0 = cudaError_t, <0 = these).
enum {
  DBX_OK = 0,
  DBX_ERR_ARG = -1,       // bad shape / alignment / null pointer
  DBX_ERR_DRIVER = -2,    // could not resolve cuTensorMapEncodeTiled
  DBX_ERR_TMAP = -3,      // cuTensorMapEncodeTiled failed
  DBX_ERR_WORKSPACE = -4, // workspace too small
  DBX_ERR_STATE = -5,     // call order (e.g. backward before forward)
};

// A view of an NHWC bf16 activation living inside a (possibly wider) channel-interleaved buffer.
struct Act {
  void* ptr;   // base of the buffer (channel 0 of pixel 0)
  int N, H, W; // logical extent
  int C;       // channels of this view (multiple of 16; TMA loads need multiples of 64 or OOB-fill)
  int cs;      // channel stride of the buffer (elements per pixel), multiple of 8
  int coff;    // first channel of this view inside the buffer, multiple of 8
};

// Pixel tile: a (tw x th x tn) box of output pixels handled as <=128 GEMM rows.
struct Tile {
  int tw, th, tn;
  int tiles_w, tiles_h, tiles_n;
  int rows() const { return tw * th * tn; }
  int count() const { return tiles_w * tiles_h * tiles_n; }
};

// Choose the box shape maximising row utilisation. need_mult16: rows must be a multiple of 16 (wgrad K).
Tile choose_tile(int W, int H, int N, bool need_mult16);

int encode_act_map(CUtensorMap* m, const Act& a, const Tile& t);                // 4D (C, W, H, N), box (64, tw, th, tn)
int encode_mat_map(CUtensorMap* m, const void* ptr, int rows, int cols, int box_rows); // 2D (cols, rows), box (64, box_rows)

int num_sms();          // of the CURRENT device (cached per device)
int tensor_sms();       // SMs the persistent tcgen05 kernels may use: num_sms() capped by set_tensor_sm_limit(), even
int set_tensor_sm_limit(int n);  // 0 = no limit; returns the previous value
bool pdl_enabled();
// A/B measurement switches (DBX_* environment variables read by the launchers) are honoured only when the process
// sets DBX_ENABLE_AB=1 before the library is first used; otherwise this returns nullptr for every name, so the kernel
// selection of a production process cannot be changed through its environment.  INTEGRATION.md lists the switches.
const char* ab_env(const char* name);
// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is a per-device setting: applied once per (kernel, device).
static constexpr int kMaxDevices = 64;
struct SmemAttrOnce { int done[kMaxDevices]; int rc[kMaxDevices]; };
int set_max_smem_once(const void* fn, int bytes, SmemAttrOnce* state);

// ---- kernels (launchers). All are asynchronous on `stream`, allocate nothing, and return an error code.

struct ConvEpilogue {
  const float* bias = nullptr; // [cout] fp32 or null
  int relu = 0;
  const void* aux = nullptr;   // bf16, pixel-aligned with the output
  int aux_cs = 0, aux_coff = 0;
  int aux_mode = 0;            // 0 none, 1 out = aux>0 ? v : 0 (ReLU backward), 2 out = v*aux (dropout scale),
                               // 3 out = v * {0,2} drawn in place from Philox (rng = device {seed, offset})
  const unsigned long long* rng = nullptr;
  int rng_channels = 0;        // channel count of the dropped activation (element index = pixel*rng_channels + ch)
  int out_fp32 = 0;            // output element type: 0 bf16, 1 fp32
  int force_cta2 = -1;         // -1 = auto, 0 = never pair CTAs (tcgen05.mma.cta_group::2)
  int epi_bufs = 0;            // 0 = auto; 2/4/8 epilogue staging boxes (short-K GEMMs with a mask want a deep ring)
  // 2x2 max-pool fused behind bias + ReLU (conv1_2 -> pool1, conv2_2 -> pool2, DenseBox.py:186-191): the epilogue
  // writes the POOLED activation (view pool_out: buffer, channel stride, channel offset) and, when pool_idx is given,
  // the 2-bit arg-max map of maxpool2x2_fwd; the full-resolution output is not stored at all (nothing downstream
  // reads it: the backward pass works from the pooled map + the arg-max map).  Falls back to conv + maxpool2x2_fwd
  // (which does store `out`) when the launch cannot use 8 x 16 column-box tiles.
  void* pool_out = nullptr;
  int pool_cs = 0, pool_coff = 0;
  void* pool_idx = nullptr;
  int pool_keep_full = 0;      // 1: store the full-resolution output too (conv3_4: the fusion buffer needs it)
  float* colsum = nullptr;     // fp32 [cout] or null: colsum[c] += sum over pixels of out[.., c] (as stored, bf16-rounded).
                               // A data-gradient launch uses it to produce the bias gradient of the layer below
                               // in its own epilogue instead of re-reading dY from HBM; requires bias == nullptr.
};

// out[n,oh,ow,:] = epi( sum_{r,s,ci} x[n,oh+r-pad,ow+s-pad,ci] * wk[co][(r*S+s)*Cin+ci] )
// x.C must be a multiple of 64 (pad with zero channels); wk is bf16 [cout_rows][R*S*x.C] K-major.
int conv_fprop(const Act& x, const void* wk, int R, int S, int pad, const Act& out, const ConvEpilogue& epi,
               int block_n, cudaStream_t stream);

// 3x3 pad-1 convolution for Cin <= 128 with column-box reuse (dbx_conv_halo.cu); DBX_ERR_ARG = shape not handled.
int conv3x3_halo(const Act& x, const void* wk, const Act& out, const ConvEpilogue& epi, cudaStream_t stream);

// dw[co][(r*S+s)*Cin+ci] += sum_{n,oh,ow} dy[n,oh,ow,co] * x[n,oh+r-pad,ow+s-pad,ci]   (fp32, red.add)
// dw has row stride R*S*x.C; rows = dy.C.
int conv_wgrad(const Act& x, const Act& dy, int R, int S, int pad, float* dw, int block_n, cudaStream_t stream);


// ---- HBM-bound kernels (dbx_elementwise.cu)
int im2col3x3_c3(const float* x, void* out, int N, int H, int W, int mode, cudaStream_t st);  // mode 2: pairs layout
int im2col3x3_c3_u8(const unsigned char* x, const float* lut, void* out, int N, int H, int W, int mode, cudaStream_t st);
int conv1_1_fold_pairs(const float* scratch, float* gw, float* gb, cudaStream_t st);
// idx (optional): u16 per (pooled pixel, 8-channel vector), 2 bits per element = position of the first maximum
int maxpool2x2_fwd(const Act& y, const Act& o, cudaStream_t st, void* idx = nullptr);
// backward from that map and the pooled activation p (no full-resolution read): dy = relu'(y) * unpool(dp)
int maxpool2x2_bwd_idx(const Act& p, const Act& dp, const void* idx, const Act& dy, cudaStream_t st, float* db = nullptr);
// db (optional, both): db[c] += column sums of the gradient written (bias gradient of the conv that produced y)
int maxpool2x2_bwd(const Act& y, const Act& dp, const Act* add, const Act& dy, cudaStream_t st, float* db = nullptr);
int upsample_bilinear_fwd(const Act& in, const Act& out, cudaStream_t st);
int upsample_bilinear_bwd(const Act& dout, const Act* relu_y, const Act& din, cudaStream_t st, float* db = nullptr);
int colsum(const Act& dy, float* db, cudaStream_t st);
// data gradient of the block-diagonal conv5_2 heads + dropout backward (+ optional bias gradient of conv5_1 in db)
int heads2_dgrad(const void* d_head, const void* wd, void* out, size_t pixels, int K, int nh, const int* ch_start,
                 int drop_mode, const void* mask, const unsigned long long* rng, float* db, cudaStream_t st);
int pack_weights(const float* src, int co, int ci, int R, int S, long s_co, long s_ci, long s_r, long s_s, void* dstK,
                 long ldK, long rowK, long kK, int cin_pad, void* dstD, long ldD, long rowD, long kD, int cout_pad,
                 float* dstF, cudaStream_t st);
int unpack_weights(const float* srcK, long ldK, long rowK, long kK, int cin_pad, float* dst, int co, int ci, int R,
                   int S, long s_co, long s_ci, long s_r, long s_s, cudaStream_t st);
int transpose_dgrad(const void* wk, void* wd, int rows, int T, int cin_pad, int kpad, cudaStream_t st);
// the same for up to 16 weight matrices in one launch (element offsets into the flat bf16 buffers)
struct TransposeBatch {
  static constexpr int kMax = 16;
  struct G { long long wk_off, wd_off; int rows, T, cin_pad, kpad, block0; };
  G g[kMax];
  int n;
};
int transpose_dgrad_multi(const void* wk_base, void* wd_base, const TransposeBatch& tb, int total_blocks,
                          cudaStream_t st);
// Every (weight, bias) pair of a network in ONE launch: mode 0 = torch tensors -> fp32 master + bf16 K-major copies,
// 1 = flat fp32 buffer (gradients or masters) -> torch tensors.  Entries address the flat buffers by element offset.
struct ParamXfer {
  static constexpr int kMax = 24;
  struct E {
    float* w; long s_co, s_ci, s_r, s_s;   // torch weight [co,ci,R,S] with element strides
    float* b; long s_b;                    // torch bias [co]
    int co, ci, R, S;
    long long w_off, b_off;                // group matrix / first bias element inside the flat buffers
    long ld, rowK, kK;                     // K-major placement: row rowK + o, column kK + (r*S+s)*cin_pad + i
    int cin_pad, dup;                      // dup: conv1_1 pairs layout, second diagonal block (rows 64.., columns 32..)
    long long elem0;                       // running element count before this entry (weights + biases)
  } e[kMax];
  int n;
  long long total;
};
int params_xfer(const ParamXfer& t, int mode, float* flat32, void* flat16, cudaStream_t st);
// head_out fp32 NHWC [pixels][HC] (+ rf_out [pixels][RC]) <-> the NCHW fp32 tensors the modules return / receive
int heads_to_nchw(const float* head, int HC, const float* rf, int RC, int N, int HW, float* score, float* loc,
                  float* lm, float* lmloc, float* rfo, cudaStream_t st);
int nchw_to_head_grads(const float* g_score, const float* g_loc, const float* g_lm, const float* g_lmloc,
                       const float* g_rf, int N, int HW, void* d_head, void* d_rf, cudaStream_t st);
int sgd_step(float* w, float* g, float* v, void* wb, size_t n, float lr, float momentum, float wd, int first,
             int zero_grad, cudaStream_t st);
int cast_bf16(const float* src, void* dst, size_t n, cudaStream_t st);
int blockdiag_mask(float* g, int rows, int ld, int nh, const int* start, cudaStream_t st);
// conv5_2 o conv5_1 as one matrix (inference only; see heads_fold_kernel): wf bf16 [HC][768], bf fp32 [HC]
int heads_fold(const float* w1, int ld1, const float* b1, const float* w2, int ld2, const float* b2, const int* start,
               int nh, int HC, void* wf, float* bf, cudaStream_t st);
int dropout_mask(void* mask, size_t n, unsigned long long seed, unsigned long long offset, cudaStream_t st);
int refine_pool_pack(const float* head, int HC, void* pooled, int N, int H, int W, cudaStream_t st);
int refine_pool_bwd(const float* head, int HC, const void* dpooled, void* dhead, int N, int H, int W, cudaStream_t st);

// ---- fused loss (dbx_loss.cu)
// One group of fp32 maps with explicit element strides (image, pixel, channel): the engine's NHWC head buffer and the
// NCHW tensors of the drop-in densebox_loss() are read (and their gradients written) through the same kernel.
struct MapRef { const float* p; long img, pix, ch; };
struct MapOut { float* p; long img, pix, ch; };
enum { LOSS_SCORE = 0, LOSS_LOC = 1, LOSS_LM = 2, LOSS_LMLOC = 3, LOSS_RF = 4 };
struct LossParams {
  const float* head; int HC;      // [B,60,60,HC] fp32: ch0 score, 1..4 loc, 5..8 landmark heat, 9..16 landmark loc
  const float* rf; int RC;        // [B,60,60,RC] fp32, ch0 = refine score (variants 1,2)
  MapRef src[5];                  // head == null: the five map groups given one by one (score 1 ch, loc 4, landmark
                                  // heat 4, landmark loc 8, refine 1); head != null: filled by the launcher
  MapOut dst[5];                  // optional fp32 gradients per group (p == null: not written)
  const float* bbox;              // [B,4] 60-space
  const float* vertices;          // [B,8] or null
  const float* labels;            // [B] or null (= all positive patches)
  const long long* rand_idx; int rand_stride;  // [B,rand_stride] injected random negatives (first `half` used)
  const long long* lm_rand_idx;   // [B,4]
  int variant;                    // 0 DenseBox, 1 DenseBoxLM, 2 DenseBoxLMLOC
  float lambda_loc, lambda_det, lambda_lm;
  int global_pos, global_batch;   // data-parallel: batch-global positive count / batch size; <0 = use this launch
  const int* global_pos_ptr;      // optional device int: overrides global_pos (filled by count_positives + allreduce)
  const unsigned long long* count_slots; int count_world;  // optional: step-tagged per-rank counts written by the peers
                                  // (count_exchange): [2][world] slots + the local step counter at [2 * world]
  int clamp_lm;                   // init_lm_heatmap_pn clamping (:1899-1907)
  int B;
  float* loss_partial;            // [B]
  float* loss;                    // [1]
  unsigned int* counter;          // zero before first use; self-resetting
  __nv_bfloat16* d_head;          // [B,60,60,64] bf16 or null
  __nv_bfloat16* d_rf;            // [B,60,60,64] bf16 or null
  float* d_head_f32;              // [B,60,60,HC] or null
  float* d_rf_f32;                // [B,60,60,RC] or null
  unsigned char* mask_out;        // [B,3600] or null
  unsigned char* lm_mask_out;     // [B,4,3600] or null
  int* info;                      // [4] = {half, pos, rand_short, 0} or null; rand_short = 1 when half > rand_stride
                                  // (fewer injected random negatives than the quota asks for)
};
int loss_fwd_bwd(const LossParams& p, cudaStream_t st);
int count_positives(const float* bbox, const float* labels, int B, int* out, cudaStream_t st);
// Data parallel without a collective on the critical path: the positives of this rank's shard are written, tagged with
// the step number, straight into the slot buffer of EVERY rank (peer stores over NVLink; peers[r] = rank r's buffer,
// local = this rank's); the loss kernel of each rank sums the slots of the step (LossParams::count_slots).
struct PeerSlots { unsigned long long* p[16]; int world, rank; };
int count_exchange(const float* bbox, const float* labels, int B, const PeerSlots& peers, unsigned long long* local,
                   cudaStream_t st);


// ---- detection post-processing (dbx_postproc.cu); maps are fp32 with explicit image/pixel/channel element strides
int decode_nms(const float* score, long s_img, long s_pix, const float* loc, long l_img, long l_pix, long l_ch,
               const float* lmloc, long m_img, long m_pix, long m_ch, int N, int H4, int W4, int K, double thresh,
               float* dets, int* keep, cudaStream_t st, int lm_heat = 0);  // lm_heat: `lmloc` = 4 heat-maps (parse_DetLM)

// cv2.warpPerspective(INTER_LINEAR, constant border) on uint8 HWC images; minv = 3 x 3 inverse map (HOST pointer)
int warp_perspective_u8(const unsigned char* src, int H, int W, int C, const double* minv, unsigned char* dst, int dH,
                        int dW, cudaStream_t st);

}  // namespace dbx
