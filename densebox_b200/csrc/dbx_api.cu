// densebox_b200 — C ABI (see include/densebox_b200.h). Plain pointers and sizes only; every call is asynchronous
// on the given stream, allocates nothing, and returns 0 or an error code. No exceptions cross this boundary.
#include "../../include/densebox_b200.h"
#include "dbx_common.h"

using namespace dbx;

static Act mk_act(const void* p, int N, int H, int W, int C, int cs, int coff) {
  Act a;
  a.ptr = const_cast<void*>(p); a.N = N; a.H = H; a.W = W; a.C = C; a.cs = cs; a.coff = coff;
  return a;
}

extern "C" {

int dbx_version(void) { return 100; }

const char* dbx_error_string(int code) {
  switch (code) {
    case DBX_OK: return "ok";
    case DBX_ERR_ARG: return "dbx: invalid argument (shape/alignment/null)";
    case DBX_ERR_DRIVER: return "dbx: cuTensorMapEncodeTiled not resolvable (no CUDA driver?)";
    case DBX_ERR_TMAP: return "dbx: cuTensorMapEncodeTiled failed";
    case DBX_ERR_WORKSPACE: return "dbx: workspace too small";
    case DBX_ERR_STATE: return "dbx: invalid call order";
    default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "dbx: unknown error";
  }
}

int dbx_conv_fprop(const void* x, int N, int H, int W, int cin, int x_cs, int x_coff, const void* wk, int R, int S,
                   int pad, int cout, const float* bias, int relu, const void* aux, int aux_cs, int aux_coff,
                   int aux_mode, void* out, int out_cs, int out_coff, int out_fp32, int block_n, void* stream) {
  Act ax = mk_act(x, N, H, W, cin, x_cs, x_coff);
  Act ao = mk_act(out, N, H + 2 * pad - R + 1, W + 2 * pad - S + 1, cout, out_cs, out_coff);
  ConvEpilogue e;
  e.bias = bias; e.relu = relu; e.aux = aux; e.aux_cs = aux_cs; e.aux_coff = aux_coff; e.aux_mode = aux_mode;
  e.out_fp32 = out_fp32;
  return conv_fprop(ax, wk, R, S, pad, ao, e, block_n, (cudaStream_t)stream);
}

int dbx_conv_wgrad(const void* x, int N, int H, int W, int cin, int x_cs, int x_coff, const void* dy, int cout,
                   int dy_cs, int dy_coff, int R, int S, int pad, float* dw, int block_n, void* stream) {
  Act ax = mk_act(x, N, H, W, cin, x_cs, x_coff);
  Act ad = mk_act(dy, N, H + 2 * pad - R + 1, W + 2 * pad - S + 1, cout, dy_cs, dy_coff);
  return conv_wgrad(ax, ad, R, S, pad, dw, block_n, (cudaStream_t)stream);
}

int dbx_loss_fwd_bwd(const float* head, int HC, const float* rf, int RC, const float* bbox, const float* vertices,
                     const float* labels, const long long* rand_idx, int rand_stride, const long long* lm_rand_idx,
                     int variant, float lambda_loc, float lambda_det, float lambda_lm, int global_pos,
                     int global_batch, const int* global_pos_ptr, int clamp_lm, int B, void* scratch, float* loss,
                     int* info, void* d_head_bf16, void* d_rf_bf16, float* d_head_f32, float* d_rf_f32,
                     unsigned char* mask_out, unsigned char* lm_mask_out, void* stream) {
  if (!scratch) return DBX_ERR_ARG;
  LossParams p{};
  p.head = head; p.HC = HC; p.rf = rf; p.RC = RC; p.bbox = bbox; p.vertices = vertices; p.labels = labels;
  p.rand_idx = rand_idx; p.rand_stride = rand_stride; p.lm_rand_idx = lm_rand_idx; p.variant = variant;
  p.lambda_loc = lambda_loc; p.lambda_det = lambda_det; p.lambda_lm = lambda_lm;
  p.global_pos = global_pos; p.global_batch = global_batch; p.global_pos_ptr = global_pos_ptr;
  p.clamp_lm = clamp_lm; p.B = B;
  p.counter = (unsigned int*)scratch; p.loss_partial = (float*)scratch + 4;
  p.loss = loss; p.info = info;
  p.d_head = (__nv_bfloat16*)d_head_bf16; p.d_rf = (__nv_bfloat16*)d_rf_bf16;
  p.d_head_f32 = d_head_f32; p.d_rf_f32 = d_rf_f32; p.mask_out = mask_out; p.lm_mask_out = lm_mask_out;
  return loss_fwd_bwd(p, (cudaStream_t)stream);
}

int dbx_loss_maps(const float* const* maps, const long* strides, float* const* grads, const float* bbox,
                  const float* vertices, const float* labels, const long long* rand_idx, int rand_stride,
                  const long long* lm_rand_idx, int variant, float lambda_loc, float lambda_det, float lambda_lm,
                  int global_pos, int global_batch, const int* global_pos_ptr, int clamp_lm, int B, void* scratch,
                  float* loss, int* info, unsigned char* mask_out, unsigned char* lm_mask_out, void* stream) {
  if (!scratch || !maps || !strides) return DBX_ERR_ARG;
  LossParams p{};
  for (int g = 0; g < 5; ++g) {
    p.src[g] = MapRef{maps[g], strides[3 * g], strides[3 * g + 1], strides[3 * g + 2]};
    p.dst[g] = MapOut{grads ? grads[g] : nullptr, strides[3 * g], strides[3 * g + 1], strides[3 * g + 2]};
  }
  p.bbox = bbox; p.vertices = vertices; p.labels = labels;
  p.rand_idx = rand_idx; p.rand_stride = rand_stride; p.lm_rand_idx = lm_rand_idx; p.variant = variant;
  p.lambda_loc = lambda_loc; p.lambda_det = lambda_det; p.lambda_lm = lambda_lm;
  p.global_pos = global_pos; p.global_batch = global_batch; p.global_pos_ptr = global_pos_ptr;
  p.clamp_lm = clamp_lm; p.B = B;
  p.counter = (unsigned int*)scratch; p.loss_partial = (float*)scratch + 4;
  p.loss = loss; p.info = info; p.mask_out = mask_out; p.lm_mask_out = lm_mask_out;
  return loss_fwd_bwd(p, (cudaStream_t)stream);
}

int dbx_count_exchange(const float* bbox, const float* labels, int B, void* const* peer_slots, int world, int rank,
                       void* local_slots, void* stream) {
  if (!peer_slots || world < 1 || world > 16) return DBX_ERR_ARG;
  PeerSlots ps{};
  ps.world = world; ps.rank = rank;
  for (int r = 0; r < world; ++r) ps.p[r] = (unsigned long long*)peer_slots[r];
  return count_exchange(bbox, labels, B, ps, (unsigned long long*)local_slots, (cudaStream_t)stream);
}

int dbx_set_tensor_sm_limit(int n) { return set_tensor_sm_limit(n); }

int dbx_count_positives(const float* bbox, const float* labels, int B, int* out, void* stream) {
  return count_positives(bbox, labels, B, out, (cudaStream_t)stream);
}

int dbx_dropout_mask(void* mask, unsigned long long n, unsigned long long seed, unsigned long long offset,
                     void* stream) {
  return dropout_mask(mask, (size_t)n, seed, offset, (cudaStream_t)stream);
}

int dbx_im2col3x3_c3(const float* x, void* out, int N, int H, int W, void* stream) {
  return im2col3x3_c3(x, out, N, H, W, 1, (cudaStream_t)stream);  // 64-channel layout, zero channels written
}
int dbx_maxpool2x2_fwd(const void* y, int N, int H, int W, int C, int y_cs, int y_coff, void* out, int o_cs, int o_coff,
                       void* stream) {
  return maxpool2x2_fwd(mk_act(y, N, H, W, C, y_cs, y_coff), mk_act(out, N, H / 2, W / 2, C, o_cs, o_coff),
                        (cudaStream_t)stream);
}
int dbx_maxpool2x2_bwd(const void* y, int N, int H, int W, int C, int y_cs, int y_coff, const void* dp, int dp_cs,
                       int dp_coff, const void* add, int add_cs, int add_coff, void* dy, int dy_cs, int dy_coff,
                       void* stream) {
  Act a = mk_act(add, N, H, W, C, add_cs, add_coff);
  return maxpool2x2_bwd(mk_act(y, N, H, W, C, y_cs, y_coff), mk_act(dp, N, H / 2, W / 2, C, dp_cs, dp_coff),
                        add ? &a : nullptr, mk_act(dy, N, H, W, C, dy_cs, dy_coff), (cudaStream_t)stream);
}
int dbx_maxpool2x2_fwd_idx(const void* y, int N, int H, int W, int C, int y_cs, int y_coff, void* out, int o_cs,
                           int o_coff, void* idx, void* stream) {
  return maxpool2x2_fwd(mk_act(y, N, H, W, C, y_cs, y_coff), mk_act(out, N, H / 2, W / 2, C, o_cs, o_coff),
                        (cudaStream_t)stream, idx);
}
int dbx_maxpool2x2_bwd_idx(const void* p, int N, int H, int W, int C, int p_cs, int p_coff, const void* dp, int dp_cs,
                           int dp_coff, const void* idx, void* dy, int dy_cs, int dy_coff, float* db, void* stream) {
  return maxpool2x2_bwd_idx(mk_act(p, N, H / 2, W / 2, C, p_cs, p_coff), mk_act(dp, N, H / 2, W / 2, C, dp_cs, dp_coff),
                            idx, mk_act(dy, N, H, W, C, dy_cs, dy_coff), (cudaStream_t)stream, db);
}
int dbx_upsample_bilinear_fwd(const void* in, int N, int h, int w, int C, int in_cs, int in_coff, void* out, int H,
                              int W, int o_cs, int o_coff, void* stream) {
  return upsample_bilinear_fwd(mk_act(in, N, h, w, C, in_cs, in_coff), mk_act(out, N, H, W, C, o_cs, o_coff),
                               (cudaStream_t)stream);
}
int dbx_upsample_bilinear_bwd(const void* dout, int N, int H, int W, int C, int d_cs, int d_coff, const void* relu_y,
                              int y_cs, int y_coff, void* din, int h, int w, int i_cs, int i_coff, void* stream) {
  Act y = mk_act(relu_y, N, h, w, C, y_cs, y_coff);
  return upsample_bilinear_bwd(mk_act(dout, N, H, W, C, d_cs, d_coff), relu_y ? &y : nullptr,
                               mk_act(din, N, h, w, C, i_cs, i_coff), (cudaStream_t)stream);
}
int dbx_colsum(const void* dy, int N, int H, int W, int C, int cs, int coff, float* db, void* stream) {
  return colsum(mk_act(dy, N, H, W, C, cs, coff), db, (cudaStream_t)stream);
}

int dbx_decode_nms(const float* score, long s_img, long s_pix, const float* loc, long l_img, long l_pix, long l_ch,
                   const float* lmloc, long m_img, long m_pix, long m_ch, int N, int H4, int W4, int K, double thresh,
                   float* dets, int* keep, void* stream) {
  return decode_nms(score, s_img, s_pix, loc, l_img, l_pix, l_ch, lmloc, m_img, m_pix, m_ch, N, H4, W4, K, thresh,
                    dets, keep, (cudaStream_t)stream);
}

int dbx_decode_nms_heat(const float* score, long s_img, long s_pix, const float* loc, long l_img, long l_pix, long l_ch,
                        const float* lmheat, long m_img, long m_pix, long m_ch, int N, int H4, int W4, int K,
                        double thresh, float* dets, int* keep, void* stream) {
  return decode_nms(score, s_img, s_pix, loc, l_img, l_pix, l_ch, lmheat, m_img, m_pix, m_ch, N, H4, W4, K, thresh,
                    dets, keep, (cudaStream_t)stream, 1);
}

int dbx_warp_perspective_u8(const unsigned char* src, int H, int W, int C, const double* minv, unsigned char* dst,
                            int dH, int dW, void* stream) {
  return warp_perspective_u8(src, H, W, C, minv, dst, dH, dW, (cudaStream_t)stream);
}

}  // extern "C"
