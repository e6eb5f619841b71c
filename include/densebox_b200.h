/* densebox_b200 — C ABI of the B200-native DenseBox hot path.
 *
 * The reference (CaptainEven/DenseBox) is pure Python on top of torch.nn; its "FFI" for this path is the set of
 * torch.nn / ATen calls made by DenseBox{,LM,LMLOC}.forward and by the train_* loop bodies.  Every entry point below
 * names the reference call site(s) it replaces (file:line into DenseBox.py @ 7340ed0).
 *
 * Conventions: all pointers are caller-owned DEVICE pointers unless stated otherwise; activations are NHWC bf16
 * inside a channel-interleaved buffer (`cs` = elements per pixel of the buffer, `coff` = first channel of the view);
 * `stream` is a cudaStream_t; calls are asynchronous, allocate nothing and never synchronise.
 * Return value: 0 = ok, >0 = cudaError_t, <0 = DBX_ERR_* (dbx_error_string() explains both).
 */
#ifndef DENSEBOX_B200_H
#define DENSEBOX_B200_H
#ifdef __cplusplus
extern "C" {
#endif

int dbx_version(void);
const char* dbx_error_string(int code);

/* nn.Conv2d forward (+bias, +ReLU) — DenseBox.py:49-140 (backbone blocks), :149-178 (1x1 heads), :399-410 (refine);
 * the same kernel is the data-gradient of loss.backward() (DenseBox.py:2925) when fed the flipped/transposed filter.
 *   x   : [N,H,W,cin] bf16 view (cin % 64 == 0)          wk : [cout][R*S*cin] bf16, K-major (tap-major, then cin)
 *   out : [N,H+2p-R+1,W+2p-S+1,cout] bf16 or fp32 view (cout % 16 == 0)
 *   aux_mode 0: none; 1: out = aux>0 ? v : 0 (ReLU backward mask); 2: out = v*aux (Dropout scale, DenseBox.py:160)
 *   block_n: output-channel tile (multiple of 16, <=256); 0 = auto. */
int dbx_conv_fprop(const void* x, int N, int H, int W, int cin, int x_cs, int x_coff, const void* wk, int R, int S,
                   int pad, int cout, const float* bias, int relu, const void* aux, int aux_cs, int aux_coff,
                   int aux_mode, void* out, int out_cs, int out_coff, int out_fp32, int block_n, void* stream);

/* nn.Conv2d weight gradient (autograd of DenseBox.py:2925):
 *   dw[co][(r*S+s)*cin+ci] += sum_{n,oh,ow} dy[n,oh,ow,co] * x[n,oh+r-p,ow+s-p,ci]   (fp32, accumulating). */
int dbx_conv_wgrad(const void* x, int N, int H, int W, int cin, int x_cs, int x_coff, const void* dy, int cout,
                   int dy_cs, int dy_coff, int R, int S, int pad, float* dw, int block_n, void* stream);

/* HBM-bound neighbours of the convolutions (NHWC bf16 views, 8-channel vectors; C % 8 == 0):
 *  im2col3x3_c3        : X fp32 NCHW [N,3,H,W] -> bf16 [N,H,W,64] (27 taps*channels + zero pad) feeding conv1_1 (:185)
 *  maxpool2x2_fwd/bwd  : nn.MaxPool2d(2,2) (:187,:191,:204) and its backward fused with the ReLU mask of the
 *                        producing conv and an optional second gradient (`add`, the concat branch of conv3_4)
 *  upsample_bilinear_* : nn.Upsample(size, 'bilinear', align_corners=True) (:213-216,:468-470); bwd optionally
 *                        applies the ReLU mask of `relu_y`
 *  colsum              : db[c] += sum over pixels (bias gradient) */
int dbx_im2col3x3_c3(const float* x, void* out, int N, int H, int W, void* stream);
int dbx_maxpool2x2_fwd(const void* y, int N, int H, int W, int C, int y_cs, int y_coff, void* out, int o_cs, int o_coff,
                       void* stream);
int dbx_maxpool2x2_bwd(const void* y, int N, int H, int W, int C, int y_cs, int y_coff, const void* dp, int dp_cs,
                       int dp_coff, const void* add, int add_cs, int add_coff, void* dy, int dy_cs, int dy_coff,
                       void* stream);
/* Same pooling, plus a compact arg-max map idx (u16 per pooled pixel and 8-channel vector, 2 bits per element =
 * position 2*dy+dx of the first maximum); and the backward that uses it together with the POOLED activation p instead
 * of re-reading the full-resolution y: dy[k] = (k == arg && p > 0) ? dp : 0; db (optional) += column sums of dy (bias
 * gradient of the conv that produced y).  H, W are the full-resolution extent in both. */
int dbx_maxpool2x2_fwd_idx(const void* y, int N, int H, int W, int C, int y_cs, int y_coff, void* out, int o_cs,
                           int o_coff, void* idx, void* stream);
int dbx_maxpool2x2_bwd_idx(const void* p, int N, int H, int W, int C, int p_cs, int p_coff, const void* dp, int dp_cs,
                           int dp_coff, const void* idx, void* dy, int dy_cs, int dy_coff, float* db, void* stream);
int dbx_upsample_bilinear_fwd(const void* in, int N, int h, int w, int C, int in_cs, int in_coff, void* out, int H,
                              int W, int o_cs, int o_coff, void* stream);
int dbx_upsample_bilinear_bwd(const void* dout, int N, int H, int W, int C, int d_cs, int d_coff, const void* relu_y,
                              int y_cs, int y_coff, void* din, int h, int w, int i_cs, int i_coff, void* stream);
int dbx_colsum(const void* dy, int N, int H, int W, int C, int cs, int coff, float* db, void* stream);

/* Detection post-processing — parse_out_MN / parse_DetLM / parse_DetLMLOC (DenseBox.py:3114-3348) + NMS (:3398-3443),
 * one CTA per image: top-K of the raw score map (no sigmoid / threshold, as in the reference), decode
 * x = (xi - loc[c, idx]) * 4, greedy NMS (areas with +1, keep while IoU <= thresh).  Maps are fp32 with explicit
 * element strides (image, pixel, channel) so NCHW tensors and the engine's NHWC buffers both work.
 * dets: [N,K,13] fp32 rows (xt,yt,xb,yb,score,x0,y0..x3,y3) in descending score order; keep: [N,K] int32 flags. */
int dbx_decode_nms(const float* score, long s_img, long s_pix, const float* loc, long l_img, long l_pix, long l_ch,
                   const float* lmloc, long m_img, long m_pix, long m_ch, int N, int H4, int W4, int K, double thresh,
                   float* dets, int* keep, void* stream);
/* parse_DetLM (DenseBox.py:3220-3300): as above, but the landmarks of every detection are the arg-max positions (x4)
 * of the four landmark HEAT-maps `lmheat` [.,4,..] (:3283-3292). */
int dbx_decode_nms_heat(const float* score, long s_img, long s_pix, const float* loc, long l_img, long l_pix, long l_ch,
                        const float* lmheat, long m_img, long m_pix, long m_ch, int N, int H4, int W4, int K,
                        double thresh, float* dets, int* keep, void* stream);

/* perspective_transform (DenseBox.py:3446-3481): cv2.warpPerspective(INTER_LINEAR, constant border 0) of a uint8 HWC
 * image, bit-identical to OpenCV's fixed-point remap.  minv: HOST pointer to the 3 x 3 INVERSE map (destination pixel
 * -> source coordinates, row-major doubles; the Python wrapper derives it from the four landmark points as
 * cv2.getPerspectiveTransform + inversion do). */
int dbx_warp_perspective_u8(const unsigned char* src, int H, int W, int C, const double* minv, unsigned char* dst,
                            int dH, int dW, void* stream);

/* The fused loss on caller-provided head maps (same semantics as dbx_net_loss below; used by the drop-in
 * densebox_loss() op).  head: fp32 [B,60,60,HC] in the channel map below, rf: fp32 [B,60,60,RC] (variants 1,2).
 * scratch: >= 16 + 4*B bytes of device memory, zeroed once before the first call.  Outputs may be NULL.
 * info: int[4] = {half, pos, rand_short, 0}. */
int dbx_loss_fwd_bwd(const float* head, int HC, const float* rf, int RC, const float* bbox, const float* vertices,
                     const float* labels, const long long* rand_idx, int rand_stride, const long long* lm_rand_idx,
                     int variant, float lambda_loc, float lambda_det, float lambda_lm, int global_pos,
                     int global_batch, const int* global_pos_ptr, int clamp_lm, int B, void* scratch, float* loss,
                     int* info, void* d_head_bf16, void* d_rf_bf16, float* d_head_f32, float* d_rf_f32,
                     unsigned char* mask_out, unsigned char* lm_mask_out, void* stream);
/* The same kernel on the five map groups given one by one with explicit element strides — what densebox_loss() calls
 * with the NCHW tensors returned by forward(): maps[5] = {score, loc, landmark heat, landmark loc, refine} (NULL where
 * the variant has none), strides[15] = (image, pixel, channel) per group, grads[5] (optional, same strides) receive
 * d(loss)/d(map) in fp32.  info: int[4] = {half, pos, rand_short, 0}; rand_short = 1 when the negative quota `half`
 * exceeds rand_stride, i.e. fewer injected random negatives were available than DenseBox.py:2888-2893 would draw. */
int dbx_loss_maps(const float* const* maps, const long* strides, float* const* grads, const float* bbox,
                  const float* vertices, const float* labels, const long long* rand_idx, int rand_stride,
                  const long long* lm_rand_idx, int variant, float lambda_loc, float lambda_det, float lambda_lm,
                  int global_pos, int global_batch, const int* global_pos_ptr, int clamp_lm, int B, void* scratch,
                  float* loss, int* info, unsigned char* mask_out, unsigned char* lm_mask_out, void* stream);
/* Positive pixels of a label shard (sum of the clipped init_score_map boxes, DenseBox.py:2864) -> *out (device). */
int dbx_count_positives(const float* bbox, const float* labels, int B, int* out, void* stream);
/* The same count exchanged WITHOUT a collective (data parallel): every rank owns a slot buffer of
 * 2 * world + 1 unsigned 64-bit words in peer-accessible memory (zeroed before first use; peer_slots[r] = rank r's
 * buffer as mapped into this process, HOST array of device pointers; local_slots = this rank's own).  The kernel
 * writes (step << 32 | count) into slot [step & 1][rank] of every rank with system-scope stores over NVLink and keeps
 * the step counter in word [2 * world] of the local buffer.  dbx_net_set_count_slots(handle, local_slots, world) makes
 * dbx_net_loss sum the slots of the current step (it spins until every rank's word carries the step's tag — they
 * were written a forward pass earlier) instead of reading global_pos / global_pos_ptr. */
int dbx_count_exchange(const float* bbox, const float* labels, int B, void* const* peer_slots, int world, int rank,
                       void* local_slots, void* stream);
int dbx_net_set_count_slots(void* handle, const void* slots, int world);
/* nn.Dropout(p=0.5) keep-mask x2 as an explicit bf16 tensor — the same Philox4x32-10 bits the conv epilogues draw
 * in place for dropout_mode 1/3 (element e: bit e&127 of philox(counter (e>>7)+offset, key seed)); n % 16 == 0. */
int dbx_dropout_mask(void* mask, unsigned long long n, unsigned long long seed, unsigned long long offset,
                     void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Network engine: one DenseBox replica (variant 0 = DenseBox :31-228, 1 = DenseBoxLM :232-473,
 * 2 = DenseBoxLMLOC :477-738) for a fixed [N,3,H,W] input, laid out inside a caller-owned device workspace.
 * The handle is a small host object; the engine never allocates device memory.
 * Head-output channel map (head_out, fp32 [N,H/4,W/4,16|32]): 0 = score, 1..4 = bbox loc, 5..8 = landmark
 * heat-maps, 9..16 = landmark offsets (variant 2);  rf_out (fp32 [N,H/4,W/4,16]): 0 = refine score. */
#include <stddef.h>
int dbx_net_workspace_bytes(int variant, int N, int H, int W, int train, size_t* bytes);
int dbx_net_create(int variant, int N, int H, int W, int train, void* workspace, size_t bytes, void* stream,
                   void** handle);
int dbx_net_destroy(void* handle);
/* Named region of the workspace ("head_out", "rf_out", "fusion", "drop", "d_head", "d_rf", "w32", "g32", ...). */
int dbx_net_buffer(void* handle, const char* name, void** ptr, size_t* bytes);
int dbx_net_head_channels(void* handle);
long long dbx_net_param_elems(void* handle);

/* state_dict plumbing (DenseBox.py:2809,2945): tensors are addressed by the reference's unique names
 * ("conv1_1".."conv4_4" without conv3_3, "conv5_1_det", "conv5_2_loc", "conv6_1_det", ...); src/dst are fp32 device
 * tensors [cout,cin,R,S] with the given element strides (bias: stride s_co only). */
int dbx_net_set_param(void* handle, const char* name, int is_bias, const float* src, long s_co, long s_ci, long s_r,
                      long s_s, void* stream);
int dbx_net_get_param(void* handle, const char* name, int is_bias, float* dst, long s_co, long s_ci, long s_r,
                      long s_s, void* stream);
int dbx_net_get_grad(void* handle, const char* name, int is_bias, float* dst, long s_co, long s_ci, long s_r,
                     long s_s, void* stream);
/* The same for a whole list of tensors in ONE kernel launch (what the drop-in modules use every step): mode 0 = torch
 * tensors -> engine (fp32 masters + bf16 GEMM copies), 1 = engine masters -> torch tensors, 2 = engine gradients ->
 * torch tensors.  names[n]; w_ptrs[n] with w_strides[4n] (co, ci, r, s element strides); b_ptrs[n], b_strides[n]. */
int dbx_net_xfer_params(void* handle, int mode, int n, const char* const* names, void* const* w_ptrs,
                        const long* w_strides, void* const* b_ptrs, const long* b_strides, void* stream);
/* forward() return values (DenseBox.py:228, :473, :738) as contiguous NCHW fp32 tensors [N,{1,4,4,8,1},H/4,W/4] copied
 * out of head_out / rf_out in one launch (NULL = not wanted), and the way back for loss.backward(): the gradients of
 * those tensors (NULL = zero) packed into the bf16 "d_head" / "d_rf" regions that dbx_net_backward consumes. */
int dbx_net_get_outputs(void* handle, float* score, float* loc, float* lm, float* lmloc, float* rf, void* stream);
int dbx_net_set_output_grads(void* handle, const float* g_score, const float* g_loc, const float* g_lm,
                             const float* g_lmloc, const float* g_rf, void* stream);
/* Rebuild the flipped/transposed bf16 filters used by the data gradients (after set_param, before backward). */
int dbx_net_refresh_dgrad(void* handle, void* stream);

/* net.forward(X) — DenseBox.py:180-228 / :412-473 / :674-738. x: fp32 NCHW [N,3,H,W] device pointer.
 * dropout_mode 0 = eval(); 1 = train(), nn.Dropout drawn inside the conv epilogues from Philox(seed, offset);
 * 2 = train() with the {0,2} bf16 mask the caller has written into the "drop" region (parity tests inject the
 * oracle's mask); 3 = like 1 but {seed, offset} are read from the "rng" region (2 x u64) as the caller left it
 * (CUDA-graph replays).
 * A net created with train = 0 (no backward pass, dropout_mode must be 0) computes the head maps through ONE folded
 * 768 -> C matrix per forward: nothing but Dropout sits between conv5_1_* and conv5_2_* (DenseBox.py:158-178), so in
 * eval mode their product is the same linear map; the "hd" region is not written by such a net. */
int dbx_net_forward(void* handle, const float* x, int dropout_mode, unsigned long long seed,
                    unsigned long long offset, void* stream);

/* The same forward pass fed by the callers' side of the path: x_u8 = the decoded image bytes, uint8
 * NHWC [N,H,W,3]; table = fp32 [3][256], table[c][u] = ToTensor + Normalize of byte u in channel c (DenseBox.py:766-772,
 * :3609-3621; densebox_b200.data.ingest_table builds it with torchvision's float32 arithmetic).  The lookup is fused
 * into the im2col kernel of conv1_1: bit-identical to normalising on the host, a quarter of the host->device bytes. */
int dbx_net_forward_u8(void* handle, const unsigned char* x_u8, const float* table, int dropout_mode,
                       unsigned long long seed, unsigned long long offset, void* stream);

/* The loop body between forward and backward — DenseBox.py:2843-2918 (variant 0), :2575-2723 (1), :2300-2456 and
 * :2023-2180 (2, `labels` != NULL selects the pos/neg-patch `_pn` helpers).  Inputs are device pointers:
 * bbox [N,4] / vertices [N,8] in 60-space floats, labels [N] (or NULL), rand_idx [N,rand_stride] int64 = the
 * np.random.choice draws (:2888-2893), lm_rand_idx [N,4] (:2676-2683).  global_pos/global_batch >= 0 give the
 * batch-global positive count / size for data-parallel shards (:2864-2868); -1 = this launch is the whole batch;
 * global_pos_ptr (device int, optional) overrides global_pos without a host sync (dbx_count_positives + allreduce).
 * Writes the scalar loss to scalars[0] ("scalars" region, fp32), {half, pos} to scalars[2..3] (int32), the bf16
 * head gradients into "d_head"/"d_rf" (train) and, when given, fp32 gradients [N,60,60,HC] / [N,60,60,16] and the
 * selected masks (uint8 [N,3600] / [N,4,3600]). */
int dbx_net_loss(void* handle, const float* bbox, const float* vertices, const float* labels,
                 const long long* rand_idx, int rand_stride, const long long* lm_rand_idx, float lambda_loc,
                 float lambda_det, float lambda_lm, int global_pos, int global_batch, const int* global_pos_ptr,
                 int clamp_lm, float* d_head_f32, float* d_rf_f32, unsigned char* mask_out,
                 unsigned char* lm_mask_out, void* stream);

/* loss.backward() — DenseBox.py:2925: consumes "d_head"/"d_rf", accumulates (+=) parameter gradients into "g32". */
int dbx_net_backward(void* handle, void* stream);
/* The same backward pass in three stages for data-parallel callers (the reference is single-device, the
 * autograd graph of :2925 is simply cut after conv4_1 and after conv3_1): stage 0 = refine branch + heads + conv4 block,
 * 1 = conv3 block, 2 = conv2 + conv1 blocks.  After stage k the gradients of bucket k are final, so the all-reduce of
 * bucket k overlaps stage k + 1 and only the last (1 MB) bucket is exposed.
 * dbx_net_grad_bucket: element range [first, first+count) of "g32" — bucket 0 = filters conv4_1..heads(+refine),
 * 1 = filters of the conv3 block, 2 = filters conv1_1..conv2_2 + every bias (the flat buffers are laid out in this
 * order).  dbx_net_join: make `stream` wait for the filter re-layout that dbx_net_forward forked onto the engine's
 * side stream (needed when forward is captured into its own CUDA graph).
 * dbx_set_tensor_sm_limit: the persistent tcgen05 kernels launched from now on use at most n SMs (0 = all) — a
 * data-parallel caller leaves a few SMs to the NCCL kernel whose all-reduce overlaps the stage (a persistent kernel
 * with a static tile schedule must be fully resident: if NCCL's CTAs hold SMs it counted on, its last CTAs start
 * late and the whole launch takes up to twice as long).  Returns the previous limit. */
int dbx_net_backward_stage(void* handle, int stage, void* stream);
int dbx_net_grad_bucket(void* handle, int bucket, long long* first, long long* count);
int dbx_net_join(void* handle, void* stream);
int dbx_set_tensor_sm_limit(int n);
int dbx_net_zero_grad(void* handle, void* stream);                       /* optimizer.zero_grad() :2858 */
/* optimizer.step() — torch.optim.SGD(momentum, weight_decay) :2821-2824, :2926; also clears g32 and refreshes the
 * bf16 filters. */
int dbx_net_sgd_step(void* handle, float lr, float momentum, float weight_decay, void* stream);
/* The same update in two launches for data-parallel callers: part 0 = gradient buckets 0 and 1, part 1 = the last
 * bucket (conv1/conv2 filters + biases).  Part 0 runs while the all-reduce of the last bucket is still in flight;
 * call part 0 then part 1, once each per step. */
int dbx_net_sgd_step_part(void* handle, int part, float lr, float momentum, float weight_decay, void* stream);

/* Measurement aids: per-launch CUDA-event timing of the engine's own kernels (eager mode; synchronise before
 * reading) and the number of kernel launches issued so far.  tags: "fprop:<layer>", "dgrad:<layer>",
 * "wgrad:<layer>", "pool_fwd", "loss", ...; flops = algorithmic (unpadded) FLOPs of that launch. */
int dbx_net_profile(void* handle, int enable);
long long dbx_net_launch_count(void* handle);
int dbx_net_profile_count(void* handle);
int dbx_net_profile_get(void* handle, int i, char* tag, int tag_bytes, double* flops, float* ms);

#ifdef __cplusplus
}
#endif
#endif
