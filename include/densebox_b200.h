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

#ifdef __cplusplus
}
#endif
#endif
