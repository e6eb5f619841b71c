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

int num_sms();

// ---- kernels (launchers). All are asynchronous on `stream`, allocate nothing, and return an error code.

struct ConvEpilogue {
  const float* bias = nullptr; // [cout] fp32 or null
  int relu = 0;
  const void* aux = nullptr;   // bf16, pixel-aligned with the output
  int aux_cs = 0, aux_coff = 0;
  int aux_mode = 0;            // 0 none, 1 out = aux>0 ? v : 0 (ReLU backward), 2 out = v*aux (dropout scale)
  int out_fp32 = 0;            // output element type: 0 bf16, 1 fp32
};

// out[n,oh,ow,:] = epi( sum_{r,s,ci} x[n,oh+r-pad,ow+s-pad,ci] * wk[co][(r*S+s)*Cin+ci] )
// x.C must be a multiple of 64 (pad with zero channels); wk is bf16 [cout_rows][R*S*x.C] K-major.
int conv_fprop(const Act& x, const void* wk, int R, int S, int pad, const Act& out, const ConvEpilogue& epi,
               int block_n, cudaStream_t stream);

// dw[co][(r*S+s)*Cin+ci] += sum_{n,oh,ow} dy[n,oh,ow,co] * x[n,oh+r-pad,ow+s-pad,ci]   (fp32, red.add)
// dw has row stride R*S*x.C; rows = dy.C.
int conv_wgrad(const Act& x, const Act& dy, int R, int S, int pad, float* dw, int block_n, cudaStream_t stream);

}  // namespace dbx
