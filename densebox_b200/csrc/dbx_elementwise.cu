// densebox_b200 — HBM-bound kernels around the tensor-core convolutions: input ingest (im2col for the 3-channel
// first layer), 2x2 max-pool forward / backward(+ReLU mask), bilinear(align_corners) upsample forward / backward
// (the upsample half of DenseBox.py:213-219; the concat half is free — conv3_4 writes straight into the fusion
// buffer), bias gradients, weight re-layout, dropout masks, refine-branch glue and the fused SGD step.
// All kernels move 16-byte (8 x bf16) vectors, one vector per thread, channel-fastest so that warps are coalesced.
#include "dbx_common.h"
#include "dbx_ptx.cuh"
#include <stdlib.h>

namespace dbx {

static inline int grid_for(size_t work, int block) { return (int)((work + block - 1) / block); }

__device__ __forceinline__ uint4 ldg16(const bf16* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
  f[0] = bf16lo(u.x); f[1] = bf16hi(u.x); f[2] = bf16lo(u.y); f[3] = bf16hi(u.y);
  f[4] = bf16lo(u.z); f[5] = bf16hi(u.z); f[6] = bf16lo(u.w); f[7] = bf16hi(u.w);
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  return make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
}
__device__ __forceinline__ bf16* at(const Act& a, int n, int y, int x, int c) {
  return reinterpret_cast<bf16*>(a.ptr) + (((size_t)n * a.H + y) * a.W + x) * a.cs + a.coff + c;
}

__device__ __forceinline__ void fold_bias_partials(const float* acc, int cc, float* __restrict__ db, float* red);

// ------------------------------------------------------------------------------------------------ ingest
// X fp32 NCHW [N,3,H,W] -> bf16 [N,H,W,64]: channel (r*3+s)*3+c holds X[n,c,y+r-1,x+s-1] (zero outside), channels
// 27..63 are zero.  conv1_1 (DenseBox.py:185) then is a K=64 1x1 GEMM on the tensor cores.
// mode 0 / 1: 64 channels per pixel (1: also write the zero channels 32..63).
// mode 2 ("pairs"): 32 channels per pixel, so one 128-byte GEMM row holds the taps of TWO horizontally adjacent
// pixels, and channel 27 is the constant 1 (its weight column is zero; in the weight gradient it collects the bias
// gradient).  Half the bytes of the 64-channel layout — this tensor is pure HBM traffic.
__global__ void im2col3x3_c3_kernel(const float* __restrict__ x, bf16* __restrict__ out, int N, int H, int W,
                                    int mode) {
  const size_t pix = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= (size_t)N * H * W) return;
  const int xw = (int)(pix % W), y = (int)((pix / W) % H), n = (int)(pix / ((size_t)W * H));
  float f[32];
#pragma unroll
  for (int k = 27; k < 32; ++k) f[k] = 0.f;
  const float* xn = x + (size_t)n * 3 * H * W;
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const int yy = y + r - 1;
    const bool yok = yy >= 0 && yy < H;
#pragma unroll
    for (int s = 0; s < 3; ++s) {
      const int xx = xw + s - 1;
      const bool ok = yok && xx >= 0 && xx < W;
#pragma unroll
      for (int c = 0; c < 3; ++c) f[(r * 3 + s) * 3 + c] = ok ? __ldg(xn + ((size_t)c * H + yy) * W + xx) : 0.f;
    }
  }
  if (mode == 2) f[27] = 1.f;
  uint4* o = reinterpret_cast<uint4*>(out + pix * (mode == 2 ? 32 : 64));
#pragma unroll
  for (int q = 0; q < 4; ++q) o[q] = pack8(f + 8 * q);
  if (mode == 1) {
#pragma unroll
    for (int q = 4; q < 8; ++q) o[q] = make_uint4(0u, 0u, 0u, 0u);
  }
}

// The same ingest from the decoded image bytes: X uint8 NHWC [N,H,W,3] (what PIL hands over), ToTensor + Normalize
// (DenseBox.py:766-772) applied through a 3 x 256 fp32 table built on the host with torchvision's own arithmetic
// (densebox_b200/data.py::ingest_table), so the result is bit-identical to normalising on the host and calling the
// fp32 kernel — and the batch crosses PCIe as 3 bytes per pixel instead of 12.
__global__ void im2col3x3_c3_u8_kernel(const unsigned char* __restrict__ x, const float* __restrict__ lut,
                                       bf16* __restrict__ out, int N, int H, int W, int mode) {
  __shared__ float tab[3 * 256];
  for (int i = threadIdx.x; i < 3 * 256; i += blockDim.x) tab[i] = lut[i];
  __syncthreads();
  const size_t pix = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= (size_t)N * H * W) return;
  const int xw = (int)(pix % W), y = (int)((pix / W) % H), n = (int)(pix / ((size_t)W * H));
  float f[32];
#pragma unroll
  for (int k = 27; k < 32; ++k) f[k] = 0.f;
  const unsigned char* xn = x + (size_t)n * H * W * 3;
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const int yy = y + r - 1;
    const bool yok = yy >= 0 && yy < H;
#pragma unroll
    for (int s = 0; s < 3; ++s) {
      const int xx = xw + s - 1;
      const bool ok = yok && xx >= 0 && xx < W;
      const unsigned char* px = xn + ((size_t)yy * W + xx) * 3;
#pragma unroll
      for (int c = 0; c < 3; ++c) f[(r * 3 + s) * 3 + c] = ok ? tab[c * 256 + __ldg(px + c)] : 0.f;
    }
  }
  if (mode == 2) f[27] = 1.f;
  uint4* o = reinterpret_cast<uint4*>(out + pix * (mode == 2 ? 32 : 64));
#pragma unroll
  for (int q = 0; q < 4; ++q) o[q] = pack8(f + 8 * q);
  if (mode == 1) {
#pragma unroll
    for (int q = 4; q < 8; ++q) o[q] = make_uint4(0u, 0u, 0u, 0u);
  }
}

int im2col3x3_c3_u8(const unsigned char* x, const float* lut, void* out, int N, int H, int W, int mode, cudaStream_t st) {
  if (!x || !lut || !out || mode < 0 || mode > 2 || (mode == 2 && (W & 1))) return DBX_ERR_ARG;
  const size_t total = (size_t)N * H * W;
  im2col3x3_c3_u8_kernel<<<grid_for(total, 128), 128, 0, st>>>(x, lut, (bf16*)out, N, H, W, mode);
  return (int)cudaGetLastError();
}

// mode 0: channels 32..63 are left untouched (the engine zeroes them once at creation); 1: written; 2: pairs layout.
int im2col3x3_c3(const float* x, void* out, int N, int H, int W, int mode, cudaStream_t st) {
  if (!x || !out || mode < 0 || mode > 2 || (mode == 2 && (W & 1))) return DBX_ERR_ARG;
  const size_t total = (size_t)N * H * W;
  im2col3x3_c3_kernel<<<grid_for(total, 128), 128, 0, st>>>(x, (bf16*)out, N, H, W, mode);
  return (int)cudaGetLastError();
}

// conv1_1 in the pairs layout is a 128 x 64 GEMM matrix [[W 0] [0 W]] (rows 0..63: even pixels, taps in columns
// 0..31; rows 64..127: odd pixels, columns 32..63).  Its weight gradient arrives in `scratch` [128][64] with the two
// diagonal blocks holding the even- and odd-pixel halves of dW, the bias gradient in column 27 of each block (the
// constant-1 tap) and cross terms in the off-diagonal blocks.  Fold: gw[both blocks][c][k] += even + odd (k != 27),
// gb[c] = gb[64 + c] += even + odd of column 27; the off-diagonal blocks and column 27 of gw receive nothing, so the
// two copies stay identical and the zero columns stay zero under SGD.
__global__ void conv1_1_fold_pairs_kernel(const float* __restrict__ scratch, float* __restrict__ gw,
                                          float* __restrict__ gb) {
  const int c = blockIdx.x, k = threadIdx.x;  // 64 blocks x 32 threads
  const float s = scratch[c * 64 + k] + scratch[(64 + c) * 64 + 32 + k];
  if (k == 27) { gb[c] += s; gb[64 + c] += s; }
  else { gw[c * 64 + k] += s; gw[(64 + c) * 64 + 32 + k] += s; }
}

int conv1_1_fold_pairs(const float* scratch, float* gw, float* gb, cudaStream_t st) {
  if (!scratch || !gw || !gb) return DBX_ERR_ARG;
  conv1_1_fold_pairs_kernel<<<64, 32, 0, st>>>(scratch, gw, gb);
  return (int)cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ max-pool 2x2
// idx (optional): 2 bits per pooled element = position (2*dy+dx) of the FIRST maximum of its window, 16 bits per
// (pooled pixel, 8-channel vector).  With it the backward pass needs the pooled map (15 + 15 bytes... one quarter of
// the bytes) instead of re-reading the full-resolution activation to find the arg-max and its ReLU mask.
__global__ void maxpool2x2_fwd_kernel(Act y, Act o, unsigned short* __restrict__ idx_out) {
  const int cc = o.C / 8;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)o.N * o.H * o.W * cc;
  if (idx >= total) return;
  const int c = (int)(idx % cc) * 8;
  const size_t pix = idx / cc;
  const int ox = (int)(pix % o.W), oy = (int)((pix / o.W) % o.H), n = (int)(pix / ((size_t)o.W * o.H));
  float a[8], b[8], c2[8], d[8];
  unpack8(ldg16(at(y, n, 2 * oy, 2 * ox, c)), a);
  unpack8(ldg16(at(y, n, 2 * oy, 2 * ox + 1, c)), b);
  unpack8(ldg16(at(y, n, 2 * oy + 1, 2 * ox, c)), c2);
  unpack8(ldg16(at(y, n, 2 * oy + 1, 2 * ox + 1, c)), d);
  unsigned int bits = 0u;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    unsigned int arg = 0u; float m = a[j];
    if (b[j] > m) { m = b[j]; arg = 1u; }
    if (c2[j] > m) { m = c2[j]; arg = 2u; }
    if (d[j] > m) { m = d[j]; arg = 3u; }
    a[j] = m;
    bits |= arg << (2 * j);
  }
  *reinterpret_cast<uint4*>(at(o, n, oy, ox, c)) = pack8(a);
  if (idx_out) idx_out[idx] = (unsigned short)bits;
}

int maxpool2x2_fwd(const Act& y, const Act& o, cudaStream_t st, void* idx) {
  if (y.H != 2 * o.H || y.W != 2 * o.W || y.C != o.C || y.N != o.N || y.C % 8) return DBX_ERR_ARG;
  const size_t total = (size_t)o.N * o.H * o.W * (o.C / 8);
  maxpool2x2_fwd_kernel<<<grid_for(total, 256), 256, 0, st>>>(y, o, (unsigned short*)idx);
  return (int)cudaGetLastError();
}

// Backward from the compact arg-max map written by maxpool2x2_fwd(idx): dy[k] = (k == arg && p > 0) ? dp : 0 — the
// ReLU mask of the producing conv only matters at the arg-max, whose value IS the pooled value p.  Reads 16 + 16 + 2
// bytes and writes 64 per item (the full-resolution read of y, 64 bytes, is gone).  db as in maxpool2x2_bwd.
__global__ void __launch_bounds__(256) maxpool2x2_bwd_idx_kernel(Act p, Act dp, const unsigned short* __restrict__ idx_in,
                                                                 Act dy, float* __restrict__ db) {
  __shared__ float red[256 * 8];
  const int cc = dp.C / 8;
  const size_t total = (size_t)dp.N * dp.H * dp.W * cc;
  float bsum[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) bsum[j] = 0.f;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(idx % cc) * 8;
    const size_t pix = idx / cc;
    const int ox = (int)(pix % dp.W), oy = (int)((pix / dp.W) % dp.H), n = (int)(pix / ((size_t)dp.W * dp.H));
    float pv[8], g[8];
    unpack8(ldg16(at(p, n, oy, ox, c)), pv);
    const uint4 gq = ldg16(at(dp, n, oy, ox, c));
    unpack8(gq, g);
    const unsigned int bits = idx_in[idx];
    float o[4][8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const unsigned int arg = (bits >> (2 * j)) & 3u;
      const float v = pv[j] > 0.f ? g[j] : 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k) o[k][j] = (unsigned int)k == arg ? v : 0.f;
      bsum[j] += v;  // dp is bf16 already: what is stored equals v exactly
    }
    *reinterpret_cast<uint4*>(at(dy, n, 2 * oy, 2 * ox, c)) = pack8(o[0]);
    *reinterpret_cast<uint4*>(at(dy, n, 2 * oy, 2 * ox + 1, c)) = pack8(o[1]);
    *reinterpret_cast<uint4*>(at(dy, n, 2 * oy + 1, 2 * ox, c)) = pack8(o[2]);
    *reinterpret_cast<uint4*>(at(dy, n, 2 * oy + 1, 2 * ox + 1, c)) = pack8(o[3]);
  }
  if (db) fold_bias_partials(bsum, cc, db, red);
}

int maxpool2x2_bwd_idx(const Act& p, const Act& dp, const void* idx, const Act& dy, cudaStream_t st, float* db) {
  if (!idx || dy.H != 2 * dp.H || dy.W != 2 * dp.W || dy.C != dp.C || p.C != dp.C || p.H != dp.H || p.W != dp.W ||
      dp.C % 8)
    return DBX_ERR_ARG;
  const int cc = dp.C / 8;
  if (256 % cc) return DBX_ERR_ARG;
  const size_t total = (size_t)dp.N * dp.H * dp.W * cc;
  int blocks = grid_for(total, 256);
  if (blocks > 16 * num_sms()) blocks = 16 * num_sms();
  maxpool2x2_bwd_idx_kernel<<<blocks, 256, 0, st>>>(p, dp, (const unsigned short*)idx, dy, db);
  return (int)cudaGetLastError();
}

// Block-level fold of per-thread channel partials into db (bias gradient): thread t owns channels 8 (t % cc) .. + 7.
__device__ __forceinline__ void fold_bias_partials(const float* acc, int cc, float* __restrict__ db, float* red) {
  // red: [blockDim.x][8] floats of shared memory
#pragma unroll
  for (int j = 0; j < 8; ++j) red[threadIdx.x * 8 + j] = acc[j];
  __syncthreads();
  const int C = cc * 8, reps = (int)blockDim.x / cc;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.f;
    for (int k = 0; k < reps; ++k) s += red[(k * cc + c / 8) * 8 + (c & 7)];
    if (s != 0.f) atomicAdd(db + c, s);
  }
}

// dy = relu'(y) * ( maxpool_bwd(dp)  [+ add] ): gradient goes to the FIRST maximum of each window in scan order
// (torch max_pool2d semantics), then the ReLU mask of the producing conv (y > 0) is applied.
// db (optional): db[c] += sum over pixels of dy[.., c] as stored — the bias gradient of the conv that produced y,
// folded here so that dy is not read from HBM a second time.  Threads keep their channel vector across the
// grid-stride loop (the stride is a multiple of C/8).
__global__ void __launch_bounds__(256) maxpool2x2_bwd_kernel(Act y, Act dp, Act add, int has_add, Act dy,
                                                             float* __restrict__ db) {
  __shared__ float red[256 * 8];
  const int cc = dp.C / 8;
  const size_t total = (size_t)dp.N * dp.H * dp.W * cc;
  float bsum[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) bsum[j] = 0.f;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(idx % cc) * 8;
    const size_t pix = idx / cc;
    const int ox = (int)(pix % dp.W), oy = (int)((pix / dp.W) % dp.H), n = (int)(pix / ((size_t)dp.W * dp.H));
    float v[4][8], g[8], o[4][8];
    unpack8(ldg16(at(y, n, 2 * oy, 2 * ox, c)), v[0]);
    unpack8(ldg16(at(y, n, 2 * oy, 2 * ox + 1, c)), v[1]);
    unpack8(ldg16(at(y, n, 2 * oy + 1, 2 * ox, c)), v[2]);
    unpack8(ldg16(at(y, n, 2 * oy + 1, 2 * ox + 1, c)), v[3]);
    unpack8(ldg16(at(dp, n, oy, ox, c)), g);
    if (has_add) {
      unpack8(ldg16(at(add, n, 2 * oy, 2 * ox, c)), o[0]);
      unpack8(ldg16(at(add, n, 2 * oy, 2 * ox + 1, c)), o[1]);
      unpack8(ldg16(at(add, n, 2 * oy + 1, 2 * ox, c)), o[2]);
      unpack8(ldg16(at(add, n, 2 * oy + 1, 2 * ox + 1, c)), o[3]);
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int j = 0; j < 8; ++j) o[k][j] = 0.f;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      int arg = 0; float m = v[0][j];
      if (v[1][j] > m) { m = v[1][j]; arg = 1; }
      if (v[2][j] > m) { m = v[2][j]; arg = 2; }
      if (v[3][j] > m) { m = v[3][j]; arg = 3; }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float r = o[k][j] + (k == arg ? g[j] : 0.f);
        o[k][j] = v[k][j] > 0.f ? r : 0.f;
      }
    }
    uint4 q[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) q[k] = pack8(o[k]);
    *reinterpret_cast<uint4*>(at(dy, n, 2 * oy, 2 * ox, c)) = q[0];
    *reinterpret_cast<uint4*>(at(dy, n, 2 * oy, 2 * ox + 1, c)) = q[1];
    *reinterpret_cast<uint4*>(at(dy, n, 2 * oy + 1, 2 * ox, c)) = q[2];
    *reinterpret_cast<uint4*>(at(dy, n, 2 * oy + 1, 2 * ox + 1, c)) = q[3];
    if (db) {  // sum what was stored (bf16-rounded), like the stand-alone colsum kernel would
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float f[8];
        unpack8(q[k], f);
#pragma unroll
        for (int j = 0; j < 8; ++j) bsum[j] += f[j];
      }
    }
  }
  if (db) fold_bias_partials(bsum, cc, db, red);
}

int maxpool2x2_bwd(const Act& y, const Act& dp, const Act* add, const Act& dy, cudaStream_t st, float* db) {
  if (y.H != 2 * dp.H || y.W != 2 * dp.W || y.C != dp.C || dy.C != y.C || dy.H != y.H || dy.W != y.W || y.C % 8)
    return DBX_ERR_ARG;
  const int cc = dp.C / 8;
  const size_t total = (size_t)dp.N * dp.H * dp.W * cc;
  if (db && (256 % cc) != 0) {  // threads could not keep their channels: separate pass
    const int rc = maxpool2x2_bwd(y, dp, add, dy, st, nullptr);
    return rc ? rc : colsum(dy, db, st);
  }
  int blocks = grid_for(total, 256);
  if (blocks > 16 * num_sms()) blocks = 16 * num_sms();
  maxpool2x2_bwd_kernel<<<blocks, 256, 0, st>>>(y, dp, add ? *add : y, add ? 1 : 0, dy, db);
  return (int)cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ bilinear upsample
// nn.Upsample(size, mode='bilinear', align_corners=True) (DenseBox.py:213-216, :468-470): src = dst*(in-1)/(out-1),
// float arithmetic identical to ATen's area_pixel_compute_source_index / linear weights.
struct Lin { int i0, i1; float l0, l1; };
__device__ __forceinline__ Lin lin_coord(int d, float scale, int in_size) {
  Lin r;
  const float src = scale * (float)d;
  r.i0 = (int)src;
  r.i1 = r.i0 + (r.i0 < in_size - 1 ? 1 : 0);
  r.l1 = src - (float)r.i0;
  r.l0 = 1.f - r.l1;
  return r;
}

// Forward.  A thread owns one 8-channel vector of a run of `len` horizontally adjacent output pixels of one output row
// and walks the run left to right holding the two VERTICALLY blended input columns it sits between in registers
// (bilinear interpolation is separable); a column is loaded once per run — about one 2 x 16-byte load pair per two
// outputs instead of four loads per output.  The round-1 kernel (4 loads + 4 unpacks + 6 FMA groups per output, an
// integer division per item, 1.6 waves of one-row blocks) issued at 72 % of the SM's instruction rate and reached
// 42 % of the HBM roofline; the instruction count per output is less than half here and the run length is chosen so
// that the grid fills whole waves.
__global__ void __launch_bounds__(256) upsample_fwd_kernel(Act in, Act out, float sh, float sw, int len, int nseg,
                                                           size_t total) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int cc = out.C / 8;
  const int cvec = (int)(idx % cc);
  size_t t = idx / cc;
  const int seg = (int)(t % nseg); t /= nseg;
  const int oy = (int)(t % out.H), n = (int)(t / out.H);
  const int c = cvec * 8;
  const Lin ly = lin_coord(oy, sh, in.H);
  const bf16* r0 = at(in, n, ly.i0, 0, c);
  const bf16* r1 = at(in, n, ly.i1, 0, c);
  bf16* orow = at(out, n, oy, 0, c);
  const int x0 = seg * len;
  int x1 = x0 + len; if (x1 > out.W) x1 = out.W;
  float va[8], vb[8];
  auto loadblend = [&](int ix, float* v) {
    float a[8], b[8];
    unpack8(ldg16(r0 + (size_t)ix * in.cs), a);
    unpack8(ldg16(r1 + (size_t)ix * in.cs), b);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = ly.l0 * a[j] + ly.l1 * b[j];
  };
  Lin lx = lin_coord(x0, sw, in.W);
  int ci0 = lx.i0, ci1 = lx.i1;
  loadblend(ci0, va);
  if (ci1 != ci0) loadblend(ci1, vb);
  else {
#pragma unroll
    for (int j = 0; j < 8; ++j) vb[j] = va[j];
  }
  for (int ox = x0; ox < x1; ++ox) {
    lx = lin_coord(ox, sw, in.W);
    if (lx.i0 != ci0) {
      if (lx.i0 == ci1) {
#pragma unroll
        for (int j = 0; j < 8; ++j) va[j] = vb[j];
      } else {
        loadblend(lx.i0, va);
      }
      ci0 = lx.i0;
    }
    if (lx.i1 != ci1) {
      if (lx.i1 == ci0) {
#pragma unroll
        for (int j = 0; j < 8; ++j) vb[j] = va[j];
      } else {
        loadblend(lx.i1, vb);
      }
      ci1 = lx.i1;
    }
    float o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = lx.l0 * va[j] + lx.l1 * vb[j];
    *reinterpret_cast<uint4*>(orow + (size_t)ox * out.cs) = pack8(o);
  }
}

// Run length for the walking kernels: as long as possible (fewer redundant loads at the run boundaries) while the grid
// still fills (nearly) whole waves of `slots` resident blocks.
static int choose_run(int W, size_t items_per_run, int block, int slots, int min_len) {
  int best = W, best_nseg = 1;
  double best_cost = 1e30;
  for (int nseg = 1; nseg <= W; ++nseg) {
    const int len = (W + nseg - 1) / nseg;
    if (len < min_len && nseg > 1) break;
    const int ns = (W + len - 1) / len;
    const double blocks = (double)((items_per_run * ns + block - 1) / block);
    const double waves = blocks / slots;
    // latency hiding needs a few resident blocks per slot over the kernel's life: below ~2.5 waves the kernel is
    // latency-bound (measured: one 60-pixel run per thread = 0.65 waves ran at 38 % warp occupancy, 40 us)
    const double quant = (waves < 2.5 ? 2.5 / (waves > 0.05 ? waves : 0.05) : 1.0) * ceil(waves) / waves;
    const double cost = quant * (1.0 + 1.5 / len);
    if (cost < best_cost) { best_cost = cost; best = len; best_nseg = ns; }
  }
  (void)best_nseg;
  return best;
}

int upsample_bilinear_fwd(const Act& in, const Act& out, cudaStream_t st) {
  if (in.C != out.C || in.N != out.N || in.C % 8) return DBX_ERR_ARG;
  const float sh = out.H > 1 ? (float)(in.H - 1) / (float)(out.H - 1) : 0.f;
  const float sw = out.W > 1 ? (float)(in.W - 1) / (float)(out.W - 1) : 0.f;
  const int cc = out.C / 8;
  static int occ[kMaxDevices] = {0};
  int dev = 0; cudaGetDevice(&dev); if (dev < 0 || dev >= kMaxDevices) dev = 0;
  if (!occ[dev]) {
    int b = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, upsample_fwd_kernel, 256, 0);
    occ[dev] = b > 0 ? b : 4;
  }
  const int len = choose_run(out.W, (size_t)out.N * out.H * cc, 256, occ[dev] * num_sms(), 4);
  const int nseg = (out.W + len - 1) / len;
  const size_t total = (size_t)out.N * out.H * nseg * cc;
  upsample_fwd_kernel<<<grid_for(total, 256), 256, 0, st>>>(in, out, sh, sw, len, nseg, total);
  return (int)cudaGetLastError();
}

// Backward.  (A register-walking adjoint of the forward kernel — two input rows per thread, vertical taps first — was
// measured in round 2 at 0.128 ms against 0.089 ms for this gather: 120 registers, 16 warps per SM and one exposed
// memory round trip per output column made it latency-bound at 25 % issue utilisation; the gather below stays.)
// din[n,iy,ix,:] = relu'(y) * sum over output pixels of their bilinear weight on (iy,ix) — a gather, so no atomics.
// One block per input row (n, iy).  The (output index, weight) lists of the row and of every input column are built
// once per block in shared memory (a few entries each: an input pixel is touched by <= ceil(2/scale)+1 outputs per
// axis), so the inner loop is loads and FMAs only; contributions are added in ascending (oy, ox) order.
static constexpr int kUpMaxTaps = 8;
static constexpr int kUpMaxW = 128;
__global__ void __launch_bounds__(256) upsample_bwd_kernel(Act dout, Act y, int has_mask, Act din, float sh, float sw,
                                                           float* __restrict__ db) {
  __shared__ float red[256 * 8];
  __shared__ int ycnt, yidx[kUpMaxTaps];
  __shared__ float ywt[kUpMaxTaps];
  __shared__ int xcnt[kUpMaxW], xidx[kUpMaxW][kUpMaxTaps];
  __shared__ float xwt[kUpMaxW][kUpMaxTaps];
  const int cc = din.C / 8;
  const int iy = (int)(blockIdx.x % din.H), n = (int)(blockIdx.x / din.H);
  // candidates: outputs whose source coordinate lies within (i - 1, i + 1), +-1 for rounding
  if (threadIdx.x == 0) {
    int k = 0, lo = 0, hi = dout.H - 1;
    if (sh > 0.f) { lo = max(0, (int)floorf((iy - 1) / sh) - 1); hi = min(dout.H - 1, (int)ceilf((iy + 1) / sh) + 1); }
    for (int oy = lo; oy <= hi; ++oy) {
      const Lin l = lin_coord(oy, sh, din.H);
      const float w = (l.i0 == iy ? l.l0 : 0.f) + (l.i1 == iy ? l.l1 : 0.f);
      if (w != 0.f && k < kUpMaxTaps) { yidx[k] = oy; ywt[k] = w; ++k; }
    }
    ycnt = k;
  }
  for (int ix = threadIdx.x; ix < din.W; ix += 256) {
    int k = 0, lo = 0, hi = dout.W - 1;
    if (sw > 0.f) { lo = max(0, (int)floorf((ix - 1) / sw) - 1); hi = min(dout.W - 1, (int)ceilf((ix + 1) / sw) + 1); }
    for (int ox = lo; ox <= hi; ++ox) {
      const Lin l = lin_coord(ox, sw, din.W);
      const float w = (l.i0 == ix ? l.l0 : 0.f) + (l.i1 == ix ? l.l1 : 0.f);
      if (w != 0.f && k < kUpMaxTaps) { xidx[ix][k] = ox; xwt[ix][k] = w; ++k; }
    }
    xcnt[ix] = k;
  }
  __syncthreads();
  const int total = din.W * cc;
  const int ny = ycnt;
  float bsum[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) bsum[j] = 0.f;
  for (int v = threadIdx.x; v < total; v += 256) {
    const int ix = v / cc, c = (v - ix * cc) * 8;
    const int nx = xcnt[ix];
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    for (int a = 0; a < ny; ++a) {
      const bf16* row = at(dout, n, yidx[a], 0, c);
      const float wy = ywt[a];
      for (int b = 0; b < nx; ++b) {
        float g[8];
        unpack8(ldg16(row + (size_t)xidx[ix][b] * dout.cs), g);
        const float w = wy * xwt[ix][b];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += w * g[j];
      }
    }
    if (has_mask) {
      float m[8];
      unpack8(ldg16(at(y, n, iy, ix, c)), m);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = m[j] > 0.f ? acc[j] : 0.f;
    }
    const uint4 q = pack8(acc);
    *reinterpret_cast<uint4*>(at(din, n, iy, ix, c)) = q;
    if (db) {  // bias gradient of the conv that produced y: sum of what was stored (256 % (C/8) == 0: fixed channels)
      float f[8];
      unpack8(q, f);
#pragma unroll
      for (int j = 0; j < 8; ++j) bsum[j] += f[j];
    }
  }
  if (db) fold_bias_partials(bsum, cc, db, red);
}

int upsample_bilinear_bwd(const Act& dout, const Act* relu_y, const Act& din, cudaStream_t st, float* db) {
  if (din.C != dout.C || din.N != dout.N || din.C % 8 || din.W > kUpMaxW) return DBX_ERR_ARG;
  // an input pixel receives from at most ceil(2 * out / in) + 1 output pixels per axis
  if (2 * ((dout.H + din.H - 1) / din.H) + 1 > kUpMaxTaps || 2 * ((dout.W + din.W - 1) / din.W) + 1 > kUpMaxTaps)
    return DBX_ERR_ARG;
  const float sh = dout.H > 1 ? (float)(din.H - 1) / (float)(dout.H - 1) : 0.f;
  const float sw = dout.W > 1 ? (float)(din.W - 1) / (float)(dout.W - 1) : 0.f;
  const int cc = din.C / 8;
  const bool fold = db && (256 % cc) == 0;
  upsample_bwd_kernel<<<din.N * din.H, 256, 0, st>>>(dout, relu_y ? *relu_y : din, relu_y ? 1 : 0, din, sh, sw,
                                                         fold ? db : nullptr);
  int rc = (int)cudaGetLastError();
  if (!rc && db && !fold) rc = colsum(din, db, st);
  return rc;
}

// ------------------------------------------------------------------------------------------------ bias gradient
// db[c] += sum over pixels of dy[pixel][c].  blockDim = (C/8) * rows; thread (r, chunk) keeps 8 fp32 partials over its
// pixels (8 loads in flight), the block folds them through shared memory and issues ONE atomic per channel: the
// atomics to the same few addresses serialise in L2, so their count (blocks x C) is what bounds the small layers —
// a shuffle-only variant with one vector red per warp was 2.4x slower (session 3).
__global__ void __launch_bounds__(256) colsum_kernel(Act dy, float* __restrict__ db, int rows, size_t pixels,
                                                     size_t pix_per_block) {
  extern __shared__ float red[];  // [rows][C]
  const int cc = dy.C / 8;
  const int chunk = threadIdx.x % cc, r = threadIdx.x / cc;
  const size_t p0 = (size_t)blockIdx.x * pix_per_block;
  size_t p1 = p0 + pix_per_block; if (p1 > pixels) p1 = pixels;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  const bf16* base = reinterpret_cast<const bf16*>(dy.ptr) + dy.coff + chunk * 8;
  size_t p = p0 + r;
  for (; p + 7 * (size_t)rows < p1; p += 8 * (size_t)rows) {
    uint4 u[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) u[q] = ldg16(base + (p + q * (size_t)rows) * dy.cs);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      float f[8];
      unpack8(u[q], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += f[j];
    }
  }
  for (; p < p1; p += rows) {
    float f[8];
    unpack8(ldg16(base + p * dy.cs), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] += f[j];
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) red[(size_t)r * dy.C + chunk * 8 + j] = acc[j];
  __syncthreads();
  for (int c = threadIdx.x; c < dy.C; c += blockDim.x) {
    float s = 0.f;
    for (int k = 0; k < rows; ++k) s += red[(size_t)k * dy.C + c];
    atomicAdd(db + c, s);
  }
}

// ---- conv5_2 o conv5_1 folded into one 768 -> HC matrix (inference: no dropout, and the reference has no
// non-linearity between the two 1x1 convolutions, DenseBox.py:158-178): wf[o][k] = sum_c w2[o][c] w1[c][k] over the 512
// hidden channels c of the branch that owns output o, bf[o] = b2[o] + sum_c w2[o][c] b1[c].  fp32 masters in, bf16
// GEMM operand + fp32 bias out.  Grid (HC, 25): y < 24 -> 32 columns k (lane) x 8 slices of the hidden channels (warp),
// 64 dependent-free FMAs per thread in batches of 8 loads (the kernel is pure load latency), summed in shared memory
// in a fixed order; y == 24 -> the bias.
struct FoldStarts { int s[5]; };
__global__ void __launch_bounds__(256)
heads_fold_kernel(const float* __restrict__ w1, int ld1, const float* __restrict__ b1, const float* __restrict__ w2,
                  int ld2, const float* __restrict__ b2, FoldStarts st, int nh, bf16* __restrict__ wf,
                  float* __restrict__ bf) {
  __shared__ float red[8][32];
  const int o = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int b = -1;
  for (int i = 0; i < nh; ++i)
    if (o >= st.s[i] && o < st.s[i + 1]) b = i;
  const float* w2r = w2 + (size_t)o * ld2 + 512 * (b < 0 ? 0 : b);
  float a = 0.f;
  if (blockIdx.y < 24) {
    const int k = blockIdx.y * 32 + lane;
    if (b >= 0) {
      const float* w1c = w1 + ((size_t)512 * b + 64 * warp) * ld1 + k;
#pragma unroll 8
      for (int c = 0; c < 64; ++c) a = fmaf(__ldg(w2r + 64 * warp + c), __ldg(w1c + (size_t)c * ld1), a);
    }
    red[warp][lane] = a;
    __syncthreads();
    if (warp == 0) {
      float v = 0.f;
#pragma unroll
      for (int g = 0; g < 8; ++g) v += red[g][lane];
      wf[(size_t)o * ld1 + k] = __float2bfloat16(v);
    }
  } else {
    if (b >= 0)
      for (int c = threadIdx.x; c < 512; c += 256) a = fmaf(__ldg(w2r + c), __ldg(b1 + 512 * b + c), a);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) a += __shfl_xor_sync(0xffffffffu, a, d);
    if (lane == 0) red[warp][0] = a;
    __syncthreads();
    if (threadIdx.x == 0) {
      float v = b >= 0 ? __ldg(b2 + o) : 0.f;
      for (int g = 0; g < 8; ++g) v += red[g][0];
      bf[o] = v;
    }
  }
}

int heads_fold(const float* w1, int ld1, const float* b1, const float* w2, int ld2, const float* b2, const int* start,
               int nh, int HC, void* wf, float* bf, cudaStream_t st) {
  if (!w1 || !b1 || !w2 || !b2 || !start || !wf || !bf || ld1 != 768 || nh < 1 || nh > 4 || HC < 1) return DBX_ERR_ARG;
  FoldStarts fs;
  for (int i = 0; i < 5; ++i) fs.s[i] = start[i];
  heads_fold_kernel<<<dim3(HC, 25), 256, 0, st>>>(w1, ld1, b1, w2, ld2, b2, fs, nh, (bf16*)wf, bf);
  return (int)cudaGetLastError();
}

int colsum(const Act& dy, float* db, cudaStream_t st) {
  if (!db || dy.C % 8 || dy.C > 2048) return DBX_ERR_ARG;
  { const char* e = ab_env("DBX_NO_COLSUM"); if (e && e[0] == '1') return DBX_OK; }  // measurement only
  const int cc = dy.C / 8;
  int rows = 256 / cc; if (rows < 1) rows = 1;
  const size_t pixels = (size_t)dy.N * dy.H * dy.W;
  int blocks = 8 * num_sms();  // measured: fewer, fatter blocks are slower even on the 29 MB layers (latency-bound)
  { const char* e = ab_env("DBX_COLSUM_BLOCKS"); if (e && atoi(e) > 0) blocks = atoi(e) * num_sms(); }
  if ((size_t)blocks * rows > pixels) blocks = (int)((pixels + rows - 1) / rows);
  if (blocks < 1) blocks = 1;
  const size_t ppb = (pixels + blocks - 1) / blocks;
  colsum_kernel<<<blocks, cc * rows, (size_t)rows * dy.C * sizeof(float), st>>>(dy, db, rows, pixels, ppb);
  return (int)cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ heads: dgrad of conv5_2
// d_hd[pix][i] = drop(pix, i) * sum_c d_head[pix][c] * W2[c][i]  (DenseBox.py:149-178 backward: conv5_2_* then
// Dropout).  W2 is block-diagonal (head h owns rows ch_start[h] .. ch_start[h+1]-1 and columns 512h .. 512h+511), so a
// column needs at most 8 products: this is a 236 MB store, not a GEMM.  As a K = 64 tensor-core launch it was paced by
// its epilogue (0.109 ms = 2.2 TB/s, plus 0.052 ms for the stand-alone column sums of d_hd that give the bias gradient
// of conv5_1).  Here a thread owns 16 channels (one 32-byte sector per store pair) of TWO pixels per iteration, the
// head's weights sit in shared memory as conflict-free float4 rows, ONE Philox call serves the 128 channels of eight
// neighbouring threads (warp shuffles; the round-1 version drew it per thread and was instruction-bound at 0.18 ms),
// dropout acts on packed bf16x2 words, and the column sums of what is stored come out of the same pass.
struct Heads2DgradParams {
  const bf16* d_head;      // [pixels][64] bf16 (channels >= HC are zero)
  const bf16* wd;          // [K][64] bf16: wd[i][c] = W2[c][i]
  bf16* out;               // [pixels][K]
  const bf16* mask;        // drop_mode 2: [pixels][K] {0,2}
  const unsigned long long* rng;  // drop_mode 3: {seed, offset}
  float* db;               // [K] or null
  int K, nh, drop_mode;
  int ch_start[5];
  size_t pixels;
};

// 16 keep-bits -> 8 packed bf16x2 words scaled by {0, 2} (x2 and x0 are exact in bf16)
__device__ __forceinline__ void dropout8(uint32_t* pk, uint32_t bits) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    // bit 2i -> 0x4000 (bf16 2.0, low half), bit 2i+1 -> 0x40000000 (high half): two shifts and one select-by-mask
    const uint32_t a = (14 - 2 * i) >= 0 ? (bits << ((14 - 2 * i) & 31)) : (bits >> ((2 * i - 14) & 31));
    const uint32_t b = bits << ((29 - 2 * i) & 31);
    const uint32_t m = ((a & 0x00004000u) | (b & 0x40000000u));
    __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&pk[i]);
    const __nv_bfloat162 mm = *reinterpret_cast<const __nv_bfloat162*>(&m);
    v = __hmul2(v, mm);
    pk[i] = *reinterpret_cast<uint32_t*>(&v);
  }
}

__global__ void __launch_bounds__(256) heads2_dgrad_kernel(const Heads2DgradParams p) {
  extern __shared__ float4 ws4[];  // [(c * 4 + q) * tpp + t16]: channels 16 t16 + 4 q .. + 3 of product row c
  __shared__ float red[256 * 8];
  const int K = p.K;
  const int tpp = K / 16;                       // threads per pixel
  const int ppb = (int)blockDim.x / tpp;        // pixel slots per block
  float* wsf = reinterpret_cast<float*>(ws4);
  for (int e = threadIdx.x; e < 8 * K; e += blockDim.x) {
    const int c = e / K, i = e - c * K, h = i >> 9;
    int c0 = 0, nc = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) if (h == k) { c0 = p.ch_start[k]; nc = p.ch_start[k + 1] - c0; }
    const float v = c < nc ? __bfloat162float(p.wd[(size_t)i * 64 + c0 + c]) : 0.f;
    wsf[(((c * 4 + ((i >> 2) & 3)) * tpp + (i >> 4)) << 2) + (i & 3)] = v;
  }
  __syncthreads();
  const int slot = (int)threadIdx.x / tpp, t16 = (int)threadIdx.x - slot * tpp, i0 = t16 * 16;
  const int h = i0 >> 9;
  int c0 = 0, nc = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) if (h == k) { c0 = p.ch_start[k]; nc = p.ch_start[k + 1] - c0; }
  unsigned long long seed = 0, off = 0;
  if (p.drop_mode == 3) { seed = p.rng[0]; off = p.rng[1]; }
  const int lane = threadIdx.x & 31;
  // Philox: one call yields the keep-bits of 128 channels = the 8 lanes of a group.  A warp handles its 512 channels
  // of EIGHT pixels per batch: lane l draws for pixel (l & 7) and channel group (l >> 3), i.e. every lane does one
  // useful call per batch (a leader-only call would still cost the whole warp its ~110 instructions), and the lanes of
  // a group fetch the draw of pixel j from lane (group * 8 + j) with shuffles.
  const int grp0 = lane & ~7;
  const uint32_t wsel = (uint32_t)(lane & 7) >> 1, wshift = (uint32_t)(lane & 1) * 16u;
  const unsigned long long wch = (unsigned long long)((t16 - lane) * 16 + 128 * (lane >> 3));  // first channel of the lane's draw
  float bsum[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) bsum[j] = 0.f;
  if (slot < ppb)
  for (size_t base = ((size_t)blockIdx.x * ppb + slot) * 8; base < p.pixels; base += (size_t)gridDim.x * ppb * 8) {
    uint4 r = make_uint4(0u, 0u, 0u, 0u);
    if (p.drop_mode == 3) {
      const unsigned long long cnt = (((base + (lane & 7)) * (unsigned long long)K + wch) >> 7) + off;
      r = philox4x32(make_uint4((uint32_t)cnt, (uint32_t)(cnt >> 32), 0u, 0u),
                     make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    }
    const bf16* g0 = p.d_head + base * 64 + c0;
    bf16* op = p.out + base * K + i0;
    const bf16* mp = p.mask ? p.mask + base * K + i0 : nullptr;
#pragma unroll 1
    for (int jp = 0; jp < 8; jp += 2, g0 += 128, op += 2 * (size_t)K) {  // two pixels per trip share the weight loads
      if (base + jp >= p.pixels) break;                                  // (one pixel per trip: 80 registers, 3 blocks
      const bool two = base + jp + 1 < p.pixels;                         //  per SM, but 0.111 ms against 0.093)
      float acc[2][16];
#pragma unroll
      for (int j = 0; j < 16; ++j) { acc[0][j] = 0.f; acc[1][j] = 0.f; }
      const bf16* g1 = g0 + (two ? 64 : 0);
      for (int c = 0; c < nc; ++c) {
        const float ga = __bfloat162float(g0[c]), gb = __bfloat162float(g1[c]);
        const float4* w4 = ws4 + (c * 4) * tpp + t16;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 w = w4[q * tpp];
          acc[0][4 * q] += ga * w.x; acc[0][4 * q + 1] += ga * w.y; acc[0][4 * q + 2] += ga * w.z; acc[0][4 * q + 3] += ga * w.w;
          acc[1][4 * q] += gb * w.x; acc[1][4 * q + 1] += gb * w.y; acc[1][4 * q + 2] += gb * w.z; acc[1][4 * q + 3] += gb * w.w;
        }
      }
#pragma unroll
      for (int s2 = 0; s2 < 2; ++s2) {
        uint32_t pk[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) pk[j] = pack_bf16x2(acc[s2][2 * j], acc[s2][2 * j + 1]);
        if (p.drop_mode == 3) {
          const int src = grp0 | (jp + s2);
          const uint32_t rx = __shfl_sync(0xffffffffu, r.x, src), ry = __shfl_sync(0xffffffffu, r.y, src);
          const uint32_t rz = __shfl_sync(0xffffffffu, r.z, src), rw = __shfl_sync(0xffffffffu, r.w, src);
          const uint32_t w = wsel == 0 ? rx : (wsel == 1 ? ry : (wsel == 2 ? rz : rw));
          dropout8(pk, (w >> wshift) & 0xFFFFu);
        } else if (p.drop_mode == 2) {
          const bf16* m = mp + (size_t)(jp + (two ? s2 : 0)) * K;
          const uint4 m0 = ldg16(m), m1 = ldg16(m + 8);
          const uint32_t mm[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            __nv_bfloat162 v = __hmul2(*reinterpret_cast<__nv_bfloat162*>(&pk[j]), *reinterpret_cast<const __nv_bfloat162*>(&mm[j]));
            pk[j] = *reinterpret_cast<uint32_t*>(&v);
          }
        }
        if (s2 == 0 || two) {
          uint4* o = reinterpret_cast<uint4*>(op + (size_t)s2 * K);
          o[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          o[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
          if (p.db) {
#pragma unroll
            for (int j = 0; j < 8; ++j) { bsum[2 * j] += bf16lo(pk[j]); bsum[2 * j + 1] += bf16hi(pk[j]); }
          }
        }
      }
    }
  }
  if (p.db) {  // fold the pixel slots of the block, then one atomic per channel
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      __syncthreads();
#pragma unroll
      for (int j = 0; j < 8; ++j) red[threadIdx.x * 8 + j] = bsum[half * 8 + j];
      __syncthreads();
      if (slot == 0) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float sum = 0.f;
          for (int k = 0; k < ppb; ++k) sum += red[((size_t)k * tpp + threadIdx.x) * 8 + j];
          if (sum != 0.f) atomicAdd(p.db + i0 + half * 8 + j, sum);
        }
      }
    }
  }
}

int heads2_dgrad(const void* d_head, const void* wd, void* out, size_t pixels, int K, int nh, const int* ch_start,
                 int drop_mode, const void* mask, const unsigned long long* rng, float* db, cudaStream_t st) {
  if (!d_head || !wd || !out || !ch_start || nh < 1 || nh > 4 || K != 512 * nh) return DBX_ERR_ARG;
  if ((drop_mode == 2 && !mask) || (drop_mode == 3 && !rng) || drop_mode == 1 || drop_mode < 0 || drop_mode > 3)
    return DBX_ERR_ARG;
  Heads2DgradParams p{};
  p.d_head = (const bf16*)d_head; p.wd = (const bf16*)wd; p.out = (bf16*)out; p.mask = (const bf16*)mask;
  p.rng = rng; p.db = db; p.K = K; p.nh = nh; p.drop_mode = drop_mode; p.pixels = pixels;
  for (int i = 0; i < 5; ++i) p.ch_start[i] = ch_start[i];
  for (int h = 0; h < nh; ++h) if (ch_start[h + 1] - ch_start[h] > 8 || ch_start[h + 1] > 64) return DBX_ERR_ARG;
  const int tpp = K / 16;                         // 32 / 64 / 96 / 128: whole warps, so the Philox shuffles stay inside a pixel
  const int threads = tpp * (256 / tpp > 0 ? 256 / tpp : 1);
  if (threads > 256 || tpp % 32) return DBX_ERR_ARG;
  const size_t smem = (size_t)8 * K * sizeof(float);
  static SmemAttrOnce attr_once;
  { const int arc = set_max_smem_once((const void*)heads2_dgrad_kernel, 8 * 2048 * 4, &attr_once); if (arc) return arc; }
  const int ppb = threads / tpp;
  int blocks = 2 * num_sms();  // 128 registers: two 256-thread blocks per SM
  { const char* e = ab_env("DBX_HEADS2_BLOCKS"); if (e && atoi(e) > 0) blocks = atoi(e) * num_sms(); }
  if ((size_t)blocks * ppb * 8 > pixels) blocks = (int)((pixels + 8 * ppb - 1) / (8 * ppb));
  heads2_dgrad_kernel<<<blocks, threads, smem, st>>>(p);
  return (int)cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ weights
// Generic strided fp32 [co,ci,R,S] -> bf16 K-major  dstK[(rowK+o)*ldK + kK + (r*S+s)*cin_pad + i]
//                                 and bf16 dgrad    dstD[(rowD+i)*ldD + kD + ((R-1-r)*S+(S-1-s))*cout_pad + o]
__global__ void pack_weights_kernel(const float* __restrict__ src, int co, int ci, int R, int S, long s_co, long s_ci,
                                    long s_r, long s_s, bf16* dstK, long ldK, long rowK, long kK, int cin_pad,
                                    bf16* dstD, long ldD, long rowD, long kD, int cout_pad, float* dstF) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)co * ci * R * S;
  if (idx >= total) return;
  const int i = (int)(idx % ci);
  const int s = (int)((idx / ci) % S);
  const int r = (int)((idx / ((size_t)ci * S)) % R);
  const int o = (int)(idx / ((size_t)ci * S * R));
  const float v = src[o * s_co + i * s_ci + r * s_r + s * s_s];
  const bf16 b = __float2bfloat16_rn(v);
  const size_t k_idx = (size_t)(rowK + o) * ldK + kK + (size_t)(r * S + s) * cin_pad + i;
  if (dstK) dstK[k_idx] = b;
  if (dstF) dstF[k_idx] = v;  // fp32 master in the same K-major layout
  if (dstD) dstD[(size_t)(rowD + i) * ldD + kD + (size_t)((R - 1 - r) * S + (S - 1 - s)) * cout_pad + o] = b;
}

int pack_weights(const float* src, int co, int ci, int R, int S, long s_co, long s_ci, long s_r, long s_s, void* dstK,
                 long ldK, long rowK, long kK, int cin_pad, void* dstD, long ldD, long rowD, long kD, int cout_pad,
                 float* dstF, cudaStream_t st) {
  if (!src) return DBX_ERR_ARG;
  const size_t total = (size_t)co * ci * R * S;
  pack_weights_kernel<<<grid_for(total, 256), 256, 0, st>>>(src, co, ci, R, S, s_co, s_ci, s_r, s_s, (bf16*)dstK, ldK,
                                                            rowK, kK, cin_pad, (bf16*)dstD, ldD, rowD, kD, cout_pad,
                                                            dstF);
  return (int)cudaGetLastError();
}

// Inverse of the K-major packing for fp32 tensors (gradients / masters back to the torch [co,ci,R,S] layout).
__global__ void unpack_weights_kernel(const float* __restrict__ srcK, long ldK, long rowK, long kK, int cin_pad,
                                      float* __restrict__ dst, int co, int ci, int R, int S, long s_co, long s_ci,
                                      long s_r, long s_s) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)co * ci * R * S;
  if (idx >= total) return;
  const int i = (int)(idx % ci);
  const int s = (int)((idx / ci) % S);
  const int r = (int)((idx / ((size_t)ci * S)) % R);
  const int o = (int)(idx / ((size_t)ci * S * R));
  dst[o * s_co + i * s_ci + r * s_r + s * s_s] = srcK[(size_t)(rowK + o) * ldK + kK + (size_t)(r * S + s) * cin_pad + i];
}

int unpack_weights(const float* srcK, long ldK, long rowK, long kK, int cin_pad, float* dst, int co, int ci, int R,
                   int S, long s_co, long s_ci, long s_r, long s_s, cudaStream_t st) {
  if (!srcK || !dst) return DBX_ERR_ARG;
  const size_t total = (size_t)co * ci * R * S;
  unpack_weights_kernel<<<grid_for(total, 256), 256, 0, st>>>(srcK, ldK, rowK, kK, cin_pad, dst, co, ci, R, S, s_co,
                                                              s_ci, s_r, s_s);
  return (int)cudaGetLastError();
}

// All parameter tensors of a network in one launch (see ParamXfer): block -> entry by the running element count.
__global__ void __launch_bounds__(256) params_xfer_kernel(const ParamXfer t, int mode, float* __restrict__ flat32,
                                                          bf16* __restrict__ flat16) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= t.total) return;
  int ei = 0;
#pragma unroll 1
  while (ei + 1 < t.n && idx >= t.e[ei + 1].elem0) ++ei;
  const ParamXfer::E& e = t.e[ei];
  long long l = idx - e.elem0;
  const long long nw = (long long)e.co * e.ci * e.R * e.S;
  if (l >= nw) {  // bias
    const int o = (int)(l - nw);
    if (o >= e.co) return;
    if (mode == 0) {
      const float v = e.b[o * e.s_b];
      flat32[e.b_off + o] = v;
      if (e.dup) flat32[e.b_off + 64 + o] = v;
    } else {
      e.b[o * e.s_b] = flat32[e.b_off + o];
    }
    return;
  }
  const int i = (int)(l % e.ci);
  const int s = (int)((l / e.ci) % e.S);
  const int r = (int)((l / ((long long)e.ci * e.S)) % e.R);
  const int o = (int)(l / ((long long)e.ci * e.S * e.R));
  const long long k = e.w_off + (long long)(e.rowK + o) * e.ld + e.kK + (long long)(r * e.S + s) * e.cin_pad + i;
  float* tp = e.w + o * e.s_co + i * e.s_ci + r * e.s_r + s * e.s_s;
  if (mode == 0) {
    const float v = *tp;
    flat32[k] = v;
    flat16[k] = __float2bfloat16_rn(v);
    if (e.dup) {
      const long long k2 = k + 64LL * e.ld + 32;
      flat32[k2] = v;
      flat16[k2] = __float2bfloat16_rn(v);
    }
  } else {
    *tp = flat32[k];
  }
}

int params_xfer(const ParamXfer& t, int mode, float* flat32, void* flat16, cudaStream_t st) {
  if (t.n < 1 || t.n > ParamXfer::kMax || !flat32 || (mode == 0 && !flat16) || mode < 0 || mode > 1) return DBX_ERR_ARG;
  params_xfer_kernel<<<grid_for((size_t)t.total, 256), 256, 0, st>>>(t, mode, flat32, (bf16*)flat16);
  return (int)cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ head maps <-> NCHW
// One thread per pixel: reads its HC interleaved floats, writes every channel plane coalesced across the warp.
__global__ void __launch_bounds__(256) heads_to_nchw_kernel(const float* __restrict__ head, int HC,
                                                            const float* __restrict__ rf, int RC, int N, int HW,
                                                            float* __restrict__ score, float* __restrict__ loc,
                                                            float* __restrict__ lm, float* __restrict__ lmloc,
                                                            float* __restrict__ rfo) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)N * HW) return;
  const size_t n = idx / HW, i = idx - n * HW;
  const float* h = head + idx * HC;
  if (score) score[idx] = h[0];
  if (loc) for (int c = 0; c < 4; ++c) loc[(n * 4 + c) * HW + i] = h[1 + c];
  if (lm) for (int c = 0; c < 4; ++c) lm[(n * 4 + c) * HW + i] = h[5 + c];
  if (lmloc) for (int c = 0; c < 8; ++c) lmloc[(n * 8 + c) * HW + i] = h[9 + c];
  if (rfo && rf) rfo[idx] = rf[idx * RC];
}

int heads_to_nchw(const float* head, int HC, const float* rf, int RC, int N, int HW, float* score, float* loc,
                  float* lm, float* lmloc, float* rfo, cudaStream_t st) {
  if (!head || N <= 0 || HW <= 0 || HC < 5 || (lm && HC < 9) || (lmloc && HC < 17) || (rfo && !rf)) return DBX_ERR_ARG;
  heads_to_nchw_kernel<<<grid_for((size_t)N * HW, 256), 256, 0, st>>>(head, HC, rf, RC, N, HW, score, loc, lm, lmloc,
                                                                      rfo);
  return (int)cudaGetLastError();
}

// The gradients autograd hands to the module's backward (fp32 NCHW, any of them absent) -> the engine's bf16
// d_head [pixels][64] (channel map of head_out, zero beyond) and d_rf [pixels][64] (channel 0).
__global__ void __launch_bounds__(256) nchw_to_head_grads_kernel(const float* __restrict__ gs, const float* __restrict__ gl,
                                                                 const float* __restrict__ gm, const float* __restrict__ gml,
                                                                 const float* __restrict__ gr, int N, int HW,
                                                                 bf16* __restrict__ d_head, bf16* __restrict__ d_rf) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)N * HW) return;
  const size_t n = idx / HW, i = idx - n * HW;
  float f[24];
#pragma unroll
  for (int c = 0; c < 24; ++c) f[c] = 0.f;
  if (gs) f[0] = gs[idx];
  if (gl) for (int c = 0; c < 4; ++c) f[1 + c] = gl[(n * 4 + c) * HW + i];
  if (gm) for (int c = 0; c < 4; ++c) f[5 + c] = gm[(n * 4 + c) * HW + i];
  if (gml) for (int c = 0; c < 8; ++c) f[9 + c] = gml[(n * 8 + c) * HW + i];
  uint4* o = reinterpret_cast<uint4*>(d_head + idx * 64);
  o[0] = pack8(f); o[1] = pack8(f + 8); o[2] = pack8(f + 16);
#pragma unroll
  for (int q = 3; q < 8; ++q) o[q] = make_uint4(0u, 0u, 0u, 0u);
  if (d_rf) {
    uint4* r = reinterpret_cast<uint4*>(d_rf + idx * 64);
    r[0] = make_uint4(pack_bf16x2(gr ? gr[idx] : 0.f, 0.f), 0u, 0u, 0u);
#pragma unroll
    for (int q = 1; q < 8; ++q) r[q] = make_uint4(0u, 0u, 0u, 0u);
  }
}

int nchw_to_head_grads(const float* g_score, const float* g_loc, const float* g_lm, const float* g_lmloc,
                       const float* g_rf, int N, int HW, void* d_head, void* d_rf, cudaStream_t st) {
  if (!d_head || N <= 0 || HW <= 0) return DBX_ERR_ARG;
  nchw_to_head_grads_kernel<<<grid_for((size_t)N * HW, 256), 256, 0, st>>>(g_score, g_loc, g_lm, g_lmloc, g_rf, N, HW,
                                                                           (bf16*)d_head, (bf16*)d_rf);
  return (int)cudaGetLastError();
}

// bf16 K-major [rows][T][cin_pad] -> bf16 dgrad layout [cin_pad][T flipped][kpad >= rows] (32x32 smem tile
// transpose; columns rows..kpad-1 stay zero from initialisation).
__global__ void transpose_dgrad_kernel(const bf16* __restrict__ wk, bf16* __restrict__ wd, int rows, int T,
                                       int cin_pad, int kpad) {
  __shared__ bf16 tile[32][34];
  const int t = blockIdx.z;
  const int o0 = blockIdx.y * 32, i0 = blockIdx.x * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
  for (int k = ty; k < 32; k += 8) {
    const int o = o0 + k, i = i0 + tx;
    tile[k][tx] = (o < rows && i < cin_pad) ? wk[((size_t)o * T + t) * cin_pad + i] : __float2bfloat16_rn(0.f);
  }
  __syncthreads();
  const int tf = T - 1 - t;  // (R-1-r)*S + (S-1-s) == T-1-(r*S+s)
  for (int k = ty; k < 32; k += 8) {
    const int i = i0 + k, o = o0 + tx;
    if (i < cin_pad && o < rows) wd[((size_t)i * T + tf) * kpad + o] = tile[tx][k];
  }
}

// All groups of a network in ONE launch: block -> (group, 32 x 32 tile, tap) through a small table passed by value.
__global__ void transpose_dgrad_multi_kernel(const bf16* __restrict__ wk_base, bf16* __restrict__ wd_base,
                                             const TransposeBatch tb) {
  __shared__ bf16 tile[32][34];
  int gi = 0;
#pragma unroll 1
  while (gi + 1 < tb.n && (int)blockIdx.x >= tb.g[gi + 1].block0) ++gi;
  const TransposeBatch::G g = tb.g[gi];
  int local = (int)blockIdx.x - g.block0;
  const int tiles_i = (g.cin_pad + 31) / 32, tiles_o = (g.rows + 31) / 32;
  const int t = local / (tiles_i * tiles_o);
  local -= t * tiles_i * tiles_o;
  const int o0 = (local / tiles_i) * 32, i0 = (local % tiles_i) * 32;
  const bf16* wk = wk_base + g.wk_off;
  bf16* wd = wd_base + g.wd_off;
  const int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
  for (int k = ty; k < 32; k += 8) {
    const int o = o0 + k, i = i0 + tx;
    tile[k][tx] = (o < g.rows && i < g.cin_pad) ? wk[((size_t)o * g.T + t) * g.cin_pad + i] : __float2bfloat16_rn(0.f);
  }
  __syncthreads();
  const int tf = g.T - 1 - t;
  for (int k = ty; k < 32; k += 8) {
    const int i = i0 + k, o = o0 + tx;
    if (i < g.cin_pad && o < g.rows) wd[((size_t)i * g.T + tf) * g.kpad + o] = tile[tx][k];
  }
}

int transpose_dgrad_multi(const void* wk_base, void* wd_base, const TransposeBatch& tb, int total_blocks,
                          cudaStream_t st) {
  if (!wk_base || !wd_base || tb.n < 1 || tb.n > TransposeBatch::kMax || total_blocks < 1) return DBX_ERR_ARG;
  transpose_dgrad_multi_kernel<<<total_blocks, dim3(32, 8), 0, st>>>((const bf16*)wk_base, (bf16*)wd_base, tb);
  return (int)cudaGetLastError();
}

int transpose_dgrad(const void* wk, void* wd, int rows, int T, int cin_pad, int kpad, cudaStream_t st) {
  if (!wk || !wd) return DBX_ERR_ARG;
  dim3 grid((cin_pad + 31) / 32, (rows + 31) / 32, T), block(32, 8);
  transpose_dgrad_kernel<<<grid, block, 0, st>>>((const bf16*)wk, (bf16*)wd, rows, T, cin_pad, kpad);
  return (int)cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ SGD
// torch.optim.SGD(momentum, weight_decay, dampening=0, nesterov=False) (DenseBox.py:2821-2824, :2926) on the flat
// fp32 master buffer; also refreshes the bf16 copy the tensor cores read and clears the gradient for the next step.
__global__ void sgd_step_kernel(float* __restrict__ w, float* __restrict__ g, float* __restrict__ v,
                                bf16* __restrict__ wb, size_t n, float lr, float momentum, float wd, int first,
                                int zero_grad) {
  const size_t i4 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i4 >= n) return;
  float4 W = *reinterpret_cast<float4*>(w + i4), G = *reinterpret_cast<float4*>(g + i4);
  float4 V = first ? make_float4(0.f, 0.f, 0.f, 0.f) : *reinterpret_cast<float4*>(v + i4);
  float wv[4] = {W.x, W.y, W.z, W.w}, gv[4] = {G.x, G.y, G.z, G.w}, vv[4] = {V.x, V.y, V.z, V.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float d = gv[j] + wd * wv[j];
    vv[j] = first ? d : momentum * vv[j] + d;
    wv[j] = wv[j] - lr * vv[j];
  }
  *reinterpret_cast<float4*>(w + i4) = make_float4(wv[0], wv[1], wv[2], wv[3]);
  *reinterpret_cast<float4*>(v + i4) = make_float4(vv[0], vv[1], vv[2], vv[3]);
  if (zero_grad) *reinterpret_cast<float4*>(g + i4) = make_float4(0.f, 0.f, 0.f, 0.f);
  if (wb) *reinterpret_cast<uint2*>(wb + i4) = make_uint2(pack_bf16x2(wv[0], wv[1]), pack_bf16x2(wv[2], wv[3]));
}

int sgd_step(float* w, float* g, float* v, void* wb, size_t n, float lr, float momentum, float wd, int first,
             int zero_grad, cudaStream_t st) {
  if (!w || !g || !v || n % 4) return DBX_ERR_ARG;
  sgd_step_kernel<<<grid_for(n / 4, 256), 256, 0, st>>>(w, g, v, (bf16*)wb, n, lr, momentum, wd, first, zero_grad);
  return (int)cudaGetLastError();
}

__global__ void cast_bf16_kernel(const float* __restrict__ src, bf16* __restrict__ dst, size_t n) {
  const size_t i4 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i4 >= n) return;
  const float4 f = *reinterpret_cast<const float4*>(src + i4);
  *reinterpret_cast<uint2*>(dst + i4) = make_uint2(pack_bf16x2(f.x, f.y), pack_bf16x2(f.z, f.w));
}

int cast_bf16(const float* src, void* dst, size_t n, cudaStream_t st) {
  if (!src || !dst || n % 4) return DBX_ERR_ARG;
  cast_bf16_kernel<<<grid_for(n / 4, 256), 256, 0, st>>>(src, (bf16*)dst, n);
  return (int)cudaGetLastError();
}

// Zero the off-diagonal blocks of the block-diagonal conv5_2 gradient: row o belongs to head `head_of_row[o]`,
// whose K range is [head*512, head*512+512).
struct HeadRows { int nh; int start[5]; };  // head h owns rows [start[h], start[h+1])
__global__ void blockdiag_mask_kernel(float* __restrict__ g, int rows, int ld, HeadRows hr) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)rows * ld) return;
  const int o = (int)(idx / ld), k = (int)(idx % ld);
  int h = -1;
  for (int j = 0; j < hr.nh; ++j) if (o >= hr.start[j] && o < hr.start[j + 1]) h = j;
  if (h < 0 || k / 512 != h) g[idx] = 0.f;
}

int blockdiag_mask(float* g, int rows, int ld, int nh, const int* start, cudaStream_t st) {
  HeadRows hr; hr.nh = nh;
  for (int j = 0; j < 5; ++j) hr.start[j] = j <= nh ? start[j] : 0;
  blockdiag_mask_kernel<<<grid_for((size_t)rows * ld, 256), 256, 0, st>>>(g, rows, ld, hr);
  return (int)cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ dropout mask
// nn.Dropout(p=0.5) in train mode (DenseBox.py:160,176) as an explicit bf16 {0,2} tensor — the same bits the conv
// epilogue draws in place (dropout_bits16); used for parity tests and by callers that want to inspect the mask.
__global__ void dropout_mask_kernel(bf16* __restrict__ mask, size_t n16, unsigned long long seed,
                                    unsigned long long offset) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n16) return;
  const uint32_t bits = dropout_bits16((unsigned long long)idx * 16ull, seed, offset);
  const uint32_t two = 0x4000u;  // bf16(2.0)
  uint32_t o[8];
#pragma unroll
  for (int j = 0; j < 8; ++j)
    o[j] = (((bits >> (2 * j)) & 1u) ? two : 0u) | (((bits >> (2 * j + 1)) & 1u) ? (two << 16) : 0u);
  uint4* dst = reinterpret_cast<uint4*>(mask) + idx * 2;
  dst[0] = make_uint4(o[0], o[1], o[2], o[3]);
  dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
}

int dropout_mask(void* mask, size_t n, unsigned long long seed, unsigned long long offset, cudaStream_t st) {
  if (!mask || n % 16) return DBX_ERR_ARG;
  dropout_mask_kernel<<<grid_for(n / 16, 256), 256, 0, st>>>((bf16*)mask, n / 16, seed, offset);
  return (int)cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ refine glue
// fusion_2 = cat(landmarks, scores) -> MaxPool2d(2) (DenseBox.py:464-465). head: fp32 [N,H,W,HC] with ch0 = score,
// ch5..8 = landmark heat-maps.  pooled: bf16 [N,H/2,W/2,64], channels 0..3 = landmarks, 4 = score, 5..63 zero.
__global__ void refine_pool_pack_kernel(const float* __restrict__ head, int HC, bf16* __restrict__ pooled, int N,
                                        int H, int W) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int PH = H / 2, PW = W / 2;
  if (idx >= (size_t)N * PH * PW) return;
  const int px = (int)(idx % PW), py = (int)((idx / PW) % PH), n = (int)(idx / ((size_t)PW * PH));
  float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const int src_ch[5] = {5, 6, 7, 8, 0};
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    float m = -INFINITY;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int y = 2 * py + (q >> 1), x = 2 * px + (q & 1);
      m = fmaxf(m, __ldg(head + (((size_t)n * H + y) * W + x) * HC + src_ch[k]));
    }
    f[k] = m;
  }
  *reinterpret_cast<uint4*>(pooled + idx * 64) = pack8(f);
}

int refine_pool_pack(const float* head, int HC, void* pooled, int N, int H, int W, cudaStream_t st) {
  if (!head || !pooled || H % 2 || W % 2 || HC < 9) return DBX_ERR_ARG;
  refine_pool_pack_kernel<<<grid_for((size_t)N * (H / 2) * (W / 2), 256), 256, 0, st>>>(head, HC, (bf16*)pooled, N, H,
                                                                                       W);
  return (int)cudaGetLastError();
}

// Backward of the above: route dpooled[:, 0..4] to the first arg-max of each 2x2 window and ADD it into the head
// gradient (bf16 [N,H,W,64], channels as in `head`) that the loss kernel has already written.
__global__ void refine_pool_bwd_kernel(const float* __restrict__ head, int HC, const bf16* __restrict__ dpooled,
                                       bf16* __restrict__ dhead, int N, int H, int W) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int PH = H / 2, PW = W / 2;
  if (idx >= (size_t)N * PH * PW) return;
  const int px = (int)(idx % PW), py = (int)((idx / PW) % PH), n = (int)(idx / ((size_t)PW * PH));
  float g[8];
  unpack8(ldg16(dpooled + idx * 64), g);
  const int src_ch[5] = {5, 6, 7, 8, 0};
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    int arg = 0; float m = -INFINITY;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int y = 2 * py + (q >> 1), x = 2 * px + (q & 1);
      const float v = __ldg(head + (((size_t)n * H + y) * W + x) * HC + src_ch[k]);
      if (v > m) { m = v; arg = q; }
    }
    const int y = 2 * py + (arg >> 1), x = 2 * px + (arg & 1);
    bf16* p = dhead + (((size_t)n * H + y) * W + x) * 64 + src_ch[k];
    *p = __float2bfloat16_rn(__bfloat162float(*p) + g[k]);
  }
}

int refine_pool_bwd(const float* head, int HC, const void* dpooled, void* dhead, int N, int H, int W, cudaStream_t st) {
  if (!head || !dpooled || !dhead || H % 2 || W % 2 || HC < 9) return DBX_ERR_ARG;
  refine_pool_bwd_kernel<<<grid_for((size_t)N * (H / 2) * (W / 2), 256), 256, 0, st>>>(
      head, HC, (const bf16*)dpooled, (bf16*)dhead, N, H, W);
  return (int)cudaGetLastError();
}

}  // namespace dbx
