// densebox_b200 — detection post-processing, one CTA per image: top-K of the raw score map, box / landmark decode
// and greedy NMS (reference: parse_out_MN / parse_DetLMLOC DenseBox.py:3114-3217, parse_DetLM :3220-3300 — landmarks =
// the arg-max of each landmark heat-map, the same four points for every detection of the image —, NMS :3398-3443).
// The reference does this on the CPU after five D2H copies per image; here only K rows leave the GPU.
#include "dbx_common.h"
#include "dbx_ptx.cuh"

namespace dbx {

static constexpr int PT = 1024, KMAX = 64;

struct MapView { const float* p; long img, pix, ch; };  // element strides: image, pixel, channel
__device__ __forceinline__ float mv(const MapView& m, int n, int idx, int c) {
  return __ldg(m.p + (size_t)n * m.img + (size_t)idx * m.pix + (size_t)c * m.ch);
}

// lm_mode: 0 = boxes only, 1 = `lm` holds 8 landmark offsets per pixel (parse_DetLMLOC), 2 = `lm` holds 4 landmark
// heat-maps whose arg-max (lowest index among equals) gives the landmark (parse_DetLM :3283-3292).
__global__ void __launch_bounds__(PT) decode_nms_kernel(MapView score, MapView loc, MapView lmloc, int lm_mode, int HW,
                                                        int W4, int K, double thresh, float* __restrict__ dets,
                                                        int* __restrict__ keep) {
  __shared__ float wv[PT / 32];
  __shared__ int wi[PT / 32];
  __shared__ int sel[KMAX];
  __shared__ float selv[KMAX];
  __shared__ double box[KMAX][4];
  __shared__ int alive[KMAX];
  __shared__ int lmarg[4];
  const int n = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int KS = K;  // row stride of the outputs
  const int has_lm = lm_mode == 1;
  if (K > KMAX) K = KMAX;
  if (K > HW) K = HW;
  if (lm_mode == 2) {  // per-landmark arg-max of the heat-maps (torch.topk(k=1) of :3285)
    for (int k = 0; k < 4; ++k) {
      float bv = -INFINITY; int bi = 0x7fffffff;
      for (int i = tid; i < HW; i += PT) {
        const float v = mv(lmloc, n, i, k);
        if (v > bv || bi == 0x7fffffff) { bv = v; bi = i; }  // ascending i: the first maximum stays
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (oi != 0x7fffffff && (bi == 0x7fffffff || ov > bv || (ov == bv && oi < bi))) { bv = ov; bi = oi; }
      }
      if (lane == 0) { wv[warp] = bv; wi[warp] = bi; }
      __syncthreads();
      if (tid == 0) {
        float v = wv[0]; int i = wi[0];
        for (int w = 1; w < PT / 32; ++w)
          if (wi[w] != 0x7fffffff && (i == 0x7fffffff || wv[w] > v || (wv[w] == v && wi[w] < i))) { v = wv[w]; i = wi[w]; }
        lmarg[k] = i;
      }
      __syncthreads();
    }
  }
  // ---- top-K by K rounds of block arg-max (ties: lowest index); torch.topk(sorted) order = descending score.
  // Thread t owns the pixels t, t + PT, ...; it caches the best of its not-yet-selected pixels, and after every round
  // only the thread that owned the winner rescans (a round is one block reduction + 64 loads of one thread instead of
  // a pass of the whole block over the map: 0.98 -> ms at 1024 x 1024 x 16, see profiles/README.md).
  float bv = -INFINITY; int bi = 0x7fffffff;
  auto rescan = [&](int nsel) {
    bv = -INFINITY; bi = 0x7fffffff;
#pragma unroll 4
    for (int i = tid; i < HW; i += PT) {
      bool taken = false;
      for (int j = 0; j < nsel; ++j) taken |= (sel[j] == i);
      if (taken) continue;
      const float v = mv(score, n, i, 0);
      if (v > bv || (v == bv && i < bi) || bi == 0x7fffffff) { bv = v; bi = i; }
    }
  };
  rescan(0);
  for (int k = 0; k < K; ++k) {
    float rv = bv; int ri = bi;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, rv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, ri, o);
      if (oi != 0x7fffffff && (ri == 0x7fffffff || ov > rv || (ov == rv && oi < ri))) { rv = ov; ri = oi; }
    }
    if (lane == 0) { wv[warp] = rv; wi[warp] = ri; }
    __syncthreads();
    if (tid == 0) {
      float v = wv[0]; int i = wi[0];
      for (int w = 1; w < PT / 32; ++w)
        if (wi[w] != 0x7fffffff && (i == 0x7fffffff || wv[w] > v || (wv[w] == v && wi[w] < i))) { v = wv[w]; i = wi[w]; }
      sel[k] = i; selv[k] = v;
    }
    __syncthreads();
    if (k + 1 < K && sel[k] % PT == tid) rescan(k + 1);
  }
  // ---- decode (:3162-3213): float32 `xi - map[idx]`, then * 4.0
  if (tid < K) {
    const int idx = sel[tid], xi = idx % W4, yi = idx / W4;
    float* d = dets + ((size_t)n * KS + tid) * 13;
    float b[4];
    for (int c = 0; c < 4; ++c) b[c] = __fsub_rn((float)((c & 1) ? yi : xi), mv(loc, n, idx, c)) * 4.0f;
    d[0] = b[0]; d[1] = b[1]; d[2] = b[2]; d[3] = b[3]; d[4] = selv[tid];
    for (int c = 0; c < 8; ++c) {
      float v = 0.f;
      if (has_lm) v = __fsub_rn((float)((c & 1) ? yi : xi), mv(lmloc, n, idx, c)) * 4.0f;
      else if (lm_mode == 2) v = (float)((c & 1) ? lmarg[c >> 1] / W4 : lmarg[c >> 1] % W4) * 4.0f;
      d[5 + c] = v;
    }
    for (int c = 0; c < 4; ++c) box[tid][c] = (double)b[c];
    alive[tid] = 1;
  }
  __syncthreads();
  // ---- greedy NMS (:3406-3441): rows are already in descending score order
  if (tid == 0) {
    for (int i = 0; i < K; ++i) {
      keep[(size_t)n * KS + i] = alive[i];
      if (!alive[i]) continue;
      const double ai = (box[i][2] - box[i][0] + 1) * (box[i][3] - box[i][1] + 1);
      for (int j = i + 1; j < K; ++j) {
        if (!alive[j]) continue;
        const double xx1 = fmax(box[i][0], box[j][0]), yy1 = fmax(box[i][1], box[j][1]);
        const double xx2 = fmin(box[i][2], box[j][2]), yy2 = fmin(box[i][3], box[j][3]);
        const double w = fmax(0.0, xx2 - xx1 + 1), h = fmax(0.0, yy2 - yy1 + 1);
        const double inter = w * h;
        const double aj = (box[j][2] - box[j][0] + 1) * (box[j][3] - box[j][1] + 1);
        const double ovr = inter / (ai + aj - inter);
        if (!(ovr <= thresh)) alive[j] = 0;
      }
    }
  }
}

int decode_nms(const float* score, long s_img, long s_pix, const float* loc, long l_img, long l_pix, long l_ch,
               const float* lmloc, long m_img, long m_pix, long m_ch, int N, int H4, int W4, int K, double thresh,
               float* dets, int* keep, cudaStream_t st, int lm_heat) {
  if (!score || !loc || !dets || !keep || N <= 0 || K <= 0 || K > KMAX) return DBX_ERR_ARG;
  if (lm_heat && !lmloc) return DBX_ERR_ARG;
  MapView s{score, s_img, s_pix, 0}, l{loc, l_img, l_pix, l_ch}, m{lmloc ? lmloc : loc, m_img, m_pix, m_ch};
  decode_nms_kernel<<<N, PT, 0, st>>>(s, l, m, lmloc ? (lm_heat ? 2 : 1) : 0, H4 * W4, W4, K, thresh, dets, keep);
  return (int)cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ perspective warp
// perspective_transform (DenseBox.py:3446-3481) = cv2.getPerspectiveTransform (host, 8 x 8 solve) + cv2.warpPerspective
// (INTER_LINEAR, constant border 0).  The kernel restates OpenCV's fixed-point bilinear remap so that the result is
// bit-identical: source coordinates in double ((M0 x + (M1 y + M2)) * 32 / w, round-half-even to 1/32 pixel), the four
// weights as 15-bit integers ((32 - a)(32 - b) * 32 etc. — exact, they always sum to 32768), (sum + 2^14) >> 15.
struct WarpMat { double m[9]; };  // INVERSE map: destination pixel -> source coordinates

__global__ void __launch_bounds__(256) warp_perspective_u8_kernel(const unsigned char* __restrict__ src, int H, int W,
                                                                  int C, WarpMat M, unsigned char* __restrict__ dst,
                                                                  int dH, int dW) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= dW || y >= dH) return;
  const double X0 = __dadd_rn(__dmul_rn(M.m[1], (double)y), M.m[2]);
  const double Y0 = __dadd_rn(__dmul_rn(M.m[4], (double)y), M.m[5]);
  const double W0 = __dadd_rn(__dmul_rn(M.m[7], (double)y), M.m[8]);
  double w = __dadd_rn(W0, __dmul_rn(M.m[6], (double)x));
  w = w != 0.0 ? __ddiv_rn(32.0, w) : 0.0;
  double fX = __dmul_rn(__dadd_rn(X0, __dmul_rn(M.m[0], (double)x)), w);
  double fY = __dmul_rn(__dadd_rn(Y0, __dmul_rn(M.m[3], (double)x)), w);
  fX = fmax(-2147483648.0, fmin(2147483647.0, fX));
  fY = fmax(-2147483648.0, fmin(2147483647.0, fY));
  const int Xi = __double2int_rn(fX), Yi = __double2int_rn(fY);
  const int sx = Xi >> 5, sy = Yi >> 5, a = Xi & 31, b = Yi & 31;
  const int w00 = (32 - a) * (32 - b) * 32, w01 = a * (32 - b) * 32, w10 = (32 - a) * b * 32, w11 = a * b * 32;
  const bool x0 = sx >= 0 && sx < W, x1 = sx + 1 >= 0 && sx + 1 < W, y0 = sy >= 0 && sy < H, y1 = sy + 1 >= 0 && sy + 1 < H;
  unsigned char* o = dst + ((size_t)y * dW + x) * C;
  for (int c = 0; c < C; ++c) {
    int acc = 0;
    if (y0 && x0) acc += w00 * (int)src[((size_t)sy * W + sx) * C + c];
    if (y0 && x1) acc += w01 * (int)src[((size_t)sy * W + sx + 1) * C + c];
    if (y1 && x0) acc += w10 * (int)src[((size_t)(sy + 1) * W + sx) * C + c];
    if (y1 && x1) acc += w11 * (int)src[((size_t)(sy + 1) * W + sx + 1) * C + c];
    const int v = (acc + (1 << 14)) >> 15;
    o[c] = (unsigned char)(v < 0 ? 0 : (v > 255 ? 255 : v));
  }
}

int warp_perspective_u8(const unsigned char* src, int H, int W, int C, const double* minv, unsigned char* dst, int dH,
                        int dW, cudaStream_t st) {
  if (!src || !dst || !minv || H <= 0 || W <= 0 || C <= 0 || C > 4 || dH <= 0 || dW <= 0) return DBX_ERR_ARG;
  WarpMat M;
  for (int i = 0; i < 9; ++i) M.m[i] = minv[i];
  dim3 grid((dW + 255) / 256, dH);
  warp_perspective_u8_kernel<<<grid, 256, 0, st>>>(src, H, W, C, M, dst, dH, dW);
  return (int)cudaGetLastError();
}

}  // namespace dbx
