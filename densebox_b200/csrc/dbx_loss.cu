// densebox_b200 — the fused multi-task loss of the DenseBox training loops, one CTA per sample.
//
// Everything the reference does between net.forward() and loss.backward() (DenseBox.py:2843-2918; LM :2575-2723;
// LMLOC :2300-2456; pos/neg :2023-2180) happens here without leaving the SM:
//   ground-truth synthesis from bbox / vertices (init_score_map :1572-1582, init_loc_map :1643-1653,
//   init_lm_heatmap :1815-1823, init_lm_locmap :1705-1718), squared errors (nn.MSELoss(reduce=False) :2818),
//   per-sample top-k hard-negative mining (torch.topk :2882, radix select here), injected random negatives
//   (np.random.choice :2888-2893), mask_by_sel :1368-1400, the gray zones (mask_gray_zone_cls :1486-1504,
//   mask_gray_zone_lm :1453-1462), the masked sums (:2914-2918) and d(loss)/d(prediction).
// The reference crosses the PCIe bus 7+ times and runs 3600*B Python iterations for the same work.
//
// Geometry arithmetic reproduces NumPy>=2 scalar semantics of the reference (float32 products of `ratio * w`,
// Python-double sums, int() truncation, Python slice clipping) — see oracle/densebox_oracle.py score_box/gray_box.
#include "dbx_common.h"
#include "dbx_ptx.cuh"

namespace dbx {

static constexpr int MAPW = 60, MAPN = 3600, LT = 1024;  // one CTA of 32 warps per sample: the kernel is latency-bound


struct Box { int x0, x1, y0, y1; };  // python slice bounds, already clipped: [x0,x1) x [y0,y1)

__device__ __forceinline__ void pyslice(int a, int b, int& lo, int& hi) {
  lo = a < 0 ? max(a + MAPW, 0) : min(a, MAPW);
  hi = b < 0 ? max(b + MAPW, 0) : min(b, MAPW);
  if (hi < lo) hi = lo;
}
__device__ __forceinline__ Box clip(int x0, int x1, int y0, int y1) {
  Box b; pyslice(x0, x1, b.x0, b.x1); pyslice(y0, y1, b.y0, b.y1); return b;
}
__device__ __forceinline__ bool inbox(const Box& b, int x, int y) { return x >= b.x0 && x < b.x1 && y >= b.y0 && y < b.y1; }

// init_score_map :1572-1582
__device__ Box score_box(const float* c) {
  const double cx = (double)__fadd_rn(c[0], c[2]) * 0.5, cy = (double)__fadd_rn(c[1], c[3]) * 0.5;
  const float rw = __fmul_rn(0.3f, __fsub_rn(c[2], c[0])), rh = __fmul_rn(0.3f, __fsub_rn(c[3], c[1]));
  const int ox = (int)(cx - (double)__fmul_rn(rw, 0.5f) + 0.5), oy = (int)(cy - (double)__fmul_rn(rh, 0.5f) + 0.5);
  const int ex = (int)((double)ox + (double)rw + 0.5), ey = (int)((double)oy + (double)rh + 0.5);
  return clip(ox, ex + 1, oy, ey + 1);
}
// mask_gray_zone_cls :1486-1504 -> outer (zeroed) and inner (re-set) boxes
__device__ void gray_boxes(const float* c, Box& outer, Box& inner) {
  const double cx = (double)__fadd_rn(c[0], c[2]) * 0.5, cy = (double)__fadd_rn(c[1], c[3]) * 0.5;
  const float rw = __fmul_rn(0.3f, __fsub_rn(c[2], c[0])), rh = __fmul_rn(0.3f, __fsub_rn(c[3], c[1]));
  const int gx = (int)(cx - (double)__fmul_rn(rw, 0.5f) - 2.0 + 0.5), gy = (int)(cy - (double)__fmul_rn(rh, 0.5f) - 2.0 + 0.5);
  const int Gx = (int)((double)gx + (double)rw + 4.0 + 0.5), Gy = (int)((double)gy + (double)rh + 4.0 + 0.5);
  outer = clip(gx, Gx, gy, Gy);
  inner = clip(gx + 2, Gx - 2 + 1, gy + 2, Gy - 2 + 1);
}

__device__ __forceinline__ double block_sum(double v, double* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double s = 0.0;
  for (int w = 0; w < LT / 32; ++w) s += red[w];
  return s;
}

// Exclusive prefix sum of one int per thread over the block (LT threads); `ws` = LT/32 ints of shared memory.
// Returns the exclusive prefix of this thread; *total = sum over the block.
__device__ __forceinline__ int block_exscan(int v, int* ws, int* total) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  __syncthreads();
  if (lane == 31) ws[w] = inc;
  __syncthreads();
  if (w == 0) {
    int x = lane < LT / 32 ? ws[lane] : 0, y = x;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, y, o);
      if (lane >= o) y += t;
    }
    if (lane < LT / 32) ws[lane] = y - x;  // exclusive warp offsets
    if (lane == 31) ws[LT / 32] = y;        // block total
  }
  __syncthreads();
  *total = ws[LT / 32];
  return ws[w] + inc - v;
}

// Mark the k largest keys (ties at the threshold: lowest index first) by setting sel[i] = 1. keys are the bit
// patterns of non-negative floats (order preserving).  All LT threads call this.  4-pass radix select on 8-bit
// digits: histogram with shared-memory atomics, then the digit of the k-th largest key is found by a block-wide
// suffix count (thread d < 256 owns digit d), no serial loops.
__device__ void topk_mark(const uint32_t* keys, int k, unsigned char* sel, uint32_t* hist, int* sc) {
  if (k <= 0) return;
  if (k > MAPN) k = MAPN;
  int* ws = sc + 4;  // LT/32 + 1 ints of scan workspace behind the two result slots
  uint32_t prefix = 0, pmask = 0;
  int remaining = k;
  for (int shift = 24; shift >= 0; shift -= 8) {
    for (int i = threadIdx.x; i < 256; i += LT) hist[i] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < MAPN; i += LT) {
      const uint32_t key = keys[i];
      if ((key & pmask) == prefix) atomicAdd(&hist[(key >> shift) & 255u], 1u);
    }
    __syncthreads();
    // thread t < 256 takes digit d = 255 - t: its exclusive prefix = number of keys with a LARGER digit
    const int mine = threadIdx.x < 256 ? (int)hist[255 - threadIdx.x] : 0;
    int total;
    const int above = block_exscan(mine, ws, &total);
    if (threadIdx.x < 256 && above < remaining && above + mine >= remaining) { sc[0] = 255 - (int)threadIdx.x; sc[1] = remaining - above; }
    if (threadIdx.x == 0 && total < remaining) { sc[0] = 0; sc[1] = remaining - (total - (int)hist[0]); }  // cannot happen (k <= MAPN)
    __syncthreads();
    prefix |= (uint32_t)sc[0] << shift;
    pmask |= 255u << shift;
    remaining = sc[1];
    __syncthreads();
  }
  // prefix == k-th largest key; `remaining` of the keys equal to it are taken, lowest indices first.
  const int per = (MAPN + LT - 1) / LT;
  const int b0 = threadIdx.x * per, b1 = min(b0 + per, MAPN);
  int cnt = 0;
  for (int i = b0; i < b1; ++i) cnt += keys[i] == prefix;
  int total;
  int rank = block_exscan(cnt, ws, &total);
  for (int i = b0; i < b1; ++i) {
    const uint32_t key = keys[i];
    if (key > prefix) sel[i] = 1;
    else if (key == prefix) { if (rank < remaining) sel[i] = 1; ++rank; }
  }
  __syncthreads();
}

// k = 1: the largest key, lowest index among equals (what topk_mark(keys, 1, ...) marks) — one arg-max reduction
// instead of a 4-pass radix select.  `ws` = 2 * LT/32 words of shared memory.  All LT threads call this.
__device__ void top1_mark(const uint32_t* keys, unsigned char* sel, uint32_t* ws) {
  uint32_t bk = 0u; int bi = MAPN;  // (key, index): larger key wins, then smaller index
  for (int i = threadIdx.x; i < MAPN; i += LT) {
    const uint32_t key = keys[i];
    if (key > bk || (key == bk && i < bi)) { bk = key; bi = i; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const uint32_t ok = __shfl_xor_sync(0xffffffffu, bk, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ok > bk || (ok == bk && oi < bi)) { bk = ok; bi = oi; }
  }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) { ws[2 * w] = bk; ws[2 * w + 1] = (uint32_t)bi; }
  __syncthreads();
  if (w == 0) {
    bk = lane < LT / 32 ? ws[2 * lane] : 0u;
    bi = lane < LT / 32 ? (int)ws[2 * lane + 1] : MAPN;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const uint32_t ok = __shfl_xor_sync(0xffffffffu, bk, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ok > bk || (ok == bk && oi < bi)) { bk = ok; bi = oi; }
    }
    if (lane == 0 && bi < MAPN) sel[bi] = 1;
  }
  __syncthreads();
}

__device__ __forceinline__ float ldm(const MapRef& m, int b, int i, int c) {
  return m.p[(size_t)b * m.img + (size_t)i * m.pix + (size_t)c * m.ch];
}
__device__ __forceinline__ void stm(const MapOut& m, int b, int i, int c, float v) {
  m.p[(size_t)b * m.img + (size_t)i * m.pix + (size_t)c * m.ch] = v;
}

__global__ void __launch_bounds__(LT) loss_kernel(const LossParams p) {
  __shared__ uint32_t keys[MAPN];
  __shared__ unsigned char mask[MAPN];
  __shared__ unsigned char lmm[4][MAPN];
  __shared__ uint32_t hist[256];
  __shared__ double red[LT / 32];
  __shared__ int sc[4 + LT / 32 + 1];
  __shared__ int lmx[4], lmy[4];

  const int b = blockIdx.x, tid = threadIdx.x;
  const float* bb = p.bbox + 4 * b;
  const bool positive = p.labels ? (p.labels[b] != 0.f) : true;
  const bool has_lm = p.variant >= 1;

  // ---- batch-global positive count -> negative quota (:2864-2876)
  int pos;
  if (p.count_slots) {
    // every rank's count of THIS step, written by the peers' count_exchange kernels at the start of their step (long
    // before this kernel runs); the tag (= step number) tells a stale slot from the current one
    const int W = p.count_world;
    const unsigned long long tag = ((const volatile unsigned long long*)p.count_slots)[2 * W];
    double cnt = 0.0;
    if (tid < W) {
      const volatile unsigned long long* slot = (const volatile unsigned long long*)p.count_slots + ((tag - 1ull) & 1ull) * W + tid;
      unsigned long long v = *slot;
      const long long t0 = clock64();
      while ((v >> 32) != (tag & 0xFFFFFFFFull)) {
        if (clock64() - t0 > 20000000000LL) {  // ~10 s: a peer died
          printf("dbx: count exchange timeout (rank slot %d, want tag %llu, have %llu)\n", tid, tag, v >> 32);
          __trap();
        }
        v = *slot;
      }
      cnt = (double)(unsigned int)(v & 0xFFFFFFFFull);
    }
    pos = (int)(block_sum(cnt, red) + 0.5);
  } else if (p.global_pos_ptr) {
    pos = *p.global_pos_ptr;
  } else if (p.global_pos >= 0) {
    pos = p.global_pos;
  } else {
    double cnt = 0.0;
    for (int i = tid; i < p.B; i += LT) {
      if (p.labels && p.labels[i] == 0.f) continue;
      const Box s = score_box(p.bbox + 4 * i);
      cnt += (double)((s.x1 - s.x0) * (s.y1 - s.y0));
    }
    pos = (int)(block_sum(cnt, red) + 0.5);
  }
  const int gbatch = p.global_batch > 0 ? p.global_batch : p.B;
  const int neg_num = (int)((double)pos / (double)gbatch + 0.5);
  const int half = (int)((double)neg_num * 0.5 + 0.5);
  if (b == 0 && tid == 0 && p.info) {
    p.info[0] = half; p.info[1] = pos;
    p.info[2] = (p.rand_idx && half > p.rand_stride) ? 1 : 0;  // the caller gave fewer random negatives than the quota
  }

  Box sbox = clip(0, 0, 0, 0), gout = sbox, gin = sbox;
  if (positive) { sbox = score_box(bb); gray_boxes(bb, gout, gin); }

  // ---- classification mask: positives + hard negatives + random negatives, then gray zone
  for (int i = tid; i < MAPN; i += LT) {
    const int y = i / MAPW, x = i - y * MAPW;
    const bool g = positive && inbox(sbox, x, y);
    const float s = ldm(p.src[LOSS_SCORE], b, i, 0);
    const float d = s - (g ? 1.f : 0.f);
    keys[i] = g ? 0u : __float_as_uint(d * d);  // (s-gt)^2 * (1-gt) >= 0
    mask[i] = g ? 1 : 0;
  }
  __syncthreads();
  topk_mark(keys, half, mask, hist, sc);
  if (p.rand_idx) {
    for (int j = tid; j < half && j < p.rand_stride; j += LT) {
      const long long idx = p.rand_idx[(size_t)b * p.rand_stride + j];
      if (idx >= 0 && idx < MAPN) mask[idx] = 1;
    }
  }
  __syncthreads();
  if (positive) {
    for (int i = tid; i < MAPN; i += LT) {
      const int y = i / MAPW, x = i - y * MAPW;
      if (inbox(gin, x, y)) mask[i] = 1;
      else if (inbox(gout, x, y)) mask[i] = 0;
    }
  }
  __syncthreads();

  // ---- landmark masks (:2660-2701)
  if (has_lm) {
    if (tid < 4) {
      int x = -1, y = -1;
      if (positive && p.vertices) {
        const float* v = p.vertices + 8 * b;
        x = (int)__fadd_rn(v[2 * tid], 0.5f);
        y = (int)__fadd_rn(v[2 * tid + 1], 0.5f);
        if (p.clamp_lm) { x = x < MAPW ? x : MAPW - 1; y = y < MAPW ? y : MAPW - 1; }
        if (x < 0 || x >= MAPW || y < 0 || y >= MAPW) { x = -1; y = -1; }  // reference would raise IndexError
      }
      lmx[tid] = x; lmy[tid] = y;
    }
    __syncthreads();
    for (int k = 0; k < 4; ++k) {
      const int px = lmx[k], py = lmy[k];
      for (int i = tid; i < MAPN; i += LT) {
        const int y = i / MAPW, x = i - y * MAPW;
        const bool g = (x == px && y == py);
        const float d = ldm(p.src[LOSS_LM], b, i, k) - (g ? 1.f : 0.f);
        keys[i] = g ? 0u : __float_as_uint(d * d);
        lmm[k][i] = g ? 1 : 0;
      }
      __syncthreads();
      top1_mark(keys, lmm[k], hist);  // torch.topk(k=1) (:2672): an arg-max
      if (tid == 0 && p.lm_rand_idx) {
        const long long idx = p.lm_rand_idx[(size_t)b * 4 + k];
        if (idx >= 0 && idx < MAPN) lmm[k][idx] = 1;
      }
      __syncthreads();
      if (px >= 0) {  // 5x5 ignore zone with python-slice clipping, centre restored
        const Box z = clip(px - 2, px + 3, py - 2, py + 3);
        for (int i = tid; i < 25; i += LT) {
          const int y = z.y0 + i / 5, x = z.x0 + i % 5;
          if (x < z.x1 && y < z.y1) lmm[k][y * MAPW + x] = 0;
        }
        __syncthreads();
        if (tid == 0) lmm[k][py * MAPW + px] = 1;
        __syncthreads();
      }
    }
  }

  // ---- masked sums and gradients (:2914-2918, :2704-2723, :2159-2180)
  const float ldet = p.variant == 0 ? 1.f : p.lambda_det;
  double acc = 0.0;
  for (int i = tid; i < MAPN; i += LT) {
    const int y = i / MAPW, x = i - y * MAPW;
    const float m = (float)mask[i];
    const float g = (positive && inbox(sbox, x, y)) ? 1.f : 0.f;
    float dh[17];
#pragma unroll
    for (int c = 0; c < 17; ++c) dh[c] = 0.f;
    float t = ldm(p.src[LOSS_SCORE], b, i, 0) - g;
    float part = m * t * t;                      // cls
    dh[0] = 2.f * ldet * m * t;
    const float mg = m * g;
    float loc_part = 0.f;
    {
      const float gt[4] = {(float)x - bb[0], (float)y - bb[1], (float)x - bb[2], (float)y - bb[3]};
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float gtc = positive ? gt[c] : 0.f;
        t = ldm(p.src[LOSS_LOC], b, i, c) - gtc;
        loc_part += mg * t * t;
        dh[1 + c] = 2.f * ldet * p.lambda_loc * mg * t;
      }
    }
    double li = (double)ldet * ((double)part + (double)p.lambda_loc * (double)loc_part);
    if (has_lm) {
      float lm_part = 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float mk = (float)lmm[k][i];
        const float gk = (x == lmx[k] && y == lmy[k]) ? 1.f : 0.f;
        t = ldm(p.src[LOSS_LM], b, i, k) - gk;
        lm_part += mk * t * t;
        dh[5 + k] = 2.f * p.lambda_lm * mk * t;
      }
      li += (double)p.lambda_lm * (double)lm_part;
      if (p.variant == 2) {
        float ll = 0.f;
        const float* v = p.vertices + 8 * b;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float gtc = positive ? (((c & 1) ? (float)y : (float)x) - v[c]) : 0.f;
          t = ldm(p.src[LOSS_LMLOC], b, i, c) - gtc;
          ll += mg * t * t;
          dh[9 + c] = 2.f * mg * t;
        }
        li += (double)ll;
      }
      const float r = ldm(p.src[LOSS_RF], b, i, 0);
      t = r - g;
      li += (double)(m * t * t);
      const float dr = 2.f * m * t;
      if (p.d_rf) {
        uint4* o = reinterpret_cast<uint4*>(p.d_rf + ((size_t)b * MAPN + i) * 64);
        o[0] = make_uint4(pack_bf16x2(dr, 0.f), 0u, 0u, 0u);
#pragma unroll
        for (int q = 1; q < 8; ++q) o[q] = make_uint4(0u, 0u, 0u, 0u);
      }
      if (p.d_rf_f32) p.d_rf_f32[((size_t)b * MAPN + i) * p.RC] = dr;
      if (p.dst[LOSS_RF].p) stm(p.dst[LOSS_RF], b, i, 0, dr);
    }
    acc += li;
    if (p.d_head) {
      uint4* o = reinterpret_cast<uint4*>(p.d_head + ((size_t)b * MAPN + i) * 64);
      o[0] = make_uint4(pack_bf16x2(dh[0], dh[1]), pack_bf16x2(dh[2], dh[3]), pack_bf16x2(dh[4], dh[5]),
                        pack_bf16x2(dh[6], dh[7]));
      o[1] = make_uint4(pack_bf16x2(dh[8], dh[9]), pack_bf16x2(dh[10], dh[11]), pack_bf16x2(dh[12], dh[13]),
                        pack_bf16x2(dh[14], dh[15]));
      o[2] = make_uint4(pack_bf16x2(dh[16], 0.f), 0u, 0u, 0u);
#pragma unroll
      for (int q = 3; q < 8; ++q) o[q] = make_uint4(0u, 0u, 0u, 0u);
    }
    if (p.d_head_f32) {
      float* o = p.d_head_f32 + ((size_t)b * MAPN + i) * p.HC;
      for (int c = 0; c < p.HC; ++c) o[c] = c < 17 ? dh[c] : 0.f;
    }
    if (p.dst[LOSS_SCORE].p) stm(p.dst[LOSS_SCORE], b, i, 0, dh[0]);
    if (p.dst[LOSS_LOC].p) {
#pragma unroll
      for (int c = 0; c < 4; ++c) stm(p.dst[LOSS_LOC], b, i, c, dh[1 + c]);
    }
    if (p.dst[LOSS_LM].p) {
#pragma unroll
      for (int c = 0; c < 4; ++c) stm(p.dst[LOSS_LM], b, i, c, dh[5 + c]);
    }
    if (p.dst[LOSS_LMLOC].p) {
#pragma unroll
      for (int c = 0; c < 8; ++c) stm(p.dst[LOSS_LMLOC], b, i, c, dh[9 + c]);
    }
    if (p.mask_out) p.mask_out[(size_t)b * MAPN + i] = mask[i];
    if (p.lm_mask_out && has_lm)
      for (int k = 0; k < 4; ++k) p.lm_mask_out[((size_t)b * 4 + k) * MAPN + i] = lmm[k][i];
  }
  const double total = block_sum(acc, red);
  if (tid == 0) {
    p.loss_partial[b] = (float)total;
    __threadfence();
    const unsigned int done = atomicAdd(p.counter, 1u);
    if (done == (unsigned int)p.B - 1) {  // last CTA: fixed-order sum over samples -> deterministic loss
      __threadfence();
      double s = 0.0;
      for (int i = 0; i < p.B; ++i) s += (double)((volatile float*)p.loss_partial)[i];
      *p.loss = (float)s;
      *p.counter = 0u;
    }
  }
}

// Positive pixels of a shard (sum of clipped score boxes) -> *out (int). One block.
__global__ void count_positives_kernel(const float* __restrict__ bbox, const float* __restrict__ labels, int B,
                                       int* __restrict__ out) {
  __shared__ double red[LT / 32];
  double cnt = 0.0;
  for (int i = threadIdx.x; i < B; i += LT) {
    if (labels && labels[i] == 0.f) continue;
    const Box s = score_box(bbox + 4 * i);
    cnt += (double)((s.x1 - s.x0) * (s.y1 - s.y0));
  }
  const double tot = block_sum(cnt, red);
  if (threadIdx.x == 0) *out = (int)(tot + 0.5);
}

// One block: this shard's positives -> slot [parity][rank] of every rank's buffer, tagged with the step number.
__global__ void count_exchange_kernel(const float* __restrict__ bbox, const float* __restrict__ labels, int B,
                                      PeerSlots peers, unsigned long long* __restrict__ local) {
  __shared__ double red[LT / 32];
  __shared__ unsigned long long word;
  double cnt = 0.0;
  for (int i = threadIdx.x; i < B; i += LT) {
    if (labels && labels[i] == 0.f) continue;
    const Box s = score_box(bbox + 4 * i);
    cnt += (double)((s.x1 - s.x0) * (s.y1 - s.y0));
  }
  const double tot = block_sum(cnt, red);
  const int W = peers.world;
  if (threadIdx.x == 0) {
    const unsigned long long tag = local[2 * W] + 1ull;   // step number, kept on the device (CUDA-graph friendly)
    local[2 * W] = tag;
    word = (tag << 32) | (unsigned long long)(unsigned int)(tot + 0.5);
  }
  __syncthreads();
  if ((int)threadIdx.x < W) {
    const unsigned long long w = word;
    unsigned long long* dst = peers.p[threadIdx.x] + (((w >> 32) - 1ull) & 1ull) * W + peers.rank;
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(dst), "l"(w) : "memory");
  }
}

int count_exchange(const float* bbox, const float* labels, int B, const PeerSlots& peers, unsigned long long* local,
                   cudaStream_t st) {
  if (!bbox || !local || B <= 0 || peers.world < 1 || peers.world > 16 || peers.rank < 0 || peers.rank >= peers.world)
    return DBX_ERR_ARG;
  for (int r = 0; r < peers.world; ++r) if (!peers.p[r]) return DBX_ERR_ARG;
  count_exchange_kernel<<<1, LT, 0, st>>>(bbox, labels, B, peers, local);
  return (int)cudaGetLastError();
}

int count_positives(const float* bbox, const float* labels, int B, int* out, cudaStream_t st) {
  if (!bbox || !out || B <= 0) return DBX_ERR_ARG;
  count_positives_kernel<<<1, LT, 0, st>>>(bbox, labels, B, out);
  return (int)cudaGetLastError();
}

int loss_fwd_bwd(const LossParams& pin, cudaStream_t st) {
  LossParams p = pin;
  if (!p.bbox || !p.loss_partial || !p.loss || !p.counter || p.B <= 0) return DBX_ERR_ARG;
  if (p.variant < 0 || p.variant > 2) return DBX_ERR_ARG;
  if (p.head) {  // one interleaved NHWC buffer (the engine's head_out / rf_out): derive the five groups
    if (p.HC < (p.variant == 2 ? 17 : (p.variant == 1 ? 9 : 5))) return DBX_ERR_ARG;
    if (p.variant >= 1 && (!p.rf || p.RC < 1)) return DBX_ERR_ARG;
    const int first[4] = {0, 1, 5, 9};
    for (int g = 0; g < 4; ++g) p.src[g] = MapRef{p.head + first[g], (long)MAPN * p.HC, (long)p.HC, 1L};
    p.src[LOSS_RF] = MapRef{p.rf, (long)MAPN * p.RC, (long)p.RC, 1L};
  } else {
    if (p.d_head_f32 || p.d_rf_f32) return DBX_ERR_ARG;  // the interleaved fp32 gradients belong to the NHWC form
  }
  if (!p.src[LOSS_SCORE].p || !p.src[LOSS_LOC].p) return DBX_ERR_ARG;
  if (p.variant >= 1 && (!p.src[LOSS_LM].p || !p.src[LOSS_RF].p || !p.vertices)) return DBX_ERR_ARG;
  if (p.variant == 2 && !p.src[LOSS_LMLOC].p) return DBX_ERR_ARG;
  loss_kernel<<<p.B, LT, 0, st>>>(p);
  return (int)cudaGetLastError();
}

}  // namespace dbx
