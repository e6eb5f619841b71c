// densebox_b200 — the two tcgen05 contraction kernels of the DenseBox hot path.
//
//   conv_fprop : implicit-GEMM convolution, rows = output pixels (a tw x th x tn box fetched by one 4-D TMA per
//                filter tap and 64-channel slice, zero padding = TMA out-of-bounds fill), cols = output channels.
//                A and B are K-major SWIZZLE_128B tiles, D lives in TMEM (double-buffered), one elected thread
//                issues tcgen05.mma, four epilogue warps drain TMEM -> bias/ReLU/mask -> global.
//                Used for every 3x3/5x5/1x1 convolution forward AND for every data-gradient (dgrad = fprop with
//                the flipped/transposed filter).  Replaces the cuDNN calls behind DenseBox.py:185-224.
//   conv_wgrad : weight gradient, rows = output channels, cols = (tap, input channel), K = pixels.  Both operands
//                are MN-major SWIZZLE_128B tiles (the NHWC tensors are used as they lie), split-K over pixel
//                boxes, fp32 red.add epilogue.  Replaces the cuDNN wgrad behind loss.backward() (DenseBox.py:2925).
//
// Both kernels are persistent (one CTA per SM, static round-robin tile schedule) and warp-specialised:
//   warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2..5 = epilogue.
#include "dbx_common.h"
#include "dbx_ptx.cuh"
#include "dbx_epilogue.cuh"
#include <mutex>
#include <stdlib.h>

namespace dbx {

static constexpr int kThreads = 192;          // wgrad: 1 TMA + 1 MMA + 4 epilogue warps
static constexpr int kFpropThreads = 320;     // fprop: 1 TMA + 1 MMA + 8 epilogue warps (2 per TMEM lane quarter)
static constexpr int kMaxStages = 8;
static constexpr int kSmemBudget = 230400;      // usable dynamic smem (227 KB - alignment slack - static barriers)
static constexpr int kEpiBuf = 16384;           // one epilogue staging buffer: 128 rows x 64 bf16 channels

// ------------------------------------------------------------------------------------------------ host helpers
// Programmatic dependent launch of the tensor kernels: opt-in (DBX_PDL=1).  Measured +0.75 % on the sustained step
// (the prologue of a persistent kernel overlaps the tail of the previous one); all parity tests pass with it, but
// the gain does not justify making early-scheduled CTAs the default before it has soaked on multi-GPU runs.
const char* ab_env(const char* name) {
  static const int enabled = [] { const char* e = getenv("DBX_ENABLE_AB"); return (e && e[0] == '1') ? 1 : 0; }();
  return enabled ? getenv(name) : nullptr;
}

bool pdl_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = ab_env("DBX_PDL"); v = (e && e[0] == '1') ? 1 : 0; }
  return v == 1;
}

static int current_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  return (dev >= 0 && dev < kMaxDevices) ? dev : 0;
}

int num_sms() {
  static int n[kMaxDevices] = {0};
  const int dev = current_device();
  if (n[dev] == 0) {
    int v = 0;
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    n[dev] = v > 0 ? v : 148;
  }
  return n[dev];
}

static int g_tensor_sm_limit = 0;
int set_tensor_sm_limit(int n) { const int old = g_tensor_sm_limit; g_tensor_sm_limit = n > 0 ? n : 0; return old; }
int tensor_sms() {
  int n = num_sms();
  if (g_tensor_sm_limit > 0 && g_tensor_sm_limit < n) n = g_tensor_sm_limit;
  if (n > 2) n &= ~1;  // CTA pairs
  return n;
}

int set_max_smem_once(const void* fn, int bytes, SmemAttrOnce* s) {
  const int dev = current_device();
  if (!s->done[dev]) {
    s->rc[dev] = (int)cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    s->done[dev] = 1;
  }
  return s->rc[dev];
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  });
  return fn;
}

int encode_act_map(CUtensorMap* m, const Act& a, const Tile& t) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return DBX_ERR_DRIVER;
  if (a.cs % 8 || a.coff % 8 || a.C <= 0) return DBX_ERR_ARG;
  cuuint64_t dims[4] = {(cuuint64_t)a.C, (cuuint64_t)a.W, (cuuint64_t)a.H, (cuuint64_t)a.N};
  cuuint64_t strides[3] = {(cuuint64_t)a.cs * 2, (cuuint64_t)a.W * a.cs * 2, (cuuint64_t)a.H * a.W * a.cs * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)t.tw, (cuuint32_t)t.th, (cuuint32_t)t.tn};
  cuuint32_t es[4] = {1, 1, 1, 1};
  void* base = (void*)((char*)a.ptr + (size_t)a.coff * 2);
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? DBX_OK : DBX_ERR_TMAP;
}

int encode_mat_map(CUtensorMap* m, const void* ptr, int rows, int cols, int box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return DBX_ERR_DRIVER;
  if (cols % 8) return DBX_ERR_ARG;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  cuuint32_t es[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)ptr, dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? DBX_OK : DBX_ERR_TMAP;
}

Tile choose_tile(int W, int H, int N, bool need_mult16) {
  Tile best{};
  double best_score = -1.0;
  for (int tw = 1; tw <= 128 && tw <= W; ++tw) {
    for (int th = 1; tw * th <= 128 && th <= H; ++th) {
      for (int tn = 1; tw * th * tn <= 128 && tn <= N; ++tn) {
        int rows = tw * th * tn;
        if (need_mult16 && (rows % 16)) continue;
        if (!need_mult16 && rows < 64 && (long)W * H * N >= 64) continue;
        long tiles = (long)((W + tw - 1) / tw) * ((H + th - 1) / th) * ((N + tn - 1) / tn);
        // fprop pays 128 MMA rows per tile whatever the box; wgrad pays the box rows (zero-filled K).
        double denom = need_mult16 ? (double)tiles * rows : (double)tiles * 128.0;
        double eff = (double)W * H * N / denom;
        double halo = (double)(tw + 2) * (th + 2) / ((double)tw * th);
        double score = eff + 1e-3 * (rows / 128.0) - 1e-4 * halo + 1e-6 * (tw / 128.0);
        if (score > best_score) {
          best_score = score;
          best.tw = tw; best.th = th; best.tn = tn;
          best.tiles_w = (W + tw - 1) / tw; best.tiles_h = (H + th - 1) / th; best.tiles_n = (N + tn - 1) / tn;
        }
      }
    }
  }
  return best;
}

static uint32_t tmem_cols_for(int n) {
  uint32_t c = 32;
  while ((int)c < n) c <<= 1;
  return c;
}

// ------------------------------------------------------------------------------------------------ fprop kernel
struct FpropParams {
  int tw, th, tn, tiles_w, tiles_h, tiles_n;
  int out_W, out_H, out_N;
  int R, S, pad;
  int cin, cin_blocks;
  int m_tiles, n_tiles, block_n;
  FastDiv fd_nt, fd_tw, fd_twh;   // division by n_tiles, tiles_w, tiles_w * tiles_h
  int cout;
  float* colsum; int csum_off;  // fused bias gradient: global fp32 [cout], smem offset of the per-CTA partial sums
  int stages, kps;         // pipeline stages; K blocks (64 channels of one tap) per stage
  uint32_t idesc, tmem_cols;
  const float* bias;
  int relu;
  const bf16* aux;
  int aux_cs, aux_coff, aux_mode;
  void* out;
  int out_cs, out_coff, out_fp32;
  const unsigned long long* rng; int rng_channels;  // aux_mode 3: in-place Philox dropout
  int cta2;                // 1: CTA pairs drive tcgen05.mma.cta_group::2 (launched with cluster size 2)
  int tma_epi, nbuf, nbuf_log2, nsb;  // bf16 outputs: epilogue staged through `nbuf` (2/4/8) smem boxes, `nsb` 64-column blocks/tile
  bf16* pool_out; int pool_cs, pool_coff; unsigned short* pool_idx; int pool_keep_full;  // fused 2x2 max-pool (ConvEpilogue)
};

// kMode: 1 / 2 / 4 = K blocks (one tap x 64 channels) per pipeline stage; 3 = column-box mode for 3x3 pad-1 layers:
// 8 x 16 pixel tiles, ONE 8 x 18 input box per (filter column s, channel slice) serves the three filter rows (a row
// shift = 8 pixels = one 1024-B swizzle atom, so the descriptor just starts r atoms further) and travels with its
// three filter tiles behind one barrier: 12 MMAs per barrier round trip, A traffic 3 x 18 KB instead of 9 x 16 KB.
// kEpi: epilogue flavour (dbx_epilogue.cuh), chosen by the launcher from ConvEpilogue.
template <bool kCta2, int kMode, int kEpi>
__global__ void __launch_bounds__(kFpropThreads, 1)
conv_fprop_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmX,
                  const FpropParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t full_bar[kMaxStages], empty_bar[kMaxStages], tfull_bar[2], tempty_bar[2], aux_bar[8];
  __shared__ uint32_t tmem_base_s;
  griddep_launch_dependents();  // persistent grid, fully resident: the next kernel may take the SMs we leave

  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // cta2: a pair of CTAs (one cluster, one TPC) works on two M-tiles with ONE tcgen05.mma.cta_group::2 per K step;
  // each CTA stages its own A tile and HALF of the B tile, which halves the weight traffic per SM.
  constexpr bool cta2 = kCta2;  // separate instantiations: a kernel holding cta_group::2 PTX must run in pairs
  uint32_t rank = 0u;
  if constexpr (cta2) rank = cluster_ctarank();
  const int u0 = cta2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;      // first work unit of this CTA (pair)
  const int ustep = cta2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int m_units = cta2 ? (p.m_tiles + 1) >> 1 : p.m_tiles;        // odd tail: the extra tile is fully out of bounds
  const int total = m_units * p.n_tiles;
  constexpr bool colbox = kMode == 3;
  constexpr uint32_t kColBox = 18u * 8u * 128u;               // 8 x 18 pixels x 64 channels
  const uint32_t a_bytes = colbox ? kColBox : (uint32_t)(p.tw * p.th * p.tn) * 128u;
  const uint32_t b_rows = cta2 ? (uint32_t)p.block_n >> 1 : (uint32_t)p.block_n;
  const uint32_t b_bytes = b_rows * 128u;
  const uint32_t slot_bytes = colbox ? kColBox + 3u * b_bytes : 16384u + b_bytes;  // one K block: A box + B box(es)
  constexpr int kps = colbox ? 1 : kMode;                     // K blocks per stage (compile time: unrolled issue)
  const uint32_t stage_bytes = (uint32_t)kps * slot_bytes;    // a stage holds kps K blocks behind ONE barrier
  const int num_kb = p.R * p.S * p.cin_blocks;
  // Unit order: output-channel tile fastest, so the CTAs running at the same time share the SAME pixel tile and the
  // activations stream from HBM once (the weights of all channel tiles stay L2-resident).  With the pixel tile
  // fastest the 1x1 head GEMM re-read its 177 MB input once per channel tile (ncu: 709 MB of DRAM reads).
#define DBX_UNIT_TILE(u, nt, mt) \
  const int mq_##mt = p.fd_nt.div(u), nt = (u) - mq_##mt * p.n_tiles, mt = cta2 ? 2 * mq_##mt + (int)rank : mq_##mt
  // pixel-tile index -> box origin (in tiles): n = mt / (tiles_w * tiles_h), h = rest / tiles_w, w = rest % tiles_w
#define DBX_TILE_ORIGIN(mt, tw_i, th_i, tn_i) \
  const int tn_i = p.fd_twh.div(mt), rem_##tn_i = (mt) - tn_i * (p.tiles_w * p.tiles_h), \
            th_i = p.fd_tw.div(rem_##tn_i), tw_i = rem_##tn_i - th_i * p.tiles_w

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < p.stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&tfull_bar[b], 1); mbar_init(&tempty_bar[b], cta2 ? 16 : 8); }
    for (int b = 0; b < 8; ++b) mbar_init(&aux_bar[b], 1);
    if (p.tma_epi) { tma_prefetch_desc(&tmO); if (p.aux_mode == 1 || p.aux_mode == 2) tma_prefetch_desc(&tmX); }
    fence_barrier_init();
  }
  if (warp == 1) {
    if constexpr (cta2) { tmem_alloc_2sm(&tmem_base_s, p.tmem_cols); tmem_relinquish_2sm(); }
    else { tmem_alloc(&tmem_base_s, p.tmem_cols); tmem_relinquish(); }
  }
  float* csum = p.colsum ? reinterpret_cast<float*>(smem + p.csum_off) : nullptr;
  if (csum) for (int c = threadIdx.x; c < 8 * p.cout; c += kFpropThreads) csum[c] = 0.f;
  tc_fence_before();
  if constexpr (cta2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  griddep_wait();  // everything above (barriers, TMEM, descriptors) overlapped the previous kernel's tail

  // Shared-window addresses are computed once: the hot loops below must stay a few dozen instructions per K block,
  // a single warp issues them back to back (round 1 profile: the old MMA loop spent ~650 cycles/K-block on address
  // arithmetic and per-instruction election loops and was the bottleneck of EVERY layer).
  uint32_t smem_base = smem_u32(smem);
  uint32_t full0 = smem_u32(&full_bar[0]), empty0 = smem_u32(&empty_bar[0]);
  uint32_t tfull0 = smem_u32(&tfull_bar[0]), tempty0 = smem_u32(&tempty_bar[0]);
  // opaque to the optimiser: otherwise ptxas re-derives every window address inside the hot loops
  // (S2R SR_CgaCtaId + LEA per barrier access, ~30 cycles of latency each on a single-warp issue loop)
  asm volatile("" : "+r"(smem_base), "+r"(full0), "+r"(empty0), "+r"(tfull0), "+r"(tempty0));

  if (warp == 0) {
    // ===================== TMA producer (whole warp walks the schedule, one elected lane issues) ===============
    int stage = 0; uint32_t phase = 0;
    for (int u = u0; u < total; u += ustep) {
      DBX_UNIT_TILE(u, nt, mt);
      DBX_TILE_ORIGIN(mt, twi, thi, tni);
      const int w0 = twi * p.tw - p.pad, h0 = thi * p.th - p.pad, n0 = tni * p.tn;
      const int brow = nt * p.block_n + (int)(rank * b_rows) * (cta2 ? 1 : 0);
      if constexpr (colbox) {
        for (int cb = 0; cb < p.cin_blocks; ++cb)
          for (int sx = 0; sx < 3; ++sx) {
            mbar_wait_a(empty0 + 8u * stage, phase ^ 1);
            if (elect_one_sync()) {
              const uint32_t fb = full0 + 8u * stage;
              const uint32_t sa = smem_base + (uint32_t)stage * stage_bytes;
              const int kc = sx * p.cin + cb * 64;              // filter column of tap (r = 0, s): + 3 * cin per row
              if constexpr (cta2) {
                if (rank == 0) mbar_arrive_expect_tx_a(fb, 2u * slot_bytes);
                tma_load_4d_2sm_a(&tmA, fb, sa, cb * 64, w0 + sx, h0, n0);
#pragma unroll
                for (int r = 0; r < 3; ++r)
                  tma_load_2d_2sm_a(&tmB, fb, sa + kColBox + (uint32_t)r * b_bytes, kc + r * 3 * p.cin, brow);
              } else {
                mbar_arrive_expect_tx_a(fb, slot_bytes);
                tma_load_4d_a(&tmA, fb, sa, cb * 64, w0 + sx, h0, n0);
#pragma unroll
                for (int r = 0; r < 3; ++r)
                  tma_load_2d_a(&tmB, fb, sa + kColBox + (uint32_t)r * b_bytes, kc + r * 3 * p.cin, brow);
              }
            }
            __syncwarp();
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
          }
        continue;
      }
      int kcol = 0, r = 0, sx = 0, cb = 0;
      for (int kb0 = 0; kb0 < num_kb; kb0 += kps) {
        int nk = num_kb - kb0; if (nk > kps) nk = kps;
        mbar_wait_a(empty0 + 8u * stage, phase ^ 1);
        if (elect_one_sync()) {
          const uint32_t fb = full0 + 8u * stage;
          uint32_t sa = smem_base + (uint32_t)stage * stage_bytes;
          if constexpr (cta2) {
            // both CTAs' bytes land on the leader's barrier; only the leader arms it
            if (rank == 0) mbar_arrive_expect_tx_a(fb, 2u * (uint32_t)nk * (a_bytes + b_bytes));
          } else {
            mbar_arrive_expect_tx_a(fb, (uint32_t)nk * (a_bytes + b_bytes));
          }
          int r2 = r, s2 = sx, cb2 = cb, kc2 = kcol;
#pragma unroll
          for (int j = 0; j < kps; ++j, sa += slot_bytes, kc2 += 64) {
            if (j >= nk) break;
            if constexpr (cta2) {
              tma_load_4d_2sm_a(&tmA, fb, sa, cb2 * 64, w0 + s2, h0 + r2, n0);
              tma_load_2d_2sm_a(&tmB, fb, sa + 16384u, kc2, brow);
            } else {
              tma_load_4d_a(&tmA, fb, sa, cb2 * 64, w0 + s2, h0 + r2, n0);
              tma_load_2d_a(&tmB, fb, sa + 16384u, kc2, brow);
            }
            if (++cb2 == p.cin_blocks) { cb2 = 0; if (++s2 == p.S) { s2 = 0; ++r2; } }
          }
        }
        __syncwarp();
#pragma unroll
        for (int j = 0; j < kps; ++j) {  // every lane tracks the (tap, channel slice) position: any lane may be elected
          if (j >= nk) break;
          kcol += 64;
          if (++cb == p.cin_blocks) { cb = 0; if (++sx == p.S) { sx = 0; ++r; } }
        }
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA; warp-uniform loop, one elected lane issues) ===============
    if (rank == 0) {
      int stage = 0; uint32_t phase = 0; int it = 0;
      const uint64_t desc_hi = umma_smem_desc_sw128(0, 16, 1024);  // constant fields of both operand descriptors
      for (int u = u0; u < total; u += ustep, ++it) {
        const int buf = it & 1; const uint32_t use = (uint32_t)(it >> 1);
        mbar_wait_a(tempty0 + 8u * buf, (use & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem + (uint32_t)(buf * p.block_n);
        if constexpr (colbox) {
          const int nst = 3 * p.cin_blocks;
          for (int st = 0; st < nst; ++st) {
            mbar_wait_a(full0 + 8u * stage, phase);
            tc_fence_after();
            if (elect_one_sync()) {
              const uint32_t a_lo = (smem_base + (uint32_t)stage * stage_bytes) >> 4;
              const uint64_t da0 = desc_hi | (uint64_t)a_lo, db0 = desc_hi | (uint64_t)(a_lo + (kColBox >> 4));
              const uint32_t bt = b_bytes >> 4;
#pragma unroll
              for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int k = 0; k < 4; ++k) {  // filter row r: the box shifted by r image rows = r atoms (64 x 16 B)
                  if constexpr (cta2)
                    umma_bf16_2sm(d_tmem, da0 + 64 * r + 2 * k, db0 + (uint64_t)r * bt + 2 * k, p.idesc,
                                  (uint32_t)((st | r | k) != 0));
                  else
                    umma_bf16(d_tmem, da0 + 64 * r + 2 * k, db0 + (uint64_t)r * bt + 2 * k, p.idesc,
                              (uint32_t)((st | r | k) != 0));
                }
              if constexpr (cta2) umma_commit_2sm_a(empty0 + 8u * stage, 3); else umma_commit_a(empty0 + 8u * stage);
            }
            __syncwarp();
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
          }
        } else
        for (int kb0 = 0; kb0 < num_kb; kb0 += kps) {
          int nk = num_kb - kb0; if (nk > kps) nk = kps;
          mbar_wait_a(full0 + 8u * stage, phase);
          tc_fence_after();
          if (elect_one_sync()) {
            uint32_t a_lo = (smem_base + (uint32_t)stage * stage_bytes) >> 4;
#pragma unroll
            for (int j = 0; j < kps; ++j, a_lo += slot_bytes >> 4) {
              if (j >= nk) break;
              const uint64_t da = desc_hi | (uint64_t)a_lo, db = desc_hi | (uint64_t)(a_lo + 1024u);  // B at +16 KB
#pragma unroll
              for (int k = 0; k < 4; ++k) {  // +32 B (2 x 16 B units) per UMMA_K = 16
                if constexpr (cta2) umma_bf16_2sm(d_tmem, da + 2 * k, db + 2 * k, p.idesc, (uint32_t)((kb0 | j | k) != 0));
                else umma_bf16(d_tmem, da + 2 * k, db + 2 * k, p.idesc, (uint32_t)((kb0 | j | k) != 0));
              }
            }
            if constexpr (cta2) umma_commit_2sm_a(empty0 + 8u * stage, 3); else umma_commit_a(empty0 + 8u * stage);
          }
          __syncwarp();
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
        if (elect_one_sync()) {
          if constexpr (cta2) umma_commit_2sm_a(tfull0 + 8u * buf, 3); else umma_commit_a(tfull0 + 8u * buf);
        }
        __syncwarp();
      }
    }
  } else {
    // ===================== epilogue (warps 2..5 -> TMEM lane quarters 2,3,0,1) =====================
    if (p.tma_epi) {
      // bf16 outputs: the shared TMA-staged epilogue (dbx_epilogue.cuh)
      EpiArgs ea;
      ea.bias = p.bias; ea.cout = p.cout; ea.relu = p.relu; ea.aux_mode = p.aux_mode;
      ea.block_n = p.block_n; ea.nsb = p.nsb; ea.nbuf_log2 = p.nbuf_log2;
      ea.tw = p.tw; ea.th = p.th; ea.box_rows = p.tw * p.th * p.tn;
      ea.out_W = p.out_W; ea.out_H = p.out_H; ea.rng = p.rng; ea.rng_channels = p.rng_channels;
      ea.csum = csum;
      ea.pool_out = p.pool_out; ea.pool_cs = p.pool_cs; ea.pool_coff = p.pool_coff; ea.pool_idx = p.pool_idx;
      ea.pool_keep_full = p.pool_keep_full;
      const int my_tiles = u0 < total ? (total - 1 - u0) / ustep + 1 : 0;
      auto tile_of = [&](int it, int& nt, int& w0, int& h0, int& n0) {
        DBX_UNIT_TILE(u0 + it * ustep, nt_, mt_);
        nt = nt_;
        DBX_TILE_ORIGIN(mt_, twi, thi, tni);
        w0 = twi * p.tw; h0 = thi * p.th; n0 = tni * p.tn;
      };
      epilogue_tma<cta2, kEpi>(ea, &tmO, &tmX, smem + (size_t)p.stages * stage_bytes, aux_bar, tfull_bar, tempty_bar, tmem,
                         my_tiles, tile_of);
    } else {
      const int q = warp & 3, half = (warp - 2) >> 2;
      const int row = q * 32 + lane;
      const int box_rows = p.tw * p.th * p.tn;
      int it = 0;
      for (int u = u0; u < total; u += ustep, ++it) {
        const int buf = it & 1; const uint32_t use = (uint32_t)(it >> 1);
        DBX_UNIT_TILE(u, nt, mt);
        DBX_TILE_ORIGIN(mt, twi, thi, tni);
        const int w = twi * p.tw + row % p.tw;
        const int h = thi * p.th + (row / p.tw) % p.th;
        const int n = tni * p.tn + row / (p.tw * p.th);
        const bool valid = row < box_rows && w < p.out_W && h < p.out_H && n < p.out_N;
        const size_t pix = ((size_t)n * p.out_H + h) * p.out_W + w;
        mbar_wait(&tfull_bar[buf], use & 1);
        tc_fence_after();
        const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * p.block_n);
        for (int c0 = half * 16; c0 < p.block_n; c0 += 32) {
          uint32_t v[16];
          tmem_ld_x16(taddr + c0, v);
          tmem_ld_wait();
          const int ch = nt * p.block_n + c0;
          if (valid && ch < p.cout) {
            float f[16];
  #pragma unroll
            for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[j]);
            if (p.bias) {
              const float4* bp = reinterpret_cast<const float4*>(p.bias + ch);
  #pragma unroll
              for (int j = 0; j < 4; ++j) {
                float4 b4 = __ldg(bp + j);
                f[4 * j] += b4.x; f[4 * j + 1] += b4.y; f[4 * j + 2] += b4.z; f[4 * j + 3] += b4.w;
              }
            }
            if (p.relu) {
  #pragma unroll
              for (int j = 0; j < 16; ++j) f[j] = fmaxf(f[j], 0.f);
            }
            if (p.aux_mode) {
              const uint4* ap = reinterpret_cast<const uint4*>(p.aux + pix * p.aux_cs + p.aux_coff + ch);
              uint4 a0 = __ldg(ap), a1 = __ldg(ap + 1);
              uint32_t au[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
  #pragma unroll
              for (int j = 0; j < 8; ++j) {
                float lo = bf16lo(au[j]), hi = bf16hi(au[j]);
                if (p.aux_mode == 1) {
                  f[2 * j] = lo > 0.f ? f[2 * j] : 0.f;
                  f[2 * j + 1] = hi > 0.f ? f[2 * j + 1] : 0.f;
                } else {
                  f[2 * j] *= lo;
                  f[2 * j + 1] *= hi;
                }
              }
            }
            if (p.out_fp32) {
              float4* op = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + pix * p.out_cs + p.out_coff + ch);
  #pragma unroll
              for (int j = 0; j < 4; ++j) op[j] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
            } else {
              uint4* op = reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(p.out) + pix * p.out_cs + p.out_coff + ch);
              op[0] = make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]),
                                 pack_bf16x2(f[6], f[7]));
              op[1] = make_uint4(pack_bf16x2(f[8], f[9]), pack_bf16x2(f[10], f[11]), pack_bf16x2(f[12], f[13]),
                                 pack_bf16x2(f[14], f[15]));
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) { if constexpr (cta2) mbar_arrive_cluster_relaxed(&tempty_bar[buf], 0); else mbar_arrive(&tempty_bar[buf]); }
      }
    }
  }
  tc_fence_before();
  __syncwarp();
  if constexpr (cta2) cluster_sync_all(); else __syncthreads();
  if (csum)  // one global atomic per channel and CTA
    for (int c = threadIdx.x; c < p.cout; c += kFpropThreads) {
      float v = 0.f;
#pragma unroll
      for (int g = 0; g < 8; ++g) v += csum[g * p.cout + c];
      if (v != 0.f) atomicAdd(p.colsum + c, v);
    }
  if (warp == 1) {
    tc_fence_after();
    if constexpr (cta2) tmem_dealloc_2sm(tmem, p.tmem_cols); else tmem_dealloc(tmem, p.tmem_cols);
  }
#undef DBX_UNIT_TILE
#undef DBX_TILE_ORIGIN
}


int conv_fprop(const Act& x, const void* wk, int R, int S, int pad, const Act& out, const ConvEpilogue& epi,
               int block_n, cudaStream_t stream) {
  if (!x.ptr || !wk || !out.ptr) return DBX_ERR_ARG;
  if (x.C % 64 || out.C % 16) return DBX_ERR_ARG;
  if (out.H != x.H + 2 * pad - R + 1 || out.W != x.W + 2 * pad - S + 1 || out.N != x.N) return DBX_ERR_ARG;
  const int out_align = epi.out_fp32 ? 4 : 8;
  if (out.cs % out_align || out.coff % out_align) return DBX_ERR_ARG;
  if ((epi.aux_mode == 1 || epi.aux_mode == 2) && (!epi.aux || epi.aux_cs % 8 || epi.aux_coff % 8)) return DBX_ERR_ARG;
  if (epi.aux_mode == 3 && (!epi.rng || epi.rng_channels % 128)) return DBX_ERR_ARG;
  if (epi.aux_mode < 0 || epi.aux_mode > 3) return DBX_ERR_ARG;
  if (R == 3 && S == 3 && pad == 1 && x.C == 64 && out.C == 64 && !epi.out_fp32 && block_n <= 0 &&
      (epi.aux_mode == 1 || epi.aux_mode == 2)) {  // masked epilogue (conv1_2 dgrad); without a mask colbox + CTA pairs wins
    const char* e = ab_env("DBX_HALO");
    if (!(e && e[0] == '0')) {  // 64->64 3x3 layers (conv1_2 fwd/dgrad): column-box kernel, resident filter (A/B: DBX_HALO=0)
      const int rc = conv3x3_halo(x, wk, out, epi, stream);
      if (rc != DBX_ERR_ARG) return rc;
    }
  }
  const bool auto_n = block_n <= 0;
  if (block_n <= 0) block_n = out.C >= 256 ? 256 : out.C;
  if (block_n % 16 || block_n > 256 || block_n < 16) return DBX_ERR_ARG;
  if (auto_n && R == 3 && S == 3 && pad == 1 && out.C >= 256 && out.C % 128 == 0 && !epi.out_fp32) {
    // Wave quantisation: a CTA pair takes whole (256-pixel x block_n) units.  On the 30 x 30 maps at B = 32 there are
    // 256 such units of width 256 for 74 pairs = 3.46 waves, i.e. 4 unit times; 128-wide units give 6.92 -> 7 half
    // unit times.  Measured (tools/bench_blockn.py): conv4_2 1 081 -> 1 475 TFLOP/s, conv4_1 956 -> 1 123; on the
    // 60 x 60 maps (6.92 waves at 256) nothing changes.  128-wide units re-read the input boxes twice: +4 % cost.
    const long m_pairs = ((long)((out.W + 7) / 8) * ((out.H + 15) / 16) * out.N + 1) / 2;
    const long pairs = tensor_sms() / 2 > 0 ? tensor_sms() / 2 : 1;
    const long u256 = m_pairs * ((out.C + 255) / 256), u128 = m_pairs * (out.C / 128);
    const double c256 = (double)((u256 + pairs - 1) / pairs) * 256.0;
    const double c128 = (double)((u128 + pairs - 1) / pairs) * 128.0 * 1.04;
    const char* e = ab_env("DBX_AUTO_BLOCK_N");
    if (c128 < c256 && !(e && e[0] == '0')) block_n = 128;
  }

  Tile t = choose_tile(out.W, out.H, out.N, false);
  int tma_epi = epi.out_fp32 ? 0 : 1;
  const bool bf16_out = !epi.out_fp32;
  { const char* e = ab_env("DBX_DIRECT_EPI"); if (e && e[0] == '1' && epi.aux_mode != 3) tma_epi = 0; }  // A/B switch
  // Column-box mode (see the kernel) for the 3x3 pad-1 layers on maps that 8 x 16 tiles cover without much waste.
  // Measured (tools/bench_colbox.py, B = 32, generic -> colbox + CTA pairs, TFLOP/s): conv2_1 dgrad 557 -> 1004,
  // conv2_2 975 -> 1322, conv3_1 dgrad 811 -> 1129, conv4_2 1161 -> 1281, conv3_2 1209 -> 1229: the generic mode is
  // bound by barrier round trips (one per 4 MMAs) and TMA delivery on the narrow layers.
  int colbox = 0;
  int colbox_max_n = 256;
  { const char* e = ab_env("DBX_COLBOX_MAX_N"); if (e) colbox_max_n = atoi(e); }
  if (R == 3 && S == 3 && pad == 1 && bf16_out && block_n <= colbox_max_n && block_n % 32 == 0 && epi.aux_mode != 3) {
    const double cover = (double)out.W * out.H / ((double)((out.W + 7) / 8 * 8) * ((out.H + 15) / 16 * 16));
    colbox = cover >= 0.87 ? 1 : 0;   // 60 x 60 and 30 x 30 maps: 0.879 (measured: still +30 % on conv3_1 dgrad)
  }
  { const char* e = ab_env("DBX_COLBOX_FPROP"); if (e && R == 3 && S == 3 && pad == 1 && bf16_out) colbox = atoi(e) != 0; }
  if (colbox) {
    t.tw = 8; t.th = 16; t.tn = 1;
    t.tiles_w = (out.W + 7) / 8; t.tiles_h = (out.H + 15) / 16; t.tiles_n = out.N;
  }
  // CTA pairs (tcgen05.mma.cta_group::2) whenever the B tile is worth halving
  // (measured round 1: +4..6 % on the N = 256 layers once the MMA issue loop was lean; before that the kernel was
  // issue-bound and pairing changed nothing.  ConvEpilogue::force_cta2 = 0 or DBX_CTA2=0 switches it off.)
  int cta2_min_n = colbox ? 64 : 256;
  { const char* e = ab_env("DBX_CTA2_MIN_N"); if (e) cta2_min_n = atoi(e); }
  const bool cta2_ok = bf16_out && block_n >= cta2_min_n && block_n % 32 == 0 && t.count() >= 2 && num_sms() >= 2;
  int cta2 = (cta2_ok && epi.force_cta2 != 0) ? 1 : 0;
  { const char* e = ab_env("DBX_CTA2"); if (e && cta2_ok) cta2 = e[0] == '1'; }  // A/B switch for measurements
  CUtensorMap tmA, tmB, tmO, tmX;
  Tile ta = t;
  if (colbox) ta.th = 18;
  int rc = encode_act_map(&tmA, x, ta);
  if (rc) return rc;
  rc = encode_mat_map(&tmB, wk, out.C, R * S * x.C, cta2 ? block_n / 2 : block_n);
  if (rc) return rc;
  if (tma_epi) {
    rc = encode_act_map(&tmO, out, t);
    if (rc) return rc;
    if (epi.aux_mode == 1 || epi.aux_mode == 2) {
      Act ax = out;
      ax.ptr = const_cast<void*>(epi.aux); ax.cs = epi.aux_cs; ax.coff = epi.aux_coff;
      rc = encode_act_map(&tmX, ax, t);
      if (rc) return rc;
    } else {
      tmX = tmO;
    }
  } else {
    if (epi.aux_mode && epi.out_fp32) return DBX_ERR_ARG;  // fp32 outputs (tiny head GEMMs) take no mask
    tmO = tmA; tmX = tmA;
  }

  FpropParams p{};
  p.tw = t.tw; p.th = t.th; p.tn = t.tn; p.tiles_w = t.tiles_w; p.tiles_h = t.tiles_h; p.tiles_n = t.tiles_n;
  p.out_W = out.W; p.out_H = out.H; p.out_N = out.N;
  p.R = R; p.S = S; p.pad = pad;
  p.cin = x.C; p.cin_blocks = x.C / 64;
  p.m_tiles = t.count(); p.n_tiles = (out.C + block_n - 1) / block_n; p.block_n = block_n;
  p.fd_nt = FastDiv::make(p.n_tiles); p.fd_tw = FastDiv::make(p.tiles_w); p.fd_twh = FastDiv::make(p.tiles_w * p.tiles_h);
  p.cout = out.C;
  p.cta2 = cta2;
  const int slot_bytes = colbox ? 18432 + 3 * (cta2 ? block_n / 2 : block_n) * 128
                                : 16384 + (cta2 ? block_n / 2 : block_n) * 128;
  // K blocks per stage: a barrier round trip of the single-warp issue loops costs ~250 ns (measured:
  // conv2_2 with neither TMA nor MMA still takes 0.113 ms of 0.152), which a 64-wide tile (128 cycles of MMA per K
  // block) cannot hide -> batch K blocks behind one barrier
  const int b_rows_cta = cta2 ? block_n / 2 : block_n;
  p.kps = b_rows_cta <= 64 ? 2 : 1;  // measured (tools/bench_kps.py): wider tiles lose more to the coarser pipeline
  { const char* e = ab_env("DBX_KPS"); if (e && atoi(e) >= 1 && atoi(e) <= 4) p.kps = atoi(e); }
  if (p.kps > R * S * (x.C / 64)) p.kps = R * S * (x.C / 64);
  if (colbox) p.kps = 1;
  p.tma_epi = tma_epi;
  p.nsb = (block_n + 63) / 64;
  p.nbuf = 4;
  if (tma_epi && (kSmemBudget - 4 * kEpiBuf) / slot_bytes < 4) p.nbuf = 2;  // keep >= 4 operand K blocks in flight
  if (colbox) p.nbuf = (block_n <= 64 && x.C <= 64) ? 4 : 2;  // measured (tools/bench_colbox.py)
  if (tma_epi && (epi.epi_bufs == 2 || epi.epi_bufs == 4 || epi.epi_bufs == 8)) p.nbuf = epi.epi_bufs;
  { const char* e = ab_env("DBX_EPI_BUFS"); if (e && tma_epi && (atoi(e) == 2 || atoi(e) == 4 || atoi(e) == 8)) p.nbuf = atoi(e); }
  p.nbuf_log2 = p.nbuf == 8 ? 3 : (p.nbuf == 4 ? 2 : 1);
  const int ring = tma_epi ? p.nbuf * kEpiBuf : 0;
  while (p.kps > 1 && (kSmemBudget - ring) / (p.kps * slot_bytes) < 2) p.kps >>= 1;
  if (p.kps == 3) p.kps = 2;
  const int stage_bytes = p.kps * slot_bytes;
  p.stages = (kSmemBudget - ring) / stage_bytes;
  if (p.stages > kMaxStages) p.stages = kMaxStages;
  { const char* e = ab_env("DBX_STAGES"); if (e && atoi(e) >= 2 && atoi(e) < p.stages) p.stages = atoi(e); }
  if (p.stages < 2) return DBX_ERR_ARG;
  // fused column sums (bias gradient of the layer below): need 8 x cout floats of spare shared memory behind the ring
  bool colsum_after = false;
  p.colsum = nullptr; p.csum_off = 0;
  if (epi.colsum) {
    if (epi.bias) return DBX_ERR_ARG;
    const int used = p.stages * stage_bytes + ring;
    const char* e = ab_env("DBX_FUSED_COLSUM");
    if (tma_epi && out.C % 2 == 0 && used + 8 * out.C * 4 <= kSmemBudget && !(e && e[0] == '0')) { p.colsum = epi.colsum; p.csum_off = used; }
    else colsum_after = true;   // no room (or fp32 output): same result from the stand-alone kernel after the launch
  }
  p.idesc = umma_idesc_bf16(cta2 ? 256 : 128, block_n, 0, 0);
  p.tmem_cols = tmem_cols_for(2 * block_n);
  p.bias = epi.bias; p.relu = epi.relu;
  p.aux = (const bf16*)epi.aux; p.aux_cs = epi.aux_cs; p.aux_coff = epi.aux_coff; p.aux_mode = epi.aux_mode;
  p.out = out.ptr; p.out_cs = out.cs; p.out_coff = out.coff; p.out_fp32 = epi.out_fp32;
  p.rng = epi.rng; p.rng_channels = epi.rng_channels;
  // fused 2x2 max-pool: needs the 8 x 16 column-box tiles (a warp of the epilogue = a 4 x 8 pixel patch), whole
  // 64-column blocks, even map sizes, no mask and no fused column sums
  bool pool_after = false;
  if (epi.pool_out) {
    const bool ok = colbox && tma_epi && epi.aux_mode == 0 && !epi.colsum && block_n % 64 == 0 && out.C % 64 == 0 &&
                    out.H % 2 == 0 && out.W % 2 == 0 && epi.pool_cs % 8 == 0 && epi.pool_coff % 8 == 0;
    if (ok) {
      p.pool_out = (bf16*)epi.pool_out; p.pool_cs = epi.pool_cs; p.pool_coff = epi.pool_coff;
      p.pool_idx = (unsigned short*)epi.pool_idx; p.pool_keep_full = epi.pool_keep_full;
    } else {
      pool_after = true;  // same result from the stand-alone kernel behind this launch (which then stores `out`)
    }
  }

  if (p.kps == 3) p.kps = 2;
  typedef void (*FpropFn)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const FpropParams);
#define DBX_FPROP_ROW(c2, e) \
  {conv_fprop_kernel<c2, 1, e>, conv_fprop_kernel<c2, 2, e>, conv_fprop_kernel<c2, 4, e>, conv_fprop_kernel<c2, 3, e>}
  static const FpropFn fns[4][2][4] = {
      {DBX_FPROP_ROW(false, kEpiPlain), DBX_FPROP_ROW(true, kEpiPlain)},
      {DBX_FPROP_ROW(false, kEpiMask), DBX_FPROP_ROW(true, kEpiMask)},
      {DBX_FPROP_ROW(false, kEpiPhilox), DBX_FPROP_ROW(true, kEpiPhilox)},
      {DBX_FPROP_ROW(false, kEpiPool), DBX_FPROP_ROW(true, kEpiPool)}};
#undef DBX_FPROP_ROW
  static SmemAttrOnce attr_once[4][2][4];
  const int fa = cta2 ? 1 : 0, fb = colbox ? 3 : (p.kps == 4 ? 2 : (p.kps == 2 ? 1 : 0));
  // epilogue flavour (the direct fp32 / DBX_DIRECT_EPI epilogue ignores it).  The pool test comes first: a pooled
  // launch has no mask (checked above).
  const int fe = p.pool_out ? kEpiPool : (epi.aux_mode == 3 ? kEpiPhilox : (epi.aux_mode ? kEpiMask : kEpiPlain));
  const FpropFn fn = fns[fe][fa][fb];
  { const int arc = set_max_smem_once((const void*)fn, kSmemBudget + 1024, &attr_once[fe][fa][fb]); if (arc) return arc; }
  const size_t smem = (size_t)p.stages * stage_bytes + ring + (p.colsum ? (size_t)8 * out.C * 4 : 0) + 1024;
  cudaLaunchConfig_t cfg{};
  cfg.blockDim = dim3(kFpropThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int grid;
  if (cta2) {
    const int units = ((p.m_tiles + 1) / 2) * p.n_tiles;
    const int pairs = tensor_sms() / 2;
    grid = 2 * (units < pairs ? units : pairs);
  } else {
    const int total = p.m_tiles * p.n_tiles;
    grid = total < tensor_sms() ? total : tensor_sms();
  }
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cta2 ? 2 : 1; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr; cfg.numAttrs = 2;
  cfg.gridDim = dim3(grid);
  rc = (int)cudaLaunchKernelEx(&cfg, fn, tmA, tmB, tmO, tmX, p);
  if (rc == 0 && colsum_after) rc = colsum(out, epi.colsum, stream);
  if (rc == 0 && pool_after) {
    Act po = out;
    po.ptr = epi.pool_out; po.H = out.H / 2; po.W = out.W / 2; po.cs = epi.pool_cs; po.coff = epi.pool_coff;
    rc = maxpool2x2_fwd(out, po, stream, epi.pool_idx);
  }
  return rc;
}

// ------------------------------------------------------------------------------------------------ wgrad kernel
struct WgradParams {
  int tw, th, tn, tiles_w, tiles_h, tiles_n;
  int R, S, pad;
  int kp;
  int cin_blocks, q_total, nb, block_n;
  int m_tiles, q_tiles, splits, boxes_total, boxes_per_split;
  FastDiv fd_tps, fd_qt, fd_tw, fd_twh, fd_cb, fd_s;  // division by tiles_per_split, q_tiles, tiles_w, tiles_w * tiles_h, cin_blocks, S
  int cout, ldw;
  float* dw;
  int stages;
  uint32_t idesc, tmem_cols;
  int colbox;  // Cin = 64, 3x3: one column-shifted 8x18 input box per stage serves the three filter rows (N = 192)
};

template <bool kCta2>
__global__ void __launch_bounds__(kThreads, 1)
conv_wgrad_kernel(const __grid_constant__ CUtensorMap tmDy, const __grid_constant__ CUtensorMap tmX,
                  const WgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t full_bar[kMaxStages], empty_bar[kMaxStages], tfull_bar[2], tempty_bar[2];
  __shared__ uint32_t tmem_base_s;
  griddep_launch_dependents();

  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // cta2: a CTA pair works on two 128-row output-channel tiles with one tcgen05.mma.cta_group::2 per K step; each CTA
  // stages its own dY boxes and HALF of the shifted-input boxes (the wgrad mainloop is bound by TMA delivery).
  constexpr bool cta2 = kCta2;
  uint32_t rank = 0u;
  if constexpr (cta2) rank = cluster_ctarank();
  const int u0 = cta2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int ustep = cta2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int m_units = cta2 ? (p.m_tiles + 1) >> 1 : p.m_tiles;
  const int nbl = cta2 ? p.nb >> 1 : p.nb;                    // input boxes staged by THIS CTA
  const uint32_t box_bytes = (uint32_t)p.kp * 128u;           // one [kp pixels][64 ch] tile
  const uint32_t stage_bytes = p.colbox ? 2u * box_bytes + 18432u : box_bytes * (2u + (uint32_t)nbl);
  const int tiles_per_split = m_units * p.q_tiles;
  const int total = tiles_per_split * p.splits;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmDy);
    tma_prefetch_desc(&tmX);
    for (int s = 0; s < p.stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&tfull_bar[b], 1); mbar_init(&tempty_bar[b], cta2 ? 8 : 4); }
    fence_barrier_init();
  }
  if (warp == 1) {
    if constexpr (cta2) { tmem_alloc_2sm(&tmem_base_s, p.tmem_cols); tmem_relinquish_2sm(); }
    else { tmem_alloc(&tmem_base_s, p.tmem_cols); tmem_relinquish(); }
  }
  tc_fence_before();
  if constexpr (cta2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  griddep_wait();

  const uint32_t smem_base = smem_u32(smem);
  const uint32_t full0 = smem_u32(&full_bar[0]), empty0 = smem_u32(&empty_bar[0]);
  const uint32_t tfull0 = smem_u32(&tfull_bar[0]), tempty0 = smem_u32(&tempty_bar[0]);

  if (warp == 0) {
    // ===================== TMA producer (warp-uniform schedule walk, one elected lane issues) =====================
    int stage = 0; uint32_t phase = 0;
    for (int t = u0; t < total; t += ustep) {
      const int split = p.fd_tps.div(t), rem = t - split * tiles_per_split;
      const int mq = p.fd_qt.div(rem);
      const int m = (cta2 ? 2 : 1) * mq + (int)rank, q0 = (rem - mq * p.q_tiles) * p.nb;
      int nvalid = p.q_total - q0; if (nvalid > p.nb) nvalid = p.nb;
      const int j0 = (int)rank * nbl;                           // first input box of this CTA
      int jn = nvalid - j0; if (jn > nbl) jn = nbl; if (jn < 0) jn = 0;
      const int b_begin = split * p.boxes_per_split;
      int b_end = b_begin + p.boxes_per_split; if (b_end > p.boxes_total) b_end = p.boxes_total;
      for (int b = b_begin; b < b_end; ++b) {
        const int nq = p.fd_twh.div(b), brem = b - nq * (p.tiles_w * p.tiles_h), hq = p.fd_tw.div(brem);
        const int w0 = (brem - hq * p.tiles_w) * p.tw, h0 = hq * p.th, n0 = nq * p.tn;
        mbar_wait_a(empty0 + 8u * stage, phase ^ 1);
        if (elect_one_sync()) {
          const uint32_t sa = smem_base + (uint32_t)stage * stage_bytes, fb = full0 + 8u * stage;
          if constexpr (cta2) {
            if (rank == 0) mbar_arrive_expect_tx_a(fb, box_bytes * (4u + (uint32_t)nvalid));  // both CTAs' bytes
            tma_load_4d_2sm_a(&tmDy, fb, sa, m * 128, w0, h0, n0);
            tma_load_4d_2sm_a(&tmDy, fb, sa + box_bytes, m * 128 + 64, w0, h0, n0);
          } else if (p.colbox) {
            // dY tile(s) + ONE input box shifted by s-1 columns, one row of halo above/below
            const bool two = m * 128 + 64 < p.cout;
            mbar_arrive_expect_tx_a(fb, (two ? 2u : 1u) * box_bytes + 18432u);
            tma_load_4d_a(&tmDy, fb, sa, m * 128, w0, h0, n0);
            if (two) tma_load_4d_a(&tmDy, fb, sa + box_bytes, m * 128 + 64, w0, h0, n0);
            // q0 = (channel slice, filter column): slice q0 / 3, column q0 % 3
            tma_load_4d_a(&tmX, fb, sa + 2u * box_bytes, (q0 / 3) * 64, w0 + q0 % 3 - 1, h0 - 1, n0);
          } else {
            mbar_arrive_expect_tx_a(fb, box_bytes * (2u + (uint32_t)nvalid));
            tma_load_4d_a(&tmDy, fb, sa, m * 128, w0, h0, n0);
            tma_load_4d_a(&tmDy, fb, sa + box_bytes, m * 128 + 64, w0, h0, n0);
          }
          for (int j = 0; j < (p.colbox ? 0 : jn); ++j) {
            const int qq = q0 + j0 + j, tap = p.fd_cb.div(qq), cb = qq - tap * p.cin_blocks;
            const int r = p.fd_s.div(tap), s = tap - r * p.S;
            if constexpr (cta2)
              tma_load_4d_2sm_a(&tmX, fb, sa + box_bytes * (2u + j), cb * 64, w0 + s - p.pad, h0 + r - p.pad, n0);
            else
              tma_load_4d_a(&tmX, fb, sa + box_bytes * (2u + j), cb * 64, w0 + s - p.pad, h0 + r - p.pad, n0);
          }
        }
        __syncwarp();
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (warp-uniform loop, one elected lane issues) =====================
    int stage = 0; uint32_t phase = 0; int it = 0;
    const int ksteps = p.kp / 16;
    // MN-major SW128: 64-channel atoms LBO apart, groups of 8 pixel rows SBO = 1024 B apart.
    const uint64_t desc_hi = umma_smem_desc_sw128(0, box_bytes, 1024);
    if (rank == 0)
    for (int t = u0; t < total; t += ustep, ++it) {
      const int buf = it & 1; const uint32_t use = (uint32_t)(it >> 1);
      const int split = p.fd_tps.div(t);
      const int b_begin = split * p.boxes_per_split;
      int b_end = b_begin + p.boxes_per_split; if (b_end > p.boxes_total) b_end = p.boxes_total;
      mbar_wait_a(tempty0 + 8u * buf, (use & 1) ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem + (uint32_t)(buf * p.block_n);
      for (int b = b_begin; b < b_end; ++b) {
        mbar_wait_a(full0 + 8u * stage, phase);
        tc_fence_after();
        if (elect_one_sync()) {
          const uint32_t a_lo = (smem_base + (uint32_t)stage * stage_bytes) >> 4;
          uint64_t da = desc_hi | (uint64_t)a_lo;
          // colbox: the three filter rows are the same box shifted by 8 pixel rows = one 1024-B atom (LBO = 1024)
          uint64_t db = (p.colbox ? umma_smem_desc_sw128(0, 1024, 1024) : desc_hi) |
                        (uint64_t)(a_lo + ((2u * box_bytes) >> 4));
          for (int ks = 0; ks < ksteps; ++ks, da += 128, db += 128) {  // 16 pixel rows = 2048 B = 128 x 16 B
            if constexpr (cta2) umma_bf16_2sm(d_tmem, da, db, p.idesc, (uint32_t)((b > b_begin) | (ks != 0)));
            else umma_bf16(d_tmem, da, db, p.idesc, (uint32_t)((b > b_begin) | (ks != 0)));
          }
          if constexpr (cta2) umma_commit_2sm_a(empty0 + 8u * stage, 3); else umma_commit_a(empty0 + 8u * stage);
        }
        __syncwarp();
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
      if (elect_one_sync()) {
        if constexpr (cta2) umma_commit_2sm_a(tfull0 + 8u * buf, 3); else umma_commit_a(tfull0 + 8u * buf);
      }
      __syncwarp();
    }
  } else {
    const int q = warp & 3;
    int it = 0;
    for (int t = u0; t < total; t += ustep, ++it) {
      const int buf = it & 1; const uint32_t use = (uint32_t)(it >> 1);
      const int rem = t - p.fd_tps.div(t) * tiles_per_split, mq = p.fd_qt.div(rem);
      const int m = (cta2 ? 2 : 1) * mq + (int)rank, q0 = (rem - mq * p.q_tiles) * p.nb;
      const int row = m * 128 + q * 32 + lane;
      const bool valid = row < p.cout;
      mbar_wait(&tfull_bar[buf], use & 1);
      tc_fence_after();
      const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * p.block_n);
      float* drow = p.dw + (size_t)row * p.ldw + (p.colbox ? (size_t)0 : (size_t)q0 * 64);
      for (int c0 = 0; c0 < p.block_n; c0 += 16) {
        uint32_t v[16];
        tmem_ld_x16(taddr + c0, v);
        tmem_ld_wait();
        // colbox: column block c0/64 is filter row r, q0 = 3 * slice + s -> tap r*3+s, channels 64*slice.. of dw[co][tap][ci]
        float* dst = p.colbox ? drow + ((c0 >> 6) * 3 + q0 % 3) * (p.cin_blocks * 64) + (q0 / 3) * 64 + (c0 & 63)
                              : drow + c0;
        if (valid && (p.colbox || (q0 + c0 / 64) < p.q_total)) {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            red_add_v4(dst + 4 * j, __uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                       __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) { if constexpr (cta2) mbar_arrive_cluster_relaxed(&tempty_bar[buf], 0); else mbar_arrive(&tempty_bar[buf]); }
    }
  }
  tc_fence_before();
  __syncwarp();
  if constexpr (cta2) cluster_sync_all(); else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if constexpr (cta2) tmem_dealloc_2sm(tmem, p.tmem_cols); else tmem_dealloc(tmem, p.tmem_cols);
  }
}

// ------------------------------------------------------------------------------------------------ wgrad, 64 -> 64
// conv1_2 (3x3 pad 1, Cin = Cout = 64): with output channels as M the 128-row MMA is half empty (round-1/2 profile:
// 849 TFLOP/s, 0.160 ms, 4 % of the step).  Here the roles are swapped: M = (tap, input channel) — the nine taps of
// the three column-shifted 8 x 18 input boxes are 4.5 tiles of 128 rows (two 64-channel atoms per tile: a row shift is
// one 1024-byte atom inside a box, the next column is the next box) — and N = Cout = 64, K = the 128 pixels of a box.
// All nine taps of a pixel box are accumulated from ONE load of (dY box + 3 input boxes): 5 MMAs of 128 x 64 per K
// step instead of 3 x (128 x 192 with half the rows wasted), 70 KB of operands per box instead of 102 KB.  Each CTA owns
// a contiguous range of pixel boxes and keeps ALL 9 x 64 x 64 partial sums in TMEM (5 x 64 columns) for its whole life:
// one epilogue per CTA (fp32 red.add into dw[co][tap * 64 + ci]).
struct Wgrad64Params {
  int tiles_w, tiles_h, boxes_total, boxes_per_cta;
  FastDiv fd_tw, fd_twh;
  float* dw;
  uint32_t idesc;
};
static constexpr uint32_t kW64Box = 18432u, kW64Dy = 16384u, kW64Stage = 3u * kW64Box + kW64Dy;  // 71 680 B
static constexpr int kW64Stages = 3;

__global__ void __launch_bounds__(kThreads, 1)
conv_wgrad64_kernel(const __grid_constant__ CUtensorMap tmDy, const __grid_constant__ CUtensorMap tmX,
                    const Wgrad64Params p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t full_bar[kW64Stages], empty_bar[kW64Stages], tfull_bar;
  __shared__ uint32_t tmem_base_s;
  griddep_launch_dependents();
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmDy);
    tma_prefetch_desc(&tmX);
    for (int s = 0; s < kW64Stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(&tfull_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(&tmem_base_s, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  griddep_wait();
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t full0 = smem_u32(&full_bar[0]), empty0 = smem_u32(&empty_bar[0]);
  const int b_begin = (int)blockIdx.x * p.boxes_per_cta;
  int b_end = b_begin + p.boxes_per_cta; if (b_end > p.boxes_total) b_end = p.boxes_total;

  if (warp == 0) {
    int stage = 0; uint32_t phase = 0;
    for (int b = b_begin; b < b_end; ++b) {
      const int nq = p.fd_twh.div(b), brem = b - nq * (p.tiles_w * p.tiles_h), hq = p.fd_tw.div(brem);
      const int w0 = (brem - hq * p.tiles_w) * 8, h0 = hq * 16;
      mbar_wait_a(empty0 + 8u * stage, phase ^ 1);
      if (elect_one_sync()) {
        const uint32_t sa = smem_base + (uint32_t)stage * kW64Stage, fb = full0 + 8u * stage;
        mbar_arrive_expect_tx_a(fb, kW64Stage);
#pragma unroll
        for (int s = 0; s < 3; ++s) tma_load_4d_a(&tmX, fb, sa + (uint32_t)s * kW64Box, 0, w0 + s - 1, h0 - 1, nq);
        tma_load_4d_a(&tmDy, fb, sa + 3u * kW64Box, 0, w0, h0, nq);
      }
      __syncwarp();
      if (++stage == kW64Stages) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1) {
    int stage = 0; uint32_t phase = 0;
    // A tiles (MN-major SW128: two 64-channel atoms LBO apart, 8-pixel groups SBO = 1024 B apart): byte offset of the
    // first atom inside the stage and the distance to the second — tile j holds taps (r, s):
    //   0: (0,0),(1,0)   1: (2,0),(0,1)   2: (1,1),(2,1)   3: (0,2),(1,2)   4: (2,2), -
    const uint32_t a_off[5] = {0u, 2048u, kW64Box + 1024u, 2u * kW64Box, 2u * kW64Box + 2048u};
    const uint32_t a_lbo[5] = {1024u, kW64Box - 2048u, 1024u, 1024u, 1024u};
    uint64_t a_hi[5];
#pragma unroll
    for (int j = 0; j < 5; ++j) a_hi[j] = umma_smem_desc_sw128(0, a_lbo[j], 1024);
    const uint64_t b_hi = umma_smem_desc_sw128(0, 1024, 1024);
    for (int b = b_begin; b < b_end; ++b) {
      mbar_wait_a(full0 + 8u * stage, phase);
      tc_fence_after();
      if (elect_one_sync()) {
        const uint32_t s_lo = (smem_base + (uint32_t)stage * kW64Stage) >> 4;
        const uint64_t db = b_hi | (uint64_t)(s_lo + ((3u * kW64Box) >> 4));
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {       // 16 pixel rows = 2048 B = 128 x 16 B per K step
#pragma unroll
          for (int j = 0; j < 5; ++j)
            umma_bf16(tmem + (uint32_t)(64 * j), (a_hi[j] | (uint64_t)(s_lo + (a_off[j] >> 4))) + 128u * ks, db + 128u * ks,
                      p.idesc, (uint32_t)((b > b_begin) | (ks != 0)));
        }
        umma_commit_a(empty0 + 8u * stage);
      }
      __syncwarp();
      if (++stage == kW64Stages) { stage = 0; phase ^= 1; }
    }
    if (elect_one_sync()) umma_commit_a(smem_u32(&tfull_bar));
    __syncwarp();
  } else if (b_begin < b_end) {
    // epilogue: TMEM lane = (tap half, input channel), column = output channel.  dw[co][tap * 64 + ci] wants
    // consecutive ci per store, a thread holds consecutive co: transpose each 128 x 64 tile through shared memory (the
    // operand stages are free once the last MMA has committed) and add with 16-byte vector reds — a quarter of the
    // atomic operations of the scalar form, which all 148 CTAs issue at the same moment.
    // Measured: 0.160 -> 0.130 ms.  Not more, because with N = 64 every MMA re-reads its 4 KB A tile and the 2 KB B tile
    // from shared memory: 6 KB per 32 tensor cycles = 192 B/clk against the 128 B/clk an SM's shared memory delivers
    // (43 FLOP per shared-memory byte where 64 are needed): the 64-channel layers are shared-memory-bound by shape.
    const int q = warp & 3, et = (int)threadIdx.x - 64;   // 128 epilogue threads
    const int row = q * 32 + lane;
    const int taps[5][2] = {{0, 3}, {6, 1}, {4, 7}, {2, 5}, {8, -1}};
    float* T = reinterpret_cast<float*>(smem);            // [64 co][128 rows] fp32 = 32 KB
    mbar_wait(&tfull_bar, 0);
    tc_fence_after();
#pragma unroll 1
    for (int j = 0; j < 5; ++j) {
      const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(64 * j);
#pragma unroll
      for (int c0 = 0; c0 < 64; c0 += 32) {
        uint32_t v[32];
        tmem_ld_x32(taddr + c0, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) T[(c0 + i) * 128 + row] = __uint_as_float(v[i]);
      }
      named_bar_sync(2, 128);
      // thread: rows 4 (et & 31) .. + 3 (same tap half), output channels (et >> 5) + 4 k
      const int r4 = (et & 31) * 4, half = r4 >> 6, tap = taps[j][half];
      if (tap >= 0) {
        float* dst = p.dw + tap * 64 + (r4 & 63);
#pragma unroll 4
        for (int k = 0; k < 16; ++k) {
          const int co = (et >> 5) + 4 * k;
          const float4 f = *reinterpret_cast<const float4*>(T + co * 128 + r4);
          red_add_v4(dst + (size_t)co * 576, f.x, f.y, f.z, f.w);
        }
      }
      named_bar_sync(2, 128);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

static int conv_wgrad64(const Act& x, const Act& dy, float* dw, cudaStream_t stream) {
  Tile t{};
  t.tw = 8; t.th = 16; t.tn = 1;
  t.tiles_w = (dy.W + 7) / 8; t.tiles_h = (dy.H + 15) / 16; t.tiles_n = dy.N;
  Tile tx = t; tx.th = 18;
  CUtensorMap tmDy, tmX;
  int rc = encode_act_map(&tmDy, dy, t);
  if (rc) return rc;
  rc = encode_act_map(&tmX, x, tx);
  if (rc) return rc;
  Wgrad64Params p{};
  p.tiles_w = t.tiles_w; p.tiles_h = t.tiles_h; p.boxes_total = t.count();
  const int ctas = p.boxes_total < tensor_sms() ? p.boxes_total : tensor_sms();
  p.boxes_per_cta = (p.boxes_total + ctas - 1) / ctas;
  p.fd_tw = FastDiv::make(p.tiles_w); p.fd_twh = FastDiv::make(p.tiles_w * p.tiles_h);
  p.dw = dw;
  p.idesc = umma_idesc_bf16(128, 64, 1, 1);
  static SmemAttrOnce attr_once;
  const int smem = kW64Stages * (int)kW64Stage + 2048;   // + alignment slack + the over-read of the last half tile
  { const int arc = set_max_smem_once((const void*)conv_wgrad64_kernel, smem, &attr_once); if (arc) return arc; }
  cudaLaunchConfig_t cfg{};
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr; cfg.numAttrs = 1;
  cfg.gridDim = dim3((p.boxes_total + p.boxes_per_cta - 1) / p.boxes_per_cta);
  return (int)cudaLaunchKernelEx(&cfg, conv_wgrad64_kernel, tmDy, tmX, p);
}

int conv_wgrad(const Act& x, const Act& dy, int R, int S, int pad, float* dw, int block_n, cudaStream_t stream) {
  const int block_n_in = block_n;
  if (!x.ptr || !dy.ptr || !dw) return DBX_ERR_ARG;
  if (x.C % 64 || dy.C % 16) return DBX_ERR_ARG;
  if (dy.H != x.H + 2 * pad - R + 1 || dy.W != x.W + 2 * pad - S + 1 || dy.N != x.N) return DBX_ERR_ARG;
  if (R == 3 && S == 3 && pad == 1 && x.C == 64 && dy.C == 64 && block_n <= 0) {  // conv1_2: roles swapped, see above
    const char* e = ab_env("DBX_WGRAD64");
    if (!(e && e[0] == '0')) return conv_wgrad64(x, dy, dw, stream);
  }
  const int q_total = R * S * (x.C / 64);
  if (block_n <= 0) {
    block_n = 256;
    if (q_total * 64 < block_n) block_n = q_total * 64;
    // small layers: prefer a column split that leaves no ragged tail (e.g. 9 taps x 64 ch -> 3 x 192)
    if (q_total % 4 && q_total % 3 == 0 && q_total >= 3) block_n = 192;
  }
  if (block_n % 64 || block_n > 256 || block_n < 64) return DBX_ERR_ARG;

  Tile t = choose_tile(dy.W, dy.H, dy.N, true);
  // 3x3 layers with Cin <= 128 (conv1_2, conv2_1, conv2_2, conv3_1): column-box variant, see WgradParams::colbox.
  // Measured (tools/bench_wgrad.py, TFLOP/s, generic -> column box): conv2_2 932 -> 1294, conv3_1 876 -> 1173; from
  // Cin = 256 up the CTA-pair path with 4 input boxes per stage wins (conv3_2 1426 vs 1287, conv4_2 1445 vs 917).
  int colbox = (R == 3 && S == 3 && pad == 1 && x.C <= 128 && block_n_in <= 0) ? 1 : 0;
  { const char* e = ab_env("DBX_COLBOX");
    if (e && e[0] == '0') colbox = 0;
    if (e && e[0] == '1' && R == 3 && S == 3 && pad == 1 && block_n_in <= 0) colbox = 1; }
  CUtensorMap tmDy, tmX;
  int rc;
  if (colbox) {
    t.tw = 8; t.th = 16; t.tn = 1;
    t.tiles_w = (dy.W + 7) / 8; t.tiles_h = (dy.H + 15) / 16; t.tiles_n = dy.N;
    block_n = 192;
    Tile tx = t; tx.th = 18;
    rc = encode_act_map(&tmDy, dy, t);
    if (rc) return rc;
    rc = encode_act_map(&tmX, x, tx);
    if (rc) return rc;
  } else {
    rc = encode_act_map(&tmDy, dy, t);
    if (rc) return rc;
    rc = encode_act_map(&tmX, x, t);
    if (rc) return rc;
  }

  WgradParams p{};
  p.tw = t.tw; p.th = t.th; p.tn = t.tn; p.tiles_w = t.tiles_w; p.tiles_h = t.tiles_h; p.tiles_n = t.tiles_n;
  p.R = R; p.S = S; p.pad = pad;
  p.kp = t.rows();
  p.cin_blocks = x.C / 64; p.q_total = q_total; p.nb = block_n / 64; p.block_n = block_n;
  p.m_tiles = (dy.C + 127) / 128; p.q_tiles = (q_total + p.nb - 1) / p.nb;
  p.colbox = colbox;
  if (colbox) { p.q_tiles = 3 * p.cin_blocks; p.nb = 1; }  // q-tile = (channel slice, filter column s)
  p.boxes_total = t.count();
  // CTA pairs (cta_group::2) share the shifted-input boxes: worth it from two output-channel tiles up
  int cta2 = (!colbox && p.m_tiles >= 2 && block_n == 256 && num_sms() >= 2) ? 1 : 0;
  { const char* e = ab_env("DBX_CTA2"); if (e && e[0] == '0') cta2 = 0; }
  const int m_units = cta2 ? (p.m_tiles + 1) / 2 : p.m_tiles;
  const int tiles = m_units * p.q_tiles;
  const int workers = cta2 ? tensor_sms() / 2 : tensor_sms();
  int splits = workers / tiles;
  if (splits < 1) splits = 1;
  if (splits > p.boxes_total) splits = p.boxes_total;
  p.boxes_per_split = (p.boxes_total + splits - 1) / splits;
  p.splits = (p.boxes_total + p.boxes_per_split - 1) / p.boxes_per_split;
  p.cout = dy.C; p.ldw = R * S * x.C; p.dw = dw;
  p.fd_tps = FastDiv::make(tiles); p.fd_qt = FastDiv::make(p.q_tiles);
  p.fd_tw = FastDiv::make(p.tiles_w); p.fd_twh = FastDiv::make(p.tiles_w * p.tiles_h);
  p.fd_cb = FastDiv::make(p.cin_blocks); p.fd_s = FastDiv::make(S);
  const int stage_bytes = colbox ? 2 * p.kp * 128 + 18432 : p.kp * 128 * (2 + (cta2 ? p.nb / 2 : p.nb));
  p.stages = kSmemBudget / stage_bytes;
  if (p.stages > kMaxStages) p.stages = kMaxStages;
  if (p.stages < 2) return DBX_ERR_ARG;
  p.idesc = umma_idesc_bf16(cta2 ? 256 : 128, block_n, 1, 1);
  p.tmem_cols = tmem_cols_for(2 * block_n);

  static SmemAttrOnce attr_once[2];
  { const int arc = set_max_smem_once(cta2 ? (const void*)conv_wgrad_kernel<true> : (const void*)conv_wgrad_kernel<false>,
                                      kSmemBudget + 1024, &attr_once[cta2 ? 1 : 0]);
    if (arc) return arc; }
  const int total = tiles * p.splits;
  const size_t smem = (size_t)p.stages * stage_bytes + 1024;
  cudaLaunchConfig_t cfg{};
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cta2 ? 2 : 1; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr; cfg.numAttrs = 2;
  cfg.gridDim = dim3((cta2 ? 2 : 1) * (total < workers ? total : workers));
  if (cta2) return (int)cudaLaunchKernelEx(&cfg, conv_wgrad_kernel<true>, tmDy, tmX, p);
  return (int)cudaLaunchKernelEx(&cfg, conv_wgrad_kernel<false>, tmDy, tmX, p);
}

}  // namespace dbx
