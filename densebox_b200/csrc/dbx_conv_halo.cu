// densebox_b200 — 3x3 (pad 1) convolution for the narrow layers (Cin <= 128: conv1_2, conv2_1, conv2_2 and their data
// gradients).  The generic implicit-GEMM kernel fetches the input tile nine times (once per filter tap); on these
// layers the K loop is so short that the kernel is bound by TMA delivery, not by the tensor cores (round-1 profile:
// the MMA warp waits on the `full` barrier for 70 % of the K blocks of conv1_2).
// Here an 8-wide x 16-high output tile loads, per 64-channel slice, THREE column-shifted boxes of 8 x 18 pixels
// (one per filter column s).  The three filter rows r are row-shifted views of the same box: a shift by one image
// row is exactly one 8-pixel group = 1024 B = one SWIZZLE_128B atom, so the UMMA descriptor simply starts r atoms
// further — standard canonical K-major layout, no unaligned starts.  A-operand TMA traffic drops from 9 x 16 KB to
// 3 x 18 KB per tile and slice.  When the whole filter fits (Cin = Cout = 64: 72 KB) it is loaded once per CTA and
// stays resident; otherwise it streams through a ring of per-tap tiles.  Epilogue = the TMA-staged epilogue of
// conv_fprop_kernel (bias / ReLU / mask in a swizzled smem box, TMA store).
#include "dbx_common.h"
#include "dbx_ptx.cuh"
#include "dbx_epilogue.cuh"

namespace dbx {

static constexpr int kHaloThreads = 320;        // 1 producer + 1 MMA + 8 epilogue warps
static constexpr int kSBox = 18 * 8 * 128;      // one column-shifted box: 18 rows x 8 pixels x 64 channels = 18 KB
static constexpr int kMaxASlots = 8;
static constexpr int kMaxBSlots = 12;
static constexpr int kHaloSmem = 230400;

struct HaloParams {
  int tiles_w, tiles_h, m_tiles, n_tiles, block_n;
  FastDiv fd_mt, fd_tw, fd_twh;   // division by m_tiles, tiles_w, tiles_w * tiles_h
  int cin_blocks, cin;     // K per tap = cin (multiple of 64)
  int cout;
  int resident;            // whole filter resident in smem (n_tiles == 1)
  int a_slots, b_slots;    // ring depths (column boxes, per-tap filter tiles)
  int nbuf, nbuf_log2, nsb;  // epilogue staging boxes (2/4/8), 64-column blocks per tile
  uint32_t idesc, tmem_cols;
  const float* bias;
  int relu, aux_mode;
  float* colsum; int csum_off;  // fused bias gradient of the layer below (see ConvEpilogue::colsum)
};

template <bool kMask>   // epilogue flavour (dbx_epilogue.cuh): ReLU-backward / dropout mask tile, or bias + ReLU only
__global__ void __launch_bounds__(kHaloThreads, 1)
conv3x3_halo_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmX,
                    const HaloParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t afull[kMaxASlots], aempty[kMaxASlots], bfull[kMaxBSlots], bempty[kMaxBSlots];
  __shared__ uint64_t tfull_bar[2], tempty_bar[2], aux_bar[8];
  __shared__ uint32_t tmem_base_s;
  griddep_launch_dependents();

  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  const int warp = threadIdx.x >> 5;
  const uint32_t b_tile = (uint32_t)p.block_n * 128u;               // one (tap, 64-channel slice) filter tile
  const int kb_per_tile = 9 * p.cin_blocks;
  const int nb = p.resident ? kb_per_tile : p.b_slots;
  uint8_t* bsm = smem + (size_t)p.a_slots * kSBox;
  uint8_t* ring = bsm + (size_t)nb * b_tile;
  const int total = p.m_tiles * p.n_tiles;
  const uint32_t a_base = smem_u32(smem), b_base = smem_u32(bsm);
  const uint32_t afull0 = smem_u32(&afull[0]), aempty0 = smem_u32(&aempty[0]);
  const uint32_t bfull0 = smem_u32(&bfull[0]), bempty0 = smem_u32(&bempty[0]);
  const uint32_t tfull0 = smem_u32(&tfull_bar[0]), tempty0 = smem_u32(&tempty_bar[0]);

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmB); tma_prefetch_desc(&tmO);
    if (p.aux_mode) tma_prefetch_desc(&tmX);
    for (int s = 0; s < kMaxASlots; ++s) { mbar_init(&afull[s], 1); mbar_init(&aempty[s], 1); }
    for (int s = 0; s < kMaxBSlots; ++s) { mbar_init(&bfull[s], 1); mbar_init(&bempty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&tfull_bar[b], 1); mbar_init(&tempty_bar[b], 8); }
    for (int b = 0; b < 8; ++b) mbar_init(&aux_bar[b], 1);
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(&tmem_base_s, p.tmem_cols); tmem_relinquish(); }
  float* csum = p.colsum ? reinterpret_cast<float*>(smem + p.csum_off) : nullptr;
  if (csum) for (int c = threadIdx.x; c < 8 * p.cout; c += kHaloThreads) csum[c] = 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  griddep_wait();

  if (warp == 0) {
    // ===================== producer: column boxes + filter tiles (warp-uniform walk, elected lane issues) =========
    int as = 0; uint32_t aph = 0; int bs = 0; uint32_t bph = 0;
    if (p.resident && elect_one_sync()) {  // n_tiles == 1: the whole filter, once
      mbar_arrive_expect_tx_a(bfull0, (uint32_t)kb_per_tile * b_tile);
      for (int kb = 0; kb < kb_per_tile; ++kb)  // slot kb = (cb, s, r) in the order the MMA consumes them
        tma_load_2d_a(&tmB, bfull0, b_base + (uint32_t)kb * b_tile,
                      (((kb % 9) % 3) * 3 + (kb % 9) / 3) * p.cin + (kb / 9) * 64, 0);
    }
    __syncwarp();
    for (int t = blockIdx.x; t < total; t += gridDim.x) {
      const int nt = p.fd_mt.div(t), mt = t - nt * p.m_tiles;
      const int n0 = p.fd_twh.div(mt), rem = mt - n0 * (p.tiles_w * p.tiles_h);
      const int hq = p.fd_tw.div(rem), w0 = (rem - hq * p.tiles_w) * 8, h0 = hq * 16;
      for (int cb = 0; cb < p.cin_blocks; ++cb)
        for (int s = 0; s < 3; ++s) {
          mbar_wait_a(aempty0 + 8u * as, aph ^ 1);
          if (elect_one_sync()) {
            mbar_arrive_expect_tx_a(afull0 + 8u * as, (uint32_t)kSBox);
            tma_load_4d_a(&tmA, afull0 + 8u * as, a_base + (uint32_t)as * kSBox, cb * 64, w0 + s - 1, h0 - 1, n0);
          }
          __syncwarp();
          if (++as == p.a_slots) { as = 0; aph ^= 1; }
          if (!p.resident) {
            for (int r = 0; r < 3; ++r) {
              mbar_wait_a(bempty0 + 8u * bs, bph ^ 1);
              if (elect_one_sync()) {
                mbar_arrive_expect_tx_a(bfull0 + 8u * bs, b_tile);
                tma_load_2d_a(&tmB, bfull0 + 8u * bs, b_base + (uint32_t)bs * b_tile, (r * 3 + s) * p.cin + cb * 64,
                              nt * p.block_n);
              }
              __syncwarp();
              if (++bs == p.b_slots) { bs = 0; bph ^= 1; }
            }
          }
        }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const uint64_t desc_hi = umma_smem_desc_sw128(0, 16, 1024);
    if (p.resident && p.cin_blocks == 1) {
      // Fully unrolled issue path for the resident-filter case (conv1_2 and its data gradient): with N = 64 one MMA
      // lasts 32 cycles, so every instruction between two UTCHMMAs counts.
      mbar_wait_a(bfull0, 0);
      tc_fence_after();
      const uint32_t a_lo0 = a_base >> 4, b_lo0 = b_base >> 4, bt = b_tile >> 4;
      int it = 0; uint32_t as = 0, aph = 0;
      for (int t = blockIdx.x; t < total; t += gridDim.x, ++it) {
        const int buf = it & 1; const uint32_t use = (uint32_t)(it >> 1);
        mbar_wait_a(tempty0 + 8u * buf, (use & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem + (uint32_t)(buf * p.block_n);
#pragma unroll
        for (int s = 0; s < 3; ++s) {
          mbar_wait_a(afull0 + 8u * as, aph);
          tc_fence_after();
          if (elect_one_sync()) {
            const uint64_t da0 = desc_hi | (uint64_t)(a_lo0 + as * (uint32_t)(kSBox >> 4));
            const uint64_t db0 = desc_hi | (uint64_t)(b_lo0 + (uint32_t)(s * 3) * bt);
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_bf16(d_tmem, da0 + 64 * r + 2 * k, db0 + (uint64_t)r * bt + 2 * k, p.idesc,
                          (uint32_t)((s | r | k) != 0));
            umma_commit_a(aempty0 + 8u * as);
            if (s == 2) umma_commit_a(tfull0 + 8u * buf);
          }
          __syncwarp();
          if (++as == (uint32_t)p.a_slots) { as = 0; aph ^= 1; }
        }
      }
    } else {
    int as = 0; uint32_t aph = 0; int bs = 0; uint32_t bph = 0; int it = 0;
    if (p.resident) { mbar_wait_a(bfull0, 0); tc_fence_after(); }
    for (int t = blockIdx.x; t < total; t += gridDim.x, ++it) {
      const int buf = it & 1; const uint32_t use = (uint32_t)(it >> 1);
      mbar_wait_a(tempty0 + 8u * buf, (use & 1) ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem + (uint32_t)(buf * p.block_n);
      for (int cb = 0; cb < p.cin_blocks; ++cb)
        for (int s = 0; s < 3; ++s) {
          mbar_wait_a(afull0 + 8u * as, aph);
          tc_fence_after();
          const uint32_t a_lo = (a_base + (uint32_t)as * kSBox) >> 4;
          for (int r = 0; r < 3; ++r) {
            uint32_t b_lo;
            if (p.resident) {
              b_lo = (b_base + (uint32_t)((cb * 3 + s) * 3 + r) * b_tile) >> 4;
            } else {
              mbar_wait_a(bfull0 + 8u * bs, bph);
              tc_fence_after();
              b_lo = (b_base + (uint32_t)bs * b_tile) >> 4;
            }
            if (elect_one_sync()) {
              // filter row r = the box shifted by r image rows = r x 8 pixels x 128 B = r swizzle atoms (64 x 16 B)
              const uint64_t da = desc_hi | (uint64_t)(a_lo + 64u * r), db = desc_hi | (uint64_t)b_lo;
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_bf16(d_tmem, da + 2 * k, db + 2 * k, p.idesc, (uint32_t)((cb | s | r | k) != 0));
              if (!p.resident) umma_commit_a(bempty0 + 8u * bs);
              if (r == 2) umma_commit_a(aempty0 + 8u * as);
            }
            __syncwarp();
            if (!p.resident) { if (++bs == p.b_slots) { bs = 0; bph ^= 1; } }
          }
          if (++as == p.a_slots) { as = 0; aph ^= 1; }
        }
      if (elect_one_sync()) umma_commit_a(tfull0 + 8u * buf);
      __syncwarp();
    }
    }
  } else {
    // ===================== epilogue: the shared TMA-staged epilogue (dbx_epilogue.cuh) =====================
    EpiArgs ea;
    ea.bias = p.bias; ea.cout = p.cout; ea.relu = p.relu; ea.aux_mode = p.aux_mode;
    ea.block_n = p.block_n; ea.nsb = p.nsb; ea.nbuf_log2 = p.nbuf_log2;
    ea.tw = 8; ea.th = 16; ea.box_rows = 128;
    ea.out_W = 0; ea.out_H = 0; ea.rng = nullptr; ea.rng_channels = 0;
    ea.csum = csum;
    const int my_tiles = (int)blockIdx.x < total ? (total - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    auto tile_of = [&](int it, int& nt, int& w0, int& h0, int& n0) {
      const int t = (int)blockIdx.x + it * (int)gridDim.x;
      nt = p.fd_mt.div(t);
      const int mt = t - nt * p.m_tiles;
      n0 = p.fd_twh.div(mt);
      const int rem = mt - n0 * (p.tiles_w * p.tiles_h), hq = p.fd_tw.div(rem);
      w0 = (rem - hq * p.tiles_w) * 8; h0 = hq * 16;
    };
    epilogue_tma<false, kMask ? kEpiMask : kEpiPlain>(ea, &tmO, &tmX, ring, aux_bar, tfull_bar, tempty_bar, tmem,
                                                      my_tiles, tile_of);
  }
  tc_fence_before();
  __syncthreads();
  if (csum)
    for (int c = threadIdx.x; c < p.cout; c += kHaloThreads) {
      float v = 0.f;
#pragma unroll
      for (int g = 0; g < 8; ++g) v += csum[g * p.cout + c];
      if (v != 0.f) atomicAdd(p.colsum + c, v);
    }
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem, p.tmem_cols); }
}

static int encode_box(CUtensorMap* m, const Act& a, int bw, int bh) {
  Tile t{}; t.tw = bw; t.th = bh; t.tn = 1;
  return encode_act_map(m, a, t);
}

// Same contract as conv_fprop for R = S = 3, pad = 1, bf16 output; returns DBX_ERR_ARG when the shape is not one
// this kernel handles (the caller then uses conv_fprop).
int conv3x3_halo(const Act& x, const void* wk, const Act& out, const ConvEpilogue& epi, cudaStream_t stream) {
  if (!x.ptr || !wk || !out.ptr || epi.out_fp32) return DBX_ERR_ARG;
  if (x.C % 64 || x.C > 128 || out.C % 64 || out.H != x.H || out.W != x.W || out.N != x.N) return DBX_ERR_ARG;
  if (epi.aux_mode && (!epi.aux || epi.aux_cs % 8 || epi.aux_coff % 8)) return DBX_ERR_ARG;
  if (epi.colsum && epi.bias) return DBX_ERR_ARG;
  const int csum_bytes = epi.colsum ? (8 * out.C * 4 + 1023) / 1024 * 1024 : 0;
  HaloParams p{};
  p.block_n = out.C >= 128 ? 128 : 64;
  p.tiles_w = (out.W + 7) / 8; p.tiles_h = (out.H + 15) / 16;
  p.m_tiles = p.tiles_w * p.tiles_h * out.N;
  p.n_tiles = (out.C + p.block_n - 1) / p.block_n;
  p.cin_blocks = x.C / 64; p.cin = x.C; p.cout = out.C;
  p.fd_mt = FastDiv::make(p.m_tiles); p.fd_tw = FastDiv::make(p.tiles_w); p.fd_twh = FastDiv::make(p.tiles_w * p.tiles_h);
  const int b_tile = p.block_n * 128;
  const int kb_per_tile = 9 * p.cin_blocks;
  p.nsb = (p.block_n + 63) / 64;
  // smem plan: resident filter (whole filter, n_tiles == 1) when it leaves room for >= 4 column boxes and 2 staging
  // boxes; the short-K resident layers are epilogue-bound (ncu: 19 B/clk of operand traffic), so they trade column
  // boxes for a 4-deep staging ring (the mask prefetch needs it: D = nbuf / 2 sub-blocks of lead).
  p.resident = (p.n_tiles == 1 && kb_per_tile * b_tile <= 80 * 1024) ? 1 : 0;
  p.nbuf = 2;
  if (p.resident) {
    p.b_slots = 0;
    int avail = kHaloSmem - kb_per_tile * b_tile - csum_bytes;
    p.nbuf = 4; p.a_slots = (avail - p.nbuf * kEpiBox) / kSBox;
    if (p.a_slots < 4) { p.nbuf = 2; p.a_slots = (avail - p.nbuf * kEpiBox) / kSBox; }
    if (p.a_slots > kMaxASlots) p.a_slots = kMaxASlots;
    if (p.a_slots < 3) return DBX_ERR_ARG;
  } else {
    p.a_slots = 6;
    int avail = kHaloSmem - p.a_slots * kSBox - p.nbuf * kEpiBox - csum_bytes;
    p.b_slots = avail / b_tile;
    if (p.b_slots > kMaxBSlots) p.b_slots = kMaxBSlots;
    if (p.b_slots < 3) return DBX_ERR_ARG;
  }
  p.nbuf_log2 = p.nbuf == 8 ? 3 : (p.nbuf == 4 ? 2 : 1);
  p.colsum = epi.colsum;
  p.csum_off = (int)((size_t)p.a_slots * kSBox + (size_t)(p.resident ? kb_per_tile : p.b_slots) * b_tile +
                     (size_t)p.nbuf * kEpiBox);
  const size_t smem = (size_t)p.csum_off + csum_bytes + 1024;
  if (smem > (size_t)kHaloSmem + 1024) return DBX_ERR_ARG;
  p.idesc = umma_idesc_bf16(128, p.block_n, 0, 0);
  p.tmem_cols = 2 * p.block_n <= 128 ? 128 : 256;
  p.bias = epi.bias; p.relu = epi.relu; p.aux_mode = epi.aux_mode;

  CUtensorMap tmA, tmB, tmO, tmX;
  int rc = encode_box(&tmA, x, 8, 18);
  if (rc) return rc;
  rc = encode_mat_map(&tmB, wk, out.C, 9 * x.C, p.block_n);
  if (rc) return rc;
  rc = encode_box(&tmO, out, 8, 16);
  if (rc) return rc;
  if (epi.aux_mode) {
    Act ax = out;
    ax.ptr = const_cast<void*>(epi.aux); ax.cs = epi.aux_cs; ax.coff = epi.aux_coff;
    rc = encode_box(&tmX, ax, 8, 16);
    if (rc) return rc;
  } else {
    tmX = tmO;
  }
  if (epi.aux_mode != 0 && epi.aux_mode != 1 && epi.aux_mode != 2) return DBX_ERR_ARG;
  typedef void (*HaloFn)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const HaloParams);
  static const HaloFn fns[2] = {conv3x3_halo_kernel<false>, conv3x3_halo_kernel<true>};
  static SmemAttrOnce attr_once[2];
  const int fi = epi.aux_mode ? 1 : 0;
  const HaloFn fn = fns[fi];
  { const int arc = set_max_smem_once((const void*)fn, kHaloSmem + 1024, &attr_once[fi]); if (arc) return arc; }
  const int total = p.m_tiles * p.n_tiles;
  const int grid = total < tensor_sms() ? total : tensor_sms();
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(kHaloThreads); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr; cfg.numAttrs = 1;
  return (int)cudaLaunchKernelEx(&cfg, fn, tmA, tmB, tmO, tmX, p);
}

}  // namespace dbx
