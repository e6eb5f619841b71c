// densebox_b200 — the TMA-staged epilogue shared by conv_fprop_kernel and conv3x3_halo_kernel.
//
// Eight epilogue warps (warp w: TMEM lane quarter w & 3, 32-column half (w - 2) >> 2) drain one 128-row x 64-column
// sub-block ("sb") of the accumulator at a time:
//     tcgen05.ld.x32 -> + bias (fp32) -> bf16x2 pack -> ReLU / ReLU-mask / dropout IN THE PACKED DOMAIN
//     -> swizzled staging box in smem -> (one named barrier) -> TMA store by the leader thread.
// The ReLU-backward / dropout-scale mask tile is TMA-prefetched into the SAME staging box D sub-blocks ahead and
// transformed in place, so no epilogue thread ever waits on a global load.
//
// What the round-1 ncu source pages said about the previous epilogue (profiles/README.md, session 2) and what this
// version does about it:
//   * ~865 SASS instructions per thread and sub-block on the dropout path (heads), ~2 650 cycles per sub-block;
//     per-element bit tests, integer divisions by the run-time ring depth, one Philox call per 32 channels.
//     -> mask / ReLU / dropout act on packed bf16x2 words (HMNMX2, HMUL2, one LOP3 per pair), the ring depth is a
//        power of two, all tile-invariant index arithmetic is hoisted, one Philox call serves 128 channels.
//   * two 256-thread named barriers per sub-block, the leader's cp.async.bulk.wait_group between them.
//     -> ONE barrier per sub-block: the leader proves the NEXT box free (and queues its mask load) before it
//        arrives, and issues the store after.
//   * mbarrier.arrive.release.cluster on the TMEM-empty barrier compiled to MEMBAR.ALL.CTA + ERRBAR (10 % of the
//     samples of the head GEMMs).  -> .relaxed: the only thing ordered is TMEM reads, which tcgen05.wait::ld and
//     tcgen05.fence::before_thread_sync already cover.
#pragma once
#include "dbx_ptx.cuh"

namespace dbx {

static constexpr int kEpiBox = 16384;  // one staging box: 128 rows x 64 bf16 channels

struct EpiArgs {
  const float* bias; int cout;
  int relu, aux_mode;
  int block_n, nsb, nbuf_log2;
  int tw, th, box_rows;                 // pixel box of one tile (rows = tw*th*tn <= 128)
  int out_W, out_H;                     // aux_mode 3: element index of the dropped activation
  const unsigned long long* rng; int rng_channels;
  float* csum = nullptr;                // smem fp32 [8][cout] (one copy per row group): running column sums of
                                        // everything this CTA stored, or null
  bf16* pool_out = nullptr;             // fused 2x2 max-pool: pooled NHWC view (see ConvEpilogue::pool_out); 8 x 16 tiles
  int pool_cs = 0, pool_coff = 0;
  unsigned short* pool_idx = nullptr;   // 2-bit arg-max map (u16 per pooled pixel and 8 channels) or null
  int pool_keep_full = 0;               // also stage + TMA-store the full-resolution tile
};

__device__ __forceinline__ uint32_t bf162_as_u32(__nv_bfloat162 h) { return *reinterpret_cast<uint32_t*>(&h); }
__device__ __forceinline__ __nv_bfloat162 u32_as_bf162(uint32_t u) { return *reinterpret_cast<__nv_bfloat162*>(&u); }

__device__ __forceinline__ void mbar_arrive_cluster_relaxed(uint64_t* bar, uint32_t rank) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [ra];\n\t}" ::"r"(smem_u32(bar)),
      "r"(rank)
      : "memory");
}

// 16 packed bf16x2 words (32 consecutive channels of one row): apply ReLU / mask / dropout in place.
//   keep-bit i of `bits` belongs to channel i (same convention as dropout_bits32 / the dropout_mask kernel).
__device__ __forceinline__ void epi_relu16(uint32_t* pk) {
  const __nv_bfloat162 z = u32_as_bf162(0u);
#pragma unroll
  for (int i = 0; i < 16; ++i) pk[i] = bf162_as_u32(__hmax2(u32_as_bf162(pk[i]), z));
}
__device__ __forceinline__ uint32_t sel32(uint32_t g, uint32_t a, uint32_t b) { return (a & g) | (b & ~g); }  // g ? a : b, bitwise

// 2x2 max-pool of a warp's 4 x 8 pixel patch (lane = 8 * row + column) over the 32 channels each lane holds as 16
// packed bf16x2 words.  The four lanes of a window (lane ^ 1: next column, lane ^ 8: next row) split the channels:
// after a two-step butterfly the lane at window position (dx, dy) owns words 8 dx + 4 dy .. + 3 of the pooled pixel.
// First maximum in scan order wins ties (torch max_pool2d), arg = 2 dy + dx as in maxpool2x2_fwd_kernel.
// Returns the 8 x 2 arg bits of the lane's eight channels.
__device__ __forceinline__ uint32_t epi_pool2x2(const uint32_t* pk, int lane, uint32_t* m) {
  const bool dx = lane & 1, dy = (lane >> 3) & 1;
  const __nv_bfloat162* dummy = nullptr; (void)dummy;
  uint32_t r[8], a1[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const uint32_t own = dx ? pk[8 + k] : pk[k], snd = dx ? pk[k] : pk[8 + k];
    const uint32_t rcv = __shfl_xor_sync(0xffffffffu, snd, 1);
    const uint32_t a = dx ? rcv : own, b = dx ? own : rcv;          // the column-0 / column-1 value of the window
    const uint32_t g = __hgt2_mask(u32_as_bf162(b), u32_as_bf162(a));
    r[k] = sel32(g, b, a);
    a1[k] = g;
  }
  uint32_t bits = 0u;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const uint32_t own = dy ? r[4 + k] : r[k], snd = dy ? r[k] : r[4 + k];
    const uint32_t oa = dy ? a1[4 + k] : a1[k], sa = dy ? a1[k] : a1[4 + k];
    const uint32_t rcv = __shfl_xor_sync(0xffffffffu, snd, 8), ra = __shfl_xor_sync(0xffffffffu, sa, 8);
    const uint32_t top = dy ? rcv : own, bot = dy ? own : rcv, ta = dy ? ra : oa, ba = dy ? oa : ra;
    const uint32_t g = __hgt2_mask(u32_as_bf162(bot), u32_as_bf162(top));
    m[k] = sel32(g, bot, top);
    const uint32_t arg = (sel32(g, ba, ta) & 0x00010001u) | ((g & 0x00010001u) << 1);   // 2 bits per 16-bit half
    bits |= ((arg & 3u) << (4 * k)) | (((arg >> 16) & 3u) << (4 * k + 2));
  }
  return bits;
}

__device__ __forceinline__ void epi_dropout16(uint32_t* pk, uint32_t bits) {
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    // {0, 2.0} per half word: bit 2i -> 0x4000, bit 2i+1 -> 0x40000000 (bf16 2.0 = 0x4000); x2 and x0 are exact
    const uint32_t a = (14 - 2 * i) >= 0 ? (bits << ((14 - 2 * i) & 31)) : (bits >> ((2 * i - 14) & 31));
    const uint32_t b = (29 - 2 * i) >= 0 ? (bits << ((29 - 2 * i) & 31)) : (bits >> ((2 * i - 29) & 31));
    const uint32_t m = (a & 0x00004000u) | (b & 0x40000000u);
    pk[i] = bf162_as_u32(__hmul2(u32_as_bf162(pk[i]), u32_as_bf162(m)));
  }
}

// tile_of(it, nt, w0, h0, n0): CTA-local tile iteration -> output-channel tile and pixel-box origin.
// kEpi selects the epilogue flavour at compile time (kEpiPlain: bias / ReLU only; kEpiMask: ReLU-backward or dropout
// mask tile prefetched by TMA, e.aux_mode 1 / 2; kEpiPhilox: dropout drawn in place, e.aux_mode 3; kEpiPool: fused
// 2x2 max-pool).  The short-K layers (conv1_1, conv1_2, conv2_1, the heads) are paced by this code, two warps per
// scheduler with every stall exposed: run-time tests of features a launch does not use (a parameter load, a compare
// and a branch each, ~15 per sub-block) measurably slow ALL of them — the unused max-pool-backward path alone cost
// 0.055 ms per step (profiles/README.md item 31).
enum { kEpiPlain = 0, kEpiMask = 1, kEpiPhilox = 2, kEpiPool = 3 };
template <bool kCta2, int kEpi, class TileFn>
__device__ __forceinline__ void epilogue_tma(const EpiArgs& e, const CUtensorMap* tmO, const CUtensorMap* tmX,
                                             uint8_t* ring, uint64_t* aux_bar, uint64_t* tfull_bar,
                                             uint64_t* tempty_bar, uint32_t tmem, int my_tiles, TileFn tile_of) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q4 = warp & 3, half = (warp - 2) >> 2;  // TMEM lane quarter, 32-column half of each 64-column block
  const int row = q4 * 32 + lane;
  const bool leader = threadIdx.x == 64;
  const int nsb = e.nsb, nbuf = 1 << e.nbuf_log2, nmask = nbuf - 1;
  constexpr bool aux_tma = kEpi == kEpiMask;
  const int D = aux_tma ? (nbuf >> 1) : 0;          // mask prefetch distance in sub-blocks
  const int total_sb = my_tiles * nsb;
  const uint32_t box_bytes = (uint32_t)e.box_rows * 128u;
  const bool live = row < e.box_rows;
  // tile-invariant pieces of this thread's addresses
  const int r_w = row % e.tw, r_h = (row / e.tw) % e.th, r_n = row / (e.tw * e.th);
  const uint32_t row_off = (uint32_t)row * 128u;
  const int sw = row & 7, cc0 = half * 4;           // first 16-byte chunk of this thread's 32 channels
  uint32_t chunk[4];
#pragma unroll
  for (int g = 0; g < 4; ++g) chunk[g] = row_off + (uint32_t)(((cc0 + g) ^ sw) << 4);

  // ---- mask prefetch head (leader only): walks the same (tile, sub-block) sequence D steps ahead
  int a_it = 0, a_j = 0, a_q = 0, a_nt = 0, a_w0 = 0, a_h0 = 0, a_n0 = 0;
  auto issue_aux_next = [&]() {
    if (a_j == 0) tile_of(a_it, a_nt, a_w0, a_h0, a_n0);
    const int b = a_q & nmask;
    mbar_arrive_expect_tx(&aux_bar[b], box_bytes);
    tma_load_4d(tmX, &aux_bar[b], ring + (size_t)b * kEpiBox, a_nt * e.block_n + a_j * 64, a_w0, a_h0, a_n0);
    ++a_q;
    if (++a_j == nsb) { a_j = 0; ++a_it; }
  };
  if (leader && aux_tma)
    for (int i = 0; i < D && a_q < total_sb; ++i) issue_aux_next();

  unsigned long long rng_seed = 0, rng_off = 0;
  if constexpr (kEpi == kEpiPhilox) { rng_seed = e.rng[0]; rng_off = e.rng[1]; }
  uint4 rnd = make_uint4(0u, 0u, 0u, 0u);
  unsigned long long rnd_ctr = ~0ull;

  uint32_t ring_s = smem_u32(ring);
  asm volatile("" : "+r"(ring_s));  // opaque: computed once, not re-derived per access
  int q = 0;
  for (int it = 0; it < my_tiles; ++it) {
    int nt, w0, h0, n0;
    tile_of(it, nt, w0, h0, n0);
    const int buf = it & 1; const uint32_t use = (uint32_t)(it >> 1);
    // element index of this row's channel 0 in the dropped activation (aux_mode 3)
    unsigned long long e_row = 0ull;
    if constexpr (kEpi == kEpiPhilox)
      e_row = (((unsigned long long)(n0 + r_n) * e.out_H + (h0 + r_h)) * e.out_W + (w0 + r_w)) *
              (unsigned long long)e.rng_channels;
    mbar_wait(&tfull_bar[buf], use & 1);
    tc_fence_after();
    const uint32_t taddr = tmem + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(buf * e.block_n);
    for (int j = 0; j < nsb; ++j, ++q) {
      uint8_t* sb = ring + (size_t)(q & nmask) * kEpiBox;
      const uint32_t sb_s = ring_s + (uint32_t)(q & nmask) * (uint32_t)kEpiBox;
      int ncols = e.block_n - j * 64; if (ncols > 64) ncols = 64;
      const int c0 = half * 32;
      const int ch = nt * e.block_n + j * 64 + c0;
      if (c0 + 32 <= ncols) {
        // ---- fast path: this warp's 32 columns in one TMEM load
        uint32_t v[32];
        tmem_ld_x32(taddr + j * 64 + c0, v);
        float4 b4[8];
        const bool vec_bias = e.bias && ch + 32 <= e.cout;
        if (vec_bias) {
          const float4* bp = reinterpret_cast<const float4*>(e.bias + ch);
#pragma unroll
          for (int i = 0; i < 8; ++i) b4[i] = __ldg(bp + i);
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            b4[i].x = (e.bias && ch + 4 * i + 0 < e.cout) ? __ldg(e.bias + ch + 4 * i + 0) : 0.f;
            b4[i].y = (e.bias && ch + 4 * i + 1 < e.cout) ? __ldg(e.bias + ch + 4 * i + 1) : 0.f;
            b4[i].z = (e.bias && ch + 4 * i + 2 < e.cout) ? __ldg(e.bias + ch + 4 * i + 2) : 0.f;
            b4[i].w = (e.bias && ch + 4 * i + 3 < e.cout) ? __ldg(e.bias + ch + 4 * i + 3) : 0.f;
          }
        }
        uint32_t bits = 0u;
        if constexpr (kEpi == kEpiPhilox) {  // one Philox call per 128 channels of a row
          const unsigned long long el = e_row + (unsigned long long)ch;
          const unsigned long long ctr = (el >> 7) + rng_off;
          if (ctr != rnd_ctr) {
            rnd_ctr = ctr;
            rnd = philox4x32(make_uint4((uint32_t)ctr, (uint32_t)(ctr >> 32), 0u, 0u),
                             make_uint2((uint32_t)rng_seed, (uint32_t)(rng_seed >> 32)));
          }
          const uint32_t wsel = (uint32_t)(el & 127ull) >> 5;
          bits = wsel == 0 ? rnd.x : (wsel == 1 ? rnd.y : (wsel == 2 ? rnd.z : rnd.w));
        }
        if (aux_tma) mbar_wait(&aux_bar[q & nmask], (uint32_t)((q >> e.nbuf_log2) & 1));
        tmem_ld_wait();
        if (live) {
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            pk[2 * i] = pack_bf16x2(__uint_as_float(v[4 * i]) + b4[i].x, __uint_as_float(v[4 * i + 1]) + b4[i].y);
            pk[2 * i + 1] = pack_bf16x2(__uint_as_float(v[4 * i + 2]) + b4[i].z, __uint_as_float(v[4 * i + 3]) + b4[i].w);
          }
          if (e.relu) epi_relu16(pk);
          if constexpr (kEpi == kEpiPhilox) {
            epi_dropout16(pk, bits);
          } else if constexpr (aux_tma) {
            const __nv_bfloat162 z = u32_as_bf162(0u);
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const uint4 a4 = ld_shared_v4(sb_s + chunk[g]);
              const uint32_t au[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                if (e.aux_mode == 1) pk[4 * g + i] &= __hgt2_mask(u32_as_bf162(au[i]), z);   // ReLU backward
                else pk[4 * g + i] = bf162_as_u32(__hmul2(u32_as_bf162(pk[4 * g + i]), u32_as_bf162(au[i])));
              }
            }
          }
          if constexpr (kEpi == kEpiPool) {
            // fused 2x2 max-pool: no staging box, no TMA store — the pooled quarter of each window goes straight out
            uint32_t m[4];
            const uint32_t bits = epi_pool2x2(pk, lane, m);
            const int px = w0 + r_w, py = h0 + r_h;
            if (px < e.out_W && py < e.out_H) {
              const int wb = ((lane & 1) ? 8 : 0) + (((lane >> 3) & 1) ? 4 : 0);
              const size_t pp = ((size_t)(n0 + r_n) * (e.out_H >> 1) + (py >> 1)) * (e.out_W >> 1) + (px >> 1);
              const int c = ch + 2 * wb;
              *reinterpret_cast<uint4*>(e.pool_out + pp * e.pool_cs + e.pool_coff + c) = make_uint4(m[0], m[1], m[2], m[3]);
              if (e.pool_idx) e.pool_idx[pp * (e.cout >> 3) + (c >> 3)] = (unsigned short)bits;
            }
          }
          if (kEpi != kEpiPool || e.pool_keep_full) {
#pragma unroll
          for (int g = 0; g < 4; ++g) st_shared_v4(sb_s + chunk[g], pk[4 * g], pk[4 * g + 1], pk[4 * g + 2], pk[4 * g + 3]);
          }
        }
      } else {
        // ---- ragged tail (block_n not a multiple of 64): 16 columns at a time, scalar guards
        if (aux_tma) mbar_wait(&aux_bar[q & nmask], (uint32_t)((q >> e.nbuf_log2) & 1));
        int cend = c0 + 32; if (cend > ncols) cend = ncols;
        for (int c = c0; c < cend; c += 16) {
          uint32_t v[16];
          tmem_ld_x16(taddr + j * 64 + c, v);
          tmem_ld_wait();
          const int chc = nt * e.block_n + j * 64 + c;
          if (live) {
            float f[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              f[i] = __uint_as_float(v[i]);
              if (e.bias && chc + i < e.cout) f[i] += __ldg(e.bias + chc + i);
              if (e.relu) f[i] = fmaxf(f[i], 0.f);
            }
            uint4* s0 = reinterpret_cast<uint4*>(sb + row_off + ((((c >> 3)) ^ sw) << 4));
            uint4* s1 = reinterpret_cast<uint4*>(sb + row_off + ((((c >> 3) + 1) ^ sw) << 4));
            if constexpr (kEpi == kEpiPhilox) {
              const uint32_t bits = dropout_bits16(e_row + (unsigned long long)chc, rng_seed, rng_off);
#pragma unroll
              for (int i = 0; i < 16; ++i) f[i] = ((bits >> i) & 1u) ? f[i] * 2.f : 0.f;
            } else if constexpr (aux_tma) {
              const uint4 a0 = *s0, a1 = *s1;
              const uint32_t au[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float lo = bf16lo(au[i]), hi = bf16hi(au[i]);
                if (e.aux_mode == 1) {
                  f[2 * i] = lo > 0.f ? f[2 * i] : 0.f;
                  f[2 * i + 1] = hi > 0.f ? f[2 * i + 1] : 0.f;
                } else {
                  f[2 * i] *= lo;
                  f[2 * i + 1] *= hi;
                }
              }
            }
            *s0 = make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]),
                             pack_bf16x2(f[6], f[7]));
            *s1 = make_uint4(pack_bf16x2(f[8], f[9]), pack_bf16x2(f[10], f[11]), pack_bf16x2(f[12], f[13]),
                             pack_bf16x2(f[14], f[15]));
          }
        }
      }
      if (kEpi == kEpiPool && !e.pool_keep_full) continue;  // nothing staged, nothing to store (whole 64-column blocks only)
      fence_proxy_async_smem();
      if (leader) {
        // Before the barrier of sub-block q the leader proves free the box that is written next: the box of
        // sub-block q+1 (no mask) or the box the mask of sub-block q+D is loaded into.  Stores are issued one per
        // sub-block, so "store (q + 1 + D' - nbuf) has read its box" = at most nbuf - 2 - D' + ... stores pending:
        //   no mask: nbuf - 2 (>= 0);   mask, D = nbuf/2: nbuf - D - 1.
        if (aux_tma) {
          if (nbuf == 8) bulk_wait_read<3>(); else if (nbuf == 4) bulk_wait_read<1>(); else bulk_wait_read<0>();
          if (a_q < total_sb) issue_aux_next();
        } else {
          if (nbuf == 8) bulk_wait_read<6>(); else if (nbuf == 4) bulk_wait_read<2>(); else bulk_wait_read<0>();
        }
      }
      named_bar_sync(1, 256);
      if (leader) {
        tma_store_4d(tmO, sb, nt * e.block_n + j * 64, w0, h0, n0);
        bulk_commit();
      }
      if (e.csum) {
        // Column sums of the finished box (bias gradient of the layer below).  Thread t: channel pair t & 31, rows
        // 16 (t >> 5) .. + 15; a warp reads the 32 words of ONE 128-byte row per instruction (conflict-free under
        // the swizzle).  Rows outside the image are exact zeros (zero-filled operands, no bias).  Each of the 8 row
        // groups owns a private copy csum[t >> 5][cout]: plain read-modify-write, no shared-memory float atomics
        // (those compile to CAS loops and cost +60 % on the short-K layers).  The box is not rewritten before every
        // thread has passed the next sub-block's barrier.
        const int t = (int)threadIdx.x - 64, cp = t & 31, rg = t >> 5, r0 = rg * 16;
        float s0 = 0.f, s1 = 0.f;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int rr = r0 + i;
          if (rr < e.box_rows) {
            const uint32_t w = *reinterpret_cast<const uint32_t*>(
                sb + (uint32_t)rr * 128u + (uint32_t)((((cp >> 2) ^ (rr & 7)) << 4) + ((cp & 3) << 2)));
            s0 += bf16lo(w); s1 += bf16hi(w);
          }
        }
        const int c = nt * e.block_n + j * 64 + 2 * cp;
        if (2 * cp < ncols && c < e.cout) {
          float2* a = reinterpret_cast<float2*>(e.csum + (size_t)rg * e.cout + c);
          float2 v = *a;
          v.x += s0; v.y += s1;
          *a = v;
        }
      }
    }
    tc_fence_before();
    __syncwarp();
    if (lane == 0) {
      if constexpr (kCta2) mbar_arrive_cluster_relaxed(&tempty_bar[buf], 0); else mbar_arrive(&tempty_bar[buf]);
    }
  }
  if (leader) bulk_wait_all();
}

}  // namespace dbx
