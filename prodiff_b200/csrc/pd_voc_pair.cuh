// One ResBlock1 round of the NSF-HiFiGAN generator as ONE kernel (modules/nsf_hifigan/models.py:62-69):
//
//     xt = c2(leaky_relu(c1(leaky_relu(x))));   x = xt + x            (c1: kernel k, dilation d;  c2: kernel k, dilation 1)
//
// for the stages with 16 / 32 / 64 channels -- 128 to 512 positions per mel frame, where a convolution per launch
// (pd_voc.cuh) is bound by moving activations, not by arithmetic: every tap of every conv re-reads its input through L2 and
// the intermediate makes a round trip through HBM.  Here a CTA keeps a haloed tile of x resident in shared memory, both
// convolutions read their taps from shared memory, the intermediate never leaves the SM, and the residual comes from the
// same resident tile: per round the stage's tensor is read once and written once.
//
// * Shared-memory operand layout = the canonical NO-swizzle K-major form with the row pitch collapsed to 16 bytes:
//   [C / 8 planes][rows][8 channels].  Row r of plane p sits at p * plane_stride + r * 16, so "8 rows x 16 bytes" core
//   matrices are contiguous (SBO = 128) for ANY starting row: a convolution tap is the same tile with the descriptor's start
//   address moved by tap * dilation * 16 bytes -- no swizzle phase to respect, any dilation, any channel count >= 16.  The two
//   K-halves of one MMA (K = 16) are two adjacent planes (LBO = plane stride).
// * Weights are packed once in exactly that form, [tap][C / 8][C_out][8], conv1 then conv2: for C <= 32 both convolutions stay
//   resident in shared memory for the whole (persistent) CTA, fetched by one bulk copy; for C = 64 taps stream through a ring.
// * Per tile of NB * 128 conv1 rows: gather x (16-byte cp.async, fully contiguous in global memory; double-buffered, the
//   next tile's gather is in flight during this tile's arithmetic) -> conv1: NB x k x C/16 MMAs (M = 128, N = C) into TMEM
//   -> epilogue 1: + bias, LeakyReLU, zero outside the utterance (conv2's zero padding), fp16, back to shared memory in the
//   same plane layout -> conv2 on that tile -> epilogue 2: + bias + raw x (from the resident tile) [+ running ResBlock sum]
//   -> scale / LeakyReLU -> fp16 rows to global memory (64 channels: through a warp-local staging pass, so every store
//   instruction writes full 128-byte rows).  A tile yields NB * 128 - (k - 1) positions.
// * The chain is serial inside a CTA; several CTAs per SM (2 - 4: 30 - 110 KB of shared memory, 64 - 128 TMEM columns each)
//   overlap one CTA's loads and epilogues with another's MMAs.  With a double-buffered x tile (16 / 32 channels) conv1 of tile
//   i + 1 is issued right behind conv2 of tile i into its own TMEM columns, so it runs under epilogue 2.
// * Warps 0 .. EW-1: gather + epilogues (thread = TMEM lane = a row of a 128-row block; EW = 8: one block per warp group);
//   warp EW: MMA issue (whole warp, uniform operands, one elected lane -- see umma_f16_lohi); warp EW + 1: weight fetch.
// * kSingle: the same machinery for ONE convolution (the rate-2 transposed convs of the narrow stages).
// * lengths: per-utterance ends of a padded batch -- both the intermediate and the result are zero past them.
#pragma once
#include "pd_voc.cuh"

namespace pd {

struct VocPairArgs {
  const __half* x;         // guarded input [B][G + T + G][C], first element; stored LeakyReLU-activated
  const __half* add;       // optional raw addend, same geometry (running sum over the ResBlocks); may alias out
  __half* out;             // output, same geometry
  float* out32;            // single-conv form only: when set, column 0 alone leaves as fp32 at out32[b * T + position] (conv_post: the waveform)
  const __half* w;         // packed weights: conv1 [k][C/8][C][8] then conv2 [k][C/8][C][8]
  const float* b1;
  const float* b2;
  long long bstride;       // (T + 2G) * C
  long long out_bstride;   // utterance stride of out / add (= bstride for a ResBlock round) and the element offset of their
  long long out_goff;      // position 0 (= G * C); the rate-2 transposed conv writes T rows of C = 2 positions x C/2 channels
  int B, T, G;
  int k, dil;
  int tiles_per_row;       // ceil(T / P), P = NB * 128 - (k - 1)
  float slope1;            // LeakyReLU between the two convolutions
  float in_inv_slope;      // raw x = x >= 0 ? x : x * in_inv_slope
  float scale;             // applied after the sums
  int act;                 // 0 none, 1 LeakyReLU(slope), 2 tanh
  float slope;
  uint32_t idesc;
  const int* lengths;      // optional [B] valid mel frames per utterance: positions >= lengths[b] * len_mul are zeros in both the
  int len_mul;             // intermediate and the result (ragged batch = every utterance computed as if alone)
  long long* trace;        // debug: CTA 0 stamps clock64() at [tile iteration * 16 + slot] for its first 64 tiles (pd_debug_set_trace)
};

#define PD_PAIR_TRACE(slot) do { if (a.trace != nullptr && blockIdx.x == 0 && it < 64u) a.trace[it * 16u + (slot)] = clock64(); } while (0)
#ifndef PD_PAIR_SLEEP_NS
#define PD_PAIR_SLEEP_NS 40   // back-off between polls of the per-tile barriers (a third of all issued instructions were polls)
#endif
constexpr int kPairRingMax = 10;   // most weight ring stages (C = 64): one tap each

// K-major operand without swizzle: start address, LBO (stride between the two 8-channel halves of a K = 16 step), SBO (stride
// between 8-row groups); descriptor version 1, layout type 0.
__device__ __forceinline__ uint64_t make_plane_desc(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
  const uint32_t lo = ((smem_addr >> 4) & 0x3FFFu) | (((lbo >> 4) & 0x3FFFu) << 16);
  const uint32_t hi = ((sbo >> 4) & 0x3FFFu) | (1u << 14);
  return (static_cast<uint64_t>(hi) << 32) | lo;
}
// tcgen05.mma with the descriptors given as (low, high) words.  All lanes of a converged warp execute the call with uniform
// operands and one elected lane issues: straight-line uniform-register code instead of the per-lane operand broadcast loop
// the compiler emits around a tcgen05.mma inside an `if (lane == 0)` branch (~60 cycles per MMA, measured -- more than a
// 128 x 16 x 16 MMA takes).
__device__ __forceinline__ void umma_f16_lohi(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t.reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}" ::"r"(d_tmem),
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
// tcgen05.commit by the same elected lane (all lanes of the converged warp execute the call)
__device__ __forceinline__ void umma_commit_elect(uint64_t* bar) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(smem_u32(bar))
      : "memory");
}
// plain bulk copy global -> shared, completion on an mbarrier
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
               "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// host + device: shared-memory plan of one launch
struct VocPairSmem {
  uint32_t xr, yr;        // allocated rows of the x tile / of the intermediate tile (odd: a plane stride of an odd number of
                          // 16-byte units spreads the 8 planes over all banks for accesses that walk planes first)
  uint32_t x_off[2], y_off, w_off, bar_off, total;
  uint32_t ring;          // weight ring stages (streamed weights)
};
template <int C, int NB, bool kRes, int kXB, int kMinCtas, bool kSingle = false>
__host__ __device__ inline VocPairSmem voc_pair_smem(int k, int dil) {
  constexpr uint32_t kBudget = (227u * 1024u) / (uint32_t)kMinCtas - 1024u;   // leaves room for kMinCtas CTAs per SM
  VocPairSmem s;
  const int h = (k - 1) / 2;
  s.xr = (uint32_t)(NB * 128 + 2 * h * dil) | 1u;
  s.yr = (uint32_t)(NB * 128 + 2 * h) | 1u;
  uint32_t o = 0;
  for (int i = 0; i < 2; ++i) {
    s.x_off[i] = o;
    if (i < kXB) o += s.xr * C * 2;
  }
  s.y_off = o;
  o += (kSingle && C < 64) ? 0u : s.yr * C * 2;   // single conv: only the 64-channel store staging uses it
  s.w_off = o;
  s.ring = 0;
  if (kRes) {
    o += (uint32_t)((kSingle ? 1 : 2) * k * C * C * 2);
  } else {   // as many one-tap stages as fit beside the other CTAs on the SM (2 .. kPairRingMax)
    const uint32_t tap = C * C * 2, fixed = o + 256u + 128u;
    uint32_t n = fixed + 2 * tap < kBudget ? (kBudget - fixed) / tap : 2u;
    if (n > (uint32_t)kPairRingMax) n = kPairRingMax;
    if (n > (uint32_t)(2 * k)) n = 2 * k;
    s.ring = n;
    o += n * tap;
  }
  s.bar_off = o;
  s.total = o + 256 + 128;   // barriers + slack for the 128-byte alignment of the base
  return s;
}

// EW = epilogue warps: 4 (each thread walks the NB row blocks of its TMEM lane quadrant) or 4 * NB (one row block per warp group)
// kSingle: ONE convolution per tile (conv1's weights, resident, single-buffered x), epilogue = + bias [+ add] -> scale / LeakyReLU
// -> store, no residual: the rate-2 transposed convolutions of the narrow stages (a 3-tap conv whose output row of C values is two
// positions of the up-sampled tensor) read their input once from a resident tile instead of gathering every tap through L2.
template <int C, int NB, bool kRes, int kXB, int EW, int kMinCtas, bool kSingle = false>
__global__ void __launch_bounds__((EW + 2) * 32, kMinCtas) voc_pair_kernel(const VocPairArgs a) {
  static_assert(!kSingle || (kRes && kXB == 1), "single-conv form: resident weights, serial tiles");
  constexpr int PL = C / 8;                      // planes = 16-byte chunks per row
  constexpr int CW = C < 32 ? C : 32;            // accumulator columns per tcgen05.ld
  constexpr int NT = EW * 32;                    // gather / epilogue threads
  constexpr int MSTEP = EW / 4;                  // row blocks are dealt round-robin to the warp groups
  constexpr bool kAhead = kXB == 2;              // conv1 of tile i+1 is issued behind conv2 of tile i (own accumulator columns)
  constexpr uint32_t kAcc2 = kAhead ? NB * C : 0;   // column offset of the conv2 accumulators
  constexpr uint32_t kCols = (kAhead ? 2 : 1) * NB * C < 32 ? 32 : (kAhead ? 2 : 1) * NB * C;
  static_assert((kCols & (kCols - 1)) == 0 && kCols <= 512, "TMEM columns");
  static_assert(EW == 4 || EW == 4 * NB, "epilogue warps");
  constexpr uint32_t kTapBytes = C * C * 2;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((128u - (raw_addr & 127u)) & 127u);
  const VocPairSmem L = voc_pair_smem<C, NB, kRes, kXB, kMinCtas, kSingle>(a.k, a.dil);
  const uint32_t XPS = L.xr * 16, YPS = L.yr * 16;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.bar_off);
  uint64_t* xfull = bars;             // [2]   EW arrivals (one per warp): a gathered x tile is visible to the tensor core
  uint64_t* acc1 = bars + 2;          //       conv1 accumulators complete
  uint64_t* yfull = bars + 3;         //       EW arrivals: the intermediate tile is in shared memory
  uint64_t* acc2 = bars + 4;          //       conv2 accumulators complete
  uint64_t* wfull = bars + 5;         // [kPairRingMax] (resident: [0] only)
  uint64_t* wempty = bars + 5 + kPairRingMax;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5 + 2 * kPairRingMax);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(&xfull[0], EW);   // one arrival per warp (after its lanes' proxy fences): an arrive is a shared-memory atomic, and 128
    mbar_init(&xfull[1], EW);   // same-address arrivals per barrier per tile from 4 CTAs kept the SM's atomic unit busy
    mbar_init(acc1, 1);
    mbar_init(yfull, EW);
    mbar_init(acc2, 1);
    for (int i = 0; i < kPairRingMax; ++i) {
      mbar_init(&wfull[i], 1);
      mbar_init(&wempty[i], 1);
    }
    fence_mbar_init();
  }
  if (warp == EW) tmem_alloc<kCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  griddep_launch();

  const int k = a.k, hd = (k - 1) / 2 * a.dil;
  const int h = kSingle ? 0 : (k - 1) / 2;          // rows a tile gives up to the second convolution's reach
  const int P = NB * 128 - 2 * h;
  const int num_tiles = a.B * a.tiles_per_row;
  const int rows_total = a.T + 2 * a.G;

  if (warp == EW + 1) {
    // ===================== weights =====================
    if (lane == 0) {
      if constexpr (kRes) {
        mbar_arrive_expect_tx(&wfull[0], (kSingle ? 1u : 2u) * k * kTapBytes);
        bulk_load(smem + L.w_off, a.w, (kSingle ? 1u : 2u) * k * kTapBytes, &wfull[0]);
      } else {
        uint32_t stage = 0, parity = 1;   // first pass: the stages start out empty
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x)
          for (int tp = 0; tp < 2 * k; ++tp) {
            mbar_wait(&wempty[stage], parity, 21);
            mbar_arrive_expect_tx(&wfull[stage], kTapBytes);
            bulk_load(smem + L.w_off + stage * kTapBytes, reinterpret_cast<const uint8_t*>(a.w) + (size_t)tp * kTapBytes, kTapBytes, &wfull[stage]);
            if (++stage == L.ring) { stage = 0; parity ^= 1u; }
          }
      }
    }
  } else if (warp == EW) {
    // ===================== MMA issue =====================
    // The whole warp runs this loop with warp-uniform values and one elected lane issues the tcgen05 instructions (inside the
    // asm), so the descriptors live in uniform registers and cost one add per MMA.  Descriptor low words count in 16-byte
    // units = tile rows: start address [0,14), LBO [16,30).
    constexpr uint32_t kDescHi = (128u >> 4) | (1u << 14);                               // SBO = 128 bytes, version 1, no swizzle
    const uint32_t sbase = __shfl_sync(0xffffffffu, (raw_addr + 127u) & ~127u, 0) >> 4;  // 32-bit shared address of `smem`, in 16-byte units
    const uint32_t x_lo0 = (sbase + (L.x_off[0] >> 4)) | (L.xr << 16), x_lo1 = (sbase + (L.x_off[1] >> 4)) | (L.xr << 16);
    const uint32_t y_lo = (sbase + (L.y_off >> 4)) | (L.yr << 16);
    const uint32_t w_lo = (sbase + (L.w_off >> 4)) | ((uint32_t)C << 16);
    constexpr uint32_t kTapUnits = kTapBytes / 16;
    const uint32_t tm = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t dil_u = (uint32_t)a.dil, idesc = a.idesc;
    uint32_t it = 0, wstage = 0, wparity = 0;
    // one convolution: k taps x NB row blocks x C/16 K steps; a_lo walks one tap (a_step rows) per iteration
    auto issue_conv = [&](uint32_t d_tm, uint32_t a_lo, const uint32_t a_ps, const uint32_t a_step, uint32_t b_lo) {
#pragma unroll 1
      for (int tp = 0; tp < k; ++tp) {
        uint32_t bt = b_lo;
        if constexpr (!kRes) {
          mbar_wait(&wfull[wstage], wparity, 25);
          tc_fence_after();
          bt = w_lo + wstage * kTapUnits;
        }
        const uint32_t acc = tp != 0 ? 1u : 0u;
#pragma unroll
        for (int m = 0; m < NB; ++m)
#pragma unroll
          for (int s = 0; s < C / 16; ++s)
            umma_f16_lohi(d_tm + (uint32_t)(m * C), a_lo + (uint32_t)(2 * s) * a_ps + (uint32_t)(m * 128), kDescHi, bt + (uint32_t)(2 * s * C), kDescHi, idesc,
                          s != 0 ? 1u : acc);
        if constexpr (!kRes) {
          umma_commit_elect(&wempty[wstage]);
          if (++wstage == L.ring) { wstage = 0; wparity ^= 1u; }
        } else {
          b_lo += kTapUnits;
        }
        a_lo += a_step;
      }
    };
    auto conv1 = [&](uint32_t itx) {   // tile iteration itx: wait for its x tile, issue, commit
      const uint32_t buf = kXB == 2 ? (itx & 1u) : 0u;
      mbar_wait(&xfull[buf], kXB == 2 ? ((itx >> 1) & 1u) : (itx & 1u), 23, PD_PAIR_SLEEP_NS);
      tc_fence_after();
      issue_conv(tm, buf != 0u ? x_lo1 : x_lo0, L.xr, dil_u, w_lo);
      umma_commit_elect(acc1);
      __syncwarp();
    };
    if constexpr (kRes) mbar_wait(&wfull[0], 0, 22);
    if (kAhead && (int)blockIdx.x < num_tiles) conv1(0);
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      if (lane == 0) PD_PAIR_TRACE(8);
      if (!kAhead) conv1(it);
      if (lane == 0) PD_PAIR_TRACE(9);
      if constexpr (!kSingle) {
        mbar_wait(yfull, it & 1u, 24, PD_PAIR_SLEEP_NS);
        tc_fence_after();
        if (lane == 0) PD_PAIR_TRACE(10);
        issue_conv(tm + kAcc2, y_lo, L.yr, 1u, w_lo + (kRes ? (uint32_t)k * kTapUnits : 0u));
        umma_commit_elect(acc2);
        if (lane == 0) PD_PAIR_TRACE(11);
        __syncwarp();
      }
      if (kAhead && tile + (int)gridDim.x < num_tiles) conv1(it + 1);   // overlaps this tile's second epilogue
    }
  } else {
    // ===================== gather + epilogues: a thread owns one row of a 128-row block (= its TMEM lane) =====================
    const int tid = threadIdx.x;
    const int q = warp & 3, m0 = MSTEP == 1 ? 0 : (warp >> 2);
    const int trow = q * 32 + lane;                                 // row inside a block
    griddep_wait();
    // x tile row i <-> buffer row row0 + i; chunk c = i * PL + p is 16 bytes at tile_src + c * 16: thread tid takes chunks
    // tid, tid + NT, ... -- same plane p every time, NT / PL rows further down
    const int n_chunks = (NB * 128 + 2 * hd) * PL;
    const uint32_t g_dst_off = (uint32_t)(tid % PL) * XPS + (uint32_t)(tid / PL) * 16u;
    auto gather = [&](int tile, uint32_t buf) {
      const int b = tile / a.tiles_per_row;
      const int p0 = (tile - b * a.tiles_per_row) * P;
      const int row0 = a.G + p0 - h - hd;                           // buffer row of tile row 0
      const __half* src = a.x + (long long)b * a.bstride + (long long)row0 * C + tid * 8;
      uint32_t dst = smem_u32(smem + L.x_off[buf]) + g_dst_off;
      if (row0 >= 0 && row0 + NB * 128 + 2 * hd <= rows_total) {
        for (int c = tid; c < n_chunks; c += NT, src += NT * 8, dst += (NT / PL) * 16)
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
      } else {                                                      // a tile that reaches past the buffer: those rows are zeros
        int row = row0 + tid / PL;
        for (int c = tid; c < n_chunks; c += NT, src += NT * 8, dst += (NT / PL) * 16, row += NT / PL) {
          const bool ok = row >= 0 && row < rows_total;
          voc_cp_async16(dst, ok ? (const void*)src : (const void*)a.x, ok);
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    const uint32_t tlane = tmem_base + ((uint32_t)(q * 32) << 16);
    const bool has_add = a.add != nullptr;
    uint32_t it = 0;
    auto publish_x = [&](uint32_t buf) {   // this thread's share of a gathered tile has landed: make it visible to the tensor core
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&xfull[buf]);
    };
    if (kAhead && (int)blockIdx.x < num_tiles) {
      gather(blockIdx.x, 0);
      publish_x(0);
      if ((int)(blockIdx.x + gridDim.x) < num_tiles) gather(blockIdx.x + gridDim.x, 1);
    }
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const uint32_t buf = kXB == 2 ? (it & 1u) : 0u;
      if (!kAhead) {
        gather(tile, 0);
        publish_x(0);
      }
      if (tid == 0) PD_PAIR_TRACE(0);
      if (tid == 0) PD_PAIR_TRACE(1);
      const int b = tile / a.tiles_per_row;
      const int p0 = (tile - b * a.tiles_per_row) * P;
      const int Tb = a.lengths != nullptr ? min(a.T, a.lengths[b] * a.len_mul) : a.T;   // this utterance's own end
      // ---- epilogue 1: intermediate tile row i <-> position p0 - h + i; rows outside the utterance are conv2's zero padding
      if constexpr (!kSingle) {
        mbar_wait(acc1, it & 1u, 26, PD_PAIR_SLEEP_NS);
        tc_fence_after();
        if (tid == 0) PD_PAIR_TRACE(2);
      }
      if constexpr (!kSingle) {
        uint8_t* yb = smem + L.y_off;
        const bool interior = p0 - h >= 0 && p0 - h + NB * 128 <= Tb;   // tile-uniform: no row needs zeroing
#pragma unroll
        for (int m = m0; m < NB; m += MSTEP) {
          const int i = m * 128 + trow;
          const int pos = p0 - h + i;
          const float keep = (interior || (pos >= 0 && pos < Tb)) ? 1.f : 0.f;
#pragma unroll
          for (int cc = 0; cc < C / CW; ++cc) {
            uint32_t r[CW];
            if constexpr (CW == 32) tmem_ld_x32(tlane + (uint32_t)(m * C + cc * CW), r);
            else tmem_ld_x16(tlane + (uint32_t)(m * C + cc * CW), r);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < CW; j += 8) {
              const float4 ba = __ldg(reinterpret_cast<const float4*>(a.b1 + cc * CW + j));
              const float4 bb = __ldg(reinterpret_cast<const float4*>(a.b1 + cc * CW + j + 4));
              const float bv[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
              float v[8];
#pragma unroll
              for (int u = 0; u < 8; ++u) {
                const float t = __uint_as_float(r[j + u]) + bv[u];
                v[u] = fmaxf(t, t * a.slope1);                      // LeakyReLU, 0 <= slope <= 1
              }
              if (!interior) {
#pragma unroll
                for (int u = 0; u < 8; ++u) v[u] *= keep;
              }
              *reinterpret_cast<uint4*>(yb + (uint32_t)(cc * CW / 8 + j / 8) * YPS + (uint32_t)i * 16u) =
                  make_uint4(Cvt<__half>::pack(v[0], v[1]), Cvt<__half>::pack(v[2], v[3]), Cvt<__half>::pack(v[4], v[5]), Cvt<__half>::pack(v[6], v[7]));
            }
          }
        }
      }
      // the addend (source branch / running ResBlock sum) does not depend on the accumulator: fetch this thread's rows of it before
      // the wait (16 channels and the single-conv form; at 32 / 64 channels the extra registers spill and the round gets slower: measured)
      constexpr bool kPreAdd = kSingle || C == 16;
      uint4 pre_add[kPreAdd ? (NB + MSTEP - 1) / MSTEP : 1][kPreAdd ? C / 8 : 1];
      if constexpr (kPreAdd) {
        if (has_add) {
#pragma unroll
          for (int m = m0, mi = 0; m < NB; m += MSTEP, ++mi) {
            const int pos = p0 + m * 128 + trow;
            const bool live = pos < a.T;
            const uint4* ap = reinterpret_cast<const uint4*>(a.add + (long long)b * a.out_bstride + a.out_goff + (long long)pos * C);
#pragma unroll
            for (int j = 0; j < C / 8; ++j) pre_add[mi][j] = live ? ap[j] : make_uint4(0u, 0u, 0u, 0u);
          }
        }
      }
      if constexpr (!kSingle) {
        if (tid == 0) PD_PAIR_TRACE(3);
        fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(yfull);
        if (kAhead && tile + (int)gridDim.x < num_tiles) publish_x(buf ^ 1u);   // the next tile, gathered during the previous iteration
        if (tid == 0) PD_PAIR_TRACE(4);
        // ---- epilogue 2: output row r <-> position p0 + r; raw x from tile row r + h + hd
        mbar_wait(acc2, it & 1u, 27, PD_PAIR_SLEEP_NS);
        tc_fence_after();
        if (tid == 0) PD_PAIR_TRACE(5);
      } else {
        mbar_wait(acc1, it & 1u, 26, PD_PAIR_SLEEP_NS);
        tc_fence_after();
      }
      {
        const uint8_t* xb = smem + L.x_off[buf];
        uint8_t* yb = smem + L.y_off;                               // free again: staging for the coalesced store (C >= 64)
#pragma unroll
        for (int m = m0; m < NB; m += MSTEP) {
          const int r0 = m * 128 + trow;
          const int pos = p0 + r0;
          const bool live = r0 < P && pos < a.T;
          const long long goff = (long long)b * a.out_bstride + a.out_goff + (long long)pos * C;   // out / add geometry
#pragma unroll
          for (int cc = 0; cc < C / CW; ++cc) {
            uint32_t r[CW];
            if constexpr (CW == 32) tmem_ld_x32(tlane + kAcc2 + (uint32_t)(m * C + cc * CW), r);
            else tmem_ld_x16(tlane + kAcc2 + (uint32_t)(m * C + cc * CW), r);
            tmem_ld_wait();
            if (live || C >= 64) {
#pragma unroll
              for (int j = 0; j < CW; j += 8) {
                const float4 ba = __ldg(reinterpret_cast<const float4*>(a.b2 + cc * CW + j));
                const float4 bb = __ldg(reinterpret_cast<const float4*>(a.b2 + cc * CW + j + 4));
                const float bv[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
                const uint4 xr = *reinterpret_cast<const uint4*>(xb + (uint32_t)(cc * CW / 8 + j / 8) * XPS + (uint32_t)(r0 + h + hd) * 16u);
                const uint32_t xw[4] = {xr.x, xr.y, xr.z, xr.w};
                float v[8];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                  float2 f = Cvt<__half>::unpack(xw[u]);
                  f.x = kSingle ? 0.f : fminf(f.x, f.x * a.in_inv_slope);   // undo the stored LeakyReLU, inv_slope >= 1
                  f.y = kSingle ? 0.f : fminf(f.y, f.y * a.in_inv_slope);   // (single conv: no residual)
                  v[2 * u] = __uint_as_float(r[j + 2 * u]) + bv[2 * u] + f.x;
                  v[2 * u + 1] = __uint_as_float(r[j + 2 * u + 1]) + bv[2 * u + 1] + f.y;
                }
                if (has_add && live) {
                  const uint4 ar = kPreAdd ? pre_add[(m - m0) / MSTEP][(cc * CW + j) / 8] : *reinterpret_cast<const uint4*>(a.add + goff + cc * CW + j);
                  const uint32_t aw[4] = {ar.x, ar.y, ar.z, ar.w};
#pragma unroll
                  for (int u = 0; u < 4; ++u) {
                    const float2 f = Cvt<__half>::unpack(aw[u]);
                    v[2 * u] += f.x;
                    v[2 * u + 1] += f.y;
                  }
                }
                if (a.scale != 1.f) {
#pragma unroll
                  for (int u = 0; u < 8; ++u) v[u] *= a.scale;
                }
                if (a.act == 1) {
#pragma unroll
                  for (int u = 0; u < 8; ++u) v[u] = fmaxf(v[u], v[u] * a.slope);
                }
                if constexpr (kSingle) {
                  if (a.out32 != nullptr) {                         // conv_post: tanh of the one real output channel, fp32
                    if (cc == 0 && j == 0 && live) a.out32[(long long)b * a.T + pos] = pos < Tb ? tanhf(v[0]) : 0.f;
                    continue;
                  }
                }
                if (pos >= Tb) {                                    // ragged batch: past this utterance's end
#pragma unroll
                  for (int u = 0; u < 8; ++u) v[u] = 0.f;
                }
                const uint4 o = make_uint4(Cvt<__half>::pack(v[0], v[1]), Cvt<__half>::pack(v[2], v[3]), Cvt<__half>::pack(v[4], v[5]), Cvt<__half>::pack(v[6], v[7]));
                if constexpr (C >= 64) *reinterpret_cast<uint4*>(yb + (uint32_t)(cc * CW / 8 + j / 8) * YPS + (uint32_t)r0 * 16u) = o;
                else *reinterpret_cast<uint4*>(a.out + goff + cc * CW + j) = o;
              }
            }
          }
          if constexpr (C >= 64) {
            // the warp's 32 rows x C channels leave as full rows: lane -> (row = idx / PL, plane = idx % PL), 32 / PL rows per store
            __syncwarp();
            const int wr0 = m * 128 + q * 32;                       // first row of this warp's slab
            int n_live = P - wr0;                                   // live rows form a prefix of the slab
            if (a.T - (p0 + wr0) < n_live) n_live = a.T - (p0 + wr0);
            __half* gbase = a.out + (long long)b * a.out_bstride + a.out_goff + (long long)(p0 + wr0) * C;
#pragma unroll
            for (int t = 0; t < PL; ++t) {
              const int idx = lane + 32 * t, row = idx / PL, pl = idx % PL;
              if (row < n_live)
                *reinterpret_cast<uint4*>(gbase + row * C + pl * 8) = *reinterpret_cast<const uint4*>(yb + (uint32_t)pl * YPS + (uint32_t)(wr0 + row) * 16u);
            }
            __syncwarp();
          }
        }
      }
      if (tid == 0) PD_PAIR_TRACE(6);
      tc_fence_before();
      named_bar_sync(1, NT);   // every reader of this x tile / every TMEM drain is done before the buffers are reused
      if (kAhead && tile + 2 * (int)gridDim.x < num_tiles) gather(tile + 2 * gridDim.x, buf);
      if (tid == 0) PD_PAIR_TRACE(7);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == EW) {
    tc_fence_after();
    tmem_dealloc<kCols>(tmem_base);
  }
}

// Conv1d weights of one ResBlock round, [C, C, k] fp32 each -> the kernel's shared-memory image, fp16:
// conv1 [k][C/8][C_out][8] then conv2 [k][C/8][C_out][8]   (element (n, c, tap) of a conv at ((tap * C/8 + c/8) * C + n) * 8 + c % 8)
__global__ void k_voc_pack_pair(__half* __restrict__ dst, const float* __restrict__ w1, const float* __restrict__ w2, int C, int k) {
  const int per = k * C * C;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 2 * per; i += gridDim.x * blockDim.x) {
    const int conv = i / per, e = i - conv * per;
    const int c8 = e & 7, n = (e >> 3) % C, pk = (e >> 3) / C;   // pk = tap * (C / 8) + plane
    const int tap = pk / (C / 8), c = (pk - tap * (C / 8)) * 8 + c8;
    const float* w = conv == 0 ? w1 : w2;
    dst[i] = __float2half_rn(w[((long long)n * C + c) * k + tap]);
  }
}

// One Conv1d weight [Cout, C, k] fp32 (Cout <= C) -> the single-conv image [k][C / 8][C rows][8], rows >= Cout zero
__global__ void k_voc_pack_single(__half* __restrict__ dst, const float* __restrict__ w, int C, int Cout, int k) {
  const int total = k * C * C;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int c8 = i & 7, n = (i >> 3) % C, pk = (i >> 3) / C;
    const int tap = pk / (C / 8), c = (pk - tap * (C / 8)) * 8 + c8;
    dst[i] = __float2half_rn(n < Cout ? w[((long long)n * C + c) * k + tap] : 0.f);
  }
}
// ConvTranspose1d weight [Cin, Cout, 2s] (stride s, padding s / 2) as the equivalent 3-tap convolution with N = s * Cout rows (see
// k_voc_pack_convT), in the resident-tile image [tap j][Cin / 8][N][8]: row (phase r, co), tap j = input position q - 1 + j.
__global__ void k_voc_pack_convT_planes(__half* __restrict__ dst, const float* __restrict__ src, int Cin, int Cout, int s) {
  const int N = s * Cout, k = 2 * s;
  const int total = 3 * Cin * N;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int c8 = i & 7, n = (i >> 3) % N, pk = (i >> 3) / N;   // pk = tap * (Cin / 8) + plane
    const int j = pk / (Cin / 8), c = (pk - j * (Cin / 8)) * 8 + c8;
    const int r = n / Cout, co = n - r * Cout;
    const int kk = r + s / 2 - (j - 1) * s;
    dst[i] = __float2half_rn((kk >= 0 && kk < k) ? src[((long long)c * Cout + co) * k + kk] : 0.f);
  }
}

}  // namespace pd
