// Conv1d as an implicit GEMM on tcgen05 for the NSF-HiFiGAN generator (modules/nsf_hifigan/models.py:39-290):
// every convolution of the generator -- conv_pre, the transposed up-sampling convs, the 90 ResBlock convs with 16..256
// channels, kernel sizes 3/7/11 and dilations 1/3/5, conv_post -- is one instance of this kernel,
//
//   D[128 positions, BN out-channels] = sum over K blocks of  A_kb[128 positions, 64] * W_kb[BN, 64]^T      (fp32 in TMEM)
//
// * Activations are fp16, channels-last, one utterance after the other with GUARD rows of zeros in front of and behind
//   each utterance: [B][G + T + G][C].  The guard rows ARE the convolution's zero padding, so a tap that reaches outside
//   the utterance simply reads zeros from memory and every conv input is stored exactly once.
// * The A operand is read through a 4-D tensor map (channel, tap, position, utterance) with strides
//   (2, dil * C * 2, C * 2, (T + 2G) * C * 2) bytes -- overlapping on purpose: element (c, tap, t) is x[t + tap * dil][c].
//   A K block is 64 channels of one tap (box 64 x 1 x 128), landing as the usual 128-byte-swizzled K-major tile.
// * For C = 16 / 32 (the late stages: 256-512 samples per mel frame) a K block is all channels of 4 / 2 consecutive taps.
//   TMA cannot build that tile (with the 128-byte swizzle every box row shorter than 128 bytes is padded to a full
//   swizzle row -- measured, tools/voc_probe.py), so four extra producer warps gather it: one thread per position copies
//   its g pieces of C channels with 16-byte cp.async straight into the swizzled rows (chunk j of row r at (j ^ (r & 7))),
//   then -- a few K blocks later, so the copies stay in flight -- waits for its own copies, issues the generic -> async
//   proxy fence and arrives on the stage's mbarrier.  Consecutive lanes read consecutive positions, i.e. contiguous memory.  That keeps the narrow
//   stages on the tensor cores without padding their channels to 64; taps past the kernel size are zero-filled.
// * The weights are packed [N][K] with K = tap * C + c (taps padded to a multiple of g with zeros), 2-D map, (64, BN) box.
// * A transposed convolution with kernel 2s and stride s is a 3-tap convolution with N = s * C_out whose output row of
//   s * C_out values IS s consecutive rows of the up-sampled channels-last tensor (no shuffle kernel); taps a phase does
//   not use carry zero weights.
// * warp 0 = TMA producer, warp 1 = MMA issuer, warps 2..9 = epilogue: (acc + bias [+ residual] [+ addend]) * scale ->
//   LeakyReLU / tanh -> fp16 (or fp32 for the waveform), every global access as full row segments through a per-warp
//   shared-memory tile.  Tensors that feed a next convolution are stored ALREADY activated (LeakyReLU is invertible, so
//   the residual path recovers the raw value with one multiply): each activation tensor exists once in HBM.
#pragma once
#include "pd_gemm.cuh"

namespace pd {

struct VocShape {
  int B, T;              // utterances, positions per utterance (the conv's output length = its input length)
  int m_tiles_per_row;   // ceil(T / 128)
  int n_tiles;           // N_total / BN
  int cc_blocks;         // 64-channel chunks per tap group (C >= 64: C / 64, else 1)
  int tap_groups;        // ceil(taps / g)
  int g;                 // taps per K block (C >= 64: 1, else 64 / C)
  uint32_t idesc;
  long long* trace;      // debug (pd_debug_set_trace): CTA 0 stamps clock64() at [tile * 16 + slot] for its first 64 tiles -- 0 / 1 first / last
                         // TMA issue, 2 accumulator stage free, 3 / 4 first / last operand block landed, 5 cycles the issuing warp spent
                         // waiting for operand blocks, 6 epilogue sees the accumulator, 7 epilogue done
  // gather path (C < 64): the guarded activation buffer itself
  const __half* a_ptr;   // first element of the buffer (front guard of utterance 0)
  int a_C, a_taps, a_dil, a_G;
};

struct VocEpi {
  __half* out16;           // fp16 output, element (b, t, col) at out16[b * out_bstride + t * out_ld + col]
  float* out32;            // fp32 output of column 0 only (conv_post: the waveform), element (b, t) at out32[b * T + t]
  long long out_bstride;
  int out_ld;
  const float* bias;       // [N_total]
  const __half* res16;     // optional residual, stored LeakyReLU-activated: raw = r >= 0 ? r : r * res_inv_slope
  long long res_bstride;
  int res_ld;
  float res_inv_slope;
  const __half* add16;     // optional raw addend (harmonic source branch / running sum over the ResBlocks); may alias out16
  long long add_bstride;
  int add_ld;
  float scale;             // applied after the sums
  int act;                 // 0 none, 1 LeakyReLU(slope), 2 tanh
  float slope;
  const int* lengths;      // optional [B] valid mel frames per utterance (ragged batch): output rows >= lengths[b] * len_mul are
  int len_mul;             // stored as zeros, so every later convolution sees the zero padding a run of that utterance alone sees
};

constexpr int kVocGatherThreads = 128;   // one per position of a 128-row tile

// scale -> activation of one output value (act is launch-uniform: the compiler hoists the branch out of the unrolled loops);
// LeakyReLU as max(x, slope * x), 0 <= slope <= 1: two instructions instead of compare / multiply / select
__device__ __forceinline__ float voc_finish(float x, float scale, int act, float slope) {
  x *= scale;
  if (act == 1) return fmaxf(x, x * slope);
  if (act == 2) return tanhf(x);
  return x;
}

__device__ __forceinline__ void voc_cp_async16(uint32_t dst_smem, const void* src, bool pred) {
  const int n = pred ? 16 : 0;   // src-size 0: the 16 bytes are zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst_smem), "l"(src), "r"(n) : "memory");
}
// kPair: two CTAs of a cluster take two adjacent 128-position tiles as ONE tcgen05.mma.cta_group::2 (M = 256); each CTA stages its
// own A tile and HALF of every weight block, so an MMA fetches 4 KB + BN * 16 B of operands per SM instead of 4 KB + BN * 32 B --
// the wide convolutions (BN = 128 / 256) run at the shared-memory operand rate, not at the tensor rate.
template <int BN, bool kPair = false> struct VocCfg {
  static constexpr int kCW = BN >= 64 ? 32 : BN;                  // columns per epilogue chunk
  static constexpr int kRowWords = kCW / 2 + 4;                   // staging row: kCW halves + 16 bytes of padding
  static constexpr uint32_t kABytes = kBlockM * kBlockK * 2;
  static constexpr uint32_t kBBytes = (kPair ? BN / 2 : BN) * kBlockK * 2;
  static constexpr uint32_t kStageBytes = kABytes + kBBytes;
  static constexpr uint32_t kStagingBytes = kEpiWarps * 32 * kRowWords * 4;
  static constexpr int kStagesFit = (int)((224u * 1024u - kStagingBytes) / kStageBytes);   // 227 KB per CTA less barriers and alignment slack
  static constexpr int kStages = kStagesFit < 8 ? kStagesFit : 8;
  static constexpr uint32_t kSmemBytes = kStages * kStageBytes + kStagingBytes + 1024 + 256;
};

template <int BN, bool kGather, bool kPair = false>
__global__ void __launch_bounds__(kGemmThreads + (kGather ? kVocGatherThreads : 0), 1)
voc_conv_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const VocShape sh, const VocEpi ep) {
  using Cfg = VocCfg<BN, kPair>;
  static_assert(!(kPair && kGather) && (!kPair || BN >= 128), "pairs: TMA-fed wide tiles only");
  constexpr int kStages = Cfg::kStages;
  constexpr int CW = Cfg::kCW;
  constexpr int RW = Cfg::kRowWords;
  static_assert(BN == 16 || BN == 32 || BN == 64 || BN == 128 || BN == 256, "tile width");

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint32_t* staging = reinterpret_cast<uint32_t*>(smem + kStages * Cfg::kStageBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(staging) + Cfg::kStagingBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kStages;
  uint64_t* tfull_bar = bars + 2 * kStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = kPair ? cluster_ctarank() : 0u;
  const bool leader = rank == 0;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], kPair ? 2 : (kGather ? 1 + kVocGatherThreads : 1));   // B by TMA (+ one arrival per gathering thread)
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], (kPair ? 2 : 1) * (BN >= 64 ? kEpiWarps : kEpiWarps / 2));   // narrow tiles: one warp group per accumulator stage
    }
    fence_mbar_init();
  }
  if constexpr (kPair) {
    cluster_sync_all();   // both CTAs alive + barriers initialised before any remote arrive / paired allocation
    if (warp == 2) tmem_alloc_2sm<kTmemCols>(tmem_slot);
    tc_fence_before();
    cluster_sync_all();
  } else {
    if (warp == 2) tmem_alloc<kTmemCols>(tmem_slot);
    tc_fence_before();
    __syncthreads();
  }
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // the generator is a chain of ~100 short launches: overlap this grid's prologue with the previous grid's tail
  griddep_launch();
  griddep_wait();

  // work items: (m-tile, n-tile), or (pair of adjacent m-tiles, n-tile) with this CTA taking m-tile 2 * pair + rank (a
  // trailing odd tile's partner has b == B: its loads are zero-filled and its epilogue stores nothing)
  const int num_m_tiles = sh.B * sh.m_tiles_per_row;
  const int num_tiles = (kPair ? (num_m_tiles + 1) / 2 : num_m_tiles) * sh.n_tiles;
  const int tile_first = kPair ? (int)(blockIdx.x >> 1) : (int)blockIdx.x, tile_step = kPair ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int k_blocks = sh.tap_groups * sh.cc_blocks;
  auto m_tile_of = [&](int tile) { return kPair ? 2 * (tile / sh.n_tiles) + (int)rank : tile / sh.n_tiles; };

  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = tile_first; tile < num_tiles; tile += tile_step) {
        const int n_tile = tile % sh.n_tiles, m_tile = m_tile_of(tile);
        const int b = m_tile / sh.m_tiles_per_row;
        const int t0 = (m_tile - b * sh.m_tiles_per_row) * kBlockM;
        for (int kb = 0; kb < k_blocks; ++kb, ++it) {
          const uint32_t stage = it % kStages, parity = (it / kStages) & 1u;
          mbar_wait(&empty_bar[stage], parity ^ 1u, 1);
          if (sh.trace != nullptr && blockIdx.x == 0 && (kb == 0 || kb == k_blocks - 1)) {
            const int tl = (tile - tile_first) / tile_step;
            if (tl < 64) sh.trace[tl * 16 + (kb == 0 ? 0 : 1)] = clock64();
          }
          uint8_t* sa = smem + stage * Cfg::kStageBytes;
          const int tg = kb / sh.cc_blocks, cc = kb - tg * sh.cc_blocks;
          if constexpr (kPair) {
            if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * Cfg::kStageBytes);
            else mbar_arrive_leader(&full_bar[stage]);
            tma_load_4d_2sm(sa, &tmA, &full_bar[stage], cc * kBlockK, tg * sh.g, t0, b);
            tma_load_2d_2sm(sa + Cfg::kABytes, &tmB, &full_bar[stage], kb * kBlockK, n_tile * BN + (int)rank * (BN / 2));
          } else {
            if constexpr (kGather) {
              mbar_arrive_expect_tx(&full_bar[stage], Cfg::kBBytes);
            } else {
              mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStageBytes);
              tma_load_4d(sa, &tmA, &full_bar[stage], cc * kBlockK, tg * sh.g, t0, b);
            }
            tma_load_2d(sa + Cfg::kABytes, &tmB, &full_bar[stage], kb * kBlockK, n_tile * BN);
          }
        }
      }
    }
  } else if (warp == 1) {
    // MMA issue.  The whole warp runs the loop with warp-uniform values and ONE elected lane issues (inside the asm): an
    // `if (lane == 0)` around tcgen05.mma makes the compiler broadcast every operand through an ELECT / R2UR loop, ~60-70 cycles
    // per MMA -- as long as an N = 128 MMA itself, so the issuing thread, not the tensor pipe, set the pace of the 128-channel
    // convolutions (ncu: the epilogue warps spent 45 % of their samples waiting for the accumulator).
    if (leader) {
      const uint32_t ring = __shfl_sync(0xffffffffu, smem_u32(smem), 0);
      const uint32_t tmem = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t idesc = sh.idesc;
      constexpr uint32_t kDescHi = (1024u >> 4) | (1u << 14) | (2u << 29);   // SBO 1024, version 1, SWIZZLE_128B
      uint32_t stage = 0, parity = 0, ti = 0;
      for (int tile = tile_first; tile < num_tiles; tile += tile_step, ++ti) {
        const uint32_t as = ti & 1u, aph = (ti >> 1) & 1u;
        mbar_wait(&tempty_bar[as], aph ^ 1u, 2);
        tc_fence_after();
        const bool tr = sh.trace != nullptr && blockIdx.x == 0 && ti < 64u && lane == 0;
        long long waited = 0;
        if (tr) sh.trace[ti * 16 + 2] = clock64();
        const uint32_t d_tmem = tmem + as * BN;
#pragma unroll 1
        for (int kb = 0; kb < k_blocks; ++kb) {
          const long long w0 = tr ? clock64() : 0;
          mbar_wait(&full_bar[stage], parity, 3);
          tc_fence_after();
          if (tr) {
            const long long w1 = clock64();
            waited += w1 - w0;
            if (kb == 0) sh.trace[ti * 16 + 3] = w1;
            if (kb == k_blocks - 1) { sh.trace[ti * 16 + 4] = w1; sh.trace[ti * 16 + 5] = waited; }
          }
          const uint32_t a_lo = ((ring + stage * Cfg::kStageBytes) >> 4) & 0x3FFFu;
          const uint32_t b_lo = ((ring + stage * Cfg::kStageBytes + Cfg::kABytes) >> 4) & 0x3FFFu;
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k)
            umma_f16_elect<kPair>(d_tmem, a_lo + 2u * k, kDescHi, b_lo + 2u * k, kDescHi, idesc, (kb | k) != 0 ? 1u : 0u);
          umma_commit_elect_t<kPair>(&empty_bar[stage]);
          if (++stage == (uint32_t)kStages) { stage = 0; parity ^= 1u; }
        }
        umma_commit_elect_t<kPair>(&tfull_bar[as]);
      }
    }
  } else if (warp >= 2 + kEpiWarps) {
    // ===================== gather producers (C < 64): thread tau builds row tau of every A tile =====================
    if constexpr (kGather) {
      const int tau = (int)threadIdx.x - kGemmThreads;
      const int C = sh.a_C, half = (sh.a_taps - 1) / 2;
      const int cshift = C == 16 ? 1 : 2;                 // 16-byte chunks per tap piece: C / 8
      const long long pitch = (long long)(sh.T + 2 * sh.a_G) * C;
      // Completion protocol: a thread's copies of K block i are known to have landed when cp.async.wait_group leaves at most
      // kLag younger groups pending; it then makes them visible to the async proxy (the tensor core reads shared memory
      // through it) and arrives on that stage's barrier.  Doing the wait + fence HERE, kLag blocks late and spread over 128
      // threads, keeps it off the single MMA-issuing thread (where a per-block proxy fence cost ~1000 cycles per K block).
      constexpr int kLag = kStages >= 6 ? 4 : (kStages >= 4 ? 2 : 1);
      uint32_t it = 0;
      for (int tile = tile_first; tile < num_tiles; tile += tile_step) {
        const int m_tile = m_tile_of(tile);
        const int b = m_tile / sh.m_tiles_per_row;
        const int pos = (m_tile - b * sh.m_tiles_per_row) * kBlockM + tau;
        const bool in_utt = pos < sh.T;
        const __half* rowp = sh.a_ptr + (long long)b * pitch + (long long)(sh.a_G + pos) * C;
        for (int kb = 0; kb < k_blocks; ++kb, ++it) {
          const uint32_t stage = it % kStages, parity = (it / kStages) & 1u;
          mbar_wait(&empty_bar[stage], parity ^ 1u, 5);
          const uint32_t dst_row = smem_u32(smem + stage * Cfg::kStageBytes) + (uint32_t)tau * 128u;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int tap = kb * sh.g + (j >> cshift);
            const bool ok = in_utt && tap < sh.a_taps;
            const __half* src = rowp + (long long)(tap - half) * sh.a_dil * C + ((j & ((1 << cshift) - 1)) << 3);
            voc_cp_async16(dst_row + (uint32_t)((j ^ (tau & 7)) << 4), ok ? (const void*)src : (const void*)sh.a_ptr, ok);
          }
          asm volatile("cp.async.commit_group;" ::: "memory");
          if (it >= (uint32_t)kLag) {
            asm volatile("cp.async.wait_group %0;" ::"n"(kLag) : "memory");
            fence_proxy_async_smem();
            mbar_arrive(&full_bar[(it - kLag) % kStages]);
          }
        }
      }
      // drain: the last min(it, kLag) blocks
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      fence_proxy_async_smem();
      for (uint32_t i = it > (uint32_t)kLag ? it - kLag : 0u; i < it; ++i) mbar_arrive(&full_bar[i % kStages]);
    }
  } else {
    // ===================== epilogue warps =====================
    const int q = warp & 3, h = (warp - 2) >> 2;
    if constexpr (BN <= 32) {
      // Narrow tiles (16 / 32 channels): a tile is 128 x BN values and its epilogue is one chain of dependent latencies
      // (side-input loads -> math -> stores), so the two warp groups each own ONE accumulator stage and work on alternate
      // tiles, and the side inputs -- which do not depend on the accumulator -- are fetched before the MMA wait.  A row is
      // BN * 2 <= 64 contiguous bytes and consecutive rows are adjacent, so row-per-thread vector access is coalesced as is.
      constexpr int NV = BN / 8;                       // uint4 per row
      uint32_t ti = (uint32_t)h;
      for (int tile = blockIdx.x + h * (int)gridDim.x; tile < num_tiles; tile += 2 * (int)gridDim.x, ti += 2) {
        const int n_tile = tile % sh.n_tiles, m_tile = tile / sh.n_tiles;
        const int b = m_tile / sh.m_tiles_per_row;
        const int t = (m_tile - b * sh.m_tiles_per_row) * kBlockM + q * 32 + lane;
        const bool live = t < sh.T;
        const bool valid = ep.lengths == nullptr || t < ep.lengths[b] * ep.len_mul;   // ragged batch: rows past the utterance's own end become zeros
        const int gcol = n_tile * BN;
        const uint32_t as = ti & 1u, aph = (ti >> 1) & 1u;
        uint4 rs[NV], ad[NV];
        if (live && ep.res16 != nullptr) {
          const uint4* p = reinterpret_cast<const uint4*>(ep.res16 + (long long)b * ep.res_bstride + (long long)t * ep.res_ld + gcol);
#pragma unroll
          for (int j = 0; j < NV; ++j) rs[j] = p[j];
        }
        if (live && ep.add16 != nullptr) {
          const uint4* p = reinterpret_cast<const uint4*>(ep.add16 + (long long)b * ep.add_bstride + (long long)t * ep.add_ld + gcol);
#pragma unroll
          for (int j = 0; j < NV; ++j) ad[j] = p[j];
        }
        mbar_wait(&tfull_bar[as], aph, 4);
        tc_fence_after();
        uint32_t r[BN];
        const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + as * BN;
        if constexpr (BN == 32) tmem_ld_x32(trow, r);
        else tmem_ld_x16(trow, r);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty_bar[as]);    // the accumulator is in registers: hand the stage back early
        if (live) {
          float v[BN];
#pragma unroll
          for (int j = 0; j < BN; j += 4) {
            const float4 bv = __ldg(reinterpret_cast<const float4*>(ep.bias + gcol + j));
            v[j] = __uint_as_float(r[j]) + bv.x;
            v[j + 1] = __uint_as_float(r[j + 1]) + bv.y;
            v[j + 2] = __uint_as_float(r[j + 2]) + bv.z;
            v[j + 3] = __uint_as_float(r[j + 3]) + bv.w;
          }
          if (ep.res16 != nullptr) {
#pragma unroll
            for (int j = 0; j < NV; ++j) {
              const uint32_t ww[4] = {rs[j].x, rs[j].y, rs[j].z, rs[j].w};
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                float2 f = Cvt<__half>::unpack(ww[u]);
                f.x = fminf(f.x, f.x * ep.res_inv_slope);   // undo the stored LeakyReLU (inverse slope >= 1)
                f.y = fminf(f.y, f.y * ep.res_inv_slope);
                v[8 * j + 2 * u] += f.x;
                v[8 * j + 2 * u + 1] += f.y;
              }
            }
          }
          if (ep.add16 != nullptr) {
#pragma unroll
            for (int j = 0; j < NV; ++j) {
              const uint32_t ww[4] = {ad[j].x, ad[j].y, ad[j].z, ad[j].w};
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                const float2 f = Cvt<__half>::unpack(ww[u]);
                v[8 * j + 2 * u] += f.x;
                v[8 * j + 2 * u + 1] += f.y;
              }
            }
          }
#pragma unroll
          for (int j = 0; j < BN; ++j) v[j] = voc_finish(v[j], ep.scale, ep.act, ep.slope);
          if (!valid) {
#pragma unroll
            for (int j = 0; j < BN; ++j) v[j] = 0.f;
          }
          if (ep.out32 != nullptr) {
            if (gcol == 0) ep.out32[(long long)b * sh.T + t] = v[0];
          } else {
            uint4* p = reinterpret_cast<uint4*>(ep.out16 + (long long)b * ep.out_bstride + (long long)t * ep.out_ld + gcol);
#pragma unroll
            for (int j = 0; j < NV; ++j)
              p[j] = make_uint4(Cvt<__half>::pack(v[8 * j], v[8 * j + 1]), Cvt<__half>::pack(v[8 * j + 2], v[8 * j + 3]),
                                Cvt<__half>::pack(v[8 * j + 4], v[8 * j + 5]), Cvt<__half>::pack(v[8 * j + 6], v[8 * j + 7]));
          }
        }
      }
    } else {
    uint32_t* stg = staging + (warp - 2) * (32 * RW);
    constexpr int LPR = CW / 8;            // lanes per row segment (16 bytes each)
    constexpr int RPI = 32 / LPR;          // rows per warp-wide access
    constexpr int kChunks = BN >= 64 ? (BN / 2) / 32 : 1;
    const bool works = BN >= 64 || h == 0;  // narrow tiles: one warp per lane quadrant takes all BN columns
    uint32_t ti = 0;
    for (int tile = tile_first; tile < num_tiles; tile += tile_step, ++ti) {
      const int n_tile = tile % sh.n_tiles, m_tile = m_tile_of(tile);
      const int b = m_tile / sh.m_tiles_per_row;
      const int t0 = (m_tile - b * sh.m_tiles_per_row) * kBlockM;
      const int t = t0 + q * 32 + lane;                       // this thread's position (accumulator row)
      const int rows_left = b < sh.B ? sh.T - (t0 + q * 32) : 0;   // rows of this warp's 32-row slab inside the utterance (0: a pair's phantom tile)
      const bool valid = ep.lengths == nullptr || b >= sh.B || t < ep.lengths[b] * ep.len_mul;   // ragged batch: rows past the utterance's own end become zeros
      const uint32_t as = ti & 1u, aph = (ti >> 1) & 1u;
      const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + as * BN;
      // The residual does not depend on the accumulator: fetch this warp's blocks of it BEFORE waiting for the MMAs (tiles of
      // <= 128 columns: 8 uint4 per lane).  Loaded inside the chunk loop its latency was exposed once per chunk -- the second
      // convolution of a round took 10-50 % longer than the first for the same MMA work (profiles/r02g_voc_conv_residentA_off.csv).
      constexpr bool kPrefetchRes = BN >= 64 && BN <= 128;
      uint4 pre_res[kPrefetchRes ? kChunks : 1][32 / RPI];
      if constexpr (kPrefetchRes) {
        if (works && ep.res16 != nullptr) {
#pragma unroll
          for (int ci = 0; ci < kChunks; ++ci) {
            const int gcol = n_tile * BN + h * (BN / 2) + ci * CW;
            const __half* base = ep.res16 + (long long)b * ep.res_bstride + (long long)(t0 + q * 32) * ep.res_ld + gcol;
#pragma unroll
            for (int i = 0; i < 32 / RPI; ++i) {
              const int rrow = i * RPI + lane / LPR, c8 = lane % LPR;
              pre_res[ci][i] = rrow < rows_left ? *reinterpret_cast<const uint4*>(base + (long long)rrow * ep.res_ld + 8 * c8) : make_uint4(0u, 0u, 0u, 0u);
            }
          }
        }
      }
      mbar_wait(&tfull_bar[as], aph, 4);
      tc_fence_after();
      const bool etr = sh.trace != nullptr && blockIdx.x == 0 && ti < 64u && threadIdx.x == 64;
      if (etr) sh.trace[ti * 16 + 6] = clock64();
      if (works) {
#pragma unroll
        for (int ci = 0; ci < kChunks; ++ci) {
          const int c = (BN >= 64 ? h * (BN / 2) : 0) + ci * CW;   // first column of this chunk inside the tile
          const int gcol = n_tile * BN + c;
          uint32_t r[CW];
          if constexpr (CW == 32) tmem_ld_x32(trow + c, r);
          else tmem_ld_x16(trow + c, reinterpret_cast<uint32_t (&)[16]>(r));
          tmem_ld_wait();
          float v[CW];
#pragma unroll
          for (int j = 0; j < CW; j += 4) {
            const float4 bv = __ldg(reinterpret_cast<const float4*>(ep.bias + gcol + j));
            v[j] = __uint_as_float(r[j]) + bv.x;
            v[j + 1] = __uint_as_float(r[j + 1]) + bv.y;
            v[j + 2] = __uint_as_float(r[j + 2]) + bv.z;
            v[j + 3] = __uint_as_float(r[j + 3]) + bv.w;
          }
          // 16-bit side inputs: the [32 rows x CW] block arrives as full row segments and is read back one row per thread
          auto add_side = [&](const __half* src, long long bstride, int ld, bool unact, const uint4* pre) {
            const __half* base = src + (long long)b * bstride + (long long)(t0 + q * 32) * ld + gcol;
#pragma unroll
            for (int i = 0; i < 32 / RPI; ++i) {
              const int rrow = i * RPI + lane / LPR, c8 = lane % LPR;
              uint4 w = make_uint4(0u, 0u, 0u, 0u);
              if (pre != nullptr) w = pre[i];
              else if (rrow < rows_left) w = *reinterpret_cast<const uint4*>(base + (long long)rrow * ld + 8 * c8);
              *reinterpret_cast<uint4*>(stg + rrow * RW + 4 * c8) = w;
            }
            __syncwarp();
#pragma unroll
            for (int j = 0; j < CW; j += 8) {
              const uint4 w = *reinterpret_cast<const uint4*>(stg + lane * RW + j / 2);
              const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                float2 f = Cvt<__half>::unpack(ww[u]);
                if (unact) {
                  f.x = fminf(f.x, f.x * ep.res_inv_slope);   // undo the stored LeakyReLU (inverse slope >= 1)
                  f.y = fminf(f.y, f.y * ep.res_inv_slope);
                }
                v[j + 2 * u] += f.x;
                v[j + 2 * u + 1] += f.y;
              }
            }
            __syncwarp();
          };
          if (ep.res16 != nullptr) add_side(ep.res16, ep.res_bstride, ep.res_ld, true, kPrefetchRes ? pre_res[ci] : nullptr);
          if (ep.add16 != nullptr) add_side(ep.add16, ep.add_bstride, ep.add_ld, false, nullptr);
#pragma unroll
          for (int j = 0; j < CW; ++j) v[j] = voc_finish(v[j], ep.scale, ep.act, ep.slope);
          if (!valid) {
#pragma unroll
            for (int j = 0; j < CW; ++j) v[j] = 0.f;
          }
          if (ep.out32 != nullptr) {
            if (gcol == 0 && t < sh.T && b < sh.B) ep.out32[(long long)b * sh.T + t] = v[0];
          } else {
#pragma unroll
            for (int j = 0; j < CW; j += 8)
              *reinterpret_cast<uint4*>(stg + lane * RW + j / 2) =
                  make_uint4(Cvt<__half>::pack(v[j], v[j + 1]), Cvt<__half>::pack(v[j + 2], v[j + 3]),
                             Cvt<__half>::pack(v[j + 4], v[j + 5]), Cvt<__half>::pack(v[j + 6], v[j + 7]));
            __syncwarp();
            __half* base = ep.out16 + (long long)b * ep.out_bstride + (long long)(t0 + q * 32) * ep.out_ld + gcol;
#pragma unroll
            for (int i = 0; i < 32 / RPI; ++i) {
              const int rrow = i * RPI + lane / LPR, c8 = lane % LPR;
              if (rrow < rows_left)
                *reinterpret_cast<uint4*>(base + (long long)rrow * ep.out_ld + 8 * c8) = *reinterpret_cast<const uint4*>(stg + rrow * RW + 4 * c8);
            }
            __syncwarp();
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (etr) sh.trace[ti * 16 + 7] = clock64();
      if (lane == 0) {
        if constexpr (kPair) mbar_arrive_leader(&tempty_bar[as]);
        else mbar_arrive(&tempty_bar[as]);
      }
    }
    }
  }

  tc_fence_before();
  if constexpr (kPair) cluster_sync_all();   // the leader's MMAs read the peer's shared memory and write its TMEM
  else __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    if constexpr (kPair) tmem_dealloc_2sm<kTmemCols>(tmem_base);
    else tmem_dealloc<kTmemCols>(tmem_base);
  }
}

// ------------------------------------------------------------------ weight packing
// Conv1d weight [Cout, Cin, k] fp32 -> [Cout rows (zero rows up to n_pad)][K = tap * Cin + c, taps padded to taps_pad] fp16.
__global__ void k_voc_pack_conv(__half* __restrict__ dst, const float* __restrict__ src, int Cout, int Cin, int k, int n_pad, int taps_pad) {
  const long long total = (long long)n_pad * taps_pad * Cin;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / ((long long)taps_pad * Cin));
    const int rem = (int)(i - (long long)r * taps_pad * Cin);
    const int tap = rem / Cin, c = rem - tap * Cin;
    const float v = (r < Cout && tap < k) ? src[((long long)r * Cin + c) * k + tap] : 0.f;
    dst[i] = __float2half_rn(v);
  }
}
// ConvTranspose1d weight [Cin, Cout, k = 2s] (stride s, padding s / 2) -> the equivalent 3-tap convolution with
// N = s * Cout rows: row (phase r, co), tap j in {0, 1, 2} = input position q - 1 + j for output position q * s + r:
//   out[q * s + r] = sum over i, kk with i * s - s/2 + kk = q * s + r  of  x[i] . W[:, :, kk]   ->  kk = r + s/2 - (j - 1) * s.
__global__ void k_voc_pack_convT(__half* __restrict__ dst, const float* __restrict__ src, int Cin, int Cout, int s, int taps_pad) {
  const int k = 2 * s;
  const long long total = (long long)s * Cout * taps_pad * Cin;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int row = (int)(i / ((long long)taps_pad * Cin));
    const int rem = (int)(i - (long long)row * taps_pad * Cin);
    const int j = rem / Cin, c = rem - j * Cin;
    const int r = row / Cout, co = row - r * Cout;
    const int kk = r + s / 2 - (j - 1) * s;
    const float v = (j < 3 && kk >= 0 && kk < k) ? src[((long long)c * Cout + co) * k + kk] : 0.f;
    dst[i] = __float2half_rn(v);
  }
}
// the bias of a transposed conv repeated per phase: [s * Cout]
__global__ void k_voc_tile_bias(float* __restrict__ dst, const float* __restrict__ src, int Cout, int n) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) dst[i] = src[i % Cout];
}
__global__ void k_voc_pad_bias(float* __restrict__ dst, const float* __restrict__ src, int n_src, int n) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) dst[i] = i < n_src ? src[i] : 0.f;
}

// ------------------------------------------------------------------ generator input / harmonic source
// c [B, M, T] fp32 (the sampler's hand-off tensor, nsf_hifigan.py:50-52) -> guarded channels-last fp16 [B][G + T + G][M]
// lengths: optional [B] valid frames per utterance; frames past it enter as zeros (whatever the caller's tensor holds there)
__global__ void k_voc_mel_in(__half* __restrict__ dst, const float* __restrict__ c, int M, int Tn, int G, const int* __restrict__ lengths) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z, m0 = blockIdx.y * 32, t0 = blockIdx.x * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
  const int Tb = lengths != nullptr ? min(Tn, lengths[b]) : Tn;
  for (int i = ty; i < 32; i += 8) {
    const int m = m0 + i, t = t0 + tx;
    tile[i][tx] = (m < M && t < Tb) ? c[((long long)b * M + m) * Tn + t] : 0.f;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int t = t0 + i, m = m0 + tx;
    if (t < Tn && m < M) dst[((long long)b * (Tn + 2 * G) + G + t) * M + m] = __float2half_rn(tile[tx][i]);
  }
}

constexpr int kVocHarm = 9;   // fundamental + 8 overtones (models.py:219, harmonic_num = 8)

// SineGen phase bookkeeping (models.py:138-170).  The reference integrates the per-sample phase increment
// rad = frac(f0 * (h + 1) / sr) -- constant over the `upp` samples of a frame, plus rand_ini on the whole first frame --
// with a float64 cumsum over samples and adds whole turns wherever a wrapped copy of the phase steps back; whole turns do
// not change sin(2 pi phase), so the phase of sample j of frame f is  P[f] + (j + 1) * rad[f]  with the frame-level
// exclusive sum P[f] = upp * sum_{f' < f} rad[f'] (float64, reduced mod 1).  One thread per (utterance, harmonic).
__global__ void k_voc_phase(double* __restrict__ P, float* __restrict__ rad_out, const float* __restrict__ f0, const float* __restrict__ rand_ini,
                            int B, int Tn, int upp, float sr) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * kVocHarm) return;
  const int b = idx / kVocHarm, h = idx - b * kVocHarm;
  double acc = 0.0;
  for (int f = 0; f < Tn; ++f) {
    const float fn = f0[(long long)b * Tn + f] * (float)(h + 1);
    float rad = fmodf(fn / sr, 1.f);
    if (f == 0) rad += rand_ini[h];
    P[((long long)b * Tn + f) * kVocHarm + h] = acc;
    rad_out[((long long)b * Tn + f) * kVocHarm + h] = rad;
    acc += (double)upp * (double)rad;
    acc -= floor(acc);
  }
}

// har_source[b, n] = tanh(l_linear(sine_waves * uv + noise)) (models.py:171-216): sine = sin(2 pi phase) * sine_amp,
// uv = f0 > 0, noise = (uv * noise_std + (1 - uv) * sine_amp / 3) * randn.  noise_in: optional injected N(0,1) draws
// [B, T * upp, 9] (parity tests); nullptr = in-kernel Philox.  One thread per sample.
__global__ void k_voc_source(float* __restrict__ har, const double* __restrict__ P, const float* __restrict__ rad, const float* __restrict__ f0,
                             const float* __restrict__ noise_in, const float* __restrict__ w_lin, const float* __restrict__ b_lin,
                             int Tn, int upp, long long total, float sine_amp, float noise_std, unsigned long long seed) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long fr = i / upp;               // b * T + f
    const int j = (int)(i - fr * upp);
    const float uv = f0[fr] > 0.f ? 1.f : 0.f;
    const float namp = uv * noise_std + (1.f - uv) * sine_amp / 3.f;
    float nz[12];
    if (noise_in != nullptr) {
#pragma unroll
      for (int h = 0; h < kVocHarm; ++h) nz[h] = noise_in[i * kVocHarm + h];
    } else {
#pragma unroll
      for (int g4 = 0; g4 < 3; ++g4) {
        const float4 n4 = philox_normal4(seed, 0x564f43u + g4, (uint32_t)i, (uint32_t)(i >> 32));
        nz[4 * g4] = n4.x; nz[4 * g4 + 1] = n4.y; nz[4 * g4 + 2] = n4.z; nz[4 * g4 + 3] = n4.w;
      }
    }
    float acc = b_lin[0];
#pragma unroll
    for (int h = 0; h < kVocHarm; ++h) {
      double ph = P[fr * kVocHarm + h] + (double)(j + 1) * (double)rad[fr * kVocHarm + h];
      ph -= floor(ph);
      const float s = sinf((float)(ph * 6.283185307179586476925)) * sine_amp;
      acc = fmaf(w_lin[h], s * uv + namp * nz[h], acc);
    }
    har[i] = tanhf(acc);
  }
}

// noise_convs[i] (models.py:233-238, :276): x_source[b, o, c] = bias[c] + sum_kk W[c, kk] * har[b, o * stride - pad + kk], zero
// padded, written raw as fp16 in the up-sampled tensor's guarded layout (it is the `add16` input of the transposed conv's
// epilogue).  wt is the weight TRANSPOSED to [k][C] (k_voc_transpose_src at pack time).  A block covers 16384 / C output
// positions x all C channels; a thread owns 8 channels x 8 positions (64 accumulators: per tap two float4 weight loads and 8
// shared-memory har samples feed 64 FMAs), its positions interleaved with the other threads' so that every store
// instruction of a warp writes whole 16-byte channel groups of adjacent positions -- contiguous memory.
constexpr int kSrcThreads = 256;
// ppt = positions per thread (8, fewer when the har window of a block would not fit in shared memory)
__host__ __device__ inline int voc_src_positions(int C, int ppt) { return ppt * (kSrcThreads / (C / 8)); }
template <bool kVec4>
__global__ void __launch_bounds__(kSrcThreads, 2)
k_voc_srcconv(__half* __restrict__ dst, const float* __restrict__ har, const float* __restrict__ wt, const float* __restrict__ bias,
              int C, int k, int stride, int pad, int To, long long Lh, int G, int ppt, const int* __restrict__ lengths, int upp) {
  extern __shared__ __align__(16) float s_h[];   // (positions - 1) * stride + k samples
  const int n_cg = C / 8, n_ps = kSrcThreads / n_cg, n_pos = ppt * n_ps;
  const int b = blockIdx.y, o0 = blockIdx.x * n_pos;
  const int win = (n_pos - 1) * stride + k;
  const long long h0 = (long long)o0 * stride - pad;
  const long long Lb = lengths != nullptr ? min(Lh, (long long)lengths[b] * upp) : Lh;   // ragged batch: the source ends with the utterance
  for (int i = threadIdx.x; i < win; i += kSrcThreads) {
    const long long n = h0 + i;
    s_h[i] = (n >= 0 && n < Lb) ? har[(long long)b * Lh + n] : 0.f;
  }
  __syncthreads();
  const int cg = threadIdx.x % n_cg, ps = threadIdx.x / n_cg;   // positions ps, ps + n_ps, ..., ps + 7 * n_ps of the block
  float acc[8][8];
  {
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + cg * 8)), b1 = __ldg(reinterpret_cast<const float4*>(bias + cg * 8 + 4));
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      acc[i][0] = b0.x; acc[i][1] = b0.y; acc[i][2] = b0.z; acc[i][3] = b0.w;
      acc[i][4] = b1.x; acc[i][5] = b1.y; acc[i][6] = b1.z; acc[i][7] = b1.w;
    }
  }
  const float* hp = s_h + ps * stride;
  const int pstep = n_ps * stride;
  if (kVec4) {   // stride % 4 == 0 and k % 4 == 0: four taps per 16-byte shared-memory load (the scalar form is LSU-bound)
    for (int kk = 0; kk < k; kk += 4) {
      float4 hv[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) hv[i] = i < ppt ? *reinterpret_cast<const float4*>(hp + i * pstep + kk) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const float4 w0 = __ldg(reinterpret_cast<const float4*>(wt + (long long)(kk + t) * C + cg * 8));
        const float4 w1 = __ldg(reinterpret_cast<const float4*>(wt + (long long)(kk + t) * C + cg * 8 + 4));
        const float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float h1 = t == 0 ? hv[i].x : (t == 1 ? hv[i].y : (t == 2 ? hv[i].z : hv[i].w));
#pragma unroll
          for (int u = 0; u < 8; ++u) acc[i][u] = fmaf(w[u], h1, acc[i][u]);
        }
      }
    }
  } else {
    for (int kk = 0; kk < k; ++kk) {
      const float4 w0 = __ldg(reinterpret_cast<const float4*>(wt + (long long)kk * C + cg * 8));
      const float4 w1 = __ldg(reinterpret_cast<const float4*>(wt + (long long)kk * C + cg * 8 + 4));
      const float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float hv = i < ppt ? hp[i * pstep + kk] : 0.f;
#pragma unroll
        for (int u = 0; u < 8; ++u) acc[i][u] = fmaf(w[u], hv, acc[i][u]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int o = o0 + ps + i * n_ps;
    if (i < ppt && o < To)
      *reinterpret_cast<uint4*>(dst + ((long long)b * (To + 2 * G) + G + o) * C + cg * 8) =
          make_uint4(Cvt<__half>::pack(acc[i][0], acc[i][1]), Cvt<__half>::pack(acc[i][2], acc[i][3]), Cvt<__half>::pack(acc[i][4], acc[i][5]),
                     Cvt<__half>::pack(acc[i][6], acc[i][7]));
  }
}
// noise_convs weight [C, 1, k] -> [k][C]
__global__ void k_voc_transpose_src(float* __restrict__ dst, const float* __restrict__ src, int C, int k) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < C * k; i += gridDim.x * blockDim.x) {
    const int kk = i / C, c = i - kk * C;
    dst[i] = src[c * k + kk];
  }
}

// zero the guard rows of up to four guarded activation buffers of one geometry ([B][G + T + G][C]): front and back G rows of every
// utterance, 16 bytes per thread (one launch per stage instead of one per buffer: the chain is launch-bound for a single utterance)
struct VocGuardBufs { __half* p[4]; int n; };
__global__ void k_voc_zero_guards(VocGuardBufs bufs, int Tn, int C, int G, int B) {
  const long long per = 2ll * G * C / 8;        // 16-byte pieces per utterance (C is a multiple of 8)
  const long long total = per * B * bufs.n;
  const int c8n = C / 8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int which = (int)(i / (per * B));
    const long long j = i - (long long)which * per * B;
    const int b = (int)(j / per);
    const long long r = j - (long long)b * per;
    const long long row = r / c8n, c8 = r - row * c8n;
    const long long prow = row < G ? row : (long long)Tn + row;   // rows [0, G) and [G + T, G + T + G)
    *reinterpret_cast<uint4*>(bufs.p[which] + ((long long)b * (Tn + 2 * G) + prow) * C + c8 * 8) = make_uint4(0u, 0u, 0u, 0u);
  }
}

}  // namespace pd
