// Skip sum + sampler tail in one kernel (sampling loop, C = 256, M = 128, L layers):
//
//   s1 = (sum_l W_skip,l . y_l + sum_l b_l) / sqrt(L)       deferred skip sum, ONE GEMM with K = L*C     wavenet.py:119
//   s2 = relu(W_sp . s1 + b_sp)                             skip_projection                              wavenet.py:120-121
//   x0 = W_out . s2 + b_out ;  x = c1 x0 + c2 x + sigma eps output_projection + q-posterior              wavenet.py:122, prodiff.py:106-121
//   xa = relu(W_in . x + b_in)                              next step's input_projection                 wavenet.py:108-111
//
// The skip-sum GEMM streams 655 MB of gated activations per launch and is bound by HBM/L2 feed (80 K blocks per tile,
// ~45 k cycles), while its epilogue warps work for ~3 k cycles per tile.  The tail (pd_tail.cuh) is bound by exactly those
// warps (Philox + posterior + three epilogues, ~20 k cycles per tile).  Here the two run in one persistent kernel: while the
// tensor core and the TMA producer walk the K loop of tile t+1, the epilogue warps take tile t through the tail chain,
// and the MMA issuer slips the three small chain GEMMs of tile t between K blocks of tile t+1 as their operands appear.
// s1 never leaves the SM.
//
// 2-CTA clusters (cta_group::2, M = 256 = two adjacent frame tiles, each CTA stages half of every weight block), 352
// threads: warp 0 = producer of the K-loop ring, warp 1 = MMA issuer (leader CTA), warps 2..9 = epilogue, warp 10 =
// producer of the chain-weight ring (its own ring, so its order is independent of the K loop's).
// Shared memory per CTA: K-loop ring 4 x {A 16 KB, B half 16 KB}; chain ring 2 x 16 KB; one 64 KB operand tile SB that
// holds, in turn, s1, s2, the 16-bit x and the staged xa.  TMEM: S (skip sum) [0, 256), chain D2 [256, 512), D3
// [256, 384), D4 [256, 512) -- each chain accumulator is drained before the next one is started.  S is single-buffered:
// the K loop of tile t+1 starts when epilogue 1 of tile t has drained it (~3 k cycles of a ~45 k-cycle tile).
#pragma once
#include "pd_tail.cuh"

namespace pd {

struct SumTailParams {
  TailParams tp;          // B, T, m_tiles_per_row, do_in, biases, posterior fields, lens, trace (idesc fields unused)
  int L;                  // layers: the K loop has 4 * L blocks
  uint32_t idesc256, idesc128;   // cta_group::2 instruction descriptors (M = 256), N = 256 / 128
  const float* b_sum;     // [256] sum of the skip biases
  float scale;            // 1 / sqrt(L)
};

constexpr int kSumTailThreads = 352;
constexpr uint32_t kStStages = 4;
constexpr uint32_t kStStageBytes = 32768;
constexpr uint32_t kStChainStages = 2;
constexpr uint32_t kSumTailSmemBytes = kStStages * kStStageBytes + kStChainStages * 16384 + 4 * 16384 + 1024 + 256;

template <typename T>
__global__ void __launch_bounds__(kSumTailThreads, 1)
sumtail_kernel(const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmWskip,
               const __grid_constant__ CUtensorMap tmWsp, const __grid_constant__ CUtensorMap tmWout,
               const __grid_constant__ CUtensorMap tmWin, const __grid_constant__ CUtensorMap tmXst, const SumTailParams sp) {
  const TailParams& tp = sp.tp;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sC = smem + kStStages * kStStageBytes;                    // chain-weight ring
  uint8_t* sB = sC + kStChainStages * 16384;                         // 4 x [128 x 64] operand / staging tile
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + 4 * 16384);
  uint64_t* full_bar = bars;                          // [4] leader: K-loop block landed in both CTAs
  uint64_t* empty_bar = full_bar + kStStages;         // [4] local : MMAs finished reading the block
  uint64_t* cfull_bar = empty_bar + kStStages;        // [2] leader: chain weight block landed in both CTAs
  uint64_t* cempty_bar = cfull_bar + kStChainStages;  // [2] local
  uint64_t* acc_bar = cempty_bar + kStChainStages;    // [4] local : S / D2 / D3 / D4 complete
  uint64_t* sfree_bar = acc_bar + 4;                  // leader: S drained by all 16 epilogue warps
  uint64_t* opnd_bar = sfree_bar + 1;                 // [3] leader: s1 / s2 / 16-bit x written in both CTAs
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(opnd_bar + 3);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const bool tr = tp.trace != nullptr && blockIdx.x == 0;
  if (tr && threadIdx.x == 0) tp.trace[15] = clock64();

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmY);
    tma_prefetch_desc(&tmWskip);
    tma_prefetch_desc(&tmWsp);
    tma_prefetch_desc(&tmWout);
    tma_prefetch_desc(&tmWin);
    tma_prefetch_desc(&tmXst);
  }
  if (warp == 1 && lane == 0) {
    for (uint32_t i = 0; i < kStStages; ++i) {
      mbar_init(&full_bar[i], 2);
      mbar_init(&empty_bar[i], 1);
    }
    for (uint32_t i = 0; i < kStChainStages; ++i) {
      mbar_init(&cfull_bar[i], 2);
      mbar_init(&cempty_bar[i], 1);
    }
    for (int i = 0; i < 4; ++i) mbar_init(&acc_bar[i], 1);
    mbar_init(sfree_bar, 2 * kEpiWarps);
    for (int i = 0; i < 3; ++i) mbar_init(&opnd_bar[i], 2 * kEpiWarps);
    fence_mbar_init();
  }
  cluster_sync_all();
  if (warp == 2) tmem_alloc_2sm<kTmemCols>(tmem_slot);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int num_m_tiles = tp.B * tp.m_tiles_per_row;
  const int num_pairs = (num_m_tiles + 1) >> 1;
  const int num_clusters = gridDim.x >> 1;
  const int cluster_id = blockIdx.x >> 1;
  const int k_blocks = 4 * sp.L;
  const int n_chain = tp.do_in ? 10 : 8;   // chain weight blocks per tile: 4 (skip_proj) + 4 (output_proj) [+ 2 (input_proj)]

  if (warp == 0) {
    // ===================== K-loop producer (both CTAs): gated activations y_l and this CTA's half of W_skip =====================
    if (lane == 0) {
      uint32_t it = 0;
      for (int pair = cluster_id; pair < num_pairs; pair += num_clusters) {
        const int m_tile = 2 * pair + (int)rank;
        const int b = m_tile / tp.m_tiles_per_row;            // == B for the dummy tile of an odd count: fully out of bounds
        const int t0 = (m_tile - b * tp.m_tiles_per_row) * kBlockM;
        // The ring holds 4 x 32 KB per SM, too little to cover HBM latency at this rate (the stand-alone skip_sum GEMM has 7
        // stages): the activation blocks are pulled into L2 kPf blocks ahead, so the ring's TMA loads only see L2 latency.
        constexpr int kPf = 12;
        for (int kb = 0; kb < kPf && kb < k_blocks; ++kb) tma_prefetch_4d(&tmY, (kb & 3) * kBlockK, t0, b, kb >> 2);
        for (int kb = 0; kb < k_blocks; ++kb, ++it) {
          if (kb + kPf < k_blocks) tma_prefetch_4d(&tmY, ((kb + kPf) & 3) * kBlockK, t0, b, (kb + kPf) >> 2);
          const uint32_t stage = it % kStStages, parity = (it / kStStages) & 1u;
          mbar_wait(&empty_bar[stage], parity ^ 1u, 1);
          if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * kStStageBytes);
          else mbar_arrive_leader(&full_bar[stage]);
          uint8_t* sa = smem + stage * kStStageBytes;
          tma_load_4d_2sm(sa, &tmY, &full_bar[stage], (kb & 3) * kBlockK, t0, b, kb >> 2);
          tma_load_2d_2sm(sa + 16384, &tmWskip, &full_bar[stage], kb * kBlockK, (int)rank * 128);
        }
      }
    }
  } else if (warp == 10) {
    // ===================== chain-weight producer (both CTAs) =====================
    if (lane == 0) {
      uint32_t it = 0;
      for (int pair = cluster_id; pair < num_pairs; pair += num_clusters) {
        for (int blk = 0; blk < n_chain; ++blk, ++it) {
          const uint32_t stage = it % kStChainStages, parity = (it / kStChainStages) & 1u;
          mbar_wait(&cempty_bar[stage], parity ^ 1u, 2);
          const uint32_t bytes = (blk >= 4 && blk < 8) ? 8192u : 16384u;   // this CTA's half: 128 / 64 / 128 weight rows
          if (leader) mbar_arrive_expect_tx(&cfull_bar[stage], 2 * bytes);
          else mbar_arrive_leader(&cfull_bar[stage]);
          uint8_t* dst = sC + stage * 16384;
          if (blk < 4) tma_load_2d_2sm(dst, &tmWsp, &cfull_bar[stage], blk * kBlockK, (int)rank * 128);
          else if (blk < 8) tma_load_2d_2sm(dst, &tmWout, &cfull_bar[stage], (blk - 4) * kBlockK, (int)rank * 64);
          else tma_load_2d_2sm(dst, &tmWin, &cfull_bar[stage], (blk - 8) * kBlockK, (int)rank * 128);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA) =====================
    if (leader && lane == 0) {
      uint32_t it = 0, cit = 0, ti = 0;
      const uint32_t sBa = smem_u32(sB);
      // chain GEMMs of the previous tile: stage 0 none pending, 1 = D2 waits for s1, 2 = D3 waits for s2, 3 = D4 waits for x
      int chain_stage = 0;
      uint32_t chain_par = 0;
      auto chain_gemm = [&](int n_blk, uint32_t d_tmem, uint32_t idesc) {
        for (int kb = 0; kb < n_blk; ++kb, ++cit) {
          const uint32_t stage = cit % kStChainStages, parity = (cit / kStChainStages) & 1u;
          mbar_wait(&cfull_bar[stage], parity, 3);
          tc_fence_after();
          const uint64_t adesc = make_sw128_kmajor_desc(sBa + kb * 16384);
          const uint64_t bdesc = make_sw128_kmajor_desc(smem_u32(sC + stage * 16384));
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k)
            umma_f16_2sm(d_tmem, desc_advance_k(adesc, k), desc_advance_k(bdesc, k), idesc, (kb | k) != 0 ? 1u : 0u);
          umma_commit_2sm(&cempty_bar[stage]);
        }
      };
      // issue the next chain GEMM of the previous tile if its operand is ready (or wait for it when `block` is set)
      auto service_chain = [&](bool block) {
        while (chain_stage != 0) {
          uint64_t* bar = &opnd_bar[chain_stage - 1];
          if (block) mbar_wait(bar, chain_par, 4);
          else if (!mbar_test(bar, chain_par)) return;   // a poll between two K blocks must not suspend the issuer
          tc_fence_after();
          if (chain_stage == 1) {
            chain_gemm(4, tmem_base + 256, sp.idesc256);      // D2 = s1 . W_sp^T
            umma_commit_2sm(&acc_bar[1]);
            if (tr) tp.trace[((ti - 1) & 7) * 16 + 2] = clock64();
            chain_stage = 2;
          } else if (chain_stage == 2) {
            chain_gemm(4, tmem_base + 256, sp.idesc128);      // D3 = s2 . W_out^T
            umma_commit_2sm(&acc_bar[2]);
            if (tr) tp.trace[((ti - 1) & 7) * 16 + 3] = clock64();
            chain_stage = tp.do_in ? 3 : 0;
          } else {
            chain_gemm(2, tmem_base + 256, sp.idesc256);      // D4 = x . W_in^T
            umma_commit_2sm(&acc_bar[3]);
            if (tr) tp.trace[((ti - 1) & 7) * 16 + 4] = clock64();
            chain_stage = 0;
          }
        }
      };
      for (int pair = cluster_id; pair < num_pairs; pair += num_clusters, ++ti) {
        const uint32_t par = ti & 1u;
        mbar_wait(sfree_bar, par ^ 1u, 5);       // epilogue 1 of the previous tile has drained S
        tc_fence_after();
        if (tr) tp.trace[(ti & 7) * 16 + 0] = clock64();
        for (int kb = 0; kb < k_blocks; ++kb, ++it) {
          service_chain(false);
          const uint32_t stage = it % kStStages, parity = (it / kStStages) & 1u;
          mbar_wait(&full_bar[stage], parity, 6);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * kStStageBytes);
          const uint64_t adesc = make_sw128_kmajor_desc(sa);
          const uint64_t bdesc = make_sw128_kmajor_desc(sa + 16384);
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k)
            umma_f16_2sm(tmem_base, desc_advance_k(adesc, k), desc_advance_k(bdesc, k), sp.idesc256, (kb | k) != 0 ? 1u : 0u);
          umma_commit_2sm(&empty_bar[stage]);
        }
        umma_commit_2sm(&acc_bar[0]);
        if (tr) tp.trace[(ti & 7) * 16 + 1] = clock64();
        service_chain(true);                     // the previous tile's chain must be fully issued before this tile's begins
        chain_stage = 1;
        chain_par = par;
      }
      service_chain(true);                       // last tile: nothing left to hide behind
    }
  } else {
    // ===================== epilogue warps (both CTAs) =====================
    const int q = warp & 3, h = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    const int M = tp.M;
    const long long Tn = tp.T;
    uint32_t ti = 0;
    // relu(acc + bias) * scale of this warp's 128 columns -> SB sub-tiles 2h, 2h + 1 (swizzled K-major operand tile)
    auto drain256 = [&](uint32_t col0, const float* bias, float scale, bool relu, bool live) {
      uint32_t rr[2][32];
      tmem_ld_x32(lane_base + col0 + h * 128, rr[0]);
#pragma unroll
      for (int ci = 0; ci < 4; ++ci) {
        const int c = h * 128 + ci * 32;
        tmem_ld_wait();
        if (ci < 3) tmem_ld_x32(lane_base + col0 + c + 32, rr[(ci + 1) & 1]);
        const uint32_t (&r)[32] = rr[ci & 1];
        uint8_t* st = sB + (c >> 6) * 16384;
        const int ch0 = (c & 63) >> 3;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint32_t ow[4];
#pragma unroll
          for (int u = 0; u < 4; u += 2) {
            const float4 bv = __ldg(reinterpret_cast<const float4*>(bias + c + 8 * j + 2 * u));
            const int col = 8 * j + 2 * u;
            float v0 = (__uint_as_float(r[col]) + bv.x) * scale, v1 = (__uint_as_float(r[col + 1]) + bv.y) * scale;
            float v2 = (__uint_as_float(r[col + 2]) + bv.z) * scale, v3 = (__uint_as_float(r[col + 3]) + bv.w) * scale;
            if (relu) {
              v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); v2 = fmaxf(v2, 0.f); v3 = fmaxf(v3, 0.f);
            }
            ow[u] = Cvt<T>::pack(v0, v1);
            ow[u + 1] = Cvt<T>::pack(v2, v3);
          }
          *sw128_chunk(st, row, ch0 + j) = live ? make_uint4(ow[0], ow[1], ow[2], ow[3]) : make_uint4(0u, 0u, 0u, 0u);
        }
      }
    };
    for (int pair = cluster_id; pair < num_pairs; pair += num_clusters, ++ti) {
      const uint32_t par = ti & 1u;
      const int m_tile = 2 * pair + (int)rank;
      const bool real = m_tile < num_m_tiles;
      const int b = m_tile / tp.m_tiles_per_row;
      const int t0 = (m_tile - b * tp.m_tiles_per_row) * kBlockM;
      const int t = t0 + row;
      const bool valid = real && t < tp.T;
      const long long n = (long long)b * tp.T + t;
      const bool act = valid && (h * 64 < M);
      const long long col0 = ((long long)b * M + h * 64) * Tn + t;

      // ---------------- epilogue 1: s1 = (S + sum b) / sqrt(L) -> SB (A operand of the skip_proj GEMM)
      if (lane == 0) tma_store_wait_read<0>();   // this warp's rows of SB were last read by its own xa store
      __syncwarp();
      mbar_wait(&acc_bar[0], par, 7);
      tc_fence_after();
      if (tr && warp == 2 && lane == 0) tp.trace[(ti & 7) * 16 + 5] = clock64();
      drain256(0, sp.b_sum, sp.scale, false, true);
      tc_fence_before();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive_leader(sfree_bar);
        mbar_arrive_leader(&opnd_bar[0]);
      }
      if (tr && warp == 2 && lane == 0) tp.trace[(ti & 7) * 16 + 6] = clock64();

      // ---------------- epilogue 2: s2 = relu(D2 + b_sp) -> SB in place (the skip_proj MMAs have finished reading s1)
      mbar_wait(&acc_bar[1], par, 8);
      tc_fence_after();
      drain256(256, tp.b_sp, 1.f, true, true);
      tc_fence_before();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(&opnd_bar[1]);
      if (tr && warp == 2 && lane == 0) tp.trace[(ti & 7) * 16 + 7] = clock64();

      // ---------------- q-posterior operand pv = c2 x_t + sigma eps (independent of the GEMMs; runs while D3 is computed)
      float pv[64];
      if (act) {
        const float* xs = tp.x_state + col0;
#pragma unroll
        for (int j = 0; j < 64; ++j) pv[j] = xs[(long long)j * Tn];
        if (tp.sigma != 0.f) {
          if (tp.noise == nullptr) {
#pragma unroll
            for (int j = 0; j < 64; j += 4) {
              const float4 nz = philox_normal4(tp.seed, tp.step_id, (uint32_t)n, (uint32_t)((h * 64 + j) >> 2));
              pv[j] = fmaf(tp.sigma, nz.x, tp.c2 * pv[j]);
              pv[j + 1] = fmaf(tp.sigma, nz.y, tp.c2 * pv[j + 1]);
              pv[j + 2] = fmaf(tp.sigma, nz.z, tp.c2 * pv[j + 2]);
              pv[j + 3] = fmaf(tp.sigma, nz.w, tp.c2 * pv[j + 3]);
            }
          } else {
            const float* nzp = tp.noise + col0;
#pragma unroll
            for (int j = 0; j < 64; ++j) pv[j] = fmaf(tp.sigma, __ldg(nzp + (long long)j * Tn), tp.c2 * pv[j]);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 64; ++j) pv[j] = tp.c2 * pv[j];
        }
      }

      if (tr && warp == 2 && lane == 0) tp.trace[(ti & 7) * 16 + 8] = clock64();
      // ---------------- epilogue 3: x0 = D3 + b_out, x = c1 x0 + pv, 16-bit x -> SB sub-tile h (A of the input_proj GEMM)
      mbar_wait(&acc_bar[2], par, 9);
      tc_fence_after();
      {
        uint8_t* st = sB + h * 16384;
#pragma unroll
        for (int ci = 0; ci < 4; ++ci) {
          const int c = h * 64 + ci * 16;
          uint32_t r[16];
          tmem_ld_x16(lane_base + 256 + c, r);
          tmem_ld_wait();
          float xn[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) xn[j] = 0.f;
          if (act) {
            const long long cbase = col0 + (long long)ci * 16 * Tn;
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
              const float4 bv = __ldg(reinterpret_cast<const float4*>(tp.b_out + c + j));
              xn[j] = __uint_as_float(r[j]) + bv.x;
              xn[j + 1] = __uint_as_float(r[j + 1]) + bv.y;
              xn[j + 2] = __uint_as_float(r[j + 2]) + bv.z;
              xn[j + 3] = __uint_as_float(r[j + 3]) + bv.w;
            }
            if (tp.x0_out) {
#pragma unroll
              for (int j = 0; j < 16; ++j) tp.x0_out[cbase + (long long)j * Tn] = xn[j];
            }
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              xn[j] = fmaf(tp.c1, xn[j], pv[ci * 16 + j]);
              if (tp.out_scale != 1.f) xn[j] = __fmul_rn(xn[j], tp.out_scale);   // vocoder hand-off, last step only
              tp.x_state[cbase + (long long)j * Tn] = xn[j];
            }
          }
          if (tp.do_in && h * 64 < M) {
            *sw128_chunk(st, row, 2 * ci) = make_uint4(Cvt<T>::pack(xn[0], xn[1]), Cvt<T>::pack(xn[2], xn[3]), Cvt<T>::pack(xn[4], xn[5]), Cvt<T>::pack(xn[6], xn[7]));
            *sw128_chunk(st, row, 2 * ci + 1) = make_uint4(Cvt<T>::pack(xn[8], xn[9]), Cvt<T>::pack(xn[10], xn[11]), Cvt<T>::pack(xn[12], xn[13]), Cvt<T>::pack(xn[14], xn[15]));
          }
        }
      }
      tc_fence_before();
      if (tp.do_in) {
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive_leader(&opnd_bar[2]);
        if (tr && warp == 2 && lane == 0) tp.trace[(ti & 7) * 16 + 9] = clock64();

        // ---------------- epilogue 4: xa = relu(D4 + b_in) -> SB -> TMA store (the next step's layer-0 input)
        mbar_wait(&acc_bar[3], par, 10);
        tc_fence_after();
        const bool live = tp.lens == nullptr || (real && t < __ldg(tp.lens + b));   // ragged batch: zero past the utterance's end
        drain256(256, tp.b_in, 1.f, true, live);
        tc_fence_before();
        fence_proxy_async_smem();
        __syncwarp();
        if (tr && warp == 2 && lane == 0) tp.trace[(ti & 7) * 16 + 10] = clock64();
        if (lane == 0 && real) {
          tma_store_4d(&tmXst, sB + (2 * h) * 16384 + q * 4096, (2 * h) * 64, t0 + q * 32, b, 0);
          tma_store_4d(&tmXst, sB + (2 * h + 1) * 16384 + q * 4096, (2 * h + 1) * 64, t0 + q * 32, b, 0);
          tma_store_commit();
        }
      }
    }
    if (lane == 0) tma_store_wait_all();
  }

  tc_fence_before();
  cluster_sync_all();   // the leader's MMAs read the peer's shared memory and write its TMEM
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_2sm<kTmemCols>(tmem_base);
  }
}

}  // namespace pd
