// Fused tail of one denoiser evaluation inside the sampling loop (C = 256, M = 128):
//
//   s2   = relu(W_sp . s1 + b_sp)                  skip_projection + ReLU           wavenet.py:120-121
//   x0   = W_out . s2 + b_out                      output_projection                wavenet.py:122
//   x    = c1 x0 + c2 x + sigma eps                q-posterior sample               prodiff.py:106-121
//   xa   = relu(W_in . x + b_in)                   NEXT step's input_projection     wavenet.py:108-111
//
// as a chain of three tcgen05 GEMMs per 128-frame tile whose intermediate operands never leave the SM: s2 and the 16-bit
// copy of x are written by the epilogue warps straight into 128B-swizzled shared-memory tiles that the next GEMM reads as
// its A operand.  Unfused, these were three launches (skip_proj, output_proj + posterior, input_proj) that moved
// s2 (2 x 32 MB) and the 16-bit x (2 x 16 MB) through HBM and paid three pipeline ramps on 3.4 tiles per SM.
//
// Roles (320 threads, one CTA per SM, persistent over frame tiles): warp 0 = TMA producer, warp 1 = MMA issuer,
// warps 2..9 = epilogue (TMEM lane quadrant = warp % 4, column half h = (warp - 2) / 4).
// Shared memory: a 3-stage ring of {A block [128 x 64] 16 KB, B block [<=256 x 64] 32 KB} + one 64 KB operand tile SB
// (four [128 x 64] sub-tiles) that holds, in turn, s2 (A of GEMM 3), the 16-bit x (A of GEMM 4, sub-tiles 0-1) and the
// staged xa tile the TMA store engine reads.  Every epilogue warp only ever rewrites the rows it stored itself, so the
// only cross-warp ordering is through the mbarriers below.
// TMEM: columns [0, 256) = D2 (skip_proj); [256, 384) = D3 (output_proj); [256, 512) = D4 (input_proj, after D3 has been
// drained) -- so the next tile's skip_proj GEMM runs while the epilogue warps are still draining D4.
// The q-posterior operand c2 x_t + sigma eps needs nothing from this tile's GEMMs: the epilogue warps fetch x_t and run
// Philox for all 64 of their mel bins right after handing s2 to the tensor core, so HBM latency and the RNG arithmetic
// overlap GEMM 3 and each other; the accumulator pass is then one fma per element.
#pragma once
#include "pd_gemm.cuh"

namespace pd {

struct TailParams {
  int B, T, m_tiles_per_row;
  int do_in;                 // 1: also produce the next step's input projection (not on the last step)
  uint32_t idesc256, idesc128;
  const float* b_sp;         // [256]
  const float* b_out;        // [128]
  const float* b_in;         // [256]
  // posterior update, same meaning as EpiArgs' EPI_OUT fields
  int M;
  float* x0_out;             // optional [B,1,M,T]
  float* x_state;            // [B,1,M,T], updated in place
  const float* noise;        // optional [B,1,M,T]; nullptr = in-kernel Philox
  float c1, c2, sigma;
  float out_scale;           // the stored x is multiplied by this (1 except on the last step of a vocoder hand-off call)
  unsigned long long seed;
  unsigned int step_id;
  const int* lens;           // optional per-utterance valid lengths: xa rows past them are written as zeros
  long long* trace;
};

constexpr uint32_t kTailStages = 3;
constexpr uint32_t kTailStageBytes = 16384 + 32768;
constexpr uint32_t kTailSmemBytes = kTailStages * kTailStageBytes + 4 * 16384 + 1024 + 256;

template <typename T>
__global__ void __launch_bounds__(kGemmThreads, 1)
tail_kernel(const __grid_constant__ CUtensorMap tmS1, const __grid_constant__ CUtensorMap tmWsp,
            const __grid_constant__ CUtensorMap tmWout, const __grid_constant__ CUtensorMap tmWin,
            const __grid_constant__ CUtensorMap tmXst, const TailParams tp) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sB = smem + kTailStages * kTailStageBytes;               // 4 x [128 x 64] operand / staging tile
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + 4 * 16384);
  uint64_t* full_bar = bars;                         // [3] TMA -> MMA
  uint64_t* empty_bar = full_bar + kTailStages;      // [3] MMA -> TMA
  uint64_t* acc_bar = empty_bar + kTailStages;       // [3] MMA -> epilogue: D2 / D3 / D4 complete
  uint64_t* s2_bar = acc_bar + 3;                    // epilogue -> MMA: s2 tile written (8 warps)
  uint64_t* xin_bar = s2_bar + 1;                    // epilogue -> MMA: 16-bit x tile written (8 warps)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(xin_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool tr = tp.trace != nullptr && blockIdx.x == 0;
  if (tr && threadIdx.x == 0) tp.trace[15] = clock64();

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmS1);
    tma_prefetch_desc(&tmWsp);
    tma_prefetch_desc(&tmWout);
    tma_prefetch_desc(&tmWin);
    tma_prefetch_desc(&tmXst);
  }
  if (warp == 1 && lane == 0) {
    for (uint32_t i = 0; i < kTailStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 3; ++i) mbar_init(&acc_bar[i], 1);
    mbar_init(s2_bar, kEpiWarps);
    mbar_init(xin_bar, kEpiWarps);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc<kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int num_tiles = tp.B * tp.m_tiles_per_row;
  auto trow_i = [&](uint32_t ti) { return (int)(ti & 3u) + 4 * tp.do_in; };   // trace row: launches with/without GEMM 4 kept apart
  const int n_blocks = tp.do_in ? 10 : 8;   // weight blocks per tile: 4 (skip_proj, with A) + 4 (output_proj) [+ 2 (input_proj)]

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int b = tile / tp.m_tiles_per_row;
        const int t0 = (tile - b * tp.m_tiles_per_row) * kBlockM;
        for (int blk = 0; blk < n_blocks; ++blk, ++it) {
          const uint32_t stage = it % kTailStages, parity = (it / kTailStages) & 1u;
          mbar_wait(&empty_bar[stage], parity ^ 1u, 1);
          uint8_t* sa = smem + stage * kTailStageBytes;
          uint8_t* sb = sa + 16384;
          if (blk < 4) {
            mbar_arrive_expect_tx(&full_bar[stage], 16384 + 32768);
            tma_load_4d(sa, &tmS1, &full_bar[stage], blk * kBlockK, t0, b, 0);
            tma_load_2d(sb, &tmWsp, &full_bar[stage], blk * kBlockK, 0);
          } else if (blk < 8) {
            mbar_arrive_expect_tx(&full_bar[stage], 16384);
            tma_load_2d(sb, &tmWout, &full_bar[stage], (blk - 4) * kBlockK, 0);
          } else {
            mbar_arrive_expect_tx(&full_bar[stage], 32768);
            tma_load_2d(sb, &tmWin, &full_bar[stage], (blk - 8) * kBlockK, 0);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      uint32_t it = 0, ti = 0;
      const uint32_t sBa = smem_u32(sB);
      auto block = [&](uint32_t d_tmem, uint64_t adesc, uint32_t idesc, bool first) {
        const uint32_t stage = it % kTailStages, parity = (it / kTailStages) & 1u;
        mbar_wait(&full_bar[stage], parity, 3);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + stage * kTailStageBytes);
        if (adesc == 0) adesc = make_sw128_kmajor_desc(sa);
        const uint64_t bdesc = make_sw128_kmajor_desc(sa + 16384);
#pragma unroll
        for (int k = 0; k < kBlockK / 16; ++k)
          umma_f16(d_tmem, desc_advance_k(adesc, k), desc_advance_k(bdesc, k), idesc, (!first || k != 0) ? 1u : 0u);
        umma_commit(&empty_bar[stage]);
        ++it;
      };
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++ti) {
        const uint32_t par = ti & 1u;
        for (int kb = 0; kb < 4; ++kb) block(tmem_base, 0, tp.idesc256, kb == 0);                 // D2 = s1 . W_sp^T
        umma_commit(&acc_bar[0]);
        if (tr) tp.trace[trow_i(ti) * 16 + 8] = clock64();
        mbar_wait(s2_bar, par, 4);
        tc_fence_after();
        for (int kb = 0; kb < 4; ++kb)                                                            // D3 = s2 . W_out^T
          block(tmem_base + 256, make_sw128_kmajor_desc(sBa + kb * 16384), tp.idesc128, kb == 0);
        umma_commit(&acc_bar[1]);
        if (tr) tp.trace[trow_i(ti) * 16 + 9] = clock64();
        if (tp.do_in) {
          mbar_wait(xin_bar, par, 5);
          tc_fence_after();
          for (int kb = 0; kb < 2; ++kb)                                                          // D4 = x . W_in^T
            block(tmem_base + 256, make_sw128_kmajor_desc(sBa + kb * 16384), tp.idesc256, kb == 0);
          umma_commit(&acc_bar[2]);
          if (tr) tp.trace[trow_i(ti) * 16 + 10] = clock64();
        }
      }
    }
  } else {
    // ===================== epilogue warps =====================
    const int q = warp & 3, h = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    const int M = tp.M;
    const long long Tn = tp.T;
    uint32_t ti = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++ti) {
      const uint32_t par = ti & 1u;
      const int b = tile / tp.m_tiles_per_row;
      const int t0 = (tile - b * tp.m_tiles_per_row) * kBlockM;
      const int t = t0 + row;
      const bool valid = t < tp.T;
      const long long n = (long long)b * tp.T + t;   // global frame index
      const bool act = valid && (h * 64 < M);
      const long long col0 = ((long long)b * M + h * 64) * Tn + t;   // element (b, mel = 64h, t) of a [B,1,M,T] tensor

      // ---------------- epilogue of D2: s2 = relu(acc + b_sp) -> SB sub-tiles 2h, 2h + 1 (A operand of GEMM 3)
      if (lane == 0) tma_store_wait_read<0>();   // this warp's rows of SB were last read by its own xa store
      __syncwarp();
      mbar_wait(&acc_bar[0], par, 6);
      tc_fence_after();
      if (tr && warp == 2 && lane == 0) tp.trace[trow_i(ti) * 16 + 0] = clock64();
      {
        uint32_t rr[2][32];
        tmem_ld_x32(lane_base + h * 128, rr[0]);
#pragma unroll
        for (int ci = 0; ci < 4; ++ci) {
          const int c = h * 128 + ci * 32;
          tmem_ld_wait();
          if (ci < 3) tmem_ld_x32(lane_base + c + 32, rr[(ci + 1) & 1]);
          const uint32_t (&r)[32] = rr[ci & 1];
          uint8_t* st = sB + (c >> 6) * 16384;
          const int ch0 = (c & 63) >> 3;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint32_t ow[4];
#pragma unroll
            for (int u = 0; u < 4; u += 2) {
              const float4 bv = __ldg(reinterpret_cast<const float4*>(tp.b_sp + c + 8 * j + 2 * u));
              const int col = 8 * j + 2 * u;
              ow[u] = Cvt<T>::pack(fmaxf(__uint_as_float(r[col]) + bv.x, 0.f), fmaxf(__uint_as_float(r[col + 1]) + bv.y, 0.f));
              ow[u + 1] = Cvt<T>::pack(fmaxf(__uint_as_float(r[col + 2]) + bv.z, 0.f), fmaxf(__uint_as_float(r[col + 3]) + bv.w, 0.f));
            }
            *sw128_chunk(st, row, ch0 + j) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
          }
        }
      }
      tc_fence_before();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(s2_bar);
      if (tr && warp == 2 && lane == 0) tp.trace[trow_i(ti) * 16 + 4] = clock64();

      // ---------------- q-posterior operand pv = c2 x_t + sigma eps for this thread's frame and 64 mel bins (prodiff.py:
      // 106-121); independent of the accumulators, so it runs while GEMM 3 does
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

      if (tr && warp == 2 && lane == 0) tp.trace[trow_i(ti) * 16 + 5] = clock64();
      // ---------------- epilogue of D3: x0 = acc + b_out, x = c1 x0 + pv, 16-bit x -> SB sub-tile h (A of GEMM 4)
      mbar_wait(&acc_bar[1], par, 7);
      tc_fence_after();
      if (tr && warp == 2 && lane == 0) tp.trace[trow_i(ti) * 16 + 1] = clock64();
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
          if (tp.do_in && h * 64 < M) {   // rows past T carry zeros: their outputs are clipped by the store anyway
            *sw128_chunk(st, row, 2 * ci) = make_uint4(Cvt<T>::pack(xn[0], xn[1]), Cvt<T>::pack(xn[2], xn[3]), Cvt<T>::pack(xn[4], xn[5]), Cvt<T>::pack(xn[6], xn[7]));
            *sw128_chunk(st, row, 2 * ci + 1) = make_uint4(Cvt<T>::pack(xn[8], xn[9]), Cvt<T>::pack(xn[10], xn[11]), Cvt<T>::pack(xn[12], xn[13]), Cvt<T>::pack(xn[14], xn[15]));
          }
        }
      }
      tc_fence_before();
      if (tp.do_in) {
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(xin_bar);
        if (tr && warp == 2 && lane == 0) tp.trace[trow_i(ti) * 16 + 6] = clock64();

        // ---------------- epilogue of D4: xa = relu(acc + b_in) -> SB sub-tiles 2h, 2h + 1 -> TMA store (layer-0 input)
        mbar_wait(&acc_bar[2], par, 8);
        tc_fence_after();
        if (tr && warp == 2 && lane == 0) tp.trace[trow_i(ti) * 16 + 2] = clock64();
        const bool live = tp.lens == nullptr || t < __ldg(tp.lens + b);   // ragged batch: layer-0 input stays zero past the end
        uint32_t rr[2][32];
        tmem_ld_x32(lane_base + 256 + h * 128, rr[0]);
#pragma unroll
        for (int ci = 0; ci < 4; ++ci) {
          const int c = h * 128 + ci * 32;
          tmem_ld_wait();
          if (ci < 3) tmem_ld_x32(lane_base + 256 + c + 32, rr[(ci + 1) & 1]);
          const uint32_t (&r)[32] = rr[ci & 1];
          uint8_t* st = sB + (c >> 6) * 16384;
          const int ch0 = (c & 63) >> 3;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint32_t ow[4];
#pragma unroll
            for (int u = 0; u < 4; u += 2) {
              const float4 bv = __ldg(reinterpret_cast<const float4*>(tp.b_in + c + 8 * j + 2 * u));
              const int col = 8 * j + 2 * u;
              ow[u] = Cvt<T>::pack(fmaxf(__uint_as_float(r[col]) + bv.x, 0.f), fmaxf(__uint_as_float(r[col + 1]) + bv.y, 0.f));
              ow[u + 1] = Cvt<T>::pack(fmaxf(__uint_as_float(r[col + 2]) + bv.z, 0.f), fmaxf(__uint_as_float(r[col + 3]) + bv.w, 0.f));
            }
            *sw128_chunk(st, row, ch0 + j) = live ? make_uint4(ow[0], ow[1], ow[2], ow[3]) : make_uint4(0u, 0u, 0u, 0u);
          }
        }
        tc_fence_before();
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_4d(&tmXst, sB + (2 * h) * 16384 + q * 4096, (2 * h) * 64, t0 + q * 32, b, 0);       // 32 frames x 64 channels
          tma_store_4d(&tmXst, sB + (2 * h + 1) * 16384 + q * 4096, (2 * h + 1) * 64, t0 + q * 32, b, 0);
          tma_store_commit();
        }
      }
      if (tr && warp == 2 && lane == 0) tp.trace[trow_i(ti) * 16 + 3] = clock64();
    }
    if (lane == 0) tma_store_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  if (tr && threadIdx.x == 0) tp.trace[14] = clock64();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<kTmemCols>(tmem_base);
  }
}

}  // namespace pd
