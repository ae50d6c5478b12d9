// Fused ResidualBlock kernel for sm_100a: one launch per layer does what wavenet.py:60-72 does for a whole batch,
//
//   y      = sigmoid(gate) * tanh(filter),  [gate|filter] = DilConv3(x + d) + cond_proj        (MMA1 + epilogue 1)
//   x_next = (x + W_res y + b_res) / sqrt(2)                                                    (MMA2 + epilogue 2)
//            where "x +" is itself part of MMA2: an identity weight block over the centre view of the resident x tile
//            initialises the accumulator (exact: bf16 x times 1.0 in fp32).  That work runs on the tensor pipe while
//            the epilogue drains the second gate chunk (when the pipe would otherwise idle), and it lets the x tile be
//            released -- and the next one prefetched -- before the epilogue has finished.
//
// and leaves y in HBM for the deferred skip GEMM.  C = 256 residual channels.
//
// B200 mapping
// * A thread-block CLUSTER of 2 CTAs (one SM pair) works on two 128-frame tiles at once with `tcgen05.mma.cta_group::2`
//   (M = 256, N = 256, K = 16): each CTA stages only HALF of every weight block (128 of 256 rows), so weight traffic
//   from L2 and shared-memory operand bandwidth per SM are halved -- the single-CTA kernel was bound by L2->SM feed.
// * The activation tile is loaded ONCE per layer with a halo (128 + 2*dilation frames, TMA zero-fills frames outside
//   the utterance = the conv's padding) and the three taps are row-shifted views of that resident tile (legal with
//   descriptor base_offset 0: the 128B swizzle is a function of absolute smem address bits; profiles/r01_probe_shift.json).
// * The gated activations are written by the epilogue straight into a swizzled K-major shared-memory tile, which is
//   both the A operand of MMA2 and the source of the TMA store to HBM; x_next is staged in the same tile afterwards.
// * TMEM (512 columns) holds two dilated-conv accumulators per tile -- channels 0..127 ("c01") and 128..255 ("c23"), each
//   one N = 256 chunk of 128 gate + 128 filter columns -- and later D2 in c01's columns.  The column assignment mirrors
//   every tile (c01 lands where the previous tile's c23 was).  The gate epilogue of c01 hides behind c23's MMAs; c23's
//   epilogue overlaps the identity part and the first two K blocks of MMA2.  (A 256 + 128 + 128 split exposes half as much
//   epilogue, but N = 128 cta_group::2 MMAs take ~100 cycles instead of 64 -- operand fetch bound -- and two-warp shared
//   slabs need named barriers: measured 5-8 % slower, removed.)
//
// Roles per CTA: warp 0 = weight-block TMA producer, warp 1 = MMA issuer (leader CTA only), warps 2..9 = epilogue,
// warp 10 = x-tile TMA producer.  (Giving the control warps the highest warp ids -- the sub-partition arbiter favours
// high ids -- was measured: MMA issue got faster but the epilogue chain, which is the longer one, got slower; net -4%.)
// Warp (q = warp & 3, h = (warp - 2) >> 2) owns rows 32q..32q+31 of Y sub-tiles h and 2 + h for the whole tile
// (channels [64h, 64h+64) and [128+64h, 128+64h+64)), so every shared-memory hazard on Y is warp-local.
#pragma once
#include "pd_gemm.cuh"

namespace pd {

constexpr int kLayerC = 256;
constexpr int kLayerStages = 5;                  // 16 KB weight blocks in flight per CTA
constexpr int kLayerMaxDil = 16;
constexpr int kLayerThreads = kGemmThreads + 32;  // + warp 10: dedicated X-tile producer (never queued behind weight blocks)
constexpr uint32_t kLayerSmemMax = 227u * 1024u;  // opt-in limit per CTA
// dynamic shared memory of one CTA: haloed x tile (4 sub-tiles of `xsub` bytes) + Y + weight ring (+ resident identity) + barriers + slack
__host__ __device__ inline uint32_t layer_smem_bytes(int xsub, int ident_resident) {
  return 4u * (uint32_t)xsub + 4u * 16384u + kLayerStages * 16384u + (ident_resident ? 4096u : 0u) + 512u + 1024u;
}

// All tensor maps of the kernel, passed by value in parameter space.
struct LayerMaps {
  CUtensorMap xh[2][5];   // haloed loads of the residual stream: [buffer (layer parity)][log2(dilation)], box (64, 128 + 2*dil)
  CUtensorMap xs[2];      // stores into the residual stream buffers, box (64, 32)
  CUtensorMap wd;         // W_dil  [L*2C][3C], 128-channel pairing, box (64, 128)
  CUtensorMap wr;         // W_res  [(L+1)*C][C] (+ identity block), box (64, 128)
  CUtensorMap ys;         // ybuf   [L][B][T][C], box (64, 32) (store)
};

struct LayerParams {
  int B, T, m_tiles_per_row;
  int layer0, n_layers;   // layers [layer0, layer0 + n_layers) are processed by this launch
  int cycle;              // dilation of layer l = 1 << (l % cycle)
  int rot;                // per-layer rotation of the tile-pair -> cluster assignment (multi-layer launches)
  unsigned int* flags;    // multi-layer launches: [L][num_m_tiles] completion counters (8 = x_next of that tile written)
  int ident_row0;         // first row of the [C x C] identity block appended to the W_res matrix
  int xsub;               // bytes reserved per 64-channel sub-tile of the haloed x tile: align_up((128 + 2*max_dil) * 128, 1024)
  int ident_resident;     // 1: the residual add uses a 4 KB identity block kept in shared memory (N = 64 MMAs, no weight
                          //    traffic); 0: identity blocks are streamed through the weight ring like W_res (N = 256)
  uint32_t idesc64;       // N = 64 instruction descriptor for the resident-identity MMAs
  uint32_t idesc;         // cta_group::2 instruction descriptor (M = 256, N = 256)
  const void* cp;         // hoisted cond projection of layer 0, tile-blocked fp16
  long long cp_layer_stride;  // 16-bit elements between layers
  const float* sb;        // step-bias table, layer 0 (layers are 4 * 2C floats apart)
  long long sb_entry_stride;
  int sb_entry_base, sb_entry_per_b;
  const float* b_res;     // residual-half bias of output_projection, [L][C]
  float scale;            // 1/sqrt(2)
  const int* lens;        // optional per-utterance valid lengths (ragged batch, see pd_plan_set_lengths); nullptr = all T
  long long* trace;
};

__device__ __forceinline__ long long global_ns() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
template <typename T>
__global__ void __launch_bounds__(kLayerThreads, 1)
layer_kernel(const __grid_constant__ LayerMaps mp,
             const LayerParams lp) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sX = smem;                                  // 4 x [<=160 frames x 64 ch]   haloed activation tile
  const uint32_t xsub = (uint32_t)lp.xsub;
  uint8_t* sY = sX + 4 * xsub;                         // 4 x [128 frames x 64 ch]     gated y, later x_next
  uint8_t* sW = sY + 4 * 16384;                        // kLayerStages x [128 rows x 64 k] weight blocks (this CTA's half)
  uint8_t* sI = sW + kLayerStages * 16384;             // [32 rows x 64 k] this CTA's half of a 64 x 64 identity (optional)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sI + (lp.ident_resident ? 4096 : 0));
  uint64_t* full_bar = bars;                           // [S] leader: weight block landed in both CTAs
  uint64_t* empty_bar = full_bar + kLayerStages;       // [S] local : MMAs finished reading the block
  uint64_t* xfull_bar = empty_bar + kLayerStages;      // leader: X landed in both CTAs
  uint64_t* xempty_bar = xfull_bar + 1;                // local : X reusable (all tensor-core reads of it completed)
  // accumulators, indexed by kind: 0 = c01 (channels 0..127), 1 = c23 (128..255), 3 = D2 (index 2 unused)
  uint64_t* afull_bar = xempty_bar + 1;                // [4] local : accumulator complete (MMA commit, both CTAs)
  uint64_t* afree_bar = afull_bar + 4;                 // [4] leader: accumulator drained by all 16 epilogue warps
  uint64_t* yready_bar = afree_bar + 4;                // [3] leader: y of chunk c01 / c23 written in both CTAs (index 2 unused)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(yready_bar + 3);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const bool tr = lp.trace != nullptr && blockIdx.x == 0;
  if (tr && threadIdx.x == 0) { lp.trace[15] = clock64(); lp.trace[1008] = global_ns(); }

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mp.wd);
    tma_prefetch_desc(&mp.wr);
    tma_prefetch_desc(&mp.ys);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kLayerStages; ++i) {
      mbar_init(&full_bar[i], 2);
      mbar_init(&empty_bar[i], 1);
    }
    mbar_init(xfull_bar, 2);
    mbar_init(xempty_bar, 1);
    for (int i = 0; i < 4; ++i) {
      mbar_init(&afull_bar[i], 1);
      mbar_init(&afree_bar[i], 2 * kEpiWarps);
    }
    for (int i = 0; i < 3; ++i) mbar_init(&yready_bar[i], 2 * kEpiWarps);
    fence_mbar_init();
  }
  cluster_sync_all();                      // both CTAs alive + barriers initialised before any remote arrive / alloc
  if (warp == 2) tmem_alloc_2sm<kTmemCols>(tmem_slot);
  if (lp.ident_resident && threadIdx.x < 256) {
    // B operand of the residual-add MMAs: rows n = 32 * rank + r of the 64 x 64 identity, K-major, 128B-swizzled
    const int r = threadIdx.x >> 3, c = threadIdx.x & 7;
    const int n = (int)rank * 32 + r;
    uint32_t w[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) w[u] = Cvt<T>::pack(n == 8 * c + 2 * u ? 1.f : 0.f, n == 8 * c + 2 * u + 1 ? 1.f : 0.f);
    *sw128_chunk(sI, r, c) = make_uint4(w[0], w[1], w[2], w[3]);
    fence_proxy_async_smem();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int num_m_tiles = lp.B * lp.m_tiles_per_row;
  const int num_pairs = (num_m_tiles + 1) >> 1;
  const int num_clusters = gridDim.x >> 1;
  const int cluster_id = blockIdx.x >> 1;
  // Every cluster walks the same weight blocks; starting each one at a different K block spreads the simultaneous
  // requests over different L2 lines instead of hammering one hot block (fp32 accumulation order differs per
  // cluster but is fixed for a given launch geometry, so results stay deterministic).
  const int rot12 = cluster_id % 12, rot4 = cluster_id & 3;
  const int l_end = lp.layer0 + lp.n_layers;
  // Multi-layer launches rotate the pair -> cluster assignment every layer, so the clusters that get the extra pair of a
  // partial wave change from layer to layer and the imbalance averages out over the stack.
  auto first_pair = [&](int l) { return (cluster_id + (l - lp.layer0) * lp.rot) % num_clusters; };

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (lane == 0) {
      uint32_t it = 0, ti = 0;
      for (int l = lp.layer0; l < l_end; ++l)
      for (int pair = first_pair(l); pair < num_pairs; pair += num_clusters, ++ti) {
        auto load_block = [&](int blk) {
          const uint32_t stage = it % kLayerStages, parity = (it / kLayerStages) & 1u;
          mbar_wait(&empty_bar[stage], parity ^ 1u, 1);
          if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * 16384);
          else mbar_arrive_leader(&full_bar[stage]);
          uint8_t* dst = sW + stage * 16384;
          if (blk < 12) {          // c01: [128 rows = this CTA's gate (rank 0) or filter (rank 1) half] x 64 K
            const int kb = (blk + rot12) % 12;
            tma_load_2d_2sm(dst, &mp.wd, &full_bar[stage], kb * kBlockK, l * 2 * kLayerC + (int)rank * 128);
          } else if (blk < 24) {   // c23: channels 128..255 as a second N = 256 chunk
            const int kb = (blk - 12 + rot12) % 12;
            tma_load_2d_2sm(dst, &mp.wd, &full_bar[stage], kb * kBlockK, l * 2 * kLayerC + 256 + (int)rank * 128);
          } else if (blk < 28) {   // identity block: D2 starts as the x tile itself
            tma_load_2d_2sm(dst, &mp.wr, &full_bar[stage], ((blk - 24 + rot4) & 3) * kBlockK, lp.ident_row0 + (int)rank * 128);
          } else {
            tma_load_2d_2sm(dst, &mp.wr, &full_bar[stage], (blk - 28) * kBlockK, l * kLayerC + (int)rank * 128);
          }
          if (tr) lp.trace[1024 + (ti & 7) * 128 + blk] = clock64();   // weight-load latency trace: issue time
          ++it;
        };
        // weight blocks per tile: 12 (c01) + 12 (c23) [+ 4 streamed identity] + 4 residual 1x1
        for (int blk = 0; blk < 32; ++blk)
          if (!(lp.ident_resident && blk >= 24 && blk < 28)) load_block(blk);
        if (tr) lp.trace[(ti & 31) * 16 + 1] = clock64();
      }
    }
  } else if (warp == 10) {
    // ===================== x-tile producer (both CTAs) =====================
    if (lane == 0) {
      uint32_t ti = 0;
      for (int l = lp.layer0; l < l_end; ++l) {
        const int di = l % lp.cycle, dil = 1 << di;
        const CUtensorMap* tmX = &mp.xh[l & 1][di];
        const uint32_t xbytes = 4u * (uint32_t)(kBlockM + 2 * dil) * 128u;     // per CTA
        for (int pair = first_pair(l); pair < num_pairs; pair += num_clusters, ++ti) {
          const int m_tile = 2 * pair + (int)rank;
          const int b = m_tile / lp.m_tiles_per_row;           // == B for the dummy tile of an odd count: fully OOB
          const int tr_i = m_tile - b * lp.m_tiles_per_row;    // tile index inside the utterance
          const int t0 = tr_i * kBlockM;
          if (lp.flags != nullptr && l > lp.layer0 && m_tile < num_m_tiles) {
            // The haloed tile of layer l reads x_l of this tile and of its neighbours inside the utterance: wait until the
            // 8 epilogue warps of each producing CTA have published their x_next stores of layer l - 1.  The same three
            // tiles are also the only readers of the buffer region layer l will overwrite, so this wait covers the
            // write-after-read hazard of the two ping-pong buffers as well.
            const unsigned int* fl = lp.flags + (size_t)(l - 1) * num_m_tiles;
            const int lo = tr_i > 0 ? m_tile - 1 : m_tile;
            const int hi = (tr_i + 1 < lp.m_tiles_per_row) ? m_tile + 1 : m_tile;
            for (int nm = lo; nm <= hi; ++nm) {
              long long t_start = clock64();
              while (ld_acquire_u32(fl + nm) < (unsigned)kEpiWarps) {
                __nanosleep(100);
                if (clock64() - t_start > PD_WATCHDOG_CYCLES) {
                  printf("[prodiff_b200] layer-flag watchdog: block %d layer %d tile %d waits for tile %d\n", (int)blockIdx.x, l, m_tile, nm);
                  __trap();
                }
              }
            }
            fence_proxy_async_all();   // order the acquired generic-proxy view before the async-proxy (TMA) reads
          }
          mbar_wait(xempty_bar, (ti & 1u) ^ 1u, 5);
          if (tr) lp.trace[(ti & 31) * 16 + 0] = clock64();
          if (leader) mbar_arrive_expect_tx(xfull_bar, 2 * xbytes);
          else mbar_arrive_leader(xfull_bar);
#pragma unroll
          for (int c = 0; c < 4; ++c) tma_load_4d_2sm(sX + c * xsub, tmX, xfull_bar, c * 64, t0 - dil, b, 0);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (leader && lane == 0) {
      uint32_t it = 0, ti = 0;
      long long wfull = 0;   // trace only: cycles this thread spent waiting for weight blocks in the current tile
      const uint32_t sXa = smem_u32(sX), sYa = smem_u32(sY), sWa = smem_u32(sW);
      for (int l = lp.layer0; l < l_end; ++l)
      for (int pair = first_pair(l); pair < num_pairs; pair += num_clusters, ++ti) {
        const int dil = 1 << (l % lp.cycle);
        const uint32_t par = ti & 1u;                       // this tile's barrier phase parity and column layout
        const uint32_t colA = par ? 256u : 0u;              // c01, later D2 (the layout mirrors every tile)
        const uint32_t colB1 = par ? 0u : 256u;             // c23
        mbar_wait(xfull_bar, par, 2);
        tc_fence_after();
        if (tr) lp.trace[(ti & 31) * 16 + 2] = clock64();
        // ---- MMA1, channels 0..127 (N = 256).  Its columns were the previous tile's c23.
        mbar_wait(&afree_bar[1], par ^ 1u, 3);
        tc_fence_after();
        for (int kk = 0; kk < 12; ++kk, ++it) {
          const int kb = (kk + rot12) % 12;
          const uint32_t stage = it % kLayerStages, parity = (it / kLayerStages) & 1u;
          { const long long w0 = tr ? clock64() : 0; mbar_wait(&full_bar[stage], parity, 4); if (tr) { const long long w1 = clock64(); wfull += w1 - w0; lp.trace[1024 + (ti & 7) * 128 + 32 + kk] = w1; lp.trace[1024 + (ti & 7) * 128 + 64 + kk] = w1 - w0; } }
          tc_fence_after();
          const uint64_t adesc = make_sw128_kmajor_desc(sXa + (kb & 3) * xsub + (uint32_t)((kb >> 2) * dil) * 128u);
          const uint64_t bdesc = make_sw128_kmajor_desc(sWa + stage * 16384);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_f16_2sm(tmem_base + colA, desc_advance_k(adesc, k), desc_advance_k(bdesc, k), lp.idesc, (kk | k) != 0 ? 1u : 0u);
          umma_commit_2sm(&empty_bar[stage]);
        }
        umma_commit_2sm(&afull_bar[0]);
        if (tr) lp.trace[(ti & 31) * 16 + 3] = clock64();
        // ---- MMA1, channels 128..255 (N = 256).  Its columns held the previous tile's D2.
        mbar_wait(&afree_bar[3], par ^ 1u, 5);
        tc_fence_after();
        {
          for (int kk = 0; kk < 12; ++kk, ++it) {
            const int kb = (kk + rot12) % 12;
            const uint32_t stage = it % kLayerStages, parity = (it / kLayerStages) & 1u;
            { const long long w0 = tr ? clock64() : 0; mbar_wait(&full_bar[stage], parity, 6); if (tr) { const long long w1 = clock64(); wfull += w1 - w0; lp.trace[1024 + (ti & 7) * 128 + 32 + 12 + kk] = w1; lp.trace[1024 + (ti & 7) * 128 + 64 + 12 + kk] = w1 - w0; } }
            tc_fence_after();
            const uint64_t adesc = make_sw128_kmajor_desc(sXa + (kb & 3) * xsub + (uint32_t)((kb >> 2) * dil) * 128u);
            const uint64_t bdesc = make_sw128_kmajor_desc(sWa + stage * 16384);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_f16_2sm(tmem_base + colB1, desc_advance_k(adesc, k), desc_advance_k(bdesc, k), lp.idesc, (kk | k) != 0 ? 1u : 0u);
            umma_commit_2sm(&empty_bar[stage]);
          }
          umma_commit_2sm(&afull_bar[1]);
          if (tr) lp.trace[(ti & 31) * 16 + 4] = lp.trace[(ti & 31) * 16 + 5] = clock64();
        }
        // ---- MMA2: D2 = x_centre * I  +  Y * W_res^T, in c01's columns once epilogue(c01) has drained them
        mbar_wait(&afree_bar[0], par, 7);
        tc_fence_after();
        if (tr) lp.trace[512 + (ti & 15) * 16 + 12] = clock64();
        if (lp.ident_resident) {
          // x_centre * I with the resident 64 x 64 identity: four N = 64 column slices, no weight traffic
          const uint64_t bdesc = make_sw128_kmajor_desc(smem_u32(sI));
#pragma unroll
          for (int cc = 0; cc < 4; ++cc) {
            const uint64_t adesc = make_sw128_kmajor_desc(sXa + cc * xsub + (uint32_t)dil * 128u);   // centre rows
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_f16_2sm(tmem_base + colA + cc * 64, desc_advance_k(adesc, k), desc_advance_k(bdesc, k), lp.idesc64, k != 0 ? 1u : 0u);
          }
        }
        for (int kb = lp.ident_resident ? 4 : 0; kb < 8; ++kb) {
          if (kb == 4) {
            umma_commit_2sm(xempty_bar);   // tensor side is done with X: the next tile's X may be prefetched now
            mbar_wait(&yready_bar[0], par, 8);
          }
          if (kb == 6) mbar_wait(&yready_bar[1], par, 8);
          if (tr) lp.trace[512 + (ti & 15) * 16 + 8 + (kb >> 1)] = clock64();
          const uint32_t stage = it % kLayerStages, parity = (it / kLayerStages) & 1u;
          { const long long w0 = tr ? clock64() : 0; mbar_wait(&full_bar[stage], parity, 9); if (tr) { const long long w1 = clock64(); wfull += w1 - w0; lp.trace[1024 + (ti & 7) * 128 + 32 + 24 + kb] = w1; lp.trace[1024 + (ti & 7) * 128 + 64 + 24 + kb] = w1 - w0; } }
          tc_fence_after();
          if (tr) lp.trace[512 + (ti & 15) * 16 + kb] = clock64();
          const uint64_t adesc = kb < 4 ? make_sw128_kmajor_desc(sXa + ((kb + rot4) & 3) * xsub + (uint32_t)dil * 128u)   // centre rows
                                        : make_sw128_kmajor_desc(sYa + (kb - 4) * 16384);
          const uint64_t bdesc = make_sw128_kmajor_desc(sWa + stage * 16384);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_f16_2sm(tmem_base + colA, desc_advance_k(adesc, k), desc_advance_k(bdesc, k), lp.idesc,
                         (lp.ident_resident || (kb | k) != 0) ? 1u : 0u);
          umma_commit_2sm(&empty_bar[stage]);
          ++it;
        }
        umma_commit_2sm(&afull_bar[3]);
        if (tr) { lp.trace[(ti & 31) * 16 + 6] = clock64(); lp.trace[(ti & 31) * 16 + 12] = wfull; wfull = 0; }
      }
    }
  } else {
    // ===================== epilogue warps (both CTAs) =====================
    const int q = warp & 3, h = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    uint32_t ti = 0;
    const uint64_t pol_stream = l2_policy_evict_first();   // cond projection (read once) and gated y (read by a later kernel)
    unsigned int* pending_flag = nullptr;   // completion counter of the x_next slab this warp stored last (lane 0 only)
    for (int l = lp.layer0; l < l_end; ++l)
    for (int pair = first_pair(l); pair < num_pairs; pair += num_clusters, ++ti) {
      const int dil = 1 << (l % lp.cycle);
      const uint16_t* cpb = static_cast<const uint16_t*>(lp.cp) + (long long)l * lp.cp_layer_stride;
      const float* b_res = lp.b_res + l * kLayerC;
      const CUtensorMap* tmXs = &mp.xs[(l + 1) & 1];
      const uint32_t par = ti & 1u;
      const uint32_t colA = par ? 256u : 0u, colB1 = par ? 0u : 256u;
      const int m_tile = 2 * pair + (int)rank;
      const int b = m_tile / lp.m_tiles_per_row;
      const int t0 = (m_tile - b * lp.m_tiles_per_row) * kBlockM;
      const int t = t0 + row;
      const bool valid = (m_tile < num_m_tiles) && (t < lp.T);
      const int entry = lp.sb_entry_base + (b < lp.B ? b : 0) * lp.sb_entry_per_b;
      // step-bias slot of this frame (all taps / no tap 0 / no tap 2 / centre tap only), see k_step_taps
      // ragged batch: the utterance ends at len_b <= T; frames past it are kept at zero in every layer (epilogue 2), so the
      // frames before it see exactly the zero padding a single-utterance run would
      const int len_b = (lp.lens != nullptr && b < lp.B) ? __ldg(lp.lens + b) : lp.T;
      const int slot = (t < dil ? 1 : 0) + (t + dil >= len_b ? 2 : 0);
      const float* sb_row = lp.sb + (long long)entry * lp.sb_entry_stride + (long long)l * (4 * 2 * kLayerC) + slot * (2 * kLayerC);
      const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);

      // ---------------- gate epilogue of a 128-channel chunk (j = 0: channels 0..127; j = 1, two-chunk schedule only:
      // channels 128..255): this warp does channels [128j + 64h, +64) -> Y sub-tile 2j + h
#pragma unroll 1
      for (int j = 0; j < 2; ++j) {
        uint4 pre[16];
        if (valid) {   // hoisted cond projection: 64 gate + 64 filter pre-activations (coalesced, blocked layout)
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            pre[u] = ldg_stream_v4(cpb + blocked_off(m_tile, 64, j * 32 + h * 8 + u, row), pol_stream);
            pre[8 + u] = ldg_stream_v4(cpb + blocked_off(m_tile, 64, j * 32 + 16 + h * 8 + u, row), pol_stream);
          }
        }
        if (valid && j == 0) {   // bring the cond-projection lines of the later sections into L2 now; they are loaded ~6-10k cycles later
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            prefetch_l2(cpb + blocked_off(m_tile, 64, 32 + h * 8 + u, row));
            prefetch_l2(cpb + blocked_off(m_tile, 64, 48 + h * 8 + u, row));
          }
        }
        const uint32_t* prew = reinterpret_cast<const uint32_t*>(pre);
        uint8_t* ytile = sY + (2 * j + h) * 16384;
        // sub-tile h rows of this warp were last read by its own x_next store of the previous tile (newest group).
        // Multi-layer launches publish that store here, one tile late: waiting for its COMPLETION right after issuing it
        // would put the store latency on the epilogue chain, whereas here the warp is about to wait for MMAs anyway.
        if (lane == 0) {
          if (j == 1) {
            tma_store_wait_read<1>();   // sub-tile 2 + h: last read by the same x_next store, now one group older than y(c01)
          } else if (pending_flag != nullptr) {
            tma_store_wait_all();
            fence_proxy_async_all();
            red_release_add_u32(pending_flag, 1u);
            pending_flag = nullptr;
          } else {
            tma_store_wait_read<0>();
          }
        }
        __syncwarp();
        mbar_wait(&afull_bar[j], par, 10);
        tc_fence_after();
        if (tr && warp == 2 && lane == 0) lp.trace[(ti & 31) * 16 + 7 + j] = clock64();
        const uint32_t trow = lane_base + (j ? colB1 : colA);
        const float* sbj = sb_row + j * 256;
        uint32_t rgf[2][2][16];   // the TMEM loads of group ci + 1 are in flight while group ci is gated
        tmem_ld_x16(trow + h * 64, rgf[0][0]);
        tmem_ld_x16(trow + 128 + h * 64, rgf[0][1]);
#pragma unroll
        for (int ci = 0; ci < 4; ++ci) {
          const int i = h * 64 + ci * 16;
          tmem_ld_wait();
          if (ci < 3) {
            tmem_ld_x16(trow + i + 16, rgf[(ci + 1) & 1][0]);
            tmem_ld_x16(trow + 128 + i + 16, rgf[(ci + 1) & 1][1]);
          }
          const uint32_t (&rg)[16] = rgf[ci & 1][0];
          const uint32_t (&rf)[16] = rgf[ci & 1][1];
          uint32_t packed[8];
          gate_chunk16<T>(rg, rf, prew + ci * 8, prew + 32 + ci * 8, sbj + i, packed);
          if (!valid) {
#pragma unroll
            for (int u = 0; u < 8; ++u) packed[u] = 0u;   // keep junk (NaN) out of MMA2's operand rows
          }
          *sw128_chunk(ytile, row, 2 * ci) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
          *sw128_chunk(ytile, row, 2 * ci + 1) = make_uint4(packed[4], packed[5], packed[6], packed[7]);
        }
        tc_fence_before();          // TMEM reads of this accumulator are done
        fence_proxy_async_smem();   // Y rows visible to the tensor core (MMA2) and the TMA store engine
        __syncwarp();
        if (tr && warp == 2 && lane == 0) lp.trace[(ti & 31) * 16 + (j ? 13 : 9)] = clock64();   // section end
        if (lane == 0) {
          mbar_arrive_leader(&afree_bar[j]);
          mbar_arrive_leader(&yready_bar[j]);
          tma_store_4d_hint(&mp.ys, ytile + q * 4096, (2 * j + h) * 64, t0 + q * 32, b, l, pol_stream);   // 32 frames x 64 channels
          tma_store_commit();
        }
      }

      // ---------------- epilogue 2: x_next = (D2 + b) / sqrt(2) (D2 already contains x), staged in Y, stored by TMA.
      // Warp (q, h) handles channels of sub-tiles h and 2 + h, whose earlier TMA stores it issued itself.
      {
        if (lane == 0) tma_store_wait_read<0>();
        __syncwarp();
        mbar_wait(&afull_bar[3], par, 12);
        tc_fence_after();
        if (tr && warp == 2 && lane == 0) lp.trace[(ti & 31) * 16 + 10] = clock64();
        const uint32_t trow = lane_base + colA;
        const float sc2 = t < len_b ? lp.scale : 0.f;
        uint32_t rr[2][32];   // the TMEM load of group i+1 is in flight while group i is converted and stored
        tmem_ld_x32(trow + h * 64, rr[0]);
#pragma unroll
        for (int jj = 0; jj < 2; ++jj) {
          uint8_t* ytile = sY + (2 * jj + h) * 16384;
#pragma unroll
          for (int ci = 0; ci < 2; ++ci) {
            const int gi = jj * 2 + ci;
            const int c = (2 * jj + h) * 64 + ci * 32;   // first channel of this 32-column group
            tmem_ld_wait();
            if (gi < 3) tmem_ld_x32(trow + (2 * ((gi + 1) >> 1) + h) * 64 + ((gi + 1) & 1) * 32, rr[(gi + 1) & 1]);
            const uint32_t (&r)[32] = rr[gi & 1];
#pragma unroll
            for (int u4 = 0; u4 < 4; ++u4) {
              uint32_t ow[4];
#pragma unroll
              for (int u = 0; u < 4; u += 2) {
                const float4 bv = __ldg(reinterpret_cast<const float4*>(b_res + c + 8 * u4 + 2 * u));
                const int col = 8 * u4 + 2 * u;
                ow[u] = Cvt<T>::pack((__uint_as_float(r[col]) + bv.x) * sc2, (__uint_as_float(r[col + 1]) + bv.y) * sc2);
                ow[u + 1] = Cvt<T>::pack((__uint_as_float(r[col + 2]) + bv.z) * sc2, (__uint_as_float(r[col + 3]) + bv.w) * sc2);
              }
              *sw128_chunk(ytile, row, ci * 4 + u4) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
            }
          }
        }
        tc_fence_before();
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive_leader(&afree_bar[3]);
          tma_store_4d(tmXs, sY + h * 16384 + q * 4096, h * 64, t0 + q * 32, b, 0);
          tma_store_4d(tmXs, sY + (2 + h) * 16384 + q * 4096, (2 + h) * 64, t0 + q * 32, b, 0);
          tma_store_commit();
          if (lp.flags != nullptr && m_tile < num_m_tiles) pending_flag = lp.flags + (size_t)l * num_m_tiles + m_tile;
        }
        if (tr && warp == 2 && lane == 0) lp.trace[(ti & 31) * 16 + 11] = clock64();
      }
    }
    if (lane == 0) {
      tma_store_wait_all();
      if (pending_flag != nullptr) {
        fence_proxy_async_all();
        red_release_add_u32(pending_flag, 1u);
      }
    }
  }

  // both CTAs must be completely done (the leader's MMAs read the peer's shared memory and write its TMEM)
  tc_fence_before();
  cluster_sync_all();
  if (tr && threadIdx.x == 0) { lp.trace[14] = clock64(); lp.trace[1009] = global_ns(); }
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_2sm<kTmemCols>(tmem_base);
  }
}

}  // namespace pd
