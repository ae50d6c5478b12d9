// Persistent, warp-specialised implicit-GEMM kernel for sm_100a (tcgen05 + TMEM + TMA).
//
// Every dense contraction on the DiffNet hot path is an instance of this one kernel:
//
//   D[128 frames, BN channels] = sum over K-blocks of  A_kb[128 frames, 64] * B_kb[BN, 64]^T       (fp32 in TMEM)
//
// * A is a channels-last activation tensor [slab][B][T][C] (16-bit) read through a 4-D TMA tensor map with a
//   (64 ch, 128 frames, 1, 1) box.  For the dilated convolution the K-blocks walk (tap, channel-chunk) and the
//   frame coordinate of the box is shifted by (tap-1)*dilation; frames before 0 / after T-1 are zero-filled by
//   TMA, which reproduces the reference's per-utterance zero padding (wavenet.py:55, padding=dilation).
//   For the deferred skip sum the K-blocks walk (layer slab, channel-chunk) instead.
// * B is a packed weight matrix [rows][K] (16-bit, K-major) read through a 2-D map with a (64, BN) box.
// * warp 0 = TMA producer, warp 1 = MMA issuer (one elected lane each), warps 2..9 = epilogue
//   (TMEM -> registers -> fused epilogue -> global).  Two TMEM accumulator stages let the MMAs of tile i+1 overlap
//   the epilogue of tile i; a kStages-deep shared-memory ring overlaps TMA with MMA.
//
// The epilogues are where the reference's elementwise graph is fused (wavenet.py:60-72,108-122; prodiff.py:106-121):
//   EPI_STORE : (acc + bias) * scale, optional ReLU             (cond / input / skip projections)
//   EPI_GATE  : sigmoid(gate) * tanh(filter) with the hoisted condition projection and the edge-aware
//               diffusion-step bias added first                 (ResidualBlock up to the gate)
//   EPI_RES   : (x + acc + bias) / sqrt(2)                      (ResidualBlock residual path)
//   EPI_OUT   : x0 = acc + bias, then the q-posterior update x_{t-1} = c1*x0 + c2*x_t + sigma*noise
//   EPI_STORE_X : EPI_STORE plus what the text encoder's GEMMs need (pd_encoder.cuh): erf-GELU, an fp32 residual added
//               after the activation, and a per-row keep mask (padding tokens come out as zeros); EPI_STORE_XH = the same
//               for 16-bit outputs without a residual (smaller staging tiles, one more ring stage for the K = 2304 FFN conv)
#pragma once
#include "pd_ptx.cuh"

namespace pd {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;  // 16-bit elements per K-block = one 128-byte swizzle atom
constexpr int kEpiWarps = 8;
constexpr int kGemmThreads = 32 * (2 + kEpiWarps);
constexpr uint32_t kTmemCols = 512;

enum EpiKind { EPI_STORE = 0, EPI_GATE = 1, EPI_RES = 2, EPI_OUT = 3, EPI_STORE_BLK = 4, EPI_STORE_X = 5, EPI_STORE_XH = 6 };
enum OutType { OUT_T = 0, OUT_F16 = 1, OUT_F32 = 2 };

struct GemmShape {
  int B, T;             // utterances, frames per utterance
  int m_tiles_per_row;  // ceil(T / 128)
  int n_tiles;          // N_total / BN
  int k_per_tap;        // 64-wide K-blocks per tap (or per slab)
  int taps;             // conv taps (3), slabs (L) or 1
  int tap_mode;         // 1: K-block group = conv tap (frame shift); 0: K-block group = A slab
  int dil;              // dilation
  int a_slab0;          // base 4th coordinate of A
  int b_row0;           // first weight row of n-tile 0
  uint32_t idesc;       // tcgen05 instruction descriptor (M=128, N=BN, operand format)
  long long* trace;     // debug: CTA 0 records clock64() at pipeline events (nullptr in production)
};

struct EpiArgs {
  void* out;
  const void* aux;    // GATE: hoisted cond projection (fp16) of this layer; RES: x_old (T)
  const float* bias;  // per output column (fp32)
  const float* sb;    // GATE: step-bias table of this layer, [entry][4][2C] = {all taps, no tap 0, no tap 2, tap 1 only}
  long long sb_entry_stride;
  int sb_entry_base, sb_entry_per_b;
  int ld;       // STORE: slab width of the output; GATE/RES: C
  int out_blocked;  // STORE: write the tile-blocked layout (see blocked_off) instead of row-major
  int out_slab;     // GATE: 4th TMA coordinate (layer) of the gated-activation store
  float scale;  // STORE: multiplier after bias; RES: 1/sqrt(2)
  const int* lens;  // optional per-utterance valid lengths (ragged batch): STORE writes zeros for frames >= lens[b]
  int relu;
  int out_type;
  // EPI_STORE_X
  int gelu;               // exact (erf) GELU after bias and scale (common_layers.py:575-576)
  const float* resid;     // optional fp32 tensor in the output's own layout, added last (may alias `out`)
  const uint8_t* keep;    // optional per-row flag [B*T]: rows with 0 are written as zeros
  // EPI_OUT
  int M, Mp;
  float* x0_out;       // optional [B,1,M,T]
  float* x_state;      // optional [B,1,M,T], updated in place
  const float* noise;  // optional [B,1,M,T]
  void* xin_next;      // optional [B*T, Mp] 16-bit copy of x_{t-1} (next step's input-projection operand)
  float c1, c2, sigma;
  float out_scale;     // the stored x is multiplied by this (pd_plan_set_output_scale; 1 otherwise)
  unsigned long long seed;
  unsigned int step_id;
};

// ------------------------------------------------------------------ small device helpers
template <typename T> struct Cvt;
template <> struct Cvt<__nv_bfloat16> {
  static __device__ __forceinline__ uint32_t pack(float a, float b) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
  }
  static __device__ __forceinline__ float2 unpack(uint32_t u) {
    return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&u));
  }
  static constexpr uint32_t kFormat = 1;  // tcgen05 kind::f16 operand format: 1 = BF16
};
template <> struct Cvt<__half> {
  static __device__ __forceinline__ uint32_t pack(float a, float b) {
    __half2 v = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
  }
  static __device__ __forceinline__ float2 unpack(uint32_t u) { return __half22float2(*reinterpret_cast<__half2*>(&u)); }
  static constexpr uint32_t kFormat = 0;  // 0 = F16
};

// sigmoid(g) * tanh(f).
//  * fp16 operands (strict-parity mode): (e^{2f} - 1) / ((1 + e^{-g}) (e^{2f} + 1)), 2 ex2 + 1 rcp, ~1e-6 accurate.
//  * bf16 operands: MUFU.TANH twice (sigmoid(g) = 0.5 tanh(g/2) + 0.5): 5 issue slots instead of 14.  tanh.approx has
//    ~2^-11 max error, below the 2^-9 rounding of the bf16 value the result is stored as.  (tanh.approx.f16x2 was
//    tried: ptxas splits it into two scalar MUFU.TANH.F16, so it saves nothing.)
template <typename T>
__device__ __forceinline__ float gate_fn(float g, float f) {
  if constexpr (sizeof(T) == 2 && Cvt<T>::kFormat == 1) {
    return fmaf(0.5f, tanh_approx(0.5f * g), 0.5f) * tanh_approx(f);
  } else {
    f = fminf(fmaxf(f, -15.f), 15.f);
    g = fmaxf(g, -30.f);
    const float e2f = __expf(2.f * f);
    const float eg = __expf(-g);
    return __fdividef(e2f - 1.f, (1.f + eg) * (e2f + 1.f));
  }
}

// Tile-blocked layout for tensors that only epilogues touch (the hoisted condition projection): within one 128-frame
// tile the 16-byte chunks of a row are strided by 128 rows, i.e. [tile][chunk][row][8 x 16-bit].  With one thread per
// row, a warp's 16-byte accesses to one chunk are 512 contiguous bytes (4 wavefronts) instead of 32 scattered lines.
__device__ __forceinline__ long long blocked_off(long long tile, int chunks_per_row, int chunk, int row) {
  return ((tile * chunks_per_row + chunk) * 128 + row) * 8;  // in 16-bit elements
}

// Philox4x32-10 (counter-based RNG) + Box-Muller: 4 standard normals per call.
__device__ __forceinline__ uint4 philox4x32(uint4 ctr, uint2 key) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const unsigned long long p0 = (unsigned long long)0xD2511F53u * ctr.x;   // one IMAD.WIDE each instead of mul.hi + mul.lo
    const unsigned long long p1 = (unsigned long long)0xCD9E8D57u * ctr.z;
    const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
    const uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += 0x9E3779B9u;
    key.y += 0xBB67AE85u;
  }
  return ctr;
}
__device__ __forceinline__ float4 philox_normal4(unsigned long long seed, uint32_t step_id, uint32_t c0, uint32_t c1) {
  const uint4 r = philox4x32(make_uint4(c0, c1, step_id, 0u), make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  const float k = 2.3283064365386963e-10f;  // 2^-32
  const float u0 = ((float)r.x + 1.0f) * k, u1 = (float)r.y * k;
  const float u2 = ((float)r.z + 1.0f) * k, u3 = (float)r.w * k;
  const float r0 = sqrt_approx(-2.f * __logf(fminf(u0, 1.f))), r1 = sqrt_approx(-2.f * __logf(fminf(u2, 1.f)));
  float s0, c0f, s1, c1f;
  __sincosf(6.283185307179586f * u1, &s0, &c0f);
  __sincosf(6.283185307179586f * u3, &s1, &c1f);
  return make_float4(r0 * c0f, r0 * s0, r1 * c1f, r1 * s1);
}

// One 16-channel step of the gate epilogue: z = acc + cond_proj + step bias, y = sigmoid(z_gate) * tanh(z_filter), packed
// to 16-bit.  rg/rf: accumulator columns of the 16 gate / filter channels; cpg/cpf: 8 packed fp16 pairs each; sb: THIS
// FRAME's step-bias row (the slot for interior / first / last frames is chosen per lane by the caller) positioned at
// the first gate channel, filter at +128.  One code path for every frame: no edge branch, no predication.
template <typename T>
__device__ __forceinline__ void gate_chunk16(const uint32_t (&rg)[16], const uint32_t (&rf)[16], const uint32_t* cpg,
                                             const uint32_t* cpf, const float* sb, uint32_t (&packed)[8]) {
#pragma unroll
  for (int j = 0; j < 16; j += 4) {
    const float4 bg = __ldg(reinterpret_cast<const float4*>(sb + j));
    const float4 bf = __ldg(reinterpret_cast<const float4*>(sb + 128 + j));
    const float zg[4] = {bg.x, bg.y, bg.z, bg.w}, zf[4] = {bf.x, bf.y, bf.z, bf.w};
#pragma unroll
    for (int u = 0; u < 4; u += 2) {
      const float2 pg = Cvt<__half>::unpack(cpg[(j + u) >> 1]);
      const float2 pf = Cvt<__half>::unpack(cpf[(j + u) >> 1]);
      const float g0 = __uint_as_float(rg[j + u]) + (pg.x + zg[u]);
      const float g1 = __uint_as_float(rg[j + u + 1]) + (pg.y + zg[u + 1]);
      const float f0 = __uint_as_float(rf[j + u]) + (pf.x + zf[u]);
      const float f1 = __uint_as_float(rf[j + u + 1]) + (pf.y + zf[u + 1]);
      packed[(j + u) >> 1] = Cvt<T>::pack(gate_fn<T>(g0, f0), gate_fn<T>(g1, f1));
    }
  }
}

// Exact (erf) GELU, 0.5 v (1 + erf(v / sqrt 2)), with erf from Abramowitz & Stegun 7.1.26 (|error| <= 1.5e-7, branch-free:
// one reciprocal, one exp, five fma) -- far below the 16-bit rounding of the activation it produces.  libdevice's erff made
// the FFN-conv GEMM of the text encoder epilogue-bound (128 values per thread per tile, ~35 k cycles against 9 k of MMAs).
__device__ __forceinline__ float gelu_erf(float v) {
  const float x = fabsf(v) * 0.70710678118654752f;
  const float t = __fdividef(1.f, fmaf(0.3275911f, x, 1.f));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  const float e = 1.f - p * t * __expf(-x * x);
  return 0.5f * v * (1.f + copysignf(e, v));
}

// kPair: two CTAs of a cluster compute two adjacent 128-frame tiles with tcgen05.mma.cta_group::2 (M = 256); each CTA
// stages its own A tile and only HALF of the weight block, which halves L2 -> SM weight traffic (the single-CTA form of
// cond_proj / skip_sum re-streams the weights per 128-row tile and sits at the L2 throughput cap).
template <int EPI, int BN, bool kPair = false> struct GemmCfg {
  static constexpr uint32_t kABytes = kBlockM * kBlockK * 2;
  static constexpr uint32_t kBBytes = (kPair ? BN / 2 : BN) * kBlockK * 2;
  static constexpr uint32_t kStageBytes = kABytes + kBBytes;
  static constexpr uint32_t kStagingBytes = (EPI == EPI_STORE_BLK) ? 65536 : 0;   // blocked-output tile staged for a bulk store
  static constexpr uint32_t kXRowWords = (EPI == EPI_STORE_XH) ? 20 : 36;                        // 16-bit only / fp32 rows (padded)
  static constexpr uint32_t kXStageBytes = (EPI == EPI_STORE_X || EPI == EPI_STORE_XH) ? kEpiWarps * 32 * kXRowWords * 4 : 0;    // per-warp 32 x 32 transposition tiles
  static constexpr uint32_t kExtraBytes = ((EPI == EPI_GATE) ? 2 * 16384 : (EPI == EPI_RES ? 4 * 16384 : 0)) + kStagingBytes + kXStageBytes;
  // smem ring depth: as many stages as fit next to the epilogue's operand/output tile (cap 8)
  static constexpr int kStagesFit = (int)((227u * 1024u - 1024u - 256u - kExtraBytes) / kStageBytes);
  static constexpr int kStagesWant = kPair ? 8 : ((EPI == EPI_RES) ? 3 : (BN == 256 ? 4 : 6));
  static constexpr int kStages = kStagesFit < kStagesWant ? kStagesFit : kStagesWant;
  static constexpr uint32_t kSmemBytes = kStages * kStageBytes + kExtraBytes + 1024 + 256;  // + alignment slack + barriers
};

// address of the 16-byte chunk `c` (0..7) of row `r` in a [128 x 64] 16-bit tile laid out with the 128-byte swizzle
__device__ __forceinline__ uint4* sw128_chunk(uint8_t* tile, int r, int c) {
  return reinterpret_cast<uint4*>(tile + r * 128 + ((c ^ (r & 7)) << 4));
}

// ------------------------------------------------------------------ the kernel
// tmA: A operand (activations); tmB: B operand (weights); tmO: TMA-store target (GATE: gated activations, RES: x_next);
// tmX: RES only, x_cur tile load.  Unused maps are passed as copies of tmA.
template <int EPI, int BN, typename T, bool kPair = false>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
            const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmX, const GemmShape sh,
            const EpiArgs ep) {
  using Cfg = GemmCfg<EPI, BN, kPair>;
  static_assert(EPI != EPI_STORE_BLK || BN == 256, "blocked-output staging is 256 columns wide");
  constexpr int kStages = Cfg::kStages;
  static_assert(BN == 256 || BN == 128, "tile width");
  static_assert(EPI != EPI_GATE || BN == 256, "gate epilogue pairs 128 gate with 128 filter columns");
  static_assert(EPI != EPI_RES || BN == 256, "residual epilogue is 256 wide");
  static_assert(EPI != EPI_OUT || BN == 128, "out epilogue is 128 wide");

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);  // SWIZZLE_128B needs 1024-byte alignment
  uint8_t* extra = smem + kStages * Cfg::kStageBytes;                   // GATE: y staging, RES: x tile
  uint64_t* bars = reinterpret_cast<uint64_t*>(extra + Cfg::kExtraBytes);
  uint64_t* full_bar = bars;                  // [kStages]  TMA -> MMA
  uint64_t* empty_bar = bars + kStages;       // [kStages]  MMA -> TMA
  uint64_t* tfull_bar = bars + 2 * kStages;   // [2]        MMA -> epilogue
  uint64_t* tempty_bar = tfull_bar + 2;       // [2]        epilogue -> MMA
  uint64_t* xfull_bar = tempty_bar + 2;       // RES: x tile landed
  uint64_t* xempty_bar = xfull_bar + 1;       // RES: x tile stored, buffer reusable
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(xempty_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = kPair ? cluster_ctarank() : 0u;
  const bool leader = rank == 0;
  if (sh.trace && blockIdx.x == 0 && threadIdx.x == 0) sh.trace[15] = clock64();  // kernel entry

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if constexpr (EPI == EPI_GATE || EPI == EPI_RES) tma_prefetch_desc(&tmO);
    if constexpr (EPI == EPI_RES) tma_prefetch_desc(&tmX);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], kPair ? 2 : 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], kPair ? 2 * kEpiWarps : kEpiWarps);
    }
    mbar_init(xfull_bar, 1);   // RES: x tile landed
    mbar_init(xempty_bar, 1);
    fence_mbar_init();
  }
  if constexpr (kPair) {
    cluster_sync_all();  // both CTAs alive + barriers initialised before any remote arrive / paired allocation
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
  if constexpr (EPI == EPI_STORE_X || EPI == EPI_STORE_XH) {
    // the encoder chains ~30 short launches: let the next one run its prologue while this grid drains, and do not touch the
    // previous kernel's output before it has completed (pd_ptx.cuh, programmatic dependent launch)
    griddep_launch();
    griddep_wait();
  }

  const int num_m_tiles = sh.B * sh.m_tiles_per_row;
  // work items: (m-tile, n-tile), or (pair of adjacent m-tiles, n-tile) with this CTA taking m-tile 2 * pair + rank
  const int num_tiles = (kPair ? (num_m_tiles + 1) / 2 : num_m_tiles) * sh.n_tiles;
  const int n_units = kPair ? (int)(gridDim.x >> 1) : (int)gridDim.x;   // CTAs or CTA pairs
  const int unit = kPair ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int tile_first = unit, tile_step = n_units, tile_end = num_tiles;
  const int k_blocks = sh.k_per_tap * sh.taps;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      uint32_t it = 0, ti = 0;
      for (int tile = tile_first; tile < tile_end; tile += tile_step, ++ti) {
        const int n_tile = tile % sh.n_tiles;
        const int m_tile = kPair ? 2 * (tile / sh.n_tiles) + (int)rank : tile / sh.n_tiles;
        const int b = m_tile / sh.m_tiles_per_row;
        const int t0 = (m_tile % sh.m_tiles_per_row) * kBlockM;
        if constexpr (EPI == EPI_RES) {
          // x_cur tile for the residual add: [128 frames x 256 ch] as four swizzled [128 x 64] sub-tiles
          mbar_wait(xempty_bar, (ti & 1u) ^ 1u, 5);
          mbar_arrive_expect_tx(xfull_bar, 4 * 16384);
#pragma unroll
          for (int j = 0; j < 4; ++j) tma_load_4d(extra + j * 16384, &tmX, xfull_bar, n_tile * BN + j * 64, t0, b, 0);
        }
        for (int kb = 0; kb < k_blocks; ++kb, ++it) {
          const uint32_t stage = it % kStages, parity = (it / kStages) & 1u;
          mbar_wait(&empty_bar[stage], parity ^ 1u, 1);
          if (sh.trace && blockIdx.x == 0 && (kb == 0 || kb == k_blocks - 1)) sh.trace[ti * 16 + (kb == 0 ? 0 : 1)] = clock64();
          uint8_t* sa = smem + stage * Cfg::kStageBytes;
          uint8_t* sb = sa + Cfg::kABytes;
          const int grp = kb / sh.k_per_tap, kc = kb - grp * sh.k_per_tap;
          const int t = sh.tap_mode ? t0 + (grp - (sh.taps >> 1)) * sh.dil : t0;
          const int slab = sh.tap_mode ? sh.a_slab0 : sh.a_slab0 + grp;
          if constexpr (kPair) {
            if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * Cfg::kStageBytes);
            else mbar_arrive_leader(&full_bar[stage]);
            tma_load_4d_2sm(sa, &tmA, &full_bar[stage], kc * kBlockK, t, b, slab);
            tma_load_2d_2sm(sb, &tmB, &full_bar[stage], kb * kBlockK, sh.b_row0 + n_tile * BN + (int)rank * (BN / 2));
          } else {
            mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStageBytes);
            tma_load_4d(sa, &tmA, &full_bar[stage], kc * kBlockK, t, b, slab);
            tma_load_2d(sb, &tmB, &full_bar[stage], kb * kBlockK, sh.b_row0 + n_tile * BN);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (pair mode: leader CTA only) =====================
    if (lane == 0 && leader) {
      uint32_t it = 0, ti = 0;
      for (int tile = tile_first; tile < tile_end; tile += tile_step, ++ti) {
        const uint32_t as = ti & 1u, aph = (ti >> 1) & 1u;
        mbar_wait(&tempty_bar[as], aph ^ 1u, 2);
        tc_fence_after();
        if (sh.trace && blockIdx.x == 0) sh.trace[ti * 16 + 2] = clock64();
        const uint32_t d_tmem = tmem_base + as * BN;
        for (int kb = 0; kb < k_blocks; ++kb, ++it) {
          const uint32_t stage = it % kStages, parity = (it / kStages) & 1u;
          mbar_wait(&full_bar[stage], parity, 3);
          tc_fence_after();
          if (sh.trace && blockIdx.x == 0 && (kb == 0 || kb == k_blocks - 1)) sh.trace[ti * 16 + (kb == 0 ? 3 : 4)] = clock64();
          const uint32_t sa = smem_u32(smem + stage * Cfg::kStageBytes);
          const uint64_t adesc = make_sw128_kmajor_desc(sa);
          const uint64_t bdesc = make_sw128_kmajor_desc(sa + Cfg::kABytes);
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k) {
            if constexpr (kPair) umma_f16_2sm(d_tmem, desc_advance_k(adesc, k), desc_advance_k(bdesc, k), sh.idesc, (kb | k) != 0 ? 1u : 0u);
            else umma_f16(d_tmem, desc_advance_k(adesc, k), desc_advance_k(bdesc, k), sh.idesc, (kb | k) != 0 ? 1u : 0u);
          }
          // smem slot reusable once these MMAs have read it (pair mode: in both CTAs)
          if constexpr (kPair) umma_commit_2sm(&empty_bar[stage]);
          else umma_commit(&empty_bar[stage]);
        }
        // accumulator complete
        if constexpr (kPair) umma_commit_2sm(&tfull_bar[as]);
        else umma_commit(&tfull_bar[as]);
      }
    }
  } else {
    // ===================== epilogue warps =====================
    const int q = warp & 3;          // TMEM lane quadrant this warp may access
    const int h = (warp - 2) >> 2;   // which half of the tile's columns this warp handles
    const int row = q * 32 + lane;
    const bool issuer = (warp == 2 && lane == 0);  // the one thread that issues TMA stores
    uint32_t ti = 0;
    for (int tile = tile_first; tile < tile_end; tile += tile_step, ++ti) {
      const int n_tile = tile % sh.n_tiles;
      const int m_tile = kPair ? 2 * (tile / sh.n_tiles) + (int)rank : tile / sh.n_tiles;
      const int b = m_tile / sh.m_tiles_per_row;
      const int t0 = (m_tile % sh.m_tiles_per_row) * kBlockM;
      const int t = t0 + row;
      const bool valid = (t < sh.T) && (m_tile < num_m_tiles);   // pair mode: an odd tile count leaves one dummy tile
      const long long n = (long long)b * sh.T + t;  // global frame index
      const uint32_t as = ti & 1u, aph = (ti >> 1) & 1u;
      const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + as * BN;
      // Global operands of the epilogue do not depend on the accumulator: issue their loads BEFORE waiting for the
      // MMAs of this tile, so HBM latency hides under the tensor work instead of serialising after it.
      uint4 pre[16];
      if constexpr (EPI == EPI_GATE) {
        if (valid) {
          const uint16_t* cpb = static_cast<const uint16_t*>(ep.aux);
          const int cpr = (2 * ep.ld) >> 3;             // 16-byte chunks per row of the cond projection
          const int cg = n_tile * 32 + h * 8;           // first gate chunk of this warp's 64 columns
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            pre[u] = __ldg(reinterpret_cast<const uint4*>(cpb + blocked_off(m_tile, cpr, cg + u, row)));
            pre[8 + u] = __ldg(reinterpret_cast<const uint4*>(cpb + blocked_off(m_tile, cpr, cg + 16 + u, row)));
          }
        }
      }
      if constexpr (EPI == EPI_OUT) {
        // x_t of the first 16 mel bins this warp handles (the other chunks are fetched one chunk ahead, below)
        if (valid && ep.x_state != nullptr && h * 64 < ep.M) {
          float* pfw = reinterpret_cast<float*>(pre);
          const float* xs = ep.x_state + ((long long)b * ep.M + h * 64) * sh.T + t;
#pragma unroll
          for (int j = 0; j < 16; ++j) pfw[j] = xs[(long long)j * sh.T];
        }
      }
      if (sh.trace && blockIdx.x == 0 && warp == 2 && lane == 0) sh.trace[ti * 16 + 5] = clock64();
      if constexpr (EPI == EPI_RES) mbar_wait(xfull_bar, ti & 1u, 6);
      mbar_wait(&tfull_bar[as], aph, 4);
      tc_fence_after();
      if (sh.trace && blockIdx.x == 0 && warp == 2 && lane == 0) sh.trace[ti * 16 + 6] = clock64();

      if constexpr (EPI == EPI_STORE_BLK) {
        // Tile-blocked 16-bit output ([tile][16-byte chunk][128 rows][8], see blocked_off): the (m-tile, n-tile) block of this
        // column half is ONE contiguous 32 KB range in global memory, so it is staged in the same order in shared memory
        // (conflict-free 16-byte stores, one thread per row) and leaves as a single bulk copy.  Direct st.global from the
        // epilogue threads made this kernel epilogue-bound (6.7 k cycles per tile against 2 k of MMA).
        uint8_t* stg = extra + h * 32768;
        const bool store_thread = (q == 0 && lane == 0);
        if (store_thread) tma_store_wait_read<0>();   // the previous tile's copy has finished reading the staging half
        named_bar_sync(1 + h, 4 * 32);
        uint32_t rr[2][32];
        tmem_ld_x32(trow + h * 128, rr[0]);
#pragma unroll
        for (int ci = 0; ci < 4; ++ci) {
          const int c = h * 128 + ci * 32;
          tmem_ld_wait();
          if (ci < 3) tmem_ld_x32(trow + c + 32, rr[(ci + 1) & 1]);
          const uint32_t (&r)[32] = rr[ci & 1];
          const int g = n_tile * BN + c;
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 bv = __ldg(reinterpret_cast<const float4*>(ep.bias + g + j));
            v[j + 0] = (__uint_as_float(r[j + 0]) + bv.x) * ep.scale;
            v[j + 1] = (__uint_as_float(r[j + 1]) + bv.y) * ep.scale;
            v[j + 2] = (__uint_as_float(r[j + 2]) + bv.z) * ep.scale;
            v[j + 3] = (__uint_as_float(r[j + 3]) + bv.w) * ep.scale;
          }
          if (ep.relu) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint4 o = ep.out_type == OUT_F16
                                ? make_uint4(Cvt<__half>::pack(v[8 * j], v[8 * j + 1]), Cvt<__half>::pack(v[8 * j + 2], v[8 * j + 3]),
                                             Cvt<__half>::pack(v[8 * j + 4], v[8 * j + 5]), Cvt<__half>::pack(v[8 * j + 6], v[8 * j + 7]))
                                : make_uint4(Cvt<T>::pack(v[8 * j], v[8 * j + 1]), Cvt<T>::pack(v[8 * j + 2], v[8 * j + 3]),
                                             Cvt<T>::pack(v[8 * j + 4], v[8 * j + 5]), Cvt<T>::pack(v[8 * j + 6], v[8 * j + 7]));
            *reinterpret_cast<uint4*>(stg + (ci * 4 + j) * 2048 + row * 16) = o;
          }
        }
        fence_proxy_async_smem();
        named_bar_sync(3 + h, 4 * 32);
        if (store_thread && m_tile < num_m_tiles) {
          const int g0 = n_tile * BN + h * 128;
          const int slab = g0 / ep.ld, within0 = g0 - slab * ep.ld;
          const long long tile_b = (long long)slab * num_m_tiles + m_tile;
          bulk_store(static_cast<uint16_t*>(ep.out) + blocked_off(tile_b, ep.ld >> 3, within0 >> 3, 0), stg, 32768);
          tma_store_commit();
        }
      } else if constexpr (EPI == EPI_STORE_X || EPI == EPI_STORE_XH) {
        // (acc + bias) * scale [-> GELU] [+ residual] [-> keep mask], fp32 or 16-bit row-major output.  One thread owns one
        // accumulator row, so direct global access would touch 32 lines per warp instruction (the in_proj GEMM wrote its
        // 39 MB at 1.1 TB/s that way): every 32-row x 32-column block goes through a per-warp shared-memory tile instead and
        // moves to / from global memory as full 128-byte (fp32) or 64-byte (16-bit) row segments.
        float* stg = reinterpret_cast<float*>(extra) + (warp - 2) * (32 * Cfg::kXRowWords);   // [32 rows][36] floats (16-bit: [32][20] words)
        const int rows_left = (m_tile < num_m_tiles) ? sh.T - (t0 + q * 32) : 0;      // rows of this warp's slab inside the utterance
        const long long nrow0 = (long long)b * sh.T + t0 + q * 32;                   // global row of slab row 0
        const bool kp = ep.keep == nullptr || (valid && ep.keep[n] != 0);
        uint32_t rr[2][32];
        tmem_ld_x32(trow + h * (BN / 2), rr[0]);
#pragma unroll
        for (int ci = 0; ci < (BN / 2) / 32; ++ci) {
          const int c = h * (BN / 2) + ci * 32;
          tmem_ld_wait();
          if (ci + 1 < (BN / 2) / 32) tmem_ld_x32(trow + c + 32, rr[(ci + 1) & 1]);
          const uint32_t (&r)[32] = rr[ci & 1];
          const int gcol = n_tile * BN + c;            // first output column of this chunk (ld = full row width)
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 bv = __ldg(reinterpret_cast<const float4*>(ep.bias + gcol + j));
            v[j + 0] = (__uint_as_float(r[j + 0]) + bv.x) * ep.scale;
            v[j + 1] = (__uint_as_float(r[j + 1]) + bv.y) * ep.scale;
            v[j + 2] = (__uint_as_float(r[j + 2]) + bv.z) * ep.scale;
            v[j + 3] = (__uint_as_float(r[j + 3]) + bv.w) * ep.scale;
          }
          if (ep.gelu) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = gelu_erf(v[j]);
          }
          if (EPI == EPI_STORE_X && ep.resid != nullptr) {
            // the residual block, loaded as full row segments (4 rows per warp instruction), read back one row per thread
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int rrow = i * 4 + (lane >> 3), c4 = lane & 7;
              float4 rv = make_float4(0.f, 0.f, 0.f, 0.f);
              if (rrow < rows_left) rv = *reinterpret_cast<const float4*>(ep.resid + (nrow0 + rrow) * ep.ld + gcol + 4 * c4);
              *reinterpret_cast<float4*>(stg + rrow * 36 + 4 * c4) = rv;
            }
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 rv = *reinterpret_cast<const float4*>(stg + lane * 36 + j);
              v[j] += rv.x; v[j + 1] += rv.y; v[j + 2] += rv.z; v[j + 3] += rv.w;
            }
            __syncwarp();
          }
          if (!kp) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = 0.f;
          }
          if (EPI == EPI_STORE_X && ep.out_type == OUT_F32) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(stg + lane * 36 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int rrow = i * 4 + (lane >> 3), c4 = lane & 7;
              if (rrow < rows_left)
                *reinterpret_cast<float4*>(static_cast<float*>(ep.out) + (nrow0 + rrow) * ep.ld + gcol + 4 * c4) =
                    *reinterpret_cast<const float4*>(stg + rrow * 36 + 4 * c4);
            }
          } else {
            uint32_t* sw = reinterpret_cast<uint32_t*>(stg);
#pragma unroll
            for (int j = 0; j < 4; ++j)
              *reinterpret_cast<uint4*>(sw + lane * 20 + 4 * j) =
                  ep.out_type == OUT_F16
                      ? make_uint4(Cvt<__half>::pack(v[8 * j], v[8 * j + 1]), Cvt<__half>::pack(v[8 * j + 2], v[8 * j + 3]),
                                   Cvt<__half>::pack(v[8 * j + 4], v[8 * j + 5]), Cvt<__half>::pack(v[8 * j + 6], v[8 * j + 7]))
                      : make_uint4(Cvt<T>::pack(v[8 * j], v[8 * j + 1]), Cvt<T>::pack(v[8 * j + 2], v[8 * j + 3]),
                                   Cvt<T>::pack(v[8 * j + 4], v[8 * j + 5]), Cvt<T>::pack(v[8 * j + 6], v[8 * j + 7]));
            __syncwarp();
            uint16_t* ob = static_cast<uint16_t*>(ep.out);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int rrow = i * 8 + (lane >> 2), c4 = lane & 3;
              if (rrow < rows_left)
                *reinterpret_cast<uint4*>(ob + (nrow0 + rrow) * ep.ld + gcol + 8 * c4) = *reinterpret_cast<const uint4*>(sw + rrow * 20 + 4 * c4);
            }
          }
          __syncwarp();   // the staging tile is reused by the next chunk
        }
      } else if constexpr (EPI == EPI_STORE) {
        const long long nfr = (long long)sh.B * sh.T;
        // two register buffers: the TMEM load of chunk i+1 is in flight while chunk i is converted and stored
        uint32_t rr[2][32];
        tmem_ld_x32(trow + h * (BN / 2), rr[0]);
#pragma unroll
        for (int ci = 0; ci < (BN / 2) / 32; ++ci) {
          const int c = h * (BN / 2) + ci * 32;
          tmem_ld_wait();
          if (ci + 1 < (BN / 2) / 32) tmem_ld_x32(trow + c + 32, rr[(ci + 1) & 1]);
          const uint32_t (&r)[32] = rr[ci & 1];
          if (valid) {
            const int g = n_tile * BN + c;
            const int slab = g / ep.ld, within = g - slab * ep.ld;
            float v[32];
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 bv = __ldg(reinterpret_cast<const float4*>(ep.bias + g + j));
              v[j + 0] = (__uint_as_float(r[j + 0]) + bv.x) * ep.scale;
              v[j + 1] = (__uint_as_float(r[j + 1]) + bv.y) * ep.scale;
              v[j + 2] = (__uint_as_float(r[j + 2]) + bv.z) * ep.scale;
              v[j + 3] = (__uint_as_float(r[j + 3]) + bv.w) * ep.scale;
            }
            if (ep.relu) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
            }
            if (ep.lens != nullptr && t >= __ldg(ep.lens + b)) {   // ragged batch: frames past this utterance's end stay zero
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = 0.f;
            }
            if (ep.out_type == OUT_F32) {
              float4* dst = reinterpret_cast<float4*>(static_cast<float*>(ep.out) + ((long long)slab * nfr + n) * ep.ld + within);
#pragma unroll
              for (int j = 0; j < 8; ++j) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            } else {
              uint4 o[4];
              if (ep.out_type == OUT_F16) {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                  o[j] = make_uint4(Cvt<__half>::pack(v[8 * j], v[8 * j + 1]), Cvt<__half>::pack(v[8 * j + 2], v[8 * j + 3]),
                                    Cvt<__half>::pack(v[8 * j + 4], v[8 * j + 5]), Cvt<__half>::pack(v[8 * j + 6], v[8 * j + 7]));
              } else {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                  o[j] = make_uint4(Cvt<T>::pack(v[8 * j], v[8 * j + 1]), Cvt<T>::pack(v[8 * j + 2], v[8 * j + 3]),
                                    Cvt<T>::pack(v[8 * j + 4], v[8 * j + 5]), Cvt<T>::pack(v[8 * j + 6], v[8 * j + 7]));
              }
              uint16_t* ob = static_cast<uint16_t*>(ep.out);
              if (ep.out_blocked) {
                const int cpr = ep.ld >> 3;
                const long long tile_b = (long long)slab * num_m_tiles + m_tile;
#pragma unroll
                for (int j = 0; j < 4; ++j)
                  *reinterpret_cast<uint4*>(ob + blocked_off(tile_b, cpr, (within >> 3) + j, row)) = o[j];
              } else {
                uint4* dst = reinterpret_cast<uint4*>(ob + ((long long)slab * nfr + n) * ep.ld + within);
#pragma unroll
                for (int j = 0; j < 4; ++j) dst[j] = o[j];
              }
            }
          }
        }
      } else if constexpr (EPI == EPI_GATE) {
        const int C2 = 2 * ep.ld;
        const int entry = ep.sb_entry_base + b * ep.sb_entry_per_b;
        const float* sb_all = ep.sb + (long long)entry * ep.sb_entry_stride + n_tile * 256;
        // step-bias slot of this frame: tap 0 reads x[t - dil], tap 2 reads x[t + dil]; taps on padding contribute nothing
        const int slot = (t < sh.dil ? 1 : 0) + (t + sh.dil >= sh.T ? 2 : 0);
        const float* sb_lane = sb_all + slot * C2;
        const uint32_t* prew = reinterpret_cast<const uint32_t*>(pre);
        uint8_t* ytile = extra + h * 16384;  // [128 frames x 64 ch] staging tile of this column half
        // the TMA store of the previous tile must have finished reading the staging tiles
        if (issuer) tma_store_wait_read<0>();
        named_bar_sync(1, kEpiWarps * 32);
#pragma unroll
        for (int ci = 0; ci < 4; ++ci) {
          const int i = h * 64 + ci * 16;
          uint32_t rg[16], rf[16];
          tmem_ld_x16(trow + i, rg);
          tmem_ld_x16(trow + 128 + i, rf);
          tmem_ld_wait();
          uint32_t packed[8];
          gate_chunk16<T>(rg, rf, prew + ci * 8, prew + 32 + ci * 8, sb_lane + i, packed);
          // rows past T hold junk here; the TMA store clips them
          *sw128_chunk(ytile, row, 2 * ci) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
          *sw128_chunk(ytile, row, 2 * ci + 1) = make_uint4(packed[4], packed[5], packed[6], packed[7]);
        }
      } else if constexpr (EPI == EPI_RES) {
#pragma unroll
        for (int ci = 0; ci < 4; ++ci) {
          const int c = h * 128 + ci * 32;
          uint8_t* xt = extra + (c >> 6) * 16384;  // sub-tile holding channels [c, c + 32)
          const int ch0 = (c & 63) >> 3;           // first 16-byte chunk inside the sub-tile
          uint32_t r[32];
          tmem_ld_x32(trow + c, r);
          tmem_ld_wait();
          const int g = n_tile * BN + c;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4* px = sw128_chunk(xt, row, ch0 + j);
            const uint4 xv = *px;
            const uint32_t xw[4] = {xv.x, xv.y, xv.z, xv.w};
            uint32_t ow[4];
#pragma unroll
            for (int u = 0; u < 4; u += 2) {
              const float4 bv = __ldg(reinterpret_cast<const float4*>(ep.bias + g + 8 * j + 2 * u));
              const float2 xa = Cvt<T>::unpack(xw[u]);
              const float2 xb = Cvt<T>::unpack(xw[u + 1]);
              const int col = 8 * j + 2 * u;
              ow[u] = Cvt<T>::pack((xa.x + __uint_as_float(r[col]) + bv.x) * ep.scale,
                                   (xa.y + __uint_as_float(r[col + 1]) + bv.y) * ep.scale);
              ow[u + 1] = Cvt<T>::pack((xb.x + __uint_as_float(r[col + 2]) + bv.z) * ep.scale,
                                       (xb.y + __uint_as_float(r[col + 3]) + bv.w) * ep.scale);
            }
            *px = make_uint4(ow[0], ow[1], ow[2], ow[3]);  // in place: x_next overwrites x_cur
          }
        }
      } else {  // EPI_OUT
        const int M = ep.M;
        const bool act = valid && (h * 64 < M);               // M is 64 or 128: a warp's 64 mel bins are all in or all out
        const long long col0 = ((long long)b * M + h * 64) * sh.T + t;   // element (b, mel = 64h, t) of a [B,1,M,T] tensor
        const long long Tn = sh.T;
        float xv[2][16];                                        // x_t of the current / next 16-bin chunk
        {
          const float* pf = reinterpret_cast<const float*>(pre);
#pragma unroll
          for (int j = 0; j < 16; ++j) xv[0][j] = pf[j];
        }
#pragma unroll
        for (int ci = 0; ci < 4; ++ci) {
          const int c = h * 64 + ci * 16;
          uint32_t r[16];
          tmem_ld_x16(trow + c, r);
          if (act && ep.x_state != nullptr && ci < 3) {        // fetch the next chunk's x_t while this one is processed
            const float* xs = ep.x_state + col0 + (long long)(ci + 1) * 16 * Tn;
#pragma unroll
            for (int j = 0; j < 16; ++j) xv[(ci + 1) & 1][j] = xs[(long long)j * Tn];
          }
          tmem_ld_wait();
          if (act) {
            const long long cbase = col0 + (long long)ci * 16 * Tn;
            float xn[16];
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
              const float4 bv = __ldg(reinterpret_cast<const float4*>(ep.bias + c + j));
              xn[j] = __uint_as_float(r[j]) + bv.x;
              xn[j + 1] = __uint_as_float(r[j + 1]) + bv.y;
              xn[j + 2] = __uint_as_float(r[j + 2]) + bv.z;
              xn[j + 3] = __uint_as_float(r[j + 3]) + bv.w;
            }
            if (ep.x0_out) {
#pragma unroll
              for (int j = 0; j < 16; ++j) ep.x0_out[cbase + (long long)j * Tn] = xn[j];
            }
            if (ep.x_state) {
#pragma unroll
              for (int j = 0; j < 16; j += 4) {
                float nzv[4] = {0.f, 0.f, 0.f, 0.f};
                if (ep.sigma != 0.f) {
                  if (ep.noise == nullptr) {
                    const float4 nz = philox_normal4(ep.seed, ep.step_id, (uint32_t)n, (uint32_t)((c + j) >> 2));
                    nzv[0] = nz.x; nzv[1] = nz.y; nzv[2] = nz.z; nzv[3] = nz.w;
                  } else {
#pragma unroll
                    for (int u = 0; u < 4; ++u) nzv[u] = __ldg(ep.noise + cbase + (long long)(j + u) * Tn);
                  }
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                  float v = fmaf(ep.c1, xn[j + u], fmaf(ep.sigma, nzv[u], ep.c2 * xv[ci & 1][j + u]));   // same order as pd_tail.cuh
                  if (ep.out_scale != 1.f) v = __fmul_rn(v, ep.out_scale);   // vocoder hand-off, last step only
                  ep.x_state[cbase + (long long)(j + u) * Tn] = v;
                  xn[j + u] = v;
                }
              }
              if (ep.xin_next) {
                uint4* dst = reinterpret_cast<uint4*>(static_cast<uint16_t*>(ep.xin_next) + n * ep.Mp + c);
                dst[0] = make_uint4(Cvt<T>::pack(xn[0], xn[1]), Cvt<T>::pack(xn[2], xn[3]), Cvt<T>::pack(xn[4], xn[5]), Cvt<T>::pack(xn[6], xn[7]));
                dst[1] = make_uint4(Cvt<T>::pack(xn[8], xn[9]), Cvt<T>::pack(xn[10], xn[11]), Cvt<T>::pack(xn[12], xn[13]), Cvt<T>::pack(xn[14], xn[15]));
              }
            }
          }
        }
      }

      // this warp has finished reading accumulator stage `as`
      if (sh.trace && blockIdx.x == 0 && warp == 2 && lane == 0) sh.trace[ti * 16 + 7] = clock64();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (kPair) mbar_arrive_leader(&tempty_bar[as]);
        else mbar_arrive(&tempty_bar[as]);
      }

      if constexpr (EPI == EPI_GATE || EPI == EPI_RES) {
        // publish the shared-memory tile to the async proxy, then one thread hands it to the TMA store engine
        fence_proxy_async_smem();
        named_bar_sync(2, kEpiWarps * 32);
        if (issuer) {
          if constexpr (EPI == EPI_GATE) {
            tma_store_4d(&tmO, extra, n_tile * 128, t0, b, ep.out_slab);
            tma_store_4d(&tmO, extra + 16384, n_tile * 128 + 64, t0, b, ep.out_slab);
            tma_store_commit();
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) tma_store_4d(&tmO, extra + j * 16384, n_tile * BN + j * 64, t0, b, 0);
            tma_store_commit();
            tma_store_wait_read<0>();   // the x tile may be overwritten by the next TMA load after this
            mbar_arrive(xempty_bar);
          }
        }
      }
    }
    if constexpr (EPI == EPI_GATE || EPI == EPI_RES) {
      if (issuer) tma_store_wait_all();  // all bulk stores complete before the CTA retires
    }
    if constexpr (EPI == EPI_STORE_BLK) {
      if (q == 0 && lane == 0) tma_store_wait_all();
    }
  }

  tc_fence_before();
  if constexpr (kPair) cluster_sync_all();  // the leader's MMAs read the peer's shared memory and write its TMEM
  else __syncthreads();
  if (sh.trace && blockIdx.x == 0 && threadIdx.x == 0) sh.trace[14] = clock64();  // all roles done
  if (warp == 2) {
    tc_fence_after();
    if constexpr (kPair) tmem_dealloc_2sm<kTmemCols>(tmem_base);
    else tmem_dealloc<kTmemCols>(tmem_base);
  }
}

}  // namespace pd
