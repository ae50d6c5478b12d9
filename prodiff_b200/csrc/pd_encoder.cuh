// Non-GEMM kernels of the FastSpeech text encoder (token-level half):
//   FastspeechEncoder.forward      /root/reference/modules/fastspeech/tts_modules.py:305-330  (embedding + positions)
//   FFTBlocks.forward              tts_modules.py:262-288                                      (masks, last LayerNorm)
//   EncSALayer.forward             /root/reference/modules/commons/common_layers.py:649-673    (pre-LN attention + conv FFN)
//   MultiheadAttention.forward     common_layers.py:242-287 -> F.multi_head_attention_forward  (bias-free, key padding mask)
// The dense contractions (packed q/k/v projection, out_proj, the kernel-9 FFN convolution, ffn_2) run on the generic
// tcgen05 GEMM kernel (pd_gemm.cuh, EPI_STORE_X epilogue: GELU, residual, keep mask).  What is left is per-token work on
// ~10^4 tokens: the embedding and LayerNorm kernels (fp32, one warp per token) and the attention (warp-level fp16 mma.sync
// with an fp32 online softmax, k_enc_attention_f16).
#pragma once
#include "pd_aux.cuh"

namespace pd {

constexpr int kEncH = 256;        // hidden size the encoder kernels are specialised for (base_config.yaml:112)
constexpr int kAttnK = 64;        // keys per shared-memory tile

// x[b,t,:] = (sqrt(H) * E[tok] + extra + pos(t)) * keep  (tts_modules.py:314-325, :269-275) and keep[b,t] = tok != 0.
// pos(t): row p of the sinusoid table (common_layers.py:111-128), p = number of non-padding tokens up to and including t
// (make_positions over ~padding_mask, tts_utils.py:6-18), row 0 (padding) = zeros; sin in [0, H/2), cos in [H/2, H).
// One warp per token; the warp counts the non-padding tokens before its own (a few hundred at most) with a strided count and a shuffle reduction.
__global__ void __launch_bounds__(256) k_enc_embed(float* __restrict__ x, uint8_t* __restrict__ keep,
                                                   const long long* __restrict__ tok, const float* __restrict__ emb_w,
                                                   const float* __restrict__ extra, int Tn, int vocab) {
  const int b = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t = blockIdx.x * 8 + warp;
  griddep_launch();   // (first kernel of the call: launched normally, but the LayerNorm behind it may start early)
  if (t >= Tn) return;
  const long long* tk = tok + (long long)b * Tn;
  const long long id = tk[t];
  const long long n = (long long)b * Tn + t;
  const bool kp = id != 0;
  if (lane == 0) keep[n] = kp ? 1 : 0;
  float* xr = x + n * kEncH;
  if (!kp) {                                               // padding: masked to zero whatever the embedding row holds
#pragma unroll
    for (int j = 0; j < kEncH / 32; ++j) xr[lane + 32 * j] = 0.f;
    return;
  }
  int cnt = 0;
  for (int i = lane; i <= t; i += 32) cnt += tk[i] != 0 ? 1 : 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  const float embed_scale = sqrtf((float)kEncH);
  const float step = logf(10000.f) / (float)(kEncH / 2 - 1);
  const float* er = emb_w + (id < vocab ? id : 0) * kEncH;
  const float pos = (float)cnt;
#pragma unroll
  for (int j = 0; j < kEncH / 32; ++j) {
    const int c = lane + 32 * j;
    const int f = c < kEncH / 2 ? c : c - kEncH / 2;
    const float ang = pos * expf((float)f * -step);
    float v = embed_scale * er[c];
    if (extra != nullptr) v += extra[n * kEncH + c];
    v += c < kEncH / 2 ? sinf(ang) : cosf(ang);
    xr[c] = v;
  }
}

// LayerNorm over H = 256 (eps 1e-5, biased variance, torch.nn.LayerNorm), one warp per token, fp32 statistics.
// out16 != nullptr: the 16-bit A operand of the next GEMM (written for EVERY token: the reference's conv and attention
// see LayerNorm(0) = bias at padding positions, not zeros).  out32 != nullptr: fp32 output times the keep mask (the
// encoder's last LayerNorm, tts_modules.py:281-282).
template <typename T>
__global__ void __launch_bounds__(256) k_enc_layernorm(T* __restrict__ out16, float* __restrict__ out32, const float* __restrict__ x,
                                                       const float* __restrict__ w, const float* __restrict__ bvec,
                                                       const uint8_t* __restrict__ keep, long long n_tok) {
  const int lane = threadIdx.x & 31;
  const long long n = blockIdx.x * 8ll + (threadIdx.x >> 5);
  griddep_launch();
  griddep_wait();
  if (n >= n_tok) return;
  const float4* xr = reinterpret_cast<const float4*>(x + n * kEncH);
  float4 v[2] = {xr[lane], xr[lane + 32]};   // channels 4*lane..+3 and 128 + 4*lane..+3
  float s = v[0].x + v[0].y + v[0].z + v[0].w + v[1].x + v[1].y + v[1].z + v[1].w;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s * (1.f / kEncH);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
    q += v[i].x * v[i].x + v[i].y * v[i].y + v[i].z * v[i].z + v[i].w * v[i].w;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float rstd = rsqrtf(q * (1.f / kEncH) + 1e-5f);
  const float k = (out32 != nullptr && keep != nullptr && !keep[n]) ? 0.f : 1.f;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int c = 4 * lane + 128 * i;
    const float4 wv = __ldg(reinterpret_cast<const float4*>(w + c)), bv = __ldg(reinterpret_cast<const float4*>(bvec + c));
    const float4 y = make_float4(v[i].x * rstd * wv.x + bv.x, v[i].y * rstd * wv.y + bv.y, v[i].z * rstd * wv.z + bv.z,
                                 v[i].w * rstd * wv.w + bv.w);
    if (out16 != nullptr) {
      uint2 o;
      o.x = Cvt<T>::pack(y.x, y.y);
      o.y = Cvt<T>::pack(y.z, y.w);
      *reinterpret_cast<uint2*>(out16 + n * kEncH + c) = o;
    }
    if (out32 != nullptr) *reinterpret_cast<float4*>(out32 + n * kEncH + c) = make_float4(y.x * k, y.y * k, y.z * k, y.w * k);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Self-attention of one (utterance, head, 64-query block): softmax(q k^T / sqrt(HD) + key padding mask) v with an online
// softmax over 64-key tiles.  T_txt x T_txt x head_dim per (utterance, head) is far too small to feed a tcgen05 pipeline (a
// 200 x 200 score matrix is 1.6 M MACs), so this uses warp-level mma.sync m16n8k16 with fp16 operands, fp32 accumulation
// and fp32 softmax.  fp16 has a 10-bit mantissa like TF32, and q, k, v (projections of a LayerNorm output) and the
// probabilities are far inside its range, so it is used whatever the encoder's GEMM operand type is.  History, all at
// 64 x 200 tokens: fp32 on the CUDA cores 128 us per launch (issue-bound: one FMA per MAC); TF32 m16n8k8 with P staged
// through shared memory 64-66 us (the legacy path's TF32 rate); this fp16 form with fp32 q|k|v converted while filling the
// tiles 46 us; this version (the in-projection GEMM stores fp16, tiles arrive by cp.async, V through ldmatrix.trans) 24 us.
// The probabilities never touch shared memory: the accumulator fragments of two adjacent 8-key blocks ARE the A fragment
// of one 16-key k-step.  CTA = 4 warps x 16 queries, three CTAs per SM; q, K and V tiles row-major [token][HD] in fp16 with
// rows padded by 8 halves (every fragment load and ldmatrix bank-conflict free).  A row whose keys are ALL padding yields
// zeros (the reference yields NaN there).
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  __half2 v = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ void mma_f16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

constexpr int kHQ = 64;   // queries per CTA (4 warps x 16)
template <int HD> struct AttnCfg {
  static constexpr int RW = (HD + 8) / 2;                  // words per staged row (q, K and V rows alike)
  static constexpr int kStageWords = 2 * kAttnK * RW;      // one K tile + one V tile
  static constexpr int kSmemBytes = 2 * kStageWords * 4;   // two stages; the q tile is staged through stage 1
};
__device__ __forceinline__ void cp_async16(void* dst, const void* src, bool pred) {
  const int n = pred ? 16 : 0;   // src-size 0: the 16 bytes are zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
// four 8x8 fp16 matrices, transposed: with lane l pointing at row (l & 7) + 8 * ((l >> 3) & 1) of a row-major [key][dim]
// tile (dims n0 .. n0 + 7 for l < 16, n0 + 8 .. n0 + 15 above), r0/r1 are the B fragment (keys 0-7 / 8-15) of output
// block n0 and r2/r3 that of block n0 + 8.
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(p)));
}

// qkv16 [B*T, 3H] fp16: the packed in-projection output as its GEMM epilogue stores it (q | k | v).  q, K and V tiles are
// plain cp.async copies (two-stage ring over the 64-key tiles: the fill of tile i + 1 overlaps the MMAs of tile i); the
// 1 / sqrt(head_dim) factor is applied to the fp32 scores.
template <typename T, int HD>
__global__ void __launch_bounds__(128, 3) k_enc_attention_f16(T* __restrict__ out16, const __half* __restrict__ qkv16,
                                                              const uint8_t* __restrict__ keep, int Tn, float scale) {
  using Cfg = AttnCfg<HD>;
  constexpr int RW = Cfg::RW;
  constexpr int KQ = HD / 16;                   // k-steps of q k^T
  constexpr int NO = HD / 8;                    // 8-wide output blocks of P v
  extern __shared__ uint32_t smw[];
  __shared__ uint8_t kvalid[2][kAttnK];
  __shared__ int s_last;
  const int b = blockIdx.z, head = blockIdx.y, q0 = blockIdx.x * kHQ;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const long long row0 = (long long)b * Tn;
  const int ld = 3 * kEncH;
  griddep_launch();
  griddep_wait();
  if (!__syncthreads_or(tid < kHQ && q0 + tid < Tn && keep[row0 + q0 + tid] != 0)) {
    // padding-only query block: no attention to compute, but out_proj's A operand must not be uninitialised memory (its
    // keep mask is a select today; zeros keep it safe if that ever becomes a multiply)
    for (int i = tid; i < kHQ * (HD / 8); i += 128) {
      const int r = i / (HD / 8), c8 = i - r * (HD / 8);
      if (q0 + r < Tn) *reinterpret_cast<uint4*>(out16 + (row0 + q0 + r) * kEncH + head * HD + 8 * c8) = make_uint4(0u, 0u, 0u, 0u);
    }
    return;
  }
  if (tid == 0) s_last = -1;
  __syncthreads();
  {
    int last = -1;
    for (int i = tid; i < Tn; i += 128)
      if (keep[row0 + i]) last = i;
    if (last >= 0) atomicMax(&s_last, last);
  }
  auto load_tile = [&](int tile, int stage) {     // K and V rows of 64 keys -> stage, 16 bytes (8 dims) per copy
    uint32_t* Kd = smw + stage * Cfg::kStageWords;
    uint32_t* Vd = Kd + kAttnK * RW;
    const int k0 = tile * kAttnK;
    for (int i = tid; i < kAttnK * (HD / 8); i += 128) {
      const int r = i / (HD / 8), c8 = i - r * (HD / 8);
      const bool in = k0 + r < Tn;
      const __half* src = qkv16 + (row0 + (in ? k0 + r : 0)) * ld + head * HD + 8 * c8;
      cp_async16(Kd + r * RW + 4 * c8, src + kEncH, in);
      cp_async16(Vd + r * RW + 4 * c8, src + 2 * kEncH, in);
    }
    if (tid < kAttnK) kvalid[stage][tid] = (k0 + tid < Tn && keep[row0 + k0 + tid]) ? 1 : 0;
  };
  load_tile(0, 0);
  cp_async_commit();
  uint32_t* Qs = smw + Cfg::kStageWords;            // q tile [64][HD + 8] through stage 1
  for (int i = tid; i < kHQ * (HD / 8); i += 128) {
    const int r = i / (HD / 8), c8 = i - r * (HD / 8);
    const bool in = q0 + r < Tn;
    cp_async16(Qs + r * RW + 4 * c8, qkv16 + (row0 + (in ? q0 + r : 0)) * ld + head * HD + 8 * c8, in);
  }
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();
  const int n_tiles = (s_last + kAttnK) / kAttnK;   // keys past the last real token contribute nothing
  uint32_t qa[KQ][4];
  {
    const uint32_t* qr = Qs + (warp * 16 + g) * RW + t;
#pragma unroll
    for (int ks = 0; ks < KQ; ++ks) {
      qa[ks][0] = qr[ks * 8];
      qa[ks][1] = qr[8 * RW + ks * 8];
      qa[ks][2] = qr[ks * 8 + 4];
      qa[ks][3] = qr[8 * RW + ks * 8 + 4];
    }
  }
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;   // rows g and g + 8 of this warp's 16 queries
  float o[NO][4];
#pragma unroll
  for (int nb = 0; nb < NO; ++nb) o[nb][0] = o[nb][1] = o[nb][2] = o[nb][3] = 0.f;
  for (int tile = 0; tile < n_tiles; ++tile) {
    const int stage = tile & 1;
    // everyone is done with the other stage (tile - 1's fragments, or the q tile): prefetch tile + 1 into it, then wait for
    // this tile's copies only
    __syncthreads();
    if (tile + 1 < n_tiles) load_tile(tile + 1, stage ^ 1);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    const uint32_t* Ks = smw + stage * Cfg::kStageWords;
    const uint32_t* Vs = Ks + kAttnK * RW;
    // S[16 x 64]: sc[nb] holds keys nb*8 + 2t, +1 for rows g (.x .y) and g + 8 (.z .w)
    float sc[kAttnK / 8][4];
#pragma unroll
    for (int nb = 0; nb < kAttnK / 8; ++nb) {
      sc[nb][0] = sc[nb][1] = sc[nb][2] = sc[nb][3] = 0.f;
      const uint32_t* kr = Ks + (nb * 8 + g) * RW + t;
#pragma unroll
      for (int ks = 0; ks < KQ; ++ks) mma_f16(sc[nb], qa[ks], kr[ks * 8], kr[ks * 8 + 4]);
    }
    float t0 = -INFINITY, t1 = -INFINITY;
#pragma unroll
    for (int nb = 0; nb < kAttnK / 8; ++nb) {
      const bool va = kvalid[stage][nb * 8 + 2 * t] != 0, vb = kvalid[stage][nb * 8 + 2 * t + 1] != 0;
      sc[nb][0] = va ? sc[nb][0] * scale : -INFINITY;
      sc[nb][2] = va ? sc[nb][2] * scale : -INFINITY;
      sc[nb][1] = vb ? sc[nb][1] * scale : -INFINITY;
      sc[nb][3] = vb ? sc[nb][3] * scale : -INFINITY;
      t0 = fmaxf(t0, fmaxf(sc[nb][0], sc[nb][1]));
      t1 = fmaxf(t1, fmaxf(sc[nb][2], sc[nb][3]));
    }
    t0 = fmaxf(t0, __shfl_xor_sync(0xffffffffu, t0, 1));
    t0 = fmaxf(t0, __shfl_xor_sync(0xffffffffu, t0, 2));
    t1 = fmaxf(t1, __shfl_xor_sync(0xffffffffu, t1, 1));
    t1 = fmaxf(t1, __shfl_xor_sync(0xffffffffu, t1, 2));
    const float n0 = fmaxf(m0, t0), n1 = fmaxf(m1, t1);
    const float c0 = m0 == -INFINITY ? 0.f : expf(m0 - n0), c1 = m1 == -INFINITY ? 0.f : expf(m1 - n1);
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int nb = 0; nb < kAttnK / 8; ++nb) {
      // exp(-inf - n) = 0 for masked keys; a row with no valid key so far has n = -inf: keep its probabilities at 0
      sc[nb][0] = n0 == -INFINITY ? 0.f : expf(sc[nb][0] - n0);
      sc[nb][1] = n0 == -INFINITY ? 0.f : expf(sc[nb][1] - n0);
      sc[nb][2] = n1 == -INFINITY ? 0.f : expf(sc[nb][2] - n1);
      sc[nb][3] = n1 == -INFINITY ? 0.f : expf(sc[nb][3] - n1);
      s0 += sc[nb][0] + sc[nb][1];
      s1 += sc[nb][2] + sc[nb][3];
    }
    s0 += __shfl_xor_sync(0xffffffffu, s0, 1);
    s0 += __shfl_xor_sync(0xffffffffu, s0, 2);
    s1 += __shfl_xor_sync(0xffffffffu, s1, 1);
    s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
    l0 = l0 * c0 + s0;
    l1 = l1 * c1 + s1;
    m0 = n0;
    m1 = n1;
#pragma unroll
    for (int nb = 0; nb < NO; ++nb) {
      o[nb][0] *= c0; o[nb][1] *= c0; o[nb][2] *= c1; o[nb][3] *= c1;
    }
    // O[16 x HD] += P[16 x 64] V[64 x HD]: the accumulator fragments of key blocks 2ks, 2ks + 1 are the A fragment of k-step
    // ks; the B fragments come from the row-major V tile through ldmatrix.trans (two 8-dim output blocks per instruction)
#pragma unroll
    for (int ks = 0; ks < kAttnK / 16; ++ks) {
      uint32_t pa[4];
      pa[0] = pack_h2(sc[2 * ks][0], sc[2 * ks][1]);
      pa[1] = pack_h2(sc[2 * ks][2], sc[2 * ks][3]);
      pa[2] = pack_h2(sc[2 * ks + 1][0], sc[2 * ks + 1][1]);
      pa[3] = pack_h2(sc[2 * ks + 1][2], sc[2 * ks + 1][3]);
      const uint32_t* vrow = Vs + (ks * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * RW + (lane >> 4) * 4;
#pragma unroll
      for (int nb = 0; nb < NO; nb += 2) {
        uint32_t vb[4];
        ldmatrix_x4_trans(vb, vrow + nb * 4);
        mma_f16(o[nb], pa, vb[0], vb[1]);
        mma_f16(o[nb + 1], pa, vb[2], vb[3]);
      }
    }
  }
  const float i0 = l0 > 0.f ? 1.f / l0 : 0.f, i1 = l1 > 0.f ? 1.f / l1 : 0.f;
  const int qa0 = q0 + warp * 16 + g, qa1 = qa0 + 8;
#pragma unroll
  for (int nb = 0; nb < NO; ++nb) {
    const int col = head * HD + nb * 8 + 2 * t;
    if (qa0 < Tn) *reinterpret_cast<uint32_t*>(out16 + (row0 + qa0) * kEncH + col) = Cvt<T>::pack(o[nb][0] * i0, o[nb][1] * i0);
    if (qa1 < Tn) *reinterpret_cast<uint32_t*>(out16 + (row0 + qa1) * kEncH + col) = Cvt<T>::pack(o[nb][2] * i1, o[nb][3] * i1);
  }
}

// ffn_1.weight [4H, H, k] fp32 -> [4H rows][K = tap * H + c] 16-bit (the implicit-GEMM K order of pd_gemm.cuh's tap mode).
template <typename T>
__global__ void k_pack_conv_taps(T* __restrict__ dst, const float* __restrict__ src, int rows, int C, int taps) {
  const long long total = (long long)rows * taps * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / (taps * C));
    const int rem = (int)(i - (long long)r * taps * C);
    const int tap = rem / C, c = rem - tap * C;
    dst[i] = from_float<T>(src[((long long)r * C + c) * taps + tap]);
  }
}

}  // namespace pd
