// Non-GEMM kernels of the FastSpeech text encoder (token-level half):
//   FastspeechEncoder.forward      /root/reference/modules/fastspeech/tts_modules.py:305-330  (embedding + positions)
//   FFTBlocks.forward              tts_modules.py:262-288                                      (masks, last LayerNorm)
//   EncSALayer.forward             /root/reference/modules/commons/common_layers.py:649-673    (pre-LN attention + conv FFN)
//   MultiheadAttention.forward     common_layers.py:242-287 -> F.multi_head_attention_forward  (bias-free, key padding mask)
// The dense contractions (packed q/k/v projection, out_proj, the kernel-9 FFN convolution, ffn_2) run on the generic
// tcgen05 GEMM kernel (pd_gemm.cuh, EPI_STORE_X epilogue: GELU, residual, keep mask).  What is left is per-token work on
// ~10^4 tokens: fp32 throughout, one warp per token, and a tiled fp32 attention with an online softmax (T_txt x T_txt x
// head_dim per head is a few GFLOP per batch -- small against the 9-tap FFN GEMM, so it stays on the CUDA cores for now).
#pragma once
#include "pd_aux.cuh"

namespace pd {

constexpr int kEncH = 256;        // hidden size the encoder kernels are specialised for (base_config.yaml:112)
constexpr int kAttnQ = 32;        // queries per CTA (8 warps x 4)
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

// Self-attention of one (utterance, head, 32-query block): softmax(q k^T / sqrt(HD) + key padding mask) v, fp32, with an
// online softmax over 64-key tiles.  qkv [B*T, 3H] fp32 = the packed in-projection output (q | k | v, head h at columns
// h*HD of each part); out16 [B*T, H] 16-bit = the A operand of out_proj.  Each warp owns 4 queries: a lane scores keys
// (lane, lane + 32) of the tile for all 4, then accumulates output dims [4*lane', ...) over the tile's probabilities
// read back from shared memory.  A row whose keys are ALL padding yields zeros (the reference yields NaN there).
template <typename T, int HD>
__global__ void __launch_bounds__(256) k_enc_attention(T* __restrict__ out16, const float* __restrict__ qkv,
                                                       const uint8_t* __restrict__ keep, int Tn, float scale) {
  constexpr int KS = HD + 4;                    // padded key row: conflict-free 128-bit reads with one key per lane
  constexpr int DL = HD / 32;                   // output dims per lane
  extern __shared__ float sm[];
  float* Ks = sm;                               // [64][HD + 4]
  float* Vs = Ks + kAttnK * KS;                 // [64][HD]
  float* Qs = Vs + kAttnK * HD;                 // [32][HD], pre-scaled
  float* Ps = Qs + kAttnQ * HD;                 // [8 warps][4 queries][64]
  __shared__ uint8_t kvalid[kAttnK];
  const int b = blockIdx.z, head = blockIdx.y, q0 = blockIdx.x * kAttnQ;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long row0 = (long long)b * Tn;
  const int ld = 3 * kEncH;
  // a query block of padding tokens only: its rows are zeroed by the keep mask after out_proj, nothing to compute
  if (!__syncthreads_or(tid < kAttnQ && q0 + tid < Tn && keep[row0 + q0 + tid] != 0)) return;
  for (int i = tid; i < kAttnQ * (HD / 4); i += 256) {
    const int r = i / (HD / 4), c4 = i - r * (HD / 4);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (q0 + r < Tn) v = *reinterpret_cast<const float4*>(qkv + (row0 + q0 + r) * ld + head * HD + 4 * c4);
    *reinterpret_cast<float4*>(Qs + r * HD + 4 * c4) = make_float4(v.x * scale, v.y * scale, v.z * scale, v.w * scale);
  }
  float m[4], l[4], o[4][DL];
#pragma unroll
  for (int qi = 0; qi < 4; ++qi) {
    m[qi] = -INFINITY;
    l[qi] = 0.f;
#pragma unroll
    for (int d = 0; d < DL; ++d) o[qi][d] = 0.f;
  }
  float* Pw = Ps + warp * 4 * kAttnK;
  for (int k0 = 0; k0 < Tn; k0 += kAttnK) {
    // previous tile fully consumed (and Qs written, first time round); a tile of padded keys only contributes nothing
    if (!__syncthreads_or(tid < kAttnK && k0 + tid < Tn && keep[row0 + k0 + tid] != 0)) continue;
    for (int i = tid; i < kAttnK * (HD / 4); i += 256) {
      const int r = i / (HD / 4), c4 = i - r * (HD / 4);
      float4 kv = make_float4(0.f, 0.f, 0.f, 0.f), vv = kv;
      if (k0 + r < Tn) {
        const float* src = qkv + (row0 + k0 + r) * ld + head * HD + 4 * c4;
        kv = *reinterpret_cast<const float4*>(src + kEncH);
        vv = *reinterpret_cast<const float4*>(src + 2 * kEncH);
      }
      *reinterpret_cast<float4*>(Ks + r * KS + 4 * c4) = kv;
      *reinterpret_cast<float4*>(Vs + r * HD + 4 * c4) = vv;
    }
    if (tid < kAttnK) kvalid[tid] = (k0 + tid < Tn && keep[row0 + k0 + tid]) ? 1 : 0;
    __syncthreads();
    // scores of keys (lane, lane + 32) for this warp's 4 queries
    float s[4][2];
#pragma unroll
    for (int qi = 0; qi < 4; ++qi) s[qi][0] = s[qi][1] = 0.f;
    const float* ka = Ks + lane * KS;
    const float* kb = Ks + (lane + 32) * KS;
    const float* qa = Qs + (warp * 4) * HD;
#pragma unroll 4
    for (int d = 0; d < HD; d += 4) {
      const float4 a = *reinterpret_cast<const float4*>(ka + d), c = *reinterpret_cast<const float4*>(kb + d);
#pragma unroll
      for (int qi = 0; qi < 4; ++qi) {
        const float4 qv = *reinterpret_cast<const float4*>(qa + qi * HD + d);
        s[qi][0] += qv.x * a.x + qv.y * a.y + qv.z * a.z + qv.w * a.w;
        s[qi][1] += qv.x * c.x + qv.y * c.y + qv.z * c.z + qv.w * c.w;
      }
    }
    const bool v0 = kvalid[lane] != 0, v1 = kvalid[lane + 32] != 0;
#pragma unroll
    for (int qi = 0; qi < 4; ++qi) {
      float tm = fmaxf(v0 ? s[qi][0] : -INFINITY, v1 ? s[qi][1] : -INFINITY);
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) tm = fmaxf(tm, __shfl_xor_sync(0xffffffffu, tm, off));
      const float mn = fmaxf(m[qi], tm);
      const float corr = (m[qi] == -INFINITY) ? 0.f : expf(m[qi] - mn);
      const float p0 = (v0 && mn != -INFINITY) ? expf(s[qi][0] - mn) : 0.f;
      const float p1 = (v1 && mn != -INFINITY) ? expf(s[qi][1] - mn) : 0.f;
      float ps = p0 + p1;
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) ps += __shfl_xor_sync(0xffffffffu, ps, off);
      l[qi] = l[qi] * corr + ps;
      m[qi] = mn;
#pragma unroll
      for (int d = 0; d < DL; ++d) o[qi][d] *= corr;
      Pw[qi * kAttnK + lane] = p0;
      Pw[qi * kAttnK + lane + 32] = p1;
    }
    __syncwarp();
    // o[qi][dims of this lane] += sum over the tile's keys of p * v
    for (int k = 0; k < kAttnK; k += 4) {
      float4 pq[4];
#pragma unroll
      for (int qi = 0; qi < 4; ++qi) pq[qi] = *reinterpret_cast<const float4*>(Pw + qi * kAttnK + k);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        float vv[DL];
        if constexpr (DL == 4) {
          const float4 t4 = *reinterpret_cast<const float4*>(Vs + (k + kk) * HD + 4 * lane);
          vv[0] = t4.x; vv[1] = t4.y; vv[2] = t4.z; vv[3] = t4.w;
        } else {
#pragma unroll
          for (int d = 0; d < DL; ++d) vv[d] = Vs[(k + kk) * HD + DL * lane + d];
        }
#pragma unroll
        for (int qi = 0; qi < 4; ++qi) {
          const float p = kk == 0 ? pq[qi].x : (kk == 1 ? pq[qi].y : (kk == 2 ? pq[qi].z : pq[qi].w));
#pragma unroll
          for (int d = 0; d < DL; ++d) o[qi][d] += p * vv[d];
        }
      }
    }
    __syncwarp();
  }
#pragma unroll
  for (int qi = 0; qi < 4; ++qi) {
    const int q = q0 + warp * 4 + qi;
    if (q >= Tn) continue;
    const float inv = l[qi] > 0.f ? 1.f / l[qi] : 0.f;
    T* dst = out16 + (row0 + q) * kEncH + head * HD + DL * lane;
#pragma unroll
    for (int d = 0; d < DL; d += 2) *reinterpret_cast<uint32_t*>(dst + d) = Cvt<T>::pack(o[qi][d] * inv, o[qi][d + 1] * inv);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// The same attention on the tensor cores.  T_txt x T_txt x head_dim per (utterance, head) is far too small to feed a
// tcgen05 pipeline (a 200 x 200 score matrix is 1.6 M MACs), so this uses warp-level mma.sync m16n8k8 with TF32 operands
// (fp32 storage, 10-bit mantissa -- no 16-bit rounding of q, k, v or the probabilities) and fp32 accumulation: the CUDA-core
// version above is bound by instruction issue (one FMA per MAC), this one issues one instruction per 2048 MACs.
// CTA = 4 warps x 16 queries; per 64-key tile each warp computes S = q k^T (fragments of q live in registers for the whole
// kernel), the online softmax on the accumulator fragments (a row is spread over the 4 lanes of a quad), P through a
// per-warp shared-memory tile (accumulator layout -> A-operand layout), then O += P v.  Shared-memory row strides are
// padded (+4 / +8 words) so that every fragment load is bank-conflict free.
__device__ __forceinline__ uint32_t f2tf32(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return r;
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

constexpr int kMmaQ = 64;   // queries per CTA (4 warps x 16)
template <int HD> struct AttnMmaCfg {
  static constexpr int QS = HD + 4, KS = HD + 4, VS = HD + 8, PS = kAttnK + 4;
  static constexpr int kQWords = kMmaQ * QS, kPWords = 4 * 16 * PS;
  static constexpr int kSmemBytes = (kAttnK * KS + kAttnK * VS + (kQWords > kPWords ? kQWords : kPWords)) * 4;
};

template <typename T, int HD>
__global__ void __launch_bounds__(128, 2) k_enc_attention_mma(T* __restrict__ out16, const float* __restrict__ qkv,
                                                              const uint8_t* __restrict__ keep, int Tn, float scale) {
  using Cfg = AttnMmaCfg<HD>;
  constexpr int QS = Cfg::QS, KS = Cfg::KS, VS = Cfg::VS, PS = Cfg::PS;
  constexpr int KQ = HD / 8;                    // k-steps of q k^T = 8-wide output blocks of P v
  extern __shared__ float sm[];
  uint32_t* Ks = reinterpret_cast<uint32_t*>(sm);           // [64][HD + 4] tf32
  uint32_t* Vs = Ks + kAttnK * KS;                           // [64][HD + 8] tf32
  float* QPs = reinterpret_cast<float*>(Vs + kAttnK * VS);   // q tile [64][HD + 4] at start, then the per-warp P tiles
  __shared__ uint8_t kvalid[kAttnK];
  const int b = blockIdx.z, head = blockIdx.y, q0 = blockIdx.x * kMmaQ;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const long long row0 = (long long)b * Tn;
  const int ld = 3 * kEncH;
  if (!__syncthreads_or(tid < kMmaQ && q0 + tid < Tn && keep[row0 + q0 + tid] != 0)) return;   // padding-only query block
  for (int i = tid; i < kMmaQ * (HD / 4); i += 128) {
    const int r = i / (HD / 4), c4 = i - r * (HD / 4);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (q0 + r < Tn) v = *reinterpret_cast<const float4*>(qkv + (row0 + q0 + r) * ld + head * HD + 4 * c4);
    *reinterpret_cast<float4*>(QPs + r * QS + 4 * c4) = make_float4(v.x * scale, v.y * scale, v.z * scale, v.w * scale);
  }
  __syncthreads();
  uint32_t qa[KQ][4];
  {
    const float* qr = QPs + (warp * 16 + g) * QS + t;
#pragma unroll
    for (int ks = 0; ks < KQ; ++ks) {
      qa[ks][0] = f2tf32(qr[ks * 8]);
      qa[ks][1] = f2tf32(qr[8 * QS + ks * 8]);
      qa[ks][2] = f2tf32(qr[ks * 8 + 4]);
      qa[ks][3] = f2tf32(qr[8 * QS + ks * 8 + 4]);
    }
  }
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;   // rows g and g + 8 of this warp's 16 queries
  float o[KQ][4];
#pragma unroll
  for (int nb = 0; nb < KQ; ++nb) o[nb][0] = o[nb][1] = o[nb][2] = o[nb][3] = 0.f;
  float* Pw = QPs + warp * 16 * PS;
  for (int k0 = 0; k0 < Tn; k0 += kAttnK) {
    // (also orders the q-fragment reads / the previous tile's reads before the shared memory is overwritten)
    if (!__syncthreads_or(tid < kAttnK && k0 + tid < Tn && keep[row0 + k0 + tid] != 0)) continue;
    for (int i = tid; i < kAttnK * (HD / 4); i += 128) {
      const int r = i / (HD / 4), c4 = i - r * (HD / 4);
      float4 kv = make_float4(0.f, 0.f, 0.f, 0.f), vv = kv;
      if (k0 + r < Tn) {
        const float* src = qkv + (row0 + k0 + r) * ld + head * HD + 4 * c4;
        kv = *reinterpret_cast<const float4*>(src + kEncH);
        vv = *reinterpret_cast<const float4*>(src + 2 * kEncH);
      }
      *reinterpret_cast<uint4*>(Ks + r * KS + 4 * c4) = make_uint4(f2tf32(kv.x), f2tf32(kv.y), f2tf32(kv.z), f2tf32(kv.w));
      *reinterpret_cast<uint4*>(Vs + r * VS + 4 * c4) = make_uint4(f2tf32(vv.x), f2tf32(vv.y), f2tf32(vv.z), f2tf32(vv.w));
    }
    if (tid < kAttnK) kvalid[tid] = (k0 + tid < Tn && keep[row0 + k0 + tid]) ? 1 : 0;
    __syncthreads();
    // S[16 x 64]: accumulator fragment sc[nb] holds keys nb*8 + 2t, +1 for rows g (.x .y) and g + 8 (.z .w)
    float sc[kAttnK / 8][4];
#pragma unroll
    for (int nb = 0; nb < kAttnK / 8; ++nb) {
      sc[nb][0] = sc[nb][1] = sc[nb][2] = sc[nb][3] = 0.f;
      const uint32_t* kr = Ks + (nb * 8 + g) * KS + t;
#pragma unroll
      for (int ks = 0; ks < KQ; ++ks) mma_tf32(sc[nb], qa[ks], kr[ks * 8], kr[ks * 8 + 4]);
    }
    float t0 = -INFINITY, t1 = -INFINITY;
#pragma unroll
    for (int nb = 0; nb < kAttnK / 8; ++nb) {
      const bool va = kvalid[nb * 8 + 2 * t] != 0, vb = kvalid[nb * 8 + 2 * t + 1] != 0;
      if (!va) sc[nb][0] = sc[nb][2] = -INFINITY;
      if (!vb) sc[nb][1] = sc[nb][3] = -INFINITY;
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
    __syncwarp();                                   // the previous tile's P reads are done
#pragma unroll
    for (int nb = 0; nb < kAttnK / 8; ++nb) {
      // exp(-inf - n) = 0 for masked keys; a row with no valid key so far has n = -inf: keep its probabilities at 0
      const float p00 = n0 == -INFINITY ? 0.f : expf(sc[nb][0] - n0), p01 = n0 == -INFINITY ? 0.f : expf(sc[nb][1] - n0);
      const float p10 = n1 == -INFINITY ? 0.f : expf(sc[nb][2] - n1), p11 = n1 == -INFINITY ? 0.f : expf(sc[nb][3] - n1);
      s0 += p00 + p01;
      s1 += p10 + p11;
      *reinterpret_cast<float2*>(Pw + g * PS + nb * 8 + 2 * t) = make_float2(p00, p01);
      *reinterpret_cast<float2*>(Pw + (g + 8) * PS + nb * 8 + 2 * t) = make_float2(p10, p11);
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
    for (int nb = 0; nb < KQ; ++nb) {
      o[nb][0] *= c0; o[nb][1] *= c0; o[nb][2] *= c1; o[nb][3] *= c1;
    }
    __syncwarp();
    // O[16 x HD] += P[16 x 64] V[64 x HD]
#pragma unroll
    for (int ks = 0; ks < kAttnK / 8; ++ks) {
      uint32_t pa[4];
      pa[0] = f2tf32(Pw[g * PS + ks * 8 + t]);
      pa[1] = f2tf32(Pw[(g + 8) * PS + ks * 8 + t]);
      pa[2] = f2tf32(Pw[g * PS + ks * 8 + t + 4]);
      pa[3] = f2tf32(Pw[(g + 8) * PS + ks * 8 + t + 4]);
      const uint32_t* vr = Vs + (ks * 8 + t) * VS + g;
#pragma unroll
      for (int nb = 0; nb < KQ; ++nb) mma_tf32(o[nb], pa, vr[nb * 8], vr[4 * VS + nb * 8]);
    }
  }
  const float i0 = l0 > 0.f ? 1.f / l0 : 0.f, i1 = l1 > 0.f ? 1.f / l1 : 0.f;
  const int qa0 = q0 + warp * 16 + g, qa1 = qa0 + 8;
#pragma unroll
  for (int nb = 0; nb < KQ; ++nb) {
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
