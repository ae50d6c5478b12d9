// Non-GEMM kernels of the hot path: weight packing (one-time), operand casts / layout changes, and the
// diffusion-step embedding chain (wavenet.py:26-38 SinusoidalPosEmb, :87-91 mlp, :56,61 diffusion_projection),
// all fp32 on CUDA cores -- they are O(C^2) per step, not per frame.
#pragma once
#include "pd_gemm.cuh"

namespace pd {

// Packed weight rows are ordered so that one 256-wide N tile holds 128 gate rows followed by the 128 filter
// rows of the SAME channels (torch.chunk pairs channel c with c + C, wavenet.py:67).
__device__ __host__ __forceinline__ int gate_row_to_orig(int p, int C) {
  const int j = p >> 8, w = p & 255;
  return (w >> 7) * C + j * 128 + (w & 127);
}

template <typename T> __device__ __forceinline__ T from_float(float v);
template <> __device__ __forceinline__ __nv_bfloat16 from_float<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ __half from_float<__half>(float v) { return __float2half_rn(v); }

// dilated_conv.weight [2C, C, 3] fp32 -> [2C (gate-paired rows), 3 (tap), C] 16-bit: K index = tap * C + c.
template <typename T>
__global__ void k_pack_dil(T* __restrict__ dst, const float* __restrict__ src, int C) {
  const long long total = 2ll * C * 3 * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int p = (int)(i / (3 * C));
    const int rem = (int)(i - (long long)p * 3 * C);
    const int tap = rem / C, c = rem - tap * C;
    const int orig = gate_row_to_orig(p, C);
    dst[i] = from_float<T>(src[((long long)orig * C + c) * 3 + tap]);
  }
}

// Generic row gather: dst[r * dst_ld + dst_col0 + c] = src[(rowmap(r) + src_row0) * cols + c], zero for rows
// beyond src_rows.  gateC > 0 selects the gate-paired row order.
template <typename T>
__global__ void k_pack_rows(T* __restrict__ dst, long long dst_ld, int dst_col0, const float* __restrict__ src,
                            int src_rows, int src_row0, int rows, int cols, int gateC) {
  const long long total = (long long)rows * cols;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / cols), c = (int)(i - (long long)r * cols);
    const int sr = (gateC > 0 ? gate_row_to_orig(r, gateC) : r) + src_row0;
    const float v = sr < src_rows ? src[(long long)sr * cols + c] : 0.f;
    dst[(long long)r * dst_ld + dst_col0 + c] = from_float<T>(v);
  }
}

template <typename T>
__global__ void k_identity(T* __restrict__ dst, int C) {
  const long long total = (long long)C * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x)
    dst[i] = from_float<T>((i / C) == (i % C) ? 1.f : 0.f);
}

// fp32 vector ops used while packing biases: dst[i] = (acc ? dst[i] : 0) + a[map(i) + a0] (+ b[map(i)]), zero padded.
__global__ void k_pack_bias(float* __restrict__ dst, const float* __restrict__ a, const float* __restrict__ b, int a0,
                            int a_len, int n, int gateC, int accumulate) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int s = (gateC > 0 ? gate_row_to_orig(i, gateC) : i) + a0;
    float v = s < a_len ? a[s] : 0.f;
    if (b != nullptr && s < a_len) v += b[s];
    dst[i] = (accumulate ? dst[i] : 0.f) + v;
  }
}

// fp32 -> 16-bit cast (condition tensor), 4 elements per thread.
template <typename T>
__global__ void k_cast(T* __restrict__ dst, const float* __restrict__ src, long long n4) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(src) + i);
    uint2 o;
    o.x = Cvt<T>::pack(v.x, v.y);
    o.y = Cvt<T>::pack(v.z, v.w);
    reinterpret_cast<uint2*>(dst)[i] = o;
  }
}

// spec [B,1,M,T] fp32 (T contiguous) -> xin [B*T, Mp] 16-bit (mel contiguous): the input-projection operand.
template <typename T>
__global__ void k_spec_to_xin(T* __restrict__ xin, const float* __restrict__ spec, int M, int Mp, int Tn) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z, m0 = blockIdx.y * 32, t0 = blockIdx.x * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
  for (int i = ty; i < 32; i += 8) {
    const int m = m0 + i, t = t0 + tx;
    tile[i][tx] = (m < M && t < Tn) ? spec[((long long)b * M + m) * Tn + t] : 0.f;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int t = t0 + i, m = m0 + tx;
    if (t < Tn && m < Mp) xin[((long long)b * Tn + t) * Mp + m] = from_float<T>(tile[tx][i]);
  }
}

// One stage of the rectified-flow integrators (reflow.py:48-84): y = x + (((sum_j w_j k_j) * mult) * dt) / div, the same
// operation order as the reference expressions (e.g. 0.0625 * (3 k1 + 9 k4) * dt, or (...) * dt / 7), with explicit
// round-to-nearest mul/add so nothing is contracted into an fma.  Writes y as fp32 in the state layout [B,1,M,T] (x_out,
// optional; may alias x) and/or as the 16-bit, frame-major input-projection operand xin [B*T, Mp] (optional).
struct RkStage {
  const float* k[6];
  float w[6];
  int n_k;
  float mult, dt, div;
};
template <typename T>
__global__ void k_rk_stage(T* __restrict__ xin, float* x_out, const float* x, RkStage st, int M, int Mp, int Tn) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z, m0 = blockIdx.y * 32, t0 = blockIdx.x * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
  for (int i = ty; i < 32; i += 8) {
    const int m = m0 + i, t = t0 + tx;
    float y = 0.f;
    if (m < M && t < Tn) {
      const long long idx = ((long long)b * M + m) * Tn + t;
      float sum = __fmul_rn(st.w[0], st.k[0][idx]);
      for (int j = 1; j < st.n_k; ++j) sum = __fadd_rn(sum, __fmul_rn(st.w[j], st.k[j][idx]));
      y = __fadd_rn(x[idx], __fdiv_rn(__fmul_rn(__fmul_rn(sum, st.mult), st.dt), st.div));
      if (x_out != nullptr) x_out[idx] = y;
    }
    tile[i][tx] = y;
  }
  if (xin == nullptr) return;
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int t = t0 + i, m = m0 + tx;
    if (t < Tn && m < Mp) xin[((long long)b * Tn + t) * Mp + m] = from_float<T>(tile[tx][i]);
  }
}

// RectifiedFlow.denorm_spec (reflow.py:106-107) in place on the state [B,1,M,T]: (x + 1) / 2 * (max - min) + min per mel bin.
__global__ void k_denorm_spec(float* __restrict__ x, const float* __restrict__ spec_min, const float* __restrict__ spec_max,
                              int n_spec, int M, int Tn, long long total, float out_scale) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int m = (int)((i / Tn) % M);
    const int j = n_spec == 1 ? 0 : m;
    const float lo = spec_min[j], hi = spec_max[j];
    const float v = __fadd_rn(__fmul_rn(__fdiv_rn(__fadd_rn(x[i], 1.f), 2.f), __fadd_rn(hi, -lo)), lo);
    x[i] = out_scale != 1.f ? __fmul_rn(v, out_scale) : v;   // vocoder hand-off (pd_plan_set_output_scale)
  }
}

// Frame-level half of ProDiffTeacher.forward_condition (prodiff_teacher.py:127-146): length regulation
// (F.pad + gather by mel2ph), + pitch embedding Linear(1,H) of log(1 + f0 / 700) (:98-101), + up to two utterance-level
// embeddings (speaker, gender: [B,H] rows, :134-137), + the sum of up to two variance embeddings Linear(1,H) (voicing,
// breath, :139-147), + an optional frame-level embedding ([B,T,H], the frame-wise speaker mixes of the web handler), all
// times the non-padding mask (mel2ph > 0, :148-149) -- in the reference's order of fp32 additions.  Writes the fp32
// condition (optional) and the 16-bit operand of the condition-projection GEMM directly, so the [B,T,H] fp32 tensor never
// has to be materialised, cast and re-read.  One thread per (frame, 4 channels).
struct CondAssemble {
  const float* enc;        // encoder_out [B, Tph, H]
  const long long* mel2ph; // [B, T], 0 = padding, i > 0 -> token i - 1
  const float* f0;         // [B, T]
  const float* w_pitch;    // [H] (Linear(1,H).weight[:, 0])
  const float* b_pitch;    // [H]
  const float* utt[2];     // optional [B, H]
  const float* var_x[2];   // optional [B, T]
  const float* var_w[2];   // [H]
  const float* var_b[2];   // [H]
  const float* frame;      // optional [B, T, H]
  int Tph, T, H, n_var;
};
template <typename T>
__global__ void k_assemble_cond(T* __restrict__ cond16, float* __restrict__ cond32, CondAssemble a, long long total4) {
  const int H4 = a.H >> 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
    const long long n = i / H4;            // frame index b * T + t
    const int h = (int)(i - n * H4) * 4;
    const int b = (int)(n / a.T);
    long long tok = a.mel2ph[n];
    if (tok > a.Tph) tok = 0;              // out of range (the reference's gather would raise): handled as padding, never read
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (tok > 0) {
      const float4 e = __ldg(reinterpret_cast<const float4*>(a.enc + ((long long)b * a.Tph + (tok - 1)) * a.H + h));
      v[0] = e.x; v[1] = e.y; v[2] = e.z; v[3] = e.w;
    }
    const float f0_mel = logf(1.f + a.f0[n] / 700.f);
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = __fadd_rn(v[u], __fadd_rn(__fmul_rn(f0_mel, a.w_pitch[h + u]), a.b_pitch[h + u]));
#pragma unroll
    for (int j = 0; j < 2; ++j)
      if (a.utt[j] != nullptr) {
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = __fadd_rn(v[u], a.utt[j][(long long)b * a.H + h + u]);
      }
    if (a.frame != nullptr) {
      const float4 f = __ldg(reinterpret_cast<const float4*>(a.frame + n * a.H + h));
      v[0] = __fadd_rn(v[0], f.x); v[1] = __fadd_rn(v[1], f.y); v[2] = __fadd_rn(v[2], f.z); v[3] = __fadd_rn(v[3], f.w);
    }
    if (a.n_var > 0) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float s = __fadd_rn(__fmul_rn(a.var_x[0][n], a.var_w[0][h + u]), a.var_b[0][h + u]);
        if (a.n_var > 1) s = __fadd_rn(s, __fadd_rn(__fmul_rn(a.var_x[1][n], a.var_w[1][h + u]), a.var_b[1][h + u]));
        v[u] = __fadd_rn(v[u], s);
      }
    }
    if (tok <= 0) {
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = 0.f;   // nonpadding mask
    }
    if (cond32 != nullptr) reinterpret_cast<float4*>(cond32)[i] = make_float4(v[0], v[1], v[2], v[3]);
    uint2 o;
    o.x = Cvt<T>::pack(v[0], v[1]);
    o.y = Cvt<T>::pack(v[2], v[3]);
    reinterpret_cast<uint2*>(cond16)[i] = o;
  }
}

struct StepVals {
  float v[64];
};
__global__ void k_set_steps(float* __restrict__ dst, StepVals vals, int n) {
  if (threadIdx.x < n) dst[threadIdx.x] = vals.v[threadIdx.x];
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// SinusoidalPosEmb (wavenet.py:31-38): emb[e] = [sin(step * f_j), cos(step * f_j)], f_j = exp(-j ln(1e4) / (C/2 - 1)).
// Accurate libm calls (no fast-math): this feeds every layer's pre-activation.
__global__ void k_step_sinusoid(const float* __restrict__ steps, int C, float* __restrict__ emb0) {
  const int e = blockIdx.x;
  const float step = steps[e];
  const int half = C / 2;
  const float neg = -(logf(10000.f) / (float)(half - 1));
  for (int j = threadIdx.x; j < half; j += blockDim.x) {
    const float a = step * expf((float)j * neg);
    emb0[(long long)e * C + j] = sinf(a);
    emb0[(long long)e * C + half + j] = cosf(a);
  }
}

// Batched fp32 mat-vec for the step chain: y[e][l][o] = act(W_l[o][:] . x[e][:] + b_l[o]).
// grid (ceil(O / 32), E, L), block 256: 8 warps x 4 outputs, the input vector staged in shared memory.
// act 0 = identity, 1 = Mish (x * tanh(softplus(x)), wavenet.py:22-24; softplus threshold 20).
__global__ void k_step_matvec(const float* __restrict__ W, const float* __restrict__ bias, const float* __restrict__ x,
                              float* __restrict__ y, int O, int I, int L, int act) {
  extern __shared__ float s_x[];
  const int e = blockIdx.y, l = blockIdx.z;
  for (int c = threadIdx.x; c < I; c += blockDim.x) s_x[c] = x[(long long)e * I + c];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* w = W + (long long)l * O * I;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int o = blockIdx.x * 32 + warp * 4 + r;
    if (o >= O) break;
    float acc = 0.f;
    for (int c = lane; c < I; c += 32) acc = fmaf(__ldg(w + (long long)o * I + c), s_x[c], acc);
    acc = warp_sum(acc);
    if (lane == 0) {
      float v = acc + bias[(long long)l * O + o];
      if (act == 1) {
        const float sp = v > 20.f ? v : log1pf(expf(v));
        v = v * tanhf(sp);
      }
      y[((long long)e * L + l) * O + o] = v;
    }
  }
}

// The reference adds d to x BEFORE the zero-padded dilated conv (wavenet.py:63-65), so the step contributes
// W_k . d through each tap k that lands inside the utterance.  Per (entry, layer) this kernel emits, in packed
// column order, the four sums a frame can need: slot 0 = all three taps (interior frames), slot 1 = taps 1+2 (the first
// `dilation` frames: tap 0 fell on padding), slot 2 = taps 0+1 (the last `dilation` frames), slot 3 = tap 1 only
// (utterances shorter than the dilation).  The gate epilogue just picks a slot per frame.  fp32 weights: exact.
// grid (2C / 8, L), block 256: one warp per packed output row, looping over all entries with the row in registers.
template <int CPL>  // C / 32
__global__ void k_step_taps(const float* __restrict__ dvec, const float* __restrict__ dil_w, float* __restrict__ sb,
                            int E, int L) {
  constexpr int C = CPL * 32;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int p = blockIdx.x * 8 + warp, l = blockIdx.y;
  const int orig = gate_row_to_orig(p, C);
  const float* w = dil_w + ((long long)l * 2 * C + orig) * C * 3;
  float w0[CPL], w1[CPL], w2[CPL];
#pragma unroll
  for (int i = 0; i < CPL; ++i) {
    const int c = lane + 32 * i;
    w0[i] = w[c * 3 + 0];
    w1[i] = w[c * 3 + 1];
    w2[i] = w[c * 3 + 2];
  }
  for (int e = 0; e < E; ++e) {
    const float* d = dvec + ((long long)e * L + l) * C;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
#pragma unroll
    for (int i = 0; i < CPL; ++i) {
      const float dv = __ldg(d + lane + 32 * i);
      a0 = fmaf(w0[i], dv, a0);
      a1 = fmaf(w1[i], dv, a1);
      a2 = fmaf(w2[i], dv, a2);
    }
    a0 = warp_sum(a0);
    a1 = warp_sum(a1);
    a2 = warp_sum(a2);
    if (lane == 0) {
      float* o = sb + ((long long)e * L + l) * 4 * 2 * C;
      o[p] = a0 + a1 + a2;
      o[2 * C + p] = a1 + a2;
      o[4 * C + p] = a0 + a1;
      o[6 * C + p] = a1;
    }
  }
}

}  // namespace pd
