// C-ABI host side of the FastSpeech text encoder (include/prodiff_b200.h, "text encoder" section); included at the end of
// pd_api.cu so that it shares the tensor-map, GEMM-launch and error helpers.  Launch sequence per call (tts_modules.py:305-312,
// 262-288; common_layers.py:649-673): embed -> L x { LN1 -> qkv GEMM -> attention -> out_proj GEMM (+x, keep) -> LN2 ->
// 9-tap conv GEMM (* k^-0.5, GELU) -> ffn_2 GEMM (+bias, +x, keep) } -> last LN (* keep).
#pragma once
#include "pd_encoder.cuh"

struct EncLayout {
  size_t w_qkv, w_out, w_f1, w_f2;        // 16-bit, per layer strides below
  size_t b_f1, b_f2, zeros;               // fp32
  size_t ln, emb;                         // fp32: [L][4][H] (ln1 w, ln1 b, ln2 w, ln2 b) + [2][H] last LN; embedding [V][H]
  size_t total;
};
static EncLayout enc_layout(const pd_encoder_config& c) {
  const size_t H = c.hidden_size, L = c.num_layers, k = c.kernel_size, V = c.vocab_size;
  EncLayout p;
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o = align_up(o + bytes, 1024); return r; };
  p.w_qkv = take(L * 3 * H * H * 2);
  p.w_out = take(L * H * H * 2);
  p.w_f1 = take(L * 4 * H * k * H * 2);
  p.w_f2 = take(L * H * 4 * H * 2);
  p.b_f1 = take(L * 4 * H * 4);
  p.b_f2 = take(L * H * 4);
  p.zeros = take(3 * H * 4);
  p.ln = take((L * 4 + 2) * H * 4);
  p.emb = take(V * H * 4);
  p.total = o;
  return p;
}

static int check_enc_config(const pd_encoder_config* c) {
  if (!c) return fail("null encoder config");
  if (c->hidden_size != kEncH) return fail("encoder hidden_size must be %d (got %d)", kEncH, c->hidden_size);
  if (c->num_heads != 2 && c->num_heads != 4) return fail("encoder num_heads must be 2 or 4 (got %d)", c->num_heads);
  if (c->num_layers < 1) return fail("encoder num_layers must be >= 1");
  if (c->kernel_size < 1 || c->kernel_size % 2 == 0 || c->kernel_size > 31) return fail("encoder kernel_size must be odd and <= 31 (got %d)", c->kernel_size);
  if (c->vocab_size < 2) return fail("encoder vocab_size must be >= 2");
  if (c->dtype != PD_BF16 && c->dtype != PD_FP16) return fail("dtype must be PD_BF16 or PD_FP16");
  return 0;
}

struct pd_encoder {
  pd_encoder_config cfg;
  EncLayout lay;
  uint8_t* packed;
  int device, sm_count;
  CUtensorMap tm_w_qkv, tm_w_out, tm_w_f1, tm_w_f2;   // all layers stacked along the row dimension
};

extern "C" int pd_encoder_state_count(const pd_encoder_config* cfg) { return cfg ? 3 + 10 * cfg->num_layers : 0; }
extern "C" size_t pd_encoder_packed_bytes(const pd_encoder_config* cfg) {
  if (check_enc_config(cfg)) return 0;
  return enc_layout(*cfg).total;
}

template <typename T>
static int enc_pack_all(const pd_encoder_config& c, const EncLayout& lay, uint8_t* pk, const float* const* st, cudaStream_t s) {
  const int H = c.hidden_size, L = c.num_layers, k = c.kernel_size;
  const int TB = 256;
  auto blocks = [&](long long n) { return (int)((n + TB - 1) / TB < 4096 ? (n + TB - 1) / TB : 4096); };
  PD_CUDA(cudaMemsetAsync(pk + lay.zeros, 0, (size_t)3 * H * 4, s));
  PD_CUDA(cudaMemcpyAsync(pk + lay.emb, st[0], (size_t)c.vocab_size * H * 4, cudaMemcpyDeviceToDevice, s));
  PD_CUDA(cudaMemcpyAsync(pk + lay.ln + (size_t)L * 4 * H * 4, st[1], (size_t)H * 4, cudaMemcpyDeviceToDevice, s));
  PD_CUDA(cudaMemcpyAsync(pk + lay.ln + (size_t)(L * 4 + 1) * H * 4, st[2], (size_t)H * 4, cudaMemcpyDeviceToDevice, s));
  for (int l = 0; l < L; ++l) {
    const float* const* q = st + 3 + 10 * l;
    float* ln = (float*)(pk + lay.ln) + (size_t)l * 4 * H;
    PD_CUDA(cudaMemcpyAsync(ln, q[0], (size_t)H * 4, cudaMemcpyDeviceToDevice, s));
    PD_CUDA(cudaMemcpyAsync(ln + H, q[1], (size_t)H * 4, cudaMemcpyDeviceToDevice, s));
    PD_CUDA(cudaMemcpyAsync(ln + 2 * H, q[4], (size_t)H * 4, cudaMemcpyDeviceToDevice, s));
    PD_CUDA(cudaMemcpyAsync(ln + 3 * H, q[5], (size_t)H * 4, cudaMemcpyDeviceToDevice, s));
    k_pack_rows<T><<<blocks(3ll * H * H), TB, 0, s>>>((T*)(pk + lay.w_qkv) + (size_t)l * 3 * H * H, H, 0, q[2], 3 * H, 0, 3 * H, H, 0);
    PD_LAUNCH_CHECK("k_pack_rows(in_proj)");
    k_pack_rows<T><<<blocks((long long)H * H), TB, 0, s>>>((T*)(pk + lay.w_out) + (size_t)l * H * H, H, 0, q[3], H, 0, H, H, 0);
    PD_LAUNCH_CHECK("k_pack_rows(out_proj)");
    k_pack_conv_taps<T><<<blocks(4ll * H * k * H), TB, 0, s>>>((T*)(pk + lay.w_f1) + (size_t)l * 4 * H * k * H, q[6], 4 * H, H, k);
    PD_LAUNCH_CHECK("k_pack_conv_taps(ffn_1)");
    k_pack_rows<T><<<blocks(4ll * H * H), TB, 0, s>>>((T*)(pk + lay.w_f2) + (size_t)l * H * 4 * H, 4 * H, 0, q[8], H, 0, H, 4 * H, 0);
    PD_LAUNCH_CHECK("k_pack_rows(ffn_2)");
    PD_CUDA(cudaMemcpyAsync(pk + lay.b_f1 + (size_t)l * 4 * H * 4, q[7], (size_t)4 * H * 4, cudaMemcpyDeviceToDevice, s));
    PD_CUDA(cudaMemcpyAsync(pk + lay.b_f2 + (size_t)l * H * 4, q[9], (size_t)H * 4, cudaMemcpyDeviceToDevice, s));
  }
  return 0;
}

extern "C" int pd_encoder_create(const pd_encoder_config* cfg, const float* const* state, int n_state, void* packed,
                                 size_t packed_bytes, void* stream, pd_encoder** out) {
  if (!out) return fail("null out pointer");
  *out = nullptr;
  if (check_enc_config(cfg)) return 1;
  if (!state || n_state != pd_encoder_state_count(cfg)) return fail("expected %d encoder state tensors, got %d", pd_encoder_state_count(cfg), n_state);
  for (int i = 0; i < n_state; ++i)
    if (!state[i]) return fail("encoder state tensor %d is null", i);
  const EncLayout lay = enc_layout(*cfg);
  if (!packed || packed_bytes < lay.total) return fail("packed buffer too small: need %zu bytes, got %zu", lay.total, packed_bytes);
  if (reinterpret_cast<uintptr_t>(packed) % 1024) return fail("packed buffer must be 1024-byte aligned");
  int device = 0, sms = 0;
  if (check_device(&device, &sms)) return 1;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  uint8_t* pk = static_cast<uint8_t*>(packed);
  const int rc = cfg->dtype == PD_FP16 ? enc_pack_all<__half>(*cfg, lay, pk, state, s) : enc_pack_all<__nv_bfloat16>(*cfg, lay, pk, state, s);
  if (rc) return rc;
  pd_encoder* e = new pd_encoder();
  e->cfg = *cfg;
  e->lay = lay;
  e->packed = pk;
  e->device = device;
  e->sm_count = sms;
  const int H = cfg->hidden_size, L = cfg->num_layers, k = cfg->kernel_size, dt = cfg->dtype;
  if (make_w_map(&e->tm_w_qkv, dt, pk + lay.w_qkv, H, L * 3 * H, 256) || make_w_map(&e->tm_w_out, dt, pk + lay.w_out, H, L * H, 256) ||
      make_w_map(&e->tm_w_f1, dt, pk + lay.w_f1, k * H, L * 4 * H, 256) || make_w_map(&e->tm_w_f2, dt, pk + lay.w_f2, 4 * H, L * H, 256)) {
    delete e;
    return 1;
  }
  *out = e;
  return 0;
}
extern "C" void pd_encoder_destroy(pd_encoder* e) { delete e; }

// Launch with the programmatic-dependent-launch attribute (pd_ptx.cuh): every encoder kernel but the first (g_enc_pdl is
// declared next to launch_gemm_t in pd_api.cu).
template <typename... KArgs, typename... Args>
static cudaError_t launch_dep(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_enc_pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

struct EncScratch {
  size_t x, xn, qkv, att, h1, keep, total;
};
static EncScratch enc_scratch(int H, long long N) {
  EncScratch w;
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o = align_up(o + bytes, 1024); return r; };
  w.x = take((size_t)N * H * 4);
  w.xn = take((size_t)N * H * 2);
  w.qkv = take((size_t)N * 3 * H * 2);
  w.att = take((size_t)N * H * 2);
  w.h1 = take((size_t)N * 4 * H * 2);
  w.keep = take((size_t)N);
  w.total = o;
  return w;
}
extern "C" size_t pd_encoder_workspace_bytes(const pd_encoder* e, int B, int T) {
  if (!e || B < 1 || T < 1) return 0;
  return enc_scratch(e->cfg.hidden_size, (long long)B * T).total;
}

template <typename T, int HD>
static int launch_attention_f16(T* out16, const __half* qkv16, const uint8_t* keep, int B, int Tn, int heads, cudaStream_t s) {
  auto kern = k_enc_attention_f16<T, HD>;
  const int smem = AttnCfg<HD>::kSmemBytes;
  static bool configured[64] = {false};
  int dev = 0;
  PD_CUDA(cudaGetDevice(&dev));
  if (!configured[dev & 63]) {
    PD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured[dev & 63] = true;
  }
  PD_CUDA(launch_dep(kern, dim3((Tn + kHQ - 1) / kHQ, heads, B), dim3(128), (size_t)smem, s, out16, qkv16, keep, Tn, sqrtf(1.0f / (float)HD)));
  PD_LAUNCH_CHECK("k_enc_attention_f16");
  return 0;
}
template <typename T>
static int enc_attention(T* out16, const __half* qkv16, const uint8_t* keep, int B, int Tn, int heads, cudaStream_t s) {
  return kEncH / heads == 128 ? launch_attention_f16<T, 128>(out16, qkv16, keep, B, Tn, heads, s)
                              : launch_attention_f16<T, 64>(out16, qkv16, keep, B, Tn, heads, s);
}

template <typename T>
static int enc_forward_t(pd_encoder* e, const long long* tok, const float* extra, float* out, uint8_t* ws, int B, int Tn, cudaStream_t s) {
  const pd_encoder_config& c = e->cfg;
  const int H = c.hidden_size, L = c.num_layers, k = c.kernel_size, dt = c.dtype, sms = e->sm_count;
  const long long N = (long long)B * Tn;
  const EncScratch w = enc_scratch(H, N);
  float* x = reinterpret_cast<float*>(ws + w.x);
  T* xn = reinterpret_cast<T*>(ws + w.xn);
  __half* qkv = reinterpret_cast<__half*>(ws + w.qkv);   // fp16 whatever the GEMM operand type (see k_enc_attention_f16)
  T* att = reinterpret_cast<T*>(ws + w.att);
  T* h1 = reinterpret_cast<T*>(ws + w.h1);
  uint8_t* keep = ws + w.keep;
  const uint8_t* pk = e->packed;
  const float* ln = reinterpret_cast<const float*>(pk + e->lay.ln);
  CUtensorMap tm_xn, tm_att, tm_h1;
  if (make_act_map(&tm_xn, dt, xn, H, Tn, B, 1) || make_act_map(&tm_att, dt, att, H, Tn, B, 1) || make_act_map(&tm_h1, dt, h1, 4 * H, Tn, B, 1))
    return 1;
  g_enc_pdl = false;
  k_enc_embed<<<dim3((Tn + 7) / 8, B), 256, 0, s>>>(x, keep, tok, reinterpret_cast<const float*>(pk + e->lay.emb), extra, Tn, c.vocab_size);
  PD_LAUNCH_CHECK("k_enc_embed");
  static int pdl = -1;   // PD_ENC_PDL=0: ordinary stream-ordered launches (A/B)
  if (pdl < 0) { const char* ev = getenv("PD_ENC_PDL"); pdl = (ev && ev[0] == '0') ? 0 : 1; }
  g_enc_pdl = pdl != 0;
  const int ln_blocks = (int)((N + 7) / 8);
  for (int l = 0; l < L; ++l) {
    const float* lnl = ln + (size_t)l * 4 * H;
    PD_CUDA(launch_dep(k_enc_layernorm<T>, dim3(ln_blocks), dim3(256), 0, s, xn, (float*)nullptr, (const float*)x, lnl, lnl + H, (const uint8_t*)nullptr, N));
    PD_LAUNCH_CHECK("k_enc_layernorm(1)");
    {  // packed in-projection, bias-free (common_layers.py:195,640-641) -> fp16 q | k | v
      GemmShape sh = base_shape(B, Tn, 3 * H / 256, H / 64);
      sh.b_row0 = l * 3 * H;
      EpiArgs ep = base_epi();
      ep.out = qkv;
      ep.bias = reinterpret_cast<const float*>(pk + e->lay.zeros);
      ep.ld = 3 * H;
      ep.out_type = OUT_F16;
      if (launch_gemm<EPI_STORE_X, 256>(dt, tm_xn, e->tm_w_qkv, sh, ep, sms, s, "gemm(enc in_proj)", 11)) return 1;
    }
    if (enc_attention<T>(att, qkv, keep, B, Tn, c.num_heads, s)) return 1;
    {  // out_proj (bias-free) + residual, then the padding mask (common_layers.py:665-667)
      GemmShape sh = base_shape(B, Tn, H / 256, H / 64);
      sh.b_row0 = l * H;
      EpiArgs ep = base_epi();
      ep.out = x;
      ep.bias = reinterpret_cast<const float*>(pk + e->lay.zeros);
      ep.ld = H;
      ep.out_type = OUT_F32;
      ep.resid = x;
      ep.keep = keep;
      if (launch_gemm<EPI_STORE_X, 256>(dt, tm_att, e->tm_w_out, sh, ep, sms, s, "gemm(enc out_proj)", 11)) return 1;
    }
    PD_CUDA(launch_dep(k_enc_layernorm<T>, dim3(ln_blocks), dim3(256), 0, s, xn, (float*)nullptr, (const float*)x, lnl + 2 * H, lnl + 3 * H, (const uint8_t*)nullptr, N));
    PD_LAUNCH_CHECK("k_enc_layernorm(2)");
    {  // ffn_1: Conv1d(H, 4H, k, padding = k // 2) as an implicit GEMM over k taps, * k^-0.5, GELU (common_layers.py:570-576)
      GemmShape sh = base_shape(B, Tn, 4 * H / 256, H / 64);
      sh.taps = k;
      sh.tap_mode = 1;
      sh.dil = 1;
      sh.b_row0 = l * 4 * H;
      EpiArgs ep = base_epi();
      ep.out = h1;
      ep.bias = reinterpret_cast<const float*>(pk + e->lay.b_f1) + (size_t)l * 4 * H;
      ep.ld = 4 * H;
      ep.scale = 1.0f / sqrtf((float)k);
      ep.gelu = 1;
      // In-kernel timeline (tools/encoder_trace.py): 630 cycles per K block = 4 MMAs of M = 128, N = 256, K = 16, whose
      // issue floor is 128 cycles each (4096 MAC/clk/SM), so the K loop runs at ~81 % of the tensor peak; the erf-GELU
      // epilogue (11-15 k cycles per tile) hides behind the next tile's K loop.  What the launch loses is at tile granularity:
      // 512 tiles on 148 CTAs are 3.46 waves (4 tiles on the busiest CTAs) and the last tile's epilogue is exposed -- 0.85
      // PFLOP/s per launch.  Measured and not adopted: the 9 taps from one resident haloed A tile (a third fewer TMA bytes:
      // K block 621 cycles, one ring stage less, 0.765 vs 0.737 ms per forward) and cta_group::2 pairs (654 cycles, 0.740 ms).
      if (launch_gemm<EPI_STORE_XH, 256>(dt, tm_xn, e->tm_w_f1, sh, ep, sms, s, "gemm(enc ffn_1)", 12)) return 1;
    }
    {  // ffn_2 + bias + residual, then the padding mask (common_layers.py:582, 669-672)
      GemmShape sh = base_shape(B, Tn, H / 256, 4 * H / 64);
      sh.b_row0 = l * H;
      EpiArgs ep = base_epi();
      ep.out = x;
      ep.bias = reinterpret_cast<const float*>(pk + e->lay.b_f2) + (size_t)l * H;
      ep.ld = H;
      ep.out_type = OUT_F32;
      ep.resid = x;
      ep.keep = keep;
      if (launch_gemm<EPI_STORE_X, 256>(dt, tm_h1, e->tm_w_f2, sh, ep, sms, s, "gemm(enc ffn_2)", 11)) return 1;
    }
  }
  const float* lnf = ln + (size_t)L * 4 * H;
  PD_CUDA(launch_dep(k_enc_layernorm<T>, dim3(ln_blocks), dim3(256), 0, s, (T*)nullptr, out, (const float*)x, lnf, lnf + H, (const uint8_t*)keep, N));
  PD_LAUNCH_CHECK("k_enc_layernorm(last)");
  g_enc_pdl = false;
  return 0;
}

extern "C" int pd_encoder_forward(pd_encoder* e, const long long* txt_tokens, const float* extra_embed, float* out, int B, int T,
                                  void* workspace, size_t workspace_bytes, void* stream) {
  if (!e || !txt_tokens || !out || !workspace) return fail("null argument");
  if (B < 1 || T < 1) return fail("B and T must be positive");
  if (B > 65535) return fail("B=%d exceeds the grid's y dimension", B);
  const size_t need = pd_encoder_workspace_bytes(e, B, T);
  if (workspace_bytes < need) return fail("encoder workspace too small: need %zu bytes, got %zu", need, workspace_bytes);
  if (reinterpret_cast<uintptr_t>(workspace) % 1024) return fail("workspace must be 1024-byte aligned");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  return e->cfg.dtype == PD_FP16 ? enc_forward_t<__half>(e, txt_tokens, extra_embed, out, ws, B, T, s)
                                 : enc_forward_t<__nv_bfloat16>(e, txt_tokens, extra_embed, out, ws, B, T, s);
}
