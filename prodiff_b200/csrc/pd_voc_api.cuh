// C-ABI host side of the NSF-HiFiGAN generator: pd_vocoder_* of include/prodiff_b200.h.
// Included at the end of pd_api.cu (uses its error / launch-count helpers and the tensor-map encoder).
//
// Reference: modules/nsf_hifigan/models.py -- Generator.forward :269-290, ResBlock1 :39-76, SineGen :103-188,
// SourceModuleHnNSF :191-216 -- and the hand-off component/vocoder/nsf_hifigan.py:49-58.

namespace {

constexpr int kVocG = 32;         // guard rows in front of / behind every utterance: >= (11 - 1) / 2 * 5 = 25, the widest reach
constexpr int kVocMaxUps = 8;
constexpr int kVocMaxRes = 4;

struct VocConv {        // one packed convolution
  size_t w_off, b_off;  // byte offsets in the packed block (fp16 weights [n_total][k_total], fp32 bias [n_total])
  int cin, n_total, taps, taps_pad, dil, bn;
  CUtensorMap tm_w;
  CUtensorMap tm_w_half;   // half-height boxes (bn / 2 rows) for the cta_group::2 form; valid when bn >= 128
};

// 4-D activation map over a guarded channels-last buffer (see pd_voc.cuh): dims (C, taps, T, B), overlapping strides.
int make_voc_a_map(CUtensorMap* m, const void* buf, int C, int taps, int dil, int T, int B) {
  if (C < 64) {   // gathered by the kernel (pd_voc.cuh); the kernel parameter still has to be a well-formed object
    memset(m, 0, sizeof(*m));
    return 0;
  }
  EncodeTiledFn enc = get_encode();
  if (!enc) return fail("cuTensorMapEncodeTiled entry point unavailable (driver too old?)");
  const int half = (taps - 1) / 2;
  const int g = C >= 64 ? 1 : 64 / C;
  cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)taps, (cuuint64_t)T, (cuuint64_t)B};
  cuuint64_t gstr[3] = {(cuuint64_t)dil * C * 2, (cuuint64_t)C * 2, (cuuint64_t)(T + 2 * kVocG) * C * 2};
  cuuint32_t box[4] = {(cuuint32_t)(C >= 64 ? 64 : C), (cuuint32_t)g, (cuuint32_t)kBlockM, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  const uint8_t* base = static_cast<const uint8_t*>(buf) + (size_t)(kVocG - half * dil) * C * 2;
  const CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<uint8_t*>(base), gdim, gstr, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled (vocoder activation map C=%d taps=%d dil=%d T=%d B=%d) failed with CUresult %d", C, taps, dil, T, B, (int)r);
  return 0;
}

template <int BN, bool kGather, bool kPair = false>
int launch_voc_t(const CUtensorMap& a, const CUtensorMap& w, const VocShape& sh, const VocEpi& ep, int sms, cudaStream_t s) {
  auto kern = voc_conv_kernel<BN, kGather, kPair>;
  using Cfg = VocCfg<BN, kPair>;
  static bool configured[64] = {false};
  int dev = 0;
  PD_CUDA(cudaGetDevice(&dev));
  if (!configured[dev & 63]) {
    PD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::kSmemBytes));
    configured[dev & 63] = true;
  }
  const int m_tiles = sh.B * sh.m_tiles_per_row;
  const int items = (kPair ? (m_tiles + 1) / 2 : m_tiles) * sh.n_tiles;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(kPair ? 2 * (items < sms / 2 ? items : sms / 2) : (items < sms ? items : sms));
  cfg.blockDim = dim3(kGemmThreads + (kGather ? kVocGatherThreads : 0));
  cfg.dynamicSmemBytes = Cfg::kSmemBytes;
  cfg.stream = s;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  attr[1].id = cudaLaunchAttributeClusterDimension;
  attr[1].val.clusterDim.x = 2;
  attr[1].val.clusterDim.y = 1;
  attr[1].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = kPair ? 2 : 1;
  prof_before(13, s);
  PD_CUDA(cudaLaunchKernelEx(&cfg, kern, a, w, sh, ep));
  prof_after(s);
  PD_LAUNCH_CHECK("voc_conv_kernel");
  return 0;
}

int g_voc_pair = -1;       // -1: read PD_VOC_PAIR from the environment on first use (default on)
bool voc_pair_mode() {
  if (g_voc_pair < 0) {
    const char* e = getenv("PD_VOC_PAIR");
    g_voc_pair = (e && e[0] == '0') ? 0 : 1;
  }
  return g_voc_pair == 1;
}

// `abuf`: the guarded activation buffer the convolution reads (its first element).  C >= 64: read through the 4-D map `a`;
// C < 64: gathered by the kernel's extra producer warps (the map is unused).  Tiles of 128 / 256 output channels with at
// least two position tiles run as cta_group::2 pairs.
int launch_voc(const VocConv& cv, const CUtensorMap& a, const __half* abuf, int B, int T, VocEpi ep, int sms, cudaStream_t s) {
  VocShape sh;
  memset(&sh, 0, sizeof(sh));
  sh.B = B;
  sh.T = T;
  sh.m_tiles_per_row = (T + kBlockM - 1) / kBlockM;
  sh.n_tiles = cv.n_total / cv.bn;
  sh.g = cv.cin >= 64 ? 1 : 64 / cv.cin;
  sh.cc_blocks = cv.cin >= 64 ? cv.cin / 64 : 1;
  sh.tap_groups = cv.taps_pad / sh.g;
  sh.idesc = make_idesc(PD_FP16, cv.bn);
  sh.a_ptr = abuf;
  sh.a_C = cv.cin;
  sh.a_taps = cv.taps;
  sh.a_dil = cv.dil;
  sh.a_G = kVocG;
  sh.trace = g_trace ? g_trace + 13 * 1024 : nullptr;
  const bool gather = cv.cin < 64;
  if (!gather && cv.bn >= 128 && voc_pair_mode() && B * sh.m_tiles_per_row >= 2) {
    sh.idesc = make_idesc(PD_FP16, cv.bn, 256);
    return cv.bn == 128 ? launch_voc_t<128, false, true>(a, cv.tm_w_half, sh, ep, sms, s) : launch_voc_t<256, false, true>(a, cv.tm_w_half, sh, ep, sms, s);
  }
  switch (cv.bn) {
    case 16: return gather ? launch_voc_t<16, true>(a, cv.tm_w, sh, ep, sms, s) : launch_voc_t<16, false>(a, cv.tm_w, sh, ep, sms, s);
    case 32: return gather ? launch_voc_t<32, true>(a, cv.tm_w, sh, ep, sms, s) : launch_voc_t<32, false>(a, cv.tm_w, sh, ep, sms, s);
    case 64: return gather ? launch_voc_t<64, true>(a, cv.tm_w, sh, ep, sms, s) : launch_voc_t<64, false>(a, cv.tm_w, sh, ep, sms, s);
    case 128: return gather ? launch_voc_t<128, true>(a, cv.tm_w, sh, ep, sms, s) : launch_voc_t<128, false>(a, cv.tm_w, sh, ep, sms, s);
    default: return gather ? launch_voc_t<256, true>(a, cv.tm_w, sh, ep, sms, s) : launch_voc_t<256, false>(a, cv.tm_w, sh, ep, sms, s);
  }
}

VocEpi voc_epi() {
  VocEpi e;
  memset(&e, 0, sizeof(e));
  e.scale = 1.f;
  e.res_inv_slope = 10.f;
  return e;
}

bool voc_pow2_ch(int c) { return c == 16 || c == 32 || c == 64 || c == 128 || c == 256 || c == 512 || c == 1024; }

// ---- fused ResBlock rounds (pd_voc_pair.cuh): stages with <= 64 channels
int g_voc_fused = -1;      // -1: read PD_VOC_FUSED from the environment on first use (default on)
bool voc_fused_enabled() {
  if (g_voc_fused < 0) {
    const char* e = getenv("PD_VOC_FUSED");
    g_voc_fused = (e && e[0] == '0') ? 0 : 1;
  }
  return g_voc_fused == 1;
}
}  // namespace
extern "C" int pd_set_voc_fused(int on) {
  const int prev = g_voc_fused;
  g_voc_fused = on ? 1 : 0;
  return prev;
}
namespace {
bool voc_pair_ok(int C, int k) { return (C == 16 || C == 32 || C == 64) && k % 2 == 1 && k <= 15; }

template <int C, int NB, bool kRes, int kXB, int EW, int kMinCtas, bool kSingle = false>
int launch_voc_pair_t(VocPairArgs a, int sms, cudaStream_t s) {
  auto kern = voc_pair_kernel<C, NB, kRes, kXB, EW, kMinCtas, kSingle>;
  constexpr int kPairThreads = (EW + 2) * 32;
  const VocPairSmem L = voc_pair_smem<C, NB, kRes, kXB, kMinCtas, kSingle>(a.k, a.dil);
  static int configured[64] = {0};
  int dev = 0;
  PD_CUDA(cudaGetDevice(&dev));
  if (configured[dev & 63] < (int)L.total) {
    PD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total));
    PD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
    configured[dev & 63] = (int)L.total;
  }
  // co-resident CTAs per SM: shared memory (227 KB per SM, 1 KB reserved per CTA), registers, TMEM columns
  static int num_regs[64] = {0};   // per instantiation, per device
  if (num_regs[dev & 63] == 0) {
    cudaFuncAttributes fa;
    PD_CUDA(cudaFuncGetAttributes(&fa, kern));
    num_regs[dev & 63] = fa.numRegs;
  }
  const int cols = std::max(32, (kXB == 2 ? 2 : 1) * NB * C);
  int occ = (int)((227u * 1024u) / (L.total + 1024u));
  const int regs_per_cta = ((num_regs[dev & 63] * 32 + 511) / 512 * 512) * (kPairThreads / 32);   // allocation unit: 512 registers per warp
  occ = std::min(occ, 65536 / regs_per_cta);
  occ = std::min(occ, 512 / cols);
  occ = std::min(occ, 8);
  if (occ < 1) return fail("voc_pair_kernel<%d>: %u bytes of shared memory do not fit", C, L.total);
  const int P = NB * 128 - (kSingle ? 0 : a.k - 1);
  a.tiles_per_row = (a.T + P - 1) / P;
  a.idesc = make_idesc(PD_FP16, C);
  a.trace = g_trace ? g_trace + 13 * 1024 : nullptr;
  const long long tiles = (long long)a.B * a.tiles_per_row;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)std::min<long long>(tiles, (long long)sms * occ));
  cfg.blockDim = dim3(kPairThreads);
  cfg.dynamicSmemBytes = L.total;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  prof_before(13, s);
  PD_CUDA(cudaLaunchKernelEx(&cfg, kern, a));
  prof_after(s);
  PD_LAUNCH_CHECK("voc_pair_kernel");
  return 0;
}
// Tile shapes (measured on B200, gpurun_out/voc_ab*.log): 16 / 32 channels keep both convolutions' weights resident, double-buffer
// the x tile and run 4 / 2 CTAs per SM; 64 channels stream one tap per ring stage (a round's weights are 48 - 176 KB) with 2 CTAs
// per SM.  Measured no faster: resident weights with one CTA per SM for 64 channels (k <= 7), 512-row tiles, 5 - 6 CTAs per SM for
// 16 channels, 3 - 4 for 32.
int launch_voc_pair(int C, const VocPairArgs& a, int sms, cudaStream_t s) {
  switch (C) {
    case 16: return launch_voc_pair_t<16, 2, true, 2, 4, 4>(a, sms, s);
    case 32: return launch_voc_pair_t<32, 2, true, 2, 8, 2>(a, sms, s);
    case 64: return launch_voc_pair_t<64, 2, false, 1, 8, 2>(a, sms, s);
    default: return fail("voc_pair_kernel: %d channels not supported", C);
  }
}
// the rate-2 transposed convolutions of the narrow stages (C_in = 2 * C_out = 32 or 64) as ONE 3-tap convolution over a resident tile
// Measured at 64 x 1000 frames: 32 -> 16 channels 1.32 -> 1.01 ms; 64 -> 32 channels 0.88 ms through the TMA-fed kernel against
// 1.02 ms here (not used); 3 instead of 4 CTAs per SM for 32 channels: no change.
bool voc_single_ok(int cin, int u) { return u == 2 && cin == 32; }
// conv_post (C -> 1, k taps) on a narrow last stage: the same form with one real output row and the fp32 / tanh epilogue
bool voc_post_ok(int cin, int k) { return cin == 16 && k % 2 == 1 && k <= 15; }
int launch_voc_single(int C, const VocPairArgs& a, int sms, cudaStream_t s) {
  if (C == 16) return launch_voc_pair_t<16, 2, true, 1, 4, 4, true>(a, sms, s);
  if (C == 32) return launch_voc_pair_t<32, 2, true, 1, 4, 4, true>(a, sms, s);
  return fail("voc_pair_kernel (single conv): %d channels not supported", C);
}

}  // namespace

struct pd_vocoder {
  pd_vocoder_config cfg;
  uint8_t* packed;
  size_t packed_total;
  int device, sm_count;
  int upp;                                   // product of the up-sampling rates (samples per mel frame)
  VocConv conv_pre, conv_post;
  VocConv ups[kVocMaxUps];
  VocConv res[kVocMaxUps][kVocMaxRes][6];    // [stage][kernel][c1_0, c2_0, c1_1, c2_1, c1_2, c2_2]
  size_t pair_w[kVocMaxUps][kVocMaxRes][3];  // fused-round weight images (pd_voc_pair.cuh) where voc_pair_ok(), else 0
  size_t ups_w[kVocMaxUps];                  // resident-tile image of the transposed conv where voc_single_ok(), else 0
  size_t post_w;                             // resident-tile image of conv_post where voc_post_ok(), else 0
  size_t src_w[kVocMaxUps], src_b[kVocMaxUps];   // fp32 noise_convs weights, transposed to [k][C] / bias [C]
  int src_k[kVocMaxUps], src_stride[kVocMaxUps], src_pad[kVocMaxUps];
  size_t lin_w, lin_b;                       // m_source.l_linear
};

static int voc_check_config(const pd_vocoder_config* c) {
  if (!c) return fail("null config");
  if (c->num_mels < 64 || c->num_mels % 64 || c->num_mels > 512) return fail("num_mels must be a multiple of 64 in [64, 512] (got %d)", c->num_mels);
  if (c->n_upsamples < 1 || c->n_upsamples > kVocMaxUps) return fail("n_upsamples out of range");
  if (c->n_resblock_kernels < 1 || c->n_resblock_kernels > kVocMaxRes) return fail("n_resblock_kernels out of range");
  if (c->resblock_type != 1) return fail("only resblock '1' (ResBlock1, models.py:39) is implemented");
  if (!voc_pow2_ch(c->upsample_initial_channel)) return fail("upsample_initial_channel must be a power of two in [16, 1024]");
  int ch = c->upsample_initial_channel;
  for (int i = 0; i < c->n_upsamples; ++i) {
    ch /= 2;
    if (ch < 16) return fail("stage %d would have %d channels (< 16)", i, ch);
    const int u = c->upsample_rates[i], k = c->upsample_kernel_sizes[i];
    if (u < 2 || u % 2 || k != 2 * u) return fail("upsample stage %d: kernel %d / rate %d not supported (need kernel == 2 * rate, rate even)", i, k, u);
  }
  for (int j = 0; j < c->n_resblock_kernels; ++j) {
    const int k = c->resblock_kernel_sizes[j];
    if (k < 1 || k % 2 == 0 || k > 15) return fail("resblock kernel size %d not supported (odd, <= 15)", k);
    for (int d = 0; d < 3; ++d) {
      const int dil = c->resblock_dilation_sizes[j][d];
      if (dil < 1 || (k - 1) / 2 * dil > kVocG) return fail("resblock kernel %d dilation %d reaches past the %d guard rows", k, dil, kVocG);
    }
  }
  if (c->sampling_rate <= 0) return fail("sampling_rate must be positive");
  return 0;
}

// walks the packed layout; fills `v` when given
static size_t voc_layout(const pd_vocoder_config& c, pd_vocoder* v) {
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o = align_up(o + bytes, 1024); return r; };
  auto conv = [&](VocConv* cv, int cin, int n_total, int taps, int dil) {
    const int g = cin >= 64 ? 1 : 64 / cin;
    const int taps_pad = (taps + g - 1) / g * g;
    const int n_pad = n_total < 16 ? 16 : n_total;
    const size_t w = take((size_t)n_pad * taps_pad * cin * 2), b = take((size_t)n_pad * 4);
    if (cv) {
      cv->w_off = w; cv->b_off = b; cv->cin = cin; cv->n_total = n_pad; cv->taps = taps; cv->taps_pad = taps_pad; cv->dil = dil;
      cv->bn = n_pad < 256 ? n_pad : 256;
    }
  };
  conv(v ? &v->conv_pre : nullptr, c.num_mels, c.upsample_initial_channel, 7, 1);
  int ch = c.upsample_initial_channel, stride_f0 = 1;
  for (int i = 0; i < c.n_upsamples; ++i) stride_f0 *= c.upsample_rates[i];
  for (int i = 0; i < c.n_upsamples; ++i) {
    const int u = c.upsample_rates[i];
    conv(v ? &v->ups[i] : nullptr, ch, u * (ch / 2), 3, 1);
    {
      const size_t uw = voc_single_ok(ch, u) ? take((size_t)3 * ch * ch * 2) : 0;
      if (v) v->ups_w[i] = uw;
    }
    ch /= 2;
    stride_f0 /= u;
    const int sk = i + 1 < c.n_upsamples ? 2 * stride_f0 : 1;
    const size_t sw = take((size_t)ch * sk * 4), sb = take((size_t)ch * 4);
    if (v) {
      v->src_w[i] = sw; v->src_b[i] = sb; v->src_k[i] = sk;
      v->src_stride[i] = i + 1 < c.n_upsamples ? stride_f0 : 1;
      v->src_pad[i] = i + 1 < c.n_upsamples ? stride_f0 / 2 : 0;
    }
    for (int j = 0; j < c.n_resblock_kernels; ++j)
      for (int d = 0; d < 3; ++d) {
        conv(v ? &v->res[i][j][2 * d] : nullptr, ch, ch, c.resblock_kernel_sizes[j], c.resblock_dilation_sizes[j][d]);
        conv(v ? &v->res[i][j][2 * d + 1] : nullptr, ch, ch, c.resblock_kernel_sizes[j], 1);
        if (voc_pair_ok(ch, c.resblock_kernel_sizes[j])) {
          const size_t pw = take((size_t)2 * c.resblock_kernel_sizes[j] * ch * ch * 2);
          if (v) v->pair_w[i][j][d] = pw;
        } else if (v) {
          v->pair_w[i][j][d] = 0;
        }
      }
  }
  conv(v ? &v->conv_post : nullptr, ch, 1, 7, 1);
  {
    const size_t pw = voc_post_ok(ch, 7) ? take((size_t)7 * ch * ch * 2) : 0;
    if (v) v->post_w = pw;
  }
  const size_t lw = take(kVocHarm * 4), lb = take(4);
  if (v) { v->lin_w = lw; v->lin_b = lb; }
  return o;
}

extern "C" int pd_vocoder_state_count(const pd_vocoder_config* c) {
  if (!c) return 0;
  return 2 + c->n_upsamples * (4 + c->n_resblock_kernels * 12) + 2 + 2;
}
extern "C" size_t pd_vocoder_packed_bytes(const pd_vocoder_config* c) {
  if (voc_check_config(c)) return 0;
  return voc_layout(*c, nullptr);
}

extern "C" int pd_vocoder_create(const pd_vocoder_config* cfg, const float* const* state, int n_state, void* packed, size_t packed_bytes,
                                 void* stream, pd_vocoder** out) {
  if (!out) return fail("null out pointer");
  *out = nullptr;
  if (voc_check_config(cfg)) return 1;
  if (!state || n_state != pd_vocoder_state_count(cfg)) return fail("expected %d state tensors, got %d", pd_vocoder_state_count(cfg), n_state);
  for (int i = 0; i < n_state; ++i)
    if (!state[i]) return fail("state tensor %d is null", i);
  int device = 0, sms = 0;
  if (check_device(&device, &sms)) return 1;
  pd_vocoder* v = new pd_vocoder();
  v->cfg = *cfg;
  v->packed = static_cast<uint8_t*>(packed);
  v->device = device;
  v->sm_count = sms;
  v->packed_total = voc_layout(*cfg, v);
  if (!packed || packed_bytes < v->packed_total || reinterpret_cast<uintptr_t>(packed) % 1024) {
    const size_t need = v->packed_total;
    delete v;
    return fail("packed buffer must be 1024-byte aligned and at least %zu bytes (got %zu)", need, packed_bytes);
  }
  v->upp = 1;
  for (int i = 0; i < cfg->n_upsamples; ++i) v->upp *= cfg->upsample_rates[i];
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  uint8_t* pk = v->packed;
  const int TB = 256;
  auto blocks = [&](long long n) { return (int)std::min<long long>((n + TB - 1) / TB, 4096); };
  auto pack_conv = [&](VocConv& cv, const float* w, const float* b, int cout) -> int {
    k_voc_pack_conv<<<blocks((long long)cv.n_total * cv.taps_pad * cv.cin), TB, 0, s>>>((__half*)(pk + cv.w_off), w, cout, cv.cin, cv.taps, cv.n_total, cv.taps_pad);
    k_voc_pad_bias<<<blocks(cv.n_total), TB, 0, s>>>((float*)(pk + cv.b_off), b, cout, cv.n_total);
    PD_LAUNCH_CHECK("k_voc_pack_conv");
    if (make_w_map(&cv.tm_w, PD_FP16, pk + cv.w_off, cv.taps_pad * cv.cin, cv.n_total, cv.bn)) return 1;
    return cv.bn >= 128 ? make_w_map(&cv.tm_w_half, PD_FP16, pk + cv.w_off, cv.taps_pad * cv.cin, cv.n_total, cv.bn / 2) : 0;
  };
  int si = 0, rc = 0;
  rc |= pack_conv(v->conv_pre, state[si], state[si + 1], cfg->upsample_initial_channel);
  si += 2;
  int ch = cfg->upsample_initial_channel;
  for (int i = 0; i < cfg->n_upsamples && !rc; ++i) {
    const int u = cfg->upsample_rates[i];
    VocConv& up = v->ups[i];
    k_voc_pack_convT<<<blocks((long long)up.n_total * up.taps_pad * up.cin), TB, 0, s>>>((__half*)(pk + up.w_off), state[si], ch, ch / 2, u, up.taps_pad);
    k_voc_tile_bias<<<blocks(up.n_total), TB, 0, s>>>((float*)(pk + up.b_off), state[si + 1], ch / 2, up.n_total);
    PD_LAUNCH_CHECK("k_voc_pack_convT");
    rc |= make_w_map(&up.tm_w, PD_FP16, pk + up.w_off, up.taps_pad * up.cin, up.n_total, up.bn);
    if (!rc && up.bn >= 128) rc |= make_w_map(&up.tm_w_half, PD_FP16, pk + up.w_off, up.taps_pad * up.cin, up.n_total, up.bn / 2);
    if (voc_single_ok(ch, u)) {
      k_voc_pack_convT_planes<<<blocks(3ll * ch * ch), TB, 0, s>>>((__half*)(pk + v->ups_w[i]), state[si], ch, ch / 2, u);
      PD_LAUNCH_CHECK("k_voc_pack_convT_planes");
    }
    ch /= 2;
    k_voc_transpose_src<<<blocks((long long)ch * v->src_k[i]), TB, 0, s>>>((float*)(pk + v->src_w[i]), state[si + 2], ch, v->src_k[i]);
    PD_LAUNCH_CHECK("k_voc_transpose_src");
    PD_CUDA(cudaMemcpyAsync(pk + v->src_b[i], state[si + 3], (size_t)ch * 4, cudaMemcpyDeviceToDevice, s));
    si += 4;
    for (int j = 0; j < cfg->n_resblock_kernels && !rc; ++j) {
      // state order per ResBlock1: convs1.{0,1,2}.{weight,bias}, convs2.{0,1,2}.{weight,bias}
      for (int d = 0; d < 3 && !rc; ++d) {
        rc |= pack_conv(v->res[i][j][2 * d], state[si + 2 * d], state[si + 2 * d + 1], ch);
        rc |= pack_conv(v->res[i][j][2 * d + 1], state[si + 6 + 2 * d], state[si + 6 + 2 * d + 1], ch);
        const int kk = cfg->resblock_kernel_sizes[j];
        if (voc_pair_ok(ch, kk)) {
          k_voc_pack_pair<<<blocks(2ll * kk * ch * ch), TB, 0, s>>>((__half*)(pk + v->pair_w[i][j][d]), state[si + 2 * d], state[si + 6 + 2 * d], ch, kk);
          PD_LAUNCH_CHECK("k_voc_pack_pair");
        }
      }
      si += 12;
    }
  }
  if (!rc) rc |= pack_conv(v->conv_post, state[si], state[si + 1], 1);
  if (!rc && voc_post_ok(ch, 7)) {
    k_voc_pack_single<<<blocks(7ll * ch * ch), TB, 0, s>>>((__half*)(pk + v->post_w), state[si], ch, 1, 7);
    PD_LAUNCH_CHECK("k_voc_pack_single");
  }
  si += 2;
  if (!rc) {
    PD_CUDA(cudaMemcpyAsync(pk + v->lin_w, state[si], kVocHarm * 4, cudaMemcpyDeviceToDevice, s));
    PD_CUDA(cudaMemcpyAsync(pk + v->lin_b, state[si + 1], 4, cudaMemcpyDeviceToDevice, s));
  }
  if (rc) {
    delete v;
    return 1;
  }
  *out = v;
  return 0;
}
extern "C" void pd_vocoder_destroy(pd_vocoder* v) { delete v; }
extern "C" int pd_vocoder_upp(const pd_vocoder* v) { return v ? v->upp : 0; }

namespace {
struct VocWs {
  size_t mel, act[7], har, P, rad, total;   // act: 0/1 = stage in/out ping-pong, 2 XL, 3 TMP, 4 CUR, 5 XS, 6 SRC
};
VocWs voc_ws_layout(const pd_vocoder* v, int B, int T) {
  const pd_vocoder_config& c = v->cfg;
  VocWs w;
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o = align_up(o + bytes, 1024); return r; };
  w.mel = take((size_t)B * (T + 2 * kVocG) * c.num_mels * 2);
  // the widest guarded activation tensor of any stage (elements): conv_pre output and every stage's tensors
  size_t mx = (size_t)(T + 2 * kVocG) * c.upsample_initial_channel;
  int ch = c.upsample_initial_channel;
  long long Ti = T;
  for (int i = 0; i < c.n_upsamples; ++i) {
    ch /= 2;
    Ti *= c.upsample_rates[i];
    mx = std::max(mx, (size_t)(Ti + 2 * kVocG) * ch);
  }
  for (int i = 0; i < 7; ++i) w.act[i] = take(mx * B * 2);
  w.har = take((size_t)B * T * v->upp * 4);
  w.P = take((size_t)B * T * kVocHarm * 8);
  w.rad = take((size_t)B * T * kVocHarm * 4);
  w.total = o;
  return w;
}
}  // namespace

extern "C" size_t pd_vocoder_workspace_bytes(const pd_vocoder* v, int B, int T) {
  if (!v || B < 1 || T < 1) return 0;
  return voc_ws_layout(v, B, T).total;
}

extern "C" int pd_vocoder_forward(pd_vocoder* v, const float* c, const float* f0, const float* rand_ini, const float* noise, uint64_t seed,
                                  const int* lengths, float* wave, int B, int T, void* workspace, size_t workspace_bytes, void* stream) {
  if (!v || !c || !f0 || !rand_ini || !wave) return fail("null argument");
  if (B < 1 || T < 1) return fail("B and T must be >= 1 (got %d, %d)", B, T);
  if ((long long)B * T * v->upp >= (1ll << 31)) return fail("B * T * upp too large for one call");
  const VocWs L = voc_ws_layout(v, B, T);
  if (!workspace || workspace_bytes < L.total || reinterpret_cast<uintptr_t>(workspace) % 1024)
    return fail("workspace must be 1024-byte aligned and at least %zu bytes (got %zu)", L.total, workspace_bytes);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const pd_vocoder_config& cf = v->cfg;
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  const uint8_t* pk = v->packed;
  const int sms = v->sm_count, G = kVocG;
  auto zero_guards = [&](long long Tn, int C, __half* b0, __half* b1 = nullptr, __half* b2 = nullptr, __half* b3 = nullptr) -> int {
    VocGuardBufs gb;
    gb.p[0] = b0; gb.p[1] = b1; gb.p[2] = b2; gb.p[3] = b3;
    gb.n = b3 ? 4 : (b2 ? 3 : (b1 ? 2 : 1));
    const long long n = 2ll * G * C / 8 * B * gb.n;
    k_voc_zero_guards<<<(int)std::min<long long>((n + 255) / 256, 2048), 256, 0, s>>>(gb, (int)Tn, C, G, B);
    PD_LAUNCH_CHECK("k_voc_zero_guards");
    return 0;
  };
  // ---- harmonic source (m_source, models.py:270): phase bookkeeping per frame, then one thread per sample
  const long long Lh = (long long)T * v->upp;
  double* P = reinterpret_cast<double*>(ws + L.P);
  float* rad = reinterpret_cast<float*>(ws + L.rad);
  float* har = reinterpret_cast<float*>(ws + L.har);
  k_voc_phase<<<(B * kVocHarm + 63) / 64, 64, 0, s>>>(P, rad, f0, rand_ini, B, T, v->upp, (float)cf.sampling_rate);
  PD_LAUNCH_CHECK("k_voc_phase");
  k_voc_source<<<sms * 8, 256, 0, s>>>(har, P, rad, f0, noise, (const float*)(pk + v->lin_w), (const float*)(pk + v->lin_b), T, v->upp,
                                       (long long)B * Lh, 0.1f, 0.003f, (unsigned long long)seed);
  PD_LAUNCH_CHECK("k_voc_source");
  // ---- conv_pre (models.py:271) on the hand-off tensor c [B, M, T]; its output is stored LeakyReLU(0.1)-activated (:273)
  __half* mel = reinterpret_cast<__half*>(ws + L.mel);
  if (zero_guards(T, cf.num_mels, mel)) return 1;
  {
    const dim3 grid((T + 31) / 32, cf.num_mels / 32, B), block(32, 8);
    k_voc_mel_in<<<grid, block, 0, s>>>(mel, c, cf.num_mels, T, G, lengths);
    PD_LAUNCH_CHECK("k_voc_mel_in");
  }
  __half* act[7];
  for (int i = 0; i < 7; ++i) act[i] = reinterpret_cast<__half*>(ws + L.act[i]);
  int cur_in = 0;
  int ch = cf.upsample_initial_channel;
  long long Ti = T;
  CUtensorMap tmA;
  {
    if (zero_guards(Ti, ch, act[cur_in])) return 1;
    if (make_voc_a_map(&tmA, mel, cf.num_mels, 7, 1, T, B)) return 1;
    VocEpi ep = voc_epi();
    ep.out16 = act[cur_in] + (size_t)G * ch;
    ep.out_bstride = (Ti + 2 * G) * ch;
    ep.out_ld = ch;
    ep.bias = (const float*)(pk + v->conv_pre.b_off);
    ep.act = 1;
    ep.slope = 0.1f;
    ep.lengths = lengths;
    ep.len_mul = 1;
    if (launch_voc(v->conv_pre, tmA, mel, B, (int)Ti, ep, sms, s)) return 1;
  }
  for (int i = 0; i < cf.n_upsamples; ++i) {
    const int u = cf.upsample_rates[i];
    const int co = ch / 2;
    const long long To = Ti * u;
    const long long obs = (To + 2 * G) * co;   // utterance stride of this stage's tensors (elements)
    __half* XL = act[2]; __half* TMP = act[3]; __half* CUR = act[4]; __half* XS = act[5]; __half* SRC = act[6];
    __half* OUT = act[cur_in ^ 1];
    // noise_convs[i](har_source) (:276)
    {
      const int k = v->src_k[i], st = v->src_stride[i];
      int ppt = 8;
      auto shm_of = [&](int p) { return ((size_t)(voc_src_positions(co, p) - 1) * st + k) * sizeof(float); };
      while (ppt > 1 && shm_of(ppt) > 48 * 1024) ppt /= 2;
      const size_t shm = shm_of(ppt);
      if (shm > 48 * 1024) return fail("noise_convs[%d]: %zu bytes of shared memory (stride %d, kernel %d) not supported", i, shm, st, k);
      const int npos = voc_src_positions(co, ppt);
      const dim3 grid((unsigned)((To + npos - 1) / npos), B);
      if (st % 4 == 0 && k % 4 == 0 && k >= 32)   // long kernels are FMA-bound once four taps come per shared-memory load; short ones are store-bound
        k_voc_srcconv<true><<<grid, kSrcThreads, shm, s>>>(SRC, har, (const float*)(pk + v->src_w[i]), (const float*)(pk + v->src_b[i]), co, k, st, v->src_pad[i],
                                                           (int)To, Lh, G, ppt, lengths, v->upp);
      else
        k_voc_srcconv<false><<<grid, kSrcThreads, shm, s>>>(SRC, har, (const float*)(pk + v->src_w[i]), (const float*)(pk + v->src_b[i]), co, k, st, v->src_pad[i],
                                                            (int)To, Lh, G, ppt, lengths, v->upp);
      PD_LAUNCH_CHECK("k_voc_srcconv");
    }
    // x = ups[i](leaky_relu(x)) + x_source (:274-277), stored activated for the ResBlocks' first convs
    if (zero_guards(To, co, XL, TMP, CUR, OUT)) return 1;
    if (voc_fused_enabled() && voc_single_ok(ch, u)) {
      VocPairArgs pa;
      memset(&pa, 0, sizeof(pa));
      pa.x = act[cur_in];
      pa.add = SRC;
      pa.out = XL;
      pa.w = (const __half*)(pk + v->ups_w[i]);
      pa.b1 = pa.b2 = (const float*)(pk + v->ups[i].b_off);
      pa.bstride = (Ti + 2 * G) * ch;
      pa.out_bstride = obs;
      pa.out_goff = (long long)G * co;
      pa.B = B;
      pa.T = (int)Ti;
      pa.G = G;
      pa.k = 3;
      pa.dil = 1;
      pa.slope1 = 0.1f;
      pa.in_inv_slope = 10.f;
      pa.scale = 1.f;
      pa.act = 1;
      pa.slope = 0.1f;
      pa.lengths = lengths;
      pa.len_mul = (int)(Ti / T);
      if (launch_voc_single(ch, pa, sms, s)) return 1;
    } else {
      if (make_voc_a_map(&tmA, act[cur_in], ch, 3, 1, (int)Ti, B)) return 1;
      VocEpi ep = voc_epi();
      ep.out16 = XL + (size_t)G * co;
      ep.out_bstride = obs;
      ep.out_ld = u * co;
      ep.bias = (const float*)(pk + v->ups[i].b_off);
      ep.add16 = SRC + (size_t)G * co;
      ep.add_bstride = obs;
      ep.add_ld = u * co;
      ep.act = 1;
      ep.slope = 0.1f;
      ep.lengths = lengths;
      ep.len_mul = (int)(Ti / T);   // rows of a transposed conv are its INPUT positions
      if (launch_voc(v->ups[i], tmA, act[cur_in], B, (int)Ti, ep, sms, s)) return 1;
    }
    // xs = sum over the ResBlocks (:278-284); each ResBlock1 is three (conv dil d -> conv dil 1 -> + x) rounds (:62-69)
    const bool last_stage = i + 1 == cf.n_upsamples;
    const bool fused = voc_fused_enabled();
    for (int j = 0; j < cf.n_resblock_kernels; ++j) {
      for (int d = 0; d < 3; ++d) {
        const VocConv& c1 = v->res[i][j][2 * d];
        const VocConv& c2 = v->res[i][j][2 * d + 1];
        const bool pair = fused && voc_pair_ok(co, c1.taps);
        // unfused: x lives in XL (d = 0) then CUR, updated in place through TMP; fused: XL -> CUR -> TMP -> result (a tile reads
        // its neighbours' rows, so a round never writes the tensor it reads)
        const __half* xin = d == 0 ? XL : (pair && d == 2 ? TMP : CUR);
        __half* xout = pair && d == 1 ? TMP : CUR;
        __half* dst = nullptr;
        const __half* add = nullptr;
        float scale = 1.f, slope = 0.1f;
        int act = 1;
        if (d < 2) {                                  // x = xt + x, kept activated for the next round's first conv
          dst = xout;
        } else if (j + 1 < cf.n_resblock_kernels) {   // this ResBlock's result joins the running sum (raw)
          dst = XS;
          if (j > 0) add = XS;
          act = 0;
        } else {                                      // last ResBlock: x = xs / num_kernels (:285), then the next consumer's LeakyReLU
          dst = OUT;
          if (j > 0) add = XS;
          scale = 1.f / (float)cf.n_resblock_kernels;
          slope = last_stage ? 0.01f : 0.1f;          // F.leaky_relu(x) before conv_post uses the default slope (:286)
        }
        if (pair) {
          VocPairArgs pa;
          memset(&pa, 0, sizeof(pa));
          pa.x = xin;
          pa.add = add;
          pa.out = dst;
          pa.w = (const __half*)(pk + v->pair_w[i][j][d]);
          pa.b1 = (const float*)(pk + c1.b_off);
          pa.b2 = (const float*)(pk + c2.b_off);
          pa.bstride = obs;
          pa.out_bstride = obs;
          pa.out_goff = (long long)G * co;
          pa.B = B;
          pa.T = (int)To;
          pa.G = G;
          pa.k = c1.taps;
          pa.dil = c1.dil;
          pa.slope1 = 0.1f;
          pa.in_inv_slope = 10.f;
          pa.scale = scale;
          pa.act = act;
          pa.slope = slope;
          pa.lengths = lengths;
          pa.len_mul = (int)(To / T);
          if (launch_voc_pair(co, pa, sms, s)) return 1;
          continue;
        }
        {
          if (make_voc_a_map(&tmA, xin, co, c1.taps, c1.dil, (int)To, B)) return 1;
          VocEpi ep = voc_epi();
          ep.out16 = TMP + (size_t)G * co;
          ep.out_bstride = obs;
          ep.out_ld = co;
          ep.bias = (const float*)(pk + c1.b_off);
          ep.act = 1;
          ep.slope = 0.1f;
          ep.lengths = lengths;
          ep.len_mul = (int)(To / T);
          if (launch_voc(c1, tmA, xin, B, (int)To, ep, sms, s)) return 1;
        }
        {
          if (make_voc_a_map(&tmA, TMP, co, c2.taps, 1, (int)To, B)) return 1;
          VocEpi ep = voc_epi();
          ep.bias = (const float*)(pk + c2.b_off);
          ep.res16 = xin + (size_t)G * co;
          ep.res_bstride = obs;
          ep.res_ld = co;
          ep.out_bstride = obs;
          ep.out_ld = co;
          ep.out16 = dst + (size_t)G * co;
          if (add) { ep.add16 = add + (size_t)G * co; ep.add_bstride = obs; ep.add_ld = co; }
          ep.scale = scale;
          ep.act = act;
          ep.slope = slope;
          ep.lengths = lengths;
          ep.len_mul = (int)(To / T);
          if (launch_voc(c2, tmA, TMP, B, (int)To, ep, sms, s)) return 1;
        }
      }
    }
    cur_in ^= 1;
    ch = co;
    Ti = To;
  }
  // ---- conv_post + tanh (:287-288) -> waveform [B, 1, T * upp] fp32
  if (voc_fused_enabled() && voc_post_ok(ch, v->conv_post.taps)) {
    VocPairArgs pa;
    memset(&pa, 0, sizeof(pa));
    pa.x = act[cur_in];
    pa.out32 = wave;
    pa.w = (const __half*)(pk + v->post_w);
    pa.b1 = pa.b2 = (const float*)(pk + v->conv_post.b_off);   // padded to 16 entries
    pa.bstride = (Ti + 2 * G) * ch;
    pa.out_bstride = pa.bstride;
    pa.out_goff = (long long)G * ch;
    pa.B = B;
    pa.T = (int)Ti;
    pa.G = G;
    pa.k = v->conv_post.taps;
    pa.dil = 1;
    pa.slope1 = 0.1f;
    pa.in_inv_slope = 10.f;
    pa.scale = 1.f;
    pa.act = 0;
    pa.lengths = lengths;
    pa.len_mul = v->upp;
    if (launch_voc_single(ch, pa, sms, s)) return 1;
  } else {
    if (make_voc_a_map(&tmA, act[cur_in], ch, 7, 1, (int)Ti, B)) return 1;
    VocEpi ep = voc_epi();
    ep.out32 = wave;
    ep.bias = (const float*)(pk + v->conv_post.b_off);
    ep.act = 2;
    ep.lengths = lengths;
    ep.len_mul = v->upp;
    if (launch_voc(v->conv_post, tmA, act[cur_in], B, (int)Ti, ep, sms, s)) return 1;
  }
  return 0;
}

// Test hook: one convolution through the generator's kernel.  x [B, Cin, T] fp32, w [Cout, Cin, k] fp32, bias [Cout] ->
// out [B, Cout, T] fp32 = conv1d(x, w, bias, dilation = dil, padding = (k - 1) / 2 * dil) with fp16 operands.  scratch:
// device bytes >= pd_voc_debug_conv_scratch(B, T, Cin, Cout, k), 1024-byte aligned.
extern "C" size_t pd_voc_debug_conv_scratch(int B, int T, int Cin, int Cout, int k) {
  const int n_pad = Cout < 16 ? 16 : Cout, g = Cin >= 64 ? 1 : 64 / Cin, tp = (k + g - 1) / g * g;
  return align_up((size_t)B * (T + 2 * kVocG) * Cin * 2, 1024) + align_up((size_t)B * (T + 2 * kVocG) * n_pad * 2, 1024) +
         align_up((size_t)n_pad * tp * Cin * 2, 1024) + align_up((size_t)n_pad * 4, 1024);
}
__global__ void k_voc_dbg_in(__half* __restrict__ dst, const float* __restrict__ x, int C, int Tn, int G, long long total) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const long long bt = i / C;
    const int t = (int)(bt % Tn), b = (int)(bt / Tn);
    dst[((long long)b * (Tn + 2 * G) + G + t) * C + c] = __float2half_rn(x[((long long)b * C + c) * Tn + t]);
  }
}
__global__ void k_voc_dbg_out(float* __restrict__ out, const __half* __restrict__ src, int C, int ld, int Tn, int G, long long total) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int t = (int)(i % Tn);
    const long long bc = i / Tn;
    const int c = (int)(bc % C), b = (int)(bc / C);
    out[i] = __half2float(src[((long long)b * (Tn + 2 * G) + G + t) * ld + c]);
  }
}
extern "C" int pd_voc_debug_conv(const float* x, int B, int T, int Cin, const float* w, const float* bias, int Cout, int k, int dil, float* out,
                                 void* scratch, size_t scratch_bytes, void* stream) {
  if (!x || !w || !bias || !out || !scratch) return fail("null argument");
  if (!voc_pow2_ch(Cin) || (Cout != 1 && !voc_pow2_ch(Cout)) || k % 2 == 0 || (k - 1) / 2 * dil > kVocG) return fail("unsupported debug conv shape");
  if (scratch_bytes < pd_voc_debug_conv_scratch(B, T, Cin, Cout, k) || reinterpret_cast<uintptr_t>(scratch) % 1024) return fail("scratch too small or misaligned");
  int device = 0, sms = 0;
  if (check_device(&device, &sms)) return 1;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  VocConv cv;
  const int n_pad = Cout < 16 ? 16 : Cout, g = Cin >= 64 ? 1 : 64 / Cin;
  cv.cin = Cin; cv.n_total = n_pad; cv.taps = k; cv.taps_pad = (k + g - 1) / g * g; cv.dil = dil; cv.bn = n_pad < 256 ? n_pad : 256;
  uint8_t* p = static_cast<uint8_t*>(scratch);
  __half* xin = reinterpret_cast<__half*>(p);
  p += align_up((size_t)B * (T + 2 * kVocG) * Cin * 2, 1024);
  __half* o16 = reinterpret_cast<__half*>(p);
  p += align_up((size_t)B * (T + 2 * kVocG) * n_pad * 2, 1024);
  __half* w16 = reinterpret_cast<__half*>(p);
  p += align_up((size_t)n_pad * cv.taps_pad * Cin * 2, 1024);
  float* b32 = reinterpret_cast<float*>(p);
  PD_CUDA(cudaMemsetAsync(xin, 0, (size_t)B * (T + 2 * kVocG) * Cin * 2, s));
  const long long nin = (long long)B * T * Cin, nout = (long long)B * T * Cout;
  k_voc_dbg_in<<<(int)std::min<long long>((nin + 255) / 256, 4096), 256, 0, s>>>(xin, x, Cin, T, kVocG, nin);
  k_voc_pack_conv<<<(int)std::min<long long>(((long long)n_pad * cv.taps_pad * Cin + 255) / 256, 4096), 256, 0, s>>>(w16, w, Cout, Cin, k, n_pad, cv.taps_pad);
  k_voc_pad_bias<<<1, 256, 0, s>>>(b32, bias, Cout, n_pad);
  PD_LAUNCH_CHECK("k_voc_dbg_in");
  CUtensorMap tmA;
  if (make_voc_a_map(&tmA, xin, Cin, k, dil, T, B) || make_w_map(&cv.tm_w, PD_FP16, w16, cv.taps_pad * Cin, n_pad, cv.bn)) return 1;
  if (cv.bn >= 128 && make_w_map(&cv.tm_w_half, PD_FP16, w16, cv.taps_pad * Cin, n_pad, cv.bn / 2)) return 1;
  VocEpi ep = voc_epi();
  ep.out16 = o16 + (size_t)kVocG * n_pad;
  ep.out_bstride = (long long)(T + 2 * kVocG) * n_pad;
  ep.out_ld = n_pad;
  ep.bias = b32;
  if (launch_voc(cv, tmA, xin, B, T, ep, sms, s)) return 1;
  k_voc_dbg_out<<<(int)std::min<long long>((nout + 255) / 256, 4096), 256, 0, s>>>(out, o16, Cout, n_pad, T, kVocG, nout);
  PD_LAUNCH_CHECK("k_voc_dbg_out");
  return 0;
}

// Test hook: one fused ResBlock round (pd_voc_pair.cuh).  x [B, C, T] fp32 (the raw x), w1 / w2 [C, C, k], b1 / b2 [C] ->
// out [B, C, T] fp32 = c2(leaky_relu(c1(leaky_relu(x, 0.1)), 0.1)) + x with c1 dilated by `dil`, fp16 operands and storage.
extern "C" size_t pd_voc_debug_pair_scratch(int B, int T, int C, int k) {
  return 2 * align_up((size_t)B * (T + 2 * kVocG) * C * 2, 1024) + align_up((size_t)2 * k * C * C * 2, 1024) + 2 * align_up((size_t)C * 4, 1024);
}
__global__ void k_voc_dbg_in_act(__half* __restrict__ dst, const float* __restrict__ x, int C, int Tn, int G, long long total, float slope) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const long long bt = i / C;
    const int t = (int)(bt % Tn), b = (int)(bt / Tn);
    const float v = x[((long long)b * C + c) * Tn + t];
    dst[((long long)b * (Tn + 2 * G) + G + t) * C + c] = __float2half_rn(v >= 0.f ? v : v * slope);
  }
}
extern "C" int pd_voc_debug_pair(const float* x, int B, int T, int C, const float* w1, const float* b1, const float* w2, const float* b2, int k,
                                 int dil, float* out, void* scratch, size_t scratch_bytes, void* stream) {
  if (!x || !w1 || !b1 || !w2 || !b2 || !out || !scratch) return fail("null argument");
  if (!voc_pair_ok(C, k) || dil < 1 || (k - 1) / 2 * (dil + 1) > kVocG) return fail("unsupported debug pair shape");
  if (scratch_bytes < pd_voc_debug_pair_scratch(B, T, C, k) || reinterpret_cast<uintptr_t>(scratch) % 1024) return fail("scratch too small or misaligned");
  int device = 0, sms = 0;
  if (check_device(&device, &sms)) return 1;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  uint8_t* p = static_cast<uint8_t*>(scratch);
  const size_t act_bytes = align_up((size_t)B * (T + 2 * kVocG) * C * 2, 1024);
  __half* xin = reinterpret_cast<__half*>(p);
  p += act_bytes;
  __half* o16 = reinterpret_cast<__half*>(p);
  p += act_bytes;
  __half* w16 = reinterpret_cast<__half*>(p);
  p += align_up((size_t)2 * k * C * C * 2, 1024);
  float* b1d = reinterpret_cast<float*>(p);
  p += align_up((size_t)C * 4, 1024);
  float* b2d = reinterpret_cast<float*>(p);
  PD_CUDA(cudaMemsetAsync(xin, 0, act_bytes, s));
  PD_CUDA(cudaMemsetAsync(o16, 0, act_bytes, s));
  const long long n = (long long)B * T * C;
  k_voc_dbg_in_act<<<(int)std::min<long long>((n + 255) / 256, 4096), 256, 0, s>>>(xin, x, C, T, kVocG, n, 0.1f);
  k_voc_pack_pair<<<(int)std::min<long long>((2ll * k * C * C + 255) / 256, 4096), 256, 0, s>>>(w16, w1, w2, C, k);
  PD_CUDA(cudaMemcpyAsync(b1d, b1, (size_t)C * 4, cudaMemcpyDeviceToDevice, s));
  PD_CUDA(cudaMemcpyAsync(b2d, b2, (size_t)C * 4, cudaMemcpyDeviceToDevice, s));
  PD_LAUNCH_CHECK("k_voc_dbg_in_act");
  VocPairArgs pa;
  memset(&pa, 0, sizeof(pa));
  pa.x = xin;
  pa.out = o16;
  pa.w = w16;
  pa.b1 = b1d;
  pa.b2 = b2d;
  pa.bstride = (long long)(T + 2 * kVocG) * C;
  pa.out_bstride = pa.bstride;
  pa.out_goff = (long long)kVocG * C;
  pa.B = B;
  pa.T = T;
  pa.G = kVocG;
  pa.k = k;
  pa.dil = dil;
  pa.slope1 = 0.1f;
  pa.in_inv_slope = 10.f;
  pa.scale = 1.f;
  const int rc = launch_voc_pair(C, pa, sms, s);
  if (rc) return 1;
  k_voc_dbg_out<<<(int)std::min<long long>((n + 255) / 256, 4096), 256, 0, s>>>(out, o16, C, C, T, kVocG, n);
  PD_LAUNCH_CHECK("k_voc_dbg_out");
  return 0;
}
