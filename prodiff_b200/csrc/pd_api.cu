// C-ABI host side of prodiff_b200 (see include/prodiff_b200.h): weight packing, scratch layout, TMA tensor maps,
// and the launch sequences behind WaveNet.forward and the K-step GaussianDiffusion reverse loop.
#include "../../include/prodiff_b200.h"
#include "pd_aux.cuh"
#include "pd_layer.cuh"
#include "pd_tail.cuh"
#include "pd_sumtail.cuh"
#include "pd_voc.cuh"
#include "pd_voc_pair.cuh"

#include <algorithm>
#include <math.h>
#include <stdarg.h>
#include <string.h>
#include <string>
#include <vector>

using namespace pd;

// ------------------------------------------------------------------ errors / bookkeeping
static thread_local char g_err[512] = "";
static thread_local long long g_launches = 0;

static int fail(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return 1;
}
#define PD_CUDA(expr)                                                                    \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess) return fail("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)
#define PD_LAUNCH_CHECK(name)                                                            \
  do {                                                                                   \
    ++g_launches;                                                                        \
    cudaError_t _e = cudaGetLastError();                                                 \
    if (_e != cudaSuccess) return fail("launch of %s failed: %s", name, cudaGetErrorString(_e)); \
  } while (0)

extern "C" const char* pd_last_error(void) { return g_err; }
extern "C" int pd_abi_version(void) { return PD_ABI_VERSION; }
extern "C" long long pd_launch_count(int reset) {
  const long long v = g_launches;
  if (reset) g_launches = 0;
  return v;
}

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// ------------------------------------------------------------------ optional per-kernel event profiling
// When enabled (pd_profile_begin), every GEMM launch is bracketed by a CUDA-event pair on its own stream; bench.py
// uses this in a separate, untimed-for-throughput pass to get the dominant kernel's average duration.
struct ProfRec {
  int kind;
  cudaEvent_t e0, e1;
};
static thread_local bool g_prof_on = false;
static thread_local std::vector<ProfRec>* g_prof = nullptr;
static thread_local size_t g_prof_used = 0;

static void prof_before(int kind, cudaStream_t s) {
  if (!g_prof_on) return;
  if (!g_prof) g_prof = new std::vector<ProfRec>();
  if (g_prof_used == g_prof->size()) {
    ProfRec r;
    r.kind = kind;
    cudaEventCreate(&r.e0);
    cudaEventCreate(&r.e1);
    g_prof->push_back(r);
  }
  (*g_prof)[g_prof_used].kind = kind;
  cudaEventRecord((*g_prof)[g_prof_used].e0, s);
}
static void prof_after(cudaStream_t s) {
  if (!g_prof_on) return;
  cudaEventRecord((*g_prof)[g_prof_used].e1, s);
  ++g_prof_used;
}
// debug: device buffer of PD_N_KINDS x 1024 slots; CTA 0 of every GEMM launch records clock64() at pipeline events
static thread_local long long* g_trace = nullptr;
extern "C" int pd_debug_set_trace(long long* device_buf) {
  g_trace = device_buf;
  return 0;
}
extern "C" int pd_profile_begin(void) {
  g_prof_on = true;
  g_prof_used = 0;
  return 0;
}
extern "C" int pd_profile_end(float* ms_by_kind, int* count_by_kind, int n_kinds) {
  g_prof_on = false;
  for (int i = 0; i < n_kinds; ++i) {
    ms_by_kind[i] = 0.f;
    count_by_kind[i] = 0;
  }
  for (size_t i = 0; i < g_prof_used; ++i) {
    ProfRec& r = (*g_prof)[i];
    PD_CUDA(cudaEventSynchronize(r.e1));
    float ms = 0.f;
    PD_CUDA(cudaEventElapsedTime(&ms, r.e0, r.e1));
    if (r.kind >= 0 && r.kind < n_kinds) {
      ms_by_kind[r.kind] += ms;
      count_by_kind[r.kind] += 1;
    }
  }
  g_prof_used = 0;
  return 0;
}

// ------------------------------------------------------------------ driver entry point for tensor maps
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
      q != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}

// 16-bit tensor, up to 4 dims (innermost first), box = (64, box1, 1, 1), 128-byte swizzle, zero OOB fill.
static int make_map(CUtensorMap* m, int dtype, const void* base, int rank, const uint64_t* dims, uint32_t box1) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return fail("cuTensorMapEncodeTiled entry point unavailable (driver too old?)");
  cuuint64_t gdim[4] = {1, 1, 1, 1}, gstr[3] = {0, 0, 0};
  cuuint32_t box[4] = {64, box1, 1, 1}, estr[4] = {1, 1, 1, 1};
  uint64_t stride = 2;
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    stride *= dims[i];
    if (i < rank - 1) gstr[i] = stride;
  }
  const CUresult r = enc(m, dtype == PD_FP16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16,
                         (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled failed with CUresult %d (rank %d dims %llu %llu)", (int)r, rank,
                                     (unsigned long long)dims[0], (unsigned long long)dims[1]);
  return 0;
}
// activation map: [slabs][B][T][Cin] 16-bit
static int make_act_map(CUtensorMap* m, int dtype, const void* base, int Cin, int T, int B, int slabs, int box_rows = kBlockM) {
  const uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)T, (uint64_t)B, (uint64_t)slabs};
  return make_map(m, dtype, base, 4, dims, (uint32_t)box_rows);
}
// weight map: [rows][K] 16-bit
static int make_w_map(CUtensorMap* m, int dtype, const void* base, int K, int rows, int bn) {
  const uint64_t dims[2] = {(uint64_t)K, (uint64_t)rows};
  return make_map(m, dtype, base, 2, dims, (uint32_t)bn);
}

// ------------------------------------------------------------------ packed weight layout
struct PackedLayout {
  // 16-bit operands (byte offsets)
  size_t w_in, w_cond, w_dil, w_res, w_skip, w_sp, w_out;
  // fp32 biases
  size_t b_in, b_cp, b_res, b_skipsum, b_sp, b_out;
  // fp32 copies for the step-embedding chain
  size_t f_mlp0w, f_mlp0b, f_mlp2w, f_mlp2b, f_dpw, f_dpb, f_dilw;
  size_t total;
};
static PackedLayout packed_layout(const pd_config& c) {
  const size_t M = c.in_dims, H = c.hidden_size, L = c.residual_layers, C = c.residual_channels;
  PackedLayout p;
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o = align_up(o + bytes, 1024); return r; };
  p.w_in = take(C * M * 2);
  p.w_cond = take(L * 2 * C * H * 2);
  p.w_dil = take(L * 2 * C * 3 * C * 2);
  p.w_res = take((L + 1) * C * C * 2);  // + one [C x C] identity block (fused layer kernel: residual add on the tensor core)
  p.w_skip = take(C * L * C * 2);
  p.w_sp = take(C * C * 2);
  p.w_out = take(128 * C * 2);
  p.b_in = take(C * 4);
  p.b_cp = take(L * 2 * C * 4);
  p.b_res = take(L * C * 4);
  p.b_skipsum = take(C * 4);
  p.b_sp = take(C * 4);
  p.b_out = take(128 * 4);
  p.f_mlp0w = take(4 * C * C * 4);
  p.f_mlp0b = take(4 * C * 4);
  p.f_mlp2w = take(4 * C * C * 4);
  p.f_mlp2b = take(C * 4);
  p.f_dpw = take(L * C * C * 4);
  p.f_dpb = take(L * C * 4);
  p.f_dilw = take(L * 2 * C * C * 3 * 4);
  p.total = o;
  return p;
}

static int check_config(const pd_config* c) {
  if (!c) return fail("null config");
  if (c->in_dims != 64 && c->in_dims != 128) return fail("in_dims must be 64 or 128 (got %d)", c->in_dims);
  if (c->hidden_size <= 0 || c->hidden_size % 64) return fail("hidden_size must be a multiple of 64 (got %d)", c->hidden_size);
  if (c->residual_channels != 256 && c->residual_channels != 512)
    return fail("residual_channels must be 256 or 512 (got %d)", c->residual_channels);
  if (c->residual_layers < 1) return fail("residual_layers must be >= 1");
  if (c->dilation_cycle_length < 1 || c->dilation_cycle_length > 16) return fail("dilation_cycle_length out of range");
  if (c->dtype != PD_BF16 && c->dtype != PD_FP16) return fail("dtype must be PD_BF16 or PD_FP16");
  return 0;
}

extern "C" int pd_state_count(const pd_config* cfg) { return cfg ? 10 + 8 * cfg->residual_layers : 0; }
extern "C" size_t pd_packed_bytes(const pd_config* cfg) {
  if (check_config(cfg)) return 0;
  return packed_layout(*cfg).total;
}

struct pd_model {
  pd_config cfg;
  PackedLayout lay;
  uint8_t* packed;
  int device, sm_count;
  CUtensorMap tm_w_in, tm_w_cond, tm_w_dil, tm_w_res, tm_w_skip, tm_w_sp, tm_w_out;
  CUtensorMap tm_w_dil128, tm_w_res128;  // 128-row boxes: one CTA's half of a cta_group::2 weight block
  CUtensorMap tm_w_in_h, tm_w_cond_h, tm_w_skip_h, tm_w_sp_h, tm_w_out_h;  // half-height boxes for the cta_group::2 GEMMs
};

static int check_device(int* device, int* sms) {
  PD_CUDA(cudaGetDevice(device));
  int major = 0;
  PD_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, *device));
  if (major != 10) return fail("prodiff_b200 needs a compute-capability 10.x device (B200); found %d.x -- no fallback path", major);
  PD_CUDA(cudaDeviceGetAttribute(sms, cudaDevAttrMultiProcessorCount, *device));
  return 0;
}

template <typename T>
static int pack_all(const pd_config& c, const PackedLayout& lay, uint8_t* pk, const float* const* st, cudaStream_t s) {
  const int M = c.in_dims, H = c.hidden_size, L = c.residual_layers, C = c.residual_channels;
  const int TB = 256;
  auto blocks = [&](long long n) { return (int)std::min<long long>((n + TB - 1) / TB, 4096); };
  // 16-bit operands
  k_pack_rows<T><<<blocks((long long)C * M), TB, 0, s>>>((T*)(pk + lay.w_in), M, 0, st[0], C, 0, C, M, 0);
  PD_LAUNCH_CHECK("k_pack_rows(in)");
  k_pack_rows<T><<<blocks((long long)C * C), TB, 0, s>>>((T*)(pk + lay.w_sp), C, 0, st[6], C, 0, C, C, 0);
  PD_LAUNCH_CHECK("k_pack_rows(sp)");
  k_pack_rows<T><<<blocks(128ll * C), TB, 0, s>>>((T*)(pk + lay.w_out), C, 0, st[8], M, 0, 128, C, 0);
  PD_LAUNCH_CHECK("k_pack_rows(out)");
  k_identity<T><<<blocks((long long)C * C), TB, 0, s>>>((T*)(pk + lay.w_res) + (size_t)L * C * C, C);
  PD_LAUNCH_CHECK("k_identity");
  k_pack_bias<<<blocks(C), TB, 0, s>>>((float*)(pk + lay.b_in), st[1], nullptr, 0, C, C, 0, 0);
  k_pack_bias<<<blocks(C), TB, 0, s>>>((float*)(pk + lay.b_sp), st[7], nullptr, 0, C, C, 0, 0);
  k_pack_bias<<<1, 128, 0, s>>>((float*)(pk + lay.b_out), st[9], nullptr, 0, M, 128, 0, 0);
  PD_LAUNCH_CHECK("k_pack_bias");
  for (int l = 0; l < L; ++l) {
    const float* const* q = st + 10 + 8 * l;
    k_pack_dil<T><<<blocks(6ll * C * C), TB, 0, s>>>((T*)(pk + lay.w_dil) + (size_t)l * 2 * C * 3 * C, q[0], C);
    PD_LAUNCH_CHECK("k_pack_dil");
    k_pack_rows<T><<<blocks(2ll * C * H), TB, 0, s>>>((T*)(pk + lay.w_cond) + (size_t)l * 2 * C * H, H, 0, q[4], 2 * C, 0, 2 * C, H, C);
    PD_LAUNCH_CHECK("k_pack_rows(cond)");
    k_pack_rows<T><<<blocks((long long)C * C), TB, 0, s>>>((T*)(pk + lay.w_res) + (size_t)l * C * C, C, 0, q[6], 2 * C, 0, C, C, 0);
    PD_LAUNCH_CHECK("k_pack_rows(res)");
    // skip half of output_projection, K-concatenated over layers: Wskip[r][l*C + c] = W_l[C + r][c]
    k_pack_rows<T><<<blocks((long long)C * C), TB, 0, s>>>((T*)(pk + lay.w_skip), (long long)L * C, l * C, q[6], 2 * C, C, C, C, 0);
    PD_LAUNCH_CHECK("k_pack_rows(skip)");
    // hoisted pre-activation bias: conditioner bias + dilated-conv bias, gate-paired order
    k_pack_bias<<<blocks(2 * C), TB, 0, s>>>((float*)(pk + lay.b_cp) + (size_t)l * 2 * C, q[5], q[1], 0, 2 * C, 2 * C, C, 0);
    k_pack_bias<<<blocks(C), TB, 0, s>>>((float*)(pk + lay.b_res) + (size_t)l * C, q[7], nullptr, 0, 2 * C, C, 0, 0);
    k_pack_bias<<<blocks(C), TB, 0, s>>>((float*)(pk + lay.b_skipsum), q[7], nullptr, C, 2 * C, C, 0, l > 0);
    PD_LAUNCH_CHECK("k_pack_bias(layer)");
    PD_CUDA(cudaMemcpyAsync(pk + lay.f_dpw + (size_t)l * C * C * 4, q[2], (size_t)C * C * 4, cudaMemcpyDeviceToDevice, s));
    PD_CUDA(cudaMemcpyAsync(pk + lay.f_dpb + (size_t)l * C * 4, q[3], (size_t)C * 4, cudaMemcpyDeviceToDevice, s));
    PD_CUDA(cudaMemcpyAsync(pk + lay.f_dilw + (size_t)l * 2 * C * C * 3 * 4, q[0], (size_t)2 * C * C * 3 * 4, cudaMemcpyDeviceToDevice, s));
  }
  PD_CUDA(cudaMemcpyAsync(pk + lay.f_mlp0w, st[2], (size_t)4 * C * C * 4, cudaMemcpyDeviceToDevice, s));
  PD_CUDA(cudaMemcpyAsync(pk + lay.f_mlp0b, st[3], (size_t)4 * C * 4, cudaMemcpyDeviceToDevice, s));
  PD_CUDA(cudaMemcpyAsync(pk + lay.f_mlp2w, st[4], (size_t)4 * C * C * 4, cudaMemcpyDeviceToDevice, s));
  PD_CUDA(cudaMemcpyAsync(pk + lay.f_mlp2b, st[5], (size_t)C * 4, cudaMemcpyDeviceToDevice, s));
  return 0;
}

extern "C" int pd_model_create(const pd_config* cfg, const float* const* state, int n_state, void* packed,
                               size_t packed_bytes, void* stream, pd_model** out) {
  if (!out) return fail("null out pointer");
  *out = nullptr;
  if (check_config(cfg)) return 1;
  if (!state || n_state != pd_state_count(cfg)) return fail("expected %d state tensors, got %d", pd_state_count(cfg), n_state);
  for (int i = 0; i < n_state; ++i)
    if (!state[i]) return fail("state tensor %d is null", i);
  const PackedLayout lay = packed_layout(*cfg);
  if (!packed || packed_bytes < lay.total) return fail("packed buffer too small: need %zu bytes, got %zu", lay.total, packed_bytes);
  if (reinterpret_cast<uintptr_t>(packed) % 1024) return fail("packed buffer must be 1024-byte aligned");
  int device = 0, sms = 0;
  if (check_device(&device, &sms)) return 1;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  uint8_t* pk = static_cast<uint8_t*>(packed);
  const int rc = cfg->dtype == PD_FP16 ? pack_all<__half>(*cfg, lay, pk, state, s) : pack_all<__nv_bfloat16>(*cfg, lay, pk, state, s);
  if (rc) return rc;

  pd_model* m = new pd_model();
  m->cfg = *cfg;
  m->lay = lay;
  m->packed = pk;
  m->device = device;
  m->sm_count = sms;
  const int M = cfg->in_dims, H = cfg->hidden_size, L = cfg->residual_layers, C = cfg->residual_channels, dt = cfg->dtype;
  if (make_w_map(&m->tm_w_in, dt, pk + lay.w_in, M, C, 256) || make_w_map(&m->tm_w_cond, dt, pk + lay.w_cond, H, L * 2 * C, 256) ||
      make_w_map(&m->tm_w_dil, dt, pk + lay.w_dil, 3 * C, L * 2 * C, 256) || make_w_map(&m->tm_w_res, dt, pk + lay.w_res, C, L * C, 256) ||
      make_w_map(&m->tm_w_skip, dt, pk + lay.w_skip, L * C, C, 256) || make_w_map(&m->tm_w_sp, dt, pk + lay.w_sp, C, C, 256) ||
      make_w_map(&m->tm_w_out, dt, pk + lay.w_out, C, 128, 128) ||
      make_w_map(&m->tm_w_dil128, dt, pk + lay.w_dil, 3 * C, L * 2 * C, 128) || make_w_map(&m->tm_w_res128, dt, pk + lay.w_res, C, (L + 1) * C, 128) ||
      make_w_map(&m->tm_w_in_h, dt, pk + lay.w_in, M, C, 128) || make_w_map(&m->tm_w_cond_h, dt, pk + lay.w_cond, H, L * 2 * C, 128) ||
      make_w_map(&m->tm_w_skip_h, dt, pk + lay.w_skip, L * C, C, 128) || make_w_map(&m->tm_w_sp_h, dt, pk + lay.w_sp, C, C, 128) ||
      make_w_map(&m->tm_w_out_h, dt, pk + lay.w_out, C, 128, 64)) {
    delete m;
    return 1;
  }
  *out = m;
  return 0;
}
extern "C" void pd_model_destroy(pd_model* m) { delete m; }

// ------------------------------------------------------------------ plan (scratch layout + activation maps)
struct Buf {
  const char* name;
  size_t off, bytes;
};
struct pd_plan {
  const pd_model* m;
  int B, T, E;
  uint8_t* ws;
  std::vector<Buf> bufs;
  CUtensorMap tm_cond, tm_xin, tm_xa, tm_xb, tm_y, tm_s1, tm_s2;
  CUtensorMap tm_xa_halo[5], tm_xb_halo[5];   // box (64, 128 + 2*2^i): haloed tile loads of the fused layer kernel
  CUtensorMap tm_xa_st, tm_xb_st, tm_y_st;    // box (64, 32): per-warp TMA stores
  bool fused_ok;
  bool tail_ok;
  bool ragged;   // pd_plan_set_lengths installed per-utterance lengths
  float out_scale;   // pd_plan_set_output_scale: multiplier of the finished sample (vocoder hand-off), 1 = off
  size_t off(const char* n) const {
    for (const Buf& b : bufs)
      if (!strcmp(b.name, n)) return b.off;
    return (size_t)-1;
  }
  template <typename U> U* ptr(const char* n) const { return reinterpret_cast<U*>(ws + off(n)); }
};

static size_t plan_layout(const pd_config& c, int B, int T, int E, std::vector<Buf>* bufs) {
  const size_t N = (size_t)B * T, M = c.in_dims, H = c.hidden_size, L = c.residual_layers, C = c.residual_channels;
  size_t o = 0;
  auto take = [&](const char* name, size_t bytes) {
    if (bufs) bufs->push_back(Buf{name, o, bytes});
    o = align_up(o + bytes, 1024);
  };
  take("cond16", N * H * 2);
  const size_t mtiles = (size_t)B * ((T + kBlockM - 1) / kBlockM);
  take("cp", L * mtiles * kBlockM * 2 * C * 2);  // tile-blocked layout, tiles padded to 128 frames
  take("xin", N * M * 2);
  take("xa", N * C * 2);
  take("xb", N * C * 2);
  take("ybuf", L * N * C * 2);
  take("s1", N * C * 2);
  take("s2", N * C * 2);
  take("steps", (size_t)E * 4);
  take("emb0", (size_t)E * C * 4);
  take("hid", (size_t)E * 4 * C * 4);
  take("emb", (size_t)E * C * 4);
  take("dvec", (size_t)E * L * C * 4);
  take("sb", (size_t)E * L * 4 * 2 * C * 4);
  take("flags", L * mtiles * 4);  // multi-layer kernel: per (layer, tile) completion counters
  take("lens", (size_t)B * 4);      // optional per-utterance valid lengths (pd_plan_set_lengths)
  return o;
}

extern "C" size_t pd_workspace_bytes(const pd_model* m, int B, int T, int max_entries) {
  if (!m || B < 1 || T < 1 || max_entries < 1) return 0;
  return plan_layout(m->cfg, B, T, max_entries, nullptr);
}

extern "C" int pd_plan_create(const pd_model* m, int B, int T, int max_entries, void* workspace, size_t workspace_bytes,
                              pd_plan** out) {
  if (!out) return fail("null out pointer");
  *out = nullptr;
  if (!m) return fail("null model");
  if (B < 1 || T < 1 || max_entries < 1) return fail("B, T and max_entries must be >= 1 (got %d, %d, %d)", B, T, max_entries);
  if ((long long)B * T >= (1ll << 31)) return fail("B*T too large");
  pd_plan* p = new pd_plan();
  p->m = m;
  p->B = B;
  p->T = T;
  p->E = max_entries;
  p->ws = static_cast<uint8_t*>(workspace);
  const size_t need = plan_layout(m->cfg, B, T, max_entries, &p->bufs);
  if (!workspace || workspace_bytes < need) {
    delete p;
    return fail("workspace too small: need %zu bytes, got %zu", need, workspace_bytes);
  }
  if (reinterpret_cast<uintptr_t>(workspace) % 1024) {
    delete p;
    return fail("workspace must be 1024-byte aligned");
  }
  const pd_config& c = m->cfg;
  const int dt = c.dtype, C = c.residual_channels;
  if (make_act_map(&p->tm_cond, dt, p->ptr<void>("cond16"), c.hidden_size, T, B, 1) ||
      make_act_map(&p->tm_xin, dt, p->ptr<void>("xin"), c.in_dims, T, B, 1) ||
      make_act_map(&p->tm_xa, dt, p->ptr<void>("xa"), C, T, B, 1) || make_act_map(&p->tm_xb, dt, p->ptr<void>("xb"), C, T, B, 1) ||
      make_act_map(&p->tm_y, dt, p->ptr<void>("ybuf"), C, T, B, c.residual_layers) ||
      make_act_map(&p->tm_s1, dt, p->ptr<void>("s1"), C, T, B, 1) || make_act_map(&p->tm_s2, dt, p->ptr<void>("s2"), C, T, B, 1)) {
    delete p;
    return 1;
  }
  // fused layer kernel (C = 256, dilations <= 16): haloed loads and 32-row store boxes
  p->fused_ok = (C == kLayerC) && (1 << (c.dilation_cycle_length - 1)) <= kLayerMaxDil;
  p->tail_ok = p->fused_ok && c.in_dims == 128;
  p->ragged = false;
  p->out_scale = 1.f;
  if (p->fused_ok) {
    for (int i = 0; i < 5 && i < c.dilation_cycle_length; ++i)
      if (make_act_map(&p->tm_xa_halo[i], dt, p->ptr<void>("xa"), C, T, B, 1, kBlockM + 2 * (1 << i)) ||
          make_act_map(&p->tm_xb_halo[i], dt, p->ptr<void>("xb"), C, T, B, 1, kBlockM + 2 * (1 << i))) {
        delete p;
        return 1;
      }
    if (make_act_map(&p->tm_xa_st, dt, p->ptr<void>("xa"), C, T, B, 1, 32) || make_act_map(&p->tm_xb_st, dt, p->ptr<void>("xb"), C, T, B, 1, 32) ||
        make_act_map(&p->tm_y_st, dt, p->ptr<void>("ybuf"), C, T, B, c.residual_layers, 32)) {
      delete p;
      return 1;
    }
  }
  *out = p;
  return 0;
}
extern "C" void pd_plan_destroy(pd_plan* p) { delete p; }

extern "C" int pd_plan_buffer(const pd_plan* p, const char* name, size_t* offset, size_t* bytes) {
  if (!p || !name) return fail("null argument");
  for (const Buf& b : p->bufs)
    if (!strcmp(b.name, name)) {
      if (offset) *offset = b.off;
      if (bytes) *bytes = b.bytes;
      return 0;
    }
  return fail("no scratch buffer named '%s'", name);
}

// ------------------------------------------------------------------ GEMM launch
static uint32_t make_idesc(int dtype, int bn, int m = kBlockM) {
  const uint32_t fmt = dtype == PD_FP16 ? 0u : 1u;
  // c_format = F32 (bit 4), a/b format (bits 7-9 / 10-12), K-major A and B, N >> 3 (bits 17-22), M >> 4 (bits 24-28)
  return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

static int g_pair = -1;  // -1: read PD_PAIR from the environment on first use
static bool pair_enabled() {
  if (g_pair < 0) {
    const char* e = getenv("PD_PAIR");
    g_pair = (e && e[0] == '0') ? 0 : 1;
  }
  return g_pair == 1;
}

static thread_local bool g_enc_pdl = false;   // set by the text encoder's launch sequence (pd_encoder_api.cuh)
template <int EPI, int BN, typename T, bool kPair>
static int launch_gemm_t(const CUtensorMap& a, const CUtensorMap& b, const CUtensorMap& o, const CUtensorMap& x,
                         const GemmShape& sh, const EpiArgs& ep, int sms, cudaStream_t s, const char* name, int kind) {
  auto kern = gemm_kernel<EPI, BN, T, kPair>;
  using Cfg = GemmCfg<EPI, BN, kPair>;
  static bool configured[64] = {false};  // per instantiation, per device
  int dev = 0;
  PD_CUDA(cudaGetDevice(&dev));
  if (!configured[dev & 63]) {
    PD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::kSmemBytes));
    configured[dev & 63] = true;
  }
  const int m_tiles = sh.B * sh.m_tiles_per_row;
  const int items = (kPair ? (m_tiles + 1) / 2 : m_tiles) * sh.n_tiles;
  cudaLaunchConfig_t cfg = {};
  const int ctas = kPair ? 2 * (items < sms / 2 ? items : sms / 2) : (items < sms ? items : sms);
  cfg.gridDim = dim3(ctas);
  cfg.blockDim = dim3(kGemmThreads);
  cfg.dynamicSmemBytes = Cfg::kSmemBytes;
  cfg.stream = s;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kPair ? 2 : 1;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;   // encoder chain only (pd_encoder_api.cuh)
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (g_enc_pdl && (EPI == EPI_STORE_X || EPI == EPI_STORE_XH)) ? 2 : 1;
  prof_before(kind, s);
  PD_CUDA(cudaLaunchKernelEx(&cfg, kern, a, b, o, x, sh, ep));
  prof_after(s);
  PD_LAUNCH_CHECK(name);
  return 0;
}
// `bh`: the weight map with half-height boxes; when given (and the problem has at least two frame tiles) the GEMM runs as
// cta_group::2 pairs.
template <int EPI, int BN>
static int launch_gemm(int dtype, const CUtensorMap& a, const CUtensorMap& b, GemmShape sh, const EpiArgs& ep, int sms,
                       cudaStream_t s, const char* name, int kind, const CUtensorMap* o = nullptr,
                       const CUtensorMap* x = nullptr, const CUtensorMap* bh = nullptr) {
  sh.trace = (g_trace && kind >= 0 && kind < PD_N_KINDS) ? g_trace + kind * 1024 : nullptr;   // the buffer holds PD_N_KINDS x 1024 slots
  const CUtensorMap& om = o ? *o : a;  // unused maps alias the A map
  const CUtensorMap& xm = x ? *x : a;
  if (bh != nullptr && pair_enabled() && sh.B * sh.m_tiles_per_row >= 2) {
    sh.idesc = make_idesc(dtype, BN, 256);
    return dtype == PD_FP16 ? launch_gemm_t<EPI, BN, __half, true>(a, *bh, om, xm, sh, ep, sms, s, name, kind)
                            : launch_gemm_t<EPI, BN, __nv_bfloat16, true>(a, *bh, om, xm, sh, ep, sms, s, name, kind);
  }
  sh.idesc = make_idesc(dtype, BN);
  return dtype == PD_FP16 ? launch_gemm_t<EPI, BN, __half, false>(a, b, om, xm, sh, ep, sms, s, name, kind)
                          : launch_gemm_t<EPI, BN, __nv_bfloat16, false>(a, b, om, xm, sh, ep, sms, s, name, kind);
}

static GemmShape base_shape(int B, int T, int n_tiles, int k_per_tap) {
  GemmShape sh;
  memset(&sh, 0, sizeof(sh));
  sh.B = B;
  sh.T = T;
  sh.m_tiles_per_row = (T + kBlockM - 1) / kBlockM;
  sh.n_tiles = n_tiles;
  sh.k_per_tap = k_per_tap;
  sh.taps = 1;
  sh.tap_mode = 0;
  sh.dil = 1;
  return sh;
}
static EpiArgs base_epi() {
  EpiArgs e;
  memset(&e, 0, sizeof(e));
  e.scale = 1.f;
  e.out_scale = 1.f;
  e.out_type = OUT_T;
  return e;
}

// ------------------------------------------------------------------ fused layer kernel launch (2-CTA clusters)
static int g_fused = -1;  // -1: read PD_FUSED from the environment on first use
extern "C" int pd_set_fused(int on) {
  const int prev = g_fused;
  g_fused = on ? 1 : 0;
  return prev;
}
static bool fused_enabled() {
  if (g_fused < 0) {
    const char* e = getenv("PD_FUSED");
    g_fused = (e && e[0] == '0') ? 0 : 1;
  }
  return g_fused == 1;
}

static int g_stack = -1;  // -1: read PD_STACK from the environment on first use
extern "C" int pd_set_stack(int on) {
  const int prev = g_stack;
  g_stack = on ? 1 : 0;
  return prev;
}
static bool stack_enabled() {
  if (g_stack < 0) {
    const char* e = getenv("PD_STACK");
    g_stack = (e && e[0] == '0') ? 0 : 1;
  }
  return g_stack == 1;
}

// Launch the fused ResidualBlock kernel for layers [lp.layer0, lp.layer0 + lp.n_layers).  n_layers > 1 is the persistent
// multi-layer form: its clusters wait on one another through global flags, so every cluster must be co-resident --
// `*ok_multi` reports whether the occupancy query allows that (the caller falls back to one launch per layer).
template <typename T>
static int launch_layer_t(const LayerMaps& mp, const LayerParams& lp, int sms, cudaStream_t s, bool* ok_multi) {
  auto kern = layer_kernel<T>;
  static bool configured[64] = {false};
  int dev = 0;
  PD_CUDA(cudaGetDevice(&dev));
  if (!configured[dev & 63]) {
    PD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kLayerSmemMax));
    configured[dev & 63] = true;
  }
  const int num_pairs = (lp.B * lp.m_tiles_per_row + 1) / 2;
  const int clusters = num_pairs < sms / 2 ? num_pairs : sms / 2;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * clusters);
  cfg.blockDim = dim3(kLayerThreads);
  cfg.dynamicSmemBytes = layer_smem_bytes(lp.xsub, lp.ident_resident);
  cfg.stream = s;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (ok_multi) {
    int max_clusters = 0;
    const cudaError_t e = cudaOccupancyMaxActiveClusters(&max_clusters, kern, &cfg);
    *ok_multi = (e == cudaSuccess) && max_clusters >= clusters;
    if (e != cudaSuccess) cudaGetLastError();
    if (!*ok_multi) return 0;
    // The clusters of a multi-layer launch wait on one another through global flags, so ALL of them must be resident at
    // once.  The occupancy query above only says the grid fits an idle GPU; a COOPERATIVE launch makes the driver schedule
    // the grid as a whole (or not at all), so another stream's or process's kernels holding SMs can delay this launch but
    // cannot leave it half-resident and spinning.  PD_COOP=0 drops the attribute (A/B only).
    // Under an injected tool (Nsight Compute sets NV_COMPUTE_PROFILER_PERFWORKS_DIR / NV_NSIGHT_INJECTION_PORT_BASE in the target
    // process, Nsight Systems and compute-sanitizer CUDA_INJECTION64_PATH) the attribute is dropped as well: ncu cannot replay a cooperative cluster launch, reports LaunchFailed and shuts the
    // application down before the fallback below could run; such tools serialise kernels, so co-residency is not at risk.
    static int coop = -1;
    if (coop < 0) {
      const char* ev = getenv("PD_COOP");
      bool tool = false;
      for (const char* name : {"NV_COMPUTE_PROFILER_PERFWORKS_DIR", "NV_NSIGHT_INJECTION_PORT_BASE", "CUDA_INJECTION64_PATH"}) {
        const char* v = getenv(name);
        tool = tool || (v && v[0] != 0);
      }
      coop = ((ev && ev[0] == '0') || (tool && !(ev && ev[0] == '1'))) ? 0 : 1;
    }
    if (coop) {
      attr[1].id = cudaLaunchAttributeCooperative;
      attr[1].val.cooperative = 1;
      cfg.numAttrs = 2;
    }
  }
  prof_before(8, s);
  const cudaError_t le = cudaLaunchKernelEx(&cfg, kern, mp, lp);
  if (le != cudaSuccess && ok_multi && cfg.numAttrs == 2) {
    // the cooperative form was refused (grid larger than what can be co-resident right now, or not supported together
    // with clusters by this driver): nothing was launched; the caller runs one launch per layer instead
    cudaGetLastError();   // (the profiling slot opened by prof_before is simply reused by the next launch)
    *ok_multi = false;
    return 0;
  }
  if (le != cudaSuccess) return fail("cudaLaunchKernelEx(layer_kernel) failed: %s", cudaGetErrorString(le));
  prof_after(s);
  PD_LAUNCH_CHECK("layer_kernel");
  return 0;
}

static int launch_layer(int dt, const LayerMaps& mp, const LayerParams& lp, int sms, cudaStream_t s, bool* ok_multi) {
  return dt == PD_FP16 ? launch_layer_t<__half>(mp, lp, sms, s, ok_multi) : launch_layer_t<__nv_bfloat16>(mp, lp, sms, s, ok_multi);
}

// ------------------------------------------------------------------ fused sampler tail launch
static int g_tail = -1;  // -1: read PD_TAIL from the environment on first use
extern "C" int pd_set_tail(int on) {
  const int prev = g_tail;
  g_tail = on ? 1 : 0;
  return prev;
}
static bool tail_enabled() {
  if (g_tail < 0) {
    const char* e = getenv("PD_TAIL");
    g_tail = (e && e[0] == '0') ? 0 : 1;
  }
  return g_tail == 1;
}
template <typename T>
static int launch_tail_t(const CUtensorMap& s1, const CUtensorMap& wsp, const CUtensorMap& wout, const CUtensorMap& win,
                         const CUtensorMap& xst, const TailParams& tp, int sms, cudaStream_t s) {
  auto kern = tail_kernel<T>;
  static bool configured[64] = {false};
  int dev = 0;
  PD_CUDA(cudaGetDevice(&dev));
  if (!configured[dev & 63]) {
    PD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTailSmemBytes));
    configured[dev & 63] = true;
  }
  const int tiles = tp.B * tp.m_tiles_per_row;
  prof_before(9, s);
  kern<<<tiles < sms ? tiles : sms, kGemmThreads, kTailSmemBytes, s>>>(s1, wsp, wout, win, xst, tp);
  prof_after(s);
  PD_LAUNCH_CHECK("tail_kernel");
  return 0;
}

static int g_sumtail = -1;  // -1: read PD_SUMTAIL from the environment on first use
extern "C" int pd_set_sumtail(int on) {
  const int prev = g_sumtail;
  g_sumtail = on ? 1 : 0;
  return prev;
}
static bool sumtail_enabled() {
  if (g_sumtail < 0) {
    // Off by default: measured equal to skip_sum + tail as two launches (0.95-0.97 vs 0.96-0.98 ms per 4-step sample;
    // 198.7 vs 134.8 + 62.7 us under ncu).  The tail chain does hide behind the next tile's K loop, but that loop runs ~20 %
    // slower next to it (4 instead of 7 ring stages, x / xa writes mixed into a read stream), S is single-buffered and
    // the last tile's chain is exposed.
    const char* e = getenv("PD_SUMTAIL");
    g_sumtail = (e && e[0] == '1') ? 1 : 0;
  }
  return g_sumtail == 1;
}
template <typename T>
static int launch_sumtail_t(const CUtensorMap& y, const CUtensorMap& wskip, const CUtensorMap& wsp, const CUtensorMap& wout,
                            const CUtensorMap& win, const CUtensorMap& xst, const SumTailParams& sp, int sms, cudaStream_t s) {
  auto kern = sumtail_kernel<T>;
  static bool configured[64] = {false};
  int dev = 0;
  PD_CUDA(cudaGetDevice(&dev));
  if (!configured[dev & 63]) {
    PD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSumTailSmemBytes));
    configured[dev & 63] = true;
  }
  const int pairs = (sp.tp.B * sp.tp.m_tiles_per_row + 1) / 2;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * (pairs < sms / 2 ? pairs : sms / 2));
  cfg.blockDim = dim3(kSumTailThreads);
  cfg.dynamicSmemBytes = kSumTailSmemBytes;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  prof_before(10, s);
  PD_CUDA(cudaLaunchKernelEx(&cfg, kern, y, wskip, wsp, wout, win, xst, sp));
  prof_after(s);
  PD_LAUNCH_CHECK("sumtail_kernel");
  return 0;
}

// ------------------------------------------------------------------ hot-path sequences
struct LenVals {
  int v[64];
};
__global__ void k_set_lens(int* __restrict__ dst, LenVals vals, int n) {
  if (threadIdx.x < n) dst[threadIdx.x] = vals.v[threadIdx.x];
}
extern "C" int pd_plan_set_lengths(pd_plan* p, const int32_t* lengths, void* stream) {
  if (!p) return fail("null argument");
  if (lengths == nullptr) {
    p->ragged = false;
    return 0;
  }
  if (!p->fused_ok) return fail("per-utterance lengths need the fused layer kernel (C = 256, dilation <= 16)");
  for (int b = 0; b < p->B; ++b)
    if (lengths[b] < 0 || lengths[b] > p->T) return fail("lengths[%d]=%d outside [0, T=%d]", b, lengths[b], p->T);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  for (int b0 = 0; b0 < p->B; b0 += 64) {   // by-value kernel arguments: no host memory is read asynchronously
    LenVals lv;
    const int n = p->B - b0 < 64 ? p->B - b0 : 64;
    for (int j = 0; j < n; ++j) lv.v[j] = lengths[b0 + j];
    k_set_lens<<<1, 64, 0, s>>>(p->ptr<int>("lens") + b0, lv, n);
    PD_LAUNCH_CHECK("k_set_lens");
  }
  p->ragged = true;
  return 0;
}

extern "C" int pd_plan_set_output_scale(pd_plan* p, float scale) {
  if (!p) return fail("null argument");
  if (!(scale == scale) || scale == 0.f) return fail("output scale must be a non-zero number, got %g", (double)scale);
  p->out_scale = scale;
  return 0;
}

static int cond_project_from16(pd_plan* p, cudaStream_t s);

extern "C" int pd_cond_project(pd_plan* p, const float* cond, void* stream) {
  if (!p || !cond) return fail("null argument");
  const pd_config& c = p->m->cfg;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const long long n4 = (long long)p->B * p->T * c.hidden_size / 4;
  const int blocks = (int)std::min<long long>((n4 + 255) / 256, 148 * 16);
  if (c.dtype == PD_FP16) k_cast<__half><<<blocks, 256, 0, s>>>(p->ptr<__half>("cond16"), cond, n4);
  else k_cast<__nv_bfloat16><<<blocks, 256, 0, s>>>(p->ptr<__nv_bfloat16>("cond16"), cond, n4);
  PD_LAUNCH_CHECK("k_cast(cond)");
  return cond_project_from16(p, s);
}

extern "C" int pd_cond_assemble(pd_plan* p, const float* encoder_out, int Tph, const long long* mel2ph, const float* f0,
                                const float* w_pitch, const float* b_pitch, const float* utt0, const float* utt1, int n_var,
                                const float* const* var_x, const float* const* var_w, const float* const* var_b,
                                const float* frame_embed, float* cond_out, void* stream) {
  if (!p || !encoder_out || !mel2ph || !f0 || !w_pitch || !b_pitch) return fail("null argument");
  if (Tph < 1 || n_var < 0 || n_var > 2) return fail("bad Tph=%d / n_var=%d", Tph, n_var);
  if (n_var > 0 && (!var_x || !var_w || !var_b)) return fail("variance embeddings: null argument");
  const pd_config& c = p->m->cfg;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  CondAssemble a;
  memset(&a, 0, sizeof(a));
  a.enc = encoder_out;
  a.mel2ph = mel2ph;
  a.f0 = f0;
  a.w_pitch = w_pitch;
  a.b_pitch = b_pitch;
  a.utt[0] = utt0;
  a.utt[1] = utt1;
  for (int j = 0; j < n_var; ++j) {
    if (!var_x[j] || !var_w[j] || !var_b[j]) return fail("variance embedding %d: null argument", j);
    a.var_x[j] = var_x[j];
    a.var_w[j] = var_w[j];
    a.var_b[j] = var_b[j];
  }
  a.frame = frame_embed;
  a.Tph = Tph;
  a.T = p->T;
  a.H = c.hidden_size;
  a.n_var = n_var;
  const long long n4 = (long long)p->B * p->T * c.hidden_size / 4;
  const int blocks = (int)std::min<long long>((n4 + 255) / 256, 148 * 16);
  if (c.dtype == PD_FP16) k_assemble_cond<__half><<<blocks, 256, 0, s>>>(p->ptr<__half>("cond16"), cond_out, a, n4);
  else k_assemble_cond<__nv_bfloat16><<<blocks, 256, 0, s>>>(p->ptr<__nv_bfloat16>("cond16"), cond_out, a, n4);
  PD_LAUNCH_CHECK("k_assemble_cond");
  return cond_project_from16(p, s);
}

// cond16 (the 16-bit [B,T,H] operand in the plan) -> hoisted condition projection of all layers
static int cond_project_from16(pd_plan* p, cudaStream_t s) {
  const pd_model* m = p->m;
  const pd_config& c = m->cfg;
  const int C2 = 2 * c.residual_channels;
  GemmShape sh = base_shape(p->B, p->T, c.residual_layers * C2 / 256, c.hidden_size / 64);
  EpiArgs ep = base_epi();
  ep.out = p->ptr<void>("cp");
  ep.bias = reinterpret_cast<const float*>(m->packed + m->lay.b_cp);
  ep.ld = C2;
  ep.out_type = OUT_F16;
  ep.out_blocked = 1;
  static int blk = -1;   // PD_CP_DIRECT=1: the earlier epilogue with direct st.global (A/B only)
  static int cp_pair = -1;   // PD_CP_PAIR=0: single-CTA tiles instead of cta_group::2 pairs
  if (cp_pair < 0) { const char* e = getenv("PD_CP_PAIR"); cp_pair = (e && e[0] == '0') ? 0 : (pair_enabled() ? 1 : 0); }
  if (blk < 0) { const char* e = getenv("PD_CP_DIRECT"); blk = (e && e[0] == '1') ? 0 : 1; }
  if (!blk) return launch_gemm<EPI_STORE, 256>(c.dtype, p->tm_cond, m->tm_w_cond, sh, ep, m->sm_count, s, "gemm(cond_proj)", 0);
  return launch_gemm<EPI_STORE_BLK, 256>(c.dtype, p->tm_cond, m->tm_w_cond, sh, ep, m->sm_count, s, "gemm(cond_proj)", 0, nullptr, nullptr,
                                         cp_pair ? &m->tm_w_cond_h : nullptr);
}

// Step-embedding chain for `E` entries whose step values are already in the plan's "steps" buffer.
static int run_step_tables(pd_plan* p, int E, cudaStream_t s) {
  const pd_model* m = p->m;
  const pd_config& c = m->cfg;
  const int C = c.residual_channels, L = c.residual_layers;
  const uint8_t* pk = m->packed;
  const PackedLayout& lay = m->lay;
  // sinusoid -> Linear(C,4C) + Mish -> Linear(4C,C) -> L x diffusion_projection (wavenet.py:112-113, :61)
  float* hid = p->ptr<float>("hid");
  k_step_sinusoid<<<E, 128, 0, s>>>(p->ptr<float>("steps"), C, p->ptr<float>("emb0"));
  PD_LAUNCH_CHECK("k_step_sinusoid");
  k_step_matvec<<<dim3(4 * C / 32, E, 1), 256, C * sizeof(float), s>>>((const float*)(pk + lay.f_mlp0w), (const float*)(pk + lay.f_mlp0b),
                                                                      p->ptr<float>("emb0"), hid, 4 * C, C, 1, 1);
  PD_LAUNCH_CHECK("k_step_matvec(mlp0)");
  k_step_matvec<<<dim3(C / 32, E, 1), 256, 4 * C * sizeof(float), s>>>((const float*)(pk + lay.f_mlp2w), (const float*)(pk + lay.f_mlp2b), hid,
                                                                      p->ptr<float>("emb"), C, 4 * C, 1, 0);
  PD_LAUNCH_CHECK("k_step_matvec(mlp2)");
  k_step_matvec<<<dim3(C / 32, E, L), 256, C * sizeof(float), s>>>((const float*)(pk + lay.f_dpw), (const float*)(pk + lay.f_dpb),
                                                                  p->ptr<float>("emb"), p->ptr<float>("dvec"), C, C, L, 0);
  PD_LAUNCH_CHECK("k_step_matvec(diffusion_projection)");
  if (C == 256) k_step_taps<8><<<dim3(2 * C / 8, L), 256, 0, s>>>(p->ptr<float>("dvec"), (const float*)(pk + lay.f_dilw), p->ptr<float>("sb"), E, L);
  else k_step_taps<16><<<dim3(2 * C / 8, L), 256, 0, s>>>(p->ptr<float>("dvec"), (const float*)(pk + lay.f_dilw), p->ptr<float>("sb"), E, L);
  PD_LAUNCH_CHECK("k_step_taps");
  return 0;
}

struct OutStage {  // what the final GEMM's epilogue does with x0
  float* x0_out;
  float* x_state;
  const float* noise;
  bool write_xin;
  bool have_xa;    // layer-0 input already produced by the previous step's fused tail: skip input_projection
  float c1, c2, sigma;
  float out_scale;
  unsigned long long seed;
  unsigned int step_id;
};

// One denoiser evaluation (wavenet.py:100-123) on the operand already in "xin"; entry selection for the step bias is
// entry = entry_base + b * entry_per_b.
static int run_denoiser(pd_plan* p, int entry_base, int entry_per_b, const OutStage& os, cudaStream_t s) {
  const pd_model* m = p->m;
  const pd_config& c = m->cfg;
  const int C = c.residual_channels, L = c.residual_layers, M = c.in_dims, B = p->B, T = p->T, dt = c.dtype, sms = m->sm_count;
  const uint8_t* pk = m->packed;
  const PackedLayout& lay = m->lay;
  // input_projection + ReLU (wavenet.py:108-111)
  if (!os.have_xa) {
    GemmShape sh = base_shape(B, T, C / 256, M / 64);
    EpiArgs ep = base_epi();
    ep.out = p->ptr<void>("xa");
    ep.bias = (const float*)(pk + lay.b_in);
    ep.ld = C;
    ep.relu = 1;
    ep.lens = p->ragged ? p->ptr<int>("lens") : nullptr;
    if (launch_gemm<EPI_STORE, 256>(dt, p->tm_xin, m->tm_w_in, sh, ep, sms, s, "gemm(input_proj)", 1)) return 1;
  }
  bool cur_is_a = true;
  const bool fused = p->fused_ok && fused_enabled();
  if (p->ragged && !fused) return fail("per-utterance lengths need the fused layer kernel (C = 256, dilation <= 16, PD_FUSED != 0)");
  for (int l = 0; l < L; ++l) {
    const int dil = 1 << (l % c.dilation_cycle_length);
    if (fused) {  // ResidualBlocks as the fused 2-CTA kernel: all layers in one persistent launch, or one launch per layer
      const size_t mtiles = (size_t)B * ((T + kBlockM - 1) / kBlockM);
      LayerMaps mp;
      for (int i = 0; i < 5; ++i) {
        const int j = i < c.dilation_cycle_length ? i : 0;
        mp.xh[0][i] = p->tm_xa_halo[j];
        mp.xh[1][i] = p->tm_xb_halo[j];
      }
      mp.xs[0] = p->tm_xa_st;
      mp.xs[1] = p->tm_xb_st;
      mp.wd = m->tm_w_dil128;
      mp.wr = m->tm_w_res128;
      mp.ys = p->tm_y_st;
      LayerParams lp;
      memset(&lp, 0, sizeof(lp));
      lp.B = B;
      lp.T = T;
      lp.m_tiles_per_row = (T + kBlockM - 1) / kBlockM;
      lp.cycle = c.dilation_cycle_length;
      lp.ident_row0 = L * C;
      {
        const int max_dil = 1 << (c.dilation_cycle_length - 1);
        lp.xsub = (int)align_up((size_t)(kBlockM + 2 * max_dil) * 128, 1024);
        static int res_opt = -1;
        if (res_opt < 0) { const char* e = getenv("PD_IDENT_RESIDENT"); res_opt = (e && e[0] == '0') ? 0 : 1; }
        lp.ident_resident = (res_opt && layer_smem_bytes(lp.xsub, 1) <= kLayerSmemMax) ? 1 : 0;
      }
      lp.idesc = (1u << 4) | ((dt == PD_FP16 ? 0u : 1u) << 7) | ((dt == PD_FP16 ? 0u : 1u) << 10) | ((256u >> 3) << 17) | ((256u >> 4) << 24);
      lp.idesc64 = (lp.idesc & ~(0x3Fu << 17)) | ((64u >> 3) << 17);
      lp.cp = p->ptr<uint16_t>("cp");
      lp.cp_layer_stride = (long long)mtiles * kBlockM * 2 * C;
      lp.sb = p->ptr<float>("sb");
      lp.sb_entry_stride = (long long)L * 4 * 2 * C;
      lp.sb_entry_base = entry_base;
      lp.sb_entry_per_b = entry_per_b;
      lp.b_res = (const float*)(pk + lay.b_res);
      lp.scale = 0.70710678118654752440f;
      lp.lens = p->ragged ? p->ptr<int>("lens") : nullptr;
      lp.trace = g_trace ? g_trace + 8 * 1024 : nullptr;
      bool done = false;
      if (stack_enabled() && L > 1) {
        // one persistent launch for the whole stack; tiles of consecutive layers synchronise through per-tile flags
        lp.layer0 = 0;
        lp.n_layers = L;
        lp.rot = 29;
        lp.flags = p->ptr<unsigned int>("flags");
        PD_CUDA(cudaMemsetAsync(lp.flags, 0, (size_t)L * mtiles * 4, s));
        bool ok = false;
        const int rc = launch_layer(dt, mp, lp, sms, s, &ok);
        if (rc) return rc;
        done = ok;
      }
      if (!done) {
        lp.flags = nullptr;
        lp.rot = 0;
        lp.n_layers = 1;
        for (int ll = 0; ll < L; ++ll) {
          lp.layer0 = ll;
          const int rc = launch_layer(dt, mp, lp, sms, s, nullptr);
          if (rc) return rc;
        }
      }
      cur_is_a = (L % 2 == 0);
      break;
    }
    const CUtensorMap& tm_cur = cur_is_a ? p->tm_xa : p->tm_xb;
    const CUtensorMap& tm_nxt = cur_is_a ? p->tm_xb : p->tm_xa;
    {  // dilated conv + cond + step bias -> gate (wavenet.py:63-68)
      GemmShape sh = base_shape(B, T, 2 * C / 256, C / 64);
      sh.taps = 3;
      sh.tap_mode = 1;
      sh.dil = dil;
      sh.b_row0 = l * 2 * C;
      EpiArgs ep = base_epi();
      const size_t mtiles = (size_t)B * ((T + kBlockM - 1) / kBlockM);
      ep.out_slab = l;  // gated activations leave through a TMA store into ybuf[l]
      ep.aux = p->ptr<uint16_t>("cp") + (size_t)l * mtiles * kBlockM * 2 * C;
      ep.sb = p->ptr<float>("sb") + (size_t)l * 4 * 2 * C;
      ep.sb_entry_stride = (long long)L * 4 * 2 * C;
      ep.sb_entry_base = entry_base;
      ep.sb_entry_per_b = entry_per_b;
      ep.ld = C;
      if (launch_gemm<EPI_GATE, 256>(dt, tm_cur, m->tm_w_dil, sh, ep, sms, s, "gemm(dilated_conv+gate)", 2, &p->tm_y, nullptr, &m->tm_w_dil128)) return 1;
    }
    {  // residual half of output_projection + (x + res)/sqrt(2) (wavenet.py:70-72)
      GemmShape sh = base_shape(B, T, C / 256, C / 64);
      sh.a_slab0 = l;
      sh.b_row0 = l * C;
      EpiArgs ep = base_epi();
      ep.bias = (const float*)(pk + lay.b_res) + (size_t)l * C;
      ep.ld = C;
      ep.scale = 0.70710678118654752440f;
      if (launch_gemm<EPI_RES, 256>(dt, p->tm_y, m->tm_w_res, sh, ep, sms, s, "gemm(res_proj)", 3, &tm_nxt, &tm_cur, &m->tm_w_res128)) return 1;
    }
    cur_is_a = !cur_is_a;
  }
  if (os.x_state != nullptr && p->tail_ok && tail_enabled() && sumtail_enabled() && pair_enabled() && B * ((T + kBlockM - 1) / kBlockM) >= 2) {
    // sampling loop: skip sum, skip_projection, output_projection, q-posterior and the next step's input_projection as one
    // kernel (pd_sumtail.cuh): the tail chain of tile t runs on the epilogue warps while the skip-sum K loop of tile t+1 streams
    SumTailParams sp;
    memset(&sp, 0, sizeof(sp));
    TailParams& tp = sp.tp;
    tp.B = B;
    tp.T = T;
    tp.m_tiles_per_row = (T + kBlockM - 1) / kBlockM;
    tp.do_in = os.write_xin ? 1 : 0;
    tp.b_sp = (const float*)(pk + lay.b_sp);
    tp.b_out = (const float*)(pk + lay.b_out);
    tp.b_in = (const float*)(pk + lay.b_in);
    tp.M = M;
    tp.x0_out = os.x0_out;
    tp.x_state = os.x_state;
    tp.noise = os.noise;
    tp.c1 = os.c1;
    tp.c2 = os.c2;
    tp.sigma = os.sigma;
    tp.out_scale = os.out_scale;
    tp.seed = os.seed;
    tp.step_id = os.step_id;
    tp.lens = p->ragged ? p->ptr<int>("lens") : nullptr;
    tp.trace = (g_trace && os.write_xin) ? g_trace + 10 * 1024 : nullptr;   // trace a launch that includes the input_proj GEMM
    sp.L = L;
    sp.idesc256 = make_idesc(dt, 256, 256);
    sp.idesc128 = make_idesc(dt, 128, 256);
    sp.b_sum = (const float*)(pk + lay.b_skipsum);
    sp.scale = 1.0f / sqrtf((float)L);
    return dt == PD_FP16 ? launch_sumtail_t<__half>(p->tm_y, m->tm_w_skip_h, m->tm_w_sp_h, m->tm_w_out_h, m->tm_w_in_h, p->tm_xa_st, sp, sms, s)
                         : launch_sumtail_t<__nv_bfloat16>(p->tm_y, m->tm_w_skip_h, m->tm_w_sp_h, m->tm_w_out_h, m->tm_w_in_h, p->tm_xa_st, sp, sms, s);
  }
  {  // sum of the L skip halves as ONE GEMM with K = L*C, then / sqrt(L)  (wavenet.py:119)
    GemmShape sh = base_shape(B, T, C / 256, C / 64);
    sh.taps = L;
    EpiArgs ep = base_epi();
    ep.out = p->ptr<void>("s1");
    ep.bias = (const float*)(pk + lay.b_skipsum);
    ep.ld = C;
    ep.scale = 1.0f / sqrtf((float)L);
    if (launch_gemm<EPI_STORE, 256>(dt, p->tm_y, m->tm_w_skip, sh, ep, sms, s, "gemm(skip_sum)", 4, nullptr, nullptr, &m->tm_w_skip_h)) return 1;
  }
  if (os.x_state != nullptr && p->tail_ok && tail_enabled()) {
    // sampling loop: skip_projection, output_projection, the q-posterior update and the next step's input_projection as
    // one kernel (pd_tail.cuh)
    TailParams tp;
    memset(&tp, 0, sizeof(tp));
    tp.B = B;
    tp.T = T;
    tp.m_tiles_per_row = (T + kBlockM - 1) / kBlockM;
    tp.do_in = os.write_xin ? 1 : 0;
    tp.idesc256 = make_idesc(dt, 256);
    tp.idesc128 = make_idesc(dt, 128);
    tp.b_sp = (const float*)(pk + lay.b_sp);
    tp.b_out = (const float*)(pk + lay.b_out);
    tp.b_in = (const float*)(pk + lay.b_in);
    tp.M = M;
    tp.x0_out = os.x0_out;
    tp.x_state = os.x_state;
    tp.noise = os.noise;
    tp.c1 = os.c1;
    tp.c2 = os.c2;
    tp.sigma = os.sigma;
    tp.out_scale = os.out_scale;
    tp.seed = os.seed;
    tp.step_id = os.step_id;
    tp.lens = p->ragged ? p->ptr<int>("lens") : nullptr;
    tp.trace = g_trace ? g_trace + 10 * 1024 : nullptr;
    return dt == PD_FP16 ? launch_tail_t<__half>(p->tm_s1, m->tm_w_sp, m->tm_w_out, m->tm_w_in, p->tm_xa_st, tp, sms, s)
                         : launch_tail_t<__nv_bfloat16>(p->tm_s1, m->tm_w_sp, m->tm_w_out, m->tm_w_in, p->tm_xa_st, tp, sms, s);
  }
  {  // skip_projection + ReLU (wavenet.py:120-121)
    GemmShape sh = base_shape(B, T, C / 256, C / 64);
    EpiArgs ep = base_epi();
    ep.out = p->ptr<void>("s2");
    ep.bias = (const float*)(pk + lay.b_sp);
    ep.ld = C;
    ep.relu = 1;
    if (launch_gemm<EPI_STORE, 256>(dt, p->tm_s1, m->tm_w_sp, sh, ep, sms, s, "gemm(skip_proj)", 5)) return 1;
  }
  {  // output_projection (wavenet.py:122) + q-posterior update (prodiff.py:106-121)
    GemmShape sh = base_shape(B, T, 1, C / 64);
    EpiArgs ep = base_epi();
    ep.bias = (const float*)(pk + lay.b_out);
    ep.M = M;
    ep.Mp = M;
    ep.x0_out = os.x0_out;
    ep.x_state = os.x_state;
    ep.noise = os.noise;
    ep.xin_next = os.write_xin ? p->ptr<void>("xin") : nullptr;
    ep.c1 = os.c1;
    ep.c2 = os.c2;
    ep.sigma = os.sigma;
    ep.out_scale = os.out_scale;
    ep.seed = os.seed;
    ep.step_id = os.step_id;
    if (launch_gemm<EPI_OUT, 128>(dt, p->tm_s2, m->tm_w_out, sh, ep, sms, s, "gemm(output_proj+posterior)", 6)) return 1;
  }
  return 0;
}

static int spec_to_xin(pd_plan* p, const float* spec, cudaStream_t s) {
  const pd_config& c = p->m->cfg;
  const dim3 grid((p->T + 31) / 32, c.in_dims / 32, p->B), block(32, 8);
  if (c.dtype == PD_FP16) k_spec_to_xin<__half><<<grid, block, 0, s>>>(p->ptr<__half>("xin"), spec, c.in_dims, c.in_dims, p->T);
  else k_spec_to_xin<__nv_bfloat16><<<grid, block, 0, s>>>(p->ptr<__nv_bfloat16>("xin"), spec, c.in_dims, c.in_dims, p->T);
  PD_LAUNCH_CHECK("k_spec_to_xin");
  return 0;
}

extern "C" int pd_denoise(pd_plan* p, const float* spec, const float* diffusion_step, float* out, void* stream) {
  if (!p || !spec || !diffusion_step || !out) return fail("null argument");
  if (p->E < p->B) return fail("plan was created with max_entries=%d < B=%d; per-row steps need one entry per row", p->E, p->B);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  PD_CUDA(cudaMemcpyAsync(p->ptr<float>("steps"), diffusion_step, (size_t)p->B * 4, cudaMemcpyDeviceToDevice, s));
  if (run_step_tables(p, p->B, s)) return 1;
  if (spec_to_xin(p, spec, s)) return 1;
  OutStage os;
  memset(&os, 0, sizeof(os));
  os.out_scale = 1.f;
  os.x0_out = out;
  return run_denoiser(p, 0, 1, os, s);
}

extern "C" int pd_sample(pd_plan* p, const float* cond, float* x, const float* noise, uint64_t seed, int K,
                         const int32_t* t_list, const float* coef1, const float* coef2, const float* logvar, int n_table,
                         float* x0_trace, void* stream) {
  if (!p || !x || !t_list || !coef1 || !coef2 || !logvar) return fail("null argument");
  if (K < 1 || K > p->E) return fail("K=%d outside [1, max_entries=%d] of this plan", K, p->E);
  for (int j = 0; j < K; ++j)
    if (t_list[j] < 0 || t_list[j] >= n_table) return fail("t_list[%d]=%d outside the %d-entry tables", j, t_list[j], n_table);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const pd_config& c = p->m->cfg;
  const size_t state_elems = (size_t)p->B * c.in_dims * p->T;
  // step values -> device (by-value kernel arguments: no host memory is read asynchronously)
  for (int j0 = 0; j0 < K; j0 += 64) {
    StepVals sv;
    const int n = K - j0 < 64 ? K - j0 : 64;
    for (int j = 0; j < n; ++j) sv.v[j] = (float)t_list[j0 + j];
    k_set_steps<<<1, 64, 0, s>>>(p->ptr<float>("steps") + j0, sv, n);
    PD_LAUNCH_CHECK("k_set_steps");
  }
  if (run_step_tables(p, K, s)) return 1;
  // once per utterance, reused by all K steps; cond == NULL: the projection is already in the plan (pd_cond_project /
  // pd_cond_assemble was called on it for this batch)
  if (cond != nullptr && pd_cond_project(p, cond, stream)) return 1;
  if (spec_to_xin(p, x, s)) return 1;
  for (int j = 0; j < K; ++j) {
    const int t = t_list[j];
    OutStage os;
    os.x0_out = x0_trace ? x0_trace + (size_t)j * state_elems : nullptr;
    os.x_state = x;
    os.noise = noise ? noise + (size_t)j * state_elems : nullptr;
    os.write_xin = j + 1 < K;
    os.have_xa = j > 0 && p->tail_ok && tail_enabled();
    os.c1 = coef1[t];
    os.c2 = coef2[t];
    os.sigma = t == 0 ? 0.f : expf(0.5f * logvar[t]);  // nonzero_mask * exp(0.5 * log_var), prodiff.py:120-121
    os.out_scale = j + 1 == K ? p->out_scale : 1.f;
    os.seed = seed;
    os.step_id = (unsigned int)j;
    if (run_denoiser(p, j, 0, os, s)) return 1;
  }
  return 0;
}

// ------------------------------------------------------------------ rectified-flow sampler (reflow.py:48-101)
namespace {
struct RkTableau {
  int stages;
  // stage i (i >= 1) evaluates the velocity at x + (((sum_j a[i][j] k_j) * mult[i]) * dt) / div[i], time t + c[i] * dt;
  // row `stages` is the final update.  Written exactly as the reference expressions are (reflow.py:49-84).
  double c[6];
  float a[7][6];
  float mult[7], div[7];
};
const RkTableau kEuler = {1, {0}, {{0}, {1}}, {1, 1}, {1, 1}};
const RkTableau kRk2 = {2, {0, 0.5}, {{0}, {1}, {0, 1}}, {1, 0.5f, 1}, {1, 1, 1}};
const RkTableau kRk4 = {4, {0, 0.5, 0.5, 1.0}, {{0}, {1}, {0, 1}, {0, 0, 1}, {1, 2, 2, 1}}, {1, 0.5f, 0.5f, 1, 1}, {1, 1, 1, 1, 6}};
const RkTableau kRk5 = {6,
                        {0, 0.25, 0.25, 0.5, 0.75, 1.0},
                        {{0}, {1}, {1, 1}, {0, -1, 2}, {3, 0, 0, 9}, {-3, 2, 12, -12, 8}, {7, 0, 32, 12, 32, 7}},
                        {1, 0.25f, 0.125f, 0.5f, 0.0625f, 1, 1},
                        {1, 1, 1, 1, 1, 7, 90}};
}  // namespace

static int rk_stage(pd_plan* p, const RkTableau& tb, int row, float* const* k, float dt, const float* x, float* x_out, bool want_xin,
                    cudaStream_t s) {
  const pd_config& c = p->m->cfg;
  RkStage st;
  memset(&st, 0, sizeof(st));
  // terms with a zero weight are skipped like the reference's expressions skip them; order of the others is kept
  for (int j = 0; j < 6; ++j)
    if (tb.a[row][j] != 0.f) {
      st.k[st.n_k] = k[j];
      st.w[st.n_k] = tb.a[row][j];
      ++st.n_k;
    }
  st.mult = tb.mult[row];
  st.dt = dt;
  st.div = tb.div[row];
  const dim3 grid((p->T + 31) / 32, (c.in_dims + 31) / 32, p->B), block(32, 8);
  if (c.dtype == PD_FP16) k_rk_stage<__half><<<grid, block, 0, s>>>(want_xin ? p->ptr<__half>("xin") : nullptr, x_out, x, st, c.in_dims, c.in_dims, p->T);
  else k_rk_stage<__nv_bfloat16><<<grid, block, 0, s>>>(want_xin ? p->ptr<__nv_bfloat16>("xin") : nullptr, x_out, x, st, c.in_dims, c.in_dims, p->T);
  PD_LAUNCH_CHECK("k_rk_stage");
  return 0;
}

extern "C" int pd_reflow_sample(pd_plan* p, const float* cond, float* x, int algorithm, int n_steps, float time_scale,
                                float* k_scratch, const float* spec_min, const float* spec_max, int n_spec, float* v_trace,
                                void* stream) {
  if (!p || !x || !k_scratch) return fail("null argument");
  if (algorithm < 0 || algorithm > 3) return fail("algorithm %d is not one of 0 euler, 1 rk2, 2 rk4, 3 rk5", algorithm);
  if (n_steps < 1) return fail("n_steps=%d must be >= 1", n_steps);
  const RkTableau& tb = algorithm == 0 ? kEuler : algorithm == 1 ? kRk2 : algorithm == 2 ? kRk4 : kRk5;
  if (p->E < tb.stages) return fail("plan was created with max_entries=%d < %d stages of this integrator", p->E, tb.stages);
  const pd_config& c = p->m->cfg;
  if ((spec_min == nullptr) != (spec_max == nullptr) || (spec_min && n_spec != 1 && n_spec < c.in_dims))
    return fail("spec_min / spec_max must both be given with 1 or >= in_dims entries");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const size_t state_elems = (size_t)p->B * c.in_dims * p->T;
  const double dt = 1.0 / (double)n_steps;   // reflow.py:89
  const float dtf = (float)dt;               // the fp32 tensor `dts` (:97)
  if (cond != nullptr && pd_cond_project(p, cond, stream)) return 1;  // once per utterance, reused by every evaluation (NULL: already projected)
  if (spec_to_xin(p, x, s)) return 1;
  for (int n = 0; n < n_steps; ++n) {
    // times of this step's evaluations, in the reference's fp32 arithmetic: time_scale * (i * dts + c * dt)
    StepVals sv;
    const float tn = (float)n * dtf;
    for (int i = 0; i < tb.stages; ++i) sv.v[i] = time_scale * (i == 0 ? tn : tn + (float)(tb.c[i] * dt));
    k_set_steps<<<1, 64, 0, s>>>(p->ptr<float>("steps"), sv, tb.stages);
    PD_LAUNCH_CHECK("k_set_steps");
    if (run_step_tables(p, tb.stages, s)) return 1;
    float* k[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    for (int i = 0; i < tb.stages; ++i)
      k[i] = v_trace ? v_trace + ((size_t)n * tb.stages + i) * state_elems : k_scratch + (size_t)i * state_elems;
    for (int i = 0; i < tb.stages; ++i) {
      OutStage os;
      memset(&os, 0, sizeof(os));
      os.out_scale = 1.f;
      os.x0_out = k[i];   // the denoiser output is the velocity (reflow.py:50)
      if (run_denoiser(p, i, 0, os, s)) return 1;
      if (i + 1 < tb.stages) {
        if (rk_stage(p, tb, i + 1, k, dtf, x, nullptr, true, s)) return 1;          // next evaluation point -> xin
      } else {
        if (rk_stage(p, tb, tb.stages, k, dtf, x, x, n + 1 < n_steps, s)) return 1;  // x += ..., and xin for the next step
      }
    }
  }
  if (spec_min != nullptr) {
    k_denorm_spec<<<p->m->sm_count * 4, 256, 0, s>>>(x, spec_min, spec_max, n_spec, c.in_dims, p->T, (long long)state_elems, p->out_scale);
    PD_LAUNCH_CHECK("k_denorm_spec");
  }
  return 0;
}

// ------------------------------------------------------------------ test hook
extern "C" int pd_debug_gemm(int dtype, const void* a, int B, int T, int Cin, int taps, int dil, const void* w, int N,
                             const float* bias, float* out, void* stream) {
  if (!a || !w || !bias || !out) return fail("null argument");
  if (N % 256 || Cin % 64 || (taps != 1 && taps != 3)) return fail("unsupported debug GEMM shape");
  int device = 0, sms = 0;
  if (check_device(&device, &sms)) return 1;
  CUtensorMap ta, tb;
  if (make_act_map(&ta, dtype, a, Cin, T, B, 1) || make_w_map(&tb, dtype, w, taps * Cin, N, 256)) return 1;
  GemmShape sh = base_shape(B, T, N / 256, Cin / 64);
  sh.taps = taps;
  sh.tap_mode = 1;
  sh.dil = dil;
  EpiArgs ep = base_epi();
  ep.out = out;
  ep.bias = bias;
  ep.ld = N;
  ep.out_type = OUT_F32;
  return launch_gemm<EPI_STORE, 256>(dtype, ta, tb, sh, ep, sms, static_cast<cudaStream_t>(stream), "gemm(debug)", 7);
}

// ------------------------------------------------------------------ experiment: row-shifted SW128 A descriptors
// One CTA: A tile [R x 64] (16-bit, TMA SWIZZLE_128B, 1024-aligned), B tile [64 x 64]; D[128 x 64] = A[shift : shift+128] * B^T.
// mode 0: descriptor base_offset = 0; mode 1: base_offset = (start_address >> 7) & 7.  Decides whether the dilated-conv
// taps can be served as shifted views of ONE resident activation tile.
__global__ void __launch_bounds__(128, 1)
k_shift_probe(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int R, int shift, int mode,
              uint32_t idesc, float* __restrict__ out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sA = smem;                 // R * 128 bytes (R <= 256)
  uint8_t* sB = smem + 256 * 128;     // 64 * 128 bytes
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + 64 * 128);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc<64>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(&bars[0], (uint32_t)(R * 128 + 64 * 128));
    tma_load_2d(sA, &tmA, &bars[0], 0, 0);
    tma_load_2d(sB, &tmB, &bars[0], 0, 0);
    mbar_wait(&bars[0], 0, 10);
    tc_fence_after();
    const uint32_t a_addr = smem_u32(sA) + (uint32_t)shift * 128u;
    uint64_t adesc = make_sw128_kmajor_desc(a_addr);
    if (mode == 1) adesc |= (uint64_t)((a_addr >> 7) & 7u) << 49;
    const uint64_t bdesc = make_sw128_kmajor_desc(smem_u32(sB));
    for (int k = 0; k < 4; ++k) umma_f16(tmem_base, desc_advance_k(adesc, k), desc_advance_k(bdesc, k), idesc, k != 0);
    umma_commit(&bars[1]);
  }
  mbar_wait(&bars[1], 0, 11);
  tc_fence_after();
  const uint32_t trow = tmem_base + ((uint32_t)(warp * 32) << 16);
  uint32_t r[32];
  for (int c = 0; c < 64; c += 32) {
    tmem_ld_x32(trow + c, r);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) out[(warp * 32 + lane) * 64 + c + j] = __uint_as_float(r[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<64>(tmem_base);
}

extern "C" int pd_debug_shift(int dtype, const void* a, int R, const void* w, int shift, int mode, float* out, void* stream) {
  if (!a || !w || !out) return fail("null argument");
  if (R < 128 || R > 256 || shift < 0 || shift + 128 > R) return fail("bad R/shift");
  int device = 0, sms = 0;
  if (check_device(&device, &sms)) return 1;
  CUtensorMap ta, tb;
  const uint64_t da[2] = {64, (uint64_t)R}, db[2] = {64, 64};
  if (make_map(&ta, dtype, a, 2, da, (uint32_t)R) || make_map(&tb, dtype, w, 2, db, 64)) return 1;
  const uint32_t fmt = dtype == PD_FP16 ? 0u : 1u;
  const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const int smem = 256 * 128 + 64 * 128 + 1024 + 64;
  PD_CUDA(cudaFuncSetAttribute(k_shift_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  k_shift_probe<<<1, 128, smem, static_cast<cudaStream_t>(stream)>>>(ta, tb, R, shift, mode, idesc, out);
  PD_LAUNCH_CHECK("k_shift_probe");
  return 0;
}

// ------------------------------------------------------------------ FastSpeech text encoder
#include "pd_encoder_api.cuh"

// ------------------------------------------------------------------ NSF-HiFiGAN generator
#include "pd_voc_api.cuh"
