/* prodiff_b200 -- C ABI of the B200-native ProDiff sampling hot path.
 *
 * The reference (T1uTiu/ProDiff) is pure Python/PyTorch and has no FFI; the boundary it offers for this path is
 * two nn.Module classes:
 *
 *   WaveNet.forward(spec, diffusion_step, cond)                       modules/decoder/wavenet.py:100-123
 *   GaussianDiffusion.forward(cond, ..., infer_step, infer=True)      modules/diffusion/prodiff.py:136-153
 *   GaussianDiffusion.p_sample / q_posterior_sample                   modules/diffusion/prodiff.py:106-126
 *
 * This header declares what a binding for those two classes calls.  Plain pointers and sizes only: no torch
 * types, no C++ types.  All device pointers are CUDA device memory owned by the caller (in the Python host they
 * are torch tensors); the library never allocates device memory and never keeps an input pointer after a call
 * returns, except the `packed` and `workspace` blocks the caller hands to pd_model_create / pd_plan_create, which
 * must outlive the handle.  Every function returns 0 on success and a non-zero code on failure;
 * pd_last_error() then returns a thread-local message.  There is no CPU fallback: a device that is not
 * compute capability 10.x is an error.
 *
 * Threading: a pd_model is immutable after creation; a pd_plan owns scratch buffers and must be used by one
 * stream at a time.  All work is enqueued on the caller's stream and is asynchronous with respect to the host.
 */
#ifndef PRODIFF_B200_H
#define PRODIFF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PD_ABI_VERSION 1

enum { PD_BF16 = 0, PD_FP16 = 1 }; /* tensor-core operand type; accumulation is always fp32 */

/* Constructor arguments of the reference WaveNet (wavenet.py:75) plus the operand type. */
typedef struct pd_config {
  int32_t in_dims;               /* M: mel bins; 64 or 128                              */
  int32_t hidden_size;           /* H: condition channels; multiple of 64               */
  int32_t residual_layers;       /* L                                                   */
  int32_t residual_channels;     /* C: multiple of 256                                  */
  int32_t dilation_cycle_length; /* dilation of layer i = 2^(i % cycle)  (wavenet.py:93) */
  int32_t dtype;                 /* PD_BF16 / PD_FP16                                   */
} pd_config;

typedef struct pd_model pd_model; /* packed weights + weight tensor maps (per device)            */
typedef struct pd_plan pd_plan;   /* scratch layout + activation tensor maps for one (B, T) shape */

/* Number of entries pd_model_create expects in `state`: 10 + 8 * L. Order (all fp32, contiguous, on device):
 *   0 input_projection.weight [C,M,1]   1 input_projection.bias [C]
 *   2 mlp.0.weight [4C,C]               3 mlp.0.bias [4C]
 *   4 mlp.2.weight [C,4C]               5 mlp.2.bias [C]
 *   6 skip_projection.weight [C,C,1]    7 skip_projection.bias [C]
 *   8 output_projection.weight [M,C,1]  9 output_projection.bias [M]
 *   then for each layer i (base = 10 + 8*i), the keys of residual_layers.{i}. in this order:
 *   +0 dilated_conv.weight [2C,C,3]            +1 dilated_conv.bias [2C]
 *   +2 diffusion_projection.weight [C,C]       +3 diffusion_projection.bias [C]
 *   +4 conditioner_projection.weight [2C,H,1]  +5 conditioner_projection.bias [2C]
 *   +6 output_projection.weight [2C,C,1]       +7 output_projection.bias [2C]
 * (the reference's state_dict keys) */
int pd_state_count(const pd_config* cfg);

const char* pd_last_error(void);
int pd_abi_version(void);

/* Bytes of device memory the packed (16-bit, tile-ordered) weights need. */
size_t pd_packed_bytes(const pd_config* cfg);

/* Pack the fp32 state into `packed` (device, >= pd_packed_bytes, 1024-byte aligned) on `stream` and build the
 * model handle.  Replaces: the reference keeping fp32 nn.Parameters and dispatching cuDNN/cuBLAS per op. */
int pd_model_create(const pd_config* cfg, const float* const* state, int n_state, void* packed,
                    size_t packed_bytes, void* stream, pd_model** out);
void pd_model_destroy(pd_model* m);

/* Scratch bytes for B utterances of T frames and up to `max_entries` distinct diffusion steps per call
 * (K for sampling, B for a per-row-step denoise call). */
size_t pd_workspace_bytes(const pd_model* m, int B, int T, int max_entries);
int pd_plan_create(const pd_model* m, int B, int T, int max_entries, void* workspace, size_t workspace_bytes,
                   pd_plan** out);
void pd_plan_destroy(pd_plan* p);

/* Hoisted condition projection: all L conditioner_projection 1x1 convs (wavenet.py:62) of cond [B,T,H] fp32
 * (physically frame-major, i.e. the tensor the reference transposes at prodiff.py:138) in one GEMM, kept in the
 * plan for every following pd_denoise call.  pd_sample calls it itself. */
int pd_cond_project(pd_plan* p, const float* cond, void* stream);

/* WaveNet.forward (wavenet.py:100-123): spec [B,1,M,T] fp32, diffusion_step [B] fp32 (device), out [B,1,M,T]
 * fp32.  Uses the condition projection already in the plan. */
int pd_denoise(pd_plan* p, const float* spec, const float* diffusion_step, float* out, void* stream);

/* GaussianDiffusion.forward(infer=True) (prodiff.py:145-153): the whole K-step reverse loop.
 *   cond   [B,T,H] fp32 (NULL: the projection is already in the plan -- pd_cond_project / pd_cond_assemble ran for this batch)
 *   x      [B,1,M,T] fp32: in = x_T (the reference draws it with torch.rand, :147), out = x_0
 *   noise  [K,B,1,M,T] fp32 or NULL: noise[j] is used by the j-th iteration (replaces torch.randn, :118);
 *          NULL = in-kernel Philox4x32-10 keyed by `seed`
 *   t_list host int[K]: the step index of each iteration (the reference uses K-1, ..., 0)
 *   coef1/coef2/logvar host float[n_table]: posterior_mean_coef1/2 and posterior_log_variance_clipped buffers
 *   x0_trace [K,B,1,M,T] fp32 or NULL: per-step denoiser outputs (for parity checks) */
int pd_sample(pd_plan* p, const float* cond, float* x, const float* noise, uint64_t seed, int K,
              const int32_t* t_list, const float* coef1, const float* coef2, const float* logvar, int n_table,
              float* x0_trace, void* stream);

/* Frame-level half of ProDiffTeacher.forward_condition (/root/reference/modules/svs/prodiff_teacher.py:127-149; SURVEY 8f
 * N1): length regulation (F.pad + gather by mel2ph, :128-130), + pitch_embed(log(1 + f0 / 700)) (:98-101, :132), + up to two
 * utterance-level embedding rows (speaker :135, gender :137), + an optional frame-level embedding (a per-frame speaker mix),
 * + the sum of up to two Linear(1,H) variance embeddings (voicing, breath, :139-147), times the non-padding mask
 * (mel2ph > 0, :148-149) -- written straight into the plan's 16-bit condition operand, followed by the condition projection
 * of every layer (what pd_cond_project does after its cast).  The token-level encoder stays with the caller.
 *   encoder_out [B,Tph,H] fp32; mel2ph [B,T] int64 (0 = padding, i > 0 = token i - 1); f0 [B,T] fp32;
 *   w_pitch / b_pitch [H] (pitch_embed.weight[:, 0], .bias); utt0 / utt1 optional [B,H]; var_x[j] [B,T], var_w[j] / var_b[j] [H]
 *   for j < n_var <= 2; frame_embed optional [B,T,H]; cond_out optional [B,T,H] fp32 (the condition forward_condition
 *   returns).  All device pointers.  Afterwards pd_sample / pd_reflow_sample are called with cond = NULL. */
int pd_cond_assemble(pd_plan* p, const float* encoder_out, int Tph, const long long* mel2ph, const float* f0,
                     const float* w_pitch, const float* b_pitch, const float* utt0, const float* utt1, int n_var,
                     const float* const* var_x, const float* const* var_w, const float* const* var_b,
                     const float* frame_embed, float* cond_out, void* stream);

/* Ragged batches (no counterpart in the reference, which pads and processes every row at the full T, SURVEY appendix A.7;
 * needed to batch its one-utterance-at-a-time offline sampling loop, component/binarizer/svs.py:193-223, without changing
 * results): lengths[b] <= T is the number of valid frames of utterance b (host array of B entries, copied by value).
 * Subsequent pd_denoise / pd_sample / pd_reflow_sample calls on this plan keep every layer's activations at zero past
 * lengths[b] and choose the step-bias edge slots from lengths[b], so frames t < lengths[b] get exactly what a run of that
 * utterance alone at T = lengths[b] computes (zero padding of x + d at its own end, wavenet.py:63-65); outputs at
 * t >= lengths[b] are unspecified.  NULL restores full-length rows.  Needs the fused layer kernel (C = 256). */
int pd_plan_set_lengths(pd_plan* p, const int32_t* lengths, void* stream);

/* Vocoder hand-off (first sub-step).  The reference's NsfHifiGAN.spec2wav_torch
 * (/root/reference/component/vocoder/nsf_hifigan.py:49-52) transposes the sampler's [B,T,M] result back to [B,M,T] and
 * multiplies it by 2.30259 (log10 -> ln mel) before Generator.forward -- two more passes over the mel.  [B,M,T] is the
 * physical layout of the sampler state x, so with scale != 1 installed the LAST posterior update of pd_sample (or the
 * denorm_spec kernel of pd_reflow_sample) stores fl(scale * x) instead of x: the state the call leaves behind is the
 * generator's input, bit-identical to `2.30259 * mel.transpose(2, 1)` on the unscaled result (one extra fp32 rounding, no
 * extra kernel, no extra HBM traffic).  scale = 1 (the default) restores the reference's return value. */
int pd_plan_set_output_scale(pd_plan* p, float scale);

/* RectifiedFlow.forward(infer=True) (/root/reference/modules/diffusion/reflow.py:32-46, 86-107): the rectified-flow ODE
 * x' = v(x, time_scale * t) integrated from t = 0 to 1 in n_steps steps of dt = 1 / n_steps with sample_euler (:48-52),
 * sample_rk2 (:54-60), sample_rk4 (:62-70) or sample_rk5 (:72-84), the velocity being the same DiffNet denoiser
 * (WaveNet.forward with a float diffusion step shared by the batch), followed by denorm_spec (:106-107).
 *   cond      [B,T,H] fp32 device; projected once and reused by every evaluation (NULL: pd_cond_project / pd_cond_assemble
 *             already ran on this plan for this batch).
 *   x         [B,1,M,T] fp32 device, in place: in = the N(0,1) start the reference draws at :88, out = denormalised result
 *             (the caller returns the [B,T,M] view x[:,0].transpose(1,2) like :100).
 *   algorithm 0 euler, 1 rk2, 2 rk4, 3 rk5 (the caller maps unknown names to euler as the reference's dict.get does, :95).
 *   k_scratch fp32 device scratch for the stage velocities: stages * B*M*T floats (stages = 1, 2, 4, 6).
 *   spec_min / spec_max  device fp32, n_spec = 1 or >= M entries; both NULL skips denorm_spec.
 *   v_trace   optional [n_steps * stages, B,1,M,T] fp32: every velocity in call order (then k_scratch is unused).
 * The plan needs max_entries >= stages. */
int pd_reflow_sample(pd_plan* p, const float* cond, float* x, int algorithm, int n_steps, float time_scale,
                     float* k_scratch, const float* spec_min, const float* spec_max, int n_spec, float* v_trace, void* stream);

/* ---- FastSpeech text encoder (token-level half) -------------------------------------------------------
 * FastspeechEncoder.forward(txt_tokens, extra_embed) in eval mode (/root/reference/modules/fastspeech/tts_modules.py:291-330,
 * FFTBlocks.forward :262-288, EncSALayer /root/reference/modules/commons/common_layers.py:625-673): what
 * ProDiffTeacher.forward_condition calls at /root/reference/modules/svs/prodiff_teacher.py:122 to produce the
 * `encoder_out` that pd_cond_assemble consumes.  Constructor arguments of the reference class plus the operand type. */
typedef struct pd_encoder_config {
  int32_t vocab_size;
  int32_t hidden_size; /* 256 */
  int32_t num_layers;
  int32_t kernel_size; /* odd; the FFN's Conv1d (enc_ffn_kernel_size, 9) */
  int32_t num_heads;   /* 2 or 4 */
  int32_t dtype;       /* PD_BF16 / PD_FP16 tensor-core operands; LayerNorm, softmax, residual stream in fp32 */
} pd_encoder_config;
typedef struct pd_encoder pd_encoder;

/* Entries pd_encoder_create expects in `state`: 3 + 10 * num_layers device fp32 tensors, the reference's state_dict keys:
 *   0 embed_tokens.weight [V,H]   1 layer_norm.weight [H]   2 layer_norm.bias [H]
 *   then for layer i (base = 3 + 10*i) the keys of layers.{i}.op. in this order:
 *   +0 layer_norm1.weight  +1 layer_norm1.bias  +2 self_attn.in_proj_weight [3H,H]  +3 self_attn.out_proj.weight [H,H]
 *   +4 layer_norm2.weight  +5 layer_norm2.bias  +6 ffn.ffn_1.weight [4H,H,k]  +7 ffn.ffn_1.bias [4H]
 *   +8 ffn.ffn_2.weight [H,4H]  +9 ffn.ffn_2.bias [H] */
int pd_encoder_state_count(const pd_encoder_config* cfg);
size_t pd_encoder_packed_bytes(const pd_encoder_config* cfg);
/* Packs the weights into `packed` (device, 1024-byte aligned, pd_encoder_packed_bytes; owned by the caller, must outlive
 * the handle).  Same ownership rules as pd_model_create. */
int pd_encoder_create(const pd_encoder_config* cfg, const float* const* state, int n_state, void* packed,
                      size_t packed_bytes, void* stream, pd_encoder** out);
void pd_encoder_destroy(pd_encoder* e);
size_t pd_encoder_workspace_bytes(const pd_encoder* e, int B, int T);
/* txt_tokens [B,T] int64 device (0 = padding, tts_modules.py:311); extra_embed [B,T,H] fp32 device or NULL (what
 * dur_embed + lang_embed add, prodiff_teacher.py:112-120); out [B,T,H] fp32 device: encoder_out, zeros at padding tokens.
 * A row whose tokens are all padding comes out as zeros (the reference produces NaN there).  workspace: device scratch of
 * pd_encoder_workspace_bytes(e, B, T), 1024-byte aligned, only used during the call's kernels. */
int pd_encoder_forward(pd_encoder* e, const long long* txt_tokens, const float* extra_embed, float* out, int B, int T,
                       void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * NSF-HiFiGAN generator: the step right after the sampler.  Replaces Generator.forward(x, f0)
 * (modules/nsf_hifigan/models.py:269-290, with ResBlock1 :39-76, SineGen :103-188 and SourceModuleHnNSF :191-216) as
 * called by NsfHifiGAN.spec2wav_torch (component/vocoder/nsf_hifigan.py:49-58).  Every convolution runs as an implicit
 * GEMM on tcgen05 (prodiff_b200/csrc/pd_voc.cuh), fp16 operands, fp32 accumulation; no cuDNN, no fallback. */
typedef struct pd_vocoder_config {
  int32_t num_mels;                  /* h.num_mels: 128 */
  int32_t upsample_initial_channel;  /* h.upsample_initial_channel: 512 */
  int32_t n_upsamples;               /* len(h.upsample_rates) <= 8 */
  int32_t upsample_rates[8];         /* e.g. 8 8 2 2 2 */
  int32_t upsample_kernel_sizes[8];  /* must be 2 * rate (e.g. 16 16 4 4 4) */
  int32_t n_resblock_kernels;        /* len(h.resblock_kernel_sizes) <= 4 */
  int32_t resblock_kernel_sizes[4];  /* e.g. 3 7 11 */
  int32_t resblock_dilation_sizes[4][3]; /* e.g. 1 3 5 each */
  int32_t resblock_type;             /* 1 = ResBlock1 (h.resblock == '1'); the only one implemented */
  int32_t sampling_rate;             /* h.sampling_rate: 44100 */
} pd_vocoder_config;
typedef struct pd_vocoder pd_vocoder;

/* State tensors (fp32, device, contiguous; weight norm already folded, i.e. what Generator.remove_weight_norm() leaves,
 * models.py:292-297), in this order:
 *   conv_pre.weight [C0, num_mels, 7], conv_pre.bias;
 *   for each up-sampling stage i: ups.i.weight [Cin, Cin/2, k], ups.i.bias, noise_convs.i.weight [Cin/2, 1, k'], noise_convs.i.bias,
 *     then for each resblock kernel j (resblocks.{i * n_kernels + j}): convs1.{0,1,2}.weight, .bias interleaved (6 tensors),
 *     convs2.{0,1,2}.weight, .bias interleaved (6 tensors);
 *   conv_post.weight [1, C_last, 7], conv_post.bias; m_source.l_linear.weight [1, 9], m_source.l_linear.bias [1]. */
int pd_vocoder_state_count(const pd_vocoder_config* cfg);
size_t pd_vocoder_packed_bytes(const pd_vocoder_config* cfg);   /* 0 = unsupported config, see pd_last_error() */
int pd_vocoder_create(const pd_vocoder_config* cfg, const float* const* state, int n_state, void* packed, size_t packed_bytes,
                      void* stream, pd_vocoder** out);
void pd_vocoder_destroy(pd_vocoder* v);
int pd_vocoder_upp(const pd_vocoder* v);                         /* samples per mel frame (prod of the rates, models.py:267) */
size_t pd_vocoder_workspace_bytes(const pd_vocoder* v, int B, int T);
/* c [B, num_mels, T] fp32 device: the generator input (2.30259 * mel.transpose(2, 1), nsf_hifigan.py:50-52 -- what
 * pd_plan_set_output_scale makes the sampler write); f0 [B, T] fp32 device; rand_ini [9] fp32 device: SineGen's random
 * initial phases (models.py:143-145; entry 0 must be 0); noise: optional [B, T * upp, 9] fp32 device N(0,1) draws standing
 * in for torch.randn_like (:186), NULL = in-kernel Philox keyed by `seed`; wave [B, T * upp] fp32 device (the reference
 * returns it as [B, 1, T * upp]).  lengths: optional [B] int32 device, valid frames per utterance of a padded batch (NULL = all
 * T): every activation past lengths[b] is kept at zero through the whole generator, so samples [0, lengths[b] * upp) of row b
 * are exactly what a call on that utterance alone returns (the reference runs one utterance per call) and the rest of the row
 * is zeros; whatever c and f0 hold past lengths[b] is ignored.  workspace: pd_vocoder_workspace_bytes(v, B, T) bytes,
 * 1024-aligned, used during the call. */
int pd_vocoder_forward(pd_vocoder* v, const float* c, const float* f0, const float* rand_ini, const float* noise, uint64_t seed,
                       const int* lengths, float* wave, int B, int T, void* workspace, size_t workspace_bytes, void* stream);
/* Test hook: one convolution through the generator's kernel (fp16 operands): out [B, Cout, T] = conv1d(x [B, Cin, T],
 * w [Cout, Cin, k], bias, dilation = dil, padding = (k - 1) / 2 * dil), all fp32 device tensors. */
size_t pd_voc_debug_conv_scratch(int B, int T, int Cin, int Cout, int k);
int pd_voc_debug_conv(const float* x, int B, int T, int Cin, const float* w, const float* bias, int Cout, int k, int dil, float* out,
                      void* scratch, size_t scratch_bytes, void* stream);
/* Test hook: one fused ResBlock1 round (models.py:62-69) through the resident-tile kernel used for the 16 / 32 / 64-channel
 * stages: out [B, C, T] = c2(leaky_relu(c1(leaky_relu(x, 0.1)), 0.1)) + x, c1 with dilation `dil`, all fp32 device tensors
 * (fp16 operands and storage inside). */
size_t pd_voc_debug_pair_scratch(int B, int T, int C, int k);
int pd_voc_debug_pair(const float* x, int B, int T, int C, const float* w1, const float* b1, const float* w2, const float* b2, int k,
                      int dil, float* out, void* scratch, size_t scratch_bytes, void* stream);

/* Number of kernels of this library launched by the calling thread since the last reset (bench evidence). */
long long pd_launch_count(int reset);

/* Per-kernel timing for bench.py's roofline leg: between begin and end every GEMM launch of the calling thread is
 * bracketed by a CUDA-event pair on its stream.  pd_profile_end synchronises and returns, per kernel kind, the summed
 * milliseconds and the launch count.  Kinds: 0 cond_proj, 1 input_proj, 2 dilated_conv+gate, 3 res_proj, 4 skip_sum,
 * 5 skip_proj, 6 output_proj+posterior, 7 debug. */
#define PD_N_KINDS 14  /* kind 8 = fused ResidualBlock layer kernel (2-CTA); kind 9 = fused sampler tail (skip_projection +
                          output_projection + q-posterior + next step's input_projection, pd_sample only); kind 10 = skip sum +
                          that tail as one kernel; kind 11 = text-encoder GEMMs with an fp32 / residual epilogue (in_proj,
                          out_proj, ffn_2); kind 12 = text-encoder FFN conv (ffn_1); kind 13 = vocoder convolutions (pd_vocoder_forward) */
int pd_profile_begin(void);
int pd_profile_end(float* ms_by_kind, int* count_by_kind, int n_kinds);

/* Introspection for tests: byte offset/size of a named scratch buffer inside the plan's workspace.
 * Names: "cond16","cp","xin","xa","xb","ybuf","s1","s2","steps","emb0","hid","emb","dvec","sb","flags","lens". Returns 0 if found. */
int pd_plan_buffer(const pd_plan* p, const char* name, size_t* offset, size_t* bytes);

/* Test hook: one implicit-GEMM launch of the production kernel with a plain fp32 store epilogue.
 *   a [slabs? no: B,T,Cin] 16-bit activations, w [N, taps*Cin] 16-bit weights (K-major, tap-major),
 *   out [B*T, N] fp32 = conv(a, w) + bias.  N multiple of 256, Cin multiple of 64, taps in {1,3}. */
int pd_debug_gemm(int dtype, const void* a, int B, int T, int Cin, int taps, int dil, const void* w, int N,
                  const float* bias, float* out, void* stream);

/* Select the fused 2-CTA ResidualBlock kernel (1, default; env PD_FUSED=0 also disables) or the unfused
 * conv+gate / res_proj pair (0).  Returns the previous setting (-1 = not yet decided). */
int pd_set_fused(int on);
/* With the fused kernel: run all layers of a denoiser call in ONE persistent launch whose tiles synchronise through
 * per-tile flags (1, default; env PD_STACK=0 also disables) or launch one kernel per layer (0). Returns the previous value.
 * The persistent form needs all its clusters resident at once.  The library (a) checks with the occupancy query that the
 * grid fits the device and (b) launches it COOPERATIVELY (cudaLaunchAttributeCooperative), so the driver schedules the
 * grid as a whole: kernels of other streams or processes that hold SMs delay the launch but cannot leave it half
 * resident.  If the cooperative launch is refused the call silently uses one launch per layer (no inter-cluster waits).
 * Different plans may therefore run on different streams of one GPU; they serialise at the persistent kernel (each one
 * fills the device anyway).  Every in-kernel wait still has a clock watchdog that traps instead of hanging.
 * Under an injected tool (Nsight Compute, Nsight Systems, compute-sanitizer: recognised by the environment variables they set in
 * the target process) the cooperative attribute is dropped
 * -- ncu cannot replay a cooperative cluster launch and would shut the application down -- those tools serialise kernels.
 * PD_COOP=0 / 1 forces it off / on. */
int pd_set_stack(int on);

/* pd_sample only: run skip_projection (wavenet.py:120-121), output_projection (:122), the q-posterior update
 * (prodiff.py:106-121) and the NEXT step's input_projection (wavenet.py:108-111) as one kernel whose intermediates stay in
 * shared memory (1, default when C = 256 and M = 128; env PD_TAIL=0 also disables) or as separate GEMM launches (0).
 * Returns the previous value. */
int pd_set_tail(int on);

/* pd_sample only: additionally run the deferred skip sum (wavenet.py:119) inside the same kernel as the tail above, the
 * tail chain of one frame tile overlapping the skip-sum K loop of the next (1; env PD_SUMTAIL=1 also enables; needs the tail,
 * cta_group::2 pairs and at least two frame tiles) or keep skip_sum a separate GEMM launch (0, default: the fused form is
 * bit-identical but measured no faster).  Returns the previous value. */
int pd_set_sumtail(int on);
/* NSF-HiFiGAN generator: run every ResBlock1 round (models.py:62-69) of the stages with <= 64 channels as one resident-tile
 * kernel (1, default; env PD_VOC_FUSED=0 disables) or as two convolution launches like the wider stages (0).  Returns the
 * previous value (-1 = not yet set). */
int pd_set_voc_fused(int on);

/* Debug: when set to a device buffer of PD_N_KINDS x 1024 int64 (a kind outside [0, PD_N_KINDS) is never traced), CTA 0 of every GEMM launch writes clock64() stamps of its
 * pipeline events at [kind * 1024 + tile * 16 + slot] (slots: 0/1 first/last TMA issue, 2 accumulator free, 3/4 first/
 * last operand stage landed, 5/6 epilogue before/after accumulator wait, 7 epilogue done, 14 exit, 15 entry). */
int pd_debug_set_trace(long long* device_buf);

/* Test hook: D[128,64] = A[shift:shift+128, 0:64] * W[64,64]^T with A read through a row-shifted view of one
 * 128B-swizzled shared-memory tile (mode 0: descriptor base_offset 0; mode 1: base_offset = (addr >> 7) & 7). */
int pd_debug_shift(int dtype, const void* a, int R, const void* w, int shift, int mode, float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PRODIFF_B200_H */
