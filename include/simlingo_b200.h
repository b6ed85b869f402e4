/* simlingo_b200 - C ABI of the B200 (sm_100a) kernels behind the SimLingo / InternVL2-1B hot path.
 *
 * The reference has no native layer: every op below replaces a call the reference makes into
 * PyTorch / transformers / flash-attn from simlingo_training/models (citations are
 * /root/reference/<file>:<line>; UPSTREAM = HF-Hub InternVL2-1B remote code, look-alike copies
 * in this image under vllm/model_executor/models/{intern_vit,internvl}.py and
 * transformers/models/qwen2/modeling_qwen2.py).
 *
 * Conventions
 *  - plain pointers + sizes, no C++/torch types; device pointers unless stated otherwise
 *  - caller owns every buffer (incl. workspace); the library never allocates on the hot path
 *  - all work is enqueued on `stream` (a cudaStream_t passed as void*), no implicit sync
 *  - return 0 on success, negative on error; message via slb_last_error()
 *  - tensors are row-major contiguous bf16 unless a leading dimension is given
 */
#ifndef SIMLINGO_B200_H_
#define SIMLINGO_B200_H_
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

int slb_version(void);
const char* slb_last_error(void);
int slb_num_sms(void);

/* ---- GEMM: out = epilogue(alpha * A B^T) -------------------------------------------------------
 * Replaces every nn.Linear / F.linear on the path: ViT qkv/proj/fc1/fc2 (UPSTREAM
 * intern_vit.py:237-247,279-284), mlp1 (internvl.py:644-655), Qwen2 q/k/v/o/gate/up/down and
 * lm_head (modeling_qwen2.py:35-48,161-204; llm.py:227), LoRA A/B (llm.py:106-118), patch-embed
 * Conv2d as implicit GEMM after slb_im2col_patch (intern_vit.py:59-64), and their backward
 * (dgrad/wgrad via the transposed-operand flags).
 * epilogue: v = alpha*acc; v += bias[n]; v = act(v); v *= scale_n[n]; v += residual[m,n]
 * swiglu:   B holds gate/up interleaved in groups of 128 rows; out[m, j] = silu(g_j) * u_j, N/2 cols */
enum { SLB_ACT_NONE = 0, SLB_ACT_GELU = 1, SLB_ACT_SILU = 2, SLB_ACT_RELU = 3 };
typedef struct {
  int32_t M, N, K;
  const void* A;   int64_t lda;   /* bf16 [M,K] (a_t=0) or [K,M] (a_t=1), row-major, ld in elements */
  const void* B;   int64_t ldb;   /* bf16 [N,K] (b_t=0) or [K,N] (b_t=1) */
  void* out;       int64_t ldo;   /* bf16 or fp32 [M, N] ([M, N/2] for swiglu) */
  const void* bias;               /* bf16 [N] or NULL */
  const void* scale_n;            /* bf16 [N] or NULL */
  const void* residual; int64_t ldr; /* same dtype as out, or NULL */
  float alpha;
  int32_t act;
  int32_t swiglu;
  int32_t out_fp32;
  int32_t a_t, b_t;
  int32_t block_n;                /* 0 = auto, else 128 or 256 */
  const void* rms_weight;         /* bf16 [K] or NULL: A rows are RMS-normalised on the fly, A' = A * rsqrt(mean(A^2) + rms_eps) * w */
  float rms_eps;                  /*   (Qwen2RMSNorm fused into the following projection; M <= 4 weight-streaming path only) */
  void* aux; int64_t ld_aux;      /* bf16 [M,N] or NULL.  aux_mode 1: the pre-activation alpha*acc + bias is also stored there (training
                                   *   forward of fc1: GELU input kept for backward); 2: v *= gelu'(aux[m,n]) (fc2 dgrad -> d pre-GELU) */
  int32_t aux_mode;
  const void* A2; int64_t lda2;   /* optional second A source, bf16 [M,K2]: logically appended to A along k, B then holds K + K2 columns per row.
                                   *   y = [x | t] [W | s B_lora]^T: the un-merged LoRA branch (t = A_lora x) rides in the base GEMM's k loop */
  int32_t K2;
  int32_t a_fp32;                 /* 1: A holds fp32 rows [M,K] (ld in fp32 elements): the fp32 residual stream of the Qwen2 inference path
                                   *   entering the fused RMSNorm prologue (requires rms_weight, M <= 4) */
} slb_gemm_args;
int slb_gemm_bf16(const slb_gemm_args* args, void* stream);

/* ---- norms ------------------------------------------------------------------------------------
 * nn.LayerNorm(1024, eps 1e-6) x2 per ViT layer (intern_vit.py:341-350); Qwen2RMSNorm
 * (modeling_qwen2.py:258-263).  mean/rstd (fp32 [rows]) are optional outputs for backward. */
int slb_layernorm_fwd(const void* x, const void* w, const void* b, void* y, int rows, int cols, float eps,
                      float* mean, float* rstd, void* stream);
int slb_rmsnorm_fwd(const void* x, const void* w, void* y, int rows, int cols, float eps, float* rstd, void* stream);
/* same with fp32 input rows (the Qwen2 inference path keeps its residual stream in fp32: rounding the stream to bf16 after each
 * of the 48 sub-layers is the dominant term of the bf16-vs-fp32 error of the features, profiles/r02_diag_rounding.log) */
int slb_rmsnorm_fwd_f32(const float* x, const void* w, void* y, int rows, int cols, float eps, float* rstd, void* stream);
/* the InternViT inference path does the same (profiles/r02_diag_rounding_e2e.log): fp32 residual stream out of the embedding
 * assembly, through LayerNorm (fp32 in, bf16 out) and the proj / fc2 epilogues (fp32 residual + out), into the projector front */
int slb_layernorm_fwd_f32(const float* x, const void* w, const void* b, void* y, int rows, int cols, float eps, void* stream);
int slb_vit_assemble_f32(const void* patch_out, const void* cls, const void* pos, float* x, int tiles, void* stream);
int slb_pixel_shuffle_ln_f32(const float* x, const void* w, const void* b, void* y, int tiles, float eps, void* stream);
/* ---- ViT embeddings (UPSTREAM InternVisionEmbeddings, intern_vit.py:103-115) ---------------------
 * im2col: pixels [T,3,448,448] -> patches [T*1024, kpad] (k = c*196 + py*14 + px, zero padded);
 * assemble: x[t,0] = cls + pos[0]; x[t,1+p] = patch_out[t*1024+p] + pos[1+p] */
/* Patch embedding as an implicit GEMM (UPSTREAM InternVisionEmbeddings.forward: patch_embedding conv + class token + position
 * embedding, intern_vit.py:59-64): the A operand is gathered from pixels bf16 [tiles,3,448,448] inside the kernel, weight = conv weight
 * flattened to [1024, 588] and zero-padded to ldw >= 640 columns; x [tiles*1025, 1024] (fp32 or bf16) receives
 * conv + bias + pos[1 + p] at row t*1025 + 1 + p and cls + pos[0] at row t*1025.  The explicit pair slb_im2col_patch + slb_gemm_bf16
 * (+ slb_vit_assemble) remains for the training forward, whose weight gradient needs the patch matrix. */
int slb_patch_embed(const void* pixels, const void* weight, int64_t ldw, const void* bias, const void* cls, const void* pos, void* x,
                    int tiles, int out_fp32, void* stream);
int slb_im2col_patch(const void* pixels, void* patches, int tiles, int kpad, void* stream);
int slb_vit_assemble(const void* patch_out, const void* cls, const void* pos, void* x, int tiles, void* stream);
/* ---- projector front: drop CLS + pixel_shuffle(0.5, v2) + LayerNorm(4096) fused
 * (UPSTREAM extract_feature / pixel_shuffle, internvl.py:657-684; closed form SURVEY 8a note 8) */
int slb_pixel_shuffle_ln(const void* x, const void* w, const void* b, void* y, int tiles, float eps, float* mean,
                         float* rstd, void* stream);
/* ---- attention ---------------------------------------------------------------------------------
 * ViT: bidirectional MHA, 16 heads x 64, packed qkv [T*n, 3*1024] -> out [T*n, 1024]
 * (intern_vit.py:237-247; replaces flash_attn_varlen_qkvpacked_func).  lse (fp32 [T,16,n]) optional. */
int slb_attn_vit_fwd(const void* qkv, void* out, float* lse, int tiles, int n_tokens, int heads, void* stream);
/* LLM: causal GQA (14 q heads / 2 kv heads x 64) over a KV cache (modeling_qwen2.py:161-204,
 * replaces flash_attention_2).  q [B, Lq, Hq*64] (row stride ldq), cache k/v [B, Hkv, Lmax, 64];
 * query i of the chunk sits at absolute position past+i and sees keys j <= past+i with
 * key_valid[b*key_valid_ld + j] != 0 (key_valid may be NULL).  out [B, Lq, Hq*64].
 * past_dev (nullable device int32): when given, the chunk position is read on the device instead of `past` (which then
 * only bounds the host-side checks), so that one captured CUDA graph serves every decode step; chunks of <= 32 queries. */
int slb_attn_gqa_fwd(const void* q, int64_t ldq, const void* kcache, const void* vcache, const uint8_t* key_valid,
                     int key_valid_ld, void* out, float* lse, int batch, int lq, int past, const int32_t* past_dev, int lmax,
                     int hq, int hkv, void* workspace, size_t workspace_bytes, void* stream);
/* device scratch (bytes) the prefill kernel needs when key_valid is given (128-bit validity words per key block); without it
 * masked prefill falls back to the first-generation kernel */
size_t slb_attn_gqa_fwd_workspace(int batch, int lq);

/* ---- RoPE + KV-cache write (modeling_qwen2.py:102-146 rotate_half convention, theta 1e6) --------
 * qkv [B*Lq, (Hq+2Hkv)*64] from the fused QKV GEMM; rotates q in place, writes rotated k and v
 * into the caches at positions past..past+Lq-1 (positions are arange over the padded sequence). */
/* Decode step (one new position per sequence) in one launch: qkv bf16 [batch, ldq] holds the un-rotated q | k | v of the new position;
 * the kernel rotates q (fp32) and the new key, writes the new key / value rows into the caches at position `past` (*past_dev when given)
 * and attends over positions 0 .. past.  Equivalent to slb_rope_kv_write + slb_attn_gqa_fwd with lq = 1 (llm.py:217-248 per token). */
int slb_attn_decode_rope(void* qkv, int64_t ldq, void* kcache, void* vcache, const uint8_t* key_valid, int key_valid_ld, void* out,
                         int batch, int past, const int32_t* past_dev, int lmax, int hq, int hkv, float theta, void* stream);
int slb_rope_kv_write(void* qkv, void* kcache, void* vcache, int batch, int lq, int past, const int32_t* past_dev, int lmax,
                      int hq, int hkv, float theta, void* stream);
/* ---- embeddings / placeholder substitution (adaptors.py:256; internvl2_model.py:54-91,119-131) --
 * out[b,l] = table[clamp(ids[b,l],0,V-1)]; rows with ids == img_id take vit[b*n_img + rank] where
 * rank counts <IMG_CONTEXT> ids before l; rows l in [wp_start[b], wp_start[b]+wp_len) take wp[b, l-wp_start[b]]. */
int slb_embed_assemble(const int64_t* ids, const void* table, const void* vit, const void* wp, const int32_t* wp_start,
                       int wp_len, void* out, int batch, int len, int hidden, int vocab, int img_id, int n_img,
                       void* stream);
int slb_gather_rows(const void* src, const int64_t* idx, void* dst, int n, int cols, int64_t src_rows, void* stream);
/* dst[idx[i]] = src[i] (the `inputs_embeds[selected] = vit_embeds` scatter of internvl2_model.py:124) */
int slb_scatter_rows(void* dst, const int64_t* idx, const void* src, int n, int cols, int64_t dst_rows, void* stream);

/* ---- camera-frame pre-processing (the step in front of the path: internvl2_utils.py:179-267, agent_simlingo.py:483-502)
 * uint8 RGB frames [B,3,H,W] -> PIL-compatible bicubic resize to (448*grid_h, 448*grid_w) -> grid_w*grid_h tiles of
 * 448x448 -> (u8/255 - mean)/std -> bf16 [B, tiles, 3, 448, 448].  The resampling tables (first input index, tap count,
 * 22-bit fixed-point taps [out, ksize] per output index; device pointers) are Pillow's precompute_coeffs /
 * normalize_coeffs_8bpc evaluated on the host.  tmp: uint8 [B,3,H,448*grid_w] scratch. */
typedef struct { const int32_t* first; const int32_t* count; const int32_t* taps; int32_t ksize; } slb_resample_table;
int slb_preprocess_frames(const uint8_t* frames, uint8_t* tmp, const slb_resample_table* horiz, const slb_resample_table* vert,
                          void* tiles_out, int batch, int height, int width, int grid_w, int grid_h, void* stream);

/* ---- post-processing behind the path ------------------------------------------------------------
 * What the reference computes on the host from the predicted route [B,n_route,2] / speed waypoints [B,n_wps,2] (fp32)
 * before its PID controllers run (team_code/agent_simlingo.py:915-1003, team_code/nav_planner.py:113-130):
 *   desired speed = 2 * |wps[wp_a] - wps[wp_b]|                              (agent_simlingo.py:944-946, float32)
 *   aim point     = PchipInterpolator(arc length, origin-prefixed route) sampled every sample_step metres, at index
 *                   min(int(clip(lookahead_scale * speed*3.6 + lookahead_offset, lookahead_min, lookahead_max)), M-1);
 *                   the last waypoint when the route is shorter than one step            (agent_simlingo.py:960-1003)
 *   heading error = wrap(atan2(aim.y, aim.x)) * 180 / pi / 90                            (nav_planner.py:123-130)
 * out: float64 [B,8] = {desired speed, heading error, aim.x, aim.y, M, look-ahead index, speed, 0} (one 64-byte row per
 * sample: everything the host-side PID needs in a single read-back).  The PID windows (state) stay on the host
 * (simlingo_b200/postprocess.py).  speed: fp32 [B] (m/s). */
typedef struct {
  int32_t wp_a, wp_b;
  float lookahead_scale, lookahead_offset, lookahead_min, lookahead_max;
  double sample_step;
} slb_control_params;
int slb_control_inputs(const float* route, const float* speed_wps, const float* speed, int batch, int n_route, int n_wps,
                       const slb_control_params* params, double* out, void* stream);
/* DrivingModel.equal_spacing_route (simlingo_training/models/driving.py:330-342) for a whole batch: the origin-prefixed
 * route re-sampled by linear interpolation at arc lengths 0,1,..,n_out-1 (np.interp semantics).  out: float64 [B,n_out,2] */
int slb_equal_spacing_route(const float* route, int batch, int n_route, int n_out, double* out, void* stream);

/* ---- elementwise helpers ---- */
int slb_silu_mul(const void* gate, const void* up, void* out, int64_t n, void* stream);
int slb_add_bf16(const void* a, const void* b, void* out, int64_t n, void* stream);
int slb_cast_f32_to_bf16(const float* x, void* y, int64_t n, void* stream);
/* ---- LM head tail: argmax over fp32 logits (llm.py:157-158) ---- */
int slb_argmax_f32(const float* logits, int64_t ld, int rows, int cols, int64_t* out_idx, float* out_margin,
                   void* stream);
/* the same + the bookkeeping of greedy_sample (llm.py:232-248) in the same launch: nxt[r] = argmax; column c = *pos - base of
 * sampled [rows, ld_sampled] receives it unless done[r]; n_gen[r] += !done[r]; done[r] |= (argmax == eos) (eos < 0: never); *step = c + 1 */
int slb_argmax_sample(const float* logits, int64_t ld, int rows, int cols, int64_t* nxt, int64_t* sampled, int64_t ld_sampled,
                      int max_new, const int32_t* pos, int base, uint8_t* done, int64_t* n_gen, int64_t* step, int64_t eos,
                      void* stream);
/* ---- greedy decode loop (llm.py:217-248: embed the last sampled token, all decoder layers against the KV cache, final norm,
 * lm_head, argmax, EOS bookkeeping - per generated token) as ONE persistent kernel: one CTA per SM, grid-wide barriers between the
 * phases of a layer, token loop / position counter / EOS test on the device.  Replaces, per token, the chain
 * slb_gather_rows -> n_layers x {slb_rmsnorm, slb_gemm_bf16 (qkv), slb_attn_decode_rope, slb_gemm_bf16 (o), slb_rmsnorm,
 * slb_gemm_bf16 (gate|up SwiGLU), slb_gemm_bf16 (down)} -> slb_rmsnorm -> slb_gemm_bf16 (lm_head) -> slb_argmax_f32.
 * Weights per layer (LoRA folded): qkv [(hq+2hkv)*64, hidden] + bias, o [hidden, hidden], gate|up interleaved in groups of 128 rows
 * [2*mlp, hidden] (the slb_gemm_bf16 swiglu layout), down [hidden, mlp], the two RMSNorm weights.
 * State (device): *pos = number of cached positions (= position of the next token), nxt[batch] = last sampled ids (in: to embed;
 * out: newest), sampled [batch, ld_sampled] / *step / done[batch] / n_gen[batch] as llm.py:232-248 keeps them; runs at most n_steps
 * tokens and stops early once every done[b] is set (eos >= 0).  *status is set non-zero if a grid barrier timed out. */
typedef struct { const void *qkv, *bqkv, *o, *gu, *d, *ln1, *ln2; } slb_decode_layer;
typedef struct {
  const slb_decode_layer* layers;   /* DEVICE array [n_layers] */
  int32_t n_layers, batch, hidden, mlp, vocab, hq, hkv, lmax, n_steps, max_new;
  int64_t emb_rows, eos /* < 0: none */, ld_sampled;
  const void *emb, *norm_w, *lm_head;   /* bf16 [emb_rows, hidden], [hidden], [vocab, hidden] */
  void *kcache, *vcache;                /* bf16 [n_layers, batch, hkv, lmax, 64] */
  int32_t* pos; int64_t* nxt; int64_t* sampled; int64_t* step; uint8_t* done; int64_t* n_gen;
  float rope_theta, rms_eps;
  void* workspace; size_t workspace_bytes;   /* slb_decode_workspace_bytes(), 256-byte aligned */
  uint32_t* status;
} slb_decode_args;
size_t slb_decode_workspace_bytes(int batch, int hidden, int mlp, int hq, int hkv);
int slb_decode_loop(const slb_decode_args* a, void* stream);
/* ---- driving heads (adaptors.py:113-115,130-132,163-180) and wp encoder (adaptors.py:64-93) ----
 * feats [B,30,896] -> route [B,20,2] (896->512->256->2 SiLU, cumsum), speed [B,10,2] (896->256->2), fp32 out */
typedef struct {
  const void *r0w, *r0b, *r2w, *r2b, *r4w; /* route_head.{0,2,4} */
  const void *s0w, *s0b, *s2w;             /* speed_wps_head.{0,2} */
} slb_heads_weights;
int slb_driving_heads(const void* feats, int64_t ld_batch, const slb_heads_weights* w, float* route, float* speed,
                      float* delta_ws /* fp32 [B,30,2] scratch */, int batch, void* stream);
typedef struct { const void *w0, *b0, *w2, *b2, *w4, *b4; } slb_wp_weights;
int slb_wp_encoder(const float* coords, const slb_wp_weights* w, void* out, int n_points, void* stream);

/* ================================ training-only entry points ===================================
 * The reference trains through torch autograd + flash-attn backward + torch.optim.AdamW under Lightning
 * (driving.py:236-271,718-732; train.py:160-217).  dgrad / wgrad GEMMs use slb_gemm_bf16 with b_t / a_t+b_t. */
/* dx_add (optional, bf16 [rows, cols]): the gradient arriving over the residual connection, added to dx before its one rounding */
int slb_layernorm_bwd(const void* dy, const void* x, const void* w, const float* mean, const float* rstd, void* dx,
                      float* dw_accum, float* db_accum, int rows, int cols, const void* dx_add, void* stream);
int slb_rmsnorm_bwd(const void* dy, const void* x, const void* w, const float* rstd, void* dx, float* dw_accum, int rows,
                    int cols, const void* dx_add, void* stream);
int slb_pixel_shuffle_ln_bwd(const void* dy, const void* x, const void* w, const float* mean, const float* rstd,
                             void* dx, float* dw_accum, float* db_accum, int tiles, void* stream);
int slb_gelu_fwd(const void* x, void* y, int64_t n, void* stream);
int slb_gelu_bwd(const void* pre, const void* dout, void* dpre, int64_t n, void* stream);
int slb_silu_mul_bwd(const void* gate, const void* up, const void* dout, void* dgate, void* dup, int64_t n, void* stream);
/* counter-based dropout (peft lora_dropout=0.1): y = keep(seed', i) ? x / (1-p) : 0; same call on the gradient = backward.
 * seed' = seed + (*seed_dev << 16) when seed_dev (device pointer) is given: a step counter living in device memory, so that
 * a CUDA graph replays with a fresh mask every step */
int slb_dropout(const void* x, void* y, int64_t n, float p, uint64_t seed, const uint64_t* seed_dev, void* stream);
/* y += dropout(x) with the same (seed, index) mask as slb_dropout */
int slb_dropout_add(const void* x, void* y, int64_t n, float p, uint64_t seed, const uint64_t* seed_dev, void* stream);
/* ---- un-merged LoRA training path (PEFT lora.Linear under llm.py:106-118): the rank-r products ride in the base GEMMs
 * (forward y = [x | t] [W | s B]^T through slb_gemm_args.A2; backward [dx | dt] = dy [W | s B], one dgrad over the concatenated
 * weight); these are the HBM-bound passes around them.
 * slb_dropout_multi: n_out (<= 4) independently masked copies of x in one read; copy j equals slb_dropout(x, seeds[j]).
 * slb_lora_pack: table_dev = int64 [n_entries][4] on the device {src bf16 [rows, rank] contiguous, dst bf16, rows, dst row
 *   stride in elements}; dst[row, 0:rank] = scale * src[row, :] for every entry, one launch.
 * slb_lora_dx: out[M, K] = in[:, 0:K] + sum_j mask_j o (in[:, K + rank*j : K + rank*(j+1)] @ A[j]) / (1 - p), A[j] bf16 [rank, K]
 *   contiguous, rank a multiple of 16 and <= 32; mask_j = the keep mask of slb_dropout(seed = seeds[j]) over the contiguous [M, K] index space (seeds == NULL or
 *   p == 0: no mask); fp32 accumulation, one rounding.  A / seeds are HOST arrays of n_adapters (<= 4) entries.
 * slb_silu_mul_cat(_bwd): SwiGLU on gate_up bf16 [rows, 2*inter] = [gate | up]: out = silu(gate) * up; dgate_up = [dgate | dup]. */
int slb_dropout_multi(const void* x, void* const* ys, const uint64_t* seeds, int n_out, int64_t n, float p,
                      const uint64_t* seed_dev, void* stream);
int slb_lora_pack(const int64_t* table_dev, int n_entries, int rank, float scale, void* stream);
/* slb_lora_wgrad_grouped: the adapters' parameter gradients of one group of linears in ONE launch (PEFT lora.Linear backward:
 * dB = s dy^T t, dA = dt^T dropout(x); llm.py:106-118 wraps q/k/v/o/gate/up/down):  out [mo, no] (+)= alpha * P^T Q  with
 * P bf16 [rows, mo] (row stride ldp), Q bf16 [rows, no] (ldq), out bf16 (ldo), mo / no multiples of 32, <= 16 problems (HOST array). */
typedef struct { const void* P; const void* Q; void* out; int64_t ldp, ldq, ldo; int32_t mo, no; float alpha; int32_t accumulate; } slb_wgrad_problem;
int slb_lora_wgrad_grouped(const slb_wgrad_problem* probs, int n, int rows, void* stream);
int slb_lora_dx(const void* in, int64_t ld_in, void* out, int64_t ld_out, const void* const* A, const uint64_t* seeds,
                int n_adapters, int M, int K, int rank, float p, const uint64_t* seed_dev, void* stream);
int slb_silu_mul_cat(const void* gate_up, void* out, int rows, int inter, void* stream);
int slb_silu_mul_cat_bwd(const void* gate_up, const void* dout, void* dgate_up, int rows, int inter, void* stream);
/* y (bf16) = (accumulate ? y : 0) + x (fp32): flush of fp32 small-parameter gradient accumulators */
int slb_flush_f32_to_bf16(const float* x, void* y, int64_t n, int accumulate, void* stream);
int slb_add_inplace_bf16(void* a, const void* b, int64_t n, void* stream);
int slb_scale_cols(const void* x, const void* s, void* out, int rows, int cols, void* stream);
/* out = res + s[col] * x  (layer-scale residual of InternViT when the branch output must be kept for backward) */
int slb_scale_cols_add(const void* x, const void* s, const void* res, void* out, int rows, int cols, void* stream);
/* acc[c] += alpha * sum_r a[r,c] * (b ? b[r,c] : 1)   (bias / layer-scale / norm-weight gradients) */
int slb_col_reduce(const void* a, int64_t lda, const void* b, int64_t ldb, float* acc, int rows, int cols, float alpha,
                   void* stream);
/* layer-scale residual backward in one pass: dbranch = dx * ls;  dls += colsum(dx * branch);  dbias += colsum(dbranch) */
int slb_layerscale_bwd(const void* dx, const void* branch, const void* ls, void* dbranch, float* dls_accum, float* dbias_accum,
                       int rows, int cols, void* stream);
int slb_vit_assemble_bwd(const void* dx, void* dpatch_out, float* dcls_accum, float* dpos_accum, int tiles, void* stream);
/* packs attention gradients (dq fp32 [B*L, Hq*64], dk/dv fp32 [B, Hkv, L, 64]) into the fused-QKV gradient
 * [B*L, (Hq+2Hkv)*64] bf16, rotating dq/dk back (theta <= 1: no rotation, the ViT case) */
int slb_rope_bwd(const float* dq, const float* dk, const float* dv, void* dqkv, int batch, int lq, int hq, int hkv,
                 float theta, void* stream);
int slb_attn_delta(const void* o, const void* dout, float* delta, int batch, int lq, int heads, void* stream);
/* Flash-attention backward.  Every key block writes its partial dQ tile into its own bf16 slab of `workspace`
 * (slb_attn_bwd_workspace bytes); a reduction kernel sums the slabs in fp32.
 * ViT: dqkv (bf16 [T*n, 3*H*64], same packing as qkv) receives dQ | dK | dV directly.
 * GQA: dq fp32 [B*L, Hq*64] (post-RoPE space, overwritten), dk / dv fp32 [B, Hkv, L, 64]; slb_rope_bwd rotates back. */
size_t slb_attn_bwd_workspace(int batch, int lq, int hq);
int slb_attn_vit_bwd(const void* qkv, const void* dout, const float* lse, const float* delta, void* dqkv, void* workspace,
                     size_t workspace_bytes, int tiles, int n_tokens, int heads, void* stream);
int slb_attn_gqa_bwd(const void* q, int64_t ldq, const void* kcache, const void* vcache, const uint8_t* key_valid,
                     int key_valid_ld, const void* dout, const float* lse, const float* delta, float* dq, float* dk,
                     float* dv, void* workspace, size_t workspace_bytes, int batch, int lq, int lmax, int hq, int hkv,
                     void* stream);
/* fused softmax cross-entropy over fp32 logits rows (adaptors.py:271-273): loss[r] = lse - logit[label];
 * dlogits (bf16, row stride ldd >= cols, zero padded) = (softmax - onehot) * grad_scale; label < 0 => ignored */
int slb_ce_fwd_bwd(const float* logits, int64_t ld, const int64_t* labels, float* loss, void* dlogits, int64_t ldd,
                   float grad_scale, int rows, int cols, void* stream);
/* fused multi-tensor AdamW + global-norm clip over flat buffers (torch.optim.AdamW at driving.py:718-724, Trainer(gradient_clip_val)
 * at train.py:206): fp32 master parameters / moments, bf16 gradients in, bf16 model copy written back in the same pass;
 * the gradient is scaled by grad_prescale (1/world under data parallelism) and by min(1, max_norm / (|g| + 1e-6)) */
int slb_grad_sqnorm(const void* grad_bf16, int64_t n, float* out_sq, void* stream);
int slb_adamw_fused(float* master, float* m, float* v, const void* grad_bf16, void* param_bf16, int64_t n, float lr,
                    float beta1, float beta2, float eps, float wd, int step, const float* grad_sqnorm, float max_norm,
                    float grad_prescale, void* stream);

/* ---- data-parallel gradient exchange (Lightning DDP / DeepSpeed ZeRO-2 in the reference: train.py:160-168,
 * train_simlingo_seed1.sh:27) -------------------------------------------------------------------------------------------
 * One NCCL communicator per process (= per GPU), created and destroyed explicitly; NCCL is resolved at run time
 * (libnccl.so.2), so nothing here is needed on a single GPU.  The 128-byte unique id comes from slb_comm_unique_id on one
 * rank and reaches the others through the host (simlingo_b200/dist.py broadcasts it with torch.distributed).
 * slb_allreduce_bucket: in-place all-reduce of `n` elements (dtype 0 = bf16, 1 = fp32; average 0 = SUM, 1 = mean) enqueued
 * on `stream`; the training runtime issues one per finished gradient bucket on its own communication stream while the
 * backward pass continues on the compute stream.  slb_comm_version: NCCL version code, 0 when NCCL cannot be loaded. */
int slb_comm_version(void);
int slb_comm_unique_id(void* out_128_bytes /* host */);
int slb_comm_init(void** comm_out, const void* unique_id_128_bytes /* host */, int rank, int world);
int slb_comm_destroy(void* comm);
int slb_allreduce_bucket(void* comm, void* buf, int64_t n, int dtype, int average, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SIMLINGO_B200_H_ */
