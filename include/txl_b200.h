/* txl_b200.h — C ABI of the B200-native Transformer-XL hot path.
 *
 * The reference (StefanHeng/Symbolic-Music-Generation) has no FFI: its boundary for this path is the
 * Python class musicnlp/models/transformer_xl.py:127-241 (MyTransfoXLLMHeadModel) whose arithmetic
 * is HF transformers==4.25.1 `modeling_transfo_xl.py` (un-vendored).  Each entry
 * point below replaces one torch-op site of that module; the HF site it replaces is cited as
 * [A.x] = SURVEY.md Appendix A section, plus the reference call site (file:line) that reaches it.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in `_host`;
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing synchronises;
 *   - return 0 on success, negative TXL_E* on error; `txl_last_error()` gives a thread-local message;
 *   - no global mutable state except a per-process cache of TMA descriptors/func attributes;
 *   - sm_100a only.  There is NO CPU fallback: without a B200 every compute entry point fails.
 *   - activations are BATCH-MAJOR: row = b*T + t (HF is time-major inside; the Python boundary converts `mems`).
 *   - dtype codes: TXL_F32 = 0, TXL_BF16 = 1.  Biases, LayerNorm affine, r_w_bias/r_r_bias, statistics,
 *     losses and gradient accumulators of parameters are always fp32.
 */
#ifndef TXL_B200_H
#define TXL_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TXL_F32 0
#define TXL_BF16 1

#define TXL_OK 0
#define TXL_EINVAL (-1)   /* bad argument / unsupported shape */
#define TXL_ECUDA (-2)    /* CUDA runtime error, see txl_last_error() */
#define TXL_ENODEV (-3)   /* no sm_100 device */

/* Integer geometry of one relative-position attention call  [A.2 step 3-5, A.4, A.5]. */
typedef struct {
  int T;           /* qlen: query rows of this segment                               */
  int mlen;        /* rows of cached memory actually present (keys 0..mlen-1)        */
  int mem_len;     /* config.mem_len (defines the same_length band)                  */
  int clamp_len;   /* config.clamp_len (<=0: no clamp)                               */
  int same_length; /* config.same_length                                             */
} TxlBand;

int txl_version(void);
const char* txl_last_error(void);
/* 0 if the current device is sm_100 (B200); TXL_ENODEV otherwise. */
int txl_device_ok(void);
/* number of kernels this library has launched in this process (bench.py's `gpu_launches`). */
unsigned long long txl_launch_count(void);

/* ---- integer index maps (bit-exact contract) -------------------------------------------------
 * masked[i*klen+j] = 1 iff key j is masked for query i          (HF uint8 triu+tril mask, [A.2-4])
 * ridx[i*klen+j]   = relative-position row used by BD at (i,j)  (HF _rel_shift + clamp, [A.4]); -1 where masked
 * lo[i], hi[i]     = first / last live key of query i (inclusive)
 * Computed on the device with the same inline functions the attention kernels use. */
int txl_relattn_index_map(const TxlBand* band, uint8_t* masked, int32_t* ridx, int32_t* lo, int32_t* hi, void* stream);

/* ---- embedding  [A.6 AdaptiveEmbedding, div_val=1; call site transformer_xl.py:163] -------------
 * out[n,:] = E[ids[n],:] * scale, then inverted dropout (p, seed, site) if p>0. */
int txl_embed_fwd(const int64_t* ids, const void* E, void* out, int64_t n_tok, int d, int V, float scale,
                  int dtype, float drop_p, uint64_t seed, uint32_t site, void* stream);
/* dE[ids[n],:] += (dOut[n,:] + dOut2[n,:]) * scale (* dropout mask)   (fp32 atomics; dOut2 may be NULL) */
int txl_embed_bwd(const int64_t* ids, const void* dOut, const void* dOut2, float* dE, int64_t n_tok, int d, int V, float scale,
                  int dtype, float drop_p, uint64_t seed, uint32_t site, void* stream);

/* ---- sinusoid table  [A.2 step 5 pos_seq + A.6 PositionalEmbedding] --------------------------------
 * HF's `pos_emb` literally: row x (0..klen-1) holds position pos = min(klen-1-x, clamp_len) (no clamp if clamp_len<=0):
 * out[x,:] = [sin(pos*f_0..), cos(pos*f_0..)], f_k = 10000^(-2k/d); inverted dropout if drop_p>0 */
int txl_posemb_table(void* out, int klen, int clamp_len, int d, int dtype, float drop_p, uint64_t seed, uint32_t site, void* stream);

/* ---- generic GEMM with epilogue  [A.3 qkv_net/r_net/o_net, A.6 CoreNet.0/3, crit.out_layers] -----
 * C[M,N] = epi( op(A)[M,K] * op(B)[K,N] ), row-major, leading dims in elements.
 *   transA=0: A is [M,K] (lda>=K); transA=1: A is stored [K,M] (lda>=M)
 *   transB=0: B is [K,N] (ldb>=N); transB=1: B is stored [N,K] (ldb>=K)   (nn.Linear weight => transB=1)
 * epilogue: v = acc; if bias: v += bias[n]; if (flags&RELU) v = max(v,0); if (flags&MASK_POS) v *= (aux[m,n]>0);
 *           if (flags&MASK_LIVE) v *= bit(live_bits, m, n)   (the same mask at one bit per element, written by the forward GEMM's EMIT_LIVE);
 *           if (flags&MASK_SCALE) v *= 1/(1-drop_p)   (aux is a post-dropout activation: its zeros already are the dropout mask,
 *                                                       so the backward of relu+dropout needs no second hash pass);
 *           if (flags&DROPOUT) v = inverted-dropout(v); if (flags&ACCUM) v += C[m,n];  C = (dtype_c) v
 * colsum (optional, fp32[N]): colsum[n] += sum_m v (before ACCUM) — bias gradients.
 * bf16 inputs with all of M,N,K and leading dims "nice" run on tcgen05 tensor cores; everything else on a
 * SIMT fp32-FMA kernel (true fp32 accumulate — the fp32 parity mode). */
#define TXL_EPI_RELU 1
#define TXL_EPI_ACCUM 2
#define TXL_EPI_MASK_POS 4
#define TXL_EPI_DROPOUT 8
#define TXL_EPI_BIAS_ROW 16   /* bias is indexed by the output ROW m (used with TRANSPOSE: y^T = W x^T + b) */
#define TXL_EPI_MASK_SCALE 64 /* with MASK_POS: also scale by 1/(1-drop_p) (aux = post-dropout activation) */
#define TXL_EPI_EMIT_LIVE 128  /* also WRITE live_bits: bit = (final v > 0) — the backward mask of relu(+dropout) at 1 bit per element */
#define TXL_EPI_MASK_LIVE 256  /* v *= live_bits bit (instead of MASK_POS's aux > 0); combines with MASK_SCALE */
#define TXL_EPI_TRANSPOSE 32  /* store C transposed: C[n*ldc + m]  (decode: features are the GEMM's M so 128-row MMA tiles stay full) */
typedef struct {
  const float* bias;     /* [N] or NULL */
  const void* aux;       /* [M,N] same dtype/ld as C, for MASK_POS */
  float* colsum;         /* [N] fp32 accumulate or NULL */
  float drop_p; uint64_t seed; uint32_t site;
  int flags;
  uint32_t* live_bits;   /* [ceil(N/32), M] words, bit n%32 of word [n/32][m] <-> element (m,n); EMIT_LIVE overwrites every word, MASK_LIVE reads */
} TxlEpilogue;
int txl_gemm(const void* A, const void* B, void* C, int64_t M, int64_t N, int64_t K,
             int64_t lda, int64_t ldb, int64_t ldc, int transA, int transB,
             int dtype_ab, int dtype_c, const TxlEpilogue* epi, void* stream);

/* ---- residual + LayerNorm  [A.3 step 8, A.6 PositionwiseFF] ---------------------------------------
 * z = x + dropout(r); y = LN(z)*gamma + beta;   mean/rstd fp32 [rows] saved for backward; z optional. */
int txl_add_ln_fwd(const void* x, const void* r, const float* gamma, const float* beta, void* y, void* z,
                   float* mean, float* rstd, int64_t rows, int d, float eps, int dtype,
                   float drop_p, uint64_t seed, uint32_t site, void* stream);
/* Linear + dropout + residual + LayerNorm in one tensor-core kernel (bf16): y = LN(resid + dropout(A W^T + bias)) * gamma + beta, the
 * o_net -> layer_norm tail of RelPartialLearnableMultiHeadAttn and the CoreNet.3 -> layer_norm tail of PositionwiseFF  [A.3 step 8, A.6].
 * A [M, K] (row pitch lda), W [N, K] (row pitch ldw), resid / y / z [M, N] contiguous; the accumulator is not rounded before the add.
 * z (the LayerNorm input), mean, rstd: what txl_add_ln_bwd reads — all three or none (evaluation).  Same dropout site indexing as
 * txl_add_ln_fwd.  *handled = 0 (and nothing launched) when the shape is not covered (N not 256 / 512, M < 256, fp32 mode, TXL_GEMM_LN=0):
 * the caller then runs txl_gemm + txl_add_ln_fwd. */
int txl_gemm_add_ln_fwd(const void* A, const void* W, const float* bias, const void* resid, const float* gamma, const float* beta,
                        void* y, void* z, float* mean, float* rstd, int64_t M, int64_t N, int64_t K, int64_t lda, int64_t ldw,
                        float eps, float drop_p, uint64_t seed, uint32_t site, void* stream, int* handled);
/* dyt = dy (+ dy2 if non-NULL: the two branches that meet at a residual node);  dz = LN'(dyt);
 * dgamma += sum dyt*xhat; dbeta += sum dyt.   dx_out = dz (+ dx_out if accumulate);  dr_out = dz * dropout-mask.  dx_out may alias dy. */
int txl_add_ln_bwd(const void* dy, const void* dy2, const void* z, const float* gamma, const float* mean, const float* rstd,
                   void* dx_out, int accumulate_dx, void* dr_out, float* dgamma, float* dbeta,
                   int64_t rows, int d, int dtype, float drop_p, uint64_t seed, uint32_t site, void* stream);
/* out[n] += sum_m X[m,n]  (fp32 accumulate; bias gradients of CoreNet.3 / crit.out_layers.0) */
int txl_colsum(const void* X, int64_t M, int64_t N, int64_t ldx, int dtype, float* out, void* stream);
/* y = dropout(x) (final `drop(core_out)`, [A.2-8]); backward is the same op on dy. */
int txl_dropout(const void* x, void* y, int64_t n, int dtype, float drop_p, uint64_t seed, uint32_t site, void* stream);

/* ---- relative-position attention  [A.3 steps 2-8 + A.4 _rel_shift + A.5 band; RelPartialLearnableMultiHeadAttn]
 * q:      [B, T, H*dh] view with row stride ldq (elements)
 * k/v:    keys 0..mlen-1 from k_mem/v_mem ([B, mlen, H*dh], row stride ldkv_mem),
 *         keys mlen..mlen+T-1 from k_cur/v_cur ([B, T, H*dh], row stride ldkv_cur)
 * r:      [klen, H*dh] = HF's r_head_k = r_net(pos_emb): row x encodes relative distance min(klen-1-x, clamp).
 *         Query i / key j use row x = (T-1-i) + j, which IS the pad/reshape `_rel_shift` (BD[i,j] = BD0[i, j+T-1-i], [A.4])
 * rwb/rrb:[H*dh] fp32 (r_w_bias, r_r_bias)
 * out:    [B, T, H*dh] (ld = H*dh);  lse: [B, H, T] fp32 log-sum-exp of the scaled scores
 * score(i,j) = ((q_i+rwb).k_j + (q_i+rrb).r[T-1-i+j]) / sqrt(dh) on the live band only. */
typedef struct {
  int B, H, dh;
  TxlBand band;
  int64_t ldq, ldkv_mem, ldkv_cur;
  int dtype;
} TxlAttnDims;
/* saved (optional, may be NULL): txl_relattn_saved_bytes(dims) bytes of forward state the tensor-core backward reuses instead of
 * recomputing the scores — the bf16 soft-max numerators of every live band tile and the running row maxima behind them (what
 * autograd keeps as `attn_prob` in HF, at half the size).  txl_relattn_saved_bytes returns 0 when no kernel would use it
 * (fp32 mode, odd shapes): pass NULL then.  Passing a buffer the selected forward kernel cannot fill is an error, never ignored. */
int64_t txl_relattn_saved_bytes(const TxlAttnDims* dims);
int txl_relattn_fwd(const void* q, const void* k_mem, const void* v_mem, const void* k_cur, const void* v_cur,
                    const void* r, const float* rwb, const float* rrb, void* out, float* lse, void* saved,
                    const TxlAttnDims* dims, void* stream);
/* Backward.  dq/dk_cur/dv_cur are written with the same strides as their forward tensors (they may be
 * slices of one [B,T,3d] buffer); dk_mem/dv_mem may be NULL (mems detached and all-zero => no wgrad term).
 * dr [klen,H*dh], drwb/drrb [H*dh] are fp32 and ACCUMULATED into.  ws: workspace of txl_relattn_bwd_workspace bytes.
 * saved: NULL, or the buffer the forward call with the same inputs and dims filled (read-only here). */
/* profiling aid (bench.py's dominant-kernel probe): dbg 16 = the tensor-core backward runs its dQ pass alone, 32 = skips the prep kernel;
 * abl = ablation bits of the dQ pass.  0, 0 = normal operation.  Returns the previous dbg.  Never set on the product path. */
int txl_relattn_bwd_probe(int dbg, int abl);
int64_t txl_relattn_bwd_workspace(const TxlAttnDims* dims);
int txl_relattn_bwd(const void* q, const void* k_mem, const void* v_mem, const void* k_cur, const void* v_cur,
                    const void* r, const float* rwb, const float* rrb, const void* out, const float* lse,
                    const void* dout, void* dq, void* dk_mem, void* dv_mem, void* dk_cur, void* dv_cur,
                    float* dr, float* drwb, float* drrb, void* ws, const void* saved, const TxlAttnDims* dims, void* stream);

/* ---- LM head: log-softmax + NLL  [A.6 ProjectedAdaptiveLogSoftmax n_clusters=0; transformer_xl.py:185,193]
 * logits [N, ldl] (dtype) are the raw h.E^T+b produced by txl_gemm; labels int64 [N] (-100 = ignore).
 * losses[n] = lse - logit[label] (0 where ignored); lse[n] saved; if logprobs != NULL: logprobs[n,v] = logit - lse (fp32, ld=V)
 * argmax (optional int64 [N]) = argmax_v logit  (preprocess_logits_for_metrics, train.py:248-252). */
int txl_logsoftmax_nll_fwd(const void* logits, int64_t ldl, const int64_t* labels, float* losses, float* lse,
                           float* logprobs, int64_t* argmax, int64_t N, int V, int dtype, void* stream);
/* dlogits[n,v] = (softmax - onehot(label)) * grow[n]  (0 rows where label ignored; pad columns v>=V zeroed).  dlogits may alias logits
 * (same dtype/pitch) or be a separate buffer of another dtype (fp32 logits -> bf16 dlogits for the tensor-core dgrad/wgrad GEMMs). */
int txl_logsoftmax_nll_bwd(const void* logits, int64_t ldl, int dtype, void* dlogits, int64_t ldd, int dtype_out, const int64_t* labels,
                           const float* lse, const float* grow, int64_t N, int V, void* stream);
/* loss = mean(losses[losses != 0]) and grow[n] = g_loss/cnt * (losses[n]!=0) + g_losses[n]  (transformer_xl.py:197-200) */
/* Adaptive softmax, cluster path (HF ProjectedAdaptiveLogSoftmax with n_clusters > 0, div_val 1: the reference's DEFAULT criterion for
 * vocabularies >= 1000, musicnlp/models/transformer_xl.py:56-66).  A logits row = V token logits followed by the n_clusters cluster logits
 * (one GEMM over [embedding ; crit.cluster_weight]); cutoffs = HF's `cutoffs` (increasing, inside (0, V)), n_clusters <= 4.
 * fwd: lse [N, 1 + n_clusters] (head, tails); losses[n] = -log p(labels[n]) (0 where the label is -100); logprobs [N, V] (optional) =
 *      head log-softmax for v < cutoffs[0], else head log-prob of the tail's cluster + tail log-softmax; argmax (optional) of those.
 * bwd: dlogits [N, ldd >= V + n_clusters] = grow[n] * d losses[n] / d logits.
 * txl_pack_losses: HF's keep_order=False ordering of the returned loss vector (positions of cluster 0 first, then cluster 1, ..., ignored
 *      labels as trailing zeros): pos_losses [B, T] -> packed [B*(T-1)], perm[k] = b*T + t of packed entry k (-1: none). */
int txl_adaptive_lsm_nll_fwd(const void* logits, int64_t ldl, const int64_t* labels, float* losses, float* lse, float* logprobs,
                             int64_t* argmax, int64_t N, int V, int n_clusters, const int* cutoffs, int dtype, void* stream);
int txl_adaptive_lsm_nll_bwd(const void* logits, int64_t ldl, int dtype, void* dlogits, int64_t ldd, int dtype_out, const int64_t* labels,
                             const float* lse, const float* grow, int64_t N, int V, int n_clusters, const int* cutoffs, void* stream);
int txl_pack_losses(const float* pos_losses, const int64_t* labels_shift, int B, int T, int V, int n_clusters, const int* cutoffs,
                    float* packed, int64_t* perm, void* stream);
int txl_masked_mean(const float* losses, int64_t N, float* loss_out, float* count_out, void* stream);
/* Next-token-prediction accuracy counts  [reference train_util_wrap.py:113-120, train.py:279-284; SURVEY §8f-2]
 * preds  [B, T] int64: greedy prediction made AT each position (the fused argmax of txl_logsoftmax_nll_fwd), row stride ld_preds
 * labels [B, T] int64 (-100 = pad).  Position t predicts token t+1:  out[0] += #{(b,t<T-1): labels[b,t+1] != -100 and preds[b,t] == labels[b,t+1]},
 * out[1] += #{labels[b,t+1] != -100}.  out is int64[2] on the device (accumulated: zero it per logging window). */
int txl_ntp_acc(const int64_t* preds, int64_t ld_preds, const int64_t* labels, int64_t ld_labels, int B, int T, int64_t* out, void* stream);

/* ---- formats either side of the path ------------------------------------------------------
 * txl_clm_labels: HF DataCollatorForLanguageModeling(mlm=False) as the reference builds it (musicnlp/trainer/train.py:360): labels = input_ids
 *   with every pad id replaced by -100 (the tokenizer already padded to max_length, musicnlp/preprocess/dataset.py:361).  n int64 elements.
 * txl_last_index_of: out[b] = last t with ids[b, t] == token, -1 if none — the cut point of MusicGenerator._truncate_last_bar
 *   (musicnlp/trainer/eval.py:178-185: ids[:last start-of-bar]). */
int txl_clm_labels(const int64_t* ids, int64_t* labels, int64_t n, int64_t pad_id, void* stream);
/* labels [B, T] int64 (row stride ld) -> shift [B*T]: shift[b*T+t] = labels[b, t+1], -100 in the last column (HF crit's `labels[..., 1:]`,
 * A.6), after the reference's all-pad first-row fix-up applied IN PLACE on the device with its own arithmetic
 * (musicnlp/models/transformer_xl.py:176-182: sum(labels[0,1:]) == (T-1)*-100  =>  labels[0,1] = eos).  bad (optional, int32, accumulated):
 * number of labels outside [0, V) that are not -100. */
int txl_shift_labels(int64_t* labels, int64_t ld, int B, int T, int64_t eos, int V, int64_t* shift, int* bad, void* stream);
int txl_last_index_of(const int64_t* ids, int64_t ld, int B, int T, int64_t token, int64_t* out, void* stream);

/* ---- parameters --------------------------------------------------------------------------------- */
int txl_cast_f32_to_bf16(const float* src, void* dst, int64_t n, void* stream);
int txl_cast_bf16_to_f32(const void* src, float* dst, int64_t n, void* stream);
/* B[c,r] = A[r,c]  (weight transposes for dgrad operands) */
int txl_transpose(const void* A, void* B, int64_t rows, int64_t cols, int dtype, void* stream);
/* fused AdamW over a flat fp32 buffer (train.py:166-190 defaults): decay applied where decay_mask[i]!=0 (NULL: all).
 * grad_scale multiplies g first (global-norm clip factor, read from device: *grad_scale_dev if non-NULL). */
int txl_adamw_step(float* p, const float* g, float* m, float* v, const uint8_t* decay_mask, int64_t n,
                   float lr, float beta1, float beta2, float eps, float weight_decay, int step,
                   const float* grad_scale_dev, void* bf16_shadow, void* stream);
/* out[0] += sum g^2 (fp32 atomics) */
int txl_sumsq(const float* g, int64_t n, float* out, void* stream);

/* ---- sampling  [A.7; HF logits warpers as used at eval.py:277-333] -------------------------------
 * scores [B, V] fp32 log-probs of the last position.  Temperature -> TopK (ties kept) -> TopP (first token
 * crossing p kept) -> renormalise -> draw.  u [B] uniform(0,1) supplied by the caller (so tests can pin draws);
 * do_sample=0 => argmax.  next [B] int64.  keep (optional uint8 [B,V]) = surviving set, warped (optional fp32
 * [B,V]) = renormalised log-probs (-inf outside keep). */
int txl_sample(const float* scores, int B, int V, int do_sample, float temperature, int top_k, float top_p,
               const float* u, int64_t* next, uint8_t* keep, float* warped, void* stream);

/* ---- decode step (batched generation; [A.3-A.5] at T=1, [A.7], [A.8'])  -----------------------------------
 * generate() keeps a private ring cache of PROJECTED keys/values kc, vc [B, H, mem_len, d_head] (HF re-projects all cached hidden
 * states every step); API-level `mems` stay hidden states.  `pos` is a device int32 = tokens appended so far (slot = pos % mem_len).
 * txl_decode_cache_init: kv_mem [B*mem_len, 2*H*dh] (k|v, from txl_gemm of the hidden-state mems) -> kc, vc.
 * txl_decode_attn: qkv [B, 3*H*dh]; appends k,v at the ring slot (overwriting the oldest entry: for T=1, mlen==mem_len the live band IS
 *   the ring after the write), then out[b,:] = softmax_s(((q+rwb).k_s + (q+rrb).r[mem_len - dist_s]) / sqrt(dh)) . v_s,
 *   dist_s = (slot - s) mod mem_len;  r [mem_len+1, H*dh] = r_head_k for klen = mem_len+1.
 * txl_skinny_gemm: C[M<=64, N] = A[M,K] W[N,K]^T (+bias)(ReLU), weight-streaming GEMM for the per-step Linears.
 * txl_decode_uniform: u[b] = U(0,1) keyed on (seed, seq_offset + b, *pos) — reproducible under any sharding of sequences over GPUs.
 * txl_decode_commit: HF sample()/greedy_search() bookkeeping on the device: finished rows emit pad, eos finishes a row, token stored at
 *   out_ids[b, col0 + *pos] and fed back in tok[b]; then *pos += 1. */
int txl_decode_cache_init(const void* kv_mem, int64_t ld, void* kc, void* vc, int B, int H, int mem_len, int dh, int dtype, void* stream);
int txl_decode_attn(const void* qkv, void* kc, void* vc, const void* r, const float* rwb, const float* rrb, void* out, const int32_t* pos,
                    int B, int H, int mem_len, int dh, int dtype, void* stream);
int txl_skinny_gemm(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias, void* C, int64_t ldc, int M, int N, int K,
                    int relu, int dtype, void* stream);
int txl_decode_uniform(float* u, int B, uint64_t seed, int64_t seq_offset, const int32_t* pos, void* stream);
int txl_decode_commit(const int64_t* next, int64_t* tok, int64_t* unfinished, int64_t* out_ids, int64_t ld_out, int col0, int32_t* pos, int B,
                      int64_t eos, int64_t pad, int use_eos, void* stream);

/* Second-generation bf16 decode kernels (csrc/decode_stream.cu): what the CUDA-graph step launches.
 * txl_dec_linear: C[M<=64, N] = A[M,K] W[N,K]^T (+bias)(ReLU), bf16 operands, fp32 accumulation; 8 or 16 output features per CTA so that
 *   64-150 CTAs stream the weight matrix through a 4-stage cp.async ring, mma.sync fragments read as 16-byte pieces.  K % 32 == 0.  C is
 *   bf16, or fp32 when out_f32 (the LM-head logits).  splits > 1: split-K, C = fp32 planes [splits][M][ldc] of partial sums (no bias / ReLU),
 *   to be summed by txl_dec_add_ln.  Replaces every nn.Linear of the T=1 step [A.3, A.6].
 * txl_dec_add_ln: y[M, d] (bf16) = LayerNorm(x + sum of `nparts` fp32 planes [M][d] + bias) * gamma + beta — the residual LayerNorm after
 *   o_net / CoreNet.3 [A.3-8, A.6] fused with the split-K reduction and the bias.
 *   Both kernels optionally issue an L2 prefetch (cp.async.bulk.prefetch.L2) of [prefetch, prefetch + prefetch_bytes), spread over their CTAs:
 *   the decode step hands them slices of the NEXT attention kernel's ring, which HBM can deliver while these latency-bound kernels run.
 * txl_decode_rtab_head_major: r [mem_len+1, H*dh] -> [H, mem_len+1, dh] (per-head rows contiguous for the bulk copies below).
 * txl_decode_cache_init_kv: kv_mem [B*mem_len, 2*H*dh] -> ring cache kvc [B, H, mem_len, 2*dh] (a key's k row followed by its v row: a stage
 *   of keys is one contiguous run).
 * txl_decode_attn_pipe: the contract of txl_decode_attn (bf16 only) over the interleaved ring kvc and the head-major r; k|v rows and r rows
 *   stream through a multi-stage shared-memory ring filled by cp.async.bulk (mbarrier full/empty), 2-4 CTAs per SM.  splits > 1: the ring
 *   is cut into `splits` ranges handled by separate CTAs (for few sequences per GPU), merged by the last CTA to arrive; needs ws of
 *   txl_decode_attn_pipe_ws_bytes and int counters[B*H] zeroed once (the kernel leaves them zero).
 * txl_set_pdl(1): launch the kernels above with programmatic stream serialization (each executes griddepcontrol.wait before touching what a
 *   predecessor produced), so launch latency, barrier set-up and weight / ring prefetch overlap the previous kernel's tail; returns the old value. */
int txl_dec_linear(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias, void* C, int64_t ldc, int M, int N, int K,
                   int relu, int out_f32, int splits, const void* prefetch, int64_t prefetch_bytes, void* stream);
int txl_dec_add_ln(const void* x, const float* part, int nparts, const float* bias, const float* gamma, const float* beta, void* y, int M, int d,
                   float eps, const void* prefetch, int64_t prefetch_bytes, void* stream);
int txl_decode_rtab_head_major(const void* r, void* out, int rows, int H, int dh, void* stream);
int64_t txl_decode_attn_pipe_ws_bytes(int B, int H, int dh, int splits);
int txl_decode_cache_init_kv(const void* kv_mem, int64_t ld, void* kvc, int B, int H, int mem_len, int dh, void* stream);
int txl_decode_attn_pipe(const void* qkv, void* kvc, const void* r_head_major, const float* rwb, const float* rrb, void* out,
                         const int32_t* pos, int B, int H, int mem_len, int dh, int splits, void* ws, int* counters, void* stream);
/* stage geometry of txl_decode_attn_pipe (keys per stage x stages at d_head 64): 0 = 32 x 4, 1 = 64 x 3, 2 = 64 x 4, 3 = 128 x 2 (default), 4 = 128 x 4,
 * 5 = 64 x 2; -1 = re-read the TXL_DECODE_ATTN_CFG environment variable at the next call.  Returns the previous setting. */
int txl_decode_attn_pipe_config(int cfg);
/* txl_decode_tail: everything after the LM-head GEMM of a bf16 decode step in one kernel, one CTA per sequence: log-softmax of logits
 * [B, ldl] fp32 (bit-identical to txl_logsoftmax_nll_fwd; also written to scores [B, V] when not NULL), the keyed uniform of
 * txl_decode_uniform, txl_sample's warpers + draw, txl_decode_commit's eos / pad bookkeeping and token store, x0[b, :] = E[token] * emb_scale
 * (the next step's input, bf16) and *pos += 1 (by the last CTA to arrive; `arrive` is one int zeroed once, left zero). */
int txl_decode_tail(const float* logits, int64_t ldl, float* scores, int B, int V, int do_sample, float temperature, int top_k, float top_p,
                    uint64_t seed, int64_t seq_offset, int64_t* tok, int64_t* unfinished, int64_t* out_ids, int64_t ld_out, int col0,
                    int32_t* pos, int* arrive, int64_t eos, int64_t pad, int use_eos, const void* E, void* x0, int d, float emb_scale,
                    void* stream);
int txl_set_pdl(int on);

/* Third-generation bf16 decode step (csrc/decode_persist.cu): embedding row in x -> all L layers -> LM-head GEMM as ONE persistent cooperative
 * kernel (one CTA per SM, the stages of a layer separated by grid barriers) over a ring of cached HIDDEN states: ring[l] [B, mem_len, d] is
 * HF's mems[l] (Appendix A.8'), slot pos % mem_len is overwritten with the layer input each step, and the per-head key / value projections
 * are absorbed into the query / output side (AC = (W_k,h^T (q_h + r_w_bias_h)) . hid_s, out_h = W_v,h sum_s p hid_s), so a cached token costs
 * d*2 bytes per layer instead of the 2*d*2 of a projected k|v cache; BD comes from one [B, H, mem_len+1] table per layer.  Replaces the T=1
 * forward of HF `sample` / `greedy_search` (A.3-A.6, A.7).  Geometry: B <= 64, d_head 64, d_model 128 or 512, H <= 8, d_inner <= 512 or a
 * multiple of 512 (txl_decode_persist_supported).  Per-layer pointers arrive as host arrays of L device pointers: wqkv (qkv_net.weight),
 * wkT ([H, d, 64]: wkT[h, c, e] = W_k[h*64+e, c]), wo, w1, w2, rtab ([mem_len+1, d] = r_net(pos_emb), row x <-> distance mem_len - x), biases
 * and LayerNorm vectors (fp32), ring.  Call once with build_table = 1 (uploads the table into ws, zeroes the barrier counter, synchronises
 * the stream), then once per token with build_table = 0: x [B, d] bf16 holds E[token]*sqrt(d) on entry and the final hidden state on exit,
 * logits [B, ldl] fp32 = x [E ; cluster_weight]^T + bias (Vx columns).  *pos must advance by one between steps (txl_decode_tail does). */
/* profiling hook: device buffer (>= 7 L + 2 uint64) receiving CTA 0's %globaltimer at the start of a step and after every grid barrier; NULL disables */
int txl_decode_persist_set_timestamps(unsigned long long* dev_buf);
int txl_decode_persist_supported(int B, int H, int dh, int d, int di, int mem_len, int L, int Vx);
int64_t txl_decode_persist_ws_bytes(int B, int H, int dh, int d, int di, int mem_len, int L, int Vx);
int txl_decode_persist_step(const void* const* wqkv, const void* const* wkT, const void* const* wo, const void* const* w1, const void* const* w2,
                            const void* const* rtab, const float* const* b1, const float* const* b2, const float* const* rwb,
                            const float* const* rrb, const float* const* ln1w, const float* const* ln1b, const float* const* ln2w,
                            const float* const* ln2b, void* const* ring, const void* E, const float* out_bias, void* x, const int32_t* pos,
                            float* logits, int64_t ldl, void* ws, int build_table, int B, int H, int dh, int d, int di, int mem_len, int L, int Vx,
                            float eps, void* stream);

/* Fourth-generation bf16 decode step (csrc/decode_cluster.cu): the contract of txl_decode_persist_step (same arguments, same hidden-state ring,
 * same absorbed projections), executed by thread-block CLUSTERS of 8 CTAs - CTA r = attention head r - each of which carries up to 8
 * sequences through all L layers and the LM head on its own: activations cross CTAs through distributed shared memory behind hardware
 * cluster barriers (8 per layer), weights stream through warp-private cp.async rings, nothing synchronises across clusters.  Geometry:
 * 8 heads of 64 (d_model 512), d_inner a multiple of 256 (<= 4096), B <= 8 x (SMs / 8).  txl_decode_cluster_set_timestamps: profiling hook
 * (CTA 0's %globaltimer at the start and after the stages of every layer, >= 7 L + 4 uint64). */
int txl_decode_cluster_set_timestamps(unsigned long long* dev_buf);
/* co-resident clusters of 8 CTAs this kernel gets on the current device (cudaOccupancyMaxActiveClusters) */
int txl_decode_cluster_max_clusters(int di);
int txl_decode_cluster_supported(int B, int H, int dh, int d, int di, int mem_len, int L, int Vx);
int64_t txl_decode_cluster_ws_bytes(int B, int H, int dh, int d, int di, int mem_len, int L, int Vx);
int txl_decode_cluster_step(const void* const* wqkv, const void* const* wkT, const void* const* wo, const void* const* w1, const void* const* w2,
                            const void* const* rtab, const float* const* b1, const float* const* b2, const float* const* rwb,
                            const float* const* rrb, const float* const* ln1w, const float* const* ln1b, const float* const* ln2w,
                            const float* const* ln2b, void* const* ring, const void* E, const float* out_bias, void* x, const int32_t* pos,
                            float* logits, int64_t ldl, void* ws, int build_table, int B, int H, int dh, int d, int di, int mem_len, int L, int Vx,
                            float eps, void* stream);

/* ---- mems ring / layout helpers  [A.8' _update_mems] ----------------------------------------------
 * time-major (L?,rows,B,d) <-> batch-major copies used at the Python boundary */
int txl_tm_to_bm(const void* src, void* dst, int rows, int B, int d, int dtype_src, int dtype_dst, void* stream);
int txl_bm_to_tm(const void* src, void* dst, int rows, int B, int d, int dtype_src, int dtype_dst, void* stream);

#ifdef __cplusplus
}
#endif
#endif
