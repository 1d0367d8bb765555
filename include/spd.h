/*
 * spd.h — C ABI of libspd.so, the B200 (sm_100a) implementation of Semi-PD's concurrent
 * prefill + decode forward path.
 *
 * Every entry point is `extern "C"`, takes raw device pointers, explicit sizes / strides and the
 * CUDA stream to launch on, performs no allocation and no host synchronisation (so the calls can
 * be captured into CUDA graphs), and returns 0 on success, a positive cudaError_t on a CUDA
 * failure, or a negative SPD_E_* code on an argument error.  Nothing here throws.  The text of
 * the most recent failure on the calling thread is available from spd_last_error_string().
 *
 * Each declaration cites the reference interface it replaces.  Paths are relative to the
 * reference checkout (infinigence/Semi-PD @ 2565f9f): S/ = python/sglang/srt/, K/ = sgl-kernel/.
 */
#ifndef SPD_H_
#define SPD_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* spd_stream_t; /* cudaStream_t */

/* element types of activations / weights / KV pool */
#define SPD_BF16 0
#define SPD_F16 1

/* argument-check error codes (CUDA errors are returned as positive cudaError_t values) */
#define SPD_E_BADARG (-1)   /* null pointer, non-positive size, bad enum */
#define SPD_E_ALIGN (-2)    /* pointer or stride not 16-byte aligned / dim not multiple of 8 */
#define SPD_E_SHAPE (-3)    /* unsupported shape (head_dim, group size, K multiple …) */
#define SPD_E_WORKSPACE (-4)/* workspace too small */
#define SPD_E_DRIVER (-5)   /* driver entry point missing (green contexts, tensor maps) */

/* ---------------------------------------------------------------------------------------------
 * library / device
 * ------------------------------------------------------------------------------------------- */
int spd_version(void);
const char* spd_last_error_string(void);
/* replaces semi_pd_ipc.get_device_sm_count (semi-pd-ipc/ipc.cpp:87-92) */
int spd_device_sm_count(int device, int* sm_count);

/* ---------------------------------------------------------------------------------------------
 * SM partitions — replaces the per-process MPS caps (CUDA_MPS_ACTIVE_THREAD_PERCENTAGE set in
 * S/entrypoints/engine.py:591-593,632-634) with two CUDA green contexts that own DISJOINT SM
 * sets inside one process.  `prefill_sms + decode_sms` may not exceed the device SM count; the
 * driver rounds each group to its SM-group granularity and reports the granted counts.
 * The returned streams belong to the partition; destroy with spd_partition_destroy.
 * ------------------------------------------------------------------------------------------- */
typedef struct spd_partition spd_partition_t;
int spd_partition_create(int device, int prefill_sms, int decode_sms, spd_partition_t** out);
int spd_partition_sm_counts(const spd_partition_t* p, int* prefill_sms, int* decode_sms);
/* which = 0: prefill stream, 1: decode stream */
int spd_partition_stream(const spd_partition_t* p, int which, spd_stream_t* stream);
int spd_partition_destroy(spd_partition_t* p);

/* ---------------------------------------------------------------------------------------------
 * elementwise ops — twins of torch.ops.sgl_kernel.* (K/csrc/torch_extension.cc:49-73)
 * ------------------------------------------------------------------------------------------- */
/* out[t,:] = x[t,:] * rsqrt(mean(x[t,:]^2) + eps) * w      (fp32 math)
 * replaces sgl_kernel.rmsnorm (K/python/sgl_kernel/elementwise.py:9-19; semantics
 * S/layers/layernorm.py:59-76).  Rows are `x_stride` / `out_stride` elements apart. */
int spd_rmsnorm(void* out, const void* x, const void* w, int64_t tokens, int64_t dim,
                int64_t x_stride, int64_t out_stride, float eps, int dtype, spd_stream_t stream);
/* residual += x (rounded to dtype, stored) ; x = rmsnorm(residual) * w — both IN PLACE.
 * replaces sgl_kernel.fused_add_rmsnorm (K/csrc/elementwise/fused_add_rms_norm_kernel.cu:24-55) */
int spd_fused_add_rmsnorm(void* x, void* residual, const void* w, int64_t tokens, int64_t dim,
                          float eps, int dtype, spd_stream_t stream);
/* out[t, :d] = silu(x[t, :d]) * x[t, d:2d]
 * replaces sgl_kernel.silu_and_mul (K/csrc/elementwise/activation.cu:36-50) */
int spd_silu_and_mul(void* out, const void* x, int64_t tokens, int64_t d, int dtype,
                     spd_stream_t stream);
/* In-place rotary embedding of q [T, Hq, D] and k [T, Hk, D] (token strides in elements) using the
 * fp32 cos_sin_cache [max_pos, rot_dim] (cos || sin) indexed by pos_ids [T] (int64).
 * interleave = 0: NeoX halves, 1: GPT-J pairs.
 * replaces sgl_kernel.apply_rope_pos_ids_cos_sin_cache (K/csrc/elementwise/rope.cu:22-89) */
int spd_rope_inplace(void* q, void* k, const float* cos_sin_cache, const int64_t* pos_ids,
                     int64_t tokens, int num_q_heads, int num_k_heads, int head_dim, int rot_dim,
                     int64_t q_stride, int64_t k_stride, int interleave, int dtype,
                     spd_stream_t stream);
/* Fused: RoPE on q and k in place (as above) AND scatter of the rotated k rows and the v rows into
 * the KV pool at out_cache_loc[t]:  k_pool[loc[t]] = k[t], v_pool[loc[t]] = v[t]  (pool rows are
 * Hk*D contiguous elements).  Replaces rope (above) + MHATokenToKVPool.set_kv_buffer
 * (S/mem_cache/memory_pool.py:316-346).  loc is int64. */
int spd_rope_kvwrite(void* q, void* k, const void* v, void* k_pool, void* v_pool,
                     const float* cos_sin_cache, const int64_t* pos_ids, const int64_t* loc,
                     int64_t tokens, int num_q_heads, int num_k_heads, int head_dim, int rot_dim,
                     int64_t q_stride, int64_t k_stride, int64_t v_stride, int interleave,
                     int dtype, spd_stream_t stream);
/* k_pool[loc[t]] = k[t]; v_pool[loc[t]] = v[t] — set_kv_buffer alone (S/mem_cache/memory_pool.py:316-346).
 * v and v_pool may both be NULL: single-buffer pools (MLATokenToKVPool.set_kv_buffer, :426-452). */
int spd_kv_write(const void* k, const void* v, void* k_pool, void* v_pool, const int64_t* loc,
                 int64_t tokens, int64_t row_elems, int64_t k_stride, int64_t v_stride, int dtype,
                 spd_stream_t stream);
/* out[i, :] = table[idx[i], :]   (row_elems 16-bit elements per row; idx int64).
 * Embedding lookup (S/layers/vocab_parallel_embedding.py:469-490) and the last-token gather of
 * the logits processor (S/layers/logits_processor.py:244-262). */
int spd_gather_rows(void* out, const void* table, const int64_t* idx, int64_t n, int64_t row_elems,
                    int64_t table_rows, int dtype, spd_stream_t stream);
/* ids[b] = argmax_v logits[b, v] (first maximal index, as torch.argmax) — greedy branch of
 * Sampler.forward (S/layers/sampler.py:73-77,171).  logits fp32 [B, V] with row stride ld. */
int spd_argmax_f32(int32_t* ids, const float* logits, int64_t batch, int64_t vocab, int64_t ld,
                   spd_stream_t stream);
/* Greedy sampling over a vocab-SHARDED lm_head (ParallelLMHead, S/layers/vocab_parallel_embedding.py
 * + the all-gather of S/layers/logits_processor.py:426-427) without moving the logits:
 * spd_argmax_partial_f32 writes one 16-byte row per request {f32 max, i32 index_offset + argmax,
 * 0, 0} for this rank's [B, vocab] shard; after an all-gather of those rows ([parts][B] rows),
 * spd_argmax_merge picks the winner (ties -> smallest index == torch.argmax of the full row). */
int spd_argmax_partial_f32(void* pairs, const float* logits, int64_t batch, int64_t vocab, int64_t ld,
                           int64_t index_offset, spd_stream_t stream);
int spd_argmax_merge(int32_t* ids, const void* pairs, int parts, int64_t batch, spd_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * batch metadata (integer work; bit-exact)
 * ------------------------------------------------------------------------------------------- */
/* positions[extend_start_loc[r] + j] = prefix_lens[r] + j ; extend_start_loc = exclusive cumsum of
 * extend_lens.  Optionally (token_req != NULL) also writes the request index of every token.
 * replaces compute_position_kernel (S/model_executor/forward_batch_info.py:393-446) */
int spd_compute_positions(int64_t* positions, int32_t* extend_start_loc, int32_t* token_req,
                          const int32_t* prefix_lens, const int32_t* extend_lens, int64_t batch,
                          spd_stream_t stream);
/* positions[b] = max(seq_lens[b] - 1, 0) — clamp_position (forward_batch_info.py:466-468) */
int spd_decode_positions(int64_t* positions, const int64_t* seq_lens, int64_t batch,
                         spd_stream_t stream);
/* kv_indices[kv_indptr[b] + j] = req_to_token[req_pool_indices[b], kv_start[b] + j], j < lens[b]
 * replaces create_flashinfer_kv_indices_triton (S/layers/attention/utils.py:5-39).
 * kv_start may be NULL (= 0). req_pool_indices / lens int64 or int32 per `idx64`. */
int spd_build_kv_indices(int32_t* kv_indices, const int32_t* req_to_token,
                         const int64_t* req_pool_indices, const int32_t* lens,
                         const int32_t* kv_indptr, const int32_t* kv_start, int64_t batch,
                         int64_t req_to_token_stride, spd_stream_t stream);
/* req_to_token[req[b], pre[b] + j] = out_cache_loc[cumsum(ext)[b] + j], j < ext[b]
 * replaces write_req_to_token_pool_triton (S/managers/schedule_batch.py:1446-1480) */
int spd_write_req_to_token(int32_t* req_to_token, const int64_t* req_pool_indices,
                           const int64_t* pre_lens, const int64_t* seq_lens,
                           const int64_t* extend_lens, const int64_t* out_cache_loc, int64_t batch,
                           int64_t req_to_token_stride, spd_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * OPT-family ops (BASELINE configs[0]: opt-125m plumbing + logit check).  The reference has no OPT
 * model; these follow torch.nn.LayerNorm / F.relu / the `residual + hidden_states` adds of
 * transformers' modeling_opt.py (OPTDecoderLayer.forward), fp32 math, 16-bit I/O.
 * spd_layernorm: out = (x - mean) / sqrt(var + eps) * gamma + beta per row (dim <= 8192).
 * spd_bias_act_residual: out = act(x + bias) + residual, rows contiguous; bias / residual may be
 * NULL; act: 0 = identity, 1 = ReLU.  out may alias x.
 * ------------------------------------------------------------------------------------------- */
int spd_layernorm(void* out, const void* x, const void* gamma, const void* beta, int64_t tokens,
                  int64_t dim, int64_t x_stride, int64_t out_stride, float eps, int dtype,
                  spd_stream_t stream);
int spd_bias_act_residual(void* out, const void* x, const void* bias, const void* residual,
                          int64_t tokens, int64_t dim, int act, int dtype, spd_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * GEMM — C[M,N] = A[M,K] * W[N,K]^T, 16-bit inputs, fp32 accumulation in TMEM (tcgen05 + TMA).
 * replaces UnquantizedLinearMethod.apply = F.linear (S/layers/linear.py:165-172) for
 * W_qkv / W_o / W_gate_up / W_down and the lm_head matmul (S/layers/logits_processor.py:416-419).
 * lda/ldw/ldc are row strides in elements.  out_f32 != 0: C is fp32 holding the dtype-rounded
 * result (== `.float()` of the 16-bit product, logits_processor.py:440).
 * epilogue: SPD_EPI_NONE, or SPD_EPI_SILU_MUL: W is [2N, K] (gate rows || up rows, the
 * MergedColumnParallelLinear order, S/models/llama.py:67-73) and C[M,N] = silu(A Wg^T) * (A Wu^T),
 * both products rounded to dtype first (== gate_up_proj followed by SiluAndMul).
 * workspace: device scratch for the stream-K partials and their "published" flags
 * (spd_gemm_workspace_bytes()), zero-initialised ONCE by the caller and then owned by one stream
 * (the kernels re-arm the flags themselves; two streams must not share one workspace).  May be
 * NULL: the call then never splits a tile.
 * sm_budget: number of SMs the call may occupy (0 = all); a worker running on a green-context
 * partition passes its partition's SM count so that persistent grids and the stream-K split match it.
 * ------------------------------------------------------------------------------------------- */
#define SPD_EPI_NONE 0
#define SPD_EPI_SILU_MUL 1
int64_t spd_gemm_workspace_bytes(void);
int spd_gemm(void* C, const void* A, const void* W, int64_t M, int64_t N, int64_t K, int64_t lda,
             int64_t ldw, int64_t ldc, int dtype, int out_f32, int epilogue, void* workspace,
             int64_t workspace_bytes, int sm_budget, spd_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * attention over the token-granular paged KV pool (page_size = 1; slot 0 is the dummy slot).
 * Pool layout per layer: k_pool, v_pool [slots, Hk, D] (S/mem_cache/memory_pool.py:231-250).
 * Slots of request r are req_to_token[req_pool_indices[r], 0 .. seq_len).
 * ------------------------------------------------------------------------------------------- */
/* Ragged chunked-prefill ("extend") attention, causal over [prefix || new]:
 * query token j of request r (row extend_start_loc[r] + j of q) attends keys 0 .. prefix_r + j.
 * The new tokens' K/V must already be in the pool.  q/out: [T, Hq, D], token stride in elements.
 * replaces FlashInferAttnBackend.forward_extend's ragged + paged + merge_state kernels
 * (S/layers/attention/flashinfer_backend.py:386-452); semantics
 * S/layers/attention/torch_native_backend.py:26-109.
 * sm_budget: SMs of the caller's partition (0 = whole device) — the kernel is persistent, one CTA
 * per SM. */
int spd_attn_prefill_ragged(void* out, const void* q, const void* k_pool, const void* v_pool,
                            const int32_t* req_to_token, const int64_t* req_pool_indices,
                            const int32_t* prefix_lens, const int32_t* extend_lens,
                            const int32_t* extend_start_loc, int64_t batch, int64_t total_tokens,
                            int64_t max_extend_len, int num_q_heads, int num_kv_heads,
                            int head_dim, int64_t q_stride, int64_t out_stride,
                            int64_t req_to_token_stride, float sm_scale, int dtype,
                            int sm_budget, spd_stream_t stream);
/* Batched paged-decode attention: one query row per request over its seq_lens[b] cached rows
 * (including the token just written).  Split-KV with an in-kernel merge; workspace from
 * spd_attn_decode_workspace_bytes().  replaces FlashInferAttnBackend.forward_decode
 * (flashinfer_backend.py:454-488); semantics torch_native_backend.py:111-179. */
int64_t spd_attn_decode_workspace_bytes(int64_t max_batch, int num_q_heads, int head_dim);
int spd_attn_decode_paged(void* out, const void* q, const void* k_pool, const void* v_pool,
                          const int32_t* req_to_token, const int64_t* req_pool_indices,
                          const int64_t* seq_lens, int64_t batch, int num_q_heads,
                          int num_kv_heads, int head_dim, int64_t q_stride, int64_t out_stride,
                          int64_t req_to_token_stride, float sm_scale, int dtype, void* workspace,
                          int64_t workspace_bytes, int sm_budget, spd_stream_t stream);
/* Plain SIMT implementations of the two calls above (no tensor cores).  Bring-up / on-device
 * cross-check only; the product path never calls them. */
int spd_attn_prefill_simt(void* out, const void* q, const void* k_pool, const void* v_pool,
                          const int32_t* req_to_token, const int64_t* req_pool_indices,
                          const int32_t* prefix_lens, const int32_t* extend_lens,
                          const int32_t* extend_start_loc, int64_t batch, int64_t total_tokens,
                          int64_t max_extend_len, int num_q_heads, int num_kv_heads, int head_dim,
                          int64_t q_stride, int64_t out_stride, int64_t req_to_token_stride,
                          float sm_scale, int dtype, spd_stream_t stream);
int spd_attn_decode_simt(void* out, const void* q, const void* k_pool, const void* v_pool,
                         const int32_t* req_to_token, const int64_t* req_pool_indices,
                         const int64_t* seq_lens, int64_t batch, int num_q_heads, int num_kv_heads,
                         int head_dim, int64_t q_stride, int64_t out_stride,
                         int64_t req_to_token_stride, float sm_scale, int dtype, void* workspace,
                         int64_t workspace_bytes, spd_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * tensor-parallel all-reduce over NVLink peer memory (one process per GPU) — the `spd_allreduce`
 * of SURVEY §8b's export list.  Replaces the custom all-reduce the reference sends every
 * RowParallelLinear output below 8 MiB through (S/layers/linear.py:1265-1266 →
 * S/distributed/parallel_state.py:376-436 → device_communicators/custom_all_reduce.py:147-160,
 * 447-480; alternative K/csrc/allreduce/trt_reduce_internal.cu:190-300); the reference sends larger
 * messages (prefill batches) to ncclAllReduce (parallel_state.py:431-436) — the host side here
 * passes them through this entry point in pieces of at most max_message_bytes.
 *
 * spd_comm_create allocates this rank's symmetric buffer (3 x max_message_bytes + 8 KB) on the
 * current device and writes its 64-byte cudaIpcMemHandle to ipc_handle_out; the caller gathers
 * all ranks' handles (rank order) with any host-side transport and passes them to
 * spd_comm_connect, which maps the peers (cudaIpcOpenMemHandle).  One communicator belongs to
 * ONE stream: the prefill and the decode worker of a GPU each create their own, like the
 * reference's separate NCCL worlds for P and D (S/managers/scheduler.py:250-259).
 *
 * spd_allreduce: out[i] = sum over ranks of in[i], i < count (16-bit elements, count % 8 == 0,
 * count*2 <= max_message_bytes), fp32 accumulation in rank order — bit-identical on all ranks.
 * `in` and `out` are ordinary device buffers (out may alias in).  All ranks must issue the same
 * sequence of calls (same counts) on a communicator.  The call is stream-ordered, allocation- and
 * sync-free and may be captured into a CUDA graph and replayed any number of times (the barrier
 * sequence numbers live in device memory).  algo: 0 = auto (the reference's dispatch: one-shot
 * below 512 KB for <= 4 ranks / 256 KB for 8 ranks, two-shot above), 1 = one-shot, 2 = two-shot.
 * A peer that does not arrive within SPD_AR_TIMEOUT_MS (default 10000) sets a sticky error that
 * spd_comm_status reports; the GPU never hangs.
 * ------------------------------------------------------------------------------------------- */
#define SPD_IPC_HANDLE_BYTES 64
typedef struct spd_comm spd_comm_t;
int spd_comm_create(int rank, int world, int64_t max_message_bytes, spd_comm_t** out,
                    void* ipc_handle_out);
int spd_comm_connect(spd_comm_t* comm, const void* all_handles);
int64_t spd_comm_max_bytes(const spd_comm_t* comm);
int spd_comm_status(spd_comm_t* comm, int* timed_out);
int spd_allreduce(spd_comm_t* comm, void* out, const void* in, int64_t count, int dtype, int algo,
                  spd_stream_t stream);
/* The all-reduce fused with its consumer (round 2): x = sum over ranks of in[rows, dim];
 * residual += x (rounded to the dtype as fused_add_rmsnorm stores it); out = RMSNorm(residual) * weight
 * — S/layers/linear.py:1265-1266 (tensor_model_parallel_all_reduce after o_proj / down_proj)
 * followed by S/layers/layernorm.py:46-57 (fused_add_rmsnorm) in ONE kernel over the peer-mapped
 * slots.  Bit-identical to spd_allreduce + spd_fused_add_rmsnorm.  dim % 8 == 0, dim <= 8192,
 * rows * dim * 2 <= max_message_bytes; out may alias in.  Same ordering / graph-capture rules. */
int spd_allreduce_add_rmsnorm(spd_comm_t* comm, void* out, const void* in, void* residual,
                              const void* weight, int64_t rows, int64_t dim, float eps, int dtype,
                              spd_stream_t stream);
/* out[r * bytes_per_rank + i] = rank r's in[i] (bytes_per_rank % 16 == 0, <= max_message_bytes):
 * the all-gather after the vocab-sharded lm_head (S/layers/logits_processor.py:426-427), here of
 * the per-shard (max, argmax) rows below — and of small logits blocks.  Same ordering /
 * graph-capture rules as spd_allreduce. */
int spd_allgather(spd_comm_t* comm, void* out, const void* in, int64_t bytes_per_rank,
                  spd_stream_t stream);
int spd_comm_destroy(spd_comm_t* comm);

/* ---------------------------------------------------------------------------------------------
 * DeepSeek-V2-Lite path (BASELINE configs[3]): absorbed MLA attention over the latent KV pool,
 * grouped GEMMs and the MoE plumbing.
 * ------------------------------------------------------------------------------------------- */
/* Grouped GEMM on the tcgen05 kernel of spd_gemm (normal regime, 128-row tiles).
 * rows mode (tile_group != NULL) — the two expert GEMMs of the fused MoE
 *   (S/layers/moe/fused_moe_triton/fused_moe.py:961-1168): A [M, K] holds the routed tokens sorted by
 *   expert in 128-row tiles (spd_moe_align / spd_moe_gather), tile_group[i] = expert of rows
 *   [128 i, 128 i + 128) or < 0 (unused tile, skipped); W = the experts' weights stacked,
 *   [num_groups * N, K] — with SPD_EPI_SILU_MUL [num_groups * 2N, K], gate rows || up rows per expert;
 *   C [M, N].  a_group_stride must be 0.
 * cols mode (tile_group == NULL) — block-diagonal products, the absorbed-MLA BMMs
 *   torch.bmm(q_nope, w_kc) / torch.bmm(attn_out, w_vc) (S/models/deepseek_v2.py:667,702):
 *   C[:, g*N:(g+1)*N] = A[:, g*a_group_stride : g*a_group_stride + K] * W[g*N:(g+1)*N, :K]^T,
 *   A [M, a_cols] with row stride lda, N % 128 == 0, K % 64 == 0.  c_group_stride > 0 places group g's
 *   N output columns at column g * c_group_stride of C instead of g * N.
 * w_group_col_stride > 0 (rows mode, no epilogue): the experts' weights are concatenated along K
 *   instead of stacked — W [N, num_groups * w_group_col_stride], expert g = columns
 *   [g * stride, g * stride + K) — the layout the decode-sized dense path multiplies as ONE GEMM. */
int spd_gemm_grouped(void* C, const void* A, const void* W, int64_t M, int64_t N, int64_t K, int64_t lda,
                     int64_t ldw, int64_t ldc, int num_groups, const int32_t* tile_group,
                     int64_t a_group_stride, int64_t a_cols, int64_t c_group_stride, int64_t w_group_col_stride,
                     int dtype, int epilogue, int sm_budget, spd_stream_t stream);
/* Router: softmax over the experts (fp32) + greedy top-k; weights = prob (renormalised over the k
 * if renormalize) * scale.  logits [tokens, num_experts] 16-bit.  replaces select_experts /
 * fused_topk (S/layers/moe/topk.py:79-120,128-163). */
int spd_moe_route(int32_t* topk_ids, float* topk_w, const void* logits, int64_t tokens, int num_experts,
                  int top_k, float scale, int renormalize, int dtype, spd_stream_t stream);
/* Expert alignment: sorts the tokens*top_k routing entries by expert, every expert padded to whole
 * 128-row tiles.  row_token[max_rows]: source token of each sorted row (-1 = padding);
 * inv_pos[tokens*top_k]: sorted row of entry (token, j); tile_group[max_rows/128]: expert per tile
 * (-1 = unused).  max_rows >= spd_moe_align_rows().  replaces moe_align_block_size
 * (K/csrc/moe/moe_align_kernel.cu:28,44). */
int64_t spd_moe_align_rows(int64_t tokens, int top_k, int num_experts);
int spd_moe_align(int32_t* row_token, int32_t* inv_pos, int32_t* tile_group, const int32_t* topk_ids,
                  int64_t tokens, int top_k, int num_experts, int64_t max_rows, spd_stream_t stream);
/* xs[r, :] = x[row_token[r], :] (zeros for padding rows) */
int spd_moe_gather(void* xs, const void* x, const int32_t* row_token, int64_t rows, int64_t dim,
                   int64_t x_stride, int dtype, spd_stream_t stream);
/* out[t, :] = (shared ? shared[t, :] : 0) + sum_j topk_w[t, j] * y[inv_pos[t*top_k + j], :]
 * (the weighted sum of fused_moe + the shared experts' output, S/models/deepseek_v2.py:205-218) */
int spd_moe_combine(void* out, const void* y, const void* shared, const int32_t* inv_pos, const float* topk_w,
                    int64_t tokens, int top_k, int64_t dim, int dtype, spd_stream_t stream);
/* Decode-sized batches run every expert as two plain spd_gemm calls over the stacked weights;
 * between them: act[t, e*inter + i] = w(t, e) * silu(gate_up[t, e, i]) * gate_up[t, e, inter + i],
 * w(t, e) = 0 for experts token t did not select.  gate_up [tokens, num_experts * 2 * inter]. */
int spd_moe_dense_act(void* act, const void* gate_up, const int32_t* topk_ids, const float* topk_w,
                      int64_t tokens, int top_k, int num_experts, int64_t inter, int dtype,
                      spd_stream_t stream);
/* Absorbed MLA attention (MQA over the latent pool), decode (seq_lens != NULL: one query per request
 * over its seq_lens[b] rows) or ragged extend (seq_lens == NULL: query token j of request r, row
 * extend_start_loc[r] + j of q, attends keys 0 .. prefix_lens[r] + j).  kv_pool [slots, kv_lora_rank +
 * rope_dim] = MLATokenToKVPool rows (S/mem_cache/memory_pool.py:379-452); K = the whole row, V = its
 * first kv_lora_rank columns; q [tokens, H, kv_lora_rank + rope_dim] (q_nope already multiplied by
 * w_kc), out [tokens, H, kv_lora_rank].  replaces DeepseekV2AttentionMLA.forward_absorb's attention
 * call (S/models/deepseek_v2.py:684-700 → S/layers/attention/triton_ops/decode_attention.py:235-470,
 * extend_attention.py:291-410). */
int64_t spd_attn_mla_workspace_bytes(int num_heads, int kv_lora_rank);
int spd_attn_mla(void* out, const void* q, const void* kv_pool, const int32_t* req_to_token,
                 const int64_t* req_pool_indices, const int64_t* seq_lens, const int32_t* prefix_lens,
                 const int32_t* extend_start_loc, int64_t batch, int64_t total_tokens, int num_heads,
                 int kv_lora_rank, int rope_dim, int64_t q_stride, int64_t out_stride,
                 int64_t req_to_token_stride, float sm_scale, int dtype, void* workspace,
                 int64_t workspace_bytes, int sm_budget, spd_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* SPD_H_ */
