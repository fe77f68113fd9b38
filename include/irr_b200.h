/*
 * irr_b200.h — C ABI of the B200-native retrieval-ranking hot path.
 *
 * This is the drop-in boundary for the similarity / top-k / embedding-loss path of
 * vitasoftAI/ImageRetrievalResearch.  The reference has no FFI: its "interface" for this path is
 * four torch calls made from Python.  Every entry point below names the reference
 * call site(s) it replaces as file:line under the reference tree.
 *
 * Conventions (all entry points):
 *   - extern "C", plain pointers and sizes, no torch / C++ types.
 *   - return irr_status: 0 = OK, negative = argument / capability error detected on the host
 *     before anything is enqueued, positive = a cudaError_t passed through unchanged.
 *   - all pointers except `losses_host`-style out-params documented as host are DEVICE pointers
 *     owned by the caller; outputs and workspaces are caller-allocated; the library never
 *     allocates or frees device memory and never synchronises the device: work is enqueued on
 *     `stream` and the call returns.
 *   - a `*_workspace_bytes` query sizes the scratch buffer of the matching call.  Workspaces of
 *     the loss entry points start with self-resetting sync words: the buffer must be zero-filled
 *     ONCE before its first use and may then be reused by any number of calls on one stream.
 *   - row-major contiguous embeddings, row stride == D elements; base pointers 16-byte aligned;
 *     D % 8 == 0 for IRR_BF16 / IRR_F16 and D % 4 == 0 for IRR_F32 (1536 / 1920 / 2560 in the reference).
 *   - re-entrant and thread-safe for distinct streams + workspaces.
 *   - kernels that wait for CTAs outside their own cluster (uncached searches of more than 512
 *     queries; the fused exchange+merge) are launched cooperatively; when the driver cannot make
 *     the grid co-resident the call takes an equivalent path that needs no co-residency.
 *   - there is no CPU fallback: a device that is not sm_100 makes the bf16 tensor path return
 *     IRR_ERR_UNSUPPORTED_DEVICE.
 */
#ifndef IRR_B200_H_
#define IRR_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define IRR_API
#else
#define IRR_API __attribute__((visibility("default")))
#endif

typedef int32_t irr_status;
/* same object as cudaStream_t / CUstream */
typedef struct CUstream_st* irr_stream_t;

enum {
  IRR_OK = 0,
  IRR_ERR_INVALID_ARG = -1,        /* null pointer, negative size, k out of range ...      */
  IRR_ERR_UNSUPPORTED_DTYPE = -2,
  IRR_ERR_ALIGNMENT = -3,          /* pointer or D violates the alignment contract above   */
  IRR_ERR_WORKSPACE_TOO_SMALL = -4,
  IRR_ERR_K_TOO_LARGE = -5,        /* k > IRR_MAX_K                                         */
  IRR_ERR_UNSUPPORTED_DEVICE = -6, /* not an sm_100 device / driver lacks tensor-map API   */
  IRR_ERR_ROW_TOO_LONG = -7        /* D too large for the shared-memory staged loss kernels */
};

/* IRR_F16: the embeddings the reference's precision=16 training produces
 * (train/train_efficient_cos_con_ce_loss.py:465), feature maps and logits produced under fp16
 * autocast.  Search / similarity entry points read the fp16 values exactly; the loss entry points
 * reproduce autocast's arithmetic on them: b - a of the contrastive distance is an fp16
 * subtraction (utils/contrastive_loss.py:56 is not on an autocast list), everything else fp32
 * (pow, sum, cosine_embedding_loss are on the fp32 list); gradients are emitted in fp16. */
typedef enum { IRR_F32 = 0, IRR_BF16 = 1, IRR_F16 = 2 } irr_dtype;

/* largest k of irr_cosine_topk; up to IRR_MAX_K_FUSED the register-resident epilogues select inside
 * the GEMM kernel (scores never written), above it a dense score block of bounded size is
 * materialised in the workspace and selected row-wise (the notebook's k=150) */
#define IRR_MAX_K 256
#define IRR_MAX_K_FUSED 16

/* which losses a pair/triplet call evaluates (bit mask) */
enum { IRR_LOSS_COSINE_EMBEDDING = 1, IRR_LOSS_CONTRASTIVE = 2 };

/* order of the four scalars of the triplet path, as logged separately by the reference
 * (train/train_efficient_cos_con_ce_loss.py:396-399) */
enum { IRR_L_COS_POS = 0, IRR_L_COS_NEG = 1, IRR_L_CON_POS = 2, IRR_L_CON_NEG = 3 };

/* number of fp32 per-row statistics the loss forward saves for the backward kernel */
#define IRR_ROW_STATS 8

IRR_API int32_t     irr_version(void);
IRR_API const char* irr_status_string(irr_status s);

/* ------------------------------------------------------------------------------------------
 * Cosine similarity + top-k (K1) — replaces, for ALL query rows at once,
 *     sim = cos(fm_ims[idx].unsqueeze(0), fm_poss); vals, inds = torch.topk(sim, k)
 *   train/train_efficient_cos_con_ce_loss.py:89,273,276,385,388
 *   inference/inference.py:169,235,240      inference/training_analysis.ipynb:187,238
 * with cos = CosineSimilarity(dim=1, eps): x1.x2 / (max(|x1|,eps) * max(|x2|,eps)).
 *
 * q [Q,D], g [N,D] of dtype `dt` (F32: FFMA path; BF16 / F16: tcgen05 kind::f16 tensor path);
 * accumulate and emit fp32.
 * g_inv_norm: optional device fp32[N] holding 1/max(|g_row|,eps) (cached by a gallery handle);
 *             NULL = computed inside the call.
 * out_val [Q,k] fp32 sorted descending, ties broken by LOWER gallery index;
 * out_idx [Q,k] int64 = local row index + idx_offset (idx_offset = first row of this shard).
 * Slots beyond N (k > N, N == 0 included) hold (-inf, -1).   1 <= k <= IRR_MAX_K.
 * Non-finite inputs follow torch: a gallery row with a NaN / Inf element scores NaN against every
 * query, NaN orders ABOVE every number (torch.topk's rule), so those rows come first, lower index
 * first; a non-finite query row yields k NaN values (which rows: unspecified, as in torch).
 * For k <= IRR_MAX_K_FUSED the Q x N score matrix is never written to memory.
 * ------------------------------------------------------------------------------------------ */
IRR_API size_t irr_cosine_topk_workspace_bytes(int64_t Q, int64_t N, int32_t D, int32_t k,
                                               irr_dtype dt);
IRR_API irr_status irr_cosine_topk(const void* q, const void* g, const float* g_inv_norm,
                                   int64_t Q, int64_t N, int32_t D, int32_t k, irr_dtype dt,
                                   float eps, int64_t idx_offset, float* out_val,
                                   int64_t* out_idx, void* workspace, size_t workspace_bytes,
                                   irr_stream_t stream);

/* ---- Measurement / test aids: NOT part of the stable ABI (no reference call site stands behind
 * them; they exist for tests/, bench.py and scripts/ and may change between versions) ----------
 *
 * irr_cosine_scores_bf16: the dense score matrix the bf16 tensor-core path ranks, out [Q,N] fp32
 * (same kernel, epilogue writes scores instead of selecting).  Small N only.
 *
 * irr_profile_next_topk (bench.py's roofline leg): arm a pair of caller-created cudaEvent_t; the
 * NEXT irr_cosine_topk call made by this host thread records them on its stream immediately before
 * and after the dominant top-k kernel (excluding any norm pre-pass and the partial-list merge),
 * then disarms.  Pass NULLs to disarm.
 *
 * irr_debug_occupy_sms: launch `ctas` CTAs that each hold `smem_bytes` of shared memory and spin
 * for `nanoseconds` (<= 2 s) on `stream` — a stand-in for a foreign kernel holding SMs, used by the
 * test that the persistent kernels neither hang nor trap when the GPU is not theirs alone. */
IRR_API irr_status irr_cosine_scores_bf16(const void* q, const void* g, int64_t Q, int64_t N,
                                          int32_t D, float eps, float* out_scores,
                                          void* workspace, size_t workspace_bytes,
                                          irr_stream_t stream);
IRR_API void irr_profile_next_topk(void* ev_start, void* ev_stop);
IRR_API irr_status irr_debug_occupy_sms(int32_t ctas, int32_t smem_bytes, int64_t nanoseconds,
                                        irr_stream_t stream);

/* 1/max(|row|,eps) for every row of x [N,D] -> out fp32[N]; what a gallery handle caches. */
IRR_API irr_status irr_row_inv_norms(const void* x, int64_t N, int32_t D, irr_dtype dt, float eps,
                                     float* out, irr_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Candidate merge (K3) — new in the sharded design: after one all-gather of
 * every rank's [Q,k] (score, global index) lists, cand_* are [G,Q,k]; emits the global top-k
 * [Q,k], descending, ties -> lower global index; entries with idx < 0 are padding.
 * ------------------------------------------------------------------------------------------ */
IRR_API irr_status irr_topk_merge(const float* cand_val, const int64_t* cand_idx, int32_t G,
                                  int64_t Q, int32_t k, float* out_val, int64_t* out_idx,
                                  irr_stream_t stream);

/* Same merge reading each rank's lists in place from an all-gather receive buffer: rank g's scores
 * start at cand_val + g*val_rank_stride (floats), its indices at cand_idx + g*idx_rank_stride
 * (int64s); both strides >= Q*k.  Lets the exchange be one collective on a packed message with no
 * pack / unpack copies. */
IRR_API irr_status irr_topk_merge_strided(const float* cand_val, int64_t val_rank_stride,
                                          const int64_t* cand_idx, int64_t idx_rank_stride,
                                          int32_t G, int64_t Q, int32_t k, float* out_val,
                                          int64_t* out_idx, irr_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Peer-memory exchange fused with the merge (K3x) — the sharded gallery's one exchange step
 *  without a collective library: every rank owns an exchange buffer of
 * irr_topk_exchange_bytes() that is mapped into all G processes of the box (CUDA VMM / symmetric
 * memory, done by the caller; zero-filled once, then a barrier, before its first use).
 * peer_bufs: HOST array of the G device pointers as seen from THIS process, peer_bufs[rank] being
 * this rank's own buffer.  One kernel stores this rank's [Q,k] list into slot `rank` of every
 * buffer over NVLink, publishes the call's epoch to each peer (st.release.sys), waits until all G
 * lists of the epoch have landed locally (ld.acquire.sys, watchdog IRR_EXCHANGE_TIMEOUT_MS,
 * default 600 s -> trap) and merges them like irr_topk_merge.  Every rank of the group must make
 * the same sequence of calls (same G, Q, k) on ONE stream per buffer; the epoch lives in the
 * buffer, so the launch is CUDA-graph capturable.  G <= IRR_MAX_PEERS.
 * mode: IRR_XCHG_FUSED the call described above; IRR_XCHG_PUSH store + publish only;
 *       IRR_XCHG_MERGE wait + merge of the epoch last pushed (PUSH then MERGE == FUSED; used to
 *       stage the protocol, e.g. several virtual ranks on one device in the tests).
 *       A stream of searches issued as "MERGE (result of search n-1), then PUSH (lists of
 *       search n)" never makes a rank wait for a peer that is less than one search behind: the
 *       rendezvous of every search is with the peers' PREVIOUS push (lagged exchange; the last
 *       search of the stream is collected with one more MERGE).
 * ------------------------------------------------------------------------------------------ */
#define IRR_MAX_PEERS 16
enum { IRR_XCHG_FUSED = 0, IRR_XCHG_PUSH = 1, IRR_XCHG_MERGE = 2 };
IRR_API size_t irr_topk_exchange_bytes(int32_t G, int64_t Q, int32_t k);
IRR_API irr_status irr_topk_exchange_merge(const float* local_val, const int64_t* local_idx,
                                           void* const* peer_bufs, int32_t G, int32_t rank,
                                           int64_t Q, int32_t k, size_t buf_bytes, int32_t mode,
                                           float* out_val, int64_t* out_idx, irr_stream_t stream);

/* Host plumbing for mapping one rank's exchange buffer into another process of the same box with
 * CUDA IPC (the alternative to symmetric memory when the caller's allocator is cudaMalloc-backed).
 * export: handle (64 bytes, cudaIpcMemHandle_t) of the allocation containing dev_ptr + dev_ptr's
 *         byte offset inside it;  import: open a handle exported by ANOTHER process (peer access
 *         is enabled lazily), *mapped_base is what irr_peer_close takes, the buffer is at
 *         *mapped_base + offset.  These three are the only entry points that touch the driver's
 *         memory-mapping state; none of them launches work. */
IRR_API irr_status irr_peer_export(const void* dev_ptr, uint8_t handle_out[64], uint64_t* offset_out);
IRR_API irr_status irr_peer_import(const uint8_t handle[64], void** mapped_base);
IRR_API irr_status irr_peer_close(void* mapped_base);

/* The whole sharded search in one call: irr_cosine_topk on this rank's gallery shard (rows
 * [idx_offset, idx_offset + N_local) of the global gallery) followed by irr_topk_exchange_merge
 * (IRR_XCHG_FUSED).  Replaces the same reference loop as irr_cosine_topk over a gallery that is
 * row-sharded across the G GPUs of one box; out_* hold the GLOBAL top-k on every rank.
 * workspace: irr_cosine_topk_sharded_workspace_bytes (local search workspace + the local lists). */
IRR_API size_t irr_cosine_topk_sharded_workspace_bytes(int64_t Q, int64_t N_local, int32_t D,
                                                       int32_t k, irr_dtype dt);
IRR_API irr_status irr_cosine_topk_sharded(const void* q, const void* g_local,
                                           const float* g_inv_norm, int64_t Q, int64_t N_local,
                                           int32_t D, int32_t k, irr_dtype dt, float eps,
                                           int64_t idx_offset, void* const* peer_bufs, int32_t G,
                                           int32_t rank, size_t buf_bytes, float* out_val,
                                           int64_t* out_idx, void* workspace,
                                           size_t workspace_bytes, irr_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * top-1 / top-k hit accounting — replaces the per-row Python tests
 *   class flavour    train/train_efficient_cos_con_ce_loss.py:279-281,390-392
 *   instance flavour inference/inference.py:237,242
 * idx [Q,k] int64 (output of irr_cosine_topk).
 * class flavour:    q_label[Q], g_label[N] int64: hit if q_label[i] == g_label[idx[i][j]].
 * instance flavour: q_label = g_label = NULL: hit if idx[i][j] == i + instance_offset.
 * out_hits int64[2] = {#rows with a hit at j==0, #rows with a hit at any j<k}.
 * ------------------------------------------------------------------------------------------ */
IRR_API irr_status irr_topk_hits(const int64_t* idx, int64_t Q, int32_t k, const int64_t* q_label,
                                 const int64_t* g_label, int64_t N, int64_t instance_offset,
                                 int64_t* out_hits, irr_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Class de-duplication of a ranked list — replaces the notebook's loop that walks the top-150
 * and keeps the first 3 DISTINCT classes, then counts top1 / top3
 *   inference/training_analysis.ipynb:240-251
 * val/idx [Q,k] (output of irr_cosine_topk), g_label int64[N], 1 <= n_distinct <= 8.
 * out_label / out_idx / out_val [Q,n_distinct]: the distinct labels in rank order, the gallery row
 * each was first seen at, and its score; unused slots hold (-1, -1, -inf).
 * q_label (optional) + out_hits int64[2] (optional): {#q_label == first label, #q_label among them}.
 * ------------------------------------------------------------------------------------------ */
IRR_API irr_status irr_topk_class_dedup(const float* val, const int64_t* idx, int64_t Q, int32_t k,
                                        const int64_t* g_label, int64_t N, int32_t n_distinct,
                                        const int64_t* q_label, int64_t* out_label,
                                        int64_t* out_idx, float* out_val, int64_t* out_hits,
                                        irr_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Row-wise cosine similarity — replaces cos(x1, x2) itself:
 *   paired scores  train/train_efficient_cos_con_ce_loss.py:377,381   ipynb:232,234
 *   one-vs-gallery train/train_efficient_cos_con_ce_loss.py:273       ipynb:238
 * x2 [N,D]; x1 is [N,D] (x1_rows == N) or a single row broadcast against x2 (x1_rows == 1).
 * out fp32[N].
 * ------------------------------------------------------------------------------------------ */
IRR_API irr_status irr_pair_cosine(const void* x1, int64_t x1_rows, const void* x2, int64_t N,
                                   int32_t D, irr_dtype dt, float eps, float* out,
                                   irr_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Embedding losses (K2) — replaces
 *   ContrastiveLoss.forward          utils/contrastive_loss.py:56-61
 *   torch.nn.CosineEmbeddingLoss     train/train_efficient_cos_con_ce_loss.py:158,230-231,334-335
 *   the four-loss composition        train/train_efficient_cos_con_ce_loss.py:230-237
 *
 * Triplet form: rows (q_i, p_i, n_i), i < B.  losses[4] (device fp32, order IRR_L_*):
 *   cos_pos = red_i (1 - c(q,p))            cos_neg = red_i max(0, c(q,n) - margin_cos)
 *   con_pos = red_i 0.5*|p-q|^2             con_neg = red_i 0.5*relu(margin_con - sqrt(|n-q|^2+1e-9))^2
 *   c(a,b) = a.b / sqrt((|a|^2+1e-12)(|b|^2+1e-12)); red = mean (reduce_mean != 0) or sum.
 * pair_cos: optional device fp32[2*B]: CosineSimilarity(dim=1, eps=pair_eps) of (q,p) then (q,n)
 *   — the cos_sims / cos_unsims the reference logs (:377-382,402-403).
 * row_stats: optional device fp32[B*IRR_ROW_STATS] saved for irr_triplet_loss_bwd.
 * dq/dp/dn: optional (all three or none) gradients of sum_j grad_scale[j]*losses[j], written in
 *   the same pass over the rows (fused forward+backward); dtype `dt`.
 * n may be NULL ("pair form": only the *_POS terms... see irr_pair_loss_* below).
 * ------------------------------------------------------------------------------------------ */
IRR_API size_t irr_triplet_loss_workspace_bytes(int64_t B, int32_t D, irr_dtype dt);
IRR_API irr_status irr_triplet_loss_fwd_bwd(const void* q, const void* p, const void* n, int64_t B,
                                            int32_t D, irr_dtype dt, float margin_cos,
                                            float margin_con, int32_t reduce_mean, float pair_eps,
                                            float* losses, float* pair_cos, float* row_stats,
                                            void* dq, void* dp, void* dn,
                                            const float grad_scale[4], void* workspace,
                                            size_t workspace_bytes, irr_stream_t stream);
/* Backward from saved row statistics; grad_out = device fp32[4], the upstream gradients of the
 * four scalars (read on the device: no host sync). */
IRR_API irr_status irr_triplet_loss_bwd(const void* q, const void* p, const void* n,
                                        const float* row_stats, const float* grad_out, int64_t B,
                                        int32_t D, irr_dtype dt, float margin_cos, float margin_con,
                                        int32_t reduce_mean, void* dq, void* dp, void* dn,
                                        irr_stream_t stream);

/* Pair form: ONE loss over rows (a_i, b_i) with a per-call or per-row label — the literal
 * signatures ContrastiveLoss(margin)(fm1, fm2, label, mean) and
 * CosineEmbeddingLoss(margin)(x1, x2, target).
 *   kind = IRR_LOSS_CONTRASTIVE:      label y in [0,1] : 0.5*(y*d + (1-y)*relu(m - sqrt(d+1e-9))^2)
 *   kind = IRR_LOSS_COSINE_EMBEDDING: target t in {1,-1}: t==1 ? 1-c : max(0, c-m)
 * label: device fp32[label_count], label_count == 1 (broadcast) or B.
 * loss: device fp32[1].  row_stats: optional fp32[B*IRR_ROW_STATS].
 * da/db: optional (both or none) gradient of grad_scale*loss (fused forward+backward). */
IRR_API size_t irr_pair_loss_workspace_bytes(int64_t B, int32_t D, irr_dtype dt);
IRR_API irr_status irr_pair_loss_fwd_bwd(const void* a, const void* b, const float* label,
                                         int64_t label_count, int64_t B, int32_t D, irr_dtype dt,
                                         int32_t kind, float margin, int32_t reduce_mean,
                                         float* loss, float* row_stats, void* da, void* db,
                                         float grad_scale, void* workspace, size_t workspace_bytes,
                                         irr_stream_t stream);
IRR_API irr_status irr_pair_loss_bwd(const void* a, const void* b, const float* label,
                                     int64_t label_count, const float* row_stats,
                                     const float* grad_out, int64_t B, int32_t D, irr_dtype dt,
                                     int32_t kind, float margin, int32_t reduce_mean, void* da,
                                     void* db, irr_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Producer / consumer steps next to the path
 *
 * get_fm — global average pool, replaces
 *     pool = AvgPool2d((fm.shape[2], fm.shape[3])); torch.reshape(pool(fm), (-1, fm.shape[1]))
 *   train/train_efficient_cos_con_ce_loss.py:103-122
 * fm: contiguous [rows = B*C, hw = H*W] of in_dt (F32 / BF16 / F16); out: [rows] of out_dt
 * (F32 / BF16) = mean over hw, i.e. the [B,C] embedding rows the path consumes.
 * Backward: grad_fm[r, :] = grad_out[r] / hw.
 * ------------------------------------------------------------------------------------------ */
IRR_API irr_status irr_avgpool_fwd(const void* fm, irr_dtype in_dt, int64_t rows, int32_t hw,
                                   void* out, irr_dtype out_dt, irr_stream_t stream);
IRR_API irr_status irr_avgpool_bwd(const void* grad_out, irr_dtype go_dt, int64_t rows, int32_t hw,
                                   void* grad_fm, irr_dtype gf_dt, irr_stream_t stream);

/* Cross-entropy of two logits tensors against one target vector, replaces
 *     loss_ce = CrossEntropyLoss()(lbl_ims, clss) + CrossEntropyLoss()(lbl_poss, clss)
 *   train/train_efficient_cos_con_ce_loss.py:160,240-242
 * a, b: [B,C] logits of dt (F32 / BF16 / F16); target int64[B] (ignore_index rows are skipped,
 * 'mean' divides by the number of kept rows).  out_loss: device fp32[3] = {sum, ce(a), ce(b)}.
 * The forward leaves what the backward needs in `workspace` (irr_ce_pair_workspace_bytes);
 * backward: grad_out = device fp32[1] upstream gradient of out_loss[0]; da / db [B,C] of dt. */
IRR_API size_t irr_ce_pair_workspace_bytes(int64_t B);
IRR_API irr_status irr_ce_pair_fwd(const void* a, const void* b, const int64_t* target, int64_t B,
                                   int32_t C, irr_dtype dt, int64_t ignore_index, float* out_loss,
                                   void* workspace, size_t workspace_bytes, irr_stream_t stream);
IRR_API irr_status irr_ce_pair_bwd(const void* a, const void* b, const int64_t* target, int64_t B,
                                   int32_t C, irr_dtype dt, int64_t ignore_index,
                                   const float* grad_out, const void* workspace, void* da, void* db,
                                   irr_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* IRR_B200_H_ */
