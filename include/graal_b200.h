/*
 * graal_b200 -- C-ABI of the B200-native GRAAL MCMC scoring path.
 *
 * Drop-in boundary: today this path sits behind PyCUDA function handles obtained with
 * module.get_function(name) (reference cuda_lib_gl.py:378-402) and launched with raw device
 * pointers.  Every entry point below names the reference kernel(s) / host routine it replaces.
 *
 * Conventions
 *   - every function returns 0 on success, a negative code on error; the message is available
 *     through graal_last_error() (thread-local).  No exceptions or C++ types cross the ABI.
 *   - all bulk buffers are caller-owned DEVICE pointers (torch tensors on the Python side); the
 *     library owns only its scratch (geometry tables, partial sums, derived level tables).
 *   - a context is bound to one GPU and one CUDA stream and must be used from one thread at a time
 *     (the reference owns its CUDA context from a single Python thread, main_gl.py:690-706).
 *   - results written to device pointers are valid after the context's stream is synchronised
 *     (graal_score_proposal / graal_score_step: after graal_join, graal_fetch or graal_sync).
 *
 * Environment switches read at graal_ctx_create (A/B measurements; the defaults are the measured best):
 *   GRAAL_LANES=n    (1..4, default 3)  proposals of a step scored concurrently; 1 = serial on the context stream
 *   GRAAL_GRAPHS=0   launch every kernel individually instead of replaying captured CUDA graphs
 *   GRAAL_PAIRING=0  score candidates 3, 5, 7 like the others instead of as deltas against 2, 4, 6
 *   GRAAL_FORK=0     contact / band passes of a proposal on one stream instead of three
 *   GRAAL_SMEM_CID=1 contig-id table of the contact pass in shared memory (math modes 0 / 1 only; slower)
 *   GRAAL_FULL_WIN=0 gather-everything contact pass instead of the windowed one; GRAAL_WIN_UNROLL / _MINB / _SUB: its variants
 *   GRAAL_DELTA_REL=0 / GRAAL_BAND_FAST=0 general delta kernels on uniform levels; GRAAL_DELTA_UNI=1 union windows in the delta contact pass
 *   GRAAL_FUSED_PROLOGUE=0 statistics + relabel as the two general launch sequences
 *   GRAAL_PUBLISH=0  graal_draw_commit hands the results over with cudaMemcpyAsync + an event instead of k_publish into mapped memory
 *   (Python side: GRAAL_DEVICE_DRAW=0 candidate draw and commit on the host after the fetch instead of graal_draw_commit)
 *   GRAAL_WIN_STAB=0|1 law table of the windowed contact pass through L1 / from a 32 KB coarse copy in shared memory (default: timed once per level)
 *   GRAAL_DELTA_SPLIT=n (1..32, default 8) slices a row is cut into when U is short; GRAAL_DELTA_MINB=2|3|4 CTAs per SM of the delta contact pass
 *   GRAAL_BAND_SPLIT=n (1..32, default 1) warps the band of one bin of a short U is cut over in the fast band delta kernels
 *
 * State layout ("slots"): the reference keeps the genome in a struct of 14 int* (kernels3.cu:9-24,
 * packed by gpustruct.py).  Here a slot is one SoA block of 14 x ld int32, field f of slot s at
 * base + (s*14 + f)*ld, field order = GRAAL_F_* below (same order as the reference struct).
 */
#ifndef GRAAL_B200_H
#define GRAAL_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct graal_ctx graal_ctx;

enum {
    GRAAL_F_POS = 0, GRAAL_F_ID_C, GRAAL_F_START_BP, GRAAL_F_LEN_BP, GRAAL_F_CIRC, GRAAL_F_ID,
    GRAAL_F_PREV, GRAAL_F_NEXT, GRAAL_F_L_CONT, GRAAL_F_L_CONT_BP, GRAAL_F_ORI, GRAAL_F_REP,
    GRAAL_F_ACTIV, GRAAL_F_ID_D, GRAAL_N_FIELDS
};

/* single structure mutations (graal_apply_move); aux = orientation (+1/-1) or `upstream` flag */
enum {
    GRAAL_OP_COPY = 0,      /* simple_copy          kernels3.cu:3755-3774 */
    GRAAL_OP_FLIP,          /* flip_frag            kernels3.cu:239-279   */
    GRAAL_OP_SWAP_ACTIV,    /* swap_activity_frag   kernels3.cu:283-326   */
    GRAAL_OP_POP_OUT,       /* pop_out_frag         kernels3.cu:329-563   */
    GRAAL_OP_POP_IN_1,      /* pop_in_frag_1        kernels3.cu:565-812   */
    GRAAL_OP_POP_IN_2,      /* pop_in_frag_2        kernels3.cu:814-1079  */
    GRAAL_OP_POP_IN_3,      /* pop_in_frag_3        kernels3.cu:1081-1265 */
    GRAAL_OP_POP_IN_4,      /* pop_in_frag_4        kernels3.cu:1267-1448 */
    GRAAL_OP_SPLIT,         /* split_contig         kernels3.cu:1451-1784 */
    GRAAL_OP_PASTE,         /* paste_contigs        kernels3.cu:1786-2070 */
    GRAAL_N_OPS
};

#define GRAAL_N_CANDIDATES 13   /* n_tmp_struct, cuda_lib_gl.py:111-112 */

/* ---- context -------------------------------------------------------------------------------- */
int  graal_ctx_create(int device, graal_ctx** out);
void graal_ctx_destroy(graal_ctx* ctx);
const char* graal_last_error(void);
int  graal_set_stream(graal_ctx* ctx, void* cuda_stream);       /* cudaStream_t; default: own stream */
int  graal_sync(graal_ctx* ctx);                                 /* graal_join + wait for the context stream */
/* Proposals scored by graal_score_proposal run on internal lanes (streams) beside the context stream, the
 * way the reference scores its candidates on 13 streams (cuda_lib_gl.py:2441-2538).  graal_join makes the
 * context stream wait for them WITHOUT blocking the host: call it before enqueueing your own work (e.g. the
 * D2H copy of d_out) on the context stream.  Every other entry point that touches the state joins by itself. */
int  graal_join(graal_ctx* ctx);
const char* graal_version(void);

/* ---- level (read-only inputs of the likelihood kernels) ---------------------------------------
 * Replaces the uploads of sampler.__init__ (cuda_lib_gl.py:123-130, 194, 210-215): collector_id,
 * dispatcher (int2[N]), id_sub_frags (int4[N]: 3 sub ids + count), len_bp_sub_frags (float3[N], kb),
 * accu_sub_frags (int3[N]) and the DENSE obsData2D, which becomes the upper triangle (row < col) of
 * the same symmetric, zero-diagonal matrix as row-segmented contact lists:
 * rowptr int64[W+1], contacts = E records {int32 col, float32 count}.
 * nfpb = mean_squared_frags_per_bin (simulation_loader.py:73). */
int graal_level_bind(graal_ctx* ctx, int n_frags, int n_new_frags, int n_sub_frags,
                     const int32_t* sub_id, const float* sub_len_kb, const int32_t* sub_accu,
                     const int32_t* collector, const int32_t* dispatcher,
                     const int64_t* rowptr, const void* contacts, int64_t n_contacts, float nfpb);

/* The preparation of the sub-level matrix in sampler.__init__ (cuda_lib_gl.py:153-172: csr + csr.T, diagonal zeroed) on the
 * device: n COO entries (rows, cols int32, counts float32; any triangle, duplicates allowed) -> the row-segmented contact
 * lists graal_level_bind takes: entries keyed (min, max), radix-sorted, duplicates summed, diagonal and zero counts dropped,
 * rows counted and prefix-summed into d_rowptr (int64[W + 1]); d_contacts must hold n records of 8 bytes; *n_contacts_out
 * (host) receives the number written.  Blacklisted rows (materialised with mean_value_trans) are added by the caller. */
int graal_coo_to_lists(graal_ctx* ctx, const int32_t* d_rows, const int32_t* d_cols, const float* d_vals, int64_t n,
                       int n_sub_frags, int64_t* d_rowptr, void* d_contacts, int64_t* n_contacts_out);

/* param_simu (kernels3.cu:26-35): kuhn, lm, c1, slope, d, d_max, fact, v_inter */
int graal_set_params(graal_ctx* ctx, const float p[8]);

/* How the expected value of an IN-BAND cis pixel (0 < s < d_max) is evaluated:
 *   0  the reference's float32 chain op for op: c1 * powf(s, slope) * expf((d-2)/(x*x+d)) * fact, max with
 *      v_inter, times norm, then log() of the float (kernels3.cu:120-133, 191-210)
 *   1  the same quantity in log space, float64: max(ln(c1*fact) + slope*ln(s) + (d-2)/(x*x+d),
 *      ln v_inter) + ln norm, from the same float32 distance s.  Differs from mode 0 by the float32
 *      roundings of the chain (<= ~1.5e-7 relative per pixel, zero mean), the order of the difference
 *      between two float32 libm implementations.
 *   2  (default) mode 1 with the law tabulated: f(s) and ln f(s) as piecewise quadratics, 512 intervals per
 *      octave of s over 2^-12 .. 2^11 kb, one 16-byte entry {double a0; float a1, a2} per interval (rebuilt on
 *      the host for every parameter set), indexed by the float bit pattern of s; interpolation error < 2e-9
 *      relative.  One gather, no log / exp / division per pixel; distances outside the table use mode 1.
 * Circular contigs always use the float32 chain (rippe_contacts_circ, kernels3.cu:135-166). */
int graal_set_math_mode(graal_ctx* ctx, int mode);

/* ---- state ---------------------------------------------------------------------------------- */
int graal_state_bind(graal_ctx* ctx, int32_t* slots_base, int ld, int n_slots);

/* relabel side effect of gl_update_pos (kernels3.cu:3848-3851) + host map of modify_gl_cuda_buffer
 * (cuda_lib_gl.py:1697-1722): ids become 0..n_contigs-1 by increasing contig length, ties by old id.
 * d_max_id (device int32, may be NULL) receives n_contigs-1; also kept inside the context. */
int graal_relabel_contigs(graal_ctx* ctx, int slot, int32_t* d_max_id);

/* one mutation kernel, src_slot -> dst_slot (persistent destination: unwritten bins keep their
 * content).  d_max_id_out (device int32, may be NULL) receives max(id_c) of the destination, i.e. the
 * ga.max(...) that follows pop_out / split in the reference (cuda_lib_gl.py:857,934,943). */
int graal_apply_move(graal_ctx* ctx, int src_slot, int dst_slot, int op, int id_fA, int id_fB,
                     int aux, int max_id_in, int32_t* d_max_id_out);

/* fused new_perform_modificationS (cuda_lib_gl.py:841-954,1045-1048): the 13 candidate structures of
 * (id_fA, id_fB) from src_slot into slots first_dst_slot .. first_dst_slot+12; bit m of mode_mask
 * enables candidate m.  max_id < 0: use the value left by the last graal_relabel_contigs. */
int graal_build_candidates(graal_ctx* ctx, int src_slot, int first_dst_slot, int id_fA, int id_fB,
                           int max_id, unsigned mode_mask);

/* copy_struct (kernels3.cu:3720-3742): commit slot src into slot dst. */
int graal_commit(graal_ctx* ctx, int dst_slot, int src_slot);

/* evaluate_likelihood (kernels3.cu:2802-3222) + ga.sum (cuda_lib_gl.py:629,1848): full
 * log-likelihood of a slot -> d_out[0] (device double).  p_override (host, 8 floats, may be NULL)
 * evaluates with test parameters (compute_likelihood_4_nuisance, cuda_lib_gl.py:1986-2019). */
int graal_full_loglik(graal_ctx* ctx, int slot, const float* p_override, double* d_out);

/* fill_sub_index_fA/fB (kernels3.cu:3225-3249) + sub_compute_likelihood (kernels3.cu:3259-3718)
 * for n_cand candidate slots against base_slot -> d_out[n_cand] (device doubles).
 * max_id: max(id_c) of base_slot (< 0: the value left by the last graal_relabel_contigs). */
int graal_delta_loglik(graal_ctx* ctx, int base_slot, int first_cand_slot, int n_cand,
                       int id_fA, int id_fB, int max_id, double* d_out);

/* stream_likelihood (cuda_lib_gl.py:2392-2546) for ONE proposal (id_fA, id_fB): the 13 candidates are
 * built into first_cand_slot.. and scored against base_slot -> d_out[13].  proposal_index (0..15) names the
 * proposal for a later graal_commit_scored.  Candidate 8 (swap activity) equals candidate 0 for unique
 * bins (reference quirk Q7) and is copied, not re-evaluated.
 * Concurrency: proposal_index selects a lane; proposals given DISJOINT candidate slot ranges overlap on the
 * GPU, proposals sharing one range are serialised (GRAAL_LANES=1 in the environment forces one lane).
 * d_out is valid after graal_join / graal_sync. */
int graal_score_proposal(graal_ctx* ctx, int base_slot, int first_cand_slot, int id_fA, int id_fB,
                         int max_id, int proposal_index, double* d_out);

/* The neighbour loop of step_max_likelihood (cuda_lib_gl.py:1866-1895) in one call: graal_score_proposal for
 * proposals x = 0..n_proposals-1 (id_fA, id_fB[x]; id_fB is a HOST array) -> d_out[13 * x + k], and, if d_dist
 * is not NULL, graal_dist_candidates -> d_dist[13 * x + k].  Proposal x is built into the candidate slots
 * first_cand_slot + 13 * (x % lanes) when the state block holds 13 * lanes candidate slots from first_cand_slot
 * on (the proposals then overlap on the GPU), else into first_cand_slot (serial). */
int graal_score_step(graal_ctx* ctx, int base_slot, int first_cand_slot, int id_fA, const int32_t* id_fB,
                     int n_proposals, int max_id, double* d_out,
                     const int32_t* init_prev, const int32_t* init_next, const int32_t* init_orientable,
                     const uint8_t* skip, double* d_dist);

/* graal_join + copy `bytes` from device to (pinned) host memory on the context stream + wait: the one device
 * round trip of a step (the reference's ga.sum / .get() calls, cuda_lib_gl.py:1848,1897). */
int graal_fetch(graal_ctx* ctx, const void* d_src, void* h_dst, size_t bytes);

/* test_copy_struct (cuda_lib_gl.py:1156-1183): rebuild candidate `mode` of (id_fA, id_fB) and commit it to
 * base_slot.  proposal_index >= 0: the proposal was scored by graal_score_proposal since base_slot last
 * changed, and the cached band total of base_slot is updated with that candidate's band delta (the next
 * graal_full_loglik then skips the band pass); < 0: plain rebuild + commit. */
int graal_commit_scored(graal_ctx* ctx, int base_slot, int first_cand_slot, int id_fA, int id_fB,
                        int max_id, int mode, int proposal_index);

/* per-step statistics of step_max_likelihood (cuda_lib_gl.py:1809-1816) -> d_out[4] (device doubles):
 * n_contigs, min l_cont, mean l_cont_bp over contig heads, max l_cont. */
int graal_state_stats(graal_ctx* ctx, int slot, double* d_out);

/* graal_state_stats followed by graal_relabel_contigs of the same slot (the head of every step, cuda_lib_gl.py:1801-1816),
 * fused into three launches when the context knows a bound on the contig ids (read back with graal_fetch after the previous
 * relabel: at most 4096 contigs); the two general sequences otherwise.  Same outputs as the two calls. */
int graal_stats_relabel(graal_ctx* ctx, int slot, double* d_stats_out, int32_t* d_max_id);

/* dist_inter_genome (cuda_lib_gl.py:475-541): d_out[0] = sum over the bins with skip[f] == 0 of their
 * distance term (3 minus the neighbour / orientation agreements with the initial genome); the caller
 * divides by 3 * (number of counted bins).  init_prev / init_next / init_orientable: int32[n] device
 * arrays (np_init_prev, np_init_next, np_init_orientable, cuda_lib_gl.py:226-233); the initial orientation
 * is +1 everywhere; skip: uint8[n], 1 for blacklisted bins and repeat copies. */
int graal_dist_genome(graal_ctx* ctx, int slot, const int32_t* init_prev, const int32_t* init_next,
                      const int32_t* init_orientable, const uint8_t* skip, double* d_out);

/* The same sum for the n_cand consecutive candidate slots first_cand_slot.. -> d_out[n_cand]: the distance
 * the genome WOULD have after committing each candidate, so that step_max_likelihood (cuda_lib_gl.py:1962)
 * needs no second device round trip after the draw.  proposal_index >= 0: the candidates are those of that
 * graal_score_proposal call and the work is queued on its lane; < 0: on the context stream. */
int graal_dist_candidates(graal_ctx* ctx, int first_cand_slot, int n_cand, int proposal_index,
                          const int32_t* init_prev, const int32_t* init_next,
                          const int32_t* init_orientable, const uint8_t* skip, double* d_out);

/* distance histogram of estimate_parameters (cuda_lib_gl.py:1236-1270) on the INITIAL sub-level
 * layout: for cis sub-frag pairs with mid-to-mid distance d < max_dist_kb, bin int(d/bin_kb):
 * d_sum[b] += contacts (zeros included through d_cnt), d_cnt[b] += 1.
 * sub_id_c / sub_start_bp / sub_len_bp / sub_pos: int32[W] device arrays (S_o_A_sub_frags). */
int graal_dist_histogram(graal_ctx* ctx, const int32_t* sub_id_c, const int32_t* sub_start_bp,
                         const int32_t* sub_len_bp, const int32_t* sub_pos,
                         double max_dist_kb, double bin_kb, int n_bins, double* d_sum, int64_t* d_cnt);

/* Host arithmetic of the candidate draw of step_max_likelihood (cuda_lib_gl.py:1899-1934) at temperature 1, float64, in
 * NumPy's operation order (pairwise sums included): score[n] (n = n_tmp x neighbours) -> id_ok[n_ok], the normalised weights
 * (work[0..n_ok)) and their normalised cumulative sums cdf[n_ok]; *id_max = argmax(score).  Returns n_ok, -1 if a weight is
 * not finite (the caller falls back to the NumPy statements).  No CUDA call. */
int graal_candidate_weights(const double* score, int n, int n_tmp, double* work, int32_t* id_ok, double* cdf, int32_t* id_max);

/* The same draw ON THE DEVICE, and the commit that follows it, without a host round trip (cuda_lib_gl.py:1899-1952):
 * score = d_delta[0 : 13 n_proposals] + likelihood (d_full[0], or likelihood_host when use_host_likelihood), the weights in
 * NumPy's operation order as in graal_candidate_weights, `u` = the NEXT uniform of the caller's RandomState -- it is consumed
 * by the draw only if more than one candidate is left (n_ok > 1): the caller advances its stream accordingly after the fetch.
 * d_sel (5 int32, device) = {sample_out, n_ok, status, id_f_sampled, op}; status 1 = NaN / non-finite weights: nothing is
 * drawn or committed, the host path decides (the reference raises there).  d_sub_score: >= 13 n_proposals doubles (the n_ok
 * normalised weights, `sampler.sub_score`); d_score: 4 doubles {score of the drawn candidate, sample_out, n_ok, status}.
 * graal_draw_commit then rebuilds candidate `op` of (id_fA, id_f_sampled) and copies it over base_slot like
 * graal_commit_scored (test_copy_struct, cuda_lib_gl.py:1156-1183), all enqueued on the context stream behind the lanes. */
int graal_draw_candidates(graal_ctx* ctx, const double* d_delta, const double* d_full, double likelihood_host, int use_host_likelihood,
                          int n_proposals, const int32_t* id_fB, double u, int32_t* d_sel, double* d_sub_score, double* d_score);
int graal_draw_commit(graal_ctx* ctx, int base_slot, int first_cand_slot, int id_fA, const int32_t* id_fB, int n_proposals, int max_id,
                      const double* d_delta, const double* d_full, double likelihood_host, int use_host_likelihood, double u,
                      int32_t* d_sel, double* d_sub_score, double* d_score, const void* d_fetch_src, void* h_fetch_dst, size_t fetch_bytes);
/* d_fetch_src / h_fetch_dst / fetch_bytes (optional): the D2H copy of the step's output block is enqueued between the draw and
 * the commit; graal_fetch_wait blocks until it has landed (the commit kernels may still be running) and does graal_fetch's
 * bookkeeping. */
int graal_fetch_wait(graal_ctx* ctx);

/* instrumentation: number of kernel launches issued by this context so far */
int64_t graal_launch_count(graal_ctx* ctx);

/* per-kernel device timers: CUDA event pairs recorded on the context's stream around the kernels
 * below (off by default).  graal_profile_read synchronises the stream and returns the accumulated
 * device time and launch count of one kernel id. */
enum {
    GRAAL_K_FULL_CONTACTS = 0,   /* contact-list pass of graal_full_loglik (the HBM-bound kernel) */
    GRAAL_K_FULL_BAND,           /* band-limited expected-mass pass of graal_full_loglik */
    GRAAL_K_DELTA_CONTACTS,      /* contact part of graal_delta_loglik */
    GRAAL_K_DELTA_BAND,          /* expected-mass part of graal_delta_loglik (both passes + reductions) */
    GRAAL_K_BUILD,               /* graal_build_candidates */
    GRAAL_K_RELABEL,             /* graal_relabel_contigs */
    GRAAL_K_FULL_WINDOWS         /* position order + per-row windows of the windowed contact pass */
};
int graal_profile_enable(graal_ctx* ctx, int on);
int graal_profile_read(graal_ctx* ctx, int kernel_id, double* total_ms, int64_t* count, int reset);
/* sizes of the proposals scored while the timers were on (what the delta kernels' roofline is computed from):
 * out[0] = stored contacts in the rows of U (contig(fA) + contig(fB)), [1] = rows (sub-frags) of U, [2] = bins of U,
 * [3] = proposals counted; summed over the proposals since the last reset. */
int graal_profile_counters(graal_ctx* ctx, int64_t out[4], int reset);

#ifdef __cplusplus
}
#endif
#endif
