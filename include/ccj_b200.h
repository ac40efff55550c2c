/* ccj_b200 -- C ABI of the B200-native CCJ MFE fill + traceback.
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++ or torch types.  The reference
 * (mateog4712/CCJ) has no FFI; its "interface" for this path is the C++ class surface of
 * src/W_final.hh:18-71 (W_final(seq,dangle), double ccj(), std::string structure, params_),
 * src/pseudo_loop.hh:13-56 and src/s_energy_matrix.hh:16-68, driven by src/CCJ.cc:44-49,58-115.
 * Each entry point below names the reference code it replaces.  The C++ shells in
 * ccj_b200/csrc/{W_final,pseudo_loop,s_energy_matrix}.hh and the CCJ command line are written on
 * top of exactly these functions.
 *
 * Error convention: every function returns 0 on success, a negative ccj_error on failure (message
 * via ccj_last_error).  Nothing here calls exit(); the reference's exit()/message behaviour of a
 * fold is reported per sequence in ccj_result and re-created by the caller (ccj_render.hpp).
 * There is no CPU fallback: without a CUDA device ccj_ctx_create fails.
 */
#ifndef CCJ_B200_H
#define CCJ_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ccj_ctx ccj_ctx;

enum ccj_error {
    CCJ_ERR_CUDA = -1,      /* CUDA runtime failure (no device, out of memory, launch error) */
    CCJ_ERR_ARG = -2,       /* bad argument */
    CCJ_ERR_PARAMS = -3,    /* parameter file unreadable / malformed; reference: "Not a valid parameter file!" */
    CCJ_ERR_SEQUENCE = -4,  /* character outside GCAUT (src/CCJ.cc:23-36) or empty sequence */
    CCJ_ERR_TOO_LARGE = -5, /* a single sequence does not fit this GPU's memory */
    CCJ_ERR_STATE = -6      /* call order (no model loaded, no batch prepared, ...) */
};

/* Per-sequence outcome.  status: 0 = the reference would print "SEQ\nSTRUCT (E)" and return 0;
 * 1 = the reference would print `msg_id`'s message on stderr and exit(1); 2 = "NOT GOOD RESTR INTER"
 * and exit(0).  n_should_not_be_here = number of "Should not be here!" lines the reference prints
 * on stdout before anything else (src/W_final.cc:714-715). */
typedef struct ccj_result {
    int32_t energy_dcal; /* W[n]; energy = energy_dcal / 100.0 (src/W_final.cc:79) */
    int32_t status;
    int32_t n_should_not_be_here;
    int32_t msg_id;      /* prefix*256 + node type character, see ccj_render.hpp */
    int32_t aux_i, aux_j;
} ccj_result;

/* One context = one GPU, one stream, one energy model.  Replaces the process-global state of the
 * reference (ViennaRNA parameter globals, noGU, pair[][], PK penalty globals). */
int ccj_ctx_create(int device, ccj_ctx **out);
void ccj_ctx_destroy(ccj_ctx *ctx);
const char *ccj_last_error(const ccj_ctx *ctx);

/* vrna_params_load(file) + scale_parameters() + model_details.dangles = dangles + make_pair_matrix()
 * (src/CCJ.cc:80-99, src/W_final.cc:20-25).  noGU as src/CCJ.cc:77. */
int ccj_model_load(ccj_ctx *ctx, const char *par_file, int dangles, int no_gu);
/* The same for a parameter set linked into the library: "dna_mathews2004" replaces
 * vrna_params_load_DNA_Mathews2004() (src/CCJ.cc:88-90, the set the reference embeds as
 * src/ViennaRNA/static/misc/dna_mathews2004.hex), "rna_turner2004" is the reference's compiled-in default set
 * (src/ViennaRNA/params/default.c).  Sections a parameter file omits keep the values of the latter, as in the
 * reference. */
int ccj_model_load_embedded(ccj_ctx *ctx, const char *name, int dangles, int no_gu);

/* W_final::W_final + W_final::ccj for `nseq` independent sequences (src/CCJ.cc:44-49).
 *   seqs     : concatenated upper-case sequences over GCAU(T), no separators
 *   offsets  : nseq+1 offsets into seqs (sequence s is seqs[offsets[s] .. offsets[s+1]))
 *   results  : nseq entries
 *   pairs    : optional (may be NULL), same offsets: 1-based partner of each nucleotide, -1 = unpaired
 *              (minimum_fold::pair, src/h_struct.hh:9-18)
 *   structs  : optional (may be NULL), same offsets: dot-bracket characters (W_final::structure)
 * Sequences are processed in waves sized to the GPU's free memory. */
int ccj_fold_batch(ccj_ctx *ctx, const char *seqs, const int64_t *offsets, int nseq, ccj_result *results,
                   int32_t *pairs, char *structs);

/* ccj_fold_batch over several GPUs of one box from ONE process: nctx contexts (one per device, each with the same
 * model loaded), one host thread each, sequences dealt dynamically in chunks; no collective, sequences are independent
 * ("one host thread + context per GPU; results gathered on host").  Per-context times of the call are
 * left in ccj_last_fill_ms / ccj_last_traceback_ms of each context. */
int ccj_fold_batch_multi(ccj_ctx **ctxs, int nctx, const char *seqs, const int64_t *offsets, int nseq, ccj_result *results,
                         int32_t *pairs, char *structs);

/* The same work split for measurement: prepare (H2D + table allocation, one wave only), fill
 * (W_final::ccj's loops, src/W_final.cc:60-77), traceback (:84-103), fetch (D2H). */
int ccj_batch_prepare(ccj_ctx *ctx, const char *seqs, const int64_t *offsets, int nseq);
int ccj_batch_fill(ccj_ctx *ctx);
int ccj_batch_traceback(ccj_ctx *ctx);
int ccj_batch_fetch(ccj_ctx *ctx, ccj_result *results, int32_t *pairs, char *structs);
/* device time of the last ccj_batch_fill / ccj_batch_traceback (CUDA events on the context's stream) */
float ccj_last_fill_ms(const ccj_ctx *ctx);
float ccj_last_traceback_ms(const ccj_ctx *ctx);
/* kernels launched by the last ccj_batch_fill */
int ccj_last_fill_launches(const ccj_ctx *ctx);
/* how many sequences of length n fit one wave on this context's GPU right now */
int64_t ccj_wave_capacity(ccj_ctx *ctx, int n);
/* the CUDA stream the context launches on (cudaStream_t), for callers that time with their own events */
void *ccj_stream(ccj_ctx *ctx);

/* Table access for parity tests (Matrix4D::get / TriangleMatrix raw values of the prepared wave).
 * Export order: 4D: i=1..n, j=i..n, k=j+2..n, l=k..n (C(n+1,4) int16); 2D: i=1..n, j=i..n (int32).
 * table ids: enum ccj_table4 / ccj_table2 in ccj_b200/csrc/ccj_types.h (same order as oracle/ref_dump.cc). */
int ccj_export_table4(ccj_ctx *ctx, int seq_index, int table, int16_t *out, int64_t out_len);
int ccj_export_table2(ccj_ctx *ctx, int seq_index, int table, int32_t *out, int64_t out_len);
int64_t ccj_table4_len(int n);
int64_t ccj_table2_len(int n);
/* one entry with the reference getters' semantics: Matrix4D::get (INF for an invalid index,
 * src/matrices.hh:177-182) / the raw TriangleMatrix or node value of (i,j), 1<=i<=j<=n */
int ccj_table4_get(ccj_ctx *ctx, int seq_index, int table, int i, int j, int k, int l, int32_t *value);
int ccj_table2_get(ccj_ctx *ctx, int seq_index, int table, int i, int j, int32_t *value);
/* FNV-1a 64 over the exported values (uint16 / uint32 each), the hash oracle/ref_dump.cc prints */
int ccj_table4_hash(ccj_ctx *ctx, int seq_index, int table, uint64_t *hash, int64_t *finite, int32_t *min_value);
int ccj_table2_hash(ccj_ctx *ctx, int seq_index, int table, uint64_t *hash, int64_t *finite, int64_t *sum);

/* What the C++ class shells need beyond whole folds (ccj_b200/csrc/{W_final,pseudo_loop,s_energy_matrix}.hh; reference
 * surface src/pseudo_loop.hh:13-56, src/s_energy_matrix.hh:16-68):
 *   ccj_model_build   host only: scale_parameters() of the loaded set as the library's model blob (model_bytes must equal
 *                     ccj_model_bytes(); layout = struct ccj_model of ccj_b200/csrc/ccj_types.h); par_file "@name" =
 *                     an embedded set
 *   ccj_model_upload  use a (possibly caller-edited) model blob for the following folds (vrna_param_t* argument of the
 *                     reference's constructors)
 *   ccj_copy_table4_raw / ccj_copy_tables2_raw   one gap table (storage order ccj_layout_index) / all 2D tables
 *                     (CCJ_NT2 x ccj_stride2(n) int32, diagonal-major) of the filled wave, for host-side getters
 *   ccj_traceback_step   ONE node of the traceback, i.e. one call of pseudo_loop::backtrack / W_final::backtrack
 *                     (src/pseudo_loop.cc:861, src/W_final.cc:175): node = {i, j, k, l, type} as in seq_interval
 *                     (src/h_struct.hh:65-92); the nodes it pushes, in push order, 5 ints each; status3 = {status,
 *                     msg_id, "Should not be here!" count} of the sequence so far
 *   ccj_fetch_fold_state minimum_fold::pair / ::type of positions 0..n as the traceback steps left them */
int ccj_model_build(const char *par_file, int dangles, int no_gu, void *model_out, size_t model_bytes, char *err, size_t err_len);
int ccj_model_upload(ccj_ctx *ctx, const void *model, size_t model_bytes);
size_t ccj_model_bytes(void);
int ccj_copy_table4_raw(ccj_ctx *ctx, int seq_index, int table, int16_t *out, int64_t out_len);
int ccj_copy_tables2_raw(ccj_ctx *ctx, int seq_index, int32_t *out, int64_t out_len);
int ccj_traceback_step(ccj_ctx *ctx, int seq_index, const int32_t *node, int32_t *pushed, int32_t cap, int32_t *n_pushed,
                       int32_t *status3);
int ccj_fetch_fold_state(ccj_ctx *ctx, int seq_index, int32_t *pair, int8_t *type);

/* ---- One sequence whose gap tables exceed one GPU (BASELINE config 5) ------------------------------------------------------
 * The reference has no counterpart (it aborts for n >= 214, src/matrices.hh:159-160); what is sharded are the loops of
 * pseudo_loop::compute_energies (src/pseudo_loop.cc:69-132) driven by W_final::ccj (src/W_final.cc:58-77): row i of
 * every gap table belongs to rank (i-1) mod world, the 12 tables that are read across rows are exchanged with one
 * in-place ncclAllGather per finished level, P(i,l) with an ncclAllReduce(min) per span (ccj_b200/csrc/ccj_shard.cu).
 * One process per GPU: rank 0 calls ccj_shard_unique_id and hands the id to the others (e.g. torch.distributed
 * broadcast); every rank then creates its shard on its own ccj_ctx (whose energy model is used), prepares the same
 * sequence and calls ccj_shard_fill(&shard, 1, ms).  The traceback runs on one rank, which first opens the other
 * ranks' memory (CUDA IPC handles from ccj_shard_ipc_handle, gathered by the caller) to read their row-local tables
 * over NVLink.  An in-process group (count == world shards on one ccj_ctx, no unique id) runs the same loop with
 * device copies as collectives: the single-GPU test of the multi-rank logic.
 * ms4 = {whole fill, compute kernels, allgather, allreduce} device ms. */
typedef struct ccj_shard ccj_shard;
size_t ccj_shard_unique_id_bytes(void);
int ccj_shard_unique_id(void *id, size_t bytes);
int ccj_shard_create(ccj_ctx *ctx, int rank, int world, const void *unique_id, ccj_shard **out);
/* all ranks inside one process, one GPU each (ncclCommInitAll; peer access between the devices is enabled): out[world] */
int ccj_shard_create_group(ccj_ctx **ctxs, int world, ccj_shard **out);
/* the whole sharded fold of one sequence over the GPUs of `ctxs` from one process: group, prepare, fill, traceback */
int ccj_shard_fold(ccj_ctx **ctxs, int nctx, const char *seq, int n, ccj_result *result, int32_t *pairs, char *structs,
                   float *ms4);
void ccj_shard_destroy(ccj_shard *shard);
const char *ccj_shard_last_error(const ccj_shard *shard);
int64_t ccj_shard_bytes(int n, int world);   /* host only: one rank's device memory for a length-n sequence */
int ccj_shard_prepare(ccj_shard *shard, const char *seq, int n);
int ccj_shard_fill(ccj_shard **shards, int count, float *ms4);
/* per step s of the last fill, 4 floats: P kernel, allreduce, 2D + gap-table kernels, allgather (device ms); and the
 * bytes one rank contributes to the allgather of a level (host only) */
int ccj_shard_level_ms(ccj_shard *shard, float *out, int64_t out_len);
int64_t ccj_shard_level_bytes(int n, int world, int level);
size_t ccj_shard_ipc_bytes(void);
int ccj_shard_ipc_handle(ccj_shard *shard, void *handle, size_t bytes);
int ccj_shard_open_peers(ccj_shard *shard, const void *handles, size_t bytes);
int ccj_shard_link_local(ccj_shard *shard, ccj_shard **all, int count);
int ccj_shard_traceback(ccj_shard *shard, ccj_result *result, int32_t *pairs, char *structs, float *ms);
int ccj_shard_energy(ccj_shard *shard, int32_t *energy_dcal);
int ccj_shard_table4_hash(ccj_shard *shard, int table, uint64_t *hash, int64_t *finite, int32_t *min_value);
int ccj_shard_table2_hash(ccj_shard *shard, int table, uint64_t *hash, int64_t *finite, int64_t *sum);
/* host only: position of cell (i,j,k,l) in the sharded layout (ccj_b200/csrc/ccj_types.h) */
int64_t ccj_shard_layout(int n, int world, int i, int j, int k, int l, int32_t *owner, int32_t *level, int64_t *level_cells,
                         int64_t *level_base);

/* Host-only helpers (no GPU needed), used by the CPU test-suite:
 * ccj_model_text writes the scaled model in the "name idx... value" text form of
 * `oracle/_ref/ccj_ref_dump params` to `out_path` (par_file "@name" = an embedded set); ccj_layout_index is the storage offset of cell
 * (i,j,k,l) inside one 4D table (-1 for an invalid index). */
int ccj_model_text(const char *par_file, int dangles, int no_gu, const char *out_path, char *err, size_t err_len);
int64_t ccj_layout_index(int n, int i, int j, int k, int l);

/* Measurement helpers.
 * ccj_batch_fill_profiled: same work as ccj_batch_fill but launched kernel by kernel (no graph) with a
 * CUDA event pair around every launch; kernel_ms[0..5] = summed device time of the gap-table split-point
 * kernel (or the whole generic level kernel), K_P, K_2D, other, gap-table window kernel, gap-table assembly.
 * ccj_count_terms (host only): the algorithmic work of one fold as SURVEY.md 8d defines it:
 * out[0]=cells C(n+1,4), out[1]=split terms, out[2]=P terms, out[3]=interior-window terms this
 * sequence evaluates (can_pair-gated), all exact. */
int ccj_batch_fill_profiled(ccj_ctx *ctx, float *kernel_ms);
int ccj_count_terms(const char *seq, int n, int no_gu, int64_t *out);
/* ccj_measure_addmin_peak: the integer ceiling SURVEY.md 8d asks to be measured on the box.  Runs a register-only
 * micro-kernel on every SM and returns min-plus candidates ("add-min pairs") per second for the instruction form
 * variant 0: int32 VIADDMNMX; 1: sign-extension of a packed int16 + VIADDMNMX (the form the split-point kernel
 * applies to a record); 2: VIADDMNMX.S16x2, two int16 cells per instruction (the window kernels' form). */
int ccj_measure_addmin_peak(ccj_ctx *ctx, int variant, double *pairs_per_s);

/* number of CUDA devices visible to the process (0 without a driver) */
int ccj_device_count(void);

/* library / build identification */
const char *ccj_version(void);

#ifdef __cplusplus
}
#endif
#endif /* CCJ_B200_H */
