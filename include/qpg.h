/*
 * qpg.h -- C ABI of libqpg_sm100.so: the B200 (sm_100a) implementation of
 * QPGesture's inference hot path.
 *
 * The reference (YoungSeng/QPGesture @ 7dd5daa) is pure Python and has no FFI;
 * its boundary for this path is the Python call surface of
 *   codebook/Speech2GestureMatching/GestureKNN.py  (CodeKNN, wavvq_distances; the legacy GestureKNN class)
 *   codebook/models/{vqvae,bottleneck,encdec,resnet}.py (VQVAE encode/decode)
 *   codebook/PAE.py (Model.forward, pose2phase)  and  process/process_bvh.py (make_bvh_GENEA2020_BT, numeric half)
 * Each entry point below names the reference lines it replaces.  The Python
 * shims in qpgesture_b200/ keep those names/arguments and call this library
 * through ctypes; INTEGRATION.md shows the stub a reference maintainer adds.
 *
 * Conventions
 *  - every pointer is a caller-owned DEVICE pointer unless the name ends in
 *    _host; nothing is allocated or freed inside the library except through
 *    the explicit qpg_*_create/destroy pairs; no torch types, no argv/env reads
 *  - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *    all work is enqueued on it and NOT synchronised
 *  - return value: 0 = ok, negative = QPG_E_* ; qpg_last_error() gives a
 *    thread-local message for the last failing call
 *  - the distance / min-by-code tables use qpg_pair_t {double d; int64 id}:
 *    d is the best distance of the bin (1e3 when the bin is empty, as
 *    GestureKNN.py:668,709), id the smallest global window id attaining it
 *    (-1 when empty) -- i.e. the result of the reference's row-major
 *    strict-< scan (GestureKNN.py:686-689, :717-720)
 */
#ifndef QPG_H_
#define QPG_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QPG_OK 0
#define QPG_E_BADARG (-1)
#define QPG_E_CUDA (-2)
#define QPG_E_UNSUPPORTED (-3)

#define QPG_CODEBOOK_SIZE 512 /* constant.py:56 */
#define QPG_ROWS_PER_GROUP 8  /* packed database: rows per tile group   */
#define QPG_CHUNK 128         /* packed database: floats per tile chunk */
#define QPG_LEV_TOKENS 11     /* GestureKNN.py:58-60: 6 past + 5 future taps */
#define QPG_LEV_STRIDE 12     /* tokens are stored 12 per row (48 B)    */

typedef struct {
  double d;
  int64_t id;
} qpg_pair_t;

int qpg_version(void);
const char* qpg_last_error(void);
/* number of kernels this library has launched in the calling process */
uint64_t qpg_launch_count(void);
/* tuning hook for sweeps (0 = automatic): compute warps per CTA (8|12), ring depth, grid size, team size */
int qpg_tune_cosine(int compute_warps, int stages, int grid, int team);
/* consecutive passes of one scan call walk the row groups in alternating direction (default on) so that a
 * pass starts on the part of the table its predecessor left in L2; results do not depend on it */
int qpg_tune_cosine_alternate(int enable);

/* ---------------- packed window database ---------------------------------
 * Row-major float32 windows [W, D] are re-laid-out once per database into
 * [G = ceil(W/8)] [NC = ceil(D/128)] [8] [128] float32 (zero padded), so that
 * each (group, chunk) tile is one contiguous 4 KiB block for a single TMA
 * bulk copy, and each row's squared L2 norm is computed in float64.
 * Replaces the materialised feature table the reference scans at
 * GestureKNN.py:685 (wavlm_train_feat[j, k]) and :716 (context_train[j, k//8]).
 */
size_t qpg_packed_bytes(int64_t W, int D);
int qpg_pack_rows_f32(const float* rows, int64_t W, int D, float* packed, double* row_sqnorm,
                      void* stream);

/* ---------------- device-side feature stacking (database / query build) -------------------------
 * out[(n*n_rows + r)*6C + i*C + c] = interp[n, row_step*r + 2i, c], i < 6 (zero past T_out), where
 * interp = F.interpolate(wavlm^T, size=T_out, mode='linear', align_corners=True)^T reproduced bit for
 * bit (data_processing.py:255-274).  Database windows: n_rows = 26, row_step = T_out/30 (GestureKNN.py:
 * 671-690); query steps: n_rows = 8, row_step = 4*T_out/30 (:528,:565).  wavlm [n_seq, T_in, C] f32. */
int qpg_stack_wavlm_rows(const float* wavlm, int64_t n_seq, int T_in, int C, int T_out, int n_rows, int row_step,
                         float* out, void* stream);

/* ---------------- candidate distance, cosine, fused min-by-start-code ----
 * For each of Q query vectors q[Q, D] (float32) and every window w < W:
 *   dist = 0.5*|| q/|q| - x_w/|x_w| ||^2      (sklearn paired cosine distance,
 *          GestureKNN.py:685,716; evaluated as (a+b)/2 - <q,x>/(|q||x|) with
 *          float64 accumulation of float32 data, a,b = 1 for non-zero vectors)
 *   table[q][labels[w]] = lexicographic min of (dist, id_offset + w)
 * `table` [Q, 512] must have been initialised with qpg_table_init (or hold a
 * partial result to be refined, e.g. another shard of the same database).
 * One launch handles up to 8 queries per pass over the database; Q larger than
 * `queries_per_pass` is processed in ceil(Q/queries_per_pass) passes
 * (0 = let the library choose from D).  Replaces CodeKNN.search_audio_cands
 * (mode 'wavlm_feat', GestureKNN.py:666-691) and search_text_cands (:708-721).
 */
int qpg_table_init(qpg_pair_t* table, int64_t n_entries, void* stream);
int qpg_cand_cosine_minbycode(const float* packed, const double* row_sqnorm, const int32_t* labels,
                              int64_t W, int D, int64_t id_offset, const float* q, int Q,
                              qpg_pair_t* table, int queries_per_pass, void* stream);
/* Same, with an explicit team size S in {1,2,3,4,6} (0 = automatic): the D range of a row
 * group is split over S warps.  S fixes the float64 summation order of a dot product, so
 * callers that merge tables from several GPUs pass the SAME S on every rank: identical rows
 * then get bit-identical distances on every shard and exact ties keep the smallest id. */
int qpg_cand_cosine_minbycode_team(const float* packed, const double* row_sqnorm, const int32_t* labels,
                                   int64_t W, int D, int64_t id_offset, const float* q, int Q,
                                   qpg_pair_t* table, int queries_per_pass, int team_size, void* stream);

/* Fused variant: rows carry two feature blocks [D1 | D2] (both multiples of 128 floats, packed as ONE
 * table with qpg_pack_rows_f32 over D1+D2 columns); q is [Q, D1+D2].  One pass produces both tables:
 * table1 from the first block (row_sqnorm1), table2 from the second (row_sqnorm2).  Replaces one
 * search_audio_cands + one search_text_cands call per query step (GestureKNN.py:549-566). */
int qpg_cand_cosine2_minbycode(const float* packed, const double* row_sqnorm1, const double* row_sqnorm2,
                               const int32_t* labels, int64_t W, int D1, int D2, int64_t id_offset,
                               const float* q, int Q, qpg_pair_t* table1, qpg_pair_t* table2, void* stream);

/* ---------------- one-pass scan for up to 64 query steps (int8-sliced table, tcgen05 kind::i8) -------------
 * The hot path of the matcher: ALL query steps of a clip against the window table in ONE pass over HBM
 * (csrc/sliced_scan.cu explains the arithmetic).  The table is kept twice: the float32 tiles of
 * qpg_pack_rows_f32 (exact re-evaluation, single-query API) and an int8-"sliced" 31-bit fixed-point copy
 * whose rows are ordered by start code (bin order):
 *   order[pos]     = source row of sorted position pos (rows with a label outside [0,512) last)
 *   bin_start[c]   = first sorted position of start code c, c = 0..512
 * Replaces one search_audio_cands + one search_text_cands call per query step (GestureKNN.py:549-566,
 * :666-721) for every step of a clip at once; results are the same (distance, id) tables.  Distances of
 * bins that needed no decision are the filter value (within 2e-7 of the float64 one); ids and the rank
 * transform are exact.
 */
typedef struct {
  double sq;   /* |q|^2, float64, fixed summation order          */
  double g;    /* 2^eq / |q|   (0 for an all-zero query)         */
  double h;    /* query part of the error bound (integer units)  */
  int32_t ex;  /* eq: |q'| < 2^eq                                */
  int32_t pad;
} qpg_qinfo_t;
typedef struct {
  double lo, hi;  /* interval that contains the bin's exact minimum distance (lo == hi: exact) */
  int64_t id;     /* global window id of the only row that can attain it, -1 = empty bin       */
  int32_t n;
  int32_t flags;  /* bit 0: lo/hi are the exact float64 distance                               */
} qpg_bin_t;
typedef struct {
  const int8_t* db_slices; /* qpg_slice_rows_i8 output of this feature block   */
  const int8_t* q_slices;  /* qpg_slice_queries_i8 output (same n_pad)         */
  int64_t* sacc;           /* int64 [n_pad][ceil(W/128)*128], zeroed by caller */
  int32_t n_kblocks;       /* ceil(D/128)                                      */
  int32_t pad;
} qpg_sliced_seg_t;
typedef struct {
  const float* q;          /* float32 [Q, ldq]                                   */
  const int8_t* col_exp;   /* the table's column exponents (NULL if none)        */
  int8_t* q_slices;        /* out, 1024-byte aligned, qpg_sliced_query_bytes()   */
  qpg_qinfo_t* q_info;     /* out [Q]                                            */
  int64_t ldq;
  int32_t D;
  int32_t pad;
} qpg_slice_job_t;
/* one table (feature block) as the bins / resolve stages see it */
typedef struct {
  const float* packed;         /* float32 tile table of qpg_pack_rows_f32 (exact re-evaluation)          */
  const double* row_sqnorm;    /* its squared norms                                                       */
  const float* q;              /* float32 queries [nq, ldq]                                               */
  const qpg_qinfo_t* q_info;   /* [nq] from qpg_slice_queries_i8                                          */
  int64_t ldq;
  int32_t D;
  int32_t pad;
  int64_t* sacc;               /* bins: scan output [n_pad][ceil(W/128)*128]                              */
  const int32_t* bin_start;    /* bins: [513]                                                             */
  const void* row_info;        /* bins: from qpg_slice_rows_i8                                            */
  const int32_t* order;        /* bins: [W]                                                               */
  qpg_bin_t* bins;             /* bins: out;  resolve: in, part 0 (part p = + p*part_stride records);     */
  int64_t bins_qstride;        /*   record of (query q, code c) at bins[q*bins_qstride + c]; 0 means 512    */
  qpg_pair_t* table;           /* resolve: out [nq][512]                                                  */
  int32_t* ranks;              /* resolve: out [nq][512]                                                  */
  int32_t* qflags;             /* resolve: out [nq] (may be NULL)                                         */
} qpg_sliced_table_t;
size_t qpg_sliced_bytes(int64_t n_rows, int D);
size_t qpg_sliced_query_bytes(int D, int n_pad);
/* rows float32 [W, D] row-major; col_exp int8 [D] optional per-column power-of-two scaling (database rows
 * are multiplied by 2^-col_exp, queries by 2^+col_exp: the dot product is unchanged, outlier feature
 * dimensions stop dominating the fixed-point range); row_sqnorm float64 [W] indexed by SOURCE row;
 * row_info 16 bytes per sorted row (opaque).  slices must be 1024-byte aligned. */
int qpg_slice_rows_i8(const float* rows, int64_t W, int D, const int32_t* order, const int8_t* col_exp,
                      const double* row_sqnorm, int8_t* slices, void* row_info, void* stream);
/* the Q query steps of one pass, 1 or 2 feature blocks in one launch; n_pad in {16,32,48,64} >= Q */
int qpg_slice_queries_i8(const qpg_slice_job_t* jobs, int n_jobs, int Q, int n_pad, void* stream);
/* the pass itself: 1 or 2 feature blocks (audio, text) of the same W rows in one launch (stream-K over
 * row tiles x k-blocks, int64 global accumulation: exact and order independent) */
int qpg_sliced_scan_i8(const qpg_sliced_seg_t* segs, int n_segs, int64_t W, int n_pad, int nq, void* stream);
/* CUDA-core evaluation of the same integer sums from the same tile images (tests / debugging): queries
 * 0, q_stride, 2*q_stride, ... < nq only */
int qpg_sliced_scan_ref(const int8_t* db_slices, const int8_t* q_slices, int n_kblocks, int64_t W, int n_pad, int nq,
                        int q_stride, int64_t* sacc, void* stream);
/* per (table, query, start code) records from sacc; bins with more than one possible winner are re-evaluated in
 * float64 right here.  Source row r of this (sliced) shard is row r + row_base of packed / row_sqnorm and has
 * global window id id_offset + r.  consume != 0: every sacc entry is zeroed once read (the next pass then needs
 * no memset).  stats (optional, uint64[2]): [0] += rows re-evaluated here, [1] += bins decided in
 * qpg_sliced_resolve */
int qpg_sliced_bins(const qpg_sliced_table_t* tabs, int n_tabs, int64_t W, int nq, int64_t id_offset, int64_t row_base,
                    int consume, uint64_t* stats, void* stream);
/* merge n_parts record sets (row shards), decide cross-shard winners and overlapping bins in float64, emit the
 * final table, its stable rank transform (= qpg_rank512) and per-query flags (bit 0: exact tie between
 * non-empty bins, i.e. the reference's argsort order is platform defined there).  packed / row_sqnorm address
 * rows by (global id - first_id). */
int qpg_sliced_resolve(const qpg_sliced_table_t* tabs, int n_tabs, int n_parts, int64_t part_stride, int nq,
                       int64_t first_id, uint64_t* stats, void* stream);

/* ---------------- candidate distance, Levenshtein, fused min-by-code -----
 * tokens [W, 12] uint32 (11 used: g0*320+g1 per tap, GestureKNN.py:58-60),
 * q_tokens [Q, 12].  dist = unit-cost edit distance (Levenshtein.distance,
 * GestureKNN.py:67), exact integers stored as double in the table.
 * Replaces search_audio_cands mode 'wavvq_feat' + wavvq_distances 'combine'.
 */
int qpg_cand_lev_minbycode(const uint32_t* tokens, const int32_t* labels, int64_t W,
                           int64_t id_offset, const uint32_t* q_tokens, int Q, qpg_pair_t* table,
                           void* stream);
/* plain pairwise distances (wavvq_distances(ls1, ls2, 'combine'), GestureKNN.py:44-67) */
int qpg_lev_distance(const uint32_t* a_tokens, const uint32_t* b_tokens, int64_t n, int32_t* out,
                     void* stream);

/* Prefetch a device buffer into L2 (one prefetch.global.L2 per 128-byte line).  The matcher uses it on the
 * small tables of the sequential tail (pos_rank, phase_amp, code), which the streaming scans evict. */
int qpg_l2_prefetch(const void* ptr, size_t bytes, void* stream);

/* ---------------- merge of per-shard tables (multi-GPU / multi-part) -----
 * out[e] = lexicographic min over p < n_parts of parts[p][e].  Used after the
 * all-gather of per-rank tables when database rows are sharded across GPUs.
 */
int qpg_table_merge(const qpg_pair_t* parts, int n_parts, int64_t n_entries, qpg_pair_t* out,
                    void* stream);

/* ---------------- rank transform ------------------------------------------
 * ranks[q][c] = position of table[q][c].d in a stable ascending sort of the
 * 512 distances (ties -> lower code first); the reference's
 * np.array(dist).argsort().argsort() (GestureKNN.py:553,574) up to the order of
 * exact ties, which NumPy leaves platform-defined.
 */
int qpg_rank512(const qpg_pair_t* table, int Q, int32_t* ranks, void* stream);
/* same, plus qflags[q] = 1 when two non-empty bins of step q hold exactly the same distance (NumPy's order of
 * those is platform defined; typical for the integer Levenshtein distances of mode B) */
int qpg_rank512_ties(const qpg_pair_t* table, int Q, int32_t* ranks, int32_t* qflags, void* stream);

/* ---------------- sequential tail of search_code_knn ----------------------
 * One thread block per clip walks its n_seg*8 steps (GestureKNN.py:528-660,
 * flag set use_phase & use_aud & use_txt):
 *   combined_x[c] = (pos_rank[last][c] + 0.05*freq_rank[c]) + rank_x[q][c]
 *   c_a, c_t      = argmin (ties -> lower code)
 *   candidate     = window aud_table[q][c_a].id / txt_table[q][c_t].id
 *   phase cosine of [prev[-5:];head[:3]] vs [prev[-3:];head[:5]] (:636,:644),
 *   audio wins ties (:646); emit 4 codes code[j, m:m+4]; prev = window tail.
 * Segments chain as predict_code_from_audio does (:791,:800).
 * Shapes: pos_rank int32 [512,512]; freq_rank int32 [512]; code int32 [N,30];
 * phase_amp float32 [N,240,16]; aud_frame/txt_frame int32 [26] = int(k/398*240)
 * per window slot; seed_code int32 [n_clips]; seed_phase float32 [n_clips,8,16];
 * tables/ranks [n_clips*n_seg*8, 512]; id_base = global id of window 0 of `code`.
 * Outputs: codes int64 [n_clips, n_seg, 30]; vote int32 [n_clips, n_seg, 8]
 * (0 = audio, 1 = text); status int32 [n_clips] (0 ok, 1 = a chosen start-code
 * had no window: the reference raises IndexError at GestureKNN.py:631).
 */
int qpg_match_tail(const qpg_pair_t* aud_table, const qpg_pair_t* txt_table, const int32_t* aud_rank,
                   const int32_t* txt_rank, const int32_t* pos_rank, const int32_t* freq_rank,
                   const int32_t* code, int64_t n_seq, const float* phase_amp, const int32_t* aud_frame,
                   const int32_t* txt_frame, const int32_t* seed_code, const float* seed_phase,
                   int n_clips, int n_seg, int64_t* codes_out, int32_t* vote_out, float* phase_out,
                   int32_t* status_out, void* stream);

/* Same state machine, one launch per run of segments [seg_begin, seg_begin+seg_count): `state`
 * (float32 [n_clips, 132], caller owned) carries (last code, failure flag, previous phase) from one
 * launch to the next, so the tail of segment g can run on a side stream as soon as that segment's
 * tables exist while later segments are still being scanned.  Tables / ranks / outputs are indexed by
 * the ABSOLUTE step (clip*n_seg*8 + seg*8 + s); seeds are read only when seg_begin == 0. */
int qpg_match_tail_segments(const qpg_pair_t* aud_table, const qpg_pair_t* txt_table, const int32_t* aud_rank,
                            const int32_t* txt_rank, const int32_t* pos_rank, const int32_t* freq_rank,
                            const int32_t* code, int64_t n_seq, const float* phase_amp, const int32_t* aud_frame,
                            const int32_t* txt_frame, const int32_t* seed_code, const float* seed_phase,
                            int n_clips, int n_seg, int seg_begin, int seg_count, float* state,
                            int64_t* codes_out, int32_t* vote_out, float* phase_out, int32_t* status_out,
                            void* stream);

/* ---------------- sequential tail, split into a parallel lookup and a short walk (csrc/match_walk.cu) -------
 * qpg_match_lookup: for every query step q < Q and every possible previous code `last` the two arg-mins of
 * GestureKNN.py:540-555,:574-576 and what hangs off the chosen windows, as 32-byte entries [Q][512].
 * pos_rank_t is the TRANSPOSED pose rank table, int16 [512 c][512 last]; ranks/tables as produced by
 * qpg_rank512 / qpg_sliced_resolve; qflags_* (optional) per-step tie flags of qpg_sliced_resolve.
 * qpg_match_walk: follows the entries (GestureKNN.py:627-660, segments chained as :791,:800).  With `trans`
 * (int16 [n_clips*n_seg*8][1024] scratch, 16-byte aligned; n_seg <= 13) the phase pick is first evaluated for
 * EVERY reachable state in parallel and the sequential part is a shared-memory table walk (lowest latency, the
 * choice for a few clips); with trans == NULL one warp per clip walks directly (two dependent loads per step,
 * the choice for many clips).  Both give identical results.  codes_out is pre-filled with -1, so a clip that
 * stops early is recognisable.
 * status_out[clip]: bit 0 (1) = a chosen start code had no window (the reference raises IndexError at :631);
 *                   bit 1 (2) = the result depended on the order of exact ties, which the reference leaves to
 *                   NumPy's unstable argsort (this implementation: lower code first). */
int qpg_match_lookup(const qpg_pair_t* aud_table, const qpg_pair_t* txt_table, const int32_t* aud_rank,
                     const int32_t* txt_rank, const int16_t* pos_rank_t, const int32_t* freq_rank, const int32_t* code,
                     int64_t n_seq, const int32_t* aud_frame, const int32_t* txt_frame, const int32_t* qflags_a,
                     const int32_t* qflags_t, int Q, void* entries, void* stream);
int qpg_match_walk(const void* entries, const int32_t* code, const float* phase_amp, const int32_t* seed_code,
                   const float* seed_phase, int n_clips, int n_seg, int16_t* trans, int64_t* codes_out,
                   int32_t* vote_out, float* phase_out, int32_t* status_out, void* stream);
/* Per-(window, table) phase statistics, built once per database: for window w = 26*sequence + m and table x
 * (0 audio, 1 text) the 72 floats at stats[(2*w + x)*72] hold the squared norms and self-dots of the rows the
 * phase pick (GestureKNN.py:636,:644) takes from that window as a candidate (rows 0..4 behind its phase frame) and
 * as the previous winner (rows 27..31), plus the four rows of the cross term.  stats: 16-byte aligned,
 * n_seq * 26 * 2 * qpg_phase_stats_floats() floats.
 * qpg_match_walk_stats = qpg_match_walk whose transition pass (trans != NULL) reads that table: the float32 filter of
 * a state is then 64 multiply-adds by eight lanes (four states per warp) instead of two 128-element cosines by a
 * whole warp; undecided states take the same float64 pick.  Identical results. */
size_t qpg_phase_stats_floats(void);
int qpg_phase_stats(const float* phase_amp, int64_t n_seq, const int32_t* aud_frame, const int32_t* txt_frame,
                    float* stats, void* stream);
int qpg_match_walk_stats(const void* entries, const int32_t* code, const float* phase_amp, const float* phase_stats,
                         const int32_t* seed_code, const float* seed_phase, int n_clips, int n_seg, int16_t* trans,
                         int64_t* codes_out, int32_t* vote_out, float* phase_out, int32_t* status_out, void* stream);

/* ---------------- VQ codebook L2 argmin -----------------------------------
 * BottleneckBlock.quantise (codebook/models/bottleneck.py:120-126):
 *   dist[m][k] = fl32( fl32(|x_m|^2 - 2*<x_m,c_k>) + |c_k|^2 ),  idx = first argmin
 * x [M, D] float32, codebook [K, D] float32; idx_out int64 [M]; min_out float32 [M]
 * (may be NULL).  <x,c> is accumulated in float64 and rounded once to float32.
 */
int qpg_vq_argmin_f32(const float* x, const float* codebook, int64_t M, int D, int K, int64_t* idx_out,
                      float* min_out, void* stream);
/* Same result, tensor-core filter + exact re-evaluation: X * C^T on tcgen05 (TF32, qpg_conv1d_taps_tf32) into
 * `scratch` (float32 [M*K + K]), then only the codes whose distance could be minimal under the TF32 error bound are
 * evaluated with the arithmetic above (typically 2-6 of 512).  Needs D % 4 == 0, K % 16 == 0. */
int qpg_vq_argmin_fast(const float* x, const float* codebook, int64_t M, int D, int K, float* scratch,
                       int64_t* idx_out, float* min_out, void* stream);
/* BottleneckBlock.dequantise (bottleneck.py:128-130): out[m,:] = codebook[idx[m],:] */
int qpg_vq_dequantise_f32(const int64_t* idx, const float* codebook, int64_t M, int D, int K, float* out,
                          void* stream);

/* ---------------- 1-D convolution stacks of the VQ-VAE ---------------------
 * Channels-last activations [B, T, C] float32.  One generic "tap GEMM":
 *   out[b, t*out_stride + out_offset, co] = bias[co] + residual[...]
 *        + sum_{tap, ci} act(in[b, t*in_stride + tap_offset[tap], ci]) * w[tap][ci][co]
 * with act = ReLU when relu_in != 0 and zero for out-of-range input frames.
 * Covers nn.Conv1d (encdec.py:20,24,39,113; resnet.py:33-36) and, as two
 * output phases, nn.ConvTranspose1d k4 s2 p1 (encdec.py:45).
 * w is [n_taps, Cin, Cout] float32 (repacked from torch's [Cout, Cin, k]).
 * precision: must be 0 (float32 FFMA, the index-parity mode); the tensor-core path is qpg_conv1d_taps_tf32.
 */
typedef struct {
  int B, T_in, T_out_total, C_in, C_out;
  int n_taps;
  int tap_offset[4];
  int in_stride;   /* input frame step per output index               */
  int out_stride;  /* output frame step (2 for transposed-conv phases) */
  int out_offset;  /* first output frame written                      */
  int n_out;       /* output indices t = 0..n_out-1 per batch item    */
  int relu_in;
  int precision;
} qpg_conv_desc_t;
int qpg_conv1d_taps_f32(const qpg_conv_desc_t* desc, const float* in, const float* w, const float* bias,
                        const float* residual, float* out, void* stream);

/* ---------------- tensor-core path of the same tap GEMM (tcgen05, kind::tf32) -------------
 *   out[b, t, n] = bias[n] + residual[b, t, n]
 *                + sum_tap sum_{k<C_in} in[b, t + row_offset[tap], chan_offset[tap] + k] * w[tap][n][k]
 * `in` is a view [B, T_view, C_view] (float32, channels-last; C_view a multiple of 4): stride-2
 * convolutions and the two phases of ConvTranspose1d(k4,s2,p1) are written on the paired-frame
 * view [B, T/2, 2C] with channel offsets.  Rows outside [0, T_view) read as zero (conv padding).
 * w is [n_taps][N_pad][K_pad] float32, zero padded (N_pad a multiple of BN, K_pad of 4).
 * Output element (b, t, n) lives at out[(b*out_rows_per_item + t)*out_ld + out_chan_offset + n];
 * `residual` uses the same addressing.  `out` and `out_relu` (= max(out, 0)) may each be NULL.
 * Operands are rounded to TF32 by the tensor cores (about 1e-3 relative): fast mode, not the
 * index-parity mode (that is qpg_conv1d_taps_f32, precision 0).
 */
typedef struct {
  int B, T_view, C_view;
  int n_out;
  int C_in, C_out;
  int K_pad, N_pad, BN;
  int n_taps;
  int row_offset[4];
  int chan_offset[4];
  int out_rows_per_item, out_ld, out_chan_offset;
} qpg_conv_tc_desc_t;
int qpg_conv1d_taps_tf32(const qpg_conv_tc_desc_t* desc, const float* in, const float* w, const float* bias,
                         const float* residual, float* out, float* out_relu, void* stream);
/* Same tap GEMM with float32-accurate products on the TF32 tensor cores (3xTF32): x*w ~ x_hi*w_hi + x_lo*w_hi +
 * x_hi*w_lo.  w_hi = w with the low 13 mantissa bits cleared, w_lo = (w - w_hi) with its low 13 bits cleared (both
 * in the layout of `w` above); the activations are split inside the kernel.  BN <= 128.  Products are accurate to
 * ~2^-21 relative (float32 FFMA: 2^-24), which is what the index-parity mode of the VQ-VAE needs. */
int qpg_conv1d_taps_3xtf32(const qpg_conv_tc_desc_t* desc, const float* in, const float* w_hi, const float* w_lo,
                           const float* bias, const float* residual, float* out, float* out_relu, void* stream);

/* ---------------- periodic auto-encoder inference: the producer of the database's phase column ----------------
 * Reference: codebook/PAE.py:50-162 (Model.forward, eval mode), :477-508 (pose2phase).  BatchNorm (eval) and the
 * convolution / linear bias are folded by the caller into one (scale, shift) pair per output channel:
 *   y = act(scale[o] * sum + shift[o]).
 *
 * qpg_pae_sliding_conv1: first convolution + BatchNorm + tanh for ALL T windows of pose2phase at once.
 *   vel_pad [T + 238, C] float32 is the zero-padded frame-difference sequence (PAE.py:481-482); window i is rows
 *   i .. i+238 behind one zero frame (:491-499).  w [O, C, K] (conv1.weight).  h1 [T, O, K + 1] receives
 *   tanh(bn(conv1(window_i))).  Shared work between overlapping windows is computed once (float64 prefix sums over
 *   the kernel taps); built for K = 240, O = 15, C <= 135.
 * qpg_pae_conv1d: plain batched Conv1d (cross-correlation, zeros padding, stride 1): x [B, Ci, Lin], w [Co, Ci, K],
 *   out [B, Co, Lin + 2 pad - K + 1]; act 0 = none, 1 = tanh.  K a multiple of 4, at most 256.
 * qpg_pae_params: latent [B, E, T] -> params [B, 4, E] = (phase, frequency, amplitude, offset) per channel
 *   (PAE.py:99-114 from the power spectrum without the DC bin, float64 DFT; :132-136 phase = atan2 of the folded
 *   Linear(T, 2) outputs with the model's own atan2 :92-97, in turns).  fcw [E, 2, T], fc_scale / fc_shift [E, 2],
 *   freqs [T / 2] and time_scale as the model holds them (:60-65).
 */
int qpg_pae_sliding_conv1(const float* vel_pad, const float* w, const float* scale, const float* shift, int T, int C,
                          int O, int K, float* h1, void* stream);
int qpg_pae_conv1d(const float* x, const float* w, const float* scale, const float* shift, int B, int Ci, int Lin,
                   int Co, int K, int pad, int act, float* out, void* stream);
int qpg_pae_params(const float* latent, const float* fcw, const float* fc_scale, const float* fc_shift,
                   const float* freqs, float time_scale, int B, int E, int T, float* params, void* stream);

/* ---------------- legacy pose-feature matcher: the `GestureKNN` class (GestureKNN.py:70-284) ----------------
 * One 8-frame step of `search_motion` (metric 0) or `search_fake_motion` (metric 1) for a BATCH of clips:
 * qpg_legacy_candidates (frame candidates per database sequence + rank transforms), qpg_legacy_pick (desired_k-th of
 * the rank-sum order; or the caller picks with np.argsort to reproduce NumPy's order among tied sums),
 * qpg_legacy_gather.
 *   feat   [n_seq, n_frames, F] float64 normalised features (audio part first, then the pose part), motion
 *          [n_seq, n_frames, J] float64, mask [n_seq, n_frames] int32 (control mask).
 *   query  [n_clips, q_ld]: metric 0 = pose feature of the previous frame (n_body values, L2, :169-171),
 *          metric 1 = audio feature of the current frame (n_aud values, sklearn cosine, :258);
 *   aud_query [n_clips, n_aud] (metric 0 only): audio feature of the current frame (:125-132).
 *   desired_k [n_clips]; j0 = first output column of this step; pred [n_clips, J, out_frames] float64.
 *   next_pose [n_clips, n_body] (nullable): pose feature of the last copied frame = the next step's query (:146).
 *   chosen_log [n_clips, n_steps, 2] (nullable): (sequence, frame) picked at step_idx.
 *   status [n_clips], caller-zeroed: bit 0 = fewer than desired_k + 1 candidates (IndexError at :144 in the
 *          reference; the clip stops), bit 1 = an exact tie (NumPy's unstable argsort decides in the reference).
 *   scratch: cands n_clips * n_seq * qpg_legacy_cand_bytes(), comb / tie_flag [n_clips, n_seq] int32, chosen
 *          [n_clips, 2] int32, n_found [n_clips] int32.
 */
size_t qpg_legacy_cand_bytes(void);
/* frame candidates per (clip, sequence) + rank sums: cands, comb [n_clips, n_seq] (-1 = no candidate), tie_flag */
int qpg_legacy_candidates(const double* feat, const int32_t* mask, int n_seq, int n_frames, int F, int n_aud, int n_body,
                          int step_sz, int n_clips, int metric, const double* query, int q_ld, const double* aud_query,
                          void* cands, int32_t* comb, int32_t* tie_flag, void* stream);
/* stable device pick: chosen [n_clips, 2] = (sequence, frame) at position desired_k of the order by (comb, sequence) */
int qpg_legacy_pick(const void* cands, const int32_t* comb, const int32_t* tie_flag, int n_seq, int n_clips,
                    const int32_t* desired_k, int32_t* chosen, int32_t* n_found, int32_t* status, void* stream);
/* copy the chosen frames into pred, feed the pose feature back, log the choice; n_found <= desired_k -> status bit 0 */
int qpg_legacy_gather(const double* feat, const double* motion, int n_seq, int n_frames, int F, int J, int n_aud,
                      int n_body, int step_sz, int n_clips, const int32_t* chosen, const int32_t* n_found,
                      const int32_t* desired_k, int j0, int out_frames, int step_idx, int n_steps, double* pred,
                      double* next_pose, int32_t* chosen_log, int32_t* status, void* stream);

/* ---------------- post-decode pose processing: the numeric half of make_bvh_GENEA2020_BT ----------------
 * (process/process_bvh.py:57-77).  qpg_savgol15_f64: Savitzky-Golay filter (window 15, order 2, scipy mode "interp")
 * of every column of x [T, C] float32 over time -> out [T, C] float64.  coef [15]: scipy.signal.savgol_coeffs(15, 2);
 * edge_first / edge_last [7, 15]: rows that evaluate the quadratic fitted to the first / last 15 frames at the
 * first / last 7 frames (built by the host shim).  T >= 15.
 * qpg_rotmat_to_euler_zxy: mats [N, 9] row-major 3x3 -> euler_deg [N, 3] = Rotation.from_matrix(m).as_euler('ZXY',
 * degrees=True): matrices that are not orthogonal (scipy's isclose test) are replaced by their orthogonal polar
 * factor U V^T first.  flags [N] (nullable): bit 0 = non-positive determinant (scipy raises ValueError), bit 1 =
 * gimbal lock (third angle set to zero, as scipy does with a warning). */
int qpg_savgol15_f64(const float* x, int T, int C, const double* coef, const double* edge_first, const double* edge_last,
                     double* out, void* stream);
int qpg_rotmat_to_euler_zxy(const double* mats, int64_t N, double* euler_deg, int32_t* flags, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* QPG_H_ */
