/*
 * metamaps_b200.h -- C ABI of the B200-native MetaMaps compute core.
 *
 * The reference (DiltheyLab/MetaMaps) has no FFI; its hot path is reached through two C++
 * constructors and one function, all driven by `metamaps mapDirectly` / `metamaps classify`:
 *
 *   skch::Sketch::Sketch(Parameters&, size_t maxMemory, std::function<void(Sketch*,size_t)>*)
 *                                                   src/map/include/winSketch.hpp:157   (index build)
 *   skch::Map::Map(const Parameters&, const Sketch&, PostProcessResultsFn_t)
 *                                                   src/map/include/computeMap.hpp:85   (L1 + L2 mapping)
 *   mapWrap::addMappingQualities(...)               src/map/mapWrap.h:215               (mapping quality)
 *   meta::doEM(const Parameters&, const std::string&)  src/meta/fEM.h:466               (EM classification)
 *
 * Each entry point below replaces the array-level work of one of those seams; the file parsing,
 * text formatting and taxonomy book-keeping stay in the C++ host (metamaps_b200/csrc/host).
 *
 * Conventions
 *   - every function returns 0 on success, a negative MM_E* code otherwise; mm_last_error() returns a
 *     thread-local message for the last failure on the calling thread;
 *   - plain pointers + sizes; the caller owns every buffer it passes; the library owns handles until the
 *     matching *_destroy;  "host" pointers are ordinary host memory, "_dev" entry points take CUDA
 *     device pointers on the context's device;
 *   - an mm_ctx is used by one host thread at a time; work is issued on the context's CUDA stream and every entry
 *     point returns after its results are complete unless stated otherwise.  Several contexts may exist per GPU
 *     (e.g. two host threads, one context each, keeping two batches in flight); a finalized mm_index is read-only
 *     and may be mapped against from any context of its device;
 *   - there is no CPU implementation: if no CUDA device is usable, mm_ctx_create fails.
 *
 * Sequences are ASCII (any case, any byte; non-ACGT bytes are hashed verbatim like the reference does,
 * commonFunc.hpp:44-51,106), concatenated, with n+1 byte offsets.
 */
#ifndef METAMAPS_B200_H
#define METAMAPS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MM_OK 0
#define MM_EINVAL (-22)
#define MM_ENOMEM (-12)
#define MM_ECUDA (-5)
#define MM_ENODEV (-19)
#define MM_ERANGE (-34)

typedef struct mm_ctx mm_ctx;
typedef struct mm_index mm_index;

const char* mm_last_error(void);
const char* mm_version(void);

/* ---- context ------------------------------------------------------------------------------------ */
int mm_ctx_create(int device, mm_ctx** out);
void mm_ctx_destroy(mm_ctx* ctx);
/* Device-time (ms, CUDA events on the context stream) and launch count of the kernels issued by the last
 * mm_* call on this context; bench.py uses these for `gpu_launches` and the roofline block. */
int mm_ctx_last_timing(mm_ctx* ctx, double* total_ms, int64_t* n_launches);
/* Free / total device memory in bytes (a host sizes its index chunks with it, like the reference's --maxmemory estimate). */
int mm_ctx_mem_info(mm_ctx* ctx, int64_t* free_bytes, int64_t* total_bytes);
/* Per-stage device time (ms, CUDA events) of the last mm_map_* call:
 *   [0] K1 sketch  [1] K3 read sketch  [2] K4 probe+gather  [3] K4 hit sort  [4] K4 candidate regions  [5] K5 setup
 *   [6] K5a classify  [7] K5b sweep  [8] K5c strand  [9] accept + summary
 *   [10] the K1 kernel alone (SketchChunkFn, inside [0])  [11] the K5b sweep kernel alone (l2_sweep_band_kernel, inside [7])
 * and the algorithmic unit counters: [0] sketch elements s_total  [1] seed hits H  [2] candidates C
 *   [3] span minimizers sum N_c  [4] accepted mappings  [5] read minimizers  [6] bases of eligible reads
 *   [7] non-ACGT bases  [8] reads whose duplicate survivor needed the std::sort replay  [9] candidates swept by the
 *   fast (shared-memory) sweep  [10] seed hits kept by the L1 contig filter  [11] K5b work items (candidate segments). */
int mm_ctx_last_map_stats(mm_ctx* ctx, double* stage_ms /*[16]*/, int64_t* counters /*[16]*/);

/* ---- K1: winnowed minimizers  (replaces CommonFunc::addMinimizers, commonFunc.hpp:92-175) --------- */
/* Sketches n_seqs sequences; results stay on the device until fetched.  *n_total = number of minimizers. */
int mm_sketch_batch(mm_ctx* ctx, const char* seqs, const int64_t* offsets, int32_t n_seqs, int k, int w,
                    int64_t* n_total);
/* counts: n_seqs+1 prefix offsets; hash/wpos/strand: n_total entries, per sequence in emission order;
 * strand is +1 / -1 (base_types.hpp:126). */
int mm_sketch_fetch(mm_ctx* ctx, int64_t* counts, uint32_t* hash, int32_t* wpos, int32_t* strand);

/* ---- K2: reference index  (replaces skch::Sketch, winSketch.hpp:157-365,452-517) ------------------ */
int mm_index_create(mm_ctx* ctx, int k, int w, mm_index** out);
/* Appends contigs in DB.fa order; contig i of the j-th call gets the next sequence id (winSketch.hpp:258-345,
 * contigs shorter than w or k keep an id but contribute no minimizers). */
int mm_index_add(mm_index* idx, const char* seqs, const int64_t* offsets, int32_t n_seqs);
int mm_index_add_dev(mm_index* idx, const void* seqs_dev, const int64_t* offsets_host, int32_t n_seqs);
/* Sorts, builds the hash->positions table and the occurrence threshold (computeFreqHist, winSketch.hpp:452-495). */
int mm_index_finalize(mm_index* idx);
int mm_index_stats(const mm_index* idx, int64_t* n_minimizers, int64_t* n_unique, int32_t* freq_threshold,
                   int32_t* n_contigs, int64_t* device_bytes);
/* Contig-range shards.  A reference too large for one GPU (or spread over the GPUs of a node) is indexed as shards of
 * consecutive contigs, one mm_index each -- the reference's own chunking (winSketch.hpp:284-329), where chunk N's
 * sequence ids restart at 0 (:308,315): results carry shard-local contig ids, the caller adds first_contig_id.
 * L1/L2 work per contig, so a read's mappings against the whole reference are the concatenation, in shard order, of its
 * mappings against every shard -- provided the shards agree on which hashes are over-frequent.  Two choices:
 *   - per-shard thresholds (nothing to call): the reference's behaviour under --maxmemory, chunk by chunk;
 *   - the threshold of the UNCHUNKED reference: call mm_index_set_shard(.., keep_counts = 1) before mm_index_finalize
 *     on every shard, then mm_index_sync_threshold on every rank (collective over the context's communicator:
 *     mm_comm_init, or mm_comm_set_allreduce + mm_comm_set_rank).  The shards' (hash, count) lists are exchanged by
 *     hash range, merged, the global histogram gives the threshold and the over-frequent hashes are flagged locally. */
int mm_index_set_shard(mm_index* idx, int32_t first_contig_id, int32_t keep_counts);
/* The reference's own chunk loop (--maxmemory) does NOT reset its occurrence histogram or threshold between chunks
 * (winSketch.hpp:302-304,452-495): chunk N's threshold comes from the histograms of chunks 0..N summed, walked against chunk N's
 * unique count, and keeps chunk N-1's value when the first bucket overshoots.  A host that walks chunks reproduces that by
 * carrying the state: mm_index_get_freq_hist after chunk N-1's finalize (count_value[i] occurrences are shared by n_hashes[i]
 * hashes; *n_buckets entries, at most cap are written; *threshold = the chunk's threshold), mm_index_set_freq_carry before chunk
 * N's finalize -- or after it (the threshold is applied at probe time), which is what ranks building their chunks concurrently
 * do once the earlier chunks' histograms are known.  Without a carry an index computes the threshold of its own contigs, i.e. chunk 0's / the unchunked behaviour. */
int mm_index_set_freq_carry(mm_index* idx, const int32_t* count_value, const int64_t* n_hashes, int32_t n_buckets, int32_t prev_threshold);
int mm_index_get_freq_hist(const mm_index* idx, int32_t* count_value, int64_t* n_hashes, int32_t cap, int32_t* n_buckets, int32_t* threshold);
int mm_index_sync_threshold(mm_index* idx, int32_t* global_threshold, int64_t* global_unique);
/* Persistent index (replaces the Boost binary archive of skch::Sketch written by `metamaps index`, mapWrap.h:358-405).
 * A GPU-native dump: fixed header + the device arrays as they lie in HBM, so loading is a sequence of plain reads and
 * host->device copies (no re-sort, no re-hash).  Not byte-compatible with the reference's archives by design; the
 * `<prefix>.index` manifest contract (first line 1 = complete, then the chunk files) is kept by the host.
 * The file records k, w and the occurrence threshold; contig names are host metadata and are not stored here. */
int mm_index_save(const mm_index* idx, const char* path);
int mm_index_load(mm_ctx* ctx, const char* path, mm_index** out);
/* k / w of an index (e.g. after mm_index_load), contig lengths in id order (n_contigs entries, may be NULL). */
int mm_index_params(const mm_index* idx, int32_t* k, int32_t* w, int32_t* contig_len);
/* Largest number of reads one mm_map_* call can take against this index: seed hits are sorted as ONE 64-bit key
 * read | contig | position (computeMap.hpp:352's order in a single radix sort), so reads * contigs * longest contig must fit
 * 64 bits; a host flushes its read batches below this (a larger batch fails with MM_ERANGE). */
int mm_index_max_batch_reads(const mm_index* idx, int64_t* max_reads);
/* minimizerIndex in (seqId,wpos) order, for parity tests. */
int mm_index_fetch(const mm_index* idx, uint32_t* hash, int32_t* seq_id, int32_t* wpos, int32_t* strand);
/* minimizerPosLookupIndex probe, for parity tests: count (0 = absent) of each hash. */
int mm_index_lookup(const mm_index* idx, const uint32_t* hashes, int64_t n, int32_t* counts);
void mm_index_destroy(mm_index* idx);

/* ---- K3-K5: map a batch of reads  (replaces skch::Map::mapSingleQuerySeq, computeMap.hpp:228-538) -- */
typedef struct mm_map_params {
  float perc_identity;   /* --pi, default 80 (parseCmdArgs.hpp:345-354) */
  int32_t min_read_len;  /* -m; reads shorter than max(w,k,m) are skipped (computeMap.hpp:137) */
  int32_t report_all;    /* --all: 1 keeps every accepted mapping; 0 keeps, per read, those with identity >= best - 1.0
                            (reportReadMappings, computeMap.hpp:551-563): the others are dropped from the accepted set */
  int32_t reserved;
} mm_map_params;

typedef struct mm_map_summary {
  int64_t n_reads;       /* reads in the batch */
  int64_t n_too_short;   /* skipped by the length rule */
  int64_t n_candidates;  /* L1 candidate loci (all evaluated by L2) */
  int64_t n_mappings;    /* candidates passing the L2 identity filter (computeMap.hpp:415) */
  int64_t n_reads_mapped;
  int64_t total_bases_mapped_reads; /* bases of reads that were long enough (the Mbp of the metric) */
} mm_map_summary;

/* Host ASCII reads in.  Results stay in the context until the next mm_map_* call. */
int mm_map_batch(mm_ctx* ctx, const mm_index* idx, const char* reads, const int64_t* offsets, int32_t n_reads,
                 const mm_map_params* params, mm_map_summary* summary);
/* Device-resident ASCII reads (offsets on the host). */
int mm_map_batch_dev(mm_ctx* ctx, const mm_index* idx, const void* reads_dev, const int64_t* offsets_host,
                     int32_t n_reads, const mm_map_params* params, mm_map_summary* summary);

/* Contig-sharded ranks (collective over the context's communicator; the index is this rank's shard): every rank passes ITS block of
 * the batch -- blocks in rank order make up the batch, n_my_reads may differ between ranks -- and sketches only that block
 * (K0, K1, K3); the sorted unique sketches, sketch sizes and read lengths of all blocks are then all-gathered (ncclAllGather of
 * padded slabs) and every rank maps ALL reads of the batch against its shard (K4, K5).  Results are those of mm_map_batch_dev on
 * the concatenated reads: read indices run over the whole batch (rank 0's block first).  first_read = index of this rank's
 * first read in the batch (may be NULL). */
int mm_map_batch_sharded_dev(mm_ctx* ctx, const mm_index* idx, const void* my_reads_dev, const int64_t* my_offsets_host, int32_t n_my_reads,
                             const mm_map_params* params, mm_map_summary* summary, int64_t* first_read);

/* Double-buffered input staging (the pinned, double-buffered H2D of a streaming host).  mm_stage_reads_async starts the
 * upload of a batch of reads on the context's copy stream and returns at once; mm_map_batch_staged maps the batch staged
 * in `slot` (0 or 1) -- the upload is awaited on the device, so staging batch i+1 before mapping batch i overlaps the
 * transfer with the kernels.  reads_host in pinned (mapped) memory: no DMA at all, the 2-bit packing kernel reads the
 * bytes over PCIe itself; pageable memory: a feeder thread copies in pieces.  Either way reads_host must stay valid and
 * unchanged until mm_map_batch_staged has returned; offsets are copied. */
int mm_stage_reads_async(mm_ctx* ctx, int slot, const char* reads_host, const int64_t* offsets, int32_t n_reads);
int mm_map_batch_staged(mm_ctx* ctx, const mm_index* idx, int slot, const mm_map_params* params, mm_map_summary* summary);

/* Per-read: sketch size s (0 for skipped reads), minimumHits, offsets into the candidate arrays (n_reads+1).
 * Any pointer may be NULL. */
int mm_map_fetch_reads(mm_ctx* ctx, int32_t* sketch_size, int32_t* minimum_hits, int64_t* cand_offsets);
/* Per L1 candidate, in the reference's order (read order, then (seqId,wpos) order): contig id, candidate range,
 * L2 result: meanOptimalPos, sharedSketchSize, strand votes (sum strandQ*strandR), accepted flag (identity
 * upper bound >= pi), valid flag (0 when no window had a shared element; the reference leaves the position
 * uninitialised there), and the optimal window [opt_start,opt_end) as minimizerIndex positions. */
int mm_map_fetch_candidates(mm_ctx* ctx, int32_t* seq_id, int32_t* range_start, int32_t* range_end,
                            int32_t* mean_optimal_pos, int32_t* shared, int32_t* strand_votes,
                            int32_t* accepted, int32_t* valid, int64_t* opt_start, int64_t* opt_end);
/* Stage outputs for parity tests: the read sketch (sorted unique hashes + strand of the surviving minimizer). */
int mm_map_fetch_sketch(mm_ctx* ctx, int64_t* offsets /*n_reads+1*/, uint32_t* hash, int32_t* strand, int64_t cap);

/* The accepted mappings of the batch (what mapDirectly writes, computeMap.hpp:546-588 with --all), compacted on the device,
 * in the reference's order (read order, then (seqId, wpos)).  Per mapping: read index in the batch, contig id,
 * meanOptimalPos (refStart), sharedSketchSize, sketch size s, strand (+1/-1), nucIdentity (float, computeMap.hpp:405-408,
 * computed on the host through glibc like the reference) and identity_parsed = the double that `classify` re-reads from
 * the 6-significant-digit text of column 10 (mapWrap.h:229, fEM.h:297).  Any output pointer may be NULL; cap = capacity
 * of the arrays (>= summary.n_mappings). */
int mm_map_fetch_mappings(mm_ctx* ctx, int32_t* read_idx, int32_t* seq_id, int32_t* ref_start, int32_t* shared,
                          int32_t* sketch, int32_t* strand, float* identity, double* identity_parsed, int64_t cap,
                          int64_t* n_mappings);

/* ---- the classify stage on the device  (mapWrap::addMappingQualities + meta::doEM on arrays that stay in HBM) -------
 * The accepted mappings of the last mm_map_* call(s) go through identity -> mapping quality (K6) -> nLoc -> EM (K7/K8)
 * without leaving the device; the host reads the finished arrays once.  Sequence of calls per batch of reads:
 *   mm_classify_setup          once per reference: contig lengths and taxa (GLOBAL contig ids), n_taxa
 *   mm_classify_begin          empties the mapping table
 *   mm_classify_add_mappings   after each mm_map_* call of the batch: appends its accepted mappings, contig ids shifted by
 *                              first_contig_id (0 for an unsharded index; the shard's first contig otherwise).  Shards must be
 *                              added in contig order: a read's lines are its mappings per chunk, in chunk order
 *                              (unifyFiles, mapWrap.h:128-132).
 *   mm_classify_next_batch     several read batches in one table (a whole sample, one EM over it like `metamaps classify`): after
 *                              a batch's mm_classify_add_mappings call(s); the next batch's reads are numbered after it
 *   mm_classify_exchange       contig-sharded multi-GPU only (collective over the context's communicator): all ranks mapped
 *                              the same reads against their own shard; the tables are all-gathered (ncclAllGather of padded
 *                              slabs) and merged on the device in shard (= rank) order; this rank keeps the reads
 *                              [read_lo, read_hi) and finalises them (mapping quality needs a read's mappings from all shards,
 *                              mapWrap.h:226-278).
 *   mm_classify_run            identity (float + its 6-digit text round trip, bit-identical to glibc's: see mm_classify.h),
 *                              mapq, nLoc, EM to the reference's stopping rule (em_max_iter == 0), for exactly em_max_iter
 *                              rounds (> 0), or no EM at all (< 0); multi-rank contexts all-reduce the taxon sums every round.  Collective when n_ranks > 1.
 *   mm_classify_fetch          any pointer may be NULL.  Per mapping (cap >= n_mappings): the arrays of mm_map_fetch_mappings
 *                              plus mapq (column 14), taxon, nloc, posterior; per mapped read (n_reads_mapped entries):
 *                              mapped_read = its index in the batch, read_off (+1 entry), best = index of its first maximal
 *                              posterior, mapq_status; f[n_taxa]; ll_hist[min(em_iters, ll_cap)]. */
typedef struct mm_classify_summary {
  int64_t n_mappings;        /* mappings in the table this rank finalised */
  int64_t n_reads_mapped;    /* reads with at least one of them */
  int32_t em_iters;          /* EM rounds run */
  int32_t n_identity_fixups; /* identities settled by the host through glibc (see mm_classify.h; normally 0) */
  double em_ms;              /* device time of the EM rounds + final pass (CUDA events) */
  double classify_ms;        /* device time of the whole stage */
} mm_classify_summary;
int mm_classify_setup(mm_ctx* ctx, const int64_t* contig_len, const int32_t* contig_taxon, int32_t n_contigs, int32_t n_taxa);
int mm_classify_begin(mm_ctx* ctx);
int mm_classify_add_mappings(mm_ctx* ctx, int32_t first_contig_id, int64_t* n_total);
int mm_classify_next_batch(mm_ctx* ctx);
int mm_classify_exchange(mm_ctx* ctx, int32_t read_lo, int32_t read_hi, int64_t* n_total);
int mm_classify_run(mm_ctx* ctx, int32_t em_max_iter, mm_classify_summary* out);
int mm_classify_fetch(mm_ctx* ctx, int32_t* read_idx, int32_t* seq_id, int32_t* ref_start, int32_t* shared, int32_t* sketch, int32_t* strand,
                      float* identity, double* identity_parsed, double* mapq, int32_t* taxon, double* nloc, double* posterior, int64_t cap,
                      int32_t* mapped_read, int64_t* read_off, int64_t* best, int32_t* mapq_status, double* f, double* ll_hist, int32_t ll_cap);

/* ---- host helpers of the classify stage (C++/OpenMP, no device work) -------------------------------- */
/* Runs of equal values in a non-decreasing array (the mappings' read indices): group_value[g], group_off[g..g+1]; arrays of
 * capacity n and n+1.  These are the reads that have lines in the mappings file, and their line ranges. */
int mm_group_sorted(const int32_t* values, int64_t n, int32_t* group_value, int64_t* group_off, int64_t* n_groups);
/* nucIdentity and its 6-significant-digit round trip for n (shared, s) pairs. */
int mm_stat_identity_batch(const int32_t* shared, const int32_t* sketch, int64_t n, int k, float* identity,
                           double* identity_parsed);
/* getMappingLocations' nLoc (fEM.h:324-348): mappings of read r are [read_off[r], read_off[r+1]); read_len[r] its length;
 * taxon[m] := contig_taxon[seq_id[m]];  nloc[m] = sum over the contigs c of that taxon of
 * (len_c >= read_len ? len_c - read_len + 1 : [c is among this read's mapped contigs]). */
int mm_nloc_batch(const int32_t* seq_id, const int64_t* read_off, const int32_t* read_len, int64_t n_reads,
                  const int64_t* contig_len, const int32_t* contig_taxon, int32_t n_contigs, int32_t n_taxa,
                  int32_t* taxon, double* nloc);

/* ---- statistics tables the host needs for text output (replace Stat::*, map_stats.hpp) ------------ */
int mm_stat_min_hits_relaxed(int s, int k, float perc_identity);          /* estimateMinimumHitsRelaxed :142 */
int mm_stat_recommended_window(double pvalue, int k, int alphabet, float perc_identity, int len_query,
                               uint64_t len_reference);                    /* recommendedWindowSize :226 */
double mm_stat_estimate_pvalue(int s, int k, int alphabet, float perc_identity, int len_query, uint64_t len_reference);
void mm_stat_identity(int shared, int s, int k, float* nuc_identity, float* nuc_identity_upper); /* computeMap.hpp:405-411 */

/* ---- K6: mapping qualities  (replaces mapWrap::addMappingQualities, mapWrap.h:215-323) ------------- */
/* identity[m] = column 10 / 100; reads delimited by read_offsets (n_reads+1); read_len per read.
 * mapq[m] = binomial likelihood normalised per read.  status[r] = 0, or 1 when the likelihood sum is 0
 * (the reference asserts, mapWrap.h:298). */
int mm_mapq_batch(mm_ctx* ctx, const double* identity, const int32_t* shared, const int32_t* sketch,
                  const int32_t* read_len, const int64_t* read_offsets, int64_t n_reads, int k,
                  double* mapq, int32_t* status);

/* ---- K7/K8: EM  (replaces the loop of meta::doEM, fEM.h:491-661, and its final pass :693-716) ------ */
/* taxon[m] in [0,T); mapq[m] = column 14; nloc[m] = possible mapping locations of that taxon for that read
 * (fEM.h:324-348).  max_iter <= 0: run to the reference's stopping rule (fEM.h:636); max_iter > 0: run exactly
 * max_iter rounds (the reference has no iteration cap; used to benchmark a fixed round count).
 * Outputs: f[T]; posterior[m]; best[r] = index of the first maximal posterior of read r (getBestMapping :217);
 * ll_hist[0..min(n_iter,ll_cap)) log-likelihood per round; *n_iter rounds run. */
int mm_em_run(mm_ctx* ctx, const int32_t* taxon, const double* mapq, const double* nloc,
              const int64_t* read_offsets, int64_t n_reads, int32_t T, int32_t max_iter,
              double* f, double* posterior, int64_t* best, double* ll_hist, int32_t ll_cap, int32_t* n_iter);

/* Multi-GPU EM: reads are partitioned over ranks; each rank passes its own mappings and the per-round
 * taxon sums + log-likelihood are all-reduced over NCCL.  id_bytes: 128-byte ncclUniqueId from
 * mm_comm_unique_id on rank 0, distributed by the caller (e.g. torch.distributed broadcast). */
int mm_comm_unique_id(void* id_bytes_128);
int mm_comm_init(mm_ctx* ctx, int n_ranks, int rank, const void* id_bytes_128);
int mm_comm_destroy(mm_ctx* ctx);
/* Alternative transport for the same exchange: a caller-supplied sum-all-reduce over a HOST buffer of n doubles
 * (e.g. MPI_Allreduce or torch.distributed/gloo).  The (T+1)-double buffer is staged through the host each EM round.
 * Used when NCCL is not wanted and by the world-size-2 gloo tests; pass fn = NULL to remove it.  NCCL takes
 * precedence when both are set.  The callback returns 0 on success. */
typedef int (*mm_allreduce_fn)(double* buf, int64_t n, void* user);
int mm_comm_set_allreduce(mm_ctx* ctx, mm_allreduce_fn fn, void* user);
/* rank / size of the host transport (NCCL contexts get theirs from mm_comm_init) */
int mm_comm_set_rank(mm_ctx* ctx, int n_ranks, int rank);

#ifdef __cplusplus
}
#endif
#endif /* METAMAPS_B200_H */
