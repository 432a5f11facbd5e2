/* hibag_b200.h -- C ABI of the B200-native HIBAG scoring path (libhibag_b200.so).
 *
 * Drop-in boundary: HIBAG's GPU-extension plugin surface `struct TypeGPUExtProc`
 * (reference inst/include/LibHLA_ext.h:357-388) -- ten plain function pointers that the
 * reference host code calls from src/LibHLA.cpp:1014,1916,1938,1961,2258,2264,2290,2433,2500,2527
 * when HLA_LIB::GPUExtProcPtr is set (src/LibHLA.cpp:193, installed by src/HIBAG.cpp:559-573).
 * `hibag_b200_get_procs()` returns a pointer to a byte-compatible struct whose hooks run on
 * the GPU; everything else in this header is the batched / multi-GPU extension surface that the
 * legacy struct cannot express and that this repo's own host driver
 * (the restatement of CVariableSelection / CAlg_EM / CAttrBag_Model) uses.
 *
 * All entry points are extern "C", take plain pointers and sizes, and never let a C++
 * exception or CUDA error cross the boundary: int-returning functions return 0 on success and
 * non-zero on failure, with the message available from hibag_b200_last_error().  The hook
 * members of the plugin struct follow the reference's own convention instead (no return codes;
 * they throw std::exception, which src/HIBAG.cpp:42-60 converts to an R error).
 *
 * There is no CPU fallback: every compute entry point fails if no CUDA device is usable.
 */
#ifndef HIBAG_B200_H
#define HIBAG_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- packed operand types (binary-compatible with the reference) -------------------- */

/* THaplotype, reference inst/include/LibHLA_ext.h:261-299 (32 bytes) */
typedef struct hibag_haplotype {
	int64_t packed[2];      /* bit i = allele of SNP i (little-endian bit order) */
	double  freq;           /* haplotype frequency */
	float   freq_f32;       /* aux.a2.Freq_f32 (unused by this implementation) */
	int32_t hla_allele;     /* aux.a2.HLA_allele: 0-based, non-decreasing along a list */
} hibag_haplotype;

/* TGenotype, reference inst/include/LibHLA_ext.h:311-352 (48 bytes) */
typedef struct hibag_genotype {
	int64_t snp1[2];        /* PackedSNP1 */
	int64_t snp2[2];        /* PackedSNP2; (s1,s2): 0->(0,0) 1->(1,0) 2->(1,1) missing->(0,1) */
	int32_t bootstrap_count;
	int32_t allele1;        /* aux_hla_type.Allele1 <= Allele2 */
	int32_t allele2;
	int32_t aux_temp;
} hibag_genotype;

/* TypeGPUExtProc, reference inst/include/LibHLA_ext.h:358-388 (member order is the ABI) */
typedef struct hibag_gpu_ext_proc {
	void      (*build_init)(int n_hla, int n_sample);                       /* :361 */
	void      (*build_done)(void);                                          /* :363 */
	void      (*build_set_bootstrap)(const int oob_cnt[]);                  /* :365 */
	uint32_t *(*build_haplomatch)(const hibag_haplotype haplo[], const size_t n_haplo[],
	              int n_snp, const hibag_genotype geno[], size_t *out_n);   /* :367-368 (size_t&) */
	void      (*build_set_haplo_geno)(const hibag_haplotype haplo[], int n_haplo,
	              const hibag_genotype geno[], int n_snp);                  /* :370-371 */
	int       (*build_acc_oob)(void);                                       /* :373 */
	double    (*build_acc_ib)(void);                                        /* :375 */
	void      (*predict_init)(int n_hla, int n_classifier,
	              const hibag_haplotype *const p_haplo[], const int n_haplo[],
	              const int n_snp[]);                                       /* :381-382 */
	void      (*predict_done)(void);                                        /* :384 */
	void      (*predict_avg_prob)(const hibag_genotype geno[], const double weight[],
	              double out_prob[], double out_match[]);                   /* :386-387 */
} hibag_gpu_ext_proc;

#define HIBAG_B200_MAX_SNP   128      /* HIBAG_MAXNUM_SNP_IN_CLASSIFIER, LibHLA_ext.h:223 */
#define HIBAG_B200_NA        INT32_MIN /* NA_INTEGER */

/* ---- library state ------------------------------------------------------------------------ */

const char *hibag_b200_version(void);
const char *hibag_b200_last_error(void);
/* number of usable CUDA devices (0 if none; never fails) */
int hibag_b200_device_count(void);
/* select the CUDA device used by subsequent calls from this process (default 0) */
int hibag_b200_set_device(int device);
/* name / SM count / clock of the selected device */
int hibag_b200_device_info(char *name, int name_len, int *sm_count, int *clock_khz);

/* ---- the drop-in plugin ----------------------------------------------------------------------
 * Pointer to the ten hooks; pass it where the reference expects a TypeGPUExtProc*
 * (8th argument of HIBAG_NewClassifiers, src/HIBAG.cpp:601; attr(cl,"proc_ptr"),
 * R/HIBAG.R:707). build_haplomatch is NULL (optional hook, the reference then uses its CPU
 * search, src/LibHLA.cpp:1074). */
hibag_gpu_ext_proc *hibag_b200_get_procs(void);
/* with_haplomatch != 0: the same hooks plus build_haplomatch (reference src/LibHLA.cpp:1014-1072,
 * replacing the CPU search _PrepHaploMatch_*, :1569-1637). A host given this hook builds its
 * haplotype-pair lists in record order rather than in its CPU scan order, so its EM sums -- and
 * the trained frequencies -- differ from the CPU path in the last bits (true of any plugin
 * that provides the hook); hibag_b200_get_procs() therefore leaves it NULL. */
hibag_gpu_ext_proc *hibag_b200_get_procs_ex(int with_haplomatch);
/* the body of that hook on host arrays: records (in-bag index, (i2 << 16) | i1) of the haplotype
 * pairs at minimum distance for every sample with bootstrap_count > 0. *out_buf is malloc'd
 * (release with hibag_b200_free): out_buf[0] = 2 * n_records, then the records; *out_n = number
 * of uint32 in the buffer. n_haplo[n_hla] = haplotypes per allele (LenPerHLA). */
int hibag_b200_haplomatch(const hibag_haplotype *haplo, const size_t *n_haplo, int n_hla, int n_snp,
	const hibag_genotype *geno, int n_samp, uint32_t **out_buf, size_t *out_n);
void hibag_b200_free(void *p);

/* ---- stateless batched scoring (kernel-level entry points; host buffers) ----------------------
 * haplo[]: n_haplo records grouped by allele (hla_allele non-decreasing), geno[]: n_geno packed
 * genotypes for the same n_snp SNPs. */

/* BestGuess per genotype, reference src/LibHLA.cpp:1639-1704; NA -> HIBAG_B200_NA */
int hibag_b200_best_guess(const hibag_haplotype *haplo, int n_haplo, int n_hla, int n_snp,
	const hibag_genotype *geno, int n_geno, int32_t *out_a1, int32_t *out_a2);
/* PostProb of each genotype's own (allele1, allele2), reference src/LibHLA.cpp:1706-1767 */
int hibag_b200_post_prob(const hibag_haplotype *haplo, int n_haplo, int n_hla, int n_snp,
	const hibag_genotype *geno, int n_geno, double *out);
/* PostProb2: normalised posterior [n_geno][n_hla*(n_hla+1)/2] + raw sums,
 * reference src/LibHLA.cpp:1769-1830 */
int hibag_b200_post_prob2(const hibag_haplotype *haplo, int n_haplo, int n_hla, int n_snp,
	const hibag_genotype *geno, int n_geno, double *out_prob, double *out_sum);

/* ---- model container + training + prediction (own host driver) -------------------------------- */

typedef struct hibag_b200_model hibag_b200_model;

/* reference CAttrBag_Model::InitTraining, src/LibHLA.cpp:2196; geno is int8 [n_samp][n_snp]
 * sample-major (0/1/2, anything else missing); h1/h2 0-based allele indices */
hibag_b200_model *hibag_b200_model_new(int n_snp, int n_hla);
void hibag_b200_model_free(hibag_b200_model *m);
int hibag_b200_model_set_training(hibag_b200_model *m, int n_samp, const int8_t *geno,
	const int32_t *h1, const int32_t *h2);

/* training options */
typedef struct hibag_b200_train_opts {
	int      nclassifier;
	int      mtry;              /* candidates per selection round, R/HIBAG.R:180-208 */
	int      prune;             /* src/LibHLA.cpp:2057 */
	int      n_threads;         /* host threads for candidate-parallel EM (<=0: all cores) */
	int64_t  seed;              /* R set.seed() value */
	int      per_classifier_seed; /* 0: one RNG stream over classifiers (what R does): the model is
	                                 seeded when it first sees this seed value; a later train call with
	                                 the same seed CONTINUES the stream (R: set.seed once, then repeated
	                                 hlaAttrBagging), a different seed or model_clear re-seeds;
	                                 1: classifier k is grown after set.seed(seed + k) */
	int      first_index;       /* global index of the first classifier built by this call */
	int      index_stride;      /* classifier indices first_index, +stride, ... (multi-GPU shard) */
	int      use_legacy_hooks;  /* 1: score through the 10-hook plugin struct with full host
	                               buffers per candidate (the reference-facing path) */
	int      verbose;
	int      n_concurrent;      /* classifiers grown concurrently on this GPU (host EM of one
	                               overlaps pair scoring of another); needs per_classifier_seed
	                               = 1 (otherwise ignored, with a warning on stderr) and is ignored
	                               with use_legacy_hooks; <= 1: one */
	int      em_on_device;      /* 1: haplotype-pair matching and the candidates' EM run on the
	                               GPU (bit-identical results); 0: on the
	                               host thread pool. Ignored with use_legacy_hooks */
	int      no_screening;      /* 0 (default): the out-of-bag / in-bag passes score only the allele-pair
	                               cells that can matter (exact screening, DESIGN.md 4.5: skipped cells
	                               are proven irrelevant, results stay bit-identical); 1: every cell.
	                               Ignored with use_legacy_hooks (always every cell) */
} hibag_b200_train_opts;

/* reference CAttrBag_Model::BuildClassifiers, src/LibHLA.cpp:2268-2305 */
int hibag_b200_model_train(hibag_b200_model *m, const hibag_b200_train_opts *opts);

/* counters of the last hibag_b200_model_train call */
typedef struct hibag_b200_train_stats {
	double   seconds_total;
	double   seconds_em;          /* summed over worker threads */
	double   seconds_gpu_wait;    /* host time blocked on the GPU */
	double   gpu_kernel_ms;       /* CUDA-event time of the scoring kernels */
	uint64_t pair_evals;          /* (sample, haplotype pair) evaluations, SURVEY.md 8d */
	uint64_t popc32_issued;       /* POPC.32 the kernels issued */
	uint64_t n_oob_evals, n_ib_evals, n_em;
	uint64_t kernel_launches;
	uint64_t h2d_bytes, d2h_bytes;
	double   cell_kernel_ms;      /* summed CUDA-event durations of the pair-scoring kernel */
	uint64_t cell_kernel_launches;
	double   seconds_prepare;     /* wall: haplotype-pair preparation per round */
	double   seconds_phase_oob;   /* wall: candidate EM + out-of-bag scoring */
	double   seconds_phase_ib;    /* wall: in-bag scoring of the candidates that need it */
	double   em_kernel_ms;        /* summed CUDA-event durations of the device EM launches */
	uint64_t n_em_host_fallback;  /* candidates re-estimated on the host (undecidable stop test) */
	uint64_t pair_evals_nominal;  /* pair evaluations the reference performs for the same passes
	                                 (pair_evals = those the GPU executed after screening) */
	uint64_t n_screen_fallback;   /* in-bag (sample, candidate) sums the screen could not certify,
	                                 rescored against every cell */
	double   gather_kernel_ms;    /* summed CUDA-event durations of cell_gather_kernel alone (the
	                                 screened passes; cell_kernel_ms spans a whole pass, bounds to
	                                 reduction) */
	uint64_t gather_kernel_launches;
	double   gather_ib_kernel_ms; /* the in-bag launches of cell_gather_kernel alone (the out-of-bag
	                                 launches score ~2.5 cells per sample and are latency-bound) */
	uint64_t gather_ib_launches;
	uint64_t gather_ib_popc32;    /* POPC.32 issued by the in-bag launches */
	uint64_t em_iterations;       /* device EM: iterations summed over the candidates */
	uint64_t em_chain_adds;       /* device EM: sum over candidates of iterations x (longest chain of
	                                 dependent fp64 adds of its M step) -- the latency floor of the
	                                 kernel in adds (x 16.9 cycles measured for a dependent DADD) */
	uint64_t em_pair_updates;     /* device EM: sum over candidates of iterations x compatible pairs */
} hibag_b200_train_stats;
int hibag_b200_model_train_stats(const hibag_b200_model *m, hibag_b200_train_stats *out);

/* search trace of the classifiers built so far: rows of 4 int64 = {global classifier index,
 * SNPs accepted so far, pair evaluations since the classifier started, candidate EM runs since
 * the classifier started}, one row per accepted SNP plus a final row per classifier (accepted
 * = -1). Returns the number of rows (copies at most max_rows). */
int hibag_b200_model_train_trace(const hibag_b200_model *m, int64_t *out, int max_rows);

int hibag_b200_model_num_classifiers(const hibag_b200_model *m);
int hibag_b200_model_clear(hibag_b200_model *m);
int hibag_b200_model_classifier_info(const hibag_b200_model *m, int k, int *n_snp,
	int *n_haplo, double *oob_acc);
/* number of bootstrap counts classifier k carries (the n_samp it was trained on or loaded with;
 * 0 when none), i.e. the ints hibag_b200_model_classifier_get writes into samp_num; < 0: error */
int hibag_b200_model_classifier_samp_num_len(const hibag_b200_model *m, int k);
/* snpidx[n_snp] 0-based, samp_num[hibag_b200_model_classifier_samp_num_len()] (may be NULL),
 * freq[n_haplo], hla[n_haplo], packed[n_haplo][2] with bits >= n_snp cleared */
int hibag_b200_model_classifier_get(const hibag_b200_model *m, int k, int32_t *snpidx,
	int32_t *samp_num, double *freq, int32_t *hla, uint64_t *packed);
/* reference HIBAG_NewClassifierHaplo / CAttrBag_Classifier::Assign, src/LibHLA.cpp:2142 */
int hibag_b200_model_add_classifier(hibag_b200_model *m, int n_snp, const int32_t *snpidx,
	const int32_t *samp_num, int n_samp, int n_haplo, const double *freq,
	const int32_t *hla, const uint64_t *packed, double oob_acc);

/* prediction outputs (any pointer may be NULL), reference CAttrBag_Model::PredictHLA,
 * src/LibHLA.cpp:2317-2412, vote_method = 1 (averaged posteriors; the only mode the
 * reference's GPU branch supports, :2433-2441) */
typedef struct hibag_b200_predict_out {
	int32_t *h1, *h2;        /* [n_samp] best guess, HIBAG_B200_NA when all weights are 0 */
	double  *max_prob;       /* [n_samp] */
	double  *matching;       /* [n_samp] */
	double  *dosage;         /* [n_samp][n_hla] */
	double  *post_prob;      /* [n_samp][n_hla*(n_hla+1)/2] */
} hibag_b200_predict_out;

/* host buffers: geno int8 [n_samp][n_snp(model)] sample-major; H2D/D2H inside */
int hibag_b200_model_predict(hibag_b200_model *m, const int8_t *geno, int n_samp,
	const hibag_b200_predict_out *out);
/* device-resident variant: geno_dev and all non-NULL outputs are DEVICE pointers on the
 * selected device (e.g. torch tensors' data_ptr()); runs on `cuda_stream` (a cudaStream_t, 0
 * for the default stream) and returns after enqueueing unless `sync` != 0 */
int hibag_b200_model_predict_device(hibag_b200_model *m, const int8_t *geno_dev, int n_samp,
	const hibag_b200_predict_out *out_dev, void *cuda_stream, int sync);

typedef struct hibag_b200_predict_stats {
	double   gpu_kernel_ms;
	double   cell_kernel_ms;      /* the dominant (pair scoring) kernel only */
	uint64_t pair_evals;
	uint64_t popc32_issued;
	uint64_t kernel_launches, cell_kernel_launches;
	uint64_t h2d_bytes, d2h_bytes;
	/* pair_evals / popc32_issued count EXECUTED work: each distinct packed genotype of a tile is
	 * scored once per classifier. Nominal = what the reference's loop nest evaluates
	 * (samples x classifiers x pairs, src/LibHLA.cpp:2451-2464); positions = (sample, classifier). */
	uint64_t pair_evals_nominal;
	uint64_t positions_scored, positions_total;
} hibag_b200_predict_stats;
int hibag_b200_model_predict_stats(const hibag_b200_model *m, hibag_b200_predict_stats *out);

/* classifier-sharded prediction: accumulate this rank's classifiers only.
 * acc_dev: DEVICE double [n_samp][n_cells + 3] = weighted posterior sums, then sum_w,
 * sum_w*match, n_used per sample; the caller all-reduces (NCCL) it over ranks and then calls
 * hibag_b200_predict_finalize_device on the reduced buffer. */
int hibag_b200_model_predict_partial_device(hibag_b200_model *m, const int8_t *geno_dev,
	int n_samp, const int32_t *snp_weight_dev, double *acc_dev, void *cuda_stream, int sync);
int hibag_b200_predict_finalize_device(int n_hla, int n_samp, const double *acc_dev,
	const hibag_b200_predict_out *out_dev, void *cuda_stream, int sync);
/* number of classifiers using each SNP (reference _GetSNPWeights, src/LibHLA.cpp:2484) */
int hibag_b200_model_snp_weights(const hibag_b200_model *m, int32_t *out_weight);

/* ---- PLINK BED import ------------------------------------------------------ */
/* reference HIBAG_ConvBED, src/HIBAG.cpp:1094-1191 (+ HIBAG_BEDFlag :1062-1083): decode the
 * 2-bit packed genotypes of a .bed file on the GPU. bed_file = the whole file (host, n_bytes,
 * 3-byte prefix included; both the SNP-major and the individual-major mode); snp_flag[n_snp] != 0
 * keeps a SNP (NULL: all); out = int8 [n_samp][n_save] sample-major, 0/1/2 = count of the .bim's
 * first allele as the reference reports it, -1 = missing (the reference's NA); *n_save = SNPs kept.
 * out may be NULL to query *n_save. Errors as the reference's: "Invalid prefix in the PLINK BED
 * file." */
int hibag_b200_bed_decode(const uint8_t *bed_file, size_t n_bytes, int n_samp, int n_snp,
	const int32_t *snp_flag, int8_t *out, int *n_save, double *kernel_ms);
/* device-resident variant: payload_dev = the bytes after the prefix, sel_dev = ascending indices of
 * the SNPs to keep (NULL: all), out_dev = int8 [n_samp][n_save]; enqueued on cuda_stream */
int hibag_b200_bed_decode_device(const uint8_t *payload_dev, int mode, int n_samp, int n_snp,
	const int32_t *sel_dev, int n_save, int8_t *out_dev, void *cuda_stream);

/* The library recycles device and pinned blocks through a process-wide cache (at most 48 GB of HBM,
 * 24 GB pinned). This gives every cached block back to the driver, e.g. before another library
 * needs the memory; returns the bytes released. */
size_t hibag_b200_trim_cache(void);

/* Page-locked host buffers from that cache, for the host-array entry points: results copied into
 * them leave the GPU at PCIe rate while the next tile is scored (a pageable destination is staged
 * by the driver and page-faults on first touch: 1 GB of posterior rows per 200,000 samples).
 * The reference has no counterpart (its hlaPredict fills an R matrix column by column,
 * src/HIBAG.cpp:680-741). NULL on failure (hibag_b200_last_error). */
void *hibag_b200_host_alloc(size_t bytes);
void hibag_b200_host_free(void *p);

/* SM-time accounting of the current device: out[16] = per kernel class the sum over its CTAs of
 * (SM cycles the CTA was resident) x 1024 / (CTAs of that launch that fit one SM), i.e. 1/1024
 * SM-cycles held. Classes: 0 out-of-bag gather, 1 in-bag gather, 2 EM, 3 screen bounds, 4 need lists,
 * 5 task prefix, 6 / 7 screened out-of-bag / in-bag reduction, 8 plain pair-scoring kernel, 10 the EM
 * kernel's plain CTA-resident cycles (x 1024).
 * Concurrent kernels overlap, so CUDA-event durations cannot attribute the GPU's time; held SM-time
 * can. reset != 0 zeroes the counters after reading. */
int hibag_b200_sm_time(uint64_t *out, int reset);

/* ---- host-only pieces exposed for the CPU test-suite (no GPU needed) ------------------------- */
/* R's Mersenne-Twister after set.seed(seed): n draws of unif_rand() */
int hibag_b200_host_unif_rand(uint32_t seed, int n, double *out);
/* the work list the scoring kernel receives for a haplotype list: out_cells int32[n_cells][8]
 * = {a_start,a_n,b_start,b_n,out_idx,diag,allele a,allele b} in launch order, out_chunks int32[<=n_cells][2]
 * = {cell_begin,cell_end}; *pairs = haplotype pairs scored per sample */
int hibag_b200_host_build_tasks(const hibag_haplotype *haplo, int n_haplo, int n_hla, int n_snp,
	int target_chunks, int32_t *out_cells, int32_t *out_chunks, int *n_chunks, uint64_t *pairs);
/* constants of the exact screening of the training passes: table[257] =
 * EXP_LOG_MIN_RARE_FREQ as the host computed it (src/LibHLA.cpp:166-183), floor_table[257] =
 * max(table, 1e-100), *bound_factor = K of bound(a,b) = U_a * U_b * K */
int hibag_b200_host_screen_constants(double *table, double *floor_table, double *bound_factor);

/* ---- microbenchmarks of the pipes that bound this path ----------------- */
/* which: 0 POPC.32, 1 LOP3, 2 DMUL+DADD, 3 DFMA, 4 LDS.64 (lane-private), 5 IADD3: lane-ops per
 * second over the whole device in *out_ops_per_s. 6 DADD, 7 DMUL+DADD: ONE warp running one
 * dependent chain per lane (32 * clock / *out_ops_per_s = dependent-issue latency in cycles). */
int hibag_b200_pipe_peak(int which, double *out_ops_per_s, double *out_ms);

#ifdef __cplusplus
}
#endif
#endif /* HIBAG_B200_H */
