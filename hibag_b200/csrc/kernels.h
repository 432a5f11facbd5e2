// kernels.h -- launch interface of the sm_100a scoring kernels (internal to libhibag_b200.so)
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace hb {

/// One allele-pair cell (a,b), a <= b, of the posterior matrix: haplotype ranges of the two
/// alleles inside the list and the position of the cell in the upper-triangular vector
/// (reference src/LibHLA.cpp:1520-1530).
struct CellTask
{
	int a_start, a_n;    // haplotypes of allele a: [a_start, a_start + a_n)
	int b_start, b_n;    // haplotypes of allele b
	int out_idx;         // H2 + H1*(2*n_hla-H1-1)/2
	int diag;            // 1 when a == b (upper triangle of pairs, first term f*f)
	int al_a, al_b;      // the two allele indices (32-byte records: two aligned vector loads)
};

/// A run of consecutive cells [cell_begin, cell_end) processed by one warp for one group of
/// samples; chunks are sorted by decreasing cost so the long fp64 chains start first.
struct Chunk
{
	int cell_begin, cell_end;
};

/// Device-side description of one scoring pass: all cells of one haplotype list against a
/// list of samples.
struct CellPass
{
	// haplotype records in kernel layout (see pack_hap_records): 16 B (<=64 SNPs) or 32 B each
	const void *hap;
	int n_hap;
	int n_snp;
	// rare-frequency table EXP_LOG_MIN_RARE_FREQ[257] (host-computed, src/LibHLA.cpp:166-183)
	const double *table;
	int n_dist;              // table entries the kernel stages (distances are clamped to n_dist-1)
	// packed genotypes, SoA 32-bit words: s1[w * geno_stride + sample], same for s2
	const uint32_t *s1, *s2;
	int geno_stride;
	// optional candidate SNP patched in at bit `cand_bit`: raw genotype column int8 [sample]
	const int8_t *cand_col;
	int cand_bit;
	// samples to score: sample index = samp_list ? samp_list[pos] : pos, pos in [0, n_pos)
	const int *samp_list;
	int n_pos;
	// optional: the number of positions is read from device memory (<= n_pos, which then only
	// sizes the launch) -- the distinct genotypes of a prediction tile are counted on the device
	const int *n_pos_dev;
	// work list
	const CellTask *cells;
	const Chunk *chunks;
	int n_chunks;
	unsigned int *task_counter;   // zeroed before launch
	// output: P[out_idx * p_stride + pos] = raw cell sum
	double *P;
	size_t p_stride;
};

/// One haplotype list of a batched launch (all lists of a batch share the samples, the SNP
/// count and the table; each has its own cells/chunks, optional candidate SNP column and
/// output matrix). Passed by value inside the kernel parameters.
struct ListDesc
{
	const void *hap;
	const CellTask *cells;
	const Chunk *chunks;
	const int8_t *cand_col;       // raw genotype column patched in at bit cand_bit, or null
	double *P;                    // P[out_idx * p_stride + pos]
	int n_hap, n_chunks, cand_bit, pad;
};

static const int MAX_BATCH_LISTS = 32;

/// SM-time accounting classes (devutil.cuh SmAcct; hibag_b200_sm_time in the C ABI)
enum { SM_ACCT_GATHER_OOB = 0, SM_ACCT_GATHER_IB = 1, SM_ACCT_EM = 2, SM_ACCT_BOUND = 3, SM_ACCT_NEED = 4,
       SM_ACCT_TASKS = 5, SM_ACCT_REDUCE_OOB = 6, SM_ACCT_REDUCE_IB = 7, SM_ACCT_CELL_PASS = 8,
       SM_ACCT_EM_PREP = 9, SM_ACCT_EM_CTA = 10, SM_ACCT_DEDUP = 11, SM_ACCT_N = 16 };
/// per-device counters [SM_ACCT_N] (device memory, zeroed at first use): sum over CTAs of resident
/// cycles x (1024 / CTAs of that launch that fit an SM)
unsigned long long *device_sm_acct();

/// Device-side description of one launch of the pair-scoring kernel: n_lists haplotype lists
/// (e.g. the candidate SNPs of one selection round) against one list of samples.
struct CellBatch
{
	const double *table;
	const uint32_t *s1, *s2;
	const int *samp_list;
	unsigned int *task_counters;  // [n_lists], zeroed before launch
	size_t p_stride;
	int n_dist, n_snp, geno_stride, n_pos;
	int n_lists, max_hap;         // max_hap = largest n_hap of the batch (shared-memory sizing)
	unsigned long long *acct;     // SM-time counters or null
	int acct_w;
	const int *n_pos_dev;         // if set: positions = min(*n_pos_dev, n_pos)
	int flat;                     // bit 0: entry-flat tasks (cell_gather_flat_kernel; task_prefix built with flat = true), bit 1: one list per CTA
	ListDesc lists[MAX_BATCH_LISTS];
};

/// bytes of one haplotype record in kernel layout for a classifier of n_snp SNPs
inline int hap_record_bytes(int n_snp) { return (n_snp <= 64) ? 16 : 32; }
/// number of 32-bit genotype words the kernels use for n_snp SNPs (1, 2 or 4)
inline int geno_words(int n_snp) { return (n_snp <= 32) ? 1 : ((n_snp <= 64) ? 2 : 4); }

/// Launch the pair-scoring kernel. samples_per_lane in {1,2,4}. Returns the number of POPC.32
/// issued per pair evaluation by the chosen instantiation (1, 2 or 4) for accounting.
int launch_cell_pass(const CellPass &p, int samples_per_lane, int sm_count, cudaStream_t st);
/// the same for a batch of lists (n_lists <= MAX_BATCH_LISTS)
int launch_cell_batch(const CellBatch &b, int samples_per_lane, int sm_count, cudaStream_t st);

/// AoS TGenotype[n] (48 B) -> SoA words + true alleles + bootstrap counts
/// rewrite SNP column snp_bit (0..127) of the SoA planes from codes (bit0 -> s1, bit1 -> s2)
void launch_patch_column(const int8_t *code, int n, uint32_t *s1, uint32_t *s2, int stride,
	int snp_bit, cudaStream_t st);
void launch_unpack_genotypes(const void *geno_aos, int n, uint32_t *s1, uint32_t *s2,
	int stride, int *a1, int *a2, int *boot, cudaStream_t st);

/// out-of-bag accuracy: per position argmax over cells (strict '<', first wins), compare with
/// the true type (src/LibHLA.cpp:912-924), integer sum into *out_count (zeroed by the caller)
/// (n_lists > 1: list l reads P + l*list_stride and adds into out_count[l])
void launch_reduce_oob(const double *P, size_t p_stride, int n_hla, const int *samp_list,
	int n_pos, const int *a1, const int *a2, int *out_count, cudaStream_t st, int n_lists = 1,
	size_t list_stride = 0);

/// in-bag: per position P_true / sum_cells (sequential sum in cell order)
/// (n_lists > 1: list l reads P + l*list_stride and writes out_ratio + l*out_stride)
void launch_reduce_ib(const double *P, size_t p_stride, int n_hla, const int *samp_list,
	int n_pos, const int *a1, const int *a2, double *out_ratio, cudaStream_t st, int n_lists = 1,
	size_t list_stride = 0, size_t out_stride = 0);

/// best guess per position written as allele pair (for hibag_b200_best_guess)
void launch_reduce_best_guess(const double *P, size_t p_stride, int n_hla, int n_pos,
	int *out_a1, int *out_a2, cudaStream_t st);

/// PostProb2 normalisation in place + raw sum per position
void launch_normalize(double *P, size_t p_stride, int n_hla, int n_pos, double *out_sum,
	cudaStream_t st);

/// transpose P[cell][pos] -> out[pos][cell]
void launch_transpose(const double *P, size_t p_stride, int n_cells, int n_pos, double *out,
	cudaStream_t st);


// ---- exact screening of the training passes (screen.cu) --------------------------------------
//
// In training every sample's true HLA type is known. One term of the true cell's chain -- the
// best haplotype of one true allele with its best partner in the other -- evaluated exactly as
// the chain evaluates it, is a lower bound x_ref of the true cell's value (a sum of non-negative
// terms is at least each term), hence of the best cell's. A cell (a,b) whose value cannot exceed
//     bound(a,b) = U_a * U_b * K,   U_a = sum_{i in a} f_i * T'[hom-SNP mismatches of h_i]
// (the heterozygous SNPs can only add distance; T' = max(T, 1e-100), K covers 2x, the table's
// rounding and the chains' rounding) is
//   * out-of-bag (argmax): irrelevant when bound < x_ref  -- it cannot be the best guess;
//   * in-bag (sequential sum): irrelevant when its bound is far below x_ref, which the
//     reduction CERTIFIES: it adds the chain once with 0 and once with the bound for every
//     skipped cell; fp64 addition is monotone, so equal results prove the full chain has that
//     value. Samples that fail the certificate are rescored without screening.
// Only the cells that survive (and always the true cell) are scored, by a gather kernel whose
// lanes are the samples that need the cell. Every value that is produced is the reference's own
// chain, bit for bit.

/// One haplotype list of a screened launch
struct GatherList
{
	const void *hap;
	const CellTask *cells;        // in decreasing-cost order (as in the blob)
	const int8_t *cand_col;
	double *P;                    // P[out_idx * p_stride + pos]
	const int *count;             // [n_cells] by out_idx: positions that need the cell
	const int *entries;           // positions, cell c at ent_off[c] .. + count[c]
	const unsigned int *task_prefix;   // [n_cells + 2]: prefix over the blob's cell order, n_tasks, log2(positions per task)
	int n_hap, cand_bit;
};

struct GatherBatch
{
	const double *table;
	const uint32_t *s1, *s2;
	const int *samp_list;
	const int *ent_off;           // [n_cells] by out_idx (shared by the lists of a launch)
	unsigned int *task_counters;  // [n_lists], zeroed before launch
	size_t p_stride;
	int n_dist, n_snp, geno_stride, n_pos;
	int n_lists, max_hap, n_cells, acct_cls;
	unsigned long long *acct;     // SM-time counters or null
	int acct_w;
	int flat;                     // bit 0: entry-flat tasks (cell_gather_flat_kernel; task_prefix built with flat = true), bit 1: one list per CTA
	GatherList lists[MAX_BATCH_LISTS];
};

/// the surviving cells of all lists in one launch. Returns POPC.32 per pair evaluation.
/// max_ctas > 0 caps the persistent grid (small passes leave room for other lanes' launches).
int launch_cell_gather(const GatherBatch &b, int sm_count, cudaStream_t st, long long max_ctas = 0);

/// arguments shared by the screening kernels of one launch (list l uses slice l of every array)
struct ScreenArgs
{
	const double *table_floor;    // T' on the device
	const double *table;          // T on the device (rescue path of the in-bag reduction)
	const uint32_t *s1, *s2;
	const int *samp_list;
	const int *a1, *a2;           // true types by sample, a1 <= a2
	int n_dist, n_snp, geno_stride, n_pos, n_hla, n_lists;
	size_t p_stride;
	double K;                     // bound factor
	double tau;                   // need a cell when bound >= tau * x_ref
	double *U;                    // [n_lists][n_hla][p_stride]
	double *xref;                 // [n_lists][p_stride] lower bound of the true cell's value
	double *P;                    // [n_lists][n_cells][p_stride]
	int *count;                   // [n_lists][n_cells]
	int *entries;                 // [n_lists][n_cells][p_stride]
	unsigned int *task_prefix;    // [n_lists][n_cells + 2]
	unsigned long long *evals;    // [n_lists] pair evaluations the gather launch will execute
	unsigned long long *rescued;  // [1] in-bag positions the reduction rescued
	int *al_tab;                  // [n_lists][n_hla][2] first haplotype and count per allele
	int force_rescue;             // test hook: > 0 rescues every force_rescue-th position
	int device_rescue;            // 1: uncertified sums are rescued inside the reduction; 0: reported
	                              // as ratio -1 and rescored by the caller with the plain kernel
	unsigned long long *acct;     // SM-time counters or null
	// second level of the in-bag screen (null / 0: off): per-allele sums split by the haplotypes' alleles
	// at the sample's first two heterozygous SNPs, and the compacted need lists the gather launch uses
	double *U2;                   // [n_lists][n_hla][p_stride][4]
	int *hetk;                    // [n_lists][p_stride] heterozygous SNPs used (0..2) | true cell << 2
	int *count2;                  // [n_lists][n_cells]
	int *entries2;                // [n_lists][n_cells][p_stride]
	double K2;                    // bound factor of the second level
	double tf[3];                 // T'[0], T'[1], T'[2]
	int u_smem;                   // 1: the need / reduction kernels keep a position's per-allele sums in shared memory
	// position classes (null: off): positions of a list whose sample has the same packed genotype (candidate
	// SNP patched in) and the same true type get the same bounds, need list, cell values and reduction --
	// only one REPRESENTATIVE of each class is screened and scored, the others copy its result
	const int *rep;               // [n_lists][p_stride] representative position of every position
	const int *rep_list;          // [n_lists][p_stride] the representatives, compacted
	const int *n_rep;             // [n_lists] how many
	int *pos_res;                 // [n_lists][p_stride] out-of-bag: correct alleles per representative
};
struct ScreenList { const void *hap; const CellTask *cells; const int8_t *cand_col; int n_hap, cand_bit; };
struct ScreenLists { ScreenList l[MAX_BATCH_LISTS]; };

/// position classes of every list of the launch: table = int [n_lists][table_size] (power of two >=
/// 2 * n_pos) followed by int [n_lists] counters, ALL ZEROED by the caller; fills rep, rep_list and the
/// counters (= a.n_rep)
void launch_screen_dedup(const ScreenArgs &a, const ScreenLists &ls, int *table, int table_size, int *rep,
	int *rep_list, int *n_rep, cudaStream_t st);
/// every position takes its representative's result: out-of-bag -- sum of pos_res into out_count[l];
/// in-bag -- ratio copied
void launch_screen_broadcast_oob(const ScreenArgs &a, int *out_count, cudaStream_t st);
void launch_screen_broadcast_ib(const ScreenArgs &a, double *out_ratio, size_t out_stride, cudaStream_t st);
/// U[l][a][pos] and xref[l][pos]
void launch_screen_bound(const ScreenArgs &a, const ScreenLists &ls, cudaStream_t st);
/// per list: task prefix over the blob's cell order from count (stride 0: one shared count array)
/// and the pair evaluations those tasks hold (added to evals[l])
void launch_screen_tasks(const ScreenLists &ls, int n_lists, int n_cells, const int *count,
	size_t count_stride, unsigned int *task_prefix, unsigned long long *evals, int target_tasks,
	int warp_slots, cudaStream_t st, bool flat = false);
/// per (list, pos): which cells are needed (appends to entries / count)
void launch_screen_need(const ScreenArgs &a, cudaStream_t st);
/// second level: every (cell, position) entry of the need lists is tested against the class bound;
/// survivors go to entries2 / count2, the others get minus their bound into P (the reduction adds it)
void launch_screen_refine(const ScreenArgs &a, cudaStream_t st);
/// screened reductions (same outputs as launch_reduce_oob / launch_reduce_ib). An in-bag position
/// whose sum is not certified is rescued inside the reduction: the lanes of its warp score its
/// skipped cells one by one, then the plain sequential sum is taken. (ratio -1 = not certified
/// is still understood by the caller, which rescores such positions with the plain kernel.)
void launch_reduce_oob_screened(const ScreenArgs &a, int *out_count, cudaStream_t st);
void launch_reduce_ib_screened(const ScreenArgs &a, const ScreenLists &ls, double *out_ratio,
	size_t out_stride, cudaStream_t st);

// ---- prediction --------------------------------------------------------------------------

/// int8 matrix transpose: in[rows][cols] -> out[cols][rows]
void launch_transpose_i8(const int8_t *in, int rows, int cols, int8_t *out, cudaStream_t st);

/// pack raw int8 genotypes for one classifier's SNP order into SoA words for a tile of samples
/// and compute the classifier's weight per sample (src/LibHLA.cpp:667-706 and 2418-2431).
/// geno_t is SNP-major: int8 [n_snp_total][n_samp_total]
void launch_pack_classifier(const int8_t *geno_t, size_t n_samp_total, int samp_begin,
	int n_tile, const int *snpidx, int n_snp, const int *snp_weight, uint32_t *s1,
	uint32_t *s2, int stride, double *weight, cudaStream_t st);

/// per sample of the tile: s = sum_cells P (sequential); acc[cell] += (P[cell]/s... ) see .cu
/// acc layout: acc[cell * acc_stride + pos]; aux[0..2][pos] = sum_w, sum_w*match, n_used
void launch_predict_accumulate(const double *P, size_t p_stride, int n_cells, int n_tile,
	const double *weight, double *acc, size_t acc_stride, double *aux, cudaStream_t st);

/// Exact de-duplication of a tile's packed genotypes (prediction). A sample's cell sums depend only
/// on its packed genotype at the classifier's SNPs, and cohorts repeat them (the haplotypes of a
/// population are few: 22 % of the 200,000 x 100 (sample, classifier) genotypes of configs[2] are
/// distinct), so each distinct genotype is scored ONCE and its column of the cell matrix is shared.
/// table: int [table_size] (power of two >= 2 * n_tile) open-addressing set keyed by the full
/// genotype words; uid[pos] = dense number of the sample's genotype, rep_list[u] = one sample that
/// carries genotype u, *n_unique = how many there are. Which of the equal samples becomes the
/// representative is a race and does not matter: equal inputs, equal bits.
void launch_dedup_genotypes(const uint32_t *s1, const uint32_t *s2, int stride, int nw, int n_tile,
	int *table, int table_size, int *repof, int *uid, int *rep_list, int *n_unique, cudaStream_t st);

/// predict_accumulate over a de-duplicated cell matrix: P[cell * p_stride + u], u = uid[pos];
/// norm[2 * u_stride]: scratch for the per-genotype 1/sum and sum. Same operations per sample, in
/// the same order, as launch_predict_accumulate.
void launch_predict_accumulate_dedup(const double *P, size_t p_stride, int n_cells, int n_tile,
	const int *uid, const int *n_unique, double *norm, size_t u_stride, const double *weight,
	double *acc, size_t acc_stride, double *aux, cudaStream_t st);

/// finalisation of a tile: normalise by sum_w, matching, best guess, max prob, dosage,
/// posterior rows. Outputs are indexed by global sample (samp_begin + pos); any may be NULL.
void launch_predict_finalize(double *acc, size_t acc_stride, const double *aux, int n_hla,
	int samp_begin, int n_tile, int *h1, int *h2, double *max_prob, double *matching,
	double *dosage, double *post_prob, cudaStream_t st);

/// classifier-sharded path: export / import the accumulators as [sample][n_cells + 3]
void launch_export_partial(const double *acc, size_t acc_stride, const double *aux,
	int n_cells, int samp_begin, int n_tile, double *out, cudaStream_t st);
void launch_finalize_from_partial(const double *partial, int n_hla, int n_samp, int *h1,
	int *h2, double *max_prob, double *matching, double *dosage, double *post_prob,
	cudaStream_t st);

// ---- PLINK BED decoding (bed.cu) ----------------------------------------------------------------
/// payload = the file's bytes after the 3-byte prefix (device); mode 0 individual-major, else
/// SNP-major; sel = indices of the SNPs to keep (device, ascending; null = all n_snp, n_save =
/// n_snp); out = int8 [n_samp][n_save] (device), 0/1/2 and -1 for missing
void launch_bed_decode(const uint8_t *payload, int mode, int n_samp, int n_snp, const int32_t *sel,
	int n_save, int8_t *out, cudaStream_t st);

// ---- microbenchmarks ---------------------------------------------------------------------
double run_pipe_peak(int which, int sm_count, double *out_ms);

}  // namespace hb
