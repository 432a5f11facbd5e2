/* oracle/hibag_oracle.h -- CPU restatement of HIBAG's haplotype-pair scoring path.
 *
 * TEST INFRASTRUCTURE. This is the parity checker for the CUDA path; it is never linked or
 * loaded by anything under hibag_b200/. Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline leg may use it.
 *
 * Parity status: PINNED. tests/test_oracle.py checks every function below against the compiled,
 * unmodified reference (oracle/_ref/libhibag_ref.so, target "base") on the reference's own
 * HapMap fixture and on seeded synthetic inputs, and the reference build itself is pinned
 * bit-for-bit to the reference's golden model inst/extdata/ModelList.RData
 * (tests/golden/modellist_a.npz).
 *
 * All `file:line` citations are relative to /root/reference/.
 */
#ifndef HIBAG_ORACLE_H
#define HIBAG_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* binary-compatible with THaplotype / TGenotype, inst/include/LibHLA_ext.h:261-299, 311-352 */
typedef struct {
	uint64_t packed[2];   /* bit i = allele of SNP i */
	double   freq;
	float    freq_f32;    /* aux.a2.Freq_f32 */
	int32_t  hla;         /* aux.a2.HLA_allele, 0-based, non-decreasing along the list */
} oracle_haplo_t;

typedef struct {
	uint64_t s1[2], s2[2]; /* (s1,s2) per SNP: 0->(0,0) 1->(1,0) 2->(1,1) missing->(0,1) */
	int32_t  boot;         /* BootstrapCount */
	int32_t  a1, a2;       /* aux_hla_type, a1 <= a2 */
	int32_t  tmp;
} oracle_geno_t;

#define ORACLE_NA_INTEGER  INT32_MIN
#define ORACLE_TABLE_LEN   257

/* src/LibHLA.cpp:166-183 */
void hibag_oracle_table(double T[ORACLE_TABLE_LEN]);
/* src/LibHLA.cpp:802-817 */
int hibag_oracle_hamming(const oracle_geno_t *g, const oracle_haplo_t *h1,
	const oracle_haplo_t *h2, int n_snp);
/* raw (un-normalised) allele-pair sums in cell order, the common body of
 * src/LibHLA.cpp:1639-1830; cells has n_hla*(n_hla+1)/2 entries */
void hibag_oracle_cells(const oracle_haplo_t *haplo, int n_haplo, int n_hla, int n_snp,
	const oracle_geno_t *g, double *cells);
/* src/LibHLA.cpp:1639-1704 */
void hibag_oracle_best_guess(const oracle_haplo_t *haplo, int n_haplo, int n_hla, int n_snp,
	const oracle_geno_t *geno, int n_geno, int32_t *out_a1, int32_t *out_a2);
/* src/LibHLA.cpp:1706-1767 (HLA type = the genotype's own a1/a2) */
void hibag_oracle_post_prob(const oracle_haplo_t *haplo, int n_haplo, int n_hla, int n_snp,
	const oracle_geno_t *geno, int n_geno, double *out);
/* src/LibHLA.cpp:1769-1830 */
void hibag_oracle_post_prob2(const oracle_haplo_t *haplo, int n_haplo, int n_hla, int n_snp,
	const oracle_geno_t *geno, int n_geno, double *out_prob, double *out_sum);
/* src/LibHLA.cpp:1934-1955 with Compare :912-924 */
int hibag_oracle_acc_oob(const oracle_haplo_t *haplo, int n_haplo, int n_hla, int n_snp,
	const oracle_geno_t *geno, int n_geno);
/* src/LibHLA.cpp:1957-1979 */
double hibag_oracle_acc_ib(const oracle_haplo_t *haplo, int n_haplo, int n_hla, int n_snp,
	const oracle_geno_t *geno, int n_geno);
/* src/LibHLA.cpp:2443-2481, vote_method = 1 */
void hibag_oracle_predict_avg(int n_hla, int n_classifier,
	const oracle_haplo_t *const *haplo, const int *n_haplo, const int *n_snp,
	const oracle_geno_t *geno, const double *weight, double *out_prob, double *out_match);
/* src/LibHLA.cpp:1549-1566 */
void hibag_oracle_best_guess_cells(const double *prob, int n_hla, int32_t *a1, int32_t *a2);
/* src/LibHLA.cpp:2387-2402 */
void hibag_oracle_dosage(const double *prob, int n_hla, double *dosage);
/* src/LibHLA.cpp:667-706 */
void hibag_oracle_int_to_snp(oracle_geno_t *out, int length, const int32_t *geno_base,
	const int32_t *index);
/* src/LibHLA.cpp:2418-2431 and 2484-2496 */
void hibag_oracle_classifier_weights(int n_classifier, const int *n_snp,
	const int32_t *const *snpidx, int n_total_snp, const int32_t *geno_row, double *weight);

/* records of a build_haplomatch plugin, src/LibHLA.cpp:1014-1072 with the selection rule of
 * _PrepHaploMatch_def :1569-1637 (see the .c file) */
long hibag_oracle_haplomatch_records(const oracle_haplo_t *haplo, const int64_t *n_haplo,
	int n_hla, int n_snp, const oracle_geno_t *geno, int n_samp, uint32_t *out, long max_records);

/* PLINK BED decoding, src/HIBAG.cpp:1094-1191 (see the .c file) */
int hibag_oracle_bed_decode(const uint8_t *file, long n_bytes, int n_samp, int n_snp,
	const int32_t *snp_flag, int n_save, int32_t *out);

#ifdef __cplusplus
}
#endif
#endif
