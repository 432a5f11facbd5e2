/* oracle/hibag_oracle.c -- see hibag_oracle.h. TEST INFRASTRUCTURE, parity pinned.
 *
 * Plain scalar C, written for clarity: every allele-pair cell is one sequential fp64 chain
 * "sum += (p * T[d])" in (i outer, j inner) order -- the order of the reference's base target
 * (src/LibHLA.cpp:1639-1830, macro ADD_FREQ_MUTANT src/LibHLA.h:222-223). Compile with
 * -ffp-contract=off so no multiply-add is fused.
 */
#include "hibag_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

static double g_table[ORACLE_TABLE_LEN];
static int g_table_ready = 0;

void hibag_oracle_table(double T[ORACLE_TABLE_LEN])
{
	/* src/LibHLA.cpp:175-183: exp(d*log(1e-5)), [0] forced to 1, non-finite -> 0 */
	for (int d = 0; d < ORACLE_TABLE_LEN; d++)
	{
		double v = exp(d * log(1e-5));
		if (d == 0) v = 1;
		if (!isfinite(v)) v = 0;
		T[d] = v;
	}
}

static const double *table(void)
{
	if (!g_table_ready) { hibag_oracle_table(g_table); g_table_ready = 1; }
	return g_table;
}

static int popc64(uint64_t v) { return __builtin_popcountll(v); }

int hibag_oracle_hamming(const oracle_geno_t *g, const oracle_haplo_t *h1,
	const oracle_haplo_t *h2, int n_snp)
{
	/* src/LibHLA.cpp:806-817; one 64-bit word when n_snp <= 64, else two (:756,768) */
	const int words = (n_snp <= 64) ? 1 : 2;
	int d = 0;
	for (int w = 0; w < words; w++)
	{
		const uint64_t H1 = h1->packed[w], H2 = h2->packed[w];
		const uint64_t S1 = g->s1[w], S2 = g->s2[w];
		const uint64_t missing = S2 & ~S1;
		const uint64_t mask = ((H1 ^ S2) | (H2 ^ S1)) & ~missing;
		d += popc64((H1 ^ S1) & mask) + popc64((H2 ^ S2) & mask);
	}
	return d;
}

/* haplotype range [start[a], start[a+1]) of every allele, derived from the per-haplotype
 * allele tag (the plugin never receives LenPerHLA, src/LibHLA.cpp:1916-1920) */
static int *allele_starts(const oracle_haplo_t *haplo, int n_haplo, int n_hla)
{
	int *st = (int *)calloc((size_t)n_hla + 1, sizeof(int));
	for (int i = 0; i < n_haplo; i++) st[haplo[i].hla + 1]++;
	for (int a = 0; a < n_hla; a++) st[a + 1] += st[a];
	return st;
}

static void cells_with_starts(const oracle_haplo_t *haplo, const int *st, int n_hla, int n_snp,
	const oracle_geno_t *g, double *cells)
{
	const double *T = table();
	double *out = cells;
	for (int a = 0; a < n_hla; a++)
	{
		/* diagonal cell (a,a): src/LibHLA.cpp:1652-1668 */
		double sum = 0;
		for (int i = st[a]; i < st[a + 1]; i++)
		{
			const double fi = haplo[i].freq;
			sum += (fi * fi) * T[hibag_oracle_hamming(g, &haplo[i], &haplo[i], n_snp)];
			const double ff = 2 * fi;
			for (int j = i + 1; j < st[a + 1]; j++)
				sum += (ff * haplo[j].freq) * T[hibag_oracle_hamming(g, &haplo[i], &haplo[j], n_snp)];
		}
		*out++ = sum;
		/* off-diagonal cells (a,b), b > a: src/LibHLA.cpp:1676-1691 */
		for (int b = a + 1; b < n_hla; b++)
		{
			sum = 0;
			for (int i = st[a]; i < st[a + 1]; i++)
			{
				const double ff = 2 * haplo[i].freq;
				for (int j = st[b]; j < st[b + 1]; j++)
					sum += (ff * haplo[j].freq) * T[hibag_oracle_hamming(g, &haplo[i], &haplo[j], n_snp)];
			}
			*out++ = sum;
		}
	}
}

void hibag_oracle_cells(const oracle_haplo_t *haplo, int n_haplo, int n_hla, int n_snp,
	const oracle_geno_t *g, double *cells)
{
	int *st = allele_starts(haplo, n_haplo, n_hla);
	cells_with_starts(haplo, st, n_hla, n_snp, g, cells);
	free(st);
}

void hibag_oracle_best_guess_cells(const double *prob, int n_hla, int32_t *a1, int32_t *a2)
{
	/* strict '<', first maximum in cell order wins, all-zero -> NA (src/LibHLA.cpp:1549-1566,
	 * same rule inside _BestGuess_def :1670,1693) */
	double best = 0;
	*a1 = *a2 = ORACLE_NA_INTEGER;
	for (int h1 = 0; h1 < n_hla; h1++)
		for (int h2 = h1; h2 < n_hla; h2++, prob++)
			if (best < *prob) { best = *prob; *a1 = h1; *a2 = h2; }
}

void hibag_oracle_best_guess(const oracle_haplo_t *haplo, int n_haplo, int n_hla, int n_snp,
	const oracle_geno_t *geno, int n_geno, int32_t *out_a1, int32_t *out_a2)
{
	const size_t nc = (size_t)n_hla * (n_hla + 1) / 2;
	int *st = allele_starts(haplo, n_haplo, n_hla);
	double *cells = (double *)malloc(nc * sizeof(double));
	for (int s = 0; s < n_geno; s++)
	{
		cells_with_starts(haplo, st, n_hla, n_snp, &geno[s], cells);
		hibag_oracle_best_guess_cells(cells, n_hla, &out_a1[s], &out_a2[s]);
	}
	free(cells); free(st);
}

static size_t cell_index(int h1, int h2, int n_hla)
{
	/* src/LibHLA.cpp:1709-1712 */
	if (h1 > h2) { int t = h1; h1 = h2; h2 = t; }
	return (size_t)h2 + (size_t)h1 * (2 * n_hla - h1 - 1) / 2;
}

void hibag_oracle_post_prob(const oracle_haplo_t *haplo, int n_haplo, int n_hla, int n_snp,
	const oracle_geno_t *geno, int n_geno, double *out)
{
	const size_t nc = (size_t)n_hla * (n_hla + 1) / 2;
	int *st = allele_starts(haplo, n_haplo, n_hla);
	double *cells = (double *)malloc(nc * sizeof(double));
	for (int s = 0; s < n_geno; s++)
	{
		cells_with_starts(haplo, st, n_hla, n_snp, &geno[s], cells);
		double sum = 0;                                  /* src/LibHLA.cpp:1739-1766 */
		for (size_t c = 0; c < nc; c++) sum += cells[c];
		out[s] = cells[cell_index(geno[s].a1, geno[s].a2, n_hla)] / sum;
	}
	free(cells); free(st);
}

void hibag_oracle_post_prob2(const oracle_haplo_t *haplo, int n_haplo, int n_hla, int n_snp,
	const oracle_geno_t *geno, int n_geno, double *out_prob, double *out_sum)
{
	const size_t nc = (size_t)n_hla * (n_hla + 1) / 2;
	int *st = allele_starts(haplo, n_haplo, n_hla);
	for (int s = 0; s < n_geno; s++)
	{
		double *cells = out_prob + nc * (size_t)s;
		cells_with_starts(haplo, st, n_hla, n_snp, &geno[s], cells);
		double sum = 0;                                  /* src/LibHLA.cpp:1823-1829 */
		for (size_t c = 0; c < nc; c++) sum += cells[c];
		const double ff = 1 / sum;
		for (size_t c = 0; c < nc; c++) cells[c] *= ff;
		out_sum[s] = sum;
	}
	free(st);
}

static int compare_types(int p1, int p2, int t1, int t2)
{
	/* number of predicted alleles found in the true unordered pair, each true allele usable
	 * once (src/LibHLA.cpp:912-924) */
	int cnt = 0;
	if (p1 == t1) { cnt = 1; t1 = -1; }
	else if (p1 == t2) { cnt = 1; t2 = -1; }
	if (p2 == t1 || p2 == t2) cnt++;
	return cnt;
}

int hibag_oracle_acc_oob(const oracle_haplo_t *haplo, int n_haplo, int n_hla, int n_snp,
	const oracle_geno_t *geno, int n_geno)
{
	int total = 0;
	for (int s = 0; s < n_geno; s++)
	{
		if (geno[s].boot != 0) continue;                 /* OOB <=> count == 0 (:1870-1873) */
		int32_t a1, a2;
		hibag_oracle_best_guess(haplo, n_haplo, n_hla, n_snp, &geno[s], 1, &a1, &a2);
		total += compare_types(a1, a2, geno[s].a1, geno[s].a2);
	}
	return total;
}

double hibag_oracle_acc_ib(const oracle_haplo_t *haplo, int n_haplo, int n_hla, int n_snp,
	const oracle_geno_t *geno, int n_geno)
{
	double loglik = 0;                                   /* src/LibHLA.cpp:1965-1977 */
	for (int s = 0; s < n_geno; s++)
	{
		if (geno[s].boot <= 0) continue;
		double p;
		hibag_oracle_post_prob(haplo, n_haplo, n_hla, n_snp, &geno[s], 1, &p);
		loglik += geno[s].boot * log(p);
	}
	return loglik * -2;
}

void hibag_oracle_predict_avg(int n_hla, int n_classifier,
	const oracle_haplo_t *const *haplo, const int *n_haplo, const int *n_snp,
	const oracle_geno_t *geno, const double *weight, double *out_prob, double *out_match)
{
	const size_t nc = (size_t)n_hla * (n_hla + 1) / 2;
	double *post = (double *)malloc(nc * sizeof(double));
	double sum_w = 0, sum_matching = 0, num_matching = 0;
	memset(out_prob, 0, nc * sizeof(double));
	for (int c = 0; c < n_classifier; c++)
	{
		const double w = weight[c];
		if (w <= 0) continue;                            /* src/LibHLA.cpp:2451 */
		double pm;
		hibag_oracle_post_prob2(haplo[c], n_haplo[c], n_hla, n_snp[c], &geno[c], 1, post, &pm);
		sum_matching += pm * w;                          /* :2458-2459 */
		num_matching += w;
		for (size_t k = 0; k < nc; k++) out_prob[k] += post[k] * w;   /* :1497-1507 */
		sum_w += w;
	}
	if (sum_w > 0)                                       /* :1509-1518 */
	{
		const double ff = 1.0 / sum_w;
		for (size_t k = 0; k < nc; k++) out_prob[k] *= ff;
	}
	*out_match = sum_matching / num_matching;            /* :2480 (NaN when no classifier) */
	free(post);
}

void hibag_oracle_dosage(const double *prob, int n_hla, double *dosage)
{
	for (int h = 0; h < n_hla; h++) dosage[h] = 0;
	for (int h1 = 0; h1 < n_hla; h1++)
	{
		dosage[h1] += 2 * (*prob++);
		for (int h2 = h1 + 1; h2 < n_hla; h2++)
		{
			const double v = *prob++;
			dosage[h1] += v; dosage[h2] += v;
		}
	}
}

void hibag_oracle_int_to_snp(oracle_geno_t *out, int length, const int32_t *geno_base,
	const int32_t *index)
{
	/* everything missing first, then one SNP at a time (src/LibHLA.cpp:671-705) */
	out->s1[0] = out->s1[1] = 0;
	out->s2[0] = out->s2[1] = ~(uint64_t)0;
	for (int i = 0; i < length; i++)
	{
		const int g = geno_base[index[i]];
		const uint64_t bit = (uint64_t)1 << (i & 63);
		const int w = i >> 6;
		if (g == 0) { out->s2[w] &= ~bit; }
		else if (g == 1) { out->s1[w] |= bit; out->s2[w] &= ~bit; }
		else if (g == 2) { out->s1[w] |= bit; }
	}
}

void hibag_oracle_classifier_weights(int n_classifier, const int *n_snp,
	const int32_t *const *snpidx, int n_total_snp, const int32_t *geno_row, double *weight)
{
	int *snp_weight = (int *)calloc((size_t)n_total_snp, sizeof(int));
	for (int c = 0; c < n_classifier; c++)               /* src/LibHLA.cpp:2484-2496 */
		for (int i = 0; i < n_snp[c]; i++) snp_weight[snpidx[c][i]]++;
	for (int c = 0; c < n_classifier; c++)               /* src/LibHLA.cpp:2418-2431 */
	{
		int nw = 0, sum = 0;
		for (int i = 0; i < n_snp[c]; i++)
		{
			const int k = snpidx[c][i];
			sum += snp_weight[k];
			if (0 <= geno_row[k] && geno_row[k] <= 2) nw += snp_weight[k];
		}
		weight[c] = (sum > 0) ? ((double)nw / sum) : 0;
	}
	free(snp_weight);
}

/* The records a build_haplomatch plugin returns (src/LibHLA.cpp:1014-1072): for every in-bag
 * sample (ascending index k within the in-bag list) the haplotype pairs of its two true alleles
 * at the minimum Hamming distance on n_snp SNPs -- the selection rule of _PrepHaploMatch_def,
 * src/LibHLA.cpp:1569-1637, on the list BEFORE doubling -- as (k, (i2 << 16) | i1), i1 outer and
 * i2 inner, upper triangle when the two alleles are the same. out must hold 2 * max_records
 * uint32; returns the number of records (or -1 if max_records is too small). */
long hibag_oracle_haplomatch_records(const oracle_haplo_t *haplo, const int64_t *n_haplo,
	int n_hla, int n_snp, const oracle_geno_t *geno, int n_samp, uint32_t *out, long max_records)
{
	int *st = (int *)malloc(sizeof(int) * (size_t)(n_hla + 1));
	long n = 0;
	int k = 0;
	st[0] = 0;
	for (int a = 0; a < n_hla; a++) st[a + 1] = st[a] + (int)n_haplo[a];
	for (int s = 0; s < n_samp; s++)
	{
		if (geno[s].boot <= 0) continue;
		const oracle_geno_t *g = &geno[s];
		const int st1 = st[g->a1], m1 = st[g->a1 + 1] - st1;
		const int st2 = st[g->a2], m2 = st[g->a2 + 1] - st2;
		const int same = (st1 == st2);
		int min_d = n_snp * 4;
		for (int i = 0; i < m1; i++)
			for (int j = same ? i : 0; j < m2; j++)
			{
				const int d = hibag_oracle_hamming(g, &haplo[st1 + i], &haplo[st2 + j], n_snp);
				if (d < min_d) min_d = d;
			}
		for (int i = 0; i < m1; i++)
			for (int j = same ? i : 0; j < m2; j++)
				if (hibag_oracle_hamming(g, &haplo[st1 + i], &haplo[st2 + j], n_snp) == min_d)
				{
					if (n >= max_records) { free(st); return -1; }
					out[2 * n] = (uint32_t)k;
					out[2 * n + 1] = ((uint32_t)j << 16) | (uint32_t)i;
					n++;
				}
		k++;
	}
	free(st);
	return n;
}


/* PLINK BED -> genotype matrix: reference HIBAG_ConvBED, src/HIBAG.cpp:1094-1191 (prefix check
 * :1113-1117, pack geometry :1120-1137, code table {2, NA, 1, 0} :1141, unpack :1153-1169,
 * individual-major store :1172-1180, SNP-major store :1181-1192). `file` is the whole file
 * including the 3-byte prefix; out is int32 [n_samp][n_save] (the reference's INTEGER matrix
 * [n_save x n_samp], column-major) with NA = INT32_MIN. Returns 0, -1 for an invalid prefix, -2 when
 * the file is too short. */
int hibag_oracle_bed_decode(const uint8_t *file, long n_bytes, int n_samp, int n_snp,
	const int32_t *snp_flag, int n_save, int32_t *out)
{
	static const int32_t cvt[4] = { 2, INT32_MIN, 1, 0 };
	if (n_bytes < 3 || file[0] != 0x6C || file[1] != 0x1B) return -1;
	const int mode = file[2];
	int n_re, n_numpack, n_pack, n_num;
	if (mode == 0) { n_re = n_snp % 4; n_numpack = n_snp / 4; n_num = n_samp; }
	else { n_re = n_samp % 4; n_numpack = n_samp / 4; n_num = n_snp; }
	n_pack = (n_re > 0) ? (n_numpack + 1) : n_numpack;
	if (n_bytes < 3 + (long)n_pack * n_num) return -2;
	int32_t *dst = (int32_t *)malloc(sizeof(int32_t) * ((size_t)n_numpack + 1) * 4);
	int i_snp = 0;
	const uint8_t *src = file + 3;
	for (int i = 0; i < n_num; i++, src += n_pack)
	{
		int32_t *p = dst;
		for (int k = 0; k < n_numpack; k++)
		{
			unsigned char g = src[k];
			*p++ = cvt[g & 0x03]; g >>= 2;
			*p++ = cvt[g & 0x03]; g >>= 2;
			*p++ = cvt[g & 0x03]; g >>= 2;
			*p++ = cvt[g & 0x03];
		}
		if (n_re > 0)
		{
			unsigned char g = src[n_numpack];
			for (int k = 0; k < n_re; k++) { *p++ = cvt[g & 0x03]; g >>= 2; }
		}
		if (mode == 0)
		{
			int32_t *pi = out + (size_t)i * n_save;
			for (int j = 0; j < n_snp; j++)
				if (snp_flag == NULL || snp_flag[j]) *pi++ = dst[j];
		} else if (snp_flag == NULL || snp_flag[i])
		{
			int32_t *pi = out + i_snp;
			i_snp++;
			for (int j = 0; j < n_samp; j++) { *pi = dst[j]; pi += n_save; }
		}
	}
	free(dst);
	return 0;
}
