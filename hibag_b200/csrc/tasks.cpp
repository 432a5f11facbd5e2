// tasks.cpp -- see tasks.h
#include "tasks.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <mutex>
#include <vector>

namespace hb {

static double g_table[2 * HIBAG_B200_MAX_SNP + 1];
static int g_first_zero = 0;
static std::once_flag g_table_once;

static double g_floor_table[2 * HIBAG_B200_MAX_SNP + 1];
static double g_bound_factor = 0, g_bound_factor2 = 0;

static void init_table()
{
	// exp(d * log(MIN_RARE_FREQ)), [0] = 1, non-finite -> 0   (src/LibHLA.cpp:175-183)
	const int n = 2 * HIBAG_B200_MAX_SNP;
	for (int d = 0; d <= n; d++)
	{
		double v = std::exp(d * std::log(1e-5));
		if (d == 0) v = 1;
		if (!std::isfinite(v)) v = 0;
		g_table[d] = v;
	}
	int z = n + 1;
	while (z > 0 && g_table[z - 1] == 0) z--;
	g_first_zero = z;
	// screening: floored table and bound factor
	for (int d = 0; d <= n; d++) g_floor_table[d] = (g_table[d] > 1e-100) ? g_table[d] : 1e-100;
	std::vector<double> suffix_max(n + 2, 0.0);
	for (int d = n; d >= 0; d--) suffix_max[d] = std::max(g_table[d], suffix_max[d + 1]);
	double kappa = 1.0;
	for (int p = 0; p <= n; p++)
		for (int q = 0; p + q <= n; q++)
		{
			const double r = suffix_max[p + q] / (g_floor_table[p] * g_floor_table[q]);
			if (r > kappa) kappa = r;
		}
	g_bound_factor = 2.0 * kappa * (1.0 + 1e-12) * (1.0 + 1e-8);
	// second level (classes of the first two heterozygous SNPs, screen.cu: screen_refine_kernel):
	// T[d] <= kappa3 * T'[p] T'[q] T'[r] for every d >= p + q + r, r <= 2
	double kappa3 = 1.0;
	for (int p = 0; p <= n; p++)
		for (int q = 0; p + q <= n; q++)
			for (int r = 0; r <= 2 && p + q + r <= n; r++)
			{
				const double v = suffix_max[p + q + r] / (g_floor_table[p] * g_floor_table[q] * g_floor_table[r]);
				if (v > kappa3) kappa3 = v;
			}
	g_bound_factor2 = 2.0 * kappa3 * (1.0 + 1e-12) * (1.0 + 1e-8);
}

const double *host_rare_freq_floor_table()
{
	std::call_once(g_table_once, init_table);
	return g_floor_table;
}

double screen_bound_factor()
{
	std::call_once(g_table_once, init_table);
	return g_bound_factor;
}

double screen_bound_factor2()
{
	std::call_once(g_table_once, init_table);
	return g_bound_factor2;
}

const double *host_rare_freq_table()
{
	std::call_once(g_table_once, init_table);
	return g_table;
}

int rare_freq_first_zero()
{
	std::call_once(g_table_once, init_table);
	return g_first_zero;
}

int table_rows_for(int n_snp)
{
	// distances range over 0..2*n_snp; every row >= first_zero is 0 and shares one row
	const int fz = rare_freq_first_zero();
	int rows = 2 * n_snp + 1;
	if (rows > fz + 1) rows = fz + 1;
	if (rows < 1) rows = 1;
	return rows;
}

static inline size_t align16(size_t x) { return (x + 15) & ~(size_t)15; }

size_t list_blob_capacity(int n_hap, int n_snp, int n_hla)
{
	const size_t n_cells = (size_t)n_hla * (n_hla + 1) / 2;
	return align16((size_t)n_hap * hap_record_bytes(n_snp)) + n_cells * sizeof(CellTask) +
		align16(n_cells * sizeof(Chunk)) + 64;
}

ListBlob build_list_blob(const hibag_haplotype *haplo, int n_hap, int n_hla, int n_snp,
	void *dst, int target_chunks)
{
	if (n_snp < 0 || n_snp > HIBAG_B200_MAX_SNP)
		throw std::runtime_error("build_list_blob: n_snp out of range");
	if (n_hla <= 0) throw std::runtime_error("build_list_blob: n_hla must be positive");

	ListBlob b;
	b.n_hap = n_hap; b.n_snp = n_snp; b.n_hla = n_hla;
	b.n_dist = table_rows_for(n_snp);

	// ---- haplotype records in kernel layout (bits >= n_snp may be garbage in the source:
	// THaplotype's ctor does not clear them, src/LibHLA.cpp:276-279; the genotype's missing
	// flags mask them in the kernel exactly as in the reference) -------------------------
	const int rec = hap_record_bytes(n_snp);
	unsigned char *out = (unsigned char *)dst;
	std::vector<int> start(n_hla + 1, 0);
	int prev = 0;
	for (int i = 0; i < n_hap; i++)
	{
		const int a = haplo[i].hla_allele;
		if (a < prev || a >= n_hla)
			throw std::runtime_error("haplotype list is not grouped by HLA allele");
		prev = a;
		start[a + 1]++;
		if (rec == 16)
		{
			memcpy(out + (size_t)i * 16, &haplo[i].packed[0], 8);
			memcpy(out + (size_t)i * 16 + 8, &haplo[i].freq, 8);
		} else {
			memcpy(out + (size_t)i * 32, &haplo[i].packed[0], 16);
			memcpy(out + (size_t)i * 32 + 16, &haplo[i].freq, 8);
			memset(out + (size_t)i * 32 + 24, 0, 8);
		}
	}
	for (int a = 0; a < n_hla; a++) start[a + 1] += start[a];

	// ---- cells -------------------------------------------------------------------------
	const size_t n_cells = (size_t)n_hla * (n_hla + 1) / 2;
	b.n_cells = (int)n_cells;
	b.off_cells = align16((size_t)n_hap * rec);
	CellTask *cells = (CellTask *)(out + b.off_cells);
	std::vector<std::pair<uint64_t, int> > order(n_cells);   // (cost, cell index)
	std::vector<CellTask> tmp(n_cells);
	uint64_t total = 0;
	{
		int idx = 0;
		for (int a = 0; a < n_hla; a++)
		{
			const int na = start[a + 1] - start[a];
			for (int c = a; c < n_hla; c++, idx++)
			{
				const int nb = start[c + 1] - start[c];
				CellTask &t = tmp[idx];
				t.a_start = start[a]; t.a_n = na;
				t.b_start = start[c]; t.b_n = nb;
				t.out_idx = idx; t.diag = (a == c) ? 1 : 0;
				t.al_a = a; t.al_b = c;
				const uint64_t pairs = (a == c) ? (uint64_t)na * (na + 1) / 2 : (uint64_t)na * nb;
				total += pairs;
				order[idx] = std::make_pair(pairs + 2 * (uint64_t)na + 4, idx);
			}
		}
	}
	b.pairs_per_sample = total;
	std::sort(order.begin(), order.end(),
		[](const std::pair<uint64_t, int> &x, const std::pair<uint64_t, int> &y)
		{ return (x.first != y.first) ? (x.first > y.first) : (x.second < y.second); });
	uint64_t cost_sum = 0;
	for (size_t k = 0; k < n_cells; k++)
	{
		cells[k] = tmp[order[k].second];
		cost_sum += order[k].first;
	}

	// ---- chunks: consecutive cells (in decreasing-cost order) up to a target cost ----------
	b.off_chunks = b.off_cells + n_cells * sizeof(CellTask);
	Chunk *chunks = (Chunk *)(out + b.off_chunks);
	if (target_chunks < 1) target_chunks = 1;
	uint64_t target = cost_sum / (uint64_t)target_chunks;
	if (target < 96) target = 96;
	int n_chunks = 0;
	size_t k = 0;
	while (k < n_cells)
	{
		uint64_t acc = 0;
		const size_t k0 = k;
		while (k < n_cells && (acc == 0 || acc + order[k].first <= target))
		{
			acc += order[k].first;
			k++;
		}
		chunks[n_chunks].cell_begin = (int)k0;
		chunks[n_chunks].cell_end = (int)k;
		n_chunks++;
	}
	b.n_chunks = n_chunks;
	b.bytes = align16(b.off_chunks + (size_t)n_chunks * sizeof(Chunk));
	return b;
}

void bind_list(const ListBlob &b, const void *dev_blob, const double *dev_table, CellPass &p)
{
	const unsigned char *d = (const unsigned char *)dev_blob;
	p.hap = d;
	p.n_hap = b.n_hap;
	p.n_snp = b.n_snp;
	p.table = dev_table;
	p.n_dist = b.n_dist;
	p.cells = (const CellTask *)(d + b.off_cells);
	p.chunks = (const Chunk *)(d + b.off_chunks);
	p.n_chunks = b.n_chunks;
}

}  // namespace hb
