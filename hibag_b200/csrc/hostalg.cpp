// hostalg.cpp -- see hostalg.h. Compiled with -ffp-contract=off: every multiply and add below is
// a separately rounded IEEE operation, as in the reference's base build.
#include "hostalg.h"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <stdexcept>

namespace hb {

// ---------------------------------------------------------------------------------------------
// RNG
// ---------------------------------------------------------------------------------------------
RRng::RRng() : idx_(625), draws_(0) { memset(mt_, 0, sizeof(mt_)); }

void RRng::sgenrand(uint32_t seed)
{
	for (int i = 0; i < 624; i++)
	{
		mt_[i] = seed & 0xffff0000u;
		seed = 69069u * seed + 1;
		mt_[i] |= (seed & 0xffff0000u) >> 16;
		seed = 69069u * seed + 1;
	}
	idx_ = 624;
}

void RRng::set_seed(uint32_t seed)
{
	// R: RNG_Init -- 50 scrambling steps of the LCG, one more for the position slot (which
	// FixupSeeds then overwrites with 624), then the 624 state words
	for (int j = 0; j < 50; j++) seed = 69069u * seed + 1;
	seed = 69069u * seed + 1;
	for (int j = 0; j < 624; j++)
	{
		seed = 69069u * seed + 1;
		mt_[j] = seed;
	}
	idx_ = 624;
	draws_ = 0;
}

double RRng::unif_rand()
{
	static const uint32_t mag01[2] = { 0x0u, 0x9908b0dfu };
	const int N = 624, M = 397;
	uint32_t y;
	if (idx_ >= N)
	{
		if (idx_ == N + 1) sgenrand(4357);
		int k;
		for (k = 0; k < N - M; k++)
		{
			y = (mt_[k] & 0x80000000u) | (mt_[k + 1] & 0x7fffffffu);
			mt_[k] = mt_[k + M] ^ (y >> 1) ^ mag01[y & 1u];
		}
		for (; k < N - 1; k++)
		{
			y = (mt_[k] & 0x80000000u) | (mt_[k + 1] & 0x7fffffffu);
			mt_[k] = mt_[k + (M - N)] ^ (y >> 1) ^ mag01[y & 1u];
		}
		y = (mt_[N - 1] & 0x80000000u) | (mt_[0] & 0x7fffffffu);
		mt_[N - 1] = mt_[M - 1] ^ (y >> 1) ^ mag01[y & 1u];
		idx_ = 0;
	}
	y = mt_[idx_++];
	y ^= (y >> 11);
	y ^= (y << 7) & 0x9d2c5680u;
	y ^= (y << 15) & 0xefc60000u;
	y ^= (y >> 18);
	draws_++;
	double x = (double)y * 2.3283064365386963e-10;
	// R's fixup(): keep the value inside the open interval (0,1)
	const double i2_32m1 = 2.328306437080797e-10;
	if (x <= 0.0) return 0.5 * i2_32m1;
	if ((1.0 - x) <= 0.0) return 1.0 - 0.5 * i2_32m1;
	return x;
}

int RRng::random_num(int n)
{
	int v = (int)(n * unif_rand());
	if (v >= n) v = n - 1;
	return v;
}

// ---------------------------------------------------------------------------------------------
// small pieces
// ---------------------------------------------------------------------------------------------
void HapList::set_tags()
{
	size_t k = 0;
	for (size_t a = 0; a < len.size(); a++)
		for (int m = len[a]; m > 0; m--, k++)
		{
			h[k].hla_allele = (int32_t)a;
			h[k].freq_f32 = (float)h[k].freq;
		}
}

int hamming(const HostGeno &g, const int64_t h1[2], const int64_t h2[2], int n_snp)
{
	const int words = (n_snp <= 64) ? 1 : 2;
	int d = 0;
	for (int w = 0; w < words; w++)
	{
		const uint64_t A = (uint64_t)h1[w], B = (uint64_t)h2[w];
		const uint64_t S1 = g.s1[w], S2 = g.s2[w];
		const uint64_t miss = S2 & ~S1;
		const uint64_t mask = ((A ^ S2) | (B ^ S1)) & ~miss;
		d += __builtin_popcountll((A ^ S1) & mask) + __builtin_popcountll((B ^ S2) & mask);
	}
	return d;
}

void SnpPool::init(int n)
{
	m_ = 0;
	idx_.resize(n);
	for (int i = 0; i < n; i++) idx_[i] = i;
}

void SnpPool::random_select(int m_try, RRng &rng)
{
	const int n = (int)idx_.size();
	if (m_try > n) m_try = n;
	if (m_try < n)      // no draw at all when everything is taken (:953)
	{
		for (int i = 0; i < m_try; i++)
		{
			const int pick = rng.random_num(n - i);
			std::swap(idx_[pick], idx_[n - i - 1]);
		}
	}
	m_ = m_try;
}

void SnpPool::remove(int i)
{
	idx_.erase(idx_.begin() + (idx_.size() - m_ + i));
}

void SnpPool::remove_selection()
{
	idx_.resize(idx_.size() - m_);
}

void SnpPool::remove_flagged()
{
	const int n = (int)idx_.size();
	for (int i = n - 1; i >= n - m_; i--)
		if (idx_[i] < 0) idx_.erase(idx_.begin() + i);
}

// ---------------------------------------------------------------------------------------------
// PrepareHaplotypes
// ---------------------------------------------------------------------------------------------
namespace {

const int PREP_GRAIN = 32;     // in-bag entries per work item

struct PrepBlock               // pairs found for entries [b*PREP_GRAIN, ...), concatenated
{
	std::vector<int> p1, p2;
	std::vector<int> count;    // pairs per entry of the block
};

struct PrepJob
{
	const HapList *cur;
	const std::vector<HostGeno> *geno;
	const std::vector<int> *a1, *a2, *inbag;
	std::vector<int> start;                          // first haplotype of each allele in cur
	std::vector<PrepBlock> *blocks;
	RoundPairs *out;
};

void prep_range(void *arg, int bbegin, int bend)
{
	PrepJob &J = *(PrepJob *)arg;
	const HapList &cur = *J.cur;
	const int n_snp = cur.n_snp;
	const int n_ib = (int)J.inbag->size();
	std::vector<short> dist;
	for (int b = bbegin; b < bend; b++)
	{
		PrepBlock &blk = (*J.blocks)[b];
		blk.p1.clear(); blk.p2.clear(); blk.count.clear();
		const int kend = std::min(n_ib, (b + 1) * PREP_GRAIN);
		for (int k = b * PREP_GRAIN; k < kend; k++)
		{
			const int s = (*J.inbag)[k];
			const HostGeno &g = (*J.geno)[s];
			const int A1 = (*J.a1)[s], A2 = (*J.a2)[s];
			const int st1 = J.start[A1], m1 = cur.len[A1];
			const int st2 = J.start[A2], m2 = cur.len[A2];
			const size_t before = blk.p1.size();
			int min_d = n_snp * 4;
			// the doubled list holds haplotype k as entries 2k (new allele 0) and 2k+1 (1); on
			// the current SNPs both copies are at the same distance: distances are taken on `cur`
			if (st1 != st2)
			{
				dist.resize((size_t)m1 * m2);
				for (int i = 0; i < m1; i++)
					for (int j = 0; j < m2; j++)
					{
						const int d = hamming(g, cur.h[st1 + i].packed, cur.h[st2 + j].packed, n_snp);
						dist[(size_t)i * m2 + j] = (short)d;
						if (d < min_d) min_d = d;
					}
				// scan order of the doubled ranges: first index outer, second inner (:1578-1604);
				// with minimum 0 the zero-distance pairs are exactly the pairs at the minimum
				for (int i = 0; i < 2 * m1; i++)
				{
					const short *row = &dist[(size_t)(i >> 1) * m2];
					for (int j = 0; j < 2 * m2; j++)
						if (row[j >> 1] == min_d)
						{
							blk.p1.push_back(2 * st1 + i); blk.p2.push_back(2 * st2 + j);
						}
				}
			} else {
				// same start => same allele range: upper triangle including i == j (:1608-1634)
				dist.resize((size_t)m1 * m1);
				for (int i = 0; i < m1; i++)
					for (int j = i; j < m1; j++)
					{
						const int d = hamming(g, cur.h[st1 + i].packed, cur.h[st1 + j].packed, n_snp);
						dist[(size_t)i * m1 + j] = (short)d;
						if (d < min_d) min_d = d;
					}
				for (int i = 0; i < 2 * m1; i++)
				{
					const short *row = &dist[(size_t)(i >> 1) * m1];
					for (int j = i; j < 2 * m1; j++)
						if (row[j >> 1] == min_d)
						{
							blk.p1.push_back(2 * st1 + i); blk.p2.push_back(2 * st1 + j);
						}
				}
			}
			blk.count.push_back((int)(blk.p1.size() - before));
		}
	}
}

void prep_copy(void *arg, int bbegin, int bend)
{
	PrepJob &J = *(PrepJob *)arg;
	RoundPairs &out = *J.out;
	for (int b = bbegin; b < bend; b++)
	{
		const PrepBlock &blk = (*J.blocks)[b];
		const size_t o = out.off[(size_t)b * PREP_GRAIN];
		if (!blk.p1.empty())
		{
			memcpy(&out.p1[o], blk.p1.data(), sizeof(int) * blk.p1.size());
			memcpy(&out.p2[o], blk.p2.data(), sizeof(int) * blk.p2.size());
		}
	}
}

}  // namespace

void prepare_round(const HapList &cur, const std::vector<HostGeno> &geno,
	const std::vector<int> &a1, const std::vector<int> &a2, const std::vector<int> &boot,
	const std::vector<int> &inbag, RoundPairs &out,
	void (*parallel_for)(void *ctx, int n, void (*fn)(void *arg, int begin, int end), void *arg),
	void *pf_ctx)
{
	if (cur.n_snp >= HIBAG_B200_MAX_SNP)
		throw std::runtime_error("prepare_round: too many SNP markers in the classifier");
	const int n_ib = (int)inbag.size();
	const int n_blocks = (n_ib + PREP_GRAIN - 1) / PREP_GRAIN;
	static thread_local std::vector<PrepBlock> blocks;     // reused across rounds (one trainer thread)
	if ((int)blocks.size() < n_blocks) blocks.resize(n_blocks);
	PrepJob J;
	J.cur = &cur; J.geno = &geno; J.a1 = &a1; J.a2 = &a2; J.inbag = &inbag;
	J.start.assign(cur.len.size() + 1, 0);
	for (size_t a = 0; a < cur.len.size(); a++) J.start[a + 1] = J.start[a] + cur.len[a];
	J.blocks = &blocks;
	J.out = &out;
	parallel_for(pf_ctx, n_blocks, prep_range, &J);

	out.n_cur = (int)cur.h.size();
	out.samp.resize(n_ib); out.boot.resize(n_ib); out.off.resize(n_ib + 1);
	size_t total = 0;
	for (int b = 0; b < n_blocks; b++)
	{
		const PrepBlock &blk = blocks[b];
		for (size_t t = 0; t < blk.count.size(); t++)
		{
			const int k = b * PREP_GRAIN + (int)t;
			out.samp[k] = inbag[k];
			out.boot[k] = boot[inbag[k]];
			out.off[k] = total;
			total += blk.count[t];
		}
	}
	out.off[n_ib] = total;
	out.p1.resize(total); out.p2.resize(total);
	parallel_for(pf_ctx, n_blocks, prep_copy, &J);
}

// ---------------------------------------------------------------------------------------------
// one candidate SNP: initial frequencies, pair flags, EM, pruning of the doubled list
// ---------------------------------------------------------------------------------------------
bool estimate_candidate(const HapList &cur, const RoundPairs &rp, const int8_t *snp_col,
	int n_samp_total, double rare_prob, EmScratch &scr, HapList &out)
{
	static const double EM_INIT_VAL_FRAC = 0.001;          // src/LibHLA.cpp:100
	static const int EM_MAX_ITER = 500;                    // :98
	const double EM_RELTOL = std::sqrt(DBL_EPSILON);       // :102
	const int n_entry = (int)rp.samp.size();
	const int n_cur = rp.n_cur;
	const int n2 = 2 * n_cur;

	// allele frequency of the new SNP in the bootstrap sample (:1136-1151)
	int allele_cnt = 0, valid_cnt = 0;
	for (int k = 0; k < n_entry; k++)
	{
		const int g = snp_col[rp.samp[k]];
		if (0 <= g && g <= 2)
		{
			allele_cnt += g * rp.boot[k];
			valid_cnt += 2 * rp.boot[k];
		}
	}
	if (allele_cnt == 0 || allele_cnt == valid_cnt) return false;

	// doubled list, initial frequencies (:444-459)
	scr.freq.resize(n2); scr.old.resize(n2);
	double *freq = scr.freq.data(), *old = scr.old.data();
	{
		const double af = double(allele_cnt) / valid_cnt;
		const double p0 = 1 - af, p1 = af;
		for (int k = 0; k < n_cur; k++)
		{
			freq[2 * k] = p0 * cur.h[k].freq + EM_INIT_VAL_FRAC;
			freq[2 * k + 1] = p1 * cur.h[k].freq + EM_INIT_VAL_FRAC;
		}
	}
	// which pairs are compatible with the new genotype (:1157-1180). The flags never change
	// during EM, so the compatible pairs are compacted once; the loops below then visit exactly
	// the pairs the reference visits with Flag == true, in the same order.
	const int *P1 = rp.p1.data(), *P2 = rp.p2.data();
	scr.cp1.clear(); scr.cp2.clear();
	scr.coff.resize(n_entry + 1);
	size_t max_pairs = 0;
	for (int k = 0; k < n_entry; k++)
	{
		const int g = snp_col[rp.samp[k]];
		const size_t b = rp.off[k], e = rp.off[k + 1];
		scr.coff[k] = scr.cp1.size();
		if (0 <= g && g <= 2)
		{
			for (size_t t = b; t < e; t++)
				if (((P1[t] & 1) + (P2[t] & 1)) == g) { scr.cp1.push_back(P1[t]); scr.cp2.push_back(P2[t]); }
		} else {
			for (size_t t = b; t < e; t++) { scr.cp1.push_back(P1[t]); scr.cp2.push_back(P2[t]); }
		}
		max_pairs = std::max(max_pairs, scr.cp1.size() - scr.coff[k]);
	}
	scr.coff[n_entry] = scr.cp1.size();
	scr.gf.resize(max_pairs + 1);
	const int *C1 = scr.cp1.data(), *C2 = scr.cp2.data();
	const size_t *coff = scr.coff.data();
	const int *boot = rp.boot.data();

	// EM (:1185-1255). The reference runs the E-step over all samples and then accumulates the
	// new frequencies sample by sample; the two loops are fused per sample here, which leaves the
	// sequence of additions into LogLik and into every Freq unchanged.
	double *gf = scr.gf.data();
	double conv_tol = 0, loglik = -1e+30;
	const double scale = 0.5 / n_samp_total;
	for (int iter = 0; iter <= EM_MAX_ITER; iter++)
	{
		const double old_loglik = loglik;
		for (int i = 0; i < n2; i++) { old[i] = freq[i]; freq[i] = 0; }
		loglik = 0;
		for (int k = 0; k < n_entry; k++)
		{
			const size_t b = coff[k], e = coff[k + 1];
			double psum = 0;
			for (size_t t = b; t < e; t++)
			{
				const int u = C1[t], v = C2[t];
				const double x = (u != v) ? (2 * old[u] * old[v]) : (old[u] * old[v]);
				gf[t - b] = x;
				psum += x;
			}
			loglik += boot[k] * std::log(psum);
			psum = boot[k] / psum;
			for (size_t t = b; t < e; t++)
			{
				const double r = gf[t - b] * psum;
				freq[C1[t]] += r; freq[C2[t]] += r;
			}
		}
		for (int i = 0; i < n2; i++) freq[i] *= scale;
		if (iter > 0)
		{
			if (std::fabs(loglik - old_loglik) <= conv_tol) break;
		} else {
			conv_tol = EM_RELTOL * (std::fabs(loglik) + EM_RELTOL);
			if (conv_tol < 0) conv_tol = 0;
		}
	}

	finish_candidate(cur, freq, rare_prob, out);
	return true;
}

void finish_candidate(const HapList &cur, const double *freq, double rare_prob, HapList &out)
{
	static const double MIN_RARE_FREQ = 1e-5;              // LibHLA_ext.h:230
	// drop / merge rare members of each doubled pair (:461-515)
	out.n_snp = cur.n_snp + 1;
	out.h.clear();
	out.len.assign(cur.len.size(), 0);
	const int bit = cur.n_snp;
	const int bw = bit >> 6;
	const uint64_t bm = (uint64_t)1 << (bit & 63);
	double sum = 0;
	int k = 0;
	for (size_t a = 0; a < cur.len.size(); a++)
	{
		int num = 0;
		for (int m = cur.len[a]; m > 0; m--, k++)
		{
			const double f0 = freq[2 * k], f1 = freq[2 * k + 1];
			const double sumfreq = f0 + f1;
			hibag_haplotype h0 = cur.h[k], h1 = cur.h[k];
			h0.packed[bw] = (int64_t)((uint64_t)h0.packed[bw] & ~bm);
			h1.packed[bw] = (int64_t)((uint64_t)h1.packed[bw] | bm);
			if (f0 < rare_prob || f1 < rare_prob)
			{
				if (sumfreq >= MIN_RARE_FREQ)
				{
					hibag_haplotype keep = (f0 >= f1) ? h0 : h1;
					keep.freq = sumfreq;
					out.h.push_back(keep);
					sum += sumfreq;
					num++;
				}
			} else {
				h0.freq = f0; h1.freq = f1;
				out.h.push_back(h0);
				out.h.push_back(h1);
				sum += sumfreq;
				num += 2;
			}
		}
		out.len[a] = num;
	}
	const double sc = 1 / sum;
	for (size_t i = 0; i < out.h.size(); i++) out.h[i].freq *= sc;
	out.set_tags();
}

}  // namespace hb
