// hostalg.h -- host side of the attribute-bagging trainer: R-compatible RNG, bootstrap,
// candidate sampling, haplotype-pair preparation, EM frequency estimation, rare-haplotype
// pruning. Own C++ restatement of the reference's host algorithm (CVariableSelection /
// CAlg_EM / CHaplotypeList, src/LibHLA.cpp:352-578, 930-1255, 1837-2122), re-organised so that
// the <= mtry candidate SNPs of one selection round can be estimated concurrently: haplotype
// pairs are index pairs into a per-round doubled list and every candidate works on its own
// frequency arrays. Arithmetic order per candidate is exactly the reference's, so results
// are bit-identical for any thread count.
#pragma once

#include <cstdint>
#include <vector>

#include "../../include/hibag_b200.h"

namespace hb {

/// R's default generator (Mersenne-Twister, R src/main/RNG.c) with set.seed() scrambling
class RRng
{
public:
	RRng();
	void set_seed(uint32_t seed);
	double unif_rand();
	/// uniform integer in [0, n): reference RandomNum, src/LibHLA.cpp:120-126
	int random_num(int n);
	uint64_t draws() const { return draws_; }
private:
	uint32_t mt_[624];
	int idx_;
	uint64_t draws_;
	void sgenrand(uint32_t seed);
};

/// packed genotype of one sample on the host (two bit planes, 128 SNP slots)
struct HostGeno
{
	uint64_t s1[2], s2[2];
};

/// haplotype list: records grouped by allele + count per allele
struct HapList
{
	int n_snp = 0;
	std::vector<hibag_haplotype> h;
	std::vector<int> len;           // haplotypes per HLA allele (LenPerHLA)
	void set_tags();                // fill hla_allele / freq_f32 (SetHaploAux_GPU, :565-578)
};

/// Hamming distance, reference src/LibHLA.cpp:802-817
int hamming(const HostGeno &g, const int64_t h1[2], const int64_t h2[2], int n_snp);

/// candidate-SNP pool with the reference's tail selection (CSamplingWithoutReplace, :930-993)
class SnpPool
{
public:
	void init(int n);
	int total() const { return (int)idx_.size(); }
	void random_select(int m_try, RRng &rng);
	int n_selected() const { return m_; }
	int &at(int i) { return idx_[idx_.size() - m_ + i]; }
	void remove(int i);
	void remove_selection();
	void remove_flagged();
private:
	std::vector<int> idx_;
	int m_ = 0;
};

/// the per-round state shared by all candidates (result of CAlg_EM::PrepareHaplotypes, :1002)
struct RoundPairs
{
	int n_cur = 0;                       // haplotypes before doubling
	std::vector<int> samp;               // in-bag sample index per entry
	std::vector<int> boot;               // its bootstrap count
	std::vector<size_t> off;             // pair range [off[k], off[k+1]) of entry k
	std::vector<int> p1, p2;             // indices into the doubled list, p1 <= p2
};

/// per-candidate scratch (owned by one worker at a time)
struct EmScratch
{
	std::vector<double> freq, old;       // doubled list frequencies
	std::vector<int> cp1, cp2;           // pairs compatible with the candidate SNP (compacted)
	std::vector<size_t> coff;            // their range per in-bag entry
	std::vector<double> gf;              // GenoFreq of one sample's pairs
};

/// pairs at minimum distance for every in-bag sample on the current SNP set
/// (reference :1076-1123 with _PrepHaploMatch_def :1569-1637). `parallel_for(n, fn)` runs
/// fn(begin, end) over index ranges; output is independent of how ranges are split.
void prepare_round(const HapList &cur, const std::vector<HostGeno> &geno,
	const std::vector<int> &a1, const std::vector<int> &a2, const std::vector<int> &boot,
	const std::vector<int> &inbag, RoundPairs &out,
	void (*parallel_for)(void *ctx, int n, void (*fn)(void *arg, int begin, int end), void *arg),
	void *pf_ctx);

/// PrepareNewSNP + ExpectationMaximization + EraseDoubleHaplos for one candidate
/// (reference :1127-1255, :444-515). snp_col: genotypes of the candidate SNP for all samples.
/// Returns false when the SNP is monomorphic in the bag (no list produced).
bool estimate_candidate(const HapList &cur, const RoundPairs &rp, const int8_t *snp_col,
	int n_samp_total, double rare_prob, EmScratch &scr, HapList &out);

/// EraseDoubleHaplos (:461-515) + renormalisation of a doubled list with frequencies
/// freq[2 * n_cur] (from the host or the device EM) into the candidate's haplotype list
void finish_candidate(const HapList &cur, const double *freq, double rare_prob, HapList &out);

}  // namespace hb
