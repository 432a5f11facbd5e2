// em.h -- device-resident selection round: haplotype-pair matching (reference
// CAlg_EM::PrepareHaplotypes / _PrepHaploMatch_def, src/LibHLA.cpp:1002-1123, 1569-1637) and
// the EM frequency estimation of all candidate SNPs of a round (PrepareNewSNP +
// ExpectationMaximization, :1127-1255) on the GPU, SURVEY.md section 8f rows 2 and 3 (internal).
//
// Bit-exactness. Every frequency accumulator receives its contributions in the reference's
// order (in-bag sample order, pair order, H1 before H2): one thread owns one haplotype and sums a
// contiguous, incidence-ordered array sequentially with un-fused fp64 adds. The only value that
// cannot be reproduced to the bit on the device is the log-likelihood (glibc log vs device log);
// it only feeds the stopping test |LL - LL_old| <= tol. A candidate whose stopping test comes
// within a relative guard band of the tolerance in any iteration is reported as ambiguous and
// is re-estimated on the host by the caller, so the iteration count -- and with it every
// frequency -- always equals the reference's.
#pragma once

#include <cstdint>
#include <vector>

#include "common.h"
#include "hostalg.h"

namespace hb {

/// status of one candidate after the device EM
enum { EM_INVALID = 0, EM_OK = 1, EM_AMBIGUOUS = 2 };

class RoundEM
{
public:
	RoundEM();
	~RoundEM();

	/// can the device EM hold a round with n_cur haplotypes and n_entry in-bag samples?
	/// (16-bit haplotype indices in the packed pairs; shared-memory budget of one CTA)
	static bool supports(int n_cur, int n_entry)
	{
		return 2 * (size_t)n_cur <= 65535 &&
			8 * (4 * (size_t)n_cur + 64 + (size_t)n_entry) + 4 * (size_t)n_entry + 16 + 16384 <= 220 * 1024;
	}

	/// Pair matching for the in-bag entries on the current SNP set + incidence structure.
	/// planes: SoA 32-bit words s1[w*stride + sample] (w < 4); ib = in-bag sample indices
	/// (device), boot = bootstrap count per SAMPLE (device). Synchronises `st` once (the pair
	/// total sizes the buffers).
	void prepare(const HapList &cur, const uint32_t *s1, const uint32_t *s2, int stride,
		const int *a1, const int *a2, const int *ib, int n_entry, const int *boot,
		cudaStream_t st);

	/// EM of m candidates (raw genotype columns geno_t + snp*n_samp) in one launch; blocks.
	/// freq(i) = final doubled-list frequencies (2*n_cur), status(i), iterations(i).
	void run_em(const int *cand_snp, int m, const int8_t *geno_t, int n_samp, cudaStream_t st);
	const double *freq(int i) const { return h_freq_.get() + (size_t)i * n2_; }
	int status(int i) const { return h_status_.get()[4 * i]; }
	/// test hook: report candidate i as undecided so that the caller's host fallback runs
	void force_ambiguous(int i) { if (h_status_.get()[4 * i] == EM_OK) h_status_.get()[4 * i] = EM_AMBIGUOUS; }
	int iterations(int i) const { return h_status_.get()[4 * i + 1]; }
	uint64_t chain_launches = 0;      // run_em calls served by em_chain_kernel
	uint64_t sum_iterations = 0, sum_chain_adds = 0, sum_pair_updates = 0;   // over run_em calls (see train stats)

	/// the pair lists as the host algorithm holds them (for the host fallback / tests)
	void fetch_pairs(RoundPairs &out, const std::vector<int> &inbag, const std::vector<int> &boot,
		cudaStream_t st);

	/// lanes sharing the GPU (> 1 selects the dense EM shape, see run_em)
	void set_lanes(int n) { n_dense_lanes_ = n; }
	int n_entry() const { return n_entry_; }
	size_t total_pairs() const { return total_pairs_; }
	size_t ell_slots() const { return n_slots_; }
	int max_chain() const { return max_chain_; }
	/// an in-bag sample without any haplotype pair (one of its alleles lost all haplotypes): the
	/// caller estimates such a round on the host, whose degenerate arithmetic is the reference's
	bool has_empty_entry() const { return has_empty_entry_; }
	double kernel_ms = 0;          // summed CUDA-event time of the EM launches
	uint64_t launches = 0, h2d_bytes = 0, d2h_bytes = 0;

private:
	int n_entry_ = 0, n_cur_ = 0, n2_ = 0, n_snp_ = 0, n_dense_lanes_ = 1;
	size_t total_pairs_ = 0;
	const int *ib_ = nullptr, *boot_ = nullptr;
	PinBuf<unsigned char> h_stage_;
	DevBuf<uint64_t> d_hap_;          // [n_cur][2]
	DevBuf<double> d_curfreq_;        // [n_cur]
	DevBuf<int> d_start_;             // [n_hla + 1]
	DevBuf<int> d_cnt_, d_off_, d_mind_;
	DevBuf<int> d_p1_, d_p2_;
	DevBuf<int> d_key_, d_val_, d_key2_, d_val2_, d_inc_off_;
	DevBuf<int> d_len_, d_hapid_, d_len2_, d_hap_sorted_, d_rank_, d_group_len_, d_group_base_;
	DevBuf<int> d_pairs4_;            // int4 {u, v, slot_u, slot_v} per pair
	size_t n_slots_ = 0;
	int max_chain_ = 0;
	bool has_empty_entry_ = false;
	DevBuf<unsigned char> d_tmp_;     // cub temporary storage
	PinBuf<int> h_total_;
	DevBuf<int> d_cand_;
	PinBuf<int> h_cand_;
	DevBuf<double> d_freq_, d_xbuf_, d_rinc_;
	DevBuf<int> d_pmap_, d_cuv_, d_coff_, d_glen_;      // per-candidate compaction (em_kernel)
	DevBuf<int> d_idxell_, d_erec_;                     // per-candidate chain records (em_chain_kernel)
	DevBuf<int> d_eid_, d_ecnt_sorted_, d_entry_sorted_, d_erank_, d_egroup_len_, d_egroup_base_;   // entry chains
	DevBuf<int> d_erec_all_, d_mrec_all_;               // per-round records of all pairs / contributions
	DevBuf<uint16_t> d_eown_, d_mown_;                  // ... and the chain (rank) each belongs to
	size_t n_eslots_ = 0;
	int max_entry_pairs_ = 0;
	DevBuf<int> d_status_;
	DevBuf<unsigned long long> d_prof_;   // HIBAG_B200_EM_PROF
	PinBuf<double> h_freq_;
	PinBuf<int> h_status_;
	Event ev0_, ev1_;
	Event ev_done_{false, true};
};

/// build_haplomatch hook body (reference src/LibHLA.cpp:1014-1072): records (in-bag index,
/// (i2 << 16) | i1) of the haplotype pairs at minimum distance, i1 outer / i2 inner, as a
/// malloc'd buffer buf[0] = 2*n_records followed by the records. geno = TGenotype[n_samp]
/// (host), ib = ascending in-bag sample indices.
uint32_t *haplomatch_records(const hibag_haplotype *haplo, const size_t *n_haplo, int n_hla,
	int n_snp, const hibag_genotype *geno, int n_samp, const std::vector<int> &ib,
	size_t *out_n);

}  // namespace hb
