// scorer.h -- device-side state for scoring one haplotype list against lists of samples
// (the body of build_set_haplo_geno / build_acc_oob / build_acc_ib) (internal)
#pragma once

#include <cstdint>
#include <mutex>
#include <vector>

#include "common.h"
#include "kernels.h"
#include "tasks.h"

namespace hb {

/// device copy of EXP_LOG_MIN_RARE_FREQ for the current device (uploaded once per device)
const double *device_rare_freq_table();
/// device copy of the floored table T' = max(T, 1e-100) the screening bounds use
const double *device_rare_freq_floor_table();

/// counters shared by the training / prediction drivers
struct ScoreStats
{
	uint64_t pair_evals = 0, popc32 = 0, launches = 0, cell_launches = 0;
	uint64_t pair_evals_nominal = 0;     // what the reference evaluates for the same passes
	uint64_t screen_fallback = 0;        // in-bag positions rescored without screening
	uint64_t h2d_bytes = 0, d2h_bytes = 0;
	double kernel_ms = 0, cell_ms = 0;
	double gather_ms = 0;                // summed CUDA-event durations of cell_gather_kernel alone
	uint64_t gather_launches = 0;
	double gather_ib_ms = 0;             // ... of the in-bag launches only
	uint64_t gather_ib_launches = 0, gather_ib_popc32 = 0;
	void add(const ScoreStats &o)
	{
		pair_evals += o.pair_evals; popc32 += o.popc32; launches += o.launches;
		cell_launches += o.cell_launches; h2d_bytes += o.h2d_bytes; d2h_bytes += o.d2h_bytes;
		pair_evals_nominal += o.pair_evals_nominal; screen_fallback += o.screen_fallback;
		kernel_ms += o.kernel_ms; cell_ms += o.cell_ms;
		gather_ms += o.gather_ms; gather_launches += o.gather_launches;
		gather_ib_ms += o.gather_ib_ms; gather_ib_launches += o.gather_ib_launches;
		gather_ib_popc32 += o.gather_ib_popc32;
	}
};

/// packed genotypes of a cohort on the device, SoA 32-bit words, plus true HLA types
struct GenoSet
{
	int n = 0;
	DevBuf<uint32_t> s1, s2;     // [4][n]
	DevBuf<int> a1, a2, boot;    // [n]
	void ensure(int n_samp)
	{
		n = n_samp;
		s1.ensure((size_t)4 * n_samp); s2.ensure((size_t)4 * n_samp);
		a1.ensure(n_samp); a2.ensure(n_samp); boot.ensure(n_samp);
	}
};

struct GenoView
{
	const uint32_t *s1 = nullptr, *s2 = nullptr;
	int stride = 0;
	const int *a1 = nullptr, *a2 = nullptr;
	const int8_t *cand_col = nullptr;    // optional candidate SNP column, int8 [n]
	int cand_bit = 0;
};

/// One in-flight evaluation: its own stream, pinned staging, device list, cell matrix and
/// result buffers. Not thread-safe; use one slot per host thread.
class EvalSlot
{
public:
	/// spin: sync() busy-waits (lowest latency; the sequential hook path) instead of sleeping on a
	/// blocking event (the trainer's workers, which must free their core)
	explicit EvalSlot(bool high_priority = false, bool spin = false);
	/// pack + cut + async upload of a haplotype list (tags must be filled, see tasks.h)
	void stage_list(const hibag_haplotype *haplo, int n_hap, int n_hla, int n_snp);
	/// score the list staged on another slot (its device blob must stay untouched until this
	/// slot's work has finished); the caller orders the streams
	void borrow_list(const EvalSlot &o) { blob_ = o.blob_; ext_blob_ = o.d_blob_.get(); }
	/// samples per lane of the following enqueue_cells (0: chosen from the pass size)
	void set_samples_per_lane(int r) { force_r_ = r; }
	/// all cells for the positions in pos_list (device int[n_pos], may be null = identity)
	void enqueue_cells(const GenoView &g, const int *pos_list, int n_pos);
	/// out-of-bag accuracy over the last enqueue_cells (src/LibHLA.cpp:1934-1955)
	void enqueue_reduce_oob(const GenoView &g, const int *pos_list, int n_pos);
	/// in-bag P_true/sum ratios over the last enqueue_cells (src/LibHLA.cpp:1957-1979)
	void enqueue_reduce_ib(const GenoView &g, const int *pos_list, int n_pos);
	/// wait for everything enqueued on this slot; folds event timings into stats
	void sync();
	int oob_count() const { return *h_count_.get(); }
	const double *ib_ratios() const { return h_ratio_.get(); }

	cudaStream_t stream() const { return st_.s; }
	const ListBlob &list() const { return blob_; }
	double *cell_matrix() const { return P_.get(); }
	size_t cell_stride() const { return p_stride_; }
	const void *dev_blob() const { return d_blob_.get(); }
	ScoreStats stats;

private:
	Stream st_;
	Event ev0_, evc_;            // slot span begin, end of the pair-scoring kernel
	Event ev1_;                  // slot span end; blocking sync (unless spin) so waiting workers free their core
	bool timing_pending_ = false;
	const unsigned char *ext_blob_ = nullptr;
	int force_r_ = 0;
	PinBuf<unsigned char> h_blob_;
	DevBuf<unsigned char> d_blob_;
	ListBlob blob_;
	DevBuf<double> P_;
	size_t p_stride_ = 0;
	DevBuf<unsigned int> counter_;
	DevBuf<int> d_count_;
	PinBuf<int> h_count_;
	DevBuf<double> d_ratio_;
	PinBuf<double> h_ratio_;
};

/// Streams of the current device on which the batched pair-scoring launches of the training
/// path are enqueued, from however many classifiers are being grown concurrently. Queue 0
/// takes every unscreened launch: those fill the GPU, run one after the other, and the CUDA
/// events around a launch time that launch alone. Screened passes (chains of small launches)
/// are spread over a few queues so that the lanes' passes overlap.
struct ScoreQueue
{
	std::mutex mu;
	Stream st;
	explicit ScoreQueue(bool high_priority) : st(high_priority) {}
	static ScoreQueue &get(int id);      // of the current device
};

/// Scores the candidate haplotype lists of one selection round in ONE launch per pass
/// (out-of-bag, then in-bag for the candidates that need it). Owned by one trainer lane.
class BatchScorer
{
public:
	BatchScorer();
	/// reserve pinned + device staging for n_lists lists of at most max_hap haplotypes
	void begin_round(int n_lists, int max_hap, int n_snp, int n_hla);
	/// pinned staging area of list i (capacity list_blob_capacity(max_hap, ...)); thread-safe
	/// for distinct i
	unsigned char *host_blob(int i) const { return h_blobs_.get() + (size_t)i * cap_; }
	void set_list(int i, const ListBlob &b, const int8_t *cand_col) { blobs_[i] = b; cols_[i] = cand_col; }
	/// async H2D of the lists in `which`
	void upload(const std::vector<int> &which);
	/// out-of-bag correct-allele counts of the lists in `which` -> counts[k] (blocks)
	void score_oob(const GenoView &g, int cand_bit, const std::vector<int> &which,
		const int *pos_list, int n_pos, std::vector<int> &counts);
	/// in-bag P_true/sum ratios of the lists in `which`; ratios(k) valid until the next call
	void score_ib(const GenoView &g, int cand_bit, const std::vector<int> &which,
		const int *pos_list, int n_pos);
	const double *ratios(int k) const { return h_ratio_.get() + (size_t)k * ratio_stride_; }
	/// Exact screening (kernels.h): the sample sets of the classifier being grown -- ascending
	/// sample indices of the out-of-bag and in-bag samples (the same order as the device lists
	/// passed to score_oob / score_ib) and every sample's true type (a1 <= a2). Turns screening
	/// on for the following score_* calls.
	void set_sample_sets(const std::vector<int> &oob, const std::vector<int> &ib,
		const std::vector<int> &a1, const std::vector<int> &a2, int n_hla);
	void disable_screening() { screen_ = false; }
	ScoreStats stats;

private:
	void run_cells(const GenoView &g, int cand_bit, const std::vector<int> &which, int first,
		int count, const int *pos_list, int n_pos, double *P, size_t p_stride, int queue = 0);
	void run_cells_screened(const GenoView &g, int cand_bit, const std::vector<int> &which,
		int first, int count, const int *pos_list, int n_pos, int kind);
	void rescore_uncertified(const GenoView &g, int cand_bit, const std::vector<int> &which);
	// screening state
	bool screen_ = false;
	int queue_id_ = 0;
	double screen_tau_ = 0x1p-70;                // in-bag: a cell is scored when bound >= tau * x_true
	std::vector<int> set_samples_[2];            // host copy of the sample lists (0 oob, 1 in-bag)
	DevBuf<int> ent_off_[2];                     // out_idx * p_stride of the set
	DevBuf<double> U_, xref_;
	double evals_per_list_[2] = { 0, 0 };        // executed by the previous pass of the kind
	DevBuf<int> cnt_, ent_, al_tab_;
	DevBuf<double> U2_;                          // second level of the in-bag screen: class sums
	DevBuf<int> hetk_, cnt2_, ent2_;             // ... heterozygous SNPs used, compacted need lists
	// position classes of a screened pass (kernels.h ScreenArgs::rep): hash sets + counters, the
	// representative of every position, the compacted representatives, out-of-bag results
	DevBuf<int> dd_table_, rep_, rep_list_, pos_res_;
	DevBuf<unsigned int> prefix_;
	DevBuf<unsigned long long> d_evals_;
	PinBuf<unsigned long long> h_evals_;
	DevBuf<int> d_fb_samp_;
	DevBuf<double> P_fb_, d_ratio_fb_;
	PinBuf<double> h_ratio_fb_;
	std::vector<uint64_t> list_pairs_;           // pairs per sample of the lists of the pass
	Stream st_;
	Event ev_up_{false}, ev0_, ev1_, ev_g0_, ev_g1_;
	void add_gather_time(bool in_bag);
	Event ev_done_{false, true};
	size_t cap_ = 0;
	int n_snp_ = 0, n_hla_ = 0, n_cells_ = 0;
	PinBuf<unsigned char> h_blobs_;
	DevBuf<unsigned char> d_blobs_;
	std::vector<ListBlob> blobs_;
	std::vector<const int8_t *> cols_;
	DevBuf<double> P_;
	size_t p_stride_ = 0;
	DevBuf<unsigned int> counters_;
	DevBuf<int> d_counts_;
	PinBuf<int> h_counts_;
	DevBuf<double> d_ratio_;
	PinBuf<double> h_ratio_;
	size_t ratio_stride_ = 0;
};

/// samples per lane for a pass over n_pos samples with n_chunks chunks
int choose_samples_per_lane(int n_pos, int n_chunks, int n_snp, int sm_count);

}  // namespace hb
