// trainer.cu -- attribute-bagging training driver on the GPU scoring path.
//
// Own restatement of the reference's host control flow (CAttrBag_Model::BuildClassifiers
// src/LibHLA.cpp:2268-2305, NewClassifierBootstrap :2220-2245, CVariableSelection::
// InitSelection :1843-1878, _InitHaplotype :1880-1911, Search :1981-2122), re-organised for
// the B200:
//   * the <= mtry candidate SNPs of a selection round are estimated concurrently on a host
//     thread pool (EM is the Amdahl term once scoring is on the GPU) and each candidate's
//     out-of-bag evaluation is submitted on its own CUDA stream as soon as its EM finishes;
//   * the genotype bit planes of the accepted SNPs live on the device; a candidate SNP is
//     patched in by the scoring kernel from the device-resident raw genotype column, so per
//     candidate only the haplotype list (KBs) crosses PCIe;
//   * decisions are applied afterwards in candidate order with the reference's exact rules
//     (including the "loss stays 0 when acc < running max" behaviour, :2031-2034), so the
//     trained model is bit-identical to the reference for any thread count.
// With use_legacy_hooks the same search scores through the ten-hook plugin struct with full
// host buffers per candidate, exactly as the reference host would call it.

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>

#include "em.h"
#include "model.h"

namespace hb {

namespace {

double now_s()
{
	using namespace std::chrono;
	return duration<double>(steady_clock::now().time_since_epoch()).count();
}

/// persistent worker threads; run(n, fn) executes fn(item, worker) for item in [0,n) with
/// dynamic scheduling. Results never depend on which worker ran an item.
class ThreadPool
{
public:
	ThreadPool(int n_threads, int device) : stop_(false), job_id_(0), n_items_(0), pending_(0)
	{
		if (n_threads < 1) n_threads = 1;
		for (int w = 1; w < n_threads; w++)
			threads_.emplace_back([this, w, device]() { cudaSetDevice(device); loop(w); });
		n_workers_ = n_threads;
	}
	~ThreadPool()
	{
		{
			std::lock_guard<std::mutex> lk(mu_);
			stop_ = true;
		}
		cv_.notify_all();
		for (auto &t : threads_) t.join();
	}
	int size() const { return n_workers_; }

	void run(int n, const std::function<void(int, int)> &fn)
	{
		if (n <= 0) return;
		if (n_workers_ == 1 || n == 1)
		{
			for (int i = 0; i < n; i++) fn(i, 0);
			return;
		}
		{
			std::lock_guard<std::mutex> lk(mu_);
			fn_ = &fn;
			n_items_ = n;
			next_.store(0);
			pending_ = n_workers_ - 1;
			error_.clear();
			job_id_++;
		}
		cv_.notify_all();
		work(0);
		std::unique_lock<std::mutex> lk(mu_);
		done_cv_.wait(lk, [this]() { return pending_ == 0; });
		fn_ = nullptr;
		if (!error_.empty()) throw std::runtime_error(error_);
	}

private:
	void work(int w)
	{
		try
		{
			for (;;)
			{
				const int i = next_.fetch_add(1);
				if (i >= n_items_) break;
				(*fn_)(i, w);
			}
		} catch (std::exception &e)
		{
			std::lock_guard<std::mutex> lk(mu_);
			if (error_.empty()) error_ = e.what();
			next_.store(n_items_);
		}
	}
	void loop(int w)
	{
		uint64_t seen = 0;
		for (;;)
		{
			{
				std::unique_lock<std::mutex> lk(mu_);
				cv_.wait(lk, [&]() { return stop_ || job_id_ != seen; });
				if (stop_) return;
				seen = job_id_;
			}
			work(w);
			{
				std::lock_guard<std::mutex> lk(mu_);
				pending_--;
			}
			done_cv_.notify_all();
		}
	}

	std::vector<std::thread> threads_;
	int n_workers_;
	std::mutex mu_;
	std::condition_variable cv_, done_cv_;
	bool stop_;
	uint64_t job_id_;
	const std::function<void(int, int)> *fn_ = nullptr;
	int n_items_;
	std::atomic<int> next_{0};
	int pending_;
	std::string error_;
};

struct PfCtx { ThreadPool *pool; };

void pool_parallel_for(void *ctx, int n, void (*fn)(void *, int, int), void *arg)
{
	ThreadPool *pool = ((PfCtx *)ctx)->pool;
	pool->run(n, [&](int b, int) { fn(arg, b, b + 1); });
}

/// per-candidate state of one selection round
struct Candidate
{
	int snp = -1;
	bool valid = false;
	int acc = 0;
	double loss = 0;
	HapList list;
	ListBlob blob;
};

class Trainer : public TrainSession
{
public:
	Trainer(hibag_b200_model &m, const hibag_b200_train_opts &o, int n_pool_threads, int n_lanes);
	~Trainer() override;
	/// grow classifiers into built_ / ts_ / trace_: next() hands out global classifier indices
	/// (< 0: none left); the lanes of a model share one dispenser, so a lane that finishes a
	/// classifier early takes the next one
	void run(const std::function<int()> &next);
	/// can this session (threads, scoring mode) serve a call with these options?
	void configure(const hibag_b200_train_opts &o)
	{
		o_ = o;
		if (o_.mtry <= 0) o_.mtry = 1;
	}
	// results of the last run()
	std::vector<std::pair<int, Classifier> > built_;
	hibag_b200_train_stats ts_;
	std::vector<int64_t> trace_;

private:
	void grow(Classifier &c);
	void set_snp_bit(int bit, int snp);      // CGenotypeList::AddSNP on the host copy
	void upload_base_geno();
	void make_aos();
	double ib_loss(const double *ratio) const;

	hibag_b200_model &m_;
	hibag_b200_train_opts o_;
	const DeviceInfo *dev_;
	std::unique_ptr<ThreadPool> pool_;
	RRng rng_;

	// cohort
	int n_samp_, n_snp_, n_hla_;
	std::vector<int> a1_, a2_;               // true types, a1 <= a2 (:1863-1868)
	std::vector<int> boot_, inbag_, oob_;
	std::vector<HostGeno> geno_;             // accepted SNPs of the classifier being grown
	std::vector<hibag_genotype> aos_;        // legacy-hook mode: TGenotype[] as the reference holds it

	// device-resident state
	DevBuf<int8_t> d_geno_t_;                // raw genotypes, SNP-major
	DevBuf<int> d_a1_, d_a2_, d_oob_, d_ib_, d_boot_;
	std::unique_ptr<RoundEM> rem_;                    // device-side pair matching + EM
	DevBuf<uint32_t> d_s1_, d_s2_;           // base bit planes [4][n_samp]
	PinBuf<uint32_t> h_planes_;
	Stream main_st_;
	std::unique_ptr<BatchScorer> scorer_;             // all candidates of a round per launch
	std::vector<EmScratch> scratch_;                  // one per worker
	std::vector<double> em_seconds_, wait_seconds_;   // per worker

	int64_t cl_global_index_ = 0;
	hibag_gpu_ext_proc *procs_ = nullptr;    // legacy-hook mode
	ScoreStats stats_;
};

Trainer::Trainer(hibag_b200_model &m, const hibag_b200_train_opts &o, int n_pool_threads, int n_lanes)
	: m_(m), o_(o)
{
	dev_ = &current_device();
	memset(&ts_, 0, sizeof(ts_));
	n_samp_ = m.n_samp; n_snp_ = m.n_snp; n_hla_ = m.n_hla;
	if (n_samp_ <= 0 || m.geno_t.empty())
		throw std::runtime_error("train: no training data (call hibag_b200_model_set_training)");
	if (o_.mtry <= 0) o_.mtry = 1;
	pool_.reset(new ThreadPool(n_pool_threads, dev_->device));
	scratch_.resize(pool_->size());
	em_seconds_.assign(pool_->size(), 0.0);
	wait_seconds_.assign(pool_->size(), 0.0);

	a1_.resize(n_samp_); a2_.resize(n_samp_);
	for (int i = 0; i < n_samp_; i++)
	{
		if (m.h1[i] < 0 || m.h1[i] >= n_hla_ || m.h2[i] < 0 || m.h2[i] >= n_hla_)
			throw std::runtime_error("train: HLA allele index out of range");
		a1_[i] = std::min(m.h1[i], m.h2[i]);
		a2_[i] = std::max(m.h1[i], m.h2[i]);
	}
	geno_.resize(n_samp_);

	if (o_.use_legacy_hooks)
	{
		procs_ = plugin_procs();
		procs_->build_init(n_hla_, n_samp_);
	} else {
		d_geno_t_.ensure(m.geno_t.size());
		HB_CUDA(cudaMemcpy(d_geno_t_.get(), m.geno_t.data(), m.geno_t.size(), cudaMemcpyHostToDevice));
		stats_.h2d_bytes += m.geno_t.size();
		d_a1_.ensure(n_samp_); d_a2_.ensure(n_samp_); d_oob_.ensure(n_samp_); d_ib_.ensure(n_samp_);
		HB_CUDA(cudaMemcpy(d_a1_.get(), a1_.data(), sizeof(int) * n_samp_, cudaMemcpyHostToDevice));
		HB_CUDA(cudaMemcpy(d_a2_.get(), a2_.data(), sizeof(int) * n_samp_, cudaMemcpyHostToDevice));
		d_s1_.ensure((size_t)4 * n_samp_); d_s2_.ensure((size_t)4 * n_samp_);
		h_planes_.ensure((size_t)8 * n_samp_);
		device_rare_freq_table();
		scorer_.reset(new BatchScorer());
		d_boot_.ensure(n_samp_);
		rem_.reset(new RoundEM());
		rem_->set_lanes(n_lanes);
	}
}

Trainer::~Trainer()
{
	if (procs_)
	{
		try { procs_->build_done(); } catch (...) {}
	}
}

void Trainer::set_snp_bit(int bit, int snp)
{
	const int8_t *col = m_.geno_t.data() + (size_t)snp * n_samp_;
	const int w = bit >> 6;
	const uint64_t b = (uint64_t)1 << (bit & 63);
	for (int i = 0; i < n_samp_; i++)
	{
		HostGeno &g = geno_[i];
		switch (col[i])       // src/LibHLA.cpp:609-622
		{
		case 0: g.s1[w] &= ~b; g.s2[w] &= ~b; break;
		case 1: g.s1[w] |= b; g.s2[w] &= ~b; break;
		case 2: g.s1[w] |= b; g.s2[w] |= b; break;
		default: g.s1[w] &= ~b; g.s2[w] |= b;
		}
	}
}

void Trainer::upload_base_geno()
{
	uint32_t *h = h_planes_.get();
	const size_t n = n_samp_;
	for (int i = 0; i < n_samp_; i++)
	{
		const HostGeno &g = geno_[i];
		for (int w = 0; w < 4; w++)
		{
			h[(size_t)w * n + i] = (uint32_t)(g.s1[w >> 1] >> ((w & 1) * 32));
			h[(4 + (size_t)w) * n + i] = (uint32_t)(g.s2[w >> 1] >> ((w & 1) * 32));
		}
	}
	HB_CUDA(cudaMemcpyAsync(d_s1_.get(), h, sizeof(uint32_t) * 4 * n, cudaMemcpyHostToDevice, main_st_.s));
	HB_CUDA(cudaMemcpyAsync(d_s2_.get(), h + 4 * n, sizeof(uint32_t) * 4 * n, cudaMemcpyHostToDevice, main_st_.s));
	stream_sync_blocking(main_st_.s);
	stats_.h2d_bytes += sizeof(uint32_t) * 8 * n;
}

void Trainer::make_aos()
{
	aos_.resize(n_samp_);
	for (int i = 0; i < n_samp_; i++)
	{
		hibag_genotype &g = aos_[i];
		g.snp1[0] = (int64_t)geno_[i].s1[0]; g.snp1[1] = (int64_t)geno_[i].s1[1];
		g.snp2[0] = (int64_t)geno_[i].s2[0]; g.snp2[1] = (int64_t)geno_[i].s2[1];
		g.bootstrap_count = boot_[i];
		g.allele1 = a1_[i]; g.allele2 = a2_[i];
		g.aux_temp = 0;
	}
}

double Trainer::ib_loss(const double *ratio) const
{
	// -2 * sum_{in-bag, ascending sample order} count * log(P_true / sum P) (:1966-1977)
	double loglik = 0;
	for (size_t k = 0; k < inbag_.size(); k++)
		loglik += boot_[inbag_[k]] * std::log(ratio[k]);
	return loglik * -2;
}

void Trainer::run(const std::function<int()> &next)
{
	const double t0 = now_s();
	// the session outlives a call: start this call's counters from zero
	stats_ = ScoreStats();
	memset(&ts_, 0, sizeof(ts_));
	built_.clear();
	trace_.clear();
	if (scorer_) scorer_->stats = ScoreStats();
	if (rem_)
	{
		rem_->kernel_ms = 0; rem_->launches = 0; rem_->h2d_bytes = 0; rem_->d2h_bytes = 0;
		rem_->sum_iterations = rem_->sum_chain_adds = rem_->sum_pair_updates = 0;
	}
	const ScoreStats plugin_before = procs_ ? plugin_build_stats() : ScoreStats();
	std::fill(em_seconds_.begin(), em_seconds_.end(), 0.0);
	std::fill(wait_seconds_.begin(), wait_seconds_.end(), 0.0);
	if (!o_.per_classifier_seed)
	{
		// reference: the user seeds R's generator once and every BuildClassifiers call continues the
		// stream (src/LibHLA.cpp:120-126). Here: seeded when the model sees this seed value for the
		// first time; a later call with the same seed continues where the last one stopped.
		if (!m_.rng_seeded || m_.rng_seed != (int64_t)o_.seed)
		{
			m_.rng.set_seed((uint32_t)o_.seed);
			m_.rng_seeded = true;
			m_.rng_seed = (int64_t)o_.seed;
		}
		rng_ = m_.rng;
	}
	for (;;)
	{
		const int global_k = next();
		if (global_k < 0) break;
		if (o_.per_classifier_seed) rng_.set_seed((uint32_t)(o_.seed + global_k));
		// bootstrap; redraw the whole sample when nobody is left out of the bag (:2229-2240)
		boot_.assign(n_samp_, 0);
		int n_unique;
		do {
			std::fill(boot_.begin(), boot_.end(), 0);
			n_unique = 0;
			for (int i = 0; i < n_samp_; i++)
			{
				const int k = rng_.random_num(n_samp_);
				if (boot_[k] == 0) n_unique++;
				boot_[k]++;
			}
		} while (n_unique >= n_samp_);

		cl_global_index_ = global_k;
		built_.emplace_back(global_k, Classifier());
		Classifier &cl = built_.back().second;
		cl.samp_num = boot_;
		if (o_.verbose)
		{
			fprintf(stderr, "=== building individual classifier %d, out-of-bag (%d/%.1f%%) ===\n",
				global_k + 1, n_samp_ - n_unique, 100.0 * (n_samp_ - n_unique) / n_samp_);
		}
		const double t_grow = now_s();
		grow(cl);
		if (getenv("HIBAG_B200_TIMING"))
			fprintf(stderr, "classifier %d: start %.3f s, %.3f s, %d SNPs\n", global_k, t_grow - t0, now_s() - t_grow,
				(int)cl.snpidx.size());
		if (o_.verbose)
		{
			fprintf(stderr, "[%d] oob acc: %0.2f%%, # of SNPs: %d, # of haplo: %d\n", global_k + 1,
				cl.oob_acc * 100, (int)cl.snpidx.size(), (int)cl.haplo.h.size());
		}
	}
	if (!o_.per_classifier_seed) m_.rng = rng_;
	// fold the counters
	if (scorer_) stats_.add(scorer_->stats);
	if (rem_)
	{
		stats_.launches += rem_->launches; stats_.kernel_ms += rem_->kernel_ms;
		stats_.h2d_bytes += rem_->h2d_bytes; stats_.d2h_bytes += rem_->d2h_bytes;
		ts_.em_kernel_ms += rem_->kernel_ms;
		ts_.em_iterations += rem_->sum_iterations;
		ts_.em_chain_adds += rem_->sum_chain_adds;
		ts_.em_pair_updates += rem_->sum_pair_updates;
	}
	if (procs_)
	{
		ScoreStats now = plugin_build_stats();
		now.pair_evals -= plugin_before.pair_evals; now.popc32 -= plugin_before.popc32;
		now.launches -= plugin_before.launches; now.cell_launches -= plugin_before.cell_launches;
		now.h2d_bytes -= plugin_before.h2d_bytes; now.d2h_bytes -= plugin_before.d2h_bytes;
		now.kernel_ms -= plugin_before.kernel_ms; now.cell_ms -= plugin_before.cell_ms;
		now.pair_evals_nominal = now.pair_evals; now.screen_fallback = 0;
		stats_.add(now);
	}
	hibag_b200_train_stats &ts = ts_;
	ts.seconds_total += now_s() - t0;
	for (double v : em_seconds_) ts.seconds_em += v;
	for (double v : wait_seconds_) ts.seconds_gpu_wait += v;
	ts.gpu_kernel_ms += stats_.kernel_ms;
	ts.pair_evals += stats_.pair_evals;
	ts.pair_evals_nominal += stats_.pair_evals_nominal;
	ts.n_screen_fallback += stats_.screen_fallback;
	ts.gather_kernel_ms += stats_.gather_ms;
	ts.gather_kernel_launches += stats_.gather_launches;
	ts.gather_ib_kernel_ms += stats_.gather_ib_ms;
	ts.gather_ib_launches += stats_.gather_ib_launches;
	ts.gather_ib_popc32 += stats_.gather_ib_popc32;
	ts.popc32_issued += stats_.popc32;
	ts.kernel_launches += stats_.launches;
	ts.cell_kernel_ms += stats_.cell_ms;
	ts.cell_kernel_launches += stats_.cell_launches;
	ts.h2d_bytes += stats_.h2d_bytes;
	ts.d2h_bytes += stats_.d2h_bytes;
}

void Trainer::grow(Classifier &cl)
{
	static const double FRACTION_HAPLO = 1.0 / 10;             // src/LibHLA.cpp:108
	static const double MIN_RARE_FREQ = 1e-5;
	static const double STOP_RELTOL_LOGLIK_ADDSNP = 0.001;     // :114
	static const double PRUNE_RELTOL_LOGLIK = 0.1;             // :116
	hibag_b200_train_stats &ts = ts_;

	// ---- InitSelection (:1843-1878) ------------------------------------------------------
	inbag_.clear(); oob_.clear();
	for (int i = 0; i < n_samp_; i++)
	{
		(boot_[i] > 0 ? inbag_ : oob_).push_back(i);
		geno_[i].s1[0] = geno_[i].s1[1] = 0;
		geno_[i].s2[0] = geno_[i].s2[1] = ~(uint64_t)0;
	}
	if (procs_)
	{
		procs_->build_set_bootstrap(boot_.data());
		make_aos();
	} else {
		HB_CUDA(cudaMemcpyAsync(d_oob_.get(), oob_.data(), sizeof(int) * oob_.size(),
			cudaMemcpyHostToDevice, main_st_.s));
		HB_CUDA(cudaMemcpyAsync(d_ib_.get(), inbag_.data(), sizeof(int) * inbag_.size(),
			cudaMemcpyHostToDevice, main_st_.s));
		HB_CUDA(cudaMemcpyAsync(d_boot_.get(), boot_.data(), sizeof(int) * (size_t)n_samp_,
			cudaMemcpyHostToDevice, main_st_.s));
		stats_.h2d_bytes += 2 * sizeof(int) * (size_t)n_samp_;
		upload_base_geno();
		if (o_.no_screening) scorer_->disable_screening();
		else scorer_->set_sample_sets(oob_, inbag_, a1_, a2_, n_hla_);
	}

	// ---- _InitHaplotype (:1880-1911): one SNP-less haplotype per allele present in the bag ---
	HapList cur;
	{
		std::vector<int> cnt(n_hla_, 0);
		int sum_cnt = 0;
		for (int s : inbag_)
		{
			cnt[a1_[s]] += boot_[s];
			cnt[a2_[s]] += boot_[s];
			sum_cnt += boot_[s];
		}
		cur.n_snp = 0;
		cur.len.assign(n_hla_, 0);
		const double scale = 0.5 / sum_cnt;
		for (int a = 0; a < n_hla_; a++)
			if (cnt[a] > 0)
			{
				cur.len[a] = 1;
				hibag_haplotype h;
				memset(&h, 0, sizeof(h));
				h.freq = cnt[a] * scale;
				cur.h.push_back(h);
			}
		cur.set_tags();
	}
	cl.snpidx.clear();

	const double rare_prob = std::max(FRACTION_HAPLO / (2 * n_samp_), MIN_RARE_FREQ);   // :1987
	const int n_oob = (int)oob_.size();
	int global_max_acc = 0;
	double global_min_loss = 1e+30;

	const int global_k = (int)cl_global_index_;
	int64_t cum_pairs = 0, cum_em = 0;
	auto list_pairs = [](const HapList &l) {
		int64_t tot = 0, rest = 0;
		for (int a = (int)l.len.size() - 1; a >= 0; a--)
		{
			const int64_t n = l.len[a];
			tot += n * (n + 1) / 2 + n * rest;
			rest += n;
		}
		return tot;
	};

	SnpPool pool;
	pool.init(n_snp_);
	RoundPairs rp;
	std::vector<Candidate> cand;
	PfCtx pf = { pool_.get() };

	GenoView view;
	view.s1 = d_s1_.get(); view.s2 = d_s2_.get(); view.stride = n_samp_;
	view.a1 = d_a1_.get(); view.a2 = d_a2_.get();

	const bool dev_em_wanted = (o_.em_on_device != 0) && !procs_;
	bool dev_em = dev_em_wanted;
	bool pairs_dirty = true;       // the pair lists depend on `cur` and the accepted SNPs only
	bool host_rp_valid = false;
	std::vector<int> cand_snps;

	while (pool.total() > 0 && (int)cl.snpidx.size() < HIBAG_B200_MAX_SNP)
	{
		const double t_prep = now_s();
		if (pairs_dirty)
		{
			dev_em = dev_em_wanted && RoundEM::supports((int)cur.h.size(), (int)inbag_.size());
			if (dev_em)
			{
				rem_->prepare(cur, d_s1_.get(), d_s2_.get(), n_samp_, d_a1_.get(), d_a2_.get(),
					d_ib_.get(), (int)inbag_.size(), d_boot_.get(), main_st_.s);
				host_rp_valid = false;
				if (rem_->has_empty_entry())
				{
					// degenerate round (an in-bag sample without pairs): host arithmetic
					rem_->fetch_pairs(rp, inbag_, boot_, main_st_.s);
					host_rp_valid = true;
					dev_em = false;
				}
			} else {
				prepare_round(cur, geno_, a1_, a2_, boot_, inbag_, rp, pool_parallel_for, &pf);
			}
			pairs_dirty = false;
		}
		ts.seconds_prepare += now_s() - t_prep;

		pool.random_select(o_.mtry, rng_);
		const int m = pool.n_selected();
		if ((int)cand.size() < m) cand.resize(m);
		const int bit = cur.n_snp;
		if (!procs_) scorer_->begin_round(m, 2 * (int)cur.h.size(), bit + 1, n_hla_);

		// ---- phase 1: EM of the candidates. Device: one launch, one CTA per candidate; the host
		// then only prunes rare haplotypes and packs the lists (in parallel on the pool), and
		// re-estimates on the host the rare candidate whose stopping test the device could not
		// decide (em.h). Host: the candidates' EM in parallel on the pool.
		const double t_p1 = now_s();
		if (dev_em)
		{
			cand_snps.resize(m);
			for (int i = 0; i < m; i++) cand_snps[i] = pool.at(i);
			const double t_w = now_s();
			rem_->run_em(cand_snps.data(), m, d_geno_t_.get(), n_samp_, main_st_.s);
			wait_seconds_[0] += now_s() - t_w;
			if (const char *e = getenv("HIBAG_B200_EM_FORCE_FALLBACK"))
			{
				// test hook: every k-th candidate takes the host re-estimation path
				const int every = std::max(1, atoi(e));
				for (int i = 0; i < m; i += every) rem_->force_ambiguous(i);
			}
			if (getenv("HIBAG_B200_EM_DEBUG"))
			{
				int it_max = 0, it_sum = 0, nv = 0;
				for (int i = 0; i < m; i++)
					if (rem_->status(i) != EM_INVALID)
					{
						it_max = std::max(it_max, rem_->iterations(i)); it_sum += rem_->iterations(i); nv++;
					}
				fprintf(stderr, "em round: n_snp %d n_cur %d pairs %zu slots %zu maxchain %d m %d valid %d iters max %d mean %.1f wall %.3f ms\n",
					bit, (int)cur.h.size(), rem_->total_pairs(), rem_->ell_slots(), rem_->max_chain(), m, nv, it_max, nv ? (double)it_sum / nv : 0.0,
					(now_s() - t_w) * 1e3);
			}
			for (int i = 0; i < m; i++)
				if (rem_->status(i) == EM_AMBIGUOUS && !host_rp_valid)
				{
					rem_->fetch_pairs(rp, inbag_, boot_, main_st_.s);
					host_rp_valid = true;
				}
		}
		const auto em_item = [&](int i, int w) {
			Candidate &cd = cand[i];
			cd.snp = pool.at(i);
			cd.acc = 0; cd.loss = 0;
			const double t_em = now_s();
			if (dev_em && rem_->status(i) != EM_AMBIGUOUS)
			{
				cd.valid = rem_->status(i) == EM_OK;
				if (cd.valid) finish_candidate(cur, rem_->freq(i), rare_prob, cd.list);
			} else {
				cd.valid = estimate_candidate(cur, rp, m_.geno_t.data() + (size_t)cd.snp * n_samp_,
					n_samp_, rare_prob, scratch_[w], cd.list);
			}
			em_seconds_[w] += now_s() - t_em;
			if (!cd.valid || procs_) return;
			cd.blob = build_list_blob(cd.list.h.data(), (int)cd.list.h.size(), n_hla_, cd.list.n_snp,
				scorer_->host_blob(i), 256);
			scorer_->set_list(i, cd.blob, d_geno_t_.get() + (size_t)cd.snp * n_samp_);
		};
		// ---- legacy hooks: sequential, exactly the reference's call sequence (:2022-2040) -----------
		int hook_running = global_max_acc;
		double hook_seconds = 0;
		const auto hook_item = [&](int i) {
			Candidate &cd = cand[i];
			if (!cd.valid) return;
			const double t_w = now_s();
			// AddSNP on the TGenotype array the hook receives (:2027, :860-874)
			const int8_t *col = m_.geno_t.data() + (size_t)cd.snp * n_samp_;
			const int w = bit >> 6;
			const uint64_t b = (uint64_t)1 << (bit & 63);
			for (int s = 0; s < n_samp_; s++)
			{
				uint64_t s1 = (uint64_t)aos_[s].snp1[w], s2 = (uint64_t)aos_[s].snp2[w];
				switch (col[s])
				{
				case 0: s1 &= ~b; s2 &= ~b; break;
				case 1: s1 |= b; s2 &= ~b; break;
				case 2: s1 |= b; s2 |= b; break;
				default: s1 &= ~b; s2 |= b;
				}
				aos_[s].snp1[w] = (int64_t)s1; aos_[s].snp2[w] = (int64_t)s2;
			}
			procs_->build_set_haplo_geno(cd.list.h.data(), (int)cd.list.h.size(), aos_.data(),
				cd.list.n_snp);
			cd.acc = procs_->build_acc_oob();
			ts.n_oob_evals++;
			if (cd.acc >= hook_running)
			{
				cd.loss = procs_->build_acc_ib();
				ts.n_ib_evals++;
			}
			if (cd.acc > hook_running) hook_running = cd.acc;
			hook_seconds += now_s() - t_w;
		};
		bool hooks_done = false;
		if (procs_ && pool_->size() > 1)
		{
			// the host EM of the later candidates runs on the pool WHILE the hooks score the earlier
			// ones: item 0 is the consumer, it calls the hooks in candidate order as the candidates'
			// lists become ready (the hook sequence itself stays strictly sequential)
			std::unique_ptr<std::atomic<int>[]> ready(new std::atomic<int>[m]);
			for (int i = 0; i < m; i++) ready[i].store(0);
			std::atomic<int> failed{0};
			pool_->run(m + 1, [&](int item, int w) {
				if (item == 0)
				{
					for (int i = 0; i < m; i++)
					{
						while (!ready[i].load(std::memory_order_acquire))
						{
							if (failed.load()) return;
							std::this_thread::yield();
						}
						if (failed.load()) return;
						hook_item(i);
					}
					return;
				}
				try { em_item(item - 1, w); }
				catch (...) { failed.store(1); ready[item - 1].store(1, std::memory_order_release); throw; }
				ready[item - 1].store(1, std::memory_order_release);
			});
			hooks_done = true;
		} else
			pool_->run(m, em_item);
		if (dev_em)
			for (int i = 0; i < m; i++) if (rem_->status(i) == EM_AMBIGUOUS) ts.n_em_host_fallback++;
		for (int i = 0; i < m; i++) if (cand[i].valid) { ts.n_em++; }
		ts.seconds_phase_oob += now_s() - t_p1;
		const double t_p2 = now_s();

		if (!procs_)
		{
			// ---- phase 2: ONE launch scores every valid candidate on the out-of-bag samples;
			// then, with all accuracies known, ONE launch scores exactly the candidates whose
			// in-bag loss the reference computes (acc >= running maximum in candidate order, :2033)
			std::vector<int> which, counts;
			for (int i = 0; i < m; i++) if (cand[i].valid) which.push_back(i);
			double t_w = now_s();
			scorer_->upload(which);
			scorer_->score_oob(view, bit, which, d_oob_.get(), n_oob, counts);
			wait_seconds_[0] += now_s() - t_w;
			std::vector<int> need;
			int running = global_max_acc;
			for (size_t k = 0; k < which.size(); k++)
			{
				Candidate &cd = cand[which[k]];
				cd.acc = counts[k];
				ts.n_oob_evals++;
				if (cd.acc >= running) { need.push_back(which[k]); ts.n_ib_evals++; }
				if (cd.acc > running) running = cd.acc;
			}
			t_w = now_s();
			scorer_->score_ib(view, bit, need, d_ib_.get(), (int)inbag_.size());
			wait_seconds_[0] += now_s() - t_w;
			pool_->run((int)need.size(), [&](int k, int) {
				cand[need[k]].loss = ib_loss(scorer_->ratios(k));
			});
		} else if (!hooks_done)
			for (int i = 0; i < m; i++) hook_item(i);
		if (procs_) wait_seconds_[0] += hook_seconds;

		ts.seconds_phase_ib += now_s() - t_p2;
		// workload accounting (independent of scheduling): what the reference would evaluate
		{
			int running = global_max_acc;
			for (int i = 0; i < m; i++)
			{
				if (!cand[i].valid) continue;
				const int64_t lp = list_pairs(cand[i].list);
				cum_em++;
				cum_pairs += lp * n_oob;
				if (cand[i].acc >= running) cum_pairs += lp * (int64_t)inbag_.size();
				if (cand[i].acc > running) running = cand[i].acc;
			}
		}

		// ---- phase 3: the reference's decisions in candidate order (:2041-2067) -------------
		int max_acc = global_max_acc;
		double min_loss = global_min_loss;
		int min_i = -1;
		for (int i = 0; i < m; i++)
		{
			const Candidate &cd = cand[i];
			if (!cd.valid) continue;
			if (cd.acc > max_acc)
			{
				min_i = i; min_loss = cd.loss; max_acc = cd.acc;
			} else if (cd.acc == max_acc)
			{
				if (cd.loss < min_loss) { min_i = i; min_loss = cd.loss; }
			}
			if (o_.prune)
			{
				if (cd.acc < global_max_acc)
					pool.at(i) = -1;
				else if (cd.acc == global_max_acc)
				{
					if ((cd.loss > global_min_loss * (1 + PRUNE_RELTOL_LOGLIK)) && (min_i != i))
						pool.at(i) = -1;
				}
			}
		}

		bool accept = false;                                         // :2072-2085
		if (max_acc > global_max_acc)
			accept = true;
		else if (max_acc == global_max_acc && min_i >= 0)
			accept = (min_loss >= STOP_RELTOL_LOGLIK_ADDSNP) &&
				(min_loss < global_min_loss * (1 - STOP_RELTOL_LOGLIK_ADDSNP));

		if (accept)
		{
			global_max_acc = max_acc;
			global_min_loss = min_loss;
			const int snp = cand[min_i].snp;
			std::swap(cur, cand[min_i].list);
			pairs_dirty = true;
			cl.snpidx.push_back(snp);
			set_snp_bit(bit, snp);
			if (procs_)
			{
				const int w = bit >> 6;
				for (int s = 0; s < n_samp_; s++)
				{
					aos_[s].snp1[w] = (int64_t)geno_[s].s1[w];
					aos_[s].snp2[w] = (int64_t)geno_[s].s2[w];
				}
			} else {
				upload_base_geno();
			}
			if (o_.prune)
			{
				pool.at(min_i) = -1;
				pool.remove_flagged();
			} else {
				pool.remove(min_i);
			}
			{
				const int64_t row[4] = { global_k, (int64_t)cl.snpidx.size(), cum_pairs, cum_em };
				trace_.insert(trace_.end(), row, row + 4);
			}
			if (o_.verbose > 1)
				fprintf(stderr, "    %2d, SNP: %d, loss: %g, oob acc: %0.2f%%, # of haplo: %d\n",
					(int)cl.snpidx.size(), snp + 1, global_min_loss,
					double(global_max_acc) / n_oob * 50, (int)cur.h.size());
		} else {
			pool.remove_selection();
			if (procs_)
			{
				// SetMissing on the hook's genotype array (:2117)
				const int w = bit >> 6;
				const uint64_t b = (uint64_t)1 << (bit & 63);
				for (int s = 0; s < n_samp_; s++)
				{
					aos_[s].snp1[w] = (int64_t)((uint64_t)aos_[s].snp1[w] & ~b);
					aos_[s].snp2[w] = (int64_t)((uint64_t)aos_[s].snp2[w] | b);
				}
			}
		}
	}

	{
		const int64_t row[4] = { global_k, -1, cum_pairs, cum_em };
		trace_.insert(trace_.end(), row, row + 4);
	}
	cl.haplo = cur;
	cl.oob_acc = 0.5 * global_max_acc / n_oob;                       // :2121
}

}  // namespace

/// the lanes of one model: each grows its own classifiers (own RNG stream, host pool, device
/// state); all share the device's scoring stream
struct LaneGroup : public TrainSession
{
	std::vector<std::unique_ptr<Trainer> > lanes;
	bool legacy = false, dev_em = false;
	int req_threads = 0, req_concurrent = 0, mtry = 0;
};

void train_model(hibag_b200_model &m, const hibag_b200_train_opts &opts)
{
	const DeviceInfo &di = current_device();      // fails here, loudly, when no CUDA device is usable
	const int mtry = opts.mtry <= 0 ? 1 : opts.mtry;
	// classifiers in flight on this GPU: independent RNG streams are required for that
	int n_lanes = 1;
	if (opts.per_classifier_seed && !opts.use_legacy_hooks && opts.n_concurrent > 1)
		n_lanes = std::min(opts.n_concurrent, std::max(opts.nclassifier, 1));
	else if (opts.n_concurrent > 1 && !opts.use_legacy_hooks)
	{
		static bool warned = false;
		if (!warned)
		{
			warned = true;
			fprintf(stderr, "hibag_b200: n_concurrent = %d ignored: with per_classifier_seed = 0 the classifiers "
				"consume ONE random stream in order (as R does) and are grown one at a time\n", opts.n_concurrent);
		}
	}
	LaneGroup *g = dynamic_cast<LaneGroup *>(m.tsession.get());
	if (!g || g->legacy != (opts.use_legacy_hooks != 0) || g->req_threads != opts.n_threads ||
		g->req_concurrent != opts.n_concurrent || g->mtry != mtry || (int)g->lanes.size() < n_lanes ||
		g->dev_em != (opts.em_on_device != 0))
		// (no_screening is read per classifier: Trainer::configure)
	{
		m.tsession.reset();
		g = new LaneGroup();
		m.tsession.reset(g);
		g->legacy = opts.use_legacy_hooks != 0;
		g->req_threads = opts.n_threads; g->req_concurrent = opts.n_concurrent; g->mtry = mtry;
		g->dev_em = opts.em_on_device != 0;
		// host threads: all cores by default, split over the lanes; one worker per candidate of
		// a round when that is a modest over-subscription (23 equal EM jobs on 16 cores finish in
		// ~1.4 job-times instead of 2)
		int nt = opts.n_threads;
		if (nt <= 0)
		{
			nt = (int)std::thread::hardware_concurrency();
			if (nt < 1) nt = 1;
			if (n_lanes == 1)
			{
				if (mtry > nt && mtry <= 2 * nt) nt = mtry;
				if (nt > std::max(mtry, 4)) nt = std::max(mtry, 4);
			} else {
				nt = std::min(mtry, std::max(2, (nt * 3 / 2 + n_lanes - 1) / n_lanes));
			}
		} else if (n_lanes > 1)
		{
			nt = std::max(1, (nt + n_lanes - 1) / n_lanes);
		}
		for (int l = 0; l < n_lanes; l++)
			g->lanes.emplace_back(new Trainer(m, opts, nt, n_lanes));
	}
	(void)di;

	const int stride = opts.index_stride > 0 ? opts.index_stride : 1;
	std::atomic<int> next_c{0};
	const std::function<int()> dispenser = [&]() {
		const int c = next_c.fetch_add(1);
		return (c < opts.nclassifier) ? opts.first_index + c * stride : -1;
	};

	const double t0 = now_s();
	std::vector<std::string> errors(n_lanes);
	auto run_lane = [&](int l) {
		try
		{
			cudaSetDevice(di.device);
			g->lanes[l]->configure(opts);
			g->lanes[l]->run(dispenser);
		} catch (std::exception &e)
		{
			errors[l] = e.what();
			if (errors[l].empty()) errors[l] = "unknown error";
		}
	};
	if (n_lanes == 1)
		run_lane(0);
	else {
		std::vector<std::thread> th;
		for (int l = 1; l < n_lanes; l++) th.emplace_back(run_lane, l);
		run_lane(0);
		for (auto &t : th) t.join();
	}
	for (int l = 0; l < n_lanes; l++)
		if (!errors[l].empty())
		{
			const std::string msg = errors[l];
			m.tsession.reset();
			throw std::runtime_error(msg);
		}

	// merge in global classifier order: the model a single lane would have built
	std::vector<std::pair<int, int> > order;       // (global index, lane)
	std::vector<size_t> next(n_lanes, 0);
	for (int l = 0; l < n_lanes; l++)
		for (size_t k = 0; k < g->lanes[l]->built_.size(); k++)
			order.emplace_back(g->lanes[l]->built_[k].first, l);
	std::sort(order.begin(), order.end());
	for (auto &o : order)
	{
		Trainer &t = *g->lanes[o.second];
		m.cls.emplace_back(std::move(t.built_[next[o.second]++].second));
	}
	hibag_b200_train_stats &ts = m.train_stats;
	ts.seconds_total += now_s() - t0;
	for (int l = 0; l < n_lanes; l++)
	{
		Trainer &t = *g->lanes[l];
		const hibag_b200_train_stats &a = t.ts_;
		ts.seconds_em += a.seconds_em; ts.seconds_gpu_wait += a.seconds_gpu_wait;
		ts.gpu_kernel_ms += a.gpu_kernel_ms; ts.pair_evals += a.pair_evals;
		ts.popc32_issued += a.popc32_issued; ts.n_oob_evals += a.n_oob_evals;
		ts.n_ib_evals += a.n_ib_evals; ts.n_em += a.n_em; ts.kernel_launches += a.kernel_launches;
		ts.h2d_bytes += a.h2d_bytes; ts.d2h_bytes += a.d2h_bytes;
		ts.cell_kernel_ms += a.cell_kernel_ms; ts.cell_kernel_launches += a.cell_kernel_launches;
		ts.seconds_prepare += a.seconds_prepare; ts.seconds_phase_oob += a.seconds_phase_oob;
		ts.seconds_phase_ib += a.seconds_phase_ib;
		ts.em_kernel_ms += a.em_kernel_ms; ts.n_em_host_fallback += a.n_em_host_fallback;
		ts.pair_evals_nominal += a.pair_evals_nominal; ts.n_screen_fallback += a.n_screen_fallback;
		ts.gather_kernel_ms += a.gather_kernel_ms; ts.gather_kernel_launches += a.gather_kernel_launches;
		ts.gather_ib_kernel_ms += a.gather_ib_kernel_ms; ts.gather_ib_launches += a.gather_ib_launches;
		ts.gather_ib_popc32 += a.gather_ib_popc32;
		ts.em_iterations += a.em_iterations; ts.em_chain_adds += a.em_chain_adds;
		ts.em_pair_updates += a.em_pair_updates;
		m.train_trace.insert(m.train_trace.end(), t.trace_.begin(), t.trace_.end());
		t.built_.clear();
	}
	m.pcache.reset();
}

}  // namespace hb

hibag_b200_model::hibag_b200_model()
{
	memset(&train_stats, 0, sizeof(train_stats));
	memset(&predict_stats, 0, sizeof(predict_stats));
}
