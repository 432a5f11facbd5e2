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
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>

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
};

class Trainer : public TrainSession
{
public:
	Trainer(hibag_b200_model &m, const hibag_b200_train_opts &o);
	~Trainer() override;
	void run();
	/// can this session (threads, scoring mode) serve a call with these options?
	bool compatible(const hibag_b200_train_opts &o) const
	{
		return (o.use_legacy_hooks != 0) == (procs_ != nullptr) && o.n_threads == requested_threads_ &&
			(o.mtry <= 0 ? 1 : o.mtry) == o_.mtry;
	}
	void configure(const hibag_b200_train_opts &o)
	{
		o_ = o;
		if (o_.mtry <= 0) o_.mtry = 1;
	}

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
	DevBuf<int> d_a1_, d_a2_, d_oob_, d_ib_;
	DevBuf<uint32_t> d_s1_, d_s2_;           // base bit planes [4][n_samp]
	PinBuf<uint32_t> h_planes_;
	Stream main_st_;
	std::vector<std::unique_ptr<EvalSlot> > slots_;   // one per candidate of a round
	std::vector<EmScratch> scratch_;                  // one per worker
	std::vector<double> em_seconds_, wait_seconds_;   // per worker

	int64_t cl_global_index_ = 0;
	int requested_threads_ = 0;
	hibag_gpu_ext_proc *procs_ = nullptr;    // legacy-hook mode
	ScoreStats stats_;
};

Trainer::Trainer(hibag_b200_model &m, const hibag_b200_train_opts &o) : m_(m), o_(o)
{
	dev_ = &current_device();
	n_samp_ = m.n_samp; n_snp_ = m.n_snp; n_hla_ = m.n_hla;
	if (n_samp_ <= 0 || m.geno_t.empty())
		throw std::runtime_error("train: no training data (call hibag_b200_model_set_training)");
	if (o_.mtry <= 0) o_.mtry = 1;
	requested_threads_ = o.n_threads;
	int nt = o_.n_threads;
	if (nt <= 0) nt = (int)std::thread::hardware_concurrency();
	if (nt < 1) nt = 1;
	if (o_.n_threads <= 0)
	{
		// one worker per candidate of a round when that is a modest over-subscription: 23
		// equal EM jobs on 16 cores finish in ~1.4 job-times instead of 2
		if (o_.mtry > nt && o_.mtry <= 2 * nt) nt = o_.mtry;
		if (nt > std::max(o_.mtry, 4)) nt = std::max(o_.mtry, 4);
	}
	pool_.reset(new ThreadPool(nt, dev_->device));
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
	HB_CUDA(cudaStreamSynchronize(main_st_.s));
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

void Trainer::run()
{
	const double t0 = now_s();
	// the session outlives a call: start this call's counters from zero
	stats_ = ScoreStats();
	for (auto &sl : slots_) if (sl) sl->stats = ScoreStats();
	const ScoreStats plugin_before = procs_ ? plugin_build_stats() : ScoreStats();
	std::fill(em_seconds_.begin(), em_seconds_.end(), 0.0);
	std::fill(wait_seconds_.begin(), wait_seconds_.end(), 0.0);
	if (!o_.per_classifier_seed) rng_.set_seed((uint32_t)o_.seed);
	const int stride = o_.index_stride > 0 ? o_.index_stride : 1;
	for (int c = 0; c < o_.nclassifier; c++)
	{
		const int global_k = o_.first_index + c * stride;
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
		m_.cls.emplace_back();
		Classifier &cl = m_.cls.back();
		cl.samp_num = boot_;
		if (o_.verbose)
		{
			fprintf(stderr, "=== building individual classifier %d, out-of-bag (%d/%.1f%%) ===\n",
				global_k + 1, n_samp_ - n_unique, 100.0 * (n_samp_ - n_unique) / n_samp_);
		}
		grow(cl);
		if (o_.verbose)
		{
			fprintf(stderr, "[%d] oob acc: %0.2f%%, # of SNPs: %d, # of haplo: %d\n", global_k + 1,
				cl.oob_acc * 100, (int)cl.snpidx.size(), (int)cl.haplo.h.size());
		}
	}
	// fold the counters
	for (auto &s : slots_) if (s) stats_.add(s->stats);
	if (procs_)
	{
		ScoreStats now = plugin_build_stats();
		now.pair_evals -= plugin_before.pair_evals; now.popc32 -= plugin_before.popc32;
		now.launches -= plugin_before.launches; now.cell_launches -= plugin_before.cell_launches;
		now.h2d_bytes -= plugin_before.h2d_bytes; now.d2h_bytes -= plugin_before.d2h_bytes;
		now.kernel_ms -= plugin_before.kernel_ms; now.cell_ms -= plugin_before.cell_ms;
		stats_.add(now);
	}
	hibag_b200_train_stats &ts = m_.train_stats;
	ts.seconds_total += now_s() - t0;
	for (double v : em_seconds_) ts.seconds_em += v;
	for (double v : wait_seconds_) ts.seconds_gpu_wait += v;
	ts.gpu_kernel_ms += stats_.kernel_ms;
	ts.pair_evals += stats_.pair_evals;
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
	hibag_b200_train_stats &ts = m_.train_stats;

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
		stats_.h2d_bytes += sizeof(int) * (size_t)n_samp_;
		upload_base_geno();
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

	while (pool.total() > 0 && (int)cl.snpidx.size() < HIBAG_B200_MAX_SNP)
	{
		const double t_prep = now_s();
		prepare_round(cur, geno_, a1_, a2_, boot_, inbag_, rp, pool_parallel_for, &pf);
		ts.seconds_prepare += now_s() - t_prep;

		pool.random_select(o_.mtry, rng_);
		const int m = pool.n_selected();
		if ((int)cand.size() < m) cand.resize(m);
		while ((int)slots_.size() < m && !procs_) slots_.emplace_back(new EvalSlot());
		const int bit = cur.n_snp;

		// ---- candidates in parallel: EM, out-of-bag accuracy, and -- speculatively -- the in-bag
		// loss. The reference computes the loss only when acc >= the running maximum over the
		// earlier candidates (:2033). A worker that finishes candidate i knows the accuracies of
		// the earlier candidates that are already done; their maximum is a lower bound of the true
		// threshold, so "acc >= that bound" never misses a loss the reference would compute. Losses
		// computed in excess are discarded below (the reference keeps loss = 0 for them).
		const double t_p1 = now_s();
		std::vector<std::atomic<int> > pub(m);          // published accuracy, -1 = not yet, -2 = skipped
		for (int i = 0; i < m; i++) pub[i].store(-1, std::memory_order_relaxed);
		std::vector<unsigned char> have_loss(m, 0);
		pool_->run(m, [&](int i, int w) {
			Candidate &cd = cand[i];
			cd.snp = pool.at(i);
			cd.acc = 0; cd.loss = 0;
			const double t_em = now_s();
			cd.valid = estimate_candidate(cur, rp, m_.geno_t.data() + (size_t)cd.snp * n_samp_,
				n_samp_, rare_prob, scratch_[w], cd.list);
			em_seconds_[w] += now_s() - t_em;
			if (!cd.valid) { pub[i].store(-2, std::memory_order_release); return; }
			if (procs_) return;
			EvalSlot &sl = *slots_[i];
			sl.stage_list(cd.list.h.data(), (int)cd.list.h.size(), n_hla_, cd.list.n_snp);
			GenoView v = view;
			v.cand_col = d_geno_t_.get() + (size_t)cd.snp * n_samp_;
			v.cand_bit = bit;
			sl.enqueue_cells(v, d_oob_.get(), n_oob);
			sl.enqueue_reduce_oob(v, d_oob_.get(), n_oob);
			double t_w = now_s();
			sl.sync();
			wait_seconds_[w] += now_s() - t_w;
			cd.acc = sl.oob_count();
			pub[i].store(cd.acc, std::memory_order_release);
			int bound = global_max_acc;
			for (int j = 0; j < i; j++)
			{
				const int a = pub[j].load(std::memory_order_acquire);
				if (a > bound) bound = a;
			}
			if (cd.acc >= bound)
			{
				sl.enqueue_cells(v, d_ib_.get(), (int)inbag_.size());
				sl.enqueue_reduce_ib(v, d_ib_.get(), (int)inbag_.size());
				t_w = now_s();
				sl.sync();
				wait_seconds_[w] += now_s() - t_w;
				cd.loss = ib_loss(sl.ib_ratios());
				have_loss[i] = 1;
			}
		});
		for (int i = 0; i < m; i++) if (cand[i].valid) { ts.n_em++; }
		ts.seconds_phase_oob += now_s() - t_p1;
		const double t_p2 = now_s();

		if (!procs_)
		{
			// ---- exact rule in candidate order: keep a loss only where the reference has one ------
			int running = global_max_acc;
			for (int i = 0; i < m; i++)
			{
				Candidate &cd = cand[i];
				if (!cd.valid) continue;
				ts.n_oob_evals++;
				const bool need = cd.acc >= running;
				if (need && !have_loss[i])
					throw std::runtime_error("internal error: in-bag loss missing for a candidate");
				if (!need) cd.loss = 0;
				else ts.n_ib_evals++;
				if (cd.acc > running) running = cd.acc;
			}
		} else {
			// ---- legacy hooks: sequential, exactly the reference's call sequence -------------
			int running = global_max_acc;
			for (int i = 0; i < m; i++)
			{
				Candidate &cd = cand[i];
				if (!cd.valid) continue;
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
				const double t_w = now_s();
				procs_->build_set_haplo_geno(cd.list.h.data(), (int)cd.list.h.size(), aos_.data(),
					cd.list.n_snp);
				cd.acc = procs_->build_acc_oob();
				ts.n_oob_evals++;
				if (cd.acc >= running)
				{
					cd.loss = procs_->build_acc_ib();
					ts.n_ib_evals++;
				}
				wait_seconds_[0] += now_s() - t_w;
				if (cd.acc > running) running = cd.acc;
			}
		}

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
				m_.train_trace.insert(m_.train_trace.end(), row, row + 4);
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
		m_.train_trace.insert(m_.train_trace.end(), row, row + 4);
	}
	cl.haplo = cur;
	cl.oob_acc = 0.5 * global_max_acc / n_oob;                       // :2121
}

}  // namespace

void train_model(hibag_b200_model &m, const hibag_b200_train_opts &opts)
{
	current_device();      // fails here, loudly, when no CUDA device is usable
	Trainer *t = dynamic_cast<Trainer *>(m.tsession.get());
	if (!t || !t->compatible(opts))
	{
		m.tsession.reset();
		t = new Trainer(m, opts);
		m.tsession.reset(t);
	}
	t->configure(opts);
	try
	{
		t->run();
	} catch (...)
	{
		m.tsession.reset();
		throw;
	}
	m.pcache.reset();
}

}  // namespace hb

hibag_b200_model::hibag_b200_model()
{
	memset(&train_stats, 0, sizeof(train_stats));
	memset(&predict_stats, 0, sizeof(predict_stats));
}
