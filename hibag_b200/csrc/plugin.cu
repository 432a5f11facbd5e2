// plugin.cu -- the drop-in boundary: HIBAG's ten TypeGPUExtProc hooks
// (reference inst/include/LibHLA_ext.h:357-388) implemented on the sm_100a kernels, plus the
// stateless batched scoring entry points of include/hibag_b200.h.
//
// Hook contract (call sites in the reference, src/LibHLA.cpp):
//   build_init :2256-2260 | build_done :2262-2266 | build_set_bootstrap :2290-2293
//   build_set_haplo_geno :1913-1921 | build_acc_oob :1938-1941 | build_acc_ib :1961-1964
//   predict_init :2498-2523 | predict_done :2525-2531 | predict_avg_prob :2433-2441
// The hooks are process-global state (one model at a time), called from one thread, and
// report failure by throwing std::exception like any other plugin of the reference would.

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <ctime>
#include <cstring>
#include <memory>
#include <vector>

#include "em.h"
#include "scorer.h"

namespace hb {

// ---------------------------------------------------------------------------------------------
// training hooks
// ---------------------------------------------------------------------------------------------
/// One of the two device-side copies of the training state. A candidate SNP is evaluated on the
/// context the previous candidate did NOT use, so that the speculative in-bag pass of the previous
/// candidate (see hook_build_acc_oob) may still be running while this one is uploaded and scored.
struct BuildCtx
{
	GenoSet geno;
	PinBuf<unsigned char> h_aos;      // pinned SHADOW of the TGenotype[] this context's planes were built from
	DevBuf<unsigned char> d_aos;
	bool shadow_valid = false;
	PinBuf<int8_t> h_code;            // [4][n_samp] 2-bit codes of up to four changed SNP columns
	DevBuf<int8_t> d_code;
	EvalSlot slot_oob{true, true};    // high-priority stream, spinning sync: the host waits on this one
	EvalSlot slot_ib{false, true};    // the speculative in-bag pass
	Event ev_ready{false};            // planes + list of this context are on the device
	bool ib_pending = false;          // an in-bag pass is enqueued on slot_ib and not yet synchronised
};

struct BuildState
{
	int n_hla = 0, n_samp = 0, n_snp = 0;
	BuildCtx ctx[2];
	int cur = 0;
	uint64_t n_full_uploads = 0, n_column_uploads = 0, n_unchanged = 0, n_ib_wasted = 0;
	double t_set = 0, t_oob = 0, t_ib = 0;        // seconds inside the three hooks
	std::vector<int> boot;            // bootstrap multiplicities currently on the device
	std::vector<int> oob, ib;         // ascending sample indices (src/LibHLA.cpp:1858-1874)
	DevBuf<int> d_oob, d_ib;
	bool have_lists = false;
	bool have_list_staged = false;
	bool speculate = true;            // HIBAG_B200_HOOK_SPECULATE=0 turns the speculative in-bag pass off
	ScoreStats retired;               // stats of slots are folded here

	void drain()
	{
		for (BuildCtx &c : ctx)
		{
			if (c.ib_pending) { c.slot_ib.sync(); c.ib_pending = false; }
			HB_CUDA(cudaStreamSynchronize(c.slot_oob.stream()));
		}
	}

	void set_bootstrap(const int *cnt)
	{
		drain();                          // the sample lists are shared by both contexts
		boot.assign(cnt, cnt + n_samp);
		oob.clear(); ib.clear();
		for (int i = 0; i < n_samp; i++)
			(cnt[i] > 0 ? ib : oob).push_back(i);
		d_oob.ensure(n_samp); d_ib.ensure(n_samp);
		cudaStream_t st = ctx[0].slot_oob.stream();
		// pageable -> device copies are synchronous with respect to the host buffer
		if (!oob.empty())
			HB_CUDA(cudaMemcpyAsync(d_oob.get(), oob.data(), sizeof(int) * oob.size(),
				cudaMemcpyHostToDevice, st));
		if (!ib.empty())
			HB_CUDA(cudaMemcpyAsync(d_ib.get(), ib.data(), sizeof(int) * ib.size(),
				cudaMemcpyHostToDevice, st));
		HB_CUDA(cudaStreamSynchronize(st));
		ctx[0].slot_oob.stats.h2d_bytes += sizeof(int) * (size_t)n_samp;
		have_lists = true;
	}

	static GenoView view(const BuildCtx &c, int n_samp)
	{
		GenoView v;
		v.s1 = c.geno.s1.get(); v.s2 = c.geno.s2.get(); v.stride = n_samp;
		v.a1 = c.geno.a1.get(); v.a2 = c.geno.a2.get();
		return v;
	}

	ScoreStats stats() const
	{
		ScoreStats s = retired;
		for (const BuildCtx &c : ctx) { s.add(c.slot_oob.stats); s.add(c.slot_ib.stats); }
		return s;
	}
};

static std::unique_ptr<BuildState> g_build;

static double hook_now()
{
	struct timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

static void hook_build_init(int n_hla, int n_sample)
{
	if (n_hla <= 0 || n_sample <= 0) throw std::runtime_error("build_init: invalid sizes");
	current_device();
	g_build.reset(new BuildState());
	g_build->n_hla = n_hla;
	g_build->n_samp = n_sample;
	for (BuildCtx &c : g_build->ctx)
	{
		c.geno.ensure(n_sample);
		c.h_aos.ensure(sizeof(hibag_genotype) * (size_t)n_sample);
		c.d_aos.ensure(sizeof(hibag_genotype) * (size_t)n_sample);
		c.h_code.ensure(4 * (size_t)n_sample);
		c.d_code.ensure(4 * (size_t)n_sample);
		int r = 2;                        // samples per lane of the single-list passes (latency-bound)
		if (const char *e = getenv("HIBAG_B200_HOOK_R")) r = atoi(e);
		c.slot_oob.set_samples_per_lane(r);
		c.slot_ib.set_samples_per_lane(r);
	}
	if (const char *e = getenv("HIBAG_B200_HOOK_SPECULATE")) g_build->speculate = atoi(e) != 0;
}

static void hook_build_done()
{
	if (g_build)
	{
		try { g_build->drain(); } catch (...) {}
		if (getenv("HIBAG_B200_HOOK_PROF"))
			fprintf(stderr, "hooks: set_haplo_geno %.3f s (%llu full, %llu column, %llu unchanged uploads), acc_oob %.3f s, "
				"acc_ib %.3f s, speculative in-bag passes not asked for: %llu\n", g_build->t_set,
				(unsigned long long)g_build->n_full_uploads, (unsigned long long)g_build->n_column_uploads,
				(unsigned long long)g_build->n_unchanged, g_build->t_oob, g_build->t_ib,
				(unsigned long long)g_build->n_ib_wasted);
	}
	g_build.reset();
}

static void hook_build_set_bootstrap(const int cnt[])
{
	if (!g_build) throw std::runtime_error("build_set_bootstrap called before build_init");
	g_build->set_bootstrap(cnt);
}

static void hook_build_set_haplo_geno(const hibag_haplotype haplo[], int n_haplo,
	const hibag_genotype geno[], int n_snp)
{
	BuildState *b = g_build.get();
	if (!b) throw std::runtime_error("build_set_haplo_geno called before build_init");
	const double t0 = hook_now();
	// the genotype array carries the bootstrap counts too; (re)build the sample lists when the
	// host did not call build_set_bootstrap or the counts changed
	bool same = b->have_lists;
	if (same)
		for (int i = 0; i < b->n_samp; i++)
			if (geno[i].bootstrap_count != b->boot[i]) { same = false; break; }
	if (!same)
	{
		std::vector<int> cnt(b->n_samp);
		for (int i = 0; i < b->n_samp; i++) cnt[i] = geno[i].bootstrap_count;
		b->set_bootstrap(cnt.data());
	}
	b->n_snp = n_snp;
	// the other context: whatever the previous candidate left running (its speculative in-bag pass)
	// is not touched; this context's own in-bag pass is two candidates old
	b->cur ^= 1;
	BuildCtx &c = b->ctx[b->cur];
	if (c.ib_pending) { c.slot_ib.sync(); c.ib_pending = false; b->n_ib_wasted++; }
	cudaStream_t st = c.slot_oob.stream();
	const size_t bytes = sizeof(hibag_genotype) * (size_t)b->n_samp;
	// The reference hands over the whole TGenotype[nSample] with every candidate SNP (240 KB at
	// 5,000 samples), but between two calls it has only rewritten the candidate's bit column
	// (CGenotypeList::AddSNP / ReduceSNP, src/LibHLA.cpp:860-881) -- or two columns after an accepted
	// SNP. This context's previous evaluation has been synchronised, so its pinned shadow is free:
	// diff against it, and when the change is confined to <= 4 bit columns upload those columns as
	// 2-bit codes (n bytes each) and patch the device planes; otherwise upload everything.
	hibag_genotype *shadow = (hibag_genotype *)c.h_aos.get();
	uint64_t d1[2] = { 0, 0 }, d2[2] = { 0, 0 };
	bool other = !c.shadow_valid;
	if (!other)
	{
		for (int i = 0; i < b->n_samp; i++)
		{
			const hibag_genotype &x = geno[i], &y = shadow[i];
			d1[0] |= (uint64_t)(x.snp1[0] ^ y.snp1[0]); d1[1] |= (uint64_t)(x.snp1[1] ^ y.snp1[1]);
			d2[0] |= (uint64_t)(x.snp2[0] ^ y.snp2[0]); d2[1] |= (uint64_t)(x.snp2[1] ^ y.snp2[1]);
			other |= (x.bootstrap_count != y.bootstrap_count) | (x.allele1 != y.allele1) | (x.allele2 != y.allele2);
		}
	}
	const uint64_t u[2] = { d1[0] | d2[0], d1[1] | d2[1] };
	const int n_changed = __builtin_popcountll(u[0]) + __builtin_popcountll(u[1]);
	if (other || n_changed > 4)
	{
		memcpy(shadow, geno, bytes);
		HB_CUDA(cudaMemcpyAsync(c.d_aos.get(), c.h_aos.get(), bytes, cudaMemcpyHostToDevice, st));
		c.slot_oob.stats.h2d_bytes += bytes;
		launch_unpack_genotypes(c.d_aos.get(), b->n_samp, c.geno.s1.get(), c.geno.s2.get(),
			b->n_samp, c.geno.a1.get(), c.geno.a2.get(), c.geno.boot.get(), st);
		c.slot_oob.stats.launches++;
		c.shadow_valid = true;
		b->n_full_uploads++;
	} else if (n_changed > 0)
	{
		int k = 0;
		for (int w = 0; w < 2; w++)
			for (uint64_t m = u[w]; m; m &= m - 1, k++)
			{
				const int bit = __builtin_ctzll(m);
				const uint64_t one = (uint64_t)1 << bit;
				int8_t *code = c.h_code.get() + (size_t)k * b->n_samp;
				for (int i = 0; i < b->n_samp; i++)
				{
					const uint64_t a = (uint64_t)geno[i].snp1[w] & one, cc = (uint64_t)geno[i].snp2[w] & one;
					code[i] = (int8_t)((a ? 1 : 0) | (cc ? 2 : 0));
					shadow[i].snp1[w] = (int64_t)(((uint64_t)shadow[i].snp1[w] & ~one) | a);
					shadow[i].snp2[w] = (int64_t)(((uint64_t)shadow[i].snp2[w] & ~one) | cc);
				}
			}
		HB_CUDA(cudaMemcpyAsync(c.d_code.get(), c.h_code.get(), (size_t)k * b->n_samp,
			cudaMemcpyHostToDevice, st));
		c.slot_oob.stats.h2d_bytes += (size_t)k * b->n_samp;
		k = 0;
		for (int w = 0; w < 2; w++)
			for (uint64_t m = u[w]; m; m &= m - 1, k++)
			{
				launch_patch_column(c.d_code.get() + (size_t)k * b->n_samp, b->n_samp, c.geno.s1.get(),
					c.geno.s2.get(), b->n_samp, 64 * w + __builtin_ctzll(m), st);
				c.slot_oob.stats.launches++;
			}
		b->n_column_uploads++;
	} else
		b->n_unchanged++;
	c.slot_oob.stage_list(haplo, n_haplo, b->n_hla, n_snp);
	HB_CUDA(cudaEventRecord(c.ev_ready.e, st));
	b->have_list_staged = true;
	b->t_set += hook_now() - t0;
}

/// optional hook (src/LibHLA.cpp:1014-1072): the host expands every record into the 3-4 doubled
/// pairs itself and free()s the buffer (:1062)
static uint32_t *hook_build_haplomatch(const hibag_haplotype haplo[], const size_t n_haplo[],
	int n_snp, const hibag_genotype geno[], size_t *out_n)
{
	BuildState *b = g_build.get();
	if (!b) throw std::runtime_error("build_haplomatch called before build_init");
	// in-bag list from the genotypes' own bootstrap counts (identical to build_set_bootstrap's)
	std::vector<int> ib;
	for (int i = 0; i < b->n_samp; i++)
		if (geno[i].bootstrap_count > 0) ib.push_back(i);
	b->ctx[0].slot_oob.stats.launches += 3;
	return haplomatch_records(haplo, n_haplo, b->n_hla, n_snp, geno, b->n_samp, ib, out_n);
}

/// enqueue the in-bag pass of the current context on its in-bag slot (after planes + list landed)
static void enqueue_ib(BuildState *b, BuildCtx &c)
{
	const int n = (int)b->ib.size();
	const GenoView v = BuildState::view(c, b->n_samp);
	HB_CUDA(cudaStreamWaitEvent(c.slot_ib.stream(), c.ev_ready.e, 0));
	c.slot_ib.borrow_list(c.slot_oob);
	c.slot_ib.enqueue_cells(v, b->d_ib.get(), n);
	c.slot_ib.enqueue_reduce_ib(v, b->d_ib.get(), n);
	c.ib_pending = true;
}

static int hook_build_acc_oob()
{
	BuildState *b = g_build.get();
	if (!b || !b->have_list_staged)
		throw std::runtime_error("build_acc_oob called before build_set_haplo_geno");
	const double t0 = hook_now();
	BuildCtx &c = b->ctx[b->cur];
	const int n = (int)b->oob.size();
	int result = 0;
	if (n > 0)
	{
		const GenoView v = BuildState::view(c, b->n_samp);
		c.slot_oob.enqueue_cells(v, b->d_oob.get(), n);
		c.slot_oob.enqueue_reduce_oob(v, b->d_oob.get(), n);
	}
	// The reference asks for the in-bag loss right after this call whenever the accuracy is not below
	// its running maximum (src/LibHLA.cpp:2031-2034) -- more than half of the candidates. The hook
	// protocol is strictly sequential and a single-list pass is bound by the latency of its longest
	// fp64 chain, not by throughput, so the in-bag pass is started NOW on a second stream beside the
	// out-of-bag pass; build_acc_ib then only waits for it. An in-bag pass that is never asked for
	// finishes in the background on this context while the next candidate uses the other one.
	if (b->speculate && !b->ib.empty()) enqueue_ib(b, c);
	if (n > 0)
	{
		c.slot_oob.sync();
		result = c.slot_oob.oob_count();
	}
	b->t_oob += hook_now() - t0;
	return result;
}

static double hook_build_acc_ib()
{
	BuildState *b = g_build.get();
	if (!b || !b->have_list_staged)
		throw std::runtime_error("build_acc_ib called before build_set_haplo_geno");
	const double t0 = hook_now();
	BuildCtx &c = b->ctx[b->cur];
	const int n = (int)b->ib.size();
	if (!c.ib_pending) enqueue_ib(b, c);
	c.slot_ib.sync();
	c.ib_pending = false;
	// log() and the in-bag-order sum stay on the host (glibc log, sequential order of
	// src/LibHLA.cpp:1966-1977)
	const double *ratio = c.slot_ib.ib_ratios();
	double loglik = 0;
	for (int i = 0; i < n; i++)
		loglik += b->boot[b->ib[i]] * std::log(ratio[i]);
	b->t_ib += hook_now() - t0;
	return loglik * -2;
}

// ---------------------------------------------------------------------------------------------
// prediction hooks (one sample per call: the granularity the reference imposes, :2433-2441)
// ---------------------------------------------------------------------------------------------
struct PredictHookState
{
	int n_hla = 0, n_cls = 0, n_cells = 0;
	std::vector<ListBlob> blobs;
	std::vector<size_t> blob_off;
	DevBuf<unsigned char> d_blobs;
	GenoSet geno;                       // "sample" c = the genotype packed for classifier c
	PinBuf<unsigned char> h_aos;
	DevBuf<unsigned char> d_aos;
	PinBuf<double> h_w;
	DevBuf<double> d_w;
	DevBuf<int> d_idx;
	DevBuf<double> P, acc, aux, d_match;
	PinBuf<double> h_out;
	DevBuf<unsigned int> counters;
	Stream st;
	ScoreStats stats;
};

static std::unique_ptr<PredictHookState> g_pred;

static void hook_predict_init(int n_hla, int n_classifier,
	const hibag_haplotype *const p_haplo[], const int n_haplo[], const int n_snp[])
{
	if (n_hla <= 0 || n_classifier < 0) throw std::runtime_error("predict_init: invalid sizes");
	current_device();
	std::unique_ptr<PredictHookState> s(new PredictHookState());
	s->n_hla = n_hla; s->n_cls = n_classifier;
	s->n_cells = n_hla * (n_hla + 1) / 2;
	size_t total = 0;
	s->blob_off.resize(n_classifier);
	for (int c = 0; c < n_classifier; c++)
	{
		s->blob_off[c] = total;
		total += (list_blob_capacity(n_haplo[c], n_snp[c], n_hla) + 255) & ~(size_t)255;
	}
	std::vector<unsigned char> host(total + 256);
	unsigned char *hbase = (unsigned char *)(((uintptr_t)host.data() + 15) & ~(uintptr_t)15);
	s->blobs.resize(n_classifier);
	for (int c = 0; c < n_classifier; c++)
		s->blobs[c] = build_list_blob(p_haplo[c], n_haplo[c], n_hla, n_snp[c],
			hbase + s->blob_off[c], 64);
	s->d_blobs.ensure(total + 256);
	if (total) HB_CUDA(cudaMemcpy(s->d_blobs.get(), hbase, total, cudaMemcpyHostToDevice));
	const int nc = n_classifier > 0 ? n_classifier : 1;
	s->geno.ensure(nc);
	s->h_aos.ensure(sizeof(hibag_genotype) * (size_t)nc);
	s->d_aos.ensure(sizeof(hibag_genotype) * (size_t)nc);
	s->h_w.ensure(nc); s->d_w.ensure(nc);
	std::vector<int> idx(nc);
	for (int c = 0; c < nc; c++) idx[c] = c;
	s->d_idx.ensure(nc);
	HB_CUDA(cudaMemcpy(s->d_idx.get(), idx.data(), sizeof(int) * nc, cudaMemcpyHostToDevice));
	s->P.ensure((size_t)s->n_cells * 32);
	s->acc.ensure((size_t)s->n_cells); s->aux.ensure(4); s->d_match.ensure(1);
	s->h_out.ensure((size_t)s->n_cells + 1);
	s->counters.ensure(nc);
	device_rare_freq_table();
	g_pred = std::move(s);
}

static void hook_predict_done()
{
	g_pred.reset();
}

static void hook_predict_avg_prob(const hibag_genotype geno[], const double weight[],
	double out_prob[], double out_match[])
{
	PredictHookState *s = g_pred.get();
	if (!s) throw std::runtime_error("predict_avg_prob called before predict_init");
	const DeviceInfo &di = current_device();
	cudaStream_t st = s->st.s;
	const int nc = s->n_cls;
	memcpy(s->h_aos.get(), geno, sizeof(hibag_genotype) * (size_t)nc);
	memcpy(s->h_w.get(), weight, sizeof(double) * (size_t)nc);
	HB_CUDA(cudaMemcpyAsync(s->d_aos.get(), s->h_aos.get(), sizeof(hibag_genotype) * (size_t)nc,
		cudaMemcpyHostToDevice, st));
	HB_CUDA(cudaMemcpyAsync(s->d_w.get(), s->h_w.get(), sizeof(double) * (size_t)nc,
		cudaMemcpyHostToDevice, st));
	launch_unpack_genotypes(s->d_aos.get(), nc, s->geno.s1.get(), s->geno.s2.get(), nc,
		s->geno.a1.get(), s->geno.a2.get(), s->geno.boot.get(), st);
	HB_CUDA(cudaMemsetAsync(s->acc.get(), 0, sizeof(double) * (size_t)s->n_cells, st));
	HB_CUDA(cudaMemsetAsync(s->aux.get(), 0, sizeof(double) * 4, st));
	HB_CUDA(cudaMemsetAsync(s->counters.get(), 0, sizeof(unsigned int) * (size_t)nc, st));
	const double *tbl = device_rare_freq_table();
	for (int c = 0; c < nc; c++)
	{
		if (!(weight[c] > 0)) continue;             // src/LibHLA.cpp:2451
		CellPass p;
		memset(&p, 0, sizeof(p));
		bind_list(s->blobs[c], s->d_blobs.get() + s->blob_off[c], tbl, p);
		p.s1 = s->geno.s1.get(); p.s2 = s->geno.s2.get(); p.geno_stride = nc;
		p.samp_list = s->d_idx.get() + c; p.n_pos = 1;
		p.task_counter = s->counters.get() + c;
		p.P = s->P.get(); p.p_stride = 32;
		const int nw = launch_cell_pass(p, 1, di.sm_count, st);
		launch_predict_accumulate(s->P.get(), 32, s->n_cells, 1, s->d_w.get() + c,
			s->acc.get(), 1, s->aux.get(), st);
		s->stats.pair_evals += s->blobs[c].pairs_per_sample;
		s->stats.popc32 += s->blobs[c].pairs_per_sample * nw;
		s->stats.launches += 2; s->stats.cell_launches++;
	}
	// aux layout for n_tile = 1: [0] = sum_w, [1] = sum_w*match, [2] = n_used
	launch_predict_finalize(s->acc.get(), 1, s->aux.get(), s->n_hla, 0, 1, nullptr, nullptr,
		nullptr, s->d_match.get(), nullptr, nullptr, st);
	HB_CUDA(cudaMemcpyAsync(s->h_out.get(), s->acc.get(), sizeof(double) * (size_t)s->n_cells,
		cudaMemcpyDeviceToHost, st));
	HB_CUDA(cudaMemcpyAsync(s->h_out.get() + s->n_cells, s->d_match.get(), sizeof(double),
		cudaMemcpyDeviceToHost, st));
	HB_CUDA(cudaStreamSynchronize(st));
	memcpy(out_prob, s->h_out.get(), sizeof(double) * (size_t)s->n_cells);
	out_match[0] = s->h_out.get()[s->n_cells];
	s->stats.launches++;
}

static hibag_gpu_ext_proc g_procs = {
	hook_build_init, hook_build_done, hook_build_set_bootstrap,
	nullptr,   // build_haplomatch: optional; the host keeps its CPU search (src/LibHLA.cpp:1074)
	hook_build_set_haplo_geno, hook_build_acc_oob, hook_build_acc_ib,
	hook_predict_init, hook_predict_done, hook_predict_avg_prob
};

// the same hooks plus build_haplomatch. Kept apart because a host that receives this hook
// builds its pair lists in record order (4 doubled pairs per record, :1052-1059) instead of its
// CPU scan order (:1569-1637): EM sums are then accumulated in a different order and its
// frequencies differ from the CPU path in the last bits -- the reference behaves the same way
// with any plugin that provides this hook.
static hibag_gpu_ext_proc g_procs_hm = {
	hook_build_init, hook_build_done, hook_build_set_bootstrap,
	hook_build_haplomatch,
	hook_build_set_haplo_geno, hook_build_acc_oob, hook_build_acc_ib,
	hook_predict_init, hook_predict_done, hook_predict_avg_prob
};

hibag_gpu_ext_proc *plugin_procs() { return &g_procs; }
hibag_gpu_ext_proc *plugin_procs_with_haplomatch() { return &g_procs_hm; }

ScoreStats plugin_build_stats()
{
	ScoreStats s;
	if (g_build) s = g_build->stats();
	return s;
}

// ---------------------------------------------------------------------------------------------
// stateless batched scoring on host arrays
// ---------------------------------------------------------------------------------------------
enum ScoreKind { SCORE_BEST_GUESS, SCORE_POST_PROB, SCORE_POST_PROB2 };

void score_host_arrays(int kind, const hibag_haplotype *haplo, int n_haplo, int n_hla,
	int n_snp, const hibag_genotype *geno, int n_geno, int32_t *out_a1, int32_t *out_a2,
	double *out_d, double *out_sum)
{
	if (n_geno <= 0) return;
	current_device();
	EvalSlot slot;
	slot.stage_list(haplo, n_haplo, n_hla, n_snp);
	const int n_cells = n_hla * (n_hla + 1) / 2;
	const int tile = 16384;
	GenoSet gs;
	gs.ensure(tile);
	DevBuf<unsigned char> d_aos;
	d_aos.ensure(sizeof(hibag_genotype) * (size_t)tile);
	DevBuf<int> d_a1, d_a2;
	DevBuf<double> d_out, d_sum;
	d_a1.ensure(tile); d_a2.ensure(tile);
	d_sum.ensure(tile);
	if (kind == SCORE_POST_PROB2) d_out.ensure((size_t)tile * n_cells);
	cudaStream_t st = slot.stream();
	for (int begin = 0; begin < n_geno; begin += tile)
	{
		const int n = (n_geno - begin < tile) ? (n_geno - begin) : tile;
		HB_CUDA(cudaMemcpyAsync(d_aos.get(), geno + begin, sizeof(hibag_genotype) * (size_t)n,
			cudaMemcpyHostToDevice, st));
		launch_unpack_genotypes(d_aos.get(), n, gs.s1.get(), gs.s2.get(), tile, gs.a1.get(),
			gs.a2.get(), gs.boot.get(), st);
		GenoView v;
		v.s1 = gs.s1.get(); v.s2 = gs.s2.get(); v.stride = tile;
		v.a1 = gs.a1.get(); v.a2 = gs.a2.get();
		slot.enqueue_cells(v, nullptr, n);
		if (kind == SCORE_BEST_GUESS)
		{
			launch_reduce_best_guess(slot.cell_matrix(), slot.cell_stride(), n_hla, n,
				d_a1.get(), d_a2.get(), st);
			HB_CUDA(cudaMemcpyAsync(out_a1 + begin, d_a1.get(), sizeof(int) * (size_t)n,
				cudaMemcpyDeviceToHost, st));
			HB_CUDA(cudaMemcpyAsync(out_a2 + begin, d_a2.get(), sizeof(int) * (size_t)n,
				cudaMemcpyDeviceToHost, st));
		} else if (kind == SCORE_POST_PROB)
		{
			launch_reduce_ib(slot.cell_matrix(), slot.cell_stride(), n_hla, nullptr, n,
				gs.a1.get(), gs.a2.get(), d_sum.get(), st);
			HB_CUDA(cudaMemcpyAsync(out_d + begin, d_sum.get(), sizeof(double) * (size_t)n,
				cudaMemcpyDeviceToHost, st));
		} else {
			launch_normalize(slot.cell_matrix(), slot.cell_stride(), n_hla, n, d_sum.get(), st);
			launch_transpose(slot.cell_matrix(), slot.cell_stride(), n_cells, n, d_out.get(), st);
			HB_CUDA(cudaMemcpyAsync(out_d + (size_t)begin * n_cells, d_out.get(),
				sizeof(double) * (size_t)n * n_cells, cudaMemcpyDeviceToHost, st));
			HB_CUDA(cudaMemcpyAsync(out_sum + begin, d_sum.get(), sizeof(double) * (size_t)n,
				cudaMemcpyDeviceToHost, st));
		}
		slot.sync();
	}
}

}  // namespace hb
