// predictor.cu -- batched ensemble prediction on the GPU scoring path.
//
// Reference semantics: CAttrBag_Model::PredictHLA / _PredictHLA, CPU branch with
// vote_method = 1 (src/LibHLA.cpp:2317-2482): per sample, per classifier with weight > 0 in
// classifier order: PostProb2 (all cells, sequential sum, normalise), weighted add into the
// ensemble sum, then normalise by the weight sum, matching = sum(w*pm)/sum(w), best guess,
// dosage, probability row. Here the loop nest is turned inside out: for a tile of samples,
// classifiers are processed one after the other on a stream (so every (sample, cell) still
// accumulates in classifier order), and genotype packing + classifier weights
// (TGenotype::IntToSNP :667-706, c_weight :2418-2431) run on the device from the raw int8
// genotype matrix, which crosses PCIe once.

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>

#include "model.h"

namespace hb {

struct PredictCache
{
	int n_cls = 0;
	std::vector<ListBlob> blobs;
	std::vector<size_t> blob_off;
	DevBuf<unsigned char> d_blobs;
	std::vector<size_t> snp_off;
	DevBuf<int> d_snpidx;
	DevBuf<int> d_snp_weight;
	std::vector<int> snp_weight;
	// tile buffers
	int tile = 0;
	DevBuf<uint32_t> s1, s2;
	DevBuf<double> weight, P, acc, aux;
	DevBuf<unsigned int> counters;
	DevBuf<int8_t> geno_t;
	Event ev_begin, ev_end;
	std::vector<std::unique_ptr<Event> > cell_ev;   // begin/end pair per cell-kernel launch
	// exact de-duplication of a tile's genotypes per classifier (kernels.h: launch_dedup_genotypes)
	DevBuf<int> dd_table, dd_repof, dd_uid, dd_rep;
	DevBuf<int> dd_count;                           // distinct genotypes per (tile, classifier) of a call
	DevBuf<double> dd_norm;
	// executed-work accounting of calls whose counts have not come back yet
	struct Pending
	{
		Event done{false};
		PinBuf<int> host;
		std::vector<uint64_t> pairs_per_sample;     // per count entry
		std::vector<int> nw;
	};
	std::vector<std::unique_ptr<Pending> > pending;
	~PredictCache()
	{
		// counts still on their way back land in page-locked blocks that return to the cache here
		for (auto &p : pending) cudaEventSynchronize(p->done.e);
	}
};

/// PREDICT_DEDUP: score each distinct packed genotype of a tile once per classifier (default on;
/// HIBAG_B200_PREDICT_DEDUP=0 scores every sample as the reference does -- same bits either way)
static bool predict_dedup_enabled()
{
	const char *e = getenv("HIBAG_B200_PREDICT_DEDUP");      // read per call: tests switch it
	return !(e && e[0] == '0');
}

/// fold the distinct-genotype counts that have arrived into the model's statistics
static void collect_pending(hibag_b200_model &m, PredictCache &pc, bool wait)
{
	hibag_b200_predict_stats &ps = m.predict_stats;
	for (size_t k = 0; k < pc.pending.size();)
	{
		PredictCache::Pending &pd = *pc.pending[k];
		if (wait) HB_CUDA(cudaEventSynchronize(pd.done.e));
		else if (cudaEventQuery(pd.done.e) != cudaSuccess) { cudaGetLastError(); k++; continue; }
		for (size_t i = 0; i < pd.pairs_per_sample.size(); i++)
		{
			const uint64_t n = (uint64_t)pd.host.get()[i];
			ps.pair_evals += pd.pairs_per_sample[i] * n;
			ps.popc32_issued += pd.pairs_per_sample[i] * n * (uint64_t)pd.nw[i];
			ps.positions_scored += n;
		}
		pc.pending.erase(pc.pending.begin() + k);
	}
}

void predict_collect_stats(hibag_b200_model &m)
{
	if (m.pcache) collect_pending(m, *m.pcache, true);
}

void snp_weights(const hibag_b200_model &m, std::vector<int> &w)
{
	w.assign(m.n_snp, 0);                         // src/LibHLA.cpp:2484-2496
	for (const Classifier &c : m.cls)
		for (int k : c.snpidx) w[k]++;
}

static PredictCache &get_cache(hibag_b200_model &m)
{
	if (m.pcache && m.pcache->n_cls == (int)m.cls.size()) return *m.pcache;
	std::shared_ptr<PredictCache> pc(new PredictCache());
	pc->n_cls = (int)m.cls.size();
	const int nc = pc->n_cls;
	pc->blobs.resize(nc); pc->blob_off.resize(nc); pc->snp_off.resize(nc + 1);
	size_t total = 0, n_idx = 0;
	for (int c = 0; c < nc; c++)
	{
		const Classifier &cl = m.cls[c];
		pc->blob_off[c] = total;
		total += (list_blob_capacity((int)cl.haplo.h.size(), (int)cl.snpidx.size(), m.n_hla) + 255) & ~(size_t)255;
		pc->snp_off[c] = n_idx;
		n_idx += cl.snpidx.size();
	}
	pc->snp_off[nc] = n_idx;
	std::vector<unsigned char> host(total + 256);
	unsigned char *hbase = (unsigned char *)(((uintptr_t)host.data() + 15) & ~(uintptr_t)15);
	std::vector<int> idx(n_idx + 1);
	for (int c = 0; c < nc; c++)
	{
		Classifier &cl = m.cls[c];
		cl.haplo.n_snp = (int)cl.snpidx.size();
		cl.haplo.set_tags();
		pc->blobs[c] = build_list_blob(cl.haplo.h.data(), (int)cl.haplo.h.size(), m.n_hla,
			(int)cl.snpidx.size(), hbase + pc->blob_off[c]);
		for (size_t k = 0; k < cl.snpidx.size(); k++)
		{
			if (cl.snpidx[k] < 0 || cl.snpidx[k] >= m.n_snp)
				throw std::runtime_error("predict: classifier SNP index out of range");
			idx[pc->snp_off[c] + k] = cl.snpidx[k];
		}
	}
	pc->d_blobs.ensure(total + 256);
	if (total) HB_CUDA(cudaMemcpy(pc->d_blobs.get(), hbase, total, cudaMemcpyHostToDevice));
	pc->d_snpidx.ensure(n_idx + 1);
	if (n_idx) HB_CUDA(cudaMemcpy(pc->d_snpidx.get(), idx.data(), sizeof(int) * n_idx, cudaMemcpyHostToDevice));
	snp_weights(m, pc->snp_weight);
	pc->d_snp_weight.ensure(m.n_snp + 1);
	HB_CUDA(cudaMemcpy(pc->d_snp_weight.get(), pc->snp_weight.data(), sizeof(int) * m.n_snp, cudaMemcpyHostToDevice));
	pc->counters.ensure(nc + 1);
	m.predict_stats.h2d_bytes += total + sizeof(int) * (n_idx + m.n_snp);
	m.pcache = pc;
	return *pc;
}

static int pick_tile(int n_samp, int n_cells)
{
	// cell matrix + accumulator of one tile: 2 * 8 * n_cells * tile bytes; keep under ~6 GB.
	// Large tiles: the more samples a tile holds, the smaller its share of distinct genotypes.
	const char *e = getenv("HIBAG_B200_PREDICT_TILE");
	const long long v = e ? atoll(e) : 0;
	long long t = (v >= 1024) ? v : 262144LL;
	while (t > 1024 && 16LL * n_cells * t > (6LL << 30)) t >>= 1;
	if (t > n_samp) t = ((n_samp + 127) / 128) * 128;
	return (int)t;
}

void predict_device(hibag_b200_model &m, const int8_t *geno_dev, int n_samp,
	const hibag_b200_predict_out &out, const int32_t *snp_weight_dev, double *partial_dev,
	cudaStream_t st, bool sync)
{
	if (n_samp <= 0) return;
	const DeviceInfo &di = current_device();
	PredictCache &pc = get_cache(m);
	const int n_hla = m.n_hla;
	const int n_cells = n_hla * (n_hla + 1) / 2;
	const int nc = pc.n_cls;
	const int tile = pick_tile(n_samp, n_cells);
	pc.s1.ensure((size_t)4 * tile); pc.s2.ensure((size_t)4 * tile);
	pc.weight.ensure(tile);
	pc.P.ensure((size_t)n_cells * tile);
	pc.acc.ensure((size_t)n_cells * tile);
	pc.aux.ensure((size_t)3 * tile);
	pc.geno_t.ensure((size_t)m.n_snp * n_samp);
	const bool dedup = predict_dedup_enabled();
	const int n_tiles = (n_samp + tile - 1) / tile;
	int table_size = 1024;
	while (table_size < 2 * tile) table_size <<= 1;
	std::unique_ptr<PredictCache::Pending> pd;
	if (dedup)
	{
		pc.dd_table.ensure((size_t)table_size);
		pc.dd_repof.ensure(tile); pc.dd_uid.ensure(tile); pc.dd_rep.ensure(tile);
		pc.dd_norm.ensure((size_t)2 * tile);
		pc.dd_count.ensure((size_t)n_tiles * nc + 1);
		collect_pending(m, pc, false);
		pd.reset(new PredictCache::Pending());
		pd->host.ensure((size_t)n_tiles * nc + 1);
	}
	const double *tbl = device_rare_freq_table();
	const int *snp_w = snp_weight_dev ? snp_weight_dev : pc.d_snp_weight.get();
	hibag_b200_predict_stats &ps = m.predict_stats;

	HB_CUDA(cudaEventRecord(pc.ev_begin.e, st));
	// sample-major -> SNP-major once, so the per-classifier gathers are coalesced
	launch_transpose_i8(geno_dev, n_samp, m.n_snp, pc.geno_t.get(), st);
	ps.kernel_launches++;

	size_t n_cell_ev = 0;
	int tile_idx = -1;
	for (int begin = 0; begin < n_samp; begin += tile)
	{
		const int nt = std::min(tile, n_samp - begin);
		tile_idx++;
		HB_CUDA(cudaMemsetAsync(pc.acc.get(), 0, sizeof(double) * (size_t)n_cells * tile, st));
		HB_CUDA(cudaMemsetAsync(pc.aux.get(), 0, sizeof(double) * 3 * (size_t)tile, st));
		HB_CUDA(cudaMemsetAsync(pc.counters.get(), 0, sizeof(unsigned int) * (size_t)(nc + 1), st));
		for (int c = 0; c < nc; c++)
		{
			const Classifier &cl = m.cls[c];
			const int n_snp_c = (int)cl.snpidx.size();
			launch_pack_classifier(pc.geno_t.get(), (size_t)n_samp, begin, nt,
				pc.d_snpidx.get() + pc.snp_off[c], n_snp_c, snp_w, pc.s1.get(), pc.s2.get(),
				tile, pc.weight.get(), st);
			CellPass p;
			memset(&p, 0, sizeof(p));
			bind_list(pc.blobs[c], pc.d_blobs.get() + pc.blob_off[c], tbl, p);
			p.s1 = pc.s1.get(); p.s2 = pc.s2.get(); p.geno_stride = tile;
			p.samp_list = nullptr; p.n_pos = nt;
			int *n_unique = nullptr;
			if (dedup)
			{
				// distinct genotypes of the tile at this classifier's SNPs: the kernel scores position
				// u < *n_unique = sample rep[u]; every other sample shares a column of the cell matrix
				n_unique = pc.dd_count.get() + (size_t)tile_idx * nc + c;
				launch_dedup_genotypes(pc.s1.get(), pc.s2.get(), tile, geno_words(n_snp_c), nt,
					pc.dd_table.get(), table_size, pc.dd_repof.get(), pc.dd_uid.get(), pc.dd_rep.get(),
					n_unique, st);
				p.samp_list = pc.dd_rep.get();
				p.n_pos_dev = n_unique;
				pd->pairs_per_sample.push_back(pc.blobs[c].pairs_per_sample);
				pd->nw.push_back(0);
				ps.kernel_launches += 2;
			}
			p.task_counter = pc.counters.get() + c;
			p.P = pc.P.get(); p.p_stride = (size_t)tile;
			const int R = choose_samples_per_lane(nt, pc.blobs[c].n_chunks, n_snp_c, di.sm_count);
			// CUDA-event bracket around every launch of the dominant kernel (roofline input)
			while (pc.cell_ev.size() < n_cell_ev + 2) pc.cell_ev.emplace_back(new Event());
			HB_CUDA(cudaEventRecord(pc.cell_ev[n_cell_ev]->e, st));
			const int nw = launch_cell_pass(p, R, di.sm_count, st);
			HB_CUDA(cudaEventRecord(pc.cell_ev[n_cell_ev + 1]->e, st));
			n_cell_ev += 2;
			if (dedup)
			{
				launch_predict_accumulate_dedup(pc.P.get(), (size_t)tile, n_cells, nt, pc.dd_uid.get(),
					n_unique, pc.dd_norm.get(), (size_t)tile, pc.weight.get(), pc.acc.get(), (size_t)tile,
					pc.aux.get(), st);
				pd->nw.back() = nw;
				ps.kernel_launches++;
			} else {
				launch_predict_accumulate(pc.P.get(), (size_t)tile, n_cells, nt, pc.weight.get(),
					pc.acc.get(), (size_t)tile, pc.aux.get(), st);
				ps.pair_evals += pc.blobs[c].pairs_per_sample * (uint64_t)nt;
				ps.popc32_issued += pc.blobs[c].pairs_per_sample * (uint64_t)nt * (uint64_t)nw;
				ps.positions_scored += (uint64_t)nt;
			}
			ps.kernel_launches += 3; ps.cell_kernel_launches++;
			ps.pair_evals_nominal += pc.blobs[c].pairs_per_sample * (uint64_t)nt;
			ps.positions_total += (uint64_t)nt;
		}
		if (partial_dev)
		{
			launch_export_partial(pc.acc.get(), (size_t)tile, pc.aux.get(), n_cells, begin, nt,
				partial_dev, st);
			ps.kernel_launches++;
		} else {
			launch_predict_finalize(pc.acc.get(), (size_t)tile, pc.aux.get(), n_hla, begin, nt,
				out.h1, out.h2, out.max_prob, out.matching, out.dosage, out.post_prob, st);
			ps.kernel_launches += out.post_prob ? 2 : 1;
		}
	}
	if (dedup)
	{
		// the counts of this call come back behind its kernels; they are folded into the statistics
		// once they have arrived (at once for a synchronous call)
		const size_t n_cnt = pd->pairs_per_sample.size();
		if (tile_idx + 1 != n_tiles || n_cnt != (size_t)n_tiles * nc)
			throw std::runtime_error("predict: internal count mismatch");
		HB_CUDA(cudaMemcpyAsync(pd->host.get(), pc.dd_count.get(), sizeof(int) * n_cnt,
			cudaMemcpyDeviceToHost, st));
		HB_CUDA(cudaEventRecord(pd->done.e, st));
		pc.pending.emplace_back(std::move(pd));
	}
	HB_CUDA(cudaEventRecord(pc.ev_end.e, st));
	if (sync)
	{
		HB_CUDA(cudaStreamSynchronize(st));
		collect_pending(m, pc, true);
		float ms = 0;
		HB_CUDA(cudaEventElapsedTime(&ms, pc.ev_begin.e, pc.ev_end.e));
		ps.gpu_kernel_ms += ms;
		for (size_t k = 0; k + 1 < n_cell_ev; k += 2)
		{
			HB_CUDA(cudaEventElapsedTime(&ms, pc.cell_ev[k]->e, pc.cell_ev[k + 1]->e));
			ps.cell_kernel_ms += ms;
		}
	}
}

void predict_host(hibag_b200_model &m, const int8_t *geno, int n_samp,
	const hibag_b200_predict_out &out)
{
	if (n_samp <= 0) return;
	current_device();
	const bool dbg = getenv("HIBAG_B200_PREDICT_DEBUG") != nullptr;
	const auto t_begin = std::chrono::steady_clock::now();
	auto since = [&]() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count(); };
	const int n_hla = m.n_hla;
	const size_t n_cells = (size_t)n_hla * (n_hla + 1) / 2;
	// Chunks of one tile: chunk c+1 is enqueued on the compute stream before the outputs of
	// chunk c are copied back on the copy stream, so the device-to-host traffic (1 GB of
	// posterior rows at 200,000 samples) hides behind the scoring of the next chunk.
	Stream st, st_copy;
	const int chunk = pick_tile(n_samp, (int)n_cells);
	const int n_chunks = (n_samp + chunk - 1) / chunk;
	DevBuf<int8_t> d_geno;
	d_geno.ensure((size_t)n_samp * m.n_snp);
	DevBuf<int> d_h1, d_h2;
	DevBuf<double> d_mp, d_mt, d_ds, d_pp;
	hibag_b200_predict_out dev;
	memset(&dev, 0, sizeof(dev));
	if (out.h1) dev.h1 = d_h1.ensure(n_samp);
	if (out.h2) dev.h2 = d_h2.ensure(n_samp);
	if (out.max_prob) dev.max_prob = d_mp.ensure(n_samp);
	if (out.matching) dev.matching = d_mt.ensure(n_samp);
	if (out.dosage) dev.dosage = d_ds.ensure((size_t)n_samp * n_hla);
	if (out.post_prob) dev.post_prob = d_pp.ensure((size_t)n_samp * n_cells);
	std::vector<std::unique_ptr<Event> > done;
	for (int c = 0; c < n_chunks; c++) done.emplace_back(new Event(false));
	size_t d2h = 0;
	auto enqueue = [&](int c) {
		const int b = c * chunk, n = std::min(chunk, n_samp - b);
		HB_CUDA(cudaMemcpyAsync(d_geno.get() + (size_t)b * m.n_snp, geno + (size_t)b * m.n_snp,
			(size_t)n * m.n_snp, cudaMemcpyHostToDevice, st.s));
		hibag_b200_predict_out o = dev;
		if (o.h1) o.h1 += b;
		if (o.h2) o.h2 += b;
		if (o.max_prob) o.max_prob += b;
		if (o.matching) o.matching += b;
		if (o.dosage) o.dosage += (size_t)b * n_hla;
		if (o.post_prob) o.post_prob += (size_t)b * n_cells;
		predict_device(m, d_geno.get() + (size_t)b * m.n_snp, n, o, nullptr, nullptr, st.s, false);
		HB_CUDA(cudaEventRecord(done[c]->e, st.s));
	};
	auto fetch = [&](int c) {
		const int b = c * chunk, n = std::min(chunk, n_samp - b);
		HB_CUDA(cudaStreamWaitEvent(st_copy.s, done[c]->e, 0));
		auto back = [&](void *dst, const void *src, size_t bytes) {
			if (!dst) return;
			HB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, st_copy.s));
			d2h += bytes;
		};
		back(out.h1 ? out.h1 + b : nullptr, dev.h1 + b, sizeof(int) * (size_t)n);
		back(out.h2 ? out.h2 + b : nullptr, dev.h2 + b, sizeof(int) * (size_t)n);
		back(out.max_prob ? out.max_prob + b : nullptr, dev.max_prob + b, sizeof(double) * (size_t)n);
		back(out.matching ? out.matching + b : nullptr, dev.matching + b, sizeof(double) * (size_t)n);
		back(out.dosage ? out.dosage + (size_t)b * n_hla : nullptr, dev.dosage + (size_t)b * n_hla,
			sizeof(double) * (size_t)n * n_hla);
		back(out.post_prob ? out.post_prob + (size_t)b * n_cells : nullptr,
			dev.post_prob + (size_t)b * n_cells, sizeof(double) * (size_t)n * n_cells);
	};
	const double t_alloc = since();
	enqueue(0);
	for (int c = 0; c < n_chunks; c++)
	{
		if (c + 1 < n_chunks) enqueue(c + 1);
		fetch(c);
	}
	const double t_enq = since();
	HB_CUDA(cudaStreamSynchronize(st.s));
	const double t_comp = since();
	HB_CUDA(cudaStreamSynchronize(st_copy.s));
	if (dbg)
		fprintf(stderr, "predict_host: %d samples, %d chunk(s): buffers %.1f ms, enqueued at %.1f, computed at %.1f, "
			"copied back at %.1f\n", n_samp, n_chunks, t_alloc, t_enq, t_comp, since());
	predict_collect_stats(m);
	m.predict_stats.h2d_bytes += (size_t)n_samp * m.n_snp;
	m.predict_stats.d2h_bytes += d2h;
}

}  // namespace hb
