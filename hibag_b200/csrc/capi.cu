// capi.cu -- the extern "C" surface declared in include/hibag_b200.h. No C++ exception or
// CUDA error crosses this boundary: everything is caught and turned into a return code plus
// hibag_b200_last_error().
#include <cstring>
#include <string>

#include <cstdlib>
#include <map>
#include <mutex>
#include <vector>

#include "em.h"
#include "model.h"

namespace {

thread_local std::string g_last_error;

template <typename F>
int guarded(F &&f)
{
	try
	{
		f();
		return 0;
	} catch (std::exception &e)
	{
		g_last_error = e.what();
	} catch (const char *e)
	{
		g_last_error = e;
	} catch (...)
	{
		g_last_error = "unknown error";
	}
	return -1;
}

void require(bool ok, const char *msg)
{
	if (!ok) throw std::runtime_error(msg);
}

}  // namespace

extern "C" {

const char *hibag_b200_version(void) { return "hibag_b200 0.1 (sm_100a)"; }
const char *hibag_b200_last_error(void) { return g_last_error.c_str(); }

int hibag_b200_device_count(void)
{
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
	return n;
}

int hibag_b200_set_device(int device)
{
	return guarded([&]() { hb::select_device(device); });
}

int hibag_b200_device_info(char *name, int name_len, int *sm_count, int *clock_khz)
{
	return guarded([&]() {
		const hb::DeviceInfo &di = hb::current_device();
		if (name && name_len > 0) { strncpy(name, di.name, name_len - 1); name[name_len - 1] = 0; }
		if (sm_count) *sm_count = di.sm_count;
		if (clock_khz) *clock_khz = di.clock_khz;
	});
}

hibag_gpu_ext_proc *hibag_b200_get_procs(void) { return hb::plugin_procs(); }
hibag_gpu_ext_proc *hibag_b200_get_procs_ex(int with_haplomatch)
{
	return with_haplomatch ? hb::plugin_procs_with_haplomatch() : hb::plugin_procs();
}

int hibag_b200_haplomatch(const hibag_haplotype *haplo, const size_t *n_haplo, int n_hla, int n_snp,
	const hibag_genotype *geno, int n_samp, uint32_t **out_buf, size_t *out_n)
{
	return guarded([&]() {
		require(haplo && n_haplo && geno && out_buf && out_n, "haplomatch: null argument");
		std::vector<int> ib;
		for (int i = 0; i < n_samp; i++) if (geno[i].bootstrap_count > 0) ib.push_back(i);
		*out_buf = hb::haplomatch_records(haplo, n_haplo, n_hla, n_snp, geno, n_samp, ib, out_n);
	});
}

void hibag_b200_free(void *p) { free(p); }

int hibag_b200_best_guess(const hibag_haplotype *haplo, int n_haplo, int n_hla, int n_snp,
	const hibag_genotype *geno, int n_geno, int32_t *out_a1, int32_t *out_a2)
{
	return guarded([&]() {
		require(haplo && geno && out_a1 && out_a2, "best_guess: null argument");
		hb::score_host_arrays(0, haplo, n_haplo, n_hla, n_snp, geno, n_geno, out_a1, out_a2,
			nullptr, nullptr);
	});
}

int hibag_b200_post_prob(const hibag_haplotype *haplo, int n_haplo, int n_hla, int n_snp,
	const hibag_genotype *geno, int n_geno, double *out)
{
	return guarded([&]() {
		require(haplo && geno && out, "post_prob: null argument");
		hb::score_host_arrays(1, haplo, n_haplo, n_hla, n_snp, geno, n_geno, nullptr, nullptr,
			out, nullptr);
	});
}

int hibag_b200_post_prob2(const hibag_haplotype *haplo, int n_haplo, int n_hla, int n_snp,
	const hibag_genotype *geno, int n_geno, double *out_prob, double *out_sum)
{
	return guarded([&]() {
		require(haplo && geno && out_prob && out_sum, "post_prob2: null argument");
		hb::score_host_arrays(2, haplo, n_haplo, n_hla, n_snp, geno, n_geno, nullptr, nullptr,
			out_prob, out_sum);
	});
}

hibag_b200_model *hibag_b200_model_new(int n_snp, int n_hla)
{
	hibag_b200_model *m = nullptr;
	guarded([&]() {
		require(n_snp > 0 && n_hla > 0, "model_new: n_snp and n_hla must be positive");
		m = new hibag_b200_model();
		m->n_snp = n_snp; m->n_hla = n_hla;
	});
	return m;
}

void hibag_b200_model_free(hibag_b200_model *m) { delete m; }

int hibag_b200_model_set_training(hibag_b200_model *m, int n_samp, const int8_t *geno,
	const int32_t *h1, const int32_t *h2)
{
	return guarded([&]() {
		require(m && geno && h1 && h2 && n_samp > 0, "set_training: invalid argument");
		m->tsession.reset();
		m->n_samp = n_samp;
		m->geno_t.resize((size_t)m->n_snp * n_samp);
		for (int s = 0; s < n_samp; s++)
			for (int k = 0; k < m->n_snp; k++)
				m->geno_t[(size_t)k * n_samp + s] = geno[(size_t)s * m->n_snp + k];
		m->h1.assign(h1, h1 + n_samp);
		m->h2.assign(h2, h2 + n_samp);
	});
}

int hibag_b200_model_train(hibag_b200_model *m, const hibag_b200_train_opts *opts)
{
	return guarded([&]() {
		require(m && opts, "train: null argument");
		require(opts->nclassifier >= 0, "train: nclassifier must be >= 0");
		hb::train_model(*m, *opts);
	});
}

int hibag_b200_model_train_stats(const hibag_b200_model *m, hibag_b200_train_stats *out)
{
	return guarded([&]() { require(m && out, "null argument"); *out = m->train_stats; });
}

int hibag_b200_model_train_trace(const hibag_b200_model *m, int64_t *out, int max_rows)
{
	if (!m) return -1;
	const int rows = (int)(m->train_trace.size() / 4);
	if (out)
		for (int i = 0; i < rows && i < max_rows; i++)
			for (int j = 0; j < 4; j++) out[4 * i + j] = m->train_trace[4 * (size_t)i + j];
	return rows;
}

int hibag_b200_model_num_classifiers(const hibag_b200_model *m)
{
	return m ? (int)m->cls.size() : -1;
}

int hibag_b200_model_clear(hibag_b200_model *m)
{
	return guarded([&]() {
		require(m != nullptr, "null argument");
		m->cls.clear(); m->pcache.reset(); m->train_trace.clear();
		m->rng_seeded = false;
		memset(&m->train_stats, 0, sizeof(m->train_stats));
		memset(&m->predict_stats, 0, sizeof(m->predict_stats));
	});
}

int hibag_b200_model_classifier_info(const hibag_b200_model *m, int k, int *n_snp,
	int *n_haplo, double *oob_acc)
{
	return guarded([&]() {
		require(m && k >= 0 && k < (int)m->cls.size(), "classifier index out of range");
		const hb::Classifier &c = m->cls[k];
		if (n_snp) *n_snp = (int)c.snpidx.size();
		if (n_haplo) *n_haplo = (int)c.haplo.h.size();
		if (oob_acc) *oob_acc = c.oob_acc;
	});
}

int hibag_b200_sm_time(uint64_t *out, int reset)
{
	return guarded([&]() {
		require(out != nullptr, "null argument");
		unsigned long long *d = hb::device_sm_acct();
		HB_CUDA(cudaDeviceSynchronize());
		unsigned long long h[hb::SM_ACCT_N];
		HB_CUDA(cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost));
		for (int k = 0; k < hb::SM_ACCT_N; k++) out[k] = (uint64_t)h[k];
		if (reset) HB_CUDA(cudaMemset(d, 0, sizeof(h)));
	});
}

size_t hibag_b200_trim_cache(void)
{
	try { return hb::pool_trim(); } catch (...) { return 0; }
}

namespace {
std::mutex g_host_mu;
std::map<void *, size_t> g_host_blocks;          // blocks handed out by hibag_b200_host_alloc
}

void *hibag_b200_host_alloc(size_t bytes)
{
	void *p = nullptr;
	const int rc = guarded([&]() {
		hb::current_device();
		size_t got = 0;
		p = hb::pool_alloc(true, bytes ? bytes : 1, &got);
		std::lock_guard<std::mutex> lk(g_host_mu);
		g_host_blocks[p] = got;
	});
	return rc == 0 ? p : nullptr;
}

void hibag_b200_host_free(void *p)
{
	if (!p) return;
	size_t bytes = 0;
	{
		std::lock_guard<std::mutex> lk(g_host_mu);
		auto it = g_host_blocks.find(p);
		if (it == g_host_blocks.end()) return;      // not ours: ignore
		bytes = it->second;
		g_host_blocks.erase(it);
	}
	try { hb::pool_free(true, p, bytes); } catch (...) {}
}

int hibag_b200_model_classifier_samp_num_len(const hibag_b200_model *m, int k)
{
	if (!m || k < 0 || k >= (int)m->cls.size()) return -1;
	return (int)m->cls[k].samp_num.size();
}

int hibag_b200_model_classifier_get(const hibag_b200_model *m, int k, int32_t *snpidx,
	int32_t *samp_num, double *freq, int32_t *hla, uint64_t *packed)
{
	return guarded([&]() {
		require(m && k >= 0 && k < (int)m->cls.size(), "classifier index out of range");
		const hb::Classifier &c = m->cls[k];
		const int ns = (int)c.snpidx.size();
		if (snpidx) for (int i = 0; i < ns; i++) snpidx[i] = c.snpidx[i];
		if (samp_num) for (size_t i = 0; i < c.samp_num.size(); i++) samp_num[i] = c.samp_num[i];
		size_t idx = 0;
		for (size_t a = 0; a < c.haplo.len.size(); a++)
			for (int n = c.haplo.len[a]; n > 0; n--, idx++)
			{
				const hibag_haplotype &h = c.haplo.h[idx];
				if (freq) freq[idx] = h.freq;
				if (hla) hla[idx] = (int32_t)a;
				if (packed)
				{
					uint64_t w0 = (uint64_t)h.packed[0], w1 = (uint64_t)h.packed[1];
					if (ns < 64) { w0 &= ns ? ((~0ULL) >> (64 - ns)) : 0ULL; w1 = 0; }
					else if (ns < 128) { w1 &= (ns > 64) ? ((~0ULL) >> (128 - ns)) : 0ULL; }
					packed[2 * idx] = w0; packed[2 * idx + 1] = w1;
				}
			}
	});
}

int hibag_b200_model_add_classifier(hibag_b200_model *m, int n_snp, const int32_t *snpidx,
	const int32_t *samp_num, int n_samp, int n_haplo, const double *freq,
	const int32_t *hla, const uint64_t *packed, double oob_acc)
{
	return guarded([&]() {
		require(m && n_snp >= 0 && n_snp <= HIBAG_B200_MAX_SNP && n_haplo >= 0, "add_classifier: invalid sizes");
		require((n_snp == 0 || snpidx) && (n_haplo == 0 || (freq && hla && packed)), "add_classifier: null argument");
		hb::Classifier c;
		c.snpidx.assign(snpidx, snpidx + n_snp);
		for (int k : c.snpidx) require(k >= 0 && k < m->n_snp, "add_classifier: SNP index out of range");
		if (samp_num && n_samp > 0) c.samp_num.assign(samp_num, samp_num + n_samp);
		c.haplo.n_snp = n_snp;
		c.haplo.len.assign(m->n_hla, 0);
		c.haplo.h.resize(n_haplo);
		int prev = 0;
		for (int i = 0; i < n_haplo; i++)
		{
			require(hla[i] >= prev && hla[i] < m->n_hla, "add_classifier: haplotypes must be grouped by allele");
			prev = hla[i];
			hibag_haplotype &h = c.haplo.h[i];
			h.packed[0] = (int64_t)packed[2 * i]; h.packed[1] = (int64_t)packed[2 * i + 1];
			h.freq = freq[i];
			c.haplo.len[hla[i]]++;
		}
		c.haplo.set_tags();
		c.oob_acc = oob_acc;
		m->cls.push_back(c);
		m->pcache.reset();
	});
}

int hibag_b200_model_predict(hibag_b200_model *m, const int8_t *geno, int n_samp,
	const hibag_b200_predict_out *out)
{
	return guarded([&]() {
		require(m && geno && out && n_samp >= 0, "predict: invalid argument");
		hb::predict_host(*m, geno, n_samp, *out);
	});
}

int hibag_b200_model_predict_device(hibag_b200_model *m, const int8_t *geno_dev, int n_samp,
	const hibag_b200_predict_out *out_dev, void *cuda_stream, int sync)
{
	return guarded([&]() {
		require(m && geno_dev && out_dev && n_samp >= 0, "predict_device: invalid argument");
		hb::predict_device(*m, geno_dev, n_samp, *out_dev, nullptr, nullptr,
			(cudaStream_t)cuda_stream, sync != 0);
	});
}

int hibag_b200_model_predict_stats(const hibag_b200_model *m, hibag_b200_predict_stats *out)
{
	return guarded([&]() { require(m && out, "null argument"); *out = m->predict_stats; });
}

int hibag_b200_model_predict_partial_device(hibag_b200_model *m, const int8_t *geno_dev,
	int n_samp, const int32_t *snp_weight_dev, double *acc_dev, void *cuda_stream, int sync)
{
	return guarded([&]() {
		require(m && geno_dev && acc_dev && n_samp >= 0, "predict_partial_device: invalid argument");
		hibag_b200_predict_out none;
		memset(&none, 0, sizeof(none));
		hb::predict_device(*m, geno_dev, n_samp, none, snp_weight_dev, acc_dev,
			(cudaStream_t)cuda_stream, sync != 0);
	});
}

int hibag_b200_predict_finalize_device(int n_hla, int n_samp, const double *acc_dev,
	const hibag_b200_predict_out *out_dev, void *cuda_stream, int sync)
{
	return guarded([&]() {
		require(acc_dev && out_dev && n_hla > 0 && n_samp >= 0, "predict_finalize_device: invalid argument");
		hb::current_device();
		hb::launch_finalize_from_partial(acc_dev, n_hla, n_samp, out_dev->h1, out_dev->h2,
			out_dev->max_prob, out_dev->matching, out_dev->dosage, out_dev->post_prob,
			(cudaStream_t)cuda_stream);
		if (sync) HB_CUDA(cudaStreamSynchronize((cudaStream_t)cuda_stream));
	});
}

int hibag_b200_model_snp_weights(const hibag_b200_model *m, int32_t *out_weight)
{
	return guarded([&]() {
		require(m && out_weight, "null argument");
		std::vector<int> w;
		hb::snp_weights(*m, w);
		for (int i = 0; i < m->n_snp; i++) out_weight[i] = w[i];
	});
}

int hibag_b200_host_unif_rand(uint32_t seed, int n, double *out)
{
	return guarded([&]() {
		require(out != nullptr && n >= 0, "invalid argument");
		hb::RRng rng;
		rng.set_seed(seed);
		for (int i = 0; i < n; i++) out[i] = rng.unif_rand();
	});
}

int hibag_b200_bed_decode(const uint8_t *bed_file, size_t n_bytes, int n_samp, int n_snp,
	const int32_t *snp_flag, int8_t *out, int *n_save, double *kernel_ms)
{
	return guarded([&]() {
		require(bed_file != nullptr && n_samp > 0 && n_snp > 0, "invalid argument");
		if (n_bytes < 3 || bed_file[0] != 0x6C || bed_file[1] != 0x1B)
			throw std::runtime_error("Invalid prefix in the PLINK BED file.");
		const int mode = bed_file[2];
		const size_t bps = (mode == 0) ? ((size_t)n_snp + 3) / 4 : ((size_t)n_samp + 3) / 4;
		const size_t payload = bps * (size_t)(mode == 0 ? n_samp : n_snp);
		require(n_bytes >= 3 + payload, "the PLINK BED file is shorter than n_samp x n_snp genotypes");
		std::vector<int32_t> sel;
		for (int s = 0; s < n_snp; s++)
			if (snp_flag == nullptr || snp_flag[s]) sel.push_back(s);
		if (n_save) *n_save = (int)sel.size();
		if (out == nullptr || sel.empty()) return;
		hb::current_device();
		hb::Stream st;
		hb::DevBuf<uint8_t> d_in;
		hb::DevBuf<int32_t> d_sel;
		hb::DevBuf<int8_t> d_out;
		const size_t n_out = (size_t)n_samp * sel.size();
		d_in.ensure(payload); d_sel.ensure(sel.size()); d_out.ensure(n_out);
		HB_CUDA(cudaMemcpyAsync(d_in.get(), bed_file + 3, payload, cudaMemcpyHostToDevice, st.s));
		HB_CUDA(cudaMemcpyAsync(d_sel.get(), sel.data(), sizeof(int32_t) * sel.size(), cudaMemcpyHostToDevice, st.s));
		hb::Event e0, e1;
		HB_CUDA(cudaEventRecord(e0.e, st.s));
		hb::launch_bed_decode(d_in.get(), mode, n_samp, n_snp, snp_flag ? d_sel.get() : nullptr,
			(int)sel.size(), d_out.get(), st.s);
		HB_CUDA(cudaEventRecord(e1.e, st.s));
		HB_CUDA(cudaMemcpyAsync(out, d_out.get(), n_out, cudaMemcpyDeviceToHost, st.s));
		HB_CUDA(cudaStreamSynchronize(st.s));
		if (kernel_ms)
		{
			float ms = 0;
			HB_CUDA(cudaEventElapsedTime(&ms, e0.e, e1.e));
			*kernel_ms = ms;
		}
	});
}

int hibag_b200_bed_decode_device(const uint8_t *payload_dev, int mode, int n_samp, int n_snp,
	const int32_t *sel_dev, int n_save, int8_t *out_dev, void *cuda_stream)
{
	return guarded([&]() {
		require(payload_dev != nullptr && out_dev != nullptr && n_samp > 0 && n_snp > 0 && n_save > 0 &&
			(sel_dev != nullptr || n_save == n_snp), "invalid argument");
		hb::current_device();
		hb::launch_bed_decode(payload_dev, mode, n_samp, n_snp, sel_dev, n_save, out_dev, (cudaStream_t)cuda_stream);
	});
}

int hibag_b200_host_screen_constants(double *table, double *floor_table, double *bound_factor)
{
	return guarded([&]() {
		require(table != nullptr && floor_table != nullptr && bound_factor != nullptr, "invalid argument");
		const size_t n = 2 * HIBAG_B200_MAX_SNP + 1;
		memcpy(table, hb::host_rare_freq_table(), sizeof(double) * n);
		memcpy(floor_table, hb::host_rare_freq_floor_table(), sizeof(double) * n);
		*bound_factor = hb::screen_bound_factor();
	});
}

int hibag_b200_host_build_tasks(const hibag_haplotype *haplo, int n_haplo, int n_hla, int n_snp,
	int target_chunks, int32_t *out_cells, int32_t *out_chunks, int *n_chunks, uint64_t *pairs)
{
	return guarded([&]() {
		require(haplo && out_cells && out_chunks && n_chunks && pairs, "null argument");
		std::vector<unsigned char> buf(hb::list_blob_capacity(n_haplo, n_snp, n_hla) + 32);
		unsigned char *base = (unsigned char *)(((uintptr_t)buf.data() + 15) & ~(uintptr_t)15);
		hb::ListBlob b = hb::build_list_blob(haplo, n_haplo, n_hla, n_snp, base, target_chunks);
		memcpy(out_cells, base + b.off_cells, sizeof(hb::CellTask) * (size_t)b.n_cells);
		memcpy(out_chunks, base + b.off_chunks, sizeof(hb::Chunk) * (size_t)b.n_chunks);
		*n_chunks = b.n_chunks;
		*pairs = b.pairs_per_sample;
	});
}

int hibag_b200_pipe_peak(int which, double *out_ops_per_s, double *out_ms)
{
	return guarded([&]() {
		require(out_ops_per_s != nullptr, "null argument");
		const hb::DeviceInfo &di = hb::current_device();
		*out_ops_per_s = hb::run_pipe_peak(which, di.sm_count, out_ms);
	});
}

}  // extern "C"
