// scorer.cu -- see scorer.h
#include "scorer.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

namespace hb {

// ---- device selection -----------------------------------------------------------------------
static std::mutex g_dev_mutex;
static int g_device = 0;
static std::map<int, DeviceInfo> g_dev_info;
static std::map<int, double *> g_dev_table;

void select_device(int device)
{
	int n = 0;
	cudaError_t e = cudaGetDeviceCount(&n);
	if (e != cudaSuccess || n <= 0)
		throw std::runtime_error("hibag_b200: no usable CUDA device (this library has no CPU fallback)");
	if (device < 0 || device >= n)
		throw std::runtime_error("hibag_b200: invalid device index");
	HB_CUDA(cudaSetDevice(device));
	std::lock_guard<std::mutex> lk(g_dev_mutex);
	g_device = device;
}

const DeviceInfo &current_device()
{
	std::lock_guard<std::mutex> lk(g_dev_mutex);
	auto it = g_dev_info.find(g_device);
	if (it != g_dev_info.end())
	{
		cudaSetDevice(g_device);      // device is per-thread state in the runtime API
		return it->second;
	}
	int n = 0;
	cudaError_t e = cudaGetDeviceCount(&n);
	if (e != cudaSuccess || n <= 0)
		throw std::runtime_error("hibag_b200: no usable CUDA device (this library has no CPU fallback)");
	HB_CUDA(cudaSetDevice(g_device));
	cudaDeviceProp prop;
	HB_CUDA(cudaGetDeviceProperties(&prop, g_device));
	DeviceInfo di;
	di.device = g_device;
	di.sm_count = prop.multiProcessorCount;
	int khz = 0;
	cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, g_device);
	di.clock_khz = khz;
	strncpy(di.name, prop.name, sizeof(di.name) - 1);
	di.name[sizeof(di.name) - 1] = 0;
	g_dev_info[g_device] = di;
	return g_dev_info[g_device];
}

const double *device_rare_freq_table()
{
	const DeviceInfo &di = current_device();
	std::lock_guard<std::mutex> lk(g_dev_mutex);
	auto it = g_dev_table.find(di.device);
	if (it != g_dev_table.end()) return it->second;
	double *d = nullptr;
	const size_t bytes = sizeof(double) * (2 * HIBAG_B200_MAX_SNP + 1);
	HB_CUDA(cudaMalloc((void **)&d, bytes));
	HB_CUDA(cudaMemcpy(d, host_rare_freq_table(), bytes, cudaMemcpyHostToDevice));
	g_dev_table[di.device] = d;
	return d;
}

static std::map<int, double *> g_dev_floor_table;

const double *device_rare_freq_floor_table()
{
	const DeviceInfo &di = current_device();
	std::lock_guard<std::mutex> lk(g_dev_mutex);
	auto it = g_dev_floor_table.find(di.device);
	if (it != g_dev_floor_table.end()) return it->second;
	double *d = nullptr;
	const size_t bytes = sizeof(double) * (2 * HIBAG_B200_MAX_SNP + 1);
	HB_CUDA(cudaMalloc((void **)&d, bytes));
	HB_CUDA(cudaMemcpy(d, host_rare_freq_floor_table(), bytes, cudaMemcpyHostToDevice));
	g_dev_floor_table[di.device] = d;
	return d;
}

static std::map<int, unsigned long long *> g_dev_acct;

unsigned long long *device_sm_acct()
{
	const DeviceInfo &di = current_device();
	std::lock_guard<std::mutex> lk(g_dev_mutex);
	auto it = g_dev_acct.find(di.device);
	if (it != g_dev_acct.end()) return it->second;
	unsigned long long *d = nullptr;
	HB_CUDA(cudaMalloc((void **)&d, sizeof(unsigned long long) * SM_ACCT_N));
	HB_CUDA(cudaMemset(d, 0, sizeof(unsigned long long) * SM_ACCT_N));
	g_dev_acct[di.device] = d;
	return d;
}

int choose_samples_per_lane(int n_pos, int n_chunks, int n_snp, int sm_count)
{
	// enough (sample group, chunk) tasks to give every SM sub-partition several warps;
	// more samples per lane amortise the warp-uniform haplotype fetch and frequency product
	const long long want = (long long)sm_count * 4 * 12;
	int R = 4;
	if (geno_words(n_snp) == 4) R = 2;
	while (R > 1)
	{
		const long long groups = (n_pos + 32 * R - 1) / (32 * R);
		if (groups * n_chunks >= want && n_pos >= 32 * R * 8) break;
		R >>= 1;
	}
	return R;
}

// ---- block cache ----------------------------------------------------------------------------------
namespace {
struct BlockPool
{
	std::mutex mu;
	std::multimap<size_t, void *> free_blocks;      // size -> block
	size_t cached_bytes = 0;
};
BlockPool g_pool[2];                                 // [0] device (per current device), [1] pinned
std::map<void *, int> g_block_device;                // device blocks: which device owns them
const size_t POOL_LIMIT[2] = { (size_t)48 << 30, (size_t)24 << 30 };  // of 180 GB HBM / host RAM
}

void *pool_alloc(bool pinned, size_t bytes, size_t *got_bytes)
{
	// size classes: next multiple of 1/8 of the leading power of two (<= 12.5 % slack), >= 512 B
	size_t cls = 512;
	while (cls < bytes) cls <<= 1;
	if (cls > 4096)
	{
		const size_t step = cls >> 4;
		cls = (bytes + step - 1) / step * step;
	}
	BlockPool &bp = g_pool[pinned ? 1 : 0];
	int dev = 0;
	if (!pinned) cudaGetDevice(&dev);
	{
		std::lock_guard<std::mutex> lk(bp.mu);
		auto range = bp.free_blocks.equal_range(cls);
		for (auto it = range.first; it != range.second; ++it)
		{
			if (!pinned && g_block_device[it->second] != dev) continue;
			void *p = it->second;
			bp.free_blocks.erase(it);
			bp.cached_bytes -= cls;
			*got_bytes = cls;
			return p;
		}
	}
	void *p = nullptr;
	cudaError_t e = pinned ? cudaMallocHost(&p, cls) : cudaMalloc(&p, cls);
	if (e != cudaSuccess)
	{
		// give the cache back and retry once
		std::vector<void *> drop;
		{
			std::lock_guard<std::mutex> lk(bp.mu);
			for (auto &kv : bp.free_blocks)
			{
				drop.push_back(kv.second);
				if (!pinned) g_block_device.erase(kv.second);
			}
			bp.free_blocks.clear();
			bp.cached_bytes = 0;
		}
		for (void *q : drop) { if (pinned) cudaFreeHost(q); else cudaFree(q); }
		cudaGetLastError();
		e = pinned ? cudaMallocHost(&p, cls) : cudaMalloc(&p, cls);
	}
	if (e != cudaSuccess)
		throw std::runtime_error(std::string("CUDA error: ") + cudaGetErrorString(e) + " allocating " +
			std::to_string(cls) + " bytes");
	if (!pinned)
	{
		std::lock_guard<std::mutex> lk(bp.mu);
		g_block_device[p] = dev;
	}
	*got_bytes = cls;
	return p;
}

void pool_free(bool pinned, void *p, size_t bytes)
{
	if (!p) return;
	BlockPool &bp = g_pool[pinned ? 1 : 0];
	{
		std::lock_guard<std::mutex> lk(bp.mu);
		if (bp.cached_bytes + bytes <= POOL_LIMIT[pinned ? 1 : 0])
		{
			bp.free_blocks.emplace(bytes, p);
			bp.cached_bytes += bytes;
			return;
		}
		if (!pinned) g_block_device.erase(p);
	}
	if (pinned) cudaFreeHost(p); else cudaFree(p);
}

size_t pool_trim()
{
	size_t freed = 0;
	cudaDeviceSynchronize();
	for (int k = 0; k < 2; k++)
	{
		BlockPool &bp = g_pool[k];
		std::vector<void *> drop;
		{
			std::lock_guard<std::mutex> lk(bp.mu);
			for (auto &kv : bp.free_blocks)
			{
				drop.push_back(kv.second);
				if (k == 0) g_block_device.erase(kv.second);
			}
			freed += bp.cached_bytes;
			bp.free_blocks.clear();
			bp.cached_bytes = 0;
		}
		for (void *q : drop) { if (k) cudaFreeHost(q); else cudaFree(q); }
	}
	cudaGetLastError();
	return freed;
}

// ---- EvalSlot ---------------------------------------------------------------------------------
EvalSlot::EvalSlot(bool high_priority, bool spin) : st_(high_priority), ev1_(true, !spin)
{
	current_device();
	counter_.ensure(4);
	d_count_.ensure(4);
	h_count_.ensure(4);
	*h_count_.get() = 0;
}

void EvalSlot::stage_list(const hibag_haplotype *haplo, int n_hap, int n_hla, int n_snp)
{
	const size_t cap = list_blob_capacity(n_hap, n_snp, n_hla);
	unsigned char *h = h_blob_.ensure(cap);
	blob_ = build_list_blob(haplo, n_hap, n_hla, n_snp, h);
	unsigned char *d = d_blob_.ensure(cap);
	HB_CUDA(cudaMemcpyAsync(d, h, blob_.bytes, cudaMemcpyHostToDevice, st_.s));
	stats.h2d_bytes += blob_.bytes;
	ext_blob_ = nullptr;
}

void EvalSlot::enqueue_cells(const GenoView &g, const int *pos_list, int n_pos)
{
	if (n_pos <= 0) return;
	const DeviceInfo &di = current_device();
	p_stride_ = ((size_t)n_pos + 31) & ~(size_t)31;
	P_.ensure(p_stride_ * (size_t)blob_.n_cells);
	CellPass p;
	memset(&p, 0, sizeof(p));
	bind_list(blob_, ext_blob_ ? (const void *)ext_blob_ : (const void *)d_blob_.get(), device_rare_freq_table(), p);
	p.s1 = g.s1; p.s2 = g.s2; p.geno_stride = g.stride;
	p.cand_col = g.cand_col; p.cand_bit = g.cand_bit;
	p.samp_list = pos_list; p.n_pos = n_pos;
	p.task_counter = counter_.get();
	p.P = P_.get(); p.p_stride = p_stride_;
	HB_CUDA(cudaMemsetAsync(counter_.get(), 0, sizeof(unsigned int), st_.s));
	int R = choose_samples_per_lane(n_pos, blob_.n_chunks, blob_.n_snp, di.sm_count);
	if (force_r_ > 0 && force_r_ < R) R = force_r_;
	HB_CUDA(cudaEventRecord(ev0_.e, st_.s));
	const int nw = launch_cell_pass(p, R, di.sm_count, st_.s);
	HB_CUDA(cudaEventRecord(evc_.e, st_.s));
	stats.launches++; stats.cell_launches++;
	stats.pair_evals += blob_.pairs_per_sample * (uint64_t)n_pos;
	stats.popc32 += blob_.pairs_per_sample * (uint64_t)n_pos * (uint64_t)nw;
	timing_pending_ = true;
}

void EvalSlot::enqueue_reduce_oob(const GenoView &g, const int *pos_list, int n_pos)
{
	HB_CUDA(cudaMemsetAsync(d_count_.get(), 0, sizeof(int), st_.s));
	launch_reduce_oob(P_.get(), p_stride_, blob_.n_hla, pos_list, n_pos, g.a1, g.a2,
		d_count_.get(), st_.s);
	HB_CUDA(cudaMemcpyAsync(h_count_.get(), d_count_.get(), sizeof(int), cudaMemcpyDeviceToHost, st_.s));
	stats.launches++; stats.d2h_bytes += sizeof(int);
}

void EvalSlot::enqueue_reduce_ib(const GenoView &g, const int *pos_list, int n_pos)
{
	double *d = d_ratio_.ensure(n_pos);
	double *h = h_ratio_.ensure(n_pos);
	launch_reduce_ib(P_.get(), p_stride_, blob_.n_hla, pos_list, n_pos, g.a1, g.a2, d, st_.s);
	HB_CUDA(cudaMemcpyAsync(h, d, sizeof(double) * (size_t)n_pos, cudaMemcpyDeviceToHost, st_.s));
	stats.launches++; stats.d2h_bytes += sizeof(double) * (size_t)n_pos;
}

void EvalSlot::sync()
{
	HB_CUDA(cudaEventRecord(ev1_.e, st_.s));
	HB_CUDA(cudaEventSynchronize(ev1_.e));
	if (timing_pending_)
	{
		float ms = 0;
		HB_CUDA(cudaEventElapsedTime(&ms, ev0_.e, ev1_.e));
		stats.kernel_ms += ms;
		HB_CUDA(cudaEventElapsedTime(&ms, ev0_.e, evc_.e));
		stats.cell_ms += ms;
		timing_pending_ = false;
	}
}

// ---- ScoreQueue / BatchScorer ------------------------------------------------------------------
static std::mutex g_queue_mutex;
static std::map<int, ScoreQueue *> g_queues;

ScoreQueue &ScoreQueue::get(int id)
{
	const DeviceInfo &di = current_device();
	std::lock_guard<std::mutex> lk(g_queue_mutex);
	const int key = di.device * 64 + (id & 63);
	auto it = g_queues.find(key);
	if (it != g_queues.end()) return *it->second;
	const char *pe = getenv("HIBAG_B200_SCORE_PRIO");
	ScoreQueue *q = new ScoreQueue(pe == nullptr || atoi(pe) != 0);       // lives for the process
	g_queues[key] = q;
	return *q;
}

static std::atomic<int> g_next_scorer{0};

BatchScorer::BatchScorer()
{
	current_device();
	// screened passes are chains of small, latency-bound launches that do not fill the GPU:
	// the lanes' passes are spread over a few streams so that they overlap
	int k = 6;
	if (const char *e = getenv("HIBAG_B200_SCORE_QUEUES")) k = std::max(1, std::min(16, atoi(e)));
	queue_id_ = g_next_scorer.fetch_add(1) % k;
	int lg = 70;
	if (const char *e = getenv("HIBAG_B200_SCREEN_TAU_LOG2")) lg = std::max(54, std::min(200, atoi(e)));
	screen_tau_ = std::ldexp(1.0, -lg);
}

void BatchScorer::begin_round(int n_lists, int max_hap, int n_snp, int n_hla)
{
	n_snp_ = n_snp; n_hla_ = n_hla; n_cells_ = n_hla * (n_hla + 1) / 2;
	cap_ = (list_blob_capacity(max_hap, n_snp, n_hla) + 255) & ~(size_t)255;
	h_blobs_.ensure(cap_ * (size_t)n_lists);
	d_blobs_.ensure(cap_ * (size_t)n_lists);
	blobs_.assign(n_lists, ListBlob());
	cols_.assign(n_lists, nullptr);
	counters_.ensure(2 * MAX_BATCH_LISTS);
	d_counts_.ensure(n_lists);
	h_counts_.ensure(n_lists);
	d_evals_.ensure(n_lists + 1);
	h_evals_.ensure(n_lists + 1);
	al_tab_.ensure(2 * (size_t)n_lists * n_hla);
}

void BatchScorer::upload(const std::vector<int> &which)
{
	for (int i : which)
	{
		HB_CUDA(cudaMemcpyAsync(d_blobs_.get() + (size_t)i * cap_, h_blobs_.get() + (size_t)i * cap_,
			blobs_[i].bytes, cudaMemcpyHostToDevice, st_.s));
		stats.h2d_bytes += blobs_[i].bytes;
	}
	HB_CUDA(cudaEventRecord(ev_up_.e, st_.s));
}

void BatchScorer::set_sample_sets(const std::vector<int> &oob, const std::vector<int> &ib,
	const std::vector<int> &a1, const std::vector<int> &a2, int n_hla)
{
	(void)a1; (void)a2;
	const int n_cells = n_hla * (n_hla + 1) / 2;
	screen_ = false;
	for (int kind = 0; kind < 2; kind++)
	{
		const std::vector<int> &s = kind ? ib : oob;
		const size_t stride = (s.size() + 31) & ~(size_t)31;
		if ((size_t)n_cells * stride > (size_t)0x7fffffff) return;    // positions are addressed with int
		set_samples_[kind] = s;
		std::vector<int> eoff(n_cells, 0);
		for (int c = 0; c < n_cells; c++) eoff[c] = (int)((size_t)c * stride);
		ent_off_[kind].ensure(n_cells);
		HB_CUDA(cudaMemcpyAsync(ent_off_[kind].get(), eoff.data(), sizeof(int) * n_cells, cudaMemcpyHostToDevice, st_.s));
		stream_sync_blocking(st_.s);                 // the host vector goes out of scope
		stats.h2d_bytes += sizeof(int) * (size_t)n_cells;
		evals_per_list_[kind] = 0;
	}
	device_rare_freq_floor_table();
	screen_ = true;
}

void BatchScorer::run_cells(const GenoView &g, int cand_bit, const std::vector<int> &which,
	int first, int count, const int *pos_list, int n_pos, double *P, size_t p_stride, int queue)
{
	const DeviceInfo &di = current_device();
	CellBatch b;
	memset(&b, 0, sizeof(b));
	b.table = device_rare_freq_table();
	b.s1 = g.s1; b.s2 = g.s2; b.geno_stride = g.stride;
	b.samp_list = pos_list; b.n_pos = n_pos;
	b.task_counters = counters_.get();
	b.p_stride = p_stride;
	b.n_snp = n_snp_;
	b.n_lists = count;
	b.acct = device_sm_acct();
	int total_chunks = 0;
	uint64_t pairs = 0;
	for (int k = 0; k < count; k++)
	{
		const int i = which[first + k];
		const ListBlob &lb = blobs_[i];
		const unsigned char *d = d_blobs_.get() + (size_t)i * cap_;
		ListDesc &L = b.lists[k];
		L.hap = d;
		L.cells = (const CellTask *)(d + lb.off_cells);
		L.chunks = (const Chunk *)(d + lb.off_chunks);
		L.cand_col = cols_[i];
		L.cand_bit = cand_bit;
		L.P = P + (size_t)(first + k) * n_cells_ * p_stride;
		L.n_hap = lb.n_hap; L.n_chunks = lb.n_chunks;
		b.n_dist = lb.n_dist;
		if (lb.n_hap > b.max_hap) b.max_hap = lb.n_hap;
		total_chunks += lb.n_chunks;
		pairs += lb.pairs_per_sample;
	}
	const int R = choose_samples_per_lane(n_pos, total_chunks, n_snp_, di.sm_count);
	ScoreQueue &q = ScoreQueue::get(queue);
	int nw;
	{
		std::lock_guard<std::mutex> lk(q.mu);
		HB_CUDA(cudaStreamWaitEvent(q.st.s, ev_up_.e, 0));
		HB_CUDA(cudaMemsetAsync(counters_.get(), 0, sizeof(unsigned int) * (size_t)count, q.st.s));
		HB_CUDA(cudaEventRecord(ev0_.e, q.st.s));
		nw = launch_cell_batch(b, R, di.sm_count, q.st.s);
		HB_CUDA(cudaEventRecord(ev1_.e, q.st.s));
	}
	HB_CUDA(cudaStreamWaitEvent(st_.s, ev1_.e, 0));
	// the next sub-batch (or pass) reuses the counters: keep its memset behind this kernel
	HB_CUDA(cudaEventRecord(ev_up_.e, st_.s));
	stats.launches++; stats.cell_launches++;
	stats.pair_evals += pairs * (uint64_t)n_pos;
	stats.popc32 += pairs * (uint64_t)n_pos * (uint64_t)nw;
}

/// second level of the in-bag screen: HIBAG_B200_SCREEN_REFINE=1 turns it on (read per
/// call: the tests switch it inside one process). Off by default: it removes 25-30 % of the in-bag pair
/// evaluations (8 % of the GPU's SM-time at config 2) and costs 3.5 % in its own kernel, but the
/// classifiers/min of the 40-lane step do not move beyond the run-to-run spread
/// (profiles/r02_screen_refine.txt), and it needs a second set of need lists (170 MB per lane)
static bool screen_refine()
{
	const char *e = getenv("HIBAG_B200_SCREEN_REFINE");
	return e ? atoi(e) != 0 : false;
}

/// Position classes (kernels.h ScreenArgs::rep): the positions of a list that carry the same packed
/// genotype and the same true type are screened and scored once. HIBAG_B200_TRAIN_DEDUP=0 turns it
/// off (read per call: the tests switch it inside one process); the trained model is the same bit for bit.
static bool train_dedup()
{
	const char *e = getenv("HIBAG_B200_TRAIN_DEDUP");
	return e ? atoi(e) != 0 : true;
}

/// One sub-batch of a screened pass, all on one of the device's scoring streams: per-allele
/// bounds and x_ref, the need lists, their tasks, the surviving cells (gather launch) and the
/// screened reduction. kind: 0 out-of-bag, 1 in-bag.
void BatchScorer::run_cells_screened(const GenoView &g, int cand_bit, const std::vector<int> &which,
	int first, int count, const int *pos_list, int n_pos, int kind)
{
	const DeviceInfo &di = current_device();
	const size_t n_cells = (size_t)n_cells_;
	ScreenArgs a;
	ScreenLists ls;
	GatherBatch gb;
	memset(&a, 0, sizeof(a)); memset(&ls, 0, sizeof(ls)); memset(&gb, 0, sizeof(gb));
	a.table_floor = device_rare_freq_floor_table();
	a.table = device_rare_freq_table();
	a.s1 = g.s1; a.s2 = g.s2; a.samp_list = pos_list; a.a1 = g.a1; a.a2 = g.a2;
	a.n_snp = n_snp_; a.geno_stride = g.stride; a.n_pos = n_pos; a.n_hla = n_hla_; a.n_lists = count;
	a.p_stride = p_stride_;
	a.K = screen_bound_factor();
	{
		// (read per call: an experiment switch) a tile of per-allele sums fits shared memory up to ~150 alleles
		const char *e = getenv("HIBAG_B200_SCREEN_USMEM");
		a.u_smem = ((e ? atoi(e) != 0 : true) && (size_t)n_hla_ * 128 * sizeof(double) <= 120 * 1024) ? 1 : 0;
	}
	a.acct = device_sm_acct();
	gb.acct = a.acct;
	gb.acct_cls = kind ? SM_ACCT_GATHER_IB : SM_ACCT_GATHER_OOB;
	a.tau = kind ? screen_tau_ : 1.0;
	a.U = U_.get() + (size_t)first * n_hla_ * p_stride_;
	a.xref = xref_.get() + (size_t)first * p_stride_;
	a.P = P_.get() + (size_t)first * n_cells * p_stride_;
	a.count = cnt_.get() + (size_t)first * n_cells;
	a.entries = ent_.get() + (size_t)first * n_cells * p_stride_;
	a.task_prefix = prefix_.get() + (size_t)first * (n_cells + 2);
	a.evals = d_evals_.get() + first;
	a.rescued = d_evals_.get() + which.size();
	a.al_tab = al_tab_.get() + 2 * (size_t)first * n_hla_;
	const bool refine = kind == 1 && screen_refine();
	if (refine)
	{
		a.U2 = U2_.get() + (size_t)first * n_hla_ * 4 * p_stride_;
		a.hetk = hetk_.get() + (size_t)first * p_stride_;
		a.count2 = cnt2_.get() + (size_t)first * n_cells;
		a.entries2 = ent2_.get() + (size_t)first * n_cells * p_stride_;
		a.K2 = screen_bound_factor2();
		const double *tf = host_rare_freq_floor_table();
		a.tf[0] = tf[0]; a.tf[1] = tf[1]; a.tf[2] = tf[2];
	}
	const bool dedup = train_dedup();
	int table_size = 1024;
	while (table_size < 2 * n_pos) table_size <<= 1;
	int *dd_table = nullptr, *dd_n_rep = nullptr;
	size_t dd_bytes = 0;
	if (dedup)
	{
		// [MAX_BATCH_LISTS][table_size] hash sets, then [MAX_BATCH_LISTS] counters: one memset
		dd_bytes = sizeof(int) * ((size_t)count * table_size + MAX_BATCH_LISTS);
		dd_table = dd_table_.ensure((size_t)MAX_BATCH_LISTS * table_size + MAX_BATCH_LISTS);
		dd_n_rep = dd_table + (size_t)count * table_size;
		a.rep = rep_.ensure((size_t)MAX_BATCH_LISTS * p_stride_);
		a.rep_list = rep_list_.ensure((size_t)MAX_BATCH_LISTS * p_stride_);
		a.n_rep = dd_n_rep;
		a.pos_res = pos_res_.ensure((size_t)MAX_BATCH_LISTS * p_stride_);
	}
	if (const char *e = getenv("HIBAG_B200_SCREEN_FORCE_RESCUE")) a.force_rescue = std::max(0, atoi(e));
	if (const char *e = getenv("HIBAG_B200_SCREEN_DEVICE_RESCUE")) a.device_rescue = atoi(e) != 0;
	gb.table = a.table;
	gb.s1 = g.s1; gb.s2 = g.s2; gb.samp_list = pos_list;
	gb.p_stride = p_stride_; gb.n_snp = n_snp_; gb.geno_stride = g.stride; gb.n_pos = n_pos;
	gb.n_lists = count; gb.n_cells = n_cells_;
	gb.ent_off = ent_off_[kind].get();
	gb.task_counters = counters_.get();
	uint64_t pairs = 0;
	for (int k = 0; k < count; k++)
	{
		const int i = which[first + k];
		const ListBlob &lb = blobs_[i];
		const unsigned char *d = d_blobs_.get() + (size_t)i * cap_;
		ScreenList &S = ls.l[k];
		S.hap = d; S.cells = (const CellTask *)(d + lb.off_cells);
		S.cand_col = cols_[i]; S.n_hap = lb.n_hap; S.cand_bit = cand_bit;
		GatherList &L = gb.lists[k];
		L.hap = d; L.cells = S.cells; L.cand_col = cols_[i];
		L.P = a.P + (size_t)k * n_cells * p_stride_;
		L.n_hap = lb.n_hap; L.cand_bit = cand_bit;
		L.task_prefix = a.task_prefix + (size_t)k * (n_cells + 2);
		L.count = (refine ? a.count2 : a.count) + (size_t)k * n_cells;
		L.entries = (refine ? a.entries2 : a.entries) + (size_t)k * n_cells * p_stride_;
		a.n_dist = gb.n_dist = lb.n_dist;
		if (lb.n_hap > gb.max_hap) gb.max_hap = lb.n_hap;
		pairs += lb.pairs_per_sample;
	}
	// tasks per list that give every warp slot of the GPU a couple of tasks
	const int target_tasks = (di.sm_count * 48 + count - 1) / count;
	// persistent CTAs of the gather launch: sized from what the previous pass of this kind
	// executed per list (the lists of consecutive rounds are alike), so that a small pass does not
	// occupy every SM and the passes of other lanes (other streams) run beside it
	long long max_ctas = 0;
	if (evals_per_list_[kind] > 0)
	{
		max_ctas = (long long)(evals_per_list_[kind] * 1.5 * count / 4e5) + di.sm_count / 2;
		if (max_ctas < di.sm_count) max_ctas = di.sm_count;
	}
	int nw;
	// out-of-bag passes: a few hundred tasks of very unequal length per list, so every CTA serves one
	// list only (no barrier between lists). HIBAG_B200_GATHER_FLAT: 0 (cell, positions) tasks walking
	// all lists, 2 the same with one list per CTA (default for out-of-bag passes), 1 / 3 entry-flat
	// tasks (cell_gather_flat_kernel, one list per CTA), +4: also for the in-bag passes
	int flat_mode = 2;                      // read per call: the tests switch forms inside one process
	if (const char *e = getenv("HIBAG_B200_GATHER_FLAT")) flat_mode = atoi(e);
	const int gform = (kind == 0 || (flat_mode & 4)) ? (flat_mode & 3) : 0;
	const bool flat = (gform & 1) != 0;
	gb.flat = flat ? 3 : gform;
	static const int excl = []() { const char *e = getenv("HIBAG_B200_GATHER_EXCL"); return e ? atoi(e) : 0; }();
	if (excl)
	{
		// the small kernels of the pass run on this lane's own stream; only the gather launch goes
		// to the device's ONE exclusive scoring stream, so that gather launches of different lanes
		// never overlap each other: each has the GPU's scoring resources to itself and the CUDA
		// events around it time it alone
		cudaStream_t s = st_.s;
		HB_CUDA(cudaMemsetAsync(counters_.get(), 0, sizeof(unsigned int) * MAX_BATCH_LISTS, s));
		HB_CUDA(cudaMemsetAsync(a.count, 0, sizeof(int) * (size_t)count * n_cells, s));
		if (dedup) HB_CUDA(cudaMemsetAsync(dd_table, 0, dd_bytes, s));
		HB_CUDA(cudaEventRecord(ev0_.e, s));
		if (dedup) launch_screen_dedup(a, ls, dd_table, table_size, (int *)a.rep, (int *)a.rep_list, dd_n_rep, s);
		launch_screen_bound(a, ls, s);
		launch_screen_need(a, s);
		if (refine) launch_screen_refine(a, s);
		launch_screen_tasks(ls, count, n_cells_, refine ? a.count2 : a.count, n_cells, a.task_prefix, a.evals, target_tasks,
			di.sm_count * 32, s, flat);
		HB_CUDA(cudaEventRecord(ev_up_.e, s));
		// HIBAG_B200_GATHER_OOB: 0 out-of-bag launches share the exclusive stream, 1 their own exclusive
		// stream (they are latency-bound and small: they run beside the in-bag launches), 2 the lane's
		// own stream (also concurrent with each other)
		static const int oob_mode = []() { const char *e = getenv("HIBAG_B200_GATHER_OOB"); return e ? atoi(e) : 0; }();
		if (kind == 0 && oob_mode == 2)
		{
			HB_CUDA(cudaEventRecord(ev_g0_.e, s));
			nw = launch_cell_gather(gb, di.sm_count, s, max_ctas);
			HB_CUDA(cudaEventRecord(ev_g1_.e, s));
		} else {
			ScoreQueue &q = ScoreQueue::get((kind == 0 && oob_mode == 1) ? 62 : 63);
			std::lock_guard<std::mutex> lk(q.mu);
			HB_CUDA(cudaStreamWaitEvent(q.st.s, ev_up_.e, 0));
			HB_CUDA(cudaEventRecord(ev_g0_.e, q.st.s));
			nw = launch_cell_gather(gb, di.sm_count, q.st.s, (excl > 1 || kind == 0) ? max_ctas : 0);
			HB_CUDA(cudaEventRecord(ev_g1_.e, q.st.s));
			HB_CUDA(cudaStreamWaitEvent(s, ev_g1_.e, 0));
		}
		if (kind == 0)
		{
			launch_reduce_oob_screened(a, d_counts_.get() + first, s);
			if (dedup) launch_screen_broadcast_oob(a, d_counts_.get() + first, s);
		} else {
			launch_reduce_ib_screened(a, ls, d_ratio_.get() + (size_t)first * ratio_stride_, ratio_stride_, s);
			if (dedup) launch_screen_broadcast_ib(a, d_ratio_.get() + (size_t)first * ratio_stride_, ratio_stride_, s);
		}
		HB_CUDA(cudaEventRecord(ev1_.e, s));
		HB_CUDA(cudaEventRecord(ev_up_.e, s));
	} else {
	ScoreQueue &q = ScoreQueue::get(queue_id_);
	{
		std::lock_guard<std::mutex> lk(q.mu);
		cudaStream_t s = q.st.s;
		HB_CUDA(cudaStreamWaitEvent(s, ev_up_.e, 0));
		HB_CUDA(cudaMemsetAsync(counters_.get(), 0, sizeof(unsigned int) * MAX_BATCH_LISTS, s));
		HB_CUDA(cudaMemsetAsync(a.count, 0, sizeof(int) * (size_t)count * n_cells, s));
		if (dedup) HB_CUDA(cudaMemsetAsync(dd_table, 0, dd_bytes, s));
		HB_CUDA(cudaEventRecord(ev0_.e, s));
		if (dedup) launch_screen_dedup(a, ls, dd_table, table_size, (int *)a.rep, (int *)a.rep_list, dd_n_rep, s);
		launch_screen_bound(a, ls, s);
		launch_screen_need(a, s);
		if (refine) launch_screen_refine(a, s);
		launch_screen_tasks(ls, count, n_cells_, refine ? a.count2 : a.count, n_cells, a.task_prefix, a.evals, target_tasks,
			di.sm_count * 32, s, flat);
		HB_CUDA(cudaEventRecord(ev_g0_.e, s));
		nw = launch_cell_gather(gb, di.sm_count, s, max_ctas);
		HB_CUDA(cudaEventRecord(ev_g1_.e, s));
		if (kind == 0)
		{
			launch_reduce_oob_screened(a, d_counts_.get() + first, s);
			if (dedup) launch_screen_broadcast_oob(a, d_counts_.get() + first, s);
		} else {
			launch_reduce_ib_screened(a, ls, d_ratio_.get() + (size_t)first * ratio_stride_, ratio_stride_, s);
			if (dedup) launch_screen_broadcast_ib(a, d_ratio_.get() + (size_t)first * ratio_stride_, ratio_stride_, s);
		}
		HB_CUDA(cudaEventRecord(ev1_.e, s));
	}
	HB_CUDA(cudaStreamWaitEvent(st_.s, ev1_.e, 0));
	HB_CUDA(cudaEventRecord(ev_up_.e, st_.s));
	}
	stats.launches += dedup ? 7 : 5; stats.cell_launches += 1;
	stats.pair_evals_nominal += pairs * (uint64_t)n_pos;
	(void)nw;
}

void BatchScorer::add_gather_time(bool in_bag)
{
	float ms = 0;
	HB_CUDA(cudaEventElapsedTime(&ms, ev_g0_.e, ev_g1_.e));
	stats.gather_ms += ms; stats.gather_launches++;
	if (in_bag) { stats.gather_ib_ms += ms; stats.gather_ib_launches++; }
}

void BatchScorer::score_oob(const GenoView &g, int cand_bit, const std::vector<int> &which,
	const int *pos_list, int n_pos, std::vector<int> &counts)
{
	const int n = (int)which.size();
	counts.assign(n, 0);
	if (n == 0 || n_pos <= 0) return;
	const bool screen = screen_ && n_pos == (int)set_samples_[0].size();
	p_stride_ = ((size_t)n_pos + 31) & ~(size_t)31;
	P_.ensure((size_t)n * n_cells_ * p_stride_);
	if (screen)
	{
		U_.ensure((size_t)n * n_hla_ * p_stride_);
		xref_.ensure((size_t)n * p_stride_);
		cnt_.ensure((size_t)n * n_cells_);
		ent_.ensure((size_t)n * n_cells_ * p_stride_);
		prefix_.ensure((size_t)n * (n_cells_ + 2));
		HB_CUDA(cudaMemsetAsync(d_evals_.get(), 0, sizeof(unsigned long long) * (size_t)(n + 1), st_.s));
	}
	HB_CUDA(cudaMemsetAsync(d_counts_.get(), 0, sizeof(int) * (size_t)n, st_.s));
	HB_CUDA(cudaEventRecord(ev_up_.e, st_.s));
	uint64_t before = stats.pair_evals;
	for (int first = 0; first < n; first += MAX_BATCH_LISTS)
	{
		const int count = std::min(MAX_BATCH_LISTS, n - first);
		if (screen) run_cells_screened(g, cand_bit, which, first, count, pos_list, n_pos, 0);
		else run_cells(g, cand_bit, which, first, count, pos_list, n_pos, P_.get(), p_stride_);
		float ms = 0;
		if (first + count < n)        // counters are reused by the next sub-batch
		{
			HB_CUDA(cudaEventSynchronize(ev1_.e));
			HB_CUDA(cudaEventElapsedTime(&ms, ev0_.e, ev1_.e));
			stats.cell_ms += ms; stats.kernel_ms += ms;
			if (screen) add_gather_time(false);
		}
	}
	if (!screen)
	{
		stats.pair_evals_nominal += stats.pair_evals - before;
		launch_reduce_oob(P_.get(), p_stride_, n_hla_, pos_list, n_pos, g.a1, g.a2, d_counts_.get(),
			st_.s, n, (size_t)n_cells_ * p_stride_);
		stats.launches++;
	} else
		HB_CUDA(cudaMemcpyAsync(h_evals_.get(), d_evals_.get(), sizeof(unsigned long long) * (size_t)(n + 1),
			cudaMemcpyDeviceToHost, st_.s));
	HB_CUDA(cudaMemcpyAsync(h_counts_.get(), d_counts_.get(), sizeof(int) * (size_t)n,
		cudaMemcpyDeviceToHost, st_.s));
	HB_CUDA(cudaEventRecord(ev_done_.e, st_.s));
	HB_CUDA(cudaEventSynchronize(ev_done_.e));
	float ms = 0;
	HB_CUDA(cudaEventElapsedTime(&ms, ev0_.e, ev1_.e));
	stats.cell_ms += ms; stats.kernel_ms += ms;
	stats.d2h_bytes += sizeof(int) * (size_t)n;
	if (screen)
	{
		add_gather_time(false);
		const int nw = geno_words(n_snp_);
		uint64_t tot = 0;
		for (int k = 0; k < n; k++) tot += h_evals_.get()[k];
		stats.pair_evals += tot;
		stats.popc32 += tot * (uint64_t)nw;
		evals_per_list_[0] = (double)tot / n;
		stats.d2h_bytes += sizeof(unsigned long long) * (size_t)n;
	}
	for (int k = 0; k < n; k++) counts[k] = h_counts_.get()[k];
}

void BatchScorer::score_ib(const GenoView &g, int cand_bit, const std::vector<int> &which,
	const int *pos_list, int n_pos)
{
	const int n = (int)which.size();
	if (n == 0 || n_pos <= 0) return;
	const bool screen = screen_ && n_pos == (int)set_samples_[1].size();
	p_stride_ = ((size_t)n_pos + 31) & ~(size_t)31;
	ratio_stride_ = p_stride_;
	P_.ensure((size_t)n * n_cells_ * p_stride_);
	d_ratio_.ensure((size_t)n * ratio_stride_);
	h_ratio_.ensure((size_t)n * ratio_stride_);
	if (screen)
	{
		U_.ensure((size_t)n * n_hla_ * p_stride_);
		xref_.ensure((size_t)n * p_stride_);
		cnt_.ensure((size_t)n * n_cells_);
		ent_.ensure((size_t)n * n_cells_ * p_stride_);
		prefix_.ensure((size_t)n * (n_cells_ + 2));
		if (screen_refine())
		{
			U2_.ensure((size_t)n * n_hla_ * 4 * p_stride_);
			hetk_.ensure((size_t)n * p_stride_);
			cnt2_.ensure((size_t)n * n_cells_);
			ent2_.ensure((size_t)n * n_cells_ * p_stride_);
		}
		HB_CUDA(cudaMemsetAsync(d_evals_.get(), 0, sizeof(unsigned long long) * (size_t)(n + 1), st_.s));
	}
	HB_CUDA(cudaEventRecord(ev_up_.e, st_.s));
	uint64_t before = stats.pair_evals;
	for (int first = 0; first < n; first += MAX_BATCH_LISTS)
	{
		const int count = std::min(MAX_BATCH_LISTS, n - first);
		if (screen) run_cells_screened(g, cand_bit, which, first, count, pos_list, n_pos, 1);
		else run_cells(g, cand_bit, which, first, count, pos_list, n_pos, P_.get(), p_stride_);
		if (first + count < n)
		{
			float ms = 0;
			HB_CUDA(cudaEventSynchronize(ev1_.e));
			HB_CUDA(cudaEventElapsedTime(&ms, ev0_.e, ev1_.e));
			stats.cell_ms += ms; stats.kernel_ms += ms;
			if (screen) add_gather_time(true);
		}
	}
	if (!screen)
	{
		stats.pair_evals_nominal += stats.pair_evals - before;
		launch_reduce_ib(P_.get(), p_stride_, n_hla_, pos_list, n_pos, g.a1, g.a2, d_ratio_.get(),
			st_.s, n, (size_t)n_cells_ * p_stride_, ratio_stride_);
		stats.launches++;
	} else
		HB_CUDA(cudaMemcpyAsync(h_evals_.get(), d_evals_.get(), sizeof(unsigned long long) * (size_t)(n + 1),
			cudaMemcpyDeviceToHost, st_.s));
	HB_CUDA(cudaMemcpyAsync(h_ratio_.get(), d_ratio_.get(), sizeof(double) * (size_t)n * ratio_stride_,
		cudaMemcpyDeviceToHost, st_.s));
	HB_CUDA(cudaEventRecord(ev_done_.e, st_.s));
	HB_CUDA(cudaEventSynchronize(ev_done_.e));
	float ms = 0;
	HB_CUDA(cudaEventElapsedTime(&ms, ev0_.e, ev1_.e));
	stats.cell_ms += ms; stats.kernel_ms += ms;
	stats.d2h_bytes += sizeof(double) * (size_t)n * ratio_stride_;
	if (screen)
	{
		add_gather_time(true);
		const int nw = geno_words(n_snp_);
		uint64_t tot = 0;
		for (int k = 0; k < n; k++) tot += h_evals_.get()[k];
		stats.pair_evals += tot;
		stats.popc32 += tot * (uint64_t)nw;
		stats.gather_ib_popc32 += tot * (uint64_t)nw;
		evals_per_list_[1] = (double)tot / n;
		stats.screen_fallback += h_evals_.get()[n];
		stats.d2h_bytes += sizeof(unsigned long long) * (size_t)(n + 1);
		rescore_uncertified(g, cand_bit, which);
	}
}

/// In-bag positions whose screened sum could not be certified (ratio -1): score them against
/// every cell with the plain kernel and patch their ratios.
void BatchScorer::rescore_uncertified(const GenoView &g, int cand_bit, const std::vector<int> &which)
{
	const int n = (int)which.size();
	const int n_pos = (int)set_samples_[1].size();
	std::vector<int> lists_fb;               // positions k in `which`
	std::vector<char> mark(n_pos, 0);
	const bool force = getenv("HIBAG_B200_SCREEN_FORCE_FALLBACK") != nullptr;   // test hook
	for (int k = 0; k < n; k++)
	{
		double *r = h_ratio_.get() + (size_t)k * ratio_stride_;
		if (force) for (int p = k % 7; p < n_pos; p += 7) r[p] = -1.0;
		bool any = false;
		for (int p = 0; p < n_pos; p++)
			if (r[p] == -1.0) { mark[p] = 1; any = true; }
		if (any) lists_fb.push_back(k);
	}
	if (lists_fb.empty()) return;
	std::vector<int> F, fb_samp;
	for (int p = 0; p < n_pos; p++)
		if (mark[p]) { F.push_back(p); fb_samp.push_back(set_samples_[1][p]); }
	const int nf = (int)F.size(), nl = (int)lists_fb.size();
	const size_t stride = ((size_t)nf + 31) & ~(size_t)31;
	d_fb_samp_.ensure(nf);
	P_fb_.ensure((size_t)nl * n_cells_ * stride);
	d_ratio_fb_.ensure((size_t)nl * stride);
	h_ratio_fb_.ensure((size_t)nl * stride);
	HB_CUDA(cudaMemcpyAsync(d_fb_samp_.get(), fb_samp.data(), sizeof(int) * (size_t)nf,
		cudaMemcpyHostToDevice, st_.s));
	stream_sync_blocking(st_.s);
	HB_CUDA(cudaEventRecord(ev_up_.e, st_.s));
	std::vector<int> which_fb(nl);
	for (int j = 0; j < nl; j++) which_fb[j] = which[lists_fb[j]];
	for (int first = 0; first < nl; first += MAX_BATCH_LISTS)
	{
		const int count = std::min(MAX_BATCH_LISTS, nl - first);
		run_cells(g, cand_bit, which_fb, first, count, d_fb_samp_.get(), nf, P_fb_.get(), stride, queue_id_);
		float ms = 0;
		HB_CUDA(cudaEventSynchronize(ev1_.e));
		HB_CUDA(cudaEventElapsedTime(&ms, ev0_.e, ev1_.e));
		stats.cell_ms += ms; stats.kernel_ms += ms;
	}
	launch_reduce_ib(P_fb_.get(), stride, n_hla_, d_fb_samp_.get(), nf, g.a1, g.a2, d_ratio_fb_.get(),
		st_.s, nl, (size_t)n_cells_ * stride, stride);
	HB_CUDA(cudaMemcpyAsync(h_ratio_fb_.get(), d_ratio_fb_.get(), sizeof(double) * (size_t)nl * stride,
		cudaMemcpyDeviceToHost, st_.s));
	HB_CUDA(cudaEventRecord(ev_done_.e, st_.s));
	HB_CUDA(cudaEventSynchronize(ev_done_.e));
	stats.launches++; stats.d2h_bytes += sizeof(double) * (size_t)nl * stride;
	stats.h2d_bytes += sizeof(int) * (size_t)nf;
	for (int j = 0; j < nl; j++)
	{
		double *r = h_ratio_.get() + (size_t)lists_fb[j] * ratio_stride_;
		const double *fb = h_ratio_fb_.get() + (size_t)j * stride;
		for (int i = 0; i < nf; i++)
			if (r[F[i]] == -1.0) { r[F[i]] = fb[i]; stats.screen_fallback++; }
	}
}

}  // namespace hb
