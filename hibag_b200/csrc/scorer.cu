// scorer.cu -- see scorer.h
#include "scorer.h"

#include <cstring>
#include <map>
#include <mutex>

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

// ---- EvalSlot ---------------------------------------------------------------------------------
EvalSlot::EvalSlot()
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
}

void EvalSlot::enqueue_cells(const GenoView &g, const int *pos_list, int n_pos)
{
	if (n_pos <= 0) return;
	const DeviceInfo &di = current_device();
	p_stride_ = ((size_t)n_pos + 31) & ~(size_t)31;
	P_.ensure(p_stride_ * (size_t)blob_.n_cells);
	CellPass p;
	memset(&p, 0, sizeof(p));
	bind_list(blob_, d_blob_.get(), device_rare_freq_table(), p);
	p.s1 = g.s1; p.s2 = g.s2; p.geno_stride = g.stride;
	p.cand_col = g.cand_col; p.cand_bit = g.cand_bit;
	p.samp_list = pos_list; p.n_pos = n_pos;
	p.task_counter = counter_.get();
	p.P = P_.get(); p.p_stride = p_stride_;
	HB_CUDA(cudaMemsetAsync(counter_.get(), 0, sizeof(unsigned int), st_.s));
	const int R = choose_samples_per_lane(n_pos, blob_.n_chunks, blob_.n_snp, di.sm_count);
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

}  // namespace hb
