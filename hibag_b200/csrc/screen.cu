// screen.cu -- exact screening of the training passes and the gather form of the pair-scoring
// kernel (see the block comment in kernels.h).
//
// What is skipped is PROVEN irrelevant for the integer / fp64 outputs the reference computes
// (out-of-bag best guess, src/LibHLA.cpp:1639-1704; in-bag P_true / sum, :1706-1767); what is
// evaluated is the reference's own sequential chain -- identical code to cell_pass_kernel's
// (kernels.cu), un-fused multiply then add in (i outer, j inner) order.

#include "kernels.h"

#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>

#include "devutil.cuh"

namespace hb {

#define CUDA_CHECK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) \
	throw std::runtime_error(std::string("CUDA error: ") + cudaGetErrorString(e_) + \
		" at " __FILE__ ":" + std::to_string(__LINE__)); } while (0)

static constexpr int GATHER_THREADS = 128;
static constexpr int NA_INT = INT32_MIN;

/// genotype bit planes of one sample, candidate SNP patched in at bit cand_bit
/// (CGenotypeList::AddSNP, src/LibHLA.cpp:609-622, 860-874); !ok -> everything missing
template <int NW>
__device__ __forceinline__ void load_geno(const uint32_t *__restrict__ s1,
	const uint32_t *__restrict__ s2, int stride, int samp, bool ok,
	const int8_t *__restrict__ cand_col, int cand_bit, uint32_t (&S1)[NW], uint32_t (&S2)[NW])
{
#pragma unroll
	for (int w = 0; w < NW; w++)
	{
		S1[w] = ok ? __ldg(s1 + (size_t)w * stride + samp) : 0u;
		S2[w] = ok ? __ldg(s2 + (size_t)w * stride + samp) : 0xffffffffu;
	}
	if (cand_col != nullptr && ok)
	{
		const int g = __ldg(cand_col + samp);
		const int cw = cand_bit >> 5;
		const uint32_t bit = 1u << (cand_bit & 31);
#pragma unroll
		for (int w = 0; w < NW; w++)
		{
			if (w == cw)
			{
				if (g == 1 || g == 2) S1[w] |= bit; else S1[w] &= ~bit;
				if (g == 0 || g == 1) S2[w] &= ~bit; else S2[w] |= bit;
			}
		}
	}
}

/// upper bound of a cell's chain value from the per-allele sums (identical in every kernel)
__device__ __forceinline__ double screen_bound(double ua, double ub, double K)
{
	return __dmul_rn(__dmul_rn(ua, ub), K);
}

__device__ __forceinline__ int geno_words_dev(int n_snp)
{
	return (n_snp <= 32) ? 1 : ((n_snp <= 64) ? 2 : 4);
}

__device__ __forceinline__ int true_cell_index(int t1, int t2, int n_hla)
{
	return t2 + t1 * (2 * n_hla - t1 - 1) / 2;           // src/LibHLA.cpp:1712, t1 <= t2
}

/// the position thread k of a launch works on: k itself, or the k-th representative of list l when
/// position classes are on (ok = there is one)
__device__ __forceinline__ int screen_pos(const ScreenArgs &a, int l, int k, bool &ok)
{
	if (a.rep_list == nullptr) { ok = k < a.n_pos; return k; }
	ok = k < __ldg(a.n_rep + l);
	return ok ? __ldg(a.rep_list + (size_t)l * a.p_stride + k) : 0;
}

// ---------------------------------------------------------------------------------------
// position classes: equal packed genotype (candidate patched in) and equal true type
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t mix64s(uint64_t x)
{
	x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
	return x;
}

template <int NW>
__global__ void __launch_bounds__(256)
screen_dedup_kernel(const ScreenArgs a, const __grid_constant__ ScreenLists ls, int *table, unsigned mask,
	int *rep, int *rep_list, int *n_rep)
{
	SmAcct acct_scope(a.acct, SM_ACCT_DEDUP, 128u);       // 256 threads: 8 CTAs fit an SM
	__shared__ int sh_cnt[8];
	__shared__ int sh_base;
	const int l = blockIdx.y;
	const ScreenList &L = ls.l[l];
	const int pos = blockIdx.x * blockDim.x + threadIdx.x;
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	int *tab = table + (size_t)l * ((size_t)mask + 1);
	bool is_rep = false;
	if (pos < a.n_pos)
	{
		const int samp = a.samp_list ? __ldg(a.samp_list + pos) : pos;
		const int t1 = __ldg(a.a1 + samp), t2 = __ldg(a.a2 + samp);
		uint32_t S1[NW], S2[NW];
		load_geno<NW>(a.s1, a.s2, a.geno_stride, samp, true, L.cand_col, L.cand_bit, S1, S2);
		uint64_t h = mix64s(((uint64_t)(uint32_t)t1 << 32) | (uint32_t)t2);
#pragma unroll
		for (int w = 0; w < NW; w++) h = mix64s(h ^ (((uint64_t)S1[w] << 32) | S2[w]));
		unsigned slot = (unsigned)h & mask;
		for (;;)
		{
			int r = *(volatile int *)(tab + slot);             // position + 1; 0 = free
			if (r == 0)
			{
				r = atomicCAS(tab + slot, 0, pos + 1);
				if (r == 0) { is_rep = true; rep[(size_t)l * a.p_stride + pos] = pos; break; }
			}
			const int rp = r - 1;
			const int rs = a.samp_list ? __ldg(a.samp_list + rp) : rp;
			bool same = (__ldg(a.a1 + rs) == t1) && (__ldg(a.a2 + rs) == t2);
			if (same)
			{
				uint32_t R1[NW], R2[NW];
				load_geno<NW>(a.s1, a.s2, a.geno_stride, rs, true, L.cand_col, L.cand_bit, R1, R2);
#pragma unroll
				for (int w = 0; w < NW; w++) same = same && (R1[w] == S1[w]) && (R2[w] == S2[w]);
			}
			if (same) { rep[(size_t)l * a.p_stride + pos] = rp; break; }
			slot = (slot + 1) & mask;
		}
	}
	// the representatives of the block, compacted in position order (neighbouring threads of the
	// later kernels then touch neighbouring columns of U and P)
	const unsigned m = __ballot_sync(0xffffffffu, is_rep);
	if (lane == 0) sh_cnt[wid] = __popc(m);
	__syncthreads();
	if (threadIdx.x == 0)
	{
		int tot = 0;
		for (int w = 0; w < 8; w++) { const int c = sh_cnt[w]; sh_cnt[w] = tot; tot += c; }
		sh_base = tot ? atomicAdd(n_rep + l, tot) : 0;
	}
	__syncthreads();
	if (is_rep)
		rep_list[(size_t)l * a.p_stride + sh_base + sh_cnt[wid] + __popc(m & ((1u << lane) - 1u))] = pos;
}

void launch_screen_dedup(const ScreenArgs &a, const ScreenLists &ls, int *table, int table_size, int *rep,
	int *rep_list, int *n_rep, cudaStream_t st)
{
	if (a.n_pos <= 0 || a.n_lists <= 0) return;
	if (table_size < 2 * a.n_pos || (table_size & (table_size - 1)))
		throw std::runtime_error("launch_screen_dedup: table size must be a power of two >= 2 * n_pos");
	dim3 grid((a.n_pos + 255) / 256, a.n_lists);
	const unsigned mask = (unsigned)table_size - 1u;
	switch (geno_words(a.n_snp))
	{
	case 1: screen_dedup_kernel<1><<<grid, 256, 0, st>>>(a, ls, table, mask, rep, rep_list, n_rep); break;
	case 2: screen_dedup_kernel<2><<<grid, 256, 0, st>>>(a, ls, table, mask, rep, rep_list, n_rep); break;
	default: screen_dedup_kernel<4><<<grid, 256, 0, st>>>(a, ls, table, mask, rep, rep_list, n_rep); break;
	}
	CUDA_CHECK(cudaGetLastError());
}

__global__ void __launch_bounds__(256)
screen_broadcast_oob_kernel(const ScreenArgs a, int *out_count)
{
	const int l = blockIdx.y;
	const int pos = blockIdx.x * blockDim.x + threadIdx.x;
	int cnt = 0;
	if (pos < a.n_pos) cnt = a.pos_res[(size_t)l * a.p_stride + __ldg(a.rep + (size_t)l * a.p_stride + pos)];
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
	if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(out_count + l, cnt);
}

void launch_screen_broadcast_oob(const ScreenArgs &a, int *out_count, cudaStream_t st)
{
	if (a.n_pos <= 0 || a.n_lists <= 0) return;
	dim3 grid((a.n_pos + 255) / 256, a.n_lists);
	screen_broadcast_oob_kernel<<<grid, 256, 0, st>>>(a, out_count);
	CUDA_CHECK(cudaGetLastError());
}

__global__ void __launch_bounds__(256)
screen_broadcast_ib_kernel(const ScreenArgs a, double *out_ratio, size_t out_stride)
{
	const int l = blockIdx.y;
	const int pos = blockIdx.x * blockDim.x + threadIdx.x;
	if (pos >= a.n_pos) return;
	const int rp = __ldg(a.rep + (size_t)l * a.p_stride + pos);
	if (rp != pos) out_ratio[(size_t)l * out_stride + pos] = out_ratio[(size_t)l * out_stride + rp];
}

void launch_screen_broadcast_ib(const ScreenArgs &a, double *out_ratio, size_t out_stride, cudaStream_t st)
{
	if (a.n_pos <= 0 || a.n_lists <= 0) return;
	dim3 grid((a.n_pos + 255) / 256, a.n_lists);
	screen_broadcast_ib_kernel<<<grid, 256, 0, st>>>(a, out_ratio, out_stride);
	CUDA_CHECK(cudaGetLastError());
}

// ---------------------------------------------------------------------------------------
// U[l][a][pos] = sum_{i in allele a} f_i * T'[min(c_i, dmax)],  c_i = mismatches of h_i on the
// sample's homozygous SNPs (the c_i of cell_pass_kernel's one-popcount distance)
// ---------------------------------------------------------------------------------------
template <int NW>
__global__ void __launch_bounds__(128)
screen_bound_kernel(const ScreenArgs a, const __grid_constant__ ScreenLists ls)
{
	extern __shared__ int sh_al[];                        // start[n_hla], n[n_hla]
	SmAcct acct_scope(a.acct, SM_ACCT_BOUND, 64u);        // 128 threads: 16 CTAs fit an SM
	const int l = blockIdx.y;
	const ScreenList &L = ls.l[l];
	int *al_start = sh_al, *al_n = sh_al + a.n_hla;
	const int n_cells = a.n_hla * (a.n_hla + 1) / 2;
	for (int k = threadIdx.x; k < n_cells; k += blockDim.x)
	{
		const int4 ca = __ldg((const int4 *)(L.cells + k));
		const int4 cb = __ldg((const int4 *)((const char *)(L.cells + k) + 16));
		if (cb.y)
		{
			al_start[cb.z] = ca.x; al_n[cb.z] = ca.y;
			if (blockIdx.x == 0)     // kept for the in-bag reduction's rescue path
			{
				a.al_tab[((size_t)l * a.n_hla + cb.z) * 2] = ca.x;
				a.al_tab[((size_t)l * a.n_hla + cb.z) * 2 + 1] = ca.y;
			}
		}
	}
	__syncthreads();

	bool ok;
	const int pos = screen_pos(a, l, blockIdx.x * blockDim.x + threadIdx.x, ok);
	if (!__any_sync(0xffffffffu, ok)) return;             // (no barrier below)
	int samp = 0, t1 = -1, t2 = -1;
	if (ok)
	{
		samp = a.samp_list ? __ldg(a.samp_list + pos) : pos;
		t1 = __ldg(a.a1 + samp); t2 = __ldg(a.a2 + samp);
	}
	uint32_t S1[NW], S2[NW], HOM[NW], G2[NW], V[NW];
	load_geno<NW>(a.s1, a.s2, a.geno_stride, samp, ok, L.cand_col, L.cand_bit, S1, S2);
#pragma unroll
	for (int w = 0; w < NW; w++)
	{
		HOM[w] = ~(S1[w] ^ S2[w]); G2[w] = S1[w] & S2[w]; V[w] = S1[w] | ~S2[w];
	}
	const int dmax = a.n_dist - 1;
	const char *hap_g = (const char *)L.hap;
	double *U = a.U + (size_t)l * a.n_hla * a.p_stride;
	// second level: the sample's first two heterozygous SNPs (word, bit); a haplotype's class = its
	// alleles there. (A SNP the sample is heterozygous at adds a mismatch to every pair of haplotypes
	// that AGREE on it: the distance the product bound ignores.)
	const bool lvl2 = a.U2 != nullptr;
	int hw0 = 0, hw1 = 0, k_eff = 0;
	uint32_t hb0 = 0u, hb1 = 0u;
	if (lvl2)
	{
#pragma unroll
		for (int w = 0; w < NW; w++)
		{
			uint32_t het = S1[w] & ~S2[w];
			while (het && k_eff < 2)
			{
				const uint32_t bit = het & (0u - het);
				if (k_eff == 0) { hw0 = w; hb0 = bit; } else { hw1 = w; hb1 = bit; }
				k_eff++;
				het ^= bit;
			}
		}
		// (the true cell of the position rides in the upper bits: one load in screen_refine_kernel)
		if (ok) a.hetk[(size_t)l * a.p_stride + pos] = k_eff | (true_cell_index(t1, t2, a.n_hla) << 2);
	}
	double *U2 = lvl2 ? a.U2 + (size_t)l * a.n_hla * 4 * a.p_stride : nullptr;
	// the haplotype of each true allele with the largest f * T'[c] (first one on ties)
	int best1 = -1, best2 = -1;
	double bu1 = -1.0, bu2 = -1.0;
	for (int al = 0; al < a.n_hla; al++)
	{
		const int i0 = al_start[al], i1 = i0 + al_n[al];
		double acc = 0.0;
		double ac0 = 0.0, ac1 = 0.0, ac2 = 0.0, ac3 = 0.0;
#pragma unroll 4
		for (int i = i0; i < i1; i++)
		{
			HapRec<NW, false> h;
			h.load(0u, hap_g, i);
			int c = 0;
			uint32_t x0 = 0u, x1 = 0u;
#pragma unroll
			for (int w = 0; w < NW; w++)
			{
				c += __popc((h.h[w] ^ G2[w]) & HOM[w]);
				if (w == hw0) x0 = h.h[w] & hb0;
				if (w == hw1) x1 = h.h[w] & hb1;
			}
			const double t = __ldg(a.table_floor + min(c, dmax));
			const double u = __dmul_rn(h.f, t);
			acc = __dadd_rn(acc, u);
			if (lvl2)
			{
				const int cls = (x0 ? 1 : 0) | (x1 ? 2 : 0);
				ac0 = __dadd_rn(ac0, cls == 0 ? u : 0.0); ac1 = __dadd_rn(ac1, cls == 1 ? u : 0.0);
				ac2 = __dadd_rn(ac2, cls == 2 ? u : 0.0); ac3 = __dadd_rn(ac3, cls == 3 ? u : 0.0);
			}
			if (al == t1 && u > bu1) { bu1 = u; best1 = i; }
			if (al == t2 && u > bu2) { bu2 = u; best2 = i; }
		}
		if (ok) U[(size_t)al * a.p_stride + pos] = acc;
		if (ok && lvl2)
		{
			// [allele][position][class]: the four classes of a (allele, position) are one 32-byte sector
			double2 *u2 = (double2 *)(U2 + ((size_t)al * a.p_stride + pos) * 4);
			u2[0] = make_double2(ac0, ac1); u2[1] = make_double2(ac2, ac3);
		}
	}
	// ---- x_ref: ONE term of the true cell's chain, evaluated exactly as the chain does (kernels.cu)
	// -- a sum of non-negative terms is at least each term, so x_ref <= the true cell's value <= the
	// best cell's. The pair: the best haplotype of one allele with its best partner in the other,
	// tried from both sides.
	double xref = 0.0;
	if (ok && best1 >= 0 && best2 >= 0)
	{
		const bool diag = (t1 == t2);
		{
			HapRec<NW, false> hi;
			hi.load(0u, hap_g, best1);
			uint32_t K[NW];
			int ci = 0;
#pragma unroll
			for (int w = 0; w < NW; w++)
			{
				K[w] = S1[w] & (S2[w] | ~hi.h[w]);
				ci += __popc((hi.h[w] ^ G2[w]) & HOM[w]);
			}
			const double ff = __dmul_rn(2.0, hi.f);
			const int j0 = al_start[t2], j1 = j0 + al_n[t2];
			for (int j = diag ? best1 : j0; j < j1; j++)
			{
				HapRec<NW, false> hj;
				hj.load(0u, hap_g, j);
				int pc = 0;
#pragma unroll
				for (int w = 0; w < NW; w++) pc += __popc((hj.h[w] ^ K[w]) & V[w]);
				const double t = __ldg(a.table + min(ci + pc, dmax));
				const double v = (diag && j == best1) ? __dmul_rn(__dmul_rn(hi.f, hi.f), t)
					: __dmul_rn(__dmul_rn(ff, hj.f), t);
				if (v > xref) xref = v;
			}
		}
		{
			HapRec<NW, false> hj;
			hj.load(0u, hap_g, best2);
			const int i0 = al_start[t1], i1 = i0 + al_n[t1];
			for (int i = i0; i < (diag ? best2 + 1 : i1); i++)
			{
				HapRec<NW, false> hi;
				hi.load(0u, hap_g, i);
				int ci = 0, pc = 0;
#pragma unroll
				for (int w = 0; w < NW; w++)
				{
					const uint32_t K = S1[w] & (S2[w] | ~hi.h[w]);
					ci += __popc((hi.h[w] ^ G2[w]) & HOM[w]);
					pc += __popc((hj.h[w] ^ K) & V[w]);
				}
				const double t = __ldg(a.table + min(ci + pc, dmax));
				const double v = (diag && i == best2) ? __dmul_rn(__dmul_rn(hi.f, hi.f), t)
					: __dmul_rn(__dmul_rn(__dmul_rn(2.0, hi.f), hj.f), t);
				if (v > xref) xref = v;
			}
		}
	}
	if (ok) a.xref[(size_t)l * a.p_stride + pos] = xref;
}

void launch_screen_bound(const ScreenArgs &a, const ScreenLists &ls, cudaStream_t st)
{
	if (a.n_pos <= 0 || a.n_lists <= 0) return;
	dim3 grid((a.n_pos + 127) / 128, a.n_lists);
	const size_t smem = sizeof(int) * 2 * (size_t)a.n_hla;
	switch (geno_words(a.n_snp))
	{
	case 1: screen_bound_kernel<1><<<grid, 128, smem, st>>>(a, ls); break;
	case 2: screen_bound_kernel<2><<<grid, 128, smem, st>>>(a, ls); break;
	default: screen_bound_kernel<4><<<grid, 128, smem, st>>>(a, ls); break;
	}
	CUDA_CHECK(cudaGetLastError());
}

// ---------------------------------------------------------------------------------------
// per list: how many positions one gather task takes (32, 64 or 128 -- small when the launch
// would otherwise have too few tasks to fill the GPU) and the exclusive prefix of
// ceil(count / that) over the blob's cell order.  task_prefix[l]: [n_cells] prefix,
// (bits 30-31 of an entry: log2(positions per task of that cell) - 5; fewer for the largest cells),
// [n_cells] = number of tasks
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
screen_tasks_kernel(const __grid_constant__ ScreenLists ls, int n_cells,
	const int *__restrict__ count, size_t count_stride, unsigned int *task_prefix,
	unsigned long long *evals, int target_tasks, int n_lists, int warp_slots, int flat)
{
	__shared__ unsigned int seg[256];
	__shared__ unsigned long long seg_ev[256];
	__shared__ unsigned int sh_shift;
	__shared__ unsigned long long sh_limit;
	const int l = blockIdx.x, tid = threadIdx.x;
	const CellTask *cells = ls.l[l].cells;
	const int *cnt = count + (size_t)l * count_stride;
	unsigned int *pre = task_prefix + (size_t)l * (n_cells + 2);
	const int per = (n_cells + 255) / 256;
	const int k0 = min(n_cells, tid * per), k1 = min(n_cells, k0 + per);
	unsigned int tot = 0;
	unsigned long long ev = 0;
	for (int k = k0; k < k1; k++)
	{
		const int4 ca = __ldg((const int4 *)(cells + k));
		const int4 cb = __ldg((const int4 *)((const char *)(cells + k) + 16));
		const int c = __ldg(cnt + cb.x);
		tot += (unsigned int)c;
		const unsigned long long pairs = cb.y ? (unsigned long long)ca.y * (ca.y + 1) / 2
			: (unsigned long long)ca.y * (unsigned long long)ca.w;
		ev += pairs * (unsigned long long)c;
	}
	seg[tid] = tot; seg_ev[tid] = ev;
	__syncthreads();
	if (flat)
	{
		// entry-flat tasks (cell_gather_flat_kernel): the plain exclusive prefix of the counts over the
		// blob's cell order; a task is 32 consecutive entries of that order, whatever cells they are in
		if (tid == 0)
		{
			unsigned int run = 0;
			for (int t = 0; t < 256; t++) { const unsigned int v = seg[t]; seg[t] = run; run += v; }
			pre[n_cells] = (run + 31u) >> 5;
			pre[n_cells + 1] = run;
		}
		__syncthreads();
		unsigned int run = seg[tid];
		for (int k = k0; k < k1; k++)
		{
			const int4 cb = __ldg((const int4 *)((const char *)(cells + k) + 16));
			pre[k] = run;
			run += (unsigned int)__ldg(cnt + cb.x);
		}
		if (ev) atomicAdd(evals + l, ev);
		return;
	}
	if (tid == 0)
	{
		unsigned int e = 0;
		unsigned long long w = 0;
		for (int t = 0; t < 256; t++) { e += seg[t]; w += seg_ev[t]; }
		// few positions in the whole list: small tasks, so that the launch still has enough of them
		unsigned int sh = 7;
		while (sh > 5 && (e >> sh) < (unsigned int)target_tasks) sh--;
		sh_shift = sh;
		// no task above ~0.4 of a warp slot's share of the launch (the lists of a launch are alike):
		// the chains of the largest cells would otherwise be the tail of the launch
		unsigned long long lim = w * (unsigned long long)n_lists * 2ull / (5ull * (unsigned long long)warp_slots);
		if (lim < 32768ull) lim = 32768ull;
		sh_limit = lim;
	}
	__syncthreads();
	const unsigned int gsh = sh_shift;
	const unsigned long long lim = sh_limit;
	unsigned int s = 0;
	for (int k = k0; k < k1; k++)
	{
		const int4 ca = __ldg((const int4 *)(cells + k));
		const int4 cb = __ldg((const int4 *)((const char *)(cells + k) + 16));
		const unsigned long long pairs = cb.y ? (unsigned long long)ca.y * (ca.y + 1) / 2
			: (unsigned long long)ca.y * (unsigned long long)ca.w;
		unsigned int sh = gsh;
		while (sh > 5 && (pairs << sh) > lim) sh--;
		s += ((unsigned int)__ldg(cnt + cb.x) + ((1u << sh) - 1u)) >> sh;
	}
	__syncthreads();
	seg[tid] = s;
	__syncthreads();
	if (tid == 0)
	{
		unsigned int run = 0;
		for (int t = 0; t < 256; t++) { const unsigned int v = seg[t]; seg[t] = run; run += v; }
		pre[n_cells] = run;
		pre[n_cells + 1] = gsh;
	}
	__syncthreads();
	unsigned int run = seg[tid];
	for (int k = k0; k < k1; k++)
	{
		const int4 ca = __ldg((const int4 *)(cells + k));
		const int4 cb = __ldg((const int4 *)((const char *)(cells + k) + 16));
		const unsigned long long pairs = cb.y ? (unsigned long long)ca.y * (ca.y + 1) / 2
			: (unsigned long long)ca.y * (unsigned long long)ca.w;
		unsigned int sh = gsh;
		while (sh > 5 && (pairs << sh) > lim) sh--;
		pre[k] = run | ((sh - 5u) << 30);            // bits 30-31: log2(positions per task) - 5
		run += ((unsigned int)__ldg(cnt + cb.x) + ((1u << sh) - 1u)) >> sh;
	}
	if (ev) atomicAdd(evals + l, ev);
}

void launch_screen_tasks(const ScreenLists &ls, int n_lists, int n_cells, const int *count,
	size_t count_stride, unsigned int *task_prefix, unsigned long long *evals, int target_tasks,
	int warp_slots, cudaStream_t st, bool flat)
{
	if (n_lists <= 0) return;
	screen_tasks_kernel<<<n_lists, 256, 0, st>>>(ls, n_cells, count, count_stride, task_prefix, evals,
		target_tasks, n_lists, warp_slots, flat ? 1 : 0);
	CUDA_CHECK(cudaGetLastError());
}

// ---------------------------------------------------------------------------------------
// which cells does a position need? lanes = consecutive positions; warp-aggregated append
// ---------------------------------------------------------------------------------------
static constexpr int NEED_TILE = 1024;       // cells per pass over the shared-memory masks

__global__ void __launch_bounds__(128)
screen_need_kernel(const ScreenArgs a)
{
	// per warp: the ballot of every cell of the tile, then the base offsets the appends start at.
	// The atomics of a tile are issued one per lane, all in flight together, instead of one
	// round trip to L2 per cell.
	__shared__ unsigned int sh_mask[4][NEED_TILE];
	__shared__ int sh_base[4][NEED_TILE];
	SmAcct acct_scope(a.acct, SM_ACCT_NEED, 147u);        // 32 KB of shared memory: 7 CTAs fit an SM
	const int l = blockIdx.y;
	bool ok;
	const int pos = screen_pos(a, l, blockIdx.x * blockDim.x + threadIdx.x, ok);
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	const int n = a.n_hla;
	const int n_cells = n * (n + 1) / 2;
	const double *U = a.U + (size_t)l * n * a.p_stride;
	int *cnt = a.count + (size_t)l * n_cells;
	int *ent = a.entries + (size_t)l * n_cells * a.p_stride;
	unsigned int *wm = sh_mask[wid];
	int *wb = sh_base[wid];
	// the position's per-allele sums: n loads into shared memory instead of n (n + 1) / 2 from L1/L2
	extern __shared__ double sU[];
	const bool us = a.u_smem != 0;
	if (us)
	{
		for (int al = 0; al < n; al++) sU[al * 128 + threadIdx.x] = ok ? U[(size_t)al * a.p_stride + pos] : 0.0;
		__syncthreads();
	}
	if (!__any_sync(0xffffffffu, ok)) return;             // (warp-level synchronisation only from here on)
	auto Uat = [&](int al) -> double { return us ? sU[al * 128 + threadIdx.x] : U[(size_t)al * a.p_stride + pos]; };
	int true_idx = -1;
	double thr = 0.0;
	if (ok)
	{
		const int samp = a.samp_list ? __ldg(a.samp_list + pos) : pos;
		true_idx = true_cell_index(__ldg(a.a1 + samp), __ldg(a.a2 + samp), n);
		thr = __dmul_rn(a.xref[(size_t)l * a.p_stride + pos], a.tau);
	}
	// a tile of cell masks is flushed: one atomic per cell per LANE reserves the slots of the warp's
	// positions, then the lanes append their positions
	auto flush = [&](int t0, int nt)
	{
		__syncwarp();
		for (int q = lane; q < nt; q += 32)
		{
			const unsigned m = wm[q];
			if (m) wb[q] = atomicAdd(cnt + t0 + q, __popc(m));
		}
		__syncwarp();
		const unsigned lt = (1u << lane) - 1u;
		for (int q = 0; q < nt; q++)
		{
			const unsigned m = wm[q];
			if ((m >> lane) & 1u)
				ent[(size_t)(t0 + q) * a.p_stride + wb[q] + __popc(m & lt)] = pos;
		}
		__syncwarp();
	};
	// cells in the blob's order (row al, columns bl >= al), eight per step so that the loads of the
	// per-allele sums overlap (one dependent L2 round trip per cell otherwise)
	int idx = 0, t0 = 0;
	for (int al = 0; al < n; al++)
	{
		const double ua = ok ? Uat(al) : 0.0;
		for (int bl = al; bl < n; bl += 8)
		{
			const int nb = min(8, n - bl);
			if (idx - t0 + nb > NEED_TILE) { flush(t0, idx - t0); t0 = idx; }
			double ub[8];
#pragma unroll
			for (int q = 0; q < 8; q++) ub[q] = (ok && q < nb) ? Uat(bl + q) : 0.0;
#pragma unroll
			for (int q = 0; q < 8; q++)
			{
				if (q < nb)                        // warp-uniform
				{
					const double bd = screen_bound(ua, ub[q], a.K);
					const bool need = ok && ((idx + q) == true_idx || (bd >= thr && bd > 0.0));
					const unsigned mask = __ballot_sync(0xffffffffu, need);
					if (lane == 0) wm[idx + q - t0] = mask;
				}
			}
			idx += nb;
		}
	}
	if (idx > t0) flush(t0, idx - t0);
}

void launch_screen_need(const ScreenArgs &a, cudaStream_t st)
{
	if (a.n_pos <= 0 || a.n_lists <= 0) return;
	dim3 grid((a.n_pos + 127) / 128, a.n_lists);
	const size_t smem = a.u_smem ? sizeof(double) * 128 * (size_t)a.n_hla : 0;
	if (smem > 16 * 1024)
		CUDA_CHECK(cudaFuncSetAttribute(screen_need_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
	screen_need_kernel<<<grid, 128, smem, st>>>(a);
	CUDA_CHECK(cudaGetLastError());
}

// ---------------------------------------------------------------------------------------
// second level of the in-bag screen. The product bound U_a U_b K ignores the heterozygous SNPs:
// d(g,i,j) = c_i + c_j + #{het SNPs on which h_i and h_j AGREE}. With the haplotypes of an allele
// split by their alleles at the sample's first k_eff <= 2 heterozygous SNPs (class s, U_a[s] from
// screen_bound_kernel), a pair of classes (s, t) agrees on k_eff - popc(s ^ t) of them, hence
//     P(a,b) <= bound2(a,b) = K2 * sum_{s,t} U_a[s] U_b[t] T'[k_eff - popc(s ^ t)]
// (K2: the factor 2, the table's rounding over THREE factors, the chains' rounding and this sum's).
// One thread per (cell, position) entry of the first level's need lists -- no divergence: a sample's
// survivors are its own entries. A survivor is appended to entries2; a discarded entry leaves
// MINUS its bound in the cell matrix, where the reduction finds it and adds it to the upper sum of
// its certificate, exactly as it adds the product bound of the cells the first level skipped.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ double screen_bound2(const double *ua, const double *ub, int k_eff,
	double K2, const double (&tf)[3])
{
	// w[j] = T'[k_eff - j] for j = popc(s ^ t) disagreements (classes that do not occur have U = 0)
	const double w0 = tf[k_eff], w1 = tf[max(k_eff - 1, 0)], w2 = tf[max(k_eff - 2, 0)];
	const double2 a01 = __ldg((const double2 *)ua), a23 = __ldg((const double2 *)ua + 1);
	const double2 b01 = __ldg((const double2 *)ub), b23 = __ldg((const double2 *)ub + 1);
	const double a4[4] = { a01.x, a01.y, a23.x, a23.y }, b4[4] = { b01.x, b01.y, b23.x, b23.y };
	double sum = 0.0;
#pragma unroll
	for (int sI = 0; sI < 4; sI++)
#pragma unroll
		for (int tI = 0; tI < 4; tI++)
		{
			const int j = __popc(sI ^ tI);
			const double w = (j == 0) ? w0 : ((j == 1) ? w1 : w2);
			sum = __dadd_rn(sum, __dmul_rn(__dmul_rn(a4[sI], b4[tI]), w));
		}
	return __dmul_rn(sum, K2);
}

__global__ void __launch_bounds__(128)
screen_refine_kernel(const ScreenArgs a)
{
	SmAcct acct_scope(a.acct, SM_ACCT_NEED, 64u);         // 128 threads: 16 CTAs fit an SM
	__shared__ int sh_keep, sh_ab[2];
	const int l = blockIdx.y, c = blockIdx.x;
	const int n = a.n_hla;
	const int n_cells = n * (n + 1) / 2;
	const int tid = threadIdx.x, lane = tid & 31;
	if (tid == 0)
	{
		sh_keep = 0;
		int row = 0, row0 = 0;                                // cell index -> (al, bl), bl >= al
		while (c >= row0 + (n - row)) { row0 += n - row; row++; }
		sh_ab[0] = row; sh_ab[1] = row + (c - row0);
	}
	__syncthreads();
	const int al = sh_ab[0], bl = sh_ab[1];
	const int cnt = a.count[(size_t)l * n_cells + c];
	const int *ent = a.entries + ((size_t)l * n_cells + c) * a.p_stride;
	int *ent2 = a.entries2 + ((size_t)l * n_cells + c) * a.p_stride;
	double *P = a.P + ((size_t)l * n_cells + c) * a.p_stride;
	const double *U2a = a.U2 + ((size_t)l * n + al) * 4 * a.p_stride;
	const double *U2b = a.U2 + ((size_t)l * n + bl) * 4 * a.p_stride;
	const int *hk = a.hetk + (size_t)l * a.p_stride;
	const double *xr = a.xref + (size_t)l * a.p_stride;
	for (int e0 = 0; e0 < cnt; e0 += 128)
	{
		const int e = e0 + tid;
		bool keep = false;
		int pos = 0;
		if (e < cnt)
		{
			pos = ent[e];
			const int hv = __ldg(hk + pos);
			const double thr = __dmul_rn(__ldg(xr + pos), a.tau);
			const double bd2 = screen_bound2(U2a + (size_t)pos * 4, U2b + (size_t)pos * 4, hv & 3, a.K2, a.tf);
			keep = (c == (hv >> 2)) || (bd2 >= thr && bd2 > 0.0);
			if (!keep) P[pos] = -((bd2 > 0.0) ? bd2 : 1e-300);      // (negative = skipped, with this bound)
		}
		const unsigned m = __ballot_sync(0xffffffffu, keep);
		int base = 0;
		if (lane == 0 && m) base = atomicAdd(&sh_keep, __popc(m));
		base = __shfl_sync(0xffffffffu, base, 0);
		if (keep) ent2[base + __popc(m & ((1u << lane) - 1u))] = pos;
	}
	__syncthreads();
	if (tid == 0) a.count2[(size_t)l * n_cells + c] = sh_keep;
}

void launch_screen_refine(const ScreenArgs &a, cudaStream_t st)
{
	if (a.n_pos <= 0 || a.n_lists <= 0 || a.U2 == nullptr) return;
	const int n_cells = a.n_hla * (a.n_hla + 1) / 2;
	dim3 grid(n_cells, a.n_lists);
	screen_refine_kernel<<<grid, 128, 0, st>>>(a);
	CUDA_CHECK(cudaGetLastError());
}

// ---------------------------------------------------------------------------------------
// screened reductions
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(64)
reduce_oob_screened_kernel(const ScreenArgs a, int *out_count)
{
	SmAcct acct_scope(a.acct, SM_ACCT_REDUCE_OOB, 32u);  // 64 threads: 32 CTAs fit an SM
	const int l = blockIdx.y;
	bool ok;
	const int pos = screen_pos(a, l, blockIdx.x * blockDim.x + threadIdx.x, ok);
	const int n = a.n_hla;
	const int n_cells = n * (n + 1) / 2;
	int cnt = 0;
	extern __shared__ double sU[];
	const bool us = a.u_smem != 0;
	if (us)
	{
		const double *Ug = a.U + (size_t)l * n * a.p_stride + pos;
		for (int al = 0; al < n; al++) sU[al * 64 + threadIdx.x] = ok ? Ug[(size_t)al * a.p_stride] : 0.0;
		__syncthreads();
	}
	if (!__any_sync(0xffffffffu, ok)) return;
	if (ok)
	{
		const double *U = a.U + (size_t)l * n * a.p_stride + pos;
		auto Uat = [&](int al) -> double { return us ? sU[al * 64 + threadIdx.x] : U[(size_t)al * a.p_stride]; };
		const double *P = a.P + (size_t)l * n_cells * a.p_stride + pos;
		const int samp = a.samp_list ? __ldg(a.samp_list + pos) : pos;
		int t1 = __ldg(a.a1 + samp), t2 = __ldg(a.a2 + samp);
		const int true_idx = true_cell_index(t1, t2, n);
		const double thr = __dmul_rn(a.xref[(size_t)l * a.p_stride + pos], a.tau);
		// strict '<' scan in cell order over the cells that were scored; a skipped cell is
		// strictly below x_ref <= the true cell's value and can never be the first maximum
		double best = 0.0;
		int p1 = NA_INT, p2 = NA_INT;
		int idx = 0;
		for (int al = 0; al < n; al++)
		{
			const double ua = Uat(al);
			for (int bl = al; bl < n; bl += 8, idx += 8)       // 8 cells per step: the loads overlap
			{
				const int nb = min(8, n - bl);
				double v[8];
				bool nd[8];
#pragma unroll
				for (int q = 0; q < 8; q++)
				{
					const double ub = (q < nb) ? Uat(bl + q) : 0.0;
					const double bd = screen_bound(ua, ub, a.K);
					nd[q] = (q < nb) && ((idx + q) == true_idx || (bd >= thr && bd > 0.0));
				}
#pragma unroll
				for (int q = 0; q < 8; q++) v[q] = nd[q] ? P[(size_t)(idx + q) * a.p_stride] : 0.0;
#pragma unroll
				for (int q = 0; q < 8; q++)
					if (nd[q] && best < v[q]) { best = v[q]; p1 = al; p2 = bl + q; }
			}
			idx -= (8 - ((n - al) & 7)) & 7;                    // undo the overshoot of the last step
		}
		// CHLATypeList::Compare (src/LibHLA.cpp:912-924)
		if (p1 == t1) { cnt = 1; t1 = -1; }
		else if (p1 == t2) { cnt = 1; t2 = -1; }
		if (p2 == t1 || p2 == t2) cnt++;
	}
	if (a.rep != nullptr)
	{
		// position classes: kept per representative, summed over ALL positions by the broadcast kernel
		if (ok) a.pos_res[(size_t)l * a.p_stride + pos] = cnt;
		return;
	}
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
	if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(out_count + l, cnt);
}

void launch_reduce_oob_screened(const ScreenArgs &a, int *out_count, cudaStream_t st)
{
	if (a.n_pos <= 0 || a.n_lists <= 0) return;
	dim3 grid((a.n_pos + 63) / 64, a.n_lists);
	const size_t smem = a.u_smem ? sizeof(double) * 64 * (size_t)a.n_hla : 0;
	if (smem > 40 * 1024)
		CUDA_CHECK(cudaFuncSetAttribute(reduce_oob_screened_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
	reduce_oob_screened_kernel<<<grid, 64, smem, st>>>(a, out_count);
	CUDA_CHECK(cudaGetLastError());
}

/// The reference's chain of one cell for one sample by ONE lane, everything read from global
/// memory: the rescue path of the in-bag reduction (rare, so simple beats fast). Same operations
/// in the same order as cell_chain below.
__device__ double rescue_chain(const char *__restrict__ hap, int nw, const double *__restrict__ table,
	int dmax, int a_start, int a_n, int b_start, int b_n, bool diag, const uint32_t (&S1)[4],
	const uint32_t (&S2)[4])
{
	const int rec = (nw <= 2) ? 16 : 32;
	double sum = 0.0;
	for (int ii = 0; ii < a_n; ii++)
	{
		const uint32_t *ri = (const uint32_t *)(hap + (size_t)(a_start + ii) * rec);
		uint32_t hi[4], K[4];
		int ci = 0;
		for (int w = 0; w < 4; w++)
		{
			hi[w] = (w < nw) ? __ldg(ri + w) : 0u;
			K[w] = S1[w] & (S2[w] | ~hi[w]);
			ci += __popc((hi[w] ^ (S1[w] & S2[w])) & ~(S1[w] ^ S2[w]));
		}
		const double fi = __ldg((const double *)((const char *)ri + ((nw <= 2) ? 8 : 16)));
		int j0 = 0;
		if (diag)
		{
			int pc = 0;
			for (int w = 0; w < 4; w++) pc += __popc((hi[w] ^ K[w]) & (S1[w] | ~S2[w]));
			sum = __dadd_rn(sum, __dmul_rn(__dmul_rn(fi, fi), __ldg(table + min(ci + pc, dmax))));
			j0 = ii + 1;
		}
		const double ff = __dmul_rn(2.0, fi);
		for (int j = j0; j < b_n; j++)
		{
			const uint32_t *rj = (const uint32_t *)(hap + (size_t)(b_start + j) * rec);
			int pc = 0;
			for (int w = 0; w < 4; w++)
			{
				const uint32_t hj = (w < nw) ? __ldg(rj + w) : 0u;
				pc += __popc((hj ^ K[w]) & (S1[w] | ~S2[w]));
			}
			const double fj = __ldg((const double *)((const char *)rj + ((nw <= 2) ? 8 : 16)));
			sum = __dadd_rn(sum, __dmul_rn(__dmul_rn(ff, fj), __ldg(table + min(ci + pc, dmax))));
		}
	}
	return sum;
}

__global__ void __launch_bounds__(64)
reduce_ib_screened_kernel(const ScreenArgs a, const __grid_constant__ ScreenLists ls,
	double *out_ratio, size_t out_stride)
{
	SmAcct acct_scope(a.acct, SM_ACCT_REDUCE_IB, 32u);
	const int l = blockIdx.y;
	bool ok;
	const int pos = screen_pos(a, l, blockIdx.x * blockDim.x + threadIdx.x, ok);
	const int lane = threadIdx.x & 31;
	const int n = a.n_hla;
	const int n_cells = n * (n + 1) / 2;
	const double *U = a.U + (size_t)l * n * a.p_stride;
	double *P = a.P + (size_t)l * n_cells * a.p_stride;
	int samp = 0, true_idx = 0;
	double x_true = 0.0, thr = 0.0, lo = 0.0, hi = 0.0;
	extern __shared__ double sU[];
	const bool us = a.u_smem != 0;
	if (us)
	{
		for (int al = 0; al < n; al++) sU[al * 64 + threadIdx.x] = ok ? U[(size_t)al * a.p_stride + pos] : 0.0;
		__syncthreads();
	}
	if (!__any_sync(0xffffffffu, ok)) return;             // (warp-level synchronisation only from here on)
	if (ok)
	{
		samp = a.samp_list ? __ldg(a.samp_list + pos) : pos;
		true_idx = true_cell_index(__ldg(a.a1 + samp), __ldg(a.a2 + samp), n);
		x_true = P[(size_t)true_idx * a.p_stride + pos];
		thr = __dmul_rn(a.xref[(size_t)l * a.p_stride + pos], a.tau);
		// the sequential sum in cell order, once with 0 and once with the bound for every skipped
		// cell: addition is monotone, so lo <= (the full chain) <= hi, and lo == hi certifies it
		const double *Up = U + pos;
		const double *Pp = P + pos;
		auto Uat = [&](int al) -> double { return us ? sU[al * 64 + threadIdx.x] : Up[(size_t)al * a.p_stride]; };
		int idx = 0;
		for (int al = 0; al < n; al++)
		{
			const double ua = Uat(al);
			for (int bl = al; bl < n; bl += 8, idx += 8)           // 8 cells per step: the loads overlap
			{
				const int nb = min(8, n - bl);
				double v[8], bd[8];
				bool nd[8];
#pragma unroll
				for (int q = 0; q < 8; q++)
				{
					const double ub = (q < nb) ? Uat(bl + q) : 0.0;
					bd[q] = screen_bound(ua, ub, a.K);
					nd[q] = (q < nb) && ((idx + q) == true_idx || (bd[q] >= thr && bd[q] > 0.0));
				}
#pragma unroll
				for (int q = 0; q < 8; q++) v[q] = nd[q] ? Pp[(size_t)(idx + q) * a.p_stride] : 0.0;
#pragma unroll
				for (int q = 0; q < 8; q++)
				{
					if (nd[q])
					{
						if (v[q] < 0.0)
							hi = __dadd_rn(hi, -v[q]);         // skipped by the second level: minus its bound
						else
						{
							lo = __dadd_rn(lo, v[q]);
							hi = __dadd_rn(hi, v[q]);
						}
					} else if (q < nb)
						hi = __dadd_rn(hi, bd[q]);
				}
			}
			idx -= (8 - ((n - al) & 7)) & 7;                        // undo the overshoot of the last step
		}
	}
	// ---- rescue: a position whose sum is not certified gets its skipped cells scored here, the
	// 32 lanes of its warp taking the cells in turn, and then the plain sequential sum -----------
	bool failed = ok && !(lo == hi);
	if (a.force_rescue && ok && ((pos + l) % a.force_rescue) == 0) failed = true;       // test hook
	if (failed && !a.device_rescue && !a.force_rescue) { lo = 0.0; hi = 1.0; }             // -> ratio -1: the caller rescores
	unsigned fm = __ballot_sync(0xffffffffu, failed && (a.device_rescue || a.force_rescue));
	if (fm)
	{
		const ScreenList &L = ls.l[l];
		const int nw = geno_words_dev(a.n_snp);
		const int dmax = a.n_dist - 1;
		const int *al_tab = a.al_tab + (size_t)l * n * 2;
		while (fm)
		{
			const int src = __ffs(fm) - 1;
			fm &= fm - 1;
			const int fpos = __shfl_sync(0xffffffffu, pos, src);
			const int fsamp = __shfl_sync(0xffffffffu, samp, src);
			const int ftrue = __shfl_sync(0xffffffffu, true_idx, src);
			const double fthr = __shfl_sync(0xffffffffu, thr, src);
			uint32_t S1[4], S2[4];
			load_geno<4>(a.s1, a.s2, a.geno_stride, fsamp, true, L.cand_col, L.cand_bit, S1, S2);
			for (int w = nw; w < 4; w++) { S1[w] = 0u; S2[w] = 0xffffffffu; }
			int row = 0, row0 = 0;                               // row `row` starts at cell row0
			for (int idx = lane; idx < n_cells; idx += 32)
			{
				while (idx >= row0 + (n - row)) { row0 += n - row; row++; }
				const int al = row, bl = row + (idx - row0);
				const double bd = screen_bound(U[(size_t)al * a.p_stride + fpos],
					U[(size_t)bl * a.p_stride + fpos], a.K);
				// already scored? (a cell the second level skipped holds minus its bound)
				if ((idx == ftrue || (bd >= fthr && bd > 0.0)) && !(P[(size_t)idx * a.p_stride + fpos] < 0.0)) continue;
				double x = 0.0;
				if (bd > 0.0)
					x = rescue_chain((const char *)L.hap, nw, a.table, dmax, al_tab[2 * al], al_tab[2 * al + 1],
						al_tab[2 * bl], al_tab[2 * bl + 1], al == bl, S1, S2);
				P[(size_t)idx * a.p_stride + fpos] = x;
			}
			__syncwarp();
			if (lane == src)
			{
				double s = 0.0;
				int idx = 0;
				for (; idx + 8 <= n_cells; idx += 8)
				{
					double v[8];
#pragma unroll
					for (int q = 0; q < 8; q++) v[q] = P[(size_t)(idx + q) * a.p_stride + pos];
#pragma unroll
					for (int q = 0; q < 8; q++) s = __dadd_rn(s, v[q]);
				}
				for (; idx < n_cells; idx++) s = __dadd_rn(s, P[(size_t)idx * a.p_stride + pos]);
				lo = hi = s;
				atomicAdd(a.rescued, 1ull);
			}
			__syncwarp();
		}
	}
	if (ok) out_ratio[(size_t)l * out_stride + pos] = (lo == hi) ? __ddiv_rn(x_true, lo) : -1.0;
}

void launch_reduce_ib_screened(const ScreenArgs &a, const ScreenLists &ls, double *out_ratio,
	size_t out_stride, cudaStream_t st)
{
	if (a.n_pos <= 0 || a.n_lists <= 0) return;
	dim3 grid((a.n_pos + 63) / 64, a.n_lists);
	const size_t smem = a.u_smem ? sizeof(double) * 64 * (size_t)a.n_hla : 0;
	if (smem > 40 * 1024)
		CUDA_CHECK(cudaFuncSetAttribute(reduce_ib_screened_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
	reduce_ib_screened_kernel<<<grid, 64, smem, st>>>(a, ls, out_ratio, out_stride);
	CUDA_CHECK(cudaGetLastError());
}

// ---------------------------------------------------------------------------------------
// the gather form of the pair-scoring kernel
// ---------------------------------------------------------------------------------------

/// the reference's chain of one cell for R samples per lane (same instruction sequence as the
/// body of cell_pass_kernel, kernels.cu)
template <int NW, int R, bool CLAMP, bool SMEM>
__device__ __forceinline__ void cell_chain(uint32_t hap_base, const char *hap_g,
	uint32_t tbl_lane, int dmax, int a_start, int a_n, int b_start, int b_n, bool diag,
	const uint32_t (*S1)[NW], const uint32_t (*S2)[NW], double *sum)
{
	// warp-uniform by construction (the task is broadcast from lane 0); the shuffles let the compiler
	// keep the loop bounds and the record addresses in uniform registers, as in cell_pass_kernel
	a_start = __shfl_sync(0xffffffffu, a_start, 0); a_n = __shfl_sync(0xffffffffu, a_n, 0);
	b_start = __shfl_sync(0xffffffffu, b_start, 0); b_n = __shfl_sync(0xffffffffu, b_n, 0);
	uint32_t V[R][NW];
#pragma unroll
	for (int r = 0; r < R; r++)
	{
		sum[r] = 0.0;
#pragma unroll
		for (int w = 0; w < NW; w++) V[r][w] = S1[r][w] | ~S2[r][w];
	}
	for (int ii = 0; ii < a_n; ii++)
	{
		HapRec<NW, SMEM> hi;
		hi.load(hap_base, hap_g, a_start + ii);
		uint32_t K[R][NW];
		uint32_t tb[R];
		int ci[R];
#pragma unroll
		for (int r = 0; r < R; r++)
		{
			int c0 = 0;
#pragma unroll
			for (int w = 0; w < NW; w++)
			{
				K[r][w] = S1[r][w] & (S2[r][w] | ~hi.h[w]);
				c0 += __popc((hi.h[w] ^ (S1[r][w] & S2[r][w])) & ~(S1[r][w] ^ S2[r][w]));
			}
			ci[r] = c0;
			tb[r] = tbl_lane + (uint32_t)c0 * 256u;
		}
		int j0 = 0;
		if (diag)
		{
			const double p2 = __dmul_rn(hi.f, hi.f);          // src/LibHLA.cpp:1658-1659
#pragma unroll
			for (int r = 0; r < R; r++)
			{
				int pc = 0;
#pragma unroll
				for (int w = 0; w < NW; w++) pc += __popc((hi.h[w] ^ K[r][w]) & V[r][w]);
				double t;
				if (CLAMP) t = lds_f64(tbl_lane + (uint32_t)min(ci[r] + pc, dmax) * 256u);
				else t = lds_f64(table_row(tb[r], pc));
				sum[r] = __dadd_rn(sum[r], __dmul_rn(p2, t));
			}
			j0 = ii + 1;
		}
		const double ff = __dmul_rn(2.0, hi.f);               // exact
		partner_loop<NW, R, CLAMP, SMEM>(hap_base, hap_g, b_start, j0, b_n, ff, K, V, ci, tb, tbl_lane, dmax, sum);
	}
}

template <int NW, bool CLAMP, bool SMEM>
__global__ void __launch_bounds__(GATHER_THREADS)
cell_gather_kernel(const __grid_constant__ GatherBatch p)
{
	extern __shared__ __align__(128) unsigned char smem_raw[];
	// [0,16): mbarrier | [16,24): flags | 128: table n_dist x 32 lanes x 8 B | task prefix | records
	const uint32_t smem_base = smem_u32(smem_raw);
	const uint32_t bar = smem_base;
	volatile int *sh_flag = (volatile int *)(smem_raw + 16);
	const uint32_t tbl_base = smem_base + 128;
	const uint32_t pre_off = 128u + (uint32_t)p.n_dist * 256u;
	const uint32_t pre_bytes = (((uint32_t)p.n_cells + 2u) * 4u + 15u) & ~15u;
	unsigned int *sh_pre = (unsigned int *)(smem_raw + pre_off);
	const uint32_t hap_base = smem_base + pre_off + pre_bytes;
	constexpr int REC = (NW <= 2) ? 16 : 32;

	const int tid = threadIdx.x;
	const int lane = tid & 31;
	SmAcct acct_scope(p.acct, p.acct_cls, (unsigned)p.acct_w);

	if (SMEM && tid == 0)
	{
		mbar_init(bar, 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	{
		double *tbl = (double *)(smem_raw + 128);
		const int n = p.n_dist * 32;
		for (int k = tid; k < n; k += GATHER_THREADS)
			tbl[k] = __ldg(p.table + (k >> 5));
	}
	__syncthreads();

	const uint32_t tbl_lane = tbl_base + lane * 8;
	const int dmax = p.n_dist - 1;
	uint32_t phase = 0;

	// p.flat & 2: a CTA serves ONE list (blockIdx mod n_lists; the grid is a multiple of n_lists). The
	// lists of an out-of-bag launch hold a few hundred tasks of very unequal length each, and a CTA that
	// walks the lists in turn waits at every list's barrier for its slowest warp (24 % of the stall
	// samples of such a launch)
	const int n_visit = (p.flat & 2) ? 1 : p.n_lists;
	for (int k = 0; k < n_visit; k++)
	{
		int l = (int)(blockIdx.x % (unsigned)p.n_lists) + k;
		if (l >= p.n_lists) l -= p.n_lists;
		const GatherList &L = p.lists[l];
		const unsigned n_tasks = __ldg(L.task_prefix + p.n_cells);
		unsigned int *counter = p.task_counters + l;

		if (tid == 0) sh_flag[k & 1] = (*(volatile unsigned int *)counter < n_tasks) ? 1 : 0;
		__syncthreads();
		if (!sh_flag[k & 1]) continue;

		if (SMEM && tid == 0)
		{
			const uint32_t total = (uint32_t)L.n_hap * REC;
			mbar_expect_tx(bar, total);
			const char *src = (const char *)L.hap;
			uint32_t off = 0;
			while (off < total)
			{
				uint32_t nb = total - off;
				if (nb > 65536u) nb = 65536u;
				tma_bulk_g2s(hap_base + off, src + off, nb, bar);
				off += nb;
			}
		}
		for (int q = tid; q < p.n_cells + 2; q += GATHER_THREADS) sh_pre[q] = __ldg(L.task_prefix + q);
		__syncthreads();
		if (SMEM)
		{
			mbar_wait(bar, phase);
			phase ^= 1u;
		}

		const char *hap_g = (const char *)L.hap;
		double *Pl = L.P;

		unsigned task = 0;
		if (lane == 0) task = atomicAdd(counter, 1u);
		task = __shfl_sync(0xffffffffu, task, 0);

		while (task < n_tasks)
		{
			unsigned next_task = 0;
			if (lane == 0) next_task = atomicAdd(counter, 1u);

			// largest q with prefix[q] <= task (cells without positions have empty ranges)
			int lo = 0, hi = p.n_cells;
			while (hi - lo > 1)
			{
				const int mid = (lo + hi) >> 1;
				if ((sh_pre[mid] & 0x3fffffffu) <= task) lo = mid; else hi = mid;
			}
			const unsigned int pk = sh_pre[lo];
			const int blk = (int)(task - (pk & 0x3fffffffu));
			const int task_shift = 5 + (int)(pk >> 30);              // log2(positions per task of this cell)
			const int task_pos = 1 << task_shift;
			const int4 ca = __ldg((const int4 *)(L.cells + lo));
			const int2 cb = __ldg((const int2 *)((const char *)(L.cells + lo) + 16));
			const int rem = min(task_pos, __ldg(L.count + cb.x) - (blk << task_shift));
			const int nr = (rem + 31) >> 5;
			const int *ent = L.entries + (size_t)__ldg(p.ent_off + cb.x) + ((size_t)blk << task_shift);

			uint32_t S1[4][NW], S2[4][NW];
			int pos[4];
			double sum[4];
#pragma unroll
			for (int r = 0; r < 4; r++)
			{
				const int e = r * 32 + lane;
				const bool ok = e < rem;
				pos[r] = ok ? __ldg(ent + e) : -1;
				int samp = 0;
				if (ok) samp = p.samp_list ? __ldg(p.samp_list + pos[r]) : pos[r];
				load_geno<NW>(p.s1, p.s2, p.geno_stride, samp, ok, L.cand_col, L.cand_bit, S1[r], S2[r]);
			}
			const bool diag = cb.y != 0;
			switch (nr)
			{
			case 1: cell_chain<NW, 1, CLAMP, SMEM>(hap_base, hap_g, tbl_lane, dmax, ca.x, ca.y, ca.z, ca.w, diag, S1, S2, sum); break;
			case 2: cell_chain<NW, 2, CLAMP, SMEM>(hap_base, hap_g, tbl_lane, dmax, ca.x, ca.y, ca.z, ca.w, diag, S1, S2, sum); break;
			case 3: cell_chain<NW, 3, CLAMP, SMEM>(hap_base, hap_g, tbl_lane, dmax, ca.x, ca.y, ca.z, ca.w, diag, S1, S2, sum); break;
			default: cell_chain<NW, 4, CLAMP, SMEM>(hap_base, hap_g, tbl_lane, dmax, ca.x, ca.y, ca.z, ca.w, diag, S1, S2, sum); break;
			}
#pragma unroll
			for (int r = 0; r < 4; r++)
				if (r < nr && pos[r] >= 0)
					Pl[(size_t)cb.x * p.p_stride + pos[r]] = sum[r];

			task = __shfl_sync(0xffffffffu, next_task, 0);
		}
		// every warp is done with this list's records / prefix before the next list overwrites them
		__syncthreads();
	}
}


// ---------------------------------------------------------------------------------------
// cell_gather_flat_kernel -- the entry-flat form for passes whose cells are needed by a handful of
// positions each (out-of-bag: ~2.5 of 595 cells per sample survive, ~8 positions per cell). In
// cell_gather_kernel such a (cell, positions) task fills a quarter of its warp and runs one dependent
// fp64 chain per lane: the launch is latency-bound (10 % of the POPC peak alone on the GPU). Here a
// task is 32 consecutive (cell, position) ENTRIES of the cost-sorted cell order and every lane walks
// the chain of ITS entry -- its own cell bounds and record addresses (plain LDS.128 instead of the
// uniform-datapath broadcast), its own (i, j) position. The lanes of a warp move in lock step
// through the inner loop for as many steps as the lane closest to the end of its row has left, then
// those lanes fetch their next row. The chain is cell_chain's, operation for operation.
// (Measured, config 2: no faster than the (cell, positions) form -- the lanes of a warp lie in
// several cells whose rows end at different steps, so the lock-step runs are short; a fully
// per-lane (i, j) iterator that re-reads the row record with every pair is 2.5 x slower, bound by
// the bank conflicts of two scattered LDS.128 per pair. Kept behind HIBAG_B200_GATHER_FLAT.)
// ---------------------------------------------------------------------------------------
template <int NW, bool CLAMP, bool SMEM>
__global__ void __launch_bounds__(GATHER_THREADS)
cell_gather_flat_kernel(const __grid_constant__ GatherBatch p)
{
	extern __shared__ __align__(128) unsigned char smem_raw[];
	// same layout as cell_gather_kernel
	const uint32_t smem_base = smem_u32(smem_raw);
	const uint32_t bar = smem_base;
	volatile int *sh_flag = (volatile int *)(smem_raw + 16);
	const uint32_t tbl_base = smem_base + 128;
	const uint32_t pre_off = 128u + (uint32_t)p.n_dist * 256u;
	const uint32_t pre_bytes = (((uint32_t)p.n_cells + 2u) * 4u + 15u) & ~15u;
	unsigned int *sh_pre = (unsigned int *)(smem_raw + pre_off);
	const uint32_t hap_base = smem_base + pre_off + pre_bytes;
	constexpr int REC = (NW <= 2) ? 16 : 32;

	const int tid = threadIdx.x;
	const int lane = tid & 31;
	SmAcct acct_scope(p.acct, p.acct_cls, (unsigned)p.acct_w);

	if (SMEM && tid == 0)
	{
		mbar_init(bar, 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	{
		double *tbl = (double *)(smem_raw + 128);
		const int n = p.n_dist * 32;
		for (int k = tid; k < n; k += GATHER_THREADS)
			tbl[k] = __ldg(p.table + (k >> 5));
	}
	__syncthreads();

	const uint32_t tbl_lane = tbl_base + lane * 8;
	const int dmax = p.n_dist - 1;
	uint32_t phase = 0;

	// a CTA serves ONE list (blockIdx mod n_lists; the grid is a multiple of n_lists): the lists of an
	// out-of-bag launch hold a few hundred tasks of very unequal length each, and a CTA that walked the
	// lists in turn would wait at every list's barrier for its slowest warp (24 % of the stall samples
	// of such a launch, profiles/r02_gather_oob_before.txt)
	{
		const int l = (int)(blockIdx.x % (unsigned)p.n_lists);
		const GatherList &L = p.lists[l];
		const unsigned n_tasks = __ldg(L.task_prefix + p.n_cells);
		unsigned int *counter = p.task_counters + l;
		if (tid == 0) *sh_flag = (*(volatile unsigned int *)counter < n_tasks) ? 1 : 0;
		__syncthreads();
		if (!*sh_flag) return;                                        // (CTA-uniform)

		if (SMEM && tid == 0)
		{
			const uint32_t total = (uint32_t)L.n_hap * REC;
			mbar_expect_tx(bar, total);
			const char *src = (const char *)L.hap;
			uint32_t off = 0;
			while (off < total)
			{
				uint32_t nb = total - off;
				if (nb > 65536u) nb = 65536u;
				tma_bulk_g2s(hap_base + off, src + off, nb, bar);
				off += nb;
			}
		}
		for (int q = tid; q < p.n_cells + 2; q += GATHER_THREADS) sh_pre[q] = __ldg(L.task_prefix + q);
		__syncthreads();
		if (SMEM)
		{
			mbar_wait(bar, phase);
			phase ^= 1u;
		}

		const char *hap_g = (const char *)L.hap;
		double *Pl = L.P;
		const unsigned n_entries = sh_pre[p.n_cells + 1];

		unsigned task = 0;
		if (lane == 0) task = atomicAdd(counter, 1u);
		task = __shfl_sync(0xffffffffu, task, 0);

		while (task < n_tasks)
		{
			unsigned next_task = 0;
			if (lane == 0) next_task = atomicAdd(counter, 1u);

			const unsigned f = task * 32u + (unsigned)lane;
			bool act = f < n_entries;
			// this lane's cell: the largest q with prefix[q] <= f (cells nobody needs have empty ranges)
			int lo = 0, hi = p.n_cells;
			while (hi - lo > 1)
			{
				const int mid = (lo + hi) >> 1;
				if (sh_pre[mid] <= f) lo = mid; else hi = mid;
			}
			const int4 ca = __ldg((const int4 *)(L.cells + lo));
			const int2 cb = __ldg((const int2 *)((const char *)(L.cells + lo) + 16));
			int pos = -1, samp = 0;
			if (act)
			{
				pos = __ldg(L.entries + (size_t)__ldg(p.ent_off + cb.x) + (f - sh_pre[lo]));
				samp = p.samp_list ? __ldg(p.samp_list + pos) : pos;
			}
			uint32_t S1[NW], S2[NW], V[NW], K[NW];
			load_geno<NW>(p.s1, p.s2, p.geno_stride, samp, act, L.cand_col, L.cand_bit, S1, S2);
#pragma unroll
			for (int w = 0; w < NW; w++) V[w] = S1[w] | ~S2[w];
			const bool diag = cb.y != 0;
			const int a_end = act ? ca.x + ca.y : ca.x;         // records [ca.x, a_end) x [ca.z, ca.z + ca.w)
			int ri = ca.x;                                      // next row
			int rj = 0, rj_end = 0;                             // current position / end in the row of partners
			double sum = 0.0, ff = 0.0;
			uint32_t tb = tbl_lane;
			int ci = 0;
			for (;;)
			{
				if (act && rj >= rj_end)
				{
					if (ri >= a_end) act = false;
					else
					{
						HapRec<NW, SMEM> hi_;
						hi_.load(hap_base, hap_g, ri);
						int c0 = 0;
#pragma unroll
						for (int w = 0; w < NW; w++)
						{
							K[w] = S1[w] & (S2[w] | ~hi_.h[w]);
							c0 += __popc((hi_.h[w] ^ (S1[w] & S2[w])) & ~(S1[w] ^ S2[w]));
						}
						ci = c0;
						tb = tbl_lane + (uint32_t)c0 * 256u;
						rj = ca.z; rj_end = ca.z + ca.w;
						if (diag)
						{
							const double p2 = __dmul_rn(hi_.f, hi_.f);          // src/LibHLA.cpp:1658-1659
							int pc = 0;
#pragma unroll
							for (int w = 0; w < NW; w++) pc += __popc((hi_.h[w] ^ K[w]) & V[w]);
							double t;
							if (CLAMP) t = lds_f64(tbl_lane + (uint32_t)min(ci + pc, dmax) * 256u);
							else t = lds_f64(table_row(tb, pc));
							sum = __dadd_rn(sum, __dmul_rn(p2, t));
							rj = ri + 1;                                        // (a diagonal cell: ca.z == ca.x)
						}
						ff = __dmul_rn(2.0, hi_.f);                            // exact
						ri++;
					}
				}
				if (!__any_sync(0xffffffffu, act)) break;
				const int n = __reduce_min_sync(0xffffffffu, act ? (rj_end - rj) : 0x7fffffff);
				if (act)
				{
					// the terms of the next four partners are fetched and multiplied before the adds of the
					// current four are issued: only the dependent adds stay on the critical path
					auto terms = [&](int j, double (&x)[4])
					{
						HapRec<NW, SMEM> hj[4];
#pragma unroll
						for (int q = 0; q < 4; q++) hj[q].load(hap_base, hap_g, j + q);
						uint32_t ad[4];
#pragma unroll
						for (int q = 0; q < 4; q++)
						{
							int pc = 0;
#pragma unroll
							for (int w = 0; w < NW; w++) pc += __popc((hj[q].h[w] ^ K[w]) & V[w]);
							ad[q] = CLAMP ? (tbl_lane + (uint32_t)min(ci + pc, dmax) * 256u) : table_row(tb, pc);
						}
#pragma unroll
						for (int q = 0; q < 4; q++) x[q] = lds_f64(ad[q]);
#pragma unroll
						for (int q = 0; q < 4; q++) x[q] = __dmul_rn(__dmul_rn(ff, hj[q].f), x[q]);
					};
					int t_ = 0;
					if (n >= 4)
					{
						double x0[4];
						terms(rj, x0);
						for (t_ = 4; t_ + 4 <= n; t_ += 4)
						{
							double x1[4];
							terms(rj + t_, x1);
#pragma unroll
							for (int q = 0; q < 4; q++) { sum = __dadd_rn(sum, x0[q]); x0[q] = x1[q]; }
						}
#pragma unroll
						for (int q = 0; q < 4; q++) sum = __dadd_rn(sum, x0[q]);
					}
					for (; t_ < n; t_++)
					{
						HapRec<NW, SMEM> hj;
						hj.load(hap_base, hap_g, rj + t_);
						const double pf = __dmul_rn(ff, hj.f);
						int pc = 0;
#pragma unroll
						for (int w = 0; w < NW; w++) pc += __popc((hj.h[w] ^ K[w]) & V[w]);
						double t;
						if (CLAMP) t = lds_f64(tbl_lane + (uint32_t)min(ci + pc, dmax) * 256u);
						else t = lds_f64(table_row(tb, pc));
						sum = __dadd_rn(sum, __dmul_rn(pf, t));
					}
					rj += n;
				}
			}
			if (pos >= 0) Pl[(size_t)cb.x * p.p_stride + pos] = sum;

			task = __shfl_sync(0xffffffffu, next_task, 0);
		}
	}
}

template <int NW, bool CLAMP>
static void launch_gather_variant(const GatherBatch &p, int sm_count, cudaStream_t st, long long max_ctas)
{
	const size_t rec = (NW <= 2) ? 16 : 32;
	const size_t fixed = 128 + (size_t)p.n_dist * 256 + ((((size_t)p.n_cells + 2) * 4 + 15) & ~(size_t)15);
	const size_t with_hap = fixed + (size_t)p.max_hap * rec;
	const size_t smem_limit = 227 * 1024;
	if (fixed > smem_limit) throw std::runtime_error("launch_cell_gather: too many cells for shared memory");
	const bool in_smem = with_hap <= smem_limit;
	const size_t smem = in_smem ? with_hap : fixed;
	int cta_per_sm = (int)((228 * 1024) / (smem + 1024));
	if (cta_per_sm > 8) cta_per_sm = 8;
	if (cta_per_sm < 1) cta_per_sm = 1;
	long long grid = (long long)sm_count * cta_per_sm;
	if (max_ctas > 0 && grid > max_ctas) grid = max_ctas;
	if (grid < p.n_lists) grid = p.n_lists;
	if (p.flat) grid = ((grid + p.n_lists - 1) / p.n_lists) * p.n_lists;   // every list the same number of CTAs
	GatherBatch q = p;
	q.acct_w = 1024 / cta_per_sm;
	if (p.flat & 1)
	{
		if (in_smem)
		{
			auto k = cell_gather_flat_kernel<NW, CLAMP, true>;
			CUDA_CHECK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_limit));
			k<<<(unsigned)grid, GATHER_THREADS, smem, st>>>(q);
		} else {
			auto k = cell_gather_flat_kernel<NW, CLAMP, false>;
			CUDA_CHECK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_limit));
			k<<<(unsigned)grid, GATHER_THREADS, smem, st>>>(q);
		}
	} else if (in_smem)
	{
		auto k = cell_gather_kernel<NW, CLAMP, true>;
		CUDA_CHECK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_limit));
		k<<<(unsigned)grid, GATHER_THREADS, smem, st>>>(q);
	} else {
		auto k = cell_gather_kernel<NW, CLAMP, false>;
		CUDA_CHECK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_limit));
		k<<<(unsigned)grid, GATHER_THREADS, smem, st>>>(q);
	}
	CUDA_CHECK(cudaGetLastError());
}

int launch_cell_gather(const GatherBatch &p, int sm_count, cudaStream_t st, long long max_ctas)
{
	if (p.n_lists < 1 || p.n_lists > MAX_BATCH_LISTS)
		throw std::runtime_error("launch_cell_gather: invalid number of lists");
	const int nw = geno_words(p.n_snp);
	const bool clamp = (2 * p.n_snp) > (p.n_dist - 1);
#define HB_GCASE(NW_) \
	if (nw == NW_) { \
		if (clamp) launch_gather_variant<NW_, true>(p, sm_count, st, max_ctas); \
		else launch_gather_variant<NW_, false>(p, sm_count, st, max_ctas); \
		return NW_; }
	HB_GCASE(1) HB_GCASE(2) HB_GCASE(4)
#undef HB_GCASE
	throw std::runtime_error("launch_cell_gather: unsupported configuration");
}

}  // namespace hb
