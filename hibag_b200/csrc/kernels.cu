// kernels.cu -- hand-written sm_100a kernels of the haplotype-pair scoring path.
//
// What is computed (reference src/LibHLA.cpp:1639-1830, macro ADD_FREQ_MUTANT LibHLA.h:222):
// for a genotype g and every unordered HLA allele pair ("cell") (a,b)
//     P(a,b) = sum_{i in a} sum_{j in b, j>=i if a==b}  (c_ij * f_i * f_j) * T[d(g,i,j)]
// as ONE sequential fp64 chain per (sample, cell) in (i outer, j inner) order, un-fused
// multiply then add -- the exact operation order of the reference's base target, so that the
// integer decisions that hang off these sums (argmax, SNP selection) are bit-identical.
//
// Mapping: a lane owns R samples; the 32 lanes of a warp walk the same cell, so haplotype
// words and frequencies are warp-uniform shared-memory broadcasts (staged once per CTA by a
// TMA bulk copy) and only the genotype masks and running sums are per lane.
//
// Distance (reference hamm_d, src/LibHLA.cpp:802-817) is evaluated in the algebraically
// identical one-popcount form: with S1,S2 the genotype bit planes,
//     V   = S1 | ~S2                 non-missing SNPs
//     K_i = S1 & (S2 | ~h_i)         per (sample, i)
//     c_i = popc((h_i ^ (S1&S2)) & ~(S1^S2))          per (sample, i)
//     d(g,i,j) = c_i + popc((h_j ^ K_i) & V)
// (per SNP: g=0 -> h_i+h_j, g=2 -> (1-h_i)+(1-h_j), g=1 -> [h_i==h_j], missing -> 0), i.e. one
// LOP3 + one POPC per 32 SNPs per pair instead of the reference formulation's four POPC.
// T[] is replicated per lane in shared memory ([d][lane]) so the data-dependent lookup is
// bank-conflict free.

#include "kernels.h"

#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>

namespace hb {

#define CUDA_CHECK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) \
	throw std::runtime_error(std::string("CUDA error: ") + cudaGetErrorString(e_) + \
		" at " __FILE__ ":" + std::to_string(__LINE__)); } while (0)

static constexpr int CELL_THREADS = 128;
static constexpr int NA_INT = INT32_MIN;

}  // namespace hb

#include "devutil.cuh"

namespace hb {


template <int NW, int R, bool CLAMP, bool SMEM>
__global__ void __launch_bounds__(CELL_THREADS)
cell_pass_kernel(const __grid_constant__ CellBatch p)
{
	extern __shared__ __align__(128) unsigned char smem_raw[];
	// [0,16): mbarrier | [16,24): list-has-work flags | table: n_dist x 32 lanes x 8 B | records
	const uint32_t smem_base = smem_u32(smem_raw);
	const uint32_t bar = smem_base;
	volatile int *sh_flag = (volatile int *)(smem_raw + 16);
	const uint32_t tbl_base = smem_base + 128;
	const uint32_t hap_base = tbl_base + (uint32_t)p.n_dist * 256u;
	constexpr int REC = (NW <= 2) ? 16 : 32;

	const int tid = threadIdx.x;
	const int lane = tid & 31;
	SmAcct acct_scope(p.acct, SM_ACCT_CELL_PASS, (unsigned)p.acct_w);

	if (SMEM && tid == 0)
	{
		mbar_init(bar, 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	// lane-replicated rare-frequency table: tbl[d][lane]
	{
		double *tbl = (double *)(smem_raw + 128);
		const int n = p.n_dist * 32;
		for (int k = tid; k < n; k += CELL_THREADS)
			tbl[k] = __ldg(p.table + (k >> 5));
	}
	__syncthreads();

	const uint32_t tbl_lane = tbl_base + lane * 8;
	const int n_pos = p.n_pos_dev ? min(__ldg(p.n_pos_dev), p.n_pos) : p.n_pos;
	const int n_groups = (n_pos + 32 * R - 1) / (32 * R);
	const int dmax = p.n_dist - 1;
	uint32_t phase = 0;

	// A CTA starts on list (blockIdx mod n_lists) and moves on to the next list when the current
	// one has no tasks left, so all CTAs converge on whatever work remains.
	for (int k = 0; k < p.n_lists; k++)
	{
		int l = (int)(blockIdx.x % (unsigned)p.n_lists) + k;
		if (l >= p.n_lists) l -= p.n_lists;
		const ListDesc &L = p.lists[l];
		const unsigned n_tasks = (unsigned)n_groups * (unsigned)L.n_chunks;
		unsigned int *counter = p.task_counters + l;

		if (p.n_lists > 1)
		{
			if (tid == 0) sh_flag[k & 1] = (*(volatile unsigned int *)counter < n_tasks) ? 1 : 0;
			__syncthreads();
			if (!sh_flag[k & 1]) continue;
		}
		if (SMEM)
		{
			if (tid == 0)
			{
				const uint32_t total = (uint32_t)L.n_hap * REC;
				mbar_expect_tx(bar, total);
				const char *src = (const char *)L.hap;
				uint32_t off = 0;
				while (off < total)            // pieces of <= 64 KB, all multiples of 16 B
				{
					uint32_t n = total - off;
					if (n > 65536u) n = 65536u;
					tma_bulk_g2s(hap_base + off, src + off, n, bar);
					off += n;
				}
			}
			mbar_wait(bar, phase);
			phase ^= 1u;
		}

		const char *hap_g = (const char *)L.hap;
		const int8_t *cand_col = L.cand_col;
		const int cand_bit = L.cand_bit;
		double *Pl = L.P;

		unsigned task = 0;
		if (lane == 0) task = atomicAdd(counter, 1u);
		task = __shfl_sync(0xffffffffu, task, 0);

		while (task < n_tasks)
		{
			unsigned next_task = 0;
			if (lane == 0) next_task = atomicAdd(counter, 1u);

			const int chunk_id = task / n_groups;
			const int group = task - chunk_id * n_groups;

			// ---- genotype bit planes of this lane's R samples ---------------------------
			uint32_t S1[R][NW], S2[R][NW], V[R][NW];
			int pos[R];
#pragma unroll
			for (int r = 0; r < R; r++)
			{
				pos[r] = group * (32 * R) + r * 32 + lane;
				const bool ok = pos[r] < n_pos;
				int samp = 0;
				if (ok) samp = p.samp_list ? __ldg(p.samp_list + pos[r]) : pos[r];
#pragma unroll
				for (int w = 0; w < NW; w++)
				{
					S1[r][w] = ok ? __ldg(p.s1 + (size_t)w * p.geno_stride + samp) : 0u;
					S2[r][w] = ok ? __ldg(p.s2 + (size_t)w * p.geno_stride + samp) : 0xffffffffu;
				}
				if (cand_col != nullptr && ok)
				{
					// CGenotypeList::AddSNP of the candidate column (src/LibHLA.cpp:609-622, 860-874)
					const int g = __ldg(cand_col + samp);
					const int cw = cand_bit >> 5;
					const uint32_t bit = 1u << (cand_bit & 31);
#pragma unroll
					for (int w = 0; w < NW; w++)
					{
						if (w == cw)
						{
							if (g == 1 || g == 2) S1[r][w] |= bit; else S1[r][w] &= ~bit;
							if (g == 0 || g == 1) S2[r][w] &= ~bit; else S2[r][w] |= bit;
						}
					}
				}
#pragma unroll
				for (int w = 0; w < NW; w++) V[r][w] = S1[r][w] | ~S2[r][w];
			}

			const int2 ch = __ldg((const int2 *)L.chunks + chunk_id);    // cell_begin, cell_end

			for (int c = ch.x; c < ch.y; c++)
			{
				const int4 ca = __ldg((const int4 *)(L.cells + c));           // a_start,a_n,b_start,b_n
				const int2 cb = __ldg((const int2 *)((const char *)(L.cells + c) + 16));  // out_idx, diag
				const int a_start = ca.x, a_n = ca.y, b_start = ca.z, b_n = ca.w;
				const bool diag = cb.y != 0;

				double sum[R];
#pragma unroll
				for (int r = 0; r < R; r++) sum[r] = 0.0;

				for (int ii = 0; ii < a_n; ii++)
				{
					HapRec<NW, SMEM> hi;
					hi.load(hap_base, hap_g, a_start + ii);

					uint32_t K[R][NW];
					uint32_t tb[R];     // shared address of T[c_i][lane]   (no clamp)
					int ci[R];          // c_i                                 (clamp)
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
					double ff;
					if (diag)
					{
						// i2 == i1: (f*f) * T[d(i,i)]   (src/LibHLA.cpp:1658-1659)
						const double p2 = __dmul_rn(hi.f, hi.f);
#pragma unroll
						for (int r = 0; r < R; r++)
						{
							int pc = 0;
#pragma unroll
							for (int w = 0; w < NW; w++)
								pc += __popc((hi.h[w] ^ K[r][w]) & V[r][w]);
							double t;
							if (CLAMP) t = lds_f64(tbl_lane + (uint32_t)min(ci[r] + pc, dmax) * 256u);
							else t = lds_f64(tb[r] + (uint32_t)pc * 256u);
							sum[r] = __dadd_rn(sum[r], __dmul_rn(p2, t));
						}
						j0 = ii + 1;
					}
					ff = __dmul_rn(2.0, hi.f);    // exact

					partner_loop<NW, R, CLAMP, SMEM>(hap_base, hap_g, b_start, j0, b_n, ff, K, V, ci, tb, tbl_lane, dmax, sum);
				}

#pragma unroll
				for (int r = 0; r < R; r++)
					if (pos[r] < n_pos)
						Pl[(size_t)cb.x * p.p_stride + pos[r]] = sum[r];
			}

			task = __shfl_sync(0xffffffffu, next_task, 0);
		}
		// every warp is done with this list's records before the next list overwrites them
		if (p.n_lists > 1) __syncthreads();
	}
}

template <int NW, int R, bool CLAMP>
static void launch_cell_variant(const CellBatch &p, int sm_count, cudaStream_t st)
{
	const size_t rec = (NW <= 2) ? 16 : 32;
	const size_t fixed = 128 + (size_t)p.n_dist * 256;
	const size_t with_hap = fixed + (size_t)p.max_hap * rec;
	const size_t smem_limit = 227 * 1024;
	const bool in_smem = with_hap <= smem_limit;
	const size_t smem = in_smem ? with_hap : fixed;

	const int n_groups = (p.n_pos + 32 * R - 1) / (32 * R);
	long long n_tasks = 0;
	for (int l = 0; l < p.n_lists; l++) n_tasks += (long long)n_groups * p.lists[l].n_chunks;
	if (n_tasks <= 0) return;
	const int warps_per_cta = CELL_THREADS / 32;
	// persistent CTAs: as many as fit by shared memory (<= 8 per SM), never more than needed
	int cta_per_sm = (int)((228 * 1024) / (smem + 1024));
	if (cta_per_sm > 8) cta_per_sm = 8;
	if (cta_per_sm < 1) cta_per_sm = 1;
	long long grid = (long long)sm_count * cta_per_sm;
	const long long need = (n_tasks + warps_per_cta - 1) / warps_per_cta;
	if (grid > need) grid = need;
	if (grid < p.n_lists) grid = p.n_lists;

	CellBatch q = p;
	q.acct_w = 1024 / cta_per_sm;
	if (in_smem)
	{
		auto k = cell_pass_kernel<NW, R, CLAMP, true>;
		CUDA_CHECK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_limit));
		k<<<(unsigned)grid, CELL_THREADS, smem, st>>>(q);
	} else {
		auto k = cell_pass_kernel<NW, R, CLAMP, false>;
		CUDA_CHECK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_limit));
		k<<<(unsigned)grid, CELL_THREADS, smem, st>>>(q);
	}
	CUDA_CHECK(cudaGetLastError());
}

int launch_cell_batch(const CellBatch &p, int samples_per_lane, int sm_count, cudaStream_t st)
{
	if (p.n_lists < 1 || p.n_lists > MAX_BATCH_LISTS)
		throw std::runtime_error("launch_cell_batch: invalid number of lists");
	const int nw = geno_words(p.n_snp);
	const bool clamp = (2 * p.n_snp) > (p.n_dist - 1);
	int R = samples_per_lane;
	if (nw == 4 && R > 2) R = 2;
#define HB_CASE(NW_, R_) \
	if (nw == NW_ && R == R_) { \
		if (clamp) launch_cell_variant<NW_, R_, true>(p, sm_count, st); \
		else launch_cell_variant<NW_, R_, false>(p, sm_count, st); \
		return NW_; }
	HB_CASE(1, 1) HB_CASE(1, 2) HB_CASE(1, 4)
	HB_CASE(2, 1) HB_CASE(2, 2) HB_CASE(2, 4)
	HB_CASE(4, 1) HB_CASE(4, 2)
#undef HB_CASE
	throw std::runtime_error("launch_cell_batch: unsupported configuration");
}

int launch_cell_pass(const CellPass &p, int samples_per_lane, int sm_count, cudaStream_t st)
{
	CellBatch b;
	memset(&b, 0, sizeof(b));
	b.table = p.table; b.s1 = p.s1; b.s2 = p.s2; b.samp_list = p.samp_list;
	b.task_counters = p.task_counter; b.p_stride = p.p_stride;
	b.n_dist = p.n_dist; b.n_snp = p.n_snp; b.geno_stride = p.geno_stride; b.n_pos = p.n_pos;
	b.n_pos_dev = p.n_pos_dev;
	b.n_lists = 1; b.max_hap = p.n_hap;
	b.acct = device_sm_acct();
	ListDesc &L = b.lists[0];
	L.hap = p.hap; L.cells = p.cells; L.chunks = p.chunks; L.cand_col = p.cand_col; L.P = p.P;
	L.n_hap = p.n_hap; L.n_chunks = p.n_chunks; L.cand_bit = p.cand_bit;
	return launch_cell_batch(b, samples_per_lane, sm_count, st);
}

// ---------------------------------------------------------------------------------------
// genotype unpacking (legacy hook path): TGenotype AoS -> SoA words
// ---------------------------------------------------------------------------------------
__global__ void unpack_genotypes_kernel(const uint32_t *__restrict__ aos, int n,
	uint32_t *s1, uint32_t *s2, int stride, int *a1, int *a2, int *boot)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const uint32_t *g = aos + (size_t)i * 12;     // 48 bytes = 12 words
#pragma unroll
	for (int w = 0; w < 4; w++)
	{
		s1[(size_t)w * stride + i] = g[w];
		s2[(size_t)w * stride + i] = g[4 + w];
	}
	boot[i] = (int)g[8];
	a1[i] = (int)g[9];
	a2[i] = (int)g[10];
}

void launch_unpack_genotypes(const void *geno_aos, int n, uint32_t *s1, uint32_t *s2,
	int stride, int *a1, int *a2, int *boot, cudaStream_t st)
{
	if (n <= 0) return;
	unpack_genotypes_kernel<<<(n + 255) / 256, 256, 0, st>>>((const uint32_t *)geno_aos, n,
		s1, s2, stride, a1, a2, boot);
	CUDA_CHECK(cudaGetLastError());
}

/// one SNP column of the bit planes rewritten from 2-bit genotype codes (legacy hook path: between
/// two build_set_haplo_geno calls only the candidate SNP's column of TGenotype[] changes)
__global__ void patch_column_kernel(const int8_t *__restrict__ code, int n, uint32_t *s1,
	uint32_t *s2, int stride, int word, uint32_t bit)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const int c = code[i];              // bit 0 = PackedSNP1 bit, bit 1 = PackedSNP2 bit
	uint32_t a = s1[(size_t)word * stride + i], b = s2[(size_t)word * stride + i];
	a = (c & 1) ? (a | bit) : (a & ~bit);
	b = (c & 2) ? (b | bit) : (b & ~bit);
	s1[(size_t)word * stride + i] = a;
	s2[(size_t)word * stride + i] = b;
}

void launch_patch_column(const int8_t *code, int n, uint32_t *s1, uint32_t *s2, int stride,
	int snp_bit, cudaStream_t st)
{
	if (n <= 0) return;
	patch_column_kernel<<<(n + 255) / 256, 256, 0, st>>>(code, n, s1, s2, stride, snp_bit >> 5,
		1u << (snp_bit & 31));
	CUDA_CHECK(cudaGetLastError());
}

// ---------------------------------------------------------------------------------------
// reductions over the cell vector of one sample (all sequential in cell order)
// ---------------------------------------------------------------------------------------

/// cell index -> (h1, h2), h1 <= h2
__device__ __forceinline__ void cell_to_pair(int idx, int n_hla, int &h1, int &h2)
{
	int row = 0, len = n_hla;
	while (idx >= len) { idx -= len; len--; row++; }
	h1 = row; h2 = row + idx;
}

/// strict '<' scan, first maximum wins, nothing positive -> -1 (src/LibHLA.cpp:1549-1566)
__device__ __forceinline__ int argmax_cells(const double *P, size_t stride, int n_cells,
	double &best)
{
	best = 0.0;
	int bi = -1;
	int c = 0;
	for (; c + 4 <= n_cells; c += 4)
	{
		const double v0 = P[(size_t)c * stride], v1 = P[(size_t)(c + 1) * stride];
		const double v2 = P[(size_t)(c + 2) * stride], v3 = P[(size_t)(c + 3) * stride];
		if (best < v0) { best = v0; bi = c; }
		if (best < v1) { best = v1; bi = c + 1; }
		if (best < v2) { best = v2; bi = c + 2; }
		if (best < v3) { best = v3; bi = c + 3; }
	}
	for (; c < n_cells; c++)
	{
		const double v = P[(size_t)c * stride];
		if (best < v) { best = v; bi = c; }
	}
	return bi;
}

__device__ __forceinline__ double seqsum_cells(const double *P, size_t stride, int n_cells)
{
	double s = 0.0;
	int c = 0;
	for (; c + 4 <= n_cells; c += 4)
	{
		const double v0 = P[(size_t)c * stride], v1 = P[(size_t)(c + 1) * stride];
		const double v2 = P[(size_t)(c + 2) * stride], v3 = P[(size_t)(c + 3) * stride];
		s = __dadd_rn(s, v0); s = __dadd_rn(s, v1); s = __dadd_rn(s, v2); s = __dadd_rn(s, v3);
	}
	for (; c < n_cells; c++) s = __dadd_rn(s, P[(size_t)c * stride]);
	return s;
}

__global__ void reduce_oob_kernel(const double *__restrict__ P, size_t p_stride, size_t list_stride,
	int n_hla, const int *__restrict__ samp_list, int n_pos, const int *__restrict__ a1,
	const int *__restrict__ a2, int *out_count)
{
	const int pos = blockIdx.x * blockDim.x + threadIdx.x;
	P += (size_t)blockIdx.y * list_stride;        // one haplotype list per blockIdx.y
	out_count += blockIdx.y;
	int cnt = 0;
	if (pos < n_pos)
	{
		const int n_cells = n_hla * (n_hla + 1) / 2;
		double best;
		const int bi = argmax_cells(P + pos, p_stride, n_cells, best);
		int p1 = NA_INT, p2 = NA_INT;
		if (bi >= 0) cell_to_pair(bi, n_hla, p1, p2);
		const int samp = samp_list ? samp_list[pos] : pos;
		int t1 = a1[samp], t2 = a2[samp];
		// CHLATypeList::Compare (src/LibHLA.cpp:912-924)
		if (p1 == t1) { cnt = 1; t1 = -1; }
		else if (p1 == t2) { cnt = 1; t2 = -1; }
		if (p2 == t1 || p2 == t2) cnt++;
	}
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
	if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(out_count, cnt);
}

void launch_reduce_oob(const double *P, size_t p_stride, int n_hla, const int *samp_list,
	int n_pos, const int *a1, const int *a2, int *out_count, cudaStream_t st, int n_lists,
	size_t list_stride)
{
	if (n_pos <= 0 || n_lists <= 0) return;
	dim3 grid((n_pos + 63) / 64, n_lists);
	reduce_oob_kernel<<<grid, 64, 0, st>>>(P, p_stride, list_stride, n_hla, samp_list, n_pos,
		a1, a2, out_count);
	CUDA_CHECK(cudaGetLastError());
}

__global__ void reduce_ib_kernel(const double *__restrict__ P, size_t p_stride, size_t list_stride,
	int n_hla, const int *__restrict__ samp_list, int n_pos, const int *__restrict__ a1,
	const int *__restrict__ a2, double *out_ratio, size_t out_stride)
{
	const int pos = blockIdx.x * blockDim.x + threadIdx.x;
	if (pos >= n_pos) return;
	P += (size_t)blockIdx.y * list_stride;
	out_ratio += (size_t)blockIdx.y * out_stride;
	const int n_cells = n_hla * (n_hla + 1) / 2;
	const double s = seqsum_cells(P + pos, p_stride, n_cells);
	const int samp = samp_list ? samp_list[pos] : pos;
	int h1 = a1[samp], h2 = a2[samp];
	if (h1 > h2) { const int t = h1; h1 = h2; h2 = t; }
	const int ix = h2 + h1 * (2 * n_hla - h1 - 1) / 2;       // src/LibHLA.cpp:1712
	out_ratio[pos] = __ddiv_rn(P[(size_t)ix * p_stride + pos], s);
}

void launch_reduce_ib(const double *P, size_t p_stride, int n_hla, const int *samp_list,
	int n_pos, const int *a1, const int *a2, double *out_ratio, cudaStream_t st, int n_lists,
	size_t list_stride, size_t out_stride)
{
	if (n_pos <= 0 || n_lists <= 0) return;
	dim3 grid((n_pos + 63) / 64, n_lists);
	reduce_ib_kernel<<<grid, 64, 0, st>>>(P, p_stride, list_stride, n_hla, samp_list, n_pos,
		a1, a2, out_ratio, out_stride);
	CUDA_CHECK(cudaGetLastError());
}

__global__ void reduce_best_guess_kernel(const double *__restrict__ P, size_t p_stride,
	int n_hla, int n_pos, int *out_a1, int *out_a2)
{
	const int pos = blockIdx.x * blockDim.x + threadIdx.x;
	if (pos >= n_pos) return;
	double best;
	const int bi = argmax_cells(P + pos, p_stride, n_hla * (n_hla + 1) / 2, best);
	int p1 = NA_INT, p2 = NA_INT;
	if (bi >= 0) cell_to_pair(bi, n_hla, p1, p2);
	out_a1[pos] = p1; out_a2[pos] = p2;
}

void launch_reduce_best_guess(const double *P, size_t p_stride, int n_hla, int n_pos,
	int *out_a1, int *out_a2, cudaStream_t st)
{
	if (n_pos <= 0) return;
	reduce_best_guess_kernel<<<(n_pos + 63) / 64, 64, 0, st>>>(P, p_stride, n_hla, n_pos,
		out_a1, out_a2);
	CUDA_CHECK(cudaGetLastError());
}

__global__ void normalize_kernel(double *P, size_t p_stride, int n_hla, int n_pos,
	double *out_sum)
{
	const int pos = blockIdx.x * blockDim.x + threadIdx.x;
	if (pos >= n_pos) return;
	const int n_cells = n_hla * (n_hla + 1) / 2;
	const double s = seqsum_cells(P + pos, p_stride, n_cells);
	const double ff = __ddiv_rn(1.0, s);                     // src/LibHLA.cpp:1827-1828
	for (int c = 0; c < n_cells; c++)
		P[(size_t)c * p_stride + pos] = __dmul_rn(P[(size_t)c * p_stride + pos], ff);
	out_sum[pos] = s;
}

void launch_normalize(double *P, size_t p_stride, int n_hla, int n_pos, double *out_sum,
	cudaStream_t st)
{
	if (n_pos <= 0) return;
	normalize_kernel<<<(n_pos + 63) / 64, 64, 0, st>>>(P, p_stride, n_hla, n_pos, out_sum);
	CUDA_CHECK(cudaGetLastError());
}

/// out[pos][cell] = P[cell][pos], 32x32 tiles through shared memory
__global__ void transpose_kernel(const double *__restrict__ P, size_t p_stride, int n_cells,
	int n_pos, double *__restrict__ out, size_t out_stride, size_t out_row0)
{
	__shared__ double tile[32][33];
	const int c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
	for (int k = threadIdx.y; k < 32; k += blockDim.y)
	{
		const int c = c0 + k, pp = p0 + threadIdx.x;
		if (c < n_cells && pp < n_pos) tile[k][threadIdx.x] = P[(size_t)c * p_stride + pp];
	}
	__syncthreads();
	for (int k = threadIdx.y; k < 32; k += blockDim.y)
	{
		const int pp = p0 + k, c = c0 + threadIdx.x;
		if (c < n_cells && pp < n_pos)
			out[(out_row0 + pp) * out_stride + c] = tile[threadIdx.x][k];
	}
}

void launch_transpose(const double *P, size_t p_stride, int n_cells, int n_pos, double *out,
	cudaStream_t st)
{
	if (n_pos <= 0 || n_cells <= 0) return;
	dim3 grid((n_pos + 31) / 32, (n_cells + 31) / 32), block(32, 8);
	transpose_kernel<<<grid, block, 0, st>>>(P, p_stride, n_cells, n_pos, out, (size_t)n_cells, 0);
	CUDA_CHECK(cudaGetLastError());
}

// ---------------------------------------------------------------------------------------
// prediction
// ---------------------------------------------------------------------------------------

/// geno_t: SNP-major int8 [n_snp_total][n_samp_total] (transposed once per predict call)
__global__ void pack_classifier_kernel(const int8_t *__restrict__ geno_t, size_t n_samp_total,
	int samp_begin, int n_tile, const int *__restrict__ snpidx, int n_snp,
	const int *__restrict__ snp_weight, uint32_t *s1, uint32_t *s2, int stride, int nw,
	double *weight)
{
	const int pos = blockIdx.x * blockDim.x + threadIdx.x;
	if (pos >= n_tile) return;
	const size_t samp = (size_t)samp_begin + pos;
	int w_all = 0, w_ok = 0;
	for (int w = 0; w < nw; w++)
	{
		uint32_t b1 = 0u, b2 = 0xffffffffu;       // everything missing (src/LibHLA.cpp:671-673)
		const int lo = w * 32, hi = min(n_snp, lo + 32);
		for (int i = lo; i < hi; i++)
		{
			const int k = __ldg(snpidx + i);
			const int g = geno_t[(size_t)k * n_samp_total + samp];
			const int sw = __ldg(snp_weight + k);
			const uint32_t bit = 1u << (i - lo);
			w_all += sw;
			if (g == 0) { b2 &= ~bit; w_ok += sw; }
			else if (g == 1) { b1 |= bit; b2 &= ~bit; w_ok += sw; }
			else if (g == 2) { b1 |= bit; w_ok += sw; }
		}
		s1[(size_t)w * stride + pos] = b1;
		s2[(size_t)w * stride + pos] = b2;
	}
	// classifier weight = share of its SNP weights that are non-missing (src/LibHLA.cpp:2418-2431)
	weight[pos] = (w_all > 0) ? __ddiv_rn((double)w_ok, (double)w_all) : 0.0;
}

void launch_pack_classifier(const int8_t *geno_t, size_t n_samp_total, int samp_begin, int n_tile,
	const int *snpidx, int n_snp, const int *snp_weight, uint32_t *s1, uint32_t *s2,
	int stride, double *weight, cudaStream_t st)
{
	if (n_tile <= 0) return;
	pack_classifier_kernel<<<(n_tile + 127) / 128, 128, 0, st>>>(geno_t, n_samp_total,
		samp_begin, n_tile, snpidx, n_snp, snp_weight, s1, s2, stride, geno_words(n_snp), weight);
	CUDA_CHECK(cudaGetLastError());
}

__global__ void transpose_i8_kernel(const int8_t *__restrict__ in, int rows, int cols,
	int8_t *__restrict__ out)
{
	__shared__ int8_t tile[32][33];
	const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
	for (int k = threadIdx.y; k < 32; k += blockDim.y)
	{
		const int r = r0 + k, c = c0 + threadIdx.x;
		if (r < rows && c < cols) tile[k][threadIdx.x] = in[(size_t)r * cols + c];
	}
	__syncthreads();
	for (int k = threadIdx.y; k < 32; k += blockDim.y)
	{
		const int c = c0 + k, r = r0 + threadIdx.x;
		if (r < rows && c < cols) out[(size_t)c * rows + r] = tile[threadIdx.x][k];
	}
}

void launch_transpose_i8(const int8_t *in, int rows, int cols, int8_t *out, cudaStream_t st)
{
	if (rows <= 0 || cols <= 0) return;
	dim3 grid((cols + 31) / 32, (rows + 31) / 32), block(32, 8);
	transpose_i8_kernel<<<grid, block, 0, st>>>(in, rows, cols, out);
	CUDA_CHECK(cudaGetLastError());
}

/// CPU branch of _PredictHLA for one classifier (src/LibHLA.cpp:2451-2464 with PostProb2's
/// normalisation :1823-1829 and AddProbToSum :1497-1507)
__global__ void predict_accumulate_kernel(const double *__restrict__ P, size_t p_stride,
	int n_cells, int n_tile, const double *__restrict__ weight, double *acc, size_t acc_stride,
	double *aux)
{
	const int pos = blockIdx.x * blockDim.x + threadIdx.x;
	if (pos >= n_tile) return;
	const double w = weight[pos];
	if (w <= 0) return;
	const double s = seqsum_cells(P + pos, p_stride, n_cells);
	const double ff = __ddiv_rn(1.0, s);
	for (int c = 0; c < n_cells; c++)
	{
		const double pr = __dmul_rn(P[(size_t)c * p_stride + pos], ff);
		double *a = acc + (size_t)c * acc_stride + pos;
		*a = __dadd_rn(*a, __dmul_rn(pr, w));
	}
	aux[pos] = __dadd_rn(aux[pos], w);                                        // sum of weights
	aux[(size_t)n_tile + pos] = __dadd_rn(aux[(size_t)n_tile + pos], __dmul_rn(s, w));
	aux[2 * (size_t)n_tile + pos] += 1.0;
}

void launch_predict_accumulate(const double *P, size_t p_stride, int n_cells, int n_tile,
	const double *weight, double *acc, size_t acc_stride, double *aux, cudaStream_t st)
{
	if (n_tile <= 0) return;
	predict_accumulate_kernel<<<(n_tile + 63) / 64, 64, 0, st>>>(P, p_stride, n_cells, n_tile,
		weight, acc, acc_stride, aux);
	CUDA_CHECK(cudaGetLastError());
}

// ---- exact de-duplication of a tile's packed genotypes (see kernels.h) ---------------------------
__device__ __forceinline__ uint64_t mix64(uint64_t x)
{
	x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
	return x;
}

template <int NW>
__global__ void __launch_bounds__(256) dedup_insert_kernel(const uint32_t *__restrict__ s1,
	const uint32_t *__restrict__ s2, int stride, int n_tile, int *table, unsigned mask, int *repof,
	int *uid, int *rep_list, int *n_unique)
{
	const int pos = blockIdx.x * blockDim.x + threadIdx.x;
	if (pos >= n_tile) return;
	uint32_t a[NW], b[NW];
	uint64_t h = 0x9e3779b97f4a7c15ull;
#pragma unroll
	for (int w = 0; w < NW; w++)
	{
		a[w] = s1[(size_t)w * stride + pos];
		b[w] = s2[(size_t)w * stride + pos];
		h = mix64(h ^ (((uint64_t)a[w] << 32) | b[w]));
	}
	unsigned slot = (unsigned)h & mask;
	for (;;)
	{
		int r = *(volatile int *)(table + slot);
		if (r < 0)
		{
			r = atomicCAS(table + slot, -1, pos);
			if (r < 0)
			{
				// this sample represents its genotype: the next dense number is its own
				const int u = atomicAdd(n_unique, 1);
				repof[pos] = pos;
				uid[pos] = u;
				rep_list[u] = pos;
				return;
			}
		}
		bool same = true;
#pragma unroll
		for (int w = 0; w < NW; w++)
			same = same && (s1[(size_t)w * stride + r] == a[w]) && (s2[(size_t)w * stride + r] == b[w]);
		if (same) { repof[pos] = r; return; }
		slot = (slot + 1) & mask;
	}
}

__global__ void dedup_resolve_kernel(const int *__restrict__ repof, int n_tile, int *uid)
{
	const int pos = blockIdx.x * blockDim.x + threadIdx.x;
	if (pos >= n_tile) return;
	const int r = repof[pos];
	if (r != pos) uid[pos] = uid[r];        // uid[r] was written by the insert launch
}

void launch_dedup_genotypes(const uint32_t *s1, const uint32_t *s2, int stride, int nw, int n_tile,
	int *table, int table_size, int *repof, int *uid, int *rep_list, int *n_unique, cudaStream_t st)
{
	if (n_tile <= 0) return;
	if (table_size < 2 * n_tile || (table_size & (table_size - 1)))
		throw std::runtime_error("launch_dedup_genotypes: table size must be a power of two >= 2 * n_tile");
	CUDA_CHECK(cudaMemsetAsync(table, 0xff, sizeof(int) * (size_t)table_size, st));
	CUDA_CHECK(cudaMemsetAsync(n_unique, 0, sizeof(int), st));
	const int grid = (n_tile + 255) / 256;
	const unsigned mask = (unsigned)table_size - 1u;
	if (nw == 1) dedup_insert_kernel<1><<<grid, 256, 0, st>>>(s1, s2, stride, n_tile, table, mask, repof, uid, rep_list, n_unique);
	else if (nw == 2) dedup_insert_kernel<2><<<grid, 256, 0, st>>>(s1, s2, stride, n_tile, table, mask, repof, uid, rep_list, n_unique);
	else dedup_insert_kernel<4><<<grid, 256, 0, st>>>(s1, s2, stride, n_tile, table, mask, repof, uid, rep_list, n_unique);
	CUDA_CHECK(cudaGetLastError());
	dedup_resolve_kernel<<<grid, 256, 0, st>>>(repof, n_tile, uid);
	CUDA_CHECK(cudaGetLastError());
}

/// per distinct genotype: PostProb2's sum and its reciprocal (src/LibHLA.cpp:1823-1829)
__global__ void predict_unique_norm_kernel(const double *__restrict__ P, size_t p_stride, int n_cells,
	const int *__restrict__ n_unique, int cap, double *norm, size_t u_stride)
{
	const int u = blockIdx.x * blockDim.x + threadIdx.x;
	if (u >= min(__ldg(n_unique), cap)) return;
	const double s = seqsum_cells(P + u, p_stride, n_cells);
	norm[u] = __ddiv_rn(1.0, s);
	norm[u_stride + u] = s;
}

/// predict_accumulate_kernel reading the cell matrix through uid[]
__global__ void predict_accumulate_dedup_kernel(const double *__restrict__ P, size_t p_stride,
	int n_cells, int n_tile, const int *__restrict__ uid, const double *__restrict__ norm,
	size_t u_stride, const double *__restrict__ weight, double *acc, size_t acc_stride, double *aux)
{
	const int pos = blockIdx.x * blockDim.x + threadIdx.x;
	if (pos >= n_tile) return;
	const double w = weight[pos];
	if (w <= 0) return;
	const int u = __ldg(uid + pos);
	const double ff = norm[u], s = norm[u_stride + u];
	const double *Pu = P + u;
	for (int c = 0; c < n_cells; c++)
	{
		const double pr = __dmul_rn(__ldg(Pu + (size_t)c * p_stride), ff);
		double *a = acc + (size_t)c * acc_stride + pos;
		*a = __dadd_rn(*a, __dmul_rn(pr, w));
	}
	aux[pos] = __dadd_rn(aux[pos], w);                                        // sum of weights
	aux[(size_t)n_tile + pos] = __dadd_rn(aux[(size_t)n_tile + pos], __dmul_rn(s, w));
	aux[2 * (size_t)n_tile + pos] += 1.0;
}

void launch_predict_accumulate_dedup(const double *P, size_t p_stride, int n_cells, int n_tile,
	const int *uid, const int *n_unique, double *norm, size_t u_stride, const double *weight,
	double *acc, size_t acc_stride, double *aux, cudaStream_t st)
{
	if (n_tile <= 0) return;
	predict_unique_norm_kernel<<<(n_tile + 63) / 64, 64, 0, st>>>(P, p_stride, n_cells, n_unique,
		n_tile, norm, u_stride);
	CUDA_CHECK(cudaGetLastError());
	predict_accumulate_dedup_kernel<<<(n_tile + 63) / 64, 64, 0, st>>>(P, p_stride, n_cells, n_tile,
		uid, norm, u_stride, weight, acc, acc_stride, aux);
	CUDA_CHECK(cudaGetLastError());
}

/// per sample: normalise (src/LibHLA.cpp:1509-1518), matching (:2480), ensemble best guess
/// (:2370-2382), dosage (:2387-2402)
__global__ void predict_finalize_kernel(double *acc, size_t acc_stride, const double *aux,
	int n_aux_stride, int n_hla, int samp_begin, int n_tile, int *h1, int *h2,
	double *max_prob, double *matching, double *dosage)
{
	const int pos = blockIdx.x * blockDim.x + threadIdx.x;
	if (pos >= n_tile) return;
	const int n_cells = n_hla * (n_hla + 1) / 2;
	const double sw = aux[pos];
	double *a = acc + pos;
	if (sw > 0)
	{
		const double ff = __ddiv_rn(1.0, sw);
		for (int c = 0; c < n_cells; c++)
			a[(size_t)c * acc_stride] = __dmul_rn(a[(size_t)c * acc_stride], ff);
	}
	const size_t g = (size_t)samp_begin + pos;
	if (matching) matching[g] = __ddiv_rn(aux[(size_t)n_aux_stride + pos], sw);
	double best;
	const int bi = argmax_cells(a, acc_stride, n_cells, best);
	int p1 = NA_INT, p2 = NA_INT;
	if (bi >= 0) cell_to_pair(bi, n_hla, p1, p2);
	if (h1) h1[g] = p1;
	if (h2) h2[g] = p2;
	if (max_prob) max_prob[g] = (bi >= 0) ? a[(size_t)bi * acc_stride] : 0.0;
	if (dosage)
	{
		for (int h = 0; h < n_hla; h++)
		{
			double d = 0.0;
			// rows above: cell (k, h), k < h, in row order
			int idx = h;                       // index of (0, h)
			for (int k = 0; k < h; k++)
			{
				d = __dadd_rn(d, a[(size_t)idx * acc_stride]);
				idx += n_hla - k - 1;          // (k+1, h) = idx(k,h) + (n_hla-k) - 1
			}
			// own row: 2*(h,h) then (h, h2 > h); idx now points at (h, h)
			d = __dadd_rn(d, __dmul_rn(2.0, a[(size_t)idx * acc_stride]));
			for (int k = h + 1; k < n_hla; k++)
			{
				idx++;
				d = __dadd_rn(d, a[(size_t)idx * acc_stride]);
			}
			dosage[g * n_hla + h] = d;
		}
	}
}

void launch_predict_finalize(double *acc, size_t acc_stride, const double *aux, int n_hla,
	int samp_begin, int n_tile, int *h1, int *h2, double *max_prob, double *matching,
	double *dosage, double *post_prob, cudaStream_t st)
{
	if (n_tile <= 0) return;
	predict_finalize_kernel<<<(n_tile + 63) / 64, 64, 0, st>>>(acc, acc_stride, aux, n_tile,
		n_hla, samp_begin, n_tile, h1, h2, max_prob, matching, dosage);
	CUDA_CHECK(cudaGetLastError());
	if (post_prob)
	{
		const int n_cells = n_hla * (n_hla + 1) / 2;
		dim3 grid((n_tile + 31) / 32, (n_cells + 31) / 32), block(32, 8);
		transpose_kernel<<<grid, block, 0, st>>>(acc, acc_stride, n_cells, n_tile, post_prob,
			(size_t)n_cells, (size_t)samp_begin);
		CUDA_CHECK(cudaGetLastError());
	}
}

/// [sample][n_cells + 3] export of the un-normalised accumulators (classifier-sharded path)
__global__ void export_partial_kernel(const double *__restrict__ acc, size_t acc_stride,
	const double *__restrict__ aux, int n_cells, int samp_begin, int n_tile, double *out)
{
	const int pos = blockIdx.x * blockDim.x + threadIdx.x;
	if (pos >= n_tile) return;
	double *o = out + ((size_t)samp_begin + pos) * (n_cells + 3);
	for (int c = 0; c < n_cells; c++) o[c] = acc[(size_t)c * acc_stride + pos];
	o[n_cells] = aux[pos];
	o[n_cells + 1] = aux[(size_t)n_tile + pos];
	o[n_cells + 2] = aux[2 * (size_t)n_tile + pos];
}

void launch_export_partial(const double *acc, size_t acc_stride, const double *aux,
	int n_cells, int samp_begin, int n_tile, double *out, cudaStream_t st)
{
	if (n_tile <= 0) return;
	export_partial_kernel<<<(n_tile + 63) / 64, 64, 0, st>>>(acc, acc_stride, aux, n_cells,
		samp_begin, n_tile, out);
	CUDA_CHECK(cudaGetLastError());
}

__global__ void finalize_from_partial_kernel(const double *__restrict__ partial, int n_hla,
	int n_samp, int *h1, int *h2, double *max_prob, double *matching, double *dosage,
	double *post_prob)
{
	const int s = blockIdx.x * blockDim.x + threadIdx.x;
	if (s >= n_samp) return;
	const int n_cells = n_hla * (n_hla + 1) / 2;
	const double *a = partial + (size_t)s * (n_cells + 3);
	const double sw = a[n_cells];
	const double ff = (sw > 0) ? __ddiv_rn(1.0, sw) : 1.0;
	if (matching) matching[s] = __ddiv_rn(a[n_cells + 1], sw);
	double best = 0.0;
	int bi = -1;
	for (int c = 0; c < n_cells; c++)
	{
		const double v = __dmul_rn(a[c], ff);
		if (post_prob) post_prob[(size_t)s * n_cells + c] = v;
		if (best < v) { best = v; bi = c; }
	}
	int p1 = NA_INT, p2 = NA_INT;
	if (bi >= 0) cell_to_pair(bi, n_hla, p1, p2);
	if (h1) h1[s] = p1;
	if (h2) h2[s] = p2;
	if (max_prob) max_prob[s] = (bi >= 0) ? best : 0.0;
	if (dosage)
	{
		for (int h = 0; h < n_hla; h++)
		{
			double d = 0.0;
			int idx = h;
			for (int k = 0; k < h; k++)
			{
				d = __dadd_rn(d, __dmul_rn(a[idx], ff));
				idx += n_hla - k - 1;
			}
			d = __dadd_rn(d, __dmul_rn(2.0, __dmul_rn(a[idx], ff)));
			for (int k = h + 1; k < n_hla; k++)
			{
				idx++;
				d = __dadd_rn(d, __dmul_rn(a[idx], ff));
			}
			dosage[(size_t)s * n_hla + h] = d;
		}
	}
}

void launch_finalize_from_partial(const double *partial, int n_hla, int n_samp, int *h1,
	int *h2, double *max_prob, double *matching, double *dosage, double *post_prob,
	cudaStream_t st)
{
	if (n_samp <= 0) return;
	finalize_from_partial_kernel<<<(n_samp + 63) / 64, 64, 0, st>>>(partial, n_hla, n_samp,
		h1, h2, max_prob, matching, dosage, post_prob);
	CUDA_CHECK(cudaGetLastError());
}

// ---------------------------------------------------------------------------------------
// pipe-rate microbenchmarks: the denominators of this path's roofline
// ---------------------------------------------------------------------------------------
template <int WHICH>
__global__ void __launch_bounds__(256) pipe_peak_kernel(int iters, uint32_t seed, uint32_t *sink)
{
	__shared__ double tbl[66 * 32];
	const int tid = threadIdx.x;
	for (int k = tid; k < 66 * 32; k += blockDim.x) tbl[k] = 1.0 + k * 1e-9;
	__syncthreads();
	uint32_t x0 = seed ^ (tid * 2654435761u), x1 = x0 * 3 + 1, x2 = x0 * 5 + 2, x3 = x0 * 7 + 3;
	uint32_t y0 = x0 ^ 0x9e3779b9u, y1 = x1 ^ 0x7f4a7c15u, y2 = x2 ^ 0x94d049bbu, y3 = x3 ^ 0xbf58476du;
	double d0 = 1.0 + tid * 1e-6, d1 = 1.000001, d2 = 0.999999, d3 = 1.0000003;
	const double m0 = 1.0000001, m1 = 0.9999999;
	const uint32_t lane_addr = smem_u32(tbl) + (tid & 31) * 8;
	for (int it = 0; it < iters; it++)
	{
#pragma unroll
		for (int u = 0; u < 16; u++)
		{
			if (WHICH == 0)          // POPC.32, 8 independent chains, 1 POPC + 1 LOP-class op each
			{
				x0 = __popc(x0 ^ y0) + y0; x1 = __popc(x1 ^ y1) + y1;
				x2 = __popc(x2 ^ y2) + y2; x3 = __popc(x3 ^ y3) + y3;
				y0 = __popc(y0 ^ x1) + x1; y1 = __popc(y1 ^ x2) + x2;
				y2 = __popc(y2 ^ x3) + x3; y3 = __popc(y3 ^ x0) + x0;
			} else if (WHICH == 1)   // LOP3
			{
				x0 = (x0 ^ y0) & y1; x1 = (x1 ^ y1) | y2;
				x2 = (x2 ^ y2) & y3; x3 = (x3 ^ y3) | y0;
				y0 = (y0 | x2) ^ x3; y1 = (y1 & x3) ^ x0;
				y2 = (y2 | x0) ^ x1; y3 = (y3 & x1) ^ x2;
			} else if (WHICH == 2)   // DMUL + DADD un-fused
			{
				d0 = __dadd_rn(__dmul_rn(d0, m0), m1); d1 = __dadd_rn(__dmul_rn(d1, m1), m0);
				d2 = __dadd_rn(__dmul_rn(d2, m0), m1); d3 = __dadd_rn(__dmul_rn(d3, m1), m0);
			} else if (WHICH == 3)   // DFMA
			{
				d0 = __fma_rn(d0, m0, m1); d1 = __fma_rn(d1, m1, m0);
				d2 = __fma_rn(d2, m0, m1); d3 = __fma_rn(d3, m1, m0);
				d0 = __fma_rn(d0, m1, m0); d1 = __fma_rn(d1, m0, m1);
				d2 = __fma_rn(d2, m1, m0); d3 = __fma_rn(d3, m0, m1);
			} else if (WHICH == 4)   // LDS.64 lane-private column, data-dependent row
			{
				d0 += lds_f64(lane_addr + ((x0 + u) & 63u) * 256u);
				d1 += lds_f64(lane_addr + ((x1 + u) & 63u) * 256u);
				d2 += lds_f64(lane_addr + ((x2 + u) & 63u) * 256u);
				d3 += lds_f64(lane_addr + ((x3 + u) & 63u) * 256u);
			} else if (WHICH == 8 || WHICH == 9 || WHICH == 10)
			{
				// LDS.64 from a COMPACT table (8-byte stride, no lane replication), data-dependent
				// row: 8 -- rows spread over a window of 8 (distinct banks); 9 -- all lanes one row;
				// 10 -- rows spread over 32 (row d and d + 16 share their banks)
				const uint32_t msk = (WHICH == 8) ? 7u : (WHICH == 9) ? 0u : 31u;
				const uint32_t base = smem_u32(tbl);
				d0 += lds_f64(base + (((x0 & msk) + u) & 63u) * 8u);
				d1 += lds_f64(base + (((x1 & msk) + u) & 63u) * 8u);
				d2 += lds_f64(base + (((x2 & msk) + u) & 63u) * 8u);
				d3 += lds_f64(base + (((x3 & msk) + u) & 63u) * 8u);
			} else if (WHICH == 6)   // one dependent DADD chain per thread (latency)
			{
				d0 = __dadd_rn(d0, m0);
			} else if (WHICH == 7)   // one dependent DMUL+DADD chain per thread (latency)
			{
				d0 = __dadd_rn(__dmul_rn(d0, m1), m0);
			} else {                 // IADD3
				x0 = x0 + y0 + y1; x1 = x1 + y1 + y2; x2 = x2 + y2 + y3; x3 = x3 + y3 + y0;
				y0 = y0 + x1 + x2; y1 = y1 + x2 + x3; y2 = y2 + x3 + x0; y3 = y3 + x0 + x1;
			}
		}
	}
	const uint32_t r = x0 ^ x1 ^ x2 ^ x3 ^ y0 ^ y1 ^ y2 ^ y3 ^
		(uint32_t)__double2loint(d0 + d1 + d2 + d3);
	if (r == 0x12345678u) sink[0] = r;
}

double run_pipe_peak(int which, int sm_count, double *out_ms)
{
	uint32_t *sink = nullptr;
	CUDA_CHECK(cudaMalloc(&sink, 4));
	const int iters = 4096;
	// 6, 7: dependent-chain latency -- one warp on one SM
	const bool lat = (which == 6 || which == 7);
	const int grid = lat ? 1 : sm_count * 8, block = lat ? 32 : 256;
	// lane-operations of the measured kind per thread per inner iteration (16x unrolled)
	static const int ops_per_u[11] = { 8, 8, 8, 8, 4, 8, 1, 1, 4, 4, 4 };
	cudaEvent_t e0, e1;
	CUDA_CHECK(cudaEventCreate(&e0)); CUDA_CHECK(cudaEventCreate(&e1));
	float best = 1e30f;
	for (int rep = 0; rep < 4; rep++)
	{
		CUDA_CHECK(cudaEventRecord(e0));
		switch (which)
		{
		case 0: pipe_peak_kernel<0><<<grid, block>>>(iters, 1u + rep, sink); break;
		case 1: pipe_peak_kernel<1><<<grid, block>>>(iters, 1u + rep, sink); break;
		case 2: pipe_peak_kernel<2><<<grid, block>>>(iters, 1u + rep, sink); break;
		case 3: pipe_peak_kernel<3><<<grid, block>>>(iters, 1u + rep, sink); break;
		case 4: pipe_peak_kernel<4><<<grid, block>>>(iters, 1u + rep, sink); break;
		case 8: pipe_peak_kernel<8><<<grid, block>>>(iters, 1u + rep, sink); break;
		case 9: pipe_peak_kernel<9><<<grid, block>>>(iters, 1u + rep, sink); break;
		case 10: pipe_peak_kernel<10><<<grid, block>>>(iters, 1u + rep, sink); break;
		default: pipe_peak_kernel<5><<<grid, block>>>(iters, 1u + rep, sink); break;
		}
		CUDA_CHECK(cudaEventRecord(e1));
		CUDA_CHECK(cudaEventSynchronize(e1));
		float ms = 0;
		CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
		if (rep > 0 && ms < best) best = ms;
	}
	CUDA_CHECK(cudaEventDestroy(e0)); CUDA_CHECK(cudaEventDestroy(e1));
	CUDA_CHECK(cudaFree(sink));
	const int w = (which < 0 || which > 10) ? 5 : which;
	const double ops = (double)grid * block * (double)iters * 16.0 * ops_per_u[w];
	if (out_ms) *out_ms = best;
	return ops / (best * 1e-3);
}

}  // namespace hb
