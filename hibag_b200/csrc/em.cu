// em.cu -- see em.h
#include "em.h"
#include "kernels.h"
#include "devutil.cuh"

#include <cooperative_groups.h>
#include <cub/cub.cuh>

#include <algorithm>
#include <cfloat>
#include <cstdio>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <condition_variable>
#include <mutex>

#include "kernels.h"

namespace cg = cooperative_groups;

namespace hb {

namespace {

constexpr int EM_THREADS_MAX = 1024;
constexpr int EM_MAX_ITER = 500;                 // src/LibHLA.cpp:98
constexpr double EM_INIT_VAL_FRAC = 0.001;       // :100
// Half-width of the band around the stopping tolerance inside which the device does not decide,
// relative to |LL|. LL = sum of n same-sign terms bc*log(psum). With u = 2^-53 = 1.11e-16: every
// term carries <= 1 ulp of log (2u; glibc and the device log alike) plus u of the product, the
// host's sequential sum at most (n-1)*u relative (the standard worst case), the device's tree sum
// at most (log2 n + 1)*u; the test compares |LL - LL_old|, two such sums on either side, so host
// and device differ by at most 2*((n + 2) + (log2 n + 4))*u*|LL| <= 2*(n + 24)*u*|LL| for
// n < 2^17. The band is 1.5 times that bound (the former 2*(1.2e-16*n + 1e-15) cleared
// the worst case by 9 % only). At n = 3,160 it is 1.1e-12, i.e. 7e-5 of the tolerance sqrt(eps).
__host__ __device__ inline double em_guard_rel(int n_entry) { return 3.0 * ((n_entry + 24) * 1.12e-16); }

// ---------------------------------------------------------------------------------------------
// haplotype-pair matching
// ---------------------------------------------------------------------------------------------

/// reference hamm_d, generic form (src/LibHLA.cpp:802-817)
__device__ __forceinline__ int hamm_ref(const uint64_t S1[2], const uint64_t S2[2],
	const uint64_t *__restrict__ A, const uint64_t *__restrict__ B, int words)
{
	int d = 0;
	for (int w = 0; w < words; w++)
	{
		const uint64_t a = A[w], b = B[w];
		const uint64_t miss = S2[w] & ~S1[w];
		const uint64_t mask = ((a ^ S2[w]) | (b ^ S1[w])) & ~miss;
		d += __popcll((a ^ S1[w]) & mask) + __popcll((b ^ S2[w]) & mask);
	}
	return d;
}

struct MatchArgs
{
	const uint64_t *hap;          // [n_cur][2]
	const int *start;             // [n_hla + 1]
	const uint32_t *s1, *s2;      // SoA planes [4][stride]
	int stride;
	const int *a1, *a2;           // true types per sample, a1 <= a2
	const int *ib;                // in-bag sample per entry
	int n_entry, n_snp;
	int *cnt;                     // per entry: pairs (doubled) or records
	int *mind;                    // per entry: minimum distance
	const int *off;               // exclusive scan of cnt
	int *p1, *p2;                 // doubled pairs
	uint32_t *rec;                // records (entry, (i2 << 16) | i1)
	int *empty_flag;              // set when an entry has no pair at all (an allele without haplotypes)
};

/// MODE 0: minimum distance + number of doubled pairs; 1: emit the doubled pairs in the
/// reference's scan order (first index outer, second inner over the doubled ranges, :1578-1634);
/// 2: number of records; 3: emit records (i1 outer, i2 inner)
template <int MODE>
__global__ void __launch_bounds__(128) haplomatch_kernel(const MatchArgs p)
{
	const int k = blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= p.n_entry) return;
	const int s = p.ib[k];
	uint64_t S1[2], S2[2];
	{
		const size_t st = (size_t)p.stride;
		S1[0] = (uint64_t)p.s1[s] | ((uint64_t)p.s1[st + s] << 32);
		S1[1] = (uint64_t)p.s1[2 * st + s] | ((uint64_t)p.s1[3 * st + s] << 32);
		S2[0] = (uint64_t)p.s2[s] | ((uint64_t)p.s2[st + s] << 32);
		S2[1] = (uint64_t)p.s2[2 * st + s] | ((uint64_t)p.s2[3 * st + s] << 32);
	}
	const int words = (p.n_snp <= 64) ? 1 : 2;
	const int A1 = p.a1[s], A2 = p.a2[s];
	const int st1 = p.start[A1], m1 = p.start[A1 + 1] - st1;
	const int st2 = p.start[A2], m2 = p.start[A2 + 1] - st2;
	const bool same = (st1 == st2);

	if (MODE == 0 || MODE == 2)
	{
		int min_d = p.n_snp * 4, n = 0;
		for (int i = 0; i < m1; i++)
		{
			const uint64_t *hi = p.hap + 2 * (size_t)(st1 + i);
			for (int j = same ? i : 0; j < m2; j++)
			{
				const int d = hamm_ref(S1, S2, hi, p.hap + 2 * (size_t)(st2 + j), words);
				if (d < min_d) { min_d = d; n = 0; }
				if (d == min_d) n += (MODE == 2) ? 1 : ((same && i == j) ? 3 : 4);
			}
		}
		p.cnt[k] = n;
		p.mind[k] = min_d;
		if (n == 0 && p.empty_flag) *p.empty_flag = 1;
		return;
	}

	const int min_d = p.mind[k];
	int o = p.off[k];
	if (MODE == 3)
	{
		for (int i = 0; i < m1; i++)
		{
			const uint64_t *hi = p.hap + 2 * (size_t)(st1 + i);
			for (int j = same ? i : 0; j < m2; j++)
				if (hamm_ref(S1, S2, hi, p.hap + 2 * (size_t)(st2 + j), words) == min_d)
				{
					p.rec[2 * (size_t)o] = (uint32_t)k;
					p.rec[2 * (size_t)o + 1] = ((uint32_t)j << 16) | (uint32_t)i;
					o++;
				}
		}
		return;
	}
	for (int i = 0; i < m1; i++)
	{
		const uint64_t *hi = p.hap + 2 * (size_t)(st1 + i);
		for (int a = 0; a < 2; a++)
		{
			const int i2 = 2 * (st1 + i) + a;
			for (int j = same ? i : 0; j < m2; j++)
				if (hamm_ref(S1, S2, hi, p.hap + 2 * (size_t)(st2 + j), words) == min_d)
				{
					for (int b = 0; b < 2; b++)
					{
						const int j2 = 2 * (st2 + j) + b;
						if (!same || j2 >= i2)
						{
							p.p1[o] = i2; p.p2[o] = j2;
							o++;
						}
					}
				}
		}
	}
}

/// keys[2t + side] = haplotype of that side of pair t, vals = 2t + side
__global__ void incidence_fill_kernel(const int *__restrict__ p1, const int *__restrict__ p2,
	int n_pairs, int *key, int *val)
{
	const int t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= n_pairs) return;
	key[2 * t] = p1[t]; key[2 * t + 1] = p2[t];
	val[2 * t] = 2 * t; val[2 * t + 1] = 2 * t + 1;
}

/// after the stable sort by haplotype: inc_off[u] = first sorted slot of haplotype u
/// (u in [0, n2]) and len[u] = number of contributions of u
__global__ void incidence_offsets_kernel(const int *__restrict__ key_sorted, int n, int n2,
	int *inc_off)
{
	const int q = blockIdx.x * blockDim.x + threadIdx.x;
	if (q > n2) return;
	int lo = 0, hi = n;                // lower_bound(key_sorted, q)
	while (lo < hi)
	{
		const int mid = (lo + hi) >> 1;
		if (key_sorted[mid] < q) lo = mid + 1; else hi = mid;
	}
	inc_off[q] = lo;
}

__global__ void incidence_len_kernel(const int *__restrict__ inc_off, int n2, int *len, int *hap)
{
	const int u = blockIdx.x * blockDim.x + threadIdx.x;
	if (u >= n2) return;
	len[u] = inc_off[u + 1] - inc_off[u];
	hap[u] = u;
}

__global__ void iota_kernel(int n, int *out)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) out[i] = i;
}

/// Haplotypes sorted by decreasing contribution count are cut into groups of 32 (one per lane
/// of a warp); group g stores its contributions interleaved ("ELL"): slot (i, lane) at
/// group_base[g] + 64*(i/2) + 2*lane + (i&1), i < group_len[g] = the longest chain of the group
/// (two consecutive rows of a lane are 16 contiguous bytes: one cp.async.cg per lane), so the M
/// step reads 512 contiguous bytes per warp step and every lane still adds its own chain in order.
/// Single block; group_base[n_groups] = total slots.
__global__ void incidence_groups_kernel(const int *__restrict__ len_sorted,
	const int *__restrict__ hap_sorted, int n2, int *rank_of, int *group_len, int *group_base)
{
	const int n_groups = (n2 + 31) / 32;
	for (int r = threadIdx.x; r < n2; r += blockDim.x) rank_of[hap_sorted[r]] = r;
	// rows padded to a multiple of 8: padding slots stay 0.0 and the M step runs guard-free
	for (int g = threadIdx.x; g < n_groups; g += blockDim.x) group_len[g] = (len_sorted[32 * g] + 7) & ~7;
	__syncthreads();
	if (threadIdx.x == 0)
	{
		int base = 0;
		for (int g = 0; g < n_groups; g++) { group_base[g] = base; base += 32 * group_len[g]; }
		group_base[n_groups] = base;
	}
}

/// pairs4[t] = {u | v << 16, entry of the pair, slot of the contribution to u, slot of the
/// contribution to v}
__global__ void incidence_slots_kernel(const int *__restrict__ key_sorted,
	const int *__restrict__ val_sorted, int n, const int *__restrict__ inc_off,
	const int *__restrict__ rank_of, const int *__restrict__ group_base, int *pairs4)
{
	const int q = blockIdx.x * blockDim.x + threadIdx.x;
	if (q >= n) return;
	const int u = key_sorted[q], e = val_sorted[q];
	const int r = rank_of[u];
	const int i = q - inc_off[u];
	const int slot = group_base[r >> 5] + 64 * (i >> 1) + 2 * (r & 31) + (i & 1);
	pairs4[4 * (size_t)(e >> 1) + 2 + (e & 1)] = slot;
}

__global__ void pairs_pack_kernel(const int *__restrict__ p1, const int *__restrict__ p2,
	const int *__restrict__ off, int n_entry, int *pairs4)
{
	const int k = blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= n_entry) return;
	for (int t = off[k]; t < off[k + 1]; t++)
	{
		pairs4[4 * (size_t)t] = p1[t] | (p2[t] << 16);
		pairs4[4 * (size_t)t + 1] = k;
	}
}

// ---------------------------------------------------------------------------------------------
// EM of the candidates of a round: one CTA per candidate
// ---------------------------------------------------------------------------------------------
struct EmArgs
{
	int n_entry, n_cur, n_samp;
	const int *ib, *boot;             // in-bag sample per entry; bootstrap count per sample
	const int *off;                   // pair range per entry
	const int4 *pairs4;               // {u | v << 16, entry, slot_u, slot_v} per pair (all pairs)
	const int *hap_sorted, *group_base;   // ELL incidence layout (groups of 32 haplotype chains)
	const int *inc_off, *inc_val;     // incidence lists sorted by haplotype: (2t + side) in order
	const double *cur_freq;
	size_t n_slots;                   // ELL slots per candidate
	const int8_t *geno_t;             // raw genotypes, SNP-major [n_snp][n_samp]
	const int *cand_snp;              // [m]
	int total_pairs;
	int m_warps;                      // warps that run the M step, each with a RING_ROWS-row ring
	// per-candidate scratch
	int *pmap;                        // [m][4 * total_pairs] pair -> compact index | entry | two ELL slots of a compatible pair
	double *rinc;                     // [m][n_slots] contributions in ELL slot order
	int *cuv;                         // [m][total_pairs] u | v << 16 of the compatible pairs
	double *xbuf;                     // [m][total_pairs] GenoFreq of the compatible pairs
	int *coff;                        // [m][n_entry + 1] compact range per entry
	int *glen;                        // [m][n_groups] ELL rows per group (multiple of 8), zeroed
	double *out_freq;                 // [m][2 * n_cur]
	int *out_status;                  // [m][4] = status, iterations, 0, 0
	double scale;                     // 0.5 / n_samp
	double em_reltol;                 // sqrt(DBL_EPSILON)
	unsigned long long *prof;         // [m][8] clock64 ticks per phase (HIBAG_B200_EM_PROF), or null
	unsigned long long *acct;         // SM-time counters (kernels.h) or null
	int acct_w;                       // 1024 / EM CTAs that fit an SM
};

constexpr int MAX_CLUSTER = 8;

__device__ __forceinline__ void cp_async16_cg(uint32_t dst, const void *src)
{
	asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit()
{
	asm volatile("cp.async.commit_group;" ::: "memory");
}
template <int N>
__device__ __forceinline__ void cp_async_wait()
{
	asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory");
}
__device__ __forceinline__ double2 lds_f64x2(uint32_t addr)
{
	double2 v;
	asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
	return v;
}
__device__ __forceinline__ int4 lds_i32x4(uint32_t addr)
{
	int4 v;
	asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];"
		: "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
	return v;
}

__device__ __forceinline__ double block_sum_f64(double v, double *scratch)
{
	// fixed-shape tree: lanes, then warps -- deterministic for a given block size
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) v = __dadd_rn(v, __shfl_xor_sync(0xffffffffu, v, o));
	const int w = threadIdx.x >> 5;
	__syncthreads();
	if ((threadIdx.x & 31) == 0) scratch[w] = v;
	__syncthreads();
	if (w == 0)
	{
		double x = (threadIdx.x < (blockDim.x >> 5)) ? scratch[threadIdx.x] : 0.0;
#pragma unroll
		for (int o = 16; o > 0; o >>= 1) x = __dadd_rn(x, __shfl_xor_sync(0xffffffffu, x, o));
		if (threadIdx.x == 0) scratch[32] = x;
	}
	__syncthreads();
	return scratch[32];
}

/// A cluster of C CTAs (C SMs) estimates one candidate. Whole in-bag entries (and their pairs)
/// are split evenly over the CTAs for the E step; the groups of haplotype chains are dealt
/// round-robin for the M step. Every CTA keeps the full frequency vector (double-buffered) and
/// the per-entry scale factors in shared memory; new frequencies, scale factors and partial
/// log-likelihoods are written into all CTAs through distributed shared memory.
///
/// Only the pairs compatible with the candidate's genotype take part in EM (:1157-1180, about a
/// third of all pairs). Once per candidate the kernel (A) compacts them in pair order and (B)
/// assigns every compatible contribution its ELL slot counted among the compatible contributions
/// of its haplotype only, recording per compatible pair the slots of its two contributions.
/// Per iteration: (1) pair pass, coalesced: GenoFreq x of every compatible pair; (2) entry pass:
/// the sum of an entry's x in list order, its log-likelihood term and scale factor count / sum;
/// (3) contribution pass, coalesced over the pairs: r = x * (count / sum) stored into the two ELL
/// slots of the pair (the slots nothing is stored into hold 0.0 from the start: s + 0.0 == s);
/// (4) M step: every lane adds its own chain in the reference's order.
/// (Round 1 history, measured with the phase clocks below at config 2, 24 lanes: a pass over the
/// ELL slots that rebuilt r per slot from shared memory cost 150 k cycles per iteration -- 48 warp
/// instructions per 32 slots of which 60 % are padding -- against 72 k for the pair-order pass.)
/// EM_THREADS threads per CTA; RING_ROWS = ELL rows (256 B each) an M-step warp keeps in flight
template <int EM_THREADS, int RING_ROWS>
__global__ void __launch_bounds__(EM_THREADS, (EM_THREADS <= 512) ? 2 : 1) em_kernel(const EmArgs p)
{
	cg::cluster_group cluster = cg::this_cluster();
	const int C = (int)cluster.num_blocks();
	const int rank = (int)cluster.block_rank();
	SmAcct acct_scope(p.acct, SM_ACCT_EM, (unsigned)p.acct_w);
	SmAcct acct_cta(p.acct, SM_ACCT_EM_CTA, 1024u);          // plain CTA-resident cycles (latency roofline)

	extern __shared__ double em_smem[];
	const int n2 = 2 * p.n_cur;
	double *fr0 = em_smem;                     // [2][n2] frequencies, double-buffered
	double *scratch = fr0 + 2 * (size_t)n2;    // [40]
	double *llp = scratch + 40;                // [2][MAX_CLUSTER] partial log-likelihoods
	int *tot = (int *)(llp + 2 * MAX_CLUSTER); // [MAX_CLUSTER] compatible pairs per CTA (8 doubles reserved)
	double *sck = llp + 3 * MAX_CLUSTER;       // [n_entry] count / sum of the entry's GenoFreq (all entries)
	double *rings = sck + ((p.n_entry + 1) & ~1);   // [m_warps][RING_ROWS][32] M-step rings (16-byte aligned)
	// [n_entry] bootstrap count of the entry << 2 | candidate genotype (3 = missing)
	int *eg = (int *)(rings + (size_t)p.m_warps * RING_ROWS * 32);
	__shared__ int sh_i[4];

	const int c = blockIdx.x / C;
	const int tid = threadIdx.x;
	const int lane = tid & 31;
	const int warp = tid >> 5;
	// phase clocks of thread 0 (it leaves every phase through the phase's barrier)
	const bool prof = (p.prof != nullptr) && tid == 0 && rank == 0;
	__shared__ long long sh_pt[6];             // [0..4] ticks per phase, [5] last clock
	if (prof) { for (int q = 0; q < 5; q++) sh_pt[q] = 0; sh_pt[5] = clock64(); }
#define EM_TICK(i_) do { if (prof) { const long long n_ = clock64(); sh_pt[i_] += n_ - sh_pt[5]; sh_pt[5] = n_; } } while (0)
	const int8_t *col = p.geno_t + (size_t)p.cand_snp[c] * p.n_samp;
	// per-candidate scratch: per original pair its index among the compatible pairs; per compatible
	// pair its entry and the ELL slots of its two contributions
	int *jmap = p.pmap + (size_t)c * 4 * p.total_pairs;
	int *ce = jmap + p.total_pairs;
	int2 *cslot = (int2 *)(ce + p.total_pairs);
	double *rinc = p.rinc + (size_t)c * p.n_slots;
	int *cuv = p.cuv + (size_t)c * p.total_pairs;
	double *xbuf = p.xbuf + (size_t)c * p.total_pairs;
	int *status = p.out_status + 4 * c;

	// allele frequency of the new SNP in the bootstrap sample (:1136-1151), integers; every CTA
	// of the cluster computes the same numbers
	{
		int ac = 0, vc = 0;
		for (int k = tid; k < p.n_entry; k += EM_THREADS)
		{
			const int s = p.ib[k];
			const int g = col[s];
			const int b = p.boot[s];
			eg[k] = (b << 2) | ((0 <= g && g <= 2) ? g : 3);
			if (0 <= g && g <= 2) { ac += g * b; vc += 2 * b; }
		}
		if (tid < 4) sh_i[tid] = 0;
		__syncthreads();
#pragma unroll
		for (int o = 16; o > 0; o >>= 1)
		{
			ac += __shfl_xor_sync(0xffffffffu, ac, o);
			vc += __shfl_xor_sync(0xffffffffu, vc, o);
		}
		if (lane == 0) { atomicAdd(&sh_i[0], ac); atomicAdd(&sh_i[1], vc); }
		__syncthreads();
	}
	const int allele_cnt = sh_i[0], valid_cnt = sh_i[1];
	if (allele_cnt == 0 || allele_cnt == valid_cnt)
	{
		if (tid == 0 && rank == 0) { status[0] = EM_INVALID; status[1] = 0; }
		return;                                 // uniform over the cluster
	}

	// doubled list, initial frequencies (:444-459)
	{
		const double af = __ddiv_rn((double)allele_cnt, (double)valid_cnt);
		const double q0 = __dsub_rn(1.0, af), q1 = af;
		for (int k = tid; k < p.n_cur; k += EM_THREADS)
		{
			const double f = p.cur_freq[k];
			fr0[2 * k] = __dadd_rn(__dmul_rn(q0, f), EM_INIT_VAL_FRAC);
			fr0[2 * k + 1] = __dadd_rn(__dmul_rn(q1, f), EM_INIT_VAL_FRAC);
		}
	}
	// this CTA's slice: entries [k_lo, k_hi) and exactly their pairs [t_lo, t_hi)
	int k_lo, k_hi;
	{
		auto first_entry_at = [&](long long target) {      // first k with off[k] >= target
			int lo = 0, hi = p.n_entry;
			while (lo < hi) { const int mid = (lo + hi) >> 1; if (p.off[mid] < target) lo = mid + 1; else hi = mid; }
			return lo;
		};
		k_lo = (rank == 0) ? 0 : first_entry_at((long long)p.total_pairs * rank / C);
		k_hi = (rank == C - 1) ? p.n_entry : first_entry_at((long long)p.total_pairs * (rank + 1) / C);
	}
	const int t_lo = p.off[k_lo], t_hi = p.off[k_hi];
	const int n_groups = (n2 + 31) >> 5;
	int *coff = p.coff + (size_t)c * (p.n_entry + 1);
	int *glen = p.glen + (size_t)c * n_groups;
	// one CTA per candidate (the dense shape with many lanes): a CTA barrier and plain shared-memory
	// stores do what the cluster barrier and the DSMEM stores do for C > 1
	auto csync = [&]() { if (C > 1) cluster.sync(); else __syncthreads(); };
	csync();                                    // all CTAs are running: remote shared memory is live

	// ---- (A) compact list of the compatible pairs of this CTA's slice, in pair order ------------
	int jb_lo, jb_hi;
	{
		const int n_slice = t_hi - t_lo;
		const int chunk = (n_slice + EM_THREADS - 1) / EM_THREADS;
		const int tb = min(t_hi, t_lo + tid * chunk), te = min(t_hi, tb + chunk);
		int cnt = 0;
		for (int t = tb; t < te; t += 4)           // four pair records in flight per thread
		{
			int4 pr[4];
#pragma unroll
			for (int q = 0; q < 4; q++) pr[q] = (t + q < te) ? __ldg(p.pairs4 + t + q) : make_int4(0, 0, 0, 0);
#pragma unroll
			for (int q = 0; q < 4; q++)
			{
				const int u = pr[q].x & 0xffff, v = (int)((unsigned)pr[q].x >> 16);
				const int g = eg[pr[q].y] & 3;
				cnt += (t + q < te && (g == 3 || ((u & 1) + (v & 1)) == g)) ? 1 : 0;
			}
		}
		// block exclusive scan of cnt (warp shuffles + one pass over the 32 warp totals)
		int incl = cnt;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1)
		{
			const int y = __shfl_up_sync(0xffffffffu, incl, o);
			if (lane >= o) incl += y;
		}
		int *wsum = (int *)scratch;             // [32] ints (scratch is 40 doubles)
		__syncthreads();
		if (lane == 31) wsum[warp] = incl;
		__syncthreads();
		if (tid < 32)
		{
			int w = (tid < (EM_THREADS >> 5)) ? wsum[tid] : 0;
#pragma unroll
			for (int o = 1; o < 32; o <<= 1)
			{
				const int y = __shfl_up_sync(0xffffffffu, w, o);
				if (tid >= o) w += y;
			}
			wsum[tid] = w;                       // inclusive over warps
		}
		__syncthreads();
		const int cta_total = wsum[31];
		const int start = incl - cnt + (warp ? wsum[warp - 1] : 0);
		if (tid < C) cluster.map_shared_rank(tot, tid)[rank] = cta_total;
		csync();
		int base = 0;
		for (int q = 0; q < rank; q++) base += tot[q];
		jb_lo = base; jb_hi = base + cta_total;
		int j = base + start;
		for (int t = tb; t < te; t += 4)
		{
			int4 pr[4];
#pragma unroll
			for (int q = 0; q < 4; q++) pr[q] = (t + q < te) ? __ldg(p.pairs4 + t + q) : make_int4(0, 0, 0, 0);
#pragma unroll
			for (int q = 0; q < 4; q++)
			{
				if (t + q < te)
				{
					const int u = pr[q].x & 0xffff, v = (int)((unsigned)pr[q].x >> 16);
					const int g = eg[pr[q].y] & 3;
					if (t + q == p.off[pr[q].y]) coff[pr[q].y] = j;      // first pair of its entry
					if (g == 3 || ((u & 1) + (v & 1)) == g)
					{
						jmap[t + q] = j; ce[j] = pr[q].y;
						cuv[j++] = pr[q].x;
					}
				}
			}
		}
		if (rank == C - 1 && tid == 0) coff[p.n_entry] = jb_hi;
		if (tid == 0) sh_i[3] = 0;
		csync();                                   // (B) reads the jmap entries of every CTA's pairs
	}
	// ---- (B) ELL slots of the compatible contributions: a warp takes one haplotype of the groups
	// dealt to this CTA at a time and walks its incidence list 128 contributions per step ---------
	{
		const int n_local_groups = (n_groups > rank) ? (n_groups - rank + C - 1) / C : 0;
		for (;;)
		{
			int idx = 0;
			if (lane == 0) idx = atomicAdd(&sh_i[3], 1);
			idx = __shfl_sync(0xffffffffu, idx, 0);
			if (idx >= n_local_groups * 32) break;
			const int gi = (idx >> 5) * C + rank, l = idx & 31;
			const int r = 32 * gi + l;
			if (r >= n2) continue;
			const int u = p.hap_sorted[r];
			const int gbase = p.group_base[gi];
			const int qb = p.inc_off[u], qe = p.inc_off[u + 1];
			int n = 0;
			for (int q0 = qb; q0 < qe; q0 += 128)
			{
				int e[4]; int4 pr[4];
#pragma unroll
				for (int w = 0; w < 4; w++)
				{
					const int q = q0 + w * 32 + lane;
					e[w] = (q < qe) ? __ldg(p.inc_val + q) : -1;
				}
#pragma unroll
				for (int w = 0; w < 4; w++)
					pr[w] = (e[w] >= 0) ? __ldg(p.pairs4 + (e[w] >> 1)) : make_int4(0, 0, 0, 0);
#pragma unroll
				for (int w = 0; w < 4; w++)
				{
					bool ok = false;
					if (e[w] >= 0)
					{
						const int uu = pr[w].x & 0xffff, vv = (int)((unsigned)pr[w].x >> 16);
						const int g = eg[pr[w].y] & 3;
						ok = (g == 3 || ((uu & 1) + (vv & 1)) == g);
					}
					const unsigned m = __ballot_sync(0xffffffffu, ok);
					if (ok)
					{
						const int rk = n + __popc(m & ((1u << lane) - 1u));
						const int slot = gbase + 64 * (rk >> 1) + 2 * l + (rk & 1);
						((int *)cslot)[2 * jmap[e[w] >> 1] + (e[w] & 1)] = slot;
					}
					n += __popc(m);
				}
			}
			if (lane == 0 && n > 0) atomicMax(&glen[gi], (n + 7) & ~7);
		}
	}
	csync();               // compact lists, entry ranges, slots and group lengths of every CTA are in place
	const int n_lg = (n_groups > rank) ? (n_groups - rank + C - 1) / C : 0;   // groups dealt to this CTA

	/// M step of one group of 32 chains by one warp: every lane adds its own chain in the
	/// reference's order, streaming the ELL rows through the warp's shared-memory ring with
	/// cp.async, RING_ROWS rows ahead of the fp64 add chain, so that the chain runs at add latency
	/// (16.9 cycles) instead of L2 latency
	auto chain_group = [&](int gi, double *fr_new)
	{
		const uint32_t ring_s = (uint32_t)__cvta_generic_to_shared(
			rings + (size_t)warp * RING_ROWS * 32) + (uint32_t)lane * 16u;
		constexpr int NB = RING_ROWS / 8;      // batches of 8 rows (4 row pairs) in flight
		const int nb = glen[gi] >> 3;
		const char *src = (const char *)(rinc + p.group_base[gi]) + lane * 16;   // + 512 per row pair
		for (int b = 0; b < NB - 1; b++)
		{
			if (b < nb)
			{
#pragma unroll
				for (int k = 0; k < 4; k++)
					cp_async16_cg(ring_s + (uint32_t)(b * 4 + k) * 512u, src + (size_t)(b * 4 + k) * 512);
			}
			cp_async_commit();
		}
		double acc = 0;
		int rb = 0;                         // b mod NB
		for (int b = 0; b < nb; b++)
		{
			const int bn = b + NB - 1;
			if (bn < nb)
			{
				const int rbn = (rb + NB - 1) & (NB - 1);
#pragma unroll
				for (int k = 0; k < 4; k++)
					cp_async16_cg(ring_s + (uint32_t)(rbn * 4 + k) * 512u, src + (size_t)(bn * 4 + k) * 512);
			}
			cp_async_commit();
			cp_async_wait<NB - 1>();
			double2 r[4];
#pragma unroll
			for (int k = 0; k < 4; k++) r[k] = lds_f64x2(ring_s + (uint32_t)(rb * 4 + k) * 512u);
#pragma unroll
			for (int k = 0; k < 4; k++) { acc = __dadd_rn(acc, r[k].x); acc = __dadd_rn(acc, r[k].y); }
			rb = (rb + 1) & (NB - 1);
		}
		cp_async_wait<0>();
		const int r = 32 * gi + lane;
		if (r < n2)
		{
			const double f = __dmul_rn(acc, p.scale);
			const int u = p.hap_sorted[r];
			const size_t o = (size_t)(fr_new - fr0) + u;
			if (C == 1) fr0[o] = f;
			else for (int q = 0; q < C; q++) cluster.map_shared_rank(fr0, q)[o] = f;
		}
	};
	EM_TICK(0);

	double conv_tol = 0, loglik = -1e+30;
	int result = EM_OK, iters = 0;
	for (int iter = 0; iter <= EM_MAX_ITER; iter++)
	{
		const double old_loglik = loglik;
		const double *fr = fr0 + (size_t)(iter & 1) * n2;          // current frequencies
		double *fr_new = fr0 + (size_t)((iter & 1) ^ 1) * n2;
		// ---- E step (:1204-1222) ---------------------------------------------------------------
		// (1) one thread per compatible pair, coalesced: GenoFreq x. Loads are issued in batches
		//     of 8 so that a thread's L2 round trips overlap.
		for (int j0 = jb_lo + tid; j0 < jb_hi; j0 += EM_THREADS * 8)
		{
			int uv[8];
#pragma unroll
			for (int q = 0; q < 8; q++)
			{
				const int j = j0 + q * EM_THREADS;
				uv[q] = (j < jb_hi) ? cuv[j] : 0;       // written in this kernel: no ld.nc
			}
#pragma unroll
			for (int q = 0; q < 8; q++)
			{
				const int j = j0 + q * EM_THREADS;
				if (j < jb_hi)
				{
					const int u = uv[q] & 0xffff, v = (int)((unsigned)uv[q] >> 16);
					xbuf[j] = (u != v) ? __dmul_rn(__dmul_rn(2.0, fr[u]), fr[v]) : __dmul_rn(fr[u], fr[v]);
				}
			}
		}
		if (tid == 0) sh_i[2] = 0;                 // group counter of the M step
		__syncthreads();
		EM_TICK(1);
		// (2) one thread per in-bag entry: sum of its pairs in list order, log-likelihood term,
		//     scale factor count / sum into every CTA of the cluster. An entry has ~6 compatible
		//     pairs: eight predicated loads in flight (one round trip to L2 for most entries), the
		//     next entry's range prefetched.
		double ll = 0;
		{
			int k = k_lo + tid;
			int b = 0, e = 0;
			if (k < k_hi) { b = coff[k]; e = coff[k + 1]; }
			while (k < k_hi)
			{
				const int kn = k + EM_THREADS;
				int bn = 0, en = 0;
				if (kn < k_hi) { bn = coff[kn]; en = coff[kn + 1]; }
				double psum = 0;
				for (int t = b; t < e; t += 8)
				{
					double x[8];
#pragma unroll
					for (int q = 0; q < 8; q++) x[q] = (t + q < e) ? xbuf[t + q] : 0.0;
#pragma unroll
					for (int q = 0; q < 8; q++) if (t + q < e) psum = __dadd_rn(psum, x[q]);
				}
				const double bc = (double)(eg[k] >> 2);
				ll = __dadd_rn(ll, __dmul_rn(bc, log(psum)));
				const double sc = __ddiv_rn(bc, psum);
				if (C == 1) sck[k] = sc;
				else for (int q = 0; q < C; q++) cluster.map_shared_rank(sck, q)[k] = sc;
				k = kn; b = bn; e = en;
			}
		}
		ll = block_sum_f64(ll, scratch);
		if (tid < C) cluster.map_shared_rank(llp, tid)[(iter & 1) * MAX_CLUSTER + rank] = ll;
		csync();               // scale factors and partial log-likelihoods of every CTA have landed
		EM_TICK(2);
		// (3) one thread per compatible pair, coalesced: its contribution r = x * (count / sum) goes
		//     to the ELL slots of both of its haplotypes
		for (int j0 = jb_lo + tid; j0 < jb_hi; j0 += EM_THREADS * 4)
		{
			double x[4];
			int en[4];
			int2 sl[4];
#pragma unroll
			for (int q = 0; q < 4; q++)
			{
				const int j = j0 + q * EM_THREADS;
				const bool ok = j < jb_hi;
				x[q] = ok ? xbuf[j] : 0.0;
				en[q] = ok ? ce[j] : 0;
				sl[q] = ok ? cslot[j] : make_int2(-1, -1);
			}
#pragma unroll
			for (int q = 0; q < 4; q++)
			{
				if (sl[q].x >= 0)
				{
					const double r = __dmul_rn(x[q], sck[en[q]]);
					rinc[sl[q].x] = r;
					rinc[sl[q].y] = r;
				}
			}
		}
		csync();               // the chains read every CTA's contributions
		EM_TICK(3);
		// ---- M step: a warp that owns a ring takes the next-longest group of 32 chains dealt to
		// this CTA -----------------------------------------------------------------------------
		if (warp < p.m_warps)
		{
			for (;;)
			{
				int i = 0;
				if (lane == 0) i = atomicAdd(&sh_i[2], 1);
				i = __shfl_sync(0xffffffffu, i, 0);
				if (i >= n_lg) break;
				chain_group(rank + C * i, fr_new);
			}
		}
		csync();               // new frequencies have landed in every CTA
		EM_TICK(4);
		iters = iter + 1;
		// ---- stopping rule (:1236-1250) with the guard band of em.h; every CTA evaluates the
		// same numbers in the same order ---------------------------------------------------------
		loglik = 0;
		for (int q = 0; q < C; q++) loglik = __dadd_rn(loglik, llp[(iter & 1) * MAX_CLUSTER + q]);
		int f = 0;
		if (iter > 0)
		{
			const double diff = fabs(__dsub_rn(loglik, old_loglik));
			if (fabs(__dsub_rn(diff, conv_tol)) <= em_guard_rel(p.n_entry) * fabs(loglik)) f = 2;
			else if (diff <= conv_tol) f = 1;
		} else {
			conv_tol = __dmul_rn(p.em_reltol, __dadd_rn(fabs(loglik), p.em_reltol));
			if (conv_tol < 0) conv_tol = 0;
		}
		if (f == 2) { result = EM_AMBIGUOUS; break; }
		if (f == 1) break;
	}
	if (rank == 0)
	{
		const double *fin = fr0 + (size_t)(iters & 1) * n2;        // written by the last M step
		double *out = p.out_freq + (size_t)c * n2;
		for (int u = tid; u < n2; u += EM_THREADS) out[u] = fin[u];
		if (tid == 0)
		{
			// longest chain of the M step in rows (= dependent fp64 adds per iteration) and the compatible
			// pairs: the kernel's latency floor and its arithmetic work, for the roofline accounting
			int longest = 0;
			for (int g = 0; g < n_groups; g++) longest = max(longest, glen[g]);
			status[0] = result; status[1] = iters; status[2] = longest; status[3] = coff[p.n_entry];
		}
	}
	if (prof)
	{
		for (int q = 0; q < 5; q++) p.prof[8 * c + q] = (unsigned long long)sh_pt[q];
		p.prof[8 * c + 5] = (unsigned long long)iters;
		p.prof[8 * c + 6] = (unsigned long long)(jb_hi - jb_lo);
	}
#undef EM_TICK
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// em_chain_kernel -- the EM of one candidate as two walks over chains of 4-byte records, state in
// shared memory, 128 threads.
//
// em_kernel (above) streams ~2.7 MB per iteration per candidate through L2 (GenoFreq buffer, 44 k
// scattered 8-byte contribution stores, the padded ELL rows read back by the M step) with 512
// threads that mostly wait at barriers: with two dozen lanes in EM the scattered stores saturate
// the L2 (233 k cycles per iteration inside the 24-lane step against 152 k alone) and the kernel
// HOLDS 30 % of the GPU's SM-time. What counts once the GPU is full is SM-time per
// candidate-iteration = (share of an SM the CTA occupies) x (duration); the arithmetic is tiny
// (22 k products and adds in the E step, 44 k contributions in the M step), everything else is
// latency. So: few threads, no stores, and both steps in ONE form -- a lane walks a chain of records
// in the reference's order and adds one recomputed term per record:
//   E step: chain = an in-bag entry, record = a compatible pair (u, v),
//           term = (2 f_u) f_v (f_u f_v when u == v); then ll += count log(sum), scale = count / sum;
//   M step: chain = a doubled haplotype h, record = {partner haplotype, entry, u == v} of a compatible
//           contribution in (sample, pair, H1-before-H2) order,
//           term = x * scale[entry] with x = (2 f_h) f_partner (f_h f_h when u == v); at the end
//           f_new[h] = sum * (0.5 / n).
// ((2 f_u) f_v and (2 f_v) f_u round the same real number once, 2 f being exact: x is the E step's
// value bit for bit whichever side h is on; the term is em_kernel's r = x * (count / sum).)
// Chains are sorted by length and cut into groups of 32 (one per lane), a group's records
// interleaved so that four rows of a lane are 16 contiguous bytes. A warp takes the next-longest
// group: long groups stream through a private cp.async ring, the terms of batch b + 1 gathered and
// multiplied before the adds of batch b are issued; short groups (<= 8 rows) are loaded into registers
// while the previous group is summed. Haplotypes and entries are addressed by their RANK in that
// order everywhere (frequencies, scale factors, records), so an iteration reads nothing from global
// memory but the records: read-only, 4 bytes a term, coalesced.
// Per round (RoundEM::prepare) every pair and every contribution gets its record with the sum
// s = (u & 1) + (v & 1) of the new SNP's alleles; per candidate the kernel only FILTERS them -- a
// record is compatible with an entry's genotype g when g is missing or g == s (:1157-1180) -- in two
// streaming passes (block scan of the flags, rank inside the chain = prefix - prefix at the chain's
// first record) that write the compatible records into the chains' columns.
// Every accumulator receives the operands em_kernel gives it, in the same order, un-fused: the two
// kernels are bit-identical (tests/test_gpu_parity.py trains the golden models with either).
// ---------------------------------------------------------------------------------------------
namespace {

constexpr int EMC_THREADS = 128;
constexpr int EMC_WARPS = EMC_THREADS / 32;
constexpr int EMC_RING_B = 8;           // batches of 4 rows (512 B per warp) a ring holds (cp.async.wait_group
                                        // with more than ~8 groups pending behaved like wait_all)
constexpr int EMC_MAX_HAP = 16382;      // 14-bit haplotype ranks, 15-bit entry ranks in the records; one
constexpr int EMC_MAX_ENTRY = 32766;    // value of each is the null rank

// E record: rank_u | rank_v << 14 | s << 28.   M record: partner rank | entry rank << 14 | (u == v) << 29 | s << 30
__host__ __device__ inline uint32_t emc_pair_record(int ru, int rv, int s) { return (uint32_t)ru | ((uint32_t)rv << 14) | ((uint32_t)s << 28); }
__host__ __device__ inline uint32_t emc_contrib_record(int rp, int re, bool diag, int s)
{
	return (uint32_t)rp | ((uint32_t)re << 14) | (diag ? (1u << 29) : 0u) | ((uint32_t)s << 30);
}

struct EmcArgs
{
	int n_entry, n_cur, n_samp, total_pairs;
	const int *ib, *boot;
	// haplotype chains (M step): rank order, group bases in record slots, all contributions of the round
	const int *hap_sorted, *hap_rank, *group_base;
	const uint32_t *mrec_all;         // [2 * total_pairs] in chain order (by haplotype, stable)
	const uint16_t *mown;             // [2 * total_pairs] rank of the chain a contribution belongs to
	size_t n_slots;
	// entry chains (E step)
	const int *entry_sorted, *egroup_base;
	const uint32_t *erec_all;         // [total_pairs] in pair order
	const uint16_t *eown;             // [total_pairs] rank of the pair's entry
	size_t n_eslots;
	const double *cur_freq;
	const int8_t *geno_t;
	const int *cand_snp;
	uint32_t *rec;                    // [m][n_slots]
	uint32_t *erec;                   // [m][n_eslots]
	double *out_freq;
	int *out_status;
	double scale, em_reltol;
	unsigned long long *acct;
	int acct_w;
	int len_words;                    // 16-bit words of the chain-length / batch-schedule region (even)
	unsigned long long *prof;         // [m][8]: cycles of set-up, E steps, M steps; iterations (HIBAG_B200_EM_DEBUG)
};

/// record slot of row i of lane l in a group that starts at base (32-bit slots)
__device__ __forceinline__ size_t emc_slot(int base, int i, int l)
{
	return (size_t)base + 128 * (size_t)(i >> 2) + 4 * (size_t)l + (size_t)(i & 3);
}

/// the term of an E-step record (a compatible pair): GenoFreq (:1208-1213). The null record
/// (both ranks = n2, where the frequency vector holds 0.0) gives +0.0.
struct EmcPairTerm
{
	const double *fr;                 // by rank
	__device__ __forceinline__ void gather(uint32_t rc, double &a, double &b) const
	{
		a = fr[rc & 0x3fffu]; b = fr[(rc >> 14) & 0x3fffu];
	}
	__device__ __forceinline__ double combine(uint32_t rc, double a, double b) const
	{
		const double t = ((rc & 0x3fffu) != ((rc >> 14) & 0x3fffu)) ? __dmul_rn(2.0, a) : a;      // 2 f_u is exact
		return __dmul_rn(t, b);
	}
};

/// the term of an M-step record: GenoFreq * count / sum (:1224-1232). The null record (partner rank
/// n2, entry rank n_entry, where the tables hold 0.0) gives +0.0.
struct EmcContribTerm
{
	const double *fr, *sck;           // by rank
	double fh2, xhh;                  // 2 f_h and f_h f_h of the lane's haplotype
	__device__ __forceinline__ void gather(uint32_t rc, double &a, double &b) const
	{
		a = fr[rc & 0x3fffu]; b = sck[(rc >> 14) & 0x7fffu];
	}
	__device__ __forceinline__ double combine(uint32_t rc, double a, double b) const
	{
		const double x = (rc & (1u << 29)) ? xhh : __dmul_rn(fh2, a);
		return __dmul_rn(x, b);
	}
};

/// terms of four consecutive rows of a lane, branch-free, in two halves: the eight shared-memory
/// gathers, then the products. (Rows past the end of a lane's chain hold null records: +0.0, and
/// s + 0.0 == s for the non-negative sums here.)
template <class Term>
__device__ __forceinline__ void emc_gather4(const int4 &q4, const Term &term, double (&a)[4], double (&b)[4])
{
	const uint32_t rc[4] = { (uint32_t)q4.x, (uint32_t)q4.y, (uint32_t)q4.z, (uint32_t)q4.w };
#pragma unroll
	for (int q = 0; q < 4; q++) term.gather(rc[q], a[q], b[q]);
}

template <class Term>
__device__ __forceinline__ void emc_combine4(const int4 &q4, const Term &term, const double (&a)[4],
	const double (&b)[4], double (&rr)[4])
{
	const uint32_t rc[4] = { (uint32_t)q4.x, (uint32_t)q4.y, (uint32_t)q4.z, (uint32_t)q4.w };
#pragma unroll
	for (int q = 0; q < 4; q++) rr[q] = term.combine(rc[q], a[q], b[q]);
}

/// The groups of one step dealt to this warp (list, longest first), as ONE stream of record batches
/// through the warp's ring (sched: where every batch of the stream lies): batch k goes to ring slot k mod 8 and is requested nine
/// batches before it is consumed, whatever groups the batches in between belong to -- a group of two
/// batches does not wait for L2 any more than the middle of a long one. Inside a group three stages
/// are in flight: the records of batch j + 2 are read from the ring and the terms of batch j + 1
/// gathered and multiplied while the adds of batch j -- the only dependent chain -- are issued.
/// done(group, chain sum) is called for every group of the list, also the empty ones.
template <class TermOf, class Done>
__device__ __forceinline__ void emc_stream(const uint16_t *list, int cnt, const uint16_t *sched, int n_sched,
	const uint32_t *records, const int *glen, uint32_t ring_s, int lane, TermOf term_of, Done done)
{
	// issuer: batch ki of the warp's schedule (the batch's first record slot / 128) goes to ring slot ki mod 8
	int ki = 0;
	const char *lane_src = (const char *)records + lane * 16;
	auto issue = [&]()
	{
		if (ki < n_sched)
			cp_async16_cg(ring_s + (uint32_t)(ki & (EMC_RING_B - 1)) * 512u, lane_src + (size_t)sched[ki] * 512);
		ki++;
		cp_async_commit();                  // (an empty group when the stream has ended: the count stays uniform)
	};
	// Three batches live in registers (terms of batch j, records of batches j + 1 and j + 2), so the
	// 8-slot ring can run NINE batches ahead: batch j + 9 goes into the slot of batch j + 1, which has
	// been read. Step j waits until at most 6 groups are pending = batch j + 3 has landed.
#pragma unroll 1
	for (int b = 0; b < EMC_RING_B; b++) issue();
	cp_async_wait<EMC_RING_B - 3>();       // stream batches 0, 1 and 2 have landed
	int4 q4 = lds_i32x4(ring_s);           // records of the next batch to be turned into terms
	int slot_n = 1;                        // ring slot of the stream batch after that one
	issue();                               // batch 8 into the slot batch 0 was just read from
#pragma unroll 1
	for (int idx = 0; idx < cnt; idx++)
	{
		const int g = (int)list[idx];
		const int nb = (glen[g] + 3) >> 2;
		double acc = 0;
		if (nb > 0)
		{
			const auto term = term_of(g);
			double rr[4];
			{
				double ga[4], gb[4];
				emc_gather4(q4, term, ga, gb);
				emc_combine4(q4, term, ga, gb, rr);
			}
			q4 = lds_i32x4(ring_s + (uint32_t)slot_n * 512u);
			slot_n = (slot_n + 1) & (EMC_RING_B - 1);
			// all batches but the last. The ring read of the batch after next and the gathers of the next
			// batch go first: the asm statements below are memory barriers for the compiler, and loads
			// placed after them would wait behind the adds
#pragma unroll 2
			for (int b = 0; b + 1 < nb; b++)
			{
				const int4 q4n = lds_i32x4(ring_s + (uint32_t)slot_n * 512u);      // (landed: see the wait below)
				slot_n = (slot_n + 1) & (EMC_RING_B - 1);
				double ga[4], gb[4];
				emc_gather4(q4, term, ga, gb);
				issue();
				cp_async_wait<EMC_RING_B - 2>();               // the stream batch three ahead has landed
				double rn[4];
				emc_combine4(q4, term, ga, gb, rn);
#pragma unroll
				for (int q = 0; q < 4; q++) { acc = __dadd_rn(acc, rr[q]); rr[q] = rn[q]; }
				q4 = q4n;
			}
			// the last batch: q4 already holds the first records of the next group
			issue();
			cp_async_wait<EMC_RING_B - 2>();
#pragma unroll
			for (int q = 0; q < 4; q++) acc = __dadd_rn(acc, rr[q]);
		}
		done(g, acc);
	}
	cp_async_wait<0>();
}

/// Streaming filter of one record array (all pairs, or all contributions, of the round) into the
/// chains' columns: a record of chain `own` is kept when the genotype of its entry is missing or
/// equals the record's allele sum; its row is its rank among the kept records of its chain. 1 024
/// records per step (8 per thread, contiguous), the next step's loads in flight.
/// pstart [n_chains] ints (scratch), len [n_chains] (zeroed by the caller), glen [n_groups] (zeroed).
template <bool M_RECORDS>
__device__ __forceinline__ int emc_filter(const uint32_t *__restrict__ all, const uint16_t *__restrict__ own, int n,
	const int *eg, const int *gbase, uint32_t *out, int *pstart, uint16_t *len, int *glen, int *wsum)
{
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	int running = 0;
	uint4 it0 = make_uint4(0, 0, 0, 0), it1 = it0, ow = it0;
	int prev_owner = -1;
	auto load = [&](int base)
	{
		const int i0 = base + 8 * tid;
		if (i0 < n)       // (the arrays are padded to a multiple of 8 records)
		{
			it0 = __ldg((const uint4 *)(all + i0)); it1 = __ldg((const uint4 *)(all + i0 + 4));
			ow = __ldg((const uint4 *)(own + i0));
			prev_owner = (i0 > 0) ? (int)__ldg(own + i0 - 1) : -1;
		}
	};
	load(0);
	for (int base = 0; base < n; base += 8 * EMC_THREADS)
	{
		const uint32_t rc[8] = { it0.x, it0.y, it0.z, it0.w, it1.x, it1.y, it1.z, it1.w };
		const uint32_t o2[4] = { ow.x, ow.y, ow.z, ow.w };
		const int pv = prev_owner;
		const int i0 = base + 8 * tid;
		if (base + 8 * EMC_THREADS < n) load(base + 8 * EMC_THREADS);
		int owner[8];
		bool keep[8];
		int cnt = 0;
#pragma unroll
		for (int q = 0; q < 8; q++)
		{
			owner[q] = (int)((o2[q >> 1] >> (16 * (q & 1))) & 0xffffu);
			const int g = eg[M_RECORDS ? (int)((rc[q] >> 14) & 0x7fffu) : owner[q]];
			const int sg = (int)(rc[q] >> (M_RECORDS ? 30 : 28)) & 3;
			keep[q] = (i0 + q < n) && (g == 3 || g == sg);
			cnt += keep[q] ? 1 : 0;
		}
		int incl = cnt;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1)
		{
			const int y = __shfl_up_sync(0xffffffffu, incl, o);
			if (lane >= o) incl += y;
		}
		if (lane == 31) wsum[warp] = incl;
		__syncthreads();
		int pre = running + incl - cnt;
		int total = 0;
#pragma unroll
		for (int w = 0; w < EMC_WARPS; w++) { const int v = wsum[w]; if (w < warp) pre += v; total += v; }
		// the first record of a chain notes the prefix its chain starts at
		{
			int pq = pre, last = pv;
#pragma unroll
			for (int q = 0; q < 8; q++)
			{
				if (i0 + q < n && owner[q] != last) pstart[owner[q]] = pq;
				last = owner[q];
				pq += keep[q] ? 1 : 0;
			}
		}
		__syncthreads();
		{
			int pq = pre, last = pv;
#pragma unroll
			for (int q = 0; q < 8; q++)
			{
				if (i0 + q < n)
				{
					if (owner[q] != last && last >= 0)
					{
						// the previous chain ended just before this record
						const int l = pq - pstart[last];
						len[last] = (uint16_t)l;
						if (l > 0) atomicMax(&glen[last >> 5], l);
					}
					if (keep[q])
						out[emc_slot(gbase[owner[q] >> 5], pq - pstart[owner[q]], owner[q] & 31)] = rc[q];
					if (i0 + q == n - 1)
					{
						const int l = pq + (keep[q] ? 1 : 0) - pstart[owner[q]];
						len[owner[q]] = (uint16_t)l;
						if (l > 0) atomicMax(&glen[owner[q] >> 5], l);
					}
				}
				last = owner[q];
				pq += keep[q] ? 1 : 0;
			}
		}
		running += total;
	}
	return running;
}

__global__ void __launch_bounds__(EMC_THREADS, 4) em_chain_kernel(const EmcArgs p)
{
	SmAcct acct_scope(p.acct, SM_ACCT_EM, (unsigned)p.acct_w);
	SmAcct acct_cta(p.acct, SM_ACCT_EM_CTA, 1024u);
	extern __shared__ double em_smem[];
	const int n2 = 2 * p.n_cur;
	const int nf = (n2 + 2) & ~1;                            // stride of a frequency buffer: n2 ranks + the null slot
	const int n_groups = (n2 + 31) >> 5, n_egroups = (p.n_entry + 31) >> 5;
	double *fr0 = em_smem;                                   // [2][nf] frequencies by rank, double-buffered; [n2] = 0.0
	double *scratch = fr0 + 2 * (size_t)nf;                  // [40]
	double *sck = scratch + 40;                              // [n_entry + 1] count / sum by rank, [n_entry] = 0.0 (set-up: genotypes, prefixes)
	uint32_t *rings = (uint32_t *)(sck + ((p.n_entry + 2) & ~1));   // [EMC_WARPS][EMC_RING_B][128] (16-byte aligned)
	int *gbase = (int *)(rings + EMC_WARPS * EMC_RING_B * 128);     // [n_groups] first record slot of the group
	int *egbase = gbase + n_groups;                          // [n_egroups]
	int *glen = egbase + n_egroups;                          // [n_groups] longest compatible chain of the group
	int *eglen = glen + n_groups;                            // [n_egroups]
	uint16_t *mlist = (uint16_t *)(eglen + n_egroups);       // [n_groups] the M step's groups, warp by warp
	uint16_t *elist = mlist + ((n_groups + 1) & ~1);         // [n_egroups] the E step's
	// set-up: chain lengths by rank; afterwards the same words hold the warps' batch schedules
	uint16_t *clen = elist + ((n_egroups + 1) & ~1);         // [n2] compatible contributions per chain
	uint16_t *elen = clen + ((n2 + 1) & ~1);                 // [n_entry] compatible pairs per entry
	uint16_t *sched_m = clen;                                // [<= n_slots / 128] first slot / 128 of every batch, warp by warp
	uint16_t *sched_e = sched_m + p.n_slots / 128;           // [<= n_eslots / 128]
	uint8_t *ebc = (uint8_t *)(clen + p.len_words);          // [n_entry] bootstrap count, by rank
	int *eg = (int *)sck;                                    // set-up: genotype of the entry (3 = missing), by rank
	int *pstart_e = eg + p.n_entry;                          // set-up: prefix at the first pair of the entry
	int *pstart_m = (int *)(fr0 + nf);                       // set-up: the same per haplotype chain (second frequency buffer)
	__shared__ int sh_i[3];                                  // allele count, valid count, oversized bootstrap count
	__shared__ int sh_w[EMC_WARPS];
	__shared__ int sh_cnt[2][EMC_WARPS + 1];                 // first group of every warp in mlist / elist
	__shared__ int sh_sch[2][EMC_WARPS + 1];                 // first batch of every warp in sched_m / sched_e

	const int c = blockIdx.x;
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int8_t *col = p.geno_t + (size_t)p.cand_snp[c] * p.n_samp;
	uint32_t *rec = p.rec + (size_t)c * p.n_slots;
	uint32_t *erec = p.erec + (size_t)c * p.n_eslots;
	int *status = p.out_status + 4 * c;

	// allele frequency of the new SNP in the bootstrap sample (:1136-1151), integers
	{
		int ac = 0, vc = 0, big = 0;
		for (int r = tid; r < p.n_entry; r += EMC_THREADS)
		{
			const int s = p.ib[__ldg(p.entry_sorted + r)];
			const int g = col[s];
			const int b = p.boot[s];
			eg[r] = (0 <= g && g <= 2) ? g : 3;
			ebc[r] = (uint8_t)b;
			elen[r] = 0;
			big |= (b > 255) ? 1 : 0;
			if (0 <= g && g <= 2) { ac += g * b; vc += 2 * b; }
		}
		if (tid < 3) sh_i[tid] = 0;
		for (int g = tid; g < n_groups; g += EMC_THREADS) { gbase[g] = p.group_base[g]; glen[g] = 0; }
		for (int g = tid; g < n_egroups; g += EMC_THREADS) { egbase[g] = p.egroup_base[g]; eglen[g] = 0; }
		for (int r = tid; r < n2; r += EMC_THREADS) clen[r] = 0;
		__syncthreads();
#pragma unroll
		for (int o = 16; o > 0; o >>= 1)
		{
			ac += __shfl_xor_sync(0xffffffffu, ac, o);
			vc += __shfl_xor_sync(0xffffffffu, vc, o);
			big |= __shfl_xor_sync(0xffffffffu, big, o);
		}
		if (lane == 0) { atomicAdd(&sh_i[0], ac); atomicAdd(&sh_i[1], vc); atomicOr(&sh_i[2], big); }
		__syncthreads();
	}
	const int allele_cnt = sh_i[0], valid_cnt = sh_i[1];
	if (allele_cnt == 0 || allele_cnt == valid_cnt)
	{
		if (tid == 0) { status[0] = EM_INVALID; status[1] = 0; status[2] = 0; status[3] = 0; }
		return;
	}
	if (sh_i[2])
	{
		// a bootstrap count that does not fit the byte table: the host re-estimates this candidate
		if (tid == 0) { status[0] = EM_AMBIGUOUS; status[1] = 0; status[2] = -1; status[3] = 0; }
		return;
	}
	// doubled list, initial frequencies (:444-459), stored by rank
	{
		const double af = __ddiv_rn((double)allele_cnt, (double)valid_cnt);
		const double q0 = __dsub_rn(1.0, af), q1 = af;
		for (int k = tid; k < p.n_cur; k += EMC_THREADS)
		{
			const double f = p.cur_freq[k];
			fr0[__ldg(p.hap_rank + 2 * k)] = __dadd_rn(__dmul_rn(q0, f), EM_INIT_VAL_FRAC);
			fr0[__ldg(p.hap_rank + 2 * k + 1)] = __dadd_rn(__dmul_rn(q1, f), EM_INIT_VAL_FRAC);
		}
		if (tid == 0) fr0[n2] = 0.0;            // the null slot of the first buffer (the second one: after the set-up)
	}
	// ---- the compatible pairs (:1157-1180) of every entry and the compatible contributions of every
	// haplotype, in the reference's order, into the chains' columns ------------------------------------
	const int n_compat = emc_filter<false>(p.erec_all, p.eown, p.total_pairs, eg, egbase, erec, pstart_e, elen, eglen, sh_w);
	__syncthreads();
	emc_filter<true>(p.mrec_all, p.mown, 2 * p.total_pairs, eg, gbase, rec, pstart_m, clen, glen, sh_w);
	__syncthreads();
	// null records behind the end of every chain, up to the last row of its group's last batch
	{
		const uint32_t null_e = emc_pair_record(n2, n2, 0), null_m = emc_contrib_record(n2, p.n_entry, false, 0);
		for (int r = tid; r < 32 * n_egroups; r += EMC_THREADS)
		{
			const int end = ((eglen[r >> 5] + 3) >> 2) << 2;
			for (int i = (r < p.n_entry) ? (int)elen[r] : 0; i < end; i++) erec[emc_slot(egbase[r >> 5], i, r & 31)] = null_e;
		}
		for (int r = tid; r < 32 * n_groups; r += EMC_THREADS)
		{
			const int end = ((glen[r >> 5] + 3) >> 2) << 2;
			for (int i = (r < n2) ? (int)clen[r] : 0; i < end; i++) rec[emc_slot(gbase[r >> 5], i, r & 31)] = null_m;
		}
	}
	// the groups of either step dealt to the warps: in rank order (longest first) each group goes to the
	// warp with the fewest batches so far
	if (tid < 2)
	{
		const int ng = tid ? n_egroups : n_groups;
		const int *len_g = tid ? eglen : glen;
		uint16_t *lst = tid ? elist : mlist;
		int load[EMC_WARPS], cnt[EMC_WARPS];
		for (int w = 0; w < EMC_WARPS; w++) { load[w] = 0; cnt[w] = 0; }
		// first pass: sizes; second pass: positions (the lists are stored warp after warp)
		for (int pass = 0; pass < 2; pass++)
		{
			int pos[EMC_WARPS];
			if (pass)
			{
				int run = 0;
				for (int w = 0; w < EMC_WARPS; w++) { sh_cnt[tid][w] = run; pos[w] = run; run += cnt[w]; load[w] = 0; }
				sh_cnt[tid][EMC_WARPS] = run;
			}
			for (int g = 0; g < ng; g++)
			{
				int best = 0;
				for (int w = 1; w < EMC_WARPS; w++) if (load[w] < load[best]) best = w;
				load[best] += ((len_g[g] + 3) >> 2) + 1;         // (+1: a group costs about a batch by itself)
				if (pass) lst[pos[best]++] = (uint16_t)g; else cnt[best]++;
			}
		}
	}
	if (tid == 0) { sck[p.n_entry] = 0.0; fr0[nf + n2] = 0.0; }
	__threadfence();           // the records are read back through L2 by other warps of this CTA
	__syncthreads();
	// the warps' batch schedules (over the chain lengths, which nobody reads any more)
	if (tid < 2)
	{
		const int *len_g = tid ? eglen : glen;
		const int *base_g = tid ? egbase : gbase;
		const uint16_t *lst = tid ? elist : mlist;
		uint16_t *sch = tid ? sched_e : sched_m;
		int k = 0;
		for (int w = 0; w < EMC_WARPS; w++)
		{
			sh_sch[tid][w] = k;
			for (int idx = sh_cnt[tid][w]; idx < sh_cnt[tid][w + 1]; idx++)
			{
				const int g = (int)lst[idx];
				const int nb = (len_g[g] + 3) >> 2, b0 = base_g[g] >> 7;
				for (int b = 0; b < nb; b++) sch[k++] = (uint16_t)(b0 + b);
			}
		}
		sh_sch[tid][EMC_WARPS] = k;
	}
	__syncthreads();
	const bool prof = (p.prof != nullptr) && tid == 0;
	long long t_last = prof ? clock64() : 0, t_e = 0, t_m = 0, t_e1 = 0;
	if (prof) p.prof[8 * c + 0] = (unsigned long long)(t_last - acct_cta.t0);
	const uint32_t ring_s = (uint32_t)__cvta_generic_to_shared(rings + (size_t)warp * EMC_RING_B * 128) +
		(uint32_t)lane * 16u;

	double conv_tol = 0, loglik = -1e+30;
	int result = EM_OK, iters = 0;
	for (int iter = 0; iter <= EM_MAX_ITER; iter++)
	{
		const double old_loglik = loglik;
		const double *fr = fr0 + (size_t)(iter & 1) * nf;
		double *fr_new = fr0 + (size_t)((iter & 1) ^ 1) * nf;
		// ---- E step (:1204-1222), pass 1: a warp takes the next-longest group of 32 entries; every lane
		// sums the GenoFreq of its entry's compatible pairs in list order ---------------------------------
		emc_stream(elist + sh_cnt[1][warp], sh_cnt[1][warp + 1] - sh_cnt[1][warp], sched_e + sh_sch[1][warp],
			sh_sch[1][warp + 1] - sh_sch[1][warp], erec, eglen, ring_s, lane,
			[&](int) { return EmcPairTerm{ fr }; },
			[&](int g, double psum) { const int r = 32 * g + lane; if (r < p.n_entry) sck[r] = psum; });
		__syncthreads();
		if (prof) { const long long n_ = clock64(); t_e1 += n_ - t_last; }
		// pass 2: a thread per entry, four at a time (four independent log / divide sequences):
		// log-likelihood term and scale factor count / sum
		double ll = 0;
		for (int r0 = tid; r0 < p.n_entry; r0 += 4 * EMC_THREADS)
		{
			double ps[4], bc[4], lg[4];
#pragma unroll
			for (int q = 0; q < 4; q++)
			{
				const int r = r0 + q * EMC_THREADS;
				const bool ok = r < p.n_entry;
				ps[q] = ok ? sck[r] : 1.0;
				bc[q] = ok ? (double)ebc[r] : 0.0;
			}
#pragma unroll
			for (int q = 0; q < 4; q++) lg[q] = log(ps[q]);
#pragma unroll
			for (int q = 0; q < 4; q++)
			{
				const int r = r0 + q * EMC_THREADS;
				if (r < p.n_entry)
				{
					ll = __dadd_rn(ll, __dmul_rn(bc[q], lg[q]));
					sck[r] = __ddiv_rn(bc[q], ps[q]);
				}
			}
		}
		ll = block_sum_f64(ll, scratch);           // (its barriers publish the scale factors)
		if (prof) { const long long n_ = clock64(); t_e += n_ - t_last; t_last = n_; }
		// ---- M step: a warp takes the next-longest group of 32 haplotype chains ---------------------------
		emc_stream(mlist + sh_cnt[0][warp], sh_cnt[0][warp + 1] - sh_cnt[0][warp], sched_m + sh_sch[0][warp],
			sh_sch[0][warp + 1] - sh_sch[0][warp], rec, glen, ring_s, lane,
			[&](int g) {
				const int r = 32 * g + lane;
				const double fh = (r < n2) ? fr[r] : 0.0;
				return EmcContribTerm{ fr, sck, __dmul_rn(2.0, fh), __dmul_rn(fh, fh) };      // 2 f_h is exact
			},
			[&](int g, double acc) { const int r = 32 * g + lane; if (r < n2) fr_new[r] = __dmul_rn(acc, p.scale); });
		__syncthreads();
		if (prof) { const long long n_ = clock64(); t_m += n_ - t_last; t_last = n_; }
		iters = iter + 1;
		// ---- stopping rule (:1236-1250) with the guard band (em_guard_rel) ---------------------------
		loglik = ll;
		int f = 0;
		if (iter > 0)
		{
			const double diff = fabs(__dsub_rn(loglik, old_loglik));
			if (fabs(__dsub_rn(diff, conv_tol)) <= em_guard_rel(p.n_entry) * fabs(loglik)) f = 2;
			else if (diff <= conv_tol) f = 1;
		} else {
			conv_tol = __dmul_rn(p.em_reltol, __dadd_rn(fabs(loglik), p.em_reltol));
			if (conv_tol < 0) conv_tol = 0;
		}
		if (f == 2) { result = EM_AMBIGUOUS; break; }
		if (f == 1) break;
	}
	{
		const double *fin = fr0 + (size_t)(iters & 1) * nf;
		double *out = p.out_freq + (size_t)c * n2;
		for (int r = tid; r < n2; r += EMC_THREADS) out[__ldg(p.hap_sorted + r)] = fin[r];
		if (tid == 0)
		{
			int longest = 0;
			for (int g = 0; g < n_groups; g++) longest = max(longest, glen[g]);
			status[0] = result; status[1] = iters; status[2] = longest; status[3] = n_compat;
			if (prof)
			{
				int longest_e = 0;
				for (int g = 0; g < n_egroups; g++) longest_e = max(longest_e, eglen[g]);
				p.prof[8 * c + 1] = (unsigned long long)t_e; p.prof[8 * c + 2] = (unsigned long long)t_m;
				p.prof[8 * c + 3] = (unsigned long long)iters; p.prof[8 * c + 4] = (unsigned long long)longest_e;
				unsigned long long se = 0, sm_ = 0, le = 0, lm = 0;
				for (int g = 0; g < n_egroups; g++) { const int nb = (eglen[g] + 3) >> 2; se += nb; le += nb > 4; }
				for (int g = 0; g < n_groups; g++) { const int nb = (glen[g] + 3) >> 2; sm_ += nb; lm += nb > 4; }
				p.prof[8 * c + 5] = se | (le << 32); p.prof[8 * c + 6] = sm_ | (lm << 32);
				p.prof[8 * c + 7] = (unsigned long long)t_e1;
			}
		}
	}
}

/// per round: the record of every pair (pair order) and of every contribution (chain order) with the
/// ranks of its haplotypes and entry, and the chain each belongs to
__global__ void emc_pair_records_kernel(const int *__restrict__ p1, const int *__restrict__ p2,
	const int *__restrict__ off, int n_entry, const int *__restrict__ hap_rank,
	const int *__restrict__ entry_rank, uint32_t *erec_all, uint16_t *eown)
{
	const int k = blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= n_entry) return;
	const int re = entry_rank[k];
	for (int t = off[k]; t < off[k + 1]; t++)
	{
		const int u = p1[t], v = p2[t];
		erec_all[t] = emc_pair_record(hap_rank[u], hap_rank[v], (u & 1) + (v & 1));
		eown[t] = (uint16_t)re;
	}
}

__global__ void emc_contrib_records_kernel(const int *__restrict__ key_sorted, const int *__restrict__ val_sorted,
	int n_inc, const int *__restrict__ pairs4, const int *__restrict__ hap_rank,
	const int *__restrict__ entry_rank, uint32_t *mrec_all, uint16_t *mown)
{
	const int q = blockIdx.x * blockDim.x + threadIdx.x;
	if (q >= n_inc) return;
	const int h = key_sorted[q], e = val_sorted[q];
	const int uv = pairs4[4 * (size_t)(e >> 1)], k = pairs4[4 * (size_t)(e >> 1) + 1];
	const int u = uv & 0xffff, v = (int)((unsigned)uv >> 16);
	const int partner = (e & 1) ? u : v;           // side 1: this haplotype is H2
	mrec_all[q] = emc_contrib_record(hap_rank[partner], entry_rank[k], u == v, (u & 1) + (v & 1));
	mown[q] = (uint16_t)hap_rank[h];
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// HIBAG_B200_EM_PROF=1: per-phase clock64 totals of the EM kernel over the process, printed at exit
namespace {
struct EmProf
{
	std::mutex mu;
	unsigned long long t[8] = {0, 0, 0, 0, 0, 0, 0, 0};
	unsigned long long cands = 0, launches = 0, slots = 0, pairs = 0;
	bool registered = false;
	static void dump();
} g_em_prof;
void EmProf::dump()
{
	EmProf &g = g_em_prof;
	if (!g.cands) return;
	const double it = (double)g.t[5];
	fprintf(stderr, "em prof: %llu launches, %llu candidates, %.0f iterations (mean %.1f), compat pairs/cand %.0f, "
		"all pairs/cand %.0f, ELL slots/cand %.0f | kcycles per candidate: setup %.1f | per iteration: "
		"pairs %.2f entries %.2f slots %.2f chains %.2f (sum %.2f)\n",
		g.launches, g.cands, it, it / g.cands, (double)g.t[6] / g.cands, (double)g.pairs / g.cands,
		(double)g.slots / g.cands, g.t[0] * 1e-3 / g.cands, g.t[1] * 1e-3 / it, g.t[2] * 1e-3 / it,
		g.t[3] * 1e-3 / it, g.t[4] * 1e-3 / it, (g.t[1] + g.t[2] + g.t[3] + g.t[4]) * 1e-3 / it);
}
}  // namespace

// HIBAG_B200_EM_GATE=n: at most n EM launches (one per lane and selection round, ~23 CTAs each) in
// flight per process. Every candidate CTA streams ~0.8 MB of scratch per iteration; with 24 lanes
// all in EM that is 440 MB -- far beyond the 126 MB L2 -- and the scattered 8-byte stores of the
// contribution pass go to HBM. 0 (default) = no limit.
namespace {
struct EmGate
{
	std::mutex mu;
	std::condition_variable cv;
	int limit = 0, in_flight = 0;
	EmGate() { if (const char *e = getenv("HIBAG_B200_EM_GATE")) limit = std::max(0, atoi(e)); }
	void enter()
	{
		if (limit <= 0) return;
		std::unique_lock<std::mutex> lk(mu);
		cv.wait(lk, [&]() { return in_flight < limit; });
		in_flight++;
	}
	void leave()
	{
		if (limit <= 0) return;
		{ std::lock_guard<std::mutex> lk(mu); in_flight--; }
		cv.notify_one();
	}
} g_em_gate;
}  // namespace

RoundEM::RoundEM() { current_device(); h_total_.ensure(8); }
RoundEM::~RoundEM() {}

void RoundEM::prepare(const HapList &cur, const uint32_t *s1, const uint32_t *s2, int stride,
	const int *a1, const int *a2, const int *ib, int n_entry, const int *boot, cudaStream_t st)
{
	if (cur.n_snp >= HIBAG_B200_MAX_SNP)
		throw std::runtime_error("prepare: too many SNP markers in the classifier");
	if (!supports((int)cur.h.size(), n_entry))
		throw std::runtime_error("prepare: list too large for the device EM");
	n_entry_ = n_entry; n_cur_ = (int)cur.h.size(); n2_ = 2 * n_cur_; n_snp_ = cur.n_snp;
	ib_ = ib; boot_ = boot;
	const int n_hla = (int)cur.len.size();
	// stage: packed haplotypes, frequencies, allele starts -> one pinned block, three copies
	const size_t b_hap = sizeof(uint64_t) * 2 * (size_t)n_cur_, b_fr = sizeof(double) * (size_t)n_cur_;
	const size_t b_st = sizeof(int) * (size_t)(n_hla + 1);
	unsigned char *h = h_stage_.ensure(b_hap + b_fr + b_st + 64);
	uint64_t *hh = (uint64_t *)h;
	double *hf = (double *)(h + b_hap);
	int *hs = (int *)(h + b_hap + b_fr);
	for (int k = 0; k < n_cur_; k++)
	{
		hh[2 * k] = (uint64_t)cur.h[k].packed[0]; hh[2 * k + 1] = (uint64_t)cur.h[k].packed[1];
		hf[k] = cur.h[k].freq;
	}
	hs[0] = 0;
	for (int a = 0; a < n_hla; a++) hs[a + 1] = hs[a] + cur.len[a];
	d_hap_.ensure(2 * (size_t)n_cur_ + 2); d_curfreq_.ensure(n_cur_ + 1); d_start_.ensure(n_hla + 1);
	HB_CUDA(cudaMemcpyAsync(d_hap_.get(), hh, b_hap, cudaMemcpyHostToDevice, st));
	HB_CUDA(cudaMemcpyAsync(d_curfreq_.get(), hf, b_fr, cudaMemcpyHostToDevice, st));
	HB_CUDA(cudaMemcpyAsync(d_start_.get(), hs, b_st, cudaMemcpyHostToDevice, st));
	h2d_bytes += b_hap + b_fr + b_st;

	d_cnt_.ensure(n_entry + 2); d_off_.ensure(n_entry + 1); d_mind_.ensure(n_entry + 1);
	MatchArgs m;
	memset(&m, 0, sizeof(m));
	m.hap = d_hap_.get(); m.start = d_start_.get();
	m.s1 = s1; m.s2 = s2; m.stride = stride; m.a1 = a1; m.a2 = a2; m.ib = ib;
	m.n_entry = n_entry; m.n_snp = cur.n_snp;
	m.cnt = d_cnt_.get(); m.mind = d_mind_.get(); m.off = d_off_.get();
	if (n_entry <= 0) throw std::runtime_error("prepare: no in-bag samples");
	const int blocks = (n_entry + 127) / 128;
	HB_CUDA(cudaMemsetAsync(d_cnt_.get() + n_entry, 0, 2 * sizeof(int), st));
	m.empty_flag = d_cnt_.get() + n_entry + 1;
	haplomatch_kernel<0><<<blocks, 128, 0, st>>>(m);
	HB_CUDA(cudaGetLastError());
	size_t tmp_bytes = 0;
	HB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, d_cnt_.get(), d_off_.get(), n_entry + 1, st));
	d_tmp_.ensure(tmp_bytes + 16);
	HB_CUDA(cub::DeviceScan::ExclusiveSum(d_tmp_.get(), tmp_bytes, d_cnt_.get(), d_off_.get(), n_entry + 1, st));
	HB_CUDA(cudaMemcpyAsync(h_total_.get(), d_off_.get() + n_entry, sizeof(int), cudaMemcpyDeviceToHost, st));
	HB_CUDA(cudaMemcpyAsync(h_total_.get() + 3, d_cnt_.get() + n_entry + 1, sizeof(int), cudaMemcpyDeviceToHost, st));
	stream_sync_blocking(st);
	total_pairs_ = (size_t)h_total_.get()[0];
	has_empty_entry_ = h_total_.get()[3] != 0;
	d2h_bytes += sizeof(int);
	launches += 2;
	if (total_pairs_ == 0) return;
	if (total_pairs_ > (size_t)500000000)
		throw std::runtime_error("prepare: too many haplotype pairs");

	d_p1_.ensure(total_pairs_); d_p2_.ensure(total_pairs_);
	m.p1 = d_p1_.get(); m.p2 = d_p2_.get();
	haplomatch_kernel<1><<<blocks, 128, 0, st>>>(m);
	HB_CUDA(cudaGetLastError());

	// incidence lists: stable sort of (haplotype, contribution) keeps every haplotype's
	// contributions in (sample, pair, H1-before-H2) order
	const int n_inc = (int)(2 * total_pairs_);
	d_key_.ensure(n_inc); d_val_.ensure(n_inc); d_key2_.ensure(n_inc); d_val2_.ensure(n_inc);
	d_inc_off_.ensure(n2_ + 2);
	incidence_fill_kernel<<<((int)total_pairs_ + 255) / 256, 256, 0, st>>>(d_p1_.get(), d_p2_.get(),
		(int)total_pairs_, d_key_.get(), d_val_.get());
	HB_CUDA(cudaGetLastError());
	int end_bit = 1;
	while ((1 << end_bit) < n2_ + 1) end_bit++;
	tmp_bytes = 0;
	HB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, d_key_.get(), d_key2_.get(),
		d_val_.get(), d_val2_.get(), n_inc, 0, end_bit, st));
	d_tmp_.ensure(tmp_bytes + 16);
	HB_CUDA(cub::DeviceRadixSort::SortPairs(d_tmp_.get(), tmp_bytes, d_key_.get(), d_key2_.get(),
		d_val_.get(), d_val2_.get(), n_inc, 0, end_bit, st));
	incidence_offsets_kernel<<<(n2_ + 1 + 255) / 256, 256, 0, st>>>(d_key2_.get(), n_inc, n2_,
		d_inc_off_.get());
	HB_CUDA(cudaGetLastError());
	// ELL layout: haplotypes by decreasing chain length, groups of 32
	const int n_groups = (n2_ + 31) / 32;
	d_len_.ensure(n2_ + 1); d_hapid_.ensure(n2_ + 1); d_len2_.ensure(n2_ + 1); d_hap_sorted_.ensure(n2_ + 1);
	d_rank_.ensure(n2_ + 1); d_group_len_.ensure(n_groups + 1); d_group_base_.ensure(n_groups + 2);
	incidence_len_kernel<<<(n2_ + 255) / 256, 256, 0, st>>>(d_inc_off_.get(), n2_, d_len_.get(), d_hapid_.get());
	HB_CUDA(cudaGetLastError());
	tmp_bytes = 0;
	HB_CUDA(cub::DeviceRadixSort::SortPairsDescending(nullptr, tmp_bytes, d_len_.get(), d_len2_.get(),
		d_hapid_.get(), d_hap_sorted_.get(), n2_, 0, 32, st));
	d_tmp_.ensure(tmp_bytes + 16);
	HB_CUDA(cub::DeviceRadixSort::SortPairsDescending(d_tmp_.get(), tmp_bytes, d_len_.get(), d_len2_.get(),
		d_hapid_.get(), d_hap_sorted_.get(), n2_, 0, 32, st));
	incidence_groups_kernel<<<1, 1024, 0, st>>>(d_len2_.get(), d_hap_sorted_.get(), n2_, d_rank_.get(),
		d_group_len_.get(), d_group_base_.get());
	HB_CUDA(cudaGetLastError());
	d_pairs4_.ensure(4 * total_pairs_ + 4);
	pairs_pack_kernel<<<blocks, 128, 0, st>>>(d_p1_.get(), d_p2_.get(), d_off_.get(), n_entry,
		d_pairs4_.get());
	HB_CUDA(cudaGetLastError());
	incidence_slots_kernel<<<(n_inc + 255) / 256, 256, 0, st>>>(d_key2_.get(), d_val2_.get(), n_inc,
		d_inc_off_.get(), d_rank_.get(), d_group_base_.get(), d_pairs4_.get());
	HB_CUDA(cudaGetLastError());
	// entry chains of em_chain_kernel's E step: entries by decreasing pair count, groups of 32, the same
	// interleaved layout (the per-candidate compatible pairs fill a prefix of every column)
	const int n_egroups = (n_entry + 31) / 32;
	d_eid_.ensure(n_entry + 1); d_ecnt_sorted_.ensure(n_entry + 1); d_entry_sorted_.ensure(n_entry + 1);
	d_erank_.ensure(n_entry + 1); d_egroup_len_.ensure(n_egroups + 1); d_egroup_base_.ensure(n_egroups + 2);
	iota_kernel<<<(n_entry + 255) / 256, 256, 0, st>>>(n_entry, d_eid_.get());
	HB_CUDA(cudaGetLastError());
	tmp_bytes = 0;
	HB_CUDA(cub::DeviceRadixSort::SortPairsDescending(nullptr, tmp_bytes, d_cnt_.get(), d_ecnt_sorted_.get(),
		d_eid_.get(), d_entry_sorted_.get(), n_entry, 0, 32, st));
	d_tmp_.ensure(tmp_bytes + 16);
	HB_CUDA(cub::DeviceRadixSort::SortPairsDescending(d_tmp_.get(), tmp_bytes, d_cnt_.get(), d_ecnt_sorted_.get(),
		d_eid_.get(), d_entry_sorted_.get(), n_entry, 0, 32, st));
	incidence_groups_kernel<<<1, 1024, 0, st>>>(d_ecnt_sorted_.get(), d_entry_sorted_.get(), n_entry, d_erank_.get(),
		d_egroup_len_.get(), d_egroup_base_.get());
	HB_CUDA(cudaGetLastError());
	// the records of every pair and of every contribution of the round (em_chain_kernel filters them
	// per candidate); padded to whole 8-record steps
	if (n2_ <= EMC_MAX_HAP && n_entry <= EMC_MAX_ENTRY)
	{
		d_erec_all_.ensure(total_pairs_ + 16); d_eown_.ensure(total_pairs_ + 16);
		d_mrec_all_.ensure((size_t)n_inc + 16); d_mown_.ensure((size_t)n_inc + 16);
		emc_pair_records_kernel<<<blocks, 128, 0, st>>>(d_p1_.get(), d_p2_.get(), d_off_.get(), n_entry,
			d_rank_.get(), d_erank_.get(), (uint32_t *)d_erec_all_.get(), d_eown_.get());
		HB_CUDA(cudaGetLastError());
		emc_contrib_records_kernel<<<(n_inc + 255) / 256, 256, 0, st>>>(d_key2_.get(), d_val2_.get(), n_inc,
			d_pairs4_.get(), d_rank_.get(), d_erank_.get(), (uint32_t *)d_mrec_all_.get(), d_mown_.get());
		HB_CUDA(cudaGetLastError());
		launches += 2;
	}
	HB_CUDA(cudaMemcpyAsync(h_total_.get() + 1, d_group_base_.get() + n_groups, sizeof(int),
		cudaMemcpyDeviceToHost, st));
	HB_CUDA(cudaMemcpyAsync(h_total_.get() + 2, d_group_len_.get(), sizeof(int), cudaMemcpyDeviceToHost, st));
	HB_CUDA(cudaMemcpyAsync(h_total_.get() + 4, d_egroup_base_.get() + n_egroups, sizeof(int),
		cudaMemcpyDeviceToHost, st));
	HB_CUDA(cudaMemcpyAsync(h_total_.get() + 5, d_egroup_len_.get(), sizeof(int), cudaMemcpyDeviceToHost, st));
	stream_sync_blocking(st);
	n_slots_ = (size_t)h_total_.get()[1];
	max_chain_ = h_total_.get()[2];
	n_eslots_ = (size_t)h_total_.get()[4];
	max_entry_pairs_ = h_total_.get()[5];
	if (getenv("HIBAG_B200_EM_DEBUG"))
	{
		std::vector<int> c(n_entry);
		HB_CUDA(cudaMemcpy(c.data(), d_cnt_.get(), sizeof(int) * (size_t)n_entry, cudaMemcpyDeviceToHost));
		std::sort(c.begin(), c.end());
		fprintf(stderr, "pairs/entry: median %d p90 %d p99 %d max %d\n", c[n_entry / 2],
			c[(size_t)n_entry * 9 / 10], c[(size_t)n_entry * 99 / 100], c[n_entry - 1]);
	}
	d2h_bytes += 4 * sizeof(int);
	launches += 11;
}

void RoundEM::run_em(const int *cand_snp, int m, const int8_t *geno_t, int n_samp, cudaStream_t st)
{
	if (m <= 0) return;
	int *hc = h_cand_.ensure(m);
	for (int i = 0; i < m; i++) hc[i] = cand_snp[i];
	d_cand_.ensure(m);
	HB_CUDA(cudaMemcpyAsync(d_cand_.get(), hc, sizeof(int) * (size_t)m, cudaMemcpyHostToDevice, st));
	const int n_groups = (n2_ + 31) / 32;
	d_freq_.ensure((size_t)m * n2_ + 2);
	d_status_.ensure(4 * (size_t)m);
	h_freq_.ensure((size_t)m * n2_ + 2);
	h_status_.ensure(4 * (size_t)m);
	// ---- em_chain_kernel (128 threads, state in shared memory, records streamed) whenever its working
	// set fits; HIBAG_B200_EM_CHAIN=0 selects em_kernel (the streaming form, also the one for rounds
	// too large for an SM) ------------------------------------------------------------------------------
	{
		int want_chain = 1;                    // read per call: the tests switch kernels inside one process
		if (const char *e = getenv("HIBAG_B200_EM_CHAIN")) want_chain = atoi(e);
		const int n_egroups = (n_entry_ + 31) / 32;
		// chain lengths during the set-up, the warps' batch schedules afterwards (at most one 16-bit word per
		// 128 record slots)
		const size_t len_words = (std::max((((size_t)n2_ + 1) & ~(size_t)1) + (((size_t)n_entry_ + 1) & ~(size_t)1),
			n_slots_ / 128 + n_eslots_ / 128) + 1) & ~(size_t)1;
		const size_t smem = sizeof(double) * (2 * (((size_t)n2_ + 2) & ~(size_t)1) + 40 + (((size_t)n_entry_ + 2) & ~(size_t)1)) +
			sizeof(uint32_t) * (size_t)EMC_WARPS * EMC_RING_B * 128 +
			sizeof(int) * 2 * ((size_t)n_groups + (size_t)n_egroups) +
			sizeof(uint16_t) * ((((size_t)n_groups + 1) & ~(size_t)1) + (((size_t)n_egroups + 1) & ~(size_t)1) + len_words) +
			(size_t)n_entry_ + 16;
		const size_t budget = (size_t)227 * 1024 - 512;     // the kernel also has a few bytes of static shared memory
		// (records hold haplotype ranks in 14 bits and entry ranks in 15; chain lengths in 16 bits)
		if (want_chain && n_entry_ <= EMC_MAX_ENTRY && n2_ <= EMC_MAX_HAP && max_chain_ <= 65535 &&
			max_entry_pairs_ <= 65535 && (n_slots_ + n_eslots_) / 128 <= 65535 && smem <= budget)
		{
			d_idxell_.ensure((size_t)m * n_slots_ + 4);
			d_erec_.ensure((size_t)m * n_eslots_ + 4);
			EmcArgs a;
			memset(&a, 0, sizeof(a));
			a.n_entry = n_entry_; a.n_cur = n_cur_; a.n_samp = n_samp; a.total_pairs = (int)total_pairs_;
			a.ib = ib_; a.boot = boot_;
			a.hap_sorted = d_hap_sorted_.get(); a.hap_rank = d_rank_.get(); a.group_base = d_group_base_.get();
			a.mrec_all = (const uint32_t *)d_mrec_all_.get(); a.mown = d_mown_.get(); a.n_slots = n_slots_;
			a.entry_sorted = d_entry_sorted_.get(); a.egroup_base = d_egroup_base_.get();
			a.erec_all = (const uint32_t *)d_erec_all_.get(); a.eown = d_eown_.get(); a.n_eslots = n_eslots_;
			a.cur_freq = d_curfreq_.get();
			a.geno_t = geno_t; a.cand_snp = d_cand_.get();
			a.rec = (uint32_t *)d_idxell_.get(); a.erec = (uint32_t *)d_erec_.get();
			a.out_freq = d_freq_.get(); a.out_status = d_status_.get();
			a.scale = 0.5 / n_samp; a.em_reltol = std::sqrt(DBL_EPSILON);
			a.acct = device_sm_acct();
			a.len_words = (int)len_words;
			const bool want_prof_r = getenv("HIBAG_B200_EM_DEBUG") != nullptr;
			if (want_prof_r)
			{
				d_prof_.ensure(8 * (size_t)m);
				HB_CUDA(cudaMemsetAsync(d_prof_.get(), 0, sizeof(unsigned long long) * 8 * (size_t)m, st));
				a.prof = d_prof_.get();
			}
			// the share of an SM a CTA holds: by shared memory (1 KB reserved per CTA) or by threads
			int per_sm = (int)((size_t)228 * 1024 / (smem + 1024));
			per_sm = std::max(1, std::min(per_sm, 2048 / EMC_THREADS));
			a.acct_w = 1024 / per_sm;
			const int max_dyn = 227 * 1024 - 256;
			g_em_gate.enter();
			auto kern = em_chain_kernel;
			cudaError_t launch_rc = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, max_dyn);
			if (launch_rc == cudaSuccess) launch_rc = cudaEventRecord(ev0_.e, st);
			if (launch_rc == cudaSuccess) { kern<<<m, EMC_THREADS, smem, st>>>(a); launch_rc = cudaGetLastError(); }
			if (launch_rc != cudaSuccess) { g_em_gate.leave(); HB_CUDA(launch_rc); }
			HB_CUDA(cudaEventRecord(ev1_.e, st));
			HB_CUDA(cudaMemcpyAsync(h_freq_.get(), d_freq_.get(), sizeof(double) * (size_t)m * n2_,
				cudaMemcpyDeviceToHost, st));
			HB_CUDA(cudaMemcpyAsync(h_status_.get(), d_status_.get(), sizeof(int) * 4 * (size_t)m,
				cudaMemcpyDeviceToHost, st));
			HB_CUDA(cudaEventRecord(ev_done_.e, st));
			const cudaError_t sync_rc = cudaEventSynchronize(ev_done_.e);
			g_em_gate.leave();
			HB_CUDA(sync_rc);
			float ms = 0;
			HB_CUDA(cudaEventElapsedTime(&ms, ev0_.e, ev1_.e));
			kernel_ms += ms;
			launches++;
			chain_launches++;
			for (int i = 0; i < m; i++)
			{
				const int *stt = h_status_.get() + 4 * i;
				if (stt[0] == EM_INVALID || stt[2] < 0) continue;
				sum_iterations += (uint64_t)stt[1];
				sum_chain_adds += (uint64_t)stt[1] * (uint64_t)stt[2];
				sum_pair_updates += (uint64_t)stt[1] * (uint64_t)stt[3];
			}
			if (want_prof_r)
			{
				long compat = 0; int nv = 0;
				for (int i = 0; i < m; i++)
				{
					const int *stt = h_status_.get() + 4 * i;
					if (stt[0] != EM_INVALID && stt[2] >= 0) { compat += stt[3]; nv++; }
				}
				std::vector<unsigned long long> hp(8 * (size_t)m);
				HB_CUDA(cudaMemcpy(hp.data(), d_prof_.get(), sizeof(unsigned long long) * hp.size(), cudaMemcpyDeviceToHost));
				double su = 0, se = 0, sm = 0, it = 0; int it_max = 0, le = 0;
				for (int i = 0; i < m; i++)
				{
					su += (double)hp[8 * i]; se += (double)hp[8 * i + 1]; sm += (double)hp[8 * i + 2]; it += (double)hp[8 * i + 3];
					it_max = std::max(it_max, (int)hp[8 * i + 3]); le = std::max(le, (int)hp[8 * i + 4]);
				}
				fprintf(stderr, "emc groups (candidate 0): E %d groups, %llu batches, %llu long; M %d groups, %llu batches, %llu long; "
					"E pass 1 %.1f kcycles per iteration\n",
					n_egroups, hp[5] & 0xffffffffull, hp[5] >> 32, n_groups, hp[6] & 0xffffffffull, hp[6] >> 32,
					hp[3] ? (double)hp[7] / (double)hp[3] * 1e-3 : 0.0);
				fprintf(stderr, "em chain: kernel %.3f ms, %d candidates, total pairs %zu, compat/cand %.0f, smem %zu (%d per SM) | "
					"kcycles: set-up %.0f per candidate, E step %.1f and M step %.1f per iteration, iterations mean %.1f max %d, "
					"longest chain %d, longest entry %d\n", ms, m, total_pairs_, nv ? (double)compat / nv : 0.0, smem, per_sm,
					su / m * 1e-3, it > 0 ? se / it * 1e-3 : 0.0, it > 0 ? sm / it * 1e-3 : 0.0, it / m, it_max,
					h_status_.get()[2], le);
			}
			h2d_bytes += sizeof(int) * (size_t)m;
			d2h_bytes += sizeof(double) * (size_t)m * n2_ + sizeof(int) * 4 * (size_t)m;
			return;
		}
	}
	d_pmap_.ensure(4 * (size_t)m * total_pairs_ + 4);
	d_rinc_.ensure((size_t)m * n_slots_ + 2);
	// slots without a compatible contribution stay 0.0
	HB_CUDA(cudaMemsetAsync(d_rinc_.get(), 0, sizeof(double) * (size_t)m * n_slots_, st));
	d_cuv_.ensure((size_t)m * total_pairs_ + 1);
	d_xbuf_.ensure((size_t)m * total_pairs_ + 2);
	d_coff_.ensure((size_t)m * (n_entry_ + 1));
	d_glen_.ensure((size_t)m * n_groups + 1);
	HB_CUDA(cudaMemsetAsync(d_glen_.get(), 0, sizeof(int) * (size_t)m * n_groups, st));
	d_freq_.ensure((size_t)m * n2_ + 2);
	d_status_.ensure(4 * (size_t)m);
	h_freq_.ensure((size_t)m * n2_ + 2);
	h_status_.ensure(4 * (size_t)m);
	EmArgs a;
	memset(&a, 0, sizeof(a));
	a.n_entry = n_entry_; a.n_cur = n_cur_; a.n_samp = n_samp;
	a.ib = ib_; a.boot = boot_;
	a.off = d_off_.get(); a.pairs4 = (const int4 *)d_pairs4_.get();
	a.hap_sorted = d_hap_sorted_.get(); a.group_base = d_group_base_.get();
	a.inc_off = d_inc_off_.get(); a.inc_val = d_val2_.get();
	a.cur_freq = d_curfreq_.get();
	a.n_slots = n_slots_;
	a.geno_t = geno_t; a.cand_snp = d_cand_.get();
	a.total_pairs = (int)total_pairs_;
	a.pmap = d_pmap_.get(); a.rinc = d_rinc_.get(); a.cuv = d_cuv_.get(); a.xbuf = d_xbuf_.get();
	a.coff = d_coff_.get(); a.glen = d_glen_.get();
	a.out_freq = d_freq_.get(); a.out_status = d_status_.get();
	a.scale = 0.5 / n_samp;
	a.em_reltol = std::sqrt(DBL_EPSILON);
	static const bool want_prof = getenv("HIBAG_B200_EM_PROF") != nullptr;
	if (want_prof)
	{
		d_prof_.ensure(8 * (size_t)m);
		HB_CUDA(cudaMemsetAsync(d_prof_.get(), 0, sizeof(unsigned long long) * 8 * (size_t)m, st));
		a.prof = d_prof_.get();
	}
	const size_t smem_base = sizeof(double) * (2 * (size_t)n2_ + 40 + 3 * MAX_CLUSTER + (((size_t)n_entry_ + 1) & ~(size_t)1)) +
		sizeof(int) * (size_t)n_entry_ + 16;
	// "dense" shape: 512 threads, 32-row rings and at most half of an SM's shared memory, so that two
	// EM CTAs (of any lanes) share an SM -- the kernel is latency-bound (~20 % issue utilisation alone)
	bool dense = n_dense_lanes_ > 1;
	if (const char *e = getenv("HIBAG_B200_EM_DENSE")) dense = atoi(e) != 0;
	size_t budget = dense ? (size_t)112 * 1024 : (size_t)220 * 1024;
	if (dense && smem_base + sizeof(double) * 32 * 32 > budget) { dense = false; budget = (size_t)220 * 1024; }
	int ring_rows = dense ? 32 : 64;
	if (const char *e = getenv("HIBAG_B200_EM_RING_ROWS")) ring_rows = (atoi(e) >= 64) ? 64 : 32;
	if (dense && smem_base + sizeof(double) * 32 * (size_t)ring_rows > budget) ring_rows = 32;
	if (!dense) ring_rows = 64;
	const size_t ring_b = sizeof(double) * 32 * (size_t)ring_rows;
	int m_warps = (smem_base < budget) ? (int)((budget - smem_base) / ring_b) : 0;
	int em_threads = dense ? 512 : 1024, max_rings = 8;
	if (const char *e = getenv("HIBAG_B200_EM_RINGS")) max_rings = std::max(1, std::min(8, atoi(e)));
	if (m_warps > max_rings) m_warps = max_rings;
	if (m_warps < 1) throw std::runtime_error("run_em: list too large for the device EM");
	a.m_warps = m_warps;
	size_t smem = smem_base + ring_b * (size_t)m_warps;
	// HIBAG_B200_EM_SMEM_KB: pad the request so that fewer EM CTAs fit an SM (116: one per SM) and the
	// scoring CTAs of other lanes always find room beside them
	static const int pad_kb = []() { const char *e = getenv("HIBAG_B200_EM_SMEM_KB"); return e ? atoi(e) : 0; }();
	if (pad_kb > 0 && smem < (size_t)pad_kb * 1024) smem = std::min((size_t)pad_kb * 1024, (size_t)220 * 1024);
	// SMs per candidate: enough pairs per CTA to pay for the cluster barriers, and the whole
	// round on at most ~half of the SMs (the other lanes' scoring launches run beside it; the
	// kernel is latency-bound, so SM-time per candidate is lowest for small clusters)
	int cluster = 1;
	while (cluster < MAX_CLUSTER && total_pairs_ / (2 * (size_t)cluster) >= 6000 &&
		(size_t)m * 2 * cluster <= (size_t)current_device().sm_count * 11 / 20) cluster *= 2;
	// many classifiers in flight: the GPU is shared by dozens of EM launches and the SM-time per
	// candidate is what counts -- one CTA per candidate (measured, config 2, 24 lanes: 783 vs 678
	// classifiers/min for clusters of 2)
	if (dense && n_dense_lanes_ >= 8) cluster = 1;
	if (const char *e = getenv("HIBAG_B200_EM_CLUSTER")) cluster = std::max(1, std::min(MAX_CLUSTER, atoi(e)));
	auto kern = dense ? (ring_rows == 64 ? em_kernel<512, 64> : em_kernel<512, 32>) : em_kernel<1024, 64>;
	a.acct = device_sm_acct();
	a.acct_w = (dense && smem <= (size_t)113 * 1024) ? 512 : 1024;
	HB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(220 * 1024)));
	g_em_gate.enter();
	HB_CUDA(cudaEventRecord(ev0_.e, st));
	{
		cudaLaunchConfig_t cfg;
		memset(&cfg, 0, sizeof(cfg));
		cfg.gridDim = dim3((unsigned)(m * cluster)); cfg.blockDim = dim3((unsigned)em_threads);
		cfg.dynamicSmemBytes = smem; cfg.stream = st;
		cudaLaunchAttribute attr[1];
		attr[0].id = cudaLaunchAttributeClusterDimension;
		attr[0].val.clusterDim.x = (unsigned)cluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
		cfg.attrs = attr; cfg.numAttrs = 1;
		HB_CUDA(cudaLaunchKernelEx(&cfg, kern, a));
	}
	HB_CUDA(cudaGetLastError());
	HB_CUDA(cudaEventRecord(ev1_.e, st));
	HB_CUDA(cudaMemcpyAsync(h_freq_.get(), d_freq_.get(), sizeof(double) * (size_t)m * n2_,
		cudaMemcpyDeviceToHost, st));
	HB_CUDA(cudaMemcpyAsync(h_status_.get(), d_status_.get(), sizeof(int) * 4 * (size_t)m,
		cudaMemcpyDeviceToHost, st));
	HB_CUDA(cudaEventRecord(ev_done_.e, st));
	const cudaError_t sync_rc = cudaEventSynchronize(ev_done_.e);
	g_em_gate.leave();
	HB_CUDA(sync_rc);
	float ms = 0;
	HB_CUDA(cudaEventElapsedTime(&ms, ev0_.e, ev1_.e));
	kernel_ms += ms;
	launches++;
	for (int i = 0; i < m; i++)
	{
		const int *stt = h_status_.get() + 4 * i;
		if (stt[0] == EM_INVALID) continue;
		sum_iterations += (uint64_t)stt[1];
		sum_chain_adds += (uint64_t)stt[1] * (uint64_t)stt[2];
		sum_pair_updates += (uint64_t)stt[1] * (uint64_t)stt[3];
	}
	if (want_prof)
	{
		std::vector<unsigned long long> hp(8 * (size_t)m);
		HB_CUDA(cudaMemcpy(hp.data(), d_prof_.get(), sizeof(unsigned long long) * hp.size(), cudaMemcpyDeviceToHost));
		std::lock_guard<std::mutex> lk(g_em_prof.mu);
		for (int i = 0; i < m; i++)
		{
			if (h_status_.get()[4 * i] == EM_INVALID) continue;
			for (int q = 0; q < 7; q++) g_em_prof.t[q] += hp[8 * (size_t)i + q];
			g_em_prof.cands++; g_em_prof.slots += n_slots_; g_em_prof.pairs += total_pairs_;
		}
		g_em_prof.launches++;
		if (!g_em_prof.registered) { g_em_prof.registered = true; atexit(EmProf::dump); }
	}
	if (getenv("HIBAG_B200_EM_DEBUG"))
	{
		std::vector<int> gl((size_t)m * n_groups), co((size_t)m * (n_entry_ + 1));
		HB_CUDA(cudaMemcpy(gl.data(), d_glen_.get(), sizeof(int) * gl.size(), cudaMemcpyDeviceToHost));
		HB_CUDA(cudaMemcpy(co.data(), d_coff_.get(), sizeof(int) * co.size(), cudaMemcpyDeviceToHost));
		long rows = 0; int gmax = 0; long compat = 0;
		for (int i = 0; i < m; i++)
		{
			if (h_status_.get()[4 * i] == EM_INVALID) continue;
			for (int g = 0; g < n_groups; g++) { rows += gl[(size_t)i * n_groups + g]; gmax = std::max(gmax, gl[(size_t)i * n_groups + g]); }
			compat += co[(size_t)i * (n_entry_ + 1) + n_entry_];
		}
		fprintf(stderr, "em compaction: cluster %d, kernel %.3f ms, compat pairs/cand %.0f of %zu, ELL rows/cand %.0f, longest chain %d (uncompacted %d)\n",
			cluster, ms, (double)compat / m, total_pairs_, (double)rows / m, gmax, max_chain_);
	}
	h2d_bytes += sizeof(int) * (size_t)m;
	d2h_bytes += sizeof(double) * (size_t)m * n2_ + sizeof(int) * 4 * (size_t)m;
}

void RoundEM::fetch_pairs(RoundPairs &out, const std::vector<int> &inbag,
	const std::vector<int> &boot, cudaStream_t st)
{
	const int n = n_entry_;
	out.n_cur = n_cur_;
	out.samp.resize(n); out.boot.resize(n); out.off.resize(n + 1);
	std::vector<int> off(n + 1);
	out.p1.resize(total_pairs_); out.p2.resize(total_pairs_);
	HB_CUDA(cudaMemcpyAsync(off.data(), d_off_.get(), sizeof(int) * (size_t)(n + 1), cudaMemcpyDeviceToHost, st));
	if (total_pairs_)
	{
		HB_CUDA(cudaMemcpyAsync(out.p1.data(), d_p1_.get(), sizeof(int) * total_pairs_, cudaMemcpyDeviceToHost, st));
		HB_CUDA(cudaMemcpyAsync(out.p2.data(), d_p2_.get(), sizeof(int) * total_pairs_, cudaMemcpyDeviceToHost, st));
	}
	stream_sync_blocking(st);
	for (int k = 0; k < n; k++)
	{
		out.samp[k] = inbag[k];
		out.boot[k] = boot[inbag[k]];
		out.off[k] = (size_t)off[k];
	}
	out.off[n] = (size_t)off[n];
	d2h_bytes += sizeof(int) * ((size_t)(n + 1) + 2 * total_pairs_);
}

// ---------------------------------------------------------------------------------------------
// build_haplomatch hook body
// ---------------------------------------------------------------------------------------------
uint32_t *haplomatch_records(const hibag_haplotype *haplo, const size_t *n_haplo, int n_hla,
	int n_snp, const hibag_genotype *geno, int n_samp, const std::vector<int> &ib, size_t *out_n)
{
	current_device();
	Stream st;
	std::vector<int> start(n_hla + 1, 0);
	for (int a = 0; a < n_hla; a++)
	{
		if (n_haplo[a] > 65535)
			throw std::runtime_error("There are too many HLA allele-specific haplotypes (# > 65535).");
		start[a + 1] = start[a] + (int)n_haplo[a];
	}
	const int n_cur = start[n_hla];
	const int n_entry = (int)ib.size();
	std::vector<uint64_t> hh(2 * (size_t)n_cur + 2);
	for (int k = 0; k < n_cur; k++)
	{
		hh[2 * k] = (uint64_t)haplo[k].packed[0]; hh[2 * k + 1] = (uint64_t)haplo[k].packed[1];
	}
	DevBuf<uint64_t> d_hap; DevBuf<int> d_start, d_ib, d_cnt, d_off, d_mind, d_a1, d_a2, d_boot;
	DevBuf<uint32_t> d_s1, d_s2, d_rec;
	DevBuf<unsigned char> d_aos, d_tmp;
	d_hap.ensure(hh.size()); d_start.ensure(n_hla + 1); d_ib.ensure(n_entry + 1);
	d_cnt.ensure(n_entry + 1); d_off.ensure(n_entry + 1); d_mind.ensure(n_entry + 1);
	d_a1.ensure(n_samp); d_a2.ensure(n_samp); d_boot.ensure(n_samp);
	d_s1.ensure(4 * (size_t)n_samp); d_s2.ensure(4 * (size_t)n_samp);
	d_aos.ensure(sizeof(hibag_genotype) * (size_t)n_samp);
	HB_CUDA(cudaMemcpyAsync(d_hap.get(), hh.data(), sizeof(uint64_t) * 2 * (size_t)n_cur, cudaMemcpyHostToDevice, st.s));
	HB_CUDA(cudaMemcpyAsync(d_start.get(), start.data(), sizeof(int) * (size_t)(n_hla + 1), cudaMemcpyHostToDevice, st.s));
	HB_CUDA(cudaMemcpyAsync(d_ib.get(), ib.data(), sizeof(int) * (size_t)n_entry, cudaMemcpyHostToDevice, st.s));
	HB_CUDA(cudaMemcpyAsync(d_aos.get(), geno, sizeof(hibag_genotype) * (size_t)n_samp, cudaMemcpyHostToDevice, st.s));
	launch_unpack_genotypes(d_aos.get(), n_samp, d_s1.get(), d_s2.get(), n_samp, d_a1.get(),
		d_a2.get(), d_boot.get(), st.s);
	MatchArgs m;
	memset(&m, 0, sizeof(m));
	m.hap = d_hap.get(); m.start = d_start.get();
	m.s1 = d_s1.get(); m.s2 = d_s2.get(); m.stride = n_samp;
	m.a1 = d_a1.get(); m.a2 = d_a2.get(); m.ib = d_ib.get();
	m.n_entry = n_entry; m.n_snp = n_snp;
	m.cnt = d_cnt.get(); m.mind = d_mind.get(); m.off = d_off.get();
	const int blocks = (n_entry + 127) / 128;
	HB_CUDA(cudaMemsetAsync(d_cnt.get() + n_entry, 0, sizeof(int), st.s));
	if (n_entry > 0)
	{
		haplomatch_kernel<2><<<blocks, 128, 0, st.s>>>(m);
		HB_CUDA(cudaGetLastError());
	}
	size_t tmp_bytes = 0;
	HB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, d_cnt.get(), d_off.get(), n_entry + 1, st.s));
	d_tmp.ensure(tmp_bytes + 16);
	HB_CUDA(cub::DeviceScan::ExclusiveSum(d_tmp.get(), tmp_bytes, d_cnt.get(), d_off.get(), n_entry + 1, st.s));
	int total = 0;
	HB_CUDA(cudaMemcpyAsync(&total, d_off.get() + n_entry, sizeof(int), cudaMemcpyDeviceToHost, st.s));
	HB_CUDA(cudaStreamSynchronize(st.s));
	uint32_t *buf = (uint32_t *)malloc(sizeof(uint32_t) * (1 + 2 * (size_t)total));
	if (!buf) throw std::runtime_error("build_haplomatch: out of memory");
	buf[0] = (uint32_t)(2 * total);
	if (total > 0)
	{
		d_rec.ensure(2 * (size_t)total);
		m.rec = d_rec.get();
		haplomatch_kernel<3><<<blocks, 128, 0, st.s>>>(m);
		HB_CUDA(cudaGetLastError());
		HB_CUDA(cudaMemcpyAsync(buf + 1, d_rec.get(), sizeof(uint32_t) * 2 * (size_t)total,
			cudaMemcpyDeviceToHost, st.s));
		HB_CUDA(cudaStreamSynchronize(st.s));
	}
	if (out_n) *out_n = 1 + 2 * (size_t)total;
	return buf;
}

}  // namespace hb
