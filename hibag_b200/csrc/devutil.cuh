// devutil.cuh -- small PTX helpers and the haplotype record fetch shared by the scoring kernels
// (kernels.cu, screen.cu) (internal)
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace hb {

// ---------------------------------------------------------------------------------------
// SM-time accounting: every CTA adds (cycles it was resident) x (its share of an SM in 1/1024) to a
// per-class counter. Overlapped CUDA-event durations cannot attribute time between concurrently
// running kernels; resident SM-time can. Classes: see SM_ACCT_* in kernels.h.
// ---------------------------------------------------------------------------------------
struct SmAcct
{
	unsigned long long *slot;
	long long t0;
	unsigned int w;
	__device__ __forceinline__ SmAcct(unsigned long long *acct, int cls, unsigned int weight) : slot(nullptr), t0(0), w(weight)
	{
		if (acct != nullptr && threadIdx.x == 0 && threadIdx.y == 0) { slot = acct + cls; t0 = clock64(); }
	}
	__device__ __forceinline__ ~SmAcct()
	{
		if (slot != nullptr) atomicAdd(slot, (unsigned long long)(clock64() - t0) * (unsigned long long)w);
	}
};

// ---------------------------------------------------------------------------------------
// small PTX helpers
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
	return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
		:: "r"(bar), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
	uint32_t done = 0;
	while (!done)
	{
		asm volatile(
			"{\n\t.reg .pred p;\n\t"
			"mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
			"selp.u32 %0, 1, 0, p;\n\t}"
			: "=r"(done) : "r"(bar), "r"(parity) : "memory");
	}
}

/// 1-D TMA bulk copy global -> shared (SASS: UBLKCP), completion counted on an mbarrier
__device__ __forceinline__ void tma_bulk_g2s(uint32_t dst, const void *src, uint32_t bytes,
	uint32_t bar)
{
	asm volatile(
		"cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
		:: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

__device__ __forceinline__ double lds_f64(uint32_t addr)
{
	double v;
	asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
	return v;
}

__device__ __forceinline__ uint4 lds_v4(uint32_t addr)
{
	uint4 v;
	asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
		: "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
	return v;
}

// ---------------------------------------------------------------------------------------
// the pair-scoring kernel
// ---------------------------------------------------------------------------------------

/// haplotype record fetch: NW 32-bit allele words + fp64 frequency.
/// layout: NW<=2 -> 16 B {w0, w1, f.lo, f.hi};  NW==4 -> 32 B {w0..w3, f.lo, f.hi, 0, 0}
template <int NW, bool SMEM>
struct HapRec
{
	uint32_t h[NW];
	double f;
	__device__ __forceinline__ void load(uint32_t smem_base, const char *gbase, int idx)
	{
		if (SMEM)
		{
			if (NW <= 2)
			{
				uint4 v = lds_v4(smem_base + idx * 16);
				h[0] = v.x; if (NW == 2) h[1] = v.y;
				f = __hiloint2double(v.w, v.z);
			} else {
				uint4 v = lds_v4(smem_base + idx * 32);
				uint4 u = lds_v4(smem_base + idx * 32 + 16);
				h[0] = v.x; h[1 % NW] = v.y; h[2 % NW] = v.z; h[3 % NW] = v.w;
				f = __hiloint2double(u.y, u.x);
			}
		} else {
			if (NW <= 2)
			{
				uint4 v = __ldg((const uint4 *)(gbase + (size_t)idx * 16));
				h[0] = v.x; if (NW == 2) h[1] = v.y;
				f = __hiloint2double(v.w, v.z);
			} else {
				uint4 v = __ldg((const uint4 *)(gbase + (size_t)idx * 32));
				uint4 u = __ldg((const uint4 *)(gbase + (size_t)idx * 32 + 16));
				h[0] = v.x; h[1 % NW] = v.y; h[2 % NW] = v.z; h[3 % NW] = v.w;
				f = __hiloint2double(u.y, u.x);
			}
		}
	}
};

/// shared address of T[c_i + pc][lane] from the row address of T[c_i][lane]: one IMAD. (Written as
/// plain C the compiler re-associates it into (c_i + pc) * 256 + base: an extra IADD per pair.)
__device__ __forceinline__ uint32_t table_row(uint32_t row_ci, int pc)
{
	uint32_t a;
	asm("mad.lo.u32 %0, %1, 256, %2;" : "=r"(a) : "r"((uint32_t)pc), "r"(row_ci));
	return a;
}

/// The inner loop of the reference's chain (src/LibHLA.cpp:1660-1668 and its _PostProb / _PostProb2
/// twins): partners j0 .. b_n-1 of row i for the lane's R samples, sum[r] += ((2 f_i) f_j) * T[d].
/// R >= 3: R independent chains per lane keep the pipes busy; the plain loop (6 instructions per pair
/// evaluation). R <= 2: the loop is bound by the latency of a whole step (record LDS.128 -> POPC ->
/// table LDS.64 -> DMUL -> DADD, ~85 cycles: the shared-memory loads are volatile asm and keep their
/// order), not by the dependent DADD. There the terms of the next 4 / R partners are fetched and
/// multiplied BEFORE the adds of the current ones are issued, so only the adds stay on the critical
/// path. Same operands, same order, un-fused.
template <int NW, int R, bool CLAMP, bool SMEM>
__device__ __forceinline__ void partner_loop(uint32_t hap_base, const char *hap_g, int b_start, int j0,
	int b_n, double ff, const uint32_t (&K)[R][NW], const uint32_t (&V)[R][NW], const int (&ci)[R],
	const uint32_t (&tb)[R], uint32_t tbl_lane, int dmax, double *sum)
{
	if (R <= 2)
	{
		constexpr int JB = (R <= 2) ? 4 / R : 1;
		auto terms = [&](int j, double (&x)[JB][R])
		{
			HapRec<NW, SMEM> hj[JB];
#pragma unroll
			for (int q = 0; q < JB; q++) hj[q].load(hap_base, hap_g, b_start + j + q);
			uint32_t ad[JB][R];
#pragma unroll
			for (int q = 0; q < JB; q++)
#pragma unroll
				for (int r = 0; r < R; r++)
				{
					int pc = 0;
#pragma unroll
					for (int w = 0; w < NW; w++) pc += __popc((hj[q].h[w] ^ K[r][w]) & V[r][w]);
					ad[q][r] = CLAMP ? (tbl_lane + (uint32_t)min(ci[r] + pc, dmax) * 256u) : table_row(tb[r], pc);
				}
#pragma unroll
			for (int q = 0; q < JB; q++)
#pragma unroll
				for (int r = 0; r < R; r++) x[q][r] = lds_f64(ad[q][r]);
#pragma unroll
			for (int q = 0; q < JB; q++)
			{
				const double pf = __dmul_rn(ff, hj[q].f);
#pragma unroll
				for (int r = 0; r < R; r++) x[q][r] = __dmul_rn(pf, x[q][r]);
			}
		};
		int j = j0;
		if (j + JB <= b_n)
		{
			double x0[JB][R];
			terms(j, x0);
			for (j += JB; j + JB <= b_n; j += JB)
			{
				double x1[JB][R];
				terms(j, x1);
#pragma unroll
				for (int q = 0; q < JB; q++)
#pragma unroll
					for (int r = 0; r < R; r++) { sum[r] = __dadd_rn(sum[r], x0[q][r]); x0[q][r] = x1[q][r]; }
			}
#pragma unroll
			for (int q = 0; q < JB; q++)
#pragma unroll
				for (int r = 0; r < R; r++) sum[r] = __dadd_rn(sum[r], x0[q][r]);
		}
		for (; j < b_n; j++)
		{
			HapRec<NW, SMEM> hj;
			hj.load(hap_base, hap_g, b_start + j);
			const double pf = __dmul_rn(ff, hj.f);
#pragma unroll
			for (int r = 0; r < R; r++)
			{
				int pc = 0;
#pragma unroll
				for (int w = 0; w < NW; w++) pc += __popc((hj.h[w] ^ K[r][w]) & V[r][w]);
				double t;
				if (CLAMP) t = lds_f64(tbl_lane + (uint32_t)min(ci[r] + pc, dmax) * 256u);
				else t = lds_f64(table_row(tb[r], pc));
				sum[r] = __dadd_rn(sum[r], __dmul_rn(pf, t));
			}
		}
	} else {
#pragma unroll 2
		for (int j = j0; j < b_n; j++)
		{
			HapRec<NW, SMEM> hj;
			hj.load(hap_base, hap_g, b_start + j);
			const double pf = __dmul_rn(ff, hj.f);
#pragma unroll
			for (int r = 0; r < R; r++)
			{
				int pc = 0;
#pragma unroll
				for (int w = 0; w < NW; w++) pc += __popc((hj.h[w] ^ K[r][w]) & V[r][w]);
				double t;
				if (CLAMP) t = lds_f64(tbl_lane + (uint32_t)min(ci[r] + pc, dmax) * 256u);
				else t = lds_f64(table_row(tb[r], pc));
				sum[r] = __dadd_rn(sum[r], __dmul_rn(pf, t));
			}
		}
	}
}

}  // namespace hb
