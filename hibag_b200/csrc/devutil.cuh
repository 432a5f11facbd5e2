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

}  // namespace hb
