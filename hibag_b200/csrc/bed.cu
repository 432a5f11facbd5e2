// bed.cu -- PLINK BED decoding on the device: the 2-bit packed genotypes
// of a .bed file become the int8 [n_samp][n_snp] matrix hibag_b200_model_set_training /
// hibag_b200_model_predict take. Reference: HIBAG_ConvBED, src/HIBAG.cpp:1094-1191 (code table
// {2, NA, 1, 0} :1141; NA is written as -1 here, the pipeline's "anything outside 0..2").
//
// HBM-bound byte work: 0.25 B read + 1 B written per genotype. SNP-major files (the PLINK
// default) need a transpose: a thread owns one selected SNP and 128 samples of it (32 packed
// bytes = one sector of its row), and the 128 threads of a CTA write 128 consecutive int8 of one
// sample row per step, so the stores -- 80 % of the traffic -- are fully coalesced.

#include "kernels.h"

#include <stdexcept>
#include <string>

namespace hb {

#define CUDA_CHECK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) \
	throw std::runtime_error(std::string("CUDA error: ") + cudaGetErrorString(e_) + \
		" at " __FILE__ ":" + std::to_string(__LINE__)); } while (0)

__device__ __forceinline__ int8_t bed_code(unsigned int g)
{
	// 00 -> 2, 01 -> missing, 10 -> 1, 11 -> 0    (src/HIBAG.cpp:1141)
	return (int8_t)((0x00'01'FF'02u >> (8u * (g & 3u))) & 0xffu);
}

/// SNP-major payload [n_snp][bps], bps = ceil(n_samp / 4)
__global__ void __launch_bounds__(128)
bed_snp_major_kernel(const uint8_t *__restrict__ payload, size_t bps, int n_samp,
	const int32_t *__restrict__ sel, int n_save, int8_t *__restrict__ out)
{
	const int k = blockIdx.x * 128 + threadIdx.x;         // selected SNP
	const int j0 = blockIdx.y * 128;                        // first sample of the tile
	const bool ok = k < n_save;
	const int n_here = min(128, n_samp - j0);
	const int n_b = (n_here + 3) >> 2;
	uint32_t w[8];
#pragma unroll
	for (int q = 0; q < 8; q++) w[q] = 0;
	if (ok)
	{
		const uint8_t *row = payload + (size_t)(sel ? __ldg(sel + k) : k) * bps + (size_t)(j0 >> 2);
#pragma unroll
		for (int b = 0; b < 32; b++)
			if (b < n_b) w[b >> 2] |= (uint32_t)__ldg(row + b) << (8 * (b & 3));
	}
	int8_t *o = out + (size_t)j0 * n_save + k;
#pragma unroll
	for (int q = 0; q < 8; q++)
	{
#pragma unroll
		for (int s = 0; s < 16; s++)
		{
			const int j = q * 16 + s;
			if (ok && j < n_here) o[(size_t)j * n_save] = bed_code(w[q] >> (2 * s));
		}
	}
}

/// individual-major payload [n_samp][bps], bps = ceil(n_snp / 4)
__global__ void __launch_bounds__(256)
bed_ind_major_kernel(const uint8_t *__restrict__ payload, size_t bps, int n_samp,
	const int32_t *__restrict__ sel, int n_save, int8_t *__restrict__ out)
{
	const int k = blockIdx.x * 256 + threadIdx.x;
	const int j = blockIdx.y;
	if (k >= n_save || j >= n_samp) return;
	const int snp = sel ? __ldg(sel + k) : k;
	const unsigned int g = __ldg(payload + (size_t)j * bps + (size_t)(snp >> 2)) >> (2 * (snp & 3));
	out[(size_t)j * n_save + k] = bed_code(g);
}

void launch_bed_decode(const uint8_t *payload, int mode, int n_samp, int n_snp, const int32_t *sel,
	int n_save, int8_t *out, cudaStream_t st)
{
	if (n_samp <= 0 || n_save <= 0) return;
	if (mode == 0)
	{
		const size_t bps = ((size_t)n_snp + 3) / 4;
		dim3 grid((n_save + 255) / 256, n_samp);
		bed_ind_major_kernel<<<grid, 256, 0, st>>>(payload, bps, n_samp, sel, n_save, out);
	} else {
		const size_t bps = ((size_t)n_samp + 3) / 4;
		dim3 grid((n_save + 127) / 128, (n_samp + 127) / 128);
		bed_snp_major_kernel<<<grid, 128, 0, st>>>(payload, bps, n_samp, sel, n_save, out);
	}
	CUDA_CHECK(cudaGetLastError());
}

}  // namespace hb
