// bed.cu -- PLINK BED decoding on the device: the 2-bit packed genotypes
// of a .bed file become the int8 [n_samp][n_snp] matrix hibag_b200_model_set_training /
// hibag_b200_model_predict take. Reference: HIBAG_ConvBED, src/HIBAG.cpp:1094-1191 (code table
// {2, NA, 1, 0} :1141; NA is written as -1 here, the pipeline's "anything outside 0..2").
//
// HBM-bound byte work: 0.25 B read + 1 B written per genotype. SNP-major files (the PLINK
// default) need a transpose: a thread owns 4 consecutive selected SNPs and 128 samples of them (32
// packed bytes = one sector of each row) and writes one 32-bit word per sample, so a warp stores
// 128 contiguous bytes of a sample row per instruction -- the stores are 80 % of the traffic.

#include "kernels.h"

#include <cstdint>
#include <stdexcept>
#include <string>

namespace hb {

#define CUDA_CHECK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) \
	throw std::runtime_error(std::string("CUDA error: ") + cudaGetErrorString(e_) + \
		" at " __FILE__ ":" + std::to_string(__LINE__)); } while (0)

// code table: 00 -> 2, 01 -> missing (-1), 10 -> 1, 11 -> 0    (src/HIBAG.cpp:1141)

/// 32 packed bytes of a row starting at any byte address, as 8 little-endian words: aligned 32-bit
/// loads + funnel shifts (the row pitch of a .bed file is arbitrary). Only words that hold at least
/// one of the n_b wanted bytes are touched.
__device__ __forceinline__ void load_row32(const uint8_t *row, int n_b, uint32_t (&w)[8])
{
	const uintptr_t a = (uintptr_t)row;
	if ((a & 15) == 0 && n_b == 32)        // a whole, 16-byte aligned sector: two 128-bit loads
	{
		const uint4 v0 = __ldg((const uint4 *)row), v1 = __ldg((const uint4 *)row + 1);
		w[0] = v0.x; w[1] = v0.y; w[2] = v0.z; w[3] = v0.w;
		w[4] = v1.x; w[5] = v1.y; w[6] = v1.z; w[7] = v1.w;
		return;
	}
	const uint32_t *p = (const uint32_t *)(a & ~(uintptr_t)3);
	const int off = (int)(a & 3);
	const int last_word = (off + n_b - 1) >> 2;
	const unsigned sh = 8u * (unsigned)off;
	uint32_t prev = (n_b > 0) ? __ldg(p) : 0u;
#pragma unroll
	for (int q = 0; q < 8; q++)
	{
		const uint32_t next = (q + 1 <= last_word) ? __ldg(p + q + 1) : 0u;
		w[q] = __funnelshift_r(prev, next, sh);
		prev = next;
	}
}

__device__ __forceinline__ uint32_t bed_code_u8(uint32_t g)
{
	return (0x0001FF02u >> (8u * (g & 3u))) & 0xffu;
}

/// SNP-major payload [n_snp][bps], bps = ceil(n_samp / 4). A thread owns 4 consecutive selected
/// SNPs x 128 samples: per sample it packs its 4 genotypes into one 32-bit store, so a warp writes
/// 128 contiguous bytes of a sample row per instruction. VEC: n_save % 4 == 0 and out 4-byte aligned.
template <bool VEC>
__global__ void __launch_bounds__(128)
bed_snp_major_kernel(const uint8_t *__restrict__ payload, size_t bps, int n_samp,
	const int32_t *__restrict__ sel, int n_save, int8_t *__restrict__ out)
{
	// packed byte -> its 4 genotypes as bytes (sample 0 in the low byte)
	__shared__ uint32_t lut[256];
	for (int v = threadIdx.x; v < 256; v += 128)
		lut[v] = bed_code_u8(v) | (bed_code_u8(v >> 2) << 8) | (bed_code_u8(v >> 4) << 16) | (bed_code_u8(v >> 6) << 24);
	__syncthreads();
	const int k = (blockIdx.x * 128 + threadIdx.x) * 4;   // first of this thread's selected SNPs
	const int j0 = blockIdx.y * 128;                        // first sample of the tile
	if (k >= n_save) return;
	const int n_here = min(128, n_samp - j0);
	const int n_b = (n_here + 3) >> 2;
	uint32_t w[4][8];
#pragma unroll
	for (int c = 0; c < 4; c++)
	{
		if (k + c < n_save)
			load_row32(payload + (size_t)(sel ? __ldg(sel + k + c) : k + c) * bps + (size_t)(j0 >> 2), n_b, w[c]);
		else
		{
#pragma unroll
			for (int q = 0; q < 8; q++) w[c][q] = 0u;
		}
	}
	int8_t *o = out + (size_t)j0 * n_save + k;
	const bool full = (k + 3 < n_save);
#pragma unroll
	for (int q = 0; q < 8; q++)
	{
#pragma unroll
		for (int bb = 0; bb < 4; bb++)
		{
			const int j = q * 16 + bb * 4;                 // 4 samples = one packed byte per SNP
			if (j < n_here)
			{
				// 4 SNPs x 4 samples: table lookup per packed byte, then a 4x4 byte transpose
				const uint32_t W0 = lut[(w[0][q] >> (8 * bb)) & 0xffu], W1 = lut[(w[1][q] >> (8 * bb)) & 0xffu];
				const uint32_t W2 = lut[(w[2][q] >> (8 * bb)) & 0xffu], W3 = lut[(w[3][q] >> (8 * bb)) & 0xffu];
				const uint32_t t0 = __byte_perm(W0, W1, 0x5140), t1 = __byte_perm(W2, W3, 0x5140);
				const uint32_t t2 = __byte_perm(W0, W1, 0x7362), t3 = __byte_perm(W2, W3, 0x7362);
				uint32_t X[4];
				X[0] = __byte_perm(t0, t1, 0x5410); X[1] = __byte_perm(t0, t1, 0x7632);
				X[2] = __byte_perm(t2, t3, 0x5410); X[3] = __byte_perm(t2, t3, 0x7632);
#pragma unroll
				for (int s = 0; s < 4; s++)
				{
					if (j + s < n_here)
					{
						if (VEC && full)
							*(uint32_t *)(o + (size_t)(j + s) * n_save) = X[s];
						else
						{
#pragma unroll
							for (int c = 0; c < 4; c++)
								if (k + c < n_save) o[(size_t)(j + s) * n_save + c] = (int8_t)(X[s] >> (8 * c));
						}
					}
				}
			}
		}
	}
}

/// individual-major payload [n_samp][bps], bps = ceil(n_snp / 4): 4 selected SNPs per thread
template <bool VEC>
__global__ void __launch_bounds__(256)
bed_ind_major_kernel(const uint8_t *__restrict__ payload, size_t bps, int n_samp,
	const int32_t *__restrict__ sel, int n_save, int8_t *__restrict__ out)
{
	const int k = (blockIdx.x * 256 + threadIdx.x) * 4;
	const int j = blockIdx.y;
	if (k >= n_save || j >= n_samp) return;
	const uint8_t *row = payload + (size_t)j * bps;
	uint32_t g = 0;
#pragma unroll
	for (int c = 0; c < 4; c++)
		if (k + c < n_save)
		{
			const int snp = sel ? __ldg(sel + k + c) : k + c;
			g |= bed_code_u8((uint32_t)__ldg(row + (snp >> 2)) >> (2 * (snp & 3))) << (8 * c);
		}
	int8_t *o = out + (size_t)j * n_save + k;
	if (VEC && k + 3 < n_save)
		*(uint32_t *)o = g;
	else
	{
#pragma unroll
		for (int c = 0; c < 4; c++)
			if (k + c < n_save) o[c] = (int8_t)(g >> (8 * c));
	}
}

void launch_bed_decode(const uint8_t *payload, int mode, int n_samp, int n_snp, const int32_t *sel,
	int n_save, int8_t *out, cudaStream_t st)
{
	if (n_samp <= 0 || n_save <= 0) return;
	const bool vec = (n_save % 4 == 0) && (((uintptr_t)out & 3) == 0);
	if (mode == 0)
	{
		const size_t bps = ((size_t)n_snp + 3) / 4;
		dim3 grid((n_save + 1023) / 1024, n_samp);
		if (vec) bed_ind_major_kernel<true><<<grid, 256, 0, st>>>(payload, bps, n_samp, sel, n_save, out);
		else bed_ind_major_kernel<false><<<grid, 256, 0, st>>>(payload, bps, n_samp, sel, n_save, out);
	} else {
		const size_t bps = ((size_t)n_samp + 3) / 4;
		dim3 grid((n_save + 511) / 512, (n_samp + 127) / 128);
		if (vec) bed_snp_major_kernel<true><<<grid, 128, 0, st>>>(payload, bps, n_samp, sel, n_save, out);
		else bed_snp_major_kernel<false><<<grid, 128, 0, st>>>(payload, bps, n_samp, sel, n_save, out);
	}
	CUDA_CHECK(cudaGetLastError());
}

}  // namespace hb
