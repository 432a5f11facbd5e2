// tasks.h -- host-side preparation of one haplotype list for the scoring kernel (internal)
#pragma once

#include <cstdint>
#include <vector>

#include "common.h"
#include "kernels.h"

namespace hb {

/// EXP_LOG_MIN_RARE_FREQ[257] computed on the host with libm exactly as the reference's static
/// initialiser does (src/LibHLA.cpp:166-183); never recomputed on the device.
const double *host_rare_freq_table();
/// smallest d with T[d'] == 0 for all d' >= d (65 with IEEE doubles)
int rare_freq_first_zero();
/// number of table rows the kernel needs for n_snp SNPs
int table_rows_for(int n_snp);

/// Screening (kernels.h): T' = max(T, 1e-100), and the factor K of the cell bound
/// U_a * U_b * K: 2 (the c_ij of off-diagonal pairs) x kappa x (1 + 1e-8), with
/// kappa = max_{p,q} max_{d >= p+q} T[d] / (T'[p] T'[q]) taken over the host table itself, so
/// that T[d(g,i,j)] <= kappa T'[c_i] T'[c_j] whenever d >= c_i + c_j; 1e-8 covers the rounding
/// of the chains (< 4e7 terms) and of the bound's own arithmetic.
const double *host_rare_freq_floor_table();
double screen_bound_factor();
/// the same for the second-level bound over the classes of two heterozygous SNPs (three table factors)
double screen_bound_factor2();

/// A haplotype list laid out for one H2D copy: [records | cells | chunks]
struct ListBlob
{
	int n_hap = 0, n_snp = 0, n_hla = 0, n_dist = 0;
	int n_cells = 0, n_chunks = 0;
	size_t off_cells = 0, off_chunks = 0, bytes = 0;
	uint64_t pairs_per_sample = 0;     // haplotype pairs (i<=j) one sample is scored against
};

/// bytes needed for a list of n_hap haplotypes / n_hla alleles
size_t list_blob_capacity(int n_hap, int n_snp, int n_hla);

/// Fill `dst` (>= list_blob_capacity bytes, 16-byte aligned) from a THaplotype array whose
/// aux.a2.HLA_allele tags are non-decreasing (src/LibHLA.cpp:565-578); derives LenPerHLA,
/// builds the cell list sorted by decreasing cost and cuts it into chunks.
/// Throws if the tags are not sorted or out of range.
ListBlob build_list_blob(const hibag_haplotype *haplo, int n_hap, int n_hla, int n_snp,
	void *dst, int target_chunks = 512);

/// Resolve device pointers of a staged blob into a CellPass
void bind_list(const ListBlob &b, const void *dev_blob, const double *dev_table, CellPass &p);

}  // namespace hb
