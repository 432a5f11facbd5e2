// ===========================================================================
// oracle/ref_driver.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// A thin extern "C" driver linked together with the UNMODIFIED reference sources
// (/root/reference/src/LibHLA.cpp + LibHLA_ext_*.cpp, compiled where they lie by
// oracle/Makefile) into oracle/_ref/libhibag_ref.so.  It replaces the R glue
// src/HIBAG.cpp (which needs the real R API) with a ctypes-friendly surface:
//
//   * R's Mersenne-Twister `unif_rand` and `set.seed` (so R seeds reproduce),
//   * CAttrBag_Model training / prediction (reference LibHLA.cpp:2268, 2317),
//   * the four scoring kernels of any compiled target by name
//     (reference LibHLA.cpp:1569-1830 and LibHLA_ext_*.cpp),
//   * installation of a TypeGPUExtProc plugin (reference LibHLA.cpp:193).
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may load this library. Nothing under hibag_b200/ links or loads it.
// ===========================================================================

#include <algorithm>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <ctime>
#include <cmath>
#include <list>
#include <stdexcept>
#include <string>
#include <vector>
#include <chrono>

// the scoring kernels and the search internals are private/protected statics of the
// reference classes; expose them to this driver only (layout is unaffected)
#define private public
#define protected public
#include "LibHLA.h"
#undef private
#undef protected

#include <R.h>

using namespace HLA_LIB;

// ---------------------------------------------------------------------------
// R's Mersenne-Twister (R: src/main/RNG.c, MT_genrand / Randomize / FixupSeeds)
// ---------------------------------------------------------------------------
namespace {

const int MT_N = 624, MT_M = 397;
uint32_t mt_state[MT_N];
int mt_idx = MT_N + 1;
unsigned long long rng_draws = 0;

void mt_sgenrand(uint32_t seed)
{
	for (int i = 0; i < MT_N; i++)
	{
		mt_state[i] = seed & 0xffff0000u;
		seed = 69069u * seed + 1;
		mt_state[i] |= (seed & 0xffff0000u) >> 16;
		seed = 69069u * seed + 1;
	}
	mt_idx = MT_N;
}

double mt_genrand()
{
	static const uint32_t mag01[2] = { 0x0u, 0x9908b0dfu };
	uint32_t y;
	if (mt_idx >= MT_N)
	{
		if (mt_idx == MT_N + 1) mt_sgenrand(4357);
		int kk;
		for (kk = 0; kk < MT_N - MT_M; kk++)
		{
			y = (mt_state[kk] & 0x80000000u) | (mt_state[kk + 1] & 0x7fffffffu);
			mt_state[kk] = mt_state[kk + MT_M] ^ (y >> 1) ^ mag01[y & 1u];
		}
		for (; kk < MT_N - 1; kk++)
		{
			y = (mt_state[kk] & 0x80000000u) | (mt_state[kk + 1] & 0x7fffffffu);
			mt_state[kk] = mt_state[kk + (MT_M - MT_N)] ^ (y >> 1) ^ mag01[y & 1u];
		}
		y = (mt_state[MT_N - 1] & 0x80000000u) | (mt_state[0] & 0x7fffffffu);
		mt_state[MT_N - 1] = mt_state[MT_M - 1] ^ (y >> 1) ^ mag01[y & 1u];
		mt_idx = 0;
	}
	y = mt_state[mt_idx++];
	y ^= (y >> 11);
	y ^= (y << 7) & 0x9d2c5680u;
	y ^= (y << 15) & 0xefc60000u;
	y ^= (y >> 18);
	return (double)y * 2.3283064365386963e-10;
}

std::string last_error;

// interrupt plumbing: lets a caller bound a training run (bench.py's bounded sample)
double interrupt_deadline = 0;          // seconds (steady clock); 0 = none
long long interrupt_after_checks = -1;  // stop after this many CheckInterrupt calls
long long interrupt_checks = 0;

double now_s()
{
	using namespace std::chrono;
	return duration<double>(steady_clock::now().time_since_epoch()).count();
}

struct Interrupted {};

}  // namespace

extern "C" double unif_rand(void)
{
	rng_draws++;
	double x = mt_genrand();
	const double i2_32m1 = 2.328306437080797e-10;
	if (x <= 0.0) return 0.5 * i2_32m1;
	if ((1.0 - x) <= 0.0) return 1.0 - 0.5 * i2_32m1;
	return x;
}

extern "C" void Rprintf(const char *fmt, ...)
{
	va_list args;
	va_start(args, fmt);
	vfprintf(stderr, fmt, args);
	va_end(args);
}

extern "C" void Rf_error(const char *fmt, ...)
{
	char buf[1024];
	va_list args;
	va_start(args, fmt);
	vsnprintf(buf, sizeof(buf), fmt, args);
	va_end(args);
	throw std::runtime_error(buf);
}

extern "C" void R_CheckUserInterrupt(void)
{
	interrupt_checks++;
	if (interrupt_after_checks >= 0 && interrupt_checks >= interrupt_after_checks)
		throw Interrupted();
	if (interrupt_deadline > 0 && now_s() >= interrupt_deadline)
		throw Interrupted();
}


// ---------------------------------------------------------------------------
// helpers
// ---------------------------------------------------------------------------
namespace {

struct RefModel
{
	CAttrBag_Model model;
	std::vector<int> geno;   // [n_samp][n_snp], sample-major (reference LibHLA.cpp:841-849)
};

#define REF_TRY      try {
#define REF_CATCH    \
	} catch (Interrupted &) { last_error = "interrupted"; return 1; } \
	catch (std::exception &e) { last_error = e.what(); return -1; } \
	catch (const char *e) { last_error = e; return -1; } \
	catch (...) { last_error = "unknown error"; return -1; } \
	return 0;

struct ScoreFuncs
{
	CAlg_Prediction::F_BestGuess best_guess;
	CAlg_Prediction::F_PostProb post_prob;
	CAlg_Prediction::F_PostProb2 post_prob2;
	CAlg_Prediction::F_PrepHaploMatch prep_match;
	bool need_aux;
};

bool pick_funcs(const char *target, ScoreFuncs &f)
{
	std::string t = target ? target : "base";
#define SET(SUFFIX, AUX) \
	{ f.best_guess = &CAlg_Prediction::_BestGuess_##SUFFIX; \
	  f.post_prob = &CAlg_Prediction::_PostProb_##SUFFIX; \
	  f.post_prob2 = &CAlg_Prediction::_PostProb2_##SUFFIX; \
	  f.prep_match = &CAlg_Prediction::_PrepHaploMatch_##SUFFIX; \
	  f.need_aux = AUX; return true; }
	if (t == "base") SET(def, false)
	if (t == "sse2") SET(sse2, false)
	if (t == "sse4") SET(sse4_2, false)
	if (t == "avx") SET(avx, true)
	if (t == "avx2") SET(avx2, true)
	if (t == "avx512f") SET(avx512f, true)
	if (t == "avx512bw") SET(avx512bw, true)
	if (t == "avx512vpopcnt") SET(avx512vpopcnt, true)
#undef SET
	return false;
}

// build a CHaplotypeList from a flat THaplotype array whose aux.a2.HLA_allele is filled
void make_haplo_list(CHaplotypeList &hl, const THaplotype *haplo, int n_haplo,
	int n_hla, int n_snp, std::vector<int64_t> &aux_h, std::vector<double> &aux_f,
	bool need_aux)
{
	hl.Num_SNP = n_snp;
	hl.ResizeHaplo(n_haplo);
	memcpy(hl.List, haplo, sizeof(THaplotype) * (size_t)n_haplo);
	hl.LenPerHLA.assign(n_hla, 0);
	for (int i = 0; i < n_haplo; i++)
		hl.LenPerHLA[haplo[i].aux.a2.HLA_allele]++;
	if (need_aux)
	{
		aux_h.resize(2 * (size_t)n_haplo + 16);
		aux_f.resize((size_t)n_haplo + 16);
		hl.SetHaploAux(&aux_h[0], &aux_f[0]);
	}
}

}  // namespace


extern "C" {

// ---- misc ----------------------------------------------------------------

const char *ref_last_error() { return last_error.c_str(); }

/// R: set.seed(seed) with the default "Mersenne-Twister" generator
void ref_set_seed(unsigned seed)
{
	uint32_t s = seed;
	for (int j = 0; j < 50; j++) s = 69069u * s + 1;
	// i_seed[0] is the 'mti' slot, overwritten by FixupSeeds with N
	s = 69069u * s + 1;
	for (int j = 0; j < MT_N; j++)
	{
		s = 69069u * s + 1;
		mt_state[j] = s;
	}
	mt_idx = MT_N;
	rng_draws = 0;
}

double ref_unif_rand() { return unif_rand(); }
unsigned long long ref_rng_draws() { return rng_draws; }

/// reference LibHLA.cpp:1266 (target strings "base","sse2","sse4","avx","avx2","avx512f",
/// "avx512bw","avx512vpopcnt","max","auto.avx2")
int ref_set_target(const char *cpu)
{
	REF_TRY
		CAlg_Prediction::Init_Target_IFunc(cpu);
	REF_CATCH
}

const char *ref_cpu_info() { return CPU_Info(); }

/// the 257-entry table built by the reference's static initialiser (LibHLA.cpp:166-183)
const double *ref_exp_log_min_rare_freq() { return EXP_LOG_MIN_RARE_FREQ; }

/// install (or clear with NULL) a GPU plugin, reference LibHLA.cpp:193 / HIBAG.cpp:559-573
void ref_set_gpu_procs(void *procs) { GPUExtProcPtr = (TypeGPUExtProc *)procs; }

void ref_set_interrupt(double seconds_from_now, long long after_checks)
{
	interrupt_deadline = (seconds_from_now > 0) ? now_s() + seconds_from_now : 0;
	interrupt_after_checks = after_checks;
	interrupt_checks = 0;
}
long long ref_interrupt_checks() { return interrupt_checks; }

int ref_sizeof_haplotype() { return (int)sizeof(THaplotype); }
int ref_sizeof_genotype() { return (int)sizeof(TGenotype); }


// ---- model: training -------------------------------------------------------

void *ref_model_new() { return new RefModel(); }
void ref_model_free(void *m) { delete (RefModel *)m; }

/// geno: int[n_samp][n_snp] sample-major, values 0/1/2, anything else = missing
int ref_model_init_training(void *mp, int n_snp, int n_samp, const int *geno,
	int n_hla, const int *H1, const int *H2)
{
	RefModel *m = (RefModel *)mp;
	REF_TRY
		m->geno.assign(geno, geno + (size_t)n_snp * n_samp);
		std::vector<int> h1(H1, H1 + n_samp), h2(H2, H2 + n_samp);
		m->model.InitTraining(n_snp, n_samp, &m->geno[0], n_hla, &h1[0], &h2[0]);
	REF_CATCH
}

/// model skeleton for prediction only (reference HIBAG.cpp:486 HIBAG_New)
int ref_model_init_predict(void *mp, int n_snp, int n_samp, int n_hla)
{
	RefModel *m = (RefModel *)mp;
	REF_TRY
		m->model.InitTraining(n_snp, n_samp, n_hla);
	REF_CATCH
}

/// reference LibHLA.cpp:2268. reseed_base < 0: one RNG stream across classifiers (what R
/// does); otherwise classifier k (global index first_index+k) is built after
/// set.seed(reseed_base + first_index + k) -- the per-classifier convention the multi-GPU
/// trainer uses.
int ref_model_build(void *mp, int nclassifier, int mtry, int prune, int verbose,
	long long reseed_base, int first_index)
{
	RefModel *m = (RefModel *)mp;
	REF_TRY
		if (reseed_base < 0)
		{
			m->model.BuildClassifiers(nclassifier, mtry, prune != 0, verbose != 0, verbose > 1);
		} else {
			for (int k = 0; k < nclassifier; k++)
			{
				ref_set_seed((unsigned)(reseed_base + first_index + k));
				m->model.BuildClassifiers(1, mtry, prune != 0, verbose != 0, verbose > 1);
			}
		}
	REF_CATCH
}

int ref_model_num_classifiers(void *mp)
{
	return (int)((RefModel *)mp)->model._ClassifierList.size();
}

void ref_model_clear(void *mp) { ((RefModel *)mp)->model.ClearClassifierList(); }

int ref_classifier_info(void *mp, int k, int *n_snp, int *n_haplo, double *oob_acc)
{
	RefModel *m = (RefModel *)mp;
	if (k < 0 || k >= (int)m->model._ClassifierList.size()) return -1;
	const CAttrBag_Classifier &c = m->model._ClassifierList[k];
	*n_snp = c.nSNP();
	*n_haplo = c.nHaplo();
	*oob_acc = c.OutOfBag_Accuracy();
	return 0;
}

/// packed: int64[n_haplo][2] with bits >= n_snp cleared
int ref_classifier_get(void *mp, int k, int *snpidx, int *bootstrap, double *freq,
	int *hla, int64_t *packed)
{
	RefModel *m = (RefModel *)mp;
	if (k < 0 || k >= (int)m->model._ClassifierList.size()) return -1;
	const CAttrBag_Classifier &c = m->model._ClassifierList[k];
	const int n_snp = c.nSNP();
	if (snpidx) std::copy(c._SNPIndex.begin(), c._SNPIndex.end(), snpidx);
	if (bootstrap) std::copy(c._BootstrapCount.begin(), c._BootstrapCount.end(), bootstrap);
	const CHaplotypeList &hl = c._Haplo;
	size_t idx = 0;
	for (size_t a = 0; a < hl.LenPerHLA.size(); a++)
	{
		for (size_t n = hl.LenPerHLA[a]; n > 0; n--, idx++)
		{
			if (freq) freq[idx] = hl.List[idx].Freq;
			if (hla) hla[idx] = (int)a;
			if (packed)
			{
				uint64_t w0 = (uint64_t)hl.List[idx].PackedHaplo[0];
				uint64_t w1 = (uint64_t)hl.List[idx].PackedHaplo[1];
				if (n_snp < 64) { w0 &= (n_snp ? ((~0ULL) >> (64 - n_snp)) : 0ULL); w1 = 0; }
				else if (n_snp < 128) { w1 &= ((n_snp > 64) ? ((~0ULL) >> (128 - n_snp)) : 0ULL); }
				packed[2 * idx] = (int64_t)w0;
				packed[2 * idx + 1] = (int64_t)w1;
			}
		}
	}
	return 0;
}

/// reference HIBAG.cpp:817 HIBAG_NewClassifierHaplo -> CAttrBag_Classifier::Assign (:2142)
/// hla[] must be non-decreasing (haplotypes grouped by allele), snpidx 0-based
int ref_model_add_classifier(void *mp, int n_snp, const int *snpidx, const int *samp_num,
	int n_haplo, const double *freq, const int *hla, const int64_t *packed, double acc)
{
	RefModel *m = (RefModel *)mp;
	REF_TRY
		std::vector<std::string> strs(n_haplo);
		std::vector<const char *> ptrs(n_haplo);
		for (int i = 0; i < n_haplo; i++)
		{
			std::string &s = strs[i];
			s.resize(n_snp);
			for (int j = 0; j < n_snp; j++)
				s[j] = (((uint64_t)packed[2 * i + (j >> 6)] >> (j & 63)) & 1) ? '1' : '0';
			ptrs[i] = s.c_str();
		}
		CAttrBag_Classifier *I = m->model.NewClassifierAllSamp();
		I->Assign(n_snp, snpidx, samp_num, n_haplo, freq, hla,
			n_haplo ? &ptrs[0] : NULL, &acc);
	REF_CATCH
}


// ---- model: prediction -------------------------------------------------------

/// reference LibHLA.cpp:2317. geno: int[n_samp][n_snp(model)] sample-major. Any output may be NULL.
int ref_model_predict(void *mp, const int *geno, int n_samp, int vote_method,
	int *out_h1, int *out_h2, double *out_maxprob, double *out_matching,
	double *out_dosage, double *out_prob)
{
	RefModel *m = (RefModel *)mp;
	REF_TRY
		m->model.PredictHLA(geno, n_samp, vote_method, out_h1, out_h2, out_maxprob,
			out_matching, out_dosage, out_prob, false);
	REF_CATCH
}


// ---- scoring kernels on raw arrays -------------------------------------------

/// BestGuess for every genotype (reference LibHLA.cpp:1639 / per-target copies).
/// haplo[].aux.a2.HLA_allele must be filled and non-decreasing.
int ref_best_guess(const char *target, const void *haplo, int n_haplo, int n_hla,
	int n_snp, const void *geno, int n_geno, int *out_a1, int *out_a2)
{
	REF_TRY
		ScoreFuncs f;
		if (!pick_funcs(target, f)) throw std::runtime_error("unknown target");
		CHaplotypeList hl;
		std::vector<int64_t> ah; std::vector<double> af;
		make_haplo_list(hl, (const THaplotype *)haplo, n_haplo, n_hla, n_snp, ah, af, f.need_aux);
		const TGenotype *g = (const TGenotype *)geno;
		for (int i = 0; i < n_geno; i++)
		{
			THLAType t = (*f.best_guess)(hl, g[i]);
			out_a1[i] = t.Allele1; out_a2[i] = t.Allele2;
		}
	REF_CATCH
}

/// PostProb of each genotype's own aux_hla_type (reference LibHLA.cpp:1706)
int ref_post_prob(const char *target, const void *haplo, int n_haplo, int n_hla,
	int n_snp, const void *geno, int n_geno, double *out)
{
	REF_TRY
		ScoreFuncs f;
		if (!pick_funcs(target, f)) throw std::runtime_error("unknown target");
		CHaplotypeList hl;
		std::vector<int64_t> ah; std::vector<double> af;
		make_haplo_list(hl, (const THaplotype *)haplo, n_haplo, n_hla, n_snp, ah, af, f.need_aux);
		const TGenotype *g = (const TGenotype *)geno;
		for (int i = 0; i < n_geno; i++)
			out[i] = (*f.post_prob)(hl, g[i], g[i].aux_hla_type);
	REF_CATCH
}

/// PostProb2: normalised posterior over all cells + raw sum (reference LibHLA.cpp:1769)
int ref_post_prob2(const char *target, const void *haplo, int n_haplo, int n_hla,
	int n_snp, const void *geno, int n_geno, double *out_prob, double *out_sum)
{
	REF_TRY
		ScoreFuncs f;
		if (!pick_funcs(target, f)) throw std::runtime_error("unknown target");
		CHaplotypeList hl;
		std::vector<int64_t> ah; std::vector<double> af;
		make_haplo_list(hl, (const THaplotype *)haplo, n_haplo, n_hla, n_snp, ah, af, f.need_aux);
		const TGenotype *g = (const TGenotype *)geno;
		const size_t nc = (size_t)n_hla * (n_hla + 1) / 2;
		for (int i = 0; i < n_geno; i++)
			out_sum[i] = (*f.post_prob2)(hl, g[i], out_prob + nc * i);
	REF_CATCH
}

/// the two training reductions exactly as the reference's CPU branch does them
/// (LibHLA.cpp:1941-1954 and 1965-1977): geno[] = ALL samples, OOB <=> BootstrapCount == 0
int ref_acc_oob(const char *target, const void *haplo, int n_haplo, int n_hla,
	int n_snp, const void *geno, int n_geno, int *out_acc)
{
	REF_TRY
		ScoreFuncs f;
		if (!pick_funcs(target, f)) throw std::runtime_error("unknown target");
		CHaplotypeList hl;
		std::vector<int64_t> ah; std::vector<double> af;
		make_haplo_list(hl, (const THaplotype *)haplo, n_haplo, n_hla, n_snp, ah, af, f.need_aux);
		const TGenotype *g = (const TGenotype *)geno;
		int cnt = 0;
		for (int i = 0; i < n_geno; i++)
		{
			if (g[i].BootstrapCount != 0) continue;
			THLAType t = (*f.best_guess)(hl, g[i]);
			cnt += CHLATypeList::Compare(t, g[i].aux_hla_type);
		}
		*out_acc = cnt;
	REF_CATCH
}

int ref_acc_ib(const char *target, const void *haplo, int n_haplo, int n_hla,
	int n_snp, const void *geno, int n_geno, double *out_loss)
{
	REF_TRY
		ScoreFuncs f;
		if (!pick_funcs(target, f)) throw std::runtime_error("unknown target");
		CHaplotypeList hl;
		std::vector<int64_t> ah; std::vector<double> af;
		make_haplo_list(hl, (const THaplotype *)haplo, n_haplo, n_hla, n_snp, ah, af, f.need_aux);
		const TGenotype *g = (const TGenotype *)geno;
		double LogLik = 0;
		for (int i = 0; i < n_geno; i++)
		{
			if (g[i].BootstrapCount <= 0) continue;
			LogLik += g[i].BootstrapCount * log((*f.post_prob)(hl, g[i], g[i].aux_hla_type));
		}
		LogLik *= -2;
		*out_loss = LogLik;
	REF_CATCH
}

/// TGenotype::IntToSNP (reference LibHLA.cpp:667)
int ref_int_to_snp(void *out_geno, int length, const int *geno_base, const int *index)
{
	REF_TRY
		((TGenotype *)out_geno)->IntToSNP(length, geno_base, index);
	REF_CATCH
}

int ref_hamming(const void *geno, const void *h1, const void *h2, int n_snp)
{
	return ((const TGenotype *)geno)->HammingDistance(n_snp, *(const THaplotype *)h1,
		*(const THaplotype *)h2);
}

/// the reference's own pair matcher (CAlg_Prediction::_PrepHaploMatch_def, LibHLA.cpp:1569-1637)
/// on two allele ranges [st1, st1+m1), [st2, st2+m2) of a haplotype list; pairs are returned as
/// offsets (i1, i2) inside the two ranges. Returns the number of pairs (<= max_pairs).
int ref_prep_haplo_match(const void *geno, const void *haplo, int st1, int m1, int st2, int m2,
	int n_snp, int *out_i1, int *out_i2, int max_pairs)
{
	THaplotype *H = (THaplotype *)haplo;
	std::vector<CAlg_EM::THaploPair> pl;
	std::vector<short> diff((size_t)m1 * (size_t)(m2 > m1 ? m2 : m1) + 1);
	CAlg_Prediction::_PrepHaploMatch_def(*(const TGenotype *)geno, H + st1, (size_t)m1, H + st2,
		(size_t)m2, (size_t)n_snp, pl, diff.data());
	int n = 0;
	for (size_t k = 0; k < pl.size() && n < max_pairs; k++, n++)
	{
		out_i1[n] = (int)(pl[k].H1 - (H + st1));
		out_i2[n] = (int)(pl[k].H2 - (H + st2));
	}
	return (int)pl.size();
}

}  // extern "C"

// CHLATypeList::Compare is declared inline in the reference's .cpp (LibHLA.cpp:912) and is
// therefore not linkable from here; restated (8 lines) for ref_acc_oob above.
inline int HLA_LIB::CHLATypeList::Compare(const THLAType &H1, const THLAType &H2)
{
	int P1 = H1.Allele1, P2 = H1.Allele2;
	int T1 = H2.Allele1, T2 = H2.Allele2;
	int cnt = 0;
	if ((P1 == T1) || (P1 == T2))
	{
		cnt = 1;
		if (P1 == T1) T1 = -1; else T2 = -1;
	}
	if ((P2 == T1) || (P2 == T2)) cnt++;
	return cnt;
}
