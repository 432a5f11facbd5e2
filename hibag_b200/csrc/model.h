// model.h -- the model container behind hibag_b200_model and its two drivers (internal)
#pragma once

#include <cstdint>
#include <memory>
#include <string>
#include <vector>

#include "hostalg.h"
#include "scorer.h"

namespace hb {

/// one individual classifier (reference CAttrBag_Classifier, src/LibHLA.h:552-597)
struct Classifier
{
	std::vector<int> snpidx;       // 0-based indices into the model's SNP list
	std::vector<int> samp_num;     // bootstrap multiplicities (may be empty for loaded models)
	HapList haplo;
	double oob_acc = 0;
};

struct PredictCache;   // device-resident copy of all classifiers (predictor.cu)

/// training state kept between hibag_b200_model_train calls (thread pool, streams, device
/// buffers); defined in trainer.cu
struct TrainSession
{
	virtual ~TrainSession() {}
};

}  // namespace hb

struct hibag_b200_model
{
	int n_snp = 0, n_hla = 0, n_samp = 0;
	std::vector<int8_t> geno_t;    // training genotypes, SNP-major [n_snp][n_samp]
	std::vector<int> h1, h2;       // training HLA types
	std::vector<hb::Classifier> cls;
	hibag_b200_train_stats train_stats;
	hibag_b200_predict_stats predict_stats;
	std::vector<int64_t> train_trace;   // rows of 4, see hibag_b200_model_train_trace
	std::shared_ptr<hb::PredictCache> pcache;
	std::shared_ptr<hb::TrainSession> tsession;
	// single-stream seeding (per_classifier_seed = 0): the RNG state is the MODEL's, seeded once per
	// seed value and continued by later train calls, as R's set.seed + repeated hlaAttrBagging do
	hb::RRng rng;
	bool rng_seeded = false;
	int64_t rng_seed = 0;
	hibag_b200_model();
};

namespace hb {

/// reference CAttrBag_Model::BuildClassifiers (src/LibHLA.cpp:2268) on the GPU scoring path
void train_model(hibag_b200_model &m, const hibag_b200_train_opts &opts);

/// batched ensemble prediction with device-resident inputs / outputs
void predict_device(hibag_b200_model &m, const int8_t *geno_dev, int n_samp,
	const hibag_b200_predict_out &out_dev, const int32_t *snp_weight_dev_or_null,
	double *partial_dev_or_null, cudaStream_t st, bool sync);
/// fold the distinct-genotype counts of finished asynchronous predict calls into m.predict_stats (waits)
void predict_collect_stats(hibag_b200_model &m);
void predict_host(hibag_b200_model &m, const int8_t *geno, int n_samp,
	const hibag_b200_predict_out &out);
void snp_weights(const hibag_b200_model &m, std::vector<int> &w);

// from plugin.cu
hibag_gpu_ext_proc *plugin_procs();
hibag_gpu_ext_proc *plugin_procs_with_haplomatch();
ScoreStats plugin_build_stats();
void score_host_arrays(int kind, const hibag_haplotype *haplo, int n_haplo, int n_hla,
	int n_snp, const hibag_genotype *geno, int n_geno, int32_t *out_a1, int32_t *out_a2,
	double *out_d, double *out_sum);

}  // namespace hb
