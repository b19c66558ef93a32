// TEST INFRASTRUCTURE — not part of the product path.
//
// C harness around the UNMODIFIED reference C++ (compiled from the sources where
// they lie under /root/reference by oracle/build.py; outputs go to oracle/_ref/).
// It exposes the reference's public operator API (Aligner::align / train,
// include/dynamont/aligner.hpp:56-85) and, for stage-level parity of the CUDA
// kernels, the private forward/backward passes (NT_aligner_api.hpp:41-66) through
// the "#define private public" trick described in SURVEY.md §8c.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may load the resulting library.
#define private public
#define protected public
#include "dynamont/NT_aligner_api.hpp"
#include "dynamont/NTK_aligner_api.hpp"
#undef private
#undef protected

#include <cmath>
#include <cstring>
#include <limits>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

using dynamont::PoreType;

namespace
{

struct RefHandle
{
	std::unique_ptr<dynamont::Aligner> aligner;
	bool basic = true;
};

void put_err(char* err, size_t errlen, const char* msg)
{
	if (err && errlen)
	{
		std::strncpy(err, msg, errlen - 1);
		err[errlen - 1] = 0;
	}
}

bool pore_from_string(const std::string& p, PoreType& out)
{
	// same table as aligner_bindings.cpp:18-32
	if (p == "rna002") out = PoreType::RNA002;
	else if (p == "rna004") out = PoreType::RNA004;
	else if (p == "dna_r9") out = PoreType::DNA_R9;
	else if (p == "dna_r10_260bps") out = PoreType::DNA_R10_260;
	else if (p == "dna_r10_400bps") out = PoreType::DNA_R10_400;
	else return false;
	return true;
}

} // namespace

extern "C"
{

void* ref_create(const char* model, const char* pore, const char* mode, int band, char* err, size_t errlen)
{
	try
	{
		PoreType pt;
		if (!pore_from_string(pore, pt))
			throw std::invalid_argument(std::string("Unknown pore type: ") + pore);
		auto h = std::make_unique<RefHandle>();
		const std::string m(mode);
		if (m == "basic" || m == "nt")
		{
			h->aligner = std::make_unique<dynamont::NTAligner>(model, pt, 1, (size_t)band);
			h->basic = true;
		}
		else if (m == "resquiggle" || m == "ntk")
		{
			h->aligner = std::make_unique<dynamont::NTKAligner>(model, pt, 1, (size_t)band);
			h->basic = false;
		}
		else
			throw std::invalid_argument("Unknown aligner mode: " + m);
		return h.release();
	}
	catch (const std::exception& e)
	{
		put_err(err, errlen, e.what());
		return nullptr;
	}
}

void ref_destroy(void* h) { delete static_cast<RefHandle*>(h); }

int ref_kmer_size(void* h) { return (int)static_cast<RefHandle*>(h)->aligner->kmerSize_; }
long ref_num_kmers(void* h) { return (long)static_cast<RefHandle*>(h)->aligner->numKmers_; }
int ref_is_rna(void* h) { return static_cast<RefHandle*>(h)->aligner->rna_ ? 1 : 0; }

// model table in native index order (aligner.cpp:136-141)
void ref_model(void* h, double* mean, double* stdev)
{
	auto& m = static_cast<RefHandle*>(h)->aligner->model_;
	for (size_t i = 0; i < m.size(); ++i)
	{
		mean[i] = m[i].mean;
		stdev[i] = m[i].stdev;
	}
}

// Aligner::align. Output arrays must hold L entries (>= Kc). polish (may be NULL)
// receives Kc fixed-width (k+1 bytes, NUL padded) strings. Returns 0 / 1 (exception, message in err).
int ref_align(void* h, const double* signal, size_t S, const char* seq, int calc_prob,
	double* Z, size_t* n_seg, size_t* seqpos, size_t* sigpos, double* prob, char* state, char* polish,
	char* err, size_t errlen)
{
	try
	{
		auto* a = static_cast<RefHandle*>(h)->aligner.get();
		dynamont::Result r = a->align(signal, S, std::string(seq), calc_prob != 0);
		*Z = r.Z;
		*n_seg = r.segments.size();
		const size_t kw = a->kmerSize_ + 1;
		for (size_t i = 0; i < r.segments.size(); ++i)
		{
			seqpos[i] = r.segments[i].sequencePosition;
			sigpos[i] = r.segments[i].signalPosition;
			prob[i] = r.segments[i].probability;
			if (state) state[i] = r.segments[i].state;
			if (polish)
			{
				std::memset(polish + i * kw, 0, kw);
				std::strncpy(polish + i * kw, r.segments[i].polish.c_str(), kw - 1);
			}
		}
		return 0;
	}
	catch (const std::exception& e)
	{
		put_err(err, errlen, e.what());
		return 1;
	}
}

// Aligner::train. mean/stdev hold numKmers entries.
int ref_train(void* h, const double* signal, size_t S, const char* seq,
	double* Z, double* trans3, double* mean, double* stdev, char* err, size_t errlen)
{
	try
	{
		auto* a = static_cast<RefHandle*>(h)->aligner.get();
		dynamont::TrainingResult r = a->train(signal, S, std::string(seq));
		*Z = r.Z;
		trans3[0] = r.transitions.m1;
		trans3[1] = r.transitions.e1;
		trans3[2] = r.transitions.e2;
		for (size_t i = 0; i < r.emissionModel.size(); ++i)
		{
			mean[i] = r.emissionModel[i].mean;
			stdev[i] = r.emissionModel[i].stdev;
		}
		return 0;
	}
	catch (const std::exception& e)
	{
		put_err(err, errlen, e.what());
		return 1;
	}
}

// Stage-level access (basic mode only): runs the reference's private computeBounds/forward/backward
// and returns Zf, Zb plus — if rows_t != NULL — the four band rows at each requested t, re-indexed
// from band storage to n = 0..N-1 (cells outside the band are -inf): out[(i*4 + {fM,fE,bM,bE})*N + n].
// Also returns the per-kmer raw sufficient statistics of runTraining (NT:490-514) if w/sx/sxx != NULL.
int ref_nt_stages(void* h, const double* signal, size_t S, const char* seq,
	double* Zf, double* Zb, const size_t* rows_t, size_t n_rows, double* rows_out,
	double* w, double* sx, double* sxx, char* err, size_t errlen)
{
	try
	{
		auto* rh = static_cast<RefHandle*>(h);
		if (!rh->basic)
			throw std::runtime_error("ref_nt_stages: basic mode only");
		auto* a = static_cast<dynamont::NTAligner*>(rh->aligner.get());
		const std::string sequence(seq);
		a->validateInput(S, sequence.size());
		std::vector<int> kmers = a->sequenceToKmers(sequence);
		const size_t T = S + 1, N = kmers.size() + 1;
		const size_t bw = std::min(a->bandwidth_, N / 2);
		const size_t B = 2 * bw + 3;
		auto bounds = a->computeBounds(T, N, bw);
		const double NEG = -std::numeric_limits<double>::infinity();
		std::vector<double> fM(T * B, NEG), fE(T * B, NEG), bM(T * B, NEG), bE(T * B, NEG);
		a->forward(signal, kmers.data(), fM.data(), fE.data(), T, B, bw, bounds);
		a->backward(signal, kmers.data(), bM.data(), bE.data(), T, N, B, bw, bounds);
		*Zf = fE[T * B - bw - 2];
		*Zb = bE[bw + 1];
		for (size_t i = 0; i < n_rows; ++i)
		{
			const size_t t = rows_t[i];
			const double* src[4] = {fM.data(), fE.data(), bM.data(), bE.data()};
			for (int m = 0; m < 4; ++m)
				for (size_t n = 0; n < N; ++n)
				{
					double v = NEG;
					if (n >= bounds[t].nStart && n < bounds[t].nEnd)
						v = src[m][(long)(t * B) + (long)n - bounds[t].start + 1];
					rows_out[(i * 4 + m) * N + n] = v;
				}
		}
		if (w && sx && sxx)
		{
			const size_t K = a->numKmers_;
			for (size_t k = 0; k < K; ++k) w[k] = sx[k] = sxx[k] = 0.0;
			const double Z = *Zb;
			for (size_t t = 1; t < T; ++t)
			{
				auto [bandStart, nStart, nEnd] = bounds[t];
				if (!nStart) nStart = 1;
				for (size_t n = nStart; n < nEnd; ++n)
				{
					const size_t idx = n + (long)(t * B) - bandStart + 1;
					const double post = std::exp(fM[idx] + bM[idx] - Z) + std::exp(fE[idx] + bE[idx] - Z);
					const int kmer = kmers[n - 1];
					const double obs = signal[t - 1];
					w[kmer] += post;
					sx[kmer] += post * obs;
					sxx[kmer] += post * obs * obs;
				}
			}
		}
		return 0;
	}
	catch (const std::exception& e)
	{
		put_err(err, errlen, e.what());
		return 1;
	}
}


// Resquiggle (NTK) mode, stages that work in the reference as shipped (NTK_aligner_api.cpp:197-441): the dense TN
// and TK pre-passes with their 95 %-mass row masks, and the sorted key list of the sparse lattice.
//   tn_mask [T*N] / tk_mask [T*K] bytes (1 = member of tnMap[t] / tkMap[t]); keys: up to keys_cap entries.
// z4 = { Zf_TN, Zb_TN, Zf_TK, Zb_TK } recomputed from the private pre-pass routines.
int ref_ntk_prepass(void* h, const double* signal, size_t S, const char* seq, unsigned char* tn_mask,
	unsigned char* tk_mask, unsigned long long* keys, size_t keys_cap, size_t* n_keys, double* z4, double* trans18,
	char* err, size_t errlen)
{
	try
	{
		auto* a = dynamic_cast<dynamont::NTKAligner*>(static_cast<RefHandle*>(h)->aligner.get());
		if (!a) throw std::runtime_error("not a resquiggle-mode aligner");
		const std::string sequence(seq);
		a->validateInput(S, sequence.size());
		const std::vector<int> kmers = a->sequenceToKmers(sequence);
		const size_t T = S + 1, N = kmers.size() + 1, K = a->numKmers_;
		dynamont::NTKAligner::ColumnMask tn, tk;
		a->preProcTN(signal, kmers.data(), tn, T, N);
		a->preProcTK(signal, tk, T, K);
		std::memset(tn_mask, 0, T * N);
		std::memset(tk_mask, 0, T * K);
		for (size_t t = 0; t < T; ++t)
		{
			for (size_t n : tn[t]) tn_mask[t * N + n] = 1;
			for (size_t k : tk[t]) tk_mask[t * K + k] = 1;
		}
		const std::vector<size_t> allowed = a->preProcTNK(signal, kmers.data(), T, N, K);
		*n_keys = allowed.size();
		for (size_t i = 0; i < allowed.size() && i < keys_cap; ++i) keys[i] = allowed[i];
		if (z4)
		{
			const double NI = -std::numeric_limits<double>::infinity();
			std::vector<double> fM(T * N, NI), fE(T * N, NI), bM(T * N, NI), bE(T * N, NI);
			a->ppForTN(signal, kmers.data(), fM.data(), fE.data(), T, N);
			a->ppBackTN(signal, kmers.data(), bM.data(), bE.data(), T, N);
			z4[0] = fE[T * N - 1];
			z4[1] = bE[0];
			std::vector<double> gM(T * K, NI), gE(T * K, NI), hM(T * K, NI), hE(T * K, NI);
			a->ppForTK(signal, gM.data(), gE.data(), T, K);
			a->ppBackTK(signal, hM.data(), hE.data(), T, K);
			double Zf = NI, Zb = NI;
			for (size_t k = 0; k < K; ++k)
			{
				Zf = dynamont::Aligner::logPlus(Zf, gE[T * K - 1 - k]);
				Zb = dynamont::Aligner::logPlus(Zb, hE[k]);
			}
			z4[2] = Zf;
			z4[3] = Zb;
		}
		if (trans18)
		{
			static const char* names[14] = {"a1", "a2", "p1", "p2", "p3", "s1", "s2", "s3", "e1", "e2", "e3", "e4", "i1", "i2"};
			for (int i = 0; i < 14; ++i) trans18[i] = a->transitions_.at(names[i]);
			trans18[14] = a->ppTNm_;
			trans18[15] = a->ppTNe_;
			trans18[16] = a->ppTKm_;
			trans18[17] = a->ppTKe_;
		}
		return 0;
	}
	catch (const std::exception& e)
	{
		put_err(err, errlen, e.what());
		return 1;
	}
}

} // extern "C"
