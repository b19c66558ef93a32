// dynamont-NT-b200 — streaming stdin/stdout front end of the C ABI (include/dynamont_b200.h).
//
// The reference snapshot no longer ships its legacy pipe binary ("dynamont-NT"); all that is left of it is
// the call site reference src/dynamont/plot/plotToolSegments.py:54,69-78 with the options m (model path),
// r (pore), p (emit probabilities) and t (threads).  The protocol is therefore defined here:
//
//   stdin, per read two lines:   <sample>,<sample>,...            (decimal, already normalised)
//                                <sequence in signal orientation>
//   stdout, per read one line:   M<basepos>,<sigstart>,<prob>;...;\tZ:<Z>      (align, -p given)
//                                Z:<Z>                                          (align, no -p)
//                                m1:<m1>;e1:<e1>;e2:<e2>\tZ:<Z>                 (--train; pooled model goes to --out-model)
//                                error:<reference message>                      (reads the reference would throw on)
//
// Reads are batched (--batch N, default 4096) so the GPU always has thousands of reads in flight; output
// order = input order.  Samples are parsed with strtod and rounded to FP32 (round-to-nearest).
#include "../../include/dynamont_b200.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <string>
#include <vector>

namespace
{

struct Options
{
	std::string model, pore = "rna002", mode = "basic", out_model;
	int threads = 1, band = 400, device = -1, batch = 4096;
	bool prob = false, train = false;
};

void usage()
{
	std::fprintf(stderr,
		"usage: dynamont-NT-b200 -m MODEL -r PORE [-p] [-t THREADS] [-b BAND] [--mode basic] [--batch N] [--device D]\n"
		"                        [--train --out-model PATH]\n");
}

bool parse(int argc, char** argv, Options& o)
{
	for (int i = 1; i < argc; ++i)
	{
		const std::string a = argv[i];
		auto next = [&](const char* what) -> const char* {
			if (i + 1 >= argc)
			{
				std::fprintf(stderr, "missing value for %s\n", what);
				std::exit(2);
			}
			return argv[++i];
		};
		if (a == "-m" || a == "--model") o.model = next("-m");
		else if (a == "-r" || a == "--pore") o.pore = next("-r");
		else if (a == "-p" || a == "--probabilities") o.prob = true;
		else if (a == "-t" || a == "--threads") o.threads = std::atoi(next("-t"));
		else if (a == "-b" || a == "--band") o.band = std::atoi(next("-b"));
		else if (a == "--mode") o.mode = next("--mode");
		else if (a == "--batch") o.batch = std::atoi(next("--batch"));
		else if (a == "--device") o.device = std::atoi(next("--device"));
		else if (a == "--train") o.train = true;
		else if (a == "--out-model") o.out_model = next("--out-model");
		else if (a == "-h" || a == "--help") return false;
		else
		{
			std::fprintf(stderr, "unknown option %s\n", a.c_str());
			return false;
		}
	}
	return !o.model.empty();
}

struct Batch
{
	std::vector<float> sig;
	std::vector<uint64_t> sig_off{0}, seq_off{0};
	std::string seq;
	uint32_t n = 0;
	void clear()
	{
		sig.clear();
		seq.clear();
		sig_off.assign(1, 0);
		seq_off.assign(1, 0);
		n = 0;
	}
};

void parse_signal(const std::string& line, std::vector<float>& out)
{
	const char* p = line.c_str();
	char* end = nullptr;
	while (*p)
	{
		const double v = std::strtod(p, &end);
		if (end == p) break;
		out.push_back((float)v);
		p = end;
		while (*p == ',' || *p == ' ' || *p == '\t' || *p == '\r') ++p;
	}
}

std::string status_text(int status, char bad)
{
	std::string msg = dyn_status_message(status);
	if (status == DYN_INVALID_NT) msg += bad;
	return msg;
}

} // namespace

int main(int argc, char** argv)
{
	Options o;
	if (!parse(argc, argv, o))
	{
		usage();
		return 2;
	}
	char err[1024] = {0};
	int kind = 0;
	dyn_aligner* A = dyn_create(o.model.c_str(), o.pore.c_str(), o.mode.c_str(), o.threads, o.band, o.device, err, sizeof err, &kind);
	if (!A)
	{
		std::fprintf(stderr, "error: %s\n", err);
		return 1;
	}
	const uint64_t K = dyn_num_kmers(A);
	std::vector<double> pw, px, pxx, pxi(2, 0.0);
	if (o.train)
	{
		pw.assign(K, 0.0);
		px.assign(K, 0.0);
		pxx.assign(K, 0.0);
	}

	std::ios::sync_with_stdio(false);
	Batch b;
	std::string sline, qline, outbuf;
	int rc = 0;

	auto flush = [&]() {
		if (!b.n) return;
		if (b.sig.empty()) b.sig.push_back(0.0f);
		if (b.seq.empty()) b.seq.push_back('A');
		outbuf.clear();
		char num[64];
		if (!o.train)
		{
			std::vector<dyn_read_result> res(b.n);
			const uint64_t nseg = dyn_count_segments(A, b.seq_off.data(), b.n);
			std::vector<uint64_t> seqpos(nseg + 1), sigpos(nseg + 1);
			std::vector<double> prob(nseg + 1);
			if (dyn_align_batch(A, b.sig.data(), b.sig_off.data(), b.seq.data(), b.seq_off.data(), b.n, o.prob ? 1 : 0,
					res.data(), seqpos.data(), sigpos.data(), prob.data()) != 0)
			{
				std::fprintf(stderr, "error: %s\n", dyn_last_error(A));
				rc = 1;
				return;
			}
			for (uint32_t r = 0; r < b.n; ++r)
			{
				if (res[r].status != DYN_OK)
				{
					outbuf += "error:" + status_text(res[r].status, res[r].bad_char) + "\n";
					continue;
				}
				for (uint64_t i = 0; i < res[r].n_segments; ++i)
				{
					const uint64_t j = res[r].seg_offset + i;
					std::snprintf(num, sizeof num, "M%llu,%llu,%.6f;", (unsigned long long)seqpos[j],
						(unsigned long long)sigpos[j], prob[j]);
					outbuf += num;
				}
				if (o.prob) outbuf += "\t";
				std::snprintf(num, sizeof num, "Z:%.10g\n", res[r].Z);
				outbuf += num;
			}
		}
		else
		{
			std::vector<dyn_train_result> res(b.n);
			if (dyn_train_batch(A, b.sig.data(), b.sig_off.data(), b.seq.data(), b.seq_off.data(), b.n, res.data(),
					pw.data(), px.data(), pxx.data(), pxi.data(), nullptr, nullptr) != 0)
			{
				std::fprintf(stderr, "error: %s\n", dyn_last_error(A));
				rc = 1;
				return;
			}
			for (uint32_t r = 0; r < b.n; ++r)
			{
				if (res[r].status != DYN_OK)
				{
					outbuf += "error:" + status_text(res[r].status, res[r].bad_char) + "\n";
					continue;
				}
				std::snprintf(num, sizeof num, "m1:%.10g;e1:%.10g;e2:%.10g\tZ:%.10g\n", res[r].m1, res[r].e1, res[r].e2, res[r].Z);
				outbuf += num;
			}
		}
		std::fwrite(outbuf.data(), 1, outbuf.size(), stdout);
		std::fflush(stdout);
		b.clear();
	};

	while (std::getline(std::cin, sline))
	{
		if (!std::getline(std::cin, qline)) break;
		while (!qline.empty() && (qline.back() == '\r' || qline.back() == ' ')) qline.pop_back();
		parse_signal(sline, b.sig);
		b.sig_off.push_back(b.sig.size());
		b.seq += qline;
		b.seq_off.push_back(b.seq.size());
		if (++b.n >= (uint32_t)o.batch) flush();
		if (rc) break;
	}
	flush();

	if (o.train && !o.out_model.empty() && rc == 0)
	{
		// pooled M-step (NT_aligner_api.cpp:519-535 applied to the pooled statistics), written as a model TSV with
		// the kmers in the aligner's native index order translated back to file orientation
		std::vector<double> mean(K), sd(K);
		dyn_model(A, mean.data(), sd.data());
		const int k = dyn_kmer_size(A);
		const bool rna = dyn_is_rna(A) != 0;
		FILE* f = std::fopen(o.out_model.c_str(), "w");
		if (!f)
		{
			std::fprintf(stderr, "error: cannot write %s\n", o.out_model.c_str());
			rc = 1;
		}
		else
		{
			std::fprintf(f, "kmer\tlevel_mean\tlevel_stdv\n");
			// file order = lexicographic over the file-orientation kmer
			for (uint64_t v = 0; v < K; ++v)
			{
				std::string kmer(k, 'A');
				uint64_t x = v;
				for (int i = k - 1; i >= 0; --i)
				{
					kmer[i] = "ACGT"[x & 3];
					x >>= 2;
				}
				std::string native = kmer;
				if (rna) native.assign(kmer.rbegin(), kmer.rend());
				uint64_t q = 0;
				for (char c : native) q = q * 4 + (c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : 3);
				double m = mean[q], s = sd[q];
				if (pw[q] > 0.0)
				{
					m = px[q] / pw[q];
					double var = pxx[q] / pw[q] - m * m;
					if (var < 1e-12) var = 1e-12;
					s = std::sqrt(var);
				}
				std::fprintf(f, "%s\t%.17g\t%.17g\n", kmer.c_str(), m, s);
			}
			std::fclose(f);
			const double norm = pxi[0] + pxi[1];
			if (norm > 0) std::fprintf(stderr, "pooled transitions: m1=%.10g e2=%.10g\n", pxi[0] / norm, pxi[1] / norm);
		}
	}
	dyn_destroy(A);
	return rc;
}
