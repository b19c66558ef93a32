/* TEST INFRASTRUCTURE — CPU oracle, not part of the product path.
 *
 * Plain-C, IEEE-double restatement of the reference's basic-mode ("NT") hot path:
 * banded 2-state log-space forward/backward, posterior, posterior-Viterbi ("MAP")
 * fill + traceback with per-segment median, and the per-read EM statistics.
 * Every function cites the reference file:line it follows (paths relative to
 * /root/reference). It is written in (t, n) lattice coordinates with an explicit
 * per-row window instead of the reference's shifted band storage, but performs the
 * same floating-point operations in the same order, so results are bit-identical
 * (pinned against the compiled reference by tests/test_oracle.py and
 * against tests/golden/*.npz which were produced by the compiled reference).
 *
 * PARITY PIN: the reference's own tests hold no golden vector for this path
 * (SURVEY.md F6); the pin is the unmodified reference compiled into oracle/_ref/.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may use this.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define NEG_INF (-INFINITY)

/* aligner.cpp:287-292 */
static double log_normal_pdf(double x, double mean, double stdev)
{
	const double diff = x - mean;
	const double z = diff / stdev;
	return -0.5 * z * z - log(stdev) - 0.5 * log(2.0 * M_PI);
}

/* aligner.cpp:276-285 */
static double log_plus(double x, double y)
{
	if (isinf(x)) return y;
	if (isinf(y)) return x;
	if (x < y) { double tmp = x; x = y; y = tmp; }
	return x + log1p(exp(y - x));
}

static int cmp_double(const void* a, const void* b)
{
	const double x = *(const double*)a, y = *(const double*)b;
	return (x > y) - (x < y);
}

/* aligner.cpp:247-263 (nth_element replaced by a full sort: same order statistics) */
static double median_of(double* v, size_t n)
{
	if (n == 0) return 0.0;
	qsort(v, n, sizeof(double), cmp_double);
	if (n % 2 == 1) return v[n / 2];
	return (v[n / 2 - 1] + v[n / 2]) / 2.0;
}

typedef struct
{
	size_t T, N, bw, W; /* W = 2*bw+1 stored columns per row */
	long* lo;           /* lo[t] = mid_t - bw (may be negative): column 0 of row t */
	size_t* nStart;     /* NT:102 */
	size_t* nEnd;       /* NT:103 */
} band_t;

/* NT:90-108 computeBounds */
static int band_init(band_t* b, size_t T, size_t N, size_t band_half)
{
	b->T = T;
	b->N = N;
	b->bw = band_half < N / 2 ? band_half : N / 2; /* NT:243 */
	b->W = 2 * b->bw + 1;
	b->lo = (long*)malloc(T * sizeof(long));
	b->nStart = (size_t*)malloc(T * sizeof(size_t));
	b->nEnd = (size_t*)malloc(T * sizeof(size_t));
	if (!b->lo || !b->nStart || !b->nEnd) return 1;
	const double ratio = (double)N / (double)T;
	for (size_t t = 0; t < T; ++t)
	{
		const size_t mid = (size_t)((double)t * ratio);
		b->lo[t] = (long)mid - (long)b->bw;
		b->nStart[t] = mid >= b->bw ? mid - b->bw : 0;
		b->nEnd[t] = mid + b->bw + 1 <= N ? mid + b->bw + 1 : N;
	}
	return 0;
}

static void band_free(band_t* b)
{
	free(b->lo);
	free(b->nStart);
	free(b->nEnd);
}

/* value of lattice cell (t, n); anything outside row t's window is -inf (NT:256-262) */
static inline double cell(const band_t* b, const double* A, size_t t, long n)
{
	const long c = n - b->lo[t];
	if (c < 0 || c >= (long)b->W) return NEG_INF;
	return A[t * b->W + (size_t)c];
}
static inline double* cellp(const band_t* b, double* A, size_t t, size_t n)
{
	return &A[t * b->W + (size_t)((long)n - b->lo[t])];
}

static double* alloc_neg_inf(size_t count)
{
	double* p = (double*)malloc(count * sizeof(double));
	if (p)
		for (size_t i = 0; i < count; ++i) p[i] = NEG_INF;
	return p;
}

typedef struct
{
	const double* x;
	const int* kmer;
	const double* mean;
	const double* stdev;
	double m1, e1, e2;
} hmm_t;

static inline double score(const hmm_t* h, size_t t, size_t j) /* aligner.cpp:241-245 */
{
	const int q = h->kmer[j];
	return log_normal_pdf(h->x[t], h->mean[q], h->stdev[q]);
}

/* NT:110-152 */
static void forward(const hmm_t* h, const band_t* b, double* fM, double* fE)
{
	*cellp(b, fE, 0, 0) = 0.0;
	for (size_t t = 1; t < b->T; ++t)
	{
		size_t n0 = b->nStart[t] ? b->nStart[t] : 1;
		for (size_t n = n0; n < b->nEnd[t]; ++n)
		{
			const double s = score(h, t - 1, n - 1);
			*cellp(b, fM, t, n) = cell(b, fE, t - 1, (long)n - 1) + s + h->m1;
			*cellp(b, fE, t, n) = log_plus(cell(b, fM, t - 1, (long)n) + s + h->e1,
				cell(b, fE, t - 1, (long)n) + s + h->e2);
		}
	}
}

/* NT:158-207 */
static void backward(const hmm_t* h, const band_t* b, double* bM, double* bE)
{
	const size_t T = b->T, N = b->N;
	*cellp(b, bE, T - 1, N - 1) = 0.0;
	for (size_t t = T - 1; t-- > 0;)
	{
		for (size_t n = b->nStart[t]; n < b->nEnd[t]; ++n)
		{
			double ext = NEG_INF;
			if (n + 1 < N)
				ext = cell(b, bM, t + 1, (long)n + 1) + score(h, t, n) + h->m1;
			if (n > 0)
			{
				const double s = score(h, t, n - 1);
				const double e_next = cell(b, bE, t + 1, (long)n);
				*cellp(b, bM, t, n) = e_next + s;
				ext = log_plus(ext, e_next + s + h->e2);
			}
			*cellp(b, bE, t, n) = ext;
		}
	}
}

/* In-band DP cells per read: forward's trip count. */
uint64_t nt_oracle_cells(size_t S, size_t Kc, size_t band)
{
	band_t b;
	if (band_init(&b, S + 1, Kc + 1, band / 2)) return 0;
	uint64_t c = 0;
	for (size_t t = 1; t < b.T; ++t)
	{
		size_t n0 = b.nStart[t] ? b.nStart[t] : 1;
		if (b.nEnd[t] > n0) c += b.nEnd[t] - n0;
	}
	band_free(&b);
	return c;
}

/* NTAligner::align (NT:230-312) + calculateSegments (NT:314-377) + decodeMAP (NT:383-456).
 * kmers: Kc native kmer indices (aligner.cpp:166-205), mean/stdev: native-order model table.
 * trans = {m1,e1,e2} in log space (NT:84-86). Outputs seqpos/sigpos/prob hold Kc entries.
 * Optional dumps (may be NULL): rows_t/n_rows/rows_out as in ref_nt_stages.
 * Returns 0 ok, 1 "Alignment failed: alignment scores do not match" (NT:289-291), 2 out of memory. */
int nt_oracle_align(const double* x, size_t S, const int* kmers, size_t Kc, int k,
	const double* mean, const double* stdev, const double* trans, size_t band, int calc_prob,
	double* Zf_out, double* Zb_out, uint64_t* seqpos, uint64_t* sigpos, double* prob,
	const size_t* rows_t, size_t n_rows, double* rows_out)
{
	const size_t T = S + 1, N = Kc + 1;
	band_t b;
	if (band_init(&b, T, N, band / 2)) return 2;
	hmm_t h = {x, kmers, mean, stdev, trans[0], trans[1], trans[2]};
	const size_t size = T * b.W;
	double* fM = alloc_neg_inf(size);
	double* fE = alloc_neg_inf(size);
	double* bM = alloc_neg_inf(size);
	double* bE = alloc_neg_inf(size);
	double *VM = NULL, *VE = NULL, *buf = NULL;
	int rc = 0;
	if (!fM || !fE || !bM || !bE) { rc = 2; goto done; }

	forward(&h, &b, fM, fE);
	backward(&h, &b, bM, bE);

	const double Zf = cell(&b, fE, T - 1, (long)N - 1); /* NT:285 */
	const double Zb = cell(&b, bE, 0, 0);               /* NT:286 */
	*Zf_out = Zf;
	*Zb_out = Zb;
	for (size_t i = 0; i < n_rows; ++i)
	{
		const size_t t = rows_t[i];
		const double* src[4] = {fM, fE, bM, bE};
		for (int m = 0; m < 4; ++m)
			for (size_t n = 0; n < N; ++n)
				rows_out[(i * 4 + m) * N + n] =
					(n >= b.nStart[t] && n < b.nEnd[t]) ? cell(&b, src[m], t, (long)n) : NEG_INF;
	}
	/* the reference divides by T*B with B = 2*bw+3 (NT:244-245,288-290) */
	if (isinf(Zf) || isinf(Zb) || fabs(Zf - Zb) / (double)(T * (2 * b.bw + 3)) > 1e-8) { rc = 1; goto done; }
	if (!calc_prob) goto done;

	/* NT:213-224: LP = f + b - Zb, in place (cells never written stay -inf; -inf + -inf - Z = -inf) */
	for (size_t i = 0; i < size; ++i)
	{
		fM[i] = fM[i] + bM[i] - Zb;
		fE[i] = fE[i] + bE[i] - Zb;
	}
	const double *LPM = fM, *LPE = fE;

	/* NT:331-363 posterior-Viterbi fill */
	VM = alloc_neg_inf(size);
	VE = alloc_neg_inf(size);
	buf = (double*)malloc(T * sizeof(double));
	if (!VM || !VE || !buf) { rc = 2; goto done; }
	*cellp(&b, VE, 0, 0) = 0.0;
	for (size_t t = 1; t < T; ++t)
	{
		size_t n0 = b.nStart[t] ? b.nStart[t] : 1;
		for (size_t n = n0; n < b.nEnd[t]; ++n)
		{
			*cellp(&b, VM, t, n) = cell(&b, VE, t - 1, (long)n - 1) + cell(&b, LPM, t, (long)n);
			const double pm = cell(&b, VM, t - 1, (long)n), pe = cell(&b, VE, t - 1, (long)n);
			*cellp(&b, VE, t, n) = (pm < pe ? pe : pm) + cell(&b, LPE, t, (long)n); /* std::max, NT:361 */
		}
	}

	/* NT:383-456 traceback */
	{
		size_t t = T - 1, n = N - 1, nbuf = 0, nseg = 0;
		int inM = 0;
		while (t && n)
		{
			if (inM)
			{
				buf[nbuf++] = exp(cell(&b, LPM, t, (long)n));
				seqpos[nseg] = n - 1 + (size_t)k / 2;
				sigpos[nseg] = t - 1;
				prob[nseg] = median_of(buf, nbuf);
				++nseg;
				nbuf = 0;
				--t;
				--n;
				inM = 0;
			}
			else
			{
				const double lp = cell(&b, LPE, t, (long)n);
				buf[nbuf++] = exp(lp);
				inM = (cell(&b, VE, t, (long)n) == cell(&b, VM, t - 1, (long)n) + lp); /* NT:448 */
				--t;
			}
		}
		/* NT:455 reverse */
		for (size_t i = 0; i < nseg / 2; ++i)
		{
			uint64_t a = seqpos[i]; seqpos[i] = seqpos[nseg - 1 - i]; seqpos[nseg - 1 - i] = a;
			a = sigpos[i]; sigpos[i] = sigpos[nseg - 1 - i]; sigpos[nseg - 1 - i] = a;
			double p = prob[i]; prob[i] = prob[nseg - 1 - i]; prob[nseg - 1 - i] = p;
		}
		if (nseg != Kc) rc = 3; /* cannot happen for a finite Z; guards the caller's buffers */
	}

done:
	free(fM); free(fE); free(bM); free(bE); free(VM); free(VE); free(buf);
	band_free(&b);
	return rc;
}

/* NTAligner::train (NT:567-639), runTraining (NT:462-561), trainTransition (NT:641-725).
 * Outputs: Z; trans_out = {m1', e1', e2'} linear; w/sx/sxx = raw per-kmer sufficient statistics (K each,
 * the pooled-training restatement sums these over reads); new_mean/new_stdev = per-read
 * M-step (K each). Returns 0 ok, 1 "Training failed: alignment scores do not match", 2 oom. */
int nt_oracle_train(const double* x, size_t S, const int* kmers, size_t Kc, size_t K,
	const double* mean, const double* stdev, const double* trans, size_t band,
	double* Z_out, double* trans_out, double* w, double* sx, double* sxx,
	double* new_mean, double* new_stdev, double* xi_out)
{
	const size_t T = S + 1, N = Kc + 1;
	band_t b;
	if (band_init(&b, T, N, band / 2)) return 2;
	hmm_t h = {x, kmers, mean, stdev, trans[0], trans[1], trans[2]};
	const size_t size = T * b.W;
	double* fM = alloc_neg_inf(size);
	double* fE = alloc_neg_inf(size);
	double* bM = alloc_neg_inf(size);
	double* bE = alloc_neg_inf(size);
	int rc = 0;
	if (!fM || !fE || !bM || !bE) { rc = 2; goto done; }
	forward(&h, &b, fM, fE);
	backward(&h, &b, bM, bE);
	const double Zf = cell(&b, fE, T - 1, (long)N - 1);
	const double Z = cell(&b, bE, 0, 0);
	*Z_out = Z;
	if (isinf(Zf) || isinf(Z) || fabs(Zf - Z) / (double)(T * (2 * b.bw + 3)) > 1e-8) { rc = 1; goto done; }

	for (size_t q = 0; q < K; ++q) w[q] = sx[q] = sxx[q] = 0.0;
	for (size_t t = 1; t < T; ++t) /* NT:494-514 */
	{
		size_t n0 = b.nStart[t] ? b.nStart[t] : 1;
		for (size_t n = n0; n < b.nEnd[t]; ++n)
		{
			const double post = exp(cell(&b, fM, t, (long)n) + cell(&b, bM, t, (long)n) - Z) +
				exp(cell(&b, fE, t, (long)n) + cell(&b, bE, t, (long)n) - Z);
			const int q = kmers[n - 1];
			const double obs = x[t - 1];
			w[q] += post;
			sx[q] += post * obs;
			sxx[q] += post * obs * obs;
		}
	}
	for (size_t q = 0; q < K; ++q) /* NT:519-535 */
	{
		if (w[q] > 0.0)
		{
			const double mu = sx[q] / w[q];
			double var = sxx[q] / w[q] - mu * mu;
			if (var < 1e-12) var = 1e-12;
			new_mean[q] = mu;
			new_stdev[q] = sqrt(var);
		}
		else
		{
			new_mean[q] = mean[q];
			new_stdev[q] = stdev[q];
		}
	}
	{ /* NT:641-725 */
		double newM1 = NEG_INF, newE2 = NEG_INF;
		for (size_t t = T - 1; t-- > 0;)
			for (size_t n = b.nStart[t]; n < b.nEnd[t]; ++n)
			{
				const double fe = cell(&b, fE, t, (long)n);
				if (n + 1 < N)
					newM1 = log_plus(newM1, fe + h.m1 + score(&h, t, n) + cell(&b, bM, t + 1, (long)n + 1));
				if (n > 0)
					newE2 = log_plus(newE2, fe + h.e2 + score(&h, t, n - 1) + cell(&b, bE, t + 1, (long)n));
			}
		const double norm = log_plus(newM1, newE2);
		if (!isinf(norm))
		{
			trans_out[0] = exp(newM1 - norm);
			trans_out[2] = exp(newE2 - norm);
		}
		else
			trans_out[0] = trans_out[2] = 0.0;
		trans_out[1] = exp(h.e1);
		if (xi_out) /* expected transition counts, Z-normalised (used by the pooled restatement) */
		{
			xi_out[0] = exp(newM1 - Z);
			xi_out[1] = exp(newE2 - Z);
		}
	}
done:
	free(fM); free(fE); free(bM); free(bE);
	band_free(&b);
	return rc;
}
