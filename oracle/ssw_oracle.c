/*
 * oracle/ssw_oracle.c -- CPU restatement of the reference's ssw_init -> ssw_align path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (libssw.so, the package
 * under complete-striped-smith-waterman-library_b200/) may include, link or
 * call this file.  Allowed users: tests/, __graft_entry__.smoke(), and
 * bench.py's cpu_baseline / --impl reference legs (as the checker / the CPU
 * arm, never as the thing shipped).
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this restatement against
 * (a) the reference's golden outputs frozen under tests/golden/ and (b) the
 * unmodified reference ssw.c compiled into oracle/_ref/libssw_ref.so, on
 * randomised inputs in every parameter regime.
 *
 * Everything here is plain scalar C (no intrinsics).  Two formulations of the
 * matrix fill are given:
 *
 *   oracle_fill_striped()  literal lane-by-lane emulation of the reference's
 *       16-lane u8 / 8-lane i16 striped loops including the lazy-F early exit
 *       (ssw.c:197-386 and ssw.c:412-588).  Exact in every gap regime.
 *
 *   oracle_fill_gotoh()    the GPU-shaped restatement: plain affine-gap
 *       recurrence over real + pad rows producing colmax[] (pass A), then the
 *       order-dependent bookkeeping (pass B).  Equal to the striped form when
 *       gapO > gapE; it is the executable spec of the CUDA kernels.
 *
 * The exported oracle_ssw_* functions mirror ssw_init / ssw_align /
 * align_destroy / init_destroy / mark_mismatch one for one.
 */
#include <stdint.h>
#include <stdlib.h>
#include <stdio.h>
#include <string.h>

#include "ssw_oracle.h"

/* ------------------------------------------------------------------------- */
/* small helpers                                                             */
/* ------------------------------------------------------------------------- */

static inline int32_t imax(int32_t a, int32_t b) { return a > b ? a : b; }
static inline int32_t imin(int32_t a, int32_t b) { return a < b ? a : b; }

/* BAM op code of a CIGAR letter (reference table ssw.c:127-160). */
static uint32_t op_code(char c) {
	switch (c) {
	case 'M': return 0; case 'I': return 1; case 'D': return 2; case 'N': return 3;
	case 'S': return 4; case 'H': return 5; case 'P': return 6; case '=': return 7;
	case 'X': return 8; default: return 0;
	}
}
static inline uint32_t pack_cigar(uint32_t len, char op) { return (len << 4) | op_code(op); }
static inline char cigar_op(uint32_t w) { return (w & 0xfu) > 8 ? 'M' : "MIDNSHP=X"[w & 0xfu]; }
static inline uint32_t cigar_len(uint32_t w) { return w >> 4; }

/* lane arithmetic of the two SSE2 flavours */
static inline int32_t u8_adds(int32_t a, int32_t b) { int32_t s = a + b; return s > 255 ? 255 : s; }
static inline int32_t u_subs(int32_t a, int32_t b) { int32_t s = a - b; return s < 0 ? 0 : s; }
static inline int32_t i16_adds(int32_t a, int32_t b) {
	int32_t s = a + b; return s > 32767 ? 32767 : (s < -32768 ? -32768 : s);
}

/* ------------------------------------------------------------------------- */
/* Fill, formulation 1: literal striped emulation                            */
/* ------------------------------------------------------------------------- */

/*
 * Follows sw_sse2_byte (ssw.c:197-386) when word == 0 and sw_sse2_word
 * (ssw.c:412-588) when word == 1.  Storage is [segment][lane]; query row of
 * (segment s, lane l) is s + l*segLen (profile layout ssw.c:176-186 / :398-408).
 */
void oracle_fill_striped(const int8_t* ref, int32_t ref_dir, int32_t refLen,
                         const int8_t* read, int32_t readLen,
                         const int8_t* mat, int32_t n,
                         int32_t gapO, int32_t gapE,
                         int32_t word, int32_t terminate, int32_t bias, int32_t maskLen,
                         oracle_fill_t* out)
{
	const int32_t lanes = word ? 8 : 16;
	const int32_t segLen = (readLen + lanes - 1) / lanes;
	const int32_t cells = segLen * lanes;
	int32_t* prof = (int32_t*)malloc(sizeof(int32_t) * (size_t)n * cells);
	int32_t* Hs = (int32_t*)calloc(cells, sizeof(int32_t));   /* pvHStore */
	int32_t* Hl = (int32_t*)calloc(cells, sizeof(int32_t));   /* pvHLoad  */
	int32_t* E  = (int32_t*)calloc(cells, sizeof(int32_t));   /* pvE      */
	int32_t* Hm = (int32_t*)calloc(cells, sizeof(int32_t));   /* pvHmax   */
	int32_t* colmax = (int32_t*)calloc(refLen > 0 ? refLen : 1, sizeof(int32_t)); /* maxColumn */
	int32_t vH[16], vF[16], vMaxCol[16], vMaxScore[16], vMaxMark[16];
	int32_t max = 0, end_read = readLen - 1, end_ref = word ? 0 : -1;   /* :218-220, :427-429 */
	int32_t i, s, l, k;

	/* query profile: qP_byte ssw.c:163-188 (u8, +bias, pad = bias), qP_word :388-410 (i16, pad = 0) */
	for (int32_t nt = 0; nt < n; ++nt)
		for (s = 0; s < segLen; ++s)
			for (l = 0; l < lanes; ++l) {
				int32_t row = s + l * segLen;
				int32_t v;
				if (word) v = row >= readLen ? 0 : mat[nt * n + read[row]];
				else v = (uint8_t)(row >= readLen ? bias : mat[nt * n + read[row]] + bias);
				prof[(nt * segLen + s) * lanes + l] = v;
			}
	for (l = 0; l < lanes; ++l) vMaxScore[l] = vMaxMark[l] = 0;

	int32_t begin = 0, end = refLen, step = 1;
	if (ref_dir == 1) { begin = refLen - 1; end = -1; step = -1; }      /* :250-254 */

	for (i = begin; i != end; i += step) {
		const int32_t* vP = prof + (size_t)ref[i] * segLen * lanes;
		int32_t changed, cm;
		/* vH = last segment of the previous column shifted up one lane (:263-264, :470-471) */
		for (l = lanes - 1; l > 0; --l) vH[l] = Hs[(segLen - 1) * lanes + l - 1];
		vH[0] = 0;
		for (l = 0; l < lanes; ++l) { vF[l] = 0; vMaxCol[l] = 0; }
		{ int32_t* t = Hl; Hl = Hs; Hs = t; }                           /* swap :268-270 */

		for (s = 0; s < segLen; ++s) {                                  /* inner loop :274-299, :482-506 */
			for (l = 0; l < lanes; ++l) {
				int32_t h = vH[l], e = E[s * lanes + l], hg;
				if (word) h = i16_adds(h, vP[s * lanes + l]);
				else { h = u8_adds(h, vP[s * lanes + l]); h = u_subs(h, bias); }
				h = imax(h, e);
				h = imax(h, vF[l]);
				vMaxCol[l] = imax(vMaxCol[l], h);
				Hs[s * lanes + l] = h;
				hg = u_subs(h, gapO);
				e = imax(u_subs(e, gapE), hg);
				E[s * lanes + l] = e;
				vF[l] = imax(u_subs(vF[l], gapE), hg);
				vH[l] = Hl[s * lanes + l];
			}
		}

		/* lazy-F (:302-315, :509-520): up to `lanes` rounds, exit when no lane's
		 * F - gapE can still beat H - gapO */
		for (k = 0; k < lanes; ++k) {
			int32_t done = 0;
			for (l = lanes - 1; l > 0; --l) vF[l] = vF[l - 1];
			vF[0] = 0;
			for (s = 0; s < segLen; ++s) {
				int32_t all = 1;
				for (l = 0; l < lanes; ++l) {
					int32_t h = imax(Hs[s * lanes + l], vF[l]);
					vMaxCol[l] = imax(vMaxCol[l], h);
					Hs[s * lanes + l] = h;
					h = u_subs(h, gapO);
					vF[l] = u_subs(vF[l], gapE);
					if (vF[l] > h) all = 0;
				}
				if (all) { done = 1; break; }
			}
			if (done) break;
		}

		/* running maximum (:318-335, :523-537) */
		changed = 0;
		for (l = 0; l < lanes; ++l) {
			vMaxScore[l] = imax(vMaxScore[l], vMaxCol[l]);
			if (vMaxScore[l] != vMaxMark[l]) changed = 1;
		}
		if (changed) {
			int32_t temp = 0;
			for (l = 0; l < lanes; ++l) { vMaxMark[l] = vMaxScore[l]; temp = imax(temp, vMaxScore[l]); }
			if (temp > max) {
				max = temp;
				if (!word && max + bias >= 255) break;                  /* overflow :329 */
				end_ref = i;
				memcpy(Hm, Hs, sizeof(int32_t) * cells);
			}
		}
		cm = 0;
		for (l = 0; l < lanes; ++l) cm = imax(cm, vMaxCol[l]);
		colmax[i] = cm;                                                 /* :338, :540 */
		if (cm == terminate) break;                                     /* :339, :541 */
	}

	/* end position on the query (:342-351, :544-553) */
	for (i = 0; i < cells; ++i)
		if (Hm[i] == max) {
			int32_t row = i / lanes + (i % lanes) * segLen;
			if (row < end_read) end_read = row;
		}

	out->score = (!word && max + bias >= 255) ? 255 : max;              /* :360, :562 */
	out->ref = end_ref;
	out->read = end_read;
	out->score2 = 0;
	out->ref2 = 0;
	{   /* second best outside the mask window (:368-381, :570-583) */
		int32_t edge = (end_ref - maskLen) > 0 ? (end_ref - maskLen) : 0;
		for (i = 0; i < edge; ++i)
			if (colmax[i] > out->score2) { out->score2 = colmax[i]; out->ref2 = i; }
		edge = (end_ref + maskLen) > refLen ? refLen : (end_ref + maskLen);
		for (i = word ? edge : edge + 1; i < refLen; ++i)
			if (colmax[i] > out->score2) { out->score2 = colmax[i]; out->ref2 = i; }
	}
	free(prof); free(Hs); free(Hl); free(E); free(Hm); free(colmax);
}

/* ------------------------------------------------------------------------- */
/* Fill, formulation 2: Gotoh recurrence + ordered bookkeeping (GPU spec)    */
/* ------------------------------------------------------------------------- */

/*
 * Pass A: exact signed arithmetic over Lp = lanes*ceil(readLen/lanes) rows
 * (pad rows score 0 against every letter: ssw.c:182, :404), one column at a
 * time in scan order.  E is opened from max(0, diag, E) only (the reference
 * updates E before lazy-F: ssw.c:288-291); F is the exact serial gap down the
 * column.  Produces colmax[] and, per column, the smallest row holding it.
 * Pass B: the reference's strict-'>' running maximum, overflow stop, early
 * termination and mask-window scan.
 * Valid for gapO > gapE (for gapO <= gapE the reference's lazy-F exit makes
 * its result layout dependent; use oracle_fill_striped there).
 */
void oracle_fill_gotoh(const int8_t* ref, int32_t ref_dir, int32_t refLen,
                       const int8_t* read, int32_t readLen,
                       const int8_t* mat, int32_t n,
                       int32_t gapO, int32_t gapE,
                       int32_t word, int32_t terminate, int32_t bias, int32_t maskLen,
                       oracle_fill_t* out)
{
	const int32_t lanes = word ? 8 : 16;
	const int32_t Lp = lanes * ((readLen + lanes - 1) / lanes);
	int32_t* H = (int32_t*)calloc(Lp + 1, sizeof(int32_t));   /* H[j+1] = H(prev column, row j) */
	int32_t* E = (int32_t*)calloc(Lp, sizeof(int32_t));
	int32_t* colmax = (int32_t*)calloc(refLen > 0 ? refLen : 1, sizeof(int32_t));
	int32_t* minrow = (int32_t*)malloc(sizeof(int32_t) * (refLen > 0 ? refLen : 1));
	int32_t i, j, c;

	/* ---- pass A ---- */
	for (c = 0; c < refLen; ++c) {
		int32_t F = 0, diag = 0, cm = 0, mr = Lp;
		i = ref_dir == 1 ? refLen - 1 - c : c;
		for (j = 0; j < Lp; ++j) {
			int32_t s = j < readLen ? mat[ref[i] * n + read[j]] : 0;
			int32_t x = imax(imax(diag + s, E[j]), 0);        /* H before the F correction */
			int32_t h = imax(x, F);
			diag = H[j + 1];
			H[j + 1] = h;
			E[j] = imax(E[j] - gapE, x - gapO);
			F = imax(F - gapE, x - gapO);                      /* == max(F - gapE, h - gapO) for gapO >= gapE */
			if (h > cm) { cm = h; mr = j; }
		}
		colmax[i] = cm;
		minrow[i] = mr;
	}

	/* ---- pass B ---- */
	{
		int32_t max = 0, end_ref = word ? 0 : -1, end_read, overflow = 0;
		int32_t limit = word ? 0x7fffffff : 255 - bias;
		for (c = 0; c < refLen; ++c) {
			i = ref_dir == 1 ? refLen - 1 - c : c;
			if (colmax[i] > max) {
				max = colmax[i];
				if (max >= limit) { overflow = 1; break; }
				end_ref = i;
			}
			if (colmax[i] == terminate) { ++c; break; }
		}
		/* columns never visited read as 0 (calloc'd maxColumn in the reference) */
		for (; c < refLen; ++c) colmax[ref_dir == 1 ? refLen - 1 - c : c] = 0;
		/* ssw.c:343-351: smallest row < readLen-1 holding max in column end_ref,
		 * else readLen-1.  With max == 0 the zeroed snapshot matches everywhere. */
		end_read = readLen - 1;
		if (max == 0) end_read = 0;
		else if (!overflow && minrow[end_ref] < end_read) end_read = minrow[end_ref];
		out->score = overflow ? 255 : max;
		out->ref = end_ref;
		out->read = end_read;
		out->score2 = 0;
		out->ref2 = 0;
		{
			int32_t edge = (end_ref - maskLen) > 0 ? (end_ref - maskLen) : 0;
			for (i = 0; i < edge; ++i)
				if (colmax[i] > out->score2) { out->score2 = colmax[i]; out->ref2 = i; }
			edge = (end_ref + maskLen) > refLen ? refLen : (end_ref + maskLen);
			for (i = word ? edge : edge + 1; i < refLen; ++i)
				if (colmax[i] > out->score2) { out->score2 = colmax[i]; out->ref2 = i; }
		}
	}
	free(H); free(E); free(colmax); free(minrow);
}

/* ------------------------------------------------------------------------- */
/* Banded traceback (ssw.c:590-783) and CIGAR re-scoring (ssw.c:785-811)     */
/* ------------------------------------------------------------------------- */

typedef struct { uint32_t* seq; int32_t length; } o_cigar;

/* index of column j inside the rolling band row of query row i (macro set_u, ssw.c:92) */
static inline int32_t band_u(int32_t w, int32_t i, int32_t j) { int32_t x = i - w; if (x < 0) x = 0; return j - x + 1; }
/* index of direction byte p of cell (i, j) inside its direction row (macro set_d, ssw.c:95) */
static inline int32_t band_d(int32_t w, int32_t i, int32_t j, int32_t p) { int32_t x = i - w; if (x < 0) x = 0; return (j - x) * 3 + p; }

static void cig_push(uint32_t** c, int32_t* cap, int32_t idx, uint32_t v) {
	if (idx >= *cap) { while (idx >= *cap) *cap *= 2; *c = (uint32_t*)realloc(*c, sizeof(uint32_t) * (size_t)*cap); }
	(*c)[idx] = v;
}

static o_cigar* oracle_banded_sw(const int8_t* ref, const int8_t* read, int32_t refLen, int32_t readLen,
                                 int32_t score, uint32_t gapO, uint32_t gapE, int32_t band_width,
                                 const int8_t* mat, int32_t n)
{
	const int32_t neg_inf = INT32_MIN / 2;
	int32_t len = imax(refLen, readLen);
	int32_t max = 0, max_i = 0, max_j = 0, width, width_d, i, j;
	int32_t *h_b = NULL, *e_b = NULL, *h_c = NULL;
	int8_t* dir = NULL;

	do {
		width = band_width * 2 + 3;
		width_d = band_width * 2 + 1;
		h_b = (int32_t*)realloc(h_b, sizeof(int32_t) * (size_t)(width + 1));
		e_b = (int32_t*)realloc(e_b, sizeof(int32_t) * (size_t)(width + 1));
		h_c = (int32_t*)realloc(h_c, sizeof(int32_t) * (size_t)(width + 1));
		dir = (int8_t*)realloc(dir, (size_t)width_d * (size_t)readLen * 3 + 16);
		for (j = 1; j < width - 1; ++j) h_b[j] = 0;                              /* :627 */
		for (i = 0; i < readLen; ++i) {
			int32_t beg = imax(0, i - band_width), end = imin(refLen - 1, i + band_width);
			int32_t edge = imin(end + 1, width - 1), u = 0, f = neg_inf;         /* :629-637 */
			int8_t* dl = dir + (size_t)width_d * i * 3;
			h_b[0] = h_b[edge] = h_c[0] = 0;
			e_b[0] = e_b[edge] = neg_inf;
			for (j = beg; j <= end; ++j) {
				int32_t e, b, d, de, df, dh, t1, t2, e1, f1;
				u = band_u(band_width, i, j);
				e = band_u(band_width, i - 1, j);
				b = band_u(band_width, i, j - 1);
				d = band_u(band_width, i - 1, j - 1);
				de = band_d(band_width, i, j, 0);
				df = band_d(band_width, i, j, 1);
				dh = band_d(band_width, i, j, 2);
				/* gap penalties are uint32_t in the reference (:595-596): the
				 * subtraction wraps, which equals signed subtraction */
				t1 = i == 0 ? (int32_t)(0u - gapO) : (int32_t)((uint32_t)h_b[e] - gapO);
				t2 = i == 0 ? neg_inf : (int32_t)((uint32_t)e_b[e] - gapE);
				e_b[u] = t1 > t2 ? t1 : t2;
				dl[de] = t1 > t2 ? 3 : 2;
				t1 = (int32_t)((uint32_t)h_c[b] - gapO);
				t2 = (int32_t)((uint32_t)f - gapE);
				f = t1 > t2 ? t1 : t2;
				dl[df] = t1 > t2 ? 5 : 4;
				e1 = e_b[u] > 0 ? e_b[u] : 0;
				f1 = f > 0 ? f : 0;
				t1 = e1 > f1 ? e1 : f1;
				t2 = h_b[d] + mat[ref[j] * n + read[i]];
				h_c[u] = t1 > t2 ? t1 : t2;
				if (h_c[u] > max) { max = h_c[u]; max_i = i; max_j = j; }
				if (t1 <= t2) dl[dh] = 1;
				else dl[dh] = e1 > f1 ? dl[de] : dl[df];
			}
			for (j = 1; j <= u; ++j) h_b[j] = h_c[j];
		}
		band_width *= 2;
	} while (max < score && band_width <= len);                                   /* :678-679 */
	band_width /= 2;
	width_d = band_width * 2 + 1;

	/* traceback (:683-762) */
	{
		int32_t cap = 16, l = 0, e = 0, state = 2;
		uint32_t* c = (uint32_t*)malloc(sizeof(uint32_t) * cap);
		char op = 'M', prev = 'M';
		o_cigar* res;
		i = max_i; j = max_j;
		while (i >= 0 && j > 0) {
			int8_t code = dir[(size_t)width_d * i * 3 + band_d(band_width, i, j, state)];
			switch (code) {
			case 1: --i; --j; state = 2; op = 'M'; break;
			case 2: --i; state = 0; op = 'I'; break;
			case 3: --i; state = 2; op = 'I'; break;
			case 4: --j; state = 1; op = 'D'; break;
			case 5: --j; state = 2; op = 'D'; break;
			default:
				free(dir); free(h_c); free(e_b); free(h_b); free(c);
				return NULL;
			}
			if (op == prev) ++e;
			else { ++l; cig_push(&c, &cap, l - 1, pack_cigar((uint32_t)e, prev)); prev = op; e = 1; }
		}
		if (op == 'M') { ++l; cig_push(&c, &cap, l - 1, pack_cigar((uint32_t)(e + 1), op)); }
		else { l += 2; cig_push(&c, &cap, l - 1, pack_cigar(1, 'M')); c[l - 2] = pack_cigar((uint32_t)e, op); }
		res = (o_cigar*)malloc(sizeof(o_cigar));
		res->seq = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)(l > 0 ? l : 1));
		for (i = 0; i < l; ++i) res->seq[i] = c[l - 1 - i];                      /* reverse :765-775 */
		res->length = l;
		free(dir); free(h_c); free(e_b); free(h_b); free(c);
		return res;
	}
}

static int32_t oracle_cigar_score(const o_cigar* path, const int8_t* ref, const int8_t* read,
                                  const int8_t* mat, int32_t n, uint32_t gapO, uint32_t gapE)
{
	int32_t score = 0, rp = 0, qp = 0, i;
	for (i = 0; i < path->length; ++i) {
		uint32_t len = cigar_len(path->seq[i]), k;
		char op = cigar_op(path->seq[i]);
		if (op == 'M') {
			for (k = 0; k < len; ++k) { score += mat[ref[rp] * n + read[qp]]; ++rp; ++qp; }
		} else {
			score -= (int32_t)(gapO + (len > 1 ? (len - 1) * gapE : 0));          /* :804 */
			if (op == 'I') qp += len; else if (op == 'D') rp += len;
		}
	}
	return score;
}

/* ------------------------------------------------------------------------- */
/* ssw_init / ssw_align mirror (ssw.c:826-977)                               */
/* ------------------------------------------------------------------------- */

struct oracle_profile {
	const int8_t* read; const int8_t* mat;
	int32_t readLen, n, bias, has_byte, has_word;
};

static int g_use_gotoh = 0;          /* tests flip this to exercise formulation 2 end to end */
void oracle_set_formulation(int32_t gotoh) { g_use_gotoh = gotoh; }

oracle_profile* oracle_ssw_init(const int8_t* read, int32_t readLen, const int8_t* mat, int32_t n, int8_t score_size)
{
	oracle_profile* p = (oracle_profile*)calloc(1, sizeof(*p));
	if (score_size == 0 || score_size == 2) {
		int32_t b = 0, i;
		for (i = 0; i < n * n; ++i) if (mat[i] < b) b = mat[i];               /* :834-838 */
		p->bias = b < 0 ? -b : b;
		p->has_byte = 1;
	}
	if (score_size == 1 || score_size == 2) p->has_word = 1;
	p->read = read; p->mat = mat; p->readLen = readLen; p->n = n;
	return p;
}

void oracle_init_destroy(oracle_profile* p) { free(p); }

static void fill(const int8_t* ref, int32_t dir, int32_t refLen, const int8_t* read, int32_t readLen,
                 const oracle_profile* p, int32_t gapO, int32_t gapE, int32_t word, int32_t terminate,
                 int32_t maskLen, oracle_fill_t* out)
{
	if (g_use_gotoh && gapO > gapE)
		oracle_fill_gotoh(ref, dir, refLen, read, readLen, p->mat, p->n, gapO, gapE, word, terminate, p->bias, maskLen, out);
	else
		oracle_fill_striped(ref, dir, refLen, read, readLen, p->mat, p->n, gapO, gapE, word, terminate, p->bias, maskLen, out);
}

oracle_align_t* oracle_ssw_align(const oracle_profile* prof, const int8_t* ref, int32_t refLen,
                                 uint8_t gapO, uint8_t gapE, uint8_t flag,
                                 uint16_t filters, int32_t filterd, int32_t maskLen)
{
	oracle_fill_t best, rev;
	int32_t word = 0, readLen = prof->readLen;
	oracle_align_t* r = (oracle_align_t*)calloc(1, sizeof(*r));
	r->ref_begin1 = -1; r->read_begin1 = -1;

	if (prof->has_byte) {                                                     /* :881-899 */
		fill(ref, 0, refLen, prof->read, readLen, prof, gapO, gapE, 0, 255, maskLen, &best);
		if (best.score == 255) {
			if (!prof->has_word) { free(r); return NULL; }
			fill(ref, 0, refLen, prof->read, readLen, prof, gapO, gapE, 1, 65535, maskLen, &best);
			word = 1;
		}
	} else if (prof->has_word) {
		fill(ref, 0, refLen, prof->read, readLen, prof, gapO, gapE, 1, 65535, maskLen, &best);
		word = 1;
	} else { free(r); return NULL; }
	if (best.score <= 0) return r;                                            /* :900-903 */

	r->score1 = (uint16_t)best.score; r->ref_end1 = best.ref; r->read_end1 = best.read;
	if (maskLen >= 15) { r->score2 = (uint16_t)best.score2; r->ref_end2 = best.ref2; }
	else { r->score2 = 0; r->ref_end2 = -1; }
	if (flag == 0 || (flag == 2 && r->score1 < filters)) return r;            /* :916 */

	{   /* begin search on the reversed prefixes (:919-936) */
		int32_t qlen = r->read_end1 + 1, k;
		int8_t* rq = (int8_t*)malloc(qlen);
		for (k = 0; k < qlen; ++k) rq[k] = prof->read[qlen - 1 - k];
		fill(ref, 1, r->ref_end1 + 1, rq, qlen, prof, gapO, gapE, word, r->score1, maskLen, &rev);
		free(rq);
		r->ref_begin1 = rev.ref;
		r->read_begin1 = r->read_end1 - rev.read;
		if (r->score1 > rev.score) r->flag = 2;
	}

	if ((7 & flag) == 0 || ((2 & flag) != 0 && r->score1 < filters) ||
	    ((4 & flag) != 0 && (r->ref_end1 - r->ref_begin1 > filterd || r->read_end1 - r->read_begin1 > filterd)))
		return r;                                                             /* :938 */

	{   /* CIGAR (:941-973) */
		int32_t rl = r->ref_end1 - r->ref_begin1 + 1, ql = r->read_end1 - r->read_begin1 + 1;
		int32_t band = abs(rl - ql) + 1, full = imax(rl, ql);
		o_cigar* path;
		for (;;) {
			path = oracle_banded_sw(ref + r->ref_begin1, prof->read + r->read_begin1, rl, ql, r->score1,
			                        gapO, gapE, band, prof->mat, prof->n);
			if (!path) break;
			if (oracle_cigar_score(path, ref + r->ref_begin1, prof->read + r->read_begin1, prof->mat, prof->n, gapO, gapE) == r->score1) break;
			free(path->seq); free(path);
			if (band >= full) { path = NULL; break; }
			band = full;
		}
		if (!path) r->flag = 1;
		else { r->cigar = path->seq; r->cigarLen = path->length; free(path); }
	}
	return r;
}

void oracle_align_destroy(oracle_align_t* a) { if (a) { free(a->cigar); free(a); } }

/* ------------------------------------------------------------------------- */
/* mark_mismatch mirror (ssw.c:1019-1074)                                    */
/* ------------------------------------------------------------------------- */

int32_t oracle_mark_mismatch(int32_t ref_begin1, int32_t read_begin1, int32_t read_end1,
                             const int8_t* ref, const int8_t* read, int32_t readLen,
                             uint32_t** cigar, int32_t* cigarLen)
{
	int32_t cap = *cigarLen * 2 + 8, p = 0, nm = 0, i;
	uint32_t* out = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)cap);
	uint32_t run_eq = 0, run_x = 0;
	ref += ref_begin1; read += read_begin1;
	/* every flush appends at most one word; cap is generous: each input word
	 * yields at most len(=/X runs) <= 2*len words, grown on demand below */
#define EMIT(len_, op_) do { if (p >= cap) { cap *= 2; out = (uint32_t*)realloc(out, sizeof(uint32_t) * (size_t)cap); } out[p++] = pack_cigar((len_), (op_)); } while (0)
	if (read_begin1 > 0) EMIT((uint32_t)read_begin1, 'S');
	for (i = 0; i < *cigarLen; ++i) {
		char op = cigar_op((*cigar)[i]);
		int32_t len = (int32_t)cigar_len((*cigar)[i]), k;
		if (op == 'M') {
			for (k = 0; k < len; ++k) {
				if (*ref != *read) {
					++nm;
					if (run_eq) { EMIT(run_eq, '='); run_eq = 0; }
					++run_x;
				} else {
					/* store_previous_m, ssw.c:994-1009: at most one of the two runs is pending */
					if (run_x) { EMIT(run_x, 'X'); run_x = 0; }
					++run_eq;
				}
				++ref; ++read;
			}
		} else if (op == 'I' || op == 'D') {
			if (op == 'I') read += len; else ref += len;
			nm += len;
			if (run_eq) { EMIT(run_eq, '='); run_eq = 0; }
			else if (run_x) { EMIT(run_x, 'X'); run_x = 0; }
			EMIT((uint32_t)len, op);
		}
	}
	if (run_eq) { EMIT(run_eq, '='); run_eq = 0; }
	else if (run_x) { EMIT(run_x, 'X'); run_x = 0; }
	if (readLen - read_end1 - 1 > 0) EMIT((uint32_t)(readLen - read_end1 - 1), 'S');
#undef EMIT
	*cigarLen = p;
	free(*cigar);
	*cigar = out;
	return nm;
}
