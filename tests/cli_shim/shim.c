/*
 * TEST INFRASTRUCTURE ONLY.  Stands in for libssw.so's batch entry points with the CPU oracle so that the
 * host-only logic of ssw_batch_cli (file parsing, option handling, BLAST/SAM formatting, strand choice) can be
 * checked against the frozen reference outputs on a machine without a GPU.  Never linked into the product.
 */
#include <stdlib.h>
#include <string.h>
#include "../../include/ssw.h"
#include "../../include/ssw_batch.h"
#include "../../oracle/ssw_oracle.h"

struct ssw_engine { int unused; };

ssw_engine* ssw_engine_create(int device) { (void)device; return (ssw_engine*)calloc(1, sizeof(struct ssw_engine)); }
void ssw_engine_destroy(ssw_engine* e) { free(e); }

int ssw_align_batch(ssw_engine* e, const ssw_batch_params* P,
                    int32_t n_queries, const int8_t* queries, const int64_t* query_off,
                    int32_t n_refs, const int8_t* refs, const int64_t* ref_off,
                    int64_t n_pairs, const int32_t* pair_query, const int32_t* pair_ref, s_align** out)
{
	(void)e; (void)n_queries;
	for (int64_t p = 0; p < n_pairs; ++p) {
		const int32_t q = pair_query ? pair_query[p] : (int32_t)(p / n_refs);
		const int32_t r = pair_ref ? pair_ref[p] : (int32_t)(p % n_refs);
		const int32_t qlen = (int32_t)(query_off[q + 1] - query_off[q]);
		oracle_profile* prof = oracle_ssw_init(queries + query_off[q], qlen, P->mat, P->n, P->score_size);
		const int32_t mask = P->mask_len < 0 ? qlen / 2 : P->mask_len;
		out[p] = (s_align*)oracle_ssw_align(prof, refs + ref_off[r], (int32_t)(ref_off[r + 1] - ref_off[r]),
		                                    P->gap_open, P->gap_extend, P->flag, P->filters, P->filterd, mask);
		oracle_init_destroy(prof);
	}
	return 0;
}

/* text variant: translate (and reverse-complement) on the host, then as above */
static int rc_letter(int c)
{
	switch (c) {
	case 'A': case 'a': return 'T';
	case 'C': case 'c': return 'G';
	case 'G': case 'g': return 'C';
	case 'T': case 't': case 'U': case 'u': return 'A';
	case 'N': case 'n': return 'N';
	default: return 4;
	}
}

int ssw_align_batch_text(ssw_engine* e, const ssw_batch_params* P, const int8_t* table, int32_t add_rc,
                         int32_t n_queries, const char* queries, const int64_t* query_off,
                         int32_t n_refs, const char* refs, const int64_t* ref_off,
                         int64_t n_pairs, const int32_t* pair_query, const int32_t* pair_ref, s_align** out)
{
	const int64_t qb = query_off[n_queries], rb = ref_off[n_refs];
	const int32_t nq = n_queries * (add_rc ? 2 : 1);
	int8_t* qc = (int8_t*)malloc((size_t)(qb * (add_rc ? 2 : 1)) + 1);
	int8_t* rcodes = (int8_t*)malloc((size_t)rb + 1);
	int64_t* qo = (int64_t*)malloc(sizeof(int64_t) * (size_t)(nq + 1));
	for (int64_t i = 0; i < qb; ++i) qc[i] = table[queries[i] & 127];
	for (int64_t i = 0; i < rb; ++i) rcodes[i] = table[refs[i] & 127];
	for (int32_t k = 0; k <= n_queries; ++k) qo[k] = query_off[k];
	if (add_rc)
		for (int32_t k = 0; k < n_queries; ++k) {
			const int64_t b = query_off[k], len = query_off[k + 1] - b;
			for (int64_t p = 0; p < len; ++p) qc[qb + b + p] = table[rc_letter(queries[b + len - 1 - p]) & 127];
			qo[n_queries + k + 1] = qb + query_off[k + 1];
		}
	const int rc = ssw_align_batch(e, P, nq, qc, qo, n_refs, rcodes, ref_off, n_pairs, pair_query, pair_ref, out);
	free(qc); free(rcodes); free(qo);
	return rc;
}

void align_destroy(s_align* a) { oracle_align_destroy((oracle_align_t*)a); }

/* ssw_align_batch + the oracle's mark_mismatch per record (the product does this step on the device) */
int ssw_align_batch_marked(ssw_engine* e, const ssw_batch_params* P,
                           int32_t n_queries, const int8_t* queries, const int64_t* query_off,
                           int32_t n_refs, const int8_t* refs, const int64_t* ref_off,
                           int64_t n_pairs, const int32_t* pair_query, const int32_t* pair_ref, s_align** out, int32_t* nm)
{
	const int rc = ssw_align_batch(e, P, n_queries, queries, query_off, n_refs, refs, ref_off, n_pairs, pair_query, pair_ref, out);
	if (rc) return rc;
	for (int64_t p = 0; p < n_pairs; ++p) {
		if (nm) nm[p] = 0;
		if (!out[p] || out[p]->cigarLen <= 0) continue;
		const int32_t q = pair_query ? pair_query[p] : (int32_t)(p / n_refs);
		const int32_t r = pair_ref ? pair_ref[p] : (int32_t)(p % n_refs);
		const int32_t v = oracle_mark_mismatch(out[p]->ref_begin1, out[p]->read_begin1, out[p]->read_end1, refs + ref_off[r], queries + query_off[q],
		                                       (int32_t)(query_off[q + 1] - query_off[q]), &out[p]->cigar, &out[p]->cigarLen);
		if (nm) nm[p] = v;
	}
	return 0;
}

int32_t mark_mismatch(int32_t ref_begin1, int32_t read_begin1, int32_t read_end1, const int8_t* ref, const int8_t* read,
                      int32_t readLen, uint32_t** cigar, int32_t* cigarLen)
{
	return oracle_mark_mismatch(ref_begin1, read_begin1, read_end1, ref, read, readLen, cigar, cigarLen);
}

/* ssw.h entry points for the C++ wrapper's host-logic test */
s_profile* ssw_init(const int8_t* read, const int32_t readLen, const int8_t* mat, const int32_t n, const int8_t score_size)
{
	return (s_profile*)oracle_ssw_init(read, readLen, mat, n, score_size);
}
void init_destroy(s_profile* p) { oracle_init_destroy((oracle_profile*)p); }
s_align* ssw_align(const s_profile* prof, const int8_t* ref, int32_t refLen, const uint8_t weight_gapO, const uint8_t weight_gapE,
                   const uint8_t flag, const uint16_t filters, const int32_t filterd, const int32_t maskLen)
{
	return (s_align*)oracle_ssw_align((const oracle_profile*)prof, ref, refLen, weight_gapO, weight_gapE, flag, filters, filterd, maskLen);
}

const uint8_t encoded_ops[128] = { ['M'] = 0, ['I'] = 1, ['D'] = 2, ['N'] = 3, ['S'] = 4, ['H'] = 5, ['P'] = 6, ['='] = 7, ['X'] = 8 };

/* device groups (ssw_batch.h): the front ends' -g N / devices != 1 paths; here one "device" answers for all of them */
struct ssw_group { int n; };
ssw_group* ssw_group_create(int32_t n_devices, const int32_t* devices) { (void)devices; ssw_group* g = (ssw_group*)calloc(1, sizeof(struct ssw_group)); g->n = n_devices > 0 ? n_devices : 1; return g; }
void ssw_group_destroy(ssw_group* g) { free(g); }
int32_t ssw_group_size(const ssw_group* g) { return g ? g->n : 0; }
int ssw_group_align_batch(ssw_group* g, const ssw_batch_params* P, const int8_t* table, int32_t add_rc,
                          int32_t n_queries, const void* queries, const int64_t* query_off,
                          int32_t n_refs, const void* refs, const int64_t* ref_off,
                          int64_t n_pairs, const int32_t* pair_query, const int32_t* pair_ref, s_align** out, int32_t marked, int32_t* nm)
{
	(void)g;
	if (table) return ssw_align_batch_text(NULL, P, table, add_rc, n_queries, (const char*)queries, query_off, n_refs, (const char*)refs, ref_off, n_pairs, pair_query, pair_ref, out);
	if (marked) return ssw_align_batch_marked(NULL, P, n_queries, (const int8_t*)queries, query_off, n_refs, (const int8_t*)refs, ref_off, n_pairs, pair_query, pair_ref, out, nm);
	return ssw_align_batch(NULL, P, n_queries, (const int8_t*)queries, query_off, n_refs, (const int8_t*)refs, ref_off, n_pairs, pair_query, pair_ref, out);
}
