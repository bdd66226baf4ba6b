/*
 * oracle/ssw_harness.c -- pthread driver for a CPU implementation of the ssw.h ABI.
 *
 * TEST INFRASTRUCTURE ONLY (same rule as ssw_oracle.c): used by tests/, by
 * __graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference /
 * parity self-check legs.  Nothing in the product links or loads it.
 *
 * It is the loop SURVEY 8(d) asks for: the reference's CLI aligns one pair per
 * blocking call in a double loop (src/main.c:462-532); the library is
 * re-entrant (no mutable globals), so independent pairs run on all host cores.
 * The implementation under test is loaded with dlopen(): oracle/_ref/libssw_ref.so
 * (the unmodified reference, symbols ssw_init ...) or oracle/libssw_oracle.so
 * (the scalar restatement, symbols oracle_ssw_init ...).  Per pair:
 * ssw_init(score_size) + ssw_align(...) + copy of the record (+ CIGAR words)
 * + align_destroy + init_destroy, exactly what main.c:477-529 does per pair.
 * Pairs are handed out through an atomic ticket (dynamic schedule).
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

/* s_align, ssw.h:55-66 (LP64: 40 bytes) */
typedef struct {
	uint16_t score1, score2;
	int32_t ref_begin1, ref_end1, read_begin1, read_end1, ref_end2;
	uint32_t* cigar;
	int32_t cigarLen;
	uint16_t flag;
} h_align;

/* same 36-byte layout as ssw_batch_result (include/ssw_batch.h), so both sides compare as byte records */
typedef struct {
	uint16_t score1, score2;
	int32_t ref_begin1, ref_end1, read_begin1, read_end1, ref_end2;
	int32_t cigar_off;      /* offset into the caller's pool, -1: none */
	int32_t cigar_len;
	uint16_t flag;
	uint16_t status;        /* 1: ssw_align returned NULL */
} harness_result;

typedef void* (*fn_init)(const int8_t*, int32_t, const int8_t*, int32_t, int8_t);
typedef void (*fn_init_destroy)(void*);
typedef h_align* (*fn_align)(const void*, const int8_t*, int32_t, uint8_t, uint8_t, uint8_t, uint16_t, int32_t, int32_t);
typedef void (*fn_align_destroy)(h_align*);

typedef struct {
	fn_init init; fn_init_destroy init_destroy; fn_align align; fn_align_destroy align_destroy;
	int64_t n_pairs;
	const int8_t* queries; const int64_t* q_off; const int32_t* pair_q;
	const int8_t* refs; const int64_t* r_off; const int32_t* pair_r;
	const int8_t* mat; int32_t n;
	int32_t gapO, gapE, flag, filters, filterd, mask_len, score_size;
	harness_result* out;
	uint32_t* pool; int64_t pool_cap;
	int64_t next;           /* ticket */
	int64_t pool_used;
	int64_t cells;
	int error;
} job_t;

static void* worker(void* arg)
{
	job_t* J = (job_t*)arg;
	int64_t cells = 0;
	for (;;) {
		const int64_t p = __atomic_fetch_add(&J->next, 1, __ATOMIC_RELAXED);
		if (p >= J->n_pairs) break;
		const int32_t q = J->pair_q ? J->pair_q[p] : (int32_t)p, r = J->pair_r ? J->pair_r[p] : 0;
		const int8_t* read = J->queries + J->q_off[q];
		const int32_t read_len = (int32_t)(J->q_off[q + 1] - J->q_off[q]);
		const int8_t* ref = J->refs + J->r_off[r];
		const int32_t ref_len = (int32_t)(J->r_off[r + 1] - J->r_off[r]);
		const int32_t mask = J->mask_len < 0 ? read_len / 2 : J->mask_len;       /* main.c:465 */
		void* prof = J->init(read, read_len, J->mat, J->n, (int8_t)J->score_size);
		h_align* a = J->align(prof, ref, ref_len, (uint8_t)J->gapO, (uint8_t)J->gapE, (uint8_t)J->flag, (uint16_t)J->filters, J->filterd, mask);
		harness_result* o = &J->out[p];
		memset(o, 0, sizeof(*o));
		o->cigar_off = -1;
		if (!a) o->status = 1;
		else {
			o->score1 = a->score1; o->score2 = a->score2; o->ref_begin1 = a->ref_begin1; o->ref_end1 = a->ref_end1;
			o->read_begin1 = a->read_begin1; o->read_end1 = a->read_end1; o->ref_end2 = a->ref_end2; o->flag = a->flag;
			if (a->cigarLen > 0 && a->cigar) {
				o->cigar_len = a->cigarLen;
				if (J->pool) {
					const int64_t off = __atomic_fetch_add(&J->pool_used, (int64_t)a->cigarLen, __ATOMIC_RELAXED);
					if (off + a->cigarLen <= J->pool_cap) {
						memcpy(J->pool + off, a->cigar, sizeof(uint32_t) * (size_t)a->cigarLen);
						o->cigar_off = (int32_t)off;
					} else J->error = 2;
				}
			}
			J->align_destroy(a);
		}
		J->init_destroy(prof);
		cells += (int64_t)read_len * ref_len;
	}
	__atomic_fetch_add(&J->cells, cells, __ATOMIC_RELAXED);
	return NULL;
}

/*
 * Run n_pairs alignments on n_threads threads with the implementation in `lib_path` (symbol prefix "" or "oracle_").
 * pair_q / pair_r NULL: pair p = (query p, reference 0).  pool may be NULL (CIGAR lengths are still recorded).
 * Returns 0 on success; *seconds = wall time of the threaded region, *cells = sum(readLen * refLen).
 */
int ssw_harness_run(const char* lib_path, const char* prefix, int32_t n_threads, int64_t n_pairs,
                    const int8_t* queries, const int64_t* q_off, const int32_t* pair_q,
                    const int8_t* refs, const int64_t* r_off, const int32_t* pair_r,
                    const int8_t* mat, int32_t n, int32_t gapO, int32_t gapE, int32_t flag, int32_t filters, int32_t filterd,
                    int32_t mask_len, int32_t score_size,
                    harness_result* out, uint32_t* pool, int64_t pool_cap, int64_t* pool_used,
                    double* seconds, int64_t* cells)
{
	void* h = dlopen(lib_path, RTLD_NOW | RTLD_LOCAL);
	if (!h) { fprintf(stderr, "ssw_harness: %s\n", dlerror()); return -1; }
	char name[128];
	job_t J;
	memset(&J, 0, sizeof(J));
	snprintf(name, sizeof(name), "%sssw_init", prefix ? prefix : ""); J.init = (fn_init)dlsym(h, name);
	snprintf(name, sizeof(name), "%sinit_destroy", prefix ? prefix : ""); J.init_destroy = (fn_init_destroy)dlsym(h, name);
	snprintf(name, sizeof(name), "%sssw_align", prefix ? prefix : ""); J.align = (fn_align)dlsym(h, name);
	snprintf(name, sizeof(name), "%salign_destroy", prefix ? prefix : ""); J.align_destroy = (fn_align_destroy)dlsym(h, name);
	if (!J.init || !J.init_destroy || !J.align || !J.align_destroy) { fprintf(stderr, "ssw_harness: missing symbols in %s\n", lib_path); dlclose(h); return -1; }
	J.n_pairs = n_pairs; J.queries = queries; J.q_off = q_off; J.pair_q = pair_q; J.refs = refs; J.r_off = r_off; J.pair_r = pair_r;
	J.mat = mat; J.n = n; J.gapO = gapO; J.gapE = gapE; J.flag = flag; J.filters = filters; J.filterd = filterd;
	J.mask_len = mask_len; J.score_size = score_size; J.out = out; J.pool = pool; J.pool_cap = pool_cap;
	if (n_threads < 1) n_threads = 1;
	if ((int64_t)n_threads > n_pairs) n_threads = n_pairs > 0 ? (int32_t)n_pairs : 1;
	pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)n_threads);
	struct timespec t0, t1;
	clock_gettime(CLOCK_MONOTONIC, &t0);
	int started = 0;
	for (int i = 1; i < n_threads; ++i) { if (pthread_create(&th[i], NULL, worker, &J) == 0) ++started; else th[i] = 0; }
	worker(&J);
	for (int i = 1; i < n_threads; ++i) if (th[i]) pthread_join(th[i], NULL);
	clock_gettime(CLOCK_MONOTONIC, &t1);
	free(th);
	(void)started;
	if (seconds) *seconds = (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
	if (cells) *cells = J.cells;
	if (pool_used) *pool_used = J.pool_used;
	dlclose(h);
	return J.error;
}
