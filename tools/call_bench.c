/* tools/call_bench.c -- latency and concurrency of the one-pair drop-in call from C, for any implementation of ssw.h.
 *   call_bench LIB.so n_calls threads ref_len read_len flag [option=value ...]
 * (options: ssw_engine_set_option(NULL, ...) of our library, e.g. tb_spec=0; ignored for libraries without that symbol)
 * Every call is ssw_init + ssw_align + align_destroy + init_destroy on the same (read, reference) pair, as a legacy
 * caller looping over reads does (main.c:462-532).  Prints microseconds per call for one thread and the wall time of
 * the same number of calls spread over `threads` pthreads (the reference is re-entrant; so is our engine pool). */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <time.h>

typedef struct { uint16_t score1, score2; int32_t rb, re, qb, qe, re2; uint32_t* cigar; int32_t cigarLen; uint16_t flag; } s_align;
typedef void* (*fn_init)(const int8_t*, int32_t, const int8_t*, int32_t, int8_t);
typedef void (*fn_idestroy)(void*);
typedef s_align* (*fn_align)(const void*, const int8_t*, int32_t, uint8_t, uint8_t, uint8_t, uint16_t, int32_t, int32_t);
typedef void (*fn_adestroy)(s_align*);
static fn_init p_init; static fn_idestroy p_idestroy; static fn_align p_align; static fn_adestroy p_adestroy;
static int8_t *g_ref, *g_read, g_mat[25];
static int g_ref_len, g_read_len, g_flag, g_calls;
static long g_sum;

static double now(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }
static void* work(void* arg)
{
	long sum = 0;
	for (int i = 0; i < g_calls; ++i) {
		void* p = p_init(g_read, g_read_len, g_mat, 5, 2);
		s_align* a = p_align(p, g_ref, g_ref_len, 3, 1, (uint8_t)g_flag, 0, 32767, g_read_len / 2);
		if (a) { sum += a->score1; p_adestroy(a); }
		p_idestroy(p);
	}
	__atomic_fetch_add(&g_sum, sum, __ATOMIC_RELAXED);
	return arg;
}

int main(int argc, char** argv)
{
	if (argc < 7) { fprintf(stderr, "usage: %s LIB.so n_calls threads ref_len read_len flag\n", argv[0]); return 2; }
	void* h = dlopen(argv[1], RTLD_NOW);
	if (!h) { fprintf(stderr, "%s\n", dlerror()); return 1; }
	p_init = (fn_init)dlsym(h, "ssw_init"); p_idestroy = (fn_idestroy)dlsym(h, "init_destroy");
	p_align = (fn_align)dlsym(h, "ssw_align"); p_adestroy = (fn_adestroy)dlsym(h, "align_destroy");
	typedef int (*fn_opt)(void*, const char*, long long);
	fn_opt p_opt = (fn_opt)dlsym(h, "ssw_engine_set_option");
	for (int a = 7; a < argc && p_opt; ++a) {
		char name[64]; long long v = 0;
		if (sscanf(argv[a], "%63[^=]=%lld", name, &v) == 2) p_opt(NULL, name, v);
	}
	const int n = atoi(argv[2]), threads = atoi(argv[3]);
	g_ref_len = atoi(argv[4]); g_read_len = atoi(argv[5]); g_flag = atoi(argv[6]);
	g_ref = (int8_t*)malloc((size_t)g_ref_len); g_read = (int8_t*)malloc((size_t)g_read_len);
	uint32_t x = 12345;
	for (int i = 0; i < g_ref_len; ++i) { x = x * 1664525u + 1013904223u; g_ref[i] = (int8_t)((x >> 24) & 3); }
	/* the read: a copy of a stretch of the reference with ~5 % substitutions and ~2 % insertions / deletions */
	for (int i = 0, rp = g_ref_len / 3; i < g_read_len; ++i) {
		x = x * 1664525u + 1013904223u;
		const unsigned u = (x >> 12) % 100;
		if (u < 2 && g_read_len > 1000) { ++rp; }                                         /* deletion */
		if (u >= 2 && u < 4 && g_read_len > 1000) { g_read[i] = (int8_t)((x >> 24) & 3); continue; }   /* insertion */
		g_read[i] = (u < 9) ? (int8_t)((x >> 24) & 3) : g_ref[rp % g_ref_len];
		++rp;
	}
	for (int i = 0; i < 25; ++i) g_mat[i] = (i / 5 == 4 || i % 5 == 4) ? 0 : (i / 5 == i % 5 ? 2 : -2);
	/* warm-up: creates engines / scratch */
	g_calls = 8;
	pthread_t th[64];
	for (int t = 0; t < threads && t < 64; ++t) pthread_create(&th[t], NULL, work, NULL);
	for (int t = 0; t < threads && t < 64; ++t) pthread_join(th[t], NULL);
	g_calls = n;
	double t0 = now();
	work(NULL);
	const double serial = now() - t0;
	g_calls = n / threads;
	t0 = now();
	for (int t = 0; t < threads && t < 64; ++t) pthread_create(&th[t], NULL, work, NULL);
	for (int t = 0; t < threads && t < 64; ++t) pthread_join(th[t], NULL);
	const double conc = now() - t0;
	printf("{\"lib\": \"%s\", \"ref_len\": %d, \"read_len\": %d, \"flag\": %d, \"calls\": %d, \"us_per_call_one_thread\": %.1f, \"threads\": %d, "
	       "\"wall_ms_one_thread\": %.2f, \"wall_ms_threads\": %.2f, \"speedup\": %.2f, \"checksum\": %ld}\n",
	       argv[1], g_ref_len, g_read_len, g_flag, n, serial / n * 1e6, threads, serial * 1e3, conc * 1e3, serial / conc, g_sum);
	return 0;
}
