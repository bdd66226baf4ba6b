/*
 * example_batch.c -- the batched C API in one page (plain C99; build: gcc -Iinclude examples/example_batch.c -Lcomplete-striped-smith-waterman-library_b200 -lssw).
 *
 * The reference ships example.c: ONE read against ONE reference through ssw_init + ssw_align (its known answer: the
 * 15-mer CTGAGCCGGTAAATC against a 39-mer scores 21 at reference 8..21, read 0..14, second best 8 ending at 4, path
 * 4=1X4=1I5= with two edits).  Here the same pair is one of several that go to the GPU in one call: letters are handed over
 * as text with the CLI's translation table (main.c:72-93), the reverse complements are made on the device, and the paths come
 * back as mark_mismatch (ssw.h:157-164) would leave them.  With a second argument N the batch is cut over N GPUs (a device
 * group); the output does not change.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "ssw.h"
#include "ssw_batch.h"

int main(int argc, char** argv)
{
	const int devices = argc > 1 ? atoi(argv[1]) : 1;
	const char* refs[2] = {"CAGCCTTTCTGACCCGGAAATCAAAATAGGCACAACAAA", "TTTGTTGTGCCTATTTTGATTTCCGGGTCAGAAAGGCTG"};    /* the second is the reverse complement of the first */
	const char* reads[3] = {"CTGAGCCGGTAAATC", "GATTTACCGGCTCAG", "ACGTNNACGT"};
	char qtext[64] = "", rtext[96] = "";
	int64_t qoff[4] = {0}, roff[3] = {0};
	int8_t table[128], mat[25];
	int i, k;
	for (i = 0; i < 128; ++i) table[i] = 4;
	table['A'] = table['a'] = 0; table['C'] = table['c'] = 1; table['G'] = table['g'] = 2; table['T'] = table['t'] = table['U'] = table['u'] = 3;
	for (i = k = 0; i < 5; ++i) for (int j = 0; j < 5; ++j) mat[k++] = (i == 4 || j == 4) ? 0 : (i == j ? 2 : -2);
	for (i = 0; i < 3; ++i) { strcat(qtext, reads[i]); qoff[i + 1] = (int64_t)strlen(qtext); }
	for (i = 0; i < 2; ++i) { strcat(rtext, refs[i]); roff[i + 1] = (int64_t)strlen(rtext); }

	ssw_batch_params P;
	memset(&P, 0, sizeof P);
	P.mat = mat; P.n = 5; P.gap_open = 3; P.gap_extend = 1;
	P.flag = 1;                 /* example.c:147: begin positions and the path, always */
	P.mask_len = 15; P.score_size = 2;

	/* 3 reads + their 3 reverse complements, each against both references: 12 pairs, one call */
	enum { NQ = 6, NR = 2 };
	ssw_batch_result res[NQ * NR];
	uint32_t pool[NQ * NR * 64];
	int32_t nm[NQ * NR];
	int64_t used = 0;
	ssw_group* g = ssw_group_create(devices, NULL);
	if (!g) return 1;           /* no GPU: there is no CPU path */
	if (ssw_group_align(g, &P, table, 1, 3, qtext, qoff, NR, rtext, roff, NQ * NR, NULL, NULL, res, pool, NQ * NR * 64, &used, 1, nm)) return 1;
	for (i = 0; i < NQ * NR; ++i) {
		printf("read %d%s x ref %d: score %d (second %d at %d)  ref %d..%d  read %d..%d  NM %d  ", i / NR % 3, i / NR >= 3 ? "-" : "+", i % NR,
		       res[i].score1, res[i].score2, res[i].ref_end2, res[i].ref_begin1, res[i].ref_end1, res[i].read_begin1, res[i].read_end1, nm[i]);
		for (k = 0; k < res[i].cigar_len; ++k)
			printf("%u%c", cigar_int_to_len(pool[res[i].cigar_off + k]), cigar_int_to_op(pool[res[i].cigar_off + k]));
		printf("\n");
	}
	ssw_group_destroy(g);
	return 0;
}
