/*
 * oracle/ssw_oracle.h -- interface of the CPU restatement (TEST INFRASTRUCTURE ONLY;
 * see the header of ssw_oracle.c for who may use it).
 */
#ifndef SSW_ORACLE_H
#define SSW_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* result of one matrix fill: the reference's alignment_end[2] (ssw.c:104-108) flattened */
typedef struct { int32_t score, ref, read, score2, ref2; } oracle_fill_t;

/* same field order and LP64 layout as s_align (ssw.h:55-66) */
typedef struct {
	uint16_t score1, score2;
	int32_t ref_begin1, ref_end1, read_begin1, read_end1, ref_end2;
	uint32_t* cigar;
	int32_t cigarLen;
	uint16_t flag;
} oracle_align_t;

typedef struct oracle_profile oracle_profile;

void oracle_fill_striped(const int8_t* ref, int32_t ref_dir, int32_t refLen, const int8_t* read, int32_t readLen,
                         const int8_t* mat, int32_t n, int32_t gapO, int32_t gapE,
                         int32_t word, int32_t terminate, int32_t bias, int32_t maskLen, oracle_fill_t* out);
void oracle_fill_gotoh(const int8_t* ref, int32_t ref_dir, int32_t refLen, const int8_t* read, int32_t readLen,
                       const int8_t* mat, int32_t n, int32_t gapO, int32_t gapE,
                       int32_t word, int32_t terminate, int32_t bias, int32_t maskLen, oracle_fill_t* out);

void oracle_set_formulation(int32_t gotoh);
oracle_profile* oracle_ssw_init(const int8_t* read, int32_t readLen, const int8_t* mat, int32_t n, int8_t score_size);
void oracle_init_destroy(oracle_profile* p);
oracle_align_t* oracle_ssw_align(const oracle_profile* prof, const int8_t* ref, int32_t refLen,
                                 uint8_t gapO, uint8_t gapE, uint8_t flag,
                                 uint16_t filters, int32_t filterd, int32_t maskLen);
void oracle_align_destroy(oracle_align_t* a);
int32_t oracle_mark_mismatch(int32_t ref_begin1, int32_t read_begin1, int32_t read_end1,
                             const int8_t* ref, const int8_t* read, int32_t readLen,
                             uint32_t** cigar, int32_t* cigarLen);
#ifdef __cplusplus
}
#endif
#endif
