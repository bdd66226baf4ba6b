/*
 * ssw_mark.cuh -- mark_mismatch (src/ssw.c:1019-1074) for a whole batch on the device: every M run of a CIGAR is split
 * into '=' and 'X' runs by comparing the codes, soft clips are added for the unaligned ends of the read, and the number
 * of mismatching + inserted + deleted bases (the NM tag the reference's SAM writer prints, main.c:226-231) is counted.
 * The sequences are the ones resident in the engine, the CIGARs the ones its traceback produced: nothing but the
 * (small) CIGAR pool crosses the bus.
 *
 * One warp per alignment.  The lanes compare 32 columns of an M run at a time; lane 0 run-length encodes the
 * equality mask (the run carries over 32-column groups and ends with the M word, as the reference's
 * store_previous_m does, ssw.c:994-1009).  Two passes: COUNT (length of the new CIGAR, NM) and WRITE.
 */
#ifndef SSW_MARK_CUH
#define SSW_MARK_CUH

#include "ssw_common.cuh"

#define SSW_MARK_THREADS 128

struct SswMarkTask {
	int64_t ref_off;        /* offset of ref[ref_begin1] in the padded reference array */
	int64_t read_off;       /* offset of read[0] in the query array */
	int64_t cig_off;        /* first word of the traceback's CIGAR in the input pool */
	int64_t out_off;        /* first word of the marked CIGAR in the output pool (set by the host between the passes) */
	int32_t read_len, read_begin1, read_end1;
	int32_t cig_len;
	int32_t out_len;        /* COUNT pass: words of the marked CIGAR */
	int32_t nm;             /* mismatches + inserted + deleted bases */
};

template <bool WRITE>
__global__ void __launch_bounds__(SSW_MARK_THREADS)
ssw_mark_kernel(SswMarkTask* __restrict__ tasks, int n_tasks, const int8_t* __restrict__ qcodes, const int8_t* __restrict__ refs,
                const uint32_t* __restrict__ cig_in, uint32_t* __restrict__ cig_out)
{
	constexpr unsigned FULL = 0xffffffffu;
	const int lane = threadIdx.x & 31;
	const int ti = (int)blockIdx.x * (SSW_MARK_THREADS / 32) + (threadIdx.x >> 5);
	if (ti >= n_tasks) return;
	const SswMarkTask T = tasks[ti];
	const int8_t* ref = refs + T.ref_off;
	const int8_t* read = qcodes + T.read_off + T.read_begin1;
	const uint32_t* cig = cig_in + T.cig_off;
	uint32_t* out = cig_out + T.out_off;
	int n_out = 0, nm = 0;                       /* meaningful in lane 0 */
	auto emit = [&](uint32_t len, uint32_t op) {  /* BAM op codes: I 1, D 2, S 4, = 7, X 8 (ssw.c:127-160) */
		if (WRITE) out[n_out] = (len << 4) | op;
		++n_out;
	};
	if (lane == 0 && T.read_begin1 > 0) emit((uint32_t)T.read_begin1, 4);
	int64_t rp = 0, qp = 0;
	for (int w = 0; w < T.cig_len; ++w) {
		const uint32_t word = cig[w];
		const int op = (int)(word & 15u), len = (int)(word >> 4);
		if (op == 0) {
			int run_type = -1, run_len = 0;       /* lane 0: 1 '=' run, 0 'X' run */
			for (int base = 0; base < len; base += 32) {
				const int k = base + lane;
				const bool valid = k < len;
				const bool eq = valid && ref[rp + k] == read[qp + k];
				const unsigned mask = __ballot_sync(FULL, eq);
				const int nv = min(32, len - base);
				if (lane == 0) {
					int pos = 0;
					while (pos < nv) {
						const int bit = (int)((mask >> pos) & 1u);
						/* length of the run of equal bits starting at pos */
						unsigned x = (bit ? ~mask : mask) >> pos;
						if (nv - pos < 32) x &= (1u << (nv - pos)) - 1u;
						const int run = x ? __ffs((int)x) - 1 : nv - pos;
						if (bit == run_type) run_len += run;
						else {
							if (run_len > 0) emit((uint32_t)run_len, run_type ? 7u : 8u);
							run_type = bit; run_len = run;
						}
						if (!bit) nm += run;
						pos += run;
					}
				}
			}
			if (lane == 0 && run_len > 0) emit((uint32_t)run_len, run_type ? 7u : 8u);
			rp += len; qp += len;
		} else if (op == 1) {
			if (lane == 0) { nm += len; emit((uint32_t)len, 1); }
			qp += len;
		} else if (op == 2) {
			if (lane == 0) { nm += len; emit((uint32_t)len, 2); }
			rp += len;
		}
	}
	if (lane == 0) {
		const int tail = T.read_len - T.read_end1 - 1;
		if (tail > 0) emit((uint32_t)tail, 4);
		if (!WRITE) { tasks[ti].out_len = n_out; tasks[ti].nm = nm; }
	}
}

#endif /* SSW_MARK_CUH */
