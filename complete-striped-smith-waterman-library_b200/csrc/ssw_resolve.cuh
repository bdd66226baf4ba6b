/*
 * ssw_resolve.cuh -- the order-dependent bookkeeping that follows a matrix fill
 * ("pass B" of SURVEY Appendix A.3).  Replaces, per alignment:
 *   - the running strict-'>' maximum / end_ref / overflow logic of the SSE2
 *     loops (src/ssw.c:318-335 byte, :523-537 word),
 *   - the end_read scan (:342-351, :544-553),
 *   - the second-best scan outside the mask window (:368-381, :570-583).
 * One warp per alignment.  Input: the per-item best cells written by the fill
 * kernel and the packed column-maximum row of the pair-task.  Items are in
 * scan order, so the first item holding the maximum also holds its first
 * column.  An item's best cell is at the same time (max, first index) of the
 * column maxima over its range, so the second-best scan reads column maxima
 * only for the few items that straddle the mask window.
 */
#ifndef SSW_RESOLVE_CUH
#define SSW_RESOLVE_CUH

#include "ssw_common.cuh"

#define SSW_RESOLVE_THREADS 128

/* candidate (v, i) beats (bv, bi): larger value, then smaller index; zero never counts */
__device__ static __forceinline__ bool ssw_second_better(int v, int i, int bv, int bi)
{
	return v > bv || (v == bv && v > 0 && i < bi);
}

template <bool SECOND>
__global__ void __launch_bounds__(SSW_RESOLVE_THREADS)
ssw_resolve_kernel(const SswAlnDesc* __restrict__ alns, int n_aln,
                   const SswItemBest* __restrict__ bests, const uint32_t* __restrict__ colmax,
                   SswFillResult* __restrict__ out)
{
	constexpr unsigned FULL = 0xffffffffu;
	const int lane = threadIdx.x & 31;
	const int idx = (int)blockIdx.x * (SSW_RESOLVE_THREADS / 32) + (threadIdx.x >> 5);
	if (idx >= n_aln) return;
	const SswAlnDesc d = alns[idx];
	const int h = d.half;

	/* best cell over the alignment's records: max score, then first scan position, then smallest row */
	int sc = 0, pos = 0x7fffffff, row = 0x7fffffff;
	for (int k = lane; k < d.n_items; k += 32) {
		const SswItemBest b = bests[d.first_item + k];
		const int s = b.score[h], p = b.pos[h], rr = b.row[h];
		if (s > 0 && (s > sc || (s == sc && (p < pos || (p == pos && rr < row))))) { sc = s; pos = p; row = rr; }
	}
#pragma unroll
	for (int off = 16; off >= 1; off >>= 1) {
		const int o_sc = __shfl_xor_sync(FULL, sc, off);
		const int o_pos = __shfl_xor_sync(FULL, pos, off), o_row = __shfl_xor_sync(FULL, row, off);
		if (o_sc > sc || (o_sc == sc && (o_pos < pos || (o_pos == pos && o_row < row)))) { sc = o_sc; pos = o_pos; row = o_row; }
	}
	if (sc == 0) { pos = 0; row = 0; }

	SswFillResult r;
	r.score = sc; r.ref = pos; r.read = row < d.read_len - 1 ? row : d.read_len - 1;
	r.score2 = 0; r.ref2 = 0; r.overflow = 0; r.pad_[0] = r.pad_[1] = 0;
	if (sc == 0) { r.ref = d.word ? 0 : -1; r.read = 0; }
	if (sc >= d.limit) { r.overflow = d.word ? 2 : 1; if (!d.word) r.score = 255; }

	if (SECOND && sc > 0 && !r.overflow && d.cm_off >= 0) {
		/* allowed columns: [0, e1) and [e2, refLen); smallest index of the largest value, values must be > 0 */
		const int e1 = max(pos - d.mask_len, 0);
		const int e2 = min(pos + d.mask_len, d.ref_len) + (d.word ? 0 : 1);
		const uint32_t* cm = colmax + d.cm_off;
		int v2 = 0, i2 = 0;
		if (d.scan_all) {
			for (int c = lane; c < d.ref_len; c += 32) {
				if (c < e1 || c >= e2) {
					const int v = half_of(cm[c], h);
					if (ssw_second_better(v, c, v2, i2)) { v2 = v; i2 = c; }
				}
			}
		} else
		for (int k0 = 0; k0 < d.n_items; k0 += 32) {
			const int k = k0 + lane;
			bool straddles = false;
			int p0 = 0, p1 = 0;
			if (k < d.n_items) {
				const SswItemBest b = bests[d.first_item + k];
				p0 = b.p0; p1 = b.p1;
				if (p1 <= e1 || p0 >= e2) { if (ssw_second_better(b.score[h], b.pos[h], v2, i2)) { v2 = b.score[h]; i2 = b.pos[h]; } }
				else straddles = true;
			}
			/* items touching the masked window: the warp scans their allowed columns element-wise */
			unsigned todo = __ballot_sync(FULL, straddles);
			while (todo) {
				const int src = __ffs((int)todo) - 1;
				todo &= todo - 1;
				const int q0 = __shfl_sync(FULL, p0, src), q1 = __shfl_sync(FULL, p1, src);
				for (int c = q0 + lane; c < q1; c += 32) {
					if (c < e1 || c >= e2) {
						const int v = half_of(cm[c], h);
						if (ssw_second_better(v, c, v2, i2)) { v2 = v; i2 = c; }
					}
				}
			}
		}
#pragma unroll
		for (int off = 16; off >= 1; off >>= 1) {
			const int o_v = __shfl_xor_sync(FULL, v2, off), o_i = __shfl_xor_sync(FULL, i2, off);
			if (ssw_second_better(o_v, o_i, v2, i2)) { v2 = o_v; i2 = o_i; }
		}
		r.score2 = v2; r.ref2 = i2;
	}
	if (lane == 0) out[idx] = r;
}

#endif /* SSW_RESOLVE_CUH */
