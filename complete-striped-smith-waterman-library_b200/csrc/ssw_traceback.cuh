/*
 * ssw_traceback.cuh -- banded affine-gap fill with direction bits, traceback
 * and CIGAR re-scoring: the follow-up kernel that replaces banded_sw
 * (src/ssw.c:590-783) and cigar_alignment_score (:785-811).
 *
 * One warp per alignment.  The band of query row i covers reference columns
 * [max(0,i-bw), min(refLen-1,i+bw)] (ssw.c:630-632).  A row is processed 32
 * columns at a time: E and the diagonal term depend only on the previous row;
 * the in-row gap F obeys  F(j+1) = max(Y(j) - gapO, F(j) - min(gapO,gapE))
 * with Y = max(max(E,0), diag + s), which is evaluated exactly with a warp
 * max-plus prefix scan.  The direction codes are pure functions of the cell
 * values, so they come out identical to the reference's scalar loop:
 *     de = 3 if H(i-1,j)-gapO >  E(i-1,j)-gapE else 2          (ssw.c:650-654)
 *     df = 5 if H(i,j-1)-gapO >  F(i,j-1)-gapE else 4          (ssw.c:656-659)
 *     dh = 1 if max(e1,f1) <= diag+s else (e1 > f1 ? de : df)  (ssw.c:661-675)
 * Reference quirks reproduced on purpose (see SURVEY Appendix A.5 and DESIGN.md):
 *   - the cell above the last band column is treated as outside the band for
 *     every row i <= bw+1 (the reference zeroes h_b[edge]/e_b[edge] with `edge`
 *     computed from the current row, ssw.c:633-637), even where the previous
 *     row did reach that column;
 *   - the running maximum and its cell persist across band doublings
 *     (ssw.c:601-602, :667-671), strict '>' in row-major order;
 *   - the traceback loop runs while i >= 0 && j > 0 and the tail rule appends
 *     one more M (ssw.c:690, :745-762);
 *   - a traceback that leaves the band reads the neighbouring row's cell, as
 *     the reference's flat direction array does (set_d, ssw.c:95).
 * Direction storage is 1 byte per band cell: bit0 de==3, bit1 df==5,
 * bits 2-3 source of H (0 diagonal, 1 E, 2 F).
 */
#ifndef SSW_TRACEBACK_CUH
#define SSW_TRACEBACK_CUH

#include <vector>
#include <functional>
#include "ssw_common.cuh"
#include "ssw_host.h"

#define SSW_TB_WARPS 4
#define SSW_TB_THREADS (SSW_TB_WARPS * 32)
#define SSW_TB_NEGINF (-(1 << 30))       /* INT32_MIN / 2, ssw.c:608 */

enum { SSW_TB_OK = 0, SSW_TB_WIDER = 1, SSW_TB_ERROR = 2, SSW_TB_MISMATCH = 3 };

struct SswTbTask {
	int64_t ref_off;     /* offset of ref[ref_begin1] in the padded reference array */
	int64_t read_off;    /* offset of read[read_begin1] in the query array */
	int32_t ref_len, read_len, score;
	int32_t bw;          /* band half-width of this round */
	int32_t max, max_i, max_j;   /* running maximum carried across band doublings */
	int32_t status, cig_len;
	int32_t init_bw;     /* band the current banded_sw call started with (retry rule, ssw.c:952-956) */
	int64_t dir_off;     /* byte offset of this task's direction cells in the scratch */
	int64_t row_off;     /* int32 offset of this task's 4 row buffers */
	int64_t cig_off;     /* word offset of this task's CIGAR buffer */
};

__device__ static __forceinline__ uint32_t ssw_tb_pack(uint32_t len, uint32_t op) { return (len << 4) | op; }  /* op: M0 I1 D2 */

/* Decide whether the band must be doubled (ssw.c:678-679); if not, walk the traceback (ssw.c:683-762) from the
 * maximum cell, run-length encode it and re-score the CIGAR (ssw.c:785-811).  One thread. */
__device__ static void ssw_tb_finish(SswTbTask& T, SswTbTask* slot, const int8_t* ref, const int8_t* read,
                                     const int8_t* __restrict__ mat, int n, int gapO, int gapE,
                                     const uint8_t* dir, uint32_t* cig_base)
{
	const int rl = T.ref_len, ql = T.read_len, bw = T.bw;
	const int W = 2 * bw + 1;
	const int len = rl > ql ? rl : ql;
	if (T.max < T.score && 2 * bw <= len) { T.status = SSW_TB_WIDER; *slot = T; return; }      /* ssw.c:678-679 */

	uint32_t* cig = cig_base + T.cig_off;
	const size_t dir_cells = (size_t)W * ql;
	int i = T.max_i, j = T.max_j, e = 0, l = 0, state = 2;
	int op = 0, prev = 0;                                         /* 0 M, 1 I, 2 D */
	bool bad = false;
	while (i >= 0 && j > 0) {
		const long long cell = (long long)W * i + (j - max(i - bw, 0));
		if (cell < -1 || cell > (long long)dir_cells) { bad = true; break; }      /* far outside the band: the reference reads unrelated memory */
		const int b = dir[cell];
		int code;
		if (state == 2) { const int hs = (b >> 2) & 3; code = hs == 0 ? 1 : (hs == 1 ? ((b & 1) ? 3 : 2) : (hs == 2 ? ((b & 2) ? 5 : 4) : 0)); }
		else if (state == 0) code = (b & 1) ? 3 : 2;
		else code = (b & 2) ? 5 : 4;
		if (code == 1) { --i; --j; state = 2; op = 0; }
		else if (code == 2) { --i; state = 0; op = 1; }
		else if (code == 3) { --i; state = 2; op = 1; }
		else if (code == 4) { --j; state = 1; op = 2; }
		else if (code == 5) { --j; state = 2; op = 2; }
		else { bad = true; break; }
		if (op == prev) ++e;
		else { cig[l++] = ssw_tb_pack((uint32_t)e, (uint32_t)prev); prev = op; e = 1; }
	}
	if (bad) { T.status = SSW_TB_ERROR; T.cig_len = 0; *slot = T; return; }
	if (op == 0) cig[l++] = ssw_tb_pack((uint32_t)(e + 1), 0);
	else { cig[l++] = ssw_tb_pack((uint32_t)e, (uint32_t)op); cig[l++] = ssw_tb_pack(1, 0); }
	for (int a = 0, b2 = l - 1; a < b2; ++a, --b2) { const uint32_t tmp = cig[a]; cig[a] = cig[b2]; cig[b2] = tmp; }

	int sc = 0, rp = 0, qp = 0;
	for (int k = 0; k < l; ++k) {
		const uint32_t clen = cig[k] >> 4, cop = cig[k] & 15;
		if (cop == 0) {
			for (uint32_t x = 0; x < clen; ++x) { sc += (int)mat[(int)ref[rp] * n + (int)read[qp]]; ++rp; ++qp; }
		} else {
			sc -= gapO + (clen > 1 ? (int)(clen - 1) * gapE : 0);
			if (cop == 1) qp += (int)clen; else rp += (int)clen;
		}
	}
	T.cig_len = l;
	T.status = sc == T.score ? SSW_TB_OK : SSW_TB_MISMATCH;
	*slot = T;
}

/*
 * Row-pipelined variant (bands up to SSW_TBP_MAXBW): one CTA per alignment, its warps take the query rows round
 * robin; row i starts a tile as soon as row i-1 has finished the same 32-column tile.  The H and E rows live in a
 * ring of shared-memory row slots (one more slot than warps, so a slot is only reused after its reader is done),
 * indexed by reference column modulo the ring width.  Cell arithmetic, quirks and outputs are those of the
 * single-warp kernel below.
 */
#define SSW_TBP_WARPS 4
#define SSW_TBP_SLOTS (SSW_TBP_WARPS + 1)
#define SSW_TBP_MAXBW 990
static int g_ssw_tb_maxbw = SSW_TBP_MAXBW;           /* "tb_maxbw" option: bands above this use the single-warp kernel (tests: 0) */

__global__ void __launch_bounds__(SSW_TBP_WARPS * 32)
ssw_banded_rows_kernel(SswTbTask* __restrict__ tasks,
                       const int8_t* __restrict__ qcodes, const int8_t* __restrict__ refs,
                       const int8_t* __restrict__ mat, int n, int gapO, int gapE,
                       uint8_t* __restrict__ dir_base, uint32_t* __restrict__ cig_base, int ring)
{
	constexpr unsigned FULL = 0xffffffffu;
	SSW_DYN_SMEM(int32_t, rows);                                  /* [slot][H|E][ring] */
	__shared__ volatile int p_row[SSW_TBP_SLOTS], p_tile[SSW_TBP_SLOTS];
	__shared__ int red_v[SSW_TBP_WARPS], red_i[SSW_TBP_WARPS], red_j[SSW_TBP_WARPS];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	SswTbTask T = tasks[blockIdx.x];
	const int8_t* ref = refs + T.ref_off;
	const int8_t* read = qcodes + T.read_off;
	const int rl = T.ref_len, ql = T.read_len, bw = T.bw;
	const int W = 2 * bw + 1, mask = ring - 1;
	const int g = gapO < gapE ? gapO : gapE;
	uint8_t* dir = dir_base + T.dir_off + 1;
	if (threadIdx.x < SSW_TBP_SLOTS) { p_row[threadIdx.x] = -1; p_tile[threadIdx.x] = 0; }
	__syncthreads();

	int bestv = 0, besti = 0, bestj = 0;
	for (int i = warp; i < ql; i += SSW_TBP_WARPS) {
		const int slot = i % SSW_TBP_SLOTS, pslot = (i + SSW_TBP_SLOTS - 1) % SSW_TBP_SLOTS;
		int32_t* Hc = rows + (size_t)slot * 2 * ring;
		int32_t* Ec = Hc + ring;
		const int32_t* Hp = rows + (size_t)pslot * 2 * ring;
		const int32_t* Ep = Hp + ring;
		const int beg = max(0, i - bw), end = min(rl - 1, i + bw);
		const int pbeg = max(0, i - 1 - bw);
		const bool top_oob = (i <= bw + 1) || (end == i + bw);
		const int rd = (int)read[i];
		uint8_t* drow = dir + (size_t)W * i - beg;
		if (lane == 0) { p_tile[slot] = beg >> 5; __threadfence_block(); p_row[slot] = i; }
		int carryF = -gapO + (beg & 31) * g;                      /* gives F(beg) = -gapO at the first active lane */
		int carryH = 0, carryFp = 0;
		for (int kt = beg >> 5; kt <= (end >> 5); ++kt) {
			const int j = kt * 32 + lane;
			const bool act = j >= beg && j <= end;
			if (i > 0) {
				if (lane == 0) { while (!(p_row[pslot] == i - 1 && p_tile[pslot] > kt)) { SSW_SPIN_PAUSE(); } }
				__syncwarp();
			}
			int Hup = 0, Eup = SSW_TB_NEGINF, Hdg = 0, s = 0;
			if (act) {
				if (i > 0) {
					if (!(j == end && top_oob)) { Hup = Hp[j & mask]; Eup = Ep[j & mask]; }
					if (j - 1 >= pbeg) Hdg = Hp[(j - 1) & mask];
				}
				s = (int)mat[(int)ref[j] * n + rd];
			}
			const int t1 = i == 0 ? -gapO : Hup - gapO;
			const int t2 = i == 0 ? SSW_TB_NEGINF : Eup - gapE;
			const int Ev = t1 > t2 ? t1 : t2;
			const int de3 = t1 > t2 ? 1 : 0;
			const int e1 = Ev > 0 ? Ev : 0;
			const int T2 = Hdg + s;
			const int Y = e1 > T2 ? e1 : T2;
			int P = act ? Y - gapO : SSW_TB_NEGINF;
#pragma unroll
			for (int d = 1; d < 32; d <<= 1) {
				const int o = __shfl_up_sync(FULL, P, d);
				if (lane >= d) P = max(P, o - d * g);
			}
			const int Pm1 = __shfl_up_sync(FULL, P, 1);
			const int Fv = lane == 0 ? carryF : max(Pm1, carryF - lane * g);
			const int Hv = Y > Fv ? Y : Fv;
			int Hl = __shfl_up_sync(FULL, Hv, 1), Fl = __shfl_up_sync(FULL, Fv, 1);
			if (lane == 0) { Hl = carryH; Fl = carryFp; }
			const int df5 = j == beg ? 1 : ((Hl - gapO > Fl - gapE) ? 1 : 0);
			const int f1 = Fv > 0 ? Fv : 0;
			const int T1 = e1 > f1 ? e1 : f1;
			const int hsel = T1 <= T2 ? 0 : (e1 > f1 ? 1 : 2);
			if (act) {
				Hc[j & mask] = Hv;
				Ec[j & mask] = Ev;
				drow[j] = (uint8_t)(de3 | (df5 << 1) | (hsel << 2));
				if (Hv > bestv) { bestv = Hv; besti = i; bestj = j; }
			}
			const int P31 = __shfl_sync(FULL, P, 31);
			carryH = __shfl_sync(FULL, Hv, 31);
			carryFp = __shfl_sync(FULL, Fv, 31);
			carryF = max(P31, carryF - 32 * g);
			__syncwarp();
			if (lane == 0) { __threadfence_block(); p_tile[slot] = kt + 1; }
		}
		if (lane == 0) { __threadfence_block(); p_tile[slot] = 0x7fffffff; }
	}
#pragma unroll
	for (int off = 16; off >= 1; off >>= 1) {
		const int ov = __shfl_xor_sync(FULL, bestv, off), oi = __shfl_xor_sync(FULL, besti, off), oj = __shfl_xor_sync(FULL, bestj, off);
		if (ov > bestv || (ov == bestv && (oi < besti || (oi == besti && oj < bestj)))) { bestv = ov; besti = oi; bestj = oj; }
	}
	if (lane == 0) { red_v[warp] = bestv; red_i[warp] = besti; red_j[warp] = bestj; }
	__threadfence_block();
	__syncthreads();
	if (threadIdx.x == 0) {
		for (int w = 1; w < SSW_TBP_WARPS; ++w)
			if (red_v[w] > bestv || (red_v[w] == bestv && (red_i[w] < besti || (red_i[w] == besti && red_j[w] < bestj)))) {
				bestv = red_v[w]; besti = red_i[w]; bestj = red_j[w];
			}
		if (bestv > T.max) { T.max = bestv; T.max_i = besti; T.max_j = bestj; }
		ssw_tb_finish(T, tasks + blockIdx.x, ref, read, mat, n, gapO, gapE, dir, cig_base);
	}
}

__global__ void __launch_bounds__(SSW_TB_THREADS)
ssw_banded_kernel(SswTbTask* __restrict__ tasks, int n_tasks,
                  const int8_t* __restrict__ qcodes, const int8_t* __restrict__ refs,
                  const int8_t* __restrict__ mat, int n, int gapO, int gapE,
                  uint8_t* __restrict__ dir_base, int32_t* __restrict__ row_base, uint32_t* __restrict__ cig_base)
{
	constexpr unsigned FULL = 0xffffffffu;
	const int lane = threadIdx.x & 31;
	const int ti = (int)blockIdx.x * SSW_TB_WARPS + (threadIdx.x >> 5);
	if (ti >= n_tasks) return;
	SswTbTask T = tasks[ti];
	const int8_t* ref = refs + T.ref_off;
	const int8_t* read = qcodes + T.read_off;
	const int rl = T.ref_len, ql = T.read_len, bw = T.bw;
	const int W = 2 * bw + 1;
	const int g = gapO < gapE ? gapO : gapE;
	uint8_t* dir = dir_base + T.dir_off + 1;                    /* one spare byte in front for x == -1 at row 0 */
	int32_t* Hrow[2] = {row_base + T.row_off + 1, row_base + T.row_off + (rl + 2) + 1};
	int32_t* Erow[2] = {row_base + T.row_off + 2 * (rl + 2) + 1, row_base + T.row_off + 3 * (rl + 2) + 1};

	/* ---------------- banded fill (ssw.c:628-676) ---------------- */
	int bestv = 0, besti = 0, bestj = 0;
	for (int i = 0; i < ql; ++i) {
		const int cur = i & 1, prv = cur ^ 1;
		const int beg = max(0, i - bw), end = min(rl - 1, i + bw);
		const int pbeg = max(0, i - 1 - bw);
		const bool top_oob = (i <= bw + 1) || (end == i + bw);   /* cell above column `end` reads as outside the band */
		const int rd = (int)read[i];
		uint8_t* drow = dir + (size_t)W * i - beg;              /* drow[j] is the cell of column j */
		int carryF = -gapO;                                      /* F of the tile's first column; F(beg) = max(0-gapO, neg_inf-gapE) */
		int carryH = 0, carryFp = 0;                             /* H, F of the column left of the tile */
		for (int j0 = beg; j0 <= end; j0 += 32) {
			const int j = j0 + lane;
			const bool act = j <= end;
			int Hup = 0, Eup = SSW_TB_NEGINF, Hdg = 0, s = 0;
			if (act) {
				if (i > 0) {
					if (!(j == end && top_oob)) { Hup = Hrow[prv][j]; Eup = Erow[prv][j]; }
					if (j - 1 >= pbeg) Hdg = Hrow[prv][j - 1];
				}
				s = (int)mat[(int)ref[j] * n + rd];
			}
			int t1 = i == 0 ? -gapO : Hup - gapO;
			int t2 = i == 0 ? SSW_TB_NEGINF : Eup - gapE;
			const int Ev = t1 > t2 ? t1 : t2;
			const int de3 = t1 > t2 ? 1 : 0;
			const int e1 = Ev > 0 ? Ev : 0;
			const int T2 = Hdg + s;
			const int Y = e1 > T2 ? e1 : T2;
			/* in-row gap: inclusive max-plus scan of A = Y - gapO with decay g per column */
			int P = act ? Y - gapO : SSW_TB_NEGINF;
#pragma unroll
			for (int d = 1; d < 32; d <<= 1) {
				const int o = __shfl_up_sync(FULL, P, d);
				if (lane >= d) P = max(P, o - d * g);
			}
			const int Pm1 = __shfl_up_sync(FULL, P, 1);
			const int Fv = lane == 0 ? carryF : max(Pm1, carryF - lane * g);
			const int Hv = Y > Fv ? Y : Fv;
			/* df from the left neighbour's H and F */
			int Hl = __shfl_up_sync(FULL, Hv, 1), Fl = __shfl_up_sync(FULL, Fv, 1);
			if (lane == 0) { Hl = carryH; Fl = carryFp; }
			int df5;
			if (j == beg) df5 = 1;                               /* 0 - gapO > neg_inf - gapE */
			else df5 = (Hl - gapO > Fl - gapE) ? 1 : 0;
			const int f1 = Fv > 0 ? Fv : 0;
			const int T1 = e1 > f1 ? e1 : f1;
			const int hsel = T1 <= T2 ? 0 : (e1 > f1 ? 1 : 2);
			if (act) {
				Hrow[cur][j] = Hv;
				Erow[cur][j] = Ev;
				drow[j] = (uint8_t)(de3 | (df5 << 1) | (hsel << 2));
				if (Hv > bestv) { bestv = Hv; besti = i; bestj = j; }
			}
			/* carries into the next tile */
			const int P31 = __shfl_sync(FULL, P, 31);
			carryH = __shfl_sync(FULL, Hv, 31);
			carryFp = __shfl_sync(FULL, Fv, 31);
			carryF = max(P31, carryF - 32 * g);
		}
		__syncwarp();
	}
	/* first row-major cell holding the maximum */
#pragma unroll
	for (int off = 16; off >= 1; off >>= 1) {
		const int ov = __shfl_xor_sync(FULL, bestv, off), oi = __shfl_xor_sync(FULL, besti, off), oj = __shfl_xor_sync(FULL, bestj, off);
		if (ov > bestv || (ov == bestv && (oi < besti || (oi == besti && oj < bestj)))) { bestv = ov; besti = oi; bestj = oj; }
	}
	if (bestv > T.max) { T.max = bestv; T.max_i = besti; T.max_j = bestj; }     /* strict, carried across doublings */
	__threadfence_block();
	__syncwarp();
	if (lane == 0) ssw_tb_finish(T, tasks + ti, ref, read, mat, n, gapO, gapE, dir, cig_base);
}

/*
 * Host driver of P3: band-doubling rounds (ssw.c:616-680) and the full-band
 * retry of ssw_align (:945-957).  `emit(i, words, len, failed)` is called once
 * per task with the final CIGAR (failed != 0: banded_sw gave up -> flag 1).
 */
static int ssw_traceback_run(cudaStream_t stream, std::vector<SswTbTask>& tasks,
                             const int8_t* d_q, const int8_t* d_r, const int8_t* d_mat, int n, int gapO, int gapE,
                             SswDevBuf* scratch, float* ms_acc, int64_t* launches,
                             const std::function<int(size_t, const uint32_t*, int32_t, int)>& emit)
{
	std::vector<size_t> active(tasks.size());
	for (size_t i = 0; i < tasks.size(); ++i) {
		SswTbTask& t = tasks[i];
		active[i] = i;
		int d = t.ref_len - t.read_len;
		t.bw = t.init_bw = (d < 0 ? -d : d) + 1;                 /* ssw.c:944 */
		t.max = 0; t.max_i = 0; t.max_j = 0; t.status = SSW_TB_WIDER; t.cig_len = 0;
	}
	size_t free_b = 0, total_b = 0;
	cudaMemGetInfo(&free_b, &total_b);
	const size_t budget = std::max<size_t>((size_t)64 << 20, (free_b + scratch->cap) / 2);
	SswTimer tm;
	std::vector<uint32_t> cig_host;
	auto ring_of = [](const SswTbTask& t) -> int {          /* 0: single-warp kernel with global row buffers */
		if (t.bw > g_ssw_tb_maxbw) return 0;
		int r = 256;
		while (r < 2 * t.bw + 66) r <<= 1;
		return r;
	};
	while (!active.empty()) {
		/* one launch = tasks of one kernel shape (ring width), as many as fit the scratch budget */
		std::stable_sort(active.begin(), active.end(), [&](size_t x, size_t y) { return ring_of(tasks[x]) < ring_of(tasks[y]); });
		const int ring = ring_of(tasks[active[0]]);
		std::vector<size_t> batch;
		size_t dir_bytes = 0, row_ints = 0, cig_words = 0;
		size_t k = 0;
		for (; k < active.size() && ring_of(tasks[active[k]]) == ring; ++k) {
			SswTbTask& t = tasks[active[k]];
			const size_t d = ((size_t)(2 * (size_t)t.bw + 1) * (size_t)t.read_len + 2 + 15) / 16 * 16;
			const size_t r = ring ? 0 : 4 * ((size_t)t.ref_len + 2);
			const size_t c = (size_t)t.ref_len + (size_t)t.read_len + 4;
			const size_t need = dir_bytes + d + 4 * (row_ints + r) + 4 * (cig_words + c) + sizeof(SswTbTask) * (batch.size() + 1) + 1024;
			if (!batch.empty() && need > budget) break;
			t.dir_off = (int64_t)dir_bytes; t.row_off = (int64_t)row_ints; t.cig_off = (int64_t)cig_words;
			dir_bytes += d; row_ints += r; cig_words += c;
			batch.push_back(active[k]);
		}
		std::vector<size_t> rest(active.begin() + k, active.end());
		const size_t off_rows = (dir_bytes + 255) / 256 * 256;
		const size_t off_cig = off_rows + (row_ints * 4 + 255) / 256 * 256;
		const size_t off_tasks = off_cig + (cig_words * 4 + 255) / 256 * 256;
		const size_t total = off_tasks + sizeof(SswTbTask) * batch.size();
		if (scratch->ensure(total)) return -1;
		std::vector<SswTbTask> bt(batch.size());
		for (size_t i = 0; i < batch.size(); ++i) bt[i] = tasks[batch[i]];
		uint8_t* base = scratch->as<uint8_t>();
		SswTbTask* d_tasks = reinterpret_cast<SswTbTask*>(base + off_tasks);
		SSW_CUDA_OK(cudaMemcpyAsync(d_tasks, bt.data(), sizeof(SswTbTask) * bt.size(), cudaMemcpyHostToDevice, stream));
		tm.start(stream);
		if (ring) {
			const size_t smem = (size_t)SSW_TBP_SLOTS * 2 * (size_t)ring * sizeof(int32_t);
			if (smem > 48 * 1024) SSW_CUDA_OK(cudaFuncSetAttribute(ssw_banded_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
			ssw_launch(ssw_banded_rows_kernel, dim3((unsigned)bt.size()), dim3(SSW_TBP_WARPS * 32), smem, stream,
			           d_tasks, d_q, d_r, d_mat, n, gapO, gapE, base, reinterpret_cast<uint32_t*>(base + off_cig), ring);
		} else {
			ssw_launch(ssw_banded_kernel, dim3(((int)bt.size() + SSW_TB_WARPS - 1) / SSW_TB_WARPS), dim3(SSW_TB_THREADS), 0, stream,
			           d_tasks, (int)bt.size(), d_q, d_r, d_mat, n, gapO, gapE,
			           base, reinterpret_cast<int32_t*>(base + off_rows), reinterpret_cast<uint32_t*>(base + off_cig));
		}
		SSW_CUDA_OK(cudaGetLastError());
		*ms_acc += tm.stop(stream);
		*launches += 1;
		SSW_CUDA_OK(cudaMemcpyAsync(bt.data(), d_tasks, sizeof(SswTbTask) * bt.size(), cudaMemcpyDeviceToHost, stream));
		cig_host.resize(cig_words);
		SSW_CUDA_OK(cudaMemcpyAsync(cig_host.data(), base + off_cig, cig_words * 4, cudaMemcpyDeviceToHost, stream));
		SSW_CUDA_OK(cudaStreamSynchronize(stream));
		std::vector<size_t> next;
		for (size_t i = 0; i < batch.size(); ++i) {
			SswTbTask& t = tasks[batch[i]];
			t = bt[i];
			const int full = t.ref_len > t.read_len ? t.ref_len : t.read_len;
			if (t.status == SSW_TB_WIDER) { t.bw *= 2; next.push_back(batch[i]); }
			else if (t.status == SSW_TB_OK) { if (emit(batch[i], cig_host.data() + t.cig_off, t.cig_len, 0)) return -1; }
			else if (t.status == SSW_TB_ERROR) { if (emit(batch[i], nullptr, 0, 1)) return -1; }
			else {                                                /* score mismatch: one retry at full band (ssw.c:952-956) */
				if (t.init_bw >= full) { if (emit(batch[i], nullptr, 0, 1)) return -1; }
				else { t.bw = t.init_bw = full; t.max = 0; t.max_i = 0; t.max_j = 0; next.push_back(batch[i]); }
			}
		}
		next.insert(next.end(), rest.begin(), rest.end());
		active.swap(next);
	}
	return 0;
}

#endif /* SSW_TRACEBACK_CUH */
