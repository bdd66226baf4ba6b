/*
 * ssw_fill.cuh -- the DP matrix fill (replaces the SSE2 loops of sw_sse2_byte,
 * src/ssw.c:258-340, and sw_sse2_word, :464-542, for both the forward pass and
 * the reverse begin-search pass, :919-926).
 *
 * Formulation (validated by oracle_fill_gotoh):
 *     X(c,r) = max(0, H(c-1,r-1) + s(c,r), E(c,r))
 *     H(c,r) = max(X(c,r), F(c,r))
 *     E(c+1,r) = max(E(c,r) - gapE, X(c,r) - gapO)
 *     F(c,r+1) = max(F(c,r) - gapE, X(c,r) - gapO)          (gapO > gapE)
 * in wrap-free signed 16-bit arithmetic, two alignments per register (s16x2
 * DPX instructions: VIADDMNMX.S16x2[.RELU], VIMNMX[3].S16x2, VIADD.16x2).
 * No lazy-F loop exists here: F is carried exactly down the column.
 *
 * Mapping.  A group of G lanes owns one item; lane t owns query rows
 * [t*R, t*R+R).  The group sweeps the reference as a skewed wavefront: in one
 * step lane t works on scan position s - t, so that the three values that
 * cross a lane boundary (bottom H, outgoing F, partial column maximum) are
 * handed down with one __shfl_up each.  Per-letter query profiles (both
 * alignments packed) live in shared memory; reference letters stream from
 * the padded reference array (null letters outside [0, refLen) keep the state
 * at exactly zero, so pipeline fill and drain need no predication).
 *
 * Outputs per item: the lexicographic best cell (score, first scan position,
 * smallest row) per half, and -- forward pass -- the column maxima over real
 * and pad rows (the reference's maxColumn[], ssw.c:338/:540) as packed s16x2
 * words.  All order-dependent semantics (strict '>' running maximum, byte
 * overflow, mask window) are applied afterwards by ssw_resolve.cuh.
 */
#ifndef SSW_FILL_CUH
#define SSW_FILL_CUH

#include "ssw_common.cuh"

#define SSW_FILL_WARPS 4
#define SSW_FILL_THREADS (SSW_FILL_WARPS * 32)

/* word offset of row k of lane `lane` inside one letter's profile block (32*R words):
 * R = 4a + rem; the first 4a rows are a uint4 segments [seg][lane], the tail is [lane][rem]. */
template <int R>
__host__ __device__ static __forceinline__ int ssw_prof_slot(int k, int lane)
{
	constexpr int A = R / 4, REM = R % 4;
	return k < 4 * A ? (k / 4) * 128 + lane * 4 + (k % 4) : A * 128 + lane * REM + (k - 4 * A);
}

/* Shared-memory reads of the profile.  On the device they are issued as ld.shared from a 32-bit shared
 * address (no generic-address conversion inside the sweep); the emulator build uses plain loads. */
#ifdef SSW_CPU_EMU
typedef const uint32_t* ssw_saddr;
__device__ static __forceinline__ ssw_saddr ssw_sbase(const uint32_t* p) { return p; }
__device__ static __forceinline__ ssw_saddr ssw_sadd(ssw_saddr a, int words) { return a + words; }
__device__ static __forceinline__ uint4 ssw_lds128(ssw_saddr a) { return *reinterpret_cast<const uint4*>(a); }
__device__ static __forceinline__ uint2 ssw_lds64(ssw_saddr a) { return *reinterpret_cast<const uint2*>(a); }
__device__ static __forceinline__ uint32_t ssw_lds32(ssw_saddr a) { return *a; }
#else
typedef uint32_t ssw_saddr;
__device__ static __forceinline__ ssw_saddr ssw_sbase(const uint32_t* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ static __forceinline__ ssw_saddr ssw_sadd(ssw_saddr a, int words) { return a + 4u * (uint32_t)words; }
__device__ static __forceinline__ uint4 ssw_lds128(ssw_saddr a)
{
	uint4 v;
	asm("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
	return v;
}
__device__ static __forceinline__ uint2 ssw_lds64(ssw_saddr a)
{
	uint2 v;
	asm("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a));
	return v;
}
__device__ static __forceinline__ uint32_t ssw_lds32(ssw_saddr a)
{
	uint32_t v;
	asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
	return v;
}
#endif

/* shared-memory bytes the fill kernel needs for an alphabet of n letters */
template <int R>
static inline size_t ssw_fill_smem_bytes(int n) { return (size_t)SSW_FILL_WARPS * (size_t)(n + 1) * 32 * R * sizeof(uint32_t); }

/*
 * MODE 0: every add is a DPX/VIADD instruction (ALU pipe).
 * MODE 1: "biased" arithmetic.  All of H, E, F, X carry a per-half bias B >= max(gapO, |min(mat)|), which makes
 *         two of the adds carry-safe as plain 32-bit operations, so they are issued as IMAD on the FMA pipe:
 *           t  = Hd + s      profile words are stored with the low half's carry pre-compensated in the high half
 *                            (low half >= 0 after the add for every real score; dead rows use -32768: never a carry)
 *           Xg = X - gapO    X >= B >= gapO in both halves: never a borrow
 *         leaving 4.5 ALU-pipe instructions per cell pair (VIMNMX3, 2 x VIADDMNMX, VIMNMX, 1/2 VIMNMX3) instead of 5.5.
 *         Scores are un-biased when they leave the kernel.
 */
template <int G, int R, int DIR, bool WRITE_CM, bool TERM, int MODE>
__global__ void __launch_bounds__(SSW_FILL_THREADS)
ssw_fill_kernel(const SswItem* __restrict__ items, int n_items,
                const int8_t* __restrict__ qcodes, const int8_t* __restrict__ refs,
                const int8_t* __restrict__ mat, int n, int gapO, int gapE,
                uint32_t* __restrict__ colmax, SswItemBest* __restrict__ bests)
{
	static_assert(G == 8 || G == 16 || G == 32, "group width");
	static_assert(R % 4 != 3 && R >= 1 && R <= 20, "rows per lane");
	static_assert(!TERM || G == 32, "early termination is per warp");
	constexpr int GPW = 32 / G;                 /* groups per warp */
	constexpr int A4 = R / 4, REM = R % 4;
	constexpr unsigned FULL = 0xffffffffu;

	SSW_DYN_SMEM(uint32_t, smem);
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int g = lane / G, t = lane % G;
	uint32_t* prof = smem + (size_t)warp * (size_t)(n + 1) * 32 * R;
	const int letter_stride = 32 * R;

	const int item_idx = ((int)blockIdx.x * SSW_FILL_WARPS + warp) * GPW + g;
	const bool live = item_idx < n_items;
	SswItem it;
	if (live) it = items[item_idx];
	else {
		it.qa.off = it.qb.off = 0; it.qa.len = it.qb.len = 0; it.qa.lp = it.qb.lp = 0; it.qa.rev = it.qb.rev = 0;
		it.ref_off = SSW_REF_PAD; it.ref_len = 0; it.cend = 0; it.p0 = it.p1 = 0; it.warm = 0; it.term_a = -1; it.cm_off = -1;
	}

	/* bias of MODE 1 (0 in MODE 0): B >= gapO and B >= -min(mat) */
	int B = 0;
	if (MODE == 1) {
		B = gapO > 1 ? gapO : 1;
		for (int i = lane; i < n * n; i += 32) B = max(B, -(int)mat[i]);
#pragma unroll
		for (int off = 16; off >= 1; off >>= 1) B = max(B, __shfl_xor_sync(FULL, B, off));
	}

	/* ---- build the packed query profile of this group (qP_byte/qP_word analogue, ssw.c:163-188/:388-410) ---- */
	{
		int ca[R], cb[R];
#pragma unroll
		for (int k = 0; k < R; ++k) {
			const int row = t * R + k;
			ca[k] = row < it.qa.len ? (int)qcodes[it.qa.off + (it.qa.rev ? it.qa.len - 1 - row : row)] : (row < it.qa.lp ? -1 : -2);
			cb[k] = row < it.qb.len ? (int)qcodes[it.qb.off + (it.qb.rev ? it.qb.len - 1 - row : row)] : (row < it.qb.lp ? -1 : -2);
		}
		for (int letter = 0; letter <= n; ++letter) {
			uint32_t* pl = prof + letter * letter_stride;
#pragma unroll
			for (int k = 0; k < R; ++k) {
				int a = SSW_NEG16, b = SSW_NEG16;
				if (letter < n) {
					a = ca[k] >= 0 ? (int)mat[letter * n + ca[k]] : (ca[k] == -1 ? 0 : SSW_NEG16);
					b = cb[k] >= 0 ? (int)mat[letter * n + cb[k]] : (cb[k] == -1 ? 0 : SSW_NEG16);
				}
				/* MODE 1: the 32-bit add Hd + s carries out of the low half exactly when the low score is a real negative one */
				if (MODE == 1 && a < 0 && a != SSW_NEG16) b -= 1;
				pl[ssw_prof_slot<R>(k, lane)] = pack2(a, b);
			}
		}
	}
	__syncwarp();

	/* ---- sweep ---- */
	const uint8_t* rp = reinterpret_cast<const uint8_t*>(refs) + it.ref_off;   /* reference column 0 */
	const uint32_t negO = pack2(-gapO, -gapO), negE = pack2(-gapE, -gapE);
	int sL = (it.p0 - it.warm - (G - 1)) & ~3;          /* scan position of the group's last lane, multiple of 4 */
	int n_body = live && it.p1 > sL ? (it.p1 - sL + 3) / 4 : 0;
#pragma unroll
	for (int off = G; off < 32; off <<= 1) n_body = max(n_body, __shfl_xor_sync(FULL, n_body, off));

	/* column cursor of this lane for step j = 0 of the current body, clamped into the padded array */
	int col = DIR > 0 ? sL + (G - 1 - t) : it.cend - (sL + (G - 1 - t));
	const int col_hi = it.ref_len + SSW_REF_PAD - 4, col_lo = -SSW_REF_PAD + 3;
	if (DIR > 0) col = min(col, col_hi); else col = max(col, col_lo);
	int sp0 = sL + (G - 1 - t);                         /* this lane's scan position at step j = 0 */

	/* per-lane shared-memory cursors: rows 0..4*A4-1 as uint4 segments, the tail behind them */
	const ssw_saddr pbase = ssw_sadd(ssw_sbase(prof), lane * 4);
	const ssw_saddr ptail = ssw_sadd(ssw_sbase(prof), A4 * 128 + lane * REM);
	uint32_t top_keep = t == 0 ? 0u : 1u;               /* lane 0 of a group takes zeros from above (multiplied in: FMA pipe) */
	const uint8_t* lptr = rp + col;                     /* letter cursor, advanced together with col */
#ifndef SSW_CPU_EMU
	asm volatile("" : "+r"(top_keep) : : "memory");     /* opaque 0/1 so that the masking stays a multiply */
#endif

	const uint32_t Bv = pack2(B, B);                    /* the zero level of every stored quantity */
	const uint32_t subO = 0u - ((uint32_t)gapO | ((uint32_t)gapO << 16));   /* MODE 1: 32-bit "- gapO" of both halves */
	const uint32_t top_add = t == 0 ? Bv : 0u;
	uint32_t one = 1u;
#ifndef SSW_CPU_EMU
	asm volatile("" : "+r"(one));                        /* opaque multiplier: keeps the MODE 1 adds as IMAD */
#endif
	uint32_t Hd[R], E[R];
#pragma unroll
	for (int k = 0; k < R; ++k) { Hd[k] = Bv; E[k] = Bv; }
	uint32_t outH = Bv, outF = Bv, outC = Bv;
	uint32_t best = Bv;
	int bpos0 = 0, bpos1 = 0, brow0 = 0, brow1 = 0;
	int stopped = 0;

	for (int body = 0; body < n_body; ++body) {
		uint32_t cmv[4];
		const bool maybe_counted = sp0 + 3 >= it.p0 && sp0 < it.p1;   /* this body touches the counted range */
#pragma unroll
		for (int j = 0; j < 4; ++j) {
			/* values crossing the lane boundary */
			const uint32_t inH = __shfl_up_sync(FULL, outH, 1, G) * top_keep + top_add;
			const uint32_t inF = __shfl_up_sync(FULL, outF, 1, G) * top_keep + top_add;
			const uint32_t inC = __shfl_up_sync(FULL, outC, 1, G) * top_keep + top_add;

			/* reference letter of this lane's scan position and its profile rows */
			int letter = (int)lptr[DIR * j];
			if (DIR < 0) { if (sp0 + j < 0) letter = n; }
			const ssw_saddr pl = ssw_sadd(pbase, letter * letter_stride);
			uint32_t s[R];
#pragma unroll
			for (int q = 0; q < A4; ++q) {
				const uint4 v = ssw_lds128(ssw_sadd(pl, q * 128));
				s[4 * q] = v.x; s[4 * q + 1] = v.y; s[4 * q + 2] = v.z; s[4 * q + 3] = v.w;
			}
			if (REM == 1) s[4 * A4] = ssw_lds32(ssw_sadd(ptail, letter * letter_stride));
			if (REM == 2) {
				const uint2 v = ssw_lds64(ssw_sadd(ptail, letter * letter_stride));
				s[4 * A4] = v.x; s[4 * A4 + 1] = v.y;
			}

			/* R cells of this lane's column */
			uint32_t F = inF, m = Bv, Hn[R];
#pragma unroll
			for (int k = 0; k < R; ++k) {
				uint32_t X, Xg;
				if (MODE == 1) {
					const uint32_t tt = Hd[k] * one + s[k];          /* IMAD: carry-compensated packed add */
					X = __vimax3_s16x2(tt, E[k], Bv);
					Xg = X * one + subO;                             /* IMAD: borrow-free packed subtract */
				} else {
					X = __viaddmax_s16x2_relu(Hd[k], s[k], E[k]);
					Xg = __vadd2(X, negO);
				}
				E[k] = __viaddmax_s16x2(E[k], negE, Xg);
				Hn[k] = __vmaxs2(X, F);
				F = __viaddmax_s16x2(F, negE, Xg);
			}
#pragma unroll
			for (int k = 0; k + 1 < R; k += 2) m = __vimax3_s16x2(m, Hn[k], Hn[k + 1]);
			if (R & 1) m = __vmaxs2(m, Hn[R - 1]);
			Hd[0] = inH;
#pragma unroll
			for (int k = 1; k < R; ++k) Hd[k] = Hn[k - 1];
			outH = Hn[R - 1];
			outF = F;
			outC = __vmaxs2(inC, m);
			cmv[j] = MODE == 1 ? outC * one + (0u - Bv) : outC;       /* un-biased column maximum (outC >= B: no borrow) */

			/* running best of this lane (strict increase only; rare path) */
			const uint32_t nb = __vmaxs2(best, m);
			if (nb != best && maybe_counted) {
				int sp = sp0 + j;
#ifndef SSW_CPU_EMU
				asm volatile("" : "+r"(sp));                /* keep the range test inside the rare path */
#endif
				if (sp >= it.p0 && sp < it.p1) {
					if (half_of(nb, 0) > half_of(best, 0)) {
						bpos0 = sp;
#pragma unroll
						for (int k = R - 1; k >= 0; --k) if (half_of(Hn[k], 0) == half_of(nb, 0)) brow0 = t * R + k;
					}
					if (half_of(nb, 1) > half_of(best, 1)) {
						bpos1 = sp;
#pragma unroll
						for (int k = R - 1; k >= 0; --k) if (half_of(Hn[k], 1) == half_of(nb, 1)) brow1 = t * R + k;
					}
					best = nb;
				}
			}
		}

		if (WRITE_CM) {
			if (t == G - 1 && sL >= it.p0 && sL < it.p1 && it.cm_off >= 0)
				*reinterpret_cast<uint4*>(colmax + it.cm_off + sL) = make_uint4(cmv[0], cmv[1], cmv[2], cmv[3]);
		}
		if (TERM) {
			/* reverse pass: stop after the first column whose maximum equals score1 (ssw.c:339/:541) */
			int hit = 0;
			if (t == G - 1 && it.term_a >= 0) {
#pragma unroll
				for (int j = 0; j < 4; ++j)
					if (sL + j >= it.p0 && sL + j < it.p1 && half_of(cmv[j], 0) == it.term_a) hit = 1;
			}
			if (__any_sync(FULL, hit)) { stopped = 1; break; }
		}

		sL += 4;
		sp0 += 4;
		{
			const int ncol = DIR > 0 ? min(col + 4, col_hi) : max(col - 4, col_lo);
			lptr += ncol - col;
			col = ncol;
		}
	}

	/* ---- reduce the group's lanes to one record per half: max score, then first position, then smallest row ---- */
	int sc0 = half_of(best, 0) - B, sc1 = half_of(best, 1) - B;
#pragma unroll
	for (int off = G / 2; off >= 1; off >>= 1) {
		const int o_sc0 = __shfl_down_sync(FULL, sc0, off, G), o_p0 = __shfl_down_sync(FULL, bpos0, off, G), o_r0 = __shfl_down_sync(FULL, brow0, off, G);
		const int o_sc1 = __shfl_down_sync(FULL, sc1, off, G), o_p1 = __shfl_down_sync(FULL, bpos1, off, G), o_r1 = __shfl_down_sync(FULL, brow1, off, G);
		if (o_sc0 > sc0 || (o_sc0 == sc0 && (o_p0 < bpos0 || (o_p0 == bpos0 && o_r0 < brow0)))) { sc0 = o_sc0; bpos0 = o_p0; brow0 = o_r0; }
		if (o_sc1 > sc1 || (o_sc1 == sc1 && (o_p1 < bpos1 || (o_p1 == bpos1 && o_r1 < brow1)))) { sc1 = o_sc1; bpos1 = o_p1; brow1 = o_r1; }
	}
	if (live && t == 0) {
		SswItemBest b;
		b.score[0] = sc0; b.pos[0] = bpos0; b.row[0] = brow0;
		b.score[1] = sc1; b.pos[1] = bpos1; b.row[1] = brow1;
		b.p0 = it.p0; b.p1 = it.p1;
		bests[item_idx] = b;
	}
}

#endif /* SSW_FILL_CUH */
