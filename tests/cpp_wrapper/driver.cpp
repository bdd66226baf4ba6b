// Test driver for the C++ wrapper (test infrastructure).  It only uses the public API shared by the reference's
// ssw_cpp.h and include/ssw_cpp.h, so the same source is compiled (a) in the build container against the UNMODIFIED
// reference wrapper + ssw.c to freeze tests/golden/cpp_wrapper.txt (make_cpp_wrapper_golden.py) and (b) against our
// wrapper.  With -DWITH_BATCH the "batch" section goes through Aligner::AlignBatch instead of a loop over Align.
#include <stdint.h>
#include <stdio.h>

#include <algorithm>
#include <string>
#include <vector>

#include "ssw_cpp.h"

using namespace StripedSmithWaterman;

static uint64_t g_state = 12345;
static uint32_t rnd() { g_state = g_state * 6364136223846793005ull + 1442695040888963407ull; return (uint32_t)(g_state >> 33); }

static std::string random_seq(size_t n, const char* alphabet, int k)
{
	std::string s(n, 'A');
	for (size_t i = 0; i < n; ++i) s[i] = alphabet[rnd() % k];
	return s;
}

static std::string mutate(const std::string& src, const char* alphabet, int k, int rate)
{
	std::string out;
	for (size_t i = 0; i < src.size(); ++i) {
		const uint32_t r = rnd() % 100;
		if ((int)r < rate) out += alphabet[rnd() % k];                                   // substitution
		else if ((int)r < rate + rate / 3) { out += alphabet[rnd() % k]; out += src[i]; } // insertion
		else if ((int)r < rate + 2 * (rate / 3)) continue;                                // deletion
		else out += src[i];
	}
	return out;
}

static void show(const char* tag, int idx, uint16_t rc, const Alignment& a)
{
	printf("%s %d rc=%d s=%d s2=%d r=[%d,%d] q=[%d,%d] r2=%d mm=%d cigar=%s n=%d", tag, idx, (int)rc, (int)a.sw_score,
	       (int)a.sw_score_next_best, a.ref_begin, a.ref_end, a.query_begin, a.query_end, a.ref_end_next_best, a.mismatches,
	       a.cigar_string.c_str(), (int)a.cigar.size());
	uint32_t h = 0;
	for (uint32_t w : a.cigar) h = h * 31 + w;
	printf(" h=%u\n", h);
}

int main()
{
	const char* dna = "ACGTN";
	const std::string ref = random_seq(3000, dna, 4);
	std::vector<std::string> queries;
	for (int i = 0; i < 48; ++i) {
		const size_t len = 20 + rnd() % 260, pos = rnd() % (ref.size() - len);
		std::string q = mutate(ref.substr(pos, len), dna, i % 7 == 0 ? 5 : 4, 3 + (i % 5) * 4);
		if (i % 11 == 0) q = random_seq(5, dna, 4) + q + random_seq(9, dna, 4);           // clipped ends
		if (i % 13 == 5) for (size_t k = 0; k < q.size(); k += 2) q[k] = (char)(q[k] + 32);  // lower case
		queries.push_back(q);
	}

	// 1. default aligner, stored reference, several filters and mask lengths
	Aligner al;
	printf("setref %d\n", (int)al.SetReferenceSequence(ref.c_str(), ref.size()));
	Filter f_all, f_begin, f_score, f_sf, f_df;
	f_begin.report_cigar = false;
	f_score.report_cigar = false; f_score.report_begin_position = false;
	f_sf.score_filter = 120;
	f_df.distance_filter = 80;
	const Filter* filters[5] = {&f_all, &f_begin, &f_score, &f_sf, &f_df};
	for (size_t i = 0; i < queries.size(); ++i) {
		Alignment a;
		const Filter& f = *filters[i % 5];
		const int32_t mask = i % 3 == 0 ? 0 : (i % 3 == 1 ? 15 : (int32_t)queries[i].size() / 2);
		const uint16_t rc = al.Align(queries[i].c_str(), queries[i].size(), f, a, mask);
		show("stored", (int)i, rc, a);
	}

	// 2. explicit reference, null-terminated overloads, other penalties
	Aligner al2(3, 4, 6, 2);
	for (size_t i = 0; i < queries.size(); i += 3) {
		Alignment a;
		const std::string sub = ref.substr((i * 53) % 1500, 900);
		const uint16_t rc = al2.Align(queries[i].c_str(), sub.c_str(), f_all, a, 20);
		show("explicit", (int)i, rc, a);
	}
	al2.SetGapPenalty(2, 1);
	{
		Alignment a;
		const uint16_t rc = al2.Align(queries[1].c_str(), queries[1].size(), ref.c_str(), ref.size(), f_all, a, 16);
		show("gap21", 1, rc, a);
	}

	// 3. empty inputs and the cleared / rebuilt aligner
	{
		Alignment a;
		printf("empty-query %d\n", (int)al.Align("", f_all, a));
		Aligner none;
		printf("no-ref %d\n", (int)none.Align("ACGT", f_all, a));
		printf("rebuild-live %d\n", (int)al.ReBuild());
		al.Clear();
		printf("cleared %d setref %d\n", (int)al.Align("ACGT", "ACGT", f_all, a), (int)al.SetReferenceSequence("ACGT"));
		printf("rebuild %d\n", (int)al.ReBuild(1, 3, 5, 2));
		const uint16_t rc = al.Align(queries[2].c_str(), ref.c_str(), f_all, a, 30);
		show("rebuilt", 2, rc, a);
	}

	// 4. custom alphabet: 6 letters with an asymmetric-free integer matrix
	{
		const char* letters = "KLMPQW";
		int8_t mat[36], table[128];
		for (int i = 0; i < 128; ++i) table[i] = 5;
		for (int i = 0; i < 6; ++i) table[(int)letters[i]] = (int8_t)i;
		for (int i = 0; i < 6; ++i)
			for (int j = 0; j < 6; ++j) mat[i * 6 + j] = (int8_t)(i == j ? 4 + i % 3 : -1 - (i + j) % 4);
		Aligner pa(mat, 6, table, 128);
		pa.SetGapPenalty(5, 1);
		const std::string pref = random_seq(1200, letters, 6);
		pa.SetReferenceSequence(pref.c_str());
		for (int i = 0; i < 12; ++i) {
			const size_t len = 30 + rnd() % 300, pos = rnd() % (pref.size() - len);
			const std::string q = mutate(pref.substr(pos, len), letters, 6, 6);
			Alignment a;
			const uint16_t rc = pa.Align(q.c_str(), f_all, a, (int32_t)q.size() / 2);
			show("custom", i, rc, a);
		}
	}

	// 5. the whole query set against the stored reference
	{
		Aligner b;
		b.SetReferenceSequence(ref.c_str(), ref.size());
		std::vector<std::string> qs = queries;
		qs.insert(qs.begin() + 7, std::string());                    // an empty query is skipped, not an error
		std::vector<Alignment> res(qs.size());
		std::vector<uint16_t> rcs(qs.size(), 0);
#ifdef WITH_BATCH
#ifdef BATCH_DEVICES          // our wrapper only: the batch cut over that many GPUs (device groups) -- same output
		const bool ok = b.AlignBatch(qs, f_all, res, &rcs, 25, BATCH_DEVICES);
#else
		const bool ok = b.AlignBatch(qs, f_all, res, &rcs, 25);
#endif
		if (!ok) printf("AlignBatch failed\n");
#else
		for (size_t i = 0; i < qs.size(); ++i) rcs[i] = b.Align(qs[i].c_str(), qs[i].size(), f_all, res[i], 25);
#endif
		for (size_t i = 0; i < qs.size(); ++i) show("batch", (int)i, rcs[i], res[i]);
	}
	return 0;
}
