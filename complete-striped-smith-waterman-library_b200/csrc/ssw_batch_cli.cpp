/*
 * ssw_batch_cli.cpp -- `ssw_batch_cli`: a batched front end with the command line and the output of the
 * reference's `ssw_test` (src/main.c), SURVEY section 8(f) rank 1.
 *
 * The reference driver re-opens and re-parses the target file for every read and issues one blocking
 * ssw_align per (read, reference) pair (main.c:462-532, :493-494).  This driver parses both files once,
 * puts every pair -- and, with -r, every reverse-complement pair -- into ONE ssw_align_batch_text call (letters are
 * translated and reverse-complemented on the device), and
 * then prints the records in the reference's order and format (BLAST-like, main.c:129-206, or SAM,
 * main.c:207-244).  Same options: -m -x -o -e -p -a FILE -c -f N -r -s -h.
 *
 * Host-only code (parsing, formatting); all alignment work happens in libssw.so on the GPU.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

#include <string>
#include <vector>

#include "../../include/ssw.h"
#include "../../include/ssw_batch.h"

namespace {

struct Record { std::string name, seq, qual; };

/* FASTA / FASTQ (optionally gzip-compressed): name = first word of the header line */
bool read_records(const char* path, std::vector<Record>& out)
{
	gzFile f = gzopen(path, "r");
	if (!f) return false;
	std::string text;
	char buf[1 << 16];
	int got;
	while ((got = gzread(f, buf, sizeof buf)) > 0) text.append(buf, (size_t)got);
	gzclose(f);
	size_t pos = 0;
	auto next_line = [&](std::string& line) -> bool {
		if (pos >= text.size()) return false;
		size_t e = text.find('\n', pos);
		if (e == std::string::npos) e = text.size();
		line.assign(text, pos, e - pos);
		if (!line.empty() && line.back() == '\r') line.pop_back();
		pos = e + 1;
		return true;
	};
	std::string line;
	bool have = next_line(line);
	while (have) {
		if (line.empty() || (line[0] != '>' && line[0] != '@')) { have = next_line(line); continue; }
		const bool fastq = line[0] == '@';
		Record r;
		size_t e = 1;
		while (e < line.size() && line[e] != ' ' && line[e] != '\t') ++e;
		r.name.assign(line, 1, e - 1);
		have = next_line(line);
		while (have && !line.empty() && line[0] != '>' && line[0] != '@' && line[0] != '+') { r.seq += line; have = next_line(line); }
		while (have && line.empty()) have = next_line(line);
		if (fastq && have && line[0] == '+') {
			have = next_line(line);
			while (have && r.qual.size() < r.seq.size()) { r.qual += line; have = next_line(line); }
		}
		out.push_back(r);
	}
	return true;
}

/* letter -> code tables of the reference driver (main.c:72-93): nucleotides A C G T(U) -> 0..3, others 4;
 * amino acids in the order ARNDCQEGHILKMFPSTWYVBZX*, unknown -> 23 */
void default_tables(int8_t nt[128], int8_t aa[128])
{
	for (int i = 0; i < 128; ++i) { nt[i] = 4; aa[i] = 23; }
	const char* n = "ACGT";
	for (int i = 0; i < 4; ++i) { nt[(int)n[i]] = (int8_t)i; nt[(int)n[i] + 32] = (int8_t)i; }
	nt['U'] = nt['u'] = 3;
	const char* a = "ARNDCQEGHILKMFPSTWYVBZX";
	for (int i = 0; a[i]; ++i) { aa[(int)a[i]] = (int8_t)i; aa[(int)a[i] + 32] = (int8_t)i; }
}

/* BLOSUM50 in the reference driver's order (main.c:43-69), built from the 24x24 upper description below */
const int8_t kBlosum50[576] = {
	5, -2, -1, -2, -1, -1, -1, 0, -2, -1, -2, -1, -1, -3, -1, 1, 0, -3, -2, 0, -2, -1, -1, -5,
	-2, 7, -1, -2, -4, 1, 0, -3, 0, -4, -3, 3, -2, -3, -3, -1, -1, -3, -1, -3, -1, 0, -1, -5,
	-1, -1, 7, 2, -2, 0, 0, 0, 1, -3, -4, 0, -2, -4, -2, 1, 0, -4, -2, -3, 5, 0, -1, -5,
	-2, -2, 2, 8, -4, 0, 2, -1, -1, -4, -4, -1, -4, -5, -1, 0, -1, -5, -3, -4, 6, 1, -1, -5,
	-1, -4, -2, -4, 13, -3, -3, -3, -3, -2, -2, -3, -2, -2, -4, -1, -1, -5, -3, -1, -3, -3, -1, -5,
	-1, 1, 0, 0, -3, 7, 2, -2, 1, -3, -2, 2, 0, -4, -1, 0, -1, -1, -1, -3, 0, 4, -1, -5,
	-1, 0, 0, 2, -3, 2, 6, -3, 0, -4, -3, 1, -2, -3, -1, -1, -1, -3, -2, -3, 1, 5, -1, -5,
	0, -3, 0, -1, -3, -2, -3, 8, -2, -4, -4, -2, -3, -4, -2, 0, -2, -3, -3, -4, -1, -2, -1, -5,
	-2, 0, 1, -1, -3, 1, 0, -2, 10, -4, -3, 0, -1, -1, -2, -1, -2, -3, 2, -4, 0, 0, -1, -5,
	-1, -4, -3, -4, -2, -3, -4, -4, -4, 5, 2, -3, 2, 0, -3, -3, -1, -3, -1, 4, -4, -3, -1, -5,
	-2, -3, -4, -4, -2, -2, -3, -4, -3, 2, 5, -3, 3, 1, -4, -3, -1, -2, -1, 1, -4, -3, -1, -5,
	-1, 3, 0, -1, -3, 2, 1, -2, 0, -3, -3, 6, -2, -4, -1, 0, -1, -3, -2, -3, 0, 1, -1, -5,
	-1, -2, -2, -4, -2, 0, -2, -3, -1, 2, 3, -2, 7, 0, -3, -2, -1, -1, 0, 1, -3, -1, -1, -5,
	-3, -3, -4, -5, -2, -4, -3, -4, -1, 0, 1, -4, 0, 8, -4, -3, -2, 1, 4, -1, -4, -4, -1, -5,
	-1, -3, -2, -1, -4, -1, -1, -2, -2, -3, -4, -1, -3, -4, 10, -1, -1, -4, -3, -3, -2, -1, -1, -5,
	1, -1, 1, 0, -1, 0, -1, 0, -1, -3, -3, 0, -2, -3, -1, 5, 2, -4, -2, -2, 0, 0, -1, -5,
	0, -1, 0, -1, -1, -1, -1, -2, -2, -1, -1, -1, -1, -2, -1, 2, 5, -3, -2, 0, 0, -1, -1, -5,
	-3, -3, -4, -5, -5, -1, -3, -3, -3, -3, -2, -3, -1, 1, -4, -4, -3, 15, 2, -3, -5, -2, -1, -5,
	-2, -1, -2, -3, -3, -1, -2, -3, 2, -1, -1, -2, 0, 4, -3, -2, -2, 2, 8, -1, -3, -2, -1, -5,
	0, -3, -3, -4, -1, -3, -3, -4, -4, 4, 1, -3, 1, -1, -3, -2, 0, -3, -1, 5, -3, -3, -1, -5,
	-2, -1, 5, 6, -3, 0, 1, -1, 0, -4, -4, 0, -3, -4, -2, 0, 0, -5, -3, -3, 6, 1, -1, -5,
	-1, 0, 0, 1, -3, 4, 5, -2, 0, -3, -3, 1, -1, -4, -1, 0, -1, -2, -2, -3, 1, 5, -1, -5,
	-1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -5,
	-5, -5, -5, -5, -5, -5, -5, -5, -5, -5, -5, -5, -5, -5, -5, -5, -5, -5, -5, -5, -5, -5, -5, 1,
};

/* weight-matrix file (-a): rows starting with a letter or '*' give the alphabet order and the scores (main.c:341-390) */
bool read_matrix(const char* path, std::vector<int8_t>& mat, int& n, int8_t aa[128])
{
	FILE* f = fopen(path, "r");
	if (!f) return false;
	char line[256];
	n = 0;
	mat.clear();
	while (fgets(line, sizeof line, f)) {
		if (!(line[0] == '*' || (line[0] >= 'A' && line[0] <= 'Z'))) continue;
		if (line[0] >= 'A' && line[0] <= 'Z') aa[(int)line[0]] = aa[(int)line[0] + 32] = (int8_t)n;
		for (char* p = line + 1; *p;) {
			if ((*p >= '0' && *p <= '9') || *p == '-') { mat.push_back((int8_t)strtol(p, &p, 10)); }
			else ++p;
		}
		++n;
	}
	fclose(f);
	return n > 0 && (int)mat.size() == n * n;
}

std::string revcomp(const std::string& s)
{
	std::string r(s.size(), 'N');
	for (size_t i = 0; i < s.size(); ++i) {
		char c = s[s.size() - 1 - i], o = 4;                     /* the reference maps unknown letters to byte 4 (main.c:97-106) */
		switch (c) {
		case 'A': case 'a': o = 'T'; break;
		case 'C': case 'c': o = 'G'; break;
		case 'G': case 'g': o = 'C'; break;
		case 'T': case 't': case 'U': case 'u': o = 'A'; break;
		case 'N': case 'n': o = 'N'; break;
		default: break;
		}
		r[i] = o;
	}
	return r;
}

/* BLAST-like record (main.c:129-206): three 60-column rows per block */
void write_blast(const s_align* a, const Record& ref, const Record& read, const std::string& qseq, const int8_t* table, bool minus)
{
	printf("target_name: %s\nquery_name: %s\noptimal_alignment_score: %d\t", ref.name.c_str(), read.name.c_str(), a->score1);
	if (a->score2 > 0) printf("suboptimal_alignment_score: %d\t", a->score2);
	printf(minus ? "strand: -\t" : "strand: +\t");
	if (a->ref_begin1 + 1) printf("target_begin: %d\t", a->ref_begin1 + 1);
	printf("target_end: %d\t", a->ref_end1 + 1);
	if (a->read_begin1 + 1) printf("query_begin: %d\t", a->read_begin1 + 1);
	printf("query_end: %d\n\n", a->read_end1 + 1);
	if (!a->cigar) return;
	/* expand the path into aligned columns, then print it 60 columns at a time */
	std::string top, mid, bot;
	std::vector<int> tpos, qpos;                          /* 0-based index of the NEXT target / query residue after each column */
	int t = a->ref_begin1, q = a->read_begin1;
	for (int c = 0; c < a->cigarLen; ++c) {
		const char op = cigar_int_to_op(a->cigar[c]);
		const uint32_t len = cigar_int_to_len(a->cigar[c]);
		for (uint32_t i = 0; i < len; ++i) {
			if (op == 'M') {
				top += ref.seq[t]; bot += qseq[q];
				mid += table[(int)ref.seq[t]] == table[(int)qseq[q]] ? '|' : '*';
				++t; ++q;
			} else if (op == 'I') { top += '-'; mid += ' '; bot += qseq[q]; ++q; }
			else { top += ref.seq[t]; mid += ' '; bot += '-'; ++t; }
			tpos.push_back(t); qpos.push_back(q);
		}
	}
	int t0 = a->ref_begin1, q0 = a->read_begin1;
	for (size_t s = 0; s < top.size(); s += 60) {
		const size_t e = s + 60 < top.size() ? s + 60 : top.size();
		printf("Target: %8d    %s    %d\n                    %s\nQuery:  %8d    %s    %d\n\n",
		       t0 + 1, top.substr(s, e - s).c_str(), tpos[e - 1], mid.substr(s, e - s).c_str(),
		       q0 + 1, bot.substr(s, e - s).c_str(), qpos[e - 1]);
		t0 = tpos[e - 1]; q0 = qpos[e - 1];
	}
}

/* SAM record (main.c:207-244); rewrites the CIGAR with mark_mismatch like the reference does */
void write_sam(s_align* a, const Record& ref, const Record& read, const std::string& qseq,
               const int8_t* ref_num, const int8_t* read_num, bool minus)
{
	printf("%s\t", read.name.c_str());
	if (a->score1 == 0) { printf("4\t*\t0\t255\t*\t*\t0\t0\t*\t*\n"); return; }
	uint32_t mapq = (uint32_t)(-4.343 * log(1 - (double)abs(a->score1 - a->score2) / (double)a->score1));
	mapq = (uint32_t)(mapq + 4.99);
	mapq = mapq < 254 ? mapq : 254;
	printf(minus ? "16\t" : "0\t");
	printf("%s\t%d\t%d\t", ref.name.c_str(), a->ref_begin1 + 1, mapq);
	const int32_t nm = mark_mismatch(a->ref_begin1, a->read_begin1, a->read_end1, ref_num, read_num, (int32_t)read.seq.size(), &a->cigar, &a->cigarLen);
	for (int c = 0; c < a->cigarLen; ++c) printf("%lu%c", (unsigned long)cigar_int_to_len(a->cigar[c]), cigar_int_to_op(a->cigar[c]));
	printf("\t*\t0\t0\t%s\t", qseq.c_str());
	if (!read.qual.empty() && minus) { for (size_t p = read.qual.size(); p-- > 0;) putchar(read.qual[p]); }
	else if (!read.qual.empty()) printf("%s", read.qual.c_str());
	else printf("*");
	printf("\tAS:i:%d\tNM:i:%d\t", a->score1, nm);
	if (a->score2 > 0) printf("ZS:i:%d\n", a->score2); else printf("\n");
}

void usage()
{
	fprintf(stderr, "\nUsage: ssw_batch_cli [options] ... <target.fasta> <query.fasta>(or <query.fastq>)\n"
	                "Options (as ssw_test): -m N  -x N  -o N  -e N  -p  -a FILE  -c  -f N  -r  -s  -h\n"
	                "         and -g N: cut the batch over N GPUs (0: all visible ones; default 1)\n\n");
}

}  // namespace

int main(int argc, char** argv)
{
	int match = 2, mismatch = 2, gap_open = 3, gap_ext = 1, filter = 0, gpus = 1;
	bool protein = false, path = false, reverse = false, sam = false, header = false;
	const char* mat_name = nullptr;
	std::vector<const char*> files;
	for (int i = 1; i < argc; ++i) {
		if (argv[i][0] != '-') { files.push_back(argv[i]); continue; }
		for (const char* c = argv[i] + 1; *c; ++c) {
			const bool has_val = i + 1 < argc && argv[i + 1][0] != '-';
			switch (*c) {
			case 'm': if (has_val) match = atoi(argv[++i]); break;
			case 'x': if (has_val) mismatch = atoi(argv[++i]); break;
			case 'o': if (has_val) gap_open = atoi(argv[++i]); break;
			case 'e': if (has_val) gap_ext = atoi(argv[++i]); break;
			case 'f': if (has_val) filter = atoi(argv[++i]); break;
			case 'a': if (has_val) mat_name = argv[++i]; break;
			case 'g': if (has_val) gpus = atoi(argv[++i]); break;
			case 'p': protein = true; break;
			case 'c': path = true; break;
			case 'r': reverse = true; break;
			case 's': sam = true; break;
			case 'h': header = true; break;
			default: break;
			}
			if (*c == 'm' || *c == 'x' || *c == 'o' || *c == 'e' || *c == 'f' || *c == 'a' || *c == 'g') break;
		}
	}
	if (files.size() < 2) { usage(); return 1; }

	int8_t nt_table[128], aa_table[128];
	default_tables(nt_table, aa_table);
	std::vector<int8_t> mat(25, 0);
	for (int i = 0, k = 0; i < 4; ++i) { for (int j = 0; j < 4; ++j) mat[k++] = (int8_t)(i == j ? match : -mismatch); mat[k++] = 0; }
	int n = 5;
	const int8_t* table = nt_table;
	if (protein && !mat_name) { mat.assign(kBlosum50, kBlosum50 + 576); n = 24; table = aa_table; }
	else if (mat_name) {
		if (!read_matrix(mat_name, mat, n, aa_table)) { fprintf(stderr, "Problem of reading the weight matrix file.\n"); return 1; }
		table = aa_table;
	}
	if (reverse && n != 5) { fprintf(stderr, "Reverse complement alignment is not available for protein sequences. \n"); return 1; }

	std::vector<Record> refs, reads;
	if (!read_records(files[0], refs)) { fprintf(stderr, "gzopen of '%s' failed.\n", files[0]); return 1; }
	if (!read_records(files[1], reads)) { fprintf(stderr, "gzopen of '%s' failed.\n", files[1]); return 1; }
	if (sam && header && path) {
		printf("@HD\tVN:1.4\tSO:queryname\n");
		for (const Record& r : refs) printf("@SQ\tSN:%s\tLN:%d\n", r.name.c_str(), (int)r.seq.size());
	} else if (sam && !path) {
		fprintf(stderr, "SAM format output is only available together with option -c.\n");
		sam = false;
	}

	/* one batch: the letters go to the device as they are; translation with `table`, the reverse complements (query
	 * n + k = reverse complement of read k) and the padded reference layout are made there */
	const bool rc = reverse && !protein;
	std::vector<std::string> rc_seq(rc ? reads.size() : 0);          /* only for printing the minus-strand records */
	std::string qtext, rtext;
	std::vector<int64_t> qoff(1, 0), roff(1, 0);
	for (const Record& r : reads) { qtext += r.seq; qoff.push_back((int64_t)qtext.size()); }
	for (const Record& r : refs) { rtext += r.seq; roff.push_back((int64_t)rtext.size()); }
	const int32_t n_q = (int32_t)reads.size() * (rc ? 2 : 1), n_r = (int32_t)refs.size();
	if (n_q == 0 || n_r == 0) return 0;

	/* -g N: the batch is cut over N GPUs of this process (0: all visible ones), see ssw_batch.h "Device groups" */
	ssw_engine* eng = gpus == 1 ? ssw_engine_create(-1) : nullptr;
	ssw_group* grp = gpus == 1 ? nullptr : ssw_group_create(gpus < 0 ? 0 : gpus, nullptr);
	if (!eng && !grp) return 1;
	ssw_batch_params P;
	memset(&P, 0, sizeof P);
	P.mat = mat.data(); P.n = n; P.gap_open = (uint8_t)gap_open; P.gap_extend = (uint8_t)gap_ext;
	P.flag = path ? 2 : 0; P.filters = (uint16_t)filter; P.filterd = 0; P.mask_len = -1; P.score_size = 2;
	std::vector<s_align*> out((size_t)n_q * n_r, nullptr);
	const int failed = eng
		? ssw_align_batch_text(eng, &P, table, rc ? 1 : 0, (int32_t)reads.size(), qtext.data(), qoff.data(), n_r, rtext.data(), roff.data(),
		                       (int64_t)n_q * n_r, nullptr, nullptr, out.data())
		: ssw_group_align_batch(grp, &P, table, rc ? 1 : 0, (int32_t)reads.size(), qtext.data(), qoff.data(), n_r, rtext.data(), roff.data(),
		                        (int64_t)n_q * n_r, nullptr, nullptr, out.data(), 0, nullptr);
	if (failed) {
		fprintf(stderr, "ssw_align_batch_text failed\n");
		return 1;
	}
	/* host-side codes are only needed by the SAM writer (mark_mismatch): translated on first use */
	std::vector<std::vector<int8_t>> ref_codes(refs.size());
	auto codes_of = [&](const std::string& text) { std::vector<int8_t> c(text.size()); for (size_t i = 0; i < text.size(); ++i) c[i] = table[(int)(text[i] & 127)]; return c; };
	auto ref_num_of = [&](int32_t ri) -> const int8_t* { if (ref_codes[ri].empty()) ref_codes[ri] = codes_of(refs[ri].seq); return ref_codes[ri].data(); };

	for (size_t qi = 0; qi < reads.size(); ++qi) {
		for (int32_t ri = 0; ri < n_r; ++ri) {
			s_align* res = out[qi * n_r + ri];
			s_align* res_rc = rc ? out[(reads.size() + qi) * n_r + ri] : nullptr;
			if (!res) {
				fprintf(stderr, "Warning: Alignment between the following sequences is failed.\nref_name: %s\nread_name: %s\n\n", refs[ri].name.c_str(), reads[qi].name.c_str());
				continue;
			}
			if (res_rc && res_rc->score1 > res->score1 && res_rc->score1 >= filter) {
				if (res_rc->flag == 2) fprintf(stderr, "Warning: The reverse compliment alignment of the following sequences may miss a small part.\nref_seq: %s\nread_seq: %s\n\n", refs[ri].name.c_str(), reads[qi].name.c_str());
				if (rc_seq[qi].empty()) rc_seq[qi] = revcomp(reads[qi].seq);
				if (sam) { const std::vector<int8_t> qn = codes_of(rc_seq[qi]); write_sam(res_rc, refs[ri], reads[qi], rc_seq[qi], ref_num_of(ri), qn.data(), true); }
				else write_blast(res_rc, refs[ri], reads[qi], rc_seq[qi], table, true);
			} else if (res->score1 > 0 && res->score1 >= filter) {
				if (res->flag == 2) fprintf(stderr, "Warning: The alignment of the following sequences may miss a small part.\nref_seq: %s\nread_seq: %s\n\n", refs[ri].name.c_str(), reads[qi].name.c_str());
				if (sam) { const std::vector<int8_t> qn = codes_of(reads[qi].seq); write_sam(res, refs[ri], reads[qi], reads[qi].seq, ref_num_of(ri), qn.data(), false); }
				else write_blast(res, refs[ri], reads[qi], reads[qi].seq, table, false);
			} else if (res->score1 <= 0) {
				fprintf(stderr, "There is no identical residue between the following reference and read seqeunces.\nref_name: %s\nread_name: %s\n\n", refs[ri].name.c_str(), reads[qi].name.c_str());
			}
		}
	}
	for (s_align* a : out) if (a) align_destroy(a);
	if (eng) ssw_engine_destroy(eng);
	if (grp) ssw_group_destroy(grp);
	return 0;
}
