"""CPU tests of the device groups (include/ssw_batch.h: ssw_group_*; csrc/ssw_group.cpp): one batch cut over several
devices of one process.  The emulator build (tests/cuda_emu, TEST INFRASTRUCTURE) reports SSW_EMU_DEVICES identical devices,
so the host logic -- cutting a grid by queries and a pair list by DP cells, scattering the records back into pair order,
concatenating the CIGAR words of the devices, reverse complements and marked CIGARs per shard -- runs here without a GPU.
Every result is compared with one engine's answer for the same pairs and with the CPU checker."""
import os
import subprocess

import numpy as np
import pytest

import common as C
from test_emulated_kernels import EMU_DIR, _pkg

FIELDS = ("score1", "score2", "ref_begin1", "ref_end1", "read_begin1", "read_end1", "ref_end2", "flag", "cigar_len", "status")


def _same(res_a, pool_a, res_b, pool_b):
    assert len(res_a) == len(res_b)
    for i in range(len(res_a)):
        a, b = res_a[i], res_b[i]
        for k in FIELDS:
            assert int(a[k]) == int(b[k]), (i, k, int(a[k]), int(b[k]))
        if a["cigar_len"] > 0:
            assert list(pool_a[a["cigar_off"]: a["cigar_off"] + a["cigar_len"]]) == list(pool_b[b["cigar_off"]: b["cigar_off"] + b["cigar_len"]]), i


@pytest.fixture(scope="module")
def lib():
    subprocess.run(["make", "-s", "-C", EMU_DIR], check=True)
    os.environ["SSW_EMU_DEVICES"] = "5"
    yield _pkg()
    os.environ.pop("SSW_EMU_DEVICES", None)


def _workload(seed, n_reads=9):
    rng = np.random.default_rng(seed)
    refs = [rng.integers(0, 4, size=int(n)).astype(np.int8) for n in (520, 260, 800)]
    reads = []
    for k in range(n_reads):
        r = refs[k % 3]
        n = int(rng.choice([30, 80, 150, 151, 260, 300]))
        n = min(n, len(r) - 20)
        reads.append(C.mutate_read(rng, r, int(rng.integers(0, len(r) - n - 5)), n, 0.06, 0.02, 0.02))
    return reads, refs


def run_group_cases(grp, one, ref_len, n_reads, n_queries, n_targets, n_check, long_ref, long_len, n_long=12):
    """The cases of tests/test_gpu_parity_group.py (full sizes on the GPU) -- run here at toy sizes on the emulator build:
    a config-2-like grid cut by queries, a config-4-like protein grid through the device-planned path of every engine, and a
    list of long and short reads with marked CIGARs cut by DP cells.  grp / one: GroupAligner / BatchAligner of the same library."""
    ref, reads = C.make_dna_workload(ref_len, n_reads, 150, seed_ref=1001, seed_reads=2002)
    mat = C.dna_matrix(2, 2)
    pq, pr = np.arange(len(reads)), np.zeros(len(reads), dtype=np.int64)
    for flag in (0, 0x0f):
        res, pool = grp.align(reads, [ref], mat, 5, 3, 1, flag=flag, filterd=32767, mask_len=75, score_size=2)
        exp, exp_pool, _, _, kind = C.cpu_batch(reads, [ref], pq, pr, mat, 5, 3, 1, flag=flag, filterd=32767, mask_len=75, score_size=2)
        bad = C.compare_records(res, pool, exp, exp_pool)
        assert bad == [], (flag, bad[:5])
        cells = [t["cells_forward"] for t in grp.timing()]
        assert min(cells) > 0, cells                       # every engine of the group took part
    print("group of %d engines: 2 x %d records equal to the %s checker" % (grp.size, len(reads), kind))
    W = C.config_workload(4, n_queries=n_queries, n_targets=n_targets)
    one.set_sequences(W["queries"], W["refs"])
    base, base_pool = one.align(C.BLOSUM50, 24, 3, 1, flag=0, mask_len=150, score_size=1)
    res, pool = grp.align(W["queries"], W["refs"], C.BLOSUM50, 24, 3, 1, flag=0, mask_len=150, score_size=1)
    assert C.compare_records(res, pool, base, base_pool) == []
    nq, nt = len(W["queries"]), len(W["refs"])
    idx = np.random.default_rng(9).choice(nq * nt, size=min(n_check, nq * nt), replace=False)
    exp, exp_pool, _, _, _ = C.cpu_batch(W["queries"], W["refs"], idx // nt, idx % nt, C.BLOSUM50, 24, 3, 1, flag=0, mask_len=150, score_size=1)
    assert C.compare_records(res, pool, exp, exp_pool, idx=idx) == []
    refL, readsL = C.make_dna_workload(long_ref, n_long, long_len, seed_ref=5005, seed_reads=5006)
    qs = readsL + reads[:n_long]
    lp = np.random.default_rng(10).integers(0, len(qs), size=40).astype(np.int32)
    lr = np.zeros(40, dtype=np.int32)
    res, pool, nm = grp.align(qs, [refL], mat, 5, 3, 1, flag=2, mask_len=40, score_size=2, pair_query=lp, pair_ref=lr, marked=True)
    one.set_sequences(qs, [refL])
    b, bp = one.align(mat, 5, 3, 1, flag=2, mask_len=40, score_size=2, pair_query=lp, pair_ref=lr)
    exp, exp_pool, _, _, _ = C.cpu_batch(qs, [refL], lp, lr, mat, 5, 3, 1, flag=2, mask_len=40, score_size=2)
    assert C.compare_records(b, bp, exp, exp_pool) == []
    mb, mbp, mnm = one.mark_mismatch(b, bp, pair_query=lp, pair_ref=lr)
    assert C.compare_records(res, pool, mb, mbp) == [] and list(nm) == list(mnm)


def test_gpu_group_cases_at_toy_size(lib, capfd):
    grp = lib.GroupAligner(devices=[0, 0], lib_dir=EMU_DIR, lib_name="libssw_emu.so")
    one = lib.BatchAligner(lib_dir=EMU_DIR, lib_name="libssw_emu.so")
    grp.set_option("latency_cols", 0)
    one.set_option("latency_cols", 0)
    run_group_cases(grp, one, ref_len=500, n_reads=6, n_queries=3, n_targets=9, n_check=20, long_ref=1000, long_len=340, n_long=3)
    grp.close()
    one.close()


def test_group_full_grid_equals_one_engine(lib, capfd):
    reads, refs = _workload(71)
    mat = C.dna_matrix(2, 2)
    one = lib.BatchAligner(lib_dir=EMU_DIR, lib_name="libssw_emu.so")
    one.set_option("latency_cols", 0)
    one.set_sequences(reads, refs)
    pq, pr = np.repeat(np.arange(len(reads)), len(refs)), np.tile(np.arange(len(refs)), len(reads))
    for flag in (0, 0x0f):
        base, base_pool = one.align(mat, 5, 3, 1, flag=flag, filterd=32767, mask_len=40, score_size=2)
        exp, exp_pool, _, _, _ = C.cpu_batch(reads, refs, pq, pr, mat, 5, 3, 1, flag=flag, filterd=32767, mask_len=40, score_size=2, threads=4)
        assert C.compare_records(base, base_pool, exp, exp_pool) == []
        for world in ((5,) if flag == 0 else (3,)):
            grp = lib.GroupAligner(n_devices=world, lib_dir=EMU_DIR, lib_name="libssw_emu.so")
            grp.set_option("latency_cols", 0)          # batch layouts: the one-pair latency path is minutes of emulator time
            assert grp.size == world
            res, pool = grp.align(reads, refs, mat, 5, 3, 1, flag=flag, filterd=32767, mask_len=40, score_size=2)
            _same(res, pool, base, base_pool)
            assert C.compare_records(res, pool, exp, exp_pool) == []
            busy = [t["fill_forward_launches"] > 0 for t in grp.timing()]
            assert sum(busy) >= min(world, 4), (world, busy)      # 9 ragged queries: a long one may leave one of five devices without a block
            grp.close()
    # more devices than queries: empty blocks are skipped
    grp = lib.GroupAligner(devices=[0, 1, 2, 3], lib_dir=EMU_DIR, lib_name="libssw_emu.so")
    grp.set_option("latency_cols", 0)
    res, pool = grp.align(reads[:2], refs, mat, 5, 3, 1, flag=2, mask_len=40, score_size=2)
    one.set_sequences(reads[:2], refs)
    base, base_pool = one.align(mat, 5, 3, 1, flag=2, mask_len=40, score_size=2)
    _same(res, pool, base, base_pool)
    # a prefix of the grid runs as an explicit list
    res, pool = grp.align(reads[:2], refs, mat, 5, 3, 1, flag=2, mask_len=40, score_size=2, n_pairs=4)
    _same(res, pool, base[:4], base_pool)
    grp.close()
    one.close()


def test_group_pair_list_is_cut_by_cells(lib, capfd):
    reads, refs = _workload(72, n_reads=9)
    rng = np.random.default_rng(5)
    mat = C.dna_matrix(2, 2)
    m = 40
    pq = rng.integers(0, len(reads), size=m).astype(np.int32)
    pr = np.sort(rng.integers(0, len(refs), size=m)).astype(np.int32)         # pairs that share a reference are neighbours
    exp, exp_pool, _, _, _ = C.cpu_batch(reads, refs, pq, pr, mat, 5, 3, 1, flag=2, filters=20, mask_len=30, score_size=2, threads=4)
    for world in (2, 3):
        grp = lib.GroupAligner(n_devices=world, lib_dir=EMU_DIR, lib_name="libssw_emu.so")
        grp.set_option("latency_cols", 0)
        res, pool = grp.align(reads, refs, mat, 5, 3, 1, flag=2, filters=20, mask_len=30, score_size=2, pair_query=pq, pair_ref=pr)
        assert C.compare_records(res, pool, exp, exp_pool) == []
        cells = [t["cells_forward"] for t in grp.timing()]
        assert min(cells) > 0 and max(cells) < 2.5 * min(cells), cells         # balanced by DP cells, not by pair count
        grp.close()
    grp = lib.GroupAligner(n_devices=2, lib_dir=EMU_DIR, lib_name="libssw_emu.so")
    with pytest.raises(RuntimeError):
        grp.align(reads, refs, mat, 5, 3, 1, flag=0, pair_query=[0, len(reads)], pair_ref=[0, 0])      # query index out of range
    grp.close()
    assert "out of range" in capfd.readouterr().err


def test_group_text_reverse_complement_and_marked(lib, capfd):
    rng = np.random.default_rng(4243)
    refs = [bytes(rng.choice(np.frombuffer(b"ACGTN", dtype=np.uint8), size=int(n), p=[0.24, 0.24, 0.24, 0.24, 0.04])) for n in (500, 260)]
    comp = {ord("A"): ord("T"), ord("C"): ord("G"), ord("G"): ord("C"), ord("T"): ord("A"), ord("N"): ord("N")}
    reads = []
    for k in range(7):
        r = refs[k % 2]
        a = int(rng.integers(0, len(r) - 90))
        piece = bytearray(r[a: a + int(rng.integers(25, 80))])
        for i in range(len(piece)):
            if rng.random() < 0.08:
                piece[i] = b"ACGT"[int(rng.integers(0, 4))]
        if k % 2:
            piece = bytearray(comp[c] for c in reversed(piece))            # minus-strand reads: found through the reverse complements
        reads.append(bytes(piece))
    table = np.full(128, 4, dtype=np.int8)
    for i, c in enumerate("ACGT"):
        table[ord(c)] = i
        table[ord(c.lower())] = i
    mat = C.dna_matrix(2, 2)
    one = lib.BatchAligner(lib_dir=EMU_DIR, lib_name="libssw_emu.so")
    one.set_option("latency_cols", 0)
    one.set_sequences_text(reads, refs, table, 5, add_reverse_complement=True)
    base, base_pool = one.align(mat, 5, 3, 1, flag=2, mask_len=15, score_size=2)
    mbase, mpool, mnm = one.mark_mismatch(base, base_pool)
    assert int(np.sum(mnm)) > 0
    for world in (1, 3):
        grp = lib.GroupAligner(n_devices=world, lib_dir=EMU_DIR, lib_name="libssw_emu.so")
        grp.set_option("latency_cols", 0)
        res, pool = grp.align(reads, refs, mat, 5, 3, 1, flag=2, mask_len=15, score_size=2, table=table, add_reverse_complement=True)
        assert len(res) == 2 * len(reads) * len(refs)
        _same(res, pool, base, base_pool)
        res, pool, nm = grp.align(reads, refs, mat, 5, 3, 1, flag=2, mask_len=15, score_size=2, table=table, add_reverse_complement=True, marked=True)
        _same(res, pool, mbase, mpool)
        assert list(nm) == list(mnm)
        # the same through an explicit list that mixes plus- and minus-strand queries
        pq = np.array([7, 0, 13, 3, 3, 8, 1], dtype=np.int32)
        pr = np.array([0, 0, 1, 1, 0, 1, 1], dtype=np.int32)
        res, pool, nm = grp.align(reads, refs, mat, 5, 3, 1, flag=2, mask_len=15, score_size=2, table=table, add_reverse_complement=True, marked=True,
                                  pair_query=pq, pair_ref=pr)
        idx = pq.astype(np.int64) * len(refs) + pr
        _same(res, pool, mbase[idx], mpool)
        assert list(nm) == list(mnm[idx])
        grp.close()
    one.close()


def test_front_ends_over_a_group(lib, tmp_path):
    """ssw_batch_cli -g 3 and Aligner::AlignBatch(..., devices = 3) over the emulator build with three emulated devices print
    what the reference's ssw_test / the reference's C++ wrapper printed (frozen goldens)."""
    import json
    link = ["-L" + EMU_DIR, "-l:libssw_emu.so", "-Wl,-rpath," + EMU_DIR, "-lm", "-lz"]
    cli = str(tmp_path / "ssw_batch_cli_emu")
    subprocess.run(["g++", "-O2", "-std=c++17", "-o", cli, os.path.join(C.PKG, "csrc", "ssw_batch_cli.cpp")] + link, check=True)
    with open(os.path.join(C.GOLDEN, "consumer_outputs.json")) as f:
        G = json.load(f)
    for name, text in G["files"].items():
        (tmp_path / name).write_text(text)
    env = dict(os.environ, SSW_EMU_DEVICES="3")
    n = 0
    for run in G["runs"]:
        if run["exe"] != "ssw_test" or "1k.fa" in run["args"]:
            continue                       # 100 reads x two strands with CIGARs: minutes on the emulator; tests/test_gpu_parity_group.py runs them with -g
        out = subprocess.run([cli, "-g", "3"] + run["args"], capture_output=True, text=True, timeout=900, cwd=str(tmp_path), env=env)
        assert out.returncode == 0, out.stderr[-500:]
        assert "\n".join(out.stdout.splitlines()) == run["stdout"], run["args"]
        n += 1
    assert n >= 6
    # more GPUs than the box has: refused, nothing printed
    out = subprocess.run([cli, "-g", "4", "r1.fa", "r1_query.fq"], capture_output=True, text=True, cwd=str(tmp_path), env=env)
    assert out.returncode != 0 and out.stdout == ""
    # Aligner::AlignBatch over a group: the same alignments as on one device
    src = tmp_path / "batch_devices.cpp"
    src.write_text(r"""
#include <stdio.h>
#include <string>
#include <vector>
#include "ssw_cpp.h"
using namespace StripedSmithWaterman;
int main() {
	unsigned long long st = 99;
	auto rnd = [&]() { st = st * 6364136223846793005ull + 1442695040888963407ull; return (unsigned)(st >> 33); };
	std::string ref(900, 'A');
	for (char& c : ref) c = "ACGT"[rnd() % 4];
	std::vector<std::string> qs;
	for (int i = 0; i < 9; ++i) {
		std::string q = ref.substr(rnd() % 700, 30 + rnd() % 120);
		for (char& c : q) if (rnd() % 10 == 0) c = "ACGT"[rnd() % 4];
		if (i == 4) q.clear();                       // empty queries are skipped like Align() skips them
		qs.push_back(q);
	}
	Aligner a;
	a.SetReferenceSequence(ref.c_str(), ref.size());
	Filter f;
	for (int devices : {1, 3}) {
		std::vector<Alignment> res;
		std::vector<uint16_t> rcs;
		if (!a.AlignBatch(qs, f, res, &rcs, 20, devices)) { printf("AlignBatch failed\n"); return 1; }
		for (size_t i = 0; i < res.size(); ++i)
			printf("%d %zu rc=%d s=%d s2=%d r=[%d,%d] q=[%d,%d] mm=%d %s\n", devices == 1 ? 1 : 0, i, (int)rcs[i], (int)res[i].sw_score, (int)res[i].sw_score_next_best,
			       res[i].ref_begin, res[i].ref_end, res[i].query_begin, res[i].query_end, res[i].mismatches, res[i].cigar_string.c_str());
	}
	return 0;
}
""")
    drv = str(tmp_path / "batch_devices")
    subprocess.run(["g++", "-O2", "-std=c++17", "-I" + os.path.join(C.ROOT, "include"), "-o", drv, str(src),
                    os.path.join(C.PKG, "csrc", "ssw_cpp.cpp")] + link, check=True)
    got = subprocess.run([drv], capture_output=True, text=True, timeout=900, env=env)
    assert got.returncode == 0, got.stderr[-500:]
    lines = got.stdout.splitlines()
    one_dev = [l[2:] for l in lines if l.startswith("1 ")]
    three_dev = [l[2:] for l in lines if l.startswith("0 ")]
    assert len(one_dev) == 9 and one_dev == three_dev and sum("s=0 " in l for l in one_dev) == 1


def test_c_example_of_the_batch_api(lib, tmp_path):
    """examples/example_batch.c (plain C99 over include/ssw_batch.h) on the emulator build: the known answer of the reference's
    example programs (score 21, second best 8 at 4, reference 8..21, read 0..14, two edits, 4=1X4=1I5=), the
    reverse-complement symmetry of its pairs, and the same output from one and from two devices."""
    exe = str(tmp_path / "example_batch")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-O1", "-I" + os.path.join(C.ROOT, "include"), "-o", exe,
                    os.path.join(C.ROOT, "examples", "example_batch.c"), "-L" + EMU_DIR, "-l:libssw_emu.so", "-Wl,-rpath," + EMU_DIR, "-lm"], check=True)
    env = dict(os.environ, SSW_EMU_DEVICES="2")
    outs = [subprocess.run([exe, str(n)], capture_output=True, text=True, timeout=300, env=env) for n in (1, 2)]
    assert all(o.returncode == 0 for o in outs), outs[0].stderr[-300:]
    lines = outs[0].stdout.splitlines()
    assert outs[0].stdout == outs[1].stdout and len(lines) == 12
    assert lines[0] == "read 0+ x ref 0: score 21 (second 8 at 4)  ref 8..21  read 0..14  NM 2  4=1X4=1I5="
    tail = lambda l: l.split(": ", 1)[1]
    assert tail(lines[6]) == tail(lines[2]) and tail(lines[7]) == tail(lines[3])          # read 0 minus strand == read 1 (its reverse complement)
    assert subprocess.run([exe, "3"], capture_output=True, env=env).returncode != 0        # more devices than there are
