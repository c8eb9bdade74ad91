"""tools/: the side-by-side harness against the real reference.  The JVM is absent here, so only the
pieces that do not need it are exercised: the corpus writer, the .result comparator, and the script's refusal."""
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))

import compare_results as cr  # noqa: E402


def _entry(doc, fn, hdr, direction, score):
    return {"docid": doc, "rank": 0, "filename": fn, "header": hdr, "sequence_direction": direction, "taxon_hierarchy": "", "score": score}


def _line(q, typ, hits, rank="unknown", name=""):
    return {"query_header": q, "query": "ACGT", "result": hits, "type": typ, "taxon_rank": rank, "taxon_name": name}


def test_compare_results_matches_by_document_fields_not_doc_ids():
    ref = {">r1": _line(">r1", "CLASSIFIED", [_entry(7, "a.fna", "A", "forward", 0.5)], "species", "s1"),
           ">r2": _line(">r2", "VAGUE", [_entry(3, "a.fna", "A", "forward", 0.25), _entry(9, "b.fna", "B", "reverse", 0.25)]),
           ">r3": _line(">r3", "UNKNOWN", [])}
    ours = json.loads(json.dumps(ref))
    ours[">r1"]["result"][0]["docid"] = 0                        # Lucene's doc ids are not canonical
    ours[">r2"]["result"] = list(reversed(ours[">r2"]["result"]))
    ours[">r2"]["result"][0]["score"] = 0.25 * (1 + 5e-6)        # inside the 1e-5 relative tolerance
    rep = cr.compare(ref, ours)
    assert rep["disagreements"] == [] and rep["missing"] == [] and rep["extra"] == [] and rep["score_mismatch"] == 0
    assert 4e-6 < rep["max_rel"] < 6e-6


def test_compare_results_lists_every_disagreement():
    ref = {">r1": _line(">r1", "CLASSIFIED", [_entry(1, "a.fna", "A", "forward", 0.5)], "species", "s1"),
           ">r2": _line(">r2", "VAGUE", [_entry(3, "a.fna", "A", "forward", 0.25), _entry(9, "b.fna", "B", "reverse", 0.25)]),
           ">r3": _line(">r3", "CLASSIFIED", [_entry(1, "a.fna", "A", "forward", 0.75)]),
           ">r4": _line(">r4", "UNKNOWN", [])}
    ours = json.loads(json.dumps(ref))
    ours[">r1"]["taxon_name"] = "s2"                             # label
    ours[">r2"]["result"] = ours[">r2"]["result"][:1]           # hit set (a broken tie)
    ours[">r2"]["type"] = "CLASSIFIED"
    ours[">r3"]["result"][0]["score"] = 0.75 * (1 + 1e-4)        # score outside the tolerance
    del ours[">r4"]                                              # dropped read
    ours[">r5"] = _line(">r5", "UNKNOWN", [])                    # a read the reference dropped
    rep = cr.compare(ref, ours)
    assert sorted(d["query_header"] for d in rep["disagreements"]) == [">r1", ">r2", ">r3"]
    assert rep["missing"] == [">r4"] and rep["extra"] == [">r5"] and rep["score_mismatch"] == 1


def test_compare_results_cli_exit_status(tmp_path):
    ref = [_line(">r%d" % i, "CLASSIFIED", [_entry(i, "a.fna", "A", "forward", 0.5)], "species", "s") for i in range(20000)]
    a, b = tmp_path / "ref.result", tmp_path / "gpu.result"
    a.write_text("".join(json.dumps(x) + "\n" for x in ref))
    ref[5]["type"] = "VAGUE"                                     # 1 of 20,000 reads: 99.995 % agreement
    b.write_text("".join(json.dumps(x) + "\n" for x in ref))
    exe = [sys.executable, os.path.join(ROOT, "tools", "compare_results.py")]
    p = subprocess.run(exe + [str(a), str(b)], capture_output=True, text=True)
    assert p.returncode == 0, p.stdout
    head = json.loads(p.stdout.splitlines()[0])
    assert head["disagreements"] == 1 and head["agreement"] > 0.9999
    assert json.loads(p.stdout.splitlines()[1])["query_header"] == ">r5"
    ref[6]["type"] = "VAGUE"; ref[7]["type"] = "VAGUE"            # 3 of 20,000: below 99.99 %
    b.write_text("".join(json.dumps(x) + "\n" for x in ref))
    assert subprocess.run(exe + [str(a), str(b)], capture_output=True).returncode == 1


def test_parity_corpus_and_script_without_jvm(tmp_path):
    w = tmp_path / "w"
    subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "make_parity_corpus.py"), str(w), "--genomes", "4",
                           "--genome-bp", "5000", "--reads", "50"], stdout=subprocess.DEVNULL)
    assert sorted(os.listdir(w / "ref")) == sorted(["%02d.fna" % g for g in range(4)] + ["%02d.fna.taxd" % g for g in range(4)])
    reads = (w / "reads" / "sample.fa").read_text().splitlines()
    assert len(reads) == 100 and reads[0].startswith(">read0|syn|") and set(reads[1]) <= set("ACGT")
    c1, c2 = json.loads((w / "conf_ref.json").read_text()), json.loads((w / "conf_gpu.json").read_text())
    assert set(c1) == {"index_path", "kmer_size", "kmer_skips", "min_strand_kmer", "query_term_min_should_match", "worker_threads",
                       "index_ram_buffer", "query_algorithm", "scoring_algorithm"}          # the reference's mapper rejects other keys
    assert {k: v for k, v in c1.items() if k != "index_path"} == {k: v for k, v in c2.items() if k != "index_path"}
    lin = json.loads((w / "ref" / "02.fna.taxd").read_text())
    assert [t["rank"] for t in lin] == ["no rank", "species", "genus"]
    if shutil.which("java") is None or shutil.which("javac") is None:
        p = subprocess.run(["bash", os.path.join(ROOT, "tools", "ref_lucene_run.sh"), "/nonexistent", str(tmp_path / "x")],
                           capture_output=True, text=True)
        assert p.returncode == 3 and "unpinned" in p.stderr


import pytest  # noqa: E402


@pytest.mark.gpu
def test_parity_corpus_through_the_cli_and_comparator(tmp_path):
    """Steps 4-5 of tools/ref_lucene_run.sh (this library's side): the CLI indexes and classifies the parity corpus, the
    comparator reads its .result file (compared with itself: full agreement), and the corpus produces both CLASSIFIED
    and VAGUE reads so that a run against the real reference exercises the tie handling."""
    w = tmp_path / "w"
    subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "make_parity_corpus.py"), str(w), "--genomes", "8",
                           "--genome-bp", "60000", "--reads", "600"], stdout=subprocess.DEVNULL)
    cli = os.path.join(ROOT, "biospectra_b200", "biospectra_gpu")
    subprocess.check_call([cli, "index", "-j", str(w / "conf_gpu.json"), "-r", str(w / "ref")], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    subprocess.check_call([cli, "lclassify", "-j", str(w / "conf_gpu.json"), "-in", str(w / "reads"), "-out", str(w / "out_gpu")],
                          stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    res = w / "out_gpu" / "sample.fa.result"
    got = cr.load(str(res))
    assert len(got) == 600
    types = {d["type"] for d in got.values()}
    assert "CLASSIFIED" in types and "VAGUE" in types, types
    assert any(d["type"] == "CLASSIFIED" and len(d["result"]) >= 2 and d["taxon_rank"] == "species" for d in got.values())
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "compare_results.py"), str(res), str(res)], capture_output=True, text=True)
    assert p.returncode == 0 and json.loads(p.stdout.splitlines()[0])["agreement"] == 1.0
