"""The drop-in seam as Nextflow drives it (polycracker/polycracker.nf:6-16, 31-110, 115-354).

The reference's pipeline is `config_polyCRACKER.txt` -> Groovy `findValue` -> string-built `polycracker <sub-command> ...`
lines run in the working directory, with files as the only hand-off.  These tests
  * port `findValue` and the derived names of nf:87-95 to Python and parse both the reference's test config and the
    config this repo ships (config_polyCRACKER.b200.txt: identical except perfect_mode = 1),
  * take the LITERAL command lines of the hot-path processes (tests/golden/nf_command_lines.txt, extracted from the .nf by
    tests/golden/make_nf_lines.py), substitute the parsed values like Nextflow does, and run them through the console
    script's click group with click.testing.CliRunner -- including `--genome=` with the near-empty value Nextflow passes
    (nf:347), the `bbmap.sh ref=` index build (nf:201; tools/bbmap.sh is the shipped stand-in) and the query-sharded
    mapping of runBlastParallel = 1 (nf:247-295: k-mer FASTA split by records, one blast_kmers per piece, SAMs concatenated),
  * compare every artefact with the fixtures the reference's own stage code produced (tests/golden/case_*).
"""
import os
import pickle
import re
import shlex
import shutil
import subprocess

import numpy as np
import pytest
import scipy.sparse as sps
from click.testing import CliRunner

from oracle import py_oracle
from polycracker_b200 import cli
from tests.test_golden import GOLD, _assert_same_npz, _cfg, _fasta, _merged, _write_sam

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF_DIR = "/root/reference/polycracker"
OUR_CONFIG = os.path.join(ROOT, "polycracker-unofficial-mirror_b200", "config_polyCRACKER.b200.txt")
NF_LINES = {int(l.split("\t")[0]): l.rstrip("\n").split("\t", 1)[1] for l in open(os.path.join(GOLD, "nf_command_lines.txt")) if l.strip()}


# ---------------------------------------------------------------------------------------------- Groovy semantics
def minus(s, sub):
    """Groovy `String - String`: the first occurrence is removed."""
    i = s.find(sub)
    return s if i < 0 else s[:i] + s[i + len(sub):]


def find_value(config_text, key):
    """polycracker.nf:6-16: first line that starts with `key` (which carries its trailing blank), minus the key, minus the
    first '= ', minus the first blank."""
    for line in config_text.split("\n"):
        if line.startswith(key):
            return minus(minus(minus(line, key), "= "), " ")
    return None


def nf_environment(config_text, working_dir):
    """Variables of polycracker.nf:31-110 that the hot-path command lines use."""
    names = dict(blastPath="blastPath ", kmercountPath="kmercountPath ", fastaPath="fastaPath ", blastMemory="blastMemory ",
                 threads="threads ", genome="genome ", splitLength="splitFastaLineLength ", kmerLength="kmerLength ",
                 k_search_length="k_search_length ", kmer_low_count="kmer_low_count ", minChunkSize="minChunkSize ",
                 removeNonChunk="removeNonChunk ", minChunkThreshold="minChunkThreshold ", lowMemory="lowMemory ",
                 highBool="use_high_count ", tfidf="tfidf ", perfect_mode="perfect_mode ", kmer_high_count="kmer_high_count ",
                 sampling_sensitivity="sampling_sensitivity ", preFilter="preFilter ")
    env = {var: find_value(config_text, key) for var, key in names.items()}
    flags = {var: int(find_value(config_text, key)) for var, key in
             dict(BB="BB ", splitFast="splitFasta ", writeKmer="writeKmer ", fromFasta="kmer2Fasta ", original="original ",
                  writeBlast="writeBlast ", runBlastParallel="runBlastParallel ", b2b="blast2bed ",
                  genMat="generateClusteringMatrix ").items()}
    genome = env["genome"]
    env["genomeSplitName"] = minus(minus(minus(genome, "_split"), ".fasta"), ".fa") + "_split.fa"     # nf:87
    env["genomeFullPath"] = env["fastaPath"] + env["genomeSplitName"]
    env["kmercountName"] = minus(env["genomeSplitName"], ".fa") + ".kcount" + ".fa"                    # nf:94
    env["blastName"] = minus(env["kmercountName"], ".fa") + ".BLASTtsv.txt"                            # nf:95
    env["blastMemStr"] = "export _JAVA_OPTIONS='-Xms5G -Xmx" + env["blastMemory"] + "G'"
    env["originalGenome"] = env["fastaPath"] + genome if flags["original"] == 1 else env["genomeFullPath"]
    env["workingDir"] = working_dir
    return env, flags


def nf_format(template, env, **extra):
    vals = dict(env, **extra)
    out = re.sub(r"\$\{(\w+)\}", lambda m: vals[m.group(1)], template)
    return out.replace("\\$cwd", vals.get("cwd", "."))


def run_cli(line):
    """One `polycracker ...` line as the bash script of a Nextflow process would run it."""
    argv = shlex.split(line)
    assert argv[0] == "polycracker"
    res = CliRunner().invoke(cli.polycracker, argv[1:], catch_exceptions=False)
    assert res.exit_code == 0, (line, res.output)
    return res


# ---------------------------------------------------------------------------------------------- CPU: parsing + formatting
def test_find_value_port_on_the_shipped_config():
    text = open(OUR_CONFIG).read()
    env, flags = nf_environment(text, "/work")
    assert env["fastaPath"] == "./test_data/test_fasta_files/" and env["genome"] == "algae.fa"
    assert env["genomeSplitName"] == "algae_split.fa" and env["kmercountName"] == "algae_split.kcount.fa"
    assert env["blastName"] == "algae_split.kcount.BLASTtsv.txt"
    assert env["splitLength"] == "50000" and env["kmerLength"] == "26" and env["kmer_low_count"] == "30"
    assert env["perfect_mode"] == "1"                          # the one value that differs from the reference's test config
    assert flags == dict(BB=1, splitFast=1, writeKmer=1, fromFasta=1, original=0, writeBlast=1, runBlastParallel=0, b2b=1, genMat=1)
    # 'kmer2Fasta ' must not be shadowed by 'kmer_low_count ' etc.: startsWith(key + blank) is what keeps the keys apart
    assert find_value(text, "kmer ") is None and find_value(text, "splitFasta ") == "1"
    line = nf_format(NF_LINES[347], env, genomeName="")
    assert "--genome= " in line and "--genome_split_name=algae_split.fa" in line and "--original_genome=algae.fa" in line
    assert nf_format(NF_LINES[318], env).endswith("--blast_file=./blast_files/algae_split.kcount.BLASTtsv.txt --bb=1 --low_memory=0")


@pytest.mark.skipif(not os.path.isdir(REF_DIR), reason="reference tree not mounted (GPU box)")
def test_shipped_config_and_command_lines_equal_the_reference():
    def settings(text):
        return [l for l in text.split("\n") if l.strip() and not l.startswith("#")]
    ref = settings(open(os.path.join(REF_DIR, "test_data", "config_polyCRACKER.txt")).read())
    ours = settings(open(OUR_CONFIG).read())
    assert len(ref) == len(ours)
    diff = [(a, b) for a, b in zip(ref, ours) if a != b]
    assert diff == [("perfect_mode = 0", "perfect_mode = 1")]
    nf = open(os.path.join(REF_DIR, "polycracker.nf")).read().split("\n")
    for no, template in NF_LINES.items():
        assert nf[no - 1].strip() == template, no
    # every config key the hot-path lines need parses the same from both files
    ref_env, ref_flags = nf_environment(open(os.path.join(REF_DIR, "test_data", "config_polyCRACKER.txt")).read(), "/w")
    our_env, our_flags = nf_environment(open(OUR_CONFIG).read(), "/w")
    assert ref_flags == our_flags and {k: v for k, v in ref_env.items() if k != "perfect_mode"} == \
        {k: v for k, v in our_env.items() if k != "perfect_mode"}


def test_cli_answers_and_bbmap_stand_in():
    r = CliRunner()
    res = r.invoke(cli.polycracker, ["--help"])
    assert res.exit_code == 0
    for cmd in ("splitFasta", "writeKmerCount", "kmer2Fasta", "blast_kmers", "blast2bed", "generate_Kmer_Matrix"):
        assert cmd in res.output
        assert r.invoke(cli.polycracker, [cmd, "--help"]).exit_code == 0
    assert r.invoke(cli.polycracker, ["blast_kmers", "--no_such_option"]).exit_code != 0
    stub = os.path.join(ROOT, "tools", "bbmap.sh")
    ok = subprocess.run([stub, "ref=./fasta_files/algae_split.fa", "k=13"], capture_output=True, text=True)
    assert ok.returncode == 0 and "skipped" in ok.stderr
    bad = subprocess.run([stub, "in=query.fa", "out=results.sam", "perfectmode=t"], capture_output=True, text=True)
    assert bad.returncode != 0 and "only" in bad.stderr


def _case_config(case, run_blast_parallel=0):
    """The shipped config re-pointed at a golden case (paths, genome, chunk length, k, thresholds of its params.txt)."""
    cfg = _cfg(case)
    text = open(OUR_CONFIG).read()
    repl = {"fastaPath": "./fasta_files/", "genome": "genome.fa", "splitFastaLineLength": cfg["chunk_size"],
            "kmerLength": cfg["k"], "kmer_low_count": cfg["kmer_low_count"], "use_high_count": cfg["high_bool"],
            "kmer_high_count": cfg["kmer_high_count"], "sampling_sensitivity": cfg["sampling"],
            "minChunkSize": cfg["chunk_size"], "removeNonChunk": cfg["remove_non_chunk"], "tfidf": cfg["tfidf"],
            "runBlastParallel": run_blast_parallel}
    for key, val in repl.items():
        text, n = re.subn(r"(?m)^%s = .*$" % key, "%s = %s" % (key, val), text)
        assert n == 1, key
    return text


@pytest.mark.parametrize("case", ["case_k23"])
def test_cpu_stages_through_the_literal_nextflow_lines(case, tmp_path, monkeypatch):
    """splitFasta, kmer2Fasta and blast2bed are host-only: run them through the nf lines without a GPU."""
    g = os.path.join(GOLD, case)
    monkeypatch.chdir(tmp_path)
    open("config_polyCRACKER.txt", "w").write(_case_config(case))
    env, flags = nf_environment(open("config_polyCRACKER.txt").read(), str(tmp_path))
    os.makedirs("fasta_files"); os.makedirs("blast_files"); os.makedirs("kmercount_files")
    shutil.copy(os.path.join(g, "fasta_files", "genome.fa"), "fasta_files/genome.fa")
    run_cli(nf_format(NF_LINES[130], env, genomeName=minus(env["genome"], "\n")))
    assert open("correspondence.bed").read() == open(os.path.join(g, "correspondence.bed")).read()
    assert open(env["genomeFullPath"]).read() == open(os.path.join(g, "fasta_files", "genome_split.fa")).read()
    shutil.copy(os.path.join(g, "kmercount_files", "genome_split.kcount"), "kmercount_files/genome_split.kcount")
    run_cli(nf_format(NF_LINES[200], env))
    assert open(env["kmercountPath"] + env["kmercountName"]).read() == open(os.path.join(g, "kmercount_files", "genome_split.kcount.fa")).read()
    kmers = [l.strip() for l in open(env["kmercountPath"] + env["kmercountName"]).readlines()[1::2]]
    _write_sam(env["blastPath"] + env["blastName"], py_oracle.blast_kmers(kmers, _fasta(env["genomeFullPath"])))
    run_cli(nf_format(NF_LINES[318], env))
    assert _merged("blasted_merged.bed") == _merged(os.path.join(g, "blasted_merged.bed"))


# ---------------------------------------------------------------------------------------------- GPU: the whole seam
@pytest.mark.gpu
@pytest.mark.parametrize("parallel", [0, 1])
@pytest.mark.parametrize("case", ["case_k23", "case_k11_high"])
def test_pipeline_through_the_literal_nextflow_lines(case, parallel, tmp_path, monkeypatch):
    g = os.path.join(GOLD, case)
    monkeypatch.chdir(tmp_path)
    monkeypatch.setenv("PATH", os.path.join(ROOT, "tools") + os.pathsep + os.environ["PATH"])
    open("config_polyCRACKER.txt", "w").write(_case_config(case, parallel))
    env, flags = nf_environment(open("config_polyCRACKER.txt").read(), str(tmp_path))
    assert flags["runBlastParallel"] == parallel
    os.makedirs("fasta_files"); os.makedirs("blast_files")
    shutil.copy(os.path.join(g, "fasta_files", "genome.fa"), "fasta_files/genome.fa")
    # splitFastaProcess (nf:115-141); its stdout is what later processes receive as genomeName
    res = run_cli(nf_format(NF_LINES[130], env, genomeName=minus(env["genome"], "\n")))
    genome_name = res.output.strip()
    run_cli(nf_format(NF_LINES[161], env))                     # writeKmerCount (nf:143-172)
    assert open("kmercount_files/genome_split.kcount").read() == open(os.path.join(g, "kmercount_files", "genome_split.kcount")).read()
    run_cli(nf_format(NF_LINES[200], env))                     # kmer2Fasta + the bbmap.sh index build (nf:200-201)
    bb = nf_format(NF_LINES[201], env)
    assert subprocess.run(["bash", "-c", bb], capture_output=True).returncode == 0
    kfa = env["kmercountPath"] + env["kmercountName"]
    assert open(kfa).read() == open(os.path.join(g, "kmercount_files", "genome_split.kcount.fa")).read()
    # BlastOff (nf:247-295): one process per piece of the k-mer FASTA (splitFasta(by: N)), results concatenated by collectFile
    lines = open(kfa).read().split("\n")
    records = ["\n".join(lines[i:i + 2]) for i in range(0, len(lines) - 1, 2) if lines[i].startswith(">")]
    by = 200 if parallel else len(records)
    sam = env["blastPath"] + env["blastName"]
    open(sam, "w").close()
    for p_i, a in enumerate(range(0, max(len(records), 1), by)):
        work = tmp_path / ("work_%d" % p_i)
        work.mkdir()
        (work / "query.fa").write_text("\n".join(records[a:a + by]) + "\n")
        run_cli(nf_format(NF_LINES[274], env, cwd=str(work)))
        with open(sam, "a") as out:
            out.write(open("results.sam").read())              # mv results.sam blast_result; collectFile(name: blastPath + blastName)
        os.remove("results.sam")
        if os.path.exists("results.sam.pcb200.npz"):
            if not parallel:
                os.rename("results.sam.pcb200.npz", sam + ".pcb200.npz")
            else:
                os.remove("results.sam.pcb200.npz")            # pieces: only the concatenated text describes the whole
    run_cli(nf_format(NF_LINES[318], env))                     # blast2bed (nf:297-326)
    assert _merged("blasted_merged.bed") == _merged(os.path.join(g, "blasted_merged.bed"))
    assert os.path.exists("blasted_merged.bed.pcb200.npz") == (not parallel)
    run_cli(nf_format(NF_LINES[347], env, genomeName=genome_name))     # genClusterMatrix (nf:328-354), --genome= near-empty
    _assert_same_npz("clusteringMatrix.npz", os.path.join(g, "clusteringMatrix.npz"))
    _assert_same_npz("clusteringMatrix.tfidf.npz", os.path.join(g, "clusteringMatrix.tfidf.npz"), exact=False)
    for name in ("kmers.p", "scaffolds.p", "originalScaffolds.p"):
        assert pickle.load(open(name, "rb")) == pickle.load(open(os.path.join(g, name), "rb")), name
    for name in ("rowNames.txt", "colNames.txt"):
        assert open(name).read() == open(os.path.join(g, name)).read(), name
    assert sps.load_npz("clusteringMatrix.npz").format == "csc"
    # perfect_mode = 0 (the reference's shipped test config) is refused loudly, not silently different
    bad = CliRunner().invoke(cli.polycracker, shlex.split(nf_format(NF_LINES[274], dict(env, perfect_mode="0"), cwd=str(tmp_path / "work_0")))[1:])
    assert bad.exit_code != 0 and "perfect_mode" in bad.output
