"""End-to-end drop-in test: our HS_call_variants (host C++ over libhsgpu, hairsplitter_b200/bin) against the
reference executable compiled from /root/reference (oracle/_ref/HS_call_variants) on the same GFA + reads
+ SAM. The .col, .vcf and error-rate files must be byte-identical (reference run with one thread, whose
contig order ours reproduces)."""
import filecmp
import os
import subprocess

import pytest

import cases
from hairsplitter_b200 import synth

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OURS = os.path.join(ROOT, "hairsplitter_b200", "bin", "HS_call_variants")
REF = os.path.join(ROOT, "oracle", "_ref", "HS_call_variants")


def _run(exe, files, tmp, tag, threads=1, amplicon=0, thr="0.33"):
    gfa, reads, sam = files
    col, vcf, err = [os.path.join(tmp, f"{tag}.{e}") for e in ("col", "vcf", "err")]
    subprocess.run([exe, gfa, reads, sam, str(threads), tmp, err, str(amplicon), "0", col, vcf, thr], check=True,
                   stdout=subprocess.DEVNULL)
    return col, vcf, err


@pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref not built")
@pytest.mark.parametrize("case", ["ont_multi", "hifi_fastq", "edges"])
def test_col_vcf_error_rate_identical_to_reference(tmp_path, case):
    assert os.path.exists(OURS), "build hairsplitter_b200/host first (python -c 'import __graft_entry__ as g; g.build()')"
    if case == "ont_multi":
        chunks = [cases.small_case(seed=91, length=20000, depth=50, mean_len=5000, error=0.06), cases.medium_case(),
                  cases.small_case(seed=5, length=3000, depth=12, mean_len=900, hard=0.4)]
        fastq = False
    elif case == "hifi_fastq":
        chunks = [cases.hifi_case(), cases.small_case(seed=12, eqx=True)]
        fastq = True
    else:
        chunks = cases.ragged_cases() + [cases.deep_case()]
        fastq = False
    for i, c in enumerate(chunks):
        c.name = f"ctg{i}"
    files = synth.write_files(chunks, os.path.join(str(tmp_path), "in"), fastq=fastq)
    ref = _run(REF, files, str(tmp_path), "ref")
    ours = _run(OURS, files, str(tmp_path), "ours", threads=4)
    for a, b in zip(ref, ours):
        assert filecmp.cmp(a, b, shallow=False), (a, b)
    assert os.path.getsize(ref[0]) > 1000


def test_usage_and_version_probe_return_zero():
    """hairsplitter.py's dependency check runs the executable with --version and expects status 0 (:229-239)"""
    r = subprocess.run([OURS, "--version"], stdout=subprocess.PIPE)
    assert r.returncode == 0 and b"Usage" in r.stdout


@pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref not built")
def test_two_gpus_give_the_same_files(tmp_path):
    """contigs sharded over two GPUs (HSGPU_NGPUS=2), heaviest first: same bytes as the reference"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    chunks = [cases.small_case(seed=91, length=20000, depth=50, mean_len=5000, error=0.06), cases.medium_case(),
              cases.hifi_case(), cases.small_case(seed=5, length=3000, depth=12, mean_len=900, hard=0.4)]
    for i, c in enumerate(chunks):
        c.name = f"ctg{i}"
    files = synth.write_files(chunks, os.path.join(str(tmp_path), "in"))
    ref = _run(REF, files, str(tmp_path), "ref")
    os.environ["HSGPU_NGPUS"] = "2"
    try:
        ours = _run(OURS, files, str(tmp_path), "ours", threads=8)
    finally:
        del os.environ["HSGPU_NGPUS"]
    for a, b in zip(ref, ours):
        assert filecmp.cmp(a, b, shallow=False), (a, b)


GLUED = os.path.join(ROOT, "oracle", "_ref", "HS_call_variants_glued")


@pytest.mark.skipif(not (os.path.exists(REF) and os.path.exists(GLUED)), reason="oracle/_ref not built")
def test_reference_main_on_libhsgpu_gives_the_reference_files(tmp_path):
    """The binding INTEGRATION.md describes, compiled and run: the reference's unmodified main() (parsers, OpenMP loop
    over contigs, merge with the automatic SNPs, output_files -- all from libhsref_cv.so) with generate_msa /
    call_variants / keep_only_robust_variants bound to integration/glue_call_variants.cpp, which only speaks the C ABI
    of include/hsgpu.h. Same .col, .vcf and error-rate bytes as the reference executable (ONT contigs with two
    strains, a HiFi contig with the stricter suspect threshold, a contig with hard clips, a contig without SNPs)."""
    chunks = [cases.small_case(seed=91, length=20000, depth=50, mean_len=5000, error=0.06), cases.medium_case(),
              cases.hifi_case(), cases.small_case(seed=5, length=3000, depth=12, mean_len=900, hard=0.4)]
    for i, c in enumerate(chunks):
        c.name = f"ctg{i}"
    files = synth.write_files(chunks, os.path.join(str(tmp_path), "in"))
    ref = _run(REF, files, str(tmp_path), "ref")
    glued = _run(GLUED, files, str(tmp_path), "glued", threads=1)
    for a, b in zip(ref, glued):
        assert filecmp.cmp(a, b, shallow=False), (a, b)
    assert open(ref[0], "rb").read().count(b"SNPS\t") > 20
    # two OpenMP threads, one hsgpu_ctx each: main() stores the contigs in the order its threads finish them
    # (SURVEY.md 8c), so the blocks are compared as a set
    two = _run(GLUED, files, str(tmp_path), "glued2", threads=2)
    assert sorted(open(ref[0], "rb").read().split(b"CONTIG\t")) == sorted(open(two[0], "rb").read().split(b"CONTIG\t"))
    assert sorted(open(ref[1], "rb").read().splitlines()) == sorted(open(two[1], "rb").read().splitlines())
