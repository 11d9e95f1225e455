"""Generates the committed golden fixtures of the HS_separate_reads path from the UNMODIFIED reference (needs
oracle/_ref, i.e. the build container with /root/reference). Run:  python tests/golden/make_golden_sr.py

sr_<case>.col  what the reference's HS_call_variants wrote for a small seeded synthetic input
sr_<case>.gro  what the reference's HS_separate_reads wrote for it with std::random_device pinned
               (oracle/ref_pin_rng.cpp, constant = oracle.pyoracle.PIN_SEED); arguments in CASES below
"""
import gzip
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))

# name -> (case factory arguments, error rate, low_memory, rarest, amplicon)
CASES = {
    "ont": (dict(seed=201, length=12000, depth=40, mean_len=4000, error=0.06), "0.06", "0", "0", "0"),
    "lowmem": (dict(seed=202, length=8000, depth=35, mean_len=3000, error=0.06), "0.06", "1", "0", "0"),
    "amplicon": (dict(seed=203, length=2000, depth=150, mean_len=1900, error=0.05), "0.05", "0", "0", "1"),
}


def main():
    import cases
    from hairsplitter_b200 import synth
    ref_cv = os.path.join(ROOT, "oracle", "_ref", "HS_call_variants")
    ref_sr = os.path.join(ROOT, "oracle", "_ref", "HS_separate_reads_pinned")
    for name, (kw, err, low, rare, amp) in CASES.items():
        tmp = tempfile.mkdtemp(prefix="hs_golden_")
        try:
            cb = cases.small_case(**kw)
            cb.name = "ctg0"
            gfa, reads, sam = synth.write_files([cb], os.path.join(tmp, "in"))
            col, gro = os.path.join(tmp, "a.col"), os.path.join(tmp, "a.gro")
            subprocess.run([ref_cv, gfa, reads, sam, "1", tmp, os.path.join(tmp, "err"), amp, "0", col,
                            os.path.join(tmp, "a.vcf"), "0.33"], check=True, stdout=subprocess.DEVNULL)
            subprocess.run([ref_sr, col, "1", err, os.path.join(tmp, "no_ploidy"), low, rare, amp, gro, "0"], check=True,
                           stdout=subprocess.DEVNULL)
            for src, ext in ((col, "col"), (gro, "gro")):
                with open(src, "rb") as f, gzip.GzipFile(os.path.join(HERE, f"sr_{name}.{ext}.gz"), "wb", mtime=0) as g:
                    shutil.copyfileobj(f, g)
            n_groups = sum(1 for l in open(gro) if l.startswith("GROUP"))
            print(name, os.path.getsize(col), "bytes of .col,", n_groups, "GROUP lines")
        finally:
            shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()
