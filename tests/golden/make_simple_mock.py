"""Packs the reference's own test fixture test/simple_mock (assembly.gfa: 4 segments and 2 links, one of them
reverse; mock_reference.fasta: the 3 haplotypes the mock reads are meant to come from) into
tests/golden/simple_mock.npz, so that BASELINE configs[0] can be realised on the GPU box, where /root/reference
does not exist. `mock_reads.fasta` is absent from the reference repository, so the reads are
simulated from the three haplotypes by hairsplitter_b200/synth.py.

Run in the build container:  python tests/golden/make_simple_mock.py
"""
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/test/simple_mock"
CODE = np.full(256, 3, np.uint8)
for i, ch in enumerate(b"ACGT"):
    CODE[ch] = i
    CODE[ch + 32] = i


def pack2(codes):
    """2-bit codes, 4 per byte (base j in bits 2*(j%4) of byte j//4)"""
    n = codes.shape[0]
    pad = np.zeros((-n) % 4, np.uint8)
    c = np.concatenate([codes, pad]).reshape(-1, 4)
    return (c[:, 0] | (c[:, 1] << 2) | (c[:, 2] << 4) | (c[:, 3] << 6)).astype(np.uint8)


def main():
    names, seqs, links = [], [], []
    for line in open(os.path.join(SRC, "assembly.gfa")):
        f = line.rstrip("\n").split("\t")
        if f[0] == "S":
            names.append(f[1])
            seqs.append(CODE[np.frombuffer(f[2].encode(), np.uint8)])
        elif f[0] == "L":
            links.append("\t".join(f))
    haps = []
    for line in open(os.path.join(SRC, "mock_reference.fasta")):
        if not line.startswith(">"):
            haps.append(CODE[np.frombuffer(line.strip().encode(), np.uint8)])
    # where every segment sits on the haplotypes (they differ from each other by substitutions only)
    offsets = []
    for s in seqs:
        best = (-1, 1 << 30)
        if s.shape[0] >= 1000:
            cand = set()
            for h in haps:  # a probe may carry a variant site: several probes, every haplotype
                hb = h.tobytes()
                for p0 in range(0, s.shape[0] - 64, max(64, s.shape[0] // 16)):
                    at = hb.find(s[p0:p0 + 64].tobytes())
                    if at >= p0:
                        cand.add(at - p0)
            for at in cand:
                if at + s.shape[0] <= haps[0].shape[0]:
                    d = int(sum(int((h[at:at + s.shape[0]] != s).sum()) for h in haps))
                    if d < best[1]:
                        best = (at, d)
            print("segment of", s.shape[0], "at", best[0], "hamming distance to the three haplotypes", best[1])
        offsets.append(best[0])
    out = {"names": np.array(names), "links": np.array(links), "offsets": np.array(offsets, np.int64),
           "lengths": np.array([s.shape[0] for s in seqs], np.int64),
           "hap_length": np.array([h.shape[0] for h in haps], np.int64)}
    for i, s in enumerate(seqs):
        out[f"seg{i}"] = pack2(s)
    for i, h in enumerate(haps):
        out[f"hap{i}"] = pack2(h)
    np.savez_compressed(os.path.join(HERE, "simple_mock.npz"), **out)
    print(names, [s.shape[0] for s in seqs], offsets, links)


if __name__ == "__main__":
    main()
