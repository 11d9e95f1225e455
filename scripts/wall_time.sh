TIMEFORMAT="wall %R s"
python - <<'PY'
import sys; sys.path.insert(0,'.')
from hairsplitter_b200 import synth
chunks, info = synth.make_config(2, scale=1.0, seed=2)
synth.write_files(chunks, '/tmp/w2')
PY
HS_TIMING=1 hairsplitter_b200/bin/HS_call_variants /tmp/w2.gfa /tmp/w2.fasta /tmp/w2.sam 16 /tmp /tmp/o.err 0 0 /tmp/o.col /tmp/o.vcf 0.33 > /dev/null 2>&1  # warm the file cache
echo "--- contexts first (HS_CTX_FIRST=1)"
for i in 1 2; do time env HS_CTX_FIRST=1 HS_TIMING=1 hairsplitter_b200/bin/HS_call_variants /tmp/w2.gfa /tmp/w2.fasta /tmp/w2.sam 16 /tmp /tmp/o.err 0 0 /tmp/o.col /tmp/o.vcf 0.33 2>&1 >/dev/null; done
echo "--- background (default)"
for i in 1 2; do time env HS_TIMING=1 hairsplitter_b200/bin/HS_call_variants /tmp/w2.gfa /tmp/w2.fasta /tmp/w2.sam 16 /tmp /tmp/o.err 0 0 /tmp/o.col /tmp/o.vcf 0.33 2>&1 >/dev/null; done
echo "--- background, 8 threads"
time env HS_TIMING=1 hairsplitter_b200/bin/HS_call_variants /tmp/w2.gfa /tmp/w2.fasta /tmp/w2.sam 8 /tmp /tmp/o.err 0 0 /tmp/o.col /tmp/o.vcf 0.33 2>&1 >/dev/null
