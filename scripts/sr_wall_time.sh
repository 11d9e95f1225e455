# phase timing of HS_separate_reads over several fresh processes (diagnostic)
TIMEFORMAT="wall %R s"
python - <<'PY'
import sys; sys.path.insert(0,'.')
from hairsplitter_b200 import synth
chunks, info = synth.make_config(2, scale=1.0, seed=2)
synth.write_files(chunks, '/tmp/w2')
PY
hairsplitter_b200/bin/HS_call_variants /tmp/w2.gfa /tmp/w2.fasta /tmp/w2.sam 16 /tmp /tmp/o.err 0 0 /tmp/o.col /tmp/o.vcf 0.33 > /dev/null 2>&1
for i in 1 2 3 4 5 6; do
  time env HS_TIMING=1 HSGPU_TIMING=1 HS_PIN_SEED=12345 hairsplitter_b200/bin/HS_separate_reads /tmp/o.col 16 0.1 /tmp/no_ploidy 0 0 0 /tmp/o.gro 0 2>&1 >/dev/null | grep -v "windows 2500"
done
