#!/bin/bash
# Microbenchmark / numerics logs quoted in DESIGN.md (one gpurun call):
#   gpurun --timeout 900 -- 'bash tools/evidence_r03.sh'
# -> gpurun_out/r03_*.log, copied into profiles/ by hand (they are small text files).
set -u
OUT=gpurun_out
mkdir -p $OUT
{ echo "# tools/umma_bench: tcgen05.mma kind::tf32 issue rate by operand layout"; ./tools/umma_bench; } > $OUT/r03_umma_bench.log 2>&1
{ echo "# tools/mma_bench: mma.sync TF32 rate"; ./tools/mma_bench; } > $OUT/r03_mma_bench.log 2>&1
{ echo "# tools/gather_bench: random row gathers, LDG.128 vs cp.async.bulk (arxiv-shaped: 256-byte rows out of 43 MB)"; ./tools/gather_bench 169343 1166173 64;
  echo "# VOC-shaped: 512-byte rows out of 251 MB"; ./tools/gather_bench 490800 2770000 128; } > $OUT/r03_gather_bench.log 2>&1
python tools/tc_error.py > $OUT/r03_tc_error.log 2>&1
python tools/time_tc.py 10 > $OUT/r03_time_tc.log 2>&1
python tools/time_pattern.py > $OUT/r03_time_pattern.log 2>&1
python tools/proj_time.py > $OUT/r03_proj_time.log 2>&1
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $OUT/r03_smoke.log 2>&1
tail -2 $OUT/r03_smoke.log
