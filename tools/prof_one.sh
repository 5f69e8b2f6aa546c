set -u
OUT=gpurun_out
python bench.py --workload pattern-gt --profile > $OUT/r2f_plain.log 2>&1 || { tail -5 $OUT/r2f_plain.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:gt_dense -s 1 -c 1 -f -o $OUT/r2f_dense python bench.py --workload pattern-gt --profile > $OUT/r2f_ncu.log 2>&1
echo rc=$?
ls -la $OUT
