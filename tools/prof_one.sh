set -u
OUT=gpurun_out
python tools/proj_time.py > $OUT/r2p_plain.log 2>&1 || { tail -5 $OUT/r2p_plain.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:proj_tf32x3 -s 2 -c 1 -f -o $OUT/r2p_proj python tools/proj_time.py > $OUT/r2p_ncu.log 2>&1
echo rc=$?
