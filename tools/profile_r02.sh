#!/bin/bash
# ncu evidence of round 2, one gpurun call (all ncu runs of a call count as one):
#   gpurun --timeout 1500 -- 'bash tools/profile_r02.sh'        (TAG=r03 by default: file prefix)
# For every GPU workload: the plain run first (must exit 0), then (a) a metrics pass over every
# conv kernel launch (duration + DRAM bytes; ours and the reference's), (b) one --set full capture
# of ONE timed step of our kernels (the launches after the warm-up step).  tools/make_traffic.py
# turns the CSVs into profiles/${TAG}_traffic.json and the .ncu-rep files into
# profiles/${TAG}_ncu_full_<workload>.txt.  gpurun brings back at most 64 MiB: keep the captures small.
set -u
TAG=${TAG:-r03}
OUT=gpurun_out
mkdir -p $OUT
OURS='regex:gat_fwd|gat_bwd|dot_fwd|gt_bwd|gt_block|gt_dense|block_attn_dense'
ALL='regex:gat_fwd|gat_bwd|dot_fwd|gt_bwd|gt_block|gt_dense|block_attn_dense|fused_|sddmm|spmm|softMax|softmax|mhsddmm|mhspmm'
for W in ${WORKLOADS:-arxiv-gat pattern-gt voc-gt reddit-gt}; do
  python bench.py --workload $W --profile --profile-ref > $OUT/${TAG}_plain_$W.log 2>&1 || { echo "plain run failed: $W"; tail -5 $OUT/${TAG}_plain_$W.log; continue; }
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
      -k "$ALL" --csv --log-file $OUT/${TAG}_traffic_$W.csv python bench.py --workload $W --profile --profile-ref > $OUT/${TAG}_ncu1_$W.log 2>&1
  echo "traffic $W rc=$?"
  # launches per step: 6 for the staged GAT path (3 staged + 3 big-tile launches), 3 for GT
  case $W in arxiv-gat) SKIP=6; CNT=6; SRC="--import-source on";; pattern-gt) SKIP=4; CNT=4; SRC="--import-source on";; *) SKIP=3; CNT=3; SRC="";; esac
  python bench.py --workload $W --profile > $OUT/${TAG}_plain2_$W.log 2>&1 &&
  ncu --set full --clock-control none $SRC -k "$OURS" -s $SKIP -c $CNT -f -o $OUT/${TAG}_full_$W \
      python bench.py --workload $W --profile > $OUT/${TAG}_ncu2_$W.log 2>&1
  echo "full $W rc=$?"
done
# launch list of the default bench command (shares of the step)
python bench.py --steps 3 --warmup 3 --no-cpu --no-ref --no-extras > $OUT/${TAG}_plain_default.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/${TAG}_launches_arxiv-gat.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu --no-ref --no-extras > $OUT/${TAG}_ncu3.log 2>&1
echo "launch list rc=$?"
du -sh $OUT
