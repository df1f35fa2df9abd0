#!/bin/bash
# ncu evidence for one mode (infer|train): launch list of two eager passes, then one `--set full` capture of a
# median launch of each of the top kernels.  Run under gpurun from the repo root; outputs land in gpurun_out/.
#   bash profiles/ncu_capture.sh infer r01c 7
set -u
MODE=${1:-infer}; TAG=${2:-r01}; TOP=${3:-6}
OUT=gpurun_out
REP=${NCU_REP_DIR:-/tmp/ncu_rep}      # the .ncu-rep files stay on the box (gpurun_out/ is capped at 64 MiB): summaries travel
mkdir -p $OUT $REP
LIST=$OUT/launches_${TAG}_${MODE}.csv
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    --csv --log-file $LIST python profiles/ncu_target.py $MODE > $OUT/ncu_list_${MODE}.log 2>&1
python profiles/pick_launches.py $LIST $TOP > $OUT/picks_${TAG}_${MODE}.txt
cat $OUT/picks_${TAG}_${MODE}.txt
while read -r ID NAME; do
  timeout 600 ncu --set full --clock-control none --import-source on -s $ID -c 1 -f \
      -o $REP/ncu_${TAG}_${MODE}_${NAME} python profiles/ncu_target.py $MODE > $REP/ncu_full_${MODE}_${NAME}.log 2>&1
  python profiles/summarize_ncu.py $REP/ncu_${TAG}_${MODE}_${NAME}.ncu-rep > $OUT/ncu_${TAG}_${MODE}_${NAME}.txt 2>/dev/null
done < $OUT/picks_${TAG}_${MODE}.txt
ls -la $OUT/ncu_${TAG}_${MODE}_*.txt
