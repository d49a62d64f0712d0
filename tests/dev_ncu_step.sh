#!/bin/bash
# Dev tool (GPU box): launch list + `ncu --set full` capture of the LAST fused mapping step of tests/dev_profile_step.py.
#   bash tests/dev_ncu_step.sh <tag> [cfg]
# writes gpurun_out/<tag>_launches.csv, gpurun_out/<tag>_step_raw.csv (ncu --page raw --csv of the step's kernels)
set -u
TAG=${1:-step}
CFG=${2:-c2}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv \
    python tests/dev_profile_step.py $CFG 3 > gpurun_out/${TAG}_launches.log 2>&1
read SKIP COUNT < <(python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/${TAG}_launches.csv")) if len(r) > 14 and r[0].isdigit()]
names=[r[4] for r in rows]
starts=[i for i,n in enumerate(names) if "clear_regions_kernel" in n]
last=starts[-1]
end=max(i for i,n in enumerate(names) if "adam_" in n)
print(last, end-last+1)
PY
)
echo "skip $SKIP count $COUNT" | tee gpurun_out/${TAG}_skip.txt
ncu --set full --clock-control none --launch-skip $SKIP --launch-count $COUNT -o gpurun_out/${TAG}_step -f \
    python tests/dev_profile_step.py $CFG 3 > gpurun_out/${TAG}_step_prof.log 2>&1
ncu -i gpurun_out/${TAG}_step.ncu-rep --page raw --csv > gpurun_out/${TAG}_step_raw.csv 2>/dev/null
rm -f gpurun_out/${TAG}_step.ncu-rep
ls -la gpurun_out/${TAG}_step_raw.csv
