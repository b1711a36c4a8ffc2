#!/bin/bash
# Build box: turn the reports tools/profile_r02.sh brought back (gpurun_out/r02_*.ncu-rep) into the tracked summaries
# under profiles/ and regenerate profiles/ncu_kernel_stats.json (stamped with the hash of each kernel's sources).
set -e
cd "$(dirname "$0")/.."
tag=${1:-r02}
declare -A CELLS=( [k1]=274877906944 [k2s]=68719476736 [k2l_1024]=1073741824 [k2l_8192]=549755813888 )
for k in k1 k2s k2l_1024 k2l_8192; do
  rep=gpurun_out/${tag}_${k}.ncu-rep
  [ -f "$rep" ] || { echo "missing $rep"; continue; }
  ncu -i "$rep" --page raw --csv > /tmp/${tag}_${k}_raw.csv
  ncu -i "$rep" --page source --csv > /tmp/${tag}_${k}_src.csv
  { echo "# ncu --set full --clock-control none --import-source on, capture ${tag}_${k} (tools/profile_r02.sh; one launch, after 3 warm-up launches)";
    python tools/ncu_summary.py /tmp/${tag}_${k}_raw.csv /tmp/${tag}_${k}_src.csv ${CELLS[$k]}; } > profiles/${k}_${tag}_ncu_summary.txt
  python tools/ncu_buckets.py /tmp/${tag}_${k}_src.csv > profiles/${k}_${tag}_ncu_source_buckets.txt 2>/dev/null || true
done
python tools/ncu_stats.py $tag
cp gpurun_out/${tag}_launches.csv profiles/${tag}_launches.csv 2>/dev/null || true
