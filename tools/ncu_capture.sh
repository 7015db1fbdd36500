#!/bin/bash
# On the GPU box: one `ncu --set full` capture of the hot kernels of the second step of a workload, summarised to text
# (the .ncu-rep is kept only when it is small enough to travel back in gpurun_out/).
#   tools/ncu_capture.sh <tag> <workload> [batch] [kernel regex] [skip] [count]
set -u
TAG=$1; WL=$2; BATCH=${3:-}; REGEX=${4:-'k_atom_bwd|k_atom_cat|k_mix_rows|k_policy|k_edge_pairs'}; SKIP=${5:-20}; COUNT=${6:-20}
OUT=gpurun_out/${TAG}
ncu --set full --clock-control none --import-source on -k "regex:${REGEX}" --launch-skip ${SKIP} --launch-count ${COUNT} \
    -o ${OUT} -f python tools/profile_step.py ${WL} 2 ${BATCH} > ${OUT}.log 2>&1
python tools/ncu_hot.py ${OUT}.ncu-rep > ${OUT}_metrics.txt 2>&1
ncu -i ${OUT}.ncu-rep --page raw --csv > ${OUT}_raw.csv 2>/dev/null
: > ${OUT}_lines.txt
for k in $(echo "${REGEX}" | tr '|' ' '); do
  python tools/ncu_lines.py ${OUT}.ncu-rep "$k" 14 >> ${OUT}_lines.txt 2>&1
  echo >> ${OUT}_lines.txt
done
SZ=$(stat -c %s ${OUT}.ncu-rep)
if [ "$SZ" -gt 30000000 ]; then rm -f ${OUT}.ncu-rep; fi
ls -la gpurun_out | tail -8
