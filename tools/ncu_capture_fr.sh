#!/bin/bash
# ncu --set full of the Fr-table kernels in their final round-2 form (run on a GPU box through gpurun): the first launch of
# each kernel on a 2^24-entry table (tools/fr_kernels.py 24); text summaries into gpurun_out/
set -u
OUT=gpurun_out
mkdir -p $OUT /tmp/ncu
full() {   # name regex skip
  local name=$1 regex=$2 skip=$3
  timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$regex" -s $skip -c 1 -f -o /tmp/ncu/$name python tools/fr_kernels.py 24 > /dev/null 2>&1
  python tools/ncu_summary.py /tmp/ncu/$name.ncu-rep "ncu --set full --clock-control none --import-source on -k regex:$regex -s $skip -c 1 python tools/fr_kernels.py 24" > $OUT/r2_final_ncu_$name.txt 2>&1
  rm -f /tmp/ncu/$name.ncu-rep
}
full sumcheck_round_lazy '^k_sumcheck_round$' 0
full sumcheck_round_direct 'k_sumcheck_round_direct' 0
full sum_rounds3 'k_sum_rounds' 0
full div_phase2 'k_div_phase2' 0
ls -la $OUT | tail -6
