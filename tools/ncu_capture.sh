#!/bin/bash
# Round-2 ncu evidence (run on a GPU box through gpurun; writes text summaries into gpurun_out/, the .ncu-rep files are dropped):
#   * launch list (gpu__time_duration) of one whole 2^20 proof
#   * ncu --set full of the kernels DESIGN.md quotes: one launch each
set -u
OUT=gpurun_out
mkdir -p $OUT /tmp/ncu
APP="python tools/hp_time.py 20 1"
export SCZ_MSM_STREAM=1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/r2_hp_ncu_launches.csv $APP > /dev/null 2>&1
full() {   # name regex skip header...
  local name=$1 regex=$2 skip=$3; shift 3
  timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$regex" -s $skip -c 1 -f -o /tmp/ncu/$name "$@" > /dev/null 2>&1
  python tools/ncu_summary.py /tmp/ncu/$name.ncu-rep "ncu --set full --clock-control none --import-source on -k regex:$regex -s $skip -c 1 $*" > $OUT/r2_ncu_$name.txt 2>&1
  rm -f /tmp/ncu/$name.ncu-rep
}
# setup of hp_time launches no k_ba / k_msm kernels before the first proof except SRS precompute; skip counts are per kernel name
full ba_phase2_level0 '^k_ba_phase2' 0 $APP
full ba_phase2_level1 '^k_ba_phase2' 1 $APP
full ba_phase1_level0 '^k_ba_phase1' 0 $APP
full ba_phase1_level1 '^k_ba_phase1' 1 $APP
full ba_accumulate '^k_ba_accumulate' 0 $APP
full ba_levels '^k_ba_levels' 0 $APP
full msm_recode_scatter 'k_msm_recode<true>|k_msm_recodeILb1' 0 $APP
full msm_recode_count 'k_msm_recode<false>|k_msm_recodeILb0' 0 $APP
full msm_tree_first 'k_msm_tree<true>|k_msm_treeILb1' 0 $APP
full msm_tree_upper 'k_msm_tree<false>|k_msm_treeILb0' 0 $APP
full msm_fixup '^k_msm_fixup\(' 0 $APP
full inv_tree_down 'k_inv_tree_down' 0 $APP
full open_fold '^k_open_fold\(' 0 $APP
full sumcheck_round '^k_sumcheck_round' 0 $APP
full div_phase2 'k_div_phase2' 0 $APP
full pss_dmsm_multi 'k_pss_dmsm_multi' 0 $APP
full sum_rounds3 'k_sum_rounds' 0 python tools/fr_kernels.py
SCZ_MSM_AFFINE=0 full msm_accumulate_xyzz '^k_msm_accumulate' 0 $APP
ls -la $OUT | tail -30
