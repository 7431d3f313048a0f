set -u
OUT=gpurun_out
mkdir -p $OUT /tmp/ncu
APP="python tools/hp_time.py 20 1"
export SCZ_MSM_STREAM=1
full() {
  local name=$1 regex=$2 skip=$3; shift 3
  timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$regex" -s $skip -c 1 -f -o /tmp/ncu/$name "$@" > /tmp/ncu/$name.log 2>&1
  python tools/ncu_summary.py /tmp/ncu/$name.ncu-rep "ncu --set full --clock-control none --import-source on -k regex:$regex -s $skip -c 1 $*" > $OUT/r2_ncu_$name.txt 2>&1
  rm -f /tmp/ncu/$name.ncu-rep
}
full msm_recode_count '^k_msm_recode$' 0 $APP
full msm_recode_scatter '^k_msm_recode$' 1 $APP
full msm_tree_first '^k_msm_tree$' 0 $APP
full msm_tree_upper '^k_msm_tree$' 1 $APP
full msm_fixup '^k_msm_fixup$' 0 $APP
full open_fold '^k_open_fold$' 0 $APP
ls -la $OUT/r2_ncu_msm_* $OUT/r2_ncu_open_fold.txt
