set -e
for cfg in "200 0" "125 0" "200 1" "125 1" "100 1"; do
  set -- $cfg
  echo "== CAP_PCT=$1 WIPE_FIRST=$2"
  DGS_RL_CAP_PCT=$1 DGS_WIPE_FIRST=$2 python tools/breakdown.py --reps 100 2>/dev/null | grep -E "raw_sample_gpu|api_sample"
  DGS_RL_CAP_PCT=$1 DGS_WIPE_FIRST=$2 DGS_BLOCKS_TRACE=1 python tools/timing_run.py 2>&1 | grep "coop trace" | tail -2
done
