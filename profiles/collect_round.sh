#!/bin/bash
# Evidence of one round, run on the GPU box: bench lines, ncu launch list of one timed step, `ncu --set full` captures of
# the kernels named in the rooflines. Writes under gpurun_out/<round>/ (copied to profiles/ by hand).
#   gpurun --timeout 2400 -- 'bash profiles/collect_round.sh r02'
R=${1:-r02}
O=gpurun_out/$R
mkdir -p $O
python bench.py > $O/bench_default.json 2> $O/bench_default.err
python bench.py --impl reference > $O/bench_reference.json 2> $O/bench_reference.err
python bench.py --batch 64 --no-sub > $O/bench_config5_batch64.json 2> $O/bench_config5.err
# launch list of exactly one timed step of the default workload
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/launches_config3.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-sub --profile-step > $O/launches_config3.log 2>&1
full() { # name, regex, skip, count, driver args
  ncu --set full --clock-control none --import-source on -k regex:"$2" --launch-skip $3 --launch-count $4 -f -o $O/$1 python profiles/prof_driver.py $5 > $O/$1.log 2>&1
  ncu -i $O/$1.ncu-rep --page raw --csv > $O/$1.csv 2>> $O/$1.log
  rm -f $O/$1.ncu-rep
}
full ncu_full_k_sst_forward_config3 '^k_sst_forward' 2 2 2
full ncu_full_k_sst_backward_config3 '^k_sst_backward' 2 2 2
full ncu_full_k_pre_config3 '^k_pre' 2 2 2
full ncu_full_k_post_config3 '^k_post' 2 2 2
full ncu_full_k_sst_factor_config3 '^k_sst_factor' 0 2 2
full ncu_full_k_flow_config2 'k_flow' 4 2 1   # one forward, one backward launch (split by profiles/split_ncu_csv.py)
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool python profiles/sanitize.py > $O/sanitizer_$tool.txt 2>&1
  tail -3 $O/sanitizer_$tool.txt
done
python - <<'PY' > $O/summary.txt 2>&1
import json, glob, sys
for f in sorted(glob.glob(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/r02/bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unreadable", e); continue
    print(f, {k: d.get(k) for k in ("impl", "metric", "value", "unit", "ms_per_step", "n_gpus", "gpu_launches")}, "e2e", (d.get("e2e") or {}).get("value"))
PY
ls -la $O
