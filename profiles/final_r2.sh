# round-2 final captures on one B200 (run from the repo root): bench lines, ncu launch list + full capture of one tf32 step
python bench.py > gpurun_out/bench_r2_final_nba.json 2> gpurun_out/bench_r2_final_nba.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r2_final_reference_arm.json 2>/dev/null
python bench.py --workload fish8 --steps 10 > gpurun_out/bench_r2_final_fish8.json 2>/dev/null
python bench.py --workload fish20 --steps 10 > gpurun_out/bench_r2_final_fish20.json 2>/dev/null
python bench.py --workload decoder --steps 5 > gpurun_out/bench_r2_final_decoder.json 2>/dev/null
python bench.py --mode train --steps 10 > gpurun_out/bench_r2_final_train.json 2>/dev/null
python bench.py --workload fishops --steps 10 > gpurun_out/bench_r2_final_fishops.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"tf32|node2edge|edge2node|corr_topk" -c 120 --csv --log-file gpurun_out/launches_r2_final_tf32.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --only > gpurun_out/ncu1.log 2>&1
ncu --set full --clock-control none -k regex:"tf32|node2edge|edge2node|corr_topk" -s 51 -c 17 -o /tmp/prof_r2_final_tf32_step python bench.py --steps 2 --warmup 3 --no-cpu-baseline --only > gpurun_out/ncu2.log 2>&1
python profiles/summarize_ncu.py /tmp/prof_r2_final_tf32_step.ncu-rep > gpurun_out/ncu_full_r2_final_tf32_nba.txt   # the report itself exceeds what gpurun copies back
python profiles/ncu_pipes.py /tmp/prof_r2_final_tf32_step.ncu-rep > gpurun_out/ncu_pipes_r2_final_tf32_nba.txt          # issue slots, pipes, stall reasons per launch
python profiles/tf32_error_margin.py > gpurun_out/tf32_error_margin_r2.txt 2>&1
for f in gpurun_out/bench_r2_final_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]).read().splitlines() if l.startswith("{")][-1]); print(sys.argv[1].split('/')[-1], d.get("dtype"), round(d["value"]), round(d["ms_per_step"],3), "e2e", round(d.get("e2e",{}).get("value",0)))
except Exception as e: print(sys.argv[1], "ERR", e)
PY
done
