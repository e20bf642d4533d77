# final captures of round 2 after the tensor-core decoder (run from the repo root on one B200)
python -m pytest tests -m gpu -q 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-300
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2_final_nba.json 2> gpurun_out/nba.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r2_final_reference_arm.json 2>/dev/null
python bench.py --workload decoder --steps 10 > gpurun_out/bench_r2_final_decoder.json 2>/dev/null
for f in gpurun_out/bench_r2_final_nba.json gpurun_out/bench_r2_final_reference_arm.json gpurun_out/bench_r2_final_decoder.json; do python - "$f" <<'PY'
import json,sys
d=json.loads([l for l in open(sys.argv[1]).read().splitlines() if l.startswith("{")][-1]); print(sys.argv[1].split('/')[-1], d.get("dtype"), round(d["value"]), round(d["ms_per_step"],3), "e2e", round(d.get("e2e",{}).get("value",0)), {k:(round(v["value"]), round(v["ms_per_step"],3)) for k,v in d.get("paths",{}).items()}, (d.get("cpu_baseline") or {}).get("value"), d.get("clocks",{}).get("sm_mhz"), d.get("clocks",{}).get("reasons"))
PY
done
