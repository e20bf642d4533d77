# final captures after the fused decoder MLP kernel (run from the repo root on one B200)
python -m pytest tests -m gpu -q 2>&1 | tail -2
python bench.py --workload decoder --steps 10 > gpurun_out/bench_r2_final_decoder.json 2>/dev/null
python profiles/inference_latency_probe.py > gpurun_out/inference_latency.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:decoder -s 12 -c 6 --csv --log-file gpurun_out/launches_r2_decoder_tc.csv python profiles/decoder_tc_step.py > gpurun_out/ncu_dec0.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:decoder -s 12 -c 3 -o /tmp/dec_tc python profiles/decoder_tc_step.py > gpurun_out/ncu_dec.log 2>&1
python profiles/summarize_ncu.py /tmp/dec_tc.ncu-rep > gpurun_out/ncu_full_r2_decoder_tc.txt
python profiles/ncu_pipes.py /tmp/dec_tc.ncu-rep > gpurun_out/ncu_pipes_r2_decoder_tc.txt
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/bench_r2_final_decoder.json").read().splitlines() if l.startswith("{")][-1])
print(d.get("dtype"), round(d["value"]), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), {k:(round(v["value"]), round(v["ms_per_step"],3), v["tflops"], round(v["e2e"]["value"]), v["kernels"]) for k,v in d.get("paths",{}).items()}, (d.get("cpu_baseline") or {}).get("value"))
PY
tail -4 gpurun_out/inference_latency.txt
