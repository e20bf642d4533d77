timeout 300 python -m pytest tests/test_gpu_decoder.py -x -q 2>&1 | tail -5
timeout 300 python bench.py --workload decoder --steps 10 --no-cpu-baseline > gpurun_out/dec.json 2> gpurun_out/dec.err; tail -3 gpurun_out/dec.err; python - <<EOF
import json
d=json.loads([l for l in open("gpurun_out/dec.json") if l.startswith("{")][-1])
print(d["dtype"], d["value"], d["ms_per_step"], d["e2e"]["value"]); print(d["kernels"])
for k,v in d["paths"].items(): print(k, v["value"], v["ms_per_step"], v["tflops"], v["e2e"], v["kernels"])
EOF
