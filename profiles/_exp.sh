mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tf32.py -x -q 2>&1 | tail -2
timeout 300 python bench.py --only --no-cpu-baseline --steps 20 > gpurun_out/exp_vm.json 2> gpurun_out/exp_vm.err
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/exp_vm.json").read().splitlines() if l.startswith("{")][-1])
print(d["value"], d["ms_per_step"])
for k,v in d.get("kernels", {}).items(): print(k, v["ms_per_step"])
PY
