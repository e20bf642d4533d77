T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
nvidia-smi -L | wc -l
for n in 2 4 8; do timeout 200 $T --nproc-per-node $n --master-port 2951$n bench.py --gpus $n --scaling strong --scenes 65536 --steps 50 --only --no-cpu-baseline > gpurun_out/bench_r2_nba_strong_${n}gpu.json 2> gpurun_out/strong_$n.err; echo strong $n rc=$?; done
timeout 200 $T --nproc-per-node 8 --master-port 29521 bench.py --gpus 8 --scaling strong --scenes 65536 --steps 50 --only --no-cpu-baseline --precision bf16 > gpurun_out/bench_r2_nba_strong_bf16_8gpu.json 2> gpurun_out/strong_bf16_8.err; echo strongbf16 rc=$?
timeout 300 $T --nproc-per-node 8 --master-port 29522 bench.py --gpus 8 --workload crowd --steps 5 > gpurun_out/bench_r2_crowd_8gpu.json 2> gpurun_out/crowd_8.err; echo crowd rc=$?
timeout 200 $T --nproc-per-node 8 --master-port 29523 bench.py --gpus 8 --mode train --steps 10 > gpurun_out/bench_r2_train_8gpu.json 2> gpurun_out/train_8.err; echo train rc=$?
timeout 200 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -3
for f in gpurun_out/bench_r2_nba_strong_*gpu.json gpurun_out/bench_r2_crowd_8gpu.json gpurun_out/bench_r2_train_8gpu.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(sys.argv[1], d["n_gpus"], round(d["value"]), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), d["e2e"].get("copy_ceiling",{}).get("value"))
except Exception as e: print(sys.argv[1], "ERR", e)
PY
done
