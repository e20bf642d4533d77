"""wall time of ONE `GroupNet.inference_simulator` call (velocities, PastEncoder, prior sample,
Decoder with 20 samples per agent, permute) at rollout batch sizes — eager drop-in pipeline vs `GraphedInference`
(one CUDA-graph replay), with a synchronize after every call, as a rollout's dependent steps see it.

    python profiles/inference_latency_probe.py
"""
import pathlib
import sys
import time
import types

import torch

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import groupnet_b200 as gb   # noqa: E402

DEV = torch.device("cuda:0")
args = types.SimpleNamespace(hidden_dim=64, hyper_scales=[5, 11], past_length=5, future_length=10, zdim=32,
                             num_decompose=2)
torch.manual_seed(1)
enc = gb.PastEncoder(args).to(DEV).eval()
dec = gb.Decoder(args).to(DEV).eval()


def wall(fn, iters=200):
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(iters):
        fn()
        torch.cuda.synchronize()
    return (time.perf_counter() - t0) * 1e6 / iters


for precision, dec_precision in (("fp32", "fp32"), ("tf32", "fp32"), ("bf16", "fp32"), ("tf32", "bf16"), ("bf16", "bf16")):
    enc._interaction_block().set_precision(precision)
    dec.set_precision(dec_precision)
    for b in (1, 8):
        data = torch.randn(b, 11, 5, 2, device=DEV)
        enc._interaction_block().set_rng("cpu-compat")
        eager = wall(lambda: gb.inference_simulator(enc, dec, data))
        g_cpu = gb.GraphedInference(enc, dec, b, 11, rng="cpu-compat")
        t_cpu = wall(lambda: g_cpu(data))
        g_dev = gb.GraphedInference(enc, dec, b, 11, rng="philox", seed=3)
        t_dev = wall(lambda: g_dev(data))
        print(f"encoder {precision:5s} decoder {dec_precision:5s} B={b}: eager {eager:8.1f} us | graph, CPU-generator noise {t_cpu:8.1f} us | "
              f"graph, device noise {t_dev:8.1f} us", flush=True)
