"""SURVEY.md §8(f) rank 4 probe: latency of ONE PastEncoder forward (front-end + fused corr/top-k + the three
MS-HGNN layers) at rollout batch sizes, launched eagerly and replayed from a captured CUDA graph.

The reference's simulator calls the model thousands of times at batch 1 (Simulator.py:231-238), where the
step is launch-bound; the library allocates nothing and launches only on the caller's stream, so one forward
is capturable as it stands.  With the seed passed by value a replay repeats the captured noise (first table);
`GraphedPastEncoder` keeps the noise fresh either from the CPU generator or from a device-resident Philox seed.

    python profiles/rollout_latency_probe.py
"""
import pathlib
import sys
import time
import types

import torch

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import groupnet_b200 as gb   # noqa: E402

DEV = torch.device("cuda:0")
N, T = 11, 5


def timed(fn, iters):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / iters, (time.perf_counter() - t0) * 1e6 / iters


def main():
    args = types.SimpleNamespace(hidden_dim=64, hyper_scales=[5, 11], past_length=T)
    torch.manual_seed(1)
    enc = gb.PastEncoder(args).to(DEV).eval()
    block = enc._interaction_block()
    for precision in ("fp32", "bf16"):
        block.set_precision(precision)
        for b in (1, 8, 64, 1024):
            x = torch.randn(b * N, T, 4, device=DEV)

            def fwd():
                block.set_rng("philox", seed=7)            # same noise every call, so results are comparable
                return enc(x, b, N)

            with torch.no_grad():
                for _ in range(5):
                    ref_f, ref_h = fwd()
                ref_f, ref_h = ref_f.clone(), ref_h.clone()
                eager_dev, eager_wall = timed(fwd, 200)
                launches = block.launches_per_forward(b, N) + 1
                line = f"{precision} B={b:5d} launches={launches:3d} eager: {eager_dev:8.1f} us device-span, {eager_wall:8.1f} us wall"
                try:
                    side = torch.cuda.Stream(DEV)
                    side.wait_stream(torch.cuda.current_stream(DEV))
                    with torch.cuda.stream(side):
                        fwd()
                    torch.cuda.current_stream(DEV).wait_stream(side)
                    graph = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(graph):
                        out_f, out_h = fwd()
                    graph.replay()
                    torch.cuda.synchronize()
                    same = torch.equal(out_f, ref_f) and torch.equal(out_h, ref_h)
                    g_dev, g_wall = timed(graph.replay, 200)
                    line += f" | graph replay: {g_dev:8.1f} us device-span, {g_wall:8.1f} us wall, bit-identical to eager: {same}"
                except Exception as exc:      # report, do not hide
                    line += f" | graph capture failed: {type(exc).__name__}: {str(exc)[:200]}"
            print(line, flush=True)
    # the packaged helper, with the reference's RNG contract (fresh torch.rand per call), one dependent call
    # after another as a rollout issues them: wall time per call including the final synchronize
    for precision in ("fp32", "bf16"):
        block.set_precision(precision)
        block.set_rng("cpu-compat")
        for b in (1, 8):
            x = torch.randn(b * N, T, 4, device=DEV)
            g = gb.GraphedPastEncoder(enc, b, N, T)
            with torch.no_grad():
                def eager_step():
                    f, _ = enc(x, b, N)
                    torch.cuda.synchronize()

                def graph_step():
                    f, _ = g(x)
                    torch.cuda.synchronize()
                for fn in (eager_step, graph_step):
                    for _ in range(20):
                        fn()
                t0 = time.perf_counter()
                for _ in range(300):
                    eager_step()
                t1 = time.perf_counter()
                for _ in range(300):
                    graph_step()
                t2 = time.perf_counter()
            print(f"{precision} B={b} cpu-compat noise, synchronised per call: eager {(t1 - t0) / 300 * 1e6:7.1f} us, "
                  f"GraphedPastEncoder {(t2 - t1) / 300 * 1e6:7.1f} us", flush=True)
            # Philox on the device: eager (seed by value) vs graph replay (seed in device memory, advanced in-graph)
            block.set_rng("philox", seed=3)
            gp = gb.GraphedPastEncoder(enc, b, N, T, rng="philox", seed=3)
            with torch.no_grad():
                def graph_philox_step():
                    f, _ = gp(x)
                    torch.cuda.synchronize()
                block.set_rng("philox", seed=3)
                for _ in range(20):
                    eager_step()
                t0 = time.perf_counter()
                for _ in range(300):
                    eager_step()
                t1 = time.perf_counter()
                gp.recapture()
                for _ in range(20):
                    graph_philox_step()
                t2 = time.perf_counter()
                for _ in range(300):
                    graph_philox_step()
                t3 = time.perf_counter()
            block.set_rng("cpu-compat")
            print(f"{precision} B={b} philox noise,     synchronised per call: eager {(t1 - t0) / 300 * 1e6:7.1f} us, "
                  f"GraphedPastEncoder {(t3 - t2) / 300 * 1e6:7.1f} us", flush=True)


if __name__ == "__main__":
    main()
