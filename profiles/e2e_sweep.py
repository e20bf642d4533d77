"""End-to-end sweep of MultiScaleInteraction.forward_host on one B200: chunk size x input-slice mode x precision.
Run on a B200: python profiles/e2e_sweep.py"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import groupnet_b200 as gb

b, n, d = 65536, 11, 64
torch.manual_seed(1234)
m = gb.MultiScaleInteraction(d, (5, 11)).cuda().eval().set_rng("philox", seed=0)
for l in m.layers():
    l.workspace_limit_bytes = 9 << 30
x = torch.randn(b, n, d).pin_memory()
of = torch.empty(b, n, m.feature_width()).pin_memory()
oh = torch.empty(b, m.incidence_rows(n), n).pin_memory()
print("threads", torch.get_num_threads(), "cpus", os.cpu_count())
for prec in ("tf32", "bf16"):
    m.set_precision(prec)
    for mode in ("host", "device"):
        for cs in (2048, 4096, 8192, 16384, 32768):
            for _ in range(2):
                m.forward_host(x, of, oh, chunk_scenes=cs, input_slice=mode)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(5):
                m.forward_host(x, of, oh, chunk_scenes=cs, input_slice=mode)
            torch.cuda.synchronize()
            ms = (time.perf_counter() - t0) / 5 * 1e3
            print(f"{prec} slice={mode:6s} chunk={cs:6d}: {ms:7.2f} ms  {b / ms / 1e3:6.2f} M scenes/s", flush=True)
