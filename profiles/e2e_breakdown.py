"""Where the end-to-end step of MultiScaleInteraction.forward_host goes on one B200: chunked compute alone, the host-side
x-slice fill alone, and the whole pipeline per chunk size.  (Filling the x slice once AFTER every chunk is
enqueued instead of per chunk measured slower: 16.7 vs 15.4 ms at 8,192-scene chunks, profiles/e2e_breakdown_r2.txt.)  Run on a B200: python profiles/e2e_breakdown.py"""
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
m.set_precision("tf32")
x = torch.randn(b, n, d).pin_memory()
of = torch.empty(b, n, m.feature_width()).pin_memory()
oh = torch.empty(b, m.incidence_rows(n), n).pin_memory()
torch.set_grad_enabled(False)
print("threads", torch.get_num_threads(), "cpus", os.cpu_count(), flush=True)


def timed(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


print(f"host fill alone (184 MB into the strided x columns): {timed(lambda: of[:, :, :d].copy_(x)):.2f} ms", flush=True)
xd = x.cuda()
fd = torch.empty(b, n, m.feature_width(), device="cuda")
hd = torch.empty(b, m.incidence_rows(n), n, device="cuda")
print(f"compute, one chunk of 65536: {timed(lambda: m.forward(xd, out_feature=fd, out_H=hd, write_input_slice=False)):.2f} ms", flush=True)
for cs in (4096, 8192, 16384):
    def chunked():
        for c in range(b // cs):
            s = slice(c * cs, (c + 1) * cs)
            m.forward(xd[s], out_feature=fd[s], out_H=hd[s], write_input_slice=False)
    print(f"compute, chunks of {cs}: {timed(chunked):.2f} ms", flush=True)
for cs in (4096, 8192, 16384):
    ms = timed(lambda: m.forward_host(x, of, oh, chunk_scenes=cs, input_slice="host"))
    print(f"forward_host slice=host   chunk={cs:6d}            : {ms:7.2f} ms  {b / ms / 1e3:6.2f} M scenes/s", flush=True)
    ms = timed(lambda: m.forward_host(x, of, oh, chunk_scenes=cs, input_slice="device"))
    print(f"forward_host slice=device chunk={cs:6d}            : {ms:7.2f} ms  {b / ms / 1e3:6.2f} M scenes/s", flush=True)
