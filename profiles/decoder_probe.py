"""Timing probe of the fp32 decoder path (gn_decoder_fwd) at the NBA inference shape: 11 agents x 20 samples per
scene, 2 DecomposeBlocks, 0.75 GFLOP per scene.   python profiles/decoder_probe.py [scenes]"""
import pathlib
import sys
import types

import torch

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import groupnet_b200 as gb   # noqa: E402

DEV = torch.device("cuda:0")


def main():
    scenes = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    args = types.SimpleNamespace(hidden_dim=64, hyper_scales=[5, 11], zdim=32, past_length=5, future_length=10,
                                 num_decompose=2)
    torch.manual_seed(1)
    dec = gb.Decoder(args).to(DEV)
    a, s = scenes * 11, 20
    pf = torch.randn(a, 256, device=DEV).repeat_interleave(s, dim=0)
    z = torch.randn(a * s, 32, device=DEV)
    past = torch.randn(a, 5, 2, device=DEV)
    cur = torch.randn(a, 1, 2, device=DEV)
    with torch.no_grad():
        for _ in range(2):
            dec(pf, z, scenes, 11, past, cur, s, mode="inference")
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        iters = 5
        for _ in range(iters):
            dec(pf, z, scenes, 11, past, cur, s, mode="inference")
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    macs_per_row_block = 5 * 3 * 96 * (32 + 96) + 5 * 32 * 6 + 2 * (384 * 512 + 512 * 256) + 256 * 10 + 256 * 20
    flop = 2.0 * macs_per_row_block * a * s * 2
    print(f"decoder fp32 path: {scenes} scenes ({a * s} rows) {ms:.3f} ms -> {scenes / ms * 1e3:,.0f} scenes/s, "
          f"{flop / ms / 1e9:.1f} TFLOP/s (algorithmic, fp32 FFMA)")


if __name__ == "__main__":
    main()
