"""ncu target: three steps of the bf16 tensor-core decoder (gn_decoder_fwd_tc) at the bench shape (1,024 scenes x 11 agents
x 20 samples, 2 DecomposeBlocks = 6 launches per step: gru_tc, mlp_fused, finish per block).  ncu -k regex:decoder -s 12 -c 6 captures the third step."""
import os
import sys
import types

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import groupnet_b200 as gb

torch.set_grad_enabled(False)
scenes, n, s = 1024, 11, 20
cfg = types.SimpleNamespace(hidden_dim=64, hyper_scales=[5, 11], zdim=32, past_length=5, future_length=10, num_decompose=2)
torch.manual_seed(1234)
dec = gb.Decoder(cfg).cuda().set_precision(sys.argv[1] if len(sys.argv) > 1 else "bf16")
a = scenes * n
pf = torch.randn(a, 256).repeat_interleave(s, dim=0).cuda()
z = torch.randn(a * s, 32).cuda()
past, cur = torch.randn(a, 5, 2).cuda(), torch.randn(a, 1, 2).cuda()
for _ in range(3):
    dec(pf, z, scenes, n, past, cur, s, mode="inference")
torch.cuda.synchronize()
print("ok")
