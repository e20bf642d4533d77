"""Generate tests/golden/inference/*.npz by running the UNMODIFIED reference `GroupNet.inference_simulator`
(/root/reference/model/GroupNet_nba.py:830-869) on CPU, on the committed real NBA windows.

Run in the build container only:  python tests/golden/make_golden_inference.py

The model is the reference `GroupNet`; only its past_encoder / decoder WEIGHTS are chosen: they are loaded from
standalone reference `PastEncoder` / `Decoder` instances built right after `torch.manual_seed` (the drop-ins
reproduce that seeded initialisation, pinned by sha256), so a fixture needs no weights.  The noise is whatever the
reference draws from the global CPU generator after `torch.manual_seed(rng_seed)`: three `torch.rand` (one per
MLP_dict_softmax) and one `torch.randn` (Normal.fixed_eps) — the drop-in's rng="cpu-compat" contract.
"""
import hashlib
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("GROUPNET_REF", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REF)
for missing in ("tkinter", "glob2"):                    # model/GroupNet_nba.py:2, model/utils.py:9
    if missing not in sys.modules:
        stub = types.ModuleType(missing)
        stub.TRUE = True
        sys.modules[missing] = stub

from model.GroupNet_nba import Decoder, GroupNet, PastEncoder  # noqa: E402


def state_sha(module) -> str:
    hsh = hashlib.sha256()
    for k, v in module.state_dict().items():
        hsh.update(k.encode())
        hsh.update(v.detach().cpu().numpy().astype(np.float32).tobytes())
    return hsh.hexdigest()


def make_case(name, data, enc_seed, dec_seed, rng_seed):
    args = types.SimpleNamespace(hidden_dim=64, hyper_scales=[5, 11], past_length=5, future_length=10, zdim=32,
                                 num_decompose=2, learn_prior=False, ztype='gaussian', sample_k=20, min_clip=2.0)
    torch.manual_seed(0)
    model = GroupNet(args, torch.device("cpu")).eval()
    torch.manual_seed(enc_seed)
    enc = PastEncoder(args)
    torch.manual_seed(dec_seed)
    dec = Decoder(args)
    model.past_encoder.load_state_dict(enc.state_dict(), strict=True)
    model.decoder.load_state_dict(dec.state_dict(), strict=True)
    torch.manual_seed(rng_seed)
    with torch.no_grad():
        pred, h = model.inference_simulator(data)
    np.savez_compressed(os.path.join(HERE, "inference", name + ".npz"), data=data.numpy(), enc_seed=enc_seed,
                        dec_seed=dec_seed, rng_seed=rng_seed, enc_sha256=state_sha(model.past_encoder),
                        dec_sha256=state_sha(model.decoder), pred=pred.numpy(), H=h.numpy())
    print(f"{name}: pred {tuple(pred.shape)} H {tuple(h.shape)} max|pred| {pred.abs().max():.4f}")


def main():
    torch.set_num_threads(4)
    os.makedirs(os.path.join(HERE, "inference"), exist_ok=True)
    traj = torch.from_numpy(np.load(os.path.join(REF, "datasets/nba/test_nba.npy"))).float() / (94 / 28)
    past = traj[:, :5].permute(0, 2, 1, 3).contiguous()            # (B, N, 5, 2)
    make_case("nba_real_b1", past[:1].contiguous(), 81, 82, 83)    # the simulator's batch-1 call
    make_case("nba_real_b3", past[3:6].contiguous(), 84, 85, 86)


if __name__ == "__main__":
    main()
