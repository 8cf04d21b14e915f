"""Ablation timing of the forward recurrent kernel (IB200_DBG flags: 1 no MMA, 2 no activations, 4 no stores, 8 no prefetch, 16 no barrier).
Needs a library built with IB200_ABLATE=1 (python -m intrepppid_b200.build --force); production builds compile the switches out."""
import os, subprocess, sys, json
code = r'''
import sys, torch, os
sys.path.insert(0, ".")
import intrepppid_b200 as ib
from intrepppid_b200 import _lib
torch.manual_seed(0)
mode = sys.argv[1]
net = ib.intrepppid_network(1, precision=mode, embedding_droprate=0.0).cuda().train()
net.encoder.check_lengths = False
g = torch.Generator().manual_seed(1)
tok = torch.randint(1, 250, (5, 80, 1500), generator=g).cuda()
for _ in range(2): z = net.encoder.forward_groups(tok)
torch.cuda.synchronize(); _lib.timing_enable(True)
for _ in range(3): z = net.encoder.forward_groups(tok)
torch.cuda.synchronize(); t = _lib.timing_read()
print(" ".join(f"{k}={v[0]/v[1]:.3f}" for k, v in t.items() if k.startswith("lstm")))
'''
for mode in ("fp32", "bf16"):
    for flags in (0, 1, 2, 4, 8):
        env = dict(os.environ, IB200_DBG=str(flags))
        r = subprocess.run([sys.executable, "-c", code, mode], env=env, capture_output=True, text=True)
        print(f"{mode} dbg={flags:2d}: {r.stdout.strip()} {r.stderr.strip()[-200:] if r.returncode else ''}", flush=True)
