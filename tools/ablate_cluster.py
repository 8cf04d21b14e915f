"""Ablation timing of the cluster forward kernel at the config-5 shape (needs an IB200_ABLATE=1 build): IB200_DBG bits
1 no remote h stores, 2 CTA barrier instead of cluster barrier, 4 no global stores, 8 no MMAs."""
import os, subprocess, sys
code = r'''
import sys, torch
sys.path.insert(0, ".")
import intrepppid_b200 as ib
from intrepppid_b200 import _lib
torch.manual_seed(0)
mode = sys.argv[1]
net = ib.intrepppid_network(1, embedding_size=256, rnn_num_layers=3, bi_reduce="mean", precision=mode).cuda().eval()
net.encoder.check_lengths = False
x = torch.randint(1, 250, (256, 2000), generator=torch.Generator().manual_seed(777)).cuda()
with torch.no_grad():
    net.encoder(x); torch.cuda.synchronize(); _lib.timing_enable(True)
    net.encoder(x); torch.cuda.synchronize(); t = _lib.timing_read()
print(" ".join(f"{k}={v[0]/v[1]:.2f}" for k, v in t.items() if k.startswith("lstm")))
'''
for mode in ("bf16", "fp32"):
    for flags in (0, 1, 2, 3, 4, 8, 15):
        env = dict(os.environ, IB200_DBG=str(flags))
        r = subprocess.run([sys.executable, "-c", code, mode], env=env, capture_output=True, text=True)
        print(f"{mode} dbg={flags:2d}: {r.stdout.strip()} {r.stderr.strip()[-200:] if r.returncode else ''}", flush=True)
