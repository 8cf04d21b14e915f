"""Where does inference time go?  Eval-mode encoder launches of several group shapes, per-family kernel times (GPU box)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import intrepppid_b200 as ib
from intrepppid_b200 import _lib

torch.manual_seed(0)
net = ib.intrepppid_network(1).cuda().eval()
net.encoder.check_lengths = False
T = 1500
g = torch.Generator().manual_seed(1)


def run(tag, tok):
    with torch.no_grad():
        net.encoder.forward_groups(tok, draw=False)
        torch.cuda.synchronize()
        _lib.timing_enable(True)
        net.encoder.forward_groups(tok, draw=False)
        fam = {k: round(v[0], 3) for k, v in _lib.timing_read().items() if v[0] > 0.01}
        _lib.timing_enable(False)
    print(f"{tag:55s} lens={net.encoder.last_lengths[1][:4].tolist()} {fam}", flush=True)


full = torch.randint(1, 250, (1024, T), generator=g).cuda()
run("G=1 B=512 full length", full[:512].view(1, 512, T))
run("G=1 B=1024 full length", full.view(1, 1024, T))
run("G=128 B=8 full length", full.view(128, 8, T))
run("G=64 B=8 full length", full[:512].view(64, 8, T))
run("G=18 B=8 full length (144 full CTAs -> HALF)", full[:144].view(18, 8, T))
run("G=256 B=4 full length", full.view(256, 4, T))
rag = full.clone()
glen = torch.linspace(1500, 900, 128).long()
for i in range(128):
    rag[i * 8:(i + 1) * 8, glen[i]:] = 0
run("G=128 B=8 group lengths 1500..900", rag.view(128, 8, T))
run("G=1 B=512 same rows as one mixed batch (pads stepped)", rag[:512].view(1, 512, T))
half = full.clone(); half[:, 750:] = 0
run("G=128 B=8 all length 750", half.view(128, 8, T))
run("G=1 B=512 all length 750", half[:512].view(1, 512, T))
