"""Small catalogue-scoring run for ncu: U users x N items, D=64."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from srfrd_b200 import evaluation as EV
U, N, D = int(sys.argv[1]) if len(sys.argv) > 1 else 4096, int(sys.argv[2]) if len(sys.argv) > 2 else 262144, 64
g = torch.Generator().manual_seed(0)
table = (torch.randn(N + 1, D, generator=g) * 0.01).cuda()
feats = torch.randn(U, D, generator=g).cuda()
idx = EV.CatalogueIndex(table, 0)
for _ in range(3):
    s, i = EV.local_topk(feats, idx, 1)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    s, i = EV.local_topk(feats, idx, 1)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"topk U={U} N={N}: {ms:.3f} ms/pass, {2.0*U*N*D/ms/1e9:.1f} TFLOP/s, {U/ms*1e3:.0f} users/s")
