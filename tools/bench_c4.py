"""C4 throughput (long-sequence variant: maxlen 200, D=256, F=16, 4 blocks, batch 1024 per GPU), single GPU.
Prints training seqs/s from CUDA-graph replays on device-resident batches plus a per-entry-point breakdown."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from srfrd_b200 import SRFR_model as M, synth, _lib
from srfrd_b200.trainer import FusedTrainer, discriminator_weights

B, L = int(os.environ.get("C4_B", 1024)), 200
data = synth.make_interactions(1239, 22363, 12101, 20, 60.0, L)
torch.manual_seed(1239)
m = M.SRFR(data.itemnum, L, 256, 16, 0.0, 4, 1, "cuda")
for _, p in m.named_parameters():
    if p.dim() >= 2:
        torch.nn.init.xavier_normal_(p.data)
m = m.to("cuda")
tr = FusedTrainer(m, use_graph=True)
smp = synth.BatchSampler(data, L, 3)
batches = [{k: torch.from_numpy(v).cuda() for k, v in smp.next_batch(B).items()} for _ in range(4)]
valid = float(np.mean([float((b["pos"] != 0).float().mean()) for b in batches]))
step = lambda i: tr.step(batches[i % 4], w_pos=discriminator_weights(batches[i % 4]["pos"], batches[i % 4]["p_fake"], "soft"))
for i in range(4):
    step(i)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 10
e0.record()
for i in range(n):
    step(i)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
print(f"C4 B={B} L={L} valid={valid:.2f}: {ms:.3f} ms/step, {B / ms * 1e3:.0f} seqs/s, loss {float(tr.loss_dev):.4f}")
recs = []
tr.use_graph = False
_lib.set_profile(recs)
tr.run_step()
torch.cuda.synchronize()
_lib.set_profile(None)
agg = {}
for name, args, a, b in recs:
    d = agg.setdefault(name, [0.0, 0]); d[0] += a.elapsed_time(b); d[1] += 1
for k, (t, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"  {k:28s} {t:8.3f} ms x{c}")
