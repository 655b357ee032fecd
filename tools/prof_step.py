"""A few eager training steps at C2 size for ncu launch lists / per-kernel captures."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from srfrd_b200 import synth
from srfrd_b200.trainer import FusedTrainer, discriminator_weights
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda", 0)
m, data = bench.make_model_and_data(dev)
tr = FusedTrainer(m, use_graph=False)
nb = synth.BatchSampler(data, 50, 100).next_batch(B)
b = {k: torch.from_numpy(v).to(dev) for k, v in nb.items()}
w = discriminator_weights(b["pos"], b["p_fake"], "soft")
for _ in range(steps):
    loss = tr.step(b, w_pos=w)
torch.cuda.synchronize()
print("loss", float(loss))
