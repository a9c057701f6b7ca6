import sys, torch
sys.path.insert(0, ".")
from torchquad_b200 import ops
from oracle import ref_oracle as O
dev = torch.device("cuda")
dt, dim, ns = torch.float64, 8, 8
C = ns**dim
dh = torch.full((C,), 1.0 / C, dtype=dt, device=dev)
nh, offsets = ops.strat_nh(dh, 100_000_000)
M = int(offsets[-1])
for _ in range(2):
    y = ops.strat_sample(offsets, ns, dim, dt, 0, M, seed=1, call_idx=0)
xe, dxe, w, c = (x.to(dev) for x in O.map_init(4096, dim, dt))
for _ in range(2):
    ops.map_forward(y, xe, dxe)
torch.cuda.synchronize()
