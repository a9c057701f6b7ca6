import sys, torch
sys.path.insert(0, ".")
from torchquad_b200 import ops
dev = torch.device("cuda")
n, dim, dt = 33, 6, torch.float64
nodes = torch.linspace(0, 1, n, dtype=dt, device=dev).repeat(dim, 1).contiguous()
P = 400_000_000
f = ops.philox_uniform(P, 1, dt, dev, 1, 0).reshape(-1)
for _ in range(3):
    ops.nc_contract(f, nodes, 0, P)
torch.cuda.synchronize()
