"""One stratified fused pass (pairs histogram) against the map size: gather locality vs same-sector reduction contention."""
import sys

sys.path.insert(0, ".")
sys.argv = [sys.argv[0], "none"]
import torch

from torchquad_b200 import integrands as F

import importlib.util

spec = importlib.util.spec_from_file_location("exp_fused_pass", "scripts/exp_fused_pass.py")
m = importlib.util.module_from_spec(spec)
spec.loader.exec_module(m)
for ni in (256, 1024, 4096, 16384, 65536, 262144, 1048576):
    m.case(f"8D f64 osc Ns=8 Ni={ni} nh=5", F.GenzOscillatory(8, a=0.5, u=0.3), 8, torch.float64, 8, ni, 5, ("pairs", "nohist"))
for ni in (1024, 4096, 16384, 65536, 262144):
    m.case(f"16D f32 peak Ns=3 Ni={ni} nh=9", F.GenzProductPeak(16, a=2.0, u=0.5), 16, torch.float32, 3, ni, 9, ("pairs", "tile", "nohist"))
