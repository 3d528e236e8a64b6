import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import numpy as np, torch
from monsoon_b200.engine import Engine
eng = Engine(0)
seeds = torch.arange(8, dtype=torch.int64, device=eng.device)
arch = np.arange(1, 25, dtype=np.uint8)
for mode, keep, q in ((0, 12, 0.0), (1, 3, 0.0), (2, 0, 0.5)):
    d, f = eng.generate_decks(seeds, 3, mode, keep, q, arch, [2, 4])
    torch.cuda.synchronize()
    print(mode, d[0].cpu().numpy().tolist(), f[0].cpu().numpy().tolist())
