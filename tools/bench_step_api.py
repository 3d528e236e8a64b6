"""GPU box: the step-per-launch API (sb_legal_mask fused into sb_step) at several batch sizes: env-steps/s and
the HBM fraction by the B_step = 1,047 B/step formula (this kernel DOES stream the state through HBM)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import torch
from monsoon_b200.engine import Engine
eng = Engine(0); dev = eng.device
for engine_id, n in [(e, n) for n in (256, 4096, 65536, 262144, 1048576) for e in (0, 1)]:
    eng.set_option("engine", engine_id)
    seeds = torch.arange(n, dtype=torch.int64, device=dev)
    st = eng.reset(seeds)
    eng.rollout_random(st, max_steps=20)            # mid-game states
    masks = eng.legal_mask(st)
    nmask = torch.empty_like(masks)
    # uniform agent on the device: lowest legal action (cheap, deterministic)
    def pick(m):
        w = m.to(torch.int64) & 0xFFFFFFFF
        out = torch.full((n,), 155, dtype=torch.int64, device=dev)
        for k in range(4, -1, -1):
            v = w[:, k]
            low = (v & -v)
            idx = torch.log2(low.clamp(min=1).to(torch.float64)).to(torch.int64) + 32 * k
            out = torch.where(v != 0, idx, out)
        return out.to(torch.uint8)
    acts = pick(masks)
    r = torch.empty(n, dtype=torch.int8, device=dev); d = torch.empty(n, dtype=torch.uint8, device=dev); e = torch.empty(n, dtype=torch.uint8, device=dev)
    times = []
    for it in range(12):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); eng.step(st, acts, r, d, e, next_masks=nmask); e1.record()
        torch.cuda.synchronize()
        if it >= 2: times.append(e0.elapsed_time(e1))
        acts = pick(nmask)
    ms = sum(times) / len(times)
    print(("thread" if engine_id == 0 else "warp  ") + " k_step n=%8d  %8.3f ms/launch  %8.1f M env-steps/s  HBM-equivalent %7.1f GB/s (%.2f %% of 6550)" % (n, ms, n / ms / 1e3, 1047 * n / ms / 1e6, 1047 * n / ms / 1e6 / 65.5), flush=True)
