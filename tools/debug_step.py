"""Debug helper (GPU box): replay one default-deck game through the oracle up to `step`, apply that
step on both sides and print the field-wise diff.  usage: python tools/debug_step.py SEED STEP"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import numpy as np, torch
import sb_oracle as o
from sb_layout import STATE_DTYPE
from monsoon_b200.engine import Engine, DEFAULT_DECKS, deck_indices

seed, target = int(sys.argv[1]), int(sys.argv[2])
D0, D1 = (deck_indices(d) for d in DEFAULT_DECKS)
st = o.new_game(seed, D0, D1, 3, 2)
eng = Engine(0)
for step in range(target + 1):
    m = o.legal_mask(st)
    legal = [a for a in range(156) if m[a >> 5] >> (a & 31) & 1]
    a = legal[o.lib().sbo_agent_pick(seed, step, len(legal))]
    if step == target:
        dev = torch.from_numpy(st.copy()[None]).to(eng.device)
        eng.step(dev, torch.tensor([a], dtype=torch.uint8, device=eng.device))
        got = dev.cpu().numpy()[0]
    o.step(st, a)
sa = np.frombuffer(st.tobytes(), dtype=STATE_DTYPE)[0]
sb = np.frombuffer(got.tobytes(), dtype=STATE_DTYPE)[0]
print("action", a)
for name in STATE_DTYPE.names:
    if name in ("pl", "tile"):
        for i in range(len(sa[name])):
            for sub in sa[name].dtype.names:
                if not np.array_equal(sa[name][i][sub], sb[name][i][sub]):
                    print("%s[%d].%s oracle=%s gpu=%s" % (name, i, sub, sa[name][i][sub], sb[name][i][sub]))
    elif not np.array_equal(sa[name], sb[name]):
        print("%s oracle=%s gpu=%s" % (name, sa[name], sb[name]))
