import sys
sys.path.insert(0, '.')
import torch
import aigar_b200.layout as lay
from aigar_b200.env import AgarBatch
for kw, E in ((dict(num_nn=1, num_greedy=1, virus=True, split=True, eject=True), 4736), (dict(num_nn=8, num_greedy=8, virus=True, split=True, eject=True), 600)):
    cfg = lay.derive_config(**kw)
    whole = AgarBatch(cfg, E, seed=3, first_env_id=0)
    h = E // 3
    parts = [AgarBatch(cfg, n, seed=3, first_env_id=f, tile_width=w) for f, n, w in ((0, h, None), (h, h, 16), (2 * h, E - 2 * h, None))]
    for b in [whole] + parts:
        b.rollout_random(40, 8, 0)
    a = whole.state_tensor()
    bcat = torch.cat([p.state_tensor() for p in parts], 0)
    print(kw.get("num_nn"), "players-config: shards equal whole:", bool(torch.equal(a, bcat)), "tiles", whole.tile_width, [p.tile_width for p in parts])
