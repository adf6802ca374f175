"""2-rank probe of the peer-memory exchange (development): small round, prints what each step sees."""
import importlib
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ibft_b200 as ib  # noqa: E402

sharding = importlib.import_module("go-ibft_b200.sharding")
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
d = np.load(os.path.join(ROOT, "tests", "golden", "config2.npz"))
items = np.ascontiguousarray(d["items"]).view(ib.ITEM_DTYPE).reshape(-1)
eng = ib.Engine(device=lr, max_items=1 << 12, max_payload_bytes=1 << 22, max_groups=8, max_table_slots=2, max_validators=4096)
eng.set_validators(0, int(d["meta"][2]), d["addrs"], d["powers"])
groups = eng.groups(len(d["groups"]))
n = len(items)
lo, hi = sharding.shard_bounds(n, world, rank)
li, la = sharding.rebase_shard(items, d["arena"], lo, hi)
st = torch.cuda.Stream()
sv = sharding.ShardedVerifier(eng, n, groups, world, rank, li, la, st, exchange="p2p")
print(rank, "peer ptrs", [hex(p) for p in sv.peer_ptrs], flush=True)
for i in range(3):
    res, bm = sv.run()
    print(rank, "round", i, "bitmap ok", bool(np.array_equal(bm, d["bitmap"])), flush=True)
sv.close()
dist.barrier()
dist.destroy_process_group()
