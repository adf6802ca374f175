"""Latency of ONE small batch on the device, per recover kernel (development tool).
Usage: python tools/latency_bench.py [n_items ...]   -- first n items of the config-3 fixture (10,000 committed seals + extras)."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ibft_b200 as ib  # noqa: E402

sizes = [int(x) for x in sys.argv[1:]] or [128, 1024, 4096, 10000, 14208, 20000]
d = np.load(os.path.join(ROOT, "tests", "golden", "config3.npz"))
base = np.ascontiguousarray(d["items"]).view(ib.ITEM_DTYPE).reshape(-1)
nmax = max(sizes)
items = np.ascontiguousarray(np.tile(base, (nmax + len(base) - 1) // len(base))[:nmax])
gold = np.unpackbits(d["bitmap"].view(np.uint8), bitorder="little")[: len(base)]
gold = np.tile(gold, (nmax + len(base) - 1) // len(base))[:nmax]
KEY_CACHE = os.environ.get("KEY_CACHE") == "1"   # known-key paths: the keys are learned by a first pass over the whole fixture
eng = ib.Engine(device=0, max_items=max(nmax, len(base), 1024), max_payload_bytes=1 << 22, max_groups=8, max_table_slots=2, max_validators=16384,
                key_cache=KEY_CACHE)
eng.set_validators(0, int(d["meta"][2]), d["addrs"], d["powers"])
eng.bind_groups(eng.groups(len(d["groups"])))
t_items = torch.from_numpy(items.view(np.uint8).reshape(-1, 128)).cuda()
t_arena = torch.from_numpy(np.ascontiguousarray(d["arena"])).cuda()
st = torch.cuda.Stream()
torch.cuda.set_stream(st)
if KEY_CACHE:
    t_all = torch.from_numpy(base.view(np.uint8).reshape(-1, 128)).cuda()
    t_bm0 = torch.zeros((len(base) + 31) // 32, dtype=torch.int32, device="cuda")
    eng.verify_device(t_all.data_ptr(), len(base), t_arena.data_ptr(), t_arena.numel(), 0, len(base), t_bm0.data_ptr(), 0, st.cuda_stream)
    torch.cuda.synchronize()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record(st)
    known = eng.refresh_key_tables()
    t1.record(st)
    torch.cuda.synchronize()
    print(json.dumps({"keys_known": known, "table_build_ms_incl_sync": round(t0.elapsed_time(t1), 2)}))
for n in sizes:
    words = (n + 31) // 32
    row = {"items": n}
    for name, path in (("auto", ib.Engine.PATH_AUTO), ("thread", ib.Engine.PATH_THREAD), ("quad", ib.Engine.PATH_QUAD), ("split", ib.Engine.PATH_SPLIT),
                       ("qsplit", ib.Engine.PATH_QSPLIT)):
        eng.set_recover_path(path)
        t_bm = torch.zeros(words, dtype=torch.int32, device="cuda")
        for _ in range(3):
            eng.verify_device(t_items.data_ptr(), n, t_arena.data_ptr(), t_arena.numel(), 0, n, t_bm.data_ptr(), 0, st.cuda_stream)
        torch.cuda.synchronize()
        got = np.unpackbits(t_bm.cpu().numpy().view(np.uint8), bitorder="little")[:n]
        ok = bool(np.array_equal(got, gold[:n]))
        reps = 20
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(st)
        for _ in range(reps):
            eng.verify_device(t_items.data_ptr(), n, t_arena.data_ptr(), t_arena.numel(), 0, n, t_bm.data_ptr(), 0, st.cuda_stream)
        b.record(st)
        torch.cuda.synchronize()
        row[name + "_us"] = round(a.elapsed_time(b) / reps * 1e3, 1)
        row[name + "_ok"] = ok
    print(json.dumps(row))
