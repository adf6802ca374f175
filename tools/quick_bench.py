"""Kernel-variant timing helper (development tool): times k_recover on a device-resident replicated config-3 batch and checks
the verdict bitmap against the golden fixture.  Usage: IBFT_LIB=path/to/variant.so python tools/quick_bench.py [log2_items]"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ibft_b200 as ib  # noqa: E402

n = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 19)
d = np.load(os.path.join(ROOT, "tests", "golden", "config3.npz"))
base = np.ascontiguousarray(d["items"]).view(ib.ITEM_DTYPE).reshape(-1)
items = np.ascontiguousarray(np.tile(base, (n + len(base) - 1) // len(base))[:n])
KEY_CACHE = os.environ.get("KEY_CACHE") == "1"  # verify against learned validator keys instead of recovering
eng = ib.Engine(device=0, max_items=n, max_payload_bytes=1 << 22, max_groups=8, max_table_slots=2, max_validators=16384, key_cache=KEY_CACHE)
eng.set_validators(0, int(d["meta"][2]), d["addrs"], d["powers"])
groups = eng.groups(len(d["groups"]))
eng.bind_groups(groups)
t_items = torch.from_numpy(items.view(np.uint8).reshape(-1, 128)).cuda()
t_arena = torch.from_numpy(np.ascontiguousarray(d["arena"])).cuda()
t_bm = torch.zeros(n // 32, dtype=torch.int32, device="cuda")
st = torch.cuda.Stream()
torch.cuda.set_stream(st)
for _ in range(2):
    eng.verify_device(t_items.data_ptr(), n, t_arena.data_ptr(), t_arena.numel(), 0, n, t_bm.data_ptr(), 0, st.cuda_stream)
torch.cuda.synchronize()
n_keys = eng.refresh_key_tables()  # (0 without KEY_CACHE) the warm-up launches learned the keys; build their tables
for _ in range(1 if KEY_CACHE else 0):
    eng.verify_device(t_items.data_ptr(), n, t_arena.data_ptr(), t_arena.numel(), 0, n, t_bm.data_ptr(), 0, st.cuda_stream)
torch.cuda.synchronize()
got = np.unpackbits(t_bm.cpu().numpy().view(np.uint8), bitorder="little")[:n]
gold = np.unpackbits(d["bitmap"].view(np.uint8), bitorder="little")[: len(base)]
ok = bool(np.array_equal(got, np.tile(gold, (n + len(base) - 1) // len(base))[:n]))
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 5
a.record(st)
for _ in range(reps):
    eng.verify_device(t_items.data_ptr(), n, t_arena.data_ptr(), t_arena.numel(), 0, n, t_bm.data_ptr(), 0, st.cuda_stream)
b.record(st)
torch.cuda.synchronize()
ms = a.elapsed_time(b) / reps
info = eng.device_info()
print(json.dumps({"lib": os.environ.get("IBFT_LIB", "default"), "items": n, "ms": ms, "verifies_per_s": n / ms * 1e3, "bitmap_ok": ok, "key_cache": KEY_CACHE, "keys_known": n_keys,
                  "regs": info["kernel_regs"], "smem": info["kernel_smem_bytes"]}))
