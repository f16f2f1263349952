#!/usr/bin/env python
"""L2 bandwidth of this GPU as a second roofline denominator for the traversal kernels (their working set, the 15 MB scene, is L2-resident):
a device-to-device copy of a buffer that fits L2 (16 MiB read + 16 MiB written of the 126 MB), repeated back to back; read+write bytes / time, best of 20,
CUDA events.  Writes profiles/<tag>_l2_peak.json."""
import json, os, sys, torch
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
n = 16 << 20
a = torch.empty(n, dtype=torch.uint8, device="cuda"); b = torch.empty_like(a)
for _ in range(10):
    b.copy_(a)
best = 0.0
for _ in range(20):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        b.copy_(a)
    e1.record(); torch.cuda.synchronize()
    best = max(best, 50 * 2 * n / (e0.elapsed_time(e1) * 1e-3) / 1e9)
big = torch.empty(1 << 30, dtype=torch.uint8, device="cuda"); big2 = torch.empty_like(big)
hbm = 0.0
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); big2.copy_(big); e1.record(); torch.cuda.synchronize()
    hbm = max(hbm, 2 * (1 << 30) / (e0.elapsed_time(e1) * 1e-3) / 1e9)
out = {"l2_copy_gbs": best, "hbm_copy_gbs_same_run": hbm, "how": "torch b.copy_(a), 16 MiB uint8 buffers (L2-resident), read+write bytes, 50 copies back to back, best of 20; HBM: 1 GiB buffers, best of 5",
       "gpu": torch.cuda.get_device_name(0)}
os.makedirs("profiles", exist_ok=True)
json.dump(out, open(f"profiles/{tag}_l2_peak.json", "w"), indent=1)
print(json.dumps(out))
