"""dev tool: does the device's stream-ordered memory pool still GROW after the first proofs of a process?  Prints, for every
proof of a fresh process (one B200, 2^20, leader mode), the host time to enqueue it, its device time and the pool's reserved /
used bytes right after the enqueue (host-side allocator state: a growth shows up at the proof whose enqueue caused it).
`SCZ_POOL_PROBE_OPPORTUNISTIC=0` switches the pool's timing-dependent reuse policy off first."""
import os
import sys
import time

os.environ.setdefault("SCZ_MSM_STREAM", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from cuda.bindings import driver as cu  # noqa: E402

import scz_b200 as scz  # noqa: E402

n = int(os.environ.get("SCZ_POOL_PROBE_LOGN", "20"))
torch.cuda.set_stream(torch.cuda.Stream(priority=-1))
ctx = scz.Context(device=0, n_parties=8)
_, dev = cu.cuDeviceGet(0)
_, pool = cu.cuDeviceGetDefaultMemPool(dev)
A = cu.CUmemPool_attribute
if os.environ.get("SCZ_POOL_PROBE_OPPORTUNISTIC") == "0":
    print("opportunistic reuse off:", cu.cuMemPoolSetAttribute(pool, A.CU_MEMPOOL_ATTR_REUSE_ALLOW_OPPORTUNISTIC, cu.cuuint64_t(0)))


def stat():
    r = cu.cuMemPoolGetAttribute(pool, A.CU_MEMPOOL_ATTR_RESERVED_MEM_CURRENT)[1]
    u = cu.cuMemPoolGetAttribute(pool, A.CU_MEMPOOL_ATTR_USED_MEM_HIGH)[1]
    return int(r) >> 20, int(u) >> 20


pp = scz.PackedSharingParams(ctx, 1)
pk = scz.PackedProvingParameters.new(ctx, n, 1, seed=1, shared_seed=0, precompute=True)
torch.cuda.synchronize()
print("after set-up: pool reserved / used-high MiB", stat(), flush=True)
for rnd in range(int(sys.argv[1]) if len(sys.argv) > 1 else 4):
    k = 6
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(k + 1)]
    rows = []
    torch.cuda.synchronize()
    ev[0].record()
    for i in range(k):
        t0 = time.perf_counter()
        scz.dhyperplonk(ctx, n, pk, pp)
        h = (time.perf_counter() - t0) * 1e3
        ev[i + 1].record()
        rows.append((h,) + stat())
    torch.cuda.synchronize()
    for i, (h, r, u) in enumerate(rows):
        print(f"round {rnd} proof {i}: device {ev[i].elapsed_time(ev[i + 1]):7.1f} ms, host enqueue {h:6.1f} ms, pool reserved {r} MiB, used-high {u} MiB", flush=True)
