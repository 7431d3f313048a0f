"""Dev tool: time the device MSM at one size for several window widths (CUDA events)."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
import scz_b200 as scz
from scz_b200.api import msm_batched

logn = int(sys.argv[1]) if len(sys.argv) > 1 else 20
cs = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0, 12, 13, 14, 15, 16]
n = 1 << logn
ctx = scz.Context(0, n_parties=8)
g = torch.Generator(device="cuda").manual_seed(1)
k = torch.randint(0, 2**62, (n, 4), dtype=torch.int64, device="cuda", generator=g)
k[:, 3] &= (1 << 60) - 1
bases = ctx.g1_generator_mul(k)
s = torch.randint(-2**63, 2**63 - 1, (n, 4), dtype=torch.int64, device="cuda", generator=g)
s[:, 3] &= (1 << 62) - 1          # < 2^254 < r: valid Montgomery representatives
torch.cuda.synchronize()
for c in cs:
    ctx.msm_set_window(c)
    for _ in range(2):
        msm_batched(ctx, [bases], [s])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    e0.record()
    for _ in range(reps):
        msm_batched(ctx, [bases], [s])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    st = ctx.msm_last_stats()
    print(f"n=2^{logn} c={c}: {ms:.3f} ms  windows={st['windows']} buckets={st['buckets']} "
          f"adds={st['bucket_adds']}  {st['bucket_adds']/ms/1e6:.3f} G adds/s  {n/ms/1e3:.2f} M pairs/s", flush=True)
