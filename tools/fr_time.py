"""dev tool: device time of the single-MLE / product sumcheck, open fold and division entry points at 2^22 and 2^24 entries"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import scz_b200 as scz  # noqa: E402

ctx = scz.Context(device=0, n_parties=8)
g = torch.Generator(device=ctx.device).manual_seed(7)
for logn in (22, 24):
    n = 1 << logn
    f = torch.randint(-2**63, 2**63 - 1, (n, 4), dtype=torch.int64, device=ctx.device, generator=g)
    f[:, 3] &= (1 << 62) - 1
    h = f.flip(0).contiguous()
    ch = f[:26].clone()
    for name, fn, nbytes in (("single-MLE sumcheck", lambda: scz.sumcheck(ctx, f, ch), (n - 1) * 96),
                             ("product sumcheck", lambda: scz.sumcheck_product(ctx, f, h, ch), (n - 1) * 192),
                             ("division", lambda: scz.fr_pointwise(ctx, "div", f, h), n * 96)):
        for _ in range(3):
            fn()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        import time
        t0 = time.perf_counter()
        a.record()
        for _ in range(5):
            fn()
        b.record()
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 5
        print(f"2^{logn} {name:22s} {ms:8.4f} ms  {nbytes / ms / 1e6:8.1f} GB/s algorithmic   (host enqueue {(t1 - t0) / 5 * 1e3:.4f} ms per call)",
              flush=True)
