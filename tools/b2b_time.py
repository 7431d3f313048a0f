"""dev tool: K proofs enqueued back to back (no synchronisation in between), device time per proof; the bench's main loop.
usage: python tools/b2b_time.py [n] [K]   (env: SCZ_MSM_AFFINE, SCZ_MSM_STREAM)"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import scz_b200 as scz  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
K = int(sys.argv[2]) if len(sys.argv) > 2 else 5
torch.cuda.set_stream(torch.cuda.Stream(priority=-1))
ctx = scz.Context(device=0, n_parties=8)
pp = scz.PackedSharingParams(ctx, 1)
pk = scz.PackedProvingParameters.new(ctx, n, 1, seed=1, precompute=True)
for _ in range(3):
    scz.dhyperplonk(ctx, n, pk, pp)
torch.cuda.synchronize()
for prof in (False, True):
    ctx.prof_enable(prof)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    t0 = time.time()
    for _ in range(K):
        scz.dhyperplonk(ctx, n, pk, pp)
    th = time.time() - t0
    b.record()
    torch.cuda.synchronize()
    print(f"prof={prof}: {a.elapsed_time(b) / K:.1f} ms per proof back to back (host enqueue {th / K * 1e3:.1f} ms per proof)", flush=True)
ctx.prof_enable(False)
for _ in range(3):
    t0 = time.time()
    scz.dhyperplonk(ctx, n, pk, pp)
    torch.cuda.synchronize()
    print(f"synchronised: {(time.time() - t0) * 1e3:.1f} ms")
