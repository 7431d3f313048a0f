"""dev tool: is the host throttled by the launch queue of ONE stream?  Host time to enqueue proof k + 1 while proof k runs,
on the same stream and on alternating streams (fenced by an event so that the proofs still run one after the other)."""
import os
import sys
import time

os.environ.setdefault("SCZ_MSM_STREAM", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import scz_b200 as scz  # noqa: E402

n = 20
s0 = torch.cuda.Stream(priority=-1)
s1 = torch.cuda.Stream(priority=-1)
torch.cuda.set_stream(s0)
ctx = scz.Context(device=0, n_parties=8)
pp = scz.PackedSharingParams(ctx, 1)
pk = scz.PackedProvingParameters.new(ctx, n, 1, seed=1, shared_seed=0, precompute=True)
for _ in range(3):
    scz.dhyperplonk(ctx, n, pk, pp)
torch.cuda.synchronize()
for mode in ("same", "alt", "same", "alt"):
    streams = [s0, s0] if mode == "same" else [s0, s1]
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record(s0)
    host, prev, keep = [], None, []
    for i in range(8):
        s = streams[i % 2]
        with torch.cuda.stream(s):
            if prev is not None:
                s.wait_event(prev)
            ctx.use_torch_stream()
            t0 = time.perf_counter()
            keep.append(scz.dhyperplonk(ctx, n, pk, pp))
            host.append((time.perf_counter() - t0) * 1e3)
            prev = torch.cuda.Event()
            prev.record(s)
    b.record(streams[7 % 2])
    torch.cuda.synchronize()
    print(f"{mode:5s}: {a.elapsed_time(b) / 8:7.2f} ms per proof; host enqueue per proof: " + " ".join(f"{h:.0f}" for h in host), flush=True)
    with torch.cuda.stream(s0):
        ctx.use_torch_stream()
