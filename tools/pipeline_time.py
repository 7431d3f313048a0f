"""dev tool: throughput of dhyperplonk (leader mode, one GPU) with several proofs in flight -- one ctx, one host thread
and one high-priority stream per prover; with SCZ_MSM_STREAM=1 every ctx runs its MSM launch sequences on a
lowest-priority side stream, so one prover's protocol kernels overlap another's bucket kernel.
usage: [SCZ_MSM_STREAM=1] python tools/pipeline_time.py [n] [in_flight] [proofs_per_prover]"""
import os
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import scz_b200 as scz  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
K = int(sys.argv[2]) if len(sys.argv) > 2 else 2
R = int(sys.argv[3]) if len(sys.argv) > 3 else 4
torch.cuda.set_device(0)
provers = []
for i in range(K):
    s = torch.cuda.Stream(priority=-1)
    with torch.cuda.stream(s):
        ctx = scz.Context(device=0, n_parties=8)
        pp = scz.PackedSharingParams(ctx, 1)
        pk = scz.PackedProvingParameters.new(ctx, n, 1, seed=1 + i, precompute=True)
        ctx.sync()
    provers.append((s, ctx, pp, pk))
print(f"{K} provers ready, {torch.cuda.memory_allocated() / 2**30:.1f} GiB, SCZ_MSM_STREAM={os.environ.get('SCZ_MSM_STREAM', '0')}", flush=True)


def run(reps):
    bar = threading.Barrier(K + 1)
    done = []

    def body(i):
        s, ctx, pp, pk = provers[i]
        torch.cuda.set_device(0)
        with torch.cuda.stream(s):
            bar.wait()
            for _ in range(reps):
                scz.dhyperplonk(ctx, n, pk, pp)
            ctx.sync()
        done.append(time.time())
    ts = [threading.Thread(target=body, args=(i,)) for i in range(K)]
    for t in ts:
        t.start()
    torch.cuda.synchronize()
    bar.wait()
    t0 = time.time()
    for t in ts:
        t.join()
    torch.cuda.synchronize()
    return time.time() - t0


run(2)   # warm-up
dt = run(R)
print(f"in flight {K}: {K * R} proofs in {dt * 1e3:.1f} ms -> {dt * 1e3 / (K * R):.1f} ms per proof, "
      f"{K * R * (1 << n) / dt:.0f} constraints/s", flush=True)
