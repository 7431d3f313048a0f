"""dev tool: time dhyperplonk (leader mode, one GPU) at circuit size 2^n with the per-kernel-class device timers.
usage: python tools/hp_time.py [n] [reps] [pre|plain] [l]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import scz_b200 as scz  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
l = int(sys.argv[4]) if len(sys.argv) > 4 else 1
ctx = scz.Context(device=0, n_parties=8 * l)
pp = scz.PackedSharingParams(ctx, l)
t0 = time.time()
pre = (sys.argv[3] if len(sys.argv) > 3 else 'pre') == 'pre'
pk = scz.PackedProvingParameters.new(ctx, n, l, seed=1, precompute=pre)
ctx.sync()
print(f"setup {time.time() - t0:.2f} s, {torch.cuda.memory_allocated() / 2**30:.2f} GiB", flush=True)
for r in range(reps):
    ctx.prof_enable(r == reps - 1)
    l0 = ctx.launches
    torch.cuda.synchronize()
    t0 = time.time()
    proof = scz.dhyperplonk(ctx, n, pk, pp)
    t_host = time.time() - t0
    torch.cuda.synchronize()
    dt = time.time() - t0
    print(f"rep {r}: {dt * 1e3:.1f} ms (host enqueue {t_host * 1e3:.1f} ms), {(1 << n) / dt:.0f} constraints/s, "
          f"{ctx.launches - l0} launches", flush=True)
tot = 0
for k in ctx.KERNEL_CLASSES:
    ms, cnt = ctx.prof_read(k)
    tot += ms
    print(f"  {k:16s} {ms:9.2f} ms  {cnt} brackets")
print(f"  sum {tot:.2f} ms")
print("comm", ctx.get_comm())
