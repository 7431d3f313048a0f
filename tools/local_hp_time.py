"""dev tool: time local_hyperplonk (hyperplonk/src/hyperplonk.rs:15-160, the reference's monolithic baseline) on one GPU.
usage: python tools/local_hp_time.py [n] [reps] [pre|plain]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import scz_b200 as scz  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
pre = (sys.argv[3] if len(sys.argv) > 3 else "pre") == "pre"
ctx = scz.Context(device=0, n_parties=8)
gen = torch.Generator(device="cuda").manual_seed(3)


def rand_fr(m):
    t = torch.randint(-2**63, 2**63 - 1, (m, 4), dtype=torch.int64, device="cuda", generator=gen)
    t[:, 3] &= (1 << 62) - 1
    return t


gc = 1 << n
tabs = {k: rand_fr(gc) for k in ("input", "q1", "q2", "eq")}
tabs.update({k: rand_fr(4 * gc) for k in ("m", "ssigma", "sid", "eq_p2")})
tabs.update(challenge=rand_fr(n), challengep2=rand_fr(n + 2), alpha_beta=rand_fr(2))
tabs["a_evals"], tabs["b_evals"], tabs["c_evals"] = (tabs["m"][:gc].contiguous(), tabs["m"][gc:2 * gc].contiguous(),
                                                     tabs["m"][2 * gc:3 * gc].contiguous())     # fix_variable(m, (0,0) / (0,1) / (1,0))
t0 = time.time()
pc = scz.PolynomialCommitment(ctx, [ctx.g1_generator_mul(rand_fr(1 << i)) for i in range(n + 3)])   # new_toy shape
if pre:
    pc.precompute()
ctx.sync()
print(f"setup {time.time() - t0:.2f} s", flush=True)
for r in range(reps):
    ctx.prof_enable(r == reps - 1)
    torch.cuda.synchronize()
    t0 = time.time()
    proof = scz.local_hyperplonk(ctx, n, tabs, pc)
    torch.cuda.synchronize()
    dt = time.time() - t0
    print(f"rep {r}: {dt * 1e3:.1f} ms, {(1 << n) / dt:.0f} constraints/s", flush=True)
for k in ctx.KERNEL_CLASSES:
    ms, cnt = ctx.prof_read(k)
    print(f"  {k:16s} {ms:9.2f} ms  {cnt} brackets")
print(ctx.msm_cum_stats())
