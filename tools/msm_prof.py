import sys, torch
sys.path.insert(0, ".")
import scz_b200 as scz
from scz_b200.api import msm_batched
ctx = scz.Context(0, n_parties=8)
g = torch.Generator(device="cuda").manual_seed(1)
for logn in (10, 13, 16, 18):
    n = 1 << logn
    k = torch.randint(0, 2**62, (n, 4), dtype=torch.int64, device="cuda", generator=g)
    bases = ctx.g1_generator_mul(k)
    s = torch.randint(-2**63, 2**63 - 1, (n, 4), dtype=torch.int64, device="cuda", generator=g); s[:, 3] &= (1 << 62) - 1
    for _ in range(3): msm_batched(ctx, [bases], [s])
    ctx.prof_enable(True)
    reps = 5
    for _ in range(reps): msm_batched(ctx, [bases], [s])
    print(logn, {k2: round(ctx.prof_read(k2)[0] / reps, 3) for k2 in ("msm_sort", "msm_accumulate", "msm_fixup", "msm_reduce", "msm_finish")}, ctx.msm_last_stats())
    ctx.prof_enable(False)
