"""dev tool: one call of each Fr-table entry point on 2^22-entry tables (for ncu captures of kernels the prover does not use,
e.g. the single-MLE sumcheck passes)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import scz_b200 as scz  # noqa: E402

ctx = scz.Context(device=0, n_parties=8)
n = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 22)
g = torch.Generator(device=ctx.device).manual_seed(7)
f = torch.randint(-2**63, 2**63 - 1, (n, 4), dtype=torch.int64, device=ctx.device, generator=g)
f[:, 3] &= (1 << 62) - 1
h = f.flip(0).contiguous()
ch = f[:24].clone()
for _ in range(2):
    scz.sumcheck(ctx, f, ch)
    scz.sumcheck_product(ctx, f, h, ch)
    scz.fr_pointwise(ctx, "div", f, h)
ctx.sync()
print("ok")
