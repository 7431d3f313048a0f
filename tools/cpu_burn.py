import sys, time, numpy as np
sys.path.insert(0, "/root/repo" if __import__("os").path.exists("/root/repo/oracle") else ".")
from oracle import oracle as orc
orc.lib()
rng = np.random.default_rng(1)
b = orc.random_g1(rng, 1024); b = np.tile(b, (64, 1)); s = orc.random_fr(rng, len(b))
t = time.time()
while time.time() - t < 40:
    orc.msm(b, s, "ark", threads=__import__("os").cpu_count())
print("burn done")
