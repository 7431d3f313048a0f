"""dev tool: hunts the sporadic slow FIRST timed step of bench.py (one B200, 2^20, leader mode).  Repeats the warm-up ->
barrier -> timed-step transition of bench.py and prints device and host time of the first step after each transition."""
import os
import sys
import threading
import time

os.environ.setdefault("SCZ_MSM_STREAM", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import scz_b200 as scz  # noqa: E402

n = 20
torch.cuda.set_stream(torch.cuda.Stream(priority=-1))
ctx = scz.Context(device=0, n_parties=8)
pp = scz.PackedSharingParams(ctx, 1)
pk = scz.PackedProvingParameters.new(ctx, n, 1, seed=1, shared_seed=0, precompute=True)
torch.cuda.synchronize()
stop = threading.Event()


def nvml_poll():
    import pynvml
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(0)
    while not stop.is_set():
        pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
        pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
        stop.wait(0.1)


for trial in range(int(sys.argv[1]) if len(sys.argv) > 1 else 10):
    mode = ("prof+nvml", "prof", "plain", "nvml")[trial % 4]
    th = None
    stop.clear()
    if "nvml" in mode:
        th = threading.Thread(target=nvml_poll, daemon=True)
        th.start()
    ctx.prof_enable("prof" in mode)
    for _ in range(3):
        scz.dhyperplonk(ctx, n, pk, pp)
    torch.cuda.synchronize()
    ctx.prof_enable("prof" in mode)
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    torch.cuda.synchronize()
    e0.record()
    t0 = time.perf_counter()
    scz.dhyperplonk(ctx, n, pk, pp)
    th1 = time.perf_counter() - t0
    e1.record()
    scz.dhyperplonk(ctx, n, pk, pp)
    e2.record()
    torch.cuda.synchronize()
    print(f"trial {trial} {mode:10s}: first step {e0.elapsed_time(e1):7.1f} ms (host {th1 * 1e3:6.1f} ms), second {e1.elapsed_time(e2):7.1f} ms", flush=True)
    stop.set()
    if th:
        th.join()
ctx.prof_enable(False)
