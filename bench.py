#!/usr/bin/env python
"""bench.py -- the driver's measurement contract for the dist-primitive hot path.

    python bench.py --gpus N --steps K --warmup W            # own arm (libscz.so, sm_100a)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU algorithm (oracle port)

Workload (BASELINE.json configs[1]): d_msm over BLS12-381 G1, 2^20 bases, l = 1 (8 parties).
A step is one d_msm (dist-primitive/src/dmsm.rs:9-43): every party's Pippenger MSM, the
gather -> leader closure (unpack2, sum, pack over G1) -> scatter round.
  N = 1      leader mode (the reference's build without `comm`): ONE party's work per step
  N = 2,4,8  the 8 parties spread over the N GPUs (8/N per GPU), rounds over NCCL
`value` counts reference-equivalent G1 additions: per party m * ceil(255 / c_ark(m)) bucket
additions, the work ark-ec's VariableBaseMSM does on the same input (c_ark = floor(log2 m * 0.69) + 2),
so both arms are scored on the same unit whatever window the implementation picks.
One JSON line on stdout (rank 0).
"""
import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_PARTIES = 8
ALG_BYTES_PER_ADD = 100          # 96 B affine base gather + 4 B sorted point index
HBM_FALLBACK_GBS = 6650.0        # /opt/skills/guides/B200_PROFILING.md fallback


def ark_window(m):
    """ark-ec 0.4.2 VariableBaseMSM window rule"""
    return 3 if m < 32 else int(math.log2(m) * 69 / 100) + 2


def ref_adds(m):
    c = ark_window(m)
    return m * ((255 + c - 1) // c)


def hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        d = json.load(open(path))
        for k in ("hbm_gbs", "hbm_gb_s", "hbm_GBps"):
            if k in d:
                return float(d[k]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        pass
    return HBM_FALLBACK_GBS, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons DURING the timed region"""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [int(r[0]) for r in self.rows if r and r[0].isdigit()]
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower() == "active"})
        return {"sm_mhz": int(statistics.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ------------------------------------------------------------------------------------------ CPU arms
def cpu_inputs(orc, m, seed):
    """bases: a pool of 1024 distinct random G1 points tiled to m (group-law cost does not depend on the
    value; generating 2^20 independent points on the CPU would take minutes), scalars uniform in Fr"""
    import numpy as np
    rng = np.random.default_rng(seed)
    pool = orc.random_g1(rng, 1024)
    bases = np.tile(pool, ((m + 1023) // 1024, 1))[:m].copy()
    return bases, orc.random_fr(rng, m)


def cpu_d_msm_leader_sim(orc, pp, bases, scalars, threads):
    """oracle restatement of dmsm.rs:9-43 in leader mode: local msm, N clones, unpack2, sum, pack, keep share 0"""
    import numpy as np
    c = orc.msm(bases, scalars, "ark", threads=threads)
    shares = np.repeat(c.reshape(1, 18), N_PARTIES, axis=0)
    sec = orc.unpack2(pp, shares, kind=1)
    return orc.pack_from_public(pp, sec, kind=1)[0]


def run_reference(args):
    """--impl reference: the reference cannot be built here (Rust + un-vendored arkworks, no cargo), so this
    times the oracle's restatement of its algorithm (ark-ec signed-digit Pippenger, arkworks' window rule) on
    all host threads.  Rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    from oracle import oracle as orc
    orc.lib()
    m = 1 << args.logn
    threads = os.cpu_count() or 1
    bases, scalars = cpu_inputs(orc, m, 11)
    pp = orc.pp_new(1)
    for _ in range(min(args.warmup, 1)):
        cpu_d_msm_leader_sim(orc, pp, bases, scalars, threads)
    steps = args.steps
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_d_msm_leader_sim(orc, pp, bases, scalars, threads)
    dt = time.perf_counter() - t0
    value = ref_adds(m) * steps / dt
    sample = (f"each step = one party's full d_msm (2^{args.logn} bases, leader-mode closure) with the ark window rule, "
              f"windows spread over {threads} host threads (arkworks `parallel`, which the reference leaves off)")
    line = {
        "impl": "reference", "metric": "d_msm G1-adds/sec", "value": value, "unit": "G1 adds/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": min(args.warmup, 1), "ms_per_step": dt / steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u32x12 Montgomery (Fq), u32x8 (Fr)", "data": "synthetic",
        "config": {"workload": f"d_msm G1 2^{args.logn} bases, l=1, leader mode", "bases_per_party": m,
                   "adds_unit": "m*ceil(255/c_ark) reference-equivalent bucket additions per party"},
        "cpu_baseline": {"value": value, "unit": "G1 adds/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "G1 adds/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def cpu_baseline_leg(logn_sample=16):
    """own arm, N = 1: the oracle port on ONE thread (what the reference does per party: ark `parallel` is off)"""
    from oracle import oracle as orc
    orc.lib()
    m = 1 << logn_sample
    bases, scalars = cpu_inputs(orc, m, 12)
    pp = orc.pp_new(1)
    t0 = time.perf_counter()
    reps = 0
    while True:
        cpu_d_msm_leader_sim(orc, pp, bases, scalars, 1)
        reps += 1
        if time.perf_counter() - t0 > 10.0 or reps >= 8:
            break
    dt = time.perf_counter() - t0
    return {"value": ref_adds(m) * reps / dt, "unit": "G1 adds/s", "cores": 1, "kind": "port",
            "sample": f"{reps} x leader-mode d_msm of 2^{logn_sample} bases (ark window c={ark_window(m)}), 1 thread, "
                      f"{dt:.1f} s of CPU work; oracle C restatement of the arkworks path"}


# ------------------------------------------------------------------------------------------ own arm
def run_own(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import scz_b200 as scz
    from scz_b200.api import msm_batched
    from scz_b200.net import TorchDistNet
    import ctypes as C

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; libscz has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    assert world in (1, 2, 4, 8), "the 8 parties of l=1 spread over 1, 2, 4 or 8 GPUs"
    P = 1 if world == 1 else N_PARTIES // world          # parties hosted by this rank
    m = 1 << args.logn

    net = TorchDistNet(dev) if world > 1 else None
    use_vtable = world == N_PARTIES                         # one party per rank: the C ABI drives NCCL itself
    ctx = scz.Context(device=local_rank, party_id=rank if use_vtable else 0, n_parties=N_PARTIES,
                      net=net if use_vtable else None)
    pp = scz.PackedSharingParams(ctx, 1)

    # synthetic inputs, resident in HBM: per hosted party TWO (bases, scalars) sets used alternately, so a
    # step never re-reads what the previous one left in L2 (one set = 128 MiB > the 126 MB L2 anyway)
    g = torch.Generator(device=dev).manual_seed(0x5CA1AB1E + rank)

    def rand_fr(n):
        s = torch.randint(-2**63, 2**63 - 1, (n, 4), dtype=torch.int64, device=dev, generator=g)
        s[:, 3] &= (1 << 62) - 1      # < 2^254 < r: a valid Montgomery representative of a uniform element
        return s

    sets = []
    for p in range(P):
        for alt in range(2):
            sets.append((ctx.g1_generator_mul(rand_fr(m)), rand_fr(m)))
    torch.cuda.synchronize()

    def step(i):
        """one d_msm of all hosted parties; returns this rank's shares"""
        if world == 1 or use_vtable:
            b, s = sets[i & 1]
            return scz.d_msm(ctx, pp, [b], [s])
        # several parties per rank: dmsm.rs:19-24 per party, then ONE gather / closure / scatter for all of them
        local = torch.cat([msm_batched(ctx, [sets[2 * p + (i & 1)][0]], [sets[2 * p + (i & 1)][1]]) for p in range(P)])
        recv = torch.empty((world, P * 18), dtype=torch.int64, device=dev) if rank == 0 else None
        net.gather_t(local.view(-1), recv.view(-1) if recv is not None else None)
        send = None
        if rank == 0:
            send = torch.empty((N_PARTIES, 18), dtype=torch.int64, device=dev)
            ctx.check(ctx.L.scz_d_msm_leader_dev(ctx.h, pp.h, C.c_void_p(recv.data_ptr()), C.c_size_t(1),
                                                 C.c_void_p(send.data_ptr())))
        out = torch.empty((P, 18), dtype=torch.int64, device=dev)
        net.scatter_t(send.view(-1) if send is not None else None, out.view(-1))
        return out

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for i in range(args.warmup):
        step(i)
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    ctx.prof_enable(True)
    launches0 = ctx.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(args.steps):
        out = step(i)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = ctx.launches - launches0
    acc_ms, acc_n = ctx.prof_read("msm_accumulate")
    prof = {k: ctx.prof_read(k)[0] / args.steps for k in ("msm_sort", "msm_accumulate", "msm_fixup", "msm_reduce",
                                                           "msm_finish", "pss")}
    ctx.prof_enable(False)
    clocks = sampler.stop() if sampler else None
    stats = ctx.msm_last_stats()
    if world > 1:
        t = torch.tensor([ms, float(launches)], dtype=torch.float64, device=dev)
        dist.all_reduce(t[0:1], op=dist.ReduceOp.MAX)
        dist.all_reduce(t[1:2], op=dist.ReduceOp.SUM)
        ms, launches = float(t[0]), int(t[1])
    parties_total = 1 if world == 1 else N_PARTIES
    value = parties_total * ref_adds(m) * args.steps / (ms * 1e-3)

    # ---- end to end: HOST buffers in, host result out, copies inside the timed region
    e2e_steps = max(1, min(args.steps, 5))
    h2d = P * m * (96 + 32)
    d2h = P * 144
    if world == 1:
        hb = torch.empty((m, 12), dtype=torch.int64).pin_memory()
        hs = torch.empty((m, 4), dtype=torch.int64).pin_memory()
        hb.copy_(sets[0][0])
        hs.copy_(sets[0][1])
        nb, ns = hb.numpy().view(np.uint64), hs.numpy().view(np.uint64)
        scz.d_msm(ctx, pp, [nb], [ns])                      # warm
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            host_out = scz.d_msm(ctx, pp, [nb], [ns])       # scz_d_msm: H2D, MSM, closure, D2H, sync
        barrier()
        e2e_s = time.perf_counter() - t0
    else:
        hsets = [(torch.empty((m, 12), dtype=torch.int64).pin_memory().copy_(sets[2 * p][0]),
                  torch.empty((m, 4), dtype=torch.int64).pin_memory().copy_(sets[2 * p][1])) for p in range(P)]

        def e2e_step():
            for p in range(P):
                sets[2 * p][0].copy_(hsets[p][0], non_blocking=True)
                sets[2 * p][1].copy_(hsets[p][1], non_blocking=True)
            return step(0).cpu()
        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            host_out = e2e_step()
        barrier()
        e2e_s = time.perf_counter() - t0
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t[0])
    e2e_value = parties_total * ref_adds(m) * e2e_steps / e2e_s

    if rank != 0:
        ctx.close()
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = hbm_peak()
    acc_launch_ms = acc_ms / max(acc_n, 1)
    adds_per_launch = stats["bucket_adds"]
    achieved = adds_per_launch * ALG_BYTES_PER_ADD / (acc_launch_ms * 1e-3) / 1e9 if acc_launch_ms > 0 else 0.0
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))["k_msm_accumulate"]["dram_bytes"]
    except Exception:
        pass
    line = {
        "metric": "d_msm G1-adds/sec", "value": value, "unit": "G1 adds/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32x12 Montgomery (Fq), u32x8 (Fr)", "data": "synthetic",
        "config": {
            "workload": f"d_msm G1 2^{args.logn} bases, l=1, " + ("leader mode (1 party)" if world == 1 else
                                                                  f"8 parties on {world} GPUs ({P} per GPU), NCCL rounds"),
            "bases_per_party": m, "parties_per_gpu": P,
            "adds_unit": "m*ceil(255/c_ark) reference-equivalent bucket additions per party",
            "kernel_window_bits": int(round(math.log2(max(stats["buckets"] // max(stats["windows"], 1), 1)))) + 1,
            "l2": "two alternating input sets per party (2 x 128 MiB) + ~0.5 GB of sort/bucket temporaries per step: "
                  "nothing survives in the 126 MB L2 between steps",
            "pairs_per_s": parties_total * m * args.steps / (ms * 1e-3),
            "kernel_ms_per_step": prof,
        },
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "G1 adds/s", "h2d_bytes_per_step": h2d * (1 if world == 1 else world),
                "d2h_bytes_per_step": d2h * (1 if world == 1 else world), "steps": e2e_steps,
                "ms_per_step": e2e_s / e2e_steps * 1e3,
                "api": "scz_d_msm (host buffers through the C ABI)" if world == 1 else
                       "pinned host -> device copies + d_msm rounds + device -> host result"},
        "gpu_launches": launches,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "kernel": "k_msm_accumulate", "peak_source": peak_src,
                     "launch_ms": acc_launch_ms, "units_per_launch": adds_per_launch,
                     "alg_bytes_per_unit": ALG_BYTES_PER_ADD,
                     "note": "the bucket kernel is bound by the INT32 multiply pipe, not HBM: ~2.9k IMAD.WIDE per "
                             "mixed add vs ~100 B of traffic; ncu shows sm__pipe_fmaheavy_cycles_active ~80-90 % "
                             "(profiles/), see DESIGN.md"},
    }
    if world == 1 and not args.no_cpu:
        line["cpu_baseline"] = cpu_baseline_leg()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--logn", type=int, default=20)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_own(args)


if __name__ == "__main__":
    main()
