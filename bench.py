#!/usr/bin/env python
"""bench.py -- the driver's measurement contract for the dist-primitive hot path.

    python bench.py --gpus N --steps K --warmup W            # own arm (libscz.so, sm_100a)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU algorithm (oracle port)

Workload (BASELINE.json `metric`, configs[4] on one box): the collaborative HyperPlonk prover
`dhyperplonk` (hyperplonk/src/dhyperplonk.rs:159-571) at 2^20 constraints, l = 1, N = 8 parties, synthetic
random witness / selector / permutation tables (PackedProvingParameters::new, :65-157).  A step is one proof:
what the reference's "Distributed HyperPlonk" timer covers (:194-561) -- ~800 MSMs (24.9 M base/scalar pairs),
~150 product sumchecks, the PST opening folds, 2^19 field inversions, the product tree and ~150 leader rounds.
  N = 1      leader mode (the reference's build without `comm`): ONE party's whole prover on one GPU
  N = 2,4,8  the 8 parties spread over the N GPUs (8/N per GPU), the reference's star rounds over NCCL
`value` = 2^n / t(one proof) at EVERY N (the circuit size over the "Distributed HyperPlonk" timer).
The 8 parties of an l = 1 run jointly prove ONE circuit and each works on full-size share tables, so 8 GPUs give
privacy and 8 provers' worth of work per unit time, not a shorter proof: that party-summed work rate is reported
beside it as `value_party_aggregate`, never as `value`.  Per-GPU work is fixed from N = 1 (one party) to N = 8 (one
party per GPU): "weak".  The d_msm figures BASELINE.json's metric also names (G1 adds/s, HBM roofline fraction of
the Pippenger bucket kernel) are measured inside the same step and reported in `d_msm` and `roofline`.  For N > 1
an untimed leg after the measurement proves a 2^10 circuit over the same live NCCL net and compares every party's
proof with the oracle's 8-party run (`parity_check`).  One JSON line on stdout (rank 0).
"""
import os as _os
_os.environ.setdefault("SCZ_MSM_STREAM", "1")   # MSM launch sequences on the ctx's low-priority stream (csrc/msm.cu)
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
METRIC = "HyperPlonk constraints/sec"
UNIT = "constraints/s"
DTYPE = "u32x8 Montgomery (Fr), u32x12 Montgomery (Fq): integer only"


def ark_window(m):
    """ark-ec 0.4.2 VariableBaseMSM window rule"""
    return 3 if m < 32 else int(math.log2(m) * 69 / 100) + 2


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


def pool_reserved_mib(device_index):
    """MiB the device's stream-ordered memory pool holds (host-side allocator state; None without cuda-python).  Read before
    and after the timed region: a pool that still grows inside it costs ~0.1 ms of stalled stream per MiB (tools/pool_probe.py)"""
    try:
        from cuda.bindings import driver as cu
        err, dev = cu.cuDeviceGet(device_index)
        err, pool = cu.cuDeviceGetDefaultMemPool(dev)
        err, v = cu.cuMemPoolGetAttribute(pool, cu.CUmemPool_attribute.CU_MEMPOOL_ATTR_RESERVED_MEM_CURRENT)
        return int(v) >> 20 if int(err) == 0 else None
    except Exception:
        return None


class ClockSampler:
    """SM clock + throttle reasons DURING the timed region.  In-process NVML (two light queries every 500 ms from a
    thread -- every query takes driver locks the launch path also needs, so the rate is kept low); `nvidia-smi --query-gpu ... -lms` is the fallback -- its full query holds driver locks long enough to slow
    a multi-threaded launcher by ~10 % (measured at N = 2: 755 vs 682 ms per step), NVML's two calls do not."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc, self.nvml = index, [], None, None
        self.t0 = self.t1 = None
        self._stop = threading.Event()

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index())
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM)
            self.thread = threading.Thread(target=self._poll_nvml, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "500"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _physical_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [x.strip() for x in vis.split(",") if x.strip()]
            if self.index < len(ids) and ids[self.index].isdigit():
                return int(ids[self.index])
        return self.index

    def _poll_nvml(self):
        n = self.nvml
        bits = [(getattr(n, "nvmlClocksEventReasonHwSlowdown", 0x8), "hw_slowdown"),
                (getattr(n, "nvmlClocksEventReasonHwThermalSlowdown", 0x40), "hw_thermal_slowdown"),
                (getattr(n, "nvmlClocksEventReasonSwThermalSlowdown", 0x20), "sw_thermal_slowdown"),
                (getattr(n, "nvmlClocksEventReasonSwPowerCap", 0x4), "sw_power_cap")]
        get_reasons = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or n.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self._stop.is_set():
            try:
                mhz = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
                r = get_reasons(self.handle)
                row = [str(mhz), str(self.max_mhz)] + ["Active" if r & b else "Not Active" for b, _ in bits]
                self.rows.append((time.perf_counter(), row))
            except Exception:
                pass
            self._stop.wait(0.5)

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [x.strip() for x in line.split(",")]))

    def wait_first(self, timeout=5.0):
        t = time.perf_counter()
        while (self.proc or self.nvml) and not self.rows and time.perf_counter() - t < timeout:
            time.sleep(0.02)

    def mark_begin(self):
        self.t0 = time.perf_counter()

    def mark_end(self):
        self.t1 = time.perf_counter()

    def stop(self):
        if not self.proc and not self.nvml:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml / nvidia-smi unavailable"], "samples": 0}
        self._stop.set()
        if self.proc:
            time.sleep(0.1)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        else:
            self.thread.join(timeout=1)
        inside = [r for t, r in self.rows if self.t0 is not None and self.t0 <= t <= (self.t1 or t) + 0.55]
        rows = inside or [r for _, r in self.rows]
        sm = [int(r[0]) for r in rows if r and r[0].isdigit()]
        mx = [int(r[1]) for r in rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in rows if len(r) >= 6 for i in range(4) if r[2 + i].lower() == "active"})
        return {"sm_mhz": int(statistics.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm), "samples_in_timed_region": len(inside),
                "source": "nvml" if self.nvml else "nvidia-smi"}


# ------------------------------------------------------------------------------------------ CPU arms
def cpu_hyperplonk_inputs(n, seed, l=1):
    """oracle pk at circuit size 2^n: random tables; SRS levels tile a pool of 1024 random G1 points (group-law
    cost does not depend on the values; drawing 2^(n+3) independent points on the CPU would take minutes)"""
    import numpy as np
    from oracle import hyperplonk as ohp
    from oracle import oracle as orc
    rng = np.random.default_rng(seed)
    pool = orc.random_g1(rng, 1024)

    def level(m):
        return np.tile(pool, ((m + 1023) // 1024, 1))[:m].copy()
    csz, dsz = ohp.srs_level_sizes(n, l, 8 * l)
    srs_c = orc.Srs.from_levels([level(m) for m in csz])
    srs_d = orc.Srs.from_levels([level(m) for m in dsz])
    return ohp.random_pk(rng, n, l, 8 * l, srs_c, srs_d)


def cpu_hyperplonk(n, pk, threads, l=1):
    """one leader-mode proof with the oracle's restatement of dhyperplonk (oracle/hyperplonk.py)"""
    from oracle import hyperplonk as ohp
    from oracle import oracle as orc
    orc.set_msm_threads(threads)
    t0 = time.perf_counter()
    ohp.dhyperplonk(n, [pk], orc.pp_new(l), orc.LEADER_SIM, 8 * l)
    dt = time.perf_counter() - t0
    orc.set_msm_threads(1)
    return dt


def run_reference(args):
    """--impl reference: the reference cannot be built here (Rust nightly + un-vendored arkworks 0.4, no cargo), so this
    times the oracle's restatement of its algorithm (arkworks' signed-digit Pippenger with arkworks' window rule,
    the same protocol schedule) with the MSM windows spread over all host threads.  Rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    from oracle import oracle as orc
    orc.lib()
    threads = os.cpu_count() or 1
    n = args.ref_logn or args.logn
    l = args.l
    pk = cpu_hyperplonk_inputs(n, 11, l)
    for _ in range(min(args.warmup, 1)):
        cpu_hyperplonk(min(n, 10), cpu_hyperplonk_inputs(min(n, 10), 12, l), threads, l)
    steps = max(1, min(args.steps, args.ref_steps))
    dt = sum(cpu_hyperplonk(n, pk, threads, l) for _ in range(steps))
    value = (1 << n) * steps / dt
    same = n == args.logn
    sample = (f"each step = one leader-mode dhyperplonk proof at 2^{n} constraints "
              + ("(the own arm's config, whole proof) " if same else f"(bounded sample of the 2^{args.logn} workload) ")
              + f"-- {steps} timed proof(s): a proof takes about a minute of CPU time at 2^20; oracle C restatement of the "
              f"arkworks path, MSM windows on {threads} host threads (arkworks `parallel`, which the reference leaves "
              f"off); everything else single-threaded like the reference")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": min(args.warmup, 1), "ms_per_step": dt / steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": DTYPE, "data": "synthetic",
        "config": {"workload": f"dhyperplonk 2^{args.logn} constraints, l={args.l}, N={8 * args.l}, leader mode (one party's prover)",
                   "packing_factor_l": args.l,
                   "log2_constraints": n, "same_config_as_own_arm": same},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit_line(line)


def cpu_baseline_leg(n=14, l=1):
    """own arm, N = 1: the oracle port on ONE thread (what the reference does per party: ark `parallel` is off)"""
    from oracle import oracle as orc
    orc.lib()
    pk = cpu_hyperplonk_inputs(n, 12, l)
    dt = cpu_hyperplonk(n, pk, 1, l)
    return {"value": (1 << n) / dt, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"one leader-mode dhyperplonk proof at 2^{n} constraints (l={l}, N={8 * l}), 1 thread, {dt:.1f} s of CPU work; "
                      f"oracle C restatement of the arkworks path (portable C field arithmetic, no assembly backend)"}


def fr_kernel_rooflines(scz, ctx, torch, peak):
    """the HBM-side kernels of the path, each timed alone with CUDA events on tables larger than L2: algorithmic bytes
    (the UNFUSED per-round figures) / time against the measured HBM peak.  Two sizes: 2^22 entries
    (128 MiB; a whole call is ~0.1-0.6 ms there, of which ~0.1 ms is the host issuing its ~15 launches and stream-ordered
    allocations -- the calls are launch-bound as much as bandwidth-bound) and 2^24 entries (512 MiB), where the
    kernels themselves dominate."""
    C = __import__("ctypes")
    vp = lambda t: C.c_void_p(t.data_ptr())   # noqa: E731
    g = torch.Generator(device=ctx.device).manual_seed(7)

    def rand_fr(m):
        t = torch.randint(-2**63, 2**63 - 1, (m, 4), dtype=torch.int64, device=ctx.device, generator=g)
        t[:, 3] &= (1 << 62) - 1
        return t
    res = []
    for logn in (22, 24):
        n = 1 << logn
        f, h, ch = rand_fr(n), rand_fr(n), rand_fr(26)
        q, val = ctx.empty(n, 4), ctx.empty(1, 4)
        out3, last = ctx.empty(3 * 26, 4), ctx.empty(2, 4)
        cases = {
            f"open_fold (dpoly_comm.rs:309-323), all {logn} rounds": (
                lambda: ctx.check(ctx.L.scz_open_fold_dev(ctx.h, vp(f), C.c_size_t(n), vp(ch), vp(q), vp(val))), (n - 1) * 128),
            f"product sumcheck rounds (dsumcheck.rs:37-85), all {logn} rounds": (
                lambda: ctx.check(ctx.L.scz_sumcheck_product_rounds_dev(ctx.h, vp(f), vp(h), C.c_size_t(n), vp(ch), vp(out3), vp(last))),
                (n - 1) * 192),
            f"single-MLE sumcheck (dsumcheck.rs:6-26), all {logn} rounds": (
                lambda: ctx.check(ctx.L.scz_sumcheck_dev(ctx.h, vp(f), C.c_size_t(n), vp(ch), vp(out3))), (n - 1) * 96),
            "pointwise a + k0*b + k1 (dhyperplonk.rs:326-337)": (
                lambda: ctx.check(ctx.L.scz_fr_pointwise_dev(ctx.h, 2, vp(f), vp(h), vp(ch), vp(q), C.c_size_t(n))), n * 96),
            "division num/den, one shared inversion (dhyperplonk.rs:339)": (
                lambda: ctx.check(ctx.L.scz_fr_pointwise_dev(ctx.h, 3, vp(f), vp(h), None, vp(q), C.c_size_t(n))), n * 96),
            "fix_variable, 2 variables (mle.rs:88-104)": (
                lambda: ctx.check(ctx.L.scz_fix_variable_dev(ctx.h, vp(f), C.c_size_t(n), vp(ch), C.c_size_t(2), vp(q))),
                (n // 2) * 96 + (n // 4) * 96),
        }
        for name, (fn, nbytes) in cases.items():
            for _ in range(3):
                fn()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(5):
                fn()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            gbs = nbytes / (ms * 1e-3) / 1e9
            res.append({"kernel": name, "table_entries": n, "alg_bytes": nbytes, "ms": ms, "achieved_gbs": gbs,
                        "frac_of_hbm_peak": gbs / peak})
        del f, h, q
    return res


# ------------------------------------------------------------------------------------------ BASELINE configs 2, 3, 4
def timed_collective(torch, dist, world, dev, run_all, reps, warm=2):
    """device time per call of `run_all` (every hosted party of every rank makes the call once): CUDA events around
    `reps` calls on this rank's stream, barriers on both sides, max over ranks"""
    for _ in range(warm):
        run_all()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        run_all()
    b.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = a.elapsed_time(b) / reps
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t[0])
    return ms


def config2_d_msm(scz, ctx, pp, pk, torch, peak, imad_peak, with_cpu):
    """BASELINE config 2: leader-mode d_msm (dmsm.rs:9-43), one MSM of 2^20 G1 bases, l = 1, on one GPU.  The bases are
    call arguments of d_msm, so the plain Pippenger path runs (level 20 of the SRS serves as 2^20 random points);
    c_commit's variant with the level's fixed-base table is timed beside it."""
    import numpy as np
    n = 1 << 20
    bases = pk.c_commitment.level(20)
    scalars = pk.t["a_evals"]
    assert len(bases) == n and len(scalars) == n
    out = {}
    for name, fn in (("d_msm (plain bases)", lambda: scz.d_msm(ctx, pp, [bases], [scalars])),
                     ("c_commit (fixed-base table of the SRS level)", lambda: pk.c_commitment.c_commit(pp, [scalars]))):
        ms = timed_collective(torch, None, 1, ctx.device, fn, reps=5)
        a0 = ctx.msm_cum_stats()["bucket_adds"]
        ctx.prof_enable(True)
        fn()
        acc_ms = ctx.prof_read("msm_accumulate")[0]
        ctx.prof_enable(False)
        adds = ctx.msm_cum_stats()["bucket_adds"] - a0
        out[name] = {"ms": ms, "pairs_per_s": n / (ms * 1e-3), "bucket_adds": adds, "g1_adds_per_s": adds / (ms * 1e-3),
                     "bucket_kernel_ms": acc_ms, "bucket_kernel_adds_per_s": adds / (acc_ms * 1e-3) if acc_ms else None,
                     "bucket_kernel_hbm_frac": adds * ALG_BYTES_PER_ADD / (acc_ms * 1e-3) / 1e9 / peak if acc_ms else None,
                     "bucket_kernel_alg_bytes_per_add": ALG_BYTES_PER_ADD}
    res = {"workload": "BASELINE config 2: d_msm G1 2^20 bases, l=1, leader mode, 1 GPU (dmsm.rs:9-43; examples/msm.rs:17-101)",
           "gpu": out}
    if with_cpu:
        from oracle import oracle as orc
        rng = np.random.default_rng(5)
        m = 1 << 18                                           # bounded sample: a quarter of the size, all host threads
        pool = orc.random_g1(rng, 1024)
        b = np.tile(pool, (m // 1024, 1))
        sc = orc.random_fr(rng, m)
        thr = os.cpu_count() or 1
        t0 = time.perf_counter()
        orc.msm(b, sc, "ark", threads=thr)
        dt = time.perf_counter() - t0
        t0 = time.perf_counter()
        orc.msm(b[: m // 4], sc[: m // 4], "ark", threads=1)
        dt1 = time.perf_counter() - t0
        res["cpu_port"] = {"pairs_per_s_all_threads": m / dt, "threads": thr, "pairs_per_s_one_thread": (m // 4) / dt1,
                           "sample": f"arkworks-style signed-digit Pippenger (oracle port) on 2^18 pairs with {thr} threads and on "
                                     "2^16 pairs with 1 thread (the reference runs G::msm on one thread per party)"}
    return res


def config34(scz, torch, dist, world, dev, parties, run_each, peak, n):
    """BASELINE configs 3 and 4 on the live net (N parties over `world` GPUs): d_sumcheck_product on 2 x 2^(n-3) plain
    slices per party (20 variables in total at n = 20, N = 8: dsumcheck.rs:359-512, examples/sumcheck.rs:94-265) and
    c_commit + c_open of one 2^n-share table per party (dpoly_comm.rs:244-267, 401-464; examples/poly_comm.rs:36-201)."""
    res = {}
    m = 1 << (n - 3)
    ms = timed_collective(torch, dist, world, dev,
                          lambda: run_each(lambda ctx, pp, pk: scz.d_sumcheck_product(ctx, pk.t["I_p"], pk.t["S1_p"], pk.t["challenge"])), reps=10)
    alg = (m - 1) * 192
    res["config3_d_sumcheck_product"] = {
        "workload": f"BASELINE config 3: d_sumcheck_product, {n} variables in total, N=8 parties on {world} GPU(s), 2 x 2^{n - 3} "
                    f"Fr per party, {n} challenges, one gather of {n - 3 + 1} triples (96 B each) to the leader + 3 leader rounds",
        "ms_per_call": ms, "constraints_per_s": (1 << n) / (ms * 1e-3), "alg_bytes_per_party": alg,
        "hbm_frac_per_gpu": alg * (len(parties)) / (ms * 1e-3) / 1e9 / peak,
        "note": "latency-bound at this size: 17 rounds of halving tables (the last 9 inside one CTA) + the gather; the HBM fraction "
                "only says how far from a bandwidth problem a 4 MiB table is"}
    ms_c = timed_collective(torch, dist, world, dev,
                            lambda: run_each(lambda ctx, pp, pk: pk.c_commitment.c_commit(pp, [pk.t["a_evals"]])), reps=5)
    ms_o = timed_collective(torch, dist, world, dev,
                            lambda: run_each(lambda ctx, pp, pk: pk.c_commitment.c_open(pp, pk.t["a_evals"], pk.t["challenge"])), reps=5)
    res["config4_dpoly_comm"] = {
        "workload": f"BASELINE config 4: c_commit and c_open of one 2^{n}-share table per party, l=1, N=8 parties on {world} GPU(s) "
                    "(one d_msm of 2^20 points; 20 fold rounds + one batched d_msm of 2^20 - 1 points + pss2ss)",
        "c_commit_ms": ms_c, "c_open_ms": ms_o, "c_commit_coeffs_per_s": (1 << n) / (ms_c * 1e-3),
        "c_open_coeffs_per_s": (1 << n) / (ms_o * 1e-3)}
    return res


def config34_cpu(n):
    """the oracle port of configs 3 and 4 on the host (bounded samples), rank 0 only"""
    import numpy as np
    from oracle import oracle as orc
    orc.lib()
    rng = np.random.default_rng(6)
    thr = os.cpu_count() or 1
    out = {}
    m = 1 << (n - 3)
    f = [orc.random_fr(rng, m) for _ in range(N_PARTIES)]
    g = [orc.random_fr(rng, m) for _ in range(N_PARTIES)]
    ch = orc.random_fr(rng, n)
    t0 = time.perf_counter()
    orc.d_sumcheck_product(orc.PARTIES, N_PARTIES, f, g, ch)
    dt = time.perf_counter() - t0
    out["config3_d_sumcheck_product"] = {"s_all_8_parties_one_thread": dt, "s_per_party": dt / N_PARTIES,
                                         "constraints_per_s_one_party_per_core": (1 << n) / (dt / N_PARTIES),
                                         "sample": "oracle port, the 8 parties' local rounds run one after the other on one thread; "
                                                   "the reference runs them on 8 machines, so the per-party time is the comparable figure"}
    ns = min(n, 18)                                             # bounded sample of config 4
    pool = orc.random_g1(rng, 1024)
    levels = [np.tile(pool, (((1 << i) + 1023) // 1024, 1))[: 1 << i].copy() for i in range(ns + 1)]
    srs = orc.Srs.from_levels(levels)
    p = orc.random_fr(rng, 1 << ns)
    u = orc.random_fr(rng, ns)
    orc.set_msm_threads(thr)
    t0 = time.perf_counter()
    orc.c_commit([srs], orc.pp_new(1), orc.LEADER_SIM, [[p]])
    dc = time.perf_counter() - t0
    t0 = time.perf_counter()
    orc.c_open([srs], orc.pp_new(1), orc.LEADER_SIM, [p], u)
    do = time.perf_counter() - t0
    orc.set_msm_threads(1)
    out["config4_dpoly_comm"] = {"c_commit_coeffs_per_s": (1 << ns) / dc, "c_open_coeffs_per_s": (1 << ns) / do, "threads": thr,
                                 "sample": f"oracle port, leader mode, 2^{ns} shares (bounded sample), MSM windows on {thr} threads"}
    return out


# ------------------------------------------------------------------------------------------ own arm
def run_own(args):
    import numpy as np  # noqa: F401
    import torch
    import torch.distributed as dist
    import scz_b200 as scz
    from scz_b200.net import HybridNet, NativeNcclNet

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; libscz has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    # everything below runs on ONE high-priority stream: the provers' short protocol kernels must outrank the MSM
    # launch sequences, which libscz puts on a lowest-priority stream of each ctx (SCZ_MSM_STREAM, csrc/msm.cu)
    torch.cuda.set_stream(torch.cuda.Stream(device=dev, priority=-1))
    assert world in (1, 2, 4, 8), "the 8 l parties spread over 1, 2, 4 or 8 GPUs"
    L_PACK = args.l                                      # packing factor l: N = 8 l parties, each holds 1 / l of every table
    N_PARTIES = 8 * L_PACK
    P = 1 if world == 1 else N_PARTIES // world          # parties hosted by this rank
    n = args.logn

    sampler = ClockSampler(local_rank) if rank == 0 and not os.environ.get("SCZ_BENCH_NO_CLOCKS") else None
    if sampler:
        sampler.start()

    # ---- setup (untimed, like PackedProvingParameters::new in the reference's bench binary)
    # N > 1: the star rounds are issued by libscz itself (NCCL hub, csrc/nccl_net.cu); SCZ_BENCH_NET=python routes them
    # through the torch.distributed callbacks of net.py instead (same proofs, for A/B measurements)
    net_kind = os.environ.get("SCZ_BENCH_NET", "native")
    hub = (NativeNcclNet(dev, P) if net_kind == "native" else HybridNet(dev, P)) if world > 1 else None
    parties = []
    for p in range(P):
        pid = rank * P + p
        ctx = scz.Context(device=local_rank, party_id=pid if world > 1 else 0, n_parties=N_PARTIES,
                          net=hub.party(p) if hub else None)
        if hub:
            hub.adopt(p, ctx)   # parties that share a GPU run on their own streams (net.py, HybridNet)
        pp = scz.PackedSharingParams(ctx, L_PACK)
        pk = scz.PackedProvingParameters.new(ctx, n, L_PACK, seed=1 + pid, shared_seed=0, precompute=not args.no_precompute)
        parties.append((ctx, pp, pk))
    torch.cuda.synchronize()

    def prove_all():
        """one proof by every hosted party (threads only when several parties share this GPU)"""
        if P == 1:
            ctx, pp, pk = parties[0]
            return [scz.dhyperplonk(ctx, n, pk, pp)]
        return hub.run_parties(lambda pid, p, net: scz.dhyperplonk(parties[p][0], n, parties[p][2], parties[p][1]))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # the per-class event brackets are on during the warm-up too: the first profiled proof creates ~2 600 CUDA events
    # (0.4 s measured inside the first timed step when they were only switched on after the warm-up)
    for ctx, _, _ in parties:
        ctx.prof_enable(True)
    def run_steps(k):
        """k proofs back to back on this rank's stream, an event after every step (no synchronisation); the warm-up goes
        through the very same statements as the timed region, so nothing in it is executed for the first time there"""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        marks, host_ms, proofs = [torch.cuda.Event(enable_timing=True) for _ in range(k)], [], None
        barrier()
        e0.record()
        for i in range(k):
            t0 = time.perf_counter()
            proofs = prove_all()
            marks[i].record()
            host_ms.append(round((time.perf_counter() - t0) * 1e3, 2))
        e1.record()
        barrier()
        return e0, e1, marks, host_ms, proofs

    run_steps(args.warmup)
    ctx0 = parties[0][0]
    # every bracket event of the timed region exists (and has been recorded once) before it starts -- a proof records ~1 300
    # brackets; without this the host creates 2 600 events per proof from step `warmup` + 1 on, while the first proofs are
    # still executing -- and the cyclic GC stays out of it.  Background: with the profiled brackets on, the host runs only
    # ~one proof ahead of the GPU (host_enqueue_ms_each_step_rank0), so a host stall longer than that leaves a hole in the
    # timeline.  One run in four had a 300 - 400 ms hole in timed step 0 or 1; the one cause found is the device memory pool
    # (a stream-ordered allocation that cannot reuse its block maps new memory at ~0.1 ms per MiB; the pool's
    # timing-dependent reuse policy is now off, csrc/ctx.cu), its size is reported around the timed region
    per_proof = max(sum(ctx0.prof_read(k)[1] for k in ctx0.KERNEL_CLASSES) // max(1, args.warmup), 1)
    for ctx, _, _ in parties:
        ctx.prof_enable(True)          # clears the warm-up's records, keeps the event pool
        ctx.prof_reserve(2 * (per_proof + 64) * (args.steps + 1))
    import gc
    gc.collect()
    gc.disable()
    launches0 = sum(c.launches for c, _, _ in parties)
    pool0 = pool_reserved_mib(local_rank)
    coll0 = dict(hub.calls) if hub else {}
    stats0 = ctx0.msm_cum_stats()
    comm0 = ctx0.get_comm()
    if sampler:
        sampler.wait_first()
        sampler.mark_begin()
    e0, e1, marks, host_enqueue_ms, proofs = run_steps(args.steps)
    gc.enable()
    pool1 = pool_reserved_mib(local_rank)
    step_ms = [round((marks[i - 1] if i else e0).elapsed_time(marks[i]), 2) for i in range(len(marks))]
    if sampler:
        sampler.mark_end()
    ms = e0.elapsed_time(e1)
    launches = sum(c.launches for c, _, _ in parties) - launches0
    coll = {k: (hub.calls[k] - coll0[k]) // args.steps for k in coll0} if hub else None
    stats1 = ctx0.msm_cum_stats()
    comm1 = ctx0.get_comm()
    prof = {k: ctx0.prof_read(k) for k in ctx0.KERNEL_CLASSES}
    for ctx, _, _ in parties:
        ctx.prof_enable(False)
    clocks = sampler.stop() if sampler else None
    if world > 1:
        t = torch.tensor([ms, float(launches)], dtype=torch.float64, device=dev)
        dist.all_reduce(t[0:1], op=dist.ReduceOp.MAX)
        dist.all_reduce(t[1:2], op=dist.ReduceOp.SUM)
        ms, launches = float(t[0]), int(t[1])
    parties_total = 1 if world == 1 else N_PARTIES
    value = (1 << n) * args.steps / (ms * 1e-3)          # 2^n / t(proof): the parties prove ONE circuit together

    # ---- the same proof with the fixed-base tables of the SRS ignored (plain Pippenger on the level's points)
    plain_ms = None
    if not args.no_precompute and not args.no_plain:
        for ctx, _, _ in parties:
            ctx.msm_use_precompute(False)
        prove_all()
        barrier()
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0.record()
        for _ in range(2):
            prove_all()
        p1.record()
        barrier()
        plain_ms = p0.elapsed_time(p1) / 2
        if world > 1:
            t = torch.tensor([plain_ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            plain_ms = float(t[0])
        for ctx, _, _ in parties:
            ctx.msm_use_precompute(True)

    # ---- end to end: the witness / selector / challenge tables start in pinned HOST memory every proof, the proof
    #      ends in host memory; the SRS (proving key, reused across proofs) stays resident.  Two resident table sets per
    #      party: the host -> device copy of proof i+1 runs on a copy stream under the MSM phase of proof i (both inside
    #      the timed region); the device -> host read of proof i is queued right behind it on a third stream and collected
    #      by the host after proof i + 1 has been enqueued (scz.ProofReader), so the GPU queue never runs dry between
    #      proofs.  Every step moves its own inputs in and its own proof out; the wall clock covers all of it, including
    #      the first proof's cold start and the last proof's read-back.
    e2e_steps = max(2, min(args.steps, 20))
    host_tabs, alt = [], []
    for ctx, pp, pk in parties:
        host_tabs.append({name: torch.empty(t.shape, dtype=torch.int64).pin_memory().copy_(t) for name, t in pk.t.items()})
        alt.append(scz.PackedProvingParameters(ctx, n, L_PACK, {k: v.clone() for k, v in pk.t.items()}, pk.c_commitment,
                                               pk.d_commitment))
    h2d = sum(t.numel() * 8 for t in host_tabs[0].values())
    readers = [scz.ProofReader(ctx, depth=2) for ctx, _, _ in parties]

    def e2e_loop(p, steps):
        ctx, pp, pk = parties[p]
        sets = [pk, alt[p]]
        main, copy = ctx.stream, torch.cuda.Stream()
        up = [torch.cuda.Event(), torch.cuda.Event()]
        with torch.cuda.stream(copy):
            sets[0].upload(host_tabs[p])
            up[0].record(copy)
        out, pending = None, None
        for i in range(steps):
            main.wait_event(up[i % 2])                       # this proof's inputs are in HBM
            proof = scz.dhyperplonk(ctx, n, sets[i % 2], pp)
            if i + 1 < steps:
                # the next proof's tables: behind the END OF THIS PROOF'S PROTOCOL PHASE, i.e. under its MSM phase.  (The
                # other table set was last read by proof i - 1, which has finished by then.)  Started at a proof boundary the
                # copy shares PCIe with the command fetches of ~1 100 short launches: +6 ms per proof, tools/e2e_probe.py
                ctx.stream_wait_protocol_phase(copy)
                with torch.cuda.stream(copy):
                    sets[(i + 1) % 2].upload(host_tabs[p])
                    up[(i + 1) % 2].record(copy)
            # device -> host: queued behind the proof on the reader's stream; the host collects proof i - 1 now, while
            # proof i is already enqueued
            ticket = proof.to_host_async(readers[p])
            if pending is not None:
                out = readers[p].collect(pending)
            pending = ticket
        return readers[p].collect(pending)

    def e2e_all(steps):
        if P == 1:
            return [e2e_loop(0, steps)]
        return hub.run_parties(lambda pid, p, net: e2e_loop(p, steps))
    out = e2e_all(2)
    d2h = sum(a.nbytes for a in out[0])
    barrier()
    t0 = time.perf_counter()
    e2e_all(e2e_steps)
    barrier()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t[0])
    e2e_value = (1 << n) * e2e_steps / e2e_s

    # one proof with nothing overlapped: host tables -> HBM, prove, proof -> host, strictly one after the other (the latency of
    # a single request with the proving key resident; `e2e.value` above is the steady state of a stream of requests)
    def single(p):
        ctx, pp, pk = parties[p]
        t0 = time.perf_counter()
        pk.upload(host_tabs[p])
        torch.cuda.current_stream().synchronize()
        proof = scz.dhyperplonk(ctx, n, pk, pp)
        proof.to_host()
        return time.perf_counter() - t0

    def single_all():
        if P == 1:
            return [single(0)]
        return hub.run_parties(lambda pid, p, net: single(p))
    barrier()
    single_s = max(max(single_all()) for _ in range(3))
    if world > 1:
        t = torch.tensor([single_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        single_s = float(t[0])
    comm = ((comm1[0] - comm0[0]) // args.steps, (comm1[1] - comm0[1]) // args.steps)

    # ---- BASELINE configs 3 and 4 as stand-alone calls on the same net (every N); config 2 at N = 1 below
    def run_each(fn):
        if P == 1:
            c_, pp_, pk_ = parties[0]
            return [fn(c_, pp_, pk_)]
        return hub.run_parties(lambda pid, p, net: fn(*parties[p]))
    peak_early, _ = hbm_peak()
    standalone = None
    if not args.no_standalone and n >= 6 and L_PACK == 1:
        standalone = config34(scz, torch, dist, world, dev, parties, run_each, peak_early, n)

    # ---- N = 1 only: TWO independent provers in flight on the same GPU (a proving service's steady state).  Each has
    #      its own ctx, high-priority stream and host thread; the MSM launch sequences run on the ctxs' low-priority
    #      streams (SCZ_MSM_STREAM), so one prover's short protocol kernels are dispatched ahead of the other's queued
    #      bucket-kernel CTAs.  Reported beside `value` (which stays one proof at a time), never instead of it.
    pipelined = None
    if world == 1 and L_PACK == 1 and not args.no_pipelined and os.environ.get("SCZ_MSM_STREAM") == "1":
        ctxA, ppA, pkA = parties[0]
        sA, sB = torch.cuda.Stream(priority=-1), torch.cuda.Stream(priority=-1)
        with torch.cuda.stream(sA):
            ctxA.use_torch_stream()
        with torch.cuda.stream(sB):
            ctxB = scz.Context(device=local_rank, party_id=0, n_parties=N_PARTIES)
            ppB = scz.PackedSharingParams(ctxB, 1)
            pkB = scz.PackedProvingParameters.new(ctxB, n, 1, seed=99, shared_seed=0, precompute=not args.no_precompute)
        torch.cuda.synchronize()
        provers = [(sA, ctxA, ppA, pkA), (sB, ctxB, ppB, pkB)]

        def in_flight(reps):
            bar = threading.Barrier(3)
            ends = [torch.cuda.Event(), torch.cuda.Event()]

            def body(i):
                st, c, pp_, pk_ = provers[i]
                torch.cuda.set_device(local_rank)
                with torch.cuda.stream(st):
                    bar.wait()
                    t_host = time.perf_counter()
                    for _ in range(reps):
                        scz.dhyperplonk(c, n, pk_, pp_)
                    ends[i].record()
                    t_enq = time.perf_counter() - t_host
                    st.synchronize()
                    print(f"[pipelined] prover {i}: {reps} proofs enqueued in {t_enq * 1e3:.0f} ms, done after "
                          f"{(time.perf_counter() - t_host) * 1e3:.0f} ms", file=sys.stderr, flush=True)
            ts = [threading.Thread(target=body, args=(i,)) for i in range(2)]
            for t_ in ts:
                t_.start()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for st, _, _, _ in provers:
                st.wait_event(a)
            bar.wait()
            for t_ in ts:
                t_.join()
            for e in ends:
                torch.cuda.current_stream().wait_event(e)
            b.record()
            torch.cuda.synchronize()
            return a.elapsed_time(b)
        in_flight(3)   # warm-up: the shared memory pool has to reach the two-prover peak before the timed region
        pms = in_flight(args.steps)
        pipelined = {"in_flight": 2, "proofs": 2 * args.steps, "ms_per_proof": pms / (2 * args.steps),
                     "value": 2 * args.steps * (1 << n) / (pms * 1e-3), "unit": UNIT,
                     "note": "two independent provers (own ctx, stream, host thread) share the GPU; device time from the first "
                             "launch to the last kernel of all proofs (CUDA events)"}
        ctxB.close()
        ctxA.use_torch_stream()

    # ---- N > 1: untimed parity leg -- a 2^10 proof by the same 8 parties over the live NCCL net, every party's output
    #      bit for bit against the oracle's 8-party run (rank 0 checks; tests/parity_util.py)
    parity = None
    if world > 1 and not args.no_parity:
        from tests.parity_util import nccl_parity_check
        try:
            parity = nccl_parity_check(local_rank, nv=args.parity_logn, l=L_PACK, net_kind=net_kind)
        except Exception as e:   # noqa: BLE001 - reported in the JSON line
            parity = {"result": f"ERROR: {e!r}"[:400]}

    if rank != 0:
        for ctx, _, _ in parties:
            ctx.close()
        if hub:
            hub.close()
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = hbm_peak()
    acc_ms, acc_n = prof["msm_accumulate"]
    adds = stats1["bucket_adds"] - stats0["bucket_adds"]        # party 0's bucket additions in the timed region
    pairs = stats1["pairs"] - stats0["pairs"]
    msm_ms = sum(prof[k][0] for k in ("msm_sort", "msm_accumulate", "msm_fixup", "msm_reduce", "msm_finish"))
    achieved = adds * ALG_BYTES_PER_ADD / (acc_ms * 1e-3) / 1e9 if acc_ms > 0 else 0.0
    affine_ran = ctx0.msm_affine_sequences() > 0
    acc_kernel = ("bucket accumulation: k_ba_levels / k_ba_phase1 / k_inv_tree_* / k_ba_phase2 / k_ba_accumulate (csrc/msm_affine.cu)"
                  if affine_ran else "k_msm_accumulate")
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))[
            "msm_accumulate_affine" if affine_ran else "k_msm_accumulate"]
    except Exception:
        pass
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": DTYPE, "data": "synthetic",
        "config": {
            "workload": f"dhyperplonk 2^{n} constraints, l={L_PACK}, N={N_PARTIES}, " + (
                "leader mode (one party's prover)" if world == 1 else
                f"{N_PARTIES} parties on {world} GPUs ({P} per GPU), star rounds over NCCL"),
            "log2_constraints": n, "parties_per_gpu": P, "packing_factor_l": L_PACK,
            "value_counts": "2^n / t(one proof), BASELINE.md 3 -- at N > 1 the 8 parties prove ONE circuit together; "
                            "the party-summed work rate is value_party_aggregate",
            "value_party_aggregate": parties_total * value,
            "proofs_per_s": args.steps / (ms * 1e-3),
            "ms_each_step_rank0": step_ms,
            "ms_per_step_median_rank0": sorted(step_ms)[len(step_ms) // 2],
            "host_enqueue_ms_each_step_rank0": host_enqueue_ms,
            "mem_pool_reserved_mib_before_after_rank0": [pool0, pool1],
            "step_outliers_note": ("one step of this run took > 1.5x the median: a hole in the GPU timeline (the kernels themselves "
                                   "at normal speed, the per-class sums unchanged) that hit one of the first two timed steps in roughly "
                                   "one run out of four before the memory pool's opportunistic reuse was switched off (DESIGN.md 6: a "
                                   "stream-ordered allocation that cannot reuse its block grows the pool at ~0.1 ms per MiB); `value` "
                                   "includes it; compare host_enqueue_ms_each_step_rank0 and mem_pool_reserved_mib_before_after_rank0"
                                   if step_ms and max(step_ms) > 1.5 * sorted(step_ms)[len(step_ms) // 2] else None),
            "srs_fixed_base_tables": (not args.no_precompute) and "window multiples of every SRS level beside the points "
                                     "(csrc/srs.cu), 12.4 GB per party, built at set-up like the SRS itself",
            "value_plain_srs": (1 << n) / (plain_ms * 1e-3) if plain_ms else None,
            "ms_per_step_plain_srs": plain_ms,
            "l2": "one proof streams > 1.2 GB of tables and bases and ~4 GB of MSM temporaries: nothing survives in "
                  "the 126 MB L2 from one step to the next (inputs larger than L2)",
            "msm_per_proof": {"msms": (stats1["segments"] - stats0["segments"]) // args.steps,
                              "launch_sequences": (stats1["sequences"] - stats0["sequences"]) // args.steps,
                              "pairs": pairs // args.steps, "bucket_adds": adds // args.steps},
            "kernel_ms_per_step": {k: v[0] / args.steps for k, v in prof.items()},
            "kernel_ms_note": ("event brackets per kernel class on the stream the class runs on; with SCZ_MSM_STREAM=1 the first MSM "
                               "sequence runs UNDER the rest of the protocol phase, so the sumcheck / open_fold brackets include the "
                               "time their kernels wait for SM slots and the classes add up to more than ms_per_step"
                               if os.environ.get("SCZ_MSM_STREAM") == "1" else "event brackets per kernel class, one stream"),
            "msm_side_stream": os.environ.get("SCZ_MSM_STREAM") == "1",
            "nccl_collectives_per_proof_rank0": coll,
            "net": (net_kind + (": libscz's NCCL hub (grouped ncclSend/ncclRecv + ncclAllGather on the ctx stream, no host "
                                "callback)" if net_kind == "native" else ": torch.distributed callbacks (net.py)")) if world > 1
                   else "leader simulator (serializing_net.rs:144-264)",
            "comm_bytes_per_proof_party0": {"upload": comm[0], "download": comm[1]},
        },
        "clocks": clocks,
        "parity_check": (parity or {}).get("result") if world > 1 else None,
        "parity_check_detail": parity,
        "pipelined": pipelined,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d * parties_total,
                "d2h_bytes_per_step": d2h * parties_total, "steps": e2e_steps, "ms_per_step": e2e_s / e2e_steps * 1e3,
                "single_proof_ms_no_overlap": single_s * 1e3,
                "api": "PackedProvingParameters.upload (pinned host tables -> HBM, on a copy stream, double-buffered so that the "
                       "copy of proof i+1 overlaps proof i) + scz_dhyperplonk_dev + proof -> pinned host buffers on a read-back stream "
                       "(ProofReader: collected one proof behind, status bits with it); wall clock over all steps"},
        "gpu_launches": launches,
        "d_msm": {"metric": "d_msm G1-adds/sec", "unit": "G1 adds/s",
                  "value_all_msm_kernels": adds / (msm_ms * 1e-3) if msm_ms > 0 else None,
                  "value_bucket_kernel": adds / (acc_ms * 1e-3) if acc_ms > 0 else None,
                  "pairs_per_s": pairs / (msm_ms * 1e-3) if msm_ms > 0 else None,
                  "note": "bucket additions of the proof's MSMs (party 0) over the device time of the MSM kernels / of "
                          "k_msm_accumulate alone, inside the timed region"},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": (traffic or {}).get("dram_bytes"), "kernel": acc_kernel, "peak_source": peak_src,
                     "launch_ms": acc_ms / args.steps, "launches": acc_n,
                     "units_per_launch": adds / args.steps, "alg_bytes_per_unit": ALG_BYTES_PER_ADD,
                     "launch_note": "per proof: the bucket accumulation of the proof's MSM sequences (event brackets around the whole "
                                    "accumulation stage of every sequence: with the batched-affine path that is k_ba_levels, 4 x "
                                    "(k_ba_phase1, k_inv_tree_*, k_ba_phase2) and k_ba_accumulate); the sequence of 192 root-opening MSMs "
                                    "of 8 points (< 0.05 ms) is folded into the same figures",
                     "traffic_note": (traffic or {}).get("source"),
                     "note": "achieved = ALGORITHMIC bytes (100 B per bucket addition, SURVEY 8d) / time, as the contract defines it.  The "
                             "accumulation is not one HBM-bound kernel: its upper levels and the XYZZ leftovers are bound by the INT32 multiply "
                             "pipe (ncu: 89 % / 64 %), its level 0 by the DRAM rate of random 96 B gathers (2.6 - 3.9 TB/s of line fills, "
                             "40 - 60 % of the HBM peak); the measured DRAM traffic (`traffic`) is ~7x the algorithmic figure because the "
                             "batched-affine levels trade multiplies for bytes (profiles/r2_ncu_ba_*.txt, DESIGN.md 3.1)"},
    }
    # the other roof of the accumulation: IMAD.WIDE issues once per 4 cycles per SM sub-partition (32 lanes / clk / SM).  An
    # XYZZ mixed addition is 2736 of them (9.5 products of 288); with the batched-affine levels a bucket addition costs
    # 0.88 x 6.1 products (affine, shared inversion) + 0.10 x 9.5 products (XYZZ leftovers) = 1823 on average (4 levels,
    # runs of ~58 entries; 1 / 58 of the entries start a bucket and cost nothing)
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    mhz = (clocks or {}).get("sm_mhz") or 1965
    imad_peak = sms * 32 * mhz * 1e6
    per_unit = 1823 if affine_ran else 2736
    imad_rate = adds * per_unit / (acc_ms * 1e-3) if acc_ms > 0 else 0.0
    line["roofline_int_pipe"] = {"kernel": acc_kernel, "bound": "INT32 multiply pipe (IMAD.WIDE)",
                                 "achieved": imad_rate / 1e12, "peak": imad_peak / 1e12, "unit": "T wide multiplies/s",
                                 "frac": imad_rate / imad_peak, "per_unit": per_unit,
                                 "per_unit_note": "estimated mix for the batched-affine path (see comment in bench.py); 2736 = one XYZZ "
                                                  "mixed addition when the affine levels are off (SCZ_MSM_AFFINE=0)",
                                 "peak_source": f"{sms} SMs x 32 lanes/clk x {mhz} MHz (issue rate measured with ncu: "
                                                "sm__pipe_fmaheavy_cycles_active 86 % for the XYZZ kernel at 2.85 G adds/s, profiles/)"}
    if standalone:
        line.update(standalone)
        if not args.no_cpu and world in (1, 8):
            cpu34 = config34_cpu(n)
            for k in cpu34:
                line[k]["cpu_port"] = cpu34[k]
    if world == 1 and not args.no_standalone and n == 20 and not args.no_precompute:
        line["config2_d_msm"] = config2_d_msm(scz, ctx0, parties[0][1], parties[0][2], torch, peak, imad_peak, not args.no_cpu)
    if world == 1:
        line["roofline_fr_kernels"] = fr_kernel_rooflines(scz, ctx0, torch, peak)
    if world == 1 and not args.no_cpu:
        line["cpu_baseline"] = cpu_baseline_leg(l=L_PACK)
    for ctx, _, _ in parties:
        ctx.close()
    if hub:
        hub.close()
    if world > 1:
        dist.destroy_process_group()
    emit_line(line)


class _StdoutToStderr:
    """Everything but the final JSON line goes to stderr, C libraries included (NCCL prints its version banner to fd 1
    when NCCL_DEBUG=VERSION): fd 1 is pointed at fd 2 for the whole run, `emit` writes to the real stdout."""

    def __init__(self):
        sys.stdout.flush()
        self.real = os.dup(1)
        os.dup2(2, 1)

    def emit(self, text):
        sys.stdout.flush()
        os.write(self.real, (text + "\n").encode())


_OUT = None


def emit_line(line):
    text = json.dumps(line)
    if _OUT is not None:
        _OUT.emit(text)
    else:
        print(text, flush=True)


def main():
    global _OUT
    _OUT = _StdoutToStderr()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--logn", type=int, default=20, help="log2 of the circuit size (BASELINE: 20)")
    ap.add_argument("--pack", dest="l", type=int, default=1, help="packing factor l (N = 8 l parties; BASELINE: 1).  l > 1 is SURVEY 8(f)3: "
                                                     "every party holds 1 / l of each table, 8 l / gpus parties share a GPU")
    ap.add_argument("--ref-logn", type=int, default=0, help="--impl reference: circuit size of the CPU run (0 = --logn: "
                                                            "the same config as the own arm, about a minute per proof)")
    ap.add_argument("--ref-steps", type=int, default=1, help="--impl reference: at most this many timed proofs")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-precompute", action="store_true", help="do not build the fixed-base tables of the SRS")
    ap.add_argument("--no-plain", action="store_true", help="skip the extra leg that times the proof with the tables ignored")
    ap.add_argument("--no-standalone", action="store_true", help="skip the stand-alone legs of BASELINE configs 2, 3, 4")
    ap.add_argument("--no-parity", action="store_true", help="N > 1: skip the untimed parity leg against the oracle")
    ap.add_argument("--parity-logn", type=int, default=10, help="N > 1: circuit size of the parity leg")
    ap.add_argument("--no-pipelined", action="store_true", help="skip the extra N = 1 leg with two provers in flight")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_own(args)


if __name__ == "__main__":
    main()
