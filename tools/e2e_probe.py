"""dev tool: where the end-to-end loop of bench.py loses time against the device-timed loop (one B200, 2^20, leader mode).
modes: dev (back to back, no copies), up (uploads only), rd (pipelined read-back only), both, sync (upload + synchronous read-back)"""
import os
import sys
import time

os.environ.setdefault("SCZ_MSM_STREAM", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import scz_b200 as scz  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
torch.cuda.set_stream(torch.cuda.Stream(priority=-1))
ctx = scz.Context(device=0, n_parties=8)
pp = scz.PackedSharingParams(ctx, 1)
pk = scz.PackedProvingParameters.new(ctx, n, 1, seed=1, shared_seed=0, precompute=True)
alt = scz.PackedProvingParameters(ctx, n, 1, {k: v.clone() for k, v in pk.t.items()}, pk.c_commitment, pk.d_commitment)
host = {name: torch.empty(t.shape, dtype=torch.int64).pin_memory().copy_(t) for name, t in pk.t.items()}
reader = scz.ProofReader(ctx, depth=2)
for _ in range(3):
    scz.dhyperplonk(ctx, n, pk, pp)
torch.cuda.synchronize()


alt2 = scz.PackedProvingParameters(ctx, n, 1, {k: v.clone() for k, v in pk.t.items()}, pk.c_commitment, pk.d_commitment)
half = dict(list(host.items())[: len(host) // 2])


def loop3(steps, prio=0, tabs=host):
    """three table sets: the upload for proof i + 2 waits for proof i - 1 and runs under proof i's MSM phase"""
    sets = [pk, alt, alt2]
    main, copy = torch.cuda.current_stream(), torch.cuda.Stream(priority=prio)
    up = [torch.cuda.Event() for _ in range(3)]
    done = [torch.cuda.Event() for _ in range(3)]
    with torch.cuda.stream(copy):
        for k in range(2):
            sets[k].upload(tabs)
            up[k].record(copy)
    for i in range(steps):
        main.wait_event(up[i % 3])
        if i + 2 < steps:
            with torch.cuda.stream(copy):
                if i >= 1:
                    copy.wait_event(done[(i + 2) % 3])
                sets[(i + 2) % 3].upload(tabs)
                up[(i + 2) % 3].record(copy)
        scz.dhyperplonk(ctx, n, sets[i % 3], pp)
        done[i % 3].record(main)
    torch.cuda.synchronize()


def loop(mode, steps):
    tabs = half if mode == "uphalf" else host
    if mode == "up3":
        return loop3(steps)
    if mode == "up3hi":
        return loop3(steps, prio=-1)
    sets = [pk, alt]
    main, copy = torch.cuda.current_stream(), torch.cuda.Stream(priority=-1 if mode == "uphi" else 0)
    up = [torch.cuda.Event(), torch.cuda.Event()]
    done = [torch.cuda.Event(), torch.cuda.Event()]
    upload = mode in ("up", "both", "sync", "uphi", "uphalf")
    if mode == "upmark":                   # upload of proof i + 1 behind the protocol-phase mark of proof i
        with torch.cuda.stream(copy):
            sets[0].upload(tabs)
            up[0].record(copy)
        for i in range(steps):
            main.wait_event(up[i % 2])
            scz.dhyperplonk(ctx, n, sets[i % 2], pp)
            if i + 1 < steps:
                ctx.stream_wait_protocol_phase(copy)
                with torch.cuda.stream(copy):
                    sets[(i + 1) % 2].upload(tabs)
                    up[(i + 1) % 2].record(copy)
        torch.cuda.synchronize()
        return
    if mode == "alt":                      # alternate the table sets, no copies
        for i in range(steps):
            scz.dhyperplonk(ctx, n, sets[i % 2], pp)
        torch.cuda.synchronize()
        return
    if mode == "d2d":                      # device -> device refresh of the next set on the copy stream (no PCIe)
        for i in range(steps):
            if i + 1 < steps:
                with torch.cuda.stream(copy):
                    if i >= 1:
                        copy.wait_event(done[(i + 1) % 2])
                    for name, t in sets[(i + 1) % 2].t.items():
                        t.copy_(sets[i % 2].t[name], non_blocking=True)
                    up[(i + 1) % 2].record(copy)
            scz.dhyperplonk(ctx, n, sets[i % 2], pp)
            done[i % 2].record(main)
            main.wait_event(up[(i + 1) % 2]) if i + 1 < steps else None
        torch.cuda.synchronize()
        return

    if upload:
        with torch.cuda.stream(copy):
            sets[0].upload(tabs)
            up[0].record(copy)
    pending = None
    for i in range(steps):
        if upload:
            main.wait_event(up[i % 2])
            if i + 1 < steps:
                with torch.cuda.stream(copy):
                    if i >= 1:
                        copy.wait_event(done[(i + 1) % 2])
                    sets[(i + 1) % 2].upload(tabs)
                    up[(i + 1) % 2].record(copy)
        proof = scz.dhyperplonk(ctx, n, sets[i % 2], pp)
        done[i % 2].record(main)
        if mode in ("rd", "both"):
            t = proof.to_host_async(reader)
            if pending is not None:
                reader.collect(pending)
            pending = t
        elif mode == "sync":
            proof.to_host()
    if pending is not None:
        reader.collect(pending)
    torch.cuda.synchronize()


for mode in ("dev", "up", "upmark", "dev", "upmark"):
    loop(mode, 2)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    loop(mode, steps)
    dt = time.perf_counter() - t0
    print(f"{mode:5s} {dt / steps * 1e3:8.2f} ms per proof (wall, {steps} proofs)", flush=True)
