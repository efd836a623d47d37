#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 hot path (see BASELINE.json / SURVEY.md 8d).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload NAME] [--batch B]

A "step" is one pass of the hot path -- CKKS Evaluator::multiply + relinearize_inplace -- over one batch of synthetic
uniform-random ciphertexts (as native/bench does: bench.h:195-270) that is already resident in HBM.  Under torchrun
every rank processes its own shard of the batch on its own GPU (weak scaling, no data-path collective; a tiny NCCL
gather of per-ciphertext digests at the end).  Rank 0 prints ONE JSON line.

  value     whole-job ciphertexts/s, inputs resident in HBM, CUDA-event timed, max over ranks
  e2e       same metric through the host-buffer C-ABI call (pinned host buffers, H2D + D2H inside the timed region)
  roofline  dominant kernel: algorithmic bytes / CUDA-event duration measured live in the timed region (sb200_profile_*)
  cpu_baseline  the reference's own CPU implementation (oracle/_ref, i.e. SEAL 4.4.3 compiled from its sources, HEXL off)
                on this box's host cores, bounded sample

--impl reference times that CPU implementation alone, on the same workload/metric (all host threads).
"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

WORKLOADS = {
    # BASELINE.json configs[4] (the config the metric is quoted on): CKKS n=65536, 32 primes, multiply+relinearize;
    # 8192 ciphertexts over 8 GPUs = 1024 per GPU.  Fits one GPU: 2 x 33.3 GB in + 33.3 GB out + scratch.
    "ckks_n65536_k32": dict(scheme=2, n=65536, bits=[55] * 32, batch=1024, e2e_batch=64, cpu_reps=1),
    # BASELINE.json configs[1]
    "ckks_n8192_k4": dict(scheme=2, n=8192, bits=[54] * 4, batch=1024, e2e_batch=256, cpu_reps=200),
    # metric text "n=2^16, L=16 primes"
    "ckks_n65536_k16": dict(scheme=2, n=65536, bits=[55] * 16, batch=1024, e2e_batch=32, cpu_reps=2),
    "ckks_n32768_k16": dict(scheme=2, n=32768, bits=[55] * 15 + [56], batch=256, e2e_batch=64, cpu_reps=4),
    # tiny shape for the CPU contract test of the reference arm (tests/test_bench_contract.py); not a bench line
    "smoke": dict(scheme=2, n=4096, bits=[40, 40, 40], batch=8, e2e_batch=4, cpu_reps=4),
}
METRIC = "CKKS multiply+relinearize ciphertexts/s"
UNIT = "ciphertexts/s"


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(kernel):
    """dram bytes per launch of the dominant kernel from the committed ncu capture summary, or None"""
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
            e = json.load(f).get(kernel)
            return e["dram_bytes_per_launch"] if e else None
    except Exception:
        return None


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.p = None

    def stop(self):
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            out, _ = self.p.communicate(timeout=5)
        except Exception:
            self.p.kill()
            out = ""
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        # median of the upper half = clocks under load (idle samples before/after the region pull the plain median down)
        load = sm[len(sm) // 2:] if sm else []
        return {"sm_mhz": (load[len(load) // 2] if load else None), "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def gpu_numa_cpus(dev):
    """(node, cpus) of the NUMA node the GPU's PCIe root port hangs off, or (None, None) when the box does not say"""
    try:
        import torch

        p = torch.cuda.get_device_properties(dev)
        addr = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{addr}/numa_node").read())
        if node < 0:
            return None, None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        return (node, cpus) if cpus else (None, None)
    except Exception:
        return None, None


def run_reference(args, wl, rank, world):
    """--impl reference: the reference's own CPU implementation (oracle/_ref) on the host cores."""
    if rank != 0:
        return
    import refseal as R

    cores = os.cpu_count() or 1
    if not R.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libsealref.so missing (build it with oracle/Makefile)"}))
        return
    mods = R.coeff_modulus_create(wl["n"], wl["bits"])
    L = len(mods) - 1
    t0 = time.time()
    rc = R.RefContext(wl["scheme"], wl["n"], mods)
    rc.relin_key()
    setup_s = time.time() - t0
    reps = wl["cpu_reps"]
    for _ in range(args.warmup if args.warmup < 2 else 1):
        rc.time_op(0, L, cores, 1)
    total_t, total_ops = 0.0, 0
    for _ in range(args.steps):
        total_t += rc.time_op(0, L, cores, reps)
        total_ops += cores * reps
    v = total_ops / total_t
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total_t / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64",
        "data": "synthetic", "config": {"workload": args.workload, "n": wl["n"], "coeff_modulus_primes": len(mods), "L": L,
                                        "ops": "Evaluator::multiply_inplace + relinearize_inplace", "step": f"{cores} threads x {reps} ciphertexts"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "reference",
                         "sample": f"{total_ops} ciphertexts on {cores} threads (SEAL 4.4.3 from oracle/_ref, HEXL off); setup {setup_s:.1f}s excluded"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="ckks_n65536_k32", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="ciphertexts per GPU per step (default: the workload's)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    wl = dict(WORKLOADS[args.workload])
    if args.batch:
        wl["batch"] = args.batch
        wl["e2e_batch"] = min(wl["e2e_batch"], args.batch)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, wl, rank, world)
        return

    import numpy as np
    import torch
    import torch.distributed as dist

    import seal_b200 as S

    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n, B = wl["n"], wl["batch"]
    mods = S.coeff_modulus_create(n, wl["bits"])
    k, L = len(mods), len(mods) - 1
    ctx = S.Context(wl["scheme"], n, mods, device=local)
    g = torch.Generator(device="cuda")
    g.manual_seed(0x5EA1 + rank)

    def rand_rows(shape_prefix, nprimes):
        """uniform residues in [0, q_i) per RNS row, generated on the device: [*shape_prefix][nprimes][n]"""
        t = torch.empty((*shape_prefix, nprimes, n), dtype=torch.int64, device="cuda")
        for i in range(nprimes):
            t[..., i, :] = torch.randint(0, mods[i], (*shape_prefix, n), generator=g, dtype=torch.int64, device="cuda")
        return t

    # synthetic key: uniform residues at the key level ([L digits][2][k][n]); key VALUES do not affect the work done
    key_host = rand_rows((L, 2), k).cpu().numpy().view(np.uint64)
    rk = ctx.load_key(key_host)
    a, b = rand_rows((B, 2), L), rand_rows((B, 2), L)
    out = torch.empty((B, 2, L, n), dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()

    def step():
        ctx.d_multiply_relinearize(a, b, rk, out, L, B)

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    launches0 = ctx.launch_count
    ctx.profile(True)
    sampler = ClockSampler(local) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    clocks = sampler.stop() if sampler else None
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    launches = ctx.launch_count - launches0
    prof = ctx.profile_read()
    ctx.profile(False)

    # end-of-run gather (the only collective): one 64-bit digest per ciphertext of this rank's last step
    from seal_b200.shard import digest as ct_digest, gather_digests

    digests = gather_digests(ct_digest(out), rank, world)  # NCCL over NVLink when world > 1
    torch.cuda.synchronize()
    assert rank != 0 or digests.numel() == B * world

    # ---- second headline metric: NTT GB/s vs the HBM roofline (Evaluator::transform_from_ntt_inplace + transform_to_ntt_inplace
    #      over the whole input slab, the reference bench's NTTForward/NTTInverse cases; 2*n*8 algorithmic bytes per row per transform)
    ntt_rows = B * 2 * L
    nb = min(B, 64)
    ntt_rows = nb * 2 * L
    ctx.d_ntt_inverse(a, L, 2, nb)
    ctx.d_ntt_forward(a, L, 2, nb)
    torch.cuda.synchronize()
    n0, n1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n0.record()
    for _ in range(args.steps):
        ctx.d_ntt_inverse(a, L, 2, nb)
        ctx.d_ntt_forward(a, L, 2, nb)  # restores the slab exactly (checked by the parity tests)
    n1.record()
    torch.cuda.synchronize()
    ntt_ms = n0.elapsed_time(n1) / (2 * args.steps)
    ntt_gbps = ntt_rows * 2 * n * 8 / (ntt_ms * 1e-3) / 1e9

    # ---- e2e through the host-buffer C-ABI entry point (pinned host memory, H2D + D2H inside the timed region)
    e2e = None
    if not args.no_e2e:
        Be = wl["e2e_batch"]
        # staging buffers on the GPU's own NUMA node (first touch by a thread bound there): every rank's H2D / D2H traffic then
        # stays on its socket instead of crossing the inter-socket link; the affinity is restored before the CPU baseline runs
        all_cpus = os.sched_getaffinity(0)
        numa_node, numa_cpus = gpu_numa_cpus(local)
        if numa_cpus:
            os.sched_setaffinity(0, numa_cpus)
        ha = torch.empty((Be, 2, L, n), dtype=torch.int64).pin_memory()
        hb = torch.empty((Be, 2, L, n), dtype=torch.int64).pin_memory()
        ho = torch.empty((Be, 2, L, n), dtype=torch.int64).pin_memory()
        ha.copy_(a[:Be])
        hb.copy_(b[:Be])
        torch.cuda.synchronize()
        na, nb_, no = ha.numpy().view(np.uint64), hb.numpy().view(np.uint64), ho.numpy().view(np.uint64)

        def e2e_step():
            S._check(S.lib().sb200_multiply_relinearize_host(ctx.h, L, Be, S._hp(na), S._hp(nb_), rk.h, S._hp(no)))

        e2e_step()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_step()  # returns after the D2H copy has completed
        t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        assert (ho.cuda() == out[:Be]).all(), "e2e result differs from the device-resident result"
        os.sched_setaffinity(0, all_cpus)
        e2e = {"value": world * Be * args.steps / float(t.item()), "unit": UNIT, "h2d_bytes_per_step": int(ha.nbytes + hb.nbytes),
               "d2h_bytes_per_step": int(ho.nbytes), "batch_per_step": Be, "host_numa_node": numa_node}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    value = world * B * args.steps / (ms_total / 1e3)
    peak, peak_src = measured_peak()
    # dominant kernel by device time; achieved = algorithmic bytes / duration
    prof.sort(key=lambda r: -r[1])
    tot_prof_ms = sum(r[1] for r in prof) or 1.0
    top = prof[0]
    achieved = top[3] / (top[1] * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": top[0], "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": ncu_traffic(top[0]), "peak_source": peak_src, "share_of_step": top[1] / tot_prof_ms,
                "launches": top[2], "avg_launch_ms": top[1] / max(top[2], 1),
                "note": "integer-ALU bound kernel (uint64 Harvey butterflies); HBM roofline reported as the contract asks, see DESIGN.md",
                "kernels": [{"name": r[0], "ms": round(r[1], 3), "launches": r[2], "alg_GBps": round(r[3] / (r[1] * 1e-3) / 1e9, 1) if r[1] > 0 else None}
                            for r in prof]}
    # whole-op roofline: fused multiply+relinearize moves 6*L*n*8 bytes per ciphertext + one key pass per chunk
    alg_per_ct = 6 * L * n * 8
    op_gbps = value / world * alg_per_ct / 1e9
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        import refseal as R

        cores = os.cpu_count() or 1
        if R.available():
            rc = R.RefContext(wl["scheme"], n, mods)
            rc.relin_key()
            reps = wl["cpu_reps"]
            rc.time_op(0, L, cores, 1)
            t = rc.time_op(0, L, cores, reps)
            cpu = {"value": cores * reps / t, "unit": UNIT, "cores": cores, "kind": "reference",
                   "sample": f"{cores * reps} ciphertexts, {cores} threads x {reps}, SEAL 4.4.3 built from its own sources (oracle/_ref), HEXL off"}
        else:
            cpu = {"value": None, "unit": UNIT, "cores": cores, "kind": "reference", "sample": "oracle/_ref/libsealref.so not present"}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64",
        "data": "synthetic",
        "config": {"workload": args.workload, "scheme": "CKKS", "n": n, "coeff_modulus_primes": k, "L": L, "prime_bits": wl["bits"][0],
                   "batch_per_gpu": B, "global_batch": B * world, "ops": "Evaluator::multiply + relinearize_inplace (fused call)",
                   "l2": f"inputs {2 * a.numel() * 8 / 2**30:.1f} GiB per GPU >> 126 MB L2 (no flush needed)",
                   "parallelism": f"batch sharded x{world}, no data-path collective"},
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
        "ntt": {"metric": "negacyclic NTT GB/s (2*n*8 B per row per transform)", "rows_per_transform_call": ntt_rows, "n": n,
                "ms_per_call": ntt_ms, "achieved_GBps": ntt_gbps, "peak_GBps": peak, "frac_of_hbm_peak": ntt_gbps / peak},
        "op_roofline": {"alg_bytes_per_ct": alg_per_ct, "achieved_GBps_per_gpu": op_gbps, "frac_of_hbm_peak": op_gbps / peak},
        "cpu_baseline": cpu,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
