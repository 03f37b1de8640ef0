#!/usr/bin/env python
"""bench.py -- IBL-NeRF kitchen-config training step throughput (rays/s) on N B200s.

    python bench.py --gpus N --steps K --warmup W            # B200-native path (this repo)
    python bench.py --impl reference --gpus N ...            # reference algorithm on the host CPU cores (oracle port)

step = one training iteration of src/train.py for the kitchen config in the full-IBL phase (BASELINE.json configs[1]):
render_decomp (64 coarse + 128 fine samples, epsilon normals, reflected ray, split-sum shading) -> phase-B losses ->
backward -> (NCCL gradient all-reduce) -> Adam, on N_rand = 4096 synthetic rays per GPU with random-init weights.
Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

METRIC = "train_rays_per_sec"
UNIT = "rays/s"
WORKLOAD = ("IBL-NeRF kitchen full-IBL training step (render_decomp + phase-B loss + backward + Adam), "
            "N_rand=%d rays/GPU, 64 coarse + 128 fine samples")
N_RAND = 4096
FLOP_FULL, FLOP_SIGMA = 1591552, 982528          # algorithmic FLOP / point (unpadded)
FLOP_PER_RAY_STEP = 2432139264                   # full-IBL training step


def synth_rays(n, seed, device="cpu", pin=False):
    g = torch.Generator().manual_seed(seed)
    o = torch.rand(n, 3, generator=g) * 2 - 1
    d = torch.randn(n, 3, generator=g)
    d = d / d.norm(dim=-1, keepdim=True) * (1.0 + 0.3 * torch.rand(n, 1, generator=g))
    tg = {k: torch.rand(n, 3, generator=g) for k in ("rgb", "rgb_1", "rgb_2", "rgb_3")}
    if pin:
        o, d = o.pin_memory(), d.pin_memory()
        tg = {k: v.pin_memory() for k, v in tg.items()}
    return o.to(device), d.to(device), {k: v.to(device) for k, v in tg.items()}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 20 ms DURING the timed regions (resident + end-to-end)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                       "-lms", "20"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(", ") for r in open(self.f.name).read().strip().splitlines() if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], None, set()
        for r in rows:
            try:
                sm.append(float(r[0])); mx = float(r[1])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.strip().lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        busy = [x for x in sm if mx and x > 0.3 * mx] or sm
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": mx, "samples": len(sm), "reasons": sorted(reasons)}


def cpu_reference_step(n_rays, threads):
    """One fwd+bwd of the reference algorithm (oracle port, plain torch CPU fp32) on n_rays rays; returns seconds."""
    import fixtures as fx
    from oracle import iblnerf_oracle as orc
    torch.set_num_threads(threads)
    torch.manual_seed(0)
    nets = []
    for _ in range(2):
        p = {}
        for name, o, i in orc.PARAM_SHAPES_INIT_ORDER:
            lin = torch.nn.Linear(i, o)
            p[name + ".weight"], p[name + ".bias"] = lin.weight, lin.bias
        nets.append(p)
    o, d, tg = synth_rays(n_rays, 1)
    rays = torch.cat([o, d, torch.full((n_rays, 1), 0.5), torch.full((n_rays, 1), 8.0), d / d.norm(dim=-1, keepdim=True)], -1)
    lut = fx.load_lut()
    t0 = time.perf_counter()
    res = orc.render_rays(rays, nets[0], nets[1], lut, perturb=1.0, approximate_radiance=True)
    loss = fx.phase_b_loss(res, tg)
    loss.backward()
    return time.perf_counter() - t0


def eager_cuda_reference(n_rays, steps=2):
    """The same reference algorithm (oracle port) in PyTorch EAGER mode on cuda:0 -- what a user of the reference gets
    on this GPU without this package ("the oracle in eager CUDA fp32 as the reference-on-B200
    bar").  A reported baseline like cpu_baseline: fp32 matmuls (parity setting) and TF32 matmuls (the setting of the
    reference authors' Ampere GPUs, torch 1.11 default).  Returns {"fp32": rays/s, "tf32": rays/s}."""
    import fixtures as fx
    from oracle import iblnerf_oracle as orc
    dev = torch.device("cuda:0")
    out = {}
    o, d, tg = synth_rays(n_rays, 1, dev)
    lut = fx.load_lut().to(dev)
    prev = torch.backends.cuda.matmul.allow_tf32
    try:
        with torch.device(dev):         # the oracle (like the reference) allocates with bare factory calls
            torch.manual_seed(0)
            nets = []
            for _ in range(2):
                p = {}
                for name, oo, ii in orc.PARAM_SHAPES_INIT_ORDER:
                    lin = torch.nn.Linear(ii, oo)
                    p[name + ".weight"], p[name + ".bias"] = lin.weight, lin.bias
                nets.append(p)
            rays = torch.cat([o, d, torch.full((n_rays, 1), 0.5), torch.full((n_rays, 1), 8.0), d / d.norm(dim=-1, keepdim=True)], -1)
            for tag, tf32 in (("fp32", False), ("tf32", True)):
                torch.backends.cuda.matmul.allow_tf32 = tf32
                ts = []
                for it in range(steps + 1):
                    for p in nets:
                        for v in p.values():
                            v.grad = None
                    torch.cuda.synchronize(dev)
                    t0 = time.perf_counter()
                    res = orc.render_rays(rays, nets[0], nets[1], lut, perturb=1.0, approximate_radiance=True)
                    loss = fx.phase_b_loss(res, tg)
                    loss.backward()
                    float(loss)
                    torch.cuda.synchronize(dev)
                    if it > 0:
                        ts.append(time.perf_counter() - t0)
                    del res, loss
                out[tag] = n_rays / (sum(ts) / len(ts))
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev
    return out


def run_reference(args, rank, world):
    """--impl reference: the reference's own algorithm on the box's host cores (rank 0 only)."""
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    sample = 1024                       # BASELINE.json configs[0]: N_rand = 1024 on the CPU reference path
    for _ in range(max(1, min(args.warmup, 1))):
        cpu_reference_step(256, threads)
    steps = max(1, min(args.steps, 3))
    ts = [cpu_reference_step(sample, threads) for _ in range(steps)]
    sec = sum(ts) / len(ts)
    val = sample / sec
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": world, "steps": steps,
            "warmup": 1, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic rays, random-init weights",
            # the product arm's workload (same string); each reference step is a bounded sample of it
            "config": {"workload": WORKLOAD % args.n_rand, "n_rand_per_gpu": args.n_rand,
                       "sample": "each step = fwd+bwd of %d of the %d rays (the reference algorithm scales linearly in rays)" % (sample, args.n_rand)},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": "%d rays x %d steps, oracle/iblnerf_oracle.py (torch CPU fp32)" % (sample, steps)},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n-rand", type=int, default=N_RAND)
    ap.add_argument("--precision", default=None, choices=[None, "bf16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-eager-baseline", action="store_true", help="skip the PyTorch-eager-on-GPU run of the oracle (N=1 only)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import fixtures as fx
    import ibl_nerf_b200 as ib
    from ibl_nerf_b200 import _lib, training

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the B200-native path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"       # the version banner goes to stdout, which carries the ONE JSON line
        dist.init_process_group("nccl", device_id=dev)
    _lib.lib()
    n = args.n_rand
    lut = fx.load_lut().to(dev)
    ts = training.TrainStep(dev, lut, precision=args.precision)
    o_d, d_d, tg_d = synth_rays(n, 100 + rank, dev)
    o_h, d_h, tg_h = synth_rays(n, 100 + rank, "cpu", pin=True)
    h2d = sum(t.numel() * 4 for t in (o_h, d_h, *tg_h.values()))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    def step_resident():
        ts.step(o_d, d_d, tg_d)

    # End-to-end step through the public API: this step's rays / targets come from pinned host memory, and every
    # step's loss is read back to the host.  The read is software-pipelined by one step (the loss of step k is copied
    # into pinned memory on the stream and consumed while step k+1 is being enqueued), the way an asynchronous
    # training logger does it, so the host never drains the GPU queue; the last loss is flushed inside the timed region.
    loss_host = torch.zeros(2, dtype=torch.float32).pin_memory()
    pending = []

    def flush_loss():
        while pending:
            ev, slot = pending.pop(0)
            ev.synchronize()
            float(loss_host[slot])

    def step_e2e():
        o = o_h.to(dev, non_blocking=True)
        d = d_h.to(dev, non_blocking=True)
        tg = {k: v.to(dev, non_blocking=True) for k, v in tg_h.items()}
        loss = ts.step(o, d, tg)
        slot = step_e2e.count & 1
        step_e2e.count += 1
        loss_host[slot:slot + 1].copy_(loss.reshape(1), non_blocking=True)      # D2H read of the loss
        ev = torch.cuda.Event()
        ev.record()
        if len(pending) >= 1:
            ev0, slot0 = pending.pop(0)
            ev0.synchronize()
            float(loss_host[slot0])
        pending.append((ev, slot))
    step_e2e.count = 0

    for _ in range(args.warmup):
        step_resident()
    # --- timed region: K steps, inputs resident in HBM.  Per-step working set (activations of 5.8 M point
    # evaluations) is far larger than the 126 MB L2, so no explicit flush is needed between iterations.
    # inside the timed region only the dominant kernel's launches are bracketed by CUDA events (roofline); every
    # entry point still counts its launches
    _lib.PROFILE = {}
    _lib.PROFILE_EVENTS = {"ibln_mlp_fwd"}
    sampler = ClockSampler(local) if rank == 0 else None
    ms_total = timed(step_resident, args.steps)
    prof = _lib.PROFILE
    _lib.PROFILE = None
    _lib.PROFILE_EVENTS = None
    ms_step = ms_total / args.steps
    value = world * n / (ms_step * 1e-3)
    # --- same metric end to end (host pinned inputs -> H2D every step, loss read back)
    for _ in range(2):
        step_e2e()
    flush_loss()

    def e2e_steps_then_flush():
        step_e2e()
        if step_e2e.count == e2e_steps_then_flush.last:
            flush_loss()
    e2e_steps_then_flush.last = step_e2e.count + args.steps
    ms_e2e = timed(e2e_steps_then_flush, args.steps) / args.steps
    e2e_val = world * n / (ms_e2e * 1e-3)
    clocks = sampler.stop() if sampler else None
    # per-entry time breakdown: two extra, untimed steps with every entry point bracketed
    _lib.PROFILE = {}
    for _ in range(2):
        step_resident()
    torch.cuda.synchronize()
    prof_all = _lib.PROFILE
    _lib.PROFILE = None

    if rank == 0:
        torch.cuda.synchronize()
        launches = sum(v["launches"] for v in prof.values())
        mlp = prof.get("ibln_mlp_fwd")
        roof = None
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        if mlp and mlp["events"]:
            dur_ms = sum(a.elapsed_time(b) for a, b in mlp["events"])
            peak = peaks.get("bf16_tflops_sustained", 1400.0)
            ach = mlp["flops"] / (dur_ms * 1e-3) / 1e12
            roof = {"bound": "tensor", "kernel": "mlp_fwd_kernel (fused encode + MLP, tcgen05)", "achieved": ach, "peak": peak,
                    "unit": "TFLOP/s", "frac": ach / peak,
                    # dram__bytes_read+write per launch, mean over the 6 launches of one step (profiles/r1_final_ncu_full.md);
                    # 99.5 % of it is the activation stash written by the two gradient launches
                    "traffic": 1.228e9, "traffic_unit": "B/launch",
                    "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained (of measured)" if peaks else "fallback 1.4 PFLOP/s sustained (of fallback)",
                    "launches": len(mlp["events"]), "share_of_step": dur_ms / ms_total}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "bf16" if (args.precision or ib.mlp.default_precision()) == "bf16" else "f32",
                "data": "synthetic rays (seeded), random-init weights of the kitchen architecture",
                "config": {"workload": WORKLOAD % n,
                           "n_rand_per_gpu": n, "parallelism": "ray-sharded dp%d, NCCL grad all-reduce" % world,
                           "l2": "per-step working set >> 126 MB L2 (no explicit flush)",
                           "step_tflops": world * n * FLOP_PER_RAY_STEP / (ms_step * 1e-3) / 1e12},
                "clocks": clocks,
                "e2e": {"value": e2e_val, "unit": UNIT, "ms_per_step": ms_e2e, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4},
                "gpu_launches": launches,
                "kernel_launches_by_entry": {k: v["launches"] for k, v in sorted(prof.items())},
                "ms_per_step_by_entry": {k: round(sum(a.elapsed_time(b) for a, b in v["events"]) / 2, 4)
                                         for k, v in sorted(prof_all.items())},
                "roofline": roof}
        if mlp and mlp["events"] and len(mlp["events"]) % args.steps == 0:
            per = len(mlp["events"]) // args.steps
            line["mlp_fwd_launch_ms"] = [round(sum(mlp["events"][i * per + j][0].elapsed_time(mlp["events"][i * per + j][1])
                                                   for i in range(args.steps)) / args.steps, 4) for j in range(per)]
        if not args.no_cpu_baseline and world == 1:     # reported at N=1 only
            threads = os.cpu_count() or 1
            sample = 1024
            cpu_reference_step(64, threads)
            secs = [cpu_reference_step(sample, threads) for _ in range(2)]
            sec = sum(secs) / len(secs)
            line["cpu_baseline"] = {"value": sample / sec, "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": "2 steps of %d rays (fwd+bwd, %.1f s), oracle/iblnerf_oracle.py, torch CPU fp32" % (sample, sum(secs))}
        if not args.no_eager_baseline and world == 1:
            try:
                eg = eager_cuda_reference(n)
                line["eager_cuda_baseline"] = {"value": eg["fp32"], "value_tf32": eg["tf32"], "unit": UNIT, "kind": "port",
                                               "sample": "2 steps of %d rays (fwd+bwd) after 1 warm-up, oracle/iblnerf_oracle.py in "
                                                         "PyTorch eager mode on cuda:0 (fp32 and TF32 matmuls)" % n}
            except Exception as e:      # a reported baseline must never take the bench line down
                line["eager_cuda_baseline"] = {"error": "%s: %s" % (type(e).__name__, str(e)[:200])}
            torch.cuda.empty_cache()
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
