#!/usr/bin/env python
"""bench.py -- headline metric of BASELINE.json: collocation points/s for one residual+grad step.

Workload (configs[1], the configuration the metric is quoted on): Burgers nu=0.01/pi, feedforward tanh
8x128, 1M collocation points per B200 (weak scaling: every rank owns 1M rows).  One step =
PDETrainer's inner step (trainer.py:577-578,689-694): zero_grad -> compute_loss (residual + 200
boundary + 100 initial rows) -> backward -> [all-reduce of the flat gradient when N>1] -> clip -> Adam.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--points P] [--no-configs]

`value`   device-timed steps with the collocation rows resident in HBM;
`e2e`     the same step through the public API from pinned HOST buffers (H2D of the rows and D2H of the
          loss inside the timed region, every step; pipelined like a loader: the next step's rows are copied on a side
          stream while this step computes, and this step's loss is read on the host during the next step);
`roofline`     dominant kernel class timed with CUDA event pairs inside libpinnk (profiling pass); the whole step against
               BOTH tensor denominators (TF32 measured in this run / 3, and MEASURED_PEAKS bf16_sustained / 6);
`cpu_baseline` the UNMODIFIED reference (pinnrl, installed to baseline/_ref; kind "reference") on the host cores -- the
               oracle port (kind "port", bit-identical arithmetic) only where baseline/_ref did not travel;
`configs`      the other BASELINE configs (C1 at 4 900 and 1 M points, C3, C4 as written and math, C5 scoring) as secondary
               entries: points/s, algorithmic FLOP/point and both roofline fractions.  With --gpus 8 they run at BASELINE's
               GLOBAL sizes (C3 4 M, C4 16 M, C5 64 M points over the 8 GPUs: strong-scaling entries);
`per_rank_step_ms`, `collective_ms`   (N > 1) min / median / max step time over the ranks and the time inside the step's one
               all-reduce, so that a scaling curve names its limiter.
`--impl reference` times the reference CPU path alone, on a bounded sample of the workload (65 536 rows).
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "collocation points/sec (residual+grad step)"
HIDDEN, LAYERS, NU = 128, 8, 0.01 / math.pi
W_ELEMS = 2 * 128 + 7 * 128 * 128 + 128          # Linear weight elements (115 072)
JET_COLS = 4                                      # u, u_x, u_xx, u_t
FLOPS_PER_POINT_STEP = 6 * JET_COLS * W_ELEMS     # fwd 2CW + dgrad 2CW + wgrad 2CW = 2.76 MFLOP
CPU_SAMPLE_POINTS = 65536                         # N_cpu for C2

# The other BASELINE configs: jet columns C, Linear weight elements W, step = 6CW (scoring 2CW) FLOP/point.
# n1: points per GPU in the default (weak) runs; n8: BASELINE's GLOBAL size, used with --gpus 8 (strong-scaling entry).
CONFIGS = [
    dict(key="c1_heat_fourier_4900", cfg="configs[0]", pde="heat", arch="fourier", hidden=128, layers=4, dim=1,
         extra={"mapping_size": 32, "scale": 10.0}, compat="reference", mode="loss", n1=4900, n8=None, C=3, W=41152,
         training={"num_collocation_points": 5000, "num_boundary_points": 500, "num_initial_points": 500}),      # README.md:99-141
    dict(key="c1_heat_fourier_1m", cfg="configs[0] network at 1 M points", pde="heat", arch="fourier", hidden=128, layers=4, dim=1,
         extra={"mapping_size": 32, "scale": 10.0}, compat="reference", mode="loss", n1=1 << 20, n8=None, C=3, W=41152,
         training={"num_collocation_points": 5000, "num_boundary_points": 500, "num_initial_points": 500}),      # (same README row counts)
    dict(key="c3_kdv_resnet_6x256", cfg="configs[2]", pde="kdv", arch="resnet", hidden=256, layers=6, dim=1,
         extra={"num_blocks": 6}, compat="reference", mode="loss", n1=1 << 19, n8=1 << 22, C=5, W=787200),
    dict(key="c4_cahn_hilliard_2d_siren_5x256_as_written", cfg="configs[3], operator as the reference evaluates it (u_t, SURVEY F2)",
         pde="cahn_hilliard", arch="siren", hidden=256, layers=5, dim=2, extra={"omega_0": 30.0}, compat="reference", mode="mse",
         n1=1 << 21, n8=1 << 24, C=2, W=263168),
    dict(key="c4_cahn_hilliard_2d_siren_5x256_math", cfg="configs[3], intended 4th-order operator (18 jet columns)",
         pde="cahn_hilliard", arch="siren", hidden=256, layers=5, dim=2, extra={"omega_0": 30.0}, compat="math", mode="mse",
         n1=1 << 20, n8=1 << 24, C=18, W=263168),
    dict(key="c5_allen_cahn_scoring", cfg="configs[4]", pde="allen_cahn", arch="feedforward", hidden=128, layers=8, dim=1,
         extra={}, compat="reference", mode="score", n1=1 << 23, n8=1 << 26, C=4, W=115072),
    # residual-adaptive refinement over BASELINE configs[4]'s 64 M-candidate pool (pde_base.py:895-935: score, then draw pool / 4
    # points with probability |r| + 1e-8): forward-only scoring + the on-device inverse-CDF sampler (the pool is past
    # torch.multinomial's 2^24 limit).  The global pool is 64 M at every GPU count (strong scaling).
    dict(key="c5_rar_draw_64m", cfg="configs[4] pool, scored and sampled (RAR)", pde="allen_cahn", arch="feedforward", hidden=128,
         layers=8, dim=1, extra={}, compat="reference", mode="rar", n1=1 << 26, n8=1 << 26, C=4, W=115072, fixed_global=True),
]
PDE_SPECS = {
    "heat": dict(domain=[[0.0, 1.0]], time=[0.0, 1.0], params={"alpha": 0.01}, bcs={"dirichlet": {"type": "dirichlet"}},
                 ic={"type": "sine", "amplitude": 1.0, "frequency": 2.0}, exact={"type": "sin_exp_decay", "amplitude": 1.0, "frequency": 2.0}),
    "kdv": dict(domain=[[-15.0, 15.0]], time=[0.0, 5.0], params={"speed": 1.0}, bcs={"dirichlet": {"value": 0.0}},
                ic={"type": "soliton", "speed": 1.0}, exact={}),
    "cahn_hilliard": dict(domain=[[0.0, 1.0]], time=[0.0, 1.0], params={"epsilon": 0.1}, bcs={"dirichlet": {"value": 0.0}},
                          ic={"type": "tanh", "epsilon": 0.1}, exact={}),
    "allen_cahn": dict(domain=[[-1.0, 1.0]], time=[0.0, 1.0], params={"epsilon": 0.1}, bcs={"dirichlet": {"value": 0.0}},
                       ic={"type": "tanh", "epsilon": 0.1}, exact={}),
}


def burgers_cfg(pk, dev):
    return pk.PDEConfig(name="burgers", domain=[[-1.0, 1.0]], time_domain=[0.0, 1.0], parameters={"nu": NU},
                        boundary_conditions={"dirichlet": {"value": 0.0}},
                        initial_condition={"type": "sine", "amplitude": -1.0, "frequency": 1.0},
                        exact_solution={}, dimension=1, device=dev)


def synth_points(n, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(n, 1, generator=g) * 2 - 1, torch.rand(n, 1, generator=g)


# --------------------------------------------------------------------------- CPU reference arm
def cpu_reference_step_fn(n_points):
    """One trainer step of the reference on the host cores, fp32: (step function, kind).
    kind "reference": the UNMODIFIED pinnrl package from baseline/_ref (its PINNModel, BurgersEquation.compute_loss,
    backward, clip, Adam -- trainer.py:577-578,689-694); kind "port": oracle/ref_port.py (bit-identical arithmetic) when
    baseline/_ref is not there."""
    x, t = synth_points(n_points, 1)
    try:
        from oracle import ref_env
        ref_env.activate()
        from pinnrl.config import Config, ModelConfig
        from pinnrl.neural_networks import PINNModel
        from pinnrl.pdes.burgers_equation import BurgersEquation
        from pinnrl.pdes.pde_base import PDEConfig
        cpu = torch.device("cpu")
        c = Config.__new__(Config)
        c.device = cpu
        c.model = ModelConfig(2, HIDDEN, 1, LAYERS, "tanh", architecture="feedforward")
        torch.manual_seed(0)
        import contextlib
        import io
        with contextlib.redirect_stdout(io.StringIO()):
            model = PINNModel(config=c, device=cpu)
            pde = BurgersEquation(config=PDEConfig(
                name="burgers", domain=[[-1.0, 1.0]], time_domain=[0.0, 1.0], parameters={"nu": NU},
                boundary_conditions={"dirichlet": {"value": 0.0}},
                initial_condition={"type": "sine", "amplitude": -1.0, "frequency": 1.0}, exact_solution={}, dimension=1, device=cpu))
        opt = torch.optim.Adam(model.parameters(), lr=1e-3)

        def step():
            opt.zero_grad()
            losses = pde.compute_loss(model, x, t)
            losses["total"].backward()
            torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
            opt.step()
            return float(losses["total"].detach())
        return step, "reference"
    except ImportError:
        pass
    from oracle import ref_port
    torch.manual_seed(0)
    model = ref_port.PINNModel("feedforward", 2, HIDDEN, LAYERS)
    fns = ref_port.boundary_condition_fns("burgers", {"dirichlet": {"value": 0.0}},
                                          {"type": "sine", "amplitude": -1.0, "frequency": 1.0}, [(-1.0, 1.0)], {"nu": NU})
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)

    def step():
        opt.zero_grad()
        r = ref_port.burgers_residual(model, x, t, nu=NU)
        losses = ref_port.base_compute_loss(model, r, [(-1.0, 1.0)], (0.0, 1.0), fns)
        losses["total"].backward()
        torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
        opt.step()
        return float(losses["total"].detach())
    return step, "port"


def time_cpu(n_points, steps, warmup):
    torch.set_num_threads(os.cpu_count() or 1)
    step, kind = cpu_reference_step_fn(n_points)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    return n_points / dt, dt, kind


def cpu_sample_text(kind, extra=""):
    what = ("the unmodified reference (pinnrl 0.3.1 from baseline/_ref: PINNModel + BurgersEquation.compute_loss + backward + "
            "clip + Adam)" if kind == "reference" else "oracle port of the reference autograd path (baseline/_ref absent)")
    return (f"{CPU_SAMPLE_POINTS} of the workload's collocation rows per step (+200 boundary +100 initial rows), {what}, "
            f"torch {torch.__version__} CPU fp32{extra}")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    pps, dt, kind = time_cpu(CPU_SAMPLE_POINTS, args.steps, args.warmup)
    cores = torch.get_num_threads()
    line = {"impl": "reference", "metric": METRIC, "value": pps, "unit": "points/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, 1),
            "cpu_baseline": {"value": pps, "unit": "points/s", "cores": cores, "kind": kind, "sample": cpu_sample_text(kind)},
            "e2e": {"value": pps, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def measure_tf32_tflops(dev, n=8192, seconds=1.0):
    """Dense TF32 tensor-core throughput of this GPU right now (cuBLAS through torch.matmul, fp32 operands with
    allow_tf32), measured the way MEASURED_PEAKS.json measures bf16: best of 10 single launches (burst) and launches back
    to back for ``seconds`` (sustained).  Denominator of the emulated-fp32 tensor roofline only; not on the product path."""
    try:
        prev = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = True
        a = torch.randn(n, n, device=dev)
        b = torch.randn(n, n, device=dev)
        c = torch.empty(n, n, device=dev)
        flops = 2.0 * n ** 3
        for _ in range(3):
            torch.matmul(a, b, out=c)
        best = float("inf")
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); torch.matmul(a, b, out=c); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        reps = max(10, int(seconds * 1e3 / best))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            torch.matmul(a, b, out=c)
        e1.record(); torch.cuda.synchronize()
        sustained = flops * reps / (e0.elapsed_time(e1) * 1e-3) / 1e12
        torch.backends.cuda.matmul.allow_tf32 = prev
        del a, b, c
        return {"burst": flops / (best * 1e-3) / 1e12, "sustained": sustained}
    except Exception:
        return None


def workload_config(args, world):
    return {"workload": "Burgers nu=0.01/pi, feedforward tanh 8x128, 1M collocation pts per B200 (BASELINE configs[1])",
            "points_per_gpu": args.points, "global_points": args.points * world, "jet_columns": JET_COLS,
            "boundary_rows": 200, "initial_rows": 100, "optimizer": "Adam lr=1e-3, clip_grad_norm 1.0 (fused in libpinnk)",
            "parallelism": f"dp{world} (rows sharded, one in-place all-reduce of the step buffer [flat gradient || loss sums])",
            "l2": "flushed between timed steps (256 MiB write); per-step CUDA events summed"}


# --------------------------------------------------------------------------- clocks
class ClockSampler:
    def __init__(self, index):
        self.index, self.proc, self.path = index, None, None

    def start(self):
        try:
            self.path = tempfile.mktemp(suffix=".csv")
            q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
                 "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
                 "clocks_event_reasons.sw_power_cap")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def count(self):
        try:
            return sum(1 for _ in open(self.path))
        except Exception:
            return 0

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in open(self.path):
            f = [c.strip() for c in ln.split(",")]
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
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def load_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


# --------------------------------------------------------------------------- secondary entries: the other BASELINE configs
def make_pde(pk, c, dev):
    s = PDE_SPECS[c["pde"]]
    cfg = pk.PDEConfig(name=c["pde"], domain=[list(d) for d in s["domain"] * c["dim"]], time_domain=list(s["time"]),
                       parameters=dict(s["params"]), boundary_conditions={k: dict(v) for k, v in s["bcs"].items()},
                       initial_condition=dict(s["ic"]), exact_solution=dict(s["exact"]), dimension=c["dim"], device=dev,
                       training=c.get("training"))
    pde = pk.create_pde(c["pde"], cfg)
    pde.compat = c["compat"]
    return pde


def run_config_entry(pk, c, dev, world, rank, tensor_tf32, tensor_bf16, flush):
    """One secondary entry: 2 warm-up + 3 timed steps (CUDA events per step, L2 flushed in between, max over ranks)."""
    import torch.distributed as dist
    from pinns_rl_pde_b200 import engine, parallel, functional as F
    strong = (world == 8 and c["n8"] is not None) or (c.get("fixed_global", False) and world > 1)
    n_global = c["n1"] if c.get("fixed_global", False) else (c["n8"] if strong else c["n1"] * world)
    lo, hi = parallel.shard_bounds(n_global, rank, world)
    n_local = hi - lo
    torch.manual_seed(0)
    model = pk.make_model(c["arch"], c["dim"] + 1, c["hidden"], c["layers"], dev, **c["extra"])
    pde = make_pde(pk, c, dev)
    g = torch.Generator(device="cuda").manual_seed(100 + rank)
    s = PDE_SPECS[c["pde"]]
    x = torch.rand(n_local, c["dim"], generator=g, device=dev) * (s["domain"][0][1] - s["domain"][0][0]) + s["domain"][0][0]
    t = torch.rand(n_local, 1, generator=g, device=dev) * (s["time"][1] - s["time"][0]) + s["time"][0]
    if c["mode"] == "loss":
        cfg = pk.TrainingConfig(learning_rate=1e-3, weight_decay=0.0, gradient_clipping=1.0, scheduler="none")
        # batches of the reference's own size (BASELINE configs[0]: 4 900 points) are launch-bound: replay the step as a CUDA graph
        use_graph = world == 1 and n_local <= 32768
        trainer = pk.PDETrainer(model, pde, config=cfg, device=dev, fused=True, graph=use_graph)

        def step():
            return trainer.train_step(x, t, n_global=n_global)["total"]
    elif c["mode"] == "mse":
        # compute_loss does not exist for dimension >= 2 in the reference: the step is
        # ((compute_residual)**2).mean().backward(), rows sharded, one in-place all-reduce of [flat gradient || sum r^2]
        dirs, kind, p0, cm = F.residual_spec(pde)
        P = F.get_program(model).grad_floats
        buf = torch.zeros(P + 4, dtype=torch.float32, device=dev)
        sums = torch.zeros(1, dtype=torch.float64, device=dev)

        def step():
            eng = engine.get_engine(model, dirs, n_local)
            buf.zero_()
            sums.zero_()
            seg = engine.Segment(kind=kind, row_start=0, row_count=n_local, component=0, weight=1.0 / n_global, p0=p0,
                                 p1=F.residual_p1(pde), compat_math=cm)
            eng.loss_step(x, t, [seg], 1, True, None, buf[:P], sums)
            buf[P:P + 1].copy_(sums)
            parallel.reduce_inplace(buf)
            return buf[P]
    elif c["mode"] == "rar":
        def step():
            xs, ts = parallel.sharded_residual_sample(pde, model, x, t, n_global // 4)
            return xs.shape[0]
    else:
        def step():
            return parallel.sharded_score(pde, model, x, t, want_abs=True)[1]

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    for _ in range(6 if (c["mode"] == "loss" and world == 1 and n_local <= 32768) else 2):     # (graph: eager warm-up calls + capture)
        step()
    evs = []
    sync()
    for _ in range(3):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        step()
        b.record()
        evs.append((a, b))
    sync()
    ms = sum(a.elapsed_time(b) for a, b in evs) / 3
    tt = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms = float(tt.item())
    flops_pt = (2 if c["mode"] in ("score", "rar") else 6) * c["C"] * c["W"]
    pps = n_global / (ms * 1e-3)
    tfl = pps * flops_pt / 1e12 / world                     # per GPU
    del model, pde
    engine._CACHE.clear() if hasattr(engine._CACHE, "clear") else None
    torch.cuda.empty_cache()
    return {"key": c["key"], "baseline_config": c["cfg"], "step": {"loss": "compute_loss + backward + clip + Adam (fused trainer step)",
                                                                    "mse": "mean(compute_residual^2) + backward", "score": "forward-only |r| + statistics",
                                                                    "loss_graph": "compute_loss + backward + clip + Adam (fused trainer step replayed as a CUDA graph)",
                                                                    "rar": "forward-only |r| of the pool + device draw of pool / 4 points (two-level inverse CDF)"}["loss_graph" if (c["mode"] == "loss" and world == 1 and n_local <= 32768) else c["mode"]],
            "global_points": n_global, "points_per_gpu": n_local, "scaling": "strong" if strong else "weak", "n_gpus": world,
            "ms_per_step": ms, "value": pps, "unit": "points/s", "jet_columns": c["C"], "flop_per_point": flops_pt,
            "achieved_tflops_per_gpu": tfl,
            "frac_vs_measured_tf32": (tfl / tensor_tf32) if tensor_tf32 else None,
            "frac_vs_bf16_sixth": tfl / tensor_bf16}


# --------------------------------------------------------------------------- our arm
def run_ours(args):
    import torch.distributed as dist
    import pinns_rl_pde_b200 as pk
    from pinns_rl_pde_b200 import _lib, parallel
    import __graft_entry__ as ge

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if rank == 0:
        ge.build()
    if world > 1:
        dist.barrier()

    torch.manual_seed(0)
    model = pk.make_model("feedforward", 2, HIDDEN, LAYERS, dev)
    pde = pk.BurgersEquation(burgers_cfg(pk, dev))
    n_local, n_global = args.points, args.points * world
    xh, th = synth_points(n_local, 1 + rank)
    xh, th = xh.pin_memory(), th.pin_memory()
    x, t = xh.to(dev), th.to(dev)
    # the package's public trainer step (mirror of PDETrainer's inner step): loss + weighted gradient in one pass per
    # row set, [all-reduce], clip_grad_norm_(1.0) + Adam in libpinnk.  --unfused uses compute_loss().backward() + torch Adam.
    cfg = pk.TrainingConfig(learning_rate=1e-3, weight_decay=0.0, gradient_clipping=1.0, scheduler="none")
    trainer = pk.PDETrainer(model, pde, config=cfg, device=dev, fused=not args.unfused)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def step(xd, td):
        return trainer.train_step(xd, td, n_global=n_global)["total"]

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        """(max over ranks of the summed per-step event times, this rank's own sum)."""
        evs = []
        sync()
        for _ in range(steps):
            flush.fill_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            evs.append((a, b))
        sync()
        ms = sum(a.elapsed_time(b) for a, b in evs)
        tt = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item()), ms

    # clocks are sampled from the first warm-up step on (same load as the timed steps): nvidia-smi needs a few hundred
    # milliseconds before its first sample, longer than a short timed region
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        step(x, t)
    t_wait = time.time()
    while True:                                  # keep every rank loaded until rank 0's sampler is live (collective-safe)
        more = torch.tensor([1 if (rank == 0 and sampler.count() == 0 and time.time() - t_wait < 3.0) else 0], device=dev)
        if world > 1:
            dist.broadcast(more, 0)
        if int(more.item()) == 0:
            break
        step(x, t)
        torch.cuda.synchronize()
    l0 = _lib.launch_count()
    parallel.TIMING = [] if world > 1 else None
    ms_total, ms_mine = timed(lambda: step(x, t), args.steps)
    launches = _lib.launch_count() - l0
    coll_ms = None
    if world > 1:
        torch.cuda.synchronize()
        coll_ms = sum(a.elapsed_time(b) for a, b in parallel.TIMING) / max(args.steps, 1)
        parallel.TIMING = None
    clocks = sampler.stop() if rank == 0 else None
    per_rank = None
    if world > 1:
        mine = torch.tensor([ms_mine / args.steps, coll_ms], dtype=torch.float64, device=dev)
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        st = sorted(float(v[0]) for v in allr)
        cl = sorted(float(v[1]) for v in allr)
        per_rank = ({"min": st[0], "median": st[len(st) // 2], "max": st[-1]},
                    {"min": cl[0], "median": cl[len(cl) // 2], "max": cl[-1],
                     "note": "CUDA events around the step's one all-reduce on each rank: includes waiting for the slowest rank "
                             "to arrive (rank skew), not only the 0.46 MB transfer"})

    if args.lite:
        if rank == 0:
            print(json.dumps({"lite": True, "ms_per_step": ms_total / args.steps, "gpu_launches": int(launches),
                              "value": n_global * args.steps / (ms_total * 1e-3)}))
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # end to end: pinned host rows -> device -> step -> loss back on the host, every step.  Pipelined the way a training loop's
    # loader is: the rows of step i + 1 are copied on a side stream while step i computes (two device buffers), and the loss of
    # step i goes to pinned host memory behind the step and is read by the host during step i + 1 -- so every step still has its
    # own H2D copy of its rows and its own D2H read inside the timed region, but neither drains the launch queue.
    copy_stream = torch.cuda.Stream(device=dev)
    bufs = [(torch.empty_like(x), torch.empty_like(t)) for _ in range(2)]
    loss_probe = step(x, t).detach().reshape(1)
    loss_h = [torch.empty(1, dtype=loss_probe.dtype).pin_memory() for _ in range(2)]
    ev_in = [torch.cuda.Event() for _ in range(2)]
    ev_free = [torch.cuda.Event() for _ in range(2)]
    ev_loss = [torch.cuda.Event() for _ in range(2)]
    e2e_state = {"i": 0, "loss": None}

    def e2e_prefetch(j):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(ev_free[j])          # the step that last read this pair of buffers has finished
            bufs[j][0].copy_(xh, non_blocking=True)
            bufs[j][1].copy_(th, non_blocking=True)
            ev_in[j].record(copy_stream)

    def e2e_step():
        i = e2e_state["i"]
        j = i & 1
        cur = torch.cuda.current_stream()
        cur.wait_event(ev_in[j])
        loss = step(bufs[j][0], bufs[j][1])
        ev_free[j].record(cur)
        loss_h[j].copy_(loss.detach().reshape(1), non_blocking=True)
        ev_loss[j].record(cur)
        e2e_prefetch(j ^ 1)
        if i > 0:
            ev_loss[j ^ 1].synchronize()
            e2e_state["loss"] = float(loss_h[j ^ 1][0])
        e2e_state["i"] = i + 1
    e2e_serial = os.environ.get("PINNK_BENCH_E2E_SERIAL", "0") == "1"      # A/B: copy, step and loss.item() back to back

    def e2e_step_serial():
        bufs[0][0].copy_(xh, non_blocking=True)
        bufs[0][1].copy_(th, non_blocking=True)
        e2e_state["loss"] = float(step(bufs[0][0], bufs[0][1]).item())
    if e2e_serial:
        e2e_step_serial()
        ms_e2e, _ = timed(e2e_step_serial, args.steps)
    else:
        e2e_prefetch(0)
        e2e_step()
        e2e_step()
        ms_e2e, _ = timed(e2e_step, args.steps)
    if not math.isfinite(e2e_state["loss"]):
        raise SystemExit("end-to-end arm: the loss read back on the host is not finite")

    # roofline: profiling pass (event pairs around every libpinnk kernel class)
    roof = None
    peaks = load_peaks()
    tensor_bf16 = peaks.get("bf16_tflops_sustained", 1418.0) / 6.0          # /2 tf32, /3 three-pass split
    tf32 = None
    # every rank runs the two profiling steps (they contain the all-reduce); only rank 0 records events
    if rank == 0:
        _lib.prof_enable(True)
    sync()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(2):
        step(x, t)
    b.record()
    sync()
    tf32 = measure_tf32_tflops(dev)                                         # (every rank: keeps the ranks in step)
    tensor_tf32 = tf32["sustained"] / 3.0 if tf32 else None
    if rank == 0:
        prof = _lib.prof_collect()
        _lib.prof_enable(False)
        step_ms = a.elapsed_time(b) / 2
        gemm = {k: prof[k] for k in ("gemm_fwd", "gemm_dgrad", "gemm_wgrad", "bwd_pair") if k in prof and prof[k][1]}
        dom = max(gemm, key=lambda k: gemm[k][0])
        ms_dom, n_dom = gemm[dom]
        from pinns_rl_pde_b200 import engine as _engine, functional as _F
        chunk = _engine.get_engine(model, _F.residual_spec(pde)[0], n_local).chunk     # the cached engine the steps used
        n_chunks = max(1, -(-n_local // chunk))
        rows_per_launch = JET_COLS * n_local / n_chunks                      # stacked jet rows one launch processes
        # algorithmic bytes per row of 128 floats (DESIGN.md "kernels"): Linear+tanh reads X and writes Y (the pre-activation
        # stash is elided); dgrad+adjoint reads dZ and the stashed Y, writes dZ_prev; wgrad reads dZ and X; the loss-fused last
        # hidden layer (own class fwd_loss_fused) reads X and writes dZ; the paired reverse kernel reads dZ and Y once, writes dZ_prev
        BYTES_PER_ROW = {"gemm_fwd": 2 * 512, "gemm_dgrad": 3 * 512, "gemm_wgrad": 2 * 512, "fwd_loss_fused": 2 * 512,
                         "bwd_pair": 3 * 512}
        # big launches per chunk in each class (value-only BC/IC launches are tiny): 6 hidden layers forward (the 7th is the
        # loss-fused launch), 7 wgrad, 6 fused dgrad+adjoint (the first hidden layer's dgrad is fused with the input layer's
        # reverse: class first_linear_bwd); with PINNK_ENABLE_PAIR=1 six pairs replace 6 dgrad + 6 wgrad
        paired = "bwd_pair" in gemm
        PER_CHUNK = {"gemm_fwd": LAYERS - 2, "gemm_dgrad": 0 if paired else LAYERS - 2, "gemm_wgrad": 1 if paired else LAYERS - 1,
                     "fwd_loss_fused": 1, "bwd_pair": LAYERS - 2}
        bytes_per_row = BYTES_PER_ROW[dom]
        flops_per_launch = 2 * rows_per_launch * HIDDEN * HIDDEN * (2 if dom == "bwd_pair" else 1)
        big = 2 * PER_CHUNK[dom] * n_chunks                                  # two profiled steps
        ms_launch = ms_dom / max(big, 1)
        hbm = peaks.get("hbm_gbs")
        hbm_src = "MEASURED_PEAKS.json hbm_gbs (of measured)"
        if hbm is None:
            hbm, hbm_src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md, of fallback)"
        tensor_peak = tensor_tf32 if tensor_tf32 else tensor_bf16
        tensor_src = (("cuBLAS TF32 8192^3 measured in this run, back to back for 1 s (%.0f TFLOP/s; burst %.0f) / 3 (3xTF32 split)"
                       % (tf32["sustained"], tf32["burst"])) if tf32 else "bf16_tflops_sustained / 2 (tf32) / 3 (3xTF32 split)")
        per_kernel = {}
        for k in BYTES_PER_ROW:
            if k in prof and prof[k][1] and PER_CHUNK[k]:
                ms_k = prof[k][0] / max(2 * PER_CHUNK[k] * n_chunks, 1)
                gbs = BYTES_PER_ROW[k] * rows_per_launch / (ms_k * 1e-3) / 1e9
                per_kernel[k] = {"launch_ms": ms_k, "achieved_gbs": gbs, "frac_hbm": gbs / hbm}
        achieved = bytes_per_row * rows_per_launch / (ms_launch * 1e-3) / 1e9
        traffic, traffic_note = None, None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
            if dom in tj:
                # per launch like `achieved`: the capture's launch processed rows_per_launch(capture) stacked rows, the
                # kernels stream every row once, so DRAM bytes scale with the rows of a launch
                traffic = tj[dom]["bytes_per_launch"] * (rows_per_launch / tj[dom]["rows_per_launch"])
                traffic_note = ("ncu --set full dram bytes of a %d-row launch (%s) x %.0f rows / launch here"
                                % (tj[dom]["rows_per_launch"], tj[dom]["capture"], rows_per_launch))
        except Exception:
            pass
        step_tflops = FLOPS_PER_POINT_STEP * n_local / (ms_total / args.steps * 1e-3) / 1e12
        roof = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm,
                "traffic": traffic, "traffic_source": traffic_note, "peak_source": hbm_src, "algorithmic_bytes_per_launch": bytes_per_row * rows_per_launch,
                "launch_ms": ms_launch, "launches_per_step": n_dom // 2, "kernels": per_kernel,
                "tensor": {"achieved_tflops": flops_per_launch / (ms_launch * 1e-3) / 1e12, "peak_tflops": tensor_peak,
                           "frac": flops_per_launch / (ms_launch * 1e-3) / 1e12 / tensor_peak,
                           "peak_source": tensor_src,
                           "step_achieved_tflops": step_tflops,
                           "frac_vs_measured_tf32": (step_tflops / tensor_tf32) if tensor_tf32 else None,
                           "frac_vs_bf16_sixth": step_tflops / tensor_bf16,
                           "denominators": {"measured_tf32_over_3": tensor_tf32, "bf16_sustained_over_6": tensor_bf16}},
                "why_hbm": "%s: %.1f algorithmic FLOP per byte (2*128*128 FLOP / %d B per row) x %.2f TB/s = %.0f TFLOP/s, "
                           "below the %.0f TFLOP/s emulated-fp32 tensor peak; the 1024 B/row kernels (forward, wgrad) sit at "
                           "%.0f TFLOP/s, i.e. at the crossover"
                           % (dom, 2.0 * HIDDEN * HIDDEN / bytes_per_row, bytes_per_row, hbm / 1e3,
                              2.0 * HIDDEN * HIDDEN / bytes_per_row * hbm / 1e3, tensor_peak,
                              2.0 * HIDDEN * HIDDEN / 1024 * hbm / 1e3),
                "share_of_step": {k: v[0] / 2 / step_ms for k, v in prof.items() if v[1]},
                "step_flops_frac_of_tensor_peak": step_tflops / tensor_peak}

    # the other BASELINE configs (secondary entries); free the headline engine's workspace first
    value = n_global * args.steps / (ms_total * 1e-3)
    e2e_value = n_global * args.steps / (ms_e2e * 1e-3)
    h2d = int(xh.numel() * 4 + th.numel() * 4)
    configs = None
    if not args.no_configs:
        from pinns_rl_pde_b200 import engine as _engine
        del trainer, model, pde, x, t, bufs
        _engine._CACHE.clear() if hasattr(_engine._CACHE, "clear") else None
        torch.cuda.empty_cache()
        configs = []
        for c in CONFIGS:
            try:
                configs.append(run_config_entry(pk, c, dev, world, rank, tensor_tf32, tensor_bf16, flush))
            except Exception as e:                       # a secondary entry must not cost the headline line
                configs.append({"key": c["key"], "error": f"{type(e).__name__}: {e}"[:300]})
                if world > 1:
                    break                                # (ranks could be out of step after a failure)

    if rank == 0:
        cpu = time_cpu(CPU_SAMPLE_POINTS, 3, 1) if world == 1 else None
        line = {"metric": METRIC, "value": value, "unit": "points/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": workload_config(args, world), "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": "points/s", "h2d_bytes_per_step": h2d,
                        "d2h_bytes_per_step": int(loss_h[0].element_size()),
                        "pipelining": "none (PINNK_BENCH_E2E_SERIAL=1)" if e2e_serial else
                                      "rows of step i+1 copied on a side stream during step i; loss of step i read on the host "
                                      "during step i+1"},
                "gpu_launches": int(launches), "roofline": roof}
        if per_rank is not None:
            line["per_rank_step_ms"], line["collective_ms"] = per_rank
        if configs is not None:
            line["configs"] = configs
        if cpu is not None:
            line["cpu_baseline"] = {"value": cpu[0], "unit": "points/s", "cores": torch.get_num_threads(), "kind": cpu[2],
                                    "sample": cpu_sample_text(cpu[2], "; 1 warm-up + 3 steps")}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# --------------------------------------------------------------------------- secondary workload: sharded scoring
def run_score(args):
    """BASELINE configs[4] (not the headline line; `--workload score`): Allen-Cahn, feed-forward 8x128, forward-only
    residual scoring of a candidate pool sharded over the ranks (parallel.sharded_score: per-rank |r| and statistics, two
    tiny collectives for the global statistics).  Weak scaling: --points candidates per GPU (default 8M = 64M / 8)."""
    import torch.distributed as dist
    import pinns_rl_pde_b200 as pk
    from pinns_rl_pde_b200 import _lib, parallel
    import __graft_entry__ as ge
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if rank == 0:
        ge.build()
    if world > 1:
        dist.barrier()
    torch.manual_seed(0)
    model = pk.make_model("feedforward", 2, HIDDEN, LAYERS, dev)
    cfg = pk.PDEConfig(name="allen_cahn", domain=[[-1.0, 1.0]], time_domain=[0.0, 1.0], parameters={"epsilon": 0.1},
                       boundary_conditions={"dirichlet": {"value": 0.0}}, initial_condition={"type": "tanh", "epsilon": 0.1},
                       exact_solution={}, dimension=1, device=dev)
    pde = pk.create_pde("allen_cahn", cfg)
    n = args.points if args.points != (1 << 20) else (1 << 23)
    xh, th = synth_points(n, 1 + rank)
    xh, th = xh.pin_memory(), th.pin_memory()
    x, t = xh.to(dev), th.to(dev)

    def step(xd, td):
        return parallel.sharded_score(pde, model, xd, td, want_abs=True)[1]

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        sync()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(steps):
            fn()
        b.record()
        sync()
        tt = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item())

    for _ in range(args.warmup):
        step(x, t)
    l0 = _lib.launch_count()
    ms = timed(lambda: step(x, t), args.steps)
    launches = _lib.launch_count() - l0
    xd, td = torch.empty_like(x), torch.empty_like(t)

    def e2e_step():
        xd.copy_(xh, non_blocking=True)
        td.copy_(th, non_blocking=True)
        return step(xd, td).cpu()
    e2e_step()
    ms_e2e = timed(e2e_step, args.steps)
    if rank == 0:
        print(json.dumps({
            "metric": "candidate points/sec (forward residual scoring)", "value": n * world * args.steps / (ms * 1e-3),
            "unit": "points/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "Allen-Cahn eps=0.1, feedforward tanh 8x128, candidate residual scoring sharded over the "
                                   "GPUs (BASELINE configs[4])", "points_per_gpu": n, "global_points": n * world,
                       "jet_columns": JET_COLS, "l2": "inputs (64 MB per GPU) and per-layer tensors (2 GB per chunk) exceed the 126 MB L2",
                       "parallelism": f"dp{world} (candidates sharded; global statistics by two 32-byte all-reduces)"},
            "e2e": {"value": n * world * args.steps / (ms_e2e * 1e-3), "unit": "points/s",
                    "h2d_bytes_per_step": int(xh.numel() * 4 + th.numel() * 4), "d2h_bytes_per_step": 32},
            "gpu_launches": int(launches)}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--points", type=int, default=1 << 20, help="collocation rows per GPU")
    ap.add_argument("--unfused", action="store_true", help="autograd route: compute_loss().backward() + torch clip/Adam")
    ap.add_argument("--lite", action="store_true", help="timed steps only (for runs under ncu): no e2e / roofline / cpu passes")
    ap.add_argument("--no-configs", action="store_true", help="skip the secondary entries for the other BASELINE configs")
    ap.add_argument("--workload", default="train", choices=["train", "score"],
                    help="train = the headline metric (BASELINE configs[1]); score = sharded candidate scoring (configs[4])")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "score":
        run_score(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
