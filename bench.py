#!/usr/bin/env python
"""bench.py -- Mcells/s per solver step of the B200-native fluid step (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl ours|reference]

One "step" = one `lib.simulate` call (advect scalar+velocity -> BCs/forces -> divergence ->
pressure solve -> velocity update -> BCs) on a synthetic grid of the named resolution.
Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every field.

Default line: BASELINE.json configs[3], the largest single-GPU configuration (4096^2 plume,
Jacobi-100: 16.8 M cells); the CNN configurations (configs[2] 1024^2 Rayleigh-Taylor ScaleNet,
configs[1] 512^2 plume ScaleNet) are measured in the same run and reported as sub-records under
"also", each with its own roofline / e2e / cpu_baseline.  `--workload NAME` runs one only.
N > 1 (torchrun): the default line is STRONG scaling (the same global grid cut into N slabs,
`--scaling weak` stacks N copies instead); a weak-scaling record rides under "also".

Timing: CUDA events on the launching stream around each step, >= 3 warm-ups, max over ranks;
between timed steps L2 is flushed when the working set is smaller than L2 (config.l2 says which).
`value` times the device-resident step; `e2e` times the same call with HOST (pinned) state:
H2D of the step's inputs + step + D2H of the results inside the timed region.
`--impl reference` times the reference's own ATen CPU path (oracle/_ref, built from
/root/reference by oracle/build_ref.py) on a bounded sample of the same workload.
A wall-clock guard (--hang-timeout) ends a stuck run with the stalled phase and Python stacks.
"""
import argparse
import datetime
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "oracle")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

METRIC = "Mcells/s per solver step (advect+project+update)"
UNIT = "Mcells/s"
L2_BYTES = 126 * 1024 * 1024

# name -> grid, pressure method, and the bounded CPU sample used for the reference arm
WORKLOADS = {
    # BASELINE.json configs[3]: pure stencil HBM-roofline run -- the largest single-GPU configuration
    "plume4096_jacobi100": dict(res=(1, 4096, 4096), method="jacobi", jacobi_iters=100, cpu_sample_res=512,
                                cpu_steps=(1, 2), baseline_config="4096x4096 2D plume, Jacobi 100 iter"),
    # BASELINE.json configs[0]
    "plume128_jacobi28": dict(res=(1, 128, 128), method="jacobi", jacobi_iters=28, cpu_sample_res=128,
                              cpu_steps=(1, 10), baseline_config="128x128 2D plume, Jacobi 28 iter"),
    "plume1024_jacobi100": dict(res=(1, 1024, 1024), method="jacobi", jacobi_iters=100, cpu_sample_res=512,
                                cpu_steps=(1, 2), baseline_config="1024x1024 2D plume, Jacobi 100 iter"),
    # BASELINE.json configs[4] with the Jacobi projection
    "plume256cube_jacobi40": dict(res=(256, 256, 256), method="jacobi", jacobi_iters=40, cpu_sample_res=None,
                                  cpu_steps=(0, 0), baseline_config="256x256x256 3D synthetic grid, Jacobi 40 iter"),
    # BASELINE.json configs[4] as written: CNN pressure on the 256^3 grid.  The reference has no 3-D model
    # (model.py:93); the slice-wise projection FluidNet.forward_fields_3d is this package's definition (parity unpinned)
    "cube256_scalenet_slicewise": dict(res=(256, 256, 256), method="convnet", jacobi_iters=0, cpu_sample_res=None,
                                       cpu_steps=(0, 0),
                                       baseline_config="256x256x256 3D synthetic grid, CNN pressure (slice-wise ScaleNet)"),
    # BASELINE.json configs[1]
    "plume512_scalenet": dict(res=(1, 512, 512), method="convnet", jacobi_iters=0, cpu_sample_res=512,
                              cpu_steps=(1, 5), baseline_config="512x512 2D plume, ScaleNet CNN pressure, fp32"),
    "plume1024_scalenet": dict(res=(1, 1024, 1024), method="convnet", jacobi_iters=0, cpu_sample_res=1024,
                               cpu_steps=(1, 2), baseline_config="1024x1024 2D plume, MultiScale CNN pressure, fp32"),
    # BASELINE.json configs[2]: Rayleigh-Taylor (rayleighTaylorConfig.yaml physics, periodic-y seam)
    "rt1024_scalenet": dict(res=(1, 1024, 1024), method="convnet", jacobi_iters=0, cpu_sample_res=1024, case="rt",
                            cpu_steps=(1, 2),
                            baseline_config="1024x1024 2D Rayleigh-Taylor, MultiScale CNN pressure, fp32"),
}
# The N=1 line of the contract = the largest single-GPU configuration of BASELINE.json.configs (its
# metric names no configuration): configs[3], 16.8 M cells.  The CNN configurations follow in "also".
DEFAULT_WORKLOAD = "plume4096_jacobi100"
ALSO_SINGLE = ("rt1024_scalenet", "plume512_scalenet")
CNN_FLOP_PER_CELL = 484476.0   # 2 * 242238 MAC over the 17 convs of the pyramid
CONFIG_KEYS = ("workload", "baseline_config", "grid", "measured_grid", "pressure", "cells_per_gpu", "parallelism",
               "l2", "launch", "arithmetic", "algorithmic_bytes_per_cell_step")


def plume_mconf(jacobi_iters, method):
    """plumeConfig.yaml physics (pytorch/plumeConfig.yaml) + the solver choice of the workload."""
    return {"dt": 0.1, "maccormackStrength": 0.6, "sampleOutsideFluid": False, "buoyancyScale": 0.25,
            "gravityScale": 0, "viscosity": 0, "correctScalar": False, "operatingDensity": 0.0,
            "gravityVec": {"x": 0.0, "y": -1.0, "z": 0.0}, "pTol": 0.0, "jacobiIter": jacobi_iters,
            "simMethod": method, "injectionDensity": 0.1, "injectionVelocity": 2, "sourceRadius": 0.145}


def rt_mconf(method):
    """rayleighTaylorConfig.yaml physics (pytorch/rayleighTaylorConfig.yaml, rayleighTaylor.py:158-159)."""
    return {"dt": 0.5, "maccormackStrength": 0.6, "sampleOutsideFluid": False, "buoyancyScale": 1.0, "gravityScale": 0,
            "viscosity": 0, "correctScalar": False, "operatingDensity": 0.0,
            "gravityVec": {"x": 0.0, "y": 1.0, "z": 0.0}, "pTol": 0.0, "jacobiIter": 200, "simMethod": method,
            "rho1": -0.01, "rho2": 0.01, "perturbThickness": 100, "perturbAmplitude": 0.01, "height": 0.5,
            "periodic-y": True, "periodic-x": False}


def workload_mconf(wl):
    return rt_mconf(wl["method"]) if wl.get("case") == "rt" else plume_mconf(wl["jacobi_iters"], wl["method"])


def init_state(fluid_mod, wl, mconf, bd, U_np, rho_np, to_tensor):
    """emptyDomain + the driver's initial / boundary conditions on zeroed fields, then the seeded
    synthetic velocity (plume: also density) -- same recipe for the GPU arm and the CPU reference."""
    fluid_mod.emptyDomain(bd["flags"])
    if wl.get("case") == "rt":
        fluid_mod.createRayleighTaylorBCs(bd, mconf, rho1=mconf["rho1"], rho2=mconf["rho2"])
        bd["U"] = to_tensor(U_np * 0.1)          # a small seeded perturbation on the quiescent RT state
        return
    fluid_mod.createPlumeBCs(bd, mconf["injectionDensity"], mconf["injectionVelocity"], mconf["sourceRadius"])
    bd["U"] = to_tensor(U_np)
    bd["density"] = to_tensor(rho_np)


def synthetic_state_numpy(D, H, W, seed=0):
    """Seeded synthetic state (SURVEY.md §8d randomised plume): U ~ N(0, 0.5^2), rho ~ U[0,1), p = 0."""
    import numpy as np
    rng = np.random.RandomState(seed)
    nc = 3 if D > 1 else 2
    U = (rng.standard_normal((1, nc, D, H, W)).astype(np.float32) * np.float32(0.5))
    rho = rng.random_sample((1, 1, D, H, W)).astype(np.float32)
    return U, rho


def pressure_name(wl):
    return f"jacobi x{wl['jacobi_iters']}" if wl["method"] == "jacobi" else "ScaleNet (MultiScaleNet, shipped weights)"


def step_bytes_per_cell(wl):
    """80 + 16 N_iter B/cell (2-D), 104 + 16 N_iter (3-D); the CNN step moves the 80 / 104
    B/cell of stages A, B, D plus the CNN's own I/O"""
    return (80 if wl["res"][0] == 1 else 104) + 16 * wl["jacobi_iters"]


def load_peaks():
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        with open(pk) as f:
            return json.load(f)
    return {}


# ---------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []      # (arrival time, csv line)
        self.t0 = self.t1 = None

    def mark_begin(self):
        self.t0 = time.perf_counter()

    def mark_end(self):
        self.t1 = time.perf_counter()

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.gpu)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        # samples that arrived inside the marked timed region (nvidia-smi reports with ~20 ms period);
        # a region shorter than two periods falls back to every sample taken while this process kept
        # the GPU under the same load (warm-up + timed + per-kernel pass), and says so
        inside = [ln for t, ln in self.lines if self.t0 is not None and self.t0 <= t <= (self.t1 or 1e300) + 0.02]
        window = "timed region"
        if len(inside) < 2:
            inside = [ln for t, ln in self.lines if self.t0 is None or t >= self.t0 - 0.5]
            window = "warm-up + timed region + per-kernel pass (timed region shorter than the sampling period)"
        for ln in inside:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons), "window": window}


def ncu_traffic(kernel_key):
    """dram bytes per launch of a kernel from this round's ncu --set full capture
    (profiles/r2_traffic.json; falls back to the round-1 file, labelled), or (None, None)."""
    for name in ("r2_traffic.json", "r1_traffic.json"):
        try:
            with open(os.path.join(ROOT, "profiles", name)) as f:
                v = json.load(f).get(kernel_key)
            if v is not None:
                return v, f"profiles/{name} (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum, per launch)"
        except OSError:
            pass
    return None, None


# ---------------------------------------------------------------------------------------------
def build_roofline(wl, steps, cells, total_ms, stage_ms, layer_recs, jacobi_launch_iters=8, rec_steps=None):
    """The roofline object of one record.  stage_ms = {"advect_forces", "pressure", "project"} totals
    (ms over `steps` steps of the per-kernel pass)."""
    peaks = load_peaks()
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    peak_src = "MEASURED_PEAKS.json hbm_gbs (measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
    D = wl["res"][0]
    iters = wl["jacobi_iters"]
    dom_ms = stage_ms.get("pressure", 0.0)
    # the stencil group the north star names (fused advect + forces/BCs/divergence + update): stages A+B+D
    grp_ms = stage_ms.get("advect_forces", 0.0) + stage_ms.get("project", 0.0)
    grp_bytes = (80.0 if D == 1 else 104.0) * cells * steps
    group = None
    if wl["method"] == "jacobi" and grp_ms > 0:
        gbs = grp_bytes / (grp_ms / 1e3) / 1e9
        group = {"kernels": ("k2_advect_clean + k2_advect + k2_forces_div + k2_project (csrc/step2d.cu)" if D == 1 else
                             "k_step_advect_fwd + k_step_advect_bwd + k_step_forces_div + k_step_project (csrc/step.cu)"),
                 "algorithmic_bytes_per_cell": 80 if D == 1 else 104, "ms_per_step": round(grp_ms / steps, 4),
                 "achieved": round(gbs, 1), "unit": "GB/s", "frac": round(gbs / hbm_peak, 4),
                 "advect_forces_ms_per_step": round(stage_ms.get("advect_forces", 0.0) / steps, 4),
                 "project_ms_per_step": round(stage_ms.get("project", 0.0) / steps, 4)}
    if wl["method"] == "jacobi":
        # dominant kernel = the Jacobi iterations: 16 B/cell/iteration algorithmic (SURVEY §8d stage C).  The 2-D
        # kernel is temporally blocked (several iterations per launch on register-resident tiles), so the
        # algorithmic figure can exceed the HBM peak; `frac_of_launch_traffic` is the same time against the
        # bytes a launch really has to move (12 R + 4 W per cell per LAUNCH).
        algo_bytes = 16.0 * cells * iters * steps
        achieved = algo_bytes / (dom_ms / 1e3) / 1e9 if dom_ms > 0 else None
        per_launch = jacobi_launch_iters if D == 1 else 1
        nlaunch = (iters + per_launch - 1) // per_launch
        launch_bytes = 16.0 * cells * nlaunch * steps
        tkey = f"k_jacobi2d_blocked @{wl['res'][1]}x{wl['res'][2]}"
        traffic, tsrc = ncu_traffic(tkey) if D == 1 else (None, None)
        roof = {"bound": "hbm",
                "kernel": f"k_jacobi2d_blocked ({per_launch} iterations per launch)" if D == 1
                else "k_jacobi3d_vec (1 iteration per launch)",
                "achieved": round(achieved, 1) if achieved else None, "peak": hbm_peak, "unit": "GB/s",
                "frac": round(achieved / hbm_peak, 4) if achieved else None,
                "traffic": traffic, "traffic_source": tsrc,
                "algorithmic_bytes_per_launch": 16.0 * cells * per_launch,
                "launch_ms": round(dom_ms / steps / nlaunch, 5) if dom_ms > 0 else None,
                "frac_of_launch_traffic": round(launch_bytes / (dom_ms / 1e3) / 1e9 / hbm_peak, 4) if dom_ms > 0 else None,
                "peak_source": peak_src, "stage_ms_per_step": round(dom_ms / steps, 4),
                "step_hbm_frac": round(step_bytes_per_cell(wl) * cells * steps / (total_ms / 1e3) / 1e9 / hbm_peak, 4),
                "stencil_group": group}
        return roof
    # dominant kernel = the tcgen05 conv layer with the largest share of the step, timed per launch with
    # CUDA events (fnx_profile_*).  achieved = ALGORITHMIC fp32-equivalent FLOPs (2*Cin*Cout*k*k*H*W) per
    # launch / launch time; the kernel executes 3x that on the tensor pipe (split-fp16: hi*hi, hi*lo,
    # lo*hi), reported beside it.  Peak = measured bf16 burst (kind::f16 and bf16 share the rate).
    tc_peak = peaks.get("bf16_tflops", 1590.0)
    peak_src_tc = ("MEASURED_PEAKS.json bf16_tflops (measured, burst)" if "bf16_tflops" in peaks
                   else "fallback 1590 TFLOP/s")
    agg = {}
    for r in layer_recs:
        key = (r.cin, r.cout, r.ksize, r.h, r.w, r.tensor)
        a = agg.setdefault(key, [0, 0.0])
        a[0] += 1; a[1] += r.ms
    layers = []
    for (cin, cout, k, h, w, tensor), (n, ms) in agg.items():
        fl = 2.0 * cin * cout * k * k * h * w
        layers.append({"layer": f"{cin}->{cout} k{k} @{h}x{w}", "tensor": bool(tensor), "launches": n,
                       "ms_per_launch": round(ms / n, 4), "algorithmic_tflops": round(fl / (ms / n / 1e3) / 1e12, 2)})
    rec_steps = rec_steps or steps             # steps the per-layer records cover
    cnn_ms = sum(v[1] for v in agg.values()) / max(rec_steps, 1)
    top = max((kv for kv in agg.items() if kv[0][5]), key=lambda kv: kv[1][1], default=None)
    if top is None:
        return {"bound": "tensor", "kernel": None, "achieved": None, "peak": round(tc_peak, 1), "unit": "TFLOP/s",
                "frac": None, "traffic": None, "peak_source": peak_src_tc}
    (cin, cout, k, h, w, _), (n, ms) = top
    fl = 2.0 * cin * cout * k * k * h * w
    achieved = fl / (ms / n / 1e3) / 1e12
    traffic, tsrc = ncu_traffic(f"k_conv_tc {cin}->{cout} k{k} @{h}x{w}")
    return {"bound": "tensor", "kernel": f"k_conv_tc (tcgen05 split-fp16 implicit GEMM) {cin}->{cout} k{k} @{h}x{w}",
            "achieved": round(achieved, 2), "peak": round(tc_peak, 1), "unit": "TFLOP/s",
            "frac": round(achieved / tc_peak, 4), "traffic": traffic, "traffic_source": tsrc,
            "peak_source": peak_src_tc, "launch_ms": round(ms / n, 4), "launches_timed": n,
            "algorithmic_flop_per_launch": fl, "executed_tensor_tflops": round(3 * achieved, 2),
            "executed_tensor_frac": round(3 * achieved / tc_peak, 4),
            "share_of_step": round((ms / rec_steps) / (total_ms / steps), 4),
            "conv_launch_ms_per_step": round(cnn_ms, 4), "stage_ms_per_step": round(dom_ms / steps, 4),
            "forward_algorithmic_tflops": round(CNN_FLOP_PER_CELL * cells * steps / (dom_ms / 1e3) / 1e12, 2)
            if dom_ms > 0 else None,
            "algorithmic_flop_per_cell": CNN_FLOP_PER_CELL, "layers": layers}


def make_record(args, name, wl, world, total_cells, cells, steps, warmup, total_ms, e2e_ms, e2e_steps, roof, h2d, d2h,
                launches, clocks, t_wall, flush, graphed, parallelism, grid, scaling):
    """One record of the contract's JSON line.  cells = cells one GPU computes per step."""
    value = total_cells * steps / (total_ms / 1e3) / 1e6
    e2e_value = total_cells * e2e_steps / (e2e_ms / 1e3) / 1e6
    rec = {
        "metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": world, "steps": steps,
        "warmup": warmup, "ms_per_step": round(total_ms / steps, 4),
        "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f32",
        "data": "synthetic (seeded N(0,0.5^2) velocity, U[0,1) density, plume inlet BCs, border obstacles)",
        "config": {"workload": name, "baseline_config": wl["baseline_config"], "grid": grid, "measured_grid": grid,
                   "pressure": pressure_name(wl), "cells_per_gpu": cells, "parallelism": parallelism,
                   "l2": "flushed between timed steps" if flush else "working set larger than L2, no flush",
                   "launch": "CUDA graph replay of the fused step" if graphed else "direct kernel launches",
                   "arithmetic": ("fp32 stencils (bit-exact vs the reference's ATen CPU path)"
                                  + ("; CNN convs fp32-equivalent on tcgen05: two-term fp16 expansion (hi+lo, 22 "
                                     "significant bits) of activations and weights, 3 MMA terms, fp32 TMEM "
                                     "accumulation -- 1e-5 parity vs torch fp32 (tests/test_gpu_cnn.py)"
                                     if wl["method"] == "convnet" else "")),
                   "algorithmic_bytes_per_cell_step": step_bytes_per_cell(wl)},
        "e2e": {"value": round(e2e_value, 3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": e2e_steps},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roof,
        "wall_s": round(t_wall, 3),
    }
    assert tuple(rec["config"]) == CONFIG_KEYS
    return rec


# ---------------------------------------------------------------------------------------------
def run_single(args, name, guard, local_rank=0, want_cpu_baseline=True):
    """One workload on one GPU through the public lib.simulate."""
    import importlib
    import torch
    from fluidnet_cxx_b200 import _native
    from fluidnet_cxx_b200.lib import fluid
    sim = importlib.import_module("fluidnet_cxx_b200.lib.simulate")

    wl = WORKLOADS[name]
    dev = torch.device("cuda", local_rank)
    lib = _native.load()
    D, H, W = wl["res"]
    cells = D * H * W
    mconf = workload_mconf(wl)
    nc = 3 if D > 1 else 2
    net = None
    if wl["method"] == "convnet":
        from fluidnet_cxx_b200.lib.pretrained import load_scalenet
        net, mconf_net = load_scalenet(dev)
        m = dict(mconf_net); m.update(mconf); mconf = m
        net.mconf = mconf; net.scale.mconf = mconf
    steps, warmup = args.steps, max(args.warmup, 3)

    # ---- synthetic state, built on the HOST (pinned) and copied in ------------------------------
    guard.beat(f"{name}: state")
    U_np, rho_np = synthetic_state_numpy(D, H, W, seed=0)
    host = {"p": torch.zeros(1, 1, D, H, W).pin_memory(), "U": torch.from_numpy(U_np).pin_memory(),
            "flags": torch.zeros(1, 1, D, H, W).pin_memory(), "density": torch.from_numpy(rho_np).pin_memory()}
    bd = {k: torch.zeros_like(v, device=dev) for k, v in host.items()}
    init_state(fluid, wl, mconf, bd, U_np, rho_np, lambda a: torch.from_numpy(a).to(dev))
    for k in ("flags", "U", "density"):
        host[k].copy_(bd[k])
    torch.cuda.synchronize()

    working_set = cells * 4 * (1 + nc + 1 + 1 + 2 * nc + 2)   # state + masks
    flush = working_set < 2 * L2_BYTES
    flush_buf = torch.empty(2 * L2_BYTES // 4, dtype=torch.float32, device=dev) if flush else None

    def one_step():
        with torch.no_grad():
            sim.simulate(mconf, bd, net, wl["method"])

    sampler = ClockSampler(local_rank)
    sampler.start()
    for i in range(warmup):
        guard.beat(f"{name}: warm-up step {i}")
        one_step()
    torch.cuda.synchronize()

    # ---- pass 1: the timed region of `value` (public API; CUDA-graph replay where the step is small) ----
    guard.beat(f"{name}: timed steps")
    step_ms = []
    torch.cuda.synchronize()
    sampler.mark_begin()
    t_wall0 = time.perf_counter()
    for _ in range(steps):
        if flush:
            flush_buf.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        one_step()
        e1.record()
        step_ms.append((e0, e1))
    torch.cuda.synchronize()
    t_wall = time.perf_counter() - t_wall0
    sampler.mark_end()
    total_ms = sum(a.elapsed_time(b) for a, b in step_ms)
    graphed = bool(sim._fusable(mconf, bd, wl["method"]) and sim._graphable(bd, net, wl["method"]))

    # ---- pass 2: the same steps issued stage by stage with CUDA events around every stage (stage hook)
    # and, for the CNN, around every conv launch (fnx_profile_*): roofline inputs.  The launch count is
    # taken here (a graph replay re-issues exactly these kernels).
    guard.beat(f"{name}: per-stage pass")
    stage_events = {}

    def hook(stage, when):
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        stage_events.setdefault(stage, []).append(ev)
    sim.set_stage_hook(hook)
    one_step()
    torch.cuda.synchronize()
    stage_events.clear()
    if wl["method"] == "convnet":
        lib.fnx_profile_enable(1)
    n0 = lib.fnx_launch_count()
    nstage = min(steps, 10 if D == 1 else 1)      # (a 3-D CNN step is 256 forwards: one profiled step is plenty)
    for _ in range(nstage):
        if flush:
            flush_buf.zero_()
        one_step()
    torch.cuda.synchronize()
    launches = (lib.fnx_launch_count() - n0) * steps // nstage
    layer_recs = []
    if wl["method"] == "convnet":
        cap = 8192
        buf = (_native.ProfileRec * cap)()
        n = lib.fnx_profile_fetch(buf, cap)
        lib.fnx_profile_enable(0)
        layer_recs = [buf[i] for i in range(max(0, min(n, cap)))]
    stage_ms = {k: sum(v[2 * i].elapsed_time(v[2 * i + 1]) for i in range(len(v) // 2)) * steps / nstage
                for k, v in stage_events.items()}
    sim.set_stage_hook(None)
    clocks = sampler.stop()

    # ---- e2e: host (pinned) state in, results out, every step -----------------------------------
    # Through lib.HostStepPipeline (lib/host_pipeline.py): every step uploads ITS inputs (U, flags, density; the
    # incoming p is never read by either projection and is not sent) and downloads ITS results (p, U, density);
    # the upload of step k+1, the kernels of step k and the download of step k-1 run on three streams.  The
    # serial figure (upload -> step -> download on one stream, all four tensors sent, as round 1 measured it) is
    # reported beside it.
    guard.beat(f"{name}: e2e steps")
    from fluidnet_cxx_b200.lib.host_pipeline import HostStepPipeline
    e2e_steps = max(3, min(steps, 10))
    out_host = [{k: torch.empty_like(host[k]).pin_memory() for k in ("p", "U", "density")} for _ in range(2)]
    masks = {k: bd[k] for k in ("UBC", "UBCInvMask", "densityBC", "densityBCInvMask") if k in bd}

    def serial_step():
        d = {k: host[k].to(dev, non_blocking=True) for k in ("p", "U", "flags", "density")}
        d.update(masks)
        with torch.no_grad():
            sim.simulate(mconf, d, net, wl["method"])
        for k in ("p", "U", "density"):
            out_host[0][k].copy_(d[k], non_blocking=True)

    serial_step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(e2e_steps):
        serial_step()
    e1.record()
    torch.cuda.synchronize()
    e2e_serial_ms = e0.elapsed_time(e1)
    serial_ref = {k: out_host[0][k].clone() for k in ("p", "U", "density")}

    pipe = HostStepPipeline(mconf, net, wl["method"], like=host, device=dev, masks=masks)
    h2d, d2h = pipe.h2d_bytes_per_step, pipe.d2h_bytes_per_step
    for i in range(2):
        pipe.submit(host, out_host[i % 2])
    pipe.flush()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for i in range(e2e_steps):
        pipe.submit(host, out_host[i % 2])
    pipe.s_out.wait_stream(torch.cuda.current_stream())
    torch.cuda.current_stream().wait_stream(pipe.s_out)     # the last download ends inside the timed region
    e1.record()
    pipe.flush()
    e2e_ms = e0.elapsed_time(e1)
    e2e_mismatch = 0        # both host result buffers hold the same step from the same host state as the serial leg
    for i in range(2):
        for k in ("p", "U", "density"):
            same = (out_host[i][k] == serial_ref[k]) | (torch.isnan(out_host[i][k]) & torch.isnan(serial_ref[k]))
            e2e_mismatch += int((~same).sum())
    if e2e_mismatch:
        print(f"[bench] {name}: {e2e_mismatch} values of the pipelined e2e results differ from the serial ones",
              file=sys.stderr)
    del pipe, serial_ref

    roof = build_roofline(wl, steps, cells, total_ms, stage_ms, layer_recs, rec_steps=nstage)
    # every stage of the per-stage pass (advect+forces[+div] | pressure solve / CNN incl. its wrapper | project), ms per step
    roof["stages_ms_per_step"] = {k: round(v / steps, 4) for k, v in stage_ms.items()}
    out = make_record(args, name, wl, 1, cells, cells, steps, warmup, total_ms, e2e_ms, e2e_steps, roof, h2d, d2h,
                      int(launches), clocks, t_wall, flush, graphed, "1 GPU", [D, H, W], args.scaling)
    out["e2e"]["path"] = ("lib.HostStepPipeline: H2D(k+1) | step(k) | D2H(k-1) on three streams, pinned host buffers; "
                          "p is not uploaded (never read)")
    out["e2e"]["values_differing_from_serial_run"] = e2e_mismatch     # 0: the pipelined results are bit-identical
    out["e2e"]["serial_value"] = round(cells * e2e_steps / (e2e_serial_ms * 1e-3) / 1e6, 3)
    out["e2e"]["serial_note"] = "upload (p, U, flags, density) -> lib.simulate -> download on ONE stream, per step"
    del bd, host, out_host, flush_buf, masks
    sim.clear_graph_cache()
    torch.cuda.empty_cache()
    if want_cpu_baseline:
        guard.beat(f"{name}: cpu_baseline")
        guard_prev, guard.timeout_s = guard.timeout_s, max(guard.timeout_s, 600.0)
        out["cpu_baseline"] = cpu_baseline(wl, *wl["cpu_steps"][::-1])
        guard.timeout_s = guard_prev
    else:
        out["cpu_baseline"] = None
    return out


# ---------------------------------------------------------------------------------------------
def run_distributed(args, name, scaling, guard, transport):
    """N > 1: slab decomposition along the outermost spatial axis (fluidnet_cxx_b200/lib/distributed.py).
    strong = the workload grid itself cut into N slabs; weak = N slabs of the workload grid stacked.
    value = global cells * steps / max-over-ranks time.  Returns the record on rank 0, None elsewhere."""
    import importlib
    import torch
    import torch.distributed as dist
    from fluidnet_cxx_b200 import _native
    from fluidnet_cxx_b200.lib import fluid
    D = importlib.import_module("fluidnet_cxx_b200.lib.distributed")

    wl = WORKLOADS[name]
    rank, world = dist.get_rank(), dist.get_world_size()
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", local_rank)
    lib = _native.load()
    Dz, H, W = wl["res"]
    is3d = Dz > 1
    mconf = workload_mconf(wl)
    net = None
    if wl["method"] == "convnet":
        from fluidnet_cxx_b200.lib.pretrained import load_scalenet
        net, mconf_net = load_scalenet(dev)
        m = dict(mconf_net); m.update(mconf); mconf = m
        net.mconf = mconf; net.scale.mconf = mconf
    ghost = (D.GHOST_CONVNET_3D if Dz > 1 else D.GHOST_CONVNET) if wl["method"] == "convnet" else D.GHOST_JACOBI
    axis = 2 if is3d else 3
    rows_owned = Dz if is3d else H
    if scaling == "strong":
        if rows_owned % world:
            raise SystemExit(f"bench.py: {rows_owned} rows do not split over {world} GPUs")
        rows_owned //= world
    ghost = min(ghost, rows_owned // 4 * 4)
    gD, gH = (rows_owned * world, H) if is3d else (1, rows_owned * world)
    decomp = D.SlabDecomposition(rows_owned * world, ghost, axis=axis, transport=transport)
    cells_global = gD * gH * W
    steps, warmup = args.steps, max(args.warmup, 3)
    tag = f"{name}/{scaling}"

    # the same seeded global state on every rank (pinned host), window rows copied in
    guard.beat(f"{tag}: state")
    U_np, rho_np = synthetic_state_numpy(gD, gH, W, seed=0)
    host = {"p": torch.zeros(1, 1, gD, gH, W).pin_memory(), "U": torch.from_numpy(U_np).pin_memory(),
            "flags": torch.zeros(1, 1, gD, gH, W).pin_memory(), "density": torch.from_numpy(rho_np).pin_memory()}
    bd = {k: torch.zeros_like(v, device=dev) for k, v in host.items()}
    init_state(fluid, wl, mconf, bd, U_np, rho_np, lambda a: torch.from_numpy(a).to(dev))
    for k in ("flags", "U", "density"):
        host[k].copy_(bd[k])
    torch.cuda.synchronize()
    D.check_reach(decomp, bd["U"], mconf["dt"])

    nc = 3 if is3d else 2
    window_cells = decomp.local_rows * (H if is3d else 1) * W
    working_set = window_cells * 4 * (1 + nc + 1 + 1 + 2 * nc + 2)
    flush = working_set < 2 * L2_BYTES
    flush_buf = torch.empty(2 * L2_BYTES // 4, dtype=torch.float32, device=dev) if flush else None

    class TimedOps(D.CudaLocalOps):
        """CUDA events around the pressure stage (Jacobi chunks / CNN) for the roofline"""
        events = []

        def _timed(self, fn, *a):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = fn(*a)
            e1.record()
            self.events.append((e0, e1))
            return out

        def jacobi(self, *a):
            return self._timed(super().jacobi, *a)

        def cnn(self, *a, **kw):
            return self._timed(lambda *x: D.CudaLocalOps.cnn(self, *x, **kw), *a)
    ops = TimedOps()
    ops.events = []

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    # pass 1 (value): the public stepper -- the step replayed as CUDA graphs when it is launch-bound;
    # pass 2 below re-issues it kernel by kernel for the roofline
    guard.beat(f"{tag}: stepper set-up / graph capture")
    # (the 3-D slice-wise CNN step is thousands of small launches: launch-bound at any size)
    graph_limit = (1 << 25) if (is3d and wl["method"] == "convnet") else (1 << 22)
    use_graph = window_cells <= graph_limit and os.environ.get("FLUIDNET_B200_GRAPHS", "1") != "0"
    stepper = D.GraphedDistributedStep(mconf, bd, net, wl["method"], decomp, use_graph=use_graph)
    graphed = stepper.graphed
    guard.beat(f"{tag}: graph-vs-direct check")
    graph_check = stepper.verify() if graphed else None     # graphs vs direct launches, owned rows
    if rank == 0 and stepper.capture_error:
        print(f"[bench] CUDA-graph capture of the distributed step failed, using direct launches: "
              f"{stepper.capture_error}", file=sys.stderr)

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for i in range(warmup):
        guard.beat(f"{tag}: warm-up step {i}")
        stepper.step()
    barrier()
    step_ms = []
    barrier()
    sampler.mark_begin()
    t_wall0 = time.perf_counter()
    for i in range(steps):
        guard.beat(f"{tag}: timed step {i}")
        if flush:
            flush_buf.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        stepper.step()
        e1.record()
        step_ms.append((e0, e1))
    barrier()
    t_wall = time.perf_counter() - t_wall0
    sampler.mark_end()
    total_ms = sum(a.elapsed_time(b) for a, b in step_ms)

    # pass 2: direct launches with per-stage / per-layer events
    guard.beat(f"{tag}: per-stage pass")
    bd = dict(stepper.state)      # a COPY of the dict: simulate_distributed re-binds its entries (see GraphedDistributedStep)

    def one_step():
        with torch.no_grad():
            D.simulate_distributed(mconf, bd, net, wl["method"], decomp, ops=ops)
    one_step()
    barrier()
    ops.events.clear()
    if wl["method"] == "convnet":
        lib.fnx_profile_enable(1)
    n0 = lib.fnx_launch_count()
    barrier()
    for i in range(steps):
        guard.beat(f"{tag}: per-stage step {i}")
        if flush:
            flush_buf.zero_()
        one_step()
    barrier()
    launches = lib.fnx_launch_count() - n0
    dom_ms = sum(a.elapsed_time(b) for a, b in ops.events)
    layer_recs = []
    if wl["method"] == "convnet":
        buf = (_native.ProfileRec * 4096)()
        n = lib.fnx_profile_fetch(buf, 4096)
        lib.fnx_profile_enable(0)
        layer_recs = [buf[i] for i in range(max(0, min(n, 4096)))]
    clocks = sampler.stop() if rank == 0 else None

    # ---- e2e: this rank's window rows in from pinned host memory, owned rows out, every step ----
    # Contiguous pinned buffers <-> contiguous device staging buffers (true asynchronous DMA), then device-side
    # copies into / out of the strided window.  (Round 1 copied the strided window views directly: torch stages
    # such copies through pageable host memory, i.e. synchronously, and that e2e leg is where the 4-GPU run of
    # round 1 stopped making progress; every buffer is also allocated up front now, no cudaHostAlloc between
    # collectives.)
    guard.beat(f"{tag}: e2e steps")
    e2e_steps = max(3, min(steps, 10))
    win, own = decomp._sl(decomp.r0, decomp.r1), decomp._sl(decomp.lo, decomp.hi)
    host_win = {k: host[k][win].contiguous().pin_memory() for k in ("p", "U", "flags", "density")}
    dev_win = {k: torch.empty_like(v, device=dev) for k, v in host_win.items()}
    out_host = {k: torch.empty(host[k][own].shape).pin_memory() for k in ("p", "U", "density")}
    dev_out = {k: torch.empty(v.shape, device=dev) for k, v in out_host.items()}
    h2d = sum(v.numel() * 4 for v in host_win.values())
    d2h = sum(v.numel() * 4 for v in out_host.values())
    barrier()

    e2e_sync = os.environ.get("FNX_E2E_MODE", "async") == "sync"

    def e2e_step():
        for k in ("p", "U", "flags", "density"):
            dev_win[k].copy_(host_win[k], non_blocking=True)
            stepper.state[k][win].copy_(dev_win[k])
        if e2e_sync:
            torch.cuda.synchronize()      # (experiment switch of the round-2 hang hunt; DESIGN.md "the 4-GPU hang")
        stepper.step()
        if e2e_sync:
            torch.cuda.synchronize()
        for k in ("p", "U", "density"):
            dev_out[k].copy_(stepper.state[k][own])
            out_host[k].copy_(dev_out[k], non_blocking=True)

    e2e_step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(e2e_steps):
        e2e_step()
    e1.record()
    barrier()
    e2e_ms = e0.elapsed_time(e1)

    t = torch.tensor([total_ms, e2e_ms, dom_ms], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms, dom_ms = t.tolist()
    tb = torch.tensor([float(h2d), float(d2h), float(launches)], dtype=torch.float64, device=dev)
    dist.all_reduce(tb)
    h2d_all, d2h_all, launches_all = (int(x) for x in tb.tolist())
    out = None
    if rank == 0:
        roof = build_roofline(wl, steps, window_cells, total_ms, {"pressure": dom_ms}, layer_recs)
        out = make_record(args, name, wl, world, cells_global, window_cells, steps, warmup, total_ms, e2e_ms, e2e_steps,
                          roof, h2d_all, d2h_all, launches_all, clocks, t_wall, flush, graphed,
                          parallelism=(f"{world} GPUs, {scaling} scaling: slab decomposition along {'D' if is3d else 'H'} "
                                       f"({rows_owned} owned + {ghost} ghost rows per interior side), halo exchange by "
                                       + ("stores into the neighbours' inboxes over NVLink peer memory"
                                          if decomp.transport == "peer" else "NCCL send/recv")
                                       + f", global grid {gD}x{gH}x{W}"),
                          grid=[gD, gH, W], scaling=scaling)
        out["cpu_baseline"] = None
        if graph_check is not None:
            out["multi_gpu"] = {"graph_vs_direct_max_abs_diff": graph_check}
    del stepper, bd, host, out_host, flush_buf, ops, host_win, dev_win, dev_out
    torch.cuda.empty_cache()
    barrier()
    return out


def plume_held_state(wl, mconf, gH, W, seed=0):
    """held(name, r0, r1) -> rows [r0, r1) of the GLOBAL initial state of a 2-D plume workload as CPU tensors
    (1, C, 1, r1-r0, W), built without ever materialising the global grid on a GPU: emptyDomain (util.py:5-49)
    and createPlumeBCs (init_conditions.py:4-83) only touch the border ring and rows 0:4."""
    import numpy as np
    import torch
    from fluidnet_cxx_b200.lib.fluid import init_conditions
    U_np, rho_np = synthetic_state_numpy(1, gH, W, seed=seed)
    small = {"p": torch.zeros(1, 1, 1, 8, W), "U": torch.zeros(1, 2, 1, 8, W), "flags": torch.zeros(1, 1, 1, 8, W),
             "density": torch.zeros(1, 1, 1, 8, W)}
    init_conditions.createPlumeBCs(small, mconf["injectionDensity"], mconf["injectionVelocity"], mconf["sourceRadius"])
    ident = {"UBC": 0.0, "UBCInvMask": 1.0, "densityBC": 0.0, "densityBCInvMask": 1.0}

    def held(name, r0, r1):
        if name == "U":
            return torch.from_numpy(np.ascontiguousarray(U_np[:, :, :, r0:r1]))
        if name == "density":
            return torch.from_numpy(np.ascontiguousarray(rho_np[:, :, :, r0:r1]))
        if name == "flags":
            f = torch.ones(1, 1, 1, r1 - r0, W)
            f[..., 0] = 2.0; f[..., W - 1] = 2.0
            if r0 == 0:
                f[:, :, :, 0] = 2.0
            if r1 == gH:
                f[:, :, :, -1] = 2.0
            return f
        if name in ident:
            c = 2 if name.startswith("U") else 1
            t = torch.full((1, c, 1, r1 - r0, W), ident[name])
            n = max(0, min(8, r1) - r0) if r0 < 8 else 0
            if n > 0:
                t[:, :, :, :n] = small[name][:, :, :, r0:r0 + n]
            return t
        return None
    return held


def run_distributed_slab(args, name, scaling, guard):
    """N > 1, 2-D Jacobi workloads: the window-based slab step (fluidnet_cxx_b200/lib/slab.py): every rank
    holds its rows + ghost rows only, halos travel by one peer-memory kernel per exchange, the whole step
    is one CUDA graph.  strong = the workload grid cut into N slabs; weak = N workload grids stacked."""
    import torch
    import torch.distributed as dist
    from fluidnet_cxx_b200 import _native
    from fluidnet_cxx_b200.lib import slab

    wl = WORKLOADS[name]
    rank, world = dist.get_rank(), dist.get_world_size()
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", local_rank)
    lib = _native.load()
    _, H, W = wl["res"]
    mconf = workload_mconf(wl)
    gH = H if scaling == "strong" else H * world
    if gH % world:
        raise SystemExit(f"bench.py: {gH} rows do not split over {world} GPUs")
    steps, warmup = args.steps, max(args.warmup, 3)
    tag = f"{name}/{scaling}/slab"
    guard.beat(f"{tag}: state")
    held = plume_held_state(wl, mconf, gH, W)
    topo = slab.ProcessTopology(dev)
    K = int(os.environ.get("FNX_SLAB_K", "3"))   # Jacobi launches per pressure exchange (measured: 1.52 / 1.43 / 1.41 ms at K = 1 / 2 / 3, 2 GPUs strong)
    step = slab.SlabJacobiStep(topo, mconf, gH, W, held, K=K)
    g = step.g
    cells_global = gH * W
    owned_cells = g["Hs"] * W
    held_cells = g["rows_held"] * W

    # pinned host copies of this rank's held rows (inputs of the e2e leg) -- allocated BEFORE any timed region
    host_in = {k: held(k, g["ya0"], g["ya1"]).pin_memory() for k in ("U", "density", "flags")}
    host_out = {k: torch.empty((1, c, 1, g["Hs"], W)).pin_memory() for k, c in (("p", 1), ("U", 2), ("density", 1))}
    working_set = held_cells * 4 * 14
    flush = working_set < 2 * L2_BYTES
    flush_buf = torch.empty(2 * L2_BYTES // 4, dtype=torch.float32, device=dev) if flush else None

    def barrier():
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()

    guard.beat(f"{tag}: direct steps + graph capture")
    for _ in range(2):
        step.step()
    barrier()
    graphed = step.capture()
    barrier()
    reach = step.check_reach()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for i in range(warmup):
        guard.beat(f"{tag}: warm-up step {i}")
        step.step()
    barrier()
    guard.beat(f"{tag}: timed steps")
    ev = []
    sampler.mark_begin()
    t_wall0 = time.perf_counter()
    for i in range(steps):
        if flush:
            flush_buf.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        step.step()
        e1.record()
        ev.append((e0, e1))
    barrier()
    t_wall = time.perf_counter() - t_wall0
    sampler.mark_end()
    total_ms = sum(a.elapsed_time(b) for a, b in ev)

    # pass 2: the same phases launched one by one with events per kind (exchange / advect+forces / jacobi / project)
    guard.beat(f"{tag}: per-stage pass")
    kinds = [op[0] for op in slab.schedule(gH, world, rank, step.iters, K)[1]]
    stage = {}
    n0 = lib.fnx_launch_count()
    nstage = max(3, min(steps, 10))
    for _ in range(nstage):
        for kind, ph in zip(kinds, step.phases()):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ph()
            e1.record()
            stage.setdefault(kind, []).append((e0, e1))
        step.advance()
    barrier()
    launches = (lib.fnx_launch_count() - n0) * steps // nstage
    stage_ms = {k: sum(a.elapsed_time(b) for a, b in v) * steps / nstage for k, v in stage.items()}
    clocks = sampler.stop() if rank == 0 else None

    # e2e: held rows in from pinned host memory, owned rows out, every step
    guard.beat(f"{tag}: e2e steps")
    e2e_steps = max(3, min(steps, 10))
    h2d = sum(v.numel() * 4 for v in host_in.values())
    d2h = sum(v.numel() * 4 for v in host_out.values())

    # Every step uploads its held rows and downloads its owned rows; the copies go through device staging buffers
    # on two copy streams, so the upload of step k+1 and the download of step k-1 overlap the kernels of step k (the
    # stepper's own double buffer cannot take the upload directly: step k+1 reads the buffer step k wrote).
    s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    stage_in = {k: torch.empty(v.shape, device=dev) for k, v in host_in.items()}
    stage_out = {k: torch.empty(v.shape, device=dev) for k, v in host_out.items()}
    ev_in, ev_consumed = torch.cuda.Event(), torch.cuda.Event()
    ev_out_ready, ev_out_done = torch.cuda.Event(), torch.cuda.Event()
    cur = torch.cuda.current_stream(dev)
    ev_consumed.record(cur)
    ev_out_done.record(cur)

    def e2e_step():
        s_in.wait_event(ev_consumed)                       # the staging buffers were read by the previous step
        with torch.cuda.stream(s_in):
            for k in ("U", "density", "flags"):
                stage_in[k].copy_(host_in[k], non_blocking=True)
            ev_in.record(s_in)
        cur.wait_event(ev_in)
        step.held("U").copy_(stage_in["U"])
        step.held("density").copy_(stage_in["density"])
        step.flags.copy_(stage_in["flags"])
        ev_consumed.record(cur)
        step.step()
        cur.wait_event(ev_out_done)                        # the previous results have left the staging buffers
        for k in ("p", "U", "density"):
            stage_out[k].copy_(step.owned(k))
        ev_out_ready.record(cur)
        s_out.wait_event(ev_out_ready)
        with torch.cuda.stream(s_out):
            for k in ("p", "U", "density"):
                host_out[k].copy_(stage_out[k], non_blocking=True)
            ev_out_done.record(s_out)

    def e2e_drain():
        cur.wait_stream(s_out)
    e2e_step()
    e2e_drain()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(e2e_steps):
        e2e_step()
    e2e_drain()                                            # the last download ends inside the timed region
    e1.record()
    barrier()
    e2e_ms = e0.elapsed_time(e1)

    t = torch.tensor([total_ms, e2e_ms] + [stage_ms.get(k, 0.0) for k in ("X", "advect", "jacobi", "project")],
                     dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms, x_ms, adv_ms, jac_ms, prj_ms = t.tolist()
    tb = torch.tensor([float(h2d), float(d2h), float(launches), float(torch.cuda.max_memory_allocated(dev))],
                      dtype=torch.float64, device=dev)
    dist.all_reduce(tb)
    h2d_all, d2h_all, launches_all, mem_all = (int(x) for x in tb.tolist())
    out = None
    if rank == 0:
        roof = build_roofline(wl, steps, owned_cells, total_ms, {"pressure": jac_ms, "advect_forces": adv_ms, "project": prj_ms},
                              [])
        roof["exchange_ms_per_step"] = round(x_ms / steps, 4)
        roof["exchanges_per_step"] = kinds.count("X")
        out = make_record(args, name, wl, world, cells_global, owned_cells, steps, warmup, total_ms, e2e_ms, e2e_steps, roof,
                          h2d_all, d2h_all, launches_all, clocks, t_wall, flush, graphed,
                          parallelism=(f"{world} GPUs, {scaling} scaling: row slabs ({g['Hs']} owned rows, {g['G']} ghost rows "
                                       f"held per interior side, every rank allocates its window only), halo rows stored "
                                       f"straight into the neighbours' ghost rows over NVLink peer memory by one kernel "
                                       f"per exchange ({kinds.count('X')} per step, no NCCL in the data path), whole step "
                                       f"replayed as one CUDA graph; global grid 1x{gH}x{W}"),
                          grid=[1, gH, W], scaling=scaling)
        out["e2e"]["path"] = ("per rank: H2D(k+1) | step(k) | D2H(k-1) on three streams through device staging buffers, pinned "
                              "host buffers; held rows in (U, density, flags), owned rows out (p, U, density)")
        out["cpu_baseline"] = None
        # (kept out of `config`: both arms of the driver's comparison carry exactly CONFIG_KEYS there)
        out["multi_gpu"] = {"max_u_dt_cells": round(reach, 3), "device_bytes_all_ranks_peak": mem_all,
                            "jacobi_launches_per_exchange": K, "ghost_rows": g["G"]}
    del step, host_in, host_out, flush_buf
    torch.cuda.empty_cache()
    barrier()
    return out



def run_ours(args):
    import torch
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback in the product path)"
    torch.cuda.set_device(local_rank)
    from fluidnet_cxx_b200.lib.hang_guard import HangGuard
    guard = HangGuard(args.hang_timeout, who=f"bench.py rank {rank}/{world}")
    name = args.workload or DEFAULT_WORKLOAD
    if world == 1:
        out = run_single(args, name, guard, local_rank, want_cpu_baseline=not args.no_cpu_baseline)
        if args.workload is None and not args.no_also:
            out["also"] = [run_single(args, w, guard, local_rank, want_cpu_baseline=not args.no_cpu_baseline)
                           for w in ALSO_SINGLE]
        guard.stop()
        emit(out)
        return
    import torch.distributed as dist
    dev = torch.device("cuda", local_rank)
    guard.beat("init_process_group")
    dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=max(60.0, 1.5 * args.hang_timeout)))
    def run_n(nm, scaling):
        w = WORKLOADS[nm]
        if w["method"] == "jacobi" and w["res"][0] == 1 and w.get("case") != "rt" and args.transport != "nccl":
            return run_distributed_slab(args, nm, scaling, guard)
        return run_distributed(args, nm, scaling, guard, "nccl" if args.transport == "window" else args.transport)
    out = run_n(name, args.scaling)
    if args.workload is None and not args.no_also:
        other = "weak" if args.scaling == "strong" else "strong"
        also = [run_n(name, other)]
        if rank == 0:
            out["also"] = also
    guard.stop()
    if rank == 0:
        emit(out)
    dist.barrier()
    dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------
def reference_step_runner(wl, res):
    """The reference's own CPU path (oracle/_ref) on a res x res sample of the workload."""
    import torch
    import ref_loader
    if not ref_loader.available():
        return None
    net = None
    mconf = workload_mconf(wl)
    if wl["method"] == "convnet":
        reflib, net, mconf_net = ref_loader.load_scalenet()
        m = dict(mconf_net); m.update(mconf); mconf = m
        net.mconf = mconf; net.scale.mconf = mconf
    else:
        reflib = ref_loader.load()
    U_np, rho_np = synthetic_state_numpy(1, res, res, seed=0)
    bd = {"p": torch.zeros(1, 1, 1, res, res), "U": torch.zeros(1, 2, 1, res, res),
          "flags": torch.zeros(1, 1, 1, res, res), "density": torch.zeros(1, 1, 1, res, res)}
    init_state(reflib.fluid, wl, mconf, bd, U_np, rho_np, torch.from_numpy)

    def step():
        with torch.no_grad():
            reflib.simulate(mconf, bd, net, wl["method"])
    return step


def _all_host_threads():
    import torch
    try:
        torch.set_num_threads(max(1, len(os.sched_getaffinity(0))))
    except (AttributeError, RuntimeError):
        pass


def sample_text(wl, res, steps, warm):
    import torch
    full = res == wl["res"][1] == wl["res"][2]
    return (f"{steps} timed step(s) (+{warm} warm-up) of the reference's ATen CPU path (patched build oracle/_ref, "
            f"torch {torch.__version__}, {os.cpu_count()} host CPUs) on a {res}x{res} grid"
            + (" = the full workload" if full else
               f": same physics, reduced grid (minutes per step at the full {wl['res'][1]}x{wl['res'][2]} size)"))


def cpu_baseline(wl, steps, warmup):
    import torch
    _all_host_threads()
    res = wl["cpu_sample_res"]
    if res is None:
        return {"value": None, "unit": UNIT, "cores": 0, "kind": "reference",
                "sample": "none: the reference asserts 3-D off (advection.py:58,108); no CPU reference exists"}
    step = reference_step_runner(wl, res)
    if step is None:
        return {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": "oracle/_ref not built"}
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    out = {"value": round(res * res * steps / dt / 1e6, 4), "unit": UNIT, "cores": torch.get_num_threads(),
           "kind": "reference", "seconds": round(dt, 2), "sample": sample_text(wl, res, steps, warmup)}
    # SURVEY.md section 8d also asks for the single-thread figure: one more step of the same sample on one thread
    n_threads = torch.get_num_threads()
    try:
        torch.set_num_threads(1)
        t0 = time.perf_counter()
        step()
        out["value_1thread"] = round(res * res / (time.perf_counter() - t0) / 1e6, 4)
    finally:
        torch.set_num_threads(n_threads)
    return out


def run_reference(args):
    """The reference arm: the reference's own ATen CPU implementation of the step (oracle/_ref) with every
    host thread, the requested --steps / --warmup, on a bounded sample of the default workload."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    name = args.workload or DEFAULT_WORKLOAD
    wl = WORKLOADS[name]
    res = wl["cpu_sample_res"]
    if res is None:
        emit({"impl": "reference", "unavailable": "the reference has no runnable 3-D path"})
        return
    import torch
    # torchrun exports OMP_NUM_THREADS=1 to its workers: the reference arm is the only work on this box while
    # it runs, give it every host thread
    _all_host_threads()
    step = reference_step_runner(wl, res)
    if step is None:
        emit({"impl": "reference", "unavailable": "oracle/_ref is not built on this box"})
        return
    steps, warm = max(1, args.steps), max(0, args.warmup)
    for _ in range(warm):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    value = res * res * steps / dt / 1e6
    sample = sample_text(wl, res, steps, warm)
    cores = torch.get_num_threads()
    out = {"impl": "reference", "metric": METRIC, "value": round(value, 4), "unit": UNIT, "n_gpus": world,
           "steps": steps, "warmup": warm, "ms_per_step": round(dt / steps * 1e3, 2), "higher_is_better": True,
           "scaling": args.scaling, "vs_baseline": None, "dtype": "f32",
           "data": "synthetic (same generator as the GPU arm)",
           "config": {"workload": name, "baseline_config": wl["baseline_config"], "grid": list(wl["res"]),
                      "measured_grid": [1, res, res], "pressure": pressure_name(wl), "cells_per_gpu": 0,
                      "parallelism": f"CPU, {cores} threads (ATen intra-op parallelism)", "l2": "n/a (CPU)",
                      "launch": "reference lib.simulate, op by op",
                      "arithmetic": "fp32 ATen CPU (the reference's own path)",
                      "algorithmic_bytes_per_cell_step": step_bytes_per_cell(wl)},
           "cpu_baseline": {"value": round(value, 4), "unit": UNIT, "cores": cores, "kind": "reference",
                            "sample": sample},
           "e2e": {"value": round(value, 4), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    assert tuple(out["config"]) == CONFIG_KEYS
    emit(out)


_RESULT_OUT = None


def claim_stdout():
    """stdout carries ONE JSON line (the driver parses it).  Libraries write there too -- NCCL prints its version
    banner on stdout when NCCL_DEBUG is set in the environment -- so file descriptor 1 is pointed at stderr for the
    whole run and the record goes out through a private duplicate of the original stdout."""
    global _RESULT_OUT
    if _RESULT_OUT is None:
        sys.stdout.flush()
        _RESULT_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(record):
    out = _RESULT_OUT or sys.stdout
    out.write(json.dumps(record) + "\n")
    out.flush()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS),
                    help=f"run this workload only (default: {DEFAULT_WORKLOAD}, plus the sub-records under 'also')")
    ap.add_argument("--no-also", action="store_true", help="default workload only, no sub-records")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--transport", default="window", choices=["window", "peer", "nccl"],
                    help="N > 1: window (default) = per-rank row windows + one peer-memory halo kernel per exchange, "
                         "whole step one CUDA graph (lib/slab.py; 2-D Jacobi workloads, others fall back to nccl); "
                         "nccl / peer = the global-array decomposition of lib/distributed.py with batched NCCL "
                         "send/recv, or stores into the neighbour's inbox")
    ap.add_argument("--scaling", default="strong", choices=["weak", "strong"],
                    help="N > 1: strong = the workload grid itself split into N slabs (default); "
                         "weak = N slabs of the workload grid stacked along H/D")
    ap.add_argument("--hang-timeout", type=float, default=120.0,
                    help="seconds without progress before the run is ended with a diagnosis (0 = off)")
    args = ap.parse_args()
    claim_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
