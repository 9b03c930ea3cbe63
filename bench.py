#!/usr/bin/env python
"""bench.py -- KPM DOS throughput (nnz * moments * vectors / s) on 1..8 B200, next to the CPU baseline.

Workload (BASELINE.json configs[1], the configuration the metric is quoted on):
    graphene.monolayer() 1000 x 1000 nm (~38.2 M sites, complex64 with a Peierls field), calc_dos,
    2050 moments (the 4k+2 number next to 2048), 64 stochastic vectors, sharded over the ranks.
One "step" = one complete moments phase (device MT19937 starters + Chebyshev recursion + reductions +
the NCCL allreduce) for all 64 vectors, i.e. what the reference's `moments_timer` covers (Core.cpp:152-156).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl reference]

For N > 1 launch with torch.distributed.run (one rank per GPU); torch is used only for rendezvous plumbing.
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (builder kwargs, num_moments, num_random, energy_range, description)
    "graphene_1000nm_c64_dos": dict(kind="graphene", size=1000.0, field=10.0, dtype="complex64",
                                    moments=2050, vectors=64, energy_range=(-8.5, 8.5),
                                    text="graphene.monolayer() 1000x1000 nm (~38M sites) calc_dos, 2050 moments, "
                                         "64 random vectors, complex64"),
    "graphene_40nm_f32_dos": dict(kind="graphene", size=40.0, field=0.0, dtype="float32",
                                  moments=1026, vectors=1, energy_range=(-8.5, 8.5),
                                  text="graphene.monolayer() 40x40 nm calc_dos, 1026 moments, 1 random vector, float32"),
    "cubic_256_f32_dos": dict(kind="cubic", size=256, dtype="float32", moments=4098, vectors=128,
                              energy_range=(-8.2, 8.2),
                              text="simple-cubic Anderson 256^3 calc_dos, 4098 moments, 128 random vectors, float32"),
    "graphene_200nm_c64_dos": dict(kind="graphene", size=200.0, field=10.0, dtype="complex64",
                                   moments=514, vectors=64, energy_range=(-8.5, 8.5),
                                   text="(reduced, for quick checks) graphene 200x200 nm calc_dos, 514 moments, 64 vectors"),
}
# BASELINE configs[2] and configs[3]: not the headline metric, each prints its own line (own metric name, own roofline)
QUANTITY_WORKLOADS = {
    "graphene_500nm_c128_ldos": dict(kind="graphene", size=500.0, field=10.0, disorder=0.5, dtype="complex128", quantity="ldos",
                                     sites=256, broadening=0.02, energy_range=(-8.8, 8.8),
                                     text="graphene 500x500 nm + onsite disorder + Peierls field, complex128, calc_ldos at 256 sites "
                                          "(16 x 16 grid), broadening 0.02 eV"),
    "graphene_500nm_c128_greens": dict(kind="graphene", size=500.0, field=10.0, disorder=0.5, dtype="complex128", quantity="greens",
                                       sites=256, broadening=0.02, energy_range=(-8.8, 8.8),
                                       text="graphene 500x500 nm + onsite disorder + Peierls field, complex128, calc_greens centre -> "
                                            "256 sites within 40 nm, broadening 0.02 eV"),
    "graphene_200nm_f64_conductivity": dict(kind="graphene", size=200.0, field=0.0, disorder=0.0, dtype="float64", quantity="conductivity",
                                            vectors=4, points=1000, energy_range=(-9.0, 9.0),
                                            text="graphene 200x200 nm calc_conductivity xx and xy, 514 x 514 Kubo-Bastin moments, "
                                                 "4 random vectors, float64, 1000 points, T = 300 K"),
}
FP64_TENSOR_PEAK = 37.1   # TFLOP/s, DMMA.8x8x4 measured on this pool's B200 (profiles/r01_fp64_mma_probe.jsonl); not in MEASURED_PEAKS.json
DEFAULT_WORKLOAD = "graphene_1000nm_c64_dos"
METRIC = "KPM nnz*moments*vectors/s (graphene DOS)"
UNIT = "nnz*moments*vectors/s"


def build_model(w):
    import pybinding_b200 as pb
    if w["kind"] == "graphene":
        return pb.graphene_rectangle(w["size"], magnetic_field=w["field"], dtype=np.dtype(w["dtype"]),
                                     disorder=w.get("disorder", 0.0), disorder_seed=0)
    return pb.cubic_anderson(w["size"], disorder=4.0, seed=0, dtype=np.dtype(w["dtype"]))


class ClockSampler:
    """Samples SM clocks / throttle reasons with nvidia-smi while the timed region runs"""
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,clocks.mem")

    def __init__(self, device):
        self.device = device
        self.proc = None
        self.lines = []
        self.begin = 0.0

    def mark_begin(self):
        """Samples that arrive from now on count.  nvidia-smi is started ahead of the timed region (its start-up takes a
        few hundred milliseconds during which driver calls of this process stall -- visible in sub-second steps) and the
        samples of the warm-up are dropped here."""
        self.begin = time.perf_counter()

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.QUERY,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def snapshot(self):
        """One sample, synchronously.  Every nvidia-smi query stalls the GPU's launch path for a moment: polled at 5 Hz
        that doubles the duration of launch-bound steps (configs[0]: 1.5 ms per moments phase without the poll, 2.4 - 4.8 ms
        with it, profiles/r02_bench_40nm_variance.log) while a 12 s step does not notice.  Timed regions shorter than two
        seconds are therefore bracketed by one sample before and one after instead of being polled."""
        try:
            out = subprocess.run(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.QUERY, "--format=csv,noheader,nounits"],
                                 capture_output=True, text=True, timeout=20).stdout
            for line in out.splitlines():
                if line.strip():
                    self.lines.append((time.perf_counter(), line.strip()))
            self.bracket = True
        except (OSError, subprocess.SubprocessError):
            pass

    def begin_region(self, expected_seconds):
        """Call right before the timed region: polls when the region is long enough not to notice, brackets it otherwise."""
        if expected_seconds >= 2.0:
            self.start()
            time.sleep(0.8)          # nvidia-smi's start-up stalls driver calls of this process: let it pass
            self.mark_begin()
        else:
            self.mark_begin()
            self.snapshot()
            self.after_snapshot = True   # the caller runs one more untimed step: the first launch after a query is slow

    def end_region(self):
        if self.proc is None and getattr(self, "bracket", False):
            self.snapshot()
            return self._summary(dict(note="timed region shorter than 2 s: bracketed by one sample before and one after "
                                           "(polling nvidia-smi perturbs launch-bound steps)"))
        return self.stop()

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def stop(self):
        if not self.proc:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        return self._summary({})

    def _summary(self, extra):
        sm, mx, mem, power, reasons = [], [], [], [], set()
        lines = [(t, l) for t, l in self.lines if t >= self.begin]
        nearest = not lines and bool(self.lines)
        if nearest:   # the timed region was shorter than the sampling period: the sample right before it stands in
            lines = self.lines[-1:]
        for stamp, line in lines:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            try:
                power.append(float(f[3]))
                mem.append(float(f[9]))
            except (ValueError, IndexError):
                pass
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        busy = [c for c in sm if c > 0]
        return dict(sm_mhz=float(np.median(busy)) if busy else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=0 if nearest else len(sm), mem_mhz=float(np.median(mem)) if mem else None,
                    power_w=float(np.median(power)) if power else None,
                    **(dict(note="timed region shorter than the 200 ms sampling period: the last sample before it is reported") if nearest else {}),
                    **extra)


def sustained_copy(device, seconds=4.0):
    """Device-to-device copy bandwidth (read + write bytes) sustained for `seconds` -- the same measurement as
    MEASURED_PEAKS.json's burst figure (torch b.copy_(a), 1 Gi bf16 elements) but held as long as a benchmark step, so it
    runs at whatever clocks the board settles to under its power cap.  Explains the gap between the step kernel's share of
    the burst peak in a short capture (ncu: 99 %) and in the long timed region; the roofline `peak` stays the burst figure."""
    try:
        import torch
        dev = torch.device("cuda", device)
        a = torch.empty(1 << 30, dtype=torch.bfloat16, device=dev)
        b = torch.empty_like(a)
        a.zero_()
        for _ in range(3):
            b.copy_(a)
        torch.cuda.synchronize(dev)
        sampler = ClockSampler(device)
        sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps, total_ms, total_reps = 200, 0.0, 0
        while total_ms < seconds * 1e3:
            e0.record()
            for _ in range(reps):
                b.copy_(a)
            e1.record()
            torch.cuda.synchronize(dev)
            total_ms += e0.elapsed_time(e1)
            total_reps += reps
        clocks = sampler.stop()
        gbs = 2.0 * a.numel() * 2 * total_reps / (total_ms * 1e-3) / 1e9
        del a, b
        torch.cuda.empty_cache()
        return dict(gbs=gbs, seconds=total_ms * 1e-3, sm_mhz=clocks.get("sm_mhz"), mem_mhz=clocks.get("mem_mhz"),
                    power_w=clocks.get("power_w"), reasons=clocks.get("reasons"),
                    how="torch b.copy_(a) over 1 Gi bf16 elements back to back, CUDA events (MEASURED_PEAKS.json's copy, sustained)")
    except Exception as e:   # explanatory extra: never fails the benchmark
        return dict(gbs=None, error=str(e))


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(workload, vectors_per_pass):
    """dram bytes per launch of the step kernel from a committed ncu capture of THIS workload at THIS number of vectors
    per pass (profiles/traffic.json: {workload: {"<R>": bytes}}), else None -- a capture of another launch shape is not
    a measurement of this run"""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        entry = json.load(open(path)).get(workload)
        if isinstance(entry, dict):
            value = entry.get(str(int(vectors_per_pass)))
            return float(value) if value is not None else None
    except Exception:
        pass
    return None


TOLERANCE = {"float32": 1e-5, "complex64": 1e-5, "float64": 1e-11, "complex128": 1e-11}   # north_star, relative to max |mu|


def parity_moments(w, nnz):
    """Number of leading moments of the timed run which are compared with the oracle: all of them when that is cheap,
    otherwise the first 10 (the 4k + 2 number next to 8)"""
    return w["moments"] if nnz * w["moments"] * w["vectors"] < 2e10 else 10


class ParityOracle:
    """The `hp` oracle (oracle/: CPU restatement of the reference, f64 accumulation) on the same Hamiltonian and the
    same MT19937 starters, run on the host cores in a background thread while the GPU loop is being timed."""

    def __init__(self, model, w, threads=0):
        from oracle.oracle import OracleKPM, hardware_threads
        self.count = parity_moments(w, model.hamiltonian.nnz)
        self.vectors = w["vectors"]
        self.threads = threads or max(1, hardware_threads() - 1)
        self.result, self.error, self.seconds = None, None, 0.0
        self._make = lambda: OracleKPM(model.hamiltonian, energy_range=w["energy_range"], num_threads=self.threads, hp=True)
        self._thread = threading.Thread(target=self._run, daemon=True)
        self._thread.start()

    def _run(self):
        t0 = time.perf_counter()
        try:
            self.result = self._make().dos_moments(self.count, self.vectors)
        except Exception as e:   # reported by check()
            self.error = e
        self.seconds = time.perf_counter() - t0

    def check(self, moments, dtype):
        """max |gpu - oracle| / max |oracle| over the compared moments, and the verdict"""
        self._thread.join()
        if self.error is not None:
            raise self.error
        got = np.asarray(moments)[:self.count]
        err = float(np.abs(got - self.result).max() / np.abs(self.result).max())
        tol = TOLERANCE[np.dtype(dtype).name]
        return dict(parity_max_rel=err, tolerance=tol, moments_compared=int(self.count), vectors=int(self.vectors),
                    oracle="hp (f64 accumulation), identical MT19937 starters, {} host threads, {:.1f} s".format(
                        self.threads, self.seconds), passed=bool(err <= tol))


def cpu_sample(model, w, threads):
    """Reference-shaped CPU run (oracle port, native accumulation) on a bounded sample of the same Hamiltonian.

    One wave of thread-pool jobs (SIMD batch x threads vectors, or all vectors if fewer) runs a shortened recursion on
    the FULL Hamiltonian with a probe around `probe_steps` recursion steps in its middle: that gives the cost per step
    with every thread busy (the asymptotic rate) and, by difference, the fixed cost per wave (starter, allocation and
    first touch of the vector blocks, r1).  Both are exactly linear in the reference, so the whole job is
    waves * (fixed + (M / 2 - 1) * t_step) and `value` is nnz * M * R divided by that.
    """
    from oracle.oracle import OracleKPM, hardware_threads
    threads = threads or hardware_threads()
    M, R = w["moments"], w["vectors"]
    batch = 32 // np.dtype(w["dtype"]).itemsize           # the reference's SIMD batch (simd.hpp:42-44)
    vectors = min(R, max(batch, threads * batch)) if R > 1 else 1
    jobs = max(1, vectors // batch + (vectors % batch))   # Compute.cpp:52-63: batches, then single vectors
    nnz = model.hamiltonian.nnz
    max_steps = max(2, (M // 2 - 4) // 2 * 2)
    probe_steps = int(min(max_steps, max(2, round(1.2e11 / (2.0 * nnz * vectors) / 2) * 2)))
    n1, n2 = 2, 2 + probe_steps
    m_run = 2 * n2 + 2
    ref = OracleKPM(model.hamiltonian, energy_range=w["energy_range"], num_threads=threads, hp=False)
    total, probe, reports = ref.time_dos_probe(m_run, vectors, threads, n1, n2, cheap_starter=True)
    t_step = probe / probe_steps
    fixed = max(0.0, total - n2 * t_step)                 # the run has n2 steps (n = 2 .. n2 + 1)
    waves = -(-R // vectors)
    job_seconds = waves * (fixed + (M // 2 - 1) * t_step)
    value = nnz * M * R / job_seconds
    slope = 2.0 * nnz * vectors / t_step
    return dict(value=value, unit=UNIT, cores=min(threads, jobs), kind="port",
                asymptotic_value=slope, fixed_seconds_per_wave=fixed, seconds_per_step=t_step, sample_seconds=total,
                extrapolated_job_seconds=job_seconds,
                sample="{} moments x {} vectors on the full Hamiltonian ({:.1f} s): oracle C++ port of the reference "
                       "CPU path (ELL, interleaved diagonal recursion, 32-byte SIMD batches of {}, thread pool over "
                       "batches, native accumulation); {} recursion steps timed with all {} jobs running: {:.3f} s per "
                       "step = {:.3e} units/s asymptotically; fixed cost per wave {:.1f} s (with a trivial +-1 fill "
                       "instead of the reference's mutex-serialised MT19937 starter, i.e. a lower bound); value = the "
                       "whole job ({} moments x {} vectors, {} wave(s)) extrapolated linearly".format(
                           m_run, vectors, total, batch, probe_steps, reports, t_step, slope, fixed, M, R, waves)), total


def run_reference(args, w):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    model = build_model(w)
    # the arm is a bounded CPU sample: at most two of them whatever --steps / --warmup say (the first one is the
    # warm-up when there are two), so that the arm ends within a few minutes
    samples = max(1, min(2, args.warmup + args.steps))
    base, seconds = None, 0.0
    for _ in range(samples):
        base, seconds = cpu_sample(model, w, 0)
    value = base["value"]
    out = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
               ms_per_step=seconds * 1e3, higher_is_better=True, scaling="strong",
               vs_baseline=None, dtype=w["dtype"], data="synthetic", impl="reference",
               config=dict(workload=w["text"], samples_run=samples,
                           note="bounded CPU sample on the full Hamiltonian, extrapolated linearly in moments and vectors "
                                "(see cpu_baseline.sample); ms_per_step is the duration of one sample"),
               cpu_baseline=base,
               e2e=dict(value=value, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    emit(out)


def run_quantity(args, name, w):
    """configs[2] / configs[3] through the public API: one step = one API call with host buffers in and curves out.
    Parity first (a bounded piece of the same calculation against the hp oracle), then W warm-up and K timed calls."""
    import pybinding_b200 as pb
    from oracle.oracle import OracleKPM, hardware_threads
    rank = int(os.environ.get("RANK", "0"))
    if int(os.environ.get("WORLD_SIZE", "1")) > 1 and rank != 0:
        return   # single-GPU lines
    model = build_model(w)
    n, nnz = model.hamiltonian.shape[0], model.hamiltonian.nnz
    tol = TOLERANCE[w["dtype"]]
    sampler = ClockSampler(0)
    q = w["quantity"]
    if q in ("ldos", "greens"):
        kpm = pb.kpm(model, energy_range=w["energy_range"], silent=True)
        ref = OracleKPM(model.hamiltonian, energy_range=w["energy_range"], hp=True)
        g = int(round(np.sqrt(w["sites"])))
        energy = np.linspace(-1, 1, 500)
        M = kpm.kernel.required_num_moments(w["broadening"] / kpm.scaling_factors[0])
        if q == "ldos":
            xs = np.linspace(-0.4 * w["size"], 0.4 * w["size"], g)
            sites = [model.system.find_nearest([x, y]) for x in xs for y in xs][:w["sites"]]
            err = float(np.abs(kpm.impl.moments_ldos(M, sites[:1]) - ref.ldos_moments(M, sites[:1])).max() / 0.5)
            call = lambda: kpm.impl._ldos_indices(sites, energy, w["broadening"])
        else:
            xs = np.linspace(-20.0, 20.0, g)
            sites = [model.system.find_nearest([x, y]) for x in xs for y in xs][:w["sites"]]
            row = model.system.find_nearest([0.0, 0.0])
            expected = ref.greens_moments(M, row, sites[:4])
            err = float(np.abs(kpm.impl.moments_greens(M, row, sites[:4]) - expected).max() / np.abs(expected).max())
            call = lambda: kpm.calc_greens(row, sites, energy, w["broadening"])
        parity = dict(parity_max_rel=err, tolerance=tol, passed=bool(err <= tol),
                      what="raw moments ({} moments) of {} against the hp oracle".format(M, "one site" if q == "ldos" else "four destinations"))
        units = lambda st: 2.0 * st.opt_nnz * st.multiplier if q == "ldos" else float(st.opt_nnz) * len(sites)
        metric = "KPM nnz*moments*{}/s on light-cone rows (graphene {})".format("sites" if q == "ldos" else "destinations", q.upper())
    else:
        kpm = pb.kpm(model, energy_range=w["energy_range"], kernel=pb.lorentz_kernel(), silent=True)
        ref = OracleKPM(model.hamiltonian, energy_range=w["energy_range"], kernel="lorentz", hp=True, num_threads=hardware_threads())
        a, _ = kpm.scaling_factors
        broadening = a * 4.0 / 512.5     # Lorentz lambda = 4 -> 514 moments
        mu = np.linspace(-1, 1, 101)
        x, y = model.system.x, model.system.y
        expected = ref.kubo_moments(66, x, y, 2)
        err = float(np.abs(kpm.impl.moments_kubo(66, x, y, 2) - expected).max() / np.abs(expected).max())
        parity = dict(parity_max_rel=err, tolerance=5 * tol, passed=bool(err <= 5 * tol),
                      what="66 x 66 Kubo-Bastin moment matrix (xy, 2 vectors) against the hp oracle")
        M = 514
        gemm = dict(ms=0.0, flops=0.0, step_ms=0.0)

        def call():
            out = []
            for direction in ("xx", "xy"):
                out.append(kpm.calc_conductivity(mu, broadening, 300.0, direction, num_random=w["vectors"], num_points=w["points"]).data)
                st = kpm.stats
                assert st.num_moments == M, st.num_moments
                gemm["ms"] += st.gemm_ms; gemm["flops"] += st.gemm_flops; gemm["step_ms"] += st.step_ms
            return out
        units = lambda st: 2.0 * float(nnz) * M * M * w["vectors"] / M   # per direction pair: 2 x M recursion steps over nnz, per vector
        metric = "KPM Kubo-Bastin sigma_xx + sigma_xy calls/s (graphene 200 nm, 514 moments, 4 vectors)"
    if not parity["passed"]:
        sys.stderr.write("bench.py: PARITY FAILED for {}: {}\n".format(name, parity))
        emit(dict(metric=metric, value=None, error="parity failed", parity=parity))
        sys.exit(1)
    t_warm = 0.0
    for _ in range(max(args.warmup, 1)):
        t0 = time.perf_counter()
        call()
        t_warm = time.perf_counter() - t0
    if q == "conductivity":
        gemm.update(ms=0.0, flops=0.0, step_ms=0.0)
    times, dev_ms, launches = [], [], 0
    sampler.begin_region(t_warm * args.steps)
    if getattr(sampler, "after_snapshot", False):
        call()
        if q == "conductivity":
            gemm.update(ms=0.0, flops=0.0, step_ms=0.0)
    for _ in range(args.steps):
        t0 = time.perf_counter()
        out = call()
        times.append(time.perf_counter() - t0)
        st = kpm.stats
        dev_ms.append(st.moments_device_ms)
        launches += st.kernel_launches
    clocks = sampler.end_region()
    st = kpm.stats
    t = float(np.mean(times))
    peak, peak_src = measured_peak()
    if q == "conductivity":
        tf = gemm["flops"] / (gemm["ms"] * 1e-3) / 1e12 if gemm["ms"] else 0.0
        roofline = dict(bound="tensor", achieved=tf, peak=FP64_TENSOR_PEAK, unit="TFLOP/s", frac=tf / FP64_TENSOR_PEAK, traffic=None,
                        peak_source="FP64 DMMA.8x8x4 peak measured with tools/probe/dmma_probe.cu (profiles/r01_fp64_mma_probe.jsonl)",
                        kernel="kubo_gemm_kernel (mu += L R^H over sites x lanes, mma.sync f64)",
                        gemm_ms_per_call=gemm["ms"] / args.steps, recursion_ms_per_call=gemm["step_ms"] / args.steps)
        value, unit = 1.0 / t, "calls/s"
    else:
        gbs = st.step_bytes / (st.step_ms * 1e-3) / 1e9 if st.step_ms else 0.0
        roofline = dict(bound="hbm", achieved=gbs, peak=peak, unit="GB/s", frac=gbs / peak, traffic=None, peak_source=peak_src,
                        kernel="cone_group_step_kernel / cheb_step on light-cone sub-systems (working sets near the L2 size: "
                               "reported against HBM for reference, see DESIGN.md)",
                        removed_by_light_cone=1.0 - st.opt_nnz / max(st.nnz, 1))
        value, unit = units(st) / t, "nnz*moments*units/s"
    emit(dict(metric=metric, value=value, unit=unit, n_gpus=1, steps=args.steps, warmup=args.warmup, ms_per_step=t * 1e3,
              higher_is_better=True, scaling="strong", vs_baseline=None, dtype=w["dtype"], data="synthetic",
              config=dict(workload=w["text"], sites=int(n), nnz=int(nnz), moments=int(M),
                          moments_device_ms=float(np.mean(dev_ms)), step_seconds=[round(x, 4) for x in times],
                          timing="wall clock around the public API call (host buffers in, curves out)"),
              roofline=roofline, cpu_baseline=None,
              e2e=dict(value=value, unit=unit, h2d_bytes_per_step=int(st.h2d_bytes), d2h_bytes_per_step=int(st.d2h_bytes), seconds=t),
              gpu_launches=int(launches), clocks=clocks, parity=parity, parity_max_rel=parity["parity_max_rel"]))


_JSON_FD = None


def claim_stdout():
    """stdout carries the ONE JSON line and nothing else: native libraries (NCCL prints its version banner there) get
    stderr for the rest of the run, the JSON line is written to the original descriptor by emit()."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit(record):
    line = (json.dumps(record) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(line.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, line)


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS) + sorted(QUANTITY_WORKLOADS))
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline sample")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle comparison (sweeps only: the line says so)")
    ap.add_argument("--max-batch", type=int, default=0)
    args = ap.parse_args()
    if args.workload in QUANTITY_WORKLOADS:
        if args.impl == "reference":
            emit(dict(impl="reference", unavailable="the reference arm is defined for the DOS workloads (BASELINE.json metric)"))
            return
        return run_quantity(args, args.workload, QUANTITY_WORKLOADS[args.workload])
    w = WORKLOADS[args.workload]
    if args.impl == "reference":
        return run_reference(args, w)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    import pybinding_b200 as pb
    from pybinding_b200 import multigpu

    model = build_model(w)
    nnz, n = model.hamiltonian.nnz, model.hamiltonian.shape[0]
    M, R = w["moments"], w["vectors"]

    def make_kpm():
        k = pb.kpm(model, energy_range=w["energy_range"], silent=True, device=local_rank, max_batch=args.max_batch)
        if world > 1:
            multigpu.attach(k, dist, rank, world, "cuda")
        return k

    def barrier():
        if world > 1:
            import torch
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # parity gate: the oracle computes the leading moments of the very run that is timed (same H, same starters) on the
    # host cores meanwhile; the line is only printed when they agree
    oracle = ParityOracle(model, w) if (rank == 0 and not args.no_parity) else None
    kpm = make_kpm()
    kpm.impl.scaling_factors  # bounds known (explicit range): nothing to compute
    # ---- device-resident timing: Hamiltonian already in HBM, one step = the whole moments phase ----
    moments = None
    sampler = ClockSampler(local_rank)
    t_warm = 0.0
    for _ in range(args.warmup):
        t0 = time.perf_counter()
        moments = kpm.impl.moments_dos(M, R)
        t_warm = time.perf_counter() - t0
    step_times, wall_times, launches, step_ms, step_bytes, step_launches, starter_ms, bulk_launches = [], [], 0, 0.0, 0.0, 0, 0.0, 0
    res_launches, persist_launches = 0, 0
    barrier()
    sampler.begin_region(max_over_ranks(t_warm) * args.steps if args.warmup else 10.0)
    if getattr(sampler, "after_snapshot", False):
        kpm.impl.moments_dos(M, R)
    for _ in range(args.steps):
        barrier()
        t0 = time.perf_counter()
        moments = kpm.impl.moments_dos(M, R)   # returns after the stream is synchronised (moments on the host)
        barrier()
        wall_times.append(max_over_ranks(time.perf_counter() - t0))
        s = kpm.stats
        # device time of the whole moments phase (CUDA events on the engine's stream), max over ranks
        step_times.append(max_over_ranks(s.moments_device_ms * 1e-3))
        launches += s.kernel_launches
        step_ms += s.step_ms
        step_bytes += s.step_bytes
        step_launches += s.step_launches
        bulk_launches += s.bulk_launches
        res_launches += s.res_launches
        persist_launches += s.persist_launches
        starter_ms += s.starter_ms
        batch = s.batch
    clocks = sampler.end_region()
    t_step = float(np.mean(step_times))
    value = nnz * M * R / t_step
    parity = oracle.check(moments, w["dtype"]) if oracle else None   # joins the oracle thread before the e2e leg

    peak, peak_src = measured_peak()
    if res_launches:
        kernel_name = "cheb_step_res (fused SpMM + moments, x rows of a tile and its halo resident in shared memory, y / matrix records staged by cp.async.bulk)"
    elif bulk_launches:
        kernel_name = "cheb_step_bulk (fused SpMM + moments, operands staged by cp.async.bulk)"
    elif persist_launches:
        kernel_name = "cheb_persist (all steps of the recursion in one cooperative launch, rows in registers)"
    else:
        kernel_name = "cheb_step (fused SpMM + moments)"
    achieved = (step_bytes / step_launches) / (step_ms / step_launches * 1e-3) / 1e9 if step_launches else 0.0
    s_item = np.dtype(w["dtype"]).itemsize
    roofline = dict(bound="hbm", achieved=achieved, peak=peak, unit="GB/s", frac=achieved / peak,
                    traffic=ncu_traffic(args.workload, batch), peak_source=peak_src,
                    kernel=kernel_name, staged_launches=int(bulk_launches), resident_tile_launches=int(res_launches),
                    persistent_launches=int(persist_launches),
                    algorithmic_bytes_per_launch=step_bytes / max(step_launches, 1),
                    launch_ms=step_ms / max(step_launches, 1), vectors_per_pass=batch,
                    bytes_model="rows*[k*(s+4) + 3*R*s], s={}".format(s_item))

    # ---- end to end through the public API with host buffers ----
    e2e = None
    if not args.no_e2e:
        energy = np.linspace(-3, 3, 1000)
        broadening = float(np.float32(np.pi)) * kpm.scaling_factors[0] / (M - 2)   # -> exactly M moments (Jackson)
        del kpm
        times, set_times, h2d, d2h = [], [], 0, 0
        k2 = make_kpm()   # context + NCCL communicator: created once per process, like a user would
        for i in range(1 + max(1, min(args.steps, 2))):
            barrier()
            t0 = time.perf_counter()
            k2.model = model                                 # host CSR -> page-locked mirror + upload, locality ordering (rank 0, broadcast)
            t_set = time.perf_counter() - t0
            dos = k2.calc_dos(energy, broadening, num_random=R)   # device layout build, moments + allreduce + reconstruction, result on the host
            barrier()
            dt = max_over_ranks(time.perf_counter() - t0)
            st = k2.stats
            assert st.num_moments == M, (st.num_moments, M)
            if i > 0:
                times.append(dt)
                set_times.append(max_over_ranks(t_set))
                h2d, d2h = st.h2d_bytes, st.d2h_bytes + dos.data.nbytes
        del k2
        e2e = dict(value=nnz * M * R / float(np.mean(times)), unit=UNIT, h2d_bytes_per_step=int(h2d),
                   d2h_bytes_per_step=int(d2h), seconds=float(np.mean(times)), set_model_seconds=float(np.mean(set_times)),
                   seconds_each=[round(x, 4) for x in times], set_model_seconds_each=[round(x, 4) for x in set_times],
                   note="kpm.model = model (host CSR mirrored into page-locked memory and uploaded, locality ordering on rank 0 + "
                        "broadcast) + calc_dos (scale / relabel / ELL / packing on the device, moments, allreduce, reconstruction, "
                        "D2H) through the public API")

    if world == 1 and n * batch * s_item > 2e9:   # long, bandwidth-bound steps only: what does a plain copy sustain on this board right now?
        roofline["sustained_copy"] = sustained_copy(local_rank)
        if roofline["sustained_copy"].get("gbs"):
            roofline["frac_of_sustained_copy"] = achieved / roofline["sustained_copy"]["gbs"]

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:   # the CPU baseline is reported at N = 1 only
        cpu = cpu_sample(model, w, 0)[0]

    if rank == 0 and parity is not None and not parity["passed"]:
        sys.stderr.write("bench.py: PARITY FAILED -- the first {moments_compared} moments of the timed run differ from the "
                         "oracle by {parity_max_rel:.3e} (tolerance {tolerance:.0e}); no result line is printed\n".format(**parity))
        emit(dict(metric=METRIC, value=None, unit=UNIT, n_gpus=world, error="parity failed", parity=parity))
        if world > 1:
            dist.destroy_process_group()
        sys.exit(1)
    if rank == 0:
        out = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup,
                   ms_per_step=t_step * 1e3, higher_is_better=True, scaling="strong", vs_baseline=None,
                   dtype=w["dtype"], data="synthetic",
                   config=dict(workload=w["text"], sites=int(n), nnz=int(nnz), moments=M, vectors=R,
                               parallelism="vectors sharded over {} rank(s), 1 ncclAllReduce".format(world),
                               l2="inputs larger than L2 ({:.1f} GB of vectors per pass)".format(
                                   2 * n * batch * s_item / 1e9),
                               starter_ms_per_step=starter_ms / max(args.steps, 1),
                               step_seconds=[round(x, 6) for x in step_times],
                               wall_ms_per_step=float(np.mean(wall_times)) * 1e3,
                               timing="CUDA events on the engine stream around the whole moments phase, max over ranks"),
                   roofline=roofline, cpu_baseline=cpu, e2e=e2e, gpu_launches=int(launches), clocks=clocks,
                   parity=parity, parity_max_rel=parity["parity_max_rel"] if parity else None,
                   moment_checksum=float(np.abs(moments).sum()))
        emit(out)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
