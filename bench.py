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
DEFAULT_WORKLOAD = "graphene_1000nm_c64_dos"
METRIC = "KPM nnz*moments*vectors/s (graphene DOS)"
UNIT = "nnz*moments*vectors/s"


def build_model(w):
    import pybinding_b200 as pb
    if w["kind"] == "graphene":
        return pb.graphene_rectangle(w["size"], magnetic_field=w["field"], dtype=np.dtype(w["dtype"]))
    return pb.cubic_anderson(w["size"], disorder=4.0, seed=0, dtype=np.dtype(w["dtype"]))


class ClockSampler:
    """Samples SM clocks / throttle reasons with nvidia-smi while the timed region runs"""
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.QUERY,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for line in self.lines:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        busy = [c for c in sm if c > 0]
        return dict(sm_mhz=float(np.median(busy)) if busy else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm))


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(workload):
    """dram bytes per launch of the step kernel from the committed ncu capture, or None"""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(path):
        try:
            return json.load(open(path)).get(workload)
        except Exception:
            return None
    return None


def cpu_sample(model, w, threads, target_seconds=15.0):
    """Reference-shaped CPU run (oracle port) on a bounded sample of the same Hamiltonian.

    The reference's cost is exactly linear in moments and vectors, so the sample keeps the full Hamiltonian and
    one SIMD batch of vectors per thread and shortens the number of moments: a short calibration run sizes it
    for about `target_seconds` of CPU work.
    """
    from oracle.oracle import OracleKPM, hardware_threads
    threads = threads or hardware_threads()
    batch = 32 // np.dtype(w["dtype"]).itemsize           # the reference's SIMD batch (simd.hpp:42-44)
    vectors = min(w["vectors"], max(batch, threads * batch))
    nnz = model.hamiltonian.nnz
    ref = OracleKPM(model.hamiltonian, energy_range=w["energy_range"], num_threads=threads, hp=False)
    moments = 6
    seconds = ref.time_dos(moments, vectors, threads, cheap_starter=True)
    for _ in range(3):  # grow the sample until it is long enough to be dominated by the recursion
        if seconds >= 0.5 * target_seconds or moments >= w["moments"]:
            break
        scale = min(8.0, target_seconds / max(seconds, 1e-9))
        moments = int(min(w["moments"], max(moments + 4, moments * scale)))
        moments = max(10, (moments - 2) // 4 * 4 + 2)
        seconds = ref.time_dos(moments, vectors, threads, cheap_starter=True)
    value = nnz * moments * vectors / seconds
    return dict(value=value, unit=UNIT, cores=threads, kind="port",
                sample="{} moments x {} vectors on the full Hamiltonian ({:.1f} s): oracle C++ port of the reference "
                       "CPU path (ELL, interleaved diagonal recursion, 32-byte SIMD batches of {}, thread pool over "
                       "batches); recursion only -- the reference's mutex-serialised MT19937 starter is replaced by "
                       "a trivial fill for this short sample, which favours the CPU".format(
                           moments, vectors, seconds, batch)), seconds, moments, vectors


def run_reference(args, w):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    model = build_model(w)
    steps = []
    base = None
    for i in range(args.warmup + args.steps):
        base, seconds, moments, vectors = cpu_sample(model, w, 0)
        if i >= args.warmup:
            steps.append((seconds, base["value"]))
    value = float(np.mean([v for _, v in steps]))
    base["value"] = value
    out = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
               ms_per_step=float(np.mean([s for s, _ in steps]) * 1e3), higher_is_better=True, scaling="strong",
               vs_baseline=None, dtype=w["dtype"], data="synthetic", impl="reference",
               config=dict(workload=w["text"], note="CPU sample extrapolates linearly in moments and vectors"),
               cpu_baseline=base,
               e2e=dict(value=value, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    emit(out)


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
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline sample")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--max-batch", type=int, default=0)
    args = ap.parse_args()
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

    kpm = make_kpm()
    kpm.impl.scaling_factors  # bounds known (explicit range): nothing to compute
    # ---- device-resident timing: Hamiltonian already in HBM, one step = the whole moments phase ----
    moments = None
    for _ in range(args.warmup):
        moments = kpm.impl.moments_dos(M, R)
    sampler = ClockSampler(local_rank)
    step_times, wall_times, launches, step_ms, step_bytes, step_launches, starter_ms, bulk_launches = [], [], 0, 0.0, 0.0, 0, 0.0, 0
    barrier()
    sampler.start()
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
        starter_ms += s.starter_ms
        batch = s.batch
    clocks = sampler.stop()
    t_step = float(np.mean(step_times))
    value = nnz * M * R / t_step

    peak, peak_src = measured_peak()
    achieved = (step_bytes / step_launches) / (step_ms / step_launches * 1e-3) / 1e9 if step_launches else 0.0
    s_item = np.dtype(w["dtype"]).itemsize
    roofline = dict(bound="hbm", achieved=achieved, peak=peak, unit="GB/s", frac=achieved / peak,
                    traffic=ncu_traffic(args.workload), peak_source=peak_src,
                    kernel="cheb_step_bulk (fused SpMM + moments, operands staged by cp.async.bulk)" if bulk_launches
                    else "cheb_step (fused SpMM + moments)", staged_launches=int(bulk_launches),
                    algorithmic_bytes_per_launch=step_bytes / max(step_launches, 1),
                    launch_ms=step_ms / max(step_launches, 1), vectors_per_pass=batch,
                    bytes_model="rows*[k*(s+4) + 3*R*s], s={}".format(s_item))

    # ---- end to end through the public API with host buffers ----
    e2e = None
    if not args.no_e2e:
        energy = np.linspace(-3, 3, 1000)
        broadening = float(np.float32(np.pi)) * kpm.scaling_factors[0] / (M - 2)   # -> exactly M moments (Jackson)
        del kpm
        times, h2d, d2h = [], 0, 0
        k2 = make_kpm()   # context + NCCL communicator: created once per process, like a user would
        for i in range(1 + max(1, min(args.steps, 2))):
            barrier()
            t0 = time.perf_counter()
            k2.model = model                                 # host CSR -> scale / order / ELL build -> H2D upload
            dos = k2.calc_dos(energy, broadening, num_random=R)   # moments + allreduce + reconstruction, result on the host
            barrier()
            dt = max_over_ranks(time.perf_counter() - t0)
            st = k2.stats
            assert st.num_moments == M, (st.num_moments, M)
            if i > 0:
                times.append(dt)
                h2d, d2h = st.h2d_bytes, st.d2h_bytes + dos.data.nbytes
        del k2
        e2e = dict(value=nnz * M * R / float(np.mean(times)), unit=UNIT, h2d_bytes_per_step=int(h2d),
                   d2h_bytes_per_step=int(d2h), seconds=float(np.mean(times)),
                   note="kpm.model = model (host CSR: scale, locality ordering, ELL build, H2D upload) + calc_dos (moments, "
                        "allreduce, reconstruction, D2H) through the public API")

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:   # the CPU baseline is reported at N = 1 only
        cpu = cpu_sample(model, w, 0)[0]

    if rank == 0:
        out = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup,
                   ms_per_step=t_step * 1e3, higher_is_better=True, scaling="strong", vs_baseline=None,
                   dtype=w["dtype"], data="synthetic",
                   config=dict(workload=w["text"], sites=int(n), nnz=int(nnz), moments=M, vectors=R,
                               parallelism="vectors sharded over {} rank(s), 1 ncclAllReduce".format(world),
                               l2="inputs larger than L2 ({:.1f} GB of vectors per pass)".format(
                                   2 * n * batch * s_item / 1e9),
                               starter_ms_per_step=starter_ms / max(args.steps, 1),
                               wall_ms_per_step=float(np.mean(wall_times)) * 1e3,
                               timing="CUDA events on the engine stream around the whole moments phase, max over ranks"),
                   roofline=roofline, cpu_baseline=cpu, e2e=e2e, gpu_launches=int(launches), clocks=clocks,
                   moment_checksum=float(np.abs(moments).sum()))
        emit(out)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
