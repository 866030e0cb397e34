#!/usr/bin/env python
"""bench.py - headline benchmark of the rosdyn::Chain hot path on B200 (BASELINE.json metric:
"6-DOF regressor+torque samples/sec (fp64)").

One step = one pass of the hot path over one batch of synthetic (q,Dq,DDq) samples of the UR10-like 6-DOF chain
(C6): getRegressor + getJointTorque for every sample.

  --workload materialise : Phi (6x70, 420 SoA planes) + tau (6 planes) written to HBM        (HBM roofline)
  --workload gram        : Phi and tau consumed in-kernel by the normal equations Phi^T Phi / Phi^T tau
                           (FP64 DMMA roofline); N>1 adds one NCCL all-reduce of the (70^2+70+1) partials per step

`value`  : device-resident inputs, CUDA-event timed, max over ranks.
`e2e`    : the same metric through the C-ABI host-buffer entry point (pinned host inputs, H2D/D2H inside the timed region).
`--impl reference` : the reference's CPU Chain loop on all host threads.  Two builds of it exist under oracle/: the plain-C restatement
                     (kind "port") and oracle/_ref, the reference's own headers compiled against stand-in Eigen/urdf/ros headers (kind
                     "reference"; its dense arithmetic is the stand-in's plain loops, so it is the slower of the two).  Both are timed; the
                     line's value is the FASTER one (the conservative baseline), the other is reported beside it.
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "6-DOF regressor+torque samples/sec (fp64)"
UNIT = "samples/s"
SEED = 0x5EED0000 + 1


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=os.environ.get("RDB_BENCH_WORKLOAD", "gram"), choices=["materialise", "gram"])
    ap.add_argument("--chain", default="c6")
    ap.add_argument("--samples", type=int, default=0, help="samples per step per GPU (0 = workload default)")
    ap.add_argument("--e2e-samples", type=int, default=0)
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target length of the bounded CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """Samples SM clock / throttle reasons during the timed region (NVML; nvidia-smi fallback)."""

    def __init__(self, index: int):
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop = threading.Event()
        self._thr = None
        self._nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nvml = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self._nvml = None

    def _loop(self):
        while not self._stop.is_set():
            try:
                if self._nvml:
                    p = self._nvml
                    self.samples.append(float(p.nvmlDeviceGetClockInfo(self._h, p.NVML_CLOCK_SM)))
                    r = p.nvmlDeviceGetCurrentClocksEventReasons(self._h) if hasattr(p, "nvmlDeviceGetCurrentClocksEventReasons") \
                        else p.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                    table = {"sw_power_cap": 0x4, "hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40,
                             "hw_power_brake_slowdown": 0x80}
                    for k, bit in table.items():
                        if r & bit:
                            self.reasons.add(k)
                else:
                    out = subprocess.run(["nvidia-smi", f"--id={self.index}", "--query-gpu=clocks.sm,clocks.max.sm", "--format=csv,noheader,nounits"],
                                         capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                    self.samples.append(float(out[0]))
                    self.max_mhz = float(out[1])
            except Exception:
                pass
            self._stop.wait(0.02)

    def __enter__(self):
        self._thr = threading.Thread(target=self._loop, daemon=True)
        self._thr.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._thr.join(timeout=2)

    def summary(self):
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def ncu_traffic_per_sample(kernel: str):
    """dram__bytes_read.sum + dram__bytes_write.sum per sample of the named kernel, from the committed `ncu --set full` capture
    (profiles/ncu_traffic.json, written by tools/ncu_summary.py); None when there is no capture for it."""
    try:
        return float(json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))[kernel]["dram_bytes_per_sample"])
    except Exception:
        return None


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)), "measured"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0}, "fallback"


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_run(chain_name: str, n: int, threads: int, reps: int = 1, kind="port"):
    """getJointTorque + getRegressor per sample on the CPU, all samples resident in host memory.
    kind "port": the plain-C restatement (oracle/rosdyn_oracle.c); "reference": the reference's own rosdyn::Chain (oracle/_ref)."""
    from oracle.oracle import OracleChain, fill_uniform
    from rosdyn_b200 import fixtures
    d = fixtures.by_name(chain_name)
    oc = OracleChain(d, fast="ref" if kind == "reference" else False)
    q, dq, ddq = (fill_uniform(d.n_inputs, n, SEED, s) for s in range(3))
    best = None
    for _ in range(reps):
        t = time.perf_counter()
        oc.regressor_torque(q, dq, ddq, nthreads=threads, store=False)
        dt = time.perf_counter() - t
        best = dt if best is None else min(best, dt)
    return n / best, best


def host_threads() -> int:
    """All host threads this process may use (torchrun exports OMP_NUM_THREADS=1, so do not ask OpenMP)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def reference_build_rate(chain_name: str, seconds: float, threads: int):
    """The reference's own headers (oracle/_ref) on the same workload; None when that build is absent."""
    try:
        from oracle import oracle
        if not oracle.have_ref():
            return None
        rate, _ = cpu_run(chain_name, 2000 * threads, threads, kind="reference")
        n = int(max(2000 * threads, min(rate * seconds, 5_000_000)))
        rate, dt = cpu_run(chain_name, n, threads, kind="reference")
        return {"value": rate, "unit": UNIT, "cores": threads, "kind": "reference",
                "sample": f"{n} samples, rosdyn::Chain::getJointTorque + getRegressor of the reference's own headers compiled against the "
                          f"stand-in Eigen of oracle/shim (one Chain::clone() per thread, OpenMP {threads} threads, {dt:.1f} s)"}
    except Exception as e:  # the checker build must never take the bench down
        return {"unavailable": repr(e)}


def cpu_baseline(chain_name: str, seconds: float):
    threads = host_threads()
    rate, _ = cpu_run(chain_name, 20000 * max(1, threads // 4), threads)       # calibration
    n = int(max(50_000, min(rate * seconds, 50_000_000)))
    rate, dt = cpu_run(chain_name, n, threads)
    out = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
           "sample": f"{n} samples of the same workload (getJointTorque + getRegressor per sample, oracle/rosdyn_oracle.c, "
                     f"OpenMP {threads} threads, {dt:.1f} s)"}
    ref = reference_build_rate(chain_name, min(4.0, seconds), threads)
    if ref is not None:
        out["reference_build"] = ref   # slower than the port (plain-loop stand-in for Eigen): the port stays the baseline
    return out


def run_reference(args):
    """--impl reference: the reference's CPU Chain loop on all host threads (the faster of the restatement and oracle/_ref)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = host_threads()
    rate, _ = cpu_run(args.chain, 20000 * max(1, threads // 4), threads)
    n = int(max(20_000, min(rate * 4.0, 20_000_000)))                           # ~4 s per step
    from oracle.oracle import OracleChain, fill_uniform
    from rosdyn_b200 import fixtures
    d = fixtures.by_name(args.chain)
    oc = OracleChain(d)
    q, dq, ddq = (fill_uniform(d.n_inputs, n, SEED, s) for s in range(3))
    for _ in range(min(args.warmup, 1)):
        oc.regressor_torque(q, dq, ddq, nthreads=threads, store=False)
    t = time.perf_counter()
    for _ in range(args.steps):
        oc.regressor_torque(q, dq, ddq, nthreads=threads, store=False)
    dt = time.perf_counter() - t
    v = n * args.steps / dt
    ref = reference_build_rate(args.chain, 4.0, threads)
    kind, note = "port", "oracle/rosdyn_oracle.c (plain-C restatement of primitives_impl.h), OpenMP"
    if ref and ref.get("value", 0.0) > v:   # report the faster CPU build: the conservative baseline
        v, kind, note = ref["value"], "reference", ref["sample"]
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": f"{d.name}: getRegressor (6x70) + getJointTorque per sample (the GPU arm's workload), evaluated by the CPU "
                               f"restatement of the reference's Chain loop; bounded sample of {n} samples/step",
                   "chain": d.name, "samples_per_step": n, "mode": args.workload},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": kind,
                         "sample": f"{n} samples/step x {args.steps} steps, {note}", "reference_build": ref},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


def bind_to_gpu_numa_node(index: int):
    """Pin this rank to the CPUs NVML reports as local to its GPU, so that the pinned host buffers of the e2e leg are first-touched on the
    GPU's NUMA node (8 ranks reading from one socket do not reach 8 x PCIe bandwidth).  Returns the previous affinity (restored before
    the CPU baseline leg) or None when NVML / the syscall is unavailable."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1}
        old = os.sched_getaffinity(0)
        cpus &= old
        if cpus:
            os.sched_setaffinity(0, cpus)
            return old
    except Exception:
        pass
    return None


# ------------------------------------------------------------------------------------------------ GPU arm
def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist
    from rosdyn_b200 import fixtures
    from rosdyn_b200.chain import Chain, fill_uniform, fp64_peak, kernel_launch_count
    from rosdyn_b200.sharding import allreduce_normal_equations

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    old_affinity = bind_to_gpu_numa_node(local) if world > 1 else None
    # rank 0 prints ONE JSON line on stdout: everything else that lands on file descriptor 1 from here on (NCCL prints its version banner
    # there) is sent to stderr, and the JSON line goes to the saved descriptor at the end
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    d = fixtures.by_name(args.chain)
    ch = Chain(d)
    n_in, P = d.n_inputs, 10 * d.n_joints
    gram = args.workload == "gram"
    S = args.samples or (8_000_000 if gram else 4_000_000)        # per GPU per step; inputs (and outputs) >> 126 MB L2
    # weak scaling: every rank owns its own shard of S samples (disjoint sample indices via the seed offset)
    q, dq, ddq = (fill_uniform(n_in, S, SEED + 1000003 * rank, s, device=dev) for s in range(3))
    if gram:
        G = torch.zeros((P, P), dtype=torch.float64, device=dev)
        b = torch.zeros((P,), dtype=torch.float64, device=dev)
        tt = torch.zeros((1,), dtype=torch.float64, device=dev)
        flat = torch.zeros((P * P + P + 1,), dtype=torch.float64, device=dev)
    else:
        phi = torch.empty((P * n_in, S), dtype=torch.float64, device=dev)
        tau = torch.empty((n_in, S), dtype=torch.float64, device=dev)

    import ctypes
    from rosdyn_b200._lib import CSamples, check, load
    lib = load()
    smp = CSamples(S, S, q.data_ptr(), dq.data_ptr(), ddq.data_ptr(), None)

    def step():
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        if gram:
            check(lib.rdb_regressor_gram_batch(ch._h, ctypes.byref(smp), None, G.data_ptr(), b.data_ptr(), tt.data_ptr(), 0, st))
            if world > 1:   # the one exchange step of the path: sum the small normal-equation partials over NVLink
                allreduce_normal_equations(G, b, tt, flat=flat)
        else:
            check(lib.rdb_regressor_batch(ch._h, ctypes.byref(smp), phi.data_ptr(), tau.data_ptr(), S, st))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    l0 = kernel_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        barrier()
        e0.record()
        for _ in range(args.steps):
            step()
        e1.record()
        barrier()
    ms = e0.elapsed_time(e1)
    launches = kernel_launch_count() - l0
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t[0])
    value = world * S * args.steps / (ms * 1e-3)

    # ------------------------------------------------------------------ e2e through the host-buffer C-ABI entry
    e2e = None
    if not args.no_e2e:
        Se = args.e2e_samples or (4_000_000 if gram else 500_000)
        hq, hdq, hddq = (torch.empty((n_in, Se), dtype=torch.float64).pin_memory() for _ in range(3))
        for k, h in enumerate((hq, hdq, hddq)):
            lib.rdb_fill_uniform_host(h.data_ptr(), n_in, Se, Se, SEED + 1000003 * rank, k)
        hs = CSamples(Se, Se, hq.data_ptr(), hdq.data_ptr(), hddq.data_ptr(), None)
        if gram:
            hG = torch.empty((P, P), dtype=torch.float64).pin_memory()
            hb = torch.empty((P,), dtype=torch.float64).pin_memory()
            ht = torch.empty((1,), dtype=torch.float64).pin_memory()
            hflat = torch.empty((P * P + P + 1,), dtype=torch.float64)

            def e2e_step():
                check(lib.rdb_regressor_gram_batch_host(ch._h, ctypes.byref(hs), None, hG.data_ptr(), hb.data_ptr(), ht.data_ptr(), 0))
                if world > 1:
                    hflat[:P * P] = hG.reshape(-1)
                    hflat[P * P:P * P + P] = hb
                    hflat[P * P + P:] = ht
                    f = hflat.to(dev)
                    dist.all_reduce(f)
                    f.cpu()
            h2d, d2h = 3 * n_in * Se * 8, (P * P + P + 1) * 8
        else:
            hphi = torch.empty((P * n_in, Se), dtype=torch.float64).pin_memory()
            htau = torch.empty((n_in, Se), dtype=torch.float64).pin_memory()

            def e2e_step():
                check(lib.rdb_regressor_batch_host(ch._h, ctypes.byref(hs), hphi.data_ptr(), htau.data_ptr(), Se))
            h2d, d2h = 3 * n_in * Se * 8, (P * n_in + n_in) * Se * 8
        for _ in range(2):
            e2e_step()
        ke = max(2, min(args.steps, 5))
        barrier()
        t0 = time.perf_counter()
        for _ in range(ke):
            e2e_step()
        barrier()
        te = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e = {"value": world * Se * ke / float(te[0]), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "samples_per_step_per_gpu": Se, "steps": ke,
               "api": "rdb_regressor_gram_batch_host" if gram else "rdb_regressor_batch_host"}

    # ------------------------------------------------------------------ roofline of the dominant kernel
    peaks, src = measured_peaks()
    if gram:
        # algorithmic FLOPs per sample, BLAS SYRK+GEMV convention: n_act*P*(P+1) + 2*n_act*P
        flop = n_in * P * (P + 1) + 2 * n_in * P
        peak, peaks64 = None, None
        if rank == 0:
            peaks64 = {"dmma_m8n8k4": fp64_peak("dmma", 3), "dfma": fp64_peak("dfma", 3), "dmma_and_dfma_interleaved": fp64_peak("mixed", 3)}
            peak = max(peaks64["dmma_m8n8k4"], peaks64["dfma"])
        ach = S * flop / (ms * 1e-3 / args.steps) / 1e12
        tps = ncu_traffic_per_sample(f"gram_fused_kernel<{d.n_joints}>:{args.chain}")
        # what the kernel actually executes: joints that never move (fixed / not an input) are folded out of the chain (fold.cpp:
        # fold_chain), and of the (P'+1)x(P'+1) augmented Gram matrix only the upper-triangular 8x8 tiles right of each row's first
        # non-zero column are multiplied (one DMMA m8n8k4 = 512 flop per tile per 4 samples)
        K = sum(1 for j in d.joints if j.input_index >= 0)
        # the row of joint j spans the first ceil((1 + 10 (K - j)) / 8) tiles (the kernel orders the columns tau, last link ... first link)
        dmma4 = sum(((1 + 10 * (K - j) + 7) // 8) * ((1 + 10 * (K - j) + 7) // 8 + 1) // 2 for j in range(K))
        ex = dmma4 * 512 / 4
        sps = S / (ms * 1e-3 / args.steps)
        roof = {"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": (ach / peak) if peak else None,
                "traffic": (tps * S) if tps else None,
                "peak_source": "own FP64 micro-benchmark on this GPU (max of DMMA m8n8k4 and DFMA); MEASURED_PEAKS.json has no FP64 figure",
                "fp64_peaks_tflops": peaks64, "flop_per_sample": flop,
                "executed": {"moving_joints": K, "dmma_per_4_samples": dmma4, "dmma_flop_per_sample": ex, "dmma_tflops": sps * ex / 1e12,
                             "dmma_frac_of_peak": (sps * ex / 1e12 / peak) if peak else None,
                             "note": "achieved/frac use the ALGORITHMIC flops of the reference's dense n_act x 10nJ regressor (SYRK+GEMV convention, "
                                     "SURVEY.md 8d); structural zeros and rigidly attached links are not multiplied, so frac can exceed 1. "
                                     "The regressor generation (~5 kflop/sample of DFMA) shares the same FP64 datapath and is not counted here."},
                "kernel": f"gram_fused_kernel<{K},slots> on the folded chain (regressor generation + DMMA normal equations)"}
    else:
        bytes_per_sample = 8 * (3 * n_in + P * n_in + n_in)          # 3552 B for C6
        ach = S * bytes_per_sample / (ms * 1e-3 / args.steps) / 1e9
        peak = float(peaks["hbm_gbs"])
        tps = ncu_traffic_per_sample(f"dyn_kernel<{d.n_joints},3>:{args.chain}")
        roof = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": (tps * S) if tps else None,
                "peak_source": f"MEASURED_PEAKS.json hbm_gbs ({src})", "bytes_per_sample": bytes_per_sample, "kernel": f"dyn_kernel<{d.n_joints},3> (regressor+torque)"}

    if old_affinity is not None:
        os.sched_setaffinity(0, old_affinity)
    if rank == 0:
        cpu = None if args.no_cpu_baseline else cpu_baseline(args.chain, args.cpu_seconds)
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": (f"{d.name}: getRegressor ({n_in}x{P}) + getJointTorque per sample, "
                                    + ("fused into Phi^T Phi / Phi^T tau normal equations" if gram else f"materialised as {P * n_in + n_in} SoA planes in HBM")),
                       "chain": d.name, "samples_per_step_per_gpu": S, "mode": args.workload,
                       "l2": "inputs (and outputs) per step are far larger than the 126 MB L2; no flush needed",
                       "parallelism": f"{world} x independent sample shards" + (" + NCCL all-reduce of the partials" if gram and world > 1 else "")},
            "e2e": e2e, "gpu_launches": int(launches), "roofline": roof, "cpu_baseline": cpu, "clocks": clk.summary(),
        }
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(out) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
