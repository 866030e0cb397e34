#!/usr/bin/env python
"""Every BASELINE.json config AT ITS STATED SIZE, one JSON line per config, each with a `clocks` record sampled during its timed region.
bench.py stays the headline line; this is the table behind DESIGN.md section 5.

  python tools/bench_configs.py [--configs 1,2,3,4,5,h,g,x] [--min-seconds 1.0]  > gpurun_out/configs.jsonl
  torchrun --nproc-per-node 8 tools/bench_configs.py --configs 4,5            (configs 4 and 5 "on 8 B200": weak shards, config 4 + all-reduce)

  1  C6, N = 1e4, getJointTorque + getRegressor on the CPU: 1 thread and all cores, -O3 and -Ofast builds of the restatement and the
     reference's own headers (recipe rosdyn_core/test/rosdyn_speed_test.cpp:106-192, flags rosdyn_core/CMakeLists.txt:88-93); one
     steady-clock interval around the whole loop
  2  C6, 1e8 samples: pose + Jacobian + twists + acceleration twists of all links + RNEA torque (1 344 B/sample), 4 chunks of 25 M
  3  C7, 1e8 samples: materialised regressor 7x70 (4 088 B/sample), 25 chunks of 4 M (409 GB of Phi do not fit: the chunk buffer is reused)
  4  C7, 1e9 samples: fused regressor -> Gram, sharded over the ranks; per rank chunks of <= 125 M samples (168 B/sample of inputs), inputs
     regenerated on the device between chunks (untimed), kernels timed with CUDA events per chunk
  5  C6, 5e7 samples: getJointInertia + getDDTwist (912 B/sample as two launches)
  h  headline chain C6: materialised regressor + torque (3 552 B/sample) over 1e8 samples; fused Gram over 1e9 samples

Timing: device-resident inputs far larger than L2, >= 3 warm-up launches, CUDA events on the launching stream, passes repeated until the timed
region is at least --min-seconds.  Under torchrun the time is the max over ranks and the rate is the whole job's."""
import argparse
import ctypes
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

SEED = 0x5EED0000


def cpu_config1(out):
    from bench import host_threads
    from oracle import oracle
    from oracle.oracle import OracleChain, fill_uniform
    from rosdyn_b200 import fixtures
    d = fixtures.by_name("c6")
    n = 10_000
    q, dq, ddq = (fill_uniform(6, n, SEED + 1, s) for s in range(3))
    threads_all = host_threads()
    builds = [("port -O3 -march=x86-64-v3 (restatement oracle/rosdyn_oracle.c)", False), ("port -Ofast -ffast-math -funroll-loops -march=x86-64-v3 (flags of rosdyn_core/CMakeLists.txt:88,92)", True)]
    if oracle.have_ref():
        builds.append(("reference's own headers (oracle/_ref, stand-in Eigen)", "ref"))
    for label, fast in builds:
        oc = OracleChain(d, fast=fast)
        for threads in (1, threads_all):
            best = None
            for _ in range(5 if fast != "ref" else 2):
                t = time.perf_counter()
                oc.regressor_torque(q, dq, ddq, nthreads=threads, store=False)
                dt = time.perf_counter() - t
                best = dt if best is None else min(best, dt)
            out({"config": "config1: C6, N = 1e4, getJointTorque + getRegressor per sample on the CPU", "build": label, "threads": threads,
                 "host_threads_available": threads_all, "samples": n, "seconds": best, "us_per_sample": best / n * 1e6, "samples_per_s": n / best,
                 "note": "best of 5 whole-loop intervals; the published reference figure for RNEA alone is 3.77 us/sample on a 2014 laptop (README.md:43)"})


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="1,2,3,4,5,h,g,x")
    ap.add_argument("--min-seconds", type=float, default=1.0)
    ap.add_argument("--scale", type=float, default=1.0, help="multiply every sample count (smoke runs)")
    args = ap.parse_args()
    want = set(args.configs.split(","))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    def out(r):
        if rank == 0:
            print(json.dumps(r), flush=True)

    if "1" in want and rank == 0:
        cpu_config1(out)
    if not (want - {"1"}):
        return

    import torch
    import torch.distributed as dist

    from bench import ClockSampler
    from rosdyn_b200 import fixtures
    from rosdyn_b200._lib import CDynamicsOut, CKinematicsOut, CSamples, check, load  # noqa: F401
    from rosdyn_b200.chain import Chain, fill_uniform, fp64_peak
    from rosdyn_b200.sharding import Group

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        sys.stdout.flush()
        real = os.dup(1)
        os.dup2(2, 1)   # NCCL banner away from the JSON lines
        dist.init_process_group("nccl", device_id=dev)
        os.dup2(real, 1)
    lib = load()
    hbm = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    fp64 = max(fp64_peak("dmma", 3), fp64_peak("dfma", 3))

    def stream():
        return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(pass_fn, samples_per_pass, between=None):
        """pass_fn(): launches one pass (samples_per_pass samples on this rank) on the current stream and returns nothing; repeated until the timed
        region reaches min-seconds.  `between` (untimed) runs before every pass (e.g. regenerate inputs).  Returns (seconds, passes, clocks)."""
        for _ in range(3):
            if between:
                between()
            pass_fn()
        sync_all()
        total_ms, passes = 0.0, 0
        wall0 = time.perf_counter()
        with ClockSampler(local) as clk:
            # at least min-seconds of timed kernels, but never more than 30 s of wall clock per config (untimed regeneration included)
            while total_ms < args.min_seconds * 1e3 and passes < 10000 and time.perf_counter() - wall0 < 30.0:
                if between:
                    between()
                    torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                pass_fn()
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1)
                if world > 1:
                    t = torch.tensor([ms], dtype=torch.float64, device=dev)
                    dist.all_reduce(t, op=dist.ReduceOp.MAX)
                    ms = float(t[0])
                total_ms += ms
                passes += 1
        return total_ms * 1e-3, passes, clk.summary()

    def report(name, samples_per_pass_rank, sec, passes, clocks, bytes_per_sample=None, flop_per_sample=None, **extra):
        rate = world * samples_per_pass_rank * passes / sec
        r = {"config": name, "n_gpus": world, "samples_per_pass": world * samples_per_pass_rank, "passes": passes, "timed_region_s": sec,
             "samples_per_s": rate, "clocks": clocks}
        if bytes_per_sample:
            gbs = rate / world * bytes_per_sample / 1e9
            r.update(bytes_per_sample=bytes_per_sample, achieved_GBps_per_gpu=gbs, hbm_peak_GBps=hbm, hbm_frac=gbs / hbm)
        if flop_per_sample:
            tf = rate / world * flop_per_sample / 1e12
            r.update(flop_per_sample=flop_per_sample, achieved_TFLOPs_per_gpu=tf, fp64_peak_TFLOPs=fp64, fp64_frac=tf / fp64)
        r.update(extra)
        out(r)

    def N(x):
        return max(1024, int(x * args.scale))

    def inputs(n_in, S, cfg, streams=3):
        return [fill_uniform(n_in, S, SEED + cfg + 1000003 * rank, s, device=dev) for s in range(streams)]

    # ------------------------------------------------------------------ config 2
    if "2" in want and world == 1:
        d = fixtures.by_name("c6")
        ch = Chain(d)
        n_in, nL, total, chunk = 6, 8, N(1e8), N(25e6)
        q, dq, ddq = inputs(n_in, total, 2)
        outs = {k: torch.empty((r, chunk), dtype=torch.float64, device=dev) for k, r in
                (("T_tool", 12), ("jacobian", 6 * n_in), ("twist", 6 * nL), ("dtwist", 6 * nL), ("torque", n_in))}
        ko = CKinematicsOut()
        ko.ld = chunk
        for k, v in outs.items():
            setattr(ko, k, v.data_ptr())
        smps = [CSamples(min(chunk, total - o), total, q[:, o:].data_ptr(), dq[:, o:].data_ptr(), ddq[:, o:].data_ptr(), None)
                for o in range(0, total, chunk)]

        def p2():
            for s in smps:
                check(lib.rdb_kinematics_batch(ch._h, ctypes.byref(s), ctypes.byref(ko), stream()))
        sec, passes, clk = timed(p2, total)
        report("config2: C6, 1e8 samples, pose + Jacobian + twist + dtwist (all links) + RNEA torque [kin_kernel<7,CFG2,NP>]", total, sec, passes, clk,
               bytes_per_sample=8 * (18 + 150), chunks=len(smps))
        del outs, q, dq, ddq
        torch.cuda.empty_cache()

    # ------------------------------------------------------------------ config 5 (1 or N GPUs: weak shards, no collective)
    if "5" in want:
        d = fixtures.by_name("c6")
        ch = Chain(d)
        n_in, nL = 6, 8
        total = N(5e7) if world == 1 else N(5e7)   # per rank: "1 and 8 B200" = 5e7 samples on every GPU (weak scaling)
        q, dq, ddq, dddq = inputs(n_in, total, 5, streams=4)
        M = torch.empty((n_in * n_in, total), dtype=torch.float64, device=dev)
        jerk = torch.empty((6 * nL, total), dtype=torch.float64, device=dev)
        ko5 = CKinematicsOut()
        ko5.ld = total
        ko5.ddtwist = jerk.data_ptr()
        smp5 = CSamples(total, total, q.data_ptr(), dq.data_ptr(), ddq.data_ptr(), dddq.data_ptr())

        def p5a():
            check(lib.rdb_inertia_batch(ch._h, ctypes.byref(smp5), M.data_ptr(), total, stream()))

        def p5b():
            check(lib.rdb_kinematics_batch(ch._h, ctypes.byref(smp5), ctypes.byref(ko5), stream()))

        def p5():
            p5a()
            p5b()
        sec, passes, clk = timed(p5a, total)
        report("config5a: C6, 5e7 samples per GPU, getJointInertia [dyn_kernel<6,INERTIA,REV>]", total, sec, passes, clk, bytes_per_sample=8 * (6 + 36),
               note="FP64-pipe bound at this byte count")
        sec, passes, clk = timed(p5b, total)
        report("config5b: C6, 5e7 samples per GPU, getDDTwist of all links [kin_kernel<7,JERK,NP>]", total, sec, passes, clk, bytes_per_sample=8 * (24 + 48))
        sec, passes, clk = timed(p5, total)
        report("config5: C6, 5e7 samples per GPU, getJointInertia + getDDTwist (two launches back to back)", total, sec, passes, clk,
               bytes_per_sample=8 * (6 + 36) + 8 * (24 + 48))
        del M, jerk, q, dq, ddq, dddq
        torch.cuda.empty_cache()

    # ------------------------------------------------------------------ config 3
    if "3" in want and world == 1:
        d7 = fixtures.by_name("c7")
        ch7 = Chain(d7)
        n_in, total, chunk = 7, N(1e8), N(4e6)
        q, dq, ddq = inputs(n_in, total, 3)
        phi = torch.empty((70 * n_in, chunk), dtype=torch.float64, device=dev)
        smps = [CSamples(min(chunk, total - o), total, q[:, o:].data_ptr(), dq[:, o:].data_ptr(), ddq[:, o:].data_ptr(), None)
                for o in range(0, total, chunk)]

        def p3():
            for s in smps:
                check(lib.rdb_regressor_batch(ch7._h, ctypes.byref(s), phi.data_ptr(), None, chunk, stream()))
        sec, passes, clk = timed(p3, total)
        report("config3: C7, 1e8 samples, materialised regressor 7x70 as 490 SoA planes [dyn_kernel<7,REGRESSOR>]", total, sec, passes, clk,
               bytes_per_sample=8 * (21 + 490), chunks=len(smps))
        del phi, q, dq, ddq
        torch.cuda.empty_cache()

    # ------------------------------------------------------------------ config 4 and the headline fused Gram: 1e9 samples over the ranks
    def gram_1e9(chain_name, cfg, label, flop):
        d = fixtures.by_name(chain_name)
        n_in, P = d.n_inputs, 10 * d.n_joints
        grp = Group.from_torch_distributed(d, local)
        per_rank = N(1e9) // world
        chunk = min(per_rank, N(125e6))
        nchunks = (per_rank + chunk - 1) // chunk
        q, dq, ddq = (torch.empty((n_in, chunk), dtype=torch.float64, device=dev) for _ in range(3))
        G = torch.zeros((P, P), dtype=torch.float64, device=dev)
        b = torch.zeros((P,), dtype=torch.float64, device=dev)
        tt = torch.zeros((1,), dtype=torch.float64, device=dev)
        state = {"k": 0}

        def regen():   # untimed: the next chunk's inputs (distinct sample indices through the seed)
            k = state["k"]
            for s, x in enumerate((q, dq, ddq)):
                check(lib.rdb_fill_uniform(x.data_ptr(), n_in, chunk, chunk, SEED + cfg + 1000003 * rank + 7919 * k, s, stream()))
            state["k"] = k + 1

        def pg():      # one chunk: fused kernel (+ the all-reduce of the partials when sharded), accumulated into G / b / tau_sq
            grp.gram([(q, dq, ddq)], out=[(G, b, tt)], accumulate=True)
        sec, passes, clk = timed(pg, chunk, between=regen)
        # passes = chunks timed; a full 1e9-sample job is nchunks chunks per rank
        job_s = sec / passes * nchunks
        report(label, chunk, sec, passes, clk, flop_per_sample=flop, chunks_per_rank_for_1e9=nchunks, seconds_for_1e9_samples=job_s,
               note="inputs regenerated on the device between chunks (untimed); every chunk accumulates into the same normal equations"
                    + ("; one ncclAllReduce of the 4 971 partials per chunk inside the C-ABI group" if world > 1 else ""))
        del grp, q, dq, ddq
        torch.cuda.empty_cache()

    if "4" in want:
        gram_1e9("c7", 4, f"config4: C7, 1e9 samples, fused regressor -> PhiT Phi / PhiT tau, {world} GPU(s) [gram_fused_kernel<7>]", 7 * 70 * 71 + 2 * 7 * 70)
    if "h" in want:
        gram_1e9("c6", 1, f"headline (fused Gram): C6, 1e9 samples, {world} GPU(s) [gram_fused_kernel<6>]", 6 * 70 * 71 + 2 * 6 * 70)
        if world == 1:
            d = fixtures.by_name("c6")
            ch = Chain(d)
            n_in, total, chunk = 6, N(1e8), N(4e6)
            q, dq, ddq = inputs(n_in, total, 1)
            phi = torch.empty((70 * n_in + n_in, chunk), dtype=torch.float64, device=dev)
            smps = [CSamples(min(chunk, total - o), total, q[:, o:].data_ptr(), dq[:, o:].data_ptr(), ddq[:, o:].data_ptr(), None)
                    for o in range(0, total, chunk)]

            def ph():
                for s in smps:
                    check(lib.rdb_regressor_batch(ch._h, ctypes.byref(s), phi.data_ptr(), phi[70 * n_in:].data_ptr(), chunk, stream()))
            sec, passes, clk = timed(ph, total)
            report("headline (materialised): C6, 1e8 samples, getRegressor 6x70 + torque as 426 SoA planes [dyn_kernel<7,REGRESSOR|TORQUE>]", total, sec,
                   passes, clk, bytes_per_sample=8 * (18 + 420 + 6), chunks=len(smps))
            # the same in the Eigen-record layout (RDB_LAYOUT_EIGEN): one dense column-major 6x70 record per sample
            do = CDynamicsOut()
            do.ld = chunk
            do.regressor, do.torque, do.layout = phi.data_ptr(), phi[70 * n_in:].data_ptr(), 1

            def pe():
                for s in smps:
                    check(lib.rdb_dynamics_batch(ch._h, ctypes.byref(s), ctypes.byref(do), stream()))
            sec, passes, clk = timed(pe, total)
            report("headline (materialised, RDB_LAYOUT_EIGEN records): C6, 1e8 samples, getRegressor + torque", total, sec, passes, clk,
                   bytes_per_sample=8 * (18 + 420 + 6), chunks=len(smps))
    # ------------------------------------------------------------------ extended model [Phi | Phi_c] (N2): one friction component per joint
    if "x" in want and world == 1:
        for cname in ("c6", "c7"):
            d = fixtures.by_name(cname)
            ch = Chain(d)
            ch.setComponents([{"type": "friction1", "joint": j, "min_velocity": 0.01, "max_velocity": 2.0} for j in ch.getActiveJointsName()])
            n_in, total = d.n_inputs, N(1e8)
            q, dq, ddq = inputs(n_in, total, 6)
            Pt = 10 * d.n_joints + 2 * n_in

            def px():
                ch.regressorGramExt(q, dq, ddq)
            sec, passes, clk = timed(px, total)
            report(f"extended model: {cname.upper()}, 1e8 samples, fused [Phi | Phi_c] -> normal equations ({Pt} x {Pt}), friction_polynomial1 on every joint "
                   f"[gram_ext_kernel<{n_in}>]", total, sec, passes, clk,
                   flop_per_sample=n_in * Pt * (Pt + 1) + 2 * n_in * Pt,
                   note="algorithmic flops of the dense n_act x (10 nJ + components) regressor (SYRK + GEMV convention)")
            del q, dq, ddq, ch
            torch.cuda.empty_cache()
    # ------------------------------------------------------------------ chains the unrolled kernels do not cover (> 8 moving joints): *_kernel_generic
    if "g" in want and world == 1:
        d = fixtures.random_chain(909, 12, p_prismatic=0.1, p_fixed=0.0)   # 12 moving joints: runtime loops, model in global memory
        ch = Chain(d)
        n_in, P, total = d.n_inputs, 10 * d.n_joints, N(2e6)
        q, dq, ddq = inputs(n_in, total, 9)
        tau = torch.empty((n_in, total), dtype=torch.float64, device=dev)
        phi = torch.empty((P * n_in, total), dtype=torch.float64, device=dev)
        smp = CSamples(total, total, q.data_ptr(), dq.data_ptr(), ddq.data_ptr(), None)
        sec, passes, clk = timed(lambda: check(lib.rdb_torque_batch(ch._h, ctypes.byref(smp), tau.data_ptr(), total, stream())), total)
        report("generic kernels: 12 moving joints, getJointTorque [dyn_kernel_generic<TORQUE>]", total, sec, passes, clk, bytes_per_sample=8 * 4 * n_in,
               note="chains with more than 8 moving joints: runtime loops, per-joint state in local memory")
        sec, passes, clk = timed(lambda: check(lib.rdb_regressor_batch(ch._h, ctypes.byref(smp), phi.data_ptr(), None, total, stream())), total)
        report("generic kernels: 12 moving joints, getRegressor 12x120 [dyn_kernel_generic<REGRESSOR>]", total, sec, passes, clk,
               bytes_per_sample=8 * (3 * n_in + P * n_in))
        G = torch.empty((P, P), dtype=torch.float64, device=dev)
        b = torch.empty((P,), dtype=torch.float64, device=dev)
        tt = torch.empty((1,), dtype=torch.float64, device=dev)
        sec, passes, clk = timed(lambda: check(lib.rdb_regressor_gram_batch(ch._h, ctypes.byref(smp), None, G.data_ptr(), b.data_ptr(), tt.data_ptr(), 0,
                                                                            stream())), total)
        report("general Gram pipeline: 12 moving joints (dyn_kernel_generic -> L2-sized workspace -> syrk_dmma_kernel)", total, sec, passes, clk,
               flop_per_sample=n_in * P * (P + 1) + 2 * n_in * P)
        del phi, tau, q, dq, ddq
        torch.cuda.empty_cache()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
