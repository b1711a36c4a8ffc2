#!/usr/bin/env python
"""bench.py -- 3D-GCUPS of the three-sequence alignment hot path on B200 (contract in the task statement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

A "step" is one pass of the hot path over one batch of synthetic ACGT triples.  Default workload (N=1) is
BASELINE.json configs[1]: 1M triples of length 64, SCORE_BITS=12 (the testbench value), on one GPU.  Under torchrun
(N>1) every rank processes its own 1M-triple shard (independent triples, no data-path collective) -> weak scaling.
`value` times the device-resident call (inputs already in HBM); `e2e` times the host-pointer C-ABI call with pinned
host buffers, H2D and D2H copies inside the timed region.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

OPS_PER_CELL = 91  # 49 adds + 42 max per cell update (PE_1cyc.v:164-218, :22-28) -- SURVEY.md 8(d)

WORKLOADS = {
    # name: (BASELINE config index, triples per GPU, length, score_bits)
    "batch64": (2, 1 << 20, 64, 12),
    "batch256": (4, 1 << 18, 256, 12),
    "long1024": (3, 1, 1024, 12),
    "long8192": (5, 1, 8192, 16),
}


def kernel_stats(workload):
    """Per-launch DRAM traffic and issued-instruction utilisation of the dominant kernel, taken from the committed
    `ncu --set full` captures (profiles/ncu_kernel_stats.json); never measured under this process."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_kernel_stats.json")) as f:
            return json.load(f).get(workload)
    except Exception:
        return None


def alu_peak():
    """Measured integer-ALU peak (tools/alu_microbench.cu on this pool's B200), else the nominal figure."""
    path = os.path.join(ROOT, "profiles", "alu_peak.json")
    try:
        with open(path) as f:
            d = json.load(f)
        return float(d["peak_s16x2_lane_ops_per_s"]), "measured (profiles/alu_peak.json: 148 SMs x 64 lanes/clk x 2 (s16x2) x measured clock)"
    except Exception:
        return 148 * 64 * 2 * 1.965e9, "nominal (148 SMs x 64 lanes/clk x 2 x 1.965 GHz)"


class ClockSampler:
    """Samples nvidia-smi during the timed region (B200_PROFILING.md's clocks line)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def host_threads():
    """Cores this process may use (torchrun exports OMP_NUM_THREADS=1, which must not shrink the CPU baseline)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def cpu_baseline(length, score_bits, target_s=22.0):
    """Times the CPU oracle (kind 'port': O1, all host threads) and the RTL cycle model on a bounded sample."""
    import numpy as np

    import oracle_bindings as ob

    threads = host_threads()
    L = min(length, 256) if length > 256 else length
    per = float(L) ** 3
    probe_n = max(4 * threads, 64) if L <= 64 else max(2, threads // 2)
    a, b, c = ob.synth_batch(0x5EED0002, 0, probe_n, L)
    rate = 0.0
    for _ in range(2):  # the first pass also pays for the OpenMP thread start-up
        t0 = time.perf_counter()
        ob.score_batch(a, b, c, score_bits, threads)
        rate = max(rate, probe_n * per / (time.perf_counter() - t0))
    n = int(max(probe_n, min(1 << 18, target_s * rate / per)))
    a, b, c = ob.synth_batch(0x5EED0002, 0, n, L)
    t0 = time.perf_counter()
    ob.score_batch(a, b, c, score_bits, threads)
    dt = time.perf_counter() - t0
    out = {"value": n * per / dt / 1e9, "unit": "GCUPS", "cores": threads, "kind": "port",
           "sample": f"{n} synthetic triples of length {L}, W={score_bits}, functional oracle O1 (OpenMP over triples), {dt:.1f} s"}
    # RTL cycle model (the 'RTL simulation' of the north star), one triple per thread, L=64 sample
    try:
        import ctypes as C

        m = max(threads, 8) * 48  # about 5 s: one triple of length 64 is 9,282 simulated clocks
        ra, rb, rc = ob.synth_batch(0x5EED0002, 0, m, 64)
        cfg = ob.RtlCfg()
        ob.lib().rtl_default_cfg(C.byref(cfg), 1)
        sc = np.zeros(m, np.int32)
        tot = C.c_int64(0)
        t0 = time.perf_counter()
        rcode = ob.lib().rtl_model_run_batch(C.byref(cfg), ra.ctypes.data_as(ob.u8p), rb.ctypes.data_as(ob.u8p),
                                             rc.ctypes.data_as(ob.u8p), m, 64, threads, sc.ctypes.data_as(ob.i32p), C.byref(tot))
        dt = time.perf_counter() - t0
        if rcode == 0:
            out["rtl_sim"] = {"kind": "hand-written cycle-accurate C++ model of TriAlign_1cyc/PE_1cyc (no Verilog simulator in the image)",
                              "gcups": m * 64.0 ** 3 / dt / 1e9, "sim_cycles_per_s": tot.value / dt, "cores": threads,
                              "sample": f"{m} triples of length 64, {dt:.1f} s"}
    except Exception as e:  # pragma: no cover
        out["rtl_sim"] = {"error": str(e)}
    return out


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU path for this metric = the RTL simulated on host cores.  No Verilog
    simulator exists here, so the functional oracle (port) is timed, with all host threads, on a bounded sample."""
    if rank != 0:
        return
    import oracle_bindings as ob

    cfg, n_full, length, w = WORKLOADS[args.workload]
    threads = host_threads()
    L = min(length, 256)
    per = float(L) ** 3
    n = max(threads * 2, 16) if L <= 64 else max(2, threads // 2)
    a, b, c = ob.synth_batch(0x5EED0000 + cfg, 0, n, L)
    t0 = time.perf_counter()
    ob.score_batch(a, b, c, w, threads)
    probe = time.perf_counter() - t0
    n = int(max(n, min(1 << 15, 6.0 / max(probe / n, 1e-9))))  # ~6 s per step
    a, b, c = ob.synth_batch(0x5EED0000 + cfg, 0, n, L)
    for _ in range(max(args.warmup, 0) and 1):
        ob.score_batch(a, b, c, w, threads)
    times = []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        ob.score_batch(a, b, c, w, threads)
        times.append(time.perf_counter() - t0)
    dt = sum(times) / len(times)
    val = n * per / dt / 1e9
    sample = f"{n} of the workload's triples (length {L}) per step, functional oracle O1, {threads} threads"
    line = {"impl": "reference", "metric": "3D-GCUPS", "value": val, "unit": "GCUPS", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int32 (wrap at W)", "data": "synthetic",
            "config": {"workload": args.workload, "baseline_config": cfg, "triples_per_gpu": n_full, "length": length,
                       "score_bits": w, "sample_triples_per_step": n, "sample_length": L},
            "cpu_baseline": {"value": val, "unit": "GCUPS", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def run_ours(args, rank, world, local_rank):
    import numpy as np
    import torch

    import trialign_b200 as ta

    cfg, n_full, length, w = WORKLOADS[args.workload]
    n = args.triples if args.triples else n_full
    reduced = bool(args.triples and args.triples != n_full)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)
    eng = ta.TriAlign(score_bits=w, devices=[local_rank])
    seed = 0x5EED0000 + cfg
    single = length > 256  # one long triple (configs 3 and 5) instead of a batch
    slab = single and world > 1
    scaling = "strong" if slab else "weak"
    first = 0 if single else rank * n  # batch: each rank owns a contiguous shard of independent triples
    da, db, dc = ta.synth_batch_cuda(seed, first, n, length, device=local_rank)
    cells_step = float(n) * float(length) ** 3
    if single:
        out = torch.zeros(8, dtype=torch.int32, device=dev)
        a1, b1, c1 = da[0], db[0], dc[0]
        if slab:
            # config 5: i-slab pipeline, rank g computes planes (la*g/N, la*(g+1)/N] and hands block states to rank g+1
            # (TRIALIGN_MULTI=slab), or the column pipeline: rank g sweeps all planes of its block-column range and streams
            # the faces crossing the range edge to rank g+1 while it runs
            use_cols = os.environ.get("TRIALIGN_MULTI", "slab") == "cols"
            pipe = (ta.ColumnPipeline(length, length, length, score_bits=w, device=local_rank) if use_cols
                    else ta.SlabPipeline(length, length, score_bits=w, device=local_rank))
            sync_t = torch.zeros(1, device=dev)

            def step():
                pipe.run(a1, b1, c1, out=out)
                dist.all_reduce(sync_t)  # stream-ordered: nobody starts run k+1 before every rank finished run k
            cells_step /= world  # per rank; the job total is restored below
        else:
            def step():
                eng.align_cuda(a1, b1, c1, out=out)
    else:
        out = torch.empty(n, dtype=torch.int32, device=dev)

        def step():
            eng.align_batch_cuda(da, db, dc, out=out)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    path = ta.last_path()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = ta.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    launches = ta.launch_count() - l0
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    clocks = sampler.stop() if rank == 0 else None
    if slab:  # the score lives on the last rank
        res = out[:1].clone().to(torch.int64)
        if rank != world - 1:
            res.zero_()
        dist.all_reduce(res)
        checksum = int(res.item())
    else:
        checksum = int(out[0].item()) if single else int(out.sum().item())

    # ---- e2e: host buffers through the C ABI (H2D + kernels + D2H inside the timed region) ----
    import ctypes as C

    L = ta.library()
    u8p, i32p = C.POINTER(C.c_uint8), C.POINTER(C.c_int32)
    ha, hb, hc = (x.cpu().pin_memory() for x in (da, db, dc))
    if slab:
        # the pipeline has no host-pointer form; e2e = device run + staging of the three sequences and the result
        hres = torch.empty(8, dtype=torch.int32).pin_memory()

        def e2e_step():
            da.copy_(ha, non_blocking=True); db.copy_(hb, non_blocking=True); dc.copy_(hc, non_blocking=True)
            step()
            hres.copy_(out, non_blocking=True)
            torch.cuda.synchronize()
        h2d, d2h = 3 * length, 32
    elif single:
        hs = C.c_int32(0)

        def e2e_step():
            rc = L.trialign_score(C.cast(ha.data_ptr(), u8p), length, C.cast(hb.data_ptr(), u8p), length,
                                  C.cast(hc.data_ptr(), u8p), length, w, 0, 1 << local_rank, C.byref(hs), None)
            if rc:
                raise RuntimeError(f"trialign_score rc={rc}")
        h2d, d2h = 3 * length, 4
    else:
        hout = torch.empty(n, dtype=torch.int32).pin_memory()

        def e2e_step():
            rc = L.trialign_score_batch(C.cast(ha.data_ptr(), u8p), C.cast(hb.data_ptr(), u8p), C.cast(hc.data_ptr(), u8p),
                                        n, length, w, 0, 1 << local_rank, C.cast(hout.data_ptr(), i32p))
            if rc:
                raise RuntimeError(f"trialign_score_batch rc={rc}")
        h2d, d2h = 3 * n * length, 4 * n

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(1, min(args.steps, 5))
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s = float(te.item())
    if single and not slab and int(hs.value) != checksum:
        raise RuntimeError("e2e score differs from the device-resident score")
    if not single and int(hout.sum().item()) != checksum:
        raise RuntimeError("e2e scores differ from the device-resident scores")

    if rank == 0:
        per_step_s = ms_max / 1e3 / args.steps
        gcups = cells_step * world / per_step_s / 1e9
        peak, peak_how = alu_peak()
        achieved = cells_step / per_step_s * OPS_PER_CELL  # per GPU, the dominant (only) kernel of the step
        line = {
            "metric": "3D-GCUPS", "value": gcups, "unit": "GCUPS", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": per_step_s * 1e3, "higher_is_better": True, "scaling": scaling,
            "vs_baseline": None, "dtype": "u16x2 (packed 16-bit integer lanes)" if path.startswith("fast") else "int16/int32",
            "data": "synthetic",
            "config": {"workload": args.workload, "baseline_config": cfg, "triples_per_gpu": n, "length": length,
                       "parallelism": (("column pipeline" if os.environ.get("TRIALIGN_MULTI", "slab") == "cols" else "i-slab pipeline") + " over %d GPUs, NVLink peer stores" % world) if slab else ("batch shard per GPU, no exchange" if world > 1 else "single GPU"),
                       "score_bits": w, "kernel_path": path, "l2": "inputs (3 x n x len bytes) larger than L2" if 3 * n * length > 126e6 else "inputs smaller than L2; compute-bound kernel, no reuse across steps assumed",
                       "reduced": reduced, "scores_checksum": checksum},
            "gpu_launches": launches,
            "e2e": {"value": cells_step * world / e2e_s / 1e9, "unit": "GCUPS", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h},
            "roofline": {"bound": "int_alu", "achieved": achieved / 1e12, "peak": peak / 1e12, "unit": "Tlane-op/s",
                         "frac": achieved / peak, "traffic": None, "peak_source": peak_how,
                         "algorithmic_ops_per_cell": OPS_PER_CELL,
                         "note": "achieved counts the RTL's 91 lane-ops per cell; the factored kernel issues fewer (see DESIGN.md), so frac can exceed 1"},
            "clocks": clocks,
        }
        ks = kernel_stats(args.workload)
        if ks:
            # DRAM bytes of one launch from the ncu capture, scaled by the number of triples when the captured launch
            # was a reduced batch (traffic is the input arrays + scores, linear in n)
            scale = float(n) / ks["triples_in_capture"] if ks.get("triples_in_capture") else 1.0
            if slab:
                scale = 1.0 / world  # every GPU of the i-slab pipeline sweeps 1/world of the planes (faces per plane are fixed)
            line["roofline"]["traffic"] = (ks["dram_bytes_read"] + ks["dram_bytes_write"]) * scale
            line["roofline"]["issued"] = {k: ks[k] for k in ("alu_pipe_pct", "issue_slot_pct", "warp_instr_per_cell_pair",
                                                              "alu_instr_per_cell_pair", "capture") if k in ks}
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline(length, w)
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="batch64", choices=sorted(WORKLOADS))
    ap.add_argument("--triples", type=int, default=0, help="override triples per GPU (development only; marks the line reduced)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
