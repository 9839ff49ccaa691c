#!/usr/bin/env python
"""bench.py -- headline benchmark of the DEDFlow FEM linear-system hot path on B200.

A STEP is one pass of the hot path the reference executes per Newton iteration (reference src/main.c:157-221):
    AssembleSystem(F)  +  AssembleSystem(J)  (tets + weak-BC faces + Dirichlet)  +  KrylovSolve(ksp, J, dx, F)
on the synthetic Kuhn box mesh of BASELINE.json configs[1] (m=55: 998,250 tets, 175,616 nodes, fp64), state B.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--m 55] [--impl ours|reference]

Prints ONE JSON line (see the keys below).  `value` = elements / second through the whole step with inputs resident
in HBM; `e2e` = the same with host buffers (pinned H2D of the nodal states, D2H of the solution inside the timed
region); `breakdown` carries the three quantities BASELINE.json names separately (assembly elems/s, SpMV GB/s and
% of measured HBM peak, Krylov solve s/step); `roofline` is the dominant kernel of the step.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "assembled elems/s"
UNIT = "elems/s"

# The contract is ONE JSON line on stdout.  Libraries (NCCL's version banner, the reference's printf logging) write to fd 1
# too, so everything else is routed to stderr while the benchmark runs and the line goes to the original stdout.
_STDOUT_FD = None


def protect_stdout():
    global _STDOUT_FD
    if _STDOUT_FD is None:
        sys.stdout.flush()
        _STDOUT_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    sys.stdout.flush()
    os.write(_STDOUT_FD if _STDOUT_FD is not None else 1, data)


def measured_hbm_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def ev_ms(torch, fn, reps):
    """median CUDA-event time of fn() in ms"""
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


def workload_text(m, E, N):
    """config.workload of BOTH arms at N = 1 (identical text: the driver compares the strings)."""
    return (f"BASELINE configs[1]: Kuhn box m={m}, {E} tets, {N} nodes, fp64, state B; step = AssembleSystem(F) + AssembleSystem(J) "
            "(tets + weak-BC faces + Dirichlet) + KrylovSolve (GMRES(120), the reference's stopping rule)")


def algorithmic_bytes(N, E, Z):
    """SURVEY.md §8(d) / DESIGN.md: compulsory bytes per launch."""
    return {
        "assemble_J": 16 * E + 48 * N + 128 * Z,
        "assemble_F": 16 * E + 160 * N,
        "spmv": 132 * Z + 4 * (N + 1) + 64 * N,          # our single-index format: 16 f64 + 1 i32 per nodal nonzero
        "spmv_reference_format": 192 * Z + 96 * N,       # four scalar-CSR blocks as cuSPARSE reads them
    }


# ----------------------------------------------------------------------------------------------------------------
def cpu_baseline(mesh, wg, dwg, with_solve=True):
    """The oracle port of the same step timed on the host cores (reported baseline, not the target)."""
    from oracle import pyoracle
    O = pyoracle.get()
    N, E = mesh.num_node, mesh.num_tet
    t_setup = time.time()
    rp, ci = O.nodal_pattern(N, mesh.ien)
    w = O.weights(pyoracle.curand_host_u32(E))
    color, nc, _ = O.color_jpl(N, mesh.ien, w)
    off, ind = O.color_batches(color)
    t_setup = time.time() - t_setup
    Z = ci.size
    F = np.zeros(6 * N)
    blocks = [np.zeros(9 * Z), np.zeros(3 * Z), np.zeros(3 * Z), np.zeros(Z)]
    f2e, forn = mesh.bound_faces(4)
    bcs = {0: (1, 1, 1), 2: (0, 1, 0), 3: (0, 0, 1), 4: (0, 0, 0)}
    t0 = time.time()
    O.assemble_tet(N, mesh.ien, mesh.xg, off, ind, wg, dwg, F=F)
    O.assemble_face(f2e, forn, N, mesh.ien, mesh.xg, color, nc, wg, dwg, F=F)
    F[4 * N:] = 0
    for b, t in bcs.items():
        O.dirichlet_vec(mesh.bound_nodes(b), np.array(t, np.int32), F)
    tF = time.time() - t0
    t0 = time.time()
    O.assemble_tet(N, mesh.ien, mesh.xg, off, ind, wg, dwg, pattern=(rp, ci), blocks=blocks)
    O.assemble_face(f2e, forn, N, mesh.ien, mesh.xg, color, nc, wg, dwg, pattern=(rp, ci), blocks=blocks)
    for b, t in bcs.items():
        O.dirichlet_mat(mesh.bound_nodes(b), np.array(t, np.int32), N, (rp, ci), blocks[0], blocks[1])
    tJ = time.time() - t0
    ts, it = 0.0, 0
    if with_solve:
        t0 = time.time()
        x, it, hist = O.gmres((rp, ci), blocks, F)
        ts = time.time() - t0
    step = tF + tJ + ts
    return {"value": E / step, "unit": UNIT, "cores": O.num_threads(), "kind": "port",
            "sample": f"1 full step (assemble F {tF:.2f}s + J {tJ:.2f}s + GMRES {it} its {ts:.2f}s) of the m={mesh.m} mesh, "
                      f"oracle/oracle.c with OpenMP",
            "assemble_elems_per_s": E / (tF + tJ), "solve_s_per_step": ts, "setup_s": t_setup}


# ----------------------------------------------------------------------------------------------------------------
def run_reference(args, rank, world):
    """--impl reference: the reference's OWN implementation of the step.  The reference has no CPU implementation of this
    path -- its only implementation is CUDA -- so when oracle/_ref (built from /root/reference/src, unmodified) loads and
    a GPU is present, that is what is timed (kind "reference"); otherwise the oracle port runs on the host cores."""
    if rank != 0:
        return
    from dedflow_b200 import boxmesh
    mesh = boxmesh.make_box(args.m)
    N, E = mesh.num_node, mesh.num_tet
    wg, dwg = boxmesh.state_random(N)
    line = {"metric": METRIC, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_text(args.m, E, N)}}
    use_cuda_ref = False
    try:
        import torch
        from oracle.ref import reflib
        use_cuda_ref = torch.cuda.is_available() and reflib.available() and not args.ref_port
    except Exception:
        use_cuda_ref = False
    if use_cuda_ref:
        R = reflib.RefProblem(mesh, patch_d1=True)
        R.color_batches()
        d_wg, d_dwg = torch.from_numpy(wg).cuda(), torch.from_numpy(dwg).cuda()
        F = torch.zeros(6 * N, dtype=torch.float64, device="cuda")
        dx = torch.zeros(6 * N, dtype=torch.float64, device="cuda")
        its = [0]

        def step():
            R.assemble(d_wg, d_dwg, F_t=F)
            R.assemble(d_wg, d_dwg, J=True)
            dx.zero_()
            h = R.solve(dx, F)
            its[0] = h[-1][0] if h else 0
        for _ in range(args.warmup):
            step()
        torch.cuda.synchronize()
        sampler = ClockSampler()
        sampler.start()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(args.steps):
            step()
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / args.steps
        clocks = sampler.stop()
        tF = ev_ms(torch, lambda: R.assemble(d_wg, d_dwg, F_t=F), 3)
        tJ = ev_ms(torch, lambda: R.assemble(d_wg, d_dwg, J=True), 3)
        xs = torch.randn(6 * N, dtype=torch.float64, device="cuda")
        ys = torch.zeros_like(xs)
        tmv = ev_ms(torch, lambda: R.matvec(xs, ys), 20)

        def ref_solve():
            dx.zero_()
            R.solve(dx, F)
        R.assemble(d_wg, d_dwg, F_t=F)
        R.assemble(d_wg, d_dwg, J=True)
        tS = ev_ms(torch, ref_solve, 3)          # KrylovSolve on its own (its own event pair, not a difference of medians)
        Z = int(R.spy1x1.contents.nnz)
        ab = algorithmic_bytes(N, E, Z)
        val = E / (ms * 1e-3)
        line.update({"value": val, "ms_per_step": ms, "clocks": clocks,
                     "cpu_baseline": {"value": val, "unit": UNIT, "cores": 1, "kind": "reference",
                                      "sample": f"{args.steps} full steps of the m={args.m} mesh by the reference's own CUDA build "
                                                "(oracle/_ref/libdedflow_ref.so, sm_100, D1 patched) on this GPU, driven by one host thread"},
                     "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                     "breakdown": {"assemble_F_ms": tF, "assemble_J_ms": tJ, "assemble_elems_per_s": E / ((tF + tJ) * 1e-3),
                                   "spmv_ms": tmv, "spmv_gbs_reference_format": ab["spmv_reference_format"] / (tmv * 1e-3) / 1e9,
                                   "solve_s_per_step": tS * 1e-3, "gmres_iters": its[0]}})
    else:
        # bounded sample on the host cores
        t0 = time.time()
        cb = None
        for _ in range(max(1, min(args.steps, 2))):
            cb = cpu_baseline(mesh, wg, dwg)
        ms = (time.time() - t0) / max(1, min(args.steps, 2)) * 1e3
        line.update({"value": cb["value"], "ms_per_step": ms, "cpu_baseline": cb,
                     "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
    emit(line)


# ----------------------------------------------------------------------------------------------------------------
def run_ours(args, rank, world):
    import torch
    from dedflow_b200 import boxmesh, lib as dlib
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product has no CPU path; use --impl reference for the host baseline)")
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local_rank)
    if world > 1:
        from dedflow_b200 import dist as ddist
        return ddist.bench_main(args, rank, world, local_rank, sys.modules[__name__])
    from dedflow_b200 import api
    mesh = boxmesh.make_box(args.m)
    N, E = mesh.num_node, mesh.num_tet
    t0 = time.time()
    fs = api.FlowSystem(mesh, device=f"cuda:{local_rank}")
    torch.cuda.synchronize()
    setup_s = time.time() - t0
    if args.timesteps > 0:
        fs.set_preconditioner(args.pc)
        return run_timesteps(args, fs, mesh, lambda t: t, 1, setup_s)
    Z = fs.nnz
    wg, dwg = boxmesh.state_random(N)
    h_wg = torch.from_numpy(wg).pin_memory()
    h_dwg = torch.from_numpy(dwg).pin_memory()
    h_dx = torch.zeros(6 * N, dtype=torch.float64).pin_memory()
    d_wg, d_dwg = h_wg.cuda(), h_dwg.cuda()
    F = torch.zeros(6 * N, dtype=torch.float64, device="cuda")
    dx = torch.zeros(6 * N, dtype=torch.float64, device="cuda")
    state = {"iters": 0, "hist": None}

    def step():
        fs.assemble_system(d_wg, d_dwg, F=F, mode=args.mode)
        fs.assemble_system(d_wg, d_dwg, J=True, mode=args.mode)
        dx.zero_()
        state["iters"], state["hist"] = fs.krylov_solve(dx, F)

    copy_stream = torch.cuda.Stream()
    ev_wg, ev_dwg = torch.cuda.Event(), torch.cuda.Event()

    def step_e2e():
        """the same step through the C ABI with HOST buffers.  The Jacobian only reads the velocities of wgalpha, so it is
        assembled first and the upload of dwgalpha rides under it on a copy stream; the download of dx is the tail."""
        cur = torch.cuda.current_stream()
        with torch.cuda.stream(copy_stream):
            d_wg[:3 * N].copy_(h_wg[:3 * N], non_blocking=True)      # the velocities: all the Jacobian reads
            ev_wg.record(copy_stream)
            d_wg[3 * N:].copy_(h_wg[3 * N:], non_blocking=True)
            d_dwg.copy_(h_dwg, non_blocking=True)
            ev_dwg.record(copy_stream)
        cur.wait_event(ev_wg)
        fs.assemble_system(d_wg, d_dwg, J=True, mode=args.mode)
        cur.wait_event(ev_dwg)
        fs.assemble_system(d_wg, d_dwg, F=F, mode=args.mode)
        dx.zero_()
        state["iters"], state["hist"] = fs.krylov_solve(dx, F)
        h_dx.copy_(dx, non_blocking=True)
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)      # nvidia-smi needs ~0.3 s to deliver its first sample: start it before the
    sampler.start()                         # priming so that samples exist for the (short) timed region
    # priming (NOT the warm-up): first-use work -- assembly plans, Krylov workspace, allocator pools -- and >= 1 s of load so
    # that the clocks are up; then EXACTLY --warmup untimed steps
    t_w = time.time()
    prime = 0
    while prime < 2 or time.time() - t_w < 1.0:
        step()
        prime += 1
    for _ in range(args.warmup):
        step()
    nw = args.warmup
    torch.cuda.synchronize()
    l0 = dlib.launch_count()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(args.steps):
        step()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / args.steps
    launches = dlib.launch_count() - l0
    # end-to-end through the C ABI with host buffers
    step_e2e()
    t0 = time.perf_counter()
    a.record()
    for _ in range(args.steps):
        step_e2e()
    b.record()
    torch.cuda.synchronize()
    ms_e2e = max(a.elapsed_time(b), (time.perf_counter() - t0) * 1e3) / args.steps
    clocks = sampler.stop()
    clocks["window"] = "warm-up + timed steps + e2e steps"

    # ---- breakdown: the three quantities BASELINE.json names, each event-timed on its own ----
    hbm, hbm_src = measured_hbm_peak()
    ab = algorithmic_bytes(N, E, Z)
    tF = ev_ms(torch, lambda: fs.assemble_system(d_wg, d_dwg, F=F, mode=args.mode), 10)
    tJ = ev_ms(torch, lambda: fs.assemble_system(d_wg, d_dwg, J=True, mode=args.mode), 10)
    import ctypes as C
    P = lambda t: C.c_void_p(t.data_ptr())
    st = fs._stream()
    tJk = ev_ms(torch, lambda: fs.L.dfb_assemble_tet(fs.plan, P(fs.xg), P(d_wg), P(d_dwg), None, P(fs.A00), P(fs.A01), P(fs.A10),
                                                    P(fs.A11), 1, 1, st), 10)
    tFk = ev_ms(torch, lambda: fs.L.dfb_assemble_tet(fs.plan, P(fs.xg), P(d_wg), P(d_dwg), P(F), None, None, None, None, 1, 1, st), 10)
    fs.assemble_system(d_wg, d_dwg, F=F, mode=args.mode)
    fs.assemble_system(d_wg, d_dwg, J=True, mode=args.mode)
    xs = torch.randn(6 * N, dtype=torch.float64, device="cuda")
    ys = torch.zeros_like(xs)

    def spmv_loop():
        for _ in range(20):
            fs.matrix_matvec(xs, ys)
    t_spmv_b2b = ev_ms(torch, spmv_loop, 5) / 20
    # the roofline number: every launch starts from a flushed L2 (a 256 MB buffer is read in between), which is the
    # state the mat-vec finds inside a solve, where ~230 MB of Krylov basis stream through L2 between two mat-vecs
    flush = torch.ones(64 * 1024 * 1024, dtype=torch.float32, device="cuda")   # 256 MB, READ so that L2 holds clean lines
    ts = []
    for _ in range(20):
        flush.sum()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fs.matrix_matvec(xs, ys)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t_spmv = float(np.median(ts))
    del flush

    def solve():
        dx.zero_()
        state["iters"], state["hist"] = fs.krylov_solve(dx, F)
    t_solve = ev_ms(torch, solve, 3)
    its = state["iters"]
    # the same solve run for the full 120 iterations (no stopping test)
    t_solve120, its120 = None, None
    tol, kept = (fs.atol, fs.rtol), dict(state)
    try:
        fs.atol = fs.rtol = 0.0
        t_solve120 = ev_ms(torch, solve, 2)
        its120 = int(state["iters"])
    except Exception:                              # a reported extra, never a reason to lose the line
        t_solve120 = None
    finally:
        fs.atol, fs.rtol = tol
        state.update(kept)                         # iteration count / history of the reference's stopping rule
    # Krylov algorithmic bytes with our SpMV format, n_eff = 4N live rows
    nl = 4 * N
    krylov_bytes = its * ab["spmv"] + sum(16 * nl * (j + 1) + 48 * nl for j in range(its)) + 8 * nl * its + 32 * nl
    roofs = {
        "k_spmv_fs": {"ms": t_spmv, "bytes": ab["spmv"]},
        "k_pairJ (assemble J, node pairs)": {"ms": tJk, "bytes": ab["assemble_J"]},
        "k_patchF+k_gatherF2 (assemble F, element patches)": {"ms": tFk, "bytes": ab["assemble_F"]},
        "KrylovSolve (all kernels)": {"ms": t_solve, "bytes": krylov_bytes},
    }
    for k, v in roofs.items():
        v["achieved"] = v["bytes"] / (v["ms"] * 1e-3) / 1e9
        v["frac"] = v["achieved"] / hbm
    # the assembly kernels are FP64-issue bound rather than HBM bound -- report them against the FP64 pipe as
    # well.  Algorithmic flops per element (FMA = 2): hoisted Jacobian 1.5 kflop, residual 1.6 kflop; peak = 148 SMs x 64
    # FP64 lanes x 2 x the SM clock sampled during the run.
    fp64_peak = 148 * 64 * 2 * (clocks.get("sm_mhz") or clocks.get("sm_max_mhz") or 1965.0) * 1e6 / 1e12
    for k, fl in (("k_pairJ (assemble J, node pairs)", 1500.0), ("k_patchF+k_gatherF2 (assemble F, element patches)", 1600.0)):
        v = roofs[k]
        v["fp64"] = {"flops": fl * E, "achieved_tflops": fl * E / (v["ms"] * 1e-3) / 1e12, "peak_tflops": fp64_peak,
                     "frac": fl * E / (v["ms"] * 1e-3) / 1e12 / fp64_peak}
    # dominant kernel of the step: the Krylov solve is >90% of it; inside it the SpMV, the multi-dot and the update each
    # stream comparable bytes.  The named kernel is the SpMV (the one BASELINE.json's metric quotes).
    spmv_share = its * t_spmv / ms
    traffic = None                      # dram__bytes_read+write per launch from the committed `ncu --set full` capture
    if args.m == 55:
        for name in ("r02_ncu_traffic.json", "r01_ncu_traffic.json"):
            tp = ROOT / "profiles" / name
            if tp.exists():
                try:
                    tj = json.loads(tp.read_text())
                    key = next(k for k in ("k_spmv_fs<8, 0, 1, 1>", "k_spmv_fs<8, 0, 0, 0>", "k_spmv_fs") if k in tj)   # the solver's launch
                    traffic = tj[key]["dram_bytes_per_launch"]
                    break
                except Exception:
                    traffic = None
    # per-kernel CUDA-event times of one more solve (events around every launch: a diagnostic, not the timed region)
    solve_kernels = None
    try:
        dlib.set_option("DFB_PROFILE", -1)
        solve()
        solve_kernels = dlib.solve_profile(fs.gmres)
    except Exception:
        solve_kernels = None
    finally:
        dlib.set_option("DFB_PROFILE", 0)
    # The roofline line: the mat-vec's average launch duration INSIDE a solve of the timed workload (one CUDA-event pair around
    # every launch, on the solver's stream; L2 in whatever state the multi-dot / update sweeps leave it -- the matrix alone is
    # 2.6x the L2).  The single launch after an explicit L2 flush (event pair around one launch, so it also carries the
    # launch-to-launch gap) stays next to it as `flushed`.
    t_in = None
    if solve_kernels and solve_kernels.get("spmv", {}).get("launches"):
        t_in = solve_kernels["spmv"]["avg_us"] * 1e-3
    t_roof = t_in if t_in else t_spmv
    roofline = {"kernel": "k_spmv_fs", "bound": "hbm", "achieved": ab["spmv"] / (t_roof * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s",
                "frac": ab["spmv"] / (t_roof * 1e-3) / 1e9 / hbm, "traffic": traffic, "peak_source": hbm_src,
                "bytes_per_launch": ab["spmv"], "ms_per_launch": t_roof,
                "timed": "in-solve average over %d launches" % solve_kernels["spmv"]["launches"] if t_in else "single launches, L2 flushed",
                "share_of_step": its * t_roof / ms,
                "flushed": {"ms_per_launch": t_spmv, "achieved": roofs["k_spmv_fs"]["achieved"], "frac": roofs["k_spmv_fs"]["frac"]},
                "back_to_back": {"ms_per_launch": t_spmv_b2b, "frac": ab["spmv"] / (t_spmv_b2b * 1e-3) / 1e9 / hbm},
                "achieved_reference_format": ab["spmv_reference_format"] / (t_roof * 1e-3) / 1e9}
    line = {
        "metric": METRIC, "value": E / (ms * 1e-3), "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": nw,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_text(args.m, E, N), "nnz": Z, "gmres_iters": its, "assembly_mode": args.mode,
                   "prime_steps": prime,
                   "l2": "step: working set (matrix 16*nnz*8 B = 328 MB + Krylov basis) exceeds the 126 MB L2, no explicit flush; "
                         "roofline kernel: timed inside the solve (roofline.flushed = one launch after a 256 MB L2 flush)"},
        "clocks": clocks,
        "e2e": {"value": E / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e, "h2d_bytes_per_step": 2 * 6 * N * 8,
                "d2h_bytes_per_step": 6 * N * 8 + 8 * (its + 1)},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "roofline_all": roofs,
        "breakdown": {"assemble_F_ms": tF, "assemble_J_ms": tJ, "assemble_elems_per_s": E / ((tF + tJ) * 1e-3),
                      "assemble_J_kernel_elems_per_s": E / (tJk * 1e-3), "spmv_ms": t_roof, "spmv_flushed_l2_ms": t_spmv, "spmv_back_to_back_ms": t_spmv_b2b, "spmv_gbs": roofline["achieved"],
                      "spmv_pct_hbm": 100 * roofline["frac"], "solve_s_per_step": t_solve * 1e-3, "gmres_iters": its,
                      "solve_fixed_iterations_s": None if t_solve120 is None else t_solve120 * 1e-3, "fixed_iterations": its120,
                      "setup_s": setup_s, "final_residual": float(state["hist"][-1]), "initial_residual": float(state["hist"][0]),
                      "solve_kernels": solve_kernels},
    }
    if not args.no_cpu:
        line["cpu_baseline"] = cpu_baseline(mesh, wg, dwg)
    fs.close()
    del fs, d_wg, d_dwg, F, dx, xs, ys
    torch.cuda.empty_cache()
    # BASELINE configs[2] on ONE GPU (the 1-GPU denominator of the strong-scaling blocks the N > 1 lines carry)
    if args.m == 55 and not args.no_strong:
        try:
            line["strong_16M"] = strong_one_gpu(139, local_rank, steps=3)
            if args.strong64:
                line["strong_64M"] = strong_one_gpu(220, local_rank, steps=2)
        except Exception as e:                     # a reported extra, never a reason to lose the line
            line["strong_16M"] = {"error": repr(e)[:200]}
    emit(line)


def strong_one_gpu(m, local_rank, steps=3, mesh=None):
    """One GPU, the whole m^3 box (BASELINE configs[2]/[3]), the reference's stopping rule: ms per step with its own clock sample."""
    import torch
    from dedflow_b200 import api, boxmesh
    mesh = mesh if mesh is not None else boxmesh.make_box(m)
    N, E = mesh.num_node, mesh.num_tet
    t0 = time.time()
    fs = api.FlowSystem(mesh, device=f"cuda:{local_rank}", with_colors=False)
    wg, dwg = boxmesh.state_random(N)
    d_wg, d_dwg = torch.from_numpy(wg).cuda(), torch.from_numpy(dwg).cuda()
    F = torch.zeros(6 * N, dtype=torch.float64, device="cuda")
    dx = torch.zeros(6 * N, dtype=torch.float64, device="cuda")
    st = {}

    def step():
        fs.assemble_system(d_wg, d_dwg, F=F)
        fs.assemble_system(d_wg, d_dwg, J=True)
        dx.zero_()
        st["iters"], st["hist"] = fs.krylov_solve(dx, F)
    step()
    torch.cuda.synchronize()
    setup_s = time.time() - t0
    sampler = ClockSampler(local_rank)
    sampler.start()
    step()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(steps):
        step()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / steps
    tF = ev_ms(torch, lambda: fs.assemble_system(d_wg, d_dwg, F=F), 3)
    tJ = ev_ms(torch, lambda: fs.assemble_system(d_wg, d_dwg, J=True), 3)
    clocks = sampler.stop()
    Z = fs.nnz
    fs.close()
    del fs, d_wg, d_dwg, F, dx
    torch.cuda.empty_cache()
    return {"m": m, "elems": E, "nodes": N, "nnz": Z, "n_gpus": 1, "steps": steps, "ms_per_step": ms, "elems_per_s": E / (ms * 1e-3),
            "gmres_iters": int(st["iters"]), "assemble_F_ms": tF, "assemble_J_ms": tJ, "setup_s": setup_s, "clocks": clocks}


def run_timesteps(args, fs, mesh_or_local, localize, world, setup_s, dist=None, rank=0, Eg=None, Ng=None):
    """BASELINE configs[4]: K time steps (predictor, Newton with reassembly of F and J in every nonlinear iteration,
    corrector: reference src/main.c:537-565 around SolveFlowSystem :77-283) from the reference's initial condition.
    A "step" of the JSON line is one NEWTON ITERATION (assemble J + KrylovSolve + state update + assemble F + norms)."""
    import torch
    from dedflow_b200 import boxmesh
    Eg = Eg or mesh_or_local.num_tet
    Ng = Ng or mesh_or_local.num_node
    state = [torch.from_numpy(localize(a)).cuda() for a in args._initial_state]
    fs.time_step(*state)                                    # warm-up step (plans, workspaces), then restart
    state = [torch.from_numpy(localize(a)).cuda() for a in args._initial_state]
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    sampler = ClockSampler(int(os.environ.get("LOCAL_RANK", 0)))
    if rank == 0:
        sampler.start()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    newton, gmres, last = 0, 0, None
    for _ in range(args.timesteps):
        hist = fs.time_step(*state)
        newton += len(hist) - 1
        gmres += sum(h[1] for h in hist)
        last = hist
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    if dist is not None:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
    if rank == 0:
        clocks = sampler.stop()
        line = {"metric": METRIC, "value": Eg * newton / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": newton, "warmup": 1,
                "ms_per_step": ms / newton, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic",
                "config": {"workload": f"BASELINE configs[4]: {args.timesteps} time steps on the Kuhn box m={args.m} ({Eg} tets, {Ng} nodes, "
                                       f"4 live DOF/node) from the reference's initial condition, reassembly of F and J in every Newton "
                                       f"iteration; a step = one Newton iteration", "preconditioner": getattr(args, "pc", "jacobi"),
                           "l2": "working set exceeds the 126 MB L2"},
                "clocks": clocks,
                "breakdown": {"time_steps": args.timesteps, "newton_iterations": newton, "gmres_iterations": gmres,
                              "s_per_time_step": ms * 1e-3 / args.timesteps, "setup_s": setup_s,
                              "last_newton_norms": [[float(x) for x in h[0]] for h in last]}}
        emit(line)
    fs.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--m", "--cells", dest="m", type=int, default=55,
                    help="cells per direction of the box mesh (55 -> 998,250 tets); use --cells under torchrun, which claims --m")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="gather", choices=["gather", "atomic", "colored", "auto"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--ref-port", action="store_true", help="--impl reference: force the host-core oracle port")
    ap.add_argument("--timesteps", type=int, default=0,
                    help="run BASELINE configs[4] instead: this many time steps with Newton reassembly (see run_timesteps)")
    ap.add_argument("--fixed-m", action="store_true", help="N > 1: keep --m (strong scaling) instead of growing the mesh with N")
    ap.add_argument("--pc", default="jacobi", choices=["jacobi", "schur2"],
                    help="--timesteps on one GPU: the reference's block-Jacobi (default) or the opt-in two-level Schur-complement preconditioner")
    ap.add_argument("--owner", default="auto", choices=["auto", "slab", "rcb"],
                    help="N > 1: node ownership (z-slabs, coordinate bisection; auto = slabs when the planes divide evenly, else bisection)")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling blocks (16M-tet mesh; 64M at 8 GPUs)")
    ap.add_argument("--strong64", action="store_true", help="N = 1: also run the 64M-tet mesh on one GPU (about a minute)")
    ap.add_argument("--no-parity", action="store_true", help="N > 1: skip the embedded parity check against the oracle")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    protect_stdout()
    if args.timesteps > 0:
        from dedflow_b200 import boxmesh as _bm
        args._initial_state = _bm.state_initial(_bm.make_box(args.m))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world)


if __name__ == "__main__":
    main()
