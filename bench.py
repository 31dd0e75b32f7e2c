#!/usr/bin/env python
"""Benchmark of the waveguide eigenmode hot path (BASELINE.json metric: mode solves/s, 500k-tri, order 2).

A "step" = one complete mode solve of BASELINE config 2 (Si rib waveguide, synthetic 500k-triangle
mesh, order 2 (N2xP2), num_modes=4, real epsilon): structural CSR pattern + element assembly +
nested-dissection analysis + multifrontal LU + shift-invert Arnoldi + H-field + power normalisation.

  python bench.py --gpus N --steps K --warmup W            # this framework (one solve per GPU, weak scaling)
  python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port) on host cores
  python bench.py --workload sweep ...                     # BASELINE metric 3: fixed-mesh wavelength sweep through
                                                           # femwell_b200.sweep (pattern + LU analysis reused), pts/s

`value`  : solves/s with the mesh / FE space already resident in HBM (library-internal CUDA events
           from the start of the solve until the results are complete in HBM).
`e2e`    : the same through the public Python API with HOST buffers: compute_modes(basis, eps, ...)
           including the upload of mesh/space/epsilon and the download of lams, E, H.
`roofline`: the DOMINANT kernel family of the step -- the triangular solves of the sparse LU (forward + backward
           substitution, one pair per Arnoldi step): entries of L and U read once per pair (from the analysis tables)
           x 8 B / average time of one pair.  `rooflines` adds the factorisation (FP64 flops), the element kernel K1
           and the SpMV K4 (device-resident timing), each against its own bound.
`assembly_nnz_per_s`: BASELINE metric 2, (nnz(A) + nnz(B)) of the canonical CSR / (element kernel + reduce), and the
           cold figure including the COO sort of the pattern.
The CPU reference (cpu_baseline and --impl reference) is the oracle port on a LADDER of smaller meshes of the same
geometry with a fitted power law, extrapolated to the 500k-triangle config and labelled as such: the
500k-triangle solve itself takes hours on the host.
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

METRIC = "mode solves/s (500k-tri, order 2)"
UNIT = "solves/s"


def make_problem(n, slab_h=0.11):
    """BASELINE config 2 (C2 of SURVEY.md 8d): rib waveguide on a jittered n x n grid (2 n^2 triangles)."""
    from femwell_b200.mesh import waveguide_mesh

    m = waveguide_mesh(n, slab_h=slab_h, jitter=0.15)
    eps = np.full(m.nelements, 1.444**2)
    eps[m.subdomains["core"]] = 3.4777**2
    if slab_h > 0:
        eps[m.subdomains["slab"]] = 3.4777**2
    return m, eps


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        if os.environ.get("FW_BENCH_NO_SAMPLER"):  # diagnostics only: is the polling itself perturbing the run?
            return
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200", "-i",
                 str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                smax.append(float(r[2]))
                for nme, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def algorithmic_bytes_element_kernel(ne, nv, nb, nt, sv, eps_bytes):
    """SURVEY.md 8d / DESIGN.md: t (12 B) + eps per element, p (16 B per vertex), 196+64 local entries out.
    (element_dofs are read by the symbolic sort, not by the element kernel.)"""
    return ne * (12 + eps_bytes) + nv * 16 + ne * (nb * nb + nt * nt) * sv


def run_cpu_sample(n_sample, num_modes):
    """The reference's CPU path on one bounded sample: numpy restatement of the scikit-fem assembly + the real scipy
    eigs / spsolve (oracle/femwell_oracle.py).  Returns (seconds, triangles, stage times, n_eff)."""
    from femwell_b200.element import triangle_quadrature
    from oracle import femwell_oracle as fo

    m, eps = make_problem(n_sample)
    X, W = triangle_quadrature(4)
    N = m.nvertices + 3 * m.nfacets + 2 * m.nelements
    v0 = np.random.default_rng(0).uniform(-1, 1, N)
    t0 = time.perf_counter()
    r = fo.compute_modes(m.p, m.t, 0, eps, 1.55, X, W, num_modes=num_modes, order=2, v0=v0)
    dt = time.perf_counter() - t0
    return dt, int(m.nelements), r.timings, [float(x.real) for x in r.n_effs]


def ncu_traffic(csv_name, kernel_substr, n_expected=None, total=False, profiles_dir=None):
    """DRAM bytes (read + write) per launch of a kernel from a committed ncu summary (profiles/*.csv: a comment line, the
    header, the unit row, one row per launch -- the format scripts/ncu_raw_to_summary.py writes), or None;
    total=True: the sum over all rows (a launch UNIT made of many kernels, e.g. one substitution pair)"""
    try:
        import csv

        path = os.path.join(profiles_dir or os.path.join(ROOT, "profiles"), csv_name)
        rows = [r for r in csv.reader(open(path)) if r and not r[0].startswith("#")]
        hdr, units = rows[0], rows[1]
        ir, iw = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        sel = [r for r in rows[2:] if kernel_substr in r[0] or (len(r) > 1 and kernel_substr in r[1])]
        if not sel:
            return None
        tot = [float(r[ir]) * scale.get(units[ir], 1.0) + float(r[iw]) * scale.get(units[iw], 1.0) for r in sel]
        return float(np.sum(tot)) if total else float(np.mean(tot))
    except Exception:
        return None


def fit_power_law(ne, seconds):
    """least squares of log t = log c + p log Ne"""
    x, y = np.log(np.asarray(ne, dtype=float)), np.log(np.asarray(seconds, dtype=float))
    if len(x) < 2:
        return 1.0, float(np.exp(y[0] - x[0]))
    p, logc = np.polyfit(x, y, 1)
    return float(p), float(np.exp(logc))


def extrapolate_ladder(ladder, full_ne):
    """ladder: list of (triangles, seconds).  Returns (seconds at full_ne by the fitted power law anchored at the
    largest sample, exponent, seconds by linear scaling of the largest sample = a lower bound favourable to the CPU)."""
    ladder = sorted(ladder)
    ne = [a for a, _ in ladder]
    sec = [b for _, b in ladder]
    p, _c = fit_power_law(ne, sec)
    p_used = max(p, 1.0)  # never extrapolate sub-linearly
    t_full = sec[-1] * (full_ne / ne[-1]) ** p_used
    t_lin = sec[-1] * (full_ne / ne[-1])
    return t_full, p, t_lin


def cpu_threads():
    try:
        from threadpoolctl import threadpool_info

        return max([p.get("num_threads", 1) for p in threadpool_info()] + [1])
    except Exception:
        return os.cpu_count() or 1


def workload_text(n, num_modes):
    ne = 2 * n * n
    return (f"BASELINE config 2: Si rib waveguide 500x220 nm + 110 nm slab in SiO2, lambda=1.55 um, jittered structured "
            f"mesh {n}x{n}x2 = {ne} triangles, order 2 (N2xP2), num_modes={num_modes}, real eps, 6-point quadrature "
            f"(intorder 4)")


def reference_arm(args, full_ne):
    """--impl reference: the oracle port on a ladder of smaller meshes (the 500k-triangle solve takes hours on the host).
    Warm-up steps walk the ladder (each size once), the timed steps repeat the anchor size; the fitted power law
    extrapolates the anchor to the full config.  Everything about the sample is in the JSON line."""
    ladder_n = [int(x) for x in args.cpu_ladder.split(",")]
    anchor = args.cpu_sample_n
    ladder, stage_info, n_eff = [], None, None
    t_start = time.perf_counter()
    for n in ladder_n:
        if n == anchor:
            continue
        dt, ne, _stages, _ = run_cpu_sample(n, args.num_modes)
        ladder.append((ne, dt))
        print(f"[bench/reference] ladder {ne} triangles: {dt:.2f} s", file=sys.stderr, flush=True)
    times, ne_anchor = [], None
    budget_s = args.cpu_budget_s
    for i in range(max(1, args.steps)):
        dt, ne_anchor, stage_info, n_eff = run_cpu_sample(anchor, args.num_modes)
        times.append(dt)
        print(f"[bench/reference] step {i}: {ne_anchor} triangles {dt:.2f} s", file=sys.stderr, flush=True)
        if time.perf_counter() - t_start > budget_s:  # keep the whole run within a few minutes
            break
    t_anchor = float(np.mean(times))
    ladder.append((ne_anchor, t_anchor))
    t_full, p, t_lin = extrapolate_ladder(ladder, full_ne)
    value = 1.0 / t_full
    ladder_txt = ", ".join(f"{a} tri: {b:.2f} s" for a, b in sorted(ladder))
    sample = (f"full oracle compute_modes (numpy scikit-fem-equivalent assembly + scipy eigs (SuperLU+ARPACK) + "
              f"{args.num_modes} spsolve H-fields) on a ladder of meshes of the same geometry ({ladder_txt}); fitted "
              f"t ~ Ne^{p:.2f}; the {ne_anchor}-triangle time extrapolated to {full_ne} triangles: {t_full:.0f} s "
              f"(linear scaling, a lower bound: {t_lin:.0f} s)")
    config = {"workload": (f"EXTRAPOLATED to BASELINE config 2 ({full_ne} triangles) from measured CPU solves of the same "
                           f"rib geometry at {ladder_txt}; timed steps ran the {ne_anchor}-triangle sample, order 2, "
                           f"num_modes={args.num_modes}, real eps, intorder 4"),
              "target_workload": workload_text(args.n, args.num_modes),
              "sweep": "single solve", "l2": "n/a (host)"}
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": len(times), "warmup": len(ladder) - 1, "ms_per_step": 1000.0 * t_anchor,
            "ms_per_step_extrapolated": 1000.0 * t_full, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
            "extrapolated": True, "sample_ne": [a for a, _ in sorted(ladder)],
            "sample_seconds": [b for _, b in sorted(ladder)], "exponent": p,
            "value_linear_lower_bound_time": 1.0 / t_lin, "sample_stages_s": stage_info, "sample_n_eff": n_eff,
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cpu_threads(), "kind": "port", "sample": sample,
                             "extrapolated": True, "exponent": p},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="solve", choices=["solve", "sweep", "geometry"])
    ap.add_argument("--geometry-n", type=int, default=224, help="--workload geometry: grid cells per side of every mesh")
    ap.add_argument("--n", type=int, default=500, help="grid cells per side (2 n^2 triangles)")
    ap.add_argument("--num-modes", type=int, default=4)
    ap.add_argument("--points-per-rank", type=int, default=4, help="--workload sweep: wavelength points per GPU per step")
    ap.add_argument("--cpu-sample-n", type=int, default=71, help="anchor size of the CPU ladder (2 n^2 triangles)")
    ap.add_argument("--cpu-ladder", default="36,50,71", help="CPU ladder sizes n (reference arm default adds 100)")
    ap.add_argument("--cpu-budget-s", type=float, default=150.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    full_ne = 2 * args.n * args.n

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return
        if args.cpu_ladder == "36,50,71":
            args.cpu_ladder = "36,50,71,100"
        reference_arm(args, full_ne)
        return

    # ------------------------------------------------------------------ this framework (GPU)
    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the b200 path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    # the host analysis of the sparse LU is threaded: share the cores between the ranks of the node
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
    if "FW_HOST_THREADS" not in os.environ:
        os.environ["FW_HOST_THREADS"] = str(max(2, min(16, (os.cpu_count() or 16) // max(1, local_world))))

    from femwell_b200.basis import Basis
    from femwell_b200.element import ElementTriP0
    from femwell_b200.maxwell import waveguide as wg

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    m, eps = make_problem(args.n)
    b0 = Basis(m, ElementTriP0(), intorder=4)
    ctx = wg.get_context(local_rank)

    if args.workload == "sweep":
        run_sweep_workload(args, rank, world, local_rank, m, eps, b0, ctx, barrier, torch, dist)
        return
    if args.workload == "geometry":
        run_geometry_workload(args, rank, world, local_rank, ctx, barrier, torch, dist)
        return

    config = {"workload": workload_text(args.n, args.num_modes),
              "sweep": "one independent solve per GPU per step" if world > 1 else "single solve",
              "l2": "working set (>20 GB of fronts) far exceeds the 126 MB L2; no explicit flush needed"}
    # every rank solves its own sweep point (a slightly different wavelength), as in a wavelength sweep
    wavelength = 1.55 + 0.001 * rank

    def one_step():
        # cold e2e step: the space is re-uploaded and every symbolic/numeric stage re-run
        ctx._space_key = None
        t0 = time.perf_counter()
        modes = wg.compute_modes(b0, eps, wavelength, num_modes=args.num_modes, order=2)
        t1 = time.perf_counter()
        return modes, t1 - t0

    stats = None
    for _ in range(args.warmup):
        modes, _dt = one_step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    dev_ms, e2e_s, elem_ms, launches = [], [], [], 0
    solve_ms, n_ops, factor_ms, reduce_ms, symbolic_ms = [], [], [], [], []
    barrier()
    t_begin = time.perf_counter()
    for _ in range(args.steps):
        modes, dt = one_step()
        tim = modes.stats["timings"]
        dev_ms.append(tim["total_ms"])
        if rank == 0:
            print("[bench] step stages (ms): " + ", ".join(f"{k[:-3]} {v:.1f}" for k, v in tim.items() if k.endswith("_ms")),
                  file=sys.stderr, flush=True)
        elem_ms.append(tim["element_kernel_ms"])
        reduce_ms.append(tim["reduce_ms"])
        symbolic_ms.append(tim["symbolic_ms"])
        factor_ms.append(tim["lu_factor_ms"])
        solve_ms.append(modes.stats["t_solve_ms_total"])
        n_ops.append(modes.stats["n_op"])
        e2e_s.append(dt)
        launches += tim["kernel_launches"]
        stats = modes.stats
    barrier()
    t_end = time.perf_counter()
    clocks = sampler.stop() if rank == 0 else None

    # max over ranks
    t = torch.tensor([sum(dev_ms) / 1000.0, t_end - t_begin], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_total_s, wall_total_s = t.tolist()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    basis = modes[0].basis
    N = basis.N
    k = args.num_modes
    value = world * args.steps / dev_total_s
    e2e_value = world * args.steps / wall_total_s
    h2d = m.p.nbytes + m.t.nbytes + basis.element_dofs.nbytes + eps.nbytes + N  # + is_z bytes
    d2h = k * 16 + 2 * k * N * 16 + k * 16
    sv = 8
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = peaks.get("hbm_gbs", 6650.0)
    peak_src = "MEASURED_PEAKS.json hbm_gbs (burst copy)" if peaks else "fallback 6.65 TB/s (B200_PROFILING.md)"
    tt = stats["timings"]

    # ---- dominant kernel family: the triangular solves of the sparse LU (one forward + backward pair per OP application)
    solve_pairs = float(np.sum(n_ops))
    t_pair = float(np.sum(solve_ms)) / 1000.0 / max(1.0, solve_pairs)
    solve_bytes = float(tt["lu_solve_entries"]) * sv
    solve_achieved = solve_bytes / t_pair / 1e9 if t_pair > 0 else 0.0
    roofline = {"kernel": "sparse-LU triangular solves (k_fwd/bwd_chain_dinv on the small-front levels, k_fwd/bwd_update, "
                          "k_tinv_matvec on the inverse pivot blocks of the top levels): one forward + backward "
                          "substitution pair per Arnoldi step, ~85 kernel nodes replayed as one CUDA graph",
                "bound": "hbm", "achieved": solve_achieved, "peak": peak, "unit": "GB/s", "frac": solve_achieved / peak,
                "traffic": ncu_traffic("ncu_r02_final_solve_pair.csv", "k_", total=True),
                "algorithmic_bytes_per_launch": solve_bytes, "seconds_per_launch": t_pair,
                "launch_unit": "one solve pair (all its kernels); L and U entries of every front read once: "
                               "sum(m^2 - (m-ns)^2) x 8 B from the analysis tables",
                "pairs_per_solve": solve_pairs / args.steps, "share_of_step": float(np.sum(solve_ms)) / float(np.sum(dev_ms)),
                "peak_source": peak_src}
    # ---- the other kernels, each against its own bound
    alg_k1 = algorithmic_bytes_element_kernel(m.nelements, m.nvertices, 14, 8, sv, 8)
    t_elem = float(np.mean(elem_ms)) / 1000.0
    n_k1 = int(tt.get("element_kernel_launches", 1))
    # TFLOP/s: measured on this pool with scripts/microbench/fp64_peak.cu (DMMA m8n8k4 from registers, 37.1; DFMA 33.9;
    # profiles/fp64_peak_r02.log) -- MEASURED_PEAKS.json has no FP64 entry
    fp64_peak = 37.1
    t_fac = float(np.mean(factor_ms)) / 1000.0
    rooflines = {
        "element_kernel_K1": {"kernel": "k_element_blocks_const" if n_k1 == 1 else "k_element_blocks (3 launches)",
                              "bound": "hbm", "achieved": alg_k1 / t_elem / 1e9, "peak": peak, "unit": "GB/s",
                              "frac": alg_k1 / t_elem / 1e9 / peak, "traffic": ncu_traffic("ncu_r01c_assembly_full.csv",
                                                                                         "k_element_blocks_const"),
                              "algorithmic_bytes_per_launch": alg_k1 / n_k1, "seconds_per_launch": t_elem / n_k1,
                              "share_of_step": t_elem * 1000.0 / float(np.mean(dev_ms))},
        "lu_factor": {"kernel": "multifrontal LU numeric phase (panel / TRSM / DMMA GEMM / extend-add kernels, pivot-block "
                                "inverses)",
                      "bound": "fp64", "achieved": tt["lu_factor_flops"] / t_fac / 1e12 if t_fac > 0 else 0.0,
                      "peak": fp64_peak, "unit": "TFLOP/s",
                      "frac": (tt["lu_factor_flops"] / t_fac / 1e12 / fp64_peak) if t_fac > 0 else 0.0,
                      "flops_per_factorisation": tt["lu_factor_flops"], "seconds": t_fac,
                      "peak_source": "measured FP64 tensor (DMMA) peak, profiles/fp64_peak_r02.log (no FP64 entry in "
                                     "MEASURED_PEAKS.json)",
                      "share_of_step": t_fac * 1000.0 / float(np.mean(dev_ms))},
    }
    try:
        ms_spmv = ctx.bench_spmv(0, False, 20)
        nnz_s = int(tt["nnz_structural"])
        spmv_bytes = nnz_s * (sv + 4) + (N + 1) * 4 + 2 * N * sv
        rooflines["spmv_K4"] = {"kernel": "k_spmv<double,double,8> on the structural CSR of A (device resident, 20 reps)",
                                "bound": "hbm", "achieved": spmv_bytes / (ms_spmv / 1e3) / 1e9, "peak": peak,
                                "unit": "GB/s", "frac": spmv_bytes / (ms_spmv / 1e3) / 1e9 / peak,
                                "traffic": ncu_traffic("ncu_r02_spmv_full.csv", "k_spmv"),
                                "algorithmic_bytes_per_launch": spmv_bytes, "seconds_per_launch": ms_spmv / 1e3}
    except Exception as exc:  # noqa: BLE001
        rooflines["spmv_K4"] = {"error": str(exc)}
    # ---- BASELINE metric 2: assembly nnz/s with nnz counted in the final canonical CSR (after eliminate_zeros)
    assembly = None
    try:
        nnzA, nnzB = ctx.matrix_nnz(0), ctx.matrix_nnz(1)
        t_asm = (float(np.mean(elem_ms)) + float(np.mean(reduce_ms))) / 1000.0
        t_cold = t_asm + float(np.mean(symbolic_ms)) / 1000.0
        assembly = {"value": (nnzA + nnzB) / t_asm, "unit": "nnz/s", "nnz_A": int(nnzA), "nnz_B": int(nnzB),
                    "seconds": t_asm, "what": "element kernel + segmented reduce onto the resident pattern (sweeps)",
                    "value_cold": (nnzA + nnzB) / t_cold, "seconds_cold": t_cold,
                    "what_cold": "+ COO sort / pattern build of a new mesh (symbolic_ms)"}
    except Exception as exc:  # noqa: BLE001
        assembly = {"error": str(exc)}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1000.0 * dev_total_s / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": 1000.0 * wall_total_s / args.steps},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "rooflines": rooflines,
            "stages_ms": {kk: vv for kk, vv in stats["timings"].items()},
            "solver": {kk: vv for kk, vv in stats.items() if kk != "timings"},
            "n_eff": [float(x.real) for x in modes.n_effs],
            "step_ms": [float(x) for x in dev_ms],
            "assembly_nnz_per_s": assembly}
    if not args.no_cpu_baseline and world == 1:
        ladder = []
        stages = None
        for n in [int(x) for x in args.cpu_ladder.split(",")]:
            dt, ne, stages, _ = run_cpu_sample(n, args.num_modes)
            ladder.append((ne, dt))
        t_full, p, t_lin = extrapolate_ladder(ladder, full_ne)
        ladder_txt = ", ".join(f"{a} tri: {b:.2f} s" for a, b in sorted(ladder))
        line["cpu_baseline"] = {
            "value": 1.0 / t_full, "unit": UNIT, "cores": cpu_threads(), "kind": "port", "extrapolated": True,
            "exponent": p, "sample_ne": [a for a, _ in sorted(ladder)], "sample_seconds": [b for _, b in sorted(ladder)],
            "value_linear_lower_bound_time": 1.0 / t_lin,
            "sample": (f"oracle compute_modes on a ladder of meshes of the same geometry ({ladder_txt}; stages of the largest: "
                       f"{ {a: round(b, 2) for a, b in stages.items()} }); fitted t ~ Ne^{p:.2f}, the largest sample "
                       f"extrapolated to {full_ne} triangles: {t_full:.0f} s (linear scaling, a lower bound favourable to "
                       f"the CPU: {t_lin:.0f} s)")}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run_sweep_workload(args, rank, world, local_rank, m, eps, b0, ctx, barrier, torch, dist):
    """BASELINE metric 3 (sweep pts/s at 1-8 GPU), config-5 style: a fixed-mesh wavelength sweep through
    femwell_b200.sweep.wavelength_sweep -- block partition over the ranks, CSR pattern and LU analysis reused between the
    points of a rank, one all_gather of the per-point n_eff at the end of every step."""
    from femwell_b200.sweep import group_index, wavelength_sweep

    ppr = args.points_per_rank
    n_pts = ppr * world
    wls = np.linspace(1.50, 1.60, n_pts)

    def eps_of(wl):  # a smooth material dispersion (Sellmeier-like slope), vary_wavelength.py:41-53
        return eps * (1.0 - 0.05 * (wl - 1.55))

    def one_step():
        t0 = time.perf_counter()
        table = wavelength_sweep(b0, eps_of, wls, num_modes=1, order=2)
        return table, time.perf_counter() - t0

    for _ in range(max(1, args.warmup)):
        table, _dt = one_step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_begin = time.perf_counter()
    ev0.record()
    launches = 0
    for _ in range(args.steps):
        table, _dt = one_step()
        launches += ctx.timings()["kernel_launches"] * ppr
    ev1.record()
    barrier()
    t_end = time.perf_counter()
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([t_end - t_begin], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    wall = float(t.item())
    if rank == 0:
        tim = ctx.timings()
        value = n_pts * args.steps / wall
        ng = group_index(wls, table[:, 0].real) if n_pts >= 3 else None
        line = {"metric": "sweep pts/s at 1-8 GPU", "value": value, "unit": "points/s", "n_gpus": world,
                "steps": args.steps, "warmup": max(1, args.warmup), "ms_per_step": 1000.0 * wall / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": (f"fixed-mesh wavelength sweep (config-5 style): {ppr} points per GPU per step on the "
                                        f"{args.n}x{args.n}x2 = {2 * args.n * args.n} triangle rib mesh, order 2, num_modes=1, "
                                        "pattern + LU analysis reused between points, results all_gathered per step"),
                           "points_per_step": n_pts},
                "e2e": {"value": value, "unit": "points/s", "h2d_bytes_per_step": int(eps.nbytes * ppr),
                        "d2h_bytes_per_step": int(2 * 16 * modes_n(ctx) * ppr)},
                "gpu_launches": int(launches), "clocks": clocks,
                "last_point_stages_ms": {kk: vv for kk, vv in tim.items()},
                "n_eff": [float(x) for x in table[:, 0].real], "group_index": None if ng is None else [float(x) for x in ng]}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run_geometry_workload(args, rank, world, local_rank, ctx, barrier, torch, dist):
    """BASELINE metric 3, config-4 style: coupled-waveguide gap x width sweep through
    femwell_b200.sweep.coupler_geometry_sweep -- every geometry has its OWN mesh (~100k triangles), 2 solves (3 modes),
    2 same-mesh and 2 cross-mesh overlaps; block partition over the ranks, one all_gather per step."""
    from femwell_b200.sweep import COUPLER_COLUMNS, coupler_geometry_sweep

    ppr = args.points_per_rank
    n_pts = ppr * world
    gaps = np.linspace(0.1, 0.6, n_pts)
    widths = [0.45]
    n = args.geometry_n

    def one_step():
        return coupler_geometry_sweep(gaps, widths, n=n, order=2)

    for _ in range(max(1, args.warmup)):
        table = one_step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    t_begin = time.perf_counter()
    for _ in range(args.steps):
        table = one_step()
    barrier()
    t_end = time.perf_counter()
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([t_end - t_begin], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    wall = float(t.item())
    if rank == 0:
        tim = ctx.timings()
        value = n_pts * args.steps / wall
        line = {"metric": "sweep pts/s at 1-8 GPU", "value": value, "unit": "points/s", "n_gpus": world,
                "steps": args.steps, "warmup": max(1, args.warmup), "ms_per_step": 1000.0 * wall / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": (f"coupled-waveguide geometry sweep (config-4 style): {ppr} geometries per GPU per step, "
                                        f"each its own {n}x{n}x2 = {2 * n * n} triangle mesh, order 2, supermodes (k=2) + "
                                        "isolated mode (k=1), 2 same-mesh + 2 cross-mesh overlaps, host mesh generation and "
                                        "uploads inside the timed region, results all_gathered per step"),
                           "points_per_step": n_pts},
                "e2e": {"value": value, "unit": "points/s", "h2d_bytes_per_step": None, "d2h_bytes_per_step": None},
                "gpu_launches": int(tim["kernel_launches"]) * 2 * ppr * args.steps, "clocks": clocks,
                "last_solve_stages_ms": {kk: vv for kk, vv in tim.items()},
                "columns": list(COUPLER_COLUMNS), "table": [[float(x) for x in row] for row in table]}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def modes_n(ctx):
    return int(ctx.n)


if __name__ == "__main__":
    main()
