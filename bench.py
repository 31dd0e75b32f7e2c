#!/usr/bin/env python
"""Benchmark of the waveguide eigenmode hot path (BASELINE.json metric: mode solves/s, 500k-tri, order 2).

A "step" = one complete mode solve of BASELINE config 2 (Si rib waveguide, synthetic 500k-triangle
mesh, order 2 (N2xP2), num_modes=4, real epsilon): structural CSR pattern + element assembly +
nested-dissection analysis + multifrontal LU + shift-invert Arnoldi + H-field + power normalisation.

  python bench.py --gpus N --steps K --warmup W            # this framework (one solve per GPU, weak scaling)
  python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port) on host cores

`value`  : solves/s with the mesh / FE space already resident in HBM (library-internal CUDA events
           from the start of the solve until the results are complete in HBM).
`e2e`    : the same through the public Python API with HOST buffers: compute_modes(basis, eps, ...)
           including the upload of mesh/space/epsilon and the download of lams, E, H.
`roofline`: K1 element kernel (assembly of the local blocks), algorithmic bytes / CUDA-event time.
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


def run_cpu_baseline(n_sample, num_modes, full_ne):
    """The reference's CPU path on a bounded sample: numpy restatement of the scikit-fem assembly + the
    real scipy eigs / spsolve (oracle/femwell_oracle.py).  Returns (seconds for the sample, info)."""
    from femwell_b200.element import triangle_quadrature
    from oracle import femwell_oracle as fo

    m, eps = make_problem(n_sample)
    X, W = triangle_quadrature(4)
    N = m.nvertices + 3 * m.nfacets + 2 * m.nelements
    v0 = np.random.default_rng(0).uniform(-1, 1, N)
    t0 = time.perf_counter()
    r = fo.compute_modes(m.p, m.t, 0, eps, 1.55, X, W, num_modes=num_modes, order=2, v0=v0)
    dt = time.perf_counter() - t0
    # conservative (favourable to the CPU) extrapolation: linear in the number of triangles, although the
    # measured SuperLU/COLAMD scaling on this pencil is ~N^1.8
    scale = full_ne / m.nelements
    return dt, {"sample_ne": int(m.nelements), "sample_seconds": dt, "scale": scale, "stages": r.timings,
                "n_eff": [float(x.real) for x in r.n_effs]}


def cpu_threads():
    try:
        from threadpoolctl import threadpool_info

        return max([p.get("num_threads", 1) for p in threadpool_info()] + [1])
    except Exception:
        return os.cpu_count() or 1


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", type=int, default=500, help="grid cells per side (2 n^2 triangles)")
    ap.add_argument("--num-modes", type=int, default=4)
    ap.add_argument("--cpu-sample-n", type=int, default=71)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    full_ne = 2 * args.n * args.n
    config = {"workload": f"BASELINE config 2: Si rib waveguide 500x220 nm + 110 nm slab in SiO2, lambda=1.55 um, "
                          f"jittered structured mesh {args.n}x{args.n}x2 = {full_ne} triangles, order 2 (N2xP2), "
                          f"num_modes={args.num_modes}, real eps, 6-point quadrature (intorder 4)",
              "sweep": "one independent solve per GPU per step" if world > 1 else "single solve",
              "l2": "working set (>20 GB of fronts) far exceeds the 126 MB L2; no explicit flush needed"}

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return
        times = []
        info = None
        for i in range(args.warmup + args.steps):
            dt, info = run_cpu_baseline(args.cpu_sample_n, args.num_modes, full_ne)
            if i >= args.warmup:
                times.append(dt)
            if sum(times) > 240:  # keep the whole run within a few minutes
                break
        t_sample = float(np.mean(times)) if times else info["sample_seconds"]
        value = 1.0 / (t_sample * info["scale"])
        sample = (f"full oracle compute_modes (numpy scikit-fem-equivalent assembly + scipy eigs (SuperLU+ARPACK) + "
                  f"{args.num_modes} spsolve H-fields) on a {info['sample_ne']}-triangle mesh of the same geometry: "
                  f"{t_sample:.2f} s, scaled LINEARLY by {info['scale']:.1f}x to {full_ne} triangles (conservative: "
                  f"measured scaling is ~N^1.8)")
        line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
                "steps": len(times), "warmup": args.warmup, "ms_per_step": 1000.0 / value, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": value, "unit": UNIT, "cores": cpu_threads(), "kind": "port",
                                 "sample": sample},
                "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return

    # ------------------------------------------------------------------ this framework (GPU)
    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the b200 path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from femwell_b200.basis import Basis
    from femwell_b200.element import ElementTriP0
    from femwell_b200.maxwell import waveguide as wg

    m, eps = make_problem(args.n)
    b0 = Basis(m, ElementTriP0(), intorder=4)
    ctx = wg.get_context(local_rank)
    # every rank solves its own sweep point (a slightly different wavelength), as in a wavelength sweep
    wavelength = 1.55 + 0.001 * rank

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

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
    alg_bytes = algorithmic_bytes_element_kernel(m.nelements, m.nvertices, 14, 8, sv, 8)
    t_elem = float(np.mean(elem_ms)) / 1000.0
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = peaks.get("hbm_gbs", 6650.0)
    achieved = alg_bytes / t_elem / 1e9
    # DRAM traffic of the same launch from the committed ncu --set full capture (bytes per launch)
    traffic = None
    n_k1 = int(stats["timings"].get("element_kernel_launches", 1))
    try:
        import csv

        rows = [r for r in csv.reader(open(os.path.join(ROOT, "profiles", "ncu_r01c_assembly_full.csv")))
                if r and not r[0].startswith("#")]
        hdr = rows[0]
        ir, iw = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        k1 = [r for r in rows[2:] if "k_element_blocks_const" in r[1]][:1]
        if n_k1 == 1 and k1:
            traffic = sum(float(r[ir]) * 1e9 + float(r[iw]) * 1e6 for r in k1)
    except Exception:
        pass
    kname = ("k_element_blocks_const (K1: aform + bform local blocks of every triangle, 1 launch per solve)"
             if n_k1 == 1 else "k_element_blocks (K1, quadrature-loop variant: 3 launches per solve)")
    roofline = {"kernel": kname,
                "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "launches_per_solve": n_k1,
                "algorithmic_bytes_per_launch": alg_bytes / n_k1, "seconds_per_launch": t_elem / n_k1,
                "algorithmic_bytes_per_solve": alg_bytes, "seconds_per_solve": t_elem,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (burst copy)" if peaks else "fallback 6.65 TB/s"}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1000.0 * dev_total_s / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": 1000.0 * wall_total_s / args.steps},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline,
            "stages_ms": {kk: vv for kk, vv in stats["timings"].items()},
            "solver": {kk: vv for kk, vv in stats.items() if kk != "timings"},
            "n_eff": [float(x.real) for x in modes.n_effs],
            "step_ms": [float(x) for x in dev_ms],
            "assembly_nnz_per_s": None}
    # assembly nnz/s (BASELINE.json metric 2): structural nnz(A)+nnz(B_tt) / (element kernel + reduce)
    try:
        tt = stats["timings"]
        line["assembly_nnz_per_s"] = {"nnz_structural_A": int(ctx.nnz_structural or 0),
                                      "element_plus_reduce_ms": tt["element_kernel_ms"] + tt["reduce_ms"],
                                      "symbolic_ms": tt["symbolic_ms"]}
    except Exception:
        pass
    if not args.no_cpu_baseline and world == 1:
        dt, info = run_cpu_baseline(args.cpu_sample_n, args.num_modes, full_ne)
        v = 1.0 / (dt * info["scale"])
        line["cpu_baseline"] = {
            "value": v, "unit": UNIT, "cores": cpu_threads(), "kind": "port",
            "sample": (f"oracle compute_modes on a {info['sample_ne']}-triangle mesh of the same geometry: {dt:.2f} s "
                       f"({ {a: round(b, 2) for a, b in info['stages'].items()} }), scaled linearly by "
                       f"{info['scale']:.1f}x to {full_ne} triangles (conservative; measured scaling ~N^1.8)")}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
