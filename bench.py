#!/usr/bin/env python
"""bench.py -- CEED BP GDOF/s (DOFs x CG iterations / s) of the B200 partial-assembly path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
                    [--bp 1..6] [--p P] [--n ELEMS_PER_SIDE_PER_GPU] [--iters I]

One "step" = one BP solve: Jacobi-preconditioned CG with a FIXED number of iterations
(rel_tol = 0, the protocol of tests/benchmarks/bench_ceed.cpp:222-243) on a synthetic
structured hex mesh of the unit cube, all-boundary Dirichlet, RHS (1, phi_i), x0 = 0.
Default workload (BASELINE.json configs[2], the config the metric is quoted on at one
GPU): BP3 diffusion, p = 4, Q = 6, 133^3 elements = 151.4 M dofs.  With --gpus N > 1
(configs[3]) each rank owns a 116^3-element block (100.5 M dofs / GPU, weak scaling) of a
2x1x1 / 2x2x1 / 2x2x2 Cartesian partition; shared dofs go over peer-memory stores (CUDA IPC
over NVLink; B2F_HALO=nccl selects ncclSend/Recv) and the PCG dot products over
ncclAllReduce.  --bp 2/4/6 are the vector (vdim = 3, byVDIM) forms.

Prints ONE JSON line (rank 0).  `value`: device-resident solves, CUDA-event timed, max
over ranks.  `e2e`: the same solves through the public API with HOST buffers -- the RHS
is copied from pinned host memory and the solution copied back inside the timed region.
`roofline`: the fused operator kernel alone (CUDA events recorded around each of its
launches inside the timed region) against MEASURED_PEAKS.json's HBM copy bandwidth.
`cpu_baseline`: the unmodified reference's CPU PA path (oracle/_ref/ref_driver, OpenMP
device over all host cores) on a bounded sample of the same workload.

--impl reference times that reference CPU path alone (the oracle side: the only place
besides tests/ and smoke() that executes anything under oracle/).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "CEED BP GDOF/s (DOFs x CG iterations / s), PA + PCG/Jacobi"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--bp", type=int, default=3, choices=[1, 2, 3, 4, 5, 6],
                    help="CEED bake-off problem; 2/4/6 are the vector (vdim=3, byVDIM) forms of 1/3/5")
    ap.add_argument("--p", type=int, default=None)
    ap.add_argument("--n", type=int, default=None, help="elements per side per GPU")
    ap.add_argument("--iters", type=int, default=32)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the reference-checksum parity check")
    ap.add_argument("--no-gpu-baseline", action="store_true",
                    help="skip timing the reference's own CUDA backend (oracle/_ref_cuda) on this GPU")
    ap.add_argument("--no-sweep", action="store_true", help="skip the BP1/BP3/BP5 x p=1..8 table after the headline")
    ap.add_argument("--sweep-dofs", type=float, default=50e6, help="dofs per sweep case (BASELINE config C2: ~50 M)")
    ap.add_argument("--no-lor", action="store_true", help="skip the LOR matrix assembly record (SURVEY 8f row 4)")
    ap.add_argument("--no-anchor", action="store_true", help="skip the 1-GPU run at the per-GPU size of --gpus N > 1")
    ap.add_argument("--cpu-n", type=int, default=None, help="elements per side of the CPU sample")
    ap.add_argument("--curved", type=int, default=None,
                    help="1: order-3 nodes displaced by the smooth map of BASELINE config C5 (default for --bp 5)")
    return ap.parse_args()


def workload(a):
    vdim = 3 if a.bp in (2, 4, 6) else 1
    sbp = a.bp - 1 if vdim == 3 else a.bp  # the scalar problem with the same integrator
    p = a.p or {1: 4, 3: 4, 5: 7}[sbp]
    if a.n:
        n = a.n
    elif a.gpus == 1:
        n = {1: round((50e6 / vdim) ** (1 / 3) / p),
             3: 133 if (p == 4 and vdim == 1) else round((150e6 / vdim) ** (1 / 3) / p),
             5: round((99e6 / vdim) ** (1 / 3) / p)}[sbp]
    else:
        n = round((100e6 / vdim) ** (1 / 3) / p)
    grid = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}[a.gpus]
    kind = {1: "mass", 3: "diffusion", 5: "diffusion"}[sbp]
    q = p + 1 if sbp == 5 else p + 2
    curved = bool(a.curved) if a.curved is not None else sbp == 5
    return dict(bp=a.bp, p=p, n=n, grid=grid, kind=kind, q1d=q, d1d=p + 1, gll=(sbp == 5), curved=curved,
                vdim=vdim)


def alg_bytes_per_apply(w, ne, nd):
    """fused L->L operator, stored D, int32 gather index."""
    k = 1 if w["kind"] == "mass" else 6
    return (k * w["q1d"] ** 3 * 8 + 4 * w["d1d"] ** 3) * ne + 16 * nd


class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.FIELDS,
                 "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        pw = [float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace(".", "", 1).isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [nm for k, nm in enumerate(names)
                   if any(len(r) > 4 + k and r[4 + k].lower() == "active" for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(self.rows), "reasons": reasons}


def ref_driver_path():
    p = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
    return p if os.path.isfile(p) and os.access(p, os.X_OK) else None


def cpu_sample_n(p):
    """elements per side of the CPU sample: ~17 M dofs (a 4-5 s solve at p=4 on 16 cores)"""
    return max(8, round(256 / p))


def run_reference(w, n_cpu, iters, reps, skip=0):
    """The unmodified reference's CPU PA path on all host cores (MFEM_DEVICE=omp), or the C
    oracle port on one core when oracle/_ref was not built."""
    cores = os.cpu_count() or 1
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        pass
    drv = ref_driver_path()
    if drv:
        env = dict(os.environ, MFEM_DEVICE="omp", OMP_NUM_THREADS=str(cores), OMP_PROC_BIND="true")
        cmd = [drv, "bench", "--mode", "pcg", "--dim", "3", "--n", str(n_cpu), "--p", str(w["p"]),
               "--kind", w["kind"], "--quad", "gll" if w["gll"] else "gl", "--curved", str(int(w["curved"])),
               "--vdim", str(w.get("vdim", 1)),
               "--iters", str(iters),
               "--reps", str(reps), "--skip", str(skip), "--jacobi", "1"]
        out = subprocess.run(cmd, check=True, stdout=subprocess.PIPE, env=env).stdout.decode()
        r = json.loads(out.strip().splitlines()[-1])
        return dict(value=r["dofs_iters_per_s"] / 1e9, seconds=r["seconds"], ndofs=r["ndofs"],
                    kind="reference", cores=cores,
                    sample=f"BP{w['bp']} p={w['p']} {n_cpu}^3 hexes ({r['ndofs']} dofs), {iters} "
                           f"Jacobi-PCG iterations, MFEM -pa, MFEM_DEVICE=omp, {cores} threads")
    # port: oracle/oracle_pa.c, scalar, one core
    import numpy as np
    from oracle import oracle as orc
    import mfem_b200 as m
    fes = m.FiniteElementSpace(m.Mesh.MakeCartesian3D(n_cpu, n_cpu, n_cpu), m.H1_FECollection(w["p"], 3))
    ir = m.IntegrationRules(w["gll"]).Get(3, 2 * w["p"] - 1 if w["gll"] else 2 * w["p"] + 3)
    B, G = m.h1_dof_to_quad(w["p"], ir.pts1d)
    gm, off, ind = orc.restriction_build(fes.ne, fes.dof, fes.ndofs, fes.elem_dof, fes.dof_map)
    Bn, Gn = m.h1_dof_to_quad(1, ir.pts1d)
    J, detJ = orc.geometric_factors(3, 2, ir.nq1d, fes.ne, Bn, Gn, fes.elem_vertices)
    D = (orc.mass_setup(ir.nq1d ** 3, fes.ne, ir.weights, detJ, [1.0]) if w["kind"] == "mass" else
         orc.diffusion_setup(3, ir.nq1d, 1, fes.ne, ir.weights, J, [1.0]))
    op = orc.Operator(3, w["d1d"], ir.nq1d, fes.ne, fes.ndofs, gm, off, ind, B, G, D,
                      is_mass=w["kind"] == "mass", ess=fes.bdr_dofs)
    b = np.ones(fes.ndofs)
    b[fes.bdr_dofs] = 0
    dinv = op.jacobi(op.diag())
    t0 = time.time()
    for _ in range(reps):
        op.pcg(b, dinv, rel_tol=0.0, max_iter=iters)
    sec = (time.time() - t0) / reps
    return dict(value=fes.ndofs * iters / sec / 1e9, seconds=sec, ndofs=fes.ndofs, kind="port", cores=1,
                sample=f"BP{w['bp']} p={w['p']} {n_cpu}^3 hexes ({fes.ndofs} dofs), {iters} "
                       f"Jacobi-PCG iterations, C oracle port, 1 thread")


def main_reference(a, w):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_cpu = a.cpu_n or cpu_sample_n(w["p"])
    # one process: `warmup` untimed solves (capped at 1: a solve is seconds long, one is enough to
    # page everything in), then `steps` timed ones, averaged
    wu = min(a.warmup, 1)
    res = run_reference(w, n_cpu, a.iters, wu + a.steps, skip=wu)
    sec = res["seconds"]
    val = res["ndofs"] * a.iters / sec / 1e9
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "GDOF/s", "n_gpus": a.gpus,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"BP{w['bp']} {w['kind']} PA + PCG/Jacobi, p={w['p']}, Q1D={w['q1d']}, "
                                   f"{a.iters} fixed iterations; CPU sample of the GPU workload",
                       "sample": res["sample"]},
            "cpu_baseline": {"value": val, "unit": "GDOF/s", "cores": res["cores"], "kind": res["kind"],
                             "sample": res["sample"]},
            "e2e": {"value": val, "unit": "GDOF/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def main():
    a = parse()
    w = workload(a)
    if a.impl == "reference":
        return main_reference(a, w)

    import numpy as np
    import torch
    import mfem_b200 as m

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != a.gpus:
        if a.gpus == 1 and world == 1:
            pass
        else:
            raise SystemExit(f"--gpus {a.gpus} needs torchrun with {a.gpus} ranks (WORLD_SIZE={world})")
    # host table generation fans out over std::threads: share the cores between the ranks of the node
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 8)
    os.environ.setdefault("B2F_HOST_THREADS", str(max(2, min(16, cores // max(world, 1)))))
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = m.Device(f"cuda:{local_rank}")

    def barrier():
        if dist:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- parity first: the reference-checksum sample on the same N ranks ---------------------
    parity = None if a.no_parity else parity_check(m, np, torch, dev, dist, rank, world, w, a.iters)

    # ---- problem ------------------------------------------------------------------------
    px, py, pz = w["grid"]
    n, p = w["n"], w["p"]
    t_setup = time.time()
    if world > 1:
        mesh = m.Mesh(3, n * px, n * py, n * pz, True, part=(px, py, pz), rank=rank)
    else:
        mesh = m.Mesh.MakeCartesian3D(n, n, n)
    if w["curved"]:
        mesh.SetCurvature(3)
        mesh.Transform(m.curve_map)
    fes = m.FiniteElementSpace(mesh, m.H1_FECollection(p, 3), vdim=w["vdim"], ordering="byVDIM")
    rules = m.IntegrationRules(gll=w["gll"])
    ir = rules.Get(3, 2 * p - 1 if w["gll"] else 2 * p + 3)
    it = (m.MassIntegrator if w["kind"] == "mass" else m.DiffusionIntegrator)(1.0, ir)
    form = m.BilinearForm(dev, fes)
    form.SetAssemblyLevel("PARTIAL")
    form.AddDomainIntegrator(it)
    form.Assemble()
    halo = None
    if world > 1:
        ids = [m.Halo.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        halo = m.Halo(dev, rank, world, fes.ndofs, fes.halo_tables, ids[0], vdim=w["vdim"])
        if os.environ.get("B2F_HALO", "p2p") != "nccl":
            halo.connect_p2p(dist)  # shared dofs: peer-memory stores over NVLink, no NCCL send/recv
        form.SetHalo(halo)
    ess = fes.GetEssentialTrueDofs()
    b_l = m.LinearForm(dev, fes).Assemble()
    if halo:
        b = dev.empty(fes.GetTrueVSize())
        halo.MultTranspose(b_l, b)
    else:
        b = b_l
    nt = fes.GetTrueVSize()
    A, _, B = form.FormLinearSystem(ess, dev.zeros(nt), b)
    M = m.OperatorJacobiSmoother(form, ess)
    cg = m.CGSolver()
    cg.SetRelTol(0.0); cg.SetAbsTol(0.0); cg.SetMaxIter(a.iters)
    cg.SetOperator(A); cg.SetPreconditioner(M)
    cg.check_every = a.iters  # fixed iteration count: the host is consulted once per solve
    X = dev.zeros(nt)
    dev.sync()
    t_setup = time.time() - t_setup
    nglob = nt
    if dist:
        tt = torch.tensor([nt], dtype=torch.int64, device=dev.tdev)
        dist.all_reduce(tt)
        nglob = int(tt.item())
    del fes.elem_dof  # host tables are no longer needed
    # host buffers of the end-to-end path
    B_h = torch.empty(nt, dtype=torch.float64).pin_memory()
    X_h = torch.empty(nt, dtype=torch.float64).pin_memory()
    B_h.copy_(B)

    def solve_resident():
        cg.Mult(B, X)

    # end to end: every step copies its right-hand side from pinned host memory and its solution back.
    # The steps are independent solves, so the copies are double-buffered on two copy streams: the
    # H2D of step k+1 and the D2H of step k-1 overlap the solve of step k (all three inside the timed
    # region; the event chain below orders buffer reuse).
    B_in = [dev.empty(nt), dev.empty(nt)]
    X_out = [dev.empty(nt), dev.empty(nt)]
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
    main = torch.cuda.current_stream()
    ev_in = [torch.cuda.Event(), torch.cuda.Event()]
    ev_solved = [torch.cuda.Event(), torch.cuda.Event()]
    ev_out = [torch.cuda.Event(), torch.cuda.Event()]
    e2e_state = {"k": 0, "primed": False}

    def h2d(slot):
        with torch.cuda.stream(s_in):
            s_in.wait_event(ev_solved[slot])  # the solve that last read this buffer is done
            B_in[slot].copy_(B_h, non_blocking=True)
            ev_in[slot].record(s_in)

    def solve_e2e():
        k = e2e_state["k"]
        slot = k % 2
        if not e2e_state["primed"]:
            h2d(slot)
            e2e_state["primed"] = True
        h2d(1 - slot)                    # next step's input, in flight during this solve
        main.wait_event(ev_in[slot])
        main.wait_event(ev_out[slot])    # the D2H that last read this output buffer is done
        cg.Mult(B_in[slot], X_out[slot])
        ev_solved[slot].record(main)
        with torch.cuda.stream(s_out):
            s_out.wait_event(ev_solved[slot])
            X_h.copy_(X_out[slot], non_blocking=True)
            ev_out[slot].record(s_out)
        e2e_state["k"] = k + 1

    def drain_e2e():
        main.wait_event(ev_out[0])
        main.wait_event(ev_out[1])

    def timed(fn, steps, warmup, kernel_timing=False):
        for _ in range(warmup):
            fn()
        barrier()
        if kernel_timing:
            capi_call("b2f_ctx_kernel_timing", dev.ctx, 1)
        dev.reset_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        if fn is solve_e2e:
            drain_e2e()  # the last solutions have reached the host buffer
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        launches = dev.launch_count()
        kt = None
        if kernel_timing:
            import ctypes as C
            tot, cnt = C.c_double(), C.c_int()
            capi_call("b2f_ctx_kernel_time", dev.ctx, C.byref(tot), C.byref(cnt))
            capi_call("b2f_ctx_kernel_timing", dev.ctx, 0)
            kt = (tot.value, cnt.value)
        barrier()
        if dist:
            tm = torch.tensor([ms], dtype=torch.float64, device=dev.tdev)
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            ms = float(tm.item())
        if dist and kt is not None:
            # the operator kernel time of the SLOWEST rank (rank 0 owns every shared dof and has no
            # boundary-element launch: it is the lightest rank, not the one to report)
            tk = torch.tensor([kt[0]], dtype=torch.float64, device=dev.tdev)
            dist.all_reduce(tk, op=dist.ReduceOp.MAX)
            kt = (float(tk.item()), kt[1], kt[0])
        return ms, launches, kt

    from mfem_b200.capi import call as capi_call

    sampler = ClockSampler(local_rank) if rank == 0 else None
    ms, launches, kt = timed(solve_resident, a.steps, a.warmup, kernel_timing=True)
    clocks = sampler.stop() if sampler else None
    assert cg.GetNumIterations() == a.iters, (cg.GetNumIterations(), a.iters)
    final_norm = cg.GetFinalNorm()
    ms_e2e, _, _ = timed(solve_e2e, a.steps, min(a.warmup, 1))

    value = nglob * a.iters * a.steps / (ms * 1e-3) / 1e9
    e2e = nglob * a.iters * a.steps / (ms_e2e * 1e-3) / 1e9

    line = None
    if rank == 0:
        peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peaks = json.load(f)
        except OSError:
            pass
        peak = peaks.get("hbm_gbs", 6650.0)
        ab = alg_bytes_per_apply(w, fes.ne, fes.GetVSize())
        # one operator application per CG iteration + the initial one; with a halo the interior /
        # boundary split makes it several launches, whose times add up to one application
        n_applies = a.steps * (a.iters + 1)
        k_ms = kt[0] / max(n_applies, 1)
        achieved = ab / (k_ms * 1e-3) / 1e9
        # DRAM traffic of the operator kernel from the committed ncu capture of this workload
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "operator_traffic.json")) as f:
                tj = json.load(f)
            key = f"bp{w['bp']}_p{p}_n{n}"
            if key in tj:
                traffic = tj[key]["dram_bytes_per_launch"]
        except (OSError, ValueError, KeyError):
            pass
        line = {
            "metric": METRIC, "value": value, "unit": "GDOF/s", "n_gpus": a.gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": ms / a.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"BP{w['bp']} {w['kind']} PA + PCG/Jacobi, p={p}, Q1D={w['q1d']}, "
                                   f"{n}^3 hexes per GPU{' (curved, Q3 nodes)' if w['curved'] else ''} on a {px}x{py}x{pz} grid, {nglob} dofs, "
                                   f"{a.iters} fixed CG iterations per step",
                       "dofs_global": nglob, "dofs_per_gpu": nglob // a.gpus, "elements_per_gpu": fes.ne,
                       "cg_iterations_per_step": a.iters,
                       "halo": ("none" if not halo else "peer-memory stores (CUDA IPC over NVLink) + ncclAllReduce dots"
                                if os.environ.get("B2F_HALO", "p2p") != "nccl" else "ncclSend/Recv + ncclAllReduce dots"),
                       "l2": "inputs (quadrature data, "
                       f"{it_bytes(w, fes.ne) / 1e9:.1f} GB per GPU) far exceed the 126 MB L2",
                       "final_norm": final_norm, "setup_seconds": round(t_setup, 1)},
            "e2e": {"value": e2e, "unit": "GDOF/s", "h2d_bytes_per_step": nt * 8 * a.gpus,
                    "d2h_bytes_per_step": nt * 8 * a.gpus, "ms_per_step": ms_e2e / a.steps,
                    "copies": "pinned host buffers; double-buffered on two copy streams: the H2D of step k+1 and the "
                              "D2H of step k-1 overlap the solve of step k, all inside the timed region"},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic,
                         "kernel": "pa_apply_3d_eo_kernel (fused gather + even-odd G^T D G + scatter)",
                         "kernel_ms": k_ms, "kernel_ms_is": "max over ranks" if a.gpus > 1 else "this GPU",
                         "kernel_ms_rank0": (kt[2] / max(n_applies, 1)) if len(kt) > 2 else k_ms,
                         "kernel_launches": kt[1],
                         "launches_per_application": kt[1] / max(n_applies, 1),
                         "kernel_share_of_step": kt[0] / ms,
                         "algorithmic_bytes_per_launch": ab,
                         "operator_gdofs": fes.GetVSize() / (k_ms * 1e-3) / 1e9,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if "hbm_gbs" in peaks
                         else "fallback 6650 GB/s"},
            "clocks": clocks,
            "parity_check": parity,
        }
        if a.gpus == 1 and not a.no_cpu_baseline:
            try:
                n_cpu = a.cpu_n or cpu_sample_n(p)
                r = run_reference(w, n_cpu, a.iters, 2, skip=1)
                line["cpu_baseline"] = {"value": r["value"], "unit": "GDOF/s", "cores": r["cores"],
                                        "kind": r["kind"], "sample": r["sample"]}
            except Exception as ex:  # the baseline is reported, never required
                line["cpu_baseline"] = {"value": None, "unit": "GDOF/s", "cores": 0, "kind": "reference",
                                        "sample": f"failed: {ex}"}
    if a.gpus == 1:
        # free the headline problem before the baselines and the table
        form.close(); it.close()
        del A, B, X, M, cg, b, b_l, B_in, X_out, B_h, X_h, form, it
        torch.cuda.empty_cache()
        if not a.no_gpu_baseline:
            line["gpu_baseline"] = run_gpu_baseline(w, a.iters)
            gb = line["gpu_baseline"]
            if "value" in gb:
                gb["b200_over_reference"] = {"pcg": line["value"] / gb["value"],
                                             "operator": line["roofline"]["operator_gdofs"] / gb["operator_gdofs"],
                                             "operator_vs_element_kernel_alone":
                                             line["roofline"]["operator_gdofs"] / gb["kernel_gdofs"]}
        if not a.no_anchor and w["bp"] == 3 and p == 4 and not a.n:
            # the per-GPU problem of the --gpus N > 1 runs (116^3 hexes), on one GPU: the equal-size
            # anchor for weak-scaling efficiency (the headline N=1 config is the larger 133^3 one)
            line["weak_anchor"] = anchor_run(m, torch, dev, w, 116, a.iters, a.steps, a.warmup)
        if not a.no_sweep:
            t_sw = time.time()
            line["sweep"] = {"what": "BP1/BP3/BP5 x p=1..8, Jacobi-PCG GDOF/s and the fused operator's share of the "
                                     f"HBM roofline, ~{a.sweep_dofs:.0e} dofs per case, {a.iters} fixed iterations, "
                                     "2 timed solves after 2 warm-up solves",
                             "rows": sweep_table(m, torch, dev, capi_call, peak, a.sweep_dofs, a.iters)}
            line["sweep"]["seconds"] = round(time.time() - t_sw, 1)
        if not a.no_lor:
            # last, and never required: nothing here can touch the numbers above
            try:
                line["lor"] = lor_record(m, torch, dev, capi_call, peak, with_ref=not a.no_gpu_baseline)
            except Exception as ex:
                line["lor"] = {"unavailable": f"{type(ex).__name__}: {ex}"}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if parity is not None and parity.get("ok") is False:
        if rank == 0:
            print(f"PARITY CHECK FAILED: {parity}", file=sys.stderr, flush=True)
        sys.exit(3)
    if dist:
        dist.barrier()
        if halo:
            halo.close()
        dist.destroy_process_group()


def ref_cuda_driver_path():
    p = os.path.join(ROOT, "oracle", "_ref_cuda", "ref_driver_cuda")
    return p if os.path.isfile(p) and os.access(p, os.X_OK) else None


REF_KERNEL = {"diffusion": "SmemPADiffusionApply3D (fem/integ/bilininteg_diffusion_kernels.hpp:989-1215)",
              "mass": "SmemPAMassApply3D (fem/integ/bilininteg_mass_kernels.hpp:815-1050)"}


def run_gpu_baseline(w, iters, n=None):
    """The UNMODIFIED reference's own CUDA backend (MFEM_USE_CUDA, generic MFEM_FORALL kernels,
    oracle/Makefile target ref_cuda) on this GPU, same BP, at the largest size its 32-bit offsets and
    its resident Jacobians allow: AddMultPA on E-vectors (the element kernel alone), BilinearForm::Mult
    (restriction + kernel + restriction^T) and the Jacobi-PCG solve.  A baseline, like cpu_baseline."""
    drv = ref_cuda_driver_path()
    if not drv:
        return {"unavailable": "oracle/_ref_cuda/ref_driver_cuda not built (make -C oracle ref_cuda)"}
    p = w["p"]
    if n is None:
        # 9 Q^3 doubles of J per element in one 32-bit sized Vector (mesh/mesh.cpp:15467), <= ~64 M dofs
        n_j = int(((2 ** 31 - 1) // (9 * w["q1d"] ** 3)) ** (1 / 3))
        n = int(min(w["n"], n_j, round(64e6 ** (1 / 3) / p)))
    cmd = [drv, "bench", "--mode", "all", "--dim", "3", "--n", str(n), "--p", str(p), "--kind", w["kind"],
           "--quad", "gll" if w["gll"] else "gl", "--curved", str(int(w["curved"])), "--iters", str(iters),
           "--reps", "3", "--jacobi", "1"]
    try:
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=900,
                           env=dict(os.environ, REF_DRIVER_DEVICE="cuda"))
        if r.returncode != 0:
            err = r.stderr.decode(errors="replace").strip().splitlines()
            why = next((ln.strip() for ln in err if "-->" in ln), err[-1] if err else "no output")
            return {"unavailable": f"the reference's CUDA backend failed at p={p}, n={n}: {why}"}
        j = json.loads(r.stdout.decode().strip().splitlines()[-1])
    except Exception as ex:  # a baseline is reported, never required
        return {"unavailable": f"{type(ex).__name__}: {ex}"}
    nd = j["ndofs"]
    return {"value": j["dofs_iters_per_s"] / 1e9, "unit": "GDOF/s",
            "operator_gdofs": nd / j["mult_seconds"] / 1e9, "kernel_gdofs": nd / j["kernel_seconds"] / 1e9,
            "operator_ms": j["mult_seconds"] * 1e3, "kernel_ms": j["kernel_seconds"] * 1e3,
            "kernel": REF_KERNEL[w["kind"]] + ("; (9,10) is not in its dispatch table: thread-per-element fallback"
                                              if p == 8 else ""),
            "config": f"BP{w['bp']} p={p} {n}^3 hexes, {nd} dofs, MFEM -pa -d cuda (sm_100), {iters} Jacobi-PCG "
                      f"iterations", "dofs": nd, "final_norm": j["final_norm"]}


def lor_record(m, torch, dev, capi_call, peak, with_ref, p=4, n=64, reps=3):
    """SURVEY 8f row 4, the assembled half: the low-order-refined matrix of the BP3 form (what
    LORSolver<AMG> is built from; BatchedLORAssembly::Assemble, fem/lor/lor_batched.cpp:712) at
    p = 4, 64^3 hexes (17 M dofs, 455 M nonzeros): seconds of one Assemble (stencil array, CSR,
    boundary elimination; CUDA events, allocations included), the algorithmic bytes (24 B per E-dof of
    vertex coordinates + 12 B per nonzero + 4 B per row) against the HBM peak, the checksum |A x| for
    x = Vector::Randomize(1), and beside it the reference's own CUDA backend on the same problem."""
    import ctypes as C
    import numpy as np
    fes = m.FiniteElementSpace(m.Mesh.MakeCartesian3D(n, n, n), m.H1_FECollection(p, 3))
    N = fes.GetNDofs()
    ess = fes.GetEssentialTrueDofs()
    fes.GetElementRestriction(dev)
    lor = m.BatchedLORAssembly(dev, fes)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    times, ktimes, A = [], [], None
    for r in range(reps + 1):
        if A is not None:
            A.close()
        torch.cuda.synchronize()
        e0.record()
        A = lor.Assemble(None, ess, mass_coeff=0.0, diff_coeff=1.0)
        e1.record()
        torch.cuda.synchronize()
        if r:
            times.append(e0.elapsed_time(e1) * 1e-3)
            ktimes.append(A.kernel_ms() * 1e-3)
    xs = np.empty(N)
    capi_call("b2f_vector_randomize_h", xs.ctypes.data_as(C.c_void_p), xs.size, 1)
    y = dev.empty(N)
    A.Mult(dev.to_device(xs), y)
    dev.sync()
    best, kbest, nnz = min(times), min(ktimes), A.NumNonZeroElems()
    alg = 24.0 * fes.ne * (p + 1) ** 3 + 12.0 * nnz + 4.0 * N
    rec = {"what": "BatchedLORAssembly::Assemble, H1 diffusion, CSR on the device", "p": p, "dofs": N, "nnz": nnz,
           "assemble_ms": best * 1e3, "assemble_ms_is": "one call, host clock on the device timeline: kernels + the "
           "call's own cudaMalloc / cudaFree of the stencil array and the CSR arrays",
           "kernel_ms": kbest * 1e3, "mdofs_per_s": N / best / 1e6, "algorithmic_gb": alg / 1e9,
           "frac_of_hbm_roofline": alg / kbest / 1e9 / peak, "checksum_norm_Ax": float(torch.linalg.norm(y))}
    A.close()
    del lor, y
    torch.cuda.empty_cache()
    drv = ref_cuda_driver_path()
    if with_ref and drv:
        try:
            r = subprocess.run([drv, "bench", "--mode", "lor", "--p", str(p), "--n", str(n), "--reps", str(reps),
                                "--lorkind", "diffusion"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=600,
                               env=dict(os.environ, REF_DRIVER_DEVICE="cuda"))
            j = json.loads(r.stdout.decode().strip().splitlines()[-1])
            rec["reference_cuda"] = {"assemble_ms": j["assemble_seconds"] * 1e3,
                                     "mdofs_per_s": j["ndofs"] / j["assemble_seconds"] / 1e6,
                                     "checksum_norm_Ax": j["ynorm"],
                                     "config": "the unmodified reference, MFEM_USE_CUDA (sm_100), same mesh, form and x; "
                                               "best of the calls after a warm-up one (its CSR arrays are reused)"}
            rec["b200_over_reference"] = j["assemble_seconds"] / best
            rec["b200_kernels_over_reference"] = j["assemble_seconds"] / kbest
            rec["checksum_rel_diff"] = abs(j["ynorm"] - rec["checksum_norm_Ax"]) / abs(j["ynorm"])
        except Exception as ex:
            rec["reference_cuda"] = {"unavailable": f"{type(ex).__name__}: {ex}"}
    return rec


def sweep_table(m, torch, dev, capi_call, peak, dofs, iters):
    """The BASELINE metric as a table (tests/benchmarks/bench_ceed.cpp:95-176, 301-357): BP1 / BP3 / BP5,
    p = 1..8, ~`dofs` dofs each: Jacobi-PCG GDOF/s and the fused operator's share of the HBM roofline."""
    import ctypes as C
    rows = []
    for bp in (1, 3, 5):
        for p in range(1, 9):
            gll = bp == 5
            n = max(2, round(dofs ** (1 / 3) / p))
            fes = m.FiniteElementSpace(m.Mesh.MakeCartesian3D(n, n, n), m.H1_FECollection(p, 3))
            ir = m.IntegrationRules(gll).Get(3, 2 * p - 1 if gll else 2 * p + 3)
            it = (m.MassIntegrator if bp == 1 else m.DiffusionIntegrator)(1.0, ir)
            form = m.BilinearForm(dev, fes)
            form.AddDomainIntegrator(it)
            form.Assemble()
            nd = fes.ndofs
            ess = fes.GetEssentialTrueDofs()
            b = m.LinearForm(dev, fes).Assemble()
            A, _, B = form.FormLinearSystem(ess, dev.zeros(nd), b)
            M = m.OperatorJacobiSmoother(form, ess)
            cg = m.CGSolver()
            cg.SetRelTol(0.0); cg.SetAbsTol(0.0); cg.SetMaxIter(iters)
            cg.SetOperator(A); cg.SetPreconditioner(M)
            cg.check_every = iters
            X = dev.zeros(nd)
            del fes.elem_dof
            for _ in range(2):
                cg.Mult(B, X)
            torch.cuda.synchronize()
            capi_call("b2f_ctx_kernel_timing", dev.ctx, 1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            steps = 2
            e0.record()
            for _ in range(steps):
                cg.Mult(B, X)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            tot, cnt = C.c_double(), C.c_int()
            capi_call("b2f_ctx_kernel_time", dev.ctx, C.byref(tot), C.byref(cnt))
            capi_call("b2f_ctx_kernel_timing", dev.ctx, 0)
            k_ms = tot.value / max(cnt.value, 1)
            k = 1 if bp == 1 else 6
            ab = (k * ir.nq1d ** 3 * 8 + 4 * (p + 1) ** 3) * fes.ne + 16 * nd
            rows.append({"bp": bp, "p": p, "n": n, "dofs": nd,
                         "pcg_gdofs": round(nd * iters * steps / (ms * 1e-3) / 1e9, 3),
                         "operator_gdofs": round(nd / (k_ms * 1e-3) / 1e9, 3), "operator_ms": round(k_ms, 4),
                         "operator_frac_of_hbm_roofline": round(ab / (k_ms * 1e-3) / 1e9 / peak, 4)})
            form.close(); it.close()
            del A, B, X, M, cg, b, form, fes
            torch.cuda.empty_cache()
    return rows


def anchor_run(m, torch, dev, w, n, iters, steps, warmup):
    """One more device-resident timing of the same BP at `n`^3 hexes on this GPU."""
    p = w["p"]
    fes = m.FiniteElementSpace(m.Mesh.MakeCartesian3D(n, n, n), m.H1_FECollection(p, 3))
    ir = m.IntegrationRules(gll=w["gll"]).Get(3, 2 * p - 1 if w["gll"] else 2 * p + 3)
    it = (m.MassIntegrator if w["kind"] == "mass" else m.DiffusionIntegrator)(1.0, ir)
    form = m.BilinearForm(dev, fes)
    form.AddDomainIntegrator(it)
    form.Assemble()
    nd = fes.ndofs
    ess = fes.GetEssentialTrueDofs()
    b = m.LinearForm(dev, fes).Assemble()
    A, _, B = form.FormLinearSystem(ess, dev.zeros(nd), b)
    M = m.OperatorJacobiSmoother(form, ess)
    cg = m.CGSolver()
    cg.SetRelTol(0.0); cg.SetAbsTol(0.0); cg.SetMaxIter(iters)
    cg.SetOperator(A); cg.SetPreconditioner(M)
    cg.check_every = iters
    X = dev.zeros(nd)
    del fes.elem_dof
    for _ in range(warmup):
        cg.Mult(B, X)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        cg.Mult(B, X)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    out = {"workload": f"BP{w['bp']} p={p} {n}^3 hexes, {nd} dofs on 1 GPU (the per-GPU block of --gpus N > 1)",
           "value": nd * iters * steps / (ms * 1e-3) / 1e9, "unit": "GDOF/s", "ms_per_step": ms / steps, "dofs": nd}
    form.close(); it.close()
    del A, B, X, M, cg, b, form, fes
    torch.cuda.empty_cache()
    return out


def serial_dof_of_lattice(m, np, n, p):
    """Reference (serial MakeCartesian3D + H1 numbering) dof id of every point of the global
    (p n + 1)^3 lattice, from the serial space's own element-dof table."""
    fes = m.FiniteElementSpace(m.Mesh.MakeCartesian3D(n, n, n), m.H1_FECollection(p, 3))
    D, L = p + 1, p * n + 1
    ev = fes.elem_vertices.reshape(fes.ne, 3, 8)
    ijk = np.rint(ev.min(axis=2) * n).astype(np.int64)  # element coordinates
    a = np.arange(D, dtype=np.int64)
    gx = ijk[:, 0, None, None, None] * p + a[None, None, None, :]
    gy = ijk[:, 1, None, None, None] * p + a[None, None, :, None]
    gz = ijk[:, 2, None, None, None] * p + a[None, :, None, None]
    lat = (gx + L * (gy + L * gz)).reshape(fes.ne, D ** 3)          # lexicographic, x fastest
    dofs = fes.elem_dof.reshape(fes.ne, D ** 3)[:, fes.dof_map]    # restriction.cpp:60-75
    s_of_lat = np.empty(L ** 3, np.int64)
    s_of_lat[lat.ravel()] = dofs.ravel()
    return s_of_lat, fes.ndofs


def parity_check(m, np, torch, dev, dist, rank, world, w, iters):
    """Driver-visible parity of the GPU arm, at every N, before anything is timed: a sample of the
    bench's own BP (partitioned over the N ranks, same halo mode) against the UNMODIFIED
    reference's checksums (tests/golden/checksums.json, written by oracle/make_checksums.py):
    ||A x||_2 for x = Vector::Randomize(1) in the reference's numbering, and ||X||_2 and
    GetFinalNorm() of `iters` fixed Jacobi-PCG iterations (bench_ceed.cpp:222-243 protocol)."""
    import ctypes as C
    from mfem_b200.capi import call
    with open(os.path.join(ROOT, "tests", "golden", "checksums.json")) as f:
        cases = json.load(f)["cases"]
    sbp = w["bp"] - 1 if w["vdim"] == 3 else w["bp"]
    px, py, pz = w["grid"]
    cand = [c for c in cases.values() if c["bp"] == sbp and c["p"] == w["p"] and bool(c["curved"]) == w["curved"]
            and c["n"] % 2 == 0 and c["pcg_iters"] == iters]
    if not cand or w["vdim"] != 1:
        return {"skipped": f"no reference checksum for BP{w['bp']} p={w['p']} curved={w['curved']}"}
    ref = max(cand, key=lambda c: c["n"])
    n, p = ref["n"], w["p"]
    if world > 1:
        mesh = m.Mesh(3, n, n, n, True, part=(px, py, pz), rank=rank)
    else:
        mesh = m.Mesh.MakeCartesian3D(n, n, n)
    if w["curved"]:
        mesh.SetCurvature(3)
        mesh.Transform(m.curve_map)
    fes = m.FiniteElementSpace(mesh, m.H1_FECollection(p, 3))
    ir = m.IntegrationRules(gll=w["gll"]).Get(3, 2 * p - 1 if w["gll"] else 2 * p + 3)
    it = (m.MassIntegrator if w["kind"] == "mass" else m.DiffusionIntegrator)(1.0, ir)
    form = m.BilinearForm(dev, fes)
    form.SetAssemblyLevel("PARTIAL")
    form.AddDomainIntegrator(it)
    form.Assemble()
    halo = None
    xs = np.empty(ref["ndofs"])
    call("b2f_vector_randomize_h", xs.ctypes.data_as(C.c_void_p), xs.size, 1)
    if world > 1:
        ids = [m.Halo.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        halo = m.Halo(dev, rank, world, fes.ndofs, fes.halo_tables, ids[0])
        if os.environ.get("B2F_HALO", "p2p") != "nccl":
            halo.connect_p2p(dist)
        form.SetHalo(halo)
        s_of_lat, nser = serial_dof_of_lattice(m, np, n, p)
        assert nser == ref["ndofs"]
        t2s = s_of_lat[fes.halo_tables["ldof_lattice"][fes.halo_tables["ltdof_ldof"]]]
        x_t = dev.to_device(np.ascontiguousarray(xs[t2s]))
    else:
        assert fes.ndofs == ref["ndofs"], (fes.ndofs, ref["ndofs"])
        x_t = dev.to_device(xs)
    nt = fes.GetTrueVSize()

    def gnorm(v):
        s = (v.double() ** 2).sum().reshape(1)
        if dist:
            dist.all_reduce(s)
        return float(s.sqrt().item())

    y_t = dev.empty(nt)
    form.Mult(x_t, y_t)
    dev.sync()
    op_norm = gnorm(y_t)
    ess = fes.GetEssentialTrueDofs()
    b_l = m.LinearForm(dev, fes).Assemble()
    if halo:
        b = dev.empty(nt)
        halo.MultTranspose(b_l, b)
    else:
        b = b_l
    A, _, B = form.FormLinearSystem(ess, dev.zeros(nt), b)
    M = m.OperatorJacobiSmoother(form, ess)
    cg = m.CGSolver()
    cg.SetRelTol(0.0); cg.SetAbsTol(0.0); cg.SetMaxIter(iters)
    cg.SetOperator(A); cg.SetPreconditioner(M)
    X = dev.zeros(nt)
    cg.Mult(B, X)
    dev.sync()
    x_norm, fin = gnorm(X), cg.GetFinalNorm()
    res = {"case": f"BP{sbp} p={p} {n}^3 hexes{' curved' if w['curved'] else ''}, {ref['ndofs']} dofs, over {world} rank(s)",
           "reference": "tests/golden/checksums.json (unmodified MFEM -pa, oracle/make_checksums.py)",
           "n_ranks": world,
           "rel_err_y": abs(op_norm - ref["op_norm"]) / ref["op_norm"],
           "x_norm_rel": abs(x_norm - ref["pcg_x_norm"]) / ref["pcg_x_norm"],
           # relative to the INITIAL norm: a solve that has converged to round-off (BP1) ends on noise
           "final_norm_rel": abs(fin - ref["pcg_final_norm"]) / ref["pcg_initial_norm"],
           "tol": {"rel_err_y": 1e-12, "x_norm_rel": 1e-9, "final_norm_rel": 1e-9}}
    res["ok"] = bool(res["rel_err_y"] < 1e-12 and res["x_norm_rel"] < 1e-9 and res["final_norm_rel"] < 1e-9
                     and cg.GetNumIterations() == ref["pcg_iters"])
    if halo:
        halo.close()
    form.close(); it.close()
    return res


def it_bytes(w, ne):
    return (1 if w["kind"] == "mass" else 6) * w["q1d"] ** 3 * 8 * ne


if __name__ == "__main__":
    main()
