#!/usr/bin/env python
"""bench.py -- cells*steps/sec of one Gen-FVGN training step (NNmodel.forward + loss.backward + Adam) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--cells C] [--net EPD|TransFVGN_v1|TransFVGN_v2]
                    [--mp G] [--precision fp32|bf16]

Workload (BASELINE.json config 5): one synthetic jittered quad mesh of `--cells` cells (default 4M) per GPU, cavity BCs,
Navier-Stokes theta, net = Encoder -> GnBlock x G -> Decoder + finite-volume PDE loss, resident batch
(the solve_with_grad_GPU.py regime).  N > 1: data parallel, one mesh per rank (weak scaling), one NCCL all-reduce of the
flat gradient per step.  Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cells", type=int, default=int(os.environ.get("FVGN_BENCH_CELLS", 4_000_000)))
    ap.add_argument("--net", default=os.environ.get("FVGN_BENCH_NET", "EPD"))
    ap.add_argument("--mp", type=int, default=int(os.environ.get("FVGN_BENCH_MP", 6)))
    ap.add_argument("--precision", default=os.environ.get("FVGN_PRECISION", "f16"), choices=["bf16", "f16", "fp32"],
                    help="bf16 / f16: tcgen05 modes (f16 = IEEE-half operands, the 11-bit significand of the reference's TF32 "
                         "GPU arithmetic); fp32: SIMT FMA mode")
    ap.add_argument("--cpu-cells", type=int, default=0,
                    help="cell count of the bounded CPU sample (cpu_baseline / --impl reference); 0 = sized from a calibration "
                         "step so that the sample takes ~20 s (cpu_baseline) / the whole run ~3 min (--impl reference)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true",
                    help="skip the sub-records (precision_modes, nets, example_meshes at N=1; cells at N>1)")
    ap.add_argument("--kernel-summary", default=None, metavar="PATH",
                    help="after the timed runs: 2 more steps under torch.profiler (CUPTI), per-kernel device time table -> PATH")
    ap.add_argument("--graph", action="store_true",
                    help="replay the whole step (fwd + bwd + Adam) as ONE CUDA graph (gen_fvgn_steady_b200.graphed); for the "
                         "launch-bound sizes of the reference's example meshes (10 k - 100 k cells), single GPU")
    ap.add_argument("--halo-layers", type=int, default=None,
                    help="--parallel cells: halo depth in cell layers; 3 = ghost refresh after every GnBlock, 6 = every 2nd "
                         "block, >= 3*mp+2 = no latent exchange at all (more redundant compute, no per-block synchronisation; the default)")
    ap.add_argument("--parallel", default="dp", choices=["dp", "cells"],
                    help="N>1: dp = one mesh of --cells cells per GPU, gradient all-reduce (weak scaling, the default the "
                         "driver measures); cells = ONE mesh of --cells cells partitioned over the GPUs with a per-GnBlock "
                         "halo exchange (strong scaling, SURVEY.md section 8(e).2)")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def alg_bytes_step(N, E, C, K, X, G):
    """ALG_BYTES(step)."""
    return G * (16.5 * N + 9 * E) * 512 + (5 * N + 2 * E) * 512 + 2.5 * (28 * X + 250 * N + 24 * K + 40 * E + 50 * C)


def make_mesh(cells, seed, device):
    """One synthetic quad mesh with ~cells cells -> the converter/loader dict (numpy) + initial field."""
    from gen_fvgn_steady_b200.mesh import synthetic as S
    n = max(int(round(cells ** 0.5)), 4)
    if device is not None and n * n > 300_000:
        from gen_fvgn_steady_b200.mesh import synthetic_torch as ST
        return ST.make_case(n, kind="quad", bc="cavity", seed=seed, device=device)
    return S.make_case(n, kind="quad", bc="cavity", seed=seed)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


# ----------------------------------------------------------------------------------------------- CPU reference arm
def cpu_reference_step_times(cells, net, mp, steps, warmup, threads):
    """The reference's algorithm on the host cores: oracle/fvgn_oracle.py (a CPU restatement pinned to the unmodified
    reference's golden vectors; the reference itself is Python that needs torch_scatter / PyG and /root/reference, none of
    which exists on the GPU box, so it cannot travel).  fwd + backward + Adam on one synthetic quad mesh of `cells` cells,
    initial weights from the same seeded initialiser as the GPU arm.  Returns ([sec per timed step], cells)."""
    from oracle import fvgn_oracle as O
    from gen_fvgn_steady_b200.FVMmodel.importer import NNmodel
    from gen_fvgn_steady_b200.utils.get_param import params as default_params
    torch.set_num_threads(threads)
    mesh, uvp = make_mesh(cells, 0, None)
    g = O.graphs_from_meshes([mesh], [uvp], torch.float32)
    torch.manual_seed(0)
    model = NNmodel(default_params(net=net, message_passing_num=mp))   # parameter container only: no kernel runs
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    leaves = {k: v.requires_grad_(True) for k, v in sd.items() if not k.startswith("node_norm.")}
    opt = torch.optim.Adam(list(leaves.values()), lr=5e-5)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        opt.zero_grad(set_to_none=True)
        res = O.nnmodel_forward(sd, g, net=net, mp_num=mp)
        loss = O.script_loss(res)
        loss.backward()
        opt.step()
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    return times, int(g["centroid"].shape[0])


def cpu_sample_cells(net, mp, threads, n_steps, budget_s, lo=20_000, hi=1_000_000):
    """Cell count of the bounded CPU sample: one calibration step at `lo` cells, then the largest mesh whose
    `n_steps` steps fit `budget_s` seconds (the CPU path is linear in the mesh size)."""
    t, c = cpu_reference_step_times(lo, net, mp, 1, 1, threads)
    rate = c / t[0]
    return int(min(max(rate * budget_s / max(n_steps, 1), lo), hi)), rate


def run_reference(args):
    """Reference arm: the reference's CPU path (the pinned port, see cpu_reference_step_times) on all host cores of the
    box, same net / G / loss / metric as the GPU arm.  Each step is a BOUNDED SAMPLE of the GPU arm's 4 M-cell workload:
    the largest mesh of the same generator for which warm-up + K steps end within ~3 minutes (stated in `config`)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    warm = max(min(args.warmup, 1), 0)
    steps = max(args.steps, 1)
    cells = args.cpu_cells
    if cells <= 0:
        cells, _ = cpu_sample_cells(args.net, args.mp, threads, steps + warm, 170.0)
    times, C = cpu_reference_step_times(cells, args.net, args.mp, steps, warm, threads)
    sec = float(np.median(times))
    val = C / sec
    sample = (f"{C}-cell synthetic quad mesh (same generator, net, G, loss as the GPU arm's {args.cells}-cell mesh; sized so "
              f"that {warm} warm-up + {steps} steps fit ~3 min), fwd+bwd+Adam, fp32, {threads} threads, median step")
    line = {"impl": "reference", "metric": "cells*steps/sec (fwd+bwd train step)", "value": val, "unit": "cells*steps/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(workload_config(args), precision="fp32 (CPU)", cpu_sample_cells=C),
            "cpu_baseline": {"value": val, "unit": "cells*steps/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "cells*steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def _gn_blocks(net, mp):
    """GnBlocks per forward: TransFVGN_v2 runs two processors of --mp blocks each (TransFVGN_v2.py:73-82)."""
    return mp * (2 if net in ("TransFVGN_v2", "TransFVGN") else 1)


def workload_config(args):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cells_mode = getattr(args, "parallel", "dp") == "cells" and world > 1
    size = f"ONE {args.cells}-cell mesh partitioned over {world} GPUs" if cells_mode else f"{args.cells} cells/GPU"
    return {"workload": f"synthetic jittered quad mesh, {size}, cavity BC, NS theta; "
                        f"{args.net} G={_gn_blocks(args.net, args.mp)} + FV PDE loss; resident batch (solve_with_grad regime)",
            "net": args.net, "gn_blocks": _gn_blocks(args.net, args.mp), "cuda_graph": bool(getattr(args, "graph", False)),
            "cells_per_gpu": args.cells // world if cells_mode else args.cells, "precision": args.precision,
            "l2": "inputs/activations (GBs) far exceed the 126 MB L2; no explicit flush"}


DTYPE_TEXT = {"fp32": "f32 (SIMT FMA)",
              "bf16": "bf16 tcgen05 operands and latent streams between the GnBlocks; f32 accumulators, LayerNorm, residual adds and "
                      "gradient streams",
              "f16": "f16 tcgen05 operands, latent streams and (power-of-two pre-scaled) gradient streams between the GnBlocks "
                     "(11-bit significand = the reference's TF32 GPU arithmetic); f32 accumulators, LayerNorm, residual adds, "
                     "reductions and parameter gradients"}


# ----------------------------------------------------------------------------------------------- our arm
class Job:
    """One resident-batch training job (model + optimizer + graphs) on this rank: step(), timed(), close()."""

    def __init__(self, dev, rank, world, graphs, net, mp, precision, cells_mode=False, halo=None, graph=False, warmup=3):
        from gen_fvgn_steady_b200.FVMmodel.importer import NNmodel
        from gen_fvgn_steady_b200.plan import GraphPlan
        from gen_fvgn_steady_b200.utils.get_param import params as default_params
        from gen_fvgn_steady_b200 import parallel
        self.dev, self.rank, self.world, self.cells_mode, self.halo = dev, rank, world, cells_mode, halo
        self.parallel = parallel
        self.graphs = graphs
        gn, gx, ge, gc, gi = graphs
        self.p = p = default_params(net=net, message_passing_num=mp, precision=precision)
        torch.manual_seed(0)
        self.model = model = NNmodel(p).to(dev)
        if cells_mode:
            model.enable_cell_partition(True)
        elif world > 1:
            model.enable_data_parallel(True)
        self.flat_grad = parallel.flatten_gradients(model)
        self.opt = torch.optim.Adam(model.parameters(), lr=p.lr, fused=True, capturable=bool(graph))
        self.plan = plan = GraphPlan.of(gn, gx, ge, gc, p.order)
        self.sizes = (plan.N, plan.E, plan.C, plan.K, int(gx.face_node_x.shape[1]))
        N = plan.N
        raw = getattr(gn, "_bench_x_raw", None)   # the model normalises graph_node.x in place: keep the loader's raw features
        if raw is None:
            raw = gn._bench_x_raw = gn.x.detach().clone()
        self.x_host = raw.cpu().pin_memory()
        self.x_dev0 = raw
        # e2e regime: the next step's x[N,12] is copied host -> device on a side stream into the other of two buffers while
        # the current step computes (what a loader with a prefetch queue does); results leave on the same side stream
        self.copy_stream = torch.cuda.Stream(device=dev)
        self.x_bufs, self.x_ready, self.k = [torch.empty_like(raw), torch.empty_like(raw)], [None, None], 0
        self._keep = None
        self.uvp_host = torch.empty((N, 3), dtype=torch.float32).pin_memory()
        self.loss_host = torch.empty((), dtype=torch.float32).pin_memory()
        self.gstep = None
        if graph:
            from gen_fvgn_steady_b200.graphed import GraphedTrainStep
            gn.x, gn.norm_uvp, gn.norm_global = self.x_dev0.clone(), True, True
            post = None
            if cells_mode:
                post = lambda: parallel.sum_gradients(self.flat_grad)
            elif world > 1:
                post = lambda: parallel.allreduce_gradients(self.flat_grad, world)
            self.gstep = GraphedTrainStep(model, self.opt, graphs, self.script_loss, warmup=max(warmup, 3),
                                          freeze_normalizer=True, post_backward=post)

    def script_loss(self, out):
        p = self.p
        return torch.mean(torch.log(p.loss_press * out[3] + p.loss_cont * out[0] + p.loss_mom * out[1] + p.loss_mom * out[2]))

    def _prefetch(self, slot):
        with torch.cuda.stream(self.copy_stream):
            self.x_bufs[slot].copy_(self.x_host, non_blocking=True)
            self.x_ready[slot] = torch.cuda.Event()
            self.x_ready[slot].record(self.copy_stream)

    def _next_x(self):
        """This step's input (prefetched during the previous step) + the H2D copy of the next step's input in flight."""
        cur = torch.cuda.current_stream()
        slot = self.k & 1
        self.k += 1
        if self.x_ready[slot] is None:
            self._prefetch(slot)
        cur.wait_event(self.x_ready[slot])
        self.x_ready[slot] = None
        self.copy_stream.wait_stream(cur)      # everything enqueued so far (the previous step) is done with the other buffer
        self._prefetch(slot ^ 1)
        return self.x_bufs[slot]

    def _results_to_host(self, uvp, loss):
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream())
        self._keep = (uvp, loss)               # alive until the next step's copies are enqueued
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(ev)
            self.uvp_host.copy_(uvp, non_blocking=True)
            self.loss_host.copy_(loss, non_blocking=True)
        uvp.record_stream(self.copy_stream)
        loss.record_stream(self.copy_stream)

    def eager_step(self, e2e):
        gn, gx, ge, gc, gi = self.graphs
        gn.x = self._next_x() if e2e else self.x_dev0.clone()   # the forward normalises graph_node.x in place
        gn.norm_uvp, gn.norm_global = True, True
        self.flat_grad.zero_()
        out = self.model(gn, gx, ge, gc, gi, is_training=True)
        loss = self.script_loss(out)
        loss.backward()
        if self.cells_mode:
            self.parallel.sum_gradients(self.flat_grad)
        elif self.world > 1:
            self.parallel.allreduce_gradients(self.flat_grad, self.world)
        self.opt.step()
        if e2e:
            self._results_to_host(out[4], loss.detach())
        return loss

    def step(self, e2e):
        if self.gstep is None:
            return self.eager_step(e2e)
        loss = self.gstep.step(self.x_host if e2e else None)
        if e2e:
            self.uvp_host.copy_(self.gstep.out[4], non_blocking=True)
            self.loss_host.copy_(loss.detach(), non_blocking=True)
        return loss

    def timed(self, nsteps, e2e, profiler_range=False):
        """ms for nsteps steps: CUDA events on the launching stream, barrier + synchronize on both sides, max over ranks."""
        import torch.distributed as dist
        if self.world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        if profiler_range:
            torch.cuda.profiler.start()  # `ncu --profile-from-start off` then lists exactly the launches of the timed steps
        for _ in range(nsteps):
            self.step(e2e)
        if profiler_range:
            torch.cuda.synchronize()
            torch.cuda.profiler.stop()
        ev1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([ev0.elapsed_time(ev1)], device=self.dev)
        if self.world > 1:
            dist.barrier()
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    def close(self):
        if self.gstep is not None:
            self.gstep.close()
        self.gstep = self.model = self.opt = self.flat_grad = self.graphs = self.plan = None
        import gc
        gc.collect()
        torch.cuda.empty_cache()


def quick_rate(dev, rank, world, graphs, net, mp, precision, steps, cells_total, graph=False, **kw):
    """cells*steps/s of a secondary configuration (sub-records of the JSON line): 3 warm-up + `steps` timed steps."""
    job = Job(dev, rank, world, graphs, net, mp, precision, graph=graph, **kw)
    for _ in range(3):
        job.step(False)
    ms = job.timed(steps, False)
    loss = float(job.step(False).detach())
    job.close()
    return {"value": cells_total / (ms / 1e3 / steps), "unit": "cells*steps/s", "ms_per_step": ms / steps, "steps": steps, "loss": loss}


def example_mesh_records(dev, precision, steps):
    """Throughput on the reference's own example meshes (BASELINE.json configs 0-3), taken from the golden fixtures
    (tests/golden/*.npz hold the meshes as parsed by the reference's COMSOL / Tecplot parsers; /root/reference is not
    needed): the default net TransFVGN_v2 (6 GnBlocks + 2 Transolver blocks), eager and as one CUDA graph per step.  The
    airfoil case is BASELINE config 3: a batch of 64 graphs (N = 1.08 M nodes) -- the two parameter draws the reference's
    sampler made for the golden case, 32 copies each."""
    from gen_fvgn_steady_b200.mesh.batching import graphs_from_meshes
    from tests import golden_util as GU
    out = {}
    for name, label, copies in (("lid_cavity_101_v2", "lid_driven_cavity_101x101-Re=100", 1),
                                ("cylinder_tri_quad_v1", "cylinder_flow_tri_quad", 1),
                                ("cylinder_poly_v1", "cylinder_flow_poly (polygon cells)", 1),
                                ("poisson_quad_tri_v2", "poisson/cavity_poisson_quad_tri", 1),
                                ("airfoil_naca0012_b2_v2", "airfoil_L=1/NACA0012, batch of 64 graphs", 32)):
        path = os.path.join(GU.GOLDEN_DIR, name + ".npz")
        if not os.path.exists(path):
            continue
        z = GU.load_case(name)
        meshes, uvps = GU.example_case_graphs(z)
        meshes, uvps = meshes * copies, uvps * copies
        rec = {}
        for graph in (False, True):
            graphs = graphs_from_meshes(meshes, uvps, dev)
            C = int(graphs[3].pos.shape[0])
            r = quick_rate(dev, 0, 1, graphs, "TransFVGN_v2", 3, precision, steps, C, graph=graph)
            rec["cuda_graph" if graph else "eager"] = {"value": r["value"], "ms_per_step": r["ms_per_step"]}
            rec.update(graphs=len(meshes), cells=C, nodes=int(graphs[0].pos.shape[0]))
            del graphs
        out[label] = rec
    return out


def grad_rec_record(dev):
    """BASELINE.json config 1: the loop of src/grad_rec_speed_test.py:118-159 (node_based_WLSQ on one mesh, scalar field,
    precomputed moments) at the size of the cavity_poisson_81x81 example and at 1 M nodes, eager and as a CUDA graph."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import grad_rec_speed as G
    hbm, _ = peaks()
    return [G.measure(n, dev, hbm, cg) for n in (80, 1000) for cg in (False, True)]


def loader_regime_record(dev, precision, steps, cells=100_000):
    """SURVEY 8(f) row f4 at the size of the reference's larger example meshes: the same EPD step (a) on a resident batch,
    (b) on a fresh batch object per step from the device pool (gen_fvgn_steady_b200.pool: sample + payback, plan found on
    the batch), (c) on a fresh batch with NEW tensors per step, as a foreign loader produces (plan found by content hash)."""
    from gen_fvgn_steady_b200.FVMmodel.importer import NNmodel
    from gen_fvgn_steady_b200.mesh.batching import graphs_from_meshes
    from gen_fvgn_steady_b200.pool import DevicePool
    from gen_fvgn_steady_b200.utils.get_param import params as default_params
    mesh, uvp = make_mesh(cells, 0, dev)
    p = default_params(net="EPD", message_passing_num=6, precision=precision, dataset_size=1)
    torch.manual_seed(0)
    model = NNmodel(p).to(dev)
    opt = torch.optim.Adam(model.parameters(), lr=p.lr, fused=True)
    pool = DevicePool([mesh], [uvp], dev)
    resident = graphs_from_meshes([mesh], [uvp], dev)
    x0 = resident[0].x.clone()

    def one(graphs):
        opt.zero_grad(set_to_none=True)
        out = model(*graphs, is_training=True)
        loss = torch.mean(torch.log(p.loss_press * out[3] + p.loss_cont * out[0] + p.loss_mom * out[1] + p.loss_mom * out[2]))
        loss.backward()
        opt.step()
        return out

    def step_resident():
        resident[0].x, resident[0].norm_uvp, resident[0].norm_global = x0.clone(), True, True
        one(resident)

    def step_pool():
        graphs, gidx = pool.sample([0])
        pool.payback(one(graphs)[4], gidx)

    def step_foreign():
        one(graphs_from_meshes([mesh], [uvp], dev))

    rec = {"cells": int(resident[3].pos.shape[0])}
    for name, fn in (("resident", step_resident), ("device_pool", step_pool), ("fresh_tensors_content_hash", step_foreign)):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(steps):
            fn()
        ev1.record()
        torch.cuda.synchronize()
        rec[name] = {"ms_per_step": ev0.elapsed_time(ev1) / steps}
    for k in ("device_pool", "fresh_tensors_content_hash"):
        rec[k]["vs_resident"] = rec[k]["ms_per_step"] / rec["resident"]["ms_per_step"]
    return rec


def run_ours(args):
    import torch.distributed as dist
    from gen_fvgn_steady_b200 import _lib
    from gen_fvgn_steady_b200.mesh.batching import graphs_from_meshes

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py --impl ours needs a CUDA device: this package has no CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # the reference's scripts run their (cuBLAS) Linear layers in TF32 (src/pre_train_Adam.py:29); only the library GEMMs
    # of the Transolver blocks of --net TransFVGN_v* are affected here, the GN / FV kernels never go through cuBLAS
    torch.set_float32_matmul_precision("high")
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def partitioned_graphs(cells, halo_layers):
        from gen_fvgn_steady_b200 import partition
        mesh, uvp = make_mesh(cells, 0, dev)            # the same global mesh on every rank
        c_global = int(mesh["cell|centroid"].shape[0])
        mesh, uvp, halo = partition.build(mesh, uvp, world, rank, halo_layers=halo_layers, device=dev)
        torch.cuda.empty_cache()
        graphs = graphs_from_meshes([mesh], [uvp], dev)
        partition.mark_partition(graphs, halo)
        return graphs, halo, c_global

    cells_mode = args.parallel == "cells" and world > 1
    gn_blocks_total = _gn_blocks(args.net, args.mp)
    if args.halo_layers is None:
        args.halo_layers = 3 * gn_blocks_total + 2
    halo = None
    if cells_mode:
        graphs, halo, C_global = partitioned_graphs(args.cells, args.halo_layers)
    else:
        mesh, uvp = make_mesh(args.cells, rank, dev)
        graphs = graphs_from_meshes([mesh], [uvp], dev)
        del mesh
    job = Job(dev, rank, world, graphs, args.net, args.mp, args.precision, cells_mode=cells_mode, halo=halo, graph=args.graph,
              warmup=args.warmup)
    N, E, C, K, X = job.sizes
    cells_total = C_global if cells_mode else world * C

    # untimed warm-up: at least 6 steps whatever --warmup says (allocator growth and lazy NCCL channel set-up take more than 3 steps
    # to settle on 8 GPUs: with 3, the first timed region of an 8-GPU run came out 4-15 % slower than the one after it)
    for _ in range(max(args.warmup, 6)):
        job.step(False)
    launches_per_replay = None
    if job.gstep is not None:  # count the kernels of one step with the eager path (the graph replays exactly those)
        l_ = _lib.launch_count
        job.eager_step(False)
        launches_per_replay = _lib.launch_count - l_
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = _lib.launch_count
    ms = job.timed(args.steps, False, profiler_range=True)
    launches = _lib.launch_count - l0 if job.gstep is None else launches_per_replay * args.steps
    job.step(True)
    ms_e2e = job.timed(args.steps, True)
    clocks = sampler.stop() if rank == 0 else None
    last_loss = float(job.loss_host)

    # kernels of this library per step, counted from one profiled step outside the timed regions (a C-ABI call launches 1-5
    # kernels; `launches` above counts the calls)
    own_kernels = lib_kernels = None
    from torch.profiler import profile, ProfilerActivity
    prof = None
    if rank == 0:
        try:
            prof = profile(activities=[ProfilerActivity.CUDA])
            prof.__enter__()
        except Exception:  # noqa: BLE001 -- e.g. another CUPTI subscriber (ncu) owns the device
            prof = None
    for _ in range(2):   # every rank steps (collectives); rank 0 records
        job.step(False)
    torch.cuda.synchronize()
    if prof is not None:
        try:
            prof.__exit__(None, None, None)
        except Exception:  # noqa: BLE001
            prof = None
    if rank == 0 and prof is not None:
        try:
            evs = prof.key_averages()
        except Exception:  # noqa: BLE001
            evs = []
        evs = [e for e in evs if e.device_time_total > 0]   # device activities only (the list also holds cudaLaunchKernel etc.)
        is_lib = lambda k: k.startswith("void at::") or k.startswith("at::") or "cutlass" in k or "cublas" in k.lower() or \
            k.startswith("Memcpy") or k.startswith("Memset") or "nccl" in k.lower() or "sm90" in k or "sm100" in k or \
            "gemv" in k or "vectorized" in k
        own_kernels = sum(e.count for e in evs if not is_lib(e.key)) // 2
        lib_kernels = sum(e.count for e in evs if is_lib(e.key) and not e.key.startswith("Mem")) // 2
    if args.kernel_summary and rank == 0 and prof is not None:
        rows = sorted(((e.key, e.count, e.device_time_total / 1e3) for e in prof.key_averages()), key=lambda r: -r[2])
        tot = sum(r[2] for r in rows)
        with open(args.kernel_summary, "w") as f:
            f.write(f"# torch.profiler (CUPTI) device time over 2 steps, in situ (concurrent, warm): total {tot:.3f} ms\n")
            for k, c, kms in rows:
                f.write(f"{kms:10.3f} ms {c:6d} {100 * kms / max(tot, 1e-9):5.1f}%  {k[:150]}\n")

    # dominant kernel: timed alone with CUDA events on the launching stream
    roof = dominant_kernel_roofline(job.model, job.plan, dev, args, job.p)
    halo_text = None
    if cells_mode:
        halo_text = (f"cells{world} (one {C_global}-cell mesh, RCB partition, {args.halo_layers}-layer halo, "
                     f"{sum(halo.wants_exchange(i, gn_blocks_total) for i in range(gn_blocks_total))} ghost refreshes per forward; "
                     f"rank 0: {halo.n_owned_cells} owned of {C} local cells)")
    job.close()

    hbm, how = peaks()
    sec = ms / 1e3 / args.steps
    ab = alg_bytes_step(N, E, C, K, X, gn_blocks_total)   # SURVEY 8(d) counts the GN + FV bytes only (Transolver blocks add none)
    line = {"metric": "cells*steps/sec (fwd+bwd train step)", "value": cells_total / sec, "unit": "cells*steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong" if cells_mode else "weak", "vs_baseline": None, "dtype": DTYPE_TEXT[args.precision],
            "data": "synthetic", "config": dict(workload_config(args), N=N, E=E, C=C, K=K, X=X,
                                                parallelism=halo_text if cells_mode else f"dp{world}", loss=last_loss),
            # kernels of this library launched in the timed region; when no profiler record is available (e.g. under ncu) the
            # number of C-ABI calls, each of which launches at least one kernel
            "clocks": clocks, "gpu_launches": own_kernels * args.steps if own_kernels else launches,
            "gpu_launches_detail": {"own_kernels_per_step": own_kernels, "c_abi_calls_in_timed_region": launches,
                                    "pytorch_elementwise_kernels_per_step": lib_kernels,
                                    "how": "kernel launches of one profiled step (CUPTI) outside the timed regions, by kernel name"},
            "e2e": {"value": cells_total / (ms_e2e / 1e3 / args.steps), "unit": "cells*steps/s",
                    "h2d_bytes_per_step": N * 12 * 4,
                    "d2h_bytes_per_step": N * 3 * 4 + 4, "ms_per_step": ms_e2e / args.steps},
            "roofline": dict(roof, peak=hbm, frac=roof["achieved"] / hbm, peak_source=how),
            "step_roofline": {"alg_bytes_per_step": ab, "achieved_gbs": ab / sec / 1e9, "frac": ab / sec / 1e9 / hbm,
                              "alg_kb_per_cell": ab / C / 1e3}}

    # ------------------------------------------------------------------ sub-records (secondary configurations)
    if not args.no_extras and not cells_mode and not args.graph:
        sub_steps = max(min(args.steps, 10), 3)
        if world == 1:
            # every tensor-core mode bench.py can time, on the same mesh (tests/test_gpu_parity.py asserts all goldens in both)
            modes = {args.precision: {"value": line["value"], "unit": "cells*steps/s", "ms_per_step": line["ms_per_step"],
                                      "steps": args.steps}}
            for prec in ("f16", "bf16"):
                if prec not in modes:
                    modes[prec] = quick_rate(dev, rank, world, graphs, args.net, args.mp, prec, sub_steps, C)
            line["precision_modes"] = modes
            # the reference's default net (get_param.py:37): 2 x (3 GnBlocks + Transolver block), same mesh
            if args.net != "TransFVGN_v2":
                r = quick_rate(dev, rank, world, graphs, "TransFVGN_v2", 3, args.precision, sub_steps, C)
                ab2 = alg_bytes_step(N, E, C, K, X, 6)
                r.update(gn_blocks=6, transolver_blocks=2, step_roofline_frac=ab2 / (r["ms_per_step"] / 1e3) / 1e9 / hbm)
                line["nets"] = {"TransFVGN_v2": r}
            del graphs
            torch.cuda.empty_cache()
            def size_sweep():
                """north_star: 'synthetic meshes scaled to 1M-16M cells' -- the sizes one GPU holds (16 M cells needs the cell
                partition: the N > 1 `cells` record); same net / precision / loss as the headline."""
                out = {}
                for c in (1_000_000, 2_000_000, 8_000_000):
                    torch.cuda.empty_cache()
                    torch.cuda.reset_peak_memory_stats(dev)
                    m_, u_ = make_mesh(c, 0, dev)
                    g_ = graphs_from_meshes([m_], [u_], dev)
                    del m_, u_
                    c_real = int(g_[3].pos.shape[0])
                    r_ = quick_rate(dev, 0, 1, g_, args.net, args.mp, args.precision, 5, c_real)
                    r_["cells"] = c_real
                    r_["peak_memory_gb"] = torch.cuda.max_memory_allocated(dev) / 1e9
                    out[f"{c // 1_000_000}M"] = r_
                    del g_
                    torch.cuda.empty_cache()
                    torch.cuda.reset_peak_memory_stats(dev)
                return out

            for key, fn in (("size_sweep", size_sweep),
                            ("example_meshes", lambda: example_mesh_records(dev, args.precision, 20)),
                            ("loader_regime", lambda: loader_regime_record(dev, args.precision, 20)),
                            ("grad_rec_speed", lambda: grad_rec_record(dev))):
                try:
                    line[key] = fn()
                except Exception as e:  # noqa: BLE001 -- a sub-record must not take the headline down
                    line[key] = {"error": repr(e)}
        else:
            # strong scaling of ONE mesh of --cells cells over the ranks: cell partition + halo
            del graphs
            torch.cuda.empty_cache()
            recs = {}
            for hl, use_graph in ((3 * gn_blocks_total + 2, False), (3 * gn_blocks_total + 2, True), (3, False)):
                g2, h2, cg = partitioned_graphs(args.cells, hl)
                owned = torch.tensor([h2.n_owned_cells, int(g2[3].pos.shape[0])], device=dev, dtype=torch.int64)
                allv = [torch.zeros_like(owned) for _ in range(world)]
                dist.all_gather(allv, owned)
                r = quick_rate(dev, rank, world, g2, args.net, args.mp, args.precision, sub_steps, cg, cells_mode=True, halo=h2,
                               graph=use_graph)
                r.update(halo_layers=hl, cuda_graph=use_graph,
                         ghost_refreshes_per_forward=sum(h2.wants_exchange(i, gn_blocks_total) for i in range(gn_blocks_total)),
                         owned_cells_per_rank=[int(v[0]) for v in allv], local_cells_per_rank=[int(v[1]) for v in allv])
                recs[f"halo{hl}" + ("_cuda_graph" if use_graph else "")] = r
                del g2, h2
                torch.cuda.empty_cache()
            line["cells"] = {"mesh_cells": cg, "scaling": "strong", "partition": "recursive coordinate bisection", **recs}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    if not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        cells_cpu = args.cpu_cells
        if cells_cpu <= 0:
            cells_cpu, _ = cpu_sample_cells(args.net, args.mp, threads, 6, 20.0)
        times, C_cpu = cpu_reference_step_times(cells_cpu, args.net, args.mp, 5, 1, threads)
        sec_cpu = float(np.median(times))
        line["cpu_baseline"] = {"value": C_cpu / sec_cpu, "unit": "cells*steps/s", "cores": threads, "kind": "port",
                                "sample": f"{C_cpu}-cell synthetic quad mesh (same generator, net, G, loss), fwd+bwd+Adam fp32, median "
                                          f"of 5 steps after 1 warm-up (oracle/fvgn_oracle.py on the host cores)"}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# DRAM bytes per edge of the kernels of one fused edge-MLP backward call (dram__bytes_read.sum + dram__bytes_write.sum of ncu
# --set full captures, 8 M edges, N = E/2, f16; the bf16 kernels move the same bytes) -> (bytes per edge, source):
#   edge-level: A 7.312526 + 2.032588 GB, B 9.492903 + 8.173909 GB = 27.011926 GB / 8e6 edges
NCU_TRAFFIC_BYTES_PER_EDGE = {
    "tc_edge_level": (27.011926e9 / 8.0e6, "profiles/r2p_ncu_tc_kernels_EDGE_8Medges.csv, per edge x E"),
    # node-level layer 1 (default): A 6.774442 + 2.031280, B<KB0=4> 8.196782 + 4.080255, dz_incidence 2.435244 + 2.033738,
    # node GEMM 3.075269 + 1.008122 GB = 29.635132 GB / 8.004e6 edges (this call ends in d(agg) [N,128]; the edge-level call
    # above is followed by a separate 5 GB incidence reduction)
    "tc_node_level": (29.635132e9 / 8.004e6, "profiles/r2y_ncu_edge_bwd_node_level_8Medges.csv, per edge x E"),
    # the same with 16-bit gradient streams in and out (f16 mode, inner blocks): A 4.739270 + 2.028844, B<KB0=4,GS=1> 6.147607 +
    # 2.035387, dz_incidence 2.304062 + 2.010787, node GEMM 3.075331 + 1.008751 GB = 23.350039 GB
    "tc_node_level_grad16": (23.350039e9 / 8.004e6, "profiles/r2ao_ncu_edge_bwd_grad16_8Medges.csv, per edge x E"),
    "fp32": None,
}


def edge_backward_runner(plan, dev, prec, params):
    """-> (run, label): one fused edge-MLP backward call of a GnBlock on random inputs of the plan's size (the product's own
    call: node-level layer 1 unless FVGN_NODE_LEVEL_LAYER1=0).  Also used by tools/edge_bwd_profile.py under ncu."""
    from gen_fvgn_steady_b200 import _lib, ops
    N, E = plan.N, plan.E
    bf = ops.is_tc(prec)
    hdt = ops.HDTYPE.get(prec)
    gen = torch.Generator(device=dev).manual_seed(1)
    agg = torch.randn((N, 128), device=dev, generator=gen)
    e = torch.randn((E, 128), device=dev, generator=gen)
    d_out = torch.randn((E, 128), device=dev, generator=gen)
    d_a1 = torch.randn((N, 64), device=dev, generator=gen)
    d_e = torch.empty((E, 128), device=dev)
    code = _lib.FVGN_MLP_EDGE
    if not bf:
        d_sr = torch.empty((E, 256), device=dev)
        return (lambda: ops.mlp_backward(code, "fp32", E, params, agg, e, plan.edge_s, plan.edge_r, d_out, d_a1, d_sr, d_e),
                "mlp_bwd_kernel<EDGE> (fused edge-MLP backward: recompute + dgrad + wgrad)")
    aggh, eh = ops.shadow(agg, dtype=hdt), ops.shadow(e, dtype=hdt)
    del agg, e
    z1 = ops.new_z1(code, prec, E, d_out)
    ops.mlp_forward(code, prec, E, params, None, None, plan.edge_s, plan.edge_r, want_out=False, want_res=False, z1=z1,
                    in0h=aggh, in1h=eh, want_outh=True)
    d_a1h = d_a1.to(hdt)
    if ops.NODE_LEVEL_LAYER1:
        d_agg = torch.empty((N, 128), device=dev, dtype=hdt)
        if prec == "f16" and ops.GRAD16 and ops.LATENTS16:   # the inner blocks of a model: 16-bit gradient streams in and out
            d_outh, d_eh = d_out.to(hdt), torch.empty((E, 128), device=dev, dtype=hdt)
            del d_out, d_e
            return (lambda: ops.mlp_backward(code, prec, E, params, None, None, plan.edge_s, plan.edge_r, None, None, None, None,
                                             z1=z1, in0h=aggh, in1h=eh, d_gatherh=d_a1h, node_path=(plan, d_agg), d_outh=d_outh,
                                             d_in1h=d_eh),
                    "fvgn_mlp_backward<EDGE>, node-level layer 1, 16-bit gradient streams = mlp_tc_bwd_a_kernel<0> + "
                    "mlp_tc_bwd_b_kernel<0,KB0=4,GS=1> + dz_incidence_kernel + mlp_tc_bwd_node_kernel (tcgen05)")
        return (lambda: ops.mlp_backward(code, prec, E, params, None, None, plan.edge_s, plan.edge_r, d_out, None, None, d_e,
                                         z1=z1, in0h=aggh, in1h=eh, d_gatherh=d_a1h, node_path=(plan, d_agg)),
                "fvgn_mlp_backward<EDGE>, node-level layer 1 = mlp_tc_bwd_a_kernel<0> + mlp_tc_bwd_b_kernel<0,KB0=4> + "
                "dz_incidence_kernel + mlp_tc_bwd_node_kernel (tcgen05)")
    d_srh = torch.empty((E, 256), device=dev, dtype=hdt)
    return (lambda: ops.mlp_backward(code, prec, E, params, None, None, plan.edge_s, plan.edge_r, d_out, None, None, d_e,
                                     z1=z1, in0h=aggh, in1h=eh, d_in0h=d_srh, d_gatherh=d_a1h),
            "fvgn_mlp_backward<EDGE> = mlp_tc_bwd_a_kernel<0> + mlp_tc_bwd_b_kernel<0> (tcgen05)")


def dominant_kernel_roofline(model, plan, dev, args, p):
    """Times the dominant call of the step -- the fused edge-MLP backward of one GnBlock (tcgen05 kernels and the
    deterministic partial reductions) -- alone, with CUDA events on its stream, on inputs of the step's own size.
    Algorithmic bytes (SURVEY.md section 8(d) accounting: fp32 volumes of the reference's tensors, every tensor row counted
    once): e and d_e_out read, d_e written (3 x 512 B per edge); agg read, d_a1 read, d(agg) written per NODE
    (512 + 256 + 512 B).  With the node-level layer 1 (default) the call really ends in d(agg) [N,128]; the edge-level
    variant (FVGN_NODE_LEVEL_LAYER1=0) ends in the per-edge stream d(agg[s]) | d(agg[r]), whose incidence reduction is then a
    separate launch outside this timing."""
    from gen_fvgn_steady_b200 import ops
    from gen_fvgn_steady_b200.FVMmodel.Models.FVGN.blocks import mlp_params
    blk = None
    for m in model.modules():
        if m.__class__.__name__ == "GnBlock":
            blk = m
            break
    N, E = plan.N, plan.E
    bf = ops.is_tc(args.precision)
    params = [q.detach() for q in mlp_params(blk.eb_module.net)]
    run, label = edge_backward_runner(plan, dev, args.precision, params)
    reps = 5
    times = []
    for i in range(reps + 2):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        run()
        ev1.record()
        torch.cuda.synchronize()
        if i > 1:
            times.append(ev0.elapsed_time(ev1))
    ms = float(np.mean(times))
    alg = E * (3 * 512) + N * (512 + 256 + 512)
    key = "fp32" if not bf else ("tc_edge_level" if not ops.NODE_LEVEL_LAYER1 else
                                 "tc_node_level_grad16" if "16-bit gradient" in label else "tc_node_level")
    tpe = NCU_TRAFFIC_BYTES_PER_EDGE.get(key)
    return {"kernel": label, "bound": "hbm",
            "achieved": alg / (ms / 1e3) / 1e9, "unit": "GB/s", "ms_per_launch": ms, "alg_bytes_per_launch": alg,
            "traffic": None if tpe is None else tpe[0] * E,
            "traffic_source": None if tpe is None else tpe[1]}


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
