#!/usr/bin/env python
"""bench.py — tet-constraint projections/s of the XPBD hot path on B200.

One "step" = one frame of the hot path (timestep_t::step: detection, 10 substeps x 10
Gauss-Seidel iterations, commit, surface update) over one batch of synthetic input.

  N = 1 : BASELINE.json configs[2] — the 1M-tet grid-tetrahedralised block dropped on an SDF
          sphere + floor, collision detection every substep ("config3").
  N > 1 : the path shards by independent scenes (north_star (3): "independent bodies or scenes
          split across GPUs"): every rank runs its own config3 scene, no data-path collective,
          weak scaling.  `--workload config4` shards the 4096-body ensemble instead
          (strong scaling, 4096/N bodies per rank); `--workload config5` runs the 8M-tet body
          on one GPU.

`value`  : projections/s with state resident in HBM (CUDA events on the launch stream,
           per-step event pairs, L2 flushed between steps, max over ranks).
`e2e`    : the same metric through the C ABI with HOST buffers (sbsb200_step_host: pinned
           host x,v -> device, step, device -> host x,v inside the timed region).
`--impl reference` times the reference's CPU algorithm (oracle/_ref when built, else the C
port in oracle/) on a bounded sample of the same workload, rank 0 only.
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "tet_constraint_projections_per_sec"
UNIT = "projections/s"
BYTES_PER_PROJECTION = 176   # SURVEY.md §8(d), fp32 build
BYTES_PER_COLLISION = 64
BYTES_PER_VERTEX_SUBSTEP = 112
BYTES_PER_SURFACE_DETECT = 16


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="config3", choices=["config1", "config2", "config3", "config4", "config5"])
    ap.add_argument("--precision", type=int, default=32, choices=[32, 64])
    ap.add_argument("--schedule", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--decompose", action="store_true",
                    help="N > 1: cut the ONE body of the workload over the N GPUs (default for config5)")
    return ap.parse_args()


def make_scene(sc, workload, rank, world, decompose=False):
    if workload == "config1":
        return sc.config1()
    if workload == "config2":
        return sc.config2()
    if workload == "config3":
        return sc.config3(seed=3 + (0 if decompose else 100 * rank))
    if workload == "config5":
        return sc.config5()
    per = 4096 // world
    return sc.config4(per, first=rank * per)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
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
        sm, smax, reasons = [], [], set()
        for r in self.rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1]))
                smax.append(float(r[2]))
            except ValueError:
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7),
                              ("sw_power_cap", 8)):
                if r[col].lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def cpu_baseline(sc, workload, rank=0, world=1):
    """Reference algorithm on the host: ONE substep (10 iterations) of the same scene, serial."""
    from oracle import oracle as O
    kind, W = "port", None
    try:
        from oracle import ref as REF
        if REF.available():
            kind, W = "reference", REF.World()
    except Exception:
        W = None
    if W is None:
        O.build()
        W = O.World()
    scene = make_scene(sc, workload, rank, world)
    scene.instantiate(W)
    dt = scene.dt / scene.substeps
    t0 = time.perf_counter()
    W.step(dt, 1, scene.iterations, False)
    sec = time.perf_counter() - t0
    proj = scene.n_tets * scene.iterations
    return {"value": proj / sec, "unit": UNIT, "cores": 1, "kind": kind,
            "sample": "1 substep (%d iterations, %d projections) of %s in %.2f s, serial Gauss-Seidel "
                      "as the reference (single-threaded), fp64" % (scene.iterations, proj, scene.name, sec),
            "seconds": sec, "host_cores_available": os.cpu_count()}, scene


def run_reference(args, rank, world):
    if rank != 0:
        return
    sc = importlib.import_module("soft-body-simulator_b200.scenes")
    from oracle import oracle as O
    kind, mk = "port", None
    try:
        from oracle import ref as REF
        if REF.available():
            kind, mk = "reference", REF.World
    except Exception:
        mk = None
    if mk is None:
        O.build()
        mk = O.World
    scene = make_scene(sc, args.workload, 0, world)
    W = mk()
    scene.instantiate(W)
    dt = scene.dt / scene.substeps
    for _ in range(args.warmup):
        W.step(dt, 1, 1, False)          # warm caches; 1 iteration each
    t0 = time.perf_counter()
    for _ in range(args.steps):
        W.step(dt, 1, scene.iterations, False)
    sec = time.perf_counter() - t0
    proj = scene.n_tets * scene.iterations * args.steps
    value = proj / sec
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sec / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": scene.name, "tets": scene.n_tets, "substeps": scene.substeps,
                       "iterations": scene.iterations,
                       "step": "bounded sample: each step is ONE substep (%d iterations) of the frame" % scene.iterations},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": kind,
                             "sample": "%d x 1 substep of %s, serial (the reference solver is single-threaded)"
                                       % (args.steps, scene.name)},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the XPBD path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    sbs = importlib.import_module("soft-body-simulator_b200")
    sc = importlib.import_module("soft-body-simulator_b200.scenes")

    decomposed = world > 1 and (args.workload == "config5" or (args.decompose and args.workload in ("config2", "config3")))
    scene = make_scene(sc, args.workload, rank, world, decomposed)
    # a dedicated non-blocking stream: the library captures the frame into a CUDA graph, which
    # the legacy default stream does not permit
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    sim = sbs.Simulation(local, args.precision, stream=stream.cuda_stream,
                         schedule=sbs.SCHED_PERSISTENT if decomposed else args.schedule)
    if decomposed:
        # ONE body cut into world x regions; every rank runs its block, shared vertices travel through
        # mailboxes in peer memory (NVLink stores issued by the substep kernel; no collective)
        ids = scene.instantiate(sim, partition=(rank, world))
        handles = [None] * world
        dist.all_gather_object(handles, sim.mailbox_handle())
        sim.connect_peers(handles)
        dist.barrier()
    else:
        ids = scene.instantiate(sim)
    stats0 = sim.stats()
    S, K = scene.substeps, scene.iterations
    proj_per_step = scene.n_tets * S * K

    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing ----------------------------------------------------------
    for _ in range(args.warmup):
        sim.step(scene.dt, S, K, scene.detect_every_substep)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    st_before = sim.stats()
    launches0 = st_before["kernels_launched"]
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    contacts = 0
    for a, b in ev:
        flush.zero_()                      # L2 flush between timed iterations (not timed)
        a.record(stream)
        sim.step(scene.dt, S, K, scene.detect_every_substep)
        b.record(stream)
    barrier()
    st_after = sim.stats()
    launches = st_after["kernels_launched"] - launches0
    kernel_ms = st_after["kernel_ms"] - st_before["kernel_ms"]
    kernel_launches = st_after["kernel_launches"] - st_before["kernel_launches"]
    ms = sum(a.elapsed_time(b) for a, b in ev)
    contacts = len(sim.contacts()[0])
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    total_proj = proj_per_step * args.steps * (world if args.workload in ("config3", "config4") and not decomposed else 1)
    value = total_proj / (ms_max * 1e-3)

    # ---- end-to-end through the C ABI with host buffers -----------------------------------
    e2e = None
    if not args.no_e2e:
        b0 = scene.tet_bodies()[0]
        nV = scene.items[b0].x0.shape[0]
        x_in = torch.from_numpy(scene.items[b0].x.copy()).pin_memory()
        v_in = torch.zeros((nV, 3), dtype=torch.float64).pin_memory()
        x_out = torch.empty((nV, 3), dtype=torch.float64).pin_memory()
        v_out = torch.empty((nV, 3), dtype=torch.float64).pin_memory()
        xin, vin, xo, vo = x_in.numpy(), v_in.numpy(), x_out.numpy(), v_out.numpy()
        for _ in range(2):
            sim.step_host(ids[b0], xin, vin, scene.dt, S, K, scene.detect_every_substep, xo, vo)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n_e2e = max(3, args.steps // 2)
        t0 = time.perf_counter()
        e0.record(stream)
        for _ in range(n_e2e):
            sim.step_host(ids[b0], xin, vin, scene.dt, S, K, scene.detect_every_substep, xo, vo)
            xin, xo = xo, xin              # next frame continues from the host copy: the two pinned
            vin, vo = vo, vin              # buffer pairs swap roles, nothing is copied on the host
        e1.record(stream)
        barrier()
        wall = time.perf_counter() - t0
        t = torch.tensor([wall], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        scale = world if args.workload in ("config3", "config4") and not decomposed else 1
        e2e = {"value": proj_per_step * n_e2e * scale / float(t.item()), "unit": UNIT,
               "h2d_bytes_per_step": int(2 * nV * 24), "d2h_bytes_per_step": int(2 * nV * 24),
               "steps": n_e2e, "timing": "wall clock around sbsb200_step_host (max over ranks)",
               "bodies_round_tripped": 1}
    clocks = sampler.stop()

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        nVtot, nVs = stats0["n_vertices"], stats0["n_surface_vertices"]
        n_det = S if scene.detect_every_substep else 1
        bytes_step = S * K * (BYTES_PER_PROJECTION * scene.n_tets + BYTES_PER_COLLISION * contacts) \
            + S * BYTES_PER_VERTEX_SUBSTEP * nVtot + n_det * BYTES_PER_SURFACE_DETECT * nVs
        if args.precision == 64:
            bytes_step = bytes_step * 336 // 176
        ms_step = ms_max / args.steps
        # per-GPU figure: a rank of a decomposed body moves 1/world of the frame's bytes
        frame_gbs = bytes_step / (world if decomposed else 1) / (ms_step * 1e-3) / 1e9
        sched = {1: "graph", 2: "persistent"}.get(stats0["schedule"], "?")
        if kernel_launches > 0:
            # dominant kernel = the substep kernel of the persistent schedule: one launch runs predict,
            # K sweeps over every tet and contact, and commit for one substep.  Timed live with CUDA
            # events around every launch of the timed region (sbsb200_stats.kernel_ms).
            share = world if decomposed else 1        # a rank of a decomposed body runs 1/world of it
            bytes_launch = (K * (BYTES_PER_PROJECTION * scene.n_tets + BYTES_PER_COLLISION * contacts)
                            + BYTES_PER_VERTEX_SUBSTEP * nVtot) // share
            if args.precision == 64:
                bytes_launch = bytes_launch * 336 // 176
            k_ms = kernel_ms / kernel_launches
            achieved = bytes_launch / (k_ms * 1e-3) / 1e9
            kernel = {"name": "k_substep_persistent", "launches_timed": int(kernel_launches),
                      "avg_ms_per_launch": k_ms, "algorithmic_bytes_per_launch": int(bytes_launch),
                      "share_of_step": kernel_ms / ms}
        else:
            # graph schedule: ~800 k_project_green launches per frame inside one CUDA graph; CUDA events
            # cannot bracket a node of a graph launch, so the frame as a whole is the timed unit
            achieved, kernel = frame_gbs, {"name": "whole frame (CUDA graph of per-colour kernels)"}
        traffic = None
        try:   # dram__bytes_read + dram__bytes_write per launch of the same kernel on the same workload (ncu)
            tr = json.load(open(os.path.join(ROOT, "profiles", "r01_ncu_traffic.json")))
            if tr["workload"] == scene.name and kernel["name"] in tr["kernel"]:
                traffic = tr["dram_bytes_read"] + tr["dram_bytes_write"]
        except Exception:
            pass
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "strong" if args.workload in ("config4", "config5") or decomposed else "weak", "vs_baseline": None,
            "dtype": "f32" if args.precision == 32 else "f64", "data": "synthetic",
            "config": {"workload": scene.name, "tets_per_gpu": scene.n_tets // (world if decomposed else 1),
                       "vertices_per_gpu": nVtot // (world if decomposed else 1),
                       "substeps": S, "iterations": K, "dt": scene.dt,
                       "detection": ("every substep" if scene.detect_every_substep else "once per frame")
                       + (", BVH broadphase" if scene.broadphase else ""),
                       "colours": stats0["n_green_colours"], "schedule": sched,
                       "regions": stats0["n_regions"], "shared_vertices": stats0["n_interface_vertices"],
                       "parallelism": ("one body decomposed over %d GPUs, shared vertices pushed to peer memory by "
                                       "the substep kernel" % world) if decomposed else
                                      "scenes sharded over %d GPU(s), no collective" % world,
                       "l2": "256 MiB write between timed steps (flush)"},
            "ms_per_frame": ms_step,
            "contacts_last_detection": contacts,
            "clocks": clocks, "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "6650 (of fallback)",
                         "kernel": kernel,
                         "frame": {"algorithmic_bytes_per_step": int(bytes_step), "achieved": frame_gbs,
                                   "frac": frame_gbs / peak, "per": "GPU"},
                         "frac_of_8TBs_nominal": achieved / 8000.0,
                         "note": "algorithmic bytes (SURVEY 8d: 176 B per projection, 64 B per contact "
                                 "projection, 112 B per vertex and substep) / CUDA-event time; the working set "
                                 "fits the 126 MB L2, so DRAM traffic is far below the algorithmic bytes "
                                 "(profiles/)"},
        }
        if e2e:
            line["e2e"] = e2e
        if not args.no_cpu_baseline and world == 1:
            try:
                line["cpu_baseline"], _ = cpu_baseline(sc, args.workload)
            except Exception as exc:  # the checker failing must not hide the GPU number
                line["cpu_baseline"] = {"error": repr(exc)}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
