#!/usr/bin/env python
"""bench.py — tet-constraint projections/s of the XPBD hot path on B200.

One "step" = one frame of the hot path (timestep_t::step: detection, 10 substeps x 10
Gauss-Seidel iterations, commit, surface update) over one batch of synthetic input.

Headline (`value`, `e2e`, `roofline`): BASELINE.json configs[2] — the 1M-tet grid-tetrahedralised block
on an SDF sphere + floor, BVH broadphase and collision detection every substep ("config3"), in sustained
contact (`contacts` reports min / mean / max over the timed frames; the run fails when the mean is 0).
With N > 1 every rank runs its own config3 scene (independent scenes, no data-path collective, weak scaling),
so the N = 1 line of a scaling run equals the single-GPU bench.

Sub-records of the same JSON line (BASELINE.json configs[3], [4], the multi-GPU rows of SURVEY 8e):
  `decomposed` : ONE 8M-tet body (config5) cut over the N ranks; shared vertices travel through mailboxes in
                 peer memory, pushed over NVLink by the substep kernel itself (no collective); strong scaling.
                 At N = 1 the same body on one GPU — the reference point for the scaling efficiency.
  `ensemble`   : 4096 independent 2k-tet bodies (config4), 4096 / N per rank, no cross-GPU traffic; strong scaling.
  `fp64`       : the headline scene in the fp64 validation build (the reference is fp64 throughout).

`value`  : projections/s with state resident in HBM (CUDA events on the launch stream,
           per-step event pairs, L2 flushed between steps, max over ranks).
`e2e`    : the same metric through the C ABI with HOST buffers (sbsb200_step_host_f32: pinned
           host x,v -> device, step, device -> host x,v inside the timed region).
`--impl reference` times the reference's CPU algorithm (oracle/_ref when built, else the C
port in oracle/) on a bounded sample of the same workload, rank 0 only.
"""
import argparse
import hashlib
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
    ap.add_argument("--region-shape", type=int, default=None, help="0 pencils (default), 1 compact")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-sub", action="store_true", help="headline only: skip the decomposed / ensemble / fp64 sub-records")
    ap.add_argument("--decompose", action="store_true",
                    help="N > 1: cut the ONE body of the headline workload over the N GPUs (config2, config3, config5)")
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
                                          "--format=csv,noheader,nounits", "-lms", "10"],
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


def reference_world():
    """(kind, constructor) of the CPU checker: the reference's own sources when built, else the C port."""
    from oracle import oracle as O
    try:
        from oracle import ref as REF
        if REF.available():
            return "reference", REF.World
    except Exception:
        pass
    O.build()
    return "port", O.World


def cpu_baseline(sc, workload, rank=0, world=1):
    """Reference algorithm on the host: ONE substep (10 iterations) of the same scene, serial."""
    kind, mk = reference_world()
    W = mk()
    scene = make_scene(sc, workload, rank, world)
    scene.instantiate(W)
    dt = scene.dt / scene.substeps
    t0 = time.perf_counter()
    W.step(dt, 1, scene.iterations, scene.detect_every_substep)
    sec = time.perf_counter() - t0
    proj = scene.n_tets * scene.iterations
    return {"value": proj / sec, "unit": UNIT, "cores": 1, "kind": kind,
            "sample": "1 substep (detection + %d iterations, %d projections, %d contacts) of %s in %.2f s, serial "
                      "Gauss-Seidel as the reference (single-threaded), fp64"
                      % (scene.iterations, proj, len(W.contacts()[0]), scene.name, sec),
            "seconds": sec, "host_cores_available": os.cpu_count()}


def run_reference(args, rank, world):
    if rank != 0:
        return
    sc = importlib.import_module("soft-body-simulator_b200.scenes")
    kind, mk = reference_world()
    scene = make_scene(sc, args.workload, 0, world)
    W = mk()
    scene.instantiate(W)
    dt = scene.dt / scene.substeps
    for _ in range(args.warmup):
        W.step(dt, 1, 1, scene.detect_every_substep)          # warm caches; 1 iteration each
    t0 = time.perf_counter()
    for _ in range(args.steps):
        W.step(dt, 1, scene.iterations, scene.detect_every_substep)
    sec = time.perf_counter() - t0
    proj = scene.n_tets * scene.iterations * args.steps
    value = proj / sec
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sec / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": scene.name, "tets": scene.n_tets, "substeps": scene.substeps,
                       "iterations": scene.iterations,
                       "step": "bounded sample: each step is ONE substep (detection + %d iterations) of the frame"
                               % scene.iterations},
            "contacts_last_detection": len(W.contacts()[0]),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": kind,
                             "sample": "%d x 1 substep of %s, serial (the reference solver is single-threaded)"
                                       % (args.steps, scene.name)},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


class Bench:
    """One scene on this rank's GPU (or its share of a decomposed one), timed as the contract says."""

    def __init__(self, env, scene, precision, schedule=0, decomposed=False, region_shape=None):
        import torch
        import torch.distributed as dist
        self.env, self.scene, self.precision, self.decomposed = env, scene, precision, decomposed
        sbs = env["sbs"]
        self.sim = sbs.Simulation(env["local"], precision, stream=env["stream"].cuda_stream,
                                  schedule=sbs.SCHED_PERSISTENT if decomposed else schedule, region_shape=region_shape)
        if decomposed:
            # ONE body cut into world x regions; every rank runs its block, shared vertices travel through mailboxes
            # in peer memory (NVLink stores issued by the substep kernel; no collective)
            self.ids = scene.instantiate(self.sim, partition=(env["rank"], env["world"]))
            handles = [None] * env["world"]
            dist.all_gather_object(handles, self.sim.mailbox_handle())
            self.sim.connect_peers(handles)
            dist.barrier()
        else:
            self.ids = scene.instantiate(self.sim)
        self.stats0 = self.sim.stats()
        torch.cuda.synchronize()

    def barrier(self):
        import torch
        import torch.distributed as dist
        if self.env["world"] > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(self, x):
        import torch
        import torch.distributed as dist
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if self.env["world"] > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def resident(self, steps, warmup):
        """State resident in HBM: per-step CUDA-event pairs, L2 flushed between steps, max over ranks."""
        import torch
        sim, scene, stream = self.sim, self.scene, self.env["stream"]
        S, K = scene.substeps, scene.iterations
        for _ in range(warmup):
            sim.step(scene.dt, S, K, scene.detect_every_substep)
        self.barrier()
        st0 = sim.stats()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        contacts = []
        for a, b in ev:
            self.env["flush"].zero_()           # L2 flush between timed iterations (not timed)
            a.record(stream)
            sim.step(scene.dt, S, K, scene.detect_every_substep)
            b.record(stream)
            if not self.decomposed:             # (a rank of a decomposed body must not block between frames)
                contacts.append(sim.contact_count())
        self.barrier()
        if self.decomposed:
            contacts.append(sim.contact_count())
        st1 = sim.stats()
        ms = sum(a.elapsed_time(b) for a, b in ev)
        return {"ms_per_step": self.max_over_ranks(ms) / steps, "steps": steps, "warmup": warmup,
                "launches": st1["kernels_launched"] - st0["kernels_launched"],
                "kernel_ms": st1["kernel_ms"] - st0["kernel_ms"],
                "kernel_launches": st1["kernel_launches"] - st0["kernel_launches"], "ms_this_rank": ms,
                "contacts": {"min": int(min(contacts)), "mean": float(sum(contacts)) / len(contacts),
                             "max": int(max(contacts)), "samples": len(contacts),
                             "what": "contacts of the last detection of every timed frame"
                                     + (" (last frame only)" if self.decomposed else "")},
                "general_route_projections": st1["green_general_calls"]}

    def end_to_end(self, steps):
        """Through the C ABI with pinned HOST buffers: H2D of x, v, the frame, D2H of x, v per step (float host buffers).
        A rank of a decomposed body round-trips the vertices it owns (sbsb200_step_host_vertices_f32), every other
        workload the whole body (sbsb200_step_host_f32)."""
        import numpy as np
        import torch
        sim, scene = self.sim, self.scene
        S, K = scene.substeps, scene.iterations
        bodies = scene.tet_bodies()
        b0 = bodies[0]
        x0 = np.concatenate([scene.items[b].x for b in bodies]).astype(np.float32)
        v0 = np.concatenate([scene.items[b].v if scene.items[b].v is not None else np.zeros_like(scene.items[b].x)
                             for b in bodies]).astype(np.float32)
        all_bodies = len(bodies) > 1           # an ensemble crosses PCIe in one copy each way (SBSB200_ALL_BODIES)
        owned = None
        if self.decomposed:
            owned = np.ascontiguousarray(np.nonzero(sim.vertex_ranks(self.ids[b0]) == self.env["rank"])[0], np.uint32)
            x0, v0 = x0[owned], v0[owned]
        nV = x0.shape[0]
        x_in = torch.from_numpy(np.ascontiguousarray(x0)).pin_memory()
        v_in = torch.from_numpy(np.ascontiguousarray(v0)).pin_memory()
        x_out = torch.empty((nV, 3), dtype=torch.float32).pin_memory()
        v_out = torch.empty((nV, 3), dtype=torch.float32).pin_memory()
        xin, vin, xo, vo = x_in.numpy(), v_in.numpy(), x_out.numpy(), v_out.numpy()

        def frame(xin, vin, xo, vo):
            if owned is None:
                sim.step_host_f32(-1 if all_bodies else self.ids[b0], xin, vin, scene.dt, S, K, scene.detect_every_substep,
                                  xo, vo)
            else:
                sim.step_host_vertices_f32(self.ids[b0], owned, xin, vin, scene.dt, S, K, scene.detect_every_substep, xo, vo)

        for _ in range(2):
            frame(xin, vin, xo, vo)
            xin, xo = xo, xin
            vin, vo = vo, vin
        self.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            frame(xin, vin, xo, vo)
            xin, xo = xo, xin              # next frame continues from the host copy: the two pinned
            vin, vo = vo, vin              # buffer pairs swap roles, nothing is copied on the host
        self.barrier()
        wall = self.max_over_ranks(time.perf_counter() - t0)
        ids_bytes = 0 if owned is None else int(4 * nV)
        return {"seconds": wall, "steps": steps, "h2d_bytes_per_step": int(2 * nV * 12) + ids_bytes,
                "d2h_bytes_per_step": int(2 * nV * 12), "contacts_last_detection": sim.contact_count(),
                "timing": "wall clock around sbsb200_step_host%s_f32 (float host buffers), max over ranks"
                          % ("_vertices" if owned is not None else ""),
                "bodies_round_tripped": len(bodies), "vertices_round_tripped_this_rank": int(nV)}

    def roofline(self, res, peak, peaks_found, share=1):
        """Algorithmic bytes / CUDA-event time, per GPU.  share: ranks one body is cut over."""
        scene, st = self.scene, self.stats0
        S, K = scene.substeps, scene.iterations
        contacts = res["contacts"]["mean"]
        scale = 336.0 / 176.0 if self.precision == 64 else 1.0
        n_det = S if scene.detect_every_substep else 1
        bytes_step = (S * K * (BYTES_PER_PROJECTION * scene.n_tets + BYTES_PER_COLLISION * contacts)
                      + S * BYTES_PER_VERTEX_SUBSTEP * st["n_vertices"] + n_det * BYTES_PER_SURFACE_DETECT
                      * st["n_surface_vertices"]) * scale / share
        frame_gbs = bytes_step / (res["ms_per_step"] * 1e-3) / 1e9
        if res["kernel_launches"] > 0:
            # dominant kernel = the substep kernel of the resident schedule: one launch runs predict, K sweeps over
            # every tet and contact, and commit for one substep.  Timed live with CUDA events around every launch
            # of the timed region (sbsb200_stats.kernel_ms).
            bytes_launch = (K * (BYTES_PER_PROJECTION * scene.n_tets + BYTES_PER_COLLISION * contacts)
                            + BYTES_PER_VERTEX_SUBSTEP * st["n_vertices"]) * scale / share
            k_ms = res["kernel_ms"] / res["kernel_launches"]
            achieved = bytes_launch / (k_ms * 1e-3) / 1e9
            kernel = {"name": "k_substep_resident", "launches_timed": int(res["kernel_launches"]),
                      "avg_ms_per_launch": k_ms, "algorithmic_bytes_per_launch": int(bytes_launch),
                      "share_of_step": res["kernel_ms"] / res["ms_this_rank"]}
        else:
            # graph schedule: ~800 k_project_green launches per frame inside one CUDA graph; CUDA events cannot
            # bracket a node of a graph launch, so the frame as a whole is the timed unit
            achieved, kernel = frame_gbs, {"name": "whole frame (CUDA graph of per-colour kernels)"}
        return {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": None,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks_found else "6650 (of fallback)",
                "kernel": kernel,
                "frame": {"algorithmic_bytes_per_step": int(bytes_step), "achieved": frame_gbs,
                          "frac": frame_gbs / peak, "per": "GPU"},
                "frac_of_8TBs_nominal": achieved / 8000.0}

    def describe(self):
        scene, st = self.scene, self.stats0
        sched = {1: "graph", 2: "resident"}.get(st["schedule"], "?")
        return {"workload": scene.name, "tets": scene.n_tets, "vertices": st["n_vertices"], "schedule": sched,
                "regions": st["n_regions"], "colours": st["n_green_colours"], "shared_vertices": st["n_shared_vertices"],
                "pulls_per_sweep": st["pulls_per_sweep"], "colour_steps_without_exchange": st["quiet_colours"],
                "schedule_note": self.sim.schedule_note()}

    def close(self):
        self.sim.close()


def lib_sha(sbs):
    """Identity of the build: hash of the sources of libsbsb200.so (soft-body-simulator_b200/build.py)."""
    try:
        return importlib.import_module("soft-body-simulator_b200.build").source_id()
    except OSError:
        return None


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
        os.environ.setdefault("NCCL_DEBUG", "WARN")   # (NCCL otherwise prints its version on stdout, next to the JSON line)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    sbs = importlib.import_module("soft-body-simulator_b200")
    sc = importlib.import_module("soft-body-simulator_b200.scenes")
    # a dedicated non-blocking stream: the library captures the frame into a CUDA graph, which
    # the legacy default stream does not permit
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    env = {"sbs": sbs, "sc": sc, "rank": rank, "world": world, "local": local, "stream": stream,
           "flush": torch.empty(256 << 20, dtype=torch.uint8, device="cuda")}     # > 126 MB L2
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))

    # ---- headline ---------------------------------------------------------------------------
    decomposed = world > 1 and (args.workload == "config5" or (args.decompose and args.workload in ("config2", "config3")))
    sharded = args.workload in ("config3", "config4") and not decomposed       # every rank has its own scene(s)
    scene = make_scene(sc, args.workload, rank, world, decomposed)
    S, K = scene.substeps, scene.iterations
    head = Bench(env, scene, args.precision, args.schedule, decomposed, args.region_shape)
    sampler = ClockSampler(local)
    sampler.start()
    res = head.resident(args.steps, args.warmup)
    clocks = sampler.stop()
    total_tets = scene.n_tets * (world if sharded else 1)
    if args.workload == "config4":
        total_tets = 4096 * (scene.n_tets // max(1, len(scene.tet_bodies())))
    value = total_tets * S * K / (res["ms_per_step"] * 1e-3)
    e2e = None
    if not args.no_e2e:
        r = head.end_to_end(max(3, args.steps // 2))
        e2e = {"value": total_tets * S * K * r["steps"] / r["seconds"], "unit": UNIT,
               "h2d_bytes_per_step": r["h2d_bytes_per_step"], "d2h_bytes_per_step": r["d2h_bytes_per_step"],
               "steps": r["steps"], "timing": r["timing"], "bodies_round_tripped": r["bodies_round_tripped"],
               "contacts_last_detection": r["contacts_last_detection"]}
    roof = head.roofline(res, peak, bool(peaks), share=world if decomposed else 1)
    try:   # dram__bytes_read + dram__bytes_write per launch of the same kernel on the same workload, from an ncu
        # capture of THIS build of the library (profiles/); a capture of another build is not quoted
        tr = json.load(open(os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")))
        if tr["workload"] == scene.name and tr.get("lib_sha16") == lib_sha(sbs) and roof["kernel"]["name"] in tr["kernel"]:
            roof["traffic"] = tr["dram_bytes_read"] + tr["dram_bytes_write"]
            roof["traffic_source"] = "profiles/r02_ncu_traffic.json (ncu --set full of this build)"
    except Exception:
        pass
    roof["note"] = ("algorithmic bytes (SURVEY 8d: 176 B per projection, 64 B per contact projection, 112 B per "
                    "vertex and substep) / CUDA-event time; the working set fits the 126 MB L2 and the vertices "
                    "live in shared memory, so DRAM traffic is far below the algorithmic bytes (profiles/)")
    head_desc = head.describe()
    if args.workload == "config3" and world == 1 and res["contacts"]["mean"] <= 0:
        raise SystemExit("bench.py: the timed frames of config3 held no contact — the scene is not the one "
                         "BASELINE.json names (collision constraints every substep)")
    head.close()

    # ---- sub-records ----------------------------------------------------------------------------
    sub = {}
    if not args.no_sub and args.workload == "config3" and args.precision == 32:
        sub_steps, sub_warm = max(3, min(args.steps, 5)), max(3, min(args.warmup, 3))
        # (1) one 8M-tet body, decomposed over the N ranks (N = 1: the same body on one GPU)
        try:
            s5 = sc.config5()
            b = Bench(env, s5, 32, sbs.SCHED_PERSISTENT if world > 1 else 0, world > 1, args.region_shape)
            r5 = b.resident(sub_steps, sub_warm)
            rec = {"config": dict(b.describe(), parallelism=(
                       "one body cut over %d GPUs; shared vertices pushed into peer memory over NVLink by the substep "
                       "kernel, no collective" % world) if world > 1 else "one GPU (reference point of the strong scaling)",
                       tets_per_gpu=s5.n_tets // world),
                   "scaling": "strong", "n_gpus": world, "steps": r5["steps"], "warmup": r5["warmup"],
                   "ms_per_step": r5["ms_per_step"], "value": s5.n_tets * S * K / (r5["ms_per_step"] * 1e-3), "unit": UNIT,
                   "gpu_launches": int(r5["launches"]), "contacts": r5["contacts"],
                   "roofline": b.roofline(r5, peak, bool(peaks), share=world)}
            if not args.no_e2e:
                r = b.end_to_end(3)
                rec["e2e"] = {"value": s5.n_tets * S * K * r["steps"] / r["seconds"], "unit": UNIT,
                              "h2d_bytes_per_step": r["h2d_bytes_per_step"], "d2h_bytes_per_step": r["d2h_bytes_per_step"],
                              "steps": r["steps"], "timing": r["timing"],
                              "vertices_round_tripped_rank0": r["vertices_round_tripped_this_rank"],
                              "note": "every rank round-trips x, v (float) of the vertices it owns through pinned host "
                                      "memory each frame (bytes are rank 0's)"}
            sub["decomposed"] = rec
            b.close()
        except Exception as exc:                      # a failing sub-record must not hide the headline
            sub["decomposed"] = {"error": repr(exc)}
            if world > 1:
                raise
        # (2) 4096 independent bodies, 4096 / N per rank
        try:
            per = 4096 // world
            s4 = sc.config4(per, first=rank * per)
            b = Bench(env, s4, 32, 0, False, args.region_shape)
            r4 = b.resident(sub_steps, sub_warm)
            tets4 = 4096 * (s4.n_tets // per)
            sub["ensemble"] = {"config": dict(b.describe(), bodies_per_gpu=per, parallelism="bodies sharded over %d GPU(s), "
                                              "no cross-GPU traffic" % world),
                               "scaling": "strong", "n_gpus": world, "steps": r4["steps"], "warmup": r4["warmup"],
                               "ms_per_step": r4["ms_per_step"], "value": tets4 * S * K / (r4["ms_per_step"] * 1e-3),
                               "unit": UNIT, "gpu_launches": int(r4["launches"]), "contacts": r4["contacts"],
                               "roofline": b.roofline(r4, peak, bool(peaks))}
            if not args.no_e2e:
                r = b.end_to_end(3)
                sub["ensemble"]["e2e"] = {"value": tets4 * S * K * r["steps"] / r["seconds"], "unit": UNIT,
                                          "h2d_bytes_per_step": r["h2d_bytes_per_step"],
                                          "d2h_bytes_per_step": r["d2h_bytes_per_step"], "steps": r["steps"],
                                          "timing": r["timing"], "bodies_round_tripped_per_rank": r["bodies_round_tripped"],
                                          "note": "every rank round-trips x, v (float) of all its bodies in one copy each "
                                                  "way (SBSB200_ALL_BODIES; bytes are rank 0's)"}
            b.close()
        except Exception as exc:
            sub["ensemble"] = {"error": repr(exc)}
        # (3) the headline scene in the fp64 validation build (the reference computes in fp64)
        if world == 1:
            try:
                b = Bench(env, make_scene(sc, "config3", rank, world), 64, args.schedule, False, args.region_shape)
                r64 = b.resident(sub_steps, sub_warm)
                sub["fp64"] = {"config": b.describe(), "dtype": "f64", "steps": r64["steps"], "warmup": r64["warmup"],
                               "ms_per_step": r64["ms_per_step"],
                               "value": scene.n_tets * S * K / (r64["ms_per_step"] * 1e-3), "unit": UNIT,
                               "contacts": r64["contacts"], "roofline": b.roofline(r64, peak, bool(peaks)),
                               "note": "336 B per projection in fp64 (SURVEY 8d)"}
                b.close()
            except Exception as exc:
                sub["fp64"] = {"error": repr(exc)}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": res["ms_per_step"], "higher_is_better": True,
            "scaling": "strong" if args.workload in ("config4", "config5") or decomposed else "weak", "vs_baseline": None,
            "dtype": "f32" if args.precision == 32 else "f64", "data": "synthetic",
            "config": dict(head_desc, tets_per_gpu=scene.n_tets // (world if decomposed else 1),
                           substeps=S, iterations=K, dt=scene.dt,
                           detection=("every substep" if scene.detect_every_substep else "once per frame")
                           + (", BVH broadphase" if scene.broadphase else ""),
                           parallelism=("one body decomposed over %d GPUs, shared vertices pushed to peer memory by "
                                        "the substep kernel" % world) if decomposed else
                           "scenes sharded over %d GPU(s), no collective" % world,
                           l2="256 MiB write between timed steps (flush)"),
            "ms_per_frame": res["ms_per_step"],
            "contacts": res["contacts"], "contacts_last_detection": res["contacts"]["max"],
            "general_route_projections": res["general_route_projections"],
            "clocks": clocks, "gpu_launches": int(res["launches"]),
            "roofline": roof, "lib_sha16": lib_sha(sbs),
        }
        if e2e:
            line["e2e"] = e2e
        line.update(sub)
        if not args.no_cpu_baseline and world == 1:
            try:
                line["cpu_baseline"] = cpu_baseline(sc, args.workload)
            except Exception as exc:  # the checker failing must not hide the GPU number
                line["cpu_baseline"] = {"error": repr(exc)}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
