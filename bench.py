#!/usr/bin/env python
"""bench.py — headline benchmark of the hot path (BASELINE.json): Mpath-segments/s and spp/s at 1920x1080 on
hyperion_rect_lights (maxdepth 3, RR from depth 2, 17 quad lights), 1..8 B200.

One "step" = one pass of the hot path over one batch: SPP_PER_STEP full-frame sample passes (every tile, reference RNG /
frameNum schedule) rendered by the CUDA wavefront pipeline through the C ABI.

  python bench.py --gpus N --steps K --warmup W            our arm (N>1: launched with torchrun, one rank per GPU)
  python bench.py --impl reference ...                      the reference's own shader text compiled for the host (oracle/glsl_ref)

The JSON line carries `roofline` (dominant kernel = closest-hit traversal, algorithmic bytes from the instrumented oracle
on a bounded sample of the same workload), `cpu_baseline`, `e2e`, `clocks`, `gpu_launches`.
"""
from __future__ import annotations
import argparse, copy, ctypes as C, json, os, subprocess, sys, threading, time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import numpy as np

SCENE = "hyperion_rect_lights"
W, H = 1920, 1080
SPP_PER_STEP = 32
METRIC = "Mpath-segments/s"
# BASELINE.json configs -> (scene blob, width, height, maxdepth override).  Tile sizes / depth come from the .scene files.
WORKLOADS = {
    "hyperion_rect_lights": ("hyperion_rect_lights", 1920, 1080, None),     # configs[2] — the headline (default)
    "hyperion_sphere_light": ("hyperion_sphere_light", 1920, 1080, None),   # configs[2]
    "cornell_box_orig": ("cornell_box_orig", 512, 512, 4),                  # configs[0]
    "ibl_spheres": ("ibl_spheres", 1920, 1080, None),                       # configs[1] stand-in (teapot meshes + HDR are missing from the checkout)
    "volume_cube": ("volume_cube", 1920, 1080, None),                       # configs[3]
    "instancing": ("instancing", 3840, 2160, None),                         # configs[4]: 10 001 instances, depth 8, 4K
}


def load_workload(scene_name=SCENE, w=None, h=None):
    import glsl_pathtracer_b200  # noqa: F401
    from glsl_pathtracer_b200 import scene_io
    blob, ww, hh, depth = WORKLOADS[scene_name]
    global W, H
    W, H = (w or ww), (h or hh)
    sc = copy.deepcopy(scene_io.load_scene(blob))
    ro = sc.renderOptions
    ro.renderResolution = (W, H); ro.windowResolution = (W, H)     # tiles stay the .scene file's (hyperion: 256x144)
    if depth is not None:
        ro.maxDepth = depth
    return sc


def workload_config(sc, extra=None):
    ro = sc.renderOptions
    c = {"workload": f"{sc.name}.scene {ro.renderResolution[0]}x{ro.renderResolution[1]}, maxdepth {ro.maxDepth}, RR depth {ro.RRDepth}, "
                     f"{len(sc.lights)} analytic lights, tiles {ro.tileWidth}x{ro.tileHeight}, {SPP_PER_STEP} spp per step",
         "scene": sc.name, "resolution": list(ro.renderResolution), "spp_per_step": SPP_PER_STEP,
         "l2_policy": "inputs larger than L2: >1 GB of path state is streamed per step (126 MB L2); the 15 MB scene is L2-resident by nature"}
    if extra:
        c.update(extra)
    return c


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True); self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i].lower().startswith("active") for r in self.rows)]
        busy = sorted(sm)[len(sm) // 2:] if sm else []
        return {"sm_mhz": (sorted(busy)[len(busy) // 2] if busy else None), "sm_max_mhz": (max(mx) if mx else None), "reasons": reasons,
                "samples": len(sm)}


def algorithmic_bytes(sc, rect=None):
    """Per-ray algorithmic fetch bytes at the reference's granularity: 12 B/node visit + 48 B/internal step +
    60 B/triangle test + 64 B/TLAS leaf, counted by the instrumented oracle on a bounded sample (1 full-frame sample pass) of the
    same workload; culled traversal (conservative) and reference-faithful unculled."""
    from oracle import binding as ob
    out = {}
    for cull in (True, False):
        o = ob.Oracle(sc, cull=cull)
        acc = np.zeros((sc.renderOptions.renderResolution[1], sc.renderOptions.renderResolution[0], 4), np.float32)
        o.render(1, 1, accum=acc, rect=rect)   # rect=None: whole frame
        st = o.stats(); o.close()
        out["culled" if cull else "unculled"] = {
            "closest": (12 * st["nodeVisits"] + 48 * st["internalSteps"] + 60 * st["triTests"] + 64 * st["tlasLeaves"]) / max(1, st["closestRays"]),
            "any": (12 * st["anyNodeVisits"] + 48 * st["anyInternalSteps"] + 60 * st["anyTriTests"] + 64 * st["anyTlasLeaves"]) / max(1, st["anyRays"]),
            "closestRays": st["closestRays"], "anyRays": st["anyRays"]}
    return out


def reference_arm(sc):
    """The reference's own shader text compiled for the host (oracle/_ref/glsl_ref, built by __graft_entry__.build() where
    /root/reference exists; the prebuilt object travels to the GPU box).  None if no object for this option set is available."""
    try:
        from oracle.glsl_ref import binding as gb
        return gb.GlslRef(sc)
    except (FileNotFoundError, OSError, RuntimeError):
        return None


def cpu_baseline(sc, spp=None):
    """CPU baseline on a bounded sample (full-frame passes, all host threads; the pass count is chosen from the first pass so that each
    arm takes about 5-10 s: 8 passes on the headline workload, 1 on the 4K depth-8 scene).  kind "reference": the reference's own
    tile.glsl (+common/*.glsl) compiled by g++ -O2 (oracle/glsl_ref); the oracle port (hand-written restatement, bit-identical
    output, so identical path segments) is timed beside it and counts the segments."""
    from oracle import binding as ob
    o = ob.Oracle(sc)
    t0 = time.time()
    o.render(1, 1)
    dt1 = time.time() - t0
    if spp is None:
        spp = max(1, min(8, int(6.0 / max(dt1, 1e-3))))
    if spp > 1:
        o.render(2, spp - 1)
    dt_port = time.time() - t0
    st = o.stats(); o.close()
    cores = ob.lib().orc_num_threads()
    port = {"value": st["closestRays"] / dt_port / 1e6, "spp_per_s": spp / dt_port}
    g = reference_arm(sc)
    note = ("reference GLSL under Mesa llvmpipe is not runnable in this image (no GL/Mesa/Xvfb); instead the reference's own shader text is "
            "compiled for the host by g++ (oracle/glsl_ref) and run on all host threads")
    if g is None:
        return {"value": port["value"], "unit": METRIC, "spp_per_s": port["spp_per_s"], "cores": cores, "kind": "port",
                "sample": f"{spp} full-frame sample passes of the same {W}x{H} workload ({dt_port:.1f} s wall)",
                "note": note + "; no prebuilt reference-shader object for this option set: oracle port timed instead"}
    t0 = time.time()
    g.render(1, spp)
    dt = time.time() - t0
    return {"value": st["closestRays"] / dt / 1e6, "unit": METRIC, "spp_per_s": spp / dt, "cores": g.num_threads(), "kind": "reference",
            "sample": f"{spp} full-frame sample passes of the same {W}x{H} workload ({dt:.1f} s wall)",
            "oracle_port": {"value": port["value"], "spp_per_s": port["spp_per_s"], "wall_s": dt_port}, "note": note}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def measured_traffic(workload):
    """DRAM bytes of the dominant kernel per step from the committed ncu capture (profiles/), headline workload only."""
    p = os.path.join(ROOT, "profiles", "r01_trace_traffic.json")
    if workload == SCENE and os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["k_trace_dram_bytes_per_step"])
    return None


def run_reference(args):
    """Reference arm: the reference's own shader text (oracle/_ref/glsl_ref) on all host threads; the oracle port only if no
    reference-shader object exists for this option set.  One bounded sample (1 full-frame pass) per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    sc = load_workload(args.workload)
    from oracle import binding as ob
    o = ob.Oracle(sc)
    g = reference_arm(sc)
    kind = "reference" if g is not None else "port"
    arm = g if g is not None else o
    spp_step = 1                       # bounded sample per step: one full-frame pass
    for _ in range(min(args.warmup, 1)):
        arm.render(1, 1, rect=(0, 0, W, 64))
    o.stats(reset=True)
    dt = 0.0
    for k in range(args.steps):
        t0 = time.time()
        arm.render(1 + k, spp_step)
        dt += time.time() - t0
        if g is not None:
            o.render(1 + k, spp_step)   # untimed: counts the path segments of the same passes (bit-identical paths)
    st = o.stats()
    val = st["closestRays"] / dt / 1e6
    cores = g.num_threads() if g is not None else ob.lib().orc_num_threads()
    what = ("the reference's tile.glsl + common/*.glsl compiled by g++ -O2 (oracle/glsl_ref)" if g is not None
            else "oracle port (no reference-shader object for this option set)")
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": METRIC, "spp_per_s": args.steps * spp_step / dt, "n_gpus": 0,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": f"scene fixture built by the reference host code ({sc.name})",
            "config": workload_config(sc, {"reference_step": "1 full-frame sample pass per step (bounded sample of the same workload)",
                                           "reference_impl": what}),
            "cpu_baseline": {"value": val, "unit": METRIC, "cores": cores, "kind": kind,
                             "sample": f"{args.steps} x 1 full-frame {W}x{H} sample pass on {cores} host threads"},
            "e2e": {"value": val, "unit": METRIC, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


def run_ours(args):
    import torch
    import torch.distributed as dist
    from glsl_pathtracer_b200 import capi
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; libptb200 has no CPU path (use --impl reference for the CPU oracle)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    sc = load_workload(args.workload)
    ctx = capi.Context(sc, device=local, samples_per_wave=args.samples_per_wave)
    ctx.set_cull(args.cull)
    stream = torch.cuda.Stream(device=local)
    ctx.set_stream(stream.cuda_stream)
    ptr, nbytes = ctx.accum_device_ptr()

    class _Wrap:   # expose the accumulation buffer to torch for the NCCL reduce
        __cuda_array_interface__ = {"shape": (H, W, 4), "typestr": "<f4", "data": (ptr, False), "version": 3}
    accum_t = torch.as_tensor(_Wrap(), device=torch.device("cuda", local))

    def step(k, e2e):
        # sample sharding: rank r renders passes {first + r + i*world}; the per-GPU work is fixed (weak scaling)
        first = 1 + k * SPP_PER_STEP * world + rank
        if e2e:
            ctx.set_camera(sc.camera)                                  # per-step inputs from the host (uniforms)
            ctx.set_options(ctx.opts)
        ctx.render_samples(first, SPP_PER_STEP, world)
        if world > 1:
            with torch.cuda.stream(stream):
                dist.reduce(accum_t, dst=0, op=dist.ReduceOp.SUM)      # one NCCL reduce per readback over NVLink
        if e2e and rank == 0:
            return ctx.read_output(1.0 / float((k + 1) * SPP_PER_STEP * world))   # tonemap + D2H RGBA8 (GetOutputBuffer)
        return None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(e2e, steps):
        ctx.reset_accum(); ctx.reset_stats()
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.time()
        ev0.record(stream)
        for k in range(steps):
            step(k, e2e)
        ev1.record(stream)
        barrier()
        wall = time.time() - t0
        ms = ev0.elapsed_time(ev1) if not e2e else wall * 1e3          # e2e includes host-side copies/syncs: wall clock bracketed by syncs
        t = torch.tensor([ms], dtype=torch.float64, device=torch.device("cuda", local))
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        st = ctx.stats()
        segs = torch.tensor([st["pathSegments"], st["shadowRays"]], dtype=torch.float64, device=torch.device("cuda", local))
        if world > 1:
            dist.all_reduce(segs, op=dist.ReduceOp.SUM)
        return float(t.item()), float(segs[0].item()), float(segs[1].item()), st

    for k in range(max(args.warmup, 3)):
        step(k, False)
    barrier()
    ctx.set_profiling(True)
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    launches0 = ctx.stats()["kernelLaunches"]
    ms, segs, shadows, st = timed(False, args.steps)
    launches = ctx.stats()["kernelLaunches"] - launches0
    clk = clocks.stop() if rank == 0 else None
    ctx.set_profiling(False)
    trace_ms_last = st["lastTraceMs"]                 # closest-hit launches of the LAST step (events recorded in-stream)
    e2e_ms, e2e_segs, _, _ = timed(True, max(2, min(args.steps, 5)))
    e2e_steps = max(2, min(args.steps, 5))

    if rank == 0:
        total_spp = args.steps * SPP_PER_STEP * world
        value = segs / (ms * 1e-3) / 1e6
        peak, peak_src = peaks()
        # bounded sample for the oracle's byte count: the whole frame up to 1080p, else a centred 1920x1080 window of it
        ab_rect = None if W * H <= 1920 * 1080 else ((W - 1920) // 2, (H - 1080) // 2, (W - 1920) // 2 + 1920, (H - 1080) // 2 + 1080)
        ab = algorithmic_bytes(sc, ab_rect)
        rays_last_step = segs / args.steps / world          # rank-0 share of one step
        bytes_per_ray = ab["culled" if args.cull else "unculled"]["closest"]
        achieved = rays_last_step * bytes_per_ray / (trace_ms_last * 1e-3) / 1e9 if trace_ms_last > 0 else None
        line = {
            "metric": METRIC, "value": value, "unit": METRIC, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": f"scene fixture built by the reference host code ({sc.name}), reference RNG/frame schedule",
            "config": workload_config(sc, {"parallelism": f"sample-range sharding x{world}, NCCL reduce per readback", "cull_boxes": bool(args.cull)}),
            "spp_per_s": total_spp / (ms * 1e-3), "mshadow_rays_per_s": shadows / (ms * 1e-3) / 1e6, "mrays_per_s": (segs + shadows) / (ms * 1e-3) / 1e6,
            "e2e": {"value": e2e_segs / (e2e_ms * 1e-3) / 1e6, "unit": METRIC, "spp_per_s": e2e_steps * SPP_PER_STEP * world / (e2e_ms * 1e-3),
                    "h2d_bytes_per_step": C.sizeof(capi.PtbCamera) + C.sizeof(capi.PtbOptions), "d2h_bytes_per_step": W * H * 4,
                    "what": "Context.set_camera + set_options (host uniforms) + render_samples + tonemapped RGBA8 readback to host per step"},
            "gpu_launches": int(launches),
            "clocks": clk,
            "roofline": {"bound": "hbm", "kernel": "k_trace (closest-hit two-level BVH traversal)", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": (achieved / peak if achieved else None), "traffic": measured_traffic(args.workload), "peak_source": peak_src,
                         "algorithmic_bytes_per_step": rays_last_step * bytes_per_ray,
                         "traffic_note": "dram bytes of all k_trace launches of one step (ncu, profiles/r01_trace_traffic.json); achieved/algorithmic are per step as well",
                         "bytes_per_ray_culled": ab["culled"]["closest"], "bytes_per_ray_unculled": ab["unculled"]["closest"],
                         "bytes_per_shadow_ray": ab["culled"]["any"], "rays_per_step": rays_last_step, "trace_ms_per_step": trace_ms_last,
                         "mrays_per_s_kernel": (rays_last_step / (trace_ms_last * 1e-3) / 1e6 if trace_ms_last > 0 else None),
                         "note": "scene (15 MB) is L2-resident: achieved = algorithmic fetch bytes served mostly by L1/L2, see profiles/ for dram bytes and L2 hit rate"},
        }
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline(sc)
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    ctx.close()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cull", type=int, default=1)
    ap.add_argument("--samples-per-wave", type=int, default=0, help="0 = library default (~16 M paths in flight)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default=SCENE, choices=sorted(WORKLOADS))
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
