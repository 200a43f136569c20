#!/usr/bin/env python
"""bench.py — Mrays/s and frame ms of the SVO ray caster on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

A step is one rendered frame. N=1: BASELINE config 2 (depth-12 sphere fractal, 1920x1080, primary rays
+ Lambert shading, one B200). N>1 (one process per GPU under torchrun): every rank renders one frame
of a flythrough batch of the same scene (frame f -> GPU f mod N, north_star "by frames for flythrough
batches") and writes its pixels straight into GPU 0's batch buffer over NVLink: weak scaling.
Prints ONE JSON line on rank 0.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

NODE_BYTES = 40          # reference node (reaction/report/main.tex:46-51) — roofline unit, SURVEY §8d
PIXEL_BYTES = 4          # one RGBA8 store


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--depth", type=int, default=12)
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--scene", default="fractal", choices=["fractal", "iso"],
                    help="fractal = gen_spheres.py scene (configs 1,2,4); iso = synthetic large volume (configs 3,5)")
    ap.add_argument("--schedule", default="auto", choices=["auto", "tiles", "persistent", "queue"])
    ap.add_argument("--smem-nodes", type=int, default=-1)
    ap.add_argument("--stack", type=int, default=-1, choices=[-1, 0, 1, 2, 4])
    ap.add_argument("--secondary", action="store_true", help="BASELINE config 4: shadow + 4 AO rays")
    ap.add_argument("--partition", default="frames", choices=["frames", "tiles", "bands"],
                    help="N>1: one frame per rank (weak), or one frame split (strong) into interleaved "
                         "blocks of --band-rows rows (tiles) or contiguous row bands (bands)")
    ap.add_argument("--band-rows", type=int, default=32)
    ap.add_argument("--gather", default="p2p", choices=["p2p", "nccl"])
    ap.add_argument("--detail", type=float, default=0.0,
                    help="SVORenderer::SetDetailCoef LOD cut-off (0 = off, the CPU tracer's behaviour; the CUDA demo used 1.0)")
    ap.add_argument("--ssna", action="store_true",
                    help="SetSSNA: BlurZ x5 + normals from the z-buffer (demo/SVORenderer.cpp:126-147); one GPU")
    ap.add_argument("--flythrough", action="store_true",
                    help="tiles partition: move the camera every step (BASELINE config 5: 64-frame flythrough)")
    ap.add_argument("--cull", action="store_true",
                    help="ablation: octant culling on (fewer node fetches, measured slower: profiles/README.md)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-l2-flush", action="store_true")
    return ap.parse_args()


# camera: eye of cell/main.cpp:25; that file's view direction (-1,-1,-1.5) was written for scene.vox and
# sees none of the sphere fractal (0 hits), so the z component is flipped to look at the fractal.
BASE_POS = (0.5, 0.5, 0.3)
BASE_DIR = (-1.0, -1.0, 1.5)
UP = (0.0, 0.0, 1.0)
FOV = 70.0


ISO_POS = (0.2, 0.15, 0.45)       # above the terrain slab of the synthetic large volume, looking across it
ISO_DIR = (0.6, 0.7, -0.45)
SCENE = "fractal"


def build_scene(yv, a, threads):
    if a.scene == "iso":
        return yv.SVOData.IsoVolume(a.depth, seed=219, iso_level=200, threads=threads)      # gen_largevol.py:8-30
    return yv.SVOData.SphereFractal(a.depth, threads=threads)                               # gen_spheres.py:8-32


def _catmull_rom(pts, u):
    """Closed uniform Catmull-Rom spline through pts (n,3) at parameter u in [0, n)."""
    n = len(pts)
    i = int(np.floor(u)) % n
    t = u - np.floor(u)
    p0, p1, p2, p3 = pts[(i - 1) % n], pts[i], pts[(i + 1) % n], pts[(i + 2) % n]
    return 0.5 * ((2 * p1) + (-p0 + p2) * t + (2 * p0 - 5 * p1 + 4 * p2 - p3) * t * t + (-p0 + 3 * p1 - 3 * p2 + p3) * t ** 3)


_ISO_PATH = None


def camera_for(frame, n_frames=64):
    """Deterministic flythrough. Sphere fractal: frame 0 is the base camera, later frames orbit the eye a
    little. Iso volume (config 5): a seeded closed Catmull-Rom path inside the cube, above the terrain slab,
    looking along the tangent with a downward pitch; frame 0 is the fixed config-3 camera."""
    global _ISO_PATH
    if SCENE == "iso":
        if frame == 0:
            return ISO_POS, ISO_DIR
        if _ISO_PATH is None:
            rng = np.random.RandomState(219)
            ang = np.sort(rng.rand(8)) * 2 * np.pi
            rad = 0.22 + 0.12 * rng.rand(8)
            _ISO_PATH = np.stack([0.5 + rad * np.cos(ang), 0.5 + rad * np.sin(ang), 0.40 + 0.08 * rng.rand(8)], axis=1)
        u = 8.0 * (frame % n_frames) / n_frames
        pos = _catmull_rom(_ISO_PATH, u)
        tan = _catmull_rom(_ISO_PATH, u + 0.05) - pos
        tan /= max(1e-9, np.linalg.norm(tan))
        d = (tan[0], tan[1], -0.55)
        return tuple(float(np.float32(v)) for v in pos), tuple(float(np.float32(v)) for v in d)
    if frame == 0:
        return BASE_POS, BASE_DIR
    a = 0.35 * frame
    pos = (BASE_POS[0] + 0.05 * np.sin(a), BASE_POS[1] + 0.05 * (1 - np.cos(a)), BASE_POS[2] + 0.01 * frame)
    d = (BASE_DIR[0] + 0.2 * np.sin(0.5 * a), BASE_DIR[1] - 0.2 * np.sin(0.3 * a), BASE_DIR[2])
    return tuple(float(v) for v in pos), tuple(float(v) for v in d)


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        deadline = time.time() + 1.0
        while not self.rows and time.time() < deadline:      # nvidia-smi needs ~0.2 s to print its first row
            time.sleep(0.05)
        rows = [r for (ts, r) in self.rows if t0 - 0.05 <= ts <= t1 + 0.1 and len(r) >= 8] or \
               [r for (_, r) in self.rows if len(r) >= 8]
        if not rows:
            self.proc.terminate()
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = sorted(float(r[1]) for r in rows)
        reasons = []
        for i, name in ((4, "hw_slowdown"), (5, "hw_thermal_slowdown"), (6, "sw_thermal_slowdown"), (7, "sw_power_cap")):
            if any(r[i].lower().startswith("active") for r in rows):
                reasons.append(name)
        self.proc.terminate()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][2]), "reasons": reasons,
                "samples": len(rows), "power_w_max": max(float(r[3]) for r in rows)}


def workload_name(a):
    scene = "gen_spheres sphere-fractal SVO" if a.scene == "fractal" else "gen_largevol-style synthetic iso-volume SVO (seed 219, iso 200)"
    return ("%s depth %d, %dx%d primary rays + Lambert%s%s" %
            (scene, a.depth, a.width, a.height, " + shadow + 4 AO" if a.secondary else "",
             ", LOD detailCoef %g" % a.detail if a.detail > 0 else "") + (" + SSNA (z-buffer BlurZ x5)" if a.ssna else ""))


def oracle_frame(svo_nodes, root, a, frame, threads, want_visits=False):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import yvo
    pos, d = camera_for(frame)
    cam = yvo.camera(pos, d, UP, FOV, a.width, a.height, detail_coef=a.detail, ssna=a.ssna,
                     ssna_voxel_size=1.0 / (1 << a.depth))
    sec = None
    if a.secondary:
        sec = yvo.secondary(shadow=1, ao_samples=4, seed=1, light_pos=(0.6, 0.4, 1.2),
                            voxel_size=1.0 / (1 << a.depth), ao_max_t=0.05)
    t0 = time.perf_counter()
    r = yvo.render(svo_nodes, root, cam, sec=sec, threads=threads, want_visits=want_visits)
    return r, time.perf_counter() - t0


class ReferenceBuild:
    """oracle/_ref/libppu_renderer_ref.so: the reference's own CPU renderer (cell/ppu_renderer.cpp, compiled unmodified
    where /root/reference exists; the built library travels with the repo). Loads the scene through SVOData::Load and
    renders through ISVORenderer, as cell/main.cpp does. Primary rays + Lambert only (that is all it has)."""

    def __init__(self, svo):
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import yvref
        self.yvref = yvref
        self.scene = None
        if not (os.path.exists(yvref.PPU_SO) or os.path.exists("/root/reference/cell/ppu_renderer.cpp")):
            return
        if not yvref.available():
            return
        import tempfile
        d = "/dev/shm" if os.path.isdir("/dev/shm") else None
        fd, path = tempfile.mkstemp(suffix=".vox", dir=d)
        os.close(fd)
        try:
            svo.Save(path)
            self.scene = yvref.Scene(path)
        finally:
            os.remove(path)

    def ok(self, a):
        return self.scene is not None and not a.secondary and not a.ssna and a.detail == 0

    def frame(self, a, frame, threaded=True):
        pos, d = camera_for(frame)
        t0 = time.perf_counter()
        img = self.yvref.ppu_frame(self.scene, pos, d, UP, FOV, a.width, a.height, threaded=threaded)
        dt = time.perf_counter() - t0
        return img.view("uint8").reshape(a.height, a.width, 4), dt


def run_reference(a, rank):
    """--impl reference: the reference's CPU tracer on the host cores, same config. The headline value is the oracle port
    (kind = "port") because it can use every host core, which makes it the stronger baseline; the reference's own
    renderer compiled from its sources (oracle/_ref, TreadedRenderer: 4 threads by construction) is timed beside it
    as `reference_build` whenever the library is present, and its frame must equal the port's."""
    if rank != 0:
        return
    import yoxel_voxel_b200 as yv
    cores = os.cpu_count() or 1
    svo = build_scene(yv, a, cores)
    nodes, root = svo.nodes(), svo.GetRoot()
    for _ in range(a.warmup):
        oracle_frame(nodes, root, a, 0, cores)
    times, rays = [], 0
    for _ in range(a.steps):
        r, dt = oracle_frame(nodes, root, a, 0, cores)
        times.append(dt)
        rays += r["stats"]["rays"]
    total = sum(times)
    val = rays / total / 1e6
    ref_build = None
    rb = ReferenceBuild(svo)
    if rb.ok(a):
        # the reference's own renderer: TreadedRenderer's thread count is a constant (ThreadNum = 4, ppu_renderer.cpp:129)
        rb.frame(a, 0)
        ts = [rb.frame(a, 0)[1] for _ in range(max(1, min(a.steps, 5)))]
        img, _ = rb.frame(a, 0)
        ref_build = {"value": a.width * a.height / (sum(ts) / len(ts)) / 1e6, "unit": "Mrays/s", "cores": 4, "kind": "reference",
                     "what": "cell/ppu_renderer.cpp TreadedRenderer compiled unmodified (oracle/_ref), whole frame, 4 threads (its constant)",
                     "frame_identical_to_port": bool((img == r["rgba"]).all())}
    line = {
        "impl": "reference", "metric": "Mrays/s", "value": val, "unit": "Mrays/s", "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": 1e3 * total / a.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(a), "camera": {"pos": camera_for(0)[0], "dir": camera_for(0)[1], "fov": FOV}},
        "cpu_baseline": {"value": val, "unit": "Mrays/s", "cores": cores, "kind": "port",
                         "sample": "whole %dx%d frame per step, %d row strips (TreadedRenderer split)" % (a.width, a.height, cores)},
        "e2e": {"value": val, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "reference_build": ref_build,
    }
    print(json.dumps(line))


def main():
    global SCENE
    a = parse()
    SCENE = a.scene
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if a.impl == "reference":
        return run_reference(a, rank)

    import torch
    import torch.distributed as dist
    import yoxel_voxel_b200 as yv
    from yoxel_voxel_b200 import multigpu

    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")      # stdout carries the one JSON line and nothing else
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    cores = os.cpu_count() or 1

    # ---- scene: replicated on every GPU ----------------------------------------------------------
    # N>1: rank 0 builds with all host cores and parks the .vox in /dev/shm, the other ranks load it
    # (SVOData::Load) — the build is host work shared by the box, the replica per GPU is not.
    t0 = time.time()
    if world > 1:
        shm = "/dev/shm/yv_%s_d%d_%d.vox" % (a.scene, a.depth, os.getuid())
        if rank == 0:
            svo = build_scene(yv, a, cores)
            svo.Save(shm)
        dist.barrier()
        if rank != 0:
            svo = yv.SVOData().Load(shm)
        dist.barrier()
        if rank == 0:
            os.unlink(shm)
    else:
        svo = build_scene(yv, a, cores)
    build_s = time.time() - t0
    dev_bytes = svo.Upload(local)
    n_rec, n_leaf = (x.shape[0] for x in svo.packed())

    r = yv.SVORenderer(local)
    schedule = a.schedule if a.schedule != "auto" else "tiles"
    r.SetOption("schedule", {"tiles": 0, "persistent": 1, "queue": 2}[schedule])
    if a.smem_nodes >= 0:
        r.SetOption("smem_nodes", a.smem_nodes)
    if a.stack >= 0:
        r.SetOption("stack", a.stack)
    r.SetScene(svo)
    r.SetResolution(a.width, a.height)
    r.SetViewUp(UP); r.SetFOV(FOV)
    r.SetDetailCoef(a.detail)
    if a.ssna:
        r.SetSSNA(True, 1.0 / (1 << a.depth))
    if a.secondary:
        r.SetSecondary(shadow=1, ao_samples=4, seed=1, light_pos=(0.6, 0.4, 1.2),
                       voxel_size=1.0 / (1 << a.depth), ao_max_t=0.05)
    stream = torch.cuda.Stream(device=dev)       # a real (non-NULL) stream: the renderer, the L2 flush and the
    torch.cuda.set_stream(stream)                # timing events all run on it
    r.SetStream(stream.cuda_stream)

    # ---- partition ----------------------------------------------------------------------------
    frame_bytes = a.width * a.height * 4
    tiles_mode = world > 1 and a.partition in ("tiles", "bands")
    my_rows = np.arange(a.height)
    if tiles_mode:
        if a.partition == "bands":
            y0, y1 = multigpu.row_band(rank, world, a.height)
            r.SetRows(y0, y1)
            my_rows = np.arange(y0, y1)
        else:
            r.SetInterleave(a.band_rows, world, rank)
            my_rows = multigpu.interleaved_rows(rank, world, a.height, a.band_rows)
        y0, y1 = 0, a.height
        my_frame = 0
        target_bytes = frame_bytes
    else:
        y0, y1 = 0, a.height
        # weak scaling: every GPU renders one whole frame of the same workload (the base camera), or, with
        # --flythrough, frame step*N+rank of the camera path
        my_frame, my_rays_px = (rank if a.flythrough else 0), a.width * a.height
        target_bytes = frame_bytes * world
    pos, d = camera_for(my_frame)
    r.SetViewPos(pos); r.SetViewDir(d)

    gather = a.gather if world > 1 else "none"
    target_ptr, target_obj, local_fb, gather_list = None, None, None, None
    if world > 1 and gather == "p2p":
        try:
            target_ptr, target_obj = multigpu.open_gather_target(dist, rank, world, local, target_bytes)
        except yv.YVError as e:                       # no IPC on this box: fall back to the NCCL baseline
            gather = "nccl"
            if rank == 0:
                print("p2p gather unavailable (%s); using nccl" % e, file=sys.stderr)
        flag = torch.tensor([1 if gather == "p2p" else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 0:
            gather = "nccl"
    if world > 1 and gather == "nccl":
        local_fb = torch.zeros(a.height, a.width, 4, dtype=torch.uint8, device=dev)
        if rank == 0:
            gather_list = [torch.zeros_like(local_fb) for _ in range(world)]
    if world == 1:
        local_fb = torch.zeros(a.height, a.width, 4, dtype=torch.uint8, device=dev)

    def dst_ptr():
        if gather == "p2p":
            return target_ptr + (0 if tiles_mode else rank * frame_bytes)
        return local_fb.data_ptr()

    def render_step(step=None):
        if a.flythrough and step is not None:
            fpos, fdir = camera_for(frame_of(step))
            r.SetViewPos(fpos); r.SetViewDir(fdir)
        r.Render(dst_ptr(), sync=False)
        if gather == "nccl":
            dist.gather(local_fb, gather_list, dst=0)

    # ---- V-bar for the roofline: the kernel's own node-visit counters on this workload -----------
    # (tests/test_gpu_parity.py::test_counters_equal_oracle_visits pins them to the oracle's count of
    # the node fetch at cell/ppu_renderer.cpp:23)
    def frame_of(step):
        return step if (tiles_mode or world == 1) else step * world + rank

    r.EnableCounters(True)
    r.SetOption("cull", 0)                                    # V-bar is the REFERENCE traversal's node-fetch count
    probe = torch.zeros(a.height, a.width, 4, dtype=torch.uint8, device=dev)
    vis_sum = pop_sum = hit_px = 0
    probe_frames = range(a.steps) if a.flythrough else [None]
    for st in probe_frames:                                   # untimed pass: per-frame node-visit counters
        if st is not None:
            fpos, fdir = camera_for(frame_of(st))
            r.SetViewPos(fpos); r.SetViewDir(fdir)
        r.Render(probe.data_ptr(), sync=True)
        visits, pops = r.GetCounters()
        vis_sum += int(visits[my_rows].sum())
        pop_sum += int(pops[my_rows].sum())
        hit_px += int((probe[torch.as_tensor(my_rows, device=dev), :, 3] == 255).sum().item())
    # what the timed kernel actually fetches for the same frame(s) (differs only with --cull)
    r.SetOption("cull", 1 if a.cull else 0)
    kern_vis = kern_pop = 0
    for st in probe_frames:
        if st is not None:
            fpos, fdir = camera_for(frame_of(st))
            r.SetViewPos(fpos); r.SetViewDir(fdir)
        r.Render(probe.data_ptr(), sync=True)
        visits, pops = r.GetCounters()
        kern_vis += int(visits[my_rows].sum()); kern_pop += int(pops[my_rows].sum())
    r.EnableCounters(False)
    n_probe = len(probe_frames)
    kern_vis //= n_probe; kern_pop //= n_probe
    my_px = len(my_rows) * a.width
    my_rays = my_px + (5 * hit_px // n_probe if a.secondary else 0)      # shadow + 4 AO per hit pixel
    vis_sum //= n_probe; pop_sum //= n_probe                  # per-step averages
    hit_frac = hit_px / float(n_probe * max(1, my_px))
    del probe

    flush = None if a.no_l2_flush else torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    sampler = ClockSampler(local) if rank == 0 else None
    warm = max(a.warmup, 3)
    for _ in range(warm):
        if flush is not None:
            flush.zero_()
        render_step()
    sync_all()

    # ---- timed region: K frames, CUDA events on the launching stream ------------------------------
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
    sync_all()
    wall0 = time.time()
    for i in range(a.steps):
        if flush is not None:
            flush.zero_()                           # evict the node pool from L2 between frames (not timed)
        evs[i][0].record(stream)
        render_step(i)
        evs[i][1].record(stream)
        if world > 1:
            torch.cuda.synchronize()
            dist.barrier()                          # the batch on GPU 0 is complete once every rank has stored
    sync_all()
    wall1 = time.time()
    launches_per_step = r.LastFrameLaunches()       # 1 (trace); 8 with SSNA (z, 5 x BlurZ, ShadeSimple)
    clocks = sampler.stop(wall0, wall1) if sampler else None
    step_ms = [e0.elapsed_time(e1) for e0, e1 in evs]
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    sums = torch.tensor([float(my_rays), float(my_px), float(vis_sum), float(pop_sum), float(kern_vis), float(kern_pop)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)          # max over ranks
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)              # units all ranks processed
    total_s = float(total_ms.item()) / 1e3
    rays_step, px_step, vis_step, pop_step, kvis_step, kpop_step = (float(v) for v in sums.tolist())
    value = rays_step * a.steps / total_s / 1e6

    # ---- e2e: the public host API with host buffers (camera in, RGBA8 frame out) -------------------
    if world == 1:
        r.SetStream(0)
        e2e_t = []
        for i in range(warm + a.steps):
            if flush is not None:
                flush.zero_()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            r.SetViewPos(pos); r.SetViewDir(d); r.SetViewUp(UP); r.SetFOV(FOV)      # host camera -> kernel params
            img = r.RenderFrame()                                                   # launch + D2H (pinned) + sync
            e2e_t.append(time.perf_counter() - t0)
        e2e_s = sum(e2e_t[warm:])
        e2e = {"value": rays_step * a.steps / e2e_s / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": 40,
               "d2h_bytes_per_step": frame_bytes, "ms_per_step": 1e3 * e2e_s / a.steps,
               "api": "yv_set_view_* + yv_render_frame: host camera in, pinned host RGBA8 frame out",
               "checksum": int(img[::16, ::16].astype(np.uint64).sum())}
        r.SetStream(stream.cuda_stream)
    else:
        # One host frame (tiles) / one host batch of N frames (frames) in shared memory, registered with every GPU:
        # each rank stores its pixels straight into it over its own PCIe link (yv_host_register +
        # yv_render_frame_device), host camera in, host pixels out. Timed per step on the host clock around the
        # synchronous call, L2 flushed and ranks aligned by a barrier outside the timed region, max over ranks.
        e2e = None
        try:
            shared = multigpu.SharedHostFrame(dist, rank, world, local, target_bytes, os.environ.get("MASTER_PORT", "0"))
        except yv.YVError as ex:
            shared = None
            if rank == 0:
                print("shared host frame unavailable (%s)" % ex, file=sys.stderr)
        okf = torch.tensor([1 if shared is not None else 0], device=dev)
        dist.all_reduce(okf, op=dist.ReduceOp.MIN)
        if int(okf.item()) == 1:
            h_dst = shared.ptr + (0 if tiles_mode else rank * frame_bytes)
            e2e_t = []
            for i in range(warm + a.steps):
                if flush is not None:
                    flush.zero_()
                torch.cuda.synchronize()
                dist.barrier()
                t0 = time.perf_counter()
                fpos, fdir = camera_for(frame_of(i - warm)) if (a.flythrough and i >= warm) else (pos, d)
                r.SetViewPos(fpos); r.SetViewDir(fdir); r.SetViewUp(UP); r.SetFOV(FOV)
                r.Render(h_dst, sync=True)
                e2e_t.append(time.perf_counter() - t0)
            e2e_s = torch.tensor([sum(e2e_t[warm:])], dtype=torch.float64, device=dev)
            dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
            dist.barrier()
            same = None
            if rank == 0 and gather == "p2p" and not a.flythrough:
                # the frame the GPUs stored into host memory == the frame they stored into GPU 0 over NVLink
                ref = np.empty(target_bytes, np.uint8)
                yv.lib().yv_copy_to_host(local, ctypes.c_void_p(ref.ctypes.data), ctypes.c_void_p(target_ptr), target_bytes)
                same = bool((ref == shared.array).all())
            e2e = {"value": rays_step * a.steps / float(e2e_s.item()) / 1e6, "unit": "Mrays/s",
                   "h2d_bytes_per_step": 40 * world, "d2h_bytes_per_step": target_bytes,
                   "ms_per_step": 1e3 * float(e2e_s.item()) / a.steps,
                   "api": "yv_set_view_* + yv_render_frame_device into one pinned host %s shared by the ranks (yv_host_register): "
                          "every GPU stores its pixels over its own PCIe link" % ("frame" if tiles_mode else "batch of %d frames" % world),
                   "identical_to_nvlink_gathered": same}
        if shared is not None:
            dist.barrier()
            shared.close()
        if e2e is None:
            # fallback: the timed region already ends with every pixel resident on GPU 0; add rank 0's D2H of the batch
            d2h_s = 0.0
            if rank == 0 and gather == "p2p":
                pinned = torch.empty(target_bytes, dtype=torch.uint8, pin_memory=True)
                t0 = time.perf_counter()
                yv.lib().yv_copy_to_host(local, ctypes.c_void_p(pinned.data_ptr()), ctypes.c_void_p(target_ptr), target_bytes)
                d2h_s = time.perf_counter() - t0
            elif rank == 0:
                t0 = time.perf_counter()
                torch.stack(gather_list).cpu()
                d2h_s = time.perf_counter() - t0
            e2e = {"value": rays_step * a.steps / (total_s + d2h_s * a.steps) / 1e6, "unit": "Mrays/s",
                   "h2d_bytes_per_step": 40 * world, "d2h_bytes_per_step": target_bytes,
                   "api": "yv_render_frame_device into GPU 0's buffer (%s) + D2H of the gathered pixels on rank 0" % gather}

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- CPU baseline + parity check (rank 0, N=1): the oracle on the host cores --------------------
    cpu_baseline, parity = None, None
    if world == 1 and not a.no_cpu_baseline:
        nodes, root = svo.nodes(), svo.GetRoot()
        o, _ = oracle_frame(nodes, root, a, 0, cores, want_visits=True)        # warm run, parity, V-bar cross-check
        o2, dt = oracle_frame(nodes, root, a, 0, cores)
        cpu_baseline = {"value": o2["stats"]["rays"] / dt / 1e6, "unit": "Mrays/s", "cores": cores, "kind": "port",
                        "sample": "one whole %dx%d frame, %d row strips (%.2f s)" % (a.width, a.height, cores, dt),
                        "ms_per_frame": 1e3 * dt}
        # the CPU tracer at 1 thread, at the reference's constant of 4 (ppu_renderer.cpp:129) and on every core
        sweep = {}
        for t in sorted({1, 4, cores}):
            if t == cores:
                sweep[str(t)] = cpu_baseline["value"]
            else:
                ot, dtt = oracle_frame(nodes, root, a, 0, t)
                sweep[str(t)] = ot["stats"]["rays"] / dtt / 1e6
        cpu_baseline["mrays_by_threads"] = sweep
        gpu_img = r.RenderFrame()
        parity = {"rgba_identical_to_oracle": bool((gpu_img == o["rgba"]).all()),
                  "rays_identical": bool(o["stats"]["rays"] == int(rays_step)),
                  "node_visits_identical": bool(o["stats"]["node_visits"] == int(vis_step))}
        rb = ReferenceBuild(svo)
        if rb.ok(a):
            # the frame of the reference's own renderer (cell/ppu_renderer.cpp compiled unmodified, oracle/_ref)
            ref_img, ref_dt = rb.frame(a, 0)
            parity["rgba_identical_to_reference_build"] = bool((gpu_img == ref_img).all())
            cpu_baseline["reference_build"] = {"value": a.width * a.height / ref_dt / 1e6, "unit": "Mrays/s", "cores": 4,
                                               "kind": "reference", "what": "TreadedRenderer, one whole frame (%.2f s)" % ref_dt}

    # ---- roofline: algorithmic bytes / measured kernel time vs the measured HBM peak ----------------
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    alg_bytes_step = vis_step * NODE_BYTES + px_step * PIXEL_BYTES        # all GPUs
    kernel_s = total_s / a.steps
    achieved = alg_bytes_step / world / kernel_s / 1e9                    # per GPU (per launch)
    traffic, l2_traffic = None, None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    # the committed ncu capture is of BASELINE config 2; other workloads have no capture and report null
    is_cfg2 = (a.scene == "fractal" and a.depth == 12 and a.width == 1920 and a.height == 1080 and not a.secondary
               and not a.ssna and a.detail == 0)
    if is_cfg2 and os.path.exists(tp):
        try:
            tj = json.load(open(tp))
            traffic = tj.get("dram_bytes_per_launch")
            l2_traffic = tj.get("l2_bytes_per_launch")
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "l2_traffic": l2_traffic, "peak_source": peak_src, "kernel": "yv::render_frame<%s>" % schedule,
                "algorithmic_bytes_per_launch": alg_bytes_step / world,
                "node_visits_per_ray": vis_step / rays_step, "pop_refetches_per_ray": pop_step / rays_step,
                "kernel_node_fetches_per_ray": kvis_step / rays_step, "kernel_pop_refetches_per_ray": kpop_step / rays_step,
                "octant_culling": bool(a.cull),
                "kernel_ms": 1e3 * kernel_s}

    line = {
        "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": a.steps, "warmup": warm,
        "ms_per_step": 1e3 * total_s / a.steps, "higher_is_better": True,
        "scaling": "strong" if tiles_mode else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(a), "camera": {"pos": camera_for(0)[0], "dir": camera_for(0)[1], "fov": FOV},
                   "schedule": schedule, "smem_nodes": r.GetOption("smem_nodes"), "stack": r.GetOption("stack"),
                   "l2": "flushed between frames (256 MiB write, untimed)" if flush is not None
                         else "not flushed; node pool %d MB > L2" % (dev_bytes >> 20),
                   "nodes": svo.nodecount, "packed_bytes": dev_bytes, "scene_build_s": round(build_s, 2),
                   "partition": (("interleaved %d-row blocks of one frame" % a.band_rows if a.partition == "tiles" else "contiguous row bands of one frame")
                                 if tiles_mode else "one frame per GPU") if world > 1 else "single GPU",
                   "gather": gather, "flythrough": bool(a.flythrough), "hit_fraction": round(hit_frac, 4)},
        "frame_ms": 1e3 * total_s / a.steps, "rays_per_step": rays_step,
        "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "clocks": clocks,
        "gpu_launches": a.steps * world * launches_per_step, "e2e_gpu_launches_per_step": r.LastFrameLaunches(), "parity": parity,
    }
    print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
