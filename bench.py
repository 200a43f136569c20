#!/usr/bin/env python
"""bench.py — Mrays/s and frame ms of the SVO ray caster on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

A step is one rendered frame.

N=1: BASELINE config 2 (depth-12 sphere fractal, 1920x1080, primary rays + Lambert shading, one B200) is the headline
line; the other single-GPU configurations of BASELINE.json (1: 512x512 depth 10; 4: config 2 + shadow + 4 AO rays;
3: 3840x2160 over the deep iso volume) are measured after it, each to the same bar (device time, roofline fraction,
end-to-end time through yv_render_frame, full-frame parity against the CPU oracle), and ride in `configs`.

N>1 (one process per GPU under torchrun), headline: a flythrough batch of the config-2 scene, frame f on GPU f mod N
(north_star "by frames for flythrough batches"), every GPU storing its pixels straight into GPU 0's batch buffer over
NVLink: weak scaling. The same batch is then rendered by GPU 0 alone (`batch_1gpu`), and one gathered frame is compared
with GPU 0's own render of it and with an oracle band (`parity`). After that rank 0 runs BASELINE config 5 through ONE
renderer handle that drives all N GPUs (yv_renderer_create_group — the SPURenderer arrangement,
cell/spu_renderer.cpp:65-90): 7680x4320, 64-frame Catmull-Rom flythrough of the deep iso volume, each frame cut into
interleaved 32-row blocks, with the 1-GPU time of the same frames measured in the same run: `strong_8k`.

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
T_START = time.time()


def parse(argv=None):
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
                    help="move the camera every step (default for N>1 with the frames partition; config 5 with tiles)")
    ap.add_argument("--same-frame", action="store_true",
                    help="N>1, frames partition: every rank renders the base camera every step (round-1 behaviour)")
    ap.add_argument("--cull", action="store_true",
                    help="ablation: octant culling on (fewer node fetches, measured slower: profiles/README.md)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-l2-flush", action="store_true")
    ap.add_argument("--no-extras", action="store_true",
                    help="N=1: skip BASELINE configs 1, 4 and 3 after the headline (config 2) line")
    ap.add_argument("--extras-depth", type=int, default=0,
                    help="depth of the iso volume of config 3 in `configs` (0 = 14 if the host has the cores and memory, else 13)")
    ap.add_argument("--no-strong", action="store_true", help="N>1: skip the strong_8k record (BASELINE config 5)")
    ap.add_argument("--strong-depth", type=int, default=0, help="iso volume depth of strong_8k (0 = auto: 14, else 13)")
    ap.add_argument("--strong-frames", type=int, default=64)
    ap.add_argument("--strong-only", action="store_true",
                    help="single process: run only the strong_8k record over --gpus GPUs through one renderer handle")
    ap.add_argument("--strong-width", type=int, default=7680)
    ap.add_argument("--strong-height", type=int, default=4320)
    ap.add_argument("--budget-s", type=float, default=420.0,
                    help="wall-clock budget of the optional records (extras, strong_8k): what does not fit is skipped and says so")
    return ap.parse_args(argv)


# camera: eye of cell/main.cpp:25; that file's view direction (-1,-1,-1.5) was written for scene.vox and
# sees none of the sphere fractal (0 hits), so the z component is flipped to look at the fractal.
BASE_POS = (0.5, 0.5, 0.3)
BASE_DIR = (-1.0, -1.0, 1.5)
UP = (0.0, 0.0, 1.0)
FOV = 70.0

ISO_POS = (0.2, 0.15, 0.45)       # above the terrain slab of the synthetic large volume, looking across it
ISO_DIR = (0.6, 0.7, -0.45)
SEC_ARGS = dict(shadow=1, ao_samples=4, seed=1, light_pos=(0.6, 0.4, 1.2), ao_max_t=0.05)


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


def camera_for(frame, scene="fractal", n_frames=64):
    """Deterministic flythrough. Sphere fractal: frame 0 is the base camera, later frames orbit the eye a
    little. Iso volume (config 5): a seeded closed Catmull-Rom path inside the cube, above the terrain slab,
    looking along the tangent with a downward pitch; frame 0 is the fixed config-3 camera."""
    global _ISO_PATH
    if scene == "iso":
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
    pos = (BASE_POS[0] + 0.05 * np.sin(a), BASE_POS[1] + 0.05 * (1 - np.cos(a)), BASE_POS[2] + 0.01 * (frame % 64))
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
                                          "--format=csv,noheader,nounits", "-lms", "100"],
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


def config_of(a):
    """The `config` object: the same keys and values in both arms (everything arm-specific goes under `setup`)."""
    pos, d = camera_for(0, a.scene)
    return {"workload": workload_name(a), "camera": {"pos": list(pos), "dir": list(d), "up": list(UP), "fov": FOV}}


def hbm_peak():
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        return float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def _yvo():
    if os.path.join(ROOT, "tests") not in sys.path:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
    import yvo
    return yvo


def oracle_frame(svo_nodes, root, a, frame, threads, want_visits=False, rows=None):
    yvo = _yvo()
    pos, d = camera_for(frame, a.scene)
    cam = yvo.camera(pos, d, UP, FOV, a.width, a.height, detail_coef=a.detail, ssna=a.ssna,
                     ssna_voxel_size=1.0 / (1 << a.depth))
    sec = None
    if a.secondary:
        sec = yvo.secondary(voxel_size=1.0 / (1 << a.depth), **SEC_ARGS)
    t0 = time.perf_counter()
    r = yvo.render(svo_nodes, root, cam, sec=sec, threads=threads, want_visits=want_visits, rows=rows)
    return r, time.perf_counter() - t0


class ReferenceBuild:
    """oracle/_ref/libppu_renderer_ref.so: the reference's own CPU renderer (cell/ppu_renderer.cpp, compiled unmodified
    where /root/reference exists; the built library travels with the repo). Loads the scene through SVOData::Load and
    renders through ISVORenderer, as cell/main.cpp does. Primary rays + Lambert only (that is all it has)."""

    def __init__(self, vox_path=None, svo=None):
        if os.path.join(ROOT, "tests") not in sys.path:
            sys.path.insert(0, os.path.join(ROOT, "tests"))
        import yvref
        self.yvref = yvref
        self.scene = None
        if not (os.path.exists(yvref.PPU_SO) or os.path.exists("/root/reference/cell/ppu_renderer.cpp")):
            return
        if not yvref.available():
            return
        if vox_path is not None:
            self.scene = yvref.Scene(vox_path)
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
        pos, d = camera_for(frame, a.scene)
        t0 = time.perf_counter()
        img = self.yvref.ppu_frame(self.scene, pos, d, UP, FOV, a.width, a.height, threaded=threaded)
        dt = time.perf_counter() - t0
        return img.view("uint8").reshape(a.height, a.width, 4), dt


def scene_vox_for_reference(a, cores):
    """The scene file of the reference arm, written once per box by a SEPARATE process (the scene builders live in the
    product library; the process that is timed never maps it) and loaded here the way cell/main.cpp loads scene.vox."""
    d = "/dev/shm" if os.path.isdir("/dev/shm") else "/tmp"
    path = os.path.join(d, "yv_bench_%s_d%d_u%d.vox" % (a.scene, a.depth, os.getuid()))
    if not os.path.exists(path):
        build = ("IsoVolume(%d, seed=219, iso_level=200, threads=%d)" % (a.depth, cores)) if a.scene == "iso" \
            else ("SphereFractal(%d, threads=%d)" % (a.depth, cores))
        code = ("import sys; sys.path.insert(0, %r); import yoxel_voxel_b200 as yv; s = yv.SVOData.%s; s.Save(%r)"
                % (ROOT, build, path + ".tmp"))
        subprocess.check_call([sys.executable, "-c", code], stdout=subprocess.DEVNULL)
        os.replace(path + ".tmp", path)
    return path


def run_reference(a, rank):
    """--impl reference: the reference's CPU tracer on the host cores, same config. The headline value is the oracle port
    (kind = "port") because it can use every host core, which makes it the stronger baseline; the reference's own
    renderer compiled from its sources (oracle/_ref, TreadedRenderer: 4 threads by construction) is timed beside it
    as `reference_build` whenever the library is present, and its frame must equal the port's."""
    if rank != 0:
        return
    yvo = _yvo()
    cores = os.cpu_count() or 1
    vox = scene_vox_for_reference(a, cores)
    root, nodes = yvo.load_vox(vox)
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
    rb = ReferenceBuild(vox_path=vox)
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
        "config": config_of(a),
        "cpu_baseline": {"value": val, "unit": "Mrays/s", "cores": cores, "kind": "port",
                         "sample": "whole %dx%d frame per step, %d row strips (TreadedRenderer split)" % (a.width, a.height, cores)},
        "e2e": {"value": val, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "reference_build": ref_build,
        "setup": {"scene_file": vox, "scene_nodes": int(len(nodes)),
                  "scene_from": "written once by a separate process (SVOData.Save), loaded here by the oracle's .vox reader"},
    }
    emit(line)


# ---------------------------------------------------------------------------------------------------------------------
# one GPU, one workload: device time, V-bar, end-to-end time, parity, CPU baseline
# ---------------------------------------------------------------------------------------------------------------------

def make_renderer(yv, a, device, svo, devices=None):
    r = yv.SVORenderer(device) if devices is None else yv.SVORenderer(devices=devices)
    schedule = a.schedule if a.schedule != "auto" else "tiles"
    r.SetOption("schedule", {"tiles": 0, "persistent": 1, "queue": 2}[schedule])
    if a.smem_nodes >= 0:
        r.SetOption("smem_nodes", a.smem_nodes)
    if a.stack >= 0:
        r.SetOption("stack", a.stack)
    r.SetOption("cull", 1 if a.cull else 0)
    r.SetScene(svo)
    r.SetResolution(a.width, a.height)
    r.SetViewUp(UP); r.SetFOV(FOV)
    r.SetDetailCoef(a.detail)
    if a.ssna:
        r.SetSSNA(True, 1.0 / (1 << a.depth))
    if a.secondary:
        r.SetSecondary(voxel_size=1.0 / (1 << a.depth), **SEC_ARGS)
    return r, schedule


def git_head():
    try:
        return subprocess.run(["git", "-C", ROOT, "rev-parse", "--short=12", "HEAD"], capture_output=True, text=True,
                              timeout=10).stdout.strip() or None
    except Exception:
        return None


def kernel_source_sha():
    """sha1 over the two files the traversal kernel is compiled from: says whether a committed ncu capture still describes
    the kernel that ran."""
    import hashlib
    h = hashlib.sha1()
    for f in ("trace_core.cuh", "render_kernels.cuh"):
        try:
            h.update(open(os.path.join(ROOT, "yoxel-voxel_b200", "csrc", f), "rb").read())
        except OSError:
            return None
    return h.hexdigest()[:16]


def traffic_record(a):
    """DRAM / L2 bytes per launch from the committed `ncu --set full` capture of config 2 (profiles/traffic.json); other
    workloads have no capture and report null. `traffic_source` names the capture and the kernel sources it was taken on."""
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    is_cfg2 = (a.scene == "fractal" and a.depth == 12 and a.width == 1920 and a.height == 1080 and not a.secondary
               and not a.ssna and a.detail == 0 and not a.cull)
    if not (is_cfg2 and os.path.exists(tp)):
        return None, None, None
    try:
        tj = json.load(open(tp))
        now = kernel_source_sha()
        src = {"file": "profiles/traffic.json", "capture": tj.get("source"), "captured_at_commit": tj.get("commit"),
               "kernel_source_sha_at_capture": tj.get("kernel_source_sha"), "kernel_source_sha_of_this_run": now,
               "same_kernel_sources": (tj.get("kernel_source_sha") == now) if tj.get("kernel_source_sha") else None,
               "this_run_commit": git_head()}
        return tj.get("dram_bytes_per_launch"), tj.get("l2_bytes_per_launch"), src
    except Exception:
        return None, None, None


def measure_one_gpu(yv, torch, a, local, svo, build_s, steps, warm, cpu_mode, clocks=True):
    """cpu_mode: "full" = CPU baseline (1 / 4 / all threads) + parity + reference build; "parity" = one oracle frame
    (parity + a one-frame CPU figure); None = neither."""
    dev = torch.device("cuda", local)
    cores = os.cpu_count() or 1
    dev_bytes = svo.Upload(local)
    r, schedule = make_renderer(yv, a, local, svo)
    stream = torch.cuda.Stream(device=dev)       # a real (non-NULL) stream: the renderer, the L2 flush and the
    torch.cuda.set_stream(stream)                # timing events all run on it
    r.SetStream(stream.cuda_stream)
    pos, d = camera_for(0, a.scene)
    r.SetViewPos(pos); r.SetViewDir(d)
    frame_bytes = a.width * a.height * 4
    fb = torch.zeros(a.height, a.width, 4, dtype=torch.uint8, device=dev)

    # V-bar for the roofline: the kernel's own node-visit counters on this workload (tests/test_gpu_parity.py::
    # test_counters_equal_oracle_visits pins them to the oracle's count of the node fetch at cell/ppu_renderer.cpp:23)
    r.EnableCounters(True)
    r.SetOption("cull", 0)
    r.Render(fb.data_ptr(), sync=True)
    visits, pops = r.GetCounters()
    vis, pop = int(visits.sum()), int(pops.sum())
    hit_px = int((fb[:, :, 3] == 255).sum().item())
    kvis, kpop = vis, pop
    if a.cull:
        r.SetOption("cull", 1)
        r.Render(fb.data_ptr(), sync=True)
        visits, pops = r.GetCounters()
        kvis, kpop = int(visits.sum()), int(pops.sum())
    r.EnableCounters(False)
    del visits, pops
    px = a.width * a.height
    rays = px + (5 * hit_px if a.secondary else 0)             # shadow + 4 AO per hit pixel

    flush = None if a.no_l2_flush else torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    sampler = ClockSampler(local) if clocks else None
    for _ in range(warm):
        if flush is not None:
            flush.zero_()
        r.Render(fb.data_ptr(), sync=False)
    torch.cuda.synchronize()
    if sampler is not None and sampler.proc:
        # nvidia-smi spends its first ~0.2 s initialising NVML, which takes driver locks and stretched one 20-step timed
        # region by 13 % (0.784 instead of 0.691 ms per frame; the same box's e2e loop and ncu agreed on 0.69): keep the
        # GPU under the same load, untimed, until the first sample is out, so that initialisation is over before the
        # timed region starts and the samples beside it are samples under load
        t_wait = time.time()
        while not sampler.rows and time.time() - t_wait < 1.5:
            for _ in range(8):
                if flush is not None:
                    flush.zero_()
                r.Render(fb.data_ptr(), sync=False)
            torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    wall0 = time.time()
    for i in range(steps):
        if flush is not None:
            flush.zero_()                           # evict the node pool from L2 between frames (not timed)
        evs[i][0].record(stream)
        r.Render(fb.data_ptr(), sync=False)
        evs[i][1].record(stream)
    torch.cuda.synchronize()
    wall1 = time.time()
    launches_per_step = r.LastFrameLaunches()       # 1 (trace); 8 with SSNA (z, 5 x BlurZ, ShadeSimple)
    clk = sampler.stop(wall0, wall1) if sampler else None
    step_ms = [e0.elapsed_time(e1) for e0, e1 in evs]
    total_s = sum(step_ms) / 1e3
    value = rays * steps / total_s / 1e6

    # e2e: the public host API with host buffers (camera in, RGBA8 frame out)
    r.SetStream(0)
    e2e_t = []
    for i in range(warm + steps):
        if flush is not None:
            flush.zero_()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r.SetViewPos(pos); r.SetViewDir(d); r.SetViewUp(UP); r.SetFOV(FOV)      # host camera -> kernel params
        img = r.RenderFrame()                                                   # launch + delivery into pinned host memory + sync
        e2e_t.append(time.perf_counter() - t0)
    e2e_s = sum(e2e_t[warm:])
    e2e = {"value": rays * steps / e2e_s / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": 40,
           "d2h_bytes_per_step": frame_bytes, "ms_per_step": 1e3 * e2e_s / steps,
           "api": "yv_set_view_* + yv_render_frame: host camera in, pinned host RGBA8 frame out",
           "checksum": int(img[::16, ::16].astype(np.uint64).sum())}

    cpu_baseline, parity = None, None
    if cpu_mode:
        nodes, root = svo.nodes(copy=False), svo.GetRoot()
        o, dt0 = oracle_frame(nodes, root, a, 0, cores, want_visits=True)        # parity, V-bar cross-check
        gpu_img = r.RenderFrame()
        parity = {"rgba_identical_to_oracle": bool((gpu_img == o["rgba"]).all()),
                  "rays_identical": bool(o["stats"]["rays"] == int(rays)),
                  "node_visits_identical": bool(o["stats"]["node_visits"] == int(vis)),
                  "sample": "the whole %dx%d frame" % (a.width, a.height)}
        if cpu_mode == "full":
            o2, dt = oracle_frame(nodes, root, a, 0, cores)
        else:
            o2, dt = o, dt0
        cpu_baseline = {"value": o2["stats"]["rays"] / dt / 1e6, "unit": "Mrays/s", "cores": cores, "kind": "port",
                        "sample": "one whole %dx%d frame, %d row strips (%.2f s)" % (a.width, a.height, cores, dt),
                        "ms_per_frame": 1e3 * dt}
        if cpu_mode == "full":
            # the CPU tracer at 1 thread, at the reference's constant of 4 (ppu_renderer.cpp:129) and on every core
            sweep = {}
            for t in sorted({1, 4, cores}):
                if t == cores:
                    sweep[str(t)] = cpu_baseline["value"]
                else:
                    ot, dtt = oracle_frame(nodes, root, a, 0, t)
                    sweep[str(t)] = ot["stats"]["rays"] / dtt / 1e6
            cpu_baseline["mrays_by_threads"] = sweep
            rb = ReferenceBuild(svo=svo)
            if rb.ok(a):
                # the frame of the reference's own renderer (cell/ppu_renderer.cpp compiled unmodified, oracle/_ref)
                ref_img, ref_dt = rb.frame(a, 0)
                parity["rgba_identical_to_reference_build"] = bool((gpu_img == ref_img).all())
                cpu_baseline["reference_build"] = {"value": a.width * a.height / ref_dt / 1e6, "unit": "Mrays/s", "cores": 4,
                                                   "kind": "reference", "what": "TreadedRenderer, one whole frame (%.2f s)" % ref_dt}
        del nodes, o, o2

    peak, peak_src = hbm_peak()
    alg_bytes = vis * NODE_BYTES + px * PIXEL_BYTES
    kernel_s = total_s / steps
    achieved = alg_bytes / kernel_s / 1e9
    traffic, l2_traffic, tsrc = traffic_record(a)
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "l2_traffic": l2_traffic, "traffic_source": tsrc, "peak_source": peak_src,
                "kernel": "yv::render_frame<%s>" % schedule, "algorithmic_bytes_per_launch": alg_bytes,
                "node_visits_per_ray": vis / rays, "pop_refetches_per_ray": pop / rays,
                "kernel_node_fetches_per_ray": kvis / rays, "kernel_pop_refetches_per_ray": kpop / rays,
                "octant_culling": bool(a.cull), "kernel_ms": 1e3 * kernel_s}
    setup = {"schedule": schedule, "smem_nodes": r.GetOption("smem_nodes"), "stack": r.GetOption("stack"),
             "l2": "flushed between frames (256 MiB write, untimed)" if flush is not None
                   else "not flushed; node pool %d MB > L2" % (dev_bytes >> 20),
             "nodes": svo.nodecount, "packed_bytes": dev_bytes, "scene_build_s": round(build_s, 2),
             "partition": "single GPU", "gather": "none", "flythrough": False, "hit_fraction": round(hit_px / float(px), 4)}
    rec = {"value": value, "ms_per_step": 1e3 * total_s / steps, "rays_per_step": float(rays), "roofline": roofline,
           "cpu_baseline": cpu_baseline, "e2e": e2e, "clocks": clk, "gpu_launches": steps * launches_per_step,
           "e2e_gpu_launches_per_step": r.LastFrameLaunches(), "parity": parity, "setup": setup}
    r.close()
    del fb, flush
    torch.cuda.set_stream(torch.cuda.default_stream(dev))
    return rec


def host_mem_available_gb():
    try:
        for line in open("/proc/meminfo"):
            if line.startswith("MemAvailable:"):
                return int(line.split()[1]) / 1048576.0
    except OSError:
        pass
    return 0.0


def auto_depth(requested):
    """Deepest iso volume this host builds in reasonable time: depth 14 is 496 M nodes = 19.8 GB of host pool; depth 13
    a quarter of that."""
    if requested:
        return requested
    cores = os.cpu_count() or 1
    return 14 if (cores >= 12 and host_mem_available_gb() >= 64.0) else 13


def run_extras(yv, torch, a, local):
    """BASELINE configs 1, 4 and 3 on one GPU, after the headline line. Each record has the headline's shape."""
    out = {}
    cores = os.cpu_count() or 1
    specs = [
        ("config1", dict(scene="fractal", depth=10, width=512, height=512, secondary=False), 50, 30.0),
        ("config4", dict(scene="fractal", depth=12, width=1920, height=1080, secondary=True), 20, 40.0),
        ("config3", dict(scene="iso", depth=auto_depth(a.extras_depth), width=3840, height=2160, secondary=False), 20, 150.0),
    ]
    for name, over, steps, need in specs:
        spent = time.time() - T_START
        if spent + need > a.budget_s:
            out[name] = {"skipped": "budget: %.0f s spent of --budget-s %.0f, this config needs ~%.0f s" % (spent, a.budget_s, need)}
            continue
        b = argparse.Namespace(**vars(a))
        for k, v in over.items():
            setattr(b, k, v)
        b.ssna, b.detail, b.cull = False, 0.0, False
        try:
            t0 = time.time()
            svo = build_scene(yv, b, cores)
            build_s = time.time() - t0
            rec = measure_one_gpu(yv, torch, b, local, svo, build_s, steps, 3, "parity", clocks=False)
            rec = dict({"metric": "Mrays/s", "unit": "Mrays/s", "steps": steps, "warmup": 3, "config": config_of(b)}, **rec)
            out[name] = rec
            del svo
        except Exception as ex:                                   # the headline line must survive
            out[name] = {"error": "%s: %s" % (type(ex).__name__, ex)}
        torch.cuda.empty_cache()
    return out


def run_single(a):
    import torch
    import yoxel_voxel_b200 as yv
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    cores = os.cpu_count() or 1
    t0 = time.time()
    svo = build_scene(yv, a, cores)
    build_s = time.time() - t0
    warm = max(a.warmup, 3)
    rec = measure_one_gpu(yv, torch, a, local, svo, build_s, a.steps, warm, None if a.no_cpu_baseline else "full")
    line = {"metric": "Mrays/s", "value": rec["value"], "unit": "Mrays/s", "n_gpus": 1, "steps": a.steps, "warmup": warm,
            "ms_per_step": rec["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": config_of(a), "setup": rec["setup"],
            "frame_ms": rec["ms_per_step"], "rays_per_step": rec["rays_per_step"], "roofline": rec["roofline"],
            "cpu_baseline": rec["cpu_baseline"], "e2e": rec["e2e"], "clocks": rec["clocks"], "gpu_launches": rec["gpu_launches"],
            "e2e_gpu_launches_per_step": rec["e2e_gpu_launches_per_step"], "parity": rec["parity"]}
    is_headline = (a.scene == "fractal" and a.depth == 12 and a.width == 1920 and a.height == 1080 and not a.secondary
                   and not a.ssna and a.detail == 0 and not a.cull)
    if is_headline and not a.no_extras:
        del svo
        line["configs"] = run_extras(yv, torch, a, local)
    line["wall_s"] = round(time.time() - T_START, 1)
    emit(line)


# ---------------------------------------------------------------------------------------------------------------------
# BASELINE config 5 through ONE renderer handle over N GPUs
# ---------------------------------------------------------------------------------------------------------------------

def nvlink_counters(n):
    """Cumulative NVLink data counters (KiB) per GPU, summed over its links: [tx, rx] per GPU. NVML field values first
    (per-link query, then the all-links scope), `nvidia-smi nvlink -gt d` as the fallback; None where neither reports."""
    out = None
    try:
        import pynvml
        pynvml.nvmlInit()
        out = []
        for i in range(n):
            h = pynvml.nvmlDeviceGetHandleByIndex(i)
            tot = [0, 0]
            got = False
            for link in range(18):
                vals = pynvml.nvmlDeviceGetFieldValues(h, [(pynvml.NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_TX, link),
                                                           (pynvml.NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_RX, link)])
                for j, v in enumerate(vals):
                    if v.nvmlReturn == 0:
                        tot[j] += int(v.value.ullVal)
                        got = True
            out.append(tot if got else [None, None])
        if all(o[0] is None for o in out):
            out = None
    except Exception:
        out = None
    if out is not None:
        return out
    try:
        import re
        out = []
        for i in range(n):
            txt = subprocess.run(["nvidia-smi", "nvlink", "-gt", "d", "-i", str(i)], capture_output=True, text=True, timeout=20).stdout
            tx = [int(m) for m in re.findall(r"Data Tx:\s*(\d+)\s*KiB", txt)]
            rx = [int(m) for m in re.findall(r"Data Rx:\s*(\d+)\s*KiB", txt)]
            out.append([sum(tx), sum(rx)] if (tx or rx) else [None, None])
        return None if all(o[0] is None for o in out) else out
    except Exception:
        return None


def run_strong_8k(yv, torch, a, ndev):
    """BASELINE config 5: 7680x4320, 64-frame Catmull-Rom flythrough of the deep iso volume. One yv_renderer over `ndev`
    GPUs (yv_renderer_create_group): scene packed on GPU 0 and replicated by peer copies, each frame cut into interleaved
    32-row blocks (blockStart / blockStride, cell/spu/trace_spu.cpp:162-177), every GPU storing its blocks into the frame.
    The same frames are rendered by a single-GPU renderer in the same run."""
    cores = os.cpu_count() or 1
    depth = auto_depth(a.strong_depth)
    W, H, F = a.strong_width, a.strong_height, a.strong_frames
    b = argparse.Namespace(**vars(a))
    b.scene, b.depth, b.width, b.height = "iso", depth, W, H
    b.secondary, b.ssna, b.detail, b.cull = False, False, 0.0, False
    rec = {"workload": workload_name(b) + ", %d-frame Catmull-Rom flythrough" % F, "n_gpus": ndev, "frames": F,
           "depth": depth, "depth_choice": ("--strong-depth" if a.strong_depth else
                                            "auto: 14 needs >= 12 host cores and >= 64 GB free host memory (have %d, %.0f GB)"
                                            % (cores, host_mem_available_gb())),
           "partition": "interleaved %d-row blocks, block b on GPU b %% %d" % (a.band_rows, ndev)}
    t0 = time.time()
    svo = build_scene(yv, b, cores)
    rec["scene_build_s"] = round(time.time() - t0, 1)
    rec["nodes"] = svo.nodecount
    frame_bytes = W * H * 4
    rays_frame = W * H
    dev0 = torch.device("cuda", 0)
    t0 = time.time()
    rec["packed_bytes"] = svo.Upload(0)
    rec["upload_pack_s"] = round(time.time() - t0, 2)

    cams = [camera_for(f, "iso", F) for f in range(F)]
    flushes = [torch.empty(256 << 20, dtype=torch.uint8, device=torch.device("cuda", k)) for k in range(ndev)]

    def flush_all(n):
        for fl in flushes[:n]:
            fl.zero_()
        for k in range(n):
            torch.cuda.synchronize(k)

    def run_frames(r, dst, n, keep=None, members=None):
        """every frame synchronous into device memory `dst`; device time from the handle's own events"""
        ms = []
        for f in range(F):
            flush_all(n)
            r.SetViewPos(cams[f][0]); r.SetViewDir(cams[f][1])
            r.Render(dst.data_ptr(), sync=True)
            ms.append(r.LastFrameMs())
            if members is not None:
                members.append(r.MemberFrameMs())
            if keep is not None and f in keep:
                keep[f] = dst.clone()
        return ms

    def e2e_pipeline(r, zero_copy=1):
        """frames in flight: camera in, pinned host frame out, three frames outstanding. zero_copy 1: the kernels store
        their pixels into the host frame; 0: frames are drawn in HBM and moved by the copy engines (every GPU its own
        blocks) while the next frame traverses"""
        r.SetOption("slots", 3)
        r.SetOption("zero_copy", zero_copy)
        tickets, last = [], None
        t0 = time.perf_counter()
        for f in range(F):
            r.SetViewPos(cams[f][0]); r.SetViewDir(cams[f][1])
            tickets.append(r.RenderFrameAsync())
            if len(tickets) == 3:
                r.WaitFrame(tickets.pop(0), as_array=False)
        while tickets:
            last = r.WaitFrame(tickets.pop(0), as_array=True)
        return time.perf_counter() - t0, last

    def e2e_sync(r):
        t0 = time.perf_counter()
        img = None
        for f in range(F):
            r.SetViewPos(cams[f][0]); r.SetViewDir(cams[f][1])
            img = r.RenderFrame()
        return time.perf_counter() - t0, img

    # V-bar (single GPU, counters on) of every 8th frame, for the roofline
    r1, _ = make_renderer(yv, b, 0, svo)
    fb1 = torch.zeros(H, W, 4, dtype=torch.uint8, device=dev0)
    r1.EnableCounters(True)
    vis_total, probe = 0, list(range(0, F, max(1, F // 8)))
    for f in probe:
        r1.SetViewPos(cams[f][0]); r1.SetViewDir(cams[f][1])
        r1.Render(fb1.data_ptr(), sync=True)
        visits, _ = r1.GetCounters()
        vis_total += int(visits.sum())
    vbar = vis_total / float(len(probe) * rays_frame)
    r1.EnableCounters(False)
    del visits

    check = sorted(set([0, F // 3, (2 * F) // 3, F - 1]))
    keep1 = {f: None for f in check}
    for _ in range(2):
        r1.Render(fb1.data_ptr(), sync=True)
    ms1 = run_frames(r1, fb1, 1, keep1)
    e2e_pipeline(r1)
    e1_pipe, _ = e2e_pipeline(r1)
    e2e_pipeline(r1, 0)
    e1_staged, _ = e2e_pipeline(r1, 0)
    r1.SetOption("zero_copy", 1)
    e1_sync, _ = e2e_sync(r1)
    r1.close()

    rN, _ = make_renderer(yv, b, 0, svo, devices=list(range(ndev)))
    rN.SetPartition("interleaved", a.band_rows)
    fbN = torch.zeros(H, W, 4, dtype=torch.uint8, device=dev0)
    for _ in range(2):
        rN.Render(fbN.data_ptr(), sync=True)
    rep_ms, rep_bytes = rN.ReplicateStats()
    keepN = {f: None for f in check}
    nv0 = nvlink_counters(ndev)
    member_all = []
    msN = run_frames(rN, fbN, ndev, keepN, member_all)
    nv1 = nvlink_counters(ndev)
    member_ms = rN.MemberFrameMs()
    rN.SetOption("group_threads", 0)                         # all N launches from one host loop (what round 2 started with)
    msN_serial = run_frames(rN, fbN, ndev)
    rN.SetOption("group_threads", 1)
    same_1gpu = all(bool(torch.equal(keep1[f], keepN[f])) for f in check)
    e2e_pipeline(rN)
    eN_pipe, last_img = e2e_pipeline(rN)
    last_dev = torch.from_numpy(np.ascontiguousarray(last_img)).to(dev0)
    e2e_pipeline(rN, 0)
    eN_staged, staged_img = e2e_pipeline(rN, 0)
    staged_dev = torch.from_numpy(np.ascontiguousarray(staged_img)).to(dev0)
    rN.SetOption("zero_copy", 1)
    eN_sync, sync_img = e2e_sync(rN)
    sync_dev = torch.from_numpy(np.ascontiguousarray(sync_img)).to(dev0)
    host_same = bool(torch.equal(last_dev, keepN[F - 1])) and bool(torch.equal(sync_dev, keepN[F - 1])) and \
        bool(torch.equal(staged_dev, keepN[F - 1]))
    del last_dev, sync_dev, staged_dev

    # oracle: sampled 16-row bands of the checked frames
    yvo = _yvo()
    nodes, root = svo.nodes(copy=False), svo.GetRoot()
    bands = [(y, y + 16) for y in range(8, H - 16, max(16, H // 6))]
    oracle_ok, rows_checked = True, 0
    for f in check:
        img = keepN[f].cpu().numpy()
        cam = yvo.camera(cams[f][0], cams[f][1], UP, FOV, W, H)
        for (y0, y1) in bands:
            o = yvo.render(nodes, root, cam, threads=cores, rows=(y0, y1))
            oracle_ok = oracle_ok and bool((o["rgba"][y0:y1] == img[y0:y1]).all())
            rows_checked += y1 - y0
    del nodes
    rN.close()

    t1, tN = sum(ms1) / 1e3, sum(msN) / 1e3
    peak, peak_src = hbm_peak()
    alg = vbar * rays_frame * NODE_BYTES + rays_frame * PIXEL_BYTES
    nvl = None
    if nv0 and nv1:
        try:
            nvl = {"gpu%d" % k: {"tx_bytes": (nv1[k][0] - nv0[k][0]) * 1024 if nv0[k][0] is not None else None,
                                 "rx_bytes": (nv1[k][1] - nv0[k][1]) * 1024 if nv0[k][1] is not None else None}
                   for k in range(ndev)}
            nvl["expected_rx_gpu0_bytes"] = int(F * frame_bytes * (ndev - 1) / ndev)
        except Exception:
            nvl = None
    imb = None
    if member_ms and min(member_ms) > 0:
        imb = {"member_ms_last_frame": [round(m, 3) for m in member_ms],
               "max_over_mean": max(member_ms) / (sum(member_ms) / len(member_ms))}
    rec.update({
        "one_gpu": {"ms_per_frame": 1e3 * t1 / F, "value": rays_frame * F / t1 / 1e6,
                    "roofline_frac": alg / (t1 / F) / 1e9 / peak,
                    "e2e_ms_per_frame_stores": 1e3 * e1_pipe / F, "e2e_ms_per_frame_copy_engine": 1e3 * e1_staged / F,
                    "e2e_ms_per_frame_sync": 1e3 * e1_sync / F},
        "ms_per_frame": 1e3 * tN / F, "value": rays_frame * F / tN / 1e6, "unit": "Mrays/s",
        "speedup": t1 / tN, "efficiency": t1 / tN / ndev,
        "launch": "every GPU's share issued by its own host thread (option group_threads 1); from one loop on the calling "
                  "thread the same frames take %.3f ms each (efficiency %.3f)"
                  % (1e3 * sum(msN_serial) / 1e3 / F, t1 / (sum(msN_serial) / 1e3) / ndev),
        "roofline_frac_per_gpu": alg / ndev / (tN / F) / 1e9 / peak, "node_visits_per_ray": vbar,
        "delivery": "device-timed frames: every GPU's kernel stores its blocks into one frame in GPU 0's HBM (peer stores over NVLink)",
        "e2e": {"value": rays_frame * F / min(eN_pipe, eN_staged) / 1e6, "unit": "Mrays/s",
                "ms_per_frame": 1e3 * min(eN_pipe, eN_staged) / F,
                "delivery": "copy engines (zero_copy 0)" if eN_staged < eN_pipe else "kernel stores (zero_copy 1)",
                "ms_per_frame_stores": 1e3 * eN_pipe / F, "ms_per_frame_copy_engine": 1e3 * eN_staged / F,
                "ms_per_frame_sync": 1e3 * eN_sync / F, "value_sync": rays_frame * F / eN_sync / 1e6,
                "h2d_bytes_per_step": 40, "d2h_bytes_per_step": frame_bytes,
                "api": "yv_set_view_* + yv_render_frame_async / yv_wait_frame, 3 frames in flight, renderer-owned pinned host "
                       "frames; stores = every GPU's kernel writes its blocks into the host frame over its own PCIe link, "
                       "copy_engine = frames drawn in HBM, every GPU's copy engine moves its blocks while the next frame "
                       "traverses; sync = yv_render_frame per frame",
                "speedup_vs_1gpu_e2e": min(e1_pipe, e1_staged) / min(eN_pipe, eN_staged)},
        "identical_to_1gpu": same_1gpu, "host_frames_identical_to_device_frames": host_same,
        "identical_to_oracle": oracle_ok,
        "oracle_sample": "%d frames x %d bands of 16 rows (%d rows of %d pixels)" % (len(check), len(bands), rows_checked, W),
        "frames_compared": check,
        "replicate": {"ms": rep_ms, "bytes": rep_bytes, "gb_per_s": (rep_bytes / 1e9) / (rep_ms / 1e3) if rep_ms > 0 else None,
                      "what": "packed pool GPU 0 -> the other GPUs, cudaMemcpyPeerAsync, all peers concurrently; "
                              "against upload_pack_s for one upload through the host + re-pack"},
        "imbalance": imb, "nvlink": nvl,
        "l2": "flushed on every GPU between frames (256 MiB write, untimed)",
    })
    if imb:
        ideal = 1e3 * t1 / F / ndev
        ok = [m for m in member_all if m and min(m) > 0]
        slowest = sum(max(m) for m in ok) / max(1, len(ok))
        mean_share = sum(sum(m) / len(m) for m in ok) / max(1, len(ok))
        imb["slowest_member_ms_mean_over_frames"] = slowest
        imb["mean_member_ms_mean_over_frames"] = mean_share
        rec["limit"] = ("per frame, mean over the %d frames: ideal (1-GPU time / N) %.3f ms; a GPU's own share takes %.3f ms on "
                        "average (a 1/N share of interleaved blocks costs more than 1/N of the frame: shorter launch, same tail) "
                        "and %.3f ms on the slowest GPU (imbalance across blocks); the frame, from the leader's first event to "
                        "the join of all members, %.3f ms (launch skew + join: %.3f ms)"
                        % (len(ok), ideal, mean_share, slowest, 1e3 * tN / F, 1e3 * tN / F - slowest))
    return rec


def run_strong_only(a):
    import torch
    import yoxel_voxel_b200 as yv
    n = max(1, min(a.gpus, torch.cuda.device_count()))
    rec = run_strong_8k(yv, torch, a, n)
    emit({"strong_8k": rec, "wall_s": round(time.time() - T_START, 1)})


# ---------------------------------------------------------------------------------------------------------------------
# N > 1: one process per GPU
# ---------------------------------------------------------------------------------------------------------------------

def run_ranks(a):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    import torch
    import torch.distributed as dist
    import yoxel_voxel_b200 as yv
    from yoxel_voxel_b200 import multigpu

    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")      # stdout carries the one JSON line and nothing else
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    idle = dist.new_group(backend="gloo")                        # host-side waits: an idle rank must not spin a kernel
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    cores = os.cpu_count() or 1
    tiles_mode = a.partition in ("tiles", "bands")
    if not tiles_mode and not a.same_frame:
        a.flythrough = True

    # ---- scene: replicated on every GPU ----------------------------------------------------------
    # rank 0 builds with all host cores and parks the .vox in /dev/shm, the other ranks load it (SVOData::Load) — the
    # build is host work shared by the box. (A process cannot read another process's device memory without IPC; the
    # single-handle renderer of strong_8k replicates the packed pool over NVLink instead.)
    t0 = time.time()
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
    build_s = time.time() - t0
    dev_bytes = svo.Upload(local)

    r, schedule = make_renderer(yv, a, local, svo)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    r.SetStream(stream.cuda_stream)

    # ---- partition ----------------------------------------------------------------------------
    frame_bytes = a.width * a.height * 4
    my_rows = np.arange(a.height)
    if tiles_mode:
        if a.partition == "bands":
            y0, y1 = multigpu.row_band(rank, world, a.height)
            r.SetRows(y0, y1)
            my_rows = np.arange(y0, y1)
        else:
            r.SetInterleave(a.band_rows, world, rank)
            my_rows = multigpu.interleaved_rows(rank, world, a.height, a.band_rows)
        target_bytes = frame_bytes
    else:
        target_bytes = frame_bytes * world

    def frame_of(step, rk=rank):
        if not a.flythrough:
            return 0
        return step if tiles_mode else step * world + rk

    pos, d = camera_for(frame_of(0), a.scene)
    r.SetViewPos(pos); r.SetViewDir(d)

    gather = a.gather
    target_ptr, target_obj, local_fb, gather_list = None, None, None, None
    if gather == "p2p":
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
    if gather == "nccl":
        local_fb = torch.zeros(a.height, a.width, 4, dtype=torch.uint8, device=dev)
        if rank == 0:
            gather_list = [torch.zeros_like(local_fb) for _ in range(world)]

    def dst_ptr():
        if gather == "p2p":
            return target_ptr + (0 if tiles_mode else rank * frame_bytes)
        return local_fb.data_ptr()

    def render_step(step=None):
        if a.flythrough and step is not None:
            fpos, fdir = camera_for(frame_of(step), a.scene)
            r.SetViewPos(fpos); r.SetViewDir(fdir)
        r.Render(dst_ptr(), sync=False)
        if gather == "nccl":
            dist.gather(local_fb, gather_list, dst=0)

    # ---- V-bar for the roofline: the kernel's own node-visit counters on this rank's frames ----------
    r.EnableCounters(True)
    r.SetOption("cull", 0)
    probe = torch.zeros(a.height, a.width, 4, dtype=torch.uint8, device=dev)
    vis_sum = pop_sum = hit_px = 0
    probe_frames = list(range(a.steps)) if a.flythrough else [None]
    rows_t = torch.as_tensor(my_rows, device=dev)
    for st in probe_frames:
        if st is not None:
            fpos, fdir = camera_for(frame_of(st), a.scene)
            r.SetViewPos(fpos); r.SetViewDir(fdir)
        r.Render(probe.data_ptr(), sync=True)
        visits, pops = r.GetCounters()
        vis_sum += int(visits[my_rows].sum())
        pop_sum += int(pops[my_rows].sum())
        hit_px += int((probe[rows_t, :, 3] == 255).sum().item())
    r.EnableCounters(False)
    r.SetOption("cull", 1 if a.cull else 0)
    rep = a.steps // len(probe_frames)                         # 1 with a flythrough, K with a fixed camera
    my_px = len(my_rows) * a.width
    my_rays_total = my_px * a.steps + (5 * hit_px * rep if a.secondary else 0)      # over the whole timed region
    vis_total, pop_total = vis_sum * rep, pop_sum * rep
    hit_frac = hit_px / float(len(probe_frames) * max(1, my_px))
    del probe

    flush = None if a.no_l2_flush else torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def sync_all():
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local) if rank == 0 else None
    warm = max(a.warmup, 3)
    for _ in range(warm):
        if flush is not None:
            flush.zero_()
        render_step()
    sync_all()
    # nvidia-smi's NVML initialisation is over before the timed region starts (see measure_one_gpu): rank 0 decides,
    # every rank keeps rendering untimed meanwhile
    t_wait = time.time()
    while True:
        go = torch.tensor([1 if (sampler is None or not sampler.proc or sampler.rows or time.time() - t_wait > 1.5) else 0], device=dev)
        dist.broadcast(go, 0)
        if int(go.item()) == 1:
            break
        for _ in range(8):
            if flush is not None:
                flush.zero_()
            render_step()
        sync_all()

    # ---- timed region: K steps, CUDA events on the launching stream ------------------------------
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
    sync_all()
    wall0 = time.time()
    for i in range(a.steps):
        if flush is not None:
            flush.zero_()                           # evict the node pool from L2 between frames (not timed)
        evs[i][0].record(stream)
        render_step(i)
        evs[i][1].record(stream)
        torch.cuda.synchronize()
        dist.barrier()                              # the batch on GPU 0 is complete once every rank has stored
    sync_all()
    wall1 = time.time()
    launches_per_step = r.LastFrameLaunches()
    clocks = sampler.stop(wall0, wall1) if sampler else None
    step_ms = [e0.elapsed_time(e1) for e0, e1 in evs]
    my_ms = sum(step_ms)
    total_ms = torch.tensor([my_ms], dtype=torch.float64, device=dev)
    sums = torch.tensor([float(my_rays_total), float(my_px * a.steps), float(vis_total), float(pop_total)],
                        dtype=torch.float64, device=dev)
    per_rank = torch.zeros(world, dtype=torch.float64, device=dev)
    per_rank[rank] = my_ms / a.steps
    dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)          # max over ranks
    dist.all_reduce(sums, op=dist.ReduceOp.SUM)              # units all ranks processed
    dist.all_reduce(per_rank, op=dist.ReduceOp.SUM)
    total_s = float(total_ms.item()) / 1e3
    rays_all, px_all, vis_all, pop_all = (float(v) for v in sums.tolist())
    value = rays_all / total_s / 1e6
    rays_step = rays_all / a.steps

    # ---- e2e: one host frame (tiles) / one host batch of N frames (frames) in shared memory, registered with every GPU:
    # each rank stores its pixels straight into it over its own PCIe link (yv_host_register + yv_render_frame_device),
    # host camera in, host pixels out. Timed per step on the host clock around the synchronous call, L2 flushed and
    # ranks aligned by a barrier outside the timed region, max over ranks.
    e2e = None
    try:
        shared = multigpu.SharedHostFrame(dist, rank, world, local, target_bytes, os.environ.get("MASTER_PORT", "0"))
    except yv.YVError as ex:
        shared = None
        if rank == 0:
            print("shared host frame unavailable (%s)" % ex, file=sys.stderr)
    okf = torch.tensor([1 if shared is not None else 0], device=dev)
    dist.all_reduce(okf, op=dist.ReduceOp.MIN)
    host_batch = None
    if int(okf.item()) == 1:
        h_dst = shared.ptr + (0 if tiles_mode else rank * frame_bytes)
        e2e_t = []
        for i in range(warm + a.steps):
            if flush is not None:
                flush.zero_()
            torch.cuda.synchronize()
            dist.barrier()
            t0 = time.perf_counter()
            fpos, fdir = camera_for(frame_of(max(0, i - warm)), a.scene)
            r.SetViewPos(fpos); r.SetViewDir(fdir); r.SetViewUp(UP); r.SetFOV(FOV)
            r.Render(h_dst, sync=True)
            e2e_t.append(time.perf_counter() - t0)
        e2e_s = torch.tensor([sum(e2e_t[warm:])], dtype=torch.float64, device=dev)
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
        dist.barrier()
        same = None
        if rank == 0:
            host_batch = shared.array.copy()
            if gather == "p2p":
                # what the GPUs stored into host memory == what they stored into GPU 0 over NVLink (both hold the last step)
                ref = np.empty(target_bytes, np.uint8)
                yv.lib().yv_copy_to_host(local, ctypes.c_void_p(ref.ctypes.data), ctypes.c_void_p(target_ptr), target_bytes)
                same = bool((ref == host_batch).all())
        e2e = {"value": rays_all / float(e2e_s.item()) / 1e6, "unit": "Mrays/s",
               "h2d_bytes_per_step": 40 * world, "d2h_bytes_per_step": target_bytes,
               "ms_per_step": 1e3 * float(e2e_s.item()) / a.steps,
               "api": "yv_set_view_* + yv_render_frame_device into one pinned host %s shared by the ranks (yv_host_register): "
                      "every GPU stores its pixels over its own PCIe link" % ("frame" if tiles_mode else "batch of %d frames" % world),
               "identical_to_nvlink_gathered": same}
        if not tiles_mode:
            # the same batch with frames in flight: every rank keeps three frames outstanding (yv_render_frame_async /
            # yv_wait_frame), drawn in HBM and moved to renderer-owned pinned host frames by the GPU's copy engine while the
            # next frame traverses. No barrier and no L2 flush inside the timed region (the kernel is insensitive to it:
            # profiles/README.md); barrier + synchronize on both sides, max over ranks.
            # (collectives stay outside the try blocks: a rank that fails still meets the others at every barrier / reduce)
            pipe_ok, pipe_t, same_local, pipe_err = 1, 0.0, 0, ""
            try:
                r.SetOption("slots", 3)
                r.SetOption("zero_copy", 0)
            except Exception as ex:
                pipe_ok, pipe_err = 0, "%s: %s" % (type(ex).__name__, ex)
            for rep in range(2):                                       # the first repetition warms the slots up
                torch.cuda.synchronize()
                dist.barrier()
                if not pipe_ok:
                    continue
                try:
                    t0 = time.perf_counter()
                    tickets, last_ptr = [], None
                    for i in range(a.steps):
                        fpos, fdir = camera_for(frame_of(i), a.scene)
                        r.SetViewPos(fpos); r.SetViewDir(fdir); r.SetViewUp(UP); r.SetFOV(FOV)
                        tickets.append(r.RenderFrameAsync())
                        if len(tickets) == 3:
                            last_ptr = r.WaitFrame(tickets.pop(0), as_array=False)
                    while tickets:
                        last_ptr = r.WaitFrame(tickets.pop(0), as_array=False)
                    pipe_t = time.perf_counter() - t0
                    if rep == 1:
                        # the last frame delivered this way == the frame this rank stored into the shared batch (same camera)
                        last_img = np.frombuffer((ctypes.c_uint8 * frame_bytes).from_address(last_ptr), np.uint8)
                        mine = np.frombuffer((ctypes.c_uint8 * frame_bytes).from_address(h_dst), np.uint8)
                        same_local = 1 if bool((mine == last_img).all()) else 0
                except Exception as ex:
                    pipe_ok, pipe_err = 0, "%s: %s" % (type(ex).__name__, ex)
            try:
                r.SetOption("zero_copy", 1)
            except Exception:
                pass
            ps = torch.tensor([pipe_t], dtype=torch.float64, device=dev)
            dist.all_reduce(ps, op=dist.ReduceOp.MAX)
            okp = torch.tensor([pipe_ok, same_local], device=dev)
            dist.all_reduce(okp, op=dist.ReduceOp.MIN)
            sync_val = e2e["value"]
            e2e["sync"] = {"value": sync_val, "ms_per_step": e2e["ms_per_step"],
                           "what": "one synchronous call per step, kernels store into the shared host batch, barrier per step"}
            if int(okp[0].item()) == 1 and float(ps.item()) > 0:
                e2e["frames_in_flight"] = {"value": rays_all / float(ps.item()) / 1e6, "ms_per_step": 1e3 * float(ps.item()) / a.steps,
                                           "identical_to_sync_delivery": bool(int(okp[1].item()) == 1),
                                           "what": "3 frames outstanding per GPU, copy-engine delivery into renderer-owned pinned "
                                                   "host frames, no barrier or L2 flush inside the timed region"}
                if e2e["frames_in_flight"]["value"] > sync_val and e2e["frames_in_flight"]["identical_to_sync_delivery"]:
                    e2e["value"] = e2e["frames_in_flight"]["value"]
                    e2e["ms_per_step"] = e2e["frames_in_flight"]["ms_per_step"]
                    e2e["api"] = ("yv_set_view_* + yv_render_frame_async / yv_wait_frame per rank, 3 frames in flight, copy-engine "
                                  "delivery into pinned host frames (the synchronous shared-batch form is in `sync`)")
            else:
                e2e["frames_in_flight"] = {"error": pipe_err or "failed on another rank"}
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
            host_batch = pinned.numpy().copy()
        elif rank == 0:
            t0 = time.perf_counter()
            host_batch = torch.stack(gather_list).cpu().numpy().reshape(-1)
            d2h_s = time.perf_counter() - t0
        e2e = {"value": rays_all / (total_s + d2h_s * a.steps) / 1e6, "unit": "Mrays/s",
               "h2d_bytes_per_step": 40 * world, "d2h_bytes_per_step": target_bytes,
               "api": "yv_render_frame_device into GPU 0's buffer (%s) + D2H of the gathered pixels on rank 0" % gather}

    # ---- rank 0 alone from here; the other ranks wait on the host (gloo), their GPUs idle ----------------------------
    if rank != 0:
        dist.barrier(group=idle)
        dist.barrier()
        dist.destroy_process_group()
        return

    batch_1gpu, parity, strong = None, None, None
    try:
        # the same batch on ONE GPU: every frame the N ranks rendered in the timed region, rendered by this GPU alone
        if tiles_mode:
            if a.partition == "bands":
                r.SetRows(0, a.height)                           # the whole frame instead of this rank's share
            else:
                r.SetInterleave(a.band_rows, 1, 0)
        one_fb = torch.zeros(a.height, a.width, 4, dtype=torch.uint8, device=dev)
        frames = [frame_of(st, rk) for st in range(a.steps) for rk in (range(world) if not tiles_mode else [0])]
        ms = 0.0
        for _ in range(3):
            r.Render(one_fb.data_ptr(), sync=True)
        for f in frames:
            if flush is not None:
                flush.zero_()
            fpos, fdir = camera_for(f, a.scene)
            r.SetViewPos(fpos); r.SetViewDir(fdir)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            r.Render(one_fb.data_ptr(), sync=False)
            e1.record(stream)
            torch.cuda.synchronize()
            ms += e0.elapsed_time(e1)
        batch_1gpu = {"frames": len(frames), "ms_total": ms, "ms_per_frame": ms / len(frames),
                      "n_gpu_ms_total": 1e3 * total_s, "speedup": ms / (1e3 * total_s), "efficiency": ms / (1e3 * total_s) / world,
                      "per_rank_ms_per_step": [round(float(v), 4) for v in per_rank.tolist()],
                      "what": "the %d frames of the timed region rendered by GPU 0 alone in the same run (device time, L2 "
                              "flushed); speedup = that time / the N-GPU time of the timed region" % len(frames)}
        # parity of the delivered pixels: the frame the LAST rank delivered at the last step against this GPU's own render
        # of it and against an oracle band
        if host_batch is not None:
            last_rank = world - 1
            f = frame_of(a.steps - 1, last_rank)
            fpos, fdir = camera_for(f, a.scene)
            r.SetViewPos(fpos); r.SetViewDir(fdir)
            r.Render(one_fb.data_ptr(), sync=True)
            mine = one_fb.cpu().numpy()
            got = host_batch.reshape(-1)[(0 if tiles_mode else last_rank * frame_bytes):][:frame_bytes].reshape(a.height, a.width, 4)
            y0 = (a.height // 2) // 8 * 8
            nodes, root = svo.nodes(copy=False), svo.GetRoot()
            o, _ = oracle_frame(nodes, root, a, f, cores, rows=(y0, y0 + 32))
            parity = {"delivered_frame_identical_to_1gpu": bool((got == mine).all()),
                      "delivered_band_identical_to_oracle": bool((got[y0:y0 + 32] == o["rgba"][y0:y0 + 32]).all()),
                      "sample": "frame %d (%s, last step): whole frame vs GPU 0's own render; rows %d..%d vs the oracle"
                                % (f, "all ranks' blocks" if tiles_mode else "rank %d" % last_rank, y0, y0 + 32)}
            del nodes
        del one_fb
    except Exception as ex:
        err = {"error": "%s: %s" % (type(ex).__name__, ex)}
        batch_1gpu = batch_1gpu or err
        parity = parity or err

    if not a.no_strong and not tiles_mode:
        spent = time.time() - T_START
        if spent > a.budget_s:
            strong = {"skipped": "budget: %.0f s spent of --budget-s %.0f" % (spent, a.budget_s)}
        else:
            try:
                flush = None
                torch.cuda.empty_cache()
                strong = run_strong_8k(yv, torch, a, world)
            except Exception as ex:
                strong = {"error": "%s: %s" % (type(ex).__name__, ex)}
    dist.barrier(group=idle)

    peak, peak_src = hbm_peak()
    alg_bytes_all = vis_all * NODE_BYTES + px_all * PIXEL_BYTES            # all GPUs, whole timed region
    achieved = alg_bytes_all / world / total_s / 1e9                       # per GPU
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": None, "l2_traffic": None, "traffic_source": None, "peak_source": peak_src,
                "kernel": "yv::render_frame<%s>" % schedule,
                "algorithmic_bytes_per_launch": alg_bytes_all / world / a.steps,
                "node_visits_per_ray": vis_all / rays_all, "pop_refetches_per_ray": pop_all / rays_all,
                "kernel_ms": 1e3 * total_s / a.steps}
    if tiles_mode:
        part = ("interleaved %d-row blocks of one frame" % a.band_rows) if a.partition == "tiles" else "contiguous row bands of one frame"
    elif a.flythrough:
        part = "one frame per GPU per step: frame step*N + rank of the flythrough"
    else:
        part = "one frame per GPU per step, every rank the base camera"
    line = {
        "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": a.steps, "warmup": warm,
        "ms_per_step": 1e3 * total_s / a.steps, "higher_is_better": True,
        "scaling": "strong" if tiles_mode else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_of(a),
        "setup": {"schedule": schedule, "smem_nodes": r.GetOption("smem_nodes"), "stack": r.GetOption("stack"),
                  "l2": "flushed between frames (256 MiB write, untimed)" if not a.no_l2_flush
                        else "not flushed; node pool %d MB > L2" % (dev_bytes >> 20),
                  "nodes": svo.nodecount, "packed_bytes": dev_bytes, "scene_build_s": round(build_s, 2),
                  "partition": part, "gather": gather, "flythrough": bool(a.flythrough), "hit_fraction": round(hit_frac, 4)},
        "frame_ms": 1e3 * total_s / a.steps, "rays_per_step": rays_step,
        "roofline": roofline, "cpu_baseline": None, "e2e": e2e, "clocks": clocks,
        "gpu_launches": a.steps * world * launches_per_step, "e2e_gpu_launches_per_step": r.LastFrameLaunches(),
        "parity": parity, "batch_1gpu": batch_1gpu, "strong_8k": strong, "wall_s": round(time.time() - T_START, 1),
    }
    emit(line)
    dist.barrier()
    dist.destroy_process_group()


_JSON_OUT = None


def emit(line):
    """The one JSON line, on the process's original stdout."""
    out = _JSON_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    # stdout carries the one JSON line and nothing else: whatever a library prints to fd 1 (NCCL's version banner when the
    # box sets NCCL_DEBUG, a loader message) goes to stderr; the line itself is written to a private copy of the original fd
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    a = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if a.impl == "reference":
        return run_reference(a, rank)
    if a.strong_only:
        return run_strong_only(a)
    if world > 1:
        return run_ranks(a)
    return run_single(a)


if __name__ == "__main__":
    main()
