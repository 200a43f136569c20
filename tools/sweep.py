#!/usr/bin/env python
"""Kernel-variant sweep on the GPU box (run through gpurun): kernel ms per frame for combinations of the
renderer's ablation knobs on one workload, L2 flushed between frames, every variant's image compared
with the first one's. Prints one line per variant and writes gpurun_out/sweep.json."""
import argparse
import itertools
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--depth", type=int, default=12)
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--frames", type=int, default=12)
    ap.add_argument("--persistent", default="0,1")
    ap.add_argument("--stack", default="0")
    ap.add_argument("--smem", default="0")
    ap.add_argument("--refill", default="20")
    ap.add_argument("--layout", default="0")
    ap.add_argument("--sec-threshold", default="-1")
    ap.add_argument("--sec-queue", default="0")
    ap.add_argument("--detail", type=float, default=0.0)
    ap.add_argument("--secondary", action="store_true")
    ap.add_argument("--scene", default="fractal", choices=["fractal", "iso"])
    ap.add_argument("--pos", default="0.5,0.5,0.3")
    ap.add_argument("--dir", default="-1,-1,1.5")
    ap.add_argument("--out", default="gpurun_out/sweep.json")
    a = ap.parse_args()
    import torch
    import yoxel_voxel_b200 as yv
    svo = yv.SVOData.SphereFractal(a.depth) if a.scene == "fractal" else yv.SVOData.IsoVolume(a.depth)
    svo.Upload(0)
    r = yv.SVORenderer(0)
    st = torch.cuda.Stream()
    r.SetStream(st.cuda_stream)
    r.SetScene(svo)
    r.SetResolution(a.width, a.height)
    r.SetViewPos([float(v) for v in a.pos.split(",")]); r.SetViewDir([float(v) for v in a.dir.split(",")])
    if a.secondary:
        r.SetSecondary(shadow=1, ao_samples=4, seed=1, light_pos=(0.6, 0.4, 1.2), voxel_size=1.0 / (1 << a.depth), ao_max_t=0.05)
    buf = torch.zeros(a.height, a.width, 4, dtype=torch.uint8, device="cuda:0")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda:0")
    ref, results = None, []
    combos = []
    for persistent, stack, smem in itertools.product([int(v) for v in a.persistent.split(",")], [int(v) for v in a.stack.split(",")],
                                                     [int(v) for v in a.smem.split(",")]):
        for refill in ([int(v) for v in a.refill.split(",")] if persistent == 1 else [20]):
            for layout in [int(v) for v in a.layout.split(",")]:
                for sect in [int(v) for v in a.sec_threshold.split(",")]:
                    for secq in [int(v) for v in a.sec_queue.split(",")]:
                        combos.append((persistent, stack, smem, refill, layout, sect, secq))
    with torch.cuda.stream(st):
        for persistent, stack, smem, refill, layout, sect, secq in combos:
            r.SetOption("persistent", persistent); r.SetOption("stack", stack); r.SetOption("smem_nodes", smem)
            r.SetOption("refill", refill); r.SetOption("layout", layout); r.SetDetailCoef(a.detail)
            r.SetOption("sec_threshold", sect); r.SetOption("sec_queue", secq)
            try:
                times = []
                for i in range(a.frames + 3):
                    flush.zero_()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(st)
                    r.Render(buf.data_ptr(), sync=False)
                    e1.record(st)
                    st.synchronize()
                    if i >= 3:
                        times.append(e0.elapsed_time(e1))
                img = buf.cpu().numpy()
                if ref is None:
                    ref = img.copy()
                same = bool((img == ref).all())
                # warm-L2 time as well (no flush)
                warm = []
                for i in range(6):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(st); r.Render(buf.data_ptr(), sync=False); e1.record(st); st.synchronize()
                    warm.append(e0.elapsed_time(e1))
                res = dict(lib=os.environ.get("YV_B200_LIB", "default"), persistent=persistent, stack=stack, smem=smem, refill=refill, layout=layout, sec_threshold=sect, sec_queue=secq, ms_median=float(np.median(times)),
                           ms_min=float(min(times)), ms_warm_l2=float(np.median(warm[2:])), same_image=same)
            except yv.YVError as e:
                res = dict(persistent=persistent, stack=stack, smem=smem, refill=refill, error=str(e))
            results.append(res)
            print(json.dumps(res), flush=True)
    os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
    json.dump(dict(args=vars(a), results=results), open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
