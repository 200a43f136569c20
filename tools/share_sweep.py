#!/usr/bin/env python
"""Why does one GPU's share of an 8K frame take longer than 1/N of the whole frame? (strong_8k at N=8: 0.98 ms per member
against 6.49 / 8 = 0.81 ms.) One GPU renders the share member 0 of an N-GPU group would get — every block size, and the
contiguous band — and the whole frame, L2 flushed or warm, on a few flythrough cameras. Run through gpurun on one B200."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--depth", type=int, default=13)
    ap.add_argument("--n", type=int, default=8)
    ap.add_argument("--frames", default="0,21,42,63")
    ap.add_argument("--out", default="gpurun_out/share_sweep.json")
    a = ap.parse_args()
    import torch
    import yoxel_voxel_b200 as yv
    W, H = 7680, 4320
    svo = yv.SVOData.IsoVolume(a.depth, seed=219, iso_level=200, threads=os.cpu_count() or 8)
    svo.Upload(0)
    r = yv.SVORenderer(0)
    r.SetScene(svo); r.SetResolution(W, H); r.SetViewUp(bench.UP); r.SetFOV(bench.FOV)
    fb = torch.zeros(H, W, 4, dtype=torch.uint8, device="cuda:0")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda:0")
    frames = [int(f) for f in a.frames.split(",")]
    cams = [bench.camera_for(f, "iso", 64) for f in frames]

    def timed(setup, flush_l2):
        tot = 0.0
        for pos, d in cams:
            r.SetViewPos(pos); r.SetViewDir(d)
            setup()
            r.Render(fb.data_ptr(), sync=True)             # warm-up of this camera (also warms L2 for the warm case)
            best = 1e9
            for _ in range(3):
                if flush_l2:
                    flush.zero_(); torch.cuda.synchronize()
                r.Render(fb.data_ptr(), sync=True)
                best = min(best, r.LastFrameMs())
            tot += best
        return tot / len(cams)

    out = {"depth": a.depth, "n": a.n, "frames": frames, "rows": []}
    for flush_l2 in (True, False):
        full = timed(lambda: r.SetRows(0, H), flush_l2)
        row = {"l2": "flushed" if flush_l2 else "warm", "full_ms": full, "ideal_share_ms": full / a.n, "share_ms": {}}
        for rows in (16, 32, 64, 128, 256):
            shares = []
            for phase in (0, a.n // 2, a.n - 1):
                shares.append(timed(lambda: r.SetInterleave(rows, a.n, phase), flush_l2))
            row["share_ms"]["interleaved_%d" % rows] = [round(s, 4) for s in shares]
        per = ((H + a.n - 1) // a.n + 7) // 8 * 8
        r.SetInterleave(32, 1, 0)
        bands = []
        for k in (0, a.n // 2, a.n - 1):
            bands.append(timed(lambda: r.SetRows(k * per, min(H, (k + 1) * per)), flush_l2))
        row["share_ms"]["band"] = [round(s, 4) for s in bands]
        r.SetRows(0, H)
        out["rows"].append(row)
        print(json.dumps(row))
    json.dump(out, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
