#!/usr/bin/env python
"""Ceiling of "N GPUs deliver pixels into one host frame": aggregate device->host bandwidth of k = 1, 2, 4, .. N GPUs
copying concurrently (copy engines, cudaMemcpyAsync from HBM into one pinned host buffer, each GPU its own slice), and
the same for one GPU alone. bench.py's N-GPU `e2e` cannot beat frame_bytes / this. Writes gpurun_out/host_ingest.json."""
import json
import sys
import time

import torch


def main():
    n = torch.cuda.device_count()
    mb = 128
    size = mb << 20
    host = torch.empty(n * size, dtype=torch.uint8).pin_memory()
    src = [torch.zeros(size, dtype=torch.uint8, device="cuda:%d" % k) for k in range(n)]
    streams = [torch.cuda.Stream(device=k) for k in range(n)]
    out = {"gpus": n, "mb_per_gpu_per_copy": mb, "rows": []}
    ks = [k for k in (1, 2, 4, 8) if k <= n]
    for k in ks:
        best = 0.0
        for rep in range(5):
            for d in range(k):
                torch.cuda.synchronize(d)
            t0 = time.perf_counter()
            for it in range(4):
                for d in range(k):
                    with torch.cuda.stream(streams[d]):
                        host[d * size:(d + 1) * size].copy_(src[d], non_blocking=True)
            for d in range(k):
                streams[d].synchronize()
            dt = time.perf_counter() - t0
            best = max(best, 4 * k * size / dt / 1e9)
        out["rows"].append({"gpus_copying": k, "aggregate_gb_per_s": round(best, 1), "per_gpu_gb_per_s": round(best / k, 1)})
        print(out["rows"][-1])
    json.dump(out, open("gpurun_out/host_ingest.json", "w"), indent=1)


if __name__ == "__main__":
    sys.exit(main())
