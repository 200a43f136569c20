#!/usr/bin/env python
"""Turn the output of tools/gpu_r2_all.sh (gpurun_out/) into the tracked evidence under profiles/:

    python tools/collect_profiles.py r02

  profiles/<tag>_bench_*.json         the bench lines as printed (N=1 headline + configs, reference arm, SSNA, 8K)
  profiles/<tag>_launches.csv         ncu launch list of the default bench command
  profiles/<tag>_prof_{primary,sec,ssna}.summary.txt / .lines.txt    from the `ncu --set full` captures
  profiles/traffic.json               DRAM / L2 bytes per launch of the primary kernel, with the git commit and the sha of
                                      the kernel sources the capture was taken on (bench.py reports both beside the
                                      sha of the sources it runs, so a stale capture shows)
Run in the container right after the gpurun call, before the kernel sources change."""
import csv
import io
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")


def sh(*cmd):
    return subprocess.run(list(cmd), capture_output=True, text=True, cwd=ROOT).stdout


def raw_page(rep):
    rows = list(csv.reader(io.StringIO(sh("ncu", "-i", rep, "--page", "raw", "--csv"))))
    return rows[0], rows[1], rows[2:]


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
    # the commit the GPU run was made on (default HEAD): the kernel-source sha is taken from that commit's blobs
    rev = sys.argv[2] if len(sys.argv) > 2 else "HEAD"
    commit = sh("git", "rev-parse", "--short=12", rev).strip()
    import hashlib
    h = hashlib.sha1()
    for f in ("trace_core.cuh", "render_kernels.cuh"):
        h.update(subprocess.run(["git", "show", "%s:yoxel-voxel_b200/csrc/%s" % (commit, f)], capture_output=True, cwd=ROOT).stdout)
    ksha = h.hexdigest()[:16]
    for name in ("bench_n1", "bench_ref", "bench_ssna", "bench_8k"):
        src = os.path.join(OUT, name + ".json")
        if os.path.exists(src) and os.path.getsize(src) > 0:
            shutil.copy(src, os.path.join(PROF, "%s_%s.json" % (tag, name)))
    if os.path.exists(os.path.join(OUT, "launches.csv")):
        shutil.copy(os.path.join(OUT, "launches.csv"), os.path.join(PROF, "%s_launches.csv" % tag))
    if os.path.exists(os.path.join(OUT, "pytest_gpu.log")):
        shutil.copy(os.path.join(OUT, "pytest_gpu.log"), os.path.join(PROF, "%s_pytest_gpu.log" % tag))
    for name in ("prof_primary", "prof_sec", "prof_ssna"):
        rep = os.path.join(OUT, name + ".ncu-rep")
        if not os.path.exists(rep):
            continue
        head = "# %s.ncu-rep, captured on commit %s, kernel sources sha %s\n" % (name, commit, ksha)
        open(os.path.join(PROF, "%s_%s.summary.txt" % (tag, name)), "w").write(
            head + sh(sys.executable, "tools/ncu_summary.py", rep))
        src_csv = os.path.join(OUT, name + ".source.csv")
        open(src_csv, "w").write(sh("ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"))
        open(os.path.join(PROF, "%s_%s.lines.txt" % (tag, name)), "w").write(
            head + sh(sys.executable, "tools/ncu_lines.py", src_csv))
    rep = os.path.join(OUT, "prof_primary.ncu-rep")
    if os.path.exists(rep):
        hdr, units, rows = raw_page(rep)
        r = rows[0]
        g = lambda k: float(r[hdr.index(k)])
        u = lambda k: units[hdr.index(k)]
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        rd = g("dram__bytes_read.sum") * scale[u("dram__bytes_read.sum")]
        wr = g("dram__bytes_write.sum") * scale[u("dram__bytes_write.sum")]
        l2r, l2w = g("lts__t_sectors_srcunit_tex_op_read.sum"), g("lts__t_sectors_srcunit_tex_op_write.sum")
        rays = 1920 * 1080
        tj = {"kernel": r[hdr.index("Kernel Name")], "commit": commit, "kernel_source_sha": ksha,
              "source": "profiles/%s_prof_primary.summary.txt (ncu --set full --clock-control none, one launch, config 2, "
                        "tools/gpu_r2_all.sh)" % tag,
              "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_launch": rd + wr,
              "l2_read_sectors": int(l2r), "l2_write_sectors": int(l2w), "l2_bytes_per_launch": int((l2r + l2w) * 32),
              "rays_per_launch": rays, "dram_bytes_per_ray": round((rd + wr) / rays, 2),
              "l2_bytes_per_ray": round((l2r + l2w) * 32 / rays, 1), "algorithmic_bytes_per_ray": 1400.7,
              "note": "lts sectors x 32 B; the L2 write traffic is the explicit stack (STL.128 write-through), the node "
                      "records are served by L1"}
        json.dump(tj, open(os.path.join(PROF, "traffic.json"), "w"), indent=1)
        print("traffic.json:", tj["dram_bytes_per_launch"], tj["l2_bytes_per_launch"], commit, ksha)
    print("collected into profiles/ with tag", tag)


if __name__ == "__main__":
    main()
