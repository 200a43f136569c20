"""Small renders through every kernel variant, for compute-sanitizer (memcheck / racecheck) runs."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import yoxel_voxel_b200 as yv

svo = yv.SVOData.SphereFractal(9)
r = yv.SVORenderer(0)
r.EnableHits(True)
r.SetScene(svo)
r.SetResolution(203, 131)
r.SetViewPos((0.5, 0.5, 0.3)); r.SetViewDir((-1, -1, 1.5))
ref = None
for sec in (False, True):
    r.SetSecondary(1, 4, 1, (0.6, 0.4, 1.2), 1.0 / 512, 0.05) if sec else r.SetSecondary(0, 0)
    for schedule in (0, 1, 2):
        for stack in (0, 4):
            for smem in (0, 73):
                for layout in (0, 1):
                    for detail in (0.0, 6.0):
                        for secq in ((0, 1) if sec else (0,)):
                            r.SetOption("schedule", schedule); r.SetOption("stack", stack); r.SetOption("smem_nodes", smem)
                            r.SetOption("layout", layout); r.SetDetailCoef(detail); r.SetOption("sec_queue", secq)
                            img = r.RenderFrame().copy()
                            key = (sec, detail)
                            ref = ref or {}
                            if key in ref:
                                assert (img == ref[key]).all(), (sec, schedule, stack, smem, layout, detail, secq)
                            else:
                                ref[key] = img
r.SetSecondary(0, 0); r.SetDetailCoef(0); r.SetOption("layout", 0); r.SetOption("schedule", 0)
r.SetLigth(0, yv.LightParams(True, (0.5, 0.5, 0.3)))
r.RenderFrame()
pos = np.random.RandomState(0).rand(500, 3).astype(np.float32)
d = np.random.RandomState(1).randn(500, 3).astype(np.float32)
r.TraceRays(pos, d)
svo.BuildRange(9, (256, 256, 300), yv.BuildMode.CLEAR, yv.MakeSphereSource(20, (200, 180, 120), True))
r.SetOption("layout", 1)
r.RenderFrame()
print("sanitize run ok: %d variants" % len(ref))
