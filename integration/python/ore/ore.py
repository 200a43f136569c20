"""ore.ore — what the reference's scripts import (`from ore.ore import *`: gen_spheres.py:2, gen_largevol.py:2,
scene_gen.py:3, qtview.py:10). The reference's own ore/ore.py is `from _ore import *` plus p3i / p3f (:6-12), written for
Python 2 (it indexes the result of map()); this is the same module for Python 3 over this repo's `_ore`."""
from ._ore import *          # noqa: F401,F403
from ._ore import point_3i, point_3f


def p3i(p):
    p = [int(v) for v in p]
    return point_3i(p[0], p[1], p[2])


def p3f(p):
    p = [float(v) for v in p]
    return point_3f(p[0], p[1], p[2])
