"""Seeded random node pools: arbitrary mixes of empty / full / leaf / node children, ragged depths, shared sub-trees
(a DAG is a legal pool: the reference only follows child ids, cell/ppu_renderer.cpp:35), garbage in unused fields.
Every pool goes through yv_svo_from_memory and is traced from random cameras; the kernel's per-ray code must agree with
the oracle bit for bit. CPU: host build of trace_core.cuh (tests/emu); GPU: the CUDA path, packed and raw layouts."""
import numpy as np
import pytest

import yve
import yvo
import yoxel_voxel_b200 as yv


def _random_pool(seed, max_depth, p_node, p_leaf, p_full, share):
    """Bottom-up: nodes of depth d may point at any already-built node of depth < d (sharing when share > 0)."""
    rng = np.random.RandomState(seed)
    nodes = []
    by_depth = {}                                        # remaining height -> node ids

    def make(height):
        n = np.zeros((), yv.NODE_DTYPE)
        n["data"] = rng.randint(0, 2 ** 32, dtype=np.uint64).astype(np.uint32)
        flags = 0
        for c in range(8):
            u = rng.rand()
            if height > 0 and u < p_node:
                pool = [i for h in range(height) for i in by_depth.get(h, [])]
                if pool and rng.rand() < share:
                    child = pool[rng.randint(len(pool))]              # shared sub-tree
                else:
                    child = make(height - 1 if rng.rand() < 0.8 else rng.randint(0, height))       # ragged depths
                n["child"][c] = child
            elif u < p_node + p_leaf:
                n["child"][c] = rng.randint(1, 2 ** 32, dtype=np.uint64).astype(np.uint32)   # inline VoxData, any bits
                flags |= 1 << c
            elif u < p_node + p_leaf + p_full:
                n["child"][c] = yv.FULL_NODE
            else:
                n["child"][c] = yv.EMPTY_NODE
        n["flags"] = flags | (int(rng.randint(0, 2 ** 16)) << 16)      # junk in the unused high half
        nodes.append(n)
        by_depth.setdefault(height, []).append(len(nodes) - 1)
        return len(nodes) - 1

    root = make(max_depth)
    return root, np.array(nodes, yv.NODE_DTYPE)


POOLS = [  # seed, max_depth, p_node, p_leaf, p_full, share
    (2, 4, 0.55, 0.20, 0.10, 0.0),
    (2, 6, 0.40, 0.15, 0.05, 0.3),
    (2, 9, 0.38, 0.10, 0.10, 0.3),
    (4, 3, 0.80, 0.15, 0.00, 0.0),
    (2, 12, 0.33, 0.06, 0.02, 0.4),
    (6, 1, 0.00, 0.50, 0.25, 0.0),
    (7, 7, 0.40, 0.02, 0.30, 0.2),
]


def _cams(seed, n):
    rng = np.random.RandomState(seed)
    out = []
    for i in range(n):
        pos = rng.uniform(-0.8, 1.8, 3) if i % 2 == 0 else rng.uniform(0.02, 0.98, 3)
        d = rng.uniform(0.2, 0.8, 3) - pos + rng.normal(0, 0.15, 3)
        if np.linalg.norm(d) < 1e-2:
            d = np.array([0.3, -1.0, 0.2])
        up = (0.0, 0.0, 1.0) if abs(d[2]) < 0.9 * np.linalg.norm(d) else (0.0, 1.0, 0.0)
        out.append((tuple(float(np.float32(v)) for v in pos), tuple(float(np.float32(v)) for v in d), up,
                    float(rng.choice([25.0, 70.0, 120.0]))))
    return out


@pytest.mark.parametrize("spec", POOLS, ids=lambda s: "seed%d_d%d_n%g" % (s[0], s[1], s[2]))
def test_emu_matches_oracle_on_random_pools(spec):
    root, nodes = _random_pool(*spec)
    svo = yv.SVOData.FromNodes(root, nodes)
    pool = svo.nodes()                                  # flags normalised by the loader
    recs, leaves = svo.packed()
    node_data = pool["data"][recs[:, 3]]
    hits = 0
    for i, (pos, d, up, fov) in enumerate(_cams(spec[0] * 17 + spec[1], 6)):
        W, H = [(48, 40), (33, 29)][i % 2]
        detail = [0.0, 4.0][i % 2]
        o = yvo.render(pool, svo.GetRoot(), yvo.camera(pos, d, up, fov, W, H, detail_coef=detail))
        d0, du, dv = yv.init_ray_dir(d, up, fov, W, H)
        half_rad = np.float32(np.float32(fov) / np.float32(2)) * np.float32(np.pi / 180.0)
        det = float(np.float32(np.float32(detail) * half_rad) / np.float32(W))
        e = yve.render(recs, leaves, 1, pos, d0, du, dv, pos, W, H, detail=det, node_data=node_data)
        assert (o["node"] == e["node"]).all() and (o["child"] == e["child"]).all(), (spec, i)
        assert o["t"].tobytes() == e["t"].tobytes() and (o["rgba"] == e["rgba"]).all(), (spec, i)
        hits += int((o["node"] != yvo.MISS_NODE).sum())
    assert hits > 0


@pytest.mark.parametrize("spec", POOLS, ids=lambda s: "seed%d_d%d_n%g" % (s[0], s[1], s[2]))
def test_emu_closed_form_descent_on_random_pools(spec):
    """lean_descend_once (trace_core.cuh) in front of primary and secondary rays on ragged pools, DAGs and pools full of
    leaves — eyes and secondary origins inside the cube, on its faces, on cell boundaries: pixels, ids, t bits equal to the
    oracle's, node fetches / pop re-fetches / stack depth equal to the general loop's, with and without the LOD cut-off."""
    root, nodes = _random_pool(*spec)
    svo = yv.SVOData.FromNodes(root, nodes)
    pool = svo.nodes()
    recs, leaves = svo.packed()
    node_data = pool["data"][recs[:, 3]]
    cams = _cams(spec[0] * 31 + spec[1], 6)
    # eyes on cell boundaries and cube faces: the closed form must hand exactly these cases back to the general loop
    cams += [((0.5, 0.5, 0.5), (0.3, -0.7, 0.2), (0.0, 0.0, 1.0), 70.0), ((0.25, 0.75, 0.5), (-0.4, 0.1, 0.9), (0.0, 1.0, 0.0), 90.0),
             ((0.0, 0.5, 0.5), (1.0, 0.1, 0.05), (0.0, 0.0, 1.0), 70.0), ((1.0, 1.0, 1.0), (-1.0, -0.9, -0.8), (0.0, 0.0, 1.0), 50.0)]
    fast_levels = 0
    for i, (pos, d, up, fov) in enumerate(cams):
        W, H = [(40, 32), (29, 23)][i % 2]
        detail = [0.0, 4.0][i % 2]
        d0, du, dv = yv.init_ray_dir(d, up, fov, W, H)
        half_rad = np.float32(np.float32(fov) / np.float32(2)) * np.float32(np.pi / 180.0)
        det = float(np.float32(np.float32(detail) * half_rad) / np.float32(W))
        o = yvo.render(pool, svo.GetRoot(), yvo.camera(pos, d, up, fov, W, H, detail_coef=detail))
        plain = yve.render(recs, leaves, 1, pos, d0, du, dv, pos, W, H, detail=det, node_data=node_data, mode=2)
        fast = yve.render(recs, leaves, 1, pos, d0, du, dv, pos, W, H, detail=det, node_data=node_data, mode=5)
        for e in (plain, fast):
            assert (o["node"] == e["node"]).all() and (o["child"] == e["child"]).all(), (spec, i)
            assert o["t"].tobytes() == e["t"].tobytes() and (o["rgba"] == e["rgba"]).all(), (spec, i)
        assert (fast["visits"], fast["fetches"], fast["max_sp"]) == (plain["visits"], plain["fetches"], plain["max_sp"]), (spec, i)
        fast_levels += fast["fast_levels"]
        # secondary rays from the same camera (no LOD): shadow + AO, origins a voxel off the surface, on it, and far off it
        for vs, amax in ((2.0 ** -spec[1], 0.05), (0.0, 0.3), (0.11, 0.0)):
            kw = dict(shadow=1, ao_samples=2, seed=3 + i, voxel_size=vs, ao_max_t=amax)
            light = (0.6, 0.4, 1.2)
            os_ = yvo.render(pool, svo.GetRoot(), yvo.camera(pos, d, up, fov, W, H),
                             sec=yvo.secondary(light_pos=light, **kw))
            p2 = yve.render(recs, leaves, 1, pos, d0, du, dv, light, W, H, mode=2, **kw)
            f2 = yve.render(recs, leaves, 1, pos, d0, du, dv, light, W, H, mode=5, **kw)
            assert (os_["rgba"] == p2["rgba"]).all() and (os_["rgba"] == f2["rgba"]).all(), (spec, i, vs)
            assert (f2["visits"], f2["fetches"], f2["max_sp"]) == (p2["visits"], p2["fetches"], p2["max_sp"]), (spec, i, vs)
            fast_levels += f2["fast_levels"]
    if spec[1] >= 4:
        assert fast_levels > 0


@pytest.mark.gpu
def test_cuda_matches_oracle_on_random_pools():
    r = yv.SVORenderer(0)
    r.EnableHits(True)
    total = 0
    try:
        for spec in POOLS:
            root, nodes = _random_pool(*spec)
            svo = yv.SVOData.FromNodes(root, nodes)
            pool = svo.nodes()
            r.SetScene(svo)
            for i, (pos, d, up, fov) in enumerate(_cams(spec[0] * 31 + spec[1], 8)):
                W, H = [(128, 96), (67, 45)][i % 2]
                detail = [0.0, 4.0][(i // 2) % 2]
                r.SetOption("layout", (i // 4) % 2)                       # packed records / raw reference pool
                r.SetOption("schedule", i % 3)
                r.SetResolution(W, H)
                r.SetViewPos(pos); r.SetViewDir(d); r.SetViewUp(up); r.SetFOV(fov); r.SetDetailCoef(detail)
                img = r.RenderFrame().copy()
                node, child, t = r.GetHits()
                o = yvo.render(pool, svo.GetRoot(), yvo.camera(pos, d, up, fov, W, H, detail_coef=detail), threads=4)
                assert (node == o["node"]).all() and (child == o["child"]).all(), (spec, i)
                assert t.tobytes() == o["t"].tobytes() and (img == o["rgba"]).all(), (spec, i)
                total += int((node != yvo.MISS_NODE).sum())
    finally:
        r.close()
    assert total > 10000
