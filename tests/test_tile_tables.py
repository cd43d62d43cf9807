"""Host logic of the DMMA band solver (no GPU): the tile ownership tables compiled into csrc/ba_solve_mma.cu obey the
rules the kernel relies on, and the 12-warp table is what tools/gen_tile_tables.py generates."""
import importlib.util
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _table(name):
    src = open(os.path.join(ROOT, "batrack_b200", "csrc", "ba_solve_mma.cu")).read()
    m = re.search(r"__constant__ unsigned char %s\[(\d+)\] = \{([^}]*)\}" % name, src)
    vals = [int(v) for v in m.group(2).replace("\n", " ").split(",") if v.strip()]
    assert len(vals) == int(m.group(1))
    return vals


def _check(px, py, nw):
    tiles = len(px) // nw
    seen = set()
    for w in range(nw):
        deg = [0] * 16
        for t in range(tiles):
            x, y = px[t * nw + w], py[t * nw + w]
            if x == 255:
                assert y == 255
                continue
            assert 0 <= x <= y <= 15
            assert (x, y) not in seen                     # every pair exactly once
            seen.add((x, y))
            for e in {x, y}:
                deg[e] += 1
        assert max(deg) <= 2                               # two e-tile slots per warp suffice
    assert len(seen) == 136
    for x in range(16):                                    # the look-ahead owner is found without a table
        assert (px[(x // nw) * nw + x % nw], py[(x // nw) * nw + x % nw]) == (x, x)


def test_tile_tables_8_and_12_warps():
    _check(_table("c_px"), _table("c_py"), 8)
    _check(_table("c_px12"), _table("c_py12"), 12)


def test_generator_reproduces_the_compiled_table():
    spec = importlib.util.spec_from_file_location("gen_tile_tables", os.path.join(ROOT, "tools", "gen_tile_tables.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    cap, tiles = gen.build(12)
    px, py = [255] * (cap * 12), [255] * (cap * 12)
    for w in range(12):
        for t, (x, y) in enumerate(tiles[w]):
            px[t * 12 + w], py[t * 12 + w] = x, y
    assert px == _table("c_px12") and py == _table("c_py12")
