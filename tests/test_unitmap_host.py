"""CPU test of the tile engine's work decomposition (csrc/unitmap.cuh, compiled for the host): every tile of the upper
triangle is produced exactly once, units are contiguous strips of one tile column, the dynamic queue order is
longest-processing-time, the per-column descriptor agrees with the unit map and with the slot formulas of phase A."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TB = 64


@pytest.fixture(scope="module")
def um():
    out = os.path.join(ROOT, "tests", "csrc", "_unitmap_host.so")
    src = os.path.join(ROOT, "tests", "csrc", "unitmap_host.cpp")
    hdr = os.path.join(ROOT, "eigensolver_gpu_b200", "csrc")
    subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-x", "c++", "-std=c++17", "-I", hdr, src, "-o", out])
    return C.CDLL(out)


def _enumerate(lib, n, G, P, rank, upc=3):
    tn = (n + TB - 1) // TB
    cap = tn * (tn + 1) // 2 + 8
    buf = np.zeros(4 * cap, dtype=np.int32)
    info = np.zeros(3, dtype=np.int32)
    cnt = lib.unitmap_enumerate(n, G, P, rank, upc, buf.ctypes.data_as(C.c_void_p), cap, info.ctypes.data_as(C.c_void_p))
    assert cnt >= 0
    return buf[:4 * cnt].reshape(cnt, 4), int(info[0]), int(info[1]), int(info[2])


@pytest.mark.parametrize("n", [1, 63, 64, 65, 200, 1000, 2049, 4096, 8192, 12345, 32768])
@pytest.mark.parametrize("P", [1, 2, 4, 8])
def test_every_tile_exactly_once(um, n, P):
    tn = (n + TB - 1) // TB
    seen = np.zeros((tn, tn), dtype=np.int32)
    for rank in range(P):
        tiles, c, total, nf = _enumerate(um, n, 148, P, rank)
        assert 1 <= c <= 8 or (tn - 1) // c <= 256
        for u, i, j, d in tiles:
            assert 0 <= i <= j < tn and j % P == rank          # only owned tile columns
            assert (d == 1) == (i == j)
            seen[i, j] += 1
        # units are contiguous column strips inside one band; a diagonal tile closes its unit
        for u in np.unique(tiles[:, 0]):
            t = tiles[tiles[:, 0] == u]
            assert len(set(t[:, 2])) == 1
            rows = t[t[:, 3] == 0][:, 1]
            if len(rows):
                assert np.array_equal(rows, np.arange(rows[0], rows[0] + len(rows))) and rows[0] // c == rows[-1] // c
            assert t[:, 3].sum() <= 1 and (t[:, 3].sum() == 0 or t[-1, 3] == 1)
            assert len(t) <= c + 1
    iu = np.triu_indices(tn)
    assert np.all(seen[iu] == 1) and seen.sum() == tn * (tn + 1) // 2


@pytest.mark.parametrize("n", [700, 3000, 8192, 16384])
def test_queue_order_is_longest_first(um, n):
    tiles, c, total, nf = _enumerate(um, n, 148, 1, 0)
    sizes = np.bincount(tiles[:, 0], minlength=total)
    assert np.all(sizes[:nf] == c)                              # F units: full strips
    d = sizes[nf:]
    assert np.all(d[:-1] >= d[1:]) and d.max() <= c and d.min() >= 1      # D units by decreasing size


@pytest.mark.parametrize("j", [1, 2, 64, 65, 777, 4095, 4096, 8191, 20000])
@pytest.mark.parametrize("G", [148, 132])
def test_descriptor_matches_unit_map_and_slot_formulas(um, j, G):
    out = np.zeros(10, dtype=np.int32)
    um.unitmap_desc(j, G, 1, 3, out.ctypes.data_as(C.c_void_p))
    jj, tn, c, rcp, kb, nf, total, r, ndj, nsj = [int(x) for x in out]
    tiles, c2, total2, nf2 = _enumerate(um, j, G, 1, 0)
    assert (jj, tn, c, total, nf) == (j, (j + TB - 1) // TB, c2, total2, nf2)
    assert all(((i * rcp) >> 16) == i // c for i in range(0, 4096))       # reciprocal division used on the device
    assert r % 8 == 0 and r * G >= j and (r - 8) * G < j                  # rows per CTA of the following phase A
    # partial-sum slots of the last row (row j-1): direct slots J > I, band slots k <= I / C
    i_last = (j - 1) // TB
    assert ndj == tn - (i_last + 1) and nsj == ndj + i_last // c + 1
    # the slots phase A reads for ANY row are exactly the ones the units write: direct (I, J) for J > I, band I / C
    written_direct = {(int(i), int(jc)) for u, i, jc, d in tiles if d == 0}
    assert written_direct == {(i, jc) for jc in range(tn) for i in range(jc)}


def _replay(tn, c, rank, p, g, ov):
    """list scheduling of the unit queue (the model behind tile::build_strip_table), in tile times"""
    owned = list(range(rank, tn, p))
    kb = (tn - 1) // c
    nf = sum(1 for k in range(kb) for j in owned if j >= (k + 1) * c)
    fin = [(nf // g + (1 if i < nf % g else 0)) * (c + ov) for i in range(g)]
    sizes = [j % c + 1 for j in owned]
    if p == 1:
        sizes.sort(reverse=True)
    for s in sizes:
        i = fin.index(min(fin))
        fin[i] += s + ov
    return max(fin)


@pytest.mark.parametrize("g,p,ov", [(148, 1, 0.5), (148, 2, 0.5), (148, 4, 1.0), (148, 8, 0.5), (7, 3, 0.5)])
def test_strip_table_minimises_the_replayed_makespan(um, g, p, ov):
    tnmax, cmax = 140, 8
    tab = np.zeros(tnmax + 1, dtype=np.uint8)
    um.unitmap_strip_table(g, p, tnmax, cmax, C.c_double(ov), tab.ctypes.data_as(C.c_void_p))
    assert tab[0] == 1 and tab.min() >= 1 and tab.max() <= cmax
    for tn in (1, 2, 5, 17, 64, 96, 124, 140):
        cost = {c: max(_replay(tn, c, r, p, g, ov) for r in range(p)) for c in range(1, cmax + 1)}
        best = min(cost.values())
        assert abs(cost[int(tab[tn])] - best) < 1e-9, (tn, tab[tn], cost)
        assert (tn - 1) // int(tab[tn]) <= 256
