"""Self-consistency of the hand-written marching-cubes triangle table (no GPU, no oracle)."""
import itertools

from icon_b200.mc_tables import CORNERS, EDGE_CORNERS, TRI_LIST, EDGE_TABLE


def _edge_on_face(e, axis, side):
    a, b = EDGE_CORNERS[e]
    return CORNERS[a][axis] == side and CORNERS[b][axis] == side


def test_edges_cross_the_surface():
    for case in range(256):
        inside = [(case >> c) & 1 for c in range(8)]
        crossing = {e for e, (a, b) in enumerate(EDGE_CORNERS) if inside[a] != inside[b]}
        assert set(TRI_LIST[case]) == crossing, case
        assert EDGE_TABLE[case] == sum(1 << e for e in crossing)


def test_complement_uses_same_edges():
    for case in range(256):
        assert set(TRI_LIST[case]) == set(TRI_LIST[255 - case])


def test_no_degenerate_triangles():
    for case in range(256):
        t = TRI_LIST[case]
        for i in range(0, len(t), 3):
            assert len({t[i], t[i + 1], t[i + 2]}) == 3


def _face_segments(case, axis, side):
    """Directed triangle edges of `case` that lie in the cube face (axis, side)."""
    segs = []
    t = TRI_LIST[case]
    for i in range(0, len(t), 3):
        tri = t[i:i + 3]
        for k in range(3):
            a, b = tri[k], tri[(k + 1) % 3]
            if _edge_on_face(a, axis, side) and _edge_on_face(b, axis, side):
                segs.append((a, b))
    return segs


def _edge_key(e, axis):
    """Identify a cube edge lying in a face by its two corners' in-face coordinates."""
    a, b = EDGE_CORNERS[e]
    drop = lambda c: tuple(v for i, v in enumerate(CORNERS[c]) if i != axis)
    return frozenset((drop(a), drop(b)))


def _net(segs):
    """Cancel a->b against b->a (two triangles of one cube meeting along an in-face edge)."""
    out = []
    for s in segs:
        r = (s[1], s[0])
        if r in out:
            out.remove(r)
        else:
            out.append(s)
    return frozenset(out)


def test_shared_faces_match_between_neighbouring_cubes():
    """Crack-free surface: the net polyline a case draws on its +axis face must be exactly the
    reverse of what every case with the same four corner states draws on its -axis face."""
    for axis in range(3):
        hi = [c for c in range(8) if CORNERS[c][axis] == 1]
        lo = [c for c in range(8) if CORNERS[c][axis] == 0]
        pair = {h: next(l for l in lo if all(CORNERS[l][i] == CORNERS[h][i] for i in range(3) if i != axis))
                for h in hi}
        seen = {}
        for case in range(256):
            pat_hi = tuple((case >> h) & 1 for h in hi)
            s_hi = _net([(_edge_key(a, axis), _edge_key(b, axis)) for a, b in _face_segments(case, axis, 1)])
            seen.setdefault(("hi", pat_hi), set()).add(s_hi)
            pat_lo = tuple((case >> pair[h]) & 1 for h in hi)
            s_lo = _net([(_edge_key(b, axis), _edge_key(a, axis)) for a, b in _face_segments(case, axis, 0)])
            seen.setdefault(("lo", pat_lo), set()).add(s_lo)
        for pat in itertools.product((0, 1), repeat=4):
            a, b = seen[("hi", pat)], seen[("lo", pat)]
            assert len(a) == 1 and len(b) == 1, (axis, pat)
            assert a == b, (axis, pat)


def test_oracle_table_is_an_independent_but_identical_copy():
    """oracle/mc_table.py must not import the product's table, and the two must agree."""
    import inspect
    import oracle.mc_table as OT
    import icon_b200.mc_tables as PT
    assert "icon_b200" not in inspect.getsource(OT).split('"""', 2)[2]
    assert OT.TRI_LIST == PT.TRI_LIST and OT.EDGE_CORNERS == PT.EDGE_CORNERS and OT.CORNERS == PT.CORNERS
    for case in range(256):
        inside = [(case >> c) & 1 for c in range(8)]
        crossing = {e for e, (a, b) in enumerate(OT.EDGE_CORNERS) if inside[a] != inside[b]}
        assert set(OT.TRI_LIST[case]) == crossing
