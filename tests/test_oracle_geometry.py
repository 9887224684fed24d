"""The collision oracle (oracle/geometry.py) against an INDEPENDENT exact-rational statement of GEOS's closed-set
semantics: two polygons `intersect` iff two edges share a point or a vertex of one lies in the (closed) other; a
segment `intersects` a polygon iff it meets an edge or an endpoint lies inside; a polygon is `contained` in a box iff
all its vertices are in the closed box.  Rotations are exact (Pythagorean triples), coordinates are small rationals,
so touching configurations are decided exactly."""

from fractions import Fraction as F
from itertools import product

import numpy as np

from oracle import geometry as G

TRIPLES = [(F(1), F(0)), (F(0), F(1)), (F(3, 5), F(4, 5)), (F(4, 5), F(3, 5)), (F(5, 13), F(12, 13)), (F(-3, 5), F(4, 5)),
           (F(8, 17), F(-15, 17)), (F(-1), F(0))]


def corners(x, y, c, s, l, w):
    return [(x + cx * c - cy * s, y + cx * s + cy * c) for cx, cy in ((l, -w), (l, w), (-l, w), (-l, -w))]


def orient(a, b, p):
    v = (b[0] - a[0]) * (p[1] - a[1]) - (b[1] - a[1]) * (p[0] - a[0])
    return (v > 0) - (v < 0)


def on_seg(a, b, p):
    return min(a[0], b[0]) <= p[0] <= max(a[0], b[0]) and min(a[1], b[1]) <= p[1] <= max(a[1], b[1])


def seg_seg(a, b, c, d):
    o1, o2, o3, o4 = orient(a, b, c), orient(a, b, d), orient(c, d, a), orient(c, d, b)
    if o1 != o2 and o3 != o4:
        return True
    return (o1 == 0 and on_seg(a, b, c)) or (o2 == 0 and on_seg(a, b, d)) or (o3 == 0 and on_seg(c, d, a)) or (o4 == 0 and on_seg(c, d, b))


def in_convex(poly, p):
    signs = [orient(poly[i], poly[(i + 1) % 4], p) for i in range(4)]
    return all(s >= 0 for s in signs) or all(s <= 0 for s in signs)


def exact_poly_poly(A, B):
    if any(seg_seg(A[i], A[(i + 1) % 4], B[j], B[(j + 1) % 4]) for i in range(4) for j in range(4)):
        return True
    return in_convex(B, A[0]) or in_convex(A, B[0])


def exact_poly_seg(A, p, q):
    return any(seg_seg(A[i], A[(i + 1) % 4], p, q) for i in range(4)) or in_convex(A, p)


def f(v):
    return float(v)


def test_obb_obb_matches_exact_definition_including_touching():
    rng = np.random.default_rng(0)
    n_touch = n_hit = n_total = 0
    for _ in range(1500):
        (ca, sa), (cb, sb) = TRIPLES[rng.integers(len(TRIPLES))], TRIPLES[rng.integers(len(TRIPLES))]
        la, wa, lb, wb = (F(int(rng.integers(1, 9)), 4) for _ in range(4))
        xa, ya = F(int(rng.integers(-8, 9)), 2), F(int(rng.integers(-8, 9)), 2)
        xb, yb = xa + F(int(rng.integers(-12, 13)), 4), ya + F(int(rng.integers(-12, 13)), 4)
        A, B = corners(xa, ya, ca, sa, la, wa), corners(xb, yb, cb, sb, lb, wb)
        want = exact_poly_poly(A, B)
        got = bool(G.obb_obb(f(xa), f(ya), f(ca), f(sa), f(la), f(wa), f(xb), f(yb), f(cb), f(sb), f(lb), f(wb)))
        # only axis-aligned / dyadic cases are exact in float; for the others skip knife-edge configurations
        exact_float = all(v.denominator in (1, 2, 4, 8) for v in (ca, sa, cb, sb))
        if exact_float:
            assert got == want, (A, B)
            n_total += 1
            n_hit += want
            # touching: shrink B slightly -> disjoint would flip if it was only touching
        elif got != want:
            # must be a knife-edge: perturbing B by 1e-9 flips the exact answer
            far = exact_poly_poly(A, corners(xb + F(1, 10**6), yb + F(1, 10**6), cb, sb, lb, wb)) != exact_poly_poly(
                A, corners(xb - F(1, 10**6), yb - F(1, 10**6), cb, sb, lb, wb))
            assert far, (A, B)
    assert n_total > 100 and 0 < n_hit < n_total
    # explicit closed-set cases: edge touch, corner touch, containment
    assert G.obb_obb(0, 0, 1, 0, 2, 1, 4, 0, 1, 0, 2, 1)          # edge to edge
    assert G.obb_obb(0, 0, 1, 0, 2, 1, 4, 2, 1, 0, 2, 1)          # corner to corner
    assert not G.obb_obb(0, 0, 1, 0, 2, 1, 4.0000001, 0, 1, 0, 2, 1)
    assert G.obb_obb(0, 0, 1, 0, 4, 4, 0.5, 0.5, 0.6, 0.8, 1, 0.5)  # wholly inside


def test_obb_segment_and_disc_predicates_match_exact_definition():
    rng = np.random.default_rng(1)
    n = 0
    for _ in range(1500):
        c, s = [(F(1), F(0)), (F(0), F(1)), (F(-1), F(0))][rng.integers(3)]
        l, w = F(int(rng.integers(1, 9)), 4), F(int(rng.integers(1, 9)), 4)
        x, y = F(int(rng.integers(-8, 9)), 2), F(int(rng.integers(-8, 9)), 2)
        p = (x + F(int(rng.integers(-16, 17)), 4), y + F(int(rng.integers(-16, 17)), 4))
        q = (p[0] + F(int(rng.integers(-16, 17)), 4), p[1] + F(int(rng.integers(-16, 17)), 4))
        A = corners(x, y, c, s, l, w)
        want = exact_poly_seg(A, p, q)
        got = bool(G.obb_segment(f(x), f(y), f(c), f(s), f(l), f(w), f(p[0]), f(p[1]), f(q[0]), f(q[1])))
        assert got == want, (A, p, q)
        n += want
        # disc vs box: exact squared distance from the centre to the box
        r = F(int(rng.integers(1, 9)), 4)
        tx, ty = p[0] - x, p[1] - y
        qx, qy = abs(tx * c + ty * s) - l, abs(ty * c - tx * s) - w
        dx, dy = max(qx, 0), max(qy, 0)
        assert bool(G.obb_circle(f(x), f(y), f(c), f(s), f(l), f(w), f(p[0]), f(p[1]), f(r))) == (dx * dx + dy * dy <= r * r)
        # disc vs segment: exact squared distance
        ddx, ddy = q[0] - p[0], q[1] - p[1]
        dd = ddx * ddx + ddy * ddy
        t = F(0) if dd == 0 else max(F(0), min(F(1), ((x - p[0]) * ddx + (y - p[1]) * ddy) / dd))
        ex, ey = x - p[0] - t * ddx, y - p[1] - t * ddy
        assert bool(G.circle_segment(f(x), f(y), f(r), f(p[0]), f(p[1]), f(q[0]), f(q[1]))) == (ex * ex + ey * ey <= r * r)
    assert 100 < n < 1400
    assert G.obb_segment(0, 0, 1, 0, 2, 1, 2, -5, 2, 5)            # grazing the front edge
    assert G.obb_segment(0, 0, 1, 0, 2, 1, -0.5, 0, 0.5, 0)        # wholly inside
    assert G.obb_segment(0, 0, 1, 0, 2, 1, 2, 1, 3, 2)             # touches one corner
    assert not G.obb_segment(0, 0, 1, 0, 2, 1, 2.001, 1, 3, 2)
    assert G.circle_circle(0, 0, 1, 2, 0, 1) and not G.circle_circle(0, 0, 1, 2.0001, 0, 1)


def test_out_of_bound_is_not_contains():
    """box.contains(pose) is closed: a pose touching the boundary from inside is contained (not out)."""
    b = (-10.0, 10.0, -5.0, 5.0)
    for (x, y, c, s, l, w), want in [((8, 0, 1, 0, 2, 1), False), ((8.0001, 0, 1, 0, 2, 1), True), ((0, 4, 1, 0, 2, 1), False),
                                    ((0, 4.5, 0, 1, 2, 1), True), ((-8, 0, 0.6, 0.8, 2, 1), False), ((-9, 0, 0.6, 0.8, 2, 1), True)]:
        ex, ey = G.extents(c, s, l, w)
        assert bool(G.out_of_bound(x, y, ex, ey, b)) == want
        cs = G.obb_corners(x, y, np.arctan2(s, c), l, w)
        inside = (cs[:, 0] >= b[0]).all() and (cs[:, 0] <= b[1]).all() and (cs[:, 1] >= b[2]).all() and (cs[:, 1] <= b[3]).all()
        assert inside == (not want) or abs(abs(cs).max() - 10) < 1e-9


def test_pose_corner_order_follows_reference_ring():
    """vehicle.py:133-140: [(+L/2,-W/2), (+L/2,+W/2), (-L/2,+W/2), (-L/2,-W/2)] rotated by the heading."""
    cs = G.obb_corners(10.0, 5.0, np.pi / 2, 2.142, 0.8995)
    np.testing.assert_allclose(cs, [[10.8995, 7.142], [9.1005, 7.142], [9.1005, 2.858], [10.8995, 2.858]], atol=1e-12)


def _winding_inside(poly, p):
    """Exact point-in-(possibly concave)-polygon by the winding number over rational coordinates; None on the boundary."""
    wn = 0
    n = len(poly)
    for i in range(n):
        a, b = poly[i], poly[(i + 1) % n]
        if orient(a, b, p) == 0 and on_seg(a, b, p):
            return None
        if a[1] <= p[1]:
            if b[1] > p[1] and orient(a, b, p) > 0:
                wn += 1
        elif b[1] <= p[1] and orient(a, b, p) < 0:
            wn -= 1
    return wn != 0


def test_pose_vs_area_polygon_matches_exact_definition():
    """oracle.scenario.events with poly_start (a pose intersects an Area polygon iff one of its edges meets the pose OR the
    pose centre lies inside; the first OBJECT hit is reported by its first segment) against the exact-rational statement of
    shapely's `pose.intersects(polygon)` for closed sets: two boundaries share a point, or one shape contains a point of the
    other - on convex and concave rings, including poses wholly inside, rings wholly inside the pose, and touching."""
    from oracle import scenario as O
    from tactics2d_b200.map import polygons_to_segments

    rings = [
        [(F(-30), F(18)), (F(30), F(18)), (F(30), F(19)), (F(-30), F(19))],
        [(F(-25), F(-20)), (F(-5), F(-20)), (F(-5), F(-4)), (F(-25), F(-4))],
        [(F(4), F(-18)), (F(16), F(-18)), (F(16), F(-6)), (F(12), F(-6)), (F(12), F(-14)), (F(4), F(-14))],    # concave
        [(F(20), F(5)), (F(41, 2), F(5)), (F(41, 2), F(11, 2)), (F(20), F(11, 2))],                             # smaller than the pose
    ]
    seg, ps = polygons_to_segments([[(float(x), float(y)) for x, y in r] for r in rings])
    table = dict(half_len=np.array([2.0], np.float32), half_wid=np.array([1.0], np.float32), radius=np.array([0.0], np.float32),
                 shape=np.array([0], np.int32), model=np.array([0], np.int32))
    for k in O.TABLE_FLOAT_FIELDS:
        table.setdefault(k, np.array([1.0], np.float32))
    rng = np.random.default_rng(1)
    n_inside_only = n_hit = n_free = 0
    cases = [(F(-15), F(-12), 2), (F(8), F(-10), 0), (F(81, 4), F(21, 4), 3), (F(0), F(37, 2), 1), (F(-3), F(-12), 0)]   # hand-placed
    for _ in range(400):
        cases.append((F(int(rng.integers(-140, 141)), 4), F(int(rng.integers(-100, 101)), 4), int(rng.integers(0, len(TRIPLES)))))
    for x, y, t in cases:
        c, s = TRIPLES[t]
        pose = corners(x, y, c, s, F(2), F(1))
        want = -1
        for p, ring in enumerate(rings):
            hit = any(seg_seg(pose[i], pose[(i + 1) % 4], ring[j], ring[(j + 1) % len(ring)]) for i in range(4) for j in range(len(ring)))
            if not hit:
                ins = _winding_inside(ring, (x, y))
                hit = bool(ins) or in_convex(pose, ring[0])          # pose inside the ring / ring inside the pose
            if hit:
                want = int(ps[p])
                break
        heading = float(np.arctan2(float(s), float(c)))
        fl, hi, hs = O.events(np.array([[f(x)]]), np.array([[f(y)]]), np.array([[heading]]), np.zeros((1, 1), np.uint8), table, seg, None, poly_start=ps)
        assert int(hs[0, 0]) == want, (x, y, t, int(hs[0, 0]), want)
        assert bool(fl[0, 0] & 2) == (want >= 0)
        plain = O.events(np.array([[f(x)]]), np.array([[f(y)]]), np.array([[heading]]), np.zeros((1, 1), np.uint8), table, seg, None)[2][0, 0]
        n_inside_only += int(want >= 0 and plain < 0)
        n_hit += int(want >= 0)
        n_free += int(want < 0)
    assert n_inside_only >= 5 and n_hit >= 40 and n_free >= 100, (n_inside_only, n_hit, n_free)


def test_pose_vs_area_with_holes_matches_exact_definition():
    """An Area with holes (``Polygon(outer, inners)``, parse_osm.py:461-510; one walkway of inD_2 has two) is one object
    whose edge range holds all its rings: the edge tests see every ring and the parity of ALL the edges a ray from the centre
    crosses is "inside the exterior, outside every hole".  Checked against the exact-rational definition - a pose parked
    wholly inside a hole touches nothing, one in the solid part is a hit, and the object after it keeps its own index."""
    from oracle import scenario as O
    from tactics2d_b200.map import Area, polygons_to_segments

    def sq(x0, y0, x1, y1):
        return [(F(x0), F(y0)), (F(x1), F(y0)), (F(x1), F(y1)), (F(x0), F(y1))]

    objects = [
        (sq(-30, -20, 10, 20), [sq(-26, -14, -12, 0), [(F(-6), F(4)), (F(6), F(4)), (F(6), F(16)), (F(0), F(10)), (F(-6), F(16))]]),   # two holes, one concave
        (sq(14, -20, 30, -4), []),
    ]
    as_float = lambda ring: np.array([[float(a), float(b)] for a, b in ring])
    areas = [Area(i, "multipolygon", "walkway", as_float(o), [as_float(h) for h in hs]) for i, (o, hs) in enumerate(objects)]
    seg, ps = polygons_to_segments(areas)
    assert ps.tolist() == [0, 4 + 4 + 5, 4 + 4 + 5 + 4]
    table = dict(half_len=np.array([2.0], np.float32), half_wid=np.array([1.0], np.float32), radius=np.array([0.0], np.float32),
                 shape=np.array([0], np.int32), model=np.array([0], np.int32))
    for k in O.TABLE_FLOAT_FIELDS:
        table.setdefault(k, np.array([1.0], np.float32))
    rng = np.random.default_rng(5)
    cases = [(F(-19), F(-7), 0), (F(0), F(8), 1), (F(-20), F(10), 2), (F(22), F(-12), 0), (F(-12), F(-7), 3), (F(40), F(0), 0)]
    for _ in range(500):
        cases.append((F(int(rng.integers(-140, 141)), 4), F(int(rng.integers(-100, 101)), 4), int(rng.integers(0, len(TRIPLES)))))
    n_in_hole = n_solid = n_edge = 0
    for x, y, t in cases:
        c, s = TRIPLES[t]
        pose = corners(x, y, c, s, F(2), F(1))
        want = -1
        for p, (outer, holes) in enumerate(objects):
            rings = [outer] + holes
            edge = any(seg_seg(pose[i], pose[(i + 1) % 4], r[j], r[(j + 1) % len(r)]) for r in rings for i in range(4) for j in range(len(r)))
            inside = lambda ring: bool(_winding_inside(ring, (x, y)))               # None (centre on an edge) is an edge hit anyway
            solid = inside(outer) and not any(inside(h) for h in holes)
            swallowed = any(in_convex(pose, r[0]) for r in rings)                  # a ring wholly inside the pose
            if p == 0:
                n_edge += int(edge)
                n_solid += int(solid and not edge)
                n_in_hole += int(not edge and not solid and inside(outer))
            if edge or solid or swallowed:
                want = int(ps[p])
                break
        heading = float(np.arctan2(float(s), float(c)))
        fl, hi, hs_ = O.events(np.array([[f(x)]]), np.array([[f(y)]]), np.array([[heading]]), np.zeros((1, 1), np.uint8), table, seg, None, poly_start=ps)
        assert int(hs_[0, 0]) == want, (x, y, t, int(hs_[0, 0]), want)
        assert bool(fl[0, 0] & 2) == (want >= 0)
    assert n_in_hole >= 10 and n_solid >= 40 and n_edge >= 40, (n_in_hole, n_solid, n_edge)

