"""The collision oracle (oracle/geometry.py) against geometry code that is NOT ours.

shapely / GEOS is absent from this image, so the collision half of the oracle cannot be pinned to the reference itself
("parity unpinned").  Two independent packages that ARE here narrow the gap:

* ``sympy.geometry`` - exact arithmetic on rationals: ``Polygon.intersection`` (all boundary points two shapes share),
  ``Polygon.encloses_point`` / ``Circle.encloses_point`` (strict interior).  GEOS's ``intersects`` on closed sets is
  "the boundaries share a point, or one shape holds a point of the other", so every predicate of the oracle is restated
  here with sympy's primitives - touching configurations included, decided exactly.
* OpenCV - its own float32 implementation of rotated-rectangle intersection (``cv2.rotatedRectangleIntersection``) and
  convex clipping (``cv2.intersectConvexConvex``): the OBB-OBB flag away from knife edges and the IoU that
  ``Arrival`` / ``NoAction`` threshold (arrival.py:42-46, no_action.py:43-46).

Rotations are Pythagorean triples and coordinates small rationals, exactly representable where the float oracle is asked
for an exact answer."""

from fractions import Fraction as F

import numpy as np
import pytest

from oracle import geometry as G

sympy = pytest.importorskip("sympy")
from sympy import Circle, Point, Polygon, Rational, Segment   # noqa: E402

DYADIC = [(F(1), F(0)), (F(0), F(1)), (F(-1), F(0)), (F(0), F(-1))]
TRIPLES = DYADIC + [(F(3, 5), F(4, 5)), (F(-4, 5), F(3, 5)), (F(5, 13), F(-12, 13)), (F(8, 17), F(15, 17))]


def R(v):
    return Rational(v.numerator, v.denominator)


def corners(x, y, c, s, l, w):
    return [(x + cx * c - cy * s, y + cx * s + cy * c) for cx, cy in ((l, -w), (l, w), (-l, w), (-l, -w))]


def spoly(pts):
    return Polygon(*[Point(R(px), R(py)) for px, py in pts])


def f(v):
    return float(v)


def test_obb_obb_against_sympy_including_touching():
    rng = np.random.default_rng(21)
    n = n_hit = n_touch = 0
    cases = []
    for _ in range(70):
        (ca, sa), (cb, sb) = DYADIC[rng.integers(4)], DYADIC[rng.integers(4)]
        la, wa, lb, wb = (F(int(rng.integers(1, 9)), 4) for _ in range(4))
        xa, ya = F(int(rng.integers(-8, 9)), 2), F(int(rng.integers(-8, 9)), 2)
        cases.append((xa, ya, ca, sa, la, wa, xa + F(int(rng.integers(-12, 13)), 4), ya + F(int(rng.integers(-12, 13)), 4), cb, sb, lb, wb))
    # hand-placed closed-set cases: edge on edge, corner on corner, corner on edge, one inside the other, a hair apart
    cases += [(F(0), F(0), F(1), F(0), F(2), F(1), F(4), F(0), F(1), F(0), F(2), F(1)),
              (F(0), F(0), F(1), F(0), F(2), F(1), F(4), F(2), F(1), F(0), F(2), F(1)),
              (F(0), F(0), F(1), F(0), F(2), F(1), F(3), F(0), F(0), F(1), F(1), F(1)),
              (F(0), F(0), F(1), F(0), F(4), F(4), F(1, 2), F(1, 2), F(0), F(1), F(1), F(1, 2)),
              (F(0), F(0), F(1), F(0), F(2), F(1), F(4) + F(1, 1024), F(0), F(1), F(0), F(2), F(1))]
    for xa, ya, ca, sa, la, wa, xb, yb, cb, sb, lb, wb in cases:
        A, B = corners(xa, ya, ca, sa, la, wa), corners(xb, yb, cb, sb, lb, wb)
        pa, pb = spoly(A), spoly(B)
        shared = pa.intersection(pb)
        want = bool(shared) or bool(pa.encloses_point(Point(R(xb), R(yb)))) or bool(pb.encloses_point(Point(R(xa), R(ya))))
        got = bool(G.obb_obb(f(xa), f(ya), f(ca), f(sa), f(la), f(wa), f(xb), f(yb), f(cb), f(sb), f(lb), f(wb)))
        assert got == want, (A, B)
        n += 1
        n_hit += want
        # touching = they meet but share no interior point: every shared piece lies on both boundaries and neither centre region overlaps
        if want and all(isinstance(p, (Point, Segment)) for p in shared) and not pa.encloses_point(Point(R(xb), R(yb))):
            n_touch += 1
    assert n >= 70 and 10 < n_hit < n - 10 and n_touch >= 3


def test_rotated_obb_obb_against_sympy_off_the_knife_edge():
    """Arbitrary (Pythagorean) rotations: the exact answer from sympy, the float answer from the oracle; they may only
    differ where a move of one part in a million flips the exact answer."""
    rng = np.random.default_rng(22)
    n = n_hit = 0
    for _ in range(60):
        (ca, sa), (cb, sb) = TRIPLES[rng.integers(len(TRIPLES))], TRIPLES[rng.integers(len(TRIPLES))]
        la, wa, lb, wb = (F(int(rng.integers(2, 9)), 4) for _ in range(4))
        xa, ya = F(int(rng.integers(-8, 9)), 2), F(int(rng.integers(-8, 9)), 2)
        xb, yb = xa + F(int(rng.integers(-14, 15)), 4), ya + F(int(rng.integers(-14, 15)), 4)

        def exact(dx):
            pa, pb = spoly(corners(xa, ya, ca, sa, la, wa)), spoly(corners(xb + dx, yb + dx, cb, sb, lb, wb))
            return bool(pa.intersection(pb)) or bool(pa.encloses_point(Point(R(xb + dx), R(yb + dx)))) or bool(pb.encloses_point(Point(R(xa), R(ya))))

        want = exact(F(0))
        got = bool(G.obb_obb(f(xa), f(ya), f(ca), f(sa), f(la), f(wa), f(xb), f(yb), f(cb), f(sb), f(lb), f(wb)))
        if got != want:
            assert exact(F(1, 10**6)) != exact(F(-1, 10**6)), (xa, ya, xb, yb)
        n += 1
        n_hit += want
    assert 10 < n_hit < n - 10


def test_obb_segment_against_sympy_including_touching():
    rng = np.random.default_rng(23)
    cases = []
    for _ in range(80):
        c, s = DYADIC[rng.integers(4)]
        l, w = F(int(rng.integers(1, 9)), 4), F(int(rng.integers(1, 9)), 4)
        x, y = F(int(rng.integers(-8, 9)), 2), F(int(rng.integers(-8, 9)), 2)
        p = (x + F(int(rng.integers(-16, 17)), 4), y + F(int(rng.integers(-16, 17)), 4))
        q = (p[0] + F(int(rng.integers(-16, 17)), 4), p[1] + F(int(rng.integers(-16, 17)), 4))
        if p != q:
            cases.append((x, y, c, s, l, w, p, q))
    cases += [(F(0), F(0), F(1), F(0), F(2), F(1), (F(2), F(-5)), (F(2), F(5))),            # along the front edge
              (F(0), F(0), F(1), F(0), F(2), F(1), (F(-1, 2), F(0)), (F(1, 2), F(0))),      # wholly inside
              (F(0), F(0), F(1), F(0), F(2), F(1), (F(2), F(1)), (F(3), F(2))),             # from one corner outwards
              (F(0), F(0), F(1), F(0), F(2), F(1), (F(2) + F(1, 512), F(1)), (F(3), F(2)))]
    n_hit = 0
    for x, y, c, s, l, w, p, q in cases:
        box = spoly(corners(x, y, c, s, l, w))
        seg = Segment(Point(R(p[0]), R(p[1])), Point(R(q[0]), R(q[1])))
        want = bool(box.intersection(seg)) or bool(box.encloses_point(seg.p1))
        got = bool(G.obb_segment(f(x), f(y), f(c), f(s), f(l), f(w), f(p[0]), f(p[1]), f(q[0]), f(q[1])))
        assert got == want, (x, y, c, s, l, w, p, q)
        n_hit += want
    assert 15 < n_hit < len(cases) - 15


def test_disc_predicates_against_sympy_including_tangency():
    rng = np.random.default_rng(24)
    n_seg = n_box = n_cc = 0
    for k in range(60):
        x, y = F(int(rng.integers(-8, 9)), 2), F(int(rng.integers(-8, 9)), 2)
        r = F(int(rng.integers(1, 9)), 4)
        p = (x + F(int(rng.integers(-12, 13)), 4), y + F(int(rng.integers(-12, 13)), 4))
        q = (p[0] + F(int(rng.integers(-12, 13)), 4), p[1] + F(int(rng.integers(-12, 13)), 4))
        disc = Circle(Point(R(x), R(y)), R(r))
        if p != q:
            seg = Segment(Point(R(p[0]), R(p[1])), Point(R(q[0]), R(q[1])))
            want = bool(disc.intersection(seg)) or bool(disc.encloses_point(seg.p1))
            assert bool(G.circle_segment(f(x), f(y), f(r), f(p[0]), f(p[1]), f(q[0]), f(q[1]))) == want, (x, y, r, p, q)
            n_seg += want
        # box vs disc (Pedestrian against Vehicle): the box sits at p
        c, s = DYADIC[rng.integers(4)]
        l, w = F(int(rng.integers(1, 7)), 4), F(int(rng.integers(1, 7)), 4)
        box = spoly(corners(p[0], p[1], c, s, l, w))
        want = bool(box.intersection(disc)) or bool(box.encloses_point(disc.center)) or bool(disc.encloses_point(Point(R(p[0]), R(p[1]))))
        assert bool(G.obb_circle(f(p[0]), f(p[1]), f(c), f(s), f(l), f(w), f(x), f(y), f(r))) == want, (p, c, s, l, w, x, y, r)
        n_box += want
        # disc vs disc
        r2 = F(int(rng.integers(1, 9)), 4)
        other = Circle(Point(R(p[0]), R(p[1])), R(r2))
        if (x, y) != p:
            # a disc strictly inside the other has its centre inside it, so boundaries + centres cover every case
            want = bool(disc.intersection(other)) or bool(disc.encloses_point(other.center)) or bool(other.encloses_point(disc.center))
            assert bool(G.circle_circle(f(x), f(y), f(r), f(p[0]), f(p[1]), f(r2))) == bool(want), (x, y, r, p, r2)
            n_cc += bool(want)
    assert n_seg >= 8 and n_box >= 8 and n_cc >= 8
    # tangency is a hit (closed sets): a disc of radius 1 against the line y = 1, another disc at distance exactly 2, a box edge
    tangent = Circle(Point(0, 0), 1)
    assert tangent.intersection(Segment(Point(-3, 1), Point(3, 1))) and G.circle_segment(0.0, 0.0, 1.0, -3.0, 1.0, 3.0, 1.0)
    assert tangent.intersection(Circle(Point(2, 0), 1)) and G.circle_circle(0.0, 0.0, 1.0, 2.0, 0.0, 1.0)
    assert spoly(corners(F(3), F(0), F(1), F(0), F(2), F(1))).intersection(tangent) and G.obb_circle(3.0, 0.0, 1.0, 0.0, 2.0, 1.0, 0.0, 0.0, 1.0)
    assert not G.circle_segment(0.0, 0.0, 1.0, -3.0, 1.0 + 1e-9, 3.0, 1.0 + 1e-9)


def test_point_in_area_with_holes_against_sympy():
    """``point_in_ring`` over all the rings of an Area at once (exterior + holes) against sympy: inside the exterior polygon
    and inside none of the holes."""
    outer = [(F(-30), F(-20)), (F(10), F(-20)), (F(10), F(20)), (F(-12), F(24)), (F(-30), F(20))]
    holes = [[(F(-26), F(-14)), (F(-12), F(-14)), (F(-12), F(0)), (F(-26), F(0))],
             [(F(-6), F(4)), (F(6), F(4)), (F(6), F(16)), (F(0), F(10)), (F(-6), F(16))]]
    so, sh = spoly(outer), [spoly(h) for h in holes]
    edges = np.concatenate([np.array([[f(r[i][0]), f(r[i][1]), f(r[(i + 1) % len(r)][0]), f(r[(i + 1) % len(r)][1])] for i in range(len(r))])
                            for r in [outer] + holes])
    rng = np.random.default_rng(25)
    n_in = n_hole = n_out = 0
    for _ in range(300):
        x, y = F(int(rng.integers(-140, 61)), 4) + F(1, 8), F(int(rng.integers(-100, 111)), 4) + F(1, 16)   # never on an edge or level with a vertex
        p = Point(R(x), R(y))
        in_hole = any(h.encloses_point(p) for h in sh)
        want = bool(so.encloses_point(p)) and not in_hole
        assert bool(G.point_in_ring(np.array([f(x)]), np.array([f(y)]), edges)[0]) == want, (x, y)
        n_in += want
        n_hole += in_hole
        n_out += (not want and not in_hole)
    assert n_in > 60 and n_hole > 15 and n_out > 40


def test_obb_obb_and_iou_against_opencv():
    """OpenCV's own rotated-rectangle code (float32): the intersect / disjoint flag wherever the two shapes are at least a
    millimetre into or away from each other, and the IoU of ``Arrival.update`` (arrival.py:42-46) to 1e-4."""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(26)
    n_hit = n_free = n_iou = 0
    for _ in range(3000):
        xa, ya, xb, yb = rng.uniform(-6, 6, 4)
        ha, hb = rng.uniform(-np.pi, np.pi, 2)
        la, lb = rng.uniform(1.0, 3.0, 2)
        wa, wb = rng.uniform(0.4, 1.2, 2)
        got = bool(G.obb_obb(xa, ya, np.cos(ha), np.sin(ha), la, wa, xb, yb, np.cos(hb), np.sin(hb), lb, wb))
        # the same test with B grown / shrunk by a millimetre: only unambiguous configurations are compared
        grown = bool(G.obb_obb(xa, ya, np.cos(ha), np.sin(ha), la, wa, xb, yb, np.cos(hb), np.sin(hb), lb + 1e-3, wb + 1e-3))
        shrunk = bool(G.obb_obb(xa, ya, np.cos(ha), np.sin(ha), la, wa, xb, yb, np.cos(hb), np.sin(hb), lb - 1e-3, wb - 1e-3))
        ra = ((float(xa), float(ya)), (float(2 * la), float(2 * wa)), float(np.degrees(ha)))
        rb = ((float(xb), float(yb)), (float(2 * lb), float(2 * wb)), float(np.degrees(hb)))
        kind, _ = cv2.rotatedRectangleIntersection(ra, rb)
        if grown == shrunk:
            assert (kind != cv2.INTERSECT_NONE) == got, (ra, rb, kind)
            n_hit += got
            n_free += not got
        A = G.obb_corners(xa, ya, ha, la, wa).astype(np.float32)
        B = G.obb_corners(xb, yb, hb, lb, wb).astype(np.float32)
        area, _ = cv2.intersectConvexConvex(A, B)
        iou_cv = area / (4 * la * wa + 4 * lb * wb - area)
        iou = G.rect_iou(xa, ya, ha, la, wa, xb, yb, hb, lb, wb)
        if shrunk:                                       # OpenCV reports area 0 for slivers; compare real overlaps
            assert abs(iou - iou_cv) < 1e-4, (ra, rb, iou, iou_cv)
            n_iou += 1
        else:
            assert iou < 2e-3
    assert n_hit > 300 and n_free > 300 and n_iou > 300
    # identical rectangles and containment
    assert abs(G.rect_iou(1.0, 2.0, 0.7, 2.0, 1.0, 1.0, 2.0, 0.7, 2.0, 1.0) - 1.0) < 1e-12
    assert abs(G.rect_iou(0.0, 0.0, 0.0, 2.0, 1.0, 0.0, 0.0, 0.0, 1.0, 0.5) - 0.25) < 1e-12
