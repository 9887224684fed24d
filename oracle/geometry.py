"""Float64 pose and closed-set predicates (TEST INFRASTRUCTURE ONLY).

The reference obtains every collision / containment result from shapely/GEOS
(third-party, ``shapely>=2.0.7,<2.1.0``, requirements.txt:18; not installable
here) at these call sites: ``traffic/event_detection/collision.py:22,40``
(``intersects``) and ``traffic/event_detection/out_bound.py:48`` (``contains``).
This module restates the documented GEOS semantics for the shapes the path uses:

* ``intersects``: the two *closed* point sets share at least one point - touching
  counts, a segment wholly inside a box counts, a box wholly inside a box counts;
* ``box.contains(pose)``: no point of ``pose`` lies in the exterior of ``box`` -
  contact with the boundary from inside is still contained.

PARITY UNPINNED for this file: no reference test pins a collision result and the
GEOS build cannot be run here.  ``tests/test_oracle_geometry.py`` cross-checks the
formulas below against an independent exact-rational definition (edges cross, or
one shape holds a vertex of the other); ``tests/test_oracle_thirdparty.py`` checks
them against code that is not ours - sympy.geometry (exact arithmetic, touching and
tangent cases included) and OpenCV's rotated-rectangle intersection / convex
clipping (flag and IoU).  Neither is GEOS: the label stays.

Pose: ``Vehicle.get_pose`` (participant/element/vehicle.py:263-281) maps the local
ring ``[(+L/2,-W/2), (+L/2,+W/2), (-L/2,+W/2), (-L/2,-W/2)]`` (:133-140) by
``x' = x + cx cos h - cy sin h ; y' = y + cx sin h + cy cos h``.  ``Cyclist``
(cyclist.py:98-105,155-175) is identical; ``Pedestrian.get_pose``
(pedestrian.py:138-149) is the circle ``((x, y), width/2)``.
"""

from __future__ import annotations

import numpy as np


def obb_corners(x, y, h, hl, hw):
    """vehicle.py:133-140,272-281 -> array [..., 4, 2] in the reference's ring order."""
    x, y, h, hl, hw = (np.asarray(a, dtype=np.float64) for a in (x, y, h, hl, hw))
    c, s = np.cos(h), np.sin(h)
    cx = np.stack([hl, hl, -hl, -hl], -1)
    cy = np.stack([-hw, hw, hw, -hw], -1)
    px = x[..., None] + cx * c[..., None] - cy * s[..., None]
    py = y[..., None] + cx * s[..., None] + cy * c[..., None]
    return np.stack([px, py], -1)


def obb_obb(xa, ya, ca, sa, la, wa, xb, yb, cb, sb, lb, wb):
    """Closed rectangles intersect <=> no face normal separates them (SAT, non-strict).

    (c, s) = cos/sin of the heading difference; the four axes are A's and B's edge
    directions.  Touching (equality on an axis) counts, as GEOS ``intersects`` does.
    """
    tx, ty = xb - xa, yb - ya
    c = ca * cb + sa * sb
    s = ca * sb - sa * cb
    ac, as_ = np.abs(c), np.abs(s)
    return ((np.abs(tx * ca + ty * sa) <= la + (lb * ac + wb * as_))
            & (np.abs(ty * ca - tx * sa) <= wa + (lb * as_ + wb * ac))
            & (np.abs(tx * cb + ty * sb) <= lb + (la * ac + wa * as_))
            & (np.abs(ty * cb - tx * sb) <= wb + (la * as_ + wa * ac)))


def obb_circle(xa, ya, ca, sa, la, wa, xc, yc, r):
    """Closed rectangle vs closed disc: distance from the disc centre to the box <= r."""
    tx, ty = xc - xa, yc - ya
    qx = np.abs(tx * ca + ty * sa) - la
    qy = np.abs(ty * ca - tx * sa) - wa
    dx, dy = np.maximum(qx, 0.0), np.maximum(qy, 0.0)
    return dx * dx + dy * dy <= r * r


def circle_circle(xa, ya, ra, xb, yb, rb):
    tx, ty = xb - xa, yb - ya
    return tx * tx + ty * ty <= (ra + rb) * (ra + rb)


def obb_segment(xa, ya, ca, sa, la, wa, x1, y1, x2, y2):
    """Closed rectangle vs closed segment, in the rectangle's frame (SAT: the two box
    axes and the segment normal)."""
    ux, uy = x1 - xa, y1 - ya
    vx, vy = x2 - xa, y2 - ya
    p1x, p1y = ux * ca + uy * sa, uy * ca - ux * sa
    p2x, p2y = vx * ca + vy * sa, vy * ca - vx * sa
    dx, dy = p2x - p1x, p2y - p1y
    return ((np.maximum(p1x, p2x) >= -la) & (np.minimum(p1x, p2x) <= la)
            & (np.maximum(p1y, p2y) >= -wa) & (np.minimum(p1y, p2y) <= wa)
            & (np.abs(p1x * dy - p1y * dx) <= la * np.abs(dy) + wa * np.abs(dx)))


def circle_segment(xc, yc, r, x1, y1, x2, y2):
    """Closed disc vs closed segment: distance from the centre to the segment <= r."""
    dx, dy = x2 - x1, y2 - y1
    ux, uy = xc - x1, yc - y1
    dd = dx * dx + dy * dy
    with np.errstate(divide="ignore", invalid="ignore"):
        t = np.where(dd > 0, (ux * dx + uy * dy) / np.where(dd > 0, dd, 1.0), 0.0)
    t = np.clip(t, 0.0, 1.0)
    ex, ey = ux - t * dx, uy - t * dy
    return ex * ex + ey * ey <= r * r


def extents(c, s, hl, hw):
    """Half-sizes of the axis-aligned box around the rotated rectangle."""
    return hl * np.abs(c) + hw * np.abs(s), hl * np.abs(s) + hw * np.abs(c)


def out_of_bound(x, y, ex, ey, bounds):
    """OutBound.update, out_bound.py:37-48: ``not Polygon(box).contains(pose)`` with
    box = (xmin, xmax, ymin, ymax) (:28-36).  ``contains`` is closed, so the pose is
    out as soon as one corner is *strictly* outside."""
    xmin, xmax, ymin, ymax = bounds
    return (x - ex < xmin) | (x + ex > xmax) | (y - ey < ymin) | (y + ey > ymax)


def rect_iou(xa, ya, ha, la, wa, xb, yb, hb, lb, wb):
    """IoU of two rotated rectangles (scalar float64): ``intersection.area / union.area`` as
    ``Arrival.update`` (arrival.py:42-46) and ``NoAction.update`` (no_action.py:43-46) compute it with
    shapely.  Intersection of two convex rings by half-plane clipping, areas by the shoelace formula."""
    A = obb_corners(xa, ya, ha, la, wa).tolist()
    B = obb_corners(xb, yb, hb, lb, wb).tolist()
    poly = A
    for e in range(4):
        (x0, y0), (x1, y1) = B[e], B[(e + 1) % 4]
        ex, ey = x1 - x0, y1 - y0
        out = []
        for i in range(len(poly)):
            p, q = poly[i], poly[(i + 1) % len(poly)]
            dp = ex * (p[1] - y0) - ey * (p[0] - x0)
            dq = ex * (q[1] - y0) - ey * (q[0] - x0)
            if dp >= 0:
                out.append(p)
            if (dp >= 0) != (dq >= 0):
                t = dp / (dp - dq)
                out.append([p[0] + t * (q[0] - p[0]), p[1] + t * (q[1] - p[1])])
        poly = out
        if not poly:
            break
    inter = 0.5 * abs(sum(poly[i][0] * poly[(i + 1) % len(poly)][1] - poly[(i + 1) % len(poly)][0] * poly[i][1]
                          for i in range(len(poly)))) if poly else 0.0
    union = 4.0 * la * wa + 4.0 * lb * wb - inter
    return inter / union if union > 0 else 0.0


def point_in_ring(px, py, ring):
    """Even-odd (crossing number) point-in-polygon for points ``px, py`` (arrays) against the closed ring given by its
    edges ``ring`` [E, 4] = (x1, y1, x2, y2): what ``Polygon.contains`` / ``intersects`` decides for a point strictly
    inside or outside (shapely's Area.geometry, area.py:13-125).  Points on the boundary are not this function's business:
    a pose whose centre is that close to an edge intersects the edge itself."""
    px, py = np.asarray(px, dtype=np.float64), np.asarray(py, dtype=np.float64)
    ring = np.asarray(ring, dtype=np.float64)
    inside = np.zeros(px.shape, bool)
    for x1, y1, x2, y2 in ring:
        straddle = (y1 > py) != (y2 > py)
        with np.errstate(divide="ignore", invalid="ignore"):
            xint = (x2 - x1) * (py - y1) / (y2 - y1) + x1
        inside ^= straddle & (px < xint)
    return inside
