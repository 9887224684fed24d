"""Static map geometry for the batched tick: Lanelet2-OSM -> collidable segment tiles.

The kernels take the scenario's static geometry as a flat tile ``float32 [S, 4] = (x1, y1, x2, y2)``
plus the out-of-bound box.  This module is the (offline, host-side) compiler from the reference's
on-disk format to that tile; it follows the reference's own parsing rules:

* node projection without pyproj: ``x = (lon - lon0) * 111320 * cos(lat0)``, ``y = (lat - lat0) * 110540``
  with the first non-deleted node as origin   (tactics2d/map/parser/parse_osm.py:259-278, :589-597);
* a Lanelet2 ``way`` is a RoadLine whose geometry is the LineString of its ``nd`` refs in order and
  whose ``type`` / ``subtype`` tags classify it                       (parse_osm.py:384-403, :82-120);
* barrier kinds (no lane change, physically blocking): type ``curbstone`` / ``road_border`` and
  ``guard_rail`` / ``wall`` / ``fence`` ...                     (tactics2d/map/element/roadline.py:107-123);
* ``Map.boundary`` = (floor(xmin), ceil(xmax), floor(ymin), ceil(ymax)) over all nodes
                                                              (tactics2d/map/element/map.py:92-167).

The tiles of the reference's 13 bundled maps (``data/{highD,inD,rounD}_map/*.osm``) are compiled once in
the build container (``python -m tactics2d_b200.map`` -> ``tactics2d_b200/map/tiles/*.npz``) because
``/root/reference`` does not exist on the GPU box.
"""

from __future__ import annotations

import math
import os
import xml.etree.ElementTree as ET
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

TILE_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tiles")

# RoadLine kinds that are physical structures a vehicle cannot drive through.  The reference gives every one of them the
# lane-change rule (False, False): type curbstone / road_border at roadline.py:109-110, subtypes guard_rail, wall, fence,
# jersey_barrier, gate, door, rail at roadline.py:111-124.  The PAINTED members of that same list (zebra_marking,
# pedestrian_marking, bike_marking, keepout, roadline.py:115-118) carry the same rule but are not obstacles and stay out.
BARRIER_TYPES = ("curbstone", "road_border", "guard_rail", "wall", "fence", "jersey_barrier", "gate", "door", "rail")


@dataclass
class Way:
    id_: int
    type_: Optional[str]
    subtype: Optional[str]
    points: np.ndarray   # float64 [P, 2]


@dataclass
class OsmMap:
    nodes: Dict[int, Tuple[float, float]]
    ways: List[Way]

    @property
    def boundary(self) -> Tuple[float, float, float, float]:
        """Map.boundary, map.py:92-167."""
        if not self.nodes:
            return (0.0, 0.0, 0.0, 0.0)
        xy = np.asarray(list(self.nodes.values()), dtype=np.float64)
        return (float(np.floor(xy[:, 0].min())), float(np.ceil(xy[:, 0].max())), float(np.floor(xy[:, 1].min())),
                float(np.ceil(xy[:, 1].max())))


def parse_osm_lanelet2(path: str) -> OsmMap:
    """OSMParser(lanelet2=True).parse without a projector (parse_osm.py:536-619): nodes and ways only."""
    root = ET.parse(path).getroot()
    all_nodes = [n for n in root.findall("node") if n.get("action") != "delete"]
    lat0 = float(all_nodes[0].attrib["lat"]) if all_nodes else 0.0
    lon0 = float(all_nodes[0].attrib["lon"]) if all_nodes else 0.0
    k = 111320.0 * math.cos(math.radians(lat0))
    nodes = {int(n.attrib["id"]): ((float(n.attrib["lon"]) - lon0) * k, (float(n.attrib["lat"]) - lat0) * 110540.0)
             for n in all_nodes}
    ways = []
    for w in root.findall("way"):
        if w.get("action") == "delete":
            continue
        tags = {t.attrib["k"]: t.attrib["v"] for t in w.findall("tag")}
        pts = np.asarray([nodes[int(nd.attrib["ref"])] for nd in w.findall("nd")], dtype=np.float64).reshape(-1, 2)
        ways.append(Way(int(w.attrib["id"]), tags.get("type"), tags.get("subtype"), pts))
    return OsmMap(nodes, ways)


def collidable_segments(map_: OsmMap, barrier_types: Sequence[str] = BARRIER_TYPES, solid_lines: Optional[bool] = None) -> np.ndarray:
    """Flatten the blocking RoadLines into segments, ways in file order, pieces in vertex order.

    ``solid_lines``: also treat ``subtype == "solid"`` markings as the road edge; default = only when the map
    has no barrier-type way at all (the highD maps are nothing but ``line_thin`` solid/dashed markings)."""
    def blocking(w):
        return (w.type_ in barrier_types) or (w.subtype in barrier_types)

    if solid_lines is None:
        solid_lines = not any(blocking(w) for w in map_.ways)
    segs = []
    for w in map_.ways:
        if not (blocking(w) or (solid_lines and w.subtype == "solid")):
            continue
        for a, b in zip(w.points[:-1], w.points[1:]):
            segs.append((a[0], a[1], b[0], b[1]))
    return np.asarray(segs, dtype=np.float32).reshape(-1, 4)


def compile_tiles(src_root: str, out_dir: str = TILE_DIR) -> List[str]:
    """Compile every ``*.osm`` below ``src_root`` into ``out_dir/<name>.npz`` (segments, bounds, counts)."""
    os.makedirs(out_dir, exist_ok=True)
    written = []
    for d, _, files in sorted(os.walk(src_root)):
        for f in sorted(files):
            if not f.endswith(".osm"):
                continue
            m = parse_osm_lanelet2(os.path.join(d, f))
            seg = collidable_segments(m)
            n_all = int(sum(max(0, len(w.points) - 1) for w in m.ways))
            name = f[:-4]
            np.savez(os.path.join(out_dir, name + ".npz"), segments=seg, bounds=np.asarray(m.boundary, dtype=np.float32),
                     n_all_segments=n_all, n_ways=len(m.ways))
            written.append(name)
    return written


def load_collidable_segments(name: str):
    """(segments float32 [S, 4], bounds (xmin, xmax, ymin, ymax)) of a packaged tile, e.g. ``"inD_1"``."""
    path = os.path.join(TILE_DIR, name + ".npz")
    if not os.path.exists(path):
        raise FileNotFoundError(f"map tile {name!r} not found in {TILE_DIR}; compile it with `python -m tactics2d_b200.map <osm dir>`")
    z = np.load(path)
    return np.ascontiguousarray(z["segments"], dtype=np.float32), tuple(float(v) for v in z["bounds"])


def list_tiles() -> List[str]:
    return sorted(f[:-4] for f in os.listdir(TILE_DIR) if f.endswith(".npz")) if os.path.isdir(TILE_DIR) else []


def polygons_to_segments(polygons, polylines=()):
    """Flatten static objects into the tile format of ``BatchedWorld.set_map`` / ``set_map_table``.

    ``polygons``: list of [V, 2] vertex arrays (``Area.geometry.exterior`` without the repeated closing vertex), in the
    order ``StaticCollision.reset`` receives the areas; ``polylines``: list of [V, 2] arrays (``RoadLine.geometry``), appended
    after them.  Returns ``(segments [S, 4] float32, poly_start [P + 1] int32)``: ring p = segments
    [poly_start[p], poly_start[p + 1]), each edge ending exactly where the next one starts."""
    segs, starts = [], [0]
    for poly in polygons:
        v = np.asarray(poly, dtype=np.float32).reshape(-1, 2)
        if len(v) >= 2 and np.array_equal(v[0], v[-1]):
            v = v[:-1]
        if len(v) < 3:
            raise ValueError("a polygon needs at least 3 vertices")
        nxt = np.roll(v, -1, axis=0)
        segs.append(np.concatenate([v, nxt], 1))
        starts.append(starts[-1] + len(v))
    for line in polylines:
        v = np.asarray(line, dtype=np.float32).reshape(-1, 2)
        if len(v) >= 2:
            segs.append(np.concatenate([v[:-1], v[1:]], 1))
    seg = np.concatenate(segs, 0).astype(np.float32) if segs else np.zeros((0, 4), np.float32)
    return np.ascontiguousarray(seg), np.asarray(starts, dtype=np.int32)
