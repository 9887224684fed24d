"""CPU-side host logic: State / Trajectory / participants / templates / type table / map tiles, written to read
like the reference's own tests (tests/test_participant.py, tests/test_physics.py)."""

import ctypes
import os

import numpy as np
import pytest

from tactics2d_b200 import TypeParams, TypeTable, _lib
from tactics2d_b200.map import collidable_segments, list_tiles, load_collidable_segments
from tactics2d_b200.participant.element import Cyclist, Obstacle, Other, Pedestrian, Vehicle
from tactics2d_b200.participant.element.participant_template import (CYCLIST_TEMPLATE, EPA_MAPPING, EURO_SEGMENT_MAPPING,
                                                                    NCAP_MAPPING, PEDESTRIAN_TEMPLATE, VEHICLE_TEMPLATE)
from tactics2d_b200.participant.trajectory import State, Trajectory
from tactics2d_b200.physics import PointMass, SingleTrackDynamics, SingleTrackKinematics
from tactics2d_b200.traffic import ScenarioStatus, TrafficStatus

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ---------------------------------------------------------------- State (reference tests/test_participant.py:158-191, 498-541)
def test_state_setters():
    state = State(frame=0, x=5.0, y=6.0, heading=0.5)
    state.x = 10.0
    state.y = 20.0
    assert state.location == (10.0, 20.0)
    state.set_heading(1.0)
    assert state.heading == 1.0
    state.set_velocity(2.0, 3.0)
    assert (state.vx, state.vy) == (2.0, 3.0)
    assert state.speed == pytest.approx(3.605551275463989)
    state.set_speed(5.0)
    assert state._speed == 5.0
    state.set_accel(1.0, 2.0)
    assert (state.ax, state.ay) == (1.0, 2.0)
    assert state.accel == pytest.approx(2.23606797749979)


def test_state_cache_invalidation():
    state = State(frame=0, x=5.0, y=6.0, heading=0.5, vx=2.0, vy=3.0)
    assert state.speed == pytest.approx(3.605551275463989)
    state._speed = None
    state.vx = 4.0
    state.vy = 0.0
    assert state.speed == 4.0
    state.ax = 1.0
    state.ay = 2.0
    assert state.accel == pytest.approx(2.23606797749979)
    state._accel = None
    state.ax = 3.0
    state.ay = 4.0
    assert state.accel == pytest.approx(5.0)
    state.set_speed(10.0)
    assert state.speed == 10.0
    state.vx = 1.0
    state.vy = 1.0
    assert state.speed == 10.0   # the stored scalar wins
    state._speed = None
    assert state.speed == pytest.approx(1.4142135623730951)


def test_state_types_and_derived():
    s = State(frame=3.0, x=1, y=2, heading=0)
    assert isinstance(s.frame, int) and isinstance(s.x, float)
    with pytest.raises(ValueError):
        State(frame=0, x="abc")
    s = State(0, 0, 0, np.pi / 2, speed=2.0)
    vx, vy = s.velocity
    assert vx == pytest.approx(0.0, abs=1e-12) and vy == pytest.approx(2.0)
    s = State(0, 0, 0, 0.0, accel=-3.0)
    assert s.acceleration == (-3.0, -0.0) and s.accel == 3.0   # accel is the norm, not the stored scalar
    assert State(0).speed is None and State(0).velocity is None and State(0).accel is None


def test_trajectory_edge_cases():
    t = Trajectory(id_=0)
    assert t.frames == [] and t.history_states == {} and t.initial_state is None and t.last_state is None
    assert t.first_frame is None and t.last_frame is None and np.isnan(t.average_speed)
    with pytest.raises(KeyError):
        t.get_state(0)
    assert t.get_trace() == []
    with pytest.raises(ValueError, match="not a valid State object"):
        t.add_state(None)
    t.add_state(State(0, 0, 0, 0, speed=1.0))
    t.add_state(State(100, 1, 0, 0, speed=1.0))
    t.add_state(State(200, 2, 0, 0, speed=3.0))
    assert t.stable_freq and len(t) == 3 and t.average_speed == pytest.approx(5 / 3)
    t.add_state(State(350, 3, 0, 0, speed=1.0))
    assert not t.stable_freq
    with pytest.raises(KeyError):
        t.add_state(State(300, 0, 0, 0))
    assert t.get_trace((100, 200)) == [(1.0, 0.0), (2.0, 0.0)]
    t.reset()
    assert len(t) == 1 and t.current_state.frame == 0
    t.reset(State(5, 9, 9, 0))
    assert t.frames == [5]


# ---------------------------------------------------------------- templates / participants
def test_templates_match_reference_values():
    assert len(VEHICLE_TEMPLATE) == 9 and len(CYCLIST_TEMPLATE) == 3 and len(PEDESTRIAN_TEMPLATE) == 4
    m = VEHICLE_TEMPLATE["medium_car"]
    assert (m["length"], m["width"], m["height"], m["front_overhang"], m["rear_overhang"]) == (4.284, 1.799, 1.452, 0.880, 0.767)
    assert (m["kerb_weight"], m["max_speed"], m["0_100_km/h"], m["max_decel"]) == (1620, 69.44, 8.9, 11.0)
    assert EURO_SEGMENT_MAPPING["C"] == "medium_car" and EPA_MAPPING["compact"] == "medium_car"
    assert NCAP_MAPPING["small_family_car"] == "medium_car" and NCAP_MAPPING["large_mpv"] == "multi_purpose_car"
    assert PEDESTRIAN_TEMPLATE["adult_male"]["width"] == 0.40 and CYCLIST_TEMPLATE["moped"]["max_steer"] == 0.35


def test_vehicle_defaults_and_template():
    v = Vehicle(id_=0)
    assert v.max_steer == 0.524 and v.speed_range == (-16.67, 55.56) and v.accel_range == (-3.0, 3.0)
    v.load_from_template("medium_car")
    assert v.length == 4.284 and v.max_accel == 3.121 and v.accel_range == (-11.0, 3.121) and v.speed_range == (-16.67, 69.44)
    v.load_from_template("C")   # EURO alias
    assert v.length == 4.284
    p = v.type_params()
    assert p.lf == 4.284 / 2 - 0.880 and p.lr == 4.284 / 2 - 0.767 and p.half_len == 2.142   # tests/test_physics.py:276-283
    v.add_state(State(0, 10.0, 5.0, np.pi / 2, speed=1.0))
    pose = v.get_pose()
    np.testing.assert_allclose(pose, [[10 + 0.8995, 5 + 2.142], [10 - 0.8995, 5 + 2.142], [10 - 0.8995, 5 - 2.142], [10 + 0.8995, 5 - 2.142]], atol=1e-12)
    v2 = Vehicle(id_=1, driven_mode="XYZ")
    assert v2.driven_mode == "FWD"
    v3 = Vehicle(id_=2, length="not a number")
    assert v3.length is None   # unconvertible -> None with a warning


def test_cyclist_pedestrian_other_obstacle():
    c = Cyclist(id_=0, type_="moped", verify=True)
    assert c.length == 2.0 and c.speed_range == (0, 13.89) and c.physics_model.lf == 1.0
    p = Pedestrian(id_=1)
    assert p.physics_model.speed_range == [0.0, 7.0]   # (-7, 7) normalised by PointMass
    p.add_state(State(0, 1.0, 2.0, 0.0, vx=1.0, vy=0.0))
    assert p.get_pose() == ((1.0, 2.0), 0.2) and p.geometry == 0.2
    assert p.type_params().radius == 0.2
    o = Other(id_=2, length=2.0)
    np.testing.assert_allclose(o.geometry, [[1, -1], [1, 1], [-1, 1], [-1, -1]])
    assert Other(id_=3).geometry is None
    with pytest.raises(TypeError):
        o.bind_trajectory(None)
    ob = Obstacle(id_=4, length=2.0, width=1.0)
    ob.add_state(State(0, 0, 0, 0))
    ob.add_state(State(100, 1, 0, 0))
    assert ob.get_state(70).frame == 100 and ob.get_state(20).frame == 0
    assert ob.is_active(50) and not ob.is_active(150)


# ---------------------------------------------------------------- physics facades (host side)
def test_physics_constructors_and_verify_state():
    k = SingleTrackKinematics(lf=1.262, lr=1.375, steer_range=0.5, speed_range=(-1, 2), accel_range=3, interval=100, delta_t=0)
    assert k.steer_range == [-0.5, 0.5] and k.speed_range == [-1.0, 2.0] and k.accel_range is None   # an int is not a float
    assert k.delta_t == 1 and k.wheel_base == pytest.approx(2.637)
    assert SingleTrackKinematics(1, 1, delta_t=50, interval=9).delta_t == 9
    d = SingleTrackDynamics(1.262, 1.375, 1620, 0.726)
    assert (d.mu, d.I_z, d.cf, d.cr) == (0.7, 1500, 20.89, 20.89) and d.type_params().model == 1
    pm = PointMass(speed_range=(-7, 7), accel_range=1.5, backend="nonsense")
    assert pm.speed_range == [0.0, 7.0] and pm.accel_range == [0.0, 1.5] and pm.backend == "newton"
    g = np.load(os.path.join(ROOT, "tests", "golden", "verify_state.npz"))
    m = SingleTrackKinematics(lf=4.284 / 2 - 0.880, lr=4.284 / 2 - 0.767, steer_range=(-0.524, 0.524), speed_range=(-16.67, 69.44),
                              accel_range=(-11.0, 3.121))
    got = [m.verify_state(State(100, r[4], r[5], r[6], speed=r[7]), State(0, r[0], r[1], r[2], speed=r[3]), 100) for r in g["inputs"]]
    assert np.array_equal(np.array(got), g["valid"])   # truth table from the reference's verify_state
    assert pm.verify_state(State(100, 0.005, 0, 0, vx=0, vy=0), State(0, 0, 0, 0, vx=0.0, vy=0.0))
    assert not pm.verify_state(State(100, 0.05, 0, 0, vx=0, vy=0), State(0, 0, 0, 0, vx=0.0, vy=0.0))


def test_status_codes():
    assert [int(s) for s in ScenarioStatus] == [1, 2, 3, 4, 5, 6]
    assert int(TrafficStatus.COLLISION_STATIC) == 3 and int(TrafficStatus.COLLISION_DYNAMIC) == 4 and len(TrafficStatus) == 10


# ---------------------------------------------------------------- type table / ABI
def test_type_table_and_c_layout():
    t = TypeTable.from_templates()
    assert len(t) == 16 and t.index("medium_car") == 2 and t.rows[12].shape == 1 and t.rows[9].lf == 0.9
    arr = t.to_c_array()
    assert ctypes.sizeof(arr) == 16 * 92 and arr[2].accel_hi == np.float32(3.121)
    o = t.as_oracle_table()
    assert o["half_len"][2] == np.float64(np.float32(2.142))
    with pytest.raises(ValueError):
        TypeTable([])


def test_abi_library_exports_every_declared_symbol():
    """The C-ABI library loads and exports every symbol include/t2d_b200.h declares (no compute without a GPU)."""
    import re

    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "t2d_b200.h")).read()
    declared = set(re.findall(r"\b(t2d_[a-z_]+)\s*\(", header))
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.t2d_version() == 100
    # argument validation happens before any CUDA call
    cfg = _lib.Config(0, 5, 0, 0)
    ctx = ctypes.c_void_p()
    assert lib.t2d_create(ctypes.byref(ctx), 0, 4, 4, ctypes.byref(cfg)) == -1
    assert b"interval_ms" in lib.t2d_last_error()
    cfg = _lib.Config(100, 5, 0, 0)
    assert lib.t2d_create(ctypes.byref(ctx), 0, 4, 129, ctypes.byref(cfg)) == -3   # T2D_E_UNSUPPORTED


# ---------------------------------------------------------------- map tiles
def test_map_tiles():
    tiles = list_tiles()
    assert len(tiles) == 13 and "inD_4" in tiles
    seg, b = load_collidable_segments("inD_4")
    assert seg.shape == (344, 4) and seg.dtype == np.float32 and b[0] < b[1] and b[2] < b[3]
    seg, b = load_collidable_segments("highD_1")
    assert len(seg) == 4 and b == (0.0, 668.0, -29.0, 0.0)
    with pytest.raises(FileNotFoundError):
        load_collidable_segments("nowhere")


def test_osm_parser_rules(tmp_path):
    osm = tmp_path / "m.osm"
    osm.write_text("""<?xml version='1.0'?><osm>
      <node id='1' lat='0.0' lon='0.0'/><node id='2' lat='0.0' lon='0.001'/><node id='3' lat='0.001' lon='0.001'/>
      <node id='4' lat='0.002' lon='0.0' action='delete'/>
      <way id='10'><nd ref='1'/><nd ref='2'/><nd ref='3'/><tag k='type' v='curbstone'/><tag k='subtype' v='low'/></way>
      <way id='11'><nd ref='1'/><nd ref='3'/><tag k='type' v='line_thin'/><tag k='subtype' v='dashed'/></way>
      <way id='12' action='delete'><nd ref='1'/><nd ref='2'/><tag k='type' v='wall'/></way></osm>""")
    from tactics2d_b200.map import parse_osm_lanelet2

    m = parse_osm_lanelet2(str(osm))
    assert len(m.nodes) == 3 and len(m.ways) == 2
    assert m.nodes[2] == pytest.approx((111.32, 0.0)) and m.nodes[3] == pytest.approx((111.32, 110.54))
    seg = collidable_segments(m)
    assert seg.shape == (2, 4) and seg[1] == pytest.approx([111.32, 0.0, 111.32, 110.54])
    assert m.boundary == (0.0, 112.0, 0.0, 111.0)


def test_osm_area_relations_chain_into_rings(tmp_path):
    """``_load_area_lanelet2`` (parse_osm.py:461-510): outer ways are chained end to end whichever way round they were drawn
    (:48-60), inner ways start a new hole whenever the current one closes, deleted relations are skipped, ways that do not
    touch raise SyntaxError."""
    from tactics2d_b200.map import Area, parse_osm_lanelet2, polygons_to_segments

    def nodes(pts):
        return "".join(f"<node id='{i + 1}' lat='{la}' lon='{lo}'/>" for i, (la, lo) in enumerate(pts))

    def way(id_, refs):
        return f"<way id='{id_}'>" + "".join(f"<nd ref='{r}'/>" for r in refs) + "<tag k='type' v='virtual'/></way>"

    # outer square 1-2-3-4 drawn as four ways in mixed directions; holes 5-6-7 (one way, closed) and 8-9-10 (two ways)
    pts = [(0, 0), (0, 0.001), (0.001, 0.001), (0.001, 0), (0.0002, 0.0002), (0.0002, 0.0004), (0.0004, 0.0003),
           (0.0006, 0.0006), (0.0006, 0.0008), (0.0008, 0.0007), (0.01, 0.01), (0.01, 0.011)]
    body = nodes(pts) + way(20, [1, 2]) + way(21, [3, 2]) + way(22, [3, 4]) + way(23, [1, 4]) + way(24, [5, 6, 7, 5]) \
        + way(25, [8, 9]) + way(26, [8, 10, 9]) + way(27, [11, 12])
    rel = ("<relation id='30'><member type='way' ref='20' role='outer'/><member type='way' ref='21' role='outer'/>"
           "<member type='way' ref='22' role='outer'/><member type='way' ref='23' role='outer'/>"
           "<member type='way' ref='24' role='inner'/><member type='way' ref='25' role='inner'/><member type='way' ref='26' role='inner'/>"
           "<member type='relation' ref='99' role='regulatory_element'/>"
           "<tag k='type' v='multipolygon'/><tag k='subtype' v='vegetation'/></relation>"
           "<relation id='31' action='delete'><member type='way' ref='20' role='outer'/><tag k='type' v='multipolygon'/></relation>"
           "<relation id='32'><member type='way' ref='20' role='left'/><member type='way' ref='22' role='right'/><tag k='type' v='lanelet'/></relation>")
    osm = tmp_path / "a.osm"
    osm.write_text("<?xml version='1.0'?><osm>" + body + rel + "</osm>")
    m = parse_osm_lanelet2(str(osm))
    assert [a.id_ for a in m.areas] == [30]
    a = m.areas[0]
    assert (a.type_, a.subtype, a.closed) == ("multipolygon", "vegetation", True)
    assert a.outer.shape == (4, 2) and [h.shape for h in a.inners] == [(3, 2), (3, 2)]
    corner = {tuple(np.round(p, 6)) for p in a.outer}
    assert corner == {(0.0, 0.0), (111.32, 0.0), (111.32, 110.54), (0.0, 110.54)}
    d = np.abs(np.roll(a.outer, -1, 0) - a.outer)
    assert ((d[:, 0] < 1e-9) ^ (d[:, 1] < 1e-9)).all()                              # consecutive vertices share a side
    seg, ps = polygons_to_segments(m.areas, [w.points for w in m.ways if w.id_ == 27])
    assert ps.tolist() == [0, 10] and len(seg) == 11
    for r0, r1 in ((0, 4), (4, 7), (7, 10)):                                        # every ring closes on itself
        assert np.array_equal(seg[r0:r1, 2:], np.roll(seg[r0:r1, :2], -1, axis=0))

    broken = tmp_path / "b.osm"
    broken.write_text("<?xml version='1.0'?><osm>" + body + "<relation id='40'><member type='way' ref='20' role='outer'/>"
                      "<member type='way' ref='27' role='outer'/><tag k='type' v='multipolygon'/></relation></osm>")
    with pytest.raises(SyntaxError):
        parse_osm_lanelet2(str(broken))
    open_ = tmp_path / "c.osm"
    open_.write_text("<?xml version='1.0'?><osm>" + body + "<relation id='41'><member type='way' ref='20' role='outer'/>"
                     "<member type='way' ref='21' role='outer'/><tag k='type' v='multipolygon'/></relation></osm>")
    assert parse_osm_lanelet2(str(open_)).areas[0].closed is False                  # the reference only warns (:491-492)


def test_packaged_maps_carry_their_areas():
    """The 96 Lanelet2 areas of the reference's inD / rounD maps (``data/{inD,rounD}_map/*.osm``; the highD maps have none),
    compiled into the packaged tiles: every exterior chains closed, the counts per subtype are the files', the one area with
    holes keeps them inside its exterior, and the whole list fits one map tile next to the collidable road lines."""
    from oracle.geometry import point_in_ring
    from tactics2d_b200.map import load_areas, polygons_to_segments

    per_map = {"inD_1": 6, "inD_2": 14, "inD_3": 11, "inD_4": 19, "rounD_0": 20, "rounD_1": 11, "rounD_2": 15}
    count = {}
    for name in list_tiles():
        areas = load_areas(name)
        assert len(areas) == per_map.get(name, 0), name
        for a in areas:
            assert a.closed and a.type_ == "multipolygon" and len(a.outer) >= 3
            count[a.subtype] = count.get(a.subtype, 0) + 1
            x, y = a.outer[:, 0], a.outer[:, 1]
            assert abs(0.5 * np.sum(x * np.roll(y, -1) - np.roll(x, -1) * y)) > 0.5          # a real region, m^2
            ring = np.concatenate([a.outer, np.roll(a.outer, -1, 0)], 1)
            for h in a.inners:
                assert point_in_ring(h[:, 0], h[:, 1], ring).all()
        if areas:
            lines, bounds = load_collidable_segments(name)
            seg, ps = polygons_to_segments(areas, [])
            assert len(ps) == len(areas) + 1 and ps[-1] == len(seg)
            assert len(seg) + len(lines) <= 32767                                            # T2D_MAX_SEGMENTS
            assert seg[:, [0, 2]].min() >= bounds[0] and seg[:, [0, 2]].max() <= bounds[1]
            assert seg[:, [1, 3]].min() >= bounds[2] and seg[:, [1, 3]].max() <= bounds[3]
    assert count == {"freespace": 23, "keepout": 9, "parking": 20, "traffic_island": 17, "vegetation": 20, "walkway": 7}
    holes = [(a.id_, len(a.inners)) for a in load_areas("inD_2") if a.inners]
    assert holes == [(30033, 2)]
    assert [a.subtype for a in load_areas("rounD_0", subtypes=("traffic_island",))] == ["traffic_island"] * len(load_areas("rounD_0", ("traffic_island",)))
    assert {a.subtype for a in load_areas("inD_4", subtypes=("vegetation", "parking"))} == {"vegetation", "parking"}


def test_parked_inside_a_real_area_is_a_static_collision():
    """StaticCollision on the reference's own map data: a small box (a pedestrian-sized pose) standing in the middle of a
    vegetation patch or a traffic island of inD_2 / rounD_0 touches none of its edges and still ``intersects`` the Area
    (collision.py:37-43).  With the areas handed over as objects the oracle reports the containing object (or an earlier one
    that overlaps it); with the same edges as bare lines it reports nothing - and a box inside one of the walkway's holes
    is free either way."""
    from oracle import scenario as O
    from oracle.geometry import point_in_ring
    from tactics2d_b200.map import load_areas, polygons_to_segments

    table = dict(half_len=np.array([0.4], np.float32), half_wid=np.array([0.3], np.float32), radius=np.array([0.0], np.float32),
                 shape=np.array([0], np.int32), model=np.array([0], np.int32))
    for k in O.TABLE_FLOAT_FIELDS:
        table.setdefault(k, np.array([1.0], np.float32))
    rng = np.random.default_rng(11)
    n_inside = n_hole = 0
    for name in ("inD_2", "rounD_0"):
        areas = load_areas(name)
        seg, ps = polygons_to_segments(areas)
        a64 = seg.astype(np.float64)
        for k, a in enumerate(areas):
            edges = a64[ps[k]:ps[k + 1]]
            lo, hi = a.outer.min(0), a.outer.max(0)
            pts = rng.uniform(lo, hi, size=(600, 2))
            # distance of each point to every edge of this area
            d = pts[:, None, :] - edges[None, :, :2]
            e = edges[None, :, 2:] - edges[None, :, :2]
            t = np.clip((d * e).sum(-1) / np.maximum((e * e).sum(-1), 1e-30), 0, 1)
            clear = np.sqrt(((d - t[..., None] * e) ** 2).sum(-1)).min(1) > 0.6          # further than the box's circumradius (0.5)
            inside = point_in_ring(pts[:, 0], pts[:, 1], edges)
            ring0 = np.concatenate([a.outer, np.roll(a.outer, -1, 0)], 1)
            in_hole = point_in_ring(pts[:, 0], pts[:, 1], ring0) & ~inside & clear
            for sel, want_hit in ((inside & clear, True), (in_hole, False)):
                q = pts[sel][:6]
                if len(q) == 0:
                    continue
                x, y = q[None, :, 0].copy(), q[None, :, 1].copy()
                h = rng.uniform(-np.pi, np.pi, x.shape)
                tid = np.zeros(x.shape, np.uint8)
                fl, _, hs = O.events(x, y, h, tid, table, seg, None, poly_start=ps)
                plain = O.events(x, y, h, tid, table, seg, None)[2]
                if want_hit:
                    assert (fl & 2).all() and (hs >= 0).all() and (hs <= ps[k]).all(), (name, a.id_)
                    assert np.isin(hs, ps[:-1]).all()                                   # an object's FIRST segment
                    own = plain < 0                                                     # clear of every edge of every area
                    assert (hs[own] <= ps[k]).all()
                    n_inside += int(own.sum())
                else:
                    later = [j for j in range(len(areas)) if j != k]
                    others = np.zeros(x.shape, bool)
                    for j in later:
                        others |= point_in_ring(x, y, a64[ps[j]:ps[j + 1]])
                    free = (plain < 0) & ~others
                    assert (hs[free] == -1).all(), (name, a.id_)
                    n_hole += int(free.sum())
    assert n_inside >= 100 and n_hole >= 4, (n_inside, n_hole)


def test_polygons_to_segments_and_pose_recovery():
    """Host helpers of round 2: static objects -> tile format; a get_pose() ring -> (centre, heading, half extents)."""
    from tactics2d_b200.map import polygons_to_segments
    from tactics2d_b200.participant.element import Vehicle
    from tactics2d_b200.participant.trajectory import State
    from tactics2d_b200.traffic.event_detection.detectors import _pose_to_rect

    seg, ps = polygons_to_segments([[(0, 0), (4, 0), (4, 3), (0, 3), (0, 0)], [(10, 10), (12, 10), (11, 12)]], [[(20, 0), (22, 0), (22, 5)]])
    assert ps.tolist() == [0, 4, 7] and seg.shape == (9, 4) and seg.dtype == np.float32
    for p0, p1 in zip(ps[:-1], ps[1:]):                       # rings chain and close
        ring = seg[p0:p1]
        assert np.array_equal(ring[:, 2:], np.roll(ring[:, :2], -1, axis=0))
    assert np.array_equal(seg[7], [20, 0, 22, 0]) and np.array_equal(seg[8], [22, 0, 22, 5])      # the open line keeps its pieces
    with pytest.raises(ValueError):
        polygons_to_segments([[(0, 0), (1, 1)]])
    v = Vehicle(id_=0)
    v.load_from_template("medium_car")
    for h in (0.3, 2.9, 4.0, 6.1):
        v.add_state(State(frame=int(h * 1000), x=12.5, y=-3.25, heading=h, speed=0.0))
        cx, cy, hh, hl, hw = _pose_to_rect(v.get_pose())
        assert abs(cx - 12.5) < 1e-12 and abs(cy + 3.25) < 1e-12
        assert abs(np.angle(np.exp(1j * (hh - h)))) < 1e-12
        assert abs(hl - v.length / 2) < 1e-12 and abs(hw - v.width / 2) < 1e-12

    class _ShapelyLike:                                       # anything with .exterior.coords (a shapely Polygon) works too
        class exterior:
            coords = [tuple(p) for p in v.get_pose()] + [tuple(v.get_pose()[0])]

    assert np.allclose(_pose_to_rect(_ShapelyLike()), _pose_to_rect(v.get_pose()))
