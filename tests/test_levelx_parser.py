"""LevelX trajectory files -> participants and initial-state pools (tactics2d_b200.dataset_parser), on synthetic CSVs written in
the datasets' schema (no LevelX data ships with the reference or this image): the restated arithmetic of
tactics2d/dataset_parser/parse_levelx.py:180-333."""

import numpy as np
import pandas as pd
import pytest

from tactics2d_b200.dataset_parser import LevelXParser, initial_state_pool
from tactics2d_b200.dataset_parser.parse_levelx import _utm_northing
from tactics2d_b200.participant.element import Cyclist, Pedestrian, Vehicle


def _write_ind(folder, fid=3):
    rows = []
    tracks = {0: ("car", 4.6, 1.9, 0, 9), 1: ("truck_bus", 11.5, 2.5, 2, 9), 2: ("bicycle", 1.8, 0.6, 0, 5), 3: ("pedestrian", 0.5, 0.5, 4, 9)}
    for tid, (cls, length, width, f0, f1) in tracks.items():
        for f in range(f0, f1 + 1):
            rows.append(dict(recordingId=fid, trackId=tid, frame=f, trackLifetime=f - f0, xCenter=10.0 * tid + 0.4 * f, yCenter=-5.0 + tid + 0.1 * f,
                             heading=90.0 * tid + 1.5 * f, width=width, length=length, xVelocity=10.0 - tid, yVelocity=0.5 * tid,
                             xAcceleration=0.1, yAcceleration=-0.2))
    pd.DataFrame(rows).to_csv(folder / f"{fid:02d}_tracks.csv", index=False)
    pd.DataFrame([dict(recordingId=fid, trackId=t, initialFrame=v[3], finalFrame=v[4], numFrames=v[4] - v[3] + 1, width=v[2], length=v[1],
                       **{"class": v[0]}) for t, v in tracks.items()]).to_csv(folder / f"{fid:02d}_tracksMeta.csv", index=False)
    pd.DataFrame([dict(recordingId=fid, locationId=2, frameRate=25)]).to_csv(folder / f"{fid:02d}_recordingMeta.csv", index=False)
    return tracks


def test_ind_schema_participants_and_pool(tmp_path):
    tracks = _write_ind(tmp_path)
    p = LevelXParser("inD")
    assert p.get_location("03_tracks.csv", str(tmp_path)) == 2
    assert p.get_time_range(3, str(tmp_path)) == (0, 360)
    parts, rng = p.parse_trajectory("03", str(tmp_path))
    assert rng == (0, 360) and set(parts) == {0, 1, 2, 3}
    assert isinstance(parts[0], Vehicle) and isinstance(parts[1], Vehicle) and isinstance(parts[2], Cyclist) and isinstance(parts[3], Pedestrian)
    assert parts[1].type_ == "bus" and parts[1].length == 11.5 and parts[1].width == 2.5
    s = parts[2].trajectory.get_state(4 * 40)
    assert s.x == pytest.approx(20.0 + 1.6) and s.y == pytest.approx(-3.0 + 0.4)
    assert s.heading == pytest.approx((180.0 + 6.0) * 2 * np.pi / 360)          # degrees -> radians (:246-249)
    assert s.vx == 8.0 and s.vy == 1.0 and parts[2].trajectory.last_frame == 200
    # a time window and an id filter
    parts, rng = p.parse_trajectory(3, str(tmp_path), time_range=(100, 250), ids=[0, 3])
    assert set(parts) == {0, 3} and rng == (120, 240) and parts[3].trajectory.first_frame == 160
    # the pool: rows at 0 ms (tracks 0, 2), 200 ms (0, 1, 2, 3), 360 ms (0, 1, 3); 3 slots per row
    pool, tid, table = initial_state_pool(p, 3, str(tmp_path), 3, [0, 200, 360])
    assert tid.shape == (3, 3) and list(tid[0] != 255) == [True, True, False] and (tid[1] != 255).all()
    np.testing.assert_allclose(pool["x"][1], [2.0, 12.0, 22.0], rtol=1e-6)         # lowest ids first: 0, 1, 2
    np.testing.assert_allclose(pool["speed"][1], np.hypot([10, 9, 8], [0, 0.5, 1.0]), rtol=1e-6)
    assert 0 <= pool["heading"].min() and pool["heading"].max() < 2 * np.pi
    rows = table.rows
    assert rows[tid[1, 1]].half_len > rows[tid[1, 0]].half_len                    # the bus got a longer template than the car
    assert rows[tid[2, 2]].shape == 1                                              # track 3 at 360 ms: a pedestrian disc


def test_highd_schema_box_centre_and_calibration(tmp_path):
    fid = 7
    rows = []
    for f in range(3):
        rows.append(dict(frame=f + 1, id=5, x=100.0 + f, y=20.0, width=4.5, height=1.8, xVelocity=30.0, yVelocity=0.6, xAcceleration=0.0, yAcceleration=0.0))
        rows.append(dict(frame=f + 1, id=6, x=300.0 - f, y=9.0, width=12.0, height=2.5, xVelocity=-25.0, yVelocity=-0.2, xAcceleration=0.1, yAcceleration=0.0))
    pd.DataFrame(rows).to_csv(tmp_path / f"{fid:02d}_tracks.csv", index=False)
    pd.DataFrame([dict(id=5, width=4.5, height=1.8, initialFrame=1, finalFrame=3, **{"class": "Car"}),
                  dict(id=6, width=12.0, height=2.5, initialFrame=1, finalFrame=3, **{"class": "Truck"})]).to_csv(tmp_path / f"{fid:02d}_tracksMeta.csv", index=False)
    pd.DataFrame([dict(id=fid, locationId=1, lowerLaneMarkings="21.0;24.9;28.8", upperLaneMarkings="8.5;12.6;16.4")]).to_csv(
        tmp_path / f"{fid:02d}_recordingMeta.csv", index=False)
    p = LevelXParser("highD")
    parts, rng = p.parse_trajectory(fid, str(tmp_path))
    assert rng == (40, 120) and parts[6].type_ == "truck" and parts[6].length == 12.0 and parts[6].width == 2.5   # highD: "width" is the length
    # calibration (:116-129): northing of the two bounding latitudes against the outermost lane markings
    lo, hi = _utm_northing(0.0, -0.00025899967), _utm_northing(0.0, 0.0)
    k = (hi - lo) / (8.5 - 28.8)
    b = hi - k * 8.5
    s = parts[5].trajectory.get_state(40)
    theta = np.round(np.arctan(0.6 / 30.0), 5)
    assert s.x == pytest.approx(100.0 + 4.5 * np.cos(theta) / 2 - 1.8 * np.sin(theta) / 2)                # :258-263
    assert s.y == pytest.approx((20.0 + 4.5 * np.sin(theta) / 2 + 1.8 * np.cos(theta) / 2) * k + b)     # :264-274
    assert s.heading == pytest.approx(np.round(np.arctan2(-0.6, 30.0), 5))                                # image y axis points down (:240-245)
    s6 = parts[6].trajectory.get_state(80)
    assert s6.heading == pytest.approx(np.round(np.arctan2(0.2, -25.0), 5))
    assert abs(k) == pytest.approx(28.6668 / 20.3, rel=1e-3)     # ~ 28.67 m of northing over 20.3 m of image: the images are ~ 1.41 px/m


def test_unknown_dataset_is_rejected():
    with pytest.raises(KeyError):
        LevelXParser("nuScenes")
    with pytest.raises(TypeError):
        LevelXParser("inD")._get_file_id(3.5)
