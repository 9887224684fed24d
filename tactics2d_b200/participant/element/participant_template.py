"""Dimension / performance templates of the traffic participants.

Same data and keys as the reference's ``tactics2d/participant/element/participant_template.py``
(:42-257): ``VEHICLE_TEMPLATE`` (9 classes), ``CYCLIST_TEMPLATE`` (3), ``PEDESTRIAN_TEMPLATE`` (4)
and the EURO / NCAP / EPA alias maps (:9-40).  Stored here as compact rows and expanded to the
reference's ``{type_name: {attribute: value}}`` dictionaries.
"""

_VEHICLE_KEYS = ("length", "width", "height", "wheel_base", "front_overhang", "rear_overhang",
                 "kerb_weight", "max_speed", "0_100_km/h", "max_decel", "driven_mode")
_VEHICLE_ROWS = {
    "mini_car": (3.540, 1.641, 1.489, 2.420, 0.585, 0.535, 1070, 44.44, 14.4, 10.0, "FWD"),
    "small_car": (4.053, 1.751, 1.461, 2.548, 0.824, 0.681, 1565, 52.78, 11.2, 10.0, "FWD"),
    "medium_car": (4.284, 1.799, 1.452, 2.637, 0.880, 0.767, 1620, 69.44, 8.9, 11.0, "FWD"),
    "large_car": (4.866, 1.832, 1.477, 2.871, 0.955, 1.040, 1735, 58.33, 8.4, 11.0, "FWD"),
    "executive_car": (5.050, 1.886, 1.475, 3.024, 0.921, 1.105, 2175, 63.89, 8.1, 11.3, "FWD"),
    "luxury_car": (5.302, 1.945, 1.488, 3.128, 0.989, 1.185, 2520, 69.44, 6.7, 11.3, "AWD"),
    "sports_coupe": (4.788, 1.916, 1.381, 2.720, 0.830, 1.238, 1740, 63.89, 5.3, 10.4, "AWD"),
    "multi_purpose_car": (5.155, 1.995, 1.740, 3.090, 0.935, 1.130, 2095, 66.67, 9.4, 10.3, "4WD"),
    "sports_utility_car": (4.828, 1.943, 1.792, 2.915, 0.959, 0.954, 2200, 88.89, 3.8, 10.29, "4WD"),
}
VEHICLE_TEMPLATE = {k: dict(zip(_VEHICLE_KEYS, row)) for k, row in _VEHICLE_ROWS.items()}

_CYCLIST_KEYS = ("length", "width", "height", "max_steer", "max_speed", "max_accel", "max_decel")
CYCLIST_TEMPLATE = {
    "cyclist": dict(zip(_CYCLIST_KEYS, (1.80, 0.65, 1.70, 1.05, 22.78, 5.8, 7.8))),
    "moped": dict(zip(_CYCLIST_KEYS, (2.00, 0.70, 1.70, 0.35, 13.89, 3.5, 7.0))),
    "motorcycle": dict(zip(_CYCLIST_KEYS, (2.40, 0.80, 1.70, 0.44, 75.00, 5.0, 10.0))),
}

_PEDESTRIAN_KEYS = ("length", "width", "height", "max_speed", "max_accel")
PEDESTRIAN_TEMPLATE = {
    "adult_male": dict(zip(_PEDESTRIAN_KEYS, (0.24, 0.40, 1.75, 7.0, 1.5))),
    "adult_female": dict(zip(_PEDESTRIAN_KEYS, (0.22, 0.37, 1.65, 6.0, 1.5))),
    "children_six_year_old": dict(zip(_PEDESTRIAN_KEYS, (0.18, 0.25, 1.16, 3.5, 1))),
    "children_ten_year_old": dict(zip(_PEDESTRIAN_KEYS, (0.20, 0.35, 1.42, 4.5, 1.0))),
}

_ORDER = ("mini_car", "small_car", "medium_car", "large_car", "executive_car", "luxury_car",
          "sports_coupe", "multi_purpose_car", "sports_utility_car")
EURO_SEGMENT_MAPPING = dict(zip(("A", "B", "C", "D", "E", "F", "S", "M", "J"), _ORDER))
NCAP_MAPPING = dict(zip(("supermini", "small_family_car", "large_family_car", "executive", "large_mpv",
                         "large_off_road"),
                        ("small_car", "medium_car", "large_car", "executive_car", "multi_purpose_car",
                         "sports_utility_car")))
EPA_MAPPING = dict(zip(("minicompact", "subcompact", "compact", "midsize", "large", "two-seater",
                        "multi_purpose_car", "standard_suv"),
                       ("mini_car", "small_car", "medium_car", "large_car", "executive_car",
                        "sports_coupe", "minivan", "sports_utility_car")))


def _table(template: dict) -> str:
    keys = sorted({k for row in template.values() for k in row})
    lines = ["\t".join(["type"] + keys)]
    for name, row in template.items():
        lines.append("\t".join([name] + [str(row.get(k, "")) for k in keys]))
    return "\n".join(lines)


def list_vehicle_templates():
    """Print the vehicle templates (the reference pretty-prints with ``tabulate``)."""
    print(_table(VEHICLE_TEMPLATE))


def list_cyclist_templates():
    print(_table(CYCLIST_TEMPLATE))


def list_pedestrian_templates():
    print(_table(PEDESTRIAN_TEMPLATE))
