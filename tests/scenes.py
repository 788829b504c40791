"""Scene lookup shared by the golden-vector tests (keys of tests/golden/solves.json -> geometry)."""
from raystrack_b200 import synthetic


def scene_for(case: str):
    tag = case[:2]
    if tag == "C1":
        return synthetic.parallel_unit_squares()
    if tag in ("C2", "C3", "V0"):
        return synthetic.street_canyon()
    if tag == "C4":
        return synthetic.unit_cube_enclosure()
    if tag == "U3":
        return synthetic.urban_block(3, 4, 8, 0)
    raise KeyError(case)


URBAN_RAY_CASES = ((0, False), (7, True), (45, False), (22, False))
