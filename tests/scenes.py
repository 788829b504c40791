"""Scene lookup shared by the golden-vector tests (keys of tests/golden/solves.json -> geometry)."""
from raystrack_b200 import synthetic


def scene_for(case: str):
    tag = case[:2]
    if tag == "C1":
        return synthetic.parallel_unit_squares()
    if tag in ("C2", "C3", "V0", "X1", "X2"):
        return synthetic.street_canyon()
    if tag in ("C4", "X3"):
        return synthetic.unit_cube_enclosure()
    if tag in ("U3", "X4"):
        return synthetic.urban_block(3, 4, 8, 0)
    if tag == "X5":
        return synthetic.tilted_pair()
    raise KeyError(case)


URBAN_RAY_CASES = ((0, False), (7, True), (45, False), (22, False))
