"""hyperion/model/tests/test_amr_checks.py restated: the AMR consistency checks of setup_grid_geometry
(src/grid/grid_geometry_amr.f90:239-314) stop the run with the reference's messages, and the log shows
them wrapped exactly as error() of fortranlib/src/lib_messages.f90:126-179 wraps them (the reference's
tests search the log for the wrapped text)."""
import io

import numpy as np
import pytest

from hyperion_b200 import runner, synthetic as syn
from hyperion_b200.flatmodel import FlatConf, FlatModel, FlatSource

WRAPPED = {
    "widths": "Grids 1 and 2 in level 1 have differing cell widths in the %s \n           direction ( 5.0000E+00 and  5.0250E+00 respectively)",
    "misaligned": "Grids 1 and 2 in level 1 have edges that are not separated by \n           an integer number of cells in the %s direction",
    "refinement": "Refinement factor in the %s direction between level 1 and \n           level 2 is not an integer (1.818)",
    "across": "Grid 1 in level 2 is not aligned with cells in level 1 in the \n           %s direction",
}


def _grids(case, d):
    """The grids of test_amr_checks.py:18-179 as (n1, n2, n3, xmin, xmax, ymin, ymax, zmin, zmax)."""
    g1 = [4, 4, 4, -10., 10., -10., 10., -10., 10.]
    lo, hi = 3 + 2 * d, 4 + 2 * d
    if case == "widths":              # :18-53  same level, one grid 20.1 wide
        g2 = list(g1)
        g2[lo] = -10.1
        return [[tuple(g1), tuple(g2)]]
    if case == "misaligned":          # :58-94  same level, shifted by 0.1
        g2 = list(g1)
        g2[lo], g2[hi] = -10.1, 9.9
        return [[tuple(g1), tuple(g2)]]
    g2 = [4, 4, 4, -5., 5., -5., 5., -5., 5.]
    if case == "refinement":          # :99-136  level 2 cell 2.75 wide: 5 / 2.75 = 1.818
        g2[lo] = -6.
        return [[tuple(g1)], [tuple(g2)]]
    g2[lo], g2[hi] = -6., 4.          # :141-179  level 2 grid starts a fifth of a parent cell off
    return [[tuple(g1)], [tuple(g2)]]


def test_error_text_is_wrapped_like_the_reference():
    msgs = {"widths": "Grids 1 and 2 in level 1 have differing cell widths in the x direction (%11.4E and %11.4E respectively)" % (5.0, 5.025),
            "misaligned": "Grids 1 and 2 in level 1 have edges that are not separated by an integer number of cells in the x direction",
            "refinement": "Refinement factor in the x direction between level 1 and level 2 is not an integer (1.818)",
            "across": "Grid 1 in level 2 is not aligned with cells in level 1 in the x direction"}
    for case, msg in msgs.items():
        buf = io.StringIO()
        runner.boxed_error("setup_grid_geometry", msg, buf)
        out = buf.getvalue()
        assert (WRAPPED[case] % "x") in out
        assert out.startswith(" " + "-" * 72 + "\n ERROR   : ") and "\n WHERE   : setup_grid_geometry\n" in out
        assert "*** Execution aborted on " in out


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["widths", "misaligned", "refinement", "across"])
@pytest.mark.parametrize("d", [0, 1, 2])
def test_amr_consistency_checks(case, d):
    from hyperion_b200.capi import Engine, HyperionError
    levels = _grids(case, d)
    n = sum(g[0] * g[1] * g[2] for lev in levels for g in lev)
    dust = syn.make_dust([3.e9, 3.e16], [0.5, 0.5], [1., 1.], n_temp=10, temp_min=0.1, temp_max=1600.)
    m = FlatModel(None, None, None, np.full((1, n), 1.e-10), [dust], [FlatSource(type=1, luminosity=1., temperature=6000.)],
                  FlatConf(), grid_type="amr", amr_levels=levels)
    eng = Engine(0)
    with pytest.raises(HyperionError) as exc:
        eng.load_model(m)
    eng.close()
    buf = io.StringIO()
    runner.boxed_error("setup_grid_geometry", str(exc.value), buf)
    assert (WRAPPED[case] % "xyz"[d]) in buf.getvalue(), buf.getvalue()
