"""the callers either side of the hot path -- box-changing routines
(sep_compress_box, sep_berendsen, sep_berendsen_iso), sep_relax_temp and sep_force_x0 / sep_spring_x0 --
through the sep_* API on host buffers, against what the reference build recorded for the same loops
(tests/golden/next_rows.npz, written by tests/golden/make_golden.py with the drivers in tests/common.py).

Tolerances: sums 1e-9 relative per step (rounding-order noise grows with the step count),
final positions/velocities 1e-7 absolute after 12-30 steps; box lengths, volume and cell counts 1e-12 / exact.
"""
import os

import numpy as np
import pytest

import common as cm
from seplib_b200 import capi

pytestmark = pytest.mark.gpu

G = np.load(os.path.join(cm.GOLDEN, "next_rows.npz"))


def _check(rec, prefix, scalar_cols, exact_cols=(), xtol=1e-7):
    ref = G[prefix + "_traj"]
    got = rec["traj"]
    assert got.shape == ref.shape
    for c in scalar_cols:
        scale = np.abs(ref[:, c]).max()
        assert np.abs(got[:, c] - ref[:, c]).max() <= 1e-9 * scale * len(ref), (prefix, c, np.abs(got[:, c] - ref[:, c]).max())
    for c in exact_cols:
        assert np.array_equal(got[:, c], ref[:, c]), (prefix, c)
    assert np.abs(rec["length"] - G[prefix + "_length"]).max() <= 1e-12 * np.abs(G[prefix + "_length"]).max()
    assert np.array_equal(rec["nsubbox"], G[prefix + "_nsubbox"])
    assert abs(rec["volume"] - float(G[prefix + "_volume"])) <= 1e-12 * float(G[prefix + "_volume"])
    L = rec["length"]
    dx = rec["x"] - G[prefix + "_x"]
    dx -= L * np.round(dx / L)                       # an atom within rounding of a face may sit on either side
    assert np.abs(dx).max() <= xtol, (prefix, np.abs(dx).max())
    assert np.abs(rec["v"] - G[prefix + "_v"]).max() <= xtol * 10, (prefix, np.abs(rec["v"] - G[prefix + "_v"]).max())


@pytest.mark.parametrize("sync", [1, 0])
def test_compress_box_list_mode(sync):
    lib = capi.load()
    lib.sep_gpu_set_sync(sync)
    rec = cm.drive_compress(lib, G["c_x0"], G["c_v0"], float(G["c_L"]))
    lib.sep_gpu_set_sync(1)
    _check(rec, "compress", scalar_cols=(0, 1, 2, 4), exact_cols=(3,))
    assert rec["traj"][0, 3] == 4 and rec["traj"][-1, 3] == 3      # the grid really changed on the way


def test_berendsen_z_brute_and_iso_list():
    lib = capi.load()
    lib.sep_gpu_set_sync(1)
    rec = cm.drive_berendsen(lib, G["b_x0"], G["b_v0"], float(G["b_L"]))
    _check(rec, "ber", scalar_cols=(0, 1, 2, 3, 4), exact_cols=(5,))
    rec = cm.drive_berendsen(lib, G["c_x0"], G["c_v0"], float(G["c_L"]), steps=12, iso=True, update=capi.SEP_LLIST_NEIGHBLIST)
    _check(rec, "beriso", scalar_cols=(0, 1, 2, 3, 4), exact_cols=(5,))


@pytest.mark.parametrize("sync", [1, 0])
def test_slit_pore_relax_temp_and_tethers(sync):
    lib = capi.load()
    lib.sep_gpu_set_sync(sync)
    rec = cm.drive_slit(lib, G["c_x0"], G["c_v0"], float(G["c_L"]))
    lib.sep_gpu_set_sync(1)
    assert np.array_equal(rec["types"], G["slit_types"])
    _check(rec, "slit", scalar_cols=(0, 1))
