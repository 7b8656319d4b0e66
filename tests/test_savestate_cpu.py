"""gridfluidsim3d_b200/savestate.py against a state file written by the unmodified reference
(tests/golden/reference_small.state, oracle/make_golden.py --only-state).  CPU only."""
import os

import numpy as np
import pytest

from gridfluidsim3d_b200 import savestate, synth

GOLD = os.path.join(os.path.dirname(__file__), "golden", "reference_small.state")


def test_reference_state_fixture_reads():
    st = savestate.read_state(GOLD)
    assert st["dims"] == (10, 8, 9) and st["dx"] == 0.25 and st["frame"] == 0
    assert st["pos"].shape == st["vel"].shape == (52, 3) and st["pos"].dtype == np.float32
    assert len(st["diffuse_pos"]) == 0 and st["brick_blob"] is None
    I, J, K = st["dims"]
    border = synth.border_material(st["dims"]).reshape(K, J, I) == synth.SOLID
    mat = savestate.material_from_state(st).reshape(K, J, I)
    assert (mat[border] == synth.SOLID).all()                                  # the simulator's solid border
    extra = np.argwhere((mat == synth.SOLID) & ~border)
    assert sorted(map(tuple, extra[:, ::-1].tolist())) == [(2, 5, 6), (6, 2, 3), (6, 3, 3), (7, 2, 3)]
    cell = np.floor(st["pos"].astype(np.float64) / st["dx"]).astype(int)
    assert ((cell >= 1) & (cell < np.array([I, J, K]) - 1)).all()             # particles sit in interior cells
    assert (mat[cell[:, 2], cell[:, 1], cell[:, 0]] != synth.SOLID).all()
    assert np.all(st["vel"] == 0)                                              # a freshly initialised simulator


def test_state_round_trip_is_byte_identical(tmp_path):
    st = savestate.read_state(GOLD)
    out = str(tmp_path / "copy.state")
    savestate.write_state(out, st["dims"], st["dx"], st["pos"], st["vel"], st["solid_ijk"], frame=st["frame"])
    assert open(out, "rb").read() == open(GOLD, "rb").read()


def test_bad_files_are_rejected(tmp_path):
    p = tmp_path / "short.state"
    p.write_bytes(b"\x00" * 10)
    with pytest.raises(ValueError):
        savestate.read_state(str(p))
    raw = open(GOLD, "rb").read()
    q = tmp_path / "cut.state"
    q.write_bytes(raw[:-100])
    with pytest.raises(ValueError):
        savestate.read_state(str(q))
