"""Checkpoint helpers of the sampler base (reference ``src/aspire/samplers/base.py:158-276``): host logic only."""
import pickle

import numpy as np
import pytest


def _sampler():
    from aspire_b200.samplers.base import Sampler

    return Sampler(lambda s: 0.0, lambda s: 0.0, 2, prior_flow=None, parameters=["a", "b"])


def test_state_keys_serialisation_and_sources(tmp_path):
    s = _sampler()
    host = {"x": np.arange(6.0).reshape(3, 2), "log_likelihood": np.zeros(3)}
    st = s.build_checkpoint_state(host, iteration=4, meta={"beta": 0.25})
    assert st["sampler"] == "Sampler" and st["iteration"] == 4 and st["meta"] == {"beta": 0.25}
    assert st["parameters"] == ["a", "b"] and st["config"]["sampler_class"] == "Sampler"
    blob = s.serialize_checkpoint(st)
    assert pickle.loads(blob)["iteration"] == 4
    s.default_checkpoint_callback(st)
    assert s.last_checkpoint_state is st and s.last_checkpoint_bytes == blob
    cb = s.default_file_checkpoint_callback(tmp_path / "c.pkl")
    cb(st)
    for src in (st, blob, str(tmp_path / "c.pkl")):
        saved, state = s.restore_from_checkpoint(src)
        np.testing.assert_array_equal(saved["x"], host["x"])
        assert state["meta"]["beta"] == 0.25
    assert s.default_file_checkpoint_callback(None) == s.default_checkpoint_callback


def test_bad_sources():
    s = _sampler()
    with pytest.raises(TypeError):
        s.restore_from_checkpoint(42)
    with pytest.raises(ValueError):
        s.restore_from_checkpoint({"iteration": 1})
