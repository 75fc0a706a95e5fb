"""Dataset containers / file formats (host-side plumbing, no GPU)."""
import numpy as np
import pytest

from neorl_industrial import datasets as D


def _fake(m=37, S=12, A=3):
    rng = np.random.default_rng(0)
    return {"observations": rng.normal(size=(m, S)).astype(np.float32), "actions": rng.uniform(-1, 1, (m, A)).astype(np.float32),
            "rewards": rng.normal(size=m).astype(np.float32), "terminals": rng.random(m) < 0.1, "timeouts": np.zeros(m, bool),
            "next_observations": rng.normal(size=(m, S)).astype(np.float32), "safety": rng.integers(0, 8, m).astype(np.uint8)}


def test_fixture_layout_fields():
    ds = _fake()
    fx = D.to_fixture_layout(ds, {"env": "ChemicalReactor-v0", "quality": "mixed"})
    assert set(D.FIXTURE_FIELDS) <= set(fx) and fx["metadata"]["n_transitions"] == 37
    assert fx["safety_violations"].dtype == bool and np.array_equal(fx["safety_violations"], ds["safety"] != 0)
    with pytest.raises(KeyError):
        D.to_fixture_layout({k: v for k, v in ds.items() if k != "safety"})


@pytest.mark.parametrize("ext", [".npz", ".pt"])
def test_save_load_roundtrip(tmp_path, ext):
    fx = D.to_fixture_layout(_fake(), {"env": "PowerGrid-v0", "seed": 3})
    path = D.save_dataset(str(tmp_path / f"ds{ext}"), fx)
    back = D.load_dataset(path)
    assert back["metadata"]["env"] == "PowerGrid-v0" and back["metadata"]["seed"] == 3
    for k in D.FIXTURE_FIELDS:
        assert back[k].dtype == fx[k].dtype and np.array_equal(back[k], fx[k]), k


def test_hdf5_needs_h5py_or_works(tmp_path):
    fx = D.to_fixture_layout(_fake())
    try:
        import h5py  # noqa: F401
    except ImportError:
        with pytest.raises(ImportError, match="h5py"):
            D.save_dataset(str(tmp_path / "ds.h5"), fx)
    else:
        back = D.load_dataset(D.save_dataset(str(tmp_path / "ds.h5"), fx))
        assert np.array_equal(back["rewards"], fx["rewards"])
    with pytest.raises(ValueError):
        D.save_dataset(str(tmp_path / "ds.csv"), fx)


def test_as_torch_views():
    import torch
    t = D.as_torch(D.to_fixture_layout(_fake()))
    assert t["observations"].dtype == torch.float32 and t["terminals"].dtype == torch.bool
    assert t["observations"].shape == (37, 12)
