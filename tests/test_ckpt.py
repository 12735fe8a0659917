"""TF checkpoint-v2 bundle writer/reader (alphafive_b200.ckpt) against the files the reference ships:
written from the 42 tensors of ckpt-6960, the .index is byte-identical and the .data file hashes equal."""
import hashlib

import numpy as np

from conftest import golden


def _tensors():
    z = golden("ckpt6960.npz")
    return {k.replace("__", "/"): z[k] for k in z.files}


def test_writer_reproduces_the_shipped_bundle_byte_for_byte(tmp_path):
    from alphafive_b200 import ckpt
    f = golden("ckpt6960_files.npz")
    prefix = str(tmp_path / "alphaFive-6960")
    ckpt.write_bundle(prefix, _tensors())
    assert open(prefix + ".index", "rb").read() == f["index"].tobytes()
    assert hashlib.sha256(open(prefix + ".data-00000-of-00001", "rb").read()).hexdigest() == str(f["data_sha256"])
    assert open(tmp_path / "checkpoint").read() == str(f["marker"])


def test_round_trip_other_shapes(tmp_path):
    from alphafive_b200 import ckpt
    rng = np.random.default_rng(0)
    w = {"a/kernel": rng.standard_normal((3, 3, 5, 7)).astype(np.float32), "a/bias": np.zeros(7, np.float32),
         "zz/fc/kernel": rng.standard_normal((300, 2)).astype(np.float32), "b": np.float32(3.5).reshape(())}
    ckpt.write_bundle(str(tmp_path / "m-12"), w)
    back = ckpt.read_bundle(str(tmp_path))
    assert set(back) == set(w)
    for k in w:
        assert back[k].shape == w[k].shape and (back[k] == w[k]).all(), k
    assert ckpt.crc32c(b"123456789") == 0xE3069283          # the CRC-32C check value


def test_reader_skips_non_float_entries_and_writes_atomically(tmp_path):
    """A checkpoint that also holds an int64 global_step (tf.train.Saver saves every variable): the float
    tensors load, the rest is skipped with a warning (the reference's load_pretrained skips what it cannot
    use, network.py:136-160); no temporary files are left; the marker keeps the history of prefixes."""
    import os
    import warnings
    from alphafive_b200 import ckpt
    w = {"bone/conv1/kernel": np.arange(12, dtype=np.float32).reshape(1, 2, 3, 2), "bone/conv1/bias": np.ones(2, np.float32),
         "global_step": np.array(6960, np.int64)}
    for step in (10, 20):
        ckpt.write_bundle(str(tmp_path / f"alphaFive-{step}"), w)
    with warnings.catch_warnings(record=True) as rec:
        warnings.simplefilter("always")
        got = ckpt.read_bundle(str(tmp_path))
    assert sorted(got) == ["bone/conv1/bias", "bone/conv1/kernel"] and (got["bone/conv1/kernel"] == w["bone/conv1/kernel"]).all()
    assert any("global_step" in str(r.message) for r in rec)
    assert not [f for f in os.listdir(tmp_path) if ".tmp" in f]
    marker = open(tmp_path / "checkpoint").read()
    assert marker.startswith('model_checkpoint_path: "alphaFive-20"') and '"alphaFive-10"' in marker
