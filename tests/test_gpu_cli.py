"""Batch CLI front end on the GPU: embed a file, verify a mixed directory."""
import os
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_cli_embed_and_batch_verify(tmp_path, capsys):
    from echoseal_b200 import cli
    from _inputs import make_clip
    key = "aa" * 32
    rng = np.random.default_rng(0)
    host = (0.05 * rng.standard_normal(48000)).astype(np.float32)
    np.save(tmp_path / "host.npy", host)
    assert cli.main(["embed", "--key", key, str(tmp_path / "host.npy"), str(tmp_path / "wm.npy")]) == 0
    wm = np.load(tmp_path / "wm.npy")
    assert wm.shape == host.shape and wm.dtype == np.float32 and np.abs(wm - host).max() > 1e-3
    cli.write_wav(str(tmp_path / "wm16.wav"), wm, 48000)
    audio, _ = make_clip("short_1s")
    np.save(tmp_path / "golden.npy", audio)
    np.save(tmp_path / "empty.npy", np.zeros(0, np.float32))
    assert cli.main(["verify", "--key", key, str(tmp_path)]) == 0
    out = capsys.readouterr().out.strip().splitlines()
    assert len(out) == 5 and all("tampered / no watermark" in l for l in out)    # the reference verdict (False)
    with pytest.raises(SystemExit):
        cli.main(["verify", "--key", "ab" * 16, str(tmp_path)])       # 128-bit key (rx_app.py:24-25)
