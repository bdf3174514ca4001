"""BASELINE.json configs[0] at its true size: bark-small dimensions, f32 GPT + f16 codec, one prompt, seed 0, the reference's CPU
path at -t 4.  tests/golden/small_f32_n12.npz holds the reference's token ids and waveform for n_steps_text_encoder = 12 (a 15 s CPU
run; made by tests/golden/make_golden_small.py; the 1.6 GB weight file comes from bark.cpp_b200/weights.py and is not committed).

The CPU-oracle check takes 3 minutes and stays opt-in (BARK_B200_SLOW_TESTS=1); the CUDA check is part of the default `-m gpu` run."""
import hashlib
import os

import numpy as np
import pytest

from conftest import GOLDEN_DIR, bits

slow = pytest.mark.skipif(os.environ.get("BARK_B200_SLOW_TESTS") != "1", reason="opt-in: BARK_B200_SLOW_TESTS=1")


def golden():
    return np.load(os.path.join(GOLDEN_DIR, "small_f32_n12.npz"))


def file_sha1(path):
    h = hashlib.sha1()
    with open(path, "rb") as f:
        for blk in iter(lambda: f.read(1 << 24), b""):
            h.update(blk)
    return h.hexdigest()


def test_fixture_is_well_formed():
    g = golden()
    assert str(g["config"]) == "small" and str(g["ftype"]) == "f32" and int(g["seed"]) == 0 and int(g["n_steps"]) == 12
    assert g["semantic"].shape == (12,) and g["coarse"].shape == (18, 2) and g["fine"].shape == (18, 8) and g["audio"].shape == (18 * 320,)
    assert "AVX2" in str(g["reference_build"]) or "mavx2" in str(g["reference_build"])


@slow
def test_oracle_reproduces_the_reference_at_bark_small_f32(orc, weights_file):
    g = golden()
    path = weights_file("small", "f32", int(g["weight_seed"]))
    assert file_sha1(path) == str(g["weights_sha1"])
    r = orc.Oracle(path, seed=0, n_steps=12).generate(str(g["prompt"]))
    for k in ("semantic", "coarse", "fine"):
        assert np.array_equal(r[k], g[k]), k
    assert np.array_equal(bits(r["audio"]), bits(g["audio"]))


@pytest.mark.gpu
def test_cuda_path_reproduces_the_reference_at_bark_small_f32(pkg, weights_file):
    g = golden()
    path = weights_file("small", "f32", int(g["weight_seed"]))
    assert file_sha1(path) == str(g["weights_sha1"])
    with pkg.Bark(path, seed=0, n_steps_text_encoder=12) as b:
        assert np.array_equal(b.tokenize(str(g["prompt"])), g["prompt_ids"])
        audio = b.generate(str(g["prompt"]))
        assert np.array_equal(b.tokens(0), g["semantic"])
        assert np.array_equal(b.tokens(1), g["coarse"])
        assert np.array_equal(b.tokens(2), g["fine"])
        assert audio.shape == g["audio"].shape
        assert float(np.abs(audio - g["audio"]).max() / np.abs(g["audio"]).max()) < 1e-3      # north_star: within 1e-3 relative
