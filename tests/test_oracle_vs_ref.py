"""CPU, build container only: the C oracle against the unmodified reference (oracle/_ref/libbark_ref.so),
bit for bit.  Skipped where the reference library was not built."""
import numpy as np
import pytest

from conftest import bits


@pytest.fixture(scope="module")
def pair(orc, weights_file):
    if not orc.have_ref():
        pytest.skip("oracle/_ref/libbark_ref.so not built here")
    path = weights_file("tiny", "f16")
    return orc.Oracle(path, seed=0, n_steps=16), orc.Ref(path, seed=0, n_steps=16)


def test_gelu_table(orc, pair):
    o, r = orc.gelu_tables()
    assert np.array_equal(o, r)


def test_tokenizer(pair):
    o, r = pair
    for text in ["hello world", "", "Hello, world! 123 café zz", "ÀÉÎõü ñ ç", "a" * 600, "x,y;z...", "日本語 text", "tab\there"]:
        assert np.array_equal(o.tokenize(text), r.tokenize(text)), text


def test_causal_eval_bit_exact(pair):
    o, r = pair
    rng = np.random.default_rng(1)
    for which, first, merge in ((0, None, True), (1, np.concatenate([rng.integers(0, 10000, 256), [12050], rng.integers(10000, 12048, 37)]).astype(np.int32), False)):
        toks = o.tokenize("hello world") if first is None else first
        po = pr = 0
        for step in range(40):
            lo, po = o.gpt_eval(which, toks, po, merge)
            lr, pr = r.gpt_eval(which, toks, pr, merge)
            assert po == pr and np.array_equal(bits(lo), bits(lr)), (which, step)
            toks = np.array([int(np.argmax(lr[:10000])) if which == 0 else 10000 + int(np.argmax(lr[10000:12048]))], np.int32)


def test_fine_eval_bit_exact(pair):
    o, r = pair
    rng = np.random.default_rng(2)
    buf = rng.integers(0, 1024, (8, 1024)).astype(np.int32)
    for nn in (2, 7):
        x = buf.copy(); x[nn:, :] = 1024
        assert np.array_equal(bits(o.fine_eval(x, nn)), bits(r.fine_eval(x, nn)))


def test_sampler(pair):
    o, r = pair
    rng = np.random.default_rng(3)
    o.reseed(9); r.reseed(9)
    for i in range(150):
        lg = (rng.standard_normal((10048, 1024)[i % 2]) * 4).astype(np.float32)
        temp = (0.7, 0.5, 0.0)[i % 3]
        assert o.sample(lg, temp) == r.sample(lg, temp)


def test_encodec_bit_exact(pair):
    o, r = pair
    rng = np.random.default_rng(4)
    for T in (7, 40):
        codes = rng.integers(0, 1024, (8, T)).astype(np.int32)
        assert np.array_equal(bits(o.encodec_decode(codes)), bits(r.encodec_decode(codes)))


def test_full_generate(pair):
    o, r = pair
    o.reseed(0); r.reseed(0)
    a, b = o.generate("hello world"), r.generate("hello world")
    for k in ("semantic", "coarse", "fine"):
        assert np.array_equal(a[k], b[k]), k
    assert np.array_equal(bits(a["audio"]), bits(b["audio"]))
