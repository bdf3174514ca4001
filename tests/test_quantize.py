"""q4_0 files (BASELINE configs[3]): the native bark_model_quantize (csrc/quantize.cu, host only) must write the very bytes the
reference tool writes (bark.cpp:2300-2377), and the C oracle's q4_0 arithmetic (quantize_row_q8_0 + ggml_vec_dot_q4_0_q8_0 of
the pinned AVX2 build, get_rows dequantisation) must reproduce the unmodified reference on such a file bit for bit."""
import ctypes as C
import hashlib
import os

import numpy as np
import pytest

from conftest import FIXTURE_DIR, bits

GGML_FTYPE_MOSTLY_Q4_0 = 2


def q4_file(pkg, weights_file, config, src_ftype):
    src = weights_file(config, src_ftype)
    dst = os.path.join(FIXTURE_DIR, f"{config}_{src_ftype}_1234_q4_0.bin")
    if not os.path.exists(dst):
        assert pkg.lib().bark_model_quantize(src.encode(), (dst + ".tmp").encode(), GGML_FTYPE_MOSTLY_Q4_0)
        os.replace(dst + ".tmp", dst)
    return src, dst


_REF_KEEPALIVE = []


def ensure_reference_tables(orc, any_model_path):
    """ggml_init fills the f16 tables the reference's quantizer relies on (examples/quantize/main.cpp:67-72); loading one model does
    that.  Once per process: the reference never frees its ggml contexts and runs out of them after 64 loads."""
    if not _REF_KEEPALIVE:
        _REF_KEEPALIVE.append(orc.Ref(any_model_path))


FTYPES = {"q4_0": 2, "q4_1": 3, "q8_0": 7, "q5_0": 8, "q5_1": 9}      # enum ggml_ftype (ggml.h:388-417), the five types the reference tool's README lists


@pytest.mark.parametrize("qname", sorted(FTYPES))
@pytest.mark.parametrize("config,src_ftype", [("tiny", "f16"), ("mini", "f32")])
def test_quantized_file_is_byte_identical_to_the_reference_tool(pkg, orc, weights_file, tmp_path, config, src_ftype, qname):
    if not orc.have_ref():
        pytest.skip("oracle/_ref/libbark_ref.so did not travel with this snapshot")
    src = weights_file(config, src_ftype)
    ours, ref_out = str(tmp_path / "ours.bin"), str(tmp_path / "ref.bin")
    assert pkg.lib().bark_model_quantize(src.encode(), ours.encode(), FTYPES[qname])
    ensure_reference_tables(orc, src)
    R = C.CDLL(orc.REF_SO)
    R.bark_model_quantize.restype = C.c_bool
    R.bark_model_quantize.argtypes = [C.c_char_p, C.c_char_p, C.c_int]
    devnull, saved = os.open(os.devnull, os.O_WRONLY), os.dup(1)
    os.dup2(devnull, 1)                                      # the reference prints one line per tensor
    try:
        assert R.bark_model_quantize(src.encode(), ref_out.encode(), FTYPES[qname])
    finally:
        os.dup2(saved, 1); os.close(devnull); os.close(saved)
    a, b = open(ours, "rb").read(), open(ref_out, "rb").read()
    assert len(a) == len(b) and hashlib.sha1(a).hexdigest() == hashlib.sha1(b).hexdigest()
    assert len(a) < os.path.getsize(src)


def test_quantize_rejects_what_it_cannot_do(pkg, weights_file, tmp_path):
    src = weights_file("tiny", "f16")
    L = pkg.lib()
    assert not L.bark_model_quantize(src.encode(), str(tmp_path / "x.bin").encode(), 12)          # q4_K: k-quants are not implemented here
    assert not L.bark_model_quantize(b"/nonexistent/in.bin", str(tmp_path / "y.bin").encode(), GGML_FTYPE_MOSTLY_Q4_0)
    bad = tmp_path / "bad.bin"; bad.write_bytes(b"\x00" * 64)
    assert not L.bark_model_quantize(str(bad).encode(), str(tmp_path / "z.bin").encode(), GGML_FTYPE_MOSTLY_Q4_0)


@pytest.mark.parametrize("qname", sorted(FTYPES))
@pytest.mark.parametrize("config,src_ftype", [("tiny", "f16"), ("mini", "f32")])
def test_quantised_oracle_matches_the_reference(pkg, orc, weights_file, tmp_path, config, src_ftype, qname):
    """Pins the oracle's quantised paths (q8_0 / q8_1 activation blocks, the 8-lane integer dots, hsum_float_8, get_rows
    dequantisation): teacher-forced logits (merged prompt, decode, ragged coarse prefill), a fine pass and a whole generation,
    oracle vs the unmodified reference on the same quantised file."""
    if not orc.have_ref():
        pytest.skip("oracle/_ref/libbark_ref.so did not travel with this snapshot")
    src = weights_file(config, src_ftype)
    path = str(tmp_path / f"{qname}.bin")
    assert pkg.lib().bark_model_quantize(src.encode(), path.encode(), FTYPES[qname])
    o, r = orc.Oracle(path, seed=0, n_steps=10), orc.Ref(path, seed=0, n_steps=10)
    try:
        assert int(o.hparams(0)[9]) % 1000 == FTYPES[qname]
        rng = np.random.default_rng(17)
        toks, po, pr = o.tokenize("Hello, world"), 0, 0
        for step in range(4):
            lo, po = o.gpt_eval(0, toks, po, True)
            lr, pr = r.gpt_eval(0, toks, pr, True)
            assert np.array_equal(bits(lo), bits(lr)), f"semantic step {step}: {int((lo != lr).sum())} logits differ, max {np.abs(lo - lr).max():.3e}"
            toks = np.array([int(np.argmax(lr[:10000]))], np.int32)
        toks = np.concatenate([rng.integers(0, 10000, 256), [12050], rng.integers(10000, 12048, 21)]).astype(np.int32)
        po = pr = 0
        for step in range(3):
            lo, po = o.gpt_eval(1, toks, po, False)
            lr, pr = r.gpt_eval(1, toks, pr, False)
            assert np.array_equal(bits(lo), bits(lr)), f"coarse step {step}: {int((lo != lr).sum())} logits differ"
            toks = np.array([10000 + int(np.argmax(lr[10000:12048]))], np.int32)
        buf = rng.integers(0, 1024, (8, 1024)).astype(np.int32); buf[:, 300:] = 1024; buf[4:, :] = 1024
        assert np.array_equal(bits(o.fine_eval(buf, 4)), bits(r.fine_eval(buf, 4)))
        go, gr = o.generate("hello world"), r.generate("hello world")
        for k in ("semantic", "coarse", "fine"):
            assert np.array_equal(go[k], gr[k]), k
        assert np.array_equal(bits(go["audio"]), bits(gr["audio"]))
    finally:
        r.close()
