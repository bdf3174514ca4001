"""The known-answer vectors the reference's OWN tests hold for ops on this path: ggml/tests/test-conv1d.cpp:233-281
(ggml_conv_1d, K=3, 10 -> 10 channels, zero padding 1) and ggml/tests/test-conv-transpose-1d.cpp:415-560 (ggml_conv_transpose_1d,
cases 0-6: strides 1/2/3, 1-3 channels), fed through the conv cores of the C oracle that the EnCodec decoder restatement uses.
The expected arrays below are copied values (test data, not code); case 7 of the transposed-conv test (32 x 1584 outputs) sums
~6.7e7-sized integers in f32 and therefore pins an accumulation order the f16 path does not share, so it is left out."""
import ctypes as C

import numpy as np
import pytest


def lib(orc):
    orc.build_oracle()
    L = C.CDLL(orc.ORACLE_SO)
    L.orc_test_conv1d.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    L.orc_test_convtr1d.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    return L


def p(a):
    return a.ctypes.data_as(C.c_void_p)


def test_conv1d_known_answer(orc):
    K, IC, OC, IL = 3, 10, 10, 8
    w = np.full(K * IC * OC, 4.5, np.float32)                 # ggml kernel [K, IC, OC]: memory [oc][ic][k]
    x = np.full(IC * IL, 2.5, np.float32)                     # [IL, IC]: memory [ic][t]
    y = np.zeros((OC, IL), np.float32)
    lib(orc).orc_test_conv1d(p(w), K, IC, OC, p(x), IL, 1, p(y))
    row = np.array([225.0] + [337.5] * 6 + [225.0], np.float32)
    assert np.array_equal(y, np.tile(row, (OC, 1)))


DATA = np.arange(16 * 32 * 32, dtype=np.float32) % 1024
CASES = [  # kernel data [Cin][Cout][k], k, Cout, Cin, input [Cin][T], T, stride, expected [Cout][(T-1)*stride+k]
    ([1, 2, 3], 3, 1, 1, [1, 2], 2, 1, [1, 4, 7, 6]),
    ([1, 2, 3, 3, 2, 1], 3, 1, 2, [2, 3, 1, 1, 3, 2], 3, 1, [5, 18, 26, 18, 5]),
    ([3, 2, 1, 1, 2, 3, 1, 2, 3, 3, 2, 1], 3, 2, 2, [2, 3, 1, 1, 3, 2], 3, 1, [7, 18, 22, 18, 7, 5, 18, 26, 18, 5]),
    ([3, 2, 1, 1, 2, 3, 1, 2, 3, 3, 2, 1], 3, 2, 2, [2, 3, 1, 1, 3, 2], 3, 2, [7, 6, 17, 12, 17, 6, 7, 5, 6, 19, 12, 19, 6, 5]),
    (DATA[:12], 2, 3, 2, DATA[:6], 3, 1, [18, 45, 59, 37, 24, 61, 83, 51, 30, 77, 107, 65]),
    (DATA[:12], 2, 3, 2, DATA[:6], 3, 2, [18, 21, 24, 29, 30, 37, 24, 27, 34, 39, 44, 51, 30, 33, 44, 49, 58, 65]),
    (DATA[:12], 2, 3, 2, DATA[:6], 3, 3, [18, 21, 0, 24, 29, 0, 30, 37, 24, 27, 0, 34, 39, 0, 44, 51, 30, 33, 0, 44, 49, 0, 58, 65]),
]


@pytest.mark.parametrize("case", range(len(CASES)))
def test_conv_transpose_1d_known_answers(orc, case):
    w, k, Cout, Cin, x, T, stride, want = CASES[case]
    w, x = np.ascontiguousarray(w, np.float32), np.ascontiguousarray(x, np.float32)
    y = np.zeros(Cout * ((T - 1) * stride + k), np.float32)
    lib(orc).orc_test_convtr1d(p(w), k, Cout, Cin, p(x), T, stride, p(y))
    assert np.array_equal(y, np.asarray(want, np.float32))
