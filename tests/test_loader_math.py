"""Known answers for the loader's transform code (c-ray_b200/host/loader/cr_math.c), the same cases the reference
pins in its own unit tests (reference tests/test_transforms.h:25-245: multiply, determinant via inverse, rotations,
translations, scales, inverse) — exact where the reference asserts equality, 5e-7 where it uses roughly_equals
(src/utils/testrunner.c:20-29)."""
import ctypes as C
import math

import numpy as np
import pytest

import crscene


class Mat4(C.Structure):
    _fields_ = [("m", (C.c_float * 4) * 4)]


class Vec3(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("z", C.c_float)]


class Xform(C.Structure):
    _fields_ = [("type", C.c_int), ("A", Mat4), ("Ainv", Mat4)]


def mat(rows):
    m = Mat4()
    for i in range(4):
        for j in range(4):
            m.m[i][j] = rows[i][j]
    return m


def arr(m):
    return np.array([[m.m[i][j] for j in range(4)] for i in range(4)], dtype=np.float32)


@pytest.fixture(scope="module")
def L():
    lib = crscene.loader()
    lib.crl_mul.restype = Mat4
    lib.crl_mul.argtypes = [C.POINTER(Mat4), C.POINTER(Mat4)]
    lib.crl_inverse.restype = Mat4
    lib.crl_inverse.argtypes = [C.POINTER(Mat4)]
    lib.crl_point.restype = Vec3
    lib.crl_point.argtypes = [Vec3, C.POINTER(Mat4)]
    for name, n in (("crl_xf_rotate_x", 1), ("crl_xf_rotate_y", 1), ("crl_xf_rotate_z", 1), ("crl_xf_translate", 3), ("crl_xf_scale", 3)):
        f = getattr(lib, name)
        f.restype = Xform
        f.argtypes = [C.c_float] * n
    return lib


def test_multiply(L):                                   # test_transforms.h:25-45
    A = mat([[5, 7, 9, 10], [2, 3, 3, 8], [8, 10, 2, 3], [3, 3, 4, 8]])
    B = mat([[3, 10, 12, 18], [12, 1, 4, 9], [9, 10, 12, 2], [3, 12, 4, 10]])
    AB = arr(L.crl_mul(C.byref(A), C.byref(B)))
    assert AB.tolist() == [[210, 267, 236, 271], [93, 149, 104, 149], [171, 146, 172, 268], [105, 169, 128, 169]]


def test_inverse(L):                                    # test_transforms.h:211-225
    N = mat([[0, 0, 0, 1], [0, 0, 1, 0], [0, 2, 0, 0], [2, 0, 0, 0]])
    assert arr(L.crl_inverse(C.byref(N))).tolist() == [[0, 0, 0, 0.5], [0, 0, 0.5, 0], [0, 1, 0, 0], [1, 0, 0, 0]]


def test_inverse_is_adjugate_over_determinant(L):       # determinant -4 case of test_transforms.h:71-89, through the inverse
    M = mat([[1, 2, 0, 0], [1, 1, 3, 0], [0, 2, -2, 0], [0, 0, 3, 1]])
    inv = arr(L.crl_inverse(C.byref(M))).astype(np.float64)
    expect = np.linalg.inv(arr(M).astype(np.float64))
    assert np.allclose(inv, expect, atol=1e-6)
    assert np.allclose(inv @ arr(M), np.eye(4), atol=1e-6)


@pytest.mark.parametrize("fn,vec,expect", [
    ("crl_xf_rotate_x", (0, 1, 0), (0, 0, 1)),           # test_transforms.h:93-103
    ("crl_xf_rotate_y", (1, 0, 0), (0, 0, -1)),          # :105-115
    ("crl_xf_rotate_z", (0, 1, 0), (-1, 0, 0)),          # :117-127
])
def test_rotations(L, fn, vec, expect):
    rads = np.float32(np.float32(90.0) * np.float32(math.pi)) / np.float32(180.0)
    t = getattr(L, fn)(rads)
    p = L.crl_point(Vec3(*vec), C.byref(t.A))
    got = (p.x, p.y, p.z)
    assert abs(math.sqrt(sum(c * c for c in got)) - 1.0) <= 5e-7
    assert all(abs(g - e) <= 5e-7 for g, e in zip(got, expect))
    assert np.allclose(arr(t.A).astype(np.float64) @ arr(t.Ainv), np.eye(4), atol=1e-6)


@pytest.mark.parametrize("fn,args,vec,expect", [
    ("crl_xf_translate", (1, 0, 0), (-10, 0, 0), (-9, 0, 0)),        # test_transforms.h:131-169
    ("crl_xf_translate", (0, 1, 0), (0, -10, 0), (0, -9, 0)),
    ("crl_xf_translate", (0, 0, 1), (0, 0, -10), (0, 0, -9)),
    ("crl_xf_translate", (-1, -10, -100), (0, 0, 0), (-1, -10, -100)),
    ("crl_xf_scale", (3, 1, 1), (-10, 0, 0), (-30, 0, 0)),           # :173-208
    ("crl_xf_scale", (1, 3, 1), (0, -10, 0), (0, -30, 0)),
    ("crl_xf_scale", (1, 1, 3), (0, 0, -10), (0, 0, -30)),
    ("crl_xf_scale", (3, 3, 3), (1, 2, 3), (3, 6, 9)),
])
def test_translate_and_scale_exact(L, fn, args, vec, expect):
    t = getattr(L, fn)(*args)
    p = L.crl_point(Vec3(*vec), C.byref(t.A))
    assert (p.x, p.y, p.z) == expect
