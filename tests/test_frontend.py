"""not gpu: host-only entry points of the front end against the reference's own functions (oracle/_ref: src/pc_loader.cpp compiled as it
is, load_mask / initialize_sensor_lut_simulation sliced out of vofod_nodelet.cpp) where they exist, and known answers everywhere."""
import numpy as np
import pytest

from frontend_cases import write_cloud_files
from vofod_b200 import abi, capi


def _ref():
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref not built here (needs /root/reference)")
    return ref


def test_load_cloud_known_answers(tmp_path):
    files = write_cloud_files(str(tmp_path))
    a = capi.load_cloud(files["mixed.xyz"])
    assert a.shape == (400 + 8, 3)                      # 8 "abc def ghi" lines parse as (0, 0, 0): atof of a non-number
    assert np.count_nonzero((a == 0).all(axis=1)) == 8
    b = capi.load_cloud(files["count.pts"])
    assert b.shape == (400, 3)                          # the count line is not a point
    assert capi.load_cloud(files["empty.txt"]).shape == (0, 3)
    assert capi.load_cloud(files["missing.xyz"]) is None


def test_load_cloud_against_reference_function(tmp_path):
    ref = _ref()
    for name, path in write_cloud_files(str(tmp_path)).items():
        want = ref.load_cloud(path)
        got = capi.load_cloud(path)
        if want is None:
            assert got is None, name
        else:
            assert got is not None and np.array_equal(got, want), name


def test_mask_mangle_against_reference_function():
    ref = _ref()
    rng = np.random.default_rng(1)
    W, H = 64, 16
    img = (rng.random((H, W)) > 0.3).astype(np.uint8) * 255
    shift = rng.integers(0, 40, size=H)
    rn = ref.RefNodelet()
    try:
        for mangle in (1, 0):
            assert np.array_equal(capi.mask_mangle(img, W, H, mangle, shift), rn.load_mask(img, W, H, mangle, shift))
        assert np.array_equal(capi.mask_mangle(img[:, :32], W, H, 1, shift), rn.load_mask(img[:, :32], W, H, 1, shift))  # wrong size -> all ones
        assert np.array_equal(capi.mask_mangle(None, W, H, 1, shift), rn.load_mask(None, W, H, 1, shift))                # missing file -> all ones
    finally:
        rn.close()


def test_mask_mangle_known_answer():
    W, H = 4, 2
    img = np.arange(8, dtype=np.uint8).reshape(H, W)
    out = capi.mask_mangle(img, W, H, 1, [1, 0])
    # row 0 shifted by one column, column-major output: index = ((v + shift[u]) % W) * H + u
    want = np.zeros(8, dtype=np.uint8)
    for u in range(H):
        for v in range(W):
            want[((v + [1, 0][u]) % W) * H + u] = img[u, v]
    assert np.array_equal(out, want)
    assert capi.mask_mangle(None, W, H, 1, [0, 0]).tolist() == [1] * 8
    with pytest.raises(capi.VofodError):
        capi.mask_mangle(img, W, H, 1, [-9, 0])         # the reference's ret.at() throws


def test_sim_xyz_lut_against_reference_function(oracle_mod):
    ref = _ref()
    rn = ref.RefNodelet()
    p = abi.default_params()
    rn.reset(p, 0.5)
    rn.set_sensor(256, 16)
    assert np.array_equal(capi.sim_xyz_lut(256, 16, p.sensor_vfov), rn.sensor_dirs())
    rn.close()
    assert np.array_equal(capi.sim_xyz_lut(256, 16, p.sensor_vfov), oracle_mod.sim_lut(256, 16, p.sensor_vfov))


def test_make_xyz_lut_formula():
    """ouster::make_xyz_lut is third-party code that is not in the reference tree (parity unpinned): checked against an independent numpy
    evaluation of its published formula and against the properties the nodelet relies on (unit directions, offsets of the beam origin)"""
    w, h = 128, 16
    rng = np.random.default_rng(2)
    az = rng.uniform(-3, 3, h)
    alt = np.linspace(45, -45, h)
    T = np.eye(4)
    T[:3, 3] = [0.0, 0.0, 36.18]
    T[0, 0] = T[1, 1] = -1.0
    d, o = capi.make_xyz_lut(w, h, az, alt, 0.001, 27.67, T)
    v, u = np.meshgrid(np.arange(w), np.arange(h))
    enc = 2 * np.pi - v * 2 * np.pi / w
    a = -az[u] * np.pi / 180
    e = alt[u] * np.pi / 180
    dn = np.stack([np.cos(enc + a) * np.cos(e), np.sin(enc + a) * np.cos(e), np.sin(e)], -1)
    on = np.stack([np.cos(enc) - dn[..., 0], np.sin(enc) - dn[..., 1], -dn[..., 2]], -1) * 27.67
    dn = dn @ T[:3, :3].T * 0.001
    on = (on @ T[:3, :3].T + T[:3, 3]) * 0.001
    dn = dn / np.linalg.norm(dn, axis=-1, keepdims=True)
    np.testing.assert_allclose(d.reshape(h, w, 3), dn, atol=2e-7)
    np.testing.assert_allclose(o.reshape(h, w, 3), on, atol=1e-8)
    np.testing.assert_allclose(np.linalg.norm(d, axis=1), 1.0, atol=2e-7)


def test_pack_ouster_points():
    rng = np.random.default_rng(4)
    pts = np.zeros(1000, dtype=capi.OUSTER_POINT_DTYPE)
    for f in ("x", "y", "z", "intensity"):
        pts[f] = rng.normal(size=1000).astype(np.float32)
    pts["range"] = rng.integers(0, 60000, size=1000)
    pts["t"], pts["ring"], pts["ambient"], pts["reflectivity"] = 7, 3, 9, 11   # ignored fields
    out = capi.pack_ouster(pts)
    for f, g in (("x", "x"), ("y", "y"), ("z", "z"), ("intensity", "intensity"), ("range", "range_mm")):
        assert np.array_equal(out[g], pts[f])
