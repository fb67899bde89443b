"""Shared helpers of the parity tests: configurations of BASELINE.json and side-by-side drivers."""
import numpy as np

from vofod_b200 import abi, synth


def params_for(size_xyz, offset_xyz=(0.0, 0.0, -1.25)):
    p = abi.default_params()
    for i in range(3):
        p.oparea_offset[i] = float(offset_xyz[i])
        p.oparea_size[i] = float(size_xyz[i])
    return p


def cfg2_params():
    """BASELINE.json configs[1]: 0.5 m voxels, 200 x 200 x 80 m map -> 401 x 401 x 161 cells."""
    return params_for((200.0, 200.0, 80.0)), 0.5


def small_params():
    """A reduced map (same voxel size, same code paths) for quick tests: 80 x 80 x 30 m."""
    return params_for((80.0, 80.0, 30.0)), 0.5


class Sensor:
    def __init__(self, W, H, vfov=np.pi / 2):
        self.W, self.H = W, H
        self.dirs = synth.sim_lut(W, H, vfov)

    def scan(self, scene, k, map_scale=1.0):
        return synth.generate(scene, k, self.W, self.H, self.dirs, map_scale)


def setup_pair(cpu, gpu, params, vs, sensor, fixed=True):
    """Reset both sides to the same state.  With fixed=True the oracle applies the raycast from the same exact
    fixed-point path-length sums as the GPU (bit-exact score parity); with fixed=False it keeps the reference's
    sequential fp32 accumulation (scores then agree to 1e-5 relative)."""
    gpu.reset(params, vs)
    gpu.set_sensor(sensor.W, sensor.H, sensor.dirs)
    cpu.reset(params, vs)
    cpu.set_sensor(sensor.W, sensor.H, sensor.dirs)
    return gpu, cpu


def rel_err(a, b, floor=1.0):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.abs(a - b) / np.maximum(np.abs(b), floor)


def assert_vox_equal(a, b):
    assert len(a) == len(b), (len(a), len(b))
    for f in ("x", "y", "z", "count"):
        assert np.array_equal(a[f], b[f]), f


def struct_close(a, b, fields, rtol=1e-5, atol=1e-6):
    for f in fields:
        np.testing.assert_allclose(np.asarray(a[f], dtype=np.float64), np.asarray(b[f], dtype=np.float64), rtol=rtol, atol=atol, err_msg=f)
