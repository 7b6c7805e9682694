"""ctypes binding of include/mcl3dl_b200.h (the C ABI of the CUDA engine).

`Engine` is the C-ABI handle (set_map / measure / measure_device / beam_status).  The reference is
compiled C++, so the host-side mirror of its plugin classes is C++ too:
mcl_3dl_b200/host/lidar_measurement_model_b200.h.  This module is what tests and bench.py call.

There is no fallback: if the CUDA library is missing or no device is usable, construction raises.
"""
import ctypes as C
import os

import numpy as np

from . import build as _build
from .synth import POINT, POSE, RESULT, STATE

ERR = {0: "ok", -1: "invalid argument", -2: "measure() before set_map()", -3: "CUDA runtime error",
       -4: "no usable CUDA device", -5: "grid exceeds 2^31-1 cells", -6: "search radius out of range"}


class EngineError(RuntimeError):
    def __init__(self, code, detail=""):
        self.code = code
        super().__init__("mcl3dl error %d (%s) %s" % (code, ERR.get(code, "?"), detail))


class LikParams(C.Structure):
    """mcl3dl_lik_params; defaults = include/mcl_3dl/parameters.h:74-76."""
    _fields_ = [("match_weight", C.c_float), ("match_dist_min", C.c_float), ("match_dist_flat", C.c_float),
                ("dist_weight", C.c_float * 3)]

    def __init__(self, match_weight=5.0, match_dist_min=0.2, match_dist_flat=0.05, dist_weight=(1.0, 1.0, 1.0)):
        super().__init__()
        self.match_weight, self.match_dist_min, self.match_dist_flat = match_weight, match_dist_min, match_dist_flat
        self.dist_weight[:] = dist_weight


class BeamParams(C.Structure):
    """mcl3dl_beam_params (derived; fill through beam_params_from_reference)."""
    _fields_ = [("map_grid_size", C.c_double * 3), ("dda_grid_size", C.c_double), ("ray_angle_half", C.c_double),
                ("hit_tolerance", C.c_double), ("hit_range_sq", C.c_float), ("sin_total_ref", C.c_float),
                ("beam_likelihood", C.c_float), ("beam_likelihood_min", C.c_float),
                ("filter_label_max", C.c_uint32), ("add_penalty_short_only_mode", C.c_int32),
                ("use_raycast_using_dda", C.c_int32), ("_reserved", C.c_int32)]

    def as_tuple(self):
        return (tuple(self.map_grid_size), self.dda_grid_size, self.ray_angle_half, self.hit_tolerance,
                self.hit_range_sq, self.sin_total_ref, self.beam_likelihood, self.beam_likelihood_min,
                self.filter_label_max, self.add_penalty_short_only_mode, self.use_raycast_using_dda)


class MapInfo(C.Structure):
    _fields_ = [("n_points", C.c_uint64), ("nn_dims", C.c_int32 * 3), ("nn_cell", C.c_float),
                ("nn_origin", C.c_float * 3), ("dda_dims", C.c_int32 * 3), ("dda_min", C.c_float * 3),
                ("dda_max", C.c_float * 3), ("device_bytes", C.c_uint64), ("build_ms", C.c_double)]


class UpdateSummary(C.Structure):
    _fields_ = [("weight_sum", C.c_float), ("entropy", C.c_float), ("match_ratio_min", C.c_float),
                ("match_ratio_max", C.c_float), ("kept", C.c_int32), ("max_index", C.c_uint32)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class ScanParams(C.Structure):
    """mcl3dl_scan_params; defaults = parameters.h:67-77,95-103 and the node's downsample_x/y/z (parameters.cpp:85-87)."""
    _fields_ = [("downsample", C.c_float * 3),
                ("lik_clip_near", C.c_float), ("lik_clip_far", C.c_float), ("lik_clip_z_min", C.c_float), ("lik_clip_z_max", C.c_float),
                ("beam_clip_near", C.c_float), ("beam_clip_far", C.c_float), ("beam_clip_z_min", C.c_float), ("beam_clip_z_max", C.c_float),
                ("lik_num_points", C.c_uint32), ("beam_num_points", C.c_uint32), ("seed", C.c_uint64)]

    def __init__(self, downsample=(0.1, 0.1, 0.1), lik_clip=(0.5, 10.0, -2.0, 2.0), beam_clip=(0.5, 4.0, -2.0, 2.0),
                 lik_num_points=96, beam_num_points=3, seed=1):
        super().__init__()
        self.downsample[:] = downsample
        self.lik_clip_near, self.lik_clip_far, self.lik_clip_z_min, self.lik_clip_z_max = lik_clip
        self.beam_clip_near, self.beam_clip_far, self.beam_clip_z_min, self.beam_clip_z_max = beam_clip
        self.lik_num_points, self.beam_num_points, self.seed = lik_num_points, beam_num_points, seed


class ScanInfo(C.Structure):
    _fields_ = [(k, C.c_uint32) for k in ("n_raw", "n_downsampled", "n_lik_clipped", "n_beam_clipped", "n_lik", "n_beam")]

    def as_dict(self):
        return {k: int(getattr(self, k)) for k, _ in self._fields_}


class Estimate(C.Structure):
    """mcl3dl_estimate"""
    _fields_ = [("mean_biased", C.c_float * 8), ("max_state", C.c_float * 8), ("max_index", C.c_uint32),
                ("weight_sum_biased", C.c_float), ("cov", C.c_float * 36)]


class WorkStats(C.Structure):
    _fields_ = [("lik_index_rows", C.c_uint64), ("lik_points_scanned", C.c_uint64),
                ("beam_cells_stepped", C.c_uint64), ("beam_cells_occupied", C.c_uint64),
                ("beam_points_tested", C.c_uint64)]

    def as_dict(self):
        return {k: int(getattr(self, k)) for k, _ in self._fields_}


EXPORTED_SYMBOLS = ["mcl3dl_measure_update", "mcl3dl_collect_stats", "mcl3dl_read_stats", "mcl3dl_abi_version", "mcl3dl_create", "mcl3dl_destroy", "mcl3dl_set_map", "mcl3dl_set_params",
                    "mcl3dl_measure", "mcl3dl_measure_device", "mcl3dl_beam_status",
                    "mcl3dl_beam_params_from_reference", "mcl3dl_get_map_info", "mcl3dl_last_timing",
                    "mcl3dl_kernel_launches", "mcl3dl_strerror", "mcl3dl_last_error_detail", "mcl3dl_near_field_info", "mcl3dl_nn_field_info", "mcl3dl_host_alloc", "mcl3dl_host_free", "mcl3dl_field_mode", "mcl3dl_field_nodes", "mcl3dl_field_upload", "mcl3dl_collect_timing",
                    "mcl3dl_exchange_create", "mcl3dl_exchange_open", "mcl3dl_measure_exchange_device", "mcl3dl_exchange_current",
                    "mcl3dl_particles_set", "mcl3dl_particles_get", "mcl3dl_particles_predict",
                    "mcl3dl_particles_measure_update", "mcl3dl_particles_resample", "mcl3dl_particles_estimate", "mcl3dl_scan_prepare", "mcl3dl_scan_get",
                    "mcl3dl_particles_measure_update_prepared"]

_LIBS = {}


def load_library(path=None):
    """Load (building if stale and nvcc is present) the in-tree CUDA library.  Raises if unavailable.
    `path` (or $MCL3DL_LIB) names an experiment variant built by build.build(out=...); libraries are cached per path."""
    path = path or os.environ.get("MCL3DL_LIB") or _build.LIB
    if path in _LIBS:
        return _LIBS[path]
    if path == _build.LIB and _build.stale():
        try:
            _build.build()
        except Exception:
            if not os.path.exists(path):
                raise
    L = C.CDLL(path)
    vp, sz = C.c_void_p, C.c_size_t
    L.mcl3dl_abi_version.restype = C.c_int
    L.mcl3dl_create.argtypes = [C.POINTER(vp), C.POINTER(C.c_int), C.c_int]
    L.mcl3dl_destroy.argtypes = [vp]
    L.mcl3dl_destroy.restype = None
    L.mcl3dl_set_map.argtypes = [vp, vp, sz, C.c_uint64, vp, vp]
    L.mcl3dl_set_params.argtypes = [vp, vp, vp]
    L.mcl3dl_measure.argtypes = [vp, vp, sz, vp, sz, vp, sz, vp, sz, vp]
    L.mcl3dl_measure_update.argtypes = [vp, vp, sz, vp, sz, vp, sz, vp, sz, vp, vp, vp, vp, vp]
    L.mcl3dl_measure_device.argtypes = [vp, vp, sz, vp, sz, vp, sz, vp, sz, vp, vp]
    L.mcl3dl_beam_status.argtypes = [vp, vp, sz, vp, sz, vp, sz, vp]
    L.mcl3dl_beam_params_from_reference.argtypes = [vp, C.c_float, C.c_float, C.c_float, sz, C.c_float, C.c_float,
                                                    C.c_uint32, C.c_float, C.c_int, C.c_int, C.c_float, C.c_float]
    L.mcl3dl_beam_params_from_reference.restype = None
    L.mcl3dl_get_map_info.argtypes = [vp, vp]
    L.mcl3dl_collect_stats.argtypes = [vp, C.c_int]
    L.mcl3dl_read_stats.argtypes = [vp, vp]
    L.mcl3dl_last_timing.argtypes = [vp] + [C.POINTER(C.c_double)] * 4
    L.mcl3dl_kernel_launches.argtypes = [vp]
    L.mcl3dl_kernel_launches.restype = C.c_uint64
    L.mcl3dl_strerror.argtypes = [C.c_int]
    L.mcl3dl_strerror.restype = C.c_char_p
    L.mcl3dl_last_error_detail.argtypes = [vp]
    L.mcl3dl_last_error_detail.restype = C.c_char_p
    L.mcl3dl_collect_timing.argtypes = [vp, C.c_int]
    L.mcl3dl_particles_set.argtypes = [vp, vp, vp, sz]
    L.mcl3dl_particles_get.argtypes = [vp, vp, vp, sz]
    L.mcl3dl_particles_predict.argtypes = [vp, vp, vp, C.c_float, C.c_float, C.c_float]
    L.mcl3dl_particles_measure_update.argtypes = [vp, vp, sz, vp, sz, vp, sz, C.c_float, vp]
    L.mcl3dl_particles_resample.argtypes = [vp, vp, vp, C.c_float, C.c_uint64]
    L.mcl3dl_particles_estimate.argtypes = [vp, vp, C.c_float, C.c_float, vp]
    L.mcl3dl_scan_prepare.argtypes = [vp, vp, sz, vp, vp]
    L.mcl3dl_scan_get.argtypes = [vp, C.c_int, vp, sz, C.POINTER(sz)]
    L.mcl3dl_particles_measure_update_prepared.argtypes = [vp, vp, sz, C.c_float, vp]
    L.mcl3dl_exchange_create.argtypes = [vp, sz, C.c_int, C.c_int, vp]
    L.mcl3dl_exchange_open.argtypes = [vp, vp]
    L.mcl3dl_measure_exchange_device.argtypes = [vp, vp, sz, vp, sz, vp, sz, vp, sz, vp, C.POINTER(vp)]
    L.mcl3dl_exchange_current.argtypes = [vp, vp, C.POINTER(vp), C.POINTER(C.c_int)]
    L.mcl3dl_near_field_info.argtypes = [vp, C.POINTER(C.c_int32), C.POINTER(C.c_uint64)]
    L.mcl3dl_nn_field_info.argtypes = [vp, C.POINTER(C.c_uint64)]
    L.mcl3dl_field_mode.argtypes = [vp, C.c_int]
    L.mcl3dl_field_nodes.argtypes = [vp, vp, C.POINTER(C.c_int32), C.POINTER(C.c_float), C.POINTER(C.c_float)]
    L.mcl3dl_field_upload.argtypes = [vp, vp, C.POINTER(C.c_int32)]
    L.mcl3dl_host_alloc.argtypes = [vp, sz, C.POINTER(vp)]
    L.mcl3dl_host_free.argtypes = [vp, vp]
    assert L.mcl3dl_abi_version() == 3
    _LIBS[path] = L
    return L


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None and a.size else None


def beam_params_from_reference(map_grid=(0.1, 0.1, 0.1), num_points_default=3, beam_likelihood_min=0.2,
                               ang_total_ref=np.pi / 6.0, filter_label_max=0xFFFFFFFF, hit_range=0.3,
                               add_penalty_short_only_mode=True, ray_angle_half=0.25 * np.pi / 180.0,
                               dda_grid_size=0.2, use_raycast_using_dda=True):
    """LidarMeasurementModelBeamParameters (parameters.h:95-111) -> derived mcl3dl_beam_params.
    (use_raycast_using_dda defaults to True here; the reference's default is False = the KD-tree caster.)"""
    L = load_library()
    bp = BeamParams()
    L.mcl3dl_beam_params_from_reference(C.byref(bp), map_grid[0], map_grid[1], map_grid[2], num_points_default,
                                        beam_likelihood_min, ang_total_ref, filter_label_max, hit_range,
                                        1 if add_penalty_short_only_mode else 0, 1 if use_raycast_using_dda else 0,
                                        ray_angle_half, dda_grid_size)
    return bp


class Engine:
    """Handle on mcl3dl_engine.  devices: list of CUDA ordinals (map replicated, particles split)."""

    def __init__(self, devices=(0,), lib_path=None):
        self.L = load_library(lib_path)
        self.h = C.c_void_p()
        ids = (C.c_int * len(devices))(*devices)
        rc = self.L.mcl3dl_create(C.byref(self.h), ids, len(devices))
        if rc != 0:
            self.h = None
            raise EngineError(rc)
        self.devices = tuple(devices)

    def _check(self, rc):
        if rc != 0:
            raise EngineError(rc, self.L.mcl3dl_last_error_detail(self.h).decode())

    def close(self):
        if getattr(self, "h", None):
            self.L.mcl3dl_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_map(self, map_pts, lik=None, beam=None, stamp=None):
        """stamp: the cloud's header.stamp, the reference's rebuild trigger (same stamp, size and params = no-op).
        None: a fresh stamp per call (as the C++ adapter does), so that a different cloud is never mistaken for the old."""
        map_pts = np.ascontiguousarray(map_pts, dtype=POINT)
        if stamp is None:
            self._stamp = getattr(self, "_stamp", 0) + 1
            stamp = (1 << 40) + self._stamp
        self._check(self.L.mcl3dl_set_map(self.h, _ptr(map_pts), len(map_pts), stamp,
                                          C.byref(lik) if lik is not None else None,
                                          C.byref(beam) if beam is not None else None))

    def set_params(self, lik=None, beam=None):
        self._check(self.L.mcl3dl_set_params(self.h, C.byref(lik) if lik is not None else None,
                                             C.byref(beam) if beam is not None else None))

    def map_info(self):
        mi = MapInfo()
        self._check(self.L.mcl3dl_get_map_info(self.h, C.byref(mi)))
        return mi

    def near_field_info(self):
        """[(k, bytes)] of the near-field screens of the staged map: [0] likelihood search, [1] KD-tree raycaster."""
        k = (C.c_int32 * 2)()
        b = (C.c_uint64 * 2)()
        self._check(self.L.mcl3dl_near_field_info(self.h, k, b))
        return [(int(k[0]), int(b[0])), (int(k[1]), int(b[1]))]

    def nn_field_info(self):
        """The NN field of the staged map: {"bytes", "candidates", "overflow_cells", "voxel_edge", "wide_cells"} (bytes 0:
        not staged)."""
        v = (C.c_uint64 * 5)()
        self._check(self.L.mcl3dl_nn_field_info(self.h, v))
        return {"bytes": int(v[0]), "candidates": int(v[1]), "overflow_cells": int(v[2]), "voxel_edge": v[3] * 1e-6,
                "wide_cells": int(v[4])}

    def field_mode(self, enable=True):
        """Opt-in, inexact: route the likelihood model through the trilinear distance volume (include/mcl3dl_b200.h)."""
        self._check(self.L.mcl3dl_field_mode(self.h, 1 if enable else 0))

    def field_nodes(self, download=True):
        """(nodes float32[nz, ny, nx] or None, origin[3], edge) of the field-mode lattice (rescaled space)."""
        dims, org, edge = (C.c_int32 * 3)(), (C.c_float * 3)(), C.c_float(0)
        self._check(self.L.mcl3dl_field_nodes(self.h, None, dims, org, C.byref(edge)))
        nodes = None
        if download:
            nodes = np.zeros((dims[2], dims[1], dims[0]), dtype=np.float32)
            self._check(self.L.mcl3dl_field_nodes(self.h, _ptr(nodes), dims, org, C.byref(edge)))
        return nodes, np.array(list(org), np.float32), float(edge.value), tuple(int(d) for d in dims)

    def field_upload(self, nodes):
        nodes = np.ascontiguousarray(nodes, dtype=np.float32)
        dims = (C.c_int32 * 3)(nodes.shape[2], nodes.shape[1], nodes.shape[0])
        self._check(self.L.mcl3dl_field_upload(self.h, _ptr(nodes), dims))

    def measure(self, poses, lik_pts=None, beam_pts=None, origins=None, out=None):
        poses = np.ascontiguousarray(poses, dtype=POSE)
        lik_pts = np.ascontiguousarray(lik_pts if lik_pts is not None else np.zeros(0, POINT), dtype=POINT)
        beam_pts = np.ascontiguousarray(beam_pts if beam_pts is not None else np.zeros(0, POINT), dtype=POINT)
        origins = np.ascontiguousarray(origins if origins is not None else np.zeros((0, 3)), dtype=np.float32)
        origins = origins.reshape(-1, 3)
        if out is None:
            out = np.zeros(len(poses), dtype=RESULT)
        self._check(self.L.mcl3dl_measure(self.h, _ptr(poses), len(poses), _ptr(lik_pts), len(lik_pts),
                                          _ptr(beam_pts), len(beam_pts), _ptr(origins), len(origins), _ptr(out)))
        return out

    def host_array(self, n, dtype):
        """A zeroed numpy array of n items in page-locked memory of this engine (mcl3dl_host_alloc): pose and record
        arrays passed to measure / bind_measure from such memory are transferred in place, without the staging memcpy.
        The memory is returned by host_free(array) or when the engine is closed - do not use the array after that."""
        dtype = np.dtype(dtype)
        nbytes = max(1, int(n) * dtype.itemsize)
        p = C.c_void_p()
        self._check(self.L.mcl3dl_host_alloc(self.h, nbytes, C.byref(p)))
        buf = (C.c_char * nbytes).from_address(p.value)
        arr = np.frombuffer(buf, dtype=dtype, count=int(n))
        arr[...] = np.zeros((), dtype)
        return arr

    def host_free(self, arr):
        self._check(self.L.mcl3dl_host_free(self.h, arr.ctypes.data))

    def bind_measure(self, poses, lik_pts, beam_pts, origins, out):
        """A repeated update over the SAME host arrays (their contents may change between calls): resolves the buffer
        addresses once and returns a zero-argument callable that runs mcl3dl_measure on them.  What a C++ caller does
        anyway; it keeps ~20 us of numpy/ctypes argument marshalling out of every call.  The arrays must already have
        the ABI dtypes and be C-contiguous (no hidden copies, or the call would read stale copies)."""
        def chk(a, dt, name):
            if a is None:
                return np.zeros(0, dt)
            if a.dtype != dt or not a.flags["C_CONTIGUOUS"]:
                raise ValueError("%s must be a C-contiguous array of %s" % (name, dt))
            return a
        poses, lik_pts, beam_pts = chk(poses, POSE, "poses"), chk(lik_pts, POINT, "lik_pts"), chk(beam_pts, POINT, "beam_pts")
        origins = chk(origins if origins is None else origins.reshape(-1, 3), np.dtype(np.float32), "origins")
        out = chk(out, RESULT, "out")
        if len(out) != len(poses):
            raise ValueError("out must hold one record per particle")
        keep = (poses, lik_pts, beam_pts, origins, out)  # the closure keeps the buffers alive
        args = (self.h, _ptr(poses), len(poses), _ptr(lik_pts), len(lik_pts), _ptr(beam_pts), len(beam_pts), _ptr(origins),
                len(origins.reshape(-1, 3)), _ptr(out))
        fn = self.L.mcl3dl_measure

        def call():
            rc = fn(*args)
            if rc != 0:
                self._check(rc)
            return keep[4]
        return call

    def measure_update(self, poses, lik_pts, beam_pts, origins, prior, extra_likelihood=None, want_records=False):
        """Fused measurement + weight update (pf::ParticleFilter::measure with the node's lambda).
        Returns (posterior float32[P], summary dict, records or None)."""
        poses = np.ascontiguousarray(poses, dtype=POSE)
        lik_pts = np.ascontiguousarray(lik_pts if lik_pts is not None else np.zeros(0, POINT), dtype=POINT)
        beam_pts = np.ascontiguousarray(beam_pts if beam_pts is not None else np.zeros(0, POINT), dtype=POINT)
        origins = np.ascontiguousarray(origins if origins is not None else np.zeros((0, 3)), dtype=np.float32)
        origins = origins.reshape(-1, 3)
        prior = np.ascontiguousarray(prior, dtype=np.float32)
        extra = np.ascontiguousarray(extra_likelihood, dtype=np.float32) if extra_likelihood is not None else None
        post = np.zeros(len(poses), dtype=np.float32)
        rec = np.zeros(len(poses), dtype=RESULT) if want_records else None
        summ = UpdateSummary()
        self._check(self.L.mcl3dl_measure_update(self.h, _ptr(poses), len(poses), _ptr(lik_pts), len(lik_pts),
                                                 _ptr(beam_pts), len(beam_pts), _ptr(origins), len(origins),
                                                 _ptr(prior), _ptr(extra) if extra is not None else None,
                                                 _ptr(post), _ptr(rec) if rec is not None else None, C.byref(summ)))
        return post, summ.as_dict(), rec

    def measure_device(self, d_poses, n_particles, d_lik, n_lik, d_beam, n_beam, d_origins, n_origins, d_out,
                       stream=0):
        """All d_* are raw device addresses (ints, e.g. torch.Tensor.data_ptr()); async on `stream`."""
        self._check(self.L.mcl3dl_measure_device(self.h, d_poses, n_particles, d_lik, n_lik, d_beam, n_beam,
                                                 d_origins, n_origins, d_out, stream))

    def beam_status(self, poses, beam_pts, origins):
        poses = np.ascontiguousarray(poses, dtype=POSE)
        beam_pts = np.ascontiguousarray(beam_pts, dtype=POINT)
        origins = np.ascontiguousarray(origins, dtype=np.float32).reshape(-1, 3)
        st = np.zeros((len(poses), len(beam_pts)), dtype=np.uint8)
        self._check(self.L.mcl3dl_beam_status(self.h, _ptr(poses), len(poses), _ptr(beam_pts), len(beam_pts),
                                              _ptr(origins), len(origins), _ptr(st)))
        return st

    def collect_stats(self, enable=True):
        self._check(self.L.mcl3dl_collect_stats(self.h, 1 if enable else 0))

    def read_stats(self):
        ws = WorkStats()
        self._check(self.L.mcl3dl_read_stats(self.h, C.byref(ws)))
        return ws.as_dict()

    # ---- resident particle set (scope row f3; prepared, see include/mcl3dl_b200.h)
    def particles_set(self, states, prob):
        states = np.ascontiguousarray(states, dtype=STATE)
        prob = np.ascontiguousarray(prob, dtype=np.float32)
        assert len(states) == len(prob)
        self._check(self.L.mcl3dl_particles_set(self.h, _ptr(states), _ptr(prob), len(states)))
        self._n_resident = len(states)

    def particles_get(self):
        states = np.zeros(self._n_resident, dtype=STATE)
        prob = np.zeros(self._n_resident, dtype=np.float32)
        self._check(self.L.mcl3dl_particles_get(self.h, _ptr(states), _ptr(prob), len(states)))
        return states, prob

    def particles_predict(self, odom_prev, odom_current, time_diff, tc_lin, tc_ang):
        a = np.ascontiguousarray(odom_prev, dtype=POSE).reshape(1)
        b = np.ascontiguousarray(odom_current, dtype=POSE).reshape(1)
        self._check(self.L.mcl3dl_particles_predict(self.h, _ptr(a), _ptr(b), time_diff, tc_lin, tc_ang))

    def particles_measure_update(self, lik_pts, beam_pts, origins, odom_err_integ_lin_sigma=0.0):
        lik_pts = np.ascontiguousarray(lik_pts if lik_pts is not None else np.zeros(0, POINT), dtype=POINT)
        beam_pts = np.ascontiguousarray(beam_pts if beam_pts is not None else np.zeros(0, POINT), dtype=POINT)
        origins = np.ascontiguousarray(origins if origins is not None else np.zeros((0, 3)), dtype=np.float32).reshape(-1, 3)
        summ = UpdateSummary()
        self._check(self.L.mcl3dl_particles_measure_update(self.h, _ptr(lik_pts), len(lik_pts), _ptr(beam_pts), len(beam_pts),
                                                           _ptr(origins), len(origins), odom_err_integ_lin_sigma,
                                                           C.byref(summ)))
        return summ.as_dict()

    def particles_resample(self, sigma_pos, sigma_rpy, initial_frac, seed):
        sp = np.ascontiguousarray(sigma_pos, dtype=np.float32)
        sr = np.ascontiguousarray(sigma_rpy, dtype=np.float32)
        self._check(self.L.mcl3dl_particles_resample(self.h, _ptr(sp), _ptr(sr), initial_frac, seed))

    # ---- scan preprocessing on the device (scope row f4; include/mcl3dl_b200.h)
    def scan_prepare(self, raw_pts, params):
        raw_pts = np.ascontiguousarray(raw_pts, dtype=POINT)
        info = ScanInfo()
        self._check(self.L.mcl3dl_scan_prepare(self.h, _ptr(raw_pts), len(raw_pts), C.byref(params), C.byref(info)))
        return info.as_dict()

    def scan_get(self, which):
        """which: 0 downsampled, 1 / 2 clipped (likelihood / beam), 3 / 4 sampled scans."""
        n = C.c_size_t(0)
        self._check(self.L.mcl3dl_scan_get(self.h, which, None, 0, C.byref(n)))
        out = np.zeros(n.value, dtype=POINT)
        if n.value:
            self._check(self.L.mcl3dl_scan_get(self.h, which, _ptr(out), len(out), C.byref(n)))
        return out

    def particles_measure_update_prepared(self, origins, odom_err_integ_lin_sigma=0.0):
        origins = np.ascontiguousarray(origins if origins is not None else np.zeros((0, 3)), dtype=np.float32).reshape(-1, 3)
        summ = UpdateSummary()
        self._check(self.L.mcl3dl_particles_measure_update_prepared(self.h, _ptr(origins), len(origins), odom_err_integ_lin_sigma,
                                                                    C.byref(summ)))
        return summ.as_dict()

    def particles_estimate(self, state_prev=None, bias_var_dist=1.0, bias_var_ang=1.0):
        """The node's pose estimate on the resident set: {"mean_biased": POSE[1], "max_state": POSE[1], "max_index",
        "weight_sum_biased", "cov": float32[6, 6]} (include/mcl3dl_b200.h: mcl3dl_particles_estimate)."""
        prev = np.ascontiguousarray(state_prev, dtype=POSE).reshape(1) if state_prev is not None else None
        e = Estimate()
        self._check(self.L.mcl3dl_particles_estimate(self.h, _ptr(prev) if prev is not None else None, bias_var_dist,
                                                     bias_var_ang, C.byref(e)))
        return {"mean_biased": np.frombuffer(bytes(e.mean_biased), dtype=POSE).copy(),
                "max_state": np.frombuffer(bytes(e.max_state), dtype=POSE).copy(), "max_index": int(e.max_index),
                "weight_sum_biased": float(e.weight_sum_biased), "cov": np.array(list(e.cov), np.float32).reshape(6, 6)}

    # ---- record exchange over peer memory (one process per GPU; see include/mcl3dl_b200.h)
    IPC_HANDLE_BYTES = 64

    def exchange_create(self, n_local, world, rank):
        """Allocates this rank's exchange buffer; returns its CUDA IPC handle (bytes) for the ranks to all-gather."""
        h = (C.c_ubyte * self.IPC_HANDLE_BYTES)()
        self._check(self.L.mcl3dl_exchange_create(self.h, n_local, world, rank, h))
        return bytes(h)

    def exchange_open(self, handles):
        """handles: the world ranks' ipc handles concatenated in rank order (bytes)."""
        buf = (C.c_ubyte * len(handles)).from_buffer_copy(handles)
        self._check(self.L.mcl3dl_exchange_open(self.h, buf))

    def measure_exchange_device(self, d_poses, n_local, d_lik, n_lik, d_beam, n_beam, d_origins, n_origins, stream=0):
        """This rank's shard + the record exchange folded into the kernels (raw device addresses, async on `stream`).
        Returns the device address that holds ALL ranks' records once the stream gets there (eager calls)."""
        d_all = C.c_void_p()
        self._check(self.L.mcl3dl_measure_exchange_device(self.h, d_poses, n_local, d_lik, n_lik, d_beam, n_beam,
                                                          d_origins, n_origins, stream, C.byref(d_all)))
        return d_all.value

    def exchange_current(self, stream=0):
        """(device address of the last completed gathered array, a peer timed out?) — synchronises `stream`."""
        d_all, f = C.c_void_p(), C.c_int(0)
        self._check(self.L.mcl3dl_exchange_current(self.h, stream, C.byref(d_all), C.byref(f)))
        return d_all.value, bool(f.value)

    def collect_timing(self, enable=True):
        """Record the per-call CUDA timing events read by last_timing() (off by default: ~28 us per update)."""
        self._check(self.L.mcl3dl_collect_timing(self.h, 1 if enable else 0))

    def last_timing(self):
        v = [C.c_double(0) for _ in range(4)]
        self.L.mcl3dl_last_timing(self.h, *[C.byref(x) for x in v])
        return {"h2d_ms": v[0].value, "lik_ms": v[1].value, "beam_ms": v[2].value, "d2h_ms": v[3].value}

    def kernel_launches(self):
        return int(self.L.mcl3dl_kernel_launches(self.h))
