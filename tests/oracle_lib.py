"""ctypes wrapper of the CPU oracle (oracle/libacn_oracle.so) and of the reference's own leaf math
(oracle/_ref/libacn_refleaf.so).  TEST INFRASTRUCTURE: imported by tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs only — never by the product package."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "libacn_oracle.so")
REFLEAF_SO = os.path.join(ORACLE_DIR, "_ref", "libacn_refleaf.so")

COUNTER_NAMES = [
    "sphere_miss", "sphere_hit", "sphere_nor", "plane", "squaroid", "squaroid_nor", "dist_step", "dist_nor",
    "side_sphere", "side_plane", "side_squaroid", "side_dist",
    "fresnel", "refract", "mirror", "cap_sample", "oren_nayar", "direct_book", "roughness", "camera", "gamma", "absorb",
    "rays_primary", "rays_reflect", "rays_chromatic", "rays_refract", "rays_path", "rays_shadow", "rays_lighthit", "diffuse_hits",
]


def build_oracle(force: bool = False):
    """Compiles the oracle (g++) and, when /root/reference exists, the reference leaf library."""
    if force or not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(os.path.join(ORACLE_DIR, "acn_oracle.cpp")):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "libacn_oracle.so"], stdout=subprocess.DEVNULL)
    if os.path.isdir("/root/reference/src") and (force or not os.path.exists(REFLEAF_SO)):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "ref"], stdout=subprocess.DEVNULL)


def _v3(v):
    return (C.c_double * 3)(float(v[0]), float(v[1]), float(v[2]))


class Oracle:
    def __init__(self):
        build_oracle()
        self.lib = C.CDLL(ORACLE_SO)
        L = self.lib
        L.oracle_render.restype = C.c_int
        L.oracle_render.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_int, C.c_double, C.c_int,
                                    C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_double), C.c_void_p]
        L.oracle_accumulate.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_void_p]
        L.oracle_counter_count.restype = C.c_int
        L.oracle_flops_per_event.restype = C.c_double
        L.oracle_flops_per_event.argtypes = [C.c_int]
        L.oracle_sf_per_event.restype = C.c_double
        L.oracle_sf_per_event.argtypes = [C.c_int]
        pd = C.POINTER(C.c_double)
        L.oracle_sphere_ray_hit.restype = C.c_double
        L.oracle_sphere_ray_hit.argtypes = [pd, C.c_double, pd, pd, C.c_double, pd]
        L.oracle_plane_ray_hit.restype = C.c_double
        L.oracle_plane_ray_hit.argtypes = [pd, pd, pd, pd, C.c_double]
        L.oracle_fresnel_reflection.restype = C.c_double
        L.oracle_fresnel_reflection.argtypes = [pd, pd, C.c_double, pd]
        L.oracle_fresnel_refraction.argtypes = [pd, pd, C.c_double, pd]
        L.oracle_con_z.argtypes = [pd, pd]
        L.oracle_random_seed.restype = C.c_uint64
        L.oracle_random_seed.argtypes = [pd, C.c_uint64]
        L.oracle_random_sphere_cap.argtypes = [C.POINTER(C.c_uint64), C.c_double, pd]
        L.oracle_cl_sat.argtypes = [pd, C.c_double, pd]
        L.oracle_oren_nayar_weight.restype = C.c_double
        L.oracle_oren_nayar_weight.argtypes = [C.c_double, C.c_double, C.c_double, C.c_double, pd, pd, pd]
        self.n_counters = L.oracle_counter_count()
        assert self.n_counters == len(COUNTER_NAMES)
        self.flops_per_event = np.array([L.oracle_flops_per_event(k) for k in range(self.n_counters)])
        self.sf_per_event = np.array([L.oracle_sf_per_event(k) for k in range(self.n_counters)])

    def render(self, flat, xy, index_base=0, seed_mode=0, eps=1e-6, threads=None, want_linear=False):
        """lum_machine_s_run on the CPU.  flat: actinon_b200.FlatScene.  Returns (rgb[n,3] float64, info)."""
        xy = np.ascontiguousarray(xy, dtype=np.float64).reshape(-1, 2)
        n = xy.shape[0]
        rgb = np.zeros((n, 3), dtype=np.float64)
        lin = np.zeros((n, 3), dtype=np.float64) if want_linear else None
        cnt = np.zeros(self.n_counters, dtype=np.uint64)
        sec = C.c_double(0)
        phase = np.zeros(4, dtype=np.float64)
        if threads is None:
            threads = os.cpu_count() or 1
        rc = self.lib.oracle_render(C.cast(flat.ptr, C.c_void_p), xy.ctypes.data, n, index_base, seed_mode, eps, threads,
                                    rgb.ctypes.data, lin.ctypes.data if lin is not None else None, cnt.ctypes.data, C.byref(sec), phase.ctypes.data)
        if rc != 0:
            raise RuntimeError(f"oracle_render failed: {rc}")
        counters = {k: int(v) for k, v in zip(COUNTER_NAMES, cnt)}
        info = {
            "seconds": sec.value, "threads": threads, "counters": counters,
            "flops": float((cnt.astype(np.float64) * self.flops_per_event).sum()),
            "sf_ops": float((cnt.astype(np.float64) * self.sf_per_event).sum()),
            "rays": int(sum(counters[k] for k in ("rays_primary", "rays_reflect", "rays_chromatic", "rays_refract", "rays_path", "rays_shadow"))),
            "linear": lin,
            # algorithmic FLOPs by phase of the ray tree = by tracer kernel: primary, spec rays, path rays, direct loop
            "phase_flops": {"primary": phase[0], "rays": phase[1], "path": phase[2], "direct": phase[3]},
        }
        return rgb, info

    def accumulate(self, xy, rgb, width, height):
        xy = np.ascontiguousarray(xy, dtype=np.float64); rgb = np.ascontiguousarray(rgb, dtype=np.float64)
        sums = np.zeros((height, width, 4), dtype=np.float64)
        self.lib.oracle_accumulate(xy.ctypes.data, rgb.ctypes.data, xy.shape[0], width, height, sums.ctypes.data)
        return sums

    # leaf functions
    def sphere_ray_hit(self, pos, r, rp, rd, eps=1e-6):
        nor = (C.c_double * 3)()
        a = self.lib.oracle_sphere_ray_hit(_v3(pos), r, _v3(rp), _v3(rd), eps, nor)
        return a, np.array(nor[:])

    def plane_ray_hit(self, pos, n, rp, rd, eps=1e-6):
        return self.lib.oracle_plane_ray_hit(_v3(pos), _v3(n), _v3(rp), _v3(rd), eps)

    def fresnel_reflection(self, d, n, trix):
        o = (C.c_double * 3)()
        r = self.lib.oracle_fresnel_reflection(_v3(d), _v3(n), trix, o)
        return r, np.array(o[:])

    def fresnel_refraction(self, d, n, trix):
        o = (C.c_double * 3)()
        self.lib.oracle_fresnel_refraction(_v3(d), _v3(n), trix, o)
        return np.array(o[:])

    def con_z(self, v):
        m = (C.c_double * 9)()
        self.lib.oracle_con_z(_v3(v), m)
        return np.array(m[:]).reshape(3, 3)

    def random_seed(self, v, rv):
        return self.lib.oracle_random_seed(_v3(v), rv)

    def random_sphere_cap(self, rv, h):
        s = C.c_uint64(rv); o = (C.c_double * 3)()
        self.lib.oracle_random_sphere_cap(C.byref(s), h, o)
        return np.array(o[:]), s.value

    def cl_sat(self, c, gamma):
        o = (C.c_double * 3)()
        self.lib.oracle_cl_sat(_v3(c), gamma, o)
        return np.array(o[:])

    def oren_nayar_weight(self, weight, theta_i, a, b, out_d, nor, prj):
        return self.lib.oracle_oren_nayar_weight(weight, theta_i, a, b, _v3(out_d), _v3(nor), _v3(prj))


class RefLeaf:
    """The reference's own vectors.h / gmath.h / gmath.c (compiled in place; see oracle/Makefile)."""

    def __init__(self):
        if not os.path.exists(REFLEAF_SO):
            build_oracle()
        if not os.path.exists(REFLEAF_SO):
            raise FileNotFoundError(REFLEAF_SO)
        self.lib = C.CDLL(REFLEAF_SO)
        L = self.lib
        pd = C.POINTER(C.c_double)
        L.ref_sphere_ray_hit.restype = C.c_double
        L.ref_sphere_ray_hit.argtypes = [pd, C.c_double, pd, pd, pd]
        L.ref_plane_ray_hit.restype = C.c_double
        L.ref_plane_ray_hit.argtypes = [pd, pd, pd, pd]
        L.ref_fresnel_reflection.restype = C.c_double
        L.ref_fresnel_reflection.argtypes = [pd, pd, C.c_double, pd]
        L.ref_fresnel_refraction.argtypes = [pd, pd, C.c_double, pd]
        L.ref_con_z.argtypes = [pd, pd]
        L.ref_random_seed.restype = C.c_uint64
        L.ref_random_seed.argtypes = [pd, C.c_uint64]
        L.ref_seed_from_f3.restype = C.c_uint64
        L.ref_seed_from_f3.argtypes = [C.c_double]
        L.ref_random_sphere_cap.argtypes = [C.POINTER(C.c_uint64), C.c_double, pd]
        L.ref_cl_sat.argtypes = [pd, C.c_double, pd]
        L.ref_rnd0.restype = C.c_double
        L.ref_rnd0.argtypes = [C.POINTER(C.c_uint64)]
        L.ref_rnd1.restype = C.c_double
        L.ref_rnd1.argtypes = [C.POINTER(C.c_uint64)]
        L.ref_eps.restype = C.c_double
        L.ref_sphere_observer_side.argtypes = [pd, C.c_double, pd]
        L.ref_plane_observer_side.argtypes = [pd, pd, pd]
        L.ref_reflection.argtypes = [pd, pd, pd]
        L.ref_con.argtypes = [pd, pd]
        L.ref_of_length.argtypes = [pd, C.c_double, pd]

    def sphere_ray_hit(self, pos, r, rp, rd):
        nor = (C.c_double * 3)()
        a = self.lib.ref_sphere_ray_hit(_v3(pos), r, _v3(rp), _v3(rd), nor)
        return a, np.array(nor[:])

    def plane_ray_hit(self, pos, n, rp, rd):
        return self.lib.ref_plane_ray_hit(_v3(pos), _v3(n), _v3(rp), _v3(rd))

    def fresnel_reflection(self, d, n, trix):
        o = (C.c_double * 3)()
        r = self.lib.ref_fresnel_reflection(_v3(d), _v3(n), trix, o)
        return r, np.array(o[:])

    def fresnel_refraction(self, d, n, trix):
        o = (C.c_double * 3)()
        self.lib.ref_fresnel_refraction(_v3(d), _v3(n), trix, o)
        return np.array(o[:])

    def con_z(self, v):
        m = (C.c_double * 9)()
        self.lib.ref_con_z(_v3(v), m)
        return np.array(m[:]).reshape(3, 3)

    def random_seed(self, v, rv):
        return self.lib.ref_random_seed(_v3(v), rv)

    def random_sphere_cap(self, rv, h):
        s = C.c_uint64(rv); o = (C.c_double * 3)()
        self.lib.ref_random_sphere_cap(C.byref(s), h, o)
        return np.array(o[:]), s.value

    def cl_sat(self, c, gamma):
        o = (C.c_double * 3)()
        self.lib.ref_cl_sat(_v3(c), gamma, o)
        return np.array(o[:])
