"""Minimal scene description replacing the Mitsuba objects the integrators query on the hot path:
perspective sensor + hdrfilm(gaussian) (python/util.py:115-138), a `constant` emitter and the diffuse
BSDF of the dummy SDF shape (constant reflectance or a trilinear reflectance grid)."""
import math

import numpy as np


def look_at(origin, target, up):
    """mitsuba ScalarTransform4f.look_at: columns (left, new_up, dir, origin)."""
    origin, target, up = (np.asarray(v, dtype=np.float64) for v in (origin, target, up))
    d = target - origin
    d /= np.linalg.norm(d)
    left = np.cross(up, d)
    left /= np.linalg.norm(left)
    new_up = np.cross(d, left)
    m = np.eye(4)
    m[:3, 0], m[:3, 1], m[:3, 2], m[:3, 3] = left, new_up, d, origin
    return m


def get_regular_camera_positions(angle_steps, radius=2.0, angle_shift=0.0, height_scale=1.0):
    """python/util.py:84-112 for the branch get_regular_cameras uses (height_steps <= 1, hemisphere, vary_height)."""
    n = angle_steps
    angles = (np.arange(n) / n + angle_shift / n) * 2 * np.pi
    elevation = 1.15 / height_scale + np.sin(angles * angle_steps / 4) * 0.5
    elevation = np.clip(elevation, 0.0, np.pi / 2 + 0.05)
    origins = np.stack([np.cos(angles) * np.sin(elevation) * radius, np.cos(elevation) * radius,
                        np.sin(angles) * np.sin(elevation) * radius], -1)
    return origins + np.array([0.5, 0.0, 0.5])


class Film:
    def __init__(self, width, height, filter_radius=2.0, filter_stddev=0.5, sample_border=True):
        self.width, self.height = width, height
        self.filter_radius, self.filter_stddev, self.sample_border = filter_radius, filter_stddev, sample_border

    def crop_size(self):
        return (self.width, self.height)

    def size(self):
        return (self.width, self.height)


class Sensor:
    def __init__(self, to_world, width=128, height=128, fov=39.0, near_clip=1e-2, far_clip=1e4):
        self.to_world = np.asarray(to_world, dtype=np.float64)
        self.fov, self.near_clip, self.far_clip = fov, near_clip, far_clip
        self._film = Film(width, height)

    def film(self):
        return self._film


def set_sensor_res(sensor, res):
    """python/util.py:146-150."""
    sensor.film().width, sensor.film().height = int(res[0]), int(res[1])


def get_regular_cameras(n_sensors, angle_shift=0.0, resx=128, resy=128, radius=2.0, height_scale=1.0):
    """python/util.py:115-138."""
    origins = get_regular_camera_positions(n_sensors, radius=radius, angle_shift=angle_shift, height_scale=height_scale)
    return [Sensor(look_at(o, [0.5, 0.5, 0.5], [0, 1, 0]), resx, resy) for o in origins]


class Scene:
    """Holds what `mi.load_file(...)` would: integrator, sensors, emitter and the dummy shape's BSDF."""

    def __init__(self, integrator, sensors, env_radiance=(1.0, 1.0, 1.0), albedo=(0.5, 0.5, 0.5), albedo_grid=None,
                 env_dist=2.0 * math.sqrt(3.0)):
        self._integrator = integrator
        self._sensors = list(sensors)
        self.env_radiance, self.albedo, self.albedo_grid, self.env_dist = env_radiance, albedo, albedo_grid, env_dist
        self.envmap, self.envmap_scale, self.envmap_version, self.envmap_to_world = None, 1.0, 0, None

    def set_envmap(self, rgb, scale=1.0, to_world=None):
        """Light the scene with an `envmap` emitter ((H, W, 3) lat-long bitmap, what the reference's scene files load from an .exr)
        instead of the constant one; None switches back.  Call again after changing the array."""
        self.envmap, self.envmap_scale, self.envmap_to_world = rgb, scale, to_world
        self.envmap_version += 1

    def integrator(self):
        return self._integrator

    def sensors(self):
        return self._sensors
