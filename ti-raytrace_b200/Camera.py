"""Camera — pinhole orbit camera (mirror of /root/reference/Camera.py:13-118).

Host maths as in the reference (numpy, f64 -> f32 view / view_inv); the device part
(get_ray_origin / get_ray_direction, Camera.py:122-142) is `camera_dir` in csrc/wavefront.cu.  The
matrices are pushed with tr_camera_set() by the integrator before each launch when they changed."""
import math
import numpy as np
import _native

FULL_HGT = 2.4      # full-frame sensor height


class Camera:
    def __init__(self, sizex, sizey, sample_count):
        self.wid, self.hgt = sizex, sizey
        self.focal = 2.0
        self.ratio = sizex / sizey
        self.fx = self.focal * sizex / FULL_HGT
        self.fy = self.fx
        self.cx, self.cy = sizex * 0.5, sizey * 0.5
        self.eye_np = np.ones((1, 3), np.float32)
        self.target = np.array([0.0, 0.0, 0.0])
        self.up = np.array([0.0, 1.0, 0.0])
        self.yaw = self.pitch = self.roll = 0.0
        self.scale = 1000.0
        self.sample_count = int(math.sqrt(sample_count))
        self.sample_dis = 1.0 / float(self.sample_count - 1)     # ZeroDivisionError for spp < 4, as upstream
        self.frame_cpu = np.zeros(1, np.int32)
        self.frame = 0
        self.fps = 30.0
        self.view_np = np.zeros((1, 4, 4), np.float32)
        self.view_inv_np = np.zeros((1, 4, 4), np.float32)
        self.dirty = True
        self.frame_gpu = _native.Field(lambda: self.frame_cpu.copy())
        self.view = _native.Field(lambda: self.view_np.copy())
        self.view_inv = _native.Field(lambda: self.view_inv_np.copy())
        self.eye = _native.Field(lambda: self.eye_np.copy())

    def yaw_cam(self, targetx, targety, targetz):
        self.target[:] = (targetx, targety, targetz)
        if self.yaw < 3.14:
            self.set_view_point(self.yaw + 0.003, 0.0, 0.0, 3.0)

    def pitch_cam(self, targetx, targety, targetz):
        self.target[:] = (targetx, targety, targetz)
        if self.pitch < 0.5:
            self.set_view_point(0.0, self.pitch + 0.003, 0.0, 3.0)

    def update(self):
        self.pitch = max(min(self.pitch, 1.57), -1.57)
        cp, sp, cy, sy = math.cos(self.pitch), math.sin(self.pitch), math.cos(self.yaw), math.sin(self.yaw)
        self.eye_np[0, 0] = self.target[0] + self.scale * cp * sy
        self.eye_np[0, 1] = self.target[1] + self.scale * sp
        self.eye_np[0, 2] = self.target[2] + self.scale * cp * cy
        self.up[:] = (-sp * sy, cp, -sp * cy)
        eye = self.eye_np[0, :]
        zaxis = eye - self.target
        zaxis = zaxis / np.linalg.norm(zaxis)
        xaxis = np.cross(self.up, zaxis)
        xaxis = xaxis / np.linalg.norm(xaxis)
        yaxis = np.cross(zaxis, xaxis)
        rows = [list(a) + [-np.dot(a, eye)] for a in (xaxis, yaxis, zaxis)]
        self.view_np[0] = np.array(rows + [[0.0, 0.0, 0.0, 1.0]])
        self.view_inv_np[:] = np.linalg.inv(self.view_np)
        self.dirty = True

    def set_view_point(self, yaw, pitch, roll, scale):
        self.pitch, self.yaw, self.roll, self.scale = pitch, yaw, roll, scale
        self.update()

    def set_target(self, targetx, targety, targetz):
        self.target[:] = (targetx, targety, targetz)
        self.update()

    def update_frame(self, n=1):
        self.frame += n
        self.frame_cpu[0] = self.frame

    def push(self, ctx):
        if self.dirty:
            ctx.camera_set(self.view_np[0], self.view_inv_np[0], self.eye_np[0], self.fx, self.fy, self.cx, self.cy)
            self.dirty = False
