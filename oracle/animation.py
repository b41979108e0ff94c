"""TEST INFRASTRUCTURE ONLY -- part of the CPU oracle (see oracle/README.md); nothing under
portal_b200/ may import this module.

Restatement of the reference's animation layer between a scene file and one frame's uniforms:

* cameras (`Cam::get`, /root/reference/src/gui/camera.rs:111-140),
* stage / real-animation initialisation (`Scene::init_stage`, src/gui/scene.rs:1180-1236;
  `StageChanging::init_stage` animation.rs:171-183; `DevStageChanging::init_stage` :223-227;
  `RealAnimationStageChanging::init_stage` :985-991),
* the time mapping and camera interpolation of `Scene::update` (scene.rs:1353-1496) with
  `get_start_cam` / `get_end_cam` (:1291-1343),
* the renderer side of it: `SceneRenderer::update` (src/main.rs:1430-1543), `teleport_camera`
  (:1217-1264) and the finite-difference `teleport_matrix` (:1174-1215) on top of the external-ray
  probe (`teleport_external_ray`, :1361-1409),
* `render_frame` (main.rs:2876-2929) and the frame/sub-frame times of `render_animation` (:1757-1830).

Everything is f64 like the reference (glam DVec3/DMat4, egui::lerp).  The egui `Memory` the reference
threads through these functions is the `memory` dict here (keys: CurrentCam, OriginalCam, OverrideCam).
parity unpinned: upstream has no tests or golden values for any of this.
"""
from __future__ import annotations

import copy
import math

from . import frontend as F


def _f(x):
    return float(x)


def egui_lerp(a, b, t):
    """emath::lerp(a..=b, t) = (1 - t) * a + t * b."""
    return (1.0 - t) * a + t * b


def _fmod(a, b):
    """Rust `%` on f64 (C fmod)."""
    if b == 0.0 or math.isinf(a) or math.isnan(a) or math.isnan(b):
        return math.nan
    return math.fmod(a, b)


def project_origin(m):
    """glam 0.13.1 DMat4::project_point3(DVec3::ZERO): x*0 + y*0 + z*0 + w, times 1/res.w."""
    res = [c * 0.0 for c in m[0]]
    res = [m[1][i] * 0.0 + res[i] for i in range(4)]
    res = [m[2][i] * 0.0 + res[i] for i in range(4)]
    res = [m[3][i] + res[i] for i in range(4)]
    r = F._rcp(res[3])
    return [res[0] * r, res[1] * r, res[2] * r]


EASINGS = {
    "Linear": lambda t: t,
    "In": F.easing_in,
    "Out": F.easing_out,
    "InOut": F.easing_in_out,
    "InOutFast": F.easing_in_out_fast,
    "ElasticOut": F.easing_elastic_out,
}


class AnimatedScene:
    """A `Scene` (oracle/frontend.py) plus its cameras, stage cameras and real animations."""

    def __init__(self, scene: F.Scene):
        self.scene = scene
        ser = scene.ser
        self.cameras: list[dict] = []
        self.camera_by_name: dict[str, int] = {}
        for item in ser.get("cameras", [[]])[0]:
            self.camera_by_name.setdefault(item["name"], len(self.cameras))
            self.cameras.append(self._cam_from_ser(item["data"]))
        # stage cameras (scene_serialized.rs:1336-1352)
        self.stage_cam: dict[str, tuple] = {}
        for item in ser.get("animation_stages", [[]])[0]:
            sc = item["data"].get("set_cam")
            if sc is None:
                self.stage_cam[item["name"]] = ("keep",)
            else:
                self.stage_cam[item["name"]] = ("set", self._cam_ref(sc))
        # real animations, in file (= visible) order (scene_serialized.rs:1374-1473)
        self.animations: list[dict] = []
        self.animation_by_name: dict[str, int] = {}
        for item in ser.get("animations", [[]])[0]:
            self.animation_by_name.setdefault(item["name"], len(self.animations))
            self.animations.append({"name": item["name"]})
        for idx, item in enumerate(ser.get("animations", [[]])[0]):
            self.animations[idx].update(self._animation_from_ser(item["data"]))
        self.current_stage = self._stage_from_ser(ser.get("current_stage"))
        self.run_animations = False
        self.prev_t_raw = 0.0

    # ------------------------------------------------------------------ loading
    def _cam_from_ser(self, d) -> dict:
        la = d["look_at"]
        if la.tag == "Coordinate":
            look_at = ("coord", [_f(x) for x in la.value[0]])
        else:
            look_at = ("matrix", self.scene._matrix_ref(la.value[0]))
        m = d.get("matrix")
        matrix = F.mat_identity() if m is None else [[_f(m[4 * c + r]) for r in range(4)] for c in range(4)]
        return {"look_at": look_at, "alpha": _f(d["alpha"]), "beta": _f(d["beta"]), "r": _f(d["r"]),
                "in_subspace": bool(d.get("in_subspace", False)), "free_movement": bool(d.get("free_movement", False)),
                "matrix": matrix}

    def _cam_ref(self, ref):
        if ref is None:
            return None
        if ref.tag == "Named":
            return self.camera_by_name.get(ref.value[0])
        if ref.tag == "Inline":
            self.cameras.append(self._cam_from_ser(ref.value[0]))
            return len(self.cameras) - 1
        raise ValueError(ref)

    def _stage_from_ser(self, s):
        if s is None or s.tag == "Dev":
            return ("dev",)
        if s.tag == "Animation":
            return ("animation", s.value[0]) if s.value[0] in self.scene.stages else ("dev",)
        if s.tag == "RealAnimation":
            return ("real", s.value[0]) if s.value[0] in self.animation_by_name else ("dev",)
        raise ValueError(s)

    def _animation_from_ser(self, d) -> dict:
        sc = self.scene

        def part(p, resolve):
            if p.tag == "CopyPrev":
                return ("copy",)
            return ("set", resolve(p.value[0]))

        a = {
            "duration": _f(d.get("duration", 0.0)),
            "stage": self._stage_from_ser(d["animation_stage"]),
            "uniforms": {k: part(v, sc._uniform_ref) for k, v in d["uniforms"][0].items() if k in sc.uniform_by_name},
            "matrices": {k: part(v, sc._matrix_ref) for k, v in d["matrices"][0].items() if k in sc.matrix_by_name},
            "use_prev_cam": bool(d.get("use_prev_cam", False)),
            "use_start_cam_as_end": bool(d.get("use_start_cam_as_end", False)),
            "cam_start": self._cam_ref(d.get("cam_start")),
            "cam_end": self._cam_ref(d.get("cam_end")),
            "use_any_cam_as_start": d.get("use_any_cam_as_start"),
            "use_any_cam_as_end": d.get("use_any_cam_as_end"),
            "cam_any_start": self.animation_by_name.get(d.get("cam_any_start")),
            "cam_any_end": self.animation_by_name.get(d.get("cam_any_end")),
            "cam_easing": d["cam_easing"].tag if d.get("cam_easing") is not None else "Linear",
            "cam_easing_uniform": None,
        }
        ceu = d.get("cam_easing_uniform")
        if ceu is not None:
            uid = sc._uniform_ref(ceu)
            a["cam_easing_uniform"] = None if uid is None else ("some", uid)   # scene_serialized.rs:1468-1471
        return a

    # ------------------------------------------------------------------ cameras
    def cam_get(self, cam_id):
        """Cam::get (camera.rs:111-140) -> CalculatedCam or None."""
        c = self.cameras[cam_id]
        if c["look_at"][0] == "coord":
            look_at = list(c["look_at"][1])
        else:
            mid = c["look_at"][1]
            m = self.scene.get_matrix(mid)
            if m is None:
                return None
            p = project_origin(m)
            look_at = [p[0] + 0.001, p[1] + 0.001, p[2] + 0.001]
        return {"look_at": look_at, "alpha": c["alpha"], "beta": c["beta"], "r": c["r"], "in_subspace": c["in_subspace"],
                "free_movement": c["free_movement"], "matrix": c["matrix"], "override_matrix": True}

    def get_start_cam(self, idx, depth=0):
        """scene.rs:1291-1317.  (depth: animations whose cameras refer to each other in a cycle would overflow the
        reference's stack; here, as in the C++ player, the chain gives up after 256 links -> no camera.)"""
        if depth > 256:
            return None
        anim = self.animations[idx]
        if anim["use_prev_cam"]:
            return self.get_end_cam(idx - 1, depth + 1) if idx >= 1 else None
        if anim["use_any_cam_as_start"] is not None:
            any_id = anim["cam_any_start"]
            if any_id is None:
                return None
            return self.get_end_cam(any_id, depth + 1) if anim["use_any_cam_as_start"] else self.get_start_cam(any_id, depth + 1)
        return anim["cam_start"]

    def get_end_cam(self, idx, depth=0):
        """scene.rs:1319-1335."""
        if depth > 256:
            return None
        anim = self.animations[idx]
        if anim["use_start_cam_as_end"]:
            return self.get_start_cam(idx, depth + 1)
        if anim["use_any_cam_as_end"] is not None:
            any_id = anim["cam_any_end"]
            if any_id is None:
                return None
            return self.get_end_cam(any_id, depth + 1) if anim["use_any_cam_as_end"] else self.get_start_cam(any_id, depth + 1)
        return anim["cam_end"]

    # ------------------------------------------------------------------ stages
    def init_stage(self, stage, memory: dict):
        """Scene::init_stage (scene.rs:1180-1236)."""
        sc = self.scene
        if stage[0] == "animation":
            sc.init_stage(stage[1])
            cam = self.stage_cam.get(stage[1], ("keep",))
            memory["CurrentCam"] = cam[1] if cam[0] == "set" else None
        elif stage[0] == "dev":
            for uname, value in sc.dev_uniforms.items():
                sc.uniforms[sc.uniform_by_name[uname]] = value
            for mname, value in sc.dev_matrices.items():
                sc.matrices[sc.matrix_by_name[mname]] = value
            memory["CurrentCam"] = None
        else:
            idx = self.animation_by_name[stage[1]]
            anim = self.animations[idx]
            if anim["stage"] != stage:
                self.init_stage(anim["stage"], memory)
            for uname, p in anim["uniforms"].items():
                if p[0] == "set" and p[1] is not None:
                    sc.uniforms[sc.uniform_by_name[uname]] = sc.uniforms[p[1]]
            for mname, p in anim["matrices"].items():
                if p[0] == "set" and p[1] is not None:
                    sc.matrices[sc.matrix_by_name[mname]] = sc.matrices[p[1]]
            cam_start = self.get_start_cam(idx)
            if cam_start is not None:
                memory["CurrentCam"] = cam_start
        self.current_stage = stage

    def init_stage_by_name(self, name, memory):
        if name not in self.scene.stages:
            raise KeyError(f"no stage named `{name}`")
        self.init_stage(("animation", name), memory)

    def init_animation_by_name(self, name, memory):
        if name not in self.animation_by_name:
            raise KeyError(f"no animation named `{name}`")
        self.init_stage(("real", name), memory)

    def total_animation_duration(self):
        s = 0.0
        for a in self.animations:
            s += a["duration"]
        return s

    # ------------------------------------------------------------------ per-frame update
    def update(self, memory: dict, time: float):
        """Scene::update (scene.rs:1353-1496): sets the formula variables `time` / `total_time` and, inside a
        real animation, leaves the interpolated camera in memory['OverrideCam']."""
        sc = self.scene
        if self.run_animations:
            total = self.total_animation_duration()
            if total > 0.0:
                time = _fmod(time, total)
                total_time = time
            else:
                time = 0.0
                total_time = 0.0
            for idx, a in enumerate(self.animations):
                if time < a["duration"]:
                    if self.current_stage != ("real", a["name"]):
                        self.init_stage(("real", a["name"]), memory)
                    time /= a["duration"]
                    break
                time -= a["duration"]
        elif self.current_stage[0] == "real":
            idx = self.animation_by_name[self.current_stage[1]]
            duration = self.animations[idx]["duration"]
            if duration > 0.0:
                local_seconds = _fmod(time, duration)
                time = local_seconds / duration
                prefix = 0.0
                for a in self.animations[:idx]:
                    prefix += a["duration"]
                total_time = prefix + local_seconds
            else:
                time = 0.0
                total_time = 0.0
        else:
            total_time = time
        sc.time = time
        sc.total_time = total_time

        if self.current_stage[0] == "real":
            idx = self.animation_by_name[self.current_stage[1]]
            anim = self.animations[idx]
            c1, c2 = self.get_start_cam(idx), self.get_end_cam(idx)
            if c1 is not None and c2 is not None:
                cam1, cam2 = self.cam_get(c1), self.cam_get(c2)
                if cam1 is None or cam2 is None:
                    raise ValueError("camera looks at a missing matrix")     # the reference unwrap()s here
                t_raw = _fmod(sc.time, 1.0)
                t = None
                if anim["cam_easing_uniform"] is not None:
                    r = sc.get_uniform(anim["cam_easing_uniform"][1])
                    if r is not None:
                        v = float(r[1])
                        if not math.isfinite(v):
                            v = 0.0
                        t = min(max(v, 0.0), 1.0)
                if t is None:
                    t = EASINGS[anim["cam_easing"]](t_raw)
                override_matrix = t_raw < self.prev_t_raw or t_raw == 0.0
                memory["OverrideCam"] = {
                    "look_at": F.vec_lerp(cam1["look_at"], cam2["look_at"], t),
                    "alpha": egui_lerp(cam1["alpha"], cam2["alpha"], t),
                    "beta": egui_lerp(cam1["beta"], cam2["beta"], t),
                    "r": egui_lerp(cam1["r"], cam2["r"], t),
                    "in_subspace": cam1["in_subspace"], "free_movement": cam1["free_movement"],
                    "matrix": cam1["matrix"], "override_matrix": override_matrix,
                }
                self.prev_t_raw = t_raw


class OrbitCam:
    """The parts of RotateAroundCam (main.rs:40-160, 278-330) the offline paths touch."""

    def __init__(self):
        self.look_at = [0.0, 0.0, 0.0]
        self.alpha = 81.0 / 180.0 * F.PI          # deg2rad(81.)
        self.beta = 64.0 / 180.0 * F.PI
        self.r = 3.5
        self.from_cam = None
        self.teleport_matrix = F.mat_identity()
        self.allow_teleport = True
        self.stop_at_objects = False
        self.prev_cam_pos = [0.0, 0.0, 0.0]
        self.in_subspace = False
        self.free_movement = False
        self.do_not_teleport_one_frame = False

    def clone(self):
        return copy.deepcopy(self)

    def get_pos_vec(self):
        return [math.sin(self.beta) * math.cos(self.alpha) * self.r, math.cos(self.beta) * self.r,
                math.sin(self.beta) * math.sin(self.alpha) * self.r]

    def get_matrix(self):
        return F.orbit_camera_matrix(self.look_at, self.alpha, self.beta, self.r, self.teleport_matrix, self.free_movement)

    def get_cam_pos(self):
        p = F.mat_mul_vec(self.get_matrix(), [0.0, 0.0, 0.0, 1.0])
        return p[:3]

    def set_cam(self, s):
        """main.rs:320-330."""
        self.look_at = [_f(x) for x in s["look_at"]]
        self.alpha, self.beta, self.r = _f(s["alpha"]), _f(s["beta"]), _f(s["r"])
        if self.free_movement:
            pv = self.get_pos_vec()
            self.look_at = [pv[i] + self.look_at[i] for i in range(3)]

    def get_calculated_cam(self):
        return {"look_at": list(self.look_at), "alpha": self.alpha, "beta": self.beta, "r": self.r,
                "in_subspace": self.in_subspace, "free_movement": self.free_movement,
                "matrix": self.teleport_matrix, "override_matrix": True}


class Player:
    """SceneRenderer's animation-facing half: new() (main.rs:1014-1062), update() (:1430-1543),
    teleport_camera() (:1217-1264), teleport_matrix() (:1174-1215).

    `probe(player, a, b) -> (pos | None, encounter_object, change_subspace)` is teleport_external_ray
    (:1361-1409) evaluated with the player's current uniforms and camera; without one the camera never
    teleports (frames whose camera does not cross a portal are unaffected)."""

    def __init__(self, scene: F.Scene, probe=None):
        self.anim = AnimatedScene(scene)
        self.scene = scene
        self.probe = probe
        self.cam = OrbitCam()
        self.prev_cam = OrbitCam()
        self.cam.set_cam(scene.cam)
        self.cam.prev_cam_pos = self.cam.get_cam_pos()
        self.memory = {"CurrentCam": None, "OriginalCam": self.cam.get_calculated_cam()}
        self.n_probes = 0
        # stereo (main.rs:1027-1030): eye cameras are placed by teleport_eye_matrices when a stereo mode is on
        self.draw_side_by_side = False
        self.draw_anaglyph = False
        self.eye_distance = 0.07
        self.swap_eyes = False
        self.left_eye_matrix, self.right_eye_matrix = F.mat_identity(), F.mat_identity()
        self.left_eye_in_subspace = self.right_eye_in_subspace = False

    # -- render_frame's option handling (main.rs:2900-2926)
    def init_stage(self, name):
        self.anim.init_stage_by_name(name, self.memory)

    def init_animation(self, name):
        self.anim.init_animation_by_name(name, self.memory)
        self.update_inner_variables(name)          # render-frame, main.rs:2906-2917

    _DEGREE_500 = ("v2.face.2", "v2.face.3", "v2.face.4", "v2.face.5", "v2.inside.1", "v2.inside.3", "v2.intro.1", "v2.normal.2",
                   "v2.normal.3", "v2.rod.2", "v2.rod.3", "v2.spiral.3", "v2.spiral.4", "v2.spiral.5", "v2.spiral.6", "v2.spiral.7",
                   "v2.spiral.9", "v2.spaaaace.0", "v4.golden.0", "v4.golden.1", "v4.golden.2", "v4.thumbnail.2")

    def update_inner_variables(self, name):
        """SceneRenderer::update_inner_variables (main.rs:1696-1755): overrides hard-coded by animation name."""
        from .ron import Tagged
        self.render_depth_override = self.fps_override = 0

        def set_degree(v):
            uid = self.scene.uniform_by_name.get("subspace_degree")
            if uid is None:
                return False                       # `find_id(..)?` leaves the function
            self.scene.uniforms[uid] = Tagged("Int", [{"min": None, "max": None, "value": v}])
            return True
        if name in self._DEGREE_500 and not set_degree(500):
            return
        if name in ("v2.spiral.4", "v2.spiral.5", "v2.spiral.6") and not set_degree(1000):
            return
        if name in ("v2.rotated.0", "v2.spiral.0", "v2.screenshot.5", "v2.screenshot.6"):
            self.render_depth_override = 100
        if name == "v2.screenshot.3":
            self.fps_override = 600

    def select_camera(self, name):
        if name not in self.anim.camera_by_name:
            raise KeyError(f"no camera named `{name}`")
        self.memory["CurrentCam"] = self.anim.camera_by_name[name]

    def _probe(self, a, b):
        self.n_probes += 1
        return self.probe(self, a, b)

    def teleport_matrix(self, matrix, start_pos, direction_pos, actual, dx):
        cols = []
        for axis in range(3):
            e = [0.0, 0.0, 0.0, 0.0]
            e[axis] = 1.0
            v = F.mat_mul_vec(matrix, e)
            v = [v[0] * dx, v[1] * dx, v[2] * dx]
            pos, _, _ = self._probe([start_pos[i] + v[i] for i in range(3)], [direction_pos[i] + v[i] for i in range(3)])
            if pos is None:
                return None
            d = [pos[i] - actual[i] for i in range(3)]
            r = F._rcp(dx)                                   # glam DVec4 / f64 = * (1/dx)
            cols.append([d[0] * r, d[1] * r, d[2] * r, 0.0 * r])
        new_mat = cols + [[0.0, 0.0, 0.0, 1.0]]
        p = F.mat_mul_vec(F.mat_mul(new_mat, F.mat_inverse(matrix)), [direction_pos[0], direction_pos[1], direction_pos[2], 1.0])
        return cols + [[actual[0] - p[0], actual[1] - p[1], actual[2] - p[2], 1.0]]

    def teleport_camera(self, prev_cam):
        cam = self.cam
        if cam.do_not_teleport_one_frame:
            cam.do_not_teleport_one_frame = False
            cam.prev_cam_pos = cam.get_cam_pos()
            return
        if not (cam.allow_teleport or cam.stop_at_objects) or self.probe is None:
            return
        cam_pos = cam.get_cam_pos()
        teleported, encounter_object, change_subspace = self._probe(cam.prev_cam_pos, cam_pos)
        if cam.stop_at_objects and encounter_object:
            self.cam = prev_cam.clone()
            return
        if teleported is not None:
            if not cam.allow_teleport:
                return
            for dx in (0.001, 0.0001, 0.00001, 0.000001):
                m = self.teleport_matrix(cam.teleport_matrix, cam.prev_cam_pos, cam_pos, teleported, dx)
                if m is None:
                    continue
                cam.teleport_matrix = m
                if change_subspace:
                    cam.in_subspace = not cam.in_subspace
                cam.prev_cam_pos = cam.get_cam_pos()
                return
            self.cam = prev_cam
        else:
            cam.prev_cam_pos = cam_pos

    def teleport_eye_matrices(self):
        """main.rs:1121-1172: each eye = the camera shifted along its x axis; if a portal lies between the camera and
        the eye, the eye's matrix is carried through it like the camera's would be."""
        cam = self.cam
        if not ((self.draw_anaglyph or self.draw_side_by_side) and cam.allow_teleport):   # main.rs:1122
            return
        ed = -self.eye_distance if self.swap_eyes else self.eye_distance

        def one(eye_x):
            start = cam.get_cam_pos()
            m = cam.get_matrix()
            dp = F.mat_mul_vec(m, [eye_x, 0.0, 0.0, 1.0])[:3]
            tr = F.mat_identity()
            tr[3] = [dp[0] - start[0], dp[1] - start[1], dp[2] - start[2], 1.0]
            matrix, sub = F.mat_mul(tr, m), cam.in_subspace
            if self.probe is None:
                return matrix, sub
            teleported, _, change_subspace = self._probe(start, dp)
            if teleported is not None:
                for dx in (0.001, 0.0001, 0.00001, 0.000001):
                    mm = self.teleport_matrix(matrix, start, dp, teleported, dx)
                    if mm is None:
                        continue
                    matrix = mm
                    if change_subspace:
                        sub = not cam.in_subspace
                    break
            return matrix, sub

        self.left_eye_matrix, self.left_eye_in_subspace = one(-ed)
        self.right_eye_matrix, self.right_eye_in_subspace = one(ed)

    def update(self, time: float):
        mem = self.memory
        self.anim.update(mem, time)
        self.scene.camera_matrix = self.cam.get_matrix()            # send_camera_object_matrix (default true)
        current = mem.get("CurrentCam")
        cam = self.cam
        if cam.from_cam != current:
            if current is not None:
                if cam.from_cam is None:
                    mem["OriginalCam"] = cam.get_calculated_cam()
                cc = self.anim.cam_get(current)
                if cc is None:
                    raise ValueError("camera looks at a missing matrix")
            else:
                cc = mem["OriginalCam"]
            cam.from_cam = current
            cam.alpha, cam.beta, cam.r = cc["alpha"], cc["beta"], cc["r"]
            cam.look_at = list(cc["look_at"])
            cam.teleport_matrix = cc["matrix"]
            cam.in_subspace = cc["in_subspace"]
            cam.free_movement = cc["free_movement"]
            if cam.free_movement:
                pv = cam.get_pos_vec()
                cam.look_at = [pv[i] + cam.look_at[i] for i in range(3)]
            cam.do_not_teleport_one_frame = True
        elif cam.from_cam is not None:
            cc = self.anim.cam_get(cam.from_cam)
            if cc is None:
                raise ValueError("camera looks at a missing matrix")
            if not cam.free_movement:
                cam.look_at = list(cc["look_at"])
        oc = mem.pop("OverrideCam", None)
        if oc is not None:
            cam.alpha, cam.beta, cam.r = oc["alpha"], oc["beta"], oc["r"]
            cam.look_at = list(oc["look_at"])
            cam.free_movement = oc["free_movement"]
            if oc["override_matrix"]:
                cam.teleport_matrix = oc["matrix"]
                cam.in_subspace = oc["in_subspace"]
                cam.do_not_teleport_one_frame = True
        if self.cam.get_matrix() != self.prev_cam.get_matrix():
            self.teleport_camera(self.prev_cam.clone())
        self.teleport_eye_matrices()
        self.prev_cam = self.cam.clone()
        self.scene.camera_matrix = self.cam.get_matrix()

    # -- what set_uniforms (main.rs:1266-1359) then sends for the camera
    def camera_state(self) -> dict:
        m = self.cam.get_matrix()
        return {"camera": [x for col in m for x in col],
                "camera_mul_inv": [x for col in F.mat_inverse(self.cam.teleport_matrix) for x in col],
                "left_eye": [x for col in self.left_eye_matrix for x in col],
                "right_eye": [x for col in self.right_eye_matrix for x in col],
                "left_eye_in_subspace": bool(self.left_eye_in_subspace),
                "right_eye_in_subspace": bool(self.right_eye_in_subspace),
                "in_subspace": bool(self.cam.in_subspace), "scale": F.camera_scale(m),
                "look_at": list(self.cam.look_at), "alpha": self.cam.alpha, "beta": self.cam.beta, "r": self.cam.r,
                "time": self.scene.time, "total_time": self.scene.total_time}


def frame_times(frame_index: int, frame_count: int, motion_blur_frames: int, duration_seconds: float, exposure: float = 0.5):
    """render_animation's sub-frame times (main.rs:1786-1797): [(aa_start, seconds)]."""
    out = []
    for j in range(motion_blur_frames):
        t = (frame_index / frame_count) + j / motion_blur_frames / frame_count * exposure
        out.append((j, t * duration_seconds))
    return out
