"""Reader for assets/model_card.bin (written by tools/compile_model.py).

The blob replaces everything the reference recompiles from XML/STL on every reset
(kinova_gripper_env.py:851-881, 1002-1003): hulls, hull graphs, ray BVHs, inertial constants and
the with_noise pose tables (kinova_gripper_env.py:1008-1054, 1245)."""
import os
import struct

import numpy as np

SHAPES = ["CubeS", "CubeB", "CylinderS", "CylinderB", "Cube45S", "Cube45B", "Cone1S", "Cone1B",
          "Cone2S", "Cone2B", "Vase1S", "Vase1B", "Vase2S", "Vase2B",
          # the medium sizes (ENV:150-178; MAIN:1271 test_sizes = ["M"]) are appended so that the ids of the 14 above never change
          "CubeM", "CylinderM", "Cube45M", "Cone1M", "Cone2M", "Vase1M", "Vase2M",
          "LemonS", "LemonM", "LemonB"]          # ENV:198-201: the single-geom object of the later additions
ORIENTATIONS = ["normal", "rotated", "top"]
DEFAULT_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "assets", "model_card.bin")
_DT = {0: "<f8", 1: "<i4", 2: "<f4"}


class ModelCard:
    def __init__(self, path=DEFAULT_PATH):
        self.path = path
        with open(path, "rb") as f:
            self.raw = f.read()
        magic, n, _ = struct.unpack_from("<8sII", self.raw, 0)
        if magic != b"KGMC0001":
            raise ValueError(f"{path}: not a model card")
        self.sections = {}
        for i in range(n):
            name, dt, cnt, off = struct.unpack_from("<24sIIQ", self.raw, 16 + 40 * i)
            self.sections[name.rstrip(b"\0").decode()] = (dt, cnt, off)

    def get(self, name):
        dt, cnt, off = self.sections[name]
        return np.frombuffer(self.raw, dtype=_DT[dt], count=cnt, offset=off)

    def pose(self, mode="train"):
        """[shape, orientation, row, 9]: obj xyz, raw hand euler xyz, euler after the 5-char truncation.
        Row r here is file line r; the reference discards line 0 (ENV:1012), i.e. its data[i] is row i+1."""
        a = self.get("pose_train" if mode != "test" else "pose_test")
        return a.reshape(len(SHAPES), len(ORIENTATIONS), -1, 9)
