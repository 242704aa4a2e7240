"""Converts the reference's Cornell Box asset into the flat scene fixture the tests and bench use
(tests/golden/cornell_emissive.npz). Run in this container only (needs /root/reference); the GPU box
uses the committed fixture. The floor's base-colour texture (BC-compressed DDS) is replaced by a
constant 0.5 grey -- see DESIGN.md."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from zetaray_b200 import scene  # noqa: E402

src = "/root/reference/Assets/CornellBox/cornell_emissive.gltf"
s = scene.load_gltf(src)
out = os.path.join(ROOT, "tests", "golden", "cornell_emissive.npz")
s.save(out)
print("wrote", out, "tris", s.num_triangles, "instances", len(s.instances), "materials", len(s.materials), "emissives", len(s.emissives))
