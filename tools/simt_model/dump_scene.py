"""Dump the triangles of one of the bench scenes for trace_model.cpp (design tooling, not product).

    python tools/simt_model/dump_scene.py cornell /tmp/cornell.bin
    python tools/simt_model/dump_scene.py heightfield /tmp/hf.bin

File layout (little endian): uint32 n_tris, uint32 n_light_tris, float32 cam[3 origin, 3 target, fov_deg],
then per triangle 13 float32: p0 p1 p2 (world space), albedo rgb, emitter flag.
"""
import sys

import numpy as np

sys.path.insert(0, __import__("os").path.join(__import__("os").path.dirname(__file__), "..", ".."))
from mitsuba3_b200 import scene as S  # noqa: E402


def main():
    which, out = sys.argv[1], sys.argv[2]
    d = S.cornell_box() if which == "cornell" else S.cornell_box_heightfield(320)
    sc = S.load_dict(d)
    rows = []
    for sh in sc.shapes:
        b = sc.bsdfs[sh.bsdf]
        alb = np.array([0.5, 0.5, 0.5], np.float32)
        try:
            t = sc.textures[b.tex[0]]
            alb = np.asarray(t.value, np.float32).reshape(-1)[:3]
        except Exception:
            pass
        P = sh.vertices[:, :3]
        for f in sh.faces:
            rows.append(np.concatenate([P[f[0]], P[f[1]], P[f[2]], alb, [1.0 if sh.emitter >= 0 else 0.0]]))
    rows = np.asarray(rows, np.float32)
    n_light = int((rows[:, 12] > 0).sum())
    with open(out, "wb") as fh:
        fh.write(np.array([len(rows), n_light], np.uint32).tobytes())
        fh.write(np.array([0, 0, 3.9, 0, 0, 0, 39.3077], np.float32).tobytes())
        fh.write(rows.tobytes())
    print(f"{which}: {len(rows)} triangles ({n_light} emitting) -> {out}")


if __name__ == "__main__":
    main()
