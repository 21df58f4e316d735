"""Fixture for packed tangent frames (mesh.cpp:355,634-665,2339-2351,2417-2429; interaction.h:571-597), written by the
UNMODIFIED reference (oracle/_ref, variant scalar_rgb):

    python tests/golden/gen_golden_tangent.py        (in the environment of oracle.ref_env.reference_env())

A PLY-loaded UV sphere carrying an ANISOTROPIC BSDF: the reference's loader then packs frame_encode(normal, tangent) into the
frame slot of the vertex records. Stored: the packed records + faces with the FaceUVFlipped bit as the live-Mitsuba extractor
(mitsuba3_b200/mitsuba_plugin.py) reads them off the mesh, 256 surface-interaction records (shading frame s, t, n and uv) of
rays shot at the sphere, and scalar_rgb renders of the env_scene setup with two anisotropic materials.
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE); sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import mitsuba as mi  # noqa: E402

mi.set_variant("scalar_rgb")
import gen_golden as gg  # noqa: E402
from mitsuba3_b200 import mitsuba_plugin as plug  # noqa: E402

MATERIALS = {
    "aniso_principled": {"type": "principled", "base_color": {"type": "rgb", "value": [0.9, 0.6, 0.2]}, "roughness": 0.35,
                         "anisotropic": 0.8, "metallic": 0.9, "specular": 0.5},
    "aniso_roughconductor": {"type": "roughconductor", "distribution": "ggx", "alpha_u": 0.05, "alpha_v": 0.3,
                             "eta": {"type": "rgb", "value": [0.2, 0.9, 1.1]}, "k": {"type": "rgb", "value": [3.9, 2.4, 2.2]}},
}
RENDERS = [("aniso_principled", 16, 5, 0, ""), ("aniso_principled", 8, 3, 4, "_m"), ("aniso_roughconductor", 16, 5, 1, "_m"), ("aniso_roughconductor", 8, 4, 2, "")]


def sphere_ply(mirror_u=False):
    n_theta, n_phi, radius, center = 12, 24, 0.6, np.array([0.0, 0.6, 0.0])
    th = (np.arange(n_theta + 1, dtype=np.float64) / n_theta) * np.pi
    ph = (np.arange(n_phi + 1, dtype=np.float64) / n_phi) * 2 * np.pi
    T_, P_ = np.meshgrid(th, ph, indexing="ij")
    nrm = np.stack([np.sin(T_) * np.cos(P_), np.cos(T_), np.sin(T_) * np.sin(P_)], -1).reshape(-1, 3)
    pos = (nrm * radius + center).astype(np.float32); nrm = nrm.astype(np.float32)
    uv = np.stack([P_ / (2 * np.pi), T_ / np.pi], -1).reshape(-1, 2).astype(np.float32)
    if mirror_u:
        uv[:, 0] = 1.0 - uv[:, 0]          # negative uv determinant on every face: FaceUVFlipped set, bitangent negated
    idx = lambda i, j: i * (n_phi + 1) + j
    faces = []
    for i in range(n_theta):
        j = np.arange(n_phi)
        a, b, c, d_ = idx(i, j), idx(i + 1, j), idx(i + 1, j + 1), idx(i, j + 1)
        if i > 0: faces.append(np.stack([a, d_, c], -1))
        if i < n_theta - 1: faces.append(np.stack([a, c, b], -1))
    faces = np.concatenate(faces, 0).astype(np.uint32)
    ply = os.path.join(tempfile.gettempdir(), "b200pt_tangent_sphere%s.ply" % ("_m" if mirror_u else ""))
    gg.write_ply(ply, pos, nrm, uv, faces)
    return ply


def scene_dict(ply, mat, spp, md):
    d = gg.env_scene(mi.ScalarTransform4f, gg.env_image(), mi.Bitmap, spp=spp, max_depth=md, area_light=True)
    del d["cube-a"], d["cube-b"]
    d["ball-mat"] = MATERIALS[mat]
    d["ball"] = {"type": "ply", "filename": ply, "bsdf": {"type": "ref", "id": "ball-mat"}}
    return gg.block_one(d, 32)


def main():
    out = {}
    rs = np.random.RandomState(11)
    n = 256
    dirs = rs.normal(size=(n, 3)); dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    o = (np.array([0.0, 0.6, 0.0]) + 2.0 * dirs).astype(np.float32)
    tgt = np.array([0.0, 0.6, 0.0]) + 0.35 * rs.uniform(-1, 1, (n, 3))
    d = (tgt - o); d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    out["si_rays"] = np.concatenate([o, d], 1)
    plys = {}
    for tag, mirror in (("", False), ("_m", True)):
        ply = plys[tag] = sphere_ply(mirror)
        sc = mi.load_dict(scene_dict(ply, "aniso_principled", 16, 5))
        ball = [s for s in sc.shapes() if s.id() == "ball"][0]
        assert ball.packs_tangent()
        host = plug.extract_scene(mi, sc)
        hb = [s for s in host.shapes if s.id == "ball"][0]
        assert hb.layout & plug.abi.LAYOUT_TANGENTS
        out["packed_vertices" + tag], out["faces" + tag], out["layout"] = hb.vertices, hb.faces, np.uint32(hb.layout)
        out["emitter_order"] = np.array(["envmap" if e.type == plug.abi.EMITTER_ENVMAP else "area" for e in host.emitters])    # Scene::emitters() of this build
        print("ball%s:" % tag, hb.vertices.shape, hb.faces.shape, "uv-flipped faces:", int((hb.faces[:, 3] >> 31).sum()))
        # surface interactions: rays from a shell around the sphere towards points inside it
        rec = np.zeros((n, 13), np.float32)
        only = mi.load_dict({"type": "scene", "ball": {"type": "ply", "filename": ply, "bsdf": MATERIALS["aniso_principled"]}})
        for i in range(n):
            si = only.ray_intersect(mi.Ray3f(mi.Point3f(*o[i]), mi.Vector3f(*d[i])))
            assert si.is_valid()
            rec[i] = [si.t, *si.sh_frame.s, *si.sh_frame.t, *si.sh_frame.n, *si.uv, float(si.prim_index)]
        out["si_records" + tag] = rec
    for mat, spp, md, seed, tag in RENDERS:
        img = np.array(mi.render(mi.load_dict(scene_dict(plys[tag], mat, spp, md)), seed=seed, spp=spp))
        out[f"{mat}{tag}_32_spp{spp}_d{md}_seed{seed}"] = img
        print(mat, tag, spp, md, seed, img.mean())
    np.savez_compressed(os.path.join(HERE, "tangent_mesh.npz"), **out)
    print("wrote tangent_mesh.npz", os.path.getsize(os.path.join(HERE, "tangent_mesh.npz")))


if __name__ == "__main__":
    import traceback
    try:
        main(); sys.stdout.flush(); os._exit(0)
    except BaseException:
        traceback.print_exc(file=sys.stdout); sys.stdout.flush(); os._exit(1)
