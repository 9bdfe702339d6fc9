#!/usr/bin/env python
"""Offline scene-asset compiler.

Reads the reference's URDF files and collision meshes (reference
``safemotions/description/{urdf,meshes}``) and writes one
``safemotionsrisk_b200/assets/scene_assets.npz`` that travels with the repo (the GPU box has no
``/root/reference``).  Only *derived data* is stored: kinematic trees (joint origins, axes, limits) and the
de-duplicated convex-hull vertex sets of every collision part, already scaled and moved into the frame of the
URDF link that carries them.  No reference source code is copied.

Bullet semantics restated here (not citable inside /root/reference):
  * a ``.stl`` collision mesh -> one convex hull of all its vertices;
  * every ``o``/``g`` group of an ``.obj`` collision mesh -> one convex hull (compound of hulls);
  * ``<cylinder>`` -> convex hull of 2 x 32 rim vertices (no implicit cylinder flag is passed by the reference,
    robot_scene_base.py:305-316);
  * ``<box>`` -> box (kept as its 8 corner vertices; the 1 mm margin shrink is applied by the scene builder);
  * mesh ``scale`` multiplies vertices; the collision ``<origin>`` is applied afterwards.

Usage: python tools/compile_assets.py [/root/reference]
"""
import os
import struct
import sys
import xml.etree.ElementTree as ET

import numpy as np

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
DESC = os.path.join(REF, "safemotions", "description")
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                   "safemotionsrisk_b200", "assets", "scene_assets.npz")


def rpy_to_matrix(rpy):
    r, p, y = rpy
    cr, sr, cp, sp, cy, sy = np.cos(r), np.sin(r), np.cos(p), np.sin(p), np.cos(y), np.sin(y)
    rx = np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]])
    ry = np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]])
    rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]])
    return rz @ ry @ rx  # URDF fixed-axis roll, pitch, yaw


def parse_origin(elem):
    xyz = np.zeros(3)
    rpy = np.zeros(3)
    if elem is not None:
        o = elem.find("origin")
        if o is not None:
            if o.get("xyz"):
                xyz = np.array([float(x) for x in o.get("xyz").split()])
            if o.get("rpy"):
                rpy = np.array([float(x) for x in o.get("rpy").split()])
    return xyz, rpy


def resolve_mesh(filename):
    prefix = "package://safemotions/description/"
    assert filename.startswith(prefix), filename
    return os.path.join(DESC, filename[len(prefix):])


def load_stl_vertices(path):
    with open(path, "rb") as f:
        data = f.read()
    n = struct.unpack_from("<I", data, 80)[0]
    assert len(data) == 84 + 50 * n, "only binary STL is expected"
    verts = np.zeros((n * 3, 3))
    for i in range(n):
        vals = struct.unpack_from("<12f", data, 84 + 50 * i)
        verts[3 * i:3 * i + 3] = np.array(vals[3:12]).reshape(3, 3)
    return [verts]


def load_obj_groups(path):
    verts, groups, cur = [], [], None
    with open(path) as f:
        for line in f:
            t = line.split()
            if not t:
                continue
            if t[0] == "v":
                verts.append([float(x) for x in t[1:4]])
            elif t[0] in ("o", "g"):
                cur = []
                groups.append(cur)
            elif t[0] == "f":
                if cur is None:
                    cur = []
                    groups.append(cur)
                cur.extend(int(x.split("/")[0]) - 1 for x in t[1:])
    verts = np.array(verts)
    return [verts[sorted(set(g))] for g in groups if g]


def dedupe(v):
    # exact duplicates only: the support function of the hull is unchanged
    _, idx = np.unique(np.round(v, 12), axis=0, return_index=True)
    return v[np.sort(idx)]


def collision_parts(link_elem):
    """All convex parts of a link, vertices in the link frame. Returns list of (kind, verts)."""
    parts = []
    for col in link_elem.findall("collision"):
        xyz, rpy = parse_origin(col)
        rot = rpy_to_matrix(rpy)
        geom = col.find("geometry")
        mesh, box, cyl, sph = geom.find("mesh"), geom.find("box"), geom.find("cylinder"), geom.find("sphere")
        if mesh is not None:
            scale = np.ones(3)
            if mesh.get("scale"):
                scale = np.array([float(x) for x in mesh.get("scale").split()])
            path = resolve_mesh(mesh.get("filename"))
            raw = load_stl_vertices(path) if path.lower().endswith(".stl") else load_obj_groups(path)
            for v in raw:
                v = dedupe(v) * scale
                parts.append(("hull", v @ rot.T + xyz))
        elif box is not None:
            half = 0.5 * np.array([float(x) for x in box.get("size").split()])
            corners = np.array([[sx, sy, sz] for sx in (-1, 1) for sy in (-1, 1) for sz in (-1, 1)]) * half
            parts.append(("box", corners @ rot.T + xyz))
        elif cyl is not None:
            radius, length = float(cyl.get("radius")), float(cyl.get("length"))
            v = []
            for i in range(32):
                ang = 2 * np.pi * (i / 32.0)
                for z in (0.5 * length, -0.5 * length):
                    v.append([radius * np.sin(ang), radius * np.cos(ang), z])
            parts.append(("hull", np.array(v) @ rot.T + xyz))
        elif sph is not None:
            # the human URDF's 1e-5 m dummy spheres: kept as a point (never in any pair list)
            parts.append(("point", np.zeros((1, 3)) + xyz))
    return parts


def compile_urdf(path, prefix, out):
    root = ET.parse(path).getroot()
    links = {l.get("name"): l for l in root.findall("link")}
    joints = root.findall("joint")
    children = {}
    child_names = set()
    for j in joints:
        children.setdefault(j.find("parent").get("link"), []).append(j)
        child_names.add(j.find("child").get("link"))
    base = [n for n in links if n not in child_names]
    assert len(base) == 1
    names, parents, jtypes, axes, origins_xyz, origins_rpy, limits = [base[0]], [-1], [0], [np.zeros(3)], \
        [np.zeros(3)], [np.zeros(3)], [np.zeros(4)]

    def inertial_of(name):
        # Bullet places a link's collision object at its inertial frame; the manifold contact-breaking
        # threshold depends on the shape's extent around that frame
        return parse_origin(links[name].find("inertial"))

    def visit(link_name, idx):
        for j in children.get(link_name, []):  # pre-order DFS = Bullet link numbering
            child = j.find("child").get("link")
            xyz, rpy = parse_origin(j)
            jt = {"fixed": 0, "revolute": 1, "continuous": 2, "prismatic": 3}[j.get("type")]
            ax = np.array([float(x) for x in j.find("axis").get("xyz").split()]) if j.find("axis") is not None \
                else np.array([1.0, 0, 0])
            lim = np.zeros(4)
            if j.find("limit") is not None:
                le = j.find("limit")
                lim = np.array([float(le.get("lower", 0)), float(le.get("upper", 0)),
                                float(le.get("effort", 0)), float(le.get("velocity", 0))])
            names.append(child)
            parents.append(idx)
            jtypes.append(jt)
            axes.append(ax)
            origins_xyz.append(xyz)
            origins_rpy.append(rpy)
            limits.append(lim)
            visit(child, len(names) - 1)

    visit(base[0], 0)
    out[prefix + "/link_names"] = np.array(names)
    out[prefix + "/parent"] = np.array(parents, dtype=np.int32)
    out[prefix + "/joint_type"] = np.array(jtypes, dtype=np.int32)
    out[prefix + "/joint_axis"] = np.array(axes)
    out[prefix + "/joint_xyz"] = np.array(origins_xyz)
    out[prefix + "/joint_rpy"] = np.array(origins_rpy)
    out[prefix + "/joint_limit"] = np.array(limits)  # lower, upper, effort, velocity
    out[prefix + "/inertial_xyz"] = np.array([inertial_of(n)[0] for n in names])
    out[prefix + "/inertial_rpy"] = np.array([inertial_of(n)[1] for n in names])
    part_link, part_kind, part_start, verts = [], [], [0], []
    for li, name in enumerate(names):
        for kind, v in collision_parts(links[name]):
            part_link.append(li)
            part_kind.append(kind)
            verts.append(v)
            part_start.append(part_start[-1] + len(v))
    out[prefix + "/part_link"] = np.array(part_link, dtype=np.int32)
    out[prefix + "/part_kind"] = np.array(part_kind)
    out[prefix + "/part_start"] = np.array(part_start, dtype=np.int32)
    out[prefix + "/part_verts"] = np.concatenate(verts) if verts else np.zeros((0, 3))


def main():
    out = {}
    urdf = os.path.join(DESC, "urdf")
    compile_urdf(os.path.join(urdf, "robot.urdf"), "robot", out)
    compile_urdf(os.path.join(urdf, "robot_ball_machine.urdf"), "robot_ball_machine", out)
    compile_urdf(os.path.join(urdf, "human.urdf"), "human", out)
    for name in ("table", "ISS", "asteroid", "basketball_red"):
        compile_urdf(os.path.join(urdf, "obstacles", name + ".urdf"), "obstacle_" + name, out)
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    np.savez_compressed(OUT, **out)
    total = sum(v.nbytes for v in out.values())
    print("wrote", OUT, "arrays:", len(out), "bytes(raw):", total)
    for k in sorted(out):
        if k.endswith("part_start"):
            print(k, np.diff(out[k]).tolist())


if __name__ == "__main__":
    main()
