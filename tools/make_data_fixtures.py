#!/usr/bin/env python
"""Convert the reference checkout's input DATA (meshes, the bunny cloud, the TX90L
kinematics and the initial/solved paths) into the compact fixtures this repo ships.

`/root/reference` does not exist on the GPU box, so every workload input named by
BASELINE.json's configs has to travel with the repo.  Only data is converted -- no
reference source code is read or copied.  Run once, in the build container:

    python tools/make_data_fixtures.py [/root/reference]

Outputs (committed):
    semiinfiniteoptimization_b200/data/geometry.npz   cube/m797/m1062/sphere (verts f64, tris i32), bunny (pts f64)
    semiinfiniteoptimization_b200/data/arc_rob.json   ARC.rob kinematics block (TParent, parents, axis, limits, mount)
    semiinfiniteoptimization_b200/data/paths.json     resources/*.path, output/*.path, *.configs
    semiinfiniteoptimization_b200/data/worlds.json    the rigidObject placement of tx90_geom_test{2,3}.xml
"""
import json
import os
import sys
import xml.etree.ElementTree as ET

import numpy as np

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "semiinfiniteoptimization_b200", "data")


def read_off(fn):
    toks = open(fn).read().split()
    assert toks[0] == "OFF", fn
    nv, nf = int(toks[1]), int(toks[2])
    k = 4
    v = np.array(toks[k:k + 3 * nv], dtype=np.float64).reshape(nv, 3)
    k += 3 * nv
    tris = []
    for _ in range(nf):
        c = int(toks[k])
        idx = [int(t) for t in toks[k + 1:k + 1 + c]]
        k += 1 + c
        for j in range(1, c - 1):          # fan-triangulate polygons
            tris.append((idx[0], idx[j], idx[j + 1]))
    return v, np.array(tris, dtype=np.int32)


def read_pcd_ascii(fn):
    lines = open(fn).read().splitlines()
    for i, l in enumerate(lines):
        if l.startswith("DATA"):
            assert "ascii" in l
            body = lines[i + 1:]
            break
    pts = np.array([[float(t) for t in l.split()[:3]] for l in body if l.strip()], dtype=np.float64)
    return pts


def read_rob(fn):
    # join continuation lines, strip comments
    txt = open(fn).read().replace("\\\n", " ")
    rob = {}
    KEYS = ("TParent", "parents", "axis", "translation", "qMinDeg", "qMaxDeg", "velMaxDeg", "accMaxDeg",
            "geometry", "mount", "mass", "automass", "torqueMax", "servoP", "servoI", "servoD",
            "dryFriction", "viscousFriction", "noselfcollision")
    lines = []
    for line in txt.splitlines():
        line = line.split("#")[0].strip()
        if not line:
            continue
        # ARC.rob ends the TParent block with a stray continuation, which glues the next
        # keyword onto it: split a line again wherever a keyword token appears mid-line.
        toks = line.split()
        cur = [toks[0]]
        for t in toks[1:]:
            if t in KEYS:
                lines.append(" ".join(cur))
                cur = [t]
            else:
                cur.append(t)
        lines.append(" ".join(cur))
    for line in lines:
        key, _, rest = line.partition(" ")
        rest = rest.strip()
        if key in ("TParent", "axis", "qMinDeg", "qMaxDeg", "velMaxDeg", "accMaxDeg", "translation"):
            rob[key] = [float(t) for t in rest.split()]
        elif key == "parents":
            rob[key] = [int(t) for t in rest.split()]
        elif key == "mount":
            toks = rest.split()
            rob["mount"] = {"link": int(toks[0]), "file": toks[1].strip('"'),
                            "T": [float(t) for t in toks[2:14]]}
        elif key == "geometry":
            rob["geometry"] = [t.strip('"') for t in rest.split()]
    return rob


def read_path(fn):
    times, ms = [], []
    for l in open(fn).read().splitlines():
        t = l.split()
        if not t:
            continue
        n = int(t[1])
        times.append(float(t[0]))
        ms.append([float(x) for x in t[2:2 + n]])
    return {"times": times, "milestones": ms}


def read_configs(fn):
    out = []
    for l in open(fn).read().splitlines():
        t = l.split()
        if t:
            out.append([float(x) for x in t[1:1 + int(t[0])]])
    return out


def read_world(fn):
    root = ET.parse(fn).getroot()
    w = {"robot": root.find("robot").attrib["file"], "rigidObjects": []}
    for ro in root.findall("rigidObject"):
        g = ro.find("geometry")
        w["rigidObjects"].append({
            "name": ro.attrib.get("name"),
            "position": [float(t) for t in ro.attrib.get("position", "0 0 0").split()],
            "mesh": g.attrib["mesh"],
            "rotateX": float(g.attrib.get("rotateX", 0)),
            "scale": float(g.attrib.get("scale", 1)),
        })
    return w


def main():
    os.makedirs(OUT, exist_ok=True)
    arrs = {}
    for name in ("cube", "m797", "m1062", "sphere"):
        v, t = read_off(os.path.join(REF, "data", name + ".off"))
        arrs[name + "_verts"] = v
        arrs[name + "_tris"] = t
        print(name, v.shape, t.shape, v.min(0), v.max(0))
    arrs["bunny_pts"] = read_pcd_ascii(os.path.join(REF, "data", "bunny.pcd"))
    print("bunny", arrs["bunny_pts"].shape)
    np.savez_compressed(os.path.join(OUT, "geometry.npz"), **arrs)
    json.dump(read_rob(os.path.join(REF, "data", "ARC.rob")), open(os.path.join(OUT, "arc_rob.json"), "w"), indent=1)
    paths = {}
    for d, fns in (("resources", ("robottrajopt_initial.path", "robottrajopt_chair_initial.path",
                                   "robottrajopt_tree_initial.path")),
                   ("output", ("robottrajopt_solved.path", "robottrajopt_tree_solved.path"))):
        for fn in fns:
            paths[d + "/" + fn] = read_path(os.path.join(REF, d, fn))
    paths["resources/robotplanopt_chair_start_goal.configs"] = read_configs(
        os.path.join(REF, "resources", "robotplanopt_chair_start_goal.configs"))
    json.dump(paths, open(os.path.join(OUT, "paths.json"), "w"))
    worlds = {fn: read_world(os.path.join(REF, "data", fn)) for fn in ("tx90_geom_test2.xml", "tx90_geom_test3.xml")}
    json.dump(worlds, open(os.path.join(OUT, "worlds.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
