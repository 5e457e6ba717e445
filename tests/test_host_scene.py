"""Host-side scene API and flattener: value semantics, transforms, CSG composites, compound push rules,
materials (objects.c:1463-1716, compound.c:140-207, container.c:376-421) and the flat layout."""
import ctypes as C
import math

import numpy as np
import pytest

import actinon_b200 as acn
from actinon_b200 import api
from actinon_b200.api import rotx, rotz


def nodes_of(flat):
    fs = flat.struct
    return [fs.nodes[i] for i in range(fs.n_nodes)], [fs.children[i] for i in range(fs.n_children)], [fs.materials[i] for i in range(fs.n_materials)]


def test_primitives_flatten_matches_appendix_c():
    flat = acn.scenes.primitives().flatten()
    nodes, children, mats = nodes_of(flat)
    fs = flat.struct
    assert fs.n_nodes == 11 and fs.light_root == 0 and fs.matter_root == 2
    light = nodes[children[nodes[0].child0]]
    assert light.kind == api.KIND_SPHERE and list(light.pos) == [0, -4, 4] and light.tail[0] == 0.5
    lm = mats[light.material]
    assert lm.radiance == 30 and list(lm.color) == [0.7, 0.7, 0.7]
    matter = [nodes[children[nodes[2].child0 + i]] for i in range(nodes[2].child1)]
    assert [n.kind for n in matter] == [1, 2, 3, 5, 3, 3, 3, 3]
    floor = matter[0]
    fm = mats[floor.material]
    assert list(floor.pos) == [0, 0, -1] and list(floor.rax)[6:] == [0, 0, 1]
    assert (fm.refractive_index, fm.fresnel_reflectivity, fm.diffuse_reflectivity, fm.sigma) == (1.2, 1.0, 1.0, 0.29)
    xs = [round(n.pos[0], 12) for n in matter[1:]]
    assert xs == [-2.0, -2.0, -0.9, -0.9, 0.1, 1.0, 2.0]
    tor = matter[3]
    assert tor.has_envelope and abs(tor.env_radius - 0.505) < 1e-12 and list(tor.env_pos) == list(tor.pos)
    assert abs(tor.tail[0] - 1 / 0.35) < 1e-12 and abs(tor.tail[1] - 0.15 / 0.35) < 1e-12 and tor.tail[2] == 200
    assert np.allclose(np.array(tor.rax).reshape(3, 3), [[1, 0, 0], [0, 0, 1], [0, -1, 0]], atol=1e-15)   # rows rotated by rot_x(90): torus axis along world y
    assert np.allclose(list(matter[2].tail), [1 / 0.09, 1 / 0.09, 4.0, -1])          # el1
    assert np.allclose(list(matter[5].tail), [25, 25, -1, -1])                        # hyperboloid1
    assert np.allclose(list(matter[6].tail), [25, 25, -0.25, 0])                      # cone
    assert np.allclose(list(matter[7].tail), [6.25, 6.25, 0, -1])                     # cylinder
    # 7 objects share one material, floor and light have their own
    assert fs.n_materials == 3


def test_value_semantics_and_transforms():
    sc = acn.Scene()
    s = sc.create_sphere(1.0)
    t = s + (1, 2, 3)            # moved clone, s untouched
    u = t * 2.0                  # scaled clone: pos and radius
    v = u * rotz(90)
    sc.push(s); sc.push(t); sc.push(u); sc.push(v)
    nodes, children, _ = nodes_of(sc.flatten())
    m = [nodes[children[nodes[1].child0 + i]] for i in range(4)]
    assert list(m[0].pos) == [0, 0, 0] and m[0].tail[0] == 1
    assert list(m[1].pos) == [1, 2, 3]
    assert list(m[2].pos) == [2, 4, 6] and m[2].tail[0] == 2
    assert np.allclose(list(m[3].pos), [-4, 2, 6]) and np.allclose(np.array(m[3].rax).reshape(3, 3)[0], [0, 1, 0])


def test_squaroid_and_distance_scaling():
    sc = acn.Scene()
    e = sc.create_ellipsoid(1, 2, 4) * 0.5          # r *= f^2 (objects.c:831)
    t = sc.create_torus(2.0, 0.5) * 3.0             # inv_scale *= 1/f, envelope scaled
    sc.push(e); sc.push(t)
    nodes, children, _ = nodes_of(sc.flatten())
    ne, nt = nodes[children[nodes[1].child0]], nodes[children[nodes[1].child0 + 1]]
    assert np.allclose(list(ne.tail), [1, 0.25, 1 / 16, -0.25])
    assert abs(nt.tail[0] - 1 / 6) < 1e-15 and abs(nt.env_radius - 2.5 * 1.01 * 3) < 1e-12


def test_csg_pairs_copy_properties_and_envelope_rules():
    sc = acn.Scene()
    a = sc.create_sphere(1.0).set_material("glass").set_envelope((0, 0, 0), 1.5)
    b = sc.create_plane().set_material("gold")
    inside, outside, neg = a & b, a | b, ~a
    sc.push(inside); sc.push(outside); sc.push(neg)
    nodes, children, mats = nodes_of(sc.flatten())
    top = [nodes[children[nodes[1].child0 + i]] for i in range(3)]
    assert [n.kind for n in top] == [api.KIND_PAIR_INSIDE, api.KIND_PAIR_OUTSIDE, api.KIND_NEG]
    # prp copied from o1 (objects.c:1014,1164,1318); pair_outside drops the envelope (objects.c:1169-1173)
    assert all(mats[n.material].refractive_index == 1.46 for n in top)
    assert top[0].has_envelope == 1 and top[1].has_envelope == 0 and top[2].has_envelope == 1
    assert nodes[top[0].child0].kind == api.KIND_SPHERE and nodes[top[0].child1].kind == api.KIND_PLANE
    # moving a pair moves its children
    inside.move((0, 0, 5)); sc.clear(); sc.push(inside)
    nodes, children, _ = nodes_of(sc.flatten())
    p = nodes[children[nodes[1].child0]]
    assert list(p.pos) == [0, 0, 5] and list(nodes[p.child0].pos) == [0, 0, 5] and list(nodes[p.child1].pos) == [0, 0, 5]
    assert list(p.env_pos) == [0, 0, 5]


def test_scale_object():
    sc = acn.Scene()
    s = sc.create_sphere(0.5).set_envelope((1, 1, 1), 2.0)
    o = s.scaled_by_vec((1.0, 0.5, 4.0))
    sc.push(o)
    nodes, children, _ = nodes_of(sc.flatten())
    n = nodes[children[nodes[1].child0]]
    assert n.kind == api.KIND_SCALE and np.allclose(list(n.tail)[:3], [1, 2, 0.25])
    assert np.allclose(list(n.env_pos), [1, 0.5, 4]) and n.env_radius == 8.0       # objects.c:1396-1400


def test_balanced_composite_and_compound_rules():
    sc = acn.Scene()
    cover = sc.create_plane()
    lst = sc.create_list([cover + (0, 0, i) for i in range(5)])
    comp = lst.create_inside_composite()
    sc.push(comp)
    nodes, children, _ = nodes_of(sc.flatten())

    def shape(i):
        n = nodes[i]
        return n.pos[2] if n.kind == api.KIND_PLANE else (shape(n.child0), shape(n.child1))
    # size 5 -> (2, 3) -> ((1,1),(1,(1,1)))   container.c:376-392
    assert shape(children[nodes[1].child0]) == ((0.0, 1.0), (2.0, (3.0, 4.0)))

    # compound without envelope dissolves into its parent, with envelope it stays a node (compound.c:166-182)
    sc2 = acn.Scene()
    sph = sc2.create_sphere(0.1)
    c1 = sc2.create_list([sph, sph + (1, 0, 0)]).create_compound()
    sc2.push(c1)
    nodes, children, _ = nodes_of(sc2.flatten())
    assert nodes[1].child1 == 2 and all(nodes[children[nodes[1].child0 + i]].kind == api.KIND_SPHERE for i in range(2))
    c1.set_auto_envelope()
    sc2.clear(); sc2.push(c1)
    nodes, children, _ = nodes_of(sc2.flatten())
    assert nodes[1].child1 == 1
    cn = nodes[children[nodes[1].child0]]
    assert cn.kind == api.KIND_COMPOUND and cn.has_envelope and cn.child1 == 2
    # the Monte-Carlo estimate (objects.c:312-363) must contain both spheres
    for i in range(2):
        s = nodes[children[cn.child0 + i]]
        assert s.has_envelope
        d = math.dist(list(s.env_pos), list(s.pos))
        assert d + 0.1 <= s.env_radius * 1.0001 and s.env_radius < 0.2
        assert math.dist(list(cn.env_pos), list(s.pos)) + 0.1 <= cn.env_radius * 1.0001


def test_scene_push_sorts_lights_from_matter_and_materials():
    sc = acn.Scene()
    lamp = sc.create_sphere(0.2).set_radiance(5)
    ball = sc.create_sphere(0.2).set_material("diffuse_polished").set_refractive_index(1.0)
    sc.push(sc.create_list([lamp, ball]))
    flat = sc.flatten()
    nodes, children, mats = nodes_of(flat)
    mr = flat.struct.matter_root
    assert nodes[0].child1 == 1 and nodes[mr].child1 == 1
    m = mats[nodes[children[nodes[mr].child0]].material]
    assert m.fresnel_reflectivity == 0.0 and m.sigma == 0.29        # set_refractive_index(1) clears Fresnel (objects.c:436-448)
    with pytest.raises(acn.AcnError):
        ball.set_material("unobtainium")


def test_pass_controller_and_image_io(tmp_path):
    sc = acn.scenes.primitives(8, 6, gradient_cycles=2)
    prm = sc.flatten().params
    img = acn.Image(8, 6)
    xy0 = img.next_pass(prm)
    assert xy0.shape == (48, 2) and np.array_equal(xy0[0], [0.5, 0.5]) and np.array_equal(xy0[9], [1.5, 1.5])
    rgb = np.zeros((48, 3), np.float32)
    rgb[8 * 2 + 3] = (1.0, 0.5, 0.25)                       # one bright pixel at (3,2)
    img.push(xy0, rgb)
    assert img.cycle == 1 and img.rval == 21943294
    xy1 = img.next_pass(prm)
    # the bright pixel and its 8 neighbours exceed the gradient threshold: 9 pixels x 2 jitter samples
    assert xy1.shape == (18, 2)
    assert np.array_equal(np.floor(xy1[:, 0]).reshape(9, 2)[:, 0], [2, 3, 4, 2, 3, 4, 2, 3, 4])
    assert np.array_equal(xy1, img.next_pass(prm))          # regenerated identically until pushed (resume semantics)
    img.push(xy1, np.full((18, 3), 0.5, np.float32))
    assert img.rval != 21943294 and img.cycle == 2
    avg = img.average()
    assert np.allclose(avg[2, 3], (1.0 + 1.0) / 3 * np.array([1, 0, 0]) + np.array([0, (0.5 + 1.0) / 3, (0.25 + 1.0) / 3]), atol=1e-6)
    # pnm: floor(c*256) clamp 255 (scene.c:76-82)
    path = str(tmp_path / "a.pnm")
    h1 = img.write_pnm(path)
    raw = open(path, "rb").read()
    assert raw.startswith(b"P6\n8 6\n255\n") and len(raw) == 11 + 8 * 6 * 3
    px = np.frombuffer(raw[11:], np.uint8).reshape(6, 8, 3)
    assert list(px[2, 3]) == [int(avg[2, 3, 0] * 256), int(avg[2, 3, 1] * 256), int(avg[2, 3, 2] * 256)]
    # checkpoint / resume
    ck = str(tmp_path / "a.lum")
    img.save(ck)
    img2 = acn.Image.load(ck)
    assert (img2.width, img2.height, img2.cycle, img2.rval) == (8, 6, 2, img.rval)
    assert img2.write_pnm(None) == h1
    assert np.array_equal(img2.next_pass(prm), img.next_pass(prm))
    img.push(img.next_pass(prm), np.zeros((len(img.next_pass(prm)), 3), np.float32))
    assert img.next_pass(prm).shape[0] == 0                 # gradient_cycles + 1 passes in total (scene.c:1103)


def test_analytic_bounding_envelopes_contain_the_monte_carlo_ones_extent():
    """a conservative bound from the shape's parameters instead of the reference's Monte-Carlo estimate
    (objects.c:312-363): exact for a ball, the largest half axis for an ellipsoid, the smaller child's bound for A&B, the
    reference's own envelope_of_pair for A|B; unbounded shapes are refused."""
    import ctypes as C
    sc = acn.Scene()

    def env_of(o):
        sc.clear(); sc.push(o)
        fs = sc.flatten().struct
        nd = fs.nodes[fs.children[fs.nodes[fs.matter_root].child0]]
        return np.array(list(nd.env_pos)), nd.env_radius, nd.has_envelope

    ball = sc.create_sphere(0.7) + (1.0, 2.0, 3.0)
    ball.clone().set_bounding_envelope()
    c, r, has = env_of(ball.clone().set_bounding_envelope())
    assert has and np.allclose(c, (1, 2, 3)) and 0.7 <= r < 0.7001
    egg = sc.create_ellipsoid(0.5, 1.5, 1.0) + (0.0, 1.0, 0.0)
    c, r, has = env_of(egg.clone().set_bounding_envelope())
    assert has and 1.5 <= r < 1.5001
    ring = sc.create_torus(2.0, 0.25)
    c, r, has = env_of(ring.clone().set_bounding_envelope())
    assert has and 2.25 <= r <= 2.25 * 1.01 + 1e-4       # create_torus sets (r1 + r2) * 1.01 itself (closures.c:568-591): kept
    lens = sc.create_sphere(1.0) & (sc.create_sphere(1.0) + (0.5, 0.0, 0.0))
    c, r, has = env_of(lens.clone().set_bounding_envelope())
    cm, rm, _ = env_of(lens.clone().set_auto_envelope())
    assert has and 1.0 <= r < 1.0001
    assert r < rm                                                   # and tighter than the estimate (1.29 around a sampled centre)
    pair = sc.create_sphere(1.0) | (sc.create_sphere(0.5) + (2.0, 0.0, 0.0))
    c, r, has = env_of(pair.clone().set_bounding_envelope())
    assert has and abs(r - 1.75) < 1e-3 and abs(c[0] - 0.75) < 1e-6     # envelope_of_pair, objects.c:113-136
    with pytest.raises(acn.AcnError) as e:
        sc.create_plane().set_bounding_envelope()
    assert e.value.code == -5
    with pytest.raises(acn.AcnError):
        (~sc.create_sphere(1.0)).set_bounding_envelope()


def test_big_csg_functions_get_packed_evaluation_programs():
    """Functions of more than 12 variables (the lamp's shade, body, chain holder ...) are cut into truth tables over stretches of
    their operator chains on the host (CsgBuilder::build_eval_program, checked there against the full program on 4096
    assignments each); no device is needed for that, and none of them may be left to the word-by-word interpreter."""
    import os, subprocess, sys, re
    ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("import sys; sys.path.insert(0, %r)\n"
            "import actinon_b200 as acn\n"
            "flat = acn.scenes.load('hanging_lamp')\n"
            "try:\n    acn.Tracer(flat, acn.Options())\nexcept acn.AcnError:\n    pass\n") % ROOT
    env = dict(os.environ, ACN_DUMP_PROGRAMS="1")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=300).stderr
    progs = re.findall(r"acn: program of node \d+: (\d+) variables, (\d+) words, (\w+)", out)
    assert len(progs) > 50
    big = [(int(v), int(w), kind) for v, w, kind in progs if int(v) > 12]
    assert big and all(kind == "packed" for _, _, kind in big)
    words = [int(x) for x in re.findall(r"packed: (\d+) evaluation words", out)]
    assert len(words) == len(big) and max(words) <= 24 and max(w for _, w, _ in big) >= 100
