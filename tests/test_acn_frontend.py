"""The .acn front-end: our interpreter must turn a scene program into the same flat scene the
scene-description API builds.  The script below is written for this test (it is not one of the reference's): it uses the
language features the shipped scripts rely on — typed closure signatures, value semantics of def / push, list methods,
`for a ( in list )`, `if / else`, the `{ ... }();` idiom, operators * + - & | ! on objects, `/` as float division,
rotx in degrees, materials — and is compared node by node with the same scene built through the Python API.
When the reference tree is present (this container, not the GPU box) the committed scenes/*.npz are also regenerated
from the reference's own scripts and compared, so the flat scenes every GPU test runs on cannot drift from the scripts."""
import os

import numpy as np
import pytest

import actinon_b200 as acn
from actinon_b200 import api

SCRIPT = r"""
/* front-end test scene */
<mclosure_s></>

def scene = scene_s;
scene.image_width  = 64;
scene.image_height = 48;
scene.gamma = 0.9;
scene.trace_depth = 12;
scene.direct_samples = 7;
scene.path_samples = 3;
scene.trace_min_intensity = 0.02;
scene.max_path_length = 2.5;
def cam = vec( 1, -8, 2 );
scene.camera_position = cam;
scene.camera_view_direction = vec( 0, 0, 0.5 ) - cam;
scene.camera_top_direction = vec( 0, 0, 1 );
scene.camera_focal_length = 10 / 4;                 // '/' is float division: 2.5
scene.background_color = color( 0.2, 0.3, 0.4 );

def make_lamp = <-( num r, num radiance ) *
{
    def lamp = obj_sphere_s * r;
    lamp.set_radiance( radiance );
    lamp;
};

def cup = <-( num outer, num inner, string material ) *
{
    // a bowl: ( ball & below-the-rim ) & !( smaller ball )
    def ball  = create_sphere( outer );
    def rim   = create_plane();                     // inside = z <= 0
    def hole  = create_sphere( inner );
    def solid = ( ball & rim ) & !hole;
    solid.set_material( material );
    solid.set_auto_envelope();
    solid;
};

def row = <-( num n ) *
{
    def l = [];
    def i = 0;
    while( i < n )
    {
        def s = create_ellipsoid( 0.2, 0.2, 0.4 );
        if( ( i % 2 ) == 0 ) s.set_material( "mirror" ) else s.set_material( "diffuse" );
        l.push( s + vecx( i * 0.7 ) );              // push clones: later edits of s do not reach the list
        i = i + 1;
    }();
    l;
};

{
    scene.clear();
    scene.push( make_lamp( 0.4, 25 ) + vec( 0, -3, 5 ) );

    def floor = create_plane();
    floor.set_material( "diffuse_polished" );
    floor.set_color( color( 0.5, 0.5, 0.4 ) );
    scene.push( floor - vecz( 1 ) );

    def c = cup( 1.0, 0.9, "glass" ) * rotx( 180 );
    scene.push( c + vec( -1.5, 0, 0.5 ) );

    def things = row( 3 );
    for a ( in things ) a.set_surface_roughness( 0.01 );
    things.move( vec( 0.5, 1, -0.6 ) );
    scene.push( things );

    def ring = create_torus( 0.5, 0.1 ) * rotx( 90 );
    ring.set_material( "gold" );
    scene.push( ring + vec( 2.2, 0, 0 ) );

    scene.create_image( "frontend_test.pnm" );
}();
"""


def python_twin():
    sc = acn.Scene()
    sc.set(image_width=64, image_height=48, gamma=0.9, trace_depth=12, direct_samples=7, path_samples=3,
           trace_min_intensity=0.02, max_path_length=2.5, camera_position=(1, -8, 2),
           camera_view_direction=(-1, 8, -1.5), camera_top_direction=(0, 0, 1), camera_focal_length=2.5,
           background_color=(0.2, 0.3, 0.4))
    lamp = sc.create_sphere(1.0) * 0.4
    lamp.set_radiance(25)
    lamp.move((0, -3, 5))
    floor = sc.create_plane().set_material("diffuse_polished").set_color((0.5, 0.5, 0.4))
    floor.move((0, 0, -1))
    solid = (sc.create_sphere(1.0) & sc.create_plane()) & ~sc.create_sphere(0.9)
    solid.set_material("glass")
    solid.set_auto_envelope()
    c = solid * api.rotx(180)
    c.move((-1.5, 0, 0.5))
    things = sc.create_list()
    for i in range(3):
        s = sc.create_ellipsoid(0.2, 0.2, 0.4).set_material("mirror" if i % 2 == 0 else "diffuse")
        s.move((i * 0.7, 0, 0))
        s.set_surface_roughness(0.01)                  # the script does this afterwards, through `for a ( in things )`
        things.push(s)
    things.move((0.5, 1, -0.6))
    ring = sc.create_torus(0.5, 0.1) * api.rotx(90)
    ring.set_material("gold")
    ring.move((2.2, 0, 0))
    sc.clear(); sc.push(lamp); sc.push(floor); sc.push(c); sc.push(things); sc.push(ring)
    return sc


def flat_tables(flat):
    st = flat.struct
    nodes = [(n.kind, n.child0, n.child1, n.material, n.has_envelope, round(n.surface_roughness, 12),
              tuple(np.round(list(n.pos), 9)), tuple(np.round(list(n.rax), 9)), tuple(np.round(list(n.tail), 9)))
             for n in (st.nodes[i] for i in range(st.n_nodes))]
    mats = [(tuple(np.round(list(m.color), 9)), round(m.radiance, 9), round(m.refractive_index, 9), m.fresnel_reflectivity,
             round(m.chromatic_reflectivity, 9), round(m.diffuse_reflectivity, 9), round(m.sigma, 9),
             tuple(np.round(list(m.transparency), 9))) for m in (st.materials[i] for i in range(st.n_materials))]
    p = flat.params
    prm = (p.image_width, p.image_height, round(p.gamma, 9), p.trace_depth, p.direct_samples, p.path_samples,
           round(p.trace_min_intensity, 9), round(p.max_path_length, 9), round(p.camera_focal_length, 9),
           tuple(np.round(list(p.camera_position), 9)), tuple(np.round(list(p.camera_view_direction), 9)),
           tuple(np.round(list(p.background_color), 9)))
    return nodes, mats, prm, [st.children[i] for i in range(st.n_children)], (st.light_root, st.matter_root)


def test_script_and_api_build_the_same_flat_scene(tmp_path):
    path = tmp_path / "frontend_test.acn"
    path.write_text(SCRIPT)
    sc = acn.Scene()
    assert sc.load_acn(str(path)) == 1                  # one create_image call recorded
    sc.select_image(0)
    a = flat_tables(sc.flatten())
    b = flat_tables(python_twin().flatten())
    assert a[2] == b[2]                                  # render parameters
    assert a[4] == b[4] and a[3] == b[3]                # roots, child lists
    assert len(a[0]) == len(b[0])
    for i, (x, y) in enumerate(zip(a[0], b[0])):
        assert x == y, f"node {i}: script {x} != api {y}"
    assert a[1] == b[1]                                  # materials


def test_syntax_errors_are_reported_not_fatal(tmp_path):
    path = tmp_path / "broken.acn"
    path.write_text("<mclosure_s></>\ndef a = vec( 1, 2 ;\n")
    sc = acn.Scene()
    with pytest.raises(acn.AcnError) as e:
        sc.load_acn(str(path))
    assert e.value.args and "broken.acn" in str(e.value)


REF = "/root/reference/src_acn"


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present (GPU box)")
@pytest.mark.parametrize("name,rel", [("primitives", "primitives.acn"), ("wine_glass", "wine_glass.acn"), ("diamond", "diamond.acn"),
                                      ("ruby_heart", "ruby_heart.acn"), ("hanging_lamp", "hanging_lamp/hanging_lamp.acn")])
def test_committed_flat_scenes_are_what_the_reference_scripts_produce(name, rel):
    sc = acn.Scene()
    assert sc.load_acn(os.path.join(REF, rel)) == 1
    sc.select_image(0)
    a = flat_tables(sc.flatten())
    b = flat_tables(acn.scenes.load(name))
    assert a[2] == b[2] and a[3] == b[3] and a[4] == b[4] and a[1] == b[1]
    assert a[0] == b[0]
