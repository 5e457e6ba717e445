"""Run-time specialisation, host side (no GPU needed: NVRTC cross-compiles for sm_100a): the source SpecGen writes for a
scene's structure, which scenes qualify, and that it compiles together with the embedded device sources."""
import re

import pytest

import actinon_b200 as acn


def test_generated_source_restates_the_scene_structure():
    flat = acn.scenes.load("wine_glass")
    src, _ = acn.spec_probe(flat)
    assert "#define ACN_SPEC_SCENE 1" in src
    # the glass (node 4: 9 variables, truth table) and the liquid (node 28) are the two composite objects of the scene
    assert "csg_spec_4(" in src and "csg_spec_28(" in src
    assert "9 variables, truth table" in src
    # light sphere, floor plane, glass, liquid: four unrolled element tests, kinds literal
    assert "4 element tests" in src
    assert len(re.findall(r"prim_hit\( sv, \d+, \d+, ray", src)) == 2
    # the three sub-envelopes of the glass gate their variables
    assert src.count("gates variables") == 3


def test_same_structure_gives_the_same_source_whatever_the_geometry():
    """Frames of the video differ in geometry only (camera / rotation): one compiled module serves all of them."""
    a, _ = acn.spec_probe(acn.scenes.load("diamond_video_000010"))
    b, _ = acn.spec_probe(acn.scenes.load("diamond_video_000080"))
    assert a == b and len(a) > 1000


def test_scene_with_a_deep_compound_tree_does_not_qualify():
    src, _ = acn.spec_probe(acn.scenes.load("many_spheres"))     # 37 449 nodes in an 8-ary tree of compounds
    assert src == ""


@pytest.mark.parametrize("name,precision", [("wine_glass", acn.PRECISION_F32), ("diamond", acn.PRECISION_F32), ("primitives", acn.PRECISION_F64)])
def test_generated_source_compiles_for_sm_100a(name, precision, tmp_path, monkeypatch):
    monkeypatch.setenv("ACN_CACHE_DIR", str(tmp_path))
    flat = acn.scenes.load(name)
    src, sec = acn.spec_probe(flat, acn.Options(precision=precision, csg_mode=acn.CSG_INTERVALS), compile=True)
    assert len(src) > 0 and sec > 0
    assert len(list(tmp_path.glob("*.cubin"))) == 1 and len(list(tmp_path.glob("*.names"))) == 1
