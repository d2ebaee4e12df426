"""install(): the run-time binding of an importable pynbody to the GPU path swaps exactly the two native
modules on the path (and the KDTree host class) and can be undone.  Exercised against a stand-in
package with pynbody's module layout (the real package needs its compiled extensions to import)."""
import sys
import types

import pytest


@pytest.fixture
def fake_pynbody():
    root = types.ModuleType("fakenbody")
    root.__path__ = []
    root.config = {"sph": {"approximate-fast-images": True, "threaded-image": True}}
    kdt = types.ModuleType("fakenbody.kdtree")
    kdt.kdmain = object()
    kdt.KDTree = type("KDTree", (), {})
    sph = types.ModuleType("fakenbody.sph")
    sph.__path__ = []
    rnd = types.ModuleType("fakenbody.sph._render")
    rnd.render_image = lambda *a: "cpu image"
    rnd.to_3d_grid = lambda *a: "cpu grid"
    rnd.render_spherical_image_core = lambda *a: "cpu healpix"
    rnd.render_something_else = lambda *a: "stays"
    filt = types.ModuleType("fakenbody.filt")
    filt.__path__ = []
    gsel = types.ModuleType("fakenbody.filt.geometry_selection")
    gsel.selection = lambda *a: "cpu selection"
    mods = {"fakenbody": root, "fakenbody.kdtree": kdt, "fakenbody.sph": sph, "fakenbody.sph._render": rnd,
            "fakenbody.filt": filt, "fakenbody.filt.geometry_selection": gsel}
    sys.modules.update(mods)
    root.kdtree, root.sph, sph._render, root.filt, filt.geometry_selection = kdt, sph, rnd, filt, gsel
    yield root
    for m in mods:
        sys.modules.pop(m, None)


def test_install_swaps_only_the_hot_path(fake_pynbody):
    from pynbody_b200 import install
    from pynbody_b200.kdtree import KDTree, kdmain
    from pynbody_b200.sph import _render
    before = (fake_pynbody.kdtree.kdmain, fake_pynbody.kdtree.KDTree, fake_pynbody.sph._render.render_image)
    install.install(fake_pynbody)
    assert fake_pynbody.kdtree.kdmain is kdmain and fake_pynbody.kdtree.KDTree is KDTree
    assert fake_pynbody.sph._render.render_image is _render.render_image
    assert fake_pynbody.sph._render.to_3d_grid is _render.to_3d_grid
    assert fake_pynbody.sph._render.render_spherical_image_core is _render.render_spherical_image_core
    assert fake_pynbody.sph._render.render_something_else() == "stays"
    from pynbody_b200.filt import geometry_selection
    assert fake_pynbody.filt.geometry_selection.selection is geometry_selection.selection
    assert fake_pynbody.config["sph"]["approximate-fast-images"] is False
    install.uninstall(fake_pynbody)
    assert (fake_pynbody.kdtree.kdmain, fake_pynbody.kdtree.KDTree, fake_pynbody.sph._render.render_image) == before
    assert fake_pynbody.sph._render.render_spherical_image_core() == "cpu healpix"
    assert fake_pynbody.filt.geometry_selection.selection() == "cpu selection"
