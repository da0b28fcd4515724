"""FluidSolver.toImage: the reference's only observable output, a plain-text PPM frame.

CPU part: the oracle's restatement against hand-computed known answers of the reference's formulas (F1:341-357: header,
`(int)((1 - d) * 255)` clamped to [0, 255], and the newline after every pixel whose index is a multiple of the solver
width -- so after pixel 0, not at the end of a row; F6:1250-1299: double-width frame with the heat panel on the left).
GPU part: the CUDA path (pixels computed on the device, text formatted by the C-ABI library) must produce the oracle's
integers and the oracle's file byte for byte on the main() scenes of every variant."""
import numpy as np
import pytest

import scenes


def test_oracle_ppm_known_answer_f1(ifl, oracle_cls, tmp_path):
    abi = ifl.abi
    o = oracle_cls(abi, abi.default_config(abi.F1_MATRIXLESS, 3, 2))
    #            shade: (int)((1-d)*255) -> 255, 127 (127.5 truncated), 0, clamp(-255)=0, clamp(510)=255, 254 (254.745)
    o.write("d", np.array([0.0, 0.5, 1.0, 2.0, -1.0, 0.001]))
    img = o.image()
    assert img.shape == (2, 3, 3)
    assert img[:, :, 0].ravel().tolist() == [255, 127, 0, 0, 255, 254]
    assert np.array_equal(img[:, :, 0], img[:, :, 1]) and np.array_equal(img[:, :, 0], img[:, :, 2])
    path = tmp_path / "f1.ppm"
    o.to_image(path)
    text = path.read_bytes().decode()
    # F1:354: `if (i % _w == 0) newline` -> after pixels 0 and 3
    assert text == "P3\n3 2\n255\n255 255 255 \n127 127 127 0 0 0 0 0 0 \n255 255 255 254 254 254 "


def test_oracle_ppm_known_answer_heat(ifl, oracle_cls, tmp_path):
    abi = ifl.abi
    o = oracle_cls(abi, abi.default_config(abi.F6_HEAT, 2, 2))
    tamb = abi.default_config(abi.F6_HEAT, 2, 2).t_ambient
    o.write("d", np.array([0.0, 1.0, 0.25, 0.0]))
    # t = (T - Tamb) / 700: 0 -> (0,0,0) ; 0.25 -> (255,127,0) ; 0.5 -> (255,255,0) ; 1 -> (255,255,255)  (volume = 1)
    o.write("t", np.array([tamb, tamb + 175.0, tamb + 350.0, tamb + 700.0]))
    img = o.image(True)
    assert img.shape == (2, 4, 3)
    assert img[0, 0].tolist() == [0, 0, 0] and img[0, 1].tolist() == [255, 127, 0]
    assert img[1, 0].tolist() == [255, 255, 0] and img[1, 1].tolist() == [255, 255, 255]
    assert img[:, 2:, 0].ravel().tolist() == [255, 0, 191, 255]          # density panel on the right: (int)(0.75*255) = 191
    path = tmp_path / "f6.ppm"
    o.to_image(path, True)
    text = path.read_bytes().decode()
    assert text.startswith("P3\n4 2\n255\n0 0 0 \n255 127 0 255 255 255 \n")    # newline after pixels 0, 2, 4, 6 (i % _w with _w = 2)
    assert o.image(False).shape == (2, 2, 3)
    o5 = oracle_cls(abi, abi.default_config(abi.F3_CONJUGATE_GRADIENTS, 4, 4))
    assert o5.lib.ifo_image_size(o5.h, 1, __import__("ctypes").byref(__import__("ctypes").c_size_t())) != 0   # renderHeat needs F6+


def test_oracle_image_solid_and_volume_variants(ifl, oracle_cls):
    """F4:678 paints SOLID cells black; F5:1142 weights the shade with the fluid volume of the cell."""
    abi = ifl.abi
    w, h = 48, 32
    for variant, spec in ((abi.F4_SOLID_BOUNDARIES, scenes.F4_BODIES), (abi.F5_CURVED_BOUNDARIES, scenes.F5_BODIES)):
        o = oracle_cls(abi, abi.default_config(variant, w, h))
        o.set_bodies(scenes.make_bodies(ifl, spec))
        o.run_stage("fill_solid_fields")
        dens = np.linspace(0.0, 0.9, w * h)
        o.write("d", dens)
        shade = o.image()[:, :, 0].ravel()
        if variant == abi.F4_SOLID_BOUNDARIES:
            solid = o.read("cell_d") == 1.0
            assert solid.any() and not solid.all()
            assert np.all(shade[solid] == 0)
            assert np.array_equal(shade[~solid], ((1.0 - dens[~solid]) * 255.0).astype(np.int64))
        else:
            vol = o.read("volume_d")
            assert (vol < 1.0).any() and (vol > 0.0).any()
            assert np.array_equal(shade, np.clip(((1.0 - dens) * vol * 255.0).astype(np.int64), 0, 255))
        o.close()


def _frames_equal(g, o, tmp_path, render_heat, tag):
    assert np.array_equal(g.image(render_heat), o.image(render_heat)), tag
    pg, po = tmp_path / f"{tag}_gpu.ppm", tmp_path / f"{tag}_ref.ppm"
    g.to_image(pg, render_heat)
    o.to_image(po, render_heat)
    assert pg.read_bytes() == po.read_bytes(), tag


@pytest.mark.gpu
@pytest.mark.parametrize("variant_name,w,h", [("F1_MATRIXLESS", 128, 128), ("F3_CONJUGATE_GRADIENTS", 100, 70)])
def test_to_image_grid_solvers(ifl, oracle_cls, tmp_path, variant_name, w, h):
    abi = ifl.abi
    variant = getattr(abi, variant_name)
    g = ifl.FluidSolver(w, h, scenes.DENSITY, variant=variant, flags=abi.FLAG_SEQUENTIAL_REDUCTIONS)
    o = oracle_cls(abi, abi.default_config(variant, w, h, flags=abi.FLAG_SEQUENTIAL_REDUCTIONS))
    for step in range(4):
        for s in (g, o):
            s.add_inflow(*scenes.inflow_for(abi, variant))
            s.update(scenes.DT)
    _frames_equal(g, o, tmp_path, False, variant_name)
    with pytest.raises(ifl.IflError):
        g.image(True)


@pytest.mark.gpu
@pytest.mark.parametrize("variant_name", ["F4_SOLID_BOUNDARIES", "F5_CURVED_BOUNDARIES", "F6_HEAT", "F7_VARIABLE_DENSITY"])
def test_to_image_solid_variants(ifl, oracle_cls, tmp_path, variant_name):
    from test_parity_solids_gpu import pair
    abi = ifl.abi
    variant = getattr(abi, variant_name)
    bodies, inflow, kw = scenes.solid_scene(abi, variant)
    g, o = pair(ifl, oracle_cls, variant, 96, 64, bodies, flags=abi.FLAG_SEQUENTIAL_REDUCTIONS, **kw)
    for step in range(3):
        for s in (g, o):
            s.add_inflow(*inflow)
            s.update(scenes.DT)
    _frames_equal(g, o, tmp_path, False, variant_name)
    if variant >= abi.F6_HEAT:
        _frames_equal(g, o, tmp_path, True, variant_name + "_heat")


@pytest.mark.gpu
def test_to_image_flip(ifl, oracle_cls, tmp_path):
    import test_parity_flip_gpu as tf
    abi = ifl.abi
    g, o = tf.pair(ifl, oracle_cls, 64, 64, scenes.F7_BODIES, flags=abi.FLAG_SEQUENTIAL_REDUCTIONS)
    for step in range(2):
        for s in (g, o):
            s.update(0.0025)
    _frames_equal(g, o, tmp_path, False, "F8")
    _frames_equal(g, o, tmp_path, True, "F8_heat")
