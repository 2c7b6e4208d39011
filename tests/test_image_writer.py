"""pt_save_image / pt_image_to_rgb8 against the reference's save path (src/main.cpp:118-139, src/image.cpp:46-88)."""
import os

import numpy as np
import pytest
from PIL import Image

from conftest import f32
from oracle_py import Ref


def test_pixels_match_the_reference_writer_golden(pt, ref_gold, tmp_path):
    k = ref_gold["save_image"]
    W, H = k["W"], k["H"]
    rgb = f32(k["rgb"]).reshape(-1, 3)
    want = np.array(k["png_pixels"], np.uint8).reshape(H, W, 3)
    assert (pt.image_to_rgb8(rgb, W, H) == want).all()
    name = pt.save_image(rgb, W, H, str(tmp_path / k["name_rule"][0]), frame=k["name_rule"][1])
    assert os.path.basename(name) == k["name_rule"][2]
    assert (np.asarray(Image.open(name).convert("RGB")) == want).all()


def test_conversion_rules(pt):
    """mirror in x, rows top-down, clamp(f*255,0,255) truncated"""
    W, H = 3, 2
    rgb = np.zeros((H, W, 3), np.float32)
    rgb[0, 0] = [1.0, 0.5, 0.25]       # buffer x=0 is the RIGHT edge of the picture
    rgb[1, 2] = [2.0, -1.0, 0.999]
    out = pt.image_to_rgb8(rgb.reshape(-1, 3), W, H)
    assert out[0, 2].tolist() == [255, 127, 63]
    assert out[1, 0].tolist() == [255, 0, 254]
    assert out.sum() == 255 + 127 + 63 + 255 + 254


def test_name_rules_and_bmp(pt, tmp_path):
    rgb = np.random.default_rng(0).random((4 * 5, 3), dtype=np.float32)
    want = pt.image_to_rgb8(rgb, 5, 4)
    n1 = pt.save_image(rgb, 5, 4, str(tmp_path / "test.bmp"), frame=0, force_png=True)   # headless default
    assert n1.endswith("test.0.png") and (np.asarray(Image.open(n1).convert("RGB")) == want).all()
    n2 = pt.save_image(rgb, 5, 4, str(tmp_path / "test.bmp"), frame=12, force_png=False)  # reference behaviour
    assert n2.endswith("test.12.bmp") and (np.asarray(Image.open(n2).convert("RGB")) == want).all()
    n3 = pt.save_image(rgb, 5, 4, str(tmp_path / "noext"), frame=1)
    assert n3.endswith("noext") and Image.open(n3).format == "PNG"
    with pytest.raises(pt.PtError):
        pt.save_image(rgb, 5, 4, str(tmp_path / "no_such_dir" / "x.png"))


@pytest.mark.skipif(not Ref.available(), reason="needs oracle/_ref")
def test_against_the_reference_writer_on_a_larger_image(pt, tmp_path):
    rng = np.random.default_rng(4)
    W, H = 67, 41
    rgb = rng.uniform(-0.1, 1.2, (W * H, 3)).astype(np.float32)
    ref_name = Ref().save_image(rgb, W, H, str(tmp_path / "r.png"), 2)
    ours = pt.save_image(rgb, W, H, str(tmp_path / "o.png"), 2)
    assert (np.asarray(Image.open(ref_name).convert("RGB")) == np.asarray(Image.open(ours).convert("RGB"))).all()
