"""Camera-frame pre-processing (the step in front of the hot path).

CPU: the numpy oracle (``oracle/preprocess.py``) against the golden fixture produced by the reference's own
``preprocess_image_batch`` (Pillow + torchvision, ``tests/golden/make_golden_preprocess.py``) - bit-exact uint8 tiles
(SHA-256 over all pixels), identical normalised float32 values, and the host-side tables of the product.
GPU: ``simlingo_b200.preprocess`` / the drop-in ``simlingo_training.utils.internvl2_utils.preprocess_image_batch``
against the oracle - bit-exact (integer resampling; the fp32 normalisation uses torch's operation order and is rounded
once to bf16)."""
import hashlib
import os

import numpy as np
import pytest
import torch

from oracle import preprocess as P

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "preprocess.npz")


@pytest.fixture(scope="module")
def golden():
    return np.load(GOLDEN)


@pytest.mark.parametrize("name", list(P.CASES))
def test_oracle_matches_reference_run(golden, name):
    h, w, seed = P.CASES[name]
    tiles = P.preprocess_tiles_u8(P.synth_camera(h, w, seed))
    assert tiles.shape == (2, 3, 448, 448)
    assert np.array_equal(tiles[:, :, ::5, ::5], golden[name + "_sub5"])
    digest = np.frombuffer(hashlib.sha256(np.ascontiguousarray(tiles).tobytes()).digest(), dtype=np.uint8)
    assert np.array_equal(digest, golden[name + "_sha256"]), "oracle tiles differ from the reference's Pillow output"
    assert np.array_equal(P.normalize(tiles)[:, :, 0, :16], golden[name + "_norm_row0"])
    assert golden[name + "_image_sizes"].tolist() == [[h, w]]
    # use_global_img=True: the reference appends the whole frame resized to one 448 x 448 tile
    with_thumb = P.preprocess_tiles_u8(P.synth_camera(h, w, seed), use_thumbnail=True)
    assert with_thumb.shape == (3, 3, 448, 448) and np.array_equal(with_thumb[:2], tiles)
    digest = np.frombuffer(hashlib.sha256(np.ascontiguousarray(with_thumb[2]).tobytes()).digest(), dtype=np.uint8)
    assert np.array_equal(digest, golden[name + "_thumb_sha256"]), "oracle thumbnail tile differs from the reference's Pillow output"


def test_tile_grid_and_tables_of_the_product_match_oracle():
    from simlingo_b200.preprocess import resample_table, tile_grid
    for (w, h) in [(1024, 359), (1024, 512), (203, 101), (448, 448), (300, 900), (896, 448)]:
        assert tile_grid(w, h) == P.tile_grid(w, h)
    for (i, o) in [(1024, 896), (359, 448), (512, 448), (203, 896), (101, 448), (2000, 448)]:
        a, b = resample_table(i, o), P.resample_coeffs(i, o)
        assert all(np.array_equal(x, y) for x, y in zip(a, b))
    f, c, t = resample_table(448, 448)
    assert f.tolist() == list(range(448)) and set(c.tolist()) == {1} and set(t.ravel().tolist()) == {1 << 22}


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(P.CASES) + ["noop_448x896", "tall_700x300"])
def test_gpu_preprocess_is_bit_exact(name):
    from simlingo_training.utils.internvl2_utils import preprocess_image_batch
    h, w, seed = P.CASES.get(name, {"noop_448x896": (448, 896, 21), "tall_700x300": (700, 300, 22)}.get(name))
    imgs = [P.synth_camera(h, w, seed + k) for k in range(3)]
    res = preprocess_image_batch([torch.from_numpy(i) for i in imgs], input_size=448, use_global_img=False, max_num_grid=2)
    pv = res["pixel_values"]
    assert pv.is_cuda and pv.dtype == torch.bfloat16 and pv.shape == (3, 2, 3, 448, 448)
    assert res["image_sizes"].tolist() == [[h, w]] * 3
    for k, img in enumerate(imgs):
        ref = torch.from_numpy(P.normalize(P.preprocess_tiles_u8(img))).to(torch.bfloat16)
        assert torch.equal(pv[k].cpu(), ref), f"{name}[{k}]: {(pv[k].cpu().float() - ref.float()).abs().max().item()}"


@pytest.mark.gpu
def test_gpu_thumbnail_tile_is_bit_exact():
    """use_global_img=True (reference internvl2_utils.py:262-265): [2 grid tiles | whole frame as a third tile]"""
    from simlingo_training.utils.internvl2_utils import preprocess_image_batch
    h, w, seed = P.CASES["agent_359x1024"]
    img = P.synth_camera(h, w, seed)
    pv = preprocess_image_batch([torch.from_numpy(img)], input_size=448, use_global_img=True, max_num_grid=2)["pixel_values"]
    ref = torch.from_numpy(P.normalize(P.preprocess_tiles_u8(img, use_thumbnail=True))).to(torch.bfloat16)
    assert pv.shape == (1, 3, 3, 448, 448) and torch.equal(pv[0].cpu(), ref)


@pytest.mark.gpu
def test_gpu_preprocess_feeds_the_model():
    """pixel_values plug straight into DrivingInput.camera_images ([B, 1, tiles, 3, 448, 448], agent_simlingo.py:497-502)."""
    from simlingo_b200.preprocess import preprocess_frames
    x = torch.from_numpy(np.stack([P.synth_camera(359, 1024, 5)])).cuda()
    pv = preprocess_frames(x)
    cam = pv.view(1, 1, 2, 3, 448, 448)
    assert cam.dtype == torch.bfloat16 and float(cam.float().abs().max()) < 3.0


def test_resample_table_properties():
    """Size-independent properties of the Pillow tables for arbitrary sizes: taps of an output sample sum to one (22-bit
    fixed point, within the rounding of its taps), windows stay inside the input and move monotonically."""
    from hypothesis import given, settings, strategies as st
    from simlingo_b200.preprocess import resample_table, tile_grid

    @settings(max_examples=40, deadline=None, derandomize=True)
    @given(st.integers(8, 1500), st.sampled_from([448, 896]))
    def check(in_size, out_size):
        first, count, taps = resample_table(in_size, out_size)
        assert first.shape == (out_size,) and taps.shape[0] == out_size
        assert (count >= 1).all() and (first >= 0).all() and (first + count <= in_size).all()
        assert (np.diff(first) >= 0).all()
        used = np.arange(taps.shape[1])[None, :] < count[:, None]
        assert (taps[~used] == 0).all()
        assert np.abs(taps.sum(1) - (1 << 22)).max() <= taps.shape[1]
        if in_size != out_size:  # equal sizes: Pillow skips the pass, the product table is the identity
            ofirst, ocount, otaps = P.resample_coeffs(in_size, out_size)
            assert np.array_equal(first, ofirst) and np.array_equal(count, ocount) and np.array_equal(taps, otaps)
        else:
            assert np.array_equal(first, np.arange(out_size)) and (count == 1).all()

    check()
    check.hypothesis.inner_test(448, 448)
    check.hypothesis.inner_test(1024, 896)

    @settings(max_examples=60, deadline=None, derandomize=True)
    @given(st.integers(16, 2000), st.integers(16, 2000))
    def grid(w, h):
        gw, gh = tile_grid(w, h)
        assert (gw, gh) == P.tile_grid(w, h) and 1 <= gw * gh <= 2

    grid()
