"""Synthetic lenslet (plenoptic) image generator -- SURVEY.md §8(d).

8-bit (or 10-bit) 4:2:0 picture: hexagonal micro-lens lattice, pitch 15 px, per-lens radial vignette
multiplying a smooth scene, plus Gaussian noise.  Chroma is flat mid-grey.  Everything is seeded, so
the reference encoder, the oracle and the CUDA path see identical inputs.
"""
import numpy as np


def lenslet_luma(width, height, seed=0, pitch=15.0, bit_depth=8, crop_of=None):
    """Return the luma plane (height x width, uint16) of the synthetic lenslet image.

    crop_of=(W, H): the top-left width x height region of the W x H image with this seed (the generator is
    point-wise except for the noise, whose stream is row-major: the first `height` rows of the full image use the
    first height*W normal draws), so a region of the Illum-size image costs what the region costs."""
    rng = np.random.default_rng(seed)
    full_w = width if crop_of is None else int(crop_of[0])
    if crop_of is not None and (width > crop_of[0] or height > crop_of[1]):
        raise ValueError("crop larger than the image")
    yy, xx = np.mgrid[0:height, 0:width].astype(np.float64)
    row_h = pitch * np.sqrt(3.0) / 2.0
    row = np.floor(yy / row_h + 0.5)
    shift = np.where((row.astype(np.int64) & 1) == 1, pitch / 2.0, 0.0)
    cx = (np.floor((xx - shift) / pitch + 0.5)) * pitch + shift
    cy = row * row_h
    r = np.sqrt((xx - cx) ** 2 + (yy - cy) ** 2)
    vignette = np.clip(1.2 - r / (pitch / 2.0), 0.0, 1.0)
    scene = 128.0 + 60.0 * np.sin(xx / 37.0) + 40.0 * np.cos(yy / 23.0)
    img = scene * vignette + rng.normal(0.0, 2.0, size=(height, full_w))[:, :width]
    img = np.clip(np.rint(img), 0, 255)
    if bit_depth > 8:
        img = img * (1 << (bit_depth - 8))
    return img.astype(np.uint16)


def write_yuv420(path, luma, bit_depth=8, append=False):
    """Write (or append) one 4:2:0 frame (flat chroma) in the raw planar layout TVideoIOYuv reads."""
    h, w = luma.shape
    mid = 1 << (bit_depth - 1)
    dt = np.uint8 if bit_depth == 8 else np.dtype("<u2")
    with open(path, "ab" if append else "wb") as f:
        f.write(luma.astype(dt).tobytes())
        c = np.full((h // 2, w // 2), mid, dtype=dt)
        f.write(c.tobytes())
        f.write(c.tobytes())
