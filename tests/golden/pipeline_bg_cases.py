"""Inputs shared by make_pipeline_bg_golden.py (reference side) and tests/test_pipeline_bg_cpu.py (this package)."""
import numpy as np

TYPE_SENTENCES = [
    "the sky is blue and the ground is green",
    "the sky is purple",
    "the ground is gray",
    "the floor is brown",
    "all things are on green grass and gray road",
    "the land is yellow and the sky is cyan",
    "the bus is orange",
    "the moon in the sky is yellow",
    "Sky: pink!",
    "color it red",
]

COMBINE_CASES = [
    ("the sky is purple", "the sky is blue and the ground is green"),
    ("the ground is gray", "the sky is blue and the ground is green"),
    ("the sky is red and the ground is black", "the sky is blue and the ground is green"),
    ("the sky is pink", "the ground is black"),
    ("the ground is brown", "the sky is cyan"),
    ("the sky is pink", "the sky is cyan"),
    ("the ground is brown", "the ground is black"),
    ("the sky is green", "the sky is blue and the ground is green"),
    ("the floor is orange", "the sky is yellow and the land is red"),
    ("the sky is gray", ""),
]


def gradient_case(name, size=64):
    """(picture uint8 [size,size,3], inner mask int32 [size,size]; 0 = background)."""
    rng = np.random.default_rng({"blue_sky": 1, "two_tone": 2, "low_horizon": 3}[name])
    img = np.zeros((size, size, 3), np.uint8)
    horizon = {"blue_sky": 28, "two_tone": 20, "low_horizon": 31}[name]
    img[:horizon] = {"blue_sky": (70, 130, 230), "two_tone": (200, 90, 160), "low_horizon": (30, 200, 190)}[name]
    img[horizon:] = (60, 160, 70)
    if name == "two_tone":                       # a second, rarer colour in the sampled rows
        img[5:7, :20] = (10, 10, 10)
    mask = np.zeros((size, size), np.int32)
    mask[10:40, 12:30] = 1                       # an instance crossing the horizon
    mask[45:60, 40:60] = 2
    img[mask == 1] = rng.integers(0, 256, (int((mask == 1).sum()), 3), dtype=np.uint8)
    img[mask == 2] = (250, 240, 20)
    return img, mask
