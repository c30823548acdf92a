"""Inputs shared by make_pipeline_golden.py (reference side) and tests/test_pipeline_cpu.py (this package)."""
import numpy as np

SENTENCES = [
    "the bus on the left is yellow with blue windows",
    "the bus is orange",
    "the red car on the right is big",
    "the person in the middle has black hair, in red shirt and blue pants.",
    "a man with blue pants has red shirt",
    "the two trees on the left are green",
    "all the people have red shirts",
    "the dog near the house is brown with white ears",
    "the cars behind the bus are blue",
    "color the road gray",
    "the house on the right has red roof with yellow walls",
    "this is a bench",
]


def road_sketch(name, size=192):
    """uint8 [size,size,3] white canvas with black (or grey) strokes."""
    s = np.full((size, size, 3), 255, dtype=np.uint8)
    if name == "two_edges_vertical":          # two vertical edges: every row crosses 2 strokes
        s[10:180, 60:63] = 0
        s[10:180, 120:123] = 0
    elif name == "two_edges_horizontal":      # two horizontal edges: every column crosses 2 strokes
        s[70:73, 8:185] = 0
        s[130:133, 8:185] = 0
    elif name == "single_line":
        s[95:98, 5:190] = 0
    elif name == "grey_edges":                # anti-aliased (grey) strokes, some above the 235 whitening threshold
        s[40:43, 10:180] = 120
        s[150:153, 10:180] = 200
        s[100:102, 10:180] = 240
    elif name == "diagonal_pair":
        for i in range(20, 170):
            s[i, i - 10:i - 7] = 0
            s[i, i + 10:i + 13] = 0
    return s
