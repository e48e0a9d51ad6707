"""Small synthetic mesh files for the reader tests (written by hand, in the formats of the reference's samples/data):
a 1-based TetGen pair with one inverted tet and rows out of order (the first row must carry index 1: that is how the
reference detects 1-based files), a 0-based pair, and an OBJ with `a/b/c` face
tokens, upper-case records, comments, texture / normal records and a degenerate (skipped) face."""
import os

ELE_ONE_BASED = """4  4  0
    1     1 2 3 4
    4     1 3 2 6
    2     2 3 4 5
    3     2 4 3 6
# trailing comment
"""
NODE_ONE_BASED = """6  3  0  0
   1    0.0  0.0  0.0
   2    1.0  0.0  0.0
   3    0.0  1.0  0.0
   4    0.0  0.0  1.0
   6    0.333333333333333  0.25  -0.7071067811865476
   5    1.1  1.2000000000000002  1.3
"""
ELE_ZERO_BASED = """2 4 0
0 0 1 2 3
1 1 2 3 4
"""
NODE_ZERO_BASED = """5 3 0 0
0 0 0 0
1 0.1 0 0
2 0 0.2 0
3 0 0 0.30000001192092896
4 0.123456789 0.987654321 0.5
"""
OBJ = """# OBJ written by hand
v -0.72375 -0.5 0
v -0.72375 -0.4 0.01
V 0.1 -0.5 0.02
v 0.1 -0.4 1e-3
vt 0.5 0.5
vn 0 0 1
v 0.5 0.25 0.125 1.0
f 1 2 3
f 2/1/1 4/1/1 3/1/1
F 3//1 4//1 5//1
f 1 2
g group
f 5 4 1 2
"""


def write_all(d):
    paths = {}
    for name, ele, node in (("one", ELE_ONE_BASED, NODE_ONE_BASED), ("zero", ELE_ZERO_BASED, NODE_ZERO_BASED)):
        with open(os.path.join(d, name + ".ele"), "w") as f:
            f.write(ele)
        with open(os.path.join(d, name + ".node"), "w") as f:
            f.write(node)
        paths[name] = (os.path.join(d, name), "elenode")
    with open(os.path.join(d, "hand.obj"), "w") as f:
        f.write(OBJ)
    paths["hand"] = (os.path.join(d, "hand.obj"), "obj")
    return paths
