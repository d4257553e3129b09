"""The composite element-wise functions of the reference's n-ary tests (tests/test_elemwise.py:252-305 upstream);
shared by tests/golden/make_golden.py (evaluated by the reference) and tests/test_api_elemwise_nary.py."""

TRINARY = [
    lambda x, y, z: (x + y) * z,
    lambda x, y, z: x * (y + z),
    lambda x, y, z: x * y * z,
    lambda x, y, z: x + y + z,
    lambda x, y, z: x + y - z,
    lambda x, y, z: x - y + z,
]
UNARY_BINARY = [
    lambda x: x * 2 + 1,
    lambda x: abs(x) ** 0.5 - x,
    lambda x, y: (x - y) * (x + y),
    lambda x, y: x * 3 > y,
]
