"""Deterministic, dependency-free input recipes shared by make_golden.py and the tests."""

WORDS = ("the of and to in is that for it as was with be by on not he this are or his from at which but have an had "
         "they you were their one all we can her has there been if more when will would who so no").split()


def _lcg(seed):
    s = seed & 0xFFFFFFFF
    while True:
        s = (s * 1664525 + 1013904223) & 0xFFFFFFFF
        yield s >> 8


def text(n, seed):
    g = _lcg(seed)
    out = bytearray()
    k = 0
    while len(out) < n:
        r = next(g)
        # square the uniform to skew towards frequent words
        idx = ((r & 0xFFF) * (r & 0xFFF) * len(WORDS)) >> 24
        out += WORDS[idx].encode()
        k += 1
        out += b"\n" if k % 12 == 0 else b" "
    return bytes(out[:n])


def rand(n, seed):
    g = _lcg(seed)
    return bytes(next(g) & 0xFF for _ in range(n))


def few_symbols(n, seed, k=3):
    g = _lcg(seed)
    return bytes(97 + next(g) % k for _ in range(n))


def binary_struct(n, seed):
    """records with high bytes (> 127) so Huffman weights need FSE compression"""
    g = _lcg(seed)
    out = bytearray()
    while len(out) < n:
        r = next(g)
        out += bytes([0xFF, 0xFE, r & 0x0F | 0xA0, 0x00, 0x00, (r >> 4) & 0xFF, 0x80 | (r >> 12) & 0x3F, 0xC3])
    return bytes(out[:n])


RECIPES = {
    "empty": lambda: b"",
    "a": lambda: b"a",
    "hello": lambda: b"hello world\n",
    "zeros1000": lambda: bytes(1000),
    "hello_zarc_x1000": lambda: b"hello zarc " * 1000,
    "text_300": lambda: text(300, 1),
    "text_1000": lambda: text(1000, 2),
    "text_5000": lambda: text(5000, 3),
    "text_20000": lambda: text(20000, 4),
    "text_150000": lambda: text(150000, 5),
    "text_300000": lambda: text(300000, 6),
    "rand_100": lambda: rand(100, 7),
    "rand_5000": lambda: rand(5000, 8),
    "three_syms_4000": lambda: few_symbols(4000, 9),
    "two_syms_900": lambda: few_symbols(900, 10, 2),
    "binary_8000": lambda: binary_struct(8000, 11),
    "mixed_30000": lambda: text(10000, 12) + rand(5000, 13) + text(10000, 12) + bytes(3000) + few_symbols(2000, 14),
    "rle_then_text": lambda: b"x" * 70000 + text(3000, 15),
    "records_rle_mode": lambda: b"".join(b"RECORD_" + bytes([v]) for v in range(255)),
    "long_repeat_200000": lambda: text(700, 16) * 286,
}


def make_input(name):
    return RECIPES[name]()
