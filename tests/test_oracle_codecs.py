"""The codec oracle (oracle/lsio_codecs_oracle.c) against liblz4 and numpy (CPU only).

The reference has no codec code (planned only, planned_design.md:83-90), so the oracle restates the
published LZ4 block format and numcodecs.Shuffle; it is pinned here by liblz4 itself (pyarrow's
"lz4_raw" codec): committed vectors compressed by liblz4 (tests/golden/codecs.json) and live
round trips in both directions.
"""
import json
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
from make_codec_vectors import RECIPES, recipe_bytes  # noqa: E402


def _golden():
    return json.load(open(os.path.join(HERE, "golden", "codecs.json")))["vectors"]


def test_lz4_golden_vectors_from_liblz4(oracle):
    vs = _golden()
    assert len(vs) == len(RECIPES)
    for v in vs:
        want = recipe_bytes(v)
        assert len(want) == v["decoded_len"] and oracle.crc32c(want) == v["decoded_crc32c"], v["kind"]
        got = oracle.lz4_decompress(bytes.fromhex(v["lz4_hex"]), v["decoded_len"])
        assert got.tobytes() == want, (v["kind"], v["n"])


def test_lz4_round_trips_against_liblz4(oracle):
    pa = pytest.importorskip("pyarrow")
    codec = pa.Codec("lz4_raw")
    rng = np.random.default_rng(0)
    cases = [recipe_bytes(r) for r in RECIPES]
    cases += [rng.integers(0, 3, n, dtype=np.uint8).tobytes() for n in range(0, 48)]  # the end-of-block rules (last 5 / 12 bytes)
    cases += [rng.integers(0, int(rng.integers(2, 200)), int(rng.integers(1, 50000)), dtype=np.uint8).tobytes() for _ in range(40)]
    for d in cases:
        mine = oracle.lz4_compress(d)
        assert oracle.lz4_decompress(mine, len(d)).tobytes() == d
        if len(d):
            assert codec.decompress(mine.tobytes(), decompressed_size=len(d)).to_pybytes() == d  # liblz4 accepts our blocks
            theirs = codec.compress(d).to_pybytes()
            assert oracle.lz4_decompress(theirs, len(d)).tobytes() == d  # and we decode liblz4's


def test_lz4_decoder_rejects_malformed_blocks(oracle):
    d = recipe_bytes({"kind": "lowent", "n": 3000, "seed": 2, "alphabet": 4})
    c = oracle.lz4_compress(d)
    with pytest.raises(ValueError) as e:
        oracle.lz4_decompress(c, len(d) - 1)  # output too small
    assert e.value.args[0] == -2
    try:  # truncated: an error, or (if the cut falls right after a literal run) a short output -- never the full one
        assert len(oracle.lz4_decompress(c[: len(c) // 2], len(d))) < len(d)
    except ValueError as e:
        assert e.args[0] == -1
    with pytest.raises(ValueError) as e:
        oracle.lz4_decompress(b"", 10)
    assert e.value.args[0] == -1
    with pytest.raises(ValueError) as e:
        oracle.lz4_decompress(bytes([0x10, 0x41, 0x00, 0x00, 0x50, 1, 2, 3, 4, 5]), 100)  # offset 0
    assert e.value.args[0] == -3
    with pytest.raises(ValueError) as e:
        oracle.lz4_decompress(bytes([0x10, 0x41, 0x02, 0x00, 0x50, 1, 2, 3, 4, 5]), 100)  # offset 2 with 1 byte of history
    assert e.value.args[0] == -3


@pytest.mark.parametrize("elem", [1, 2, 3, 4, 8, 16])
def test_shuffle_matches_numpy_transpose(oracle, elem):
    rng = np.random.default_rng(elem)
    for n in (0, 1, elem, elem * 7, elem * 1000 + elem - 1 if elem > 1 else 1000, 4096):
        d = rng.integers(0, 256, n, dtype=np.uint8)
        cnt = n // elem
        want = np.concatenate([d[: cnt * elem].reshape(cnt, elem).T.ravel(), d[cnt * elem:]])  # numcodecs.Shuffle
        got = oracle.shuffle(d, elem)
        assert np.array_equal(got, want), (elem, n)
        assert np.array_equal(oracle.unshuffle(got, elem), d)


def test_crc32c_trailer_residue(oracle, lsio):
    """Zarr v3 crc32c codec: CRC-32C(data || LE32(CRC-32C(data))) is a constant."""
    import struct

    rng = np.random.default_rng(1)
    for n in (0, 1, 9, 4096, 100001):
        d = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        stored = d + struct.pack("<I", oracle.crc32c(d))
        assert oracle.crc32c(stored) == lsio.codecs.CRC32C_RESIDUE
        assert lsio.codecs.crc32c_trailer_ok(oracle.crc32c(stored))
        if n:
            bad = bytearray(stored)
            bad[n // 2] ^= 1
            assert not lsio.codecs.crc32c_trailer_ok(oracle.crc32c(bytes(bad)))
