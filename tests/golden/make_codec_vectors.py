"""Generates tests/golden/codecs.json: LZ4 block vectors compressed by liblz4 ITSELF (through pyarrow's
"lz4_raw" codec), so the C oracle decoder (oracle/lsio_codecs_oracle.c) is pinned by something that is
not itself.  The reference holds no codec code or vectors (lsio_codecs is planned only:
planned_design.md:83-90), so there is nothing to import from it.

Inputs are generated from small seeds/recipes so the fixture stays small; each vector stores the recipe,
the hex of the liblz4-compressed block, and CRC-32C + length of the decoded bytes.
Run:  python tests/golden/make_codec_vectors.py
"""
import json
import os
import sys

import numpy as np
import pyarrow as pa

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))


def recipe_bytes(r):
    """Deterministic input for a recipe dict (shared with tests/test_oracle_codecs.py)."""
    kind, n, seed = r["kind"], r["n"], r.get("seed", 0)
    rng = np.random.default_rng(seed)
    if kind == "zeros":
        return bytes(n)
    if kind == "text":
        unit = b"light speed io lands byte ranges in HBM; "
        return (unit * (n // len(unit) + 1))[:n]
    if kind == "random":
        return rng.integers(0, 256, n, dtype=np.uint8).tobytes()
    if kind == "lowent":
        return rng.integers(0, r.get("alphabet", 4), n, dtype=np.uint8).tobytes()
    if kind == "walk_f32":
        return np.cumsum(rng.integers(-3, 4, n // 4)).astype("<f4").tobytes()
    if kind == "walk_f32_shuffled":
        a = np.frombuffer(np.cumsum(rng.integers(-3, 4, n // 4)).astype("<f4").tobytes(), dtype=np.uint8)
        return a.reshape(-1, 4).T.copy().tobytes()
    if kind == "period":
        p = r["period"]
        unit = rng.integers(0, 256, p, dtype=np.uint8).tobytes()
        return (unit * (n // p + 1))[:n]
    if kind == "far":  # a match at the maximum offset, 65535: 2000 random bytes, zeros up to 65535, then the start again
        unit = rng.integers(0, 256, 2000, dtype=np.uint8).tobytes() + bytes(63535)
        return (unit * (n // 65535 + 1))[:n]
    raise ValueError(kind)


RECIPES = [
    {"kind": "text", "n": 1},
    {"kind": "text", "n": 13},
    {"kind": "text", "n": 300},
    {"kind": "zeros", "n": 70000},  # one match with a 270-byte length extension
    {"kind": "random", "n": 1000, "seed": 1},  # incompressible: one long literal run
    {"kind": "lowent", "n": 3000, "seed": 2, "alphabet": 4},
    {"kind": "lowent", "n": 2000, "seed": 3, "alphabet": 2},
    {"kind": "walk_f32", "n": 4096, "seed": 4},
    {"kind": "walk_f32_shuffled", "n": 4096, "seed": 4},
    {"kind": "period", "n": 5000, "seed": 5, "period": 1},
    {"kind": "period", "n": 5000, "seed": 6, "period": 3},
    {"kind": "period", "n": 5000, "seed": 7, "period": 33},
    {"kind": "far", "n": 70000, "seed": 8},
]


def main():
    from oracle import oracle

    codec = pa.Codec("lz4_raw")
    out = []
    for r in RECIPES:
        d = recipe_bytes(r)
        c = codec.compress(d).to_pybytes()
        assert codec.decompress(c, decompressed_size=len(d)).to_pybytes() == d
        v = dict(r)
        v["decoded_len"] = len(d)
        v["decoded_crc32c"] = oracle.crc32c(d)
        v["lz4_hex"] = c.hex()
        out.append(v)
    with open(os.path.join(HERE, "codecs.json"), "w") as f:
        json.dump({"generator": f"pyarrow {pa.__version__} Codec('lz4_raw') (liblz4)", "vectors": out}, f, indent=0)
    print(len(out), "vectors,", os.path.getsize(os.path.join(HERE, "codecs.json")), "bytes")


if __name__ == "__main__":
    main()
