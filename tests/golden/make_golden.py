"""Generates tests/golden/checksums.json and tests/golden/pattern.json.

Independent sources, so the C oracle is pinned by something that is not itself:
  * CRC-32C: RFC 3720 appendix B.4 known answers (hard-coded below) + a pure-Python bitwise
    CRC over the other inputs;
  * XXH64: the `xxhash` Python package (python-xxhash, wraps the upstream C library);
  * synthetic pattern: a pure-Python splitmix64.
The reference itself (Rust) stores no golden byte vectors: its tests use random data
(crates/lsio_uring/tests/integration_test.rs:22-31), so there is nothing to import from it.
Run:  python tests/golden/make_golden.py
"""
import json
import os

import xxhash

HERE = os.path.dirname(os.path.abspath(__file__))
POLY = 0x82F63B78
M64 = (1 << 64) - 1


def crc32c_py(data, crc=0):
    c = crc ^ 0xFFFFFFFF
    for b in data:
        c ^= b
        for _ in range(8):
            c = (c >> 1) ^ (POLY & -(c & 1))
    return c ^ 0xFFFFFFFF


def splitmix64(x):
    z = (x + 0x9E3779B97F4A7C15) & M64
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M64
    return z ^ (z >> 31)


def pattern_byte(file_id, off):
    w = splitmix64(0x4C53494F ^ ((file_id * 0x9E3779B97F4A7C15) & M64) ^ (off >> 3))
    return (w >> (8 * (off & 7))) & 0xFF


def inputs():
    yield "empty", b""
    yield "123456789", b"123456789"
    yield "zeros32", bytes(32)
    yield "ones32", b"\xff" * 32
    yield "inc32", bytes(range(32))
    yield "dec32", bytes(range(31, -1, -1))
    for n in (1, 2, 3, 4, 5, 7, 8, 15, 16, 17, 31, 32, 33, 63, 64, 65, 100, 511, 512, 513, 4095, 4096, 4097):
        yield f"pattern_f3_o11_{n}", bytes(pattern_byte(3, 11 + i) for i in range(n))
    yield "lsio_bench_block_16k", bytes(i & 0xFF for i in range(16384))  # `i as u8`, lsio_bench main.rs:102


RFC3720 = {"zeros32": 0x8A9136AA, "ones32": 0x62A8AB43, "inc32": 0x46DD794E, "dec32": 0x113FDB5C, "123456789": 0xE3069283, "empty": 0}

vec = []
for name, data in inputs():
    c = crc32c_py(data)
    if name in RFC3720:
        assert c == RFC3720[name], (name, hex(c))
    vec.append({"name": name, "hex": data.hex(), "crc32c": c, "xxh64": xxhash.xxh64(data, seed=0).intdigest(),
                "xxh64_seed7": xxhash.xxh64(data, seed=7).intdigest()})
# the full lsio_bench block: stored by value, input regenerated in the test
blk = bytes(i & 0xFF for i in range(262144))
vec.append({"name": "lsio_bench_block_256k", "gen": "i_as_u8", "len": 262144, "crc32c": 0x1CA1B1D2, "xxh64": xxhash.xxh64(blk).intdigest()})
assert xxhash.xxh64(blk).intdigest() == 0xC1034584416AA22C
json.dump(vec, open(os.path.join(HERE, "checksums.json"), "w"), indent=0)

pat = []
for file_id, off, n in [(0, 0, 64), (1, 0, 64), (7, 5, 40), (499, 41214400 - 33, 33), (9999, 1 << 20, 24), (2**32 + 5, 2**33 + 3, 32)]:
    pat.append({"file_id": file_id, "off": off, "hex": bytes(pattern_byte(file_id, off + i) for i in range(n)).hex()})
json.dump(pat, open(os.path.join(HERE, "pattern.json"), "w"), indent=0)
print("wrote", len(vec), "checksum vectors and", len(pat), "pattern vectors")
