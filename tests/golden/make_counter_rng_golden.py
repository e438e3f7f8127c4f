"""Generates tests/golden/counter_rng.json: known answers of the harness's index-addressable generator, computed by a
pure-Python splitmix64 that shares no code with oracle/ndarray_oracle.cpp or ndarray_b200/csrc/rng.cu.

    python tests/golden/make_counter_rng_golden.py

value(seed, i) = splitmix64 output number i of a generator seeded with `seed`:
    z = seed + (i + 1) * 0x9E3779B97F4A7C15;  z = (z ^ z >> 30) * 0xBF58476D1CE4E5B9;  z = (z ^ z >> 27) * 0x94D049BB133111EB;  z ^= z >> 31
f32: (z >> 40) * 2^-23 - 1     f64: (z >> 11) * 2^-52 - 1     i32: top 32 bits     i64: all 64 bits (two's complement)
The ndarray reference is Rust and cannot be imported here; its own golden vectors are transcribed as literals (with file:line)
in tests/test_oracle_golden.py.  This fixture pins the one piece of test infrastructure that is NOT ndarray's: the data generator
both the device kernel and the oracle must agree on, so that they cannot drift together.
"""
import json
import struct
from pathlib import Path

M = (1 << 64) - 1


def splitmix(seed, i):
    z = (seed + (i + 1) * 0x9E3779B97F4A7C15) & M
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M
    return z ^ (z >> 31)


def f32_bits(x):
    return struct.unpack("<I", struct.pack("<f", x))[0]


def main():
    cases = []
    for seed in (0x5EED0001, 0x5EED0002, 7, 0):
        for first in (0, 1, 4095, 2 ** 32 - 2, 2 ** 40 + 12345):
            idx = [first + k for k in range(6)]
            h = [splitmix(seed, i) for i in idx]
            cases.append({
                "seed": seed, "first": first,
                "f32_bits": [f32_bits((z >> 40) * 2.0 ** -23 - 1.0) for z in h],
                "f64": [((z >> 11) * 2.0 ** -52 - 1.0).hex() for z in h],
                "i32": [((z >> 32) ^ (1 << 31)) - (1 << 31) for z in h],
                "i64": [(z ^ (1 << 63)) - (1 << 63) for z in h],
            })
    out = Path(__file__).with_name("counter_rng.json")
    out.write_text(json.dumps({"generator": "splitmix64 output i, see make_counter_rng_golden.py", "cases": cases}, indent=1) + "\n")
    print("wrote", out, len(cases), "cases")


if __name__ == "__main__":
    main()
