"""Generates tests/golden/ecdc_bitpack.npz from the UNMODIFIED reference's binary.BitPacker (run in the build container,
where /root/reference is mounted): random code frames -> the bytes the reference writes for them."""
import io
import os
import sys

import numpy as np

sys.dont_write_bytecode = True
sys.path.insert(0, "/root/reference/encodec")
import binary  # noqa: E402  (the reference's own module)

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = [(8, 75, 10), (32, 150, 10), (2, 113, 10), (16, 45, 10), (3, 37, 7), (5, 19, 13), (1, 9, 1), (4, 64, 8), (7, 33, 16)]


def main():
    out = {}
    rng = np.random.default_rng(20261018)
    for i, (k, t, bits) in enumerate(CASES):
        codes = rng.integers(0, 2 ** bits, size=(k, t), dtype=np.int64)
        fo = io.BytesIO()
        packer = binary.BitPacker(bits, fo)
        for tt in range(t):                      # compress.py:79-88
            for value in codes[:, tt].tolist():
                packer.push(value)
        packer.flush()
        out[f"codes_{i}"] = codes
        out[f"bytes_{i}"] = np.frombuffer(fo.getvalue(), dtype=np.uint8)
        out[f"bits_{i}"] = np.int64(bits)
    out["n_cases"] = np.int64(len(CASES))
    np.savez_compressed(os.path.join(HERE, "..", "tests", "golden", "ecdc_bitpack.npz"), **out)
    print("wrote", len(CASES), "cases")


if __name__ == "__main__":
    main()
