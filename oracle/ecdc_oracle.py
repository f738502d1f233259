"""CPU restatement of the reference's n-bit packer -- TEST INFRASTRUCTURE ONLY.

binary.BitPacker / BitUnpacker (/root/reference/encodec/binary.py:55-122) as driven by compress_to_file /
decompress_from_file with use_lm=False (compress.py:66-89,130-155): values time-major (for t: for k), ``bits`` each,
LSB first, last byte zero-padded. Pinned to the reference's own BitPacker by tests/golden/ecdc_bitpack.npz
(oracle/make_golden_ecdc.py imports the unmodified reference to produce it).
"""
import numpy as np


def pack_frame(codes: np.ndarray, bits: int) -> bytes:
    """codes [K, T] -> the bytes BitPacker writes (push per value, flush at the end)."""
    cur, nbits, out = 0, 0, bytearray()
    for t in range(codes.shape[1]):
        for k in range(codes.shape[0]):
            cur += int(codes[k, t]) << nbits          # binary.py:71
            nbits += bits
            while nbits >= 8:                          # binary.py:73-77
                out.append(cur & 0xFF)
                nbits -= 8
                cur >>= 8
    if nbits:                                          # flush, binary.py:82-85
        out.append(cur)
    return bytes(out)


def unpack_frame(data: bytes, n_codebooks: int, n_frames: int, bits: int) -> np.ndarray:
    """BitUnpacker.pull (binary.py:104-122), K values per time step -> codes [K, T]."""
    codes = np.zeros((n_codebooks, n_frames), dtype=np.int64)
    cur, nbits, pos, mask = 0, 0, 0, (1 << bits) - 1
    for t in range(n_frames):
        for k in range(n_codebooks):
            while nbits < bits:
                if pos >= len(data):
                    raise EOFError("The stream ended sooner than expected.")
                cur += data[pos] << nbits
                pos += 1
                nbits += 8
            codes[k, t] = cur & mask
            cur >>= bits
            nbits -= bits
    return codes
