"""Stand-in for python-lz4 0.7.0 (`lz4.dumps` / `lz4.loads`, utils/compress_utils.py:289-294), which cannot be installed
here.  PASS-THROUGH: the bytes are framed with the 4-byte little-endian size python-lz4 0.7 writes and not compressed at
all -- bench.py selects `--basic_compressor lz4` to time the reference WITHOUT an entropy coder (its README rates lz4 at
300x the speed of bzip2, so this flatters the reference by a fraction of a millisecond per frame at most)."""
import struct


def dumps(data):
    raw = bytes(memoryview(data).cast("B")) if not isinstance(data, (bytes, bytearray)) else bytes(data)
    return struct.pack("<I", len(raw)) + raw


def loads(data):
    n = struct.unpack_from("<I", data, 0)[0]
    return bytes(data[4:4 + n])


compress, decompress = dumps, loads
