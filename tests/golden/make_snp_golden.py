"""Writes tests/golden/snp_unphased_tiny.snpdat.hex: a `.snpdat` file derived BY HAND from the layout the reference documents
(adelie_core/io/io_snp_unphased.ipp:88-110 header, :164-262 columns) -- every byte below is spelled out, no writer is called.

calldata (n=5, p=2), column-major:   col 0 = [1, 0, -9, 2, 1]     col 1 = [2, -9, -9, 0, 2]
  nnz = [4, 4]   nnm = [4, 3]   impute = [(1+2+1)/4, (2+2)/3] = [1.0, 4/3]
  preamble = 1 + 8 + 8 + 3*8*2 + 8*3 = 89 bytes;  column 0 = 55 bytes, column 1 = 50 bytes;  outer = [89, 144, 194]
"""
import os
import struct

u64 = lambda *x: struct.pack("<%dQ" % len(x), *x)
u32 = lambda x: struct.pack("<I", x)
u8 = lambda *x: bytes(x)

header = (
    u8(0)                       # little endian
    + u64(5) + u64(2)           # n, p
    + u64(4, 4)                 # nnz
    + u64(4, 3)                 # nnm
    + struct.pack("<2d", 1.0, 4.0 / 3.0)   # impute
    + u64(89, 144, 194)         # outer
)
col0 = (
    u64(24, 34, 45)             # category offsets relative to the column start
    + u32(1) + u32(0) + u8(0) + u8(2)            # missing: 1 chunk; chunk 0, 1 entry (stored 0), row 2
    + u32(1) + u32(0) + u8(1) + u8(0, 4)         # ones: chunk 0, 2 entries, rows 0 and 4
    + u32(1) + u32(0) + u8(0) + u8(3)            # twos: chunk 0, 1 entry, row 3
)
col1 = (
    u64(24, 35, 39)
    + u32(1) + u32(0) + u8(1) + u8(1, 2)         # missing: rows 1, 2
    + u32(0)                                      # ones: no chunk
    + u32(1) + u32(0) + u8(1) + u8(0, 4)         # twos: rows 0, 4
)
blob = header + col0 + col1
assert len(header) == 89 and len(col0) == 55 and len(col1) == 50 and len(blob) == 194
here = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(here, "snp_unphased_tiny.snpdat.hex"), "w") as f:
    f.write(blob.hex() + "\n")
print("wrote", len(blob), "bytes")
