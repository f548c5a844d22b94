"""Writes tests/golden/snp_phased_tiny.snpdat.hex: a phased-ancestry `.snpdat` file derived BY HAND from the layout documented in
adelie_core/io/io_snp_phased_ancestry.ipp (header :170-186, SNP blocks :268-342) -- every byte is spelled out, no writer is called.

n = 3, s = 1, A = 2.   calldata (n, 2) = [[1, 0], [1, 1], [0, 1]],  ancestries (n, 2) = [[0, 1], [1, 1], [0, 0]]
  hap 0 (column 0): rows 0 (ancestry 0), 1 (ancestry 1) carry the mutation;  hap 1 (column 1): rows 1 (ancestry 1), 2 (ancestry 0)
  matrix (n, s*A) = [[1, 0], [0, 2], [1, 0]];  nnz0 = [1, 1], nnz1 = [1, 1]
  preamble = 1 + 8 + 8 + 1 + 16*2 + 8*2 = 66;  SNP block = 2*8 + 2 * (16 + 10 + 10) = 88;  outer = [66, 154]
"""
import os
import struct

u64 = lambda *x: struct.pack("<%dQ" % len(x), *x)
u32 = lambda x: struct.pack("<I", x)
u8 = lambda *x: bytes(x)

header = u8(0) + u64(3) + u64(1) + u8(2) + u64(1, 1) + u64(1, 1) + u64(66, 154)
hap = lambda row: u32(1) + u32(0) + u8(0) + u8(row)          # one chunk (index 0), one entry, the row
anc0 = u64(16, 26) + hap(0) + hap(2)                          # ancestry 0: hap 0 hits row 0, hap 1 hits row 2
anc1 = u64(16, 26) + hap(1) + hap(1)                          # ancestry 1: hap 0 hits row 1, hap 1 hits row 1
snp = u64(16, 16 + 36) + anc0 + anc1
blob = header + snp
assert len(header) == 66 and len(anc0) == 36 and len(snp) == 88 and len(blob) == 154
here = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(here, "snp_phased_tiny.snpdat.hex"), "w") as f:
    f.write(blob.hex() + "\n")
print("wrote", len(blob), "bytes")
