"""Host half of the device-side deflate of edit-distance tables (SURVEY.md "next" row N1).

The reference stores every table as a compressed ``.npz`` (make_data_tables.py:186-188).
Once the table itself takes a quarter of a millisecond, the deflate is the whole cost of
``make_edit_distance``; the device does it (csrc/iiv_deflate.cu), so that only the
compressed bytes cross PCIe and the host just writes ZIP records.

What the device needs from the host, and what this module computes, is small and
sequential: from the symbol histograms of one table, length-limited Huffman codes for the
literal/length and distance alphabets (RFC 1951 section 3.2), the dynamic-block header that
announces them, and -- afterwards -- the CRC-32 of the whole member from the CRCs of the
device's blocks (zlib's ``crc32_combine`` construction).

The device's blocks are ordinary deflate: a dynamic-Huffman block of literal bytes and of
matches at a few fixed distances (runs of repeated entries, i.e. the zeros above the
diagonal, and repeats of an earlier column), followed by an empty stored block that
byte-aligns it (the pigz construction npz_io.py uses too).  Any inflater reads the result.
"""

import zlib
from typing import List, Sequence, Tuple

import numpy as np

N_LITLEN, N_DIST, N_CLEN = 286, 30, 19
HIST_STRIDE = 320          # uint32 counters per code table: 286 literal/length + 30 distance
MAX_BITS = 15
CLEN_ORDER = (16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15)


def limited_lengths(freqs: Sequence[int], max_len: int) -> List[int]:
    """Optimal prefix-code lengths with no code longer than ``max_len`` (package-merge).
    Symbols with zero frequency get length 0; a single used symbol gets length 1."""
    used = [(int(f), s) for s, f in enumerate(freqs) if f > 0]
    lengths = [0] * len(freqs)
    if not used:
        return lengths
    if len(used) == 1:
        lengths[used[0][1]] = 1
        return lengths
    if len(used) > (1 << max_len):
        raise ValueError("too many symbols for %d-bit codes" % max_len)
    used.sort()
    leaves = [(f, (s,)) for f, s in used]
    packages = list(leaves)
    for _ in range(max_len - 1):
        # pair up, then merge with the leaves again
        paired = [(packages[k][0] + packages[k + 1][0], packages[k][1] + packages[k + 1][1])
                  for k in range(0, len(packages) - 1, 2)]
        packages = sorted(leaves + paired, key=lambda x: x[0])
    for _, syms in packages[:2 * len(used) - 2]:
        for s in syms:
            lengths[s] += 1
    return lengths


def canonical_codes(lengths: Sequence[int]) -> List[int]:
    """RFC 1951 3.2.2 canonical codes, bit-REVERSED (deflate packs Huffman codes starting
    from their most significant bit while everything else goes LSB first)."""
    bl_count = [0] * (MAX_BITS + 2)
    for n in lengths:
        if n:
            bl_count[n] += 1
    code, next_code = 0, [0] * (MAX_BITS + 2)
    for bits in range(1, MAX_BITS + 1):
        code = (code + bl_count[bits - 1]) << 1
        next_code[bits] = code
    out = []
    for n in lengths:
        if n == 0:
            out.append(0)
            continue
        c = next_code[n]
        next_code[n] += 1
        out.append(int(format(c, "0%db" % n)[::-1], 2))
    return out


class _Bits:
    def __init__(self):
        self.value, self.n = 0, 0

    def put(self, v: int, n: int):
        self.value |= (v & ((1 << n) - 1)) << self.n
        self.n += n


def dynamic_header(litlen_lengths: Sequence[int], dist_lengths: Sequence[int],
                   final: bool = False) -> Tuple[int, int]:
    """Bits of a BTYPE=2 block up to (not including) its first symbol: BFINAL, BTYPE, HLIT,
    HDIST, HCLEN, the code-length code and the run-length coded code lengths (RFC 1951
    3.2.7).  Returns (bits as an integer, first bit = bit 0; number of bits)."""
    hlit = max(257, max((s + 1 for s, n in enumerate(litlen_lengths) if n), default=257))
    hdist = max(1, max((s + 1 for s, n in enumerate(dist_lengths) if n), default=1))
    seq = list(litlen_lengths[:hlit]) + list(dist_lengths[:hdist])
    # run-length code the lengths: 16 = repeat previous 3-6, 17 = zeros 3-10, 18 = zeros 11-138
    syms = []       # (symbol, extra value, extra bits)
    k = 0
    while k < len(seq):
        v = seq[k]
        run = 1
        while k + run < len(seq) and seq[k + run] == v:
            run += 1
        if v == 0 and run >= 3:
            take = min(run, 138)
            syms.append((18, take - 11, 7) if take >= 11 else (17, take - 3, 3))
            k += take
        elif v != 0 and run >= 4:
            syms.append((v, 0, 0))
            take = min(run - 1, 6)
            syms.append((16, take - 3, 2))
            k += 1 + take
        else:
            syms.append((v, 0, 0))
            k += 1
    cl_freq = [0] * N_CLEN
    for s, _, _ in syms:
        cl_freq[s] += 1
    cl_len = limited_lengths(cl_freq, 7)
    cl_code = canonical_codes(cl_len)
    hclen = max(4, max((k + 1 for k, s in enumerate(CLEN_ORDER) if cl_len[s]), default=4))
    b = _Bits()
    b.put(1 if final else 0, 1)
    b.put(2, 2)
    b.put(hlit - 257, 5)
    b.put(hdist - 1, 5)
    b.put(hclen - 4, 4)
    for s in CLEN_ORDER[:hclen]:
        b.put(cl_len[s], 3)
    for s, extra, nbits in syms:
        b.put(cl_code[s], cl_len[s])
        if nbits:
            b.put(extra, nbits)
    return b.value, b.n


# length / distance code tables of RFC 1951 3.2.5
LENGTH_BASE = (3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83,
               99, 115, 131, 163, 195, 227, 258)
LENGTH_EXTRA = (0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5,
                5, 0)
DIST_BASE = (1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025,
             1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577)
DIST_EXTRA = (0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11,
              12, 12, 13, 13)


def length_symbol(length: int) -> Tuple[int, int, int]:
    """match length 3..258 -> (literal/length symbol 257..285, extra value, extra bits)."""
    for k in range(len(LENGTH_BASE) - 1, -1, -1):
        if length >= LENGTH_BASE[k]:
            return 257 + k, length - LENGTH_BASE[k], LENGTH_EXTRA[k]
    raise ValueError(length)


def distance_symbol(dist: int) -> Tuple[int, int, int]:
    for k in range(len(DIST_BASE) - 1, -1, -1):
        if dist >= DIST_BASE[k]:
            return k, dist - DIST_BASE[k], DIST_EXTRA[k]
    raise ValueError(dist)


class CodeTable:
    """Everything the encode kernel needs for one table of codes, as one uint32 array:
      [0, 286)      literal/length: reversed code | length << 16
      [286, 316)    distance:       reversed code | length << 16
      [316]         header bit count;  [317 ..]  header bits, 32 per word, LSB first."""

    HEADER_WORDS = 96

    def __init__(self, hist: np.ndarray):
        hist = np.asarray(hist, dtype=np.int64)
        lit = hist[:N_LITLEN].copy()
        dist = hist[N_LITLEN:N_LITLEN + N_DIST].copy()
        lit[256] = max(lit[256], 1)              # every block ends with end-of-block
        # every symbol the encoder could meet must have a code: the histogram comes from the
        # same tokens, but blocks that fall back to another table's codes must stay valid
        lit[:256] = np.maximum(lit[:256], 1)
        lit[257:] = np.maximum(lit[257:], 1)
        dist[:] = np.maximum(dist, 1)
        self.litlen_lengths = limited_lengths(lit.tolist(), MAX_BITS)
        self.dist_lengths = limited_lengths(dist.tolist(), MAX_BITS)
        self.litlen_codes = canonical_codes(self.litlen_lengths)
        self.dist_codes = canonical_codes(self.dist_lengths)
        self.header_bits, self.header_nbits = dynamic_header(self.litlen_lengths,
                                                             self.dist_lengths)
        if self.header_nbits > 32 * self.HEADER_WORDS:
            raise ValueError("dynamic block header of %d bits" % self.header_nbits)

    def words(self) -> np.ndarray:
        out = np.zeros(317 + self.HEADER_WORDS, dtype=np.uint32)
        for s in range(N_LITLEN):
            out[s] = self.litlen_codes[s] | (self.litlen_lengths[s] << 16)
        for s in range(N_DIST):
            out[N_LITLEN + s] = self.dist_codes[s] | (self.dist_lengths[s] << 16)
        out[316] = self.header_nbits
        v = self.header_bits
        for k in range(self.HEADER_WORDS):
            out[317 + k] = v & 0xffffffff
            v >>= 32
        return out


# ---- CRC-32 of a concatenation from the parts' CRCs (zlib's crc32_combine) ----------------

def _gf2_times(mat: List[int], vec: int) -> int:
    out, k = 0, 0
    while vec:
        if vec & 1:
            out ^= mat[k]
        vec >>= 1
        k += 1
    return out


def _gf2_square(mat: List[int]) -> List[int]:
    return [_gf2_times(mat, mat[k]) for k in range(32)]


def zero_operator(n_bytes: int) -> List[int]:
    """32x32 GF(2) matrix (as 32 column words) that maps crc32(A) to the contribution of A
    in crc32(A + n_bytes of anything)."""
    if n_bytes <= 0:
        return [1 << k for k in range(32)]
    odd = [0xedb88320] + [1 << k for k in range(31)]      # one zero BIT
    even = _gf2_square(odd)                                 # two bits
    odd = _gf2_square(even)                                 # four bits
    # now square up to bytes, multiplying in the operators the bits of n_bytes ask for
    result = [1 << k for k in range(32)]
    op = odd
    n = n_bytes
    while True:
        op = _gf2_square(op)            # first pass: 8 bits = one byte
        if n & 1:
            result = [_gf2_times(op, result[k]) for k in range(32)]
        n >>= 1
        if not n:
            break
    return result


def crc32_combine(crc1: int, crc2: int, len2: int) -> int:
    return _gf2_times(zero_operator(len2), crc1) ^ crc2


def crc32_of_equal_parts(crcs: np.ndarray, part_bytes: int) -> int:
    """CRC-32 of the concatenation of len(crcs) parts of ``part_bytes`` bytes each, given
    each part's CRC-32: pairwise combination, vectorised over the parts, while the count is
    even; the rest one by one."""
    crcs = np.asarray(crcs, dtype=np.uint32).copy()
    length = int(part_bytes)
    while len(crcs) > 1 and len(crcs) % 2 == 0:
        op = np.array(zero_operator(length), dtype=np.uint32)
        left, right = crcs[0::2], crcs[1::2]
        shifted = np.zeros_like(left)
        for k in range(32):
            shifted ^= np.where((left >> np.uint32(k)) & np.uint32(1), op[k], np.uint32(0))
        crcs = shifted ^ right
        length *= 2
    total = int(crcs[0])
    if len(crcs) > 1:
        op = zero_operator(length)
        for c in crcs[1:]:
            total = _gf2_times(op, total) ^ int(c)
    return total


def crc32_of_groups(crcs: np.ndarray, part_bytes: int, group: int) -> np.ndarray:
    """CRC-32 of every run of ``group`` (a power of two) consecutive equal-length parts."""
    crcs = np.asarray(crcs, dtype=np.uint32).reshape(-1, group).copy()
    length = int(part_bytes)
    while crcs.shape[1] > 1:
        op = np.array(zero_operator(length), dtype=np.uint32)
        left, right = crcs[:, 0::2], crcs[:, 1::2]
        shifted = np.zeros_like(left)
        for k in range(32):
            shifted ^= np.where((left >> np.uint32(k)) & np.uint32(1), op[k], np.uint32(0))
        crcs = shifted ^ right
        length *= 2
    return crcs[:, 0]


def level_operators(first_bytes: int, levels: int) -> np.ndarray:
    """Operators for a binary combination tree whose leaves are ``first_bytes`` long:
    uint32[levels][32]; level k joins two neighbours of first_bytes << k bytes."""
    return np.array([zero_operator(first_bytes << k) for k in range(levels)], dtype=np.uint32)


def _crc_table() -> np.ndarray:
    """The byte-at-a-time table of CRC-32 (reflected 0xEDB88320)."""
    t = np.zeros(256, dtype=np.uint32)
    for n in range(256):
        c = n
        for _ in range(8):
            c = (c >> 1) ^ 0xedb88320 if c & 1 else c >> 1
        t[n] = c
    return t


CRC_TABLE = _crc_table()
