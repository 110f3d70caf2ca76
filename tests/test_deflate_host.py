"""Host half of the device-side deflate (iivision_b200/deflate.py, next row N1): the Huffman
code construction, the dynamic-block header and the CRC-32 algebra, checked against zlib
without a GPU.  The reader of the files is np.load (screen.py:352), i.e. zlib's inflate."""

import heapq
import importlib.util
import os
import zlib

import numpy as np
import pytest


def _load():
    path = os.path.join(os.path.dirname(__file__), "..", "iivision_b200", "deflate.py")
    spec = importlib.util.spec_from_file_location("deflate_under_test", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _huffman_cost(freqs):
    """Total bits of an unconstrained optimal prefix code."""
    h = [f for f in freqs if f > 0]
    if len(h) < 2:
        return sum(h)
    heapq.heapify(h)
    cost = 0
    while len(h) > 1:
        a, b = heapq.heappop(h), heapq.heappop(h)
        cost += a + b
        heapq.heappush(h, a + b)
    return cost


@pytest.mark.parametrize("seed", range(6))
def test_limited_lengths_are_optimal_prefix_codes(seed):
    d = _load()
    rng = np.random.default_rng(seed)
    n = int(rng.integers(2, 286))
    freqs = (rng.pareto(1.2, size=n) * 50).astype(np.int64)
    freqs[rng.random(n) < 0.2] = 0
    if (freqs > 0).sum() < 2:
        freqs[:2] = 1
    lengths = d.limited_lengths(freqs.tolist(), 15)
    used = [l for l, f in zip(lengths, freqs) if f > 0]
    assert all(1 <= l <= 15 for l in used)
    assert all(l == 0 for l, f in zip(lengths, freqs) if f == 0)
    assert sum(2.0 ** -l for l in used) == 1.0                      # complete code
    cost = sum(int(f) * l for f, l in zip(freqs, lengths))
    free = _huffman_cost(freqs.tolist())
    assert cost >= free
    if max(used) < 15:
        assert cost == free       # the limit did not bind: as good as Huffman's own


def test_length_limit_binds_on_fibonacci_frequencies():
    d = _load()
    fib = [1, 1]
    while len(fib) < 30:
        fib.append(fib[-1] + fib[-2])
    lengths = d.limited_lengths(fib, 15)            # Huffman would go 29 deep
    assert max(lengths) == 15 and sum(2.0 ** -l for l in lengths) == 1.0
    assert d.limited_lengths([0, 7, 0], 15) == [0, 1, 0]     # a lone symbol still needs a bit
    assert d.limited_lengths([0, 0], 7) == [0, 0]


def _emit(d, table, tokens):
    """A complete raw-deflate stream of one dynamic block from (literal | (length, distance))
    tokens, coded the way the encode kernel does it: reversed code | length << 16."""
    words = table.words()
    value, nbits = 0, 0

    def put(v, n):
        nonlocal value, nbits
        value |= v << nbits
        nbits += n
    hv = 0
    for k in range(d.CodeTable.HEADER_WORDS):
        hv |= int(words[317 + k]) << (32 * k)
    put(hv & ((1 << int(words[316])) - 1), int(words[316]))
    for tok in tokens:
        if isinstance(tok, int):
            c = int(words[tok])
            put(c & 0xffff, c >> 16)
        else:
            length, dist = tok
            s, x, nb = d.length_symbol(length)
            c = int(words[s])
            put((c & 0xffff) | (x << (c >> 16)), (c >> 16) + nb)
            s, x, nb = d.distance_symbol(dist)
            c = int(words[286 + s])
            put((c & 0xffff) | (x << (c >> 16)), (c >> 16) + nb)
    c = int(words[256])
    put(c & 0xffff, c >> 16)                      # end of block
    put(0, 3)                                     # empty stored block, not final
    pad = (-nbits) % 8
    put(0, pad)
    put(0xffff0000, 32)                           # LEN = 0, NLEN = 0xffff
    return value.to_bytes(nbits // 8, "little") + b"\x01\x00\x00\xff\xff"


def test_code_table_and_header_make_a_stream_zlib_inflates():
    d = _load()
    rng = np.random.default_rng(11)
    raw = bytearray()
    tokens = []
    for _ in range(4000):
        if len(raw) > 8 and rng.random() < 0.3:
            dist = int(rng.choice([2, 8, 96, 4096, int(rng.integers(1, len(raw) + 1))]))
            dist = min(dist, len(raw), 32768)
            length = int(rng.integers(3, 259))
            tokens.append((length, dist))
            for _ in range(length):
                raw.append(raw[-dist])
        else:
            b = int(rng.integers(0, 8)) if rng.random() < 0.5 else int(rng.integers(0, 256))
            tokens.append(b)
            raw.append(b)
    hist = np.zeros(d.HIST_STRIDE, dtype=np.int64)
    for tok in tokens:
        if isinstance(tok, int):
            hist[tok] += 1
        else:
            hist[d.length_symbol(tok[0])[0]] += 1
            hist[d.N_LITLEN + d.distance_symbol(tok[1])[0]] += 1
    table = d.CodeTable(hist)
    assert max(table.litlen_lengths) <= 15 and min(table.litlen_lengths) >= 1   # every symbol coded
    stream = _emit(d, table, tokens)
    assert zlib.decompressobj(-15).decompress(stream) == bytes(raw)
    # the symbol maps agree with RFC 1951's tables at every length and at the distance edges
    for length in range(3, 259):
        s, x, nb = d.length_symbol(length)
        assert d.LENGTH_BASE[s - 257] + x == length and x < (1 << nb) or nb == 0 and x == 0
    for dist in (1, 2, 3, 4, 5, 8, 9, 4096, 4097, 24577, 32768):
        s, x, nb = d.distance_symbol(dist)
        assert d.DIST_BASE[s] + x == dist and (x < (1 << nb) or (nb == 0 and x == 0))


def test_crc32_algebra_matches_zlib():
    d = _load()
    rng = np.random.default_rng(3)
    a, b = rng.bytes(1000), rng.bytes(777)
    assert d.crc32_combine(zlib.crc32(a), zlib.crc32(b), len(b)) == zlib.crc32(a + b)
    assert d.crc32_combine(zlib.crc32(a), zlib.crc32(b""), 0) == zlib.crc32(a)
    part = 256
    for n_parts in (1, 2, 6, 8, 96):
        data = rng.bytes(part * n_parts)
        crcs = np.array([zlib.crc32(data[k * part:(k + 1) * part]) for k in range(n_parts)],
                        dtype=np.uint32)
        assert d.crc32_of_equal_parts(crcs, part) == zlib.crc32(data)
    data = rng.bytes(part * 32)
    crcs = np.array([zlib.crc32(data[k * part:(k + 1) * part]) for k in range(32)], dtype=np.uint32)
    groups = d.crc32_of_groups(crcs, part, 8)
    assert [int(g) for g in groups] == [zlib.crc32(data[k * 8 * part:(k + 1) * 8 * part])
                                        for k in range(4)]
    # the operators the survey kernel's combination tree uses: level k appends 256 << k zeros
    ops = d.level_operators(256, 7)
    x = zlib.crc32(a)
    for k in range(7):
        want = d.crc32_combine(x, 0, 256 << k)          # crc of a + zeros, crc2 = 0 contribution
        got = 0
        for bit in range(32):
            if (x >> bit) & 1:
                got ^= int(ops[k][bit])
        assert got == want
    assert int(d.CRC_TABLE[1]) == 0x77073096
