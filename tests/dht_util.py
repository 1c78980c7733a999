"""Parses a dynamic Huffman table as the NX job descriptor carries it (cpb.in_dht: the RFC 1951 §3.2.7
block header from HLIT on, LSB first) into code lengths, and costs a histogram with it."""
CL_ORDER = [16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15]
LEXT = [0] * 8 + [1] * 4 + [2] * 4 + [3] * 4 + [4] * 4 + [5] * 4 + [0]
DEXT = [0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13]


class Bits:
    def __init__(self, data, nbits):
        self.v = int.from_bytes(data, "little")
        self.n, self.p = nbits, 0

    def get(self, k):
        assert self.p + k <= self.n, "table shorter than its own description"
        r = (self.v >> self.p) & ((1 << k) - 1)
        self.p += k
        return r


def canonical(lens):
    """code length list -> {(length, code): symbol}; asserts the code is not over-subscribed"""
    maxl = max(lens) if lens else 0
    count = [0] * (maxl + 2)
    for l in lens:
        count[l] += 1
    count[0] = 0
    code, nxt = 0, [0] * (maxl + 2)
    for l in range(1, maxl + 1):
        code = (code + count[l - 1]) << 1
        nxt[l] = code
    table = {}
    for s, l in enumerate(lens):
        if l:
            table[(l, nxt[l])] = s
            nxt[l] += 1
    return table


def kraft(lens):
    return sum(2.0 ** -l for l in lens if l)


def parse_dht(data, nbits):
    b = Bits(data, nbits)
    hlit, hdist, hclen = b.get(5) + 257, b.get(5) + 1, b.get(4) + 4
    cl = [0] * 19
    for i in range(hclen):
        cl[CL_ORDER[i]] = b.get(3)
    assert kraft(cl) <= 1.0 + 1e-12
    table = canonical(cl)
    lens = []
    while len(lens) < hlit + hdist:
        l, code = 0, 0
        while True:
            code = (code << 1) | b.get(1)
            l += 1
            assert l <= 7, "bad code-length code"
            if (l, code) in table:
                sym = table[(l, code)]
                break
        if sym < 16:
            lens.append(sym)
        elif sym == 16:
            lens += [lens[-1]] * (3 + b.get(2))
        elif sym == 17:
            lens += [0] * (3 + b.get(3))
        else:
            lens += [0] * (11 + b.get(7))
    assert len(lens) == hlit + hdist and b.p == nbits, (len(lens), hlit + hdist, b.p, nbits)
    ll = lens[:hlit] + [0] * (286 - hlit)
    dd = lens[hlit:] + [0] * (30 - hdist)
    return ll, dd


def block_cost(counts, ll, dd, hdr_bits):
    """bits of one dynamic block with this table for this histogram (3 header bits + table + symbols + extra bits)"""
    bits = 3 + hdr_bits
    for s in range(286):
        if counts[s]:
            assert ll[s], f"lit/len symbol {s} has a count but no code"
            bits += counts[s] * (ll[s] + (LEXT[s - 257] if s > 256 else 0))
    for s in range(30):
        if counts[286 + s]:
            assert dd[s], f"distance symbol {s} has a count but no code"
            bits += counts[286 + s] * (dd[s] + DEXT[s])
    return bits
