"""The fixed-Huffman code arithmetic of K6 (csrc/deflate.cu token_bits), restated in Python and checked against zlib: a token
stream packed with these formulas must inflate to the bytes the tokens stand for (RFC 1951 3.2.5 / 3.2.6).  The CUDA kernel
uses the same closed forms (no tables); its output is checked against zlib / the oracle / K1 on the GPU in tests/test_gpu_deflate.py."""
import random
import zlib


def brev(x, n):
    return int(format(x, f"0{n}b")[::-1], 2)


def token_bits(lit, length, dist):
    if length == 0:
        c, nb = (0x30 + lit, 8) if lit < 144 else (0x190 + lit - 144, 9)
        return brev(c, nb), nb
    y = length - 3
    if y < 8:
        sym, eb, ev = 257 + y, 0, 0
    elif length == 258:
        sym, eb, ev = 285, 0, 0
    else:
        nb = y.bit_length() - 1
        eb = nb - 2
        sym = 257 + 4 * (nb - 1) + ((y >> eb) & 3)
        ev = y & ((1 << eb) - 1)
    cb = 7 if sym < 280 else 8
    c = sym - 256 if sym < 280 else 0xC0 + sym - 280
    v = brev(c, cb)
    v |= ev << cb
    cb += eb
    x = dist - 1
    if x < 4:
        dsym, deb, dev = x, 0, 0
    else:
        nb = x.bit_length() - 1
        deb = nb - 1
        dsym = 2 * nb + ((x >> deb) & 1)
        dev = x & ((1 << deb) - 1)
    v |= brev(dsym, 5) << cb
    cb += 5
    v |= dev << cb
    cb += deb
    assert cb <= 31
    return v, cb


def pack(tokens):
    acc, n = 3, 3                       # BFINAL = 1, BTYPE = 01
    for t in tokens:
        v, nb = token_bits(*t)
        acc |= v << n
        n += nb
    n += 7                              # end of block: seven zero bits
    return acc.to_bytes((n + 7) // 8, "little")


def test_every_length_and_distance_class():
    out = bytearray(b"x" * 40000)       # history: literals first
    tokens = [(ord("x"), 0, 0)] * 40000
    for length in range(3, 259):
        for dist in (1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145,
                     8193, 12289, 16385, 24577, 32768, 6, 8, 100, 1000, 30000):
            tokens.append((0, length, dist))
            for _ in range(length):
                out.append(out[-dist])
        tokens.append((length & 0xFF, 0, 0))
        out.append(length & 0xFF)
    assert zlib.decompress(pack(tokens), -15) == bytes(out)


def test_random_token_streams():
    rnd = random.Random(11)
    for _ in range(20):
        out, tokens = bytearray(), []
        while len(out) < 20000:
            if out and rnd.random() < 0.5:
                dist = min(len(out), rnd.choice([1, 2, 3, 4, rnd.randrange(1, 300), rnd.randrange(1, 32769)]))
                length = rnd.choice([3, 4, 5, 10, 11, 18, 19, 34, 35, 66, 67, 130, 131, 257, 258, rnd.randrange(3, 259)])
                tokens.append((0, length, dist))
                for _ in range(length):
                    out.append(out[-dist])
            else:
                b = rnd.randrange(256)
                tokens.append((b, 0, 0))
                out.append(b)
        assert zlib.decompress(pack(tokens), -15) == bytes(out)
