"""The host range coder and the bitstream container (csrc/codec.cu, sga_b200/codec.py) on hand-made tables: no GPU needed.
Reference seam: mbt2018.py:84-85 (`string`, `side_string`), :211-221 (PackedTensors), :248-295 (decompress)."""
import numpy as np
import pytest


def _random_table(rng, L):
    """L regular symbols + escape, every frequency >= 1, sum 65536."""
    w = rng.random(L + 1) ** 3 + 1e-4
    w[-1] *= 0.02
    f = np.floor(w / w.sum() * (65536 - (L + 1))).astype(np.int64) + 1
    f[np.argmax(f)] += 65536 - f.sum()
    assert f.sum() == 65536 and f.min() >= 1
    return f.astype(np.uint32)


def _codec(rng):
    from sga_b200.codec import HostCodec
    ty = [(-2, _random_table(rng, 5)), (-150, _random_table(rng, 301)), (0, _random_table(rng, 1)), (-700, _random_table(rng, 1401))]
    tz = [(-96, _random_table(rng, 193)), (-96, _random_table(rng, 193))]
    return HostCodec.from_tables(ty, tz), ty, tz


def test_range_coder_round_trip_with_escapes_and_size_near_entropy():
    rng = np.random.default_rng(0)
    cd, ty, tz = _codec(rng)
    n = 200000
    tab = rng.integers(0, len(ty), n).astype(np.int32)
    sym = np.empty(n, dtype=np.int32)
    bits = 0.0
    for t, (mn, f) in enumerate(ty):
        idx = np.nonzero(tab == t)[0]
        L = len(f) - 1
        p = f[:L].astype(np.float64) / f[:L].sum()
        q = rng.choice(L, size=idx.size, p=p)
        sym[idx] = q + mn
        bits += float(-np.log2(f[q] / 65536.0).sum())
    out_of_range = rng.choice(n, 300, replace=False)          # escapes: beyond both ends, some very far
    for i in out_of_range:
        mn, f = ty[tab[i]]
        L = len(f) - 1
        far = int(rng.choice([1, 2, 7, 100, 12345, 2 ** 20]))
        sym[i] = mn - far if rng.random() < 0.5 else mn + L - 1 + far
    data = cd.encode(0, tab, sym)
    back = cd.decode(0, tab, data)
    assert np.array_equal(back, sym)
    assert len(data) * 8 <= bits * 1.002 + 300 * 64 + 64          # within 0.2 % of the tables' entropy (+ the escapes)
    assert len(data) * 8 >= bits * 0.98
    # z tables, short and empty inputs
    tabz = rng.integers(0, 2, 17).astype(np.int32)
    symz = rng.integers(-120, 121, 17).astype(np.int32)
    assert np.array_equal(cd.decode(1, tabz, cd.encode(1, tabz, symz)), symz)
    empty = np.zeros(0, dtype=np.int32)
    assert cd.decode(0, empty, cd.encode(0, empty, empty)).size == 0
    one = cd.encode(0, np.array([2], np.int32), np.array([0], np.int32))       # a 1-symbol table costs ~nothing
    assert np.array_equal(cd.decode(0, np.array([2], np.int32), one), [0])
    cd.close()


def test_range_coder_rejects_bad_tables_and_indices():
    from sga_b200 import _lib
    from sga_b200.codec import HostCodec
    bad = np.full(4, 16384, dtype=np.uint32)
    bad[0] = 0
    with pytest.raises(_lib.SgaError):
        HostCodec.from_tables([(0, bad)], [])
    with pytest.raises(_lib.SgaError):
        HostCodec.from_tables([(0, np.full(4, 100, dtype=np.uint32))], [])     # does not sum to 65536
    rng = np.random.default_rng(1)
    cd, ty, tz = _codec(rng)
    with pytest.raises(_lib.SgaError):
        cd.encode(0, np.array([99], np.int32), np.array([0], np.int32))
    cd.close()


def test_container_round_trip_and_corruption():
    from sga_b200 import codec
    streams = [(b"zz0", b"yyyy0"), (b"", b"y1"), (b"z2", b"")]
    blob = codec.pack(streams, codec.MODE_MEAN_CENTRED, 192, 768, 512)
    s2, mode, C, H, W = codec.unpack(blob)
    assert s2 == streams and (mode, C, H, W) == (1, 192, 768, 512)
    with pytest.raises(ValueError):
        codec.unpack(b"XXXX" + blob[4:])
    with pytest.raises(ValueError):
        codec.unpack(blob[:-2])
